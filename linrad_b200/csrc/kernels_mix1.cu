// kernels_mix1.cu -- explicit launchers for mix1_kernel
#include <cuda_runtime.h>
#include "mix1.cuh"
using namespace lb;

// points per thread (log2) of the back-transform for mix1.size 512 and 1024
#ifndef LB_MIX1_LE_MID
#define LB_MIX1_LE_MID 3
#endif
// ... and for mix1.size 2048 .. 8192
#ifndef LB_MIX1_LE_BIG
#define LB_MIX1_LE_BIG 4
#endif

// transforms side by side in one CTA, at most 8: about 512 threads for small mix1.size, 256 from
// 1024 points up (measured: two half-size CTAs per SM overlap each other's barrier phases better
// than one -- cfg4 mix1 75 -> 60 us, cfg2 84 -> 76 us -- while mix1.size 512 loses with 256)
template <int LOG2M, int LOG2E, int NCH>
struct Mix1Par {
  static constexpr int LANE = NCH << (LOG2M - LOG2E);
  static constexpr int RAW = (LOG2M >= 10 ? 256 : 512) / LANE;
  static constexpr int value = RAW < 1 ? 1 : (RAW > 8 ? 8 : RAW);
};

template <int LOG2M, int LOG2E, int NCH, int STAGE>
static cudaError_t launch_mix1_stage(const Mix1K& k, int grid, cudaStream_t s)
{
  constexpr int PAR = Mix1Par<LOG2M, LOG2E, NCH>::value;
  constexpr size_t smem = STAGE == 0 ? mix1_smem<LOG2M, NCH, PAR>() : mix1_smem_split<LOG2M, NCH, PAR>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mix1_kernel<LOG2M, LOG2E, NCH, PAR, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  mix1_kernel<LOG2M, LOG2E, NCH, PAR, STAGE><<<grid, (PAR * NCH) << (LOG2M - LOG2E), smem, s>>>(k);
  return cudaGetLastError();
}

// stage 0: one kernel (sizes whose predecessor tail fits on chip)
template <int LOG2M, int LOG2E, int NCH>
static cudaError_t launch_mix1(const Mix1K& k, int grid, int stage, cudaStream_t s)
{
  if (stage != 0) return cudaErrorNotSupported;
  return launch_mix1_stage<LOG2M, LOG2E, NCH, 0>(k, grid, s);
}
// stages 1 and 2: the two-launch form of the sizes that do not fit (k.ybuf_g between them)
template <int LOG2M, int LOG2E, int NCH>
static cudaError_t launch_mix1_split(const Mix1K& k, int grid, int stage, cudaStream_t s)
{
  if (stage == 1) return launch_mix1_stage<LOG2M, LOG2E, NCH, 1>(k, grid, s);
  if (stage == 2) return launch_mix1_stage<LOG2M, LOG2E, NCH, 2>(k, grid, s);
  return cudaErrorNotSupported;
}

typedef cudaError_t (*mix1_launch_t)(const Mix1K&, int grid, int stage, cudaStream_t);

#define LB_MCASE1(LM, LE, NC)                                                  \
  {                                                                            \
    *par = Mix1Par<LM, LE, NC>::value;                                         \
    *threads = (*par * NC) << (LM - LE);                                       \
    *smem = mix1_smem<LM, NC, Mix1Par<LM, LE, NC>::value>();                   \
    return launch_mix1<LM, LE, NC>;                                            \
  }
#define LB_MCASE(LM, LE)                                                       \
  if (log2m == LM) {                                                           \
    if (nch == 1) LB_MCASE1(LM, LE, 1)                                         \
    LB_MCASE1(LM, LE, 2)                                                       \
  }

// mix1.size 8 .. 8192 (one channel) / 8 .. 4096 (two channels): what fits the 227 KB of shared memory with the
// predecessor's tail kept on chip (*split = 0).  16384 / 8192: the two-launch form (*split = 1).  (The reference
// allows up to 32768, buf.c:856.)
mix1_launch_t lb_get_mix1(int log2m, int nch, int* threads, size_t* smem, int* par, int* split)
{
  *split = 0;
  LB_MCASE(3, 3) LB_MCASE(4, 3) LB_MCASE(5, 3) LB_MCASE(6, 3) LB_MCASE(7, 3) LB_MCASE(8, 3) LB_MCASE(9, LB_MIX1_LE_MID)
  LB_MCASE(10, LB_MIX1_LE_MID) LB_MCASE(11, LB_MIX1_LE_BIG) LB_MCASE(12, LB_MIX1_LE_BIG)
  if (log2m == 13 && nch == 1) LB_MCASE1(13, 4, 1)
  *split = 1;
  if (log2m == 14 && nch == 1) {
    *par = 1; *threads = 512; *smem = mix1_smem_split<14, 1, 1>();
    return launch_mix1_split<14, 5, 1>;
  }
  if (log2m == 13 && nch == 2) {
    *par = 1; *threads = 512; *smem = mix1_smem_split<13, 2, 1>();
    return launch_mix1_split<13, 5, 2>;
  }
  return nullptr;
}
