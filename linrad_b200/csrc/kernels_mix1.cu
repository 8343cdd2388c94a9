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

template <int LOG2M, int LOG2E, int NCH>
static cudaError_t launch_mix1(const Mix1K& k, int grid, cudaStream_t s)
{
  constexpr int PAR = Mix1Par<LOG2M, LOG2E, NCH>::value;
  constexpr size_t smem = mix1_smem<LOG2M, NCH, PAR>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mix1_kernel<LOG2M, LOG2E, NCH, PAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  mix1_kernel<LOG2M, LOG2E, NCH, PAR><<<grid, (PAR * NCH) << (LOG2M - LOG2E), smem, s>>>(k);
  return cudaGetLastError();
}

typedef cudaError_t (*mix1_launch_t)(const Mix1K&, int grid, cudaStream_t);

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

// mix1.size 8 .. 8192 (one channel) / 8 .. 4096 (two channels): what fits the 227 KB of shared
// memory with the predecessor's tail kept on chip.  (The reference allows up to 32768, buf.c:856.)
mix1_launch_t lb_get_mix1(int log2m, int nch, int* threads, size_t* smem, int* par)
{
  LB_MCASE(3, 3) LB_MCASE(4, 3) LB_MCASE(5, 3) LB_MCASE(6, 3) LB_MCASE(7, 3) LB_MCASE(8, 3) LB_MCASE(9, LB_MIX1_LE_MID)
  LB_MCASE(10, LB_MIX1_LE_MID) LB_MCASE(11, LB_MIX1_LE_BIG) LB_MCASE(12, LB_MIX1_LE_BIG)
  if (log2m == 13 && nch == 1) LB_MCASE1(13, 4, 1)
  return nullptr;
}
