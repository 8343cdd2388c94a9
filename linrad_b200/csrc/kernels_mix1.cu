// kernels_mix1.cu -- explicit launchers for mix1_kernel
#include <cuda_runtime.h>
#include "mix1.cuh"
using namespace lb;

template <int LOG2M, int LOG2E, int NCH>
static cudaError_t launch_mix1(const Mix1K& k, int grid, cudaStream_t s)
{
  constexpr size_t smem = mix1_smem<LOG2M, NCH>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mix1_kernel<LOG2M, LOG2E, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  mix1_kernel<LOG2M, LOG2E, NCH><<<grid, NCH << (LOG2M - LOG2E), smem, s>>>(k);
  return cudaGetLastError();
}

typedef cudaError_t (*mix1_launch_t)(const Mix1K&, int grid, cudaStream_t);

#define LB_MCASE(LM, LE)                                                       \
  if (log2m == LM) {                                                           \
    *threads = nch << (LM - LE);                                               \
    if (nch == 1) { *smem = mix1_smem<LM, 1>(); return launch_mix1<LM, LE, 1>; } \
    *smem = mix1_smem<LM, 2>(); return launch_mix1<LM, LE, 2>;                 \
  }

mix1_launch_t lb_get_mix1(int log2m, int nch, int* threads, size_t* smem)
{
  LB_MCASE(3, 3) LB_MCASE(4, 3) LB_MCASE(5, 3) LB_MCASE(6, 3) LB_MCASE(7, 3) LB_MCASE(8, 3) LB_MCASE(9, 3)
  LB_MCASE(10, 3) LB_MCASE(11, 4) LB_MCASE(12, 4)
  return nullptr;
}
