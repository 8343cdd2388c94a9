// kernels_dispatch.cu -- format dispatch for the per-format translation units
#include <cuda_runtime.h>
#include "fft1_small.cuh"
#include "plan.h"
using namespace lb;
typedef cudaError_t (*fft1_small_launch_t)(const Fft1K&, int grid, cudaStream_t);
fft1_small_launch_t lb_get_fft1_small_fmt0(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt1(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt2(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt3(int, int, int*, size_t*);

fft1_small_launch_t lb_get_fft1_small(int log2n, int fmt, int variant, int* threads, size_t* smem)
{
  switch (fmt) {
    case 0: return lb_get_fft1_small_fmt0(log2n, variant, threads, smem);
    case 1: return lb_get_fft1_small_fmt1(log2n, variant, threads, smem);
    case 2: return lb_get_fft1_small_fmt2(log2n, variant, threads, smem);
    case 3: return lb_get_fft1_small_fmt3(log2n, variant, threads, smem);
  }
  return nullptr;
}

bool lb_fft1_large_supported(int) { return false; }
cudaError_t lb_launch_fft1_large(lb200_plan*, const Fft1K&) { return cudaErrorNotSupported; }
