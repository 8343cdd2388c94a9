// Microbenchmark: issue throughput of scalar FFMA vs packed FFMA2/FADD2 (sm_100a), and of a
// complex butterfly written both ways.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 16
__global__ void k_ffma(float* out, float a, float b, int iters)
{
  float x[2 * CHAINS];
#pragma unroll
  for (int i = 0; i < 2 * CHAINS; i++) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; i++) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 2 * CHAINS; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b, int iters)
{
  float2 x[CHAINS];
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < CHAINS; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x[i] = __ffma2_rn(x[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// complex butterfly lo = e + w*o, hi = e - w*o on CHAINS/2 pairs: scalar (6 FFMA) vs packed
__global__ void k_bfly(float* out, float wr, float wi, int iters)
{
  float2 x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i += 2) {
      float2 e = x[i], o = x[i + 1];
      const float lx = fmaf(-o.y, wi, fmaf(o.x, wr, e.x));
      const float ly = fmaf(o.x, wi, fmaf(o.y, wr, e.y));
      x[i + 1] = make_float2(fmaf(2.0f, e.x, -lx), fmaf(2.0f, e.y, -ly));
      x[i] = make_float2(lx, ly);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bfly2(float* out, float wr, float wi, int iters)
{
  float2 x[CHAINS];
  const float2 wrr = make_float2(wr, wr), wii = make_float2(-wi, wi), two = make_float2(2.f, 2.f);
#pragma unroll
  for (int i = 0; i < CHAINS; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i += 2) {
      float2 e = x[i], o = x[i + 1];
      float2 lo = __ffma2_rn(o, wrr, e);                               // e + o*wr
      lo = __ffma2_rn(make_float2(o.y, o.x), wii, lo);                 // + (-o.y*wi, o.x*wi)
      x[i + 1] = __ffma2_rn(two, e, make_float2(-lo.x, -lo.y));        // 2e - lo
      x[i] = lo;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K>
static float run(K kern, float* d, int iters, float p0, float p1)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  kern<<<148 * 4, 256>>>(d, p0, p1, iters);
  cudaEventRecord(a);
  kern<<<148 * 4, 256>>>(d, p0, p1, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms;
}
int main()
{
  float* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  const int iters = 20000;
  const double threads = 148.0 * 4 * 256;
  float t1 = run(k_ffma, d, iters, 0.999f, 0.001f);
  float t2 = run(k_ffma2, d, iters, 0.999f, 0.001f);
  float t3 = run(k_bfly, d, iters, 0.6f, 0.8f);
  float t4 = run(k_bfly2, d, iters, 0.6f, 0.8f);
  printf("FFMA  : %.3f ms  %.1f Gfma/s (scalar lanes)\n", t1, threads * iters * 2 * CHAINS / t1 / 1e6);
  printf("FFMA2 : %.3f ms  %.1f Gfma/s (scalar lanes)\n", t2, threads * iters * 2 * CHAINS / t2 / 1e6);
  printf("bfly scalar : %.3f ms  %.1f Gbfly/s\n", t3, threads * iters * (CHAINS / 2) / t3 / 1e6);
  printf("bfly packed : %.3f ms  %.1f Gbfly/s\n", t4, threads * iters * (CHAINS / 2) / t4 / 1e6);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
