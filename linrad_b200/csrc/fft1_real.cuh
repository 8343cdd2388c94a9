// fft1_real.cuh -- second half of the real-input fft1 (fft1 version 2, fft1_re.c:32-232).
//
// The reference scatters 2N windowed real samples through fft1_permute and runs
// fft_real_to_hermitian (fft0.c:33-159), giving X_k = sum_n x[n] w[n] exp(-2 pi i n k / 2N),
// k = 0..N, then maps it to fft1_float (fft1_re.c:100-130):
//   fft1_direction > 0 : bin k = (Im X_k, Re X_k), k = max(first,1)..last;  bin 0 = (X_N, X_0)
//   fft1_direction < 0 : bin N-ia = (Re X_ia, Im X_ia), ia = max(N-1-last,1) .. N-first, where the
//                        "imaginary part" read for ia = N is tmp[N] = X_N;  bin N-1 = (X_0, X_N)
//                        is written first and overwritten by ia = 1 when the loop reaches it
// Here the transform kernels (fft1_small_kernel / four-step) have left the N-point complex
// spectrum Z of the packed sequence z[m] = x[2m] w[2m] + i x[2m+1] w[2m+1] in zbuf; this kernel
// untangles   X_k = (Z_k + conj Z_{N-k})/2 - (i/2) exp(-i pi k/N) (Z_k - conj Z_{N-k}),
// applies the mapping above, then fft1_c's arithmetic (fft1.c:4115-4200: filtercorr multiply,
// |z|^2 summed over channels and over the averaging group into fft1_sumsq).
// One thread owns one bin for all transforms of one averaging group, so the fft1_sumsq row is
// accumulated in a register in the reference's order and written once.
#pragma once
#include "fft1_small.cuh"

namespace lb {

LB_D float2 real_untangle(const float2* __restrict__ Z, const float2* __restrict__ Wre, int k, int N)
{
  // L2 loads: inside the persistent kernel Z is a ring whose slots other SMs rewrite (no stale L1 lines)
  const float2 a = __ldcg(Z + (k & (N - 1)));      // Z_N == Z_0
  const float2 zb = __ldcg(Z + ((N - k) & (N - 1)));
  const float2 b = make_float2(zb.x, -zb.y);
  const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
  const float2 o = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y - b.y));
  const float2 w = Wre[k];                         // (cos, -sin)(pi k/N)
  const float c = w.x, s = -w.y;
  // X = E + (-s - i c) * O
  return make_float2(e.x - s * o.x + c * o.y, e.y - c * o.x - s * o.y);
}

// which X output bin j shows (fft1_re.c:100-130): spectrum index k >= 0, -1 = the special bin (0 / N-1), -2 = a bin
// fft1_b does not write (the reference leaves the ring's old contents, and fft1_c never touches such a bin either)
LB_D int real_bin_source(const Fft1K& p, int j, int N)
{
  if (p.direction > 0) {
    const int lo = p.first_point < 1 ? 1 : p.first_point;
    return (j >= lo && j <= p.last_point) ? j : (j == 0 ? -1 : -2);
  }
  int kk = N - 1 - p.last_point;
  const int m = 1 + kk + p.last_point - p.first_point;
  if (kk == 0) kk = 1;
  const int ia = N - j;
  return (ia >= kk && ia <= m) ? ia : (j == N - 1 ? -1 : -2);
}

// one channel's value of output bin j (k = real_bin_source(j) >= -1) after the mapping and fft1_c's filter
// correction; adds |z|^2 to pw when the bin is inside [first_point, last_point]
LB_D float2 real_bin_value(const Fft1K& p, const float2* __restrict__ Z, int j, int k, int N, int c, int MM, bool inr, float& pw)
{
  float2 o;
  if (k >= 0) {
    const float2 x = real_untangle(Z, p.Wre, k, N);
    if (p.direction > 0) o = make_float2(x.y, x.x);
    else o = (k == N) ? make_float2(x.x, x.x) : x;
  } else {
    const float2 z0 = __ldcg(Z);
    const float x0 = z0.x + z0.y, xn = z0.x - z0.y;
    o = p.direction > 0 ? make_float2(xn, x0) : make_float2(x0, xn);
  }
  if (p.fc_mode != 0 && inr) {
    float2 f;
    if (p.fc_mode == 2 || j < p.fc_edge || j >= N - p.fc_edge)
      f = *reinterpret_cast<const float2*>(p.filtercorr + (size_t)j * MM + 2 * c);
    else
      f = make_float2(p.fc_gain, 0.0f);
    const float re = o.x * f.x - o.y * f.y;       // fft1.c:4121-4125
    const float im = o.y * f.x + o.x * f.y;
    o = make_float2(re, im);
    pw += re * re + im * im;
  }
  return o;
}

template <int NCH>
__global__ void __launch_bounds__(256)
fft1_real_post_kernel(const Fft1K p, int log2n, int b_first, int b_count, int g_first, int g_count)
{
  constexpr int MM = 2 * NCH;
  const int N = 1 << log2n;
  const int group_size = p.power_rows ? 1 : p.avg1num;
  const int c0 = p.power_rows ? 0 : p.counter0;
  const int chunks = (N + blockDim.x - 1) / blockDim.x;
  const int nwork = g_count * chunks;
  for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int g = g_first + w / chunks;
    const int j = (w % chunks) * blockDim.x + threadIdx.x;      // output bin
    if (j >= N) continue;
    int b0 = g * group_size - c0;
    int b1 = b0 + group_size;
    if (b0 < 0) b0 = 0;
    if (b1 > p.nblocks) b1 = p.nblocks;
    const int k = real_bin_source(p, j, N);
    const bool inr = (j >= p.first_point) && (j <= p.last_point);
    float acc = 0.0f;
    for (int b = b0; b < b1; b++) {
      float* outb = p.out + ((p.out_pa + (uint32_t)b * (uint32_t)(MM * N)) & p.out_mask);
      float pw = 0.0f;
      if (k >= -1) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          const float2* Z = p.zbuf + ((size_t)(b - p.zb_first) * NCH + c) * N;
          const float2 o = real_bin_value(p, Z, j, k, N, c, MM, inr, pw);
          *reinterpret_cast<float2*>(outb + (size_t)j * MM + 2 * c) = o;
        }
      }
      if (p.fc_mode != 0) {
        if (p.power_rows) p.power_rows[(size_t)b * N + j] = inr ? pw : 0.0f;
        acc = (b == b0) ? pw : acc + pw;
      }
    }
    if (p.sumsq && !p.power_rows && p.fc_mode != 0 && b1 > b0 && inr) {
      float* row = p.sumsq + ((p.sumsq_pa + (uint32_t)g * (uint32_t)N) & p.sumsq_mask);
      const bool continuing = (g == 0 && p.counter0 > 0);
      row[j] = continuing ? row[j] + acc : acc;
    }
  }
}

}  // namespace lb
