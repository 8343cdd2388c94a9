// timf2.cuh -- the front end of the second FFT: make_timf2 (timf2.c:31-208) for the float path
// (swfloat): split every fft1 transform into a weak and a strong spectrum by liminfo, transform both
// back to the time domain (fft1back_one / fft1back_two, timf2.c:210-968: a plain forward DFT of the
// split spectrum, probed from the compiled reference) and lay the result into the timf2 ring
// (fft1back_fp_finish, timf2.c:970-1127): straight copy without a window, overlap-add of the two
// halves with the sin^2 window, centre portion times the inverted window otherwise; |weak|^2 into
// timf2_pwr_float.
//
// Two kernels per call: timf2_back_kernel (one CTA per transform and signal: gather + split + DFT
// into an L2-resident scratch) and timf2_finish_kernel (one thread per output sample and signal
// set: the sequential += of the reference becomes "this transform's first half + the previous
// transform's second half", the very first one taking the half the previous call parked in the ring).
#pragma once
#include "fft32_core.cuh"

namespace lb {

struct Timf2K {
  const float* fft1;       // fft1_float ring
  uint32_t fft1_mask;      // floats
  uint32_t fft1_px;        // first transform
  int nblocks;
  const float* liminfo;    // fft1_size floats (device)
  int first_point, last_point;
  float2* tmp;             // [nblocks][2*NCH][N]: timf2_tmp of every transform
  const float2* Wn;        // exp(-2 pi i m / N)
  const float4* tab1;      // pass-1 twiddles of the 32-points-per-thread plan (N > 1024)
  // finish
  float* timf2;            // timf2_float ring
  uint32_t timf2_mask;     // floats
  uint32_t timf2_pa;
  float* pwr;              // timf2_pwr_float: one float per ring sample
  float ampfac;            // 1 / (1 << genparm[FIRST_BCKFFT_ATT_N])
  const float* invwin;     // fft1_inverted_window (mode 2)
  int interleave;          // fft1_interleave_points
  int mode;                // 0 no window, 1 sin^2 (interleave == N/2), 2 inverted window
};

// ---- split + back transform -------------------------------------------------------------------
template <int LOG2N, int NCH>
__global__ void __launch_bounds__(LOG2N >= 10 ? (1 << (LOG2N - 5)) : (1 << (LOG2N - 3)))
timf2_back_kernel(const Timf2K p)
{
  constexpr int N = 1 << LOG2N, MM = 2 * NCH, S = 2 * NCH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* xch = reinterpret_cast<float2*>(smem_raw);
  const int t = threadIdx.x;
  const int sig = blockIdx.x % S;          // strong * NCH + channel
  const int b = blockIdx.x / S;
  const int c = sig % NCH;
  const bool strong = sig >= NCH;
  const float* x = p.fft1 + ((p.fft1_px + (uint32_t)b * (uint32_t)(MM * N)) & p.fft1_mask);
  float2* out = p.tmp + ((size_t)b * S + sig) * N;
  auto pick = [&](int k) {
    float2 z = make_float2(0.f, 0.f);
    if (k >= p.first_point && k <= p.last_point) {
      const bool is_weak = p.liminfo[k] == 0.0f;          // timf2.c:46, 86
      if (is_weak != strong) z = *reinterpret_cast<const float2*>(x + (size_t)k * MM + 2 * c);
    }
    return z;
  };
  if constexpr (LOG2N >= 10) {
    using P = Plan32<LOG2N>;
    constexpr int T = P::T;
    float4* tab1 = reinterpret_cast<float4*>(smem_raw + sizeof(float2) * P::XCH);
    if (P::NPASS == 3)
      for (int i = t; i < P::TAB1; i += T) tab1[i] = p.tab1[i];
    float2 wb[5];
#pragma unroll
    for (int j = 0; j < 5; j++) wb[j] = p.Wn[t << j];
    float2 v[32];
#pragma unroll
    for (int e = 0; e < 32; e++) v[e] = pick(t + T * e);
    pass0<P::R0>(v);
    exch1_store<LOG2N>(v, xch, t);
    __syncthreads();
    exch1_load<LOG2N>(v, xch, t);
    if (P::NPASS == 3) {
      float2 w32[32];
      const float4* tp = tab1 + (t & (P::R0 - 1));
#pragma unroll
      for (int q = 0; q < 16; q++) {
        const float4 f = tp[q * P::R0];
        w32[2 * q] = make_float2(f.x, f.y);
        w32[2 * q + 1] = make_float2(f.z, f.w);
      }
      radix32_table(v, w32);
      __syncthreads();
      exch2_store<LOG2N>(v, xch, t);
      __syncthreads();
      exch2_load<LOG2N>(v, xch, t);
    }
    radix32_gen(v, wb);
#pragma unroll
    for (int e = 0; e < 32; e++) out[t + T * e] = v[e];
  } else {
    using P = Plan<LOG2N, 3>;
    constexpr int E = P::E, T = P::T;
    Twiddles<P> tw;
    load_twiddles<P>(tw, p.Wn, t);
    float2 v[E];
#pragma unroll
    for (int e = 0; e < E; e++) v[e] = pick(t + T * e);
    fft_forward<P>(v, xch, t, tw);
#pragma unroll
    for (int e = 0; e < E; e++) out[t + T * e] = v[e];
  }
}

// ---- fft1back_fp_finish -----------------------------------------------------------------------
// One thread per ring sample written by this call: nblocks * new_points samples, plus (sin^2 window)
// the N/2 samples of the last transform's second half that are parked for the next call.
template <int NCH>
__global__ void __launch_bounds__(256) timf2_finish_kernel(const Timf2K p, int log2n)
{
  constexpr int S = 2 * NCH, SF = 4 * NCH;           // signals, floats per sample
  const int N = 1 << log2n;
  const int newp = N - p.interleave;
  const long total = (long)p.nblocks * newp + (p.mode == 1 ? N / 2 : 0);
  for (long w = (long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long)gridDim.x * blockDim.x) {
    const int b = (int)(w / newp);
    const int i = (int)(w - (long)b * newp);
    const uint32_t pos = (p.timf2_pa + (uint32_t)w * (uint32_t)SF) & p.timf2_mask;   // consecutive samples, consecutive ring slots
    float* dst = p.timf2 + pos;
    float o[SF];
    if (p.mode == 1) {
      if (b < p.nblocks) {
        // timf2.c:1001-1011: first half added onto what the previous transform (or call) parked there
        const float2* cur = p.tmp + (size_t)b * S * N + i;
#pragma unroll
        for (int s = 0; s < S; s++) {
          const float2 a = cur[(size_t)s * N];
          float2 old;
          if (b > 0) {
            const float2 q = p.tmp[((size_t)(b - 1) * S + s) * N + N / 2 + i];
            old = make_float2(__fmul_rn(q.x, p.ampfac), __fmul_rn(q.y, p.ampfac));
          } else {
            old = *reinterpret_cast<const float2*>(dst + 2 * s);
          }
          o[2 * s] = __fadd_rn(old.x, __fmul_rn(p.ampfac, a.x));
          o[2 * s + 1] = __fadd_rn(old.y, __fmul_rn(p.ampfac, a.y));
        }
      } else {
        // timf2.c:1012-1020: the last transform's second half, stored for the next call to add onto
        const float2* last = p.tmp + (size_t)(p.nblocks - 1) * S * N + N / 2 + i;
#pragma unroll
        for (int s = 0; s < S; s++) {
          const float2 q = last[(size_t)s * N];
          o[2 * s] = __fmul_rn(q.x, p.ampfac);
          o[2 * s + 1] = __fmul_rn(q.y, p.ampfac);
        }
      }
    } else {
      int src;
      float fac;
      if (p.mode == 0) {
        src = i;                                          // timf2.c:986-997
        fac = p.ampfac;
      } else {
        // timf2.c:1027-1056: centre portion; the window index runs ia..ib-1, then ib..ia+1
        const int ia = p.interleave / 2, ib = N / 2;
        src = ia + i;
        const int wi = i < ib - ia ? ia + i : ib - (i - (ib - ia));
        fac = __fmul_rn(p.invwin[wi], p.ampfac);
      }
      const float2* cur = p.tmp + (size_t)b * S * N + src;
#pragma unroll
      for (int s = 0; s < S; s++) {
        const float2 a = cur[(size_t)s * N];
        o[2 * s] = __fmul_rn(fac, a.x);
        o[2 * s + 1] = __fmul_rn(fac, a.y);
      }
    }
#pragma unroll
    for (int f = 0; f < SF; f += 4) *reinterpret_cast<float4*>(dst + f) = make_float4(o[f], o[f + 1], o[f + 2], o[f + 3]);
    if (b < p.nblocks) {
      // |weak|^2, summed in the reference's order (timf2.c:990-993, 1048-1055)
      float pw = __fmul_rn(o[0], o[0]);
#pragma unroll
      for (int f = 1; f < 2 * NCH; f++) pw = __fadd_rn(pw, __fmul_rn(o[f], o[f]));
      p.pwr[pos / SF] = pw;
    }
  }
}

}  // namespace lb
