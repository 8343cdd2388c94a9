// fft32_core.cuh -- 32-points-per-thread Stockham plan for N = 2^10 .. 2^14 (sm_100a).
//
// One transform of N = 32*T points is spread over T threads, each holding 32 points in
// registers with the invariant  v[e] == data[t + T*e]  before the first and after the last
// pass (unit-stride global access, no shuffle step).  N = R0 * 32 [* 32]:
//     pass 0   radix R0 = N/1024 (2..16; 32 for N = 1024), no twiddles, 32/R0 butterflies/thread
//     pass 1   radix 32, Ns = R0: only R0 distinct twiddle sets -> exact table in shared memory
//     pass 2   radix 32, Ns = N/32 = T: one twiddle set per thread, generated from the five
//              correctly rounded binary powers w^1,w^2,w^4,w^8,w^16 the thread keeps in registers
//     (N = 1024 has no middle pass)
// Between passes the points go through shared memory (Stockham auto-sort scatter); the first
// exchange is written with 128-bit stores.  Replaces the reference's bulk_of_dit / bulk_of_dif
// (fft0.c:1590-1769, 161-195) like fft_core.cuh does, at about 60 % of its instruction count:
//   * every non-trivial radix-2 butterfly is  lo = e + w*o (4 FFMA),  hi = 2e - lo (2 FFMA)
//     instead of complex multiply + add + subtract (8 instructions);
//   * the external twiddle of input r+16 is folded into the first butterfly level the same way.
// Sign convention: forward, exp(-2 pi i n k / N), unnormalised.
#pragma once
#include "fft_core.cuh"

namespace lb {

// lo = a + w*b, hi = 2a - lo for a run-time twiddle w
LB_HD void bfly_w(float2& a, float2& b, float2 w)
{
  const float lx = fmaf(-b.y, w.y, fmaf(b.x, w.x, a.x));
  const float ly = fmaf(b.x, w.y, fmaf(b.y, w.x, a.y));
  b = make_float2(fmaf(2.0f, a.x, -lx), fmaf(2.0f, a.y, -ly));
  a = make_float2(lx, ly);
}

// In-place forward DFT of the R points x[0], x[S], ..., natural order in and out (radix-2
// decimation in time; all indices are compile-time constants after unrolling).  With SKIP2
// the innermost level (pairs x[i], x[i+R/2]) has already been done by the caller.
template <int R, int S, bool SKIP2>
struct Dft32 {
  static LB_HD void run(float2* x)
  {
    Dft32<R / 2, 2 * S, SKIP2>::run(x);
    Dft32<R / 2, 2 * S, SKIP2>::run(x + S);
    float2 lo[R / 2], hi[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; k++) {
      float2 e = x[2 * S * k], o = x[S + 2 * S * k];
      bfly32(e, o, k * (32 / R));
      lo[k] = e;
      hi[k] = o;
    }
#pragma unroll
    for (int k = 0; k < R / 2; k++) {
      x[S * k] = lo[k];
      x[S * (k + R / 2)] = hi[k];
    }
  }
};
template <int S, bool SKIP2>
struct Dft32<2, S, SKIP2> {
  static LB_HD void run(float2* x)
  {
    if (!SKIP2) bfly32(x[0], x[S], 0);
  }
};
template <int S, bool SKIP2>
struct Dft32<1, S, SKIP2> {
  static LB_HD void run(float2*) {}
};

// ---- pass 0: Q = 32/R0 untwiddled radix-R0 butterflies; butterfly q works on v[q + r*Q]
template <int R0>
LB_HD void pass0(float2 (&v)[32])
{
  constexpr int Q = 32 / R0;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    float2 x[R0];
#pragma unroll
    for (int r = 0; r < R0; r++) x[r] = v[q + r * Q];
    Dft32<R0, 1, false>::run(x);
#pragma unroll
    for (int r = 0; r < R0; r++) v[q + r * Q] = x[r];
  }
}

// ---- twiddled radix-32 butterfly, all 31 twiddles given (w[r] = w^r, w[0] unused)
LB_HD void radix32_table(float2 (&x)[32], const float2 (&w)[32])
{
#pragma unroll
  for (int i = 0; i < 16; i++) {
    float2 a = x[i];
    if (i > 0) a = cmul(a, w[i]);
    float2 b = x[i + 16];
    bfly_w(a, b, w[i + 16]);
    x[i] = a;
    x[i + 16] = b;
  }
  Dft32<32, 1, true>::run(x);
}

// ---- twiddled radix-32 butterfly, twiddles generated from the exact binary powers
// wb[j] = w^(2^j), j = 0..4: w^r for r < 16 is (w^4h)*(w^l), r = 4h+l, and w^(r+16) = w^r * w^16,
// so no power is more than three rounded products away from an exact table value.
LB_HD void radix32_gen(float2 (&x)[32], const float2 (&wb)[5])
{
  float2 lo[4], hi[4];
  lo[1] = wb[0];
  lo[2] = wb[1];
  lo[3] = cmul(wb[0], wb[1]);
  hi[1] = wb[2];
  hi[2] = wb[3];
  hi[3] = cmul(wb[2], wb[3]);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int l = i & 3, h = i >> 2;
    float2 a = x[i], b = x[i + 16];
    float2 wr, wq;
    if (i == 0) {
      wq = wb[4];
    } else {
      if (h == 0) wr = lo[l];
      else if (l == 0) wr = hi[h];
      else wr = cmul(hi[h], lo[l]);
      a = cmul(a, wr);
      wq = cmul(wr, wb[4]);
    }
    bfly_w(a, b, wq);
    x[i] = a;
    x[i + 16] = b;
  }
  Dft32<32, 1, true>::run(x);
}

// ---- shared-memory layouts of the two exchanges (each is rewritten completely, so each can
// have the padding that makes its own scatter conflict-free)
// exchange 1 is written with 128-bit stores, which the LSU serves a quarter warp (8 lanes) at a
// time: lane j of a quarter writes at j*R0 slots, so one 16-byte skew per 16 slots (per 32 for
// R0 = 32) spreads the eight lanes over the eight 16-byte bank groups.
template <int SH>
LB_HD int pad2(int i) { return i + ((i >> SH) << 1); }
LB_HD int pad1(int i) { return i + (i >> 5); }           // exchange 2 (64-bit stores)

template <int LOG2N>
struct Plan32 {
  static constexpr int N = 1 << LOG2N;
  static constexpr int T = N / 32;
  static constexpr int NPASS = LOG2N > 10 ? 3 : 2;
  static constexpr int R0 = LOG2N > 10 ? (N >> 10) : 32;
  static constexpr int Q = 32 / R0;
  static constexpr int SH1 = R0 == 32 ? 5 : 4;             // exchange-1 skew granularity (log2 slots)
  static constexpr int XCH = N + N / 8 + 16;               // float2 slots of the exchange buffer
  static constexpr int TAB1 = NPASS == 3 ? 16 * R0 : 0;    // float4 entries of the pass-1 table
};

// exchange 1 (after pass 0, Ns = 1): butterfly j = t + T*q writes its R0 outputs to j*R0 + r
template <int LOG2N>
LB_HD void exch1_store(const float2 (&v)[32], float2* sm, int t)
{
  using P = Plan32<LOG2N>;
  // pad2((t + T*q)*R0) = pad2(t*R0) + q*(T*R0 + 2*T*R0/2^SH1): one base, compile-time offsets
  float2* p = sm + pad2<P::SH1>(t * P::R0);
#pragma unroll
  for (int q = 0; q < P::Q; q++) {
#pragma unroll
    for (int r = 0; r < P::R0; r += 2) {
      const float2 a = v[q + r * P::Q], b = v[q + (r + 1) * P::Q];
      *reinterpret_cast<float4*>(p + q * (P::T * P::R0 + ((P::T * P::R0) >> P::SH1) * 2) + r) = make_float4(a.x, a.y, b.x, b.y);
    }
  }
}
template <int LOG2N>
LB_HD void exch1_load(float2 (&v)[32], const float2* sm, int t)
{
  using P = Plan32<LOG2N>;
  const float2* p = sm + pad2<P::SH1>(t);
#pragma unroll
  for (int e = 0; e < 32; e++) v[e] = p[e * (P::T + (P::T >> P::SH1) * 2)];
}
// exchange 2 (after pass 1, Ns = R0): output r of butterfly t goes to (t-k)*32 + k + r*R0
template <int LOG2N>
LB_HD void exch2_store(const float2 (&v)[32], float2* sm, int t)
{
  using P = Plan32<LOG2N>;
  // pad1((t-k)*32 + k + r*R0) = 33*(t-k) + k + r*R0 + (r*R0)/32 because (r*R0 mod 32) + k < 32
  const int k = t & (P::R0 - 1);
  float2* p = sm + 33 * (t - k) + k;
#pragma unroll
  for (int r = 0; r < 32; r++) p[r * P::R0 + (r * P::R0) / 32] = v[r];
}
template <int LOG2N>
LB_HD void exch2_load(float2 (&v)[32], const float2* sm, int t)
{
  using P = Plan32<LOG2N>;
  const float2* p = sm + pad1(t);
#pragma unroll
  for (int e = 0; e < 32; e++) v[e] = p[e * (P::T + P::T / 32)];
}

}  // namespace lb
