// fft_core.cuh -- register-resident Stockham FFT building blocks for sm_100a.
//
// One transform of N = E*T points is spread over T threads, each holding E points in
// registers as float2 v[E] with the invariant
//        v[e] == data[t + T*e]                       (t = thread index within the transform)
// both before the first pass (natural-order input) and after the last pass (natural-order
// output), so global loads and stores are unit-stride across threads with no shuffle step.
// A pass of radix R (R | E) does E/R butterflies per thread; between passes the points are
// re-distributed through shared memory (the Stockham auto-sort scatter).  This replaces the
// reference's table-driven in-place CPU kernels: bulk_of_dit (fft0.c:1590-1769, radix-4 DIT
// + fft1_permute gather), bulk_of_dif (fft0.c:161-195) and fftback (fft0.c:481-533).
// Sign convention: forward, X[k] = sum_n x[n] exp(-2*pi*i*n*k/N), unnormalised (what the
// probed reference kernels compute).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef LB_HD
#define LB_HD __host__ __device__ __forceinline__
#endif
#ifndef LB_D
#define LB_D __device__ __forceinline__
#endif

namespace lb {

// ---------------------------------------------------------------- small complex helpers
LB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
LB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
LB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// cos(2*pi*k/32), k = 0..31 (folds to a literal once k is a compile-time constant)
LB_HD constexpr float cos32(int k)
{
  k &= 31;
  if (k > 16) k = 32 - k;
  switch (k) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323044913f;
    case 2: return 0.92387953251128675613f;
    case 3: return 0.83146961230254523708f;
    case 4: return 0.70710678118654752440f;
    case 5: return 0.55557023301960222474f;
    case 6: return 0.38268343236508977173f;
    case 7: return 0.19509032201612826785f;
    case 8: return 0.0f;
    case 9: return -0.19509032201612826785f;
    case 10: return -0.38268343236508977173f;
    case 11: return -0.55557023301960222474f;
    case 12: return -0.70710678118654752440f;
    case 13: return -0.83146961230254523708f;
    case 14: return -0.92387953251128675613f;
    case 15: return -0.98078528040323044913f;
    default: return -1.0f;
  }
}
LB_HD constexpr float sin32(int k) { return cos32(k - 8); }

// a * exp(-2*pi*i*k32/32) with the trivial rotations done without multiplies
LB_HD float2 rot32(float2 a, int k32)
{
  k32 &= 31;
  switch (k32) {
    case 0: return a;
    case 8: return make_float2(a.y, -a.x);
    case 16: return make_float2(-a.x, -a.y);
    case 24: return make_float2(-a.y, a.x);
    case 4: return make_float2((a.x + a.y) * 0.70710678118654752440f, (a.y - a.x) * 0.70710678118654752440f);
    case 12: return make_float2((a.y - a.x) * 0.70710678118654752440f, -(a.x + a.y) * 0.70710678118654752440f);
    case 20: return make_float2(-(a.x + a.y) * 0.70710678118654752440f, (a.x - a.y) * 0.70710678118654752440f);
    case 28: return make_float2((a.x - a.y) * 0.70710678118654752440f, (a.x + a.y) * 0.70710678118654752440f);
    default: {
      const float c = cos32(k32), s = sin32(k32);      // exp(-i th) = c - i s
      return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
    }
  }
}

// e' = e + W*o, o' = e - W*o with W = exp(-2 pi i k32/32), k32 in [0,16)
LB_HD void bfly32(float2& e, float2& o, int k32)
{
  if (k32 == 0) {
    const float2 t = e;
    e = make_float2(t.x + o.x, t.y + o.y);
    o = make_float2(t.x - o.x, t.y - o.y);
  } else if (k32 == 8) {                 // W = -i:  W*o = (o.y, -o.x)
    const float2 t = e, u = o;
    e = make_float2(t.x + u.y, t.y - u.x);
    o = make_float2(t.x - u.y, t.y + u.x);
  } else {                               // W = c - i s:  W*o = (o.x c + o.y s, o.y c - o.x s)
    const float c = cos32(k32), s = sin32(k32);
    const float lx = fmaf(o.y, s, fmaf(o.x, c, e.x));
    const float ly = fmaf(-o.x, s, fmaf(o.y, c, e.y));
    o = make_float2(fmaf(2.0f, e.x, -lx), fmaf(2.0f, e.y, -ly));
    e = make_float2(lx, ly);
  }
}

// ---------------------------------------------------------------- DFT_R on registers
// In-place forward DFT of the R points x[0], x[S], ..., x[(R-1)S]; natural order in and out.
// Radix-2 decimation in time; every index is a compile-time constant after unrolling so the
// "temporary" arrays are pure register renaming.
template <int R, int S>
struct DftS {
  static LB_HD void run(float2* x)
  {
    DftS<R / 2, 2 * S>::run(x);
    DftS<R / 2, 2 * S>::run(x + S);
    float2 lo[R / 2], hi[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; k++) {
      float2 e = x[2 * S * k], o = x[S + 2 * S * k];
      bfly32(e, o, k * (32 / R));            // lo = e + W o (4 FFMA), hi = 2e - lo (2 FFMA)
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
template <int S>
struct DftS<1, S> {
  static LB_HD void run(float2*) {}
};

// ---------------------------------------------------------------- one Stockham pass
// Geometry of a pass: N points, T threads, E = N/T points per thread, radix R, Ns = product
// of the radices of the earlier passes.  Butterfly q of thread t is Stockham butterfly
// j = t + T*q; it consumes v[q + r*(E/R)], r = 0..R-1 (== data[j + r*N/R]) and its r-th
// output belongs at index (j-k)*R + k + r*Ns with k = j mod Ns.
//
// Twiddles: input r is multiplied by w^r, w = exp(-2*pi*i*k/(Ns*R)).  Raising one rounded w
// to the r-th power multiplies its angle error by r (measured: 4e-7 relative rms on the whole
// transform instead of 1.4e-7), so the binary powers w^1, w^2, w^4, ... are each fetched
// correctly rounded from the table (wb[b] = w^(2^b)) and w^r is assembled from at most
// log2(R) of them with a two-level split r = hi*2^LB + lo that keeps few twiddles live.
template <int R>
struct Log2 { static constexpr int value = 1 + Log2<R / 2>::value; };
template <>
struct Log2<1> { static constexpr int value = 0; };

template <int E, int R, int T>
LB_HD void pass_butterflies(float2 (&v)[E], const float2* wb, bool twiddled)
{
  constexpr int Q = E / R;
  constexpr int LR = Log2<R>::value;
  constexpr int LB = LR / 2;               // low bits
  constexpr int NLO = 1 << LB, NHI = R >> LB;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    float2 x[R];
#pragma unroll
    for (int r = 0; r < R; r++) x[r] = v[q + r * Q];
    if (twiddled && R > 1) {
      const float2* w2 = wb + q * (LR > 0 ? LR : 1);
      float2 lo[NLO > 1 ? NLO : 2], hi[NHI > 1 ? NHI : 2];
      // lo[i] = w^i (i < 2^LB), hi[j] = w^(j*2^LB): binary products of the exact powers
#pragma unroll
      for (int i = 1; i < NLO; i++) {
        const int hb = 31 - __builtin_clz(i);            // highest set bit
        if ((i & (i - 1)) == 0) lo[i] = w2[hb];
        else lo[i] = cmul(w2[hb], lo[i - (1 << hb)]);
      }
#pragma unroll
      for (int j = 1; j < NHI; j++) {
        const int hb = 31 - __builtin_clz(j);
        if ((j & (j - 1)) == 0) hi[j] = w2[LB + hb];
        else hi[j] = cmul(w2[LB + hb], hi[j - (1 << hb)]);
      }
#pragma unroll
      for (int r = 1; r < R; r++) {
        const int l = r & (NLO - 1), h = r >> LB;
        float2 w;
        if (h == 0) w = lo[l];
        else if (l == 0) w = hi[h];
        else w = cmul(hi[h], lo[l]);
        x[r] = cmul(x[r], w);
      }
    }
    DftS<R, 1>::run(x);
#pragma unroll
    for (int r = 0; r < R; r++) v[q + r * Q] = x[r];
  }
}

// index of w^(2^b) for butterfly q of thread t inside the table W[m] = exp(-2*pi*i*m/N):
// m = 2^b * k * N/(Ns*R)   (always < N because 2^b <= R/2 and k < Ns)
template <int E, int R, int T>
LB_HD int tw_index(int t, int q, int Ns, int b)
{
  constexpr int N = E * T;
  const int j = t + T * q;
  const int k = j & (Ns - 1);
  return (k * (N / (Ns * R))) << b;
}

// shared-memory slot for logical index i; one pad slot every 32 keeps the strided scatter of
// the first exchange spread over the banks.
template <int LOG2PAD>
LB_HD int padded(int i) { return LOG2PAD > 0 ? i + (i >> LOG2PAD) : i; }

template <int E, int R, int T, int LOG2PAD, int STRIDE = 1>
LB_HD void exchange_store(const float2 (&v)[E], float2* sm, int t, int Ns)
{
  constexpr int Q = E / R;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int j = t + T * q;
    const int k = j & (Ns - 1);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; r++) sm[padded<LOG2PAD>(base + r * Ns) * STRIDE] = v[q + r * Q];
  }
}

template <int E, int T, int LOG2PAD, int STRIDE = 1>
LB_HD void exchange_load(float2 (&v)[E], const float2* sm, int t)
{
#pragma unroll
  for (int e = 0; e < E; e++) v[e] = sm[padded<LOG2PAD>(t + T * e) * STRIDE];
}

// v[e] *= b * s^e, e = 0..E-1, with s given by its exact binary powers sb[j] = s^(2^j)
// (the inter-step twiddle of the four-step transform, same accuracy argument as above)
template <int E>
LB_HD void apply_power_twiddles(float2 (&v)[E], float2 b, const float2* sb)
{
  constexpr int LE = Log2<E>::value;
  constexpr int LB = LE / 2;
  constexpr int NLO = 1 << LB, NHI = E >> LB;
  float2 lo[NLO > 1 ? NLO : 2], hi[NHI > 1 ? NHI : 2];
  lo[0] = b;
#pragma unroll
  for (int i = 1; i < NLO; i++) {
    const int hb = 31 - __builtin_clz(i);
    lo[i] = cmul(sb[hb], lo[i - (1 << hb)]);
  }
#pragma unroll
  for (int j = 1; j < NHI; j++) {
    const int hb = 31 - __builtin_clz(j);
    if ((j & (j - 1)) == 0) hi[j] = sb[LB + hb];
    else hi[j] = cmul(sb[LB + hb], hi[j - (1 << hb)]);
  }
#pragma unroll
  for (int e = 0; e < E; e++) {
    const int l = e & (NLO - 1), h = e >> LB;
    const float2 w = (h == 0) ? lo[l] : cmul(hi[h], lo[l]);
    v[e] = cmul(v[e], w);
  }
}

// ---------------------------------------------------------------- radix plans
// Radix sequence for N = 2^LOG2N with E points per thread: as many radix-E passes as fit,
// preceded by one smaller pass for the remainder (the first pass needs no twiddles, so the
// odd-sized one goes first).
template <int LOG2N, int LOG2E>
struct Plan {
  static constexpr int N = 1 << LOG2N;
  static constexpr int E = 1 << LOG2E;
  static constexpr int T = N / E;
  static constexpr int REM = LOG2N % LOG2E;                 // log2 of the leading small radix
  static constexpr int NPASS = LOG2N / LOG2E + (REM ? 1 : 0);
  static constexpr int R0 = REM ? (1 << REM) : E;           // radix of pass 0
  // radix of pass p, and Ns before pass p
  static LB_HD constexpr int radix(int p) { return p == 0 ? R0 : E; }
  static LB_HD constexpr int ns(int p) { return p == 0 ? 1 : (R0 << (LOG2E * (p - 1))); }
  static constexpr int MAXQ = E / R0;                       // butterflies per thread in pass 0
  static constexpr int NTW = NPASS - 1;                     // twiddled passes (each has Q = 1 ... or E/R)
};

// Per-thread twiddle bases: for every pass p >= 1 (all of radix E, one butterfly per thread)
// the LOG2E exact binary powers of its w.
template <class P>
struct Twiddles {
  static constexpr int LE = Log2<P::E>::value;
  float2 w[(P::NTW > 0 ? P::NTW : 1) * (LE > 0 ? LE : 1)];
};

template <class P>
LB_D void load_twiddles(Twiddles<P>& tw, const float2* __restrict__ Wn, int t)
{
  constexpr int LE = Twiddles<P>::LE;
#pragma unroll
  for (int p = 1; p < P::NPASS; p++) {
    const int Ns = P::ns(p);
#pragma unroll
    for (int b = 0; b < LE; b++) tw.w[(p - 1) * LE + b] = Wn[tw_index<P::E, P::E, P::T>(t, 0, Ns, b)];
  }
}

#ifdef __CUDACC__
// The whole transform.  On entry v[e] = x[t + T*e]; on exit v[e] = X[t + T*e].
// `sm` must hold N + N/32 float2 (STRIDE==1) or N*STRIDE float2 with `sm` already offset by the
// batch lane (STRIDE>1: STRIDE transforms interleaved element by element); every thread of the CTA must call this (it contains
// __syncthreads), but several independent transforms may run side by side in one CTA as
// long as each gets its own `sm` slice and its own t in [0,T).
template <class P, int STRIDE = 1>
LB_D void fft_forward(float2 (&v)[P::E], float2* sm, int t, const Twiddles<P>& tw)
{
  constexpr int E = P::E, T = P::T;
  constexpr int LP = STRIDE == 1 ? 5 : 0;     // interleaved batches need no padding
  pass_butterflies<E, P::R0, T>(v, nullptr, false);
#pragma unroll
  for (int p = 1; p < P::NPASS; p++) {
    const int NsPrev = P::ns(p - 1);
    if (p == 1) exchange_store<E, P::R0, T, LP, STRIDE>(v, sm, t, NsPrev);
    else exchange_store<E, E, T, LP, STRIDE>(v, sm, t, NsPrev);
    __syncthreads();
    exchange_load<E, T, LP, STRIDE>(v, sm, t);
    __syncthreads();
    pass_butterflies<E, E, T>(v, &tw.w[(p - 1) * Twiddles<P>::LE], true);
  }
}
#endif

}  // namespace lb
