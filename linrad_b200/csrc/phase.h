// phase.h -- exact emulation of the reference's float phase accumulator.
//
// do_mix1 advances the mixer phase with a running single-precision sum, one addition per
// output sample (mix1.c:146-153, 172-186, 225-258: "t1+=t2"), and carries the result to the
// next transform (mix1.c:154,187,260 -> mix1_phase[ss]).  To reproduce timf3 for an arbitrary
// sample of an arbitrary block without iterating, phase_advance(x, d, n) returns exactly the
// value of  `for(i=0;i<n;i++) x+=d;`  in IEEE binary32 round-to-nearest-even:
// while x stays inside one binade every addition moves it by the same whole number of ulps,
// so the run can be skipped in one step; the additions that cross a power of two are done
// one at a time with a real float add.
#pragma once
#include <math.h>
#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#if defined(__CUDACC__)
#define LBP_HD __host__ __device__ inline
#else
#define LBP_HD static inline
#endif

LBP_HD float lb_float_add(float a, float b)
{
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);     // never contracted, never reassociated
#else
  volatile float r = a + b;
  return r;
#endif
}

LBP_HD uint32_t lb_f2u(float f)
{
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
LBP_HD double lb_pow2(int k)          /* 2^k as a double, -1022 <= k <= 1023 */
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)(1023 + k) << 52);
#else
  union { double d; uint64_t u; } c; c.u = (uint64_t)(1023 + k) << 52; return c.d;
#endif
}

LBP_HD float lb_phase_advance(float x, float d, int n)
{
  if ((lb_f2u(d) & 0x7fffffffu) == 0u)      /* d = +-0 (a selection on an exact bin): x never moves */
    return (n > 0) ? lb_float_add(x, d) : x;   /* one add settles the sign of a zero x */
  while (n > 0) {
    bool jumped = false;
    const uint32_t bx = lb_f2u(x), bd = lb_f2u(d);
    const int ex = (int)((bx >> 23) & 0xffu);          /* biased exponent: |x| in [2^(ex-127), 2^(ex-126)) */
    /* x normal and not tiny (stay clear of denormals), d finite and non-zero */
    if (ex > 27 && ex != 255 && ((bd >> 23) & 0xffu) != 255u && (bd & 0x7fffffffu) != 0u) {
      const long long m = (long long)((bx & 0x7fffffu) | 0x800000u);   /* integer in [2^23, 2^24) */
      const bool neg = (bx >> 31) != 0u;
      /* step in ulps of x, counted towards +|x| when positive (exact: a power-of-two scaling) */
      const double qd = (neg ? -(double)d : (double)d) * lb_pow2(150 - ex);
      if (qd < 8388608.0 && qd > -8388608.0) {
        long long fli = (long long)qd;                 /* floor(qd) */
        if ((double)fli > qd) fli -= 1;
        const double frac = qd - (double)fli;
        long long qi;
        bool ok = true;
        if (frac == 0.5) {
          /* tie: result goes to the even neighbour; only stable once m is even */
          if (m & 1) ok = false;
          qi = (fli & 1) ? fli + 1 : fli;
        } else {
          qi = (frac < 0.5) ? fli : fli + 1;
        }
        /* at the bottom edge of the binade a step towards zero lands on the finer grid below */
        if (qd < 0 && m < 8388609) ok = false;
        if (ok) {
          if (qi == 0) return x;                       /* x has stopped moving */
          /* longest run that keeps the exact sum inside the binade (the division is only
             needed when the run would leave it) */
          long long smax = (long long)n;
          const long long end = m + smax * qi;
          if (qi > 0) { if (end > 16777216) smax = (16777216 - m) / qi; }
          else { if (end < 8388609) smax = (m - 8388609) / (-qi); }
          if (smax >= 1) {
            const long long s = smax;
            const double r = (double)(m + s * qi) * lb_pow2(ex - 150);
            x = (float)(neg ? -r : r);
            n -= (int)s;
            jumped = true;
          }
        }
      }
    }
    if (!jumped) {
      x = lb_float_add(x, d);
      n -= 1;
    }
  }
  return x;
}

#if defined(__cplusplus)
// Host-side stepper for long runs with one fixed increment (the job table of a bulk lb200_mix1
// call advances every selection's phase once per transform): for a fixed d the whole-ulp step
// and its tie/edge flags depend only on the sign and exponent of x, so they are cached per
// (sign, exponent) and an advance inside a binade is a handful of integer operations.
// Returns exactly what lb_phase_advance returns.
struct lb_phase_stepper {
  float d;
  int qi[512];              // step in ulps towards +|x|, per (sign<<8 | biased exponent)
  unsigned char st[512];    // 0 unknown, 1 usable, 2 tie (m must be even), 3 not usable
  explicit lb_phase_stepper(float step) : d(step) { for (int i = 0; i < 512; i++) st[i] = 0; }
  void fill(int key)
  {
    const int ex = key & 255;
    const bool neg = (key >> 8) != 0;
    const uint32_t bd = lb_f2u(d);
    st[key] = 3;
    qi[key] = 0;
    if (!(ex > 27 && ex != 255 && ((bd >> 23) & 0xffu) != 255u && (bd & 0x7fffffffu) != 0u)) return;
    const double qd = (neg ? -(double)d : (double)d) * lb_pow2(150 - ex);
    if (!(qd < 8388608.0 && qd > -8388608.0)) return;
    long long fli = (long long)qd;
    if ((double)fli > qd) fli -= 1;
    const double frac = qd - (double)fli;
    if (frac == 0.5) { st[key] = 2; qi[key] = (int)((fli & 1) ? fli + 1 : fli); }
    else { st[key] = 1; qi[key] = (int)((frac < 0.5) ? fli : fli + 1); }
  }
  float advance(float x, int n)
  {
    if ((lb_f2u(d) & 0x7fffffffu) == 0u) return (n > 0) ? lb_float_add(x, d) : x;
    while (n > 0) {
      const uint32_t bx = lb_f2u(x);
      const int key = (int)(bx >> 23);
      if (st[key] == 0) fill(key);
      bool jumped = false;
      const long long m = (long long)((bx & 0x7fffffu) | 0x800000u);
      const int s_ = st[key];
      const long long q = qi[key];
      if (s_ != 3 && !(s_ == 2 && (m & 1)) && !(m < 8388609 && lb_toward_zero(bx, d))) {
        if (q == 0) return x;
        long long smax = (long long)n;
        const long long end = m + smax * q;
        if (q > 0) { if (end > 16777216) smax = (16777216 - m) / q; }
        else { if (end < 8388609) smax = (m - 8388609) / (-q); }
        if (smax >= 1) {
          const long long m2 = m + smax * q;
          uint32_t br;
          if (m2 >= 16777216) br = (bx & 0x80000000u) | ((uint32_t)((key & 255) + 1) << 23);   // exactly 2^(e+1)
          else br = (bx & 0xff800000u) | ((uint32_t)m2 & 0x7fffffu);
          union { uint32_t u; float f; } c; c.u = br; x = c.f;
          n -= (int)smax;
          jumped = true;
        }
      }
      if (!jumped) { x = lb_float_add(x, d); n -= 1; }
    }
    return x;
  }
  // qd < 0  <=>  the step moves x towards zero
  static bool lb_toward_zero(uint32_t bx, float dd) { return ((bx >> 31) != 0u) != (dd < 0.0f); }
};
#endif
