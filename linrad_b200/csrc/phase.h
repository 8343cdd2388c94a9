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

LBP_HD float lb_phase_advance(float x, float d, int n)
{
  while (n > 0) {
    bool jumped = false;
    if (d != 0.0f && x != 0.0f && isfinite(x) && isfinite(d)) {
      const float ax = fabsf(x);
      int e;
      (void)frexpf(ax, &e);              // ax = f * 2^e, f in [0.5,1)  ->  ax in [2^(e-1), 2^e)
      e -= 1;                            // ax in [2^e, 2^(e+1))
      if (e > -100) {                    // stay clear of denormals
        const double sgn = x < 0 ? -1.0 : 1.0;
        const double m = ldexp((double)ax, 23 - e);          // integer in [2^23, 2^24)
        const double qd = sgn * ldexp((double)d, 23 - e);    // step in ulps, towards +|x| if >0
        if (fabs(qd) < 8388608.0) {
          const double fl = floor(qd);
          const double frac = qd - fl;
          double qi;
          bool ok = true;
          if (frac == 0.5) {
            // tie: result goes to the even neighbour; only stable once m is even
            if (fmod(m, 2.0) != 0.0) ok = false;
            qi = (fmod(fl, 2.0) == 0.0) ? fl : fl + 1.0;
          } else {
            qi = (frac < 0.5) ? fl : fl + 1.0;
          }
          // at the bottom edge of the binade a step towards zero lands on the finer grid below
          if (qd < 0 && m < 8388609.0) ok = false;
          if (ok) {
            if (qi == 0.0) return x;       // x has stopped moving
            double smax;
            if (qi > 0) smax = floor((16777216.0 - m) / qi);
            else smax = floor((m - 8388609.0) / (-qi));   // keep the exact sum inside the binade
            if (smax >= 1.0) {
              double s = smax < (double)n ? smax : (double)n;
              const double m2 = m + s * qi;
              x = (float)(sgn * ldexp(m2, e - 23));
              n -= (int)s;
              jumped = true;
            }
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
