// Fuzz: lb_phase_stepper::advance (table-driven, used for the bulk mix1 job table) and
// lb_phase_advance (generic, used on the device) both equal the reference's running float sum
// `for(i=0;i<n;i++) x+=d;` (mix1.c:146-153) bit for bit.  Prints "bad=<count>".
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../../linrad_b200/csrc/phase.h"
int main()
{
  srand(1);
  long bad = 0;
  for (int it = 0; it < 60000; it++) {
    float x = (float)((rand() / (double)RAND_MAX - 0.5) * 20.0);
    float d = (float)((rand() / (double)RAND_MAX - 0.5) * 1.6 * pow(10.0, -(rand() % 7)));
    if (it % 50 == 0) d = ldexpf(1.0f, -(rand() % 30));
    if (it % 77 == 0) x = ldexpf(1.0f, (rand() % 8) - 3);
    if (it % 91 == 0) d = -ldexpf(1.0f, -(rand() % 26));
    if (it % 1001 == 0) d = 0.0f;
    if (it % 1003 == 0) x = 0.0f;
    const int n = rand() % 3000;
    lb_phase_stepper st(d);
    const float a = lb_phase_advance(x, d, n), b = st.advance(x, n);
    float c = x;
    for (int i = 0; i < n; i++) c = lb_float_add(c, d);
    if (lb_f2u(a) != lb_f2u(c) || lb_f2u(b) != lb_f2u(c)) {
      if (bad < 5) printf("mismatch x=%a d=%a n=%d: generic %a stepper %a sum %a\n", x, d, n, a, b, c);
      bad++;
    }
  }
  printf("bad=%ld\n", bad);
  return bad ? 1 : 0;
}
