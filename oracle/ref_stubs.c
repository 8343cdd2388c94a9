/* oracle/ref_stubs.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Link-time stubs for the handful of functions that the reference's hot-path
 * objects (fft0.c fft1.c fft1_re.c mix1.c + the *var.c files, compiled
 * unmodified from /root/reference by oracle/Makefile) call into the GUI / OS
 * layers of Linrad.  None of them takes part in the arithmetic of the path
 * except new_fft1_averages, which is restated from wide_graph.c:1003-1051
 * (wide_graph.c itself cannot be linked: it drags in the whole screen layer).
 */
#include <stdio.h>
#include <stdlib.h>
#include "osnum.h"
#include "globdef.h"
#include "uidef.h"
#include "fft1def.h"
#include "screendef.h"

int ref_last_lirerr = 0;

void lirerr(int errcod)            /* lxsys.c:495 -- record, never spin */
{
  if (ref_last_lirerr == 0) ref_last_lirerr = errcod;
  fprintf(stderr, "[ref oracle] lirerr(%d)\n", errcod);
}
void lir_sched_yield(void) {}
void lir_sleep(int us) { (void)us; }
void lir_mutex_lock(int no) { (void)no; }
void lir_mutex_unlock(int no) { (void)no; }
void lir_text(int x, int y, char *txt) { (void)x; (void)y; (void)txt; }
void lir_pixwrite(int x, int y, char *s) { (void)x; (void)y; (void)s; }
void settextcolor(unsigned char color) { (void)color; }
void awake_screen(void) {}
/* fft3.c (make_fft3_all): only its transform part is driven, these belong to the display half */
void add_mix1_cursor(int n) { (void)n; }
void clear_thread_times(int n) { (void)n; }
void lir_await_event(int n) { (void)n; }
void parabolic_fit(float *amp, float *pos, float yy1, float yy2, float yy3) { (void)yy1; (void)yy2; (void)yy3; *amp = 0; *pos = 0; }
void eliminate_spurs(void) {}
void spursearch_spectrum_cleanup(void) {}
void expand_foldcorr(float *x, float *tmp) { (void)x; (void)tmp; }
void normalise_fft1_filtercorr(float *xyzq) { (void)xyzq; }

int make_power_of_two(int *i)
{
  int k = 1, n = 0;
  while (k < *i) { k <<= 1; n++; }
  *i = k;
  return n;
}

/* Restatement of wide_graph.c:1003-1051: rebuild fft1_slowsum over bins
 * [ia,ib] as the sum of the latest wg_fft_avg2num rows of fft1_sumsq, the
 * newest of which starts at ptr.  (correlation spectra are not on the path) */
void new_fft1_averages(int ptr, int ia, int ib)
{
  int row, bin, src;
  latest_wg_spectrum++;
  change_fft1_flag = FALSE;
  if (ia < 0 || ib < ia || ib >= fft1_size) { lirerr(521233); return; }
  src = (ptr - (wg_fft_avg2num - 1) * fft1_size + fft1_sumsq_bufsize) & fft1_sumsq_mask;
  for (bin = ia; bin <= ib; bin++) fft1_slowsum[bin] = fft1_sumsq[src + bin];
  for (row = 1; row < wg_fft_avg2num; row++) {
    src = (src + fft1_size) & fft1_sumsq_mask;
    for (bin = ia; bin <= ib; bin++) {
      fft1_slowsum[bin] += fft1_sumsq[src + bin];
      if (fft1_slowsum[bin] < FFT1_SMALL) fft1_slowsum[bin] = FFT1_SMALL;
    }
  }
  if (fft1_correlation_flag == 1) {                       /* wide_graph.c:1031-1050 */
    src = (ptr - (wg_fft_avg2num - 1) * fft1_size + fft1_sumsq_bufsize) & fft1_sumsq_mask;
    for (bin = ia; bin <= ib; bin++) {
      fft1_slowcorr[2 * bin] = fft1_corrsum[2 * (src + bin)];
      fft1_slowcorr[2 * bin + 1] = fft1_corrsum[2 * (src + bin) + 1];
    }
    for (row = 1; row < wg_fft_avg2num; row++) {
      src = (src + fft1_size) & fft1_sumsq_mask;
      for (bin = ia; bin <= ib; bin++) {
        fft1_slowcorr[2 * bin] += fft1_corrsum[2 * (src + bin)];
        fft1_slowcorr[2 * bin + 1] += fft1_corrsum[2 * (src + bin) + 1];
      }
    }
  }
}
/* timf2.c: the short-int / MMX halves of the split and the back transform are never taken (swfloat = 1) */
void split_one(void) { lirerr(99001); }
void split_two(void) { lirerr(99002); }

