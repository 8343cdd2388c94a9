/* oracle/ref_driver.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Drives the reference's own hot-path functions (compiled unmodified from
 * /root/reference by oracle/Makefile into oracle/_ref/libref_oracle.so):
 *     fft1_b -> fft1_c -> fft1_waterfall -> fft1_mix1_fixed
 * on a synthetic timf1 ring, the way wideband_dsp()/narrowband_dsp() do
 * (wcw.c:1036-1085, wcw.c:1706-1716).  This file restates only the SIZING and
 * table-initialisation glue of buf.c (which cannot be linked, it pulls in the
 * GUI): get_wideband_sizes buf.c:139-332, timf3 sizes buf.c:645-657, table
 * init buf.c:1297-1300,1403-1456, prepare_mixer buf.c:55-111, and
 * make_wg_yfac wide_graph.c:956-1001.  All arithmetic on samples is done by
 * the reference functions themselves.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <signal.h>
#include <execinfo.h>
#include <unistd.h>
#include "osnum.h"
#include "globdef.h"
#include "uidef.h"
#include "fft1def.h"
#include "fft2def.h"
#include "fft3def.h"
#include "screendef.h"
#include "seldef.h"
#include "sigdef.h"
#include "thrdef.h"
#include "graphcal.h"

#define REF_MAX_SEL 64

typedef struct {
  int input_mode;        /* ui.rx_input_mode bits: DWORD_INPUT=1 TWO_CHANNELS=2 IQ_DATA=4 */
  int rf_channels;       /* ui.rx_rf_channels */
  int ad_speed;          /* ui.rx_ad_speed */
  int fft1_n;            /* log2(fft1_size) */
  int fft1_version;      /* index into fft_cntrl[] (6, 7, 2, 1, 10 ...) */
  int sinpow;            /* genparm[FIRST_FFT_SINPOW] */
  int fft1_gain;         /* genparm[FIRST_FFT_GAIN] */
  int mix1_red_n;        /* genparm[MIX1_BANDWIDTH_REDUCTION_N] */
  int avg1num;           /* wg.fft_avg1num */
  int avg2num;           /* wg_fft_avg2num */
  int waterfall_avgnum;  /* wg.waterfall_avgnum */
  int direction;         /* fft1_direction (+1/-1) */
  int n_sel;             /* number of mix1 selections driven by this harness */
  int first_xpoint;      /* wg.first_xpoint */
  int xpoints;           /* wg.xpoints */
  int xpoints_per_pixel; /* wg.xpoints_per_pixel */
  int pixels_per_xpoint; /* wg.pixels_per_xpoint */
  int wf_lines;          /* waterfall ring lines */
  int sample_shift;      /* ui.sample_shift */
  int correlation;       /* genparm[FFT1_CORRELATION_SPECTRUM] -> fft1_correlation_flag (two channels only, buf.c:1223-1224) */
  int afc;               /* fft1afc_flag: AFC from fft1 (second FFT off), fft1_c keeps fft1_power / fft1_xypower */
  int afc_mix;           /* drive fft1_mix1_afc with a synthetic AFC track instead of fft1_mix1_fixed (one selection) */
} ref_cfg;

typedef struct {
  float phase, phase_step, phase_rot, old_phase;
  int point, old_point;
  double selfreq;
} sel_state;

/* -DLB200_USE_SHIM: the same harness with the three hot-path calls routed through
 * linrad_b200/host/lb200_shim.c (i.e. liblinrad_b200.so on the GPU) -- the integration test of
 * the drop-in boundary: the shim reads the reference's OWN globals and tables. */
#ifdef LB200_USE_SHIM
int lb200_shim_open(int no_of_threads);
void lb200_shim_close(void);
void lb200_shim_fft1_b(int timf1p_ref, float *out, float *tmp, int gpu_handle_number);
void lb200_shim_fft1_c(void);
void lb200_shim_mix1_fixed(void);
void lb200_shim_mix1_afc(void);
extern void (*lb200_shim_afc_tables)(int ss);
#define HOT_MIX1_AFC lb200_shim_mix1_afc
#define HOT_FFT1_B lb200_shim_fft1_b
#define HOT_FFT1_C lb200_shim_fft1_c
#define HOT_MIX1_FIXED lb200_shim_mix1_fixed
void lb200_shim_make_timf2(void);
#define HOT_MAKE_TIMF2 lb200_shim_make_timf2
#else
#define HOT_MAKE_TIMF2 make_timf2
#define HOT_FFT1_B fft1_b
#define HOT_FFT1_C fft1_c
#define HOT_MIX1_FIXED fft1_mix1_fixed
#define HOT_MIX1_AFC fft1_mix1_afc
#endif
void fft1_mix1_afc(void);
void do_mix1_afc(int ss);        /* mix1.c:648, external linkage but no prototype in the headers */

/* the synthetic AFC track of the afc_mix harness: the selected frequency wobbling slowly */
static double afc_track(int ss, long tno);
#ifdef LB200_USE_SHIM
/* Stand-in for the split mix1.c a Linrad maintainer would make (see lb200_shim.c): run the
 * reference's own do_mix1_afc for its table maintenance with the mixer's output and phase state
 * held away, so that nothing of its CPU do_mix1 reaches the results. */
static float *afc_dummy_timf3;
static void afc_tables_via_reference(int ss)
{
  float *keep = timf3_float;
  float ph = mix1_phase[ss], oph = mix1_old_phase[ss];
  if (!afc_dummy_timf3) afc_dummy_timf3 = calloc((size_t)2 * timf3_size + 64, sizeof(float));
  timf3_float = afc_dummy_timf3;
  do_mix1_afc(ss);
  timf3_float = keep;
  mix1_phase[ss] = ph;
  mix1_old_phase[ss] = oph;
}
#endif

#if HAVE_CUFFT == 1
/* The reference's own GPU path (fft_cntrl row 19, fft1.c:3531-3553): the handles and device
 * buffers live in wcw.c:72-76, which is the thread body and is not compiled here, so the harness
 * owns them and creates them the way wcw.c:553-576 does.  Built only into _ref/libref_cufft.so. */
#include <cufft.h>
#include <cuda_runtime.h>
cufftHandle cufft_handle_fft1b[MAX_FFT1_THREADS];
cuFloatComplex *cuda_in[MAX_FFT1_THREADS];
cuFloatComplex *cuda_out[MAX_FFT1_THREADS];
static int cufft_ready = 0;
static void cufft_close(void)
{
  if (!cufft_ready) return;
  cufftDestroy(cufft_handle_fft1b[0]);
  cudaFree(cuda_in[0]); cudaFree(cuda_out[0]);
  cuda_in[0] = cuda_out[0] = NULL;
  cufft_ready = 0;
}
#endif
static int gpu_batch_n = 4;            /* gpu.fft1_batch_n (buf.c:251): 16 transforms per cuFFT call */
void ref_set_gpu_batch_n(int n) { gpu_batch_n = n; }
int ref_has_cufft(void)
{
#if HAVE_CUFFT == 1
  return 1;
#else
  return 0;
#endif
}
/* debugging aid: LB200_REF_BACKTRACE=1 prints the C frames when the compiled reference crashes */
static void segv_backtrace(int sig)
{
  void *frames[48];
  int n = backtrace(frames, 48);
  backtrace_symbols_fd(frames, n, 2);
  _exit(139);
}
static ref_cfg C;
static sel_state SEL[REF_MAX_SEL];
static float *timf3_all;        /* n_sel regions of 2*timf3_size floats */
static int frame_bytes;
static int inited = 0;
extern int ref_last_lirerr;

/* a few globals live in reference files that cannot be linked */
/* (everything else comes from the reference's own *var.c objects) */

static void *zalloc(size_t n) { void *p = calloc(n + 64, 1); if (!p) { fprintf(stderr, "oom\n"); exit(3);} return p; }

int ref_fft1_size(void) { return fft1_size; }
int ref_fft1_block(void) { return fft1_block; }
int ref_fft1_muln(void) { return fft1_muln; }
int ref_interleave_points(void) { return fft1_interleave_points; }
int ref_new_points(void) { return fft1_new_points; }
int ref_timf1_blockbytes(void) { return timf1_blockbytes; }
int ref_mix1_size(void) { return (int)mix1.size; }
int ref_mix1_interleave(void) { return (int)mix1.interleave_points; }
int ref_mix1_new_points(void) { return (int)mix1.new_points; }
int ref_mix1_crossover(void) { return (int)mix1.crossover_points; }
int ref_timf3_block(void) { return timf3_block; }
int ref_timf3_size(void) { return timf3_size; }
int ref_wg_xpixels(void) { return wg_xpixels; }
int ref_sumsq_bufsize(void) { return fft1_sumsq_bufsize; }
int ref_sumsq_pa(void) { return fft1_sumsq_pa; }
int ref_sumsq_counter(void) { return fft1_sumsq_counter; }
int ref_waterf_ptr(void) { return wg_waterf_ptr; }
int ref_waterf_sum_counter(void) { return wg_waterf_sum_counter; }
int ref_sumsq_recalc(void) { return fft1_sumsq_recalc; }
int ref_sumsq_pwg(void) { return fft1_sumsq_pwg; }
int ref_latest_wg_spectrum(void) { return latest_wg_spectrum; }
int ref_wg_first_point(void) { return wg_first_point; }
int ref_wg_last_point(void) { return wg_last_point; }
int ref_first_fft_bandwidth(void) { return genparm[FIRST_FFT_BANDWIDTH]; }
void ref_set_change_fft1_flag(int v) { change_fft1_flag = v; }      /* wide_graph.c:648 sets it on every redraw */
int ref_waterf_size(void) { return wg_waterf_size; }
int ref_first_point(void) { return fft1_first_point; }
int ref_last_point(void) { return fft1_last_point; }
int ref_lirerr(void) { return ref_last_lirerr; }
float ref_points_per_hz(void) { return fftx_points_per_hz; }
float ref_filtercorr_start(void) { return fft1_filtercorr_start; }
const float *ref_window(void) { return fft1_window; }
const float *ref_filtercorr(void) { return fft1_filtercorr; }
const float *ref_desired(void) { return fft1_desired; }
const float *ref_sumsq(void) { return fft1_sumsq; }
const float *ref_slowsum(void) { return fft1_slowsum; }
const float *ref_corrsum(void) { return fft1_corrsum; }
const float *ref_fft1_power(void) { return fft1_power; }
const float *ref_fft1_xypower(void) { return (const float *)fft1_xypower; }
int ref_max_fft1n(void) { return max_fft1n; }
const float *ref_slowcorr(void) { return fft1_slowcorr; }
const double *ref_slowcorr_tot(void) { return fft1_slowcorr_tot; }
int ref_slowcorr_tot_avgnum(void) { return slowcorr_tot_avgnum; }
const short *ref_waterf(void) { return wg_waterf; }
const float *ref_waterf_yfac(void) { return wg_waterf_yfac; }
const float *ref_waterf_sum(void) { return wg_waterf_sum; }
const float *ref_mix1_fqwin(void) { return mix1_fqwin; }
const float *ref_mix1_window(void) { return mix1.window; }
const float *ref_mix1_cos2win(void) { return mix1.cos2win; }
const float *ref_mix1_sin2win(void) { return mix1.sin2win; }
const float *ref_timf3(int ss) { return timf3_all + (size_t)ss * 2 * timf3_size; }
int ref_timf3_pa(void) { return timf3_pa; }
void ref_sel_state(int ss, float *f4, int *i2)
{
  f4[0] = SEL[ss].phase; f4[1] = SEL[ss].phase_step; f4[2] = SEL[ss].phase_rot; f4[3] = SEL[ss].old_phase;
  i2[0] = SEL[ss].point; i2[1] = SEL[ss].old_point;
}

/* the reference's own table builders, exposed for oracle-port validation */
void ref_make_window(int mo, int sz, int n, float *win) { make_window(mo, sz, n, win); }
void ref_fftback(int size, int n, float *x)
{
  COSIN_TABLE *tab = zalloc(sizeof(COSIN_TABLE) * size);
  unsigned short *perm = zalloc(sizeof(unsigned short) * size * 2);
  init_fft(0, n, size, tab, perm);
  fftback(size, n, x, tab, perm, 0);
  free(tab); free(perm);
}

static void free_all(void)
{
  /* buffers are intentionally leaked between re-inits of this test harness
     except the big ones */
  free(timf1_char); timf1_char = NULL;
  free(fft1_char); fft1_char = NULL;
  free(fft1_sumsq); fft1_sumsq = NULL;
  free(timf3_all); timf3_all = NULL;
  free(wg_waterf); wg_waterf = NULL;
}

/* I/Q mirror-image calibration (fft1.c:3607-3657, 3941-4026): install a foldcorr table and raise
 * CALIQ; NULL switches it off again.  Channel-2 phasing (fft1.c:4064-4080, pol_graph.c:165-173). */
void ref_set_foldcorr(const float *table)
{
  if (table) {
    memcpy(fft1_foldcorr, table, sizeof(float) * twice_rxchan * fft1_size);
    fft1_calibrate_flag |= CALIQ;
  } else {
    fft1_calibrate_flag &= ~CALIQ;
  }
}
void ref_set_ch2_phasing(float c1, float c2) { pg_ch2_c1 = c1; pg_ch2_c2 = c2; }

void ref_set_selfreq(int ss, double hz)
{
  SEL[ss].selfreq = hz;
  SEL[ss].point = -1;         /* wide_graph.c:174 */
}

int ref_init(const ref_cfg *cfg, int timf1_bytes_req, int max_fft1n_req)
{
  int i, j, k, v, mode_row;
  unsigned int ui_, uj, uk;
  float t1;
#ifdef LB200_USE_SHIM
  if (inited) lb200_shim_close();
#endif
  if (inited) free_all();
  if (getenv("LB200_REF_BACKTRACE")) signal(SIGSEGV, segv_backtrace);
  C = *cfg;
  ref_last_lirerr = 0;
  memset(&ui, 0, sizeof(ui));
  memset(genparm, 0, sizeof(int) * (MAX_GENPARM + 2));
  memset(&wg, 0, sizeof(wg));
  ui.rx_input_mode = C.input_mode;
  ui.rx_rf_channels = C.rf_channels;
  ui.rx_ad_channels = (C.input_mode & IQ_DATA) ? 2 * C.rf_channels : C.rf_channels;
  ui.rx_ad_speed = C.ad_speed;
  ui.sample_shift = C.sample_shift;
  ui.network_flag = 0;
  ui.operator_skil = OPERATOR_SKIL_NEWCOMER;
  rx_mode = 0;
  kill_all_flag = 0;
  internal_generator_flag = 0;
  genparm[FIRST_FFT_SINPOW] = C.sinpow;
  genparm[FIRST_FFT_GAIN] = C.fft1_gain;
  genparm[MIX1_BANDWIDTH_REDUCTION_N] = C.mix1_red_n;
  genparm[MIX1_NO_OF_CHANNELS] = 1;      /* harness loops selections itself, see ref_process */
  genparm[SECOND_FFT_ENABLE] = 0;
  genparm[MAX_NO_OF_SPURS] = 0;
  genparm[FIRST_FFT_BANDWIDTH] = 100;
  /* buf.c:149 + fft1var.c:74-79: find the VERNR column that maps to the wanted fft_cntrl row */
  fft1mode = (ui.rx_input_mode & (TWO_CHANNELS + IQ_DATA)) / 2;
  mode_row = fft1mode;
  v = -1;
  for (i = 0; i < MAX_FFT1_VERNR; i++) if (fft1_version[mode_row][i] == C.fft1_version) v = i;
  if (v < 0) { fprintf(stderr, "[ref oracle] version %d not legal for fft1mode %d\n", C.fft1_version, fft1mode); return -1; }
  genparm[FIRST_FFT_VERNR] = v;

  /* ---- sizes, buf.c:165-332 with fft1_n given directly ---- */
  twice_rxchan = 2 * ui.rx_rf_channels;
  sw_onechan = (ui.rx_rf_channels == 1);
  fft1_n = C.fft1_n;
  fft1_size = 1 << fft1_n;
  fft1_block = twice_rxchan * fft1_size;
  fft1_use_gpu = 0;
  fft1_muln = (fft_cntrl[FFT1_CURMODE].real2complex + 1) * fft_cntrl[FFT1_CURMODE].parall_fft;
  if (fft_cntrl[FFT1_CURMODE].gpu == GPU_CUDA) {          /* buf.c:238-258 */
#if HAVE_CUFFT == 1
    fft1_use_gpu = GPU_CUDA;
    gpu_fft1_batch_size = 1 << gpu_batch_n;
    fft1_muln = gpu_fft1_batch_size;
#else
    fprintf(stderr, "[ref oracle] fft_cntrl row %d needs the cuFFT build (_ref/libref_cufft.so)\n", C.fft1_version);
    return -3;
#endif
  } else if (fft_cntrl[FFT1_CURMODE].gpu != 0) return -3;
  fft1_mulblock = fft1_block * fft1_muln;
  j = fft1_size * (fft_cntrl[FFT1_CURMODE].real2complex + 1);
  fft1_permute_size = j; fft1_window_size = j; fft1_costab_size = j / 2;
  if (fft_cntrl[FFT1_CURMODE].permute == 2) { fft1_costab_size *= 2; fft1_permute_size *= 2; fft1_window_size += 16; }
  if (!fft1_use_gpu && fft_cntrl[FFT1_CURMODE].doub == 0 && fft1_permute_size > 0x10000) { fprintf(stderr, "[ref oracle] N too big for float path\n"); return -2; }
  fft1_blockbytes = fft1_block * (int)sizeof(float);
  frame_bytes = 2 * ui.rx_ad_channels;
  if (ui.rx_input_mode & DWORD_INPUT) frame_bytes *= 2;
  /* buf.c:113-136 make_interleave_ratio (stored in a float) */
  if (C.sinpow == 0) fft1_interleave_ratio = 0;
  else if (C.sinpow == 9) fft1_interleave_ratio = 0.625;
  else if (C.sinpow == 8) fft1_interleave_ratio = 0.8;
  else fft1_interleave_ratio = 2 * asin(pow(0.5, 1.0 / C.sinpow)) / PI_L;
  mix1.n = fft1_n - C.mix1_red_n;
  if (mix1.n < 3) mix1.n = 3;
  mix1.size = 1 << mix1.n;
  mix1.interleave_points = fft1_interleave_ratio * mix1.size;
  mix1.interleave_points &= 0xfffffffe;
  fft1_interleave_points = mix1.interleave_points * (fft1_size / mix1.size);
  fft1_new_points = fft1_size - fft1_interleave_points;
  mix1.new_points = mix1.size - mix1.interleave_points;
  timf1_blockbytes = fft1_new_points * frame_bytes;
  if ((ui.rx_input_mode & IQ_DATA) == 0) timf1_blockbytes *= 2;
  timf1_blockbytes *= fft1_muln;                 /* buf.c:616-617 */
  timf1_sampling_speed = ui.rx_ad_speed;
  if ((ui.rx_input_mode & IQ_DATA) == 0) timf1_sampling_speed *= 0.5;   /* buf.c:48-51 */
  fft1_hz_per_point = (float)ui.rx_ad_speed / fft1_size;
  if ((ui.rx_input_mode & IQ_DATA) == 0) fft1_hz_per_point /= 2;
  fftx_points_per_hz = 1 / fft1_hz_per_point;
  timf3_sampling_speed = timf1_sampling_speed / fft1_size * mix1.size;
  timf3_block = twice_rxchan * mix1.new_points;
  timf3_size = 16 * mix1.size * twice_rxchan;            /* pow2, includes the 2C factor (buf.c:645-655) */
  timf3_mask = timf3_size - 1;
  timf3_totsiz = timf3_size;
  yieldflag_wdsp_fft1 = 0; yieldflag_ndsp_mix1 = 0;
  fft1_calibrate_flag = 0;
  fft1_direction = C.direction;
  pg_ch2_c1 = 1; pg_ch2_c2 = 0;
  fft1afc_flag = C.afc ? 1 : 0; no_of_spurs = 0;
  fft1_correlation_flag = (C.correlation == 1 && ui.rx_rf_channels == 2) ? 1 : 0;   /* buf.c:1223-1224 */

  /* ---- rings ---- */
  timf1_bytes = timf1_bytes_req;
  timf1_bytemask = timf1_bytes - 1;
  timf1_char = zalloc(timf1_bytes);
  timf1_short_int = (short int *)timf1_char; timf1_int = (int *)timf1_char; timf1_float = (float *)timf1_char;
  timf1p_pa = timf1p_pb = timf1p_px = 0;
  max_fft1n = max_fft1n_req;
  fft1n_mask = max_fft1n - 1;
  fft1_bytes = max_fft1n * fft1_blockbytes;
  fft1_mask = max_fft1n * fft1_block - 1;
  fft1_char = zalloc(2 * (size_t)fft1_bytes);
  fft1_float = (float *)fft1_char;
  fft1_pa = fft1_pb = fft1_px = 0; fft1_na = fft1_nb = fft1_nx = fft1_nm = 0;
  fft1_tmp_bytes = fft1_blockbytes * (fft_cntrl[FFT1_CURMODE].real2complex + 1) * fft_cntrl[FFT1_CURMODE].parall_fft;
  if (fft_cntrl[FFT1_CURMODE].doub) fft1_tmp_bytes *= 2;
  if (fft1_use_gpu) fft1_tmp_bytes = fft1_blockbytes * gpu_fft1_batch_size;      /* buf.c:794-797 */
  /* buf.c:901-908 lays timf2_tmp (2*fft1_tmp_bytes of scratch) directly behind fftw_tmp.  That
   * matters: when log2(2*fft1_size) is even, fft_real_to_hermitian's stray length-2 butterfly
   * after its first loop (fft0.c:63-65) lands at z[2*size-2], i.e. outside fftw_tmp and inside
   * timf2_tmp -- harmless in Linrad, heap corruption if fftw_tmp is allocated on its own. */
  fftw_tmp = zalloc(3 * (size_t)fft1_tmp_bytes + 64);
  fft1_sumsq_bufsize = 16 * fft1_size; if (C.avg2num + 2 > 16) { k = C.avg2num + 2; make_power_of_two(&k); fft1_sumsq_bufsize = k * fft1_size; }
  fft1_sumsq_mask = fft1_sumsq_bufsize - 1;
  fft1_sumsq = zalloc(sizeof(float) * fft1_sumsq_bufsize);
  fft1_slowsum = zalloc(sizeof(float) * fft1_size);
  if (C.afc_mix) {                                        /* buf.c:1089-1092, 1255-1258, 1598 */
    int nafc = REF_MAX_SEL * max_fft1n;
    mix1_fq_mid = zalloc(sizeof(float) * nafc);
    mix1_fq_start = zalloc(sizeof(float) * nafc);
    mix1_fq_curv = zalloc(sizeof(float) * nafc);
    mix1_fq_slope = zalloc(sizeof(float) * nafc);
    for (i = 0; i < nafc; i++) { mix1_fq_mid[i] = -1; mix1_fq_start[i] = -1; }
    fftxn_mask = fft1n_mask;
    baseband_bw_hz = 0.05f * fft1_hz_per_point * (float)mix1.size;
#ifdef LB200_USE_SHIM
    lb200_shim_afc_tables = afc_tables_via_reference;
#endif
  }
  if (fft1afc_flag > 0) {                                 /* buf.c:935-940 */
    fft1_power = zalloc(sizeof(float) * fft1_size * max_fft1n);
    fft1_xypower = zalloc(sizeof(TWOCHAN_POWER) * fft1_size * max_fft1n);
  }
  if (fft1_correlation_flag == 1) {                       /* buf.c:1229-1232, clear_fft1_correlation fft1.c:5386 */
    fft1_corrsum = zalloc(sizeof(float) * 2 * fft1_sumsq_bufsize);
    fft1_slowcorr = zalloc(sizeof(double) * 2 * fft1_size);
    fft1_slowcorr_tot = zalloc(sizeof(double) * 2 * fft1_size);
    slowcorr_tot_avgnum = 0;
    correlation_reset_flag = fft1corr_reset_flag;
  }
  fft1_sumsq_pa = 0; fft1_sumsq_counter = 0; fft1_sumsq_pwg = 0;
  ag_pa = 0; ag_mask = 255;

  /* ---- tables, buf.c:1403-1456 ---- */
  i = (fft_cntrl[FFT1_CURMODE].permute == 2) ? 2 : 1;
  k = fft_cntrl[FFT1_CURMODE].real2complex ? 2 * fft1_size : fft1_size;
  fft1tab = zalloc(sizeof(COSIN_TABLE) * (fft1_costab_size + 16));
  d_fft1tab = zalloc(sizeof(D_COSIN_TABLE) * (fft1_costab_size + 16));
  fft1_permute = zalloc(sizeof(unsigned short) * (fft1_permute_size + 16));
  fft1_bigpermute = zalloc(sizeof(unsigned int) * (fft1_permute_size + 16));
  fft1_window = zalloc(sizeof(float) * (fft1_window_size + 32));
  d_fft1_window = zalloc(sizeof(double) * (fft1_window_size + 32));
  if (!fft1_use_gpu) make_sincos(i, k, fft1tab);        /* buf.c:1408, 1434: no CPU tables for the GPU rows */
  if (fft1_use_gpu) {
#if HAVE_CUFFT == 1
    /* wcw.c:553-576 */
    int ncu[1];
    ncu[0] = fft1_size;
    cufft_close();
    if (cufftPlanMany(&cufft_handle_fft1b[0], 1, ncu, NULL, 1, fft1_size, NULL, 1, fft1_size, CUFFT_C2C, gpu_fft1_batch_size) != CUFFT_SUCCESS) {
      fprintf(stderr, "[ref oracle] cufftPlanMany failed (lirerr 1461)\n");
      return 1461;
    }
    if (cudaMalloc((void **)&cuda_in[0], (size_t)gpu_fft1_batch_size * fft1_size * sizeof(cuFloatComplex)) != cudaSuccess ||
        cudaMalloc((void **)&cuda_out[0], (size_t)gpu_fft1_batch_size * fft1_size * sizeof(cuFloatComplex)) != cudaSuccess) return 1461;
    cufft_ready = 1;
#endif
  } else if (fft_cntrl[FFT1_CURMODE].doub) {
    make_d_sincos(i, k, d_fft1tab);
    make_bigpermute(fft_cntrl[FFT1_CURMODE].permute, fft_cntrl[FFT1_CURMODE].real2complex ? fft1_n + 1 : fft1_n, k, fft1_bigpermute);
    make_d_window(fft_cntrl[FFT1_CURMODE].window, k, C.sinpow, d_fft1_window);
  } else {
    make_permute(fft_cntrl[FFT1_CURMODE].permute, fft_cntrl[FFT1_CURMODE].real2complex ? fft1_n + 1 : fft1_n, k, fft1_permute);
  }
  make_window(fft_cntrl[FFT1_CURMODE].window, k, C.sinpow, fft1_window);

  /* ---- wide graph range + endpoints (fft1.c:4607) ---- */
  wg.first_xpoint = C.first_xpoint;
  wg.xpoints = C.xpoints;
  wg.fft_avg1num = C.avg1num;
  wg.waterfall_avgnum = C.waterfall_avgnum;
  wg.xpoints_per_pixel = C.xpoints_per_pixel;
  wg.pixels_per_xpoint = C.pixels_per_xpoint;
  wg_fft_avg2num = C.avg2num;
  change_fft1_flag = 0;
  lir_status = 0;
  fft1_filtercorr = zalloc(sizeof(float) * twice_rxchan * fft1_size + 64);
  fft1_desired = zalloc(sizeof(float) * fft1_size);
  fft1_foldcorr = zalloc(sizeof(float) * twice_rxchan * fft1_size);
  clear_fft1_filtercorr();                       /* fft1.c:4673, uncalibrated defaults */
  set_fft1_endpoints();                          /* fft1.c:4607 */

  /* ---- waterfall (wide_graph.c:956-1001, 1374-1389) ---- */
  if (wg.xpoints_per_pixel > 1) wg_xpixels = wg.xpoints / wg.xpoints_per_pixel;
  else if (wg.pixels_per_xpoint > 1) wg_xpixels = wg.xpoints * wg.pixels_per_xpoint;
  else wg_xpixels = wg.xpoints;
  if (wg.xpoints_per_pixel == 1 || wg.pixels_per_xpoint == 1) { if (wg_xpixels + wg.first_xpoint > fft1_size) wg_xpixels = fft1_size - wg.first_xpoint; }
  wg_waterf_size = wg_xpixels * C.wf_lines;
  wg_waterf = zalloc(sizeof(short) * (wg_waterf_size + wg_xpixels + 64));
  for (i = 0; i < wg_waterf_size; i++) wg_waterf[i] = (short int)0x8000;
  wg_waterf_ptr = 0;
  wg_waterf_sum = zalloc(sizeof(float) * (fft1_size + 16));
  for (i = 0; i < fft1_size; i++) wg_waterf_sum[i] = 0.00001F;
  wg_waterf_sum_counter = 0;
  wg_waterf_yfac = zalloc(sizeof(float) * (fft1_size + 16));
  t1 = (float)FFT1_WATERFALL_ZERO / (float)wg.waterfall_avgnum;
  if (wg.xpoints_per_pixel > 1) t1 *= (float)ui.rx_rf_channels;
  for (i = 0; i < fft1_size; i++) {
    if (fft1_desired[i] > 0.3162278) wg_waterf_yfac[i] = t1 / (float)pow(fft1_desired[i], 2.0);
    else wg_waterf_yfac[i] = t1 * 10;
  }
  wg_waterf_yfac[0] = t1; wg_waterf_yfac[fft1_size - 1] = t1;
  audio_dump_flag = 0;

  /* ---- mix1 (buf.c:982,1297-1300; prepare_mixer buf.c:55-111) ---- */
  mix1_fqwin = zalloc(sizeof(float) * (mix1.size / 2 + 16));
  make_window(5, mix1.size, 4, mix1_fqwin);
  mix1.window = zalloc(sizeof(float) * (mix1.size + 16));
  mix1.cos2win = zalloc(sizeof(float) * (mix1.size + 16));
  mix1.sin2win = zalloc(sizeof(float) * (mix1.size + 16));
  mix1.permute = zalloc(sizeof(unsigned short) * (mix1.size + 16));
  mix1.table = zalloc(sizeof(COSIN_TABLE) * (mix1.size + 16));
  if (C.sinpow != 0 && C.sinpow != 2) make_window(3, mix1.size, C.sinpow, mix1.window);
  init_fft(0, mix1.n, mix1.size, mix1.table, mix1.permute);
  mix1.crossover_points = 0;
  if (C.sinpow != 0 && C.sinpow != 2) {
    if (C.sinpow == 9) mix1.crossover_points = mix1.size / 8;
    else if (C.sinpow == 8) mix1.crossover_points = mix1.size / 16;
    else {
      ui_ = mix1.interleave_points / 2;
      t1 = mix1.window[ui_];
      while (mix1.window[ui_] < 30 * t1 && ui_ > 0) { ui_--; mix1.crossover_points++; }
      if (mix1.crossover_points > 0.75 * mix1.new_points) mix1.crossover_points = 0.75 * mix1.new_points;
      if (mix1.crossover_points > mix1.interleave_points / 2) mix1.crossover_points = mix1.interleave_points / 2;
    }
    t1 = 0.25 * PI_L / mix1.crossover_points;
    uj = (mix1.size - mix1.new_points) / 2;
    uk = uj + mix1.crossover_points / 2;
    uj -= mix1.crossover_points / 2;
    for (ui_ = 0; ui_ < mix1.crossover_points; ui_++) {
      mix1.cos2win[ui_] = mix1.window[uk] * pow(cos(t1), 2.0);
      mix1.sin2win[ui_] = mix1.window[uj] * pow(sin(t1), 2.0);
      uk--; uj++;
      t1 += 0.5 * PI_L / mix1.crossover_points;
    }
  }
  fftn_tmp = zalloc(sizeof(float) * (4 * mix1.size * ui.rx_rf_channels + 64));
  timf3_all = zalloc(sizeof(float) * 2 * (size_t)timf3_size * (C.n_sel > 0 ? C.n_sel : 1));
  timf3_float = timf3_all;
  timf3_pa = 0;
  mix1_lowest_fq = (fft1_first_point + 1) * fft1_hz_per_point;      /* wide_graph.c:1336-1341 */
  mix1_highest_fq = (fft1_last_point - 1) * fft1_hz_per_point;
  old_mix1_selfreq = -1;
  for (i = 0; i < REF_MAX_SEL; i++) {
    memset(&SEL[i], 0, sizeof(sel_state));
    SEL[i].selfreq = -1; SEL[i].point = -1;
  }
  inited = 1;
#ifdef LB200_USE_SHIM
  if (lb200_shim_open(1) != 0) return ref_last_lirerr ? ref_last_lirerr : -1;
#endif
  return ref_last_lirerr;
}
int ref_uses_shim(void)
{
#ifdef LB200_USE_SHIM
  return 1;
#else
  return 0;
#endif
}

/* append nblocks*timf1_blockbytes bytes of raw samples to the timf1 ring and
 * run the path once per block.  Outputs (any may be NULL):
 *   fft1_out : nblocks * fft1_block floats  (post-fft1_c contents of fft1_float)
 *   raw_out  : nblocks * fft1_block floats  (fft1_b output BEFORE fft1_c)
 *   timf3_out: nblocks * n_sel * timf3_block floats (the block at timf3_pa after each call)
 */
int ref_process(const void *data, int nblocks, float *fft1_out, float *raw_out, float *timf3_out)
{
  int b, ss, i, sub;
  const char *src = (const char *)data;
  for (b = 0; b < nblocks; b++) {
    {                                              /* the block goes into the ring in at most two pieces */
      size_t first = (size_t)(timf1_bytemask + 1 - timf1p_pa);
      if (first > (size_t)timf1_blockbytes) first = timf1_blockbytes;
      memcpy(timf1_char + timf1p_pa, src + (size_t)b * timf1_blockbytes, first);
      memcpy(timf1_char, src + (size_t)b * timf1_blockbytes + first, (size_t)timf1_blockbytes - first);
    }
    timf1p_pa = (timf1p_pa + timf1_blockbytes) & timf1_bytemask;
    timf1p_pb = timf1p_pa;
    /* wcw.c:1036-1047 */
    HOT_FFT1_B(timf1p_px, &fft1_float[fft1_pa], fftw_tmp, 0);
    timf1p_px = (timf1p_px + timf1_blockbytes) & timf1_bytemask;
    if (raw_out) memcpy(raw_out + (size_t)b * fft1_mulblock, &fft1_float[fft1_pa], sizeof(float) * fft1_mulblock);
    fft1_pa = (fft1_pa + fft1_mulblock) & fft1_mask;
    fft1_na = fft1_pa / fft1_block;
    /* wcw.c:1067-1072 */
    sub = 0;
    while (fft1_na != fft1_nb) {
      int at = fft1_nb * fft1_block;
      size_t tno = (size_t)b * fft1_muln + sub;      /* running transform number */
      HOT_FFT1_C();
      fft1_waterfall();
      if (fft1_out) memcpy(fft1_out + tno * fft1_block, &fft1_float[at], sizeof(float) * fft1_block);
      /* wcw.c:1706-1716, one fft1_mix1_fixed per transform; selections looped here (MAX_MIX1==1) */
      if (C.n_sel > 0) {
        int pa0 = timf3_pa, nx0 = fft1_nx, px0 = fft1_px;
        for (ss = 0; ss < C.n_sel; ss++) {
          timf3_pa = pa0; fft1_nx = nx0; fft1_px = px0;
          timf3_float = timf3_all + (size_t)ss * 2 * timf3_size;
          mix1_selfreq[0] = SEL[ss].selfreq;
          mix1_phase[0] = SEL[ss].phase; mix1_phase_step[0] = SEL[ss].phase_step;
          mix1_phase_rot[0] = SEL[ss].phase_rot; mix1_old_phase[0] = SEL[ss].old_phase;
          mix1_point[0] = SEL[ss].point; mix1_old_point[0] = SEL[ss].old_point;
          if (C.afc_mix) {
            /* what the AFC would have put there: this transform's and the next one's frequency */
            if (mix1_fq_mid[fft1_nx] < 0) mix1_fq_mid[fft1_nx] = (float)afc_track(ss, (long)tno);
            if (mix1_fq_start[fft1_nx] < 0) mix1_fq_start[fft1_nx] = mix1_fq_mid[fft1_nx];
            mix1_fq_mid[(fft1_nx + 1) & fft1n_mask] = (float)afc_track(ss, (long)tno + 1);
            HOT_MIX1_AFC();
          } else {
            HOT_MIX1_FIXED();
          }
          SEL[ss].phase = mix1_phase[0]; SEL[ss].phase_step = mix1_phase_step[0];
          SEL[ss].phase_rot = mix1_phase_rot[0]; SEL[ss].old_phase = mix1_old_phase[0];
          SEL[ss].point = mix1_point[0]; SEL[ss].old_point = mix1_old_point[0];
          if (timf3_out) {
            float *dst = timf3_out + (tno * C.n_sel + ss) * timf3_block;
            for (i = 0; i < timf3_block; i++) dst[i] = timf3_float[(pa0 + i) & timf3_mask];
          }
        }
      } else {
        fft1_nx = (fft1_nx + 1) & fft1n_mask;
        fft1_px = (fft1_px + fft1_block) & fft1_mask;
      }
      sub++;
    }
    if (ref_last_lirerr) return ref_last_lirerr;
  }
  return 0;
}

static double afc_track(int ss, long tno)
{
  return SEL[ss].selfreq + 0.4 * fft1_hz_per_point * sin(0.35 * (double)tno);
}

/* timing leg for bench.py --impl reference / cpu_baseline: same loop, no copies */
int ref_process_timed(const void *data, int nblocks) { return ref_process(data, nblocks, NULL, NULL, NULL); }

/* ---------------------------------------------------------------------------------------------
 * Second-FFT front end: the reference's own make_timf2 (timf2.c:31-208: strong/weak split by
 * liminfo, fft1back_one/two, fft1back_fp_finish) run on the fft1_float blocks the harness holds.
 * Sizes and tables as buf.c does them for the float path (swfloat): buf.c:396-401 (timf2 ring),
 * 962-1036 (buffers), 1313-1325 (inverted window, back scramble, back table), 1499-1508 (initial
 * contents).  genparm[SECOND_FFT_ENABLE] stays 0: the harness calls make_timf2 itself. */
void make_timf2(void);
static int timf2_ready;
int ref_timf2_setup(int att_n, int pow_size)
{
  int i;
  genparm[FIRST_BCKFFT_VERNR] = 0;                       /* fft_cntrl rows 10 / 13: the C back transforms */
  genparm[FIRST_BCKFFT_ATT_N] = att_n;
  swfloat = 1;
  swmmx_fft2 = 0;
  ampinfo_flag = 0;
  yieldflag_timf2_fft1 = 0;
  fft1_split_float = zalloc(sizeof(float) * 4 * ui.rx_rf_channels * fft1_size + 64);
  liminfo = zalloc(sizeof(float) * 2 * fft1_size + 64);
  fft1_back_scramble = zalloc(sizeof(short int) * fft1_size + 64);
  make_permute(fft_cntrl[FFT1_BCKCURMODE].permute, fft1_n, fft1_size, fft1_back_scramble);
  if (fft_cntrl[FFT1_CURMODE].permute == 2 || fft_cntrl[FFT1_CURMODE].real2complex == 1) {
    fft1_backtab = zalloc(sizeof(COSIN_TABLE) * fft1_size / 2 + 64);
    make_sincos(0, fft1_size, fft1_backtab);
  } else {
    fft1_backtab = fft1tab;
  }
  if (genparm[FIRST_FFT_SINPOW] != 0 && genparm[FIRST_FFT_SINPOW] != 2) {
    fft1_inverted_window = zalloc(sizeof(float) * (16 + fft1_size / 2) + 64);
    make_window(3, fft1_size, genparm[FIRST_FFT_SINPOW], fft1_inverted_window);
  }
  timf2pow_size = pow_size;
  timf2pow_mask = timf2pow_size - 1;
  timf2_size = 4 * ui.rx_rf_channels * timf2pow_size;
  timf2_mask = timf2_size - 1;
  timf2_float = zalloc(sizeof(float) * timf2_size + 64);
  timf2_pwr_float = zalloc(sizeof(float) * timf2pow_size + 64);
  for (i = 0; i < timf2pow_size; i++) timf2_pwr_float[i] = 0.5F;
  timf2_input_block = (fft1_size - fft1_interleave_points) * 4 * ui.rx_rf_channels;
  timf2_pa = 0;
  fft1_lowlevel_fraction = .75F;
  /* timf2_tmp needs 4*channels*fft1_size floats; the fft1 scratch behind fftw_tmp is too small for it */
  timf2_tmp = zalloc(sizeof(float) * 4 * ui.rx_rf_channels * fft1_size + 256);
  timf2_ready = 1;
#ifdef LB200_USE_SHIM
  /* the plan is created from Linrad's tables: fft1_inverted_window exists only now */
  lb200_shim_close();
  if (lb200_shim_open(1) != 0) return ref_last_lirerr ? ref_last_lirerr : -1;
#endif
  return 0;
}
void ref_set_liminfo(const float *v) { memcpy(liminfo, v, sizeof(float) * fft1_size); }
/* make_timf2 on `nblocks` consecutive blocks of fft1_float starting at float index px */
int ref_make_timf2(int px, int nblocks)
{
  int b;
  if (!timf2_ready) return -1;
  fft1_px = px & fft1_mask;
  for (b = 0; b < nblocks; b++) HOT_MAKE_TIMF2();
  return ref_last_lirerr;
}
float *ref_timf2_float(void) { return timf2_float; }
float *ref_timf2_pwr_float(void) { return timf2_pwr_float; }
int ref_timf2_size(void) { return timf2_size; }
int ref_timf2_pa(void) { return timf2_pa; }
int ref_timf2_input_block(void) { return timf2_input_block; }
float ref_lowlevel_fraction(void) { return fft1_lowlevel_fraction; }
int ref_lowlevel_points(void) { return fft1_lowlevel_points; }
float *ref_inverted_window(void) { return fft1_inverted_window; }
void ref_set_fft1_block(int block_index, const float *v) { memcpy(&fft1_float[(size_t)block_index * fft1_block], v, sizeof(float) * fft1_block); }



/* ---- third FFT, transform half of make_fft3_all (fft3.c:215-470) ------------------------------------
 * Tables as baseb_graph.c:3679-3680 builds them (init_fft(1,..), make_window(1,..)), buffers as
 * baseb_graph.c:3481-3487.  With THREAD_FFT3 not ACTIVE make_fft3_all returns right after the transforms
 * (fft3.c:470-471), before the baseband-graph power / waterfall half and before it advances timf3_px and
 * fft3_pa, so the harness steps the indices.  MAX_MIX1 == 1: one selection (ss = 0, poffs = 0). */
#ifdef LB200_USE_SHIM
int lb200_shim_fft3_open(void);
void lb200_shim_fft3_transforms(void);
#endif
static float *fft3_test_ring;
static int fft3_new_points_req;
void ref_set_fft3_new_points(int n) { fft3_new_points_req = n; }
int ref_fft3_setup(int n, int sinpow, int ring_floats)
{
  if (!inited) return -1;
  fft3_n = n;
  fft3_size = 1 << n;
  fft3_block = fft3_size * 2 * ui.rx_rf_channels * MAX_MIX1;          /* baseb_graph.c:3409 */
  fft3_totsiz = 4 * fft3_block * 4;
  fft3_mask = fft3_totsiz - 1;
  fft3 = zalloc(sizeof(float) * fft3_totsiz);
  fft3_tmp = zalloc(4 * (size_t)fft3_size * ui.rx_rf_channels * sizeof(float) + 64);
  fft3_tab = zalloc(fft3_size * sizeof(COSIN_TABLE) / 2 + 64);
  fft3_permute = zalloc(fft3_size * sizeof(short int) + 64);
  fft3_window = zalloc(fft3_size * sizeof(float) + 64);
  genparm[THIRD_FFT_SINPOW] = sinpow;
  init_fft(1, fft3_n, fft3_size, fft3_tab, fft3_permute);
  make_window(1, fft3_size, sinpow, fft3_window);
  fft3_new_points = fft3_new_points_req > 0 ? fft3_new_points_req : fft3_size / 2;
  fft3_pa = 0;
  yieldflag_ndsp_fft3 = 0;
  thread_command_flag[THREAD_FFT3] = THRFLAG_IDLE;
  genparm[MIX1_NO_OF_CHANNELS] = 1;
  mix1_selfreq[0] = 1000.0;
  old_mix1_selfreq = mix1_selfreq[0];
  free(fft3_test_ring);
  fft3_test_ring = zalloc(sizeof(float) * (size_t)ring_floats);
  timf3_float = fft3_test_ring;
  timf3_size = ring_floats;
  timf3_mask = ring_floats - 1;
#ifdef LB200_USE_SHIM
  if (lb200_shim_fft3_open() != 0) return ref_last_lirerr ? ref_last_lirerr : -1;
#endif
  return 0;
}
/* one make_fft3_all on the given ring contents with timf3_px = px; out = twice_rxchan*fft3_size floats */
int ref_make_fft3(const float *ring, int px, float *out)
{
  memcpy(fft3_test_ring, ring, sizeof(float) * (size_t)timf3_size);
  timf3_float = fft3_test_ring;
  timf3_px = px;
  fft3_pa = 0;
#ifdef LB200_USE_SHIM
  /* Linrad with the library plugged in: the transform half of make_fft3_all through the shim */
  lb200_shim_fft3_transforms();
#else
  make_fft3_all();
#endif
  memcpy(out, &fft3[fft3_pa], sizeof(float) * twice_rxchan * fft3_size);
  return ref_last_lirerr;
}
float *ref_fft3_window(void) { return fft3_window; }
