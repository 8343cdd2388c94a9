/* lb200_shim.c -- the Linrad-side binding of liblinrad_b200.so.
 *
 * Linrad has no FFI on this path; its boundary is a handful of C functions working on global
 * ring buffers (fft1def.h:359-364,375).  This file is what a Linrad maintainer
 * adds to the source tree (next to cuda.c) and lists in Makefile.in; it is compiled against
 * Linrad's own headers and forwards, one call each, to the C ABI of include/linrad_b200.h.
 * See INTEGRATION.md for the call-site edits in wcw.c.  Nothing else of Linrad changes: the
 * buffers stay where buf.c put them, the ring indices are advanced by the same code as before.
 *
 *   lb200_shim_open()          at the cufftPlanMany site, wcw.c:552-576 (one plan per fft1b thread)
 *   lb200_shim_close()         at wcw.c:1174-1184
 *   lb200_shim_fft1_b()        in place of fft1_b()           (wcw.c:1036, do_fft1b wcw.c:500)
 *   lb200_shim_fft1_c()        in place of fft1_c()           (wcw.c:338,366,423,1069,1098)
 *   lb200_shim_mix1_fixed()    in place of fft1_mix1_fixed()  (wcw.c:1700,1712)
 *   lb200_shim_make_timf2()    in place of make_timf2()       (timf2.c:31; second FFT enabled, float path)
 *   lb200_shim_fft3_transforms() in place of the transform half of make_fft3_all() (fft3.c:215-470)
 *
 * fft1_b and fft1_c stay two calls on two threads, as in Linrad, but the arithmetic of both
 * runs in ONE kernel: lb200_shim_fft1_b asks the library for the filter-corrected spectrum and
 * the per-transform power row; lb200_shim_fft1_c only folds that row into fft1_sumsq in the
 * reference's own order and does fft1_c's index bookkeeping (fft1.c:4507-4523).
 */
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include "globdef.h"
#include "uidef.h"
#include "fft1def.h"
#include "fft3def.h"
#include "seldef.h"
#include "screendef.h"
#include "thrdef.h"
#include "linrad_b200.h"

#define LB200_SHIM_MAX_THREADS MAX_FFT1_THREADS      /* thrdef.h:106 */

static lb200_plan *shim_plan[LB200_SHIM_MAX_THREADS];
/* A plan is not re-entrant (linrad_b200.h), and Linrad drives plan 0 from two threads: fft1_b on the
 * wideband thread (wcw.c:1036) and fft1_mix1_fixed/_afc on the narrowband thread (wcw.c:1700,1712).
 * They share the plan on purpose -- mix1 finds the spectra fft1_b left in the plan's device mirror --
 * so every entry takes the plan's lock. */
static pthread_mutex_t shim_lock[LB200_SHIM_MAX_THREADS];
static int shim_locks_ready;
static float *shim_power;          /* max_fft1n rows of fft1_size floats, indexed like fft1_float blocks */
static float *shim_window;         /* natural-order copy of fft1_window */
static float *shim_corr;           /* fft1_correlation_flag == 1: rows of 2*fft1_size floats, like shim_power */
static float *shim_xy;             /* fft1afc_flag > 0, two channels: rows of fft1_size TWOCHAN_POWER */

static void shim_fail(int code)
{
/* LB200_ERR_* are new lirerr numbers (texts for errors.lir in INTEGRATION.md); 1211/1212 are */
/* the codes set_mix1_phases raises itself (mix1.c:787-796). */
lirerr(code);
}

int lb200_shim_open(int no_of_threads)
{
int i, k, mo;
lb200_config c;
if(shim_locks_ready == 0)
  {
  for(i=0; i<LB200_SHIM_MAX_THREADS; i++)pthread_mutex_init(&shim_lock[i],NULL);
  shim_locks_ready=1;
  }
/* What the library reproduces: the fft1 versions that take their N frames fft1_interleave_points */
/* before timf1p_ref (fft1.c:700, 425; fft1_re.c:44).  The "twin"/"quad" versions that run two */
/* transforms side by side start I samples earlier (fft1.c:2980, 1047: fft_cntrl[].parall_fft == 2), */
/* and of the real-input versions only the split-radix one (window mode 2, fft1_re.c) is covered. */
if(fft_cntrl[FFT1_CURMODE].parall_fft != 1){shim_fail(LB200_ERR_UNSUPPORTED); return -1;}
if( (ui.rx_input_mode&IQ_DATA) == 0 && fft_cntrl[FFT1_CURMODE].window != 2)
  {
  shim_fail(LB200_ERR_UNSUPPORTED);
  return -1;
  }
/* fft1_correlation_flag >= 2 is the double precision mixer of mix1.c:275-452 */
if(fft1_correlation_flag >= 2){shim_fail(LB200_ERR_UNSUPPORTED); return -1;}
memset(&c,0,sizeof(c));
c.abi_version=LB200_ABI_VERSION;
c.device=0;
c.rx_input_mode=ui.rx_input_mode;
c.rx_rf_channels=ui.rx_rf_channels;
c.sample_shift=ui.sample_shift;
c.fft1_n=fft1_n;
c.fft1_interleave_points=fft1_interleave_points;
c.fft1_direction=fft1_direction;
c.fft1_first_point=fft1_first_point;
c.fft1_last_point=fft1_last_point;
/* make_window() layouts differ per fft version (fft0.c:812-921): hand over natural order */
k=fft1_size;
if( (ui.rx_input_mode&IQ_DATA) == 0)k*=2;
shim_window=NULL;
if(genparm[FIRST_FFT_SINPOW] != 0)
  {
  shim_window=malloc((size_t)k*sizeof(float));
  if(shim_window == NULL){shim_fail(LB200_ERR_BAD_CONFIG); return -1;}
  mo=fft_cntrl[FFT1_CURMODE].window;
  lb200_window_to_natural(mo, fft1_size, fft1_window, shim_window);
  }
c.fft1_window=shim_window;
c.fft1_filtercorr=fft1_filtercorr;
c.fft1_foldcorr=(fft1_calibrate_flag&CALIQ) ? fft1_foldcorr : NULL;
c.fft_avg1num=wg.fft_avg1num;
c.mix1_n=(int)mix1.n;
c.mix1_interleave_points=(int)mix1.interleave_points;
c.mix1_crossover_points=(int)mix1.crossover_points;
c.mix1_fqwin=mix1_fqwin;
c.mix1_window=mix1.window;
c.mix1_cos2win=mix1.cos2win;
c.mix1_sin2win=mix1.sin2win;
c.fftx_points_per_hz=fftx_points_per_hz;
c.mix1_lowest_fq=mix1_lowest_fq;
c.mix1_highest_fq=mix1_highest_fq;
c.max_batch=1;
c.pg_ch2_c1=pg_ch2_c1;
c.pg_ch2_c2=pg_ch2_c2;
/* second FFT (float path): the inverted window of fft1back_fp_finish, buf.c:1005,1313 */
c.fft1_inverted_window=NULL;
if(genparm[FIRST_FFT_SINPOW] != 0 && genparm[FIRST_FFT_SINPOW] != 2)c.fft1_inverted_window=fft1_inverted_window;
shim_power=malloc((size_t)(fft1n_mask+1)*(size_t)fft1_size*sizeof(float));
if(shim_power == NULL){shim_fail(LB200_ERR_BAD_CONFIG); return -1;}
shim_xy=NULL;
if(fft1afc_flag > 0 && ui.rx_rf_channels == 2)
  {
  shim_xy=malloc((size_t)(fft1n_mask+1)*4*(size_t)fft1_size*sizeof(float));
  if(shim_xy == NULL){shim_fail(LB200_ERR_BAD_CONFIG); return -1;}
  }
shim_corr=NULL;
if(fft1_correlation_flag == 1)
  {
  shim_corr=malloc((size_t)(fft1n_mask+1)*2*(size_t)fft1_size*sizeof(float));
  if(shim_corr == NULL){shim_fail(LB200_ERR_BAD_CONFIG); return -1;}
  }
for(i=0; i<no_of_threads && i<LB200_SHIM_MAX_THREADS; i++)
  {
  k=lb200_create(&c,&shim_plan[i]);
  if(k != LB200_OK){shim_fail(k); return -1;}
  }
return 0;
}

void lb200_shim_close(void)
{
int i;
for(i=0; i<LB200_SHIM_MAX_THREADS; i++)
  {
  if(shim_plan[i])lb200_destroy(shim_plan[i]);
  shim_plan[i]=NULL;
  }
free(shim_power); shim_power=NULL;
free(shim_window); shim_window=NULL;
free(shim_corr); shim_corr=NULL;
free(shim_xy); shim_xy=NULL;
}

/* Same signature as fft1_b (fft1def.h:363).  `tmp` is not needed.  `out` is &fft1_float[fft1_pa]. */
void lb200_shim_fft1_b(int timf1p_ref, float *out, float *tmp, int gpu_handle_number)
{
lb200_fft1_args a;
int rc;
(void)tmp;
memset(&a,0,sizeof(a));
a.timf1.base=timf1_char;
a.timf1.size=(size_t)timf1_bytemask+1;
a.timf1p_ref=(uint32_t)timf1p_ref;
a.nblocks=1;
a.fft1_float.base=fft1_float;
a.fft1_float.size=(size_t)fft1_mask+1;
a.fft1_pa=(uint32_t)(out-fft1_float);
a.apply_filtercorr=1;
a.power_rows=&shim_power[(size_t)((out-fft1_float)/fft1_block)*(size_t)fft1_size];
if(fft1_correlation_flag == 1)
  a.corr_rows=&shim_corr[(size_t)((out-fft1_float)/fft1_block)*2*(size_t)fft1_size];
if(shim_xy != NULL)
  a.xypower_rows=&shim_xy[(size_t)((out-fft1_float)/fft1_block)*4*(size_t)fft1_size];
pthread_mutex_lock(&shim_lock[gpu_handle_number]);
rc=lb200_fft1(shim_plan[gpu_handle_number],&a);
pthread_mutex_unlock(&shim_lock[gpu_handle_number]);
if(rc != LB200_OK)shim_fail(rc);
}

/* fft1_c (fft1.c:4085): the filtercorr multiply and |z|^2 were done on the GPU; what is */
/* left is the accumulation order of fft1.c:4115-4200 and the bookkeeping of fft1.c:4507-4523. */
void lb200_shim_fft1_c(void)
{
int ia;
float *sum, *pwr;
sum=&fft1_sumsq[fft1_sumsq_pa];
pwr=&shim_power[(size_t)fft1_nb*(size_t)fft1_size];
if(fft1afc_flag > 0)
  {
/* fft1.c:4203-4426: AFC runs from fft1, so fft1_c also leaves the per-transform powers behind */
/* (fft1_power, or fft1_xypower for two channels).  Spur elimination works on fft1_float on */
/* the host between the filter and the power step and is not part of the library. */
/* The spur search of fft1.c:4426-4505 (spursearch_powersum / spursearch_xysum) is not done here */
/* either, so with it enabled no spur would ever be found: refuse instead of silently disabling it. */
  if(no_of_spurs > 0 || genparm[MAX_NO_OF_SPURS] > 0){shim_fail(LB200_ERR_UNSUPPORTED); return;}
  ffts_na=fft1_nb;
  ffts_nm=fft1_nm;
  if(ui.rx_rf_channels == 1)
    {
    float *fp;
    fp=&fft1_power[(size_t)fft1_nb*(size_t)fft1_size];
    for(ia=fft1_first_point; ia <= fft1_last_point; ia++)fp[ia]=pwr[ia];
    }
  else
    {
    TWOCHAN_POWER *pxy, *src;
    pxy=&fft1_xypower[(size_t)fft1_nb*(size_t)fft1_size];
    src=(TWOCHAN_POWER*)&shim_xy[(size_t)fft1_nb*4*(size_t)fft1_size];
    for(ia=fft1_first_point; ia <= fft1_last_point; ia++)
      {
      pxy[ia]=src[ia];
      pwr[ia]=pxy[ia].x2+pxy[ia].y2;         /* fft1.c:4365 */
      }
    }
  }
if(fft1_sumsq_counter == 0)
  {
  for(ia=fft1_first_point; ia <= fft1_last_point; ia++)sum[ia]=pwr[ia];
  }
else
  {
  for(ia=fft1_first_point; ia <= fft1_last_point; ia++)sum[ia]+=pwr[ia];
  }
if(fft1_correlation_flag == 1)
  {
/* fft1.c:4146-4152 / 4189-4195 */
  float *cor, *cs;
  cor=&shim_corr[(size_t)fft1_nb*2*(size_t)fft1_size];
  cs=&fft1_corrsum[2*fft1_sumsq_pa];
  if(fft1_sumsq_counter == 0)
    {
    for(ia=2*fft1_first_point; ia <= 2*fft1_last_point+1; ia++)cs[ia]=cor[ia];
    }
  else
    {
    for(ia=2*fft1_first_point; ia <= 2*fft1_last_point+1; ia++)cs[ia]+=cor[ia];
    }
  }
fft1_sumsq_counter++;
if(fft1_sumsq_counter >= wg.fft_avg1num)
  {
  fft1_sumsq_counter=0;
  update_fft1_slowsum();
  if(genparm[SECOND_FFT_ENABLE] != 0)fft1_liminfo_cnt++;
  fft1_sumsq_pa=(fft1_sumsq_pa+fft1_size)&fft1_sumsq_mask;
  }
fft1_nb=(fft1_nb+1)&fft1n_mask;
fft1_pb=fft1_nb*fft1_block;
if(genparm[SECOND_FFT_ENABLE] == 0)ag_pa=(ag_pa+1)&ag_mask;
}

/* fft1_mix1_fixed (mix1.c:995): set_mix1_phases + gather + do_mix1 for every selection. */
void lb200_shim_mix1_fixed(void)
{
lb200_mix1_args a;
lb200_mix1_state st[MAX_MIX1];
int ss, rc, k;
k=genparm[MIX1_NO_OF_CHANNELS];
for(ss=0; ss<k; ss++)
  {
  st[ss].mix1_selfreq=mix1_selfreq[ss];
  st[ss].mix1_phase=mix1_phase[ss];
  st[ss].mix1_phase_step=mix1_phase_step[ss];
  st[ss].mix1_phase_rot=mix1_phase_rot[ss];
  st[ss].mix1_old_phase=mix1_old_phase[ss];
  st[ss].mix1_point=mix1_point[ss];
  st[ss].mix1_old_point=mix1_old_point[ss];
  }
memset(&a,0,sizeof(a));
a.fft1_float.base=fft1_float;
a.fft1_float.size=(size_t)fft1_mask+1;
a.fft1_px=(uint32_t)fft1_px;
a.nblocks=1;
a.no_of_channels=k;
a.state=st;
a.timf3_float.base=timf3_float;
a.timf3_float.size=(size_t)timf3_size;
a.timf3_pa=(uint32_t)timf3_pa;
pthread_mutex_lock(&shim_lock[0]);
rc=lb200_mix1(shim_plan[0],&a);
pthread_mutex_unlock(&shim_lock[0]);
if(rc != LB200_OK){shim_fail(rc); return;}
for(ss=0; ss<k; ss++)
  {
  mix1_phase[ss]=st[ss].mix1_phase;
  mix1_phase_step[ss]=st[ss].mix1_phase_step;
  mix1_phase_rot[ss]=st[ss].mix1_phase_rot;
  mix1_old_phase[ss]=st[ss].mix1_old_phase;
  mix1_point[ss]=st[ss].mix1_point;
  mix1_old_point[ss]=st[ss].mix1_old_point;
  }
timf3_pa=(timf3_pa+timf3_block)&timf3_mask;         /* mix1.c:1038-1040 */
fft1_nx=(fft1_nx+1)&fft1n_mask;
fft1_px=(fft1_px+fft1_block)&fft1_mask;
}

/* fft1_mix1_afc (mix1.c:1044-1096): the mixer frequency of every transform comes from the AFC
 * track mix1_fq_mid[] instead of mix1_selfreq[]; the arithmetic is fft1_mix1_fixed's.  What
 * do_mix1_afc (mix1.c:648-768) does besides calling do_mix1 -- stepping mix1_fq_slope / _curv /
 * _start and bending future mix1_fq_mid entries -- is host logic on Linrad's AFC tables.  It must
 * be reachable without the CPU mixer: split do_mix1_afc in mix1.c at its last line into
 *     void mix1_afc_tables(int ss)   (everything up to, not including, do_mix1(ss,t2-t1))
 * and point lb200_shim_afc_tables at it. */
void (*lb200_shim_afc_tables)(int ss);

void lb200_shim_mix1_afc(void)
{
lb200_mix1_args a;
lb200_mix1_state st[MAX_MIX1];
int ss, rc, k;
if(lb200_shim_afc_tables == NULL){shim_fail(LB200_ERR_UNSUPPORTED); return;}
k=genparm[MIX1_NO_OF_CHANNELS];
for(ss=0; ss<k; ss++)
  {
  st[ss].mix1_selfreq=-1;
  if(mix1_selfreq[ss] >= 0)st[ss].mix1_selfreq=mix1_fq_mid[ss*max_fft1n+fft1_nx];     /* mix1.c:1059 */
  st[ss].mix1_phase=mix1_phase[ss];
  st[ss].mix1_phase_step=mix1_phase_step[ss];
  st[ss].mix1_phase_rot=mix1_phase_rot[ss];
  st[ss].mix1_old_phase=mix1_old_phase[ss];
  st[ss].mix1_point=mix1_point[ss];
  st[ss].mix1_old_point=mix1_old_point[ss];
  }
memset(&a,0,sizeof(a));
a.fft1_float.base=fft1_float;
a.fft1_float.size=(size_t)fft1_mask+1;
a.fft1_px=(uint32_t)fft1_px;
a.nblocks=1;
a.no_of_channels=k;
a.state=st;
a.timf3_float.base=timf3_float;
a.timf3_float.size=(size_t)timf3_size;
a.timf3_pa=(uint32_t)timf3_pa;
pthread_mutex_lock(&shim_lock[0]);
rc=lb200_mix1(shim_plan[0],&a);
pthread_mutex_unlock(&shim_lock[0]);
if(rc != LB200_OK){shim_fail(rc); return;}
for(ss=0; ss<k; ss++)
  {
  mix1_phase[ss]=st[ss].mix1_phase;
  mix1_phase_step[ss]=st[ss].mix1_phase_step;
  mix1_phase_rot[ss]=st[ss].mix1_phase_rot;
  mix1_old_phase[ss]=st[ss].mix1_old_phase;
  mix1_point[ss]=st[ss].mix1_point;
  mix1_old_point[ss]=st[ss].mix1_old_point;
  if(mix1_selfreq[ss] >= 0)lb200_shim_afc_tables(ss);
  }
timf3_pa=(timf3_pa+timf3_block)&timf3_mask;         /* mix1.c:1093-1095 */
fft1_nx=(fft1_nx+1)&fft1n_mask;
fft1_px=(fft1_px+fft1_block)&fft1_mask;
}

/* make_timf2 (timf2.c:31-208), float path: strong/weak split by liminfo, back transform and */
/* fft1back_fp_finish on the GPU; the index bookkeeping of timf2.c:117-118, 205-207 stays here. */
void lb200_shim_make_timf2(void)
{
lb200_timf2_args a;
int rc;
if(!swfloat){shim_fail(LB200_ERR_UNSUPPORTED); return;}   /* the short-int / MMX back transform is not reproduced */
memset(&a,0,sizeof(a));
a.fft1_float.base=fft1_float;
a.fft1_float.size=(size_t)fft1_mask+1;
a.fft1_px=(uint32_t)fft1_px;
a.nblocks=1;
a.liminfo=liminfo;
a.timf2_float.base=timf2_float;
a.timf2_float.size=(size_t)timf2_mask+1;
a.timf2_pwr_float=timf2_pwr_float;
a.timf2_pa=(uint32_t)timf2_pa;
a.first_bckfft_att_n=genparm[FIRST_BCKFFT_ATT_N];
a.fft1_lowlevel_points=&fft1_lowlevel_points;
pthread_mutex_lock(&shim_lock[0]);
rc=lb200_make_timf2(shim_plan[0],&a);
pthread_mutex_unlock(&shim_lock[0]);
if(rc != LB200_OK){shim_fail(rc); return;}
fft1_px=(fft1_px+fft1_block)&fft1_mask;
fft1_nx=(fft1_nx+1)&fft1n_mask;
fft1_lowlevel_fraction=0.02*(49*fft1_lowlevel_fraction+fft1_lowlevel_points/
             ((float)(fft1_last_point-fft1_first_point)));
timf2_pa=(timf2_pa+timf2_input_block)&timf2_mask;
}



/* ---- third FFT: the transforms of make_fft3_all (fft3.c:215-470) -------------------------------------
 * A second plan whose input ring is timf3_float (float IQ frames, LB200_FLOAT_INPUT) and whose output ring
 * is fft3.  Call lb200_shim_fft3_open() where baseb_graph.c:3679-3680 has just built fft3_tab / fft3_window
 * (init_basebmem), lb200_shim_fft3_close() where it frees them.  make_fft3_all is split in two in Linrad:
 *   void make_fft3_all(void){ if(fft1_use_gpu == GPU_LB200)lb200_shim_fft3_transforms(); else <the loop over ss>;
 *                             <fft3.c:471 on: powers, slowsum, waterfall, index bookkeeping, unchanged> }
 */
static lb200_plan *shim_fft3_plan;
static float *shim_fft3_window;
void lb200_shim_fft3_close(void)
{
if(shim_fft3_plan)lb200_destroy(shim_fft3_plan);
shim_fft3_plan=NULL;
free(shim_fft3_window); shim_fft3_window=NULL;
}
int lb200_shim_fft3_open(void)
{
lb200_config c;
int rc;
lb200_shim_fft3_close();
if(fft1_correlation_flag > 1){shim_fail(LB200_ERR_UNSUPPORTED); return -1;}   /* the double precision branch, fft3.c:281-421 */
memset(&c,0,sizeof(c));
c.abi_version=LB200_ABI_VERSION;
c.device=0;
c.rx_input_mode=IQ_DATA|DWORD_INPUT|LB200_FLOAT_INPUT;
if(ui.rx_rf_channels == 2)c.rx_input_mode|=TWO_CHANNELS;
c.rx_rf_channels=ui.rx_rf_channels;
c.fft1_n=fft3_n;
c.fft1_interleave_points=fft3_size-fft3_new_points;
c.fft1_direction=1;
c.fft1_first_point=0;
c.fft1_last_point=fft3_size-1;
shim_fft3_window=malloc((size_t)fft3_size*sizeof(float));
if(shim_fft3_window == NULL){shim_fail(LB200_ERR_BAD_CONFIG); return -1;}
lb200_window_to_natural(1, fft3_size, fft3_window, shim_fft3_window);      /* make_window(1,...), baseb_graph.c:3680 */
c.fft1_window=shim_fft3_window;
c.fft_avg1num=1;
c.max_batch=1;
c.pg_ch2_c1=1;
rc=lb200_create(&c,&shim_fft3_plan);
if(rc != LB200_OK){shim_fft3_plan=NULL; shim_fail(rc); return -1;}
return 0;
}
void lb200_shim_fft3_transforms(void)
{
lb200_fft1_args a;
int ss, rc, mm;
mm=twice_rxchan;
for(ss=0; ss<genparm[MIX1_NO_OF_CHANNELS]; ss++)
  {
  if(mix1_selfreq[ss] < 0)continue;
  memset(&a,0,sizeof(a));
  a.timf1.base=&timf3_float[ss*mm*timf3_size];                 /* poffs, fft3.c:232 */
  a.timf1.size=((size_t)timf3_mask+1)*sizeof(float);
  /* lb200_fft1 reads the fft3_size frames that begin fft1_interleave_points frames before the reference */
  a.timf1p_ref=(uint32_t)(((size_t)timf3_px*sizeof(float)+
                 (size_t)(fft3_size-fft3_new_points)*mm*sizeof(float))&(a.timf1.size-1));
  a.nblocks=1;
  a.fft1_float.base=fft3;
  a.fft1_float.size=(size_t)fft3_mask+1;
  a.fft1_pa=(uint32_t)((fft3_pa+ss*mm*fft3_size)&fft3_mask);   /* z, fft3.c:233 */
  a.apply_filtercorr=0;
  rc=lb200_fft1(shim_fft3_plan,&a);
  if(rc != LB200_OK){shim_fail(rc); return;}
  }
}

/* layout guard for the harness: the argument structures this object was compiled against */
int lb200_shim_sizeof_args(int which)
{
switch(which)
  {
  case 0: return (int)sizeof(lb200_config);
  case 1: return (int)sizeof(lb200_fft1_args);
  case 2: return (int)sizeof(lb200_mix1_args);
  case 3: return (int)sizeof(lb200_timf2_args);
  }
return -1;
}
