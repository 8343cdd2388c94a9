/* linrad_b200.h -- C ABI of liblinrad_b200.so, the B200 (sm_100a) replacement for Linrad's
 * wideband DSP hot path:  timf1 -> (unpack, window) -> fft1 -> fft1_float -> |X|^2 ->
 * fft1_sumsq -> mix1 -> timf3, with the wide-graph consumers of fft1_sumsq (slowsum, waterfall), the input codecs
 * of recorded files, and the front ends of the second and third FFT (make_timf2, the transforms of make_fft3_all).
 *
 * The reference (fventuri/linrad) has no FFI for this path: its boundary is a set of C
 * functions working on global ring buffers.  Every entry point
 * below names the reference function it replaces; argument names are the reference's own
 * global names so that the host-side shim (linrad_b200/host/lb200_shim.c)
 * is a one-line forward per function.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Two flavours of every compute call:
 *   *_dev : all buffers are DEVICE pointers (bulk processing, several streams per GPU)
 *   plain : buffers are HOST pointers exactly as Linrad owns them (buf.c mem() list); the
 *           library stages them through its own device mirrors (H2D/D2H inside the call).
 * All calls on one plan are serialised on the plan's CUDA stream; different plans are
 * independent (one plan per fft1b worker thread, like the cuFFT handles of wcw.c:552-576).
 * A plan is NOT re-entrant: it must not be entered from two host threads at the same time (its
 * mirrors, event pool and counters are unguarded).  A host that drives one plan from two threads
 * -- Linrad's wideband thread calls fft1_b while the narrowband thread calls fft1_mix1_fixed --
 * takes a lock per plan around every entry, as lb200_shim.c does.
 *
 * Return value: 0 on success, else an LB200_ERR_* code.  The shim forwards non-zero codes
 * to lirerr() (lxsys.c:495); texts for errors.lir are listed in INTEGRATION.md.
 */
#ifndef LINRAD_B200_H
#define LINRAD_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LB200_ABI_VERSION 3

/* ui.rx_input_mode bits used by the path (globdef.h:277-279) */
#define LB200_DWORD_INPUT 1
#define LB200_TWO_CHANNELS 2
#define LB200_IQ_DATA 4
#define LB200_FLOAT_INPUT 64      /* globdef.h:283 FLOAT_INPUT; with IQ_DATA|DWORD_INPUT: frames of floats [re, im] (x channels).
                                      This is how the third FFT's transforms (make_fft3_all, fft3.c:215-470) run: a second
                                      plan whose timf1 ring is Linrad's timf3_float and whose fft1_float ring is fft3 */

/* error codes (new lirerr numbers; 3100-3119 are unused in errors.lir) */
#define LB200_OK 0
#define LB200_ERR_NO_DEVICE 3100      /* no CUDA device / driver */
#define LB200_ERR_CUDA 3101           /* a CUDA runtime call or kernel launch failed */
#define LB200_ERR_BAD_CONFIG 3102     /* inconsistent sizes in lb200_config */
#define LB200_ERR_UNSUPPORTED 3103    /* size / mode outside what the kernels cover */
#define LB200_ERR_BAD_ARG 3104        /* null pointer, misaligned offset, ring too small */
#define LB200_ERR_MIX1_RANGE_LOW 1211 /* same codes set_mix1_phases raises (mix1.c:787-796) */
#define LB200_ERR_MIX1_RANGE_HIGH 1212

#define LB200_MAX_MIX1 64             /* selections per plan (reference MAX_MIX1 is 1, globdef.h:169) */

typedef struct lb200_plan lb200_plan; /* opaque */

/* Everything the path reads from Linrad's setup (sizes from get_wideband_sizes buf.c:139,
 * tables from get_buffers buf.c:1297-1456).  Tables are HOST pointers copied at create. */
typedef struct lb200_config {
  int abi_version;              /* LB200_ABI_VERSION */
  int device;                   /* CUDA device ordinal */
  /* --- input format ------------------------------------------------------------- */
  int rx_input_mode;            /* ui.rx_input_mode (DWORD_INPUT|TWO_CHANNELS|IQ_DATA) */
  int rx_rf_channels;           /* ui.rx_rf_channels: 1 or 2 */
  int sample_shift;             /* ui.sample_shift (I/Q skew in samples, IQ input only: < 0 takes Q
                                   from |shift| frames earlier, > 0 takes I from shift frames
                                   earlier, fft1.c:770-790, 2041-2247) */
  /* --- fft1 geometry -------------------------------------------------------------- */
  int fft1_n;                   /* fft1_n;  fft1_size = 1 << fft1_n (bins) */
  int fft1_interleave_points;   /* fft1_interleave_points (buf.c:303,327) */
  int fft1_direction;           /* fft1_direction: +1 or -1 (fft1.c:3660-3680) */
  int fft1_first_point;         /* fft1_first_point (fft1.c:4607-4651) */
  int fft1_last_point;          /* fft1_last_point */
  /* --- tables ----------------------------------------------------------------------- */
  const float *fft1_window;     /* NATURAL-order window w[0..fft1_size) (2*fft1_size for real
                                   input); NULL when genparm[FIRST_FFT_SINPOW]==0.  Use
                                   lb200_window_to_natural() to convert a make_window() table. */
  const float *fft1_filtercorr; /* twice_rxchan*fft1_size floats, layout of fft1.c:4691-4692 */
  const float *fft1_foldcorr;   /* twice_rxchan*fft1_size floats, layout of fft1_filtercorr, or NULL
                                   when (fft1_calibrate_flag & CALIQ) == 0: the I/Q mirror-image
                                   correction of fft1_b (fft1.c:3607-3657, 3941-4026), IQ input only */
  /* --- power spectra --------------------------------------------------------------- */
  int fft_avg1num;              /* wg.fft_avg1num */
  /* --- mix1 ------------------------------------------------------------------------- */
  int mix1_n;                   /* mix1.n ; mix1.size = 1 << mix1_n ; 0 = mix1 not used */
  int mix1_interleave_points;   /* mix1.interleave_points */
  int mix1_crossover_points;    /* mix1.crossover_points */
  const float *mix1_fqwin;      /* mix1.size/2+1 floats, make_window(5,...) buf.c:1297 */
  const float *mix1_window;     /* mix1.size floats or NULL (sinpow 0 or 2) */
  const float *mix1_cos2win;    /* crossover_points floats or NULL */
  const float *mix1_sin2win;    /* crossover_points floats or NULL */
  float fftx_points_per_hz;     /* fftx_points_per_hz */
  float mix1_lowest_fq;         /* mix1_lowest_fq  (wide_graph.c:1336-1341) */
  float mix1_highest_fq;        /* mix1_highest_fq */
  /* --- bulk-mode capacity ------------------------------------------------------------ */
  int max_batch;                /* largest nblocks per call (sizes staging buffers) */
  /* --- channel-2 phasing (ABI 2) ---------------------------------------------------- */
  float pg_ch2_c1, pg_ch2_c2;   /* pol_graph.c:165-173; (1, 0) = off.  Two-channel IQ input only:
                                   ch2 *= (c1 - i c2) on bins [first_sym_point, N - first_sym_point)
                                   (fft1.c:4064-4080) */
  /* --- second FFT front end (ABI 3) -------------------------------------------------- */
  const float *fft1_inverted_window; /* fft1_inverted_window, make_window(3,...) buf.c:1313: fft1_size/2+1 floats, or
                                   NULL (no second FFT, or FIRST_FFT_SINPOW 0 / 2 where it is not used) */
} lb200_config;

/* Ring-buffer descriptor: base pointer + power-of-two size (the reference's xxx_mask+1).
 * Device rings (the *_dev entry points) must be 16-byte aligned (Linrad's own buffers are,
 * buf.c:2105; the host-ring entry points stage through aligned mirrors). */
typedef struct lb200_ring {
  void *base;
  size_t size;                  /* bytes for timf1, floats for fft1_float/fft1_sumsq/timf3_float */
} lb200_ring;

/* One call of the fused fft1_b + fft1_c over `nblocks` consecutive transforms. */
typedef struct lb200_fft1_args {
  lb200_ring timf1;             /* timf1_char, timf1_bytemask+1 */
  uint32_t timf1p_ref;          /* byte offset of the first NEW sample of the first transform
                                   (fft1_b's timf1p_ref == timf1p_px, wcw.c:1036) */
  int nblocks;
  lb200_ring fft1_float;        /* fft1_float, fft1_mask+1 (floats) */
  uint32_t fft1_pa;             /* float index where the first transform goes (wcw.c:1036) */
  int apply_filtercorr;         /* 0: raw fft1_b output; 1: fft1_c's filtercorr applied too */
  lb200_ring fft1_sumsq;        /* fft1_sumsq ring or base==NULL to skip power */
  uint32_t fft1_sumsq_pa;       /* fft1_sumsq_pa (floats) */
  int fft1_sumsq_counter;       /* fft1_sumsq_counter on entry (fft1.c:4115,4507) */
  float *power_rows;            /* optional: nblocks rows of fft1_size floats = per-transform
                                   |z|^2 (lets a host fft1_c keep the reference's own
                                   accumulation order); NULL to skip */
  int flags;                    /* LB200_FFT1_* (ABI 2) */
  /* fft1_correlation_flag == 1 (two RF channels, fft1.c:4146-4152, 4189-4195): the cross spectrum
   * 2*z1*conj(z2) of the filter-corrected channels, summed like fft1_sumsq (ABI 2) */
  lb200_ring fft1_corrsum;      /* fft1_corrsum ring, 2*fft1_sumsq.size floats ([re,im] at 2*(fft1_sumsq
                                   index)); needs fft1_sumsq; base==NULL to skip */
  float *corr_rows;             /* with power_rows: nblocks rows of 2*fft1_size floats = per-transform
                                   cross spectrum; NULL to skip */
  float *xypower_rows;          /* with power_rows, two RF channels: nblocks rows of fft1_size
                                   TWOCHAN_POWER {x2, y2, im_xy, re_xy} (globdef.h:1371-1376) = what
                                   fft1_c stores in fft1_xypower when fft1afc_flag > 0
                                   (fft1.c:4349-4368); NULL to skip.  One channel: fft1_power is
                                   power_rows itself. */
  /* Several input rings in one call (ABI 3; lb200_fft1_dev, plans with LB200_FLOAT_INPUT, apply_filtercorr = 0):
   * ring r = 0..no_of_rings-1 is read at timf1.base + r*timf1_ring_stride (bytes, same size and same timf1p_ref)
   * and written at fft1_pa + r*fft1_pa_stride (floats, same fft1_float ring).  This is the loop over ss of
   * make_fft3_all: timf3_float selections 2*timf3_size floats apart, fft3 blocks mm*fft3_size apart
   * (fft3.c:232-233).  0 or 1 = one ring. */
  int no_of_rings;
  size_t timf1_ring_stride;
  uint32_t fft1_pa_stride;
} lb200_fft1_args;
/* lb200_fft1 (host rings) only: leave fft1_float in the plan's device mirror of the ring and do
 * not write the host ring.  For set-ups where nothing on the host reads the spectrum (second FFT
 * and AFC off, no fft1 network output: its only consumer is mix1, and lb200_mix1 finds the
 * transforms in the mirror).  Saves 8*C*N bytes of device-to-host traffic per transform. */
#define LB200_FFT1_SPECTRUM_STAYS_ON_DEVICE 1

/* per-selection mix1 state, the reference's per-ss globals (selvar.c:213-220) */
typedef struct lb200_mix1_state {
  double mix1_selfreq;          /* mix1_selfreq[ss], <0 = not selected */
  float mix1_phase;             /* mix1_phase[ss] */
  float mix1_phase_step;        /* mix1_phase_step[ss] */
  float mix1_phase_rot;         /* mix1_phase_rot[ss] */
  float mix1_old_phase;         /* mix1_old_phase[ss] */
  int mix1_point;               /* mix1_point[ss], -1 right after selection */
  int mix1_old_point;           /* mix1_old_point[ss] */
} lb200_mix1_state;

typedef struct lb200_mix1_args {
  lb200_ring fft1_float;        /* source spectra (post fft1_c) */
  uint32_t fft1_px;             /* float index of the first transform to mix (mix1.c:1019) */
  int nblocks;
  int no_of_channels;           /* genparm[MIX1_NO_OF_CHANNELS] */
  lb200_mix1_state *state;      /* HOST array[no_of_channels], updated like set_mix1_phases/do_mix1 do */
  lb200_ring timf3_float;       /* per-selection ring: selection ss lives at base + ss*2*size
                                   (mix1.c:109 poffs=ss*timf3_size; see DESIGN.md on ss>0) */
  uint32_t timf3_pa;            /* timf3_pa on entry */
} lb200_mix1_args;

/* ---------------------------------------------------------------------------------------- */
int lb200_create(const lb200_config *cfg, lb200_plan **plan); /* ~ cufftPlanMany site wcw.c:552-576 */
void lb200_destroy(lb200_plan *plan);                         /* ~ wcw.c:1174-1184 */
const char *lb200_strerror(int code);
int lb200_abi_version(void);
/* plan-owned CUDA stream (cudaStream_t) so callers can order their own copies against it */
void *lb200_stream(lb200_plan *plan);
int lb200_synchronize(lb200_plan *plan);
/* counters for bench.py: kernels launched / bytes copied by this plan since creation */
uint64_t lb200_launch_count(const lb200_plan *plan);
uint64_t lb200_h2d_bytes(const lb200_plan *plan);
uint64_t lb200_d2h_bytes(const lb200_plan *plan);

/* fft1_b (fft1.c:3302) fused with fft1_c's arithmetic (fft1.c:4085-4200): device buffers */
int lb200_fft1_dev(lb200_plan *plan, const lb200_fft1_args *a);
/* same on Linrad's host buffers: copies the needed timf1 span H2D, runs, copies the new
 * fft1_float blocks (and sumsq rows / power rows) back D2H */
int lb200_fft1(lb200_plan *plan, const lb200_fft1_args *a);

/* fft1_mix1_fixed (mix1.c:995) incl. set_mix1_phases (mix1.c:781) and do_mix1 (mix1.c:55) */
int lb200_mix1_dev(lb200_plan *plan, const lb200_mix1_args *a);
int lb200_mix1(lb200_plan *plan, const lb200_mix1_args *a);

/* ---- second FFT front end: make_timf2 (timf2.c:31-208), float path (swfloat) ----------------------
 * Every transform of fft1_float is split by liminfo (0: weak, else strong; fft1_update_liminfo keeps
 * that table on the host, sellim.c:738), both halves are transformed back to the time domain
 * (fft1back_one / fft1back_two) and laid into the timf2 ring by fft1back_fp_finish (timf2.c:970):
 * sample layout [weak re, im, strong re, im] (two channels: [w1 re, im, w2 re, im, s1 re, im, s2 re, im]),
 * |weak|^2 into timf2_pwr_float.  With the sin^2 window the first half of a transform is ADDED onto
 * the second half of its predecessor, which the previous call left at timf2_pa.  fft1 sizes 2^7..2^14. */
typedef struct lb200_timf2_args {
  lb200_ring fft1_float;        /* source spectra (post fft1_c) */
  uint32_t fft1_px;             /* float index of the first transform (timf2.c:37) */
  int nblocks;
  const float *liminfo;         /* fft1_size floats: HOST for lb200_make_timf2, DEVICE for lb200_make_timf2_dev */
  lb200_ring timf2_float;       /* timf2_float, timf2_mask+1 floats */
  float *timf2_pwr_float;       /* timf2_float.size / (4*rx_rf_channels) floats */
  uint32_t timf2_pa;            /* timf2_pa on entry; the caller advances it by nblocks*timf2_input_block */
  int first_bckfft_att_n;       /* genparm[FIRST_BCKFFT_ATT_N] */
  int *fft1_lowlevel_points;    /* HOST, optional: fft1_lowlevel_points (the same for every transform of the call;
                                   host-buffer variant only) */
} lb200_timf2_args;
int lb200_make_timf2_dev(lb200_plan *plan, const lb200_timf2_args *a);
int lb200_make_timf2(lb200_plan *plan, const lb200_timf2_args *a);

/* ---- wide-graph consumers of fft1_sumsq -------------------------------------------------- */
/* what update_fft1_slowsum / fft1_waterfall read from the wide-graph setup (WG_PARMS
 * globdef.h:907-926, screendef.h; wg_first/last_point fft1.c:4612-4614) */
typedef struct lb200_wg_config {
  int wg_fft_avg2num;           /* wg_fft_avg2num */
  int waterfall_avgnum;         /* wg.waterfall_avgnum */
  int first_xpoint;             /* wg.first_xpoint */
  int xpoints;                  /* wg.xpoints */
  int wg_first_point;           /* wg_first_point */
  int wg_last_point;            /* wg_last_point */
  int wg_xpixels;               /* wg_xpixels */
  int xpoints_per_pixel;        /* wg.xpoints_per_pixel */
  int pixels_per_xpoint;        /* wg.pixels_per_xpoint */
  int first_fft_bandwidth;      /* genparm[FIRST_FFT_BANDWIDTH] (selects fresh_recalc, fft1.c:4547) */
} lb200_wg_config;

/* the scalar globals the two functions advance */
typedef struct lb200_wg_state {
  int fft1_sumsq_pwg;           /* fft1_sumsq_pwg (floats) */
  int fft1_sumsq_recalc;        /* fft1_sumsq_recalc */
  int change_fft1_flag;         /* change_fft1_flag */
  int wg_waterf_sum_counter;    /* wg_waterf_sum_counter */
  int wg_waterf_ptr;            /* wg_waterf_ptr */
  int latest_wg_spectrum;       /* latest_wg_spectrum */
} lb200_wg_state;

typedef struct lb200_wg_args {
  lb200_ring fft1_sumsq;        /* fft1_sumsq, fft1_sumsq_mask+1 (floats) */
  uint32_t fft1_sumsq_pa;       /* the FIRST newly completed row (fft1_sumsq_pa when fft1_c
                                   called update_fft1_slowsum for it, fft1.c:4540) */
  int nrows;                    /* newly completed rows, in ring order */
  float *fft1_slowsum;          /* fft1_size floats */
  float *wg_waterf_sum;         /* fft1_size floats */
  const float *wg_waterf_yfac;  /* fft1_size floats (make_wg_yfac wide_graph.c:956-1001) */
  short *wg_waterf;             /* wg_waterf_size shorts (+ pixels_per_xpoint+1 slack when
                                   interpolating, like the reference's own allocation) */
  int wg_waterf_size;
  lb200_wg_state *state;        /* HOST, updated */
} lb200_wg_args;

/* update_fft1_slowsum (fft1.c:4526) incl. new_fft1_averages (wide_graph.c:1003), once per row */
int lb200_update_fft1_slowsum_dev(lb200_plan *plan, const lb200_wg_config *wg, const lb200_wg_args *a);
int lb200_update_fft1_slowsum(lb200_plan *plan, const lb200_wg_config *wg, const lb200_wg_args *a);
/* fft1_waterfall (fft1.c:115) incl. update_wg_waterf (fft1.c:104): drains the rows from
 * state->fft1_sumsq_pwg up to fft1_sumsq_pa + nrows*fft1_size */
int lb200_fft1_waterfall_dev(lb200_plan *plan, const lb200_wg_config *wg, const lb200_wg_args *a);
int lb200_fft1_waterfall(lb200_plan *plan, const lb200_wg_config *wg, const lb200_wg_args *a);

/* ---- input codecs of file playback (bit-exact) -------------------------------------------- */
/* expand_rawdat (getiq64.s:158-220): 18-bit packed -> int32; out_bytes multiple of 16, the
 * packed input holds 9*out_bytes/16 bytes */
int lb200_expand_rawdat_dev(lb200_plan *plan, const void *packed, void *out, size_t out_bytes);
int lb200_expand_rawdat(lb200_plan *plan, const void *packed, void *out, size_t out_bytes);
/* 24-bit PCM -> left-justified int32 (rxin.c:1603-1614); nsamples multiple of 4 */
int lb200_widen_24bit_dev(lb200_plan *plan, const void *pcm24, void *out, size_t nsamples);
int lb200_widen_24bit(lb200_plan *plan, const void *pcm24, void *out, size_t nsamples);
/* 8-bit unsigned PCM -> int16: (byte << 8) - 32640 (rxin.c:1573-1583); nsamples multiple of 4 */
int lb200_widen_8bit_dev(lb200_plan *plan, const void *pcm8, void *out, size_t nsamples);
int lb200_widen_8bit(lb200_plan *plan, const void *pcm8, void *out, size_t nsamples);
/* 32-bit float wav samples -> int32: 0x7fffffff * z truncated, out of range / NaN -> 0x80000000 like
 * the reference's x86 conversion (rxin.c:1624-1634); nsamples multiple of 4 */
int lb200_float_to_int32_dev(lb200_plan *plan, const void *f32, void *out, size_t nsamples);
int lb200_float_to_int32(lb200_plan *plan, const void *f32, void *out, size_t nsamples);

/* ---- Linrad .raw recordings (modesub.c:656-733 reader, 1519-1640 writer) -------------------
 * Little-endian, no padding:
 *   int32 first          >= 0: this IS rx_input_mode (old format; time 0, centre 0, direction +1)
 *                        <  0: REMEMBER_* tag (-1 unknown, -2 nothing, -3 Perseus, -4 SDR-14; the
 *                              last two are followed by int32 chunk_size + chunk), then
 *                              double diskread_time, double passband_center,
 *                              int32 passband_direction (+1/-1 -> fft1_direction), int32 rx_input_mode
 *   int32 rx_rf_channels (2 sets TWO_CHANNELS in rx_input_mode), int32 rx_ad_channels (1..4, equal to
 *   or twice rx_rf_channels), int32 rx_ad_speed, uint8 save_init_flag (bit 0: filter calibration
 *   follows, bit 1: I/Q calibration follows; fft1.c:5203, 4780)
 * Payload: blocks of timf1 frames; int16 verbatim, DWORD_INPUT recordings 18-bit packed
 * (18*block_bytes/32 bytes per block, buf.c:599, rxin.c:1643) for lb200_expand_rawdat. */
#define LB200_REMEMBER_UNKNOWN (-1)
#define LB200_REMEMBER_NOTHING (-2)
#define LB200_REMEMBER_PERSEUS (-3)
#define LB200_REMEMBER_SDR14 (-4)
typedef struct lb200_raw_header {
  int remember_tag;             /* remember_proprietery_chunk[0]; NOTHING for old-format files */
  int chunk_size;               /* remember_proprietery_chunk[1] (Perseus / SDR-14), else 0 */
  uint64_t chunk_offset;        /* file offset of that chunk, 0 when there is none */
  double diskread_time;
  double passband_center;       /* fg.passband_center */
  int passband_direction;       /* fg.passband_direction == fft1_direction */
  int rx_input_mode;            /* ui.rx_input_mode, TWO_CHANNELS or-ed in like the reader does */
  int rx_rf_channels;
  int rx_ad_channels;
  int rx_ad_speed;
  int save_init_flag;
  uint64_t payload_offset;      /* first byte behind save_init_flag */
} lb200_raw_header;
/* open_savefile's header logic on a memory image of the file head.  LB200_OK, or
 * LB200_ERR_BAD_ARG where the reference says "File corrupted" (short file, unknown tag,
 * direction not +-1, rx_input_mode >= MODEPARM_MAX, bad channel counts). */
int lb200_raw_header_parse(const void *bytes, size_t nbytes, lb200_raw_header *out);
/* bytes one block of `block_bytes` timf1 bytes occupies in the file (save_rw_bytes, buf.c:599) */
size_t lb200_raw_block_bytes(const lb200_raw_header *h, size_t block_bytes);

/* ---- multi-GPU: sum of the averaged power spectra ------------------------
 * The hot path shards over independent receiver streams / time-block ranges, one process per
 * GPU, with no data-path collective; what one Linrad instance's wide graph needs from all of
 * them is the SUM of their fft1_sumsq rows (the reference sums rows of one stream in fft1_c,
 * fft1.c:4507-4520; it has no multi-device path).  Every rank deposits its rows in the root's
 * mailbox with the copy engine over NVLink (IPC-mapped peer memory), the root adds them with
 * one small kernel: no collective shares the SMs with the persistent fft1 kernels.
 *   create -> export (64-byte cudaIpcMemHandle_t) -> caller all-gathers the handles ->
 *   connect -> per round: push on every rank, sum on the root.
 * push/sum are queued on a side stream owned by the reducer and ordered against the plan's
 * stream by events; nothing here synchronises with the host. */
typedef struct lb200_reduce lb200_reduce; /* opaque */
#define LB200_IPC_HANDLE_BYTES 64
int lb200_reduce_create(lb200_plan *plan, int rank, int world, size_t floats, int depth, lb200_reduce **out);
int lb200_reduce_export(lb200_reduce *r, void *handle64);
int lb200_reduce_connect(lb200_reduce *r, const void *handles /* world x 64 bytes, by rank */, int root);
/* every rank: its rows (device pointer, `floats` of create) for the next round */
int lb200_reduce_push(lb200_reduce *r, const float *rows);
/* plan's stream waits until the last pushed rows have been copied out (call before rewriting them) */
int lb200_reduce_rows_released(lb200_reduce *r);
/* root: out[i] = sum over ranks (in rank order) of the next round's rows */
int lb200_reduce_sum(lb200_reduce *r, float *out);
/* plan's stream waits for the last sum */
int lb200_reduce_result_ready(lb200_reduce *r);
int lb200_reduce_synchronize(lb200_reduce *r);
void lb200_reduce_destroy(lb200_reduce *r);

/* Host helpers shared by the shim and the tests (pure integer / scalar logic) */
/* set_mix1_phases (mix1.c:781-861) for one selection and one transform; returns 0 or 1211/1212 */
int lb200_set_mix1_phases(const lb200_config *cfg, lb200_mix1_state *s, float fq);
/* value of mix1_phase[ss] after do_mix1's running sum of `count` float additions of `rot`
 * (mix1.c:146-153,172-186), computed without iterating when the exponent does not change */
float lb200_phase_advance(float phase, float rot, int count);
/* convert a reference window table (make_window mo=1 interleaved / mo=4 natural / mo=2 half,
 * fft0.c:812-921) to natural order */
void lb200_window_to_natural(int mo, int size, const float *win, float *natural);

#ifdef __cplusplus
}
#endif
#endif
