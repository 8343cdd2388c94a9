// lb200_wg.cu -- C-ABI entry points for the wide-graph consumers (wide_graph.cuh) and the
// input codecs (codecs.cuh).
#include <cstdio>
#include <cstring>
#include <vector>
#include "plan.h"
#include "wide_graph.cuh"
#include "codecs.cuh"

using namespace lb;

static bool wg_pow2(size_t v) { return v && !(v & (v - 1)); }

static int wg_fill(lb200_plan* plan, const lb200_wg_config* c, const lb200_wg_args* a, WgK& k)
{
  if (!plan || !c || !a || !a->state || !a->fft1_sumsq.base) return LB200_ERR_BAD_ARG;
  if (!wg_pow2(a->fft1_sumsq.size) || a->nrows < 0) return LB200_ERR_BAD_ARG;
  if (c->wg_fft_avg2num < 1 || c->waterfall_avgnum < 1 || c->wg_xpixels < 0) return LB200_ERR_BAD_CONFIG;
  if ((size_t)(c->wg_fft_avg2num + 1) * plan->N > a->fft1_sumsq.size) return LB200_ERR_BAD_ARG;
  memset(&k, 0, sizeof(k));
  k.sumsq = (const float*)a->fft1_sumsq.base;
  k.sumsq_mask = (uint32_t)(a->fft1_sumsq.size - 1);
  k.pa0 = a->fft1_sumsq_pa;
  k.nrows = a->nrows;
  k.N = plan->N;
  k.slowsum = a->fft1_slowsum;
  k.first_point = plan->cfg.fft1_first_point;
  k.last_point = plan->cfg.fft1_last_point;
  k.wg_first_point = c->wg_first_point;
  k.wg_last_point = c->wg_last_point;
  k.avg2num = c->wg_fft_avg2num;
  k.xpoints = c->xpoints;
  // fft1.c:4547-4567
  if (c->first_fft_bandwidth > 200) k.fresh_recalc = c->wg_fft_avg2num < 100 ? 4 : 8;
  else k.fresh_recalc = 2;
  k.recalc0 = a->state->fft1_sumsq_recalc;
  k.change_flag0 = a->state->change_fft1_flag;
  // the window recurrence stepped up to 24 rows before the end of the call (integers only)
  k.tail_row = 0;
  k.tail_recalc = k.recalc0;
  if (a->nrows > 48) {
    int recalc = k.recalc0;
    bool change = k.change_flag0 != 0;
    const int step = k.xpoints / k.fresh_recalc;
    const int t0 = a->nrows - 24;
    for (int r = 0; r < t0; r++) {
      if (change) { change = false; continue; }
      if (recalc == k.last_point) recalc = k.first_point;
      recalc += step;
      if (recalc > k.last_point) recalc = k.last_point;
    }
    k.tail_row = t0;
    k.tail_recalc = recalc;
  }
  k.wsum = a->wg_waterf_sum;
  k.yfac = a->wg_waterf_yfac;
  k.waterf = a->wg_waterf;
  k.waterf_size = a->wg_waterf_size;
  k.waterf_ptr0 = a->state->wg_waterf_ptr;
  k.counter0 = a->state->wg_waterf_sum_counter;
  k.avg1num = plan->cfg.fft_avg1num;
  k.waterfall_avgnum = c->waterfall_avgnum;
  k.first_xpoint = c->first_xpoint;
  k.xpixels = c->wg_xpixels;
  k.xpp = c->xpoints_per_pixel;
  k.ppx = c->pixels_per_xpoint;
  k.pwg0 = (uint32_t)a->state->fft1_sumsq_pwg;
  const uint32_t pa_end = (a->fft1_sumsq_pa + (uint32_t)a->nrows * (uint32_t)plan->N) & k.sumsq_mask;
  k.wrows = (int)(((pa_end - k.pwg0) & k.sumsq_mask) / (uint32_t)plan->N);     // fft1.c:119
  return 0;
}

// scalar state machines, repeated on the host so the caller gets the reference's globals back
static void wg_advance_slowsum(const WgK& k, lb200_wg_state* s)
{
  for (int r = 0; r < k.nrows; r++) {
    s->latest_wg_spectrum++;                          // wide_graph.c:1006
    if (s->change_fft1_flag) { s->change_fft1_flag = 0; continue; }
    if (s->fft1_sumsq_recalc == k.last_point) s->fft1_sumsq_recalc = k.first_point;
    s->fft1_sumsq_recalc += k.xpoints / k.fresh_recalc;
    if (s->fft1_sumsq_recalc > k.last_point) s->fft1_sumsq_recalc = k.last_point;
  }
}
static int wg_advance_waterfall(const WgK& k, lb200_wg_state* s, std::vector<int>* lines)
{
  int nlines = 0;
  for (int r = 0; r < k.wrows; r++) {
    s->fft1_sumsq_pwg = (int)(((uint32_t)s->fft1_sumsq_pwg + (uint32_t)k.N) & k.sumsq_mask);
    s->wg_waterf_sum_counter += k.avg1num;
    if (s->wg_waterf_sum_counter >= k.waterfall_avgnum) {
      if (lines) lines->push_back(s->wg_waterf_ptr);
      s->wg_waterf_ptr -= k.xpixels;
      if (s->wg_waterf_ptr < 0) s->wg_waterf_ptr += k.waterf_size;
      s->wg_waterf_sum_counter = 0;
      nlines++;
    }
  }
  return nlines;
}

static int wg_mode(const WgK& k, int* nunits)
{
  const int nb = k.wg_last_point - k.first_xpoint + 1;
  if (k.xpp == 1 || k.ppx == 1) { *nunits = k.N; return 0; }                     // fft1.c:136
  if (k.xpp == 0) {                                                                // fft1.c:149
    const int m = k.xpixels - k.ppx;
    const int groups = m > 0 ? (m + k.ppx - 1) / k.ppx : 0;
    int n = groups + 2;
    if (nb > n) n = nb;
    if (n > k.N - k.first_xpoint) n = k.N - k.first_xpoint;
    *nunits = n;
    return 2;
  }
  int n = (nb + k.xpp - 1) / k.xpp;
  if (k.xpixels > n) n = k.xpixels;
  *nunits = n;
  return 1;
}

extern "C" int lb200_update_fft1_slowsum_dev(lb200_plan* plan, const lb200_wg_config* c, const lb200_wg_args* a)
{
  WgK k;
  int rc = wg_fill(plan, c, a, k);
  if (rc) return rc;
  if (!a->fft1_slowsum) return LB200_ERR_BAD_ARG;
  if (a->nrows == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  slowsum_kernel<<<(plan->N + 255) / 256, 256, 0, plan->stream>>>(k);
  LB_CUDA(cudaGetLastError());
  plan->launches++;
  wg_advance_slowsum(k, a->state);
  return LB200_OK;
}

extern "C" int lb200_fft1_waterfall_dev(lb200_plan* plan, const lb200_wg_config* c, const lb200_wg_args* a)
{
  WgK k;
  int rc = wg_fill(plan, c, a, k);
  if (rc) return rc;
  if (!a->wg_waterf_sum || !a->wg_waterf_yfac || !a->wg_waterf || a->wg_waterf_size < c->wg_xpixels) return LB200_ERR_BAD_ARG;
  // the ring is a whole number of lines (wg_waterf_size = wg_xpixels*wg_waterf_lines, wide_graph.c:1374)
  if (c->wg_xpixels > 0 && a->wg_waterf_size % c->wg_xpixels) return LB200_ERR_BAD_ARG;
  if (k.wrows == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  int nunits = 0;
  const int mode = wg_mode(k, &nunits);
  if (nunits > 0) {
    // launch 1: the complete lines of the call; launch 2: the rows behind the last line -> wg_waterf_sum
    const WgLines lg = wg_lines(k.counter0, k.avg1num, k.waterfall_avgnum, k.wrows);
    const int gx = (nunits + 255) / 256;
    bool fused = false;
    if (lg.nl > 0 && mode == 2) {
      // interpolation: the tail of a line reaches into the first pixels of the line written before it
      // (fft1.c:190-205), so the lines have to land in time order: one launch each
      for (int L = 0; L < lg.nl; L++) {
        waterfall_kernel<<<gx, 256, 0, plan->stream>>>(k, mode, nunits, 0, L, 1);
        LB_CUDA(cudaGetLastError());
        plan->launches++;
      }
    } else if (lg.nl > 0) {
      // lines that land on the same place of the waterfall ring: only the last one survives the sequential walk,
      // the earlier ones are not computed
      int first = 0;
      const int ring_lines = k.waterf_size / k.xpixels;
      if (lg.nl > ring_lines) first = lg.nl - ring_lines;
      int gy = 262144 / nunits;                            // enough threads for the GPU, else one per unit
      if (gy < 1) gy = 1;
      if (gy > lg.nl - first) gy = lg.nl - first;
      if (gy > 32768) gy = 32768;
      const int lps = (lg.nl - first + gy - 1) / gy;
      fused = mode == 0 && gy == 1;                        // one thread per bin: lines and hand-back in one walk
      waterfall_kernel<<<dim3(gx, gy), 256, 0, plan->stream>>>(k, mode, nunits, fused ? 2 : 0, first, lps);
      LB_CUDA(cudaGetLastError());
      plan->launches++;
    }
    if (!fused) {
      waterfall_kernel<<<gx, 256, 0, plan->stream>>>(k, mode, nunits, 1, 0, 0);
      LB_CUDA(cudaGetLastError());
      plan->launches++;
    }
  }
  wg_advance_waterfall(k, a->state, nullptr);
  return LB200_OK;
}

// ---- host-buffer variants: stage what the kernels touch through plan-owned device mirrors
static int wg_mirror(lb200_plan* plan, HostMirror& m, size_t bytes)
{
  if (m.d && m.bytes >= bytes) return 0;
  if (m.d) cudaFree(m.d);
  m.d = nullptr;
  LB_CUDA(cudaMalloc(&m.d, bytes));
  m.bytes = bytes;
  return 0;
}
static int wg_copy(lb200_plan* plan, void* dst, const void* src, size_t bytes, bool h2d)
{
  LB_CUDA(cudaMemcpyAsync(dst, src, bytes, h2d ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, plan->stream));
  if (h2d) plan->h2d += bytes; else plan->d2h += bytes;
  return 0;
}
// rows [pa - back*N, pa + nrows*N) of the host ring -> the same places of the device mirror
static int wg_stage_rows(lb200_plan* plan, const lb200_wg_args* a, uint32_t first_row_pa, int rows)
{
  const size_t size = a->fft1_sumsq.size;
  int rc;
  if ((rc = wg_mirror(plan, plan->m_wg_sumsq, size * 4))) return rc;
  for (int r = 0; r < rows; r++) {
    const size_t off = ((size_t)first_row_pa + (size_t)r * plan->N) & (size - 1);
    if ((rc = wg_copy(plan, (float*)plan->m_wg_sumsq.d + off, (const float*)a->fft1_sumsq.base + off, (size_t)plan->N * 4, true))) return rc;
  }
  return 0;
}

extern "C" int lb200_update_fft1_slowsum(lb200_plan* plan, const lb200_wg_config* c, const lb200_wg_args* a)
{
  if (!plan || !c || !a || !a->fft1_sumsq.base || !a->fft1_slowsum) return LB200_ERR_BAD_ARG;
  if (!wg_pow2(a->fft1_sumsq.size)) return LB200_ERR_BAD_ARG;
  if (a->nrows <= 0) return LB200_OK;
  cudaSetDevice(plan->device);
  int rc;
  const size_t size = a->fft1_sumsq.size;
  const int back = c->wg_fft_avg2num;                       // oldest row read: pa - avg2num*N
  int rows = back + a->nrows;
  if ((size_t)rows * plan->N > size) rows = (int)(size / plan->N);
  if ((rc = wg_stage_rows(plan, a, (uint32_t)((a->fft1_sumsq_pa + size - (size_t)back * plan->N) & (size - 1)), rows))) return rc;
  if ((rc = wg_mirror(plan, plan->m_wg_slowsum, (size_t)plan->N * 4))) return rc;
  if ((rc = wg_copy(plan, plan->m_wg_slowsum.d, a->fft1_slowsum, (size_t)plan->N * 4, true))) return rc;
  lb200_wg_args d = *a;
  d.fft1_sumsq.base = plan->m_wg_sumsq.d;
  d.fft1_slowsum = (float*)plan->m_wg_slowsum.d;
  if ((rc = lb200_update_fft1_slowsum_dev(plan, c, &d))) return rc;
  if ((rc = wg_copy(plan, a->fft1_slowsum, plan->m_wg_slowsum.d, (size_t)plan->N * 4, false))) return rc;
  LB_CUDA(cudaStreamSynchronize(plan->stream));
  return LB200_OK;
}

extern "C" int lb200_fft1_waterfall(lb200_plan* plan, const lb200_wg_config* c, const lb200_wg_args* a)
{
  WgK k;
  int rc = wg_fill(plan, c, a, k);
  if (rc) return rc;
  if (!a->wg_waterf_sum || !a->wg_waterf_yfac || !a->wg_waterf || a->wg_waterf_size < c->wg_xpixels) return LB200_ERR_BAD_ARG;
  // the ring is a whole number of lines (wg_waterf_size = wg_xpixels*wg_waterf_lines, wide_graph.c:1374)
  if (c->wg_xpixels > 0 && a->wg_waterf_size % c->wg_xpixels) return LB200_ERR_BAD_ARG;
  if (k.wrows == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  if ((rc = wg_stage_rows(plan, a, k.pwg0, k.wrows))) return rc;
  const size_t nb = (size_t)plan->N * 4;
  const size_t slack = (size_t)(c->pixels_per_xpoint > 1 ? c->pixels_per_xpoint + 1 : 0);
  const size_t wbytes = ((size_t)a->wg_waterf_size + slack) * sizeof(short);
  if ((rc = wg_mirror(plan, plan->m_wg_wsum, nb))) return rc;
  if ((rc = wg_mirror(plan, plan->m_wg_yfac, nb))) return rc;
  if ((rc = wg_mirror(plan, plan->m_wg_waterf, wbytes))) return rc;
  if ((rc = wg_copy(plan, plan->m_wg_wsum.d, a->wg_waterf_sum, nb, true))) return rc;
  if ((rc = wg_copy(plan, plan->m_wg_yfac.d, a->wg_waterf_yfac, nb, true))) return rc;
  // the lines this call will write (plus the interpolation overrun): the mirror gets the host's
  // current pixels first so that pixels the reference leaves untouched come back untouched
  lb200_wg_state st0 = *a->state;
  std::vector<int> lines;
  wg_advance_waterfall(k, &st0, &lines);
  auto line_copy = [&](bool h2d) -> int {
    for (int p0 : lines) {
      size_t n = (size_t)c->wg_xpixels + slack;
      if ((size_t)p0 + n > (size_t)a->wg_waterf_size + slack) n = (size_t)a->wg_waterf_size + slack - p0;
      int r2;
      if (h2d) r2 = wg_copy(plan, (short*)plan->m_wg_waterf.d + p0, a->wg_waterf + p0, n * sizeof(short), true);
      else r2 = wg_copy(plan, a->wg_waterf + p0, (short*)plan->m_wg_waterf.d + p0, n * sizeof(short), false);
      if (r2) return r2;
    }
    return 0;
  };
  if ((rc = line_copy(true))) return rc;
  lb200_wg_args d = *a;
  d.fft1_sumsq.base = plan->m_wg_sumsq.d;
  d.wg_waterf_sum = (float*)plan->m_wg_wsum.d;
  d.wg_waterf_yfac = (const float*)plan->m_wg_yfac.d;
  d.wg_waterf = (short*)plan->m_wg_waterf.d;
  if ((rc = lb200_fft1_waterfall_dev(plan, c, &d))) return rc;
  if ((rc = wg_copy(plan, a->wg_waterf_sum, plan->m_wg_wsum.d, nb, false))) return rc;
  if ((rc = line_copy(false))) return rc;
  LB_CUDA(cudaStreamSynchronize(plan->stream));
  return LB200_OK;
}

// ------------------------------------------------------------------------------------------
// input codecs
extern "C" int lb200_expand_rawdat_dev(lb200_plan* plan, const void* packed, void* out, size_t out_bytes)
{
  if (!plan || !packed || !out || (out_bytes & 15)) return LB200_ERR_BAD_ARG;
  if (out_bytes == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  const size_t groups = out_bytes / 16;
  int grid = (int)((groups + 255) / 256);
  if (grid > plan->sm_count * 16) grid = plan->sm_count * 16;
  expand_rawdat_kernel<<<grid, 256, 0, plan->stream>>>((const uint8_t*)packed, (int4*)out, groups);
  LB_CUDA(cudaGetLastError());
  plan->launches++;
  return LB200_OK;
}

extern "C" int lb200_widen_24bit_dev(lb200_plan* plan, const void* pcm24, void* out, size_t nsamples)
{
  if (!plan || !pcm24 || !out || (nsamples & 3)) return LB200_ERR_BAD_ARG;
  if (nsamples == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  const size_t groups = nsamples / 4;
  int grid = (int)((groups + 255) / 256);
  if (grid > plan->sm_count * 16) grid = plan->sm_count * 16;
  widen_24bit_kernel<<<grid, 256, 0, plan->stream>>>((const uint32_t*)pcm24, (int4*)out, groups);
  LB_CUDA(cudaGetLastError());
  plan->launches++;
  return LB200_OK;
}

extern "C" int lb200_widen_8bit_dev(lb200_plan* plan, const void* pcm8, void* out, size_t nsamples)
{
  if (!plan || !pcm8 || !out || (nsamples & 3)) return LB200_ERR_BAD_ARG;
  if (nsamples == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  const size_t groups = nsamples / 4;
  int grid = (int)((groups + 255) / 256);
  if (grid > plan->sm_count * 16) grid = plan->sm_count * 16;
  widen_8bit_kernel<<<grid, 256, 0, plan->stream>>>((const uint32_t*)pcm8, (uint2*)out, groups);
  LB_CUDA(cudaGetLastError());
  plan->launches++;
  return LB200_OK;
}

extern "C" int lb200_float_to_int32_dev(lb200_plan* plan, const void* f32, void* out, size_t nsamples)
{
  if (!plan || !f32 || !out || (nsamples & 3)) return LB200_ERR_BAD_ARG;
  if (nsamples == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  const size_t groups = nsamples / 4;
  int grid = (int)((groups + 255) / 256);
  if (grid > plan->sm_count * 16) grid = plan->sm_count * 16;
  float_to_int32_kernel<<<grid, 256, 0, plan->stream>>>((const float4*)f32, (int4*)out, groups);
  LB_CUDA(cudaGetLastError());
  plan->launches++;
  return LB200_OK;
}

static int codec_host(lb200_plan* plan, const void* in, size_t in_bytes, void* out, size_t out_bytes, int which, size_t count)
{
  int rc;
  if ((rc = wg_mirror(plan, plan->m_codec_in, in_bytes))) return rc;
  if ((rc = wg_mirror(plan, plan->m_codec_out, out_bytes))) return rc;
  if ((rc = wg_copy(plan, plan->m_codec_in.d, in, in_bytes, true))) return rc;
  switch (which) {
    case 0: rc = lb200_expand_rawdat_dev(plan, plan->m_codec_in.d, plan->m_codec_out.d, count); break;
    case 1: rc = lb200_widen_24bit_dev(plan, plan->m_codec_in.d, plan->m_codec_out.d, count); break;
    case 2: rc = lb200_widen_8bit_dev(plan, plan->m_codec_in.d, plan->m_codec_out.d, count); break;
    default: rc = lb200_float_to_int32_dev(plan, plan->m_codec_in.d, plan->m_codec_out.d, count); break;
  }
  if (rc) return rc;
  if ((rc = wg_copy(plan, out, plan->m_codec_out.d, out_bytes, false))) return rc;
  LB_CUDA(cudaStreamSynchronize(plan->stream));
  return LB200_OK;
}

extern "C" int lb200_expand_rawdat(lb200_plan* plan, const void* packed, void* out, size_t out_bytes)
{
  if (!plan || !packed || !out || (out_bytes & 15)) return LB200_ERR_BAD_ARG;
  if (out_bytes == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  return codec_host(plan, packed, out_bytes / 16 * 9, out, out_bytes, 0, out_bytes);
}

extern "C" int lb200_widen_24bit(lb200_plan* plan, const void* pcm24, void* out, size_t nsamples)
{
  if (!plan || !pcm24 || !out || (nsamples & 3)) return LB200_ERR_BAD_ARG;
  if (nsamples == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  return codec_host(plan, pcm24, nsamples * 3, out, nsamples * 4, 1, nsamples);
}

extern "C" int lb200_widen_8bit(lb200_plan* plan, const void* pcm8, void* out, size_t nsamples)
{
  if (!plan || !pcm8 || !out || (nsamples & 3)) return LB200_ERR_BAD_ARG;
  if (nsamples == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  return codec_host(plan, pcm8, nsamples, out, nsamples * 2, 2, nsamples);
}

extern "C" int lb200_float_to_int32(lb200_plan* plan, const void* f32, void* out, size_t nsamples)
{
  if (!plan || !f32 || !out || (nsamples & 3)) return LB200_ERR_BAD_ARG;
  if (nsamples == 0) return LB200_OK;
  cudaSetDevice(plan->device);
  return codec_host(plan, f32, nsamples * 4, out, nsamples * 4, 3, nsamples);
}
