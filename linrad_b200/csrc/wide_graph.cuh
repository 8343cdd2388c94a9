// wide_graph.cuh -- the two consumers of fft1_sumsq that feed Linrad's wide graph:
//   update_fft1_slowsum (fft1.c:4526-4605) + new_fft1_averages (wide_graph.c:1003-1051):
//       fft1_slowsum = sliding sum of the latest wg_fft_avg2num rows, kept by add-new/subtract-old
//       with a moving window of bins recomputed from scratch each time, clamped at FFT1_SMALL
//   fft1_waterfall (fft1.c:115-223) + update_wg_waterf (fft1.c:104-113):
//       wg_waterf_sum += row; every wg.waterfall_avgnum transforms one line of
//       short = clamp(1000*log10(sum*wg_waterf_yfac)) in one of three pixel mappings.
// Both are sequential over rows and independent over bins, so one thread owns one bin (or one
// pixel) and walks the new rows in order with its running value in a register: the arithmetic
// order per bin is the reference's.  The scalar state machines (recalc window, line counter,
// line pointer) are cheap recurrences every thread repeats; the host repeats them once more to
// hand the updated state back.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lb {

#define LB_FFT1_SMALL 1.0e-20f      /* fft1def.h: FFT1_SMALL */

struct WgK {
  const float* sumsq;       // fft1_sumsq ring
  uint32_t sumsq_mask;      // floats
  uint32_t pa0;             // first new completed row
  int nrows;
  int N;
  // slowsum
  float* slowsum;
  int first_point, last_point;          // fft1_first_point, fft1_last_point
  int wg_first_point, wg_last_point;
  int avg2num, xpoints, fresh_recalc;
  int recalc0, change_flag0;
  int tail_row, tail_recalc;            // the recalculation window's state on entering row tail_row (host-stepped): the
                                        // kernel scans only the last rows of a long call for a bin's last fresh value
  // waterfall
  float* wsum;
  const float* yfac;
  short* waterf;
  int waterf_size, waterf_ptr0, counter0, avg1num, waterfall_avgnum;
  int first_xpoint, xpixels, xpp, ppx;  // wg.xpoints_per_pixel, wg.pixels_per_xpoint
  uint32_t pwg0;            // fft1_sumsq_pwg on entry
  int wrows;                // rows the waterfall drains
};

__device__ __forceinline__ float wg_fresh(const WgK& p, uint32_t pa, int i)
{
  uint32_t p0 = (pa - (uint32_t)(p.avg2num - 1) * (uint32_t)p.N) & p.sumsq_mask;    // wide_graph.c:1016
  float s = p.sumsq[p0 + i];
  for (int m = 1; m < p.avg2num; m++) {
    p0 = (p0 + p.N) & p.sumsq_mask;
    s = __fadd_rn(s, p.sumsq[p0 + i]);
    if (s < LB_FFT1_SMALL) s = LB_FFT1_SMALL;
  }
  return s;
}

__global__ void __launch_bounds__(256) slowsum_kernel(const WgK p)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N) return;
  // A row in whose recalculation window the bin lies (fft1.c:4568-4574) or the change-flag row
  // (fft1.c:4541-4546) REPLACES the sum by a fresh one that does not depend on its history.  First walk the
  // window recurrence alone (integers, no memory) to find the last such row of this call for this bin, then do
  // the arithmetic from there: the value is the reference's, the work no longer grows with the rows of a call.
  const bool change0 = p.change_flag0 != 0;
  int last = -1;
  const int step = p.xpoints / p.fresh_recalc;
  auto scan = [&](int r_first, int recalc, bool change) {
    for (int r = r_first; r < p.nrows; r++) {
      if (change) {
        change = false;
        if (i >= p.wg_first_point && i <= p.wg_last_point) last = r;
        continue;
      }
      if (recalc == p.last_point) recalc = p.first_point;
      const int ia = recalc;
      recalc += step;
      if (recalc > p.last_point) recalc = p.last_point;
      if (i >= ia && i <= recalc) last = r;
    }
  };
  // every in-band bin is recalculated every fresh_recalc (2..8) rows, so the last rows of the call decide; the whole
  // call is scanned only if they do not (they always do for in-band bins when the tail is longer than a sweep)
  if (p.tail_row > 0) scan(p.tail_row, p.tail_recalc, false);
  if (last < 0) scan(0, p.recalc0, change0);
  const bool inband = i >= p.first_point && i <= p.last_point;
  if (last < 0 && !inband) return;                       // untouched by this call
  float s;
  int r0;
  if (last >= 0) {
    s = wg_fresh(p, (p.pa0 + (uint32_t)last * (uint32_t)p.N) & p.sumsq_mask, i);
    r0 = last + 1;
  } else {
    s = p.slowsum[i];
    r0 = change0 ? 1 : 0;                                 // the change-flag row adds nothing outside the wide graph
  }
  if (inband) {
    // the remaining rows: add the new row, subtract the one that leaves the window (fft1.c:4576,4581);
    // their loads are issued twelve rows at a time, the arithmetic stays in row order
    constexpr int G = 12;
    for (int rb = r0; rb < p.nrows; rb += G) {
      float add[G], sub[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        add[g] = 0.f;
        sub[g] = 0.f;
        if (rb + g < p.nrows) {
          const uint32_t pa = (p.pa0 + (uint32_t)(rb + g) * (uint32_t)p.N) & p.sumsq_mask;
          const uint32_t pb = (pa - (uint32_t)p.avg2num * (uint32_t)p.N) & p.sumsq_mask;   // fft1.c:4568
          add[g] = __ldg(p.sumsq + pa + i);
          sub[g] = __ldg(p.sumsq + pb + i);
        }
      }
#pragma unroll
      for (int g = 0; g < G; g++) {
        if (rb + g < p.nrows) {
          s = __fadd_rn(s, __fsub_rn(add[g], sub[g]));
          if (s < LB_FFT1_SMALL) s = LB_FFT1_SMALL;
        }
      }
    }
  }
  p.slowsum[i] = s;
}

__device__ __forceinline__ float wg_log(float v)      // 1000*(float)log10(v), fft1.c:141
{
  return 1000.0f * (float)log10((double)v);
}
__device__ __forceinline__ short wg_clamp(float y)    // fft1.c:142-144
{
  if (y < -32767.0f) y = -32767.0f;
  if (y > 32767.0f) y = 32767.0f;
  return (short)y;
}
__device__ __forceinline__ short wg_clamp_interp(float y)   // fft1.c:169-171: low values go to +32767
{
  if (y < -32767.0f) y = 32767.0f;
  if (y > 32767.0f) y = 32767.0f;
  return (short)y;
}

// Geometry of the waterfall lines of one call (update_wg_waterf, fft1.c:104-113): the counter grows by
// avg1num per row and a line is written when it reaches waterfall_avgnum, after which the sums of the bins on
// the wide graph start again from 0.00001.  So line L depends only on its own rows (line 0 also on the sums
// carried in wg_waterf_sum): lines are independent work.
struct WgLines {
  int rf;       // rows until the first line of the call is complete
  int rl;       // rows per further line
  int nl;       // complete lines in this call
};
__host__ __device__ inline WgLines wg_lines(int counter0, int avg1num, int waterfall_avgnum, int wrows)
{
  WgLines g;
  int need = waterfall_avgnum - counter0;
  g.rf = need <= 0 ? 1 : (need + avg1num - 1) / avg1num;
  if (g.rf < 1) g.rf = 1;
  g.rl = (waterfall_avgnum + avg1num - 1) / avg1num;
  if (g.rl < 1) g.rl = 1;
  g.nl = wrows >= g.rf ? 1 + (wrows - g.rf) / g.rl : 0;
  return g;
}

// One thread per unit u and segment.  mode 0 (1:1): unit = bin u.  mode 1 (xpoints_per_pixel > 1): unit =
// pixel u = bins first_xpoint + u*xpp ...  mode 2 (interpolation, pixels_per_xpoint > 1): unit =
// bin first_xpoint + u; it owns wsum of that bin and the ppx pixels that end on it.
// Launch 1 (handback == 0): blockIdx.y owns the `lps` consecutive lines from line_first + blockIdx.y*lps on and
// walks them in order from the start state of the first.  Launch 2 (handback == 1) = the rows after the last
// complete line: hands the running sums back to wg_waterf_sum (a launch of its own because line 0 reads what
// it overwrites).  The host chooses lps so that there are enough threads when a call brings hundreds of
// rows for few bins, and one thread per unit when the bins alone fill the GPU.
__global__ void __launch_bounds__(256) waterfall_kernel(const WgK p, int mode, int nunits, int handback, int line_first, int lps)
{
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nunits) return;
  int b0, nb;                       // bins whose wsum this thread carries
  if (mode == 0) { b0 = u; nb = 1; }
  else if (mode == 1) { b0 = p.first_xpoint + u * p.xpp; nb = p.xpp; }
  else { b0 = p.first_xpoint + u - 1; nb = 2; }     // [previous bin, own bin]
  const WgLines lg = wg_lines(p.counter0, p.avg1num, p.waterfall_avgnum, p.wrows);
  auto in_wg = [&](int b) { return b >= p.wg_first_point && b <= p.wg_last_point && b < p.N && b >= 0; };
  auto src = [&](uint32_t row, int b) { return p.sumsq[row + (uint32_t)(p.first_xpoint + (b - p.wg_first_point))]; };   // fft1.c:121-126
  // lines [L0, L1) are completed by this walk (none in the hand-back launch)
  // handback == 2: one walk does both (1:1 mapping with one thread per unit: it reads and writes only its own sum)
  const int L0 = handback == 1 ? lg.nl : line_first + (int)blockIdx.y * lps;
  int L1 = handback == 1 ? lg.nl : L0 + lps;
  if (L1 > lg.nl) L1 = lg.nl;
  if (!handback && L0 >= L1) return;
  // rows [rs, re) and the state the sequential walk has on entering row rs
  const int rs = L0 == 0 ? 0 : lg.rf + (L0 - 1) * lg.rl;
  const int re = handback ? p.wrows : lg.rf + (L1 - 1) * lg.rl;
  int counter = L0 == 0 ? p.counter0 : 0;
  int ptr = p.waterf_ptr0 - (int)(((long long)L0 * p.xpixels) % p.waterf_size);
  if (ptr < 0) ptr += p.waterf_size;
  float acc[2];
  if (mode != 1) {
    for (int k = 0; k < nb; k++) {
      const int b = b0 + k;
      acc[k] = (b >= 0 && b < p.N) ? ((L0 == 0 || !in_wg(b)) ? p.wsum[b] : 0.00001f) : 0.f;
    }
  }
  uint32_t row = (p.pwg0 + (uint32_t)rs * (uint32_t)p.N) & p.sumsq_mask;
  int seg_start = rs;               // first row of the current (unfinished) line, mode 1
  // modes 0 and 2: the rows' values are fetched twelve rows at a time (loads in flight together),
  // the sums are formed in row order
  constexpr int G = 12;
  float pre[2][G];
  for (int r = rs; r < re; r++) {
    if (mode != 1) {
      const int g = (r - rs) % G;
      if (g == 0) {
        uint32_t rw = row;
#pragma unroll
        for (int gg = 0; gg < G; gg++) {
#pragma unroll
          for (int k = 0; k < 2; k++) pre[k][gg] = (k < nb && r + gg < re && in_wg(b0 + k)) ? __ldg(&p.sumsq[rw + (uint32_t)(p.first_xpoint + (b0 + k - p.wg_first_point))]) : 0.f;
          rw = (rw + p.N) & p.sumsq_mask;
        }
      }
#pragma unroll
      for (int k = 0; k < 2; k++) {
        if (k < nb && in_wg(b0 + k)) {
          float x = pre[k][0];
#pragma unroll
          for (int gg = 1; gg < G; gg++) x = g == gg ? pre[k][gg] : x;
          acc[k] = __fadd_rn(acc[k], x);
        }
      }
    }
    row = (row + p.N) & p.sumsq_mask;
    counter += p.avg1num;
    if (counter >= p.waterfall_avgnum) {
      if (mode == 0) {
        const int ix = u - p.first_xpoint;
        if (ix >= 0 && ix < p.xpixels) p.waterf[ptr + ix] = wg_clamp(wg_log(acc[0] * p.yfac[u]));
        if (in_wg(u)) acc[0] = 0.00001f;
      } else if (mode == 1) {
        float t1 = 0.f;
        for (int k = 0; k < nb; k++) {
          const int b = b0 + k;
          if (b >= p.N) break;
          float v = p.wsum[b];
          if (in_wg(b)) {
            if (seg_start > 0) v = 0.00001f;          // reset by the previous line
            uint32_t rr = (p.pwg0 + (uint32_t)seg_start * (uint32_t)p.N) & p.sumsq_mask;
            for (int q = seg_start; q <= r; q++) { v = __fadd_rn(v, src(rr, b)); rr = (rr + p.N) & p.sumsq_mask; }
          }
          const float t2 = v * p.yfac[b];
          if (t2 > t1) t1 = t2;
        }
        if (u < p.xpixels) p.waterf[ptr + u] = wg_clamp(wg_log(t1));
        seg_start = r + 1;
      } else {
        // fft1.c:158-205: pixel 0 from the first bin; then ppx pixels per further bin, a running
        // float sum from the previous bin's level to this bin's
        const int m = p.xpixels - p.ppx;
        if (u == 0) {
          p.waterf[ptr] = wg_clamp(wg_log(acc[1] * p.yfac[p.first_xpoint]));
        } else {
          const int ix = (u - 1) * p.ppx;
          const int i = p.first_xpoint + u;
          const bool regular = ix < m;
          const int groups = (m + p.ppx - 1) / p.ppx;            // iterations of the ix loop
          const bool tail = (u - 1 == (groups > 0 ? groups : 0)) && i < p.N;     // fft1.c:190
          if ((regular || tail) && i < p.N) {
            float yval = wg_log(acc[0] * p.yfac[i - 1]);
            const float t1 = wg_log(acc[1] * p.yfac[i]);
            const float der = (t1 - yval) / (float)p.ppx;
            for (int k = ix + 1; k <= ix + p.ppx; k++) {
              yval = __fadd_rn(yval, der);
              p.waterf[ptr + k] = wg_clamp_interp(yval);
            }
          }
        }
        for (int k = 0; k < nb; k++)
          if (in_wg(b0 + k)) acc[k] = 0.00001f;
      }
      counter = 0;                                              // update_wg_waterf, fft1.c:104-113
      ptr -= p.xpixels;
      if (ptr < 0) ptr += p.waterf_size;
    }
  }
  if (!handback) return;
  // ---- hand the running sums back (the rows after the last complete line)
  if (mode == 0) {
    if (in_wg(u)) p.wsum[u] = acc[0];
  } else if (mode == 2) {
    const int b = p.first_xpoint + u;
    if (in_wg(b)) p.wsum[b] = acc[1];
  } else {
    for (int k = 0; k < nb; k++) {
      const int b = b0 + k;
      if (b >= p.N || !in_wg(b)) continue;
      float v = p.wsum[b];
      if (seg_start > 0) v = 0.00001f;
      uint32_t rr = (p.pwg0 + (uint32_t)seg_start * (uint32_t)p.N) & p.sumsq_mask;
      for (int q = seg_start; q < p.wrows; q++) { v = __fadd_rn(v, src(rr, b)); rr = (rr + p.N) & p.sumsq_mask; }
      p.wsum[b] = v;
    }
  }
}

}  // namespace lb
