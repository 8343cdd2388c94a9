// mix1.cuh -- first mixer: select M bins around mix1_point, taper, back-transform to the
// decimated baseband and overlap into timf3.  Replaces, for a batch of transforms and any
// number of selections,
//   fft1_mix1_fixed   (mix1.c:995-1041: bin gather with the first/last point clamps)
//   do_mix1           (mix1.c:55-272 one channel, 453-645 two channels): mix1_fqwin taper,
//                     fftback (fft0.c:481-533; 2-ch dual_fftback fft0.c:341), phase rotation
//                     and the three overlap schemes (none / sin^2 50 % / crossover windows)
//   mix1_clear        (mix1.c:770-779)
// set_mix1_phases (mix1.c:781-861) stays on the host (plan.cu builds one Mix1Job per
// transform and selection, carrying the reference's float phase state bit-exactly).
//
// One CTA walks a *run* of consecutive transforms of one selection, PAR of them at a time side
// by side, so that the raw tail of transform b-1 (which the reference parks in timf3 and
// re-reads, mix1.c:178-194) is found in shared memory; only the first transform of a run has to
// rebuild its predecessor.
#pragma once
#include "fft_core.cuh"
#include "phase.h"

namespace lb {

struct Mix1Job {          // one (transform, selection)
  uint32_t src;           // float index of the transform's block in fft1_float
  uint32_t dst;           // timf3_pa for this transform (float index inside the selection's ring)
  int point;              // mix1_point[ss]; <0: selection empty -> mix1_clear
  float t1, t2;           // mix1_phase[ss] after set_mix1_phases, mix1_phase_rot[ss]
  float r1, r2;           // mix1_old_phase[ss], r2 of mix1.c:167
};

struct Mix1K {
  const float* fft1;      // fft1_float ring
  uint32_t fft1_mask;     // floats
  const Mix1Job* jobs;    // [nsel][nblocks]
  int nblocks, nsel, runlen;
  float* timf3;           // selection ss at timf3 + ss*sel_stride
  size_t sel_stride;      // floats between selections (2*timf3_size)
  uint32_t timf3_mask;    // floats
  const float2* Wm;       // exp(-2 pi i m / M)
  const float* fqwin;
  const float* window;    // inverse window (crossover mode)
  const float* cos2win;
  const float* sin2win;
  int first_point, last_point;
  int Mi, Mn, cross;      // mix1.interleave_points, new_points, crossover_points
  int mode;               // 0 none, 1 sin^2 (Mi==Mn), 2 crossover
  float2* ybuf_g;         // two-launch form (STAGE 1 / 2): back-transformed blocks [nsel][nblocks][NCH][M]
};

// sin/cos of a mixer phase (the reference takes sin()/cos() in double of the float phase,
// mix1.c:147-148): two-constant reduction to [-pi, pi], then the SFU approximations, absolute
// error below 5e-7 -- far inside the 2e-5 baseband tolerance.
LB_D void mix1_sincos(float x, float* s, float* c)
{
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, x);           // float(2 pi)
  r = fmaf(-k, -1.7484555e-7f, r);                      // 2 pi - float(2 pi)
  *s = __sinf(r);
  *c = __cosf(r);
}

// [start, start+bytes) of fft1_float -> L2 (16-byte granules; a transform's block does not wrap)
LB_D void mix1_l2_prefetch(const unsigned char* base, uint32_t start, uint32_t bytes)
{
#if defined(__CUDA_ARCH__)
  const uint32_t a = start & ~15u;
  const uint32_t len = (bytes + (start & 15u) + 15u) & ~15u;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + a), "r"(len) : "memory");
#endif
}

// taper index of do_mix1 (mix1.c:113-135 / 455-491, including the doubled factors of the
// two-channel loop at i==M-1 and i==M/2)
template <int NCH>
LB_D float mix1_taper(const float* fqwin, int i, int M)
{
  const int h = M / 2;
  float w;
  if (i == 0) w = fqwin[h - 1];
  else if (i <= h) w = fqwin[h - i];
  else w = fqwin[i - h];
  if (NCH == 2) {
    if (i == M - 1) w *= fqwin[h - 1];
    if (i == h) w *= fqwin[0];
  }
  return w;
}

// Threads per CTA: PAR transforms side by side, each by NCH*T threads.
// STAGE 0: everything in one kernel (runs of consecutive transforms, the predecessor's tail found in shared
// memory).  mix1.size 16384 (8192 with two channels) does not leave room for that: there the two halves are two
// launches with every transform its own work item -- STAGE 1 = gather + taper + back transform into ybuf_g,
// STAGE 2 = phase rotation and overlap from ybuf_g into timf3.  (For the sizes that fit, the two-launch form was
// measured slower than STAGE 0: configs[3] pass 0.234 vs 0.231 ms, profiles/r2_notes.txt.)
template <int LOG2M, int LOG2E, int NCH, int PAR, int STAGE>
__global__ void __launch_bounds__((PAR * NCH) << (LOG2M - LOG2E), (((PAR * NCH) << (LOG2M - LOG2E)) <= 512 && LOG2M < 13 + (NCH == 1) ? 2 : 1))
mix1_kernel(const Mix1K p)
{
  using P = Plan<LOG2M, LOG2E>;
  constexpr int M = P::N, E = P::E, T = P::T, MM = 2 * NCH;
  constexpr int LANE_THREADS = NCH * T;
  constexpr int SLOTS = PAR + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: xch[PAR][NCH][M+M/32+32] | ybuf[SLOTS][NCH][M]; after its transform a lane reuses
  // its first exchange slice for the two phase chains (ph_t | ph_r, one pad float per 16: a thread fills 16
  // consecutive phases, so the pad puts the threads of a warp 17 floats apart = on 32 different banks)
  float2* xch_all = reinterpret_cast<float2*>(smem_raw);
  constexpr int XCH = M + M / 32 + 32;
  float2* ybuf_all = xch_all + PAR * NCH * XCH;

  const int tid = threadIdx.x;
  const int lane = tid / LANE_THREADS;          // which transform of the chunk
  const int lt = tid - lane * LANE_THREADS;
  const int ch = lt / T;
  const int t = lt - ch * T;
  float2* xch = xch_all + (lane * NCH + ch) * XCH;
  float* ph_t = reinterpret_cast<float*>(xch_all + lane * NCH * XCH);
  float* ph_r = ph_t + M + M / 16;
  auto pp = [](int i) { return i + (i >> 4); };
  Twiddles<P> tw;
  if (STAGE != 2) load_twiddles<P>(tw, p.Wm, t);

  const int runs_per_sel = (p.nblocks + p.runlen - 1) / p.runlen;
  const int nruns = runs_per_sel * p.nsel;
  const int carry_len = p.mode == 1 ? M / 2 : (p.mode == 2 ? p.cross : 0);
  const int yoff = p.mode == 2 ? (p.Mi / 2 - p.cross / 2) : 0;
  const int carry_src = p.mode == 1 ? M / 2 : p.Mn + yoff;
  const int nt = p.Mn;                         // rotated samples per transform
  const int nr = carry_len;

  for (int run = blockIdx.x; run < nruns; run += gridDim.x) {
    const int ss = run / runs_per_sel;
    const int bfirst = (run - ss * runs_per_sel) * p.runlen;
    int blast = bfirst + p.runlen;
    if (blast > p.nblocks) blast = p.nblocks;
    const Mix1Job* jobs = p.jobs + (size_t)ss * p.nblocks;
    float* t3 = p.timf3 + (size_t)ss * p.sel_stride;
    // ystart = bfirst-1 rebuilds the predecessor's tail when this run does not start the call
    const int ystart = (STAGE == 0 && bfirst > 0 && carry_len > 0) ? bfirst - 1 : bfirst;
    // the spectra this run will gather (fft1_float has long left L2 when the batch is large): one thread
    // per transform asks for its M bins now, so that only the first gather of the run waits for DRAM
    if (STAGE != 2 && tid < blast - ystart && tid < 32) {
      const Mix1Job pj = jobs[ystart + tid];
      if (pj.point >= 0) {
        int lo = pj.point - M / 2, hi = pj.point + M / 2;
        if (lo < p.first_point) lo = p.first_point;
        if (hi > p.last_point) hi = p.last_point;
        if (hi > lo) {
          const uint32_t start = ((pj.src & p.fft1_mask) + (uint32_t)lo * MM) * 4u;
          mix1_l2_prefetch(reinterpret_cast<const unsigned char*>(p.fft1), start, (uint32_t)(hi - lo) * MM * 4u);
        }
      }
    }
    for (int y0 = ystart; y0 < blast; y0 += PAR) {
      const int b = y0 + lane;                 // this lane's transform
      const bool active = b < blast;
      const bool emit = active && b >= bfirst; // the predecessor only lends its tail
      Mix1Job job;
      job.point = -1;
      if (active) job = jobs[b];
      float2* ybuf = STAGE == 0 ? ybuf_all + (((b - ystart) % SLOTS) * NCH + ch) * M
                                : p.ybuf_g + (((size_t)ss * p.nblocks + (active ? b : 0)) * NCH + ch) * M;
      if (STAGE != 2) {
      // ---- gather + taper (mix1.c:1015-1030, 113-135); idle lanes run the transform on zeros
      float2 v[E];
      if (job.point >= 0) {
        const float* src = p.fft1 + (job.src & p.fft1_mask);
#pragma unroll
        for (int e = 0; e < E; e++) {
          const int i = t + T * e;
          const int bin = (i < M / 2) ? job.point + i : job.point - M + i;
          const bool ok = (i < M / 2) ? (bin < p.last_point) : (bin >= p.first_point);
          float2 z = make_float2(0.f, 0.f);
          if (ok) z = *reinterpret_cast<const float2*>(src + (size_t)bin * MM + 2 * ch);
          const float w = mix1_taper<NCH>(p.fqwin, i, M);
          v[e] = make_float2(z.x * w, z.y * w);
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) v[e] = make_float2(0.f, 0.f);
      }
      fft_forward<P>(v, xch, t, tw);           // fftback: sum_k y_k exp(-2 pi i n k / M)
      if (active) {
#pragma unroll
        for (int e = 0; e < E; e++) ybuf[t + T * e] = v[e];
      }
      }
      if (STAGE == 1) {
        __syncthreads();                       // the exchange buffer is free for the next chunk
        continue;
      }
      // ---- exact float phase chains for this transform (mix1.c:143-153,164-186)
      if (emit && job.point >= 0) {
        // one closed-form jump per 16 samples (it costs ~300 instructions), then plain float adds
        int chunk_t = (nt + LANE_THREADS - 1) / LANE_THREADS;
        if (chunk_t < 16) chunk_t = 16;
        int i0 = lt * chunk_t;
        if (i0 < nt) {
          float x = lb_phase_advance(job.t1, job.t2, i0);
          int i1 = i0 + chunk_t; if (i1 > nt) i1 = nt;
          for (int i = i0; i < i1; i++) { ph_t[pp(i)] = x; x = lb_float_add(x, job.t2); }
        }
        int chunk_r = (nr + LANE_THREADS - 1) / LANE_THREADS;
        if (chunk_r < 16 && nr > 0) chunk_r = 16;
        // the second chain is dealt from the other end of the lane: with 16 samples per thread each chain keeps
        // only part of the threads busy, and this way they are different threads
        i0 = (LANE_THREADS - 1 - lt) * chunk_r;
        if (chunk_r > 0 && i0 < nr) {
          float x = lb_phase_advance(job.r1, job.r2, i0);
          int i1 = i0 + chunk_r; if (i1 > nr) i1 = nr;
          for (int i = i0; i < i1; i++) { ph_r[pp(i)] = x; x = lb_float_add(x, job.r2); }
        }
      }
      __syncthreads();
      if (emit) {
        const float2* yb = STAGE == 0 ? ybuf_all + (((b - ystart) % SLOTS) * NCH) * M          // [NCH][M]
                                      : p.ybuf_g + (((size_t)ss * p.nblocks + b) * NCH) * M;
        const float2* cb = STAGE == 0 ? ybuf_all + (((b - 1 - ystart + SLOTS) % SLOTS) * NCH) * M   // predecessor
                                      : p.ybuf_g + (((size_t)ss * p.nblocks + (b > 0 ? b - 1 : 0)) * NCH) * M;
        const bool from_ring = (b == 0);      // first transform of the call: tail is in timf3
        if (job.point < 0) {
          // mix1_clear: zero timf3_block floats
          for (int s = lt; s < p.Mn * MM; s += LANE_THREADS) t3[(job.dst + s) & p.timf3_mask] = 0.f;
        } else {
          for (int s = lt; s < nt; s += LANE_THREADS) {
            float st, ct;
            mix1_sincos(ph_t[pp(s)], &st, &ct);
            float sr = 0.f, cr = 1.f;
            if (s < nr) mix1_sincos(ph_r[pp(s)], &sr, &cr);
            const uint32_t o = (job.dst + (uint32_t)s * MM) & p.timf3_mask;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
              const float2 y = yb[c * M + s + yoff];
              float re = ct * y.x - st * y.y;
              float im = ct * y.y + st * y.x;
              if (p.mode == 1) {
                float2 a;
                if (from_ring) a = *reinterpret_cast<const float2*>(t3 + o + 2 * c);
                else a = cb[c * M + carry_src + s];
                re = cr * a.x - sr * a.y + re;           // mix1.c:180-181
                im = cr * a.y + sr * a.x + im;
              } else if (p.mode == 2) {
                if (s < p.cross) {
                  float2 a;
                  if (from_ring) a = *reinterpret_cast<const float2*>(t3 + o + 2 * c);
                  else a = cb[c * M + carry_src + s];
                  const float w1 = p.sin2win[s], w2 = p.cos2win[s];
                  a.x *= w2; a.y *= w2;
                  re = cr * a.x - sr * a.y + re * w1;    // mix1.c:219-224
                  im = cr * a.y + sr * a.x + im * w1;
                } else {
                  const int sb = p.Mn / 2 + 1 + p.cross / 2;
                  const int j = (s < sb) ? (yoff + s) : (yoff + 2 * sb - 2 - s);
                  const float w = p.window[j];
                  re *= w; im *= w;                      // mix1.c:237-239,253-255
                }
              }
              *reinterpret_cast<float2*>(t3 + o + 2 * c) = make_float2(re, im);
            }
          }
        }
        // the raw tail of the LAST transform of the call is parked in timf3 for the next call
        if (b == p.nblocks - 1 && carry_len > 0) {
          for (int s = lt; s < carry_len; s += LANE_THREADS) {
            const uint32_t o = (job.dst + (uint32_t)(p.Mn + s) * MM) & p.timf3_mask;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
              float2 y = make_float2(0.f, 0.f);
              if (job.point >= 0) y = yb[c * M + carry_src + s];
              // after mix1_clear the reference leaves whatever the ring held; a cleared
              // selection has no defined tail, zeros keep the next blend finite
              *reinterpret_cast<float2*>(t3 + o + 2 * c) = y;
            }
          }
        }
      }
      __syncthreads();                         // slots and phase chains are free for the next chunk
    }
  }
}

template <int LOG2M, int NCH, int PAR>
constexpr size_t mix1_smem()
{
  return sizeof(float2) * (size_t)(PAR * NCH * ((1 << LOG2M) + (1 << LOG2M) / 32 + 32) + (PAR + 1) * NCH * (1 << LOG2M));
}

// the two-launch form keeps only the exchange slices (STAGE 2: the phase chains) in shared memory
template <int LOG2M, int NCH, int PAR>
constexpr size_t mix1_smem_split()
{
  return sizeof(float2) * (size_t)(PAR * NCH * ((1 << LOG2M) + (1 << LOG2M) / 32 + 32));
}

}  // namespace lb
