// fft1_fused.cuh -- the fused fft1 kernel for 2^10 <= N <= 2^14 (one CTA per transform, 32 points
// per thread).  One launch replaces, for a batch of consecutive time blocks,
//   fft1win_dit_one / fft1win_dif_chan   (fft1.c:684-1028, 2041-2247: ring gather, int->float, window)
//   bulk_of_dit / bulk_of_dif + permute  (fft0.c:1590-1769, 161-195; fft1.c:637-682)
//   the direction flip of fft1_b         (fft1.c:3660-3680, 4029-4060)
//   fft1_c                               (fft1.c:4115-4200: filtercorr multiply, |z|^2, fft1_sumsq)
// Output convention (probed from the compiled reference):
//   fft1_float[k] = conj( sum_n w[n] x[n] exp(-2 pi i n ((k+N/2) mod N)/N) ) * filtercorr[k]
//
// How the conventions are folded away so that they cost no instructions:
//   * the (k+N/2) rotation is (-1)^n on the input: the sign is stored in the window table;
//   * conj(DFT(z)) = swap(DFT(swap(conj z))) with swap(a+ib) = b+ia, so the kernel loads
//     (-Q*w, I*w), runs the forward transform and stores (im, re); for fft1_direction < 0 the
//     reference reverses the spectrum and swaps re/im, which is DFT(conj z) stored as (im, re):
//     load (I*w, -Q*w).  Negated operands are free in FMUL.
//   * FC == 1: an uncalibrated fft1_filtercorr (clear_fft1_filtercorr, fft1.c:4673-4724) is one
//     real gain except on the 16 outermost bins at each end; the gain is folded into the window
//     table and the edge bins are corrected by their ratio to it.
//
// Work decomposition: one CTA owns one channel of one averaging group (the wg.fft_avg1num
// consecutive transforms summed into one fft1_sumsq row, fft1.c:4507-4520) and walks through its
// transforms in time order, so the row is accumulated on chip in the reference's own order; with
// two channels the two CTAs of a group add their shares into the (host-zeroed) row.  While transform b is computed the new timf1 bytes of transform
// b+1 are pulled into L2 with one TMA bulk prefetch (cp.async.bulk.prefetch.L2).
#pragma once
#include "fft32_core.cuh"
#include "fft1_small.cuh"

namespace lb {

enum FcMode { FC_RAW = 0, FC_FOLDED = 1, FC_TABLE = 2 };

LB_D void l2_prefetch_span(const uint8_t* ring, uint32_t mask, uint32_t start, uint32_t bytes)
{
#if defined(__CUDA_ARCH__)
  // [start, start+bytes) on the ring, widened to 16-byte granules, split at the wrap
  uint32_t a = start & mask & ~15u;
  uint32_t len = (bytes + (start & 15u) + 15u) & ~15u;
  const uint32_t size = mask + 1u;
  while (len > 0) {
    uint32_t n = size - a;
    if (n > len) n = len;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ring + a), "r"(n) : "memory");
    len -= n;
    a = 0;
  }
#endif
}

// fft1_float is written once and not read again by this kernel: stream it through L2
// (evict-first) so that it does not push the input span and the channel-0 scratch row out
#ifndef LB_NO_STREAM_HINTS
LB_D void lb_store_stream(float2* p, float2 v) { __stcs(p, v); }
LB_D void lb_store_stream(float4* p, float4 v) { __stcs(p, v); }
LB_D float2 lb_load_last(const float2* p) { return __ldcs(p); }
#else
LB_D void lb_store_stream(float2* p, float2 v) { *p = v; }
LB_D void lb_store_stream(float4* p, float4 v) { *p = v; }
LB_D float2 lb_load_last(const float2* p) { return *p; }
#endif

// Split-phase CTA barrier on an mbarrier in shared memory: a thread ARRIVES as soon as it has
// finished reading an exchange buffer and only WAITS right before it overwrites that buffer, so
// the write-after-read hand-over costs no stall when the other warps are already past it.
LB_D void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
LB_D void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
LB_D void mbar_wait(uint64_t* bar, uint32_t parity)
{
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}

template <int FMT>
LB_D float2 cvt_iq(const uint8_t* p)
{
  if (FMT == FMT_I16_1CH || FMT == FMT_I16_2CH) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
    return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
  } else {
    const int2 w = *reinterpret_cast<const int2*>(p);
    return make_float2((float)w.x, (float)w.y);
  }
}

// ---- asynchronous staging of the raw timf1 frames through the exchange buffer ----------------
// The slots of the exchange buffer a thread reads in the LAST exchange of a transform are its own
// (nobody else reads or writes them until the next transform's first exchange store).  Right after
// that read the thread issues cp.async (LDGSTS) copies of its 32 raw frames of the NEXT transform
// into those very slots: the global-load latency then runs under the last radix-32 pass, the
// epilogue and the stores of the current transform, and costs neither registers nor LSU stalls.
// 8-byte channels (int32 I/Q) use one slot per frame, 4-byte channels (int16 I/Q) share a slot
// between frames e and e+1.
LB_D void cp_async_8(void* smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
LB_D void cp_async_4(void* smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
LB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
LB_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// float2 slot index (relative to the thread's base slot) of element e in the last exchange's load pattern
template <int LOG2N>
LB_HD constexpr int last_xch_stride()
{
  using P = Plan32<LOG2N>;
  return P::NPASS == 3 ? (P::T + P::T / 32) : (P::T + (P::T >> P::SH1) * 2);
}
template <int LOG2N>
LB_HD int last_xch_base(int t)
{
  using P = Plan32<LOG2N>;
  return P::NPASS == 3 ? pad1(t) : pad2<P::SH1>(t);
}

template <int LOG2N, int FMT>
LB_D void raw_issue(float2* slot0, const uint8_t* ring, uint32_t mask, uint32_t start, int t, int c)
{
  using P = Plan32<LOG2N>;
  constexpr int T = P::T;
  constexpr int FRAME = FmtInfo<FMT>::FRAME, NCH = FmtInfo<FMT>::NCH, CHB = FRAME / NCH;
  constexpr int STR = last_xch_stride<LOG2N>();
  const uint32_t span = (uint32_t)P::N * FRAME;
  if (start + span <= mask + 1u) {
#if defined(LB_EXP) && (LB_EXP & 8)
    const uint8_t* srcx = ring + start + (uint32_t)t * CHB + c * (P::N * CHB);
#pragma unroll
    for (int e = 0; e < 32; e++) {
      if (CHB == 8) cp_async_8(slot0 + e * STR, srcx + (size_t)e * (T * CHB));
      else cp_async_4(reinterpret_cast<uint32_t*>(slot0 + (e >> 1) * STR) + (e & 1), srcx + (size_t)e * (T * CHB));
    }
    cp_async_commit();
    return;
#endif
    const uint8_t* src = ring + start + (uint32_t)t * FRAME + c * CHB;
#pragma unroll
    for (int e = 0; e < 32; e++) {
      if (CHB == 8) cp_async_8(slot0 + e * STR, src + (size_t)e * (T * FRAME));
      else cp_async_4(reinterpret_cast<uint32_t*>(slot0 + (e >> 1) * STR) + (e & 1), src + (size_t)e * (T * FRAME));
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; e++) {
      const uint32_t off = (start + (uint32_t)(t + T * e) * FRAME) & mask;
      if (CHB == 8) cp_async_8(slot0 + e * STR, ring + off + c * CHB);
      else cp_async_4(reinterpret_cast<uint32_t*>(slot0 + (e >> 1) * STR) + (e & 1), ring + off + c * CHB);
    }
  }
  cp_async_commit();
}

// raw frames -> float, window (sign and, for FC_FOLDED, gain are in the table), conj/direction swap
template <int LOG2N, int FMT>
LB_D void raw_take(float2 (&v)[32], const float2* slot0, const float* wtab, int direction)
{
  using P = Plan32<LOG2N>;
  constexpr int T = P::T;
  constexpr int FRAME = FmtInfo<FMT>::FRAME, NCH = FmtInfo<FMT>::NCH, CHB = FRAME / NCH;
  constexpr int STR = last_xch_stride<LOG2N>();
  if (CHB == 8) {
#pragma unroll
    for (int e = 0; e < 32; e++) {
      const int2 r = *reinterpret_cast<const int2*>(slot0 + e * STR);
      v[e] = make_float2((float)r.x, (float)r.y);
    }
  } else {
#pragma unroll
    for (int h = 0; h < 16; h++) {
      const uint2 r = *reinterpret_cast<const uint2*>(slot0 + h * STR);
      v[2 * h] = make_float2((float)(short)(r.x & 0xffffu), (float)(short)(r.x >> 16));
      v[2 * h + 1] = make_float2((float)(short)(r.y & 0xffffu), (float)(short)(r.y >> 16));
    }
  }
  if (direction > 0) {
#pragma unroll
    for (int e = 0; e < 32; e++) {
      const float w = wtab[e * T];
      v[e] = make_float2(v[e].y * -w, v[e].x * w);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; e++) {
      const float w = wtab[e * T];
      v[e] = make_float2(v[e].x * w, v[e].y * -w);
    }
  }
}

template <int LOG2N, int FMT, int FC>
__global__ void __launch_bounds__(1 << (LOG2N - 5), 512 >> (LOG2N - 5))
fft1_fused_kernel(const Fft1K p)
{
  using P = Plan32<LOG2N>;
  constexpr int N = P::N, T = P::T;
  constexpr int FRAME = FmtInfo<FMT>::FRAME, NCH = FmtInfo<FMT>::NCH, MM = 2 * NCH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* xch = reinterpret_cast<float2*>(smem_raw);
  float4* tab1 = reinterpret_cast<float4*>(smem_raw + sizeof(float2) * P::XCH);
  float* wsm = reinterpret_cast<float*>(smem_raw + sizeof(float2) * P::XCH + sizeof(float4) * P::TAB1);
  const int t = threadIdx.x;

  // pass-1 twiddle table [pair][k] and the window table -> shared memory; last-pass exact powers
  // -> registers
  if (P::NPASS == 3)
    for (int i = t; i < P::TAB1; i += T) tab1[i] = p.tab1[i];
  for (int i = t; i < N; i += T) wsm[i] = p.wtab[i];
  float2 wb[5];
#pragma unroll
  for (int j = 0; j < 5; j++) wb[j] = p.Wn[t << j];
  const float* wtab = wsm + t;
  __shared__ uint64_t war_bar[2];            // [0]: exchange 1 has been read, [1]: the staged raw frames have been read
  if (t == 0) {
    mbar_init(&war_bar[0], T);
    mbar_init(&war_bar[1], T);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t par0 = 0, par1 = 0;
  // All CTAs of a launch run the same load / transform / store cycle with the same period.  Started
  // together they stay in step for the whole launch: the memory system sees every SM storing at
  // once, then nothing.  A pseudo-random start offset per work pair spreads the phases.
  if (t == 0 && p.stagger_ns) {
    const uint32_t h = (((uint32_t)blockIdx.x / NCH) * 2654435761u) >> 22;      // 10 bits
    const uint64_t d = ((uint64_t)p.stagger_ns * h) >> 10;
    uint64_t t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      __nanosleep(256);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < d);
  }
  __syncthreads();

  const int group_size = p.power_rows ? 1 : p.avg1num;
  const int c0 = p.power_rows ? 0 : p.counter0;
  const int ngroups = (c0 + p.nblocks + group_size - 1) / group_size;
  const uint32_t span = (uint32_t)N * FRAME;
  // work item = (averaging group, channel).  The two channels of a group go to neighbouring CTAs:
  // they read the same timf1 frames at the same time (one DRAM fetch, the second CTA hits L2) and
  // each writes its own 8-byte half of every output slot; the halves of a 32-byte sector arrive
  // within microseconds of each other and leave L2 as whole sectors.
  const int nwork = ngroups * NCH;
  int w = blockIdx.x;
  if (w >= nwork) return;
  // block range [b0, b1) of work item w
  auto range_of = [&](int ww, int& bb0, int& bb1) {
    const int gg = ww / NCH;
    bb0 = gg * group_size - c0;
    bb1 = bb0 + group_size;
    if (bb0 < 0) bb0 = 0;
    if (bb1 > p.nblocks) bb1 = p.nblocks;
  };
  auto start_of = [&](int bb) { return (p.ref0 + (uint32_t)bb * p.blockbytes - p.pre_bytes) & p.ring_mask; };
  int b0, b1;
  range_of(w, b0, b1);
  int b = b0;
  float2* slot0 = xch + last_xch_base<LOG2N>(t);
  raw_issue<LOG2N, FMT>(slot0, p.timf1, p.ring_mask, start_of(b), t, w % NCH);

  while (true) {
    const int g = w / NCH;
    const int c = w - g * NCH;
    // the transform after this one (same group, or the first of this CTA's next work item)
    int nw = w, nb = b + 1, nb0 = b0, nb1 = b1;
    bool have_next = true;
    if (nb >= b1) {
      nw = w + (int)gridDim.x;
      have_next = nw < nwork;
      if (have_next) {
        range_of(nw, nb0, nb1);
        nb = nb0;
      }
    }
    float* outb = p.out + ((p.out_pa + (uint32_t)b * (uint32_t)(MM * N)) & p.out_mask) + (size_t)t * MM + 2 * c;
    {
      float2 v[32];
      cp_async_wait_all();                       // my own raw frames have landed in my own slots
      raw_take<LOG2N, FMT>(v, slot0, wtab, p.direction);
      mbar_arrive(&war_bar[1]);                  // my slots may be overwritten by the first exchange
      // ---- transform
      pass0<P::R0>(v);
      mbar_wait(&war_bar[1], par1);              // everybody has taken their raw frames
      par1 ^= 1;
      exch1_store<LOG2N>(v, xch, t);
      __syncthreads();
      exch1_load<LOG2N>(v, xch, t);
      if (P::NPASS == 3) {
        mbar_arrive(&war_bar[0]);
        {
          float2 w32[32];
          const float4* tp = tab1 + (t & (P::R0 - 1));
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const float4 f = tp[q * P::R0];
            w32[2 * q] = make_float2(f.x, f.y);
            w32[2 * q + 1] = make_float2(f.z, f.w);
          }
          radix32_table(v, w32);
        }
        mbar_wait(&war_bar[0], par0);            // everybody has read exchange 1
        par0 ^= 1;
        exch2_store<LOG2N>(v, xch, t);
        __syncthreads();
        exch2_load<LOG2N>(v, xch, t);
      }
      // ---- the slots just read are free: stage the next transform's raw frames into them
      if (have_next) {
        raw_issue<LOG2N, FMT>(slot0, p.timf1, p.ring_mask, start_of(nb), t, nw % NCH);
        if (t == 0 && (nw % NCH) == 0) {
          // pull what comes after that into L2: the new bytes of the following transform of the
          // same group, or the whole span of the first transform of the work item after it
          if (nb + 1 < nb1) {
            l2_prefetch_span(p.timf1, p.ring_mask, start_of(nb) + span, p.blockbytes);
          } else if (nw + (int)gridDim.x < nwork) {
            int fb0, fb1;
            range_of(nw + (int)gridDim.x, fb0, fb1);
            if (fb0 < fb1) l2_prefetch_span(p.timf1, p.ring_mask, start_of(fb0), span);
          }
        }
      }
      radix32_gen(v, wb);
      // ---- epilogue: bin k = t + T*e; v holds (im, re) of the output value.  |z|^2 goes
      // straight to the fft1_sumsq row in L2 (fft1.c:4507-4520 sums the transforms of a group in
      // time order; so do the reductions of one thread on one address): the first transform of a
      // one-channel group stores, everything else is a fire-and-forget RED.ADD.  Two-channel rows
      // are zeroed by the host because the two channel CTAs of a group add into the same row.
      if (FC != FC_RAW) {
        const bool rows = p.power_rows != nullptr;
        float* prow = rows ? p.power_rows + (size_t)b * N + t
                           : (p.sumsq ? p.sumsq + ((p.sumsq_pa + (uint32_t)g * (uint32_t)N) & p.sumsq_mask) + t : nullptr);
        const bool plain = (NCH == 1) && (rows || (b == b0 && !(g == 0 && p.counter0 > 0)));
        if (FC == FC_FOLDED) {                   // host guarantees the full bin range here
          if (t < 16) {                          // bins 0..15 and N-16..N-1 carry the taper of fft1.c:4703-4722
            const float2 f = p.edge[t];
            v[0] = make_float2(v[0].x * f.x + v[0].y * f.y, v[0].y * f.x - v[0].x * f.y);
          }
          if (t >= T - 16) {
            const float2 f = p.edge[16 + t - (T - 16)];
            v[31] = make_float2(v[31].x * f.x + v[31].y * f.y, v[31].y * f.x - v[31].x * f.y);
          }
#if defined(LB_EXP) && (LB_EXP & 4)
          if (prow && v[3].x == 1.2345f) {
#else
          if (prow) {
#endif
            if (plain) {
#pragma unroll
              for (int e = 0; e < 32; e++) prow[e * T] = fmaf(v[e].x, v[e].x, v[e].y * v[e].y);
            } else {
#pragma unroll
              for (int e = 0; e < 32; e++) atomicAdd(prow + e * T, fmaf(v[e].x, v[e].x, v[e].y * v[e].y));
            }
          }
        } else {
          // general path: full filtercorr table and/or a limited bin range (fft1.c:4115-4131)
          const float* fcp = p.filtercorr + (size_t)t * MM + 2 * c;
#pragma unroll
          for (int e = 0; e < 32; e++) {
            const int k = t + T * e;
            if (k >= p.first_point && k <= p.last_point) {
              const float2 f = *reinterpret_cast<const float2*>(fcp + (size_t)e * (T * MM));
              const float re = v[e].y * f.x - v[e].x * f.y;
              const float im = v[e].x * f.x + v[e].y * f.y;
              v[e] = make_float2(im, re);
              const float pw = re * re + im * im;
              if (prow) {
                if (plain) prow[e * T] = pw;
                else atomicAdd(prow + e * T, pw);
              }
            } else if (rows && NCH == 1) {
              prow[e * T] = 0.0f;
            }
          }
        }
      }
#if defined(LB_EXP) && (LB_EXP & 2)
      if (v[0].x == 1.2345f) outb[0] = v[5].y;
#elif defined(LB_EXP) && (LB_EXP & 1)
      {
        float* ob = p.out + ((p.out_pa + (uint32_t)b * (uint32_t)(MM * N)) & p.out_mask) + (size_t)t * 2 + (size_t)c * N * 2;
#pragma unroll
        for (int e = 0; e < 32; e++) lb_store_stream(reinterpret_cast<float2*>(ob + (size_t)e * (T * 2)), make_float2(v[e].y, v[e].x));
      }
#else
#pragma unroll
      for (int e = 0; e < 32; e++) lb_store_stream(reinterpret_cast<float2*>(outb + (size_t)e * (T * MM)), make_float2(v[e].y, v[e].x));
#endif
    }
    if (!have_next) break;
    w = nw;
    b = nb;
    b0 = nb0;
    b1 = nb1;
  }
}

template <int LOG2N>
constexpr size_t fft1_fused_smem()
{
  return sizeof(float2) * Plan32<LOG2N>::XCH + sizeof(float4) * Plan32<LOG2N>::TAB1 + sizeof(float) * (1 << LOG2N);
}

}  // namespace lb
