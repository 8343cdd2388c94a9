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

// ---- asynchronous output: results leave through the exchange buffer and the TMA unit ---------
// STG costs the SM one 32-byte sector per clock (measured: 4.2 cycles per 128-byte line), the
// issuing warps stall on the LSU queue meanwhile, and a two-channel transform fills only half of
// every sector (8 bytes of each 16-byte [re1,im1,re2,im2] slot).  So once the last exchange has
// been read the threads write their results into the exchange buffer in the layout of fft1_float
// and one thread hands the block to the TMA unit (cp.async.bulk shared -> global).
// The copy drains while the CTA loads and starts the next transform.  Two-channel formats keep
// direct stores (see the comment at the store).
LB_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
LB_D void bulk_store(void* gdst, const void* ssrc, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
// TMA bulk load global -> shared, completion counted in bytes on an mbarrier
LB_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
LB_D void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// the N-frame span starting at ring offset `start` (16-byte aligned) -> dst, split at the ring wrap
LB_D void bulk_load_span(void* dst, const uint8_t* ring, uint32_t mask, uint32_t start, uint32_t bytes, uint64_t* bar)
{
  const uint32_t a = start & mask;
  const uint32_t size = mask + 1u;
  const uint32_t n1 = size - a < bytes ? size - a : bytes;
  mbar_expect_tx(bar, bytes);
  bulk_load(dst, ring + a, n1, bar);
  if (n1 < bytes) bulk_load(reinterpret_cast<unsigned char*>(dst) + n1, ring, bytes - n1, bar);
}
LB_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// ---- 2-CTA cluster helpers (two-channel output path)
LB_D uint32_t cluster_ctarank()
{
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address -> the same offset in CTA `rank` of the cluster (shared::cluster window)
LB_D uint32_t cluster_map(const void* p, uint32_t rank)
{
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(rank));
  return r;
}
LB_D void st_cluster_f2(uint32_t addr, float2 v)
{
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
LB_D void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
LB_D void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
LB_D void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
LB_D void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// |z|^2 of elements [E0, E1) -> fft1_sumsq row / power row, after the filtercorr multiply (fft1_c,
// fft1.c:4115-4200).  v holds (im, re).
template <int FC, int NCH, int T, int E0, int E1>
LB_D void fused_epilogue(float2 (&v)[32], const Fft1K& p, float* prow, bool plain, bool rows, int t, int c)
{
  constexpr int MM = 2 * NCH;
  if (FC == FC_FOLDED) {                     // host guarantees the full bin range here
    if (prow) {
      if (plain) {
#pragma unroll
        for (int e = E0; e < E1; e++) prow[e * T] = fmaf(v[e].x, v[e].x, v[e].y * v[e].y);
      } else {
#pragma unroll
        for (int e = E0; e < E1; e++) atomicAdd(prow + e * T, fmaf(v[e].x, v[e].x, v[e].y * v[e].y));
      }
    }
  } else if (FC == FC_TABLE) {
    // general path: full filtercorr table and/or a limited bin range (fft1.c:4115-4131)
    const float* fcp = p.filtercorr + (size_t)t * MM + 2 * c;
#pragma unroll
    for (int e = E0; e < E1; e++) {
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

template <int LOG2N, int FMT, int FC>
__global__ void __launch_bounds__(1 << (LOG2N - 5), 512 >> (LOG2N - 5))
fft1_fused_kernel(const Fft1K p)
{
  using P = Plan32<LOG2N>;
  constexpr int N = P::N, T = P::T;
  constexpr int FRAME = FmtInfo<FMT>::FRAME, NCH = FmtInfo<FMT>::NCH, MM = 2 * NCH;
  constexpr int CHB = FRAME / NCH;                         // bytes of one channel's IQ pair
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* xch = reinterpret_cast<float2*>(smem_raw);
  float4* tab1 = reinterpret_cast<float4*>(smem_raw + sizeof(float2) * P::XCH);
  float* wsm = reinterpret_cast<float*>(smem_raw + sizeof(float2) * P::XCH + sizeof(float4) * P::TAB1);
  const int t = threadIdx.x;

  // pass-1 twiddle table [pair][k] and the window table -> shared memory; last-pass exact powers
  // -> registers
  // int16 one-channel IQ: a span of N frames is 4N contiguous bytes, exactly the size of the
  // window table.  Then the table stays in global memory (it lives in L1) and its place takes the
  // raw frames of the NEXT transform, fetched by the TMA unit while this one is computed.
  const bool stage = (FMT == FMT_I16_1CH) && p.stage_raw;
  if (P::NPASS == 3)
    for (int i = t; i < P::TAB1; i += T) tab1[i] = p.tab1[i];
  if (!stage)
    for (int i = t; i < N; i += T) wsm[i] = p.wtab[i];
  float2 wb[5];
#pragma unroll
  for (int j = 0; j < 5; j++) wb[j] = p.Wn[t << j];
  const float* wtab = wsm + t;
  // split-phase barriers (arrive when done with a buffer, wait right before it is overwritten):
  // [0] exchange 1 has been read, [1] the last exchange has been read, [2] a staging round has
  // been written (consumer: thread 0), [3] the TMA unit has read the staged round (producer: thread 0)
  // [4] the staged raw span has landed (TMA, bytes), [5] it has been read by everybody
  __shared__ uint64_t bars[6];
  if (t == 0) {
    mbar_init(&bars[0], T);
    mbar_init(&bars[1], T);
    mbar_init(&bars[2], T);
    mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1);
    mbar_init(&bars[5], T);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t par0 = 0, par1 = 0, par2 = 0, par3 = 0, par4 = 0, par5 = 0;
  bool staged = false;                       // the TMA unit may still be reading the exchange buffer
  bool first_xch = true;
  __syncthreads();

  const int group_size = p.power_rows ? 1 : p.avg1num;
  const int c0 = p.power_rows ? 0 : p.counter0;
  const int ngroups = (c0 + p.nblocks + group_size - 1) / group_size;
  const uint32_t span = (uint32_t)N * FRAME;
  // work item = (averaging group, channel).  The two channels of a group go to neighbouring CTAs:
  // they read the same timf1 frames at the same time (one DRAM fetch, the second CTA hits L2).
  const int nwork = ngroups * NCH;
  if (stage && t == 0 && (int)blockIdx.x < nwork) {
    int fb = ((int)blockIdx.x / NCH) * group_size - c0;
    if (fb < 0) fb = 0;
    bulk_load_span(wsm, p.timf1, p.ring_mask, p.ref0 + (uint32_t)fb * p.blockbytes - p.pre_bytes, span, &bars[4]);
  }
  for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int g = w / NCH;
    const int c = w - g * NCH;
    int b0 = g * group_size - c0;
    int b1 = b0 + group_size;
    if (b0 < 0) b0 = 0;
    if (b1 > p.nblocks) b1 = p.nblocks;
    for (int b = b0; b < b1; b++) {
      const uint32_t start = (p.ref0 + (uint32_t)b * p.blockbytes - p.pre_bytes) & p.ring_mask;
      // the transform this CTA does next: b+1 of the same group, or the first of its next work item
      uint32_t next_start;
      bool has_next = true, next_same_group = (b + 1 < b1);
      if (next_same_group) {
        next_start = start + p.blockbytes;
      } else {
        const int gn = (w + (int)gridDim.x) / NCH;
        int bn = gn * group_size - c0;
        if (bn < 0) bn = 0;
        has_next = (gn < ngroups && bn < p.nblocks);
        next_start = p.ref0 + (uint32_t)bn * p.blockbytes - p.pre_bytes;
      }
      if (t == 0 && c == 0 && has_next && !stage) {
        // pull it into L2: only the new bytes when the overlap half has just been read
        if (next_same_group) l2_prefetch_span(p.timf1, p.ring_mask, start + span, p.blockbytes);
        else l2_prefetch_span(p.timf1, p.ring_mask, next_start, span);
      }
      float* out_block = p.out + ((p.out_pa + (uint32_t)b * (uint32_t)(MM * N)) & p.out_mask);
      const bool wraps = (start + span > p.ring_mask + 1u) || (p.skew_i | p.skew_q);
      float2 v[32];
      // ---- load, int -> float, window (sign and, for FC_FOLDED, gain are in the table)
      if (stage) {
        mbar_wait(&bars[4], par4);               // the TMA unit has delivered this transform's span
        par4 ^= 1;
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(wsm) + t;
        const float* wg = p.wtab + t;
#pragma unroll
        for (int e = 0; e < 32; e++) {
          const uint32_t x = rw[e * T];
          const float2 s = make_float2((float)(short)(x & 0xffffu), (float)(short)(x >> 16));
          const float wv = __ldg(wg + e * T);
          v[e] = p.direction > 0 ? make_float2(s.y * -wv, s.x * wv) : make_float2(s.x * wv, s.y * -wv);
        }
        mbar_arrive(&bars[5]);                   // my reads of the span are done
        if (t == 0 && has_next) {
          mbar_wait(&bars[5], par5);             // everybody's are: fetch the next span into its place
          bulk_load_span(wsm, p.timf1, p.ring_mask, next_start, span, &bars[4]);
        }
        par5 ^= 1;
      } else if (!wraps) {
        const uint8_t* src = p.timf1 + start + (uint32_t)t * FRAME + c * CHB;
        if (p.direction > 0) {
#pragma unroll
          for (int e = 0; e < 32; e++) {
            const float2 s = cvt_iq<FMT>(src + (size_t)e * (T * FRAME));
            const float wv = wtab[e * T];
            v[e] = make_float2(s.y * -wv, s.x * wv);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e++) {
            const float2 s = cvt_iq<FMT>(src + (size_t)e * (T * FRAME));
            const float wv = wtab[e * T];
            v[e] = make_float2(s.x * wv, s.y * -wv);
          }
        }
      } else {
        // ring wrap inside the span, or ui.sample_shift (I and Q words from different frames)
#pragma unroll
        for (int e = 0; e < 32; e++) {
          const uint32_t off = (start + (uint32_t)(t + T * e) * FRAME) & p.ring_mask;
          float2 s;
          if ((FMT == FMT_I16_1CH || FMT == FMT_I32_1CH) && (p.skew_i | p.skew_q))
            s = load_iq_skew<FMT>(p.timf1, p.ring_mask, off, p.skew_i, p.skew_q);
          else
            s = cvt_iq<FMT>(p.timf1 + off + c * CHB);
          const float wv = wtab[e * T];
          v[e] = p.direction > 0 ? make_float2(s.y * -wv, s.x * wv) : make_float2(s.x * wv, s.y * -wv);
        }
      }
      // ---- transform
      pass0<P::R0>(v);
      if (NCH == 2 && !first_xch) {            // the previous transform's last exchange has been read by everybody
        mbar_wait(&bars[1], par1);
        par1 ^= 1;
      }
      first_xch = false;
      if (staged) {                            // the previous transform's output has left the exchange buffer
        if (t == 0) {
          bulk_wait_read();
          mbar_arrive(&bars[3]);
        }
        mbar_wait(&bars[3], par3);
        par3 ^= 1;
        staged = false;
      }
      exch1_store<LOG2N>(v, xch, t);
      __syncthreads();
      exch1_load<LOG2N>(v, xch, t);
      if (P::NPASS == 3) {
        mbar_arrive(&bars[0]);
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
        mbar_wait(&bars[0], par0);             // everybody has read exchange 1
        par0 ^= 1;
        exch2_store<LOG2N>(v, xch, t);
        __syncthreads();
        exch2_load<LOG2N>(v, xch, t);
      }
      mbar_arrive(&bars[1]);                   // my reads of the exchange buffer are done
      if (NCH == 2 && p.cluster2) cluster_arrive();   // ... which my sibling will write into (waited for before the staging)
      radix32_gen(v, wb);
      // ---- epilogue: bin k = t + T*e; v holds (im, re) of the output value.  |z|^2 goes
      // straight to the fft1_sumsq row in L2 (fft1.c:4507-4520 sums the transforms of a group in
      // time order; so do the reductions of one thread on one address): the first transform of a
      // one-channel group stores, everything else is a fire-and-forget RED.ADD.  Two-channel rows
      // are zeroed by the host because the two channel CTAs of a group add into the same row.
      const bool rows = p.power_rows != nullptr;
      float* prow = nullptr;
      if (FC != FC_RAW) {
        prow = rows ? p.power_rows + (size_t)b * N + t
                    : (p.sumsq ? p.sumsq + ((p.sumsq_pa + (uint32_t)g * (uint32_t)N) & p.sumsq_mask) + t : nullptr);
      }
      const bool plain = (NCH == 1) && (rows || (b == b0 && !(g == 0 && p.counter0 > 0)));
      if (FC == FC_FOLDED) {
        if (t < 16) {                          // bins 0..15 and N-16..N-1 carry the taper of fft1.c:4703-4722
          const float2 f = p.edge[t];
          v[0] = make_float2(v[0].x * f.x + v[0].y * f.y, v[0].y * f.x - v[0].x * f.y);
        }
        if (t >= T - 16) {
          const float2 f = p.edge[16 + t - (T - 16)];
          v[31] = make_float2(v[31].x * f.x + v[31].y * f.y, v[31].y * f.x - v[31].x * f.y);
        }
      }
      if (NCH == 1) {
        // one round: bin k at byte 8k of the exchange buffer
        fused_epilogue<FC, NCH, T, 0, 32>(v, p, prow, plain, rows, t, c);
        mbar_wait(&bars[1], par1);             // everybody has read the last exchange
        par1 ^= 1;
#pragma unroll
        for (int e = 0; e < 32; e++) xch[t + T * e] = make_float2(v[e].y, v[e].x);
        fence_async_smem();
        mbar_arrive(&bars[2]);
        if (t == 0) {
          mbar_wait(&bars[2], par2);
#pragma unroll
          for (int q = 0; q < 4; q++) bulk_store(out_block + q * (N / 2), xch + q * (N / 4), N * 2);
          bulk_commit();
        }
        par2 ^= 1;
        staged = true;
      } else {
        // two channels: this CTA owns 8 bytes of every 16-byte slot.  The masked TMA copy
        // (cp.async.bulk ... .cp_mask) was measured: like STG it moves one half-used sector per
        // clock, and with room for only N/2 staged bins the second round has to wait for the
        // first, so plain streaming stores are faster here (profiles/r1_v4_notes.txt).
        fused_epilogue<FC, NCH, T, 0, 32>(v, p, prow, plain, rows, t, c);
        if (p.cluster2) {
          // The two channel CTAs of the group are one cluster and walk the same transforms in step.  Bins
          // [0, N/2) of BOTH channels are collected in rank 0's exchange buffer, bins [N/2, N) in rank 1's, in
          // fft1_float's own slot layout [re1, im1, re2, im2]: each CTA writes half of its values into its own
          // buffer and half into its sibling's (st.shared::cluster), and then hands 8N contiguous bytes to the TMA
          // unit: every global sector is written whole, once, off the LSU path.
          const uint32_t rank = cluster_ctarank();
          cluster_wait();                      // both CTAs have read their last exchange (arrived below)
          const uint32_t lo = cluster_map(xch, 0), hi = cluster_map(xch, 1);
#pragma unroll
          for (int e = 0; e < 16; e++) st_cluster_f2(lo + (uint32_t)((2 * (t + T * e) + c) * 8), make_float2(v[e].y, v[e].x));
#pragma unroll
          for (int e = 16; e < 32; e++) st_cluster_f2(hi + (uint32_t)((2 * (t + T * (e - 16)) + c) * 8), make_float2(v[e].y, v[e].x));
          cluster_arrive();                    // my slots, here and over there, are written
          cluster_wait();                      // ... and so are my sibling's
          if (t == 0) {
            fence_async_smem();
            float* dst = out_block + (size_t)rank * (N / 2) * MM;
#pragma unroll
            for (int q = 0; q < 4; q++) bulk_store(dst + q * (N / 2), xch + q * (N / 4), N * 2);
            bulk_commit();
          }
          staged = true;
        } else {
          float* outb = out_block + (size_t)t * MM + 2 * c;
#pragma unroll
          for (int e = 0; e < 32; e++) lb_store_stream(reinterpret_cast<float2*>(outb + (size_t)e * (T * MM)), make_float2(v[e].y, v[e].x));
        }
      }
    }
  }
  if (t == 0) bulk_wait_all();                 // the exchange buffer must outlive the last copy
}

template <int LOG2N>
constexpr size_t fft1_fused_smem()
{
  return sizeof(float2) * Plan32<LOG2N>::XCH + sizeof(float4) * Plan32<LOG2N>::TAB1 + sizeof(float) * (1 << LOG2N);
}

}  // namespace lb
