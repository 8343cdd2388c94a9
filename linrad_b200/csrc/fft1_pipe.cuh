// fft1_pipe.cuh -- fft1 for 2^15 <= N <= 2^20 as ONE persistent, asynchronously staged kernel.
//
// Same four-step mathematics as fft1_large.cuh (N = N1*N2, column transforms over n1, inter-step
// twiddle W_N^(n2*k1), row transforms over n2, bin k = k1 + N1*k2), same reference functions
// replaced (fft1win_* + bulk_of_dif/dit of fft1.c:684-2247 / fft0.c:161-195,1590-1769, the
// direction flip fft1.c:3660-3680, fft1_c fft1.c:4115-4200), rebuilt so that nothing waits:
//
//   * one launch per call.  Work items "columns of transform b" (role A) and "rows of transform b"
//     (role B) sit in one queue, A(b) ahead of B(b) by `lag` transforms; CTAs claim items in queue
//     order with an atomic counter, so a dependency always points to an item that a running CTA
//     already holds (no deadlock, no co-residency requirement).  B(b) waits for the columns of b
//     (doneA[b]), A(b) waits until the rows of b-nslots have left the slot it writes (doneB).
//     The intermediate Y lives in a ring of `nslots` transforms (a few MB): it never leaves L2.
//   * Y is kept TRANSPOSED, Y[n2][k1].  The one transposition the four-step scheme needs is done
//     where the data is smallest: the raw int16/int32 tile (TA adjacent columns x N1 rows) is
//     staged in shared memory by cp.async (16-byte pieces, padded pitch) while the previous item
//     is computed, and each thread picks the rows of its own column out of it.
//   * role A: a column transform of N1 = 32*T1 points is done by T1 <= 32 lanes of ONE warp
//     (32 points per lane, two passes, warp-private exchange, __syncwarp only): the eight warps of
//     a CTA never wait for each other inside an item.  Results leave as 128-byte runs (lanes along
//     k1).  The window table is stored transposed (wT[n2][n1]) so that it is read the same way.
//   * role B: 256/T2 adjacent rows (k1) side by side with the row index fastest across lanes; the
//     Y tile [N2][TB] arrives by ONE TMA tensor load (cp.async.bulk.tensor.3d, mbarrier
//     complete_tx) issued while the previous item is computed; bins k1..k1+TB-1 of one k2 leave
//     as one 128-byte run, either as streaming stores or (one channel) staged and written by TMA
//     tensor stores (cp.async.bulk.tensor.3d shared -> global).
//   * the new timf1 bytes of a later transform are pulled into L2 by cp.async.bulk.prefetch.L2.
#pragma once
#include <cuda.h>
#ifdef LB_PIPE_STATS
#include <cstdio>
#endif
#include "fft1_fused.cuh"

#ifndef LB_PIPE_THREADS
#define LB_PIPE_THREADS 256     // threads per CTA = 32-point slices per item
#endif

namespace lb {

struct Fft1PipeK {
  Fft1K k;               // same parameter block as the other fft1 kernels
  float2* Y;             // intermediate: [nslots * NCH][N2][N1]
  const void* wT;        // transposed window: IQ float[N2][N1] with (-1)^n folded in; real input float2[N2][N1]
  const float2* Wn1;     // exp(-2 pi i m / N1)
  const float2* Wn2;     // exp(-2 pi i m / N2)
  const float2* Wbig;    // exp(-2 pi i m / N)
  int* sync;             // [0] queue head, [1] error flag, [2 .. 2+nblocks) doneA, [2+nblocks .. 2+2 nblocks) doneB
  int nslots;            // transforms in the Y ring
  int lag;               // B(b) is queued `lag` transforms behind A(b)
  int tma_in;            // role B input by TMA tensor load (else cp.async)
  int tma_out;           // role B output by TMA tensor store (one channel / zbuf only; else streaming stores)
  int prefetch_ahead;    // L2 prefetch distance in transforms (0 = off)
  uint32_t out_blk0;     // tma_out: index of the call's first output block in the tensor map's outermost dimension
  uint32_t out_nblk;     // ... and the extent of that dimension (ring wrap)
};

// ---- small PTX wrappers -----------------------------------------------------------------------
LB_D void cp_async16(void* sdst, const void* gsrc)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
// the executing thread's earlier cp.async copies arrive on the mbarrier when they have landed
LB_D void cp_async_arrive_noinc(uint64_t* bar)
{
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
LB_D void tma_load_3d(void* sdst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(sdst)),
               "l"(map), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
LB_D void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* ssrc)
{
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// fire-and-forget float add into L2 (RED, no return value)
LB_D void red_add(float* p, float v)
{
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
LB_D int ld_relaxed(const int* p)
{
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
LB_D int ld_acquire(const int* p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until *ctr >= target.  Never hangs: after about a second (or once another waiter has given
// up) the error flag is set and the wait returns; the host reports LB200_ERR_CUDA for the call.
LB_D void pipe_wait(const int* ctr, int target, int* err)
{
  if (ld_relaxed(ctr) >= target) return;
  const long long t0 = clock64();
  while (ld_relaxed(ctr) < target) {
    __nanosleep(64);
    if (*reinterpret_cast<volatile int*>(err)) return;
    if (clock64() - t0 > (1ll << 31)) { atomicExch(err, 1); return; }
  }
}

// mbarrier wait that cannot hang either; the waiting warp is suspended by the hardware
LB_D void pipe_mbar_wait(uint64_t* bar, uint32_t parity, int* err)
{
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(a), "r"(parity) : "memory");
  if (done) return;
  const long long t0 = clock64();
  for (uint32_t n = 1;; n++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity), "r"(20000u) : "memory");
    if (done) return;
    if ((n & 255u) == 0) {
      if (*reinterpret_cast<volatile int*>(err)) return;
      if (clock64() - t0 > (1ll << 31)) { atomicExch(err, 2); return; }
    }
  }
}

// ---- geometry -----------------------------------------------------------------------------------
template <int LN1, int LN2, int FMT>
struct PipeCfg {
  static constexpr int N1 = 1 << LN1, N2 = 1 << LN2, N = N1 * N2;
  static constexpr int FRAME = FmtInfo<FMT>::FRAME, NCH = FmtInfo<FMT>::NCH;
  static constexpr bool REAL = FmtInfo<FMT>::REAL;
  static constexpr int NTHREADS = LB_PIPE_THREADS, NWARPS = NTHREADS / 32;
  static constexpr int TILE_BYTES = 32 * NTHREADS * 8;      // one item: 32 points per thread
  // role A: T1 lanes per column, CW columns per warp, TA columns per item
  static constexpr int T1 = N1 / 32, LT1 = LN1 - 5, CW = 32 / T1, TA = NWARPS * CW, TILES_A = N2 / TA;
  static constexpr int Q1 = 32 / T1;                       // pass 0 does Q1 radix-T1 butterflies per lane
  static constexpr int PITCH_A = TA * FRAME + 16;          // bytes; the pad rotates the rows over the banks
  static constexpr int CPR_A = TA * FRAME / 16;            // 16-byte pieces per raw row
  static constexpr int XP = T1 + 2;                        // exchange pitch (float2 slots), 16-byte skew per row
  static constexpr int AREA_A = CW * T1 * XP * 8;          // warp-private exchange bytes
  // role B: T2 threads per row, TB rows per item
  static constexpr int T2 = N2 / 32, LT2 = LN2 - 5, TB = NTHREADS / T2, TILES_B = N1 / TB;
  static constexpr int Q2 = 32 / T2;
  static constexpr int ROUND_B = TILE_BYTES / Q2;           // one exchange round, all rows
  static constexpr int STAGE_B = TILE_BYTES / 2;                   // half an output tile (TMA store rounds)
  static constexpr int BOX_IN = N2 < 256 ? N2 : 256;       // rows per input box
  static constexpr int BOX_OUT = N2 / 2 < 256 ? N2 / 2 : 256;
  static constexpr int IA = TILES_A * NCH, IB = TILES_B * NCH;   // items per transform
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  static constexpr int IN_BYTES = (cmax(N1 * PITCH_A, TILE_BYTES) + 127) & ~127;
  static constexpr int WORK_BYTES = (cmax(cmax(NWARPS * AREA_A, ROUND_B), STAGE_B) + 127) & ~127;
  static constexpr int cmin(int a, int b) { return a < b ? a : b; }
  static constexpr int ROUNDS_PER_BARRIER = cmin(Q2, cmax(1, WORK_BYTES / ROUND_B));
  static_assert(Q2 % ROUNDS_PER_BARRIER == 0, "exchange rounds");
  static constexpr int TAB_BYTES = (T1 + T2) * 5 * 8;
  static constexpr int SMEM = IN_BYTES + WORK_BYTES + TAB_BYTES;
  static constexpr int MINB = cmax(1, (65536 / (128 * NTHREADS)) < (227 * 1024 / (SMEM + 1024)) ? (65536 / (128 * NTHREADS)) : (227 * 1024 / (SMEM + 1024)));
};

struct PipeItem {
  int role;              // 0 = columns (A), 1 = rows (B), -1 = queue empty
  int b;                 // transform within the call
  int j;                 // tile * NCH + channel
  int ready;             // input may be fetched right away
};

// queue position -> item.  Phases: `lag` phases of A only, then A(p) + B(p-lag), then B only.
LB_HD PipeItem pipe_decode(int i, int nb, int lag, int IA, int IB)
{
  PipeItem it;
  it.ready = 0;
  const int p1 = nb < lag ? nb : lag;
  const int p2 = nb > lag ? nb - lag : 0;
  const int n1 = p1 * IA, n2 = p2 * (IA + IB), n3 = p1 * IB;
  if (i < n1) {
    it.role = 0; it.b = i / IA; it.j = i - it.b * IA;
  } else if (i < n1 + n2) {
    i -= n1;
    const int ph = i / (IA + IB);
    const int j = i - ph * (IA + IB);
    if (j < IA) { it.role = 0; it.b = lag + ph; it.j = j; }
    else { it.role = 1; it.b = ph; it.j = j - IA; }
  } else if (i < n1 + n2 + n3) {
    i -= n1 + n2;
    const int ph = i / IB;
    it.role = 1; it.b = p2 + ph; it.j = i - ph * IB;
  } else {
    it.role = -1; it.b = 0; it.j = 0;
  }
  return it;
}

// ---- the two-pass transform of M = 32*T points held by T lanes (32 points each) -----------------
// pass 0: Q = 32/T radix-T butterflies without twiddles (lane t, butterfly q: v[q + r*Q]);
// exchange: butterfly q of lane t owns row t of a T x T matrix, the next pass needs column t:
//           u[q*T + e] = (row e, column t);
// pass 1: one radix-32 butterfly with w = exp(-2 pi i t / M) (radix32_gen).
// The exchange is the only part that differs between the roles (warp-private / CTA-wide).

// role A exchange, one column: `area` holds T rows of XP = T+2 float2 slots
template <int T>
LB_HD void colx_store(const float2 (&v)[32], float2* area, int t, int q)
{
  constexpr int Q = 32 / T;
  float2* row = area + t * (T + 2);
#pragma unroll
  for (int r = 0; r < T; r += 2) {
    const float2 a = v[q + r * Q], b = v[q + (r + 1) * Q];
    *reinterpret_cast<float4*>(row + r) = make_float4(a.x, a.y, b.x, b.y);
  }
}
template <int T>
LB_HD void colx_load(float2 (&u)[32], const float2* area, int t, int q)
{
#pragma unroll
  for (int e = 0; e < T; e++) u[q * T + e] = area[e * (T + 2) + t];
}
// role B exchange: TB rows interleaved element by element (row index fastest)
template <int T, int TB>
LB_HD void rowx_store(const float2 (&v)[32], float2* buf, int t, int r, int q)
{
  constexpr int Q = 32 / T;
#pragma unroll
  for (int rr = 0; rr < T; rr++) buf[(t * T + rr) * TB + r] = v[q + rr * Q];
}
template <int T, int TB>
LB_HD void rowx_load(float2 (&u)[32], const float2* buf, int t, int r, int q)
{
#pragma unroll
  for (int e = 0; e < T; e++) u[q * T + e] = buf[(e * T + t) * TB + r];
}

#ifdef __CUDACC__
// raw frame at a shared-memory address -> the complex point of channel c (same conversions as load_iq)
template <int FMT>
LB_D float2 cvt_raw(const unsigned char* p, int c)
{
  return load_iq<FMT>(p, 0u, c);
}

template <int LN1, int LN2, int FMT>
__global__ void __launch_bounds__(PipeCfg<LN1, LN2, FMT>::NTHREADS, PipeCfg<LN1, LN2, FMT>::MINB)
fft1_large_pipe_kernel(const Fft1PipeK q, const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapOut)
{
  using C = PipeCfg<LN1, LN2, FMT>;
  constexpr int N1 = C::N1, N2 = C::N2, N = C::N, NCH = C::NCH, MM = 2 * NCH, FRAME = C::FRAME;
  constexpr int T1 = C::T1, CW = C::CW, TA = C::TA, Q1 = C::Q1;
  constexpr int T2 = C::T2, TB = C::TB, Q2 = C::Q2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* const in = smem_raw;
  unsigned char* const work = smem_raw + C::IN_BYTES;
  float2* const wbt = reinterpret_cast<float2*>(smem_raw + C::IN_BYTES + C::WORK_BYTES);
  __shared__ uint64_t bar_in;
  __shared__ PipeItem items[2];
  __shared__ int a_arrived;
  const Fft1K& p = q.k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* const head = q.sync;
  int* const err = q.sync + 1;
  int* const doneA = q.sync + 2;
  int* const doneB = q.sync + 2 + p.nblocks;
  const int nb = p.nblocks;
  const int total = nb * (C::IA + C::IB);

  // last-pass twiddles: the five exact binary powers of w = exp(-2 pi i t / M) per lane position
  for (int i = tid; i < T1 * 5; i += C::NTHREADS) wbt[i] = q.Wn1[(i / 5) << (i % 5)];
  for (int i = tid; i < T2 * 5; i += C::NTHREADS) wbt[T1 * 5 + i] = q.Wn2[(i / 5) << (i % 5)];
  if (tid == 0) {
    a_arrived = 0;
    mbar_init(&bar_in, C::NTHREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- helpers ---------------------------------------------------------------------------------
  auto claim = [&](PipeItem& dst) {               // thread 0 only
    const int i = atomicAdd(head, 1);
    PipeItem it = i < total ? pipe_decode(i, nb, q.lag, C::IA, C::IB) : pipe_decode(total, nb, q.lag, C::IA, C::IB);
    if (it.role == 0) it.ready = 1;
    else if (it.role == 1) it.ready = ld_relaxed(doneA + it.b) >= C::IA * C::NWARPS ? 1 : 0;
    dst = it;
  };
  auto slot_of = [&](int b) { return b % q.nslots; };
  // fetch the input of an item into `in`; called by all threads, the item's dependency is satisfied
  auto issue_load = [&](const PipeItem& it) {
    const int tile = it.j / NCH;
    if (it.role == 0) {
      const uint32_t start = p.ref0 + (uint32_t)it.b * p.blockbytes - p.pre_bytes;
      const uint32_t base = start + (uint32_t)(tile * TA) * FRAME;
#pragma unroll 4
      for (int ch = tid; ch < N1 * C::CPR_A; ch += C::NTHREADS) {
        const int row = ch / C::CPR_A, cc = ch - row * C::CPR_A;
        const uint32_t off = (base + (uint32_t)row * (uint32_t)(N2 * FRAME) + (uint32_t)cc * 16u) & p.ring_mask;
        cp_async16(in + row * C::PITCH_A + cc * 16, p.timf1 + off);
      }
      cp_async_arrive_noinc(&bar_in);
      // the new bytes of a later transform -> L2, one slice per column tile
      if (tid == 0 && q.prefetch_ahead > 0 && it.j % NCH == 0 && it.b + q.prefetch_ahead < nb) {
        const uint32_t piece = (p.blockbytes / C::TILES_A) & ~15u;
        if (piece >= 16u) {
          const uint32_t s2 = p.ref0 + (uint32_t)(it.b + q.prefetch_ahead) * p.blockbytes - p.pre_bytes + (uint32_t)N * FRAME - p.blockbytes;
          l2_prefetch_span(p.timf1, p.ring_mask, s2 + (uint32_t)tile * piece, piece);
        }
      }
    } else {
      const int c = it.j - tile * NCH;
      const int plane = slot_of(it.b) * NCH + c;
      if (q.tma_in) {
        if (tid == 0) {
          // Y was written through the generic proxy (other SMs, observed by an acquire): order it
          // before the TMA unit's reads
          asm volatile("fence.proxy.async;" ::: "memory");
          mbar_expect_tx(&bar_in, (uint32_t)C::TILE_BYTES);
#pragma unroll
          for (int bx = 0; bx < N2 / C::BOX_IN; bx++)
            tma_load_3d(in + bx * (C::BOX_IN * TB * 8), &mapY, 2 * tile * TB, bx * C::BOX_IN, plane, &bar_in);
        } else {
          mbar_arrive(&bar_in);
        }
      } else {
        constexpr int CPR = TB * 8 / 16;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(q.Y + (size_t)plane * N + (size_t)tile * TB);
#pragma unroll 4
        for (int ch = tid; ch < N2 * CPR; ch += C::NTHREADS) {
          const int row = ch / CPR, cc = ch - row * CPR;
          cp_async16(in + row * (TB * 8) + cc * 16, src + (size_t)row * (N1 * 8) + cc * 16);
        }
        cp_async_arrive_noinc(&bar_in);
      }
    }
  };

  if (tid == 0) {
    claim(items[0]);
    items[0].ready = 0;                          // the first item goes through the deferred path below
  }
  __syncthreads();
  PipeItem cur = items[0];
  if (cur.role >= 0) {
    if (cur.role == 1 && tid == 0) pipe_wait(doneA + cur.b, C::IA * C::NWARPS, err);
    __syncthreads();
    issue_load(cur);
  }
#ifdef LB_PIPE_STATS
  long long st_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // thread 0: cycles in [A wait, A rest, B wait, B rest, deferred, claim, slot wait] + counts
  long long st_t = clock64();
  const long long st_begin = st_t;
#define LB_ST(i) do { if (tid == 0) { const long long n_ = clock64(); st_[i] += n_ - st_t; st_t = n_; } } while (0)
#define LB_CNT(i) do { if (tid == 0) st_[i] += 1; } while (0)
#else
#define LB_ST(i) do { } while (0)
#define LB_CNT(i) do { } while (0)
#endif
  uint32_t par = 0;
  int s = 0;
  bool stores_pending = false;                   // thread 0: TMA stores may still be reading `work`

  while (cur.role >= 0) {
    LB_ST(9);
    if (tid == 0) claim(items[s ^ 1]);
    LB_ST(5);
    const int tile = cur.j / NCH;
    const int c = cur.j - tile * NCH;
    float2 v[32];
    if (cur.role == 0) {
      // =============================== role A: TA columns of transform cur.b ===================
      const int t = lane & (T1 - 1), cw = lane >> C::LT1;
      const int col = warp * CW + cw;
      const int n2 = tile * TA + col;
      // window (transposed table: lanes along n1), in flight during the wait
      float wv[C::REAL ? 64 : 32];
      if (C::REAL) {
        const float2* wp = reinterpret_cast<const float2*>(q.wT) + (size_t)n2 * N1 + t;
#pragma unroll
        for (int e = 0; e < 32; e++) {
          const float2 w2 = __ldg(wp + e * T1);
          wv[2 * e] = w2.x;
          wv[2 * e + 1] = w2.y;
        }
      } else {
        const float* wp = reinterpret_cast<const float*>(q.wT) + (size_t)n2 * N1 + t;
#pragma unroll
        for (int e = 0; e < 32; e++) wv[e] = __ldg(wp + e * T1);
      }
      LB_ST(1);
      pipe_mbar_wait(&bar_in, par, err);
      LB_ST(0);
      LB_CNT(7);
      par ^= 1;
      {
        const unsigned char* rp = in + t * C::PITCH_A + col * FRAME;
        const float dq = p.direction < 0 ? -1.0f : 1.0f;
#pragma unroll
        for (int e = 0; e < 32; e++) {
          const float2 sm = cvt_raw<FMT>(rp + e * (T1 * C::PITCH_A), c);
          if (C::REAL) v[e] = make_float2(sm.x * wv[2 * e], sm.y * wv[2 * e + 1]);
          else v[e] = make_float2(sm.x * wv[e], sm.y * (wv[e] * dq));
        }
      }
      if (tid == 0 && stores_pending) {          // role B's TMA stores read `work`: done before anybody rewrites it
        bulk_wait_read();
        stores_pending = false;
      }
      __syncthreads();                            // the raw tile is consumed; items[s^1] is visible
      const PipeItem nxt = items[s ^ 1];
      if (nxt.role >= 0 && nxt.ready) issue_load(nxt);
      // inter-step twiddle W_N^(n2*(t+T1*e)) = base * step^e, step given by exact binary powers
      const float2 tw_base = __ldg(q.Wbig + n2 * t);
      float2 tw_sb[5];
#pragma unroll
      for (int j = 0; j < 5; j++) tw_sb[j] = __ldg(q.Wbig + ((n2 * T1) << j));
      // ---- column transform, warp-private
      pass0<T1>(v);
      {
        float2* area = reinterpret_cast<float2*>(work + warp * C::AREA_A) + cw * (T1 * C::XP);
        float2 u[32];
#pragma unroll
        for (int qq = 0; qq < Q1; qq++) {
          colx_store<T1>(v, area, t, qq);
          __syncwarp();
          colx_load<T1>(u, area, t, qq);
          __syncwarp();
        }
#pragma unroll
        for (int e = 0; e < 32; e++) v[e] = u[e];
      }
      {
        float2 wb[5];
#pragma unroll
        for (int j = 0; j < 5; j++) wb[j] = wbt[t * 5 + j];
        radix32_gen(v, wb);
      }
      apply_power_twiddles<32>(v, tw_base, tw_sb);
      // ---- the slot must have been read by the rows of transform b - nslots
      LB_ST(1);
      if (cur.b >= q.nslots) {
        if (lane == 0) pipe_wait(doneB + (cur.b - q.nslots), C::IB, err);
        __syncwarp();
      }
      LB_ST(6);
      {
        float2* Yp = q.Y + (size_t)(slot_of(cur.b) * NCH + c) * N + (size_t)n2 * N1 + t;
#pragma unroll
        for (int e = 0; e < 32; e++) Yp[e * T1] = v[e];
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        const int old = atomicAdd(&a_arrived, 1);
        __threadfence_block();
        if (old == C::NWARPS - 1) {
          a_arrived = 0;
          __threadfence();
          atomicAdd(doneA + cur.b, C::NWARPS);
        }
      }
      LB_ST(1);
      if (nxt.role >= 0 && !nxt.ready) {
        if (nxt.role == 1 && tid == 0) pipe_wait(doneA + nxt.b, C::IA * C::NWARPS, err);
        __syncthreads();
        issue_load(nxt);
      }
      LB_ST(4);
      cur = nxt;
    } else {
      // =============================== role B: TB rows of transform cur.b ======================
      const int r = tid & (TB - 1), t = tid / TB;
      LB_ST(3);
      pipe_mbar_wait(&bar_in, par, err);
      LB_ST(2);
      LB_CNT(8);
      par ^= 1;
      {
        const float2* ip = reinterpret_cast<const float2*>(in) + t * TB + r;
#pragma unroll
        for (int e = 0; e < 32; e++) v[e] = ip[e * (T2 * TB)];
      }
      if (tid == 0 && stores_pending) {
        bulk_wait_read();
        stores_pending = false;
      }
      __syncthreads();                            // the Y tile is in registers; items[s^1] is visible
      const PipeItem nxt = items[s ^ 1];
      if (tid == 0) {                             // this tile's share of the slot may be overwritten
        atomicAdd(doneB + cur.b, 1);
      }
      if (nxt.role >= 0 && nxt.ready) issue_load(nxt);
      // ---- row transforms
      pass0<T2>(v);
      {
        float2* buf = reinterpret_cast<float2*>(work);
        float2 u[32];
        // a round moves T2 values per thread; as many rounds as fit into `work` share one pair of barriers
        constexpr int RPB = C::ROUNDS_PER_BARRIER, RELEMS = C::ROUND_B / 8;
#pragma unroll
        for (int q0 = 0; q0 < Q2; q0 += RPB) {
          if (q0 > 0) __syncthreads();            // the previous rounds have been read
#pragma unroll
          for (int j = 0; j < RPB; j++) rowx_store<T2, TB>(v, buf + j * RELEMS, t, r, q0 + j);
          __syncthreads();
#pragma unroll
          for (int j = 0; j < RPB; j++) rowx_load<T2, TB>(u, buf + j * RELEMS, t, r, q0 + j);
        }
#pragma unroll
        for (int e = 0; e < 32; e++) v[e] = u[e];
      }
      {
        float2 wb[5];
#pragma unroll
        for (int j = 0; j < 5; j++) wb[j] = wbt[T1 * 5 + t * 5 + j];
        radix32_gen(v, wb);
      }
      // ---- epilogue: bin k = k1 + N1*k2, k2 = t + T2*e (fft1_b direction flip, fft1_c)
      const int b = cur.b;
      const int k1 = tile * TB + r;
      const bool use_tma_out = q.tma_out != 0;
      if (p.zbuf) {
        // real input: the plain packed spectrum Z, finished by fft1_real_post_kernel
        if (!use_tma_out) {
          float2* zp = p.zbuf + ((size_t)(b - p.zb_first) * NCH + c) * N + k1 + (size_t)t * N1;
#pragma unroll
          for (int e = 0; e < 32; e++) zp[(size_t)e * (T2 * N1)] = v[e];
        }
      } else {
        const int group_size = p.power_rows ? 1 : p.avg1num;
        const int c0 = p.power_rows ? 0 : p.counter0;
        const int g = (b + c0) / group_size;
        float* rowp = (p.sumsq && !p.power_rows) ? p.sumsq + ((p.sumsq_pa + (uint32_t)g * (uint32_t)N) & p.sumsq_mask) : nullptr;
        float* outb = p.out + ((p.out_pa + (uint32_t)b * (uint32_t)(MM * N)) & p.out_mask);
        // the tile's bins are k1 + N1*k2 for all k2: inside [first_point, last_point] and clear of the
        // tapered edge bins (k < fc_edge needs k2 == 0, k >= N - fc_edge needs k2 == N2-1) for all
        // but the outermost tiles of a full-range set-up -> one uniform real gain, no table, no tests
        const int klo = tile * TB, khi = tile * TB + TB - 1;
        const bool fast = p.fc_mode == 1 && !p.power_rows && klo >= p.first_point && khi + N1 * (N2 - 1) <= p.last_point &&
                          p.fc_edge <= N1 && klo >= p.fc_edge && khi < N1 - p.fc_edge;
        if (fast) {
          const float gain = p.fc_gain;           // fft1.c:4121-4125 with filtercorr = (gain, 0)
          if (p.direction < 0) {
#pragma unroll
            for (int e = 0; e < 32; e++) v[e] = make_float2(v[e].y * gain, v[e].x * gain);
          } else {
#pragma unroll
            for (int e = 0; e < 32; e++) v[e] = make_float2(v[e].x * gain, -v[e].y * gain);
          }
          if (rowp) {
            float* rp = rowp + k1 + N1 * t;
#pragma unroll
            for (int e = 0; e < 32; e++) red_add(rp + e * (N1 * T2), fmaf(v[e].x, v[e].x, v[e].y * v[e].y));
          }
          if (!use_tma_out) {
            float2* op = reinterpret_cast<float2*>(outb + (size_t)(k1 + N1 * t) * MM + 2 * c);
#pragma unroll
            for (int e = 0; e < 32; e++) __stcs(op + (size_t)e * (N1 * T2 * NCH), v[e]);
          }
        } else {
          float* prow = p.power_rows ? p.power_rows + (size_t)b * N : nullptr;
#pragma unroll
          for (int e = 0; e < 32; e++) {
            const int k = k1 + N1 * (t + T2 * e);
            const float2 z = v[e];
            float2 ov = p.direction < 0 ? make_float2(z.y, z.x) : make_float2(z.x, -z.y);
            if (p.fc_mode != 0) {
              const bool inr = (k >= p.first_point) && (k <= p.last_point);
              if (inr) {
                float2 f;
                if (p.fc_mode == 2 || k < p.fc_edge || k >= N - p.fc_edge)
                  f = *reinterpret_cast<const float2*>(p.filtercorr + (size_t)k * MM + 2 * c);
                else
                  f = make_float2(p.fc_gain, 0.0f);
                const float re = ov.x * f.x - ov.y * f.y;      // fft1.c:4121-4125
                const float im = ov.y * f.x + ov.x * f.y;
                ov = make_float2(re, im);
                const float pw = fmaf(re, re, im * im);
                if (prow) {
                  if (NCH == 1) prow[k] = pw;
                  else red_add(prow + k, pw);                  // two channel items add into the host-zeroed row
                } else if (rowp) {
                  red_add(rowp + k, pw);
                }
              } else if (prow && NCH == 1) {
                prow[k] = 0.0f;
              }
            }
            v[e] = ov;
            if (!use_tma_out) __stcs(reinterpret_cast<float2*>(outb + (size_t)k * MM + 2 * c), ov);
          }
        }
      }
      if (use_tma_out) {
        // two rounds of N2/2 bins-of-k2 each: [k2][TB] tile in `work`, written by the TMA unit
        const int plane = p.zbuf ? (b - p.zb_first) * NCH + c
                                 : (int)((q.out_blk0 + (uint32_t)b) % q.out_nblk);
        float2* st = reinterpret_cast<float2*>(work) + t * TB + r;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (h > 0 && tid == 0) bulk_wait_read();           // round 0 has left `work`
          __syncthreads();                                   // exchange reads / round 0 are over
#pragma unroll
          for (int e = 0; e < 16; e++) st[e * (T2 * TB)] = v[16 * h + e];
          fence_async_smem();
          __syncthreads();
          if (tid == 0) {
#pragma unroll
            for (int bx = 0; bx < (N2 / 2) / C::BOX_OUT; bx++)
              tma_store_3d(&mapOut, 2 * tile * TB, h * (N2 / 2) + bx * C::BOX_OUT, plane, work + bx * (C::BOX_OUT * TB * 8));
            bulk_commit();
            stores_pending = true;
          }
        }
      }
      LB_ST(3);
      if (nxt.role >= 0 && !nxt.ready) {
        if (nxt.role == 1 && tid == 0) pipe_wait(doneA + nxt.b, C::IA * C::NWARPS, err);
        __syncthreads();
        issue_load(nxt);
      }
      LB_ST(4);
      cur = nxt;
    }
    s ^= 1;
  }
  if (tid == 0) bulk_wait_all();                  // shared memory must outlive the last TMA store
#ifdef LB_PIPE_STATS
  if (tid == 0 && (blockIdx.x % 37) == 0 && nb >= 60)
    printf("PIPESTAT cta %d total %lld Await %lld Arest %lld Bwait %lld Brest %lld deferred %lld claim %lld slotwait %lld other %lld nA %lld nB %lld\n", (int)blockIdx.x,
           clock64() - st_begin, st_[0], st_[1], st_[2], st_[3], st_[4], st_[5], st_[6], st_[9], st_[7], st_[8]);
#endif
}
#endif  // __CUDACC__

}  // namespace lb
