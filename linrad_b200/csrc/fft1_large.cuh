// fft1_large.cuh -- fft1 for transforms that do not fit one CTA (2^15 <= N <= 2^20): the
// four-step decomposition N = N1*N2 in two kernels with the intermediate kept in a scratch
// buffer that is sized to stay resident in the 126 MB L2.
//
//   step A (columns): for each n2, FFT over n1 of x[n1*N2+n2] (ring gather, int->float, window
//                     fused), times the inter-step twiddle W_N^(n2*k1)  ->  Y[k1][n2]
//   step B (rows):    for each k1, FFT over n2 of Y[k1][.]  ->  X[k1 + N1*k2], then the same
//                     epilogue as the single-CTA kernel (direction, filtercorr, |z|^2, sumsq)
//
// This replaces what the reference can only do through its double precision path (fft1
// version 20: d_fft1win_dif_chan fft1.c:2145, d_bulk_of_dif fft0.c:197, dif_bigpermute_chan
// fft1.c:668) or through cuFFT/clFFT (versions 18/19, fft1.c:3519-3553): the float CPU
// versions stop at 65536 points (buf.c:285-290).
//
// Step A runs TA adjacent columns side by side with the column index fastest across lanes, so
// the strided gather from timf1 and the store to Y are both made of TA-element contiguous runs.
// Step B transposes its finished tile through shared memory so that bins k1..k1+TB-1 of one k2
// leave the SM as one contiguous run.
#pragma once
#include "fft1_small.cuh"

namespace lb {

// tile widths (log2): adjacent columns per CTA in step A, adjacent rows per CTA in step B.  Eight
// float2 = 64 bytes = two full sectors per run; small tiles keep several CTAs resident per SM so
// that the load, exchange and store phases of different CTAs overlap.
#ifndef LB_LARGE_LT
#define LB_LARGE_LT 3
#endif
// separate widths for the two steps (default: both LB_LARGE_LT)
#ifndef LB_LARGE_LTA
#define LB_LARGE_LTA LB_LARGE_LT
#endif
#ifndef LB_LARGE_LTB
#define LB_LARGE_LTB LB_LARGE_LT
#endif
// points per thread (log2) of the column / row transforms
#ifndef LB_LARGE_LE
#define LB_LARGE_LE 4
#endif
// resident CTAs per SM asked of the compiler for CTAs of more than 256 threads
#ifndef LB_LARGE_MINB
#define LB_LARGE_MINB 1
#endif

struct Fft1LargeK {
  Fft1K k;               // same parameter block as the single-CTA kernel
  float2* scratch;       // Y: [slot][channel][N]
  const float2* Wn1;     // exp(-2 pi i m / N1)
  const float2* Wn2;     // exp(-2 pi i m / N2)
  const float2* Wbig;    // exp(-2 pi i m / N)
  int b_first;           // first transform of this sub-batch (index into the call's batch)
  int b_count;           // transforms in this sub-batch
  int g_first;           // first averaging group covered by this sub-batch
  int g_count;
};

// ------------------------------------------------------------------------------ step A
template <int LOG2N1, int LOG2N2, int LOG2E, int LOG2TA, int FMT>
__global__ void __launch_bounds__(1 << (LOG2N1 - LOG2E + LOG2TA), (1 << (LOG2N1 - LOG2E + LOG2TA)) <= 256 ? 3 : LB_LARGE_MINB)
fft1_large_cols_kernel(const Fft1LargeK q)
{
  using P = Plan<LOG2N1, LOG2E>;
  constexpr int N1 = 1 << LOG2N1, N2 = 1 << LOG2N2, E = P::E, T = P::T, TA = 1 << LOG2TA;
  constexpr int FRAME = FmtInfo<FMT>::FRAME, NCH = FmtInfo<FMT>::NCH;
  constexpr int N = N1 * N2;
  constexpr int TILES = N2 / TA;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* xch = reinterpret_cast<float2*>(smem_raw);
  const Fft1K& p = q.k;
  const int col = threadIdx.x & (TA - 1);
  const int t = threadIdx.x >> LOG2TA;

  Twiddles<P> tw;
  load_twiddles<P>(tw, q.Wn1, t);

  const int nwork = q.b_count * NCH * TILES;
  for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int tile = w % TILES;
    const int c = (w / TILES) % NCH;
    const int slot = w / (TILES * NCH);
    const int b = q.b_first + slot;
    const int n2 = tile * TA + col;
    // inter-step twiddle W_N^(n2*(t+T*e)) = base * step^e, step given by exact binary powers
    const float2 base = q.Wbig[n2 * t];
    float2 sb[LOG2E];
#pragma unroll
    for (int j = 0; j < LOG2E; j++) sb[j] = q.Wbig[(n2 * T) << j];
    const float sgn = (n2 & 1) ? -1.0f : 1.0f;
    const float qs = p.direction < 0 ? -sgn : sgn;
    const uint32_t start = p.ref0 + (uint32_t)b * p.blockbytes - p.pre_bytes;
    float2 v[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int n = (t + T * e) * N2 + n2;
      const uint32_t off = (start + (uint32_t)n * FRAME) & p.ring_mask;
      float2 s;
      if ((FMT == FMT_I16_1CH || FMT == FMT_I32_1CH) && (p.skew_i | p.skew_q))
        s = load_iq_skew<FMT>(p.timf1, p.ring_mask, off, p.skew_i, p.skew_q);
      else
        s = load_iq<FMT>(p.timf1, off, c);
      if (FmtInfo<FMT>::REAL) {            // packed real pair, plain transform (fft1_re.c)
        const float2 wv = p.window ? reinterpret_cast<const float2*>(p.window)[n] : make_float2(1.0f, 1.0f);
        v[e] = make_float2(s.x * wv.x, s.y * wv.y);
      } else {
        const float wv = p.window ? p.window[n] : 1.0f;
        v[e] = make_float2(s.x * (wv * sgn), s.y * (wv * qs));
      }
    }
    __syncthreads();                       // previous work item's exchange reads are done
    fft_forward<P, TA>(v, xch + col, t, tw);
    apply_power_twiddles<E>(v, base, sb);
    float2* Y = q.scratch + ((size_t)slot * NCH + c) * N;
#pragma unroll
    for (int e = 0; e < E; e++) Y[(size_t)(t + T * e) * N2 + n2] = v[e];
  }
}

// ------------------------------------------------------------------------------ step B
template <int LOG2N1, int LOG2N2, int LOG2E, int LOG2TB, int NCH>
__global__ void __launch_bounds__(1 << (LOG2N2 - LOG2E + LOG2TB), (1 << (LOG2N2 - LOG2E + LOG2TB)) <= 256 ? 3 : LB_LARGE_MINB)
fft1_large_rows_kernel(const Fft1LargeK q)
{
  using P = Plan<LOG2N2, LOG2E>;
  constexpr int N1 = 1 << LOG2N1, N2 = 1 << LOG2N2, E = P::E, T = P::T, TB = 1 << LOG2TB;
  constexpr int N = N1 * N2, MM = 2 * NCH;
  constexpr int TILES = N1 / TB;
  constexpr int NTHREADS = TB * T;
  constexpr int XCH = N2 + N2 / 32 + 32;          // per-row exchange slice
  constexpr int TILE_PTS = TB * N2;
  constexpr int TP = TB + 1;                      // padded row of the transpose tile (conflict-free scatter)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* xch_all = reinterpret_cast<float2*>(smem_raw);      // TB rows; reused as the transpose tile
  const Fft1K& p = q.k;
  const int t = threadIdx.x & (T - 1);
  const int row = threadIdx.x / T;
  float2* xch = xch_all + row * XCH;

  Twiddles<P> tw;
  load_twiddles<P>(tw, q.Wn2, t);

  // work item = (transform, tile of TB rows).  |z|^2 is added straight into the fft1_sumsq row
  // (host-zeroed unless it continues a partial group): the transforms of a group are spread over
  // many CTAs, which keeps every SM busy even when a sub-batch holds only a few groups.
  const int group_size = p.power_rows ? 1 : p.avg1num;
  const int c0 = p.power_rows ? 0 : p.counter0;
  const int nwork = q.b_count * TILES;
  for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int tile = w % TILES;
    const int slot = w / TILES;
    const int b = q.b_first + slot;
    const int g = (b + c0) / group_size;
    const int k1 = tile * TB + row;
    float* outb = p.out + ((p.out_pa + (uint32_t)b * (uint32_t)(MM * N)) & p.out_mask);
    float* rowp = (p.sumsq && !p.power_rows) ? p.sumsq + ((p.sumsq_pa + (uint32_t)g * (uint32_t)N) & p.sumsq_mask) : nullptr;
    float* prow = p.power_rows ? p.power_rows + (size_t)b * N : nullptr;
#pragma unroll 1
    for (int c = 0; c < NCH; c++) {
      const float2* Y = q.scratch + ((size_t)slot * NCH + c) * N + (size_t)k1 * N2;
      float2 v[E];
#pragma unroll
      for (int e = 0; e < E; e++) v[e] = Y[t + T * e];
      __syncthreads();                   // the transpose tile of the previous pass is consumed
      fft_forward<P>(v, xch, t, tw);
      __syncthreads();
      // transpose: tile[k2][row]
#pragma unroll
      for (int e = 0; e < E; e++) xch_all[(t + T * e) * TP + row] = v[e];
      __syncthreads();
#pragma unroll 4
      for (int o = threadIdx.x; o < TILE_PTS; o += NTHREADS) {
        const int r = o & (TB - 1), k2 = o >> LOG2TB;
        const int k = tile * TB + r + N1 * k2;
        const float2 z = xch_all[k2 * TP + r];
        if (p.zbuf) {                    // real input: plain Z, finished by fft1_real_post_kernel
          p.zbuf[((size_t)(b - p.zb_first) * NCH + c) * N + k] = z;
          continue;
        }
        float2 ov = p.direction < 0 ? make_float2(z.y, z.x) : make_float2(z.x, -z.y);
        const bool inr = (k >= p.first_point) && (k <= p.last_point);
        if (p.fc_mode != 0) {
          if (inr) {
            float2 f;
            if (p.fc_mode == 2 || k < p.fc_edge || k >= N - p.fc_edge)
              f = *reinterpret_cast<const float2*>(p.filtercorr + (size_t)k * MM + 2 * c);
            else
              f = make_float2(p.fc_gain, 0.0f);
            const float re = ov.x * f.x - ov.y * f.y;
            const float im = ov.y * f.x + ov.x * f.y;
            ov = make_float2(re, im);
            const float pw = re * re + im * im;
            if (prow) {                  // the same thread owns bin k for both channels
              if (c == 0) prow[k] = pw;
              else prow[k] += pw;
            } else if (rowp) {
              atomicAdd(rowp + k, pw);
            }
          } else if (prow && c == 0) {
            prow[k] = 0.0f;
          }
        }
        __stcs(reinterpret_cast<float2*>(outb + (size_t)k * MM + 2 * c), ov);
      }
    }
  }
}

template <int LOG2N1, int LOG2TA>
constexpr size_t fft1_large_cols_smem() { return sizeof(float2) * ((size_t)1 << (LOG2N1 + LOG2TA)); }
template <int LOG2N2, int LOG2TB>
constexpr size_t fft1_large_rows_smem()
{
  const size_t xch = (size_t)(1 << LOG2TB) * ((1 << LOG2N2) + (1 << LOG2N2) / 32 + 32);
  const size_t tile = (size_t)((1 << LOG2TB) + 1) << LOG2N2;
  return sizeof(float2) * (xch > tile ? xch : tile);
}

}  // namespace lb
