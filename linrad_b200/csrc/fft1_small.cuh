// fft1_small.cuh -- fused fft1 kernel for transforms that fit one CTA (N <= 2^14 per channel).
//
// One launch replaces, for a batch of consecutive time blocks,
//   fft1win_dit_one / fft1win_dif_chan   (fft1.c:684-1028, 2041-2247: ring gather, int->float, window)
//   bulk_of_dit / bulk_of_dif + permute  (fft0.c:1590-1769, 161-195; fft1.c:637-682)
//   the direction flip of fft1_b         (fft1.c:3660-3680, 4029-4060)
//   fft1_c                               (fft1.c:4115-4200: filtercorr multiply, |z|^2, fft1_sumsq)
// Output convention (probed from the compiled reference):
//   fft1_float[k] = conj( sum_n w[n] x[n] exp(-2 pi i n ((k+N/2) mod N)/N) ) * filtercorr[k]
// The (k+N/2) rotation is obtained for free by negating odd-numbered input samples.
//
// Work decomposition: one CTA owns one *averaging group* (the wg.fft_avg1num consecutive
// transforms that are summed into one fft1_sumsq row, fft1.c:4507-4520) and walks through its
// transforms and channels in time order, so the row is accumulated on chip in the reference's
// own order and written exactly once.
#pragma once
#include "fft_core.cuh"

namespace lb {

// Input formats.  IQ: one timf1 frame per complex point.  Real input (fft1_re.c): two
// consecutive frames are packed as one complex point z[m] = x[2m] + i x[2m+1]; FRAME is then
// the byte size of the PAIR of frames.
enum InFmt { FMT_I16_1CH = 0, FMT_I16_2CH = 1, FMT_I32_1CH = 2, FMT_I32_2CH = 3,
             FMT_R16_1CH = 4, FMT_R16_2CH = 5, FMT_R32_1CH = 6, FMT_R32_2CH = 7,
             // float IQ frames [re, im] / [re1, im1, re2, im2]: the timf3 baseband ring as input of the
             // third FFT (make_fft3_all, fft3.c:215-470) -- same frame geometry as the int32 formats
             FMT_F32_1CH = 8, FMT_F32_2CH = 9 };
template <int FMT> struct FmtInfo;
template <> struct FmtInfo<FMT_I16_1CH> { static constexpr int FRAME = 4, NCH = 1; static constexpr bool REAL = false; };
template <> struct FmtInfo<FMT_I16_2CH> { static constexpr int FRAME = 8, NCH = 2; static constexpr bool REAL = false; };
template <> struct FmtInfo<FMT_I32_1CH> { static constexpr int FRAME = 8, NCH = 1; static constexpr bool REAL = false; };
template <> struct FmtInfo<FMT_I32_2CH> { static constexpr int FRAME = 16, NCH = 2; static constexpr bool REAL = false; };
template <> struct FmtInfo<FMT_R16_1CH> { static constexpr int FRAME = 4, NCH = 1; static constexpr bool REAL = true; };
template <> struct FmtInfo<FMT_R16_2CH> { static constexpr int FRAME = 8, NCH = 2; static constexpr bool REAL = true; };
template <> struct FmtInfo<FMT_R32_1CH> { static constexpr int FRAME = 8, NCH = 1; static constexpr bool REAL = true; };
template <> struct FmtInfo<FMT_R32_2CH> { static constexpr int FRAME = 16, NCH = 2; static constexpr bool REAL = true; };
template <> struct FmtInfo<FMT_F32_1CH> { static constexpr int FRAME = 8, NCH = 1; static constexpr bool REAL = false; };
template <> struct FmtInfo<FMT_F32_2CH> { static constexpr int FRAME = 16, NCH = 2; static constexpr bool REAL = false; };

struct Fft1K {
  const uint8_t* timf1;     // device ring
  uint32_t ring_mask;       // bytes
  uint32_t ref0;            // byte offset of first new sample of transform 0
  uint32_t blockbytes;      // timf1_blockbytes
  uint32_t pre_bytes;       // fft1_interleave_points * frame bytes
  int nblocks;
  const float* window;      // natural order, nullptr = rectangular
  const float2* Wn;         // exp(-2 pi i m / N)
  const float* filtercorr;  // mm*N floats
  int fc_mode;              // 0 none (raw fft1_b), 1 uniform real gain except edges, 2 full table
  float fc_gain;            // the uniform gain for fc_mode 1
  int fc_edge;              // bins [0,fc_edge) and [N-fc_edge,N) always read the table
  float* out;               // fft1_float ring
  uint32_t out_mask;        // floats
  uint32_t out_pa;          // floats
  float* sumsq;             // fft1_sumsq ring or nullptr
  uint32_t sumsq_mask;
  uint32_t sumsq_pa;
  int counter0;             // fft1_sumsq_counter on entry
  int avg1num;
  float* power_rows;        // per-transform |z|^2 rows or nullptr
  int first_point, last_point;
  int direction;
  // fft1_fused_kernel only
  const float* wtab;        // window * (-1)^n [* uniform gain for FC_FOLDED], natural order
  const float2* edge;       // FC_FOLDED: filtercorr/gain for bins 0..15 and N-16..N-1 (32 entries)
  float2* scratch2;         // 2-channel formats: gridDim.x rows of N float2 (channel 0 parked until channel 1 is done)
  const float4* tab1;       // pass-1 twiddles [pair][k] = (w^(2q), w^(2q+1)), w = exp(-2 pi i k/(32 R0))
  // real input (fft1_re.c): the transform kernels leave the plain packed spectrum Z in zbuf
  // ([transform - zb_first][channel][N] float2, L2-resident) and fft1_real_post_kernel finishes
  float2* zbuf;
  int zb_first;
  const float2* Wre;        // exp(-i pi k / N), k = 0..N
  uint32_t stagger_ns;      // fft1_fused_kernel: CTAs start spread over this many ns (0 = together)
  // ui.sample_shift (one-channel IQ only, fft1.c:770-790): byte offsets of the frame the I word
  // and the frame the Q word of sample n are taken from, relative to frame n (both 0 = off)
  int skew_i, skew_q;
  int stage_raw;            // fft1_fused_kernel, int16 one-channel IQ: raw spans by TMA bulk load (16-byte aligned spans only)
  // several rings in one launch (fft1_small_kernel, raw output only): ring r = blockIdx.y reads at timf1 + r*ring_stride
  // (bytes) and writes at out_pa + r*out_pa_stride (floats, same output ring) -- the selections of the third FFT
  int nrings;
  size_t ring_stride;
  uint32_t out_pa_stride;
  // fft1_fused_kernel, two-channel formats: the two channel CTAs of a group are launched as one 2-CTA cluster and
  // hand whole [re1,im1,re2,im2] slots to the TMA unit (see the kernel); 0 = independent CTAs with 8-byte stores
  int cluster2;
};

// I from the frame at off + skew_i, Q from the frame at off + skew_q (one-channel IQ formats)
template <int FMT>
LB_D float2 load_iq_skew(const uint8_t* ring, uint32_t mask, uint32_t off, int skew_i, int skew_q)
{
  const uint32_t oi = (off + (uint32_t)skew_i) & mask, oq = (off + (uint32_t)skew_q) & mask;
  if (FMT == FMT_I16_1CH)
    return make_float2((float)*reinterpret_cast<const short*>(ring + oi), (float)*reinterpret_cast<const short*>(ring + oq + 2));
  if (FMT == FMT_F32_1CH)
    return make_float2(*reinterpret_cast<const float*>(ring + oi), -*reinterpret_cast<const float*>(ring + oq + 4));
  return make_float2((float)*reinterpret_cast<const int*>(ring + oi), (float)*reinterpret_cast<const int*>(ring + oq + 4));
}

template <int FMT>
LB_D float2 load_iq(const uint8_t* ring, uint32_t off, int c)
{
  if (FMT == FMT_I16_1CH) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(ring + off);
    return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
  } else if (FMT == FMT_I16_2CH) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(ring + off + 4 * c);
    return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
  } else if (FMT == FMT_I32_1CH) {
    const int2 w = *reinterpret_cast<const int2*>(ring + off);
    return make_float2((float)w.x, (float)w.y);
  } else if (FMT == FMT_I32_2CH) {
    const int2 w = *reinterpret_cast<const int2*>(ring + off + 8 * c);
    return make_float2((float)w.x, (float)w.y);
  } else if (FMT == FMT_F32_1CH) {
    // make_fft3_all transforms x, fft1_b transforms conj(x) (probed on the compiled reference: fft3 =
    // sum x w exp(+2 pi i n (k - N/2)/N), fft1 the same sum over conj(x)): conjugate on the way in
    const float2 z = *reinterpret_cast<const float2*>(ring + off);
    return make_float2(z.x, -z.y);
  } else if (FMT == FMT_F32_2CH) {
    const float2 z = *reinterpret_cast<const float2*>(ring + off + 8 * c);
    return make_float2(z.x, -z.y);
  } else if (FMT == FMT_R16_1CH) {                     // [x(2m), x(2m+1)]
    const uint32_t w = *reinterpret_cast<const uint32_t*>(ring + off);
    return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
  } else if (FMT == FMT_R16_2CH) {                     // [x1(2m), x2(2m), x1(2m+1), x2(2m+1)]
    const uint2 w = *reinterpret_cast<const uint2*>(ring + off);
    const uint32_t a = c ? (w.x >> 16) : (w.x & 0xffffu), b = c ? (w.y >> 16) : (w.y & 0xffffu);
    return make_float2((float)(short)a, (float)(short)b);
  } else if (FMT == FMT_R32_1CH) {
    const int2 w = *reinterpret_cast<const int2*>(ring + off);
    return make_float2((float)w.x, (float)w.y);
  } else {                                             // FMT_R32_2CH
    const int4 w = *reinterpret_cast<const int4*>(ring + off);
    return c ? make_float2((float)w.y, (float)w.w) : make_float2((float)w.x, (float)w.z);
  }
}

template <int LOG2N, int LOG2E, int FMT>
__global__ void __launch_bounds__(1 << (LOG2N - LOG2E))
fft1_small_kernel(const Fft1K p)
{
  using P = Plan<LOG2N, LOG2E>;
  constexpr int N = P::N, E = P::E, T = P::T;
  constexpr int FRAME = FmtInfo<FMT>::FRAME, NCH = FmtInfo<FMT>::NCH, MM = 2 * NCH;
  constexpr bool REAL = FmtInfo<FMT>::REAL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* xch = reinterpret_cast<float2*>(smem_raw);
  float* acc = reinterpret_cast<float*>(smem_raw + sizeof(float2) * (N + N / 32 + 32));
  const int t = threadIdx.x;

  Twiddles<P> tw;
  load_twiddles<P>(tw, p.Wn, t);
  const float sgn = (t & 1) ? -1.0f : 1.0f;          // (-1)^n: rotates the spectrum by N/2
  const float qs = p.direction < 0 ? -sgn : sgn;     // conj(input) reverses the frequency axis

  const int group_size = p.power_rows ? 1 : p.avg1num;
  const int c0 = p.power_rows ? 0 : p.counter0;
  const int ngroups = (c0 + p.nblocks + group_size - 1) / group_size;
  const uint8_t* const timf1_r = p.timf1 + (size_t)blockIdx.y * p.ring_stride;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    int b0 = g * group_size - c0;
    int b1 = b0 + group_size;
    if (b0 < 0) b0 = 0;
    if (b1 > p.nblocks) b1 = p.nblocks;
    for (int b = b0; b < b1; b++) {
      const uint32_t start = p.ref0 + (uint32_t)b * p.blockbytes - p.pre_bytes;
      float* outb = p.out + ((p.out_pa + (uint32_t)blockIdx.y * p.out_pa_stride + (uint32_t)b * (uint32_t)(MM * N)) & p.out_mask);
#pragma unroll 1
      for (int c = 0; c < NCH; c++) {
        float2 v[E];
#pragma unroll
        for (int e = 0; e < E; e++) {
          const int idx = t + T * e;
          const uint32_t off = (start + (uint32_t)idx * FRAME) & p.ring_mask;
          float2 s;
          if ((FMT == FMT_I16_1CH || FMT == FMT_I32_1CH) && (p.skew_i | p.skew_q))
            s = load_iq_skew<FMT>(timf1_r, p.ring_mask, off, p.skew_i, p.skew_q);
          else
            s = load_iq<FMT>(timf1_r, off, c);
          if (REAL) {
            const float2 w = p.window ? reinterpret_cast<const float2*>(p.window)[idx] : make_float2(1.0f, 1.0f);
            v[e] = make_float2(s.x * w.x, s.y * w.y);
          } else {
            const float w = p.window ? p.window[idx] : 1.0f;
            v[e] = make_float2(s.x * (w * sgn), s.y * (w * qs));
          }
        }
        fft_forward<P>(v, xch, t, tw);
        if (REAL) {
          float2* z = p.zbuf + ((size_t)(b - p.zb_first) * NCH + c) * N;
#pragma unroll
          for (int e = 0; e < E; e++) z[t + T * e] = v[e];
          continue;
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
          const int k = t + T * e;
          // direction>0: conj(Y).  direction<0: j*conj(Y') with Y' from the conjugated input.
          float2 o = p.direction < 0 ? make_float2(v[e].y, v[e].x) : make_float2(v[e].x, -v[e].y);
          const bool inr = (k >= p.first_point) && (k <= p.last_point);
          if (p.fc_mode != 0 && inr) {
            float2 f;
            if (p.fc_mode == 2 || k < p.fc_edge || k >= N - p.fc_edge)
              f = *reinterpret_cast<const float2*>(p.filtercorr + (size_t)k * MM + 2 * c);
            else
              f = make_float2(p.fc_gain, 0.0f);
            const float re = o.x * f.x - o.y * f.y;       // fft1.c:4121-4125
            const float im = o.y * f.x + o.x * f.y;
            o = make_float2(re, im);
            const float pw = re * re + im * im;
            if (b == b0 && c == 0) acc[k] = pw;
            else acc[k] += pw;
          }
          *reinterpret_cast<float2*>(outb + (size_t)k * MM + 2 * c) = o;
        }
      }
      if (!REAL && p.power_rows && p.fc_mode != 0) {
#pragma unroll
        for (int e = 0; e < E; e++) {
          const int k = t + T * e;
          const bool inr = (k >= p.first_point) && (k <= p.last_point);
          p.power_rows[(size_t)b * N + k] = inr ? acc[k] : 0.0f;
        }
      }
    }
    if (!REAL && p.sumsq && !p.power_rows && p.fc_mode != 0 && b1 > b0) {
      float* row = p.sumsq + ((p.sumsq_pa + (uint32_t)g * (uint32_t)N) & p.sumsq_mask);
      const bool continuing = (g == 0 && p.counter0 > 0);
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int k = t + T * e;
        if (k >= p.first_point && k <= p.last_point) {
          float val = acc[k];
          if (continuing) val = row[k] + val;
          row[k] = val;
        }
      }
    }
  }
}

template <int LOG2N, int LOG2E>
constexpr size_t fft1_small_smem()
{
  return sizeof(float2) * ((1 << LOG2N) + (1 << LOG2N) / 32 + 32) + sizeof(float) * (1 << LOG2N);
}

}  // namespace lb
