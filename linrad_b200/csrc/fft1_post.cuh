// fft1_post.cuh -- the rarely used tail of fft1_b for IQ input, as a second kernel behind the
// transform kernels (which then run without their fft1_c epilogue):
//   * I/Q mirror-image correction with the calibration table fft1_foldcorr
//       (fft1.c:3607-3657 one channel, 3941-4026 two channels), per channel, for ib = m .. N/2-1, ic = N-ib:
//         z'[ib] = z[ib] - conj(z[ic]) * f[ic]        z'[ic] = z[ic] - conj(z[ib] * f[ib])
//       bins 0, N/2 and those outside [m, N-m] pass unchanged (m = max(1, fft1_first_sym_point));
//   * for fft1_direction < 0 the reference then reverses and swaps re/im IN THE SAME LOOP, so a
//     calibrated transform is run with direction +1 and flipped here:
//         out[ib] = swap(z'[ic]), out[ic] = swap(z'[ib]), bins 0 and N/2 swap re/im in place
//       (uncalibrated transforms get their flip for free inside the transform kernels);
//   * channel-2 phasing ch2 *= (c1 - i c2) on bins [fft1_first_sym_point, N - fft1_first_sym_point)
//       (fft1.c:4064-4080, pol_graph.c:165-173);
//   * fft1_c (fft1.c:4115-4200): filtercorr multiply, |z|^2 of all channels -> fft1_sumsq / power rows.
// One thread owns the bin pair (ib, N-ib) of all channels for a whole averaging group and walks
// the group's transforms in time order, so the power row is summed in the reference's order and
// written once.
#pragma once
#include "fft1_small.cuh"

namespace lb {

struct Fft1PostK {
  Fft1K k;                  // out ring, sumsq / power rows, filtercorr, first/last point, counters
  const float* foldcorr;    // mm*N floats or nullptr
  int flip;                 // 1: apply the direction < 0 reversal here
  int first_sym;            // fft1_first_sym_point (fft1.c:4647-4650)
  float c1, c2;             // pg_ch2_c1 / pg_ch2_c2
  int phasing;              // 0/1
  int N;
  // fft1_correlation_flag == 1: cross spectrum 2*z1*conj(z2) (fft1.c:4146-4152)
  float* corrsum;           // ring indexed like fft1_sumsq, two floats per bin; or nullptr
  float* corr_rows;         // per-transform rows of 2N floats (with power_rows); or nullptr
  float* xypower_rows;      // per-transform rows of N TWOCHAN_POWER {x2, y2, im_xy, re_xy} (fft1.c:4361-4364); or nullptr
};

template <int NCH>
__global__ void __launch_bounds__(256) fft1_iqpost_kernel(const Fft1PostK q)
{
  constexpr int MM = 2 * NCH;
  const Fft1K& p = q.k;
  const int N = q.N, H = N / 2;
  const int group_size = p.power_rows ? 1 : p.avg1num;
  const int c0 = p.power_rows ? 0 : p.counter0;
  const int ngroups = (c0 + p.nblocks + group_size - 1) / group_size;
  const int chunks = (H + 1 + 255) / 256;             // pair index j = 0..H: bins (j, N-j); j = 0 and j = H are single bins
  const int mcal = q.first_sym < 1 ? 1 : q.first_sym;
  for (int w = blockIdx.x; w < ngroups * chunks; w += gridDim.x) {
    const int g = w / chunks;
    const int j = (w - g * chunks) * 256 + threadIdx.x;
    if (j > H) continue;
    const int ib = j, ic = (j == 0 || j == H) ? j : N - j;
    const bool pair = ic != ib;
    int b0 = g * group_size - c0;
    int b1 = b0 + group_size;
    if (b0 < 0) b0 = 0;
    if (b1 > p.nblocks) b1 = p.nblocks;
    float accb = 0.f, accc = 0.f;
    float2 corb = make_float2(0.f, 0.f), corc = make_float2(0.f, 0.f);
    for (int b = b0; b < b1; b++) {
      float* out = p.out + ((p.out_pa + (uint32_t)b * (uint32_t)(MM * N)) & p.out_mask);
      float2 zb[NCH], zc[NCH];
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        zb[c] = *reinterpret_cast<const float2*>(out + (size_t)ib * MM + 2 * c);
        zc[c] = pair ? *reinterpret_cast<const float2*>(out + (size_t)ic * MM + 2 * c) : zb[c];
      }
      // ---- mirror-image correction
      if (q.foldcorr && pair && ib >= mcal) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          const float2 fb = *reinterpret_cast<const float2*>(q.foldcorr + (size_t)ib * MM + 2 * c);
          const float2 fcx = *reinterpret_cast<const float2*>(q.foldcorr + (size_t)ic * MM + 2 * c);
          const float2 b_ = zb[c], c_ = zc[c];
          const float t1 = b_.x * fb.x - b_.y * fb.y;
          const float t2 = b_.x * fb.y + b_.y * fb.x;
          zb[c] = make_float2(b_.x - (c_.x * fcx.x + c_.y * fcx.y), b_.y - (c_.x * fcx.y - c_.y * fcx.x));
          zc[c] = make_float2(c_.x - t1, c_.y + t2);
        }
      }
      // ---- direction < 0 (only when the transform kernel did not do it)
      if (q.flip) {
        if (pair) {
          if (ib >= mcal) {
#pragma unroll
            for (int c = 0; c < NCH; c++) {
              const float2 nb = make_float2(zc[c].y, zc[c].x), nc = make_float2(zb[c].y, zb[c].x);
              zb[c] = nb;
              zc[c] = nc;
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < NCH; c++) zb[c] = make_float2(zb[c].y, zb[c].x);    // bins 0 and N/2 (fft1.c:3649-3654)
        }
      }
      // ---- channel-2 phasing
      if (NCH == 2 && q.phasing) {
        if (ib >= q.first_sym && ib < N - q.first_sym) {
          const float2 z = zb[NCH - 1];
          zb[NCH - 1] = make_float2(z.x * q.c1 + z.y * q.c2, z.y * q.c1 - z.x * q.c2);
        }
        if (pair && ic >= q.first_sym && ic < N - q.first_sym) {
          const float2 z = zc[NCH - 1];
          zc[NCH - 1] = make_float2(z.x * q.c1 + z.y * q.c2, z.y * q.c1 - z.x * q.c2);
        }
      }
      // ---- fft1_c
      float pwb = 0.f, pwc = 0.f;
      float2 xb = make_float2(0.f, 0.f), xc = make_float2(0.f, 0.f);
      if (p.fc_mode != 0) {
        if (ib >= p.first_point && ib <= p.last_point) {
#pragma unroll
          for (int c = 0; c < NCH; c++) {
            const float2 f = *reinterpret_cast<const float2*>(p.filtercorr + (size_t)ib * MM + 2 * c);
            const float2 z = zb[c];
            zb[c] = make_float2(z.x * f.x - z.y * f.y, z.y * f.x + z.x * f.y);
            pwb += zb[c].x * zb[c].x + zb[c].y * zb[c].y;
          }
          if (NCH == 2) xb = make_float2(2.0f * (zb[0].x * zb[NCH - 1].x + zb[0].y * zb[NCH - 1].y),
                                         2.0f * (zb[0].y * zb[NCH - 1].x - zb[0].x * zb[NCH - 1].y));
        }
        if (pair && ic >= p.first_point && ic <= p.last_point) {
#pragma unroll
          for (int c = 0; c < NCH; c++) {
            const float2 f = *reinterpret_cast<const float2*>(p.filtercorr + (size_t)ic * MM + 2 * c);
            const float2 z = zc[c];
            zc[c] = make_float2(z.x * f.x - z.y * f.y, z.y * f.x + z.x * f.y);
            pwc += zc[c].x * zc[c].x + zc[c].y * zc[c].y;
          }
          if (NCH == 2) xc = make_float2(2.0f * (zc[0].x * zc[NCH - 1].x + zc[0].y * zc[NCH - 1].y),
                                         2.0f * (zc[0].y * zc[NCH - 1].x - zc[0].x * zc[NCH - 1].y));
        }
      }
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        *reinterpret_cast<float2*>(out + (size_t)ib * MM + 2 * c) = zb[c];
        if (pair) *reinterpret_cast<float2*>(out + (size_t)ic * MM + 2 * c) = zc[c];
      }
      if (p.power_rows && p.fc_mode != 0) {
        p.power_rows[(size_t)b * N + ib] = pwb;
        if (pair) p.power_rows[(size_t)b * N + ic] = pwc;
      }
      if (NCH == 2 && q.xypower_rows && p.fc_mode != 0) {
        // xb = 2 (re_xy, im_xy); bins outside [first_point, last_point] give zeros
        const bool inb = ib >= p.first_point && ib <= p.last_point, inc = ic >= p.first_point && ic <= p.last_point;
        const float4 vb = inb ? make_float4(zb[0].x * zb[0].x + zb[0].y * zb[0].y, zb[1].x * zb[1].x + zb[1].y * zb[1].y, 0.5f * xb.y, 0.5f * xb.x)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(q.xypower_rows + ((size_t)b * N + ib) * 4) = vb;
        if (pair) {
          const float4 vc = inc ? make_float4(zc[0].x * zc[0].x + zc[0].y * zc[0].y, zc[1].x * zc[1].x + zc[1].y * zc[1].y, 0.5f * xc.y, 0.5f * xc.x)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(q.xypower_rows + ((size_t)b * N + ic) * 4) = vc;
        }
      }
      if (NCH == 2 && q.corr_rows && p.fc_mode != 0) {
        *reinterpret_cast<float2*>(q.corr_rows + ((size_t)b * N + ib) * 2) = xb;
        if (pair) *reinterpret_cast<float2*>(q.corr_rows + ((size_t)b * N + ic) * 2) = xc;
      }
      accb = (b == b0) ? pwb : accb + pwb;
      accc = (b == b0) ? pwc : accc + pwc;
      corb = (b == b0) ? xb : make_float2(corb.x + xb.x, corb.y + xb.y);
      corc = (b == b0) ? xc : make_float2(corc.x + xc.x, corc.y + xc.y);
    }
    if (p.sumsq && !p.power_rows && p.fc_mode != 0 && b1 > b0) {
      float* row = p.sumsq + ((p.sumsq_pa + (uint32_t)g * (uint32_t)N) & p.sumsq_mask);
      const bool continuing = (g == 0 && p.counter0 > 0);
      if (ib >= p.first_point && ib <= p.last_point) row[ib] = continuing ? row[ib] + accb : accb;
      if (pair && ic >= p.first_point && ic <= p.last_point) row[ic] = continuing ? row[ic] + accc : accc;
      if (NCH == 2 && q.corrsum) {
        float2* crow = reinterpret_cast<float2*>(q.corrsum) + ((p.sumsq_pa + (uint32_t)g * (uint32_t)N) & p.sumsq_mask);
        if (ib >= p.first_point && ib <= p.last_point) {
          const float2 o = crow[ib];
          crow[ib] = continuing ? make_float2(o.x + corb.x, o.y + corb.y) : corb;
        }
        if (pair && ic >= p.first_point && ic <= p.last_point) {
          const float2 o = crow[ic];
          crow[ic] = continuing ? make_float2(o.x + corc.x, o.y + corc.y) : corc;
        }
      }
    }
  }
}

}  // namespace lb
