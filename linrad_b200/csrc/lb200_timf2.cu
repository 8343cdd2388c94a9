// lb200_timf2.cu -- C-ABI entry points of make_timf2 (timf2.cuh)
#include <cstdio>
#include <cstring>
#include <vector>
#include "plan.h"
#include "timf2.cuh"

using namespace lb;

static bool t2_pow2(size_t v) { return v && !(v & (v - 1)); }

template <int LOG2N, int NCH>
static cudaError_t launch_back(const Timf2K& k, cudaStream_t s)
{
  constexpr int N = 1 << LOG2N;
  constexpr size_t smem = LOG2N >= 10 ? sizeof(float2) * Plan32<LOG2N >= 10 ? LOG2N : 10>::XCH + sizeof(float4) * Plan32<LOG2N >= 10 ? LOG2N : 10>::TAB1
                                      : sizeof(float2) * (N + N / 32 + 32);
  constexpr int threads = LOG2N >= 10 ? (1 << (LOG2N - 5)) : (1 << (LOG2N - 3));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(timf2_back_kernel<LOG2N, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  timf2_back_kernel<LOG2N, NCH><<<k.nblocks * 2 * NCH, threads, smem, s>>>(k);
  return cudaGetLastError();
}

template <int NCH>
static cudaError_t launch_back_n(int log2n, const Timf2K& k, cudaStream_t s)
{
  switch (log2n) {
    case 7: return launch_back<7, NCH>(k, s);
    case 8: return launch_back<8, NCH>(k, s);
    case 9: return launch_back<9, NCH>(k, s);
    case 10: return launch_back<10, NCH>(k, s);
    case 11: return launch_back<11, NCH>(k, s);
    case 12: return launch_back<12, NCH>(k, s);
    case 13: return launch_back<13, NCH>(k, s);
    case 14: return launch_back<14, NCH>(k, s);
  }
  return cudaErrorNotSupported;
}

static int timf2_mode(const lb200_plan* plan)
{
  const int il = plan->cfg.fft1_interleave_points;
  return il == 0 ? 0 : (il == plan->N / 2 ? 1 : 2);
}

static int check_timf2(lb200_plan* plan, const lb200_timf2_args* a)
{
  if (!plan || !a || !a->fft1_float.base || !a->timf2_float.base || !a->timf2_pwr_float || !a->liminfo) return LB200_ERR_BAD_ARG;
  if (!t2_pow2(a->fft1_float.size) || !t2_pow2(a->timf2_float.size)) return LB200_ERR_BAD_ARG;
  if (a->first_bckfft_att_n < 0 || a->first_bckfft_att_n > 30) return LB200_ERR_BAD_ARG;
  if (plan->cfg.fft1_n < 7 || plan->cfg.fft1_n > 14) return LB200_ERR_UNSUPPORTED;       // back transform: single-CTA sizes
  if (timf2_mode(plan) == 2 && !plan->d_invwin) return LB200_ERR_BAD_CONFIG;
  const size_t sf = 4 * (size_t)plan->nch;
  const size_t newp = (size_t)plan->N - plan->cfg.fft1_interleave_points;
  // the whole call (plus the parked half) must fit the ring without lapping itself
  if (sf * (newp * a->nblocks + plan->N / 2) > a->timf2_float.size) return LB200_ERR_BAD_ARG;
  if ((a->timf2_pa % sf) || (a->fft1_px % plan->fft1_block)) return LB200_ERR_BAD_ARG;
  return 0;
}

// fft1_lowlevel_points of timf2.c:38-51 (the same for every transform of a call: liminfo is fixed)
static int lowlevel_points(const lb200_plan* plan, const float* liminfo)
{
  int n = 0;
  for (int i = plan->cfg.fft1_first_point; i <= plan->cfg.fft1_last_point; i++)
    if (liminfo[i] == 0.0f) n++;
  return n;
}

static int run_timf2(lb200_plan* plan, const lb200_timf2_args* a, const float* d_fft1, const float* d_liminfo, float* d_timf2, float* d_pwr)
{
  const int N = plan->N, S = 2 * plan->nch;
  const size_t need = (size_t)a->nblocks * S * N;
  if (plan->timf2_tmp_elems < need) {
    if (plan->d_timf2_tmp) cudaFree(plan->d_timf2_tmp);
    plan->d_timf2_tmp = nullptr;
    plan->timf2_tmp_elems = 0;
    LB_CUDA(cudaMalloc((void**)&plan->d_timf2_tmp, need * sizeof(float2)));
    plan->timf2_tmp_elems = need;
  }
  Timf2K k;
  memset(&k, 0, sizeof(k));
  k.fft1 = d_fft1;
  k.fft1_mask = (uint32_t)(a->fft1_float.size - 1);
  k.fft1_px = a->fft1_px;
  k.nblocks = a->nblocks;
  k.liminfo = d_liminfo;
  k.first_point = plan->cfg.fft1_first_point;
  k.last_point = plan->cfg.fft1_last_point;
  k.tmp = plan->d_timf2_tmp;
  k.Wn = plan->d_Wn;
  k.tab1 = plan->d_tab1_any;
  k.timf2 = d_timf2;
  k.timf2_mask = (uint32_t)(a->timf2_float.size - 1);
  k.timf2_pa = a->timf2_pa;
  k.pwr = d_pwr;
  k.ampfac = 1.0f / (float)(1 << a->first_bckfft_att_n);
  k.invwin = plan->d_invwin;
  k.interleave = plan->cfg.fft1_interleave_points;
  k.mode = timf2_mode(plan);
  if (plan->nch == 1) LB_CUDA(launch_back_n<1>(plan->cfg.fft1_n, k, plan->stream));
  else LB_CUDA(launch_back_n<2>(plan->cfg.fft1_n, k, plan->stream));
  const long total = (long)a->nblocks * (N - k.interleave) + (k.mode == 1 ? N / 2 : 0);
  long grid = (total + 255) / 256;
  if (grid > (long)plan->sm_count * 16) grid = (long)plan->sm_count * 16;
  if (plan->nch == 1) timf2_finish_kernel<1><<<(int)grid, 256, 0, plan->stream>>>(k, plan->cfg.fft1_n);
  else timf2_finish_kernel<2><<<(int)grid, 256, 0, plan->stream>>>(k, plan->cfg.fft1_n);
  LB_CUDA(cudaGetLastError());
  plan->launches += 2;
  return LB200_OK;
}

extern "C" int lb200_make_timf2_dev(lb200_plan* plan, const lb200_timf2_args* a)
{
  int rc = check_timf2(plan, a);
  if (rc) return rc;
  if (a->nblocks <= 0) return LB200_OK;
  cudaSetDevice(plan->device);
  if (a->fft1_lowlevel_points) return LB200_ERR_BAD_ARG;      // liminfo is on the device here: count on the host side of the caller
  return run_timf2(plan, a, (const float*)a->fft1_float.base, a->liminfo, (float*)a->timf2_float.base, a->timf2_pwr_float);
}

static int t2_mirror(lb200_plan* plan, HostMirror& m, size_t bytes)
{
  if (m.d && m.bytes >= bytes) return 0;
  if (m.d) cudaFree(m.d);
  m.d = nullptr;
  LB_CUDA(cudaMalloc(&m.d, bytes));
  LB_CUDA(cudaMemsetAsync(m.d, 0, bytes, plan->stream));
  m.bytes = bytes;
  return 0;
}

// copy [off, off+len) floats of a power-of-two ring, wrapping
static int t2_ring(lb200_plan* plan, float* dev, float* host, size_t size, size_t off, size_t len, bool h2d)
{
  off &= size - 1;
  while (len > 0) {
    size_t n = size - off;
    if (n > len) n = len;
    if (h2d) {
      LB_CUDA(cudaMemcpyAsync(dev + off, host + off, n * 4, cudaMemcpyHostToDevice, plan->stream));
      plan->h2d += n * 4;
    } else {
      LB_CUDA(cudaMemcpyAsync(host + off, dev + off, n * 4, cudaMemcpyDeviceToHost, plan->stream));
      plan->d2h += n * 4;
    }
    len -= n;
    off = (off + n) & (size - 1);
  }
  return 0;
}

extern "C" int lb200_make_timf2(lb200_plan* plan, const lb200_timf2_args* a)
{
  int rc = check_timf2(plan, a);
  if (rc) return rc;
  if (a->nblocks <= 0) return LB200_OK;
  cudaSetDevice(plan->device);
  const int N = plan->N;
  const size_t sf = 4 * (size_t)plan->nch;
  const size_t newp = (size_t)N - plan->cfg.fft1_interleave_points;
  const int mode = timf2_mode(plan);
  if (a->fft1_lowlevel_points) *a->fft1_lowlevel_points = lowlevel_points(plan, a->liminfo);
  if ((rc = t2_mirror(plan, plan->m_t2_fft1, a->fft1_float.size * 4))) return rc;
  if ((rc = t2_mirror(plan, plan->m_t2_ring, a->timf2_float.size * 4))) return rc;
  if ((rc = t2_mirror(plan, plan->m_t2_pwr, a->timf2_float.size / sf * 4))) return rc;
  if ((rc = t2_mirror(plan, plan->m_t2_lim, (size_t)N * 4))) return rc;
  LB_CUDA(cudaMemcpyAsync(plan->m_t2_lim.d, a->liminfo, (size_t)N * 4, cudaMemcpyHostToDevice, plan->stream));
  plan->h2d += (size_t)N * 4;
  if ((rc = t2_ring(plan, (float*)plan->m_t2_fft1.d, (float*)a->fft1_float.base, a->fft1_float.size, a->fft1_px, (size_t)plan->fft1_block * a->nblocks, true))) return rc;
  // the half the previous call parked in the ring (sin^2 window only)
  if (mode == 1)
    if ((rc = t2_ring(plan, (float*)plan->m_t2_ring.d, (float*)a->timf2_float.base, a->timf2_float.size, a->timf2_pa, sf * (N / 2), true))) return rc;
  lb200_timf2_args d = *a;
  d.fft1_lowlevel_points = nullptr;
  if ((rc = run_timf2(plan, &d, (const float*)plan->m_t2_fft1.d, (const float*)plan->m_t2_lim.d, (float*)plan->m_t2_ring.d, (float*)plan->m_t2_pwr.d))) return rc;
  const size_t out_samples = newp * a->nblocks + (mode == 1 ? N / 2 : 0);
  if ((rc = t2_ring(plan, (float*)plan->m_t2_ring.d, (float*)a->timf2_float.base, a->timf2_float.size, a->timf2_pa, sf * out_samples, false))) return rc;
  if ((rc = t2_ring(plan, (float*)plan->m_t2_pwr.d, a->timf2_pwr_float, a->timf2_float.size / sf, a->timf2_pa / sf, newp * a->nblocks, false))) return rc;
  LB_CUDA(cudaStreamSynchronize(plan->stream));
  return LB200_OK;
}
