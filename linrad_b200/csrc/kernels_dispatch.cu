// kernels_dispatch.cu -- format dispatch for the per-format translation units
#include <cuda_runtime.h>
#include "fft1_small.cuh"
#include "plan.h"
using namespace lb;
typedef cudaError_t (*fft1_small_launch_t)(const Fft1K&, int grid, cudaStream_t);
fft1_small_launch_t lb_get_fft1_small_fmt0(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt1(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt2(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt3(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt4(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt5(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt6(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt7(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt8(int, int, int*, size_t*);
fft1_small_launch_t lb_get_fft1_small_fmt9(int, int, int*, size_t*);

fft1_small_launch_t lb_get_fft1_small(int log2n, int fmt, int variant, int* threads, size_t* smem)
{
  switch (fmt) {
    case 0: return lb_get_fft1_small_fmt0(log2n, variant, threads, smem);
    case 1: return lb_get_fft1_small_fmt1(log2n, variant, threads, smem);
    case 2: return lb_get_fft1_small_fmt2(log2n, variant, threads, smem);
    case 3: return lb_get_fft1_small_fmt3(log2n, variant, threads, smem);
    case 4: return lb_get_fft1_small_fmt4(log2n, variant, threads, smem);
    case 5: return lb_get_fft1_small_fmt5(log2n, variant, threads, smem);
    case 6: return lb_get_fft1_small_fmt6(log2n, variant, threads, smem);
    case 7: return lb_get_fft1_small_fmt7(log2n, variant, threads, smem);
    case 8: return lb_get_fft1_small_fmt8(log2n, variant, threads, smem);
    case 9: return lb_get_fft1_small_fmt9(log2n, variant, threads, smem);
  }
  return nullptr;
}


#define LB_DECL_FUSED(F, C) fft1_small_launch_t lb_get_fft1_fused_fmt##F##_fc##C(int, int*, size_t*);
LB_DECL_FUSED(0, 0) LB_DECL_FUSED(0, 1) LB_DECL_FUSED(0, 2) LB_DECL_FUSED(1, 0) LB_DECL_FUSED(1, 1) LB_DECL_FUSED(1, 2)
LB_DECL_FUSED(2, 0) LB_DECL_FUSED(2, 1) LB_DECL_FUSED(2, 2) LB_DECL_FUSED(3, 0) LB_DECL_FUSED(3, 1) LB_DECL_FUSED(3, 2)

fft1_small_launch_t lb_get_fft1_fused(int log2n, int fmt, int fc, int* threads, size_t* smem)
{
  typedef fft1_small_launch_t (*getter_t)(int, int*, size_t*);
  static const getter_t g[4][3] = {
      {lb_get_fft1_fused_fmt0_fc0, lb_get_fft1_fused_fmt0_fc1, lb_get_fft1_fused_fmt0_fc2},
      {lb_get_fft1_fused_fmt1_fc0, lb_get_fft1_fused_fmt1_fc1, lb_get_fft1_fused_fmt1_fc2},
      {lb_get_fft1_fused_fmt2_fc0, lb_get_fft1_fused_fmt2_fc1, lb_get_fft1_fused_fmt2_fc2},
      {lb_get_fft1_fused_fmt3_fc0, lb_get_fft1_fused_fmt3_fc1, lb_get_fft1_fused_fmt3_fc2}};
  if (fmt < 0 || fmt > 3 || fc < 0 || fc > 2) return nullptr;
  return g[fmt][fc](log2n, threads, smem);
}

// ---------------------------------------------------------------------------------------------
// four-step path: sub-batches of whole averaging groups, sized so that the intermediate Y stays
// in L2 between step A and step B
#include "fft1_large.cuh"
#include <cstdlib>
cudaError_t lb_launch_fft1_pipe(lb200_plan* plan, const Fft1K& k);
bool lb_fft1_pipe_supported(const lb200_plan* plan, const Fft1K& k);
cudaError_t lb_large_launch_fmt0(int, int, const Fft1LargeK&, int, cudaStream_t);
cudaError_t lb_large_launch_fmt1(int, int, const Fft1LargeK&, int, cudaStream_t);
cudaError_t lb_large_launch_fmt2(int, int, const Fft1LargeK&, int, cudaStream_t);
cudaError_t lb_large_launch_fmt3(int, int, const Fft1LargeK&, int, cudaStream_t);
cudaError_t lb_large_launch_fmt4(int, int, const Fft1LargeK&, int, cudaStream_t);
cudaError_t lb_large_launch_fmt5(int, int, const Fft1LargeK&, int, cudaStream_t);
cudaError_t lb_large_launch_fmt6(int, int, const Fft1LargeK&, int, cudaStream_t);
cudaError_t lb_large_launch_fmt7(int, int, const Fft1LargeK&, int, cudaStream_t);
typedef cudaError_t (*large_fn_t)(int, int, const Fft1LargeK&, int, cudaStream_t);
static large_fn_t large_fn(int fmt)
{
  static const large_fn_t t[8] = {lb_large_launch_fmt0, lb_large_launch_fmt1, lb_large_launch_fmt2, lb_large_launch_fmt3,
                                  lb_large_launch_fmt4, lb_large_launch_fmt5, lb_large_launch_fmt6, lb_large_launch_fmt7};
  return (fmt >= 0 && fmt < 8) ? t[fmt] : nullptr;
}

// zero the fft1_sumsq rows a launch will ADD into (all but a row that continues a partial group)
cudaError_t lb_zero_sumsq_rows(lb200_plan* plan, const Fft1K& k, int ngroups)
{
  const int g0 = k.counter0 > 0 ? 1 : 0;
  const size_t N = (size_t)plan->N;
  size_t off = (k.sumsq_pa + (size_t)g0 * N) & k.sumsq_mask;
  size_t len = (size_t)(ngroups - g0) * N;
  const size_t size = (size_t)k.sumsq_mask + 1;
  while (len > 0) {                                            // at most two pieces: the rows are consecutive on the ring
    size_t n = size - off;
    if (n > len) n = len;
    cudaError_t e = cudaMemsetAsync(k.sumsq + off, 0, sizeof(float) * n, plan->stream);
    if (e != cudaSuccess) return e;
    len -= n;
    off = 0;
  }
  return cudaSuccess;
}

bool lb_fft1_large_supported(int log2n) { return log2n >= 15 && log2n <= 20; }

static void large_split(int log2n, int* ln1, int* ln2)
{
  *ln1 = (log2n + 1) / 2;
  if (log2n == 15) *ln1 = 8;
  *ln2 = log2n - *ln1;
}

cudaError_t lb_launch_fft1_large(lb200_plan* plan, const Fft1K& k)
{
  const int log2n = plan->cfg.fft1_n;
  int ln1, ln2;
  large_split(log2n, &ln1, &ln2);
  const size_t N = (size_t)1 << log2n;
  const int nch = plan->nch;
  const int group = k.power_rows ? 1 : k.avg1num;
  const int c0 = k.power_rows ? 0 : k.counter0;
  const int ngroups = (c0 + k.nblocks + group - 1) / group;
  const char* env = getenv("LB200_SCRATCH_MB");
  const size_t budget = (size_t)(env ? atoi(env) : 48) << 20;
  size_t per_group = (size_t)group * nch * N * sizeof(float2);
  int gps = (int)(budget / per_group);
  if (gps < 1) gps = 1;
  const size_t need = (size_t)gps * group * nch * N;
  if (plan->scratch_elems < need) {
    if (plan->d_scratch) cudaFree(plan->d_scratch);
    plan->d_scratch = nullptr;
    plan->scratch_elems = 0;
    cudaError_t e = cudaMalloc((void**)&plan->d_scratch, need * sizeof(float2));
    if (e != cudaSuccess) return e;
    plan->scratch_elems = need;
  }
  Fft1LargeK q;
  q.k = k;
  q.scratch = plan->d_scratch;
  q.Wn1 = plan->d_Wn1;
  q.Wn2 = plan->d_Wn2;
  q.Wbig = plan->d_Wn;
  large_fn_t fn = large_fn(plan->fmt);
  const int tilesA = (1 << ln2) >> LB_LARGE_LTA, tilesB = (1 << ln1) >> LB_LARGE_LTB;
  if (k.sumsq && !k.power_rows && k.fc_mode != 0) {
    cudaError_t e = lb_zero_sumsq_rows(plan, k, ngroups);    // step B adds |z|^2 into the rows
    if (e != cudaSuccess) return e;
  }
  for (int g0 = 0; g0 < ngroups; g0 += gps) {
    const int g1 = g0 + gps < ngroups ? g0 + gps : ngroups;
    int b_first = g0 * group - c0;
    int b_last = g1 * group - c0;
    if (b_first < 0) b_first = 0;
    if (b_last > k.nblocks) b_last = k.nblocks;
    q.b_first = b_first;
    q.b_count = b_last - b_first;
    q.g_first = g0;
    q.g_count = g1 - g0;
    int gridA = q.b_count * nch * tilesA;
    int gridB = q.b_count * tilesB;
    const int cap = plan->sm_count * 16;
    if (gridA > cap) gridA = cap;
    if (gridB > cap) gridB = cap;
    cudaError_t e = fn(log2n, 0, q, gridA, plan->stream);
    if (e != cudaSuccess) return e;
    e = fn(log2n, 1, q, gridB, plan->stream);
    if (e != cudaSuccess) return e;
    plan->launches += 2;
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// real input (fft1_re.c): packed complex transform into an L2-resident Z buffer, then the
// untangle / output-mapping / fft1_c kernel, in sub-batches of whole averaging groups
#include "fft1_real.cuh"

static cudaError_t ensure_buf(float2** buf, size_t* have, size_t need)
{
  if (*have >= need) return cudaSuccess;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *have = 0;
  cudaError_t e = cudaMalloc((void**)buf, need * sizeof(float2));
  if (e == cudaSuccess) *have = need;
  return e;
}

cudaError_t lb_launch_fft1_real(lb200_plan* plan, const Fft1K& k)
{
  const int log2n = plan->cfg.fft1_n;
  const bool large = log2n > 14;
  const size_t N = (size_t)1 << log2n;
  const int nch = plan->nch;
  const int group = k.power_rows ? 1 : k.avg1num;
  const int c0 = k.power_rows ? 0 : k.counter0;
  const int ngroups = (c0 + k.nblocks + group - 1) / group;
  // Z of a sub-batch.  With the persistent four-step kernel fewer, longer launches win over keeping Z in L2
  // (configs[2], 740 transforms: 0.367 ms in 4 sub-batches of 48 MB, 0.337 in 2, 0.328 in one of 194 MB), so the
  // default budget is 256 MB there; the two-kernel path keeps its L2-sized sub-batches.
  const bool piped = large && lb_fft1_pipe_supported(plan, k);
  const char* env = getenv("LB200_SCRATCH_MB");
  const size_t budget = (size_t)(env ? atoi(env) : (piped ? 256 : 48)) << 20;
  const size_t per_group = (size_t)group * nch * N * sizeof(float2) * ((large && !piped) ? 2 : 1);
  int gps = (int)(budget / per_group);
  if (gps < 1) gps = 1;
  if (gps > ngroups) gps = ngroups;                    // no more than this call needs
  const size_t need = (size_t)gps * group * nch * N;
  cudaError_t e = ensure_buf(&plan->d_zbuf, &plan->zbuf_elems, need);
  if (e != cudaSuccess) return e;
  if (large && !lb_fft1_pipe_supported(plan, k)) {
    e = ensure_buf(&plan->d_scratch, &plan->scratch_elems, need);
    if (e != cudaSuccess) return e;
  }
  int ln1 = 0, ln2 = 0;
  if (large) large_split(log2n, &ln1, &ln2);
  for (int g0 = 0; g0 < ngroups; g0 += gps) {
    const int g1 = g0 + gps < ngroups ? g0 + gps : ngroups;
    int b_first = g0 * group - c0;
    int b_last = g1 * group - c0;
    if (b_first < 0) b_first = 0;
    if (b_last > k.nblocks) b_last = k.nblocks;
    const int b_count = b_last - b_first;
    if (b_count <= 0) continue;
    // ---- transform: every block on its own, plain Z to zbuf
    Fft1K k1 = k;
    k1.fc_mode = 0;
    k1.sumsq = nullptr;
    k1.power_rows = nullptr;
    k1.avg1num = 1;
    k1.counter0 = 0;
    k1.zbuf = plan->d_zbuf;
    if (!large) {
      int threads = 0;
      size_t smem = 0;
      fft1_small_launch_t fn = lb_get_fft1_small(log2n, plan->fmt, 0, &threads, &smem);
      if (!fn) return cudaErrorNotSupported;
      k1.ref0 = k.ref0 + (uint32_t)b_first * k.blockbytes;
      k1.nblocks = b_count;
      k1.zb_first = 0;
      int ctas = (int)((227 * 1024) / (smem + 1024));
      if (ctas > 2048 / threads) ctas = 2048 / threads;
      if (ctas < 1) ctas = 1;
      int grid = b_count;
      if (grid > plan->sm_count * ctas) grid = plan->sm_count * ctas;
      e = fn(k1, grid, plan->stream);
      if (e != cudaSuccess) return e;
      plan->launches += 1;
    } else if (lb_fft1_pipe_supported(plan, k)) {
      k1.ref0 = k.ref0 + (uint32_t)b_first * k.blockbytes;
      k1.nblocks = b_count;
      k1.zb_first = 0;
      e = lb_launch_fft1_pipe(plan, k1);    // counts its own launch
      if (e != cudaSuccess) return e;
    } else {
      Fft1LargeK q;
      k1.zb_first = b_first;
      q.k = k1;
      q.scratch = plan->d_scratch;
      q.Wn1 = plan->d_Wn1;
      q.Wn2 = plan->d_Wn2;
      q.Wbig = plan->d_Wn;
      q.b_first = b_first;
      q.b_count = b_count;
      q.g_first = b_first;          // groups of one transform (avg1num = 1, counter0 = 0)
      q.g_count = b_count;
      const int tilesA = (1 << ln2) >> LB_LARGE_LTA, tilesB = (1 << ln1) >> LB_LARGE_LTB;
      int gridA = b_count * nch * tilesA, gridB = b_count * tilesB;
      const int cap = plan->sm_count * 8;
      if (gridA > cap) gridA = cap;
      if (gridB > cap) gridB = cap;
      large_fn_t fn = large_fn(plan->fmt);
      e = fn(log2n, 0, q, gridA, plan->stream);
      if (e != cudaSuccess) return e;
      e = fn(log2n, 1, q, gridB, plan->stream);
      if (e != cudaSuccess) return e;
      plan->launches += 2;
    }
    // ---- untangle + output mapping + fft1_c
    Fft1K kp = k;
    kp.zbuf = plan->d_zbuf;
    kp.zb_first = b_first;
    const int chunks = (int)((N + 255) / 256);
    int grid = (g1 - g0) * chunks;
    if (grid > plan->sm_count * 16) grid = plan->sm_count * 16;
    if (nch == 1) fft1_real_post_kernel<1><<<grid, 256, 0, plan->stream>>>(kp, log2n, b_first, b_count, g0, g1 - g0);
    else fft1_real_post_kernel<2><<<grid, 256, 0, plan->stream>>>(kp, log2n, b_first, b_count, g0, g1 - g0);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    plan->launches += 1;
  }
  return cudaSuccess;
}
