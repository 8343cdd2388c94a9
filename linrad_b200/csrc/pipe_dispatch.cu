// pipe_dispatch.cu -- host side of the persistent four-step kernel (fft1_pipe.cuh): Y ring, queue
// counters, tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no link
// against libcuda), choice of queue lag / ring depth, and the per-format launchers.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "fft1_pipe.cuh"
#include "plan.h"
using namespace lb;

#define LB_PIPE_DECL(F) cudaError_t lb_pipe_launch_fmt##F(int, int*, const Fft1PipeK*, const CUtensorMap*, const CUtensorMap*, int, int*, cudaStream_t);
LB_PIPE_DECL(0) LB_PIPE_DECL(1) LB_PIPE_DECL(2) LB_PIPE_DECL(4) LB_PIPE_DECL(5) LB_PIPE_DECL(6)
typedef cudaError_t (*pipe_fn_t)(int, int*, const Fft1PipeK*, const CUtensorMap*, const CUtensorMap*, int, int*, cudaStream_t);

static pipe_fn_t pipe_fn(int fmt)
{
  switch (fmt) {
    case 0: return lb_pipe_launch_fmt0;
    case 1: return lb_pipe_launch_fmt1;
    case 2: return lb_pipe_launch_fmt2;
    case 4: return lb_pipe_launch_fmt4;
    case 5: return lb_pipe_launch_fmt5;
    case 6: return lb_pipe_launch_fmt6;
  }
  return nullptr;                       // 16-byte frames (two-channel int32): the raw tile does not fit, legacy kernels
}

static int env_i(const char* name, int dflt)
{
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_fn_t get_encode()
{
  static encode_fn_t fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (encode_fn_t)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// [planes][N2][N1] float2 seen as 32-bit words: dims (2*N1, N2, planes), box (2*TB, rows, 1)
static bool encode_map(void* out128, void* base, int ln1, int ln2, size_t planes, int tb, int box_rows)
{
  encode_fn_t enc = get_encode();
  if (!enc || planes == 0 || planes > 0xffffffffull || ((uintptr_t)base & 15u)) return false;
  const cuuint64_t n1 = 1ull << ln1, n2 = 1ull << ln2;
  cuuint64_t dims[3] = {2 * n1, n2, (cuuint64_t)planes};
  cuuint64_t strides[2] = {n1 * 8, n1 * n2 * 8};
  cuuint32_t box[3] = {(cuuint32_t)(2 * tb), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  memcpy(out128, &m, sizeof(m));
  return true;
}

bool lb_fft1_pipe_supported(const lb200_plan* plan, const Fft1K& k)
{
  if (env_i("LB200_LARGE_LEGACY", 0)) return false;
  const int ln = plan->cfg.fft1_n;
  if (ln < 15 || ln > 20 || !pipe_fn(plan->fmt) || !plan->d_wT) return false;
  if (k.skew_i | k.skew_q) return false;                         // ui.sample_shift: legacy kernels
  // the raw tile is fetched in 16-byte pieces
  if (((k.ref0 - k.pre_bytes) & 15u) || (k.blockbytes & 15u) || ((uintptr_t)k.timf1 & 15u) || k.ring_mask < 15u) return false;
  return true;
}

// after a synchronisation: did a dependency wait of the last launches give up?
int lb_fft1_pipe_status(lb200_plan* plan)
{
  if (!plan->d_pipe_sync || plan->pipe_checked) return 0;
  int flag = 0;
  if (cudaMemcpy(&flag, plan->d_pipe_sync + 1, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  plan->pipe_checked = true;
  return flag;
}

cudaError_t lb_launch_fft1_pipe(lb200_plan* plan, const Fft1K& k)
{
  const int ln = plan->cfg.fft1_n;
  pipe_fn_t fn = pipe_fn(plan->fmt);
  if (!fn) return cudaErrorNotSupported;
  int geo[8];
  cudaError_t e = fn(ln, geo, nullptr, nullptr, nullptr, 0, nullptr, plan->stream);
  if (e != cudaSuccess) return e;
  const int IA = geo[0], IB = geo[1], TB = geo[2], box_in = geo[3], box_out = geo[4], ln1 = geo[5], ln2 = geo[6];
  const size_t N = (size_t)1 << ln;
  const int nch = plan->nch;
  const int nb = k.nblocks;

  // ---- queue lag and ring depth: a dependency should be long satisfied when its consumer is claimed.
  // About 2 items per resident CTA are in flight (the one computed and the one prefetched).
  const int resident = plan->sm_count * (512 / geo[7]);       // sixteen warps per SM (128 registers per thread)
  const int per_phase = IA + IB;
  int lag = env_i("LB200_PIPE_LAG", (3 * resident / 2 + per_phase - 1) / per_phase + 1);
  if (lag < 1) lag = 1;
  int slots = env_i("LB200_PIPE_SLOTS", 2 * lag);
  if (slots < lag + 1) slots = lag + 1;
  if (slots > nb) slots = nb;                                  // a short call never wraps the ring
  if (slots < 1) slots = 1;
  if (plan->pipe_slots < slots || !plan->d_pipe_y) {
    if (plan->d_pipe_y) cudaFree(plan->d_pipe_y);
    plan->d_pipe_y = nullptr;
    plan->pipe_slots = 0;
    // allocate the steady-state depth at once so that the map is not rebuilt call after call
    int want = slots;
    const int full = 2 * ((3 * resident / 2 + per_phase - 1) / per_phase + 1);
    if (want < full && !getenv("LB200_PIPE_SLOTS")) want = full;
    e = cudaMalloc((void**)&plan->d_pipe_y, (size_t)want * nch * N * sizeof(float2));
    if (e != cudaSuccess) return e;
    plan->pipe_slots = want;
    if (!encode_map(plan->map_y, plan->d_pipe_y, ln1, ln2, (size_t)want * nch, TB, box_in)) memset(plan->map_y, 0, sizeof(plan->map_y));
  }
  const size_t need_ints = 2 + 2 * (size_t)nb;
  if (plan->pipe_sync_ints < need_ints) {
    if (plan->d_pipe_sync) {
      if (lb_fft1_pipe_status(plan)) fprintf(stderr, "[lb200] four-step pipeline: a dependency wait timed out in an earlier call\n");
      cudaFree(plan->d_pipe_sync);
    }
    plan->d_pipe_sync = nullptr;
    plan->pipe_sync_ints = 0;
    e = cudaMalloc((void**)&plan->d_pipe_sync, need_ints * sizeof(int) * 2);
    if (e != cudaSuccess) return e;
    plan->pipe_sync_ints = need_ints * 2;
    e = cudaMemsetAsync(plan->d_pipe_sync, 0, 2 * sizeof(int), plan->stream);
    if (e != cudaSuccess) return e;
  }
  // head and counters start at zero; the error flag [1] is sticky until read back
  e = cudaMemsetAsync(plan->d_pipe_sync, 0, sizeof(int), plan->stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(plan->d_pipe_sync + 2, 0, 2 * (size_t)nb * sizeof(int), plan->stream);
  if (e != cudaSuccess) return e;
  plan->pipe_checked = false;

  Fft1PipeK q;
  memset(&q, 0, sizeof(q));
  q.k = k;
  q.Y = plan->d_pipe_y;
  q.wT = plan->d_wT;
  q.Wn1 = plan->d_Wn1;
  q.Wn2 = plan->d_Wn2;
  q.Wbig = plan->d_Wn;
  q.sync = plan->d_pipe_sync;
  q.nslots = slots;
  q.lag = lag;
  q.prefetch_ahead = env_i("LB200_PIPE_PREFETCH", 3);
  static const unsigned char zero_map[128] = {0};
  const bool have_map_y = memcmp(plan->map_y, zero_map, 128) != 0;
  q.tma_in = (have_map_y && env_i("LB200_PIPE_TMA_IN", 1)) ? 1 : 0;
  // ---- output by TMA tensor stores: planar targets only (one channel, or the packed spectrum of real input)
  q.tma_out = 0;
  if (env_i("LB200_PIPE_TMA_OUT", 0) && (k.zbuf || nch == 1)) {   // measured: streaming stores are faster (0.150 vs 0.167 ms at configs[3], profiles/r2_notes.txt)
    void* base = k.zbuf ? (void*)k.zbuf : (void*)k.out;
    const size_t planes = k.zbuf ? plan->zbuf_elems / N : ((size_t)k.out_mask + 1) / (2 * N);
    if (planes >= 1 && (k.zbuf || (k.out_pa % (2 * N)) == 0)) {
      if (plan->map_out_base != base || plan->map_out_planes != planes) {
        if (encode_map(plan->map_out, base, ln1, ln2, planes, TB, box_out)) {
          plan->map_out_base = base;
          plan->map_out_planes = planes;
        } else {
          plan->map_out_base = nullptr;
          plan->map_out_planes = 0;
        }
      }
      if (plan->map_out_base == base && plan->map_out_planes == planes) {
        q.tma_out = 1;
        q.out_blk0 = k.zbuf ? 0u : (uint32_t)((k.out_pa & k.out_mask) / (2 * N));
        q.out_nblk = (uint32_t)planes;
      }
    }
  }
  // |z|^2 is ADDED into the rows
  extern cudaError_t lb_zero_sumsq_rows(lb200_plan*, const Fft1K&, int);
  if (!k.zbuf && k.fc_mode != 0) {
    if (k.power_rows && nch == 2) {
      e = cudaMemsetAsync(k.power_rows, 0, sizeof(float) * N * nb, plan->stream);
      if (e != cudaSuccess) return e;
    } else if (k.sumsq && !k.power_rows) {
      const int group = k.avg1num;
      const int ngroups = (k.counter0 + nb + group - 1) / group;
      e = lb_zero_sumsq_rows(plan, k, ngroups);
      if (e != cudaSuccess) return e;
    }
  }
  CUtensorMap my, mo;
  memcpy(&my, plan->map_y, sizeof(my));
  memcpy(&mo, plan->map_out, sizeof(mo));
  int grid[2] = {0, 0};
  e = fn(ln, nullptr, &q, &my, &mo, plan->sm_count, grid, plan->stream);
  if (e != cudaSuccess) return e;
  plan->launches += 1;
  return cudaSuccess;
}
