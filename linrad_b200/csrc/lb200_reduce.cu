// lb200_reduce.cu -- sum of the averaged power spectra of several GPUs (the
// reference has no multi-GPU path: this is the only exchange the sharded hot path needs, the
// spectrum the wide graph of ONE Linrad instance draws from all receiver streams).
//
// One process per GPU.  Nothing runs on the SMs except one small add kernel on the root:
//   push : every rank copies its rows into its slot of the ROOT's mailbox with the copy engine
//          (cudaMemcpyAsync device -> peer device over NVLink, IPC-mapped), then stores the round
//          number into its flag word there (one-thread kernel);
//   sum  : the root waits for the round's flags with stream memory operations
//          (cuStreamWaitValue32: no host synchronisation, no spinning CTA), adds the world slots
//          and acknowledges the round in every peer's mailbox (flow control for the slot ring).
// Everything is queued on a side stream of the reducer, ordered against the plan's stream by
// events, so the persistent fft1 kernels of the next batch never share their SMs with a collective.
// The mailbox handles (cudaIpcMemHandle_t, 64 bytes) are exchanged by the caller over whatever
// transport it has (torch.distributed in bench.py / the tests, a pipe or MPI in a Linrad-side host).
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "plan.h"

struct lb200_reduce {
  lb200_plan* plan = nullptr;
  int rank = 0, world = 1, root = 0, depth = 2;
  size_t floats = 0;
  // this rank's mailbox: [world flags (64 B apart)][ack (64 B)][world * depth * floats]
  unsigned char* box = nullptr;
  size_t box_bytes = 0;
  std::vector<unsigned char*> peer;       // mapped mailboxes (peer[rank] == box)
  cudaStream_t rs = nullptr;
  cudaEvent_t e_rows = nullptr, e_pushed = nullptr, e_sum = nullptr;
  uint32_t round = 0;                     // rounds pushed by this rank
  uint32_t summed = 0;                    // rounds summed (root)
  bool connected = false;
  bool have_pushed = false, have_sum = false;
};

static const size_t kHdr = 64;
static size_t hdr_bytes(int world) { return kHdr * (size_t)(world + 1); }
static uint32_t* flag_of(unsigned char* box, int r) { return reinterpret_cast<uint32_t*>(box + kHdr * (size_t)r); }
static uint32_t* ack_of(unsigned char* box, int world) { return reinterpret_cast<uint32_t*>(box + kHdr * (size_t)world); }
static float* slot_of(unsigned char* box, int world, int depth, size_t floats, int r, uint32_t round)
{
  return reinterpret_cast<float*>(box + hdr_bytes(world)) + ((size_t)r * depth + (round % (uint32_t)depth)) * floats;
}

typedef CUresult (*wait32_fn_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static wait32_fn_t get_wait32()
{
  static wait32_fn_t fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (wait32_fn_t)p;
    else
      cudaGetLastError();
  }
  return fn;
}

__global__ void reduce_store_flag(uint32_t* dst, uint32_t value)
{
  __threadfence_system();
  *reinterpret_cast<volatile uint32_t*>(dst) = value;
  __threadfence_system();
}
// fallback when stream memory operations are unavailable: one thread polls (bounded)
__global__ void reduce_wait_flag(const uint32_t* src, uint32_t value)
{
  const long long t0 = clock64();
  while ((int32_t)(*reinterpret_cast<const volatile uint32_t*>(src) - value) < 0) {
    __nanosleep(200);
    if (clock64() - t0 > (1ll << 33)) return;
  }
}
// out[i] = sum over ranks of their slot, in rank order (deterministic)
__global__ void __launch_bounds__(256) reduce_add_kernel(float* __restrict__ out, const float* __restrict__ slots, size_t floats, size_t rank_stride,
                                                          int world)
{
  const size_t n4 = floats / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = __ldcs(reinterpret_cast<const float4*>(slots) + i);
    for (int r = 1; r < world; r++) {
      const float4 b = __ldcs(reinterpret_cast<const float4*>(slots + (size_t)r * rank_stride) + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t i = n4 * 4; i < floats; i++) {
      float a = slots[i];
      for (int r = 1; r < world; r++) a += slots[(size_t)r * rank_stride + i];
      out[i] = a;
    }
}

static int wait_value(lb200_reduce* r, uint32_t* addr, uint32_t value)
{
  lb200_plan* plan = r->plan;
  wait32_fn_t w = getenv("LB200_REDUCE_SPIN") ? nullptr : get_wait32();
  if (w) {
    if (w((CUstream)r->rs, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
      plan->last_cuda_error = -1;
      return LB200_ERR_CUDA;
    }
    return 0;
  }
  reduce_wait_flag<<<1, 1, 0, r->rs>>>(addr, value);
  LB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lb200_reduce_create(lb200_plan* plan, int rank, int world, size_t floats, int depth, lb200_reduce** out)
{
  if (!plan || !out || world < 1 || rank < 0 || rank >= world || floats == 0 || depth < 1 || depth > 64) return LB200_ERR_BAD_ARG;
  *out = nullptr;
  cudaSetDevice(plan->device);
  lb200_reduce* r = new lb200_reduce();
  r->plan = plan;
  r->rank = rank;
  r->world = world;
  r->depth = depth;
  r->floats = floats;
  r->box_bytes = hdr_bytes(world) + (size_t)world * depth * floats * sizeof(float);
  auto fail = [&](int code) { if (r->box) cudaFree(r->box); delete r; return code; };
  if (cudaMalloc((void**)&r->box, r->box_bytes) != cudaSuccess) return fail(LB200_ERR_CUDA);
  if (cudaMemset(r->box, 0, hdr_bytes(world)) != cudaSuccess) return fail(LB200_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&r->rs, cudaStreamNonBlocking) != cudaSuccess) return fail(LB200_ERR_CUDA);
  cudaEventCreateWithFlags(&r->e_rows, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&r->e_pushed, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&r->e_sum, cudaEventDisableTiming);
  r->peer.assign(world, nullptr);
  r->peer[rank] = r->box;
  if (world == 1) { r->connected = true; r->root = 0; }
  *out = r;
  return LB200_OK;
}

extern "C" int lb200_reduce_export(lb200_reduce* r, void* handle64)
{
  if (!r || !handle64) return LB200_ERR_BAD_ARG;
  lb200_plan* plan = r->plan;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaSetDevice(plan->device);
  cudaIpcMemHandle_t h;
  LB_CUDA(cudaIpcGetMemHandle(&h, r->box));
  memcpy(handle64, &h, 64);
  return LB200_OK;
}

extern "C" int lb200_reduce_connect(lb200_reduce* r, const void* handles, int root)
{
  if (!r || !handles || root < 0 || root >= r->world) return LB200_ERR_BAD_ARG;
  lb200_plan* plan = r->plan;
  cudaSetDevice(plan->device);
  r->root = root;
  // the root writes acknowledgements into every mailbox, the others only write into the root's
  for (int i = 0; i < r->world; i++) {
    if (i == r->rank) continue;
    if (r->rank != root && i != root) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)handles + 64 * (size_t)i, 64);
    void* p = nullptr;
    LB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    r->peer[i] = (unsigned char*)p;
  }
  r->connected = true;
  return LB200_OK;
}

// every rank: `rows` (device, floats of create) as this rank's contribution to the next round.
// The rows are read behind everything queued on the plan's stream so far; they must stay
// untouched until lb200_reduce_rows_released() has been queued in front of their next writer.
extern "C" int lb200_reduce_push(lb200_reduce* r, const float* rows)
{
  if (!r || !rows || !r->connected) return LB200_ERR_BAD_ARG;
  lb200_plan* plan = r->plan;
  cudaSetDevice(plan->device);
  // the root's own slot ring is protected by call order only: it may run `depth` rounds ahead of its sums
  if (r->rank == r->root && r->round >= r->summed + (uint32_t)r->depth) return LB200_ERR_BAD_ARG;
  const uint32_t k = ++r->round;
  LB_CUDA(cudaEventRecord(r->e_rows, plan->stream));
  LB_CUDA(cudaStreamWaitEvent(r->rs, r->e_rows, 0));
  // flow control: the slot of this round was last used by round k - depth, which the root must have summed
  if (k > (uint32_t)r->depth && r->rank != r->root) {
    int rc = wait_value(r, ack_of(r->box, r->world), k - (uint32_t)r->depth);
    if (rc) return rc;
  }
  unsigned char* rb = r->peer[r->root];
  LB_CUDA(cudaMemcpyAsync(slot_of(rb, r->world, r->depth, r->floats, r->rank, k), rows, r->floats * sizeof(float), cudaMemcpyDefault, r->rs));
  reduce_store_flag<<<1, 1, 0, r->rs>>>(flag_of(rb, r->rank), k);
  LB_CUDA(cudaGetLastError());
  LB_CUDA(cudaEventRecord(r->e_pushed, r->rs));
  r->have_pushed = true;
  plan->launches++;
  return LB200_OK;
}

// queue on the plan's stream: wait until the last pushed rows have been copied out
extern "C" int lb200_reduce_rows_released(lb200_reduce* r)
{
  if (!r) return LB200_ERR_BAD_ARG;
  lb200_plan* plan = r->plan;
  if (r->have_pushed) LB_CUDA(cudaStreamWaitEvent(plan->stream, r->e_pushed, 0));
  return LB200_OK;
}

// root: out = sum over ranks of the next round's rows (in rank order).  Queued on the reducer's
// stream; lb200_reduce_result_ready() orders the plan's stream behind it.
extern "C" int lb200_reduce_sum(lb200_reduce* r, float* out)
{
  if (!r || !out || !r->connected || r->rank != r->root) return LB200_ERR_BAD_ARG;
  lb200_plan* plan = r->plan;
  cudaSetDevice(plan->device);
  const uint32_t k = ++r->summed;
  // a previous result in `out` may still be read by work queued on the plan's stream
  LB_CUDA(cudaEventRecord(r->e_rows, plan->stream));
  LB_CUDA(cudaStreamWaitEvent(r->rs, r->e_rows, 0));
  for (int i = 0; i < r->world; i++) {
    int rc = wait_value(r, flag_of(r->box, i), k);
    if (rc) return rc;
  }
  const float* slots = slot_of(r->box, r->world, r->depth, r->floats, 0, k);
  int grid = (int)((r->floats / 4 + 255) / 256);
  if (grid > plan->sm_count * 4) grid = plan->sm_count * 4;
  if (grid < 1) grid = 1;
  reduce_add_kernel<<<grid, 256, 0, r->rs>>>(out, slots, r->floats, (size_t)r->depth * r->floats, r->world);
  LB_CUDA(cudaGetLastError());
  plan->launches++;
  for (int i = 0; i < r->world; i++) {
    if (i == r->rank) continue;
    reduce_store_flag<<<1, 1, 0, r->rs>>>(ack_of(r->peer[i], r->world), k);
    LB_CUDA(cudaGetLastError());
  }
  LB_CUDA(cudaEventRecord(r->e_sum, r->rs));
  r->have_sum = true;
  return LB200_OK;
}

extern "C" int lb200_reduce_result_ready(lb200_reduce* r)
{
  if (!r) return LB200_ERR_BAD_ARG;
  lb200_plan* plan = r->plan;
  if (r->have_sum) LB_CUDA(cudaStreamWaitEvent(plan->stream, r->e_sum, 0));
  return LB200_OK;
}

extern "C" int lb200_reduce_synchronize(lb200_reduce* r)
{
  if (!r) return LB200_ERR_BAD_ARG;
  lb200_plan* plan = r->plan;
  LB_CUDA(cudaStreamSynchronize(r->rs));
  return LB200_OK;
}

extern "C" void lb200_reduce_destroy(lb200_reduce* r)
{
  if (!r) return;
  cudaSetDevice(r->plan->device);
  if (r->rs) cudaStreamSynchronize(r->rs);
  for (int i = 0; i < r->world; i++)
    if (i != r->rank && r->peer[i]) cudaIpcCloseMemHandle(r->peer[i]);
  if (r->box) cudaFree(r->box);
  if (r->e_rows) cudaEventDestroy(r->e_rows);
  if (r->e_pushed) cudaEventDestroy(r->e_pushed);
  if (r->e_sum) cudaEventDestroy(r->e_sum);
  if (r->rs) cudaStreamDestroy(r->rs);
  delete r;
}
