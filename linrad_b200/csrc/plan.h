// plan.h -- internal plan object behind the opaque lb200_plan of include/linrad_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <map>
#include "../../include/linrad_b200.h"

struct HostMirror {            // device mirror of one of Linrad's host rings
  void* d = nullptr;
  size_t bytes = 0;
  const void* host = nullptr;
  bool registered = false;
};

struct lb200_plan {
  lb200_config cfg;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  // derived geometry
  int N = 0;                   // fft1_size
  int nch = 1;                 // rx_rf_channels
  int mm = 2;                  // twice_rxchan
  int frame = 4;               // bytes per input frame
  int fmt = 0;                 // lb::InFmt
  bool iq = true;
  int fft1_block = 0;          // mm*N floats
  int new_points = 0;          // P
  uint32_t blockbytes = 0;     // timf1_blockbytes
  uint32_t pre_bytes = 0;      // bytes of the overlap span in front of timf1p_ref
  int M = 0;                   // mix1.size
  // device tables
  float* d_window = nullptr;
  float2* d_Wn = nullptr;      // fft1 twiddles, N entries
  float* d_filtercorr = nullptr;
  float* d_foldcorr = nullptr; // CALIQ mirror-image correction table (mm*N floats) or nullptr
  bool phasing = false;        // pg_ch2_c1/c2 != (1, 0)
  int first_sym = 0;           // fft1_first_sym_point (fft1.c:4647-4650)
  int shift_i = 0, shift_q = 0; // ui.sample_shift as frame offsets of the I and the Q word
  int fc_mode = 2;
  float fc_gain = 0.f;
  int fc_edge = 0;
  // fft1_fused_kernel tables (2^10 <= N <= 2^14)
  float* d_wsign = nullptr;    // window * (-1)^n
  float* d_wsign_g = nullptr;  // window * (-1)^n * fc_gain (FC_FOLDED)
  float2* d_edge = nullptr;    // filtercorr/fc_gain on the 16 outermost bins at each end
  float4* d_tab1 = nullptr;    // pass-1 twiddle table
  float2* d_scratch2 = nullptr; // fused kernel, 2 channels: one row of N float2 per resident CTA
  bool fc_foldable = false;    // fc_mode==1, all channels alike, gain != 0
  float2* d_Wm = nullptr;      // mix1 twiddles, M entries
  float* d_fqwin = nullptr;
  float* d_mixwin = nullptr;
  float* d_cos2win = nullptr;
  float* d_sin2win = nullptr;
  // large-N (four-step) scratch
  float2* d_scratch = nullptr;
  size_t scratch_elems = 0;
  float* d_powtmp = nullptr;   // small-batch path: per-transform |z|^2 rows
  size_t powtmp_bytes = 0;
  float2* d_zbuf = nullptr;    // real input: packed spectrum Z of one sub-batch
  size_t zbuf_elems = 0;
  float2* d_Wre = nullptr;     // real input: exp(-i pi k / N), k = 0..N
  float2* d_Wn1 = nullptr;     // four-step: exp(-2 pi i m / N1)
  float2* d_Wn2 = nullptr;     // four-step: exp(-2 pi i m / N2)
  // persistent four-step kernel (fft1_pipe.cuh)
  void* d_wT = nullptr;        // window transposed to [n2][n1] (IQ: float with (-1)^n folded in; real input: float2)
  int* d_pipe_sync = nullptr;  // queue head, error flag, per-transform completion counters
  size_t pipe_sync_ints = 0;
  float2* d_pipe_y = nullptr;  // the Y ring: [pipe_slots * nch][N2][N1]
  int pipe_slots = 0;          // transforms the Y ring holds
  alignas(64) unsigned char map_y[128];      // CUtensorMap of the Y ring
  alignas(64) unsigned char map_out[128];    // CUtensorMap of the output ring / zbuf of the last call
  const void* map_out_base = nullptr;
  size_t map_out_planes = 0;
  bool pipe_checked = false;   // the error flag of the last launch has been read back
  // make_timf2 (timf2.cuh)
  float* d_invwin = nullptr;        // fft1_inverted_window (windows other than none / sin^2)
  float4* d_tab1_any = nullptr;     // pass-1 twiddles of the 32-points-per-thread plan (any input kind)
  float2* d_timf2_tmp = nullptr;    // timf2_tmp of a whole call
  size_t timf2_tmp_elems = 0;
  void* d_mix_y = nullptr;          // mix1.size 16384 / 8192 (two channels): back-transformed blocks between the two launches
  size_t mix_y_bytes = 0;
  HostMirror m_t2_fft1, m_t2_ring, m_t2_pwr, m_t2_lim;
  // mix1 per-call staging: a small ring of pinned/device job tables so that consecutive
  // calls never wait for each other on the host
  static constexpr int kJobSlots = 4;
  void* d_mixjobs[kJobSlots] = {nullptr, nullptr, nullptr, nullptr};
  void* h_mixjobs[kJobSlots] = {nullptr, nullptr, nullptr, nullptr};   // pinned
  size_t mixjobs_bytes[kJobSlots] = {0, 0, 0, 0};
  cudaEvent_t mixjobs_done[kJobSlots] = {nullptr, nullptr, nullptr, nullptr};
  int mixjobs_next = 0;
  // host-pointer API mirrors
  HostMirror m_timf1, m_fft1, m_sumsq, m_timf3, m_power, m_corrsum, m_corr, m_xy;
  HostMirror m_wg_sumsq, m_wg_slowsum, m_wg_wsum, m_wg_yfac, m_wg_waterf, m_codec_in, m_codec_out;
  std::map<const void*, size_t> registered;
  // pipelined host path: copy streams, event pool, validity of the fft1_float mirror per block
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> events;
  std::vector<uint8_t> fft1_valid;
  std::vector<uint8_t> fft1_host_stale;   // block produced with LB200_FFT1_SPECTRUM_STAYS_ON_DEVICE: the host ring does not have it
  const void* fft1_valid_host = nullptr;
  // counters
  uint64_t launches = 0, h2d = 0, d2h = 0;
  int last_cuda_error = 0;
};

#define LB_CUDA(call)                                                 \
  do {                                                                \
    cudaError_t e_ = (call);                                          \
    if (e_ != cudaSuccess) {                                          \
      plan->last_cuda_error = (int)e_;                                \
      fprintf(stderr, "[lb200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return LB200_ERR_CUDA;                                          \
    }                                                                 \
  } while (0)
