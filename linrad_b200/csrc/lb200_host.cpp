// lb200_host.cpp -- scalar host logic of the C ABI (no CUDA): the parts of the reference
// path that are sequential state updates rather than data-parallel work.
#include <math.h>
#include <string.h>
#include "../../include/linrad_b200.h"
#include "phase.h"

#define LB_PI 3.1415926535897932   /* PI_L, globdef.h:93 */

extern "C" int lb200_abi_version(void) { return LB200_ABI_VERSION; }

extern "C" const char* lb200_strerror(int code)
{
  switch (code) {
    case LB200_OK: return "ok";
    case LB200_ERR_NO_DEVICE: return "lb200: no usable CUDA device";
    case LB200_ERR_CUDA: return "lb200: CUDA runtime call or kernel launch failed";
    case LB200_ERR_BAD_CONFIG: return "lb200: inconsistent sizes in lb200_config";
    case LB200_ERR_UNSUPPORTED: return "lb200: size or mode not covered by the sm_100a kernels";
    case LB200_ERR_BAD_ARG: return "lb200: bad argument (null pointer, misaligned offset, ring too small)";
    case LB200_ERR_MIX1_RANGE_LOW: return "mix1 frequency below mix1_lowest_fq";
    case LB200_ERR_MIX1_RANGE_HIGH: return "mix1 frequency above mix1_highest_fq";
    default: return "lb200: unknown error";
  }
}

extern "C" float lb200_phase_advance(float phase, float rot, int count)
{
  return lb_phase_advance(phase, rot, count);
}

// set_mix1_phases, mix1.c:781-861 (single-precision branch), for one selection.
// Every intermediate keeps the reference's type: t1,t2 float; the products with PI_L are
// formed in double and rounded once when stored to the float state variables.
extern "C" int lb200_set_mix1_phases(const lb200_config* cfg, lb200_mix1_state* s, float fq)
{
  const unsigned int msize = 1u << cfg->mix1_n;
  const unsigned int mnew = msize - (unsigned int)cfg->mix1_interleave_points;
  if (fq < cfg->mix1_lowest_fq) return LB200_ERR_MIX1_RANGE_LOW;     /* mix1.c:787-791 */
  if (fq > cfg->mix1_highest_fq) return LB200_ERR_MIX1_RANGE_HIGH;   /* mix1.c:792-796 */
  volatile float t1 = fq * cfg->fftx_points_per_hz;                  /* mix1.c:799 */
  const int fftx_pnt = (int)(t1 + 0.5);                              /* mix1.c:800 */
  int k = (int)((unsigned int)fftx_pnt % msize);                     /* mix1.c:839 */
  volatile float t2 = (float)(msize * ((unsigned int)fftx_pnt / msize)); /* mix1.c:840 */
  t2 = t1 - t2 - (float)k;                                           /* mix1.c:841 */
  t2 = t2 - (float)(int)(t2);                                        /* mix1.c:842 */
  s->mix1_phase_rot = (float)(t2 * 2 * LB_PI / msize);               /* mix1.c:843 */
  k = (int)(((unsigned int)k * mnew) % msize);                       /* mix1.c:846 */
  s->mix1_old_phase = s->mix1_phase;                                 /* mix1.c:847 */
  {
    volatile float ph = s->mix1_phase + s->mix1_phase_step;          /* mix1.c:848 */
    s->mix1_phase = ph;
  }
  s->mix1_phase_step = (float)((unsigned int)(k * 2) * LB_PI / msize); /* mix1.c:849 */
  if (s->mix1_point != -1) s->mix1_old_point = s->mix1_point;        /* mix1.c:850-857 */
  else s->mix1_old_point = fftx_pnt;
  s->mix1_point = fftx_pnt;
  /* mix1.c:859-860 -- both tests compare against +PI_L (reference quirk, kept) */
  if ((double)s->mix1_phase > LB_PI) s->mix1_phase = (float)((double)s->mix1_phase - 2 * LB_PI);
  if ((double)s->mix1_phase < LB_PI) s->mix1_phase = (float)((double)s->mix1_phase + 2 * LB_PI);
  return LB200_OK;
}

// make_window layouts, fft0.c:812-921: mo=4 natural (size floats); mo=1 interleaved
// w[2i]=w(i), w[2i+1]=w(size/2+i) (fft0.c:905-920); mo=2 first half of a 2*size window
// (size+1 floats, symmetric).
extern "C" void lb200_window_to_natural(int mo, int size, const float* win, float* natural)
{
  int i;
  if (mo == 1) {
    for (i = 0; i < size / 2; i++) {
      natural[i] = win[2 * i];
      natural[size / 2 + i] = win[2 * i + 1];
    }
  } else if (mo == 2) {
    /* window of 2*size real samples, value i and its mirror 2*size-1-i share win[i] (fft1_re.c:48-57) */
    for (i = 0; i < size; i++) {
      natural[i] = win[i];
      natural[2 * size - 1 - i] = win[i];
    }
  } else {
    memcpy(natural, win, sizeof(float) * (size_t)size);
  }
}

// ---- Linrad .raw header (open_savefile, modesub.c:656-733) -----------------------------------
namespace {
struct Cursor {
  const unsigned char* p;
  size_t n, at;
  bool get(void* dst, size_t k)
  {
    if (at + k > n) return false;
    memcpy(dst, p + at, k);
    at += k;
    return true;
  }
};
}  // namespace

extern "C" int lb200_raw_header_parse(const void* bytes, size_t nbytes, lb200_raw_header* out)
{
  if (!bytes || !out) return LB200_ERR_BAD_ARG;
  Cursor c{(const unsigned char*)bytes, nbytes, 0};
  lb200_raw_header h;
  memset(&h, 0, sizeof(h));
  int first;
  if (!c.get(&first, 4)) return LB200_ERR_BAD_ARG;
  if (first < 0) {
    h.remember_tag = first;
    switch (first) {
      case LB200_REMEMBER_UNKNOWN:
      case LB200_REMEMBER_NOTHING:
        break;
      case LB200_REMEMBER_PERSEUS:
      case LB200_REMEMBER_SDR14:
        if (!c.get(&h.chunk_size, 4)) return LB200_ERR_BAD_ARG;
        if (h.chunk_size < 0 || c.at + (size_t)h.chunk_size > c.n) return LB200_ERR_BAD_ARG;
        h.chunk_offset = c.at;
        c.at += (size_t)h.chunk_size;
        break;
      default:
        return LB200_ERR_BAD_ARG;             // "This Linrad version is too old"
    }
    if (!c.get(&h.diskread_time, 8)) return LB200_ERR_BAD_ARG;
    if (!c.get(&h.passband_center, 8)) return LB200_ERR_BAD_ARG;
    if (!c.get(&h.passband_direction, 4)) return LB200_ERR_BAD_ARG;
    if (h.passband_direction != 1 && h.passband_direction != -1) return LB200_ERR_BAD_ARG;
    if (!c.get(&h.rx_input_mode, 4)) return LB200_ERR_BAD_ARG;
  } else {
    h.remember_tag = LB200_REMEMBER_NOTHING;
    h.rx_input_mode = first;
    h.passband_direction = 1;
  }
  if (h.rx_input_mode >= 256) return LB200_ERR_BAD_ARG;          // MODEPARM_MAX, globdef.h:285
  if (!c.get(&h.rx_rf_channels, 4)) return LB200_ERR_BAD_ARG;
  if (h.rx_rf_channels == 2) h.rx_input_mode |= LB200_TWO_CHANNELS;
  if (!c.get(&h.rx_ad_channels, 4)) return LB200_ERR_BAD_ARG;
  if (h.rx_ad_channels > 4 || h.rx_ad_channels < 1) return LB200_ERR_BAD_ARG;
  if (h.rx_ad_channels != h.rx_rf_channels && h.rx_ad_channels != 2 * h.rx_rf_channels) return LB200_ERR_BAD_ARG;
  if (!c.get(&h.rx_ad_speed, 4)) return LB200_ERR_BAD_ARG;
  unsigned char flag;
  if (!c.get(&flag, 1)) return LB200_ERR_BAD_ARG;
  h.save_init_flag = flag;
  h.payload_offset = c.at;
  *out = h;
  return LB200_OK;
}

extern "C" size_t lb200_raw_block_bytes(const lb200_raw_header* h, size_t block_bytes)
{
  if (!h) return 0;
  return (h->rx_input_mode & LB200_DWORD_INPUT) ? 18 * block_bytes / 32 : block_bytes;
}
