/* oracle/oracle_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the integer/byte stages of the path (the float stages are restated
 * in oracle/port.py with numpy).  Parity pin: the 18-bit codec has no compiled reference here
 * (the authoritative version is NASM, getiq64.s, and nasm is not installed), so it is pinned
 * by the compress->expand round-trip identity and by hand-derived vectors in
 * tests/test_oracle_cpu.py; the float chain is pinned against oracle/_ref (mix1 phase state).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

/* expand_rawdat, getiq64.s:158-220: every 9 input bytes hold four 18-bit samples: bytes
 * 2i,2i+1 are bits 16..31 of word i, byte 8 carries bits 14,15 of the four words (word i in
 * bits 2i,2i+1).  Half an LSB (0x2000) is added to undo the truncation bias. */
void port_expand_rawdat(const uint8_t *packed, int32_t *out, size_t out_bytes)
{
  size_t g, groups = out_bytes / 16;
  for (g = 0; g < groups; g++) {
    const uint8_t *p = packed + 9 * g;
    unsigned b8 = p[8];
    int i;
    for (i = 0; i < 4; i++) {
      uint32_t hi = (uint32_t)p[2 * i] | ((uint32_t)p[2 * i + 1] << 8);
      uint32_t w = (hi << 16) | (((b8 >> (2 * i)) & 3u) << 14);
      w += 0x2000u;
      out[4 * g + i] = (int32_t)w;
    }
  }
}

/* compress_rawdat_disk / _net, getiq64.s:39-96,99-155: the inverse packing (bits 0..13 are
 * dropped). */
void port_compress_rawdat(const int32_t *in, uint8_t *packed, size_t in_bytes)
{
  size_t g, groups = in_bytes / 16;
  for (g = 0; g < groups; g++) {
    uint8_t *p = packed + 9 * g;
    unsigned b8 = 0;
    int i;
    for (i = 0; i < 4; i++) {
      uint32_t w = (uint32_t)in[4 * g + i];
      p[2 * i] = (uint8_t)(w >> 16);
      p[2 * i + 1] = (uint8_t)(w >> 24);
      b8 |= ((w >> 14) & 3u) << (2 * i);
    }
    p[8] = (uint8_t)b8;
  }
}

/* 24-bit PCM widening of rx_file_input, rxin.c:1603-1614: 3 little-endian bytes -> int32
 * left-justified (low byte zero). */
void port_widen_24bit(const uint8_t *in, int32_t *out, size_t nsamples)
{
  size_t i;
  for (i = 0; i < nsamples; i++) {
    uint32_t w = ((uint32_t)in[3 * i] << 8) | ((uint32_t)in[3 * i + 1] << 16) | ((uint32_t)in[3 * i + 2] << 24);
    out[i] = (int32_t)w;
  }
}

/* 8-bit PCM widening of rx_file_input, rxin.c:1573-1583, with the reference's own types
 * (rxin_char is a plain char pointer, rxin_isho a short pointer, fft1def.h:143-145):
 *   rxin_isho[j]=(rxin_char[j]<<8)-32640; */
void port_widen_8bit(const char *rxin_char, short int *rxin_isho, size_t nsamples)
{
  size_t j = nsamples;
  while (j > 0) {
    j--;
    rxin_isho[j] = (short int)((rxin_char[j] << 8) - 32640);
  }
}

/* float wav samples, rxin.c:1624-1634:  rxin_int[j]=0x7fffffff*z[j];  (int * float -> float,
 * converted back with the host's truncating conversion; out of range gives 0x80000000 on x86-64) */
void port_float_to_int32(const float *z, int *rxin_int, size_t nsamples)
{
  size_t j = nsamples;
  while (j > 0) {
    j--;
    rxin_int[j] = 0x7fffffff * z[j];
  }
}

/* the running single-precision phase sum of do_mix1 (mix1.c:146-153,172-186), literally */
float port_phase_chain(float phase, float rot, int count, float *trace)
{
  volatile float t1 = phase;
  int i;
  for (i = 0; i < count; i++) {
    if (trace) trace[i] = t1;
    t1 = t1 + rot;
  }
  return t1;
}
