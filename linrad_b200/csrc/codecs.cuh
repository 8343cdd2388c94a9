// codecs.cuh -- input sample codecs of Linrad's file playback, bit-exact integer work:
//   expand_rawdat  (asm, getiq64.s:158-220): 18-bit packed .raw -> left-justified int32.  Every
//       9 input bytes hold four samples: bytes 2i,2i+1 are bits 16..31 of word i, byte 8 carries
//       bits 14,15 of the four words (word i in its bits 2i,2i+1); 0x2000 (half an 18-bit LSB)
//       is added to remove the truncation bias.
//   24-bit PCM widening (rxin.c:1603-1614): 3 little-endian bytes -> int32 << 8, low byte zero.
// Both are pure streaming: 9 -> 16 and 12 -> 16 bytes per four samples, HBM-bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lb {

// One tile = 256 groups = 2304 packed bytes, staged through shared memory with aligned 32-bit
// loads (2304 is a multiple of 4; the caller's buffer is 4-byte aligned), then one group per
// thread and one 16-byte store per thread.
__global__ void __launch_bounds__(256) expand_rawdat_kernel(const uint8_t* __restrict__ packed, int4* __restrict__ out, size_t groups)
{
  __shared__ uint32_t tile[576 + 4];
  const size_t ntiles = (groups + 255) / 256;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const size_t g0 = t * 256;
    const size_t n = groups - g0 < 256 ? groups - g0 : 256;
    const size_t bytes = n * 9;
    const uint8_t* src = packed + g0 * 9;
    __syncthreads();
    if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
      const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
      for (int i = threadIdx.x; i < (int)(bytes / 4); i += 256) tile[i] = s4[i];
      uint8_t* tb = reinterpret_cast<uint8_t*>(tile);
      for (int i = (int)(bytes & ~(size_t)3) + threadIdx.x; i < (int)bytes; i += 256) tb[i] = src[i];
    } else {
      uint8_t* tb = reinterpret_cast<uint8_t*>(tile);
      for (int i = threadIdx.x; i < (int)bytes; i += 256) tb[i] = src[i];
    }
    __syncthreads();
    if (threadIdx.x < n) {
      const uint8_t* p = reinterpret_cast<const uint8_t*>(tile) + 9 * threadIdx.x;
      const uint32_t b8 = p[8];
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t hi = (uint32_t)p[2 * i] | ((uint32_t)p[2 * i + 1] << 8);
        w[i] = ((hi << 16) | (((b8 >> (2 * i)) & 3u) << 14)) + 0x2000u;
      }
      out[g0 + threadIdx.x] = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
    }
  }
}

// four samples (three 32-bit words) per thread
__global__ void __launch_bounds__(256) widen_24bit_kernel(const uint32_t* __restrict__ in, int4* __restrict__ out, size_t groups)
{
  for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < groups; g += (size_t)gridDim.x * 256) {
    const uint32_t w0 = in[3 * g], w1 = in[3 * g + 1], w2 = in[3 * g + 2];
    const uint32_t o0 = w0 << 8;
    const uint32_t o1 = ((w0 >> 16) & 0xff00u) | (w1 << 16);
    const uint32_t o2 = ((w1 >> 8) & 0xffff00u) | (w2 << 24);
    const uint32_t o3 = w2 & 0xffffff00u;
    out[g] = make_int4((int)o0, (int)o1, (int)o2, (int)o3);
  }
}

// 8-bit unsigned PCM -> int16 (rxin.c:1573-1583): rxin_isho[j] = (rxin_char[j] << 8) - 32640,
// stored to a short (the reference's signed-char promotion and the unsigned form agree modulo
// 2^16).  Four samples per thread.
__global__ void __launch_bounds__(256) widen_8bit_kernel(const uint32_t* __restrict__ in, uint2* __restrict__ out, size_t groups)
{
  for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < groups; g += (size_t)gridDim.x * 256) {
    const uint32_t w = in[g];
    const uint32_t s0 = (((w & 0xffu) << 8) - 32640u) & 0xffffu;
    const uint32_t s1 = ((((w >> 8) & 0xffu) << 8) - 32640u) & 0xffffu;
    const uint32_t s2 = ((((w >> 16) & 0xffu) << 8) - 32640u) & 0xffffu;
    const uint32_t s3 = (((w >> 24) << 8) - 32640u) & 0xffffu;
    out[g] = make_uint2(s0 | (s1 << 16), s2 | (s3 << 16));
  }
}

// 32-bit float samples -> int32 (rxin.c:1624-1634): rxin_int[j] = 0x7fffffff * z[j], i.e. the float
// product with (float)0x7fffffff = 2^31 truncated toward zero; what does not fit (|product| >= 2^31,
// NaN) becomes 0x80000000, the "integer indefinite" the reference's cvttss2si returns on x86-64.
__global__ void __launch_bounds__(256) float_to_int32_kernel(const float4* __restrict__ in, int4* __restrict__ out, size_t groups)
{
  for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < groups; g += (size_t)gridDim.x * 256) {
    const float4 z = in[g];
    const float v[4] = {z.x, z.y, z.z, z.w};
    int r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float t = __fmul_rn(2147483648.0f, v[i]);
      r[i] = (t >= 2147483648.0f || t < -2147483648.0f || t != t) ? (int)0x80000000u : __float2int_rz(t);
    }
    out[g] = make_int4(r[0], r[1], r[2], r[3]);
  }
}

}  // namespace lb
