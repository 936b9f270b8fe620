// Shared device helpers for the mdbn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mdbn_b200.h"

namespace mdbn {

// ---- math (full-precision: the fp32 parity bar is 1e-5 relative) -------------
__device__ __forceinline__ float sigmoidf_(float x) {
  // Theano nnet.sigmoid saturates to exactly 0 / 1 at the extremes; so does this.
  return 1.0f / (1.0f + expf(-x));
}
__device__ __forceinline__ float softplusf_(float x) {
  return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}

// ---- Philox4x32-10 ---------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
  uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c[0], c[1], c[2], c[3]};
}
// 24-bit uniform strictly inside (0,1), exactly representable in fp32
__host__ __device__ __forceinline__ float u24(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// Randomness of one segment (one sampling site of one step).
struct RngSeg {
  int mode;              // MDBN_RNG_*
  const float* seg;      // BUFFER: first value of this segment
  uint32_t k0, k1;       // PHILOX key
  uint32_t c1, c2, c3;   // PHILOX: segment ordinal, offset lo, offset hi
};
__host__ inline RngSeg make_seg(const mdbn_rng& r, long long buf_off, uint32_t ordinal) {
  RngSeg s;
  s.mode = r.mode;
  s.seg = (r.mode == MDBN_RNG_BUFFER && r.buffer) ? r.buffer + buf_off : nullptr;
  s.k0 = (uint32_t)r.seed; s.k1 = (uint32_t)(r.seed >> 32);
  s.c1 = ordinal; s.c2 = (uint32_t)r.offset; s.c3 = (uint32_t)(r.offset >> 32);
  return s;
}
__device__ __forceinline__ float rng_uniform(const RngSeg& s, long long e) {
  if (s.mode == MDBN_RNG_BUFFER) return __ldg(s.seg + e);
  Philox4 p = philox4x32_10((uint32_t)(e >> 2), s.c1, s.c2, s.c3, s.k0, s.k1);
  uint32_t lane = (uint32_t)e & 3u;
  uint32_t x = lane == 0 ? p.x : lane == 1 ? p.y : lane == 2 ? p.z : p.w;
  return u24(x);
}
__device__ __forceinline__ float rng_normal(const RngSeg& s, long long e) {
  if (s.mode == MDBN_RNG_BUFFER) return __ldg(s.seg + e);
  // Box-Muller on two uniforms of this element's own Philox block (c0 = e, upper ordinal bit set)
  Philox4 p = philox4x32_10((uint32_t)e, s.c1 | 0x80000000u, s.c2, s.c3 ^ (uint32_t)(e >> 32), s.k0, s.k1);
  float u1 = u24(p.x), u2 = u24(p.y);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ---- reductions --------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// deterministic block sum (fixed tree); result valid in thread 0. `red` >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    r = lane < nw ? red[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace mdbn
