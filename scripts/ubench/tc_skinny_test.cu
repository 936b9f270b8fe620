// Standalone validation of the tcgen05 building blocks for the skinny (batch <= 16) kernel:
//   U: out[j, b] = sum_i W[i][j] * v[b][i]    A = W^T (MN-major, SW128 / 32B atoms, TMA), B = v (K-major, no swizzle, SIMT-written)
//   D: out[i, b] = sum_j W[i][j] * h[b][j]    A = W   (K-major, SW128, TMA),                B = h (K-major, no swizzle)
// with the fp32 operands split hi + lo (hi = what the tensor core keeps after truncation, lo = exact
// remainder in a second shared-memory tile) so the result is fp32-accurate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_skinny_test tc_skinny_test.cu ; run on a B200.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int R = 128, H = 400, HP = 512, NB = 16, BREAL = 10;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// idesc: D f32, A/B tf32, M=128, N=16
__device__ constexpr uint32_t make_idesc(bool a_mn, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// K-major, no swizzle ("interleave") B panel: element (n, k) of a [16][K] matrix
__device__ __host__ inline int bpanel_off(int n, int k, int K) { return (n >> 3) * (K / 4) * 32 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3); }   // in floats

// mode 0: U (A = W^T, MN-major), mode 1: D (A = W, K-major)
__global__ void __launch_bounds__(192) test_kernel(const __grid_constant__ CUtensorMap tmMN, const __grid_constant__ CUtensorMap tmK,
                                                   const float* __restrict__ vpan_g, const float* __restrict__ hpan_g, float* __restrict__ outU,
                                                   float* __restrict__ outD) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* st_hi = reinterpret_cast<float*>(smem);              // 16 KB stage
  float* st_lo = reinterpret_cast<float*>(smem + 16384);      // 16 KB lo twin
  float* vp_hi = reinterpret_cast<float*>(smem + 32768);      // [16][128] 8 KB
  float* vp_lo = reinterpret_cast<float*>(smem + 32768 + 8192);
  float* hp_hi = reinterpret_cast<float*>(smem + 49152);      // [16][HP=512] 32 KB (binary: no lo)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t bar_full = smem_u32(bars), bar_mma = smem_u32(bars + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(bar_full, 1); mbar_init(bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // operand panels: hi = raw value, lo = value - trunc_tf32(value)
  for (int e = tid; e < 16 * 128; e += 192) {
    float x = vpan_g[e];
    float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    vp_hi[e] = x;
    vp_lo[e] = x - hi;
  }
  for (int e = tid; e < 16 * HP; e += 192) hp_hi[e] = hpan_g[e];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  uint32_t ph_full = 0, ph_mma = 0;

  // ================= U: 4 M-tiles (128 j each) x 4 row quarters (32 rows each) =================
  for (int mt = 0; mt < 4; ++mt) {
    for (int rq = 0; rq < 4; ++rq) {
      if (tid == 0) {
        mbar_expect_tx(bar_full, 16384);
        for (int c = 0; c < 4; ++c) tma_load_2d(smem_u32(st_hi) + c * 4096, &tmMN, bar_full, mt * 128 + 32 * c, rq * 32);
      }
      mbar_wait(bar_full, ph_full); ph_full ^= 1;
      for (int e = tid; e < 4096; e += 192) {
        float x = st_hi[e];
        st_lo[e] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc(true, NB);
        for (int kk = 0; kk < 4; ++kk) {
          const int k = rq * 32 + kk * 8;    // row index inside the 128-row slab = K index of the B panel
          const uint64_t a_hi = make_desc(smem_u32(st_hi) + kk * 1024, 4096, 512, 1);
          const uint64_t a_lo = make_desc(smem_u32(st_lo) + kk * 1024, 4096, 512, 1);
          // B panel [16][128]: 8-row groups are (128/4)*128 B = 4096 B apart, core matrices 128 B apart along K
          const uint64_t b_hi = make_desc(smem_u32(vp_hi) + (k / 4) * 128, 128, 4096, 0);
          const uint64_t b_lo = make_desc(smem_u32(vp_lo) + (k / 4) * 128, 128, 4096, 0);
          const uint32_t first = (rq == 0 && kk == 0) ? 0u : 1u;
          umma_tf32(tmem + mt * NB, a_hi, b_hi, idesc, first);
          umma_tf32(tmem + mt * NB, a_lo, b_hi, idesc, 1u);
          umma_tf32(tmem + mt * NB, a_hi, b_lo, idesc, 1u);
        }
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;      // everybody waits: the stage buffer is reused
      __syncthreads();
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp >= 2) {
    const int quarter = warp & 3;
    for (int mt = 0; mt < 4; ++mt) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + mt * NB, r);
      const int j = mt * 128 + quarter * 32 + lane;
      for (int b = 0; b < 16; ++b) outU[j * 16 + b] = __uint_as_float(r[b]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();

  // ================= D: 13 column chunks (32 j each) of the 128-row slab =================
  for (int c = 0; c < 13; ++c) {
    if (tid == 0) {
      mbar_expect_tx(bar_full, 16384);
      tma_load_2d(smem_u32(st_hi), &tmK, bar_full, c * 32, 0);
    }
    mbar_wait(bar_full, ph_full); ph_full ^= 1;
    for (int e = tid; e < 4096; e += 192) {
      float x = st_hi[e];
      st_lo[e] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 32) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = make_idesc(false, NB);
      for (int kk = 0; kk < 4; ++kk) {
        const int k = c * 32 + kk * 8;      // column index = K index of the h panel [16][HP]
        const uint64_t a_hi = make_desc(smem_u32(st_hi) + kk * 32, 16, 1024, 2);
        const uint64_t a_lo = make_desc(smem_u32(st_lo) + kk * 32, 16, 1024, 2);
        const uint64_t b_hi = make_desc(smem_u32(hp_hi) + (k / 4) * 128, 128, (HP / 4) * 128, 0);
        const uint32_t first = (c == 0 && kk == 0) ? 0u : 1u;
        umma_tf32(tmem + 64, a_hi, b_hi, idesc, first);
        umma_tf32(tmem + 64, a_lo, b_hi, idesc, 1u);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp >= 2) {
    const int quarter = warp & 3;
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + 64, r);
    const int i = quarter * 32 + lane;
    for (int b = 0; b < 16; ++b) outD[i * 16 + b] = __uint_as_float(r[b]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  std::vector<float> W(R * H), vpan(16 * 128, 0.f), hpan(16 * HP, 0.f), v(16 * R, 0.f), h(16 * H, 0.f);
  srand(1);
  for (auto& x : W) x = (rand() / (float)RAND_MAX - 0.5f) * 0.4f;
  for (int b = 0; b < BREAL; ++b) for (int i = 0; i < R; ++i) { v[b * R + i] = (rand() / (float)RAND_MAX - 0.5f) * 3.f; vpan[bpanel_off(b, i, 128)] = v[b * R + i]; }
  for (int b = 0; b < BREAL; ++b) for (int j = 0; j < H; ++j) { h[b * H + j] = (rand() & 1) ? 1.f : 0.f; hpan[bpanel_off(b, j, HP)] = h[b * H + j]; }
  float *dW, *dv, *dh, *dU, *dD;
  CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dv, vpan.size() * 4)); CK(cudaMalloc(&dh, hpan.size() * 4));
  CK(cudaMalloc(&dU, 512 * 16 * 4)); CK(cudaMalloc(&dD, 128 * 16 * 4));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, vpan.data(), vpan.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dh, hpan.data(), hpan.size() * 4, cudaMemcpyHostToDevice));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  CUtensorMap tmMN, tmK;
  cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)H * 4}; cuuint32_t es[2] = {1, 1};
  cuuint32_t boxMN[2] = {32, 32}, boxK[2] = {32, 128};
  CUresult r1 = enc(&tmMN, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dW, dims, strides, boxMN, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dW, dims, strides, boxK, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", r1, r2); return 1; }
  const int smem_bytes = 49152 + 32768 + 256 + 1024;
  CK(cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  test_kernel<<<1, 192, smem_bytes>>>(tmMN, tmK, dv, dh, dU, dD);
  CK(cudaDeviceSynchronize());
  std::vector<float> oU(512 * 16), oD(128 * 16);
  CK(cudaMemcpy(oU.data(), dU, oU.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(oD.data(), dD, oD.size() * 4, cudaMemcpyDeviceToHost));
  double maxeU = 0, maxsU = 0, maxeD = 0, maxsD = 0;
  for (int j = 0; j < H; ++j) for (int b = 0; b < BREAL; ++b) {
    double s = 0, sa = 0; for (int i = 0; i < R; ++i) { s += (double)W[i * H + j] * v[b * R + i]; sa += fabs((double)W[i * H + j] * v[b * R + i]); }
    maxeU = fmax(maxeU, fabs(oU[j * 16 + b] - s)); maxsU = fmax(maxsU, sa);
  }
  for (int i = 0; i < R; ++i) for (int b = 0; b < BREAL; ++b) {
    double s = 0, sa = 0; for (int j = 0; j < H; ++j) { s += (double)W[i * H + j] * h[b * H + j]; sa += fabs((double)W[i * H + j] * h[b * H + j]); }
    maxeD = fmax(maxeD, fabs(oD[i * 16 + b] - s)); maxsD = fmax(maxsD, sa);
  }
  printf("U: max abs err %.3e  (sum|terms| %.3e)  rel %.3e   sample %f\n", maxeU, maxsU, maxeU / maxsU, oU[5 * 16 + 3]);
  printf("D: max abs err %.3e  (sum|terms| %.3e)  rel %.3e   sample %f\n", maxeD, maxsD, maxeD / maxsD, oD[7 * 16 + 2]);
  return 0;
}
