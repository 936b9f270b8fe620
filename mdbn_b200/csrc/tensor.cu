// Large-batch / many-chain path: TF32 GEMMs on the 5th-generation tensor cores.
//
//   tcgen05.mma.cta_group::1.kind::tf32  (one elected thread issues; SASS: UTCHMMA-family)
//   operands staged in shared memory by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle; UTMALDG)
//   accumulator 128 x 128 fp32 in tensor memory (TMEM), read back with tcgen05.ld (LDTM)
//   warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-9 =
//   epilogue (two per TMEM lane quarter, one column half each; MUFU sigmoid), 3-stage mbarrier ring, 2 CTAs per SM so one tile's epilogue overlaps another's mainloop.
//
// The three GEMMs of a CD step map onto ONE kernel template by operand major-ness:
//   propup    H[B,H]  = X[B,V] W[V,H]          A K-major  (X rows),  B MN-major (W rows are K)
//   propdown  V[B,V]  = Hs[B,H] W[V,H]^T       A K-major,            B K-major  (W rows are N)
//   stats     G[V,H]  = [v0;nv]^T [ph;nh] with the nv/nh half negated through the
//             instruction descriptor's a_negate bit     A MN-major,  B MN-major, K = 2B
// with the bias + sigmoid (or linear GRBM mean) + Bernoulli / Gaussian sampling fused into the
// TMEM epilogue, so pre-activations never go to HBM.  TF32 keeps 10 mantissa bits of W and of
// real-valued activations; {0,1} samples are exact.  Tolerance bar: 2e-3 relative.
#include <cuda.h>
#include "ctx.h"

namespace mdbn {

int apply_update(mdbn_ctx* c, const mdbn_cd_args& a, const float* G, int rows, cudaStream_t st);   // generic.cu

namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;     // BK floats = 128 bytes = one swizzle row
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NTHREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter: column halves)
constexpr int TMEM_COLS = 128;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

enum { EPI_ACT = 0, EPI_PART = 1 };
enum { ACT_SIGMOID = 0, ACT_LINEAR = 1 };
enum { SMP_NONE = 0, SMP_BERNOULLI = 1, SMP_MEAN = 2, SMP_GAUSS = 3 };

struct EpiParams {
  const float* bias;
  int act, smp;
  RngSeg rs;
  float *pre, *mean, *sample;
  long long ld_pre, ld_mean, ld_sample;
  float* part;          // EPI_PART: [splits][M][N]
  int vec4;             // all epilogue pointers 16-byte aligned, strides and N multiples of 4
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// sigmoid on the MUFU pipe (ex2.approx + rcp.approx, a few ulp — far inside the 2e-3 bar of the TF32 path)
__device__ __forceinline__ float sigmoid_mufu(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + __expf(-x)));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// shared-memory matrix descriptor (SM100 UMMA).  K-major tiles use the plain 128-byte swizzle (16-byte
// chunks); MN-major TF32 operands only exist in the 128-byte swizzle with 32-byte atoms (4-row period).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);             // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;    // [16,30) leading byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;    // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                               // [46,48) descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;                          // [61,64) 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptor: D=f32, A=B=tf32, M=128, N=128
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn, bool a_neg) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_neg ? 1u : 0u) << 13) | ((a_mn ? 1u : 0u) << 15) |
         ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(NTHREADS) tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                                                           int kb_per_split, int kb_neg, EpiParams ep) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb_total = (K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(nkb_total, kb_begin + kb_per_split);
  const int nkb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES, it = i / STAGES;
      mbar_wait(empty0 + 8 * s, (it & 1) ^ 1);
      const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
      const uint32_t bar = full0 + 8 * s;
      mbar_expect_tx(bar, STAGE_BYTES);
      const int k0 = (kb_begin + i) * BK;
      if (A_MN) {
#pragma unroll
        for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * (BK * 128), &tmA, bar, m0 + 32 * c, k0);
      } else {
        tma_load_2d(sa, &tmA, bar, k0, m0);
      }
      if (B_MN) {
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) tma_load_2d(sb + c * (BK * 128), &tmB, bar, n0 + 32 * c, k0);
      } else {
        tma_load_2d(sb, &tmB, bar, k0, n0);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES, it = i / STAGES;
      mbar_wait(full0 + 8 * s, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
      const uint32_t idesc = make_idesc(A_MN, B_MN, (kb_begin + i) >= kb_neg);
#pragma unroll
      for (int kk = 0; kk < BK / 8; ++kk) {
        // K-major: 8 floats = 32 bytes along the swizzled row; MN-major: 8 K-rows = 1024 bytes
        const uint64_t ad = A_MN ? make_desc(sa + kk * 1024, BK * 128, 512, 1) : make_desc(sa + kk * 32, 16, 1024, 2);
        const uint64_t bd = B_MN ? make_desc(sb + kk * 1024, BK * 128, 512, 1) : make_desc(sb + kk * 32, 16, 1024, 2);
        umma_tf32(tmem_base, ad, bd, idesc, (i > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(empty0 + 8 * s);          // frees the smem stage when these MMAs have read it
    }
    umma_commit(tfull);                     // accumulator complete
  } else if (warp >= 2) {
    // ===== epilogue: TMEM -> registers -> bias / activation / sampling -> global =====
    mbar_wait(tfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quarter = warp & 3;           // a warp may only touch its own 32 TMEM lanes
    const int chalf = (warp - 2) >> 2;      // ... and the two warps of a quarter split the columns
    const int m = m0 + quarter * 32 + lane;
#pragma unroll 1
    for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c * 32, r);
      const int nb = n0 + c * 32;
      if (m < M && nkb > 0) {
        if (EPI == EPI_PART) {
          float* dst = ep.part + ((size_t)blockIdx.z * M + m) * N + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < N) dst[j] = __uint_as_float(r[j]);
        } else if (ep.vec4 && nb + 32 <= N) {
          // 4 columns per step: one Philox block feeds 4 draws, 16-byte stores
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = nb + j;
            const float4 bv = *reinterpret_cast<const float4*>(ep.bias + n);
            float pre[4] = {__uint_as_float(r[j]) + bv.x, __uint_as_float(r[j + 1]) + bv.y,
                            __uint_as_float(r[j + 2]) + bv.z, __uint_as_float(r[j + 3]) + bv.w};
            float mu[4], x[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) mu[t] = ep.act == ACT_SIGMOID ? sigmoid_mufu(pre[t]) : pre[t];
            if (ep.pre) *reinterpret_cast<float4*>(ep.pre + m * ep.ld_pre + n) = make_float4(pre[0], pre[1], pre[2], pre[3]);
            if (ep.mean) *reinterpret_cast<float4*>(ep.mean + m * ep.ld_mean + n) = make_float4(mu[0], mu[1], mu[2], mu[3]);
            if (ep.sample) {
              const long long e = (long long)m * N + n;      // multiple of 4 on this path
              if (ep.smp == SMP_BERNOULLI) {
                float u[4];
                if (ep.rs.mode == MDBN_RNG_BUFFER) {
                  const float4 uv = __ldg(reinterpret_cast<const float4*>(ep.rs.seg + e));
                  u[0] = uv.x; u[1] = uv.y; u[2] = uv.z; u[3] = uv.w;
                } else {
                  Philox4 ph = philox4x32_10((uint32_t)(e >> 2), ep.rs.c1, ep.rs.c2, ep.rs.c3, ep.rs.k0, ep.rs.k1);
                  u[0] = u24(ph.x); u[1] = u24(ph.y); u[2] = u24(ph.z); u[3] = u24(ph.w);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) x[t] = u[t] < mu[t] ? 1.f : 0.f;
              } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) x[t] = ep.smp == SMP_GAUSS ? mu[t] + rng_normal(ep.rs, e + t) : mu[t];
              }
              *reinterpret_cast<float4*>(ep.sample + m * ep.ld_sample + n) = make_float4(x[0], x[1], x[2], x[3]);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nb + j;
            if (n < N) {
              const float pre = __uint_as_float(r[j]) + ep.bias[n];
              const float mu = ep.act == ACT_SIGMOID ? sigmoid_mufu(pre) : pre;
              if (ep.pre) ep.pre[m * ep.ld_pre + n] = pre;
              if (ep.mean) ep.mean[m * ep.ld_mean + n] = mu;
              if (ep.sample) {
                const long long e = (long long)m * N + n;
                float x;
                if (ep.smp == SMP_BERNOULLI) x = rng_uniform(ep.rs, e) < mu ? 1.f : 0.f;
                else if (ep.smp == SMP_GAUSS) x = mu + rng_normal(ep.rs, e);
                else x = mu;
                ep.sample[m * ep.ld_sample + n] = x;
              }
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  }
  return fn;
}

// Encoding a tensor map costs the host 3-5 us, and a CD step issues 5 to 2k+3 GEMMs on the SAME operands step after step
// (parameters, context scratch): a small per-thread cache keyed by everything that enters the descriptor.
struct MapKey {
  const float* ptr;
  long long inner, outer, ld;
  int box_inner, box_outer, mn;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer && mn == o.mn;
  }
};
struct MapCache {
  static constexpr int N = 64;
  MapKey key[N];
  CUtensorMap map[N];
  int used = 0, next = 0;
};
static int make_map_uncached(CUtensorMap* tm, const float* ptr, long long inner, long long outer, long long ld,
                             int box_inner, int box_outer, bool mn_major);
// operand matrix in memory: [outer][inner] row-major with row stride ld (floats)
static int make_map(CUtensorMap* tm, const float* ptr, long long inner, long long outer, long long ld, int box_inner,
                    int box_outer, bool mn_major) {
  static thread_local MapCache cache;
  const MapKey kq{ptr, inner, outer, ld, box_inner, box_outer, mn_major ? 1 : 0};
  for (int i = 0; i < cache.used; ++i)
    if (cache.key[i] == kq) { *tm = cache.map[i]; return 0; }
  MDBN_TRY(make_map_uncached(tm, ptr, inner, outer, ld, box_inner, box_outer, mn_major));
  const int slot = cache.used < MapCache::N ? cache.used++ : (cache.next++ % MapCache::N);
  cache.key[slot] = kq;
  cache.map[slot] = *tm;
  return 0;
}
static int make_map_uncached(CUtensorMap* tm, const float* ptr, long long inner, long long outer, long long ld,
                             int box_inner, int box_outer, bool mn_major) {
  EncodeFn enc = get_encode();
  MDBN_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  MDBN_CHECK(((uintptr_t)ptr & 15) == 0 && (ld * 4) % 16 == 0, "TMA operand must be 16-byte aligned (ptr %p ld %lld)",
             (const void*)ptr, ld);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MDBN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

struct Operand {
  const float* ptr;
  long long ld;
  bool mn_major;      // true: memory is [K][MN]; false: memory is [MN][K]
};

template <bool A_MN, bool B_MN, int EPI>
static int launch_gemm(mdbn_ctx* c, const Operand& A, const Operand& Bo, int M, int N, int K, int splits, int kneg,
                       const EpiParams& ep, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  if (A_MN) MDBN_TRY(make_map(&tmA, A.ptr, M, K, A.ld, 32, BK, true)); else MDBN_TRY(make_map(&tmA, A.ptr, K, M, A.ld, BK, BM, false));
  if (B_MN) MDBN_TRY(make_map(&tmB, Bo.ptr, N, K, Bo.ld, 32, BK, true)); else MDBN_TRY(make_map(&tmB, Bo.ptr, K, N, Bo.ld, BK, BN, false));
  auto kfn = tc_gemm_kernel<A_MN, B_MN, EPI>;
  static bool configured[64] = {};
  if (!configured[c->device]) {
    MDBN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured[c->device] = true;
  }
  const int nkb = (K + BK - 1) / BK;
  int kbps = (nkb + splits - 1) / splits;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, (nkb + kbps - 1) / kbps);
  kfn<<<grid, NTHREADS, SMEM_BYTES, st>>>(tmA, tmB, M, N, K, kbps, kneg / BK, ep);
  c->launches++;
  MDBN_CUDA(cudaGetLastError());
  return 0;
}

// ---- small ld-aware helpers of the tensor path ----------------------------------
__global__ void gather_rows_ld_kernel(const float* __restrict__ data, long long ld, const int* __restrict__ idx, int B,
                                      int V, float* __restrict__ out, long long ldo, float* __restrict__ xi) {
  long long total = (long long)B * V;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int b = (int)(e / V), i = (int)(e % V);
    long long r = idx ? idx[b] : b;
    float x = data[r * ld + i];
    out[b * ldo + i] = x;
    if (xi) xi[b * ldo + i] = roundf(x);
  }
}
// the same, four columns per thread (V, strides and pointers multiples of 4 / 16 bytes): one row per blockIdx.y
__global__ void gather_rows_ld4_kernel(const float* __restrict__ data, long long ld, const int* __restrict__ idx, int B,
                                       int V4, float* __restrict__ out, long long ldo, float* __restrict__ xi) {
  const int b = blockIdx.y;
  const long long r = idx ? idx[b] : b;
  const float4* src = reinterpret_cast<const float4*>(data + r * ld);
  float4* dst = reinterpret_cast<float4*>(out + (size_t)b * ldo);
  float4* dx = xi ? reinterpret_cast<float4*>(xi + (size_t)b * ldo) : nullptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V4; i += gridDim.x * blockDim.x) {
    const float4 x = __ldg(src + i);
    dst[i] = x;
    if (dx) dx[i] = make_float4(roundf(x.x), roundf(x.y), roundf(x.z), roundf(x.w));
  }
}
// raw column sums of (top half - bottom half) of a [2B, N] matrix, two deterministic stages:
// row chunks in parallel, then a fixed-order sum of the chunk partials
__global__ void col_diff_partial_kernel(const float* __restrict__ X, long long ld, int B, int N, int rows_per_chunk,
                                        float* __restrict__ partial) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int b0 = blockIdx.y * rows_per_chunk, b1 = min(B, b0 + rows_per_chunk);
  float p = 0.f, q = 0.f;
  for (int b = b0; b < b1; ++b) { p += X[(size_t)b * ld + n]; q += X[(size_t)(B + b) * ld + n]; }
  partial[(size_t)blockIdx.y * N + n] = p - q;
}
__global__ void col_diff_final_kernel(const float* __restrict__ partial, int chunks, int N, float* __restrict__ out) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;   // one warp per column
  if (n >= N) return;
  float s = 0.f;
  for (int c = lane; c < chunks; c += 32) s += partial[(size_t)c * N + n];
  s = warp_sum(s);
  if (lane == 0) out[n] = s;
}
static int col_diff_sum(mdbn_ctx* c, const float* X, long long ld, int B, int N, float* out, cudaStream_t st) {
  const int rpc = 32, chunks = (B + rpc - 1) / rpc;
  float* part = (float*)ws_get(c, WS_MISC, (size_t)chunks * N * sizeof(float));
  if (!part) return 3;
  col_diff_partial_kernel<<<dim3((N + 255) / 256, chunks), 256, 0, st>>>(X, ld, B, N, rpc, part);
  col_diff_final_kernel<<<(N + 7) / 8, 256, 0, st>>>(part, chunks, N, out);
  c->launches += 2;
  return 0;
}
__global__ void recon_cost_ld_kernel(const float* __restrict__ prev, const float* __restrict__ v0, long long ld, int B,
                                     int V, int kind, float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  long long n = (long long)B * V;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int b = (int)(e / V), i = (int)(e % V);
    float p = prev[b * ld + i], t = v0[b * ld + i];
    if (kind == MDBN_GRBM) { float d = sigmoidf_(p) - t; s += d * d; }
    else s += t * softplusf_(-p) + (1.f - t) * softplusf_(p);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void pl_row_ld_kernel(const float* __restrict__ prex, long long ldy, int H, const float* __restrict__ xi,
                                 long long ldx, int V, const float* __restrict__ W, int ldw, const float* __restrict__ vb,
                                 const int* __restrict__ bit_idx, int kind, float* __restrict__ partial) {
  __shared__ float red[32];
  int b = blockIdx.x, idx = *bit_idx;
  float x = xi[(size_t)b * ldx + idx];
  float d = 1.f - 2.f * x;
  float h0 = 0.f, h1 = 0.f;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float p = prex[(size_t)b * ldy + j];
    h0 += softplusf_(p);
    h1 += softplusf_(p + d * W[(size_t)idx * ldw + j]);
  }
  h0 = block_sum(h0, red);
  h1 = block_sum(h1, red);
  if (threadIdx.x == 0) {
    float vterm;
    if (kind == MDBN_GRBM) { float a = x - vb[idx], c = (1.f - x) - vb[idx]; vterm = 0.5f * (a * a - c * c); }
    else vterm = d * vb[idx];
    partial[b] = -(float)V * softplusf_((h1 - h0) + vterm);
  }
}
__global__ void sum_tree_kernel(const float* __restrict__ partial, int n, float* __restrict__ out, float rows,
                                float* __restrict__ rows_out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) { *out = s; *rows_out = rows; }
}
__global__ void bump_bit_kernel(int* bit_idx, int V) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *bit_idx = (*bit_idx + 1) % V;
}
__global__ void reduce_parts_kernel(const float* __restrict__ part, int splits, long long n, float* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(size_t)z * n + e];
    out[e] = s;
  }
}
__global__ void copy_rows_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd,
                                 int B, int N) {
  long long total = (long long)B * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int b = (int)(e / N), j = (int)(e % N);
    dst[b * ldd + j] = src[b * lds + j];
  }
}

static int vec_ok(const EpiParams& ep, int N) {
  uintptr_t a = (uintptr_t)ep.bias | (uintptr_t)ep.pre | (uintptr_t)ep.mean | (uintptr_t)ep.sample |
                (ep.rs.mode == MDBN_RNG_BUFFER ? (uintptr_t)ep.rs.seg : 0);
  return N % 4 == 0 && ep.ld_pre % 4 == 0 && (a & 15) == 0;
}
static int up(mdbn_ctx* c, const float* W, int ldw, const float* hb, int B, int V, int H, const float* x, long long ldx,
              float* pre, float* mean, float* sample, long long ldo, const RngSeg& rs, cudaStream_t st) {
  EpiParams ep{};
  ep.bias = hb; ep.act = ACT_SIGMOID; ep.smp = sample ? SMP_BERNOULLI : SMP_NONE; ep.rs = rs;
  ep.pre = pre; ep.mean = mean; ep.sample = sample; ep.ld_pre = ep.ld_mean = ep.ld_sample = ldo;
  ep.vec4 = vec_ok(ep, H);
  return launch_gemm<false, true, EPI_ACT>(c, Operand{x, ldx, false}, Operand{W, ldw, true}, B, H, V, 1, 1 << 30, ep, st);
}
static int down(mdbn_ctx* c, const float* W, int ldw, const float* vb, int B, int V, int H, int kind, int noisy,
                const float* h, long long ldh, float* pre, float* mean, float* sample, long long ldo, const RngSeg& rs,
                cudaStream_t st) {
  EpiParams ep{};
  ep.bias = vb; ep.act = kind == MDBN_GRBM ? ACT_LINEAR : ACT_SIGMOID;
  ep.smp = !sample ? SMP_NONE : (kind == MDBN_GRBM ? (noisy ? SMP_GAUSS : SMP_MEAN) : SMP_BERNOULLI);
  ep.rs = rs;
  ep.pre = pre; ep.mean = mean; ep.sample = sample; ep.ld_pre = ep.ld_mean = ep.ld_sample = ldo;
  ep.vec4 = vec_ok(ep, V);
  return launch_gemm<false, false, EPI_ACT>(c, Operand{h, ldh, false}, Operand{W, ldw, false}, B, V, H, 1, 1 << 30, ep,
                                            st);
}

}  // namespace tc

bool tensor_phase_supported(const void* W, int ldw, const void* x, long long ldx) {
  return ldw % 4 == 0 && ldx % 4 == 0 && (((uintptr_t)W | (uintptr_t)x) & 15) == 0 && tc::get_encode() != nullptr;
}
int tensor_propup(mdbn_ctx* c, const float* W, int ldw, const float* hb, const float* v, int ldv, int B, int V, int H,
                  float* pre, float* mean, float* sample, const RngSeg& rs, cudaStream_t st) {
  return tc::up(c, W, ldw, hb, B, V, H, v, ldv, pre, mean, sample, H, rs, st);
}
int tensor_propdown(mdbn_ctx* c, const float* W, int ldw, const float* vb, const float* h, int ldh, int B, int V, int H,
                    int kind, int noisy, float* pre, float* mean, float* sample, const RngSeg& rs, cudaStream_t st) {
  return tc::down(c, W, ldw, vb, B, V, H, kind, noisy, h, ldh, pre, mean, sample, V, rs, st);
}

bool tensor_supported(const mdbn_ctx*, const mdbn_cd_args& a) {
  if (a.B < 32 || a.B % tc::BK != 0) return false;          // the nv/nh half is negated per 32-row K block
  if (a.ldw % 4 != 0 || ((uintptr_t)a.W & 15)) return false;
  if (a.phase == MDBN_PHASE_APPLY) return false;
  return tc::get_encode() != nullptr;
}

int tensor_cd_step(mdbn_ctx* c, const mdbn_cd_args& a, cudaStream_t st) {
  using namespace tc;
  const int B = a.B, V = a.V, H = a.H, k = a.k;
  const long long VH = (long long)V * H;
  const long long ldx = (V + 3) & ~3, ldy = (H + 3) & ~3;
  float* G = a.phase == MDBN_PHASE_FULL ? (float*)ws_get(c, WS_G, (size_t)(VH + H + V + 2) * sizeof(float))
                                        : a.stats_buf;
  MDBN_CHECK(G != nullptr, "cd_step: stats buffer missing");
  float* XV = (float*)ws_get(c, WS_XV, (size_t)2 * B * ldx * sizeof(float));
  float* YH = (float*)ws_get(c, WS_YH, (size_t)2 * B * ldy * sizeof(float));
  // PCD chain state: the caller's [B, H] array IS the chain buffer when its row stride suits TMA (H % 4 == 0, 16-byte
  // aligned): the Gibbs steps read and overwrite it in place, no copy in, no copy out (src/rbm.py:308-311, :369)
  const bool chain_in_place = a.persistent != nullptr && ldy == H && (((uintptr_t)a.persistent) & 15) == 0;
  float* HS = chain_in_place ? a.persistent : (float*)ws_get(c, WS_HS, (size_t)B * ldy * sizeof(float));
  float* VS = (float*)ws_get(c, WS_VS, (size_t)B * ldx * sizeof(float));
  float* PREV = (float*)ws_get(c, WS_PREV, (size_t)B * ldx * sizeof(float));
  float* RED = (float*)ws_get(c, WS_RED, (size_t)(B > 4096 ? B : 4096) * sizeof(float));
  if (!XV || !YH || !HS || !VS || !PREV || !RED) return 3;
  const bool pcd = a.persistent != nullptr;
  float *XI = nullptr, *PREX = nullptr;
  if (pcd) {
    XI = (float*)ws_get(c, WS_XI, (size_t)B * ldx * sizeof(float));
    PREX = (float*)ws_get(c, WS_PREX, (size_t)B * ldy * sizeof(float));
    if (!XI || !PREX) return 3;
  }
  ULayout ul = u_layout(a.kind, a.noisy, B, V, H);
  const int eb = 4 * c->num_sms;
  if (V % 4 == 0 && a.ld_data % 4 == 0 && (((uintptr_t)a.data) & 15) == 0)
    gather_rows_ld4_kernel<<<dim3((V / 4 + 255) / 256, B), 256, 0, st>>>(a.data, a.ld_data, a.indices, B, V / 4, XV, ldx, XI);
  else
    gather_rows_ld_kernel<<<eb, 256, 0, st>>>(a.data, a.ld_data, a.indices, B, V, XV, ldx, XI);
  c->launches++;
  // positive phase
  MDBN_TRY(up(c, a.W, a.ldw, a.hbias, B, V, H, XV, ldx, nullptr, YH, pcd ? nullptr : HS, ldy,
              make_seg(a.rng, ul.off_h0, 0), st));
  if (pcd) {
    MDBN_TRY(up(c, a.W, a.ldw, a.hbias, B, V, H, XI, ldx, PREX, nullptr, nullptr, ldy, make_seg(a.rng, 0, 0), st));
    if (!chain_in_place) {
      copy_rows_kernel<<<eb, 256, 0, st>>>(a.persistent, H, HS, ldy, B, H);      // chain state, padded stride for TMA
      c->launches++;
    }
  }
  float* nv_mean = XV + (size_t)B * ldx;
  float* nh_mean = YH + (size_t)B * ldy;
  for (int s = 0; s < k; ++s) {
    long long base = (long long)B * H + s * ul.step_stride;
    MDBN_TRY(down(c, a.W, a.ldw, a.vbias, B, V, H, a.kind, a.noisy, HS, ldy, PREV, nv_mean,
                  a.kind == MDBN_RBM ? VS : nullptr, ldx, make_seg(a.rng, base + ul.off_v, ord_v(s)), st));
    const float* v_in = a.kind == MDBN_GRBM ? nv_mean : VS;
    MDBN_TRY(up(c, a.W, a.ldw, a.hbias, B, V, H, v_in, ldx, nullptr, nh_mean, HS, ldy,
                make_seg(a.rng, base + ul.off_h, ord_h(s)), st));
  }
  // statistics: G = [v0;nv]^T (+/-) [ph;nh] over K = 2B, split-K across CTAs when V*H has few tiles
  {
    const int tiles = ((V + BM - 1) / BM) * ((H + BN - 1) / BN);
    const int nkb = 2 * B / BK;
    int splits = (2 * c->num_sms + tiles - 1) / tiles;
    if (splits > nkb / 2) splits = nkb / 2 > 0 ? nkb / 2 : 1;
    if (splits > 32) splits = 32;
    if (splits < 1) splits = 1;
    const int kbps = (nkb + splits - 1) / splits;
    const int zs = (nkb + kbps - 1) / kbps;
    float* part = (float*)ws_get(c, WS_PART, (size_t)zs * VH * sizeof(float));
    if (!part) return 3;
    EpiParams ep{};
    ep.part = part;
    MDBN_TRY((launch_gemm<true, true, EPI_PART>(c, Operand{XV, ldx, true}, Operand{YH, ldy, true}, V, H, 2 * B, splits, B,
                                                ep, st)));
    reduce_parts_kernel<<<eb, 256, 0, st>>>(part, zs, VH, G);
    c->launches++;
    if (c->ev_stats_w) { MDBN_CUDA(cudaEventRecord(c->ev_stats_w, st)); c->ev_stats_w_done = true; }
  }
  MDBN_TRY(col_diff_sum(c, YH, ldy, B, H, G + VH, st));
  MDBN_TRY(col_diff_sum(c, XV, ldx, B, V, G + VH + H, st));
  if (pcd) {
    pl_row_ld_kernel<<<B, 128, 0, st>>>(PREX, ldy, H, XI, ldx, V, a.W, a.ldw, a.vbias, a.bit_i_idx, a.kind, RED);
    c->launches++;
    sum_tree_kernel<<<1, 256, 0, st>>>(RED, B, G + VH + H + V, (float)B, G + VH + H + V + 1);
    c->launches++;
    bump_bit_kernel<<<1, 32, 0, st>>>(a.bit_i_idx, V);
    c->launches++;
    if (!chain_in_place) {
      copy_rows_kernel<<<eb, 256, 0, st>>>(HS, ldy, a.persistent, H, B, H);
      c->launches++;
    }
  } else {
    const int nb = 256;
    recon_cost_ld_kernel<<<nb, 256, 0, st>>>(PREV, XV, ldx, B, V, a.kind, RED);
    c->launches++;
    sum_tree_kernel<<<1, 256, 0, st>>>(RED, nb, G + VH + H + V, (float)B, G + VH + H + V + 1);
    c->launches++;
  }
  MDBN_CUDA(cudaGetLastError());
  if (a.phase == MDBN_PHASE_STATS) return 0;
  return apply_update(c, a, G, B, st);
}

}  // namespace mdbn
