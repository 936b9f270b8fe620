// Batch > 20 path (mid batch, large batch, many chains): the GEMMs of a CD step on the 5th-generation tensor cores.
//
//   tcgen05.mma.cta_group::1.kind::tf32  (one elected thread issues; SASS: UTCHMMA-family)
//   operands staged in shared memory by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle; UTMALDG)
//   accumulator 128 x 128 fp32 in tensor memory (TMEM), read back with tcgen05.ld (LDTM)
//   warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-9 = epilogue (two per TMEM
//   lane quarter, one column half each; MUFU sigmoid), 3-stage mbarrier ring.
//
// The three GEMMs of a CD step map onto ONE kernel template by operand major-ness:
//   propup    H[B,H]  = X[B,V] W[V,H]          A K-major  (X rows),  B MN-major (W rows are K)
//   propdown  V[B,V]  = Hs[B,H] W[V,H]^T       A K-major,            B K-major  (W rows are N)
//   stats     G[V,H]  = [v0;nv]^T [ph;nh] with the nv/nh half negated through the
//             instruction descriptor's a_negate bit     A MN-major,  B MN-major, K = 2B
// with the bias + sigmoid (or linear GRBM mean) + Bernoulli / Gaussian sampling fused into the TMEM epilogue of the
// first two (pre-activations never go to HBM) and the lambda_1 / lambda_2 / weight-cost / momentum update of W and
// W_speed fused into the epilogue of the third (EPI_UPDATE: the statistics never go to HBM either, src/rbm.py:347-365).
//
// Two arithmetic modes (template parameter SPLIT):
//   SPLIT = false  plain TF32: the tensor core truncates the fp32 operands to 10 mantissa bits.  Bar 2e-3, {0,1}
//                  samples exact.  96 KB of stages -> 2 CTAs per SM (one tile's epilogue overlaps another's mainloop).
//   SPLIT = true   fp32-exact ("3xTF32"): every real-valued operand tile gets a lo twin  x - trunc_tf32(x)  written
//                  next to it in shared memory by the eight epilogue warps (idle during the mainloop) and each k-step
//                  issues A*B + A_lo*B + A*B_lo into the same accumulator; {0,1} operands need no twin.  Error
//                  ~2^-22 relative to sum|terms| -> the 1e-5 bar of the fp32 paths.  This is the DEFAULT for B > 20
//                  (the W-streaming SIMT kernel's register budget ends at B = 20); tf32 = 1 selects the plain mode.
// Skinny shapes (B <= 128 rows against 10^4 visible units) split K over the SMs: partial tiles, then one kernel
// that sums them in fixed order and applies the same fused epilogue.
#include <cuda.h>
#include "ctx.h"

namespace mdbn {

int apply_update(mdbn_ctx* c, const mdbn_cd_args& a, const float* G, int rows, cudaStream_t st);   // generic.cu
int free_energy_from_parts(mdbn_ctx* c, const float* part, int splits, int B, int H, const float* hb, const float* v,
                           long long ldv, int V, const float* vb, int kind, float* F, cudaStream_t st);   // generic.cu

namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;     // BK floats = 128 bytes = one swizzle row
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
constexpr int LO_OFF = A_BYTES + B_BYTES;      // SPLIT: the lo twins of A and B sit behind the pair
__host__ __device__ constexpr int stage_bytes(bool split) { return (A_BYTES + B_BYTES) * (split ? 2 : 1); }
// Thread layout.  Plain mode: warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter: column halves);
// 2 CTAs per SM, one tile per CTA.  SPLIT mode: four more warps (2-5) build the lo twins, warps 6-13 are the epilogue;
// one PERSISTENT CTA per SM walks the tiles with the accumulator double-buffered in TMEM, so the epilogue of a tile
// (global loads and stores of the update) overlaps the mainloop of the next one.
__host__ __device__ constexpr int n_threads(bool split) { return split ? 448 : 320; }
__host__ __device__ constexpr int epi_warp0(bool split) { return split ? 6 : 2; }
constexpr int N_XFORM = 128;                   // transform threads (SPLIT)
constexpr int TLD = 20;                        // row stride (floats) of an epilogue warp's 32 x 16 parking tile
constexpr int T_BYTES = 8 * 32 * TLD * 4;      // eight epilogue warps
__host__ __device__ constexpr int tmem_cols(bool split) { return split ? 256 : 128; }
__host__ __device__ constexpr int smem_bytes(bool split) {
  // the plain mode parks the epilogue tiles in the (by then idle) stages; the persistent mode needs its own room
  return STAGES * stage_bytes(split) + 1024 /*align*/ + 256 /*barriers*/ + (split ? T_BYTES : 0);
}

enum { EPI_ACT = 0, EPI_PART = 1, EPI_UPDATE = 2 };
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
  // EPI_UPDATE: the accumulator tile is v0^T ph - nv^T nh of rows m (visible) x columns n (hidden)
  float *uW, *uS;
  const float* uSnap;
  int uldw;
  UpdateScalars u;
  // ... with one extra row (m == uV: the visible "unit" that is always 1 -> sum_b ph - nh, the hidden-bias gradient) and
  // one extra column (n == uH: the hidden unit that is always 1 -> sum_b v0 - nv, the visible-bias gradient)
  int uV, uH;
  float *uhb, *uShb, *uvb, *uSvb;
  float inv_rows;
};
// bias + speed from the raw gradient sum d (src/rbm.py:416-417, :361-364)
__device__ __forceinline__ void update_bias_one(const EpiParams& ep, float* b, float* S, float d) {
  const float g = d * ep.inv_rows, s = *S;
  *S = g + (s - g) * ep.u.mom;
  *b = *b + s * ep.u.lr;
}

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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptor: D=f32, A=B=tf32, M=128, N=128
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn, bool a_neg) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_neg ? 1u : 0u) << 13) | ((a_mn ? 1u : 0u) << 15) |
         ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---- the fused epilogue of propup / propdown for 4 consecutive columns (n % 4 == 0, all pointers 16-byte aligned)
//      and for one element; shared by the TMEM epilogue and by the split-K reduction kernel ----
__device__ __forceinline__ void epi_quad(const EpiParams& ep, int m, int n, int N, const float (&acc)[4]) {
  const float4 bv = *reinterpret_cast<const float4*>(ep.bias + n);
  const float pre[4] = {acc[0] + bv.x, acc[1] + bv.y, acc[2] + bv.z, acc[3] + bv.w};
  float mu[4], x[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) mu[t] = ep.act == ACT_SIGMOID ? sigmoid_mufu(pre[t]) : pre[t];
  if (ep.pre) *reinterpret_cast<float4*>(ep.pre + m * ep.ld_pre + n) = make_float4(pre[0], pre[1], pre[2], pre[3]);
  if (ep.mean) *reinterpret_cast<float4*>(ep.mean + m * ep.ld_mean + n) = make_float4(mu[0], mu[1], mu[2], mu[3]);
  if (ep.sample) {
    const long long e = (long long)m * N + n;      // multiple of 4 on this path: one Philox block feeds 4 draws
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
__device__ __forceinline__ void epi_one(const EpiParams& ep, int m, int n, int N, float acc) {
  const float pre = acc + ep.bias[n];
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
// lo_mask (SPLIT only): bit 0 = A has a lo twin (real-valued operand), bit 1 = B has one.
// Work items = (n tile, m tile, K slice), n fastest.  Plain mode: one item per CTA (3-D grid).  SPLIT mode: a 1-D grid
// of persistent CTAs, item = blockIdx.x + i * gridDim.x; the smem ring and its mbarrier phases run on across items.
template <bool A_MN, bool B_MN, int EPI, bool SPLIT>
__global__ void __launch_bounds__(n_threads(SPLIT), SPLIT ? 1 : 2)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                   int kb_per_split, int kb_neg, int lo_mask, int mn3d, int nbx, int nby, int nbz, EpiParams ep) {
  constexpr int STAGE_BYTES = stage_bytes(SPLIT);
  constexpr int EW0 = epi_warp0(SPLIT);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), ready0 = smem_u32(bars + 2 * STAGES),
                 tfull0 = smem_u32(bars + 3 * STAGES), tempty0 = smem_u32(bars + 3 * STAGES + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb_total = (K + BK - 1) / BK;
  const bool twins = SPLIT && lo_mask != 0;
  const int n_items = nbx * nby * nbz;
  const int item0 = SPLIT ? (int)blockIdx.x : (int)(blockIdx.x + nbx * (blockIdx.y + nby * blockIdx.z));
  const int item_step = SPLIT ? (int)gridDim.x : n_items;
  auto decode = [&](int item, int& m0, int& n0, int& bz, int& kb_begin, int& nkb) {
    const int bx = item % nbx, r = item / nbx, by = r % nby;
    bz = r / nby;
    m0 = by * BM;
    n0 = bx * BN;
    kb_begin = bz * kb_per_split;
    nkb = min(nkb_total, kb_begin + kb_per_split) - kb_begin;
  };

  // The W / W_speed (/ W_snap) tile an update item will read-modify-write: pulled into L2 while its mainloop runs, one row
  // per thread of a group of `nthr` otherwise idle threads (NOT the producer warp: the bulk-prefetch instruction runs on
  // the uniform datapath, lane after lane, and 256 of them in front of an item's TMA loads cost it 6-9 us)
  auto prefetch_tile = [&](int m0, int n0, int t0, int nthr) {
    const int ncols = min(BN, ep.uldw - n0);
    if (ncols <= 0) return;
    for (int r = t0; r < BM && m0 + r < ep.uV; r += nthr) {
      const size_t o = (size_t)(m0 + r) * ep.uldw + n0;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ep.uW + o), "r"(ncols * 4) : "memory");
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ep.uS + o), "r"(ncols * 4) : "memory");
      if (ep.uSnap) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ep.uSnap + o), "r"(ncols * 4) : "memory");
    }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
      mbar_init(ready0 + 8 * s, N_XFORM / 32);      // one arrival per transform warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 8);                // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols(SPLIT))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (lane 0) =====
    int g = 0;      // k-blocks issued so far: ring position and phase
    for (int item = item0; item < n_items; item += item_step) {
      int m0, n0, bz, kb_begin, nkb;
      decode(item, m0, n0, bz, kb_begin, nkb);
      if (lane == 0) {
        for (int i = 0; i < nkb; ++i, ++g) {
          const int s = g % STAGES, it = g / STAGES;
          mbar_wait(empty0 + 8 * s, (it & 1) ^ 1);
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
          const uint32_t bar = full0 + 8 * s;
          mbar_expect_tx(bar, A_BYTES + B_BYTES);
          const int k0 = (kb_begin + i) * BK;
          if (A_MN) {
            // MN-major tile = BM/32 chunks of {32 columns x BK rows}: one 3-D box when the operand is library scratch
            // (view [chunk][k][32]; a ragged last chunk wraps into the next row: rows m >= M, discarded), else a box
            // per chunk
            if (mn3d & 1) {
              tma_load_3d(sa, &tmA, bar, 0, k0, m0 >> 5);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * (BK * 128), &tmA, bar, m0 + 32 * c, k0);
            }
          } else {
            tma_load_2d(sa, &tmA, bar, k0, m0);
          }
          if (B_MN) {
            if (mn3d & 2) {
              tma_load_3d(sb, &tmB, bar, 0, k0, n0 >> 5);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 32; ++c) tma_load_2d(sb + c * (BK * 128), &tmB, bar, n0 + 32 * c, k0);
            }
          } else {
            tma_load_2d(sb, &tmB, bar, k0, n0);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int g = 0, t = 0;
      for (int item = item0; item < n_items; item += item_step, ++t) {
        int m0, n0, bz, kb_begin, nkb;
        decode(item, m0, n0, bz, kb_begin, nkb);
        const int ab = SPLIT ? (t & 1) : 0;
        const uint32_t tacc = tmem_base + ab * BN;
        if (SPLIT) {
          mbar_wait(tempty0 + 8 * ab, ((t >> 1) & 1) ^ 1);        // the epilogue has drained this accumulator
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int i = 0; i < nkb; ++i, ++g) {
          const int s = g % STAGES, it = g / STAGES;
          mbar_wait((twins ? ready0 : full0) + 8 * s, it & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
          const uint32_t idesc = make_idesc(A_MN, B_MN, (kb_begin + i) >= kb_neg);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            // K-major: 8 floats = 32 bytes along the swizzled row; MN-major: 8 K-rows = 1024 bytes
            const uint32_t oa = A_MN ? kk * 1024 : kk * 32, ob = B_MN ? kk * 1024 : kk * 32;
            auto adesc = [&](uint32_t base) { return A_MN ? make_desc(base + oa, BK * 128, 512, 1) : make_desc(base + oa, 16, 1024, 2); };
            auto bdesc = [&](uint32_t base) { return B_MN ? make_desc(base + ob, BK * 128, 512, 1) : make_desc(base + ob, 16, 1024, 2); };
            umma_tf32(tacc, adesc(sa), bdesc(sb), idesc, (i > 0 || kk > 0) ? 1u : 0u);
            if (SPLIT) {
              // x = trunc(x) + lo(x): the two cross terms restore the bits the tensor core drops
              if (lo_mask & 1) umma_tf32(tacc, adesc(sa + LO_OFF), bdesc(sb), idesc, 1u);
              if (lo_mask & 2) umma_tf32(tacc, adesc(sa), bdesc(sb + LO_OFF), idesc, 1u);
            }
          }
          umma_commit(empty0 + 8 * s);          // frees the smem stage when these MMAs have read it
        }
        umma_commit(tfull0 + 8 * ab);           // accumulator complete
      }
    }
  } else if (SPLIT && warp < EW0) {
    // ===== lo twins: x - trunc_tf32(x), element by element (the swizzle is the same on both sides) =====
    if (twins) {
      const int te = threadIdx.x - 64;
      int g = 0;
      for (int item = item0; item < n_items; item += item_step) {
        int m0, n0, bz, kb_begin, nkb;
        decode(item, m0, n0, bz, kb_begin, nkb);
        if (EPI == EPI_UPDATE) prefetch_tile(m0, n0, te, N_XFORM);
        for (int i = 0; i < nkb; ++i, ++g) {
          const int s = g % STAGES, it = g / STAGES;
          mbar_wait(full0 + 8 * s, it & 1);
          float4* st = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (!(lo_mask & (1 << half))) continue;
            const float4* src = st + half * (A_BYTES / 16);
            float4* dst = st + (LO_OFF + half * A_BYTES) / 16;
#pragma unroll
            for (int r = 0; r < A_BYTES / 16 / N_XFORM; ++r) {
              const float4 x = src[te + r * N_XFORM];
              dst[te + r * N_XFORM] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> the MMA's async-proxy reads
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ready0 + 8 * s) : "memory");
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (transposed through shared memory) -> bias / activation / sampling or
    //       the W update -> global.  tcgen05.ld hands every thread one ROW of the tile; stored like that, a warp
    //       instruction would touch 32 different cache lines.  Each warp therefore parks 32 rows x 16 columns in
    //       shared memory and re-reads them as quads: 4 lanes cover a 64-byte row segment, a warp instruction covers 8
    //       rows — whole sectors, coalesced. =====
    const int ew = warp - EW0;
    const int quarter = warp & 3;           // a warp may only touch its own 32 TMEM lanes
    const int chalf = ew >> 2;              // ... and the two warps of a quarter split the columns
    float* T = reinterpret_cast<float*>(SPLIT ? smem + STAGES * STAGE_BYTES + 256 : smem) + ew * (32 * TLD);
    const int trow = lane >> 2, tcol = 4 * (lane & 3);
    int t = 0;
    for (int item = item0; item < n_items; item += item_step, ++t) {
      int m0, n0, bz, kb_begin, nkb;
      decode(item, m0, n0, bz, kb_begin, nkb);
      const int ab = SPLIT ? (t & 1) : 0;
      if (EPI == EPI_UPDATE && !SPLIT) prefetch_tile(m0, n0, (int)threadIdx.x - 32 * EW0, 256);
      mbar_wait(tfull0 + 8 * ab, SPLIT ? ((t >> 1) & 1) : 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c = chalf * (BN / 32); c < (chalf + 1) * (BN / 32); ++c) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * BN + c * 16, r);
        if (SPLIT && c == (chalf + 1) * (BN / 32) - 1) {
          // last read of this accumulator by this warp: hand it back to the MMA issuer
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8 * ab) : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<uint4*>(T + lane * TLD + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        __syncwarp();
        const int n = n0 + c * 16 + tcol;
        if (nkb <= 0) continue;
        if (EPI == EPI_UPDATE) {
          // The tile is rows m <= V (m == V: the all-ones visible unit) x columns n <= H (n == H: the all-ones hidden
          // unit).  W loads of all row segments are issued before the first use (the tile comes from L2 at best, HBM
          // at worst); padding columns of W (zeros) are left alone.
          if (n <= ep.uH) {
            const bool hs = ep.uSnap != nullptr, whole = n + 4 <= ep.uH, wquad = n < ep.uldw;
            float4 w4[4], s4[4], n4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + quarter * 32 + 8 * i + trow;
              const size_t o = (size_t)min(m, ep.uV - 1) * ep.uldw + n;
              w4[i] = wquad ? *reinterpret_cast<const float4*>(ep.uW + o) : make_float4(0.f, 0.f, 0.f, 0.f);
              s4[i] = wquad ? *reinterpret_cast<const float4*>(ep.uS + o) : make_float4(0.f, 0.f, 0.f, 0.f);
              n4[i] = (hs && wquad) ? *reinterpret_cast<const float4*>(ep.uSnap + o) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = 8 * i + trow;
              const int m = m0 + quarter * 32 + row;
              if (m >= M) continue;
              const float4 a4v = *reinterpret_cast<const float4*>(T + row * TLD + tcol);
              const float a4[4] = {a4v.x, a4v.y, a4v.z, a4v.w};
              if (m == ep.uV) {                                   // hidden-bias row
#pragma unroll
                for (int tt = 0; tt < 4; ++tt)
                  if (n + tt < ep.uH) update_bias_one(ep, ep.uhb + n + tt, ep.uShb + n + tt, a4[tt]);
                continue;
              }
              float4 wo, so;
              update_one(ep.u, a4v.x, w4[i].x, s4[i].x, n4[i].x, hs, wo.x, so.x);
              update_one(ep.u, a4v.y, w4[i].y, s4[i].y, n4[i].y, hs, wo.y, so.y);
              update_one(ep.u, a4v.z, w4[i].z, s4[i].z, n4[i].z, hs, wo.z, so.z);
              update_one(ep.u, a4v.w, w4[i].w, s4[i].w, n4[i].w, hs, wo.w, so.w);
              const size_t o = (size_t)m * ep.uldw + n;
              if (whole) {
                *reinterpret_cast<float4*>(ep.uW + o) = wo;
                *reinterpret_cast<float4*>(ep.uS + o) = so;
              } else {
                const float wf[4] = {wo.x, wo.y, wo.z, wo.w}, sf[4] = {so.x, so.y, so.z, so.w};
#pragma unroll
                for (int tt = 0; tt < 4; ++tt) {
                  if (n + tt < ep.uH) { ep.uW[o + tt] = wf[tt]; ep.uS[o + tt] = sf[tt]; }
                  else if (n + tt == ep.uH) update_bias_one(ep, ep.uvb + m, ep.uSvb + m, a4[tt]);      // visible-bias column
                }
              }
            }
          }
          continue;
        }
#pragma unroll 2
        for (int i = 0; i < 4; ++i) {
          const int row = 8 * i + trow;
          const int m = m0 + quarter * 32 + row;
          if (m >= M) continue;
          const float4 a4v = *reinterpret_cast<const float4*>(T + row * TLD + tcol);
          const float a4[4] = {a4v.x, a4v.y, a4v.z, a4v.w};
          if (EPI == EPI_PART) {
            float* dst = ep.part + ((size_t)bz * M + m) * N + n;
            if (ep.vec4 && n + 4 <= N) {
              *reinterpret_cast<float4*>(dst) = a4v;
            } else {
#pragma unroll
              for (int tt = 0; tt < 4; ++tt)
                if (n + tt < N) dst[tt] = a4[tt];
            }
          } else if (ep.vec4 && n + 4 <= N) {
            epi_quad(ep, m, n, N, a4);
          } else {
#pragma unroll
            for (int tt = 0; tt < 4; ++tt)
              if (n + tt < N) epi_one(ep, m, n + tt, N, a4[tt]);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols(SPLIT)) : "memory");
  }
}

// split-K tail of propup / propdown: partial tiles summed in fixed order, then the same fused epilogue
__global__ void part_act_kernel(const float* __restrict__ part, int splits, int M, int N, EpiParams ep) {
  const long long total = (long long)M * N;
  if (ep.vec4) {
    const long long nq = total >> 2;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
      const long long e = q << 2;
      const int m = (int)(e / N), n = (int)(e - (long long)m * N);
      float4 s = *reinterpret_cast<const float4*>(part + e);
      for (int z = 1; z < splits; ++z) {
        const float4 t = *reinterpret_cast<const float4*>(part + (size_t)z * total + e);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      const float a4[4] = {s.x, s.y, s.z, s.w};
      epi_quad(ep, m, n, N, a4);
    }
  } else {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
      const int m = (int)(e / N), n = (int)(e - (long long)m * N);
      float s = 0.f;
      for (int z = 0; z < splits; ++z) s += part[(size_t)z * total + e];
      epi_one(ep, m, n, N, s);
    }
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
  bool exact;         // every value is exactly representable in TF32 ({0,1} samples): no lo twin in SPLIT mode
  bool scratch3d;     // MN-major library scratch (slack behind the last row): whole-tile 3-D TMA boxes
};
// MN-major operand [K][MN] viewed as [chunk = MN/32][k][32]: ONE box {32, BK, tile/32} per k-block lands in shared
// memory exactly like the per-chunk 2-D boxes (chunk-major, 128-byte rows, 32-byte-atom swizzle).  The last chunk of a
// ragged MN reads on into the next row (columns the epilogue discards) and, on the last row, up to 124 bytes past the
// matrix — only for scratch buffers, which have that slack.
static int make_map3(CUtensorMap* tm, const float* ptr, long long mn, long long k, long long ld, int tile) {
  static thread_local MapCache cache;
  const MapKey kq{ptr, mn, k, ld, tile, -3, 1};
  for (int i = 0; i < cache.used; ++i)
    if (cache.key[i] == kq) { *tm = cache.map[i]; return 0; }
  EncodeFn enc = get_encode();
  MDBN_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  MDBN_CHECK(((uintptr_t)ptr & 15) == 0 && (ld * 4) % 16 == 0, "TMA operand must be 16-byte aligned (ptr %p ld %lld)",
             (const void*)ptr, ld);
  cuuint64_t dims[3] = {32, (cuuint64_t)k, (cuuint64_t)((mn + 31) / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)BK, (cuuint32_t)(tile / 32)};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MDBN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with %d", (int)r);
  const int slot = cache.used < MapCache::N ? cache.used++ : (cache.next++ % MapCache::N);
  cache.key[slot] = kq;
  cache.map[slot] = *tm;
  return 0;
}

// split K over the SMs when the output has too few tiles to occupy them (skinny shapes); 1 = no split
static int pick_splits(const mdbn_ctx* c, int M, int N, int K, bool split3) {
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN), nkb = (K + BK - 1) / BK;
  const int slots = c->num_sms * (split3 ? 1 : 2);      // resident CTAs
  if (2 * tiles > slots || nkb < 8) return 1;
  int s = slots / tiles;
  if (s > nkb / 4) s = nkb / 4;
  return s < 1 ? 1 : s;
}

template <bool A_MN, bool B_MN, int EPI, bool SPLIT>
static int launch_gemm_t(mdbn_ctx* c, const Operand& A, const Operand& Bo, int M, int N, int K, int splits, int kneg,
                         const EpiParams& ep, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  const int mn3d = ((A_MN && A.scratch3d) ? 1 : 0) | ((B_MN && Bo.scratch3d) ? 2 : 0);
  if (mn3d & 1) MDBN_TRY(make_map3(&tmA, A.ptr, M, K, A.ld, BM));
  else if (A_MN) MDBN_TRY(make_map(&tmA, A.ptr, M, K, A.ld, 32, BK, true));
  else MDBN_TRY(make_map(&tmA, A.ptr, K, M, A.ld, BK, BM, false));
  if (mn3d & 2) MDBN_TRY(make_map3(&tmB, Bo.ptr, N, K, Bo.ld, BN));
  else if (B_MN) MDBN_TRY(make_map(&tmB, Bo.ptr, N, K, Bo.ld, 32, BK, true));
  else MDBN_TRY(make_map(&tmB, Bo.ptr, K, N, Bo.ld, BK, BN, false));
  auto kfn = tc_gemm_kernel<A_MN, B_MN, EPI, SPLIT>;
  static bool configured[64] = {};
  if (!configured[c->device]) {
    MDBN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(SPLIT)));
    configured[c->device] = true;
  }
  const int nkb = (K + BK - 1) / BK;
  int kbps = (nkb + splits - 1) / splits;
  const int nbx = (N + BN - 1) / BN, nby = (M + BM - 1) / BM, nbz = (nkb + kbps - 1) / kbps;
  const int lo_mask = SPLIT ? ((A.exact ? 0 : 1) | (Bo.exact ? 0 : 2)) : 0;
  // SPLIT: persistent CTAs, one per SM, walking the work items; plain: one CTA per item, two resident per SM
  const dim3 grid = SPLIT ? dim3(nbx * nby * nbz < c->num_sms ? nbx * nby * nbz : c->num_sms) : dim3(nbx, nby, nbz);
  kfn<<<grid, n_threads(SPLIT), smem_bytes(SPLIT), st>>>(tmA, tmB, M, N, K, kbps, kneg / BK, lo_mask, mn3d, nbx, nby, nbz, ep);
  c->launches++;
  MDBN_CUDA(cudaGetLastError());
  return 0;
}
template <bool A_MN, bool B_MN, int EPI>
static int launch_gemm(mdbn_ctx* c, bool split3, const Operand& A, const Operand& Bo, int M, int N, int K, int splits,
                       int kneg, const EpiParams& ep, cudaStream_t st) {
  return split3 ? launch_gemm_t<A_MN, B_MN, EPI, true>(c, A, Bo, M, N, K, splits, kneg, ep, st)
                : launch_gemm_t<A_MN, B_MN, EPI, false>(c, A, Bo, M, N, K, splits, kneg, ep, st);
}
// number of K slices a launch with `splits` really has
static int n_slices(int K, int splits) {
  const int nkb = (K + BK - 1) / BK, kbps = (nkb + splits - 1) / splits;
  return (nkb + kbps - 1) / kbps;
}

// ---- small ld-aware helpers of the tensor path ----------------------------------
// rows b >= B of the batch tile (the statistics GEMM negates the nv/nh half per 32-row K block, so a minibatch that is
// not a multiple of 32 is padded with zero rows) are written as zeros
// The gather also plants the all-ones units of the statistics GEMM: column V of both halves of XV and column H of both
// halves of YH are 1 in the rows of the minibatch (0 in the padding rows), so that row V / column H of
// [v0;nv]^T (+/-) [ph;nh] are the bias gradients sum_b (ph - nh) and sum_b (v0 - nv).
__device__ __forceinline__ void plant_ones(int b, int B, int Bp, float* XV, long long ldx, int V, float* YH, long long ldy,
                                           int H) {
  const float one = b < B ? 1.f : 0.f;
  XV[(size_t)b * ldx + V] = one;
  XV[(size_t)(Bp + b) * ldx + V] = one;
  YH[(size_t)b * ldy + H] = one;
  YH[(size_t)(Bp + b) * ldy + H] = one;
}
__global__ void gather_rows_ld_kernel(const float* __restrict__ data, long long ld, const int* __restrict__ idx, int B,
                                      int Bp, int V, float* __restrict__ out, long long ldo, float* __restrict__ xi,
                                      float* __restrict__ YH, long long ldy, int H) {
  const int b = blockIdx.y;      // rows b >= B are the zero padding
  const bool real = b < B;
  const float* src = data + (real ? (idx ? idx[b] : b) : 0) * ld;
  if (blockIdx.x == 0 && threadIdx.x == 0) plant_ones(b, B, Bp, out, ldo, V, YH, ldy, H);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V; i += gridDim.x * blockDim.x) {
    const float x = real ? __ldg(src + i) : 0.f;
    out[(size_t)b * ldo + i] = x;
    if (xi) xi[(size_t)b * ldo + i] = roundf(x);
  }
}
// the same, four columns per thread (V, strides and pointers multiples of 4 / 16 bytes): one row per blockIdx.y
__global__ void gather_rows_ld4_kernel(const float* __restrict__ data, long long ld, const int* __restrict__ idx, int B,
                                       int Bp, int V4, float* __restrict__ out, long long ldo, float* __restrict__ xi,
                                       float* __restrict__ YH, long long ldy, int H) {
  const int b = blockIdx.y;
  const bool real = b < B;
  const long long r = real ? (idx ? idx[b] : b) : 0;
  const float4* src = reinterpret_cast<const float4*>(data + r * ld);
  float4* dst = reinterpret_cast<float4*>(out + (size_t)b * ldo);
  float4* dx = xi ? reinterpret_cast<float4*>(xi + (size_t)b * ldo) : nullptr;
  if (blockIdx.x == 0 && threadIdx.x == 0) plant_ones(b, B, Bp, out, ldo, 4 * V4, YH, ldy, H);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V4; i += gridDim.x * blockDim.x) {
    const float4 x = real ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    dst[i] = x;
    if (dx) dx[i] = make_float4(roundf(x.x), roundf(x.y), roundf(x.z), roundf(x.w));
  }
}
// zero the padding rows [B, Bp) of the nv half of XV and of both halves of YH (scratch memory: a stale NaN times a
// zero row of the other operand would poison the statistics)
__global__ void zero_pad_rows_kernel(float* __restrict__ nv, long long ldx, int V, float* __restrict__ ph,
                                     float* __restrict__ nh, long long ldy, int H, int npad) {
  const long long nx = (long long)npad * ldx, ny = (long long)npad * ldy;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nx + 2 * ny;
       e += (long long)gridDim.x * blockDim.x) {
    if (e < nx) nv[e] = 0.f;
    else if (e < nx + ny) ph[e - nx] = 0.f;
    else nh[e - nx - ny] = 0.f;
  }
}
// raw column sums of (top half - bottom half) of a [Bp + B, N] matrix (the bottom half starts at row Bp), two
// deterministic stages: row chunks in parallel, then a fixed-order sum of the chunk partials
__global__ void col_diff_partial_kernel(const float* __restrict__ X, long long ld, int B, int Bp, int N, int rows_per_chunk,
                                        float* __restrict__ partial) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int b0 = blockIdx.y * rows_per_chunk, b1 = min(B, b0 + rows_per_chunk);
  float p = 0.f, q = 0.f;
  for (int b = b0; b < b1; ++b) { p += X[(size_t)b * ld + n]; q += X[(size_t)(Bp + b) * ld + n]; }
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
static int col_diff_sum(mdbn_ctx* c, const float* X, long long ld, int B, int Bp, int N, float* out, cudaStream_t st) {
  const int rpc = 32, chunks = (B + rpc - 1) / rpc;
  float* part = (float*)ws_get(c, WS_MISC, (size_t)chunks * N * sizeof(float));
  if (!part) return 3;
  col_diff_partial_kernel<<<dim3((N + 255) / 256, chunks), 256, 0, st>>>(X, ld, B, Bp, N, rpc, part);
  col_diff_final_kernel<<<(N + 7) / 8, 256, 0, st>>>(part, chunks, N, out);
  c->launches += 2;
  return 0;
}
// (plain TF32 mode only: the split mode gets the bias gradients from the statistics GEMM)
// Both bias gradients of a step with a moderate batch in ONE launch: column n < H is a hidden unit (rows of YH), the
// rest are visible units (rows of XV); sum over the rows of (positive - negative), then either the raw sum goes to the
// packed statistics (gsum != NULL) or the bias and its speed are updated in place (src/rbm.py:416-417, :361-364).
// Block = 32 columns x 8 row groups (coalesced 128-byte reads, 8 loads in flight per column), combined in fixed order.
__global__ void bias_tail_kernel(const float* __restrict__ YH, long long ldy, int H, const float* __restrict__ XV,
                                 long long ldx, int V, int B, int Bp, const float* __restrict__ presum,
                                 float* __restrict__ gsum, float* __restrict__ hb,
                                 float* __restrict__ Shb, float* __restrict__ vb, float* __restrict__ Svb, float inv_rows,
                                 float mom, float lr) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const bool ok = n < H + V;
  const bool hid = n < H;
  float d = 0.f;
  if (ok && !presum) {
    const float* X = hid ? YH + n : XV + (n - H);
    const long long ld = hid ? ldy : ldx;
    float p = 0.f, q = 0.f;
    for (int b = ty; b < B; b += 8) { p += X[(size_t)b * ld]; q += X[(size_t)(Bp + b) * ld]; }
    d = p - q;
  }
  red[ty][tx] = d;
  __syncthreads();
  if (ty != 0 || !ok) return;
  if (presum) {
    d = presum[n];          // (large batch: the sums were built by the two-stage column reduction)
  } else {
    d = red[0][tx];
#pragma unroll
    for (int t = 1; t < 8; ++t) d += red[t][tx];
  }
  if (gsum) { gsum[n] = d; return; }
  float* bias = hid ? hb + n : vb + (n - H);
  float* S = hid ? Shb + n : Svb + (n - H);
  const float g = d * inv_rows, s = *S;
  *S = g + (s - g) * mom;
  *bias = *bias + s * lr;
}
// reconstruction-cost numerators, one minibatch row per blockIdx.y (src/rbm.py:479-480 CE; :697 MSE with sigma of the
// linear mean): partial[blockIdx.y * gridDim.x + blockIdx.x]
__global__ void recon_cost_ld_kernel(const float* __restrict__ prev, const float* __restrict__ v0, long long ld, int B,
                                     int V, int kind, float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  const int b = blockIdx.y;
  const float* pr = prev + (size_t)b * ld;
  const float* tr = v0 + (size_t)b * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V; i += gridDim.x * blockDim.x) {
    const float p = pr[i], t = tr[i];
    if (kind == MDBN_GRBM) { float d = sigmoidf_(p) - t; s += d * d; }
    else s += t * softplusf_(-p) + (1.f - t) * softplusf_(p);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
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
// fixed-order sum of the cost partials; the packed statistics get (numerator, rows), a full step its cost
// (numerator * inv_den); the pseudo-likelihood cursor advances here too (src/rbm.py:445)
__global__ void sum_tree_kernel(const float* __restrict__ partial, int n, float* __restrict__ out, float rows,
                                float* __restrict__ rows_out, float* __restrict__ cost_out, float inv_den,
                                int* __restrict__ bit_idx, int V) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    if (out) { *out = s; *rows_out = rows; }
    if (cost_out) *cost_out = s * inv_den;
    if (bit_idx) *bit_idx = (*bit_idx + 1) % V;
  }
}
// Split-K statistics [splits][V+1][H+1] (row V / column H: the bias gradients) summed in fixed order.
// STATS (data-parallel shard): scattered into the packed buffer [V*H | H | V].
__global__ void reduce_parts_kernel(const float* __restrict__ part, int splits, int V, int H, int ones, float* __restrict__ G) {
  const int H1 = H + ones;
  const long long n = (long long)(V + ones) * H1;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / H1), j = (int)(e - (long long)i * H1);
    float g = 0.f;
    for (int z = 0; z < splits; ++z) g += part[(size_t)z * n + e];
    if (i < V && j < H) G[(size_t)i * H + j] = g;
    else if (i == V && j < H) G[(size_t)V * H + j] = g;
    else if (j == H && i < V) G[(size_t)V * H + H + i] = g;
  }
}
// Full step: the update of W / W_speed and of both biases rides on the reduction of the partials.
__global__ void reduce_update_kernel(const float* __restrict__ part, int splits, int V, int H, int ones, EpiParams ep) {
  const int H1 = H + ones;
  const long long total = (long long)(V + ones) * H1;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / H1), j = (int)(e - (long long)i * H1);
    float g = 0.f;
    for (int z = 0; z < splits; ++z) g += part[(size_t)z * total + e];
    if (i < V && j < H) {
      const size_t o = (size_t)i * ep.uldw + j;
      float wo, so;
      update_one(ep.u, g, ep.uW[o], ep.uS[o], ep.uSnap ? ep.uSnap[o] : 0.f, ep.uSnap != nullptr, wo, so);
      ep.uW[o] = wo;
      ep.uS[o] = so;
    } else if (i == V && j < H) {
      update_bias_one(ep, ep.uhb + j, ep.uShb + j, g);
    } else if (j == H && i < V) {
      update_bias_one(ep, ep.uvb + i, ep.uSvb + i, g);
    }
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
  return N % 4 == 0 && ep.ld_pre % 4 == 0 && ep.ld_mean % 4 == 0 && ep.ld_sample % 4 == 0 && (a & 15) == 0;
}
// one propagation: the GEMM with the fused epilogue, or — skinny shapes — split-K partials + the reduction kernel
// that applies the same epilogue
template <bool B_MN>
static int propagate(mdbn_ctx* c, bool split3, const Operand& A, const Operand& Bo, int M, int N, int K, EpiParams ep,
                     cudaStream_t st) {
  ep.vec4 = vec_ok(ep, N);
  const int splits = pick_splits(c, M, N, K, split3);
  if (splits <= 1) return launch_gemm<false, B_MN, EPI_ACT>(c, split3, A, Bo, M, N, K, 1, 1 << 30, ep, st);
  const int zs = n_slices(K, splits);
  float* part = (float*)ws_get(c, WS_PART, (size_t)zs * M * N * sizeof(float));
  if (!part) return 3;
  EpiParams pp{};
  pp.part = part;
  pp.vec4 = N % 4 == 0;
  MDBN_TRY((launch_gemm<false, B_MN, EPI_PART>(c, split3, A, Bo, M, N, K, splits, 1 << 30, pp, st)));
  const long long work = ep.vec4 ? ((long long)M * N) >> 2 : (long long)M * N;
  int blocks = (int)((work + 255) / 256);
  if (blocks > 4 * c->num_sms) blocks = 4 * c->num_sms;
  part_act_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(part, zs, M, N, ep);
  c->launches++;
  MDBN_CUDA(cudaGetLastError());
  return 0;
}
static int up(mdbn_ctx* c, bool split3, const float* W, int ldw, const float* hb, int B, int V, int H, const float* x,
              long long ldx, bool x_exact, float* pre, float* mean, float* sample, long long ldo, long long ld_s,
              const RngSeg& rs, cudaStream_t st) {
  EpiParams ep{};
  ep.bias = hb; ep.act = ACT_SIGMOID; ep.smp = sample ? SMP_BERNOULLI : SMP_NONE; ep.rs = rs;
  ep.pre = pre; ep.mean = mean; ep.sample = sample; ep.ld_pre = ep.ld_mean = ldo; ep.ld_sample = ld_s;
  return propagate<true>(c, split3, Operand{x, ldx, false, x_exact, false}, Operand{W, ldw, true, false, false}, B, H, V, ep, st);
}
static int down(mdbn_ctx* c, bool split3, const float* W, int ldw, const float* vb, int B, int V, int H, int kind, int noisy,
                const float* h, long long ldh, bool h_exact, float* pre, float* mean, float* sample, long long ldo,
                const RngSeg& rs, cudaStream_t st) {
  EpiParams ep{};
  ep.bias = vb; ep.act = kind == MDBN_GRBM ? ACT_LINEAR : ACT_SIGMOID;
  ep.smp = !sample ? SMP_NONE : (kind == MDBN_GRBM ? (noisy ? SMP_GAUSS : SMP_MEAN) : SMP_BERNOULLI);
  ep.rs = rs;
  ep.pre = pre; ep.mean = mean; ep.sample = sample; ep.ld_pre = ep.ld_mean = ep.ld_sample = ldo;
  return propagate<false>(c, split3, Operand{h, ldh, false, h_exact, false}, Operand{W, ldw, false, false, false}, B, V, H, ep, st);
}

}  // namespace tc

bool tensor_phase_supported(const void* W, int ldw, const void* x, long long ldx) {
  return ldw % 4 == 0 && ldx % 4 == 0 && (((uintptr_t)W | (uintptr_t)x) & 15) == 0 && tc::get_encode() != nullptr;
}
// single-phase calls: fp32-exact (SPLIT) unless the context allows plain TF32 (mdbn_set_tf32_phases)
int tensor_propup(mdbn_ctx* c, const float* W, int ldw, const float* hb, const float* v, int ldv, int B, int V, int H,
                  float* pre, float* mean, float* sample, const RngSeg& rs, cudaStream_t st) {
  return tc::up(c, !c->tf32_phases, W, ldw, hb, B, V, H, v, ldv, false, pre, mean, sample, H, H, rs, st);
}
int tensor_propdown(mdbn_ctx* c, const float* W, int ldw, const float* vb, const float* h, int ldh, int B, int V, int H,
                    int kind, int noisy, float* pre, float* mean, float* sample, const RngSeg& rs, cudaStream_t st) {
  return tc::down(c, !c->tf32_phases, W, ldw, vb, B, V, H, kind, noisy, h, ldh, false, pre, mean, sample, V, rs, st);
}
// F(v): the v W product on the tensor cores as split-K partials, softplus + row reduction in the tail kernel
int tensor_free_energy(mdbn_ctx* c, const float* W, int ldw, const float* hb, const float* vb, const float* v, int ldv,
                       int B, int V, int H, int kind, float* F, cudaStream_t st) {
  using namespace tc;
  const bool split3 = !c->tf32_phases;
  const int splits = pick_splits(c, B, H, V, split3), zs = n_slices(V, splits);
  float* part = (float*)ws_get(c, WS_PART, (size_t)zs * B * H * sizeof(float));
  if (!part) return 3;
  EpiParams pp{};
  pp.part = part;
  pp.vec4 = H % 4 == 0;
  MDBN_TRY((launch_gemm<false, true, EPI_PART>(c, split3, Operand{v, ldv, false, false, false}, Operand{W, ldw, true, false, false}, B, H,
                                               V, splits, 1 << 30, pp, st)));
  return free_energy_from_parts(c, part, zs, B, H, hb, v, ldv, V, vb, kind, F, st);
}

bool tensor_supported(const mdbn_ctx*, const mdbn_cd_args& a) {
  if (a.ldw % 4 != 0 || ((uintptr_t)a.W & 15) || ((uintptr_t)a.W_speed & 15) || ((uintptr_t)a.W_snap & 15)) return false;
  if (a.phase == MDBN_PHASE_APPLY) return false;
  return tc::get_encode() != nullptr;
}

int tensor_cd_step(mdbn_ctx* c, const mdbn_cd_args& a, cudaStream_t st) {
  using namespace tc;
  const int B = a.B, V = a.V, H = a.H, k = a.k;
  const int Bp = (B + BK - 1) / BK * BK;       // the nv/nh half of the statistics GEMM is negated per 32-row K block
  const bool split3 = !a.tf32;                 // fp32-exact unless the caller allows plain TF32
  const bool full = a.phase == MDBN_PHASE_FULL;
  const long long VH = (long long)V * H;
  // activation scratch: one spare column behind the V visible / H hidden units for the all-ones unit (plant_ones)
  const long long ldx = (V + 1 + 3) & ~3, ldy = (H + 1 + 3) & ~3, ldhs = (H + 3) & ~3;
  float* G = full ? nullptr : a.stats_buf;
  MDBN_CHECK(full || G != nullptr, "cd_step: stats buffer missing");
  float* XV = (float*)ws_get(c, WS_XV, (size_t)2 * Bp * ldx * sizeof(float));
  float* YH = (float*)ws_get(c, WS_YH, (size_t)2 * Bp * ldy * sizeof(float));
  // PCD chain state: the caller's [B, H] array IS the chain buffer when its row stride suits TMA (H % 4 == 0, 16-byte
  // aligned): the Gibbs steps read and overwrite it in place, no copy in, no copy out (src/rbm.py:308-311, :369)
  const bool chain_in_place = a.persistent != nullptr && ldhs == H && (((uintptr_t)a.persistent) & 15) == 0;
  float* HS = chain_in_place ? a.persistent : (float*)ws_get(c, WS_HS, (size_t)B * ldhs * sizeof(float));
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
  float* nv_mean = XV + (size_t)Bp * ldx;
  float* nh_mean = YH + (size_t)Bp * ldy;
  // v0 = data[indices] (+ round(v0) for the pseudo-likelihood) and the all-ones units; padding rows are written as zeros
  if (V % 4 == 0 && a.ld_data % 4 == 0 && (((uintptr_t)a.data) & 15) == 0)
    gather_rows_ld4_kernel<<<dim3((V / 4 + 255) / 256, pcd ? B : Bp), 256, 0, st>>>(a.data, a.ld_data, a.indices, B, Bp, V / 4,
                                                                                    XV, ldx, XI, YH, ldy, H);
  else
    gather_rows_ld_kernel<<<dim3((V + 1023) / 1024, pcd ? B : Bp), 256, 0, st>>>(a.data, a.ld_data, a.indices, B, Bp, V, XV, ldx,
                                                                                 XI, YH, ldy, H);
  c->launches++;
  if (Bp != B) {
    if (pcd) MDBN_CUDA(cudaMemsetAsync(XV + (size_t)B * ldx, 0, (size_t)(Bp - B) * ldx * sizeof(float), st));   // (XI has B rows)
    zero_pad_rows_kernel<<<64, 256, 0, st>>>(nv_mean + (size_t)B * ldx, ldx, V, YH + (size_t)B * ldy, nh_mean + (size_t)B * ldy,
                                             ldy, H, Bp - B);
    c->launches++;
  }
  // positive phase
  MDBN_TRY(up(c, split3, a.W, a.ldw, a.hbias, B, V, H, XV, ldx, false, nullptr, YH, pcd ? nullptr : HS, ldy, ldhs,
              make_seg(a.rng, ul.off_h0, 0), st));
  if (pcd) {
    MDBN_TRY(up(c, split3, a.W, a.ldw, a.hbias, B, V, H, XI, ldx, false, PREX, nullptr, nullptr, ldy, ldhs,
                make_seg(a.rng, 0, 0), st));
    if (!chain_in_place) {
      copy_rows_kernel<<<eb, 256, 0, st>>>(a.persistent, H, HS, ldhs, B, H);      // chain state, padded stride for TMA
      c->launches++;
    }
  }
  for (int s = 0; s < k; ++s) {
    long long base = (long long)B * H + s * ul.step_stride;
    // (the chain state is {0,1} once this step has sampled it; a caller-provided persistent chain is not assumed to be)
    MDBN_TRY(down(c, split3, a.W, a.ldw, a.vbias, B, V, H, a.kind, a.noisy, HS, ldhs, !(pcd && s == 0),
                  a.kind == MDBN_GRBM ? nullptr : PREV, nv_mean,
                  a.kind == MDBN_RBM ? VS : nullptr, ldx, make_seg(a.rng, base + ul.off_v, ord_v(s)), st));
    const float* v_in = a.kind == MDBN_GRBM ? nv_mean : VS;
    MDBN_TRY(up(c, split3, a.W, a.ldw, a.hbias, B, V, H, v_in, ldx, a.kind == MDBN_RBM, nullptr, nh_mean, HS, ldy, ldhs,
                make_seg(a.rng, base + ul.off_h, ord_h(s)), st));
  }
  // monitoring cost on the OLD parameters (src/rbm.py:367-374), before anything is updated
  {
    float* num = full ? nullptr : G + VH + H + V;
    const float den = (!pcd && a.kind == MDBN_GRBM) ? (float)B * (float)V : (float)B;      // :697 / :479-482
    float* cost = full ? a.cost_out : nullptr;
    if (pcd) {
      pl_row_ld_kernel<<<B, 128, 0, st>>>(PREX, ldy, H, XI, ldx, V, a.W, a.ldw, a.vbias, a.bit_i_idx, a.kind, RED);
      sum_tree_kernel<<<1, 1024, 0, st>>>(RED, B, num, (float)B, num ? num + 1 : nullptr, cost, 1.0f / den, a.bit_i_idx, V);
      c->launches += 2;
      if (!chain_in_place) {
        copy_rows_kernel<<<eb, 256, 0, st>>>(HS, ldhs, a.persistent, H, B, H);
        c->launches++;
      }
    } else {
      // (the linear Gaussian mean IS the pre-activation: no separate array for it)
      int nbx = (V + 1023) / 1024;
      while ((long long)nbx * B > 4096) nbx = (nbx + 1) / 2;
      recon_cost_ld_kernel<<<dim3(nbx, B), 256, 0, st>>>(a.kind == MDBN_GRBM ? nv_mean : PREV, XV, ldx, B, V, a.kind, RED);
      sum_tree_kernel<<<1, 1024, 0, st>>>(RED, nbx * B, num, (float)B, num ? num + 1 : nullptr, cost, 1.0f / den, nullptr, V);
      c->launches += 2;
    }
  }
  // statistics: [v0;nv;1]^T (+/-) [ph;nh;1] over K = 2 Bp.  In the fp32-exact mode the extra row and column (the all-ones
  // units) are the bias gradients; plain TF32 would truncate v0 / nv / ph / nh inside those sums — differences of nearly
  // equal numbers — so that mode keeps its exact fp32 column sums (bias_tail_kernel / col_diff_sum).  A full step with
  // enough output tiles applies the update of W, W_speed (and the biases) in the GEMM's epilogue; few tiles (small
  // layers, large batch) split K across the SMs and the update rides on the reduction of the partials.  STATS
  // (data-parallel shard): the packed buffer gets the raw sums.
  {
    const int ones = split3 ? 1 : 0;
    const int V1 = V + ones, H1 = H + ones;
    const int tiles = ((V1 + BM - 1) / BM) * ((H1 + BN - 1) / BN);
    const int nkb = 2 * Bp / BK;
    const int slots = c->num_sms * (split3 ? 1 : 2);
    int splits = 1;
    if (2 * tiles <= slots) {
      splits = slots / tiles;          // (floor: one round of resident CTAs, never a ragged second one)
      if (splits > nkb / 2) splits = nkb / 2 > 0 ? nkb / 2 : 1;
      if (splits > 32) splits = 32;
    }
    const Operand Ao{XV, ldx, true, false, true}, Bo{YH, ldy, true, false, true};
    EpiParams ep{};
    ep.uW = a.W; ep.uS = a.W_speed; ep.uSnap = a.weightcost != 0.f ? a.W_snap : nullptr; ep.uldw = a.ldw;
    ep.u = make_update_scalars(a);
    ep.uV = V; ep.uH = ones ? H : H + (1 << 20);      // (no all-ones column: n == uH never happens)
    ep.uhb = a.hbias; ep.uShb = a.hbias_speed; ep.uvb = a.vbias; ep.uSvb = a.vbias_speed;
    ep.inv_rows = 1.0f / (float)B;
    if (full && splits == 1) {
      MDBN_TRY((launch_gemm<true, true, EPI_UPDATE>(c, split3, Ao, Bo, V1, H1, 2 * Bp, 1, Bp, ep, st)));
    } else {
      const int zs = n_slices(2 * Bp, splits);
      float* part = (float*)ws_get(c, WS_PART, (size_t)zs * V1 * H1 * sizeof(float));
      if (!part) return 3;
      EpiParams pp{};
      pp.part = part;
      pp.vec4 = H1 % 4 == 0;
      MDBN_TRY((launch_gemm<true, true, EPI_PART>(c, split3, Ao, Bo, V1, H1, 2 * Bp, splits, Bp, pp, st)));
      if (full) reduce_update_kernel<<<eb, 256, 0, st>>>(part, zs, V, H, ones, ep);
      else reduce_parts_kernel<<<eb, 256, 0, st>>>(part, zs, V, H, ones, G);
      c->launches++;
      if (!full && c->ev_stats_w) { MDBN_CUDA(cudaEventRecord(c->ev_stats_w, st)); c->ev_stats_w_done = true; }
    }
    if (!ones) {
      // plain TF32: exact fp32 bias gradients, raw sums into the packed buffer (STATS) or the in-place update (full step)
      if (B <= 1024) {
        bias_tail_kernel<<<(H + V + 31) / 32, 256, 0, st>>>(YH, ldy, H, XV, ldx, V, B, Bp, nullptr, full ? nullptr : G + VH,
                                                            a.hbias, a.hbias_speed, a.vbias, a.vbias_speed, 1.0f / (float)B,
                                                            a.momentum, a.lr);
        c->launches++;
      } else {
        float* gs = full ? (float*)ws_get(c, WS_G, (size_t)(H + V) * sizeof(float)) : G + VH;
        if (!gs) return 3;
        MDBN_TRY(col_diff_sum(c, YH, ldy, B, Bp, H, gs, st));
        MDBN_TRY(col_diff_sum(c, XV, ldx, B, Bp, V, gs + H, st));
        if (full) {
          bias_tail_kernel<<<(H + V + 31) / 32, 256, 0, st>>>(YH, ldy, H, XV, ldx, V, B, Bp, gs, nullptr, a.hbias,
                                                              a.hbias_speed, a.vbias, a.vbias_speed, 1.0f / (float)B,
                                                              a.momentum, a.lr);
          c->launches++;
        }
      }
    }
  }
  MDBN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mdbn
