// Skinny-batch CD-k / PCD-k step, tcgen05 edition (B <= 16, H <= 512): ONE persistent cooperative
// kernel per step whose two skinny GEMMs run on the 5th-generation tensor cores.
//
// The batch is far below tcgen05's minimum M (64/128), so the operand roles are swapped: W supplies
// the M side and the (padded-to-16) batch is N:
//   propup   out[j, b] = sum_i W[i][j] v[b][i]   A = W^T  MN-major, TMA boxes {32 j x 32 rows}, SW128/32B atoms
//   propdown out[i, b] = sum_j W[i][j] h[b][j]   A = W    K-major,  TMA boxes {32 j x 32 rows}, SW128
//   B operand = activation panel [16][K] in shared memory (K-major, no swizzle), written by the SIMT
//   epilogues; accumulators in TMEM (128 lanes x 16 columns per tile).
// fp32 accuracy on TF32 hardware: the tensor core truncates fp32 operands to TF32, so every staged W
// tile gets a "lo twin" (x - trunc(x)) written by transform warps, every panel has a lo panel, and each
// product is 2-3 MMAs (hi*hi + lo*hi [+ hi*lo]); {0,1} samples need no lo term.  Measured on the
// building blocks (scripts/ubench/tc_skinny_test.cu): 4-7e-7 relative to sum|terms|.
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-5 TMEM epilogue,
// warps 6-11 lo-twin transform; mbarrier ring full -> ready -> empty.  Around the two streaming passes the
// step keeps the structure of skinny.cu: per-CTA row slabs, partial hidden sums -> deterministic
// cross-CTA reduction between two grid barriers, fused statistics + update pass (SIMT).
#include <cuda.h>
#include <stdlib.h>
#include "ctx.h"

namespace mdbn {
namespace sktc {

constexpr int NT = 384, NWARP = NT / 32;
constexpr int NB = 16;                 // MMA N: batch rows padded to 16
constexpr int TILE = 16384;            // one staged W tile (TMA); its lo twin lives in a separate 2-deep ring
constexpr int MAX_STG = 8;             // hi tiles in flight: sized for HBM latency, not for the MMA hop
constexpr int NLO = 3;
constexpr int N_TRANSFORM_WARPS = 6;   // warps 6..11
constexpr int MAX_SBAR = 6;

struct Params {
  float *W, *S;
  const float* Wsnap;
  int ldw;
  float *hb, *vb, *Shb, *Svb;
  const float* data;
  long long ld_data;
  const int* idx;
  float* P;
  int* bit_idx;
  float* cost_out;
  int kind, noisy, B, V, H, k, pcd;
  float inv_bnom, inv_b, wc, c1, decay, mom, lr, cost_scale;
  int rng_mode;
  const float* ubuf;
  uint32_t k0, k1, c2, c3;
  long long u_step_stride, u_off_v, u_off_h;
  int rows_per_cta, n_active, CQ, GW, G, nstg, KH, KR, nmt;
  float* part;                  // [n_active][2][B][ldw]
  float *PH, *NH, *HS, *PREX;   // [NB][ldw]
  float* cost_part;
  unsigned long long* bar;
  unsigned long long* dbg;
  int off_hp, off_hplo, off_v0, off_v0lo, off_x, off_vin, off_vinlo, off_nv, off_vb, off_bars, off_misc;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;     // 0 none, 1 SW128 with 32B atoms (MN-major tf32), 2 SW128
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// D f32, A/B tf32, M = 128, N = n (16/32/48: stacked panels), B K-major
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// activation panel [16][KP], K-major, no swizzle: 8x4-float core matrices, 128 B apart along K,
// the two 8-row groups KP*32 bytes apart.  Offset in floats.
__device__ __forceinline__ int poff(int n, int k, int KP) { return (n >> 3) * KP * 8 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3); }
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ bool tf32_exact(float x) { return (__float_as_uint(x) & 0x1FFFu) == 0u; }

__device__ __forceinline__ void grid_sync(unsigned long long* bar, unsigned long long& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1ULL);
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

template <int BT>
__global__ void __launch_bounds__(NT, 1) cd_skinny_tc_kernel(const __grid_constant__ CUtensorMap tmMN32,
                                                             const __grid_constant__ CUtensorMap tmMN8,
                                                             const __grid_constant__ CUtensorMap tmK32,
                                                             const __grid_constant__ CUtensorMap tmK8, const Params p) {
  constexpr int BTP = (BT + 3) / 4 * 4;
  extern __shared__ __align__(1024) unsigned char smem[];
  float* hp = reinterpret_cast<float*>(smem + p.off_hp);        // h panel [16][KH]
  float* hplo = reinterpret_cast<float*>(smem + p.off_hplo);
  float* v0p = reinterpret_cast<float*>(smem + p.off_v0);       // row panels [16][KR]
  float* v0lo = reinterpret_cast<float*>(smem + p.off_v0lo);
  float* xp = reinterpret_cast<float*>(smem + p.off_x);
  float* vinp = reinterpret_cast<float*>(smem + p.off_vin);
  float* vinlo = reinterpret_cast<float*>(smem + p.off_vinlo);
  float* nvp = reinterpret_cast<float*>(smem + p.off_nv);
  float* vbs = reinterpret_cast<float*>(smem + p.off_vb);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);
  float* misc = reinterpret_cast<float*>(smem + p.off_misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;
  const int ldw = p.ldw, KH = p.KH, KR = p.KR, nmt = p.nmt, nstg = p.nstg;
  const int B = p.B, V = p.V, H = p.H;
  const int row0 = cta * p.rows_per_cta;
  const int rows = max(0, min(p.rows_per_cta, V - row0));
  const int rows8 = (rows + 7) & ~7;
  const int nrq = (rows8 + 31) >> 5;
  const int tail8 = rows8 - 32 * (nrq - 1);            // rows (multiple of 8) in the last 32-row quarter
  const int nch = KH >> 5;
  const int q = tid % p.GW, g_ = tid / p.GW;
  const bool col_ok = g_ < p.G && q < p.CQ;
  unsigned long long bar_target = 0;

  // barrier addresses
  const uint32_t b_full = smem_u32(bars), b_ready = b_full + 8 * MAX_STG, b_empty = b_ready + 8 * MAX_STG;
  const uint32_t b_lofree = b_empty + 8 * MAX_STG;
  const uint32_t b_accfull = b_lofree + 8 * NLO, b_dfull = b_accfull + 8, b_dfree = b_dfull + 16;
  const uint32_t b_stats = b_dfree + 16;

  int dbg_i = 0;
  auto mark = [&]() {
    if (p.dbg && cta == 0 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[dbg_i++] = t;
    }
  };
  auto stamp = [&](int role, uint32_t n) {      // per-stage role timeline of the first 16 stages (debug)
    if (p.dbg && cta == 0 && n < 16) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[32 + role * 16 + n] = t;
    }
  };
  mark();

  if (tid == 0) {
    for (int i = 0; i < MAX_STG; ++i) {
      mbar_init(b_full + 8 * i, 1);
      mbar_init(b_ready + 8 * i, N_TRANSFORM_WARPS);
      mbar_init(b_empty + 8 * i, 1);
    }
    for (int i = 0; i < NLO; ++i) mbar_init(b_lofree + 8 * i, 1);
    mbar_init(b_accfull, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(b_dfull + 8 * i, 1); mbar_init(b_dfree + 8 * i, 4); }
    for (int i = 0; i < MAX_SBAR; ++i) mbar_init(b_stats + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // ---- gather: v0 panel (+ lo, + rounded for the pseudo-likelihood), zero the other row panels ----
  int ex0 = 1, ex1 = 1;
  for (int e = tid; e < NB * KR; e += NT) {
    const int b = e / KR, r = e - b * KR;
    float x = 0.f;
    if (b < B && r < rows) {
      const long long dr = p.idx ? p.idx[b] : b;
      x = p.data[dr * p.ld_data + row0 + r];
    }
    const int o = poff(b, r, KR);
    const float xr = p.pcd ? roundf(x) : 0.f;      // src/rbm.py:428
    v0p[o] = x;
    v0lo[o] = tf32_lo(x);
    xp[o] = xr;
    vinp[o] = 0.f;
    vinlo[o] = 0.f;
    nvp[o] = 0.f;
    ex0 &= tf32_exact(x) ? 1 : 0;
    ex1 &= tf32_exact(xr) ? 1 : 0;
  }
  for (int r = tid; r < KR; r += NT) vbs[r] = r < rows ? p.vb[row0 + r] : 0.f;
  for (int e = tid; e < NB * KH; e += NT) { hp[e] = 0.f; hplo[e] = 0.f; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  const bool v0_exact = __syncthreads_and(ex0) != 0;
  (void)__syncthreads_and(ex1);     // rounded inputs beyond +-2048 keep only their tf32 part in the monitor
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // TMEM columns.  Per 128-column m-tile of the propup: [0,48) = A_hi x {src_lo, src, x}, [48,80) = A_lo x {src, x};
  // per propdown accumulator: [0,32) = A_hi x {h_lo, h}, [32,48) = A_lo x h.
  constexpr int UC = 80, DC = 48;
  const uint32_t t_d = tmem + nmt * UC;
  mark();   // gather done

  // ring / pipeline state (each role keeps its own counters; the schedules are deterministic)
  uint32_t stage_n = 0;              // stages processed so far by this role
  uint32_t ph_acc = 0;               // accfull parity (epilogue warps)
  uint32_t n_d[2] = {0, 0};          // uses of dfull[i] / dfree[i]
  const uint32_t ring = smem_u32(smem);

  auto seg = [&](long long off, uint32_t ordinal) {
    RngSeg s;
    s.mode = p.rng_mode;
    s.seg = p.ubuf ? p.ubuf + off : nullptr;
    s.k0 = p.k0; s.k1 = p.k1; s.c1 = ordinal; s.c2 = p.c2; s.c3 = p.c3;
    return s;
  };

  // lo twin of one staged 16 KB tile (transform warps)
  auto transform_stage = [&](uint32_t n) {
    const int slot = n % nstg, ls = n % NLO;
    mbar_wait(b_full + 8 * slot, (n / nstg) & 1);
    if (tid == NT - 32 * N_TRANSFORM_WARPS) stamp(1, n);
    mbar_wait(b_lofree + 8 * ls, ((n / NLO) & 1) ^ 1);
    if (tid == NT - 32 * N_TRANSFORM_WARPS) stamp(2, n);
    const float4* hi = reinterpret_cast<const float4*>(smem + (size_t)slot * TILE);
    float4* lo = reinterpret_cast<float4*>(smem + (size_t)(nstg + ls) * TILE);
    for (int e = tid - (NT - 32 * N_TRANSFORM_WARPS); e < TILE / 16; e += 32 * N_TRANSFORM_WARPS) {
      const float4 x = hi[e];
      lo[e] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(b_ready + 8 * slot);
    if (tid == NT - 32 * N_TRANSFORM_WARPS) stamp(3, n);
  };

  // =========================================================================================
  // propup pass: acc[mt][j, b] (+)= sum_i W[i][j] * src[b][i] over this CTA's rows
  // =========================================================================================
  auto u_pass = [&](const float* src_hi, const float* src_lo, bool src_exact, bool dual) {
    if (rows > 0) {
      if (warp == 0) {
        if (lane == 0) {
          for (int rq = 0; rq < nrq; ++rq) {
            const int nr8 = (rq == nrq - 1) ? tail8 : 32;
            for (int mt = 0; mt < nmt; ++mt, ++stage_n) {
              const int slot = stage_n % nstg;
              mbar_wait(b_empty + 8 * slot, ((stage_n / nstg) & 1) ^ 1);
              const uint32_t dst = ring + slot * TILE, bar = b_full + 8 * slot;
              stamp(0, stage_n);
              mbar_expect_tx(bar, 4u * nr8 * 128u);
              if (nr8 == 32) {
#pragma unroll
                for (int c = 0; c < 4; ++c) tma_load_2d(dst + c * 4096, &tmMN32, bar, mt * 128 + 32 * c, row0 + rq * 32);
              } else {
                for (int c = 0; c < 4; ++c)
                  for (int t = 0; t < nr8 / 8; ++t)
                    tma_load_2d(dst + c * 4096 + t * 1024, &tmMN8, bar, mt * 128 + 32 * c, row0 + rq * 32 + 8 * t);
              }
            }
          }
        }
      } else if (warp == 1) {
        if (lane == 0) {
          // B operands: stacked panels [src_lo | src | x]; src_exact drops the lo panel, dual adds x
          const uint32_t s_b1 = smem_u32(src_exact ? src_hi : src_lo), s_b2 = smem_u32(src_hi);
          const int n1 = (src_exact ? 16 : 32) + (dual ? 16 : 0), n2 = dual ? 32 : 16;
          const int c1 = src_exact ? 16 : 0;
          const uint32_t id1 = make_idesc(true, n1), id2 = make_idesc(true, n2);
          const uint64_t db1 = make_desc(s_b1, 128, KR * 32, 0), db2 = make_desc(s_b2, 128, KR * 32, 0);
          for (int rq = 0; rq < nrq; ++rq) {
            const int nkk = ((rq == nrq - 1) ? tail8 : 32) >> 3;
            for (int mt = 0; mt < nmt; ++mt, ++stage_n) {
              const int slot = stage_n % nstg;
              const uint32_t a0 = ring + slot * TILE, l0 = ring + (nstg + stage_n % NLO) * TILE;
              // the hi products only need the TMA tile: they run while the transform warps make the lo twin
              mbar_wait(b_full + 8 * slot, (stage_n / nstg) & 1);
              stamp(4, stage_n);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              // descriptors differ only in their start-address field: +64 (1024 B >> 4) per k-step on the
              // W tile, +16 (2 core matrices of 128 B) on the panels
              {
                const uint64_t ad = make_desc(a0, 4096, 512, 1);
                const uint64_t bd = db1 + (uint64_t)(rq * 64);
                const uint32_t acc0 = rq == 0 ? 0u : 1u;
                for (int kk = 0; kk < nkk; ++kk)
                  umma_tf32(tmem + mt * UC + c1, ad + (uint64_t)(kk * 64), bd + (uint64_t)(kk * 16), id1, kk == 0 ? acc0 : 1u);
              }
              mbar_wait(b_ready + 8 * slot, (stage_n / nstg) & 1);
              stamp(5, stage_n);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              {
                const uint64_t ad = make_desc(l0, 4096, 512, 1);
                const uint64_t bd = db2 + (uint64_t)(rq * 64);
                const uint32_t acc0 = rq == 0 ? 0u : 1u;
                for (int kk = 0; kk < nkk; ++kk)
                  umma_tf32(tmem + mt * UC + 48, ad + (uint64_t)(kk * 64), bd + (uint64_t)(kk * 16), id2, kk == 0 ? acc0 : 1u);
              }
              umma_commit(b_empty + 8 * slot);
              umma_commit(b_lofree + 8 * (stage_n % NLO));
            }
          }
          umma_commit(b_accfull);
        }
      } else if (warp >= NWARP - N_TRANSFORM_WARPS) {
        for (int rq = 0; rq < nrq; ++rq)
          for (int mt = 0; mt < nmt; ++mt, ++stage_n) transform_stage(stage_n);
      } else {
        // epilogue warps 2..5: accumulators -> this CTA's partial [B][ldw] in global scratch
        mbar_wait(b_accfull, ph_acc);
        ph_acc ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;
        for (int mt = 0; mt < nmt; ++mt) {
          const uint32_t tb = tmem + ((uint32_t)(quarter * 32) << 16) + mt * UC;
          uint32_t r0[16], r1[16], r2[16];
          tmem_ld16(tb + 16, r0);                     // A_hi x src
          tmem_ld16(tb + 48, r1);                     // A_lo x src
          if (!src_exact) tmem_ld16(tb, r2);          // A_hi x src_lo
          const int j = mt * 128 + quarter * 32 + lane;
          float* dst = p.part + ((size_t)cta * 2) * B * ldw;
          if (j < ldw) {
#pragma unroll
            for (int b = 0; b < NB; ++b)
              if (b < B) {
                float v = __uint_as_float(r1[b]);
                if (!src_exact) v += __uint_as_float(r2[b]);
                __stcg(dst + (size_t)b * ldw + j, __uint_as_float(r0[b]) + v);
              }
          }
          if (dual) {
            tmem_ld16(tb + 32, r0);                   // A_hi x x
            tmem_ld16(tb + 64, r1);                   // A_lo x x
            dst += (size_t)B * ldw;
            if (j < ldw) {
#pragma unroll
              for (int b = 0; b < NB; ++b)
                if (b < B) __stcg(dst + (size_t)b * ldw + j, __uint_as_float(r0[b]) + __uint_as_float(r1[b]));
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
    }
    __syncthreads();
  };

  // =========================================================================================
  // propdown pass: v[b, i] = act( sum_j W[i][j] h[b][j] + vb[i] ) for this CTA's rows, into the vin panels
  // =========================================================================================
  float cost_acc = 0.f;
  // tensor groups of 128 rows; a remainder of <= 32 rows after at least one full group goes to SIMT
  const int ng_full = rows8 >> 7, rem_rows = rows8 - 128 * ng_full;
  const bool simt_tail = ng_full > 0 && rem_rows > 0 && rem_rows <= 32;
  const int ngt = ng_full + ((rem_rows > 0 && !simt_tail) ? 1 : 0);
  const int tail_r0 = 128 * ng_full, tail_n = simt_tail ? max(0, rows - tail_r0) : 0;
  auto d_pass = [&](bool h_exact, const RngSeg& rs_v, bool last, const float* hsrc, int hld) {
    if (rows > 0) {
      if (warp == 0) {
        if (lane == 0) {
          for (int g = 0; g < ngt; ++g) {
            const int gr8 = min(128, rows8 - 128 * g), n32 = gr8 >> 5, t8 = (gr8 & 31) >> 3;
            for (int c = 0; c < nch; ++c, ++stage_n) {
              const int slot = stage_n % nstg;
              mbar_wait(b_empty + 8 * slot, ((stage_n / nstg) & 1) ^ 1);
              const uint32_t dst = ring + slot * TILE, bar = b_full + 8 * slot;
              mbar_expect_tx(bar, (uint32_t)gr8 * 128u);
              for (int t = 0; t < n32; ++t) tma_load_2d(dst + t * 4096, &tmK32, bar, 32 * c, row0 + 128 * g + 32 * t);
              for (int u = 0; u < t8; ++u)
                tma_load_2d(dst + n32 * 4096 + u * 1024, &tmK8, bar, 32 * c, row0 + 128 * g + 32 * n32 + 8 * u);
            }
          }
        }
      } else if (warp == 1) {
        if (lane == 0) {
          const uint32_t s_b1 = smem_u32(h_exact ? hp : hplo), s_b2 = smem_u32(hp);
          const int c1 = h_exact ? 16 : 0;
          const uint32_t id1 = make_idesc(false, h_exact ? 16 : 32), id2 = make_idesc(false, 16);
          const uint64_t db1 = make_desc(s_b1, 128, KH * 32, 0), db2 = make_desc(s_b2, 128, KH * 32, 0);
          for (int g = 0; g < ngt; ++g) {
            const int di = g & 1;
            if (n_d[di] > 0) mbar_wait(b_dfree + 8 * di, (n_d[di] - 1) & 1);    // epilogue drained the previous use
            for (int c = 0; c < nch; ++c, ++stage_n) {
              const int slot = stage_n % nstg;
              const uint32_t a0 = ring + slot * TILE, l0 = ring + (nstg + stage_n % NLO) * TILE;
              mbar_wait(b_full + 8 * slot, (stage_n / nstg) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              {
                const uint64_t ad = make_desc(a0, 16, 1024, 2);       // +2 (32 B >> 4) per k-step along the swizzled row
                const uint64_t bd = db1 + (uint64_t)(c * 64);
                const uint32_t acc0 = c == 0 ? 0u : 1u;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_tf32(t_d + di * DC + c1, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 16), id1, kk == 0 ? acc0 : 1u);
              }
              mbar_wait(b_ready + 8 * slot, (stage_n / nstg) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              {
                const uint64_t ad = make_desc(l0, 16, 1024, 2);
                const uint64_t bd = db2 + (uint64_t)(c * 64);
                const uint32_t acc0 = c == 0 ? 0u : 1u;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_tf32(t_d + di * DC + 32, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 16), id2, kk == 0 ? acc0 : 1u);
              }
              umma_commit(b_empty + 8 * slot);
              umma_commit(b_lofree + 8 * (stage_n % NLO));
            }
            umma_commit(b_dfull + 8 * di);
            n_d[di]++;
          }
        }
      } else if (warp >= NWARP - N_TRANSFORM_WARPS) {
        for (int g = 0; g < ngt; ++g)
          for (int c = 0; c < nch; ++c, ++stage_n) transform_stage(stage_n);
      } else {
        // epilogue warps: bias, activation, sampling (src/rbm.py:226-240 / :650-660) -> vin panels
        const int quarter = warp & 3;
        auto emit = [&](int r, int b, float sum) {
          float vin = 0.f, mean = 0.f;
          if (b < B) {
            const float pre = sum + vbs[r];
            if (p.kind == MDBN_GRBM) {
              mean = pre;
              vin = pre;          // mean-field visible: h given v_MEAN (src/rbm.py:669)
            } else {
              mean = sigmoidf_(pre);
              vin = rng_uniform(rs_v, (long long)b * V + row0 + r) < mean ? 1.f : 0.f;
            }
            if (last && !p.pcd) {
              const float t = v0p[poff(b, r, KR)];
              if (p.kind == MDBN_GRBM) { const float d = sigmoidf_(pre) - t; cost_acc += d * d; }   // :697
              else cost_acc += t * softplusf_(-pre) + (1.f - t) * softplusf_(pre);                  // :479-480
            }
          }
          const int o = poff(b, r, KR);
          vinp[o] = vin;
          vinlo[o] = tf32_lo(vin);
          if (last) nvp[o] = mean;
        };
        // the few rows beyond the last full 128-row group: plain fp32 dot products, done while the
        // tensor pipeline streams the full groups (these warps would otherwise wait)
        for (int rr = warp - 2; rr < tail_n; rr += 4) {
          const int r = tail_r0 + rr;
          const float* wrow = p.W + (size_t)(row0 + r) * ldw;
          float w[16];                               // H <= 512: this lane's columns lane, lane+32, ...
#pragma unroll
          for (int t = 0; t < 16; ++t) w[t] = (lane + 32 * t < H) ? __ldg(wrow + lane + 32 * t) : 0.f;
          float mine = 0.f;
          for (int b = 0; b < B; ++b) {
            const float* hrow = hsrc + (size_t)b * hld;
            float hv[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) hv[t] = (lane + 32 * t < H) ? __ldcg(hrow + lane + 32 * t) : 0.f;
            float a = 0.f;
#pragma unroll
            for (int t = 0; t < 16; ++t) a = fmaf(hv[t], w[t], a);
            a = warp_sum(a);
            if (lane == b) mine = a;
          }
          if (lane < NB) emit(r, lane, mine);
        }
        for (int g = 0; g < ngt; ++g) {
          const int di = g & 1;
          mbar_wait(b_dfull + 8 * di, n_d[di] & 1);
          n_d[di]++;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tb = tmem + ((uint32_t)(quarter * 32) << 16) + nmt * UC + di * DC;
          uint32_t r0[16], r1[16], r2[16];
          tmem_ld16(tb + 16, r0);                     // A_hi x h
          tmem_ld16(tb + 32, r1);                     // A_lo x h
          if (!h_exact) tmem_ld16(tb, r2);            // A_hi x h_lo
          const int r = 128 * g + quarter * 32 + lane;
          if (r < rows) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              float v = __uint_as_float(r1[b]);
              if (!h_exact) v += __uint_as_float(r2[b]);
              emit(r, b, __uint_as_float(r0[b]) + v);
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(b_dfree + 8 * di);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
    }
    __syncthreads();
  };

  // ---- distributed reduction of the hidden pre-activations + epilogue (as skinny.cu) ----
  auto reduce_hidden = [&](int nsets, float* mean_out, const RngSeg& rs, bool write_hs, bool write_p) {
    const int ldw4 = ldw >> 2;
    const int nq = nsets * B * p.CQ;
    const int per = (nq + gridDim.x - 1) / gridDim.x;
    const int o0 = cta * per, o1 = min(nq, o0 + per);
    for (int o = o0 + warp; o < o1; o += NWARP) {
      const int set = o / (B * p.CQ), rem = o % (B * p.CQ);
      const int b = rem / p.CQ, qq = rem % p.CQ;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 t8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = lane + 32 * u;
        t8[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < p.n_active)
          t8[u] = __ldcg(reinterpret_cast<const float4*>(p.part + (((size_t)c * 2 + set) * B + b) * ldw) + qq);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += t8[u].x; s.y += t8[u].y; s.z += t8[u].z; s.w += t8[u].w; }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, off);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, off);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, off);
        s.w += __shfl_xor_sync(0xffffffffu, s.w, off);
      }
      if (lane < 4) {
        const float sv = lane == 0 ? s.x : lane == 1 ? s.y : lane == 2 ? s.z : s.w;
        const int j = qq * 4 + lane;
        if (j < H) {
          const float pre = sv + p.hb[j];
          if (set == 1) {
            __stcg(&p.PREX[b * ldw + j], pre);
          } else {
            const float mean = sigmoidf_(pre);
            float smp = 0.f;
            if (write_hs || write_p) smp = rng_uniform(rs, (long long)b * H + j) < mean ? 1.f : 0.f;
            __stcg(&mean_out[b * ldw + j], mean);
            if (write_hs) __stcg(&p.HS[b * ldw + j], smp);
            if (write_p) __stcg(&p.P[(size_t)b * H + j], smp);
          }
        }
      }
    }
  };

  // chain state [B][H] -> h panels (hi, lo); returns whether it is exact in tf32 ({0,1} samples are)
  auto load_h = [&](const float* src, int ld_src) {
    int pred = 1;
    for (int e = tid; e < B * (KH >> 2); e += NT) {
      const int b = e / (KH >> 2), j = (e - b * (KH >> 2)) * 4;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (j + t < H) x[t] = __ldcg(src + (size_t)b * ld_src + j + t);
      const int o = poff(b, j, KH);
      *reinterpret_cast<float4*>(hp + o) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(hplo + o) = make_float4(tf32_lo(x[0]), tf32_lo(x[1]), tf32_lo(x[2]), tf32_lo(x[3]));
      pred &= (tf32_exact(x[0]) && tf32_exact(x[1]) && tf32_exact(x[2]) && tf32_exact(x[3])) ? 1 : 0;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    return __syncthreads_and(pred) != 0;
  };

  // =============================== pass 0: positive phase ===============================
  u_pass(v0p, v0lo, v0_exact, p.pcd != 0);
  mark();   // pass-0 done (partials written)
  mark();
  grid_sync(p.bar, bar_target);
  mark();
  reduce_hidden(p.pcd ? 2 : 1, p.PH, seg(0, 0), !p.pcd, false);      // CD: chain starts from the fresh sample
  mark();
  grid_sync(p.bar, bar_target);
  bool h_exact = p.pcd ? load_h(p.P, H) : load_h(p.HS, ldw);          // PCD: from the persistent chain (:308-311)
  mark();

  // pseudo-likelihood monitor (src/rbm.py:421-447) — CTA 0, pre-update W, hb, vb
  if (p.pcd && cta == 0) {
    const int bit = *p.bit_idx;
    for (int b = warp; b < B; b += NWARP) {
      const long long dr = p.idx ? p.idx[b] : b;
      const float x = roundf(p.data[dr * p.ld_data + bit]);
      const float d = 1.f - 2.f * x;
      float h0 = 0.f, h1 = 0.f;
      for (int j = lane; j < H; j += 32) {
        const float pre = __ldcg(&p.PREX[b * ldw + j]);
        h0 += softplusf_(pre);
        h1 += softplusf_(pre + d * p.W[(size_t)bit * ldw + j]);
      }
      h0 = warp_sum(h0);
      h1 = warp_sum(h1);
      if (lane == 0) {
        const float vbv = p.vb[bit];
        float vterm;
        if (p.kind == MDBN_GRBM) { const float a = x - vbv, c = (1.f - x) - vbv; vterm = 0.5f * (a * a - c * c); }
        else vterm = d * vbv;
        misc[32 + b] = -(float)V * softplusf_((h1 - h0) + vterm);
      }
    }
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += misc[32 + b];
      misc[63] = s * p.cost_scale;
    }
    __syncthreads();
  }

  // =============================== Gibbs steps ===============================
  for (int s = 0; s < p.k; ++s) {
    const bool last = (s == p.k - 1);
    const long long ubase = (long long)B * H + (long long)s * p.u_step_stride;
    const RngSeg rs_v = seg(ubase + p.u_off_v, 1u + 2u * s);
    const RngSeg rs_h = seg(ubase + p.u_off_h, 2u + 2u * s);
    d_pass(h_exact, rs_v, last, (s == 0 && p.pcd) ? p.P : p.HS, (s == 0 && p.pcd) ? H : ldw);
    u_pass(vinp, vinlo, p.kind == MDBN_RBM, false);
    if (last) mark();
    if (last && !p.pcd) {
      const float c = block_sum(cost_acc, misc);
      if (tid == 0) __stcg(&p.cost_part[cta], c);
    }
    grid_sync(p.bar, bar_target);
    if (last) mark();
    reduce_hidden(1, p.NH, rs_h, !last, last && p.pcd);
    grid_sync(p.bar, bar_target);
    if (last) mark();
    if (!last) h_exact = load_h(p.HS, ldw);
  }

  // =============================== statistics + update (SIMT) ===============================
  {
    const int narr = p.wc != 0.f ? 3 : 2;
    const int TRS = 8, slot_s = (TRS * ldw * 4 + 127) & ~127;
    int depth = ((nstg + NLO) * TILE) / (narr * slot_s);
    depth = depth > MAX_SBAR ? MAX_SBAR : depth;
    const int ntiles_s = (rows + TRS - 1) / TRS;
    uint32_t sphase = 0;
    auto issue = [&](int j) {       // thread 0
      const int r0 = j * TRS;
      if (r0 >= rows) return;
      const int st = j % depth, nr = min(TRS, rows - r0);
      const uint32_t bytes = (uint32_t)nr * ldw * 4u, bar = b_stats + 8 * st;
      mbar_expect_tx(bar, bytes * narr);
      const uint32_t dst = ring + st * narr * slot_s;
      const size_t goff = (size_t)(row0 + r0) * ldw;
      bulk_g2s(dst, p.W + goff, bytes, bar);
      bulk_g2s(dst + slot_s, p.S + goff, bytes, bar);
      if (narr > 2) bulk_g2s(dst + 2 * slot_s, p.Wsnap + goff, bytes, bar);
    };
    if (tid == 0) for (int j = 0; j < depth; ++j) issue(j);
    const int ldw4 = ldw >> 2;
    float4 ph[BT], nh[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      ph[b] = nh[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col_ok && b < B) {
        ph[b] = __ldcg(reinterpret_cast<const float4*>(p.PH) + b * ldw4 + q);
        nh[b] = __ldcg(reinterpret_cast<const float4*>(p.NH) + b * ldw4 + q);
      }
    }
    const int ncol = min(4, H - 4 * q);
    for (int j = 0, stg = 0; j < ntiles_s; ++j, stg = (stg + 1 == depth ? 0 : stg + 1)) {
      mbar_wait(b_stats + 8 * stg, (sphase >> stg) & 1u);
      sphase ^= (1u << stg);
      const unsigned char* sb = smem + (size_t)stg * narr * slot_s;
      const float4* wt = reinterpret_cast<const float4*>(sb);
      const float4* st4 = reinterpret_cast<const float4*>(sb + slot_s);
      const float4* sn4 = reinterpret_cast<const float4*>(sb + 2 * (size_t)slot_s);
      const int nr = min(TRS, rows - j * TRS);
      if (col_ok) {
        for (int r = g_; r < nr; r += p.G) {
          const int lr_ = j * TRS + r;
          const float4 w = wt[r * ldw4 + q], sp = st4[r * ldw4 + q];
          float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            const int o = poff(b, lr_, KR);
            const float a = v0p[o], n = nvp[o];
            gs.x = fmaf(a, ph[b].x, gs.x); gs.x = fmaf(-n, nh[b].x, gs.x);
            gs.y = fmaf(a, ph[b].y, gs.y); gs.y = fmaf(-n, nh[b].y, gs.y);
            gs.z = fmaf(a, ph[b].z, gs.z); gs.z = fmaf(-n, nh[b].z, gs.z);
            gs.w = fmaf(a, ph[b].w, gs.w); gs.w = fmaf(-n, nh[b].w, gs.w);
          }
          const float wv[4] = {w.x, w.y, w.z, w.w}, sv[4] = {sp.x, sp.y, sp.z, sp.w}, gv[4] = {gs.x, gs.y, gs.z, gs.w};
          float snv[4] = {0.f, 0.f, 0.f, 0.f};
          if (narr > 2) { const float4 t4 = sn4[r * ldw4 + q]; snv[0] = t4.x; snv[1] = t4.y; snv[2] = t4.z; snv[3] = t4.w; }
          float wo[4], so[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float gw = gv[c] * p.inv_bnom - p.wc * snv[c];                // src/rbm.py:411-415
            float mult = p.decay;
            if (p.c1 != 0.f) {
              const float invD = __fdividef(1.0f, 1.0f + p.c1 * __fdividef(1.0f, fabsf(wv[c]) + 0.001f));   // :347-350
              gw *= invD;
              mult *= invD;                                               // :353-356
            }
            so[c] = gw + (sv[c] - gw) * p.mom;                            // :361
            wo[c] = wv[c] * mult + sv[c] * p.lr;                          // :364 (OLD speed)
          }
          const size_t go = (size_t)(row0 + lr_) * ldw + 4 * q;
          if (ncol == 4) {
            *reinterpret_cast<float4*>(p.W + go) = make_float4(wo[0], wo[1], wo[2], wo[3]);
            *reinterpret_cast<float4*>(p.S + go) = make_float4(so[0], so[1], so[2], so[3]);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < ncol) { p.W[go + c] = wo[c]; p.S[go + c] = so[c]; }
          }
        }
      }
      __syncthreads();
      if (tid == 0) issue(j + depth);
    }
    for (int r = tid; r < rows; r += NT) {           // visible bias, src/rbm.py:417
      float gsum = 0.f;
      for (int b = 0; b < B; ++b) { const int o = poff(b, r, KR); gsum += v0p[o] - nvp[o]; }
      const float gb = gsum * p.inv_b, sv = p.Svb[row0 + r];
      p.Svb[row0 + r] = gb + (sv - gb) * p.mom;
      p.vb[row0 + r] = p.vb[row0 + r] + sv * p.lr;
    }
    if (cta == gridDim.x - 1) {                      // hidden bias, src/rbm.py:416
      for (int j = tid; j < H; j += NT) {
        float gsum = 0.f;
        for (int b = 0; b < B; ++b) gsum += __ldcg(&p.PH[b * ldw + j]) - __ldcg(&p.NH[b * ldw + j]);
        const float gb = gsum * p.inv_b, sv = p.Shb[j];
        p.Shb[j] = gb + (sv - gb) * p.mom;
        p.hb[j] = p.hb[j] + sv * p.lr;
      }
    }
    if (cta == 0 && tid == 0) {
      float c;
      if (p.pcd) {
        c = misc[63];
        *p.bit_idx = (*p.bit_idx + 1) % V;                                 // :445
      } else {
        c = 0.f;
        for (int i = 0; i < p.n_active; ++i) c += __ldcg(&p.cost_part[i]);
        c *= p.cost_scale;
      }
      if (p.cost_out) *p.cost_out = c;
    }
  }
  mark();

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
  if (tid == 0) {
    __threadfence();
    const unsigned long long prev = atomicAdd(p.bar + 1, 1ULL);
    if (prev == gridDim.x - 1) {
      p.bar[0] = 0ULL;
      p.bar[1] = 0ULL;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------------
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
static int make_map(CUtensorMap* tm, const float* W, int V, int ldw, int box_rows, bool mn) {
  cuuint64_t dims[2] = {(cuuint64_t)ldw, (cuuint64_t)V};
  cuuint64_t strides[1] = {(cuuint64_t)ldw * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)W, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE,
                            mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MDBN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

struct Geometry {
  int BT, rows_per_cta, n_active, CQ, GW, G, nstg, KH, KR, nmt, grid;
  int off_hp, off_hplo, off_v0, off_v0lo, off_x, off_vin, off_vinlo, off_nv, off_vb, off_bars, off_misc;
  size_t smem;
  bool ok;
};

static Geometry plan(const mdbn_ctx* c, const mdbn_cd_args& a) {
  Geometry g{};
  g.ok = false;
  if (a.B > NB || a.phase != MDBN_PHASE_FULL) return g;
  g.BT = a.B <= 10 ? 10 : 16;
  const int ncols8 = (a.H + 7) & ~7;
  if (a.ldw != ncols8 || ncols8 > 512) return g;
  if (((uintptr_t)a.W | (uintptr_t)a.W_speed | (uintptr_t)a.W_snap) & 15) return g;
  if (!get_encode()) return g;
  // only worth it when streaming W dominates: small layers stay on the mma.sync kernel
  if ((long long)a.V * a.ldw < 512 * 1024) return g;
  g.CQ = a.ldw / 4;
  g.GW = (g.CQ + 31) / 32 * 32;
  if (g.GW > NT) return g;
  g.G = NT / g.GW;
  g.grid = c->num_sms;
  g.rows_per_cta = (((a.V + g.grid - 1) / g.grid) + 7) & ~7;
  g.n_active = (a.V + g.rows_per_cta - 1) / g.rows_per_cta;
  g.KR = (g.rows_per_cta + 31) & ~31;
  g.KH = (ncols8 + 31) & ~31;
  g.nmt = (ncols8 + 127) / 128;
  auto up128 = [](size_t x) { return (x + 127) & ~(size_t)127; };
  const size_t hp_b = (size_t)NB * g.KH * 4, rp_b = (size_t)NB * g.KR * 4, vb_b = up128((size_t)g.KR * 4);   // KH, KR % 32 == 0
  const size_t fixed = 2 * hp_b + 6 * rp_b + vb_b + 512 + 512;
  const size_t smem_max = 227 * 1024;
  if (fixed + (size_t)(2 + NLO) * TILE > smem_max) return g;
  g.nstg = (int)((smem_max - fixed) / TILE) - NLO;
  if (g.nstg > MAX_STG) g.nstg = MAX_STG;
  // the statistics pass reuses the ring: W and S (and W_snap) tiles of 8 rows must fit at least once
  const int narr = a.weightcost != 0.f ? 3 : 2;
  if ((size_t)narr * ((8 * a.ldw * 4 + 127) & ~127) > (size_t)(g.nstg + NLO) * TILE) return g;
  size_t off = (size_t)(g.nstg + NLO) * TILE;
  auto take = [&](size_t bytes) { size_t o = off; off += bytes; return (int)o; };
  // stacked-N operands: [lo | hi | x] panels are contiguous so one MMA covers several of them
  g.off_hplo = take(hp_b); g.off_hp = take(hp_b);
  g.off_v0lo = take(rp_b); g.off_v0 = take(rp_b); g.off_x = take(rp_b);
  g.off_vinlo = take(rp_b); g.off_vin = take(rp_b); g.off_nv = take(rp_b);
  g.off_vb = take(vb_b);
  g.off_bars = take(512);
  g.off_misc = take(512);
  g.smem = off;
  g.ok = g.smem <= smem_max;
  return g;
}

template <int BT>
static int launch(mdbn_ctx* c, const CUtensorMap* tms, const Params& p, const Geometry& g, cudaStream_t st) {
  static bool configured[64] = {};
  auto kfn = cd_skinny_tc_kernel<BT>;
  if (!configured[c->device]) {
    MDBN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[c->device] = true;
  }
  void* args[] = {(void*)&tms[0], (void*)&tms[1], (void*)&tms[2], (void*)&tms[3], (void*)&p};
  MDBN_CUDA(cudaLaunchCooperativeKernel((void*)kfn, dim3(g.grid), dim3(NT), args, g.smem, st));
  c->launches++;
  return 0;
}

}  // namespace sktc

bool skinny_tc_supported(const mdbn_ctx* c, const mdbn_cd_args& a) {
  // opt-in (MDBN_SKINNY_TC=1): fp32-exact via split TF32, but measured slower than the SIMT kernel — for a
  // 16-wide N every tcgen05.mma re-reads 4 KB of A from shared memory (~130 clk), see DESIGN.md
  static const bool on = getenv("MDBN_SKINNY_TC") != nullptr;
  return on && sktc::plan(c, a).ok;
}

int skinny_tc_cd_step(mdbn_ctx* c, const mdbn_cd_args& a, cudaStream_t st) {
  using namespace sktc;
  Geometry g = plan(c, a);
  MDBN_CHECK(g.ok, "skinny tcgen05 path: unsupported shape");
  Params p{};
  p.W = a.W; p.S = a.W_speed; p.Wsnap = a.weightcost != 0.f ? a.W_snap : nullptr; p.ldw = a.ldw;
  p.hb = a.hbias; p.vb = a.vbias; p.Shb = a.hbias_speed; p.Svb = a.vbias_speed;
  p.data = a.data; p.ld_data = a.ld_data; p.idx = a.indices;
  p.P = a.persistent; p.bit_idx = a.bit_i_idx; p.cost_out = a.cost_out;
  p.kind = a.kind; p.noisy = a.noisy; p.B = a.B; p.V = a.V; p.H = a.H; p.k = a.k; p.pcd = a.persistent != nullptr;
  p.inv_bnom = 1.0f / (float)a.B_nom;
  p.inv_b = 1.0f / (float)a.B;
  p.wc = a.weightcost;
  p.c1 = (2.0f * a.lr) * a.lambda_1;
  p.decay = 1.0f - (2.0f * a.lr) * a.lambda_2;
  p.mom = a.momentum; p.lr = a.lr;
  p.cost_scale = (!p.pcd && a.kind == MDBN_GRBM) ? 1.0f / ((float)a.B * (float)a.V) : 1.0f / (float)a.B;
  p.rng_mode = a.rng.mode;
  p.ubuf = a.rng.mode == MDBN_RNG_BUFFER ? a.rng.buffer : nullptr;
  p.k0 = (uint32_t)a.rng.seed; p.k1 = (uint32_t)(a.rng.seed >> 32);
  p.c2 = (uint32_t)a.rng.offset; p.c3 = (uint32_t)(a.rng.offset >> 32);
  ULayout ul = u_layout(a.kind, a.noisy, a.B, a.V, a.H);
  p.u_step_stride = ul.step_stride; p.u_off_v = ul.off_v; p.u_off_h = ul.off_h;
  p.rows_per_cta = g.rows_per_cta; p.n_active = g.n_active; p.CQ = g.CQ; p.GW = g.GW; p.G = g.G;
  p.nstg = g.nstg; p.KH = g.KH; p.KR = g.KR; p.nmt = g.nmt;
  p.off_hp = g.off_hp; p.off_hplo = g.off_hplo; p.off_v0 = g.off_v0; p.off_v0lo = g.off_v0lo; p.off_x = g.off_x;
  p.off_vin = g.off_vin; p.off_vinlo = g.off_vinlo; p.off_nv = g.off_nv; p.off_vb = g.off_vb;
  p.off_bars = g.off_bars; p.off_misc = g.off_misc;

  const size_t hb_f = (size_t)NB * a.ldw;
  const size_t part_f = (size_t)g.n_active * 2 * (size_t)a.B * a.ldw;
  const size_t total_f = part_f + 4 * hb_f + (size_t)g.grid + 64;
  mdbn_ctx::Buf& wb = c->ws[WS_TENSOR];
  const void* before = wb.p;
  const size_t before_n = wb.n;
  float* base = (float*)ws_get(c, WS_TENSOR, total_f * sizeof(float));
  if (!base) return 3;
  static thread_local unsigned long long last_key = 0;
  const unsigned long long key = ((unsigned long long)a.B << 48) ^ ((unsigned long long)a.ldw << 24) ^
                                 (unsigned long long)g.n_active ^ ((unsigned long long)(uintptr_t)base << 1);
  if (before != wb.p || before_n != wb.n || key != last_key) {
    MDBN_CUDA(cudaMemsetAsync(base, 0, wb.n, st));
    last_key = key;
  }
  p.part = base;
  p.PH = base + part_f;
  p.NH = p.PH + hb_f;
  p.HS = p.NH + hb_f;
  p.PREX = p.HS + hb_f;
  p.cost_part = p.PREX + hb_f;
  p.bar = reinterpret_cast<unsigned long long*>(c->barrier);
  static const bool want_timing = getenv("MDBN_SKINNY_TIMING") != nullptr;
  p.dbg = want_timing ? reinterpret_cast<unsigned long long*>(c->barrier) + 8 : nullptr;

  CUtensorMap tms[4];
  MDBN_TRY(make_map(&tms[0], a.W, a.V, a.ldw, 32, true));
  MDBN_TRY(make_map(&tms[1], a.W, a.V, a.ldw, 8, true));
  MDBN_TRY(make_map(&tms[2], a.W, a.V, a.ldw, 32, false));
  MDBN_TRY(make_map(&tms[3], a.W, a.V, a.ldw, 8, false));
  int rc = g.BT == 10 ? launch<10>(c, tms, p, g, st) : launch<16>(c, tms, p, g, st);
  if (rc == 0 && p.dbg) {
    unsigned long long t[32 + 6 * 16];
    MDBN_CUDA(cudaStreamSynchronize(st));
    MDBN_CUDA(cudaMemcpy(t, p.dbg, sizeof(t), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[skinny-tc timeline us] V=%d H=%d B=%d k=%d nstg=%d:", a.V, a.H, a.B, a.k, g.nstg);
    for (int i = 1; i < 13; ++i) fprintf(stderr, " %.1f", (double)(t[i] - t[0]) * 1e-3);
    fprintf(stderr, "\n");
    const char* names[6] = {"tma issue", "xform full", "xform lofree", "xform done", "mma full", "mma ready"};
    for (int r = 0; r < 6; ++r) {
      fprintf(stderr, "   %-12s", names[r]);
      for (int i = 0; i < 16; ++i) fprintf(stderr, " %6.2f", (double)(t[32 + r * 16 + i] - t[0]) * 1e-3);
      fprintf(stderr, "\n");
    }
  }
  return rc;
}

}  // namespace mdbn
