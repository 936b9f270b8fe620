// Skinny-batch CD-k / PCD-k step as ONE persistent cooperative kernel (B <= 20).
//
// Regime: batch 10-20 on 10^2..2*10^4-wide layers is 0.36*B flop/byte ->
// bound by streaming W, not by math.  Design:
//   * every CTA owns a contiguous slab of W rows (visible units) for the whole step;
//   * W row tiles are staged in shared memory by 1-D bulk async copies
//     (cp.async.bulk -> UBLKCP, one per tile: the TMA op rate, ~150 ns per copy per SM, rules
//     out per-row copies) completing on mbarriers;
//   * the two skinny GEMMs run on the tensor cores as mma.sync m16n8k8 TF32 with the fp32
//     operands split hi + lo (3 MMAs per product: hi*hi + lo*hi + hi*lo), which keeps fp32-level
//     accuracy (the 1e-5 parity bar) at a fraction of the issue slots of FFMA + shuffle
//     reductions (the tcgen05 shapes, M >= 64, do not fit a 10-20 row batch);
//   * pass 0      : partial  v0 W          (and round(v0) W for the pseudo-likelihood)
//   * pass 1..k   : FUSED propdown + propup from the SAME staged tile: v_i = h . W[i,:] is
//     complete inside the owning CTA (no cross-CTA traffic), its bias/sigmoid/Bernoulli
//     epilogue runs in place and the tile is immediately reused for  h' += v_i W[i,:];
//     so a Gibbs step reads W once, not twice;
//   * hidden pre-activations need all rows: per-CTA partials -> global scratch -> grid
//     barrier -> each CTA reduces a slice in fixed order (deterministic) + bias + sigmoid +
//     sample -> grid barrier -> every CTA reloads the full [B,H] hidden state;
//   * last pass  : statistics + lambda_1/lambda_2/momentum update fused: W and W_speed tiles
//     are read once and written once; v0 and nv slabs never left shared memory.
// HBM traffic per step: (k+1) reads of W + read W,S + write W,S (+ read W_snap) versus the
// (2k+1)+4 of an unfused implementation.
#include <stdlib.h>
#include "ctx.h"

namespace mdbn {
namespace sk {

constexpr int MAX_SLOTS = 6;
constexpr int MAX_TR = 32;
constexpr int MAX_NTD = MAX_TR / 8;

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
  // geometry
  int rows_per_cta, rows_alloc, n_active, CQ, GW, G, TR, nslots, ldp, ldh, slot_bytes;
  // global scratch
  float* part;        // [n_active][2][BT][ldw]
  float *PH, *NH, *HS, *PREX;   // [BT][ldw], zero-initialised, padded columns never written
  float* cost_part;   // [gridDim]
  unsigned long long* bar;   // [0] barrier counter, [1] exit counter
  unsigned long long* dbg;   // optional phase timeline (MDBN_SKINNY_TIMING=1), CTA 0 only
  // smem byte offsets
  int off_hs, off_v0, off_nv, off_vt, off_dred, off_bars, off_misc, off_vb;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Device-wide barrier.  All CTAs are co-resident (cooperative launch).  The counter is
// monotonic within a launch and reset by the last CTA to leave the kernel.
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

// ---- split-TF32 tensor-core helpers -------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  // hi = x truncated to tf32 (what the tensor core would do anyway), lo = exact remainder, itself
  // truncated by the hardware: |x - hi - tf32(lo)| <= 2^-21 |x|.  Bit masks, not cvt.rna: the
  // conversions run on the quarter-rate XU pipe and were the bottleneck of this loop.
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += A * B with A = a_hi + a_lo, B = b_hi + b_lo (lo*lo dropped: 2^-22 relative)
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                       uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(c, al, bh0, bh1);
  mma_tf32(c, ah, bl0, bl1);
  mma_tf32(c, ah, bh0, bh1);
}

template <int BT>
struct Cfg {
  static constexpr int BTP = (BT + 3) / 4 * 4;
  static constexpr int MT = BT > 16 ? 2 : 1;     // 16-row m-tiles of the batch
  static constexpr int MB = 16 * MT;             // batch rows seen by the MMAs (zero padded)
  static constexpr int BTS = MT == 1 ? 24 : 40;  // slab row stride in floats, = 8 or 24 mod 32
};

template <int BT, int NPW, int NT>
__global__ void __launch_bounds__(NT, 1) cd_skinny_kernel(const Params p) {
  using C = Cfg<BT>;
  constexpr int NWARP = NT / 32;
  constexpr int BTP = C::BTP, MT = C::MT, MB = C::MB, BTS = C::BTS;
  constexpr bool WIDE = NPW * NWARP * 8 > 512;   // H > 512: short tiles (TR <= 16)
  constexpr int NTD = WIDE ? 2 : MAX_NTD;
  constexpr bool ALLOW_DUAL = !(WIDE && MT == 2);   // register budget (plan() routes that case away)
  extern __shared__ __align__(1024) unsigned char smem[];
  float* hs = reinterpret_cast<float*>(smem + p.off_hs);      // [MB][ldh] chain state
  float* v0s = reinterpret_cast<float*>(smem + p.off_v0);     // [rows_alloc][BTS]
  float* nvs = reinterpret_cast<float*>(smem + p.off_nv);     // [rows_alloc][BTS]
  float* vt = reinterpret_cast<float*>(smem + p.off_vt);      // [TR][BTS] visible tile -> propup input
  float* dred = reinterpret_cast<float*>(smem + p.off_dred);  // [NWARP][MB][TR] propdown k-split partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);
  float* misc = reinterpret_cast<float*>(smem + p.off_misc);  // [64]: block_sum scratch, pl cost
  float* vbs = reinterpret_cast<float*>(smem + p.off_vb);     // [rows_alloc] visible bias of the owned rows

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lq = lane >> 2, lr = lane & 3;     // mma fragment coordinates
  const int cta = blockIdx.x;
  const int ldw = p.ldw, ldp = p.ldp, ldp4 = ldp >> 2, ldh = p.ldh;
  const int B = p.B, V = p.V, H = p.H;
  const int row0 = cta * p.rows_per_cta;
  const int rows = max(0, min(p.rows_per_cta, V - row0));
  const int ntiles = (rows + p.TR - 1) / p.TR;
  const int ncols8 = (H + 7) & ~7;
  // SIMT mapping of the statistics pass: thread -> (row group g, column quad q)
  const int q = tid % p.GW, g = tid / p.GW;
  const bool col_ok = g < p.G && q < p.CQ;
  unsigned long long bar_target = 0;
  uint32_t phase_bits = 0;
  int dbg_i = 0, dbg_j = 16;
  auto mark2 = [&]() {
    if (p.dbg && cta == 0 && tid == 0 && dbg_j < 32) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[dbg_j++] = t;
    }
  };
  auto mark = [&]() {
    if (p.dbg && cta == 0 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[dbg_i++] = t;
    }
  };
  mark();

  if (tid == 0) {
    for (int i = 0; i < MAX_SLOTS; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the ring once: the pad columns [ldw, ldp) of every staged row are never written by the
  // copies and are read (times zero) by the last 8-column fragment
  for (int e = tid; e < p.nslots * (p.slot_bytes >> 4); e += NT)
    reinterpret_cast<float4*>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = tid; e < MB * ldh; e += NT) hs[e] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  // ---- tile pipeline --------------------------------------------------------------
  // job j of a pass loads `narr` arrays (W [, S [, Wsnap]]) of tile j into consecutive slots, one
  // bulk copy per row (padded destination stride).  Called by all lanes of warp 0.
  auto issue = [&](int j, int narr, int depth, int tr, int slot_b) {      // lane 0 of warp 0
    const int r0 = j * tr;
    if (r0 >= rows || lane != 0) return;
    const int st = j % depth;
    const int nr = min(tr, rows - r0);
    const uint32_t bytes = (uint32_t)nr * ldw * 4u;       // rows are contiguous: ONE bulk copy per array
    uint64_t* bar = &bars[st];
    mbar_expect_tx(bar, bytes * narr);
    unsigned char* dst = smem + (size_t)st * narr * slot_b;
    const size_t goff = (size_t)(row0 + r0) * ldw;
    bulk_g2s(dst, p.W + goff, bytes, bar);
    if (narr > 1) bulk_g2s(dst + slot_b, p.S + goff, bytes, bar);
    if (narr > 2) bulk_g2s(dst + 2 * (size_t)slot_b, p.Wsnap + goff, bytes, bar);
  };
  auto wait_stage = [&](int st) {
    mbar_wait(&bars[st], (phase_bits >> st) & 1u);
    phase_bits ^= (1u << st);
  };

  // ---- randomness -------------------------------------------------------------
  auto seg = [&](long long off, uint32_t ordinal) {
    RngSeg s;
    s.mode = p.rng_mode;
    s.seg = p.ubuf ? p.ubuf + off : nullptr;
    s.k0 = p.k0; s.k1 = p.k1; s.c1 = ordinal; s.c2 = p.c2; s.c3 = p.c3;
    return s;
  };

  // ---- gather v0 slab: v0s[r][b] = data[idx[b]][row0 + r]; rows >= `rows` and b >= B are zero -----
  if (warp == 0) issue(0, 1, p.nslots, p.TR, p.slot_bytes);     // start streaming W while the minibatch is gathered
  for (int r = tid; r < p.rows_alloc; r += NT) vbs[r] = r < rows ? p.vb[row0 + r] : 0.f;
  int exact_pred = 1, exact_pred2 = 1;
  for (int e = tid; e < p.rows_alloc * MB; e += NT) {
    int b = e / p.rows_alloc, r = e % p.rows_alloc;
    float x = 0.f;
    if (b < B && r < rows) {
      long long dr = p.idx ? p.idx[b] : b;
      x = p.data[dr * p.ld_data + row0 + r];
    }
    const float xr = p.pcd ? roundf(x) : 0.f;     // src/rbm.py:428; the nv slab is free until the last Gibbs step
    v0s[r * BTS + b] = x;
    nvs[r * BTS + b] = xr;
    exact_pred &= ((__float_as_uint(x) & 0x1FFFu) == 0u) ? 1 : 0;       // representable in tf32?
    exact_pred2 &= ((__float_as_uint(xr) & 0x1FFFu) == 0u) ? 1 : 0;
  }
  const bool v0_exact = __syncthreads_and(exact_pred) != 0;    // binary / small-integer data: no low term
  const bool x_exact = __syncthreads_and(exact_pred2) != 0;
  mark();   // gather done

  // ---- propup of one staged tile on the tensor cores ------------------------------------
  // out[b, j] += sum_i src[i][b] * W[i][j]; warp w owns the 8-column n-tiles w, w+8, ...;
  // DUAL shares the W fragments with a second input slab.
  auto up_mma = [&](const float* __restrict__ tile, const float* __restrict__ src, const float* __restrict__ src2,
                    int nr8, float (&acc)[MT][NPW][4], float (&acc2)[MT][NPW][4], bool dual, bool a_exact,
                    bool a2_exact) {
    for (int i0 = 0; i0 < nr8; i0 += 8) {
      uint32_t ah[MT][4], al[MT][4], a2h[MT][4], a2l[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float* s0 = src + (i0 + lr) * BTS + mt * 16 + lq;
        split_tf32(s0[0], ah[mt][0], al[mt][0]);
        split_tf32(s0[8], ah[mt][1], al[mt][1]);
        split_tf32(s0[4 * BTS], ah[mt][2], al[mt][2]);
        split_tf32(s0[4 * BTS + 8], ah[mt][3], al[mt][3]);
        if (dual) {
          const float* t0 = src2 + (i0 + lr) * BTS + mt * 16 + lq;
          split_tf32(t0[0], a2h[mt][0], a2l[mt][0]);
          split_tf32(t0[8], a2h[mt][1], a2l[mt][1]);
          split_tf32(t0[4 * BTS], a2h[mt][2], a2l[mt][2]);
          split_tf32(t0[4 * BTS + 8], a2h[mt][3], a2l[mt][3]);
        }
      }
      // B fragments of all n-tiles first, then the three split terms term-major so that
      // back-to-back MMAs never hit the same accumulator (dependent-issue latency)
      uint32_t bh[NPW][2], bl[NPW][2];
#pragma unroll
      for (int nt = 0; nt < NPW; ++nt) {
        const int n0 = (warp + nt * NWARP) * 8;
        if (n0 < ncols8) {
          const float* w0 = tile + (i0 + lr) * ldp + n0 + lq;
          split_tf32(w0[0], bh[nt][0], bl[nt][0]);
          split_tf32(w0[4 * ldp], bh[nt][1], bl[nt][1]);
        }
      }
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        if (term == 2 && a_exact && (!dual || a2_exact)) continue;   // no low parts (binary / integer inputs)
#pragma unroll
        for (int nt = 0; nt < NPW; ++nt) {
          const int n0 = (warp + nt * NWARP) * 8;
          if (n0 < ncols8) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              if (term == 0) mma_tf32(acc[mt][nt], ah[mt], bh[nt][0], bh[nt][1]);
              else if (term == 1) mma_tf32(acc[mt][nt], ah[mt], bl[nt][0], bl[nt][1]);
              else if (!a_exact) mma_tf32(acc[mt][nt], al[mt], bh[nt][0], bh[nt][1]);
              if (dual) {     // second input: round(v0) is integer-valued -> exact in tf32, no low term
                if (term == 0) mma_tf32(acc2[mt][nt], a2h[mt], bh[nt][0], bh[nt][1]);
                else if (term == 1) mma_tf32(acc2[mt][nt], a2h[mt], bl[nt][0], bl[nt][1]);
                else if (!a2_exact) mma_tf32(acc2[mt][nt], a2l[mt], bh[nt][0], bh[nt][1]);
              }
            }
          }
        }
      }
    }
  };

  // ---- this CTA's partial [BT][ldw] straight from the accumulator fragments to global scratch ----
  auto flush_partial = [&](float (&acc)[MT][NPW][4], int set) {
    if (rows <= 0) return;
    float* dst = p.part + ((size_t)cta * 2 + set) * BT * ldw;
#pragma unroll
    for (int nt = 0; nt < NPW; ++nt) {
      const int j = (warp + nt * NWARP) * 8 + 2 * lr;
      if (j < ldw) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int b = mt * 16 + lq;
          if (b < BT) __stcg(reinterpret_cast<float2*>(dst + b * ldw + j), make_float2(acc[mt][nt][0], acc[mt][nt][1]));
          if (b + 8 < BT)
            __stcg(reinterpret_cast<float2*>(dst + (b + 8) * ldw + j), make_float2(acc[mt][nt][2], acc[mt][nt][3]));
        }
      }
    }
  };

  // ---- distributed reduction of the hidden pre-activations + epilogue ----------------
  //  set 0: pre = sum + hb -> mean (sigmoid) -> mean_out, sample -> HS (and P on the last PCD step)
  //  set 1: PREX = sum + hb (pre-activation of round(v0), pseudo-likelihood)
  auto reduce_hidden = [&](int nsets, float* mean_out, const RngSeg& rs, bool write_hs, bool write_p) {
    const int ldw4 = ldw >> 2;
    const int nq = nsets * BT * p.CQ;
    const int per = (nq + gridDim.x - 1) / gridDim.x;
    const int o0 = cta * per, o1 = min(nq, o0 + per);
    for (int o = o0 + warp; o < o1; o += NWARP) {
      int set = o / (BT * p.CQ), rem = o % (BT * p.CQ);
      int b = rem / p.CQ, qq = rem % p.CQ;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      // up to 8 partials per lane (grids up to 256 CTAs): all loads in flight at once, summed in order
      float4 t8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = lane + 32 * u;
        t8[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < p.n_active)
          t8[u] = __ldcg(reinterpret_cast<const float4*>(p.part + ((size_t)c * 2 + set) * BT * ldw) + b * ldw4 + qq);
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
        float sv = lane == 0 ? s.x : lane == 1 ? s.y : lane == 2 ? s.z : s.w;
        int j = qq * 4 + lane;
        if (j < H) {
          float pre = sv + p.hb[j];
          if (set == 1) {
            __stcg(&p.PREX[b * ldw + j], pre);
          } else {
            float mean = 0.f, smp = 0.f;
            if (b < B) {
              mean = sigmoidf_(pre);
              if (write_hs || write_p) smp = rng_uniform(rs, (long long)b * H + j) < mean ? 1.f : 0.f;
            }
            __stcg(&mean_out[b * ldw + j], mean);
            if (write_hs) __stcg(&p.HS[b * ldw + j], smp);
            if (write_p && b < B) __stcg(&p.P[(size_t)b * H + j], smp);
          }
        }
      }
    }
  };

  auto load_hs = [&](const float* src, int ld_src, int nrows_src) {
    // [BT][ldw] chain state from L2 into the padded shared-memory panel.  float4 loads, four
    // independent loads in flight per thread (a dependent chain of L2 round trips was 3-7 us here).
    int pred = 1;
    const int ldw4 = ldw >> 2, n4 = BT * ldw4;
    const bool vec = (ld_src & 3) == 0 && (((uintptr_t)src) & 15) == 0;
    for (int e0 = tid; e0 < n4; e0 += 4 * NT) {
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * NT;
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < n4) {
          const int b = e / ldw4, j = (e - b * ldw4) * 4;
          if (b < nrows_src) {
            const float* sp = src + (size_t)b * ld_src + j;
            if (vec && j + 3 < H) x[u] = __ldcg(reinterpret_cast<const float4*>(sp));
            else {
              if (j < H) x[u].x = __ldcg(sp);
              if (j + 1 < H) x[u].y = __ldcg(sp + 1);
              if (j + 2 < H) x[u].z = __ldcg(sp + 2);
              if (j + 3 < H) x[u].w = __ldcg(sp + 3);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * NT;
        if (e < n4) {
          const int b = e / ldw4, j = (e - b * ldw4) * 4;
          *reinterpret_cast<float4*>(hs + b * ldh + j) = x[u];
          pred &= (((__float_as_uint(x[u].x) | __float_as_uint(x[u].y) | __float_as_uint(x[u].z) |
                     __float_as_uint(x[u].w)) & 0x1FFFu) == 0u) ? 1 : 0;
        }
      }
    }
    return __syncthreads_and(pred) != 0;      // {0,1} chain states are exact in tf32
  };

  // =============================== pass 0: positive phase ===============================
  {
    const int depth = p.nslots;
    if (warp == 0) for (int j = 1; j < depth; ++j) issue(j, 1, depth, p.TR, p.slot_bytes);
    float acc[MT][NPW][4], acc2[MT][NPW][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NPW; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[mt][nt][c] = acc2[mt][nt][c] = 0.f;
    for (int j = 0, st = 0; j < ntiles; ++j, st = (st + 1 == depth ? 0 : st + 1)) {
      wait_stage(st);
      const float* tile = reinterpret_cast<const float*>(smem + (size_t)st * p.slot_bytes);
      const int nr = min(p.TR, rows - j * p.TR), nr8 = (nr + 7) & ~7;
      up_mma(tile, v0s + (size_t)j * p.TR * BTS, nvs + (size_t)j * p.TR * BTS, nr8, acc, acc2, ALLOW_DUAL && p.pcd != 0, v0_exact,
             x_exact);
      __syncthreads();
      if (warp == 0) issue(j + depth, 1, depth, p.TR, p.slot_bytes);
    }
    mark();   // pass-0 tiles done
    flush_partial(acc, 0);
    if (ALLOW_DUAL && p.pcd) flush_partial(acc2, 1);
  }
  mark();
  grid_sync(p.bar, bar_target);
  mark();
  // CD: chain starts from the fresh sample; PCD: from the persistent chain (src/rbm.py:308-311)
  reduce_hidden(p.pcd ? 2 : 1, p.PH, seg(0, 0), !p.pcd, false);
  mark();   // reduce 0 done
  grid_sync(p.bar, bar_target);
  if (warp == 0) issue(0, 1, p.nslots, p.TR, p.slot_bytes);     // W is unchanged until the update: prefetch the next pass now
  bool h_exact = p.pcd ? load_hs(p.P, H, B) : load_hs(p.HS, ldw, BT);
  mark();

  // pseudo-likelihood monitor (src/rbm.py:421-447) — CTA 0, uses the pre-update W, hb, vb
  if (p.pcd && cta == 0) {
    const int bit = *p.bit_idx;
    for (int b = warp; b < B; b += NWARP) {
      long long dr = p.idx ? p.idx[b] : b;
      float x = roundf(p.data[dr * p.ld_data + bit]);
      float d = 1.f - 2.f * x;
      float h0 = 0.f, h1 = 0.f;
      for (int j = lane; j < H; j += 32) {
        float pre = __ldcg(&p.PREX[b * ldw + j]);
        h0 += softplusf_(pre);
        h1 += softplusf_(pre + d * p.W[(size_t)bit * ldw + j]);
      }
      h0 = warp_sum(h0);
      h1 = warp_sum(h1);
      if (lane == 0) {
        float vbv = p.vb[bit], vterm;
        if (p.kind == MDBN_GRBM) { float a = x - vbv, c = (1.f - x) - vbv; vterm = 0.5f * (a * a - c * c); }
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

  // =============================== passes 1..k: fused Gibbs steps ===============================
  float cost_acc = 0.f;
  for (int s = 0; s < p.k; ++s) {
    const bool last = (s == p.k - 1);
    const long long ubase = (long long)B * H + (long long)s * p.u_step_stride;
    const RngSeg rs_v = seg(ubase + p.u_off_v, 1u + 2u * s);
    const RngSeg rs_h = seg(ubase + p.u_off_h, 2u + 2u * s);
    const int depth = p.nslots;
    if (warp == 0) for (int j = 1; j < depth; ++j) issue(j, 1, depth, p.TR, p.slot_bytes);   // job 0 was prefetched
    float acc[MT][NPW][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NPW; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;

    for (int j = 0, st = 0; j < ntiles; ++j, st = (st + 1 == depth ? 0 : st + 1)) {
      if (j < 3) mark2();
      wait_stage(st);
      if (j < 3) mark2();
      const float* tile = reinterpret_cast<const float*>(smem + (size_t)st * p.slot_bytes);
      const int nr = min(p.TR, rows - j * p.TR), nr8 = (nr + 7) & ~7, ntd = nr8 >> 3;
      // ---- propdown of the tile rows: out[b, i] = sum_j h[b][j] W[i][j]; the 8 warps split j ----
      // mma.sync has a long dependent-issue latency on sm_100: every split term (and, for one
      // m-tile, every other k-step) gets its own accumulator so ~3*KI*ntd MMAs are in flight
      {
        constexpr int KI = MT == 1 ? 2 : 1;
        float dacc[KI][3][MT][NTD][4];
#pragma unroll
        for (int ki = 0; ki < KI; ++ki)
#pragma unroll
          for (int t3 = 0; t3 < 3; ++t3)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
              for (int nt = 0; nt < NTD; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) dacc[ki][t3][mt][nt][c] = 0.f;
        for (int jb = warp * 8; jb < ncols8; jb += NWARP * 8 * KI) {
#pragma unroll
          for (int ki = 0; ki < KI; ++ki) {
            const int j0 = jb + ki * NWARP * 8;
            if (j0 < ncols8) {
              uint32_t ah[MT][4], al[MT][4];
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const float* h0 = hs + (mt * 16 + lq) * ldh + j0 + lr;
                split_tf32(h0[0], ah[mt][0], al[mt][0]);
                split_tf32(h0[8 * ldh], ah[mt][1], al[mt][1]);
                split_tf32(h0[4], ah[mt][2], al[mt][2]);
                split_tf32(h0[8 * ldh + 4], ah[mt][3], al[mt][3]);
              }
#pragma unroll
              for (int nt = 0; nt < NTD; ++nt) {
                if (nt < ntd) {
                  const float* w0 = tile + (nt * 8 + lq) * ldp + j0 + lr;
                  uint32_t bh0, bl0, bh1, bl1;
                  split_tf32(w0[0], bh0, bl0);
                  split_tf32(w0[4], bh1, bl1);
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    mma_tf32(dacc[ki][0][mt][nt], ah[mt], bh0, bh1);
                    mma_tf32(dacc[ki][1][mt][nt], ah[mt], bl0, bl1);
                    if (!h_exact) mma_tf32(dacc[ki][2][mt][nt], al[mt], bh0, bh1);
                  }
                }
              }
            }
          }
        }
#pragma unroll
        for (int nt = 0; nt < NTD; ++nt) {
          if (nt < ntd) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              float r4[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                float small = dacc[0][1][mt][nt][c] + dacc[0][2][mt][nt][c];
                float big = dacc[0][0][mt][nt][c];
                if (KI == 2) {
                  small += dacc[KI - 1][1][mt][nt][c] + dacc[KI - 1][2][mt][nt][c];
                  big += dacc[KI - 1][0][mt][nt][c];
                }
                r4[c] = big + small;
              }
              float* d0 = dred + ((size_t)warp * MB + mt * 16 + lq) * p.TR + nt * 8 + 2 * lr;
              *reinterpret_cast<float2*>(d0) = make_float2(r4[0], r4[1]);
              *reinterpret_cast<float2*>(d0 + 8 * p.TR) = make_float2(r4[2], r4[3]);
            }
          }
        }
      }
      if (j < 3) mark2();
      __syncthreads();
      if (j < 3) mark2();
      // ---- visible epilogue: bias, activation, sampling (src/rbm.py:226-240 / :650-660) ----
      for (int it = tid; it < nr8 * MB; it += NT) {
        const int r = it / MB, b = it % MB;
        float vin = 0.f, mean = 0.f;
        if (r < nr && b < B) {
          float sum = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < NWARP; ++w2) sum += dred[((size_t)w2 * MB + b) * p.TR + r];
          const int gi = row0 + j * p.TR + r;
          const float pre = sum + vbs[j * p.TR + r];
          if (p.kind == MDBN_GRBM) {
            mean = pre;
            vin = pre;        // mean-field visible: h given v_MEAN (src/rbm.py:669)
          } else {
            mean = sigmoidf_(pre);
            vin = rng_uniform(rs_v, (long long)b * V + gi) < mean ? 1.f : 0.f;
          }
          if (last && !p.pcd) {
            const float t = v0s[(j * p.TR + r) * BTS + b];
            if (p.kind == MDBN_GRBM) { float d = sigmoidf_(pre) - t; cost_acc += d * d; }   // :697
            else cost_acc += t * softplusf_(-pre) + (1.f - t) * softplusf_(pre);          // :479-480
          }
        }
        vt[r * BTS + b] = vin;
        if (last) nvs[(j * p.TR + r) * BTS + b] = mean;
      }
      __syncthreads();
      if (j < 3) mark2();
      // ---- propup accumulation from the same tile ----
      up_mma(tile, vt, vt, nr8, acc, acc, false, p.kind == MDBN_RBM, true);
      __syncthreads();
      if (warp == 0) issue(j + depth, 1, depth, p.TR, p.slot_bytes);
    }
    if (last) mark();   // Gibbs tiles done
    flush_partial(acc, 0);
    if (last && !p.pcd) {
      float c = block_sum(cost_acc, misc);
      if (tid == 0) __stcg(&p.cost_part[cta], c);
    }
    grid_sync(p.bar, bar_target);
    if (last) mark();
    reduce_hidden(1, p.NH, rs_h, !last, last && p.pcd);
    grid_sync(p.bar, bar_target);
    if (last) mark();
    if (!last) {
      if (warp == 0) issue(0, 1, p.nslots, p.TR, p.slot_bytes);
      h_exact = load_hs(p.HS, ldw, BT);
    }
  }

  // =============================== statistics + update ===============================
  {
    // short tiles (8 rows) for this pass: it streams 2-3 arrays and wants a deep pipeline
    const int narr = p.wc != 0.f ? 3 : 2;
    const int TRS = 8, slot_s = (TRS * ldp * 4 + 127) & ~127;
    int depth = (p.nslots * p.slot_bytes) / (narr * slot_s);
    depth = depth > MAX_SLOTS ? MAX_SLOTS : depth;
    const int ntiles_s = (rows + TRS - 1) / TRS;
    if (warp == 0) for (int j = 0; j < depth; ++j) issue(j, narr, depth, TRS, slot_s);
    const int ldw4 = ldw >> 2;
    float4 ph[BT], nh[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      if (col_ok) {
        ph[b] = __ldcg(reinterpret_cast<const float4*>(p.PH) + b * ldw4 + q);
        nh[b] = __ldcg(reinterpret_cast<const float4*>(p.NH) + b * ldw4 + q);
      } else {
        ph[b] = nh[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const int ncol = min(4, H - 4 * q);
    for (int j = 0, stg = 0; j < ntiles_s; ++j, stg = (stg + 1 == depth ? 0 : stg + 1)) {
      wait_stage(stg);
      const unsigned char* sb = smem + (size_t)stg * narr * slot_s;
      const float4* wt = reinterpret_cast<const float4*>(sb);
      const float4* st = reinterpret_cast<const float4*>(sb + slot_s);
      const float4* sn = reinterpret_cast<const float4*>(sb + 2 * (size_t)slot_s);
      const int nr = min(TRS, rows - j * TRS);
      if (col_ok) {
        for (int r = g; r < nr; r += p.G) {
          const int lr_ = j * TRS + r;
          float4 w = wt[r * ldp4 + q], sp = st[r * ldp4 + q];
          const float4* a4 = reinterpret_cast<const float4*>(v0s + lr_ * BTS);
          const float4* n4 = reinterpret_cast<const float4*>(nvs + lr_ * BTS);
          float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int b4 = 0; b4 < BTP / 4; ++b4) {
            float4 av = a4[b4], nv = n4[b4];
            float as[4] = {av.x, av.y, av.z, av.w}, ns[4] = {nv.x, nv.y, nv.z, nv.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int b = b4 * 4 + t;
              if (b < BT) {
                gs.x = fmaf(as[t], ph[b].x, gs.x); gs.x = fmaf(-ns[t], nh[b].x, gs.x);
                gs.y = fmaf(as[t], ph[b].y, gs.y); gs.y = fmaf(-ns[t], nh[b].y, gs.y);
                gs.z = fmaf(as[t], ph[b].z, gs.z); gs.z = fmaf(-ns[t], nh[b].z, gs.z);
                gs.w = fmaf(as[t], ph[b].w, gs.w); gs.w = fmaf(-ns[t], nh[b].w, gs.w);
              }
            }
          }
          float wv[4] = {w.x, w.y, w.z, w.w}, sv[4] = {sp.x, sp.y, sp.z, sp.w}, gv[4] = {gs.x, gs.y, gs.z, gs.w};
          float snv[4] = {0.f, 0.f, 0.f, 0.f};
          if (narr > 2) { float4 t4 = sn[r * ldp4 + q]; snv[0] = t4.x; snv[1] = t4.y; snv[2] = t4.z; snv[3] = t4.w; }
          float wo[4], so[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float gw = gv[c] * p.inv_bnom - p.wc * snv[c];                // src/rbm.py:411-415
            float mult = p.decay;
            if (p.c1 != 0.f) {
              // D = 1 + 2 lr lambda_1 / (|W| + eps); MUFU reciprocals (<= 2 ulp) instead of IEEE division
              float invD = __fdividef(1.0f, 1.0f + p.c1 * __fdividef(1.0f, fabsf(wv[c]) + 0.001f));   // :347-350
              gw *= invD;
              mult *= invD;                                               // :353-356
            }
            so[c] = gw + (sv[c] - gw) * p.mom;                            // :361
            wo[c] = wv[c] * mult + sv[c] * p.lr;                          // :364 (OLD speed)
          }
          size_t go = (size_t)(row0 + lr_) * ldw + 4 * q;
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
      if (warp == 0) issue(j + depth, narr, depth, TRS, slot_s);
    }
    // visible bias (rows owned by this CTA)  src/rbm.py:417
    for (int r = tid; r < rows; r += NT) {
      float gsum = 0.f;
      for (int b = 0; b < B; ++b) gsum += v0s[r * BTS + b] - nvs[r * BTS + b];
      float gb = gsum * p.inv_b, sv = p.Svb[row0 + r];
      p.Svb[row0 + r] = gb + (sv - gb) * p.mom;
      p.vb[row0 + r] = p.vb[row0 + r] + sv * p.lr;
    }
    // hidden bias  src/rbm.py:416 — one CTA (the last: it owns the fewest rows)
    if (cta == gridDim.x - 1) {
      for (int j = tid; j < H; j += NT) {
        float gsum = 0.f;
        for (int b = 0; b < B; ++b) gsum += __ldcg(&p.PH[b * ldw + j]) - __ldcg(&p.NH[b * ldw + j]);
        float gb = gsum * p.inv_b, sv = p.Shb[j];
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

  mark();   // stats + update done
  // reset the barrier for the next launch: the last CTA out switches off the lights
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    unsigned long long prev = atomicAdd(p.bar + 1, 1ULL);
    if (prev == gridDim.x - 1) {
      p.bar[0] = 0ULL;
      p.bar[1] = 0ULL;
      __threadfence();
    }
  }
}

struct Geometry {
  int BT, NPW, NT, rows_per_cta, rows_alloc, n_active, CQ, GW, G, TR, nslots, grid, ldp, ldh, slot_bytes;
  int off_hs, off_v0, off_nv, off_vt, off_dred, off_bars, off_misc, off_vb;
  size_t smem;
  bool ok;
};

static int pad_mod32(int n, int want) { return n + ((want - n % 32) + 32) % 32; }

static Geometry plan(const mdbn_ctx* c, const mdbn_cd_args& a) {
  Geometry g{};
  g.ok = false;
  g.BT = a.B <= 10 ? 10 : (a.B <= 20 ? 20 : 0);
  if (!g.BT || a.ldw % 4 != 0) return g;
  if (((uintptr_t)a.W | (uintptr_t)a.W_speed | (uintptr_t)a.W_snap) & 15) return g;
  const int MT = g.BT > 16 ? 2 : 1, MB = 16 * MT, BTS = MT == 1 ? 24 : 40;
  const bool pcd = a.persistent != nullptr;
  g.CQ = a.ldw / 4;
  const int ncols8 = (a.H + 7) & ~7;
  if (ncols8 > 1024) return g;
  // 512 threads (16 warps) was measured slower (Gibbs pass 40 vs 31 us): register spills at 128 regs/thread
  g.NT = 256;
  if (g.CQ > g.NT) return g;
  const bool wide = ncols8 > 512;
  g.NPW = (g.NT == 512) ? (wide ? 8 : 4) : (wide ? 16 : 8);
  if (wide && MT == 2 && pcd) return g;     // register budget of the dual accumulators
  if (g.CQ <= 32) { g.GW = 1; while (g.GW < g.CQ) g.GW <<= 1; } else g.GW = (g.CQ + 31) / 32 * 32;
  g.G = g.NT / g.GW;
  g.grid = c->num_sms;
  g.rows_per_cta = (a.V + g.grid - 1) / g.grid;
  g.rows_alloc = (g.rows_per_cta + 7) & ~7;
  g.n_active = (a.V + g.rows_per_cta - 1) / g.rows_per_cta;
  if (a.ldw != ncols8) return g;      // rows padded to 8 floats (the Python host allocates W that way)
  g.ldp = a.ldw;                      // staged rows keep the global stride: one bulk copy per tile

  g.ldh = pad_mod32(ncols8, 20);
  int tr = (40 * 1024) / (g.ldp * 4);
  tr &= ~7;
  if (tr < 8) return g;
  if (tr > MAX_TR) tr = MAX_TR;
  if (wide && tr > 16) tr = 16;
  // no point in tiles taller than the slab
  while (tr > 8 && tr - 8 >= g.rows_alloc) tr -= 8;
  g.TR = tr;
  g.slot_bytes = (g.TR * g.ldp * 4 + 127) & ~127;
  auto up128 = [](size_t x) { return (x + 127) & ~(size_t)127; };
  const size_t hs_b = up128((size_t)MB * g.ldh * 4), slab_b = up128((size_t)g.rows_alloc * BTS * 4),
               vt_b = up128((size_t)MAX_TR * BTS * 4), dred_b = up128((size_t)(g.NT / 32) * MB * g.TR * 4);
  const size_t vb_b = up128((size_t)g.rows_alloc * 4);
  const size_t fixed = hs_b + 2 * slab_b + vt_b + dred_b + 128 + 256 + vb_b;
  const size_t smem_max = 227 * 1024;
  const int narr = a.weightcost != 0.f ? 3 : 2;
  if (fixed + (size_t)narr * g.slot_bytes > smem_max) return g;
  g.nslots = (int)((smem_max - fixed) / g.slot_bytes);
  if (g.nslots > MAX_SLOTS) g.nslots = MAX_SLOTS;
  if (g.nslots < narr || g.nslots < 2) return g;
  size_t off = (size_t)g.nslots * g.slot_bytes;
  auto take = [&](size_t bytes) { size_t o = off; off += bytes; return (int)o; };
  g.off_hs = take(hs_b);
  g.off_v0 = take(slab_b);
  g.off_nv = take(slab_b);
  g.off_vt = take(vt_b);
  g.off_dred = take(dred_b);
  g.off_bars = take(128);
  g.off_misc = take(256);
  g.off_vb = take(vb_b);
  g.smem = off;
  g.ok = g.smem <= smem_max;
  return g;
}

template <int BT, int NPW, int NT>
static int launch(mdbn_ctx* c, const Params& p, const Geometry& g, cudaStream_t st) {
  static bool configured[64] = {};
  auto kfn = cd_skinny_kernel<BT, NPW, NT>;
  if (!configured[c->device]) {
    MDBN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[c->device] = true;
  }
  void* args[] = {(void*)&p};
  MDBN_CUDA(cudaLaunchCooperativeKernel((void*)kfn, dim3(g.grid), dim3(NT), args, g.smem, st));
  c->launches++;
  return 0;
}

}  // namespace sk

bool skinny_supported(const mdbn_ctx* c, const mdbn_cd_args& a) {
  if (a.phase != MDBN_PHASE_FULL) return false;
  return sk::plan(c, a).ok;
}

int skinny_cd_step(mdbn_ctx* c, const mdbn_cd_args& a, cudaStream_t st) {
  sk::Geometry g = sk::plan(c, a);
  MDBN_CHECK(g.ok, "skinny path: unsupported shape");
  sk::Params p{};
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
  p.rows_per_cta = g.rows_per_cta; p.rows_alloc = g.rows_alloc; p.n_active = g.n_active;
  p.CQ = g.CQ; p.GW = g.GW; p.G = g.G; p.TR = g.TR; p.nslots = g.nslots;
  p.ldp = g.ldp; p.ldh = g.ldh; p.slot_bytes = g.slot_bytes;
  p.off_hs = g.off_hs; p.off_v0 = g.off_v0; p.off_nv = g.off_nv; p.off_vt = g.off_vt; p.off_dred = g.off_dred;
  p.off_bars = g.off_bars; p.off_misc = g.off_misc; p.off_vb = g.off_vb;

  // scratch: [part | PH | NH | HS | PREX | cost_part]; zero-filled whenever (re)allocated so that
  // padded columns of the [BT][ldw] buffers stay zero
  const size_t hb_f = (size_t)g.BT * a.ldw;
  const size_t part_f = (size_t)g.n_active * 2 * hb_f;
  const size_t total_f = part_f + 4 * hb_f + (size_t)g.grid + 64;
  mdbn_ctx::Buf& wb = c->ws[WS_SKINNY];
  const void* before = wb.p;
  const size_t before_n = wb.n;
  float* base = (float*)ws_get(c, WS_SKINNY, total_f * sizeof(float));
  if (!base) return 3;
  // layout depends on (BT, ldw, n_active): re-zero when the allocation or the layout key changes
  static thread_local unsigned long long last_key = 0;
  unsigned long long key = ((unsigned long long)g.BT << 48) ^ ((unsigned long long)a.ldw << 24) ^
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
  p.dbg = want_timing ? reinterpret_cast<unsigned long long*>(c->barrier) + 8 : nullptr;   // needs 32 slots
  int rc = 2;
  if (g.BT == 10 && g.NPW == 8) rc = sk::launch<10, 8, 256>(c, p, g, st);
  else if (g.BT == 10 && g.NPW == 16) rc = sk::launch<10, 16, 256>(c, p, g, st);
  else if (g.BT == 20 && g.NPW == 8) rc = sk::launch<20, 8, 256>(c, p, g, st);
  else if (g.BT == 20 && g.NPW == 16) rc = sk::launch<20, 16, 256>(c, p, g, st);
  else set_error("skinny path: no kernel for BT=%d NPW=%d", g.BT, g.NPW);
  if (rc == 0 && p.dbg) {
    unsigned long long t[32];
    MDBN_CUDA(cudaStreamSynchronize(st));
    MDBN_CUDA(cudaMemcpy(t, p.dbg, sizeof(t), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[skinny timeline us] V=%d H=%d B=%d k=%d:", a.V, a.H, a.B, a.k);
    for (int i = 1; i < 13; ++i) fprintf(stderr, " %.1f", (double)(t[i] - t[0]) * 1e-3);
    fprintf(stderr, "\n   gibbs tiles (wait< wait> D sync epi | ...):");
    for (int i = 16; i < 31; ++i) fprintf(stderr, " %.2f", (double)(t[i] - t[0]) * 1e-3);
    fprintf(stderr, "\n");
  }
  return rc;
}

}  // namespace mdbn
