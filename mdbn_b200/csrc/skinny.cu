// Skinny-batch CD-k / PCD-k step as ONE persistent cooperative kernel (B <= 20).
//
// Regime: batch 10-20 on 10^2..2*10^4-wide layers is 0.36*B flop/byte ->
// bound by streaming W, not by math.  Design:
//   * every CTA owns a contiguous slab of W rows (visible units) for the whole step;
//   * W row tiles are staged in shared memory by TMA (cp.async.bulk.tensor.2d, boxes of 32 columns x
//     R rows, 128-byte swizzle) completing on mbarriers, a ring of <= 56 KB stages.  The swizzle is
//     what makes BOTH access patterns of the step bank-conflict free from one copy of the tile:
//       propup   (thread = column quad, walks rows)   -> the 8 quads of a box row hit 8 distinct 16-byte
//                                                        bank groups
//       propdown (lane = ROW, walks the columns)       -> the 8 rows of an octet hit 8 distinct groups
//     With lane = row a propdown dot product lives in ONE thread: no cross-lane reduction at all (the
//     shuffle reductions / mma.sync fragments of the earlier versions were the bottleneck);
//   * pass 0      : partial  v0 W          (and round(v0) W for the pseudo-likelihood, sharing the W reads)
//   * pass 1..k   : FUSED propdown + propup from the SAME staged tile: v_i = h . W[i,:] is
//     complete inside the owning CTA, its bias/sigmoid/Bernoulli epilogue runs in place and the tile is
//     immediately reused for  h' += v_i W[i,:];  a Gibbs step reads W once, not twice;
//   * hidden pre-activations need all rows: ONE grid barrier per pass.  Every CTA adds its [B,H] partial to a
//     fixed-point (2^-32, int64) accumulator in L2 with red.global.add.u64 (integer sums commute: the result is
//     bitwise reproducible), arrives at the barrier, and afterwards rebuilds the whole chain state itself
//     (bias, sigmoid, element-indexed Philox draw) in compact loops over shared memory;
//   * last pass  : statistics + lambda_1/lambda_2/momentum update fused: W and W_speed tiles
//     are read once and written once; v0 and nv slabs never left shared memory.
// Several steps can be chained in one launch (n_steps, CHAIN instantiation): a CTA only ever reads its own rows
// of W / W_speed / vb, everything that crosses CTAs is ordered by the barriers of the next step.
// All arithmetic is plain fp32 FFMA / FFMA2.  HBM traffic per step: (k+1) reads of W + read W,S + write W,S
// (+ read W_snap) versus the (2k+1)+4 of an unfused implementation.
#include <cuda.h>
#include <stdlib.h>
#include "ctx.h"

namespace mdbn {
namespace sk {

constexpr int NT = 256;
constexpr int NWARP = NT / 32;
constexpr int MAX_SLOTS = 6;

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
  int rows_per_cta, rows_small, n_big, rows_alloc, CQ, GW, G, R, nbox, nslots, ldh, slot_bytes, ring_bytes;
  // global scratch: fixed-point (2^-32) accumulators of the hidden pre-activation sums, [5][BT][ldw]:
  // 0 = positive phase, 1 = round(v0) (pseudo-likelihood), 2..4 = Gibbs steps (rotating).  Zero on entry;
  // `acc_other` is the set of the previous launch, cleared by this one.
  unsigned long long *acc, *acc_other;   // set of step 0 / the other one; they alternate from step to step
  int n_acc;          // BT * ldw
  int n_steps;        // minibatches processed by this launch (idx [n_steps][B], cost_out [n_steps])
  float* cost_part;   // [max(gridDim, BT)]
  float* PHf;         // [BT][ldw] positive-phase hidden means as fp32 (written in slices after barrier 0)
  unsigned long long* bar;   // [0] barrier counter, [1] exit counter
  unsigned long long* dbg;   // optional phase timeline (MDBN_SKINNY_TIMING=1), CTA 0 only
  int dbg_flags;             // MDBN_SKINNY_DEBUG: skip parts of the passes (timing experiments only; results are wrong)
  // smem byte offsets
  int off_hs, off_v0, off_nv, off_vt, off_dred, off_bars, off_misc, off_vb, off_hb;
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
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

// Fixed-point image of a partial sum: integer additions commute, so the grid-wide sum built with
// red.global.add.u64 is bitwise reproducible whatever order the CTAs arrive in.  2^-32 resolution, |x| < 2^31.
__device__ __forceinline__ unsigned long long to_fixed(float x) {
  return (unsigned long long)__float2ll_rn(x * 4294967296.0f);
}
__device__ __forceinline__ float from_fixed(long long s) { return __ll2float_rn(s) * 2.3283064365386963e-10f; }
// sigmoid on the MUFU pipe (ex2.approx + rcp.approx, a few ulp): for the hidden layer rebuilt by every CTA
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast_(float x) { return rcp_approx(1.0f + __expf(-x)); }
__device__ __forceinline__ void red_add_u64(unsigned long long* addr, unsigned long long v) {
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
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

// Split form: arrive as soon as this CTA's contribution is published, keep working, wait later.
__device__ __forceinline__ void grid_arrive(unsigned long long* bar, unsigned long long& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1ULL);
  }
}
__device__ __forceinline__ void grid_wait(unsigned long long* bar, unsigned long long target) {
  if (threadIdx.x == 0) {
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// CHAIN = false is the single-step instantiation (the step loop folds away: same code as before chaining
// existed, which measures ~4 % faster per step than the looped build); CHAIN = true runs p.n_steps steps.
template <int BT, bool CHAIN>
__global__ void __launch_bounds__(NT, 1) cd_skinny_kernel(const __grid_constant__ CUtensorMap tmR,
                                                          const __grid_constant__ CUtensorMap tm8, const Params p) {
  constexpr int BTP = (BT + 3) / 4 * 4, BTS = BTP;
  extern __shared__ __align__(1024) unsigned char smem[];
  float* hs = reinterpret_cast<float*>(smem + p.off_hs);      // [BT][ldh] chain state, zero padded to nbox*32 columns
  float* v0s = reinterpret_cast<float*>(smem + p.off_v0);     // [rows_alloc][BTS]
  float* nvs = reinterpret_cast<float*>(smem + p.off_nv);     // [rows_alloc][BTS]
  float* vt = reinterpret_cast<float*>(smem + p.off_vt);      // [R][BTS] visible tile -> propup input
  float* dred = reinterpret_cast<float*>(smem + p.off_dred);  // [NWARP * 32/R][BT][R] propdown column-split partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);
  float* misc = reinterpret_cast<float*>(smem + p.off_misc);  // [64]: block_sum scratch, pl cost
  float* vbs = reinterpret_cast<float*>(smem + p.off_vb);     // [rows_alloc] visible bias of the owned rows
  float* hbs = reinterpret_cast<float*>(smem + p.off_hb);     // [ldh] hidden bias as it was on entry

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;
  const int ldw = p.ldw, ldh = p.ldh, R = p.R, nbox = p.nbox;
  const int B = p.B, V = p.V, H = p.H;
  // CTAs [0, n_big) own rows_per_cta rows each, the rest rows_small (PCD: the monitor CTAs get fewer rows)
  const int row0 = cta < p.n_big ? cta * p.rows_per_cta : p.n_big * p.rows_per_cta + (cta - p.n_big) * p.rows_small;
  const int rows = max(0, min(cta < p.n_big ? p.rows_per_cta : p.rows_small, V - row0));
  const int ntiles = (rows + R - 1) / R;
  const int box_bytes = R * 128;
  // propup / statistics mapping: thread -> (row group g, column quad q)
  const int q = tid % p.GW, g = tid / p.GW;
  const bool col_ok = g < p.G && q < p.CQ;
  // propdown mapping: lane -> (row of the tile, which of the 32/R boxes handled together)
  const int drow = lane & (R - 1), dsub = lane / R, SUBS = 32 / R;
  unsigned long long bar_target = 0;
  uint32_t phase_bits = 0;
  int dbg_i = 0;
  const int F = p.dbg_flags;
  int step = 0;                                  // minibatch of this launch being processed
  auto mark = [&]() {
    if (p.dbg && cta == 0 && tid == 0 && step == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[dbg_i++] = t;
    }
  };
  mark();
  if (p.dbg && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.dbg[32 + cta] = t; }

  if (tid == 0) {
    for (int i = 0; i < MAX_SLOTS; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid; e < BT * ldh; e += NT) hs[e] = 0.f;
  __syncthreads();

  // ---- tile pipeline --------------------------------------------------------------
  // W-only passes: stage j = rows [j*R, j*R+R) x all columns as nbox swizzled boxes (TMA).  Called by ALL
  // warps: lane 0 of warp w issues boxes w, w+NWARP, ... (one thread issuing a whole stage costs ~1 us per
  // tile on the critical path); warp 0 also posts the byte count.  A complete_tx that overtakes the
  // expect_tx is legal: the phase cannot complete before the arrival that carries the expect_tx.
  auto issue = [&](int j, int depth) {
    const int r0 = j * R;
    if (r0 >= rows || lane != 0) return;
    const int st = j % depth;
    const int nr8 = (min(R, rows - r0) + 7) & ~7;
    uint64_t* bar = &bars[st];
    unsigned char* dst = smem + (size_t)st * p.slot_bytes;
    if (warp == 0) mbar_expect_tx(bar, (uint32_t)nbox * nr8 * 128u);
    if (nr8 == R) {
      for (int bx = warp; bx < nbox; bx += NWARP) tma_load_2d(dst + bx * box_bytes, &tmR, bar, 32 * bx, row0 + r0);
    } else {
      for (int bx = warp; bx < nbox; bx += NWARP)
        for (int t = 0; t < (nr8 >> 3); ++t)
          tma_load_2d(dst + bx * box_bytes + t * 1024, &tm8, bar, 32 * bx, row0 + r0 + 8 * t);
    }
  };
  // statistics pass: plain row tiles of W, S (, W_snap): one 1-D bulk copy per array
  auto issue_rows = [&](int j, int narr, int depth, int tr, int slot_b) {
    const int r0 = j * tr;
    if (r0 >= rows || lane != 0) return;
    const int st = j % depth;
    const int nr = min(tr, rows - r0);
    const uint32_t bytes = (uint32_t)nr * ldw * 4u;
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
    // Philox offset of this step = rng.offset + step (recomputed here: nothing per-step is kept live across
    // the streaming loops, which run at 255 registers)
    const unsigned long long off64 = (((unsigned long long)p.c3 << 32) | p.c2) + (unsigned long long)step;
    s.k0 = p.k0; s.k1 = p.k1; s.c1 = ordinal; s.c2 = (uint32_t)off64; s.c3 = (uint32_t)(off64 >> 32);
    return s;
  };

  // ---- propup of one staged tile: acc[b][4q..4q+3] += src[r][b] * W[r, 4q..4q+3]; DUAL shares the W loads.
  //      Packed FFMA2: an accumulator pair is (row b, row b+1) of one column, the v pair comes straight out
  //      of the LDS.128, the weight is duplicated.  Two rows are in flight per iteration: with two warps per
  //      scheduler the shared-memory latency is otherwise exposed (scripts/ubench/up_rate*.cu) ----
  auto fma_row = [&](const float4& w, const float4 (&v)[BTP / 4], float2 (&acc)[BTP / 2][4]) {
    const float2 wd[4] = {make_float2(w.x, w.x), make_float2(w.y, w.y), make_float2(w.z, w.z), make_float2(w.w, w.w)};
#pragma unroll
    for (int b4 = 0; b4 < BTP / 4; ++b4) {
      const float2 p0 = make_float2(v[b4].x, v[b4].y), p1 = make_float2(v[b4].z, v[b4].w);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[2 * b4][c] = __ffma2_rn(p0, wd[c], acc[2 * b4][c]);
        if (4 * b4 + 2 < BT) acc[2 * b4 + 1][c] = __ffma2_rn(p1, wd[c], acc[2 * b4 + 1][c]);
      }
    }
  };
  auto up_tile = [&](const unsigned char* __restrict__ tile, const float* __restrict__ src,
                     const float* __restrict__ src2, int nr, float2 (&acc)[BTP / 2][4], float2 (&acc2)[BTP / 2][4],
                     bool dual) {
    if (!col_ok) return;
    const unsigned char* bp = tile + (q >> 3) * box_bytes;
    const int c = q & 7, G = p.G;
    auto ldw_ = [&](int r) { return *reinterpret_cast<const float4*>(bp + r * 128 + ((c ^ (r & 7)) << 4)); };
    int r = g;
    for (; r + G < nr; r += 2 * G) {
      const float4 w0 = ldw_(r), w1 = ldw_(r + G);
      float4 v0[BTP / 4], v1[BTP / 4], x0[BTP / 4], x1[BTP / 4];
#pragma unroll
      for (int b4 = 0; b4 < BTP / 4; ++b4) {
        v0[b4] = reinterpret_cast<const float4*>(src + r * BTS)[b4];
        v1[b4] = reinterpret_cast<const float4*>(src + (r + G) * BTS)[b4];
      }
      if (dual) {
#pragma unroll
        for (int b4 = 0; b4 < BTP / 4; ++b4) {
          x0[b4] = reinterpret_cast<const float4*>(src2 + r * BTS)[b4];
          x1[b4] = reinterpret_cast<const float4*>(src2 + (r + G) * BTS)[b4];
        }
      }
      fma_row(w0, v0, acc);
      fma_row(w1, v1, acc);
      if (dual) {
        fma_row(w0, x0, acc2);
        fma_row(w1, x1, acc2);
      }
    }
    if (r < nr) {
      const float4 w0 = ldw_(r);
      float4 v0[BTP / 4];
#pragma unroll
      for (int b4 = 0; b4 < BTP / 4; ++b4) v0[b4] = reinterpret_cast<const float4*>(src + r * BTS)[b4];
      fma_row(w0, v0, acc);
      if (dual) {
#pragma unroll
        for (int b4 = 0; b4 < BTP / 4; ++b4) v0[b4] = reinterpret_cast<const float4*>(src2 + r * BTS)[b4];
        fma_row(w0, v0, acc2);
      }
    }
  };

  // ---- CTA partial [B][H] -> grid-wide sum: the G row groups are added in shared memory (fixed order, hs is
  //      the staging buffer: it is rebuilt from the sums afterwards anyway), then every element goes to the
  //      fixed-point accumulator with one red.global.add.u64; CTAs start at staggered offsets ----
  auto flush_sums = [&](float2 (&acc)[BTP / 2][4], unsigned long long* dst) {
    if (p.G <= 2) {
      // wide layers (one or two row groups): add in place in hs, group after group
      float4* stage = reinterpret_cast<float4*>(hs);
      const int ldh4 = ldh >> 2;
      for (int gg = 0; gg < p.G; ++gg) {
        if (col_ok && g == gg) {
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            float4 a = (b & 1) ? make_float4(acc[b >> 1][0].y, acc[b >> 1][1].y, acc[b >> 1][2].y, acc[b >> 1][3].y)
                               : make_float4(acc[b >> 1][0].x, acc[b >> 1][1].x, acc[b >> 1][2].x, acc[b >> 1][3].x);
            if (gg > 0) {
              const float4 o = stage[b * ldh4 + q];
              a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
            }
            stage[b * ldh4 + q] = a;
          }
        }
        __syncthreads();
      }
    } else {
      // narrow layers have 4-32 row groups: every (group, quad) thread parks its rows in the idle tile ring,
      // then each (b, quad) is summed over the groups in fixed order — two block barriers whatever G is
      float4* park = reinterpret_cast<float4*>(smem);
      if (col_ok) {
#pragma unroll
        for (int b = 0; b < BT; ++b)
          park[(g * BT + b) * p.CQ + q] =
              (b & 1) ? make_float4(acc[b >> 1][0].y, acc[b >> 1][1].y, acc[b >> 1][2].y, acc[b >> 1][3].y)
                      : make_float4(acc[b >> 1][0].x, acc[b >> 1][1].x, acc[b >> 1][2].x, acc[b >> 1][3].x);
      }
      __syncthreads();
      for (int e = tid; e < BT * p.CQ; e += NT) {
        const int b = e / p.CQ, qq = e - b * p.CQ;
        float4 a = park[b * p.CQ + qq];
        for (int gg = 1; gg < p.G; ++gg) {
          const float4 o = park[(gg * BT + b) * p.CQ + qq];
          a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
        }
        *reinterpret_cast<float4*>(hs + b * ldh + 4 * qq) = a;
      }
      __syncthreads();
    }
    if (rows > 0) {
      // flat walk over the B*H elements from a CTA-specific offset (spreads the CTAs over the addresses);
      // (b, j) advance incrementally: no division in the loop
      const int total = B * H;
      int e = tid + (int)(((long long)cta * total) / gridDim.x);
      if (e >= total) e -= total;
      int b = e / H, j = e - b * H;
      for (int i = tid; i < total; i += NT) {
        red_add_u64(dst + b * ldw + j, to_fixed(hs[b * ldh + j]));
        j += NT;
        while (j >= H) { j -= H; if (++b == B) b = 0; }
      }
    }
    __syncthreads();
  };

  // ---- summed pre-activations -> hs as fp32 (no bias yet): [B][H] quads, eight L2 loads in flight per
  //      thread.  Rows b >= B are zero.  The consumers below are COMPACT loops over shared memory: this code
  //      runs once per pass, so straight-line unrolled math would be bound by instruction fetch ----
  auto sums_to_hs = [&](const unsigned long long* src) {
    constexpr int UB = 8;
    const int n = BT * p.CQ;
    for (int e0 = tid; e0 < n; e0 += UB * NT) {
      longlong2 t[UB][2];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int e = e0 + u * NT, b = e / p.CQ, qq = e - b * p.CQ;
        t[u][0] = t[u][1] = make_longlong2(0, 0);
        if (e < n && b < B) {
          const longlong2* sp = reinterpret_cast<const longlong2*>(src + b * ldw + 4 * qq);
          t[u][0] = __ldcg(sp);
          t[u][1] = __ldcg(sp + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int e = e0 + u * NT, b = e / p.CQ, qq = e - b * p.CQ;
        if (e < n)
          *reinterpret_cast<float4*>(hs + b * ldh + 4 * qq) =
              make_float4(from_fixed(t[u][0].x), from_fixed(t[u][0].y), from_fixed(t[u][1].x), from_fixed(t[u][1].y));
      }
    }
    __syncthreads();
  };
  // ---- hidden layer from the summed pre-activations: mean = sigmoid(sum + hb), sample ~ Bernoulli(mean) into
  //      the shared-memory chain state.  EVERY CTA does this for the whole [B][H] (the draws are indexed by
  //      element, so all CTAs get the same sample); the last CTA also stores the persistent chain ----
  auto hidden_from_sums = [&](const unsigned long long* src, const RngSeg& rs, bool write_p) {
    sums_to_hs(src);
    const bool quad_rng = rs.mode != MDBN_RNG_BUFFER && (H & 3) == 0;
    const int n = B * p.CQ;
#pragma unroll 2
    for (int e = tid; e < n; e += NT) {
      const int b = e / p.CQ, j0 = 4 * (e - b * p.CQ);
      float4* hp = reinterpret_cast<float4*>(hs + b * ldh + j0);
      const float4 x = *hp, hb4 = *reinterpret_cast<const float4*>(hbs + j0);
      const float mean[4] = {sigmoid_fast_(x.x + hb4.x), sigmoid_fast_(x.y + hb4.y), sigmoid_fast_(x.z + hb4.z),
                             sigmoid_fast_(x.w + hb4.w)};
      float u[4];
      if (quad_rng) {
        const long long e0 = (long long)b * H + j0;     // multiple of 4: one Philox block serves the quad
        const Philox4 ph4 = philox4x32_10((uint32_t)(e0 >> 2), rs.c1, rs.c2, rs.c3, rs.k0, rs.k1);
        u[0] = u24(ph4.x); u[1] = u24(ph4.y); u[2] = u24(ph4.z); u[3] = u24(ph4.w);
      } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) u[t] = j0 + t < H ? rng_uniform(rs, (long long)b * H + j0 + t) : 2.f;
      }
      float sm4[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        sm4[t] = (j0 + t < H && u[t] < mean[t]) ? 1.f : 0.f;
        if (write_p && j0 + t < H) p.P[(size_t)b * H + j0 + t] = sm4[t];
      }
      *hp = make_float4(sm4[0], sm4[1], sm4[2], sm4[3]);
    }
    __syncthreads();
  };
  // PCD: chain state from the persistent chain [B][H] (src/rbm.py:308-311)
  auto load_chain = [&]() {
    constexpr int UB = 8;
    const bool vec = (H & 3) == 0 && (((uintptr_t)p.P) & 15) == 0;
    const int n = BT * p.CQ;
    for (int e0 = tid; e0 < n; e0 += UB * NT) {
      float4 x[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int e = e0 + u * NT, b = e / p.CQ, j0 = 4 * (e - b * p.CQ);
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < n && b < B) {
          const float* sp = p.P + (size_t)b * H + j0;
          if (vec && j0 + 3 < H) x[u] = __ldcg(reinterpret_cast<const float4*>(sp));
          else {
            if (j0 < H) x[u].x = __ldcg(sp);
            if (j0 + 1 < H) x[u].y = __ldcg(sp + 1);
            if (j0 + 2 < H) x[u].z = __ldcg(sp + 2);
            if (j0 + 3 < H) x[u].w = __ldcg(sp + 3);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int e = e0 + u * NT, b = e / p.CQ, j0 = 4 * (e - b * p.CQ);
        if (e < n) *reinterpret_cast<float4*>(hs + b * ldh + j0) = x[u];
      }
    }
    __syncthreads();
  };

  // the pseudo-likelihood bit index advances by one per step (src/rbm.py:445): read once, written once
  const int pl_b = (int)gridDim.x - 1 - cta;
  const bool pl_cta = p.pcd && pl_b < B;
  const int bit0 = (p.pcd && (cta == 0 || pl_cta)) ? *p.bit_idx : 0;

  // ======================= one CD-k / PCD-k step per iteration =======================
  // Steps of one launch need no extra synchronisation: a CTA only ever reads ITS rows of W, W_speed and vb;
  // what crosses CTAs (hidden bias, persistent chain, accumulators, cost partials) is written after the last
  // barrier of step s and read after the first barrier of step s+1.
  const int n_steps = CHAIN ? p.n_steps : 1;
  for (step = 0; step < n_steps; ++step) {
  // accumulator set of this step / of the previous one (alternating), recomputed where they are needed
  auto acc_set = [&]() { return (step & 1) ? p.acc_other : p.acc; };
  auto GA = [&](int s) { return acc_set() + (size_t)(2 + s % 3) * p.n_acc; };
#define A0 (acc_set())
#define A1 (acc_set() + p.n_acc)

  // ---- gather v0 slab: v0s[r][b] = data[idx[b]][row0 + r]; rows >= `rows` and b >= B are zero -----
  if (!(F & 2)) issue(0, p.nslots);     // start streaming W while the minibatch is gathered
  for (int r = tid; r < p.rows_alloc; r += NT) vbs[r] = r < rows ? __ldcg(&p.vb[row0 + r]) : 0.f;
  // the minibatch row numbers first (one dependent load for everybody), then eight gathers in flight per thread
  int* sidx = reinterpret_cast<int*>(misc) + 32;
  const int* idxp = p.idx ? p.idx + (size_t)step * B : nullptr;      // row numbers of this step's minibatch
  if (tid < BTS) sidx[tid] = tid < B ? (idxp ? idxp[tid] : tid) : -1;
  // small vectors that are only needed after the first barrier or in the last pass: pull them into L2 now
  // (with the weights of several layers in rotation they have been evicted since the previous step)
  {
    auto pf = [](const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); };
    const int hb_lines = (H * 4 + 127) / 128, sv_lines = (rows * 4 + 127) / 128 + 1;
    if (tid < hb_lines) pf(reinterpret_cast<const char*>(p.hb) + tid * 128);
    else if (tid < hb_lines + sv_lines && rows > 0)
      pf(reinterpret_cast<const char*>(p.Svb + row0) + min((tid - hb_lines) * 128, rows * 4 - 4));
    else if (cta == (int)gridDim.x - 1 && tid < 2 * hb_lines + sv_lines)
      pf(reinterpret_cast<const char*>(p.Shb) + (tid - hb_lines - sv_lines) * 128);
    if (p.pcd) {
      const int p_lines = (B * H * 4 + 127) / 128;
      for (int i = cta * NT + tid; i < p_lines; i += gridDim.x * NT) pf(reinterpret_cast<const char*>(p.P) + (size_t)i * 128);
    }
  }
  // pseudo-likelihood monitor: the CTA that will compute it for minibatch row pl_b fetches its scalars now
  // (three dependent loads that would otherwise sit on that CTA's critical path)
  if (pl_cta && tid == NT - 1) {
    const int bit = (bit0 + step) % V;                                 // src/rbm.py:445, one advance per step
    const long long dr = idxp ? idxp[pl_b] : pl_b;
    misc[60] = __int_as_float(bit);
    misc[61] = roundf(p.data[dr * p.ld_data + bit]);
    // (vb[bit] belongs to another CTA, which may still be updating it for the previous step of this launch:
    //  it is read after the first barrier of the step, below)
  }
  __syncthreads();
  {
    constexpr int UG = 8;
    const int n = p.rows_alloc * BTS;
    for (int e0 = tid; e0 < n; e0 += UG * NT) {
      float x[UG];
#pragma unroll
      for (int u = 0; u < UG; ++u) {
        const int e = e0 + u * NT, b = e / p.rows_alloc, r = e - b * p.rows_alloc;
        x[u] = 0.f;
        if (e < n && r < rows) {
          const int dr = sidx[b];
          if (dr >= 0) x[u] = __ldg(&p.data[(long long)dr * p.ld_data + row0 + r]);
        }
      }
#pragma unroll
      for (int u = 0; u < UG; ++u) {
        const int e = e0 + u * NT, b = e / p.rows_alloc, r = e - b * p.rows_alloc;
        if (e < n) {
          v0s[r * BTS + b] = x[u];
          nvs[r * BTS + b] = p.pcd ? roundf(x[u]) : 0.f;   // src/rbm.py:428; the nv slab is free until the last Gibbs step
        }
      }
    }
  }
  __syncthreads();
  mark();   // gather done

  // =============================== pass 0: positive phase ===============================
  {
    const int depth = p.nslots;
    if (!(F & 2)) for (int j = 1; j < depth; ++j) issue(j, depth);
    float2 acc[BTP / 2][4], acc2[BTP / 2][4];
#pragma unroll
    for (int b = 0; b < BTP / 2; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[b][c] = acc2[b][c] = make_float2(0.f, 0.f);
    for (int j = 0, st = 0; j < ntiles; ++j, st = (st + 1 == depth ? 0 : st + 1)) {
      if (!(F & 2)) wait_stage(st);
      const unsigned char* tile = smem + (size_t)st * p.slot_bytes;
      const int nr = min(R, rows - j * R);
      if (!(F & 1)) up_tile(tile, v0s + (size_t)j * R * BTS, nvs + (size_t)j * R * BTS, nr, acc, acc2, p.pcd != 0);
      __syncthreads();
      if (!(F & 2)) issue(j + depth, depth);
    }
    mark();   // pass-0 tiles done
    flush_sums(acc, A0);
    grid_arrive(p.bar, bar_target);
    // the round(v0) sums only feed the pseudo-likelihood monitor: published behind the main barrier on a
    // counter of their own (bar[2]) that only the CTAs computing the monitor ever wait for
    if (p.pcd) flush_sums(acc2, A1);
  }
  mark();
  grid_wait(p.bar, bar_target);
  mark();
  if (!(F & 32)) issue(0, p.nslots);     // W is unchanged until the update: prefetch the next pass now
  // every CTA is past the previous step now: the hidden bias it wrote is final, and nobody reads the
  // accumulator set of the previous step (or launch) any more -> clear it for the next one
  for (int j = tid; j < ldh; j += NT) hbs[j] = j < H ? __ldcg(&p.hb[j]) : 0.f;
  {
    unsigned long long* acc_prev = (step & 1) ? p.acc : p.acc_other;
    for (int i = cta * NT + tid; i < 5 * p.n_acc; i += gridDim.x * NT) __stcg(&acc_prev[i], 0ULL);
  }
  // positive-phase means as fp32 for the statistics pass: every CTA converts one slice (published by the
  // barriers that follow), so that pass does not pay a sum -> mean round trip.  Loads first, the chain
  // state is rebuilt while they are in flight.
  const int ph_per = (BT * p.CQ + (int)gridDim.x - 1) / (int)gridDim.x;   // <= NT (plan())
  const int ph_e = cta * ph_per + tid;
  const bool ph_mine = tid < ph_per && ph_e < BT * p.CQ;
  const int ph_b = ph_e / p.CQ, ph_q = ph_e - ph_b * p.CQ;
  longlong2 ph_s01 = make_longlong2(0, 0), ph_s23 = ph_s01;
  if (ph_mine && ph_b < B) {
    const longlong2* sp = reinterpret_cast<const longlong2*>(A0 + ph_b * ldw + 4 * ph_q);
    ph_s01 = __ldcg(sp);
    ph_s23 = __ldcg(sp + 1);
  }
  // CD: chain starts from the fresh sample; PCD: from the persistent chain (src/rbm.py:308-311)
  if (p.pcd) {
    load_chain();
    unsigned long long aux = 0;
    grid_arrive(p.bar + 2, aux);     // round(v0) sums of this CTA: landed long ago, the fence is free here
  } else {
    hidden_from_sums(A0, seg(0, 0), false);
  }
  if (ph_mine) {
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ph_b < B) {
      const float4 hb4 = *reinterpret_cast<const float4*>(hbs + 4 * ph_q);
      m = make_float4(sigmoid_fast_(from_fixed(ph_s01.x) + hb4.x), sigmoid_fast_(from_fixed(ph_s01.y) + hb4.y),
                      sigmoid_fast_(from_fixed(ph_s23.x) + hb4.z), sigmoid_fast_(from_fixed(ph_s23.y) + hb4.w));
    }
    __stcg(reinterpret_cast<float4*>(p.PHf + ph_b * ldw + 4 * ph_q), m);
  }
  mark();   // chain state ready

  // pseudo-likelihood monitor (src/rbm.py:421-447), pre-update W, hb, vb: one minibatch row per CTA, taken
  // from the END of the grid (the last CTA owns the fewest rows); its loads overlap the tile prefetch above
  if (pl_cta) {
    const int bit = __float_as_int(misc[60]);
    const float x = misc[61], d = 1.f - 2.f * x;
    const float vbv = tid == 0 ? __ldcg(&p.vb[bit]) : 0.f;
    // W[bit, :] does not depend on the barrier: in flight while thread 0 polls it
    float wrow[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int j = tid + u * NT; wrow[u] = j < H ? __ldcg(&p.W[(size_t)bit * ldw + j]) : 0.f; }
    grid_wait(p.bar + 2, (unsigned long long)(step + 1) * gridDim.x);
    float pre[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int j = tid + u * NT; pre[u] = j < H ? from_fixed((long long)__ldcg(&A1[pl_b * ldw + j])) + hbs[j] : 0.f; }
    float h0 = 0.f, h1 = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (tid + u * NT < H) {
        h0 += softplusf_(pre[u]);
        h1 += softplusf_(pre[u] + d * wrow[u]);
      }
    }
    h0 = block_sum(h0, misc);
    h1 = block_sum(h1, misc);
    if (tid == 0) {
      float vterm;
      if (p.kind == MDBN_GRBM) { const float a = x - vbv, c = (1.f - x) - vbv; vterm = 0.5f * (a * a - c * c); }
      else vterm = d * vbv;
      __stcg(&p.cost_part[pl_b], -(float)V * softplusf_((h1 - h0) + vterm));
    }
  }

  // =============================== passes 1..k: fused Gibbs steps ===============================
  float cost_acc = 0.f;
  for (int s = 0; s < p.k; ++s) {
    const bool last = (s == p.k - 1);
    const long long ubase = (long long)B * H + (long long)s * p.u_step_stride;
    const RngSeg rs_v = seg(ubase + p.u_off_v, 1u + 2u * s);
    const RngSeg rs_h = seg(ubase + p.u_off_h, 2u + 2u * s);
    const int depth = p.nslots;
    if (!(F & 32)) for (int j = 1; j < depth; ++j) issue(j, depth);   // job 0 was prefetched
    float2 acc[BTP / 2][4];
#pragma unroll
    for (int b = 0; b < BTP / 2; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[b][c] = make_float2(0.f, 0.f);

    for (int j = 0, st = 0; j < ntiles; ++j, st = (st + 1 == depth ? 0 : st + 1)) {
      if (!(F & 32)) wait_stage(st);
      const unsigned char* tile = smem + (size_t)st * p.slot_bytes;
      const int nr = min(R, rows - j * R);
      // ---- propdown of the tile rows, lane = row: out[b] = sum_j h[b][j] W[row][j] stays in one thread;
      //      the warps (and lane groups) split the boxes of 32 columns.  This phase is bound by the
      //      shared-memory wavefronts of the h broadcasts, so with full tiles (R = 32) a lane takes TWO rows
      //      (l and l+16; the half-warps work on different boxes): every h load then feeds 8 FMAs ----
      if (!(F & 4)) {
        if (R == 32) {
          const int lr = lane & 15, half = lane >> 4;
          float da[BT], db[BT];
#pragma unroll
          for (int b = 0; b < BT; ++b) da[b] = db[b] = 0.f;
          for (int bx = warp * 2 + half; bx < nbox; bx += NWARP * 2) {
            const unsigned char* bpa = tile + bx * box_bytes + lr * 128;
            const float* hb0 = hs + bx * 32;
            // NOT fully unrolled: a box is visited once per tile, straight-line code this long is bound by
            // instruction fetch (ncu: stall_no_instruction); the 2-chunk body is re-run from the i-cache
#pragma unroll 2
            for (int c = 0; c < 8; ++c) {
              const int sw = (c ^ (lr & 7)) << 4;                                  // rows l and l+16 swizzle alike
              const float4 wa = *reinterpret_cast<const float4*>(bpa + sw);
              const float4 wb = *reinterpret_cast<const float4*>(bpa + 16 * 128 + sw);
#pragma unroll
              for (int b = 0; b < BT; ++b) {
                const float4 h4 = *reinterpret_cast<const float4*>(hb0 + b * ldh + c * 4);   // broadcast per half-warp
                da[b] = fmaf(h4.x, wa.x, fmaf(h4.y, wa.y, fmaf(h4.z, wa.z, fmaf(h4.w, wa.w, da[b]))));
                db[b] = fmaf(h4.x, wb.x, fmaf(h4.y, wb.y, fmaf(h4.z, wb.z, fmaf(h4.w, wb.w, db[b]))));
              }
            }
          }
          // the two half-warps hold partials of the same 32 rows: add them, lanes 0-15 store both rows
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            da[b] += __shfl_xor_sync(0xffffffffu, da[b], 16);
            db[b] += __shfl_xor_sync(0xffffffffu, db[b], 16);
          }
          if (half == 0) {
            float* d0 = dred + (size_t)warp * BT * 32 + lr;
#pragma unroll
            for (int b = 0; b < BT; ++b) { d0[b * 32] = da[b]; d0[b * 32 + 16] = db[b]; }
          }
        } else {
          float dacc[BT];
#pragma unroll
          for (int b = 0; b < BT; ++b) dacc[b] = 0.f;
          for (int bx = warp * SUBS + dsub; bx < nbox; bx += NWARP * SUBS) {
            const unsigned char* bp = tile + bx * box_bytes + drow * 128;
            const float* hb0 = hs + bx * 32;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 w = *reinterpret_cast<const float4*>(bp + ((c ^ (drow & 7)) << 4));
#pragma unroll
              for (int b = 0; b < BT; ++b) {
                const float4 h4 = *reinterpret_cast<const float4*>(hb0 + b * ldh + c * 4);     // warp-broadcast
                dacc[b] = fmaf(h4.x, w.x, fmaf(h4.y, w.y, fmaf(h4.z, w.z, fmaf(h4.w, w.w, dacc[b]))));
              }
            }
          }
          float* d0 = dred + (size_t)(warp * SUBS + dsub) * BT * R + drow;
#pragma unroll
          for (int b = 0; b < BT; ++b) d0[b * R] = dacc[b];
        }
      }
      __syncthreads();
      // ---- visible epilogue: bias, activation, sampling (src/rbm.py:226-240 / :650-660); consecutive threads
      //      take consecutive rows so that the partial sums are read without bank conflicts ----
      const int nparts = R == 32 ? NWARP : NWARP * SUBS;
      for (int it = tid; it < R * BTS && !(F & 8); it += NT) {
        const int b = it / R, r = it & (R - 1);
        if (r >= nr) continue;
        float vin = 0.f, mean = 0.f;
        if (b < B) {
          float sum = 0.f;
          for (int w2 = 0; w2 < nparts; ++w2) sum += dred[((size_t)w2 * BT + b) * R + r];
          const int lr_ = j * R + r;
          const float pre = sum + vbs[lr_];
          if (p.kind == MDBN_GRBM) {
            mean = pre;
            vin = pre;        // mean-field visible: h given v_MEAN (src/rbm.py:669)
          } else {
            mean = sigmoidf_(pre);
            vin = rng_uniform(rs_v, (long long)b * V + row0 + lr_) < mean ? 1.f : 0.f;
          }
          if (last && !p.pcd) {
            const float t = v0s[lr_ * BTS + b];
            if (p.kind == MDBN_GRBM) { const float d = sigmoidf_(pre) - t; cost_acc += d * d; }   // :697
            else cost_acc += t * softplusf_(-pre) + (1.f - t) * softplusf_(pre);                  // :479-480
          }
        }
        vt[r * BTS + b] = vin;
        if (last) nvs[(j * R + r) * BTS + b] = mean;
      }
      __syncthreads();
      // ---- propup accumulation from the same tile ----
      if (!(F & 16)) up_tile(tile, vt, vt, nr, acc, acc, false);
      __syncthreads();
      if (!(F & 32)) issue(j + depth, depth);
    }
    if (last) mark();   // Gibbs tiles done
    // accumulator of step s+1 was last used by step s-2 of THIS launch: everybody finished reading it before
    // the previous barrier, nobody adds to it before the next one
    if (s >= 2 && s + 1 < p.k) {
      unsigned long long* z = GA(s + 1);
      for (int i = cta * NT + tid; i < p.n_acc; i += gridDim.x * NT) __stcg(&z[i], 0ULL);
    }
    flush_sums(acc, GA(s));
    if (last && !p.pcd) {
      const float c = block_sum(cost_acc, misc);
      if (tid == 0) __stcg(&p.cost_part[cta], c);
    }
    if (last) mark();
    grid_sync(p.bar, bar_target);
    if (last) mark();
    if (!last) {
      if (!(F & 32)) issue(0, p.nslots);
      hidden_from_sums(GA(s), rs_h, false);
    } else if (p.pcd && cta == (int)gridDim.x - 1) {
      hidden_from_sums(GA(s), rs_h, true);     // new persistent chain (src/rbm.py:372)
    }
  }

  // =============================== statistics + update ===============================
  {
    // short tiles (8 rows) for this pass: it streams 2-3 arrays and wants a deep pipeline
    const int narr = p.wc != 0.f ? 3 : 2;
    const int TRS = 8, slot_s = (TRS * ldw * 4 + 127) & ~127;
    int depth = (p.nslots * p.slot_bytes) / (narr * slot_s);
    depth = depth > MAX_SLOTS ? MAX_SLOTS : depth;
    const int ntiles_s = (rows + TRS - 1) / TRS;
    if (warp == 0) for (int j = 0; j < depth; ++j) issue_rows(j, narr, depth, TRS, slot_s);
    const int ldw4x = ldw >> 2;
    float4 ph[BT], nh[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b)
      ph[b] = col_ok ? __ldcg(reinterpret_cast<const float4*>(p.PHf + b * ldw + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (BT <= 10) {
      // every thread converts the sums of its own columns: all loads in flight, no shared-memory round
      const unsigned long long* GL = GA(p.k - 1);
      const float4 hb4 = *reinterpret_cast<const float4*>(hbs + 4 * (col_ok ? q : 0));
      longlong2 t[BT][2];
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        t[b][0] = t[b][1] = make_longlong2(0, 0);
        if (col_ok && b < B) {
          const longlong2* sp = reinterpret_cast<const longlong2*>(GL + b * ldw + 4 * q);
          t[b][0] = __ldcg(sp);
          t[b][1] = __ldcg(sp + 1);
        }
      }
#pragma unroll
      for (int b = 0; b < BT; ++b)
        nh[b] = (col_ok && b < B)
                    ? make_float4(sigmoid_fast_(from_fixed(t[b][0].x) + hb4.x), sigmoid_fast_(from_fixed(t[b][0].y) + hb4.y),
                                  sigmoid_fast_(from_fixed(t[b][1].x) + hb4.z), sigmoid_fast_(from_fixed(t[b][1].y) + hb4.w))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      // (the BT = 20 instantiation has no registers for 40 loads in flight next to ph and nh)
      sums_to_hs(GA(p.k - 1));
      const int n = B * p.CQ;
#pragma unroll 1
      for (int e = tid; e < n; e += NT) {
        const int b = e / p.CQ, j0 = 4 * (e - b * p.CQ);
        float4* hp = reinterpret_cast<float4*>(hs + b * ldh + j0);
        const float4 x = *hp, hb4 = *reinterpret_cast<const float4*>(hbs + j0);
        *hp = make_float4(sigmoid_fast_(x.x + hb4.x), sigmoid_fast_(x.y + hb4.y), sigmoid_fast_(x.z + hb4.z),
                          sigmoid_fast_(x.w + hb4.w));
      }
      __syncthreads();
#pragma unroll
      for (int b = 0; b < BT; ++b)
        nh[b] = col_ok ? *reinterpret_cast<const float4*>(hs + b * ldh + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
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
        // (measured: packed FFMA2 and two rows in flight are both SLOWER here than this plain loop)
        for (int r = g; r < nr; r += p.G) {
          const int lr_ = j * TRS + r;
          float4 w = wt[r * ldw4x + q], sp = st[r * ldw4x + q];
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
          if (narr > 2) { float4 t4 = sn[r * ldw4x + q]; snv[0] = t4.x; snv[1] = t4.y; snv[2] = t4.z; snv[3] = t4.w; }
          float wo[4], so[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float gw = gv[c] * p.inv_bnom - p.wc * snv[c];                // src/rbm.py:411-415
            float mult = p.decay;
            if (p.c1 != 0.f) {
              // 1/D, D = 1 + 2 lr lambda_1 / (|W| + eps)  ==  (|W| + eps) / (|W| + eps + 2 lr lambda_1):
              // ONE MUFU reciprocal (<= 1 ulp) instead of two IEEE divisions                      :347-350
              const float t = fabsf(wv[c]) + 0.001f;
              const float invD = t * rcp_approx(t + p.c1);
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
      if (warp == 0) issue_rows(j + depth, narr, depth, TRS, slot_s);
    }
    // visible bias (rows owned by this CTA)  src/rbm.py:417
    for (int r = tid; r < rows; r += NT) {
      float gsum = 0.f;
      for (int b = 0; b < B; ++b) gsum += v0s[r * BTS + b] - nvs[r * BTS + b];
      float gb = gsum * p.inv_b, sv = p.Svb[row0 + r];
      p.Svb[row0 + r] = gb + (sv - gb) * p.mom;
      p.vb[row0 + r] = vbs[r] + sv * p.lr;
    }
    // hidden bias  src/rbm.py:416 — one CTA (the last: it owns the fewest rows), from the means already in
    // the registers of row group 0
    if (cta == (int)gridDim.x - 1 && col_ok && g == 0) {
      float gs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        if (b < B) {
          gs4[0] += ph[b].x - nh[b].x; gs4[1] += ph[b].y - nh[b].y;
          gs4[2] += ph[b].z - nh[b].z; gs4[3] += ph[b].w - nh[b].w;
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = 4 * q + t;
        if (j < H) {
          const float gb = gs4[t] * p.inv_b, sv = p.Shb[j];
          p.Shb[j] = gb + (sv - gb) * p.mom;
          p.hb[j] = hbs[j] + sv * p.lr;
        }
      }
    }
    if (cta == 0 && warp == 0) {
      // cost = fixed-order sum of the per-CTA (CD) / per-row (PCD) partials: the loads of one lane are
      // independent and in flight together, the lanes are combined by the fixed shuffle tree
      const int n = p.pcd ? B : (int)gridDim.x;
      float c = 0.f;
      for (int i = lane; i < n; i += 32) c += __ldcg(&p.cost_part[i]);
      c = warp_sum(c) * p.cost_scale;
      if (lane == 0) {
        if (p.pcd && step == n_steps - 1) *p.bit_idx = (bit0 + n_steps) % V;     // :445
        if (p.cost_out) p.cost_out[step] = c;
      }
    }
  }

  mark();   // stats + update done
  // the rows this CTA just wrote with ordinary stores are read by TMA (async proxy) in the next step
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  }   // step
#undef A0
#undef A1

  // reset the barrier for the next launch: the last CTA out switches off the lights
  __syncthreads();
  if (p.dbg && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.dbg[32 + 256 + cta] = t; }
  if (tid == 0) {
    __threadfence();
    const unsigned long long prev = atomicAdd(p.bar + 1, 1ULL);
    if (prev == gridDim.x - 1) {
      p.bar[0] = 0ULL;
      p.bar[2] = 0ULL;
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
static int make_map(CUtensorMap* tm, const float* W, int V, int ldw, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)ldw, (cuuint64_t)V};
  cuuint64_t strides[1] = {(cuuint64_t)ldw * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)W, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MDBN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

struct Geometry {
  int BT, rows_per_cta, rows_small, n_big, rows_alloc, CQ, GW, G, R, nbox, nslots, grid, ldh, slot_bytes, ring_bytes;
  int off_hs, off_v0, off_nv, off_vt, off_dred, off_bars, off_misc, off_vb, off_hb;
  size_t smem;
  bool ok;
};

static Geometry plan(const mdbn_ctx* c, const mdbn_cd_args& a) {
  Geometry g{};
  g.ok = false;
  g.BT = a.B <= 10 ? 10 : (a.B <= 20 ? 20 : 0);
  if (!g.BT || a.ldw % 4 != 0 || !get_encode()) return g;
  if (((uintptr_t)a.W | (uintptr_t)a.W_speed | (uintptr_t)a.W_snap) & 15) return g;
  const int BTS = (g.BT + 3) / 4 * 4;
  g.CQ = a.ldw / 4;
  if (g.CQ > NT) return g;
  if (g.CQ <= 32) { g.GW = 1; while (g.GW < g.CQ) g.GW <<= 1; } else g.GW = (g.CQ + 31) / 32 * 32;
  g.G = NT / g.GW;
  g.grid = c->num_sms;
  if (g.BT * g.CQ > g.grid * NT || g.grid < g.BT || a.ldw > 4 * NT) return g;
  g.rows_per_cta = (a.V + g.grid - 1) / g.grid;
  g.rows_small = g.rows_per_cta;
  g.n_big = g.grid;
  if (a.persistent && a.V >= 8 * g.grid) {
    // PCD: the last B CTAs also compute the pseudo-likelihood monitor (~2.5 us); they own ~14 % fewer rows
    g.n_big = g.grid - a.B;
    g.rows_per_cta = (int)((100LL * a.V + (100LL * g.grid - 14LL * a.B) - 1) / (100LL * g.grid - 14LL * a.B));
    const int rest = a.V - g.n_big * g.rows_per_cta;
    g.rows_small = rest > 0 ? (rest + a.B - 1) / a.B : 0;
    if (g.rows_small > g.rows_per_cta) { g.rows_per_cta = (a.V + g.grid - 1) / g.grid; g.rows_small = g.rows_per_cta; g.n_big = g.grid; }
  }
  g.rows_alloc = (g.rows_per_cta + 7) & ~7;
  g.nbox = (a.ldw + 31) / 32;
  g.ldh = g.nbox * 32;
  // stage = R rows x all columns as nbox swizzled boxes, at most 56 KB
  g.R = 32;
  while (g.R > 8 && (g.nbox * g.R * 128 > 56 * 1024 || g.R / 2 >= g.rows_alloc)) g.R >>= 1;
  if (g.nbox * g.R * 128 > 64 * 1024) return g;
  g.slot_bytes = g.nbox * g.R * 128;
  auto up128 = [](size_t x) { return (x + 127) & ~(size_t)127; };
  const size_t hs_b = up128((size_t)g.BT * g.ldh * 4), slab_b = up128((size_t)g.rows_alloc * BTS * 4),
               vt_b = up128((size_t)32 * BTS * 4), dred_b = up128((size_t)NWARP * 32 * g.BT * 4),
               vb_b = up128((size_t)g.rows_alloc * 4), hb_b = up128((size_t)g.ldh * 4);
  const size_t fixed = hs_b + 2 * slab_b + vt_b + dred_b + 128 + 256 + vb_b + hb_b;
  const size_t smem_max = 227 * 1024;
  if (fixed + 2 * (size_t)g.slot_bytes > smem_max) return g;
  g.nslots = (int)((smem_max - fixed) / g.slot_bytes);
  if (g.nslots > MAX_SLOTS) g.nslots = MAX_SLOTS;
  g.ring_bytes = g.nslots * g.slot_bytes;
  {
    // the flush parks G x BT x CQ float4 partials in the ring
    const size_t park = g.G > 2 ? (size_t)g.G * g.BT * g.CQ * 16 : 0;
    if ((size_t)g.ring_bytes < park) g.ring_bytes = (int)((park + 1023) & ~(size_t)1023);
    if (fixed + (size_t)g.ring_bytes > smem_max) return g;
  }
  const int narr = a.weightcost != 0.f ? 3 : 2;
  if ((size_t)narr * ((8 * a.ldw * 4 + 127) & ~127) > (size_t)g.ring_bytes) return g;
  size_t off = (size_t)g.ring_bytes;
  auto take = [&](size_t bytes) { size_t o = off; off += bytes; return (int)o; };
  g.off_hs = take(hs_b);
  g.off_v0 = take(slab_b);
  g.off_nv = take(slab_b);
  g.off_vt = take(vt_b);
  g.off_dred = take(dred_b);
  g.off_bars = take(128);
  g.off_misc = take(256);
  g.off_vb = take(vb_b);
  g.off_hb = take(hb_b);
  g.smem = off;
  g.ok = g.smem <= smem_max;
  return g;
}

template <int BT, bool CHAIN>
static int launch(mdbn_ctx* c, const CUtensorMap* tms, const Params& p, const Geometry& g, cudaStream_t st) {
  static bool configured[64] = {};
  auto kfn = cd_skinny_kernel<BT, CHAIN>;
  if (!configured[c->device]) {
    MDBN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[c->device] = true;
  }
  void* args[] = {(void*)&tms[0], (void*)&tms[1], (void*)&p};
  MDBN_CUDA(cudaLaunchCooperativeKernel((void*)kfn, dim3(g.grid), dim3(NT), args, g.smem, st));
  c->launches++;
  return 0;
}

}  // namespace sk

bool skinny_supported(const mdbn_ctx* c, const mdbn_cd_args& a) {
  if (a.phase != MDBN_PHASE_FULL) return false;
  return sk::plan(c, a).ok;
}

int skinny_cd_step(mdbn_ctx* c, const mdbn_cd_args& a, cudaStream_t st) { return skinny_cd_steps(c, a, 1, st); }

// n_steps consecutive steps in ONE launch: a.indices is [n_steps][B], a.cost_out [n_steps], step s draws with
// rng.offset + s (PHILOX only).  Same results as n_steps single-step calls.
int skinny_cd_steps(mdbn_ctx* c, const mdbn_cd_args& a, int n_steps, cudaStream_t st) {
  sk::Geometry g = sk::plan(c, a);
  MDBN_CHECK(n_steps >= 1, "skinny path: n_steps must be >= 1");
  MDBN_CHECK(n_steps == 1 || a.rng.mode == MDBN_RNG_PHILOX, "skinny path: multi-step launches need the PHILOX generator");
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
  p.rows_per_cta = g.rows_per_cta; p.rows_small = g.rows_small; p.n_big = g.n_big; p.rows_alloc = g.rows_alloc;
  p.CQ = g.CQ; p.GW = g.GW; p.G = g.G; p.R = g.R; p.nbox = g.nbox; p.nslots = g.nslots;
  p.ldh = g.ldh; p.slot_bytes = g.slot_bytes; p.ring_bytes = g.ring_bytes;
  p.off_hs = g.off_hs; p.off_v0 = g.off_v0; p.off_nv = g.off_nv; p.off_vt = g.off_vt; p.off_dred = g.off_dred;
  p.off_bars = g.off_bars; p.off_misc = g.off_misc; p.off_vb = g.off_vb; p.off_hb = g.off_hb;

  // scratch: two sets of 5 fixed-point accumulators [BT][ldw] + cost partials.  A launch works in one set
  // (zero on entry) and clears the other; everything is re-zeroed when the layout changes.
  const size_t n_acc = (size_t)g.BT * a.ldw;
  const size_t total_b = 2 * 5 * n_acc * sizeof(unsigned long long) + ((size_t)g.grid + 64 + n_acc) * sizeof(float);
  mdbn_ctx::Buf& wb = c->ws[WS_SKINNY];
  const void* before = wb.p;
  const size_t before_n = wb.n;
  unsigned long long* base = (unsigned long long*)ws_get(c, WS_SKINNY, total_b);
  if (!base) return 3;
  const unsigned long long key = ((unsigned long long)g.BT << 48) ^ ((unsigned long long)a.ldw << 24) ^
                                 ((unsigned long long)(uintptr_t)base << 1) ^ 1ULL;
  if (before != wb.p || before_n != wb.n || key != c->skinny_key) {
    MDBN_CUDA(cudaMemsetAsync(base, 0, wb.n, st));
    c->skinny_key = key;
    c->skinny_parity = 0;
  }
  p.n_acc = (int)n_acc;
  p.acc = base + (size_t)c->skinny_parity * 5 * n_acc;
  p.acc_other = base + (size_t)(1 - c->skinny_parity) * 5 * n_acc;
  c->skinny_parity = (c->skinny_parity + (unsigned)n_steps) & 1u;
  p.n_steps = n_steps;
  p.cost_part = reinterpret_cast<float*>(base + 2 * 5 * n_acc);
  p.PHf = p.cost_part + g.grid + 64;
  p.bar = reinterpret_cast<unsigned long long*>(c->barrier);
  static const bool want_timing = getenv("MDBN_SKINNY_TIMING") != nullptr;
  static const int dbg_flags = getenv("MDBN_SKINNY_DEBUG") ? atoi(getenv("MDBN_SKINNY_DEBUG")) : 0;
  p.dbg_flags = dbg_flags;
  p.dbg = want_timing ? reinterpret_cast<unsigned long long*>(c->barrier) + 8 : nullptr;
  CUtensorMap tms[2];
  MDBN_TRY(sk::make_map(&tms[0], a.W, a.V, a.ldw, g.R));
  MDBN_TRY(sk::make_map(&tms[1], a.W, a.V, a.ldw, 8));
  int rc = 2;
  if (g.BT == 10) rc = n_steps > 1 ? sk::launch<10, true>(c, tms, p, g, st) : sk::launch<10, false>(c, tms, p, g, st);
  else if (g.BT == 20) rc = n_steps > 1 ? sk::launch<20, true>(c, tms, p, g, st) : sk::launch<20, false>(c, tms, p, g, st);
  else set_error("skinny path: no kernel for BT=%d", g.BT);
  if (rc == 0 && p.dbg) {
    unsigned long long t[16];
    MDBN_CUDA(cudaStreamSynchronize(st));
    MDBN_CUDA(cudaMemcpy(t, p.dbg, sizeof(t), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[skinny timeline us] V=%d H=%d B=%d k=%d R=%d nslots=%d:", a.V, a.H, a.B, a.k, g.R, g.nslots);
    for (int i = 1; i < 13; ++i) fprintf(stderr, " %.1f", (double)(t[i] - t[0]) * 1e-3);
    fprintf(stderr, "\n");
    {
      static unsigned long long se[512];
      MDBN_CUDA(cudaMemcpy(se, p.dbg + 32, (256 + g.grid) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      double smin = 1e30, smax = -1e30, emin = 1e30, emax = -1e30;
      for (int i = 0; i < g.grid; ++i) {
        const double s0 = (double)((long long)(se[i] - t[0])) * 1e-3, e0 = (double)((long long)(se[256 + i] - t[0])) * 1e-3;
        smin = s0 < smin ? s0 : smin; smax = s0 > smax ? s0 : smax; emin = e0 < emin ? e0 : emin; emax = e0 > emax ? e0 : emax;
      }
      fprintf(stderr, "[skinny cta spread us] start %.1f..%.1f  end %.1f..%.1f\n", smin, smax, emin, emax);
    }
  }
  return rc;
}

}  // namespace mdbn
