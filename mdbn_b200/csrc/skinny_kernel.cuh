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
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>
#include "ctx.h"

namespace mdbn {
namespace sk {

// threads per CTA by batch tile: B <= 10 runs 12 warps (3 per scheduler, <= 170 registers), B <= 20 needs the full
// 255 registers for the hidden means of the statistics pass and stays at 8 warps
#ifndef MDBN_SK_NT10
#define MDBN_SK_NT10 256
#endif
__host__ __device__ constexpr int nt_of(int BT) { return BT <= 10 ? MDBN_SK_NT10 : 256; }
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
  float* PHf;         // [BT/2][ldw][2] positive-phase hidden means as fp32, minibatch rows 2i and 2i+1 of a column
                      // interleaved: the statistics pass loads them as ready-made FFMA2 operand pairs
  unsigned long long* bar;   // [0] barrier counter, [1] exit counter
  unsigned long long* dbg;   // optional phase timeline (MDBN_SKINNY_TIMING=1), CTA 0 only
#ifdef MDBN_SKINNY_DEBUG
  int dbg_flags;             // skip parts of the passes (timing experiments only; results are wrong)
#endif
  // smem byte offsets
  int slab_bytes;            // one [rows_alloc][BTS] slab: v0 at off_v0, nv behind it
                             // fused chains: two more slabs behind them, v0 and round(v0) of the NEXT minibatch
  int off_hs, off_v0, off_vt, off_dred, off_bars, off_misc, off_vb, off_hb;
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

// Two packed fp32 values (lo, hi) held in ONE 64-bit register, i.e. an aligned register pair: the operand form of
// FFMA2 (fma.rn.f32x2).  Typed as an integer so that ptxas keeps the pair together — float2 values are split
// into two independent registers and re-paired with MOVs in front of every packed instruction.
typedef unsigned long long P2;
__device__ __forceinline__ P2 pack2(float lo, float hi) {
  P2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo2(P2 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(P2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ P2 ffma2(P2 a, P2 b, P2 c) {
  P2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ P2 add2(P2 a, P2 b) {
  P2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ P2 sub2(P2 a, P2 b) {
  P2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ P2 mul2(P2 a, P2 b) {
  P2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// shared memory in its own state space: the address is a 32-bit register, offsets fold into the instruction
__device__ __forceinline__ ulonglong2 lds128(uint32_t a) {
  ulonglong2 v;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ P2 lds64(uint32_t a) {
  P2 v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
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
//
// FUSE (chained launches, B <= 10): the positive phase of step s+1 is computed INSIDE the statistics pass of step s,
// on the freshly updated W rows while they are still in registers (a CTA owns its rows, so the updated row is
// exactly what pass 0 of the next minibatch needs): the step then sweeps W twice instead of three times and the
// gather of the next minibatch (cp.async, issued before the Gibbs pass) is off the critical path.  The row -> thread
// assignment is by SLAB row (r mod G), the same in pass 0 and in the fused form, so both give the same bits.
#ifdef MDBN_SKINNY_DEBUG
#define SKF (p.dbg_flags)
#else
#define SKF 0
#endif
template <int BT, bool CHAIN>
__global__ void __launch_bounds__(nt_of(BT), 1) cd_skinny_kernel(const __grid_constant__ CUtensorMap tmR,
                                                                 const __grid_constant__ CUtensorMap tm8, const Params p) {
  constexpr int NT = nt_of(BT), NWARP = NT / 32;
  constexpr int BTP = (BT + 3) / 4 * 4, BTS = BTP;
  extern __shared__ __align__(1024) unsigned char smem[];
  float* hs = reinterpret_cast<float*>(smem + p.off_hs);      // [BT][ldh] chain state, zero padded to nbox*32 columns
  constexpr bool FUSE = CHAIN && BT <= 10;
  float* v0s = reinterpret_cast<float*>(smem + p.off_v0);                    // [rows_alloc][BTS]
  float* nvs = reinterpret_cast<float*>(smem + p.off_v0 + p.slab_bytes);     // [rows_alloc][BTS] (round(v0) in pass 0)
  float* v0n = reinterpret_cast<float*>(smem + p.off_v0 + 2 * p.slab_bytes);    // FUSE only: v0 of the next step
  float* xn = reinterpret_cast<float*>(smem + p.off_v0 + 3 * p.slab_bytes);     // FUSE only: its round(v0)
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
  const int F = SKF;
  int step = 0;                                  // minibatch of this launch being processed
  const int dbg_step = (CHAIN && p.n_steps > 2) ? 2 : 0;      // the step whose phases CTA 0 stamps
  auto mark = [&]() {
    if (p.dbg && cta == 0 && tid == 0 && step == dbg_step) {
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
  if (FUSE) for (int e = tid; e < p.slab_bytes; e += NT) reinterpret_cast<float*>(smem + p.off_v0)[e] = 0.f;   // 4 slabs
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
  auto fma_row = [&](const float4& w, const ulonglong2 (&v)[BTP / 4], P2 (&acc)[BTP / 2][4]) {
    const P2 wd[4] = {pack2(w.x, w.x), pack2(w.y, w.y), pack2(w.z, w.z), pack2(w.w, w.w)};
#pragma unroll
    for (int b4 = 0; b4 < BTP / 4; ++b4) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[2 * b4][c] = ffma2(v[b4].x, wd[c], acc[2 * b4][c]);
        if (4 * b4 + 2 < BT) acc[2 * b4 + 1][c] = ffma2(v[b4].y, wd[c], acc[2 * b4 + 1][c]);
      }
    }
  };
  // thread (g, q) owns the SLAB rows r with r mod G == g.  `rem` = (slab row of the tile's first row) mod G, kept
  // incrementally by the tile loops (advance_rem): no division on the tile path
  auto first_row = [&](int rem) { return g >= rem ? g - rem : g - rem + p.G; };
  const int remR = R % p.G, rem8 = 8 % p.G, rem16 = 16 % p.G;     // tile heights mod G
  auto advance_rem = [&](int& rem, int step_mod) { rem += step_mod; if (rem >= p.G) rem -= p.G; };
  auto up_tile = [&](const unsigned char* __restrict__ tile, const float* __restrict__ src,
                     const float* __restrict__ src2, int rem, int nr, P2 (&acc)[BTP / 2][4],
                     P2 (&acc2)[BTP / 2][4], bool dual) {
    if (!col_ok) return;
    const unsigned char* bp = tile + (q >> 3) * box_bytes;
    const int c = q & 7, G = p.G;
    auto ldw_ = [&](int r) { return *reinterpret_cast<const float4*>(bp + r * 128 + ((c ^ (r & 7)) << 4)); };
    int r = first_row(rem);
    for (; r + G < nr; r += 2 * G) {
      const float4 w0 = ldw_(r), w1 = ldw_(r + G);
      ulonglong2 v0[BTP / 4], v1[BTP / 4], x0[BTP / 4], x1[BTP / 4];
#pragma unroll
      for (int b4 = 0; b4 < BTP / 4; ++b4) {
        v0[b4] = reinterpret_cast<const ulonglong2*>(src + r * BTS)[b4];
        v1[b4] = reinterpret_cast<const ulonglong2*>(src + (r + G) * BTS)[b4];
      }
      if (dual) {
#pragma unroll
        for (int b4 = 0; b4 < BTP / 4; ++b4) {
          x0[b4] = reinterpret_cast<const ulonglong2*>(src2 + r * BTS)[b4];
          x1[b4] = reinterpret_cast<const ulonglong2*>(src2 + (r + G) * BTS)[b4];
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
      ulonglong2 v0[BTP / 4];
#pragma unroll
      for (int b4 = 0; b4 < BTP / 4; ++b4) v0[b4] = reinterpret_cast<const ulonglong2*>(src + r * BTS)[b4];
      fma_row(w0, v0, acc);
      if (dual) {
#pragma unroll
        for (int b4 = 0; b4 < BTP / 4; ++b4) v0[b4] = reinterpret_cast<const ulonglong2*>(src2 + r * BTS)[b4];
        fma_row(w0, v0, acc2);
      }
    }
  };

  // ---- CTA partial [B][H] -> grid-wide sum: the G row groups are added in shared memory (fixed order, hs is
  //      the staging buffer: it is rebuilt from the sums afterwards anyway), then every element goes to the
  //      fixed-point accumulator with one red.global.add.u64; CTAs start at staggered offsets ----
  auto flush_sums = [&](P2 (&acc)[BTP / 2][4], unsigned long long* dst) {
    if (p.G <= 2) {
      // wide layers (one or two row groups): add in place in hs, group after group
      float4* stage = reinterpret_cast<float4*>(hs);
      const int ldh4 = ldh >> 2;
      for (int gg = 0; gg < p.G; ++gg) {
        if (col_ok && g == gg) {
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            float4 a = (b & 1) ? make_float4(hi2(acc[b >> 1][0]), hi2(acc[b >> 1][1]), hi2(acc[b >> 1][2]), hi2(acc[b >> 1][3]))
                               : make_float4(lo2(acc[b >> 1][0]), lo2(acc[b >> 1][1]), lo2(acc[b >> 1][2]), lo2(acc[b >> 1][3]));
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
              (b & 1) ? make_float4(hi2(acc[b >> 1][0]), hi2(acc[b >> 1][1]), hi2(acc[b >> 1][2]), hi2(acc[b >> 1][3]))
                      : make_float4(lo2(acc[b >> 1][0]), lo2(acc[b >> 1][1]), lo2(acc[b >> 1][2]), lo2(acc[b >> 1][3]));
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
  if (step == dbg_step && step > 0) mark();
  // accumulator set of this step / of the previous one (alternating), recomputed where they are needed
  auto acc_set = [&]() { return (step & 1) ? p.acc_other : p.acc; };
  auto GA = [&](int s) { return acc_set() + (size_t)(2 + s % 3) * p.n_acc; };
#define A0 (acc_set())
#define A1 (acc_set() + p.n_acc)

  // later steps of a fused chain: the minibatch was gathered during the previous step
  if (FUSE && step > 0) {
    for (int e = tid; e < rows * (BTS / 4); e += NT)
      reinterpret_cast<float4*>(v0s)[e] = reinterpret_cast<const float4*>(v0n)[e];
    __syncthreads();
  }
  const int* idxp = p.idx ? p.idx + (size_t)step * B : nullptr;      // row numbers of this step's minibatch
  if (!FUSE || step == 0) {
  // ---- gather v0 slab: v0s[r][b] = data[idx[b]][row0 + r]; rows >= `rows` and b >= B are zero -----
  if (!(F & 2)) issue(0, p.nslots);     // start streaming W while the minibatch is gathered
  for (int r = tid; r < p.rows_alloc; r += NT) vbs[r] = r < rows ? __ldcg(&p.vb[row0 + r]) : 0.f;
  // the minibatch row numbers first (one dependent load for everybody), then eight gathers in flight per thread
  int* sidx = reinterpret_cast<int*>(misc) + 32;
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
    P2 acc[BTP / 2][4], acc2[BTP / 2][4];
#pragma unroll
    for (int b = 0; b < BTP / 2; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[b][c] = acc2[b][c] = 0ULL;
    int rem = 0;
    for (int j = 0, st = 0; j < ntiles; ++j, st = (st + 1 == depth ? 0 : st + 1)) {
      if (!(F & 2)) wait_stage(st);
      const unsigned char* tile = smem + (size_t)st * p.slot_bytes;
      const int nr = min(R, rows - j * R);
      if (!(F & 1)) up_tile(tile, v0s + (size_t)j * R * BTS, nvs + (size_t)j * R * BTS, rem, nr, acc, acc2, p.pcd != 0);
      advance_rem(rem, remR);
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
  }   // !FUSE || step == 0 (later steps of a fused chain got their positive phase from the previous statistics pass)
  // pseudo-likelihood monitor: the CTA that will compute it for minibatch row pl_b fetches its scalars now
  // (three dependent loads that would otherwise sit on that CTA's critical path; they overlap the barrier wait)
  if (pl_cta && tid == NT - 1) {
    const int bit = (bit0 + step) % V;                                 // src/rbm.py:445, one advance per step
    const long long dr = idxp ? idxp[pl_b] : pl_b;
    misc[60] = __int_as_float(bit);
    misc[61] = roundf(p.data[dr * p.ld_data + bit]);
    // (vb[bit] belongs to another CTA, which may still be updating it for the previous step of this launch:
    //  it is read after the first barrier of the step, below)
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
    float* dstp = p.PHf + ((size_t)(ph_b >> 1) * ldw + 4 * ph_q) * 2 + (ph_b & 1);
    __stcg(dstp, m.x); __stcg(dstp + 2, m.y); __stcg(dstp + 4, m.z); __stcg(dstp + 6, m.w);
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

  // FUSE: the minibatch of the NEXT step is gathered into the other slab set while this step's Gibbs passes run
  // (cp.async, element by element: v0n[r][b] = data[idx[b]][row0 + r]); it is consumed by the statistics pass
  const bool fuse_next = FUSE && step + 1 < n_steps;
  if (fuse_next) {
    const int* idxn = p.idx ? p.idx + (size_t)(step + 1) * B : nullptr;
    const int n = B * rows;
    for (int e = tid; e < n; e += NT) {
      const int b = e / rows, r = e - b * rows;
      const long long dr = idxn ? __ldg(idxn + b) : b;
      const float* src = p.data + dr * p.ld_data + row0 + r;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(v0n + r * BTS + b)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
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
    P2 acc[BTP / 2][4];
#pragma unroll
    for (int b = 0; b < BTP / 2; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[b][c] = 0ULL;

    int rem = 0;
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
          // packed FMAs: an accumulator pair holds the partial sums over the even and the odd columns of a row
          P2 da2[BT], db2[BT];
#pragma unroll
          for (int b = 0; b < BT; ++b) da2[b] = db2[b] = 0ULL;
          const uint32_t s_hs = smem_u32(hs);
          // Work unit = one PAIR of adjacent 16-byte chunks of a box (its half-warps take one chunk each, so the two h
          // addresses of a broadcast load fall into one 128-byte line: one wavefront; different boxes would conflict).
          // The 4 * nbox units go round-robin over the warps: 52 units on 8 warps = 7,7,7,7,6,6,6,6 — and every
          // scheduler (warps w and w+4) gets 13, where whole boxes gave 16 / 12 / 12 / 12.
          // NOT unrolled: straight-line code this long is bound by instruction fetch (ncu: stall_no_instruction)
#pragma unroll 1
          for (int u = warp; u < 4 * nbox; u += NWARP) {
            const int bx = u >> 2, c2 = u & 3;
            const uint32_t bpa = smem_u32(tile) + bx * box_bytes + lr * 128;
            const uint32_t hb0 = s_hs + bx * 128;
            {
              const int c = 2 * c2 + half;
              const uint32_t sw = (c ^ (lr & 7)) << 4;                             // rows l and l+16 swizzle alike
              const ulonglong2 wa = lds128(bpa + sw), wb = lds128(bpa + 16 * 128 + sw);
#pragma unroll
              for (int b = 0; b < BT; ++b) {
                const ulonglong2 h4 = lds128(hb0 + (b * ldh + c * 4) * 4);          // broadcast per half-warp
                da2[b] = ffma2(h4.x, wa.x, ffma2(h4.y, wa.y, da2[b]));
                db2[b] = ffma2(h4.x, wb.x, ffma2(h4.y, wb.y, db2[b]));
              }
            }
          }
          float da[BT], db[BT];
#pragma unroll
          for (int b = 0; b < BT; ++b) { da[b] = lo2(da2[b]) + hi2(da2[b]); db[b] = lo2(db2[b]) + hi2(db2[b]); }
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
      if (!(F & 16)) up_tile(tile, vt, vt, rem, nr, acc, acc, false);
      advance_rem(rem, remR);
      __syncthreads();
      if (!(F & 32)) issue(j + depth, depth);
    }
    if (last) mark();   // Gibbs tiles done
    if (last && fuse_next) {
      // the gather of the next minibatch has long landed: round(v0) for its pseudo-likelihood sums (the block
      // barriers of the flush below publish both slabs to the statistics pass)
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
      if (p.pcd) for (int e = tid; e < rows * BTS; e += NT) xn[e] = roundf(v0n[e]);     // src/rbm.py:428
    }
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

  // =============================== statistics + update (+ positive phase of the next step) =======================
  {
    // short tiles for this pass: it streams 2-3 arrays and wants a deep pipeline (16 rows when three stages of
    // them fit: half as many tile barriers, else 8)
    const int narr = p.wc != 0.f ? 3 : 2;
    const int TRS = (p.nslots * p.slot_bytes) / (narr * ((16 * ldw * 4 + 127) & ~127)) >= 3 ? 16 : 8;
    const int slot_s = (TRS * ldw * 4 + 127) & ~127;
    int depth = (p.nslots * p.slot_bytes) / (narr * slot_s);
    depth = depth > MAX_SLOTS ? MAX_SLOTS : depth;
    const int ntiles_s = (rows + TRS - 1) / TRS;
    if (warp == 0) for (int j = 0; j < depth; ++j) issue_rows(j, narr, depth, TRS, slot_s);
    const int ldw4x = ldw >> 2;
    // next minibatch (FUSE): the cp.async gather has long landed; round(v0) for the pseudo-likelihood sums
    // hidden means of the minibatch rows b and b+1 of one column travel as a pair: packed FFMA2 with the
    // (v[b], v[b+1]) pairs of the slabs, even and odd rows summed at the end -> 8 independent chains of BT/2
    P2 ph2[BT / 2][4], nh2[BT / 2][4];      // nh2 holds -nh: one packed add joins the two chains
#pragma unroll
    for (int b2 = 0; b2 < BT / 2; ++b2) {
      const ulonglong2* src = reinterpret_cast<const ulonglong2*>(p.PHf + ((size_t)b2 * ldw + 4 * q) * 2);
      const ulonglong2 lo = col_ok ? __ldcg(src) : make_ulonglong2(0ULL, 0ULL);
      const ulonglong2 hi = col_ok ? __ldcg(src + 1) : make_ulonglong2(0ULL, 0ULL);
      ph2[b2][0] = lo.x; ph2[b2][1] = lo.y; ph2[b2][2] = hi.x; ph2[b2][3] = hi.y;
    }
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
      float4 m4[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b)
        m4[b] = (col_ok && b < B)
                    ? make_float4(sigmoid_fast_(from_fixed(t[b][0].x) + hb4.x), sigmoid_fast_(from_fixed(t[b][0].y) + hb4.y),
                                  sigmoid_fast_(from_fixed(t[b][1].x) + hb4.z), sigmoid_fast_(from_fixed(t[b][1].y) + hb4.w))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b2 = 0; b2 < BT / 2; ++b2) {
        nh2[b2][0] = pack2(-m4[2 * b2].x, -m4[2 * b2 + 1].x); nh2[b2][1] = pack2(-m4[2 * b2].y, -m4[2 * b2 + 1].y);
        nh2[b2][2] = pack2(-m4[2 * b2].z, -m4[2 * b2 + 1].z); nh2[b2][3] = pack2(-m4[2 * b2].w, -m4[2 * b2 + 1].w);
      }
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
      for (int b2 = 0; b2 < BT / 2; ++b2) {
        const float4 m0 = col_ok ? *reinterpret_cast<const float4*>(hs + (2 * b2) * ldh + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 m1 = col_ok ? *reinterpret_cast<const float4*>(hs + (2 * b2 + 1) * ldh + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        nh2[b2][0] = pack2(-m0.x, -m1.x); nh2[b2][1] = pack2(-m0.y, -m1.y);
        nh2[b2][2] = pack2(-m0.z, -m1.z); nh2[b2][3] = pack2(-m0.w, -m1.w);
      }
      __syncthreads();      // hs is the staging buffer of the fused flush below
    }
    // positive-phase sums of the next step (FUSE): same thread -> (rows, column quad) assignment and the same
    // packed FMAs as pass 0 (the weight enters FFMA2 as a broadcast scalar), on the updated row still in registers
    constexpr int NF = BTP / 2;     // (dead registers in the instantiations that never fuse)
    P2 facc[NF][4], facc2[NF][4];
#pragma unroll
    for (int b = 0; b < NF; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) facc[b][c] = facc2[b][c] = 0ULL;
    const int ncol = min(4, H - 4 * q);
    const P2 k_bnom = pack2(p.inv_bnom, p.inv_bnom), k_nwc = pack2(-p.wc, -p.wc), k_c1 = pack2(p.c1, p.c1),
             k_decay = pack2(p.decay, p.decay), k_mom = pack2(p.mom, p.mom), k_lr = pack2(p.lr, p.lr);
    mark();   // statistics prologue done
    long long t_wait = 0, t_sync = 0;
    // The tile loop, specialised at compile time on what the step needs (lambda_1 term, weight cost, fused positive
    // phase, pseudo-likelihood sums): no per-row branches, and every address of a row is one register plus an
    // immediate — shared memory is addressed in its own state space, global rows by bumping two pointers.
    const uint32_t s_v0 = smem_u32(v0s), s_ring = smem_u32(smem);
    const uint32_t o_nv = (uint32_t)p.slab_bytes;
    auto tiles = [&](auto L1c, auto WCc, auto FNc, auto DUc) {
      constexpr bool L1 = decltype(L1c)::value, WC = decltype(WCc)::value, FN = decltype(FNc)::value,
                     DU = decltype(DUc)::value;
      const uint32_t row_w = (uint32_t)p.G * (uint32_t)ldw * 4u, row_v = (uint32_t)p.G * BTS * 4u;
      const size_t row_g = (size_t)p.G * ldw;
      const ptrdiff_t s_off = p.S - p.W;
      int rem = 0;
      for (int j = 0, stg = 0; j < ntiles_s; ++j, stg = (stg + 1 == depth ? 0 : stg + 1)) {
        const long long c0 = p.dbg ? clock64() : 0;
        wait_stage(stg);
        if (p.dbg) t_wait += clock64() - c0;
        const int nr = min(TRS, rows - j * TRS);
        if (col_ok) {
          int r = first_row(rem);
          uint32_t a_w = s_ring + (uint32_t)stg * narr * slot_s + ((uint32_t)r * ldw + 4u * q) * 4u;
          uint32_t a_v = s_v0 + (uint32_t)(j * TRS + r) * BTS * 4u;
          float* gw = p.W + (size_t)(row0 + j * TRS + r) * ldw + 4 * q;
          for (; r < nr; r += p.G, a_w += row_w, a_v += row_v, gw += row_g) {
            const ulonglong2 w = lds128(a_w), sp = lds128(a_w + slot_s);
            P2 gp[4], gn[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) gp[c] = gn[c] = 0ULL;
            if (!(F & 64))
#pragma unroll
            for (int b4 = 0; b4 < BTP / 4; ++b4) {
              // (the last quarter of a 10- or 20-row slab line holds one pair: a 64-bit load)
              const bool half = 4 * b4 + 2 >= BT;
              ulonglong2 av, nv;
              if (half) { av.x = lds64(a_v + 16 * b4); nv.x = lds64(a_v + o_nv + 16 * b4); av.y = nv.y = 0ULL; }
              else { av = lds128(a_v + 16 * b4); nv = lds128(a_v + o_nv + 16 * b4); }
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                gp[c] = ffma2(av.x, ph2[2 * b4][c], gp[c]);
                gn[c] = ffma2(nv.x, nh2[2 * b4][c], gn[c]);
              }
              if (!half) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  gp[c] = ffma2(av.y, ph2[2 * b4 + 1][c], gp[c]);
                  gn[c] = ffma2(nv.y, nh2[2 * b4 + 1][c], gn[c]);
                }
              }
            }
            // v0^T ph - nv^T nh of the four columns, as two column pairs
            P2 g2[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) g2[c] = add2(gp[c], gn[c]);
            const P2 gsum[2] = {pack2(lo2(g2[0]) + hi2(g2[0]), lo2(g2[1]) + hi2(g2[1])),
                                pack2(lo2(g2[2]) + hi2(g2[2]), lo2(g2[3]) + hi2(g2[3]))};
            const P2 wv[2] = {w.x, w.y}, sv[2] = {sp.x, sp.y};
            P2 wo[2], so[2];
            ulonglong2 snap = make_ulonglong2(0ULL, 0ULL);
            if (WC) snap = lds128(a_w + 2 * slot_s);
#pragma unroll
            for (int cp = 0; cp < 2; ++cp) {
              P2 gw2 = mul2(gsum[cp], k_bnom);                                 // src/rbm.py:411-415
              if (WC) gw2 = ffma2(cp ? snap.y : snap.x, k_nwc, gw2);
              P2 mult = k_decay;
              if (L1 && !(F & 128)) {
                // 1/D, D = 1 + 2 lr lambda_1 / (|W| + eps)  ==  (|W| + eps) / (|W| + eps + 2 lr lambda_1):
                // ONE MUFU reciprocal (<= 1 ulp) per element instead of two IEEE divisions          :347-350
                const float t0 = fabsf(lo2(wv[cp])) + 0.001f, t1 = fabsf(hi2(wv[cp])) + 0.001f;
                const P2 t = pack2(t0, t1), u = add2(t, k_c1);
                const P2 invD = mul2(t, pack2(rcp_approx(lo2(u)), rcp_approx(hi2(u))));
                gw2 = mul2(gw2, invD);
                mult = mul2(mult, invD);                                       // :353-356
              }
              so[cp] = ffma2(sub2(sv[cp], gw2), k_mom, gw2);                   // :361
              wo[cp] = ffma2(wv[cp], mult, mul2(sv[cp], k_lr));                // :364 (OLD speed)
            }
            if (F & 256) {
            } else if (ncol == 4) {
              *reinterpret_cast<ulonglong2*>(gw) = make_ulonglong2(wo[0], wo[1]);
              *reinterpret_cast<ulonglong2*>(gw + s_off) = make_ulonglong2(so[0], so[1]);
            } else {
              const float wf[4] = {lo2(wo[0]), hi2(wo[0]), lo2(wo[1]), hi2(wo[1])};
              const float sf[4] = {lo2(so[0]), hi2(so[0]), lo2(so[1]), hi2(so[1])};
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c < ncol) { gw[c] = wf[c]; gw[s_off + c] = sf[c]; }
              // padding columns of the quad: keep them zero in the fused positive phase below
              wo[0] = pack2(lo2(wo[0]), ncol > 1 ? hi2(wo[0]) : 0.f);
              wo[1] = pack2(ncol > 2 ? lo2(wo[1]) : 0.f, 0.f);
            }
            if constexpr (FUSE && FN) {
              if (!(F & 512)) {
                // the updated row is still in registers: v0(next) . W' for this row
                const float4 wn = make_float4(lo2(wo[0]), hi2(wo[0]), lo2(wo[1]), hi2(wo[1]));
                ulonglong2 vn[BTP / 4];
#pragma unroll
                for (int b4 = 0; b4 < BTP / 4; ++b4) {
                  if (4 * b4 + 2 >= BT) { vn[b4].x = lds64(a_v + 2 * o_nv + 16 * b4); vn[b4].y = 0ULL; }
                  else vn[b4] = lds128(a_v + 2 * o_nv + 16 * b4);
                }
                fma_row(wn, vn, facc);
                if (DU) {
#pragma unroll
                  for (int b4 = 0; b4 < BTP / 4; ++b4) {
                    if (4 * b4 + 2 >= BT) { vn[b4].x = lds64(a_v + 3 * o_nv + 16 * b4); vn[b4].y = 0ULL; }
                    else vn[b4] = lds128(a_v + 3 * o_nv + 16 * b4);
                  }
                  fma_row(wn, vn, facc2);
                }
              }
            }
          }
        }
        advance_rem(rem, TRS == 16 ? rem16 : rem8);
        const long long c1 = p.dbg ? clock64() : 0;
        __syncthreads();
        if (p.dbg) t_sync += clock64() - c1;
        if (warp == 0) issue_rows(j + depth, narr, depth, TRS, slot_s);
      }
    };
    {
      using T = std::true_type;
      using N = std::false_type;
      const bool l1 = p.c1 != 0.f, wc = narr > 2, du = p.pcd != 0;
      auto pick2 = [&](auto L1c, auto WCc) {
        if constexpr (FUSE) {
          if (fuse_next) {
            if (du) tiles(L1c, WCc, T{}, T{});
            else tiles(L1c, WCc, T{}, N{});
            return;
          }
        }
        tiles(L1c, WCc, N{}, N{});
      };
      if (l1 && !wc) pick2(T{}, N{});
      else if (!l1 && wc) pick2(N{}, T{});
      else if (l1 && wc) pick2(T{}, T{});
      else pick2(N{}, N{});
    }
    mark();   // statistics tiles done
    if (p.dbg && cta == 0 && tid == 0 && step == dbg_step) { p.dbg[24] = (unsigned long long)t_wait; p.dbg[25] = (unsigned long long)t_sync; }
    // visible bias (rows owned by this CTA)  src/rbm.py:417
    for (int r = tid; r < rows; r += NT) {
      float gsum = 0.f;
      for (int b = 0; b < B; ++b) gsum += v0s[r * BTS + b] - nvs[r * BTS + b];
      float gb = gsum * p.inv_b, sv = p.Svb[row0 + r];
      p.Svb[row0 + r] = gb + (sv - gb) * p.mom;
      const float nvb = vbs[r] + sv * p.lr;
      p.vb[row0 + r] = nvb;
      if (FUSE) vbs[r] = nvb;      // the next step of a fused chain does not reload its rows of vb
    }
    // hidden bias  src/rbm.py:416 — one CTA (the last: it owns the fewest rows), from the means already in
    // the registers of row group 0
    if (cta == (int)gridDim.x - 1 && col_ok && g == 0) {
      float gs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int b2 = 0; b2 < BT / 2; ++b2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (2 * b2 < B) gs4[c] += lo2(ph2[b2][c]) + lo2(nh2[b2][c]);
          if (2 * b2 + 1 < B) gs4[c] += hi2(ph2[b2][c]) + hi2(nh2[b2][c]);
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
    mark();   // stats + update done
    // the rows this CTA just wrote with ordinary stores are read by TMA (async proxy) in the next step
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if constexpr (FUSE) {
      if (fuse_next) {
        // positive-phase sums of the next step -> its accumulator set; this is the first barrier of that step
        unsigned long long* an = ((step + 1) & 1) ? p.acc_other : p.acc;
        flush_sums(facc, an);
        grid_arrive(p.bar, bar_target);
        if (p.pcd) flush_sums(facc2, an + p.n_acc);
      }
    }
  }
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

struct Geometry {
  int BT, rows_per_cta, rows_small, n_big, rows_alloc, CQ, GW, G, R, nbox, nslots, grid, ldh, slot_bytes, ring_bytes;
  int slab_bytes, dslab_bytes;
  int off_hs, off_v0, off_vt, off_dred, off_bars, off_misc, off_vb, off_hb;
  size_t smem;
  bool ok;
};


template <int BT, bool CHAIN>
int launch(mdbn_ctx* c, const CUtensorMap* tms, const Params& p, const Geometry& g, cudaStream_t st) {
  static bool configured[64] = {};
  auto kfn = cd_skinny_kernel<BT, CHAIN>;
  if (!configured[c->device]) {
    MDBN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[c->device] = true;
  }
  void* args[] = {(void*)&tms[0], (void*)&tms[1], (void*)&p};
  MDBN_CUDA(cudaLaunchCooperativeKernel((void*)kfn, dim3(g.grid), dim3(nt_of(BT)), args, g.smem, st));
  c->launches++;
  return 0;
}

}  // namespace sk
}  // namespace mdbn
