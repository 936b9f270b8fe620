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
//   * hidden pre-activations need all rows: per-CTA partials -> global scratch -> grid
//     barrier -> each CTA reduces a slice in fixed order (deterministic) + bias + sigmoid +
//     sample -> grid barrier -> every CTA reloads the full [B,H] hidden state;
//   * last pass  : statistics + lambda_1/lambda_2/momentum update fused: W and W_speed tiles
//     are read once and written once; v0 and nv slabs never left shared memory.
// All arithmetic is plain fp32 FFMA.  HBM traffic per step: (k+1) reads of W + read W,S + write W,S
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
  int rows_per_cta, rows_alloc, n_active, CQ, GW, G, R, nbox, nslots, ldh, slot_bytes, ring_bytes;
  // global scratch
  float* part;        // [n_active][2][BT][ldw]
  float *PH, *NH, *HS, *PREX;   // [BT][ldw], zero-initialised, padded columns never written
  float* cost_part;   // [gridDim]
  unsigned long long* bar;   // [0] barrier counter, [1] exit counter
  unsigned long long* dbg;   // optional phase timeline (MDBN_SKINNY_TIMING=1), CTA 0 only
  int dbg_flags;             // MDBN_SKINNY_DEBUG: skip parts of the passes (timing experiments only; results are wrong)
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

template <int BT>
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

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;
  const int ldw = p.ldw, ldh = p.ldh, R = p.R, nbox = p.nbox;
  const int B = p.B, V = p.V, H = p.H;
  const int row0 = cta * p.rows_per_cta;
  const int rows = max(0, min(p.rows_per_cta, V - row0));
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
  for (int e = tid; e < BT * ldh; e += NT) hs[e] = 0.f;
  __syncthreads();

  // ---- tile pipeline --------------------------------------------------------------
  // W-only passes: stage j = rows [j*R, j*R+R) x all columns as nbox swizzled boxes (TMA).  lane 0 of warp 0.
  auto issue = [&](int j, int depth) {
    const int r0 = j * R;
    if (r0 >= rows || lane != 0) return;
    const int st = j % depth;
    const int nr8 = (min(R, rows - r0) + 7) & ~7;
    uint64_t* bar = &bars[st];
    unsigned char* dst = smem + (size_t)st * p.slot_bytes;
    mbar_expect_tx(bar, (uint32_t)nbox * nr8 * 128u);
    if (nr8 == R) {
      for (int bx = 0; bx < nbox; ++bx) tma_load_2d(dst + bx * box_bytes, &tmR, bar, 32 * bx, row0 + r0);
    } else {
      for (int bx = 0; bx < nbox; ++bx)
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
    s.k0 = p.k0; s.k1 = p.k1; s.c1 = ordinal; s.c2 = p.c2; s.c3 = p.c3;
    return s;
  };

  // ---- gather v0 slab: v0s[r][b] = data[idx[b]][row0 + r]; rows >= `rows` and b >= B are zero -----
  if (warp == 0 && !(F & 2)) issue(0, p.nslots);     // start streaming W while the minibatch is gathered
  for (int r = tid; r < p.rows_alloc; r += NT) vbs[r] = r < rows ? p.vb[row0 + r] : 0.f;
  for (int e = tid; e < p.rows_alloc * BTS; e += NT) {
    const int b = e / p.rows_alloc, r = e - b * p.rows_alloc;
    float x = 0.f;
    if (b < B && r < rows) {
      const long long dr = p.idx ? p.idx[b] : b;
      x = p.data[dr * p.ld_data + row0 + r];
    }
    v0s[r * BTS + b] = x;
    nvs[r * BTS + b] = p.pcd ? roundf(x) : 0.f;   // src/rbm.py:428; the nv slab is free until the last Gibbs step
  }
  __syncthreads();
  mark();   // gather done

  // ---- propup of one staged tile: acc[b] += src[r][b] * W[r, 4q..4q+3]; DUAL shares the W loads ----
  auto up_tile = [&](const unsigned char* __restrict__ tile, const float* __restrict__ src,
                     const float* __restrict__ src2, int nr, float4 (&acc)[BT], float4 (&acc2)[BT], bool dual) {
    if (!col_ok) return;
    const unsigned char* bp = tile + (q >> 3) * box_bytes;
    const int c = q & 7;
    for (int r = g; r < nr; r += p.G) {
      const float4 w = *reinterpret_cast<const float4*>(bp + r * 128 + ((c ^ (r & 7)) << 4));
      const float4* vr = reinterpret_cast<const float4*>(src + r * BTS);
#pragma unroll
      for (int b4 = 0; b4 < BTP / 4; ++b4) {
        const float4 vv = vr[b4];
        const float xs[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int b = b4 * 4 + t;
          if (b < BT) {
            acc[b].x = fmaf(xs[t], w.x, acc[b].x);
            acc[b].y = fmaf(xs[t], w.y, acc[b].y);
            acc[b].z = fmaf(xs[t], w.z, acc[b].z);
            acc[b].w = fmaf(xs[t], w.w, acc[b].w);
          }
        }
      }
      if (dual) {
        const float4* xr = reinterpret_cast<const float4*>(src2 + r * BTS);
#pragma unroll
        for (int b4 = 0; b4 < BTP / 4; ++b4) {
          const float4 vv = xr[b4];
          const float xs[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int b = b4 * 4 + t;
            if (b < BT) {
              acc2[b].x = fmaf(xs[t], w.x, acc2[b].x);
              acc2[b].y = fmaf(xs[t], w.y, acc2[b].y);
              acc2[b].z = fmaf(xs[t], w.z, acc2[b].z);
              acc2[b].w = fmaf(xs[t], w.w, acc2[b].w);
            }
          }
        }
      }
    }
  };

  // ---- CTA partial [BT][ldw]: sum over the G row groups (fixed order), then to global scratch ---
  // uses hs as the staging accumulator (it is reloaded after the reduction anyway)
  auto flush_partial = [&](float4 (&acc)[BT], int set) {
    float4* stage = reinterpret_cast<float4*>(hs);
    const int ldh4 = ldh >> 2;
    for (int gg = 0; gg < p.G; ++gg) {
      if (col_ok && g == gg) {
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          float4 a = acc[b];
          if (gg > 0) {
            const float4 o = stage[b * ldh4 + q];
            a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
          }
          stage[b * ldh4 + q] = a;
        }
      }
      __syncthreads();
    }
    if (rows > 0) {
      float4* dst = reinterpret_cast<float4*>(p.part + ((size_t)cta * 2 + set) * BT * ldw);
      const int ldw4 = ldw >> 2;
      for (int e = tid; e < BT * p.CQ; e += NT) {
        const int b = e / p.CQ, qq = e - b * p.CQ;
        __stcg(&dst[b * ldw4 + qq], stage[b * ldh4 + qq]);
      }
    }
    __syncthreads();
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

  // chain state [BT][ldw] from L2 into the shared-memory panel (float4, four loads in flight per thread)
  auto load_hs = [&](const float* src, int ld_src, int nrows_src) {
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
        }
      }
    }
    __syncthreads();
  };

  // =============================== pass 0: positive phase ===============================
  {
    const int depth = p.nslots;
    if (warp == 0 && !(F & 2)) for (int j = 1; j < depth; ++j) issue(j, depth);
    float4 acc[BT], acc2[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = acc2[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0, st = 0; j < ntiles; ++j, st = (st + 1 == depth ? 0 : st + 1)) {
      if (!(F & 2)) wait_stage(st);
      const unsigned char* tile = smem + (size_t)st * p.slot_bytes;
      const int nr = min(R, rows - j * R);
      if (!(F & 1)) up_tile(tile, v0s + (size_t)j * R * BTS, nvs + (size_t)j * R * BTS, nr, acc, acc2, p.pcd != 0);
      __syncthreads();
      if (warp == 0 && !(F & 2)) issue(j + depth, depth);
    }
    mark();   // pass-0 tiles done
    flush_partial(acc, 0);
    if (p.pcd) flush_partial(acc2, 1);
  }
  mark();
  grid_sync(p.bar, bar_target);
  mark();
  // CD: chain starts from the fresh sample; PCD: from the persistent chain (src/rbm.py:308-311)
  reduce_hidden(p.pcd ? 2 : 1, p.PH, seg(0, 0), !p.pcd, false);
  mark();   // reduce 0 done
  grid_sync(p.bar, bar_target);
  if (warp == 0 && !(F & 32)) issue(0, p.nslots);     // W is unchanged until the update: prefetch the next pass now
  if (p.pcd) load_hs(p.P, H, B); else load_hs(p.HS, ldw, BT);
  mark();

  // pseudo-likelihood monitor (src/rbm.py:421-447), pre-update W, hb, vb: one minibatch row per CTA, taken
  // from the END of the grid (the last CTA owns the fewest rows); its loads overlap the tile prefetch above
  if (p.pcd) {
    const int bit = *p.bit_idx;
    for (int b = (int)gridDim.x - 1 - cta; b < B; b += gridDim.x) {
      const long long dr = p.idx ? p.idx[b] : b;
      const float x = roundf(p.data[dr * p.ld_data + bit]);
      const float d = 1.f - 2.f * x;
      float h0 = 0.f, h1 = 0.f;
      for (int j = tid; j < H; j += NT) {
        const float pre = __ldcg(&p.PREX[b * ldw + j]);
        h0 += softplusf_(pre);
        h1 += softplusf_(pre + d * p.W[(size_t)bit * ldw + j]);
      }
      h0 = block_sum(h0, misc);
      h1 = block_sum(h1, misc);
      if (tid == 0) {
        const float vbv = p.vb[bit];
        float vterm;
        if (p.kind == MDBN_GRBM) { const float a = x - vbv, c = (1.f - x) - vbv; vterm = 0.5f * (a * a - c * c); }
        else vterm = d * vbv;
        __stcg(&p.cost_part[b], -(float)V * softplusf_((h1 - h0) + vterm));
      }
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
    if (warp == 0 && !(F & 32)) for (int j = 1; j < depth; ++j) issue(j, depth);   // job 0 was prefetched
    float4 acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int j = 0, st = 0; j < ntiles; ++j, st = (st + 1 == depth ? 0 : st + 1)) {
      if (!(F & 32)) wait_stage(st);
      const unsigned char* tile = smem + (size_t)st * p.slot_bytes;
      const int nr = min(R, rows - j * R);
      // ---- propdown of the tile rows, lane = row: out[b] = sum_j h[b][j] W[row][j] stays in one thread;
      //      the warps (and, for short tiles, the lane groups) split the boxes of 32 columns ----
      if (!(F & 4)) {
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
      __syncthreads();
      // ---- visible epilogue: bias, activation, sampling (src/rbm.py:226-240 / :650-660) ----
      for (int it = tid; it < nr * BTS && !(F & 8); it += NT) {
        const int r = it / BTS, b = it - r * BTS;
        float vin = 0.f, mean = 0.f;
        if (b < B) {
          float sum = 0.f;
          for (int w2 = 0; w2 < NWARP * SUBS; ++w2) sum += dred[((size_t)w2 * BT + b) * R + r];
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
      if (warp == 0 && !(F & 32)) issue(j + depth, depth);
    }
    if (last) mark();   // Gibbs tiles done
    flush_partial(acc, 0);
    if (last && !p.pcd) {
      const float c = block_sum(cost_acc, misc);
      if (tid == 0) __stcg(&p.cost_part[cta], c);
    }
    grid_sync(p.bar, bar_target);
    if (last) mark();
    reduce_hidden(1, p.NH, rs_h, !last, last && p.pcd);
    grid_sync(p.bar, bar_target);
    if (last) mark();
    if (!last) {
      if (warp == 0 && !(F & 32)) issue(0, p.nslots);
      load_hs(p.HS, ldw, BT);
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
    const int ldw4 = ldw >> 2, ldw4x = ldw >> 2;
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
      if (warp == 0) issue_rows(j + depth, narr, depth, TRS, slot_s);
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
        c = 0.f;
        for (int b = 0; b < B; ++b) c += __ldcg(&p.cost_part[b]);
        c *= p.cost_scale;
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
  int BT, rows_per_cta, rows_alloc, n_active, CQ, GW, G, R, nbox, nslots, grid, ldh, slot_bytes, ring_bytes;
  int off_hs, off_v0, off_nv, off_vt, off_dred, off_bars, off_misc, off_vb;
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
  g.rows_per_cta = (a.V + g.grid - 1) / g.grid;
  g.rows_alloc = (g.rows_per_cta + 7) & ~7;
  g.n_active = (a.V + g.rows_per_cta - 1) / g.rows_per_cta;
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
               vb_b = up128((size_t)g.rows_alloc * 4);
  const size_t fixed = hs_b + 2 * slab_b + vt_b + dred_b + 128 + 256 + vb_b;
  const size_t smem_max = 227 * 1024;
  if (fixed + 2 * (size_t)g.slot_bytes > smem_max) return g;
  g.nslots = (int)((smem_max - fixed) / g.slot_bytes);
  if (g.nslots > MAX_SLOTS) g.nslots = MAX_SLOTS;
  g.ring_bytes = g.nslots * g.slot_bytes;
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
  g.smem = off;
  g.ok = g.smem <= smem_max;
  return g;
}

template <int BT>
static int launch(mdbn_ctx* c, const CUtensorMap* tms, const Params& p, const Geometry& g, cudaStream_t st) {
  static bool configured[64] = {};
  auto kfn = cd_skinny_kernel<BT>;
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
  p.CQ = g.CQ; p.GW = g.GW; p.G = g.G; p.R = g.R; p.nbox = g.nbox; p.nslots = g.nslots;
  p.ldh = g.ldh; p.slot_bytes = g.slot_bytes; p.ring_bytes = g.ring_bytes;
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
  static thread_local unsigned long long last_key = 0;
  const unsigned long long key = ((unsigned long long)g.BT << 48) ^ ((unsigned long long)a.ldw << 24) ^
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
  static const int dbg_flags = getenv("MDBN_SKINNY_DEBUG") ? atoi(getenv("MDBN_SKINNY_DEBUG")) : 0;
  p.dbg_flags = dbg_flags;
  p.dbg = want_timing ? reinterpret_cast<unsigned long long*>(c->barrier) + 8 : nullptr;
  CUtensorMap tms[2];
  MDBN_TRY(sk::make_map(&tms[0], a.W, a.V, a.ldw, g.R));
  MDBN_TRY(sk::make_map(&tms[1], a.W, a.V, a.ldw, 8));
  int rc = 2;
  if (g.BT == 10) rc = sk::launch<10>(c, tms, p, g, st);
  else if (g.BT == 20) rc = sk::launch<20>(c, tms, p, g, st);
  else set_error("skinny path: no kernel for BT=%d", g.BT);
  if (rc == 0 && p.dbg) {
    unsigned long long t[16];
    MDBN_CUDA(cudaStreamSynchronize(st));
    MDBN_CUDA(cudaMemcpy(t, p.dbg, sizeof(t), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[skinny timeline us] V=%d H=%d B=%d k=%d R=%d nslots=%d:", a.V, a.H, a.B, a.k, g.R, g.nslots);
    for (int i = 1; i < 13; ++i) fprintf(stderr, " %.1f", (double)(t[i] - t[0]) * 1e-3);
    fprintf(stderr, "\n");
  }
  return rc;
}

}  // namespace mdbn
