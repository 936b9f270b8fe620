// Skinny-batch CD-k / PCD-k step as ONE persistent cooperative kernel (B <= 20).
//
// Regime: batch 10-20 on 10^2..2*10^4-wide layers is 0.36*B flop/byte ->
// bound by streaming W, not by math.  Design:
//   * every CTA owns a contiguous slab of W rows (visible units) for the whole step;
//   * W row tiles are staged in shared memory by 1-D bulk async copies (cp.async.bulk ->
//     UBLKCP) completing on mbarriers, a ring of 16 KB slots;
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

constexpr int NT = 256;
constexpr int NWARP = NT / 32;
constexpr int SLOT = 16384;
constexpr int MAX_SLOTS = 9;
constexpr int MAX_TR = 64;
constexpr int DRED_FLOATS = 4096;

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
  int rows_per_cta, n_active, CQ, GW, G, NWG, TR, nslots;
  // global scratch
  float* part;        // [n_active][2*BT*ldw]
  float *PH, *NH, *HS, *PREX;   // [BT][ldw], zero-initialised, padded columns never written
  float* cost_part;   // [gridDim]
  unsigned long long* bar;   // [0] barrier counter, [1] exit counter
  unsigned long long* dbg;   // optional phase timeline (MDBN_SKINNY_TIMING=1), CTA 0 only
  // smem byte offsets
  int off_hs, off_v0, off_nv, off_vt, off_dred, off_bars, off_misc;
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

__host__ __device__ constexpr int half_up(int n) { return (n + 1) / 2; }
__host__ __device__ constexpr int lvl_size(int n, int l) { return l == 0 ? n : half_up(lvl_size(n, l - 1)); }

// Sum N per-lane values across the 32 lanes by recursive halving: ~N shuffles in total
// instead of 5N.  Afterwards the lane holds `n` finished sums for indices base..base+n-1.
template <int N, int L>
struct HalvingLevel {
  static __device__ __forceinline__ void run(float (&x)[N], int lane, int& base, int& n) {
    constexpr int nl = lvl_size(N, L), cnt = half_up(nl), off = 16 >> L;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      float lo = x[i];
      float hi = (cnt + i < nl) ? x[cnt + i] : 0.f;
      float send = upper ? lo : hi;
      float keep = upper ? hi : lo;
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    if (upper) { base += cnt; n = max(n - cnt, 0); } else { n = min(n, cnt); }
    if constexpr (L < 4) HalvingLevel<N, L + 1>::run(x, lane, base, n);
  }
};
template <int N>
__device__ __forceinline__ void warp_halving_sum(float (&x)[N], int lane, int& base, int& n) {
  base = 0;
  n = N;
  HalvingLevel<N, 0>::run(x, lane, base, n);
}

struct Ring {
  unsigned char* base;
  uint64_t* bars;
  uint32_t phase_bits;
};

template <int BT>
struct Cfg {
  static constexpr int BTP = (BT + 3) / 4 * 4;
  static constexpr int R = (BT <= 10) ? 4 : (BT <= 16 ? 2 : 1);   // rows per propdown chunk (register budget)
  static constexpr int N = R * BT;                  // values per halving reduction
  static constexpr int NFIN = lvl_size(N, 5);
};

template <int BT>
__global__ void __launch_bounds__(NT, 1) cd_skinny_kernel(const Params p) {
  using C = Cfg<BT>;
  constexpr int BTP = C::BTP;
  extern __shared__ __align__(1024) unsigned char smem[];
  float* hs = reinterpret_cast<float*>(smem + p.off_hs);      // [BT][ldw] chain state
  float* v0s = reinterpret_cast<float*>(smem + p.off_v0);     // [rows][BTP]
  float* nvs = reinterpret_cast<float*>(smem + p.off_nv);     // [rows][BTP]
  float* vt = reinterpret_cast<float*>(smem + p.off_vt);      // [TR][BTP] visible tile -> propup input
  float* dred = reinterpret_cast<float*>(smem + p.off_dred);  // propdown cross-warp partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);
  float* misc = reinterpret_cast<float*>(smem + p.off_misc);  // [64]: block_sum scratch, pl cost

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;
  const int ldw = p.ldw, ldw4 = ldw >> 2;
  const int B = p.B, V = p.V, H = p.H;
  const int row0 = cta * p.rows_per_cta;
  const int rows = max(0, min(p.rows_per_cta, V - row0));
  const int ntiles = (rows + p.TR - 1) / p.TR;
  // thread -> (row group g, column quad q)
  const int q = tid % p.GW, g = tid / p.GW;
  const bool grp_ok = g < p.G;                 // warp-uniform when GW >= 32
  const bool col_ok = grp_ok && q < p.CQ;
  const int wg = (p.GW >= 32) ? (warp % (p.GW >> 5)) : 0;   // warp index inside its group
  unsigned long long bar_target = 0;
  int dbg_i = 0;
  auto mark = [&]() {
    if (p.dbg && cta == 0 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[dbg_i++] = t;
    }
  };
  mark();

  Ring ring{smem, bars, 0u};
  if (tid == 0) {
    for (int i = 0; i < MAX_SLOTS; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- tile pipeline helpers -------------------------------------------------
  // job j of a pass loads `narr` arrays (W [, S [, Wsnap]]) of tile j into consecutive slots.
  auto issue = [&](int j, int narr, int depth) {      // thread 0 only; jobs are issued in order j = 0,1,2,...
    if (j >= ntiles) return;
    int st = j % depth;
    int r0 = j * p.TR, nr = min(p.TR, rows - r0);
    uint32_t bytes = (uint32_t)nr * ldw * 4u;
    uint64_t* bar = &ring.bars[st];
    mbar_expect_tx(bar, bytes * narr);
    size_t goff = (size_t)(row0 + r0) * ldw;
    unsigned char* dst = ring.base + (size_t)st * narr * SLOT;
    bulk_g2s(dst, p.W + goff, bytes, bar);
    if (narr > 1) bulk_g2s(dst + SLOT, p.S + goff, bytes, bar);
    if (narr > 2) bulk_g2s(dst + 2 * SLOT, p.Wsnap + goff, bytes, bar);
  };
  auto wait_stage = [&](int st) {
    mbar_wait(&ring.bars[st], (ring.phase_bits >> st) & 1u);
    ring.phase_bits ^= (1u << st);
  };

  // ---- randomness -------------------------------------------------------------
  auto seg = [&](long long off, uint32_t ordinal) {
    RngSeg s;
    s.mode = p.rng_mode;
    s.seg = p.ubuf ? p.ubuf + off : nullptr;
    s.k0 = p.k0; s.k1 = p.k1; s.c1 = ordinal; s.c2 = p.c2; s.c3 = p.c3;
    return s;
  };

  // ---- gather v0 slab: v0s[r][b] = data[idx[b]][row0 + r] ---------------------------
  for (int e = tid; e < rows * BTP; e += NT) {
    int b = e / rows, r = e % rows;
    float x = 0.f;
    if (b < B) {
      long long dr = p.idx ? p.idx[b] : b;
      x = p.data[dr * p.ld_data + row0 + r];
    }
    v0s[r * BTP + b] = x;
    if (p.pcd) nvs[r * BTP + b] = roundf(x);   // src/rbm.py:428; slab is free until the last Gibbs step
  }
  __syncthreads();

  // ---- propup accumulation of one staged tile: acc[b] += src[r][b] * W[r, 4q..4q+3] -------
  // DUAL: a second input slab (round(v0), pseudo-likelihood) shares the W loads.
  auto up_tile = [&](const float4* __restrict__ tile, const float* __restrict__ src, const float* __restrict__ src2,
                     int nr, float4 (&acc)[BT], float4 (&acc2)[BT], bool dual) {
    if (!col_ok) return;
    for (int r = g; r < nr; r += p.G) {
      float4 w = tile[r * ldw4 + q];
      const float4* vr = reinterpret_cast<const float4*>(src + r * BTP);
#pragma unroll
      for (int b4 = 0; b4 < BTP / 4; ++b4) {
        float4 vv = vr[b4];
        float xs[4] = {vv.x, vv.y, vv.z, vv.w};
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
        const float4* xr = reinterpret_cast<const float4*>(src2 + r * BTP);
#pragma unroll
        for (int b4 = 0; b4 < BTP / 4; ++b4) {
          float4 vv = xr[b4];
          float xs[4] = {vv.x, vv.y, vv.z, vv.w};
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
    for (int gg = 0; gg < p.G; ++gg) {
      if (col_ok && g == gg) {
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          float4 a = acc[b];
          if (gg > 0) {
            float4 o = stage[b * ldw4 + q];
            a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
          }
          stage[b * ldw4 + q] = a;
        }
      }
      __syncthreads();
    }
    if (rows > 0) {
      float4* dst = reinterpret_cast<float4*>(p.part + ((size_t)cta * 2 + set) * BT * ldw);
      for (int e = tid; e < BT * p.CQ; e += NT) {
        int b = e / p.CQ, qq = e % p.CQ;
        __stcg(&dst[b * ldw4 + qq], stage[b * ldw4 + qq]);
      }
    }
    __syncthreads();
  };

  // ---- distributed reduction of the hidden pre-activations + epilogue ----------------
  //  set 0: pre = sum + hb -> mean (sigmoid) -> mean_out, sample -> HS (and P on the last PCD step)
  //  set 1: PREX = sum + hb (pre-activation of round(v0), pseudo-likelihood)
  auto reduce_hidden = [&](int nsets, float* mean_out, const RngSeg& rs, bool write_hs, bool write_p) {
    const int nq = nsets * BT * p.CQ;
    const int per = (nq + gridDim.x - 1) / gridDim.x;
    const int o0 = cta * per, o1 = min(nq, o0 + per);
    for (int o = o0 + warp; o < o1; o += NWARP) {
      int set = o / (BT * p.CQ), rem = o % (BT * p.CQ);
      int b = rem / p.CQ, qq = rem % p.CQ;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = lane; c < p.n_active; c += 32) {
        const float4* src = reinterpret_cast<const float4*>(p.part + ((size_t)c * 2 + set) * BT * ldw);
        float4 t = __ldcg(&src[b * ldw4 + qq]);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
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
    for (int e = tid; e < BT * ldw; e += NT) {
      int b = e / ldw, j = e % ldw;
      float x = 0.f;
      if (b < nrows_src && j < H) x = __ldcg(&src[(size_t)b * ld_src + j]);
      hs[e] = x;
    }
    __syncthreads();
  };

  mark();   // gather done
  // =============================== pass 0: positive phase ===============================
  {
    const int depth = p.nslots;
    if (tid == 0) for (int j = 0; j < depth; ++j) issue(j, 1, depth);
    float4 acc[BT], acc2[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = acc2[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0, st = 0; j < ntiles; ++j, st = (st + 1 == depth ? 0 : st + 1)) {
      wait_stage(st);
      const float4* tile = reinterpret_cast<const float4*>(ring.base + (size_t)st * SLOT);
      int nr = min(p.TR, rows - j * p.TR);
      up_tile(tile, v0s + (size_t)j * p.TR * BTP, nvs + (size_t)j * p.TR * BTP, nr, acc, acc2, p.pcd != 0);
      __syncthreads();
      if (tid == 0) issue(j + depth, 1, depth);
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
  if (p.pcd) load_hs(p.P, H, B); else load_hs(p.HS, ldw, BT);
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
    if (tid == 0) for (int j = 0; j < depth; ++j) issue(j, 1, depth);
    // chain state of this thread's columns
    float4 hreg[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b)
      hreg[b] = col_ok ? reinterpret_cast<const float4*>(hs)[b * ldw4 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int j = 0, st = 0; j < ntiles; ++j, st = (st + 1 == depth ? 0 : st + 1)) {
      wait_stage(st);
      const float4* tile = reinterpret_cast<const float4*>(ring.base + (size_t)st * SLOT);
      const int nr = min(p.TR, rows - j * p.TR);
      // ---- propdown of the tile rows: partial dot over this thread's 4 columns, then across lanes
      if (grp_ok) {
        // uniform trip count for every lane: the reductions below are warp-synchronous
        for (int rb0 = 0; rb0 < nr; rb0 += p.G * C::R) {
          const int rb = rb0 + g * C::R;
          float x[C::N];
#pragma unroll
          for (int rr = 0; rr < C::R; ++rr) {
            int r = rb + rr;
            float4 w = (col_ok && r < nr) ? tile[r * ldw4 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int b = 0; b < BT; ++b)
              x[rr * BT + b] = fmaf(hreg[b].x, w.x, fmaf(hreg[b].y, w.y, fmaf(hreg[b].z, w.z, hreg[b].w * w.w)));
          }
          if (p.GW >= 32) {
            int base, n;
            warp_halving_sum<C::N>(x, lane, base, n);
#pragma unroll
            for (int t = 0; t < C::NFIN; ++t) {
              if (t < n) {
                int id = base + t, rr = id / BT, b = id % BT, r = rb + rr;
                if (r < nr) dred[(r * BT + b) * p.NWG + wg] = x[t];
              }
            }
          } else {
            // narrow layers: several row groups share a warp -> butterfly inside the GW lanes
            for (int off = p.GW >> 1; off > 0; off >>= 1) {
#pragma unroll
              for (int i = 0; i < C::N; ++i) x[i] += __shfl_xor_sync(0xffffffffu, x[i], off);
            }
            if (q == 0) {
#pragma unroll
              for (int i = 0; i < C::N; ++i) {
                int rr = i / BT, b = i % BT, r = rb + rr;
                if (r < nr) dred[(r * BT + b)] = x[i];
              }
            }
          }
        }
      }
      __syncthreads();
      // ---- visible epilogue: bias, activation, sampling (src/rbm.py:226-240 / :650-660) ----
      for (int it = tid; it < nr * BT; it += NT) {
        int r = it / BT, b = it % BT;
        float sum = 0.f;
        for (int w2 = 0; w2 < p.NWG; ++w2) sum += dred[it * p.NWG + w2];
        int gi = row0 + j * p.TR + r;
        float pre = sum + p.vb[gi];
        float mean = 0.f, vin = 0.f;
        if (b < B) {
          if (p.kind == MDBN_GRBM) {
            mean = pre;
            vin = pre;        // mean-field visible: h given v_MEAN (src/rbm.py:669)
          } else {
            mean = sigmoidf_(pre);
            vin = rng_uniform(rs_v, (long long)b * V + gi) < mean ? 1.f : 0.f;
          }
          if (last && !p.pcd) {
            float t = v0s[(j * p.TR + r) * BTP + b];
            if (p.kind == MDBN_GRBM) { float d = sigmoidf_(pre) - t; cost_acc += d * d; }   // :697
            else cost_acc += t * softplusf_(-pre) + (1.f - t) * softplusf_(pre);          // :479-480
          }
        }
        vt[r * BTP + b] = vin;
        if (last) nvs[(j * p.TR + r) * BTP + b] = mean;
      }
      if constexpr (BTP > BT) {
        constexpr int PADB = BTP - BT;
        for (int it = tid; it < nr * PADB; it += NT) {
          int r = it / PADB, b = BT + it % PADB;
          vt[r * BTP + b] = 0.f;
          if (last) nvs[(j * p.TR + r) * BTP + b] = 0.f;
        }
      }
      __syncthreads();
      // ---- propup accumulation from the same tile ----
      up_tile(tile, vt, vt, nr, acc, acc, false);
      __syncthreads();
      if (tid == 0) issue(j + depth, 1, depth);
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
    if (!last) load_hs(p.HS, ldw, BT);
  }

  // =============================== statistics + update ===============================
  {
    const int narr = p.wc != 0.f ? 3 : 2;
    const int depth = p.nslots / narr;
    if (tid == 0) for (int j = 0; j < depth; ++j) issue(j, narr, depth);
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
    for (int j = 0, stg = 0; j < ntiles; ++j, stg = (stg + 1 == depth ? 0 : stg + 1)) {
      wait_stage(stg);
      const unsigned char* sb = ring.base + (size_t)stg * narr * SLOT;
      const float4* wt = reinterpret_cast<const float4*>(sb);
      const float4* st = reinterpret_cast<const float4*>(sb + SLOT);
      const float4* sn = reinterpret_cast<const float4*>(sb + 2 * SLOT);
      const int nr = min(p.TR, rows - j * p.TR);
      if (col_ok) {
        for (int r = g; r < nr; r += p.G) {
          const int lr_ = j * p.TR + r;
          float4 w = wt[r * ldw4 + q], sp = st[r * ldw4 + q];
          const float4* a4 = reinterpret_cast<const float4*>(v0s + lr_ * BTP);
          const float4* n4 = reinterpret_cast<const float4*>(nvs + lr_ * BTP);
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
          if (narr > 2) { float4 t4 = sn[r * ldw4 + q]; snv[0] = t4.x; snv[1] = t4.y; snv[2] = t4.z; snv[3] = t4.w; }
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
      if (tid == 0) issue(j + depth, narr, depth);
    }
    // visible bias (rows owned by this CTA)  src/rbm.py:417
    for (int r = tid; r < rows; r += NT) {
      float gsum = 0.f;
      for (int b = 0; b < B; ++b) gsum += v0s[r * BTP + b] - nvs[r * BTP + b];
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
  int BT, rows_per_cta, n_active, CQ, GW, G, NWG, TR, nslots, grid;
  int off_hs, off_v0, off_nv, off_vt, off_dred, off_bars, off_misc;
  size_t smem;
  bool ok;
};

static int pick_bt(int B) {
  const int opts[] = {4, 8, 10, 12, 16, 20};
  for (int o : opts)
    if (B <= o) return o;
  return 0;
}

static Geometry plan(const mdbn_ctx* c, const mdbn_cd_args& a) {
  Geometry g{};
  g.ok = false;
  g.BT = pick_bt(a.B);
  if (!g.BT || a.ldw % 4 != 0) return g;
  if (((uintptr_t)a.W | (uintptr_t)a.W_speed | (uintptr_t)a.W_snap) & 15) return g;
  const int BTP = (g.BT + 3) / 4 * 4;
  g.CQ = a.ldw / 4;
  if (g.CQ > NT) return g;
  if (g.CQ <= 32) { g.GW = 1; while (g.GW < g.CQ) g.GW <<= 1; } else g.GW = (g.CQ + 31) / 32 * 32;
  g.G = NT / g.GW;
  g.NWG = g.GW >= 32 ? g.GW / 32 : 1;
  g.grid = c->num_sms;
  g.rows_per_cta = (a.V + g.grid - 1) / g.grid;
  g.n_active = (a.V + g.rows_per_cta - 1) / g.rows_per_cta;
  int tr = SLOT / (a.ldw * 4);
  if (tr < 1) return g;
  g.TR = tr > MAX_TR ? MAX_TR : tr;
  if (g.TR * g.BT * g.NWG > DRED_FLOATS) g.TR = DRED_FLOATS / (g.BT * g.NWG);
  if (g.TR < 1) return g;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 127) & ~(size_t)127; return (int)o; };
  size_t fixed_after_ring;
  // ring first (1024-aligned), sized last: compute the fixed part, then give the ring what is left
  size_t hs_b = (size_t)g.BT * a.ldw * 4, slab_b = (size_t)g.rows_per_cta * BTP * 4, vt_b = (size_t)MAX_TR * BTP * 4;
  fixed_after_ring = ((hs_b + 127) & ~127) + 2 * ((slab_b + 127) & ~127) + ((vt_b + 127) & ~127) +
                     DRED_FLOATS * 4 + 128 + 256 + 1024;
  const size_t smem_max = 227 * 1024;
  if (fixed_after_ring + 4 * SLOT > smem_max) return g;
  g.nslots = (int)((smem_max - fixed_after_ring) / SLOT);
  if (g.nslots > MAX_SLOTS) g.nslots = MAX_SLOTS;
  const int narr = a.weightcost != 0.f ? 3 : 2;
  if (g.nslots < narr) return g;
  off = (size_t)g.nslots * SLOT;
  g.off_hs = take(hs_b);
  g.off_v0 = take(slab_b);
  g.off_nv = take(slab_b);
  g.off_vt = take(vt_b);
  g.off_dred = take(DRED_FLOATS * 4);
  g.off_bars = take(128);
  g.off_misc = take(256);
  g.smem = off;
  g.ok = g.smem <= smem_max;
  return g;
}

template <int BT>
static int launch(mdbn_ctx* c, const Params& p, const Geometry& g, cudaStream_t st) {
  static bool configured[64] = {};
  auto kfn = cd_skinny_kernel<BT>;
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
  if (a.persistent && a.B != a.B_nom && false) return false;
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
  p.rows_per_cta = g.rows_per_cta; p.n_active = g.n_active; p.CQ = g.CQ; p.GW = g.GW; p.G = g.G; p.NWG = g.NWG;
  p.TR = g.TR; p.nslots = g.nslots;
  p.off_hs = g.off_hs; p.off_v0 = g.off_v0; p.off_nv = g.off_nv; p.off_vt = g.off_vt; p.off_dred = g.off_dred;
  p.off_bars = g.off_bars; p.off_misc = g.off_misc;

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
  p.dbg = want_timing ? reinterpret_cast<unsigned long long*>(c->barrier) + 8 : nullptr;
  int rc = 2;
  switch (g.BT) {
    case 4: rc = sk::launch<4>(c, p, g, st); break;
    case 8: rc = sk::launch<8>(c, p, g, st); break;
    case 10: rc = sk::launch<10>(c, p, g, st); break;
    case 12: rc = sk::launch<12>(c, p, g, st); break;
    case 16: rc = sk::launch<16>(c, p, g, st); break;
    case 20: rc = sk::launch<20>(c, p, g, st); break;
    default: set_error("skinny path: no kernel for BT=%d", g.BT);
  }
  if (rc == 0 && p.dbg) {
    unsigned long long t[16];
    MDBN_CUDA(cudaStreamSynchronize(st));
    MDBN_CUDA(cudaMemcpy(t, p.dbg, sizeof(t), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[skinny timeline us] V=%d H=%d B=%d k=%d:", a.V, a.H, a.B, a.k);
    for (int i = 1; i < 13; ++i) fprintf(stderr, " %.1f", (double)(t[i] - t[0]) * 1e-3);
    fprintf(stderr, "\n");
  }
  return rc;
}

}  // namespace mdbn
