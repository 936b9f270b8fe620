// Small layers (the upper layers of every DBN of the reference: 559->40, 400->40, 200->20, 100->24 ...):
// CD-k / PCD-k steps on ONE THREAD-BLOCK CLUSTER with the weights resident in shared memory.
//
// On such layers a step is pure latency: the persistent grid kernel (skinny.cu) spends ~9 us per Gibbs
// iteration in L2 round trips (reduction atomics, grid barrier, read-back) for a few hundred thousand FMAs.
// Here the 8 CTAs of one cluster each own V/8 rows of W and of its momentum and keep them in shared memory
// for the whole launch (an epoch of chained steps touches global memory only for the minibatch rows);
// the hidden pre-activations are all-reduced through DISTRIBUTED SHARED MEMORY (every CTA reads the eight
// partials with ld.shared::cluster and adds them in rank order: deterministic) behind one
// barrier.cluster per propagation; bias, sigmoid and the Bernoulli draw are then computed by every CTA
// for the whole [B, H] (the draws are indexed by element, so all CTAs agree).  Same arithmetic contract
// as the other paths: fp32, src/rbm.py semantics, Philox or caller-supplied uniforms.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include "ctx.h"

namespace mdbn {
namespace tn {

constexpr int NT = 512;
constexpr int CL = 8;            // CTAs per cluster (portable maximum)
constexpr int MH = 20;           // rows of M = W^T W a thread of the register edition of the local Gibbs steps holds

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
  int kind, B, V, H, k, pcd, n_steps;
  float inv_bnom, inv_b, wc, c1, decay, mom, lr, cost_scale;
  int rng_mode;
  const float* ubuf;
  uint32_t k0, k1, c2, c3;
  long long u_step_stride, u_off_v, u_off_h;
  int rows_per_cta, rows_alloc, BTS, CQ, G, lds;
  int use_m;                 // GRBM with k > 1: Gibbs steps 0..k-2 run on M = W^T W (see below)
  int off_M, off_Mp, off_U;
  unsigned long long* dbg;   // optional phase timeline (MDBN_TINY_TIMING=1), rank 0, first step
  // shared-memory byte offsets
  int off_park, off_W, off_S, off_v0, off_nv, off_vin, off_hs, off_pc, off_ph, off_nh, off_part, off_vb, off_svb, off_hb,
      off_shb, off_misc;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// float4 from the same shared-memory offset of CTA `rank` of this cluster
__device__ __forceinline__ float4 ld_remote4(const void* local_ptr, uint32_t rank) {
  uint32_t ra;
  float4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(ra)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_remote1(const void* local_ptr, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast_(float x) { return rcp_approx(1.0f + __expf(-x)); }

__global__ void __launch_bounds__(NT, 1) cd_tiny_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* Ws = reinterpret_cast<float*>(smem + p.off_W);       // [rows_alloc][lds]  resident weights of the owned rows (lds/4 odd: conflict-free row walks)
  float* Ss = reinterpret_cast<float*>(smem + p.off_S);       // [rows_alloc][lds]  resident momentum
  float* v0s = reinterpret_cast<float*>(smem + p.off_v0);     // [rows_alloc][BTS]  data
  float* nvs = reinterpret_cast<float*>(smem + p.off_nv);     // [rows_alloc][BTS]  round(v0) (PCD pass 0), then nv mean
  float* vin = reinterpret_cast<float*>(smem + p.off_vin);    // [rows_alloc][BTS]  visible input of the next propup
  float* hs = reinterpret_cast<float*>(smem + p.off_hs);      // [BTS][ldw]  chain state
  float* pc = reinterpret_cast<float*>(smem + p.off_pc);      // [BTS][ldw]  persistent chain (PCD)
  float* phs = reinterpret_cast<float*>(smem + p.off_ph);     // [BTS][ldw]  positive-phase means
  float* nhs = reinterpret_cast<float*>(smem + p.off_nh);     // [BTS][ldw]  last negative means / PL pre-activations
  float* part = reinterpret_cast<float*>(smem + p.off_part);  // [2 parity][2 sets][BTS][ldw] partial hidden sums
  float* vbs = reinterpret_cast<float*>(smem + p.off_vb);     // [rows_alloc]
  float* svbs = reinterpret_cast<float*>(smem + p.off_svb);   // [rows_alloc]
  float* hbs = reinterpret_cast<float*>(smem + p.off_hb);     // [ldw]
  float* shbs = reinterpret_cast<float*>(smem + p.off_shb);   // [ldw]
  float* Ms = reinterpret_cast<float*>(smem + p.off_M);       // [H+1][ldw]  W^T W of ALL rows; row H = vb . W
  float* Mp = reinterpret_cast<float*>(smem + p.off_Mp);      // [H+1][ldw]  this CTA's share of it
  float* Us = reinterpret_cast<float*>(smem + p.off_U);       // [k-1][B][CQ][4] uniforms of the local Gibbs steps (use_m == 2)
  float* park = reinterpret_cast<float*>(smem + p.off_park);  // [G][BTS][ldw] row-group partials of a propup
  float* misc = reinterpret_cast<float*>(smem + p.off_misc);  // [64]: block_sum scratch, [40] cost partial, [48..] sidx

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_rank();
  const int ldw = p.ldw, lds = p.lds, B = p.B, V = p.V, H = p.H, BTS = p.BTS, CQ = p.CQ;
  const int row0 = (int)rank * p.rows_per_cta;
  const int rows = max(0, min(p.rows_per_cta, V - row0));
  const int nhid = BTS * ldw;                       // floats of one [BTS][ldw] panel
  int parity = 0;
  int dbg_i = 0;
  auto mark = [&]() {
    if (p.dbg && rank == 0 && tid == 0 && dbg_i < 16) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[dbg_i++] = t;
    }
  };
  mark();

  // ---- load the resident state ----
  for (int e = tid; e < p.rows_alloc * CQ; e += NT) {
    const int r = e / CQ, q = e - r * CQ;
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f), s = w;
    if (r < rows) {
      w = *reinterpret_cast<const float4*>(p.W + (size_t)(row0 + r) * ldw + 4 * q);
      s = *reinterpret_cast<const float4*>(p.S + (size_t)(row0 + r) * ldw + 4 * q);
    }
    *reinterpret_cast<float4*>(Ws + r * lds + 4 * q) = w;
    *reinterpret_cast<float4*>(Ss + r * lds + 4 * q) = s;
  }
  for (int r = tid; r < p.rows_alloc; r += NT) {
    vbs[r] = r < rows ? p.vb[row0 + r] : 0.f;
    svbs[r] = r < rows ? p.Svb[row0 + r] : 0.f;
  }
  for (int j = tid; j < ldw; j += NT) {
    hbs[j] = j < H ? p.hb[j] : 0.f;
    shbs[j] = j < H ? p.Shb[j] : 0.f;
  }
  for (int e = tid; e < nhid; e += NT) {
    const int b = e / ldw, j = e - b * ldw;
    hs[e] = phs[e] = nhs[e] = 0.f;
    pc[e] = (p.pcd && b < B && j < H) ? p.P[(size_t)b * H + j] : 0.f;
  }
  const int bit0 = p.pcd ? *p.bit_idx : 0;
  __syncthreads();
  mark();   // state loaded

  auto seg = [&](long long off, uint32_t ordinal, int step) {
    RngSeg s;
    const unsigned long long off64 = (((unsigned long long)p.c3 << 32) | p.c2) + (unsigned long long)step;
    s.mode = p.rng_mode;
    s.seg = p.ubuf ? p.ubuf + off : nullptr;
    s.k0 = p.k0; s.k1 = p.k1; s.c1 = ordinal; s.c2 = (uint32_t)off64; s.c3 = (uint32_t)(off64 >> 32);
    return s;
  };

  // ---- partial propup of the owned rows: out[b][j] = sum_r src[r][b] * W[r][j].  Item = (column quad, 4 rows
  //      of the minibatch, row group g): the G row groups keep all threads busy on narrow layers; their
  //      partials are added in group order ----
  auto up_partial = [&](const float* __restrict__ src, float* __restrict__ out) {
    const int nb4 = BTS >> 2, nq = CQ * nb4, G = p.G;
    for (int it = tid; it < nq * G; it += NT) {
      const int g = it / nq, rem = it - g * nq, q = rem % CQ, b4 = rem / CQ;
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll 2
      for (int r = g; r < rows; r += G) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + r * lds + 4 * q);
        const float4 v = *reinterpret_cast<const float4*>(src + r * BTS + 4 * b4);
        a0.x = fmaf(v.x, w.x, a0.x); a0.y = fmaf(v.x, w.y, a0.y); a0.z = fmaf(v.x, w.z, a0.z); a0.w = fmaf(v.x, w.w, a0.w);
        a1.x = fmaf(v.y, w.x, a1.x); a1.y = fmaf(v.y, w.y, a1.y); a1.z = fmaf(v.y, w.z, a1.z); a1.w = fmaf(v.y, w.w, a1.w);
        a2.x = fmaf(v.z, w.x, a2.x); a2.y = fmaf(v.z, w.y, a2.y); a2.z = fmaf(v.z, w.z, a2.z); a2.w = fmaf(v.z, w.w, a2.w);
        a3.x = fmaf(v.w, w.x, a3.x); a3.y = fmaf(v.w, w.y, a3.y); a3.z = fmaf(v.w, w.z, a3.z); a3.w = fmaf(v.w, w.w, a3.w);
      }
      float* o = (G > 1 ? park + (size_t)g * nhid : out) + (4 * b4) * ldw + 4 * q;
      *reinterpret_cast<float4*>(o) = a0;
      *reinterpret_cast<float4*>(o + ldw) = a1;
      *reinterpret_cast<float4*>(o + 2 * ldw) = a2;
      *reinterpret_cast<float4*>(o + 3 * ldw) = a3;
    }
    if (G > 1) {
      __syncthreads();
      for (int e = tid; e < BTS * CQ; e += NT) {
        float4 a = *reinterpret_cast<const float4*>(park + 4 * e);
        for (int g = 1; g < G; ++g) {
          const float4 o = *reinterpret_cast<const float4*>(park + (size_t)g * nhid + 4 * e);
          a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
        }
        *reinterpret_cast<float4*>(out + 4 * e) = a;
      }
      __syncthreads();
    }
  };
  // ---- cluster-wide sum of a partial panel (rank order) + hidden bias; fn(b, j0, pre[4]) per quad ----
  auto all_reduce = [&](const float* my_panel, auto&& fn) {
    for (int it = tid; it < B * CQ; it += NT) {
      const int b = it / CQ, q = it - b * CQ;
      const float* lp = my_panel + b * ldw + 4 * q;
      float4 s = ld_remote4(lp, 0);
#pragma unroll
      for (uint32_t c = 1; c < CL; ++c) {
        const float4 o = ld_remote4(lp, c);
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
      }
      const float4 hb4 = *reinterpret_cast<const float4*>(hbs + 4 * q);
      const float pre[4] = {s.x + hb4.x, s.y + hb4.y, s.z + hb4.z, s.w + hb4.w};
      fn(b, 4 * q, pre);
    }
  };
  auto sample4 = [&](const RngSeg& rs, int b, int j0, const float (&mean)[4], float (&out)[4]) {
    float u[4];
    if (rs.mode != MDBN_RNG_BUFFER && (H & 3) == 0) {
      const long long e0 = (long long)b * H + j0;
      const Philox4 x = philox4x32_10((uint32_t)(e0 >> 2), rs.c1, rs.c2, rs.c3, rs.k0, rs.k1);
      u[0] = u24(x.x); u[1] = u24(x.y); u[2] = u24(x.z); u[3] = u24(x.w);
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) u[t] = j0 + t < H ? rng_uniform(rs, (long long)b * H + j0 + t) : 2.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) out[t] = (j0 + t < H && u[t] < mean[t]) ? 1.f : 0.f;
  };

  int* sidx = reinterpret_cast<int*>(misc) + 48;
  // minibatch rows whose local Gibbs steps this CTA runs (use_m == 2)
  const int RPR = (B + CL - 1) / CL, b_lo = (int)rank * RPR, nb = max(0, min(RPR, B - b_lo)), nbq = nb * CQ;

  for (int step = 0; step < p.n_steps; ++step) {
    const int* idxp = p.idx ? p.idx + (size_t)step * B : nullptr;
    // ---- gather the minibatch rows of the owned visible units ----
    if (tid < BTS) sidx[tid] = tid < B ? (idxp ? idxp[tid] : tid) : -1;
    __syncthreads();
    for (int e = tid; e < p.rows_alloc * BTS; e += NT) {
      const int b = e / p.rows_alloc, r = e - b * p.rows_alloc;
      float x = 0.f;
      if (r < rows && sidx[b] >= 0) x = __ldg(&p.data[(long long)sidx[b] * p.ld_data + row0 + r]);
      v0s[r * BTS + b] = x;
      nvs[r * BTS + b] = p.pcd ? roundf(x) : 0.f;
    }
    __syncthreads();
    mark();   // gathered

    // =============================== positive phase ===============================
    {
      float* my = part + (size_t)parity * 2 * nhid;
      up_partial(v0s, my);
      if (p.pcd) up_partial(nvs, my + nhid);
      if (p.use_m) {
        // Gaussian visibles are mean-field (src/rbm.py:669): v = h W^T + vb feeds the next propup unsampled, so
        //   pre_h' = (h W^T + vb) W + hb = h (W^T W) + (vb W + hb).
        // With k > 1 the first k-1 Gibbs steps therefore run on the H x H matrix M = W^T W, locally in every
        // CTA and without any exchange; only the last step (whose visible means enter the statistics) goes
        // through the V visibles.  M does not depend on the minibatch: its partial over the owned rows is
        // exchanged behind the same cluster barrier as the positive phase.
        // 4 x 4 blocks of M, the owned rows dealt over four adjacent lanes (two quad loads per 16 FMAs; the scalar x quad
        // form was bound by the shared-memory pipe), joined by two butterfly steps
        for (int it = tid; it < CQ * CQ * 4; it += NT) {
          const int blk = it >> 2, rg = it & 3, iq = blk / CQ, q = blk - iq * CQ;
          float acc[4][4];
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
#pragma unroll 4
          for (int r0 = 0; r0 < rows; r0 += 4) {                 // (same trip count in the four lanes)
            const int r = min(r0 + rg, rows - 1);
            const float4 a4 = r0 + rg < rows ? *reinterpret_cast<const float4*>(Ws + r * lds + 4 * iq) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b4 = *reinterpret_cast<const float4*>(Ws + r * lds + 4 * q);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
          }
          const unsigned am = 0xFu << (lane & ~3);               // the four lanes of this block
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              acc[x][y] += __shfl_xor_sync(am, acc[x][y], 1);
              acc[x][y] += __shfl_xor_sync(am, acc[x][y], 2);
            }
          if (rg == 0) {
#pragma unroll
            for (int x = 0; x < 4; ++x)
              if (4 * iq + x < H)
                *reinterpret_cast<float4*>(Mp + (4 * iq + x) * ldw + 4 * q) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
          }
        }
        // row H: vb . W (on the last threads of the CTA, which have no or few blocks)
        for (int q = NT - 1 - tid; q < CQ; q += NT) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
          for (int r = 0; r < rows; ++r) {
            const float wi = vbs[r];
            const float4 w4 = *reinterpret_cast<const float4*>(Ws + r * lds + 4 * q);
            a.x = fmaf(wi, w4.x, a.x); a.y = fmaf(wi, w4.y, a.y); a.z = fmaf(wi, w4.z, a.z); a.w = fmaf(wi, w4.w, a.w);
          }
          *reinterpret_cast<float4*>(Mp + H * ldw + 4 * q) = a;
        }
      }
      mark();
      cluster_sync();
      mark();
      if (p.use_m) {
        // (use_m == 2) the uniforms of this CTA's rows for the k - 1 local Gibbs steps are drawn HERE, between the remote
        // loads and their sum: a Philox block per thread rides on the DSMEM latency
        const int n_u = p.use_m == 2 ? (p.k - 1) * nbq : 0;
        auto draw = [&](int ui) {
          const int s = ui / nbq, e = ui - s * nbq, ub = b_lo + e / CQ, uq = e % CQ;
          const long long ubase = (long long)B * H + (long long)s * p.u_step_stride;
          const RngSeg rs_h = seg(ubase + p.u_off_h, 2u + 2u * s, step);
          float u[4];
          if (rs_h.mode != MDBN_RNG_BUFFER && (H & 3) == 0) {
            const long long e0 = (long long)ub * H + 4 * uq;
            const Philox4 x = philox4x32_10((uint32_t)(e0 >> 2), rs_h.c1, rs_h.c2, rs_h.c3, rs_h.k0, rs_h.k1);
            u[0] = u24(x.x); u[1] = u24(x.y); u[2] = u24(x.z); u[3] = u24(x.w);
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) u[t] = 4 * uq + t < H ? rng_uniform(rs_h, (long long)ub * H + 4 * uq + t) : 2.f;
          }
          *reinterpret_cast<float4*>(Us + (size_t)ui * 4) = make_float4(u[0], u[1], u[2], u[3]);
        };
        int ui = tid;
        for (int it = tid; it < (H + 1) * CQ; it += NT) {
          const float* lp = Mp + 4 * it + (it / CQ) * (ldw - 4 * CQ);
          float4 v[CL];
#pragma unroll
          for (uint32_t c = 0; c < CL; ++c) v[c] = ld_remote4(lp, c);
          if (ui < n_u) { draw(ui); ui += NT; }
          float4 sm = v[0];
#pragma unroll
          for (uint32_t c = 1; c < CL; ++c) { sm.x += v[c].x; sm.y += v[c].y; sm.z += v[c].z; sm.w += v[c].w; }
          *reinterpret_cast<float4*>(Ms + (lp - Mp)) = sm;
        }
        for (; ui < n_u; ui += NT) draw(ui);
      }
      const RngSeg rs0 = seg(0, 0, step);
      all_reduce(my, [&](int b, int j0, const float (&pre)[4]) {
        const float mean[4] = {sigmoid_fast_(pre[0]), sigmoid_fast_(pre[1]), sigmoid_fast_(pre[2]), sigmoid_fast_(pre[3])};
        *reinterpret_cast<float4*>(phs + b * ldw + j0) = make_float4(mean[0], mean[1], mean[2], mean[3]);
        float smp[4];
        if (p.pcd) {                                             // chain starts from the persistent state (src/rbm.py:308-311)
          const float4 c4 = *reinterpret_cast<const float4*>(pc + b * ldw + j0);
          smp[0] = c4.x; smp[1] = c4.y; smp[2] = c4.z; smp[3] = c4.w;
        } else {
          sample4(rs0, b, j0, mean, smp);
        }
        *reinterpret_cast<float4*>(hs + b * ldw + j0) = make_float4(smp[0], smp[1], smp[2], smp[3]);
      });
      if (p.pcd && rank == 0)                                    // pre-activations of round(v0) for the monitor
        all_reduce(my + nhid, [&](int b, int j0, const float (&pre)[4]) {
          *reinterpret_cast<float4*>(nhs + b * ldw + j0) = make_float4(pre[0], pre[1], pre[2], pre[3]);
        });
      parity ^= 1;
      __syncthreads();
      mark();   // positive phase done
    }
    // pseudo-likelihood monitor (src/rbm.py:421-447) on CTA 0: pre-update W[bit,:] and vb[bit] come from their owner
    if (p.pcd && rank == 0) {
      const int bit = (bit0 + step) % V;
      const uint32_t owner = (uint32_t)(bit / p.rows_per_cta);
      const int lr_ = bit - (int)owner * p.rows_per_cta;
      const float vbv = ld_remote1(vbs + lr_, owner);
      for (int b = warp; b < B; b += NT / 32) {
        const float x = roundf(__ldg(&p.data[(long long)sidx[b] * p.ld_data + bit]));
        const float d = 1.f - 2.f * x;
        float h0 = 0.f, h1 = 0.f;
        for (int j = lane; j < H; j += 32) {
          const float pre = nhs[b * ldw + j];
          h0 += softplusf_(pre);
          h1 += softplusf_(pre + d * ld_remote1(Ws + lr_ * lds + j, owner));
        }
        h0 = warp_sum(h0);
        h1 = warp_sum(h1);
        if (lane == 0) {
          float vterm;
          if (p.kind == MDBN_GRBM) { const float a = x - vbv, c = (1.f - x) - vbv; vterm = 0.5f * (a * a - c * c); }
          else vterm = d * vbv;
          misc[16 + b] = -(float)V * softplusf_((h1 - h0) + vterm);
        }
      }
      __syncthreads();
      if (tid == 0) {
        float c = 0.f;
        for (int b = 0; b < B; ++b) c += misc[16 + b];
        if (p.cost_out) p.cost_out[step] = c * p.cost_scale;
      }
    }

    // =============================== k Gibbs steps ===============================
    float cost_acc = 0.f;
    int s_begin = 0;
    if (p.use_m == 2) {
      // Register edition of the local steps below (H <= 2 MH): thread = (minibatch row, column quad, half of the rows of
      // M).  Its 4 x MH block of M = W^T W stays in registers for all k - 1 steps, so a step reads only the chain state
      // from shared memory (5 LDS.128 instead of 40 LDS.32 + 40 LDS.128 per thread: the loop was bound by the
      // shared-memory pipe) and the two halves of a dot product are joined with one shuffle.
      // The chains of different minibatch rows are independent: every CTA of the cluster runs the k - 1 local steps for
      // ITS rows only (ceil(B / 8) of them: an eighth of the draws and of the dot products) and the final states are
      // exchanged once through distributed shared memory.
      // one warp per row (2 CQ <= 32 lanes: column quad x half): a row of the chain state is private to its warp, so the
      // steps need no block barrier
      const int half = lane & 1;
      const bool mine = warp < nb && (lane >> 1) < CQ;
      const int lb = mine ? warp : 0, q = mine ? (lane >> 1) : 0, b = mine ? b_lo + lb : 0, pr = lb * CQ + q;
      const int HH = ((H + 7) >> 3) << 2;                        // rows per half (multiple of 4, <= MH)
      float4 mreg[MH];
#pragma unroll
      for (int i = 0; i < MH; ++i) {
        const int row = HH * half + i;
        mreg[i] = (i < HH && row < H) ? *reinterpret_cast<const float4*>(Ms + row * ldw + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f);               // (vb W + hb) for the quad, on the first half only
      if (half == 0) {
        const float4 c4 = *reinterpret_cast<const float4*>(Ms + H * ldw + 4 * q);
        const float4 hb4 = *reinterpret_cast<const float4*>(hbs + 4 * q);
        a0 = make_float4(c4.x + hb4.x, c4.y + hb4.y, c4.z + hb4.z, c4.w + hb4.w);
      }
      // chunk offsets of this half's part of the chain state (a chunk beyond the half, or beyond the row, re-reads
      // chunk 0 against zero rows of M: no branches in the loop)
      int hoff[MH / 4];
#pragma unroll
      for (int i4 = 0; i4 < MH / 4; ++i4) hoff[i4] = (4 * i4 < HH && HH * half + 4 * i4 < ldw) ? 4 * i4 : 0;
      const float* hrow = hs + b * ldw + HH * half;
      const float4* up = reinterpret_cast<const float4*>(Us) + (mine ? pr : 0);
      mark();   // registers of the local steps ready (the uniforms were drawn under the all-reduce of M)
      for (int s = 0; s < (warp < nb ? p.k - 1 : 0); ++s) {
        float4 a = a0;
#pragma unroll
        for (int i4 = 0; i4 < MH / 4; ++i4) {
          const float4 h4 = *reinterpret_cast<const float4*>(hrow + hoff[i4]);
          const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float4 m4 = mreg[4 * i4 + t];
            a.x = fmaf(hv[t], m4.x, a.x); a.y = fmaf(hv[t], m4.y, a.y); a.z = fmaf(hv[t], m4.z, a.z); a.w = fmaf(hv[t], m4.w, a.w);
          }
        }
        const float4 u4 = mine ? up[(size_t)s * nbq] : make_float4(2.f, 2.f, 2.f, 2.f);
        a.x += __shfl_xor_sync(0xffffffffu, a.x, 1); a.y += __shfl_xor_sync(0xffffffffu, a.y, 1);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, 1); a.w += __shfl_xor_sync(0xffffffffu, a.w, 1);
        float4 smp;
        smp.x = (4 * q + 0 < H && u4.x < sigmoid_fast_(a.x)) ? 1.f : 0.f;
        smp.y = (4 * q + 1 < H && u4.y < sigmoid_fast_(a.y)) ? 1.f : 0.f;
        smp.z = (4 * q + 2 < H && u4.z < sigmoid_fast_(a.z)) ? 1.f : 0.f;
        smp.w = (4 * q + 3 < H && u4.w < sigmoid_fast_(a.w)) ? 1.f : 0.f;
        __syncwarp();
        if (mine && half == 0) *reinterpret_cast<float4*>(hs + b * ldw + 4 * q) = smp;
        __syncwarp();
        if (s == 0) mark();   // first local step
      }
      // every CTA fetches the rows the other CTAs ran (the next cluster barrier — in the last Gibbs step — comes after
      // these reads, before anybody overwrites its chain state again)
      cluster_sync();
      for (int it = tid; it < B * CQ; it += NT) {
        const int rb = it / CQ, rq = it - rb * CQ;
        const uint32_t owner = (uint32_t)(rb / RPR);
        if (owner != rank) *reinterpret_cast<float4*>(hs + rb * ldw + 4 * rq) = ld_remote4(hs + rb * ldw + 4 * rq, owner);
      }
      __syncthreads();
      s_begin = p.k - 1;
      mark();   // local Gibbs steps done
    }
    for (int s = s_begin; s < p.k; ++s) {
      const bool last = (s == p.k - 1);
      const long long ubase = (long long)B * H + (long long)s * p.u_step_stride;
      const RngSeg rs_v = seg(ubase + p.u_off_v, 1u + 2u * s, step);
      const RngSeg rs_h = seg(ubase + p.u_off_h, 2u + 2u * s, step);
      if (p.use_m && !last) {
        // h' ~ sigmoid(h M + vb W + hb): every CTA for the whole [B, H], no exchange (B * CQ <= NT)
        float smp[4] = {0.f, 0.f, 0.f, 0.f};
        const int b = tid / CQ, q = tid - b * CQ;
        const bool mine = tid < B * CQ;
        if (mine) {
          const float4 c4 = *reinterpret_cast<const float4*>(Ms + H * ldw + 4 * q);
          const float4 hb4 = *reinterpret_cast<const float4*>(hbs + 4 * q);
          float4 a = make_float4(c4.x + hb4.x, c4.y + hb4.y, c4.z + hb4.z, c4.w + hb4.w);
          const float* hrow = hs + b * ldw;
#pragma unroll 8
          for (int i = 0; i < H; ++i) {
            const float hv = hrow[i];
            const float4 m4 = *reinterpret_cast<const float4*>(Ms + i * ldw + 4 * q);
            a.x = fmaf(hv, m4.x, a.x); a.y = fmaf(hv, m4.y, a.y); a.z = fmaf(hv, m4.z, a.z); a.w = fmaf(hv, m4.w, a.w);
          }
          const float mean[4] = {sigmoid_fast_(a.x), sigmoid_fast_(a.y), sigmoid_fast_(a.z), sigmoid_fast_(a.w)};
          sample4(rs_h, b, 4 * q, mean, smp);
        }
        __syncthreads();
        if (mine) *reinterpret_cast<float4*>(hs + b * ldw + 4 * q) = make_float4(smp[0], smp[1], smp[2], smp[3]);
        __syncthreads();
        continue;
      }
      // propdown of the owned rows (complete dot products: every CTA holds the whole chain state) + epilogue
      const int nb4 = BTS >> 2;
      for (int it = tid; it < rows * nb4; it += NT) {
        const int r = it % rows, b4 = it / rows;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* wr = Ws + r * lds;
        const float* h0p = hs + (4 * b4) * ldw;
#pragma unroll 2
        for (int q = 0; q < CQ; ++q) {
          const float4 w = *reinterpret_cast<const float4*>(wr + 4 * q);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float4 h4 = *reinterpret_cast<const float4*>(h0p + t * ldw + 4 * q);
            acc[t] = fmaf(h4.x, w.x, fmaf(h4.y, w.y, fmaf(h4.z, w.z, fmaf(h4.w, w.w, acc[t]))));
          }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int b = 4 * b4 + t;
          float v_in = 0.f, mean = 0.f;
          if (b < B) {
            const float pre = acc[t] + vbs[r];
            if (p.kind == MDBN_GRBM) {
              mean = pre;
              v_in = pre;                                        // mean-field visible (src/rbm.py:669)
            } else {
              mean = sigmoidf_(pre);
              v_in = rng_uniform(rs_v, (long long)b * V + row0 + r) < mean ? 1.f : 0.f;
            }
            if (last && !p.pcd) {
              const float t0 = v0s[r * BTS + b];
              if (p.kind == MDBN_GRBM) { const float d = sigmoidf_(pre) - t0; cost_acc += d * d; }      // :697
              else cost_acc += t0 * softplusf_(-pre) + (1.f - t0) * softplusf_(pre);                    // :479-480
            }
          }
          vin[r * BTS + b] = v_in;
          if (last) nvs[r * BTS + b] = mean;
        }
      }
      __syncthreads();
      if (s == 0) mark();   // propdown + epilogue
      float* my = part + (size_t)parity * 2 * nhid;
      up_partial(vin, my);
      if (s == 0) mark();
      if (last && !p.pcd) {
        const float c = block_sum(cost_acc, misc);
        if (tid == 0) misc[40] = c;
      }
      cluster_sync();
      if (s == 0) mark();
      all_reduce(my, [&](int b, int j0, const float (&pre)[4]) {
        const float mean[4] = {sigmoid_fast_(pre[0]), sigmoid_fast_(pre[1]), sigmoid_fast_(pre[2]), sigmoid_fast_(pre[3])};
        float smp[4];
        sample4(rs_h, b, j0, mean, smp);
        const float4 s4 = make_float4(smp[0], smp[1], smp[2], smp[3]);
        *reinterpret_cast<float4*>(hs + b * ldw + j0) = s4;
        if (last) {
          *reinterpret_cast<float4*>(nhs + b * ldw + j0) = make_float4(mean[0], mean[1], mean[2], mean[3]);
          if (p.pcd) *reinterpret_cast<float4*>(pc + b * ldw + j0) = s4;     // new persistent chain (src/rbm.py:372)
        }
      });
      parity ^= 1;
      __syncthreads();
      if (s == 0) mark();   // Gibbs step 0 done
    }
    mark();   // all Gibbs steps done
    if (!p.pcd && rank == 0 && warp == 0) {                      // reconstruction cost: rank-ordered sum of the partials
      float c = lane < CL ? ld_remote1(misc + 40, (uint32_t)lane) : 0.f;
      c = warp_sum(c);
      if (lane == 0 && p.cost_out) p.cost_out[step] = c * p.cost_scale;
    }

    // =============================== statistics + update (owned rows, in shared memory) ===============================
    // thread = (pair of owned rows, column quad): the hidden means of a minibatch row are read once for the two rows and
    // one round covers the slab (one row per thread left 512 threads 700 items: two rounds, the second a third full)
    {
      const int half = (rows + 1) >> 1;
      auto finish = [&](int r, int q, const float (&gs)[4]) {
        float4 w4 = *reinterpret_cast<float4*>(Ws + r * lds + 4 * q), s4 = *reinterpret_cast<float4*>(Ss + r * lds + 4 * q);
        float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.wc != 0.f) n4 = __ldg(reinterpret_cast<const float4*>(p.Wsnap + (size_t)(row0 + r) * ldw + 4 * q));
        float wv[4] = {w4.x, w4.y, w4.z, w4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w}, nv4[4] = {n4.x, n4.y, n4.z, n4.w};
        float wo[4], so[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float gw = gs[c] * p.inv_bnom - p.wc * nv4[c];                  // src/rbm.py:411-415
          float mult = p.decay;
          if (p.c1 != 0.f) {
            const float t = fabsf(wv[c]) + 0.001f;                        // :347-350, one reciprocal
            const float invD = t * rcp_approx(t + p.c1);
            gw *= invD;
            mult *= invD;                                                 // :353-356
          }
          so[c] = gw + (sv[c] - gw) * p.mom;                              // :361
          wo[c] = 4 * q + c < H ? wv[c] * mult + sv[c] * p.lr : wv[c];    // :364 (OLD speed); padding columns stay zero
          if (4 * q + c >= H) so[c] = sv[c];
        }
        *reinterpret_cast<float4*>(Ws + r * lds + 4 * q) = make_float4(wo[0], wo[1], wo[2], wo[3]);
        *reinterpret_cast<float4*>(Ss + r * lds + 4 * q) = make_float4(so[0], so[1], so[2], so[3]);
      };
      for (int it = tid; it < half * CQ; it += NT) {
        const int r0 = it / CQ, q = it - r0 * CQ, r1 = r0 + half;
        const bool two = r1 < rows;
        const float* a0p = v0s + r0 * BTS; const float* n0p = nvs + r0 * BTS;
        const float* a1p = v0s + (two ? r1 : r0) * BTS; const float* n1p = nvs + (two ? r1 : r0) * BTS;
        float g0[4] = {0.f, 0.f, 0.f, 0.f}, g1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int b = 0; b < B; ++b) {
          const float4 ph4 = *reinterpret_cast<const float4*>(phs + b * ldw + 4 * q);
          const float4 nh4 = *reinterpret_cast<const float4*>(nhs + b * ldw + 4 * q);
          const float a0 = a0p[b], n0 = -n0p[b], a1 = a1p[b], n1 = -n1p[b];
          g0[0] = fmaf(a0, ph4.x, g0[0]); g0[0] = fmaf(n0, nh4.x, g0[0]);
          g0[1] = fmaf(a0, ph4.y, g0[1]); g0[1] = fmaf(n0, nh4.y, g0[1]);
          g0[2] = fmaf(a0, ph4.z, g0[2]); g0[2] = fmaf(n0, nh4.z, g0[2]);
          g0[3] = fmaf(a0, ph4.w, g0[3]); g0[3] = fmaf(n0, nh4.w, g0[3]);
          g1[0] = fmaf(a1, ph4.x, g1[0]); g1[0] = fmaf(n1, nh4.x, g1[0]);
          g1[1] = fmaf(a1, ph4.y, g1[1]); g1[1] = fmaf(n1, nh4.y, g1[1]);
          g1[2] = fmaf(a1, ph4.z, g1[2]); g1[2] = fmaf(n1, nh4.z, g1[2]);
          g1[3] = fmaf(a1, ph4.w, g1[3]); g1[3] = fmaf(n1, nh4.w, g1[3]);
        }
        finish(r0, q, g0);
        if (two) finish(r1, q, g1);
      }
    }
    // biases on the LAST warps of the CTA (they have no or few row pairs): visible bias of the owned rows
    // (src/rbm.py:417), hidden bias (:416 — every CTA, identical result); nothing above writes what they read
    for (int r = NT - 1 - tid; r < rows; r += NT) {
      float gsum = 0.f;
      for (int b = 0; b < B; ++b) gsum += v0s[r * BTS + b] - nvs[r * BTS + b];
      const float gb = gsum * p.inv_b, sv = svbs[r];
      svbs[r] = gb + (sv - gb) * p.mom;
      vbs[r] = vbs[r] + sv * p.lr;
    }
    for (int j = (NT - 1 - tid + NT - rows % NT) % NT; j < H; j += NT) {      // (threads in front of the visible-bias ones)
      float gsum = 0.f;
      for (int b = 0; b < B; ++b) gsum += phs[b * ldw + j] - nhs[b * ldw + j];
      const float gb = gsum * p.inv_b, sv = shbs[j];
      shbs[j] = gb + (sv - gb) * p.mom;
      hbs[j] = hbs[j] + sv * p.lr;
    }
    __syncthreads();
    mark();   // update done
  }   // step

  // ---- write the resident state back ----
  for (int e = tid; e < rows * CQ; e += NT) {
    const int r = e / CQ, q = e - r * CQ;
    *reinterpret_cast<float4*>(p.W + (size_t)(row0 + r) * ldw + 4 * q) = *reinterpret_cast<const float4*>(Ws + r * lds + 4 * q);
    *reinterpret_cast<float4*>(p.S + (size_t)(row0 + r) * ldw + 4 * q) = *reinterpret_cast<const float4*>(Ss + r * lds + 4 * q);
  }
  for (int r = tid; r < rows; r += NT) {
    p.vb[row0 + r] = vbs[r];
    p.Svb[row0 + r] = svbs[r];
  }
  if (rank == 0) {
    for (int j = tid; j < H; j += NT) {
      p.hb[j] = hbs[j];
      p.Shb[j] = shbs[j];
    }
    if (p.pcd) {
      for (int e = tid; e < B * H; e += NT) {
        const int b = e / H, j = e - b * H;
        p.P[e] = pc[b * ldw + j];
      }
      if (tid == 0) *p.bit_idx = (bit0 + p.n_steps) % V;          // :445, one advance per step
    }
  }
  cluster_sync();      // nobody leaves while a neighbour may still read its shared memory
}

struct Geometry {
  int BTS, CQ, rows_per_cta, rows_alloc, G, lds, use_m;
  int off_U, off_M, off_Mp, off_park, off_W, off_S, off_v0, off_nv, off_vin, off_hs, off_pc, off_ph, off_nh, off_part, off_vb, off_svb, off_hb, off_shb,
      off_misc;
  size_t smem;
  bool ok;
};

static Geometry plan(const mdbn_cd_args& a) {
  Geometry g{};
  g.ok = false;
  if (a.B < 1 || a.B > 20 || a.ldw % 4 != 0 || a.ldw > 256 || a.V < 1) return g;
  if (((uintptr_t)a.W | (uintptr_t)a.W_speed | (uintptr_t)a.W_snap) & 15) return g;
  g.BTS = (a.B + 3) / 4 * 4;
  g.CQ = a.ldw / 4;
  g.rows_per_cta = (a.V + CL - 1) / CL;
  g.rows_alloc = (g.rows_per_cta + 3) & ~3;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 127) & ~(size_t)127; return (int)o; };
  g.lds = ((a.ldw / 4) & 1) ? a.ldw : a.ldw + 4;
  const size_t slab = (size_t)g.rows_alloc * g.lds * 4, vis = (size_t)g.rows_alloc * g.BTS * 4,
               hid = (size_t)g.BTS * a.ldw * 4;
  g.G = NT / (g.CQ * (g.BTS / 4));
  g.G = g.G < 1 ? 1 : (g.G > 8 ? 8 : g.G);
  while (g.G > 1 && (size_t)g.G * hid > 48 * 1024) --g.G;
  g.use_m = a.kind == MDBN_GRBM && a.k > 1 && a.ldw <= 64 && a.B * g.CQ <= NT;
  static const bool no_mreg = getenv("MDBN_TINY_NO_MREG") != nullptr;
  if (g.use_m && a.ldw <= 2 * tn::MH && 2 * g.CQ <= 32 && (a.B + CL - 1) / CL <= NT / 32 && !no_mreg) g.use_m = 2;      // M in registers
  g.off_M = take(g.use_m ? (size_t)(a.H + 1) * a.ldw * 4 : 0);
  g.off_Mp = take(g.use_m ? (size_t)(a.H + 1) * a.ldw * 4 : 0);
  g.off_park = take(g.G > 1 ? (size_t)g.G * hid : 0);
  g.off_W = take(slab); g.off_S = take(slab);
  g.off_v0 = take(vis); g.off_nv = take(vis); g.off_vin = take(vis);
  g.off_hs = take(hid); g.off_pc = take(hid); g.off_ph = take(hid); g.off_nh = take(hid);
  g.off_part = take(4 * hid);
  g.off_vb = take((size_t)g.rows_alloc * 4); g.off_svb = take((size_t)g.rows_alloc * 4);
  g.off_hb = take((size_t)a.ldw * 4); g.off_shb = take((size_t)a.ldw * 4);
  g.off_misc = take(512);
  g.off_U = 0;
  if (g.use_m == 2) {      // uniforms of the k - 1 local Gibbs steps, [k-1][B][CQ] quads
    const size_t ub = (size_t)(a.k - 1) * a.B * g.CQ * 16;
    if (off + ub <= 200 * 1024) g.off_U = take(ub);
    else g.use_m = 1;
  }
  g.smem = off;
  g.ok = g.smem <= 200 * 1024;
  return g;
}

}  // namespace tn

bool tiny_supported(const mdbn_ctx*, const mdbn_cd_args& a) {
  if (a.phase != MDBN_PHASE_FULL) return false;
  return tn::plan(a).ok;
}

// n_steps consecutive steps in one launch (a.indices [n_steps][B], a.cost_out [n_steps]; PHILOX when n_steps > 1)
int tiny_cd_steps(mdbn_ctx* c, const mdbn_cd_args& a, int n_steps, cudaStream_t st) {
  tn::Geometry g = tn::plan(a);
  MDBN_CHECK(g.ok, "tiny path: unsupported shape");
  MDBN_CHECK(n_steps >= 1, "tiny path: n_steps must be >= 1");
  MDBN_CHECK(n_steps == 1 || a.rng.mode == MDBN_RNG_PHILOX, "tiny path: chained steps need the PHILOX generator");
  tn::Params p{};
  p.W = a.W; p.S = a.W_speed; p.Wsnap = a.weightcost != 0.f ? a.W_snap : nullptr; p.ldw = a.ldw;
  p.hb = a.hbias; p.vb = a.vbias; p.Shb = a.hbias_speed; p.Svb = a.vbias_speed;
  p.data = a.data; p.ld_data = a.ld_data; p.idx = a.indices;
  p.P = a.persistent; p.bit_idx = a.bit_i_idx; p.cost_out = a.cost_out;
  p.kind = a.kind; p.B = a.B; p.V = a.V; p.H = a.H; p.k = a.k; p.pcd = a.persistent != nullptr; p.n_steps = n_steps;
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
  p.rows_per_cta = g.rows_per_cta; p.rows_alloc = g.rows_alloc; p.BTS = g.BTS; p.CQ = g.CQ; p.G = g.G; p.lds = g.lds;
  p.use_m = g.use_m; p.off_M = g.off_M; p.off_Mp = g.off_Mp; p.off_U = g.off_U;
  p.off_park = g.off_park;
  p.off_W = g.off_W; p.off_S = g.off_S; p.off_v0 = g.off_v0; p.off_nv = g.off_nv; p.off_vin = g.off_vin;
  p.off_hs = g.off_hs; p.off_pc = g.off_pc; p.off_ph = g.off_ph; p.off_nh = g.off_nh; p.off_part = g.off_part;
  p.off_vb = g.off_vb; p.off_svb = g.off_svb; p.off_hb = g.off_hb; p.off_shb = g.off_shb; p.off_misc = g.off_misc;

  static bool configured[64] = {};
  if (!configured[c->device]) {
    MDBN_CUDA(cudaFuncSetAttribute(tn::cd_tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[c->device] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tn::CL);
  cfg.blockDim = dim3(tn::NT);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tn::CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const bool want_timing = getenv("MDBN_TINY_TIMING") != nullptr;
  p.dbg = want_timing ? reinterpret_cast<unsigned long long*>(c->barrier) + 32 : nullptr;
  MDBN_CUDA(cudaLaunchKernelEx(&cfg, tn::cd_tiny_kernel, p));
  c->launches++;
  if (p.dbg) {
    unsigned long long t[16];
    MDBN_CUDA(cudaStreamSynchronize(st));
    MDBN_CUDA(cudaMemcpy(t, p.dbg, sizeof(t), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[tiny timeline us] V=%d H=%d B=%d k=%d G=%d:", a.V, a.H, a.B, a.k, g.G);
    for (int i = 1; i < 13; ++i) fprintf(stderr, " %.1f", (double)(t[i] - t[0]) * 1e-3);
    fprintf(stderr, "\n");
  }
  return 0;
}

}  // namespace mdbn
