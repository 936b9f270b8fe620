// Medium layers at skinny batch (784->500, 1000->1000, 1686->200 ... with B <= 20): CD-k / PCD-k steps as ONE persistent
// cooperative kernel whose all-reduces are BROADCASTS instead of reductions.
//
// On these layers W (1-4 MB) lives in L2 and a step is a few microseconds of streaming; the row-slab kernel (skinny.cu)
// spends its time in the three all-reduces of the [B, H] pre-activations — B*H 64-bit atomics PER CTA (the SM issues
// ~0.75 of them per clock: 6.7 us at 784->500, B = 20) and a [B, H] chain rebuild by EVERY CTA (7.7 us).  Here the work is
// cut the other way round:
//   propup   : a CTA owns a slice of hidden COLUMNS and sums over ALL visible units (it holds the whole [B, V] visible state
//              in shared memory and reads its columns of W from L2): complete pre-activations, no cross-CTA sum; bias,
//              sigmoid and the Bernoulli draw are computed once, by the owner, and published as [B, H] in global memory;
//   propdown : a CTA owns a slice of visible ROWS and sums over all hidden units (whole [B, H] sample in shared memory,
//              its rows of W contiguous): complete again; the result is published as [V, B];
//   between two propagations: ONE grid barrier, then every CTA copies the 40-80 KB panel it needs from L2;
//   update   : the row owner updates its rows of W / W_speed (it kept its rows of v0 and of the negative means), reading
//              the positive / negative hidden means of all columns from the published panels; biases by their owners.
// No atomics, no redundant rebuild; 2k + 2 barriers per step.  Same arithmetic contract as the other paths (fp32,
// src/rbm.py semantics, Philox draws indexed by element or the caller's uniform buffer).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include "ctx.h"

namespace mdbn {
namespace md {

constexpr int NT = 256, NWARP = NT / 32;
constexpr int MAXB = 20;

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
  UpdateScalars u;
  float inv_b, cost_scale;
  int rng_mode;
  const float* ubuf;
  uint32_t k0, k1, c2, c3;
  long long u_step_stride, u_off_v, u_off_h;
  int BTS, CQ, ldh, NQ, NR, HC;     // batch tile, column quads, padded H, own quads / rows per CTA, statistics column chunk
  float *Hg, *PHg, *NHg;            // [BTS][ldh] published hidden sample / positive means / negative means
  float* Vg;                        // [V][BTS]  published visible state (transposed)
  float* cost_part;                 // [grid]
  float* pl_part;                   // [grid][BTS][2]
  unsigned long long* bar;
  int off_vbuf, off_hs, off_v0o, off_nvo, off_red, off_preo, off_pho, off_nho, off_misc, off_wcol;
  unsigned long long* dbg;          // optional phase timeline (MDBN_MID_TIMING=1), CTA 0
};

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast_(float x) { return rcp_approx(1.0f + __expf(-x)); }

// Device-wide barrier (all CTAs co-resident: cooperative launch); the counter is monotonic within a launch and reset by
// the last CTA to leave the kernel.
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

__global__ void __launch_bounds__(NT, 1) cd_mid_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* vbuf = reinterpret_cast<float*>(smem + p.off_vbuf);   // [V][BTS] whole visible state; the statistics pass parks ph / nh chunks here
  float* hs = reinterpret_cast<float*>(smem + p.off_hs);       // [BTS][ldh] whole hidden sample
  float* v0o = reinterpret_cast<float*>(smem + p.off_v0o);     // [NR][BTS] own rows of v0
  float* nvo = reinterpret_cast<float*>(smem + p.off_nvo);     // [NR][BTS] own rows of the last negative visible mean
  float* red = reinterpret_cast<float*>(smem + p.off_red);     // propup partials [chunk][b group][4 rows][4 columns]
  float* preo = reinterpret_cast<float*>(smem + p.off_preo);   // [NQ][BTS][4] pre-activations of the own columns (no bias)
  float* pho = reinterpret_cast<float*>(smem + p.off_pho);     // [NQ][BTS][4] positive means of the own columns
  float* nho = reinterpret_cast<float*>(smem + p.off_nho);     // [NQ][BTS][4] last negative means
  float* misc = reinterpret_cast<float*>(smem + p.off_misc);   // [64] block_sum scratch; [32..] minibatch row numbers
  float4* wcol = reinterpret_cast<float4*>(smem + p.off_wcol); // [V] one column quad of W, staged for a propup

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, G = gridDim.x;
  const int B = p.B, V = p.V, H = p.H, BTS = p.BTS, ldh = p.ldh, ldw = p.ldw;
  const int q0 = cta * p.NQ, nq = max(0, min(p.NQ, p.CQ - q0));      // own column quads
  const int i0 = cta * p.NR, nr = max(0, min(p.NR, V - i0));         // own visible rows
  unsigned long long bar_target = 0;
  int* sidx = reinterpret_cast<int*>(misc) + 32;
  const int bit0 = p.pcd ? *p.bit_idx : 0;
  int dbg_i = 0;
  auto mark = [&]() {
    if (p.dbg && cta == 0 && tid == 0 && dbg_i < 24) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[dbg_i++] = t;
    }
  };
  mark();
  // global -> shared copy of n4 float4, eight L2 loads in flight per thread
  auto copy_f4 = [&](float* dst, const float* src, int n4) {
    constexpr int U = 8;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int e0 = tid; e0 < n4; e0 += U * NT) {
      float4 t[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * NT;
        t[u] = e < n4 ? __ldcg(s4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * NT;
        if (e < n4) d4[e] = t[u];
      }
    }
  };

  auto seg = [&](long long off, uint32_t ordinal, int step) {
    RngSeg s;
    const unsigned long long off64 = (((unsigned long long)p.c3 << 32) | p.c2) + (unsigned long long)step;
    s.mode = p.rng_mode;
    s.seg = p.ubuf ? p.ubuf + off : nullptr;
    s.k0 = p.k0; s.k1 = p.k1; s.c1 = ordinal; s.c2 = (uint32_t)off64; s.c3 = (uint32_t)(off64 >> 32);
    return s;
  };

  // ---- pre-activations of the own columns from the visible state in vbuf: preo[qq][b][c] = sum_i v[i][b] W[i][4(q0+qq)+c].
  //      Thread = (group of 4 minibatch rows, row chunk); the chunks are added in fixed order ----
  auto propup = [&](bool rounded) {
    const int nb4 = BTS >> 2, NCH = NT / nb4;
    const int bg = tid % nb4, ch = tid / nb4;
    for (int qq = 0; qq < nq; ++qq) {
      float acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
      // the column quad of all V rows: every thread fetches its share at once (one L2 round trip), then the sums run
      // out of shared memory
      {
        const float* wp = p.W + 4 * (q0 + qq);
        for (int i = tid; i < V; i += NT) wcol[i] = __ldcg(reinterpret_cast<const float4*>(wp + (size_t)i * ldw));
      }
      __syncthreads();
      if (ch < NCH) {
#pragma unroll 4
        for (int i = ch; i < V; i += NCH) {
          const float4 w = wcol[i];
          float4 v = *reinterpret_cast<const float4*>(vbuf + i * BTS + 4 * bg);
          if (rounded) v = make_float4(roundf(v.x), roundf(v.y), roundf(v.z), roundf(v.w));      // src/rbm.py:428
          const float vv[4] = {v.x, v.y, v.z, v.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(vv[a], ww[c], acc[a][c]);
        }
        float4* rp = reinterpret_cast<float4*>(red + (size_t)(ch * nb4 + bg) * 16);
#pragma unroll
        for (int a = 0; a < 4; ++a) rp[a] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      }
      __syncthreads();
      // 80 outputs x 3 interleaved chunk ranges, joined in fixed order
      const int nout = BTS * 4;
      float* red2 = red + NT * 16;
      if (tid < 3 * nout) {
        const int part = tid / nout, o = tid - part * nout, b = o >> 2, c = o & 3;
        float s = 0.f;
        for (int h = part; h < NCH; h += 3) s += red[(size_t)(h * nb4 + (b >> 2)) * 16 + (b & 3) * 4 + c];
        red2[part * nout + o] = s;
      }
      __syncthreads();
      if (tid < nout) preo[qq * nout + tid] = (red2[tid] + red2[nout + tid]) + red2[2 * nout + tid];
      __syncthreads();
    }
  };
  // whole [B][H] hidden panel from global memory (row stride lds) into hs
  auto load_hidden = [&](const float* src, int lds) {
    if (lds == ldh && (((uintptr_t)src) & 15) == 0) {      // published panel (or a chain whose rows are ldh wide): straight copy
      copy_f4(hs, src, (B * ldh) >> 2);
      __syncthreads();
      return;
    }
    const bool vec = (lds & 3) == 0 && (((uintptr_t)src) & 15) == 0;
    for (int e = tid; e < BTS * p.CQ; e += NT) {
      const int b = e / p.CQ, j0 = 4 * (e - b * p.CQ);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < B) {
        const float* sp = src + (size_t)b * lds + j0;
        if (vec && j0 + 3 < lds) x = __ldcg(reinterpret_cast<const float4*>(sp));
        else {
          if (j0 < H) x.x = __ldcg(sp);
          if (j0 + 1 < H) x.y = __ldcg(sp + 1);
          if (j0 + 2 < H) x.z = __ldcg(sp + 2);
          if (j0 + 3 < H) x.w = __ldcg(sp + 3);
        }
        if (j0 + 3 >= H) {      // padding columns stay zero
          if (j0 >= H) x.x = 0.f;
          if (j0 + 1 >= H) x.y = 0.f;
          if (j0 + 2 >= H) x.z = 0.f;
          x.w = 0.f;
        }
      }
      *reinterpret_cast<float4*>(hs + b * ldh + j0) = x;
    }
    __syncthreads();
  };

  for (int step = 0; step < p.n_steps; ++step) {
    const int* idxp = p.idx ? p.idx + (size_t)step * B : nullptr;
    const int bit = p.pcd ? (bit0 + step) % V : 0;
    // ---- the whole minibatch into vbuf (transposed), the own rows into v0o ----
    if (tid < BTS) sidx[tid] = tid < B ? (idxp ? idxp[tid] : tid) : -1;
    __syncthreads();
    for (int i = tid; i < V; i += NT) {
#pragma unroll
      for (int b0 = 0; b0 < MAXB; b0 += 4) {
        if (b0 < BTS) {
          const long long r0 = sidx[b0], r1 = sidx[b0 + 1], r2 = sidx[b0 + 2], r3 = sidx[b0 + 3];
          float4 x;
          x.x = r0 >= 0 ? __ldg(&p.data[r0 * p.ld_data + i]) : 0.f;
          x.y = r1 >= 0 ? __ldg(&p.data[r1 * p.ld_data + i]) : 0.f;
          x.z = r2 >= 0 ? __ldg(&p.data[r2 * p.ld_data + i]) : 0.f;
          x.w = r3 >= 0 ? __ldg(&p.data[r3 * p.ld_data + i]) : 0.f;
          *reinterpret_cast<float4*>(vbuf + i * BTS + b0) = x;
        }
      }
    }
    __syncthreads();
    for (int e = tid; e < nr * BTS; e += NT) v0o[e] = vbuf[i0 * BTS + e];

    mark();   // gathered
    // =============================== positive phase (own columns) ===============================
    propup(false);
    {
      const RngSeg rs0 = seg(0, 0, step);
      for (int e = tid; e < nq * BTS * 4; e += NT) {
        const int qq = e / (BTS * 4), b = (e >> 2) % BTS, c = e & 3, j = 4 * (q0 + qq) + c;
        float mean = 0.f, smp = 0.f;      // padding rows / columns of the published panels are written as zeros
        if (b < B && j < H) {
          mean = sigmoid_fast_(preo[e] + __ldcg(&p.hb[j]));
          if (!p.pcd) smp = rng_uniform(rs0, (long long)b * H + j) < mean ? 1.f : 0.f;
        }
        __stcg(&p.PHg[b * ldh + j], mean);
        if (!p.pcd) __stcg(&p.Hg[b * ldh + j], smp);
        pho[e] = mean;
      }
    }
    if (p.pcd) {
      // pseudo-likelihood monitor (src/rbm.py:421-447), pre-update parameters: partial sums over the own columns of
      // softplus(pre_j) and softplus(pre_j + d W[bit][j]) for the rounded input, one pair per minibatch row
      __syncthreads();
      propup(true);
      float h0 = 0.f, h1 = 0.f;      // thread = (b, c) of the final reduction layout
      const int pb = tid >> 2, pc = tid & 3;
      if (tid < BTS * 4 && pb < B) {
        const float d = 1.f - 2.f * roundf(vbuf[bit * BTS + pb]);
        for (int qq = 0; qq < nq; ++qq) {
          const int j = 4 * (q0 + qq) + pc;
          if (j < H) {
            const float pre = preo[(qq * BTS + pb) * 4 + pc] + __ldcg(&p.hb[j]);
            h0 += softplusf_(pre);
            h1 += softplusf_(pre + d * __ldcg(&p.W[(size_t)bit * ldw + j]));
          }
        }
      }
      // the four column lanes of a row are adjacent threads (whole warps take part in the shuffles)
      h0 += __shfl_xor_sync(0xffffffffu, h0, 1); h0 += __shfl_xor_sync(0xffffffffu, h0, 2);
      h1 += __shfl_xor_sync(0xffffffffu, h1, 1); h1 += __shfl_xor_sync(0xffffffffu, h1, 2);
      if (tid < BTS * 4 && pc == 0) {
        __stcg(&p.pl_part[((size_t)cta * BTS + pb) * 2], h0);
        __stcg(&p.pl_part[((size_t)cta * BTS + pb) * 2 + 1], h1);
      }
    }
    mark();   // positive phase done
    grid_sync(p.bar, bar_target);
    mark();
    if (p.pcd && cta == 0 && warp == 0) {
      // monitor: fixed-order sum of the partials over the CTAs, one lane per minibatch row
      float cb = 0.f;
      if (lane < B) {
        float h0 = 0.f, h1 = 0.f;
        for (int g = 0; g < G; ++g) {
          h0 += __ldcg(&p.pl_part[((size_t)g * BTS + lane) * 2]);
          h1 += __ldcg(&p.pl_part[((size_t)g * BTS + lane) * 2 + 1]);
        }
        const float x = roundf(vbuf[bit * BTS + lane]), d = 1.f - 2.f * x, vbv = __ldcg(&p.vb[bit]);
        float vterm;
        if (p.kind == MDBN_GRBM) { const float a = x - vbv, c = (1.f - x) - vbv; vterm = 0.5f * (a * a - c * c); }
        else vterm = d * vbv;
        cb = -(float)V * softplusf_((h1 - h0) + vterm);
      }
      cb = warp_sum(cb);
      if (lane == 0 && p.cost_out) p.cost_out[step] = cb * p.cost_scale;
    }

    // =============================== k Gibbs steps ===============================
    float cost_acc = 0.f;
    for (int s = 0; s < p.k; ++s) {
      const bool last = (s == p.k - 1);
      const long long ubase = (long long)B * H + (long long)s * p.u_step_stride;
      const RngSeg rs_v = seg(ubase + p.u_off_v, 1u + 2u * s, step);
      const RngSeg rs_h = seg(ubase + p.u_off_h, 2u + 2u * s, step);
      // chain state: CD starts from the fresh sample, PCD from the persistent chain (src/rbm.py:308-311)
      if (p.pcd && s == 0) load_hidden(p.P, H);
      else load_hidden(p.Hg, ldh);
      // ---- propdown of the own rows: one warp per row, lanes over the column quads ----
      for (int r = warp; r < nr; r += NWARP) {
        const int i = i0 + r;
        float acc[MAXB];
#pragma unroll
        for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
        const float* wr = p.W + (size_t)i * ldw;
        float4 wreg[8];                                          // ldw <= 1024: all loads of the row in flight at once
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int q = lane + 32 * t;
          wreg[t] = q < p.CQ ? __ldcg(reinterpret_cast<const float4*>(wr + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float vb_i = __ldcg(&p.vb[i]);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int q = lane + 32 * t;
          if (q < p.CQ) {
            const float4 w = wreg[t];
#pragma unroll
            for (int b = 0; b < MAXB; ++b) {
              if (b < BTS) {
                const float4 h4 = *reinterpret_cast<const float4*>(hs + b * ldh + 4 * q);
                acc[b] = fmaf(h4.x, w.x, fmaf(h4.y, w.y, fmaf(h4.z, w.z, fmaf(h4.w, w.w, acc[b]))));
              }
            }
          }
        }
        float mine = 0.f;
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
          if (b < BTS) {
            const float t = warp_sum(acc[b]);
            if (lane == b) mine = t;
          }
        }
        if (lane < BTS) {
          const int b = lane;
          float v_in = 0.f, mean = 0.f;
          if (b < B) {
            const float pre = mine + vb_i;
            if (p.kind == MDBN_GRBM) {
              mean = pre;
              v_in = pre;                                        // mean-field visible (src/rbm.py:669)
            } else {
              mean = sigmoidf_(pre);
              v_in = rng_uniform(rs_v, (long long)b * V + i) < mean ? 1.f : 0.f;
            }
            if (last && !p.pcd) {
              const float t0 = v0o[r * BTS + b];
              if (p.kind == MDBN_GRBM) { const float d = sigmoidf_(pre) - t0; cost_acc += d * d; }      // :697
              else cost_acc += t0 * softplusf_(-pre) + (1.f - t0) * softplusf_(pre);                    // :479-480
            }
          }
          __stcg(&p.Vg[(size_t)i * BTS + b], v_in);
          if (last) nvo[r * BTS + b] = mean;
        }
      }
      mark();   // hidden panel + propdown
      if (last && !p.pcd) {
        const float c = block_sum(cost_acc, misc);
        if (tid == 0) __stcg(&p.cost_part[cta], c);
      }
      grid_sync(p.bar, bar_target);
      mark();
      // ---- the whole visible state, then the hidden units of the own columns ----
      copy_f4(vbuf, p.Vg, (V * BTS) >> 2);
      __syncthreads();
      propup(false);
      for (int e = tid; e < nq * BTS * 4; e += NT) {
        const int qq = e / (BTS * 4), b = (e >> 2) % BTS, c = e & 3, j = 4 * (q0 + qq) + c;
        float mean = 0.f, smp = 0.f;
        const bool real = b < B && j < H;
        if (real) {
          mean = sigmoid_fast_(preo[e] + __ldcg(&p.hb[j]));
          smp = rng_uniform(rs_h, (long long)b * H + j) < mean ? 1.f : 0.f;
        }
        if (!last) __stcg(&p.Hg[b * ldh + j], smp);
        else {
          __stcg(&p.NHg[b * ldh + j], mean);
          if (p.pcd && real) __stcg(&p.P[(size_t)b * H + j], smp);      // new persistent chain (src/rbm.py:372)
          nho[e] = mean;
        }
      }
      mark();   // visible panel + propup
      grid_sync(p.bar, bar_target);
      mark();
    }
    if (!p.pcd && cta == 0 && warp == 0) {                       // reconstruction cost: fixed-order sum of the CTA partials
      float c = 0.f;
      for (int g = lane; g < G; g += 32) c += __ldcg(&p.cost_part[g]);
      c = warp_sum(c);
      if (lane == 0 && p.cost_out) p.cost_out[step] = c * p.cost_scale;
    }

    // =============================== statistics + update of the own rows ===============================
    // positive / negative hidden means of ALL columns come from the published panels, HC columns at a time (parked in the
    // visible-state buffer, dead by now); v0 and the negative visible means of the own rows never left this CTA
    for (int c0 = 0; c0 < ldh; c0 += p.HC) {
      const int ncq = min(p.HC, ldh - c0) >> 2;          // quads in this chunk
      float* phc = vbuf;
      float* nhc = vbuf + BTS * p.HC;
      __syncthreads();
      if (p.HC == ldh) {            // one chunk: the panels are contiguous
        copy_f4(phc, p.PHg, (BTS * ldh) >> 2);
        copy_f4(nhc, p.NHg, (BTS * ldh) >> 2);
      } else {
        for (int e = tid; e < BTS * ncq; e += NT) {
          const int b = e / ncq, qd = e - b * ncq;
          reinterpret_cast<float4*>(phc + b * p.HC)[qd] = __ldcg(reinterpret_cast<const float4*>(p.PHg + b * ldh + c0) + qd);
          reinterpret_cast<float4*>(nhc + b * p.HC)[qd] = __ldcg(reinterpret_cast<const float4*>(p.NHg + b * ldh + c0) + qd);
        }
      }
      __syncthreads();
      const int groups = ncq <= 128 ? 2 : 1;             // narrow chunks: two rows in flight
      const int qd = tid % (NT / groups), rg = tid / (NT / groups);
      if (qd < ncq) {
        const int j0 = c0 + 4 * qd;
        for (int r = rg; r < nr; r += groups) {
          const size_t o = (size_t)(i0 + r) * ldw + j0;
          const float4 w4 = __ldcg(reinterpret_cast<const float4*>(p.W + o)), s4 = __ldcg(reinterpret_cast<const float4*>(p.S + o));
          const float4 n4 = p.Wsnap ? __ldcg(reinterpret_cast<const float4*>(p.Wsnap + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float g4[4] = {0.f, 0.f, 0.f, 0.f};
          for (int b = 0; b < B; ++b) {
            const float a = v0o[r * BTS + b], n = nvo[r * BTS + b];
            const float4 ph4 = *reinterpret_cast<const float4*>(phc + b * p.HC + 4 * qd);
            const float4 nh4 = *reinterpret_cast<const float4*>(nhc + b * p.HC + 4 * qd);
            g4[0] = fmaf(a, ph4.x, g4[0]); g4[0] = fmaf(-n, nh4.x, g4[0]);
            g4[1] = fmaf(a, ph4.y, g4[1]); g4[1] = fmaf(-n, nh4.y, g4[1]);
            g4[2] = fmaf(a, ph4.z, g4[2]); g4[2] = fmaf(-n, nh4.z, g4[2]);
            g4[3] = fmaf(a, ph4.w, g4[3]); g4[3] = fmaf(-n, nh4.w, g4[3]);
          }
          {
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w}, nv[4] = {n4.x, n4.y, n4.z, n4.w};
            float wo[4], so[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              wo[c] = wv[c]; so[c] = sv[c];
              if (j0 + c < H) update_one(p.u, g4[c], wv[c], sv[c], nv[c], p.Wsnap != nullptr, wo[c], so[c]);
            }
            *reinterpret_cast<float4*>(p.W + o) = make_float4(wo[0], wo[1], wo[2], wo[3]);
            *reinterpret_cast<float4*>(p.S + o) = make_float4(so[0], so[1], so[2], so[3]);
          }
        }
      }
    }
    // visible bias of the own rows (src/rbm.py:417), hidden bias of the own columns (:416)
    for (int r = tid; r < nr; r += NT) {
      float gs = 0.f;
      for (int b = 0; b < B; ++b) gs += v0o[r * BTS + b] - nvo[r * BTS + b];
      const float gb = gs * p.inv_b, sv = __ldcg(&p.Svb[i0 + r]);
      p.Svb[i0 + r] = gb + (sv - gb) * p.u.mom;
      p.vb[i0 + r] = __ldcg(&p.vb[i0 + r]) + sv * p.u.lr;
    }
    for (int e = tid; e < nq * 4; e += NT) {
      const int qq = e >> 2, c = e & 3, j = 4 * (q0 + qq) + c;
      if (j < H) {
        float gs = 0.f;
        for (int b = 0; b < B; ++b) gs += pho[(qq * BTS + b) * 4 + c] - nho[(qq * BTS + b) * 4 + c];
        const float gb = gs * p.inv_b, sv = __ldcg(&p.Shb[j]);
        p.Shb[j] = gb + (sv - gb) * p.u.mom;
        p.hb[j] = __ldcg(&p.hb[j]) + sv * p.u.lr;
      }
    }
    mark();   // statistics + update done
    // the next step of a chained launch reads columns of W that other CTAs have just written, and overwrites the panels
    if (step + 1 < p.n_steps) grid_sync(p.bar, bar_target);
  }
  if (p.pcd && cta == 0 && tid == 0) *p.bit_idx = (bit0 + p.n_steps) % V;     // src/rbm.py:445

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

struct Geometry {
  int BTS, CQ, ldh, NQ, NR, HC, grid;
  int off_vbuf, off_hs, off_v0o, off_nvo, off_red, off_preo, off_pho, off_nho, off_misc, off_wcol;
  size_t smem;
  bool ok;
};

static Geometry plan(const mdbn_ctx* c, const mdbn_cd_args& a) {
  Geometry g{};
  g.ok = false;
  if (a.B < 1 || a.B > MAXB || a.ldw % 4 != 0) return g;     // (a noisy GRBM still feeds the visible MEAN upwards, src/rbm.py:669)
  if (((uintptr_t)a.W | (uintptr_t)a.W_speed | (uintptr_t)a.W_snap) & 15) return g;
  static const int og = getenv("MDBN_MID_GRID") ? atoi(getenv("MDBN_MID_GRID")) : 0;
  g.grid = og > 0 && og < c->num_sms ? og : c->num_sms;
  g.BTS = (a.B + 3) & ~3;
  g.CQ = a.ldw / 4;
  g.ldh = a.ldw;
  g.NQ = (g.CQ + g.grid - 1) / g.grid;
  g.NR = (a.V + g.grid - 1) / g.grid;
  g.HC = g.ldh < 512 ? g.ldh : 512;
  if (g.NQ > 4 || g.NR > 32) return g;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 127) & ~(size_t)127; return (int)o; };
  const size_t vis = (size_t)a.V * g.BTS * 4, chunks = (size_t)2 * g.BTS * g.HC * 4;
  g.off_vbuf = take(vis > chunks ? vis : chunks);
  g.off_hs = take((size_t)g.BTS * g.ldh * 4);
  g.off_v0o = take((size_t)g.NR * g.BTS * 4);
  g.off_nvo = take((size_t)g.NR * g.BTS * 4);
  g.off_red = take((size_t)NT * 16 * 4 + 3 * MAXB * 4 * 4);
  g.off_preo = take((size_t)g.NQ * g.BTS * 16);
  g.off_pho = take((size_t)g.NQ * g.BTS * 16);
  g.off_nho = take((size_t)g.NQ * g.BTS * 16);
  g.off_misc = take(512);
  g.off_wcol = take((size_t)a.V * 16);
  g.smem = off;
  g.ok = g.smem <= 200 * 1024;
  return g;
}

}  // namespace md

bool mid_supported(const mdbn_ctx* c, const mdbn_cd_args& a) {
  if (a.phase != MDBN_PHASE_FULL) return false;
  return md::plan(c, a).ok;
}

// n_steps consecutive steps in one launch (a.indices [n_steps][B], a.cost_out [n_steps]; PHILOX when n_steps > 1)
int mid_cd_steps(mdbn_ctx* c, const mdbn_cd_args& a, int n_steps, cudaStream_t st) {
  md::Geometry g = md::plan(c, a);
  MDBN_CHECK(g.ok, "mid path: unsupported shape");
  MDBN_CHECK(n_steps >= 1, "mid path: n_steps must be >= 1");
  MDBN_CHECK(n_steps == 1 || a.rng.mode == MDBN_RNG_PHILOX, "mid path: chained steps need the PHILOX generator");
  md::Params p{};
  p.W = a.W; p.S = a.W_speed; p.Wsnap = a.weightcost != 0.f ? a.W_snap : nullptr; p.ldw = a.ldw;
  p.hb = a.hbias; p.vb = a.vbias; p.Shb = a.hbias_speed; p.Svb = a.vbias_speed;
  p.data = a.data; p.ld_data = a.ld_data; p.idx = a.indices;
  p.P = a.persistent; p.bit_idx = a.bit_i_idx; p.cost_out = a.cost_out;
  p.kind = a.kind; p.B = a.B; p.V = a.V; p.H = a.H; p.k = a.k; p.pcd = a.persistent != nullptr; p.n_steps = n_steps;
  p.u = make_update_scalars(a);
  p.inv_b = 1.0f / (float)a.B;
  p.cost_scale = (!p.pcd && a.kind == MDBN_GRBM) ? 1.0f / ((float)a.B * (float)a.V) : 1.0f / (float)a.B;
  p.rng_mode = a.rng.mode;
  p.ubuf = a.rng.mode == MDBN_RNG_BUFFER ? a.rng.buffer : nullptr;
  p.k0 = (uint32_t)a.rng.seed; p.k1 = (uint32_t)(a.rng.seed >> 32);
  p.c2 = (uint32_t)a.rng.offset; p.c3 = (uint32_t)(a.rng.offset >> 32);
  ULayout ul = u_layout(a.kind, a.noisy, a.B, a.V, a.H);
  p.u_step_stride = ul.step_stride; p.u_off_v = ul.off_v; p.u_off_h = ul.off_h;
  p.BTS = g.BTS; p.CQ = g.CQ; p.ldh = g.ldh; p.NQ = g.NQ; p.NR = g.NR; p.HC = g.HC;
  p.off_vbuf = g.off_vbuf; p.off_hs = g.off_hs; p.off_v0o = g.off_v0o; p.off_nvo = g.off_nvo; p.off_red = g.off_red;
  p.off_preo = g.off_preo; p.off_pho = g.off_pho; p.off_nho = g.off_nho; p.off_misc = g.off_misc; p.off_wcol = g.off_wcol;
  // published panels + cost partials
  const size_t hid = (size_t)g.BTS * g.ldh, vis = (size_t)a.V * g.BTS;
  const size_t total = (3 * hid + vis + (size_t)g.grid + (size_t)g.grid * g.BTS * 2 + 64) * sizeof(float);
  float* base = (float*)ws_get(c, WS_TENSOR, total);
  if (!base) return 3;
  p.Hg = base; p.PHg = base + hid; p.NHg = base + 2 * hid; p.Vg = base + 3 * hid;
  p.cost_part = p.Vg + vis;
  p.pl_part = p.cost_part + ((g.grid + 3) & ~3);
  p.bar = reinterpret_cast<unsigned long long*>(c->barrier);
  static const bool want_timing = getenv("MDBN_MID_TIMING") != nullptr;
  p.dbg = want_timing ? reinterpret_cast<unsigned long long*>(c->barrier) + 64 : nullptr;
  static bool configured[64] = {};
  if (!configured[c->device]) {
    MDBN_CUDA(cudaFuncSetAttribute(md::cd_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[c->device] = true;
  }
  void* args[] = {(void*)&p};
  MDBN_CUDA(cudaLaunchCooperativeKernel((void*)md::cd_mid_kernel, dim3(g.grid), dim3(md::NT), args, g.smem, st));
  c->launches++;
  if (p.dbg) {
    unsigned long long t[24];
    MDBN_CUDA(cudaStreamSynchronize(st));
    MDBN_CUDA(cudaMemcpy(t, p.dbg, sizeof(t), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[mid timeline us] V=%d H=%d B=%d k=%d grid=%d NQ=%d NR=%d:", a.V, a.H, a.B, a.k, g.grid, g.NQ, g.NR);
    for (int i = 1; i < 12; ++i) fprintf(stderr, " %.1f", (double)((long long)(t[i] - t[0])) * 1e-3);
    fprintf(stderr, "\n");
  }
  return 0;
}

}  // namespace mdbn
