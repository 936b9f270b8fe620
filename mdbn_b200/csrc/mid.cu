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
//   between two propagations: ONE grid barrier, then every CTA copies the 40-135 KB panel it needs from L2 (cp.async);
//   minibatch: every CTA gathers only its OWN rows and publishes them as rows of the visible panel (in a chained launch
//              ahead of the barrier that ends the previous step); the whole minibatch then arrives as a panel copy;
//   update   : the row owner updates its rows of W / W_speed (it kept its rows of v0 and of the negative means), reading
//              the positive / negative hidden means of all columns from the published panels; biases by their owners;
//   PCD      : pseudo-likelihood partials per CTA, one minibatch row summed on each of the last CTAs of the grid.
// No atomics, no redundant rebuild; 2k + 2 barriers per step (+ 1 at the start of a launch), fixed summation orders.  Same
// arithmetic contract as the other paths (fp32, src/rbm.py semantics, Philox draws indexed by element or the caller's
// uniform buffer).  Measured: DESIGN.md 4.4, profiles/r2_mid_kernel_*.txt.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include "ctx.h"

namespace mdbn {
namespace md {

constexpr int NT = 256, NWARP = NT / 32;
constexpr int MAXB = 20;
constexpr size_t MAX_SMEM = 227 * 1024;

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
  int wcol_cached;                  // the own column quads of W stay in shared memory for a whole step
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

// Asynchronous global -> shared copies (LDGSTS): every request of a panel is in flight at once, no register staging.
// .cg reads through L2 (panels and weights other CTAs have just written), .ca may keep the read-only dataset in L1.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ float4 lds128v(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}

// Sum over the 32 lanes of v[0..MAXB) at once: lane L returns the warp total of v[L] (0 for L >= MAXB).  31 shuffles in
// five levels instead of 5 per value; fixed order.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int s = 0; s < w; ++s) {
      const float send = up ? v[s] : v[s + w], keep = up ? v[s + w] : v[s];
      v[s] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

// Device-wide barrier (all CTAs co-resident: cooperative launch); the counter is monotonic within a launch and reset by
// the last CTA to leave the kernel.
__device__ __forceinline__ void grid_sync(unsigned long long* bar, unsigned long long& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");      // no round trip before the spin
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

  float* hbo = misc + 32;                                      // [NQ][4] hidden biases of the own columns (this step's)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, G = gridDim.x;
  const int B = p.B, V = p.V, H = p.H, BTS = p.BTS, ldh = p.ldh, ldw = p.ldw;
  const int q0 = cta * p.NQ, nq = max(0, min(p.NQ, p.CQ - q0));      // own column quads
  const int i0 = cta * p.NR, nr = max(0, min(p.NR, V - i0));         // own visible rows
  unsigned long long bar_target = 0;
  const int bit0 = p.pcd ? *p.bit_idx : 0;
  int dbg_i = 0;
  auto mark = [&]() {
    if (p.dbg && cta == 0 && tid == 0 && dbg_i < 24) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[dbg_i++] = t;
    }
  };
  if (p.dbg && cta == 0 && tid < 24) p.dbg[tid] = 0ULL;
  mark();
  // global -> shared copy of n4 float4 (both 16-byte aligned), all requests in flight at once; cp_async_wait() completes it
  auto copy_f4 = [&](float* dst, const float* src, int n4) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
    for (int e = tid; e < n4; e += NT) cp_async16(d4 + e, s4 + e);
  };

  // The own rows of minibatch `step` (v0o, and published as rows of the [V][BTS] visible panel).  Every CTA reading the
  // whole minibatch out of the dataset measured ~5 us whatever the access width; the panel copy after a barrier is 0.6 us,
  // and in a chained launch the barrier is the one that ends the previous step anyway.
  auto gather_slice = [&](int step) {
    const int* idxp = p.idx ? p.idx + (size_t)step * B : nullptr;
    for (int e = tid; e < nr * BTS; e += NT) {
      const int r = e / BTS, b = e - r * BTS;
      float x = 0.f;
      if (b < B) {
        const long long row = idxp ? __ldg(idxp + b) : b;
        x = __ldg(p.data + row * p.ld_data + i0 + r);
      }
      v0o[e] = x;
      __stcg(&p.Vg[(size_t)i0 * BTS + e], x);
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
      // the column quad of all V rows out of shared memory: staged once per step (stage_columns) where the own quads fit,
      // else fetched here, every thread its share at once (one L2 round trip)
      const float4* wq = wcol + (p.wcol_cached ? (size_t)qq * V : 0);
      if (!p.wcol_cached) {
        const float* wp = p.W + 4 * (q0 + qq);
        for (int i = tid; i < V; i += NT) wcol[i] = __ldcg(reinterpret_cast<const float4*>(wp + (size_t)i * ldw));
        __syncthreads();
      }
      if (ch < NCH) {
#pragma unroll 4
        for (int i = ch; i < V; i += NCH) {
          const float4 w = wq[i];
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
      cp_async_wait();
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

  // Propdown work items: (row, part of its column quads) — a whole row per warp when the CTA owns >= 5 rows, else the row
  // split over 2 or 4 warps.  A lane holds quads lane + 32 t of its part.  The first item of every warp goes into registers
  // ahead of the grid barrier in front of a propdown (W does not change between the statistics passes).
  const int NTQ = (p.CQ + 31) >> 5;                                 // 32-quad blocks in a row
  const int SPLIT = nr > 4 ? 1 : nr > 2 ? 2 : 4, TPS = (NTQ + SPLIT - 1) / SPLIT, nitem = nr * SPLIT;
  float* part = red;                                               // [nitem][BTS] partial dot products
  auto load_item = [&](int it, float4 (&w)[8]) {
    if (it < nitem) {
      const int r = it / SPLIT, t0 = (it - r * SPLIT) * TPS;
      const float* wr = p.W + (size_t)(i0 + r) * ldw;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int q = lane + 32 * (t0 + t);
        w[t] = (t < TPS && q < p.CQ) ? __ldcg(reinterpret_cast<const float4*>(wr + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  float4 wreg[8];
  const uint32_t hs_u32 = (uint32_t)__cvta_generic_to_shared(hs);
  float vb_pre = 0.f;                                              // visible bias of this thread's first propdown output

  for (int step = 0; step < p.n_steps; ++step) {
    const int bit = p.pcd ? (bit0 + step) % V : 0;
    // ---- the minibatch (published slice by slice before the barrier that ended the previous step) and the own column
    //      quads of W into shared memory ----
    if (step == 0) {
      gather_slice(0);
      mark();   // own rows of the minibatch published
      grid_sync(p.bar, bar_target);
      mark();
    }
    if (p.wcol_cached) {
      for (int qq = 0; qq < nq; ++qq) {
        const float* wp = p.W + 4 * (q0 + qq);
        float4* wq = wcol + (size_t)qq * V;
#pragma unroll 4
        for (int i = tid; i < V; i += NT) cp_async16(wq + i, wp + (size_t)i * ldw);
      }
    }
    copy_f4(vbuf, p.Vg, (V * BTS) >> 2);
    if (tid < nq * 4) hbo[tid] = 4 * q0 + tid < H ? __ldcg(&p.hb[4 * q0 + tid]) : 0.f;
    cp_async_wait();
    __syncthreads();

    mark();   // gathered
    // =============================== positive phase (own columns) ===============================
    propup(false);
    {
      const RngSeg rs0 = seg(0, 0, step);
      for (int e = tid; e < nq * BTS * 4; e += NT) {
        const int qq = e / (BTS * 4), b = (e >> 2) % BTS, c = e & 3, j = 4 * (q0 + qq) + c;
        float mean = 0.f, smp = 0.f;      // padding rows / columns of the published panels are written as zeros
        if (b < B && j < H) {
          mean = sigmoid_fast_(preo[e] + hbo[qq * 4 + c]);
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
            const float pre = preo[(qq * BTS + pb) * 4 + pc] + hbo[qq * 4 + pc];
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
    load_item(warp, wreg);
    if (tid < nr * BTS) vb_pre = __ldcg(&p.vb[i0 + tid / BTS]);
    grid_sync(p.bar, bar_target);
    mark();
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
      mark();   // hidden panel
      if (p.pcd && s == 0 && warp == NWARP - 1) {
        // monitor: minibatch row pl_b on one of the LAST CTAs of the grid (they own no or few hidden columns), on the
        // warp that has a propdown item only on CTAs with 8 rows or more: the partials of all CTAs, lane-strided and
        // joined by a butterfly (fixed order); v0 is still in vbuf
        const int pl_b = G - 1 - cta;
        if (pl_b < B) {
          float h0 = 0.f, h1 = 0.f;
          for (int g = lane; g < G; g += 32) {
            const float2 t = __ldcg(reinterpret_cast<const float2*>(p.pl_part + ((size_t)g * BTS + pl_b) * 2));
            h0 += t.x; h1 += t.y;
          }
          h0 = warp_sum(h0); h1 = warp_sum(h1);
          if (lane == 0) {
            const float x = roundf(vbuf[bit * BTS + pl_b]), d = 1.f - 2.f * x, vbv = __ldcg(&p.vb[bit]);
            float vterm;
            if (p.kind == MDBN_GRBM) { const float a = x - vbv, c = (1.f - x) - vbv; vterm = 0.5f * (a * a - c * c); }
            else vterm = d * vbv;
            __stcg(&p.cost_part[pl_b], -(float)V * softplusf_((h1 - h0) + vterm));
          }
        }
      }
      // ---- propdown of the own rows: the lanes of a warp sum their quads of W against the hidden panel for every
      //      minibatch row and are joined by ONE transposing reduction per item; split rows meet in shared memory ----
      for (int it = warp; it < nitem; it += NWARP) {
        const int t0 = (it % SPLIT) * TPS;
        if (it != warp) load_item(it, wreg);                       // (only CTAs with more than 8 rows: V > 1184)
        float acc[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) acc[b] = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int q = lane + 32 * (t0 + t);
          if (t < TPS && 32 * (t0 + t) < p.CQ) {                     // (warp-uniform)
            const float4 w = wreg[t];
            const uint32_t ha = hs_u32 + 16u * (uint32_t)(q < p.CQ ? q : 0);      // (w is zero beyond the last quad)
#pragma unroll
            for (int bh = 0; bh < MAXB; bh += MAXB / 2) {          // ten loads in flight (volatile: issued back to back)
              float4 h[MAXB / 2];
#pragma unroll
              for (int b = 0; b < MAXB / 2; ++b)
                if (bh + b < BTS) h[b] = lds128v(ha + (uint32_t)((bh + b) * ldh) * 4u);
#pragma unroll
              for (int b = 0; b < MAXB / 2; ++b)
                if (bh + b < BTS)
                  acc[bh + b] = fmaf(h[b].x, w.x, fmaf(h[b].y, w.y, fmaf(h[b].z, w.z, fmaf(h[b].w, w.w, acc[bh + b]))));
            }
          }
        }
        const float tot = warp_transpose_sum(acc, lane);
        if (lane < BTS) part[it * BTS + lane] = tot;
      }
      mark();   // propdown sums
      __syncthreads();
      for (int e = tid; e < nr * BTS; e += NT) {
        const int r = e / BTS, b = e - r * BTS, i = i0 + r;
        float v_in = 0.f, mean = 0.f;
        if (b < B) {
          float pre = 0.f;
          for (int sp = 0; sp < SPLIT; ++sp) pre += part[(r * SPLIT + sp) * BTS + b];
          pre += e == tid ? vb_pre : __ldcg(&p.vb[i]);
          if (p.kind == MDBN_GRBM) {
            mean = pre;
            v_in = pre;                                        // mean-field visible (src/rbm.py:669)
          } else {
            mean = sigmoidf_(pre);
            v_in = rng_uniform(rs_v, (long long)b * V + i) < mean ? 1.f : 0.f;
          }
          if (last && !p.pcd) {
            const float t0 = v0o[e];
            if (p.kind == MDBN_GRBM) { const float d = sigmoidf_(pre) - t0; cost_acc += d * d; }      // :697
            // cross entropy t0 softplus(-pre) + (1 - t0) softplus(pre) (:479-480), with softplus(-x) = softplus(x) - x
            else cost_acc += softplusf_(pre) - t0 * pre;
          }
        }
        __stcg(&p.Vg[(size_t)i0 * BTS + e], v_in);
        if (last) nvo[e] = mean;
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
      cp_async_wait();
      __syncthreads();
      mark();   // visible panel
      propup(false);
      for (int e = tid; e < nq * BTS * 4; e += NT) {
        const int qq = e / (BTS * 4), b = (e >> 2) % BTS, c = e & 3, j = 4 * (q0 + qq) + c;
        float mean = 0.f, smp = 0.f;
        const bool real = b < B && j < H;
        if (real) {
          mean = sigmoid_fast_(preo[e] + hbo[qq * 4 + c]);
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
      if (!last) load_item(warp, wreg);
      grid_sync(p.bar, bar_target);
      mark();
    }
    if (cta == 0 && warp == 0) {      // reconstruction cost: fixed-order sum of the CTA partials; PCD: of the monitor's rows
      float c = 0.f;
      for (int g = lane; g < (p.pcd ? B : G); g += 32) c += __ldcg(&p.cost_part[g]);
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
      const int groups = ncq <= 128 ? 2 : 1;             // narrow chunks: two rows in flight
      const int qd = tid % (NT / groups), rg = tid / (NT / groups);
      const int j0 = c0 + 4 * qd;
      const bool active = qd < ncq;
      // A thread takes RB rows at a time for its column quad: the W / W_speed / snapshot quads of the first block (all the
      // rows of a thread up to V = 1184) are requested ahead of the panel copy; the hidden means of a minibatch row are read once
      // for the RB rows, v0 / nv of a row as quads over the minibatch
      constexpr int RB = 4;
      float4 w4[RB], s4[RB], n4[RB];
      auto load_block = [&](int m0, float4 (&w)[RB], float4 (&sp)[RB], float4 (&n)[RB]) {
#pragma unroll
        for (int m = 0; m < RB; ++m) {
          const int r = rg + groups * (m0 + m);
          n[m] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (active && r < nr) {
            const size_t o = (size_t)(i0 + r) * ldw + j0;
            w[m] = __ldcg(reinterpret_cast<const float4*>(p.W + o));
            sp[m] = __ldcg(reinterpret_cast<const float4*>(p.S + o));
            if (p.Wsnap) n[m] = __ldcg(reinterpret_cast<const float4*>(p.Wsnap + o));
          }
        }
      };
      load_block(0, w4, s4, n4);
      __syncthreads();
      if (p.HC == ldh) {            // one chunk: the panels are contiguous
        copy_f4(phc, p.PHg, (BTS * ldh) >> 2);
        copy_f4(nhc, p.NHg, (BTS * ldh) >> 2);
      } else {
        for (int e = tid; e < BTS * ncq; e += NT) {
          const int b = e / ncq, qe = e - b * ncq;
          cp_async16(reinterpret_cast<float4*>(phc + b * p.HC) + qe, reinterpret_cast<const float4*>(p.PHg + b * ldh + c0) + qe);
          cp_async16(reinterpret_cast<float4*>(nhc + b * p.HC) + qe, reinterpret_cast<const float4*>(p.NHg + b * ldh + c0) + qe);
        }
      }
      if (c0 == 0) {
        // under the panel copy (different warps, so that their L2 round trips overlap): visible bias of the own rows (src/rbm.py:417), hidden bias of the own columns (:416)
        for (int r = tid; r < nr; r += NT) {
          const float sv = __ldcg(&p.Svb[i0 + r]), vbo = __ldcg(&p.vb[i0 + r]);
          float gs = 0.f;
          for (int b0 = 0; b0 < BTS; b0 += 4) {      // (rows b >= B are zeros in both slabs)
            const float4 a4 = *reinterpret_cast<const float4*>(v0o + r * BTS + b0);
            const float4 m4 = *reinterpret_cast<const float4*>(nvo + r * BTS + b0);
            gs += ((a4.x - m4.x) + (a4.y - m4.y)) + ((a4.z - m4.z) + (a4.w - m4.w));
          }
          const float gb = gs * p.inv_b;
          p.Svb[i0 + r] = gb + (sv - gb) * p.u.mom;
          p.vb[i0 + r] = vbo + sv * p.u.lr;
        }
        for (int e = tid - NT / 2; e >= 0 && e < nq * 4; e += NT) {      // (nq <= 4)
          const int qq = e >> 2, c = e & 3, j = 4 * (q0 + qq) + c;
          if (j < H) {
            const float sv = __ldcg(&p.Shb[j]);
            float gs = 0.f;
#pragma unroll 4
            for (int b = 0; b < B; ++b) gs += pho[(qq * BTS + b) * 4 + c] - nho[(qq * BTS + b) * 4 + c];
            const float gb = gs * p.inv_b;
            p.Shb[j] = gb + (sv - gb) * p.u.mom;
            p.hb[j] = hbo[e] + sv * p.u.lr;
          }
        }
      }
      cp_async_wait();
      __syncthreads();
      mark();   // mean panels
      if (active) {
        for (int m0 = 0; rg + groups * m0 < nr; m0 += RB) {
          if (m0 > 0) load_block(m0, w4, s4, n4);
          float g[RB][4];
#pragma unroll
          for (int m = 0; m < RB; ++m) g[m][0] = g[m][1] = g[m][2] = g[m][3] = 0.f;
          for (int b0 = 0; b0 < BTS; b0 += 4) {      // (rows b >= B of every panel are zeros)
            float av[RB][4], nv[RB][4];
#pragma unroll
            for (int m = 0; m < RB; ++m) {
              const int r = min(rg + groups * (m0 + m), nr - 1);
              const float4 a4 = *reinterpret_cast<const float4*>(v0o + r * BTS + b0);
              const float4 m4 = *reinterpret_cast<const float4*>(nvo + r * BTS + b0);
              av[m][0] = a4.x; av[m][1] = a4.y; av[m][2] = a4.z; av[m][3] = a4.w;
              nv[m][0] = m4.x; nv[m][1] = m4.y; nv[m][2] = m4.z; nv[m][3] = m4.w;
            }
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
              const float4 ph4 = *reinterpret_cast<const float4*>(phc + (b0 + bb) * p.HC + 4 * qd);
              const float4 nh4 = *reinterpret_cast<const float4*>(nhc + (b0 + bb) * p.HC + 4 * qd);
#pragma unroll
              for (int m = 0; m < RB; ++m) {
                g[m][0] = fmaf(av[m][bb], ph4.x, g[m][0]); g[m][0] = fmaf(-nv[m][bb], nh4.x, g[m][0]);
                g[m][1] = fmaf(av[m][bb], ph4.y, g[m][1]); g[m][1] = fmaf(-nv[m][bb], nh4.y, g[m][1]);
                g[m][2] = fmaf(av[m][bb], ph4.z, g[m][2]); g[m][2] = fmaf(-nv[m][bb], nh4.z, g[m][2]);
                g[m][3] = fmaf(av[m][bb], ph4.w, g[m][3]); g[m][3] = fmaf(-nv[m][bb], nh4.w, g[m][3]);
              }
            }
          }
#pragma unroll
          for (int m = 0; m < RB; ++m) {
            const int r = rg + groups * (m0 + m);
            if (r < nr) {
              const size_t o = (size_t)(i0 + r) * ldw + j0;
              const float wv[4] = {w4[m].x, w4[m].y, w4[m].z, w4[m].w}, sv[4] = {s4[m].x, s4[m].y, s4[m].z, s4[m].w};
              const float sn4[4] = {n4[m].x, n4[m].y, n4[m].z, n4[m].w};
              float wo[4], so[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                wo[c] = wv[c]; so[c] = sv[c];
                if (j0 + c < H) update_one(p.u, g[m][c], wv[c], sv[c], sn4[c], p.Wsnap != nullptr, wo[c], so[c]);
              }
              *reinterpret_cast<float4*>(p.W + o) = make_float4(wo[0], wo[1], wo[2], wo[3]);
              *reinterpret_cast<float4*>(p.S + o) = make_float4(so[0], so[1], so[2], so[3]);
            }
          }
        }
      }
    }
    mark();   // statistics + update done
    // the next step of a chained launch reads columns of W that other CTAs have just written, and overwrites the panels
    if (step + 1 < p.n_steps) {
      __syncthreads();                 // v0o is dead (bias update done)
      gather_slice(step + 1);
      grid_sync(p.bar, bar_target);
    }
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
  int BTS, CQ, ldh, NQ, NR, HC, grid, wcol_cached;
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
  g.grid = og >= MAXB && og < c->num_sms ? og : c->num_sms;      // (the monitor of PCD takes one CTA per minibatch row)
  if (g.grid < MAXB) return g;
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
  const size_t red_up = (size_t)NT * 16 * 4 + 3 * MAXB * 4 * 4;                      // propup partials
  const size_t red_down = (size_t)(g.NR > 4 ? g.NR : 4 * g.NR) * g.BTS * 4;           // propdown partials
  g.off_red = take(red_up > red_down ? red_up : red_down);
  g.off_preo = take((size_t)g.NQ * g.BTS * 16);
  g.off_pho = take((size_t)g.NQ * g.BTS * 16);
  g.off_nho = take((size_t)g.NQ * g.BTS * 16);
  g.off_misc = take(512);
  // all own column quads where they fit (they are then fetched once per step), else one quad restaged per propagation
  const size_t fixed = off;
  g.wcol_cached = fixed + (size_t)g.NQ * a.V * 16 <= MAX_SMEM ? 1 : 0;
  g.off_wcol = take((size_t)(g.wcol_cached ? g.NQ : 1) * a.V * 16);
  g.smem = off;
  g.ok = g.smem <= MAX_SMEM;
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
  p.wcol_cached = g.wcol_cached;
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
    MDBN_CUDA(cudaFuncSetAttribute(md::cd_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)md::MAX_SMEM));
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
    // own minibatch rows published, barrier, minibatch + W columns in place, positive, barrier, hidden panel, propdown sums,
    // propdown epilogue, barrier, visible panel, propup, barrier, mean panels, statistics  (k = 1, one statistics chunk)
    for (int i = 1; i < 15 && t[i] >= t[0]; ++i) fprintf(stderr, " %.1f", (double)((long long)(t[i] - t[0])) * 1e-3);
    fprintf(stderr, "\n");
  }
  return 0;
}

}  // namespace mdbn
