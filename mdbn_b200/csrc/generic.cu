// Generic fp32 path: any B, V, H, any strides.  Classic smem-tiled SIMT GEMMs with
// split-K + small fused epilogue kernels.  This is the fallback for shapes the
// persistent skinny kernel (skinny.cu) and the tcgen05 path (tensor.cu) do not
// take, and the implementation of the single-phase API calls (propup, propdown,
// free_energy) at arbitrary batch.  Deterministic: no atomics anywhere.
#include "ctx.h"

namespace mdbn {

// ---------------------------------------------------------------------------
// C_part[z] = opA(A) * opB(Bm) over this split's K range.
//   TA=false: A[m*lda+k]   TA=true: A[k*lda+m], rows k >= kneg enter with a minus sign
//   TB=false: Bm[k*ldb+n]  TB=true: Bm[n*ldb+k]
// ---------------------------------------------------------------------------
constexpr int GBM = 64, GBN = 64, GBK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, long long lda,
                                                    const float* __restrict__ Bm, long long ldb,
                                                    float* __restrict__ part, int M, int N, int K,
                                                    int k_per_split, int kneg) {
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
  float acc[4][4] = {};
  for (int k0 = kbeg; k0 < kend; k0 += GBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int m, kk;
      if (TA) { kk = e >> 6; m = e & 63; } else { m = e >> 4; kk = e & 15; }
      int gm = m0 + m, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < kend) {
        v = TA ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk];
        if (TA && gk >= kneg) v = -v;
      }
      As[kk][m] = v;
      int n;
      if (TB) { n = e >> 4; kk = e & 15; } else { kk = e >> 6; n = e & 63; }
      int gn = n0 + n;
      gk = k0 + kk;
      float w = 0.f;
      if (gn < N && gk < kend) w = TB ? Bm[(long long)gn * ldb + gk] : Bm[(long long)gk * ldb + gn];
      Bs[kk][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = part + (size_t)blockIdx.z * M * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn < N) out[(size_t)gm * N + gn] = acc[i][j];
    }
  }
}

static int pick_splits(const mdbn_ctx* c, int M, int N, int K) {
  int tiles = ((M + GBM - 1) / GBM) * ((N + GBN - 1) / GBN);
  int want = (2 * c->num_sms + tiles - 1) / tiles;
  int maxs = (K + 4 * GBK - 1) / (4 * GBK);
  int s = want < 1 ? 1 : want;
  if (s > maxs) s = maxs;
  if (s > 64) s = 64;
  return s < 1 ? 1 : s;
}

template <bool TA, bool TB>
static int launch_sgemm(mdbn_ctx* c, const float* A, long long lda, const float* Bm, long long ldb, float* part,
                        int M, int N, int K, int splits, int kneg, cudaStream_t st) {
  int kps = (K + splits - 1) / splits;
  kps = ((kps + GBK - 1) / GBK) * GBK;
  dim3 grid((N + GBN - 1) / GBN, (M + GBM - 1) / GBM, splits);
  sgemm_kernel<TA, TB><<<grid, 256, 0, st>>>(A, lda, Bm, ldb, part, M, N, K, kps, kneg);
  c->launches++;
  MDBN_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// epilogue: sum split-K partials + bias, activation, sampling
// ---------------------------------------------------------------------------
enum { ACT_SIGMOID = 0, ACT_LINEAR = 1 };
enum { SMP_NONE = 0, SMP_BERNOULLI = 1, SMP_MEAN = 2, SMP_GAUSS = 3 };

__global__ void act_epilogue_kernel(const float* __restrict__ part, int splits, int M, int N,
                                    const float* __restrict__ bias, int act, int smp, RngSeg rs,
                                    float* __restrict__ pre, long long ld_pre, float* __restrict__ mean,
                                    long long ld_mean, float* __restrict__ sample, long long ld_sample) {
  long long total = (long long)M * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int m = (int)(e / N), n = (int)(e % N);
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(size_t)z * total + e];
    s += bias[n];
    float mu = act == ACT_SIGMOID ? sigmoidf_(s) : s;
    if (pre) pre[m * ld_pre + n] = s;
    if (mean) mean[m * ld_mean + n] = mu;
    if (sample) {
      float x;
      if (smp == SMP_BERNOULLI) x = rng_uniform(rs, e) < mu ? 1.f : 0.f;
      else if (smp == SMP_GAUSS) x = mu + rng_normal(rs, e);
      else x = mu;
      sample[m * ld_sample + n] = x;
    }
  }
}

static int launch_epilogue(mdbn_ctx* c, const float* part, int splits, int M, int N, const float* bias, int act,
                           int smp, const RngSeg& rs, float* pre, long long ld_pre, float* mean, long long ld_mean,
                           float* sample, long long ld_sample, cudaStream_t st) {
  long long total = (long long)M * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * c->num_sms) blocks = 4 * c->num_sms;
  if (blocks < 1) blocks = 1;
  act_epilogue_kernel<<<blocks, 256, 0, st>>>(part, splits, M, N, bias, act, smp, rs, pre, ld_pre, mean, ld_mean,
                                              sample, ld_sample);
  c->launches++;
  MDBN_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// single-phase API
// ---------------------------------------------------------------------------
int generic_propup(mdbn_ctx* c, const float* W, int ldw, const float* hb, const float* v, int ldv, int B, int V,
                   int H, float* pre, float* mean, float* sample, const RngSeg& rs, cudaStream_t st) {
  int splits = pick_splits(c, B, H, V);
  float* part = (float*)ws_get(c, WS_PART, (size_t)splits * B * H * sizeof(float));
  if (!part) return 3;
  MDBN_TRY((launch_sgemm<false, false>(c, v, ldv, W, ldw, part, B, H, V, splits, 1 << 30, st)));
  return launch_epilogue(c, part, splits, B, H, hb, ACT_SIGMOID, sample ? SMP_BERNOULLI : SMP_NONE, rs, pre, H,
                         mean, H, sample, H, st);
}

int generic_propdown(mdbn_ctx* c, const float* W, int ldw, const float* vb, const float* h, int ldh, int B, int V,
                     int H, int kind, int noisy, float* pre, float* mean, float* sample, const RngSeg& rs,
                     cudaStream_t st) {
  int splits = pick_splits(c, B, V, H);
  float* part = (float*)ws_get(c, WS_PART, (size_t)splits * B * V * sizeof(float));
  if (!part) return 3;
  MDBN_TRY((launch_sgemm<false, true>(c, h, ldh, W, ldw, part, B, V, H, splits, 1 << 30, st)));
  int act = kind == MDBN_GRBM ? ACT_LINEAR : ACT_SIGMOID;
  int smp = !sample ? SMP_NONE : (kind == MDBN_GRBM ? (noisy ? SMP_GAUSS : SMP_MEAN) : SMP_BERNOULLI);
  return launch_epilogue(c, part, splits, B, V, vb, act, smp, rs, pre, V, mean, V, sample, V, st);
}

// F[b]: one block per row
__global__ void free_energy_kernel(const float* __restrict__ part, int splits, int B, int H,
                                   const float* __restrict__ hb, const float* __restrict__ v, long long ldv, int V,
                                   const float* __restrict__ vb, int kind, float* __restrict__ F) {
  __shared__ float red[32];
  int b = blockIdx.x;
  float hid = 0.f, vis = 0.f;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[((size_t)z * B + b) * H + j];
    hid += softplusf_(s + hb[j]);
  }
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float x = v[b * ldv + i];
    if (kind == MDBN_GRBM) { float d = x - vb[i]; vis += 0.5f * d * d; } else vis += x * vb[i];
  }
  hid = block_sum(hid, red);
  vis = block_sum(vis, red);
  if (threadIdx.x == 0) F[b] = kind == MDBN_GRBM ? -hid + vis : -hid - vis;
}

// (the tcgen05 path produces the same [splits][B][H] partials)
int free_energy_from_parts(mdbn_ctx* c, const float* part, int splits, int B, int H, const float* hb, const float* v,
                           long long ldv, int V, const float* vb, int kind, float* F, cudaStream_t st) {
  free_energy_kernel<<<B, 256, 0, st>>>(part, splits, B, H, hb, v, ldv, V, vb, kind, F);
  c->launches++;
  MDBN_CUDA(cudaGetLastError());
  return 0;
}

int generic_free_energy(mdbn_ctx* c, const float* W, int ldw, const float* hb, const float* vb, const float* v,
                        int ldv, int B, int V, int H, int kind, float* F, cudaStream_t st) {
  int splits = pick_splits(c, B, H, V);
  float* part = (float*)ws_get(c, WS_PART, (size_t)splits * B * H * sizeof(float));
  if (!part) return 3;
  MDBN_TRY((launch_sgemm<false, false>(c, v, ldv, W, ldw, part, B, H, V, splits, 1 << 30, st)));
  free_energy_kernel<<<B, 256, 0, st>>>(part, splits, B, H, hb, v, ldv, V, vb, kind, F);
  c->launches++;
  MDBN_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// CD step pieces
// ---------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ data, long long ld, const int* __restrict__ idx, int B,
                                   int V, float* __restrict__ out, float* __restrict__ xi) {
  long long total = (long long)B * V;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int b = (int)(e / V), i = (int)(e % V);
    long long r = idx ? idx[b] : b;
    float x = data[r * ld + i];
    out[e] = x;
    if (xi) xi[e] = roundf(x);   // half away from zero == Theano tensor.round default (src/rbm.py:428)
  }
}

// raw column sums of (top half - bottom half) of a [2B,N] matrix
__global__ void col_diff_sum_kernel(const float* __restrict__ X, int B, int N, float* __restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float p = 0.f, q = 0.f;
  for (int b = 0; b < B; ++b) { p += X[(size_t)b * N + n]; q += X[(size_t)(B + b) * N + n]; }
  out[n] = p - q;
}

// reconstruction-cost numerators (src/rbm.py:479-480 CE; :697 MSE with sigma of the linear mean)
__global__ void recon_cost_partial_kernel(const float* __restrict__ prev, const float* __restrict__ v0, long long n,
                                          int kind, float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float p = prev[e], t = v0[e];
    if (kind == MDBN_GRBM) { float d = sigmoidf_(p) - t; s += d * d; }
    else s += t * softplusf_(-p) + (1.f - t) * softplusf_(p);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// pseudo-likelihood numerator per row (src/rbm.py:421-447); one block per row
__global__ void pl_row_kernel(const float* __restrict__ prex, int H, const float* __restrict__ xi, int V,
                              const float* __restrict__ W, int ldw, const float* __restrict__ vb,
                              const int* __restrict__ bit_idx, int kind, float* __restrict__ partial) {
  __shared__ float red[32];
  int b = blockIdx.x, idx = *bit_idx;
  float x = xi[(size_t)b * V + idx];
  float d = 1.f - 2.f * x;   // xi_flip - xi at column idx
  float h0 = 0.f, h1 = 0.f;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float p = prex[(size_t)b * H + j];
    h0 += softplusf_(p);
    h1 += softplusf_(p + d * W[(size_t)idx * ldw + j]);
  }
  h0 = block_sum(h0, red);
  h1 = block_sum(h1, red);
  if (threadIdx.x == 0) {
    float vterm;
    if (kind == MDBN_GRBM) {
      float a = x - vb[idx], c = (1.f - x) - vb[idx];
      vterm = 0.5f * (a * a - c * c);          // vis(xi) - vis(xi_flip)
    } else {
      vterm = d * vb[idx];                     // -xi.vb + xi_flip.vb
    }
    float diff = (h1 - h0) + vterm;            // F(xi) - F(xi_flip)
    partial[b] = -(float)V * softplusf_(diff);
  }
}

__global__ void sum_small_kernel(const float* __restrict__ partial, int n, float* __restrict__ out, float rows,
                                 float* __restrict__ rows_out) {
  // single thread, fixed order: n is at most a few hundred
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += partial[i];
    *out = s;
    *rows_out = rows;
  }
}

__global__ void bump_bit_idx_kernel(int* bit_idx, int V) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *bit_idx = (*bit_idx + 1) % V;
}

// W, S update (src/rbm.py:347-365); G holds raw v0^T ph - nv^T nh sums, dense ld = H
__global__ void update_w_kernel(float* __restrict__ W, float* __restrict__ S, const float* __restrict__ Wsnap,
                                int ldw, const float* __restrict__ G, int V, int H, UpdateScalars u) {
  long long total = (long long)V * H;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int i = (int)(e / H), j = (int)(e % H);
    size_t o = (size_t)i * ldw + j;
    float wo, so;      // (one formula for every multi-kernel path: the data-parallel APPLY equals the fused full step)
    update_one(u, G[e], W[o], S[o], Wsnap ? Wsnap[o] : 0.f, Wsnap != nullptr, wo, so);
    S[o] = so;
    W[o] = wo;
  }
}

__global__ void update_bias_kernel(float* __restrict__ b, float* __restrict__ S, const float* __restrict__ gsum,
                                   int n, float inv_rows, float mom, float lr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = gsum[i] * inv_rows, s = S[i];
  S[i] = g + (s - g) * mom;
  b[i] = b[i] + s * lr;
}

__global__ void finalize_cost_kernel(const float* __restrict__ num, float inv_den, float* __restrict__ cost) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *cost = *num * inv_den;
}

int apply_update(mdbn_ctx* c, const mdbn_cd_args& a, const float* G, int rows, cudaStream_t st) {
  const int V = a.V, H = a.H;
  const UpdateScalars u = make_update_scalars(a);
  long long total = (long long)V * H;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 8 * c->num_sms) blocks = 8 * c->num_sms;
  update_w_kernel<<<blocks, 256, 0, st>>>(a.W, a.W_speed, a.weightcost != 0.f ? a.W_snap : nullptr, a.ldw, G, V, H, u);
  c->launches++;
  float inv_rows = 1.0f / (float)rows;
  update_bias_kernel<<<(H + 255) / 256, 256, 0, st>>>(a.hbias, a.hbias_speed, G + total, H, inv_rows, a.momentum, a.lr);
  c->launches++;
  update_bias_kernel<<<(V + 255) / 256, 256, 0, st>>>(a.vbias, a.vbias_speed, G + total + H, V, inv_rows, a.momentum,
                                                      a.lr);
  c->launches++;
  if (a.cost_out) {
    // CE / PL: mean over rows; GRBM MSE: mean over rows*V (src/rbm.py:697)
    float den = (a.persistent == nullptr && a.kind == MDBN_GRBM) ? (float)rows * (float)V : (float)rows;
    finalize_cost_kernel<<<1, 32, 0, st>>>(G + total + H + V, 1.0f / den, a.cost_out);
    c->launches++;
  }
  MDBN_CUDA(cudaGetLastError());
  return 0;
}

int generic_cd_step(mdbn_ctx* c, const mdbn_cd_args& a, cudaStream_t st) {
  const int B = a.B, V = a.V, H = a.H, k = a.k;
  const long long VH = (long long)V * H;
  float* G = a.phase == MDBN_PHASE_FULL ? (float*)ws_get(c, WS_G, (size_t)(VH + H + V + 2) * sizeof(float))
                                        : a.stats_buf;
  MDBN_CHECK(G != nullptr, "cd_step: stats buffer missing");
  if (a.phase == MDBN_PHASE_APPLY) return apply_update(c, a, G, a.B_total, st);

  float* XV = (float*)ws_get(c, WS_XV, (size_t)2 * B * V * sizeof(float));
  float* YH = (float*)ws_get(c, WS_YH, (size_t)2 * B * H * sizeof(float));
  float* HS = (float*)ws_get(c, WS_HS, (size_t)B * H * sizeof(float));
  float* VS = (float*)ws_get(c, WS_VS, (size_t)B * V * sizeof(float));
  float* PREV = (float*)ws_get(c, WS_PREV, (size_t)B * V * sizeof(float));
  float* RED = (float*)ws_get(c, WS_RED, (size_t)(B > 4096 ? B : 4096) * sizeof(float));   // pl_row_kernel writes one partial per row
  if (!XV || !YH || !HS || !VS || !PREV || !RED) return 3;
  const bool pcd = a.persistent != nullptr;
  float *XI = nullptr, *PREX = nullptr;
  if (pcd) {
    XI = (float*)ws_get(c, WS_XI, (size_t)B * V * sizeof(float));
    PREX = (float*)ws_get(c, WS_PREX, (size_t)B * H * sizeof(float));
    if (!XI || !PREX) return 3;
  }
  ULayout ul = u_layout(a.kind, a.noisy, B, V, H);

  {  // v0 = data[indices]
    long long total = (long long)B * V;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4 * c->num_sms) blocks = 4 * c->num_sms;
    gather_rows_kernel<<<blocks, 256, 0, st>>>(a.data, a.ld_data, a.indices, B, V, XV, XI);
    c->launches++;
  }
  // positive phase (src/rbm.py:303)
  MDBN_TRY(generic_propup(c, a.W, a.ldw, a.hbias, XV, V, B, V, H, nullptr, YH, pcd ? nullptr : HS,
                          make_seg(a.rng, ul.off_h0, 0), st));
  if (pcd) {
    RngSeg none = make_seg(a.rng, 0, 0);
    MDBN_TRY(generic_propup(c, a.W, a.ldw, a.hbias, XI, V, B, V, H, PREX, nullptr, nullptr, none, st));
  }
  const float* h_in = pcd ? a.persistent : HS;   // :308-311
  float* nv_mean = XV + (size_t)B * V;
  float* nh_mean = YH + (size_t)B * H;
  for (int s = 0; s < k; ++s) {                  // :328-336
    long long base = (long long)B * H + s * ul.step_stride;
    MDBN_TRY(generic_propdown(c, a.W, a.ldw, a.vbias, h_in, H, B, V, H, a.kind, a.noisy, PREV, nv_mean,
                              a.kind == MDBN_RBM ? VS : nullptr, make_seg(a.rng, base + ul.off_v, ord_v(s)), st));
    const float* v_in = a.kind == MDBN_GRBM ? nv_mean : VS;   // GRBM: h given v_MEAN (:669)
    MDBN_TRY(generic_propup(c, a.W, a.ldw, a.hbias, v_in, V, B, V, H, nullptr, nh_mean, HS,
                            make_seg(a.rng, base + ul.off_h, ord_h(s)), st));
    h_in = HS;
  }
  // statistics (:411-417): G = [v0;nv]^T (+/-) [ph;nh]
  MDBN_TRY((launch_sgemm<true, false>(c, XV, V, YH, H, G, V, H, 2 * B, 1, B, st)));
  if (c->ev_stats_w) { MDBN_CUDA(cudaEventRecord(c->ev_stats_w, st)); c->ev_stats_w_done = true; }
  col_diff_sum_kernel<<<(H + 255) / 256, 256, 0, st>>>(YH, B, H, G + VH);
  c->launches++;
  col_diff_sum_kernel<<<(V + 255) / 256, 256, 0, st>>>(XV, B, V, G + VH + H);
  c->launches++;
  // monitoring cost numerator (old parameters)
  if (pcd) {
    pl_row_kernel<<<B, 128, 0, st>>>(PREX, H, XI, V, a.W, a.ldw, a.vbias, a.bit_i_idx, a.kind, RED);
    c->launches++;
    sum_small_kernel<<<1, 32, 0, st>>>(RED, B, G + VH + H + V, (float)B, G + VH + H + V + 1);
    c->launches++;
    bump_bit_idx_kernel<<<1, 32, 0, st>>>(a.bit_i_idx, V);
    c->launches++;
    MDBN_CUDA(cudaMemcpyAsync(a.persistent, HS, (size_t)B * H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    int nb = 256;
    recon_cost_partial_kernel<<<nb, 256, 0, st>>>(PREV, XV, (long long)B * V, a.kind, RED);
    c->launches++;
    sum_small_kernel<<<1, 32, 0, st>>>(RED, nb, G + VH + H + V, (float)B, G + VH + H + V + 1);
    c->launches++;
  }
  MDBN_CUDA(cudaGetLastError());
  if (a.phase == MDBN_PHASE_STATS) return 0;
  return apply_update(c, a, G, B, st);
}

}  // namespace mdbn
