// Host-side context: scratch memory + launch accounting.  One per (device, user).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string>
#include "common.cuh"

struct mdbn_ctx {
  int device = 0;
  int num_sms = 0;
  int l2_bytes = 0;
  unsigned long long launches = 0;
  // named scratch arenas (grown on demand, never shrunk)
  struct Buf { void* p = nullptr; size_t n = 0; };
  Buf ws[16];
  unsigned int* barrier = nullptr;   // grid barrier word for the persistent kernel
  int tf32_phases = 0;               // single-phase calls use the tcgen05 TF32 path (mdbn_set_tf32_phases)
  unsigned skinny_parity = 0;        // which of the two accumulator sets the next skinny launch uses
  unsigned long long skinny_key = 0; // layout of the accumulator scratch (re-zeroed when it changes)
  // data-parallel step: event the statistics path records once the V*H block of the packed buffer is complete
  cudaEvent_t ev_stats_w = nullptr;
  bool ev_stats_w_done = false;
  // The scratch arenas, the accumulators and the grid-barrier words are shared by every call on this context: a call on
  // another stream than the previous one first waits for the previous call (ctx_enter), so that users on different
  // streams are serialised instead of racing
  cudaStream_t last_stream = nullptr;
  bool has_last_stream = false;
  cudaEvent_t ev_order = nullptr;
};

namespace mdbn {

void set_error(const char* fmt, ...);

#define MDBN_CUDA(x)                                                                   \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      mdbn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

#define MDBN_CHECK(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      mdbn::set_error(__VA_ARGS__);    \
      return 2;                        \
    }                                  \
  } while (0)

#define MDBN_TRY(x)        \
  do {                     \
    int r_ = (x);          \
    if (r_) return r_;     \
  } while (0)

enum WsSlot {
  WS_PART = 0,   // split-K partials
  WS_XV,         // [2B,V]  v0 ; nv_mean
  WS_YH,         // [2B,H]  ph_mean ; nh_mean
  WS_HS,         // [B,H]   hidden sample (chain state)
  WS_VS,         // [B,V]   visible input of the next propup
  WS_PREV,       // [B,V]   last pre-sigmoid visible
  WS_G,          // [V*H + H + V + 2] packed statistics
  WS_RED,        // small reduction partials
  WS_XI,         // [B,V]   rounded input (pseudo-likelihood)
  WS_PREX,       // [B,H]   pre-activation of rounded input
  WS_SKINNY,     // persistent-kernel scratch
  WS_TENSOR,     // tcgen05 path scratch
  WS_MISC,
};

// orders this call behind the previous call on the context when the stream changed (see mdbn_ctx::last_stream)
void ctx_enter(mdbn_ctx* c, cudaStream_t st);

// grow-only scratch; returns nullptr (and sets error) on failure
void* ws_get(mdbn_ctx* c, int slot, size_t bytes);

// ---- generic (any shape, fp32-exact) path: generic.cu -----------------------
int generic_propup(mdbn_ctx*, const float* W, int ldw, const float* hb, const float* v, int ldv, int B, int V, int H,
                   float* pre, float* mean, float* sample, const RngSeg& rs, cudaStream_t st);
int generic_propdown(mdbn_ctx*, const float* W, int ldw, const float* vb, const float* h, int ldh, int B, int V, int H,
                     int kind, int noisy, float* pre, float* mean, float* sample, const RngSeg& rs, cudaStream_t st);
int generic_free_energy(mdbn_ctx*, const float* W, int ldw, const float* hb, const float* vb, const float* v, int ldv,
                        int B, int V, int H, int kind, float* F, cudaStream_t st);
int generic_cd_step(mdbn_ctx*, const mdbn_cd_args& a, cudaStream_t st);

// ---- skinny persistent path (B <= 32): skinny.cu ----------------------------
bool skinny_supported(const mdbn_ctx*, const mdbn_cd_args& a);
int skinny_cd_step(mdbn_ctx*, const mdbn_cd_args& a, cudaStream_t st);
int skinny_cd_steps(mdbn_ctx*, const mdbn_cd_args& a, int n_steps, cudaStream_t st);

// ---- small layers on one thread-block cluster, weights resident in shared memory: tiny.cu ----
bool tiny_supported(const mdbn_ctx*, const mdbn_cd_args& a);
int tiny_cd_steps(mdbn_ctx*, const mdbn_cd_args& a, int n_steps, cudaStream_t st);

// ---- medium layers at skinny batch: column / row sliced propagations around broadcasts, W in L2: mid.cu ----
bool mid_supported(const mdbn_ctx*, const mdbn_cd_args& a);
int mid_cd_steps(mdbn_ctx*, const mdbn_cd_args& a, int n_steps, cudaStream_t st);

// ---- tcgen05 / TMA path (large batch, TF32): tensor.cu ----------------------
bool tensor_supported(const mdbn_ctx*, const mdbn_cd_args& a);
int tensor_cd_step(mdbn_ctx*, const mdbn_cd_args& a, cudaStream_t st);
bool tensor_phase_supported(const void* W, int ldw, const void* x, long long ldx);
int tensor_propup(mdbn_ctx*, const float* W, int ldw, const float* hb, const float* v, int ldv, int B, int V, int H,
                  float* pre, float* mean, float* sample, const RngSeg& rs, cudaStream_t st);
int tensor_propdown(mdbn_ctx*, const float* W, int ldw, const float* vb, const float* h, int ldh, int B, int V, int H,
                    int kind, int noisy, float* pre, float* mean, float* sample, const RngSeg& rs, cudaStream_t st);
int tensor_free_energy(mdbn_ctx*, const float* W, int ldw, const float* hb, const float* vb, const float* v, int ldv,
                       int B, int V, int H, int kind, float* F, cudaStream_t st);
int apply_update(mdbn_ctx* c, const mdbn_cd_args& a, const float* G, int rows, cudaStream_t st);

// ---- data-parallel step over NCCL: comm.cu ----
int comm_cd_step(mdbn_ctx* c, const mdbn_cd_args& a, cudaStream_t st);

// hyper-parameters of the W / W_speed update (src/rbm.py:347-365) as the kernels take them
struct UpdateScalars {
  float inv_bnom, wc, c1 /*2*lr*lambda_1*/, decay /*1-2*lr*lambda_2*/, mom, lr;
};
inline UpdateScalars make_update_scalars(const mdbn_cd_args& a) {
  UpdateScalars u;
  u.inv_bnom = 1.0f / (float)a.B_nom;
  u.wc = a.weightcost;
  u.c1 = (2.0f * a.lr) * a.lambda_1;
  u.decay = 1.0f - (2.0f * a.lr) * a.lambda_2;
  u.mom = a.momentum;
  u.lr = a.lr;
  return u;
}

// W, W_speed update of one element from its raw statistic g = (v0^T ph - nv^T nh)[i][j]   src/rbm.py:347-365, :411-415
#ifdef __CUDACC__
// (explicit rounding intrinsics: no context-dependent FMA contraction, so the fused epilogue, the reduction kernel and the
//  data-parallel APPLY kernel produce the same bits from the same statistics)
__device__ __forceinline__ void update_one(const UpdateScalars& u, float graw, float w, float s, float snap, bool has_snap,
                                           float& w_out, float& s_out) {
  float g = __fmul_rn(graw, u.inv_bnom);
  if (has_snap) g = __fmaf_rn(-u.wc, snap, g);
  float mult = u.decay;
  if (u.c1 != 0.f) {
    // 1/D, D = 1 + 2 lr lambda_1 / (|W| + eps)  ==  (|W| + eps) / (|W| + eps + 2 lr lambda_1): ONE MUFU reciprocal
    // (~1 ulp) instead of three IEEE divisions (they were 30 % of the fused kernel's instructions)      src/rbm.py:347-356
    const float t = __fadd_rn(fabsf(w), 0.001f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(t, u.c1)));
    const float invD = __fmul_rn(t, r);
    g = __fmul_rn(g, invD);
    mult = __fmul_rn(mult, invD);
  }
  s_out = __fmaf_rn(__fsub_rn(s, g), u.mom, g);
  w_out = __fmaf_rn(w, mult, __fmul_rn(s, u.lr));     // OLD speed: Theano updates are simultaneous (App. C-1)
}
#endif

// layout of the App. A random buffer
struct ULayout {
  long long off_h0;
  long long step_stride;   // floats per Gibbs step
  long long off_v;         // within a step (valid if has_v)
  long long off_h;         // within a step
  bool has_v;
};
inline ULayout u_layout(int kind, int noisy, int B, int V, int H) {
  ULayout u;
  u.off_h0 = 0;
  u.has_v = (kind == MDBN_RBM) || noisy;
  u.off_v = 0;
  u.off_h = u.has_v ? (long long)B * V : 0;
  u.step_stride = u.off_h + (long long)B * H;
  return u;
}
// segment ordinals (Philox counter word 1): 0 = positive phase, then 1+2s (visible), 2+2s (hidden)
inline uint32_t ord_v(int s) { return 1u + 2u * (uint32_t)s; }
inline uint32_t ord_h(int s) { return 2u + 2u * (uint32_t)s; }

}  // namespace mdbn
