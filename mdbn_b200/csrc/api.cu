// C ABI (include/mdbn_b200.h): context management, argument validation, path dispatch.
#include <stdlib.h>
#include <string.h>
#include "ctx.h"

namespace mdbn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void* ws_get(mdbn_ctx* c, int slot, size_t bytes) {
  mdbn_ctx::Buf& b = c->ws[slot];
  if (bytes <= b.n && b.p) return b.p;
  // Growing frees the old block: make sure nothing enqueued still uses it.
  if (b.p) {
    cudaDeviceSynchronize();
    cudaFree(b.p);
    b.p = nullptr;
    b.n = 0;
  }
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    set_error("scratch slot %d: cudaMalloc(%zu) -> %s", slot, want, cudaGetErrorString(e));
    b.p = nullptr;
    return nullptr;
  }
  b.n = want;
  return b.p;
}

}  // namespace mdbn

namespace mdbn {
void ctx_enter(mdbn_ctx* c, cudaStream_t st) {
  if (c->has_last_stream && c->last_stream != st) {
    if (!c->ev_order && cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming) != cudaSuccess) c->ev_order = nullptr;
    // (a previous stream the caller has destroyed since has nothing left to wait for: errors are dropped)
    if (c->ev_order && cudaEventRecord(c->ev_order, c->last_stream) == cudaSuccess) cudaStreamWaitEvent(st, c->ev_order, 0);
    cudaGetLastError();
  }
  c->last_stream = st;
  c->has_last_stream = true;
}
}  // namespace mdbn

using namespace mdbn;

extern "C" {

int mdbn_abi_version(void) { return MDBN_ABI_VERSION; }

const char* mdbn_last_error(void) { return g_err; }

int mdbn_create(mdbn_ctx** out, int device) {
  MDBN_CHECK(out != nullptr, "mdbn_create: out is NULL");
  int n = 0;
  MDBN_CUDA(cudaGetDeviceCount(&n));
  MDBN_CHECK(device >= 0 && device < n, "mdbn_create: device %d not present (%d devices)", device, n);
  MDBN_CUDA(cudaSetDevice(device));
  cudaDeviceProp p;
  MDBN_CUDA(cudaGetDeviceProperties(&p, device));
  MDBN_CHECK(p.major == 10, "mdbn_b200 is built for sm_100a only; device %d is sm_%d%d", device, p.major, p.minor);
  mdbn_ctx* c = new mdbn_ctx();
  c->device = device;
  c->num_sms = p.multiProcessorCount;
  c->l2_bytes = p.l2CacheSize;
  MDBN_CUDA(cudaMalloc(&c->barrier, 4096));
  MDBN_CUDA(cudaMemset(c->barrier, 0, 4096));
  *out = c;
  return 0;
}

int mdbn_destroy(mdbn_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& b : c->ws)
    if (b.p) cudaFree(b.p);
  if (c->barrier) cudaFree(c->barrier);
  if (c->ev_order) cudaEventDestroy(c->ev_order);
  delete c;
  return 0;
}

unsigned long long mdbn_launch_count(const mdbn_ctx* c) { return c ? c->launches : 0; }

int mdbn_copy_async(void* dst, const void* src, unsigned long long bytes, void* stream) {
  MDBN_CHECK(dst && src, "copy_async: NULL pointer");
  MDBN_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

int mdbn_set_tf32_phases(mdbn_ctx* c, int enable) {
  MDBN_CHECK(c != nullptr, "ctx is NULL");
  c->tf32_phases = enable != 0;
  return 0;
}

long long mdbn_stats_size(int V, int H) { return (long long)V * H + H + V + 2; }

// The single-phase calls run on the tensor cores (fp32-exact split-TF32 by default, plain TF32 after
// mdbn_set_tf32_phases) whenever the operands suit TMA; layers of a few thousand weights stay on the SIMT kernels
// (one small launch either way).
static bool tensor_phase_wanted(int V, int H) { return (long long)V * H >= 4096; }

// AUTO takes the broadcast kernel (mid.cu) where the whole visible state of a minibatch fits in shared memory and the
// layer is small enough for its column-sliced reads of W to stay in L2; wider layers stream row slabs (skinny.cu)
static bool mid_wanted(const mdbn_ctx* c, const mdbn_cd_args& a) {
  static const bool off = getenv("MDBN_NO_MID") != nullptr;
  return !off && (long long)a.V * a.ldw <= (4LL << 20) && mid_supported(c, a);
}

static int check_common(const mdbn_ctx* c, const void* W, int ldw, int B, int V, int H) {
  MDBN_CHECK(c != nullptr, "ctx is NULL");
  MDBN_CHECK(W != nullptr, "W is NULL");
  MDBN_CHECK(B > 0 && V > 0 && H > 0, "bad shape B=%d V=%d H=%d", B, V, H);
  MDBN_CHECK(ldw >= H, "ldw=%d < H=%d", ldw, H);
  return 0;
}

int mdbn_propup(mdbn_ctx* c, const float* W, int ldw, const float* hbias, const float* v, int ldv, int B, int V,
                int H, float* pre_out, float* mean_out, float* sample_out, const mdbn_rng* rng, void* stream) {
  MDBN_TRY(check_common(c, W, ldw, B, V, H));
  MDBN_CHECK(hbias && v && ldv >= V, "propup: bad hbias/v/ldv");
  MDBN_CHECK(!sample_out || (rng && rng->mode != MDBN_RNG_NONE), "propup: sample_out needs an rng");
  MDBN_CHECK(!(rng && rng->mode == MDBN_RNG_BUFFER && sample_out) || rng->buffer, "propup: rng buffer is NULL");
  MDBN_CUDA(cudaSetDevice(c->device));
  ctx_enter(c, (cudaStream_t)stream);
  mdbn_rng none = {MDBN_RNG_NONE, nullptr, 0, 0};
  if (tensor_phase_wanted(V, H) && tensor_phase_supported(W, ldw, v, ldv))
    return tensor_propup(c, W, ldw, hbias, v, ldv, B, V, H, pre_out, mean_out, sample_out,
                         make_seg(rng ? *rng : none, 0, 0), (cudaStream_t)stream);
  return generic_propup(c, W, ldw, hbias, v, ldv, B, V, H, pre_out, mean_out, sample_out,
                        make_seg(rng ? *rng : none, 0, 0), (cudaStream_t)stream);
}

int mdbn_forward(mdbn_ctx* c, const float* W, int ldw, const float* b, const float* x, int ldx, int B, int V, int H,
                 float* out, void* stream) {
  MDBN_CHECK(out != nullptr, "forward: out is NULL");
  return mdbn_propup(c, W, ldw, b, x, ldx, B, V, H, nullptr, out, nullptr, nullptr, stream);
}

int mdbn_propdown(mdbn_ctx* c, const float* W, int ldw, const float* vbias, const float* h, int ldh, int B, int V,
                  int H, int kind, int noisy, float* pre_out, float* mean_out, float* sample_out,
                  const mdbn_rng* rng, void* stream) {
  MDBN_TRY(check_common(c, W, ldw, B, V, H));
  MDBN_CHECK(vbias && h && ldh >= H, "propdown: bad vbias/h/ldh");
  MDBN_CHECK(kind == MDBN_RBM || kind == MDBN_GRBM, "propdown: bad kind %d", kind);
  bool needs_rng = sample_out && (kind == MDBN_RBM || noisy);
  MDBN_CHECK(!needs_rng || (rng && rng->mode != MDBN_RNG_NONE), "propdown: sample_out needs an rng");
  MDBN_CHECK(!(needs_rng && rng->mode == MDBN_RNG_BUFFER) || rng->buffer, "propdown: rng buffer is NULL");
  MDBN_CUDA(cudaSetDevice(c->device));
  ctx_enter(c, (cudaStream_t)stream);
  mdbn_rng none = {MDBN_RNG_NONE, nullptr, 0, 0};
  if (tensor_phase_wanted(V, H) && tensor_phase_supported(W, ldw, h, ldh))
    return tensor_propdown(c, W, ldw, vbias, h, ldh, B, V, H, kind, noisy, pre_out, mean_out, sample_out,
                           make_seg(rng ? *rng : none, 0, 1), (cudaStream_t)stream);
  return generic_propdown(c, W, ldw, vbias, h, ldh, B, V, H, kind, noisy, pre_out, mean_out, sample_out,
                          make_seg(rng ? *rng : none, 0, 1), (cudaStream_t)stream);
}

int mdbn_free_energy(mdbn_ctx* c, const float* W, int ldw, const float* hbias, const float* vbias, const float* v,
                     int ldv, int B, int V, int H, int kind, float* F_out, void* stream) {
  MDBN_TRY(check_common(c, W, ldw, B, V, H));
  MDBN_CHECK(hbias && vbias && v && F_out && ldv >= V, "free_energy: bad arguments");
  MDBN_CUDA(cudaSetDevice(c->device));
  ctx_enter(c, (cudaStream_t)stream);
  if (tensor_phase_wanted(V, H) && tensor_phase_supported(W, ldw, v, ldv))
    return tensor_free_energy(c, W, ldw, hbias, vbias, v, ldv, B, V, H, kind, F_out, (cudaStream_t)stream);
  return generic_free_energy(c, W, ldw, hbias, vbias, v, ldv, B, V, H, kind, F_out, (cudaStream_t)stream);
}

int mdbn_cd_step(mdbn_ctx* c, const mdbn_cd_args* a, void* stream) {
  MDBN_CHECK(a != nullptr, "cd_step: args is NULL");
  MDBN_TRY(check_common(c, a->W, a->ldw, a->B, a->V, a->H));
  MDBN_CHECK(a->kind == MDBN_RBM || a->kind == MDBN_GRBM, "cd_step: bad kind %d", a->kind);
  MDBN_CHECK(a->hbias && a->vbias && a->W_speed && a->hbias_speed && a->vbias_speed, "cd_step: NULL parameter/state");
  MDBN_CHECK(a->B_nom > 0, "cd_step: batch_size (B_nom) must be given (src/rbm.py:413)");
  MDBN_CHECK(a->phase >= MDBN_PHASE_FULL && a->phase <= MDBN_PHASE_APPLY, "cd_step: bad phase");
  MDBN_CHECK(a->weightcost == 0.f || a->W_snap, "cd_step: weightcost != 0 needs W_snap");
  if (a->phase != MDBN_PHASE_APPLY) {
    MDBN_CHECK(a->k >= 1, "cd_step: k must be >= 1");
    MDBN_CHECK(a->data && a->ld_data >= a->V, "cd_step: bad data/ld_data");
    MDBN_CHECK(a->rng.mode == MDBN_RNG_BUFFER || a->rng.mode == MDBN_RNG_PHILOX, "cd_step: rng mode required");
    MDBN_CHECK(a->rng.mode != MDBN_RNG_BUFFER || a->rng.buffer, "cd_step: rng buffer is NULL");
    MDBN_CHECK(!a->persistent || a->bit_i_idx, "cd_step: PCD needs bit_i_idx");
  }
  if (a->phase != MDBN_PHASE_FULL) MDBN_CHECK(a->stats_buf, "cd_step: STATS/APPLY need stats_buf");
  if (a->phase == MDBN_PHASE_APPLY) MDBN_CHECK(a->B_total > 0, "cd_step: APPLY needs B_total");
  MDBN_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  ctx_enter(c, st);
  if (a->comm) {
    MDBN_CHECK(a->phase == MDBN_PHASE_FULL, "cd_step: comm goes with full steps only");
    return comm_cd_step(c, *a, st);
  }
  int path = a->path;
  if (path == MDBN_PATH_AUTO) {
    if (a->phase == MDBN_PHASE_FULL && tiny_supported(c, *a)) path = MDBN_PATH_TINY;
    else if (a->phase == MDBN_PHASE_FULL && mid_wanted(c, *a)) path = MDBN_PATH_MID;
    else if (a->phase == MDBN_PHASE_FULL && skinny_supported(c, *a)) path = MDBN_PATH_SKINNY;
    else if ((a->B > 20 || a->tf32) && tensor_supported(c, *a)) path = MDBN_PATH_TENSOR;   // fp32-exact unless tf32
    else path = MDBN_PATH_GENERIC;
  }
  switch (path) {
    case MDBN_PATH_GENERIC:
      return generic_cd_step(c, *a, st);
    case MDBN_PATH_SKINNY:
      MDBN_CHECK(skinny_supported(c, *a), "cd_step: skinny path does not take B=%d V=%d H=%d ldw=%d phase=%d", a->B,
                 a->V, a->H, a->ldw, a->phase);
      return skinny_cd_step(c, *a, st);
    case MDBN_PATH_TINY:
      MDBN_CHECK(tiny_supported(c, *a), "cd_step: tiny path does not take B=%d V=%d H=%d ldw=%d phase=%d", a->B, a->V, a->H,
                 a->ldw, a->phase);
      return tiny_cd_steps(c, *a, 1, st);
    case MDBN_PATH_MID:
      MDBN_CHECK(mid_supported(c, *a), "cd_step: mid path does not take B=%d V=%d H=%d ldw=%d phase=%d", a->B, a->V, a->H,
                 a->ldw, a->phase);
      return mid_cd_steps(c, *a, 1, st);
    case MDBN_PATH_TENSOR:
      MDBN_CHECK(tensor_supported(c, *a), "cd_step: tensor path does not take B=%d V=%d H=%d ldw=%d", a->B, a->V, a->H,
                 a->ldw);
      return tensor_cd_step(c, *a, st);
    default:
      set_error("cd_step: bad path %d", path);
      return 2;
  }
}

int mdbn_cd_steps(mdbn_ctx* c, const mdbn_cd_args* a, int n_steps, void* stream) {
  MDBN_CHECK(a != nullptr, "cd_steps: args is NULL");
  MDBN_CHECK(n_steps >= 1, "cd_steps: n_steps must be >= 1");
  if (n_steps == 1) return mdbn_cd_step(c, a, stream);
  MDBN_CHECK(a->phase == MDBN_PHASE_FULL, "cd_steps: only full steps can be chained");
  MDBN_CHECK(a->rng.mode == MDBN_RNG_PHILOX, "cd_steps: the PHILOX generator is required (one offset per step)");
  MDBN_CHECK(a->indices, "cd_steps: indices [n_steps][B] is required");
  // one launch when the persistent kernel takes the shape (validation happens in the first single step otherwise)
  const bool want_tiny = a->path == MDBN_PATH_AUTO || a->path == MDBN_PATH_TINY;
  const bool want_skinny = a->path == MDBN_PATH_AUTO || a->path == MDBN_PATH_SKINNY;
  const bool use_tiny = want_tiny && c && a->W && tiny_supported(c, *a);
  const bool use_mid = !use_tiny && c && a->W && ((a->path == MDBN_PATH_AUTO && mid_wanted(c, *a)) ||
                                                  (a->path == MDBN_PATH_MID && mid_supported(c, *a)));
  if (use_tiny || use_mid || (want_skinny && c && a->W && skinny_supported(c, *a))) {
    MDBN_TRY(check_common(c, a->W, a->ldw, a->B, a->V, a->H));
    MDBN_CHECK(a->kind == MDBN_RBM || a->kind == MDBN_GRBM, "cd_steps: bad kind %d", a->kind);
    MDBN_CHECK(a->hbias && a->vbias && a->W_speed && a->hbias_speed && a->vbias_speed, "cd_steps: NULL parameter/state");
    MDBN_CHECK(a->B_nom > 0, "cd_steps: batch_size (B_nom) must be given (src/rbm.py:413)");
    MDBN_CHECK(a->weightcost == 0.f || a->W_snap, "cd_steps: weightcost != 0 needs W_snap");
    MDBN_CHECK(a->k >= 1, "cd_steps: k must be >= 1");
    MDBN_CHECK(a->data && a->ld_data >= a->V, "cd_steps: bad data/ld_data");
    MDBN_CHECK(!a->persistent || a->bit_i_idx, "cd_steps: PCD needs bit_i_idx");
    MDBN_CUDA(cudaSetDevice(c->device));
    ctx_enter(c, (cudaStream_t)stream);
    if (use_tiny) return tiny_cd_steps(c, *a, n_steps, (cudaStream_t)stream);
    if (use_mid) return mid_cd_steps(c, *a, n_steps, (cudaStream_t)stream);
    return skinny_cd_steps(c, *a, n_steps, (cudaStream_t)stream);
  }
  for (int s = 0; s < n_steps; ++s) {
    mdbn_cd_args one = *a;
    one.indices = a->indices + (size_t)s * a->B;
    one.cost_out = a->cost_out ? a->cost_out + s : nullptr;
    one.rng.offset = a->rng.offset + (unsigned long long)s;
    MDBN_TRY(mdbn_cd_step(c, &one, stream));
  }
  return 0;
}

}  // extern "C"
