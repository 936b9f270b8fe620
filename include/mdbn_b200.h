/* mdbn_b200 — C ABI of the B200-native RBM / GRBM contrastive-divergence path.
 *
 * The reference (glgerard/MDBN) has no FFI: its hot path is Python methods that
 * build Theano graphs.  Each entry point below replaces the Theano-compiled
 * function behind one of those methods; the citation after "replaces:" is the
 * reference interface (path:line relative to the reference root).
 *
 * Conventions
 *   - plain C, no exceptions; every call returns 0 on success, non-zero on error,
 *     message via mdbn_last_error() (thread-local).
 *   - all tensor arguments are CALLER-OWNED DEVICE pointers, fp32, row-major;
 *     `ld*` are row strides in elements.  W is [V,H] (src/rbm.py:100-109).
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, not waited on.
 *   - the context owns only scratch memory (arenas, accumulators, grid-barrier words).  Calls on ONE context from
 *     different streams are serialised by the library (a call on another stream than the previous one waits for it with
 *     an event); users that want concurrency between streams create one context per stream.  Not thread-safe: one host
 *     thread at a time per context.
 *   - samples are fp32 0.0/1.0 (src/rbm.py:210-212).
 */
#ifndef MDBN_B200_H
#define MDBN_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define MDBN_ABI_VERSION 2

typedef struct mdbn_ctx mdbn_ctx;
typedef struct mdbn_comm mdbn_comm;   /* NCCL communicator of the data-parallel step (below) */

enum { MDBN_RBM = 0, MDBN_GRBM = 1 };                       /* src/rbm.py:46 / :631 */
enum { MDBN_RNG_NONE = 0, MDBN_RNG_BUFFER = 1, MDBN_RNG_PHILOX = 2 };
/* AUTO picks by shape (full steps): TINY (one cluster, weights in shared memory: small layers), MID (medium layers,
 * B <= 20: column / row sliced propagations around broadcasts, W in L2), SKINNY (persistent row-slab grid kernel,
 * B <= 20, wide layers), TENSOR (tcgen05 GEMMs with fused epilogues, any B > 20: fp32-exact split-TF32 arithmetic, or
 * plain TF32 with tf32 = 1), else GENERIC (fp32 SIMT: operands that do not suit TMA). */
enum { MDBN_PATH_AUTO = 0, MDBN_PATH_GENERIC = 1, MDBN_PATH_SKINNY = 2, MDBN_PATH_TENSOR = 3, MDBN_PATH_TINY = 4, MDBN_PATH_MID = 5 };
enum { MDBN_PHASE_FULL = 0, MDBN_PHASE_STATS = 1, MDBN_PHASE_APPLY = 2 };

/* Source of randomness for one call.
 * BUFFER: `buffer` points at this call's fp32 values, laid out as SURVEY.md App. A
 *         ([U_h0 B*H] then per Gibbs step [U_v B*V (RBM) | N_v B*V (noisy GRBM)] [U_h B*H]);
 *         for the single-phase calls it is just the B*H or B*V values of that phase.
 *         Values are uniforms in (0,1) except N_v (standard normals).  Comparison is u < p.
 * PHILOX: Philox4x32-10, key = seed, counter = (element/4, segment ordinal, offset lo, offset hi);
 *         the caller advances `offset` by one per call. */
typedef struct {
  int mode;
  const float* buffer;
  unsigned long long seed;
  unsigned long long offset;
} mdbn_rng;

int mdbn_abi_version(void);
const char* mdbn_last_error(void);
int mdbn_create(mdbn_ctx** out, int device);
int mdbn_destroy(mdbn_ctx* ctx);
/* number of kernels this context has launched (diagnostic; bench.py's gpu_launches) */
unsigned long long mdbn_launch_count(const mdbn_ctx* ctx);
/* The single-phase calls below (propup / propdown / free_energy / forward) run on the tcgen05 tensor-core path
 * when their operands are 16-byte aligned (row strides multiples of 4) and the layer has >= 4096 weights:
 * fp32-exact split-TF32 arithmetic by default (tolerance 1e-5); enable != 0 selects plain TF32 (tolerance 2e-3,
 * one MMA per k-step instead of three). */
int mdbn_set_tf32_phases(mdbn_ctx* ctx, int enable);

/* pre = v W + hbias ; mean = sigmoid(pre) ; sample = (u < mean).  Any of the three
 * outputs may be NULL (sample requires rng->mode != NONE).
 * replaces: RBM.propup src/rbm.py:187-199, RBM.sample_h_given_v src/rbm.py:201-213,
 *           HiddenLayer output src/mlp.py:103-107 (mean only). */
int mdbn_propup(mdbn_ctx* ctx, const float* W, int ldw, const float* hbias,
                const float* v, int ldv, int B, int V, int H,
                float* pre_out, float* mean_out, float* sample_out,
                const mdbn_rng* rng, void* stream);

/* out = sigmoid(x W + b): the deterministic up-pass of one DBN layer (mean only, no sampling).
 * replaces: HiddenLayer.output src/mlp.py:103-107 as used by DBN.get_output src/dbn.py:214-236 and by the input of the
 *           upper layers during pretraining (src/dbn.py:146). */
int mdbn_forward(mdbn_ctx* ctx, const float* W, int ldw, const float* b, const float* x, int ldx, int B, int V, int H,
                 float* out, void* stream);

/* pre = h W^T + vbias.
 * kind RBM : mean = sigmoid(pre), sample = (u < mean)              src/rbm.py:215-240
 * kind GRBM: mean = pre (linear, unit variance); sample = mean, or mean + n with
 *            n ~ N(0,1) when noisy != 0 (error_free=False)          src/rbm.py:647-660 */
int mdbn_propdown(mdbn_ctx* ctx, const float* W, int ldw, const float* vbias,
                  const float* h, int ldh, int B, int V, int H, int kind, int noisy,
                  float* pre_out, float* mean_out, float* sample_out,
                  const mdbn_rng* rng, void* stream);

/* F[b] = -sum_j softplus(v W + hbias)_j - v.vbias            (RBM,  src/rbm.py:166-171)
 *      = -sum_j softplus(v W + hbias)_j + 0.5 sum_i (v-vbias)^2 (GRBM, src/rbm.py:684-688) */
int mdbn_free_energy(mdbn_ctx* ctx, const float* W, int ldw, const float* hbias, const float* vbias,
                     const float* v, int ldv, int B, int V, int H, int kind,
                     float* F_out, void* stream);

/* One CD-k / PCD-k parameter update — the Theano function compiled from
 * RBM.get_cost_updates (src/rbm.py:258-376) incl. compute_rbm_grad (:392-419), the
 * lambda_1/lambda_2/momentum update (:347-365) and the monitoring cost
 * (:421-482, :690-699). */
typedef struct {
  int kind;              /* MDBN_RBM | MDBN_GRBM */
  int noisy;             /* GRBM: error_free == False                                  */
  int B;                 /* rows in this minibatch                                      */
  int B_nom;             /* the batch_size argument: divisor of the W statistics (:413) */
  int V, H, k;
  float* W;              /* [V,ldw]  updated in place */
  int ldw;
  float* hbias;          /* [H] */
  float* vbias;          /* [V] */
  float* W_speed;        /* [V,ldw]  src/rbm.py:153-162 */
  float* hbias_speed;
  float* vbias_speed;
  const float* W_snap;   /* [V,ldw] frozen copy multiplied by weightcost (:414-415); NULL iff weightcost==0 */
  const float* data;     /* dataset [N,ld_data]; v0 = data[indices] (givens, src/dbn.py:307) */
  long long ld_data;
  const int* indices;    /* device int32 [B], or NULL for rows 0..B-1 */
  float* persistent;     /* PCD chain state [B,H] (ld = H) updated in place (:369), or NULL for CD */
  int* bit_i_idx;        /* device int: pseudo-likelihood column cursor (:425,445); required iff persistent */
  float lr, momentum, lambda_1, lambda_2, weightcost;
  mdbn_rng rng;
  float* cost_out;       /* device float: reconstruction cost (CD) / pseudo-likelihood (PCD) */
  int path;              /* MDBN_PATH_* (AUTO picks by shape) */
  int tf32;              /* TENSOR path arithmetic: 0 = fp32-exact split TF32 (three MMAs per k-step, tolerance 1e-5),
                            1 = plain TF32 (tolerance 2e-3) */
  int phase;             /* MDBN_PHASE_FULL, or STATS (fill stats_buf, no update) / APPLY (update from stats_buf) */
  float* stats_buf;      /* data-parallel packing: [sum v0^T ph - nv^T nh (V*H) | sum(ph-nh) (H) | sum(v0-nv) (V) |
                            cost numerator (1) | rows (1)], raw sums over this rank's rows */
  int B_total;           /* APPLY, or FULL with comm: rows of the whole minibatch over all ranks (bias means, cost mean) */
  mdbn_comm* comm;       /* FULL only; non-NULL: data-parallel step.  B / indices / persistent are THIS rank's rows and chains,
                            B_nom the batch_size argument, B_total the rows of the whole minibatch: the library runs the
                            statistics on the shard, all-reduces the packed buffer over NCCL (two chunks on a side stream,
                            the V*H block overlapping the tail kernels) and applies the identical update on every rank.
                            No counterpart in the reference (single device); the sums it shards are src/rbm.py:411-417. */
} mdbn_cd_args;

/* Staging helper of the host-streaming train step (TrainFn.step_from_host): cudaMemcpyAsync(cudaMemcpyDefault)
 * on `stream`; host buffers should be pinned. */
int mdbn_copy_async(void* dst, const void* src, unsigned long long bytes, void* stream);

int mdbn_cd_step(mdbn_ctx* ctx, const mdbn_cd_args* args, void* stream);

/* n_steps consecutive full steps (an epoch, or the minibatches up to the next validation point of
 * src/dbn.py:343-353): args->indices is [n_steps][B] row-major, args->cost_out [n_steps] (or NULL); step s
 * draws with rng.offset + s (PHILOX generator only).  Same results as n_steps calls of mdbn_cd_step with
 * those arguments; on the skinny path it is ONE kernel launch, otherwise a loop of single steps. */
int mdbn_cd_steps(mdbn_ctx* ctx, const mdbn_cd_args* args, int n_steps, void* stream);

/* size in floats of stats_buf for a layer */
long long mdbn_stats_size(int V, int H);

/* Data parallelism inside the library (nothing to replace in the reference, which is single device).
 * Rank 0 creates an id (ncclUniqueId, MDBN_COMM_ID_BYTES bytes) and hands it to the other ranks by whatever
 * channel the host has; every rank then calls mdbn_comm_init with its device.  NCCL is loaded at run time. */
#define MDBN_COMM_ID_BYTES 128
int mdbn_comm_unique_id(unsigned char* id_out);
int mdbn_comm_init(mdbn_comm** out, const unsigned char* id, int rank, int world, int device);
int mdbn_comm_destroy(mdbn_comm* comm);
/* in-place sum over the ranks, enqueued on `stream` (used for the small exchange of modality activations) */
int mdbn_comm_all_reduce(mdbn_comm* comm, float* buf, unsigned long long count, void* stream);

#ifdef __cplusplus
}
#endif
#endif
