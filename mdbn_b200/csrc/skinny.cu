// Host side of the skinny-batch persistent kernel (skinny_kernel.cuh): geometry, scratch, launch.
// The kernel instantiations are compiled in skinny_i10.cu / skinny_i20.cu (parallel build).
#include "skinny_kernel.cuh"

namespace mdbn {
namespace sk {

extern template int launch<10, false>(mdbn_ctx*, const CUtensorMap*, const Params&, const Geometry&, cudaStream_t);
extern template int launch<10, true>(mdbn_ctx*, const CUtensorMap*, const Params&, const Geometry&, cudaStream_t);
extern template int launch<20, false>(mdbn_ctx*, const CUtensorMap*, const Params&, const Geometry&, cudaStream_t);
extern template int launch<20, true>(mdbn_ctx*, const CUtensorMap*, const Params&, const Geometry&, cudaStream_t);

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

static Geometry plan(const mdbn_ctx* c, const mdbn_cd_args& a, bool chain = false) {
  Geometry g{};
  g.ok = false;
  g.BT = a.B <= 10 ? 10 : (a.B <= 20 ? 20 : 0);
  if (!g.BT || a.ldw % 4 != 0 || !get_encode()) return g;
  if (((uintptr_t)a.W | (uintptr_t)a.W_speed | (uintptr_t)a.W_snap) & 15) return g;
  const int BTS = (g.BT + 3) / 4 * 4;
  const int NT = nt_of(g.BT), NWARP = NT / 32;
  g.CQ = a.ldw / 4;
  if (g.CQ > NT) return g;
  if (g.CQ <= 32) { g.GW = 1; while (g.GW < g.CQ) g.GW <<= 1; } else g.GW = (g.CQ + 31) / 32 * 32;
  g.G = NT / g.GW;
  g.grid = c->num_sms;      // (fewer, fatter CTAs on small layers measured SLOWER: the per-CTA O(B*H) flush and rebuild stay)
  if (g.BT * g.CQ > g.grid * NT || g.grid < g.BT || a.ldw > 4 * NT) return g;
  g.rows_per_cta = (a.V + g.grid - 1) / g.grid;
  g.rows_small = g.rows_per_cta;
  g.n_big = g.grid;
  if (a.persistent && a.V >= 8 * g.grid) {
    // PCD: the last B CTAs also compute the pseudo-likelihood monitor (~2.5 us); they own ~14 % fewer rows
    g.n_big = g.grid - a.B;
    g.rows_per_cta = (int)((100LL * a.V + (100LL * g.grid - 14LL * a.B) - 1) / (100LL * g.grid - 14LL * a.B));
    const int rest = a.V - g.n_big * g.rows_per_cta;
    g.rows_small = rest > 0 ? (rest + a.B - 1) / a.B : 0;
    if (g.rows_small > g.rows_per_cta) { g.rows_per_cta = (a.V + g.grid - 1) / g.grid; g.rows_small = g.rows_per_cta; g.n_big = g.grid; }
  }
  g.rows_alloc = (g.rows_per_cta + 7) & ~7;
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
               vb_b = up128((size_t)g.rows_alloc * 4), hb_b = up128((size_t)g.ldh * 4);
  // fused chains (B <= 10) also hold v0 and round(v0) of the next minibatch
  const size_t dslab_b = (chain && g.BT <= 10) ? slab_b : 0;
  g.slab_bytes = (int)slab_b;
  g.dslab_bytes = (int)dslab_b;
  const size_t fixed = hs_b + 2 * slab_b + 2 * dslab_b + vt_b + dred_b + 128 + 256 + vb_b + hb_b;
  const size_t smem_max = 227 * 1024;
  if (fixed + 2 * (size_t)g.slot_bytes > smem_max) return g;
  g.nslots = (int)((smem_max - fixed) / g.slot_bytes);
  if (g.nslots > MAX_SLOTS) g.nslots = MAX_SLOTS;
  g.ring_bytes = g.nslots * g.slot_bytes;
  {
    // the flush parks G x BT x CQ float4 partials in the ring
    const size_t park = g.G > 2 ? (size_t)g.G * g.BT * g.CQ * 16 : 0;
    if ((size_t)g.ring_bytes < park) g.ring_bytes = (int)((park + 1023) & ~(size_t)1023);
    if (fixed + (size_t)g.ring_bytes > smem_max) return g;
  }
  const int narr = a.weightcost != 0.f ? 3 : 2;
  if ((size_t)narr * ((8 * a.ldw * 4 + 127) & ~127) > (size_t)g.ring_bytes) return g;
  size_t off = (size_t)g.ring_bytes;
  auto take = [&](size_t bytes) { size_t o = off; off += bytes; return (int)o; };
  g.off_hs = take(hs_b);
  g.off_v0 = take(2 * slab_b + 2 * dslab_b);
  g.off_vt = take(vt_b);
  g.off_dred = take(dred_b);
  g.off_bars = take(128);
  g.off_misc = take(256);
  g.off_vb = take(vb_b);
  g.off_hb = take(hb_b);
  g.smem = off;
  g.ok = g.smem <= smem_max;
  return g;
}

}  // namespace sk

bool skinny_supported(const mdbn_ctx* c, const mdbn_cd_args& a) {
  if (a.phase != MDBN_PHASE_FULL) return false;
  return sk::plan(c, a).ok;
}

int skinny_cd_step(mdbn_ctx* c, const mdbn_cd_args& a, cudaStream_t st) { return skinny_cd_steps(c, a, 1, st); }

// n_steps consecutive steps in ONE launch: a.indices is [n_steps][B], a.cost_out [n_steps], step s draws with
// rng.offset + s (PHILOX only).  Same results as n_steps single-step calls.
int skinny_cd_steps(mdbn_ctx* c, const mdbn_cd_args& a, int n_steps, cudaStream_t st) {
  sk::Geometry g = sk::plan(c, a, n_steps > 1);
  MDBN_CHECK(n_steps >= 1, "skinny path: n_steps must be >= 1");
  MDBN_CHECK(n_steps == 1 || a.rng.mode == MDBN_RNG_PHILOX, "skinny path: multi-step launches need the PHILOX generator");
  if (n_steps > 1 && ((!g.ok && sk::plan(c, a, false).ok) || (g.ok && g.BT > 10 && (size_t)a.V * a.ldw * 4 > ((size_t)8 << 20)))) {
    // one launch per step when the second slab set of the fused chain does not fit next to this layer's tiles, and
    // for B > 10 on large layers, where the looped instantiation cannot fuse (register budget) and measures slower
    // than single launches (19937x400, B = 20: 129 vs 121 us per step; small layers still gain from one launch)
    for (int s = 0; s < n_steps; ++s) {
      mdbn_cd_args one = a;
      one.indices = a.indices ? a.indices + (size_t)s * a.B : nullptr;
      one.cost_out = a.cost_out ? a.cost_out + s : nullptr;
      one.rng.offset = a.rng.offset + (unsigned long long)s;
      MDBN_TRY(skinny_cd_steps(c, one, 1, st));
    }
    return 0;
  }
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
  p.rows_per_cta = g.rows_per_cta; p.rows_small = g.rows_small; p.n_big = g.n_big; p.rows_alloc = g.rows_alloc;
  p.CQ = g.CQ; p.GW = g.GW; p.G = g.G; p.R = g.R; p.nbox = g.nbox; p.nslots = g.nslots;
  p.ldh = g.ldh; p.slot_bytes = g.slot_bytes; p.ring_bytes = g.ring_bytes;
  p.slab_bytes = g.slab_bytes;
  p.off_hs = g.off_hs; p.off_v0 = g.off_v0; p.off_vt = g.off_vt; p.off_dred = g.off_dred;
  p.off_bars = g.off_bars; p.off_misc = g.off_misc; p.off_vb = g.off_vb; p.off_hb = g.off_hb;

  // scratch: two sets of 5 fixed-point accumulators [BT][ldw] + cost partials.  A launch works in one set
  // (zero on entry) and clears the other; everything is re-zeroed when the layout changes.
  const size_t n_acc = (size_t)g.BT * a.ldw;
  const size_t n_cost = ((size_t)g.grid + 64 + 3) & ~(size_t)3;     // keeps PHf 16-byte aligned whatever the SM count
  const size_t total_b = 2 * 5 * n_acc * sizeof(unsigned long long) + (n_cost + n_acc) * sizeof(float);
  mdbn_ctx::Buf& wb = c->ws[WS_SKINNY];
  const void* before = wb.p;
  const size_t before_n = wb.n;
  unsigned long long* base = (unsigned long long*)ws_get(c, WS_SKINNY, total_b);
  if (!base) return 3;
  const unsigned long long key = ((unsigned long long)g.BT << 48) ^ ((unsigned long long)a.ldw << 24) ^
                                 ((unsigned long long)(uintptr_t)base << 1) ^ 1ULL;
  if (before != wb.p || before_n != wb.n || key != c->skinny_key) {
    MDBN_CUDA(cudaMemsetAsync(base, 0, wb.n, st));
    c->skinny_key = key;
    c->skinny_parity = 0;
  }
  p.n_acc = (int)n_acc;
  p.acc = base + (size_t)c->skinny_parity * 5 * n_acc;
  p.acc_other = base + (size_t)(1 - c->skinny_parity) * 5 * n_acc;
  c->skinny_parity = (c->skinny_parity + (unsigned)n_steps) & 1u;
  p.n_steps = n_steps;
  p.cost_part = reinterpret_cast<float*>(base + 2 * 5 * n_acc);
  p.PHf = p.cost_part + n_cost;
  p.bar = reinterpret_cast<unsigned long long*>(c->barrier);
  static const bool want_timing = getenv("MDBN_SKINNY_TIMING") != nullptr;
#ifdef MDBN_SKINNY_DEBUG
  static const int dbg_flags = getenv("MDBN_SKINNY_DEBUG") ? atoi(getenv("MDBN_SKINNY_DEBUG")) : 0;
  p.dbg_flags = dbg_flags;      // timing experiments only (skips parts of the passes): never in the release build
#endif
  p.dbg = want_timing ? reinterpret_cast<unsigned long long*>(c->barrier) + 8 : nullptr;
  CUtensorMap tms[2];
  MDBN_TRY(sk::make_map(&tms[0], a.W, a.V, a.ldw, g.R));
  MDBN_TRY(sk::make_map(&tms[1], a.W, a.V, a.ldw, 8));
  int rc = 2;
  if (g.BT == 10) rc = n_steps > 1 ? sk::launch<10, true>(c, tms, p, g, st) : sk::launch<10, false>(c, tms, p, g, st);
  else if (g.BT == 20) rc = n_steps > 1 ? sk::launch<20, true>(c, tms, p, g, st) : sk::launch<20, false>(c, tms, p, g, st);
  else set_error("skinny path: no kernel for BT=%d", g.BT);
  if (rc == 0 && p.dbg) {
    unsigned long long t[32];
    MDBN_CUDA(cudaStreamSynchronize(st));
    MDBN_CUDA(cudaMemcpy(t, p.dbg, sizeof(t), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[skinny timeline us] V=%d H=%d B=%d k=%d R=%d nslots=%d:", a.V, a.H, a.B, a.k, g.R, g.nslots);
    for (int i = 1; i < 13; ++i) fprintf(stderr, " %.1f", (double)((long long)(t[i] - t[0])) * 1e-3);
    fprintf(stderr, "  | stats: TMA wait %.1f us, tile barrier %.1f us (thread 0, @1.965 GHz)\n", (double)t[24] / 1965.0, (double)t[25] / 1965.0);
    {
      static unsigned long long se[512];
      MDBN_CUDA(cudaMemcpy(se, p.dbg + 32, (256 + g.grid) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      double smin = 1e30, smax = -1e30, emin = 1e30, emax = -1e30;
      for (int i = 0; i < g.grid; ++i) {
        const double s0 = (double)((long long)(se[i] - t[0])) * 1e-3, e0 = (double)((long long)(se[256 + i] - t[0])) * 1e-3;
        smin = s0 < smin ? s0 : smin; smax = s0 > smax ? s0 : smax; emin = e0 < emin ? e0 : emin; emax = e0 > emax ? e0 : emax;
      }
      fprintf(stderr, "[skinny cta spread us] start %.1f..%.1f  end %.1f..%.1f\n", smin, smax, emin, emax);
    }
  }
  return rc;
}

}  // namespace mdbn
