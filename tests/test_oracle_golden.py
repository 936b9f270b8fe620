"""The oracle (oracle/rbm_oracle.py) against the golden vectors produced by the
reference's own source under the Theano shim (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import rbm_oracle as O
from oracle import shared_u

from conftest import GOLDEN

TOL = dict(rtol=1e-11, atol=1e-12)


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def layer_from(g, dtype=np.float64, W="W"):
    V, H = g[W].shape
    L = O.Layer(V, H, int(g["kind"]), W=np.array(g[W], dtype), dtype=dtype, error_free=bool(g["error_free"]))
    return L


@pytest.mark.parametrize("name", ["phases_rbm", "phases_grbm", "phases_grbm_noisy"])
def test_phases(name):
    g = load(name)
    L = layer_from(g)
    L.hbias[...] = g["hbias"]
    L.vbias[...] = g["vbias"]
    v, hid, uh, uv, nv = g["v"], g["hid"], g["uh"].astype(float), g["uv"].astype(float), g["nv"].astype(float)
    vdraw = nv if L.kind == O.GRBM else uv
    np.testing.assert_allclose(np.stack(O.propup(L, v)), g["propup"], **TOL)
    pd = O.propdown(L, hid)
    if L.kind == O.RBM:
        np.testing.assert_allclose(np.stack(pd), g["propdown"], **TOL)
    else:
        # GRBM does not override propdown (src/rbm.py:215-227): the inherited one applies sigma
        np.testing.assert_allclose(pd[0], g["propdown"][0], **TOL)
        np.testing.assert_allclose(O.sigmoid(pd[0]), g["propdown"][1], **TOL)
    np.testing.assert_allclose(np.stack(O.sample_h_given_v(L, v, uh)), g["sample_h_given_v"], **TOL)
    np.testing.assert_allclose(np.stack(O.sample_v_given_h(L, hid, vdraw)), g["sample_v_given_h"], **TOL)
    r = O.gibbs_hvh(L, hid, vdraw, uh)
    np.testing.assert_allclose(np.stack(r[:3]), g["gibbs_hvh_v"], **TOL)
    np.testing.assert_allclose(np.stack(r[3:]), g["gibbs_hvh_h"], **TOL)
    r = O.gibbs_vhv(L, v, uh, vdraw)
    np.testing.assert_allclose(np.stack(r[:3]), g["gibbs_vhv_h"], **TOL)
    np.testing.assert_allclose(np.stack(r[3:]), g["gibbs_vhv_v"], **TOL)
    np.testing.assert_allclose(O.free_energy(L, v), g["free_energy"], **TOL)
    np.testing.assert_allclose(O.free_energy_gap(L, v, g["v2"]), g["free_energy_gap"], **TOL)


CD_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "cd_*.npz")))


def run_cd_case(g, dtype=np.float64):
    L = layer_from(g, dtype, W="W0")
    V, H, k, B_nom = int(g["V"]), int(g["H"]), int(g["k"]), int(g["B_nom"])
    snap = L.W.copy()
    P = np.zeros((B_nom, H), dtype) if bool(g["pcd"]) else None
    costs, states = [], []
    for t, (lo, hi) in enumerate(g["rows"]):
        v0 = g["data"][lo:hi].astype(dtype)
        U = shared_u.step_buffer(int(g["seed_u"]), int(g["layer_id"]), t, L.kind, L.error_free, hi - lo, V, H, k)
        costs.append(O.cd_step(L, v0, U, lr=float(g["lr"]), k=k, lambda_1=float(g["lambda_1"]),
                               lambda_2=float(g["lambda_2"]), weightcost=float(g["weightcost"]),
                               batch_size=B_nom, momentum=float(g["momentum"][t]), persistent=P, W_snap=snap))
        states.append(np.concatenate([L.W.ravel(), L.hbias, L.vbias]))
    return L, P, np.array(costs), np.stack(states)


@pytest.mark.parametrize("name", CD_CASES)
def test_cd_sequences_f64(name):
    g = load(name)
    L, P, costs, states = run_cd_case(g)
    np.testing.assert_allclose(costs, g["costs"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(states, g["states"], **TOL)
    for n in ("W", "hbias", "vbias", "W_speed", "hbias_speed", "vbias_speed"):
        np.testing.assert_allclose(getattr(L, n), g[n], **TOL)
    if P is not None:
        np.testing.assert_array_equal(P, g["persistent"])


@pytest.mark.parametrize("name", CD_CASES)
def test_cd_sequences_f32_tracks(name):
    """float32 oracle stays within fp32 noise of the float64 truth (bounds the
    noise the CUDA fp32 path is allowed)."""
    g = load(name)
    L, P, costs, states = run_cd_case(g, np.float32)
    # a Bernoulli flip at |u-p| ~ 1e-7 is possible in principle; none occurs on these seeds
    np.testing.assert_allclose(L.W, g["W"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(costs, g["costs"], rtol=1e-3)


def test_minibatches():
    g = load("minibatches")
    for n, b in ((170, 20), (23, 5), (20, 20), (7, 10)):
        np.random.seed(n * 100 + b)
        rng_idx, mbs = O.get_minibatches_idx(n, b, shuffle=True)
        assert [len(m) for m in mbs] == list(g["n%d_b%d_lens" % (n, b)])
        np.testing.assert_array_equal(np.concatenate(mbs), g["n%d_b%d" % (n, b)])
        assert mbs[0].dtype == np.int32
    _, mbs = O.get_minibatches_idx(0, 5)
    assert mbs == []


@pytest.mark.parametrize("name", ["dbn_gauss", "dbn_bern"])
def test_dbn_training_loop(name):
    g = load(name)
    sizes = [int(s) for s in g["sizes"]]
    k, B = int(g["k"]), int(g["B"])
    d = O.DBN(numpy_rng=np.random.RandomState(int(g["seed"])), n_ins=int(g["n_ins"]), gauss=bool(g["gauss"]),
              hidden_layers_sizes=sizes[:-1], n_outs=sizes[-1])
    for i, L in enumerate(d.layers):
        np.testing.assert_array_equal(L.W, g["W0_%d" % i])     # same RandomState draw order (src/dbn.py:114,155)

    def u_provider(layer, call, b):
        L = d.layers[layer]
        return shared_u.step_buffer(int(g["seed_u"]), layer, call, L.kind, True, b, L.n_visible, L.n_hidden, k)
    np.random.seed(int(g["shuffle_seed"]))
    hist = d.training(g["train"], B, k, [int(e) for e in g["epochs"]], [float(x) for x in g["lrs"]],
                      lambda_1=float(g["lambda_1"]), lambda_2=float(g["lambda_2"]),
                      validation_x=g["val"] if "val" in g else None, u_provider=u_provider)
    assert [h["calls"] for h in hist] == list(g["n_calls"])    # early stopping fires at the same iteration
    printed = [c for h in hist for (_, c, _) in h["validations"]]
    np.testing.assert_allclose(printed, g["printed_costs"], rtol=1e-6)   # '%s' of a float64: 12+ digits
    fegs = [f for h in hist for (_, _, f) in h["validations"] if f is not None]
    np.testing.assert_allclose(fegs, g["printed_fegs"], rtol=1e-6)
    for i, L in enumerate(d.layers):
        np.testing.assert_allclose(L.W, g["W_%d" % i], **TOL)
        np.testing.assert_allclose(L.hbias, g["b_%d" % i], **TOL)
        np.testing.assert_allclose(L.vbias, g["vb_%d" % i], **TOL)
    np.testing.assert_allclose(d.get_output(g["train"]), g["out_train"], **TOL)


@pytest.mark.parametrize("name", ["train_rbm_pcd", "train_rbm_cd2", "train_grbm"])
def test_standalone_training_loop(name):
    """RBM.training / GRBM.training / learn_model (src/rbm.py:484-629, 701-728) as run by the reference itself:
    PCD default with a chain of zeros, momentum switch at 0-based epoch 6, per-epoch mean cost and free-energy gap."""
    g = load(name)
    kind, V, H, B, k = int(g["kind"]), int(g["V"]), int(g["H"]), int(g["B"]), int(g["k"])
    rng = np.random.RandomState(int(g["seed"]))
    rng.randint(2 ** 30)                                        # the Theano stream seed is drawn first (src/rbm.py:92)
    L = O.Layer(V, H, kind, numpy_rng=rng)
    np.testing.assert_array_equal(L.W, g["W0"])
    kw = {key[3:]: g[key].item() for key in g if key.startswith("kw_")}
    kw.pop("k", None)
    np.random.seed(int(g["shuffle_seed"]))
    hist = O.rbm_training(L, g["train"], g["val"], int(g["epochs"]), batch_size=B, k=k,
                          u_provider=lambda call, b: shared_u.step_buffer(int(g["seed_u"]), 0, call, kind, True, b, V, H, k),
                          **kw)
    np.testing.assert_allclose([h[0] for h in hist], g["costs"], rtol=1e-9)
    np.testing.assert_allclose([h[1] for h in hist], g["fegs"], rtol=1e-9)
    for name_ in ("W", "hbias", "vbias", "W_speed", "hbias_speed", "vbias_speed"):
        np.testing.assert_allclose(getattr(L, name_), g[name_], **TOL)


def test_free_energy_bruteforce():
    """F(v) == -log sum_h exp(-E(v,h)) for a tiny Bernoulli RBM."""
    rs = np.random.RandomState(0)
    V, H = 6, 5
    L = O.Layer(V, H, O.RBM, numpy_rng=rs)
    L.hbias[...] = rs.randn(H)
    L.vbias[...] = rs.randn(V)
    v = (rs.rand(4, V) < 0.5).astype(float)
    hs = np.array([[(i >> j) & 1 for j in range(H)] for i in range(2 ** H)], float)
    E = -(v @ L.vbias)[:, None] - (hs @ L.hbias)[None, :] - v @ L.W @ hs.T
    np.testing.assert_allclose(O.free_energy(L, v), -np.log(np.exp(-E).sum(1)), rtol=1e-12)


def test_rbm_grad_is_free_energy_gradient():
    """compute_rbm_grad with samples in place of means and wc=0 equals the finite-difference
    gradient of mean F(v_k) - mean F(v0)  (src/rbm.py:386-389 vs :411-412)."""
    rs = np.random.RandomState(1)
    V, H, B = 5, 4, 3
    L = O.Layer(V, H, O.RBM, numpy_rng=rs)
    v0 = (rs.rand(B, V) < 0.5).astype(float)
    vk = (rs.rand(B, V) < 0.5).astype(float)
    cost = lambda: O.free_energy(L, vk).mean() - O.free_energy(L, v0).mean()
    g = (v0.T @ O.propup(L, v0)[1] - vk.T @ O.propup(L, vk)[1]) / B
    num = np.zeros_like(L.W)
    for i in range(V):
        for j in range(H):
            w = L.W[i, j]
            L.W[i, j] = w + 1e-6
            a = cost()
            L.W[i, j] = w - 1e-6
            b = cost()
            L.W[i, j] = w
            num[i, j] = (a - b) / 2e-6
    np.testing.assert_allclose(g, num, atol=1e-8)


def test_symbolic_gradient_identity_grbm():
    """the reference's alternative `compute_symbolic_grad` (src/rbm.py:378-390) differentiates
    mean F(v0) - mean F(chain_end) with the chain end held constant.  For the error-free GRBM (mean-field visibles:
    chain end == negative visible mean) that gradient IS the statistics `compute_rbm_grad` builds (:392-419).  Checked on
    the oracle in fp64 with central differences of its own free_energy against the packed statistics of cd_stats."""
    rs = np.random.RandomState(11)
    V, H, B, k = 7, 5, 4, 2
    L = O.Layer(V, H, O.GRBM, W=rs.randn(V, H) * 0.3, hbias=rs.randn(H) * 0.1, vbias=rs.randn(V) * 0.1)
    v0 = rs.randn(B, V)
    U = rs.uniform(size=O.u_size(O.GRBM, True, B, V, H, k))
    packed = O.cd_stats(L, v0, U, k=k)
    gW, ghb, gvb = packed[:V * H].reshape(V, H) / B, packed[V * H:V * H + H] / B, packed[V * H + H:V * H + H + V] / B
    # chain end, replayed exactly as cd_stats runs it
    u = O._views(U, O.GRBM, True, B, V, H, k)
    _, _, h = O.sample_h_given_v(L, v0, u["hpos"])
    for s in range(k):
        _, vk, _, _, _, h = O.gibbs_hvh(L, h, u.get("v%d" % s), u["h%d" % s])

    def cost():
        return O.free_energy(L, v0).mean() - O.free_energy(L, vk).mean()

    def numgrad(arr):
        g = np.zeros_like(arr)
        it = np.nditer(arr, flags=["multi_index"])
        for _ in it:
            i = it.multi_index
            old = arr[i]
            arr[i] = old + 1e-6
            cp = cost()
            arr[i] = old - 1e-6
            cm = cost()
            arr[i] = old
            g[i] = (cp - cm) / 2e-6
        return g
    # the update direction is MINUS the gradient of the cost (src/rbm.py:386-390)
    np.testing.assert_allclose(-numgrad(L.W), gW, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(-numgrad(L.hbias), ghb, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(-numgrad(L.vbias), gvb, rtol=1e-6, atol=1e-8)
