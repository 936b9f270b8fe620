#!/usr/bin/env python
"""Regenerate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN SOURCE
(/root/reference/src/{rbm,dbn,MDBN,mlp,utils}.py, unmodified, imported in place)
under oracle/theano_shim.  Only runnable in the build container (the reference
tree does not travel to the GPU box); the .npz files it writes are committed.

    python tests/golden/make_golden.py

All cases run with floatX=float64 ("truth"); randomness comes from
oracle/shared_u.step_buffer (shared-uniform-buffer mode).
"""
import io
import os
import sys
from contextlib import redirect_stdout, redirect_stderr

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import theano_shim as shim          # noqa: E402
from oracle import rbm_oracle as O              # noqa: E402
from oracle import shared_u                     # noqa: E402

theano = shim.install("/root/reference/src", floatX="float64")
import rbm as ref_rbm                            # noqa: E402  (reference source)
import dbn as ref_dbn                            # noqa: E402
import MDBN as ref_mdbn                          # noqa: E402
import utils as ref_utils                        # noqa: E402
from theano import tensor                        # noqa: E402
from theano.sandbox.rng_mrg import MRG_RandomStreams as RandomStreams  # noqa: E402

SEED_U = 20161230


class Provider:
    """Routes every random node of a compiled function to its slice of the
    per-step shared buffer."""

    def __init__(self):
        self.fns = {}       # id(fn) -> spec
        self.explicit = {}  # ordinal -> array (ad-hoc phase functions)

    def register(self, fn, base, layer, kind, ef, V, H, k):
        self.fns[id(fn)] = dict(base=base, layer=layer, kind=kind, ef=ef, V=V, H=H, k=k)

    def __call__(self, ordinal, what, shape, fn, call_idx):
        if id(fn) not in self.fns:
            return self.explicit[ordinal]
        s = self.fns[id(fn)]
        B = shape[0]
        buf = shared_u.step_buffer(SEED_U, s["layer"], call_idx, s["kind"], s["ef"], B, s["V"], s["H"], s["k"])
        lay, _ = O.u_layout(s["kind"], s["ef"], B, s["V"], s["H"], s["k"])
        name, off, sh = lay[ordinal - s["base"]]
        assert tuple(sh) == tuple(shape), (name, sh, shape)
        return buf[off:off + sh[0] * sh[1]].reshape(sh).astype(np.float64)


PROV = Provider()
shim.set_rng_provider(PROV)


def make_data(rs, n, V, kind, real01=False):
    if kind == O.GRBM:
        x = rs.randn(n, V)
        return (x - x.mean(0)) / x.std(0)
    if real01:
        return rs.rand(n, V)
    return (rs.rand(n, V) < 0.3).astype(np.float64)


def build(kind, V, H, seed, ef=True):
    x = tensor.matrix("x")
    rng = np.random.RandomState(seed)
    trng = RandomStreams(rng.randint(2 ** 30))
    if kind == O.GRBM:
        m = ref_rbm.GRBM(input=x, n_visible=V, n_hidden=H, numpy_rng=rng, theano_rng=trng, error_free=ef)
    else:
        m = ref_rbm.RBM(input=x, n_visible=V, n_hidden=H, numpy_rng=rng, theano_rng=trng)
    return x, m


# ---------------------------------------------------------------------------
# A. phase-level functions (propup ... free_energy)   src/rbm.py:166-256, 647-688
# ---------------------------------------------------------------------------
def phases_case(name, kind, V, H, B, seed, ef=True):
    x, m = build(kind, V, H, seed, ef)
    rs = np.random.RandomState(seed + 1)
    # non-trivial biases
    m.hbias.set_value(rs.randn(H) * 0.3)
    m.vbias.set_value(rs.randn(V) * 0.3)
    h = tensor.matrix("h")
    v = make_data(rs, B, V, kind)
    hid = (rs.rand(B, H) < 0.5).astype(np.float64)
    g = shared_u._gen(SEED_U, 99, seed)
    uh, uv = shared_u.uniforms(g, B * H).reshape(B, H), shared_u.uniforms(g, B * V).reshape(B, V)
    nv = g.standard_normal((B, V)).astype(np.float32)
    out = dict(kind=kind, error_free=ef, W=m.W.get_value(), hbias=m.hbias.get_value(),
               vbias=m.vbias.get_value(), v=v, hid=hid, uh=uh, uv=uv, nv=nv)

    def run(expr, inp, val, draws):
        base = m.theano_rng.n_nodes - len(draws)
        PROV.explicit = {base + i: d.astype(np.float64) for i, d in enumerate(draws)}
        return theano.function([inp], expr)(val)

    vdraw = nv if kind == O.GRBM else uv
    n_v = 1 if (kind == O.RBM or not ef) else 0
    out["propup"] = np.stack(run(m.propup(x), x, v, []))
    out["propdown"] = np.stack(run(m.propdown(h), h, hid, []))
    out["sample_h_given_v"] = np.stack(run(m.sample_h_given_v(x), x, v, [uh]))
    out["sample_v_given_h"] = np.stack(run(m.sample_v_given_h(h), h, hid, [vdraw][:n_v]))
    r = run(m.gibbs_hvh(h), h, hid, [vdraw][:n_v] + [uh])
    out["gibbs_hvh_v"], out["gibbs_hvh_h"] = np.stack(r[:3]), np.stack(r[3:])
    r = run(m.gibbs_vhv(x), x, v, [uh] + [vdraw][:n_v])
    out["gibbs_vhv_h"], out["gibbs_vhv_v"] = np.stack(r[:3]), np.stack(r[3:])
    out["free_energy"] = run(m.free_energy(x), x, v, [])
    v2 = make_data(rs, B + 2, V, kind)
    out["v2"] = v2
    x2 = tensor.matrix("x2")
    out["free_energy_gap"] = theano.function([x, x2], m.free_energy_gap(x, x2))(v, v2)
    fa, fb = theano.function([x, x2], m.free_energies(x, x2))(v, v2)
    out["free_energies_a"], out["free_energies_b"] = fa, fb
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name)


# ---------------------------------------------------------------------------
# B. CD-k / PCD-k step sequences             src/rbm.py:258-376, 392-447
# ---------------------------------------------------------------------------
def cd_case(name, kind, V, H, B_nom, k, n_steps, seed, lr, momentum, lambda_1=0.0, lambda_2=0.0,
            weightcost=0.0, pcd=False, ef=True, real01=False, tail=None, layer_id=0):
    x, m = build(kind, V, H, seed, ef)
    rs = np.random.RandomState(seed + 7)
    N = B_nom * n_steps
    data = make_data(rs, N, V, kind, real01)
    W0 = m.W.get_value()
    persistent = theano.shared(np.zeros((B_nom, H)), borrow=True) if pcd else None
    base = m.theano_rng.n_nodes
    cost, updates = m.get_cost_updates(lr=lr, k=k, lambda_1=lambda_1, lambda_2=lambda_2,
                                       weightcost=weightcost, batch_size=B_nom, persistent=persistent)
    mom = tensor.scalar("momentum")
    fn = theano.function([x, mom], cost, updates=updates, givens={m.momentum: mom})
    PROV.register(fn, base, layer_id, kind, ef, V, H, k)
    costs, states = [], []
    moms = momentum if isinstance(momentum, (list, tuple)) else [momentum] * n_steps
    rows = []
    for t in range(n_steps):
        lo, hi = t * B_nom, (t + 1) * B_nom
        if tail is not None and t == n_steps - 1:
            hi = lo + tail                       # ragged tail minibatch (src/utils.py:71-73)
        rows.append((lo, hi))
        costs.append(float(fn(data[lo:hi], moms[t])))
        states.append(np.concatenate([m.W.get_value().ravel(), m.hbias.get_value(), m.vbias.get_value()]))
    out = dict(kind=kind, error_free=ef, V=V, H=H, B_nom=B_nom, k=k, lr=lr, momentum=np.array(moms),
               lambda_1=lambda_1, lambda_2=lambda_2, weightcost=weightcost, pcd=pcd, layer_id=layer_id,
               seed_u=SEED_U, data=data, rows=np.array(rows), W0=W0, costs=np.array(costs),
               states=np.stack(states),
               W=m.W.get_value(), hbias=m.hbias.get_value(), vbias=m.vbias.get_value(),
               W_speed=m.W_speed.get_value(), hbias_speed=m.hbias_speed.get_value(),
               vbias_speed=m.vbias_speed.get_value())
    if pcd:
        out["persistent"] = persistent.get_value()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "costs", np.round(costs, 5))


# ---------------------------------------------------------------------------
# C. DBN greedy loop with early stopping      src/dbn.py:64-204, 238-517
# ---------------------------------------------------------------------------
def register_dbn(d, k, layer_base=0):
    """Hook the compiled train fns of a reference DBN to the provider.  The
    shim's RandomStreams numbers nodes in creation order: per layer 1+k (GRBM,
    error_free) or 1+2k (RBM) nodes, in layer order (src/dbn.py:280-294)."""
    return d, k, layer_base


def dbn_case(name, n_ins, sizes, N, n_val, B, k, epochs, lrs, lambda_1, lambda_2, gauss, seed, shuffle_seed):
    rs = np.random.RandomState(seed + 3)
    kind0 = O.GRBM if gauss else O.RBM
    train = make_data(rs, N, n_ins, kind0)
    val = make_data(rs, n_val, n_ins, kind0) if n_val else None
    rng = np.random.RandomState(seed)
    d = ref_dbn.DBN(numpy_rng=rng, n_ins=n_ins, gauss=gauss, hidden_layers_sizes=sizes[:-1], n_outs=sizes[-1])
    W0 = [L.W.get_value() for L in d.rbm_layers]
    # intercept training_functions so the compiled fns can be registered
    orig_tf = d.training_functions
    captured = {}

    def tf(*a, **kw):
        fns, fegs = orig_tf(*a, **kw)
        base = 0
        for i, (f, L) in enumerate(zip(fns, d.rbm_layers)):
            kd = O.GRBM if isinstance(L, ref_rbm.GRBM) else O.RBM
            PROV.register(f, base, i, kd, True, L.n_visible, L.n_hidden, k)
            base += (1 + k) if kd == O.GRBM else (1 + 2 * k)
        captured["fns"] = fns
        return fns, fegs
    d.training_functions = tf
    np.random.seed(shuffle_seed)
    buf = io.StringIO()
    with redirect_stdout(buf), redirect_stderr(io.StringIO()):
        d.training(theano.shared(train, borrow=True), B, k, epochs, lrs, lambda_1=lambda_1,
                   lambda_2=lambda_2, validation_set_x=(theano.shared(val, borrow=True) if n_val else None))
    log = buf.getvalue()
    # parse the reference's own printed monitor values (src/dbn.py:463-504)
    costs, fegs = [], []
    for line in log.splitlines():
        if line.startswith("Pre-training cost"):
            costs.append(float(line.split(":")[1]))
        if line.startswith("Free energy gap"):
            fegs.append(float(line.split(":")[1]))
    out = dict(n_ins=n_ins, sizes=np.array(sizes), B=B, k=k, epochs=np.array(epochs), lrs=np.array(lrs),
               lambda_1=lambda_1, lambda_2=lambda_2, gauss=gauss, seed=seed, shuffle_seed=shuffle_seed,
               seed_u=SEED_U, train=train, n_calls=np.array([f.n_calls for f in captured["fns"]]),
               printed_costs=np.array(costs), printed_fegs=np.array(fegs),
               out_train=d.get_output(theano.shared(train)))
    if n_val:
        out["val"] = val
    for i, L in enumerate(d.rbm_layers):
        out["W0_%d" % i] = W0[i]
        out["W_%d" % i] = L.W.get_value()
        out["b_%d" % i] = L.hbias.get_value()
        out["vb_%d" % i] = L.vbias.get_value()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "calls", out["n_calls"], "n_printed", len(costs), len(fegs))


# ---------------------------------------------------------------------------
# D. utils.get_minibatches_idx                 src/utils.py:54-75
# ---------------------------------------------------------------------------
def minibatch_case():
    out = {}
    for n, b in ((170, 20), (23, 5), (20, 20), (7, 10)):
        np.random.seed(n * 100 + b)
        _, mbs = ref_utils.get_minibatches_idx(n, b, shuffle=True)
        out["n%d_b%d" % (n, b)] = np.concatenate(mbs)
        out["n%d_b%d_lens" % (n, b)] = np.array([len(m) for m in mbs])
    np.savez_compressed(os.path.join(HERE, "minibatches.npz"), **out)
    print("wrote minibatches")


# ---------------------------------------------------------------------------
# E. MDBN.train_bottom_layer + train_top        src/MDBN.py:31-76
# ---------------------------------------------------------------------------
def mdbn_case(name, seed=5):
    rs = np.random.RandomState(seed)
    N = 24
    mods = {"ME": (15, [6], [40], [0.005], 2, 0.01, 0.01),
            "GE": (31, [10, 6], [60, 30], [0.005, 0.1], 1, 0.01, 0.1)}
    rng = np.random.RandomState(123)
    out = dict(N=N, seed_u=SEED_U)
    tops = []
    for li, (mn, (V, sizes, ep, lr, k, l1, l2)) in enumerate(mods.items()):
        data = make_data(rs, N, V, O.GRBM)
        out[mn + "_data"] = data
        # hook: DBN is constructed inside train_bottom_layer; wrap the class
        orig = ref_mdbn.DBN

        class Hooked(orig):
            def training_functions(self, *a, **kw):
                fns, fegs = orig.training_functions(self, *a, **kw)
                base = 0
                for i, (f, L) in enumerate(zip(fns, self.rbm_layers)):
                    kd = O.GRBM if isinstance(L, ref_rbm.GRBM) else O.RBM
                    PROV.register(f, base, 10 * (li + 1) + i, kd, True, L.n_visible, L.n_hidden, kw["k"])
                    base += (1 + kw["k"]) if kd == O.GRBM else (1 + 2 * kw["k"])
                return fns, fegs
        ref_mdbn.DBN = Hooked
        np.random.seed(1000 + li)
        with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
            d, o_tr, _ = ref_mdbn.train_bottom_layer(theano.shared(data, borrow=True), None, batch_size=5, k=k,
                                                     layers_sizes=sizes, pretraining_epochs=ep, pretrain_lr=lr,
                                                     lambda_1=l1, lambda_2=l2, rng=rng)
        ref_mdbn.DBN = orig
        tops.append(o_tr)
        out[mn + "_out"] = o_tr
        for i, L in enumerate(d.rbm_layers):
            out["%s_W_%d" % (mn, i)] = L.W.get_value()
            out["%s_b_%d" % (mn, i)] = L.hbias.get_value()
    joint = np.concatenate(tops, axis=1)
    out["joint"] = joint
    orig = ref_mdbn.DBN

    class HookedTop(orig):
        def training_functions(self, *a, **kw):
            fns, fegs = orig.training_functions(self, *a, **kw)
            base = 0
            for i, (f, L) in enumerate(zip(fns, self.rbm_layers)):
                PROV.register(f, base, 90 + i, O.RBM, True, L.n_visible, L.n_hidden, kw["k"])
                base += 1 + 2 * kw["k"]
            return fns, fegs
    ref_mdbn.DBN = HookedTop
    np.random.seed(2000)
    with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
        top = ref_mdbn.train_top(5, False, theano.shared(joint, borrow=True), None, rng)
    ref_mdbn.DBN = orig
    for i, L in enumerate(top.rbm_layers):
        out["top_W_%d" % i] = L.W.get_value()
        out["top_b_%d" % i] = L.hbias.get_value()
    out["top_out"] = top.get_output(theano.shared(joint))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name)



# ---------------------------------------------------------------------------
# F. RBM.training / GRBM.training / learn_model   src/rbm.py:484-629, 701-728
# ---------------------------------------------------------------------------
def training_case(name, kind, V, H, N, n_val, B, epochs, seed, shuffle_seed, **kw):
    """The standalone epoch loop: PCD by default for the RBM (chain of zeros, :490-498), always CD for the GRBM
    (`persistent` is accepted and not forwarded, :711-717), momentum switch at 0-based epoch 6 (:584), per-epoch
    mean cost and free-energy gap on train[0:n_val] vs the validation set (:597-600)."""
    x, m = build(kind, V, H, seed)
    rs = np.random.RandomState(seed + 5)
    train, val = make_data(rs, N, V, kind), make_data(rs, n_val, V, kind)
    W0 = m.W.get_value()
    base = m.theano_rng.n_nodes
    k = kw.get("k", 1)
    orig_function = theano.function

    def hooked(*a, **k2):
        f = orig_function(*a, **k2)
        if k2.get("name") == "train_rbm":
            PROV.register(f, base, 0, kind, True, V, H, k)
        return f
    theano.function = hooked
    ref_rbm.theano.function = hooked
    np.random.seed(shuffle_seed)
    buf = io.StringIO()
    try:
        with redirect_stdout(buf), redirect_stderr(io.StringIO()):
            m.training(theano.shared(train, borrow=True), theano.shared(val, borrow=True), epochs, batch_size=B, **kw)
    finally:
        theano.function = orig_function
        ref_rbm.theano.function = orig_function
    costs, fegs = [], []
    for line in buf.getvalue().splitlines():
        if line.startswith("Training epoch"):
            costs.append(float(line.split("cost is")[1]))
        if line.startswith("Free energy gap is"):
            fegs.append(float(line.split("is")[1]))
    assert len(costs) == epochs and len(fegs) == epochs, (len(costs), len(fegs))
    out = dict(kind=kind, V=V, H=H, N=N, n_val=n_val, B=B, epochs=epochs, seed=seed, shuffle_seed=shuffle_seed,
               seed_u=SEED_U, k=k, train=train, val=val, W0=W0, costs=np.array(costs), fegs=np.array(fegs),
               W=m.W.get_value(), hbias=m.hbias.get_value(), vbias=m.vbias.get_value(),
               W_speed=m.W_speed.get_value(), hbias_speed=m.hbias_speed.get_value(),
               vbias_speed=m.vbias_speed.get_value())
    for key, v in kw.items():
        out["kw_" + key] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "costs", np.round(costs, 4), "feg", np.round(fegs, 4))


def io_case():
    """the table loader / pre-processing before the path, the class extraction after it and the
    .npz checkpoint written by the experiment scripts — all produced by the reference's own functions."""
    import tempfile
    import shutil
    import AMLsm as ref_aml                       # reference experiment script (save_network / load_network)
    tmp = tempfile.mkdtemp()
    out = {}
    try:
        rs = np.random.RandomState(99)
        n_feat, n_pers = 40, 12
        table = rs.lognormal(1.0, 0.8, size=(n_feat, n_pers))
        table[7] = 3.25                            # constant feature -> NaN z-score -> dropped
        table[19] = table[19].round(1)
        with open(os.path.join(tmp, "toy.tsv"), "w") as f:
            f.write("gene\t" + "\t".join("P%02d" % i for i in range(n_pers)) + "\n")
            for i in range(n_feat):
                f.write("G%03d\t" % i + "\t".join("%.6f" % x for x in table[i]) + "\n")
        out["table_text"] = np.frombuffer(open(os.path.join(tmp, "toy.tsv"), "rb").read(), dtype=np.uint8)
        cases = {"a": dict(holdout=0.25, repeats=1, clip=None, shuffle=True, seed=11, transform=False),
                 "b": dict(holdout=0.2, repeats=3, clip=(-1.0, 1.0), shuffle=False, seed=12, transform=False),
                 "c": dict(holdout=0.0, repeats=2, clip=None, shuffle=True, seed=13, transform=True)}
        for name, c in cases.items():
            np.random.seed(c["seed"])
            tr, va = ref_utils.load_n_preprocess_data("toy.tsv", dtype="float64", holdout=c["holdout"], clip=c["clip"],
                                                      transform_fn=np.power if c["transform"] else None, exponent=0.5,
                                                      repeats=c["repeats"], shuffle=c["shuffle"], datadir=tmp)
            out["prep_%s_train" % name] = tr.get_value()
            out["prep_%s_val" % name] = va.get_value() if va is not None else np.zeros((0, 0))
        # class extraction
        bits = (rs.rand(40, 5) > 0.45).astype(np.float64)
        bits[5:12] = bits[0]
        bits[20:29] = bits[1]
        labels, dist = ref_utils.find_unique_classes(bits)
        out["cls_bits"], out["cls_labels"], out["cls_dist"] = bits, labels, dist
        for n in (2, 3, 5):
            out["cls_remap_%d" % n] = ref_utils.remap_class(labels.astype(int), dist, n)
        # checkpoint written by the reference
        rng = np.random.RandomState(5)

        def net(n_ins, sizes, gauss=True):
            with redirect_stdout(io.StringIO()):
                d = ref_dbn.DBN(numpy_rng=rng, n_ins=n_ins, gauss=gauss, hidden_layers_sizes=sizes[:-1], n_outs=sizes[-1])
            for p_ in d.params:
                if p_.name == "b":
                    p_.set_value(rng.randn(*p_.get_value().shape))
            return d
        me, ge, sm, top = net(9, [4]), net(13, [6, 3]), net(7, [5]), net(12, [6, 2], gauss=False)
        classes = np.arange(6) % 3
        ref_aml.save_network(classes, ge, me, sm, None, top, 0.1, "ref_ckpt.npz", tmp, 2)
        shutil.copy(os.path.join(tmp, "ref_ckpt.npz"), os.path.join(HERE, "ref_checkpoint.npz"))
        x = rng.randn(4, 13)
        out["ckpt_ge_in"], out["ckpt_ge_out"] = x, ge.get_output(theano.shared(x))
        # and read back by the reference's own loader
        np_load = np.load                          # the reference predates allow_pickle=False (NumPy 1.16.3)
        np.load = lambda *a, **k: np_load(*a, **dict(k, allow_pickle=True))
        try:
            with redirect_stdout(io.StringIO()):
                me2, ge2, sm2, _, top2 = ref_aml.load_network("ref_ckpt.npz", tmp)
        finally:
            np.load = np_load
        out["ckpt_roundtrip_ok"] = np.array([np.array_equal(a.get_value(), b.get_value())
                                             for a, b in zip(ge.params + top.params, ge2.params + top2.params)])
        # MNIST idx loader / normalisation / tilings of the demo (src/MNIST.py is Python 2: give it xrange,
        # numpy.int and the float type it asks theano for)
        import gzip, struct
        import MNIST as ref_mnist
        ref_mnist.xrange = range
        if not hasattr(np, "int"):
            np.int = int
        nimg, sy, sx = 7, 4, 4      # square, like MNIST: the reference reshapes to (sizeX, sizeY) into a (sizeY, sizeX) slot
        pix = rs.randint(0, 256, size=(nimg, sy * sx)).astype(np.uint8)
        lab = np.array([3, 1, 4, 1, 5, 2, 6], dtype=np.uint8)
        img_bytes = struct.pack(">IIII", 2051, nimg, sy, sx) + pix.tobytes()
        lab_bytes = struct.pack(">II", 2049, nimg) + lab.tobytes()
        with gzip.open(os.path.join(tmp, "img.gz"), "wb") as f:
            f.write(img_bytes)
        with gzip.open(os.path.join(tmp, "lab.gz"), "wb") as f:
            f.write(lab_bytes)
        out["mnist_img_bytes"] = np.frombuffer(img_bytes, dtype=np.uint8)
        out["mnist_lab_bytes"] = np.frombuffer(lab_bytes, dtype=np.uint8)
        with redirect_stdout(io.StringIO()):
            mn = ref_mnist.MNIST("img.gz", "lab.gz", tmp)
        out["mnist_images"], out["mnist_labels"] = mn.images, mn.labels
        out["mnist_meta"] = np.array([mn.n_images, mn.sizeY, mn.sizeX, mn.n_levels])
        out["mnist_norm"] = mn.normalize(mn.images)
        Wd = rs.randn(sy * sx, 7)
        out["mnist_W"] = Wd
        out["mnist_tiles_w7"] = mn.display_weigths(Wd, 7)
        out["mnist_tiles_w4"] = mn.display_weigths(Wd[:, :4], 4)
        smp = [[rs.rand(sy * sx) for _ in range(3)] for _ in range(2)]
        out["mnist_samples"] = np.array(smp)
        out["mnist_tiles_s"] = mn.display_samples(smp)
    finally:
        shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(HERE, "io_cases.npz"), **out)
    print("io_cases:", sorted(out))

if __name__ == "__main__":
    if sys.argv[1:] == ["io"]:          # only the 8f rows (the other files are unchanged by it)
        io_case()
        sys.exit(0)
    if sys.argv[1:] == ["training"]:    # only the standalone epoch loops
        training_case("train_rbm_pcd", O.RBM, 24, 10, 45, 6, 5, 8, seed=61, shuffle_seed=881, learning_rate=0.1, k=1,
                      initial_momentum=0.5, final_momentum=0.9, weightcost=0.0002)
        training_case("train_rbm_cd2", O.RBM, 24, 10, 43, 6, 5, 7, seed=62, shuffle_seed=882, learning_rate=0.05, k=2,
                      initial_momentum=0.5, final_momentum=0.9, weightcost=0.0002, persistent=False)
        training_case("train_grbm", O.GRBM, 24, 10, 45, 6, 5, 8, seed=63, shuffle_seed=883, learning_rate=0.01, k=1,
                      initial_momentum=0.0, final_momentum=0.5, lambda_1=0.01, lambda_2=0.1, persistent=True)
        sys.exit(0)
    phases_case("phases_rbm", O.RBM, 13, 7, 5, seed=11)
    phases_case("phases_grbm", O.GRBM, 13, 7, 5, seed=12)
    phases_case("phases_grbm_noisy", O.GRBM, 13, 7, 5, seed=13, ef=False)

    cd_case("cd_rbm_cd1", O.RBM, 13, 7, 5, k=1, n_steps=4, seed=21, lr=0.1, momentum=[0.6, 0.6, 0.9, 0.9],
            weightcost=0.0002)
    cd_case("cd_rbm_cd3", O.RBM, 13, 7, 5, k=3, n_steps=3, seed=22, lr=0.1, momentum=0.5, weightcost=0.0002)
    cd_case("cd_rbm_pcd2", O.RBM, 13, 7, 5, k=2, n_steps=4, seed=23, lr=0.1, momentum=0.6, weightcost=0.0002,
            pcd=True)
    cd_case("cd_rbm_pcd1_real", O.RBM, 13, 7, 5, k=1, n_steps=3, seed=24, lr=0.05, momentum=0.0, pcd=True,
            real01=True)
    cd_case("cd_rbm_tail", O.RBM, 13, 7, 5, k=1, n_steps=3, seed=25, lr=0.1, momentum=0.6, weightcost=0.0002,
            tail=3)
    cd_case("cd_grbm_cd1", O.GRBM, 13, 7, 5, k=1, n_steps=4, seed=31, lr=0.005, momentum=0.0, lambda_1=0.01,
            lambda_2=0.1)
    cd_case("cd_grbm_cd3_mom", O.GRBM, 13, 7, 5, k=3, n_steps=3, seed=32, lr=0.01, momentum=0.7, lambda_1=0.02,
            lambda_2=0.05, weightcost=0.001)
    cd_case("cd_grbm_noisy_cd2", O.GRBM, 13, 7, 5, k=2, n_steps=3, seed=33, lr=0.005, momentum=0.0,
            lambda_1=0.01, lambda_2=0.1, ef=False)
    cd_case("cd_grbm_pcd2", O.GRBM, 13, 7, 5, k=2, n_steps=3, seed=34, lr=0.005, momentum=0.0, lambda_1=0.01,
            lambda_2=0.1, pcd=True)
    cd_case("cd_grbm_tail", O.GRBM, 13, 7, 5, k=1, n_steps=3, seed=35, lr=0.005, momentum=0.0, lambda_1=0.01,
            lambda_2=0.1, tail=2)
    cd_case("cd_rbm_mid", O.RBM, 96, 40, 20, k=1, n_steps=3, seed=41, lr=0.1, momentum=0.6, weightcost=0.0002)
    cd_case("cd_grbm_mid", O.GRBM, 200, 24, 10, k=2, n_steps=3, seed=42, lr=0.005, momentum=0.0, lambda_1=0.01,
            lambda_2=0.1, pcd=True)

    dbn_case("dbn_gauss", n_ins=12, sizes=[8, 4], N=23, n_val=4, B=5, k=1, epochs=[40, 30], lrs=[0.005, 0.1],
             lambda_1=0.01, lambda_2=0.1, gauss=True, seed=51, shuffle_seed=777)
    dbn_case("dbn_bern", n_ins=12, sizes=[8, 5, 3], N=20, n_val=0, B=5, k=2, epochs=[30, 30, 30],
             lrs=[0.1, 0.1, 0.1], lambda_1=0.0, lambda_2=0.1, gauss=False, seed=52, shuffle_seed=778)
    minibatch_case()
    mdbn_case("mdbn_small")
    training_case("train_rbm_pcd", O.RBM, 24, 10, 45, 6, 5, 8, seed=61, shuffle_seed=881, learning_rate=0.1, k=1,
                  initial_momentum=0.5, final_momentum=0.9, weightcost=0.0002)
    training_case("train_rbm_cd2", O.RBM, 24, 10, 43, 6, 5, 7, seed=62, shuffle_seed=882, learning_rate=0.05, k=2,
                  initial_momentum=0.5, final_momentum=0.9, weightcost=0.0002, persistent=False)
    training_case("train_grbm", O.GRBM, 24, 10, 45, 6, 5, 8, seed=63, shuffle_seed=883, learning_rate=0.01, k=1,
                  initial_momentum=0.0, final_momentum=0.5, lambda_1=0.01, lambda_2=0.1, persistent=True)
    io_case()
