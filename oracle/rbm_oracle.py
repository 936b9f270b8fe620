"""CPU oracle: NumPy restatement of the reference's RBM / GRBM CD-k hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (mdbn_b200/) may
import this module; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do.

PARITY PIN STATUS: the reference (glgerard/MDBN) ships no tests, golden vectors
or fixtures for this path and Theano itself (third-party, unpinned, ~0.8.2/0.9
by the notebooks' banners) cannot be installed here, so this oracle is
"parity unpinned" at the binary level.  It IS pinned at the source level:
tests/golden/*.npz are produced by executing the reference's own unmodified
src/rbm.py / src/dbn.py / src/MDBN.py under oracle/theano_shim (see
tests/golden/make_golden.py) and tests/test_oracle_golden.py checks every
function below against them to ~1e-12 in float64.

Every function cites the reference lines (relative to /root/reference/) it
restates.  Quirks reproduced on purpose are listed in SURVEY.md App. C.
"""
from __future__ import annotations

import numpy as np

RBM = 0
GRBM = 1

EPSILON = 0.001  # src/rbm.py:347


# ---------------------------------------------------------------------------
# elementwise helpers (Theano semantics)
# ---------------------------------------------------------------------------
def sigmoid(x):
    x = np.asarray(x)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    ex = np.exp(x[~pos])
    out[~pos] = ex / (1.0 + ex)
    return out


def softplus(x):
    x = np.asarray(x)
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def round_half_away(x):
    """tensor.round default mode (src/rbm.py:428)."""
    x = np.asarray(x)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


# ---------------------------------------------------------------------------
# parameters + optimiser state          src/rbm.py:49-164
# ---------------------------------------------------------------------------
def init_W(numpy_rng, n_visible, n_hidden, dtype=np.float64):
    """U(+-4*sqrt(6/(H+V))) of shape [V,H]   src/rbm.py:100-107, src/dbn.py:155-159"""
    bound = 4.0 * np.sqrt(6.0 / (n_hidden + n_visible))
    return np.asarray(numpy_rng.uniform(low=-bound, high=bound, size=(n_visible, n_hidden)),
                      dtype=dtype)


class Layer:
    """State of one RBM/GRBM: W [V,H], hbias [H], vbias [V] and the three
    momentum 'speeds' (src/rbm.py:138-164).  W/hbias may alias another
    layer's arrays (DBN weight tying, src/dbn.py:193-202)."""

    def __init__(self, n_visible, n_hidden, kind=RBM, W=None, hbias=None, vbias=None,
                 numpy_rng=None, dtype=np.float64, error_free=True):
        self.n_visible, self.n_hidden, self.kind, self.error_free = n_visible, n_hidden, kind, error_free
        self.dtype = np.dtype(dtype)
        if numpy_rng is None:
            numpy_rng = np.random.RandomState(1234)          # src/rbm.py:89
        if W is None:
            W = init_W(numpy_rng, n_visible, n_hidden, dtype)
        self.W = W
        self.hbias = np.zeros(n_hidden, dtype) if hbias is None else hbias
        self.vbias = np.zeros(n_visible, dtype) if vbias is None else vbias
        self.W_speed = np.zeros((n_visible, n_hidden), dtype)
        self.hbias_speed = np.zeros(n_hidden, dtype)
        self.vbias_speed = np.zeros(n_visible, dtype)
        self.bit_i_idx = 0                                    # src/rbm.py:425


# ---------------------------------------------------------------------------
# propagation + sampling                 src/rbm.py:187-256, 647-682
# ---------------------------------------------------------------------------
def propup(L, vis):
    pre = vis @ L.W + L.hbias                                 # :198
    return pre, sigmoid(pre)                                  # :199


def propdown(L, hid):
    pre = hid @ L.W.T + L.vbias                               # :226 / :650
    if L.kind == GRBM:
        return pre, pre                                       # linear mean, slot0 == slot1 (:660)
    return pre, sigmoid(pre)                                  # :227


def sample_h_given_v(L, v, u):
    pre, mean = propup(L, v)
    sample = (u < mean).astype(mean.dtype)                    # :210-212, strict <
    return pre, mean, sample


def sample_v_given_h(L, h, u=None):
    """RBM: u are uniforms.  GRBM: u are N(0,1) draws, used only if not error_free."""
    pre, mean = propdown(L, h)
    if L.kind == GRBM:
        sample = mean if L.error_free else mean + u           # :652-658
        return mean, mean, sample                             # :660
    sample = (u < mean).astype(mean.dtype)                    # :237-239
    return pre, mean, sample


def gibbs_hvh(L, h0, u_v, u_h):
    pre_v, v_mean, v_sample = sample_v_given_h(L, h0, u_v)
    v_in = v_mean if L.kind == GRBM else v_sample             # :669 vs :246
    pre_h, h_mean, h_sample = sample_h_given_v(L, v_in, u_h)
    return pre_v, v_mean, v_sample, pre_h, h_mean, h_sample


def gibbs_vhv(L, v0, u_h, u_v):
    pre_h, h_mean, h_sample = sample_h_given_v(L, v0, u_h)
    h_in = h_mean if L.kind == GRBM else h_sample             # :680 vs :254
    pre_v, v_mean, v_sample = sample_v_given_h(L, h_in, u_v)
    return pre_h, h_mean, h_sample, pre_v, v_mean, v_sample


# ---------------------------------------------------------------------------
# energies + monitors                    src/rbm.py:166-185, 421-482, 684-699
# ---------------------------------------------------------------------------
def free_energy(L, v):
    wx_b = v @ L.W + L.hbias
    hidden_term = softplus(wx_b).sum(axis=1)
    if L.kind == GRBM:
        return -hidden_term + 0.5 * np.square(v - L.vbias).sum(axis=1)   # :684-688
    return -hidden_term - v @ L.vbias                                     # :166-171


def free_energy_gap(L, train, test):
    return free_energy(L, test).mean() - free_energy(L, train).mean()     # :173-180


def reconstruction_cost(L, pre_sigmoid_nv, v0):
    if L.kind == GRBM:
        return np.square(sigmoid(pre_sigmoid_nv) - v0).mean()             # :697 (sigma of linear mean)
    ce = v0 * softplus(-pre_sigmoid_nv) + (1.0 - v0) * softplus(pre_sigmoid_nv)
    return ce.sum(axis=1).mean()                                          # :479-480


def pseudo_likelihood_cost(L, v0, bit_i_idx):
    xi = round_half_away(v0)                                              # :428
    fe_xi = free_energy(L, xi)
    xi_flip = xi.copy()
    xi_flip[:, bit_i_idx] = 1 - xi[:, bit_i_idx]                          # :436
    fe_flip = free_energy(L, xi_flip)
    return -np.mean(L.n_visible * softplus(fe_xi - fe_flip))              # :442


# ---------------------------------------------------------------------------
# shared-uniform buffer layout (defined by this build)
# ---------------------------------------------------------------------------
def u_layout(kind, error_free, B, V, H, k):
    """[(name, offset, shape)] of the flat per-step fp32 random buffer:
    [U_h0 (B*H)] then for s<k: [U_v[s] (B*V) RBM | N_v[s] (B*V) noisy GRBM] [U_h[s] (B*H)]."""
    out, off = [("hpos", 0, (B, H))], B * H
    for s in range(k):
        if kind == RBM or not error_free:
            out.append(("v%d" % s, off, (B, V)))
            off += B * V
        out.append(("h%d" % s, off, (B, H)))
        off += B * H
    return out, off


def u_size(kind, error_free, B, V, H, k):
    return u_layout(kind, error_free, B, V, H, k)[1]


def _views(U, kind, error_free, B, V, H, k):
    lay, n = u_layout(kind, error_free, B, V, H, k)
    assert U.size >= n
    return {name: U[off:off + sh[0] * sh[1]].reshape(sh) for name, off, sh in lay}


# ---------------------------------------------------------------------------
# one CD-k / PCD-k step                  src/rbm.py:258-376 + 392-419
# ---------------------------------------------------------------------------
def cd_step(L, v0, U, lr=0.1, k=1, lambda_1=0.0, lambda_2=0.0, weightcost=0.0,
            batch_size=None, momentum=0.0, persistent=None, W_snap=None, trace=None):
    """In-place update of L (and of `persistent`); returns the monitoring cost.

    batch_size is the NOMINAL batch (divisor of the W statistics, :413) — the
    bias statistics use the true row mean (:416-417).  W_snap is the frozen copy
    of W that the weight-decay term multiplies (:414-415, App. C-2); required when
    weightcost != 0.  `trace`, if a dict, receives every intermediate."""
    dt = L.W.dtype
    B, V, H = v0.shape[0], L.n_visible, L.n_hidden
    u = _views(np.asarray(U), L.kind, L.error_free, B, V, H, k)
    f = lambda x: np.asarray(x, dtype=dt)
    lr, momentum, lambda_1, lambda_2, weightcost = map(dt.type, (lr, momentum, lambda_1, lambda_2, weightcost))

    # positive phase :303
    pre_h0, ph_mean, ph_sample = sample_h_given_v(L, v0, f(u["hpos"]))
    h = ph_sample if persistent is None else persistent       # :308-311
    if trace is not None:
        trace.update(pre_h0=pre_h0, ph_mean=ph_mean, ph_sample=ph_sample, chain=[])
    # k Gibbs steps :328-336
    for s in range(k):
        uv = f(u["v%d" % s]) if ("v%d" % s) in u else None
        h_in = h
        pre_v, nv_mean, nv_sample, pre_h, nh_mean, h = gibbs_hvh(L, h_in, uv, f(u["h%d" % s]))
        if trace is not None:
            trace["chain"].append(dict(h_in=np.array(h_in), pre_v=pre_v, nv_mean=nv_mean,
                                       nv_sample=nv_sample, pre_h=pre_h, nh_mean=nh_mean,
                                       nh_sample=h))
    nh_sample = h

    # statistics :411-417
    if weightcost != 0:
        assert W_snap is not None
        decay = weightcost * W_snap
    else:
        decay = 0
    gW = (v0.T @ ph_mean - nv_mean.T @ nh_mean) / dt.type(batch_size) - decay
    ghb = np.mean(ph_mean - nh_mean, axis=0)
    gvb = np.mean(v0 - nv_mean, axis=0)

    # lambda_1 scaling + multipliers :347-356
    D = 1 + 2 * lr * lambda_1 / (np.abs(L.W) + dt.type(EPSILON))
    gW = gW / D
    mult_W = (1 - 2 * lr * lambda_2) / D

    # monitoring cost uses OLD parameters (Theano evaluates outputs and updates
    # against the pre-update shared values)
    if persistent is not None:
        cost = pseudo_likelihood_cost(L, v0, L.bit_i_idx)     # :367-371
        L.bit_i_idx = (L.bit_i_idx + 1) % V                   # :445
    else:
        cost = reconstruction_cost(L, pre_v, v0)              # :374

    if trace is not None:
        trace.update(gW=gW, ghb=ghb, gvb=gvb, cost=cost)

    # simultaneous update: param uses the OLD speed :358-365 (App. C-1)
    newS_W = gW + (L.W_speed - gW) * momentum
    newS_hb = ghb + (L.hbias_speed - ghb) * momentum
    newS_vb = gvb + (L.vbias_speed - gvb) * momentum
    new_W = L.W * mult_W + L.W_speed * lr
    new_hb = L.hbias + L.hbias_speed * lr
    new_vb = L.vbias + L.vbias_speed * lr
    L.W[...], L.hbias[...], L.vbias[...] = new_W, new_hb, new_vb
    L.W_speed[...], L.hbias_speed[...], L.vbias_speed[...] = newS_W, newS_hb, newS_vb
    if persistent is not None:
        persistent[...] = nh_sample                            # :369
    return dt.type(cost)


# ---------------------------------------------------------------------------
# the same step split at the data-parallel cut: raw row sums -> packed buffer ->
# update.  cd_apply(cd_stats(...)) == cd_step(...) up to summation order.
# ---------------------------------------------------------------------------
def cd_stats(L, v0, U, k=1, persistent=None):
    """[v0^T ph - nv^T nh (V*H) | sum(ph-nh) (H) | sum(v0-nv) (V) | cost numerator | rows] for these rows;
    updates `persistent` (rows of the chain owned by the caller) but not the parameters."""
    dt = L.W.dtype
    B, V, H = v0.shape[0], L.n_visible, L.n_hidden
    u = _views(np.asarray(U), L.kind, L.error_free, B, V, H, k)
    f = lambda x: np.asarray(x, dtype=dt)
    pre_h0, ph_mean, ph_sample = sample_h_given_v(L, v0, f(u["hpos"]))
    h = ph_sample if persistent is None else persistent
    for s in range(k):
        uv = f(u["v%d" % s]) if ("v%d" % s) in u else None
        pre_v, nv_mean, nv_sample, pre_h, nh_mean, h = gibbs_hvh(L, h, uv, f(u["h%d" % s]))
    if persistent is not None:
        xi = round_half_away(v0)
        fe = free_energy(L, xi)
        xf = xi.copy()
        xf[:, L.bit_i_idx] = 1 - xi[:, L.bit_i_idx]
        num = -np.sum(V * softplus(fe - free_energy(L, xf)))
        persistent[...] = h
    elif L.kind == GRBM:
        num = np.square(sigmoid(pre_v) - v0).sum()
    else:
        num = (v0 * softplus(-pre_v) + (1.0 - v0) * softplus(pre_v)).sum()
    return np.concatenate([(v0.T @ ph_mean - nv_mean.T @ nh_mean).ravel(), (ph_mean - nh_mean).sum(0),
                           (v0 - nv_mean).sum(0), [num, B]]).astype(dt)


def cd_apply(L, packed, lr, lambda_1=0.0, lambda_2=0.0, weightcost=0.0, batch_size=None, momentum=0.0,
             W_snap=None, pcd=False):
    dt = L.W.dtype
    V, H = L.n_visible, L.n_hidden
    rows = float(packed[-1])
    lr, momentum, lambda_1, lambda_2, weightcost = map(dt.type, (lr, momentum, lambda_1, lambda_2, weightcost))
    gW = packed[:V * H].reshape(V, H) / dt.type(batch_size) - (weightcost * W_snap if weightcost != 0 else 0)
    ghb = packed[V * H:V * H + H] / dt.type(rows)
    gvb = packed[V * H + H:V * H + H + V] / dt.type(rows)
    D = 1 + 2 * lr * lambda_1 / (np.abs(L.W) + dt.type(EPSILON))
    gW = gW / D
    mult_W = (1 - 2 * lr * lambda_2) / D
    den = rows * V if (not pcd and L.kind == GRBM) else rows
    cost = packed[-2] / den
    new = (L.W * mult_W + L.W_speed * lr, L.hbias + L.hbias_speed * lr, L.vbias + L.vbias_speed * lr,
           gW + (L.W_speed - gW) * momentum, ghb + (L.hbias_speed - ghb) * momentum,
           gvb + (L.vbias_speed - gvb) * momentum)
    L.W[...], L.hbias[...], L.vbias[...], L.W_speed[...], L.hbias_speed[...], L.vbias_speed[...] = new
    if pcd:
        L.bit_i_idx = (L.bit_i_idx + 1) % V
    return dt.type(cost)


# ---------------------------------------------------------------------------
# batching                               src/utils.py:54-75
# ---------------------------------------------------------------------------
def get_minibatches_idx(n, batch_size, shuffle=False, rng=None):
    """rng=None -> the global numpy RNG exactly as the reference (App. C-10)."""
    idx_list = np.arange(n, dtype="int32")
    if shuffle:
        (np.random if rng is None else rng).shuffle(idx_list)
    minibatches, start = [], 0
    for _ in range(n // batch_size):
        minibatches.append(idx_list[start:start + batch_size])
        start += batch_size
    if start != n:
        minibatches.append(idx_list[start:])                   # ragged tail
    return range(len(minibatches)), minibatches


# ---------------------------------------------------------------------------
# standalone epoch loop                  src/rbm.py:484-629 (RBM.training / learn_model), :701-728 (GRBM.training)
# ---------------------------------------------------------------------------
def rbm_training(L, train, val, training_epochs, batch_size=10, learning_rate=None, k=1, initial_momentum=0.0,
                 final_momentum=0.0, weightcost=0.0, lambda_1=0.0, lambda_2=None, persistent=None, u_provider=None):
    """RBM.training (:484-520): PCD with a chain of zeros by default (`persistent=True`), `lambda_2` accepted and NOT
    forwarded (:504-509).  GRBM.training (:701-728): always CD (`persistent` accepted and not forwarded, :711-717),
    lambda_1 / lambda_2 forwarded, defaults lr 0.01 / lambda_2 0.1.  Then learn_model (:522-629): momentum switches
    to `final_momentum` at 0-based epoch 6 (:584), shuffled minibatches from the GLOBAL numpy RNG (:587-589),
    per epoch the mean cost and the free-energy gap of train[0:n_val] vs val (:597-600).
    u_provider(call_index, B) -> the shared random buffer of that step.  Returns [(mean cost, gap)] per epoch."""
    if L.kind == GRBM:
        lr = 0.01 if learning_rate is None else learning_rate
        l1, l2 = lambda_1, (0.1 if lambda_2 is None else lambda_2)
        chain = None
    else:
        lr = 0.1 if learning_rate is None else learning_rate
        l1, l2 = 0.0, 0.0
        use_pcd = True if persistent is None else bool(persistent)
        chain = np.zeros((batch_size, L.n_hidden), dtype=L.W.dtype) if use_pcd else None
    W_snap = L.W.copy()                                        # :414-415, captured by get_cost_updates
    n, n_val = train.shape[0], val.shape[0]
    momentum, history, call = initial_momentum, [], 0
    for epoch in range(training_epochs):
        if epoch == 6:                                         # :584
            momentum = final_momentum
        _, minibatches = get_minibatches_idx(n, batch_size, shuffle=True)
        costs = []
        for mb in minibatches:
            U = u_provider(call, len(mb))
            costs.append(cd_step(L, train[mb], U, lr=lr, k=k, lambda_1=l1, lambda_2=l2, weightcost=weightcost,
                                 batch_size=batch_size, momentum=momentum, persistent=chain, W_snap=W_snap))
            call += 1
        history.append((float(np.mean(costs)), float(free_energy_gap(L, train[:n_val], val))))
    return history


# ---------------------------------------------------------------------------
# DBN stack + greedy loop                src/dbn.py:64-204, 238-332, 334-517
# ---------------------------------------------------------------------------
class DBN:
    def __init__(self, numpy_rng=None, n_ins=784, gauss=True, hidden_layers_sizes=(400,),
                 n_outs=40, W_list=None, b_list=None, dtype=np.float64):
        self.n_ins = n_ins
        self.sizes = list(hidden_layers_sizes) + [n_outs]      # :105
        self.n_layers = len(self.sizes)
        self.dtype = np.dtype(dtype)
        if numpy_rng is None:
            numpy_rng = np.random.RandomState(123)             # :111
        self.theano_seed = numpy_rng.randint(2 ** 30)          # :114 (consumes one draw)
        self.layers = []
        for i in range(self.n_layers):
            n_in = n_ins if i == 0 else self.sizes[i - 1]
            n_out = self.sizes[i]
            W = init_W(numpy_rng, n_in, n_out, dtype) if W_list is None else np.array(W_list[i], dtype)
            b = np.zeros(n_out, dtype) if b_list is None else np.array(b_list[i], dtype)
            kind = GRBM if (i == 0 and gauss) else RBM          # :187
            self.layers.append(Layer(n_in, n_out, kind, W=W, hbias=b, dtype=dtype))

    def get_output(self, x, layer=-1):                          # :214-236
        n = self.n_layers if layer == -1 else layer + 1
        for L in self.layers[:n]:
            x = sigmoid(x @ L.W + L.hbias)                      # src/mlp.py:103-107
        return x

    def layer_hyper(self, i, lambda_1, lambda_2):               # :284-294
        if self.layers[i].kind == GRBM:
            return dict(lambda_1=lambda_1, lambda_2=lambda_2, weightcost=0.0)
        return dict(lambda_1=0.0, lambda_2=0.0, weightcost=0.0002)

    def training(self, train_x, batch_size, k, pretraining_epochs, pretrain_lr,
                 lambda_1=0.0, lambda_2=0.1, validation_x=None,
                 u_provider=None, shuffle_rng=None, log=None):
        """u_provider(layer, call_idx, B) -> flat U buffer for that step.
        Returns per-layer list of (iter, cost) at validation points."""
        assert batch_size > 1                                   # :276
        n_data = train_x.shape[0]
        snaps = [L.W.copy() for L in self.layers]               # W_snap at training_functions() :415
        idx_mb, _ = get_minibatches_idx(n_data, batch_size, True, shuffle_rng)   # :420
        n_train_batches = idx_mb[-1] + 1
        history = []
        for i, L in enumerate(self.layers):
            hyp = self.layer_hyper(i, lambda_1, lambda_2)
            momentum = 0.0 if L.kind == GRBM else 0.6           # :430-433
            best_cost, epoch, done, calls = np.inf, 0, False, 0
            patience = pretraining_epochs[i]                    # :440
            vf = min(20 * n_train_batches, patience // 2)       # :441
            hist = []
            while epoch < pretraining_epochs[i] and not done:   # :444
                epoch += 1
                _, minibatches = get_minibatches_idx(n_data, batch_size, True, shuffle_rng)
                if L.kind != GRBM and epoch == 6:
                    momentum = 0.9                              # :452-453
                for mb, minibatch in enumerate(minibatches):
                    x = train_x[minibatch]
                    v0 = self.get_output(x, i - 1) if i > 0 else x
                    U = u_provider(i, calls, v0.shape[0])
                    cost = cd_step(L, v0, U, lr=pretrain_lr[i], k=k, batch_size=batch_size,
                                   momentum=momentum, W_snap=snaps[i], **hyp)
                    calls += 1
                    it = (epoch - 1) * n_train_batches + mb     # :460
                    if (it + 1) % vf == 0:                      # :462
                        feg = None
                        if cost < best_cost:                    # :477
                            if cost < best_cost * 0.995:
                                patience = max(patience, it * 2)  # :483
                            best_cost = cost
                            if validation_x is not None:        # :488-501
                                nv = validation_x.shape[0]
                                tin = train_x if i == 0 else self.get_output(train_x[np.arange(nv)], i - 1)
                                vin = validation_x if i == 0 else self.get_output(validation_x, i - 1)
                                feg = free_energy(L, vin).mean() - free_energy(L, tin).mean()
                        hist.append((it, float(cost), feg))
                    if patience <= it:                          # :506-508
                        done = True
                        break
            history.append(dict(epochs=epoch, calls=calls, validations=hist))
        return history
