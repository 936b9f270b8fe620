"""Minimal eager NumPy stand-in for the handful of Theano features that the
reference's hot path (src/rbm.py, src/dbn.py, src/mlp.py, src/MDBN.py,
src/utils.py) touches.  TEST INFRASTRUCTURE ONLY.

Why it exists: Theano cannot be installed here (no network, CPython 3.12), so
the reference cannot be run as-is.  This shim lets the reference's *own,
unmodified source files* be imported from /root/reference/src and executed, so
that the golden vectors under tests/golden/ come from the reference's algebra
(its expression graphs, its update dictionaries, its training loops) rather
than from a hand restatement.  What the shim supplies is only Theano's
*semantics* for the ops used there:

  * lazy expression graph, evaluated by ``theano.function``;
  * ``updates`` are simultaneous: every new value is computed from the OLD
    values of all shared variables, then committed together;
  * ``givens`` substitute graph nodes by identity;
  * a NumPy array multiplied into a graph becomes a constant captured at
    graph-build time (copy -> "snapshot"; see ``config.constant_alias``);
  * ``scan(fn, outputs_info=[None..., init], n_steps=k)`` is a sequential loop
    returning the stacked per-step outputs (the reference indexes ``[-1]``);
  * ``MRG_RandomStreams.binomial(n=1, p) == (uniform < p)`` (strict ``<``) and
    ``normal == avg + std * n``; the uniforms/normals come from an injected
    provider (the shared-uniform-buffer mode of SURVEY.md App. A) — bit-exact
    MRG31k3p streams are NOT reproduced and not needed for parity;
  * ``sigmoid``/``softplus`` in their numerically stable forms, and
    ``binary_crossentropy(sigmoid(x), t)`` evaluated in the softplus form that
    Theano's optimiser rewrites it to.

Nothing in the product package imports this.  Use ``install()`` to inject the
fake ``theano`` (+ matplotlib / scipy.misc stubs) into ``sys.modules`` and to
put the reference ``src`` directory on ``sys.path``.
"""
from __future__ import annotations

import sys
import types
from collections import OrderedDict

import numpy as np


# --------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------
class _Config:
    floatX = "float64"
    mode = "FAST_RUN"
    # False: ndarray constants are copied when they enter a graph (snapshot,
    # what the GPU backend the author used necessarily does: get_value() is a
    # device->host copy).  True: alias the live array (possible on Theano's
    # CPU backend with borrow=True).  SURVEY.md App. C-2.
    constant_alias = False


config = _Config()

# provider(ordinal, kind, shape, fn, call_idx) -> ndarray ; kind in {"uniform","normal"}
_rng_provider = None


def set_rng_provider(fn):
    global _rng_provider
    _rng_provider = fn


# --------------------------------------------------------------------------
# graph
# --------------------------------------------------------------------------
class Variable:
    __array_ufunc__ = None  # ndarray <op> Variable defers to Variable.__r<op>__
    __array_priority__ = 1000

    def __init__(self, op, inputs=(), name=None, payload=None):
        self.op = op
        self.inputs = tuple(inputs)
        self.name = name
        self.payload = payload

    # Theano: variables are truthy unless they come from a comparison
    def __bool__(self):
        return True

    __hash__ = object.__hash__

    def __repr__(self):
        return "<%s %s>" % (self.op, self.name or hex(id(self)))

    # arithmetic -----------------------------------------------------------
    def __add__(self, o): return _apply("add", self, o)
    def __radd__(self, o): return _apply("add", o, self)
    def __sub__(self, o): return _apply("sub", self, o)
    def __rsub__(self, o): return _apply("sub", o, self)
    def __mul__(self, o): return _apply("mul", self, o)
    def __rmul__(self, o): return _apply("mul", o, self)
    def __truediv__(self, o): return _apply("div", self, o)
    def __rtruediv__(self, o): return _apply("div", o, self)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__
    def __mod__(self, o): return _apply("mod", self, o)
    def __neg__(self): return _apply("neg", self)
    def __pow__(self, o): return _apply("pow", self, o)

    @property
    def T(self): return _apply("transpose", self)

    @property
    def shape(self): return _apply("shape", self)

    def sum(self, axis=None): return _apply("sum", self, payload=axis)
    def mean(self, axis=None): return _apply("mean", self, payload=axis)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        dyn = [i for i in idx if isinstance(i, Variable)]
        return Variable("subtensor", (self, *dyn), payload=idx)


class SharedVariable(Variable):
    def __init__(self, value, name=None, borrow=False):
        super().__init__("shared", (), name=name)
        value = np.asarray(value)
        self.value = value if borrow else value.copy()

    def get_value(self, borrow=False):
        return self.value if borrow else self.value.copy()

    def set_value(self, v, borrow=False):
        v = np.asarray(v)
        self.value = v if borrow else v.copy()


def shared(value=None, name=None, borrow=False, **_):
    return SharedVariable(value, name=name, borrow=borrow)


def _as_var(x):
    if isinstance(x, Variable):
        return x
    if isinstance(x, np.ndarray):
        data = x if config.constant_alias else x.copy()
        return Variable("constant", (), payload=data)
    return Variable("constant", (), payload=x)  # python scalar stays weakly typed


def _apply(op, *inputs, payload=None):
    return Variable(op, tuple(_as_var(i) for i in inputs), payload=payload)


def _sigmoid(x):
    x = np.asarray(x)
    out = np.empty_like(x, dtype=np.result_type(x, np.float32))
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    ex = np.exp(x[~pos])
    out[~pos] = ex / (1.0 + ex)
    return out


def _softplus(x):
    x = np.asarray(x)
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def _round_half_away(x):
    # Theano's tensor.round default mode is "half_away_from_zero"
    x = np.asarray(x)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


class _Evaluator:
    def __init__(self, givens, inputs, fn, call_idx):
        self.givens = givens
        self.memo = dict(inputs)
        self.fn = fn
        self.call_idx = call_idx

    def __call__(self, v):
        v = self.givens.get(v, v)
        k = id(v)
        if k in self.memo:
            return self.memo[k]
        r = self._eval(v)
        self.memo[k] = r
        return r

    def _eval(self, v):
        op = v.op
        if op == "shared":
            return v.value
        if op == "constant":
            return v.payload
        if op == "input":
            raise KeyError("missing value for input %r" % (v.name,))
        a = [self(i) for i in v.inputs] if op not in ("subtensor", "set_subtensor") else None
        if op == "add": return a[0] + a[1]
        if op == "sub": return a[0] - a[1]
        if op == "mul": return a[0] * a[1]
        if op == "div": return a[0] / a[1]
        if op == "mod": return a[0] % a[1]
        if op == "pow": return a[0] ** a[1]
        if op == "neg": return -a[0]
        if op == "transpose": return np.asarray(a[0]).T
        if op == "shape": return np.asarray(a[0]).shape
        if op == "dot": return np.dot(a[0], a[1])
        if op == "abs": return np.abs(a[0])
        if op == "sqr": return np.square(a[0])
        if op == "log": return np.log(a[0])
        if op == "tanh": return np.tanh(a[0])
        if op == "round": return _round_half_away(a[0])
        if op == "sigmoid": return _sigmoid(a[0])
        if op == "softplus": return _softplus(a[0])
        if op == "sum": return np.sum(a[0], axis=v.payload)
        if op == "mean": return np.mean(a[0], axis=v.payload)
        if op == "cast": return np.asarray(a[0]).astype(v.payload)
        if op == "bce":
            # binary_crossentropy(sigmoid(x), t) after Theano's
            # log(sigmoid(x)) -> -softplus(-x), log(1-sigmoid(x)) -> -softplus(x)
            out_v, tgt = v.inputs
            t = self(tgt)
            out_v = self.givens.get(out_v, out_v)
            if out_v.op == "sigmoid":
                x = self(out_v.inputs[0])
                return t * _softplus(-x) + (1.0 - t) * _softplus(x)
            o = self(out_v)
            return -(t * np.log(o) + (1.0 - t) * np.log(1.0 - o))
        if op == "subtensor":
            base = self(v.inputs[0])
            idx = tuple(self(i) if isinstance(i, Variable) else i for i in v.payload)
            idx = tuple(np.asarray(i) if isinstance(i, (list, range)) else i for i in idx)
            return np.asarray(base)[idx if len(idx) > 1 else idx[0]]
        if op == "set_subtensor":
            sub, val = v.inputs
            base = np.array(self(sub.inputs[0]), copy=True)
            idx = tuple(self(i) if isinstance(i, Variable) else i for i in sub.payload)
            base[idx if len(idx) > 1 else idx[0]] = self(val)
            return base
        if op == "rng_binomial":
            p = np.asarray(a[1])
            u = _rng_provider(v.payload["ordinal"], "uniform", p.shape, self.fn, self.call_idx)
            return (np.asarray(u) < p).astype(v.payload["dtype"])
        if op == "rng_normal":
            shape = tuple(int(s) for s in a[0])
            n = _rng_provider(v.payload["ordinal"], "normal", shape, self.fn, self.call_idx)
            return (v.payload["avg"] + v.payload["std"] * np.asarray(n)).astype(v.payload["dtype"])
        raise NotImplementedError(op)


class In:
    def __init__(self, variable, **_):
        self.variable = variable


class Function:
    def __init__(self, inputs, outputs=None, updates=None, givens=None, name=None, mode=None, **_):
        self.inputs = [i.variable if isinstance(i, In) else i for i in inputs]
        self.outputs = outputs
        self.updates = list(updates.items()) if updates is not None and hasattr(updates, "items") \
            else list(updates or [])
        self.givens = {k: _as_var(v) for k, v in dict(givens or {}).items()}
        self.name = name
        self.n_calls = 0

    def __call__(self, *args, **kwargs):
        vals = {}
        for var, a in zip(self.inputs, args):
            vals[id(var)] = self._coerce(var, a)
        for k, a in kwargs.items():
            var = next(v for v in self.inputs if v.name == k)
            vals[id(var)] = self._coerce(var, a)
        ev = _Evaluator(self.givens, vals, self, self.n_calls)
        self.n_calls += 1
        if self.outputs is None:
            out = None
        elif isinstance(self.outputs, (list, tuple)):
            out = [np.asarray(ev(o)) for o in self.outputs]
        else:
            out = np.asarray(ev(self.outputs))
        # simultaneous updates: evaluate everything against the old state first
        new_vals = [(sv, ev(_as_var(expr))) for sv, expr in self.updates]
        for sv, nv in new_vals:
            sv.value = np.array(nv, dtype=sv.value.dtype, copy=True)
        return out

    @staticmethod
    def _coerce(var, a):
        dt = var.payload
        return np.asarray(a, dtype=dt) if dt is not None else np.asarray(a)


def function(inputs, outputs=None, updates=None, givens=None, name=None, mode=None, **kw):
    return Function(inputs, outputs, updates, givens, name, mode, **kw)


def scan(fn, outputs_info=None, n_steps=None, name=None, **_):
    """Sequential loop; returns (list of per-output step lists, updates)."""
    carry = [o for o in outputs_info if o is not None]
    n_out = len(outputs_info)
    steps = [[] for _ in range(n_out)]
    for _s in range(int(n_steps)):
        outs = fn(*carry)
        if not isinstance(outs, (list, tuple)):
            outs = [outs]
        for j, o in enumerate(outs):
            steps[j].append(o)
        carry = [outs[j] for j in range(n_out) if outputs_info[j] is not None]
    return steps, OrderedDict()


# --------------------------------------------------------------------------
# module tree
# --------------------------------------------------------------------------
def _input(dtype_default):
    def make(name=None, dtype=None):
        dt = dtype or dtype_default()
        return Variable("input", (), name=name, payload=dt)
    return make


def _build_modules():
    theano = types.ModuleType("theano")
    tensor = types.ModuleType("theano.tensor")
    nnet = types.ModuleType("theano.tensor.nnet")
    compile_ = types.ModuleType("theano.compile")
    nanguard = types.ModuleType("theano.compile.nanguardmode")
    sandbox = types.ModuleType("theano.sandbox")
    rng_mrg = types.ModuleType("theano.sandbox.rng_mrg")
    srs = types.ModuleType("theano.tensor.shared_randomstreams")

    theano.config = config
    theano.shared = shared
    theano.function = function
    theano.scan = scan
    theano.In = In
    theano.tensor = tensor
    theano.compile = compile_
    theano.sandbox = sandbox
    theano.__shim__ = True

    fx = lambda: config.floatX
    tensor.matrix = _input(fx)
    tensor.vector = _input(fx)
    tensor.scalar = _input(fx)
    tensor.dmatrix = _input(lambda: "float64")
    tensor.lvector = _input(lambda: "int64")
    tensor.ivector = _input(lambda: "int32")
    tensor.dot = lambda a, b: _apply("dot", a, b)
    tensor.cast = lambda x, dtype: _apply("cast", x, payload=dtype)
    tensor.abs_ = lambda x: _apply("abs", x)
    tensor.sqr = lambda x: _apply("sqr", x)
    tensor.log = lambda x: _apply("log", x)
    tensor.tanh = lambda x: _apply("tanh", x)
    tensor.round = lambda x: _apply("round", x)
    tensor.mean = lambda x, axis=None: _apply("mean", x, payload=axis)
    tensor.sum = lambda x, axis=None: _apply("sum", x, payload=axis)
    tensor.set_subtensor = lambda sub, val: _apply("set_subtensor", sub, val)

    def grad(*a, **k):
        raise NotImplementedError("tensor.grad is outside the shim (symbolic_grad is never enabled)")
    tensor.grad = grad
    tensor.nnet = nnet
    tensor.shared_randomstreams = srs

    nnet.sigmoid = lambda x: _apply("sigmoid", x)
    nnet.softplus = lambda x: _apply("softplus", x)
    nnet.binary_crossentropy = lambda o, t: _apply("bce", o, t)

    class RandomStreams:
        def __init__(self, seed=None):
            self.seed = seed
            self.n_nodes = 0

        def _ordinal(self):
            n = self.n_nodes
            self.n_nodes += 1
            return n

        def binomial(self, size=None, n=1, p=None, dtype="int64", **_):
            assert n == 1
            return _apply("rng_binomial", size, p,
                          payload={"ordinal": self._ordinal(), "dtype": dtype})

        def normal(self, size=None, avg=0.0, std=1.0, dtype=None, **_):
            return _apply("rng_normal", size,
                          payload={"ordinal": self._ordinal(), "avg": avg, "std": std,
                                   "dtype": dtype or config.floatX})

    rng_mrg.MRG_RandomStreams = RandomStreams
    srs.RandomStreams = RandomStreams
    sandbox.rng_mrg = rng_mrg

    class _Mode:
        def __init__(self, *a, **k):
            pass
    nanguard.NanGuardMode = _Mode
    compile_.nanguardmode = nanguard
    compile_.MonitorMode = _Mode

    mods = {
        "theano": theano, "theano.tensor": tensor, "theano.tensor.nnet": nnet,
        "theano.compile": compile_, "theano.compile.nanguardmode": nanguard,
        "theano.sandbox": sandbox, "theano.sandbox.rng_mrg": rng_mrg,
        "theano.tensor.shared_randomstreams": srs,
    }
    return mods


def _stub_plotting():
    mods = {}
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            for n in ("figure", "ion", "clf", "subplot", "imshow", "draw", "pause", "close",
                      "axis", "title"):
                setattr(plt, n, lambda *a, **k: None)
            mpl.pyplot = plt
            mods["matplotlib"] = mpl
            mods["matplotlib.pyplot"] = plt
    import scipy
    if not hasattr(scipy, "misc") or not hasattr(getattr(scipy, "misc", None), "imsave"):
        misc = types.ModuleType("scipy.misc")
        misc.imsave = lambda *a, **k: None
        mods["scipy.misc"] = misc
        scipy.misc = misc
    return mods


def install(reference_src="/root/reference/src", floatX="float64"):
    """Inject the shim as ``theano`` and make the reference importable."""
    config.floatX = floatX
    for k, m in {**_build_modules(), **_stub_plotting()}.items():
        sys.modules[k] = m
    if reference_src not in sys.path:
        sys.path.insert(0, reference_src)
    return sys.modules["theano"]
