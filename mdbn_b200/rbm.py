"""RBM / GRBM with the reference's class surface (src/rbm.py:46-728), backed by the
sm_100a kernels behind include/mdbn_b200.h.  Eager: where the reference returns
symbolic Theano expressions, these methods take arrays and return device tensors.

Arrays in:  numpy arrays, torch tensors or `Shared`; out: fp32 CUDA tensors
(`.cpu().numpy()` for host values).  Parameters live in `Shared` objects
(get_value / set_value like theano.shared) and are updated IN PLACE by training.
"""
from __future__ import print_function, division

import ctypes
import timeit

import numpy
import torch

from . import _lib
from .rng import RandomStreams, BufferStreams, ExplicitBuffer
from .utils import Shared, as_device_matrix, default_device, get_minibatches_idx, IndexFeeder

_PAD = 8   # W rows padded to 8 floats: 16-byte aligned bulk copies + whole 8-column MMA tiles


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream on the current device (the raw accessor: torch.cuda.current_stream()
    builds a Stream object, ~13 us per call on the launch path)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class CDUpdates(dict):
    """What get_cost_updates returns in place of Theano's `updates` dictionary: the
    hyper-parameters of one CD-k / PCD-k step plus the state it will overwrite
    (keys = the Shared objects, like the reference's shared-variable keys)."""

    def __init__(self, rbm, **hyper):
        super().__init__()
        self.rbm = rbm
        self.hyper = hyper
        for s in rbm.params + rbm.params_speed:
            self[s] = "in-place"
        if hyper.get("persistent") is not None:
            self[hyper["persistent"]] = "in-place"


class Cost:
    """Handle for the monitoring cost of a step (src/rbm.py:367-374)."""

    def __init__(self, updates):
        self.updates = updates
        self.kind = "pseudo_likelihood" if updates.hyper.get("persistent") is not None else "reconstruction"


class TrainFn:
    """The compiled `train_rbm(indexes, momentum)` of src/rbm.py:533-544 /
    `fn(indexes=, momentum=, lr=)` of src/dbn.py:302-312: one mdbn_cd_step per call."""

    def __init__(self, rbm, updates, dataset, layer_id=0, input_fn=None, path="auto", tf32=False):
        self.rbm, self.updates, self.layer_id = rbm, updates, layer_id
        self.device = rbm.device
        self.path, self.tf32 = path, tf32
        self._dataset_arg, self._input_fn = dataset, input_fn
        self._data, self._data_key = None, None
        self.cost_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._cost_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self._keep = None
        self.n_calls = 0
        self.sync = True            # return a Python float (like the reference); False -> device scalar
        self.dp = None              # parallel.DataParallel: shard the minibatch rows + all-reduce the statistics
        self._stats = None

    def data(self):
        """Device-resident input matrix of this layer.  For a stacked layer it is the
        deterministic up-pass of the frozen layers below (src/dbn.py:146), cached and
        recomputed only when their parameters change."""
        key = self._input_fn.version() if self._input_fn is not None else 0
        if self._data is None or key != self._data_key:
            x = as_device_matrix(self._dataset_arg, self.device)
            self._data = self._input_fn(x) if self._input_fn is not None else x
            self._data_key = key
        return self._data

    def __call__(self, indexes=None, momentum=0.0, lr=None, rows=None):
        if self.dp is not None and getattr(self.dp, "comm", None) is not None:
            # data-parallel step inside the library: this rank's rows + the communicator, one mdbn_cd_step
            from .parallel import shard_rows
            mine = shard_rows(indexes, self.dp.rank, self.dp.world)
            return self._call(mine, momentum, lr, rows_total=len(indexes), comm=self.dp.comm)
        if self.dp is not None and self.dp.world > 1:
            return self._call_dp(indexes, momentum, lr)
        return self._call(indexes, momentum, lr)

    def _call_dp(self, indexes, momentum, lr):
        """Large-batch CD split over ranks: STATS on this rank's rows -> all-reduce -> APPLY."""
        from .parallel import stats_size
        r = self.rbm
        if self._stats is None:
            self._stats = torch.zeros(stats_size(r.n_visible, r.n_hidden), dtype=torch.float32, device=self.device)
        n_total = len(indexes)

        def stats_fn(mine):
            self._call(mine, momentum, lr, phase=_lib.PHASE_STATS)
            return self._stats

        def apply_fn(buf, rows_total):
            return self._call(None, momentum, lr, phase=_lib.PHASE_APPLY, rows_total=rows_total)
        return self.dp.step(indexes, stats_fn, apply_fn)

    def step_from_host(self, host_batch, momentum=0.0, lr=None, next_host_batch=None, lag=0):
        """One step on a minibatch that lives in (pinned) HOST memory — the streaming form of the train
        function for datasets kept off the device.  Two device staging buffers and a copy stream: the
        host->device copy of `next_host_batch` is enqueued right after this step's kernel and overlaps it.
        lag=0: the cost of THIS step is read back and returned (same contract as __call__ with sync=True).
        lag=1: every step's cost is still copied to the host, but the call returns the cost of the PREVIOUS
        step (None on the first call; `flush()` returns the last one) so that the next step is enqueued
        before the host blocks — an epoch loop that only averages the costs (src/dbn.py:343-353) is unchanged."""
        B, V = int(host_batch.shape[0]), int(host_batch.shape[1])
        st = getattr(self, "_feed", None)
        if st is None or st["shape"] != (B, V):
            main = torch.cuda.current_stream()       # looked up once: torch.cuda.current_stream() costs ~13 us
            copy = torch.cuda.Stream(device=self.device)
            st = self._feed = {
                "shape": (B, V), "cur": 0, "staged": [None, None], "nbytes": B * V * 4,
                "buf": [torch.empty((B, V), dtype=torch.float32, device=self.device) for _ in range(2)],
                "ready": [torch.cuda.Event(), torch.cuda.Event()], "free": [torch.cuda.Event(), torch.cuda.Event()],
                "copy": copy, "copy_h": ctypes.c_void_p(copy.cuda_stream),
                "main": main, "main_h": ctypes.c_void_p(main.cuda_stream),
                "rows": torch.arange(B, dtype=torch.int32, device=self.device)}
        lib = self.rbm.ctx.lib
        main, copy = st["main"], st["copy"]

        def stage(slot, batch):
            if batch.dtype != torch.float32 or not batch.is_contiguous():
                raise TypeError("step_from_host takes contiguous float32 [B, V] host tensors (pinned for overlap)")
            if st["staged"][slot] is not None:
                copy.wait_event(st["free"][slot])
            _lib.check(lib.mdbn_copy_async(st["buf"][slot].data_ptr(), batch.data_ptr(), st["nbytes"], st["copy_h"]))
            st["ready"][slot].record(copy)
            st["staged"][slot] = batch
        cur = st["cur"]
        if st["staged"][cur] is not host_batch:                  # not prefetched by the previous call
            stage(cur, host_batch)
        main.wait_event(st["ready"][cur])
        # the cost of step n goes to slot n%2 of a small device ring and is read back on a THIRD stream: a
        # 4-byte D2H copy in the main stream would sit between two kernels (~10 us per step)
        if "lag" not in st:
            d2h = torch.cuda.Stream(device=self.device)
            st["lag"] = {"host": [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)],
                         "dev": torch.zeros(2, dtype=torch.float32, device=self.device),
                         "ev": [torch.cuda.Event(), torch.cuda.Event()], "n": 0, "pending": None,
                         "d2h": d2h, "d2h_h": ctypes.c_void_p(d2h.cuda_stream)}
        lg = st["lag"]
        k = lg["n"] & 1
        if lg["n"] >= 2:
            main.wait_event(lg["ev"][k])         # slot k was read back two steps ago (long done)
        sync, self.sync = self.sync, False
        try:
            self._call(st["rows"], momentum, lr, data_override=st["buf"][cur], stream=st["main_h"],
                       costs=lg["dev"][k:k + 1])
        finally:
            self.sync = sync
        st["free"][cur].record(main)
        nxt = 1 - cur
        if next_host_batch is not None:
            stage(nxt, next_host_batch)
        else:
            st["staged"][nxt] = None
        st["cur"] = nxt
        lg["d2h"].wait_event(st["free"][cur])
        _lib.check(lib.mdbn_copy_async(lg["host"][k].data_ptr(), lg["dev"].data_ptr() + 4 * k, 4, lg["d2h_h"]))
        lg["ev"][k].record(lg["d2h"])
        prev, lg["n"] = lg["pending"], lg["n"] + 1
        if not lag:
            lg["pending"] = None
            lg["ev"][k].synchronize()
            return float(lg["host"][k][0])
        lg["pending"] = k
        if prev is None:
            return None
        lg["ev"][prev].synchronize()
        return float(lg["host"][prev][0])

    def flush(self):
        """Cost of the last lagged step (step_from_host(..., lag=1)); None when nothing is pending."""
        lg = getattr(self, "_feed", {}).get("lag") if getattr(self, "_feed", None) else None
        if not lg or lg["pending"] is None:
            return None
        k, lg["pending"] = lg["pending"], None
        lg["ev"][k].synchronize()
        return float(lg["host"][k][0])

    def run_steps_from_host(self, host_chunk, momentum=0.0, lr=None, next_host_chunk=None, lag=0):
        """`n` consecutive steps on minibatches that live in (pinned) HOST memory: host_chunk is [n, B, V] float32 —
        the streaming form of run_steps (an epoch, or the minibatches up to the next validation point) for datasets
        kept off the device.  The chunk is copied into one of two device staging buffers on a copy stream (the copy
        of `next_host_chunk` is enqueued right behind this chunk's launch and overlaps it), the n steps run as ONE
        mdbn_cd_steps launch on it, and the n costs are copied back on a third stream.
        lag=0: returns the n costs of THIS chunk (list of floats).  lag=1: returns those of the PREVIOUS call (None on
        the first; `flush_chunk()` returns the last) so that the host never blocks between launches."""
        n, B, V = (int(x) for x in host_chunk.shape)
        st = getattr(self, "_feedc", None)
        if st is None or st["shape"] != (n, B, V):
            main = torch.cuda.current_stream()
            copy, d2h = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
            st = self._feedc = {
                "shape": (n, B, V), "cur": 0, "staged": [None, None], "nbytes": n * B * V * 4,
                "buf": [torch.empty((n * B, V), dtype=torch.float32, device=self.device) for _ in range(2)],
                "ready": [torch.cuda.Event(), torch.cuda.Event()], "free": [torch.cuda.Event(), torch.cuda.Event()],
                "copy": copy, "copy_h": ctypes.c_void_p(copy.cuda_stream), "main": main,
                "main_h": ctypes.c_void_p(main.cuda_stream), "d2h": d2h, "d2h_h": ctypes.c_void_p(d2h.cuda_stream),
                "rows": torch.arange(n * B, dtype=torch.int32, device=self.device).view(n, B),
                "cost_dev": [torch.zeros(n, dtype=torch.float32, device=self.device) for _ in range(2)],
                "cost_host": [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(2)],
                "ev": [torch.cuda.Event(), torch.cuda.Event()], "n": 0, "pending": None}
        lib = self.rbm.ctx.lib
        main, copy = st["main"], st["copy"]

        def stage(slot, chunk):
            if chunk.dtype != torch.float32 or not chunk.is_contiguous():
                raise TypeError("run_steps_from_host takes contiguous float32 [n, B, V] host tensors (pinned for overlap)")
            if st["staged"][slot] is not None:
                copy.wait_event(st["free"][slot])
            _lib.check(lib.mdbn_copy_async(st["buf"][slot].data_ptr(), chunk.data_ptr(), st["nbytes"], st["copy_h"]))
            st["ready"][slot].record(copy)
            st["staged"][slot] = chunk
        cur = st["cur"]
        if st["staged"][cur] is not host_chunk:                  # not prefetched by the previous call
            stage(cur, host_chunk)
        main.wait_event(st["ready"][cur])
        k = st["n"] & 1
        if st["n"] >= 2:
            main.wait_event(st["ev"][k])                         # cost slot k was read back two calls ago
        sync, self.sync = self.sync, False
        try:
            self._call(st["rows"], momentum, lr, data_override=st["buf"][cur], stream=st["main_h"],
                       n_steps=n, costs=st["cost_dev"][k])
        finally:
            self.sync = sync
        st["free"][cur].record(main)
        nxt = 1 - cur
        if next_host_chunk is not None:
            stage(nxt, next_host_chunk)
        else:
            st["staged"][nxt] = None
        st["cur"] = nxt
        st["d2h"].wait_event(st["free"][cur])
        _lib.check(lib.mdbn_copy_async(st["cost_host"][k].data_ptr(), st["cost_dev"][k].data_ptr(), 4 * n, st["d2h_h"]))
        st["ev"][k].record(st["d2h"])
        prev, st["n"] = st["pending"], st["n"] + 1
        if not lag:
            st["pending"] = None
            st["ev"][k].synchronize()
            return [float(c) for c in st["cost_host"][k]]
        st["pending"] = k
        if prev is None:
            return None
        st["ev"][prev].synchronize()
        return [float(c) for c in st["cost_host"][prev]]

    def flush_chunk(self):
        """Costs of the last lagged chunk (run_steps_from_host(..., lag=1)); None when nothing is pending."""
        st = getattr(self, "_feedc", None)
        if not st or st["pending"] is None:
            return None
        k, st["pending"] = st["pending"], None
        st["ev"][k].synchronize()
        return [float(c) for c in st["cost_host"][k]]

    def run_steps(self, index_matrix, momentum=0.0, lr=None):
        """`n` consecutive steps from an [n, B] matrix of row numbers — an epoch, or the minibatches up to the
        next validation point (src/dbn.py:343-353).  Same parameters, chains and costs as n single calls;
        with the in-kernel generator on the skinny path it is ONE kernel launch (mdbn_cd_steps).
        Returns the n costs: a device tensor when sync is False, else a list of floats."""
        if isinstance(index_matrix, torch.Tensor):
            idx = index_matrix.to(device=self.device, dtype=torch.int32)
        else:
            idx = torch.as_tensor(numpy.asarray(index_matrix, dtype=numpy.int32)).to(self.device)
        idx = idx.reshape(1, -1) if idx.dim() == 1 else idx.contiguous()
        n = int(idx.shape[0])
        chained = (n > 1 and getattr(self.rbm.theano_rng, "mode", None) == _lib.RNG_PHILOX
                   and self.dp is None)
        if chained:
            costs = torch.empty(n, dtype=torch.float32, device=self.device)
            self._call(idx, momentum, lr, n_steps=n, costs=costs)
        else:
            sync, self.sync = self.sync, False
            try:
                costs = torch.stack([self(idx[s], momentum, lr).reshape(()).clone() for s in range(n)])
            finally:
                self.sync = sync
        if not self.sync:
            return costs
        return [float(c) for c in costs.cpu()]

    def _call(self, indexes, momentum=0.0, lr=None, phase=_lib.PHASE_FULL, rows_total=0, data_override=None,
              n_steps=1, costs=None, stream=None, comm=None):
        r, h = self.rbm, self.updates.hyper
        data = self.data() if data_override is None else data_override
        if indexes is None:                      # APPLY phase: no rows of its own
            idx = torch.zeros(1, dtype=torch.int32, device=self.device)
        elif isinstance(indexes, torch.Tensor):
            idx = indexes.to(device=self.device, dtype=torch.int32)
        else:
            idx = torch.as_tensor(numpy.asarray(indexes, dtype=numpy.int32)).to(self.device)
        B = int(idx.numel()) // int(n_steps)
        persistent = h.get("persistent")
        if persistent is not None and persistent.shape[0] != B and phase != _lib.PHASE_APPLY:
            raise ValueError("PCD chain has %d rows but the minibatch has %d (the reference fails the same way)"
                             % (persistent.shape[0], B))
        if h["batch_size"] is None:
            raise TypeError("batch_size=None: the W statistics are divided by batch_size (src/rbm.py:413)")
        if phase == _lib.PHASE_APPLY:
            rng, keep = _lib.Rng(_lib.RNG_PHILOX, None, 0, 0), None
        else:
            rng, keep = r.theano_rng.next_rng(id(self), self.device, layer_id=self.layer_id, B=B, n_steps=n_steps)
            if self.dp is not None and self.dp.world > 1 and rng.mode == _lib.RNG_PHILOX and phase != _lib.PHASE_APPLY:
                # data-parallel shards index their draws by the LOCAL row: give every rank its own Philox key so
                # that row b of two shards never sees the same uniforms (ranks must still agree on the minibatch)
                rng.seed = (rng.seed ^ (0x9E3779B97F4A7C15 * (self.dp.rank + 1))) & (2 ** 64 - 1)
        a = _lib.CdArgs()
        a.kind, a.noisy = r.kind, int(not getattr(r, "error_free", True))
        a.B, a.B_nom, a.V, a.H, a.k = B, int(h["batch_size"]), r.n_visible, r.n_hidden, int(h["k"])
        a.W, a.ldw = r.W.storage.data_ptr(), r.W.ld
        a.hbias, a.vbias = r.hbias.data.data_ptr(), r.vbias.data.data_ptr()
        a.W_speed = r.W_speed.storage.data_ptr()
        a.hbias_speed, a.vbias_speed = r.hbias_speed.data.data_ptr(), r.vbias_speed.data.data_ptr()
        snap = h.get("W_snap")
        a.W_snap = snap.data_ptr() if snap is not None else None
        a.data, a.ld_data, a.indices = data.data_ptr(), data.stride(0), idx.data_ptr()
        a.persistent = persistent.data.data_ptr() if persistent is not None else None
        a.bit_i_idx = r.bit_i_idx.data_ptr() if persistent is not None else None
        a.lr = float(h["lr"] if lr is None else lr)
        a.momentum = float(momentum)
        a.lambda_1, a.lambda_2, a.weightcost = float(h["lambda_1"]), float(h["lambda_2"]), float(h["weightcost"])
        a.rng = rng
        a.cost_out = self.cost_dev.data_ptr() if costs is None else costs.data_ptr()
        a.path, a.tf32, a.phase = _lib.PATHS[self.path], int(self.tf32), phase
        if comm is not None:
            a.comm, a.B_total = comm.handle, int(rows_total)
        if phase != _lib.PHASE_FULL:
            a.stats_buf, a.B_total = self._stats.data_ptr(), int(rows_total)
            if a.path in (_lib.PATH_SKINNY, _lib.PATH_TINY) or phase == _lib.PHASE_APPLY:
                a.path = _lib.PATH_AUTO      # the update from reduced statistics is one elementwise kernel
        stream = _stream() if stream is None else stream
        if n_steps > 1:
            _lib.check(r.ctx.lib.mdbn_cd_steps(r.ctx.handle, ctypes.byref(a), int(n_steps), stream))
        else:
            _lib.check(r.ctx.lib.mdbn_cd_step(r.ctx.handle, ctypes.byref(a), stream))
        self._keep = (keep, idx, data, costs)
        self.n_calls += int(n_steps)
        if phase == _lib.PHASE_STATS:
            return None
        for s in (r.W, r.hbias, r.vbias):
            s.version += 1
        if costs is not None:
            return costs
        if not self.sync:
            return self.cost_dev
        self._cost_host.copy_(self.cost_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self._cost_host[0])


class RBM(object):
    """Restricted Boltzmann Machine (src/rbm.py:46)."""
    kind = _lib.RBM

    def __init__(self, input=None, n_visible=784, n_hidden=500, W=None, hbias=None, vbias=None,
                 numpy_rng=None, theano_rng=None, device=None):
        self.n_visible = n_visible
        self.n_hidden = n_hidden
        self.device = torch.device(device) if device is not None else default_device()
        self.ctx = _lib.context(self.device.index if self.device.index is not None else torch.cuda.current_device())

        if numpy_rng is None:
            numpy_rng = numpy.random.RandomState(1234)                       # src/rbm.py:89
        if theano_rng is None:
            theano_rng = RandomStreams(numpy_rng.randint(2 ** 30))           # :92 (consumes one draw)
        if W is None:
            bound = 4 * numpy.sqrt(6. / (n_hidden + n_visible))              # :100-107
            W = numpy.asarray(numpy_rng.uniform(low=-bound, high=bound, size=(n_visible, n_hidden)),
                              dtype=numpy.float32)
        if not isinstance(W, Shared):
            W = Shared(W, name='W', device=self.device, ld_pad=_PAD)
        elif W.ld % _PAD != 0:
            # a user-made shared(W) is dense: move it into row-padded storage IN PLACE so that every holder
            # of this Shared (weight tying, src/dbn.py:193-202) keeps seeing the same tensor
            W.repad(_PAD)
        if hbias is None:
            hbias = numpy.zeros(n_hidden, dtype=numpy.float32)
        if not isinstance(hbias, Shared):
            hbias = Shared(hbias, name='hbias', device=self.device)
        if vbias is None:
            vbias = numpy.zeros(n_visible, dtype=numpy.float32)
        if not isinstance(vbias, Shared):
            vbias = Shared(vbias, name='vbias', device=self.device)
        assert W.shape == (n_visible, n_hidden) and hbias.shape == (n_hidden,) and vbias.shape == (n_visible,)

        self.input = input            # None, or a callable giving this layer's input from the DBN input
        self.W, self.hbias, self.vbias = W, hbias, vbias
        self.theano_rng = theano_rng
        self.params = [self.W, self.hbias, self.vbias]
        self.momentum = 0.0                                                   # :151
        self.W_speed = Shared(numpy.zeros((n_visible, n_hidden), numpy.float32), name='W_speed',
                              device=self.device, ld_pad=_PAD)               # :153-162
        if self.W_speed.ld != self.W.ld:
            raise ValueError("W (row stride %d) and W_speed (%d) must share a row stride" % (self.W.ld, self.W_speed.ld))
        self.hbias_speed = Shared(numpy.zeros(n_hidden, numpy.float32), name='hbias_speed', device=self.device)
        self.vbias_speed = Shared(numpy.zeros(n_visible, numpy.float32), name='vbias_speed', device=self.device)
        self.params_speed = [self.W_speed, self.hbias_speed, self.vbias_speed]
        self.bit_i_idx = torch.zeros(1, dtype=torch.int32, device=self.device)   # :425

    @property
    def Wt(self):
        return self.W.data.t()                                                # :139 (a view, never materialised)

    # ---- helpers -----------------------------------------------------------
    def _mat(self, x, n):
        t = as_device_matrix(x, self.device)
        if t.dim() != 2 or t.shape[1] != n:
            raise ValueError("expected a [B,%d] matrix, got %s" % (n, tuple(t.shape)))
        return t

    def _rng_for(self, u, site):
        if u is not None:
            return ExplicitBuffer(u).next_rng(site, self.device)
        if isinstance(self.theano_rng, BufferStreams):
            raise ValueError("BufferStreams drives whole CD steps; pass u=... to the single-phase sampling calls")
        return self.theano_rng.next_rng((id(self), site), self.device)

    def _propup(self, vis, want_pre=True, want_mean=True, u=None, sample=False):
        v = self._mat(vis, self.n_visible)
        B, H = v.shape[0], self.n_hidden
        mk = lambda on: torch.empty((B, H), dtype=torch.float32, device=self.device) if on else None
        pre, mean, smp = mk(want_pre), mk(want_mean), mk(sample)
        rng, keep = self._rng_for(u, "h") if sample else (None, None)
        _lib.check(self.ctx.lib.mdbn_propup(self.ctx.handle, self.W.storage.data_ptr(), self.W.ld,
                                            self.hbias.data.data_ptr(), v.data_ptr(), v.stride(0), B,
                                            self.n_visible, H, _ptr(pre), _ptr(mean), _ptr(smp),
                                            ctypes.byref(rng) if rng is not None else None, _stream()))
        return pre, mean, smp

    def _propdown(self, hid, kind, noisy=False, u=None, sample=False):
        h = self._mat(hid, self.n_hidden)
        B, V = h.shape[0], self.n_visible
        mk = lambda on: torch.empty((B, V), dtype=torch.float32, device=self.device) if on else None
        pre, mean, smp = mk(True), mk(True), mk(sample)
        needs = sample and (kind == _lib.RBM or noisy)
        rng, keep = self._rng_for(u, "v") if needs else (None, None)
        _lib.check(self.ctx.lib.mdbn_propdown(self.ctx.handle, self.W.storage.data_ptr(), self.W.ld,
                                              self.vbias.data.data_ptr(), h.data_ptr(), h.stride(0), B, V,
                                              self.n_hidden, kind, int(noisy), _ptr(pre), _ptr(mean), _ptr(smp),
                                              ctypes.byref(rng) if rng is not None else None, _stream()))
        return pre, mean, smp

    # ---- reference surface ---------------------------------------------------
    def free_energy(self, v_sample):
        """src/rbm.py:166-171 (GRBM: :684-688) -> [B]"""
        v = self._mat(v_sample, self.n_visible)
        F = torch.empty(v.shape[0], dtype=torch.float32, device=self.device)
        _lib.check(self.ctx.lib.mdbn_free_energy(self.ctx.handle, self.W.storage.data_ptr(), self.W.ld,
                                                 self.hbias.data.data_ptr(), self.vbias.data.data_ptr(),
                                                 v.data_ptr(), v.stride(0), v.shape[0], self.n_visible,
                                                 self.n_hidden, self.kind, F.data_ptr(), _stream()))
        return F

    def free_energy_gap(self, train, test):
        """mean F(test) - mean F(train)   src/rbm.py:173-180"""
        return self.free_energy(test).mean() - self.free_energy(train).mean()

    def free_energies(self, train, test):
        return self.free_energy(train), self.free_energy(test)               # :182-185

    def propup(self, vis):
        pre, mean, _ = self._propup(vis)                                      # :187-199
        return [pre, mean]

    def sample_h_given_v(self, v0_sample, u=None):
        pre, mean, smp = self._propup(v0_sample, u=u, sample=True)            # :201-213
        return [pre, mean, smp]

    def propdown(self, hid):
        pre, mean, _ = self._propdown(hid, _lib.RBM)                          # :215-227 (sigmoid, also for GRBM)
        return [pre, mean]

    def sample_v_given_h(self, h0_sample, u=None):
        pre, mean, smp = self._propdown(h0_sample, _lib.RBM, u=u, sample=True)   # :229-240
        return [pre, mean, smp]

    def gibbs_hvh(self, h0_sample, u_v=None, u_h=None):
        pre_v, v_mean, v_sample = self.sample_v_given_h(h0_sample, u=u_v)     # :242-248
        pre_h, h_mean, h_sample = self.sample_h_given_v(v_sample, u=u_h)
        return [pre_v, v_mean, v_sample, pre_h, h_mean, h_sample]

    def gibbs_vhv(self, v0_sample, u_h=None, u_v=None):
        pre_h, h_mean, h_sample = self.sample_h_given_v(v0_sample, u=u_h)     # :250-256
        pre_v, v_mean, v_sample = self.sample_v_given_h(h_sample, u=u_v)
        return [pre_h, h_mean, h_sample, pre_v, v_mean, v_sample]

    def make_sample_fn(self, persistent_vis_chain, plot_every=500):
        """The `sample_fn` of the reference's sampling demo (src/rbm.py:806-853): every call runs `plot_every`
        steps of `gibbs_vhv` on the persistent visible chains [n_chains, V] (a Shared, updated in place with
        the last visible sample) and returns (vis_mf, vis_sample) of the last step as numpy arrays.
        Works for GRBM as well (its gibbs_vhv feeds the hidden MEAN down, src/rbm.py:673-682)."""
        chain = persistent_vis_chain if isinstance(persistent_vis_chain, Shared) else \
            Shared(persistent_vis_chain, name='persistent_vis_chain', device=self.device)

        def sample_fn():
            v = chain.data
            v_mean = v
            for _ in range(int(plot_every)):
                _, _, _, _, v_mean, v = self.gibbs_vhv(v)
            chain.data.copy_(v)
            return v_mean.cpu().numpy(), v.cpu().numpy()
        sample_fn.chain = chain
        return sample_fn

    def get_cost_updates(self, lr=0.1, k=1, lambda_1=0.0, lambda_2=0.0, weightcost=0.0,
                         batch_size=None, persistent=None, symbolic_grad=False):
        """One step of CD-k / PCD-k (src/rbm.py:258-376).  Returns (cost, updates): `updates`
        describes the in-place step; `make_train_fn(dataset, cost, updates)` compiles it.
        W_snap (the constant `weightcost` multiplies, :414-415) is captured HERE, like the
        reference captures `self.W.get_value()` at graph-build time."""
        if symbolic_grad:
            raise NotImplementedError("symbolic_grad=True is never enabled by any caller of the reference "
                                      "(src/rbm.py:341-342) and has no kernel")
        if persistent is not None and not isinstance(persistent, Shared):
            persistent = Shared(persistent, name='persistent', device=self.device)
        snap = self.W.storage.clone() if weightcost != 0 else None
        updates = CDUpdates(self, lr=lr, k=k, lambda_1=lambda_1, lambda_2=lambda_2, weightcost=weightcost,
                            batch_size=batch_size, persistent=persistent, W_snap=snap)
        return Cost(updates), updates

    def make_train_fn(self, train_set_x, cost, updates, layer_id=0, input_fn=None, path="auto", tf32=False):
        assert cost.updates is updates
        return TrainFn(self, updates, train_set_x, layer_id=layer_id, input_fn=input_fn, path=path, tf32=tf32)

    def training(self, train_set_x, validation_set_x, training_epochs, batch_size=10, learning_rate=0.1, k=1,
                 initial_momentum=0.0, final_momentum=0.0, weightcost=0.0, lambda_2=0.0, persistent=True,
                 display_fn=None, graph_output=False):
        """src/rbm.py:484-520.  NB: `lambda_2` is accepted and not forwarded, and the default
        is PCD with a chain of zeros, exactly like the reference (App. C-6/7)."""
        if persistent:
            persistent_chain = Shared(numpy.zeros((batch_size, self.n_hidden), numpy.float32), device=self.device)
        else:
            persistent_chain = None
        cost, updates = self.get_cost_updates(lr=learning_rate, k=k, weightcost=weightcost,
                                              batch_size=batch_size, persistent=persistent_chain)
        return self.learn_model(train_set_x=train_set_x, validation_set_x=validation_set_x,
                                training_epochs=training_epochs, batch_size=batch_size,
                                initial_momentum=initial_momentum, final_momentum=final_momentum,
                                cost=cost, updates=updates, display_fn=display_fn, graph_output=graph_output)

    def learn_model(self, train_set_x, validation_set_x, training_epochs, batch_size,
                    initial_momentum, final_momentum, cost, updates, display_fn, graph_output, verbose=True):
        """Epoch loop of src/rbm.py:522-629.  Returns [(mean cost, free-energy gap)] per epoch."""
        train_rbm = self.make_train_fn(train_set_x, cost, updates)
        train_rbm.sync = False
        train = train_rbm.data()
        val = as_device_matrix(validation_set_x, self.device)
        n_train_data, n_val = train.shape[0], val.shape[0]
        start_time = timeit.default_timer()
        momentum = initial_momentum
        history = []
        feeder = IndexFeeder(self.device, n_train_data)
        for epoch in range(training_epochs):
            if epoch == 6:                                                    # :584 (0-based)
                momentum = final_momentum
            _, minibatches = get_minibatches_idx(n_train_data, batch_size, shuffle=True)
            # one H2D copy of the whole epoch's index list instead of one per step
            flat = feeder.upload(minibatches)
            costs = torch.empty(len(minibatches), dtype=torch.float32, device=self.device)
            lo, i = 0, 0
            while i < len(minibatches):          # runs of equal-length minibatches: one chained launch each
                B0, n = len(minibatches[i]), 1
                while i + n < len(minibatches) and len(minibatches[i + n]) == B0:
                    n += 1
                costs[i:i + n].copy_(train_rbm.run_steps(flat[lo:lo + n * B0].view(n, B0), momentum))
                lo += n * B0
                i += n
            feg = float(self.free_energy_gap(train[:n_val], val))             # :597, :549-558
            mean_cost = float(costs.double().mean())
            history.append((mean_cost, feg))
            if verbose:
                print('Training epoch %d, cost is ' % epoch, mean_cost)
                print('Free energy gap is ', feg)
            if display_fn is not None:
                display_fn(self.W.get_value(borrow=True), self.n_hidden)
        if verbose:
            print('Training took %f minutes' % ((timeit.default_timer() - start_time) / 60.))
        return history


class GRBM(RBM):
    """Gaussian-Bernoulli RBM (src/rbm.py:631-728): linear unit-variance visibles."""
    kind = _lib.GRBM

    def __init__(self, input=None, n_visible=784, n_hidden=500, W=None, hbias=None, vbias=None,
                 numpy_rng=None, theano_rng=None, error_free=True, device=None):
        super(GRBM, self).__init__(input, n_visible, n_hidden, W, hbias, vbias, numpy_rng, theano_rng, device=device)
        self.error_free = error_free

    def sample_v_given_h(self, h0_sample, u=None):
        """[v1_mean, v1_mean, v1_sample] — slot 0 is NOT a pre-sigmoid (src/rbm.py:647-660)."""
        pre, mean, smp = self._propdown(h0_sample, _lib.GRBM, noisy=not self.error_free, u=u, sample=True)
        return [mean, mean, smp]

    def gibbs_hvh(self, h0_sample, u_v=None, u_h=None):
        pre_v, v_mean, v_sample = self.sample_v_given_h(h0_sample, u=u_v)     # :662-671
        pre_h, h_mean, h_sample = self.sample_h_given_v(v_mean, u=u_h)        # h given the MEAN
        return [pre_v, v_mean, v_sample, pre_h, h_mean, h_sample]

    def gibbs_vhv(self, v0_sample, u_h=None, u_v=None):
        pre_h, h_mean, h_sample = self.sample_h_given_v(v0_sample, u=u_h)     # :673-682
        pre_v, v_mean, v_sample = self.sample_v_given_h(h_mean, u=u_v)        # v given the MEAN
        return [pre_h, h_mean, h_sample, pre_v, v_mean, v_sample]

    def training(self, train_set_x, validation_set_x, training_epochs, batch_size=10, learning_rate=0.01, k=1,
                 initial_momentum=0.0, final_momentum=0.0, weightcost=0.0, lambda_1=0.0, lambda_2=0.1,
                 persistent=False, display_fn=None, graph_output=False):
        """src/rbm.py:701-728; `persistent` is accepted and ignored like the reference (always CD)."""
        cost, updates = self.get_cost_updates(lr=learning_rate, k=k, lambda_1=lambda_1, lambda_2=lambda_2,
                                              weightcost=weightcost, batch_size=batch_size)
        return self.learn_model(train_set_x=train_set_x, validation_set_x=validation_set_x,
                                training_epochs=training_epochs, batch_size=batch_size,
                                initial_momentum=initial_momentum, final_momentum=final_momentum,
                                cost=cost, updates=updates, display_fn=display_fn, graph_output=graph_output)
