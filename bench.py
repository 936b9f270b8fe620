#!/usr/bin/env python
"""bench.py — RBM CD-k training samples/sec/layer on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one full CD-k / PCD-k parameter update (positive phase, k Gibbs steps, statistics,
lambda_1/lambda_2/momentum update, monitoring cost) of one layer on one minibatch.

Default workload (N=1): BASELINE.json configs[1] — Gaussian-Bernoulli RBM 19937->400 (AML
gene-expression shape), PCD-1, batch 10, on synthetic z-scored data [170,19937].
With N>1 every rank trains its own layer of that shape (the path shards by independent
layers/modalities; no data-path collective) -> weak scaling.

The timed region issues what the reference's training loops issue through this package (DBN.training /
RBM.learn_model -> TrainFn.run_steps): the minibatches of an epoch in ONE launch (mdbn_cd_steps).  Between
launches the parameter sets rotate over 4 independent layers (4 x (W + W_speed) = 255 MB > the 126 MB L2),
so every launch starts with its weights in HBM; `single_launch` is the same step as one launch per minibatch.

Timing hygiene: W >= 3 warm-up steps; CUDA events on the launching stream; barrier + synchronize on both
sides (a ~1 ms spin kernel in front of the first event absorbs the host's enqueue latency); max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RBM CD-k training samples/sec/layer"
WORKLOADS = {
    # name: kind, V, H, B, k, pcd, N rows, lr, momentum, lambda_1, lambda_2, weightcost
    "ge_grbm_pcd1_b10": dict(kind=1, V=19937, H=400, B=10, k=1, pcd=True, N=170, lr=0.005, mom=0.0, l1=0.01,
                             l2=0.1, wc=0.0),
    "ge_grbm_cd1_b20": dict(kind=1, V=19937, H=400, B=20, k=1, pcd=False, N=170, lr=0.005, mom=0.0, l1=0.01,
                            l2=0.1, wc=0.0),
    "mnist_rbm_cd1_b20": dict(kind=0, V=784, H=500, B=20, k=1, pcd=False, N=49940, lr=0.1, mom=0.9, l1=0.0,
                              l2=0.0, wc=0.0002),
    "mnist_rbm_pcd1_b20": dict(kind=0, V=784, H=500, B=20, k=1, pcd=True, N=49940, lr=0.1, mom=0.9, l1=0.0,
                               l2=0.0, wc=0.0002),
    # batch 21..128 (north_star (c): "batch 10-100"): path=auto -> the tcgen05 path in fp32-exact split-TF32 arithmetic
    "ge_grbm_pcd1_b32": dict(kind=1, V=19937, H=400, B=32, k=1, pcd=True, N=416, lr=0.005, mom=0.0, l1=0.01, l2=0.1, wc=0.0),
    "ge_grbm_cd1_b50": dict(kind=1, V=19937, H=400, B=50, k=1, pcd=False, N=400, lr=0.005, mom=0.0, l1=0.01, l2=0.1, wc=0.0),
    "ge_grbm_pcd1_b100": dict(kind=1, V=19937, H=400, B=100, k=1, pcd=True, N=400, lr=0.005, mom=0.0, l1=0.01, l2=0.1,
                              wc=0.0),
    "mnist_rbm_cd1_b100": dict(kind=0, V=784, H=500, B=100, k=1, pcd=False, N=50000, lr=0.1, mom=0.9, l1=0.0, l2=0.0,
                               wc=0.0002),
    # BASELINE.json configs[4]: large-batch / many-chain PCD on the tcgen05 TF32 path; with --gpus N the
    # minibatch rows (and chains) are sharded over ranks and the packed statistics all-reduced (strong scaling)
    "rbm_784x500_b8192_pcd1_tf32": dict(kind=0, V=784, H=500, B=8192, k=1, pcd=True, N=65536, lr=0.1, mom=0.9,
                                        l1=0.0, l2=0.0, wc=0.0002, tensor=True, dp=True),
    "rbm_784x500_b1024_pcd1_tf32": dict(kind=0, V=784, H=500, B=1024, k=1, pcd=True, N=65536, lr=0.1, mom=0.9,
                                        l1=0.0, l2=0.0, wc=0.0002, tensor=True, dp=True),
    "rbm_784x500_b8192_pcd10_tf32": dict(kind=0, V=784, H=500, B=8192, k=10, pcd=True, N=65536, lr=0.1, mom=0.9,
                                         l1=0.0, l2=0.0, wc=0.0002, tensor=True, dp=True),
    "grbm_19937x400_b2048_cd1_tf32": dict(kind=1, V=19937, H=400, B=2048, k=1, pcd=False, N=4096, lr=0.005, mom=0.0,
                                          l1=0.01, l2=0.1, wc=0.0, tensor=True, dp=True),
}
DEFAULT = "ge_grbm_pcd1_b10"


def synth(kind, n, V, seed):
    rs = np.random.RandomState(seed)
    if kind == 1:
        x = rs.randn(n, V).astype(np.float32)
        return ((x - x.mean(0)) / x.std(0)).astype(np.float32)     # per-feature z-score (src/utils.py:96)
    return (rs.rand(n, V) < 0.13).astype(np.float32)


def algorithmic_bytes(w):
    """4*[(2k+5) V H + B V + (2k+1)(V+H) + 4 (V+H)] bytes per CD-k step."""
    V, H, B, k = w["V"], w["H"], w["B"], w["k"]
    return 4 * ((2 * k + 5) * V * H + B * V + (2 * k + 1) * (V + H) + 4 * (V + H))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def algorithmic_flops(w):
    """2 B V H (2k+3) flops per CD-k step (monitor GEMMs excluded)."""
    return 2.0 * w["B"] * w["V"] * w["H"] * (2 * w["k"] + 3)


def tf32_peak_tflops(torch, dev):
    """TF32 is not in MEASURED_PEAKS.json: measured the way that file measures bf16 — cuBLAS
    torch.matmul 8192^3 with TF32 enabled, best of 10 (burst)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=dev)
    b = torch.randn(8192, 8192, device=dev)
    best = 1e9
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2 * 8192.0 ** 3 / (best * 1e-3) / 1e12


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return None
        hi = [s for s in sm if s > 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons)}


def config_of(wname, w):
    """The workload description — byte-identical in both arms (b200 and --impl reference)."""
    return {"workload": wname, "layer": "%s %d->%d" % ("GRBM" if w["kind"] else "RBM", w["V"], w["H"]),
            "batch": w["B"], "k": w["k"], "pcd": bool(w["pcd"])}


def host_threads():
    """All host cores for the BLAS of the CPU arm, whatever OMP_NUM_THREADS the launcher exported (torchrun
    sets it to 1)."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n), n
    except Exception:
        return None, n


def cpu_step_rate(w, seconds=15.0, max_steps=60, warmup=1):
    """The oracle port (NumPy fp32, OpenBLAS on all host threads) timed on this host's cores."""
    from oracle import rbm_oracle as O, shared_u
    limiter, cores = host_threads()
    try:
        from threadpoolctl import threadpool_info
        cores = max([i.get("num_threads", 1) for i in threadpool_info()] or [cores])
    except Exception:
        pass
    kind, V, H, B, k = w["kind"], w["V"], w["H"], w["B"], w["k"]
    data = synth(kind, max(w["N"] if w["N"] < 1000 else 1000, B), V, 1)
    L = O.Layer(V, H, kind, numpy_rng=np.random.RandomState(123), dtype=np.float32)
    snap = L.W.copy()
    P = np.zeros((B, H), np.float32) if w["pcd"] else None
    U = shared_u.step_buffer(1, 0, 0, kind, True, B, V, H, k)
    rs = np.random.RandomState(0)
    times = []
    t_end = time.perf_counter() + seconds
    n = 0
    while n < max_steps + warmup and (time.perf_counter() < t_end or n < warmup + 3):
        idx = rs.randint(0, data.shape[0], B)
        t0 = time.perf_counter()
        O.cd_step(L, data[idx], U, lr=w["lr"], k=k, lambda_1=w["l1"], lambda_2=w["l2"], weightcost=w["wc"],
                  batch_size=B, momentum=w["mom"], persistent=P, W_snap=snap)
        dt = time.perf_counter() - t0
        if n >= warmup:
            times.append(dt)
        n += 1
    med = float(np.median(times))
    return B / med, cores, len(times), med


def run_reference(args, w, wname):
    """--impl reference: the reference's CPU path for this step.  Theano cannot be installed here
    (SURVEY.md 8c), so this is the oracle PORT (NumPy restatement pinned to the reference source,
    oracle/rbm_oracle.py) on ALL host threads (set explicitly: torchrun exports OMP_NUM_THREADS=1); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = 150.0
    v, cores, n, med = cpu_step_rate(w, seconds=budget, max_steps=max(args.steps, 3), warmup=min(args.warmup, 3))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(wname, w),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "%d CD steps of the workload timed (median), NumPy/OpenBLAS fp32 oracle port on %d "
                                   "threads; Theano itself is not installable" % (n, cores)},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def dp_extra(torch, dist, M, dev, world, rank, k, steps=12):
    """Large-batch data-parallel CD (BASELINE.json configs[4]): RBM 784->500, B = 8192, PCD-k on the tcgen05 TF32
    path; the minibatch rows and chains are sharded over the ranks and the packed statistics all-reduced.
    Rank 0 first times the same step alone (the N = 1 figure of THIS run), then all ranks time the sharded one."""
    from mdbn_b200.parallel import DataParallel
    V, H, B = 784, 500, 8192
    data = torch.from_numpy(synth(0, 16384, V, 1)).to(dev)

    def make(rows, dp):
        m = M.RBM(n_visible=V, n_hidden=H, numpy_rng=np.random.RandomState(123), theano_rng=M.RandomStreams(1000))
        P = M.shared(np.zeros((rows, H), np.float32))
        cost, upd = m.get_cost_updates(lr=0.1, k=k, weightcost=0.0002, batch_size=B, persistent=P)
        fn = m.make_train_fn(data, cost, upd, path="tensor", tf32=True)
        fn.sync = False
        if dp:
            fn.dp = DataParallel()
        return fn
    idx = [torch.arange(i * B, (i + 1) * B, dtype=torch.int32, device=dev) for i in range(2)]

    def timed(fn, n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for s in range(n):
            fn(idx[s & 1], 0.9)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / n
    ms1 = None
    if rank == 0:
        f1 = make(B, False)
        timed(f1, 6)
        ms1 = min(timed(f1, steps) for _ in range(3))
        del f1
    dist.barrier()
    fN = make(B // world, True)
    timed(fN, 10)                  # (the communicator's first collectives set up their channels)
    msN = 1e9
    for _ in range(3):             # best of three windows on both arms
        dist.barrier()
        msN = min(msN, timed(fN, steps))
    t = torch.tensor([msN], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    msN = float(t.item())
    if rank != 0:
        return None
    v1, vN = B / (ms1 * 1e-3), B / (msN * 1e-3)
    return {"workload": "rbm_784x500_b8192_pcd%d_tf32" % k, "value": vN, "unit": "samples/s", "ms_per_step": msN,
            "value_n1_same_run": v1, "efficiency_vs_n1": vN / (world * v1), "scaling": "strong",
            "parallelism": "minibatch rows and chains sharded over %d ranks + all-reduce of the packed statistics" % world}


def mdbn_wallclock(torch, dist, world, rank, scale):
    """MDBN pretrain wall-clock (BASELINE.json metric, part 2): AML-shaped synthetic modalities, one DBN per
    modality per GPU (round-robin), joint DBN on rank 0; the reference's early-stopping logic unchanged."""
    from mdbn_b200.parallel import train_modalities, aml_synthetic_specs
    specs = aml_synthetic_specs(scale)
    # warm-up, untimed like every other leg: the same run with 0.2 % of the patience budgets loads the kernels of every
    # layer shape on every rank (a fresh process otherwise pays ~0.5 s of lazy module loading inside the timed region)
    train_modalities(aml_synthetic_specs(0.002 * scale), batch_size=20, top=True)
    np.random.seed(20161230 + rank)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    train_modalities(specs, batch_size=20, top=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=204)
    ap.add_argument("--warmup", type=int, default=34)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the MDBN wall-clock and data-parallel legs")
    ap.add_argument("--mdbn-scale", type=float, default=1.0)
    ap.add_argument("--replicas", type=int, default=4)
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w, args.workload)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import mdbn_b200 as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    kind, V, H, B, k = w["kind"], w["V"], w["H"], w["B"], w["k"]
    tensor, dp = bool(w.get("tensor")), bool(w.get("dp")) and world > 1
    path = "tensor" if tensor else "auto"
    if dp:
        assert B % world == 0
    data_h = synth(kind, min(w["N"], 4096 if not tensor else 65536), V, 1 + (0 if dp else rank))
    data = torch.from_numpy(data_h).to(dev)
    cls = M.GRBM if kind == 1 else M.RBM
    R = max(1, args.replicas) if not tensor else 1     # large-batch working sets already exceed L2
    Bl = B // world if dp else B                       # rows (and chains) this rank owns

    def make_fn(r, dataset):
        m = cls(n_visible=V, n_hidden=H, numpy_rng=np.random.RandomState(123 + r),
                theano_rng=M.RandomStreams(1000 + 17 * r + rank))
        P = M.shared(np.zeros((Bl, H), np.float32)) if w["pcd"] else None
        cost, upd = m.get_cost_updates(lr=w["lr"], k=k, lambda_1=w["l1"], lambda_2=w["l2"], weightcost=w["wc"],
                                       batch_size=B, persistent=P)
        fn = m.make_train_fn(dataset, cost, upd, path=path, tf32=tensor)
        fn.sync = False
        if dp:
            from mdbn_b200.parallel import DataParallel
            fn.dp = DataParallel()
        return m, fn
    layers, fns = zip(*[make_fn(r, data) for r in range(R)])
    ctx = layers[0].ctx
    n_rows = data.shape[0]
    n_mb = n_rows // B
    rs = np.random.RandomState(5)
    perm = torch.from_numpy(rs.permutation(n_rows)[: n_mb * B].astype(np.int32)).to(dev)
    mbs = [perm[i * B:(i + 1) * B] for i in range(n_mb)]
    # an epoch per launch, as DBN.training / RBM.learn_model issue it (src/dbn.py:444-455 is the loop they chain)
    chained = not tensor and not dp and n_mb >= 2
    chain = min(n_mb, 32) if chained else 1
    idx_mat = perm[: chain * B].view(chain, B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def issue(n_steps, rotate, per_launch):
        """n_steps CD steps: `per_launch` minibatches per launch, the parameter sets rotate between launches."""
        done, l = 0, 0
        while done < n_steps:
            n = min(per_launch, n_steps - done)
            f = fns[l % R if rotate else 0]
            if per_launch > 1:
                f.run_steps(idx_mat[:n], w["mom"])
            else:
                f(mbs[done % n_mb], w["mom"])
            done += n
            l += 1
        return l

    def timed(n_steps, rotate, per_launch):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        # pre-roll: a spin kernel keeps the stream busy (~1 ms) while the host enqueues the launches behind it, so the
        # events bracket device time of exactly n_steps steps and not the host's enqueue latency of the first launch
        # (a training loop of thousands of launches runs with the host ahead of the device)
        torch.cuda._sleep(2000000)
        ev0.record()
        issue(n_steps, rotate, per_launch)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    timed(max(args.warmup, 2 * R * chain if chained else args.warmup), True, chain)   # warm-up (untimed): every set, twice
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.15)
    l0 = ctx.launches
    ms = timed(args.steps, True, chain)                        # ---- the timed region ----
    launches = ctx.launches - l0
    # keep the GPU busy a little longer so that the 50 ms clock sampler sees it under load
    t_hold = time.time() + 0.6
    while time.time() < t_hold:
        timed(2 * chain if chained else 50, True, chain)
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    value = (1 if dp else world) * B / (ms_per_step * 1e-3)
    # beside the headline: one launch per minibatch (cold rotation), and the chained form on ONE set (L2-warm)
    single = warm = None
    if chained:
        timed(max(args.warmup, 2 * R), True, 1)
        ms_single = timed(args.steps, True, 1) / args.steps
        ms_warm = timed(args.steps, False, chain) / args.steps
        single = {"ms_per_step": ms_single, "value": world * B / (ms_single * 1e-3)}
        warm = {"ms_per_step": ms_warm, "value": world * B / (ms_warm * 1e-3)}

    # ---- end to end: host minibatches -> pinned -> H2D -> steps -> D2H costs, every launch ----
    if chained:
        n_host = 2 * R
        host_chunks = [torch.from_numpy(np.ascontiguousarray(data_h[rs.randint(0, n_rows, chain * B)]).reshape(chain, B, V)).pin_memory()
                       for _ in range(n_host)]
        stage = torch.empty((chain * B, V), dtype=torch.float32, device=dev)
        e2e_fns = [make_fn(r, stage)[1] for r in range(R)]
        n_e2e_launches = max(2, -(-min(args.steps, 408) // chain))
        n_e2e = n_e2e_launches * chain

        def e2e_loop(nl):
            # public streaming call: every launch's minibatches are copied host -> device inside the loop (double
            # buffered on a copy stream, overlapping the previous launch) and the costs of every step are read back
            acc = 0.0
            for l in range(nl):
                c = e2e_fns[l % R].run_steps_from_host(host_chunks[l % n_host], w["mom"],
                                                       next_host_chunk=host_chunks[(l + R) % n_host], lag=1)
                acc += sum(c) if c is not None else 0.0
            for f in e2e_fns:
                c = f.flush_chunk()
                acc += sum(c) if c is not None else 0.0
            return acc
        e2e_loop(3 * R)                                   # every rotating function warmed (set-up, pinned buffers, streams)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(n_e2e_launches)
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        e2e_path = ("RBM.make_train_fn(...).run_steps_from_host -> mdbn_cd_steps (C ABI): an epoch of pinned host "
                    "minibatches per launch, double-buffered H2D on a copy stream, every step's cost copied D2H")
    else:
        n_host = 2
        host_mb = [torch.from_numpy(np.ascontiguousarray(data_h[rs.randint(0, n_rows, B)])).pin_memory() for _ in range(n_host)]
        stage = torch.empty((B, V), dtype=torch.float32, device=dev)
        e2e_fns = []
        for r in range(R):
            f2 = make_fn(r, stage)[1]
            f2.sync = True
            e2e_fns.append(f2)
        rows = torch.arange(B, dtype=torch.int32, device=dev)
        n_e2e = min(args.steps, 400)

        def e2e_loop(n):
            acc = 0.0
            for s in range(n):
                stage.copy_(host_mb[s % n_host], non_blocking=True)
                acc += e2e_fns[s % R](rows, w["mom"])
            return acc
        e2e_loop(max(3, 2 * R))
        barrier()
        t0 = time.perf_counter()
        e2e_loop(n_e2e)
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        e2e_path = "TrainFn.__call__ on a staging buffer filled from pinned host memory every step; cost read back every step"
    if world > 1:
        t = torch.tensor([t_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = (1 if dp else world) * B * n_e2e / t_e2e

    # ---- the two partitions of the path that DO exchange data, reported beside the headline ----
    extras = {}
    if not args.no_extras and not tensor:
        del e2e_fns
        torch.cuda.empty_cache()
        try:
            extras["mdbn_aml_wallclock_s"] = mdbn_wallclock(torch, dist, world, rank, args.mdbn_scale)
            extras["mdbn_aml"] = {"config": "AML-shaped synthetic: ME 559->40 (k=10), GE 19937->400->40, SM 1686->200->20, "
                                            "joint 100->24->3, N=170, batch 20 (BASELINE.json configs[3])",
                                  "scale": args.mdbn_scale,
                                  "parallelism": "one modality DBN per GPU (round-robin over %d), joint DBN on rank 0" % world}
        except Exception as e:      # never lose the headline line to an extra
            extras["mdbn_aml_error"] = repr(e)[:300]
        if world > 1:
            try:
                extras["dp"] = [dp_extra(torch, dist, M, dev, world, rank, kk) for kk in (1, 10)]
            except Exception as e:
                extras["dp_error"] = repr(e)[:300]

    if rank == 0:
        peak, how = peaks()
        abytes = algorithmic_bytes(w)
        achieved = abytes / (ms_per_step * 1e-3) / 1e9
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get(args.workload)
            traffic_note = tj.get("_how")
        if tensor:
            tpeak = tf32_peak_tflops(torch, dev)
            aflops = algorithmic_flops(w) / (world if dp else 1)
            ach = aflops / (ms_per_step * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                    "traffic": None, "kernel": "tc_gemm_kernel (tcgen05 tf32)",
                    "peak_source": "measured here: torch.matmul 8192^3 TF32, best of 10 (MEASURED_PEAKS.json has no TF32 row)",
                    "algorithmic_flops_per_step_per_gpu": aflops}
        else:
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_how": traffic_note,
                    # B <= 20: the broadcast kernel where the layer fits it (api.cu mid_wanted: V * ldw <= 4M), else
                    # the row-slab kernel
                    "kernel": ("cd_mid_kernel" if V * ((H + 7) // 8 * 8) <= (4 << 20) else "cd_skinny_kernel") if B <= 20 else
                              "tc_gemm_kernel<SPLIT> (tcgen05, fp32-exact split TF32; statistics GEMM with the update fused "
                              "in its epilogue)",
                    "peak_source": how,
                    "algorithmic_bytes_per_step": abytes, "steps_per_launch": chain,
                    "algorithmic_bytes_per_launch": abytes * chain}
            if single:
                roof["single_launch"] = {"ms_per_step": single["ms_per_step"], "value": single["value"],
                                         "frac": abytes / (single["ms_per_step"] * 1e-3) / 1e9 / peak}
                roof["l2_warm"] = {"ms_per_step": warm["ms_per_step"], "value": warm["value"],
                                   "frac": abytes / (warm["ms_per_step"] * 1e-3) / 1e9 / peak}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if dp else "weak",
            "vs_baseline": None, "dtype": "tf32" if tensor else ("f32" if B <= 20 else "f32 (3xTF32 on tcgen05)"), "data": "synthetic",
            "config": config_of(args.workload, w),
            "notes": {"rng": "philox4x32-10 in-kernel",
                      "launches": ("one launch per epoch of %d minibatches (TrainFn.run_steps -> mdbn_cd_steps), as "
                                   "DBN.training issues them" % chain) if chained else "one launch sequence per step",
                      "l2": ("inputs larger than L2: activations + dataset of a %d-row batch" % B) if tensor else
                            ("the parameter sets rotate over %d independent layers between launches (%.0f MB > L2): every "
                             "launch starts with its weights in HBM" % (R, R * 2 * V * H * 4 / 1e6)),
                      "parallelism": ("minibatch rows sharded over %d ranks + NCCL all-reduce of the packed statistics" % world)
                      if dp else "one independent layer per GPU (modality-parallel), no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": B * V * 4, "d2h_bytes_per_step": 4,
                    "steps": n_e2e, "path": e2e_path},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        line.update(extras)
        if not args.no_cpu_baseline:
            v, cores, n, med = cpu_step_rate(w)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": "%d CD steps of the same workload (median %.1f ms/step), NumPy/OpenBLAS "
                                              "fp32 oracle port standing in for Theano" % (n, med * 1e3)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
