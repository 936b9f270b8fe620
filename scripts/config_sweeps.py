#!/usr/bin/env python
"""BASELINE.json configs[2] and configs[4] as measured tables.

    python scripts/config_sweeps.py                                              # 1 GPU
    torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/config_sweeps.py  # config 5 sharded over N ranks

config 5: RBM 784->500, PCD-k with B chains, lr 0.1, momentum 0.9, weightcost 0.0002, B in {128 ... 8192} x k in {1, 2, 5, 10},
          TF32 tensor path; with N ranks the rows / chains are sharded and the packed statistics all-reduced through the
          library's NCCL communicator.  Reports samples/s, us per step, fraction of the TF32 peak measured here.
config 3: DBN 784-1000-1000-1000, greedy CD-1, B = 20, lr 0.01 (src/dbn.py:623-629 shape of the demo), on synthetic
          binarised data [50000, 784]: wall-clock of DBN.training for a fixed budget of iterations per layer (rank 0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import mdbn_b200 as M
from mdbn_b200.parallel import DataParallel

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)


def tf32_peak():
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
    best = 1e9
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        if i >= 2: best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = False
    return 2 * 8192.0 ** 3 / (best * 1e-3) / 1e12


peak = tf32_peak()
V, H = 784, 500
data = torch.from_numpy((np.random.RandomState(1).rand(16384, V) < 0.13).astype(np.float32)).to(dev)
dp = DataParallel() if world > 1 else None
rows = []
for B in (128, 512, 2048, 8192):
    for k in (1, 2, 5, 10):
        Bl = B // world
        if Bl < 32 or Bl % 32:
            continue
        m = M.RBM(n_visible=V, n_hidden=H, numpy_rng=np.random.RandomState(123), theano_rng=M.RandomStreams(1000))
        P = M.shared(np.zeros((Bl, H), np.float32))
        cost, upd = m.get_cost_updates(lr=0.1, k=k, weightcost=0.0002, batch_size=B, persistent=P)
        fn = m.make_train_fn(data, cost, upd, path="tensor", tf32=True)
        fn.sync = False
        fn.dp = dp
        idx = [torch.arange(i * B, (i + 1) * B, dtype=torch.int32, device=dev) for i in range(2)]
        for s in range(4):
            fn(idx[s & 1], 0.9)
        n = 30 if B <= 2048 else 12
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        e0.record()
        for s in range(n):
            fn(idx[s & 1], 0.9)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        flops = 2.0 * B * V * H * (2 * k + 3) / world
        rows.append({"B": B, "k": k, "n_gpus": world, "us_per_step": round(ms * 1e3, 1), "samples_per_s": round(B / (ms * 1e-3)),
                     "tf32_frac_per_gpu": round(flops / (ms * 1e-3) / 1e12 / peak, 3)})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
out = {"config5": rows, "tf32_peak_tflops_measured": round(peak, 1)}
if rank == 0:
    # ---- config 3: DBN 784-1000-1000-1000, B = 20, CD-1 ----
    N = 50000
    x = (np.random.RandomState(0).rand(N, 784) < 0.13).astype(np.float32)
    budget = [5000, 5000, 5000]            # iterations per layer (patience is counted in iterations, src/dbn.py:440)
    d = M.DBN(numpy_rng=np.random.RandomState(123), n_ins=784, gauss=False, hidden_layers_sizes=[1000, 1000], n_outs=1000, verbose=False)
    np.random.seed(7)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hist = d.training(x, 20, 1, budget, [0.01, 0.01, 0.01], validation_set_x=None)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    calls = [h["calls"] for h in hist]
    out["config3"] = {"layers": "784-1000-1000-1000 (Bernoulli), B=20, CD-1, lr 0.01", "cd_steps_per_layer": calls, "wallclock_s": round(dt, 3),
                      "us_per_step_overall": round(dt / sum(calls) * 1e6, 1), "samples_per_s": round(20 * sum(calls) / dt),
                      "note": "pretraining only: the reference has no fine-tuning code (src/mlp.py holds HiddenLayer alone)"}
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
