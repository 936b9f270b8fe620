#!/usr/bin/env python
"""MDBN pretrain wall-clock on AML-shaped synthetic data (BASELINE.json metric, part 2; configs[3]).

    python scripts/mdbn_aml_wallclock.py                      # 1 GPU: modalities one after the other (like the reference)
    torchrun --nproc-per-node 3 scripts/mdbn_aml_wallclock.py # one modality DBN per GPU, joint DBN on rank 0

Shapes and hyper-parameters: SURVEY.md 8(d) config 4 (src/AMLsm.py:38-62,207-339, src/AMLsm2.py:308-339,
src/MDBN.py:31-41): ME 559->40 (k=10, patience 80000), GE 19937->400->40, SM 1686->200->20, joint 100->24->3,
N=170, batch 20, the reference's early-stopping logic unchanged.  --scale shrinks every patience/epoch
budget by that factor (for quick checks; a scaled run is NOT the named config)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from mdbn_b200.parallel import train_modalities

def zs(x):
    return ((x - x.mean(0)) / x.std(0)).astype(np.float32)
N = 170
me = zs(np.random.RandomState(2).randn(N, 559))
ge = zs(np.random.RandomState(3).randn(N, 19937))
rs = np.random.RandomState(4)
sm_raw = (rs.rand(N, 1686) < 0.007) * rs.choice([1, 2, 3], size=(N, 1686), p=[0.985, 0.0146, 0.0004])
sm_raw[0] += (sm_raw.sum(0) == 0)          # no zero-variance columns (the reference drops them, src/utils.py:97)
sm = zs(sm_raw.astype(np.float64))
sc = lambda xs: [max(2, int(round(x * args.scale))) for x in xs]
specs = {
    "ME": dict(data=me, layers_sizes=[40], pretraining_epochs=sc([80000]), pretrain_lr=[0.005], k=10, lambda_1=0.01, lambda_2=0.01),
    "GE": dict(data=ge, layers_sizes=[400, 40], pretraining_epochs=sc([8000, 800]), pretrain_lr=[0.005, 0.1], k=1, lambda_1=0.01, lambda_2=0.1),
    "SM": dict(data=sm, layers_sizes=[200, 20], pretraining_epochs=sc([8000, 800]), pretrain_lr=[0.005, 0.1], k=1, lambda_1=0.01, lambda_2=0.01),
}
np.random.seed(20161230 + rank)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
dbns, joint, top = train_modalities(specs, batch_size=20, top=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
calls = {n: [h["calls"] for h in d.history] for n, d in dbns.items()}
if world > 1:
    allc = [None] * world
    dist.all_gather_object(allc, calls)
    calls = {k: v for c in allc for k, v in c.items()}
if rank == 0:
    calls["top"] = [h["calls"] for h in top.history]
    print(json.dumps({"metric": "MDBN pretrain wall-clock", "value": dt, "unit": "s", "n_gpus": world, "higher_is_better": False,
                      "scale": args.scale, "cd_steps_per_layer": calls,
                      "config": "AML-shaped synthetic: ME 559->40 (k=10), GE 19937->400->40, SM 1686->200->20, joint 100->24->3, N=170, batch 20",
                      "parallelism": "one modality DBN per GPU (round-robin), joint DBN on rank 0"}), flush=True)
if world > 1:
    dist.destroy_process_group()
