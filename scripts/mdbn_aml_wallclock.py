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
ap.add_argument("--cold", action="store_true", help="no warm-up run: include context set-up and lazy module loading")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from mdbn_b200.parallel import train_modalities

from mdbn_b200.parallel import aml_synthetic_specs
specs = aml_synthetic_specs(args.scale)
if not args.cold:
    # warm-up, untimed (as bench.py does for every leg): the same run with 0.2 % of the patience budgets loads the kernels
    # of every layer shape on every rank; --cold times a fresh process including the lazy module loading (~0.5 s)
    train_modalities(aml_synthetic_specs(0.002 * args.scale), batch_size=20, top=True)
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
                      "scale": args.scale, "warmup": "none (cold process)" if args.cold else "scaled-down run (0.2 % of the budgets), untimed",
                      "cd_steps_per_layer": calls,
                      "config": "AML-shaped synthetic: ME 559->40 (k=10), GE 19937->400->40, SM 1686->200->20, joint 100->24->3, N=170, batch 20",
                      "parallelism": "one modality DBN per GPU (round-robin), joint DBN on rank 0"}), flush=True)
if world > 1:
    dist.destroy_process_group()
