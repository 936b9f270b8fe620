"""Wall-clock of the ME DBN alone (559->40, k = 10, 100 k CD steps) against its kernel time: is the epoch loop host-bound?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mdbn_b200.parallel import train_modalities, aml_synthetic_specs
specs = {"ME": aml_synthetic_specs(1.0)["ME"]}
warm = {"ME": aml_synthetic_specs(0.002)["ME"]}
train_modalities(warm, batch_size=20, top=False)
np.random.seed(20161230)
torch.cuda.synchronize(); t0 = time.perf_counter()
dbns = train_modalities(specs, batch_size=20, top=False)[0]
torch.cuda.synchronize(); dt = time.perf_counter() - t0
calls = dbns["ME"].history[0]["calls"]
print("ME alone: %.3f s for %d steps = %.1f us per step wall (kernel: 22.9 us per chained step)" % (dt, calls, dt / calls * 1e6))
