"""us per CD step for batches between the skinny kernel (B <= 20) and the large-batch regime, GRBM 19937->400 CD-1:
generic fp32 path vs tensor (TF32) path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mdbn_b200 as M
V, H = 19937, 400
data = torch.from_numpy(np.random.RandomState(0).randn(256, V).astype(np.float32)).cuda()
for B in (20, 32, 50, 64, 100, 128):
    for path, tf32 in (("auto", False), ("tensor", True)):
        if path == "tensor" and B % 32:
            continue
        r = M.GRBM(n_visible=V, n_hidden=H, numpy_rng=np.random.RandomState(1), theano_rng=M.RandomStreams(2))
        cost, upd = r.get_cost_updates(lr=0.005, k=1, lambda_1=0.01, lambda_2=0.1, batch_size=B)
        f = r.make_train_fn(data, cost, upd, path=path, tf32=tf32)
        f.sync = False
        idx = torch.arange(B, dtype=torch.int32).cuda()
        for _ in range(5):
            f(idx, 0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50):
            f(idx, 0.0)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        ab = 4 * (7 * V * H + B * V + 3 * (V + H) + 4 * (V + H))
        print(json.dumps({"B": B, "path": path, "us_per_step": round(us, 1), "samples_per_s": round(B / us * 1e6), "hbm_frac": round(ab / (us * 1e-6) / 1e9 / 6524.3, 3)}), flush=True)
