"""us per CD step for batches between the W-streaming kernel (B <= 20) and the large-batch regime:
GRBM 19937->400 and RBM 784->500, CD-1 / PCD-1, the tcgen05 path in its fp32-exact split-TF32 mode (path=auto),
in plain TF32 (tf32=True) and the generic SIMT path it replaced.  Launches per step from the library's counter."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mdbn_b200 as M

HBM = 6524.3
LAYERS = {"ge": (M.GRBM, 19937, 400, dict(lr=0.005, lambda_1=0.01, lambda_2=0.1)),
          "mnist": (M.RBM, 784, 500, dict(lr=0.1, weightcost=0.0002))}
names = sys.argv[1:] or ["ge", "mnist"]
for name in names:
    cls, V, H, kw = LAYERS[name]
    data = torch.from_numpy(np.random.RandomState(0).randn(512, V).astype(np.float32)).cuda()
    if cls is M.RBM:
        data = (data > 1.0).float()
    for B in (20, 21, 32, 50, 64, 100, 128):
        for pcd in (False, True):
            for path, tf32 in (("auto", False), ("tensor", True), ("generic", False)):
                if B == 20 and path != "auto":
                    continue
                R = 4 if V * H > 4e6 else 1     # rotate parameter sets on the big layer: weights start in HBM
                fns = []
                for i in range(R):
                    r = cls(n_visible=V, n_hidden=H, numpy_rng=np.random.RandomState(1 + i), theano_rng=M.RandomStreams(2 + i))
                    P = M.shared(np.zeros((B, H), np.float32)) if pcd else None
                    cost, upd = r.get_cost_updates(k=1, batch_size=B, persistent=P, **kw)
                    f = r.make_train_fn(data, cost, upd, path=path, tf32=tf32)
                    f.sync = False
                    fns.append(f)
                idx = torch.arange(B, dtype=torch.int32).cuda()
                for i in range(2 * R):
                    fns[i % R](idx, 0.0)
                n = 48
                l0 = fns[0].rbm.ctx.launches
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                for i in range(n):
                    fns[i % R](idx, 0.0)
                e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / n
                ab = 4 * (7 * V * H + B * V + 3 * (V + H) + 4 * (V + H))
                print(json.dumps({"layer": "%d->%d" % (V, H), "B": B, "pcd": pcd, "path": path + ("+tf32" if tf32 else ""),
                                  "us_per_step": round(us, 1), "samples_per_s": round(B / us * 1e6),
                                  "hbm_frac": round(ab / (us * 1e-6) / 1e9 / HBM, 3),
                                  "launches_per_step": (fns[0].rbm.ctx.launches - l0) / n}), flush=True)
                del fns
