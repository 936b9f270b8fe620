"""A few CD steps of one mid-batch configuration (for an ncu launch list / --set full capture):
python scripts/ncu_mid.py <ge|mnist> <B> <pcd 0|1> [tf32 0|1] [k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mdbn_b200 as M
name, B, pcd = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
tf32 = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
k = int(sys.argv[5]) if len(sys.argv) > 5 else 1
cls, V, H, kw = {"ge": (M.GRBM, 19937, 400, dict(lr=0.005, lambda_1=0.01, lambda_2=0.1)),
                 "mnist": (M.RBM, 784, 500, dict(lr=0.1, weightcost=0.0002))}[name]
data = torch.from_numpy(np.random.RandomState(0).randn(max(512, B), V).astype(np.float32)).cuda()
if cls is M.RBM:
    data = (data > 1.0).float()
r = cls(n_visible=V, n_hidden=H, numpy_rng=np.random.RandomState(1), theano_rng=M.RandomStreams(2))
P = M.shared(np.zeros((B, H), np.float32)) if pcd else None
cost, upd = r.get_cost_updates(k=k, batch_size=B, persistent=P, **kw)
f = r.make_train_fn(data, cost, upd, path="auto", tf32=tf32)
f.sync = False
idx = torch.arange(B, dtype=torch.int32).cuda()
for _ in range(3):
    f(idx, 0.0)
torch.cuda.synchronize()
