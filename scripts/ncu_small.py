"""A few chained launches (8 CD steps each) of one small / medium layer for an ncu capture of cd_mid_kernel / cd_tiny_kernel:
python scripts/ncu_small.py <mnist|dbn1000|sm|me>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mdbn_b200 as M
name = sys.argv[1]
cls, V, H, B, k, kw = {"mnist": (M.RBM, 784, 500, 20, 1, dict(lr=0.1, weightcost=0.0002)),
                       "dbn1000": (M.RBM, 1000, 1000, 20, 1, dict(lr=0.01, weightcost=0.0002)),
                       "sm": (M.GRBM, 1686, 200, 20, 1, dict(lr=0.005, lambda_1=0.01, lambda_2=0.01)),
                       "me": (M.GRBM, 559, 40, 20, 10, dict(lr=0.005, lambda_1=0.01, lambda_2=0.01))}[name]
data = torch.from_numpy(np.random.RandomState(0).randn(512, V).astype(np.float32)).cuda()
if cls is M.RBM:
    data = (data > 1.0).float()
r = cls(n_visible=V, n_hidden=H, numpy_rng=np.random.RandomState(1), theano_rng=M.RandomStreams(2))
cost, upd = r.get_cost_updates(k=k, batch_size=B, **kw)
f = r.make_train_fn(data, cost, upd)
f.sync = False
idx = torch.arange(8 * B, dtype=torch.int32).cuda().view(8, B)
for _ in range(4):
    f.run_steps(idx, 0.0)
torch.cuda.synchronize()
print("ok")
