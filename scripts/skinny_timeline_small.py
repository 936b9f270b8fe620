import os, sys
sys.path.insert(0, "/root/repo")
os.environ["MDBN_SKINNY_TIMING"] = "1"; os.environ["MDBN_TINY_TIMING"] = "1"
import numpy as np, torch
import mdbn_b200 as M
for (cls, V, H, B, k) in ((M.RBM, 100, 24, 20, 1), (M.RBM, 400, 40, 20, 1), (M.GRBM, 559, 40, 20, 10), (M.GRBM, 1686, 200, 20, 1)):
    data = np.random.RandomState(0).randn(170, V).astype(np.float32)
    r = cls(n_visible=V, n_hidden=H, theano_rng=M.RandomStreams(1))
    cost, upd = r.get_cost_updates(lr=0.005, k=k, lambda_1=0.01, lambda_2=0.1, batch_size=B)
    fn = r.make_train_fn(data, cost, upd)
    for t in range(4):
        fn(np.arange(B, dtype=np.int32), 0.0)
