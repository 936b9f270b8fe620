"""Phase timeline of the broadcast kernel for medium layers (MDBN_MID_TIMING=1 makes CTA 0 stamp %globaltimer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MDBN_MID_TIMING"] = "1"
import numpy as np, torch
import mdbn_b200 as M
for (cls, V, H, B, pcd, kw) in ((M.RBM, 784, 500, 20, False, dict(lr=0.1, weightcost=0.0002)),
                                (M.GRBM, 1686, 200, 20, False, dict(lr=0.005, lambda_1=0.01, lambda_2=0.01)),
                                (M.RBM, 1000, 1000, 20, False, dict(lr=0.01, weightcost=0.0002))):
    data = np.random.RandomState(0).randn(200, V).astype(np.float32)
    r = cls(n_visible=V, n_hidden=H, theano_rng=M.RandomStreams(1))
    P = M.shared(np.zeros((B, H), np.float32)) if pcd else None
    cost, upd = r.get_cost_updates(k=1, batch_size=B, persistent=P, **kw)
    fn = r.make_train_fn(data, cost, upd, path="mid")
    for t in range(4):
        fn(np.arange(B, dtype=np.int32), 0.0)
    torch.cuda.synchronize()
