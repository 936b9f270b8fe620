"""Phase timeline of the persistent skinny kernel (MDBN_SKINNY_TIMING=1 makes CTA 0 stamp %globaltimer).
COLD=1 rotates over 4 independent layers (255 MB of W + speed > L2) so that weights stream from HBM."""
import os, sys
sys.path.insert(0, "/root/repo")
os.environ["MDBN_SKINNY_TIMING"] = "1"
import numpy as np, torch
import mdbn_b200 as M
nrep = 4 if os.environ.get("COLD") else 1
for (cls, V, H, B, pcd, kw) in ((M.GRBM, 19937, 400, 10, True, dict(lr=0.005, lambda_1=0.01, lambda_2=0.1)),
                                (M.GRBM, 19937, 400, 20, False, dict(lr=0.005, lambda_1=0.01, lambda_2=0.1)),
                                (M.RBM, 784, 500, 20, False, dict(lr=0.1, weightcost=0.0002))):
    data = np.random.RandomState(0).randn(170, V).astype(np.float32)
    fns = []
    for i in range(nrep):
        r = cls(n_visible=V, n_hidden=H, theano_rng=M.RandomStreams(1 + i))
        P = M.shared(np.zeros((B, H), np.float32)) if pcd else None
        cost, upd = r.get_cost_updates(k=1, batch_size=B, persistent=P, **kw)
        fns.append(r.make_train_fn(data, cost, upd))
    for t in range(4 * nrep):
        fns[t % nrep](np.arange(B, dtype=np.int32), 0.0)
    torch.cuda.synchronize()
