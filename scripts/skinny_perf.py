"""us per CD step of the persistent skinny kernel: single-step launches and chained launches (an epoch per
launch, TrainFn.run_steps), both with 4 parameter sets in rotation BETWEEN launches (255 MB > L2: the first
sweep of a launch streams from HBM), plus the L2-warm single-set figures.  CUDA events, after warm-up."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mdbn_b200 as M

CASES = {
    "ge_b10_pcd1": (M.GRBM, 19937, 400, 10, 1, True, dict(lr=0.005, lambda_1=0.01, lambda_2=0.1)),
    "ge_b20_cd1": (M.GRBM, 19937, 400, 20, 1, False, dict(lr=0.005, lambda_1=0.01, lambda_2=0.1)),
    "ge_b10_cd1": (M.GRBM, 19937, 400, 10, 1, False, dict(lr=0.005, lambda_1=0.01, lambda_2=0.1)),
    "ge_b10_pcd5": (M.GRBM, 19937, 400, 10, 5, True, dict(lr=0.005, lambda_1=0.01, lambda_2=0.1)),
    "mnist_b20_cd1": (M.RBM, 784, 500, 20, 1, False, dict(lr=0.1, weightcost=0.0002)),
    "dbn1000_b20_cd1": (M.RBM, 1000, 1000, 20, 1, False, dict(lr=0.01, weightcost=0.0002)),
    "mnist_b10_cd1": (M.RBM, 784, 500, 10, 1, False, dict(lr=0.1, weightcost=0.0002)),
    "dbn784x1000_b10_cd1": (M.RBM, 784, 1000, 10, 1, False, dict(lr=0.01, weightcost=0.0002)),
    "dbn1000_b10_cd1": (M.RBM, 1000, 1000, 10, 1, False, dict(lr=0.01, weightcost=0.0002)),
    "dbn1000_b10_pcd5": (M.RBM, 1000, 1000, 10, 5, True, dict(lr=0.01, weightcost=0.0002)),
    "sm_b20_cd1": (M.GRBM, 1686, 200, 20, 1, False, dict(lr=0.005, lambda_1=0.01, lambda_2=0.01)),
}
names = sys.argv[1:] or ["ge_b10_pcd1", "ge_b20_cd1"]
HBM = 6524.3


def timeit(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


for name in names:
    cls, V, H, B, k, pcd, kw = CASES[name]
    N = 170 if V > 5000 else 4000
    data = torch.from_numpy(np.random.RandomState(0).randn(N, V).astype(np.float32)).cuda()
    if cls is M.RBM:
        data = (data > 1.0).float()
    R = 4
    fns = []
    for i in range(R):
        r = cls(n_visible=V, n_hidden=H, numpy_rng=np.random.RandomState(123 + i), theano_rng=M.RandomStreams(1 + i))
        P = M.shared(np.zeros((B, H), np.float32)) if pcd else None
        cost, upd = r.get_cost_updates(k=k, batch_size=B, persistent=P, **kw)
        f = r.make_train_fn(data, cost, upd, path=os.environ.get("MDBN_PATH", "auto"))
        f.sync = False
        fns.append(f)
    n_mb = N // B
    chain = min(n_mb, 17)
    perm = torch.from_numpy(np.random.RandomState(5).permutation(N)[: chain * B].astype(np.int32)).cuda()
    idx_mat = perm.view(chain, B)
    mom = 0.0 if cls is M.GRBM else 0.9

    def single(n, rot=True):
        for s in range(n):
            fns[s % R if rot else 0](idx_mat[s % chain], mom)

    def chained(n, rot=True):
        for l in range(n):
            fns[l % R if rot else 0].run_steps(idx_mat, mom)
    single(8); chained(8)
    abytes = 4 * ((2 * k + 5) * V * H + B * V + (2 * k + 1) * (V + H) + 4 * (V + H))
    out = {"case": name}
    for label, f, n, per in (("single_cold", lambda n: single(n), 200, 1), ("single_warm", lambda n: single(n, False), 200, 1),
                             ("chain_cold", lambda n: chained(n), 24, chain), ("chain_warm", lambda n: chained(n, False), 24, chain)):
        best = min(timeit(f, n) for _ in range(3)) / (n * per)
        out[label + "_us"] = round(best, 2)
        out[label + "_frac"] = round(abytes / (best * 1e-6) / 1e9 / HBM, 3)
    print(json.dumps(out), flush=True)
