"""us per CD step on the small layers of the AML configuration (single launches and chained launches)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mdbn_b200 as M
N = 170
for (name, cls, V, H, B, k) in (("ME 559->40 k=10", M.GRBM, 559, 40, 20, 10), ("SM 1686->200", M.GRBM, 1686, 200, 20, 1),
                                ("GE2 400->40", M.RBM, 400, 40, 20, 1), ("top 100->24", M.RBM, 100, 24, 20, 1),
                                ("GE 19937->400", M.GRBM, 19937, 400, 20, 1)):
    data = np.random.RandomState(0).randn(N, V).astype(np.float32)
    r = cls(n_visible=V, n_hidden=H, theano_rng=M.RandomStreams(1))
    cost, upd = r.get_cost_updates(lr=0.005, k=k, lambda_1=0.01, lambda_2=0.1, batch_size=B)
    fn = r.make_train_fn(data, cost, upd); fn.sync = False
    idx = torch.arange(8 * B, dtype=torch.int32, device="cuda").view(8, B) % N
    for _ in range(20): fn(idx[0], 0.0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 2000
    for s in range(n): fn(idx[s & 7], 0.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for s in range(n // 8): fn.run_steps(idx, 0.0)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%-18s single %.1f us/step   chained(8) %.1f us/step" % (name, (t1 - t0) / n * 1e6, (t2 - t1) / n * 1e6))
