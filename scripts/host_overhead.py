"""Host time per train-fn call versus device time per step (is the step loop launch-bound?)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mdbn_b200 as M
V, H, B = 19937, 400, 10
data = torch.from_numpy(np.random.RandomState(0).randn(4096, V).astype(np.float32)).cuda()
fns = []
for i in range(4):
    r = M.GRBM(n_visible=V, n_hidden=H, theano_rng=M.RandomStreams(1 + i))
    P = M.shared(np.zeros((B, H), np.float32))
    cost, upd = r.get_cost_updates(lr=0.005, k=1, lambda_1=0.01, lambda_2=0.1, batch_size=B, persistent=P)
    fn = r.make_train_fn(data, cost, upd)
    fn.sync = False
    fns.append(fn)
idx = torch.arange(B, dtype=torch.int32, device="cuda")
for n in (200, 2000):
    for f in fns: f(idx, 0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(n):
        fns[s & 3](idx, 0.0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("n=%d host %.1f us/call, total %.1f us/step" % (n, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
