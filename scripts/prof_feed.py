import sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mdbn_b200 as M
V, H, B = 19937, 400, 10
r = M.GRBM(n_visible=V, n_hidden=H, theano_rng=M.RandomStreams(1))
P = M.shared(np.zeros((B, H), np.float32))
cost, upd = r.get_cost_updates(lr=0.005, k=1, lambda_1=0.01, lambda_2=0.1, batch_size=B, persistent=P)
fn = r.make_train_fn(np.zeros((B, V), np.float32), cost, upd)
host = [torch.randn(B, V).pin_memory() for _ in range(8)]
def loop(n):
    for s in range(n):
        fn.step_from_host(host[s % 8], 0.0, next_host_batch=host[(s + 1) % 8], lag=1)
    fn.flush()
loop(50); torch.cuda.synchronize()
t0 = time.perf_counter(); loop(500); torch.cuda.synchronize(); print("us/step", (time.perf_counter() - t0) / 500 * 1e6)
pr = cProfile.Profile(); pr.enable(); loop(300); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
