import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mdbn_b200 as M
V, H, B, k = 559, 40, 20, 10
data = np.random.RandomState(0).randn(170, V).astype(np.float32)
r = M.GRBM(n_visible=V, n_hidden=H, theano_rng=M.RandomStreams(1))
cost, upd = r.get_cost_updates(lr=0.005, k=k, lambda_1=0.01, lambda_2=0.01, batch_size=B)
fn = r.make_train_fn(data, cost, upd); fn.sync = False
idx = torch.arange(8 * B, dtype=torch.int32, device="cuda").view(8, B) % 170
for _ in range(6): fn.run_steps(idx, 0.0)
torch.cuda.synchronize()
