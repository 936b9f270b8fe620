"""A few chained launches (17 CD steps each) of the skinny kernel for an ncu capture:
   ncu --set full --import-source on -k regex:cd_skinny -s 2 -c 1 -o gpurun_out/chain python scripts/ncu_chain.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mdbn_b200 as M
V, H, B, chain = 19937, 400, int(os.environ.get("B", "10")), 17
pcd = os.environ.get("PCD", "1") == "1"
data = torch.from_numpy(np.random.RandomState(0).randn(chain * B, V).astype(np.float32)).cuda()
r = M.GRBM(n_visible=V, n_hidden=H, numpy_rng=np.random.RandomState(123), theano_rng=M.RandomStreams(1))
P = M.shared(np.zeros((B, H), np.float32)) if pcd else None
cost, upd = r.get_cost_updates(lr=0.005, k=1, lambda_1=0.01, lambda_2=0.1, batch_size=B, persistent=P)
f = r.make_train_fn(data, cost, upd)
f.sync = False
idx = torch.arange(chain * B, dtype=torch.int32).cuda().view(chain, B)
for _ in range(4):
    f.run_steps(idx, 0.0)
torch.cuda.synchronize()
print("ok")
