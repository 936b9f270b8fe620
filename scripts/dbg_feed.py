import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mdbn_b200 as m
V, H, B, n_mb = 300, 72, 10, 5
data = np.random.RandomState(33).randn(B * n_mb, V).astype(np.float32)
W0 = (np.random.RandomState(9).randn(V, H) * 0.01).astype(np.float32)
def make(dataset):
    r = m.GRBM(n_visible=V, n_hidden=H, W=W0, theano_rng=m.RandomStreams(77))
    P = m.shared(np.zeros((B, H), np.float32))
    cost, upd = r.get_cost_updates(lr=0.05, k=1, lambda_1=0.01, lambda_2=0.1, batch_size=B, persistent=P)
    return r, r.make_train_fn(dataset, cost, upd)
r_dev, fn_dev = make(data)
r_host, fn_host = make(np.zeros((B, V), np.float32))
r_dev2, fn_dev2 = make(data)
host = [torch.from_numpy(np.ascontiguousarray(data[i * B:(i + 1) * B])).pin_memory() for i in range(n_mb)]
for i in range(n_mb):
    c_dev = fn_dev(np.arange(i * B, (i + 1) * B, dtype=np.int32), 0.5)
    c_dev2 = fn_dev2(np.arange(i * B, (i + 1) * B, dtype=np.int32), 0.5)
    nxt = host[i + 1] if (i + 1 < n_mb and i % 2 == 0) else None
    c_host = fn_host.step_from_host(host[i], 0.5, next_host_batch=nxt)
    print(i, c_dev, c_dev2, c_host, np.abs(r_dev.W.get_value() - r_host.W.get_value()).max(), np.abs(r_dev.W.get_value() - r_dev2.W.get_value()).max(),
          int(r_dev.bit_i_idx.item()), int(r_host.bit_i_idx.item()))
