import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mdbn_b200 as m
def run(kind, pcd, V, H, B, steps, seed=77, path="auto"):
    data = np.random.RandomState(33).randn(B * steps, V).astype(np.float32)
    if not kind: data = (data > 0).astype(np.float32)
    W0 = (np.random.RandomState(9).randn(V, H) * 0.01).astype(np.float32)
    cls = m.GRBM if kind else m.RBM
    r = cls(n_visible=V, n_hidden=H, W=W0, theano_rng=m.RandomStreams(seed))
    P = m.shared(np.zeros((B, H), np.float32)) if pcd else None
    cost, upd = r.get_cost_updates(lr=0.05, k=1, lambda_1=0.01, lambda_2=0.1, batch_size=B, persistent=P)
    fn = r.make_train_fn(data, cost, upd, path=path)
    out = []
    for i in range(steps):
        c = fn(np.arange(i * B, (i + 1) * B, dtype=np.int32), 0.5)
        out.append((c, r.W_speed.get_value().copy(), r.hbias_speed.get_value().copy(), r.vbias_speed.get_value().copy(),
                    P.get_value().copy() if pcd else None))
    return out
for path in ("skinny", "generic"):
  for (kind, pcd, V, H, B) in ((1, True, 300, 72, 10), (0, False, 300, 72, 10), (1, False, 19937, 400, 20), (1, False, 19937, 400, 10)):
    a = run(kind, pcd, V, H, B, 1, path=path)
    for rep in range(3):
        b = run(kind, pcd, V, H, B, 1, path=path)
        x, y = a[0], b[0]
        print(path, kind, pcd, V, H, B, "cost", x[0] == y[0], x[0], y[0], "S diff", int((x[1] != y[1]).sum()), "hbS", int((x[2] != y[2]).sum()),
              "vbS", int((x[3] != y[3]).sum()), "P", int((x[4] != y[4]).sum()) if pcd else None)
