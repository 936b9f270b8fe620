import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mdbn_b200 as M
np.set_printoptions(linewidth=200, precision=3, suppress=True)
B, V, H = 128, 256, 128
rs = np.random.RandomState(0)
W = np.zeros((V, H), np.float32)
W[:] = (np.arange(V)[:, None] * 1000 + np.arange(H)[None, :]).astype(np.float32)   # W[i,j] = 1000 i + j
r = M.RBM(n_visible=V, n_hidden=H, W=W)
r.ctx.set_tf32_phases(True)
# propup with one-hot rows: v[b] = e_{b} -> pre[b, :] = W[b, :]
v = np.zeros((B, V), np.float32); v[np.arange(B), np.arange(B)] = 1
pre, mean = r.propup(v)
pre = pre.cpu().numpy()
print("propup one-hot: max err", np.abs(pre - W[:B]).max())
print(pre[:4, :8]); print(pre[8:10, :8]); print(pre[:2, 30:40])
# propdown with one-hot: h[b] = e_b -> pre[b, :] = W[:, b]
h = np.zeros((B, H), np.float32); h[np.arange(B), np.arange(B) % H] = 1
pre2, _ = r.propdown(h)
pre2 = pre2.cpu().numpy()
print("propdown one-hot: max err", np.abs(pre2 - W[:, :B].T[:B]).max())
print(pre2[:4, :8]); print(pre2[:2, 30:40])
