"""torchrun --nproc-per-node 2 scripts/dp_check.py : NCCL data-parallel CD step (mdbn_cd_args.comm; DP_HOST=1 for the
host-level torch.distributed form) == single-GPU step,
and modality-parallel MDBN pretraining == sequential (same weights)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import mdbn_b200 as M
from mdbn_b200.parallel import DataParallel, train_modalities
from oracle import rbm_oracle as O, shared_u

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
V, H, B, k = 784, 500, 256, 1
kind = O.RBM
data = (np.random.RandomState(0).rand(B, V) < 0.13).astype(np.float32)
W0 = O.init_W(np.random.RandomState(1), V, H).astype(np.float32)
U = shared_u.step_buffer(8, 0, 0, kind, True, B, V, H, k)
lay, _ = O.u_layout(kind, True, B, V, H, k)
per = B // world
lo, hi = rank * per, (rank + 1) * per
Uloc = np.concatenate([U[o:o + sh[0] * sh[1]].reshape(sh)[lo:hi].ravel() for _, o, sh in lay])
for path, tf32 in (("generic", False), ("tensor", True)):
    r = M.RBM(n_visible=V, n_hidden=H, W=W0, theano_rng=M.BufferStreams(lambda l, c, b: Uloc))
    cost, upd = r.get_cost_updates(lr=0.1, k=k, weightcost=0.0002, batch_size=B)
    fn = r.make_train_fn(data, cost, upd, path=path, tf32=tf32)
    fn.dp = DataParallel(c_abi=os.environ.get("DP_HOST") is None)      # default: the communicator inside libmdbn_b200.so
    c = fn(np.arange(B, dtype=np.int32), 0.5)
    # reference: the whole minibatch on this GPU alone
    r1 = M.RBM(n_visible=V, n_hidden=H, W=W0, theano_rng=M.BufferStreams(lambda l, c, b: U))
    cost1, upd1 = r1.get_cost_updates(lr=0.1, k=k, weightcost=0.0002, batch_size=B)
    fn1 = r1.make_train_fn(data, cost1, upd1, path=path, tf32=tf32)
    c1 = fn1(np.arange(B, dtype=np.int32), 0.5)
    err = np.abs(r.W_speed.get_value() - r1.W_speed.get_value()).max() / np.abs(r1.W_speed.get_value()).max()
    print("[rank %d] %s dp cost %.5f single %.5f  W_speed rel err %.2e" % (rank, path, c, c1, err), flush=True)
    assert err < (5e-3 if tf32 else 1e-5) and abs(c - c1) < 1e-3 * abs(c1)
# modality-parallel
rs = np.random.RandomState(3)
specs = {"ME": dict(data=rs.randn(40, 64).astype(np.float32), layers_sizes=[16], pretraining_epochs=[40], pretrain_lr=[0.005], k=2, lambda_1=0.01, lambda_2=0.01),
         "GE": dict(data=rs.randn(40, 256).astype(np.float32), layers_sizes=[32, 16], pretraining_epochs=[40, 20], pretrain_lr=[0.005, 0.1], k=1, lambda_1=0.01, lambda_2=0.1)}
np.random.seed(7)
dbns, joint, top = train_modalities(specs, batch_size=10, top=False)
print("[rank %d] trained %s joint %s" % (rank, sorted(dbns), None if joint is None else joint.shape), flush=True)
dist.barrier()
dist.destroy_process_group()
