"""Host-side logic of the multi-GPU paths on CPU: world_size-2 gloo process groups.

The compute kernels need a B200, so the per-rank step is injected: the oracle's cd_stats / cd_apply
stand in for the STATS / APPLY phases of mdbn_cd_step.  What is under test is the product's host
logic in mdbn_b200/parallel.py: row sharding, the packed-statistics all-reduce and its layout, the
rows bookkeeping, modality placement and the RNG replay that keeps initial weights identical."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rbm_oracle as O
from oracle import shared_u


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, kind, pcd, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mdbn_b200.parallel import DataParallel, shard_rows, stats_size
    V, H, B, k = 23, 11, 12, 2
    rs = np.random.RandomState(5)
    data = rs.randn(40, V) if kind == O.GRBM else (rs.rand(40, V) < 0.4).astype(np.float64)
    L = O.Layer(V, H, kind, numpy_rng=np.random.RandomState(9))
    snap = L.W.copy()
    P_full = np.zeros((B, H))
    dp = DataParallel()
    assert stats_size(V, H) == V * H + H + V + 2
    for step in range(3):
        idx = np.arange(step * B, (step + 1) * B)
        U = shared_u.step_buffer(3, 0, step, kind, True, B, V, H, k)
        lay, _ = O.u_layout(kind, True, B, V, H, k)
        rows_of = {r: shard_rows(list(range(B)), r, world) for r in range(world)}

        def stats_fn(mine):
            loc = [int(i) - step * B for i in mine]
            # this rank's rows of every segment of the shared random buffer
            Uloc = np.concatenate([U[o:o + sh[0] * sh[1]].reshape(sh)[loc].ravel() for _, o, sh in lay])
            Ploc = P_full[loc] if pcd else None
            buf = O.cd_stats(L, data[list(mine)], Uloc, k=k, persistent=Ploc)
            if pcd:
                P_full[loc] = Ploc
            return torch.from_numpy(buf.copy())

        def apply_fn(buf, rows_total):
            assert rows_total == B
            return O.cd_apply(L, buf.numpy(), lr=0.05, lambda_1=0.01, lambda_2=0.1, weightcost=0.0002, batch_size=B,
                              momentum=0.5, W_snap=snap, pcd=pcd)
        cost = dp.step(idx, stats_fn, apply_fn)
    if rank == 0:
        out.put((L.W.copy(), L.hbias.copy(), L.vbias.copy(), float(cost)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,pcd", [(O.RBM, False), (O.GRBM, False), (O.RBM, True)])
def test_data_parallel_step_equals_single_process(kind, pcd):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, kind, pcd, q)) for r in range(world)]
    for p in procs:
        p.start()
    W, hb, vb, cost = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, whole minibatch
    V, H, B, k = 23, 11, 12, 2
    rs = np.random.RandomState(5)
    data = rs.randn(40, V) if kind == O.GRBM else (rs.rand(40, V) < 0.4).astype(np.float64)
    L = O.Layer(V, H, kind, numpy_rng=np.random.RandomState(9))
    snap = L.W.copy()
    P = np.zeros((B, H)) if pcd else None
    for step in range(3):
        U = shared_u.step_buffer(3, 0, step, kind, True, B, V, H, k)
        c = O.cd_step(L, data[step * B:(step + 1) * B], U, lr=0.05, k=k, lambda_1=0.01, lambda_2=0.1,
                      weightcost=0.0002, batch_size=B, momentum=0.5, persistent=P, W_snap=snap)
    np.testing.assert_allclose(W, L.W, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(hb, L.hbias, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(vb, L.vbias, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(cost, c, rtol=1e-10)


def test_shard_rows_and_placement():
    from mdbn_b200.parallel import shard_rows, place_modalities
    idx = list(range(10))
    parts = [shard_rows(idx, r, 4) for r in range(4)]
    assert sum(parts, []) == idx and [len(p) for p in parts] == [3, 3, 2, 2]
    assert shard_rows(idx, 0, 1) == idx
    assert place_modalities(["ME", "GE", "SM"], 8) == {"ME": 0, "GE": 1, "SM": 2}
    assert place_modalities(["ME", "GE", "SM"], 2) == {"ME": 0, "GE": 1, "SM": 0}


def test_modality_rng_replay_matches_sequential_construction():
    """Ranks that do not own a modality must consume exactly the draws DBN.__init__ would
    (src/dbn.py:114,155-159), so that every DBN starts from the weights of a sequential run."""
    dims = {"ME": [15, 6], "GE": [31, 10, 6], "SM": [20, 8, 4]}
    seq = np.random.RandomState(123)
    ref = {}
    for n, d in dims.items():
        seq.randint(2 ** 30)
        ref[n] = [O.init_W(seq, a, b) for a, b in zip(d[:-1], d[1:])]
    for owner in dims:
        rng = np.random.RandomState(123)
        for n, d in dims.items():
            if n == owner:
                rng.randint(2 ** 30)
                got = [O.init_W(rng, a, b) for a, b in zip(d[:-1], d[1:])]
                for g, r in zip(got, ref[n]):
                    np.testing.assert_array_equal(g, r)
            else:
                rng.randint(2 ** 30)
                for a, b in zip(d[:-1], d[1:]):
                    rng.uniform(size=(a, b))
