"""Multi-GPU use of the path, only where it shards naturally:

1. data parallelism inside one CD step — every rank runs the positive phase and the Gibbs chain on
   its slice of the minibatch rows with replicated parameters; the sufficient statistics are sums
   over rows (src/rbm.py:411-417), so ONE all-reduce(sum) of the packed buffer
   [v0^T ph - nv^T nh (V*H) | sum(ph-nh) (H) | sum(v0-nv) (V) | cost numerator | rows]
   per step, after which every rank applies the identical in-place update;
2. modality parallelism — the per-modality DBNs of an MDBN are independent until their top
   activations are concatenated (src/AMLsm.py:38-83): one GPU per modality, one tensor collective of the
   [N, H_top] activations to the rank that trains the joint DBN (src/MDBN.py:31-42).

torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is the plumbing."""
import numpy
import torch
import torch.distributed as dist


def shard_rows(indexes, rank, world):
    """Contiguous, balanced slice of a minibatch's row indices for `rank`."""
    n = len(indexes)
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return indexes[lo:hi]


def stats_size(V, H):
    return V * H + H + V + 2


class DataParallel:
    """Data-parallel CD step.  Two forms:

    * inside the library (default on GPUs): an NCCL communicator owned by libmdbn_b200.so (mdbn_comm_init); TrainFn
      passes it in mdbn_cd_args.comm and ONE mdbn_cd_step does shard statistics -> chunked all-reduce -> update;
    * host-level (`c_abi=False`, and the CPU tests): `stats_fn(rows) -> buffer` fills this rank's packed buffer,
      torch.distributed all-reduces it, `apply_fn(buffer, rows_total)` applies the update."""

    def __init__(self, group=None, c_abi=None, device=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.comm = None
        if c_abi is None:
            c_abi = torch.cuda.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl"
        if c_abi:
            from . import _lib
            dev = torch.cuda.current_device() if device is None else int(device)

            def exchange(uid):
                if self.world == 1:
                    return uid
                box = [uid]
                dist.broadcast_object_list(box, src=0, group=group)
                return box[0]
            self.comm = _lib.Comm(self.rank, self.world, dev, exchange)

    def all_reduce(self, buf):
        if self.world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf

    def step(self, indexes, stats_fn, apply_fn):
        mine = shard_rows(indexes, self.rank, self.world)
        buf = stats_fn(mine)
        self.all_reduce(buf)
        rows_total = int(round(float(buf[-1])))
        assert rows_total == len(indexes), "ranks disagree on the minibatch (%d vs %d rows)" % (rows_total, len(indexes))
        return apply_fn(buf, rows_total)


def place_modalities(names, world):
    """modality -> rank.  Round-robin in the given order; with fewer GPUs than modalities several
    modalities share a rank (they then train one after the other, like the reference)."""
    return {n: i % world for i, n in enumerate(names)}


def train_modalities(specs, rank=None, world=None, group=None, rng_seed=123, batch_size=20, device=None,
                     verbose=False, top=True):
    """Modality-parallel MDBN pretraining.

    specs: ordered dict  name -> dict(data=ndarray [N,V], layers_sizes=[...], pretraining_epochs=[...],
           pretrain_lr=[...], k=1, lambda_1=..., lambda_2=...)   (the train_ME/GE/SM arguments of
           src/AMLsm.py:207-339).
    The reference threads ONE numpy RandomState(123) through all DBN constructors in order
    (src/AMLsm.py:31-92): every rank replays that sequence on the host so initial weights are
    identical to a sequential run, then trains only the modalities placed on it.
    Returns (dbns on this rank, joint activations [N, sum H_top] on rank 0, top DBN on rank 0)."""
    from . import MDBN
    from .dbn import DBN
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    names = list(specs)
    where = place_modalities(names, world)
    rng = numpy.random.RandomState(rng_seed)
    dbns, tops = {}, {}
    for name in names:
        sp = specs[name]
        V = sp["data"].shape[1]
        if where[name] == rank:
            dbn, out, _ = MDBN.train_bottom_layer(sp["data"], None, batch_size=batch_size, k=sp.get("k", 1),
                                                  layers_sizes=sp["layers_sizes"],
                                                  pretraining_epochs=sp["pretraining_epochs"],
                                                  pretrain_lr=sp["pretrain_lr"], lambda_1=sp.get("lambda_1", 0.0),
                                                  lambda_2=sp.get("lambda_2", 0.1), rng=rng, device=device,
                                                  verbose=verbose)
            dbns[name], tops[name] = dbn, out
        else:
            # consume exactly the draws DBN.__init__ would (src/dbn.py:114,155-159) to stay in step
            rng.randint(2 ** 30)
            dims = [V] + list(sp["layers_sizes"])
            for a, b in zip(dims[:-1], dims[1:]):
                rng.uniform(size=(a, b))
    # one exchange: every rank writes the [N, H_top] activations of ITS modalities into their column block of a
    # zero [N, sum H_top] tensor; a tensor all-reduce(sum) then IS the concatenation (each block has one owner)
    widths = [int(specs[n]["layers_sizes"][-1]) for n in names]
    offs = numpy.concatenate([[0], numpy.cumsum(widths)]).astype(int)
    n_rows = int(specs[names[0]]["data"].shape[0])
    joint = None
    if world > 1:
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        buf = torch.zeros((n_rows, int(offs[-1])), dtype=torch.float32, device=dev)
        for i, n in enumerate(names):
            if n in tops:
                buf[:, offs[i]:offs[i + 1]] = torch.as_tensor(numpy.asarray(tops[n], dtype=numpy.float32)).to(dev)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        if rank == 0:
            joint = buf.cpu().numpy()
    else:
        joint = numpy.concatenate([tops[n] for n in names], axis=1)
    top_dbn = None
    if top and rank == 0:
        top_dbn = MDBN.train_top(batch_size, False, joint.astype(numpy.float32), None, rng, device=device, verbose=verbose)
    return dbns, joint, top_dbn


def aml_synthetic_specs(scale=1.0, n=170):
    """AML-shaped synthetic MDBN workload (BASELINE.json configs[3]): shapes from the run
    recorded in the reference's notebooks (ME 559->40, GE 19937->400->40, SM 1686->200->20, N = 170), hyper-parameters
    from src/AMLsm.py:38-62,207-339 and src/AMLsm2.py:308-339.  `scale` shrinks every patience budget (quick checks;
    a scaled run is NOT the named configuration)."""
    def zs(x):
        return ((x - x.mean(0)) / x.std(0)).astype(numpy.float32)
    me = zs(numpy.random.RandomState(2).randn(n, 559))
    ge = zs(numpy.random.RandomState(3).randn(n, 19937))
    rs = numpy.random.RandomState(4)
    sm_raw = (rs.rand(n, 1686) < 0.007) * rs.choice([1, 2, 3], size=(n, 1686), p=[0.985, 0.0146, 0.0004])
    sm_raw[0] += (sm_raw.sum(0) == 0)          # no zero-variance columns (the reference drops them, src/utils.py:97)
    sm = zs(sm_raw.astype(numpy.float64))
    sc = lambda xs: [max(2, int(round(x * scale))) for x in xs]
    return {
        "ME": dict(data=me, layers_sizes=[40], pretraining_epochs=sc([80000]), pretrain_lr=[0.005], k=10,
                   lambda_1=0.01, lambda_2=0.01),
        "GE": dict(data=ge, layers_sizes=[400, 40], pretraining_epochs=sc([8000, 800]), pretrain_lr=[0.005, 0.1], k=1,
                   lambda_1=0.01, lambda_2=0.1),
        "SM": dict(data=sm, layers_sizes=[200, 20], pretraining_epochs=sc([8000, 800]), pretrain_lr=[0.005, 0.1], k=1,
                   lambda_1=0.01, lambda_2=0.01),
    }
