"""The data formats either side of the hot path: the TCGA table loader and
pre-processing that feeds it, the class extraction that consumes its output, and the `.npz`
checkpoint schema of the experiment scripts.  Plain NumPy/SciPy on the host, restated from the
reference so that its notebooks can read runs produced here and vice versa.

  import_TCGA_data / load_n_preprocess_data   src/utils.py:34-52, :77-119
  find_unique_classes / remap_class           src/utils.py:124-177
  save_network / load_network                 src/AMLsm.py:112-205 (+ DBN(W_list=, b_list=), src/dbn.py:154-166)
  MNIST (idx loader, normalize, tilings)      src/MNIST.py:33-124 (the sampling demo's data and pictures)
"""
import gzip
import os

import numpy

from .utils import get_minibatches_idx


# --------------------------------------------------------------------------- before the path
def import_TCGA_data(file, datadir='data', dtype='float32'):
    """Tab-separated table, first row = header, first column = row names, one COLUMN per person.
    Returns (n_persons, n_persons, data[features, persons]) exactly as src/utils.py:34-52 does (its
    first two values are the same number computed two ways)."""
    path = os.path.join(datadir, file)
    opener = gzip.open if file.endswith('.gz') else open
    with opener(path, 'rt') as f:
        ncols = len(f.readline().split('\t'))
    data = numpy.loadtxt(path, dtype=dtype, delimiter='\t', skiprows=1, usecols=range(1, ncols))
    if data.ndim == 1:
        data = data.reshape(1, -1)
    return data.shape[1], ncols - 1, data


def load_n_preprocess_data(datafile, dtype='float32', holdout=0.1, clip=None, transform_fn=None, exponent=1.0,
                           repeats=10, shuffle=True, datadir='data'):
    """src/utils.py:77-119: per-feature z-score over the population (ddof = 0), features with a NaN dropped,
    persons to rows, optional clip, every person repeated `repeats` times, hold-out split.

    Quirk kept on purpose: the split indexes are drawn over the n_cols ORIGINAL persons and applied to the
    repeated matrix, so with repeats > 1 only the first n_cols rows (copies of the first persons) are used.
    Returns (train, validation-or-None) as host arrays; every entry point of the path takes them as they are
    (the reference wraps them in theano.shared)."""
    n_data, n_cols, data = import_TCGA_data(datafile, datadir, dtype)
    if transform_fn is not None:
        data = transform_fn(data, exponent)
    with numpy.errstate(invalid='ignore', divide='ignore'):
        mean = data.mean(axis=1, keepdims=True)
        std = data.std(axis=1, keepdims=True)
        z = (data - mean) / std
    z = z[~numpy.isnan(z).any(axis=1)].T
    if clip is not None:
        z = numpy.clip(z, clip[0], clip[1])
    if repeats > 1:
        z = numpy.repeat(z, repeats=repeats, axis=0)
    n_val = int(n_cols * holdout)
    _, parts = get_minibatches_idx(n_cols, n_cols - n_val, shuffle=shuffle)
    train = numpy.ascontiguousarray(z[parts[0]])
    val = numpy.ascontiguousarray(z[parts[1]]) if n_val > 0 else None
    return train, val


# --------------------------------------------------------------------------- after the path
def find_unique_classes(dbn_output):
    """src/utils.py:161-177.  Classes are the distinct rows (node patterns) of the thresholded output, numbered
    in the byte-wise order numpy.unique gives a void view of the rows; returns (class of every sample as a
    float vector, Hamming distance matrix between the class patterns)."""
    from scipy.spatial import distance
    out = numpy.ascontiguousarray(dbn_output)
    row_bytes = out.view(numpy.dtype((numpy.void, out.dtype.itemsize * out.shape[1])))
    _, first = numpy.unique(row_bytes, return_index=True)
    patterns = out[first]
    dist = distance.cdist(patterns, patterns, metric='hamming')
    labels = numpy.zeros(out.shape[0])
    for c, pat in enumerate(patterns):
        labels = labels + (numpy.sum(out == pat, axis=1) == out.shape[1]) * c
    return labels, dist


def remap_class(classified_samples, distance_matrix, n_classes):
    """src/utils.py:124-158: keep the n_classes most frequent classes (relabelled 0.. by rank), fold every other
    class into a kept one.  Two quirks of the reference are reproduced: the scan over the kept classes by
    increasing Hamming distance has no early exit, so the LAST qualifying one (the farthest) wins, and the
    self-exclusion test compares a frequency rank with a class id."""
    a = numpy.asarray(classified_samples)
    n_initial = int(numpy.max(a)) + 1
    freq = [int(numpy.sum(a == c)) for c in range(n_initial)]
    by_rank = list(reversed(numpy.argsort(freq).tolist()))          # class id at every frequency rank
    rank_of = {c: r for r, c in enumerate(by_rank)}
    new_label = {by_rank[r]: r for r in range(n_classes)}
    D = numpy.asarray(distance_matrix)
    for c in [by_rank[r] for r in range(n_classes, D.shape[0])]:
        for i in numpy.argsort(D[c]):
            r = rank_of[int(i)]
            if r < n_classes and r != c:
                new_label[c] = r
    return numpy.array([new_label[int(c)] for c in a])


# --------------------------------------------------------------------------- checkpoints
_REFERENCE_CONFIGS = {       # the hyper-parameters src/AMLsm.py:121-157 writes next to the weights
    'me': dict(epochs=[8000], learning_rate=[0.005], batch_size=20, k=10),
    'ge': dict(epochs=[8000, 800], learning_rate=[0.005, 0.1], batch_size=20, k=1),
    'sm': dict(epochs=[8000], learning_rate=[0.005], batch_size=20, k=10),
    'dm': dict(epochs=[8000, 800], learning_rate=[0.005, 0.1], batch_size=20, k=1),
    'top': dict(epochs=[800, 800], learning_rate=[0.1, 0.1], batch_size=20, k=1),
}


def network_arrays(networks, classes, holdout, repeats, configs=None):
    """The keyword arguments of the reference's numpy.savez call: `<name>_config` dicts, `<name>_params` lists of
    one-entry {parameter name: array} dicts in DBN.params order (W, b per layer), classes, holdout, repeats."""
    out = dict(holdout=holdout, repeats=repeats, classes=classes)
    for name, net in networks.items():
        if net is None:
            continue
        cfg = dict(number_of_nodes=net.number_of_nodes())
        cfg.update((configs or {}).get(name, _REFERENCE_CONFIGS.get(name, {})))
        out[name + '_config'] = cfg
        out[name + '_params'] = [{p.name: numpy.asarray(p.get_value())} for p in net.params]
    return out


def save_network(classes, ge_DBN, me_DBN, sm_DBN, dm_DBN, top_DBN, holdout, output_file, output_folder, repeats,
                 configs=None):
    """src/AMLsm.py:112-163, same argument order.  (The reference never writes the DM network; here it is written
    when one is passed.)"""
    os.makedirs(output_folder, exist_ok=True)
    nets = dict(me=me_DBN, ge=ge_DBN, sm=sm_DBN, dm=dm_DBN, top=top_DBN)
    arrays = network_arrays(nets, classes, holdout, repeats, configs)
    numpy.savez(os.path.join(output_folder, output_file), **{k: _as_savez_value(v) for k, v in arrays.items()})


def _as_savez_value(v):
    if isinstance(v, (dict, list)):
        a = numpy.empty((), dtype=object) if isinstance(v, dict) else numpy.empty(len(v), dtype=object)
        if isinstance(v, dict):
            a[()] = v
        else:
            for i, x in enumerate(v):
                a[i] = x
        return a
    return v


def read_network_file(input_file, input_folder='.'):
    """{name: (config dict, [W per layer], [b per layer])} of every network in a checkpoint, plus the scalars."""
    npz = numpy.load(os.path.join(input_folder, input_file), allow_pickle=True)
    nets, extra = {}, {}
    for key in npz.files:
        if key.endswith('_config'):
            name = key[:-len('_config')]
            params = list(npz[name + '_params'])
            W = [d['W'] for d in params if 'W' in d]
            b = [d['b'] for d in params if 'b' in d]
            nets[name] = (npz[key].tolist(), W, b)
        elif not key.endswith('_params'):
            extra[key] = npz[key]
    return nets, extra


def load_network(input_file, input_folder='.', dbn_factory=None):
    """src/AMLsm.py:165-205: (me_DBN, ge_DBN, sm_DBN, dm_DBN-or-None, top_DBN) rebuilt through
    DBN(n_ins, hidden_layers_sizes, n_outs, W_list=, b_list=) — the top network Bernoulli (gauss=False)."""
    if dbn_factory is None:
        from .dbn import DBN as dbn_factory
    nets, _ = read_network_file(input_file, input_folder)
    built = {}
    for name, (cfg, W, b) in nets.items():
        sizes = cfg['number_of_nodes']
        kw = dict(n_ins=sizes[0], hidden_layers_sizes=list(sizes[1:-1]), n_outs=sizes[-1], W_list=W, b_list=b)
        if name == 'top':
            kw['gauss'] = False
        built[name] = dbn_factory(**kw)
    return built.get('me'), built.get('ge'), built.get('sm'), built.get('dm'), built.get('top')


# --------------------------------------------------------------------------- the demo's data set and pictures
class MNIST(object):
    """src/MNIST.py:33-124: idx-format images and labels (gzip), the guideTR 13.2 normalisation and the two image
    tilings of the RBM demo.  The files must already be in `datadir` (the reference downloads them; no network
    here).  Attributes as in the reference: images [n, sizeY*sizeX], labels [n, 1], n_levels, sizeX, sizeY."""

    def __init__(self, datafile='train-images-idx3-ubyte.gz', targetfile='train-labels-idx1-ubyte.gz', datadir='data',
                 dtype='float32'):
        self.dtype = dtype
        self.n_images = self.load(datafile, targetfile, datadir)

    def load(self, datafile, targetfile, datadir):
        import struct
        for f in (datafile, targetfile):
            if not os.path.isfile(os.path.join(datadir, f)):
                raise IOError("%s is not in %s (fetch it from http://yann.lecun.com/exdb/mnist)" % (f, datadir))
        with gzip.open(os.path.join(datadir, datafile), 'rb') as f:
            raw = f.read()
        _, n_images, self.sizeY, self.sizeX = struct.unpack(">IIII", raw[:16])
        size = self.sizeY * self.sizeX
        self.images = numpy.frombuffer(raw, dtype=numpy.uint8, count=n_images * size, offset=16) \
            .reshape(n_images, size).astype(self.dtype)
        with gzip.open(os.path.join(datadir, targetfile), 'rb') as f:
            raw = f.read()
        _, n_labels = struct.unpack(">II", raw[:8])
        self.labels = numpy.frombuffer(raw, dtype=numpy.uint8, count=n_labels, offset=8) \
            .reshape(n_labels, 1).astype(self.dtype)
        self.n_levels = int(self.labels.max() - self.labels.min() + 1)
        return n_images

    def normalize(self, X):
        """Zero mean, about unit deviation: (X - 128) / 128 over the GLOBAL standard deviation (:90-96)."""
        X = (X - 128.0) / 128.0
        return X / numpy.std(X)

    def _tile(self, rows_of_images):
        gap_r, gap_c = self.sizeY + 1, self.sizeX + 1
        Y = numpy.zeros((len(rows_of_images) * gap_r, max(len(r) for r in rows_of_images) * gap_c), dtype=self.dtype)
        for r, row in enumerate(rows_of_images):
            for c, img in enumerate(row):
                Y[r * gap_r:(r + 1) * gap_r - 1, c * gap_c:(c + 1) * gap_c - 1] = \
                    numpy.asarray(img).reshape(self.sizeX, self.sizeY)          # (sic: X before Y, :108)
        return Y

    def display_weigths(self, X, n_hidden):
        """One tile per hidden unit on a round(sqrt(n_hidden))-wide square grid (:98-111; the reference's
        spelling; units that do not fit the square are not drawn, a short last row stays black)."""
        W = numpy.asarray(X).T
        n = int(numpy.round(numpy.sqrt(n_hidden)))
        rows = [[W[r * n + c] for c in range(n) if r * n + c < n_hidden] for r in range(n)]
        Y = self._tile([r for r in rows if r] or [[]])
        full = numpy.zeros((n * (self.sizeY + 1), n * (self.sizeX + 1)), dtype=self.dtype)
        full[:Y.shape[0], :Y.shape[1]] = Y
        return full

    def display_samples(self, samples):
        """One row of tiles per sampling round, one column per chain (:113-122)."""
        return self._tile([list(s) for s in samples])

