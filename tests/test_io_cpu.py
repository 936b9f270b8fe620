"""SURVEY 8f rows either side of the hot path: TCGA table pre-processing, class extraction, .npz checkpoint
schema.  Expected values were produced by the reference's own functions (tests/golden/make_golden.py io)."""
import os

import numpy as np
import pytest

from mdbn_b200 import io as mio

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "io_cases.npz"))
REF_CKPT = os.path.join(HERE, "golden", "ref_checkpoint.npz")


@pytest.fixture()
def table_dir(tmp_path):
    (tmp_path / "toy.tsv").write_bytes(G["table_text"].tobytes())
    return str(tmp_path)


@pytest.mark.parametrize("name,kw", [
    ("a", dict(holdout=0.25, repeats=1, clip=None, shuffle=True, seed=11, transform=False)),
    ("b", dict(holdout=0.2, repeats=3, clip=(-1.0, 1.0), shuffle=False, seed=12, transform=False)),
    ("c", dict(holdout=0.0, repeats=2, clip=None, shuffle=True, seed=13, transform=True))])
def test_preprocessing_matches_reference(table_dir, name, kw):
    np.random.seed(kw["seed"])
    tr, va = mio.load_n_preprocess_data("toy.tsv", dtype="float64", holdout=kw["holdout"], clip=kw["clip"],
                                        transform_fn=np.power if kw["transform"] else None, exponent=0.5,
                                        repeats=kw["repeats"], shuffle=kw["shuffle"], datadir=table_dir)
    np.testing.assert_allclose(tr, G["prep_%s_train" % name], rtol=1e-12, atol=1e-12)
    if G["prep_%s_val" % name].size:
        np.testing.assert_allclose(va, G["prep_%s_val" % name], rtol=1e-12, atol=1e-12)
    else:
        assert va is None


def test_preprocessing_gz_and_shape_triplet(table_dir):
    import gzip
    with gzip.open(os.path.join(table_dir, "toy.tsv.gz"), "wb") as f:
        f.write(G["table_text"].tobytes())
    n_data, n_cols, data = mio.import_TCGA_data("toy.tsv.gz", table_dir, "float32")
    assert (n_data, n_cols) == (12, 12) and data.shape == (40, 12) and data.dtype == np.float32


def test_class_extraction_matches_reference():
    labels, dist = mio.find_unique_classes(G["cls_bits"])
    assert np.array_equal(labels, G["cls_labels"])
    assert np.array_equal(dist, G["cls_dist"])
    for n in (2, 3, 5):
        assert np.array_equal(mio.remap_class(labels.astype(int), dist, n), G["cls_remap_%d" % n])


class _Param:
    def __init__(self, name, value):
        self.name, self._v = name, value

    def get_value(self):
        return self._v


class _Net:
    def __init__(self, sizes, W, b):
        self._sizes = list(sizes)
        self.params = [p for w, bb in zip(W, b) for p in (_Param("W", w), _Param("b", bb))]

    def number_of_nodes(self):
        return self._sizes


def test_checkpoint_schema_round_trip(tmp_path):
    nets, extra = mio.read_network_file(REF_CKPT)                       # written by the reference's save_network
    assert sorted(nets) == ["ge", "me", "sm", "top"]
    assert nets["ge"][0] == {"number_of_nodes": [13, 6, 3], "epochs": [8000, 800], "learning_rate": [0.005, 0.1],
                             "batch_size": 20, "k": 1}
    assert [w.shape for w in nets["ge"][1]] == [(13, 6), (6, 3)] and [b.shape for b in nets["ge"][2]] == [(6,), (3,)]
    stubs = {k: _Net(cfg["number_of_nodes"], W, b) for k, (cfg, W, b) in nets.items()}
    mio.save_network(extra["classes"], stubs["ge"], stubs["me"], stubs["sm"], None, stubs["top"], float(extra["holdout"]),
                     "mine.npz", str(tmp_path), int(extra["repeats"]))
    ref = np.load(REF_CKPT, allow_pickle=True)
    mine = np.load(os.path.join(str(tmp_path), "mine.npz"), allow_pickle=True)
    assert sorted(mine.files) == sorted(ref.files)
    for key in ref.files:
        if key.endswith("_config"):
            assert mine[key].tolist() == ref[key].tolist()
        elif key.endswith("_params"):
            # read the way the reference's load_network reads it: params[i]['W'] / params[i + 1]['b']
            assert len(mine[key]) == len(ref[key])
            for a, b in zip(mine[key], ref[key]):
                assert list(a) == list(b) and np.array_equal(a[list(a)[0]], b[list(b)[0]])
        else:
            assert np.array_equal(mine[key], ref[key])


def test_load_network_rebuilds_with_reference_arguments():
    calls = []

    def factory(**kw):
        calls.append(kw)
        return kw
    me, ge, sm, dm, top = mio.load_network(REF_CKPT, dbn_factory=factory)
    assert dm is None
    assert (ge["n_ins"], ge["hidden_layers_sizes"], ge["n_outs"]) == (13, [6], 3) and "gauss" not in ge
    assert top["gauss"] is False and len(top["W_list"]) == 2 and len(top["b_list"]) == 2
    assert me["W_list"][0].shape == (9, 4) and sm["b_list"][0].shape == (5,)


def test_mnist_loader_and_tilings_match_reference(tmp_path):
    import gzip
    with gzip.open(str(tmp_path / "img.gz"), "wb") as f:
        f.write(G["mnist_img_bytes"].tobytes())
    with gzip.open(str(tmp_path / "lab.gz"), "wb") as f:
        f.write(G["mnist_lab_bytes"].tobytes())
    mn = mio.MNIST("img.gz", "lab.gz", str(tmp_path), dtype="float64")
    assert [mn.n_images, mn.sizeY, mn.sizeX, mn.n_levels] == G["mnist_meta"].tolist()
    assert np.array_equal(mn.images, G["mnist_images"]) and np.array_equal(mn.labels, G["mnist_labels"])
    np.testing.assert_allclose(mn.normalize(mn.images), G["mnist_norm"], rtol=1e-13)
    assert np.array_equal(mn.display_weigths(G["mnist_W"], 7), G["mnist_tiles_w7"])
    assert np.array_equal(mn.display_weigths(G["mnist_W"][:, :4], 4), G["mnist_tiles_w4"])
    assert np.array_equal(mn.display_samples(list(G["mnist_samples"])), G["mnist_tiles_s"])
    with pytest.raises(IOError):
        mio.MNIST("missing.gz", "lab.gz", str(tmp_path))
