"""mdbn_b200 — B200-native RBM / GRBM contrastive-divergence path behind the class
surface of glgerard/MDBN's src/rbm.py, src/dbn.py, src/mlp.py and src/MDBN.py."""
from . import _lib
from .utils import Shared, shared, get_minibatches_idx
from .rng import RandomStreams, BufferStreams
from .rbm import RBM, GRBM
from .mlp import HiddenLayer
from .dbn import DBN
from . import MDBN
from . import io

__all__ = ["RBM", "GRBM", "DBN", "HiddenLayer", "MDBN", "io", "Shared", "shared", "get_minibatches_idx",
           "RandomStreams", "BufferStreams"]
