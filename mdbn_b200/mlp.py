"""HiddenLayer (src/mlp.py:36-110): output = sigmoid(input W + b).  In a DBN its W, b are the
same Shared objects as the RBM's W, hbias (src/dbn.py:193-202)."""
import ctypes

import numpy
import torch

from . import _lib
from .utils import Shared, as_device_matrix, default_device


class HiddenLayer(object):
    def __init__(self, rng, input, n_in, n_out, W=None, b=None, activation="sigmoid", device=None):
        self.device = torch.device(device) if device is not None else default_device()
        self.ctx = _lib.context(self.device.index if self.device.index is not None else torch.cuda.current_device())
        if activation not in ("sigmoid", None):
            raise NotImplementedError("the reference only ever builds sigmoid layers (src/dbn.py:174)")
        self.input = input
        self.n_in, self.n_out = n_in, n_out
        if W is None:
            W_values = numpy.asarray(rng.uniform(low=-numpy.sqrt(6. / (n_in + n_out)),
                                                 high=numpy.sqrt(6. / (n_in + n_out)),
                                                 size=(n_in, n_out)), dtype=numpy.float32)
            if activation == "sigmoid":
                W_values *= 4                                            # src/mlp.py:90-91
            W = W_values
        if not isinstance(W, Shared):
            W = Shared(W, name='W', device=self.device, ld_pad=8)
        if b is None:
            b = numpy.zeros((n_out,), dtype=numpy.float32)
        if not isinstance(b, Shared):
            b = Shared(b, name='b', device=self.device)
        self.W, self.b = W, b
        self.activation = activation
        self.params = [self.W, self.b]

    def output(self, x):
        """sigmoid(x W + b) as a device tensor (src/mlp.py:103-107)."""
        x = as_device_matrix(x, self.device)
        out = torch.empty((x.shape[0], self.n_out), dtype=torch.float32, device=self.device)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.activation is not None:
            _lib.check(self.ctx.lib.mdbn_forward(self.ctx.handle, self.W.storage.data_ptr(), self.W.ld,
                                                 self.b.data.data_ptr(), x.data_ptr(), x.stride(0), x.shape[0],
                                                 self.n_in, self.n_out, out.data_ptr(), st))
        else:
            _lib.check(self.ctx.lib.mdbn_propup(self.ctx.handle, self.W.storage.data_ptr(), self.W.ld,
                                                self.b.data.data_ptr(), x.data_ptr(), x.stride(0), x.shape[0],
                                                self.n_in, self.n_out, ctypes.c_void_p(out.data_ptr()), None, None, None, st))
        return out
