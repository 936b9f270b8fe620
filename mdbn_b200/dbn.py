"""Deep Belief Network with the reference's surface (src/dbn.py:51-536): a stack of
HiddenLayer + RBM/GRBM pairs sharing W and hbias, trained greedily with CD-k."""
from __future__ import print_function, division

import sys
import timeit

import numpy
import torch

from .mlp import HiddenLayer
from .rbm import RBM, GRBM
from .rng import RandomStreams
from .utils import Shared, as_device_matrix, default_device, get_minibatches_idx, IndexFeeder


class _LowerStack:
    """Deterministic up-pass through layers [0, n) (src/dbn.py:146) + a version stamp."""

    def __init__(self, layers):
        self.layers = layers

    def version(self):
        return tuple(p.version for l in self.layers for p in l.params)

    def __call__(self, x):
        for l in self.layers:
            x = l.output(x)
        return x


class DBN(object):
    def __init__(self, numpy_rng=None, theano_rng=None, n_ins=784, gauss=True,
                 hidden_layers_sizes=[400], n_outs=40, W_list=None, b_list=None, device=None, verbose=True):
        self.device = torch.device(device) if device is not None else default_device()
        self.verbose = verbose
        self.n_ins = n_ins
        self.sigmoid_layers = []
        self.rbm_layers = []
        self.params = []
        self.stacked_layers_sizes = list(hidden_layers_sizes) + [n_outs]
        self.n_layers = len(self.stacked_layers_sizes)
        self.train_path, self.tf32 = "auto", False
        assert self.n_layers > 0
        if numpy_rng is None:
            numpy_rng = numpy.random.RandomState(123)                        # src/dbn.py:111
        if theano_rng is None:
            theano_rng = RandomStreams(numpy_rng.randint(2 ** 30))           # :114
        self.theano_rng = theano_rng

        for i in range(self.n_layers):
            n_in = n_ins if i == 0 else self.stacked_layers_sizes[i - 1]
            n_out = self.stacked_layers_sizes[i]
            if self.verbose:
                print('Adding a layer with %i input and %i outputs' % (n_in, n_out))
            if W_list is None:
                W = numpy.asarray(numpy_rng.uniform(low=-4. * numpy.sqrt(6. / (n_in + n_out)),
                                                    high=4. * numpy.sqrt(6. / (n_in + n_out)),
                                                    size=(n_in, n_out)), dtype=numpy.float32)   # :155-159
            else:
                W = W_list[i]
            b = numpy.zeros((n_out,), dtype=numpy.float32) if b_list is None else b_list[i]
            sigmoid_layer = HiddenLayer(rng=numpy_rng, input=None, n_in=n_in, n_out=n_out,
                                        W=Shared(W, name='W', device=self.device, ld_pad=8),
                                        b=Shared(b, name='b', device=self.device), device=self.device)
            self.sigmoid_layers.append(sigmoid_layer)
            self.params.extend(sigmoid_layer.params)
            cls = GRBM if (i == 0 and gauss) else RBM                        # :187
            lower = _LowerStack(self.sigmoid_layers[:i]) if i > 0 else None
            rbm_layer = cls(numpy_rng=numpy_rng, theano_rng=theano_rng, input=lower, n_visible=n_in,
                            n_hidden=n_out, W=sigmoid_layer.W, hbias=sigmoid_layer.b, device=self.device)
            self.rbm_layers.append(rbm_layer)

    def number_of_nodes(self):
        return [self.n_ins] + self.stacked_layers_sizes                      # :206-212

    def get_output(self, input, layer=-1):
        """Output of MLP layer `layer` as a numpy array, None for None (src/dbn.py:214-236)."""
        if input is None:
            return None
        n = self.n_layers if layer == -1 else (layer % self.n_layers) + 1
        return _LowerStack(self.sigmoid_layers[:n])(as_device_matrix(input, self.device)).cpu().numpy()

    def training_functions(self, train_set_x, batch_size, k, lambda_1=0.0, lambda_2=0.1, monitor=False):
        """(train_fns, free_energy_gap_fns) — src/dbn.py:238-332.  train_fns[i](indexes=, momentum=,
        lr=) -> cost; free_energy_gap_fns[i](train, test) -> (F_train, F_test)."""
        assert batch_size > 1                                                # :276
        train_fns, free_energy_gap_fns = [], []
        for i, rbm in enumerate(self.rbm_layers):
            if isinstance(rbm, GRBM):
                cost, updates = rbm.get_cost_updates(0.0, lambda_1=lambda_1, lambda_2=lambda_2,
                                                     batch_size=batch_size, persistent=None, k=k)   # :285-289
            else:
                cost, updates = rbm.get_cost_updates(0.0, weightcost=0.0002, batch_size=batch_size,
                                                     persistent=None, k=k)                          # :291-294
            fn = rbm.make_train_fn(train_set_x, cost, updates, layer_id=i, input_fn=rbm.input,
                                   path=self.train_path, tf32=self.tf32)
            train_fns.append(fn)

            def feg(train, test, rbm=rbm):
                a, b = rbm.free_energies(train, test)
                return a.cpu().numpy(), b.cpu().numpy()
            free_energy_gap_fns.append(feg)
        return train_fns, free_energy_gap_fns

    pretraining_functions = training_functions   # the upstream tutorial's / BASELINE.json's name

    def training(self, train_set_x, batch_size, k, pretraining_epochs, pretrain_lr, lambda_1=0.0, lambda_2=0.1,
                 validation_set_x=None, monitor=False, graph_output=False):
        """Greedy layer-wise pretraining with the reference's patience logic (src/dbn.py:334-517)."""
        log = print if self.verbose else (lambda *a, **k: None)
        train = as_device_matrix(train_set_x, self.device)
        val = as_device_matrix(validation_set_x, self.device) if validation_set_x is not None else None
        log('... getting the pretraining functions')
        log('Training set sample size %i' % train.shape[0])
        if val is not None:
            log('Validation set sample size %i' % val.shape[0])
        training_fns, free_energy_gap_fns = self.training_functions(train_set_x=train, batch_size=batch_size, k=k,
                                                                    lambda_1=lambda_1, lambda_2=lambda_2,
                                                                    monitor=monitor)
        log('... pre-training the model')
        start_time = timeit.default_timer()
        n_data = train.shape[0]
        patience_increase = 2
        improvement_threshold = 0.995
        idx_minibatches, minibatches = get_minibatches_idx(n_data, batch_size, shuffle=True)      # :420
        n_train_batches = idx_minibatches[-1] + 1
        self.history = []
        feeder = IndexFeeder(self.device, n_data)
        for i in range(self.n_layers):
            momentum = 0.0 if isinstance(self.rbm_layers[i], GRBM) else 0.6                       # :430-433
            best_cost = numpy.inf
            epoch = 0
            done_looping = False
            patience = pretraining_epochs[i]                                                      # :440
            validation_frequency = min(20 * n_train_batches, patience // 2)                       # :441
            log('Validation frequency: %d' % validation_frequency)
            fn = training_fns[i]
            fn.sync = False          # the cost is only looked at on validation iterations
            hist = []
            while (epoch < pretraining_epochs[i]) and (not done_looping):
                epoch = epoch + 1
                idx_minibatches, minibatches = get_minibatches_idx(n_data, batch_size, shuffle=True)
                flat = feeder.upload(minibatches)      # one asynchronous H2D per epoch (pinned ring: no stream sync)
                if not isinstance(self.rbm_layers[i], GRBM) and epoch == 6:
                    momentum = 0.9                                                                # :452-453
                lo, mb = 0, 0
                while mb < len(minibatches):
                    # Chain the minibatches up to the next iteration at which the reference's loop looks at
                    # anything: a validation iteration, or the one where the patience runs out (patience only
                    # changes on validation iterations, so it is constant inside a chunk).  Same steps, same
                    # order; with the in-kernel generator one launch per chunk (TrainFn.run_steps).
                    iter0 = (epoch - 1) * n_train_batches + mb
                    B0 = len(minibatches[mb])
                    next_val = (iter0 // validation_frequency + 1) * validation_frequency - 1
                    stop = min(next_val, max(patience, iter0), (epoch - 1) * n_train_batches + len(minibatches) - 1)
                    n = 1
                    while n < stop - iter0 + 1 and len(minibatches[mb + n]) == B0:
                        n += 1
                    costs_dev = fn.run_steps(flat[lo:lo + n * B0].view(n, B0), momentum=momentum, lr=pretrain_lr[i])
                    lo += n * B0
                    mb += n
                    iter = iter0 + n - 1
                    if (iter + 1) % validation_frequency == 0:
                        current_cost = float(costs_dev[n - 1].item())
                        log('Pre-training cost (layer %i, epoch %d): ' % (i, epoch), end=' ')
                        log(current_cost)
                        feg = None
                        if current_cost < best_cost:
                            if current_cost < best_cost * improvement_threshold:
                                patience = max(patience, iter * patience_increase)
                            best_cost = current_cost
                            if val is not None:
                                if i == 0:
                                    input_t_set, input_v_set = train, val                         # :490-492
                                else:
                                    low = _LowerStack(self.sigmoid_layers[:i])
                                    input_t_set = low(train[:val.shape[0]])                       # :494-496
                                    input_v_set = low(val)
                                f_train, f_test = free_energy_gap_fns[i](input_t_set, input_v_set)
                                feg = float(f_test.mean() - f_train.mean())
                                log('Free energy gap (layer %i, epoch %i): ' % (i, epoch), end=' ')
                                log(feg)
                        hist.append((iter, current_cost, feg))
                    if patience <= iter:                                                          # :506-508
                        done_looping = True
                        break
            self.history.append(dict(epochs=epoch, calls=fn.n_calls, validations=hist))
        end_time = timeit.default_timer()
        torch.cuda.synchronize(self.device)
        if self.verbose:
            print('The pretraining code ran for %.2fm' % ((end_time - start_time) / 60.), file=sys.stderr)
        return self.history
