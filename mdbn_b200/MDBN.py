"""Multimodal DBN composition (src/MDBN.py:31-76): one DBN per modality with a Gaussian
bottom layer, then a joint Bernoulli DBN [sum of top widths -> 24 -> 3]."""
from __future__ import print_function, division

from .dbn import DBN
from .utils import as_device_matrix


def _n_cols(x):
    if hasattr(x, "get_value"):
        return x.get_value(borrow=True).shape[1]
    return x.shape[1]


def train_top(batch_size, graph_output, joint_train_set, joint_val_set, rng, device=None, verbose=True):
    top_DBN = DBN(numpy_rng=rng, n_ins=_n_cols(joint_train_set), gauss=False, hidden_layers_sizes=[24], n_outs=3,
                  device=device, verbose=verbose)
    top_DBN.training(joint_train_set, batch_size, k=1, pretraining_epochs=[800, 800], pretrain_lr=[0.1, 0.1],
                     validation_set_x=joint_val_set, graph_output=graph_output)
    return top_DBN


def train_bottom_layer(train_set, validation_set, batch_size=20, k=1, layers_sizes=[40], pretraining_epochs=[800],
                       pretrain_lr=[0.005], lambda_1=0.0, lambda_2=0.1, rng=None, graph_output=False,
                       device=None, verbose=True):
    if verbose:
        print('Visible nodes: %i' % _n_cols(train_set))
        print('Output nodes: %i' % layers_sizes[-1])
    dbn = DBN(numpy_rng=rng, n_ins=_n_cols(train_set), hidden_layers_sizes=layers_sizes[:-1],
              n_outs=layers_sizes[-1], device=device, verbose=verbose)
    dbn.training(train_set, batch_size, k=k, pretraining_epochs=pretraining_epochs, pretrain_lr=pretrain_lr,
                 lambda_1=lambda_1, lambda_2=lambda_2, validation_set_x=validation_set, graph_output=graph_output)
    output_train_set = dbn.get_output(train_set)
    output_val_set = dbn.get_output(validation_set) if validation_set is not None else None
    return dbn, output_train_set, output_val_set
