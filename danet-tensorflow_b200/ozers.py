"""Optimisers behind `hparams.register_optimizer` (reference: app/ozers.py:9-18).  Each entry returns the
keyword arguments of the fused clip + update kernel; `lr_decay` arguments are ignored as in the reference."""
from .hparams import hparams


@hparams.register_optimizer('adam')
def adam(learn_rate, lr_decay=None):
    return dict(kind='adam', lr=learn_rate, beta1=.9, beta2=.999, eps=1e-8)      # tf.train.AdamOptimizer defaults


@hparams.register_optimizer('sgd')
def sgd(learn_rate, lr_decay=None):
    # plain gradient descent = Adam kernel is not applicable; Model.apply_gradients handles kind == 'sgd'
    return dict(kind='sgd', lr=learn_rate)
