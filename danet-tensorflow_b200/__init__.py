"""
danet-tensorflow_b200: the DANet separation hot path (STFT -> log-magnitude -> BiLSTM encoder ->
attractor estimation -> mask x complex mixture -> iSTFT, + PIT-MSE) as hand-written sm_100a CUDA
kernels behind a C-ABI (include/danet.h), under the Encoder / Estimator / Separator plugin
surface of khaotik/DaNet-Tensorflow.  Import name: `danet_tensorflow_b200` (see the loader
module of that name at the repository root; the directory name carries a hyphen).
"""
from . import _lib, build, kernels, shard
from .hparams import hparams, Hyperparameter
from . import modules, datasets, ozers
from .modules import Encoder, Estimator, Separator, ModelModule
from .model import Model
from .streaming import StreamingSeparator

__all__ = ['hparams', 'Hyperparameter', 'kernels', 'modules', 'Model', 'Encoder', 'Estimator',
           'Separator', 'ModelModule', 'build', '_lib', 'StreamingSeparator']
