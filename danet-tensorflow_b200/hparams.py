"""
Hyper-parameters and plugin registries -- TF-free mirror of the reference's
app/hparams.py:15-130 (same keys, `load`/`load_json`/`digest`, the five class-level
registries and their `register_*` / `get_*` accessors) plus default.json:2-40.
"""
import json
import re
import types

import numpy as np
import scipy.signal.windows

DEFAULTS = {   # default.json:2-40
    'FLOATX': 'float32', 'INTX': 'int32',
    'FFT_SIZE': 256, 'FFT_STRIDE': 64,
    'FFT_WND': 'np.sqrt(scipy.signal.hann(self.FFT_SIZE)).astype(self.FLOATX)',
    'SMPRATE': 8000,
    'BATCH_SIZE': 32, 'MAX_N_SIGNAL': 2,
    'LENGTH_ALIGN': 4, 'MAX_TRAIN_LEN': 128, 'EMBED_SIZE': 20,
    'RELU_LEAKAGE': 0.3, 'EPS': 1e-7, 'DROPOUT_KEEP_PROB': 1.0,
    'REG_SCALE': 1e-2, 'REG_TYPE': 'L2', 'LR': 3e-4, 'LR_DECAY': 0.8,
    'LR_DECAY_TYPE': None, 'NUM_EPOCH_PER_LR_DECAY': 10, 'GRAD_CLIP_THRES': 100.0,
    'TRAIN_ESTIMATOR_METHOD': 'truth-weighted', 'INFER_ESTIMATOR_METHOD': 'anchor',
    'NUM_ANCHOR': 6,
    'ENCODER_TYPE': 'toy', 'SEPARATOR_TYPE': 'dot-sigmoid-orig',
    'OPTIMIZER_TYPE': 'adam', 'DATASET_TYPE': 'toy',
    'SUMMARY_DIR': './logs', 'SUMMARY_TITLE': 'Test 1',
    'DEBUG': False,
}


def _window_namespace(self):
    # `scipy.signal.hann` (default.json:7) was removed from scipy; same symmetric window
    sig = types.SimpleNamespace(hann=scipy.signal.windows.hann, hanning=scipy.signal.windows.hann,
                                windows=scipy.signal.windows)
    return {'__builtins__': {}, 'np': np, 'scipy': types.SimpleNamespace(signal=sig), 'self': self}


class Hyperparameter:
    pattern = r'[A-Z_]+'
    encoder_registry = {}
    estimator_registry = {}
    separator_registry = {}
    ozer_registry = {}
    dataset_registry = {}

    def __init__(self):
        self.__dict__.update(DEFAULTS)

    def digest(self):
        """app/hparams.py:29-42: derive COMPLEXX, FEATURE_SIZE, evaluate FFT_WND"""
        self.COMPLEXX = dict(float32='complex64', float64='complex128')[self.FLOATX]
        self.FEATURE_SIZE = 1 + self.FFT_SIZE // 2
        assert isinstance(self.DROPOUT_KEEP_PROB, float)
        assert 0. < self.DROPOUT_KEEP_PROB <= 1.
        if isinstance(self.FFT_WND, str):
            self.FFT_WND = eval(self.FFT_WND, _window_namespace(self))   # noqa: S307 (as the reference)

    def load(self, di):
        assert isinstance(di, dict)
        pat = re.compile(self.pattern)
        for k, v in di.items():
            if pat.fullmatch(k) is None:
                raise NameError(k)
            assert isinstance(v, (str, int, float, bool, type(None)))
        self.__dict__.update(di)

    def load_json(self, file_):
        if isinstance(file_, (str, bytes)):
            with open(file_, 'r') as f:
                di = json.load(f)
        else:
            di = json.load(file_)
        self.load(di)

    @classmethod
    def _register(cls_, registry, name):
        def wrapper(obj):
            registry[name] = obj
            return obj
        return wrapper

    @classmethod
    def register_encoder(cls_, name):
        return cls_._register(cls_.encoder_registry, name)

    def get_encoder(self):
        return type(self).encoder_registry[self.ENCODER_TYPE]

    @classmethod
    def register_estimator(cls_, name):
        return cls_._register(cls_.estimator_registry, name)

    def get_estimator(self, name):
        return type(self).estimator_registry[name]

    @classmethod
    def register_separator(cls_, name):
        return cls_._register(cls_.separator_registry, name)

    def get_separator(self, name):
        return type(self).separator_registry[name]

    @classmethod
    def register_optimizer(cls_, name):
        return cls_._register(cls_.ozer_registry, name)

    def get_optimizer(self):
        return type(self).ozer_registry[self.OPTIMIZER_TYPE]

    @classmethod
    def register_dataset(cls_, name):
        return cls_._register(cls_.dataset_registry, name)

    def get_dataset(self):
        return type(self).dataset_registry[self.DATASET_TYPE]


hparams = Hyperparameter()
