"""
Datasets behind `hparams.register_dataset` (reference: app/datasets/dataset.py:8-63).  `epoch(subset,
batch_size)` yields tuples whose first entry is `[batch_size = B*C, T, F]` spectra (main.py:414-421).
  toy        -- the reference's WhiteNoiseData: uniform noise, 10 batches of 128 frames (dataset.py:43-63)
  synthetic  -- shaped-noise waveforms through the device STFT: stands in for the licensed
                TIMIT / WSJ0 corpora, which are not redistributable
"""
import numpy as np
import torch

from . import kernels as K
from .hparams import hparams


class Dataset(object):
    """app/datasets/dataset.py:8-40"""
    def __init__(self):
        self.is_loaded = False

    def epoch(self, subset, batch_size, shuffle=False):
        raise NotImplementedError()

    def install_and_load(self):
        raise NotImplementedError()


@hparams.register_dataset('toy')
class WhiteNoiseData(Dataset):
    def epoch(self, subset, batch_size, shuffle=False):
        if not self.is_loaded:
            raise RuntimeError('Dataset is not loaded.')
        for _ in range(10):
            yield (np.random.rand(batch_size, 128, hparams.FEATURE_SIZE).astype(hparams.FLOATX),)

    def install_and_load(self):
        self.is_loaded = True


@hparams.register_dataset('synthetic')
class SyntheticSpeechLike(Dataset):
    N_BATCHES = 10
    SECONDS = 4.

    def epoch(self, subset, batch_size, shuffle=False):
        if not self.is_loaded:
            raise RuntimeError('Dataset is not loaded.')
        n = int(self.SECONDS * hparams.SMPRATE)
        seed = dict(train=0, valid=1, test=2).get(subset, 3)
        g = torch.Generator(device='cuda').manual_seed(1337 + seed)
        t = torch.arange(n, device='cuda') / float(hparams.SMPRATE)
        for _ in range(self.N_BATCHES):
            x = torch.randn(batch_size, n, device='cuda', generator=g)
            ph = torch.rand(batch_size, 1, device='cuda', generator=g) * 2 * np.pi
            x = x * (0.5 - 0.5 * torch.cos(2 * np.pi * 4. * t + ph))
            x = x * (1000. / x.pow(2).mean(-1, keepdim=True).sqrt())
            yield (K.stft(x),)            # complex64 [batch_size, T, F] on the device

    def install_and_load(self):
        self.is_loaded = True
