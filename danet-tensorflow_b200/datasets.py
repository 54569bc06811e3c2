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


def random_zeropad(x, padlen, axis=-1, rs=np.random):
    """app/utils.py:78-92: zero padding of total length `padlen` on `axis`, split at random between the two ends"""
    if padlen == 0:
        return x
    left = rs.randint(0, padlen + 1)               # the reference's random.randint is inclusive on both ends
    pad = [(0, 0)] * x.ndim
    pad[axis % x.ndim] = (left, padlen - left)
    return np.pad(x, pad, mode='constant')


def _data_dir(name):
    import os
    env = os.environ.get('DANET_%s_DIR' % name.upper())
    return env or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', name)


@hparams.register_dataset('timit')
class TimitDataset(Dataset):
    """Reader of the reference's preprocessed TIMIT files (app/datasets/timit.py:90-113): `train_set.pkl` and
    `test_set.pkl`, each THREE consecutive pickles -- a list of complex64 spectra [T_i, 129] (scipy.signal.stft of the
    8 kHz utterance, what danet_stft_fwd computes; TIMIT/process.py:85-104), a list of phoneme index arrays and a list
    of character index arrays.  `valid` is the test set (timit.py:116-118).  The directory is $DANET_TIMIT_DIR or
    <package>/data/TIMIT; the corpus itself is licensed and not redistributable (`preprocess_wavs` below builds the
    spectra list from wav files on the GPU).

    `epoch` reproduces the reference's batching (timit.py:31-88), quirks included: full batches start at
    range(0, n - batch, batch), so when n is a multiple of the batch size the last full batch is skipped; a ragged tail
    re-uses the LAST `batch` utterances, zero-padded at the end; utterances of a full batch are zero-padded at random
    on both ends of the time axis to the batch maximum.  Yields `(spectra [batch, T, 129], (text_indices, text_values,
    text_shape))` -- main.py consumes only element 0 (main.py:414-421)."""
    CHARSET = 'abcdefghijklmnopqrstuvwxyz '

    def install_and_load(self):
        import gc
        import os
        import pickle
        self.subset = {}
        for subset in ('train', 'test'):
            path = os.path.join(_data_dir('TIMIT'), '%s_set.pkl' % subset)
            if not os.path.exists(path):
                raise IOError('Did not found TIMIT file "%s", make sure you download and install the dataset' % path)
            with open(path, 'rb') as f:
                gc.disable()
                try:
                    self.subset[subset] = [pickle.load(f) for _ in range(3)]      # spectra, phonemes, texts
                finally:
                    gc.enable()
        self.subset['valid'] = self.subset['test']
        self.is_loaded = True

    @staticmethod
    def _sparse_text(texts, batch_size):
        n = sum(len(t) for t in texts)
        idx = np.empty((n, 2), dtype=hparams.INTX)
        pos = 0
        for j, t in enumerate(texts):
            idx[pos:pos + len(t), 0] = j
            idx[pos:pos + len(t), 1] = np.arange(len(t))
            pos += len(t)
        values = np.concatenate(texts) if texts else np.zeros((0,), dtype=hparams.INTX)
        return idx, values, (batch_size, max((len(t) for t in texts), default=0))

    def epoch(self, subset, batch_size, shuffle=False):
        if not self.is_loaded:
            raise RuntimeError('Dataset is not loaded.')
        if subset not in self.subset:
            raise KeyError('Unknown subset "%s", valid options are %s' % (subset, list(self.subset.keys())))
        signals, phonemes, texts = self.subset[subset]
        n = len(signals)
        assert n == len(phonemes) == len(texts)
        order = np.random.permutation(n) if shuffle else np.arange(n)
        for i in range(0, n - batch_size, batch_size):
            sel = order[i:i + batch_size]
            sig = [signals[j] for j in sel]
            longest = max(len(s) for s in sig)
            batch = np.stack([random_zeropad(s, longest - len(s), axis=-2) for s in sig])
            yield batch, self._sparse_text([texts[j] for j in sel], batch_size)
        if n % batch_size:
            sel = order[-batch_size:]
            sig = [signals[j] for j in sel]
            # the reference pads to len(signals[-1]) (timit.py:73), which only works when that utterance is the longest
            # of the tail; the batch maximum is the same value whenever the reference's own call succeeds
            longest = max(len(signals[-1]), max(len(s) for s in sig))
            batch = np.stack([np.pad(s, ((0, longest - len(s)), (0, 0)), mode='constant') for s in sig])
            yield batch, self._sparse_text([texts[j] for j in sel], batch_size)

    @classmethod
    def encode_from_str(cls, s):
        return np.asarray([cls.CHARSET.index(c) for c in s], dtype='int32')

    @classmethod
    def decode_to_str(cls, arr):
        return ''.join(cls.CHARSET[int(i)] for i in arr)


@hparams.register_dataset('wsj0')
class Wsj0Dataset(Dataset):
    """Reader of the reference's `wsj0-danet.hdf5` (written by WSJ0/process.py:146-223, read through fuel's H5PYDataset at
    app/datasets/wsj0.py:22-59): per split a variable-length complex64 dataset `<split>_spectra` (one flattened [T_i, 129]
    spectrum per utterance), `<split>_spectra_shapes` int32 [n, 2] and the file attribute `split` whose rows carry the
    (start, stop) example range.  Read here with h5py alone -- fuel is a thin index layer over exactly these arrays.
    h5py is not part of this image, so `install_and_load` raises ImportError without it; the batching below (wrap-around
    indices so every batch is full, optional shuffle, random zero padding to the batch maximum: wsj0.py:37-59) is
    covered by the CPU tests through an in-memory stand-in for the h5py file."""

    def install_and_load(self, h5file=None):
        import os
        if h5file is None:
            try:
                import h5py
            except ImportError as e:
                raise ImportError('the wsj0 dataset needs h5py to read wsj0-danet.hdf5 (%s)' % e)
            path = os.path.join(_data_dir('WSJ0'), 'wsj0-danet.hdf5')
            if not os.path.exists(path):
                raise IOError('Did not found WSJ0 file "%s"' % path)
            h5file = h5py.File(path, 'r')
        self.h5file = h5file
        self.sizes = {}
        for row in h5file.attrs['split']:
            name = row['split']
            name = name.decode('utf8') if isinstance(name, bytes) else str(name)
            self.sizes[name] = (int(row['start']), int(row['stop']))
        self.is_loaded = True

    def epoch(self, subset, batch_size, shuffle=False):
        if not self.is_loaded:
            raise RuntimeError('Dataset is not loaded.')
        start, stop = self.sizes[subset]
        n = stop - start
        spectra, shapes = self.h5file['%s_spectra' % subset], self.h5file['%s_spectra_shapes' % subset]
        indices = np.arange((n + batch_size - 1) // batch_size * batch_size) % n
        if shuffle:
            np.random.shuffle(indices)
        for i in range(0, len(indices), batch_size):
            items = []
            for j in indices[i:i + batch_size]:
                t, f = (int(v) for v in shapes[start + j])
                items.append(np.asarray(spectra[start + j]).reshape(t, f))
            longest = max(len(x) for x in items)
            yield (np.stack([random_zeropad(x, longest - len(x), axis=-2) for x in items]),)


def preprocess_wavs(paths, smprate=None):
    """TIMIT/process.py:35-57, 85-104 and WSJ0/process.py:175-179 with the transform on the GPU: wav files -> list of
    complex64 spectra [T_i, 129].  Rate conversion as the reference does it (block mean for integer factors, FFT
    resampling otherwise); the STFT is danet_stft_fwd, the kernel pinned against scipy.signal.stft."""
    import scipy.io.wavfile
    import scipy.signal
    smprate = smprate or hparams.SMPRATE
    out = []
    for path in paths:
        rate, data = scipy.io.wavfile.read(path)
        if rate == smprate:
            data = data.astype(np.float32)
        elif rate % smprate == 0:
            k = rate // smprate
            data = np.pad(data, [(0, (-len(data)) % k)], mode='constant').reshape(-1, k).astype(np.float32).mean(axis=1)
        else:
            data = scipy.signal.resample(data, int(np.ceil(len(data) * smprate / rate))).astype(np.float32)
        out.append(K.stft(torch.from_numpy(np.ascontiguousarray(data[None])).cuda())[0].cpu().numpy())
    return out
