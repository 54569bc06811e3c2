"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol that
include/danet.h declares, the hyper-parameter / registry surface behaves like the reference's
(app/hparams.py), and the product refuses to run without the device library."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def D():
    import danet_tensorflow_b200 as D
    D.build.build()
    return D


def test_library_exports_every_declared_symbol(D):
    header = open(os.path.join(ROOT, 'include', 'danet.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(danet_\w+)\s*\(', header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(D._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(D._lib.PROTOTYPES), declared ^ set(D._lib.PROTOTYPES)
    assert D._lib.load().danet_version() >= 100


def test_sizes_queries_without_gpu(D):
    lib = D._lib.load()
    assert lib.danet_stft_num_frames(32000) == 501
    assert lib.danet_stft_num_frames(64000) == 1001
    assert lib.danet_stft_num_frames(240000) == 3751
    assert lib.danet_stft_num_frames(31999) == 501
    assert lib.danet_stft_num_frames(256) == 5
    assert lib.danet_stft_num_frames(255) < 0
    assert lib.danet_anchor_num_subsets(6, 2) == 15
    assert lib.danet_anchor_num_subsets(6, 3) == 20
    assert lib.danet_attractor_workspace_bytes(32, 30, 20) == 32 * 32 * 30 * 6 * 16
    assert lib.danet_pit_workspace_bytes(32, 2) > 0
    assert lib.danet_lstm_seq_workspace_bytes(2, 32, 300) >= 256


def test_hparams_defaults_and_digest(D):
    hp = D.Hyperparameter()
    assert hp.FFT_SIZE == 256 and hp.FFT_STRIDE == 64 and hp.EMBED_SIZE == 20 and hp.NUM_ANCHOR == 6
    assert hp.TRAIN_ESTIMATOR_METHOD == 'truth-weighted' and hp.INFER_ESTIMATOR_METHOD == 'anchor'
    hp.digest()
    assert hp.FEATURE_SIZE == 129 and hp.COMPLEXX == 'complex64'
    assert hp.FFT_WND.shape == (256,) and hp.FFT_WND.dtype.name == 'float32'
    assert hp.FFT_WND[0] == 0. and abs(float(hp.FFT_WND.sum()) - 162.336) < 1e-2


def test_hparams_load_rejects_bad_keys(D):
    hp = D.Hyperparameter()
    with pytest.raises(NameError):
        hp.load({'lower_case': 1})
    hp.load({'BATCH_SIZE': 8})
    assert hp.BATCH_SIZE == 8


def test_registries_hold_reference_names(D):
    H = D.Hyperparameter
    assert {'lstm-orig', 'bilstm-orig'} <= set(H.encoder_registry)
    assert {'truth', 'truth-threshold', 'truth-weighted', 'anchor', 'kmeans'} <= set(H.estimator_registry)
    assert {'dot-sigmoid-orig', 'dot-softmax-orig'} <= set(H.separator_registry)
    with pytest.raises(KeyError):
        D.hparams.get_estimator('no-such-estimator')
    assert H.estimator_registry['anchor'].USE_TRUTH is False
    assert H.estimator_registry['truth'].USE_TRUTH is True

    @H.register_separator('unit-test-sep')
    class Dummy(D.Separator):
        pass
    assert D.hparams.get_separator('unit-test-sep') is Dummy
    with pytest.raises(NotImplementedError):
        Dummy(None, 'x')(None, None, None)
    del H.separator_registry['unit-test-sep']


def test_no_cpu_path(D):
    import torch
    with pytest.raises(ValueError):
        D.kernels.stft(torch.zeros(1, 1000))
    with pytest.raises(ValueError):
        D.kernels.center(torch.zeros(2, 8))


def test_missing_library_fails_loudly(D, monkeypatch):
    monkeypatch.setattr(D._lib, '_lib', None)
    monkeypatch.setattr(D._lib, 'LIB_PATH', '/nonexistent/libdanet_sm100.so')
    with pytest.raises(RuntimeError, match='no CPU or PyTorch fallback'):
        D._lib.load()
