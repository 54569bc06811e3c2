"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol that
include/danet.h declares, the hyper-parameter / registry surface behaves like the reference's
(app/hparams.py), and the product refuses to run without the device library."""
import ctypes
import os
import re

import numpy as np
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


def test_wide_recurrent_kernel_sizes_without_gpu(D):
    """384 < H <= 608 (lstm-orig): the pack / workspace size queries answer for the wide tcgen05 kernel, nothing beyond"""
    lib = D._lib.load()
    n19 = 19 * (3 * 19 * 128 * 8 + 128) * 4                   # 19 CTAs x (57 slices x 128 rows x 8 words + 128 row scales)
    assert lib.danet_lstm_pack_wh_bytes(1, 600) == n19
    assert lib.danet_lstm_pack_wh_bytes(2, 608) == 2 * n19
    assert lib.danet_lstm_pack_wh_bytes(1, 640) == 0 and lib.danet_lstm_pack_wh_bytes(1, 300) > 0
    # exchange buffer (2 parities x 19 producers x 128 LL words of 8 bytes per group of 8 utterances) + room for an image
    assert lib.danet_lstm_seq_workspace_bytes(1, 32, 600) >= 4 * 2 * 19 * 128 * 8 + n19


def test_training_slices(D):
    """Model.TRAIN_GROUPS / TRAIN_GROUP_MIN: how many stream groups a training batch is cut into"""
    m = D.Model.__new__(D.Model)
    assert [m._train_group_count(b) for b in (1, 8, 15, 16, 32, 256)] == [1, 1, 1, 2, 2, 2]
    m.TRAIN_GROUPS, m.TRAIN_GROUP_MIN = 4, 4
    assert [m._train_group_count(b) for b in (3, 8, 19, 32)] == [1, 2, 4, 4]
    m.TRAIN_GROUPS = 0
    assert m._train_group_count(32) == 1


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
    assert {'toy', 'lstm-orig', 'bilstm-orig', 'conv-bilstm-v1'} <= set(H.encoder_registry)   # app/modules.py: every registered encoder
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


# ---------------------------------------------------------------- TensorFlow checkpoint bundles (tf_bundle.py)
def test_tf_bundle_primitives_known_answers():
    from danet_tensorflow_b200 import tf_bundle as TB
    assert TB._crc32c_py(b'123456789') == 0xE3069283           # the CRC-32C check value (RFC 3720 appendix B.4)
    assert TB._crc32c_py(b'\x00' * 32) == 0x8A9136AA            # RFC 3720: 32 bytes of zeros
    assert TB._crc32c_py(b'\xff' * 32) == 0x62A8AB43            # RFC 3720: 32 bytes of ones
    blob = bytes(range(256)) * 5
    assert TB.crc32c(blob) == TB._crc32c_py(blob)              # the library's host helper agrees with the table walk
    for c in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert TB.unmask_crc(TB.mask_crc(c)) == c
    assert TB.mask_crc(0) == 0xA282EAD8                        # leveldb/util/crc32c.h kMaskDelta
    assert TB.put_varint(300) == b'\xac\x02' and TB.get_varint(b'\xac\x02', 0) == (300, 2)
    assert TB.MAGIC == 0xdb4775248b80fb57                      # leveldb/table/format.h kTableMagicNumber
    raw = b'hello hello hello hello'
    # a hand-assembled snappy stream: literal "hello " then a copy of 17 bytes at offset 6
    comp = TB.put_varint(len(raw)) + bytes([(6 - 1) << 2]) + b'hello ' + bytes([((17 - 1) << 2) | 2, 6, 0])
    assert TB._snappy_uncompress(comp) == raw


def test_tf_bundle_round_trip(tmp_path):
    from danet_tensorflow_b200 import tf_bundle as TB
    rs = np.random.RandomState(0)
    t = {'global/encoder/lstm%d_fwd/LSTM/linear/W' % i: rs.standard_normal((17 + i, 8)).astype(np.float32) for i in range(150)}
    t['global/train_estimator/anchors'] = rs.standard_normal((6, 20)).astype(np.float32)
    t['learn_rate'] = np.float32(3e-4)
    t['global_step'] = np.int64(7)
    prefix = str(tmp_path / 'sub' / 'ckpt')
    TB.write_bundle(prefix, t, entries_per_block=16)           # several data blocks + an index block with several entries
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    with open(prefix + '.index', 'rb') as f:
        assert f.read()[-8:] == (0xdb4775248b80fb57).to_bytes(8, 'little')
    r = TB.read_bundle(prefix)
    assert set(r) == set(t)
    for k in t:
        assert r[k].dtype == np.asarray(t[k]).dtype and r[k].shape == np.asarray(t[k]).shape
        assert np.array_equal(r[k], t[k])
    # corruption is detected: flip one byte of a tensor
    with open(prefix + '.data-00000-of-00001', 'r+b') as f:
        f.seek(5)
        b = f.read(1)
        f.seek(5)
        f.write(bytes([b[0] ^ 0x40]))
    with pytest.raises(IOError):
        TB.read_bundle(prefix)
