"""CPU tests of the on-disk dataset readers: the reference's TIMIT pickle triple
(app/datasets/timit.py:31-113) and its WSJ0 HDF5 layout (app/datasets/WSJ0/process.py:146-223, wsj0.py:22-59),
on small files written here in exactly those formats."""
import os
import pickle

import numpy as np
import pytest


@pytest.fixture()
def D():
    import danet_tensorflow_b200 as D
    D.hparams.__dict__.clear()
    D.hparams.__dict__.update(D.Hyperparameter().__dict__)
    D.hparams.digest()
    return D


def _utterances(rs, lengths):
    return [(rs.standard_normal((t, 129)) + 1j * rs.standard_normal((t, 129))).astype(np.complex64) for t in lengths]


def _write_timit(dirname, rs, lengths):
    sig = _utterances(rs, lengths)
    pho = [rs.randint(0, 60, size=rs.randint(3, 9)) for _ in lengths]
    txt = [rs.randint(0, 27, size=rs.randint(2, 12)).astype(np.int32) for _ in lengths]
    for subset in ('train', 'test'):
        with open(os.path.join(dirname, '%s_set.pkl' % subset), 'wb') as f:       # three consecutive pickles (process.py)
            for part in (sig, pho, txt):
                pickle.dump(part, f)
    return sig, txt


def test_timit_reader_batches_like_the_reference(D, tmp_path, monkeypatch):
    rs = np.random.RandomState(0)
    lengths = [7, 12, 9, 5, 11, 8, 10, 6, 13, 4, 9]                # 11 utterances
    sig, txt = _write_timit(str(tmp_path), rs, lengths)
    monkeypatch.setenv('DANET_TIMIT_DIR', str(tmp_path))
    assert 'timit' in D.Hyperparameter.dataset_registry
    ds = D.Hyperparameter.dataset_registry['timit']()
    with pytest.raises(RuntimeError):
        next(ds.epoch('train', 4))
    ds.install_and_load()
    assert ds.subset['valid'] is ds.subset['test']
    with pytest.raises(KeyError):
        next(ds.epoch('nope', 4))
    batches = list(ds.epoch('train', 4))
    # range(0, 11 - 4, 4) -> two full batches, then the ragged tail re-uses the LAST four utterances
    assert len(batches) == 3
    for k, (spec, (idx, val, shape)) in enumerate(batches[:2]):
        members = list(range(4 * k, 4 * k + 4))
        assert spec.shape == (4, max(lengths[j] for j in members), 129) and spec.dtype == np.complex64
        for row, j in zip(spec, members):
            # random split of the zero padding: the utterance appears once, contiguously, the rest is zero
            nz = np.flatnonzero(np.abs(row).sum(-1))
            assert len(nz) == lengths[j] and nz[-1] - nz[0] + 1 == lengths[j]
            assert np.array_equal(row[nz[0]:nz[-1] + 1], sig[j])
        assert shape == (4, max(len(txt[j]) for j in members))
        assert np.array_equal(val, np.concatenate([txt[j] for j in members]))
        assert np.array_equal(idx[:len(txt[members[0]]), 1], np.arange(len(txt[members[0]]))) and idx[-1, 0] == 3
    tail = batches[2][0]
    assert tail.shape == (4, max(lengths[-4:]), 129)
    for row, j in zip(tail, range(7, 11)):
        assert np.array_equal(row[:lengths[j]], sig[j]) and not np.abs(row[lengths[j]:]).any()   # padded at the end
    # a multiple of the batch size: the reference's range() drops the last full batch (timit.py:44)
    ds.subset['train'] = [part[:8] for part in ds.subset['train']]
    assert len(list(ds.epoch('train', 4))) == 1
    order = [b[0].shape for b in ds.epoch('test', 5, shuffle=True)]
    assert len(order) == 3
    assert ds.decode_to_str(ds.encode_from_str('hello world')) == 'hello world'


def test_timit_missing_file_raises_ioerror(D, tmp_path, monkeypatch):
    monkeypatch.setenv('DANET_TIMIT_DIR', str(tmp_path))
    with pytest.raises(IOError):
        D.Hyperparameter.dataset_registry['timit']().install_and_load()


class _FakeH5(dict):
    """the slice of the h5py.File interface the reader uses: item access by dataset name and `.attrs`"""
    def __init__(self, arrays, attrs):
        super().__init__(arrays)
        self.attrs = attrs


def test_wsj0_reader_on_the_reference_layout(D):
    rs = np.random.RandomState(1)
    arrays, rows = {}, []
    ref = {}
    for split, lengths in (('train', [6, 9, 4, 7, 5]), ('valid', [3, 8]), ('test', [5])):
        utt = _utterances(rs, lengths)
        ref[split] = utt
        arrays['%s_spectra' % split] = [u.reshape(-1) for u in utt]            # vlen complex64, flattened (process.py:181)
        arrays['%s_spectra_shapes' % split] = np.array([[len(u), 129] for u in utt], dtype=np.int32)
        rows.append((split.encode('utf8'), ('%s_spectra' % split).encode('utf8'), 0, len(utt)))
    split_attr = np.array(rows, dtype=[('split', 'S5'), ('source', 'S15'), ('start', np.int64), ('stop', np.int64)])
    ds = D.Hyperparameter.dataset_registry['wsj0']()
    ds.install_and_load(_FakeH5(arrays, {'split': split_attr}))
    batches = list(ds.epoch('train', 2))
    assert len(batches) == 3                                    # ceil(5 / 2) full batches, indices wrap around
    seen = []
    for (spec,), members in zip(batches, ([0, 1], [2, 3], [4, 0])):
        assert spec.shape == (2, max(len(ref['train'][j]) for j in members), 129) and spec.dtype == np.complex64
        for row, j in zip(spec, members):
            nz = np.flatnonzero(np.abs(row).sum(-1))
            assert np.array_equal(row[nz[0]:nz[-1] + 1], ref['train'][j])
            seen.append(j)
    assert sorted(set(seen)) == [0, 1, 2, 3, 4]
    assert [b[0].shape[0] for b in ds.epoch('valid', 2, shuffle=True)] == [2]
    assert len(list(ds.epoch('test', 3))) == 1


def test_wsj0_without_h5py_says_so(D, monkeypatch, tmp_path):
    try:
        import h5py  # noqa: F401
        pytest.skip('h5py is installed here')
    except ImportError:
        pass
    with pytest.raises(ImportError):
        D.Hyperparameter.dataset_registry['wsj0']().install_and_load()
