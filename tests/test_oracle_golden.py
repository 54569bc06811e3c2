"""
Pins oracle/danet_oracle.py against fixtures produced by the reference's own
Python (tests/golden/make_golden.py) and against scipy.signal.stft live.
CPU only.
"""
import glob
import json
import os

import numpy as np
import pytest
import scipy.signal
import torch

from oracle import danet_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def rel(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


@pytest.fixture(scope='module')
def audio():
    return np.load(os.path.join(GOLDEN, 'audio.npz'))


def test_window_matches_reference(audio):
    assert np.array_equal(O.fft_window(256), audio['window'])
    w = O.fft_window(256, np.float64)
    assert w[0] == 0. and w[255] == 0.
    assert abs(w.sum() - 162.336) < 1e-2


@pytest.mark.parametrize('n', [32000, 31999, 4096, 777, 256])
def test_stft_golden(audio, n):
    Z = O.stft(audio['wav_%d' % n])
    assert Z.shape == audio['stft_%d' % n].shape == (O.num_frames(n), 129)
    # the fixture wavs are float32, so scipy ran its float32 pocketfft: 1e-6, not 1e-12
    assert rel(Z, audio['stft_%d' % n]) < 1e-6


def test_stft_scipy_live():
    rs = np.random.RandomState(0)
    for n in (256, 257, 320, 1000, 8000):
        w = rs.standard_normal(n) * 1000.
        Z = scipy.signal.stft(w, window=O.fft_window(256), nperseg=256, noverlap=192)[2].T
        assert rel(O.stft(w), Z) < 1e-12


def test_stft_too_short_raises():
    with pytest.raises(ValueError):
        O.stft(np.zeros(100))


def test_stft_sinusoid_known_answer(audio):
    # bin-centred cosine, amplitude 1000 at bin 16: interior frames have
    # |Z[16]| = A/2 * sum(w)/sum(w) ... scaled by the window sum -> A/2
    Z = O.stft(audio['wav_sin'])
    assert rel(Z, audio['stft_sin']) < 1e-6
    mid = np.abs(Z[8:20])
    # symmetric (non-periodic) sqrt-hann leaks a little: 499.889 instead of 500
    assert np.allclose(mid[:, 16], 500., rtol=5e-4)
    assert mid[:, 24:].max() < 1e-2 * 500 and np.all(np.argmax(mid, axis=1) == 16)


@pytest.mark.parametrize('key', ['4096', '777', 'rand'])
def test_istft_golden(audio, key):
    if key == 'rand':
        X, ref = audio['istft_rand_in'], audio['istft_rand_out']
    else:
        X, ref = audio['stft_' + key].astype(np.complex64), audio['istft_' + key]
    out = O.istft(X)
    assert out.shape == ref.shape == (64 * X.shape[0],)
    assert rel(out, ref) < 1e-6


def test_istft_is_delayed_scaled_identity(audio):
    # y[128+k] * sum(w) == x[k]
    x = audio['wav_4096'].astype(np.float64)
    y = O.istft(O.stft(x))
    sw = O.fft_window(256, np.float64).sum()
    n = 64 * (y.shape[0] // 64 - 4) - 256
    assert np.allclose(y[128 + 64:128 + n] * sw, x[64:n], rtol=0, atol=1e-6 * np.abs(x).max() * sw / 100)


def test_toy_dataset_first_batch(audio):
    np.random.seed(1337)
    a = np.random.rand(4, 128, 129).astype(np.float32)
    assert np.array_equal(a, audio['toy_first_batch'])


def test_ops_golden():
    d = np.load(os.path.join(GOLDEN, 'ops.npz'))
    x, y = torch.from_numpy(d['pit_x']), torch.from_numpy(d['pit_y'])
    loss, perms, idx, _ = O.pit_mse_loss(x, y)
    assert abs(float(loss) - float(d['pit_c_loss'])) < 1e-12
    assert np.array_equal(perms.numpy(), d['pit_c_perms'])
    assert np.array_equal(idx.numpy(), d['pit_c_idx'])
    loss, _, idx, _ = O.pit_mse_loss(x.abs(), y.abs())
    assert abs(float(loss) - float(d['pit_r_loss'])) < 1e-12
    assert np.array_equal(idx.numpy(), d['pit_r_idx'])
    assert rel(O.batch_snr(x, y).numpy(), d['snr_c']) < 1e-12
    assert rel(O.batch_snr(x.abs(), y.abs()).numpy(), d['snr_r']) < 1e-12
    a = torch.from_numpy(d['comb_in'])
    assert np.array_equal(a[O.combinations(6, 2)].numpy(), d['comb_2'])
    assert np.array_equal(a[O.combinations(6, 3)].numpy(), d['comb_3'])
    c1, h1 = O.lstm_cell(torch.from_numpy(d['lstm_x']), torch.from_numpy(d['lstm_c0']),
                         torch.from_numpy(d['lstm_h0']), torch.from_numpy(d['lstm_W']),
                         torch.from_numpy(d['lstm_B']))
    assert rel(c1.numpy(), d['lstm_c1']) < 1e-13
    assert rel(h1.numpy(), d['lstm_h1']) < 1e-13


def test_pit_swap_invariance():
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.standard_normal((4, 2, 6, 5)))
    y = x.flip(1) + 0.01 * torch.from_numpy(rs.standard_normal((4, 2, 6, 5)))
    loss, perms, idx, _ = O.pit_mse_loss(x, y)
    assert np.all(idx.numpy() == 1)
    l2, _, idx2, _ = O.pit_mse_loss(x, y.flip(1))
    assert np.all(idx2.numpy() == 0) and abs(float(loss) - float(l2)) < 1e-15


MODEL_FILES = sorted(glob.glob(os.path.join(GOLDEN, 'model_*.npz')))


def load_case(path):
    d = np.load(path)
    meta = json.loads(str(d['meta']))
    over = meta['over']
    est = [v[0].split('/')[1] for v in meta['var_order'] if v[0].endswith('anchors:0')]
    P = O.reference_init(
        meta['seed'], encoder=over.get('ENCODER_TYPE', 'toy'),
        embed=over.get('EMBED_SIZE', 20), estimators=tuple(est))
    return d, meta, over, P


@pytest.mark.parametrize('path', MODEL_FILES, ids=[os.path.basename(p)[6:-4] for p in MODEL_FILES])
def test_model_forward_golden(path):
    d, meta, over, P = load_case(path)
    src = torch.from_numpy(d['src']).to(torch.complex128)
    for p in P.values():
        p.requires_grad_(True)
    out = O.model_forward(
        src, P, encoder=over.get('ENCODER_TYPE', 'toy'),
        train_est=over.get('TRAIN_ESTIMATOR_METHOD', 'truth-weighted'),
        infer_est=over.get('INFER_ESTIMATOR_METHOD', 'anchor'),
        sep=over.get('SEPARATOR_TYPE', 'dot-sigmoid-orig'),
        embed=over.get('EMBED_SIZE', 20))
    tol = 1e-9
    assert rel(out['embed'].detach().numpy(), d['dbg_embed']) < tol
    assert rel(out['attrs'].detach().numpy(), d['dbg_attrs']) < tol
    # the separator runs twice when train/infer estimators differ and the second call
    # overwrites debug_fetches['masks'] (modules.py:570-571, main.py:276-278)
    assert rel(out['masks_valid'].detach().numpy(), d['dbg_masks']) < tol
    assert rel(out['output'].detach().numpy(), d['dbg_output']) < tol
    assert rel(out['infer_signals'].detach().numpy(), d['infer_signals']) < tol
    for k in ('train_loss', 'train_snr', 'valid_loss', 'valid_snr'):
        assert abs(float(out[k].detach()) - float(d[k])) <= tol * max(1., abs(float(d[k]))), k
    # gradients of the train loss (main.py:357-358), sampled
    names = meta['grad_names']
    keys = [n.replace('global/', '').replace(':0', '') for n in names]
    grads = torch.autograd.grad(out['train_loss'], [P[k] for k in keys])
    for n, g in zip(names, grads):
        key = n.replace('/', '.').replace(':0', '')
        g = g.numpy()
        ref_l2 = float(d['grad_l2.' + key])
        assert abs(np.sqrt((g * g).sum()) - ref_l2) <= 1e-8 * max(ref_l2, 1e-30), n
        got = g.reshape(-1)[d['grad_idx.' + key]]
        assert np.abs(got - d['grad_val.' + key]).max() <= 1e-8 * max(np.abs(d['grad_val.' + key]).max(), 1e-30), n


def test_clip_adam_golden():
    d, meta, over, P = load_case(os.path.join(GOLDEN, 'model_bilstm_anchor_softmax_c2.npz'))
    src = torch.from_numpy(d['src']).to(torch.complex128)
    for p in P.values():
        p.requires_grad_(True)
    out = O.model_forward(src, P, encoder='bilstm-orig', train_est='anchor', infer_est='anchor',
                          sep='dot-softmax-orig')
    keys = ['encoder/output/W', 'train_estimator/anchors']
    grads = torch.autograd.grad(out['train_loss'], [P[k] for k in keys])
    params = {k: P[k].detach().clone() for k in keys}
    m = {k: torch.zeros_like(params[k]) for k in keys}
    v = {k: torch.zeros_like(params[k]) for k in keys}
    O.clip_adam_step(params, dict(zip(keys, grads)), m, v, step=1)
    for k in keys:
        key = 'global.' + k.replace('/', '.')
        got = params[k].numpy().reshape(-1)[d['adam_idx.' + key]]
        assert np.abs(got - d['adam_val.' + key]).max() < 1e-12
