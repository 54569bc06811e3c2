"""
Generates the golden fixtures in this directory by running the UNMODIFIED
reference code (`/root/reference`: main.Model.build, app.modules, app.ops,
app.utils) on seeded inputs under the eager TF1 shim in `oracle/tf1_shim.py`
(TensorFlow itself cannot be installed here), plus
`scipy.signal.stft` -- the very function the reference calls for its STFT.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The GPU box never runs this; it reads the committed .npz files.

Weights are NOT stored (9 M parameters): the shim draws them from
np.random.RandomState(seed) in variable-creation order, which
`oracle.danet_oracle.reference_init` reproduces.  Gradients are stored as
(sum, l2, 256 strided samples) per variable.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import tf1_shim  # noqa: E402

tf1_shim.install()
tf1_shim.DEFAULT_FLOAT = 'float64'
import scipy.signal  # noqa: E402

from app.hparams import hparams  # noqa: E402  (reference module)
import app.utils as ref_utils  # noqa: E402
import app.datasets.dataset as ref_dataset  # noqa: E402
import main as ref_main  # noqa: E402


def shaped_noise(rs, n, rms=1000.):
    """int16-scale coloured noise with a slow envelope (SURVEY §8d cfg 2)"""
    w = rs.standard_normal(n)
    y = np.zeros(n)
    acc = 0.
    for i in range(n):
        acc = 0.9 * acc + w[i]
        y[i] = acc
    env = 0.5 - 0.5 * np.cos(2 * np.pi * (4. * np.arange(n) / 8000. + rs.rand()))
    y = y * env
    return (y * (rms / np.sqrt(np.mean(y * y) + 1e-12))).astype(np.float32)


def ref_stft(wav, wnd):
    # exactly the call at app/utils.py:117-122
    Z = scipy.signal.stft(wav, window=wnd, nperseg=256, noverlap=256 - 64)[2]
    return Z.T


def load_hparams(over):
    hparams.load_json(os.path.join(REF, 'default.json'))
    hparams.load(over)
    hparams.digest()


def sample(a, n=256):
    a = np.asarray(a).reshape(-1)
    idx = np.linspace(0, a.size - 1, num=min(n, a.size)).astype(np.int64)
    return idx, a[idx]


def run_model_case(tag, over, B, C, T, seed, scale=1000., toy_input=False):
    over = dict(over)
    over.update(BATCH_SIZE=B, MAX_N_SIGNAL=C, DEBUG=True, FLOATX='float64')
    load_hparams(over)
    tf1_shim.reset(seed)
    rs = np.random.RandomState(seed + 1)
    if toy_input:
        # app/datasets/dataset.py:55-59 : uniform [0,1) real "spectra"
        src = rs.rand(B * C, T, hparams.FEATURE_SIZE).astype(np.float32)
        src = src.reshape(B, C, T, -1).astype(np.complex128)
    else:
        n = 64 * (T - 1)
        wavs = np.stack([shaped_noise(rs, n, scale) for _ in range(B * C)])
        src = np.stack([ref_stft(w, hparams.FFT_WND) for w in wavs])
        assert src.shape[1] == T, src.shape
        src = src.reshape(B, C, T, -1).astype(np.complex128)
    tf1_shim.FEEDS['source_signal'] = src
    tf1_shim.FEEDS['dropout_keep'] = np.float64(1.)
    model = ref_main.Model(name=tag)
    model.build()
    sess = ref_main.g_sess
    out = dict(src=src.astype(np.complex64))
    dbg = sess.run(model.debug_fetches)
    for k in ('embed', 'attrs', 'output', 'masks', 'asets', 'anchors', 'subset_choice'):
        if k in dbg:
            out['dbg_' + k] = np.asarray(dbg[k])
    tr = sess.run(model.train_fetches)[1]
    va = sess.run(model.valid_fetches)[1]
    out['train_loss'] = np.float64(tr['loss'])
    out['train_snr'] = np.float64(tr['SNR'])
    out['valid_loss'] = np.float64(va['loss'])
    out['valid_snr'] = np.float64(va['SNR'])
    out['infer_signals'] = np.asarray(sess.run(model.infer_fetches)['signals'])
    names = []
    for name, g in tf1_shim.GRADS.items():
        key = name.replace('/', '.').replace(':0', '')
        names.append(name)
        idx, val = sample(g)
        out['grad_sum.' + key] = np.float64(g.sum())
        out['grad_l2.' + key] = np.float64(np.sqrt((g * g).sum()))
        out['grad_idx.' + key] = idx
        out['grad_val.' + key] = val
    oz = tf1_shim.LAST_OPTIMIZER[0]
    if oz is not None:
        for name in list(oz.updated)[-2:]:   # last two variables: output/W (+anchors)
            key = name.replace('/', '.').replace(':0', '')
            idx, val = sample(oz.updated[name])
            out['adam_idx.' + key] = idx
            out['adam_val.' + key] = val
    var_order = [(k, list(v.v.shape), bool(v.trainable))
                 for k, v in tf1_shim.VARIABLES.items()]
    out['meta'] = np.array(json.dumps(dict(
        tag=tag, over=over, B=B, C=C, T=T, seed=seed, var_order=var_order,
        grad_names=names)))
    for k in list(out):
        a = out[k]
        if isinstance(a, np.ndarray) and a.dtype == np.complex128 and k != 'src':
            out[k] = a  # keep f64 truth
    path = os.path.join(HERE, 'model_%s.npz' % tag)
    np.savez_compressed(path, **out)
    print('%-28s loss=%.6g snr=%.4g vloss=%.6g  -> %s (%.0f KB)' % (
        tag, out['train_loss'], out['train_snr'], out['valid_loss'],
        os.path.basename(path), os.path.getsize(path) / 1024.))


def audio_cases():
    load_hparams({})
    rs = np.random.RandomState(7)
    out = {}
    for n in (32000, 31999, 4096, 777, 256):
        w = shaped_noise(rs, n)
        out['wav_%d' % n] = w
        out['stft_%d' % n] = ref_stft(w, hparams.FFT_WND).astype(np.complex128)
    # bin-centred sinusoid known-answer: bin 16 of a 256-point frame
    t = np.arange(2048)
    w = (1000. * np.cos(2 * np.pi * 16. * t / 256.)).astype(np.float32)
    out['wav_sin'] = w
    out['stft_sin'] = ref_stft(w, hparams.FFT_WND).astype(np.complex128)
    # reference iSTFT (app/utils.py:53-75) on reference STFTs
    for n in (4096, 777):
        X = out['stft_%d' % n].astype(np.complex64)
        out['istft_%d' % n] = ref_utils.istft(X, hparams.FFT_STRIDE, hparams.FFT_WND)
    X = (rs.standard_normal((9, 129)) + 1j * rs.standard_normal((9, 129))).astype(np.complex64)
    out['istft_rand_in'] = X
    out['istft_rand_out'] = ref_utils.istft(X, hparams.FFT_STRIDE, hparams.FFT_WND)
    out['window'] = np.asarray(hparams.FFT_WND)
    # toy dataset (app/datasets/dataset.py:43-63) with numpy's global seed
    np.random.seed(1337)
    ds = ref_dataset.WhiteNoiseData()
    ds.install_and_load()
    first = next(iter(ds.epoch('train', 4)))[0]
    out['toy_first_batch'] = first
    path = os.path.join(HERE, 'audio.npz')
    np.savez_compressed(path, **out)
    print('audio -> %s (%.0f KB)' % (os.path.basename(path), os.path.getsize(path) / 1024.))


def ops_cases():
    """direct calls of reference ops on tiny tensors (pins pit/snr/combinations)"""
    import app.ops as ref_ops
    from oracle.tf1_shim import TT
    import torch
    load_hparams(dict(BATCH_SIZE=3, MAX_N_SIGNAL=3, FLOATX='float64'))
    tf1_shim.reset(5)
    rs = np.random.RandomState(11)
    x = rs.standard_normal((3, 3, 5, 7)) + 1j * rs.standard_normal((3, 3, 5, 7))
    y = rs.standard_normal((3, 3, 5, 7)) + 1j * rs.standard_normal((3, 3, 5, 7))
    out = dict(pit_x=x, pit_y=y)
    loss, perms, idx = ref_ops.pit_mse_loss(TT(torch.from_numpy(x)), TT(torch.from_numpy(y)))
    out['pit_c_loss'] = loss.numpy()
    out['pit_c_perms'] = perms.numpy()
    out['pit_c_idx'] = idx.numpy()
    loss, perms, idx = ref_ops.pit_mse_loss(
        TT(torch.from_numpy(np.abs(x))), TT(torch.from_numpy(np.abs(y))))
    out['pit_r_loss'] = loss.numpy()
    out['pit_r_idx'] = idx.numpy()
    out['snr_c'] = ref_ops.batch_snr(TT(torch.from_numpy(x)), TT(torch.from_numpy(y))).numpy()
    out['snr_r'] = ref_ops.batch_snr(
        TT(torch.from_numpy(np.abs(x))), TT(torch.from_numpy(np.abs(y)))).numpy()
    a = rs.standard_normal((6, 4))
    out['comb_in'] = a
    out['comb_2'] = ref_ops.combinations(TT(torch.from_numpy(a)), 2).numpy()
    out['comb_3'] = ref_ops.combinations(TT(torch.from_numpy(a)), 3).numpy()
    # one LSTM cell step, W given
    load_hparams(dict(BATCH_SIZE=2, FLOATX='float64'))
    tf1_shim.reset(9)
    xs = rs.standard_normal((2, 5))
    c0 = rs.standard_normal((2, 4))
    h0 = rs.standard_normal((2, 4))
    c1, h1 = ref_ops.lyr_lstm_flat(
        'cell', TT(torch.from_numpy(xs)), TT(torch.from_numpy(c0)), TT(torch.from_numpy(h0)),
        w_init=tf1_shim.random_uniform_initializer(-.5, .5),
        b_init=tf1_shim.random_uniform_initializer(-.5, .5))
    out.update(lstm_x=xs, lstm_c0=c0, lstm_h0=h0, lstm_c1=c1.numpy(), lstm_h1=h1.numpy(),
               lstm_W=tf1_shim.VARIABLES['cell/linear/W:0'].numpy(),
               lstm_B=tf1_shim.VARIABLES['cell/linear/B:0'].numpy())
    path = os.path.join(HERE, 'ops.npz')
    np.savez_compressed(path, **out)
    print('ops -> %s (%.0f KB)' % (os.path.basename(path), os.path.getsize(path) / 1024.))


def conv_case():
    # the experimental CNN-LSTM hybrid (app/modules.py:263-379); T must be a multiple of 4 (two 2x2 max-pools)
    run_model_case('convbilstm_anchor_softmax', dict(
        ENCODER_TYPE='conv-bilstm-v1', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
        SEPARATOR_TYPE='dot-softmax-orig'), B=2, C=2, T=12, seed=11)


def main():
    if '--only-conv' in sys.argv:
        conv_case()
        return
    audio_cases()
    ops_cases()
    bl = dict(ENCODER_TYPE='bilstm-orig')
    run_model_case('bilstm_tw_anchor_sigmoid', dict(bl), B=2, C=2, T=8, seed=1337)
    run_model_case('bilstm_truth_softmax_toy', dict(
        bl, TRAIN_ESTIMATOR_METHOD='truth', INFER_ESTIMATOR_METHOD='truth',
        SEPARATOR_TYPE='dot-softmax-orig'), B=2, C=2, T=8, seed=1337, toy_input=True)
    run_model_case('bilstm_anchor_softmax_c3', dict(
        bl, TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
        SEPARATOR_TYPE='dot-softmax-orig', EMBED_SIZE=12), B=2, C=3, T=6, seed=21)
    run_model_case('bilstm_thresh_softmax', dict(
        bl, TRAIN_ESTIMATOR_METHOD='truth-threshold', INFER_ESTIMATOR_METHOD='anchor',
        SEPARATOR_TYPE='dot-softmax-orig'), B=2, C=2, T=8, seed=5, scale=60.)
    run_model_case('bilstm_anchor_softmax_c2', dict(
        bl, TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
        SEPARATOR_TYPE='dot-softmax-orig'), B=3, C=2, T=8, seed=33)
    run_model_case('toy_defaults', {}, B=2, C=2, T=8, seed=3, toy_input=True)
    run_model_case('lstm_tw_anchor_sigmoid', dict(ENCODER_TYPE='lstm-orig'),
                   B=2, C=2, T=6, seed=8)
    conv_case()


if __name__ == '__main__':
    main()
