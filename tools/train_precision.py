"""Training step with the forward recurrence's state as fp16 (Model.TRAIN_RECURRENT_FP16) against bf16x3: gradient error
of every variable against float64 autograd on the oracle (B = 4, 128 frames: the reference's default crop), and the step
time at cfg 2 (B = 32, T = 501)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import danet_tensorflow_b200 as D
import bench
from oracle import danet_oracle as O
K = D.kernels
D.hparams.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
                    SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=4)); D.hparams.digest()
srcw = bench.synth_sources(4, 64 * 127, 7)
src_np = np.stack([[O.stft(w) for w in u] for u in srcw]).astype(np.complex64)
P = O.reference_init(1337, estimators=('train_estimator',), dtype=torch.float64)
Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
ref = O.model_forward(torch.from_numpy(src_np).to(torch.complex128), Pg, encoder='bilstm-orig', train_est='anchor',
                      infer_est='anchor', sep='dot-softmax-orig', embed=20)
ref['train_loss'].backward()
for fp16 in (0, 1):
    D.Model.TRAIN_RECURRENT_FP16 = bool(fp16)
    m = D.Model('tp', 'cuda:0').build()
    m.load_params(P)
    out = m.train_forward_backward(torch.from_numpy(src_np).cuda())
    worst = 0.
    rows = []
    for k, v in Pg.items():
        g = m.grads[k].double().cpu()
        e = float((g - v.grad).abs().max() / (v.grad.abs().max() + 1e-300))
        rows.append((e, k))
        worst = max(worst, e)
    rows.sort(reverse=True)
    print('forward state fp16 = %d: loss rel err %.2e, worst gradient max-norm rel err %.2e (%s); next %.2e (%s)' % (
        fp16, abs(float(out['loss']) - float(ref['train_loss'])) / abs(float(ref['train_loss'])), rows[0][0], rows[0][1],
        rows[1][0], rows[1][1]))
D.hparams.load(dict(BATCH_SIZE=32)); D.hparams.digest()
src = K.stft(torch.from_numpy(bench.synth_sources(32, 32000, 1337)).cuda())
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for fp16 in (0, 1, 0, 1):
    D.Model.TRAIN_RECURRENT_FP16 = bool(fp16)
    m = D.Model('tt', 'cuda:0').build()
    for _ in range(3):
        m.train_step(src)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.train_step(src); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print('forward state fp16 = %d: train step %.3f ms (%.0f mixtures/s)' % (fp16, np.median(ts), 32 / np.median(ts) * 1e3))
    del m
D.Model.TRAIN_RECURRENT_FP16 = None
