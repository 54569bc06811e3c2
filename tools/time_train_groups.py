"""Training step at cfg 2 (B = 32, T = 501) with the batch in 1 / 2 / 4 stream groups (Model.TRAIN_GROUPS): ms per step and
the deviation of the summed gradient from the single-pass one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import danet_tensorflow_b200 as D
K = D.kernels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hp = D.hparams
hp.load(dict(ENCODER_TYPE=os.environ.get('ENC', 'bilstm-orig'), TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=B)); hp.digest()
g = torch.Generator(device='cuda').manual_seed(0)
src = K.stft(torch.randn(B, 2, 32000, device='cuda', generator=g) * 1000.)
ref = None
for groups in (1, 2, 4):
    D.Model.TRAIN_GROUPS = groups
    m = D.Model('t%d' % groups, 'cuda:0', seed=1337).build()
    out = m.train_forward_backward(src)
    torch.cuda.synchronize()
    grad = m._flat['grad'].clone()
    if ref is None:
        ref = grad
    dev = float((grad - ref).abs().max() / ref.abs().max())
    for _ in range(3):
        m.train_step(src)
    torch.cuda.synchronize()
    steps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = m.train_step(src)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print('TRAIN_GROUPS %d: %.3f ms per step -> %6.0f mixtures/s   loss %.6g   gradient vs one pass: %.2e of the largest entry'
          % (groups, ms, B / ms * 1e3, float(out['loss']), dev))
