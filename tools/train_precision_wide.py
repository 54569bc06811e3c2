"""lstm-orig (4 x 600) training gradients: the wide tcgen05 kernels (forward: fp16 state + fp8 residual weights; backward:
fp16 weights, scaled fp16 hi/lo da) against the exact fp32 kernels, per variable, relative to the variable's largest entry."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import danet_tensorflow_b200 as D
K = D.kernels
B = 8
hp = D.hparams
hp.load(dict(ENCODER_TYPE='lstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=B)); hp.digest()
g = torch.Generator(device='cuda').manual_seed(0)
src = K.stft(torch.randn(B, 2, 32000, device='cuda', generator=g) * 1000.)
D.Model.TRAIN_GROUPS, D.Model.TRAIN_GRAPH = 1, False
grads = {}
for fp16 in (False, True):
    D.Model.TRAIN_RECURRENT_FP16 = fp16
    m = D.Model('p%d' % fp16, 'cuda:0', seed=1337).build()
    out = m.train_forward_backward(src)
    torch.cuda.synchronize()
    grads[fp16] = {k: v.clone() for k, v in m.grads.items()}
    print('fp16 kernels %d: loss %.6f' % (fp16, float(out['loss'])))
worst = 0.
for k in grads[False]:
    a, b = grads[True][k], grads[False][k]
    if float(b.abs().max()) == 0.:
        continue
    e = float((a - b).abs().max() / b.abs().max())
    l2 = float((a - b).norm() / b.norm())
    worst = max(worst, e)
    print('%-40s max %.2e of the largest entry   L2 %.2e' % (k, e, l2))
print('worst %.2e' % worst)
D.Model.TRAIN_RECURRENT_FP16 = None
