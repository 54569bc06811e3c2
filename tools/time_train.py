"""Timing of one training step (forward + backward + clip/Adam) at cfg 2: B = 32, 2 speakers, T = 501."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import danet_tensorflow_b200 as D
K = D.kernels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hp = D.hparams
hp.load(dict(ENCODER_TYPE=os.environ.get('ENC', 'bilstm-orig'), TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=B)); hp.digest()
D.Model.TRAIN_GRAPH, D.Model.TRAIN_GROUPS = False, 1      # per-kernel events need eager one-pass steps
model = D.Model('t', 'cuda:0').build()
g = torch.Generator(device='cuda').manual_seed(0)
wav = torch.randn(B, 2, 32000, device='cuda', generator=g) * 1000.
src = K.stft(wav)
events = collections.OrderedDict(); on = [False]
def wrap(name):
    f = getattr(K, name)
    def w(*a, **kw):
        if not on[0]: return f(*a, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **kw); e1.record()
        events.setdefault(name, []).append((e0, e1)); return r
    setattr(K, name, w)
for n in ('mix_features', 'center', 'linear', 'lstm_seq', 'attractor_anchor', 'mask_cmul', 'pit_mse', 'head_bwd', 'gemm',
          'lstm_seq_bwd', 'colsum', 'clip_adam'):
    wrap(n)
for _ in range(2):
    out = model.train_step(src)
torch.cuda.synchronize()
print('loss', float(out['loss']), 'snr', float(out['snr']), 'params', model.parameter_count())
on[0] = True
steps = 3
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(steps):
    out = model.train_step(src)
t1.record(); torch.cuda.synchronize()
for k, ev in events.items():
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    print('%-18s calls/step %3d  ms/step %8.3f' % (k, len(ev) // steps, ms))
ms = t0.elapsed_time(t1) / steps
print('train step %.3f ms -> %.1f mixtures/s; loss %.5g' % (ms, B / ms * 1e3, float(out['loss'])))
print('peak mem GB', torch.cuda.max_memory_allocated() / 2**30)
