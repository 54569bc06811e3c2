"""Per-call CUDA-event timing of every C-ABI call in one cfg-2 inference step (B = 32, 4 s)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import danet_tensorflow_b200 as D
import bench
K = D.kernels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hp = D.hparams
hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=B))
hp.digest()
model = D.Model('t', 'cuda:0').build()
wav = torch.from_numpy(bench.synth_mixtures(B, 32000, 1)).cuda()
events = collections.OrderedDict()
on = [False]
def wrap(name):
    f = getattr(K, name)
    def g(*a, **kw):
        if not on[0]:
            return f(*a, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **kw); e1.record()
        key = name
        if name == 'linear':
            key = 'linear M%d N%d K%d' % (a[0].shape[0], a[1].shape[1], a[0].shape[1])
        events.setdefault(key, []).append((e0, e1))
        return r
    setattr(K, name, g)
for n in ("stft", "istft", "center", "linear", "lstm_seq", "attractor_anchor", "mask_cmul", "mask_cmul_istft", "gemm_split", "mean", "mix_features"):
    wrap(n)
for _ in range(3):
    model.separate(wav, groups=int(os.environ.get("GROUPS", "1")))
torch.cuda.synchronize()
on[0] = True
steps = 5
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(steps):
    model.separate(wav, groups=int(os.environ.get("GROUPS", "1")))
t1.record()
torch.cuda.synchronize()
tot = 0.
for k, ev in events.items():
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    tot += ms
    print('%-28s calls/step %2d  ms/step %7.3f  ms/call %7.3f' % (k, len(ev) // steps, ms, ms * steps / len(ev)))
print('sum %.3f ms   step wall %.3f ms' % (tot, t0.elapsed_time(t1) / steps))
