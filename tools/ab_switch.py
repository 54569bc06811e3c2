"""In-process A/B of Model class switches on the bench step (B = 32, graph replay, L2 flushed between steps), device-resident
and end to end (pinned host in/out).  Usage: python tools/ab_switch.py GROUP_PRIORITIES=0 GROUP_PRIORITIES=1 [...]
Every argument is one variant: comma-separated NAME=int assignments to Model class attributes; each variant gets its own
Model (streams and their priorities are created per model), the variants' steps are interleaved."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import danet_tensorflow_b200 as D
K = D.kernels
hp = D.hparams
hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=32)); hp.digest()
B = 32
wav_np = bench.synth_mixtures(B, 32000, 1)
wav_dev = torch.from_numpy(wav_np).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
variants = sys.argv[1:] or ['GROUP_PRIORITIES=0', 'GROUP_PRIORITIES=1']
defaults = {}
models = []


def apply(spec):
    for kv in spec.split(','):
        k, v = kv.split('=')
        defaults.setdefault(k, getattr(D.Model, k))
        setattr(D.Model, k, type(defaults[k])(int(v)))


for spec in variants:
    for k, v in defaults.items():
        setattr(D.Model, k, v)
    apply(spec)
    m = D.Model('ab', 'cuda:0', seed=1337).build()
    hin = torch.from_numpy(wav_np).pin_memory()
    hout = torch.empty((B, 2, 64 * 501), dtype=torch.float32).pin_memory()
    for _ in range(3):
        m.separate_graphed(wav_dev)
        m.separate_host(hin, hout)
    torch.cuda.synchronize()
    models.append((spec, m, hin, hout))
res = {spec: ([], []) for spec, _, _, _ in models}
for rep in range(7):
    for spec, m, hin, hout in models:
        apply(spec)
        for which, fn in ((0, lambda: m.separate_graphed(wav_dev)), (1, lambda: m.separate_host(hin, hout))):
            ts = []
            for _ in range(5):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[spec][which].append(float(np.mean(ts)))
ref = None
for spec, (dev, e2e) in res.items():
    a, b = float(np.median(dev)), float(np.median(e2e))
    print('%-40s device %.3f ms (%6.0f mixtures/s)   e2e %.3f ms (%6.0f mixtures/s)' % (spec, a, B / a * 1e3, b, B / b * 1e3))
if len(models) > 1:
    o0, o1 = models[0][3], models[1][3]
    print('outputs of the first two variants: max abs diff %.3g' % float((o0 - o1).abs().max()))
