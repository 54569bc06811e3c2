"""A/B of the inference schedule variants inside ONE process (box-to-box variance is ~2 %)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import danet_tensorflow_b200 as D
hp = D.hparams
hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=32)); hp.digest()
wav = torch.from_numpy(bench.synth_mixtures(32, 32000, 1)).cuda()
variants = [('plain', False, False), ('packed', True, False), ('packed+fold', True, True)]
models = {}
for name, packed, fold in variants:
    D.Model.USE_PACKED, D.Model.USE_CENTER_FOLD = packed, fold
    m = D.Model(name, 'cuda:0').build()
    for _ in range(3):
        m.separate_graphed(wav)
    models[name] = m
torch.cuda.synchronize()
res = {n: [] for n, _, _ in variants}
for rep in range(6):
    for name, packed, fold in variants:
        m = models[name]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m.separate_graphed(wav)
        e1.record(); torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / 10)
for n, v in res.items():
    print('%-12s %s  median %.3f ms' % (n, ' '.join('%.3f' % x for x in v), sorted(v)[len(v) // 2]))
