"""A/B of the stream-group count of Model.separate inside ONE process (graph replay, B = 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import danet_tensorflow_b200 as D
hp = D.hparams
hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=32)); hp.digest()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
wav = torch.from_numpy(bench.synth_mixtures(B, 32000, 1)).cuda()
D.Model.LSTM_PRIORITY_STREAM = bool(int(os.environ.get('AB_LSTM_PRIO', '0')))
D.Model.USE_FUSED_K4 = bool(int(os.environ.get('AB_FUSED', '1')))
m = D.Model('ab', 'cuda:0').build()
variants = [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else '1,2,3,4'.split(','))]
for g in variants:
    for _ in range(3):
        m.separate_graphed(wav, groups=g)
torch.cuda.synchronize()
res = {g: [] for g in variants}
for rep in range(5):
    for g in variants:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m.separate_graphed(wav, groups=g)
        e1.record(); torch.cuda.synchronize()
        res[g].append(e0.elapsed_time(e1) / 10)
for g, v in res.items():
    med = sorted(v)[len(v) // 2]
    print('groups %d: %s  median %.3f ms  -> %.0f mixtures/s' % (g, ' '.join('%.3f' % x for x in v), med, B / med * 1e3))
