"""Inference rate (Model.separate, graph replay, B = 32, ~4 s mixtures) of every registered encoder."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import danet_tensorflow_b200 as D
hp = D.hparams
B, N = 32, 499 * 64            # T = 500 frames: a multiple of 4, as conv-bilstm-v1 needs
wav = torch.from_numpy(bench.synth_mixtures(B, N, 1)).cuda()
for enc in ('bilstm-orig', 'lstm-orig', 'conv-bilstm-v1', 'toy'):
    hp.load(dict(ENCODER_TYPE=enc, TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
                 SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=B)); hp.digest()
    m = D.Model(enc, 'cuda:0').build()
    for _ in range(3):
        m.separate_graphed(wav)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m.separate_graphed(wav)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 5)
    med = sorted(ts)[len(ts) // 2]
    print('%-16s %8.3f ms per batch of %d  -> %8.0f mixtures/s   (%d parameters)' % (enc, med, B, B / med * 1e3, m.parameter_count()))
    del m
