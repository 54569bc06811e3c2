"""Inference step (B = 32, graph replay, L2 flushed) with the batch in 4 / 5 / 6 / 8 stream groups (Model.separate(groups=...)):
more, smaller groups shorten every group's share of the throughput-bound end of the step but take more SMs for recurrences."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import danet_tensorflow_b200 as D
hp = D.hparams
hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=32)); hp.digest()
B = 32
wav = torch.from_numpy(bench.synth_mixtures(B, 32000, 1)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
m = D.Model('g', 'cuda:0', seed=1337).build()
ref = None
for groups in (4, 5, 6, 8):
    for _ in range(3):
        out = m.separate_graphed(wav, groups=groups)
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    ts = []
    for _ in range(30):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = m.separate_graphed(wav, groups=groups); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts))
    print('%d groups: %.3f ms per step (%6.0f mixtures/s)   max abs diff vs 4 groups %.3g' % (groups, t, B / t * 1e3, float((out - ref).abs().max())))
