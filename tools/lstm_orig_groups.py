import os, sys
sys.path.insert(0, '/root/repo')
import torch, bench
import danet_tensorflow_b200 as D
hp = D.hparams
B, N = 32, 499 * 64
wav = torch.from_numpy(bench.synth_mixtures(B, N, 1)).cuda()
hp.load(dict(ENCODER_TYPE='lstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=B)); hp.digest()
m = D.Model('l', 'cuda:0').build()
for g in (1, 2, 4):
    for _ in range(2):
        m.separate_graphed(wav, groups=g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.separate_graphed(wav, groups=g)
    e1.record(); torch.cuda.synchronize()
    print('lstm-orig groups %d: %.3f ms' % (g, e0.elapsed_time(e1) / 3))
