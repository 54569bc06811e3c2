import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import danet_tensorflow_b200 as D
hp = D.hparams
hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=32)); hp.digest()
model = D.Model('t', 'cuda:0').build()
wav = torch.from_numpy(bench.synth_mixtures(32, 32000, 1)).cuda()
for g in (1, 2, 4):
    for _ in range(3): model.separate(wav, groups=g)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): model.separate(wav, groups=g)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for _ in range(2): model.separate_graphed(wav, groups=g)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    for _ in range(5): y = model.separate_graphed(wav, groups=g)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    ref = model.separate(wav, groups=1)
    print('groups %d graphed: %.3f ms/step, max abs diff vs eager %.3g' % (g, (t4 - t3) / 5 * 1e3, float((y - ref).abs().max())))
    print('groups %d: host enqueue %.3f ms/step, total %.3f ms/step' % (g, (t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
