"""Timeline of one CUDA-graph replay of Model.separate (4 stream groups): %globaltimer stamps recorded by
1-thread kernels between the stages of every group; prints, per group, when each stage finished (us since the
first stamp) and how long it took."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import danet_tensorflow_b200 as D
K = D.kernels
hp = D.hparams
hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=32)); hp.digest()
model = D.Model('t', 'cuda:0').build()
wav = torch.from_numpy(bench.synth_mixtures(32, 32000, 1)).cuda()
for _ in range(2):
    model.separate(wav)
buf = torch.zeros(4096, dtype=torch.int64, device='cuda')
labels = []
K._timeline = (buf, labels)
graph = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    with torch.cuda.graph(graph, stream=side):
        y = model.separate(wav)
K._timeline = None
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
t = buf[:len(labels)].cpu().numpy().astype(np.int64)
t0 = t.min()
rows = {}
order = []
cur = None
for lab, ts in zip(labels, t):
    if lab.endswith('start'):
        cur = lab.split()[0]
        order.append(cur)
        rows[cur] = []
    rows[cur].append((lab, (ts - t0) / 1e3))
for g in order:
    prev = None
    print('--- group', g)
    for lab, ts in rows[g]:
        print('  %-28s at %8.1f us  (+%7.1f)' % (lab, ts, ts - prev if prev is not None else 0.))
        prev = ts
print('total %.1f us' % ((t.max() - t0) / 1e3))
