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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32000          # samples per utterance (T = N/64 + 1 frames)
HOST = len(sys.argv) > 2 and sys.argv[2] == 'host'             # pinned host buffers in and out (the e2e call)
wav = torch.from_numpy(bench.synth_mixtures(32, N, 1))
wav = wav.pin_memory() if HOST else wav.cuda()
out_host = torch.empty((32, 2, 64 * K.num_frames(N)), dtype=torch.float32).pin_memory() if HOST else None
for _ in range(2):
    model.separate(wav, out=out_host)
buf = torch.zeros(4096, dtype=torch.int64, device='cuda')
labels = []
K._timeline = (buf, labels)
graph = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    with torch.cuda.graph(graph, stream=side):
        y = model.separate(wav, out=out_host)
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
print('N = %d samples, %s: total %.1f us' % (N, 'pinned host in/out' if HOST else 'device resident', (t.max() - t0) / 1e3))
