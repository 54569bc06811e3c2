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
for kv in os.environ.get('DANET_AB', '').split(','):          # e.g. DANET_AB=LSTM_HEADSTART_US=4,TIME_MAJOR_HANDOVER=0
    if kv:
        k, v = kv.split('=')
        setattr(D.Model, k, type(getattr(D.Model, k))(int(v)))
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
PROF = bool(os.environ.get('DANET_LSTM_PROFILE'))
if PROF:
    K._prof_ws = []
graph = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    with torch.cuda.graph(graph, stream=side):
        y = model.separate(wav, out=out_host)
K._timeline = None
prof_ws, K._prof_ws = K._prof_ws, None
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

if PROF:
    # in-kernel stamps of CTA (0,0,0) of every pipelined recurrence (launch order = group-major, 4 layers each):
    # %globaltimer at kernel entry / exit, SM clock at entry / prologue end / step s
    print('recurrent kernels, CTA (0,0,0): entry and exit by %globaltimer (us on the timeline above), clock stamps')
    for i, ws in enumerate(prof_ws):
        p = ws[:501 * 16 * 8].view(torch.int64).view(-1, 16).cpu().numpy()
        ent, ext = (p[0, 14] - t0) / 1e3, (p[0, 15] - t0) / 1e3
        ck = lambda a: a / 1840.
        e = p[0, 10]
        if os.environ['DANET_LSTM_PROFILE'] == '2':       # entry / exit only: the undisturbed schedule
            print('  #%02d (group %d, layer %d) entry %8.1f  exit %8.1f  = %6.1f us in the kernel, prologue %4.1f us'
                  % (i, i // 4, i % 4, ent, ext, ext - ent, ck(p[0, 11] - e)))
            continue
        if i in (1, 6, 11):        # the phases of a step INSIDE the full schedule (cycles, steps 100-400), as tools/lstm_profile.py prints them alone
            q = p[100:400]
            names = ['mma:wait_begin', 'mma:h_full', 'mma:issued', 'epi:step_begin', 'epi:acc_full', 'epi:tmem_ld', 'epi:gathered',
                     'epi:activated', 'epi:bar', 'copies_issued']
            print('     period %.1f cycles; relative to epi:step_begin: ' % np.diff(p[100:400, 3]).mean() +
                  ', '.join('%s %+.0f' % (n, (q[:, j] - q[:, 3]).mean()) for j, n in enumerate(names)))
            print('     copies_issued -> next h_full %.0f, h_full -> issued %.0f, issued -> acc_full %.0f' % (
                (p[101:401, 1] - p[100:400, 9]).mean(), (q[:, 2] - q[:, 1]).mean(), (q[:, 4] - q[:, 2]).mean()))
        print('  #%02d entry %8.1f exit %8.1f | prologue %5.1f us, step 1 done +%5.1f, step 16 +%5.1f, step 32 +%5.1f, step 64 +%5.1f, step 250 +%6.1f, loop end +%6.1f'
              % (i, ent, ext, ck(p[0, 11] - e), ck(p[1, 7] - e), ck(p[16, 7] - e), ck(p[32, 7] - e), ck(p[64, 7] - e),
                 ck(p[250, 7] - e), ck(p[0, 12] - e)))
