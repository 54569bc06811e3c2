"""Consecutive inference steps overlapped: step i+1 is replayed on a second stream while step i drains (its recurrences
free their SMs one group at a time, and the end of a step is a throughput-bound convoy of projections).  Compares K steps
back to back on ONE stream with the same K steps alternating over N streams (N graphs with their own buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import danet_tensorflow_b200 as D
hp = D.hparams
hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
             SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=32)); hp.digest()
B, K_STEPS = 32, 24
m = D.Model('ov', 'cuda:0', seed=1337).build()
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wavs = [torch.from_numpy(bench.synth_mixtures(B, 32000, 10 + i)).cuda() for i in range(NS)]
outs = [torch.empty((B, 2, 64 * 501), dtype=torch.float32, device='cuda') for _ in range(NS)]
hin = [torch.from_numpy(bench.synth_mixtures(B, 32000, 10 + i)).pin_memory() for i in range(NS)]
hout = [torch.empty((B, 2, 64 * 501), dtype=torch.float32).pin_memory() for _ in range(NS)]
streams = [torch.cuda.Stream() for _ in range(NS)]
for i in range(NS):
    for _ in range(2):
        m.separate_graphed(wavs[i], out=outs[i])
        m.separate_host(hin[i], hout[i])
torch.cuda.synchronize()
ref = [o.clone() for o in outs]


def timed(fn):
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K_STEPS)
    return best


def serial(host):
    for k in range(K_STEPS):
        i = k % NS
        m.separate_host(hin[i], hout[i]) if host else m.separate_graphed(wavs[i], out=outs[i])


def overlapped(host):
    main = torch.cuda.current_stream()
    for st in streams:
        st.wait_stream(main)
    for k in range(K_STEPS):
        i = k % NS
        with torch.cuda.stream(streams[i]):
            m.separate_host(hin[i], hout[i]) if host else m.separate_graphed(wavs[i], out=outs[i])
    for st in streams:
        main.wait_stream(st)


for host in (False, True):
    a, b = timed(lambda: serial(host)), timed(lambda: overlapped(host))
    print('%-16s one stream %.3f ms/step (%6.0f mixtures/s)   %d streams %.3f ms/step (%6.0f mixtures/s)'
          % ('pinned host i/o' if host else 'device resident', a, B / a * 1e3, NS, b, B / b * 1e3))
print('outputs unchanged:', all(torch.equal(o, r) for o, r in zip(outs, ref)))
