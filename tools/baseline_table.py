"""Rows of BASELINE.md section 4 that bench.py's line does not carry: cfg 1 (toy dataset, training step), cfg 4 (3 speakers,
8 s, E = 40: training step with the anchor estimator -- k-means has no backward -- and inference with k-means), each next
to the CPU restatement on a bounded sample.  Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import danet_tensorflow_b200 as D
import bench
from oracle import danet_oracle as O      # CPU baseline / checker only
K = D.kernels
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
threads = os.cpu_count() or 1
torch.set_num_threads(threads)


def configure(**kw):
    D.hparams.__dict__.clear()
    D.hparams.__dict__.update(D.Hyperparameter().__dict__)
    D.hparams.load(kw)
    D.hparams.digest()


def time_steps(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def cpu_train(src_np, est, sep, embed, reps=2):
    P = O.reference_init(1337, embed=embed, estimators=('train_estimator',) if est == 'anchor' else ('infer_estimator',),
                         dtype=torch.float32)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    src = torch.from_numpy(src_np).to(torch.complex64)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = O.model_forward(src, Pg, encoder='bilstm-orig', train_est=est, infer_est='anchor', sep=sep, embed=embed)
        out['train_loss'].backward()
        for p in Pg.values():
            p.grad = None
        best = min(best, time.perf_counter() - t0)
    return src.shape[0] / best, best


res = {}
# ---- cfg 1: the reference's own CPU-runnable case (toy white-noise dataset, B = 2, truth estimator, default separator)
configure(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='truth', INFER_ESTIMATOR_METHOD='anchor',
          SEPARATOR_TYPE='dot-sigmoid-orig', BATCH_SIZE=2)
src_np = np.random.RandomState(1337).rand(4, 128, 129).astype(np.float32).astype(np.complex64).reshape(2, 2, 128, 129)
m = D.Model('cfg1', dev, seed=1337).build()
src = torch.from_numpy(src_np).to(dev)
ms = time_steps(lambda: m.train_step(src), 20)
rate, secs = cpu_train(src_np, 'truth', 'dot-sigmoid-orig', 20)
res['cfg1_train'] = dict(ms_per_step=ms, mixtures_per_s=2 / ms * 1e3, cpu_mixtures_per_s=rate, cpu_cores=threads,
                         cpu_sample='the whole batch of 2, best of 2 (%.2f s)' % secs)
# ---- cfg 4: 3 speakers, 8 s, E = 40
B4, n4, C4, E4 = 16, 64000, 3, 40
configure(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
          SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=B4, MAX_N_SIGNAL=C4, EMBED_SIZE=E4)
m4 = D.Model('cfg4', dev, seed=1337).build()
srcw = bench.synth_sources(B4, n4, 4242, C4)
src4 = K.stft(torch.from_numpy(srcw).to(dev))
ms = time_steps(lambda: m4.train_step(src4), 5, warm=2)
nb = 1
src4_np = np.stack([[O.stft(w) for w in u] for u in srcw[:nb]]).astype(np.complex64)
rate, secs = cpu_train(src4_np, 'anchor', 'dot-softmax-orig', E4, reps=1)
res['cfg4_train_anchor'] = dict(ms_per_step=ms, mixtures_per_s=B4 / ms * 1e3, cpu_mixtures_per_s=rate, cpu_cores=threads,
                                cpu_sample='%d of the %d mixtures, one step (%.1f s)' % (nb, B4, secs))
print(json.dumps(res))
