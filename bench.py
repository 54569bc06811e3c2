#!/usr/bin/env python
"""
bench.py -- mixtures/sec of the DANet separation hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1], inference form): a batch of 32 synthetic 2-speaker mixtures,
4 s @ 8 kHz, FFT 256 / hop 64 (T = 501 frames), BiLSTM encoder 4 x (300+300), EMBED 20, anchor
estimator (6 anchors), softmax separator:  wav -> STFT -> log-magnitude -> BiLSTM -> anchor
attractors -> mask x complex mixture -> iSTFT -> 2 separated wavs per mixture.  One step = one
pass over one batch.  Per-GPU work is fixed (weak scaling); utterances shard across ranks with no
data-path collective.

  value  : device-resident inputs, CUDA-event time of K steps (max over ranks)
  e2e    : the public call Model.separate_host(): pinned host wavs -> H2D -> ... -> D2H wavs
  roofline: the dominant kernel (the persistent BiLSTM recurrence), timed live with CUDA events
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle/, fp32 torch-CPU,
           all host threads) on a bounded sample of the same workload.  TensorFlow 1.x cannot be
           installed in this image, so the restatement stands in for the TF1 reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = 'mixtures/sec (2-spk, 4s@8kHz, FFT256)'
N_SAMPLES = 32000
N_SPK = 2
EMBED = 20
HDIM = 300
N_LAYERS = 4


def synth_mixtures(batch, n_samples, seed):
    """per source white noise through a one-pole low-pass (a = 0.9) x a 4 Hz
    raised-cosine envelope with random phase, RMS 1000 (int16 scale); mixture = sum of sources."""
    return synth_sources(batch, n_samples, seed).sum(1).astype(np.float32)


def synth_sources(batch, n_samples, seed, n_spk=None):
    """the per-source waveforms [B, C, N] behind synth_mixtures (training consumes the sources)"""
    n_spk = n_spk or N_SPK
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((batch, n_spk, n_samples)).astype(np.float64)
    y = np.empty_like(x)
    acc = np.zeros((batch, n_spk))
    for i in range(n_samples):
        acc = 0.9 * acc + x[..., i]
        y[..., i] = acc
    t = np.arange(n_samples) / 8000.
    ph = rs.uniform(0, 2 * np.pi, (batch, n_spk, 1))
    y *= 0.5 - 0.5 * np.cos(2 * np.pi * 4. * t + ph)
    y *= 1000. / np.sqrt((y ** 2).mean(-1, keepdims=True))
    return y.astype(np.float32)


def reference_params(seed=1337):
    from oracle import danet_oracle as O
    return O.reference_init(seed, encoder='bilstm-orig', embed=EMBED, estimators=('infer_estimator',),
                            dtype=torch.float32)


def cpu_reference_rate(wav, threads, repeats=1):
    """mixtures/sec of the CPU restatement (oracle) on `wav` [b, N], its separated waveforms and intermediate tensors"""
    from oracle import danet_oracle as O
    torch.set_num_threads(threads)
    P = reference_params()
    O.separate_waveforms(wav[:1, :2048], P, dtype=torch.float32)      # warm the thread pool
    best = float('inf')
    out = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out, sig, aux = O.separate_waveforms(wav, P, dtype=torch.float32)
        best = min(best, time.perf_counter() - t0)
    aux = dict(aux, sig=sig)
    return wav.shape[0] / best, best, np.asarray(out), aux


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled through NVML every 20 ms during the timed region
    (same quantities as the nvidia-smi line of B200_PROFILING.md, without a process spawn per sample)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        nv = self.nv
        if nv is None:
            return
        while not self._halt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, 'nvmlDeviceGetCurrentClocksEventReasons') else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(sm), int(rs)))
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        nv = self.nv
        if nv is None or not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable'], 'samples': 0}
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
        return {'sm_mhz': float(np.median([c for c, _ in self.rows])), 'sm_max_mhz': self.max_sm,
                'reasons': reasons, 'samples': len(self.rows)}


def cpu_reference_train_rate(src_np, threads, repeats=2):
    """mixtures/sec of ONE training step of the CPU restatement on complex source spectra [b,C,T,F]: forward of the train
    graph (main.py:233-337), torch autograd in place of tf.gradients (main.py:357-358), clip + Adam (main.py:359-363)"""
    from oracle import danet_oracle as O
    torch.set_num_threads(threads)
    P = O.reference_init(1337, encoder='bilstm-orig', embed=EMBED, estimators=('train_estimator',), dtype=torch.float32)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    m = {k: torch.zeros_like(v) for k, v in P.items()}
    v2 = {k: torch.zeros_like(v) for k, v in P.items()}
    src = torch.from_numpy(src_np).to(torch.complex64)
    best = float('inf')
    for it in range(repeats):
        t0 = time.perf_counter()
        out = O.model_forward(src, Pg, encoder='bilstm-orig', train_est='anchor', infer_est='anchor',
                              sep='dot-softmax-orig', embed=EMBED)
        out['train_loss'].backward()
        with torch.no_grad():
            cur = {k: p.detach().clone() for k, p in Pg.items()}
            O.clip_adam_step(cur, {k: p.grad for k, p in Pg.items()}, m, v2, it + 1)
            for k, p in Pg.items():
                p.copy_(cur[k])
                p.grad = None
        best = min(best, time.perf_counter() - t0)
    return src.shape[0] / best, best


def source_spectra(batch, seed):
    """complex source spectra [b,C,T,F] of the synthetic sources (host STFT of the oracle: the training step's input)"""
    from oracle import danet_oracle as O
    w = synth_sources(batch, N_SAMPLES, seed)
    return np.stack([[O.stft(x) for x in u] for u in w]).astype(np.complex64)


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores.  TensorFlow 1.x cannot be
    installed in this image (no wheel, no network), so the op-for-op torch-CPU restatement in oracle/
    stands in for it (`cpu_baseline.kind = "port"`).  Same config / metric / unit as the repo arm, --steps / --warmup
    honoured; every step is one full batch of the workload."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    b = args.batch
    wav = synth_mixtures(b, N_SAMPLES, 1337)
    times = []
    P = reference_params()
    from oracle import danet_oracle as O
    torch.set_num_threads(threads)
    O.separate_waveforms(wav[:1, :2048], P, dtype=torch.float32)
    for _ in range(max(args.warmup, 0)):
        O.separate_waveforms(wav, P, dtype=torch.float32)
    for _ in range(args.steps):
        t0 = time.perf_counter()
        O.separate_waveforms(wav, P, dtype=torch.float32)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    val = b / (ms / 1e3)
    train = None
    if args.train_steps > 0:
        nb = args.cpu_train_mixtures
        rate, secs = cpu_reference_train_rate(source_spectra(nb, 1337), threads)
        train = {'value': rate, 'unit': 'mixtures/s', 'ms_per_step': 1e3 * secs * b / nb, 'cores': threads, 'kind': 'port',
                 'sample': '%d of the %d mixtures of one training step, best of 2 (%.1f s each): oracle forward + torch '
                           'autograd + clip/Adam' % (nb, b, secs)}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'mixtures/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(b),
        'where': 'host cores (torch-CPU fp32, %d threads)' % threads,
        'cpu_baseline': {'value': val, 'unit': 'mixtures/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d mixtures of 4 s per step (one full batch), %d steps after %d warm-up; torch-CPU fp32 '
                                   'restatement of the TF1 graph (TensorFlow 1.x is not installable here)'
                                   % (b, args.steps, args.warmup)},
        'e2e': {'value': val, 'unit': 'mixtures/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'train': train,
    }
    print(json.dumps(line), flush=True)


def workload_config(batch):
    """ONE dict for both arms (the driver compares them key by key); what differs between the arms -- where it runs, the
    arithmetic, the schedule -- is reported under separate top-level keys"""
    return {'workload': 'cfg2-infer: wav->STFT->logmag->BiLSTM 4x(300+300)->anchor(6)->softmax mask x mix->iSTFT',
            'batch_per_gpu': batch, 'n_speakers': N_SPK, 'samples': N_SAMPLES, 'frames': 501, 'fft': 256,
            'hop': 64, 'embed': EMBED, 'estimator': 'anchor', 'separator': 'dot-softmax-orig',
            'parallelism': 'utterance-sharded, no collective',
            'l2': 'flushed between timed steps (GPU arm: 256 MB write before every step)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=32, help='mixtures per GPU per step')
    ap.add_argument('--backend', type=int, default=None, help='0 = fp32 SIMT, 1 = tcgen05')
    ap.add_argument('--cpu-baseline-mixtures', type=int, default=32)
    ap.add_argument('--cpu-train-mixtures', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--graph', type=int, default=1, help='1 = replay the step from a CUDA graph (default), 0 = eager')
    ap.add_argument('--train-steps', type=int, default=10, help='timed training steps for the "train" key (0 = skip)')
    ap.add_argument('--extras', type=int, default=1,
                    help='1 = also time BASELINE.json configs[3] (3 speakers, 8 s, E = 40, k-means) and configs[4] (one 30 s '
                         'stream, latency) on rank 0 and report them under "other_configs"')
    ap.add_argument('--recurrent-fp16', type=int, default=1,
                    help='1 (default) = inference carries h into the recurrent product as fp16 (C-ABI backend 2, ~1e-4 of the '
                         'embedding scale); 0 = bf16 hi/lo everywhere (~1e-5).  The other setting is timed as well and '
                         'reported under "precision_ab"')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import danet_tensorflow_b200 as D
    K = D.kernels
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (no CPU fallback for the product path)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    full_affinity = os.sched_getaffinity(0)
    numa = D.shard.bind_to_local_numa(local)          # before any pinned allocation: host buffers on the GPU's own node
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    if args.backend is not None:
        K.DEFAULT_BACKEND = args.backend
    D.Model.RECURRENT_FP16 = bool(args.recurrent_fp16)
    D.build.build()
    D._lib.check(D._lib.load().danet_check_device(), 'check_device')

    hp = D.hparams
    hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
                 SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=args.batch, EMBED_SIZE=EMBED, MAX_N_SIGNAL=N_SPK))
    hp.digest()
    model = D.Model('bench', dev, seed=1337).build()
    B = args.batch
    wav_np = synth_mixtures(B, N_SAMPLES, 1337 + rank)
    wav_host = torch.from_numpy(wav_np).pin_memory()
    wav_dev = wav_host.to(dev)
    T = K.num_frames(N_SAMPLES)
    out_host = torch.empty((B, N_SPK, 64 * T), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    # instrument our kernels with events on the launching stream (used in the eager passes only)
    kernel_events = {}

    class Timed(object):
        on = False

    def instrument(name):
        raw = getattr(K, name)

        def timed(*a, **kw):
            if not Timed.on:
                return raw(*a, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = raw(*a, **kw)
            e1.record()
            kernel_events.setdefault(name, []).append((e0, e1))
            return r
        setattr(K, name, timed)
    for nm in ('lstm_seq', 'lstm_seq_bwd', 'stft', 'istft', 'attractor_anchor', 'mask_cmul', 'mask_cmul_istft', 'gemm_split',
               'proj_anchor', 'split_operand', 'mean'):
        instrument(nm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sm_khz = torch.cuda.get_device_properties(dev).clock_rate if hasattr(torch.cuda.get_device_properties(dev), 'clock_rate') else 1965000

    def timed_loop(step_fn, steps, phase_ms=0.):
        """K steps, CUDA events around each, max over ranks of the sum.  phase_ms > 0: this rank's first step starts that
        much after the barrier (a device-side wait ahead of the first event), so that the ranks' steps -- equally long and
        otherwise in lockstep -- end at different moments and their device-to-host bursts do not meet on the host side"""
        evs = []
        barrier()
        if phase_ms > 0.:
            torch.cuda._sleep(int(phase_ms * sm_khz))
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        return D.shard.max_over_ranks(ms, dev)

    def step_dev():
        return model.separate_graphed(wav_dev) if args.graph else model.separate(wav_dev)

    def step_e2e():
        return model.separate_host(wav_host, out_host, graphed=bool(args.graph))

    for _ in range(max(args.warmup, 3)):
        step_dev()
        step_e2e()
        model.separate(wav_dev, groups=1)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev = timed_loop(step_dev, args.steps)
    ms_e2e_lockstep = timed_loop(step_e2e, args.steps)
    if world > 1:
        # N ranks of one box share the host side of their PCIe links: measured (tools/copy_contention.py) a rank's 8.2 MB of
        # results take 152 us alone and 245 us when a second rank copies at the same moment.  Equal steps keep the ranks in
        # lockstep, so every step ends in N simultaneous bursts; shifting rank r by r/N of a step interleaves them.
        ms_e2e = timed_loop(step_e2e, args.steps, phase_ms=rank * (ms_dev / args.steps) / world)
    else:
        ms_e2e = ms_e2e_lockstep
    # the same K steps once more, eagerly on one stream, with CUDA events around every launch of the dominant
    # kernel (events cannot be read back from inside a replayed graph) and our launch counter running
    K.launches = 0
    Timed.on = True
    ms_eager = timed_loop(lambda: model.separate(wav_dev, groups=1), args.steps)
    Timed.on = False
    per_kernel = {k: [a.elapsed_time(b) for a, b in v] for k, v in kernel_events.items()}
    # our launches in one step of the TIMED schedule (4 stream groups; the graph replays exactly these): counted on one
    # eager call of the same grouped step
    K.launches = 0
    model.separate(wav_dev)
    torch.cuda.synchronize()
    launches_per_step = K.launches
    launches = launches_per_step * args.steps
    lstm_ms = per_kernel.get('lstm_seq', [])
    clocks = sampler.stop()
    got_e2e = out_host.clone()

    # ---- the other precision setting of the recurrent product, timed the same way (device-resident and end to end)
    precision_ab = None
    if K.DEFAULT_BACKEND == 1 and args.graph:
        other = not D.Model.RECURRENT_FP16
        D.Model.RECURRENT_FP16 = other
        model._invalidate()
        for _ in range(3):
            step_dev()
            step_e2e()
        torch.cuda.synchronize()
        ab_dev = timed_loop(step_dev, args.steps)
        ab_e2e = timed_loop(step_e2e, args.steps)
        precision_ab = {'recurrent_fp16': int(other), 'value': B * world * args.steps / (ab_dev / 1e3),
                        'e2e': B * world * args.steps / (ab_e2e / 1e3), 'unit': 'mixtures/s',
                        'ms_per_step': ab_dev / args.steps,
                        'what': 'the same step with the recurrent product ' +
                                ('taking h as one fp16 value' if other else 'on bf16 hi/lo pairs only (bf16x3 everywhere)')}
        D.Model.RECURRENT_FP16 = not other
        model._invalidate()

    # ---- parity of every tensor north_star names, on the timed batch: stage-by-stage eager pass of the product path
    def product_stages(n):
        mix, logmag = K.stft(wav_dev[:n], want_logmag=True)
        embed = model.encoder(logmag)
        flat = embed.view(n, T * 129, EMBED)
        attrs = model.infer_estimator(embed, s_embed_flat=flat)
        o = model.separator(None, attrs, flat, s_mixed_signals=mix, want=('sep', 'masks'))
        return embed.cpu().numpy(), o['masks'].cpu().numpy(), torch.view_as_real(o['sep']).cpu().numpy()

    stages = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        stages = product_stages(min(args.cpu_baseline_mixtures, B, 8))     # before the training steps move the weights

    # ---- the training step of the same config (forward + backward + bucketed gradient all-reduce + clip/Adam)
    train = None
    if args.train_steps > 0:
        src = K.stft(torch.from_numpy(synth_sources(B, N_SAMPLES, 1337 + rank)).to(dev))     # [B,C,T,F] complex
        for _ in range(4):                       # two eager steps, the capture of forward + backward, one replay
            model.train_step(src)
        model.TIME_ALLREDUCE = True
        model._ar_events = []
        ms_train = timed_loop(lambda: model.train_step(src), args.train_steps)
        model.TIME_ALLREDUCE = False
        ar_ms = [a.elapsed_time(b) for a, b in model._ar_events]
        ar_exposed = D.shard.max_over_ranks(float(np.mean(ar_ms)) if ar_ms else 0., dev)
        # the timed steps replay forward + backward from a CUDA graph (two stream groups): events cannot be read back from
        # inside a replay and our launch counter does not run, so (a) the launches of the TIMED schedule are counted on one
        # eager step of the same grouped schedule, (b) the BPTT / forward recurrence kernels are timed on eager one-pass steps
        groups_timed = model._train_group_count(B)
        model.TRAIN_GRAPH = False
        K.launches = 0
        model.train_step(src)
        torch.cuda.synchronize()
        train_launches = K.launches
        model.TRAIN_GROUPS = 1
        model.train_step(src)                     # the one-pass shapes are new to the caching allocator: not under the events
        torch.cuda.synchronize()
        kernel_events.clear()
        Timed.on = True
        for _ in range(3):
            model.train_step(src)
        torch.cuda.synchronize()
        Timed.on = False
        del model.TRAIN_GROUPS, model.TRAIN_GRAPH          # back to the class defaults
        bptt = [a.elapsed_time(b) for a, b in kernel_events.get('lstm_seq_bwd', [])]
        fwd_l = [a.elapsed_time(b) for a, b in kernel_events.get('lstm_seq', [])]
        n_param = int(model._flat['grad'].numel())
        train = {'value': B * world * args.train_steps / (ms_train / 1e3), 'unit': 'mixtures/s',
                 'ms_per_step': ms_train / args.train_steps, 'steps': args.train_steps, 'n_gpus': world,
                 'global_batch': B * world, 'gpu_launches': train_launches, 'stream_groups': groups_timed,
                 'allreduce_ms_exposed': ar_exposed, 'allreduce_bytes': 4 * n_param,
                 'allreduce': ('NCCL, one exchange of the flat gradient buffer after the stream groups\' gradients are summed; '
                               'exposed = its whole duration' if groups_timed > 1 else
                               'NCCL, 5 buckets in backward completion order (projection + anchors, L3 .. L0), each started '
                               'on the stream that produced it; exposed = what the step still waits for after the backward'),
                 'what': 'spectra resident in HBM -> forward, PIT-MSE, backward (tcgen05 cluster BPTT + tcgen05 dW/dX '
                         'products) in %d stream groups replayed from a CUDA graph, gradient all-reduce (N > 1), fused '
                         'clip + Adam' % groups_timed}
        if bptt:
            bflops = 2. * 2 * B * HDIM * 4 * HDIM * T           # dh = da * Wh^T, both directions, per launch
            bms = float(np.mean(bptt))
            train['roofline'] = {'bound': 'tensor', 'kernel': 'lstm_seq_bwd (BPTT, one launch per layer)',
                                 'achieved': bflops / (bms * 1e-3) / 1e12, 'unit': 'TFLOP/s', 'ms_per_launch': bms,
                                 'launches_per_step': N_LAYERS, 'forward_recurrence_ms_per_launch': float(np.mean(fwd_l)) if fwd_l else None}

    # ---- extra: the other single-GPU configurations of BASELINE.json (parity cases in tests/, timed here for the record)
    other = None
    if args.extras and rank == 0:
        other = []
        for name, b2, n2, c2, e2_, est, enc in (
                ('configs[3]: 3 spk, 8 s, E = 40, k-means (5 iterations)', 16, 64000, 3, 40, 'kmeans', 'bilstm-orig'),
                ('configs[4]: one 30 s stream, anchor estimator', 1, 240000, 2, 20, 'anchor', 'bilstm-orig'),
                ('configs[1] shape with the lstm-orig encoder (4 x 600 unidirectional: wide tcgen05 recurrent kernel)',
                 args.batch, N_SAMPLES, N_SPK, EMBED, 'anchor', 'lstm-orig')):
            hp.load(dict(ENCODER_TYPE=enc, TRAIN_ESTIMATOR_METHOD=est, INFER_ESTIMATOR_METHOD=est,
                         SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=b2, EMBED_SIZE=e2_, MAX_N_SIGNAL=c2))
            hp.digest()
            m2 = D.Model('other', dev, seed=1337).build()
            w2 = torch.from_numpy(synth_sources(b2, n2, 4242, c2).sum(1).astype(np.float32)).pin_memory()
            o2 = torch.empty((b2, c2, 64 * K.num_frames(n2)), dtype=torch.float32).pin_memory()
            for _ in range(3):
                m2.separate_host(w2, o2)
            torch.cuda.synchronize()
            lat = []
            for _ in range(7):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                m2.separate_host(w2, o2)
                e1.record()
                torch.cuda.synchronize()
                lat.append(e0.elapsed_time(e1))
            med = float(np.median(lat))
            other.append({'config': name, 'batch': b2, 'samples': n2, 'frames': K.num_frames(n2), 'ms_per_step_e2e': med,
                          'mixtures_per_s_e2e': b2 / med * 1e3, 'audio_seconds_per_second': b2 * n2 / 8000. / (med * 1e-3),
                          'what': 'pinned host waveforms in -> separated waveforms in pinned host memory, CUDA graph replay, '
                                  'median of 7, L2 flushed'})
            del m2
        hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
                     SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=args.batch, EMBED_SIZE=EMBED, MAX_N_SIGNAL=N_SPK))
        hp.digest()

    total = B * world
    value = total * args.steps / (ms_dev / 1e3)
    e2e = total * args.steps / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    # algorithmic flops of one recurrent launch (one BiLSTM layer): h_{t-1} * Wh for both directions
    flops = 2. * 2 * B * HDIM * 4 * HDIM * T
    lstm_avg_ms = float(np.mean(lstm_ms)) if lstm_ms else None
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.)
    achieved = flops / (lstm_avg_ms * 1e-3) / 1e12 if lstm_avg_ms else None
    roofline = {'bound': 'tensor', 'kernel': 'lstm_seq (BiLSTM recurrence, one launch per layer)',
                'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': achieved / peak_tf if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of one 8-utterance launch (profiles/r02_ncu_final_lstm.txt:
                # 41.85 + 1.04 MB; the input projections stream in once, outputs stay in L2), scaled to this launch's batch
                'traffic': 42.19e6 * B / 8.,       # profiles/r02_ncu_final_lstm.txt: 41.85 MB read + 0.33 MB written per launch
                'traffic_source': 'ncu --set full, lstm_tc2_kernel<1>, 8 utterances per launch, scaled by B / 8',
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the step)'
                if peaks else 'fallback',
                'ms_per_launch': lstm_avg_ms, 'launches_per_step': N_LAYERS,
                'share_of_step': (sum(lstm_ms) / ms_eager) if lstm_ms else None,
                'timed_in': 'eager single-stream pass of the same %d steps (%.3f ms/step); the headline value '
                            'replays the step from a CUDA graph with 4 staggered stream groups' % (args.steps, ms_eager / args.steps),
                'note': 'algorithmic fp32 flops 2*n_dir*B*H*4H*T; the kernel is bound by the latency of T dependent steps '
                        '(MMA issue, epilogue and DSMEM exchange of h in series; profiles/r02_lstm_phase_cycles*.txt), '
                        'not by tensor throughput'}
    if train is not None and 'roofline' in train:
        train['roofline'].update(peak=peak_tf, frac=train['roofline']['achieved'] / peak_tf)

    # the other kernels of the step against their own rooflines (algorithmic bytes / flops from DESIGN.md section 6)
    hbm = peaks.get('hbm_gbs', 6650.)
    TF, E = T * 129, EMBED

    def avg_ms(name):
        v = per_kernel.get(name, [])
        return float(np.mean(v)) if v else None
    kernels = []
    for name, bound, work, note in (
            ('stft', 'hbm', B * (4. * N_SAMPLES + 12. * TF), 'wav in, complex spectrum + log-magnitude out'),
            ('attractor_anchor', 'hbm', B * 4. * TF * E, 'one read of the embedding (fp32 SIMT products: 336 MAC per bin)'),
            ('mask_cmul', 'hbm', B * (4. * TF * E + 8. * TF + 8. * N_SPK * TF), 'embedding + mixture in, separated spectra out'),
            ('istft', 'hbm', B * N_SPK * (8. * TF + 4. * 64 * T), 'spectra in, waveforms out'),
            ('mask_cmul_istft', 'hbm', B * (4. * TF * E + 8. * TF + 4. * N_SPK * 64 * T),
             'K4 fused: embedding + mixture in, separated waveforms out (masked spectra stay in shared memory)'),
            ('proj_anchor', 'tensor', 2. * B * T * 600 * 2580,
             'output projection (N = 2580) with the anchor estimator\'s sums in its epilogue (replaces gemm + attractor_anchor); '
             'also streams 4*T*F*E bytes of embedding out'),
            ('gemm_split', 'tensor', None, 'hoisted input projections (N = 2400)' +
             ('' if per_kernel.get('proj_anchor') else ' and the output projection (N = 2580)') + ', bf16x3')):
        ms = avg_ms(name)
        if ms is None:
            continue
        if bound == 'hbm':
            ach = work / (ms * 1e-3) / 1e9
            kernels.append({'kernel': name, 'bound': 'hbm', 'ms_per_launch': ms, 'achieved': ach, 'peak': hbm, 'unit': 'GB/s',
                            'frac': ach / hbm, 'note': note})
        elif work is not None:
            ach = work / (ms * 1e-3) / 1e12
            kernels.append({'kernel': name, 'bound': 'tensor', 'ms_per_launch': ms, 'achieved': ach, 'issued': 3 * ach,
                            'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf, 'frac_issued': 3 * ach / peak_tf,
                            'note': note})
        else:
            n_calls = len(per_kernel[name]) // args.steps
            gflops = 2. * B * T * (129 * 2400 + 3 * 600 * 2400 + (0 if per_kernel.get('proj_anchor') else 600 * 2580))
            tot_ms = ms * n_calls
            ach = gflops / (tot_ms * 1e-3) / 1e12
            kernels.append({'kernel': name, 'bound': 'tensor', 'ms_per_step': tot_ms, 'launches_per_step': n_calls,
                            'achieved': ach, 'issued': 3 * ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
                            'frac_issued': 3 * ach / peak_tf,
                            'note': note + '; "issued" counts the 3 bf16 products behind every fp32-grade product'})

    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline and world == 1:       # "on rank 0 at N = 1 only": all host cores belong to this process
        os.sched_setaffinity(0, full_affinity)
        threads = os.cpu_count() or 1
        nb = min(args.cpu_baseline_mixtures, B)
        rate, secs, ref_out, aux = cpu_reference_rate(wav_np[:nb], threads, repeats=3)
        cpu_baseline = {'value': rate, 'unit': 'mixtures/s', 'cores': threads, 'kind': 'port',
                        'sample': '%d of the %d mixtures of one step, best of 3 (%.1f s each); torch-CPU fp32 restatement of '
                                  'the TF1 graph' % (nb, B, secs)}
        # the checker at work on the timed configuration: every tensor north_star names, product vs the fp32 CPU oracle
        # on the same mixtures (the fp64 comparison is tests/test_gpu_fullsize.py)
        npar = min(nb, 8)
        g_embed, g_masks, g_sep = stages
        r_embed, r_masks = aux['embed'][:npar].numpy(), aux['masks'][:npar].numpy()
        r_sep = torch.view_as_real(aux['sig'][:npar]).numpy()
        got = got_e2e[:nb].numpy()

        def mx(a, b, relative=True):
            return float(np.abs(a - b).max() / (np.abs(b).max() if relative else 1.))
        parity = {'tolerance': 1e-3, 'recurrent_fp16': int(D.Model.RECURRENT_FP16),
                  'embedding': mx(g_embed, r_embed), 'masks_abs': mx(g_masks, r_masks, False),
                  'spectra': mx(g_sep, r_sep), 'waveforms_e2e': mx(got, ref_out), 'max_rel_err': mx(got, ref_out),
                  'what': 'max-norm (relative to the largest reference entry; masks absolute) against the fp32 CPU oracle: '
                          'embedding / masks / separated spectra on the first %d mixtures of the timed batch, waveforms of '
                          'the LAST timed e2e step on %d' % (npar, nb)}
        if train is not None:
            nt = args.cpu_train_mixtures
            trate, tsecs = cpu_reference_train_rate(source_spectra(nt, 1337), threads)
            train['cpu_baseline'] = {'value': trate, 'unit': 'mixtures/s', 'cores': threads, 'kind': 'port',
                                     'sample': '%d of the %d mixtures of one training step, best of 2 (%.1f s each): oracle '
                                               'forward + torch autograd + clip/Adam' % (nt, B, tsecs)}

    fp16_on = bool(args.recurrent_fp16) and K.DEFAULT_BACKEND == 1
    line = {
        'metric': METRIC, 'value': value, 'unit': 'mixtures/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': ('f32 (tcgen05 products on bf16 hi/lo splits, fp32 accumulate' +
                  ('; recurrent state enters its product as fp16)' if fp16_on else ')'))
        if K.DEFAULT_BACKEND == 1 else 'f32',
        'data': 'synthetic', 'config': workload_config(B),
        'where': 'cuda (one process per GPU%s)' % (', host buffers bound to NUMA node %s' % numa if numa is not None else ''),
        'schedule': 'CUDA graph, 4 stream groups of 8 utterances, staggered',
        'arithmetic': 'fp32 in / out; tensor-core products on bf16 hi/lo splits (3 products, fp32 accumulate)' +
                      ('; the recurrent product takes h_{t-1} as one fp16 value x fp16 hi/lo weights' if fp16_on else ''),
        'backend': K.DEFAULT_BACKEND, 'cuda_graph': bool(args.graph),
        'kernels': kernels, 'other_configs': other, 'precision_ab': precision_ab,
        'clocks': clocks, 'gpu_launches': launches, 'gpu_launches_per_step': launches_per_step,
        'cpu_baseline': cpu_baseline, 'roofline': roofline, 'parity': parity,
        'e2e': {'value': e2e, 'unit': 'mixtures/s', 'h2d_bytes_per_step': int(wav_host.numel() * 4),
                'd2h_bytes_per_step': int(out_host.numel() * 4), 'ms_per_step': ms_e2e / args.steps,
                'lockstep_value': total * args.steps / (ms_e2e_lockstep / 1e3),
                'ranks': 'phase-shifted by rank/N of a step (their host-bound result copies interleave); lockstep_value = '
                         'all ranks started together' if world > 1 else 'single rank'},
        # last on purpose: log tails keep the end of the line
        'train': train,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
