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
    """mixtures/sec of the CPU restatement (oracle) on `wav` [b, N], and its separated waveforms"""
    from oracle import danet_oracle as O
    torch.set_num_threads(threads)
    P = reference_params()
    O.separate_waveforms(wav[:1, :2048], P, dtype=torch.float32)      # warm the thread pool
    best = float('inf')
    out = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = O.separate_waveforms(wav, P, dtype=torch.float32)[0]
        best = min(best, time.perf_counter() - t0)
    return wav.shape[0] / best, best, np.asarray(out)


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


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    b = args.ref_batch
    wav = synth_mixtures(b, N_SAMPLES, 1337)
    times = []
    P = reference_params()
    from oracle import danet_oracle as O
    torch.set_num_threads(threads)
    O.separate_waveforms(wav[:1, :2048], P, dtype=torch.float32)
    steps = max(1, min(args.steps, args.ref_steps))
    for _ in range(steps):
        t0 = time.perf_counter()
        O.separate_waveforms(wav, P, dtype=torch.float32)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    val = b / (ms / 1e3)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'mixtures/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': 1, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(b, 'cpu'),
        'cpu_baseline': {'value': val, 'unit': 'mixtures/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d mixtures of 4 s per step, %d steps; torch-CPU fp32 restatement of the TF1 '
                                   'graph (TensorFlow 1.x is not installable here)' % (b, steps)},
        'e2e': {'value': val, 'unit': 'mixtures/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(batch, where, recurrent_fp16=False):
    return {'workload': 'cfg2-infer: wav->STFT->logmag->BiLSTM 4x(300+300)->anchor(6)->softmax mask x mix->iSTFT',
            'batch_per_gpu': batch, 'n_speakers': N_SPK, 'samples': N_SAMPLES, 'frames': 501, 'fft': 256,
            'hop': 64, 'embed': EMBED, 'estimator': 'anchor', 'separator': 'dot-softmax-orig',
            'parallelism': 'utterance-sharded, no collective', 'l2': 'flushed between timed steps', 'where': where,
            'schedule': 'CUDA graph, 4 stream groups of 8 utterances, staggered',
            'arithmetic': 'fp32 in / out; tensor-core products on bf16 hi/lo splits (3 products, fp32 accumulate)' +
                          ('; the recurrent product takes h_{t-1} as one fp16 value x fp16 hi/lo weights'
                           if recurrent_fp16 else '')}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=32, help='mixtures per GPU per step')
    ap.add_argument('--backend', type=int, default=None, help='0 = fp32 SIMT, 1 = tcgen05')
    ap.add_argument('--ref-batch', type=int, default=32)
    ap.add_argument('--ref-steps', type=int, default=5)
    ap.add_argument('--cpu-baseline-mixtures', type=int, default=32)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--graph', type=int, default=1, help='1 = replay the step from a CUDA graph (default), 0 = eager')
    ap.add_argument('--train-steps', type=int, default=3, help='timed training steps for the extra "train" key (0 = skip)')
    ap.add_argument('--extras', type=int, default=1,
                    help='1 = also time BASELINE.json configs[3] (3 speakers, 8 s, E = 40, k-means) and configs[4] (one 30 s '
                         'stream, latency) on rank 0 and report them under "other_configs"')
    ap.add_argument('--recurrent-fp16', type=int, default=1,
                    help='1 (default) = inference carries h into the recurrent product as fp16 (C-ABI backend 2, ~1e-4 of the '
                         'embedding scale); 0 = bf16 hi/lo everywhere (~1e-5)')
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

    # instrument our kernels with events on the launching stream (used in the eager pass only)
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
    for nm in ('lstm_seq', 'stft', 'istft', 'attractor_anchor', 'mask_cmul', 'mask_cmul_istft', 'gemm_split', 'split_operand', 'mean'):
        instrument(nm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn, steps):
        evs = []
        barrier()
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
    ms_e2e = timed_loop(step_e2e, args.steps)
    # the same K steps once more, eagerly on one stream, with CUDA events around every launch of the dominant
    # kernel (events cannot be read back from inside a replayed graph) and our launch counter running
    K.launches = 0
    Timed.on = True
    ms_eager = timed_loop(lambda: model.separate(wav_dev, groups=1), args.steps)
    Timed.on = False
    launches = K.launches
    per_kernel = {k: [a.elapsed_time(b) for a, b in v] for k, v in kernel_events.items()}
    lstm_ms = per_kernel.get('lstm_seq', [])
    clocks = sampler.stop()

    # ---- extra: the training step of the same config (forward + backward + gradient all-reduce + clip/Adam)
    train = None
    if args.train_steps > 0:
        src = K.stft(torch.from_numpy(synth_sources(B, N_SAMPLES, 1337 + rank)).to(dev))     # [B,C,T,F] complex
        for _ in range(2):
            model.train_step(src)
        K.launches = 0
        ms_train = timed_loop(lambda: model.train_step(src), args.train_steps)
        train = {'value': B * world * args.train_steps / (ms_train / 1e3), 'unit': 'mixtures/s',
                 'ms_per_step': ms_train / args.train_steps, 'steps': args.train_steps,
                 'gpu_launches': K.launches,
                 'what': 'spectra resident in HBM -> forward, PIT-MSE, backward (tcgen05 cluster BPTT + tcgen05 dW/dX products), '
                         'one NCCL all-reduce of the flat gradient buffer (N > 1), fused clip + Adam'}

    # ---- extra: the other single-GPU configurations of BASELINE.json (parity cases in tests/, timed here for the record)
    other = None
    if args.extras and rank == 0:
        other = []
        for name, b2, n2, c2, e2_, est in (('configs[3]: 3 spk, 8 s, E = 40, k-means (5 iterations)', 16, 64000, 3, 40, 'kmeans'),
                                           ('configs[4]: one 30 s stream, anchor estimator', 1, 240000, 2, 20, 'anchor')):
            hp.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD=est, INFER_ESTIMATOR_METHOD=est,
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
                # dram__bytes_read.sum + dram__bytes_write.sum of one 8-utterance launch (profiles/r01_ncu_v7_lstm.txt:
                # 41.40 + 1.72 MB; the input projections stream in once, outputs stay in L2), scaled to this launch's batch
                'traffic': 43.12e6 * B / 8.,
                'traffic_source': 'ncu --set full, lstm_tc2_kernel<1>, 8 utterances per launch, scaled by B / 8',
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the step)'
                if peaks else 'fallback',
                'ms_per_launch': lstm_avg_ms, 'launches_per_step': N_LAYERS,
                'share_of_step': (sum(lstm_ms) / ms_eager) if lstm_ms else None,
                'timed_in': 'eager single-stream pass of the same %d steps (%.3f ms/step); the headline value '
                            'replays the step from a CUDA graph with 4 staggered stream groups' % (args.steps, ms_eager / args.steps),
                'note': 'algorithmic fp32 flops 2*n_dir*B*H*4H*T; the kernel is bound by the latency of T dependent steps '
                        '(per step: ~540 cycles of MMAs streaming Wh out of tensor memory, ~550 of epilogue, ~620 of DSMEM '
                        'exchange of h; profiles/r01_lstm_phase_cycles_v4_gen2.txt), not by tensor throughput'}

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
            ('gemm_split', 'tensor', None, 'hoisted input projections (N = 2400) and the output projection (N = 2580), bf16x3')):
        ms = avg_ms(name)
        if ms is None:
            continue
        if bound == 'hbm':
            ach = work / (ms * 1e-3) / 1e9
            kernels.append({'kernel': name, 'bound': 'hbm', 'ms_per_launch': ms, 'achieved': ach, 'peak': hbm, 'unit': 'GB/s',
                            'frac': ach / hbm, 'note': note})
        else:
            n_calls = len(per_kernel[name]) // args.steps
            flops = 2. * B * T * (129 * 2400 + 3 * 600 * 2400 + 600 * 2580)          # logical fp32 flops per step
            tot_ms = ms * n_calls
            ach = flops / (tot_ms * 1e-3) / 1e12
            kernels.append({'kernel': name, 'bound': 'tensor', 'ms_per_step': tot_ms, 'launches_per_step': n_calls,
                            'achieved': ach, 'issued': 3 * ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
                            'frac_issued': 3 * ach / peak_tf,
                            'note': note + '; "issued" counts the 3 bf16 products behind every fp32-grade product'})

    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        nb = args.cpu_baseline_mixtures
        rate, secs, ref_out = cpu_reference_rate(wav_np[:nb], threads, repeats=3)
        cpu_baseline = {'value': rate, 'unit': 'mixtures/s', 'cores': threads, 'kind': 'port',
                        'sample': '%d of the %d mixtures of one step, best of 3 (%.1f s each); torch-CPU fp32 restatement of '
                                  'the TF1 graph' % (nb, B, secs)}
        # the checker at work on the timed configuration: separated waveforms of the LAST timed e2e step vs the oracle
        got = out_host[:nb].numpy()
        parity = {'max_rel_err': float(np.abs(got - ref_out).max() / np.abs(ref_out).max()), 'tolerance': 1e-3,
                  'what': 'separated waveforms of the e2e step vs the CPU oracle (fp32) on the same %d mixtures, '
                          'max-norm relative' % nb}

    line = {
        'metric': METRIC, 'value': value, 'unit': 'mixtures/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': ('f32 (tcgen05 products on bf16 hi/lo splits, fp32 accumulate' +
                  ('; recurrent state enters its product as fp16)' if args.recurrent_fp16 else ')'))
        if K.DEFAULT_BACKEND == 1 else 'f32',
        'data': 'synthetic', 'config': workload_config(B, 'cuda', bool(args.recurrent_fp16) and K.DEFAULT_BACKEND == 1),
        'e2e': {'value': e2e, 'unit': 'mixtures/s', 'h2d_bytes_per_step': int(wav_host.numel() * 4),
                'd2h_bytes_per_step': int(out_host.numel() * 4), 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu_baseline, 'clocks': clocks,
        'backend': K.DEFAULT_BACKEND, 'cuda_graph': bool(args.graph), 'train': train, 'kernels': kernels,
        'parity': parity, 'other_configs': other,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
