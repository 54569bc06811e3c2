"""
Streaming front and back end of the demo path (BASELINE.json configs[4]: one 30 s stream).

The reference's demo (main.py:660-695) loads the whole file, transforms it on the host (app/utils.py:111-122), runs the
infer graph and inverse-transforms every source in a Python loop (app/utils.py:53-75).  What CAN move ahead of the end
of the utterance is the front end: frame t of scipy.signal.stft covers samples [64 t - 128, 64 t + 128), so it is final
as soon as those samples have arrived.  `StreamingSeparator.feed()` therefore copies each chunk of audio to the device
and transforms every frame that has become final (K1, danet_stft_fwd), while the audio is still coming in; when the
stream ends only the last few frames are left.  The encoder itself cannot start early: the first thing it does is
subtract the mean over ALL frames (app/modules.py:209-210), its backward direction starts at the LAST frame
(app/modules.py:128-137), and the attractors are sums over every bin (app/modules.py:513-537).  The back end hands the
audio out in blocks: the fused mask x mixture -> iSTFT kernel (K4) writes the separated waveforms once, and they travel to
pinned host memory block by block on a copy stream, each block with its own event, so a consumer can start playing /
writing the first block while the rest is still crossing the link.

Results are bit-identical to `Model.separate` on the whole waveform: the per-frame arithmetic of K1 does not depend on
the frame's position, and blocks are cut at even frame indices so the kernel pairs the same frames into one complex FFT.
"""
import torch

from . import kernels as K
from .hparams import hparams

HOP = K.FFT_STRIDE
HALF = K.FFT_SIZE // 2


class StreamingSeparator(object):
    def __init__(self, model, max_seconds=60., out_block_frames=512):
        self.model = model
        self.dev = model.device
        cap = int(max_seconds * hparams.SMPRATE)
        self.cap_frames = K.num_frames(cap)
        self.wav = torch.zeros((1, cap + HOP), dtype=torch.float32, device=self.dev)
        self.mix = torch.empty((1, self.cap_frames, K.FEATURE), dtype=torch.complex64, device=self.dev)
        self.logmag = torch.empty((1, self.cap_frames, K.FEATURE), dtype=torch.float32, device=self.dev)
        self.out_block = int(out_block_frames) // 2 * 2
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.reset()

    def reset(self):
        self.n = 0                   # samples received
        self.t_done = 0              # frames [0, t_done) are final and transformed (always even until finish)

    def _transform(self, t_lo, t_hi, n_end):
        """frames [t_lo, t_hi) of the stream from the samples received so far (`n_end` of them are valid): K1 on the
        slice that starts two frames earlier -- its own left zero padding only reaches its first two frames, which are
        dropped -- and extends far enough that the wanted frames never see the right padding"""
        if t_hi <= t_lo:
            return
        lead = min(t_lo, 2)
        s0 = HOP * (t_lo - lead)
        spec, lm = K.stft(self.wav[:, s0:n_end], want_logmag=True)
        self.mix[:, t_lo:t_hi] = spec[:, lead:lead + t_hi - t_lo]
        self.logmag[:, t_lo:t_hi] = lm[:, lead:lead + t_hi - t_lo]

    def feed(self, chunk):
        """append audio (1-D float32, host or device); transforms every frame the new samples complete"""
        chunk = torch.as_tensor(chunk, dtype=torch.float32).reshape(-1)
        m = chunk.numel()
        if self.n + m > self.wav.shape[1] - HOP:
            raise ValueError('stream longer than the %d samples this separator was built for' % (self.wav.shape[1] - HOP))
        self.wav[0, self.n:self.n + m].copy_(chunk, non_blocking=True)
        self.n += m
        # frame t is final when sample 64 t + 127 has arrived; keep t_done even (frame pairing of the kernel) and leave
        # two more frames of margin so the slice handed to K1 ends beyond the last wanted frame's support
        t_final = (self.n - HALF) // HOP + 1 if self.n >= HALF else 0
        t_hi = max(self.t_done, (t_final - 2) // 2 * 2)
        if t_hi - self.t_done >= 2 and self.n >= K.FFT_SIZE:
            self._transform(self.t_done, t_hi, self.n)
            self.t_done = t_hi
        return self.t_done

    def finish(self, out_host=None):
        """end of stream: the remaining frames, then encoder -> attractors -> K4, and the device-to-host copy in blocks;
        returns (separated waveforms [C, 64 T] in pinned host memory, list of (first sample, CUDA event) per block --
        block i is in `out_host` once its event has completed)"""
        if self.n < K.FFT_SIZE:
            raise ValueError('window is longer than input signal (%d < %d)' % (self.n, K.FFT_SIZE))
        model = self.model
        T = K.num_frames(self.n)
        Cn = hparams.MAX_N_SIGNAL
        self._transform(self.t_done, T, self.n)        # the tail sees the true right padding: same call as the batch path
        mix, logmag = self.mix[:, :T].contiguous(), self.logmag[:, :T].contiguous()
        if out_host is None:
            out_host = torch.empty((Cn, HOP * T), dtype=torch.float32).pin_memory()
        events = []
        cur = torch.cuda.current_stream()
        # the same call Model.separate makes per stream group: encoder (with the estimator's sums fused into its output
        # projection when the shapes allow) -> attractors -> K4.  utils.istft semantics (last 4 frames unused, division by
        # the window power) depend on the absolute frame index, so K4 runs once over the utterance (tens of microseconds)
        # and the blocks are cut from its output
        wav_dev = model.infer(mix, logmag=logmag, want='wav')[0]
        step = HOP * self.out_block
        self.copy_stream.wait_event(cur.record_event())
        with torch.cuda.stream(self.copy_stream):
            for s0 in range(0, HOP * T, step):
                s1 = min(HOP * T, s0 + step)
                out_host[:, s0:s1].copy_(wav_dev[:, s0:s1], non_blocking=True)
                events.append((s0, self.copy_stream.record_event()))
        cur.wait_stream(self.copy_stream)
        return out_host, events
