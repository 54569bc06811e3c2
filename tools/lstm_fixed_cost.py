"""Fixed (launch + prologue + tail) cost of the tcgen05 recurrent kernel: time per launch for several T, packed weights,
backend 2, no profiling stamps.  The slope is the step time, the intercept everything else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import danet_tensorflow_b200 as D
K = D.kernels
B, H, I = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 300, 600
torch.manual_seed(0)
r = .75 / np.sqrt(H)
Ws = [(torch.rand(I + H, 4 * H, device='cuda') * 2 - 1) * r for _ in range(2)]
packed = K.lstm_pack_wh(Ws, I, H)
for backend in (2, 1):
    rows = []
    for T in (1, 2, 11, 101, 501):
        pre = torch.randn(T, B, 2, 4 * H, device='cuda')
        for want_split in (False, True):
            for _ in range(3):
                K.lstm_seq(pre, Ws, I, T, B, H, backend=backend, interleaved=True, want_split=want_split, wh_packed=packed)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                K.lstm_seq(pre, Ws, I, T, B, H, backend=backend, interleaved=True, want_split=want_split, wh_packed=packed)
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            rows.append((T, want_split, min(ts)))
    print('backend', backend, ' '.join('T=%d%s: %.1f us' % (t, '+split' if w else '', v) for t, w, v in rows))
