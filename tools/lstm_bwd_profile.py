"""Per-step phase timing of the tcgen05 BPTT kernel (DANET_LSTM_PROFILE=1): SM-clock stamps of CTA (0,0,0):
epilogue thread 0 slots 0-7, MMA thread slots 8-9."""
import os, sys
os.environ['DANET_LSTM_PROFILE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import danet_tensorflow_b200 as D
K = D.kernels
lib = D._lib.load()
B, T, H, I = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 501, 300, 600
torch.manual_seed(0)
r = .75 / np.sqrt(H)
Ws = [(torch.rand(I + H, 4 * H, device='cuda') * 2 - 1) * r for _ in range(2)]
pre = torch.randn(2, T, B, 4 * H, device='cuda')
out, cell = K.lstm_seq(pre, Ws, I, T, B, H, backend=1, keep_cell=True, keep_gates=True)
gates0 = pre.clone()
dout = torch.randn(B, T, 2 * H, device='cuda')
ptrs = (C.c_void_p * 2)(*[w.data_ptr() + I * 4 * H * 4 for w in Ws])
ws = torch.zeros(1 << 20, dtype=torch.uint8, device='cuda')
times = []
for it in range(4):
    g = gates0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.danet_lstm_seq_bwd(C.c_void_p(dout.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(cell.data_ptr()), ptrs, 4 * H,
                                2, T, B, H, C.c_void_p(ws.data_ptr()), ws.numel(), 1, None)
    assert rc == 0, lib.danet_last_error_string()
    e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e3)
prof = ws[:T * 16 * 8].view(torch.int64).view(T, 16).cpu().numpy()
names = ['epi:step_begin', 'epi:loads_issued', 'epi:p_full', 'epi:summed', 'epi:math', 'epi:staged(b_full)', 'epi:acc_full',
         'epi:slices_staged', 'mma:b_full', 'mma:issued']
s0, s1 = 100, 400
rel = prof[s0:s1, :10] - prof[s0:s1, 0:1]
print('kernel %.1f us (best of 3 warm), step period %.1f cycles' % (min(times[1:]), np.diff(prof[s0:s1, 0]).mean()))
for i, n in enumerate(names):
    print('   %-22s %8.1f' % (n, rel[:, i].mean()))
print('   slices_staged -> next p_full: %.1f' % (prof[s0 + 1:s1 + 1, 2] - prof[s0:s1, 7]).mean())
ref = K.lstm_seq_bwd(dout, gates0.clone(), cell, Ws, I, T, B, H, backend=0)
print('da vs fp32 cooperative kernel: max abs diff %.3g (scale %.3g)' % ((g - ref).abs().max().item(), ref.abs().max().item()))
