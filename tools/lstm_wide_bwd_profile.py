"""Where a step of the wide BPTT kernel goes (csrc/lstm_wide_bwd_tc.cu, DANET_LSTM_PROFILE=1): cycles of CTA (0,0,0)'s first
epilogue thread, averaged over the steps: gather of the partial slices / da, scale and staging / wait for the accumulators /
TMEM loads + publishing."""
import os, sys
os.environ['DANET_LSTM_PROFILE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import danet_tensorflow_b200 as D
K = D.kernels
lib = D._lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H, T, I = 600, 501, 600
torch.manual_seed(0)
r = 1.15 / np.sqrt(H)
W = (torch.rand(I + H, 4 * H, device='cuda') * 2 - 1) * r
pre = torch.randn(1, T, B, 4 * H, device='cuda')
out, cell = K.lstm_seq(pre, [W], I, T, B, H, backend=2, keep_cell=True, keep_gates=True)
gates0 = pre.clone()
dout = torch.randn(B, T, H, device='cuda') * 1e-3
need = lib.danet_lstm_seq_bwd_workspace_bytes(1, B, H)
ws = torch.zeros(need, dtype=torch.uint8, device='cuda')
ptrs = (C.c_void_p * 1)(W.data_ptr() + I * 4 * H * 4)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
times = []
for it in range(4):
    g = gates0.clone()
    ev[0].record()
    rc = lib.danet_lstm_seq_bwd(C.c_void_p(dout.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(cell.data_ptr()), ptrs, 4 * H,
                                1, T, B, H, C.c_void_p(ws.data_ptr()), ws.numel(), 2, None)
    assert rc == 0, lib.danet_last_error_string()
    ev[1].record(); torch.cuda.synchronize()
    times.append(ev[0].elapsed_time(ev[1]) * 1e3)
groups, ncta = (B + 7) // 8, (H + 31) // 32
off = groups * 2 * ncta * ncta * 256 * 8
prof = ws[off:off + 40].view(torch.int64).cpu().numpy() / float(T - 1)
print('   the last source\'s slice for CTA 0 arrives %.0f cycles after CTA 0 finished publishing (mean over the steps)' % prof[4])
prof = prof[:4]
print('wide BPTT kernel, B = %d, H = %d: %.1f us per launch (best of 3 warm) = %.0f ns per step' % (B, H, min(times[1:]), min(times[1:]) * 1e3 / T))
print('   cycles per step: gather %.0f, da / scale / staging %.0f, wait for the accumulators %.0f, TMEM loads + publishing %.0f, sum %.0f'
      % (prof[0], prof[1], prof[2], prof[3], prof.sum()))
ref = K.lstm_seq_bwd(dout, gates0.clone(), cell, [W], I, T, B, H, backend=0)
print('da vs the fp32 kernel: max abs diff %.3g of max %.3g' % (float((g - ref).abs().max()), float(ref.abs().max())))
