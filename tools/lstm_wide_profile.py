"""Per-step phase timing of the wide recurrent kernel (csrc/lstm_wide_tc.cu, DANET_LSTM_PROFILE=1): SM-clock stamps of
CTA (0,0,0).  MMA thread: 0 wait begin, 1 h_full, 2 MMAs issued + committed.  Epilogue thread 0: 3 accumulator ready,
4 own slice of h published, 5 all slices gathered + arrived.  Row T: 0 kernel entry, 1 prologue done, 2 loop done."""
import os, sys
os.environ['DANET_LSTM_PROFILE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import danet_tensorflow_b200 as D
K = D.kernels
lib = D._lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 600
T, I = 501, 600
torch.manual_seed(0)
pre = torch.randn(1, T, B, 4 * H, device='cuda')
r = 1.15 / np.sqrt(H)
Ws = [(torch.rand(I + H, 4 * H, device='cuda') * 2 - 1) * r]
ptrs = (C.c_void_p * 1)(Ws[0].data_ptr() + I * 4 * H * 4)
need = lib.danet_lstm_seq_workspace_bytes(1, B, H)
ws = torch.zeros(need + (T + 1) * 64 + 4096, dtype=torch.uint8, device='cuda')
packed = K.lstm_pack_wh(Ws, I, H)
out = torch.empty(B, T, H, device='cuda')
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
times = []
for it in range(4):
    ev[0].record()
    rc = lib.danet_lstm_seq_fwd_packed(C.c_void_p(pre.data_ptr()), 0, 0, ptrs, 4 * H, C.c_void_p(packed.data_ptr()),
                                       C.c_void_p(out.data_ptr()), None, None, None, 0, 1, T, B, H,
                                       C.c_void_p(ws.data_ptr()), ws.numel(), 2, None)
    assert rc == 0, lib.danet_last_error_string()
    ev[1].record()
    torch.cuda.synchronize()
    times.append(ev[0].elapsed_time(ev[1]) * 1e3)
xch = (2 * ((B + 7) // 8) * ((H + 31) // 32) * 128 * 8 + 255) // 256 * 256
off = xch + lib.danet_lstm_pack_wh_bytes(1, H)
prof = ws[off:off + (T + 1) * 64].view(torch.int64).view(T + 1, 8).cpu().numpy()
s0, s1 = 100, 400
period = np.diff(prof[s0:s1, 3]).mean()
print('== wide kernel, B = %d, H = %d: %.1f us per launch (best of 3 warm), step period %.1f cycles' % (B, H, min(times[1:]), period))
p = prof
print('   h_full -> MMAs issued      %8.1f' % (p[s0:s1, 2] - p[s0:s1, 1]).mean())
print('   issued -> acc ready (epi)  %8.1f' % (p[s0:s1, 3] - p[s0:s1, 2]).mean())
print('   acc ready -> h published   %8.1f' % (p[s0:s1, 4] - p[s0:s1, 3]).mean())
print('   published -> gathered      %8.1f' % (p[s0:s1, 5] - p[s0:s1, 4]).mean())
print('   gathered(t0) -> h_full(MMA)%8.1f' % (p[s0 + 1:s1 + 1, 1] - p[s0:s1, 5]).mean())
print('   prologue %d cycles, loop %d' % (p[T, 1] - p[T, 0], p[T, 2] - p[T, 1]))
ref = K.lstm_seq(pre, Ws, I, T, B, H, backend=0)
print('vs fp32 SIMT kernel: max abs diff %.3g' % (out - ref).abs().max().item())
