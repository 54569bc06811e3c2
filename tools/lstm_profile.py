"""Per-step phase timing of the tcgen05 LSTM kernel (DANET_LSTM_PROFILE=1): SM-clock stamps of
CTA (0,0,0): MMA thread slots 0-2, epilogue thread 0 slots 3-9."""
import os, sys
os.environ['DANET_LSTM_PROFILE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import danet_tensorflow_b200 as D
K = D.kernels
lib = D._lib.load()
B, T, H, I = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 501, 300, 600
torch.manual_seed(0)
pre = torch.randn(2, T, B, 4 * H, device='cuda')
r = .75 / np.sqrt(H)
Ws = [(torch.rand(I + H, 4 * H, device='cuda') * 2 - 1) * r for _ in range(2)]
ptrs = (C.c_void_p * 2)(*[w.data_ptr() + I * 4 * H * 4 for w in Ws])
out = torch.empty(B, T, 2 * H, device='cuda')
ws = torch.zeros(1 << 20, dtype=torch.uint8, device='cuda')
for _ in range(2):
    rc = lib.danet_lstm_seq_fwd(C.c_void_p(pre.data_ptr()), 0, 0, ptrs, 4 * H, C.c_void_p(out.data_ptr()), None, None, None, 0, 2, T, B, H,
                                C.c_void_p(ws.data_ptr()), ws.numel(), 1, None)
    assert rc == 0, lib.danet_last_error_string()
torch.cuda.synchronize()
prof = ws[:T * 16 * 8].view(torch.int64).view(T, 16).cpu().numpy()
names = ['mma:wait_begin', 'mma:h_full', 'mma:issued', 'epi:step_begin', 'epi:acc_full', 'epi:tmem_ld', 'epi:gathered',
         'epi:activated', 'epi:bar', 'epi:copies_issued']
s0, s1 = 100, 400
base = prof[s0:s1, 3:4]
rel = prof[s0:s1, :10] - base
print('step period (cycles):', np.diff(prof[s0:s1, 3]).mean())
for i, n in enumerate(names):
    print('%-20s %8.1f' % (n, rel[:, i].mean()))
# next step's h_full relative to this step's copies issued
print('copies_issued -> next h_full:', (prof[s0 + 1:s1 + 1, 1] - prof[s0:s1, 9]).mean())
print('h_full -> issued:', (prof[s0:s1, 2] - prof[s0:s1, 1]).mean(), ' issued -> acc_full:', (prof[s0:s1, 4] - prof[s0:s1, 2]).mean())
