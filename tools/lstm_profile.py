"""Per-step phase timing of the tcgen05 LSTM kernels (DANET_LSTM_PROFILE=1): SM-clock stamps of CTA (0,0,0):
MMA thread slots 0-2, epilogue thread 0 slots 3-8, first sender slot 9; row 0 slots 10-13 = kernel entry, prologue done,
loop done, exit.  Runs the first-generation kernel (DANET_LSTM_V=1), the second generation reading fp32 weights, and the
second generation on pre-packed weights, and checks that all three agree."""
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
ws = torch.zeros(1 << 20, dtype=torch.uint8, device='cuda')
packed = K.lstm_pack_wh(Ws, I, H)
names = ['mma:wait_begin', 'mma:h_full', 'mma:issued', 'epi:step_begin', 'epi:acc_full', 'epi:tmem_ld', 'epi:gathered',
         'epi:activated', 'epi:bar', 'copies_issued']


def run(label, env, wh_packed, backend=1):
    for k in ('DANET_LSTM_V', 'DANET_LSTM_NOPACK'):
        os.environ.pop(k, None)
    os.environ.update(env)
    out = torch.empty(B, T, 2 * H, device='cuda')
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for it in range(4):
        ev[0].record()
        rc = lib.danet_lstm_seq_fwd_packed(C.c_void_p(pre.data_ptr()), 0, 0, ptrs, 4 * H,
                                           C.c_void_p(wh_packed.data_ptr()) if wh_packed is not None else None,
                                           C.c_void_p(out.data_ptr()), None, None, None, 0, 2, T, B, H,
                                           C.c_void_p(ws.data_ptr()), ws.numel(), backend, None)
        assert rc == 0, lib.danet_last_error_string()
        ev[1].record()
        torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]) * 1e3)
    prof = ws[:T * 16 * 8].view(torch.int64).view(T, 16).cpu().numpy()
    s0, s1 = 100, 400
    rel = prof[s0:s1, :10] - prof[s0:s1, 3:4]
    period = np.diff(prof[s0:s1, 3]).mean()
    print('== %s: kernel %.1f us (memset + launch, best of 3 warm), step period %.1f cycles -> %.1f us for %d steps'
          % (label, min(times[1:]), period, period * (T - 1) / 1965., T - 1))
    for i, n in enumerate(names):
        print('   %-20s %8.1f' % (n, rel[:, i].mean()))
    print('   copies_issued -> next h_full: %.1f   h_full -> issued: %.1f   issued -> acc_full: %.1f' % (
        (prof[s0 + 1:s1 + 1, 1] - prof[s0:s1, 9]).mean(), (prof[s0:s1, 2] - prof[s0:s1, 1]).mean(),
        (prof[s0:s1, 4] - prof[s0:s1, 2]).mean()))
    if prof[0, 10]:
        print('   prologue %d cycles, loop %d, tail %d' % (prof[0, 11] - prof[0, 10], prof[0, 12] - prof[0, 11],
                                                          prof[0, 13] - prof[0, 12]))
        ns = prof[0, 15] - prof[0, 14]
        print('   CTA (0,0,0) lived %.1f us by %%globaltimer = %d cycles -> SM clock %.0f MHz during the kernel'
              % (ns / 1e3, prof[0, 13] - prof[0, 10], (prof[0, 13] - prof[0, 10]) / (ns / 1e3)))
    return out


o1 = run('gen 1 (3 MMAs per K16, smem regroup)', {'DANET_LSTM_V': '1'}, None)
o2 = run('gen 2, fp32 weights', {}, None)
o3 = run('gen 2, packed weights', {}, packed)
o4 = run('gen 2, packed weights, fp16 recurrent state (backend 2)', {}, packed, backend=2)
ref = K.lstm_seq(pre, Ws, I, T, B, H, backend=0)
for nm, o in (('gen1', o1), ('gen2', o2), ('gen2 packed', o3), ('gen2 fp16-h', o4)):
    print('%s vs fp32 SIMT kernel: max abs diff %.3g' % (nm, (o - ref).abs().max().item()))
print('gen2 packed == gen2:', torch.equal(o2, o3))
