"""Anchor estimator alone (B utterances of T x 129 bins, E = 20, two sources): the tensor-core kernel against the
register-tiled SIMT kernel (DANET_ATTRACTOR_SIMT=1), cold L2 and warm; prints us per launch and GB/s of the one read of V."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import danet_tensorflow_b200 as D
K = D.kernels
T = 501
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for B in (8, 32):
    g = torch.Generator(device='cuda').manual_seed(1)
    V = torch.randn(B, T * 129, 20, device='cuda', generator=g) * 2.
    an = torch.randn(6, 20, device='cuda', generator=g)
    res = {}
    for simt in (0, 1):
        if simt:
            os.environ['DANET_ATTRACTOR_SIMT'] = '1'
        else:
            os.environ.pop('DANET_ATTRACTOR_SIMT', None)
        for _ in range(3):
            out = K.attractor_anchor(V, an, 2, return_all=True)
        torch.cuda.synchronize()
        for cold in (1, 0):
            ts = []
            for _ in range(7):
                if cold:
                    flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); K.attractor_anchor(V, an, 2); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            us = float(np.median(ts))
            print('B %2d %-5s %-4s %7.1f us  %6.0f GB/s' % (B, 'SIMT' if simt else 'MMA', 'cold' if cold else 'warm', us,
                                                         V.numel() * 4 / us / 1e3))
        res[simt] = out
    for i, nm in enumerate(('attractors', 'sets', 'sims')):
        a, b = res[0][i], res[1][i]
        print('   %s: MMA vs SIMT max rel diff %.2e' % (nm, float((a - b).abs().max() / b.abs().max())))
    print('   choice equal:', bool(torch.equal(res[0][3], res[1][3])))
os.environ.pop('DANET_ATTRACTOR_SIMT', None)
