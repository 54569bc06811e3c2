"""Output projection + anchor estimator: the fused kernel (danet_proj_anchor_fwd) against the unfused pair
(danet_gemm_split + danet_attractor_anchor_fwd) at the bench shapes, one stream group (B = 8) and the whole batch (B = 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import danet_tensorflow_b200 as D
K = D.kernels
T, F, E, Kd = 501, 129, 20, 600
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for B in (8, 32):
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(B, T, Kd, device='cuda', generator=g) * .3
    W = (torch.rand(Kd, F * E, device='cuda', generator=g) * 2 - 1) * 1.85
    an = torch.randn(6, E, device='cuda', generator=g)
    a2, w2 = K.split_operand(x.view(B * T, Kd), False), K.split_operand(W, True)
    mu, cs = K.mean(x), K.colsum(W)

    def fused():
        return K.proj_anchor(a2, w2, B, T, F, E, Kd, an, row_mu=mu, col_s=cs)

    def unfused():
        v = K.gemm_split(a2, w2, B * T, F * E, Kd, row_mu=mu, col_s=cs, rows_per_mu=T)
        return v, K.attractor_anchor(v.view(B, T * F, E), an, 2)

    def gemm_only():
        return K.gemm_split(a2, w2, B * T, F * E, Kd, row_mu=mu, col_s=cs, rows_per_mu=T)

    for name, fn in (('fused proj+anchor', fused), ('gemm + attractor', unfused), ('gemm alone', gemm_only)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = float(np.median(ts))
        print('B %2d %-20s %7.1f us   %6.0f TFLOP/s issued (3 bf16 products)' % (B, name, us, 3 * 2. * B * T * Kd * F * E / us / 1e6))
    e, a = fused()
    v, a_ref = unfused()
    print('   embed max rel diff %.2e, attractors %.2e' % (float((e.view_as(v) - v).abs().max() / v.abs().max()),
                                                          float((a - a_ref).abs().max() / a_ref.abs().max())))
