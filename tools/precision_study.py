"""Which operand formats can the RECURRENT product h_{t-1} * Wh use and stay inside the 1e-3 parity gate?
CPU emulation on the oracle (fp64 everywhere except the emulated rounding of the recurrent operands; the hoisted
input products and the output projection stay exact, as the bf16x3 GEMMs are ~1e-5).  Prints the max-norm relative
error of the embedding and of the separated spectra against the fp64 oracle, for the bench's synthetic mixtures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import danet_oracle as O
import bench

torch.set_num_threads(os.cpu_count())
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32000
wav = bench.synth_mixtures(Bn, N, 1337)
P = O.reference_init(1337, estimators=('infer_estimator',), dtype=torch.float64)
mix = torch.from_numpy(np.stack([O.stft(w) for w in wav]))            # [B,T,F] complex
x0 = torch.log1p(mix.abs()).double()


def rnd(x, fmt):
    if fmt == 'exact':
        return x
    if fmt == 'fp16':
        return x.to(torch.float32).to(torch.float16).to(torch.float64)
    if fmt == 'bf16':
        return x.to(torch.float32).to(torch.bfloat16).to(torch.float64)
    if fmt == 'bf16x2':     # hi + lo, both bf16
        x32 = x.to(torch.float32)
        hi = x32.to(torch.bfloat16).to(torch.float32)
        lo = (x32 - hi).to(torch.bfloat16).to(torch.float32)
        return (hi + lo).to(torch.float64)
    if fmt == 'fp16x2':
        x32 = x.to(torch.float32)
        hi = x32.to(torch.float16).to(torch.float32)
        lo = (x32 - hi).to(torch.float16).to(torch.float32)
        return (hi + lo).to(torch.float64)
    if fmt == 'tf32':
        i = x.to(torch.float32).view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32).to(torch.float64)
    raise KeyError(fmt)


def lstm_layer(x, W, B, hfmt, wfmt, xfmt='exact'):
    Bsz, T, I = x.shape
    H = B.shape[0] // 4
    pre = rnd(x, xfmt) @ W[:I] + B          # hoisted input projection: activations in `xfmt`, weights exact (hi + lo)
    Wh = rnd(W[I:], wfmt)
    c = x.new_zeros(Bsz, H); h = x.new_zeros(Bsz, H)
    out = []
    for t in range(T):
        a = pre[:, t] + rnd(h, hfmt) @ Wh
        g = a[:, :H]
        i, f, o = torch.sigmoid(a[:, H:]).split(H, dim=-1)
        c = i * g + f * c
        h = (o * torch.tanh(c)).to(torch.float32).to(torch.float64)     # h is produced in fp32
        out.append(h)
    return torch.stack(out, 1)


def encoder(x, hfmt, wfmt, xfmt='exact', pfmt='exact'):
    x = x - x.mean(dim=(1, 2), keepdim=True)
    for l in range(4):
        n = 'encoder/lstm%d_%s/LSTM/linear/'
        hf = lstm_layer(x, P[n % (l, 'fwd') + 'W'], P[n % (l, 'fwd') + 'B'], hfmt, wfmt, xfmt)
        hb = torch.flip(lstm_layer(torch.flip(x, [1]), P[n % (l, 'bwd') + 'W'], P[n % (l, 'bwd') + 'B'], hfmt, wfmt, xfmt), [1])
        x = torch.cat([hf, hb], -1)
    # output projection on the UNcentred operand (the centring is a rank-1 epilogue term): activations in `pfmt`
    mu = x.mean(dim=(1, 2), keepdim=True)
    W = P['encoder/output/W']
    return (rnd(x, pfmt) @ W - mu * W.sum(0)).reshape(x.shape[0], x.shape[1], 129, 20)


def separated(V):
    A = O.estimator_anchor(V, P['infer_estimator/anchors'], 2)
    pw = O.separator(mix.abs().double(), A, V.reshape(V.shape[0], -1, 20), 'dot-softmax-orig')
    return pw


ref = encoder(x0, 'exact', 'exact')
ref_s = separated(ref)
print('T = %d, B = %d; max-norm relative error vs fp64' % (x0.shape[1], Bn))
print('%-34s %12s %12s' % ('recurrent operands (h / Wh)', 'embedding', 'spectra'))
for hfmt, wfmt in (('bf16x2', 'bf16x2'), ('fp16', 'bf16x2'), ('fp16', 'fp16x2'), ('fp16', 'fp16'), ('tf32', 'tf32'),
                   ('bf16', 'bf16x2'), ('bf16', 'bf16')):
    V = encoder(x0, hfmt, wfmt)
    e = ((V - ref).abs().max() / ref.abs().max()).item()
    s = separated(V)
    es = ((s - ref_s).abs().max() / ref_s.abs().max()).item()
    print('%-34s %12.3g %12.3g' % (hfmt + ' / ' + wfmt, e, es))
print('dense products too: activations of the hoisted input products (x) / of the output projection (p) as one fp16 value')
for hfmt, wfmt, xfmt, pfmt in (('fp16', 'fp16x2', 'fp16', 'exact'), ('fp16', 'fp16x2', 'fp16', 'fp16'),
                               ('fp16', 'fp16', 'fp16', 'fp16'), ('exact', 'exact', 'fp16', 'fp16')):
    V = encoder(x0, hfmt, wfmt, xfmt, pfmt)
    e = ((V - ref).abs().max() / ref.abs().max()).item()
    s = separated(V)
    es = ((s - ref_s).abs().max() / ref_s.abs().max()).item()
    print('%-34s %12.3g %12.3g' % ('%s / %s, x %s, p %s' % (hfmt, wfmt, xfmt, pfmt), e, es))
