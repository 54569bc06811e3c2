"""
Torch-tensor front of the C-ABI (include/danet.h).  PyTorch is plumbing here: it owns
device memory and the stream; every computation is a call into libdanet_sm100.so with raw
device pointers.  Argument problems raise ValueError, device problems RuntimeError.
"""
import ctypes as C
import os

import torch

from . import _lib

FEATURE = 129
FFT_SIZE = 256
FFT_STRIDE = 64

# 0 = exact fp32 SIMT kernels, 1 = tcgen05 (bf16x3 split operands, fp32 accumulate)
DEFAULT_BACKEND = int(os.environ.get('DANET_BACKEND', '1'))

TC_LSTM_MAX_H = 384
# 384 < H <= 608 (the `lstm-orig` encoder): backend 2 runs the wide tcgen05 kernel (csrc/lstm_wide_tc.cu: groups of
# ceil(H/32) CTAs exchanging h through L2); backend 1 has no kernel there, backend 0 is the exact fp32 one
TC_WIDE_MAX_H = 608

# launch counter: bench.py reports how many of OUR kernels ran inside the timed region
launches = 0


def _count(n=1):
    global launches
    launches += n


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, name, dtype=torch.float32, dim=None):
    if not isinstance(t, torch.Tensor):
        raise ValueError('%s: expected a torch.Tensor' % name)
    if not t.is_cuda:
        raise ValueError('%s: expected a CUDA tensor (there is no CPU path)' % name)
    if t.dtype != dtype:
        raise ValueError('%s: expected dtype %s, got %s' % (name, dtype, t.dtype))
    if dim is not None and t.dim() != dim:
        raise ValueError('%s: expected %d dims, got shape %s' % (name, dim, tuple(t.shape)))
    return t.contiguous()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


_timeline = None      # (int64 device tensor, list of labels) while tools/timeline.py records
_prof_ws = None       # list collecting the pipelined recurrences' workspaces (their in-kernel stamps) for tools/timeline.py


def stamp(label):
    """profiling aid: record %globaltimer on the current stream (no-op unless a timeline is being recorded)"""
    if _timeline is None:
        return
    buf, labels = _timeline
    i = len(labels)
    labels.append(label)
    _lib.check(_lib.load().danet_timestamp(C.c_void_p(buf.data_ptr() + 8 * i), _stream()), 'timestamp')


def num_frames(n_samples):
    """T = ceil(N/64) + 1 (scipy.signal.stft padding as used at app/utils.py:117-122)"""
    t = _lib.load().danet_stft_num_frames(int(n_samples))
    if t < 0:
        raise ValueError('window is longer than input signal (%d < 256)' % n_samples)
    return t


def stft(wav, want_logmag=False):
    """wav f32 [..., N] -> spec c64 [..., T, 129] (+ log1p|spec| f32)   [app/utils.py:117-122]"""
    wav = _req(wav, 'wav')
    lead, n = wav.shape[:-1], wav.shape[-1]
    T = num_frames(n)
    n_sig = int(torch.Size(lead).numel())
    spec = torch.empty(lead + (T, FEATURE), dtype=torch.complex64, device=wav.device)
    logmag = torch.empty(lead + (T, FEATURE), dtype=torch.float32, device=wav.device) if want_logmag else None
    _lib.check(_lib.load().danet_stft_fwd(_p(wav), n_sig, n, _p(spec), _p(logmag), _stream()), 'stft')
    _count()
    return (spec, logmag) if want_logmag else spec


def istft(spec, out=None):
    """spec c64 [..., T, 129] -> wav f32 [..., 64*T] with utils.istft semantics [app/utils.py:53-75]"""
    spec = _req(spec, 'spec', torch.complex64)
    if spec.dim() < 2 or spec.shape[-1] != FEATURE:
        raise ValueError('spec: expected [..., T, 129], got %s' % (tuple(spec.shape),))
    lead, T = spec.shape[:-2], spec.shape[-2]
    n_sig = int(torch.Size(lead).numel())
    if out is None:
        wav = torch.empty(lead + (FFT_STRIDE * T,), dtype=torch.float32, device=spec.device)
    else:
        wav = out
        if tuple(wav.shape) != tuple(lead) + (FFT_STRIDE * T,) or wav.dtype != torch.float32 or not wav.is_contiguous():
            raise ValueError('istft: out must be a contiguous float32 %s' % (tuple(lead) + (FFT_STRIDE * T,),))
    _lib.check(_lib.load().danet_istft_fwd(_p(spec), n_sig, T, _p(wav), _stream()), 'istft')
    _count()
    return wav


def mix_features(src, want=('mix', 'src_pwr', 'mix_pwr', 'logmag')):
    """src c64 [B,C,T,F] -> dict(mix c64 [B,T,F], src_pwr [B,C,T,F], mix_pwr, logmag [B,T,F])
    [main.py:233-240]"""
    src = _req(src, 'src', torch.complex64, 4)
    B, Cn, T, F = src.shape
    dev = src.device
    out = {
        'mix': torch.empty((B, T, F), dtype=torch.complex64, device=dev) if 'mix' in want else None,
        'src_pwr': torch.empty((B, Cn, T, F), dtype=torch.float32, device=dev) if 'src_pwr' in want else None,
        'mix_pwr': torch.empty((B, T, F), dtype=torch.float32, device=dev) if 'mix_pwr' in want else None,
        'logmag': torch.empty((B, T, F), dtype=torch.float32, device=dev) if 'logmag' in want else None,
    }
    _lib.check(_lib.load().danet_mix_features_fwd(
        _p(src), B, Cn, T * F, _p(out['mix']), _p(out['src_pwr']), _p(out['mix_pwr']), _p(out['logmag']),
        _stream()), 'mix_features')
    _count()
    return out


def center(x, out=None):
    """x [B, ...] minus its per-utterance mean over all other axes [app/modules.py:209-210, 244-245]"""
    x = _req(x, 'x')
    B = x.shape[0]
    n_per = x[0].numel() if B else 1
    y = torch.empty_like(x) if out is None else out
    lib = _lib.load()
    ws = _ws(lib.danet_center_workspace_bytes(B), x.device)
    _lib.check(lib.danet_center_fwd(_p(x), B, n_per, _p(y), _p(ws), _stream()), 'center')
    _count(2)
    return y


def mean(x):
    """per-utterance mean over all other axes -> [B] (the statistic of `center`)"""
    x = _req(x, 'x')
    B = x.shape[0]
    out = torch.empty((B,), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    ws = _ws(lib.danet_center_workspace_bytes(B), x.device)
    _lib.check(lib.danet_mean_fwd(_p(x), B, x[0].numel() if B else 1, _p(out), _p(ws), _stream()), 'mean')
    _count(2)
    return out


def leaky_relu(x, leak=0., inplace=True):
    """max(x*leak, x) [app/ops.py:93-107]"""
    x = _req(x, 'x')
    y = x if inplace else torch.empty_like(x)
    _lib.check(_lib.load().danet_leaky_relu_fwd(_p(x), _p(y), x.numel(), float(leak), _stream()), 'leaky_relu')
    _count()
    return y


def leaky_relu_bwd(y, dy, leak=0.):
    """in place on dy: dy * (y > 0 ? 1 : leak)"""
    y, dy = _req(y, 'y'), _req(dy, 'dy')
    _lib.check(_lib.load().danet_leaky_relu_bwd(_p(y), _p(dy), _p(dy), y.numel(), float(leak), _stream()), 'leaky_relu_bwd')
    _count()
    return dy


def linear(a, w, bias=None, time_major_T=0, backend=None, k_rows=None, row_offset=0, out=None):
    """
    a [M,K] @ w[row_offset : row_offset+K, :] (+ bias) -> [M,N]   [app/ops.py:72-89]
    `w` is the reference's stacked [I+H, 4H] matrix or a plain [K,N]; rows are selected without
    a copy.  time_major_T > 0 writes logical row b*T+t to row t*(M/T)+b.
    """
    a = _req(a, 'a', dim=2)
    w = _req(w, 'w', dim=2)
    M, K = a.shape
    N = w.shape[1]
    if k_rows is None:
        k_rows = w.shape[0] - row_offset
    if k_rows != K or row_offset + K > w.shape[0]:
        raise ValueError('linear: a is [%d,%d] but w rows %d..%d of %d' % (M, K, row_offset, row_offset + k_rows, w.shape[0]))
    if bias is not None:
        bias = _req(bias, 'bias', dim=1)
        if bias.shape[0] != N:
            raise ValueError('linear: bias has %d entries, N = %d' % (bias.shape[0], N))
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    elif tuple(out.shape) != (M, N) or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError('linear: out must be a contiguous float32 [%d,%d]' % (M, N))
    wp = C.c_void_p(w.data_ptr() + row_offset * N * 4)
    be = DEFAULT_BACKEND if backend is None else backend
    lib = _lib.load()
    ws = _ws(lib.danet_linear_workspace_bytes(M, N, K, be), a.device)
    _lib.check(lib.danet_linear_fwd(_p(a), K, wp, N, _p(bias), _p(out), M, N, K, int(time_major_T),
                                    _p(ws), ws.numel(), be, _stream()), 'linear')
    _count(3 if be == 1 else 1)
    return out


def gemm(a, b, trans_a=False, trans_b=False, bias=None, out=None, ldc=None, perm_a_T=0, shift_a=0, out_perm_T=0,
         accumulate=False, m=None, n=None, k=None, lda=None, ldb=None):
    """
    General tensor-core product of the training step: C[M,N] (+)= A'[M,K] @ B'[K,N] (+ bias).
    `a` / `b` are 2-D float32 tensors (or views with a row stride); sizes default to their shapes.
    See include/danet.h (danet_gemm) for perm / shift / out_perm.
    """
    a = _req_strided(a, 'a')
    b = _req_strided(b, 'b')
    lda = a.stride(0) if lda is None else lda
    ldb = b.stride(0) if ldb is None else ldb
    M = (a.shape[1] if trans_a else a.shape[0]) if m is None else m
    Kd = (a.shape[0] if trans_a else a.shape[1]) if k is None else k
    N = (b.shape[0] if trans_b else b.shape[1]) if n is None else n
    kb = b.shape[1] if trans_b else b.shape[0]
    if k is None and kb != Kd:
        raise ValueError('gemm: reduction sizes differ (%d vs %d)' % (Kd, kb))
    if out is None:
        if accumulate:
            raise ValueError('gemm: accumulate needs out')
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ldc = out.stride(0) if ldc is None else ldc
    lib = _lib.load()
    ws = _ws(lib.danet_gemm_workspace_bytes(M, N, Kd), a.device)
    _lib.check(lib.danet_gemm(_p(a), lda, int(trans_a), int(perm_a_T), int(shift_a), _p(b), ldb, int(trans_b),
                              _p(bias), _p(out), ldc, M, N, Kd, int(out_perm_T), int(accumulate), _p(ws),
                              ws.numel(), _stream()), 'gemm')
    _count(3)
    return out


def _req_strided(t, name):
    """2-D float32 CUDA tensor whose rows are contiguous (row stride >= width): views are fine"""
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2:
        raise ValueError('%s: expected a 2-D float32 CUDA tensor' % name)
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def lstm_pack_wh(w_list, in_dim, H):
    """recurrent rows of the stacked [I+H,4H] matrices -> the tcgen05 recurrent kernel's TMEM image (bf16 hi/lo),
    or None when H is outside that backend's range.  Valid until the weights change."""
    lib = _lib.load()
    n_dir = len(w_list)
    nbytes = lib.danet_lstm_pack_wh_bytes(n_dir, H)
    if nbytes == 0:
        return None
    ptrs = (C.c_void_p * n_dir)()
    for d, w in enumerate(w_list):
        w = _req(w, 'W[%d]' % d, dim=2)
        ptrs[d] = w.data_ptr() + in_dim * 4 * H * 4
    packed = torch.empty((nbytes,), dtype=torch.uint8, device=w_list[0].device)
    _lib.check(lib.danet_lstm_pack_wh(ptrs, 4 * H, n_dir, H, _p(packed), nbytes, _stream()), 'lstm_pack_wh')
    _count()
    return packed


def lstm_seq(pre, w_list, in_dim, T, B, H, backend=None, keep_cell=False, keep_gates=False, interleaved=False,
             want_split=False, wh_packed=None):
    """
    pre [n_dir,T,B,4H]; w_list = the reference's stacked [I+H,4H] matrices, one per direction
    (recurrent rows start at `in_dim`) -> hidden [B,T,n_dir*H] (+ cell [n_dir,T,B,H]).
    keep_gates: the post-activation gates [g|i|f|o] overwrite `pre` in place (training).
    interleaved: `pre` is [T,B,n_dir,4H] (one product for both directions) instead of [n_dir,T,B,4H].
    want_split: also return the hidden sequence as a ready-made tensor-core operand (split_rows layout).
    [main.py:76-132; app/ops.py:139-147; app/modules.py:120-137]
    """
    pre = _req(pre, 'pre', dim=4)
    n_dir = len(w_list)
    want_shape = (T, B, n_dir, 4 * H) if interleaved else (n_dir, T, B, 4 * H)
    if tuple(pre.shape) != want_shape:
        raise ValueError('lstm_seq: pre is %s, expected %s' % (tuple(pre.shape), want_shape))
    dir_stride, row_stride = (4 * H, n_dir * 4 * H) if interleaved else (0, 0)
    ptrs = (C.c_void_p * n_dir)()
    keep = []
    for d, w in enumerate(w_list):
        w = _req(w, 'W[%d]' % d, dim=2)
        if tuple(w.shape) != (in_dim + H, 4 * H):
            raise ValueError('lstm_seq: W[%d] is %s, expected %s' % (d, tuple(w.shape), (in_dim + H, 4 * H)))
        keep.append(w)
        ptrs[d] = w.data_ptr() + in_dim * 4 * H * 4
    out = torch.empty((B, T, n_dir * H), dtype=torch.float32, device=pre.device)
    cell = torch.empty((n_dir, T, B, H), dtype=torch.float32, device=pre.device) if keep_cell else None
    lib = _lib.load()
    nws = lib.danet_lstm_seq_workspace_bytes(n_dir, B, H)
    ws = _ws(nws, pre.device)
    if backend is None:
        # the tcgen05 cluster kernel keeps Wh in one cluster's shared memory: H <= 320
        be = DEFAULT_BACKEND if H <= TC_LSTM_MAX_H else 0
    else:
        be = backend
    out_split, kp = None, 0
    if want_split:
        kp = (n_dir * H + 63) // 64 * 64
        out_split = torch.empty((2, B * T, kp), dtype=torch.bfloat16, device=pre.device)
    _lib.check(lib.danet_lstm_seq_fwd_packed(_p(pre), dir_stride, row_stride, ptrs, 4 * H,
                                             _p(wh_packed) if be >= 1 else None, _p(out), _p(cell),
                                             _p(pre) if keep_gates else None, _p(out_split), kp, n_dir, T, B, H,
                                             _p(ws), ws.numel(), be, _stream()), 'lstm_seq')
    _count()
    if want_split and keep_cell:
        return out, cell, out_split
    if want_split:
        return out, out_split
    return (out, cell) if keep_cell else out


PIPELINE_MAX_ROWS = 64 * 128


def pipeline_flags(device, sets=None):
    """the 64 completion flags of gemm_split_pipelined (or `sets` x 64 of them), cleared on the current stream
    (danet_zero_async)"""
    flags = torch.empty((64,) if sets is None else (int(sets), 64), dtype=torch.int32, device=device)
    _lib.check(_lib.load().danet_zero_async(_p(flags), flags.numel() * 4, _stream()), 'zero_async')
    _count()
    return flags


def gemm_split_pipelined(a2, b2, m, n, k, T, flags, bias=None, rows_tm=False):
    """gemm_split(out_perm_T=T) that publishes its row tiles through `flags` (int32[64], zeroed by the caller) in the order
    a forward and a backward scan need them -> (C [m,n], flag_need)  [include/danet.h: danet_gemm_split_pipelined].
    rows_tm: the rows of a2 are already time-major (t*B + b)."""
    out = torch.empty((m, n), dtype=torch.float32, device=a2.device)
    need = C.c_int(0)
    _lib.check(_lib.load().danet_gemm_split_pipelined(_p(a2), _p(b2), _p(bias), _p(out), n, m, n, k, int(T), int(rows_tm),
                                                      _p(flags), C.byref(need), _stream()), 'gemm_split_pipelined')
    _count()
    return out, need.value


def lstm_seq_pipelined(pre, w_list, in_dim, T, B, H, flags, flag_need, backend=2, wh_packed=None, pre_tm=False,
                       split_tm=False, programmatic=False):
    """lstm_seq(interleaved=True, want_split=True) on input projections that are still being produced by
    gemm_split_pipelined on another stream -> (hidden [B,T,n_dir*H], its split operand).
    pre_tm: the producer ran with rows_tm; split_tm: emit the split operand with time-major rows (t*B + b);
    programmatic: launch as a programmatic dependent of the previous kernel on the current stream (include/danet.h)."""
    pre = _req(pre, 'pre', dim=4)
    n_dir = len(w_list)
    if tuple(pre.shape) != (T, B, n_dir, 4 * H):
        raise ValueError('lstm_seq_pipelined: pre is %s, expected %s' % (tuple(pre.shape), (T, B, n_dir, 4 * H)))
    ptrs = (C.c_void_p * n_dir)()
    for d, w in enumerate(w_list):
        w = _req(w, 'W[%d]' % d, dim=2)
        ptrs[d] = w.data_ptr() + in_dim * 4 * H * 4
    out = torch.empty((B, T, n_dir * H), dtype=torch.float32, device=pre.device)
    kp = (n_dir * H + 63) // 64 * 64
    out_split = torch.empty((2, B * T, kp), dtype=torch.bfloat16, device=pre.device)
    lib = _lib.load()
    ws = _ws(lib.danet_lstm_seq_workspace_bytes(n_dir, B, H), pre.device)
    if _prof_ws is not None:             # tools/timeline.py with DANET_LSTM_PROFILE=1: room for the kernel's clock stamps
        ws = torch.zeros(T * 16 * 8 + 256, dtype=torch.uint8, device=pre.device)
        _prof_ws.append(ws)
    _lib.check(lib.danet_lstm_seq_fwd_pipelined(_p(pre), 4 * H, n_dir * 4 * H, ptrs, 4 * H, _p(wh_packed), _p(out),
                                                _p(out_split), kp, n_dir, T, B, H, _p(flags), int(flag_need), int(pre_tm),
                                                int(split_tm), int(programmatic), _p(ws), ws.numel(), int(backend),
                                                _stream()),
               'lstm_seq_pipelined')
    _count()
    return out, out_split


def split_operand(x, k_major_rows, out=None, row0=0, rows_total=None):
    """fp32 [rows,K] (k_major_rows=False) or [K,rows] (True) -> bf16 [2*rows_total, Kp] hi/lo operand"""
    x = _req_strided(x, 'x')
    rows, Kd = (x.shape[1], x.shape[0]) if k_major_rows else (x.shape[0], x.shape[1])
    rows_total = rows if rows_total is None else rows_total
    kp = (Kd + 63) // 64 * 64
    if out is None:
        out = torch.empty((2 * rows_total, kp), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().danet_split_operand(_p(x), x.stride(0), int(k_major_rows), rows, Kd, _p(out), row0, rows_total,
                                               _stream()), 'split_operand')
    _count()
    return out


def split_operand_time_major(x, T):
    """batch-major fp32 [B*T, K] -> bf16 [2*B*T, Kp] hi/lo operand with TIME-major rows (t*B + b)"""
    x = _req_strided(x, 'x')
    rows, Kd = x.shape
    kp = (Kd + 63) // 64 * 64
    out = torch.empty((2 * rows, kp), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().danet_split_operand_time_major(_p(x), x.stride(0), rows, Kd, int(T), _p(out), _stream()),
               'split_operand_time_major')
    _count()
    return out


def split_operand_paired(x, perm_T, shift=0, out=None, row0=0, rows_total=None):
    """batch-major activation x [B*T, rows] (row stride allowed) -> rows row0.. of a bf16 [2*rows_total, Kp] hi/lo operand
    whose reduction index is time-major (k = t*B + b), optionally shifted by one step (zero filled): the A operand of
    the weight-gradient products dW = sum_t x_t^T da_t / h_{t-+1}^T da_t"""
    x = _req_strided(x, 'x')
    Kd, rows = x.shape
    rows_total = rows if rows_total is None else rows_total
    kp = (Kd + 63) // 64 * 64
    if out is None:
        out = torch.empty((2 * rows_total, kp), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().danet_split_operand_paired(_p(x), x.stride(0), rows, Kd, int(perm_T), int(shift), _p(out),
                                                      row0, rows_total, _stream()), 'split_operand_paired')
    _count()
    return out


def gemm_split(a2, b2, m, n, k, bias=None, out=None, out_perm_T=0, accumulate=False, row_mu=None, col_s=None,
               rows_per_mu=1):
    """C[m,n] (+)= A2 @ B2^T (+ bias) (- row_mu[row // rows_per_mu] * col_s[n]) on operands already in the split
    bf16 layout"""
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a2.device)
    _lib.check(_lib.load().danet_gemm_split(_p(a2), _p(b2), _p(bias), _p(row_mu), _p(col_s), int(rows_per_mu), _p(out),
                                            out.stride(0), m, n, k, int(out_perm_T), int(accumulate), _stream()),
               'gemm_split')
    _count()
    return out


PROJ_ANCHOR_E = 20


def proj_anchor(a2, w2, B, T, F, E, k, anchors, row_mu=None, col_s=None, return_all=False):
    """
    Output projection with the anchor estimator's sums taken in its epilogue [app/modules.py:244-259 + 501-545, two
    sources]: split operands a2 [2*B*T, Kp], w2 [2*F*E, Kp] -> (embed [B,T,F,E], attractors [B,2,E]
    (+ sets [B,P,2,E], sims [B,P], choice [B] int32)).  E = 20 and at most 6 anchors; otherwise ValueError.
    """
    anchors = _req(anchors, 'anchors', dim=2)
    A = anchors.shape[0]
    if anchors.shape[1] != E:
        raise ValueError('proj_anchor: anchors %s, E = %d' % (tuple(anchors.shape), E))
    dev = a2.device
    P = A * (A - 1) // 2
    embed = torch.empty((B, T, F, E), dtype=torch.float32, device=dev)
    attrs = torch.empty((B, 2, E), dtype=torch.float32, device=dev)
    sets = torch.empty((B, P, 2, E), dtype=torch.float32, device=dev) if return_all else None
    sims = torch.empty((B, P), dtype=torch.float32, device=dev) if return_all else None
    choice = torch.empty((B,), dtype=torch.int32, device=dev) if return_all else None
    lib = _lib.load()
    ws = _ws(lib.danet_proj_anchor_workspace_bytes(B, T, F, E), dev)
    _lib.check(lib.danet_proj_anchor_fwd(_p(a2), _p(w2), _p(row_mu), _p(col_s), _p(anchors), _p(embed), _p(attrs),
                                         _p(sets), _p(sims), _p(choice), B, T, F, E, int(k), A, _p(ws), ws.numel(),
                                         _stream()), 'proj_anchor')
    _count(2)
    return (embed, attrs, sets, sims, choice) if return_all else (embed, attrs)


def _embed_flat(embed):
    embed = _req(embed, 'embed')
    if embed.dim() == 4:
        B, T, F, E = embed.shape
        return embed, B, T * F, E
    if embed.dim() == 3:
        B, TF, E = embed.shape
        return embed, B, TF, E
    raise ValueError('embed: expected [B,T,F,E] or [B,T*F,E], got %s' % (tuple(embed.shape),))


TRUTH_MODES = {'truth': 0, 'truth-threshold': 1, 'truth-weighted': 2}


def attractor_truth(embed, src_pwr, mix_pwr, mode, return_den=False):
    """[app/modules.py:390-412 / 425-450 / 462-487] -> [B,C,E] (+ den [B,C], the per-class weight sums)"""
    embed, B, TF, E = _embed_flat(embed)
    src_pwr = _req(src_pwr, 'src_pwr', dim=4)
    Cn = src_pwr.shape[1]
    if src_pwr.shape[0] != B or src_pwr.shape[2] * src_pwr.shape[3] != TF:
        raise ValueError('attractor_truth: src_pwr %s does not match embed' % (tuple(src_pwr.shape),))
    m = TRUTH_MODES[mode] if isinstance(mode, str) else int(mode)
    if m != 0:
        mix_pwr = _req(mix_pwr, 'mix_pwr', dim=3)
        if mix_pwr.shape[0] != B or mix_pwr.shape[1] * mix_pwr.shape[2] != TF:
            raise ValueError('attractor_truth: mix_pwr %s does not match embed' % (tuple(mix_pwr.shape),))
    else:
        mix_pwr = None
    out = torch.empty((B, Cn, E), dtype=torch.float32, device=embed.device)
    den = torch.empty((B, Cn), dtype=torch.float32, device=embed.device) if return_den else None
    lib = _lib.load()
    ws = _ws(lib.danet_attractor_workspace_bytes(B, Cn, E), embed.device)
    _lib.check(lib.danet_attractor_truth_fwd(_p(embed), _p(src_pwr), _p(mix_pwr), _p(out), _p(den), B, Cn, TF, E, m,
                                             _p(ws), ws.numel(), _stream()), 'attractor_truth')
    _count(2)
    return (out, den) if return_den else out


def attractor_anchor(embed, anchors, n_signal, return_all=False, return_den=False):
    """[app/modules.py:501-545] -> [B,C,E] (+ sets [B,P,C,E], sims [B,P], choice [B] int32 (+ den [B,P,C]))"""
    embed, B, TF, E = _embed_flat(embed)
    anchors = _req(anchors, 'anchors', dim=2)
    A = anchors.shape[0]
    if anchors.shape[1] != E:
        raise ValueError('attractor_anchor: anchors %s, E = %d' % (tuple(anchors.shape), E))
    lib = _lib.load()
    P = lib.danet_anchor_num_subsets(A, n_signal)
    dev = embed.device
    out = torch.empty((B, n_signal, E), dtype=torch.float32, device=dev)
    return_all = return_all or return_den
    sets = torch.empty((B, P, n_signal, E), dtype=torch.float32, device=dev) if return_all else None
    sims = torch.empty((B, P), dtype=torch.float32, device=dev) if return_all else None
    choice = torch.empty((B,), dtype=torch.int32, device=dev) if return_all else None
    den = torch.empty((B, P, n_signal), dtype=torch.float32, device=dev) if return_den else None
    ws = _ws(lib.danet_attractor_workspace_bytes(B, max(P, 1) * n_signal, E), dev)
    _lib.check(lib.danet_attractor_anchor_fwd(_p(embed), _p(anchors), _p(out), _p(sets), _p(sims), _p(choice),
                                              _p(den), B, n_signal, TF, E, A, _p(ws), ws.numel(), _stream()),
               'attractor_anchor')
    _count(2)
    if return_den:
        return out, sets, sims, choice, den
    return (out, sets, sims, choice) if return_all else out


def attractor_kmeans(embed, init, n_iter=5):
    """Lloyd iterations from `init` [B,C,E] (new plugin, no reference twin) -> [B,C,E]"""
    embed, B, TF, E = _embed_flat(embed)
    cen = _req(init, 'init', dim=3).clone()
    if cen.shape[0] != B or cen.shape[2] != E:
        raise ValueError('attractor_kmeans: init %s does not match embed' % (tuple(cen.shape),))
    Cn = cen.shape[1]
    lib = _lib.load()
    ws = _ws(lib.danet_attractor_workspace_bytes(B, Cn, E), embed.device)
    _lib.check(lib.danet_attractor_kmeans_fwd(_p(embed), _p(cen), B, Cn, TF, E, int(n_iter), _p(ws),
                                              ws.numel(), _stream()), 'attractor_kmeans')
    _count(2 * int(n_iter))
    return cen


SEPARATOR_KINDS = {'dot-softmax-orig': 0, 'dot-sigmoid-orig': 1}


def mask_cmul(embed, attractors, mix, kind, want=('sep_pwr', 'sep', 'masks'), mix_pwr=None):
    """
    [app/modules.py:548-603; main.py:281-284] embed [B,TF,E], attractors [B,C,E], mix c64 [B,T,F]
    and/or mix_pwr f32 [B,T,F] -> dict(sep_pwr [B,C,T,F], sep c64 [B,C,T,F], masks [B,T,F,C])
    """
    embed, B, TF, E = _embed_flat(embed)
    attractors = _req(attractors, 'attractors', dim=3)
    if mix is None and mix_pwr is None:
        raise ValueError('mask_cmul: need the complex mixture or its magnitude')
    if mix is not None:
        mix = _req(mix, 'mix', torch.complex64, 3)
    if mix_pwr is not None:
        mix_pwr = _req(mix_pwr, 'mix_pwr', dim=3)
    if mix is None:
        want = tuple(w for w in want if w != 'sep')
    Cn = attractors.shape[1]
    if attractors.shape[0] != B or attractors.shape[2] != E:
        raise ValueError('mask_cmul: attractors %s do not match embed' % (tuple(attractors.shape),))
    for nm, m in (('mix', mix), ('mix_pwr', mix_pwr)):
        if m is not None and (m.shape[0] != B or m.shape[1] * m.shape[2] != TF):
            raise ValueError('mask_cmul: %s %s does not match embed' % (nm, tuple(m.shape)))
    T, F = (mix if mix is not None else mix_pwr).shape[1:]
    dev = embed.device
    k = SEPARATOR_KINDS[kind] if isinstance(kind, str) else int(kind)
    out = {
        'sep_pwr': torch.empty((B, Cn, T, F), dtype=torch.float32, device=dev) if 'sep_pwr' in want else None,
        'sep': torch.empty((B, Cn, T, F), dtype=torch.complex64, device=dev) if 'sep' in want else None,
        'masks': torch.empty((B, T, F, Cn), dtype=torch.float32, device=dev) if 'masks' in want else None,
    }
    _lib.check(_lib.load().danet_mask_cmul_fwd(_p(embed), _p(attractors), _p(mix), _p(mix_pwr), _p(out['sep_pwr']),
                                               _p(out['sep']), _p(out['masks']), B, Cn, TF, E, k, _stream()),
               'mask_cmul')
    _count()
    return out


def conv2d(x, w_hwio, bias=None, leak=-1.):
    """tf.layers.conv2d(channels_first, padding='same') + bias + leaky relu max(leak*v, v) (leak < 0: none)
    [app/modules.py:289-369; app/ops.py:103-106].  x [B,Cin,H,W]; w in TensorFlow's [k,k,Cin,Cout] layout."""
    x = _req(x, 'x', dim=4)
    w_hwio = _req(w_hwio, 'w', dim=4)
    B, Cin, H, W = x.shape
    k, k2, ci, Cout = w_hwio.shape
    if k != k2 or ci != Cin:
        raise ValueError('conv2d: kernel %s does not match input %s' % (tuple(w_hwio.shape), tuple(x.shape)))
    if bias is not None:
        bias = _req(bias, 'bias', dim=1)
    y = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().danet_conv2d_fwd(_p(x), _p(w_hwio), _p(bias), _p(y), B, Cin, Cout, H, W, k, float(leak),
                                            _stream()), 'conv2d')
    _count()
    return y


def maxpool2x2(x):
    """tf.layers.max_pooling2d((2,2),(2,2), channels_first) [app/modules.py:299-300, 312-313]: [B,C,H,W] -> [B,C,H/2,W/2]"""
    x = _req(x, 'x', dim=4)
    B, Cn, H, W = x.shape
    if H < 2 or W < 2:
        raise ValueError('maxpool2x2: input %s is smaller than the window' % (tuple(x.shape),))
    y = torch.empty((B, Cn, H // 2, W // 2), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().danet_maxpool2x2_fwd(_p(x), _p(y), B * Cn, H, W, _stream()), 'maxpool2x2')
    _count()
    return y


def conv2d_bwd(x, w_hwio, y, dy, leak=-1., need_dx=True):
    """backward of conv2d (+ leaky relu when leak >= 0: `y` is the layer's output) -> (dx | None, dw [k,k,Cin,Cout], dbias)
    [TF autodiff of app/modules.py:289-369 at main.py:357-358]"""
    x, w_hwio, dy = _req(x, 'x', dim=4), _req(w_hwio, 'w', dim=4), _req(dy, 'dy', dim=4)
    B, Cin, H, W = x.shape
    k, _, _, Cout = w_hwio.shape
    if tuple(dy.shape) != (B, Cout, H, W):
        raise ValueError('conv2d_bwd: dy %s does not match output %s' % (tuple(dy.shape), (B, Cout, H, W)))
    if leak >= 0.:
        dy = leaky_relu_bwd(_req(y, 'y', dim=4), dy.clone(), leak)
    lib = _lib.load()
    dw = torch.empty_like(w_hwio)
    db = torch.empty((Cout,), dtype=torch.float32, device=x.device)
    _lib.check(lib.danet_conv2d_bwd_weights(_p(x), _p(dy), _p(dw), _p(db), B, Cin, Cout, H, W, k, _stream()), 'conv2d_bwd_weights')
    _count()
    dx = None
    if need_dx:
        dx = torch.empty_like(x)
        _lib.check(lib.danet_conv2d_bwd_data(_p(dy), _p(w_hwio), _p(dx), B, Cin, Cout, H, W, k, _stream()), 'conv2d_bwd_data')
        _count()
    return dx, dw, db


def maxpool2x2_bwd(x, dy):
    """gradient of maxpool2x2 w.r.t. its input x [B,C,H,W] given dy [B,C,H/2,W/2]"""
    x, dy = _req(x, 'x', dim=4), _req(dy, 'dy', dim=4)
    B, Cn, H, W = x.shape
    if tuple(dy.shape) != (B, Cn, H // 2, W // 2):
        raise ValueError('maxpool2x2_bwd: dy %s does not match %s' % (tuple(dy.shape), (B, Cn, H // 2, W // 2)))
    dx = torch.empty_like(x)
    _lib.check(_lib.load().danet_maxpool2x2_bwd(_p(x), _p(dy), _p(dx), B * Cn, H, W, _stream()), 'maxpool2x2_bwd')
    _count()
    return dx


def add(a, b):
    """a + b (the residual connection at app/modules.py:335)"""
    a = _req(a, 'a')
    b = _req(b, 'b')
    if a.shape != b.shape:
        raise ValueError('add: shapes %s and %s differ' % (tuple(a.shape), tuple(b.shape)))
    out = torch.empty_like(a)
    _lib.check(_lib.load().danet_add_fwd(_p(a), _p(b), _p(out), a.numel(), _stream()), 'add')
    _count()
    return out


def mask_cmul_istft(embed, attractors, mix, kind, out=None):
    """
    K4 in one launch [app/modules.py:548-603; main.py:281-284; app/utils.py:53-75]: embed [B,TF,E], attractors [B,C,E],
    mix c64 [B,T,129] -> separated waveforms f32 [B,C,64*T]; the separated spectra are never written to HBM.
    Equals istft(mask_cmul(...)['sep']).
    """
    embed, B, TF, E = _embed_flat(embed)
    attractors = _req(attractors, 'attractors', dim=3)
    mix = _req(mix, 'mix', torch.complex64, 3)
    Cn = attractors.shape[1]
    T = mix.shape[1]
    if attractors.shape[0] != B or attractors.shape[2] != E:
        raise ValueError('mask_cmul_istft: attractors %s do not match embed' % (tuple(attractors.shape),))
    if mix.shape[0] != B or mix.shape[2] != FEATURE or T * FEATURE != TF:
        raise ValueError('mask_cmul_istft: mix %s does not match embed' % (tuple(mix.shape),))
    if out is None:
        out = torch.empty((B, Cn, FFT_STRIDE * T), dtype=torch.float32, device=embed.device)
    elif tuple(out.shape) != (B, Cn, FFT_STRIDE * T) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError('mask_cmul_istft: out must be a contiguous float32 %s' % ((B, Cn, FFT_STRIDE * T),))
    k = SEPARATOR_KINDS[kind] if isinstance(kind, str) else int(kind)
    _lib.check(_lib.load().danet_mask_cmul_istft_fwd(_p(embed), _p(attractors), _p(mix), _p(out), B, Cn, T, E, k,
                                                     _stream()), 'mask_cmul_istft')
    _count()
    return out


FUSED_K4_MAX_C, FUSED_K4_MAX_E = 4, 64


def pit_mse(x, y):
    """
    [app/ops.py:374-431, 191-222; main.py:293-309]  x, y [B,C,T,F] both complex64 or both float32
    -> dict(loss [1], perm_idx [B] int32, perm_losses [B,C!], cross [B,C,C], snr [B])
    """
    cplx = x.dtype == torch.complex64
    x = _req(x, 'x', torch.complex64 if cplx else torch.float32, 4)
    y = _req(y, 'y', torch.complex64 if cplx else torch.float32, 4)
    if x.shape != y.shape:
        raise ValueError('pit_mse: shapes differ %s vs %s' % (tuple(x.shape), tuple(y.shape)))
    B, Cn, T, F = x.shape
    nperm = 1
    for c in range(2, Cn + 1):
        nperm *= c
    dev = x.device
    out = {
        'cross': torch.empty((B, Cn, Cn), dtype=torch.float32, device=dev),
        'perm_losses': torch.empty((B, nperm), dtype=torch.float32, device=dev),
        'perm_idx': torch.empty((B,), dtype=torch.int32, device=dev),
        'loss': torch.empty((1,), dtype=torch.float32, device=dev),
        'snr': torch.empty((B,), dtype=torch.float32, device=dev),
    }
    lib = _lib.load()
    ws = _ws(lib.danet_pit_workspace_bytes(B, Cn), dev)
    _lib.check(lib.danet_pit_mse_fwd(_p(x), _p(y), B, Cn, T * F, 1 if cplx else 0, _p(out['cross']),
                                     _p(out['perm_losses']), _p(out['perm_idx']), _p(out['loss']),
                                     _p(out['snr']), _p(ws), ws.numel(), _stream()), 'pit_mse')
    _count(2)
    return out


def head_bwd(embed, attractors, mix, src, perm_idx, kind, est_mode, src_pwr=None, mix_pwr=None, anchors=None,
             choice=None, den=None):
    """
    Backward of loss = pit_mse(src, mask*mix) through the separator and the estimator
    (TF autodiff at main.py:357-358) -> dict(d_embed [B,TF,E], d_attractors [B,C,E], d_anchors [A,E] | None).
    est_mode: 'truth' | 'truth-threshold' | 'truth-weighted' | 'anchor'.
    """
    embed, B, TF, E = _embed_flat(embed)
    attractors = _req(attractors, 'attractors', dim=3)
    mix = _req(mix, 'mix', torch.complex64, 3)
    src = _req(src, 'src', torch.complex64, 4)
    perm_idx = _req(perm_idx, 'perm_idx', torch.int32, 1)
    den = _req(den, 'den')
    Cn = attractors.shape[1]
    k = SEPARATOR_KINDS[kind] if isinstance(kind, str) else int(kind)
    mode = 3 if est_mode == 'anchor' else TRUTH_MODES[est_mode]
    dev = embed.device
    lib = _lib.load()
    ws = _ws(lib.danet_head_bwd_workspace_bytes(B, Cn, E), dev)
    d_attr = torch.empty((B, Cn, E), dtype=torch.float32, device=dev)
    _lib.check(lib.danet_head_bwd_attractors(_p(embed), _p(attractors), _p(mix), _p(src), _p(perm_idx), _p(d_attr),
                                             B, Cn, TF, E, k, _p(ws), ws.numel(), _stream()), 'head_bwd_attractors')
    d_embed = torch.empty((B, TF, E), dtype=torch.float32, device=dev)
    d_anchors, n_anchor = None, 0
    if mode == 3:
        anchors = _req(anchors, 'anchors', dim=2)
        choice = _req(choice, 'choice', torch.int32, 1)
        n_anchor = anchors.shape[0]
        d_anchors = torch.empty_like(anchors)
    else:
        src_pwr = _req(src_pwr, 'src_pwr', dim=4)
        if mode != 0:
            mix_pwr = _req(mix_pwr, 'mix_pwr', dim=3)
    _lib.check(lib.danet_head_bwd_embed(_p(embed), _p(attractors), _p(mix), _p(src), _p(perm_idx), _p(d_attr), mode,
                                        _p(src_pwr), _p(mix_pwr), _p(anchors), _p(choice), _p(den), _p(d_embed),
                                        _p(d_anchors), B, Cn, TF, E, k, n_anchor, _p(ws), ws.numel(), _stream()),
               'head_bwd_embed')
    _count(4)
    return {'d_embed': d_embed, 'd_attractors': d_attr, 'd_anchors': d_anchors}


def lstm_seq_bwd(d_out, gates, cell, w_list, in_dim, T, B, H, backend=None):
    """
    Backward through time [TF autodiff of main.py:125-131]: d_out [B,T,n_dir*H], gates [n_dir,T,B,4H]
    (post-activation, from lstm_seq(keep_gates=True)) are overwritten IN PLACE with the pre-activation
    gradients da; returns `gates` (now da).  backend 2: the tcgen05 kernel for wide layers (TC_LSTM_MAX_H < H <=
    TC_WIDE_MAX_H, recurrent weights as one fp16 value in the product); None: 1 up to TC_LSTM_MAX_H, else the exact fp32 one.
    """
    d_out = _req(d_out, 'd_out', dim=3)
    gates = _req(gates, 'gates', dim=4)
    cell = _req(cell, 'cell', dim=4)
    n_dir = len(w_list)
    if tuple(gates.shape) != (n_dir, T, B, 4 * H) or tuple(cell.shape) != (n_dir, T, B, H) or \
            tuple(d_out.shape) != (B, T, n_dir * H):
        raise ValueError('lstm_seq_bwd: shapes %s %s %s' % (tuple(d_out.shape), tuple(gates.shape), tuple(cell.shape)))
    ptrs = (C.c_void_p * n_dir)()
    for d, w in enumerate(w_list):
        w = _req(w, 'W[%d]' % d, dim=2)
        ptrs[d] = w.data_ptr() + in_dim * 4 * H * 4
    lib = _lib.load()
    ws = _ws(lib.danet_lstm_seq_bwd_workspace_bytes(n_dir, B, H), gates.device)
    be = (DEFAULT_BACKEND if H <= TC_LSTM_MAX_H else 0) if backend is None else backend
    _lib.check(lib.danet_lstm_seq_bwd(_p(d_out), _p(gates), _p(cell), ptrs, 4 * H, n_dir, T, B, H, _p(ws), ws.numel(),
                                      be, _stream()), 'lstm_seq_bwd')
    _count(2)
    return gates


def colsum(x, out=None, accumulate=False):
    """column sums of a 2-D tensor (bias gradients)"""
    x = _req_strided(x, 'x')
    rows, n = x.shape
    if out is None:
        out = torch.empty((n,), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    ws = _ws(lib.danet_colsum_workspace_bytes(n), x.device)
    _lib.check(lib.danet_colsum(_p(x), x.stride(0), rows, n, _p(out), int(accumulate), _p(ws), ws.numel(), _stream()),
               'colsum')
    _count(2)
    return out


def clip_adam(param, grad, m, v, step, lr=3e-4, clip=100., beta1=.9, beta2=.999, eps=1e-8, grad_scale=1.):
    """in place: clip_by_value then Adam [main.py:359-363; app/ozers.py:15-18]"""
    for t, nm in ((param, 'param'), (grad, 'grad'), (m, 'm'), (v, 'v')):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError('clip_adam: %s must be a contiguous float32 CUDA tensor' % nm)
        if t.numel() != param.numel():
            raise ValueError('clip_adam: %s has %d elements, param %d' % (nm, t.numel(), param.numel()))
    _lib.check(_lib.load().danet_clip_adam(_p(param), _p(grad), _p(m), _p(v), param.numel(), grad_scale,
                                           clip if clip is not None else 0., lr, beta1, beta2, eps, int(step),
                                           _stream()), 'clip_adam')
    _count()


def clip_sgd(param, grad, lr, clip=100., grad_scale=1.):
    """in place: clip_by_value then gradient descent [app/ozers.py:9-12]"""
    _lib.check(_lib.load().danet_clip_sgd(_p(param), _p(grad), param.numel(), grad_scale,
                                          clip if clip is not None else 0., lr, _stream()), 'clip_sgd')
    _count()
