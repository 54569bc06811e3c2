"""world_size-2 gloo test (CPU) of the N>1 host logic: balanced utterance sharding, max-over-ranks
timing, and rank-ordered gather -- the only cross-rank operations of the inference path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import danet_tensorflow_b200 as D
    lo, hi = D.shard.shard_bounds(n_total, rank, world)
    batch = torch.arange(n_total * 3, dtype=torch.float32).reshape(n_total, 3)
    local = batch[lo:hi] * 2.                      # stand-in for "separate my shard"
    full = D.shard.gather_rows(local, n_total)
    t = D.shard.max_over_ranks(10. + rank)
    out[rank] = (lo, hi, full.numpy().copy(), t)
    dist.destroy_process_group()


def test_two_rank_sharding():
    world, n_total = 2, 7
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    assert (out[0][0], out[0][1]) == (0, 4) and (out[1][0], out[1][1]) == (4, 7)
    ref = np.arange(n_total * 3, dtype=np.float32).reshape(n_total, 3) * 2.
    for r in range(world):
        assert np.array_equal(out[r][2], ref)
        assert out[r][3] == 11.


def test_shard_bounds_cover_everything():
    import danet_tensorflow_b200 as D
    for n in (0, 1, 7, 32, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [D.shard.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _toy_problem(seed=5, B=4, C=2, T=6, F=129, E=20):
    """a tiny full batch of the real training graph: complex spectra + reference-initialised weights (2 layers, fp64)"""
    from oracle import danet_oracle as O
    rs = np.random.RandomState(seed)
    src = torch.from_numpy((rs.standard_normal((B, C, T, F)) + 1j * rs.standard_normal((B, C, T, F))) * 30.)
    P = O.reference_init(1337, embed=E, estimators=('train_estimator',), dtype=torch.float64, n_layers=2, hdim=16)
    return src, P


def _oracle_grads(src, P):
    """d(train loss)/d(every variable) by torch autograd on the oracle (main.py:289, 357-358), in P's order"""
    from oracle import danet_oracle as O
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    x = torch.log1p(src.sum(1).abs())
    V = O.encoder_bilstm(x, Pg, 20, n_layers=2)
    A = O.estimator_anchor(V, Pg['train_estimator/anchors'], src.shape[1])
    mix = src.sum(1)
    sep_pwr = O.separator(mix.abs(), A, V.reshape(V.shape[0], -1, 20), 'dot-softmax-orig')
    ph = torch.atan2(mix.imag, mix.real).unsqueeze(1)
    sep = torch.complex(torch.cos(ph) * sep_pwr, torch.sin(ph) * sep_pwr)
    loss = O.pit_mse_loss(src, sep)[0]
    loss.backward()
    return {k: v.grad for k, v in Pg.items()}


def _grad_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import danet_tensorflow_b200 as D
    src, P = _toy_problem()
    lo, hi = D.shard.shard_bounds(src.shape[0], rank, world)
    g = _oracle_grads(src[lo:hi], P)                      # stand-in for the CUDA backward of this rank's shard
    names = list(P)
    offs, total = D.shard.flat_layout({k: P[k].numel() for k in names})
    flat = torch.zeros(total, dtype=torch.float64)
    for k in names:
        flat[offs[k]:offs[k] + P[k].numel()] = g[k].reshape(-1)
    buckets = D.shard.GradientBuckets(flat, {k: (offs[k], offs[k] + P[k].numel()) for k in names})
    # the order Model / _RecurrentEncoder.backward uses: projection + anchors, then the layers top down;
    # layer 0 is deliberately left to finish() (an encoder without bucket hooks must still be exchanged completely)
    buckets.reduce(['encoder/output/W', 'train_estimator/anchors'])
    plan1 = buckets.plan(['encoder/lstm1_%s/LSTM/linear/%s' % (d, v) for d in ('fwd', 'bwd') for v in 'WB'])
    buckets.reduce(['encoder/lstm1_%s/LSTM/linear/%s' % (d, v) for d in ('fwd', 'bwd') for v in 'WB'])
    scale = buckets.finish()
    out[rank] = (flat.numpy().copy(), scale, plan1, offs, total)
    dist.destroy_process_group()


def test_sharded_gradients_equal_the_full_batch_gradient():
    """N-way sharded gradients, summed by the bucketed all-reduce and scaled by 1/N, equal the
    single-process full-batch gradient (1e-5; here fp64 so ~1e-12), and the clip is applied AFTER the mean
    (main.py:358-363 clips the full-batch gradient)"""
    from oracle import danet_oracle as O
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_grad_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    src, P = _toy_problem()
    full = _oracle_grads(src, P)
    flat0, scale0, plan1, offs, total = out[0]
    assert scale0 == 0.5 and np.array_equal(out[0][0], out[1][0])            # every rank holds the same sum
    assert len(plan1) == 1                                                   # one layer = ONE contiguous slice
    mean = {k: torch.from_numpy(flat0[offs[k]:offs[k] + P[k].numel()] * scale0).reshape(P[k].shape) for k in P}
    for k in P:
        denom = float(full[k].abs().max()) + 1e-30
        assert float((mean[k] - full[k]).abs().max()) / denom < 1e-5, k
    # clip after the mean: one Adam step from the exchanged gradient == one step from the full-batch gradient,
    # with a threshold low enough to bite, and != clipping each shard's gradient before the exchange
    clip = float(np.median([float(v.abs().max()) for v in full.values()])) * 0.25
    def step(grads):
        p = {k: v.clone() for k, v in P.items()}
        m = {k: torch.zeros_like(v) for k, v in P.items()}
        v2 = {k: torch.zeros_like(v) for k, v in P.items()}
        O.clip_adam_step(p, grads, m, v2, 1, clip=clip)
        return p, m
    p_mean, m_mean = step(mean)
    p_full, m_full = step(full)
    for k in P:
        assert float((m_mean[k] - m_full[k]).abs().max()) <= 1e-9 * (float(m_full[k].abs().max()) + 1e-30), k
        assert float((p_mean[k] - p_full[k]).abs().max()) < 1e-9, k
    assert any(float(v.abs().max()) > clip for v in full.values())


def _model_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import danet_tensorflow_b200 as D
    model = D.Model('ddp', device='cpu')
    # Model.all_reduce_grads on the flat gradient buffer of a 2-variable model; rank r holds gradient (r + 1) everywhere
    flat = torch.full((128,), float(rank + 1))
    model._flat = dict(grad=flat)
    model._buckets = D.shard.GradientBuckets(flat, {'a': (0, 64), 'b': (64, 128)})
    model.grads_ready(['b'])
    scale = model.all_reduce_grads()
    out[rank] = (flat.numpy().copy(), scale)
    dist.destroy_process_group()


def test_gradient_all_reduce_is_a_mean_over_ranks():
    """Model.grads_ready / all_reduce_grads: bucket `b` goes early, `a` at the end; sum everywhere, 1/world for Adam"""
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_model_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        g, scale = out[r]
        assert np.all(g == 3.) and scale == 0.5          # (1 + 2) summed; mean = 1.5 after the scale
