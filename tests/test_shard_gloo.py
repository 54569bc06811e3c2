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


def _grad_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import danet_tensorflow_b200 as D
    model = D.Model('ddp', device='cpu')
    # the flat gradient buffer of a 2-variable model; rank r holds gradient (r + 1) everywhere
    flat = torch.full((128,), float(rank + 1))
    model._flat = dict(grad=flat)
    scale = model.all_reduce_grads()
    out[rank] = (flat.numpy().copy(), scale)
    dist.destroy_process_group()


def test_gradient_all_reduce_is_a_mean_over_ranks():
    """the one collective of the training step: sum of the flat gradient buffer, 1/world folded into Adam"""
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_grad_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        g, scale = out[r]
        assert np.all(g == 3.) and scale == 0.5          # (1 + 2) summed; mean = 1.5 after the scale
