"""
Two-process NCCL checks of the training step's one collective, run under `-m gpu` when the box has at
least two devices (skipped otherwise): rank r computes the gradient of ITS shard with the CUDA path, the bucketed
all-reduce (shard.GradientBuckets, started per layer from the backward pass) sums them, and every rank must then hold
the full-batch gradient and, after clip + Adam, identical parameters.
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import danet_tensorflow_b200 as D
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
K = D.kernels
D.hparams.load(dict(ENCODER_TYPE='bilstm-orig', TRAIN_ESTIMATOR_METHOD='anchor', INFER_ESTIMATOR_METHOD='anchor',
                    SEPARATOR_TYPE='dot-softmax-orig', BATCH_SIZE=8))
D.hparams.digest()
g = torch.Generator(device='cuda').manual_seed(11)          # same seed on every rank: the same global batch
src = K.stft(torch.randn(8, 2, 4000, device='cuda', generator=g) * 1000.)
lo, hi = D.shard.shard_bounds(8, rank, world)
full = D.Model('full', 'cuda', seed=1337).build()
full.BUCKETED_ALLREDUCE = False
full.train_forward_backward(src)                            # full batch, no exchange
gfull = full._flat['grad'].clone()
full.apply_gradients(1.)
m = D.Model('shard', 'cuda', seed=1337).build()
for bucketed in (True, False):
    m.BUCKETED_ALLREDUCE = bucketed
    m.train_forward_backward(src[lo:hi].contiguous())
    scale = m.all_reduce_grads()
    assert scale == 1. / world
    mean = m._flat['grad'] * scale
    err = float((mean - gfull).abs().max()) / float(gfull.abs().max())
    assert err < 5e-5, (bucketed, err)          # bf16x3 products, shard vs full batch tiling
m.apply_gradients(scale)
# Adam's first step moves every weight by ~lr * sign(g): where the gradient is numerically zero the sign is noise, so the
# parameters are compared where the gradient is significant (and bounded by 2 lr everywhere)
diff = (m._flat['param'] - full._flat['param']).abs()
sig = gfull.abs() > 1e-3 * gfull.abs().max()
perr = float(diff[sig].max())
assert perr < 2e-6, perr
assert float(diff.max()) <= 2.1 * D.hparams.LR
# every rank ends with bit-identical parameters (the exchanged gradient is the same everywhere)
ref = m._flat['param'].clone()
dist.broadcast(ref, 0)
assert torch.equal(ref, m._flat['param'])
if rank == 0:
    print('MULTI_OK grad err %%.2e param err %%.2e' %% (err, perr))
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_rank_nccl_sharded_training_step(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % dict(root=ROOT))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(_free_port()), str(script)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode()
    assert r.returncode == 0 and 'MULTI_OK' in out, out[-3000:]
