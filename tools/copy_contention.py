"""Why does end-to-end scaling dip at N = 8 (SCALE_r01: 0.943)?  Every rank of the box moves one step's worth of pinned
host traffic (4.1 MB in, 8.2 MB out, as bench.py's e2e step) -- first ONE RANK AT A TIME, then ALL RANKS TOGETHER after a
barrier -- and reports the per-rank copy time.  Run under torchrun (--nproc-per-node N).  If the together-times grow with N
the host side (root complex / memory / IOMMU of the VM) is shared; if not, the dip is elsewhere (launch path, CPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import danet_tensorflow_b200 as D
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
bind = os.environ.get('BIND', '1') == '1'
numa = D.shard.bind_to_local_numa(local) if bind else None
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
hin = torch.empty(32 * 32000, dtype=torch.float32).pin_memory()
hout = torch.empty(32 * 2 * 64 * 501, dtype=torch.float32).pin_memory()
din, dout = torch.empty_like(hin, device='cuda'), torch.empty_like(hout, device='cuda')


def once():
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); din.copy_(hin, non_blocking=True); e1.record(); hout.copy_(dout, non_blocking=True); e2.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(5):
    once()
alone = None
for r in range(world):
    barrier()
    if r == rank:
        alone = np.median([once() for _ in range(20)], axis=0)
barrier()
together = []
for _ in range(20):
    barrier()
    together.append(once())
together = np.median(together, axis=0)
res = torch.tensor([alone[0], alone[1], together[0], together[1]], dtype=torch.float64, device='cuda')
allr = [torch.zeros_like(res) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    print('N = %d ranks, affinity binding %s, cpus of rank 0: %d, numa node %s' % (world, bind, len(os.sched_getaffinity(0)), numa))
    print('rank   H2D alone  D2H alone | H2D together  D2H together   (us; 4.1 MB in, 8.2 MB out)')
    for r, t in enumerate(allr):
        a = t.cpu().numpy()
        print('%4d   %8.1f  %9.1f | %11.1f  %12.1f' % (r, a[0], a[1], a[2], a[3]))
    m = np.stack([t.cpu().numpy() for t in allr])
    print('mean   %8.1f  %9.1f | %11.1f  %12.1f   -> together / alone: H2D %.2fx, D2H %.2fx' % (
        m[:, 0].mean(), m[:, 1].mean(), m[:, 2].mean(), m[:, 3].mean(), m[:, 2].mean() / m[:, 0].mean(), m[:, 3].mean() / m[:, 1].mean()))
if world > 1:
    dist.destroy_process_group()
