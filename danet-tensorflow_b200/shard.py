"""
Utterance sharding for N GPUs: every stage of the separation path is per utterance,
so ranks split the batch axis and exchange NOTHING on the data path.  The only cross-rank traffic is
bookkeeping: the max-over-ranks step time and the gather of per-rank results on request.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced split: the first n_items % world ranks carry one extra utterance."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('rank %d of world %d' % (rank, world))
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device='cpu'):
    """max of a python float over all ranks (device-side for NCCL, host for gloo)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local, n_total, device=None):
    """Concatenate per-rank row blocks (sized by shard_bounds) on every rank, in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def flat_layout(numels, align=64):
    """offsets of variables packed into one flat buffer in the given (creation) order, each view aligned to `align`
    elements (256 bytes in fp32) -> ({name: offset}, total elements)"""
    offs, total = {}, 0
    for k, n in numels.items():
        offs[k] = total
        total += (int(n) + align - 1) // align * align
    return offs, total


class GradientBuckets(object):
    """
    The training step's one collective (main.py:357-363 computes the gradient of the FULL batch, so the
    shards' gradients are summed and scaled by 1/world before the clip): every gradient lives in ONE flat buffer, and a
    bucket is a contiguous slice of it.  `reduce(names)` is called by the backward pass as soon as a layer's gradients
    have been queued -- on the stream that produces them -- and starts an asynchronous all-reduce of the slice(s) that
    cover those variables, so the exchange of layer l runs under the backward recurrence of layer l-1.  `finish()`
    reduces whatever no bucket covered and makes the current stream wait for all of it.
    Works on any backend (NCCL on the GPUs, gloo in the CPU tests).
    """

    def __init__(self, flat, spans, pad=64):
        self.flat = flat
        self.spans = dict(spans)              # variable name -> (lo, hi) element range of the flat buffer
        self.pad = int(pad)                   # views are aligned to this many elements: smaller gaps are padding
        self._done, self._works = [], []

    @staticmethod
    def active():
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _merge(self, spans):
        out = []
        for lo, hi in sorted(spans):
            if out and lo - out[-1][1] < self.pad:
                out[-1][1] = max(out[-1][1], hi)
            else:
                out.append([lo, hi])
        return [(lo, hi) for lo, hi in out]

    def plan(self, names):
        """the contiguous slices an all-reduce of `names` touches (exposed for the tests)"""
        return self._merge(self.spans[n] for n in names if n in self.spans)

    def reduce(self, names):
        if not self.active():
            return
        for lo, hi in self.plan(names):
            self._works.append(dist.all_reduce(self.flat[lo:hi], async_op=True))
            self._done.append((lo, hi))

    def finish(self):
        """-> the scale that turns the summed gradient into the full-batch mean (1 / world)"""
        if not self.active():
            self._done, self._works = [], []
            return 1.
        covered, pos = self._merge(self._done), 0
        for lo, hi in covered + [(self.flat.numel(), self.flat.numel())]:
            if lo - pos >= 1 and any(s < lo and e > pos for s, e in self.spans.values()):
                self._works.append(dist.all_reduce(self.flat[pos:lo], async_op=True))
            pos = max(pos, hi)
        for w in self._works:
            w.wait()
        self._done, self._works = [], []
        return 1. / dist.get_world_size()


def _cpulist(text):
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_local_numa(local_rank, world_local=None):
    """One process per GPU: run this rank's host threads -- and therefore first-touch its pinned staging buffers -- on the
    GPU's own NUMA node (sysfs `numa_node` of the device's PCI function), so N ranks do not all stage through node 0.
    When the platform reports no locality (one node, or -1 as on most virtual machines) the ranks of the box still get
    disjoint core ranges, which keeps eight copy-issuing host threads off each other's cores.  Returns the node (or
    None) for the record; never raises -- affinity is an optimisation, not a requirement."""
    import os
    try:
        ncpu = os.cpu_count() or 1
        world_local = world_local or int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1')))
        node = None
        if torch.cuda.is_available():
            bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
            dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
            path = '/sys/bus/pci/devices/%04x:%02x:00.0/numa_node' % (dom, bus)
            if os.path.exists(path):
                n = int(open(path).read().strip())
                if n >= 0 and os.path.exists('/sys/devices/system/node/node%d/cpulist' % n):
                    node = n
        allowed = sorted(os.sched_getaffinity(0))
        if node is not None:
            cpus = [c for c in _cpulist(open('/sys/devices/system/node/node%d/cpulist' % node).read()) if c in allowed]
            peers = max(1, world_local // max(1, len([d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')])))
        else:
            cpus, peers = allowed, world_local
        if world_local > 1 and len(cpus) >= 2 * peers:
            k = len(cpus) // peers
            j = local_rank % peers
            cpus = cpus[j * k:(j + 1) * k]
        if cpus:
            os.sched_setaffinity(0, cpus)
            torch.set_num_threads(max(1, min(torch.get_num_threads(), len(cpus))))
        return node
    except Exception:
        return None
