"""Multi-GPU plumbing for batches of independent frames: one process per GPU, frames sharded
contiguously across ranks, calibration maps broadcast once, no collective in the steady state
(SURVEY.md §8e).  torch.distributed is used for the rendezvous, the one-off broadcast and the
final timing reduction only."""
import numpy as np


def shard_range(n_frames, world_size, rank):
    """contiguous share [lo, hi) of rank `rank`; the first n % world ranks get one frame more"""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError('bad rank %r / world size %r' % (rank, world_size))
    base, extra = divmod(int(n_frames), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_calibration(maps, src=0, device=None):
    """Broadcast a dict of calibration arrays (dark / flat float32 maps, lens 3x3 / 1x5 float64) from rank
    `src` to every rank.  Every rank passes a dict with the same keys; non-source ranks pass None values or
    arrays to be overwritten.  Shapes / dtypes travel first (object broadcast), then one tensor broadcast per
    array — over NCCL/NVLink when the process group is NCCL and `device` is a CUDA device, over gloo on CPU.
    Returns {key: torch tensor on `device`} (None entries stay None)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    keys = sorted(maps)
    meta = None
    if rank == src:
        meta = []
        for k in keys:
            v = maps[k]
            if v is None:
                meta.append(None)
            else:
                a = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
                meta.append((tuple(a.shape), str(a.dtype).replace('torch.', '')))
    box = [meta]
    dist.broadcast_object_list(box, src=src)
    meta = box[0]
    out = {}
    dev = torch.device('cpu') if device is None else torch.device(device)
    for k, m in zip(keys, meta):
        if m is None:
            out[k] = None
            continue
        shape, dtype = m
        dt = getattr(torch, dtype)
        if rank == src:
            v = maps[k]
            t = (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))).to(dev, dt).contiguous()
        else:
            t = torch.empty(shape, dtype=dt, device=dev)
        dist.broadcast(t, src=src)
        out[k] = t
    return out


def reduce_max(value, device=None):
    """max over ranks of a python float (timing is the slowest rank's)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the format of sysfs local_cpulist)"""
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def bind_host_to_gpu(device_index, sysfs='/sys/bus/pci/devices'):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs local_cpulist of the GPU's PCI
    function), so that the page-locked staging buffers allocated afterwards are node-local: with one process per
    GPU the host<->device streams of the ranks then do not all cross the same socket interconnect.  Returns a
    dict describing what was done; never raises (a box without the sysfs entries is left as it is)."""
    import os
    info = {'bound': False}
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        info['pci'] = bdf
        with open(os.path.join(sysfs, bdf, 'local_cpulist')) as f:
            cpus = parse_cpulist(f.read())
        try:
            with open(os.path.join(sysfs, bdf, 'numa_node')) as f:
                info['numa_node'] = int(f.read().strip())
        except (OSError, ValueError):
            pass
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            info['bound'] = True
        info['cpus'] = len(allowed)
    except Exception as e:                      # noqa: BLE001 - best effort by design
        info['error'] = '%s: %s' % (type(e).__name__, e)
    return info
