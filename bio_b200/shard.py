"""Multi-GPU plumbing of the sketching path: records shard embarrassingly (no cross-record state in any
reference constructor: sketches/iterator.go:615-655, sketches/sketch.go:85-202), so rank r sketches a
contiguous range of reads on its own GPU; the only exchange is an optional gather of the per-GPU uint64
hash arrays (rank order == read order, so no sort is needed).  torch.distributed is plumbing here: NCCL
over NVLink on the GPUs, gloo in the CPU tests."""
import os
import time

import numpy as np


def shard_bounds(n_reads, world):
    """Contiguous equal-count shards: rank r owns reads [b[r], b[r+1])."""
    return [n_reads * r // world for r in range(world + 1)]


def shard_bounds_by_bases(read_off, world):
    """Contiguous shards balanced by cumulative bases (long, skewed reads: C4)."""
    read_off = np.asarray(read_off, dtype=np.uint64)
    n = len(read_off) - 1
    total = int(read_off[-1] - read_off[0])
    bounds = [0]
    for r in range(1, world):
        target = int(read_off[0]) + total * r // world
        bounds.append(int(np.searchsorted(read_off, np.uint64(target), side="left")))
    bounds.append(n)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [min(b, n) for b in bounds]


def gather_hashes(local, dist, dst=0):
    """Gather variable-length 1-D tensors to rank dst, concatenated in rank order.
    Returns (tensor on dst | None, counts list)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    cnt = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    if rank == dst:
        out = torch.empty(sum(counts), dtype=local.dtype, device=local.device)
        offs = np.concatenate([[0], np.cumsum(counts)])
        ops = []
        for r in range(world):
            seg = out[offs[r]:offs[r + 1]]
            if r == dst:
                seg.copy_(local)
            elif counts[r]:
                ops.append(dist.P2POp(dist.irecv, seg, r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out, counts
    if local.numel():
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst)]):
            w.wait()
    return None, counts


def timed_gather(part, dist, dev, reps=3):
    """Time the NCCL gather of one uint64 hash array per GPU to rank 0 (device timing, max over ranks)."""
    import torch
    gather_hashes(part[:1024], dist)  # warm up the P2P channels
    torch.cuda.synchronize()
    best = None
    nbytes = 0
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, counts = gather_hashes(part, dist)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        nbytes = (sum(counts) - counts[0]) * part.element_size()
        best = ms if best is None else min(best, ms)
        del out
    return {"bytes_over_nvlink": nbytes, "ms": best, "GBps": nbytes / (best * 1e-3) / 1e9 if best else None,
            "what": "gather of each GPU's uint64 minimizer array (1/world of its output) to rank 0 over NCCL"}


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_device(device):
    """Run the calling thread (and the threads it starts) on the CPUs of the NUMA node GPU `device` hangs off, so
    that the pinned batch buffers it allocates next are local to that GPU's PCIe root (sysfs `local_cpulist` of the
    GPU's PCI function).  With several ranks on one host, every rank's H2D/D2H traffic then stays on its own socket
    instead of crossing the inter-socket link.  Returns (previous affinity, description); a host without the sysfs
    entry is left as it is."""
    import torch
    prev = os.sched_getaffinity(0)
    try:
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        cpus = _parse_cpulist(open(base + "/local_cpulist").read()) & prev
        node = open(base + "/numa_node").read().strip()
        if not cpus:
            return prev, "unbound (no local cpus in the allowed set for %s)" % bdf
        os.sched_setaffinity(0, cpus)
        return prev, "numa node %s of %s, %d cpus" % (node, bdf, len(cpus))
    except Exception as e:  # no sysfs / no such attribute: nothing to bind to
        return prev, "unbound (%s)" % type(e).__name__


def restore_host_binding(prev):
    try:
        os.sched_setaffinity(0, prev)
    except Exception:
        pass
