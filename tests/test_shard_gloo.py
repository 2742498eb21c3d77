"""N>1 host logic on CPU: world_size-2 gloo.  Each rank takes its contiguous shard of the reads, sketches
it (the oracle stands in for the GPU here -- test infrastructure only), and the uint64 arrays are gathered
to rank 0 in rank order; the result must equal sketching the whole batch at once."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from bio_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = synth.ont_like_lengths(60, 3, mean=2000)
    bases, off = synth.ragged_reads(lens, 3)
    ok = True
    for bounds in (shard.shard_bounds(len(lens), world), shard.shard_bounds_by_bases(off, world)):
        lo, hi = bounds[rank], bounds[rank + 1]
        sub_off = off[lo:hi + 1] - off[lo]
        sub = bases[int(off[lo]):int(off[hi])]
        r = oracle.run_batch(sub, sub_off, oracle.MODE_MINIMIZER, k=21, w=11)
        local = torch.from_numpy(r["val"].view(np.int64).copy())
        out, counts = shard.gather_hashes(local, dist, dst=0)
        if rank == 0:
            whole = oracle.run_batch(bases, off, oracle.MODE_MINIMIZER, k=21, w=11)
            ok = ok and np.array_equal(out.numpy().view(np.uint64), whole["val"])
            ok = ok and sum(counts) == len(whole["val"])
    if rank == 0:
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_bounds():
    sys.path.insert(0, ROOT)
    from bio_b200 import shard
    assert shard.shard_bounds(10, 4) == [0, 2, 5, 7, 10]
    off = np.array([0, 100, 100, 1100, 1200, 1300], dtype=np.uint64)
    b = shard.shard_bounds_by_bases(off, 2)
    assert b[0] == 0 and b[-1] == 5 and b == sorted(b)
    b = shard.shard_bounds_by_bases(off, 8)
    assert len(b) == 9 and b == sorted(b)


def test_host_binding_helpers():
    """cpulist parsing; binding is a no-op (and says so) where there is no GPU / sysfs entry to bind to."""
    import os
    from bio_b200 import shard
    assert shard._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert shard._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    prev, what = shard.bind_host_to_device(0)
    assert prev == before and isinstance(what, str)
    shard.restore_host_binding(prev)
    assert os.sched_getaffinity(0) == before
