"""Multi-GPU side of the C ABI (include/b200sketch.h): shard boundaries (host logic, CPU), the device
group (one process, a context + worker thread per device), the gather buffer over CUDA IPC with the sketching kernel
storing straight into the root's memory, and the in-place segment compaction."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from bio_b200 import _cabi as cabi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_by_bases_host_logic():
    lens = synth.ont_like_lengths(5000, 3)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    for world in (1, 2, 3, 8):
        cut = cabi.shard_by_bases(off, world)
        assert cut[0] == 0 and cut[-1] == len(lens) and np.all(np.diff(cut.astype(np.int64)) >= 0)
        per = np.diff(off[cut.astype(np.int64)].astype(np.int64))
        assert per.max() - per.min() <= 2 * int(lens.max())  # balanced up to one read either side
    # more shards than reads: empty shards, still a partition
    cut = cabi.shard_by_bases(np.array([0, 10, 30], dtype=np.uint64), 8)
    assert cut[0] == 0 and cut[-1] == 2 and np.all(np.diff(cut.astype(np.int64)) >= 0)
    # bench.py's equal-count shards: block-aligned for 1/2/4/8 ranks, a partition for any count
    sys.path.insert(0, ROOT)
    import bench
    for n in (100_000_000, 10_000_000, 1234567, 3):
        for world in (1, 2, 4, 8):
            b = bench.shard_bounds(n, world)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
    b = bench.shard_bounds(100_000_000, 8)
    assert all(x % bench.BLOCK == 0 for x in b[:-1])


def _devices():
    import torch
    return [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]


def _same(res, ref):
    assert np.array_equal(res["off"], ref["off"])
    assert np.array_equal(res["val"], ref["val"])
    assert np.array_equal(res["pos"], ref["pos"])
    assert np.array_equal(res["status"], ref["status"])


@pytest.mark.gpu
def test_group_one_process_several_contexts():
    """b200sk_group_run: two contexts driven from two threads of one process (two devices when the box has them,
    else two contexts on device 0), reads sharded by bases, results assembled in read order."""
    g = cabi.Group(_devices())
    assert g.size() == 2
    b, o = synth.uniform_reads(30000, 150, 5)
    res = g.run(cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150), b, o)
    _same(res, oracle.run_batch(b, o, oracle.MODE_MINIMIZER, k=21, w=11, threads=8))
    lens = np.array([0, 5, 30, 31, 32, 150, 0, 0, 400, 20, 31, 1000, 3, 151, 20000] * 40)
    b, o = synth.ragged_reads(lens, 7, alphabet=b"ACGTNacgt")
    for mode, omode, kw in ((cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, dict(k=21, w=11)),
                            (cabi.MODE_SYNCMER, oracle.MODE_SYNCMER, dict(k=21, s=11)),
                            (cabi.MODE_NTHASH, oracle.MODE_NTHASH, dict(k=21))):
        res = g.run(cabi.make_params(mode, **kw), b, o)
        _same(res, oracle.run_batch(b, o, omode, threads=8, **kw))
    # fewer reads than devices, and none at all
    b1, o1 = synth.uniform_reads(1, 150, 9)
    _same(g.run(cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11), b1, o1),
          oracle.run_batch(b1, o1, oracle.MODE_MINIMIZER, k=21, w=11))
    res = g.run(cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11), np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert res["total"] == 0 and res["off"].tolist() == [0]
    with pytest.raises(cabi.SketchError) as e:
        g.run(cabi.make_params(cabi.MODE_MINIMIZER, 21, w=0), b1, o1)
    assert e.value.code == cabi.ERR_INVALID_W
    assert g.kernel_launches() > 0
    g.close()


def _view(addr, n, dev):
    import torch

    class M:
        pass
    m = M()
    m.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (addr, False), "version": 2}
    return torch.as_tensor(m, device=dev)


@pytest.mark.gpu
def test_compact_segments_in_place():
    import torch
    ctx = cabi.Context(0)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    for counts, gaps in (([1000, 5, 70000, 0, 333], [0, 17, 1, 90000, 4]), ([10], [0]), ([0, 0, 7], [3, 3, 3]),
                         ([300000, 300001, 299999], [0, 1, 100000])):
        base, cur = [], 0
        for c, g in zip(counts, gaps):
            cur += g
            base.append(cur)
            cur += c
        base[0] = 0
        handle, addr = ctx.gather_create(cur + 8)
        buf = _view(addr, cur + 8, dev)
        buf.fill_(-1)
        want = []
        for b, c in zip(base, counts):
            v = torch.from_numpy(rng.integers(0, 2**62, size=c, dtype=np.int64)).to(dev)
            buf[b:b + c] = v
            want.append(v)
        ctx.compact_segments(addr, base, counts)
        torch.cuda.synchronize()
        assert torch.equal(buf[:sum(counts)], torch.cat(want))
        del buf
        ctx.gather_close(addr, True)
    ctx.close()


@pytest.mark.gpu
def test_ipc_gather_two_processes():
    """Root (this process) exports the gather buffer; a second PROCESS maps it and its sketching kernel writes the
    second shard's minimizers straight into its segment; after compaction the buffer is the single-GPU output."""
    import torch
    ndev = torch.cuda.device_count()
    child_dev = 1 if ndev >= 2 else 0
    dev = torch.device("cuda", 0)
    n_total, r_split = 24000, 11000
    b, o = synth.uniform_reads(n_total, 150, 77)
    ref = oracle.run_batch(b, o, oracle.MODE_MINIMIZER, k=21, w=11, threads=8)
    ctx = cabi.Context(0)
    p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150)
    L = cabi.lib()
    cap0 = int(L.b200sk_output_bound(C.byref(p), r_split * 150, r_split, 0))
    cap1 = int(L.b200sk_output_bound(C.byref(p), (n_total - r_split) * 150, n_total - r_split, 0))
    handle, gaddr = ctx.gather_create(cap0 + cap1)
    child = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gather_child.py"), handle.hex(), str(cap0),
                              str(cap1), str(child_dev), str(n_total), str(r_split), str(n_total)],
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    bases = torch.from_numpy(np.concatenate([b[:r_split * 150], np.zeros(64, dtype=np.uint8)])).to(dev)
    off = torch.from_numpy(o[:r_split + 1].astype(np.int64)).to(dev)
    pos = torch.empty(cap0, dtype=torch.int32, device=dev)
    ooff = torch.empty(r_split + 1, dtype=torch.int64, device=dev)
    st = torch.empty(r_split, dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    ctx.enqueue_device_raw(p, bases, off, r_split * 150, gaddr, cap0, pos, ooff, st, flags)
    torch.cuda.synchronize()
    out, err = child.communicate(timeout=300)
    assert child.returncode == 0, err[-2000:]
    c1 = int([ln for ln in out.splitlines() if ln.startswith("COUNT")][0].split()[1])
    c0 = int(ooff[r_split].item())
    assert c0 + c1 == len(ref["val"])
    ctx.compact_segments(gaddr, [0, cap0], [c0, c1])
    torch.cuda.synchronize()
    got = _view(gaddr, c0 + c1, dev).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, ref["val"])
    ctx.gather_close(gaddr, True)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode,chunk", [("minimizer", 64), ("minimizer", 1024), ("syncmer", 96)])
def test_sharded_chain_two_processes(mode, chunk):
    """b200sk_enqueue_device_sharded: two ranks (two processes, two GPUs) share ONE output chain -- their kernels'
    look-back runs through both GPUs' status words and every flush stores at its exact place in the root's arrays.
    The root must end up holding exactly the single-GPU output (values, uint8 positions, offsets, statuses)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (two kernels that wait for each other must run at the same time)")
    dev = torch.device("cuda", 0)
    n_total = 20000 + 17
    b, o = synth.uniform_reads(n_total, 150, 91)
    b = b.copy()
    b[150 * 40:150 * 40 + 30] = ord("N")  # a tile that leaves the all-ACGT fast path
    if mode == "syncmer":
        p = cabi.make_params(cabi.MODE_SYNCMER, 21, s=11, max_read_len=150, pos_width=1)
        ref = oracle.run_batch(b, o, oracle.MODE_SYNCMER, k=21, s=11, threads=8)
    else:
        p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150, pos_width=1)
        ref = oracle.run_batch(b, o, oracle.MODE_MINIMIZER, k=21, w=11, threads=8)
    ctx = cabi.Context(0)
    cap = int(cabi.lib().b200sk_output_bound(C.byref(p), n_total * 150, n_total, 0))
    n_tiles = (n_total + 31) // 32
    bufs = [ctx.gather_create(cap), ctx.gather_create(cap // 8 + 8), ctx.gather_create(n_total + 1),
            ctx.gather_create(n_total // 2 + 8), ctx.gather_create(n_tiles + 1)]
    (hv, val), (hp, pos), (ho, off), (hs, status), (ht, state0) = bufs
    child = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_sharded_child.py"), "1", str(n_total), str(chunk),
                              mode, hv.hex(), hp.hex(), ho.hex(), hs.hex(), ht.hex(), str(cap)],
                             stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    line = child.stdout.readline()
    assert line.startswith("HANDLE"), child.stderr.read()[-2000:]
    state1 = ctx.gather_open(bytes.fromhex(line.split()[1]))
    mine = [c for c in range((n_total + chunk - 1) // chunk) if c % 2 == 0]
    lb = np.concatenate([b[c * chunk * 150:min(n_total, (c + 1) * chunk) * 150] for c in mine] + [np.zeros(64, np.uint8)])
    n = (len(lb) - 64) // 150
    bases = torch.from_numpy(lb).to(dev)
    loff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 150
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    child.stdin.write("GO\n")
    child.stdin.flush()
    for epoch in (1, 2):
        _view(val, cap, dev).fill_(-1)
        torch.cuda.synchronize()
        spec = cabi.ShardSpec()
        spec.rank, spec.n_ranks, spec.chunk_reads, spec.epoch, spec.n_reads_global = 0, 2, chunk, epoch, n_total
        spec.state[0], spec.state[1] = state0, state1
        ctx.enqueue_device_sharded(p, spec, bases, loff, n * 150, val, pos, off, status, cap, flags)
        torch.cuda.synchronize()
        assert child.stdout.readline().startswith("DONE"), child.stderr.read()[-2000:]
        assert int(flags.item()) == 0
        total = len(ref["val"])
        g_off = _view(off, n_total + 1, dev).cpu().numpy().view(np.uint64)
        assert np.array_equal(g_off, ref["off"])
        assert np.array_equal(_view(val, total, dev).cpu().numpy().view(np.uint64), ref["val"])
        g_pos = _view(pos, (total + 7) // 8, dev).cpu().numpy().view(np.uint8)[:total]
        assert np.array_equal(g_pos, ref["pos"].astype(np.uint8))
        g_st = _view(status, (n_total + 1) // 2, dev).cpu().numpy().view(np.int32)[:n_total]
        assert np.array_equal(g_st, ref["status"])
        child.stdin.write("NEXT\n")
        child.stdin.flush()
    assert child.wait(timeout=120) == 0, child.stderr.read()[-2000:]
    ctx.gather_close(state1, False)
    for _, a in bufs:
        ctx.gather_close(a, True)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["minimizer", "syncmer"])
@pytest.mark.parametrize("pw,want_pos", [(1, True), (2, True), (0, True), (0, False)])
def test_sharded_chain_single_rank_bulk_stores(mode, pw, want_pos):
    """The SHARD kernels on ONE rank (a chain of one GPU): their flush leaves as bulk shared->global stores with a
    16-byte-aligned body and per-lane head / tail elements -- every position width, output arrays that start at odd
    element addresses, and a second batch behind the first (out_base-free, same arrays)."""
    import torch
    dev = torch.device("cuda", 0)
    n_total = 9000 + 5
    b, o = synth.uniform_reads(n_total, 150, 123)
    b = b.copy()
    b[150 * 77:150 * 77 + 40] = ord("N")
    b[150 * 500:150 * 501] = ord("A")  # a low-complexity read: list overflow path
    kw = dict(k=21, s=11) if mode == "syncmer" else dict(k=21, w=11)
    omode = oracle.MODE_SYNCMER if mode == "syncmer" else oracle.MODE_MINIMIZER
    p = cabi.make_params(cabi.MODE_SYNCMER if mode == "syncmer" else cabi.MODE_MINIMIZER, max_read_len=150,
                         pos_width=pw, want_pos=want_pos, **kw)
    ref = oracle.run_batch(b, o, omode, threads=8, **kw)
    total = len(ref["val"])
    ctx = cabi.Context(0)
    cap = int(cabi.lib().b200sk_output_bound(C.byref(p), n_total * 150, n_total, 0))
    bases = torch.from_numpy(np.concatenate([b, np.zeros(64, np.uint8)])).to(dev)
    loff = torch.arange(n_total + 1, dtype=torch.int64, device=dev) * 150
    n_tiles = (n_total + 31) // 32
    pdt = {1: torch.uint8, 2: torch.int16}.get(pw, torch.int32)
    for shift in (0, 1, 3):  # output arrays starting `shift` elements into their allocations
        val = torch.full((cap + 8,), -1, dtype=torch.int64, device=dev)
        pos = torch.zeros(cap + 8, dtype=pdt, device=dev)
        off = torch.zeros(n_total + 1, dtype=torch.int64, device=dev)
        status = torch.zeros(n_total, dtype=torch.int32, device=dev)
        state = torch.zeros(n_tiles + 1, dtype=torch.int64, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        for epoch in (1, 2):
            spec = cabi.ShardSpec()
            spec.rank, spec.n_ranks, spec.chunk_reads, spec.epoch, spec.n_reads_global = 0, 1, 96, epoch, n_total
            spec.state[0] = state.data_ptr()
            ctx.enqueue_device_sharded(p, spec, bases, loff, n_total * 150, val[shift:].data_ptr(),
                                       pos[shift:].data_ptr() if want_pos else 0, off.data_ptr(), status.data_ptr(),
                                       cap, flags)
            torch.cuda.synchronize()
            assert int(flags.item()) == 0
            assert np.array_equal(off.cpu().numpy().view(np.uint64), ref["off"])
            assert np.array_equal(val[shift:shift + total].cpu().numpy().view(np.uint64), ref["val"])
            assert int(val[shift + total].item()) == -1 and (shift == 0 or int(val[shift - 1].item()) == -1)
            if want_pos:
                npdt = {1: np.uint8, 2: np.uint16}.get(pw, np.uint32)
                g = pos[shift:shift + total].cpu().numpy().view(npdt)
                assert np.array_equal(g, ref["pos"].astype(npdt))
                assert int(pos[shift + total].item()) == 0
            assert np.array_equal(status.cpu().numpy(), ref["status"])
    ctx.close()
