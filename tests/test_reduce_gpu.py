"""FracMinHash scale filter + sort + unique on the resident hash arrays.  Oracle = numpy on the
oracle's own stream (np.unique of the values <= MaxUint64 // scale, the rule of sketches/iterator.go:180-185)."""
import numpy as np
import pytest

import oracle
from bio_b200 import _cabi as cabi, synth

M64 = (1 << 64) - 1


def test_scale_max_hash_is_the_references():
    # iterator.go:184: maxHash = math.MaxUint64 / uint64(scale)
    L = cabi.lib()
    for s in (1, 2, 5, 10, 1000, 2**31):
        assert int(L.b200sk_scale_max_hash(s)) == M64 // s
    assert int(L.b200sk_scale_max_hash(0)) == M64


def _reduce(ctx, vals, scale, unique, cap=None):
    import torch
    dev = torch.device("cuda", 0)
    d = torch.from_numpy(vals.view(np.int64).copy()).to(dev)
    out = torch.empty(len(vals) + 1 if cap is None else cap, dtype=torch.int64, device=dev)
    rc, n = ctx.reduce_device(d, len(vals), out, scale=scale, unique=unique)
    return rc, n, out[:n].cpu().numpy().view(np.uint64) if rc == 0 else None


def _want(vals, scale, unique):
    keep = vals[vals <= np.uint64(M64 // max(scale, 1))]
    return np.unique(keep) if unique else np.sort(keep)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 2, 4095, 4096, 4097, 100_003, 3_000_000])
def test_sort_unique_random(gpu_ctx, n):
    rng = np.random.default_rng(n)
    vals = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    if n > 10:
        vals[rng.integers(0, n, size=n // 3)] = vals[rng.integers(0, n, size=n // 3)]  # duplicates
    for scale, unique in ((1, True), (1, False), (10, True), (1000, True), (3, False), (2**20, True)):
        rc, m, got = _reduce(gpu_ctx, vals, scale, unique)
        want = _want(vals, scale, unique)
        assert rc == 0 and m == len(want), (n, scale, unique)
        assert np.array_equal(got, want), (n, scale, unique)


@pytest.mark.gpu
def test_sort_degenerate_keys(gpu_ctx):
    for vals in (np.zeros(10000, dtype=np.uint64), np.full(10000, M64, dtype=np.uint64),
                 np.arange(50000, dtype=np.uint64)[::-1].copy(), (np.arange(70000, dtype=np.uint64) % 7) << np.uint64(56),
                 np.array([5, 5, 5, 1, 1, M64, 0, 0], dtype=np.uint64)):
        for unique in (True, False):
            rc, m, got = _reduce(gpu_ctx, vals, 1, unique)
            assert rc == 0 and np.array_equal(got, _want(vals, 1, unique))


@pytest.mark.gpu
def test_capacity_is_reported(gpu_ctx):
    vals = np.random.default_rng(3).integers(0, 2**64, size=50000, dtype=np.uint64)
    rc, need, _ = _reduce(gpu_ctx, vals, 2, True, cap=100)
    assert rc == cabi.ERR_CAPACITY and need == int(np.count_nonzero(vals <= np.uint64(M64 // 2)))
    rc, need, _ = _reduce(gpu_ctx, vals, 1, True, cap=100)
    assert rc == cabi.ERR_CAPACITY and need == 50000


@pytest.mark.gpu
def test_minimizer_stream_reduced(gpu_ctx):
    """the whole f4 chain on a real stream: sketch on the device, reduce there, compare with sort|uniq of the oracle"""
    import ctypes as C
    import torch
    dev = torch.device("cuda", 0)
    b, o = synth.uniform_reads(40000, 150, 21)
    ref = oracle.run_batch(b, o, oracle.MODE_MINIMIZER, k=21, w=11, threads=8)
    p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150, want_pos=False)
    n, nb = len(o) - 1, len(b)
    cap = int(cabi.lib().b200sk_output_bound(C.byref(p), nb, n, 0))
    bases = torch.from_numpy(np.concatenate([b, np.zeros(64, np.uint8)])).to(dev)
    off = torch.from_numpy(o.astype(np.int64)).to(dev)
    val = torch.empty(cap, dtype=torch.int64, device=dev)
    ooff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    rc, total = gpu_ctx.run_device(p, bases, off, nb, val, None, ooff, st)
    assert rc == 0 and total == len(ref["val"])
    for scale in (1, 8, 200):
        v = val[:total].clone()
        out = torch.empty(total + 1, dtype=torch.int64, device=dev)
        rc, m = gpu_ctx.reduce_device(v, total, out, scale=scale, unique=True)
        want = _want(ref["val"], scale, True)
        assert rc == 0 and m == len(want)
        assert np.array_equal(out[:m].cpu().numpy().view(np.uint64), want)


@pytest.mark.gpu
def test_run_reduced_host_path(gpu_ctx, monkeypatch):
    """b200sk_run_reduced: host pointers in, only the reduced sketch back; several sub-batches through the pipeline"""
    monkeypatch.setenv("B200SK_SUB_BYTES", "300000")  # ~2000 reads per sub-batch
    b, o = synth.uniform_reads(30000, 150, 33)
    ref = oracle.run_batch(b, o, oracle.MODE_MINIMIZER, k=21, w=11, threads=8)
    p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150)
    for scale, unique in ((1, True), (1, False), (50, True), (1000, True)):
        got = gpu_ctx.run_reduced(p, b, o, scale=scale, unique=unique)
        assert np.array_equal(got, _want(ref["val"], scale, unique)), (scale, unique)
    lens = np.array([0, 5, 30, 31, 150, 400, 20000, 3] * 30)
    b, o = synth.ragged_reads(lens, 7, alphabet=b"ACGTN")
    ref = oracle.run_batch(b, o, oracle.MODE_SYNCMER, k=21, s=11, threads=8)
    got = gpu_ctx.run_reduced(cabi.make_params(cabi.MODE_SYNCMER, 21, s=11), b, o, scale=4)
    assert np.array_equal(got, _want(ref["val"], 4, True))
    got = gpu_ctx.run_reduced(p, np.zeros(0, np.uint8), np.zeros(1, np.uint64), scale=4)
    assert len(got) == 0
    # the plain host path still works on the same context afterwards
    b, o = synth.uniform_reads(3000, 150, 34)
    res = gpu_ctx.run(p, b, o)
    ref = oracle.run_batch(b, o, oracle.MODE_MINIMIZER, k=21, w=11, threads=8)
    assert np.array_equal(res["val"], ref["val"]) and np.array_equal(res["pos"], ref["pos"])
