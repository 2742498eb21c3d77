"""Parity of the CUDA path (through the C ABI) with the CPU oracle: bit-exact values, positions,
offsets and per-read status.  Run on the B200 box: pytest -m gpu."""
import ctypes
import os

import numpy as np
import pytest

import oracle
from bio_b200 import _cabi as cabi
from bio_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

OMODE = {cabi.MODE_SIMHASH: oracle.MODE_SIMHASH, cabi.MODE_PROTEIN_MINIMIZER: oracle.MODE_PROTEIN_MINIMIZER, cabi.MODE_KMER: oracle.MODE_KMER, cabi.MODE_NTHASH: oracle.MODE_NTHASH,
         cabi.MODE_MINIMIZER: oracle.MODE_MINIMIZER, cabi.MODE_SYNCMER: oracle.MODE_SYNCMER,
         cabi.MODE_PROTEIN: oracle.MODE_PROTEIN}


def assert_same(res, ref, label=""):
    assert np.array_equal(res["status"], ref["status"]), label + " status"
    assert np.array_equal(res["off"], ref["off"]), label + " offsets"
    assert res["total"] == len(ref["val"]), label + " count"
    assert np.array_equal(res["val"], ref["val"]), label + " values"
    if res["pos"] is not None and ref["pos"] is not None:
        assert np.array_equal(res["pos"], ref["pos"]), label + " positions"


def run_both(ctx, mode, bases, off, hint=0, **kw):
    p = cabi.make_params(mode, max_read_len=hint, **kw)
    res = ctx.run(p, bases, off)
    ref = oracle.run_batch(bases, off, OMODE[mode], threads=8, **kw)
    return res, ref


def test_reference_golden_minimizer_on_gpu(gpu_ctx, golden):
    # sketches/sketch_test.go:67-72 through the CUDA path
    g = golden["sketch"]["minimizer"]
    s = np.frombuffer(g["seq"].encode(), dtype=np.uint8)
    off = np.array([0, len(s)], dtype=np.uint64)
    res = gpu_ctx.run(cabi.make_params(cabi.MODE_MINIMIZER, g["k"], w=g["w"]), s, off)
    assert [int(v) for v in res["val"]] == g["values"]
    assert list(res["pos"]) == [0, 1, 4, 7, 8]


def test_reference_syncmer_input_on_gpu(gpu_ctx, golden):
    g = golden["sketch"]["syncmer"]
    s = np.frombuffer(g["seq"].encode(), dtype=np.uint8)
    off = np.array([0, len(s)], dtype=np.uint64)
    res = gpu_ctx.run(cabi.make_params(cabi.MODE_SYNCMER, g["k"], s=g["s"]), s, off)
    assert list(res["pos"]) == [0, 3, 5, 8, 11]
    assert int(res["val"][-1]) == 1955511966892880774


FIXTURE_CASES = {
    "nthash_k21": (cabi.MODE_NTHASH, dict(k=21)),
    "nthash_k21_fwd": (cabi.MODE_NTHASH, dict(k=21, canonical=False)),
    "minimizer_k21_w11": (cabi.MODE_MINIMIZER, dict(k=21, w=11)),
    "minimizer_k5_w3": (cabi.MODE_MINIMIZER, dict(k=5, w=3)),
    "syncmer_k21_s11": (cabi.MODE_SYNCMER, dict(k=21, s=11)),
    "kmer_k21": (cabi.MODE_KMER, dict(k=21)),
    "kmer_k5_both": (cabi.MODE_KMER, dict(k=5, canonical=False)),
    "protein_k11_f1": (cabi.MODE_PROTEIN, dict(k=11, frame=1)),
    "protein_k11_fm2": (cabi.MODE_PROTEIN, dict(k=11, frame=-2)),
}


@pytest.mark.parametrize("name", sorted(FIXTURE_CASES))
def test_committed_fixture(gpu_ctx, name):
    z = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))
    mode, kw = FIXTURE_CASES[name]
    res = gpu_ctx.run(cabi.make_params(mode, **kw), z["bases"], z["off"])
    ref = dict(val=z[name + "/val"], pos=z[name + "/pos"], off=z[name + "/off"], status=z[name + "/status"])
    assert_same(res, ref, name)


@pytest.mark.parametrize("mode,kw", [
    (cabi.MODE_NTHASH, dict(k=21)),
    (cabi.MODE_NTHASH, dict(k=31, canonical=False)),
    (cabi.MODE_MINIMIZER, dict(k=21, w=11)),
    (cabi.MODE_MINIMIZER, dict(k=31, w=15)),
    (cabi.MODE_MINIMIZER, dict(k=21, w=1)),
    (cabi.MODE_SYNCMER, dict(k=21, s=11)),
    (cabi.MODE_SYNCMER, dict(k=31, s=16)),
    (cabi.MODE_SYNCMER, dict(k=21, s=21)),
    (cabi.MODE_SYNCMER, dict(k=21, s=20)),
    (cabi.MODE_KMER, dict(k=21)),
    (cabi.MODE_KMER, dict(k=31, canonical=False)),
    (cabi.MODE_PROTEIN, dict(k=11, frame=1)),
    (cabi.MODE_PROTEIN, dict(k=11, frame=-3)),
])
def test_150bp_reads(gpu_ctx, mode, kw):
    b, o = synth.uniform_reads(30000, 150, 42)
    res, ref = run_both(gpu_ctx, mode, b, o, hint=150, **kw)
    assert_same(res, ref)
    res, ref = run_both(gpu_ctx, mode, b, o, hint=0, **kw)  # library measures the longest read itself
    assert_same(res, ref)


@pytest.mark.parametrize("mode,kw", [
    (cabi.MODE_NTHASH, dict(k=21)),
    (cabi.MODE_MINIMIZER, dict(k=21, w=11)),
    (cabi.MODE_SYNCMER, dict(k=21, s=11)),
    (cabi.MODE_KMER, dict(k=21)),
    (cabi.MODE_KMER, dict(k=15, canonical=False)),
    (cabi.MODE_PROTEIN, dict(k=11, frame=2)),
    (cabi.MODE_PROTEIN, dict(k=11, frame=-1)),
])
def test_ont_like_long_reads(gpu_ctx, mode, kw):
    L = synth.ont_like_lengths(400, 44)
    b, o = synth.ragged_reads(L, 44)
    res, ref = run_both(gpu_ctx, mode, b, o, **kw)
    assert_same(res, ref)
    res, ref = run_both(gpu_ctx, mode, b, o, hint=int(L.max()), **kw)
    assert_same(res, ref)


@pytest.mark.parametrize("read_len", [128, 160, 192, 256, 384])
@pytest.mark.parametrize("mode,kw", [
    (cabi.MODE_MINIMIZER, dict(k=21, w=11)),
    (cabi.MODE_MINIMIZER, dict(k=31, w=15)),
    (cabi.MODE_MINIMIZER, dict(k=21, w=20)),
    (cabi.MODE_SYNCMER, dict(k=21, s=11)),
    (cabi.MODE_SYNCMER, dict(k=15, s=11)),
])
def test_uniform_reads_of_a_multiple_of_32_bytes(gpu_ctx, mode, kw, read_len):
    # the lanes of a tile sit one read apart in shared memory: such tiles are staged with a word of skew per lane
    # (k_sparse_warp<.., SKEW>, restage_skewed).  Clean tiles, tiles with N / IUPAC bytes (staged twice:
    # fast-path bytes, then codes), a last tile of ONE read, then the same reads behind an 8-byte record (tiles no
    # longer start on a 16-byte boundary: bulk-copy path) and a batch whose last read is shorter / longer.
    n = 32 * 40 + 1
    b, o = synth.uniform_reads(n, read_len, 5)
    b = b.copy()
    rng = np.random.default_rng(6)
    dirty = rng.integers(64 * read_len, 200 * read_len, size=300)   # tiles 2..6
    b[dirty] = np.frombuffer(b"NnRYacgt", dtype=np.uint8)[rng.integers(0, 8, size=300)]
    for hint in (read_len, 0):
        res, ref = run_both(gpu_ctx, mode, b, o, hint=hint, **kw)
        assert_same(res, ref, f"uniform {read_len} hint {hint}")
    lens = np.array([8] + [read_len] * 100, dtype=np.uint64)
    b2, o2 = synth.ragged_reads(lens, 7)
    res, ref = run_both(gpu_ctx, mode, b2, o2, hint=read_len, **kw)
    assert_same(res, ref, "behind an 8-byte record")
    for last in (read_len - 28, 60):
        lens = np.array([read_len] * 63 + [last], dtype=np.uint64)
        b3, o3 = synth.ragged_reads(lens, 8)
        res, ref = run_both(gpu_ctx, mode, b3, o3, hint=read_len, **kw)
        assert_same(res, ref, f"last read of {last}")
    lens = np.array([read_len] * 31 + [read_len + 128 if read_len < 384 else 300] + [read_len] * 32, dtype=np.uint64)
    b4, o4 = synth.ragged_reads(lens, 9)
    res, ref = run_both(gpu_ctx, mode, b4, o4, hint=int(lens.max()), **kw)
    assert_same(res, ref, "one longer read at the end of a tile")


@pytest.mark.parametrize("seed", range(6))
def test_random_parameters_ragged_inputs(gpu_ctx, seed):
    """Random k/w/s over ragged batches that include empty and too-short reads, lowercase, N, IUPAC
    codes and arbitrary bytes (ntHash treats the forward strand by full byte and the reverse strand
    by byte & 7 -- SURVEY.md 8c)."""
    rng = np.random.default_rng(1000 + seed)
    alphabets = [b"ACGT", b"ACGTN", b"ACGTacgtNn", b"ACGTRYKMSWBDHVN", bytes(range(256)), b"AC", b"A"]
    lens = rng.integers(0, 700, size=300)
    lens[rng.integers(0, 300, size=40)] = 0
    b, o = synth.ragged_reads(lens, seed, alphabet=alphabets[seed % len(alphabets)])
    k = int(rng.integers(1, 40))
    w = int(rng.integers(1, 36))
    s = int(rng.integers(1, k + 1))
    for mode, kw in ((cabi.MODE_NTHASH, dict(k=k, canonical=bool(seed & 1))),
                     (cabi.MODE_MINIMIZER, dict(k=k, w=w)),
                     (cabi.MODE_SYNCMER, dict(k=k, s=s))):
        res, ref = run_both(gpu_ctx, mode, b, o, **kw)
        assert_same(res, ref, f"mode={mode} {kw}")


@pytest.mark.parametrize("mode,kw", [
    (cabi.MODE_NTHASH, dict(k=21)),
    (cabi.MODE_MINIMIZER, dict(k=21, w=11)),
    (cabi.MODE_SYNCMER, dict(k=21, s=11)),
    (cabi.MODE_KMER, dict(k=21)),
])
def test_circular(gpu_ctx, mode, kw):
    lens = [150] * 200 + [0, 10, 20, 21, 30, 31, 40, 41, 3000]
    b, o = synth.ragged_reads(lens, 3)
    res, ref = run_both(gpu_ctx, mode, b, o, circular=True, **kw)
    assert_same(res, ref)


def test_low_complexity_reads_overflow_staging(gpu_ctx):
    """poly-A / dinucleotide reads: every window has a new leftmost minimum, so items emit far more
    than the staged-list capacity and take the direct-write path."""
    b, o = synth.ragged_reads([150] * 700 + [5000] * 10, 9, alphabet=b"A")
    res, ref = run_both(gpu_ctx, cabi.MODE_MINIMIZER, b, o, k=21, w=11)
    assert_same(res, ref)
    assert ref["ties"] > 0
    b, o = synth.ragged_reads([5000] * 40, 9, alphabet=b"AC")
    res, ref = run_both(gpu_ctx, cabi.MODE_SYNCMER, b, o, k=21, s=11)
    assert_same(res, ref)


def test_empty_and_tiny_batches(gpu_ctx):
    p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11)
    res = gpu_ctx.run(p, np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert res["total"] == 0 and len(res["off"]) == 1
    b, o = synth.ragged_reads([0, 0, 0], 1)
    res = gpu_ctx.run(p, b, o)
    assert res["total"] == 0 and list(res["status"]) == [cabi.ERR_SHORT_SEQ] * 3
    b, o = synth.ragged_reads([31], 1)
    res, ref = run_both(gpu_ctx, cabi.MODE_MINIMIZER, b, o, k=21, w=11)
    assert_same(res, ref)
    assert res["total"] == 1


def test_kmer_illegal_base(gpu_ctx):
    """NextKmer stops at the first k-mer holding an illegal base (iterator.go:730-748): the codes before
    it are emitted, the read's status is ErrIllegalBase."""
    b, o = synth.ragged_reads([150] * 64 + [3000] * 4, 5)
    b = b.copy()
    for r, p in ((3, 0), (7, 100), (10, 149), (20, 20), (21, 21), (64, 1500), (65, 2999), (66, 5)):
        b[int(o[r]) + p] = ord("-")
    for canonical in (True, False):
        res, ref = run_both(gpu_ctx, cabi.MODE_KMER, b, o, k=21, canonical=canonical)
        assert_same(res, ref)
        assert (res["status"] == cabi.ERR_ILLEGAL_BASE).sum() == 8


def test_device_capacity_and_hint_errors(gpu_ctx):
    import torch
    dev = torch.device("cuda:0")
    b, o = synth.uniform_reads(5000, 150, 8)
    db = torch.zeros(len(b) + 64, dtype=torch.uint8, device=dev)
    db[:len(b)] = torch.from_numpy(b)
    do = torch.from_numpy(o.astype(np.int64)).to(dev)
    ooff = torch.empty(len(o), dtype=torch.int64, device=dev)
    st = torch.empty(len(o) - 1, dtype=torch.int32, device=dev)
    p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150)
    small = torch.empty(1000, dtype=torch.int64, device=dev)
    rc, need = gpu_ctx.run_device(p, db, do, len(b), small, None, ooff, st)
    assert rc == cabi.ERR_CAPACITY and need > 1000
    val = torch.empty(need, dtype=torch.int64, device=dev)
    pos = torch.empty(need, dtype=torch.int32, device=dev)
    rc, total = gpu_ctx.run_device(p, db, do, len(b), val, pos, ooff, st)
    assert rc == 0 and total == need
    ref = oracle.run_batch(b, o, oracle.MODE_MINIMIZER, k=21, w=11, threads=4)
    assert np.array_equal(val.cpu().numpy().view(np.uint64), ref["val"])
    assert np.array_equal(pos.cpu().numpy().view(np.uint32), ref["pos"])
    bad = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=100)  # hint below the real length
    with pytest.raises(cabi.SketchError) as e:
        gpu_ctx.run_device(bad, db, do, len(b), val, pos, ooff, st)
    assert e.value.code == cabi.ERR_BAD_ARG


def test_c2_c3_scale_properties(gpu_ctx):
    """2M x 150 bp on the device (C2/C3 geometry): size-independent properties + a sampled oracle check."""
    import torch
    dev = torch.device("cuda:0")
    n, L, k, w = 2_000_000, 150, 21, 11
    db, do = synth.device_uniform_reads(n, L, 43, dev)
    nb = n * L
    ooff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    # dense: exactly L-k+1 hashes per read
    ph = cabi.make_params(cabi.MODE_NTHASH, k, max_read_len=L, want_pos=False)
    hv = torch.empty(n * (L - k + 1), dtype=torch.int64, device=dev)
    rc, total = gpu_ctx.run_device(ph, db, do, nb, hv, None, ooff, st)
    assert rc == 0 and total == n * (L - k + 1)
    assert torch.equal(ooff, torch.arange(n + 1, device=dev, dtype=torch.int64) * (L - k + 1))
    # minimizers: values are the hashes at their positions; positions strictly increase inside a read
    # and consecutive ones are at most w apart; deterministic across runs
    pm = cabi.make_params(cabi.MODE_MINIMIZER, k, w=w, max_read_len=L)
    cap = int(cabi.lib().b200sk_output_bound(__import__("ctypes").byref(pm), nb, n, 0))
    mv = torch.empty(cap, dtype=torch.int64, device=dev)
    mp = torch.empty(cap, dtype=torch.int32, device=dev)
    moff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    rc, mt = gpu_ctx.run_device(pm, db, do, nb, mv, mp, moff, st)
    assert rc == 0 and int(st.abs().sum()) == 0
    mv, mp = mv[:mt], mp[:mt].long()
    read_of = torch.repeat_interleave(torch.arange(n, device=dev), moff[1:] - moff[:-1])
    assert torch.equal(hv[read_of * (L - k + 1) + mp], mv)
    same = read_of[1:] == read_of[:-1]
    dpos = mp[1:] - mp[:-1]
    assert bool((dpos[same] > 0).all()) and bool((dpos[same] <= w).all())
    first = moff[:-1]
    assert bool((mp[first] < w).all())  # the first window's minimum lies in the first w k-mers
    chk1 = int(mv.sum().item())
    mv2 = torch.empty(cap, dtype=torch.int64, device=dev)
    rc, mt2 = gpu_ctx.run_device(pm, db, do, nb, mv2, None, moff, st)
    assert mt2 == mt and int(mv2[:mt].sum().item()) == chk1
    # sampled oracle check: 20k reads taken out of the big batch
    idx = np.sort(np.random.default_rng(0).choice(n, 20000, replace=False))
    hb = db[:nb].view(n, L)[torch.from_numpy(idx).to(dev)].cpu().numpy().reshape(-1)
    ho = np.arange(len(idx) + 1, dtype=np.uint64) * np.uint64(L)
    ref = oracle.run_batch(hb, ho, oracle.MODE_MINIMIZER, k=k, w=w, threads=8)
    o_np = moff.cpu().numpy()
    mv_np = mv.cpu().numpy().view(np.uint64)
    got = np.concatenate([mv_np[o_np[i]:o_np[i + 1]] for i in idx])
    assert np.array_equal(got, ref["val"])


def test_host_pipeline_many_subbatches(gpu_ctx, monkeypatch):
    """b200sk_run splits the batch into sub-batches that flow through two device slots; force tiny
    sub-batches (ragged boundaries, unaligned starts) and compare the stitched result with the oracle."""
    monkeypatch.setenv("B200SK_SUB_BYTES", "20000")
    lens = np.concatenate([np.full(700, 150), synth.ont_like_lengths(30, 5, mean=4000), [0, 7, 31, 150, 33]])
    b, o = synth.ragged_reads(lens, 21)
    for mode, kw in ((cabi.MODE_MINIMIZER, dict(k=21, w=11)), (cabi.MODE_SYNCMER, dict(k=21, s=11)),
                     (cabi.MODE_NTHASH, dict(k=21)), (cabi.MODE_KMER, dict(k=21, canonical=False)),
                     (cabi.MODE_PROTEIN, dict(k=11, frame=-2)), (cabi.MODE_MINIMIZER, dict(k=21, w=11, circular=True))):
        res, ref = run_both(gpu_ctx, mode, b, o, **kw)
        assert_same(res, ref, f"mode={mode}")
    # a view that does not start at offset 0 of the buffer
    res = gpu_ctx.run(cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11), b, o[100:])
    ref = oracle.run_batch(b, o[100:], oracle.MODE_MINIMIZER, k=21, w=11, threads=4)
    assert np.array_equal(res["val"], ref["val"]) and np.array_equal(res["pos"], ref["pos"])
    assert np.array_equal(res["off"], ref["off"])


@pytest.mark.parametrize("w", list(range(1, 27)) + [33, 40])
def test_every_minimizer_window_size(gpu_ctx, w):
    """w = 2..24 run the register-window kernel (one template instantiation each), w = 1 the dense kernel,
    larger w the generic shared-memory-ring kernel."""
    lens = np.concatenate([np.full(300, 150), np.array([0, 19, 20 + w - 1, 20 + w, 400, 1500, 37, 251, 276, 277])])
    b, o = synth.ragged_reads(lens, 100 + w, alphabet=b"ACGTACGTACGTN")
    res, ref = run_both(gpu_ctx, cabi.MODE_MINIMIZER, b, o, k=21, w=w)
    assert_same(res, ref, f"w={w}")
    res, ref = run_both(gpu_ctx, cabi.MODE_MINIMIZER, b[: 300 * 150], o[:301], hint=150, k=21, w=w)
    assert_same(res, ref, f"w={w} hint")


@pytest.mark.parametrize("d", list(range(0, 15)))
def test_every_syncmer_window_size(gpu_ctx, d):
    """k-s = 1..12 run the register-window kernel, 0 the dense kernel, larger the generic kernel."""
    k = 24
    s = k - d
    lens = np.concatenate([np.full(300, 150), np.array([0, 2 * k - s - 2, 2 * k - s - 1, 2 * k - s, 400, 1500, 37, 251])])
    b, o = synth.ragged_reads(lens, 200 + d, alphabet=b"ACGTACGTACGTN")
    res, ref = run_both(gpu_ctx, cabi.MODE_SYNCMER, b, o, k=k, s=s)
    assert_same(res, ref, f"k-s={d}")
    res, ref = run_both(gpu_ctx, cabi.MODE_SYNCMER, b[: 300 * 150], o[:301], hint=150, k=k, s=s)
    assert_same(res, ref, f"k-s={d} hint")


@pytest.mark.parametrize("k", [1, 2, 3, 8, 31, 32, 33, 63, 64, 65, 100])
def test_kmer_sizes(gpu_ctx, k):
    """Small and large k, including k >= 64 where the rotations wrap (unverified against the Go module;
    GPU == oracle)."""
    b, o = synth.ragged_reads([150] * 200 + [0, k - 1 if k > 1 else 0, k, k + 1, 700], 300 + k)
    res, ref = run_both(gpu_ctx, cabi.MODE_NTHASH, b, o, k=k)
    assert_same(res, ref, f"nthash k={k}")
    res, ref = run_both(gpu_ctx, cabi.MODE_MINIMIZER, b, o, k=k, w=5)
    assert_same(res, ref, f"minimizer k={k}")
    if k <= 32:
        res, ref = run_both(gpu_ctx, cabi.MODE_KMER, b, o, k=k, canonical=bool(k & 1))
        assert_same(res, ref, f"kmer k={k}")
    if k >= 3:
        res, ref = run_both(gpu_ctx, cabi.MODE_SYNCMER, b, o, k=k, s=k - 2)
        assert_same(res, ref, f"syncmer k={k}")


@pytest.mark.parametrize("frame", [1, 2, 3, -1, -2, -3])
@pytest.mark.parametrize("k,w", [(10, 5), (11, 1), (5, 12)])
def test_protein_minimizer(gpu_ctx, frame, k, w):
    """ProteinMinimizerSketch (sketches/sketch-protein.go): SURVEY.md 8f row 2."""
    lens = np.concatenate([np.full(200, 150), synth.ont_like_lengths(12, 3, mean=3000),
                           [0, 3 * k - 1, 3 * k, 3 * k + w - 2, 3 * k + w - 1, 3 * k + w + 5, 3 * (k + w), 700]])
    b, o = synth.ragged_reads(lens, 400 + k + w, alphabet=b"ACGTACGTACGTN")
    res, ref = run_both(gpu_ctx, cabi.MODE_PROTEIN_MINIMIZER, b, o, k=k, w=w, frame=frame)
    assert_same(res, ref, f"frame={frame} k={k} w={w}")
    res, ref = run_both(gpu_ctx, cabi.MODE_PROTEIN_MINIMIZER, b[: 200 * 150], o[:201], hint=150, k=k, w=w, frame=frame)
    assert_same(res, ref, f"hint frame={frame} k={k} w={w}")


def test_protein_minimizer_amino_acid_input(gpu_ctx):
    aa = b"ACDEFGHIKLMNPQRSTVWY*X"
    lens = [0, 10, 29, 30, 34, 35, 60, 500, 2000] * 20
    b, o = synth.ragged_reads(lens, 77, alphabet=aa)
    p = cabi.make_params(cabi.MODE_PROTEIN_MINIMIZER, 10, w=5, alphabet=cabi.ALPHABET_PROTEIN)
    res = gpu_ctx.run(p, b, o)
    ref = oracle.run_batch(b, o, oracle.MODE_PROTEIN_MINIMIZER, k=10, w=5, alphabet=5, threads=4)
    assert_same(res, ref)
    p = cabi.make_params(cabi.MODE_PROTEIN, 10, alphabet=cabi.ALPHABET_PROTEIN)
    res = gpu_ctx.run(p, b, o)
    # amino-acid input through the oracle: hash every 10-mer of reads with at least 30 residues
    vals = []
    for i, L in enumerate(lens):
        if L >= 30:
            s = b[int(o[i]):int(o[i + 1])]
            vals += [oracle.wyhash(s[j:j + 10], 1) for j in range(L - 10 + 1)]
    assert [int(v) for v in res["val"]] == vals


@pytest.mark.parametrize("k,m,scale,canonical", [(21, 5, 5, True), (31, 5, 5, True), (31, 7, 1, False), (21, 21, 1, True),
                                                  (64, 4, 3, True), (100, 10, 8, False), (16, 5, 12, True)])
def test_simhash(gpu_ctx, k, m, scale, canonical):
    """SimHashIterator (sketches/iterator.go:113-612): SURVEY.md 8f row 3."""
    lens = np.concatenate([np.full(200, 150), synth.ont_like_lengths(10, 9, mean=2500), [0, k - 1, k, k + 1, 700]])
    b, o = synth.ragged_reads(lens, 500 + k + m, alphabet=b"ACGTACGTACGTNacgt")
    res, ref = run_both(gpu_ctx, cabi.MODE_SIMHASH, b, o, k=k, m=m, scale=scale, canonical=canonical)
    assert_same(res, ref, f"k={k} m={m} scale={scale}")
    res, ref = run_both(gpu_ctx, cabi.MODE_SIMHASH, b[: 200 * 150], o[:201], hint=150, k=k, m=m, scale=scale,
                        canonical=canonical, circular=True)
    assert_same(res, ref, "hint+circular")


@pytest.mark.parametrize("pw", [1, 2])
def test_narrow_positions(gpu_ctx, pw):
    """pos_width = 1 / 2: out_pos as uint8 / uint16 (same values), for every kernel family."""
    b, o = synth.uniform_reads(3000, 150, 77)
    for mode, kw in ((cabi.MODE_MINIMIZER, dict(k=21, w=11)), (cabi.MODE_MINIMIZER, dict(k=21, w=30)),
                     (cabi.MODE_SYNCMER, dict(k=21, s=11)), (cabi.MODE_NTHASH, dict(k=21)),
                     (cabi.MODE_KMER, dict(k=21, canonical=False)), (cabi.MODE_PROTEIN, dict(k=11, frame=-1))):
        p = cabi.make_params(mode, max_read_len=150, pos_width=pw, **kw)
        res = gpu_ctx.run(p, b, o)
        ref = oracle.run_batch(b, o, OMODE[mode], threads=4, **kw)
        assert res["pos"].dtype == (np.uint8 if pw == 1 else np.uint16)
        assert_same(res, ref, f"pw={pw} mode={mode}")
    # the hint is mandatory and must cover the positions
    with pytest.raises(cabi.SketchError):
        gpu_ctx.run(cabi.make_params(cabi.MODE_NTHASH, 21, pos_width=pw), b, o)
    if pw == 1:
        with pytest.raises(cabi.SketchError):
            gpu_ctx.run(cabi.make_params(cabi.MODE_NTHASH, 21, pos_width=1, max_read_len=300), b, o)


@pytest.mark.parametrize("k,canonical", [(21, True), (21, False), (5, True), (31, True), (64, True), (2, False)])
def test_nthash_values_only_kernel(gpu_ctx, k, canonical):
    """want_pos=False takes the warp-tile ntHash kernel (b200sk_nthash.cu): all-ACGT fast path with pair tables,
    general path for any other byte, whole and partial 16-step blocks, long reads in chunks, circular."""
    cases = []
    cases.append(synth.uniform_reads(20000, 150, 5) + (150,))                       # fast path, 8 blocks + tail of 2
    cases.append(synth.uniform_reads(3000, 16 + k - 1, 6) + (16 + k - 1,))          # exactly one whole block
    cases.append(synth.ragged_reads(np.random.default_rng(7).integers(0, 400, size=3000), 7) + (0,))
    cases.append(synth.ragged_reads(np.random.default_rng(8).integers(0, 300, size=2000), 8,
                                    alphabet=b"ACGTNacgtRYKMSWBDHVU-*") + (0,))      # general tables
    cases.append(synth.ragged_reads([150] * 999 + [0, 3, 20], 9, alphabet=b"ACGTacgt") + (150,))  # lower case, fast
    L = synth.ont_like_lengths(200, 45)
    cases.append(synth.ragged_reads(L, 45) + (0,))                                   # chunked items
    for b, o, hint in cases:
        for circular in (False, True):
            p = cabi.make_params(cabi.MODE_NTHASH, k, canonical=canonical, circular=circular, max_read_len=hint,
                                 want_pos=False)
            res = gpu_ctx.run(p, b, o)
            ref = oracle.run_batch(b, o, oracle.MODE_NTHASH, threads=8, k=k, canonical=canonical, circular=circular)
            assert res["pos"] is None
            assert_same(res, ref, f"k={k} canonical={canonical} circular={circular} hint={hint}")


@pytest.mark.parametrize("k", [1, 5, 21, 31, 32])
def test_kmer_values_only_kernel(gpu_ctx, k):
    """Canonical k-mer codes with want_pos=False take the warp-tile kernel: IUPAC codes (first base), illegal bases
    (the read stops before the first k-mer holding one, status ErrIllegalBase), ragged and long reads."""
    cases = [synth.uniform_reads(20000, 150, 15) + (150,),
             synth.ragged_reads(np.random.default_rng(16).integers(0, 400, size=3000), 16) + (0,),
             synth.ragged_reads(np.random.default_rng(17).integers(0, 300, size=2000), 17,
                                alphabet=b"ACGTNacgtRYKMSWBDHVU") + (0,),
             synth.ragged_reads(np.random.default_rng(18).integers(30, 300, size=2000), 18,
                                alphabet=b"ACGTACGTACGTACGTACGTACGT-*X") + (0,),
             synth.ragged_reads(synth.ont_like_lengths(150, 46), 46) + (0,)]
    for b, o, hint in cases:
        p = cabi.make_params(cabi.MODE_KMER, k, canonical=True, max_read_len=hint, want_pos=False)
        res = gpu_ctx.run(p, b, o)
        ref = oracle.run_batch(b, o, oracle.MODE_KMER, threads=8, k=k, canonical=True)
        assert_same(res, ref, f"k={k} hint={hint}")


@pytest.mark.parametrize("frame", [1, 2, 3, -1, -2, -3])
@pytest.mark.parametrize("k", [1, 3, 8, 9, 11, 16])
def test_protein_values_only_kernel(gpu_ctx, frame, k):
    """ProteinIterator with want_pos=False and k <= 16 on short reads takes the warp-tile kernel: amino acids in a
    register window, 64-entry tables for all-ACGT tiles, CodonTable.Get for any other byte."""
    cases = [synth.uniform_reads(6000, 150, 25) + (150,),
             synth.ragged_reads(np.random.default_rng(26).integers(0, 380, size=3000), 26) + (380,),
             synth.ragged_reads(np.random.default_rng(27).integers(0, 300, size=2000), 27,
                                alphabet=b"ACGTNacgtRYKMSWBDHVU-*") + (300,),
             synth.ragged_reads([150] * 500 + [0, 2, 33, 34, 35], 28, alphabet=b"ACGTacgt") + (150,)]
    for table in (1, 11):
        for b, o, hint in cases:
            p = cabi.make_params(cabi.MODE_PROTEIN, k, frame=frame, codon_table=table, max_read_len=hint, want_pos=False)
            res = gpu_ctx.run(p, b, o)
            ref = oracle.run_batch(b, o, oracle.MODE_PROTEIN, threads=8, k=k, frame=frame, codon_table=table)
            assert_same(res, ref, f"k={k} frame={frame} table={table} hint={hint}")


def test_protein_values_only_amino_acid_input(gpu_ctx):
    aa = b"ACDEFGHIKLMNPQRSTVWY*X"
    lens = [0, 10, 29, 30, 34, 35, 60, 200, 380] * 20
    b, o = synth.ragged_reads(lens, 78, alphabet=aa)
    for k in (5, 10, 16):
        p = cabi.make_params(cabi.MODE_PROTEIN, k, alphabet=cabi.ALPHABET_PROTEIN, max_read_len=380, want_pos=False)
        res = gpu_ctx.run(p, b, o)
        vals = []
        for i, L in enumerate(lens):
            if L >= 3 * k:  # iterator-protein.go:50: the length check is on 3k even for amino-acid input
                s = b[int(o[i]):int(o[i + 1])]
                vals += [oracle.wyhash(s[j:j + k], 1) for j in range(L - k + 1)]
        assert [int(v) for v in res["val"]] == vals


def test_protein_minimizer_low_complexity_overflow(gpu_ctx):
    """Homopolymer reads translate to one repeated residue: every window moves its leftmost minimum, the staged
    lists overflow and the items are walked again straight to global memory (register-window kernel, k <= 16)."""
    b, o = synth.ragged_reads([150] * 300 + [900] * 40, 31, alphabet=b"A")
    b2, o2 = synth.ragged_reads([150] * 64, 32)
    bases = np.concatenate([b, b2])
    off = np.concatenate([o, o2[1:] + o[-1]])
    for k, w, frame in ((10, 5, 1), (7, 3, -2), (16, 24, 3)):
        res, ref = run_both(gpu_ctx, cabi.MODE_PROTEIN_MINIMIZER, bases, off, k=k, w=w, frame=frame)
        assert_same(res, ref, f"k={k} w={w} frame={frame}")


def test_two_streams_one_context(gpu_ctx):
    """One context, batches enqueued alternately on two streams without any host synchronisation in between: the
    context's scratch words serve one batch at a time, so the library orders the batches on the device
    (include/b200sketch.h, b200sk_ctx) -- every batch must still be bit-exact."""
    import torch
    dev = torch.device("cuda:0")
    p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150)
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    jobs = []
    for i in range(6):
        b, o = synth.uniform_reads(40000 + 5000 * i, 150, 900 + i)
        db = torch.from_numpy(np.concatenate([b, np.zeros(64, np.uint8)])).to(dev)
        do = torch.from_numpy(o.astype(np.int64)).to(dev)
        n = len(o) - 1
        cap = int(cabi.lib().b200sk_output_bound(ctypes.byref(p), len(b), n, 0))
        val = torch.zeros(cap, dtype=torch.int64, device=dev)
        pos = torch.zeros(cap, dtype=torch.int32, device=dev)
        ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        st = torch.zeros(n, dtype=torch.int32, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        jobs.append((b, o, db, do, val, pos, ooff, st, flags))
    torch.cuda.synchronize()
    for i, (b, o, db, do, val, pos, ooff, st, flags) in enumerate(jobs):
        gpu_ctx.enqueue_device(p, db, do, len(b), val, pos, ooff, st, flags, stream=streams[i & 1].cuda_stream)
    torch.cuda.synchronize()
    for i, (b, o, db, do, val, pos, ooff, st, flags) in enumerate(jobs):
        ref = oracle.run_batch(b, o, oracle.MODE_MINIMIZER, threads=8, k=21, w=11)
        t = int(ooff[-1].item())
        assert int(flags.item()) == 0
        res = dict(val=val[:t].cpu().numpy().view(np.uint64), pos=pos[:t].cpu().numpy().view(np.uint32),
                   off=ooff.cpu().numpy().view(np.uint64), status=st.cpu().numpy(), total=t)
        assert_same(res, ref, f"batch {i} on stream {i & 1}")


@pytest.mark.parametrize("k,table,alphabet,hint", [(11, 1, b"ACGT", 150), (11, 11, b"ACGTNacgtRYKMSWBDHVU", 150),
                                                    (16, 1, b"ACGT", 150), (5, 4, b"ACGTacgt", 384),
                                                    (20, 1, b"ACGT", 150), (11, 1, b"ACGT", 0)])
def test_protein_six_frames_one_call(gpu_ctx, k, table, alphabet, hint):
    """b200sk_enqueue_device_frames: ProteinIterator over the six frames of every read (BASELINE.json config 5) -- the
    fused kernel (hint <= 384, k <= 16) and the six-batch path behind the same entry point, against the oracle frame
    by frame: values, offsets, statuses; ragged reads (too short for some frames, empty), IUPAC and lower case."""
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(31 + k)
    top = hint if hint else 150
    lens = np.concatenate([np.full(3000, top), rng.integers(0, top + 1, size=2000), [0, 1, 2, 3 * k - 1, 3 * k, 3 * k + 1, 3 * k + 2]])
    if hint == 0:
        lens = np.concatenate([lens, synth.ont_like_lengths(20, 5, mean=1500)])
    b, o = synth.ragged_reads(lens, 77 + k, alphabet=alphabet)
    n = len(o) - 1
    p = cabi.make_params(cabi.MODE_PROTEIN, k, codon_table=table, max_read_len=hint, want_pos=False)
    db = torch.from_numpy(np.concatenate([b, np.zeros(64, np.uint8)])).to(dev)
    do = torch.from_numpy(o.astype(np.int64)).to(dev)
    cap = int(cabi.lib().b200sk_output_bound(ctypes.byref(p), len(b), n, 1)) + 8
    vals = [torch.full((cap,), -1, dtype=torch.int64, device=dev) for _ in range(6)]
    offs = [torch.zeros(n + 1, dtype=torch.int64, device=dev) for _ in range(6)]
    sts = [torch.full((n,), 99, dtype=torch.int32, device=dev) for _ in range(6)]
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    l0 = gpu_ctx.kernel_launches()
    gpu_ctx.enqueue_device_frames(p, db, do, len(b), vals, offs, sts, flags)
    torch.cuda.synchronize()
    launches = gpu_ctx.kernel_launches() - l0
    assert int(flags.item()) == 0
    if hint and k <= 16:
        assert launches == 7  # six offset scans + ONE sketching launch
    for fi, frame in enumerate((1, 2, 3, -1, -2, -3)):
        ref = oracle.run_batch(b, o, oracle.MODE_PROTEIN, threads=8, k=k, codon_table=table, frame=frame)
        t = int(offs[fi][-1].item())
        res = dict(val=vals[fi][:t].cpu().numpy().view(np.uint64), pos=None, off=offs[fi].cpu().numpy().view(np.uint64),
                   status=sts[fi].cpu().numpy(), total=t)
        assert_same(res, ref, f"frame {frame}")
        assert int(vals[fi][t].item()) == -1


@pytest.mark.parametrize("k,hint", [(11, 150), (7, 0)])
def test_protein_six_frames_host_entry(gpu_ctx, k, hint):
    """b200sk_run_frames: the host form of the six-frame call (one copy in, six sketches out), a batch that does not
    start at offset 0 included."""
    lens = np.concatenate([np.full(2000, 150), np.random.default_rng(3).integers(0, 151, size=1500), [0, 3 * k - 1, 3 * k]])
    b, o = synth.ragged_reads(lens, 321, alphabet=b"ACGTACGTACGTN")
    p = cabi.make_params(cabi.MODE_PROTEIN, k, max_read_len=hint, want_pos=False)
    for skip in (0, 5):  # skip > 0: read_off[0] != 0
        res = gpu_ctx.run_frames(p, b, o[skip:])
        for fi, frame in enumerate((1, 2, 3, -1, -2, -3)):
            ref = oracle.run_batch(b[int(o[skip]):], o[skip:] - o[skip], oracle.MODE_PROTEIN, threads=8, k=k, frame=frame)
            assert_same(res[fi], ref, f"frame {frame} skip {skip}")
    with pytest.raises(cabi.SketchError):
        gpu_ctx.run_frames(cabi.make_params(cabi.MODE_NTHASH, 21), b, o)
