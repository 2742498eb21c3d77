"""The reference's own tests for the path, transcribed against the host-side mirror of its interface
(bio_b200.sketches).  Same inputs, same assertions (sketches/iterator_test.go, sketches/sketch_test.go,
sketches/iterator-protein_test.go), plus the values the CPU oracle gives where the reference asserts only counts."""
import numpy as np
import pytest

import oracle
from bio_b200 import sketches as sk

pytestmark = pytest.mark.gpu


def test_TestKmerIterator(golden):
    # sketches/iterator_test.go:31-66
    g = golden["iterator"]["TestKmerIterator"]
    sequence = sk.NewSeq(sk.DNA, g["seq"])
    k = g["k"]
    it = sk.NewKmerIterator(sequence, k, True, False)
    codes = []
    while True:
        code, ok, err = it.Next()
        assert err is None
        if not ok:
            break
        codes.append(code)
    assert len(codes) == len(g["seq"]) - k + 1
    assert codes[:3] == [49027, 196109, 784436]
    ref, _, _ = oracle.kmer_iterator(g["seq"], k, canonical=True)
    assert codes == [int(x) for x in ref]


def test_TestHashIterator(golden):
    # sketches/iterator_test.go:68-103
    g = golden["iterator"]["TestHashIterator"]
    sequence = sk.NewSeq(sk.DNA, g["seq"])
    it = sk.NewHashIterator(sequence, g["k"], True, False)
    codes = []
    while True:
        code, ok = it.NextHash()
        if not ok:
            break
        codes.append(code)
        assert it.Index() == len(codes) - 1
    assert len(codes) == len(g["seq"]) - g["k"] + 1
    ref, _ = oracle.hash_iterator(g["seq"], g["k"], canonical=True)
    assert codes == [int(x) for x in ref]


def test_TestMinimizer(golden):
    # sketches/sketch_test.go:33-76 -- the known-answer test of the path
    g = golden["sketch"]["minimizer"]
    sequence = sk.NewSeq(sk.DNA, g["seq"])
    sketch = sk.NewMinimizerSketch(sequence, g["k"], g["w"], False)
    codes, idxs = [], []
    while True:
        code, ok = sketch.NextMinimizer()
        if not ok:
            break
        idxs.append(sketch.Index())
        codes.append(code)
    assert len(codes) == 5
    assert codes == [973456138564179607, 2645801399420473919, 1099502864234245338, 6763474888237448943,
                     2737971715116251183]
    assert idxs == [0, 1, 4, 7, 8]


def test_TestSyncmer(golden):
    # sketches/sketch_test.go:78-117 (the reference asserts nothing; values from the oracle)
    g = golden["sketch"]["syncmer"]
    sequence = sk.NewSeq(sk.DNA, g["seq"])
    sketch = sk.NewSyncmerSketch(sequence, g["k"], g["s"], False)
    codes, idxs = [], []
    while True:
        code, ok = sketch.NextSyncmer()
        if not ok:
            break
        idxs.append(sketch.Index())
        codes.append(code)
    rv, ri, _, _ = oracle.syncmer(g["seq"], g["k"], g["s"])
    assert codes == [int(x) for x in rv] and idxs == [int(x) for x in ri]


def test_TestProteinIterator():
    # sketches/iterator-protein_test.go:29-62: count only in the reference
    s = "AAGTTTGAATCATTCAACTATCTAGTTTTCAGAGAACAATGTTCTCTAAAGAATAGAAAAGAGTCATTGTGCGGTGATGATGGCGGGAAGGATCCACCTG"
    sequence = sk.NewSeq(sk.DNA, s)
    k = 10
    it = sk.NewProteinIterator(sequence, k, 1, 1)
    codes = []
    while True:
        code, ok = it.Next()
        if not ok:
            break
        codes.append(code)
        assert it.Index() == len(codes) - 1
    assert len(codes) == len(s) // 3 - k + 1
    ref, _ = oracle.protein_iterator(s, k, 1, 1)
    assert codes == [int(x) for x in ref]


def test_constructor_errors():
    s = sk.NewSeq(sk.DNA, "ACGTACGTAC")
    with pytest.raises(sk.ErrInvalidK):
        sk.NewHashIterator(s, 0, True, False)
    with pytest.raises(sk.ErrShortSeq):
        sk.NewHashIterator(s, 11, True, False)
    with pytest.raises(sk.ErrInvalidW):
        sk.NewMinimizerSketch(s, 5, 0, False)
    with pytest.raises(sk.ErrShortSeq):
        sk.NewMinimizerSketch(s, 5, 7, False)
    with pytest.raises(sk.ErrInvalidS):
        sk.NewSyncmerSketch(s, 5, 6, False)
    with pytest.raises(sk.ErrInvalidS):
        sk.NewSyncmerSketch(s, 5, 0, False)
    with pytest.raises(sk.ErrKOverflow):
        sk.NewKmerIterator(sk.NewSeq(sk.DNA, "ACGT" * 20), 33, True, False)
    with pytest.raises(sk.ErrInvalidM):
        sk.NewSimHashIterator(s, 8, 3, 1, True, False)
    with pytest.raises(sk.ErrInvalidM):
        sk.NewSimHashIterator(s, 8, 9, 1, True, False)
    with pytest.raises(sk.ErrInvalidScale):
        sk.NewSimHashIterator(s, 8, 5, 5, True, False)
    with pytest.raises(sk.ErrKTooLarge):
        sk.NewSimHashIterator(s, 65535, 5, 1, True, False)
    with pytest.raises(sk.ErrShortSeq):
        sk.NewProteinIterator(s, 4, 1, 1)


def test_illegal_base_surfaces_from_NextKmer():
    # iterator.go:730-748: the codes before the failing k-mer come out, then the error
    it = sk.NewKmerIterator(sk.NewSeq(sk.DNA, "ACGTAC-GTACGT"), 4, True, False)
    codes = []
    while True:
        code, ok, err = it.NextKmer()
        if not ok:
            assert isinstance(err, sk.ErrIllegalBase)
            break
        codes.append(code)
    ref, _, _ = oracle.kmer_iterator("ACGTAC-GTACGT", 4)
    assert codes == [int(x) for x in ref] and len(codes) == 3


def test_noncanonical_kmers_walk_both_strands():
    # iterator.go:713-723: forward strand, then the reverse-complemented sequence
    s = "AAGTTTGAATCATTCAACTATCTAGTTTTCAGRYACN"
    for alphabet in (sk.DNAredundant, sk.DNA):
        it = sk.NewKmerIterator(sk.NewSeq(alphabet, s), 5, False, False)
        codes, idxs = [], []
        while True:
            code, ok, err = it.NextKmer()
            if not ok:
                break
            codes.append(code)
            idxs.append(it.Index())
        ref, _, _ = oracle.kmer_iterator(s, 5, canonical=False, alphabet=0 if alphabet == sk.DNAredundant else 1)
        n = len(s) - 5 + 1
        assert codes == [int(x) for x in ref] and len(codes) == 2 * n
        assert idxs == list(range(n)) + list(range(n))


def test_batch_replay_equals_single_calls():
    rng = np.random.default_rng(3)
    reads = ["".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(20, 400, size=40)]
    b = sk.Batch()
    for r in reads:
        b.Add(r)
    res = b.MinimizerSketch(21, 11, False)
    for i, r in enumerate(reads):
        if len(r) < 31:
            with pytest.raises(sk.ErrShortSeq):
                res.iterator(i)
            continue
        it = res.iterator(i)
        rv, ri, _, _ = oracle.minimizer(r, 21, 11)
        got = []
        while True:
            v, ok = it.Next()
            if not ok:
                break
            got.append((v, it.Index()))
        assert got == [(int(v), int(p)) for v, p in zip(rv, ri)]


def test_TestProteinMinimizer():
    # sketches/sketch-protein_test.go:29-58 (the reference asserts nothing; values from the oracle)
    s = "AAGTTTGAATCATTCAACTATCTAGTTTTCAGAGAACAATGTTCTCTAAAGAATAGAAAAGAGTCATTGTGCGGTGATGATGGCGGGAAGGATCCACCTG"
    sequence = sk.NewSeq(sk.DNA, s)
    k, w = 10, 3
    sketch = sk.NewProteinMinimizerSketch(sequence, k, 1, 1, w)
    got = []
    while True:
        code, ok = sketch.Next()
        if not ok:
            break
        got.append((code, sketch.Index()))
    rv, ri, err, _ = oracle.protein_minimizer(s, k, w, 1, 1)
    assert err == 0 and got == [(int(v), int(i)) for v, i in zip(rv, ri)] and len(got) > 0
    with pytest.raises(sk.ErrInvalidW):
        sk.NewProteinMinimizerSketch(sequence, k, 1, 1, 0)
    with pytest.raises(sk.ErrShortSeq):
        sk.NewProteinMinimizerSketch(sk.NewSeq(sk.DNA, s[:31]), k, 1, 1, w)


def test_TestSimHashIterator():
    # sketches/iterator_test.go:105-145: two 21-bp strings (one with a lowercase base), k=21, m=5, scale=5;
    # the reference asserts the count only, the values come from the oracle
    for _s in ("GAACAATGTTCTCTAAAATTG", "GcACAATGTTCTCTAAAATTG"):
        sequence = sk.NewSeq(sk.DNA, _s)
        k = 21
        it = sk.NewSimHashIterator(sequence, k, 5, 5, True, False)
        codes = []
        while True:
            code, ok = it.NextSimHash()
            if not ok:
                break
            codes.append(code)
        assert len(codes) == len(_s) - k + 1
        ref, err = oracle.simhash_iterator(_s, k, 5, 5, True, False)
        assert err == 0 and codes == [int(x) for x in ref]


def test_batch_protein_frames_equals_six_iterators():
    """Batch.ProteinFrames (one call, b200sk_run_frames) replays like six NewProteinIterator loops per record
    (sketches/iterator-protein.go:46-90), frame 1, 2, 3, -1, -2, -3; too-short records raise ErrShortSeq."""
    rng = np.random.default_rng(11)
    reads = ["".join(rng.choice(list("ACGTN"), size=int(n), p=[.24, .24, .24, .24, .04])) for n in rng.integers(10, 200, size=60)]
    b = sk.Batch()
    for r in reads:
        b.Add(r)
    k = 7
    frames = b.ProteinFrames(k, 1)
    assert len(frames) == 6
    for fi, frame in enumerate((1, 2, 3, -1, -2, -3)):
        for i, r in enumerate(reads):
            if len(r) < 3 * k:
                with pytest.raises(sk.ErrShortSeq):
                    frames[fi].iterator(i)
                continue
            it = frames[fi].iterator(i)
            single = sk.NewProteinIterator(sk.NewSeq(sk.DNAredundant, r), k, 1, frame)
            got, want = [], []
            while True:
                v, ok = it.Next()
                if not ok:
                    break
                got.append((v, it.Index()))
            while True:
                v, ok = single.Next()
                if not ok:
                    break
                want.append((v, single.Index()))
            assert got == want, (frame, i)
