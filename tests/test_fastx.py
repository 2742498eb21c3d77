"""Record feeder: oracle = literal restatement of seqio/fastx.Reader.Read/parseRecord
(oracle/fastx_oracle.c); GPU = b200sk_fastx_parse_device / b200sk_run_fastx through the C ABI."""
import os

import numpy as np
import pytest

import oracle

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fastx")  # copies of the reference's fixtures


def make_fastq(n, read_len, seed, crlf=False, at_quals=False, tail_newline=True, lower=False):
    rng = np.random.default_rng(seed)
    nl = b"\r\n" if crlf else b"\n"
    out = []
    seqs = []
    for i in range(n):
        L = int(read_len if np.isscalar(read_len) else read_len[i])
        s = bytes(np.frombuffer(b"acgtn" if lower else b"ACGTN", dtype=np.uint8)[rng.integers(0, 5 if i % 7 == 0 else 4, size=L)])
        q = bytearray(rng.integers(33, 74, size=L, dtype=np.uint8).tobytes())
        if at_quals and L and i % 3 == 0:
            q[0] = ord("@")
        seqs.append(s)
        out.append(b"@r%d desc %d" % (i, i) + nl + s + nl + b"+" + nl + bytes(q) + nl)
    text = b"".join(out)
    if not tail_newline and text:
        text = text[:-len(nl)]
    return text, seqs


def make_fasta(n, seed, width=60, crlf=False, blank_lines=False):
    rng = np.random.default_rng(seed)
    nl = b"\r\n" if crlf else b"\n"
    out, seqs = [], []
    for i in range(n):
        L = int(rng.integers(0, 700))
        s = bytes(np.frombuffer(b"ACGTacgtN>@", dtype=np.uint8)[rng.integers(0, 9 if i % 5 else 11, size=L)])
        # a '>' or '@' inside a line is data, but never at a line start
        rows = [s[j:j + width] for j in range(0, L, width)]
        rows = [(b"A" + r[1:]) if r[:1] in (b">", b"@") else r for r in rows]
        seqs.append(b"".join(rows))
        out.append(b">seq%d some description" % i + nl)
        for r in rows:
            out.append(r + nl)
            if blank_lines and rng.integers(0, 6) == 0:
                out.append(nl)
    return b"".join(out), seqs


# ------------------------------------------------------------------ oracle (CPU)
def test_oracle_pinned_by_reference_fixtures():
    """reader_test.go:84,105,125,130-158: record counts of the reference's own fixtures."""
    want = {"test.fa": 6, "test.fq": 8, "test2.fq": 5, "test3.fq": 3}
    for name, n in want.items():
        r = oracle.fastx_parse(open(os.path.join(REF, name), "rb").read())
        assert r["status"] == 0 and r["n_records"] == n, name
    r = oracle.fastx_parse(open(os.path.join(REF, "test3.fq"), "rb").read())
    lens = np.diff(r["read_off"].astype(np.int64))
    assert len(set(lens.tolist())) == 1  # reader_test.go:130-158: equal-length records, '@'-leading quality lines


def test_oracle_rules():
    r = oracle.fastx_parse(b"\n\n@r1 d\nACGT\n+\n@III\n@r2\nAC\r\n+\r\nII\r\n")
    assert r["status"] == 0 and r["format"] == oracle.FASTX_FASTQ and r["n_records"] == 2
    assert bytes(r["bases"]) == b"ACGTAC" and r["read_off"].tolist() == [0, 4, 6]
    assert r["rec_off"].tolist() == [2, 20] and r["qual_off"].tolist() == [15, 31]
    r = oracle.fastx_parse(b">a b\nAC\nGT\n\n>b\nTT>A\nC")
    assert r["n_records"] == 2 and bytes(r["bases"]) == b"ACGTTT>AC" and r["name_len"].tolist() == [3, 1]
    assert oracle.fastx_parse(b"hello\n")["status"] == oracle.ERR_NOT_FASTX
    assert oracle.fastx_parse(b"\r\n>a\nAC\n")["status"] == oracle.ERR_NOT_FASTX  # only '\n' may lead (reader.go:286)
    assert oracle.fastx_parse(b"@r\nACGT\n+\nII\n")["status"] == oracle.ERR_BAD_FASTQ
    # multi-line FASTQ is legal for the reference (reader.go:396-412)
    r = oracle.fastx_parse(b"@r\nAC\nGT\n+\nII\nII\n@s\nA\n+\nI\n")
    assert r["status"] == 0 and r["n_records"] == 2 and bytes(r["bases"]) == b"ACGTA"
    text, seqs = make_fastq(50, 37, 3, at_quals=True)
    r = oracle.fastx_parse(text)
    assert r["n_records"] == 50 and bytes(r["bases"]) == b"".join(seqs)
    text, seqs = make_fasta(40, 4, blank_lines=True, crlf=True)
    r = oracle.fastx_parse(text)
    assert r["n_records"] == 40 and bytes(r["bases"]) == b"".join(seqs)


def test_head_id_desc():
    from bio_b200.fastx import parse_head_id_and_desc as f
    assert f(b"id desc more") == (b"id", b"desc more")
    assert f(b"id\t  desc") == (b"id", b"desc")
    assert f(b"id") == (b"id", b"")
    assert f(b"id   ") == (b"id", b"")


# ------------------------------------------------------------------ GPU parity
def _ctx():
    from bio_b200 import _cabi as cabi
    return cabi, cabi.Context(0)


def _parse_gpu(ctx, text, fmt=0, final=True):
    import torch
    n = len(text)
    host = np.zeros((n + 15) // 16 * 16 + 16, dtype=np.uint8)
    host[:n] = np.frombuffer(text, dtype=np.uint8)
    d = torch.from_numpy(host).cuda()
    info = ctx.fastx_parse_device(d, n, fmt, final)
    return info, ctx.fastx_fetch(info)


def _same(g, o, fastq):
    assert g["n_records"] == o["n_records"]
    assert np.array_equal(g["read_off"], o["read_off"])
    assert np.array_equal(g["bases"], o["bases"])
    assert np.array_equal(g["rec_off"][:-1], o["rec_off"])
    if fastq:
        assert np.array_equal(g["qual_off"], o["qual_off"])


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(crlf=True), dict(at_quals=True), dict(tail_newline=False),
                                dict(lower=True, at_quals=True, crlf=True)])
def test_fastq_parity(kw):
    cabi, ctx = _ctx()
    for n, L, seed in ((1, 150, 1), (1000, 150, 2), (333, 1, 3), (5000, 151, 4)):
        text, seqs = make_fastq(n, L, seed, **kw)
        for lead in (b"", b"\n\n\n"):
            info, g = _parse_gpu(ctx, lead + text)
            o = oracle.fastx_parse(lead + text)
            assert o["status"] == 0 and g["format"] == cabi.FASTX_FASTQ
            _same(g, o, True)
            assert g["consumed"] == len(lead + text) and g["max_read_len"] == L
    ctx.close()


@pytest.mark.gpu
def test_fastq_ragged_and_long():
    cabi, ctx = _ctx()
    rng = np.random.default_rng(9)
    lens = rng.integers(0, 3000, size=400)
    lens[5] = 70000
    text, seqs = make_fastq(len(lens), lens, 11, at_quals=True)
    info, g = _parse_gpu(ctx, text + b"\n\n")  # trailing blank lines are not a record
    o = oracle.fastx_parse(text + b"\n\n")
    _same(g, o, True)
    assert bytes(g["bases"]) == b"".join(seqs)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(crlf=True), dict(blank_lines=True), dict(width=7), dict(width=100000)])
def test_fasta_parity(kw):
    cabi, ctx = _ctx()
    for n, seed in ((1, 1), (300, 2), (2000, 3)):
        text, seqs = make_fasta(n, seed, **kw)
        for tail in (b"", b"\n"):
            t = (text + tail) if tail else text[:-1] if text.endswith(b"\n") and not kw.get("crlf") else text
            info, g = _parse_gpu(ctx, t)
            o = oracle.fastx_parse(t)
            assert o["status"] == 0 and g["format"] == cabi.FASTX_FASTA
            _same(g, o, False)
    ctx.close()


@pytest.mark.gpu
def test_errors_and_edges():
    cabi, ctx = _ctx()
    for bad in (b"hello\n", b"\r\n>a\nAC\n", b" \n"):  # b" \n" = blank.fx, reader_test.go:160
        with pytest.raises(cabi.SketchError) as e:
            _parse_gpu(ctx, bad)
        assert e.value.code == cabi.ERR_NOT_FASTX
    for bad in (b"@r\nACGT\n+\nII\n", b"@r\nAC\nGT\n+\nII\nIII\n", b"@r\nACGT\n+\nIIII\n@s\nAC\n"):
        with pytest.raises(cabi.SketchError) as e:
            _parse_gpu(ctx, bad)
        assert e.value.code == cabi.ERR_BAD_FASTQ
    # a broken record whose sequence line is far longer than its quality line must not be copied past the base
    # buffer (sized n_bytes / 2): the call fails cleanly and the context stays usable
    big = b"@a\n" + b"ACGT" * 300_000 + b"\n+\nI\n"
    with pytest.raises(cabi.SketchError) as e:
        _parse_gpu(ctx, big)
    assert e.value.code == cabi.ERR_BAD_FASTQ
    with pytest.raises(cabi.SketchError) as e:
        _parse_gpu(ctx, b"@ok\nAC\n+\nII\n" + big)
    assert e.value.code == cabi.ERR_BAD_FASTQ
    info, g = _parse_gpu(ctx, b"@ok\nACGT\n+\nIIII\n")
    assert g["n_records"] == 1 and bytes(g["bases"]) == b"ACGT"
    info, g = _parse_gpu(ctx, b"\n\n\n")
    assert g["n_records"] == 0
    # reader.go:286-294: leading blank lines are tolerated up to byte 10240 only
    for lead in (10241, 10242, 20000, 2_000_000):
        for tail in (b">a\nACGT\n", b""):
            text = b"\n" * lead + tail
            o = oracle.fastx_parse(text)
            if o["status"] == 0:
                info, g = _parse_gpu(ctx, text)
                assert g["n_records"] == o["n_records"]
            else:
                assert o["status"] == cabi.ERR_NOT_FASTX
                with pytest.raises(cabi.SketchError) as e:
                    _parse_gpu(ctx, text)
                assert e.value.code == cabi.ERR_NOT_FASTX
    for name, want in (("blank.fx", cabi.ERR_NOT_FASTX), ("blank1.fx", 0), ("empty.fx", 0)):  # reader_test.go:160-197
        text = open(os.path.join(REF, name), "rb").read()
        o = oracle.fastx_parse(text)
        assert o["status"] == want and o["n_records"] == 0
    info, g = _parse_gpu(ctx, b">only header")
    o = oracle.fastx_parse(b">only header")
    assert g["n_records"] == o["n_records"] == 1 and g["read_off"].tolist() == [0, 0]
    # the reference's test4.fa: a record without sequence between two others, a two-line sequence
    t4 = b">a\nATC\n>b\n>123\nATCGN\n>abcdefg\nATCGN\nGCCTN\n"
    info, g = _parse_gpu(ctx, t4)
    o = oracle.fastx_parse(t4)
    _same(g, o, False)
    assert g["read_off"].tolist() == [0, 3, 3, 8, 18] and bytes(g["bases"]) == b"ATCATCGNATCGNGCCTN"
    ctx.close()


@pytest.mark.gpu
def test_streaming_chunks_and_reader():
    """A text cut at arbitrary places: consumed/carry-over reproduces the whole-file parse; Reader.Read()
    yields the reference's Record fields."""
    from bio_b200 import fastx
    cabi, ctx = _ctx()
    for text, seqs, fq in (make_fastq(700, 150, 5, at_quals=True) + (True,), make_fasta(300, 6, blank_lines=True) + (False,)):
        o = oracle.fastx_parse(text)
        rd = fastx.Reader(text, ctx=ctx, chunk_bytes=10007)
        recs = list(rd)
        assert len(recs) == o["n_records"] and rd.IsFastq == fq
        assert [r.Seq for r in recs] == seqs
        for i in (0, len(recs) // 2, len(recs) - 1):
            start = int(o["rec_off"][i])
            name = text[start + 1:start + 1 + int(o["name_len"][i])]
            assert recs[i].Name == name and recs[i].ID == name.split(b" ")[0]
            if fq:
                q = int(o["qual_off"][i])
                assert recs[i].Qual == text[q:q + len(seqs[i])]
    if os.path.isdir(REF):
        for name, n in {"test.fa": 6, "test.fq": 8, "test2.fq": 5, "test3.fq": 3}.items():
            assert len(list(fastx.Reader(os.path.join(REF, name), ctx=ctx))) == n
    ctx.close()


@pytest.mark.gpu
def test_run_fastx_matches_sketch_of_oracle_records():
    """text -> records -> minimizers in one C-ABI call == oracle parse + oracle NextMinimizer."""
    cabi, ctx = _ctx()
    text, seqs = make_fastq(3000, 150, 21)
    o = oracle.fastx_parse(text)
    for mode, omode, kw in ((cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, dict(k=21, w=11)),
                            (cabi.MODE_NTHASH, oracle.MODE_NTHASH, dict(k=21)),
                            (cabi.MODE_SYNCMER, oracle.MODE_SYNCMER, dict(k=21, s=11))):
        res = ctx.run_fastx(cabi.make_params(mode, **kw), text)
        ref = oracle.run_batch(o["bases"], o["read_off"], omode, threads=4, **kw)
        assert np.array_equal(res["val"], ref["val"]) and np.array_equal(res["off"], ref["off"])
        assert np.array_equal(res["pos"], ref["pos"]) and np.array_equal(res["status"], ref["status"])
    text, seqs = make_fasta(200, 8)
    o = oracle.fastx_parse(text)
    res = ctx.run_fastx(cabi.make_params(cabi.MODE_MINIMIZER, k=21, w=11), text)
    ref = oracle.run_batch(o["bases"], o["read_off"], oracle.MODE_MINIMIZER, threads=4, k=21, w=11)
    assert np.array_equal(res["val"], ref["val"]) and np.array_equal(res["off"], ref["off"])
    ctx.close()


@pytest.mark.gpu
def test_fxstream_equals_one_shot():
    """The pipelined reader (b200sk_fxstream, two slots / two worker threads) over a text cut every few KB returns,
    chunk after chunk, exactly the records and sketches of the oracle on the whole text; rewind reuses the stream;
    a record larger than the chunk grows the chunk; errors arrive in order."""
    cabi, _ = _ctx()
    cases = [make_fastq(4000, 150, 31, at_quals=True) + (dict(k=21, w=11), 10007),
             make_fastq(900, 150, 32, crlf=True, tail_newline=False) + (dict(k=21, w=11), 64),
             make_fasta(400, 33, blank_lines=True) + (dict(k=15, w=7), 4099),
             make_fastq(50, 150, 34) + (dict(k=21, w=11), 0)]
    stream = None
    for text, seqs, kw, chunk in cases:
        p = cabi.make_params(cabi.MODE_MINIMIZER, **kw)
        o = oracle.fastx_parse(text)
        ref = oracle.run_batch(o["bases"], o["read_off"], oracle.MODE_MINIMIZER, threads=4, **kw)
        for rep in range(2):
            if stream is None or rep == 0:
                if stream is not None:
                    stream.close()
                stream = cabi.FastxStream(p, text, chunk_bytes=chunk)
            else:
                stream.rewind()  # same text again through the same buffers
            nrec, nval, nchunk, pos_in_text = 0, 0, 0, 0
            for c in stream:
                n, t = int(c["info"].n_records), c["total"]
                assert np.array_equal(c["off"], ref["off"][nrec:nrec + n + 1] - ref["off"][nrec])
                assert np.array_equal(c["val"], ref["val"][nval:nval + t])
                assert np.array_equal(c["pos"], ref["pos"][nval:nval + t])
                assert np.array_equal(c["status"], ref["status"][nrec:nrec + n])
                nrec, nval, nchunk = nrec + n, nval + t, nchunk + 1
                pos_in_text += int(c["info"].consumed)
            assert nrec == o["n_records"] == len(seqs) and nval == len(ref["val"])
            assert pos_in_text == len(text)
            assert nchunk > 1 or chunk == 0
            assert stream.next() is None  # stays at the end
        assert stream.kernel_launches > 0
    # empty text: one empty chunk, then the end
    stream.rewind(b"")
    c = stream.next()
    assert c is not None and c["total"] == 0 and int(c["info"].n_records) == 0 and stream.next() is None
    # a broken record in the middle: the chunks before it arrive, then the error, then nothing
    good, _ = make_fastq(300, 150, 35)
    bad = good + b"@broken\nACGT\n+\nII\n" + good
    stream.rewind(bad)
    got = 0
    with pytest.raises(cabi.SketchError) as ei:
        for c in stream:
            got += int(c["info"].n_records)
    assert ei.value.code == cabi.ERR_BAD_FASTQ and got <= 300
    assert stream.next() is None
    stream.rewind(b"not a fastx file\n")
    with pytest.raises(cabi.SketchError) as ei:
        stream.next()
    assert ei.value.code == cabi.ERR_NOT_FASTX
    stream.close()


def make_multiline_fastq(n, seed, width=60, crlf=False, blank_between=False, at_quals=True):
    rng = np.random.default_rng(seed)
    nl = b"\r\n" if crlf else b"\n"
    out, seqs, quals = [], [], []
    for i in range(n):
        L = int(rng.integers(0, 400))
        s = bytes(np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=L)])
        q = bytearray(rng.integers(33, 74, size=L, dtype=np.uint8).tobytes())
        wq = int(rng.integers(7, 90))
        for j in range(0, L, wq):  # '@' and '+' at the start of quality lines
            if at_quals and rng.integers(0, 3) == 0:
                q[j] = ord("@") if rng.integers(0, 2) else ord("+")
        seqs.append(s)
        quals.append(bytes(q))
        out.append(b"@r%d d" % i + nl)
        out += [s[j:j + width] + nl for j in range(0, L, width)]
        out.append(b"+" + (b"r%d" % i if i % 2 else b"") + nl)
        out += [bytes(q[j:j + wq]) + nl for j in range(0, L, wq)]
        if blank_between and rng.integers(0, 3) == 0:
            out.append(nl)
    return b"".join(out), seqs, quals


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(crlf=True), dict(blank_between=True), dict(width=7), dict(at_quals=False)])
def test_multiline_fastq_parity(kw):
    """reader.go:308-345,396-417: records of any line structure; the oracle follows the reference literally."""
    cabi, ctx = _ctx()
    for n, seed in ((1, 1), (40, 2), (1500, 3)):
        text, seqs, quals = make_multiline_fastq(n, seed, **kw)
        for t in (text, text[:-1] if not kw.get("crlf") else text, b"\n\n" + text):
            o = oracle.fastx_parse(t)
            assert o["status"] == 0 and o["n_records"] == n
            info, g = _parse_gpu(ctx, t)
            assert g["format"] == cabi.FASTX_FASTQ and g["consumed"] == len(t)
            _same(g, o, True)
            assert bytes(g["bases"]) == b"".join(seqs)
    # four-line records with blank lines between them (the parallel path hands over to the general rule)
    text, seqs = make_fastq(300, 80, 5)
    recs = text.split(b"\n@")
    t = b"\n\n@".join(recs)
    o = oracle.fastx_parse(t)
    info, g = _parse_gpu(ctx, t)
    assert o["status"] == 0 and o["n_records"] == 300
    _same(g, o, True)
    # the reader mirror joins multi-line quality
    from bio_b200 import fastx
    text, seqs, quals = make_multiline_fastq(200, 9)
    got = list(fastx.Reader(text, ctx=ctx, chunk_bytes=3000))
    assert [r.Seq for r in got] == seqs and [r.Qual for r in got] == quals
    # a quality longer than its sequence at the next record start is ErrBadFASTQFormat; a short one at the end too
    for bad in (b"@a\nAC\nGT\n+\nIII\nIII\n@b\nA\n+\nI\n", b"@a\nAC\nGT\n+\nII\nI\n"):
        assert oracle.fastx_parse(bad)["status"] == oracle.ERR_BAD_FASTQ
        with pytest.raises(cabi.SketchError) as e:
            _parse_gpu(ctx, bad)
        assert e.value.code == cabi.ERR_BAD_FASTQ
    ctx.close()


@pytest.mark.gpu
def test_multiline_fastq_streaming_chunks():
    """consumed / carry-over with records that span many lines: any cut reproduces the whole-file parse"""
    cabi, ctx = _ctx()
    text, seqs, quals = make_multiline_fastq(120, 4, width=25)
    want = oracle.fastx_parse(text)
    for chunk in (500, 2000, 7919):
        got, pos, fmt = [], 0, 0
        while pos < len(text):
            end = min(len(text), pos + chunk)
            final = end == len(text)
            while True:
                info, g = _parse_gpu(ctx, text[pos:end], fmt, final)
                if g["consumed"] or final:
                    break
                end = min(len(text), end + chunk)
                final = end == len(text)
            fmt = g["format"]
            for i in range(g["n_records"]):
                got.append(bytes(g["bases"][int(g["read_off"][i]):int(g["read_off"][i + 1])]))
            pos += g["consumed"] if not final else end - pos
        assert got == seqs and len(got) == want["n_records"]
    ctx.close()


def test_oracle_alphabet_rules():
    # seq/alphabet.go:411-452
    g = oracle.guess_alphabet_less_conservatively
    assert g(b"ACGTNacgt") == "DNAredundant" and g(b"ACGU") == "RNAredundant" and g(b"ACGTRYK") == "DNAredundant"
    assert g(b"ACGURY") == "RNAredundant" and g(b"MKVLAAGIVGLE") == "Protein" and g(b"ACGT!") == "Unlimit"
    assert g(b"") == "Unlimit" and g(b"ACGT" * 3000 + b"!") == "DNAredundant"  # only the first 10 000 letters are looked at
    assert oracle.alphabet_is_valid("DNAredundant", b"ACGTN-.") and not oracle.alphabet_is_valid("DNAredundant", b"ACGTJ")
    assert oracle.alphabet_is_valid("Unlimit", b"!!") and oracle.alphabet_is_valid("Protein", b"MKV*")


@pytest.mark.gpu
def test_alphabet_guess_and_validation():
    """reader.go:430-452: the alphabet comes from the first record, every record is checked against it"""
    cabi, ctx = _ctx()
    from bio_b200 import fastx
    names = fastx.ALPHABET_NAMES
    rng = np.random.default_rng(5)
    cases = [(b"ACGT", {}), (b"ACGTN", {}), (b"ACGU", {}), (b"ACGTRYKM", {}), (b"MKVLAGIE", {}),
             (b"ACGT", {3: b"J", 17: b"!", 18: b"u"}), (b"ACGU", {5: b"T"}), (b"ACGT!", {})]
    for letters, spoil in cases:
        seqs = []
        for i in range(40):
            L = int(rng.integers(1, 300))
            sq = bytearray(np.frombuffer(letters, dtype=np.uint8)[rng.integers(0, len(letters), size=L)].tobytes())
            if i in spoil:
                sq[int(rng.integers(0, L))] = spoil[i][0]
            seqs.append(bytes(sq))
        seqs[0] = letters * 3  # the first record shows every letter of the set
        for fastq in (False, True):
            if fastq:
                text = b"".join(b"@r%d\n" % i + sq + b"\n+\n" + b"I" * len(sq) + b"\n" for i, sq in enumerate(seqs))
            else:
                text = b"".join(b">r%d\n" % i + sq + b"\n" for i, sq in enumerate(seqs))
            info, g = _parse_gpu(ctx, text)
            want_alpha = oracle.guess_alphabet_less_conservatively(seqs[0])
            assert names[g["alphabet"]] == want_alpha
            want_bad = [not oracle.alphabet_is_valid(want_alpha, sq) for sq in seqs]
            assert g["invalid"].astype(bool).tolist() == want_bad
            first = want_bad.index(True) if any(want_bad) else (1 << 64) - 1
            assert g["first_invalid"] == first
            recs = list(fastx.Reader(text, ctx=ctx, chunk_bytes=1500))  # the alphabet travels with the later chunks
            assert [r.Err is not None for r in recs] == want_bad and all(r.Alphabet == want_alpha for r in recs)
    ctx.close()


@pytest.mark.gpu
def test_fxstream_long_run_of_blank_lines_mid_file():
    """More than 1 MiB of blank lines between two records of a file read in small chunks: a chunk that continues the
    file may start with any number of newlines (the format is known by then), so the search for its first record byte
    goes to the end of the chunk -- no record behind the run may be lost."""
    cabi, _ = _ctx()
    a, sa = make_fasta(150, 41)
    b, sb = make_fasta(150, 42)
    text = a + b"\n" * (3 << 19) + b
    kw = dict(k=15, w=7)
    p = cabi.make_params(cabi.MODE_MINIMIZER, **kw)
    o = oracle.fastx_parse(text)
    assert o["n_records"] == 300
    ref = oracle.run_batch(o["bases"], o["read_off"], oracle.MODE_MINIMIZER, threads=4, **kw)
    stream = cabi.FastxStream(p, text, chunk_bytes=64 << 10)
    nrec, nval, consumed = 0, 0, 0
    for c in stream:
        n, t = int(c["info"].n_records), c["total"]
        assert np.array_equal(c["val"], ref["val"][nval:nval + t])
        assert np.array_equal(c["status"], ref["status"][nrec:nrec + n])
        nrec, nval, consumed = nrec + n, nval + t, consumed + int(c["info"].consumed)
    assert nrec == 300 and nval == len(ref["val"]) and consumed == len(text)
    stream.close()
