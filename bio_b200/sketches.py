"""Host-side mirror of the reference's `sketches` package interface for the hot path
(shenwei356/bio sketches/iterator.go, sketches/sketch.go, sketches/iterator-protein.go): the same
constructor names, argument meaning and error behaviour, with the work done by libb200sketch.so on the GPU.

The reference toolchain (Go) is absent from the build image, so this module plays the part of the cgo shim
(go/sketchesgpu) in Python: constructors sketch the sequence through the C ABI and `Next()` / `Index()` replay
the result.  One sequence per call is the reference's shape; `Batch` is the shape that performs (one C-ABI
call for a whole chunk of records -- see INTEGRATION.md).

There is no CPU fallback: every constructor needs a CUDA device.
"""
import numpy as np

from . import _cabi as cabi

# ---- errors (sketches/iterator.go:34-53, sketches/sketch.go:32-42) ---------------------------------------


class SketchesError(Exception):
    code = 0


class ErrInvalidK(SketchesError):
    code = cabi.ERR_INVALID_K

    def __init__(self):
        super().__init__("sketches: invalid k-mer size")


class ErrShortSeq(SketchesError):
    code = cabi.ERR_SHORT_SEQ

    def __init__(self):
        super().__init__("sketches: sequence too short")


class ErrIllegalBase(SketchesError):
    code = cabi.ERR_ILLEGAL_BASE

    def __init__(self):
        super().__init__("sketches: illegal base")


class ErrKTooLarge(SketchesError):
    code = cabi.ERR_K_TOO_LARGE

    def __init__(self):
        super().__init__("sketches: k-mer size is too large")


class ErrKOverflow(SketchesError):
    """kmers.ErrKOverflow: what kmers.Encode returns for k > 32 (iterator.go:742)."""
    code = cabi.ERR_K_OVERFLOW

    def __init__(self):
        super().__init__("unikmer: k-mer size (1-32) overflow")


class ErrInvalidM(SketchesError):
    code = cabi.ERR_INVALID_M

    def __init__(self):
        super().__init__("sketches: invalid m-mer size, should be in range of [4, k]")


class ErrInvalidScale(SketchesError):
    code = cabi.ERR_INVALID_SCALE

    def __init__(self):
        super().__init__("sketches: invalid scale, should be in range of [1, k-m+1]")


class ErrInvalidS(SketchesError):
    code = cabi.ERR_INVALID_S

    def __init__(self):
        super().__init__("kmers: invalid s-mer size")


class ErrInvalidW(SketchesError):
    code = cabi.ERR_INVALID_W

    def __init__(self):
        super().__init__("kmers: invalid minimimzer window")


_BY_CODE = {c.code: c for c in (ErrInvalidK, ErrShortSeq, ErrIllegalBase, ErrKTooLarge, ErrKOverflow, ErrInvalidS,
                                ErrInvalidW, ErrInvalidM, ErrInvalidScale)}


def _raise(code):
    if code in _BY_CODE:
        raise _BY_CODE[code]()
    raise cabi.SketchError(code)


# ---- seq.Seq (seq/seq.go:29-34): only what the path touches ----------------------------------------------
DNA, DNAredundant, RNA, RNAredundant, Unlimit, Protein = "DNA", "DNAredundant", "RNA", "RNAredundant", "Unlimit", "Protein"
_ALPHA = {DNAredundant: cabi.ALPHABET_DNA_REDUNDANT, DNA: cabi.ALPHABET_DNA, RNAredundant: cabi.ALPHABET_RNA_REDUNDANT,
          RNA: cabi.ALPHABET_RNA, Unlimit: cabi.ALPHABET_UNLIMIT, Protein: cabi.ALPHABET_PROTEIN}


class Seq:
    def __init__(self, alphabet, seq):
        self.Alphabet = alphabet
        self.Seq = seq.encode() if isinstance(seq, str) else bytes(seq)


def NewSeq(alphabet, seq):
    return Seq(alphabet, seq)


_ctx = None


def default_context():
    global _ctx
    if _ctx is None:
        _ctx = cabi.Context(0)
    return _ctx


def _run_one(seq_, **kw):
    s = seq_
    p = cabi.make_params(alphabet=_ALPHA.get(s.Alphabet, cabi.ALPHABET_DNA_REDUNDANT), **kw)
    rc = cabi.lib().b200sk_check_params(__import__("ctypes").byref(p))
    if rc != 0:
        _raise(rc)
    bases = np.frombuffer(s.Seq, dtype=np.uint8)
    off = np.array([0, len(bases)], dtype=np.uint64)
    res = default_context().run(p, bases, off)
    return res["val"], res["pos"], int(res["status"][0])


class Iterator:
    """sketches.Iterator (iterator.go:60): k-mer code or ntHash iterator."""

    def __init__(self, val, pos, deferred=0):
        self._val, self._pos, self._i, self._deferred = val, pos, 0, deferred

    def NextKmer(self):
        """(code, ok, err) -- iterator.go:708; an illegal base surfaces after the codes before it."""
        if self._i >= len(self._val):
            if self._deferred:
                d, self._deferred = self._deferred, 0
                return 0, False, _BY_CODE[d]()
            return 0, False, None
        v = int(self._val[self._i])
        self._i += 1
        return v, True, None

    def NextHash(self):
        """(hash, ok) -- iterator.go:658."""
        if self._i >= len(self._val):
            return 0, False
        v = int(self._val[self._i])
        self._i += 1
        return v, True

    def NextSimHash(self):
        """(simhash, ok) -- iterator.go:191."""
        return self.NextHash()

    def Next(self):
        """(code, ok, err) -- iterator.go:762."""
        return self.NextKmer()

    def Index(self):
        """0-based index of the last element (iterator.go:776)."""
        return int(self._pos[self._i - 1])


class Sketch:
    """sketches.Sketch (sketch.go:45): minimizer / syncmer iterator."""

    def __init__(self, val, pos, minimizer):
        self._val, self._pos, self._i, self._minimizer = val, pos, 0, minimizer

    def Next(self):
        if self._i >= len(self._val):
            return 0, False
        v = int(self._val[self._i])
        self._i += 1
        return v, True

    NextMinimizer = Next
    NextSyncmer = Next

    def Index(self):
        return int(self._pos[self._i - 1])


class ProteinIterator(Sketch):
    """sketches.ProteinIterator (iterator-protein.go:35)."""


def NewKmerIterator(s, k, canonical, circular):
    """iterator.go:668"""
    val, pos, st = _run_one(s, mode=cabi.MODE_KMER, k=k, canonical=canonical, circular=circular)
    if st not in (0, cabi.ERR_ILLEGAL_BASE):
        _raise(st)
    return Iterator(val, pos, deferred=st)


def NewHashIterator(s, k, canonical, circular):
    """iterator.go:615"""
    val, pos, st = _run_one(s, mode=cabi.MODE_NTHASH, k=k, canonical=canonical, circular=circular)
    if st:
        _raise(st)
    return Iterator(val, pos)


def NewSimHashIterator(s, k, m, scale, canonical, circular):
    """iterator.go:113"""
    val, pos, st = _run_one(s, mode=cabi.MODE_SIMHASH, k=k, m=m, scale=scale, canonical=canonical, circular=circular)
    if st:
        _raise(st)
    return Iterator(val, pos)


def NewMinimizerSketch(S, k, w, circular):
    """sketch.go:85"""
    if w > (1 << 31) - 1:
        raise ErrInvalidW()
    val, pos, st = _run_one(S, mode=cabi.MODE_MINIMIZER, k=k, w=w, circular=circular)
    if st:
        _raise(st)
    return Sketch(val, pos, True)


def NewSyncmerSketch(S, k, s, circular):
    """sketch.go:142"""
    val, pos, st = _run_one(S, mode=cabi.MODE_SYNCMER, k=k, s=s, circular=circular)
    if st:
        _raise(st)
    return Sketch(val, pos, False)


def NewProteinIterator(s, k, codonTable, frame):
    """iterator-protein.go:46"""
    val, pos, st = _run_one(s, mode=cabi.MODE_PROTEIN, k=k, codon_table=codonTable, frame=frame)
    if st:
        _raise(st)
    return ProteinIterator(val, pos, False)


class ProteinMinimizerSketch(Sketch):
    """sketches.ProteinMinimizerSketch (sketch-protein.go:32)."""


def NewProteinMinimizerSketch(S, k, codonTable, frame, w):
    """sketch-protein.go:62"""
    if w > (1 << 31) - 1:
        raise ErrInvalidW()
    val, pos, st = _run_one(S, mode=cabi.MODE_PROTEIN_MINIMIZER, k=k, codon_table=codonTable, frame=frame, w=w)
    if st:
        _raise(st)
    return ProteinMinimizerSketch(val, pos, True)


# ---- the batch shape -------------------------------------------------------------------------------------
class Batch:
    """Records of one fastx chunk packed into concatenated bases + offsets; one C-ABI call sketches them all.
    `result.iterator(i)` replays read i like the reference's per-record loop."""

    def __init__(self, ctx=None, alphabet=DNAredundant):
        self.ctx = ctx or default_context()
        self.alphabet = alphabet
        self._chunks, self._off = [], [0]
        self._max = 0

    def Add(self, seq):
        b = seq.Seq if isinstance(seq, Seq) else (seq.encode() if isinstance(seq, str) else bytes(seq))
        self._chunks.append(b)
        self._off.append(self._off[-1] + len(b))
        self._max = max(self._max, len(b))

    def _run(self, **kw):
        p = cabi.make_params(alphabet=_ALPHA[self.alphabet], max_read_len=self._max, **kw)
        rc = cabi.lib().b200sk_check_params(__import__("ctypes").byref(p))
        if rc != 0:
            _raise(rc)
        bases = np.frombuffer(b"".join(self._chunks), dtype=np.uint8)
        return BatchResult(self.ctx.run(p, bases, np.array(self._off, dtype=np.uint64)), kw["mode"])

    def KmerIterator(self, k, canonical, circular):
        return self._run(mode=cabi.MODE_KMER, k=k, canonical=canonical, circular=circular)

    def HashIterator(self, k, canonical, circular):
        return self._run(mode=cabi.MODE_NTHASH, k=k, canonical=canonical, circular=circular)

    def SimHashIterator(self, k, m, scale, canonical, circular):
        return self._run(mode=cabi.MODE_SIMHASH, k=k, m=m, scale=scale, canonical=canonical, circular=circular)

    def MinimizerSketch(self, k, w, circular):
        return self._run(mode=cabi.MODE_MINIMIZER, k=k, w=w, circular=circular)

    def SyncmerSketch(self, k, s, circular):
        return self._run(mode=cabi.MODE_SYNCMER, k=k, s=s, circular=circular)

    def ProteinIterator(self, k, codonTable, frame):
        return self._run(mode=cabi.MODE_PROTEIN, k=k, codon_table=codonTable, frame=frame)

    def ProteinMinimizerSketch(self, k, codonTable, frame, w):
        return self._run(mode=cabi.MODE_PROTEIN_MINIMIZER, k=k, codon_table=codonTable, frame=frame, w=w)

    def ProteinFrames(self, k, codonTable):
        """Six NewProteinIterator loops -- frame 1, 2, 3, -1, -2, -3 -- over every record through ONE call
        (b200sk_run_frames; go/sketchesgpu: Batch.ProteinFrames).  Returns six BatchResults in that order."""
        p = cabi.make_params(mode=cabi.MODE_PROTEIN, k=k, codon_table=codonTable, frame=1, want_pos=False,
                             alphabet=_ALPHA[self.alphabet], max_read_len=self._max)
        rc = cabi.lib().b200sk_check_params(__import__("ctypes").byref(p))
        if rc != 0:
            _raise(rc)
        bases = np.frombuffer(b"".join(self._chunks), dtype=np.uint8)
        out = []
        for res in self.ctx.run_frames(p, bases, np.array(self._off, dtype=np.uint64)):
            # Index() of a dense iterator is the running position
            res["pos"] = np.concatenate([np.arange(int(res["off"][i + 1] - res["off"][i]), dtype=np.uint32)
                                         for i in range(len(res["status"]))]) if res["total"] else np.zeros(0, np.uint32)
            out.append(BatchResult(res, cabi.MODE_PROTEIN))
        return out


class BatchResult:
    def __init__(self, res, mode):
        self.res, self.mode = res, mode

    def __len__(self):
        return len(self.res["status"])

    def iterator(self, i):
        st = int(self.res["status"][i])
        lo, hi = int(self.res["off"][i]), int(self.res["off"][i + 1])
        val, pos = self.res["val"][lo:hi], self.res["pos"][lo:hi]
        if self.mode == cabi.MODE_KMER:
            if st not in (0, cabi.ERR_ILLEGAL_BASE):
                _raise(st)
            return Iterator(val, pos, deferred=st)
        if st:
            _raise(st)
        if self.mode in (cabi.MODE_NTHASH, cabi.MODE_SIMHASH):
            return Iterator(val, pos)
        return Sketch(val, pos, self.mode == cabi.MODE_MINIMIZER)
