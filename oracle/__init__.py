"""ctypes loader for the CPU oracle (oracle/sketch_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(bio_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsketch_oracle.so")

OK = 0
ERR_INVALID_K = -1
ERR_SHORT_SEQ = -2
ERR_INVALID_W = -3
ERR_INVALID_S = -4
ERR_ILLEGAL_BASE = -5
ERR_K_OVERFLOW = -6
ERR_INVALID_FRAME = -7
ERR_CODON_TABLE = -8
ERR_TRANSLATE_SHORT = -9
ERR_INVALID_CODON = -10
ERR_INVALID_M = -11
ERR_INVALID_SCALE = -12
ERR_K_TOO_LARGE = -13

SORT_STABLE = 0
SORT_GO14 = 1

MODE_KMER, MODE_NTHASH, MODE_MINIMIZER, MODE_SYNCMER, MODE_PROTEIN, MODE_PROTEIN_MINIMIZER, MODE_SIMHASH = range(7)


def build():
    """Compile the oracle with gcc (no GPU, no reference sources needed)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


class _Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "mode", "k", "w", "s", "canonical", "circular", "codon_table", "frame",
        "alphabet", "sort_policy", "m", "scale")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, f) for f in ("sketch_oracle.c", "fastx_oracle.c")]
        if (not os.path.exists(_LIB_PATH)
                or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs)):
            build()
        L = C.CDLL(_LIB_PATH)
        u8p, u64p, i64p, ip = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint64),
                               C.POINTER(C.c_int64), C.POINTER(C.c_int))
        L.ora_hash_iterator.restype = C.c_int64
        L.ora_hash_iterator.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, ip]
        L.ora_kmer_iterator.restype = C.c_int64
        L.ora_kmer_iterator.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, u64p, ip, i64p]
        for f in (L.ora_minimizer, L.ora_syncmer):
            f.restype = C.c_int64
            f.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, u64p, i64p, ip, ip]
        for f in (L.ora_minimizer_closed, L.ora_syncmer_closed):
            f.restype = C.c_int64
            f.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, i64p, ip]
        L.ora_translate.restype = C.c_int64
        L.ora_translate.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, ip]
        L.ora_wyhash.restype = C.c_uint64
        L.ora_wyhash.argtypes = [u8p, C.c_uint64, C.c_uint64]
        L.ora_protein_iterator.restype = C.c_int64
        L.ora_protein_iterator.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, ip]
        L.ora_protein_minimizer.restype = C.c_int64
        L.ora_protein_minimizer.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            u64p, i64p, ip, ip]
        L.ora_simhash_iterator.restype = C.c_int64
        L.ora_simhash_iterator.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u64p, ip]
        L.ora_pair_lut.restype = None
        L.ora_pair_lut.argtypes = [C.c_int, u8p]
        L.ora_run_batch.restype = C.c_uint64
        L.ora_run_batch.argtypes = [C.POINTER(_Params), u8p, u64p, C.c_uint64, C.c_int,
                                    u64p, C.POINTER(C.c_int32), u64p, u64p,
                                    C.POINTER(C.c_uint32), u64p]
        _lib = L
    return _lib


def _as_u8(seq):
    if isinstance(seq, str):
        seq = seq.encode()
    if isinstance(seq, (bytes, bytearray)):
        a = np.frombuffer(bytes(seq), dtype=np.uint8)
    else:
        a = np.ascontiguousarray(seq, dtype=np.uint8)
    return a


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def hash_iterator(seq, k, canonical=True, circular=False):
    """sketches.NewHashIterator + NextHash loop -> (values, err)."""
    s = _as_u8(seq)
    out = np.zeros(len(s) + max(k, 1) + 2, dtype=np.uint64)
    err = C.c_int(0)
    n = lib().ora_hash_iterator(_p(s, C.c_uint8), len(s), k, int(canonical), int(circular),
                                _p(out, C.c_uint64), C.byref(err))
    return out[:n].copy(), err.value


def kmer_iterator(seq, k, canonical=True, circular=False, alphabet=0):
    """sketches.NewKmerIterator + NextKmer loop -> (codes, err, err_idx)."""
    s = _as_u8(seq)
    out = np.zeros(2 * (len(s) + max(k, 1)) + 4, dtype=np.uint64)
    err = C.c_int(0)
    eidx = C.c_int64(-1)
    n = lib().ora_kmer_iterator(_p(s, C.c_uint8), len(s), k, int(canonical), int(circular), alphabet,
                                _p(out, C.c_uint64), C.byref(err), C.byref(eidx))
    return out[:n].copy(), err.value, eidx.value


def _sketch(fn, seq, k, x, circular, policy):
    s = _as_u8(seq)
    cap = len(s) + max(k, 1) + 2
    val = np.zeros(cap, dtype=np.uint64)
    idx = np.zeros(cap, dtype=np.int64)
    err = C.c_int(0)
    tie = C.c_int(0)
    n = fn(_p(s, C.c_uint8), len(s), k, x, int(circular), policy,
           _p(val, C.c_uint64), _p(idx, C.c_int64), C.byref(err), C.byref(tie))
    return val[:n].copy(), idx[:n].copy(), err.value, bool(tie.value)


def minimizer(seq, k, w, circular=False, policy=SORT_STABLE):
    """sketches.NewMinimizerSketch + NextMinimizer/Index loop -> (vals, idxs, err, first_window_tie)."""
    return _sketch(lib().ora_minimizer, seq, k, w, circular, policy)


def syncmer(seq, k, s, circular=False, policy=SORT_STABLE):
    """sketches.NewSyncmerSketch + NextSyncmer/Index loop -> (vals, idxs, err, first_window_tie)."""
    return _sketch(lib().ora_syncmer, seq, k, s, circular, policy)


def _closed(fn, seq, k, x, circular):
    s = _as_u8(seq)
    cap = len(s) + max(k, 1) + 2
    val = np.zeros(cap, dtype=np.uint64)
    idx = np.zeros(cap, dtype=np.int64)
    err = C.c_int(0)
    n = fn(_p(s, C.c_uint8), len(s), k, x, int(circular),
           _p(val, C.c_uint64), _p(idx, C.c_int64), C.byref(err))
    return val[:n].copy(), idx[:n].copy(), err.value


def minimizer_closed(seq, k, w, circular=False):
    return _closed(lib().ora_minimizer_closed, seq, k, w, circular)


def syncmer_closed(seq, k, s, circular=False):
    return _closed(lib().ora_syncmer_closed, seq, k, s, circular)


def translate(seq, table=1, frame=1, trim=False, clean=False, allow_unknown=True):
    """seq.Seq.Translate (markInitCodonAsM=false) -> (aa bytes, err)."""
    s = _as_u8(seq)
    out = np.zeros(len(s) // 3 + 4, dtype=np.uint8)
    err = C.c_int(0)
    n = lib().ora_translate(_p(s, C.c_uint8), len(s), table, frame, int(trim), int(clean),
                            int(allow_unknown), _p(out, C.c_uint8), C.byref(err))
    return out[:n].tobytes(), err.value


def wyhash(data, seed=1):
    s = _as_u8(data)
    return int(lib().ora_wyhash(_p(s, C.c_uint8), len(s), seed))


def protein_iterator(seq, k, table=1, frame=1):
    """sketches.NewProteinIterator + Next loop -> (hashes, err)."""
    s = _as_u8(seq)
    out = np.zeros(len(s) // 3 + 4, dtype=np.uint64)
    err = C.c_int(0)
    n = lib().ora_protein_iterator(_p(s, C.c_uint8), len(s), k, table, frame,
                                   _p(out, C.c_uint64), C.byref(err))
    return out[:n].copy(), err.value


def protein_minimizer(seq, k, w, table=1, frame=1, protein_input=False, policy=SORT_STABLE):
    """sketches.NewProteinMinimizerSketch + Next/Index loop -> (vals, idxs, err, first_window_tie)."""
    s = _as_u8(seq)
    cap = len(s) + 4
    val = np.zeros(cap, dtype=np.uint64)
    idx = np.zeros(cap, dtype=np.int64)
    err, tie = C.c_int(0), C.c_int(0)
    n = lib().ora_protein_minimizer(_p(s, C.c_uint8), len(s), k, table, frame, w, int(protein_input), policy,
                                    _p(val, C.c_uint64), _p(idx, C.c_int64), C.byref(err), C.byref(tie))
    return val[:n].copy(), idx[:n].copy(), err.value, bool(tie.value)


def simhash_iterator(seq, k, m, scale, canonical=True, circular=False):
    """sketches.NewSimHashIterator + NextSimHash loop -> (codes, err)."""
    s = _as_u8(seq)
    out = np.zeros(len(s) + max(k, 1) + 2, dtype=np.uint64)
    err = C.c_int(0)
    n = lib().ora_simhash_iterator(_p(s, C.c_uint8), len(s), k, m, scale, int(canonical), int(circular),
                                   _p(out, C.c_uint64), C.byref(err))
    return out[:n].copy(), err.value


def pair_lut(alphabet=0):
    lut = np.zeros(256, dtype=np.uint8)
    lib().ora_pair_lut(alphabet, _p(lut, C.c_uint8))
    return lut


def run_batch(bases, off, mode, k, w=0, s=0, canonical=True, circular=False, codon_table=1,
              frame=1, alphabet=0, sort_policy=SORT_STABLE, threads=1, want_output=True,
              want_pos=True, m=0, scale=1):
    """The reference's per-record pull loop over a concatenated batch.

    Returns dict(counts, status, off, val, pos, ties, checksum); val/pos None if
    want_output is False (timing mode: one pass, no allocation of outputs).
    """
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = len(off) - 1
    p = _Params(mode, k, w, s, int(canonical), int(circular), codon_table, frame, alphabet, sort_policy, m, scale)
    counts = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    ties = C.c_uint64(0)
    L = lib()
    cks = L.ora_run_batch(C.byref(p), _p(bases, C.c_uint8), _p(off, C.c_uint64), n, threads,
                          _p(counts, C.c_uint64), _p(status, C.c_int32), None, None, None,
                          C.byref(ties))
    res = dict(counts=counts, status=status, ties=int(ties.value), checksum=int(cks),
               off=None, val=None, pos=None)
    if want_output:
        out_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(counts, out=out_off[1:])
        total = int(out_off[-1])
        val = np.zeros(max(total, 1), dtype=np.uint64)
        pos = np.zeros(max(total, 1), dtype=np.uint32) if want_pos else None
        L.ora_run_batch(C.byref(p), _p(bases, C.c_uint8), _p(off, C.c_uint64), n, threads,
                        None, None, _p(out_off, C.c_uint64), _p(val, C.c_uint64),
                        _p(pos, C.c_uint32) if want_pos else None, None)
        res.update(off=out_off, val=val[:total], pos=pos[:total] if want_pos else None)
    return res


# ---------------------------------------------------------------- record feeder oracle (fastx_oracle.c)
FASTX_FASTA, FASTX_FASTQ = 1, 2
ERR_NOT_FASTX, ERR_BAD_FASTQ = -20, -21


def fastx_parse(text):
    """seqio/fastx.Reader.Read until EOF over `text` (bytes): the records' sequences packed, their offsets,
    the text offset of every record delimiter / quality line, and the header length."""
    L = lib()
    u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    f = L.ora_fastx_parse
    f.restype = C.c_int
    f.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_int), u64p, u64p, C.POINTER(u8p), C.POINTER(u64p),
                  C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u64p)]
    L.ora_fastx_free.argtypes = [C.c_void_p]
    fmt = C.c_int(0)
    n_rec, n_bases = C.c_uint64(0), C.c_uint64(0)
    bases = u8p()
    ro, rc, qo, nl = u64p(), u64p(), u64p(), u64p()
    text = bytes(text)
    st = f(text, len(text), C.byref(fmt), C.byref(n_rec), C.byref(n_bases), C.byref(bases), C.byref(ro),
           C.byref(rc), C.byref(qo), C.byref(nl))
    n, nb = n_rec.value, n_bases.value
    out = {
        "status": st, "format": fmt.value, "n_records": n,
        "bases": np.ctypeslib.as_array(bases, shape=(max(nb, 1),))[:nb].copy(),
        "read_off": np.ctypeslib.as_array(ro, shape=(n + 1,)).copy(),
        "rec_off": np.ctypeslib.as_array(rc, shape=(n + 1,))[:n].copy(),
        "qual_off": np.ctypeslib.as_array(qo, shape=(n + 1,))[:n].copy(),
        "name_len": np.ctypeslib.as_array(nl, shape=(n + 1,))[:n].copy(),
    }
    for p in (bases, ro, rc, qo, nl):
        L.ora_fastx_free(C.cast(p, C.c_void_p))
    return out


# ---------------------------------------------------------------- alphabet guess / validation (numpy restatement)
# seq/alphabet.go:353-399: letters + gap + ambiguous of every alphabet the reader can guess
ALPHABET_LETTERS = {
    "DNA": b"acgtACGT -.nN",
    "DNAredundant": b"acgtryswkmbdhvACGTRYSWKMBDHV -.nN",
    "RNA": b"acguACGU -.nN",
    "RNAredundant": b"acguryswkmbdhvACGURYSWKMBDHV -.nN",
    "Protein": b"abcdefghijklmnopqrstuvwyzABCDEFGHIJKLMNOPQRSTUVWYZ -xX*_.",
}


def guess_alphabet_less_conservatively(seq_bytes, threshold=10000):
    """seq.GuessAlphabetLessConservatively (seq/alphabet.go:411-452) over the first record's sequence."""
    s = bytes(seq_bytes)
    if len(s) == 0:
        return "Unlimit"
    present = set(s[:threshold] if threshold and len(s) > threshold else s)
    if any(b >= 128 for b in present):
        return "Unlimit"
    for name in ("DNA", "RNA", "DNAredundant", "RNAredundant", "Protein"):
        if present <= set(ALPHABET_LETTERS[name]):
            return {"DNA": "DNAredundant", "RNA": "RNAredundant"}.get(name, name)
    return "Unlimit"


def alphabet_is_valid(name, seq_bytes):
    """Alphabet.IsValid (seq/alphabet.go:234-300): every letter among the alphabet's letters; Unlimit takes all."""
    if name == "Unlimit" or len(seq_bytes) == 0:
        return True
    ok = np.zeros(256, dtype=bool)
    ok[np.frombuffer(ALPHABET_LETTERS[name], dtype=np.uint8)] = True
    return bool(ok[np.frombuffer(bytes(seq_bytes), dtype=np.uint8)].all())
