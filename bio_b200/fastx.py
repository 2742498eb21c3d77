"""Host-side mirror of the reference's record stream for the sketching path (shenwei356/bio
seqio/fastx/reader.go, seqio/fastx/records.go): `Reader.Read()` yields `Record`s with the reference's field
names, the records being found on the GPU a chunk of text at a time (b200sk_fastx_parse_device) instead of one
`Read()` at a time on the CPU.

Shape that performs: `Reader.batches()` hands out (FastxInfo, text chunk) pairs whose packed bases and
offsets are already resident in HBM -- feed `info.d_bases / info.d_read_off` straight to
`Context.run_device`, or call `Context.run_fastx` to go text -> sketches in one C-ABI call; `SketchStream`
(b200sk_fxstream) does that for a whole text with the chunks pipelined -- the `ChunkChan` shape
(seqio/fastx/reader.go:556-603).

There is no CPU fallback: the reader needs a CUDA device.
"""
import io

import numpy as np

from . import _cabi as cabi

SketchStream = cabi.FastxStream  # for chunk in SketchStream(params, text): chunk["val"], chunk["pos"], chunk["off"]


class ErrNotFASTXFormat(Exception):  # seqio/fastx/reader.go:16
    def __init__(self):
        super().__init__("fastx: invalid FASTA/Q format")


class ErrBadFASTQFormat(Exception):  # seqio/fastx/reader.go:19
    def __init__(self):
        super().__init__("fastx: bad fastq format")


ALPHABET_NAMES = {cabi.ALPHABET_DNA_REDUNDANT: "DNAredundant", cabi.ALPHABET_DNA: "DNA",
                  cabi.ALPHABET_RNA_REDUNDANT: "RNAredundant", cabi.ALPHABET_RNA: "RNA",
                  cabi.ALPHABET_UNLIMIT: "Unlimit", cabi.ALPHABET_PROTEIN: "Protein"}  # seq/alphabet.go:353-399


class Record:
    """seqio/fastx/records.go:13-18 (ID, Name, Desc, Seq); Seq.Seq / Seq.Qual flattened to seq / qual.
    Err: what Read() returns next to the record in the reference -- None, or the alphabet check's complaint
    (parseRecord, reader.go:450-452: `seq: invalid <alphabet> letter`)."""
    __slots__ = ("ID", "Name", "Desc", "Seq", "Qual", "Alphabet", "Err")

    def __init__(self, name, seq, qual, alphabet=None, err=None):
        self.Name = name
        self.ID, self.Desc = parse_head_id_and_desc(name)
        self.Seq = seq
        self.Qual = qual
        self.Alphabet = alphabet
        self.Err = err


def parse_head_id_and_desc(head):
    """parseHeadIDAndDesc with the default ID regexp (reader.go:486-525): ID up to the first blank or tab,
    description after the run of blanks/tabs that follows."""
    i_tab, i_space = head.find(b"\t"), head.find(b" ")
    if i_space >= 0:
        i = i_space if not (0 <= i_tab < i_space) else i_tab
    elif i_tab >= 0:
        i = i_tab
    else:
        return head, b""
    j = i + 1
    while j < len(head) and head[j] in b" \t":
        j += 1
    return head[:i], head[j:]


def _line(text, start, end_excl):
    """text[start:end_excl] without its '\\n' and one trailing '\\r' (dropCR, reader.go:535-541)."""
    b = text[start:end_excl]
    if b.endswith(b"\n"):
        b = b[:-1]
    if b.endswith(b"\r"):
        b = b[:-1]
    return b


class Reader:
    """fastx.NewReader(t, file, idRegexp) + Read() (reader.go:130, :233), reading `chunk_bytes` of text per
    GPU call.  `source` is a path, bytes, or a binary file object."""

    def __init__(self, source, ctx=None, chunk_bytes=64 << 20):
        import torch
        self._ctx = ctx or cabi.Context(0)
        self._dev = torch.device("cuda", self._ctx.device)
        if isinstance(source, (bytes, bytearray, memoryview)):
            self._fh = io.BytesIO(bytes(source))
        elif hasattr(source, "read"):
            self._fh = source
        else:
            self._fh = open(source, "rb")
        self._chunk = int(chunk_bytes)
        self._carry = b""
        self._eof = False
        self._format = 0
        self._alphabet = None  # guessed from the first record of the file (reader.go:430-435), then kept
        self.IsFastq = False
        self._pending = iter(())

    def batches(self):
        """Yield (info, text) per chunk: info.d_bases / d_read_off hold the chunk's records in HBM."""
        import torch
        while not self._eof or self._carry:
            fresh = b"" if self._eof else self._fh.read(self._chunk)
            if not fresh:
                self._eof = True
            text = self._carry + fresh
            if not text:
                return
            n = len(text)
            host = np.zeros((n + 15) // 16 * 16 + 16, dtype=np.uint8)
            host[:n] = np.frombuffer(text, dtype=np.uint8)
            d_text = torch.from_numpy(host).to(self._dev)
            try:
                fmt = self._format | ((self._alphabet + 1) << 8 if self._alphabet is not None else 0)
                info = self._ctx.fastx_parse_device(d_text, n, fmt, final=self._eof)
            except cabi.SketchError as e:
                if e.code == cabi.ERR_NOT_FASTX:
                    raise ErrNotFASTXFormat() from None
                if e.code == cabi.ERR_BAD_FASTQ:
                    raise ErrBadFASTQFormat() from None
                raise
            self._format = int(info.format) or self._format
            if self._alphabet is None and int(info.n_records):
                self._alphabet = int(info.alphabet)
            self.IsFastq = self._format == cabi.FASTX_FASTQ
            used = int(info.consumed)
            if not self._eof and used == 0 and len(fresh) == 0:
                raise ErrBadFASTQFormat()
            self._carry = text[used:] if not self._eof else b""
            if not self._eof and used == 0:
                # no complete record yet: read more before parsing again
                self._chunk *= 2
                continue
            yield info, text
            if self._eof:
                return

    def _records(self):
        for info, text in self.batches():
            r = self._ctx.fastx_fetch(info)
            n = r["n_records"]
            ro, rc, qo = r["read_off"], r["rec_off"], r["qual_off"]
            bases = r["bases"].tobytes()
            alpha = ALPHABET_NAMES.get(r["alphabet"], "Unlimit")
            for i in range(n):
                start = int(rc[i])
                nl = text.find(b"\n", start)
                name = _line(text, start + 1, (nl + 1) if nl >= 0 else len(text))
                seq = bases[int(ro[i]):int(ro[i + 1])]
                qual = b""
                if self.IsFastq:
                    q = int(qo[i])
                    qual = text[q:q + len(seq)]
                    if b"\n" in qual or b"\r" in qual:  # multi-line quality: its lines joined (reader.go:403-410)
                        parts, have = [], 0
                        while have < len(seq) and q < len(text):
                            e = text.find(b"\n", q)
                            e = len(text) if e < 0 else e + 1
                            ln = _line(text, q, e)
                            parts.append(ln)
                            have += len(ln)
                            q = e
                        qual = b"".join(parts)
                err = "seq: invalid %s letter" % alpha if r["invalid"][i] else None
                yield Record(name, seq, qual, alpha, err)

    def Read(self):
        """One record, or raises EOFError (io.EOF)."""
        try:
            return next(self._pending)
        except StopIteration:
            pass
        if getattr(self, "_gen", None) is None:
            self._gen = self._records()
        try:
            return next(self._gen)
        except StopIteration:
            raise EOFError("EOF") from None

    def __iter__(self):
        while True:
            try:
                yield self.Read()
            except EOFError:
                return
