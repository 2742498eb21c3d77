// Package sketchesgpu is the cgo shim that puts libb200sketch.so behind the method set of
// github.com/shenwei356/bio/sketches (Iterator / Sketch / ProteinIterator): constructors with the same
// names, arguments and errors, Next()/Index() with the same meaning.
//
// The reference API is a pull iterator over ONE sequence; a cgo call per element (~100 ns) would be an order
// of magnitude slower than the Go loop it replaces, so the shim is batch-shaped underneath:
//
//	b := sketchesgpu.NewBatch(ctx)            // pinned staging buffer, C-owned
//	for chunk := range reader.ChunkChan(...)  // seqio/fastx/reader.go:562
//	    for _, rec := range chunk.Data { b.Add(rec.Seq) }
//	res, _ := b.MinimizerSketch(k, w, false)  // ONE cgo call: the whole batch is sketched on the GPU
//	for i := range chunk.Data {
//	    sk := res.Sketch(i)                   // *sketchesgpu.Sketch with the reference's method set
//	    for { v, ok := sk.Next(); if !ok { break }; _ = sk.Index() }
//	}
//
// NOTE: this file is authored against include/b200sketch.h but has NOT been compiled: the build image has
// no Go toolchain (INTEGRATION.md).  The Python mirror (bio_b200/sketches.py) exercises the same C ABI and
// is what the parity tests run.
package sketchesgpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../bio_b200/lib -lb200sketch -Wl,-rpath,${SRCDIR}/../../bio_b200/lib
#include <stdlib.h>
#include <string.h>
#include "b200sketch.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"unsafe"

	"github.com/shenwei356/bio/seq"
)

// The reference's exported errors (sketches/iterator.go:34-53, sketches/sketch.go:32-42).
var (
	ErrInvalidK    = fmt.Errorf("sketches: invalid k-mer size")
	ErrShortSeq    = fmt.Errorf("sketches: sequence too short")
	ErrIllegalBase = errors.New("sketches: illegal base")
	ErrKTooLarge   = fmt.Errorf("sketches: k-mer size is too large")
	ErrInvalidM    = fmt.Errorf("sketches: invalid m-mer size, should be in range of [4, k]")
	ErrInvalidScale = fmt.Errorf("sketches: invalid scale, should be in range of [1, k-m+1]")
	ErrInvalidS    = fmt.Errorf("kmers: invalid s-mer size")
	ErrInvalidW    = fmt.Errorf("kmers: invalid minimimzer window")
)

func codeToError(rc C.int) error {
	switch rc {
	case C.B200SK_OK:
		return nil
	case C.B200SK_ERR_INVALID_K:
		return ErrInvalidK
	case C.B200SK_ERR_SHORT_SEQ:
		return ErrShortSeq
	case C.B200SK_ERR_INVALID_W:
		return ErrInvalidW
	case C.B200SK_ERR_INVALID_S:
		return ErrInvalidS
	case C.B200SK_ERR_ILLEGAL_BASE:
		return ErrIllegalBase
	case C.B200SK_ERR_K_TOO_LARGE:
		return ErrKTooLarge
	case C.B200SK_ERR_INVALID_M:
		return ErrInvalidM
	case C.B200SK_ERR_INVALID_SCALE:
		return ErrInvalidScale
	default:
		return errors.New(C.GoString(C.b200sk_strerror(rc)))
	}
}

// Context owns one GPU (one per goroutine; not goroutine-safe, like the reference's iterators).
type Context struct{ h *C.b200sk_ctx }

// NewContext binds CUDA device `device`.  There is no CPU fallback: without a device this fails.
func NewContext(device int) (*Context, error) {
	var h *C.b200sk_ctx
	if rc := C.b200sk_create(&h, C.int(device)); rc != 0 {
		return nil, codeToError(rc)
	}
	return &Context{h}, nil
}

// Close releases the device buffers.
func (c *Context) Close() { C.b200sk_destroy(c.h); c.h = nil }

// Batch packs record.Seq.Seq bytes into a C-owned pinned buffer (cgo must not let C keep Go pointers).
type Batch struct {
	ctx    *Context
	bases  unsafe.Pointer // pinned, capBases bytes
	nBases int
	cap    int
	off    []C.uint64_t // read offsets, len = reads+1
	maxLen int
	alpha  C.int32_t
}

// NewBatch allocates a staging buffer of capBytes (grown on demand).
func NewBatch(ctx *Context, capBytes int) *Batch {
	return &Batch{ctx: ctx, bases: C.b200sk_alloc_pinned(C.size_t(capBytes)), cap: capBytes,
		off: []C.uint64_t{0}, alpha: C.B200SK_ALPHABET_DNA_REDUNDANT}
}

// Add appends one sequence (the bytes are copied: fastx.Reader reuses its buffer, reader.go:229-232).
func (b *Batch) Add(s *seq.Seq) {
	n := len(s.Seq)
	if b.nBases+n+64 > b.cap {
		ncap := 2*(b.nBases+n) + 64
		nb := C.b200sk_alloc_pinned(C.size_t(ncap))
		C.memcpy(nb, b.bases, C.size_t(b.nBases))
		C.b200sk_free_pinned(b.bases)
		b.bases, b.cap = nb, ncap
	}
	if n > 0 {
		C.memcpy(unsafe.Add(b.bases, b.nBases), unsafe.Pointer(&s.Seq[0]), C.size_t(n))
	}
	b.nBases += n
	b.off = append(b.off, C.uint64_t(b.nBases))
	if n > b.maxLen {
		b.maxLen = n
	}
	switch s.Alphabet {
	case seq.DNA:
		b.alpha = C.B200SK_ALPHABET_DNA
	case seq.RNA:
		b.alpha = C.B200SK_ALPHABET_RNA
	case seq.RNAredundant:
		b.alpha = C.B200SK_ALPHABET_RNA_REDUNDANT
	case seq.Unlimit:
		b.alpha = C.B200SK_ALPHABET_UNLIMIT
	case seq.Protein:
		b.alpha = C.B200SK_ALPHABET_PROTEIN
	}
}

// Reset empties the batch, keeping the buffer.
func (b *Batch) Reset() { b.nBases, b.off, b.maxLen = 0, b.off[:1], 0 }

// Free releases the pinned buffer.
func (b *Batch) Free() { C.b200sk_free_pinned(b.bases); b.bases = nil }

// Result holds the library-owned output arrays of one run (valid until the next run on the context).
type Result struct {
	val    []uint64
	pos    []uint32
	off    []uint64
	status []int32
}

func (b *Batch) run(p *C.b200sk_params) (*Result, error) {
	if rc := C.b200sk_check_params(p); rc != 0 { // what the reference constructor returns before looking at a sequence
		return nil, codeToError(rc)
	}
	p.alphabet = b.alpha
	p.want_pos = 1
	p.pos_width = 4 // a production shim picks 1 or 2 when b.maxLen allows and widens in Index()
	p.max_read_len = C.uint32_t(b.maxLen)
	n := len(b.off) - 1
	var v *C.uint64_t
	var ps *C.uint32_t
	var o *C.uint64_t
	var st *C.int32_t
	var total C.uint64_t
	rc := C.b200sk_run(b.ctx.h, p, (*C.uint8_t)(b.bases), &b.off[0], C.uint64_t(n), &v, &ps, &o, &st, &total)
	if rc != 0 {
		return nil, codeToError(rc)
	}
	return &Result{
		val:    unsafe.Slice((*uint64)(unsafe.Pointer(v)), int(total)),
		pos:    unsafe.Slice((*uint32)(unsafe.Pointer(ps)), int(total)),
		off:    unsafe.Slice((*uint64)(unsafe.Pointer(o)), n+1),
		status: unsafe.Slice((*int32)(unsafe.Pointer(st)), n),
	}, nil
}

// KmerIterator == sketches.NewKmerIterator over every read (iterator.go:668).
func (b *Batch) KmerIterator(k int, canonical, circular bool) (*Result, error) {
	p := C.b200sk_params{mode: C.B200SK_MODE_KMER, k: C.int32_t(k), canonical: cbool(canonical), circular: cbool(circular)}
	return b.run(&p)
}

// HashIterator == sketches.NewHashIterator (iterator.go:615).
func (b *Batch) HashIterator(k int, canonical, circular bool) (*Result, error) {
	p := C.b200sk_params{mode: C.B200SK_MODE_NTHASH, k: C.int32_t(k), canonical: cbool(canonical), circular: cbool(circular)}
	return b.run(&p)
}

// SimHashIterator == sketches.NewSimHashIterator (iterator.go:113).
func (b *Batch) SimHashIterator(k, m, scale int, canonical, circular bool) (*Result, error) {
	p := C.b200sk_params{mode: C.B200SK_MODE_SIMHASH, k: C.int32_t(k), m: C.int32_t(m), scale: C.int32_t(scale),
		canonical: cbool(canonical), circular: cbool(circular)}
	return b.run(&p)
}

// MinimizerSketch == sketches.NewMinimizerSketch (sketch.go:85).
func (b *Batch) MinimizerSketch(k, w int, circular bool) (*Result, error) {
	if w > (1<<31)-1 {
		return nil, ErrInvalidW
	}
	p := C.b200sk_params{mode: C.B200SK_MODE_MINIMIZER, k: C.int32_t(k), w: C.int32_t(w), circular: cbool(circular)}
	return b.run(&p)
}

// SyncmerSketch == sketches.NewSyncmerSketch (sketch.go:142).
func (b *Batch) SyncmerSketch(k, s int, circular bool) (*Result, error) {
	p := C.b200sk_params{mode: C.B200SK_MODE_SYNCMER, k: C.int32_t(k), s: C.int32_t(s), circular: cbool(circular)}
	return b.run(&p)
}

// ProteinIterator == sketches.NewProteinIterator (iterator-protein.go:46).
func (b *Batch) ProteinIterator(k, codonTable, frame int) (*Result, error) {
	p := C.b200sk_params{mode: C.B200SK_MODE_PROTEIN, k: C.int32_t(k), codon_table: C.int32_t(codonTable), frame: C.int32_t(frame)}
	return b.run(&p)
}

// ProteinMinimizerSketch == sketches.NewProteinMinimizerSketch (sketch-protein.go:62).
func (b *Batch) ProteinMinimizerSketch(k, codonTable, frame, w int) (*Result, error) {
	if w > (1<<31)-1 {
		return nil, ErrInvalidW
	}
	p := C.b200sk_params{mode: C.B200SK_MODE_PROTEIN_MINIMIZER, k: C.int32_t(k), w: C.int32_t(w),
		codon_table: C.int32_t(codonTable), frame: C.int32_t(frame)}
	return b.run(&p)
}

func cbool(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
}

// Iterator replays one read's slice with the reference's method set
// (sketches.Iterator / sketches.Sketch / sketches.ProteinIterator: Next, Index).
type Iterator struct {
	val []uint64
	pos []uint32
	i   int
	err error // the constructor / NextKmer error of this read
}

// Sketch is the same replay type under the reference's other name.
type Sketch = Iterator

// Iterator returns read i's iterator, or the error the reference constructor returns for that read
// (ErrShortSeq ...).  For k-mer codes an illegal base is reported by Next after the codes before it
// (iterator.go:730-748).
func (r *Result) Iterator(i int) (*Iterator, error) {
	st := C.int(r.status[i])
	it := &Iterator{val: r.val[r.off[i]:r.off[i+1]], pos: r.pos[r.off[i]:r.off[i+1]]}
	if st == C.B200SK_ERR_ILLEGAL_BASE {
		it.err = ErrIllegalBase
		return it, nil
	}
	if st != 0 {
		return nil, codeToError(st)
	}
	return it, nil
}

// Sketch is Iterator under the name used for minimizers and syncmers.
func (r *Result) Sketch(i int) (*Sketch, error) { return r.Iterator(i) }

// Next returns the next element (sketches.Iterator.NextHash / Sketch.Next / ProteinIterator.Next).
func (it *Iterator) Next() (uint64, bool) {
	if it.i >= len(it.val) {
		return 0, false
	}
	v := it.val[it.i]
	it.i++
	return v, true
}

// NextKmer mirrors (*Iterator).NextKmer (iterator.go:708): the deferred illegal-base error comes last.
func (it *Iterator) NextKmer() (uint64, bool, error) {
	if it.i >= len(it.val) {
		return 0, false, it.err
	}
	v := it.val[it.i]
	it.i++
	return v, true, nil
}

// Index returns the 0-based position of the last returned element (iterator.go:776, sketch.go:488).
func (it *Iterator) Index() int { return int(it.pos[it.i-1]) }

// ---- record feeder: seqio/fastx.Reader.Read over a chunk of text (seqio/fastx/reader.go:233-471) -----------

// ErrNotFASTXFormat / ErrBadFASTQFormat are the reference's errors (seqio/fastx/reader.go:16,19).
var (
	ErrNotFASTXFormat = errors.New("fastx: invalid FASTA/Q format")
	ErrBadFASTQFormat = errors.New("fastx: bad fastq format")
)

// TextChunk is one chunk of FASTA/FASTQ text in C-owned pinned memory (fill Bytes()[:n] from the file or from
// xopen's decompressor, prefixed with the Carry of the previous chunk).
type TextChunk struct {
	ctx  *Context
	buf  unsafe.Pointer
	cap_ int
}

// NewTextChunk allocates a pinned chunk buffer.
func NewTextChunk(ctx *Context, capBytes int) *TextChunk {
	return &TextChunk{ctx: ctx, buf: C.b200sk_alloc_pinned(C.size_t(capBytes)), cap_: capBytes}
}

// Bytes is the chunk buffer.
func (t *TextChunk) Bytes() []byte { return unsafe.Slice((*byte)(t.buf), t.cap_) }

// Free releases the buffer.
func (t *TextChunk) Free() { C.b200sk_free_pinned(t.buf); t.buf = nil }

// FastxResult is a sketched chunk: the Result of its records plus what locates every record in the text.
type FastxResult struct {
	Result
	Format   int    // 1 FASTA, 2 FASTQ (pass it to the next chunk of the same file)
	Records  int    // complete records in this chunk
	Consumed int    // text[Consumed:n] is the cut record that opens the next chunk
	IsFastq  bool
}

// MinimizerSketchText replaces
//     for { rec, err := reader.Read(); sk, _ := sketches.NewMinimizerSketch(rec.Seq, k, w, false); for sk.Next() ... }
// over text[:n]: ONE cgo call copies the chunk to the GPU, splits it into records there, sketches them and
// brings the sketches back; the bases never exist as Go slices.
func (t *TextChunk) MinimizerSketchText(n, format int, final bool, k, w int) (*FastxResult, error) {
	p := C.b200sk_params{mode: C.B200SK_MODE_MINIMIZER, k: C.int32_t(k), w: C.int32_t(w), want_pos: 1, pos_width: 4}
	var info C.b200sk_fastx_info
	var v *C.uint64_t
	var ps *C.uint32_t
	var o *C.uint64_t
	var st *C.int32_t
	var total C.uint64_t
	rc := C.b200sk_run_fastx(t.ctx.h, &p, (*C.uint8_t)(t.buf), C.uint64_t(n), C.int(format), C.int(cbool(final)),
		&info, &v, &ps, &o, &st, &total)
	switch rc {
	case 0:
	case C.B200SK_ERR_NOT_FASTX:
		return nil, ErrNotFASTXFormat
	case C.B200SK_ERR_BAD_FASTQ:
		return nil, ErrBadFASTQFormat
	default:
		return nil, codeToError(rc)
	}
	nrec := int(info.n_records)
	return &FastxResult{
		Result: Result{
			val:    unsafe.Slice((*uint64)(unsafe.Pointer(v)), int(total)),
			pos:    unsafe.Slice((*uint32)(unsafe.Pointer(ps)), int(total)),
			off:    unsafe.Slice((*uint64)(unsafe.Pointer(o)), nrec+1),
			status: unsafe.Slice((*int32)(unsafe.Pointer(st)), nrec),
		},
		Format: int(info.format), Records: nrec, Consumed: int(info.consumed), IsFastq: info.format == C.B200SK_FASTX_FASTQ,
	}, nil
}

// FastxStream is the pipelined reader over a whole FASTA/FASTQ text in C-owned pinned memory (a slurped or
// decompressed file): what `for chunk := range reader.ChunkChan(bufferSize, chunkSize)` is to Read()
// (seqio/fastx/reader.go:556-603).  The library runs two slots with a worker thread each, so the copy and parse of
// chunk j+1 overlap the sketching and copy back of chunk j and the Go code consuming chunk j-1.
type FastxStream struct{ h *C.b200sk_fxstream }

// NewMinimizerStream opens a stream of minimizer sketches over text.Bytes()[:n]; chunkBytes = 0 means 256 MiB.
func NewMinimizerStream(device int, text *TextChunk, n, format, k, w int, chunkBytes uint64) (*FastxStream, error) {
	p := C.b200sk_params{mode: C.B200SK_MODE_MINIMIZER, k: C.int32_t(k), w: C.int32_t(w), want_pos: 1, pos_width: 4}
	var h *C.b200sk_fxstream
	if rc := C.b200sk_fxstream_open(&h, C.int(device), &p, (*C.uint8_t)(text.buf), C.uint64_t(n), C.int(format),
		C.uint64_t(chunkBytes)); rc != 0 {
		return nil, codeToError(rc)
	}
	return &FastxStream{h: h}, nil
}

// Next returns the next chunk of records in file order, io.EOF after the last one.  The slices of a chunk are
// valid until the next call.
func (s *FastxStream) Next() (*FastxResult, error) {
	var info C.b200sk_fastx_info
	var v *C.uint64_t
	var ps *C.uint32_t
	var o *C.uint64_t
	var st *C.int32_t
	var total C.uint64_t
	rc := C.b200sk_fxstream_next(s.h, &info, &v, &ps, &o, &st, &total)
	switch rc {
	case 0:
	case C.B200SK_FXSTREAM_END:
		return nil, io.EOF
	case C.B200SK_ERR_NOT_FASTX:
		return nil, ErrNotFASTXFormat
	case C.B200SK_ERR_BAD_FASTQ:
		return nil, ErrBadFASTQFormat
	default:
		return nil, codeToError(rc)
	}
	nrec := int(info.n_records)
	return &FastxResult{
		Result: Result{
			val:    unsafe.Slice((*uint64)(unsafe.Pointer(v)), int(total)),
			pos:    unsafe.Slice((*uint32)(unsafe.Pointer(ps)), int(total)),
			off:    unsafe.Slice((*uint64)(unsafe.Pointer(o)), nrec+1),
			status: unsafe.Slice((*int32)(unsafe.Pointer(st)), nrec),
		},
		Format: int(info.format), Records: nrec, Consumed: int(info.consumed), IsFastq: info.format == C.B200SK_FASTX_FASTQ,
	}, nil
}

// Rewind points the stream at another text (the device and pinned buffers are kept).
func (s *FastxStream) Rewind(text *TextChunk, n, format int) error {
	if rc := C.b200sk_fxstream_rewind(s.h, (*C.uint8_t)(text.buf), C.uint64_t(n), C.int(format)); rc != 0 {
		return codeToError(rc)
	}
	return nil
}

// Close stops the workers and frees both slots.
func (s *FastxStream) Close() { C.b200sk_fxstream_close(s.h); s.h = nil }
