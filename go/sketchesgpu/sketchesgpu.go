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
//	    sk, err := res.Sketch(i)              // *sketchesgpu.Sketch with the reference's method set, or the
//	    if err != nil { continue }            // error NewMinimizerSketch returns for that record (ErrShortSeq)
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
type Context struct {
	h   *C.b200sk_ctx
	gen uint64 // runs so far: a Result is valid while its gen equals this
}

// NewContext binds CUDA device `device`.  There is no CPU fallback: without a device this fails.
func NewContext(device int) (*Context, error) {
	var h *C.b200sk_ctx
	if rc := C.b200sk_create(&h, C.int(device)); rc != 0 {
		return nil, codeToError(rc)
	}
	return &Context{h: h}, nil
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

// ErrMixedAlphabet: the records of one batch share one b200sk_params, so they must share one alphabet.
var ErrMixedAlphabet = errors.New("sketchesgpu: records with different alphabets in one batch")

func alphabetCode(a *seq.Alphabet) C.int32_t {
	switch a {
	case seq.DNA:
		return C.B200SK_ALPHABET_DNA
	case seq.RNA:
		return C.B200SK_ALPHABET_RNA
	case seq.RNAredundant:
		return C.B200SK_ALPHABET_RNA_REDUNDANT
	case seq.Unlimit:
		return C.B200SK_ALPHABET_UNLIMIT
	case seq.Protein:
		return C.B200SK_ALPHABET_PROTEIN
	}
	return C.B200SK_ALPHABET_DNA_REDUNDANT
}

// Add appends one sequence (the bytes are copied: fastx.Reader reuses its buffer, reader.go:229-232).
func (b *Batch) Add(s *seq.Seq) error {
	a := alphabetCode(s.Alphabet)
	if len(b.off) == 1 {
		b.alpha = a
	} else if a != b.alpha {
		return ErrMixedAlphabet
	}
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
	return nil
}

// Reset empties the batch, keeping the buffer.
func (b *Batch) Reset() { b.nBases, b.off, b.maxLen = 0, b.off[:1], 0 }

// Free releases the pinned buffer.
func (b *Batch) Free() { C.b200sk_free_pinned(b.bases); b.bases = nil }

// Result holds the library-owned output arrays of one run.  They alias memory the NEXT run on the same context
// overwrites: every Result remembers the context's run counter and its accessors panic once it is stale.
type Result struct {
	val    []uint64
	pos    []uint32
	off    []uint64
	status []int32
	mode   C.int32_t
	ctx    *Context
	gen    uint64
}

func (r *Result) check() {
	if r.ctx != nil && r.ctx.gen != r.gen {
		panic("sketchesgpu: Result used after a later run on the same Context (its arrays were overwritten)")
	}
}

func (b *Batch) run(p *C.b200sk_params) (*Result, error) {
	// everything the checks look at is set first (a Protein batch, for one, skips the codon-table / frame checks:
	// iterator-protein.go:68)
	p.alphabet = b.alpha
	p.want_pos = 1
	p.pos_width = 4 // a production shim picks 1 or 2 when b.maxLen allows and widens in Index()
	p.max_read_len = C.uint32_t(b.maxLen)
	if rc := C.b200sk_check_params(p); rc != 0 { // what the reference constructor returns before looking at a sequence
		return nil, codeToError(rc)
	}
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
	b.ctx.gen++
	return &Result{
		val:    unsafe.Slice((*uint64)(unsafe.Pointer(v)), int(total)),
		pos:    unsafe.Slice((*uint32)(unsafe.Pointer(ps)), int(total)),
		off:    unsafe.Slice((*uint64)(unsafe.Pointer(o)), n+1),
		status: unsafe.Slice((*int32)(unsafe.Pointer(st)), n),
		mode:   p.mode, ctx: b.ctx, gen: b.ctx.gen,
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

// ProteinFrames == six sketches.NewProteinIterator loops, frame 1, 2, 3, -1, -2, -3 in this order, over every read
// (iterator-protein.go:46-90; BASELINE.json config 5) through ONE library call: the batch crosses PCIe once and, for
// reads of at most 384 bases and k <= 16, is fetched and decoded once on the device (b200sk_run_frames).
// Results[i].ProteinIterator(r) replays frame i of read r.  The six Results share the context's output arrays and
// become stale together at the next run.
func (b *Batch) ProteinFrames(k, codonTable int) ([6]*Result, error) {
	var out [6]*Result
	p := C.b200sk_params{mode: C.B200SK_MODE_PROTEIN, k: C.int32_t(k), codon_table: C.int32_t(codonTable), frame: 1}
	p.alphabet = b.alpha
	p.max_read_len = C.uint32_t(b.maxLen)
	if rc := C.b200sk_check_params(&p); rc != 0 {
		return out, codeToError(rc)
	}
	n := len(b.off) - 1
	var v, o [6]*C.uint64_t
	var st [6]*C.int32_t
	var total [6]C.uint64_t
	rc := C.b200sk_run_frames(b.ctx.h, &p, (*C.uint8_t)(b.bases), &b.off[0], C.uint64_t(n), &v[0], &o[0], &st[0], &total[0])
	if rc != 0 {
		return out, codeToError(rc)
	}
	b.ctx.gen++
	for i := range out {
		out[i] = &Result{
			val:    unsafe.Slice((*uint64)(unsafe.Pointer(v[i])), int(total[i])),
			off:    unsafe.Slice((*uint64)(unsafe.Pointer(o[i])), n+1),
			status: unsafe.Slice((*int32)(unsafe.Pointer(st[i])), n),
			mode:   p.mode, ctx: b.ctx, gen: b.ctx.gen, // pos == nil: Index() is the running position (replay.index)
		}
	}
	return out, nil
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

// ---- replay types: the reference's method sets over one read's slice of a Result -------------------------------
//
//	sketches.Iterator               NextKmer :708, NextHash :658, NextSimHash :191, Next :762 (3 values), Index :776
//	sketches.Sketch                 NextMinimizer sketch.go:205, NextSyncmer :312, Next :480 (2 values), Index :488
//	sketches.ProteinIterator        Next iterator-protein.go:76, Index :93
//	sketches.ProteinMinimizerSketch Next sketch-protein.go:106, Index :213
//	sketches.IdxValue               sketch.go:496

// IdxValue is sketches.IdxValue (sketch.go:496-499).
type IdxValue struct {
	Idx int    // index
	Val uint64 // hash
}

type replay struct {
	val []uint64
	pos []uint32
	i   int
	res *Result
}

func (it *replay) next() (uint64, bool) {
	it.res.check()
	if it.i >= len(it.val) {
		return 0, false
	}
	v := it.val[it.i]
	it.i++
	return v, true
}

// index: what Index() returns after the last Next().  Results without a position array (ProteinFrames: a dense mode,
// values only) count it: the i-th element of a dense iterator sits at position i.
func (it *replay) index() int {
	if it.pos == nil {
		return it.i - 1
	}
	return int(it.pos[it.i-1])
}

// IdxValues returns the read's whole slice as the reference's IdxValue pairs (what its users collect from a Next loop).
func (it *replay) IdxValues() []IdxValue {
	it.res.check()
	out := make([]IdxValue, len(it.val))
	for j := range it.val {
		idx := j
		if it.pos != nil {
			idx = int(it.pos[j])
		}
		out[j] = IdxValue{Idx: idx, Val: it.val[j]}
	}
	return out
}

// Iterator mirrors sketches.Iterator (k-mer codes, ntHash values, SimHash codes).
type Iterator struct {
	replay
	mode C.int32_t
	err  error // NextKmer's deferred illegal-base error of this read (iterator.go:730-748)
}

// Sketch mirrors sketches.Sketch (minimizers, syncmers).
type Sketch struct {
	replay
	syncmer bool
}

// ProteinIterator mirrors sketches.ProteinIterator.
type ProteinIterator struct{ replay }

// ProteinMinimizerSketch mirrors sketches.ProteinMinimizerSketch.
type ProteinMinimizerSketch struct{ replay }

func (r *Result) slice(i int) (replay, error) {
	r.check()
	st := C.int(r.status[i])
	rp := replay{val: r.val[r.off[i]:r.off[i+1]], res: r}
	if r.pos != nil {
		rp.pos = r.pos[r.off[i]:r.off[i+1]]
	}
	if st != 0 && st != C.B200SK_ERR_ILLEGAL_BASE {
		return rp, codeToError(st) // what the reference constructor returns for this read (ErrShortSeq ...)
	}
	return rp, nil
}

// Iterator returns read i's iterator (batches made by KmerIterator / HashIterator / SimHashIterator), or the error
// the reference constructor returns for that read.  For k-mer codes an illegal base is reported by NextKmer / Next
// after the codes before it, as in the reference.
func (r *Result) Iterator(i int) (*Iterator, error) {
	rp, err := r.slice(i)
	if err != nil {
		return nil, err
	}
	it := &Iterator{replay: rp, mode: r.mode}
	if C.int(r.status[i]) == C.B200SK_ERR_ILLEGAL_BASE {
		it.err = ErrIllegalBase
	}
	return it, nil
}

// Sketch returns read i's sketch (batches made by MinimizerSketch / SyncmerSketch).
func (r *Result) Sketch(i int) (*Sketch, error) {
	rp, err := r.slice(i)
	if err != nil {
		return nil, err
	}
	return &Sketch{replay: rp, syncmer: r.mode == C.B200SK_MODE_SYNCMER}, nil
}

// ProteinIterator returns read i's iterator (batches made by ProteinIterator).
func (r *Result) ProteinIterator(i int) (*ProteinIterator, error) {
	rp, err := r.slice(i)
	if err != nil {
		return nil, err
	}
	return &ProteinIterator{rp}, nil
}

// ProteinMinimizerSketch returns read i's sketch (batches made by ProteinMinimizerSketch).
func (r *Result) ProteinMinimizerSketch(i int) (*ProteinMinimizerSketch, error) {
	rp, err := r.slice(i)
	if err != nil {
		return nil, err
	}
	return &ProteinMinimizerSketch{rp}, nil
}

// NextKmer mirrors (*Iterator).NextKmer (iterator.go:708): the deferred illegal-base error comes after the codes
// before the k-mer that holds the illegal base.
func (it *Iterator) NextKmer() (code uint64, ok bool, err error) {
	code, ok = it.next()
	if !ok {
		return 0, false, it.err
	}
	return code, true, nil
}

// NextHash mirrors (*Iterator).NextHash (iterator.go:658).
func (it *Iterator) NextHash() (code uint64, ok bool) { return it.next() }

// NextSimHash mirrors (*Iterator).NextSimHash (iterator.go:191).
func (it *Iterator) NextSimHash() (code uint64, ok bool) { return it.next() }

// Next mirrors (*Iterator).Next (iterator.go:762-773): three values; only k-mer iterators ever return an error.
func (it *Iterator) Next() (code uint64, ok bool, err error) {
	if it.mode == C.B200SK_MODE_KMER {
		return it.NextKmer()
	}
	code, ok = it.next()
	return code, ok, nil
}

// Index mirrors (*Iterator).Index (iterator.go:776): 0-based position of the last element returned.
func (it *Iterator) Index() int { return it.index() }

// NextMinimizer mirrors (*Sketch).NextMinimizer (sketch.go:205).
func (s *Sketch) NextMinimizer() (code uint64, ok bool) { return s.next() }

// NextSyncmer mirrors (*Sketch).NextSyncmer (sketch.go:312).
func (s *Sketch) NextSyncmer() (code uint64, ok bool) { return s.next() }

// Next mirrors (*Sketch).Next (sketch.go:480): two values.
func (s *Sketch) Next() (uint64, bool) { return s.next() }

// Index mirrors (*Sketch).Index (sketch.go:488).
func (s *Sketch) Index() int { return s.index() }

// Next mirrors (*ProteinIterator).Next (iterator-protein.go:76).
func (it *ProteinIterator) Next() (code uint64, ok bool) { return it.next() }

// Index mirrors (*ProteinIterator).Index (iterator-protein.go:93).
func (it *ProteinIterator) Index() int { return it.index() }

// Next mirrors (*ProteinMinimizerSketch).Next (sketch-protein.go:106).
func (s *ProteinMinimizerSketch) Next() (code uint64, ok bool) { return s.next() }

// Index mirrors (*ProteinMinimizerSketch).Index (sketch-protein.go:213).
func (s *ProteinMinimizerSketch) Index() int { return s.index() }

// ---- multi-GPU and the reduced sketch ---------------------------------------------------------------------------

// Group drives several devices from one process (b200sk_group_*): one library context and worker thread per device,
// the batch sharded over them by cumulative bases, results in read order.
type Group struct {
	h   *C.b200sk_group
	ctx Context // carries the run counter the Results check
}

// NewGroup binds the given CUDA devices (1, 2, 4 or 8 of one box).
func NewGroup(devices []int) (*Group, error) {
	d := make([]C.int, len(devices))
	for i, v := range devices {
		d[i] = C.int(v)
	}
	var h *C.b200sk_group
	if rc := C.b200sk_group_create(&h, &d[0], C.int(len(d))); rc != 0 {
		return nil, codeToError(rc)
	}
	return &Group{h: h}, nil
}

// Close releases every device of the group.
func (g *Group) Close() { C.b200sk_group_destroy(g.h); g.h = nil }

// MinimizerSketch is Batch.MinimizerSketch over all devices of the group.
func (g *Group) MinimizerSketch(b *Batch, k, w int, circular bool) (*Result, error) {
	p := C.b200sk_params{mode: C.B200SK_MODE_MINIMIZER, k: C.int32_t(k), w: C.int32_t(w), circular: cbool(circular),
		alphabet: b.alpha, want_pos: 1, pos_width: 4, max_read_len: C.uint32_t(b.maxLen)}
	if rc := C.b200sk_check_params(&p); rc != 0 {
		return nil, codeToError(rc)
	}
	n := len(b.off) - 1
	var v *C.uint64_t
	var ps *C.uint32_t
	var o *C.uint64_t
	var st *C.int32_t
	var total C.uint64_t
	if rc := C.b200sk_group_run(g.h, &p, (*C.uint8_t)(b.bases), &b.off[0], C.uint64_t(n), &v, &ps, &o, &st, &total); rc != 0 {
		return nil, codeToError(rc)
	}
	g.ctx.gen++
	return &Result{
		val:    unsafe.Slice((*uint64)(unsafe.Pointer(v)), int(total)),
		pos:    unsafe.Slice((*uint32)(unsafe.Pointer(ps)), int(total)),
		off:    unsafe.Slice((*uint64)(unsafe.Pointer(o)), n+1),
		status: unsafe.Slice((*int32)(unsafe.Pointer(st)), n),
		mode:   p.mode, ctx: &g.ctx, gen: g.ctx.gen,
	}, nil
}

// MinimizerSet replaces the consumer loop of a FracMinHash / unique-k-mer tool,
//
//	for each record { for sk.Next() { if h <= math.MaxUint64/scale { set[h] = struct{}{} } } }; sort(keys(set))
//
// (the scale rule of sketches/iterator.go:180-185): the per-read arrays never leave the GPU, only the sorted distinct
// values come back (b200sk_run_reduced).  scale <= 1 keeps every minimizer.
func (b *Batch) MinimizerSet(k, w int, scale uint32) ([]uint64, error) {
	p := C.b200sk_params{mode: C.B200SK_MODE_MINIMIZER, k: C.int32_t(k), w: C.int32_t(w), alphabet: b.alpha,
		max_read_len: C.uint32_t(b.maxLen)}
	if rc := C.b200sk_check_params(&p); rc != 0 {
		return nil, codeToError(rc)
	}
	var v *C.uint64_t
	var total C.uint64_t
	if rc := C.b200sk_run_reduced(b.ctx.h, &p, C.uint32_t(scale), 1, (*C.uint8_t)(b.bases), &b.off[0],
		C.uint64_t(len(b.off)-1), &v, &total); rc != 0 {
		return nil, codeToError(rc)
	}
	b.ctx.gen++
	return unsafe.Slice((*uint64)(unsafe.Pointer(v)), int(total)), nil
}

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
			mode:   C.B200SK_MODE_MINIMIZER, // (valid until the next call on this chunk / stream)
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
			mode:   C.B200SK_MODE_MINIMIZER, // (valid until the next call on this chunk / stream)
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
