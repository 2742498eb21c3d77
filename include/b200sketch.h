/*
 * b200sketch.h -- C ABI of libb200sketch.so: the batch-shaped drop-in boundary
 * for the sketching hot path of shenwei356/bio (reference @ 7b48836e).
 *
 * The reference has no FFI; its boundary for this path is the exported Go
 * method set of package `sketches` (pull iterators, one uint64 per Next()).
 * A per-element cgo call (~100 ns) would dwarf the 8 ns/base the Go loop
 * needs, so the ABI is batch-shaped: the host side packs the records of a
 * fastx chunk into one pinned byte buffer + offsets, one call sketches the
 * whole batch on the GPU, and the Go shim (go/sketchesgpu)
 * replays the per-read slices through Next()/Index().
 *
 * Each entry point cites the reference interface it replaces (paths relative
 * to the reference root).  Plain pointers and sizes only -- no torch types.
 *
 * Output contract (all modes): for read r the emitted elements are
 * out_val[out_off[r] .. out_off[r+1]) in exactly the order the reference's
 * Next() loop yields them; out_pos holds what Index() returns after each
 * Next().  read_status[r] carries the error the reference constructor (or
 * NextKmer) would have returned for that read (0 = ok); a read whose
 * constructor fails emits nothing.
 */
#ifndef B200SKETCH_H
#define B200SKETCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error / status codes ------------------------------------------------ */
#define B200SK_OK 0
#define B200SK_ERR_INVALID_K (-1)       /* sketches/iterator.go:35  ErrInvalidK    */
#define B200SK_ERR_SHORT_SEQ (-2)       /* sketches/iterator.go:41  ErrShortSeq    */
#define B200SK_ERR_INVALID_W (-3)       /* sketches/sketch.go:36    ErrInvalidW    */
#define B200SK_ERR_INVALID_S (-4)       /* sketches/sketch.go:33    ErrInvalidS    */
#define B200SK_ERR_ILLEGAL_BASE (-5)    /* sketches/iterator.go:44  ErrIllegalBase */
#define B200SK_ERR_K_OVERFLOW (-6)      /* kmers.ErrKOverflow (k > 32), iterator.go:742 */
#define B200SK_ERR_INVALID_FRAME (-7)   /* seq/codon_tables.go:209-211             */
#define B200SK_ERR_CODON_TABLE (-8)     /* seq/seq.go:685 Translate: unknown table */
#define B200SK_ERR_TRANSLATE_SHORT (-9) /* seq/codon_tables.go:206-208             */
#define B200SK_ERR_INVALID_CODON (-10)  /* seq.ErrInvalidDNABase (not reachable: allowUnknownCodon=true) */
#define B200SK_ERR_INVALID_M (-11)      /* sketches/iterator.go:50  ErrInvalidM     */
#define B200SK_ERR_INVALID_SCALE (-12)  /* sketches/iterator.go:53  ErrInvalidScale */
#define B200SK_ERR_K_TOO_LARGE (-13)    /* sketches/iterator.go:47  ErrKTooLarge (k >= 65535) */
#define B200SK_ERR_NOT_FASTX (-20)      /* seqio/fastx/reader.go:37  ErrNotFASTXFormat                      */
#define B200SK_ERR_BAD_FASTQ (-21)      /* seqio/fastx/reader.go:40,43 ErrBadFASTQFormat / ErrUnequalSeqAndQual */
/* library-level conditions (no reference counterpart) */
#define B200SK_ERR_CUDA (-100)          /* a CUDA runtime call failed; see b200sk_last_error */
#define B200SK_ERR_NO_DEVICE (-101)     /* no CUDA device: there is NO CPU fallback */
#define B200SK_ERR_UNSUPPORTED (-102)   /* parameter outside what the kernels implement */
#define B200SK_ERR_CAPACITY (-103)      /* caller-provided device output too small; *n_out = needed */
#define B200SK_ERR_BAD_ARG (-104)
#define B200SK_ERR_NOMEM (-105)

/* ---- modes: which reference constructor the batch stands for ------------- */
#define B200SK_MODE_KMER 0      /* NewKmerIterator      sketches/iterator.go:668 + NextKmer :708   */
#define B200SK_MODE_NTHASH 1    /* NewHashIterator      sketches/iterator.go:615 + NextHash :658   */
#define B200SK_MODE_MINIMIZER 2 /* NewMinimizerSketch   sketches/sketch.go:85    + NextMinimizer :205 */
#define B200SK_MODE_SYNCMER 3   /* NewSyncmerSketch     sketches/sketch.go:142   + NextSyncmer :312   */
#define B200SK_MODE_PROTEIN 4   /* NewProteinIterator   sketches/iterator-protein.go:46 + Next :76   */
#define B200SK_MODE_SIMHASH 6   /* NewSimHashIterator   sketches/iterator.go:113 + NextSimHash :191 (k, m, scale, canonical, circular) */
#define B200SK_MODE_PROTEIN_MINIMIZER 5 /* NewProteinMinimizerSketch sketches/sketch-protein.go:62 + Next :106 (k, w, codon_table, frame) */

/* seq.Alphabet of the records (only NextKmer's non-canonical second strand
 * depends on it: RevComInplace, sketches/iterator.go:719, seq/seq.go:350). */
#define B200SK_ALPHABET_DNA_REDUNDANT 0 /* seq/alphabet.go:361 */
#define B200SK_ALPHABET_DNA 1           /* seq/alphabet.go:353 */
#define B200SK_ALPHABET_RNA_REDUNDANT 2 /* seq/alphabet.go:377 */
#define B200SK_ALPHABET_RNA 3           /* seq/alphabet.go:369 */
#define B200SK_ALPHABET_UNLIMIT 4       /* seq/alphabet.go:399: complement is a no-op */
#define B200SK_ALPHABET_PROTEIN 5       /* MODE_PROTEIN only: records are amino acids already (iterator-protein.go:68) */

/* Constructor arguments, shared by every read of the batch. */
typedef struct b200sk_params {
    int32_t mode;        /* B200SK_MODE_*                                             */
    int32_t k;           /* k-mer size (amino acids for MODE_PROTEIN)                 */
    int32_t w;           /* minimizer window (MODE_MINIMIZER)                         */
    int32_t s;           /* s-mer size (MODE_SYNCMER)                                 */
    int32_t canonical;   /* KMER / NTHASH only; minimizer and syncmer are always canonical (sketch.go:212,319) */
    int32_t circular;    /* append seq[0:k-1] (iterator.go:642-646, sketch.go:106-110,163-167)  */
    int32_t codon_table; /* MODE_PROTEIN: NCBI transl_table id (seq/codon_tables.go:431-640)    */
    int32_t frame;       /* MODE_PROTEIN: 1,2,3,-1,-2,-3 (seq/codon_tables.go:205)              */
    int32_t alphabet;    /* B200SK_ALPHABET_*                                         */
    int32_t want_pos;    /* 0: do not produce out_pos (dense modes: pos is just the running index) */
    uint32_t max_read_len; /* optional hint: length of the longest read (0 = unknown; the library then
                              measures it on the device, which costs one extra pass over read_off) */
    int32_t m;           /* MODE_SIMHASH: m-mer size, range [4, k] (iterator.go:121)             */
    int32_t scale;       /* MODE_SIMHASH: FracMinHash scale of the m-mers, range [1, k-m+1] (:124) */
    int32_t pos_width;   /* bytes per out_pos element: 0 or 4 = uint32 (default), 1 = uint8, 2 = uint16.  Narrow
                            positions need the max_read_len hint (<= 256 resp. 65536): every Index() must fit.  The
                            out_pos pointers then address uint8_t / uint16_t arrays (a third less PCIe traffic for
                            150-bp reads). */
    int32_t reserved[2];
} b200sk_params;

/* One per caller thread and device: the HOST side of a context is not thread-safe.  On the device a context owns one
 * set of scratch words (ticket, look-back status words, item tables) that serve ONE batch at a time: the `stream`
 * argument of the device entry points may change from call to call -- a batch enqueued on another stream than the
 * batch before it is ordered behind it on the device (cudaStreamWaitEvent inside the call), never raced.  For batches
 * that should overlap on the device use one context per stream. */
typedef struct b200sk_ctx b200sk_ctx;

/* Library / device lifetime.  device = CUDA ordinal.  There is no CPU
 * fallback: without a usable device this returns B200SK_ERR_NO_DEVICE. */
int b200sk_create(b200sk_ctx **ctx, int device);
void b200sk_destroy(b200sk_ctx *ctx);

/* C-owned pinned staging: the Go side packs record.Seq.Seq bytes straight into
 * this buffer (cgo must not let C retain Go pointers across calls). */
void *b200sk_alloc_pinned(size_t bytes);
void b200sk_free_pinned(void *p);

/* The constructor checks every New*() performs before looking at a sequence
 * (iterator.go:616,669; sketch.go:86-91,143-148; iterator-protein.go:47;
 * codon_tables.go:209).  Returns 0 or the B200SK_ERR_* the constructor returns. */
int b200sk_check_params(const b200sk_params *p);

/* Upper bound on the number of elements a batch can emit (for sizing device
 * output in b200sk_run_device).  exact=0 returns the default (expected-density)
 * capacity the host path starts with. */
uint64_t b200sk_output_bound(const b200sk_params *p, uint64_t n_bases, uint64_t n_reads, int exact);

/* Host entry point: replaces the per-record loop
 *     it, err := sketches.NewXxx(record.Seq, ...); for { v, ok := it.Next(); i := it.Index() }
 * over the n_reads records of a batch.  bases = concatenated record.Seq.Seq
 * bytes (ASCII as delivered by seqio/fastx.Reader.Read, reader.go:233),
 * read_off[n_reads+1] = byte offsets.  Host<->device copies are pipelined
 * inside.  Outputs are library-owned pinned host arrays, valid until the next
 * b200sk_run / b200sk_destroy on this ctx. */
int b200sk_run(b200sk_ctx *ctx, const b200sk_params *p,
               const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads,
               uint64_t **out_val, uint32_t **out_pos, uint64_t **out_off,
               int32_t **read_status, uint64_t *n_out);

/* Device entry point: same contract with every buffer already resident in HBM
 * (what a multi-stage GPU pipeline or the multi-GPU shard driver calls).
 * d_bases must be readable up to the next 16-byte boundary past its end (TMA
 * bulk copies move 16-byte units).  d_out_pos may be NULL.  capacity = elements
 * available in d_out_val/d_out_pos.  Work is enqueued on `stream` (a
 * cudaStream_t passed as void*); *n_out is valid after the call returns (the
 * call synchronises the stream once to read it).  On B200SK_ERR_CAPACITY
 * nothing was written to d_out_val/pos and *n_out holds the required capacity. */
int b200sk_run_device(b200sk_ctx *ctx, const b200sk_params *p,
                      const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads,
                      uint64_t n_bases,
                      uint64_t *d_out_val, uint32_t *d_out_pos, uint64_t *d_out_off,
                      int32_t *d_read_status, uint64_t capacity, void *stream, uint64_t *n_out);

/* Same, but fully asynchronous: no host synchronisation; the element count is
 * left in d_out_off[n_reads] and a capacity overflow is reported through
 * *d_flags (bit 0) on the device.  Used inside timed regions. */
int b200sk_enqueue_device(b200sk_ctx *ctx, const b200sk_params *p,
                          const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads,
                          uint64_t n_bases,
                          uint64_t *d_out_val, uint32_t *d_out_pos, uint64_t *d_out_off,
                          int32_t *d_read_status, uint64_t capacity, void *stream,
                          uint32_t *d_flags);

/* ---- record feeder: seqio/fastx.Reader.Read (seqio/fastx/reader.go:233-471) as a batch operation ----------
 * The caller side of the path.  The reference reads one record per Read() call: finds the
 * next delimiter that follows a newline, strips line ends, joins the sequence lines (parseRecord,
 * reader.go:372-471).  Here a whole chunk of FASTA/FASTQ text is split into records on the device and its
 * sequences land packed in HBM in exactly the layout b200sk_run_device takes (bases + read_off), so a
 * chunk goes text -> records -> sketches without the bases ever visiting the host.
 * FASTA and FASTQ with any line structure (four-line FASTQ records take a parallel path, anything else the
 * reference's general record rule, reader.go:308-345,396-417); LF or CRLF line ends; leading blank lines; a last
 * line without newline.  The alphabet is guessed from the first record and every letter checked against it
 * (reader.go:430-452): b200sk_fastx_info.alphabet / first_invalid / d_invalid.  The reference's Read() returns the
 * offending record together with an error; here the chunk is parsed to its end and the caller decides. */
#define B200SK_FASTX_FASTA 1
#define B200SK_FASTX_FASTQ 2
typedef struct b200sk_fastx_info {
    int32_t format;        /* B200SK_FASTX_FASTA / _FASTQ (reader.go:271-304: first byte that is not a newline) */
    int32_t status;        /* B200SK_OK or the error also returned                                              */
    uint64_t n_records;    /* complete records found                                                            */
    uint64_t n_bases;      /* bytes in d_bases = d_read_off[n_records]                                          */
    uint64_t n_lines;      /* lines looked at                                                                   */
    uint64_t consumed;     /* text bytes the records cover: feed text[consumed:] + the next chunk next time     */
    uint64_t bad_record;   /* B200SK_ERR_BAD_FASTQ: index of the first record that is not header/seq/+/qual     */
    uint32_t max_read_len; /* FASTQ: longest sequence (the max_read_len hint of b200sk_params); FASTA: 0        */
    uint32_t reserved;
    /* library-owned device arrays, valid until the next feeder call on this ctx */
    uint8_t *d_bases;      /* record.Seq.Seq of every record, concatenated                                      */
    uint64_t *d_read_off;  /* [n_records+1] offsets into d_bases                                                */
    uint64_t *d_rec_off;   /* [n_records+1] text offset of each record's delimiter; [n_records] = consumed      */
    uint64_t *d_qual_off;  /* FASTQ: [n_records] text offset of the quality line; FASTA: NULL                   */
    uint64_t *d_line_off;  /* [n_lines+1] text offset of every line start                                       */
    int32_t alphabet;      /* B200SK_ALPHABET_*: seq.GuessAlphabetLessConservatively over the first record's first
                              10 000 letters (reader.go:430-435, seq/alphabet.go:411-452), or the one passed in      */
    int32_t reserved2;
    uint64_t first_invalid; /* first record holding a letter outside that alphabet (Alphabet.IsValid,
                              seq/alphabet.go:234-300: the error Read() returns with that record), ~0 = none        */
    uint8_t *d_invalid;    /* [n_records] 1 = the record holds such a letter                                     */
} b200sk_fastx_info;

/* d_text: the chunk in HBM, 16-byte aligned, readable up to the next 16-byte boundary past n_bytes.
 * format: 0 = detect, else B200SK_FASTX_* (a chunk that continues a file starts at a record and passes
 * the file's format; it also passes the file's alphabet, as (1 + info.alphabet of the first chunk) << 8, so that the
 * guess is made once per file as in the reference).  final: 1 = the text ends here (the last record is complete), 0 = more follows (the
 * last, possibly cut, record is left for the next call: see consumed). */
int b200sk_fastx_parse_device(b200sk_ctx *ctx, const uint8_t *d_text, uint64_t n_bytes, int format, int final,
                              void *stream, b200sk_fastx_info *info);

/* Host entry point for the whole front of the path: replaces
 *     for { record, err := reader.Read(); it, _ := sketches.NewXxx(record.Seq, ...); for it.Next() ... }
 * over one chunk of FASTA/FASTQ text in host memory (pinned for full copy speed): copy to the device,
 * split into records, sketch, copy the sketches back.  Outputs as b200sk_run. */
int b200sk_run_fastx(b200sk_ctx *ctx, const b200sk_params *p, const uint8_t *text, uint64_t n_bytes, int format,
                     int final, b200sk_fastx_info *info, uint64_t **out_val, uint32_t **out_pos, uint64_t **out_off,
                     int32_t **read_status, uint64_t *n_out);

/* Pipelined reader: the loop above over a WHOLE FASTA/FASTQ text in host memory (an mmap'd or slurped file;
 * pinned for full copy speed), handed back chunk by chunk in file order -- what
 *     for chunk := range reader.ChunkChan(bufferSize, chunkSize) { ... }     (seqio/fastx/reader.go:556-603)
 * is to Read().  Two slots, each with its own device buffers, pinned arrays, stream and worker thread: the copy
 * and parse of chunk j+1 overlap the sketching and copy back of chunk j and the caller consuming chunk j-1.
 * Chunks are cut every chunk_bytes of text (0 = 256 MiB) and end on a record boundary (b200sk_fastx_info
 * .consumed); a chunk that holds no complete record grows until it does.  format: 0 = detect.
 * b200sk_fxstream_next returns 0 and one chunk of records (info and arrays as b200sk_run_fastx; valid until the
 * next call on this stream), B200SK_FXSTREAM_END after the last chunk, or the B200SK_ERR_* of the chunk that
 * failed (chunks before it are returned first; none after it).  One caller thread per stream.
 * b200sk_fxstream_rewind points an open stream at another text (the buffers are kept). */
#define B200SK_FXSTREAM_END 1
typedef struct b200sk_fxstream b200sk_fxstream;
int b200sk_fxstream_open(b200sk_fxstream **s, int device, const b200sk_params *p, const uint8_t *text,
                         uint64_t n_bytes, int format, uint64_t chunk_bytes);
int b200sk_fxstream_next(b200sk_fxstream *s, b200sk_fastx_info *info, uint64_t **out_val, uint32_t **out_pos,
                         uint64_t **out_off, int32_t **read_status, uint64_t *n_out);
int b200sk_fxstream_rewind(b200sk_fxstream *s, const uint8_t *text, uint64_t n_bytes, int format);
void b200sk_fxstream_close(b200sk_fxstream *s);
uint64_t b200sk_fxstream_kernel_launches(const b200sk_fxstream *s);
const char *b200sk_fxstream_last_error(const b200sk_fxstream *s);

/* ---- multi-GPU -------------------------------------------------------------------------------
 * Records carry no cross-record state (sketches/iterator.go:615-655, sketches/sketch.go:85-202): a batch shards
 * over GPUs by contiguous read ranges and the only exchange is the gather of the per-GPU uint64 arrays; rank order
 * == read order, so the gathered array is exactly what one GPU would have produced.
 *
 * One process per GPU: the root allocates the gather buffer and exports it over CUDA IPC; every other rank maps
 * it and passes a pointer INTO it as d_out_val of b200sk_enqueue_device / b200sk_run_device -- the sketching
 * kernel's own flush then stores its elements straight into the root's HBM over NVLink, tile by tile while the
 * rest of the shard is still being walked (no staging copy, no collective on the data path).  Rank r's segment
 * starts at a base the caller picks from b200sk_output_bound of the shards before it; the element counts
 * (d_out_off[n_reads] of every rank, 8 bytes each) travel through whatever the host side uses.
 * b200sk_compact_segments then closes the gaps on the root, in place.  The handle is a cudaIpcMemHandle_t. */
#define B200SK_IPC_HANDLE_BYTES 64
int b200sk_gather_create(b200sk_ctx *ctx, uint64_t capacity_elems, uint8_t *handle /*[64] out*/, uint64_t **d_buf);
int b200sk_gather_open(b200sk_ctx *ctx, const uint8_t *handle /*[64]*/, uint64_t **d_buf);
int b200sk_gather_close(b200sk_ctx *ctx, uint64_t *d_buf, int is_owner);
/* seg_base / seg_count: host arrays, segment r = d_buf[seg_base[r] .. seg_base[r] + seg_count[r]), ordered,
 * disjoint, seg_base[0] == 0.  Afterwards d_buf[0 .. sum of counts) is the gathered array. */
int b200sk_compact_segments(b200sk_ctx *ctx, uint64_t *d_buf, const uint64_t *seg_base, const uint64_t *seg_count,
                            int n_seg, void *stream);
/* cut[n_shards + 1]: shard d owns reads [cut[d], cut[d+1]), balanced by cumulative bases (long, skewed reads). */
void b200sk_shard_by_bases(const uint64_t *read_off, uint64_t n_reads, int n_shards, uint64_t *cut);

/* The fused form: ONE ordered output chain across the GPUs.  The batch is dealt out in chunks of chunk_reads
 * consecutive reads (chunk c belongs to rank c % n_ranks; a rank holds its chunks back to back in its own HBM),
 * every rank runs its sketching kernel over its chunks, and the kernels' single-pass output allocation (the
 * decoupled look-back over per-tile status words) runs across all ranks at once: a tile publishes its count into
 * every rank's copy of the status words (n-1 posted 8-byte stores over NVLink) and polls only its own copy.  Each
 * flush therefore knows its exact place in the ROOT's arrays and stores there directly -- d_out_val, d_out_pos,
 * d_out_off ([n_reads_global + 1]) and d_read_status ([n_reads_global]) are the root's buffers (peer-mapped on the
 * other ranks), indexed by the global read / element.  When every rank's kernel has finished, the root holds
 * exactly what one GPU would have produced for the whole batch: no staging copy, no compaction, no counts to
 * exchange.  Minimizer / syncmer batches of reads of one item each (max_read_len given, <= 384), not circular.
 * state[r]: rank r's status-word array -- ceil(n_reads_global / 32) + 1 uint64, zeroed once when allocated, never
 * again -- as mapped in THIS process (b200sk_gather_create / _open work for any buffer).  epoch: the same on every
 * rank, not 0, and different mod 16384 from the step before; every rank must have finished step e before any rank
 * enqueues step e + 1 (a stream-ordered barrier of the host's choice). */
typedef struct b200sk_shard_spec {
    int32_t rank, n_ranks;   /* 1 <= n_ranks <= 8 */
    uint32_t chunk_reads;    /* a multiple of 32 */
    uint32_t epoch;
    uint64_t n_reads_global;
    uint64_t *state[8];
} b200sk_shard_spec;
int b200sk_enqueue_device_sharded(b200sk_ctx *ctx, const b200sk_params *p, const b200sk_shard_spec *spec,
                                  const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases,
                                  uint64_t *d_out_val, uint32_t *d_out_pos, uint64_t *d_out_off, int32_t *d_read_status,
                                  uint64_t capacity, void *stream, uint32_t *d_flags);

/* All six reading frames of a batch in one call: replaces the six loops
 *     for _, frame := range []int{1, 2, 3, -1, -2, -3} {
 *         it, err := sketches.NewProteinIterator(record.Seq, k, codonTable, frame); for { v, ok := it.Next() } }
 * (sketches/iterator-protein.go:46-90; the frames of seq/codon_tables.go:205-285) -- BASELINE.json config 5.
 * p->mode must be B200SK_MODE_PROTEIN; p->frame and p->want_pos are ignored (Index() of a dense mode is the running
 * position).  [i] of every array argument belongs to frame 1, 2, 3, -1, -2, -3 in this order: d_out_val[i] (capacity
 * elements each), d_out_off[i] (n_reads + 1), d_read_status[i] (n_reads; the array or single entries may be null).
 * With a max_read_len hint of at most 384, k <= 16 and nucleotide input the reads are fetched and decoded ONCE and
 * walked six times by one kernel; any other batch runs as six ordinary batches behind this entry point. */
int b200sk_enqueue_device_frames(b200sk_ctx *ctx, const b200sk_params *p, const uint8_t *d_bases,
                                 const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases,
                                 uint64_t *const *d_out_val, uint64_t *const *d_out_off, int32_t *const *d_read_status,
                                 uint64_t capacity, void *stream, uint32_t *d_flags);

/* The same from host memory (the contract of b200sk_run: record bytes + offsets in, library-owned pinned arrays out,
 * valid until the next b200sk_run* on this context): the batch crosses PCIe once for all six frames.
 * out_val[i] / out_off[i] / read_status[i] / n_out[i]: frame 1, 2, 3, -1, -2, -3 (read_status may be null). */
int b200sk_run_frames(b200sk_ctx *ctx, const b200sk_params *p, const uint8_t *bases, const uint64_t *read_off,
                      uint64_t n_reads, uint64_t **out_val, uint64_t **out_off, int32_t **read_status, uint64_t *n_out);

/* One process, several devices: what SURVEY.md 8b calls b200sk_create(ctx**, devices, n).  One context, stream and
 * worker thread per device; b200sk_group_run has the contract of b200sk_run (host pointers in, library-owned pinned
 * arrays in read order out, valid until the next call on the group) with the reads sharded over every device of
 * the group by cumulative bases.  One caller thread per group. */
typedef struct b200sk_group b200sk_group;
int b200sk_group_create(b200sk_group **g, const int *devices, int n_devices);
void b200sk_group_destroy(b200sk_group *g);
int b200sk_group_size(const b200sk_group *g);
int b200sk_group_run(b200sk_group *g, const b200sk_params *p, const uint8_t *bases, const uint64_t *read_off,
                     uint64_t n_reads, uint64_t **out_val, uint32_t **out_pos, uint64_t **out_off,
                     int32_t **read_status, uint64_t *n_out);
const char *b200sk_group_last_error(const b200sk_group *g);
uint64_t b200sk_group_kernel_launches(const b200sk_group *g);

/* ---- downstream reduction of the hash arrays -----------------------------------------------
 * What the tools built on package sketches (kmcp, unikmer: sketches/README.md:14-15,42) do with the uint64 stream:
 * keep the FracMinHash fraction h <= MaxUint64 / scale -- the rule sketches/iterator.go:180-185,281,443 applies to
 * its m-mer hashes --, sort ascending, drop duplicates.  Done on the device, so that only the reduced sketch crosses
 * PCIe / NVLink.  d_val[0..n) is consumed as scratch; d_out (capacity elements, must not alias d_val; capacity >= n
 * when scale <= 1) receives the result; *n_out = its length (or the length needed, with B200SK_ERR_CAPACITY).
 * scale <= 1: no filter.  unique = 0: sorted, duplicates kept. */
uint64_t b200sk_scale_max_hash(uint32_t scale); /* MaxUint64 / scale, iterator.go:184 */
int b200sk_reduce_device(b200sk_ctx *ctx, uint64_t *d_val, uint64_t n, uint32_t scale, int unique, uint64_t *d_out,
                         uint64_t capacity, uint64_t *n_out, void *stream);
/* Host entry point with the reduction inside: b200sk_run's input contract, but the per-read arrays never leave the
 * device -- every sub-batch's values are filtered into one accumulation array there, and sort | unique of that array
 * is all that comes back (library-owned pinned array, valid until the next call on this ctx).  Replaces
 *     for each record { for it.Next() { if h <= maxHash { set[h] = struct{}{} } } }; sort(keys(set))
 * on the host side of a FracMinHash / unique-k-mer consumer.  D2H traffic falls from 9-12 bytes per emitted element to
 * 8 bytes per KEPT DISTINCT element. */
int b200sk_run_reduced(b200sk_ctx *ctx, const b200sk_params *p, uint32_t scale, int unique, const uint8_t *bases,
                       const uint64_t *read_off, uint64_t n_reads, uint64_t **out_val, uint64_t *n_out);

/* Synchronous copy of a library-owned device array (the feeder's tables) into host memory. */
int b200sk_copy_to_host(b200sk_ctx *ctx, void *dst, const void *d_src, uint64_t bytes);

/* Error text: the reference's error strings for the codes that mirror them
 * (iterator.go:34-53, sketch.go:32-42), library text otherwise. */
const char *b200sk_strerror(int code);
const char *b200sk_last_error(const b200sk_ctx *ctx); /* CUDA error detail */

/* Introspection for the bench / tests. */
uint64_t b200sk_kernel_launches(const b200sk_ctx *ctx); /* kernels launched so far on this ctx */
/* Kernel timing: when enabled, every batch brackets its main sketching kernel with CUDA events on the
 * stream it is launched on.  b200sk_timing_collect waits for the pending pairs, adds their durations
 * (milliseconds) into *sum_ms, their number into *n, and forgets them. */
void b200sk_timing_enable(b200sk_ctx *ctx, int on);
int b200sk_timing_collect(b200sk_ctx *ctx, double *sum_ms, uint64_t *n);
int b200sk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SKETCH_H */
