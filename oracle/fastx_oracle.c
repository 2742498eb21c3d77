/*
 * fastx_oracle.c -- CPU restatement of seqio/fastx.Reader.Read + parseRecord
 * (reference: seqio/fastx/reader.go:233-471, @ 7b48836e) over one in-memory text.
 *
 * TEST INFRASTRUCTURE ONLY: the checker the record feeder (bio_b200/csrc/b200sk_fastx.cu)
 * is compared with.  Nothing under bio_b200/ links, imports or executes it.
 *
 * It follows the Go control flow literally -- the delimiter search with the "previous byte is a
 * newline" rule (reader.go:308-350), the retry when a FASTQ candidate leaves the quality shorter than
 * the sequence (:328-340), the last-part branch (:352-364) and parseRecord's line loops (:372-428),
 * dropCR / dropLF (:535-547) -- with the reader's 64 KiB refill loop collapsed (the text is one buffer;
 * nothing in the record logic depends on where a refill falls).  Not restated: alphabet guessing and
 * per-letter validation (:430-452), ID/description splitting (:486-525, a pure function of the header
 * line that the host shim applies lazily).
 *
 * Pinned by: the record counts of the reference's own fixtures (seqio/fastx/reader_test.go:84,105,125,
 * 130-158: test.fa = 6, test.fq = 8, test2.fq = 5, test3.fq = 3 records, the last with quality lines that
 * start with '@'), checked by tests/test_fastx.py when /root/reference is present, and by hand-written
 * vectors for each rule above.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FX_OK 0
#define FX_ERR_NOT_FASTX (-20)
#define FX_ERR_BAD_FASTQ (-21)
#define FX_FASTA 1
#define FX_FASTQ 2

typedef struct {
    uint8_t *p;
    uint64_t *src; /* text offset every buffered byte came from */
    size_t len, cap;
} fxbuf;

static void fx_push(fxbuf *b, uint8_t c, uint64_t src) {
    if (b->len == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 4096;
        b->p = (uint8_t *)realloc(b->p, b->cap);
        b->src = (uint64_t *)realloc(b->src, b->cap * sizeof(uint64_t));
    }
    b->p[b->len] = c;
    b->src[b->len] = src;
    b->len++;
}
static void fx_write(fxbuf *b, const uint8_t *t, uint64_t from, uint64_t to) {
    for (uint64_t i = from; i < to; i++) fx_push(b, t[i], i);
}
/* dropCR, reader.go:535-541: end of the slice [s, e) without one trailing '\r' */
static size_t drop_cr(const uint8_t *p, size_t s, size_t e) { return (e > s && p[e - 1] == '\r') ? e - 1 : e; }

typedef struct {
    uint8_t *bases;
    uint64_t n_bases, cap_bases;
    uint64_t *read_off, *rec_off, *qual_off, *name_len;
    uint64_t n_rec, cap_rec;
} fxout;

static void out_base(fxout *o, uint8_t c) {
    if (o->n_bases == o->cap_bases) {
        o->cap_bases = o->cap_bases ? o->cap_bases * 2 : 4096;
        o->bases = (uint8_t *)realloc(o->bases, o->cap_bases);
    }
    o->bases[o->n_bases++] = c;
}

/* parseRecord, reader.go:372-471.  Returns 0 ok, 1 unequal (shorter = seq longer than qual), 2 io.EOF.
 * On success the record's sequence has been appended to o->bases. */
static int parse_record(const fxbuf *b, int is_fastq, fxout *o, uint64_t rec_start, int *shorter) {
    const uint8_t *p = b->p;
    const size_t n = b->len;
    const uint64_t seq_begin = o->n_bases;
    size_t head_len = 0;
    uint64_t qual_src = 0;
    int have_qual_src = 0;
    size_t seq_len = 0, qual_len = 0;
    size_t j = 0;
    while (j < n && p[j] != '\n') j++;
    if (j < n && j > 0) { /* bytes.IndexByte(p, '\n') > 0, :379 */
        head_len = drop_cr(p, 0, j);
        size_t r = j + 1;
        if (!is_fastq) { /* :383-393 */
            for (;;) {
                size_t k = r;
                while (k < n && p[k] != '\n') k++;
                if (k < n) {
                    const size_t e = drop_cr(p, r, k);
                    for (size_t i = r; i < e; i++) out_base(o, p[i]);
                    r = k + 1;
                    continue;
                }
                const size_t e = drop_cr(p, r, n);
                for (size_t i = r; i < e; i++) out_base(o, p[i]);
                break;
            }
            seq_len = o->n_bases - seq_begin;
        } else { /* :395-417 */
            int is_qual = 0;
            for (;;) {
                size_t k = r;
                while (k < n && p[k] != '\n') k++;
                if (k < n) {
                    if (k > r && p[r] == '+' && !is_qual) {
                        is_qual = 1;
                    } else if (is_qual) {
                        if (!have_qual_src) { qual_src = b->src[r < n ? r : n - 1]; have_qual_src = 1; }
                        qual_len += drop_cr(p, r, k) - r;
                    } else {
                        const size_t e = drop_cr(p, r, k);
                        for (size_t i = r; i < e; i++) out_base(o, p[i]);
                    }
                    r = k + 1;
                    continue;
                }
                if (is_qual) {
                    if (!have_qual_src && r < n) { qual_src = b->src[r]; have_qual_src = 1; }
                    qual_len += drop_cr(p, r, n) - r;
                }
                break;
            }
            seq_len = o->n_bases - seq_begin;
            if (seq_len != qual_len) { /* :415-417 */
                *shorter = seq_len > qual_len;
                o->n_bases = seq_begin;
                return 1;
            }
        }
    } else { /* :420-424: head = dropCR(dropLF(p)) */
        size_t e = n;
        if (e > 0 && p[e - 1] == '\n') e--;
        head_len = drop_cr(p, 0, e);
    }
    if (head_len == 0 && seq_len == 0) { /* :437-439 */
        o->n_bases = seq_begin;
        return 2;
    }
    if (o->n_rec + 1 >= o->cap_rec) {
        o->cap_rec = o->cap_rec ? o->cap_rec * 2 : 1024;
        o->read_off = (uint64_t *)realloc(o->read_off, (o->cap_rec + 1) * 8);
        o->rec_off = (uint64_t *)realloc(o->rec_off, (o->cap_rec + 1) * 8);
        o->qual_off = (uint64_t *)realloc(o->qual_off, (o->cap_rec + 1) * 8);
        o->name_len = (uint64_t *)realloc(o->name_len, (o->cap_rec + 1) * 8);
    }
    o->read_off[o->n_rec] = seq_begin;
    o->rec_off[o->n_rec] = rec_start;
    o->qual_off[o->n_rec] = have_qual_src ? qual_src : 0;
    o->name_len[o->n_rec] = head_len;
    o->n_rec++;
    o->read_off[o->n_rec] = o->n_bases;
    return 0;
}

/* Read() until EOF, reader.go:233-369.  Returns the status; *format, *n_rec, *n_bases and the malloc'd arrays
 * (caller frees with fx_free) describe the records read before the status arose. */
int ora_fastx_parse(const uint8_t *t, uint64_t n, int *format, uint64_t *n_rec, uint64_t *n_bases, uint8_t **bases,
                    uint64_t **read_off, uint64_t **rec_off, uint64_t **qual_off, uint64_t **name_len) {
    fxout o;
    memset(&o, 0, sizeof(o));
    o.cap_rec = 1024;
    o.read_off = (uint64_t *)malloc((o.cap_rec + 1) * 8);
    o.rec_off = (uint64_t *)malloc((o.cap_rec + 1) * 8);
    o.qual_off = (uint64_t *)malloc((o.cap_rec + 1) * 8);
    o.name_len = (uint64_t *)malloc((o.cap_rec + 1) * 8);
    o.read_off[0] = 0;
    int status = FX_OK, is_fastq = 0;
    uint8_t delim = 0;
    uint64_t r = 0;
    *format = 0;
    /* :271-304 */
    {
        int found = 0;
        uint64_t pn = 0;
        for (uint64_t i = 0; i < n && !found; i++) {
            switch (t[i]) {
            case '>': is_fastq = 0; delim = '>'; r = i + 1; found = 1; break;
            case '@': is_fastq = 1; delim = '@'; r = i + 1; found = 1; break;
            case '\n':
                pn++;
                if (pn > 100 && i > 10240) { status = FX_ERR_NOT_FASTX; found = 2; }
                break;
            default: status = FX_ERR_NOT_FASTX; found = 2; break;
            }
        }
        if (found != 1) { /* nothing but newlines: no record (the Go reader ends with io.EOF) */
            goto done;
        }
    }
    *format = is_fastq ? FX_FASTQ : FX_FASTA;
    {
        fxbuf b;
        memset(&b, 0, sizeof(b));
        uint64_t rec_start = r - 1;
        for (;;) {
            /* bytes.IndexByte(buf[r:], delim), :310 */
            uint64_t j = r;
            while (j < n && t[j] != delim) j++;
            if (j < n) {
                const uint64_t i = j - r;
                uint8_t last;
                if (i > 0) last = t[j - 1];
                else last = b.len ? b.p[b.len - 1] : 0;
                if (last == '\n') {
                    if (i > 0) {
                        const uint64_t e = r + drop_cr(t + r, 0, (size_t)(i - 1));
                        fx_write(&b, t, r, e);
                    } else fx_push(&b, '\n', j ? j - 1 : 0);
                    int shorter = 0;
                    const int pr = parse_record(&b, is_fastq, &o, rec_start, &shorter);
                    if (is_fastq && pr == 1) {
                        if (shorter) { /* the '@' opened a quality line, :330-336 */
                            fx_push(&b, '\n', j - 1);
                            fx_push(&b, delim, j);
                            r = j + 1;
                            continue;
                        }
                        status = FX_ERR_BAD_FASTQ; /* ErrBadFASTQFormat, :338 */
                        break;
                    }
                    b.len = 0;
                    r = j + 1;
                    if (pr == 2) break; /* io.EOF returned with the record, :344: the caller stops */
                    rec_start = j;
                    continue;
                }
                fx_write(&b, t, r, j + 1); /* inline > / @, :346-349 */
                r = j + 1;
                continue;
            }
            fx_write(&b, t, r, n);
            {
                int shorter = 0;
                const int pr = parse_record(&b, is_fastq, &o, rec_start, &shorter); /* last part, :353-363 */
                if (pr == 1) status = FX_ERR_BAD_FASTQ; /* ErrUnequalSeqAndQual: "no any chance" */
            }
            break;
        }
        free(b.p);
        free(b.src);
    }
done:
    *n_rec = o.n_rec;
    *n_bases = o.n_bases;
    *bases = o.bases ? o.bases : (uint8_t *)malloc(1);
    *read_off = o.read_off;
    *rec_off = o.rec_off;
    *qual_off = o.qual_off;
    *name_len = o.name_len;
    return status;
}

void ora_fastx_free(void *p) { free(p); }
