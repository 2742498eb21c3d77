/*
 * oracle/sketch_oracle.c -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT.
 *
 * CPU restatement (plain C) of the sketching hot path of shenwei356/bio
 * (reference @ 7b48836e).  It exists so that the CUDA path can be checked
 * bit-for-bit; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (bio_b200/lib/libb200sketch.so) never links, loads or calls this file.
 *
 * Every state machine below follows the reference line by line (file:line
 * cited at each function, paths relative to /root/reference).  The arithmetic
 * that lives in un-vendored Go modules is restated from their published
 * algorithms:
 *   - github.com/will-rowe/nthash v0.4.0 (go.mod:15)   ntHash-1     PINNED by
 *     sketches/sketch_test.go:67-72 (tests/test_oracle_golden.py)
 *   - github.com/shenwei356/kmers v0.1.0 (go.mod:11)   Encode/MustRevComp
 *     parity unpinned by reference tests (count-only), arithmetic determined
 *     by sketches/iterator.go:736,740,754
 *   - github.com/twotwotwo/sorts@bf5c1f2b8553 (go.mod:14) Quicksort: comparator
 *     is Val only (sketches/sketch.go:506); TIE ORDER IN THE FIRST WINDOW IS
 *     PARITY-UNPINNED.  Policy here: stable (leftmost first); every call
 *     reports whether the first window held equal values so the parity report
 *     can state how many reads are affected.  ORA_SORT_GO14 selects a
 *     restatement of the Go<=1.5 stdlib quicksort the module is believed to
 *     derive from (unverified) for sensitivity analysis only.
 *   - github.com/zeebo/wyhash v0.0.1 (go.mod:16)       Hash(b, seed)
 *     PARITY UNPINNED: no reference test checks a protein hash value
 *     (sketches/iterator-protein_test.go:58 is count-only); restated from the
 *     published wyhash v1 layout.
 *   - seq.Translate / codon tables: PINNED by seq/codon_tables_test.go:26-134.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200sk_codon_data.h"

/* error codes: same numeric values as include/b200sketch.h */
#define ORA_OK 0
#define ORA_ERR_INVALID_K (-1)      /* sketches/iterator.go:35 ErrInvalidK   */
#define ORA_ERR_SHORT_SEQ (-2)      /* sketches/iterator.go:41 ErrShortSeq   */
#define ORA_ERR_INVALID_W (-3)      /* sketches/sketch.go:36  ErrInvalidW    */
#define ORA_ERR_INVALID_S (-4)      /* sketches/sketch.go:33  ErrInvalidS    */
#define ORA_ERR_ILLEGAL_BASE (-5)   /* sketches/iterator.go:44 ErrIllegalBase */
#define ORA_ERR_K_OVERFLOW (-6)     /* kmers.ErrKOverflow (k > 32)           */
#define ORA_ERR_INVALID_FRAME (-7)  /* seq/codon_tables.go:209-211           */
#define ORA_ERR_CODON_TABLE (-8)    /* seq/seq.go Translate: unknown table   */
#define ORA_ERR_TRANSLATE_SHORT (-9)/* seq/codon_tables.go:206-208           */
#define ORA_ERR_INVALID_CODON (-10) /* seq.ErrInvalidDNABase w/o allowUnknown */
#define ORA_ERR_INVALID_M (-11)     /* sketches/iterator.go:50 ErrInvalidM     */
#define ORA_ERR_INVALID_SCALE (-12) /* sketches/iterator.go:53 ErrInvalidScale */
#define ORA_ERR_K_TOO_LARGE (-13)   /* sketches/iterator.go:47 ErrKTooLarge    */

#define ORA_SORT_STABLE 0
#define ORA_SORT_GO14 1

/* ------------------------------------------------------------------ */
/* ntHash-1 (will-rowe/nthash v0.4.0), spec: SURVEY.md 8c              */
/* ------------------------------------------------------------------ */

#define SEED_A 0x3c8bfbb395c60474ULL
#define SEED_C 0x3193c18562a02b4cULL
#define SEED_G 0x20323ed082572324ULL
#define SEED_T 0x295549f54be24456ULL
#define SEED_N 0x0000000000000000ULL

static uint64_t seed_tab[256];
static int tables_ready = 0;

static inline uint64_t rol64(uint64_t v, unsigned n) {
    n &= 63u;
    return n ? (v << n) | (v >> (64u - n)) : v;
}
static inline uint64_t ror64(uint64_t v, unsigned n) {
    n &= 63u;
    return n ? (v >> n) | (v << (64u - n)) : v;
}

static void build_codon_tables(void);

static void ora_init_tables(void) {
    if (tables_ready) return;
    memset(seed_tab, 0, sizeof(seed_tab));
    /* indices 0..7: {N,T,N,G,A,A,N,C}: the complement lookup is seed_tab[b & 7] */
    seed_tab[0] = SEED_N; seed_tab[1] = SEED_T; seed_tab[2] = SEED_N; seed_tab[3] = SEED_G;
    seed_tab[4] = SEED_A; seed_tab[5] = SEED_A; seed_tab[6] = SEED_N; seed_tab[7] = SEED_C;
    seed_tab['A'] = seed_tab['a'] = SEED_A;
    seed_tab['C'] = seed_tab['c'] = SEED_C;
    seed_tab['G'] = seed_tab['g'] = SEED_G;
    seed_tab['T'] = seed_tab['t'] = SEED_T;
    seed_tab['U'] = seed_tab['u'] = SEED_T;
    build_codon_tables();
    tables_ready = 1;
}

__attribute__((constructor)) static void ora_ctor(void) { ora_init_tables(); }

typedef struct {
    const uint8_t *seq;
    size_t len;
    unsigned k;
    uint64_t fh, rh;
    size_t cur, max; /* cur = index of the next k-mer; max = len-k+1 */
} nthi_t;

/* nthash.NewHasher: error when k > len; first hash computed in O(k). */
static int nthi_init(nthi_t *h, const uint8_t *seq, size_t len, unsigned k) {
    if ((size_t)k > len) return -1;
    uint64_t fh = 0, rh = 0;
    for (unsigned i = 0; i < k; i++) {
        fh = rol64(fh, 1) ^ seed_tab[seq[i]];
        rh = rol64(rh, 1) ^ seed_tab[seq[k - 1 - i] & 7];
    }
    h->seq = seq; h->len = len; h->k = k; h->fh = fh; h->rh = rh;
    h->cur = 0; h->max = len - (k - 1);
    return 0;
}

/* (*NTHi).Next(canonical): yields len-k+1 values, then ok=false. */
static inline int nthi_next(nthi_t *h, int canonical, uint64_t *out) {
    if (h->cur >= h->max) { *out = 0; return 0; }
    if (h->cur != 0) {
        uint8_t prev = h->seq[h->cur - 1];
        uint8_t end = h->seq[h->cur + h->k - 1];
        h->fh = rol64(h->fh, 1) ^ rol64(seed_tab[prev], h->k) ^ seed_tab[end];
        h->rh = ror64(h->rh, 1) ^ ror64(seed_tab[prev & 7], 1) ^ rol64(seed_tab[end & 7], h->k - 1);
    }
    h->cur++;
    if (canonical) *out = (h->rh < h->fh) ? h->rh : h->fh;
    else *out = h->fh;
    return 1;
}

/* circular: "seq2 = S + S[0:k-1]" (sketches/iterator.go:642-646, sketch.go:106-110,163-167) */
static uint8_t *make_seq2(const uint8_t *seq, size_t len, int k, int circular, size_t *len2) {
    size_t extra = circular ? (size_t)(k - 1) : 0;
    uint8_t *s2 = (uint8_t *)malloc(len + extra + 1);
    memcpy(s2, seq, len);
    if (extra) memcpy(s2 + len, seq, extra);
    *len2 = len + extra;
    return s2;
}

/* ------------------------------------------------------------------ */
/* NewHashIterator / NextHash  (sketches/iterator.go:615-665)          */
/* ------------------------------------------------------------------ */
int64_t ora_hash_iterator(const uint8_t *seq, size_t len, int k, int canonical, int circular,
                          uint64_t *out, int *err) {
    *err = ORA_OK;
    if (k < 1) { *err = ORA_ERR_INVALID_K; return 0; }        /* :616 */
    if (len < (size_t)k) { *err = ORA_ERR_SHORT_SEQ; return 0; } /* :619 */
    size_t len2;
    uint8_t *s2 = make_seq2(seq, len, k, circular, &len2);
    nthi_t h;
    nthi_init(&h, s2, len2, (unsigned)k);
    int64_t n = 0;
    uint64_t v;
    while (nthi_next(&h, canonical, &v)) out[n++] = v;
    free(s2);
    return n;
}

/* ------------------------------------------------------------------ */
/* NewSimHashIterator / NextSimHash (sketches/iterator.go:113-612)      */
/* One 64-bit SimHash per k-mer over its k-m+1 m-mer ntHashes, after the */
/* FracMinHash filter (hash > MaxUint64/scale -> dropped, :281,443).     */
/* int16 counters and the sign-bit decode are kept literally (:359-424). */
/* ------------------------------------------------------------------ */
int64_t ora_simhash_iterator(const uint8_t *seq, size_t len, int k, int m, int scale, int canonical,
                             int circular, uint64_t *out, int *err) {
    *err = ORA_OK;
    if (k < 1) { *err = ORA_ERR_INVALID_K; return 0; }                    /* :114 */
    if (k >= 65535) { *err = ORA_ERR_K_TOO_LARGE; return 0; }             /* :117 */
    if (m < 4 || m > k) { *err = ORA_ERR_INVALID_M; return 0; }           /* :121 */
    if (scale < 1 || scale > k - m + 1) { *err = ORA_ERR_INVALID_SCALE; return 0; } /* :124 */
    if (len < (size_t)k) { *err = ORA_ERR_SHORT_SEQ; return 0; }          /* :128 */
    size_t length;
    uint8_t *s2 = make_seq2(seq, len, k, circular, &length);
    int64_t end = (int64_t)length - k + 1, idx = 0, n = 0;                /* :151 */
    nthi_t h;
    nthi_init(&h, s2, length, (unsigned)m);                               /* :159 */
    int16_t sum[64];
    memset(sum, 0, sizeof(sum));
    int16_t n_pos = 0, thr;
    int em = k - m, pre_i = 0, first = 1;
    uint64_t *hashes = (uint64_t *)calloc((size_t)em + 1, sizeof(uint64_t));
    const int frac = scale > 1;                                           /* :180 */
    const uint64_t max_hash = frac ? UINT64_MAX / (uint64_t)scale : UINT64_MAX;
    uint64_t hv, code;
    while (idx != end) {                                                  /* :196 */
        if (!first) {
            uint64_t pre = hashes[pre_i];                                 /* :205 */
            if (pre > 0) {
                n_pos--;
                for (int b = 0; b < 64; b++) sum[b] -= (int16_t)(pre >> (63 - b) & 1);
            }
            nthi_next(&h, canonical, &hv);                                /* :279 */
            if (frac && hv > max_hash) hv = 0;                            /* :281 */
            else if (hv > 0) n_pos++;
            hashes[pre_i] = hv;                                           /* :288 */
            if (hv > 0) for (int b = 0; b < 64; b++) sum[b] += (int16_t)(hv >> (63 - b) & 1);
            pre_i = (pre_i == em) ? 0 : pre_i + 1;                        /* :432-436 */
        } else {
            n_pos = 0;
            for (int j = 0; j <= em; j++) {                               /* :438-520 */
                nthi_next(&h, canonical, &hv);
                if (frac && hv > max_hash) { hashes[j] = 0; continue; }
                hashes[j] = hv;
                if (hv == 0) continue;
                n_pos++;
                for (int b = 0; b < 64; b++) sum[b] += (int16_t)(hv >> (63 - b) & 1);
            }
            pre_i = 0;
            first = 0;
        }
        code = 0;
        thr = (int16_t)((n_pos + 1) / 2);                                 /* :357,528 */
        if (n_pos > 0)
            for (int b = 0; b < 64; b++)
                code |= (uint64_t)(((int16_t)(sum[b] - thr) >> 15 & 1) ^ 1) << (63 - b);
        out[n++] = code;
        idx++;
    }
    free(hashes); free(s2);
    return n;
}

/* ------------------------------------------------------------------ */
/* base2bit (sketches/kmers.go:23-40), kmers.Encode / MustRevComp      */
/* ------------------------------------------------------------------ */
static inline unsigned base2bit(uint8_t b) {
    switch (b) {
    case 'A': case 'a': case 'D': case 'd': case 'H': case 'h': case 'M': case 'm':
    case 'N': case 'n': case 'R': case 'r': case 'V': case 'v': case 'W': case 'w':
        return 0;
    case 'B': case 'b': case 'C': case 'c': case 'S': case 's': case 'Y': case 'y':
        return 1;
    case 'G': case 'g': case 'K': case 'k':
        return 2;
    case 'T': case 't': case 'U': case 'u':
        return 3;
    default:
        return 4;
    }
}

/* kmers.Encode: fold code = code<<2 | bits, first base most significant;
 * error on illegal base or k > 32. */
static int kmers_encode(const uint8_t *kmer, int k, uint64_t *code) {
    if (k <= 0 || k > 32) return ORA_ERR_K_OVERFLOW;
    uint64_t c = 0;
    for (int i = 0; i < k; i++) {
        unsigned b = base2bit(kmer[i]);
        if (b == 4) return ORA_ERR_ILLEGAL_BASE;
        c = (c << 2) | b;
    }
    *code = c;
    return ORA_OK;
}

/* kmers.MustRevComp: complement (3 - bits) and reverse the k 2-bit groups. */
static uint64_t kmers_revcomp(uint64_t code, int k) {
    uint64_t c = 0;
    for (int i = 0; i < k; i++) {
        c = (c << 2) | ((code & 3ULL) ^ 3ULL);
        code >>= 2;
    }
    return c;
}

/* seq.Alphabet.PairLetter tables (seq/alphabet.go:313-325, 353-399).
 * alphabet: 0 DNAredundant, 1 DNA, 2 RNAredundant, 3 RNA, 4 Unlimit.
 * Letters outside the alphabet are returned unchanged (error ignored,
 * seq/seq.go:389-391). */
static void build_pair_lut(int alphabet, uint8_t lut[256]) {
    for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i;
    const char *l = 0, *p = 0;
    switch (alphabet) {
    case 0: l = "acgtryswkmbdhvACGTRYSWKMBDHV"; p = "tgcayrswmkvhdbTGCAYRSWMKVHDB"; break;
    case 1: l = "acgtACGT"; p = "tgcaTGCA"; break;
    case 2: l = "acguryswkmbdhvACGURYSWKMBDHV"; p = "ugcayrswmkvhdbUGCAYRSWMKVHDB"; break;
    case 3: l = "acguACGU"; p = "ugcaUGCA"; break;
    default: return;
    }
    for (int i = 0; l[i]; i++) lut[(uint8_t)l[i]] = (uint8_t)p[i];
}

void ora_pair_lut(int alphabet, uint8_t *lut256) { build_pair_lut(alphabet, lut256); }

/* ------------------------------------------------------------------ */
/* NewKmerIterator / NextKmer (sketches/iterator.go:668-759)           */
/* Non-canonical mode walks the forward strand, then reverse-complements
 * the sequence and walks it again (:713-723).  The caller's buffer is
 * NOT modified here (we work on a copy); the mutation is documented.   */
/* Returns number of codes emitted; on illegal base *err is set and
 * *err_idx is the k-mer index (on the strand being walked) that failed. */
/* ------------------------------------------------------------------ */
int64_t ora_kmer_iterator(const uint8_t *seq, size_t len, int k, int canonical, int circular,
                          int alphabet, uint64_t *out, int *err, int64_t *err_idx) {
    *err = ORA_OK;
    if (err_idx) *err_idx = -1;
    if (k < 1) { *err = ORA_ERR_INVALID_K; return 0; }          /* :669 */
    if (len < (size_t)k) { *err = ORA_ERR_SHORT_SEQ; return 0; } /* :672 */
    size_t length;
    uint8_t *s = make_seq2(seq, len, k, circular, &length);
    int64_t end = (int64_t)length - k + 1;                       /* :693 */
    int kp1 = k - 1;
    /* Go: mask1 = (1 << (kP1*2)) - 1; a shift count >= 64 yields 0 in Go */
    uint64_t mask1 = (kp1 * 2 >= 64) ? (uint64_t)0 - 1 : ((1ULL << (kp1 * 2)) - 1); /* :699 */
    unsigned mask2 = (unsigned)kp1 * 2;                          /* :700 */
    int finished = 0, revcom = 0, first = 1;
    int64_t idx = 0, n = 0;
    uint64_t pre = 0, pre_rc = 0, code = 0, code_rc = 0;
    uint8_t pair[256];
    build_pair_lut(alphabet, pair);
    while (!finished) {
        if (idx == end) {                                        /* :713 */
            if (canonical || revcom) { finished = 1; break; }
            /* RevComInplace: reverse then complement (seq/seq.go:350-352) */
            for (size_t i = 0, j = length - 1; i < j; i++, j--) { uint8_t t = s[i]; s[i] = s[j]; s[j] = t; }
            for (size_t i = 0; i < length; i++) s[i] = pair[s[i]];
            idx = 0; revcom = 1; first = 1;
        }
        const uint8_t *kmer = s + idx;
        int e = ORA_OK;
        if (!first) {
            uint64_t cb = base2bit(kmer[kp1]);                   /* :728 */
            if (cb == 4) e = ORA_ERR_ILLEGAL_BASE;
            code = ((pre & mask1) << 2) | cb;                    /* :736 */
            code_rc = ((mask2 >= 64) ? 0 : ((cb ^ 3) << mask2)) | (pre_rc >> 2); /* :740 */
        } else {
            e = kmers_encode(kmer, k, &code);                    /* :742 */
            if (e == ORA_OK) code_rc = kmers_revcomp(code, k);   /* :743 */
            first = 0;
        }
        if (e != ORA_OK) {                                       /* :746-748 */
            *err = e;
            if (err_idx) *err_idx = idx;
            break;
        }
        pre = code; pre_rc = code_rc; idx++;
        uint64_t o = code;
        if (canonical && code > code_rc) o = code_rc;            /* :754-756 */
        out[n++] = o;
    }
    free(s);
    return n;
}

/* ------------------------------------------------------------------ */
/* first-window sort (twotwotwo/sorts Quicksort; comparator = Val)     */
/* ------------------------------------------------------------------ */
typedef struct { int64_t idx; uint64_t val; } iv_t;

static int has_equal_values(const iv_t *b, int n) {
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            if (b[i].val == b[j].val) return 1;
    return 0;
}

static void sort_stable(iv_t *b, int n) { /* insertion sort: stable, leftmost first */
    for (int i = 1; i < n; i++) {
        iv_t x = b[i];
        int j = i;
        while (j > 0 && x.val < b[j - 1].val) { b[j] = b[j - 1]; j--; }
        b[j] = x;
    }
}

/* Go <= 1.5 sort.Sort restated (believed ancestor of twotwotwo/sorts; unverified). */
#define LESS(i, j) (d[i].val < d[j].val)
#define SWAP(i, j) do { iv_t _t = d[i]; d[i] = d[j]; d[j] = _t; } while (0)
static void go14_insertion(iv_t *d, int a, int b) {
    for (int i = a + 1; i < b; i++)
        for (int j = i; j > a && LESS(j, j - 1); j--) SWAP(j, j - 1);
}
static void go14_sift(iv_t *d, int lo, int hi, int first) {
    int root = lo;
    for (;;) {
        int child = 2 * root + 1;
        if (child >= hi) return;
        if (child + 1 < hi && LESS(first + child, first + child + 1)) child++;
        if (!LESS(first + root, first + child)) return;
        SWAP(first + root, first + child);
        root = child;
    }
}
static void go14_heapsort(iv_t *d, int a, int b) {
    int first = a, lo = 0, hi = b - a;
    for (int i = (hi - 1) / 2; i >= 0; i--) go14_sift(d, i, hi, first);
    for (int i = hi - 1; i >= 0; i--) { SWAP(first, first + i); go14_sift(d, lo, i, first); }
}
static void go14_med3(iv_t *d, int a, int b, int c) {
    int m0 = b, m1 = a, m2 = c;
    if (LESS(m1, m0)) SWAP(m1, m0);
    if (LESS(m2, m1)) SWAP(m2, m1);
    if (LESS(m1, m0)) SWAP(m1, m0);
}
static void go14_swaprange(iv_t *d, int a, int b, int n) {
    for (int i = 0; i < n; i++) SWAP(a + i, b + i);
}
static void go14_pivot(iv_t *d, int lo, int hi, int *midlo, int *midhi) {
    int m = lo + (hi - lo) / 2;
    if (hi - lo > 40) {
        int s = (hi - lo) / 8;
        go14_med3(d, lo, lo + s, lo + 2 * s);
        go14_med3(d, m, m - s, m + s);
        go14_med3(d, hi - 1, hi - 1 - s, hi - 1 - 2 * s);
    }
    go14_med3(d, lo, m, hi - 1);
    int pivot = lo;
    int a = lo + 1, b = lo + 1, c = hi, e = hi;
    for (;;) {
        while (b < c) {
            if (LESS(b, pivot)) b++;
            else if (!LESS(pivot, b)) { SWAP(a, b); a++; b++; }
            else break;
        }
        while (b < c) {
            if (LESS(pivot, c - 1)) c--;
            else if (!LESS(c - 1, pivot)) { SWAP(c - 1, e - 1); c--; e--; }
            else break;
        }
        if (b >= c) break;
        SWAP(b, c - 1);
        b++; c--;
    }
    int n = (b - a < a - lo) ? b - a : a - lo;
    go14_swaprange(d, lo, b - n, n);
    n = (hi - e < e - c) ? hi - e : e - c;
    go14_swaprange(d, c, hi - n, n);
    *midlo = lo + b - a;
    *midhi = hi - (e - c);
}
static void go14_quicksort(iv_t *d, int a, int b, int depth) {
    while (b - a > 7) {
        if (depth == 0) { go14_heapsort(d, a, b); return; }
        depth--;
        int mlo, mhi;
        go14_pivot(d, a, b, &mlo, &mhi);
        if (mlo - a < b - mhi) { go14_quicksort(d, a, mlo, depth); a = mhi; }
        else { go14_quicksort(d, mhi, b, depth); b = mlo; }
    }
    if (b - a > 1) go14_insertion(d, a, b);
}
#undef LESS
#undef SWAP

static void first_window_sort(iv_t *b, int n, int policy) {
    if (policy == ORA_SORT_GO14) {
        int depth = 0;
        for (int i = n; i > 0; i >>= 1) depth++;
        go14_quicksort(b, 0, n, depth * 2);
    } else {
        sort_stable(b, n);
    }
}

/* The hand-rolled binary search + insert of sketches/sketch.go:261-295 and
 * :371-405.  buf holds r elements on entry, r+1 on exit. */
static void insert_sorted(iv_t *buf, int r, int64_t idx, uint64_t code) {
    int flag = 0, i = 0;
    int b = 0, e = r - 1, t;
    for (;;) {
        t = b + (e - b) / 2;
        if (code < buf[t].val) {
            e = t - 1;
            if (e <= b) { flag = 1; i = b; break; }
        } else {
            b = t + 1;
            if (b >= r) { flag = 0; break; }
            if (b >= e) { flag = 1; i = e; break; }
        }
    }
    if (!flag) {
        buf[r].idx = idx; buf[r].val = code;          /* biggest: append */
    } else {
        if (code >= buf[i].val) i++;                   /* "have to check again" */
        memmove(&buf[i + 1], &buf[i], (size_t)(r - i) * sizeof(iv_t));
        buf[i].idx = idx; buf[i].val = code;
    }
}

static inline void fifo_pop(int64_t *f, size_t *n) {
    for (size_t i = 1; i < *n; i++) f[i - 1] = f[i];
    if (*n) (*n)--;
}

/* evict the element whose Idx == target (sketch.go:250-258, 355-363); len r+1 -> r */
static void evict_idx(iv_t *buf, int r, int64_t target) {
    for (int i = 0; i <= r; i++) {
        if (buf[i].idx == target) {
            if (i < r) memmove(&buf[i], &buf[i + 1], (size_t)(r - i) * sizeof(iv_t));
            return;
        }
    }
}

/* ------------------------------------------------------------------ */
/* NewMinimizerSketch / NextMinimizer (sketches/sketch.go:85-138,205-309) */
/* out_idx receives Index() (= mI, sketch.go:488-491) per emitted value. */
/* ------------------------------------------------------------------ */
int64_t ora_minimizer(const uint8_t *seq, size_t len, int k, int w, int circular, int sort_policy,
                      uint64_t *out_val, int64_t *out_idx, int *err, int *first_window_tie) {
    *err = ORA_OK;
    if (first_window_tie) *first_window_tie = 0;
    if (k < 1) { *err = ORA_ERR_INVALID_K; return 0; }                    /* :86 */
    if (w < 1) { /* w > 2^31-1 cannot be represented in the int32 ABI */ *err = ORA_ERR_INVALID_W; return 0; } /* :89 */
    if ((int64_t)len < (int64_t)k + w - 1) { *err = ORA_ERR_SHORT_SEQ; return 0; }  /* :92 */
    size_t len2;
    uint8_t *s2 = make_seq2(seq, len, k, circular, &len2);
    int64_t idx = 0, end = (int64_t)len2 - 1;                              /* :115-116 */
    int r = w - 1;                                                         /* :117 */
    int skip = (w == 1);                                                   /* :103 */
    nthi_t h;
    nthi_init(&h, s2, len2, (unsigned)k);
    iv_t *buf = (iv_t *)malloc(((size_t)w + 2) * sizeof(iv_t));
    int blen = 0;
    int64_t pre_min_idx = -1, n = 0;
    uint64_t code;
    for (;;) {
        if (idx > end) break;                                              /* :207 */
        if (!nthi_next(&h, 1, &code)) break;                               /* :212-216 */
        if (skip) {                                                        /* :218-222 */
            out_val[n] = code; out_idx[n] = idx; n++; idx++;
            continue;
        }
        if (idx < r) {                                                     /* :225-230 */
            buf[blen].idx = idx; buf[blen].val = code; blen++;
            idx++;
            continue;
        }
        if (idx == r) {                                                    /* :233-245 */
            buf[blen].idx = idx; buf[blen].val = code; blen++;
            if (first_window_tie && has_equal_values(buf, blen)) *first_window_tie = 1;
            first_window_sort(buf, blen, sort_policy);
            out_val[n] = buf[0].val; out_idx[n] = buf[0].idx; n++;
            pre_min_idx = buf[0].idx;
            idx++;
            continue;
        }
        evict_idx(buf, r, idx - w);                                        /* :250-258 */
        insert_sorted(buf, r, idx, code);                                  /* :261-295 */
        if (buf[0].idx == pre_min_idx) { idx++; continue; }                /* :297-301 */
        out_val[n] = buf[0].val; out_idx[n] = buf[0].idx; n++;             /* :303-307 */
        pre_min_idx = buf[0].idx;
        idx++;
    }
    free(buf);
    free(s2);
    return n;
}

/* ------------------------------------------------------------------ */
/* NewSyncmerSketch / NextSyncmer (sketches/sketch.go:142-202,312-477)  */
/* out_idx receives Index() (= idx-1 at return, sketch.go:492).         */
/* ------------------------------------------------------------------ */
int64_t ora_syncmer(const uint8_t *seq, size_t len, int k, int s, int circular, int sort_policy,
                    uint64_t *out_val, int64_t *out_idx, int *err, int *first_window_tie) {
    *err = ORA_OK;
    if (first_window_tie) *first_window_tie = 0;
    if (k < 1) { *err = ORA_ERR_INVALID_K; return 0; }                    /* :143 */
    if (s > k || s == 0) { *err = ORA_ERR_INVALID_S; return 0; }           /* :146 */
    if (s < 0) { *err = ORA_ERR_INVALID_S; return 0; }                     /* Go: uint(s) would make NewHasher fail */
    if ((int64_t)len < (int64_t)k * 2 - s - 1) { *err = ORA_ERR_SHORT_SEQ; return 0; } /* :149 */
    size_t len2;
    uint8_t *s2 = make_seq2(seq, len, k, circular, &len2);
    int64_t idx = 0;
    int64_t end = (int64_t)len2 - 2 * (int64_t)k + s + 1;                  /* :173 */
    int r = 2 * k - s - 1 - s;                                             /* :174 */
    int kms = k - s, w = k - s;                                            /* :175-176 */
    int skip = (s == k);                                                   /* :160 */
    nthi_t h, hs;
    if (nthi_init(&h, s2, len2, (unsigned)k) != 0) { *err = ORA_ERR_SHORT_SEQ; free(s2); return 0; }
    nthi_init(&hs, s2, len2, (unsigned)s);
    iv_t *buf = (iv_t *)malloc(((size_t)(2 * kms) + 2) * sizeof(iv_t));
    int blen = 0;
    /* preMinIdxs FIFO: at most a handful of entries; grow as needed */
    size_t fcap = 64, flen = 0;
    int64_t *fifo = (int64_t *)malloc(fcap * sizeof(int64_t));
    int64_t pre_min_idx = -1, n = 0;
    uint64_t code, v;
    for (;;) {
        if (idx > end) break;                                              /* :314 */
        if (!nthi_next(&h, 1, &code)) break;                               /* :319-323 */
        if (skip) { out_val[n] = code; out_idx[n] = idx; n++; idx++; continue; } /* :328-331 */
        int late = (flen > 0 && idx == fifo[0]);                           /* :333-338 */
        if (idx == 0) {                                                    /* :341-351 */
            int fail = 0;
            for (int64_t i = idx; i <= idx + r; i++) {
                if (!nthi_next(&hs, 1, &v)) { fail = 1; break; }
                buf[blen].idx = i; buf[blen].val = v; blen++;
            }
            if (fail) break;
            if (first_window_tie && has_equal_values(buf, blen)) *first_window_tie = 1;
            first_window_sort(buf, blen, sort_policy);
        } else {
            evict_idx(buf, r, idx - 1);                                    /* :355-363 */
            if (!nthi_next(&hs, 1, &v)) break;                             /* :367-370 */
            insert_sorted(buf, r, idx + r, v);                             /* :371-405 */
        }
        int64_t mI = buf[0].idx;                                           /* :408-409 */
        int64_t bidx = (mI - idx < w) ? mI : mI - kms;                     /* :414-420 */

        if (flen > 0 && bidx == fifo[0]) {                                 /* :425-440 */
            if (late) {
                fifo_pop(fifo, &flen);
                idx++; pre_min_idx = bidx;
                out_val[n] = code; out_idx[n] = idx - 1; n++;
                continue;
            }
            idx++;
            continue;
        }
        if (late) {                                                        /* :442-455 */
            fifo_pop(fifo, &flen);
            if (pre_min_idx != bidx) {
                if (flen == fcap) { fcap *= 2; fifo = (int64_t *)realloc(fifo, fcap * sizeof(int64_t)); }
                fifo[flen++] = bidx;
            }
            idx++; pre_min_idx = bidx;
            out_val[n] = code; out_idx[n] = idx - 1; n++;
            continue;
        }
        if (bidx == idx) {                                                 /* :458-468 */
            if (flen > 0) { fifo_pop(fifo, &flen); }
            idx++; pre_min_idx = bidx;
            out_val[n] = code; out_idx[n] = idx - 1; n++;
            continue;
        }
        if (pre_min_idx != bidx) {                                         /* :470-472 */
            if (flen == fcap) { fcap *= 2; fifo = (int64_t *)realloc(fifo, fcap * sizeof(int64_t)); }
            fifo[flen++] = bidx;
        }
        idx++; pre_min_idx = bidx;                                         /* :474-475 */
    }
    free(fifo);
    free(buf);
    free(s2);
    return n;
}

/* ------------------------------------------------------------------ */
/* Independent closed forms -- second implementation used */
/* by property tests: literal state machine == closed form.            */
/* ------------------------------------------------------------------ */
int64_t ora_minimizer_closed(const uint8_t *seq, size_t len, int k, int w, int circular,
                             uint64_t *out_val, int64_t *out_idx, int *err) {
    *err = ORA_OK;
    if (k < 1) { *err = ORA_ERR_INVALID_K; return 0; }
    if (w < 1) { *err = ORA_ERR_INVALID_W; return 0; }
    if ((int64_t)len < (int64_t)k + w - 1) { *err = ORA_ERR_SHORT_SEQ; return 0; }
    size_t len2;
    uint8_t *s2 = make_seq2(seq, len, k, circular, &len2);
    int64_t nk = (int64_t)len2 - k + 1;
    uint64_t *h = (uint64_t *)malloc((size_t)nk * sizeof(uint64_t));
    int e2;
    ora_hash_iterator(s2, len2, k, 1, 0, h, &e2);
    int64_t n = 0, prev = -1;
    for (int64_t i = 0; i + w <= nk; i++) {
        int64_t p = i;
        for (int64_t j = i + 1; j < i + w; j++) if (h[j] < h[p]) p = j; /* leftmost min */
        if (p != prev) { out_val[n] = h[p]; out_idx[n] = p; n++; prev = p; }
    }
    free(h); free(s2);
    return n;
}

int64_t ora_syncmer_closed(const uint8_t *seq, size_t len, int k, int s, int circular,
                           uint64_t *out_val, int64_t *out_idx, int *err) {
    *err = ORA_OK;
    if (k < 1) { *err = ORA_ERR_INVALID_K; return 0; }
    if (s > k || s <= 0) { *err = ORA_ERR_INVALID_S; return 0; }
    if ((int64_t)len < (int64_t)k * 2 - s - 1) { *err = ORA_ERR_SHORT_SEQ; return 0; }
    size_t len2;
    uint8_t *s2 = make_seq2(seq, len, k, circular, &len2);
    int64_t nk = (int64_t)len2 - k + 1, ns = (int64_t)len2 - s + 1;
    int64_t end = (int64_t)len2 - 2 * (int64_t)k + s + 1;
    uint64_t *hk = (uint64_t *)malloc((size_t)(nk > 0 ? nk : 1) * sizeof(uint64_t));
    uint64_t *hs = (uint64_t *)malloc((size_t)ns * sizeof(uint64_t));
    int e2;
    int64_t n = 0;
    if (nk <= 0) { free(hk); free(hs); free(s2); return 0; }
    ora_hash_iterator(s2, len2, k, 1, 0, hk, &e2);
    ora_hash_iterator(s2, len2, s, 1, 0, hs, &e2);
    if (s == k) {
        for (int64_t i = 0; i <= end && i < nk; i++) { out_val[n] = hk[i]; out_idx[n] = i; n++; }
    } else {
        int d = k - s;
        int64_t prev = -1;
        for (int64_t idx = 0; idx <= end; idx++) {
            int64_t m = idx;
            for (int64_t j = idx + 1; j < idx + 2 * d; j++) if (hs[j] < hs[m]) m = j;
            int64_t b = (m - idx < d) ? m : m - d;
            if (b != prev && b <= end) { out_val[n] = hk[b]; out_idx[n] = b; n++; }
            prev = b;
        }
    }
    free(hk); free(hs); free(s2);
    return n;
}

/* ------------------------------------------------------------------ */
/* Codon tables / Translate (seq/codon_tables.go, seq/ambiguous_bases.go) */
/* ------------------------------------------------------------------ */
static uint8_t codon_mat[32][16][16][16]; /* indexed by table id */
static uint8_t codon_have[32];

/* base2code (seq/ambiguous_bases.go:28-67); -1 = ErrInvalidDNABase */
static int base2code(uint8_t b) {
    switch (b) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 4;
    case 'T': case 't': case 'U': case 'u': return 8;
    case 'N': case 'n': return 15;
    case 'M': case 'm': return 3;
    case 'R': case 'r': return 5;
    case 'W': case 'w': return 9;
    case 'S': case 's': return 6;
    case 'Y': case 'y': return 10;
    case 'K': case 'k': return 12;
    case 'V': case 'v': return 7;
    case 'H': case 'h': return 11;
    case 'D': case 'd': return 13;
    case 'B': case 'b': return 14;
    case ' ': case '*': case '-': return 0;
    default: return -1;
    }
}

/* AmbCodes2Codes (seq/ambiguous_bases.go:178-197): only these keys exist;
 * the value set is every non-zero sub-mask of the key. */
static int amb_key_exists(int c) { return c >= 1 && c <= 15; }

/* codonTableFromText (seq/codon_tables.go:316-427) */
static void expand_axis(uint8_t t[16][16][16], int axis) {
    for (int i = 1; i < 16; i++) {
        for (int j = 1; j < 16; j++) {
            /* group the third coordinate by amino acid */
            int mask_of[256];
            memset(mask_of, 0, sizeof(mask_of));
            for (int k = 1; k < 16; k++) {
                uint8_t aa = axis == 3 ? t[i][j][k] : axis == 2 ? t[i][k][j] : t[k][i][j];
                if (aa) mask_of[aa] |= k; /* Codes2AmbCode = OR of codes */
            }
            for (int aa = 1; aa < 256; aa++) {
                int amb = mask_of[aa];
                if (!amb || !amb_key_exists(amb)) continue;
                for (int c = 1; c < 16; c++) {
                    if ((c & amb) != c) continue; /* sub-masks of amb */
                    if (axis == 3) t[i][j][c] = (uint8_t)aa;
                    else if (axis == 2) t[i][c][j] = (uint8_t)aa;
                    else t[c][i][j] = (uint8_t)aa;
                }
            }
        }
    }
}

static void build_codon_tables(void) {
    static const char order[4] = {'T', 'C', 'A', 'G'};
    memset(codon_mat, 0, sizeof(codon_mat));
    memset(codon_have, 0, sizeof(codon_have));
    for (int ti = 0; ti < B200SK_N_CODON_ROWS; ti++) {
        int id = B200SK_CODON_ROWS[ti].id;
        const char *aas = B200SK_CODON_ROWS[ti].aas;
        uint8_t(*t)[16][16] = codon_mat[id];
        for (int c = 0; c < 64; c++) {
            int i = base2code((uint8_t)order[c >> 4]);
            int j = base2code((uint8_t)order[(c >> 2) & 3]);
            int k = base2code((uint8_t)order[c & 3]);
            t[i][j][k] = (uint8_t)aas[c];
        }
        expand_axis(t, 3); /* base3 (:345-371) */
        expand_axis(t, 2); /* base2 (:373-399) */
        expand_axis(t, 1); /* base1 (:401-425) */
        codon_have[id] = 1;
    }
}

/* CodonTable.Get (seq/codon_tables.go:152-170) */
static int codon_get(int table, const uint8_t codon[3], int allow_unknown, uint8_t *aa) {
    int i = base2code(codon[0]), j = base2code(codon[1]), k = base2code(codon[2]);
    if (i < 0 || j < 0 || k < 0) {
        if (allow_unknown) { *aa = 'X'; return ORA_OK; }
        return ORA_ERR_INVALID_CODON;
    }
    if (codon[0] == '-' && codon[1] == '-' && codon[2] == '-') { *aa = '-'; return ORA_OK; }
    uint8_t a = codon_mat[table][i][j][k];
    if (a == 0) a = 'X';
    *aa = a;
    return ORA_OK;
}

/* CodonTable.Translate (seq/codon_tables.go:205-285), markInitCodonAsM=false.
 * Reverse frames complement with the DNA alphabet's PairLetter, leaving
 * letters outside "acgtACGT -.nN" unchanged (:219-226). */
int64_t ora_translate(const uint8_t *seq, size_t len, int table, int frame, int trim, int clean,
                      int allow_unknown, uint8_t *out, int *err) {
    ora_init_tables();
    *err = ORA_OK;
    if (table < 0 || table >= 32 || !codon_have[table]) { *err = ORA_ERR_CODON_TABLE; return 0; }
    if (len < 3) { *err = ORA_ERR_TRANSLATE_SHORT; return 0; }
    if (frame < -3 || frame > 3 || frame == 0) { *err = ORA_ERR_INVALID_FRAME; return 0; }
    int64_t n = 0;
    uint8_t aa, codon[3];
    if (frame < 0) {
        uint8_t pair[256];
        build_pair_lut(1 /* DNA */, pair);
        for (int64_t i = (int64_t)len + frame; i >= 2; i -= 3) {
            codon[0] = pair[seq[i]]; codon[1] = pair[seq[i - 1]]; codon[2] = pair[seq[i - 2]];
            int e = codon_get(table, codon, allow_unknown, &aa);
            if (e) { *err = e; return 0; }
            if (trim && (aa == 'X' || aa == '*')) break;
            if (clean && aa == '*') aa = 'X';
            out[n++] = aa;
        }
    } else {
        for (int64_t i = frame - 1; i < (int64_t)len - 2; i += 3) {
            int e = codon_get(table, seq + i, allow_unknown, &aa);
            if (e) { *err = e; return 0; }
            if (trim && (aa == 'X' || aa == '*')) break;
            if (clean && aa == '*') aa = 'X';
            out[n++] = aa;
        }
    }
    return n;
}

/* ------------------------------------------------------------------ */
/* wyhash (zeebo/wyhash v0.0.1 Hash(b, seed)) -- PARITY UNPINNED        */
/* Published wyhash v1 layout: 32-byte blocks, switch(len&31) tail,     */
/* final mum(seed, len ^ p5).                                          */
/* ------------------------------------------------------------------ */
#define WYP0 0xa0761d6478bd642fULL
#define WYP1 0xe7037ed1a0b428dbULL
#define WYP2 0x8ebc6af09c88c6e3ULL
#define WYP3 0x589965cc75374cc3ULL
#define WYP4 0x1d8e4e27c47d124fULL
#define WYP5 0xeb44accab455d165ULL

static inline uint64_t wymum(uint64_t a, uint64_t b) {
    __uint128_t r = (__uint128_t)a * b;
    return (uint64_t)(r >> 64) ^ (uint64_t)r;
}
static inline uint64_t wyr08(const uint8_t *p) { return p[0]; }
static inline uint64_t wyr16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint64_t wyr32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t wyr64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
/* the tail reads 8 bytes as two 32-bit halves, first half high */
static inline uint64_t wyr64s(const uint8_t *p) { return (wyr32(p) << 32) | wyr32(p + 4); }

uint64_t ora_wyhash(const uint8_t *p, uint64_t len, uint64_t seed) {
    uint64_t i;
    for (i = 0; i + 32 <= len; i += 32, p += 32)
        seed = wymum(seed ^ WYP0, wymum(wyr64(p) ^ WYP1, wyr64(p + 8) ^ WYP2) ^
                                      wymum(wyr64(p + 16) ^ WYP3, wyr64(p + 24) ^ WYP4));
    seed ^= WYP0;
    switch (len & 31) {
    case 0: break;
    case 1: seed = wymum(seed, wyr08(p) ^ WYP1); break;
    case 2: seed = wymum(seed, wyr16(p) ^ WYP1); break;
    case 3: seed = wymum(seed, ((wyr16(p) << 8) | wyr08(p + 2)) ^ WYP1); break;
    case 4: seed = wymum(seed, wyr32(p) ^ WYP1); break;
    case 5: seed = wymum(seed, ((wyr32(p) << 8) | wyr08(p + 4)) ^ WYP1); break;
    case 6: seed = wymum(seed, ((wyr32(p) << 16) | wyr16(p + 4)) ^ WYP1); break;
    case 7: seed = wymum(seed, ((wyr32(p) << 24) | (wyr16(p + 4) << 8) | wyr08(p + 6)) ^ WYP1); break;
    case 8: seed = wymum(seed, wyr64s(p) ^ WYP1); break;
    case 9: seed = wymum(wyr64s(p) ^ seed, wyr08(p + 8) ^ WYP2); break;
    case 10: seed = wymum(wyr64s(p) ^ seed, wyr16(p + 8) ^ WYP2); break;
    case 11: seed = wymum(wyr64s(p) ^ seed, ((wyr16(p + 8) << 8) | wyr08(p + 10)) ^ WYP2); break;
    case 12: seed = wymum(wyr64s(p) ^ seed, wyr32(p + 8) ^ WYP2); break;
    case 13: seed = wymum(wyr64s(p) ^ seed, ((wyr32(p + 8) << 8) | wyr08(p + 12)) ^ WYP2); break;
    case 14: seed = wymum(wyr64s(p) ^ seed, ((wyr32(p + 8) << 16) | wyr16(p + 12)) ^ WYP2); break;
    case 15: seed = wymum(wyr64s(p) ^ seed, ((wyr32(p + 8) << 24) | (wyr16(p + 12) << 8) | wyr08(p + 14)) ^ WYP2); break;
    case 16: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2); break;
    case 17: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(seed, wyr08(p + 16) ^ WYP3); break;
    case 18: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(seed, wyr16(p + 16) ^ WYP3); break;
    case 19: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(seed, ((wyr16(p + 16) << 8) | wyr08(p + 18)) ^ WYP3); break;
    case 20: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(seed, wyr32(p + 16) ^ WYP3); break;
    case 21: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(seed, ((wyr32(p + 16) << 8) | wyr08(p + 20)) ^ WYP3); break;
    case 22: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(seed, ((wyr32(p + 16) << 16) | wyr16(p + 20)) ^ WYP3); break;
    case 23: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(seed, ((wyr32(p + 16) << 24) | (wyr16(p + 20) << 8) | wyr08(p + 22)) ^ WYP3); break;
    case 24: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(seed, wyr64s(p + 16) ^ WYP3); break;
    case 25: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(wyr64s(p + 16) ^ seed, wyr08(p + 24) ^ WYP4); break;
    case 26: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(wyr64s(p + 16) ^ seed, wyr16(p + 24) ^ WYP4); break;
    case 27: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(wyr64s(p + 16) ^ seed, ((wyr16(p + 24) << 8) | wyr08(p + 26)) ^ WYP4); break;
    case 28: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(wyr64s(p + 16) ^ seed, wyr32(p + 24) ^ WYP4); break;
    case 29: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(wyr64s(p + 16) ^ seed, ((wyr32(p + 24) << 8) | wyr08(p + 28)) ^ WYP4); break;
    case 30: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(wyr64s(p + 16) ^ seed, ((wyr32(p + 24) << 16) | wyr16(p + 28)) ^ WYP4); break;
    case 31: seed = wymum(wyr64s(p) ^ seed, wyr64s(p + 8) ^ WYP2) ^ wymum(wyr64s(p + 16) ^ seed, ((wyr32(p + 24) << 24) | (wyr16(p + 28) << 8) | wyr08(p + 30)) ^ WYP4); break;
    }
    return wymum(seed, len ^ WYP5);
}

/* ------------------------------------------------------------------ */
/* NewProteinIterator / Next (sketches/iterator-protein.go:46-90)      */
/* Nucleotide input only (Alphabet != Protein branch, :62-67).          */
/* ------------------------------------------------------------------ */
int64_t ora_protein_iterator(const uint8_t *seq, size_t len, int k, int table, int frame,
                             uint64_t *out, int *err) {
    *err = ORA_OK;
    if (k < 1) { *err = ORA_ERR_INVALID_K; return 0; }                 /* :47 */
    if ((int64_t)len < (int64_t)k * 3) { *err = ORA_ERR_SHORT_SEQ; return 0; } /* :50 */
    uint8_t *aa = (uint8_t *)malloc(len / 3 + 4);
    int64_t na = ora_translate(seq, len, table, frame, 0, 0, 1, aa, err); /* :63 */
    if (*err) { free(aa); return 0; }
    int64_t end = na - k, n = 0;                                          /* :70 */
    for (int64_t idx = 0; idx <= end; idx++)                              /* :81-88 */
        out[n++] = ora_wyhash(aa + idx, (uint64_t)k, 1);
    free(aa);
    return n;
}

/* ------------------------------------------------------------------ */
/* NewProteinMinimizerSketch / Next (sketches/sketch-protein.go:62-210) */
/* The same sorted-buffer state machine as NextMinimizer over the       */
/* wyhash stream of the translated frame.  protein_input != 0: the      */
/* sequence already is amino acids (Alphabet == Protein, :86-87).       */
/* out_idx receives Index() (= mI, :213-215).                           */
/* ------------------------------------------------------------------ */
int64_t ora_protein_minimizer(const uint8_t *seq, size_t len, int k, int table, int frame, int w,
                              int protein_input, int sort_policy, uint64_t *out_val, int64_t *out_idx,
                              int *err, int *first_window_tie) {
    *err = ORA_OK;
    if (first_window_tie) *first_window_tie = 0;
    if (k < 1) { *err = ORA_ERR_INVALID_K; return 0; }                          /* :63 */
    if ((int64_t)len < (int64_t)k * 3) { *err = ORA_ERR_SHORT_SEQ; return 0; }   /* :66 */
    if (w < 1) { *err = ORA_ERR_INVALID_W; return 0; }                          /* :70 */
    if ((int64_t)len < (int64_t)k * 3 + w - 1) { *err = ORA_ERR_SHORT_SEQ; return 0; } /* :73 */
    uint8_t *aa = (uint8_t *)malloc(len + 4);
    int64_t na;
    if (!protein_input) {
        na = ora_translate(seq, len, table, frame, 0, 0, 1, aa, err);           /* :84 */
        if (*err) { free(aa); return 0; }
    } else {
        memcpy(aa, seq, len);
        na = (int64_t)len;
    }
    int64_t idx = 0, end0 = na - k;                                              /* :92-93 */
    int r = w - 1;                                                               /* :97 */
    int skip = (w == 1);                                                         /* :95 */
    iv_t *buf = (iv_t *)malloc(((size_t)w + 2) * sizeof(iv_t));
    int blen = 0;
    int64_t pre_min_idx = -1, n = 0;
    for (;;) {
        if (idx > end0) break;                                                   /* :113 */
        uint64_t code = ora_wyhash(aa + idx, (uint64_t)k, 1);                    /* :118 */
        if (skip) { out_val[n] = code; out_idx[n] = idx; n++; idx++; continue; } /* :120-124 */
        if (idx < r) { buf[blen].idx = idx; buf[blen].val = code; blen++; idx++; continue; } /* :127-132 */
        if (idx == r) {                                                          /* :135-147 */
            buf[blen].idx = idx; buf[blen].val = code; blen++;
            if (first_window_tie && has_equal_values(buf, blen)) *first_window_tie = 1;
            first_window_sort(buf, blen, sort_policy);
            out_val[n] = buf[0].val; out_idx[n] = buf[0].idx; n++;
            pre_min_idx = buf[0].idx;
            idx++;
            continue;
        }
        evict_idx(buf, r, idx - w);                                              /* :152-160 */
        insert_sorted(buf, r, idx, code);                                        /* :163-197 */
        if (buf[0].idx == pre_min_idx) { idx++; continue; }                      /* :199-203 */
        out_val[n] = buf[0].val; out_idx[n] = buf[0].idx; n++;                   /* :205-209 */
        pre_min_idx = buf[0].idx;
        idx++;
    }
    free(buf); free(aa);
    return n;
}

/* ------------------------------------------------------------------ */
/* Batch drivers (concatenated reads + offsets), multi-threaded.        */
/* These are what bench.py times as the CPU baseline: the reference's   */
/* per-record pull loop, one thread per contiguous shard of reads.      */
/* mode: 0 kmer, 1 nthash, 2 minimizer, 3 syncmer, 4 protein,           */
/*       5 protein minimizer (alphabet 5 = amino-acid input), 6 simhash */
/* Pass 1 (out_val == NULL) only counts; pass 2 writes at out_off[r].   */
/* ------------------------------------------------------------------ */
typedef struct {
    int mode, k, w, s, canonical, circular, codon_table, frame, alphabet, sort_policy, m, scale;
} ora_params;

typedef struct {
    const ora_params *p;
    const uint8_t *bases;
    const uint64_t *off;
    uint64_t r0, r1;
    uint64_t *counts;       /* per read */
    int32_t *status;        /* per read */
    const uint64_t *out_off;/* per read (pass 2) */
    uint64_t *out_val;
    uint32_t *out_pos;
    uint64_t ties;          /* reads with first-window ties */
    uint64_t sum;           /* checksum so the work cannot be optimised away */
} ora_job;

static int64_t ora_max_out(const ora_params *p, size_t len) {
    size_t ext = len + (p->circular ? (size_t)(p->k > 0 ? p->k - 1 : 0) : 0);
    if (p->mode == 0 && !p->canonical) return 2 * (int64_t)ext + 2;
    return (int64_t)ext + 2;
}

static void *ora_worker(void *arg) {
    ora_job *j = (ora_job *)arg;
    const ora_params *p = j->p;
    size_t cap = 1024;
    uint64_t sum = 0, ties = 0; /* thread-local: the job structs share cache lines */
    uint64_t *val = (uint64_t *)malloc(cap * sizeof(uint64_t));
    int64_t *idx = (int64_t *)malloc(cap * sizeof(int64_t));
    for (uint64_t r = j->r0; r < j->r1; r++) {
        const uint8_t *s = j->bases + j->off[r];
        size_t len = (size_t)(j->off[r + 1] - j->off[r]);
        int64_t need = ora_max_out(p, len);
        if ((size_t)need > cap) {
            cap = (size_t)need * 2;
            val = (uint64_t *)realloc(val, cap * sizeof(uint64_t));
            idx = (int64_t *)realloc(idx, cap * sizeof(int64_t));
        }
        int err = 0, tie = 0;
        int64_t n = 0, eidx;
        int have_idx = 0;
        switch (p->mode) {
        case 0: n = ora_kmer_iterator(s, len, p->k, p->canonical, p->circular, p->alphabet, val, &err, &eidx); break;
        case 1: n = ora_hash_iterator(s, len, p->k, p->canonical, p->circular, val, &err); break;
        case 2: n = ora_minimizer(s, len, p->k, p->w, p->circular, p->sort_policy, val, idx, &err, &tie); have_idx = 1; break;
        case 3: n = ora_syncmer(s, len, p->k, p->s, p->circular, p->sort_policy, val, idx, &err, &tie); have_idx = 1; break;
        case 4: n = ora_protein_iterator(s, len, p->k, p->codon_table, p->frame, val, &err); break;
        case 6: n = ora_simhash_iterator(s, len, p->k, p->m, p->scale, p->canonical, p->circular, val, &err); break;
        case 5: n = ora_protein_minimizer(s, len, p->k, p->codon_table, p->frame, p->w, p->alphabet == 5,
                                          p->sort_policy, val, idx, &err, &tie); have_idx = 1; break;
        default: err = ORA_ERR_INVALID_K;
        }
        ties += (uint64_t)tie;
        if (j->status) j->status[r] = err;
        if (j->counts) j->counts[r] = (uint64_t)n;
        for (int64_t i = 0; i < n; i++) sum += val[i];
        if (j->out_val) {
            uint64_t o = j->out_off[r];
            memcpy(j->out_val + o, val, (size_t)n * sizeof(uint64_t));
            if (j->out_pos) {
                for (int64_t i = 0; i < n; i++) {
                    int64_t pos = have_idx ? idx[i] : i;
                    if (p->mode == 0 && !p->canonical) {
                        /* Index() restarts at 0 on the reverse strand */
                        int64_t per = (int64_t)len + (p->circular ? p->k - 1 : 0) - p->k + 1;
                        if (i >= per) pos = i - per;
                    }
                    j->out_pos[o + (uint64_t)i] = (uint32_t)pos;
                }
            }
        }
    }
    free(val); free(idx);
    j->sum = sum; j->ties = ties;
    return 0;
}

/* Returns checksum; fills counts/status (may be NULL); when out_val != NULL
 * writes values (and positions if out_pos != NULL) at out_off[r]. */
uint64_t ora_run_batch(const ora_params *p, const uint8_t *bases, const uint64_t *off, uint64_t n_reads,
                       int n_threads, uint64_t *counts, int32_t *status, const uint64_t *out_off,
                       uint64_t *out_val, uint32_t *out_pos, uint64_t *ties_out) {
    ora_init_tables();
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_reads && n_reads > 0) n_threads = (int)n_reads;
    ora_job *jobs = (ora_job *)calloc((size_t)n_threads, sizeof(ora_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        jobs[t].p = p; jobs[t].bases = bases; jobs[t].off = off;
        jobs[t].r0 = n_reads * (uint64_t)t / (uint64_t)n_threads;
        jobs[t].r1 = n_reads * (uint64_t)(t + 1) / (uint64_t)n_threads;
        jobs[t].counts = counts; jobs[t].status = status;
        jobs[t].out_off = out_off; jobs[t].out_val = out_val; jobs[t].out_pos = out_pos;
        if (n_threads == 1) ora_worker(&jobs[t]);
        else pthread_create(&th[t], 0, ora_worker, &jobs[t]);
    }
    uint64_t sum = 0, ties = 0;
    for (int t = 0; t < n_threads; t++) {
        if (n_threads > 1) pthread_join(th[t], 0);
        sum += jobs[t].sum; ties += jobs[t].ties;
    }
    if (ties_out) *ties_out = ties;
    free(jobs); free(th);
    return sum;
}
