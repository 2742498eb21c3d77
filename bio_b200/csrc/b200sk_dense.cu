// Dense modes: every position of a read emits exactly one element, so the output offsets follow from
// the read lengths alone (k_scan_reads<true> writes out_off[] and read_status[] before this kernel runs).
//
//   MODE_NTHASH   NextHash        sketches/iterator.go:658-665   (also minimizer w==1 / syncmer s==k)
//   MODE_KMER     NextKmer        sketches/iterator.go:708-759   2-bit codes, canonical or both strands
//   MODE_PROTEIN  ProteinIterator sketches/iterator-protein.go:46-90  translate one frame + wyhash(seed 1)
//
// Same tiling as the sparse kernels (tile of blockDim.x items, one TMA bulk copy per tile, one thread walks
// one item).  Each thread produces 8 bytes per step; a warp collects 16 steps per lane in shared memory and
// then writes every lane's 128-byte run with coalesced stores.
#include "../../include/b200sk_codon_data.h"
#include "b200sk_tile.cuh"
#include "b200sk_protein.cuh"

namespace b200sk {

// steps staged per lane between flushes: 16 (128-byte runs) except both-strand k-mers, which keep two
// staging areas per warp and gain more from the occupancy of 8-step rows.  Must match dense_steps() in
// b200sk_api.cu.
#define DENSE_S (MODE == B200SK_MODE_KMER ? 8 : 16)
#define DENSE_ROW (DENSE_S * 8 + 8)         /* bytes per lane row (+8: rows land on different banks) */
#define DENSE_WARP_STAGE (32 * DENSE_ROW + 512) /* one staging area per warp (two for both-strand k-mers): 32 rows +
                                                   32 x 16 B of per-lane flush descriptors */

// ------------------------------------------------------------------ 2-bit base code, pair letters
// sketches/kmers.go:23-40 (IUPAC codes map to their first base; 4 = illegal)
__device__ __host__ inline uint32_t base2bit_of(uint32_t b) {
    switch (b) {
    case 'A': case 'a': case 'D': case 'd': case 'H': case 'h': case 'M': case 'm':
    case 'N': case 'n': case 'R': case 'r': case 'V': case 'v': case 'W': case 'w': return 0;
    case 'B': case 'b': case 'C': case 'c': case 'S': case 's': case 'Y': case 'y': return 1;
    case 'G': case 'g': case 'K': case 'k': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
    }
}
// seq.Alphabet.PairLetter (seq/alphabet.go:313-325, letters/pairs :353-399); letters outside the
// alphabet come back unchanged (seq/seq.go:389-391)
__device__ __host__ inline uint32_t pair_letter(int alphabet, uint32_t b) {
    const char *l, *p;
    switch (alphabet) {
    case B200SK_ALPHABET_DNA_REDUNDANT: l = "acgtryswkmbdhvACGTRYSWKMBDHV"; p = "tgcayrswmkvhdbTGCAYRSWMKVHDB"; break;
    case B200SK_ALPHABET_DNA: l = "acgtACGT"; p = "tgcaTGCA"; break;
    case B200SK_ALPHABET_RNA_REDUNDANT: l = "acguryswkmbdhvACGURYSWKMBDHV"; p = "ugcayrswmkvhdbUGCAYRSWMKVHDB"; break;
    case B200SK_ALPHABET_RNA: l = "acguACGU"; p = "ugcaUGCA"; break;
    default: return b;
    }
    for (int i = 0; l[i]; i++)
        if ((uint32_t)(uint8_t)l[i] == b) return (uint32_t)(uint8_t)p[i];
    return b;
}

// first illegal base of every read (0xffffffff: none) -- one warp per read
__global__ void k_first_illegal(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ off,
                                uint64_t n_reads, uint32_t *ill) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n_reads; r += nwarps) {
        const uint64_t s0 = off[r], L = off[r + 1] - s0;
        uint32_t best = 0xffffffffu;
        for (uint64_t i0 = 0; i0 < L && best == 0xffffffffu; i0 += 32) {
            const uint64_t i = i0 + lane;
            const bool bad = i < L && base2bit_of(bases[s0 + i]) == 4;
            const unsigned m = __ballot_sync(0xffffffffu, bad);
            if (m) best = (uint32_t)(i0 + (uint64_t)(__ffs(m) - 1));
        }
        if (lane == 0) ill[r] = best;
    }
}

// ------------------------------------------------------------------ translation of one frame of every read
// seq.Seq.Translate(table, frame, trim=false, clean=false, allowUnknownCodon=true, markInitCodonAsM=false)
// (sketch-protein.go:84, codon_tables.go:205-285): aa[aa_off[r] + t], eight lanes per read.
__global__ void __launch_bounds__(256) k_translate(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ off,
                                                   const uint64_t *__restrict__ aa_off, uint64_t n_reads, int frame,
                                                   const uint8_t *__restrict__ aux, uint8_t *aa) {
    __shared__ uint8_t tab[4608];
    for (uint32_t i = threadIdx.x; i < 4608; i += blockDim.x) tab[i] = aux[i];
    __syncthreads();
    // eight lanes per read (four reads per warp): a 150-bp frame is 50 residues, 32 lanes per read would idle
    const unsigned sub = threadIdx.x & 7u;
    const uint64_t grp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 3;
    const uint64_t ngrp = ((uint64_t)gridDim.x * blockDim.x) >> 3;
    const uint8_t *pl = tab + 4352;
    for (uint64_t r = grp; r < n_reads; r += ngrp) {
        const uint8_t *s = bases + off[r];
        const uint64_t L = off[r + 1] - off[r];
        const uint64_t o = aa_off[r], n = aa_off[r + 1] - o;
        if (frame > 0) {
            for (uint64_t t = sub; t < n; t += 8) {
                const uint64_t i = (uint64_t)(frame - 1) + 3 * t;
                aa[o + t] = (uint8_t)codon_aa(tab, s[i], s[i + 1], s[i + 2]);
            }
        } else {
            for (uint64_t t = sub; t < n; t += 8) {
                const uint64_t i = L - (uint64_t)(-frame) - 3 * t;
                aa[o + t] = (uint8_t)codon_aa(tab, pl[s[i]], pl[s[i - 1]], pl[s[i - 2]]);
            }
        }
    }
}

cudaError_t launch_translate(const uint8_t *bases, const uint64_t *off, const uint64_t *aa_off, uint64_t n_reads,
                             int frame, const uint8_t *aux, uint8_t *aa, cudaStream_t st) {
    uint64_t cb = (n_reads + 31) / 32;
    if (cb > 148 * 16) cb = 148 * 16;
    if (cb == 0) cb = 1;
    k_translate<<<(unsigned)cb, 256, 0, st>>>(bases, off, aa_off, n_reads, frame, aux, aa);
    return cudaGetLastError();
}

// N consecutive bytes starting at an arbitrary shared-memory address, fetched as aligned 32-bit words
template <int N> struct ByteWords {
    static constexpr int NWORD = (N + 3 + 3) / 4;
    static constexpr int NG = (N + 3) / 4;
    uint32_t x[NG];
    __device__ __forceinline__ void load(const uint8_t *p) {
        const uint32_t phi = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
        const uint32_t *a = reinterpret_cast<const uint32_t *>(p - phi);
        const uint32_t sel = 0x3210u + 0x1111u * phi;
        uint32_t w[NWORD + 1];
#pragma unroll
        for (int i = 0; i < NWORD; i++) w[i] = a[i];
        w[NWORD] = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) x[g] = __byte_perm(w[g], w[g + 1 < NWORD ? g + 1 : NWORD], sel);
    }
    __device__ __forceinline__ uint32_t byte(const int j) const { return __byte_perm(x[j >> 2], 0u, 0x4440u | (j & 3)); }
};

// ------------------------------------------------------------------ staged, coalesced output
// Each lane filled `n` (<= 16) entries of its row; entry e goes to out[b + d*e] (d = +1, or -1 for the
// second strand of both-strand k-mers), position p0 + d*e.  Two lanes' rows leave per iteration.
template <int MODE>
__device__ __forceinline__ void flush_rows(uint8_t *stage, uint64_t *out_val, void *out_pos, uint32_t pw,
                                           uint64_t b, uint32_t n, int d, uint32_t p0, unsigned lane) {
    // every lane publishes where its row goes; the copy loop then reads the two rows' descriptors with one
    // broadcast LDS.128 each (cheaper on the shared-memory pipe than four shuffles per iteration)
    uint4 *info = reinterpret_cast<uint4 *>(stage + 32 * DENSE_ROW);
    info[lane] = make_uint4((uint32_t)b, (uint32_t)(b >> 32), n | (d < 0 ? 0x80000000u : 0u), p0);
    __syncwarp();
    constexpr int ROWS = 32 / DENSE_S; // rows that leave per iteration, DENSE_S lanes each
#pragma unroll 4
    for (int j = 0; j < 32; j += ROWS) {
        const int src = j + (int)(lane / DENSE_S);
        const uint32_t e = lane % DENSE_S;
        const uint4 inf = info[src];
        const uint32_t nn = inf.z & 0x7fffffffu;
        if (e < nn) {
            const uint64_t v = *reinterpret_cast<const uint64_t *>(stage + src * DENSE_ROW + e * 8);
            const uint64_t bb = ((uint64_t)inf.y << 32) | inf.x;
            const bool neg = (inf.z >> 31) != 0;
            const uint64_t idx = neg ? bb - e : bb + e;
            out_val[idx] = v;
            if (out_pos) store_pos(out_pos, pw, idx, neg ? inf.w - e : inf.w + e);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------ kernel
struct DItem {
    uint64_t r, gb0, obase; // obase: global element index of the item's first element
    uint32_t nb, nstep, p0, np;
    uint64_t L;
    bool valid, both; // both: k-mer second strand is emitted too
};

// item index -> (read, chunk) -> what the item computes and where its elements go
template <int MODE>
__device__ __forceinline__ void dense_item_geometry(const KArgs &a, const ReadGeom &g, uint64_t item_idx,
                                                    uint64_t n_items, DItem &it) {
    const int k = a.k;
    it.valid = item_idx < n_items;
    it.r = 0; it.gb0 = 0; it.obase = 0; it.nb = 0; it.nstep = 0; it.p0 = 0; it.np = 0; it.L = 0; it.both = false;
    if (!it.valid) return;
    {
        const uint64_t item = item_idx;
            uint64_t r = item;
            uint32_t c = 0;
            if (a.item_first) {
                uint64_t lo = 0, hi = a.n_reads;
                if (a.tile_read) { lo = a.tile_read[item >> 5]; hi = lo + 32 < a.n_reads ? lo + 32 : a.n_reads; }
                while (hi - lo > 1) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if (a.item_first[mid] <= item) lo = mid; else hi = mid;
                }
                r = lo;
                c = (uint32_t)(item - a.item_first[r]);
            }
            it.r = r;
            const uint64_t o0 = a.off[r];
            it.L = a.off[r + 1] - o0;
            const uint64_t orig = a.off_orig ? a.off_orig[r + 1] - a.off_orig[r] : it.L;
            int32_t st;
            it.np = read_positions(g, r, it.L, orig, &st);
            it.gb0 = o0;
            if (it.np) {
                // reverse frames read the sequence downwards: hand the chunks out last-first so that item
                // order is still address order and a tile stays one compact byte range
                if (MODE == B200SK_MODE_PROTEIN && a.frame < 0 && !g.protein_input) c = chunks_of(it.np, a.C) - 1 - c;
                it.p0 = c * a.C;
                it.nstep = min(it.np, it.p0 + a.C) - it.p0;
                it.obase = a.out_off[r] - a.out_base + it.p0;
                if (MODE == B200SK_MODE_PROTEIN) {
                    const uint32_t naa = it.nstep + (uint32_t)k - 1; // amino acids the item reads
                    if (g.protein_input) {
                        it.gb0 = o0 + it.p0;
                        it.nb = naa;
                    } else if (a.frame > 0) {
                        it.gb0 = o0 + (uint32_t)(a.frame - 1) + 3ull * it.p0;
                        it.nb = 3 * naa;
                    } else { // reverse frames walk down from base L-|f|
                        const uint64_t top = it.L - (uint64_t)(-a.frame) - 3ull * it.p0; // first codon's base i
                        it.nb = 3 * naa;
                        it.gb0 = o0 + top + 1 - it.nb;
                    }
                } else {
                    it.nb = it.nstep + (uint32_t)k - 1;
                    it.gb0 = o0 + it.p0;
                    it.both = MODE == B200SK_MODE_KMER && !a.canonical && st == B200SK_OK;
                }
            }
    }
}

// Bit-sliced counters for SimHash: plane p holds bit p of the 64 per-bit-position counts.
template <int NB> struct BitCounters {
    uint64_t pl[NB];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int p = 0; p < NB; p++) pl[p] = 0;
    }
    __device__ __forceinline__ void add(uint64_t v) { // count[b] += bit b of v
#pragma unroll
        for (int p = 0; p < NB; p++) { const uint64_t t = pl[p] & v; pl[p] ^= v; v = t; }
    }
    __device__ __forceinline__ void sub(uint64_t v) { // count[b] -= bit b of v
#pragma unroll
        for (int p = 0; p < NB; p++) { const uint64_t t = ~pl[p] & v; pl[p] ^= v; v = t; }
    }
    __device__ __forceinline__ uint64_t ge(uint32_t thr) const { // bit b = (count[b] >= thr)
        uint64_t gt = 0, eq = ~0ull;
#pragma unroll
        for (int p = NB - 1; p >= 0; p--) {
            const uint64_t tb = (thr >> p) & 1u ? ~0ull : 0ull;
            gt |= eq & pl[p] & ~tb;
            eq &= ~(pl[p] ^ tb);
        }
        return gt | eq;
    }
};

template <int MODE, int NB = 1>
__global__ void __launch_bounds__(128) k_dense(const KArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, T = blockDim.x, lane = tid & 31u, wid = tid >> 5;
    uint8_t *tab = smem; // 8 KB: ntHash tables by byte / k-mer LUT / codon tables
    TileCtl *ctl = reinterpret_cast<TileCtl *>(smem + 8192);
    uint8_t *tilebuf = smem + a.sm_tile;
    uint8_t *stage = smem + a.sm_ring + wid * (MODE == B200SK_MODE_KMER ? 2 : 1) * DENSE_WARP_STAGE;
    uint8_t *row = stage + lane * DENSE_ROW;
    const int k = a.k;
    if (MODE == B200SK_MODE_NTHASH || MODE == B200SK_MODE_SIMHASH) {
        const int hk = MODE == B200SK_MODE_SIMHASH ? a.w : k; // SimHash hashes m-mers (a.w carries m)
        ulonglong2 *tIn = reinterpret_cast<ulonglong2 *>(tab), *tOut = tIn + 256;
        for (uint32_t b = tid; b < 256; b += T) {
            const uint64_t f = fwd_seed(b), r = rev_seed(b);
            tIn[b] = make_ulonglong2(f, rol64(r, (unsigned)(hk - 1)));
            tOut[b] = make_ulonglong2(rol64(f, (unsigned)hk), ror64(r, 1));
        }
        if (MODE == B200SK_MODE_NTHASH && tid < 20) { // pair tables of the all-ACGT fast path
            const char letter[4] = {'A', 'C', 'T', 'G'}; // class = (byte >> 1) & 3
            const uint32_t ci = tid & 3u, co = tid >> 2;
            uint64_t x = fwd_seed((uint32_t)letter[ci]), y = rol64(rev_seed((uint32_t)letter[ci]), (unsigned)(k - 1));
            if (co < 4) {
                x ^= rol64(fwd_seed((uint32_t)letter[co]), (unsigned)k);
                y ^= ror64(rev_seed((uint32_t)letter[co]), 1);
            }
            reinterpret_cast<uint64_t *>(smem + 8192 + 256)[tid] = x;
            reinterpret_cast<uint64_t *>(smem + 8192 + 256 + 160)[tid] = y;
        }
    } else if (MODE == B200SK_MODE_KMER) {
        for (uint32_t b = tid; b < 256; b += T) {
            const uint32_t bit = base2bit_of(b);
            const uint32_t cb = base2bit_of(pair_letter(a.alphabet, b));
            tab[b] = (uint8_t)((bit & 3u) | ((cb & 3u) << 2) | (bit == 4 ? 0x10u : 0u));
        }
    } else {
        for (uint32_t i = tid; i < 4608; i += T) tab[i] = a.aux[i];
    }
    if (tid == 0) {
        mbar_init(&ctl->mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint64_t n_items = a.n_items_dev ? *a.n_items_dev : a.n_items;
    const uint64_t total = a.out_off[a.n_reads] - a.out_base;
    const bool fits = total <= a.capacity;
    if (!fits) {
        if (blockIdx.x == 0 && tid == 0) atomicOr(a.flags, B200SK_FLAG_CAPACITY);
        return;
    }
    const ReadGeom g = a.geom();
    uint32_t parity = 0;
    for (;;) {
        if (tid == 0) ctl->tile = atomicAdd(a.ticket, 1ULL);
        __syncthreads();
        const uint64_t tile = ctl->tile;
        const uint64_t item0 = tile * T;
        if (item0 >= n_items) break;
        // ---- geometry
        DItem it;
        dense_item_geometry<MODE>(a, g, item0 + tid, n_items, it);
        // tile byte range = [min gb0, max gb0+nb) over the items (reverse frames walk a read downwards, so
        // item order is not address order here)
        if (tid == 0) { ctl->lo = ~0ULL; ctl->hi = 0; }
        __syncthreads();
        {
            unsigned long long lo = it.nb ? it.gb0 : ~0ULL, hi = it.nb ? it.gb0 + it.nb : 0ULL;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if (lane == 0) {
                atomicMin(reinterpret_cast<unsigned long long *>(&ctl->lo), lo);
                atomicMax(reinterpret_cast<unsigned long long *>(&ctl->hi), hi);
            }
        }
        __syncthreads();
        const uint64_t lo_al = ctl->hi ? ctl->lo & ~15ULL : 0;
        const uint64_t span = ctl->hi > lo_al ? ctl->hi - lo_al : 0;
        const uint32_t bytes = (uint32_t)((span + 15ULL) & ~15ULL);
        const bool span_ok = bytes <= a.sm_tile_bytes;
        if (tid == 0 && bytes && span_ok) {
            mbar_expect_tx(&ctl->mbar, bytes);
            tma_load_1d(tilebuf, a.bases + lo_al, bytes, &ctl->mbar);
        }
        if (!span_ok && tid == 0) atomicOr(a.flags, B200SK_FLAG_SPAN);
        if (bytes && span_ok) {
            mbar_wait(&ctl->mbar, parity);
            parity ^= 1u;
        }
        // ntHash fast path: when every byte of the tile is one of ACGTacgt the bytes are rewritten as 2-bit
        // classes (a=0 c=1 t=2 g=3) and each step needs ONE pair-table lookup per strand instead of two
        // 16-byte lookups.  Any other byte: the tile is fetched again and walks the general 256-entry tables.
        bool fast = false;
        if (MODE == B200SK_MODE_NTHASH && bytes && span_ok) {
            uint32_t bad = 0;
            for (uint32_t o = tid * 16u; o < bytes; o += T * 16u) {
                uint4 v = *reinterpret_cast<uint4 *>(tilebuf + o);
                uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t x = w[i] | 0x20202020u;             // lower case
                    const uint32_t t = (x >> 1) & 0x03030303u;         // class of every byte
                    const uint32_t u2 = t | (t >> 4);                  // nibble pairs in bytes 0 and 2
                    const uint32_t sel = __byte_perm(u2, 0u, 0x4420u); // four nibbles = PRMT selector
                    const uint32_t want = __byte_perm(0x67746361u, 0u, sel); // 'a','c','t','g' by class
                    bad |= x ^ want;
                    w[i] = t;
                }
                *reinterpret_cast<uint4 *>(tilebuf + o) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            // bytes outside [lo, hi) (alignment slop) may be anything: only the tile's own range decides.
            // Slop bytes belong to neighbouring reads or padding; a false alarm only costs the slow path.
            fast = __syncthreads_or(bad) == 0;
            if (!fast) {
                if (tid == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&ctl->mbar, bytes);
                    tma_load_1d(tilebuf, a.bases + lo_al, bytes, &ctl->mbar);
                }
                mbar_wait(&ctl->mbar, parity);
                parity ^= 1u;
            }
        }
        const uint32_t nstep = span_ok ? it.nstep : 0u;
        const uint8_t *sb = tilebuf + (uint32_t)(it.gb0 - lo_al);
        // warp-uniform trip count
        uint32_t maxn = nstep;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxn = max(maxn, __shfl_xor_sync(0xffffffffu, maxn, o));

        if (MODE == B200SK_MODE_NTHASH) {
            const ulonglong2 *tIn = reinterpret_cast<const ulonglong2 *>(tab), *tOut = tIn + 256;
            const bool canonical = a.canonical != 0;
            uint64_t fh = 0, rh = 0;
            if (nstep && !fast)
                for (int j = 0; j < k - 1; j++) {
                    const ulonglong2 e = tIn[sb[j]];
                    fh = rol1(fh) ^ e.x;
                    rh = ror1(rh) ^ e.y;
                }
            const uint8_t *sbs = nstep ? sb : tilebuf + 4;                          // a safe base for idle lanes
            if (fast) {
                // pair tables: X[cin + 4 cout] = A[in] ^ rol(A[out], k), Y = rol(B[in], k-1) ^ ror(B[out], 1);
                // cout = 4: no outgoing base (the folds of the first k-mer)
                const uint8_t *tX = smem + 8192 + 256, *tY = tX + 160; // 20 x 8 B each, after the control block
                if (nstep)
                    for (int j = 0; j < k - 1; j++) {
                        const uint32_t o8 = (uint32_t)sb[j] * 8u + 128u;
                        fh = rol1(fh) ^ *reinterpret_cast<const uint64_t *>(tX + o8);
                        rh = ror1(rh) ^ *reinterpret_cast<const uint64_t *>(tY + o8);
                    }
                const uint32_t last_block = nstep ? ((nstep - 1) / DENSE_S) * DENSE_S : 0u;
                for (uint32_t u0 = 0; u0 < maxn; u0 += DENSE_S) {
                    ByteWords<DENSE_S> win, wout;
                    const uint32_t ul = min(u0, last_block);
                    win.load(sbs + ul + k - 1);
                    wout.load(sbs + ul - 1);
#pragma unroll
                    for (uint32_t e = 0; e < DENSE_S; e++) {
                        const uint32_t u = u0 + e;
                        if (u < nstep) {
                            const uint32_t co = (e == 0 && u0 == 0) ? 4u : wout.byte(e);
                            const uint32_t o8 = win.byte(e) * 8u + co * 32u;
                            fh = rol1(fh) ^ *reinterpret_cast<const uint64_t *>(tX + o8);
                            rh = ror1(rh) ^ *reinterpret_cast<const uint64_t *>(tY + o8);
                            *reinterpret_cast<uint64_t *>(row + e * 8) = (canonical && rh < fh) ? rh : fh;
                        }
                    }
                    const uint32_t n = u0 < nstep ? min((uint32_t)DENSE_S, nstep - u0) : 0u;
                    flush_rows<MODE>(stage, a.out_val, a.out_pos, a.pos_width, it.obase + u0, n, 1, it.p0 + u0, lane);
                }
                __syncthreads();
                continue;
            }
            const uint32_t last_block = nstep ? ((nstep - 1) / DENSE_S) * DENSE_S : 0u; // first step of the last block
            // the 16 incoming and 16 outgoing bases of a staging block are fetched as aligned words
            // (ByteWords), a quarter of the shared-memory wavefronts of byte loads
            for (uint32_t u0 = 0; u0 < maxn; u0 += DENSE_S) {
                // lanes whose item is already finished (or absent) re-read their own last block: the loads stay
                // unconditional (schedulable) and inside the tile
                ByteWords<DENSE_S> win, wout;
                const uint32_t ul = min(u0, last_block);
                win.load(sbs + ul + k - 1);
                wout.load(sbs + ul - 1);
#pragma unroll
                for (uint32_t e = 0; e < DENSE_S; e++) {
                    const uint32_t u = u0 + e;
                    if (u < nstep) {
                        const ulonglong2 in = tIn[win.byte(e)];
                        ulonglong2 o = make_ulonglong2(0, 0);
                        if (u) o = tOut[wout.byte(e)];
                        fh = rol1(fh) ^ o.x ^ in.x;
                        rh = ror1(rh) ^ o.y ^ in.y;
                        *reinterpret_cast<uint64_t *>(row + e * 8) = (canonical && rh < fh) ? rh : fh; // iterator.go:659
                    }
                }
                const uint32_t n = u0 < nstep ? min((uint32_t)DENSE_S, nstep - u0) : 0u;
                flush_rows<MODE>(stage, a.out_val, a.out_pos, a.pos_width, it.obase + u0, n, 1, it.p0 + u0, lane);
            }
        } else if (MODE == B200SK_MODE_KMER) {
            // iterator.go:736,740,754: code = (pre & mask1) << 2 | bit; rc = (bit ^ 3) << 2(k-1) | preRC >> 2.
            // Second strand (non-canonical, :713-723): the codes of RevComInplace(seq), i.e. for k-mer i the
            // pair-letter reverse complement, emitted at index np-1-i after the np forward codes.
            const int sh = 2 * (k - 1);
            const uint64_t mask1 = (1ull << sh) - 1ull; // iterator.go:699 (sh <= 62)
            uint64_t code = 0, rc3 = 0, rcq = 0;
            const bool canonical = a.canonical != 0;
            uint8_t *row2 = row + DENSE_WARP_STAGE;
            if (nstep)
                for (int j = 0; j < k - 1; j++) {
                    const uint32_t v = tab[sb[j]];
                    code = (code << 2) | (v & 3u);
                    rc3 = (rc3 >> 2) | ((uint64_t)((v & 3u) ^ 3u) << sh);
                    rcq = (rcq >> 2) | ((uint64_t)((v >> 2) & 3u) << sh);
                }
            for (uint32_t u0 = 0; u0 < maxn; u0 += DENSE_S) {
#pragma unroll 4
                for (uint32_t e = 0; e < DENSE_S; e++) {
                    const uint32_t u = u0 + e;
                    if (u < nstep) {
                        const uint32_t v = tab[sb[u + k - 1]];
                        code = ((code & mask1) << 2) | (v & 3u);
                        rc3 = (rc3 >> 2) | ((uint64_t)((v & 3u) ^ 3u) << sh);
                        rcq = (rcq >> 2) | ((uint64_t)((v >> 2) & 3u) << sh);
                        *reinterpret_cast<uint64_t *>(row + e * 8) = (canonical && rc3 < code) ? rc3 : code;
                        *reinterpret_cast<uint64_t *>(row2 + e * 8) = rcq;
                    }
                }
                const uint32_t n = u0 < nstep ? min((uint32_t)DENSE_S, nstep - u0) : 0u;
                flush_rows<MODE>(stage, a.out_val, a.out_pos, a.pos_width, it.obase + u0, n, 1, it.p0 + u0, lane);
                // strand 2: k-mer i = p0+u0+e lands at out_off[r] + np + (np-1-i), position np-1-i
                const uint32_t i0 = it.p0 + u0;
                const uint64_t b2 = it.obase - it.p0 + 2ull * it.np - 1 - i0;
                flush_rows<MODE>(stage + DENSE_WARP_STAGE, a.out_val, a.out_pos, a.pos_width, b2, it.both ? n : 0u, -1,
                           it.np - 1 - i0, lane);
            }
        } else if (MODE == B200SK_MODE_SIMHASH) {
            // NextSimHash (iterator.go:191-612): per k-mer, the per-bit majority over its n = k-m+1 m-mer
            // hashes that pass the FracMinHash filter.  The window slides by one m-mer per k-mer: subtract the
            // hash that leaves (ring in shared memory), add the one that enters, compare every bit count with
            // the threshold (nPos+1)/2 -- all 64 bit positions at once in bit-sliced counters.
            const int m = a.w, n = k - m + 1;
            const ulonglong2 *tIn = reinterpret_cast<const ulonglong2 *>(tab), *tOut = tIn + 256;
            const bool canonical = a.canonical != 0;
            const uint64_t maxh = a.s > 1 ? 0xffffffffffffffffull / (uint64_t)a.s : 0xffffffffffffffffull; // :180-185
            uint64_t *ring = reinterpret_cast<uint64_t *>(smem + a.sm_listv) + tid; // [n][T]
            BitCounters<NB> cnt;
            cnt.clear();
            uint32_t npos = 0, slot = 0;
            uint64_t fh = 0, rh = 0;
            if (nstep) {
                for (int j = 0; j < m - 1; j++) {
                    const ulonglong2 e = tIn[sb[j]];
                    fh = rol1(fh) ^ e.x;
                    rh = ror1(rh) ^ e.y;
                }
                // the first n-1 m-mers only fill the window
                for (int t = 0; t < n - 1; t++) {
                    const ulonglong2 in = tIn[sb[t + m - 1]];
                    ulonglong2 o = make_ulonglong2(0, 0);
                    if (t) o = tOut[sb[t - 1]];
                    fh = rol1(fh) ^ o.x ^ in.x;
                    rh = ror1(rh) ^ o.y ^ in.y;
                    uint64_t hv = (canonical && rh < fh) ? rh : fh;
                    if (hv > maxh) hv = 0;
                    ring[(uint32_t)t * T] = hv;
                    cnt.add(hv);
                    npos += hv != 0;
                }
                slot = (uint32_t)(n - 1);
            }
            for (uint32_t u0 = 0; u0 < maxn; u0 += DENSE_S) {
                for (uint32_t e = 0; e < DENSE_S; e++) {
                    const uint32_t u = u0 + e;
                    if (u < nstep) {
                        const uint32_t t = u + (uint32_t)n - 1; // m-mer entering the window of k-mer u
                        const ulonglong2 in = tIn[sb[t + m - 1]];
                        ulonglong2 o = make_ulonglong2(0, 0);
                        if (t) o = tOut[sb[t - 1]];
                        fh = rol1(fh) ^ o.x ^ in.x;
                        rh = ror1(rh) ^ o.y ^ in.y;
                        uint64_t hv = (canonical && rh < fh) ? rh : fh;
                        if (hv > maxh) hv = 0;                          // iterator.go:281
                        const uint64_t old = u ? ring[slot * T] : 0ull; // the m-mer of k-mer u-1 that left (:205)
                        ring[slot * T] = hv;
                        slot = slot + 1 == (uint32_t)n ? 0 : slot + 1;
                        cnt.sub(old);
                        npos -= old != 0;
                        cnt.add(hv);
                        npos += hv != 0;
                        *reinterpret_cast<uint64_t *>(row + e * 8) = npos ? cnt.ge((npos + 1) >> 1) : 0ull; // :357-424
                    }
                }
                const uint32_t nn = u0 < nstep ? min((uint32_t)DENSE_S, nstep - u0) : 0u;
                flush_rows<MODE>(stage, a.out_val, a.out_pos, a.pos_width, it.obase + u0, nn, 1, it.p0 + u0, lane);
            }
        } else { // PROTEIN
            uint8_t *aab = smem + a.sm_listv + tid * a.lcap; // lcap = per-thread amino-acid buffer stride
            const uint32_t naa = nstep ? nstep + (uint32_t)k - 1 : 0u;
            ByteSrc src;
            if (g.protein_input) {
                src.p = sb;
            } else {
                src.p = aab;
                if (a.frame > 0) {
                    for (uint32_t t = 0; t < naa; t++)
                        aab[t] = (uint8_t)codon_aa(tab, sb[3 * t], sb[3 * t + 1], sb[3 * t + 2]);
                } else {
                    const uint8_t *pl = tab + 4352;
                    const uint32_t top = it.nb - 1; // the item's bases end at the first codon's base i
                    for (uint32_t t = 0; t < naa; t++) {
                        const uint32_t i = top - 3 * t;
                        aab[t] = (uint8_t)codon_aa(tab, pl[sb[i]], pl[sb[i - 1]], pl[sb[i - 2]]); // codon_tables.go:222-226
                    }
                }
            }
            for (uint32_t u0 = 0; u0 < maxn; u0 += DENSE_S) {
                for (uint32_t e = 0; e < DENSE_S; e++) {
                    const uint32_t u = u0 + e;
                    if (u < nstep) {
                        ByteSrc s2;
                        s2.p = src.p + u;
                        *reinterpret_cast<uint64_t *>(row + e * 8) = wyhash_dev(s2, (uint32_t)k, 1ull); // iterator-protein.go:87
                    }
                }
                const uint32_t n = u0 < nstep ? min((uint32_t)DENSE_S, nstep - u0) : 0u;
                flush_rows<MODE>(stage, a.out_val, a.out_pos, a.pos_width, it.obase + u0, n, 1, it.p0 + u0, lane);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host: codon tables
// codonTableFromText (seq/codon_tables.go:316-427): the 64 standard codons, then every ambiguous codon
// whose expansions agree on one amino acid, axis by axis (third base, second, first).
static int base2code_host(int b) { // seq/ambiguous_bases.go:28-67; -1 invalid
    switch (b) {
    case 'A': case 'a': return 1;  case 'C': case 'c': return 2;  case 'G': case 'g': return 4;
    case 'T': case 't': case 'U': case 'u': return 8;  case 'N': case 'n': return 15;
    case 'M': case 'm': return 3;  case 'R': case 'r': return 5;  case 'W': case 'w': return 9;
    case 'S': case 's': return 6;  case 'Y': case 'y': return 10; case 'K': case 'k': return 12;
    case 'V': case 'v': return 7;  case 'H': case 'h': return 11; case 'D': case 'd': return 13;
    case 'B': case 'b': return 14; case ' ': case '*': case '-': return 0;
    default: return -1;
    }
}

static void merge_axis(uint8_t (*t)[16][16], int axis) {
    for (int i = 1; i < 16; i++)
        for (int j = 1; j < 16; j++) {
            int mask_of[256] = {0};
            for (int c = 1; c < 16; c++) {
                const uint8_t aa = axis == 3 ? t[i][j][c] : axis == 2 ? t[i][c][j] : t[c][i][j];
                if (aa) mask_of[aa] |= c;
            }
            for (int aa = 1; aa < 256; aa++) {
                const int amb = mask_of[aa];
                if (!amb) continue;
                for (int c = 1; c < 16; c++) {
                    if ((c & amb) != c) continue;
                    if (axis == 3) t[i][j][c] = (uint8_t)aa;
                    else if (axis == 2) t[i][c][j] = (uint8_t)aa;
                    else t[c][i][j] = (uint8_t)aa;
                }
            }
        }
}

// Fills aux[4608] for transl_table `id`; false if the reference does not register that table.
bool build_codon_aux(int id, uint8_t *aux) {
    const char *aas = nullptr;
    for (int i = 0; i < B200SK_N_CODON_ROWS; i++)
        if (B200SK_CODON_ROWS[i].id == id) aas = B200SK_CODON_ROWS[i].aas;
    if (!aas) return false;
    static const char order[4] = {'T', 'C', 'A', 'G'};
    uint8_t(*t)[16][16] = reinterpret_cast<uint8_t(*)[16][16]>(aux);
    memset(aux, 0, 4608);
    for (int c = 0; c < 64; c++)
        t[base2code_host(order[c >> 4])][base2code_host(order[(c >> 2) & 3])][base2code_host(order[c & 3])] =
            (uint8_t)aas[c];
    merge_axis(t, 3);
    merge_axis(t, 2);
    merge_axis(t, 1);
    for (int b = 0; b < 256; b++) {
        const int c = base2code_host(b);
        aux[4096 + b] = c < 0 ? 0xff : (uint8_t)c;
        aux[4352 + b] = (uint8_t)pair_letter(B200SK_ALPHABET_DNA, (uint32_t)b);
    }
    return true;
}

// ------------------------------------------------------------------ launch
// The same per 16-byte vector of the whole base array (read boundaries ignored until a hit): every word is
// first tested for "all of ACGTacgt" with the PRMT class trick; only a word that fails is looked at byte by
// byte (IUPAC codes and U are legal, sketches/kmers.go:23-40), and only a truly illegal byte searches its read
// and lowers ill[read] -- illegal bases are rare, so the pass is a plain streaming read.
__global__ void __launch_bounds__(256) k_first_illegal_vec(const uint8_t *__restrict__ bases, uint64_t n_bases,
                                                           const uint64_t *__restrict__ off, uint64_t n_reads,
                                                           uint32_t *ill) {
    const uint64_t b_lo = off[0], b_hi = off[n_reads]; // the bytes that belong to reads
    (void)n_bases;
    const uint64_t nvec = (b_hi + 15) / 16;
    for (uint64_t v = b_lo / 16 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec;
         v += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 x = reinterpret_cast<const uint4 *>(bases)[v];
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t lc = w[i] | 0x20202020u;
            const uint32_t t = (lc >> 1) & 0x03030303u;
            const uint32_t u2 = t | (t >> 4);
            const uint32_t sel = __byte_perm(u2, 0u, 0x4420u);
            if (lc == __byte_perm(0x67746361u, 0u, sel)) continue; // four of a, c, g, t
            for (int j = 0; j < 4; j++) {
                const uint64_t pos = v * 16 + (uint64_t)i * 4 + j;
                if (pos >= b_hi) break;
                if (pos < b_lo) continue;
                if (base2bit_of((w[i] >> (8 * j)) & 0xffu) != 4) continue;
                uint64_t lo = 0, hi = n_reads; // read that holds byte pos: largest r with off[r] <= pos
                while (hi - lo > 1) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if (off[mid] <= pos) lo = mid; else hi = mid;
                }
                if (pos < off[lo + 1]) atomicMin(ill + lo, (uint32_t)(pos - off[lo]));
            }
        }
    }
}

cudaError_t launch_first_illegal(const uint8_t *bases, const uint64_t *off, uint64_t n_reads, uint32_t *ill,
                                 cudaStream_t st, uint64_t n_bases) {
    if (n_bases && (reinterpret_cast<uintptr_t>(bases) & 15u) == 0) {
        // off[0] may be > 0 and the array is readable up to the next 16-byte boundary (the TMA contract)
        cudaError_t e = cudaMemsetAsync(ill, 0xff, n_reads * 4, st);
        if (e != cudaSuccess) return e;
        const uint64_t nvec = (n_bases + 15) / 16;
        uint64_t cb = (nvec + 255) / 256;
        if (cb > 148 * 16) cb = 148 * 16;
        k_first_illegal_vec<<<(unsigned)cb, 256, 0, st>>>(bases, n_bases, off, n_reads, ill);
        return cudaGetLastError();
    }
    uint64_t cb = (n_reads + 7) / 8;
    if (cb > 148 * 16) cb = 148 * 16;
    if (cb == 0) cb = 1;
    k_first_illegal<<<(unsigned)cb, 256, 0, st>>>(bases, off, n_reads, ill);
    return cudaGetLastError();
}

template <int MODE, int NB = 1>
static cudaError_t launch_dense_mode(const KArgs &a, int threads, int blocks, cudaStream_t st, int *occ) {
    const void *fn = (const void *)k_dense<MODE, NB>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.sm_total);
    if (e != cudaSuccess) return e;
    if (occ) {
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, a.sm_total);
        *occ = nb < 1 ? 1 : nb;
        return e;
    }
    k_dense<MODE, NB><<<blocks, threads, a.sm_total, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_nthash_warp(const KArgs &a, cudaStream_t st, int *occ); // b200sk_nthash.cu
bool nthash_warp_fits(uint32_t span_max);

cudaError_t launch_dense(const KArgs &a, int threads, int blocks, cudaStream_t st, int *occ) {
    switch (a.mode) {
    case B200SK_MODE_NTHASH:
        if (nthash_warp_fits(a.span_max)) return launch_nthash_warp(a, st, occ); // the warp-tile kernel
        return launch_dense_mode<B200SK_MODE_NTHASH>(a, threads, blocks, st, occ);
    case B200SK_MODE_KMER:
        // canonical codes of reads that are one item each: the warp-tile kernel (a long read cut short by an illegal
        // base would leave a hole of untouched bytes in a tile's byte range); both strands: the generic dense kernel
        if (a.canonical && !a.item_first && nthash_warp_fits(a.span_max)) return launch_nthash_warp(a, st, occ);
        return launch_dense_mode<B200SK_MODE_KMER>(a, threads, blocks, st, occ);
    case B200SK_MODE_PROTEIN:
        // k <= 16, one item per read: the warp-tile kernel (amino acids in a register window)
        if (a.k <= 16 && !a.item_first && nthash_warp_fits(a.span_max)) return launch_nthash_warp(a, st, occ);
        return launch_dense_mode<B200SK_MODE_PROTEIN>(a, threads, blocks, st, occ);
    case B200SK_MODE_SIMHASH: { // counter planes: enough bits for n = k-m+1 (a.w carries m)
        const int n = a.k - a.w + 1;
        if (n < 16) return launch_dense_mode<B200SK_MODE_SIMHASH, 4>(a, threads, blocks, st, occ);
        if (n < 32) return launch_dense_mode<B200SK_MODE_SIMHASH, 5>(a, threads, blocks, st, occ);
        if (n < 64) return launch_dense_mode<B200SK_MODE_SIMHASH, 6>(a, threads, blocks, st, occ);
        return launch_dense_mode<B200SK_MODE_SIMHASH, 8>(a, threads, blocks, st, occ);
    }
    default: return cudaErrorInvalidValue;
    }
}

} // namespace b200sk
