// NextHash (sketches/iterator.go:615-665; ntHash-1 of will-rowe/nthash v0.4.0): one tile of 32 items per WARP,
// no block-wide barrier, no ordering between tiles (the output
// offsets of a dense mode follow from the read lengths: k_scan_reads wrote out_off before this kernel runs).
//
// Per tile:
//   1. one 1-D TMA bulk copy brings the tile's byte range into the warp's shared-memory buffer;
//   2. if every byte is one of ACGTacgt the tile is rewritten as cls * 40 (cls = (byte >> 1) & 3), so that
//      (in & 0x18) | (out & 0x60) is the byte offset into 16-entry pair tables
//          X[in, out] = A[in] ^ rol(A[out], k),   Y[in, out] = rol(B[in], k-1) ^ ror(B[out], 1)
//      and a rolling step is one PRMT, two conflict-free LDS.64 and one three-input XOR per 32-bit half; the
//      first k-1 bases fold two at a time through F2X / F2Y.  Any other byte: the tile is fetched again and
//      walks the general 256-entry tables (forward seed by byte, reverse seed by byte & 7);
//   3. every lane walks its item 16 steps at a time into its row of the staging area (row stride 136 B: the
//      lanes' 8-byte stores and the half-warp row reads below are both 2 wavefronts, the minimum for 64-bit
//      accesses); whole blocks run as straight-line code so that the table loads of all 16 steps are in
//      flight together;
//   4. the 32 rows leave two per store instruction -- lanes 0-15 one row, lanes 16-31 the next, each a
//      contiguous 128-byte run of out_val.
// (Tried: one cp.async.bulk shared -> global per lane and row instead of step 4.  The TMA engine takes such
// 128-byte stores at the full HBM write bandwidth -- scripts/ubench/bulkstore.cu: 6.4 TB/s -- but UBLKCP is a
// uniform-datapath instruction: with per-lane addresses the compiler emits a loop over the 32 lanes, 12
// instructions per row against 6 per warp-step here.)
// The same kernel also serves canonical NextKmer (sketches/iterator.go:668-759) on reads of one item each
// (KIND_KMER): no tables beyond the 256-byte base2bit LUT (sketches/kmers.go:23-40), the rolling update of
// iterator.go:736,740 and the canonical minimum of :754-756; reads with an illegal base stop before the first
// k-mer that holds it (k_first_illegal + read_positions).  Both-strand k-mers stay with the generic dense kernel.
// Index() of a dense mode is the running position: a second store loop writes it when the caller wants it.
// And ProteinIterator.Next (sketches/iterator-protein.go:46-90) for k <= 16 on reads of one item each
// (KIND_PROTEIN): the frame is translated codon by codon as the lane walks (all-ACGT tiles: 2-bit classes and a
// 64-entry amino-acid table per strand; any other byte: CodonTable.Get over the 4-bit IUPAC matrix), the last 16
// amino acids live in a 128-bit register window, and wyhash(seed 1) of the k newest is computed from registers.
#include "b200sk_protein.cuh"
#include "b200sk_tile.cuh"

namespace b200sk {

namespace {

#define NH_ROW 136u              // staging row stride (16 values + 8 B of padding)
#define NH_STAGE (32u * NH_ROW)  // per warp
#define NH_DESC 384u             // 32 row destinations (8 B each) + 32 first positions (4 B each)
#define NH_TAB_GENERAL 8192u     // tIn[256], tOut[256] (16 B entries)
#define NH_TAB_FAST 8192u        // fast tables start here (640 B used)
#define NH_TABLES (8192u + 1024u)
// fast-table offsets: X 0, Y 128, F2X 256, F2Y 384, F1X 512, F1Y 544

__device__ __forceinline__ uint64_t lds64(const uint8_t *sm, uint32_t o) { return *reinterpret_cast<const uint64_t *>(sm + o); }
__device__ __forceinline__ uint32_t lds8(const uint8_t *sm, uint32_t o) { return sm[o]; }

__device__ __forceinline__ uint32_t fast_word(uint32_t w, uint32_t &bad) {
    const uint32_t x = w | 0x20202020u;                // lower case
    const uint32_t t = (x >> 1) & 0x03030303u;         // class of every byte: a=0 c=1 t=2 g=3
    const uint32_t u2 = t | (t >> 4);                  // nibble pairs in bytes 0 and 2
    const uint32_t sel = __byte_perm(u2, 0u, 0x4420u); // four nibbles = PRMT selector
    bad |= x ^ __byte_perm(0x67746361u, 0u, sel);      // 'a','c','t','g' by class
    return t * 40u;                                    // (cls << 3) | (cls << 5)
}

// 16 consecutive bytes from an arbitrary shared-memory offset, as aligned words lined up with PRMT
struct Bytes16 {
    uint32_t x[4];
    __device__ __forceinline__ void load(const uint8_t *sm, uint32_t p) {
        const uint32_t a = p & ~3u, sel = 0x3210u + 0x1111u * (p & 3u);
        uint32_t w[5];
#pragma unroll
        for (int i = 0; i < 5; i++) w[i] = *reinterpret_cast<const uint32_t *>(sm + a + 4u * i);
#pragma unroll
        for (int g = 0; g < 4; g++) x[g] = __byte_perm(w[g], w[g + 1], sel);
    }
    __device__ __forceinline__ uint32_t byte(const int j) const { return __byte_perm(x[j >> 2], 0u, 0x4440u | (j & 3)); }
};

// 16 virtual steps of one lane into its staging row (slot e = virtual step v0 + e).  FAST: pair tables over
// fast bytes, else the general tables over the original bytes.  FULL: every slot is a real step, and slot 0 is
// the item's first k-mer (no outgoing base) iff FIRST0; otherwise slots [lo, hi) are real and slot `first`
// (16 = none) is the first k-mer.
template <bool CANON, bool FAST, bool FULL, bool FIRST0>
__device__ __forceinline__ void block16(uint8_t *smem, uint32_t FT, const Bytes16 &win, const Bytes16 &wout, uint32_t lo,
                                        uint32_t hi, uint32_t first, uint32_t s_row, uint64_t &fh, uint64_t &rh) {
    uint32_t po[4];
    if (FAST) {
#pragma unroll
        for (int gq = 0; gq < 4; gq++) po[gq] = (win.x[gq] & 0x18181818u) | (wout.x[gq] & 0x60606060u);
    }
    const ulonglong2 *tIn = reinterpret_cast<const ulonglong2 *>(smem), *tOut = tIn + 256;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        if (FULL || ((uint32_t)e >= lo && (uint32_t)e < hi)) {
            const bool is_first = FULL ? (FIRST0 && e == 0) : ((uint32_t)e == first);
            if (FAST) {
                const uint32_t o = __byte_perm(po[e >> 2], 0u, 0x4440u | (e & 3));
                if (is_first) {
                    fh = rol1(fh) ^ lds64(smem, FT + 512u + (o & 0x18u));
                    rh = ror1(rh) ^ lds64(smem, FT + 544u + (o & 0x18u));
                } else {
                    fh = rol1(fh) ^ lds64(smem, FT + o);
                    rh = ror1(rh) ^ lds64(smem, FT + 128u + o);
                }
            } else {
                const ulonglong2 in = tIn[win.byte(e)];
                ulonglong2 o = make_ulonglong2(0, 0);
                if (!is_first) o = tOut[wout.byte(e)];
                fh = rol1(fh) ^ o.x ^ in.x;
                rh = ror1(rh) ^ o.y ^ in.y;
            }
            *reinterpret_cast<uint64_t *>(smem + s_row + e * 8) = (CANON && rh < fh) ? rh : fh; // iterator.go:659
        }
    }
}

// KIND_KMER: 16 virtual steps of 2-bit codes (iterator.go:736,740,754).  lut: 256-byte base2bit table.
template <bool CANON, bool FULL>
__device__ __forceinline__ void block16_kmer(uint8_t *smem, const Bytes16 &win, uint32_t lo, uint32_t hi, uint32_t s_row,
                                             uint64_t mask1, uint32_t sh, uint64_t &code, uint64_t &rc) {
#pragma unroll
    for (int e = 0; e < 16; e++) {
        if (FULL || ((uint32_t)e >= lo && (uint32_t)e < hi)) {
            const uint64_t bit = smem[win.byte(e)] & 3u;
            code = ((code & mask1) << 2) | bit;
            rc = (rc >> 2) | ((bit ^ 3ull) << sh);
            *reinterpret_cast<uint64_t *>(smem + s_row + e * 8) = (CANON && rc < code) ? rc : code;
        }
    }
}

#define KIND_NTHASH 0
#define KIND_KMER 1
#define KIND_PROTEIN 2

// KIND_PROTEIN: 16 virtual steps of amino-acid k-mer hashes.  Slot e is amino-acid k-mer u = v0 + e - shift, whose
// newest amino acid has index u + k - 1.  DIR 0: the records are amino acids already; 1: forward frame; 2:
// reverse frame (codons run down the read, complemented).  FAST (DIR 1, 2): the tile holds 2-bit classes and the
// 64-entry tables at smem + 4608 (forward) / + 4672 (reverse) give the amino acid; else CodonTable.Get over the
// copy of the aux block at smem + 0.  cb: shared offset of the first base of amino acid 0 (forward) or of the base
// the first codon starts from (reverse).
// FAST == 2 (k_protein6_warp): the tile holds, at every position p, the index of the codon that STARTS there,
// c[p] * 16 + c[p+1] * 4 + c[p+2] -- one byte load per amino acid; the reverse strand reads the codon that starts two
// bases down through a table with the fields swapped (smem + 4736).
template <int DIR, int FAST>
__device__ __forceinline__ uint32_t protein_aa(const uint8_t *smem, uint32_t cb, uint32_t t) {
    if (DIR == 0) return lds8(smem, cb + t);
    if (FAST == 2) return DIR == 1 ? smem[4608u + lds8(smem, cb + 3u * t)] : smem[4736u + lds8(smem, cb - 3u * t - 2u)];
    if (DIR == 1) {
        const uint32_t p = cb + 3u * t;
        const uint32_t b0 = lds8(smem, p), b1 = lds8(smem, p + 1), b2 = lds8(smem, p + 2);
        if (FAST) return smem[4608u + b0 * 16u + b1 * 4u + b2];
        return codon_aa(smem, b0, b1, b2);
    }
    const uint32_t p = cb - 3u * t;
    const uint32_t b0 = lds8(smem, p), b1 = lds8(smem, p - 1), b2 = lds8(smem, p - 2);
    if (FAST) return smem[4672u + b0 * 16u + b1 * 4u + b2];
    const uint8_t *pl = smem + 4352; // DNA pair letters; other bytes pass through (codon_tables.go:222-226)
    return codon_aa(smem, pl[b0], pl[b1], pl[b2]);
}
template <int DIR, int FAST, bool FULL>
__device__ __forceinline__ void block16_protein(uint8_t *smem, uint32_t cb, uint32_t t0, uint32_t lo, uint32_t hi,
                                                uint32_t s_row, uint32_t k, uint64_t &wlo, uint64_t &whi) {
#pragma unroll 4 // ~60 instructions per step: fully unrolled, the variants of this block do not fit the instruction cache
    for (int e = 0; e < 16; e++) {
        if (FULL || ((uint32_t)e >= lo && (uint32_t)e < hi)) {
            const uint64_t aa = protein_aa<DIR, FAST>(smem, cb, t0 + (uint32_t)e);
            wlo = (wlo >> 8) | (whi << 56);
            whi = (whi >> 8) | (aa << 56);
            *reinterpret_cast<uint64_t *>(smem + s_row + e * 8) = wyhash_window(wlo, whi, k); // iterator-protein.go:87
        }
    }
}

template <int DIR, int FAST>
__device__ __forceinline__ void protein_warm(const uint8_t *smem, uint32_t cb, uint32_t k, uint64_t &wlo, uint64_t &whi) {
    for (uint32_t t = 0; t + 1 < k; t++) { // the k-1 amino acids before the first k-mer is complete
        const uint64_t aa = protein_aa<DIR, FAST>(smem, cb, t);
        wlo = (wlo >> 8) | (whi << 56);
        whi = (whi >> 8) | (aa << 56);
    }
}
struct NItem {
    uint64_t gb0, obase;
    uint32_t nb, nstep, p0; // p0: Index() of the item's first element
};

template <int KIND>
__device__ __forceinline__ void nthash_item(const KArgs &a, const ReadGeom &g, uint64_t item, uint64_t n_items, NItem &it) {
    it.gb0 = 0; it.obase = 0; it.nb = 0; it.nstep = 0; it.p0 = 0;
    if (item >= n_items) return;
    uint64_t r = item;
    uint32_t c = 0;
    if (a.item_first) {
        uint64_t lo = 0, hi = a.n_reads; // largest r with item_first[r] <= item
        if (a.tile_read) { lo = a.tile_read[item >> 5]; hi = lo + 32 < a.n_reads ? lo + 32 : a.n_reads; }
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (a.item_first[mid] <= item) lo = mid; else hi = mid;
        }
        r = lo;
        c = (uint32_t)(item - a.item_first[r]);
    }
    const uint64_t o0 = a.off[r], L = a.off[r + 1] - o0;
    const uint64_t orig = a.off_orig ? a.off_orig[r + 1] - a.off_orig[r] : L;
    int32_t st;
    const uint32_t np = read_positions(g, r, L, orig, &st);
    it.gb0 = o0;
    if (np == 0) return;
    const uint32_t p0 = c * a.C;
    it.nstep = min(np, p0 + a.C) - p0;
    it.p0 = p0;
    it.obase = a.out_off[r] - a.out_base + p0;
    if (KIND == KIND_PROTEIN) { // one item per read: the bases (or amino acids) of the frame's naa codons
        const uint32_t naa = it.nstep + (uint32_t)a.k - 1;
        if (g.protein_input) { it.nb = naa; it.gb0 = o0; }
        else if (a.frame > 0) { it.nb = 3u * naa; it.gb0 = o0 + (uint32_t)(a.frame - 1); }
        else { it.nb = 3u * naa; it.gb0 = o0 + (L - (uint64_t)(-a.frame)) + 1 - it.nb; } // codons run down from base L-|f|
        return;
    }
    it.nb = it.nstep + (uint32_t)a.k - 1;
    it.gb0 = o0 + p0;
}

// KC: KIND_PROTEIN only -- the k-mer size as a compile-time constant (1..16; 0 = a.k), see k_protein6_warp
template <int KIND, bool CANON, int KC = 0>
__global__ void __launch_bounds__(768, 1) k_nthash_warp(const KArgs a, uint32_t tile_bytes_cap, uint32_t warp_stride) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const int k = KC ? KC : a.k;
    if (KIND == KIND_PROTEIN) {
        // the aux block (codon matrix over IUPAC codes, base2code, pair letters), then the 64-entry tables of the
        // all-ACGT fast path: class = (byte >> 1) & 3 -> a, c, t, g; complement = class ^ 2
        for (uint32_t i = tid; i < 4608; i += blockDim.x) smem[i] = a.aux[i];
        __syncthreads();
        if (tid < 64) {
            const char letter[4] = {'A', 'C', 'T', 'G'};
            const uint32_t c0 = tid >> 4, c1 = (tid >> 2) & 3u, c2 = tid & 3u;
            smem[4608 + tid] = (uint8_t)codon_aa(smem, (uint32_t)letter[c0], (uint32_t)letter[c1], (uint32_t)letter[c2]);
            smem[4672 + tid] = (uint8_t)codon_aa(smem, (uint32_t)letter[c0 ^ 2u], (uint32_t)letter[c1 ^ 2u], (uint32_t)letter[c2 ^ 2u]);
        }
    } else if (KIND == KIND_KMER) {
        for (uint32_t b = tid; b < 256; b += blockDim.x) {
            uint32_t bit; // sketches/kmers.go:23-40 (IUPAC codes map to their first base)
            switch (b) {
            case 'A': case 'a': case 'D': case 'd': case 'H': case 'h': case 'M': case 'm':
            case 'N': case 'n': case 'R': case 'r': case 'V': case 'v': case 'W': case 'w': bit = 0; break;
            case 'B': case 'b': case 'C': case 'c': case 'S': case 's': case 'Y': case 'y': bit = 1; break;
            case 'G': case 'g': case 'K': case 'k': bit = 2; break;
            case 'T': case 't': case 'U': case 'u': bit = 3; break;
            default: bit = 0; break; // illegal bases never reach a step (read_positions stops before them)
            }
            smem[b] = (uint8_t)bit;
        }
    } else {
        ulonglong2 *tIn = reinterpret_cast<ulonglong2 *>(smem), *tOut = tIn + 256;
        for (uint32_t b = tid; b < 256; b += blockDim.x) {
            const uint64_t f = fwd_seed(b), r = rev_seed(b);
            tIn[b] = make_ulonglong2(f, rol64(r, (unsigned)(k - 1)));
            tOut[b] = make_ulonglong2(rol64(f, (unsigned)k), ror64(r, 1));
        }
        if (tid < 16) {
            const char letter[4] = {'A', 'C', 'T', 'G'}; // class = (byte >> 1) & 3
            const uint32_t lo = tid & 3u, hi = tid >> 2;
            const uint64_t Alo = fwd_seed((uint32_t)letter[lo]), Ahi = fwd_seed((uint32_t)letter[hi]);
            const uint64_t Blo = rev_seed((uint32_t)letter[lo]), Bhi = rev_seed((uint32_t)letter[hi]);
            uint64_t *q = reinterpret_cast<uint64_t *>(smem + NH_TAB_FAST);
            q[tid] = Alo ^ rol64(Ahi, (unsigned)k);                                          // X
            q[16 + tid] = rol64(Blo, (unsigned)(k - 1)) ^ ror64(Bhi, 1);                     // Y
            q[32 + tid] = rol64(Alo, 1) ^ Ahi;                                               // F2X: lo first, hi second
            q[48 + tid] = ror64(rol64(Blo, (unsigned)(k - 1)), 1) ^ rol64(Bhi, (unsigned)(k - 1)); // F2Y
            if (tid < 4) {
                q[64 + tid] = Alo;                                  // F1X
                q[68 + tid] = rol64(Blo, (unsigned)(k - 1));        // F1Y
            }
        }
    }
    const uint32_t region = NH_TABLES + wid * warp_stride;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + region);
    const uint32_t s_tile = region + 16;
    uint8_t *tilebuf = smem + s_tile;
    const uint32_t s_stage = s_tile + tile_bytes_cap;
    const uint32_t s_desc = s_stage + NH_STAGE;
    if (lane == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads(); // tables + barriers ready; the only block-wide barrier
    const uint64_t n_items = a.n_items_dev ? *a.n_items_dev : a.n_items;
    const uint64_t total = a.out_off[a.n_reads] - a.out_base;
    if (total > a.capacity) {
        if (blockIdx.x == 0 && tid == 0) atomicOr(a.flags, B200SK_FLAG_CAPACITY);
        return;
    }
    const ReadGeom g = a.geom();
    const uint32_t FT = NH_TAB_FAST;
    uint32_t parity = 0;
    for (;;) {
        uint64_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.ticket, 1ULL);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        const uint64_t item0 = tile * 32ull;
        if (item0 >= n_items) break;
        const uint32_t nvalid = (uint32_t)min((uint64_t)32, n_items - item0);
        NItem it;
        nthash_item<KIND>(a, g, item0 + lane, n_items, it);
        const uint64_t lo = __shfl_sync(0xffffffffu, it.gb0, 0);
        const uint64_t hi = __shfl_sync(0xffffffffu, it.gb0 + it.nb, (int)nvalid - 1);
        const uint64_t lo_al = lo & ~15ULL;
        const uint64_t span = hi > lo_al ? hi - lo_al : 0;
        const uint32_t bytes = (uint32_t)((span + 15ULL) & ~15ULL);
        const bool span_ok = bytes + 16u <= tile_bytes_cap; // + the word-granular look-ahead of the last block
        if (!span_ok && lane == 0) atomicOr(a.flags, B200SK_FLAG_SPAN);
        bool fast = false;
        if (bytes && span_ok) {
            if (lane == 0) {
                fence_proxy_async(); // the previous tile's generic-proxy accesses precede the async write
                mbar_expect_tx(mbar, bytes);
                tma_load_1d(tilebuf, a.bases + lo_al, bytes, mbar);
            }
            mbar_wait(mbar, parity);
            parity ^= 1u;
            uint32_t bad = 0;
            const bool try_fast = KIND == KIND_NTHASH || (KIND == KIND_PROTEIN && !g.protein_input);
            if (try_fast)
            for (uint32_t o = lane * 16u; o < bytes; o += 512u) {
                uint4 v = *reinterpret_cast<uint4 *>(tilebuf + o);
                v.x = fast_word(v.x, bad); v.y = fast_word(v.y, bad);
                v.z = fast_word(v.z, bad); v.w = fast_word(v.w, bad);
                if (KIND == KIND_PROTEIN) { // plain classes 0..3 (cls * 40 = cls << 3 | cls << 5)
                    v.x = (v.x >> 3) & 0x03030303u; v.y = (v.y >> 3) & 0x03030303u;
                    v.z = (v.z >> 3) & 0x03030303u; v.w = (v.w >> 3) & 0x03030303u;
                }
                *reinterpret_cast<uint4 *>(tilebuf + o) = v;
            }
            fast = try_fast && !__any_sync(0xffffffffu, bad != 0);
            if (try_fast && !fast) { // some other byte (alignment slop included): the original bytes again, general tables
                __syncwarp();
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(mbar, bytes);
                    tma_load_1d(tilebuf, a.bases + lo_al, bytes, mbar);
                }
                mbar_wait(mbar, parity);
                parity ^= 1u;
            }
        }
        __syncwarp();
        const uint32_t nstep = span_ok ? it.nstep : 0u;
        // Virtual steps: v = u + shift with shift = (element index of the item's first hash in out_val) mod 4, so
        // that every staging row starts on a 32-byte sector of global memory: rows that straddle sectors cost
        // a third of the write bandwidth (scripts/ubench/bulkstore.cu: 2.96 vs 4.0 TB/s for this pattern).
        uint64_t *g0 = a.out_val + it.obase;
        const uint32_t shift = nstep ? (uint32_t)((reinterpret_cast<uintptr_t>(g0) >> 3) & 3u) : 0u;
        const uint32_t vend = shift + nstep; // one past the last virtual step
        // where the lane's rows go: address of virtual step 0
        *reinterpret_cast<uint64_t *>(smem + s_desc + lane * 8u) = reinterpret_cast<uint64_t>(g0 - shift);
        *reinterpret_cast<uint32_t *>(smem + s_desc + 256u + lane * 4u) = it.p0 - shift; // Index() of virtual step 0
        uint32_t maxv = vend;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxv = max(maxv, __shfl_xor_sync(0xffffffffu, maxv, o));
        const uint32_t sb = nstep ? s_tile + (uint32_t)(it.gb0 - lo_al) : s_tile + 8u; // idle lanes: a safe base
        const uint32_t sbv = sb - shift;                                               // byte of virtual base 0
        uint64_t fh = 0, rh = 0; // KIND_KMER: code and reverse-complement code
        const uint32_t ksh = 2u * (uint32_t)(k - 1);
        const uint64_t kmask1 = (1ull << ksh) - 1ull; // iterator.go:699
        // KIND_PROTEIN: (fh, rh) = the 128-bit window (low, high) of the last 16 amino acids; cb as in protein_aa
        const int pdir = KIND != KIND_PROTEIN ? 0 : g.protein_input ? 0 : a.frame > 0 ? 1 : 2;
        const uint32_t cb = pdir == 2 ? sb + it.nb - 1u : sb;
        if (KIND == KIND_PROTEIN) {
            if (nstep) {
                if (pdir == 0) protein_warm<0, false>(smem, cb, (uint32_t)k, fh, rh);
                else if (pdir == 1) { if (fast) protein_warm<1, true>(smem, cb, (uint32_t)k, fh, rh); else protein_warm<1, false>(smem, cb, (uint32_t)k, fh, rh); }
                else { if (fast) protein_warm<2, true>(smem, cb, (uint32_t)k, fh, rh); else protein_warm<2, false>(smem, cb, (uint32_t)k, fh, rh); }
            }
        } else if (KIND == KIND_KMER) {
            if (nstep)
                for (int j = 0; j < k - 1; j++) {
                    const uint64_t bit = smem[lds8(smem, sb + j)] & 3u;
                    fh = (fh << 2) | bit;
                    rh = (rh >> 2) | ((bit ^ 3ull) << ksh);
                }
        } else if (nstep) {
            if (fast) {
                int j = 0;
                for (; j + 1 < k - 1; j += 2) {
                    const uint32_t o = (lds8(smem, sb + j) & 0x18u) | (lds8(smem, sb + j + 1) & 0x60u);
                    fh = rol64(fh, 2) ^ lds64(smem, FT + 256u + o);
                    rh = ror64(rh, 2) ^ lds64(smem, FT + 384u + o);
                }
                if (j < k - 1) {
                    const uint32_t o = lds8(smem, sb + j) & 0x18u;
                    fh = rol1(fh) ^ lds64(smem, FT + 512u + o);
                    rh = ror1(rh) ^ lds64(smem, FT + 544u + o);
                }
            } else {
                const ulonglong2 *tIn = reinterpret_cast<const ulonglong2 *>(smem);
                for (int j = 0; j < k - 1; j++) {
                    const ulonglong2 e = tIn[lds8(smem, sb + j)];
                    fh = rol1(fh) ^ e.x;
                    rh = ror1(rh) ^ e.y;
                }
            }
        }
        const uint32_t last_block = vend ? ((vend - 1) / 16u) * 16u : 0u; // first virtual step of the lane's last block
        const uint32_t s_row = s_stage + lane * NH_ROW;
        __syncwarp();
        uint64_t *rowdst[8]; // destinations (virtual step 0) of rows 4i + lane/8
#pragma unroll
        for (int i = 0; i < 8; i++) rowdst[i] = reinterpret_cast<uint64_t *>(lds64(smem, s_desc + (4u * i + (lane >> 3)) * 8u));
        for (uint32_t v0 = 0; v0 < maxv; v0 += 16u) {
            // lanes whose item is finished (or absent) re-read their own last block: the loads stay
            // unconditional and inside the tile
            const uint32_t vl = min(v0, last_block);
            Bytes16 win, wout;
            win.load(smem, sbv + vl + (uint32_t)k - 1);
            wout.load(smem, sbv + vl - 1);
            // real slots of this block: [lo, hi); the item's first k-mer is virtual step `shift`
            const uint32_t lo = v0 == 0 ? shift : 0u;
            const uint32_t hi = vend > v0 ? min(16u, vend - v0) : 0u;
            const uint32_t first = v0 == 0 ? shift : 16u;
            const bool full = __all_sync(0xffffffffu, lo == 0u && hi == 16u && shift == 0u) ||
                              (v0 != 0 && __all_sync(0xffffffffu, hi == 16u));
            // straight-line code for whole blocks; the predicated variant only for a block some lane does not fill
            if (KIND == KIND_PROTEIN) {
                const uint32_t t0 = v0 - shift + (uint32_t)k - 1u; // amino-acid index of slot 0's newest amino acid
                const uint32_t kk = (uint32_t)k;
#define B200SK_PBLOCK(DIR, FAST_)                                                                          \
    if (full) block16_protein<DIR, FAST_, true>(smem, cb, t0, lo, hi, s_row, kk, fh, rh);                  \
    else block16_protein<DIR, FAST_, false>(smem, cb, t0, lo, hi, s_row, kk, fh, rh);
                if (pdir == 0) { B200SK_PBLOCK(0, false) }
                else if (pdir == 1) { if (fast) { B200SK_PBLOCK(1, true) } else { B200SK_PBLOCK(1, false) } }
                else { if (fast) { B200SK_PBLOCK(2, true) } else { B200SK_PBLOCK(2, false) } }
#undef B200SK_PBLOCK
            } else if (KIND == KIND_KMER) {
                if (full) block16_kmer<CANON, true>(smem, win, lo, hi, s_row, kmask1, ksh, fh, rh);
                else block16_kmer<CANON, false>(smem, win, lo, hi, s_row, kmask1, ksh, fh, rh);
            } else if (fast) {
                if (full) { if (v0 == 0) block16<CANON, true, true, true>(smem, FT, win, wout, lo, hi, first, s_row, fh, rh);
                            else block16<CANON, true, true, false>(smem, FT, win, wout, lo, hi, first, s_row, fh, rh); }
                else block16<CANON, true, false, false>(smem, FT, win, wout, lo, hi, first, s_row, fh, rh);
            } else {
                if (full) { if (v0 == 0) block16<CANON, false, true, true>(smem, FT, win, wout, lo, hi, first, s_row, fh, rh);
                            else block16<CANON, false, true, false>(smem, FT, win, wout, lo, hi, first, s_row, fh, rh); }
                else block16<CANON, false, false, false>(smem, FT, win, wout, lo, hi, first, s_row, fh, rh);
            }
            __syncwarp();
            // flush: rows 2i and 2i+1 per store instruction
            const uint32_t half = lane >> 4, e = lane & 15u;
            if (full) {
                // whole, sector-aligned rows: four rows per store instruction, eight lanes per row, 16 bytes per
                // lane (rows start on 32-byte sectors, so the 16-byte stores are aligned); the eight rows a lane
                // serves keep their destinations in registers
                const uint32_t q = lane >> 3, e2 = (lane & 7u) * 2u;
#pragma unroll
                for (uint32_t i = 0; i < 8u; i++) {
                    const uint32_t src = 4u * i + q;
                    ulonglong2 v;
                    v.x = lds64(smem, s_stage + src * NH_ROW + e2 * 8u);
                    v.y = lds64(smem, s_stage + src * NH_ROW + e2 * 8u + 8u);
                    *reinterpret_cast<ulonglong2 *>(rowdst[i] + v0 + e2) = v;
                }
            } else {
                const uint32_t lohi = lo | (hi << 8);
                for (uint32_t i = 0; i < 16u; i++) {
                    const uint32_t src = 2u * i + half;
                    const uint32_t lh = __shfl_sync(0xffffffffu, lohi, (int)src);
                    if (e >= (lh & 0xffu) && e < (lh >> 8)) {
                        uint64_t *dst = reinterpret_cast<uint64_t *>(lds64(smem, s_desc + src * 8u));
                        dst[v0 + e] = lds64(smem, s_stage + src * NH_ROW + e * 8u);
                    }
                }
            }
            if (a.out_pos) { // Index() of a dense mode is the running position (iterator.go:776)
                const uint32_t lohi = lo | (hi << 8);
                for (uint32_t i = 0; i < 16u; i++) {
                    const uint32_t src = 2u * i + half;
                    const uint32_t lh = __shfl_sync(0xffffffffu, lohi, (int)src);
                    if (e >= (lh & 0xffu) && e < (lh >> 8)) {
                        const uint64_t *dst = reinterpret_cast<const uint64_t *>(lds64(smem, s_desc + src * 8u));
                        const uint32_t p0v = *reinterpret_cast<const uint32_t *>(smem + s_desc + 256u + src * 4u);
                        store_pos(a.out_pos, a.pos_width, (uint64_t)(dst - a.out_val) + v0 + e, p0v + v0 + e);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ProteinIterator.Next (sketches/iterator-protein.go:46-90) for ALL SIX frames of every read in one launch
// (BASELINE.json config 5: "ProteinIterator k=11 over 6-frame-translated reads").  Same tiles, tables, register
// window and row flush as KIND_PROTEIN above, but a tile's reads are fetched ONCE (one TMA bulk copy of the whole
// reads), rewritten to 2-bit classes once, and walked six times -- frame 1, 2, 3 down the forward strand, -1, -2, -3
// down the reverse complement (seq/codon_tables.go:205-285) -- each frame into its own value array at the offsets
// k_scan_reads wrote for it.  Against six launches: a sixth of the input traffic, tile set-up and conversion.
// Reads of one item each (the caller checks), k <= 16, values only.
// K: the k-mer size as a compile-time constant (1..16), so that wyhash's byte shuffles, shifts and its length word fold
// into immediates (the run-time-k block spends a fifth of its instructions on them, an indirect branch per k-mer included).
template <int K>
__global__ void __launch_bounds__(768, 1) k_protein6_warp(const KArgs a, uint32_t tile_bytes_cap, uint32_t warp_stride) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    constexpr uint32_t k = (uint32_t)K;
    for (uint32_t i = tid; i < 4608; i += blockDim.x) smem[i] = a.aux[i];
    __syncthreads();
    if (tid < 64) {
        const char letter[4] = {'A', 'C', 'T', 'G'};
        const uint32_t c0 = tid >> 4, c1 = (tid >> 2) & 3u, c2 = tid & 3u;
        smem[4608 + tid] = (uint8_t)codon_aa(smem, (uint32_t)letter[c0], (uint32_t)letter[c1], (uint32_t)letter[c2]);
        smem[4672 + tid] = (uint8_t)codon_aa(smem, (uint32_t)letter[c0 ^ 2u], (uint32_t)letter[c1 ^ 2u], (uint32_t)letter[c2 ^ 2u]);
        // reverse strand by the index of the codon two bases DOWN (fields swapped: c0 is that codon's third base)
        smem[4736 + tid] = (uint8_t)codon_aa(smem, (uint32_t)letter[c2 ^ 2u], (uint32_t)letter[c1 ^ 2u], (uint32_t)letter[c0 ^ 2u]);
    }
    const uint32_t region = NH_TABLES + wid * warp_stride;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + region);
    const uint32_t s_tile = region + 16;
    uint8_t *tilebuf = smem + s_tile;
    const uint32_t s_stage = s_tile + tile_bytes_cap;
    const uint32_t s_desc = s_stage + NH_STAGE;
    if (lane == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint64_t n_reads = a.n_reads;
    bool over = false;
#pragma unroll
    for (int fi = 0; fi < 6; fi++) over |= a.fr_off[fi][n_reads] > a.capacity;
    if (over) {
        if (blockIdx.x == 0 && tid == 0) atomicOr(a.flags, B200SK_FLAG_CAPACITY);
        return;
    }
    ReadGeom g = a.geom();
    uint32_t parity = 0;
    for (;;) {
        uint64_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.ticket, 1ULL);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        const uint64_t r0 = tile * 32ull;
        if (r0 >= n_reads) break;
        const uint32_t nvalid = (uint32_t)min((uint64_t)32, n_reads - r0);
        const uint64_t r = r0 + lane;
        const bool valid = lane < nvalid;
        const uint64_t o0 = valid ? a.off[r] : 0, L = valid ? a.off[r + 1] - o0 : 0;
        const uint64_t lo = __shfl_sync(0xffffffffu, o0, 0);
        const uint64_t hi = __shfl_sync(0xffffffffu, o0 + L, (int)nvalid - 1);
        const uint64_t lo_al = lo & ~15ULL;
        const uint64_t span = hi > lo_al ? hi - lo_al : 0;
        const uint32_t bytes = (uint32_t)((span + 15ULL) & ~15ULL);
        const bool span_ok = bytes + 16u <= tile_bytes_cap;
        if (!span_ok && lane == 0) atomicOr(a.flags, B200SK_FLAG_SPAN);
        bool fast = false;
        if (bytes && span_ok) {
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(mbar, bytes);
                tma_load_1d(tilebuf, a.bases + lo_al, bytes, mbar);
            }
            mbar_wait(mbar, parity);
            parity ^= 1u;
            uint32_t bad = 0;
            for (uint32_t o = lane * 16u; o < bytes; o += 512u) {
                uint4 v = *reinterpret_cast<uint4 *>(tilebuf + o);
                v.x = fast_word(v.x, bad); v.y = fast_word(v.y, bad);
                v.z = fast_word(v.z, bad); v.w = fast_word(v.w, bad);
                v.x = (v.x >> 3) & 0x03030303u; v.y = (v.y >> 3) & 0x03030303u; // plain classes 0..3
                v.z = (v.z >> 3) & 0x03030303u; v.w = (v.w >> 3) & 0x03030303u;
                *reinterpret_cast<uint4 *>(tilebuf + o) = v;
            }
            fast = !__any_sync(0xffffffffu, bad != 0);
            if (fast) {
                // classes -> codon indices in place: position p gets c[p] * 16 + c[p+1] * 4 + c[p+2] (bytes <= 63: no
                // carries between the bytes of a word).  Every lane reads its 16 bytes and the next word before any
                // lane writes; the word behind the tile's last chunk is slack (its two codons are never looked up).
                for (uint32_t base = 0; base < bytes; base += 512u) { // (every lane takes every trip: __syncwarp)
                    const uint32_t o = base + lane * 16u;
                    const bool mine = o < bytes;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    uint32_t nx = 0;
                    if (mine) {
                        v = *reinterpret_cast<const uint4 *>(tilebuf + o);
                        nx = *reinterpret_cast<const uint32_t *>(tilebuf + o + 16u);
                    }
                    __syncwarp();
                    if (mine) {
                        uint4 c;
                        c.x = v.x * 16u + __funnelshift_r(v.x, v.y, 8) * 4u + __funnelshift_r(v.x, v.y, 16);
                        c.y = v.y * 16u + __funnelshift_r(v.y, v.z, 8) * 4u + __funnelshift_r(v.y, v.z, 16);
                        c.z = v.z * 16u + __funnelshift_r(v.z, v.w, 8) * 4u + __funnelshift_r(v.z, v.w, 16);
                        c.w = v.w * 16u + __funnelshift_r(v.w, nx, 8) * 4u + __funnelshift_r(v.w, nx, 16);
                        *reinterpret_cast<uint4 *>(tilebuf + o) = c;
                    }
                    __syncwarp();
                }
            }
            if (!fast) { // some other byte: the original bytes again, CodonTable.Get over the IUPAC matrix
                __syncwarp();
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(mbar, bytes);
                    tma_load_1d(tilebuf, a.bases + lo_al, bytes, mbar);
                }
                mbar_wait(mbar, parity);
                parity ^= 1u;
            }
        }
        __syncwarp();
        const uint32_t sb0 = s_tile + (uint32_t)(o0 - lo_al); // the read's first base
        const uint32_t s_row = s_stage + lane * NH_ROW;
        for (int fi = 0; fi < 6; fi++) {
            const int frame = fi < 3 ? fi + 1 : 2 - fi; // 1, 2, 3, -1, -2, -3
            g.frame = frame;
            int32_t st;
            const uint32_t nstep = (valid && span_ok) ? read_positions(g, r, L, L, &st) : 0u;
            uint64_t *g0 = a.fr_val[fi] + (valid ? a.fr_off[fi][r] : 0);
            const uint32_t shift = nstep ? (uint32_t)((reinterpret_cast<uintptr_t>(g0) >> 3) & 3u) : 0u;
            const uint32_t vend = shift + nstep;
            *reinterpret_cast<uint64_t *>(smem + s_desc + lane * 8u) = reinterpret_cast<uint64_t>(g0 - shift);
            uint32_t maxv = vend;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) maxv = max(maxv, __shfl_xor_sync(0xffffffffu, maxv, o));
            // cb: the first base of amino acid 0 (forward) / the base the first codon runs down from (reverse)
            const uint32_t cb = !nstep ? s_tile + 8u : frame > 0 ? sb0 + (uint32_t)(frame - 1) : sb0 + (uint32_t)L - (uint32_t)(-frame);
            uint64_t wlo = 0, whi = 0;
            if (nstep) {
                if (frame > 0) { if (fast) protein_warm<1, 2>(smem, cb, k, wlo, whi); else protein_warm<1, 0>(smem, cb, k, wlo, whi); }
                else { if (fast) protein_warm<2, 2>(smem, cb, k, wlo, whi); else protein_warm<2, 0>(smem, cb, k, wlo, whi); }
            }
            const uint32_t last_block = vend ? ((vend - 1) / 16u) * 16u : 0u;
            __syncwarp();
            uint64_t *rowdst[8];
#pragma unroll
            for (int i = 0; i < 8; i++) rowdst[i] = reinterpret_cast<uint64_t *>(lds64(smem, s_desc + (4u * i + (lane >> 3)) * 8u));
            for (uint32_t v0 = 0; v0 < maxv; v0 += 16u) {
                const uint32_t vl = min(v0, last_block); // finished lanes re-read their own last block
                const uint32_t lo_s = v0 == 0 ? shift : 0u;
                const uint32_t hi_s = vend > v0 ? min(16u, vend - v0) : 0u;
                const bool full = __all_sync(0xffffffffu, lo_s == 0u && hi_s == 16u && shift == 0u) ||
                                  (v0 != 0 && __all_sync(0xffffffffu, hi_s == 16u));
                const uint32_t t0 = vl - shift + k - 1u; // amino-acid index of slot 0's newest amino acid
#define B200SK_PBLOCK(DIR, FAST_)                                                                          \
    if (full) block16_protein<DIR, FAST_, true>(smem, cb, t0, lo_s, hi_s, s_row, k, wlo, whi);             \
    else block16_protein<DIR, FAST_, false>(smem, cb, t0, lo_s, hi_s, s_row, k, wlo, whi);
                if (frame > 0) { if (fast) { B200SK_PBLOCK(1, 2) } else { B200SK_PBLOCK(1, 0) } }
                else { if (fast) { B200SK_PBLOCK(2, 2) } else { B200SK_PBLOCK(2, 0) } }
#undef B200SK_PBLOCK
                __syncwarp();
                const uint32_t half = lane >> 4, e = lane & 15u;
                if (full) {
                    const uint32_t q = lane >> 3, e2 = (lane & 7u) * 2u;
#pragma unroll
                    for (uint32_t i = 0; i < 8u; i++) {
                        const uint32_t src = 4u * i + q;
                        ulonglong2 v;
                        v.x = lds64(smem, s_stage + src * NH_ROW + e2 * 8u);
                        v.y = lds64(smem, s_stage + src * NH_ROW + e2 * 8u + 8u);
                        *reinterpret_cast<ulonglong2 *>(rowdst[i] + v0 + e2) = v;
                    }
                } else {
                    const uint32_t lohi = lo_s | (hi_s << 8);
                    for (uint32_t i = 0; i < 16u; i++) {
                        const uint32_t src = 2u * i + half;
                        const uint32_t lh = __shfl_sync(0xffffffffu, lohi, (int)src);
                        if (e >= (lh & 0xffu) && e < (lh >> 8)) {
                            uint64_t *dst = reinterpret_cast<uint64_t *>(lds64(smem, s_desc + src * 8u));
                            dst[v0 + e] = lds64(smem, s_stage + src * NH_ROW + e * 8u);
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
}

} // namespace

cudaError_t launch_protein6_warp(const KArgs &a, cudaStream_t st) {
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
    }
    const uint32_t tile_cap = ((32u * a.span_max + 32u + 15u) & ~15u) + 16u;
    const uint32_t stride = (16u + tile_cap + NH_STAGE + NH_DESC + 15u) & ~15u;
    int nw = (int)((227u * 1024u - NH_TABLES) / stride);
    if (nw > 24) nw = 24;
    if (nw < 1) return cudaErrorInvalidValue;
    const uint32_t sm_total = NH_TABLES + (uint32_t)nw * stride;
    void (*fn)(const KArgs, uint32_t, uint32_t) = nullptr;
    switch (a.k) {
#define B200SK_P6(K) case K: fn = k_protein6_warp<K>; break;
        B200SK_P6(1) B200SK_P6(2) B200SK_P6(3) B200SK_P6(4) B200SK_P6(5) B200SK_P6(6) B200SK_P6(7) B200SK_P6(8)
        B200SK_P6(9) B200SK_P6(10) B200SK_P6(11) B200SK_P6(12) B200SK_P6(13) B200SK_P6(14) B200SK_P6(15) B200SK_P6(16)
#undef B200SK_P6
    default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_total);
    if (e != cudaSuccess) return e;
    fn<<<sm_count, nw * 32, sm_total, st>>>(a, tile_cap, stride);
    return cudaGetLastError();
}

// occ != nullptr: the grid is sized here, report 1.
cudaError_t launch_nthash_warp(const KArgs &a, cudaStream_t st, int *occ) {
    if (occ) { *occ = 1; return cudaSuccess; }
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
    }
    const uint32_t tile_cap = ((32u * a.span_max + 32u + 15u) & ~15u) + 16u;
    const uint32_t stride = (16u + tile_cap + NH_STAGE + NH_DESC + 15u) & ~15u;
    int nw = (int)((227u * 1024u - NH_TABLES) / stride);
    if (nw > 24) nw = 24;
    if (nw < 1) return cudaErrorInvalidValue;
    const uint32_t sm_total = NH_TABLES + (uint32_t)nw * stride;
    void (*fn)(const KArgs, uint32_t, uint32_t);
    if (a.mode == B200SK_MODE_PROTEIN) {
        switch (a.k) {
#define B200SK_PK(K) case K: fn = k_nthash_warp<KIND_PROTEIN, true, K>; break;
            B200SK_PK(1) B200SK_PK(2) B200SK_PK(3) B200SK_PK(4) B200SK_PK(5) B200SK_PK(6) B200SK_PK(7) B200SK_PK(8)
            B200SK_PK(9) B200SK_PK(10) B200SK_PK(11) B200SK_PK(12) B200SK_PK(13) B200SK_PK(14) B200SK_PK(15) B200SK_PK(16)
#undef B200SK_PK
        default: return cudaErrorInvalidValue;
        }
    }
    else if (a.mode == B200SK_MODE_KMER) fn = k_nthash_warp<KIND_KMER, true>;
    else fn = a.canonical ? k_nthash_warp<KIND_NTHASH, true> : k_nthash_warp<KIND_NTHASH, false>;
    cudaError_t e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_total);
    if (e != cudaSuccess) return e;
    fn<<<sm_count, nw * 32, sm_total, st>>>(a, tile_cap, stride);
    return cudaGetLastError();
}

bool nthash_warp_fits(uint32_t span_max) {
    const uint32_t tile_cap = ((32u * span_max + 32u + 15u) & ~15u) + 16u;
    const uint32_t stride = (16u + tile_cap + NH_STAGE + NH_DESC + 15u) & ~15u;
    return (227u * 1024u - NH_TABLES) / stride >= 4u;
}

} // namespace b200sk
