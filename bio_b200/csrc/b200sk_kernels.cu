// Sketching kernels, hand-written for sm_100a (B200).
//
// Work decomposition (DESIGN.md "kernels"): the batch is a list of ITEMS; an
// item is one read, or -- for reads with more than C output positions -- one
// chunk of C consecutive positions of a read (with the halo of bases its
// first window needs).  A CTA takes a TILE of blockDim.x consecutive items
// from a ticket counter, pulls the tile's contiguous byte range from HBM into
// shared memory with ONE 1-D TMA bulk copy, and every thread then walks its own
// item sequentially with the O(1) rolling ntHash update (sequential rolling is
// ~10x fewer integer ops per base than a lane-parallel XOR scan).  Emitted
// elements are staged per thread in shared memory, the tile's output range is
// allocated in global order by a decoupled look-back over per-tile status
// words, and the staged elements leave through a coalesced ordered copy.
//
// Reference behaviour restated here (bit-exact, checked against oracle/):
//   ntHash roll            will-rowe/nthash v0.4.0 via sketches/iterator.go:659, sketch.go:212,319,367
//   NextHash               sketches/iterator.go:658-665
//   NextMinimizer          sketches/sketch.go:205-309   (== leftmost window minimum, de-duplicated by position)
//   NextSyncmer            sketches/sketch.go:312-477   (== bounded closed syncmer closed form)
#include "b200sk_tile.cuh"
#include "b200sk_protein.cuh"

namespace b200sk {

// ------------------------------------------------------------------ small kernels

// Per-read chunk count, total number of items and the longest read.
// meta[0] += items, meta[1] = max(meta[1], L)
__global__ void k_prepass(const uint64_t *__restrict__ off, const uint64_t *__restrict__ off_orig,
                          uint64_t n_reads, const ReadGeom g, uint32_t C, unsigned long long *meta) {
    unsigned long long items = 0, maxlen = 0;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads;
         r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t L = off[r + 1] - off[r];
        const uint64_t orig = off_orig ? off_orig[r + 1] - off_orig[r] : L;
        int32_t st;
        const uint32_t np = read_positions(g, r, L, orig, &st);
        items += chunks_of(np, C);
        maxlen = L > maxlen ? L : maxlen;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        items += __shfl_xor_sync(0xffffffffu, items, o);
        unsigned long long m = __shfl_xor_sync(0xffffffffu, maxlen, o);
        maxlen = m > maxlen ? m : maxlen;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&meta[0], items);
        atomicMax(&meta[1], maxlen);
    }
}

// Exclusive scan over reads, single pass with look-back, 1024 reads per tile (256 threads x 4).
//   COUNTS == false: chunks per read  -> item_first[0..n_reads]
//   COUNTS == true : elements per read (dense modes) -> out[0..n_reads] (+ base), and the per-read status
template <bool COUNTS>
__global__ void __launch_bounds__(256) k_scan_reads(const uint64_t *__restrict__ off,
                                                    const uint64_t *__restrict__ off_orig, uint64_t n_reads,
                                                    const ReadGeom g, uint32_t C, uint64_t base,
                                                    uint64_t *item_first, int32_t *status_out,
                                                    uint64_t *tile_state, unsigned long long *ticket,
                                                    uint64_t *tile_read = nullptr) {
    __shared__ uint32_t warp_sums[34];
    __shared__ uint64_t sh_tile, sh_base;
    for (;;) {
        if (threadIdx.x == 0) sh_tile = atomicAdd(ticket, 1ULL);
        __syncthreads();
        const uint64_t tile = sh_tile;
        const uint64_t r0 = tile * 1024ULL + threadIdx.x * 4ULL;
        if (tile * 1024ULL >= n_reads) break;
        uint32_t c[4], sum = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            c[i] = 0;
            const uint64_t r = r0 + i;
            if (r < n_reads) {
                const uint64_t L = off[r + 1] - off[r];
                const uint64_t orig = off_orig ? off_orig[r + 1] - off_orig[r] : L;
                int32_t st;
                const uint32_t np = read_positions(g, r, L, orig, &st);
                if (COUNTS) {
                    c[i] = (uint32_t)dense_count(g, np, st);
                    if (status_out) status_out[r] = st;
                } else {
                    c[i] = chunks_of(np, C);
                }
            }
            sum += c[i];
        }
        uint32_t total;
        uint32_t excl = block_excl_scan(sum, warp_sums, &total);
        if (threadIdx.x < 32) {
            uint64_t b = lookback_exclusive(tile_state, tile, total);
            if (threadIdx.x == 0) sh_base = b;
        }
        __syncthreads();
        uint64_t run = base + sh_base + excl;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint64_t r = r0 + i;
            if (r < n_reads) item_first[r] = run;
            if (!COUNTS && tile_read && r < n_reads) // the read that holds item 32 t, for every t this read covers
                for (uint64_t t = (run + 31) >> 5; (t << 5) < run + c[i]; t++) tile_read[t] = r;
            run += c[i];
            if (r + 1 == n_reads) item_first[n_reads] = run;
        }
        __syncthreads();
    }
}

// circular=true: "seq2 = S + S[0:k-1]" (iterator.go:642-646, sketch.go:106-110,163-167).
// off2[r] = off[r] - off[0] + sum_{i<r} ext_i, ext_i = k-1 (0 for reads shorter than k-1, which
// every constructor rejects before extending).
__global__ void __launch_bounds__(256) k_circ_offsets(const uint64_t *__restrict__ off, uint64_t n_reads, int k,
                                                      uint64_t *off2, uint64_t *tile_state,
                                                      unsigned long long *ticket) {
    __shared__ uint32_t warp_sums[34];
    __shared__ uint64_t sh_tile, sh_base;
    const uint64_t base0 = off[0];
    for (;;) {
        if (threadIdx.x == 0) sh_tile = atomicAdd(ticket, 1ULL);
        __syncthreads();
        const uint64_t tile = sh_tile;
        const uint64_t r0 = tile * 1024ULL + threadIdx.x * 4ULL;
        if (tile * 1024ULL >= n_reads) break;
        uint32_t c[4], sum = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            c[i] = 0;
            const uint64_t r = r0 + i;
            if (r < n_reads) c[i] = (off[r + 1] - off[r] >= (uint64_t)(k - 1)) ? (uint32_t)(k - 1) : 0u;
            sum += c[i];
        }
        uint32_t total;
        uint32_t excl = block_excl_scan(sum, warp_sums, &total);
        if (threadIdx.x < 32) {
            uint64_t b = lookback_exclusive(tile_state, tile, total);
            if (threadIdx.x == 0) sh_base = b;
        }
        __syncthreads();
        uint64_t run = sh_base + excl;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint64_t r = r0 + i;
            if (r < n_reads) off2[r] = off[r] - base0 + run;
            run += c[i];
            if (r + 1 == n_reads) off2[n_reads] = off[n_reads] - base0 + run;
        }
        __syncthreads();
    }
}

// one warp per read: copy the read, then its first k-1 bases again
__global__ void k_circ_copy(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ off,
                            const uint64_t *__restrict__ off2, uint64_t n_reads, uint8_t *bases2) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n_reads; r += nwarps) {
        const uint64_t s0 = off[r], L = off[r + 1] - s0, d0 = off2[r], L2 = off2[r + 1] - d0;
        for (uint64_t i = lane; i < L2; i += 32) bases2[d0 + i] = bases[s0 + (i < L ? i : i - L)];
    }
}

// ------------------------------------------------------------------ window minimum
// Deque-free sliding-window minimum (block decomposition): the stream is cut
// into blocks of w; P = running prefix minimum of the current block, S[j] =
// suffix minimum of the previous block from slot j; the minimum of the window
// ending at the current element is min(S[j+1], P).  Leftmost on ties: P takes
// only strictly smaller values, S takes smaller-or-equal ones walking left, and
// S (older) beats P on equality.  State: one ring of w (value, index) slots per
// thread in shared memory, laid out [slot][thread] (conflict-free).
struct WinMin {
    uint64_t *rv; // ring values, this thread's column
    uint16_t *ru; // ring stream indices
    uint32_t stride;
    int w, j;
    uint64_t Pv;
    uint32_t Pu;
    __device__ __forceinline__ void init(uint64_t *v, uint16_t *u, uint32_t st, int w_) {
        rv = v; ru = u; stride = st; w = w_; j = 0; Pv = 0; Pu = 0;
    }
    // Push element (h, u).  Returns true when a full window ends here; (mv, mu) = its leftmost minimum.
    __device__ __forceinline__ bool push(uint64_t h, uint32_t u, uint64_t &mv, uint32_t &mu) {
        if (j == 0 || h < Pv) { Pv = h; Pu = u; }
        rv[j * stride] = h;
        bool full = u + 1 >= (uint32_t)w;
        if (full) {
            mv = Pv; mu = Pu;
            if (j != w - 1) {
                const uint64_t sv = rv[(j + 1) * stride];
                if (!(Pv < sv)) { mv = sv; mu = ru[(j + 1) * stride]; }
            }
        }
        if (j == w - 1) {
            uint64_t sv = h;
            uint32_t su = u;
            ru[j * stride] = (uint16_t)su;
            for (int jj = w - 2; jj >= 1; jj--) {
                const uint64_t hh = rv[jj * stride];
                if (hh <= sv) { sv = hh; su = u - (uint32_t)(w - 1 - jj); }
                rv[jj * stride] = sv;
                ru[jj * stride] = (uint16_t)su;
            }
            j = 0;
        } else {
            j++;
        }
        return full;
    }
};

// Where one item's emitted elements go.
template <bool DIRECT> struct Sink;
template <> struct Sink<false> { // staged in shared memory, [slot][thread]
    uint64_t *lv;
    uint16_t *lp;
    uint32_t stride, cap, cnt;
    __device__ __forceinline__ void emit(uint64_t v, uint32_t rel) {
        if (cnt < cap) { lv[cnt * stride] = v; lp[cnt * stride] = (uint16_t)rel; }
        cnt++;
    }
};
template <> struct Sink<true> { // straight to the final global position (overflow path)
    uint64_t *gv;
    void *gp; // out_pos array (whole), element index gi + cnt
    uint64_t gi;
    uint32_t pw, posbase, cnt;
    __device__ __forceinline__ void emit(uint64_t v, uint32_t rel) {
        gv[cnt] = v;
        if (gp) store_pos(gp, pw, gi + cnt, posbase + rel);
        cnt++;
    }
};

// NextMinimizer over one item.  sb = the item's bases in shared memory.
template <bool DIRECT>
__device__ __forceinline__ void minimizer_item(const uint8_t *sb, const Item &it, int k, int w,
                                               const ulonglong2 *tabIn, const ulonglong2 *tabOut, WinMin &wm,
                                               Sink<DIRECT> &sink) {
    uint64_t fh = 0, rh = 0;
    for (int j = 0; j < k - 1; j++) {
        const ulonglong2 e = tabIn[sb[j]];
        fh = rol1(fh) ^ e.x;
        rh = ror1(rh) ^ e.y;
    }
    uint32_t prev = 0xffffffffu;
    const uint32_t first_own = it.p0 - it.q0; // window index (relative) of the first owned window
    for (uint32_t u = 0; u < it.nstep; u++) {
        const ulonglong2 e = tabIn[sb[u + k - 1]];
        ulonglong2 o = make_ulonglong2(0, 0);
        if (u) o = tabOut[sb[u - 1]];
        fh = rol1(fh) ^ o.x ^ e.x;
        rh = ror1(rh) ^ o.y ^ e.y;
        const uint64_t h = rh < fh ? rh : fh; // canonical (sketch.go:212)
        uint64_t mv;
        uint32_t mu;
        if (wm.push(h, u, mv, mu)) {
            const uint32_t win = u + 1 - (uint32_t)w;
            if (mu != prev && win >= first_own) sink.emit(mv, mu);
            prev = mu;
        }
    }
}

// ProteinMinimizerSketch.Next (sketch-protein.go:106-210) over one item: sb = the item's AMINO ACIDS (the
// frame was translated by k_translate); every window of w consecutive wyhash(k residues, seed 1) values.
template <bool DIRECT>
__device__ __forceinline__ void protein_minimizer_item(const uint8_t *sb, const Item &it, int k, int w, WinMin &wm,
                                                       Sink<DIRECT> &sink) {
    uint32_t prev = 0xffffffffu;
    const uint32_t first_own = it.p0 - it.q0;
    for (uint32_t u = 0; u < it.nstep; u++) {
        ByteSrc src;
        src.p = sb + u;
        const uint64_t h = wyhash_dev(src, (uint32_t)k, 1ull); // sketch-protein.go:118
        uint64_t mv;
        uint32_t mu;
        if (wm.push(h, u, mv, mu)) {
            const uint32_t win = u + 1 - (uint32_t)w;
            if (mu != prev && win >= first_own) sink.emit(mv, mu);
            prev = mu;
        }
    }
}

// NextSyncmer over one item (bounded closed syncmers; s < k here, s == k is the dense path).
template <bool DIRECT>
__device__ __forceinline__ void syncmer_item(const uint8_t *sb, const Item &it, int k, int s,
                                             const ulonglong2 *tabInS, const ulonglong2 *tabOutS,
                                             const uint64_t *tabInK, const ulonglong2 *tabOutK, WinMin &wm,
                                             uint64_t *kring, uint32_t stride, Sink<DIRECT> &sink) {
    const int d = k - s;
    uint64_t fs = 0, rs = 0, fk = 0, rk = 0;
    for (int j = 0; j < s - 1; j++) {
        const unsigned b = sb[j];
        const ulonglong2 e = tabInS[b];
        fs = rol1(fs) ^ e.x;
        rs = ror1(rs) ^ e.y;
        fk = rol1(fk) ^ e.x;
        rk = ror1(rk) ^ tabInK[b];
    }
    uint32_t prev = 0xffffffffu;
    const uint32_t first_own = it.p0 - it.q0;
    int kslot = 0;
    for (uint32_t u = 0; u < it.nstep; u++) {
        const unsigned b = sb[u + s - 1];
        const ulonglong2 e = tabInS[b];
        ulonglong2 os = make_ulonglong2(0, 0), ok = make_ulonglong2(0, 0);
        if (u) os = tabOutS[sb[u - 1]];
        if (u > (uint32_t)d) ok = tabOutK[sb[u - d - 1]];
        fs = rol1(fs) ^ os.x ^ e.x;
        rs = ror1(rs) ^ os.y ^ e.y;
        fk = rol1(fk) ^ ok.x ^ e.x;
        rk = ror1(rk) ^ ok.y ^ tabInK[b];
        if (u >= (uint32_t)d) { // k-mer at stream position u-d is complete
            kring[kslot * stride] = rk < fk ? rk : fk;
            kslot = kslot + 1 == d ? 0 : kslot + 1;
        }
        const uint64_t hs = rs < fs ? rs : fs;
        uint64_t mv;
        uint32_t mu;
        if (wm.push(hs, u, mv, mu)) {
            const uint32_t idx = u + 1 - 2u * (uint32_t)d;          // window start (relative)
            const uint32_t b_rel = (mu - idx < (uint32_t)d) ? mu : mu - (uint32_t)d; // sketch.go:414-420
            if (b_rel != prev && idx >= first_own && it.q0 + b_rel <= it.end) {
                // k-mer b_rel lives in ring slot (b_rel mod d); kslot is the slot of k-mer (u-d+1)
                int slot = kslot - (int)(u - (uint32_t)d + 1 - b_rel);
                if (slot < 0) slot += d;
                sink.emit(kring[slot * stride], b_rel);
            }
            prev = b_rel;
        }
    }
}

// ------------------------------------------------------------------ sparse kernel (generic window size)
template <int MODE>
__global__ void __launch_bounds__(128) k_sparse(const KArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, T = blockDim.x;
    // tables: [0,4K) in, [4K,8K) out; syncmer adds [8K,12K) outK, [12K,14K) inK
    ulonglong2 *tabIn = reinterpret_cast<ulonglong2 *>(smem);
    ulonglong2 *tabOut = tabIn + 256;
    ulonglong2 *tabOutK = tabOut + 256;
    uint64_t *tabInK = reinterpret_cast<uint64_t *>(tabOutK + 256);
    TileCtl *ctl = reinterpret_cast<TileCtl *>(smem + 14336);
    uint8_t *tilebuf = smem + a.sm_tile;
    uint64_t *ringv = reinterpret_cast<uint64_t *>(smem + a.sm_ring);
    uint64_t *listv = reinterpret_cast<uint64_t *>(smem + a.sm_listv);
    uint16_t *listp = reinterpret_cast<uint16_t *>(smem + a.sm_listp);

    const int hk = MODE == B200SK_MODE_SYNCMER ? a.s : a.k; // size of the streamed hash
    const int ww = MODE == B200SK_MODE_SYNCMER ? 2 * (a.k - a.s) : a.w;
    for (uint32_t b = tid; b < 256; b += T) {
        const uint64_t f = fwd_seed(b), r = rev_seed(b);
        tabIn[b] = make_ulonglong2(f, rol64(r, (unsigned)(hk - 1)));
        tabOut[b] = make_ulonglong2(rol64(f, (unsigned)hk), ror64(r, 1));
        if (MODE == B200SK_MODE_SYNCMER) {
            tabOutK[b] = make_ulonglong2(rol64(f, (unsigned)a.k), ror64(r, 1));
            tabInK[b] = rol64(r, (unsigned)(a.k - 1));
        }
    }
    if (tid == 0) {
        mbar_init(&ctl->mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    // ring layout: values [ww][T] u64, then k-mer ring [d][T] u64 (syncmer), then indices [ww][T] u16
    const int d = MODE == B200SK_MODE_SYNCMER ? a.k - a.s : 0;
    uint64_t *kring = ringv + (size_t)ww * T;
    uint16_t *ringu = reinterpret_cast<uint16_t *>(kring + (size_t)d * T);

    const uint64_t n_items = a.n_items_dev ? *a.n_items_dev : a.n_items;
    uint32_t parity = 0;
    for (;;) {
        if (tid == 0) ctl->tile = atomicAdd(a.ticket, 1ULL);
        __syncthreads();
        const uint64_t tile = ctl->tile;
        const uint64_t item0 = tile * T;
        if (item0 >= n_items) break;
        const uint32_t nvalid = (uint32_t)min((uint64_t)T, n_items - item0);
        Item it;
        item_geometry<MODE>(a, item0 + tid, n_items, it);
        if (tid == 0) { ctl->lo = it.gb0; ctl->any_overflow = 0; }
        if (tid == nvalid - 1) ctl->hi = it.gb0 + it.nb;
        __syncthreads();
        const uint64_t lo_al = ctl->lo & ~15ULL;
        const uint64_t span = ctl->hi > lo_al ? ctl->hi - lo_al : 0;
        const uint32_t bytes = (uint32_t)((span + 15ULL) & ~15ULL);
        const bool span_ok = bytes <= a.sm_tile_bytes;
        if (tid == 0 && bytes && span_ok) {
            fence_proxy_async(); // earlier generic-proxy writes to this buffer (ordered copy) vs the async write
            mbar_expect_tx(&ctl->mbar, bytes);
            tma_load_1d(tilebuf, a.bases + lo_al, bytes, &ctl->mbar);
        }
        if (!span_ok && tid == 0) atomicOr(a.flags, B200SK_FLAG_SPAN);
        // per-read bookkeeping that does not need the bases
        if (it.valid && it.first_chunk && a.status) a.status[it.r] = it.status;
        if (bytes && span_ok) {
            mbar_wait(&ctl->mbar, parity);
            parity ^= 1u;
        }
        Sink<false> sink;
        sink.lv = listv + tid; sink.lp = listp + tid; sink.stride = T; sink.cap = a.lcap; sink.cnt = 0;
        const uint8_t *sb = tilebuf + (uint32_t)(it.gb0 - lo_al);
        if (it.valid && it.nstep && span_ok) {
            WinMin wm;
            wm.init(ringv + tid, ringu + tid, T, ww);
            if (MODE == B200SK_MODE_MINIMIZER)
                minimizer_item<false>(sb, it, a.k, a.w, tabIn, tabOut, wm, sink);
            else if (MODE == B200SK_MODE_PROTEIN_MINIMIZER)
                protein_minimizer_item<false>(sb, it, a.k, a.w, wm, sink);
            else
                syncmer_item<false>(sb, it, a.k, a.s, tabIn, tabOut, tabInK, tabOutK, wm, kring + tid, T, sink);
        }
        const uint32_t cnt = sink.cnt;
        const bool overflow = cnt > a.lcap;
        if (overflow) ctl->any_overflow = 1;
        uint32_t total;
        const uint32_t excl = block_excl_scan(cnt, ctl->warp_sums, &total);
        if (tid < 32) {
            const uint64_t b = lookback_exclusive(a.tile_state, tile, total);
            if (tid == 0) ctl->base = b;
        }
        __syncthreads();
        const uint64_t tb = ctl->base;
        const uint64_t mine = tb + excl;
        if (it.valid && it.first_chunk) a.out_off[it.r] = a.out_base + mine;
        if (it.valid && it.last_item) a.out_off[a.n_reads] = a.out_base + mine + cnt;
        const bool fits = tb + total <= a.capacity;
        if (!fits && tid == 0) atomicOr(a.flags, B200SK_FLAG_CAPACITY);
        if (fits && total) {
            // ordered copy: scatter the staged lists into one contiguous buffer (the bases and the
            // rings are dead now), then stream it out coalesced.
            const uint32_t ob_bytes = a.sm_listv - a.sm_tile;
            const uint32_t OB = ob_bytes / 12u;
            uint64_t *obv = reinterpret_cast<uint64_t *>(tilebuf);
            uint32_t *obp = reinterpret_cast<uint32_t *>(tilebuf + (size_t)OB * 8u);
            const uint32_t ncopy = overflow ? 0u : cnt;
            for (uint32_t r0 = 0; r0 < total; r0 += OB) {
                for (uint32_t j = 0; j < ncopy; j++) {
                    const uint32_t o = excl + j - r0;
                    if (o < OB) { // unsigned: also rejects excl + j < r0
                        obv[o] = listv[j * T + tid];
                        obp[o] = it.q0 + listp[j * T + tid];
                    }
                }
                __syncthreads();
                const uint32_t n = min(OB, total - r0);
                uint64_t *gv = a.out_val + tb + r0;
                for (uint32_t i = tid; i < n; i += T) gv[i] = obv[i];
                if (a.out_pos)
                    for (uint32_t i = tid; i < n; i += T) store_pos(a.out_pos, a.pos_width, tb + r0 + i, obp[i]);
                __syncthreads();
            }
            if (ctl->any_overflow) {
                // rare: an item emitted more than the staging list holds (low-complexity reads).
                // Re-walk those items writing straight to their final positions.  The tile's bases
                // were overwritten by the ordered copy, so read them from global memory instead.
                if (overflow) {
                    Sink<true> ds;
                    ds.gv = a.out_val + mine; ds.gp = a.out_pos; ds.gi = mine; ds.pw = a.pos_width;
                    ds.posbase = it.q0; ds.cnt = 0;
                    WinMin wm;
                    wm.init(ringv + tid, ringu + tid, T, ww);
                    const uint8_t *gb = a.bases + it.gb0;
                    if (MODE == B200SK_MODE_MINIMIZER)
                        minimizer_item<true>(gb, it, a.k, a.w, tabIn, tabOut, wm, ds);
                    else if (MODE == B200SK_MODE_PROTEIN_MINIMIZER)
                        protein_minimizer_item<true>(gb, it, a.k, a.w, wm, ds);
                    else
                        syncmer_item<true>(gb, it, a.k, a.s, tabIn, tabOut, tabInK, tabOutK, wm, kring + tid, T, ds);
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ launch helpers (host)
cudaError_t launch_prepass(const KArgs &a, unsigned long long *meta, cudaStream_t st) {
    const int threads = 256;
    uint64_t blocks = (a.n_reads + threads - 1) / threads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks == 0) blocks = 1;
    k_prepass<<<(unsigned)blocks, threads, 0, st>>>(a.off, a.off_orig, a.n_reads, a.geom(), a.C, meta);
    return cudaGetLastError();
}

cudaError_t launch_scan_items(const KArgs &a, uint64_t *item_first, uint64_t *tile_state,
                              unsigned long long *ticket, cudaStream_t st, uint64_t *tile_read) {
    uint64_t tiles = (a.n_reads + 1023) / 1024;
    uint64_t blocks = tiles < 148 * 8 ? tiles : 148 * 8;
    if (blocks == 0) blocks = 1;
    k_scan_reads<false><<<(unsigned)blocks, 256, 0, st>>>(a.off, a.off_orig, a.n_reads, a.geom(), a.C, 0, item_first,
                                                         nullptr, tile_state, ticket, tile_read);
    return cudaGetLastError();
}

// generic per-read count scan (used for the amino-acid offsets of a translated frame)
cudaError_t launch_scan_geom(const uint64_t *off, uint64_t n_reads, const ReadGeom &g, uint64_t *out,
                             uint64_t *tile_state, unsigned long long *ticket, cudaStream_t st) {
    uint64_t tiles = (n_reads + 1023) / 1024;
    uint64_t blocks = tiles < 148 * 8 ? tiles : 148 * 8;
    if (blocks == 0) blocks = 1;
    k_scan_reads<true><<<(unsigned)blocks, 256, 0, st>>>(off, nullptr, n_reads, g, 1, 0, out, nullptr, tile_state, ticket);
    return cudaGetLastError();
}

// dense modes: out_off[0..n_reads] (+ out_base) and read_status straight from the lengths
cudaError_t launch_scan_counts(const KArgs &a, uint64_t *tile_state, unsigned long long *ticket, cudaStream_t st) {
    uint64_t tiles = (a.n_reads + 1023) / 1024;
    uint64_t blocks = tiles < 148 * 8 ? tiles : 148 * 8;
    if (blocks == 0) blocks = 1;
    k_scan_reads<true><<<(unsigned)blocks, 256, 0, st>>>(a.off, a.off_orig, a.n_reads, a.geom(), a.C, a.out_base,
                                                        a.out_off, a.status, tile_state, ticket);
    return cudaGetLastError();
}

cudaError_t launch_circularize(const uint8_t *bases, const uint64_t *off, uint64_t n_reads, int k,
                               uint8_t *bases2, uint64_t *off2, uint64_t *tile_state,
                               unsigned long long *ticket, cudaStream_t st) {
    uint64_t tiles = (n_reads + 1023) / 1024;
    uint64_t blocks = tiles < 148 * 8 ? tiles : 148 * 8;
    if (blocks == 0) blocks = 1;
    k_circ_offsets<<<(unsigned)blocks, 256, 0, st>>>(off, n_reads, k, off2, tile_state, ticket);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    uint64_t cb = (n_reads + 7) / 8; // 8 warps per block
    if (cb > 148 * 16) cb = 148 * 16;
    if (cb == 0) cb = 1;
    k_circ_copy<<<(unsigned)cb, 256, 0, st>>>(bases, off, off2, n_reads, bases2);
    return cudaGetLastError();
}

static cudaError_t set_smem(const void *fn, uint32_t bytes) {
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_main(const KArgs &a, int threads, int blocks, cudaStream_t st) {
    cudaError_t e;
    switch (a.mode) {
    case B200SK_MODE_MINIMIZER:
        if ((e = set_smem((const void *)k_sparse<B200SK_MODE_MINIMIZER>, a.sm_total)) != cudaSuccess) return e;
        k_sparse<B200SK_MODE_MINIMIZER><<<blocks, threads, a.sm_total, st>>>(a);
        break;
    case B200SK_MODE_SYNCMER:
        if ((e = set_smem((const void *)k_sparse<B200SK_MODE_SYNCMER>, a.sm_total)) != cudaSuccess) return e;
        k_sparse<B200SK_MODE_SYNCMER><<<blocks, threads, a.sm_total, st>>>(a);
        break;
    case B200SK_MODE_PROTEIN_MINIMIZER:
        if ((e = set_smem((const void *)k_sparse<B200SK_MODE_PROTEIN_MINIMIZER>, a.sm_total)) != cudaSuccess) return e;
        k_sparse<B200SK_MODE_PROTEIN_MINIMIZER><<<blocks, threads, a.sm_total, st>>>(a);
        break;
    default:
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int main_kernel_occupancy(const KArgs &a, int threads) {
    int nb = 0;
    const void *fn = a.mode == B200SK_MODE_MINIMIZER ? (const void *)k_sparse<B200SK_MODE_MINIMIZER>
                     : a.mode == B200SK_MODE_SYNCMER ? (const void *)k_sparse<B200SK_MODE_SYNCMER>
                                                     : (const void *)k_sparse<B200SK_MODE_PROTEIN_MINIMIZER>;
    set_smem(fn, a.sm_total);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, a.sm_total) != cudaSuccess) return 1;
    return nb < 1 ? 1 : nb;
}

} // namespace b200sk
