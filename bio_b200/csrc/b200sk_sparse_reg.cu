// Minimizer / closed-syncmer kernels with the sliding-window state in REGISTERS
// (window size a template parameter, the block loop fully unrolled) -- the fast
// path for window sizes 2..32; larger windows use the generic shared-memory ring
// kernel in b200sk_kernels.cu.
//
// Per tile (blockDim.x consecutive items, see b200sk_kernels.cu):
//   1. one 1-D TMA bulk copy brings the tile's byte range into shared memory;
//   2. a cooperative 16-byte-vector pass rewrites each ASCII byte as a 6-bit
//      CODE that carries exactly what ntHash needs of it (see code_of_byte);
//   3. every thread walks its item: 2 byte loads + 2 LDS.128 table lookups +
//      the two rolling 64-bit hashes + canonical select per base, then the
//      window minimum by block decomposition over registers;
//   4. emitted (value, position-delta) pairs are staged per thread in shared
//      memory, the tile's output range comes from the look-back, and the
//      elements leave in global order through a coalesced copy.
//
// Reference behaviour: sketches/sketch.go:205-309 (NextMinimizer), :312-477
// (NextSyncmer), ntHash via will-rowe/nthash v0.4.0 (sketch.go:212,319,367).
#include "b200sk_tile.cuh"

namespace b200sk {

// ------------------------------------------------------------------ shared-memory access
// All hot-loop accesses index the kernel's dynamic shared array with 32-bit offsets, so the compiler
// emits LDS/STS with register+immediate addressing (unrolled steps fold their offsets).
__device__ __forceinline__ uint32_t lds_u8(const uint8_t *sm, uint32_t o) { return sm[o]; }
__device__ __forceinline__ ulonglong2 lds_v2u64(const uint8_t *sm, uint32_t o) {
    return *reinterpret_cast<const ulonglong2 *>(sm + o);
}
__device__ __forceinline__ void sts_u64(uint8_t *sm, uint32_t o, uint64_t v) {
    *reinterpret_cast<uint64_t *>(sm + o) = v;
}
__device__ __forceinline__ void sts_u8(uint8_t *sm, uint32_t o, uint32_t v) { sm[o] = (uint8_t)v; }

// ------------------------------------------------------------------ 6-bit codes
// ntHash looks at a base twice: forward seed = seedTab[b], reverse-complement seed = seedTab[b & 7].
// Bytes in 0x40..0x7f (all letters) are represented by b & 31 (case folds, b & 7 is preserved);
// the five low bytes with a non-zero forward seed keep their own entries; every other byte has a
// zero forward seed and maps to the letter 'H'..'O' that shares its low three bits.  The mapping is
// exact for all 256 byte values, so there is no slow path for unusual input.
__device__ __host__ __forceinline__ uint32_t code_of_byte(uint32_t b) {
    if ((b & 0xC0u) == 0x40u) return b & 31u;
    if (b == 1 || b == 3 || b == 4 || b == 5 || b == 7) return 32u + b;
    return 8u | (b & 7u);
}
__device__ __host__ __forceinline__ uint32_t byte_of_code(uint32_t c) { return c < 32u ? (0x40u | c) : c - 32u; }

__device__ __forceinline__ uint32_t codes_of_word(uint32_t w) {
    return code_of_byte(w & 0xff) | (code_of_byte((w >> 8) & 0xff) << 8) | (code_of_byte((w >> 16) & 0xff) << 16) |
           (code_of_byte(w >> 24) << 24);
}

// ------------------------------------------------------------------ compare-select blocks
// One setp feeds every select of a (value, position) update: written as PTX blocks so that the 64-bit
// compare is materialised once (two ISETP) instead of once per polarity.
// The 64-bit "a < b" is computed ONCE as an all-ones/zero mask through the borrow chain
// (sub.cc / subc), and every select of the update is a single LOP3 bit-mux on that mask.  (With
// setp + selp ptxas re-evaluates the two-instruction 64-bit compare for each polarity it needs.)
__device__ __forceinline__ uint32_t lt_mask(uint64_t a, uint64_t b) { // a < b ? 0xffffffff : 0
    uint32_t m;
    asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %3;\n\tsubc.cc.u32 t, %2, %4;\n\tsubc.u32 %0, 0, 0;\n\t}"
        : "=r"(m)
        : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
    return m;
}
__device__ __forceinline__ uint32_t mux32(uint32_t m, uint32_t a, uint32_t b) { return m ? a : b; } // m ? a : b
__device__ __forceinline__ uint64_t mux64(uint32_t m, uint64_t a, uint64_t b) {
    const uint32_t lo = mux32(m, (uint32_t)a, (uint32_t)b), hi = mux32(m, (uint32_t)(a >> 32), (uint32_t)(b >> 32));
    return ((uint64_t)hi << 32) | lo;
}
// if (h < v) { v = h; p = c; }
__device__ __forceinline__ void take_if_less(uint64_t &v, uint32_t &p, uint64_t h, uint32_t c) {
    const uint32_t m = lt_mask(h, v);
    v = mux64(m, h, v);
    p = mux32(m, c, p);
}
// (mv, mu) = (pv < sv) ? (pv, pu) : (sv, su)      -- the older element (S) wins ties
__device__ __forceinline__ void pick_min(uint64_t &mv, uint32_t &mu, uint64_t pv, uint32_t pu, uint64_t sv,
                                         uint32_t su) {
    const uint32_t m = lt_mask(pv, sv);
    mv = mux64(m, pv, sv);
    mu = mux32(m, pu, su);
}
// suffix minimum step: (h, ph) <- min((h, c), (v, pv)), the left element (h) wins ties
__device__ __forceinline__ void suffix_step(uint64_t &h, uint32_t &ph, uint32_t c, uint64_t v, uint32_t pv) {
    const uint32_t m = lt_mask(v, h);
    h = mux64(m, v, h);
    ph = mux32(m, pv, c);
}
__device__ __forceinline__ uint64_t min_u64(uint64_t a, uint64_t b) { return mux64(lt_mask(a, b), a, b); }

// ------------------------------------------------------------------ sinks
// emit(p, v, delta): record (v, delta) when p.  The slot address and the operands are computed
// unconditionally; only the two stores and the counter depend on p, so the compiler predicates them
// instead of branching (some lane of a warp emits on almost every step).
struct ListSink { // staged in shared memory, [slot][thread]; slot `cap` is a scratch slot for overflow
    uint32_t av, ap, sv, sp, cap, cnt; // av/ap: 32-bit shared-window addresses of this thread's slot 0
    __device__ __forceinline__ void emit(uint32_t mu, uint32_t prev, uint64_t v) {
        const uint32_t slot = min(cnt, cap);
        const uint32_t ov = av + slot * sv, op = ap + slot * sp;
        const uint32_t delta = mu - prev;
        // predicated stores (no branch): record iff the minimum moved
        asm volatile(
            "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, %2;\n\t@q st.shared.u64 [%3], %4;\n\t"
            "@q st.shared.u8 [%5], %6;\n\t@q add.u32 %0, %0, 1;\n\t}"
            : "+r"(cnt)
            : "r"(mu), "r"(prev), "r"(ov), "l"(v), "r"(op), "r"(delta));
    }
};
struct GlobalSink { // straight to the final position (tiles where some item overflowed its list)
    uint64_t *gv;
    uint32_t *gp;
    uint32_t pos, cnt, skip;
    __device__ __forceinline__ void emit(uint32_t mu, uint32_t prev, uint64_t v) {
        if (mu != prev) {
            pos += mu - prev;
            if (cnt >= skip) {
                gv[cnt - skip] = v;
                if (gp) gp[cnt - skip] = pos;
            }
            cnt++;
        }
    }
};

// ------------------------------------------------------------------ hashing step
struct Roll {
    uint64_t f, r;
    __device__ __forceinline__ void fold(const ulonglong2 e) {
        f = rol1(f) ^ e.x;
        r = ror1(r) ^ e.y;
    }
    __device__ __forceinline__ void roll(const ulonglong2 in, const ulonglong2 out) {
        f = rol1(f) ^ out.x ^ in.x;
        r = ror1(r) ^ out.y ^ in.y;
    }
    __device__ __forceinline__ uint64_t canonical() const { return min_u64(r, f); }
};

// ------------------------------------------------------------------ window minimum in registers
// Block decomposition over blocks of W stream elements.  hs[j] holds the current block's element j
// until the block ends, then the suffix minimum S[j] of that block; sp[j] = block-relative index of
// S[j].  Positions are tracked relative to the start of the PREVIOUS block (frame), so all selects use
// compile-time constants: S positions are 0..W-1, P positions W..2W-1.
template <int W> struct WinReg {
    uint64_t hs[W];
    uint32_t sp[W];
    uint64_t pv;
    uint32_t pj;
    uint32_t prev;  // frame-relative position of the previous window's minimum
    uint32_t wbase; // == W, but a run-time value: position constants are formed as wbase + J in a register
                    // (IMAD, fma pipe) so that both selects of an update share ONE predicate; with an
                    // immediate operand ptxas re-evaluates the 64-bit compare for the other polarity

    __device__ __forceinline__ void init(uint32_t w_runtime) { prev = W - 1; pv = 0; pj = 0; wbase = w_runtime; }

    // element j of the current block.  FIRST: block 0 (no previous block: only the last element closes a window)
    template <class SinkT> __device__ __forceinline__ void push(const int J, const bool FIRST, uint64_t h, SinkT &sink) {
        if (J == 0) { pv = h; pj = W; }
        else take_if_less(pv, pj, h, wbase + (uint32_t)J); // strictly less: the leftmost stays on ties
        if (!FIRST || J == W - 1) {
            uint64_t mv = pv;
            uint32_t mu = pj;
            if (J != W - 1) pick_min(mv, mu, pv, pj, hs[J + 1], sp[J + 1]);
            sink.emit(mu, prev, mv);
            prev = mu;
        }
        hs[J] = h;
    }
    // after element W-1: turn hs[] into suffix minima, shift the frame by one block
    __device__ __forceinline__ void close_block() {
        sp[W - 1] = W - 1;
#pragma unroll
        for (int jj = W - 2; jj >= 1; jj--)
            suffix_step(hs[jj], sp[jj], wbase - (uint32_t)(W - jj), hs[jj + 1], sp[jj + 1]);
        prev -= W;
    }
};

// NextMinimizer over one item; its codes start at shared offset sb; tabIn/tabOut = table offsets.
template <int W, class SinkT>
__device__ __forceinline__ void minimizer_item_reg(const uint8_t *sm, uint32_t sb, uint32_t nstep, int k,
                                                   uint32_t w_runtime, uint32_t tabIn, uint32_t tabOut,
                                                   SinkT &sink) {
#define B200SK_IN(o) lds_v2u64(sm, tabIn + lds_u8(sm, (o)) * 16u)
#define B200SK_OUT(o) lds_v2u64(sm, tabOut + lds_u8(sm, (o)) * 16u)
    Roll h;
    h.f = 0; h.r = 0;
    for (int j = 0; j < k - 1; j++) h.fold(B200SK_IN(sb + j));
    WinReg<W> wm;
    wm.init(w_runtime);
    uint32_t pin = sb + (uint32_t)k - 1; // next incoming code
    uint32_t pout = sb - 1;              // next outgoing code is pout + 1 ... (block 0 starts one step late)
    // block 0: the first k-mer has no outgoing base
    h.fold(B200SK_IN(pin));
    wm.push(0, true, h.canonical(), sink);
#pragma unroll
    for (int j = 1; j < W; j++) {
        h.roll(B200SK_IN(pin + j), B200SK_OUT(pout + j));
        wm.push(j, true, h.canonical(), sink);
    }
    wm.close_block();
    pin += W; pout += W;
    uint32_t u0 = W;
    // full blocks
    while (u0 + W <= nstep) {
#pragma unroll
        for (int j = 0; j < W; j++) {
            h.roll(B200SK_IN(pin + j), B200SK_OUT(pout + j));
            wm.push(j, false, h.canonical(), sink);
        }
        wm.close_block();
        pin += W; pout += W;
        u0 += W;
    }
    // tail: fewer than W elements left
    const uint32_t rem = nstep - u0;
#pragma unroll
    for (int j = 0; j < W - 1; j++) {
        if ((uint32_t)j >= rem) break;
        h.roll(B200SK_IN(pin + j), B200SK_OUT(pout + j));
        wm.push(j, false, h.canonical(), sink);
    }
#undef B200SK_IN
#undef B200SK_OUT
}

// ------------------------------------------------------------------ kernel
template <int W>
__global__ void __launch_bounds__(128, 4) k_minimizer_reg(const KArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, T = blockDim.x;
    // [0,1K) in-table, [1K,2K) out-table (64 codes x 16 B), then TileCtl, then tile / lists
    ulonglong2 *tIn = reinterpret_cast<ulonglong2 *>(smem);
    ulonglong2 *tOut = tIn + 64;
    TileCtl *ctl = reinterpret_cast<TileCtl *>(smem + 2048);
    uint8_t *tilebuf = smem + a.sm_tile;
    uint64_t *listv = reinterpret_cast<uint64_t *>(smem + a.sm_listv);
    uint8_t *listp = smem + a.sm_listp;
    for (uint32_t c = tid; c < 64; c += T) {
        const uint32_t b = byte_of_code(c);
        const uint64_t f = fwd_seed(b), r = rev_seed(b);
        tIn[c] = make_ulonglong2(f, rol64(r, (unsigned)(a.k - 1)));
        tOut[c] = make_ulonglong2(rol64(f, (unsigned)a.k), ror64(r, 1));
    }
    if (tid == 0) {
        mbar_init(&ctl->mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t s_tIn = 0, s_tOut = 1024, s_tile = a.sm_tile;
    const uint32_t s_lv = a.sm_listv + tid * 8u, s_lp = a.sm_listp + tid;
    const uint32_t smem_base = smem_u32(smem);
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
        item_geometry<B200SK_MODE_MINIMIZER>(a, item0 + tid, n_items, it);
        if (tid == 0) { ctl->lo = it.gb0; ctl->any_overflow = 0; }
        if (tid == nvalid - 1) ctl->hi = it.gb0 + it.nb;
        __syncthreads();
        const uint64_t lo_al = ctl->lo & ~15ULL;
        const uint64_t span = ctl->hi > lo_al ? ctl->hi - lo_al : 0;
        const uint32_t bytes = (uint32_t)((span + 15ULL) & ~15ULL);
        const bool span_ok = bytes <= a.sm_tile_bytes;
        if (tid == 0 && bytes && span_ok) {
            fence_proxy_async(); // the previous tile's generic-proxy writes to this buffer precede the async write
            mbar_expect_tx(&ctl->mbar, bytes);
            tma_load_1d(tilebuf, a.bases + lo_al, bytes, &ctl->mbar);
        }
        if (!span_ok && tid == 0) atomicOr(a.flags, B200SK_FLAG_SPAN);
        if (it.valid && it.first_chunk && a.status) a.status[it.r] = it.status;
        if (bytes && span_ok) {
            mbar_wait(&ctl->mbar, parity);
            parity ^= 1u;
            // ASCII -> codes, 16 bytes per thread per trip
            for (uint32_t o = tid * 16u; o < bytes; o += T * 16u) {
                uint4 v = *reinterpret_cast<uint4 *>(tilebuf + o);
                const uint32_t orr = v.x | v.y | v.z | v.w, andd = v.x & v.y & v.z & v.w;
                if ((orr & 0x80808080u) == 0 && (andd & 0x40404040u) == 0x40404040u) {
                    v.x &= 0x1f1f1f1fu; v.y &= 0x1f1f1f1fu; v.z &= 0x1f1f1f1fu; v.w &= 0x1f1f1f1fu;
                } else {
                    v.x = codes_of_word(v.x); v.y = codes_of_word(v.y);
                    v.z = codes_of_word(v.z); v.w = codes_of_word(v.w);
                }
                *reinterpret_cast<uint4 *>(tilebuf + o) = v;
            }
        }
        __syncthreads();
        ListSink sink;
        sink.av = smem_base + s_lv; sink.ap = smem_base + s_lp; sink.sv = T * 8u; sink.sp = T; sink.cap = a.lcap; sink.cnt = 0;
        const uint32_t sb = s_tile + (uint32_t)(it.gb0 - lo_al);
        const bool run = it.valid && it.nstep && span_ok;
        if (run) minimizer_item_reg<W>(smem, sb, it.nstep, a.k, (uint32_t)a.w, s_tIn, s_tOut, sink);
        // a non-first chunk walks one window more (the one before its first own window) to seed the
        // de-duplication; that window always emits first and is dropped here
        const uint32_t skip = (run && it.q0 != it.p0) ? 1u : 0u;
        const uint32_t cnt = sink.cnt - skip;
        const bool overflow = sink.cnt > a.lcap;
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
            if (!ctl->any_overflow) {
                // ordered copy: scatter the staged lists into one contiguous buffer (the codes are dead
                // now), then stream it out coalesced
                const uint32_t OB = (a.sm_listv - a.sm_tile) / 12u;
                uint64_t *obv = reinterpret_cast<uint64_t *>(tilebuf);
                uint32_t *obp = reinterpret_cast<uint32_t *>(tilebuf + (size_t)OB * 8u);
                for (uint32_t r0 = 0; r0 < total; r0 += OB) {
                    __syncthreads();
                    uint32_t pos = it.q0 - 1u;
                    for (uint32_t j = 0; j < sink.cnt; j++) {
                        pos += listp[j * T + tid];
                        const uint32_t o = excl + j - skip - r0;
                        if (j >= skip && o < OB) { // unsigned compare also rejects entries before r0
                            obv[o] = listv[j * T + tid];
                            obp[o] = pos;
                        }
                    }
                    __syncthreads();
                    const uint32_t n = min(OB, total - r0);
                    uint64_t *gv = a.out_val + tb + r0;
                    for (uint32_t i = tid; i < n; i += T) gv[i] = obv[i];
                    if (a.out_pos) {
                        uint32_t *gp = a.out_pos + tb + r0;
                        for (uint32_t i = tid; i < n; i += T) gp[i] = obp[i];
                    }
                }
            } else {
                // rare (low-complexity reads): some item emitted more than its list holds.  Items that fit
                // write their lists straight to their final range; the others walk their item again with
                // the global sink (the codes are still in shared memory).
                if (!overflow) {
                    uint32_t pos = it.q0 - 1u;
                    for (uint32_t j = 0; j < sink.cnt; j++) {
                        pos += listp[j * T + tid];
                        if (j >= skip) {
                            a.out_val[mine + j - skip] = listv[j * T + tid];
                            if (a.out_pos) a.out_pos[mine + j - skip] = pos;
                        }
                    }
                } else {
                    GlobalSink gs;
                    gs.gv = a.out_val + mine; gs.gp = a.out_pos ? a.out_pos + mine : nullptr;
                    gs.pos = it.q0 - 1u; gs.cnt = 0; gs.skip = skip;
                    minimizer_item_reg<W>(smem, sb, it.nstep, a.k, (uint32_t)a.w, s_tIn, s_tOut, gs);
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ launch
template <int W> static cudaError_t launch_w(const KArgs &a, int threads, int blocks, cudaStream_t st, int *occ) {
    const void *fn = (const void *)k_minimizer_reg<W>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.sm_total);
    if (e != cudaSuccess) return e;
    if (occ) {
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, a.sm_total);
        *occ = nb < 1 ? 1 : nb;
        return e;
    }
    k_minimizer_reg<W><<<blocks, threads, a.sm_total, st>>>(a);
    return cudaGetLastError();
}

// Window sizes with a register-resident instantiation.  B200SK_FAST_BUILD (development) keeps a handful.
#ifdef B200SK_FAST_BUILD
#define B200SK_WLIST B200SK_W(3) B200SK_W(5) B200SK_W(11) B200SK_W(15) B200SK_W(20)
#else
#define B200SK_WLIST                                                                                         \
    B200SK_W(2) B200SK_W(3) B200SK_W(4) B200SK_W(5) B200SK_W(6) B200SK_W(7) B200SK_W(8) B200SK_W(9)          \
    B200SK_W(10) B200SK_W(11) B200SK_W(12) B200SK_W(13) B200SK_W(14) B200SK_W(15) B200SK_W(16) B200SK_W(17)  \
    B200SK_W(18) B200SK_W(19) B200SK_W(20) B200SK_W(21) B200SK_W(22) B200SK_W(23) B200SK_W(24)
#endif
// occ != nullptr: only report the occupancy
cudaError_t launch_minimizer_reg(const KArgs &a, int threads, int blocks, cudaStream_t st, int *occ) {
    switch (a.w) {
#define B200SK_W(W) case W: return launch_w<W>(a, threads, blocks, st, occ);
        B200SK_WLIST
#undef B200SK_W
    default: return cudaErrorInvalidValue;
    }
}

bool minimizer_reg_supported(int w) {
    switch (w) {
#define B200SK_W(W) case W: return true;
        B200SK_WLIST
#undef B200SK_W
    default: return false;
    }
}

} // namespace b200sk
