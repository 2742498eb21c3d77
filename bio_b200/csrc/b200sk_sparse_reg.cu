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
#include "b200sk_protein.cuh"
#include "b200sk_tile.cuh"

// B200SK_EXPERIMENTS (off): the two timing experiments of DESIGN.md 5.1 -- output ranges handed out in completion order
// (B200SK_UNORDERED=1, wrong order, same work) and a sleeping look-back poll (B200SK_SPIN_NS).  Compiled out of the
// product: their run-time branches alone cost the headline kernel 1 %.
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
// emit_if(p, v, delta): record (v, delta) when p.  The slot address and the operands are computed
// unconditionally; only the two stores and the counter depend on p and they are PREDICATED, not branched
// around (some lane of a warp emits on almost every step).
struct ListSink { // staged in shared memory, [slot][lane]; slot `cap` is a scratch slot for overflow
    uint32_t av, ap, cap, cnt; // av/ap: 32-bit shared-window addresses of this lane's slot 0
    __device__ __forceinline__ void emit_if(uint32_t pred, uint64_t v, uint32_t delta) {
        const uint32_t slot = min(cnt, cap);
        const uint32_t ov = av + slot * 256u, op = ap + slot * 32u; // 32 lanes x 8 B / x 1 B per slot
        asm volatile(
            "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q st.shared.u64 [%2], %3;\n\t"
            "@q st.shared.u8 [%4], %5;\n\t@q add.u32 %0, %0, 1;\n\t}"
            : "+r"(cnt)
            : "r"(pred), "r"(ov), "l"(v), "r"(op), "r"(delta));
    }
};
struct GlobalSink { // straight to the final position (tiles where some item overflowed its list)
    uint64_t *gv;
    void *gp; // out_pos array (whole), element index gi + n
    uint64_t gi;
    uint32_t pw, pos, cnt, skip;
    __device__ __forceinline__ void emit_if(uint32_t pred, uint64_t v, uint32_t delta) {
        if (pred) {
            pos += delta;
            if (cnt >= skip) {
                gv[cnt - skip] = v;
                if (gp) store_pos(gp, pw, gi + cnt - skip, pos);
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
// S[j].  Positions are tracked relative to the start of the PREVIOUS block (the frame), so the position
// operands are constants: S positions are 0..W-1, P positions W..2W-1.
template <int W> struct WinReg {
    uint64_t hs[W];
    uint32_t sp[W];
    uint64_t pv;
    uint32_t pj;
    uint32_t wbase; // == W, but a run-time value: position constants are formed as wbase + J in a register
                    // (IMAD, fma pipe); with an immediate operand ptxas re-evaluates the 64-bit compare for
                    // the other predicate polarity

    __device__ __forceinline__ void init(uint32_t w_runtime) { pv = 0; pj = 0; wbase = w_runtime; }

    // element J of the current block.  FIRST: block 0 (no previous block: only the last element closes a
    // window).  Returns true (compile-time) when a window closes here; (mv, mu) = its leftmost minimum.
    __device__ __forceinline__ bool push(const int J, const bool FIRST, uint64_t h, uint64_t &mv, uint32_t &mu) {
        if (J == 0) { pv = h; pj = W; }
        else take_if_less(pv, pj, h, wbase + (uint32_t)J); // strictly less: the leftmost stays on ties
        const bool closes = !FIRST || J == W - 1;
        if (closes) {
            mv = pv;
            mu = pj;
            if (J != W - 1) pick_min(mv, mu, pv, pj, hs[J + 1], sp[J + 1]);
        }
        hs[J] = h;
        return closes;
    }
    // after element W-1: turn hs[] into suffix minima (the frame then moves on by one block)
    __device__ __forceinline__ void close_block() {
        sp[W - 1] = W - 1;
#pragma unroll
        for (int jj = W - 2; jj >= 1; jj--)
            suffix_step(hs[jj], sp[jj], wbase - (uint32_t)(W - jj), hs[jj + 1], sp[jj + 1]);
    }
};

#define B200SK_LD128(tab, o) lds_v2u64(sm, (tab) + lds_u8(sm, (o)) * 16u)

// The codes of one block of W steps, fetched as aligned 32-bit words (a quarter of the shared-memory
// wavefronts of per-step byte loads, and a quarter of the cost of any bank conflict between lanes), then
// lined up with one PRMT per 4 codes and picked out with one PRMT per code.
template <int W> struct CodeWords {
    static constexpr int NWORD = (W + 3 + 3) / 4; // words that cover bytes phi .. phi+W-1 for phi <= 3
    static constexpr int NG = (W + 3) / 4;        // groups of 4 consecutive codes
    uint32_t x[NG];
    __device__ __forceinline__ void load(const uint8_t *sm, uint32_t p) { // p: shared offset of step 0's code
        const uint32_t a = p & ~3u;
        const uint32_t sel = 0x3210u + 0x1111u * (p & 3u);
        uint32_t w[NWORD + 1];
#pragma unroll
        for (int i = 0; i < NWORD; i++) w[i] = *reinterpret_cast<const uint32_t *>(sm + a + 4u * i);
        w[NWORD] = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) x[g] = __byte_perm(w[g], w[g + 1 < NWORD ? g + 1 : NWORD], sel);
    }
    __device__ __forceinline__ uint32_t code(const int j) const { return __byte_perm(x[j >> 2], 0u, 0x4440u | (j & 3)); }
};
#define B200SK_TAB(tab, c) lds_v2u64(sm, (tab) + (c) * 16u)

// ------------------------------------------------------------------ all-ACGT fast path
// When every byte of a tile is one of ACGTacgt the tile is rewritten as cls * 40 = (cls << 3) | (cls << 5)
// with cls = (byte >> 1) & 3 (a=0 c=1 t=2 g=3).  (in & 0x18) | (out & 0x60) is then directly the byte
// offset into 16-entry PAIR tables of 8-byte entries,
//     X[in, out] = A[in] ^ rol(A[out], h),   Y[in, out] = rol(B[in], h-1) ^ ror(B[out], 1),
// so a rolling step is ONE table offset (one PRMT), two conflict-free LDS.64 and one three-input XOR per
// 32-bit half -- against two code extractions, two address multiplies, two LDS.128 and two more XORs per
// half with the general 64-entry tables.  The k-1 (s-1) initial folds go two bases at a time through
//     F2X[c0, c1] = rol(A[c0], 1) ^ A[c1],   F2Y[c0, c1] = ror(RB[c0], 1) ^ RB[c1],   RB[c] = rol(B[c], h-1).
// Fast-table block (byte offsets from its base FT): X 0, Y 128, F2X 256, F2Y 384, F1X 512 (A[c], 4 entries),
// F1Y 544 (RB[c]); syncmer k-mer hasher: XK 576, YK 704, F2YK 832, F1YK 960.  992 bytes.
#define FT_X 0u
#define FT_Y 128u
#define FT_F2X 256u
#define FT_F2Y 384u
#define FT_F1X 512u
#define FT_F1Y 544u
#define FT_XK 576u
#define FT_YK 704u
#define FT_F2YK 832u
#define FT_F1YK 960u
#define FT_BYTES 1024u
__device__ __forceinline__ uint64_t lds_u64(const uint8_t *sm, uint32_t o) {
    return *reinterpret_cast<const uint64_t *>(sm + o);
}
__device__ __forceinline__ void build_fast_tables(uint8_t *ft, uint32_t t, int h, int kk, bool sync) {
    if (t >= 16) return;
    const char letter[4] = {'A', 'C', 'T', 'G'}; // class = (byte >> 1) & 3
    const uint32_t lo = t & 3u, hi = t >> 2;
    const uint64_t Alo = fwd_seed((uint32_t)letter[lo]), Ahi = fwd_seed((uint32_t)letter[hi]);
    const uint64_t Blo = rev_seed((uint32_t)letter[lo]), Bhi = rev_seed((uint32_t)letter[hi]);
    uint64_t *q = reinterpret_cast<uint64_t *>(ft);
    q[FT_X / 8 + t] = Alo ^ rol64(Ahi, (unsigned)h);
    q[FT_Y / 8 + t] = rol64(Blo, (unsigned)(h - 1)) ^ ror64(Bhi, 1);
    q[FT_F2X / 8 + t] = rol64(Alo, 1) ^ Ahi;
    q[FT_F2Y / 8 + t] = ror64(rol64(Blo, (unsigned)(h - 1)), 1) ^ rol64(Bhi, (unsigned)(h - 1));
    if (t < 4) {
        q[FT_F1X / 8 + t] = Alo;
        q[FT_F1Y / 8 + t] = rol64(Blo, (unsigned)(h - 1));
    }
    if (sync) {
        q[FT_XK / 8 + t] = Alo ^ rol64(Ahi, (unsigned)kk);
        q[FT_YK / 8 + t] = rol64(Blo, (unsigned)(kk - 1)) ^ ror64(Bhi, 1);
        q[FT_F2YK / 8 + t] = ror64(rol64(Blo, (unsigned)(kk - 1)), 1) ^ rol64(Bhi, (unsigned)(kk - 1));
        if (t < 4) q[FT_F1YK / 8 + t] = rol64(Blo, (unsigned)(kk - 1));
    }
}
// ASCII word -> four fast-path bytes; bad accumulates a non-zero value when a byte is not one of ACGTacgt
__device__ __forceinline__ uint32_t fast_word(uint32_t w, uint32_t &bad) {
    const uint32_t x = w | 0x20202020u;                       // lower case
    const uint32_t t = (x >> 1) & 0x03030303u;                // class of every byte
    const uint32_t u2 = t | (t >> 4);                         // nibble pairs in bytes 0 and 2
    const uint32_t sel = __byte_perm(u2, 0u, 0x4420u);        // four nibbles = PRMT selector
    bad |= x ^ __byte_perm(0x67746361u, 0u, sel);             // 'a','c','t','g' by class
    return t * 40u;
}
// pair-table offsets of one block of W steps: in bytes from pin, out bytes from pout
template <int W> struct PairWords {
    static constexpr int NWORD = (W + 3 + 3) / 4, NG = (W + 3) / 4;
    uint32_t x[NG];
    __device__ __forceinline__ void load(const uint8_t *sm, uint32_t pin, uint32_t pout) {
        const uint32_t ai = pin & ~3u, ao = pout & ~3u;
        const uint32_t si = 0x3210u + 0x1111u * (pin & 3u), so = 0x3210u + 0x1111u * (pout & 3u);
        uint32_t wi[NWORD + 1], wo[NWORD + 1];
#pragma unroll
        for (int i = 0; i < NWORD; i++) {
            wi[i] = *reinterpret_cast<const uint32_t *>(sm + ai + 4u * i);
            wo[i] = *reinterpret_cast<const uint32_t *>(sm + ao + 4u * i);
        }
        wi[NWORD] = 0; wo[NWORD] = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int g1 = g + 1 < NWORD ? g + 1 : NWORD;
            x[g] = (__byte_perm(wi[g], wi[g1], si) & 0x18181818u) | (__byte_perm(wo[g], wo[g1], so) & 0x60606060u);
        }
    }
    __device__ __forceinline__ uint32_t off(const int j) const { return __byte_perm(x[j >> 2], 0u, 0x4440u | (j & 3)); }
};
// the h-1 initial folds of a hasher over fast-path bytes starting at shared offset sb
__device__ __forceinline__ void fast_fold(const uint8_t *sm, uint32_t ft, uint32_t sb, int n, uint64_t &f, uint64_t &r) {
    int j = 0;
    for (; j + 1 < n; j += 2) {
        const uint32_t o = (lds_u8(sm, sb + j) & 0x18u) | (lds_u8(sm, sb + j + 1) & 0x60u);
        f = rol64(f, 2) ^ lds_u64(sm, ft + FT_F2X + o);
        r = ror64(r, 2) ^ lds_u64(sm, ft + FT_F2Y + o);
    }
    if (j < n) {
        const uint32_t o = lds_u8(sm, sb + j) & 0x18u;
        f = rol1(f) ^ lds_u64(sm, ft + FT_F1X + o);
        r = ror1(r) ^ lds_u64(sm, ft + FT_F1Y + o);
    }
}

// NextMinimizer (sketch.go:205-309) over one item; its codes start at shared offset sb.

// FAST: the tile holds fast-path bytes and ft is the fast-table block; else 6-bit codes and tabIn/tabOut.
template <int W, bool FAST, class SinkT>
__device__ __forceinline__ void minimizer_item_reg(const uint8_t *sm, uint32_t sb, uint32_t nstep, int k,
                                                   uint32_t w_runtime, uint32_t tabIn, uint32_t tabOut,
                                                   uint32_t ft, SinkT &sink) {
    Roll h;
    h.f = 0; h.r = 0;
    if (FAST) fast_fold(sm, ft, sb, k - 1, h.f, h.r);
    else
        for (int j = 0; j < k - 1; j++) h.fold(B200SK_LD128(tabIn, sb + j));
    WinReg<W> wm;
    wm.init(w_runtime);
    uint32_t prev = W - 1; // frame-relative position of the previous window's minimum (none yet)
    uint32_t pin = sb + (uint32_t)k - 1; // incoming code of step 0
    uint32_t pout = sb - 1;              // outgoing code of step j is pout + j (step 0 has none)
    uint64_t mv;
    uint32_t mu;
    CodeWords<FAST ? 1 : W> cin, cout;
    PairWords<FAST ? W : 1> pw;
#define B200SK_LOAD_BLOCK()                                                \
    if (FAST) pw.load(sm, pin, pout);                                      \
    else { cin.load(sm, pin); cout.load(sm, pout); }
#define B200SK_ROLL(J)                                                     \
    if (FAST) {                                                            \
        const uint32_t o = pw.off(J);                                      \
        h.f = rol1(h.f) ^ lds_u64(sm, ft + FT_X + o);                      \
        h.r = ror1(h.r) ^ lds_u64(sm, ft + FT_Y + o);                      \
    } else h.roll(B200SK_TAB(tabIn, cin.code(J)), B200SK_TAB(tabOut, cout.code(J)));
#define B200SK_MIN_STEP(J, FIRST)                                          \
    if (wm.push(J, FIRST, h.canonical(), mv, mu)) {                        \
        sink.emit_if(mu != prev, mv, mu - prev); /* sketch.go:297-307 */     \
        prev = mu;                                                         \
    }
    B200SK_LOAD_BLOCK()
    if (FAST) { // the first k-mer has no outgoing base
        const uint32_t o = pw.off(0) & 0x18u;
        h.f = rol1(h.f) ^ lds_u64(sm, ft + FT_F1X + o);
        h.r = ror1(h.r) ^ lds_u64(sm, ft + FT_F1Y + o);
    } else h.fold(B200SK_TAB(tabIn, cin.code(0)));
    B200SK_MIN_STEP(0, true)
#pragma unroll
    for (int j = 1; j < W; j++) {
        B200SK_ROLL(j)
        B200SK_MIN_STEP(j, true)
    }
    wm.close_block();
    prev -= W;
    pin += W; pout += W;
    uint32_t u0 = W;
    while (u0 + W <= nstep) { // full blocks
        B200SK_LOAD_BLOCK()
#pragma unroll
        for (int j = 0; j < W; j++) {
            B200SK_ROLL(j)
            B200SK_MIN_STEP(j, false)
        }
        wm.close_block();
        prev -= W;
        pin += W; pout += W;
        u0 += W;
    }
    const uint32_t rem = nstep - u0; // tail: fewer than W elements left
    B200SK_LOAD_BLOCK()
#pragma unroll
    for (int j = 0; j < W - 1; j++) {
        if ((uint32_t)j >= rem) break;
        B200SK_ROLL(j)
        B200SK_MIN_STEP(j, false)
    }
#undef B200SK_MIN_STEP
#undef B200SK_ROLL
#undef B200SK_LOAD_BLOCK
}

// ------------------------------------------------------------------ keyed window minimum (W <= 16)
// The same walk with the window minimum kept on 32-bit KEYS instead of (64-bit value, position) pairs:
//     key = (top 26 bits of the canonical hash) << 6 | frame position,
// so a leftmost minimum is ONE unsigned min (VIMNMX / VIMNMX3 / VIADDMNMX) instead of a 64-bit compare
// and three selects.  Truncating the hash is made exact by detection, not by luck: every min the window
// logic takes also feeds "later key - earlier key" into an accumulator (acc = min(acc, difference), fused
// into one VIADDMNMX with a negated register operand); two keys whose 26 hash bits agree differ by their
// positions only, i.e. by 1..63, and no other pair that can decide a window gives a difference that small
// unless the hashes are that close -- acc <= 63 at the end of the item means "some comparison was decided by
// position bits": the walk returns false and the caller walks the item again with the exact 64-bit window
// (minimizer_item_reg).  On random reads that happens for ~2e-5 of the reads.
//
// The emitted VALUES never travel through the mins: the canonical hashes of the last W elements stay in W
// 64-bit registers (pv), the window logic only marks the frame position of every window minimum in a bit
// mask, and an element is written out when it leaves the window (W-1 steps after it was hashed) if its bit
// is set -- by then its register index is a compile-time constant.  Positions are not staged at all: the
// mask word of every block goes to shared memory and the flush turns set bits back into Index() values.
struct KeySink {
    uint32_t ov;    // shared-window address of this lane's next value slot (slot stride 256 B)
    uint32_t ovlim; // address of the scratch slot behind the last real one: ov sticks there on overflow
    uint32_t cw;    // position byte of the slot at ov lives at (ov >> 3) + cw  ([slot][lane] bytes, slot stride 32 B)
    uint32_t tcnt;  // elements written out so far (counted per block from the mask, so it survives an overflow)
    // pos: stream index of the element (its Index() is q0 + pos), < 256
    __device__ __forceinline__ void emit_if(uint32_t pred, uint64_t v, uint32_t pos) {
        asm volatile(
            "{\n\t.reg .pred q;\n\t.reg .u32 t, a;\n\tsetp.ne.u32 q, %1, 0;\n\t@q st.shared.u64 [%0], %3;\n\t"
            "shr.u32 a, %0, 3;\n\tadd.u32 a, a, %4;\n\t@q st.shared.u8 [a], %5;\n\t"
            "selp.u32 t, 256, 0, q;\n\tadd.u32 t, %0, t;\n\tmin.u32 %0, t, %2;\n\t}"
            : "+r"(ov)
            : "r"(pred), "r"(ovlim), "l"(v), "r"(cw), "r"(pos));
    }
    __device__ __forceinline__ void block_done(uint32_t bits) { tcnt += __popc(bits); }
};
__device__ __forceinline__ uint32_t key_sub_min(uint32_t acc, uint32_t later, uint32_t earlier) { // min(acc, later - earlier)
    uint32_t d;
    asm("{\n\t.reg .u32 t;\n\tsub.u32 t, %1, %2;\n\tmin.u32 %0, t, %3;\n\t}" : "=r"(d) : "r"(later), "r"(earlier), "r"(acc));
    return d;
}
__device__ __forceinline__ uint32_t bit_of_key(uint32_t key) { // 1 << (key & 31)
    uint32_t b;
    asm("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(b) : "r"(0u), "r"(1u), "r"(key));
    return b;
}
// key = (hi & kmask) | pos in ONE LOP3: kmask (0xffffffc0) comes in as a run-time value -- with an immediate mask
// ptxas needs two instructions, and it would also re-derive "key - W" as a second OR instead of folding the
// subtraction into the suffix minimum (VIADDMNMX)
__device__ __forceinline__ uint32_t make_key(uint32_t hi, uint32_t kmask, uint32_t pos) {
    uint32_t k;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(k) : "r"(hi), "r"(kmask), "r"(pos));
    return k;
}

// returns false when a comparison was (or may have been) decided by the position bits: walk again exactly
template <int W, bool FAST>
__device__ __forceinline__ bool minimizer_item_key(const uint8_t *sm, uint32_t sb, uint32_t nstep, int k,
                                                   uint32_t tabIn, uint32_t tabOut, uint32_t ft, uint32_t kmask,
                                                   KeySink &sink) {
    static_assert(W >= 2 && W <= 16, "frame positions 0..2W-1 must fit the 32-bit mask");
    Roll h;
    h.f = 0; h.r = 0;
    if (FAST) fast_fold(sm, ft, sb, k - 1, h.f, h.r);
    else
        for (int j = 0; j < k - 1; j++) h.fold(B200SK_LD128(tabIn, sb + j));
    uint64_t pv[W]; // canonical hash of the newest W elements, slot = element index within its block
    uint32_t kk[W] = {}; // [J]: key of element J of the current block once hashed; before that the suffix minimum
                    // (already in this frame's coordinates) of the previous block from element J on
    uint32_t pk = 0, mask = 0, acc = 0xffffffffu;
    uint32_t fpos = 0u - (uint32_t)W;    // stream index of frame position 0
    uint32_t pin = sb + (uint32_t)k - 1; // incoming code of step 0
    uint32_t pout = sb - 1;              // outgoing code of step j is pout + j (step 0 has none)
    CodeWords<FAST ? 1 : W> cin, cout;
    PairWords<FAST ? W : 1> pw;
#define B200SK_LOAD_BLOCK()                                                \
    if (FAST) pw.load(sm, pin, pout);                                      \
    else { cin.load(sm, pin); cout.load(sm, pout); }
#define B200SK_ROLL(J)                                                     \
    if (FAST) {                                                            \
        const uint32_t o = pw.off(J);                                      \
        h.f = rol1(h.f) ^ lds_u64(sm, ft + FT_X + o);                      \
        h.r = ror1(h.r) ^ lds_u64(sm, ft + FT_Y + o);                      \
    } else h.roll(B200SK_TAB(tabIn, cin.code(J)), B200SK_TAB(tabOut, cout.code(J)));
    // element J of the current block (frame position W + J).  Not FIRST: the window [J+1, W+J] closes here.
#define B200SK_KEY_STEP(J, FIRST)                                                                     \
    {                                                                                                 \
        const uint64_t c = h.canonical();                                                             \
        const uint32_t key = make_key((uint32_t)(c >> 32), kmask, (uint32_t)(W + (J)));               \
        if ((J) == 0) pk = key;                                                                       \
        else { acc = key_sub_min(acc, key, pk); pk = min(pk, key); }                                  \
        if (!(FIRST) || (J) == W - 1) {                                                               \
            uint32_t wk = pk;                                                                         \
            if (!(FIRST) && (J) != W - 1) { acc = key_sub_min(acc, pk, kk[(J) + 1]); wk = min(pk, kk[(J) + 1]); } \
            mask |= bit_of_key(wk);                                                                   \
        }                                                                                             \
        kk[J] = key;                                                                                  \
        pv[J] = c;                                                                                    \
        /* the leftmost element of this window, frame position J+1, can never be a minimum again */     \
        if (!(FIRST) || (J) == W - 1) sink.emit_if(mask & (2u << (J)), pv[((J) + 1) % W], fpos + (uint32_t)((J) + 1)); \
    }
    // after element W-1: the block's keys become suffix minima in the next frame's coordinates (-W)
#define B200SK_KEY_CLOSE()                                                                            \
    {                                                                                                 \
        kk[W - 1] -= (uint32_t)W;                                                                     \
        _Pragma("unroll") for (int jj = W - 2; jj >= 1; jj--) {                                       \
            const uint32_t t = kk[jj + 1] - kk[jj]; /* + W: kk[jj] is still in the old frame */         \
            acc = min(acc, t + (uint32_t)W);                                                          \
            kk[jj] = min(kk[jj] - (uint32_t)W, kk[jj + 1]);                                           \
        }                                                                                             \
        sink.block_done(mask & (((1u << W) - 1u) << 1)); /* frame positions 1..W: written out in this block */ \
        mask >>= W;                                                                                   \
        fpos += (uint32_t)W;                                                                          \
    }
    B200SK_LOAD_BLOCK()
    if (FAST) { // the first k-mer has no outgoing base
        const uint32_t o = pw.off(0) & 0x18u;
        h.f = rol1(h.f) ^ lds_u64(sm, ft + FT_F1X + o);
        h.r = ror1(h.r) ^ lds_u64(sm, ft + FT_F1Y + o);
    } else h.fold(B200SK_TAB(tabIn, cin.code(0)));
    B200SK_KEY_STEP(0, true)
#pragma unroll
    for (int j = 1; j < W; j++) {
        B200SK_ROLL(j)
        B200SK_KEY_STEP(j, true)
    }
    B200SK_KEY_CLOSE()
    pin += W; pout += W;
    uint32_t u0 = W;
    while (u0 + W <= nstep) { // full blocks
        B200SK_LOAD_BLOCK()
#pragma unroll
        for (int j = 0; j < W; j++) {
            B200SK_ROLL(j)
            B200SK_KEY_STEP(j, false)
        }
        B200SK_KEY_CLOSE()
        pin += W; pout += W;
        u0 += W;
    }
    const uint32_t rem = nstep - u0; // tail: fewer than W elements left
    B200SK_LOAD_BLOCK()
#pragma unroll
    for (int j = 0; j < W - 1; j++) {
        if ((uint32_t)j >= rem) break;
        B200SK_ROLL(j)
        B200SK_KEY_STEP(j, false)
    }
    // drain: the last window was [rem, W+rem-1]; positions rem+1 .. W+rem-1 are still inside it
#pragma unroll
    for (int q = 1; q <= 2 * W - 2; q++)
        if ((uint32_t)q > rem && (uint32_t)q < (uint32_t)W + rem) sink.emit_if(mask & (1u << q), pv[q % W], fpos + (uint32_t)q);
    sink.block_done(mask & ~1u);
#undef B200SK_KEY_CLOSE
#undef B200SK_KEY_STEP
#undef B200SK_ROLL
#undef B200SK_LOAD_BLOCK
    return acc > 63u;
}

// The exact walk (minimizer_item_reg) staged in the keyed format (values + absolute position bytes), for the items
// the keyed walk hands back.
struct PosListSink {
    uint8_t *sm;
    uint32_t ov0, op0, cap, cnt, sidx; // sidx: stream index of the last element + 1
    __device__ __forceinline__ void emit_if(uint32_t pred, uint64_t v, uint32_t delta) {
        if (pred) {
            sidx += delta;
            const uint32_t slot = min(cnt, cap);
            *reinterpret_cast<uint64_t *>(sm + ov0 + slot * 256u) = v;
            sm[op0 + slot * 32u] = (uint8_t)(sidx - 1u);
            cnt++;
        }
    }
};

// NextSyncmer (sketch.go:312-477, bounded closed syncmers, s < k) over one item.  W = 2(k-s): the window of
// s-mer hashes.  For the window starting at idx the leftmost minimum s-mer m anchors k-mer b = m if
// m - idx < k-s, else m - (k-s) (sketch.go:414-420); a k-mer is emitted when b changes and b <= end.
// The last k-s k-mer hashes wait in a shared-memory ring (slot = position mod (k-s)).
// tabs: offsets of tInS {A, rolB_{s-1}}, tOutS {rolA_s, rorB_1}, tOutK {rolA_k, rorB_1} (16 B entries) and
// tInK {rolB_{k-1}} (8 B entries); FAST: the fast-table block ft instead.  lim0 = end - q0 (stream index
// of the last emittable k-mer).
template <int W, bool FAST, class SinkT>
__device__ __forceinline__ void syncmer_item_reg(uint8_t *sm, uint32_t sb, uint32_t nstep, int s,
                                                 uint32_t w_runtime, uint32_t tInS, uint32_t tOutS, uint32_t tInK,
                                                 uint32_t tOutK, uint32_t ft, uint32_t kring, int32_t lim0,
                                                 uint32_t halo, SinkT &sink) {
    constexpr int D = W / 2;
    Roll hs_, hk_; // s-mer and k-mer hashers
    hs_.f = hs_.r = hk_.f = hk_.r = 0;
    if (FAST) {
        int j = 0;
        for (; j + 1 < s - 1; j += 2) {
            const uint32_t o = (lds_u8(sm, sb + j) & 0x18u) | (lds_u8(sm, sb + j + 1) & 0x60u);
            const uint64_t x = lds_u64(sm, ft + FT_F2X + o);
            hs_.f = rol64(hs_.f, 2) ^ x;
            hs_.r = ror64(hs_.r, 2) ^ lds_u64(sm, ft + FT_F2Y + o);
            hk_.r = ror64(hk_.r, 2) ^ lds_u64(sm, ft + FT_F2YK + o);
        }
        if (j < s - 1) {
            const uint32_t o = lds_u8(sm, sb + j) & 0x18u;
            hs_.f = rol1(hs_.f) ^ lds_u64(sm, ft + FT_F1X + o);
            hs_.r = ror1(hs_.r) ^ lds_u64(sm, ft + FT_F1Y + o);
            hk_.r = ror1(hk_.r) ^ lds_u64(sm, ft + FT_F1YK + o);
        }
        hk_.f = hs_.f; // both hashers fold the same bases with the same forward seeds
    } else
        for (int j = 0; j < s - 1; j++) {
            const uint32_t c = lds_u8(sm, sb + j);
            const ulonglong2 e = lds_v2u64(sm, tInS + c * 16u);
            hs_.fold(e);
            hk_.fold(make_ulonglong2(e.x, *reinterpret_cast<const uint64_t *>(sm + tInK + c * 8u)));
        }
    WinReg<W> wm;
    wm.init(w_runtime);
    uint32_t prevb = W - 1; // frame-relative position of the previous window's k-mer (none yet: stream -1)
    int32_t lim = lim0 + W;       // frame-relative version of lim0 (frame starts at -W in block 0)
    uint32_t pin = sb + (uint32_t)s - 1; // incoming code of step 0
    uint32_t pos = sb - 1;               // s-mer outgoing code of step j is pos + j
    uint32_t pok = sb - 1 - D;           // k-mer outgoing code of step j is pok + j (steps <= D have none)
    uint64_t mv;
    uint32_t mu;
#define B200SK_SYNC_HASH(J, FIRST)                                                                         \
    if (FAST) {                                                                                            \
        const uint32_t ci = lds_u8(sm, pin + (J)) & 0x18u;                                                 \
        if ((FIRST) && (J) == 0) {                                                                         \
            hs_.f = rol1(hs_.f) ^ lds_u64(sm, ft + FT_F1X + ci);                                           \
            hs_.r = ror1(hs_.r) ^ lds_u64(sm, ft + FT_F1Y + ci);                                           \
        } else {                                                                                           \
            const uint32_t o = ci | (lds_u8(sm, pos + (J)) & 0x60u);                                       \
            hs_.f = rol1(hs_.f) ^ lds_u64(sm, ft + FT_X + o);                                              \
            hs_.r = ror1(hs_.r) ^ lds_u64(sm, ft + FT_Y + o);                                              \
        }                                                                                                  \
        if ((FIRST) && (J) <= D) {                                                                         \
            hk_.f = rol1(hk_.f) ^ lds_u64(sm, ft + FT_F1X + ci);                                           \
            hk_.r = ror1(hk_.r) ^ lds_u64(sm, ft + FT_F1YK + ci);                                          \
        } else {                                                                                           \
            const uint32_t o = ci | (lds_u8(sm, pok + (J)) & 0x60u);                                       \
            hk_.f = rol1(hk_.f) ^ lds_u64(sm, ft + FT_XK + o);                                             \
            hk_.r = ror1(hk_.r) ^ lds_u64(sm, ft + FT_YK + o);                                             \
        }                                                                                                  \
    } else {                                                                                               \
        const uint32_t c = lds_u8(sm, pin + (J));                                                          \
        const ulonglong2 e = lds_v2u64(sm, tInS + c * 16u);                                                \
        const uint64_t ek = *reinterpret_cast<const uint64_t *>(sm + tInK + c * 8u);                       \
        if ((FIRST) && (J) == 0) hs_.fold(e); else hs_.roll(e, B200SK_LD128(tOutS, pos + (J)));           \
        if ((FIRST) && (J) <= D) hk_.fold(make_ulonglong2(e.x, ek));                                       \
        else hk_.roll(make_ulonglong2(e.x, ek), B200SK_LD128(tOutK, pok + (J)));                           \
    }
#define B200SK_SYNC_STEP(J, FIRST)                                                                         \
    {                                                                                                      \
        B200SK_SYNC_HASH(J, FIRST)                                                                         \
        *reinterpret_cast<uint64_t *>(sm + kring + ((J) % D) * 256u) = hk_.canonical(); /* k-mer (J-D) */  \
        if (wm.push(J, FIRST, hs_.canonical(), mv, mu)) {                                                  \
            const uint32_t off = mu - (uint32_t)((J) + 1);        /* m - idx */                            \
            const uint32_t b = off < (uint32_t)D ? mu : mu - (uint32_t)D;                                  \
            const uint64_t kv = *reinterpret_cast<const uint64_t *>(sm + kring + (mu % (uint32_t)D) * 256u); \
            uint32_t ok = (b != prevb) & ((int32_t)b <= lim);                                              \
            if ((FIRST) && (J) == W - 1) ok |= halo; /* a non-first chunk's seeding window: always staged, dropped later */ \
            sink.emit_if(ok, kv, b - prevb);                                                                  \
            prevb = b;                                                                                     \
        }                                                                                                  \
    }
#pragma unroll
    for (int j = 0; j < W; j++) B200SK_SYNC_STEP(j, true)
    wm.close_block();
    prevb -= W; lim -= W;
    pin += W; pos += W; pok += W;
    uint32_t u0 = W;
    while (u0 + W <= nstep) {
#pragma unroll
        for (int j = 0; j < W; j++) B200SK_SYNC_STEP(j, false)
        wm.close_block();
        prevb -= W; lim -= W;
        pin += W; pos += W; pok += W;
        u0 += W;
    }
    const uint32_t rem = nstep - u0;
#pragma unroll
    for (int j = 0; j < W - 1; j++) {
        if ((uint32_t)j >= rem) break;
        B200SK_SYNC_STEP(j, false)
    }
#undef B200SK_SYNC_STEP
#undef B200SK_SYNC_HASH
}

// ProteinMinimizerSketch.Next (sketch-protein.go:106-210) over one item of amino acids (the frame was translated
// by k_translate): wyhash(seed 1) of every amino-acid k-mer (k <= 16) from a 128-bit register window of the last
// 16 residues, then the same window minimum as NextMinimizer.
template <int W, class SinkT>
__device__ __forceinline__ void protmin_item_reg(const uint8_t *sm, uint32_t sb, uint32_t nstep, int k,
                                                 uint32_t w_runtime, SinkT &sink) {
    uint64_t wlo = 0, whi = 0;
    for (int j = 0; j < k - 1; j++) {
        const uint64_t aa = lds_u8(sm, sb + j);
        wlo = (wlo >> 8) | (whi << 56);
        whi = (whi >> 8) | (aa << 56);
    }
    WinReg<W> wm;
    wm.init(w_runtime);
    const WyPlan wp = wy_plan((uint32_t)k);
    uint32_t prev = W - 1; // frame-relative position of the previous window's minimum (none yet)
    uint32_t pin = sb + (uint32_t)k - 1;
    uint64_t mv;
    uint32_t mu;
#define B200SK_PM_STEP(J, FIRST)                                           \
    {                                                                      \
        const uint64_t aa = lds_u8(sm, pin + (J));                         \
        wlo = (wlo >> 8) | (whi << 56);                                    \
        whi = (whi >> 8) | (aa << 56);                                     \
        if (wm.push(J, FIRST, wyhash_window(wlo, whi, wp), mv, mu)) {          \
            sink.emit_if(mu != prev, mv, mu - prev);                       \
            prev = mu;                                                     \
        }                                                                  \
    }
#pragma unroll
    for (int j = 0; j < W; j++) B200SK_PM_STEP(j, true)
    wm.close_block();
    prev -= W;
    pin += W;
    uint32_t u0 = W;
    while (u0 + W <= nstep) {
#pragma unroll
        for (int j = 0; j < W; j++) B200SK_PM_STEP(j, false)
        wm.close_block();
        prev -= W;
        pin += W;
        u0 += W;
    }
    const uint32_t rem = nstep - u0;
#pragma unroll
    for (int j = 0; j < W - 1; j++) {
        if ((uint32_t)j >= rem) break;
        B200SK_PM_STEP(j, false)
    }
#undef B200SK_PM_STEP
}

// ------------------------------------------------------------------ skewed staging (uniform reads of n x 32 bytes)
// The tile's rows of `rowlen` bytes (a multiple of 32) lie back to back at `landing` (where the bulk copy put them:
// the staged-list area, dead between two tiles); they move to tilebuf at a row stride of rowlen + 4 and are rewritten
// on the way: FASTB -> fast-path bytes (returns false when some byte is not one of ACGTacgt: the caller calls again
// for 6-bit codes -- the landing area still holds the ASCII).  nvec 16-byte vectors cover the rows.
// (Lanes fetching the rows from global memory themselves was tried first: 16 dependent trips per lane for a 256-byte
// row tile, 289 Gbases/s; with the loads batched the registers of W >= 11 spill.)
template <bool FASTB>
__device__ __forceinline__ bool restage_skewed(uint8_t *tilebuf, const uint8_t *landing, uint32_t nvec, uint32_t rowlen,
                                               uint32_t lane) {
    const uint32_t vpr = rowlen >> 4; // vectors per row
    uint32_t bad = 0;
    for (uint32_t v = lane; v < nvec; v += 32u) {
        uint4 x = *reinterpret_cast<const uint4 *>(landing + v * 16u);
        const uint32_t row = v / vpr, col = v - row * vpr;
        if (FASTB) {
            x.x = fast_word(x.x, bad); x.y = fast_word(x.y, bad);
            x.z = fast_word(x.z, bad); x.w = fast_word(x.w, bad);
        } else {
            x.x = codes_of_word(x.x); x.y = codes_of_word(x.y);
            x.z = codes_of_word(x.z); x.w = codes_of_word(x.w);
        }
        uint32_t *d = reinterpret_cast<uint32_t *>(tilebuf + row * (rowlen + 4u) + col * 16u);
        d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
    }
    return !__any_sync(0xffffffffu, bad != 0);
}

// ------------------------------------------------------------------ kernel: one tile per WARP
// A tile is 32 consecutive items; every warp runs its own ticket -> TMA -> walk -> look-back -> ordered
// copy loop with no block-wide barrier, so warps drift freely and cover each other's latencies.
// Shared memory: tables at 0; warp w owns [sm_tile + w*stride, +stride):
//   +0 mbarrier, +16 tile bytes, +sm_ring k-mer ring (syncmer), +sm_listv values, +sm_listp position deltas.
// KEYED (minimizers, W <= 16): the keyed walk above with the exact walk behind it; the lists then hold values only
// and the position bytes are absolute stream indices instead of deltas.
// SHARD: this rank's tiles are chunks of ONE tile chain shared by several GPUs (b200sk_enqueue_device_sharded): the
// look-back runs over the ranks' replicated status words, the output arrays are the root's, indexed globally.  A
// template parameter, not a run-time branch: kept in the single-GPU kernel it cost 5 % (registers 118 -> 122).
// Warps per CTA an instantiation is launched with at most.  The register file is split over the four schedulers: with
// four warps on a scheduler a thread gets 128 registers whatever the CTA size, with three 168.  Wide syncmer windows
// (3 W words of window state + two hashers) spill at 128, and shared memory holds only 13-14 of their warps anyway --
// 13 or 14 warps run no faster than 12 (the ordered chain moves at the pace of the schedulers that hold four).
template <int MODE, int W> constexpr int max_warps() { return MODE == B200SK_MODE_SYNCMER && W >= 16 ? 12 : 16; }
// SKEW: the instantiation the host picks for batches whose longest read is a multiple of 32 bytes (a.skew): tiles of
// equally long reads are then staged with a word of skew per lane (restage_skewed).  Its own instantiation, so that the
// kernel every other batch runs stays instruction for instruction what it was (in one kernel the extra code cost the
// 150-bp headline 1.7 %: 416 vs 424 Gbases/s, profiles/r02bb_readlen.txt).
template <int MODE, int W, bool KEYED, bool SHARD, bool SKEW = false>
__global__ void __launch_bounds__(32 * max_warps<MODE, W>(), 1) k_sparse_warp(const KArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    constexpr bool SYNC = MODE == B200SK_MODE_SYNCMER;
    constexpr bool PROT = MODE == B200SK_MODE_PROTEIN_MINIMIZER; // items are amino acids: no tables, no rewriting
    // tables (64 codes): [0,1K) in {A, rolB_{h-1}}, [1K,2K) out {rolA_h, rorB_1} for the streamed hash
    // (h = k for minimizers, s for syncmers); syncmer adds [2K,3K) k-mer out table, [3K,3.5K) rolB_{k-1}
    const int hk = SYNC ? a.s : a.k;
    constexpr uint32_t FT = SYNC ? 3584u : 2048u; // fast-table block (1 KB)
    if (!PROT) {
        ulonglong2 *tIn = reinterpret_cast<ulonglong2 *>(smem), *tOut = tIn + 64, *tOutK = tIn + 128;
        uint64_t *tInK = reinterpret_cast<uint64_t *>(smem + 3072);
        for (uint32_t c = tid; c < 64; c += blockDim.x) {
            const uint32_t b = byte_of_code(c);
            const uint64_t f = fwd_seed(b), r = rev_seed(b);
            tIn[c] = make_ulonglong2(f, rol64(r, (unsigned)(hk - 1)));
            tOut[c] = make_ulonglong2(rol64(f, (unsigned)hk), ror64(r, 1));
            if (SYNC) {
                tOutK[c] = make_ulonglong2(rol64(f, (unsigned)a.k), ror64(r, 1));
                tInK[c] = rol64(r, (unsigned)(a.k - 1));
            }
        }
        build_fast_tables(smem + FT, tid, hk, a.k, SYNC);
    }
    const uint32_t region = a.sm_tile + wid * a.sm_ring_bytes;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + region);
    uint8_t *tilebuf = smem + region + 16;
    const uint32_t s_tile = region + 16;
    const uint32_t s_kring = region + a.sm_ring + lane * 8u;
    uint64_t *listv = reinterpret_cast<uint64_t *>(smem + region + a.sm_listv);
    uint8_t *listp = smem + region + a.sm_listp;
    if (lane == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads(); // tables + barriers ready; the only block-wide barrier
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t n_items = a.n_items_dev ? *a.n_items_dev : a.n_items;
    uint32_t parity = 0;
    for (;;) {
        uint64_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.ticket, 1ULL);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        const uint64_t item0 = tile * 32ull;
        if (item0 >= n_items) {
            if (SHARD && lane == 0) bulk_wait_all(); // every bulk store of this warp has been written
            break;
        }
        // sharded batch: local tile -> tile of the global order (reads alike, 32 per tile).  Recomputed where it is
        // needed instead of being kept across the walk: the single-GPU kernels must not pay registers for it.
        auto global_tile = [&]() -> uint64_t {
            const uint32_t ct = a.shard_chunk_tiles;
            return ((tile / ct) * a.shard.n + a.shard.rank) * ct + tile % ct;
        };
        const uint32_t nvalid = (uint32_t)min((uint64_t)32, n_items - item0);
        Item it;
        item_geometry<MODE>(a, item0 + lane, n_items, it);
        const uint64_t lo = __shfl_sync(0xffffffffu, it.gb0, 0);
        const uint64_t hi = __shfl_sync(0xffffffffu, it.gb0 + it.nb, (int)nvalid - 1);
        const uint64_t lo_al = lo & ~15ULL;
        const uint64_t span = hi > lo_al ? hi - lo_al : 0;
        const uint32_t bytes = (uint32_t)((span + 15ULL) & ~15ULL);
        const bool span_ok = bytes <= a.sm_tile_bytes;
        // Reads of ONE length that is a multiple of 128 bytes would start 32 lanes on the same shared-memory bank
        // (lanes sit a read length apart): every word load of the walk a 32-way conflict, 235 instead of 425 Gbases/s
        // at 128 bp (profiles/r02ba_readlen.txt); multiples of 64 and 32 bytes share 2 and 4 banks.  For such a tile
        // the bulk copy lands in the staged-list area and the rewriting pass moves row r to r * (length + 4) of the
        // tile buffer: length / 4 + 1 words between the lanes is odd, so the 32 lanes sit on 32 different banks.
        // (The bulk copy cannot skew by itself: its addresses move in 16-byte units, which would leave 4-way conflicts.)
        uint64_t rowlen = 0;
        bool skew = false;
        if constexpr (SKEW) {
            rowlen = __shfl_sync(0xffffffffu, it.gb0, 1) - lo;
            const bool rows = __all_sync(0xffffffffu, !it.valid || it.gb0 == lo + (uint64_t)lane * rowlen);
            skew = rows && nvalid > 1u && (lo & 15ULL) == 0 && rowlen != 0 && (rowlen & 31ULL) == 0 && bytes != 0 &&
                   hi <= lo + (uint64_t)nvalid * rowlen && 32ULL * (rowlen + 4ULL) <= a.sm_tile_bytes &&
                   bytes <= (a.lcap + 1u) * 288u; // the bulk copy lands in the list area (values + position bytes)
        }
        if (SHARD) { // the previous tile's bulk stores must have read the buffer
            if (lane == 0) bulk_wait_read();
            __syncwarp();
        }
        if (lane == 0 && bytes && span_ok) {
            fence_proxy_async(); // the previous tile's generic-proxy writes to this buffer precede the async write
            mbar_expect_tx(mbar, bytes);
            tma_load_1d(SKEW && skew ? reinterpret_cast<uint8_t *>(listv) : tilebuf, a.bases + lo_al, bytes, mbar);
        }
        if (!span_ok && lane == 0) atomicOr(a.flags, B200SK_FLAG_SPAN);
        if (it.valid && it.first_chunk && a.status) a.status[SHARD ? global_tile() * 32ull + lane : it.r] = it.status;
        bool fast = false;
        if (SKEW && skew) {
            mbar_wait(mbar, parity);
            parity ^= 1u;
            const uint8_t *landing = reinterpret_cast<const uint8_t *>(listv);
            fast = restage_skewed<true>(tilebuf, landing, bytes >> 4, (uint32_t)rowlen, lane);
            if (!fast) restage_skewed<false>(tilebuf, landing, bytes >> 4, (uint32_t)rowlen, lane);
        } else if (bytes && span_ok) {
            mbar_wait(mbar, parity);
            parity ^= 1u;
            if (!PROT) {
                // ASCII -> fast bytes, 16 bytes per lane per trip; any byte outside ACGTacgt (alignment slop
                // included: a false alarm only costs the general path) -> fetch again and write 6-bit codes
                uint32_t bad = 0;
                for (uint32_t o = lane * 16u; o < bytes; o += 512u) {
                    uint4 v = *reinterpret_cast<uint4 *>(tilebuf + o);
                    v.x = fast_word(v.x, bad); v.y = fast_word(v.y, bad);
                    v.z = fast_word(v.z, bad); v.w = fast_word(v.w, bad);
                    *reinterpret_cast<uint4 *>(tilebuf + o) = v;
                }
                fast = !__any_sync(0xffffffffu, bad != 0);
                if (!fast) {
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
            if (!fast && !PROT) {
                // ASCII -> codes, 16 bytes per lane per trip
                for (uint32_t o = lane * 16u; o < bytes; o += 512u) {
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
        }
        __syncwarp();
        ListSink sink;
        sink.av = smem_base + region + a.sm_listv + lane * 8u;
        sink.ap = smem_base + region + a.sm_listp + lane;
        sink.cap = a.lcap; sink.cnt = 0;
        const uint32_t sb = s_tile + (SKEW && skew ? lane * ((uint32_t)rowlen + 4u) : (uint32_t)(it.gb0 - lo_al));
        const bool run = it.valid && it.nstep && span_ok;
        const int32_t lim0 = (int32_t)(it.end - it.q0);
        const uint32_t halo = it.q0 != it.p0 ? 1u : 0u;
        if constexpr (KEYED) {
            if (run) {
                KeySink ks;
                ks.ov = sink.av; ks.ovlim = sink.av + a.lcap * 256u;
                ks.cw = sink.ap - (sink.av >> 3); ks.tcnt = 0;
                const bool exact = fast ? minimizer_item_key<W, true>(smem, sb, it.nstep, a.k, 0u, 1024u, FT, a.key_mask, ks)
                                        : minimizer_item_key<W, false>(smem, sb, it.nstep, a.k, 0u, 1024u, FT, a.key_mask, ks);
                sink.cnt = ks.tcnt;
                if ((!exact && !(a.keyed & 4u)) || (a.keyed & 2u)) { // some window was decided by position bits: the exact 64-bit walk, same staging
                                                  // (keyed & 2: testing knob, every item takes this path)
                    PosListSink ps;
                    ps.sm = smem; ps.ov0 = sink.av - smem_base; ps.op0 = sink.ap - smem_base; ps.cap = a.lcap; ps.cnt = 0; ps.sidx = 0;
                    if (fast) minimizer_item_reg<W, true>(smem, sb, it.nstep, a.k, (uint32_t)a.w, 0u, 1024u, FT, ps);
                    else minimizer_item_reg<W, false>(smem, sb, it.nstep, a.k, (uint32_t)a.w, 0u, 1024u, FT, ps);
                    sink.cnt = ps.cnt;
                    if (a.rewalks) atomicAdd(a.rewalks, 1ULL);
                }
            }
        } else if (run) {
            if (PROT) protmin_item_reg<W>(smem, sb, it.nstep, a.k, (uint32_t)a.w, sink);
            else if (SYNC)
                if (fast) syncmer_item_reg<W, true>(smem, sb, it.nstep, a.s, 2u * (uint32_t)(a.k - a.s), 0u, 1024u, 3072u, 2048u,
                                                     FT, s_kring, lim0, halo, sink);
                else syncmer_item_reg<W, false>(smem, sb, it.nstep, a.s, 2u * (uint32_t)(a.k - a.s), 0u, 1024u, 3072u, 2048u,
                                                FT, s_kring, lim0, halo, sink);
            else
                if (fast) minimizer_item_reg<W, true>(smem, sb, it.nstep, a.k, (uint32_t)a.w, 0u, 1024u, FT, sink);
                else minimizer_item_reg<W, false>(smem, sb, it.nstep, a.k, (uint32_t)a.w, 0u, 1024u, FT, sink);
        }
        __syncwarp();
        // a non-first chunk walks one window more (the one before its first own window) to seed the
        // de-duplication; that window always emits first and is dropped here
        const uint32_t skip = (run && halo) ? 1u : 0u;
        const uint32_t cnt = sink.cnt - skip;
        const bool overflow = sink.cnt > a.lcap;
        const bool any_overflow = __any_sync(0xffffffffu, overflow);
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        const uint32_t excl = inc - cnt;
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        // Ordered allocation in two halves: publish the tile's count, then -- while the tiles before this one
        // finish -- scatter the staged lists into one contiguous buffer (the codes are dead now; nothing here
        // needs the prefix), and only then wait for the prefix.
        if (SHARD) lookback_publish_multi(a.shard, global_tile(), total);
#ifdef B200SK_EXPERIMENTS
        else if (!a.unordered) lookback_publish(a.tile_state, tile, total);
#else
        else lookback_publish(a.tile_state, tile, total);
#endif
        const uint32_t OB = a.sm_tile_bytes / 12u;
        uint64_t *obv = reinterpret_cast<uint64_t *>(tilebuf);
        uint32_t *obp = reinterpret_cast<uint32_t *>(tilebuf + (size_t)OB * 8u);
        auto scatter = [&](uint32_t r0) {
            uint32_t pos = it.q0 - 1u;
            for (uint32_t j = 0; j < sink.cnt; j++) {
                if (KEYED) pos = it.q0 + listp[j * 32u + lane]; // absolute stream index
                else pos += listp[j * 32u + lane];
                const uint32_t o = excl + j - skip - r0;
                if (j >= skip && o < OB) { // unsigned compare also rejects entries before r0
                    obv[o] = listv[j * 32u + lane];
                    obp[o] = pos;
                }
            }
            __syncwarp();
        };
        if (!SHARD && !any_overflow && total) scatter(0); // SHARD scatters after the prefix is known (alignment, below)
        uint64_t tb;
#ifdef B200SK_EXPERIMENTS
        if (a.unordered) { // timing experiment only (B200SK_UNORDERED=1): ranges in completion order, WRONG output order
            unsigned long long t0 = 0;
            if (lane == 0) t0 = atomicAdd(a.unordered, (unsigned long long)total);
            tb = __shfl_sync(0xffffffffu, t0, 0);
        } else tb = lookback_resolve(a.tile_state, tile, total, a.spin_ns);
#else
        if (SHARD) tb = lookback_resolve_multi(a.shard, global_tile(), total);
        else tb = lookback_resolve(a.tile_state, tile, total);
#endif
        const uint64_t mine = tb + excl;
        if (SHARD) { // the root's offset table, indexed by the global read
            const uint64_t gread = global_tile() * 32ull + lane;
            if (it.valid) a.out_off[gread] = a.out_base + mine;
            if (it.valid && gread + 1 == a.shard_n_reads) a.out_off[a.shard_n_reads] = a.out_base + mine + cnt;
        } else {
            if (it.valid && it.first_chunk) a.out_off[it.r] = a.out_base + mine;
            if (it.valid && it.last_item) a.out_off[a.n_reads] = a.out_base + mine + cnt;
        }
        const bool fits = tb + total <= a.capacity;
        if (!fits && lane == 0) atomicOr(a.flags, B200SK_FLAG_CAPACITY);
        if (fits && total) {
            if (!any_overflow) {
                // stream the buffer out coalesced; a tile with more elements than the buffer holds takes
                // further scatter + copy rounds
                if constexpr (SHARD) {
                    // The arrays are the root's (peer memory for every other rank).  Plain st.global into a saturated
                    // NVLink backs up into the LSU pipe the walking warps load their tables through, so the tile leaves
                    // as BULK stores (cp.async.bulk shared -> global: the copy engine's queue, not the LSU's): the
                    // elements are scattered so that shared and global addresses agree modulo 16 bytes, a lane stores
                    // the few elements before / after the 16-byte-aligned body, lane 0 issues one bulk store per array.
                    const uint32_t pw = a.out_pos ? a.pos_width : 0u;
                    const uint32_t OBS = (a.sm_tile_bytes - 48u) / (8u + pw); // entries per round (+1 value, +16 B of positions of shift)
                    uint8_t *pb = tilebuf + (((OBS + 1u) * 8u + 15u) & ~15u);
                    for (uint32_t r0 = 0; r0 < total; r0 += OBS) {
                        const uint32_t n = min(OBS, total - r0);
                        const uint64_t g0 = tb + r0; // global element index of entry 0 of this round
                        const uint32_t vs = (uint32_t)((uintptr_t)(a.out_val + g0) >> 3) & 1u;
                        const uint32_t pmask = pw ? 16u / pw - 1u : 0u;
                        const uint32_t ps = pw ? (uint32_t)(((uintptr_t)a.out_pos + g0 * pw) & 15u) / pw : 0u;
                        if (lane == 0) bulk_wait_read(); // the round before has left the buffer
                        __syncwarp();
                        {
                            uint32_t pos = it.q0 - 1u;
                            for (uint32_t j = 0; j < sink.cnt; j++) {
                                pos += listp[j * 32u + lane];
                                const uint32_t o = excl + j - skip - r0;
                                if (j >= skip && o < OBS) {
                                    obv[o + vs] = listv[j * 32u + lane];
                                    if (pw == 1) pb[o + ps] = (uint8_t)pos;
                                    else if (pw == 2) reinterpret_cast<uint16_t *>(pb)[o + ps] = (uint16_t)pos;
                                    else if (pw == 4) reinterpret_cast<uint32_t *>(pb)[o + ps] = pos;
                                }
                            }
                        }
                        fence_proxy_async(); // these generic-proxy writes precede the bulk store's reads
                        __syncwarp();
                        // values: head (g0 odd), 16-byte-aligned body, tail
                        uint64_t *gv = a.out_val + g0;
                        const uint32_t vh = min(n, vs), vb = (n - vh) & ~1u;
                        if (lane == 0 && vh) gv[0] = obv[vs];
                        if (lane == 1 && vh + vb < n) gv[n - 1] = obv[vs + n - 1];
                        if (lane == 0 && vb) bulk_store(gv + vh, obv + vs + vh, vb * 8u);
                        if (pw) {
                            uint8_t *gp = reinterpret_cast<uint8_t *>(a.out_pos) + g0 * pw;
                            const uint32_t ph = min(n, (pmask + 1u - ps) & pmask), pbody = (n - ph) & ~pmask;
                            // head and tail elements: at most 15 each, one per lane, moved as bytes
                            const uint32_t hb = ph * pw, tb0 = (ph + pbody) * pw, tbn = n * pw - tb0;
                            for (uint32_t i = lane; i < hb; i += 32u) gp[i] = pb[ps * pw + i];
                            for (uint32_t i = lane; i < tbn; i += 32u) gp[tb0 + i] = pb[ps * pw + tb0 + i];
                            if (lane == 0 && pbody) bulk_store(gp + hb, pb + ps * pw + hb, pbody * pw);
                        }
                        if (lane == 0) bulk_commit();
                    }
                } else
                for (uint32_t r0 = 0; r0 < total; r0 += OB) {
                    if (r0) scatter(r0);
                    const uint32_t n = min(OB, total - r0);
                    uint64_t *gv = a.out_val + tb + r0;
                    for (uint32_t i = lane; i < n; i += 32u) gv[i] = obv[i];
                    if (a.out_pos)
                        for (uint32_t i = lane; i < n; i += 32u) store_pos(a.out_pos, a.pos_width, tb + r0 + i, obp[i]);
                    __syncwarp();
                }
            } else {
                // rare (low-complexity reads): some item emitted more than its list holds.  Items that fit
                // write their lists straight to their final range; the others walk their item again with
                // the global sink (the codes are still in shared memory).
                if (!overflow) {
                    uint32_t pos = it.q0 - 1u;
                    for (uint32_t j = 0; j < sink.cnt; j++) {
                        if (KEYED) pos = it.q0 + listp[j * 32u + lane];
                        else pos += listp[j * 32u + lane];
                        if (j >= skip) {
                            a.out_val[mine + j - skip] = listv[j * 32u + lane];
                            if (a.out_pos) store_pos(a.out_pos, a.pos_width, mine + j - skip, pos);
                        }
                    }
                } else {
                    GlobalSink gs;
                    gs.gv = a.out_val + mine; gs.gp = a.out_pos; gs.gi = mine; gs.pw = a.pos_width;
                    gs.pos = it.q0 - 1u; gs.cnt = 0; gs.skip = skip;
                    if (PROT) protmin_item_reg<W>(smem, sb, it.nstep, a.k, (uint32_t)a.w, gs);
                    else if (SYNC)
                        if (fast) syncmer_item_reg<W, true>(smem, sb, it.nstep, a.s, 2u * (uint32_t)(a.k - a.s), 0u, 1024u, 3072u,
                                                             2048u, FT, s_kring, lim0, halo, gs);
                        else syncmer_item_reg<W, false>(smem, sb, it.nstep, a.s, 2u * (uint32_t)(a.k - a.s), 0u, 1024u, 3072u,
                                                        2048u, FT, s_kring, lim0, halo, gs);
                    else
                        if (fast) minimizer_item_reg<W, true>(smem, sb, it.nstep, a.k, (uint32_t)a.w, 0u, 1024u, FT, gs);
                        else minimizer_item_reg<W, false>(smem, sb, it.nstep, a.k, (uint32_t)a.w, 0u, 1024u, FT, gs);
                }
                __syncwarp();
            }
        }
    }
}

// ------------------------------------------------------------------ launch
// The keyed walk is an EXPERIMENT kept for A/B (B200SK_WALKER=keyed): bit-exact, 6 % fewer main-loop
// instructions, not faster -- the kernel is bound by its total integer-ALU instruction count, which the drain and the
// detection give back, and every item handed to the exact walk stalls the ordered output of all later tiles.
// Instantiated for the window sizes the A/B script uses.
template <int MODE, int W> constexpr bool has_keyed() {
    return MODE == B200SK_MODE_MINIMIZER && (W == 3 || W == 5 || W == 11 || W == 15);
}
template <int MODE, int W, bool SHARD>
static cudaError_t launch_w(const KArgs &a, int threads, int blocks, cudaStream_t st, int *occ) {
    const void *fn = (const void *)k_sparse_warp<MODE, W, false, SHARD>;
    if constexpr (has_keyed<MODE, W>() && !SHARD) {
        if (a.keyed) fn = (const void *)k_sparse_warp<MODE, W, true, false>;
    }
    if constexpr (MODE != B200SK_MODE_PROTEIN_MINIMIZER && !SHARD) { // single-GPU minimizers and syncmers
        if (a.skew && !a.keyed) fn = (const void *)k_sparse_warp<MODE, W, false, false, true>;
    }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.sm_total);
    if (e != cudaSuccess) return e;
    if (occ) {
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, a.sm_total);
        *occ = nb < 1 ? 1 : nb;
        return e;
    }
    void *args[] = {(void *)&a};
    return cudaLaunchKernel(fn, dim3((unsigned)blocks), dim3((unsigned)threads), args, a.sm_total, st);
}

// Window sizes with a register-resident instantiation.  The instantiations are spread over four translation
// units (this file compiled with -DB200SK_PART=0..3, see the Makefile) so that they build in parallel;
// part 0 also holds the dispatcher.  B200SK_FAST_BUILD (development) keeps a handful.
#ifndef B200SK_PART
#error "compile with -DB200SK_PART=0..8"
#endif
#ifdef B200SK_FAST_BUILD
#define B200SK_LIST_0 B200SK_W(B200SK_MODE_MINIMIZER, 3) B200SK_W(B200SK_MODE_MINIMIZER, 5) B200SK_W(B200SK_MODE_MINIMIZER, 11)
#define B200SK_LIST_1 B200SK_W(B200SK_MODE_MINIMIZER, 15) B200SK_W(B200SK_MODE_MINIMIZER, 20)
#define B200SK_LIST_2 B200SK_W(B200SK_MODE_SYNCMER, 2) B200SK_W(B200SK_MODE_SYNCMER, 10)
#define B200SK_LIST_3 B200SK_W(B200SK_MODE_SYNCMER, 20)
#define B200SK_LIST_4 B200SK_W(B200SK_MODE_PROTEIN_MINIMIZER, 5)
#else
#define B200SK_M(W) B200SK_W(B200SK_MODE_MINIMIZER, W)
#define B200SK_S(W) B200SK_W(B200SK_MODE_SYNCMER, W)
#define B200SK_LIST_0                                                                                        \
    B200SK_M(2) B200SK_M(3) B200SK_M(4) B200SK_M(5) B200SK_M(6) B200SK_M(7) B200SK_M(8) B200SK_M(9)          \
    B200SK_M(10) B200SK_M(11) B200SK_M(12) B200SK_M(13) B200SK_M(14) B200SK_M(15)
#define B200SK_LIST_1                                                                                        \
    B200SK_M(16) B200SK_M(17) B200SK_M(18) B200SK_M(19) B200SK_M(20) B200SK_M(21) B200SK_M(22) B200SK_M(23)  \
    B200SK_M(24)
// syncmer windows 2(k-s): even sizes
#define B200SK_LIST_2                                                                                        \
    B200SK_S(2) B200SK_S(4) B200SK_S(6) B200SK_S(8) B200SK_S(10) B200SK_S(12) B200SK_S(14) B200SK_S(16)
#define B200SK_LIST_3 B200SK_S(18) B200SK_S(20) B200SK_S(22) B200SK_S(24)
#define B200SK_P(W) B200SK_W(B200SK_MODE_PROTEIN_MINIMIZER, W)
#define B200SK_LIST_4                                                                                        \
    B200SK_P(2) B200SK_P(3) B200SK_P(4) B200SK_P(5) B200SK_P(6) B200SK_P(7) B200SK_P(8) B200SK_P(9)          \
    B200SK_P(10) B200SK_P(11) B200SK_P(12) B200SK_P(13) B200SK_P(14) B200SK_P(15) B200SK_P(16) B200SK_P(17)  \
    B200SK_P(18) B200SK_P(19) B200SK_P(20) B200SK_P(21) B200SK_P(22) B200SK_P(23) B200SK_P(24)
#endif
#define B200SK_CAT2(a, b) a##b
#define B200SK_CAT(a, b) B200SK_CAT2(a, b)
// parts 5..8 hold the SHARD instantiations of the lists of parts 0..3 (minimizers and syncmers)
#if B200SK_PART == 5
#define B200SK_MY_LIST B200SK_LIST_0
#elif B200SK_PART == 6
#define B200SK_MY_LIST B200SK_LIST_1
#elif B200SK_PART == 7
#define B200SK_MY_LIST B200SK_LIST_2
#elif B200SK_PART == 8
#define B200SK_MY_LIST B200SK_LIST_3
#else
#define B200SK_MY_LIST B200SK_CAT(B200SK_LIST_, B200SK_PART)
#endif
#if B200SK_PART >= 5
#define B200SK_IS_SHARD true
#else
#define B200SK_IS_SHARD false
#endif

// One part: launch (or, occ != nullptr, report the occupancy of) the instantiation for (mode, window) if
// this part holds it; *has tells whether it does.
cudaError_t B200SK_CAT(launch_sparse_reg_part, B200SK_PART)(const KArgs &a, int window, int threads, int blocks,
                                                            cudaStream_t st, int *occ, bool *has) {
    *has = true;
#define B200SK_W(MODE, W) if (a.mode == MODE && window == W) return launch_w<MODE, W, B200SK_IS_SHARD>(a, threads, blocks, st, occ);
    B200SK_MY_LIST
#undef B200SK_W
    *has = false;
    return cudaSuccess;
}
bool B200SK_CAT(sparse_reg_has_part, B200SK_PART)(int mode, int window) {
#define B200SK_W(MODE, W) if (mode == MODE && window == W) return true;
    B200SK_MY_LIST
#undef B200SK_W
    return false;
}

#if B200SK_PART == 0
cudaError_t launch_sparse_reg_part1(const KArgs &, int, int, int, cudaStream_t, int *, bool *);
cudaError_t launch_sparse_reg_part2(const KArgs &, int, int, int, cudaStream_t, int *, bool *);
cudaError_t launch_sparse_reg_part3(const KArgs &, int, int, int, cudaStream_t, int *, bool *);
cudaError_t launch_sparse_reg_part4(const KArgs &, int, int, int, cudaStream_t, int *, bool *);
cudaError_t launch_sparse_reg_part5(const KArgs &, int, int, int, cudaStream_t, int *, bool *);
cudaError_t launch_sparse_reg_part6(const KArgs &, int, int, int, cudaStream_t, int *, bool *);
cudaError_t launch_sparse_reg_part7(const KArgs &, int, int, int, cudaStream_t, int *, bool *);
cudaError_t launch_sparse_reg_part8(const KArgs &, int, int, int, cudaStream_t, int *, bool *);
bool sparse_reg_has_part1(int, int);
bool sparse_reg_has_part2(int, int);
bool sparse_reg_has_part3(int, int);
bool sparse_reg_has_part4(int, int);

static int window_of(int mode, int k, int w, int s) { return mode == B200SK_MODE_SYNCMER ? 2 * (k - s) : w; }

// occ != nullptr: only report the occupancy
cudaError_t launch_sparse_reg(const KArgs &a, int threads, int blocks, cudaStream_t st, int *occ) {
    const int window = window_of(a.mode, a.k, a.w, a.s);
    bool has = false;
    if (a.shard.n) { // one tile chain across several GPUs: the SHARD instantiations
        cudaError_t es = launch_sparse_reg_part5(a, window, threads, blocks, st, occ, &has);
        if (has) return es;
        es = launch_sparse_reg_part6(a, window, threads, blocks, st, occ, &has);
        if (has) return es;
        es = launch_sparse_reg_part7(a, window, threads, blocks, st, occ, &has);
        if (has) return es;
        es = launch_sparse_reg_part8(a, window, threads, blocks, st, occ, &has);
        if (has) return es;
        return cudaErrorInvalidValue;
    }
    cudaError_t e = launch_sparse_reg_part0(a, window, threads, blocks, st, occ, &has);
    if (has) return e;
    e = launch_sparse_reg_part1(a, window, threads, blocks, st, occ, &has);
    if (has) return e;
    e = launch_sparse_reg_part2(a, window, threads, blocks, st, occ, &has);
    if (has) return e;
    e = launch_sparse_reg_part3(a, window, threads, blocks, st, occ, &has);
    if (has) return e;
    e = launch_sparse_reg_part4(a, window, threads, blocks, st, occ, &has);
    if (has) return e;
    return cudaErrorInvalidValue;
}

int sparse_reg_max_warps(int mode, int k, int w, int s) {
    // protein minimizers: measured, not derived -- 16 warps 4.08 ms, 14 4.35, 12 3.63, 10 3.81 per 10 M reads
    // (profiles/r02af_warps.txt); an odd number of warps per scheduler pair always loses to the next multiple of four
    if (mode == B200SK_MODE_PROTEIN_MINIMIZER) return 12;
    return mode == B200SK_MODE_SYNCMER && window_of(mode, k, w, s) >= 16 ? 12 : 16; // max_warps<MODE, W>()
}

bool sparse_reg_supported(int mode, int k, int w, int s) {
    if (mode != B200SK_MODE_MINIMIZER && mode != B200SK_MODE_SYNCMER && mode != B200SK_MODE_PROTEIN_MINIMIZER) return false;
    if (mode == B200SK_MODE_PROTEIN_MINIMIZER && k > 16) return false; // the register window holds 16 residues
    const int window = window_of(mode, k, w, s);
    return sparse_reg_has_part0(mode, window) || sparse_reg_has_part1(mode, window) ||
           sparse_reg_has_part2(mode, window) || sparse_reg_has_part3(mode, window) ||
           sparse_reg_has_part4(mode, window);
}
#endif

} // namespace b200sk
