// Protein pieces shared by the dense protein kernel and the protein-minimizer kernel:
// wyhash (zeebo/wyhash v0.0.1) and the codon lookup of CodonTable.Get.
#pragma once
#include "b200sk_device.cuh"

namespace b200sk {

// ------------------------------------------------------------------ wyhash (zeebo/wyhash v0.0.1 Hash(b, seed))
// Published wyhash v1 layout: 32-byte blocks, tail by len & 31, final mum(seed, len ^ p5).  The tail
// reads 8 bytes as two 32-bit halves with the first half high.  Reference parity of this function is
// unpinned (no reference test checks a protein hash value); it is bit-exact against oracle/.
#define WYP0 0xa0761d6478bd642fULL
#define WYP1 0xe7037ed1a0b428dbULL
#define WYP2 0x8ebc6af09c88c6e3ULL
#define WYP3 0x589965cc75374cc3ULL
#define WYP4 0x1d8e4e27c47d124fULL
#define WYP5 0xeb44accab455d165ULL
__device__ __forceinline__ uint64_t wymum(uint64_t a, uint64_t b) { return __umul64hi(a, b) ^ (a * b); }

struct ByteSrc { // little-endian reads of an unaligned byte string in shared memory
    const uint8_t *p;
    __device__ __forceinline__ uint64_t r8(uint32_t i) const { return p[i]; }
    __device__ __forceinline__ uint64_t r16(uint32_t i) const { return r8(i) | (r8(i + 1) << 8); }
    __device__ __forceinline__ uint64_t r32(uint32_t i) const { return r16(i) | (r16(i + 2) << 16); }
    __device__ __forceinline__ uint64_t r64(uint32_t i) const { return r32(i) | (r32(i + 4) << 32); }
    __device__ __forceinline__ uint64_t r64s(uint32_t i) const { return (r32(i) << 32) | r32(i + 4); }
};

__device__ __forceinline__ uint64_t wy_tail_word(const ByteSrc &s, uint32_t o, uint32_t n) { // n in 1..8
    switch (n) {
    case 1: return s.r8(o);
    case 2: return s.r16(o);
    case 3: return (s.r16(o) << 8) | s.r8(o + 2);
    case 4: return s.r32(o);
    case 5: return (s.r32(o) << 8) | s.r8(o + 4);
    case 6: return (s.r32(o) << 16) | s.r16(o + 4);
    case 7: return (s.r32(o) << 24) | (s.r16(o + 4) << 8) | s.r8(o + 6);
    default: return s.r64s(o);
    }
}

__device__ __forceinline__ uint64_t wyhash_dev(const ByteSrc &s, uint32_t len, uint64_t seed) {
    uint32_t o = 0;
    for (uint32_t i = 0; i + 32 <= len; i += 32, o += 32)
        seed = wymum(seed ^ WYP0, wymum(s.r64(o) ^ WYP1, s.r64(o + 8) ^ WYP2) ^
                                      wymum(s.r64(o + 16) ^ WYP3, s.r64(o + 24) ^ WYP4));
    seed ^= WYP0;
    const uint32_t t = len & 31u;
    if (t == 0) {
    } else if (t <= 8) {
        seed = wymum(seed, wy_tail_word(s, o, t) ^ WYP1);
    } else if (t <= 16) {
        seed = wymum(s.r64s(o) ^ seed, wy_tail_word(s, o + 8, t - 8) ^ WYP2);
    } else if (t <= 24) {
        seed = wymum(s.r64s(o) ^ seed, s.r64s(o + 8) ^ WYP2) ^ wymum(seed, wy_tail_word(s, o + 16, t - 16) ^ WYP3);
    } else {
        seed = wymum(s.r64s(o) ^ seed, s.r64s(o + 8) ^ WYP2) ^
               wymum(s.r64s(o + 16) ^ seed, wy_tail_word(s, o + 24, t - 24) ^ WYP4);
    }
    return wymum(seed, (uint64_t)len ^ WYP5);
}

// wyhash (zeebo/wyhash v0.0.1 layout, b200sk_protein.cuh) of the k <= 16 newest bytes of a 128-bit window whose
// top byte is the newest: the k-mer's bytes are the top k bytes of (whi:wlo).
__device__ __forceinline__ uint64_t wy_tail_of(uint64_t v, uint32_t n) { // v: n bytes, first byte lowest; n in 1..8
    switch (n) {
    case 1: case 2: case 4: return v;
    case 3: return ((v & 0xffffull) << 8) | ((v >> 16) & 0xffull);
    case 5: return ((v & 0xffffffffull) << 8) | ((v >> 32) & 0xffull);
    case 6: return ((v & 0xffffffffull) << 16) | ((v >> 32) & 0xffffull);
    case 7: return ((v & 0xffffffffull) << 24) | (((v >> 32) & 0xffffull) << 8) | ((v >> 48) & 0xffull);
    default: return (v << 32) | (v >> 32);
    }
}
__device__ __forceinline__ uint64_t wyhash_window(uint64_t wlo, uint64_t whi, uint32_t k) {
    const uint64_t seed = 1ull ^ WYP0;
    uint64_t h;
    if (k <= 8) {
        const uint64_t v = whi >> (8u * (8u - k));
        h = wymum(seed, wy_tail_of(v, k) ^ WYP1);
    } else {
        const uint32_t sft = 8u * (16u - k); // < 64
        const uint64_t first8 = sft ? ((wlo >> sft) | (whi << (64u - sft))) : wlo;
        const uint64_t rest = whi >> sft;
        h = wymum(((first8 << 32) | (first8 >> 32)) ^ seed, wy_tail_of(rest, k - 8u) ^ WYP2);
    }
    return wymum(h, (uint64_t)k ^ WYP5);
}

// The same for a k known only at run time, without a switch per k-mer: the tail word's byte shuffle (wy_tail_of) is a
// permutation of the word's bytes with zero fill, i.e. two PRMTs whose selectors depend on the tail length alone --
// computed once per item.  Selector nibble 7 picks the word's top byte, which is zero whenever the tail is shorter
// than 8 bytes.
struct WyPlan {
    uint32_t k, sft, sel_lo, sel_hi;
};
__device__ __forceinline__ WyPlan wy_plan(uint32_t k) {
    WyPlan p;
    p.k = k;
    p.sft = 8u * ((k <= 8u ? 8u : 16u) - k);
    switch (k <= 8u ? k : k - 8u) {
    case 3: p.sel_lo = 0x7102u; p.sel_hi = 0x7777u; break;
    case 5: p.sel_lo = 0x2104u; p.sel_hi = 0x7773u; break;
    case 6: p.sel_lo = 0x1054u; p.sel_hi = 0x7732u; break;
    case 7: p.sel_lo = 0x0546u; p.sel_hi = 0x7321u; break;
    case 8: p.sel_lo = 0x7654u; p.sel_hi = 0x3210u; break;
    default: p.sel_lo = 0x3210u; p.sel_hi = 0x7654u; break; // 1, 2, 4: as is
    }
    return p;
}
__device__ __forceinline__ uint64_t wy_tail_perm(uint64_t v, const WyPlan &p) {
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    return ((uint64_t)__byte_perm(lo, hi, p.sel_hi) << 32) | __byte_perm(lo, hi, p.sel_lo);
}
__device__ __forceinline__ uint64_t wyhash_window(uint64_t wlo, uint64_t whi, const WyPlan &p) {
    const uint64_t seed = 1ull ^ WYP0;
    uint64_t h;
    if (p.k <= 8u) {
        h = wymum(seed, wy_tail_perm(whi >> p.sft, p) ^ WYP1);
    } else {
        const uint64_t first8 = p.sft ? ((wlo >> p.sft) | (whi << (64u - p.sft))) : wlo;
        h = wymum(((first8 << 32) | (first8 >> 32)) ^ seed, wy_tail_perm(whi >> p.sft, p) ^ WYP2);
    }
    return wymum(h, (uint64_t)p.k ^ WYP5);
}


// ------------------------------------------------------------------ codon lookup
// aux layout: [0,4096) matrix[i][j][k] over 4-bit IUPAC codes, [4096,4352) base2code (0xff = invalid),
// [4352,4608) DNA pair letters.  CodonTable.Get: seq/codon_tables.go:152-170 with allowUnknownCodon=true.
__device__ __forceinline__ uint32_t codon_aa(const uint8_t *tab, uint32_t b0, uint32_t b1, uint32_t b2) {
    const uint32_t c0 = tab[4096 + b0], c1 = tab[4096 + b1], c2 = tab[4096 + b2];
    if ((c0 | c1 | c2) & 0x80u) return 'X'; // invalid base, unknown codons allowed
    if (b0 == '-' && b1 == '-' && b2 == '-') return '-';
    const uint32_t aa = tab[(c0 << 8) | (c1 << 4) | c2];
    return aa ? aa : 'X';
}


} // namespace b200sk
