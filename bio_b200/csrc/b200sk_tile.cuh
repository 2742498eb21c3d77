// Tile/item bookkeeping shared by the sketching kernels.
#pragma once
#include "b200sk_device.cuh"
#include "b200sk_kernels.cuh"

namespace b200sk {

struct Item {
    uint64_t r;     // read
    uint64_t gb0;   // global byte offset of the first base the item touches
    uint32_t nb;    // bases touched
    uint32_t nstep; // stream steps (k-mers for minimizer, s-mers for syncmer)
    uint32_t q0;    // read-relative position of stream element 0
    uint32_t p0;    // first owned position (a window starting before it only seeds the de-duplication)
    uint32_t end;   // syncmer: last emittable k-mer position (sketch.go:173)
    int32_t status;
    bool first_chunk, last_item, valid;
};

struct TileCtl {
    uint64_t mbar;
    uint64_t tile;
    uint64_t lo, hi;
    uint64_t base;
    uint32_t warp_sums[34];
    uint32_t any_overflow;
};

// item index -> (read, chunk) -> the stream range the item walks
template <int MODE>
__device__ __forceinline__ void item_geometry(const KArgs &a, uint64_t item, uint64_t n_items, Item &it) {
    it.valid = item < n_items;
    it.nb = 0; it.nstep = 0; it.q0 = 0; it.p0 = 0; it.end = 0; it.status = 0;
    it.first_chunk = false; it.last_item = false; it.gb0 = 0; it.r = 0;
    if (!it.valid) return;
    uint64_t r = item;
    uint32_t c = 0;
    if (a.item_first) {
        uint64_t lo = 0, hi = a.n_reads; // largest r with item_first[r] <= item
        if (a.tile_read) { // the read of item 32 * (item / 32) is known; every read has at least one item
            lo = a.tile_read[item >> 5];
            hi = lo + 32 < a.n_reads ? lo + 32 : a.n_reads;
        }
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (a.item_first[mid] <= item) lo = mid; else hi = mid;
        }
        r = lo;
        c = (uint32_t)(item - a.item_first[r]);
    }
    it.r = r;
    it.first_chunk = c == 0;
    it.last_item = item + 1 == n_items;
    const uint64_t o0 = a.off[r], o1 = a.off[r + 1];
    const uint64_t L = o1 - o0;
    const uint64_t orig = a.off_orig ? a.off_orig[r + 1] - a.off_orig[r] : L;
    const uint32_t np = read_positions(a.geom(), r, L, orig, &it.status);
    it.gb0 = o0;
    if (np == 0) return;
    const uint32_t p0 = c * a.C;
    const uint32_t p1 = min(np, p0 + a.C);
    const uint32_t q0 = p0 - (c > 0 ? 1u : 0u);
    it.p0 = p0;
    it.q0 = q0;
    it.gb0 = o0 + q0;
    if (MODE == B200SK_MODE_MINIMIZER || MODE == B200SK_MODE_PROTEIN_MINIMIZER) {
        it.nstep = p1 - q0 + (uint32_t)a.w - 1;
        it.nb = it.nstep + (uint32_t)a.k - 1;
    } else { // SYNCMER
        it.nstep = p1 - q0 + 2u * (uint32_t)(a.k - a.s) - 1;
        it.nb = it.nstep + (uint32_t)a.s - 1;
        it.end = np - 1;
    }
}

} // namespace b200sk
