// Kernel argument block + per-read geometry shared by host planner and kernels.
#pragma once
#include <stdint.h>

#include "../../include/b200sketch.h"
#include "b200sk_device.cuh"

namespace b200sk {

// flags word written by the kernels
#define B200SK_FLAG_CAPACITY 1u /* output capacity exceeded: values not written       */
#define B200SK_FLAG_SPAN 2u     /* a read was longer than the max_read_len hint        */

// What the per-read geometry depends on (a slice of KArgs, also used by the small pre-pass kernels).
struct ReadGeom {
    int32_t mode, k, w, s, frame, canonical, protein_input;
    const uint32_t *ill; // MODE_KMER: first illegal base per read (>= length: none); may be null on the host
};

struct KArgs {
    const uint8_t *bases;
    const uint64_t *off;      // n_reads+1 (circular: offsets into the extended copy)
    const uint64_t *off_orig; // circular only: original offsets (length checks use the un-extended length)
    uint64_t n_reads;
    const uint64_t *item_first; // chunked: exclusive scan of chunks per read (n_reads+1); else nullptr
    const uint64_t *tile_read;  // chunked: read that holds item 32 t, per group of 32 items (may be nullptr)
    uint64_t n_items;           // == item_first[n_reads] when chunked (read on device), else n_reads
    const uint64_t *n_items_dev;
    uint64_t *out_val;
    void *out_pos;      // u8 / u16 / u32 elements (pos_width bytes each)
    uint32_t pos_width; // 1, 2 or 4
    uint64_t *out_off;
    int32_t *status;
    uint64_t capacity;
    uint64_t out_base; // added to every out_off entry (host pipeline: running total of earlier sub-batches)
    uint64_t *tile_state;
    unsigned long long *ticket;
    uint32_t *flags;
    const uint32_t *ill; // MODE_KMER: first illegal base per read (0xffffffff = none)
    const uint8_t *aux;  // MODE_PROTEIN: codon matrix (4096) + base2code (256) + pair LUT (256); KMER: luts
    int32_t mode, k, w, s, canonical, frame, alphabet;
    __host__ __device__ ReadGeom geom() const {
        ReadGeom g;
        g.mode = mode; g.k = k; g.w = w; g.s = s; g.frame = frame; g.canonical = canonical;
        g.protein_input = alphabet == 5; g.ill = ill;
        return g;
    }
    // one batch sharded over several GPUs (b200sk_enqueue_device_sharded): this rank's tiles are chunks of
    // shard_chunk_tiles consecutive tiles of the GLOBAL tile order (chunk c belongs to rank c mod n); out_val /
    // out_pos / out_off / status are the ROOT's arrays, indexed globally.  shard.n == 0: a single GPU.
    PeerStates shard;
    uint32_t shard_chunk_tiles;
    uint64_t shard_n_reads; // reads of the whole batch
    unsigned long long *rewalks; // keyed walk: items handed to the exact walk are counted here (may be null)
    uint32_t keyed;    // minimizers, W <= 16: window minimum on 32-bit keys (b200sk_sparse_reg.cu)
    uint32_t key_mask; // 0xffffffc0, as a run-time value (see make_key)
    uint32_t spin_ns;  // look-back poll interval (0 = busy poll)
    uint32_t skew;     // the longest read is a multiple of 32 bytes: the SKEW instantiation of k_sparse_warp
    unsigned long long *unordered; // timing experiment: allocate output ranges in completion order (null = ordered)
    // all six frames of ProteinIterator in one launch (b200sk_enqueue_device_frames): [i] = frame 1, 2, 3, -1, -2, -3
    uint64_t *fr_val[6];
    const uint64_t *fr_off[6];
    uint32_t C;        // positions per chunk
    uint32_t span_max; // max bases one item touches
    uint32_t lcap;     // staged outputs per item (sparse modes)
    // dynamic shared memory layout (byte offsets)
    uint32_t sm_tile, sm_tile_bytes, sm_ring, sm_ring_bytes, sm_listv, sm_listp, sm_total;
};

// out_pos element store (the width is uniform across the grid)
__device__ __forceinline__ void store_pos(void *base, uint32_t width, uint64_t idx, uint32_t v) {
    if (width == 4) reinterpret_cast<uint32_t *>(base)[idx] = v;
    else if (width == 1) reinterpret_cast<uint8_t *>(base)[idx] = (uint8_t)v;
    else reinterpret_cast<uint16_t *>(base)[idx] = (uint16_t)v;
}

// amino acids a frame of a read of length L translates to (seq/codon_tables.go:219,255)
__host__ __device__ inline uint32_t frame_aa_count(uint64_t L, int frame) {
    const uint64_t f = (uint64_t)(frame < 0 ? -frame : frame);
    return L >= f + 2 ? (uint32_t)((L - f - 2) / 3 + 1) : 0u;
}

// Number of output-candidate positions of a read of (extended) length L, and the status the reference
// returns for it.  orig = un-extended length, r = read index (for ill[]).
//   NTHASH   : k-mers                         iterator.go:616-621
//   KMER     : k-mers before the first one holding an illegal base      iterator.go:669-674,730-748
//   MINIMIZER: windows = L-k-w+2              sketch.go:86-94 (length check on the un-extended length)
//   SYNCMER  : idx in [0, end], end=L-2k+s+1  sketch.go:143-151,173
//   PROTEIN  : amino-acid k-mers of the frame iterator-protein.go:47-52,70
__host__ __device__ inline uint32_t read_positions(const ReadGeom &g, uint64_t r, uint64_t L, uint64_t orig,
                                                   int32_t *status) {
    *status = B200SK_OK;
    const int k = g.k;
    switch (g.mode) {
    case B200SK_MODE_NTHASH:
    case B200SK_MODE_SIMHASH: // one SimHash per k-mer (iterator.go:128,151)
        if (orig < (uint64_t)k) { *status = B200SK_ERR_SHORT_SEQ; return 0; }
        return (uint32_t)(L - (uint64_t)k + 1);
    case B200SK_MODE_KMER: {
        if (orig < (uint64_t)k) { *status = B200SK_ERR_SHORT_SEQ; return 0; }
        const uint32_t np = (uint32_t)(L - (uint64_t)k + 1);
        if (g.ill) {
            const uint32_t bad = g.ill[r];
            if ((uint64_t)bad < orig) {
                *status = B200SK_ERR_ILLEGAL_BASE;
                return bad >= (uint32_t)k - 1 ? bad - ((uint32_t)k - 1) : 0u;
            }
        }
        return np;
    }
    case B200SK_MODE_MINIMIZER:
        if (orig < (uint64_t)k + (uint64_t)g.w - 1) { *status = B200SK_ERR_SHORT_SEQ; return 0; }
        return (uint32_t)(L - (uint64_t)k - (uint64_t)g.w + 2);
    case B200SK_MODE_SYNCMER: {
        const int64_t need = 2 * (int64_t)k - g.s - 1;
        if ((int64_t)orig < need || L < (uint64_t)k) { *status = B200SK_ERR_SHORT_SEQ; return 0; }
        return (uint32_t)(L - 2 * (uint64_t)k + (uint64_t)g.s + 2);
    }
    case B200SK_MODE_PROTEIN: {
        if (orig < 3ull * (uint64_t)k) { *status = B200SK_ERR_SHORT_SEQ; return 0; }
        const uint32_t naa = g.protein_input ? (uint32_t)L : frame_aa_count(L, g.frame);
        return naa >= (uint32_t)k ? naa - (uint32_t)k + 1 : 0u;
    }
    case B200SK_MODE_PROTEIN_MINIMIZER: {
        // sketch-protein.go:66,73: both length checks look at the un-translated length; here L is the
        // amino-acid length (the frame was translated by k_translate) and orig the record's length
        if (orig < 3ull * (uint64_t)k || orig < 3ull * (uint64_t)k + (uint64_t)g.w - 1) {
            *status = B200SK_ERR_SHORT_SEQ;
            return 0;
        }
        const uint64_t nk = L >= (uint64_t)k ? L - (uint64_t)k + 1 : 0;
        return nk >= (uint64_t)g.w ? (uint32_t)(nk - (uint64_t)g.w + 1) : 0u;
    }
    case 100: // internal: amino acids of the frame (k_translate's output length)
        return frame_aa_count(L, g.frame);
    default: return 0;
    }
}

// elements the read emits in a dense mode
__host__ __device__ inline uint64_t dense_count(const ReadGeom &g, uint32_t np, int32_t status) {
    if (g.mode == B200SK_MODE_KMER && !g.canonical && status == B200SK_OK) return 2ull * np;
    return np;
}

__host__ __device__ inline uint32_t chunks_of(uint32_t npos, uint32_t C) {
    return npos == 0 ? 1u : (npos + C - 1) / C;
}

} // namespace b200sk
