// Kernel argument block + per-read geometry shared by host planner and kernels.
#pragma once
#include <stdint.h>

#include "../../include/b200sketch.h"

namespace b200sk {

// flags word written by the kernels
#define B200SK_FLAG_CAPACITY 1u /* output capacity exceeded: values not written       */
#define B200SK_FLAG_SPAN 2u     /* a read was longer than the max_read_len hint        */

struct KArgs {
    const uint8_t *bases;
    const uint64_t *off;      // n_reads+1 (circular: offsets into the extended copy)
    const uint64_t *off_orig; // circular only: original offsets (length checks use the un-extended length)
    uint64_t n_reads;
    const uint64_t *item_first; // chunked: exclusive scan of chunks per read (n_reads+1); else nullptr
    uint64_t n_items;           // == item_first[n_reads] when chunked (read on device), else n_reads
    const uint64_t *n_items_dev;
    uint64_t *out_val;
    uint32_t *out_pos;
    uint64_t *out_off;
    int32_t *status;
    uint64_t capacity;
    uint64_t out_base; // added to every out_off entry (host pipeline: running total of earlier sub-batches)
    uint64_t *tile_state;
    unsigned long long *ticket;
    uint32_t *flags;
    const uint32_t *ill; // MODE_KMER: first illegal base per read (0xffffffff = none)
    const uint8_t *aux;  // MODE_PROTEIN: codon matrix (4096) + base2code (256) + pair LUT (256); KMER: luts
    int32_t mode, k, w, s, canonical, frame;
    uint32_t C;        // positions per chunk
    uint32_t span_max; // max bases one item touches
    uint32_t lcap;     // staged outputs per item (sparse modes)
    // dynamic shared memory layout (byte offsets)
    uint32_t sm_tile, sm_tile_bytes, sm_ring, sm_ring_bytes, sm_listv, sm_listp, sm_total;
};

// Number of output-candidate positions of a read of (extended) length L, and the
// status the reference constructor returns for it.  orig = un-extended length.
//   NTHASH  : k-mers                      iterator.go:616-621
//   KMER    : k-mers                      iterator.go:669-674
//   MINIMIZER: windows = L-k-w+2          sketch.go:86-94 (length check on the un-extended length)
//   SYNCMER : idx in [0, end], end=L-2k+s+1   sketch.go:143-151,173
__host__ __device__ inline uint32_t read_positions(int mode, uint64_t L, uint64_t orig, int k, int w, int s,
                                                   int32_t *status) {
    *status = B200SK_OK;
    switch (mode) {
    case B200SK_MODE_KMER:
    case B200SK_MODE_NTHASH:
        if (orig < (uint64_t)k) { *status = B200SK_ERR_SHORT_SEQ; return 0; }
        return (uint32_t)(L - (uint64_t)k + 1);
    case B200SK_MODE_MINIMIZER:
        if (orig < (uint64_t)k + (uint64_t)w - 1) { *status = B200SK_ERR_SHORT_SEQ; return 0; }
        return (uint32_t)(L - (uint64_t)k - (uint64_t)w + 2);
    case B200SK_MODE_SYNCMER: {
        const int64_t need = 2 * (int64_t)k - s - 1;
        if ((int64_t)orig < need || L < (uint64_t)k) { *status = B200SK_ERR_SHORT_SEQ; return 0; }
        return (uint32_t)(L - 2 * (uint64_t)k + (uint64_t)s + 2);
    }
    default: return 0;
    }
}

__host__ __device__ inline uint32_t chunks_of(uint32_t npos, uint32_t C) {
    return npos == 0 ? 1u : (npos + C - 1) / C;
}

} // namespace b200sk
