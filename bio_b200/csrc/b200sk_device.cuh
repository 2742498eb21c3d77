// Device-side building blocks shared by the sketching kernels (sm_100a only).
//
//  * ntHash-1 seeds / rotations (arithmetic of will-rowe/nthash v0.4.0 as used at
//    sketches/iterator.go:649,659 and sketches/sketch.go:120,179,184,212,319,344,367)
//  * 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) + mbarrier wrappers
//  * single-pass ordered allocation of output ranges across tiles
//    (decoupled look-back over a per-tile status word)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sk {

// ------------------------------------------------------------------ ntHash
__device__ __host__ __forceinline__ uint64_t rol64(uint64_t v, unsigned n) {
    n &= 63u;
    return n ? (v << n) | (v >> (64u - n)) : v;
}
__device__ __host__ __forceinline__ uint64_t ror64(uint64_t v, unsigned n) {
    n &= 63u;
    return n ? (v >> n) | (v << (64u - n)) : v;
}
__device__ __forceinline__ uint64_t rol1(uint64_t v) { return (v << 1) | (v >> 63); }
__device__ __forceinline__ uint64_t ror1(uint64_t v) { return (v >> 1) | (v << 63); }

#define B200SK_SEED_A 0x3c8bfbb395c60474ULL
#define B200SK_SEED_C 0x3193c18562a02b4cULL
#define B200SK_SEED_G 0x20323ed082572324ULL
#define B200SK_SEED_T 0x295549f54be24456ULL

// seedTab[b] of the hasher: nonzero for A,C,G,T,U (either case) and for the
// byte values 1,3,4,5,7 which the complement lookup seedTab[b & 7] lands on.
__device__ __host__ __forceinline__ uint64_t seed_of_byte(unsigned b) {
    switch (b) {
    case 'A': case 'a': case 4: case 5: return B200SK_SEED_A;
    case 'C': case 'c': case 7: return B200SK_SEED_C;
    case 'G': case 'g': case 3: return B200SK_SEED_G;
    case 'T': case 't': case 'U': case 'u': case 1: return B200SK_SEED_T;
    default: return 0;
    }
}
__device__ __host__ __forceinline__ uint64_t fwd_seed(unsigned b) { return seed_of_byte(b & 0xffu); }
__device__ __host__ __forceinline__ uint64_t rev_seed(unsigned b) { return seed_of_byte(b & 7u); }

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk global->shared copy, completion counted in bytes on `bar`.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 1-D bulk shared->global copy (dst/src 16-byte aligned, bytes a multiple of 16), tracked by the issuing thread's
// bulk async-groups: commit, then wait until the sources have been READ (the buffer may be reused) or until the
// copies are complete.
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ ordered allocation
// status word per tile: [63:62] flag, [61:0] value.
#define B200SK_FLAG_EMPTY 0ULL
#define B200SK_FLAG_AGG 1ULL
#define B200SK_FLAG_INC 2ULL
#define B200SK_VAL_MASK ((1ULL << 62) - 1)

__device__ __forceinline__ uint64_t ld_state(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by one full warp.  Publishes `total` for `tile` and returns the sum of
// the totals of all tiles before it (exclusive prefix).  Tiles are handed out by
// a ticket counter, so every predecessor is already running: no deadlock.
__device__ __forceinline__ uint64_t lookback_exclusive(uint64_t *state, uint64_t tile, uint64_t total) {
    const unsigned lane = threadIdx.x & 31u;
    if (tile == 0) {
        if (lane == 0) st_state(state, (B200SK_FLAG_INC << 62) | total);
        return 0;
    }
    if (lane == 0) st_state(state + tile, (B200SK_FLAG_AGG << 62) | total);
    uint64_t excl = 0;
    int64_t idx = (int64_t)tile - 1 - (int64_t)lane;
    while (true) {
        uint64_t v = (B200SK_FLAG_INC << 62); // lanes before tile 0 read as "inclusive 0"
        if (idx >= 0) {
            do {
                v = ld_state(state + idx);
            } while ((v >> 62) == B200SK_FLAG_EMPTY);
        }
        const unsigned inc = __ballot_sync(0xffffffffu, (v >> 62) == B200SK_FLAG_INC);
        uint64_t contrib = v & B200SK_VAL_MASK;
        if (inc) {
            const unsigned first = __ffs(inc) - 1;
            if (lane > first) contrib = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (inc) break;
        idx -= 32;
    }
    if (lane == 0) st_state(state + tile, (B200SK_FLAG_INC << 62) | (excl + total));
    return excl;
}

// The same in two halves, so that work which does not need the prefix can sit between them: publish the
// tile's own total first (later tiles can already add it up), resolve the exclusive prefix afterwards.
__device__ __forceinline__ void lookback_publish(uint64_t *state, uint64_t tile, uint64_t total) {
    if ((threadIdx.x & 31u) == 0)
        st_state(state + tile, ((tile == 0 ? B200SK_FLAG_INC : B200SK_FLAG_AGG) << 62) | total);
}
__device__ __forceinline__ uint64_t lookback_resolve(uint64_t *state, uint64_t tile, uint64_t total, uint32_t spin_ns = 0) {
    const unsigned lane = threadIdx.x & 31u;
    if (tile == 0) return 0;
    uint64_t excl = 0;
    int64_t idx = (int64_t)tile - 1 - (int64_t)lane;
    while (true) {
        uint64_t v = (B200SK_FLAG_INC << 62); // lanes before tile 0 read as "inclusive 0"
        if (idx >= 0) {
            v = ld_state(state + idx);
            while ((v >> 62) == B200SK_FLAG_EMPTY) {
                if (spin_ns) __nanosleep(spin_ns);
                v = ld_state(state + idx);
            }
        }
        const unsigned inc = __ballot_sync(0xffffffffu, (v >> 62) == B200SK_FLAG_INC);
        uint64_t contrib = v & B200SK_VAL_MASK;
        if (inc) {
            const unsigned first = __ffs(inc) - 1;
            if (lane > first) contrib = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (inc) break;
        idx -= 32;
    }
    if (lane == 0) st_state(state + tile, (B200SK_FLAG_INC << 62) | (excl + total));
    return excl;
}

// ------------------------------------------------------------------ the same chain ACROSS GPUs
// One batch sharded over n GPUs in chunks of consecutive tiles (chunk c belongs to rank c mod n): the tiles of all
// ranks form ONE ordered chain, so every rank's flush lands at its exact offset in the root's output arrays.
// Every rank keeps a full copy of the status words and a tile publishes into ALL copies (n-1 posted 8-byte stores
// over NVLink); the look-back itself only ever polls the local copy.  Words carry an epoch so that a step never
// has to wait for the other ranks' memset: [63:62] flag, [61:48] epoch, [47:0] value.
#define B200SK_MAX_RANKS 8
#define B200SK_MVAL_MASK ((1ULL << 48) - 1)
struct PeerStates {
    uint64_t *copy[B200SK_MAX_RANKS]; // [r]: rank r's copy as mapped here ([rank] is the local one)
    uint32_t n, rank, epoch;
    uint32_t poll_ns; // back-off between two polls of a status word that is not there yet ...
    uint32_t poll_free; // ... after this many polls without one
};
__device__ __forceinline__ uint64_t ld_state_sys(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state_sys(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t mstate_word(uint64_t flag, uint32_t epoch, uint64_t value) {
    return (flag << 62) | ((uint64_t)(epoch & 0x3fffu) << 48) | (value & B200SK_MVAL_MASK);
}
__device__ __forceinline__ void mstate_store_all(const PeerStates &ps, uint64_t tile, uint64_t word) {
    for (uint32_t r = 0; r < ps.n; r++) st_state_sys(ps.copy[r] + tile, word);
}
__device__ __forceinline__ void lookback_publish_multi(const PeerStates &ps, uint64_t tile, uint64_t total) {
    if ((threadIdx.x & 31u) == 0)
        mstate_store_all(ps, tile, mstate_word(tile == 0 ? B200SK_FLAG_INC : B200SK_FLAG_AGG, ps.epoch, total));
}
__device__ __forceinline__ uint64_t lookback_resolve_multi(const PeerStates &ps, uint64_t tile, uint64_t total) {
    const unsigned lane = threadIdx.x & 31u;
    if (tile == 0) return 0;
    const uint64_t *state = ps.copy[ps.rank];
    const uint64_t want_epoch = (uint64_t)(ps.epoch & 0x3fffu);
    uint64_t excl = 0;
    int64_t idx = (int64_t)tile - 1 - (int64_t)lane;
    while (true) {
        uint64_t v = mstate_word(B200SK_FLAG_INC, ps.epoch, 0); // lanes before tile 0 read as "inclusive 0"
        if (idx >= 0) {
            v = ld_state_sys(state + idx);
            // a rank that is ahead of the chain waits here with every warp it has: back off between polls, or thousands
            // of warps hammer the few L2 lines of the chain's front -- the very lines the peers' status stores must reach
            // (the first polls go out back to back: a predecessor that is merely a little late -- the only kind of wait
            // there is on one GPU -- should not cost a sleep)
            uint32_t tries = 0;
            while ((v >> 62) == B200SK_FLAG_EMPTY || ((v >> 48) & 0x3fffu) != want_epoch) {
                if (++tries > ps.poll_free) __nanosleep(ps.poll_ns);
                v = ld_state_sys(state + idx);
            }
        }
        const unsigned inc = __ballot_sync(0xffffffffu, (v >> 62) == B200SK_FLAG_INC);
        uint64_t contrib = v & B200SK_MVAL_MASK;
        if (inc) {
            const unsigned first = __ffs(inc) - 1;
            if (lane > first) contrib = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (inc) break;
        idx -= 32;
    }
    if (lane == 0) mstate_store_all(ps, tile, mstate_word(B200SK_FLAG_INC, ps.epoch, excl + total));
    return excl;
}

// Block-wide exclusive scan of one uint32 per thread (blockDim.x <= 1024, multiple of 32).
// warp_sums: shared scratch of >= 33 uint32.  Returns exclusive prefix; *block_total = sum.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *warp_sums, uint32_t *block_total) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < nw ? warp_sums[lane] : 0;
        uint32_t si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= (unsigned)o) si += t;
        }
        if (lane < nw) warp_sums[lane] = si - s;
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    uint32_t excl = warp_sums[wid] + inc - v;
    *block_total = warp_sums[32];
    __syncthreads();
    return excl;
}

} // namespace b200sk
