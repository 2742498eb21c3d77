// Downstream consumers of the resident hash arrays: what the tools built on package `sketches`
// (kmcp, unikmer: sketches/README.md:14-15,42) do with the uint64 stream on the host -- keep the FracMinHash
// fraction h <= MaxUint64 / scale (the rule of sketches/iterator.go:180-185,281,443), sort, drop duplicates -- done
// on the device, so that only the reduced sketch crosses PCIe / NVLink.
//
//   k_filter_scale   stream compaction of the values <= max_hash (order not kept: a sort follows)
//   k_radix_hist     per-tile digit histograms, laid out [digit][tile]
//   k_scan_hist      exclusive scan of that table (single pass, decoupled look-back) -> 64-bit bases
//   k_radix_scatter  stable LSD radix pass: the tile is ranked in shared memory (warp match + per-warp digit
//                    counters), reordered there, and leaves in runs of equal digits
//   k_unique         ordered compaction of the first element of every run of equal values (look-back)
//
// HBM-bound integer/byte work: every pass reads and writes the keys once (+ one more read for the histogram);
// digits whose 8 bits are constant zero under the scale filter are skipped.  No tensor cores.
#include <cuda_runtime.h>
#include <stdio.h>

#include <algorithm>

#include "b200sk_device.cuh"
#include "../../include/b200sketch.h"

namespace b200sk {
int ctx_device(b200sk_ctx *ctx);
void ctx_set_error(b200sk_ctx *ctx, const char *msg);
void ctx_add_launches(b200sk_ctx *ctx, uint64_t n);
void **ctx_reduce_slot(b200sk_ctx *ctx);

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // 4096 keys per tile
constexpr int SC_ITEMS = 8;                        // scan: counters per thread
constexpr int SC_TILE = RS_THREADS * SC_ITEMS;

struct ReduceState {
    void *tmp = nullptr; size_t tmp_cap = 0;       // the other half of the sort's double buffer
    void *hist = nullptr; size_t hist_cap = 0;     // uint32 [256][tiles]
    void *base = nullptr; size_t base_cap = 0;     // uint64 [256][tiles]
    void *state = nullptr; size_t state_cap = 0;   // look-back words
    unsigned long long *meta = nullptr;            // [0] ticket, [1] count
    unsigned long long *h_meta = nullptr;          // pinned
};

cudaError_t reserve(void *&p, size_t &cap, size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t w = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, w);
    if (e == cudaSuccess) cap = w;
    return e;
}

// ---------------------------------------------------------------- FracMinHash filter
__global__ void __launch_bounds__(256) k_filter_scale(const uint64_t *__restrict__ in, uint64_t n, uint64_t max_hash,
                                                      uint64_t *__restrict__ out, uint64_t capacity,
                                                      unsigned long long *count) {
    __shared__ uint32_t warp_cnt[8];
    __shared__ unsigned long long s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint64_t tiles = (n + 2047) / 2048;
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        uint64_t v[8];
        uint32_t keep = 0, mine = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint64_t idx = t * 2048 + (uint64_t)i * 256 + tid;
            v[i] = idx < n ? in[idx] : ~0ULL;
            const bool k = idx < n && v[i] <= max_hash;
            keep |= (k ? 1u : 0u) << i;
            mine += k;
        }
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += x;
        }
        if (lane == 31) warp_cnt[wid] = inc;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = warp_cnt[w]; warp_cnt[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(count, (unsigned long long)tot) : 0ULL;
        }
        __syncthreads();
        uint64_t o = s_base + warp_cnt[wid] + inc - mine;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if ((keep >> i) & 1u) {
                if (o < capacity) out[o] = v[i];
                o++;
            }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- LSD radix sort, 8 bits per pass
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const uint64_t *__restrict__ in, uint64_t n, int shift,
                                                           uint32_t *__restrict__ hist, uint64_t tiles) {
    __shared__ uint32_t h[256];
    const uint32_t tid = threadIdx.x;
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        h[tid] = 0;
        __syncthreads();
        const uint64_t b = t * RS_TILE;
#pragma unroll 4
        for (int i = 0; i < RS_ITEMS; i++) {
            const uint64_t idx = b + (uint64_t)i * RS_THREADS + tid;
            if (idx < n) atomicAdd(&h[(uint32_t)(in[idx] >> shift) & 255u], 1u);
        }
        __syncthreads();
        hist[(uint64_t)tid * tiles + t] = h[tid];
        __syncthreads();
    }
}

// exclusive scan of `m` uint32 counters into uint64 bases, tiles of SC_TILE handed out by a ticket
__global__ void __launch_bounds__(RS_THREADS) k_scan_hist(const uint32_t *__restrict__ in, uint64_t m, uint64_t *__restrict__ out,
                                                          uint64_t *state, unsigned long long *ticket) {
    __shared__ uint32_t warp_sums[34];
    __shared__ uint64_t s_tile, s_base;
    const uint32_t tid = threadIdx.x;
    const uint64_t tiles = (m + SC_TILE - 1) / SC_TILE;
    for (;;) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1ULL);
        __syncthreads();
        const uint64_t t = s_tile;
        if (t >= tiles) break;
        uint32_t v[SC_ITEMS], sum = 0;
        const uint64_t b = t * SC_TILE + (uint64_t)tid * SC_ITEMS;
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) {
            v[i] = b + i < m ? in[b + i] : 0u;
            sum += v[i];
        }
        uint32_t total;
        const uint32_t excl = block_excl_scan(sum, warp_sums, &total);
        if (tid < 32) {
            const uint64_t p = lookback_exclusive(state, t, total);
            if (tid == 0) s_base = p;
        }
        __syncthreads();
        uint64_t run = s_base + excl;
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) {
            if (b + i < m) out[b + i] = run;
            run += v[i];
        }
        __syncthreads();
    }
}

// One stable pass.  Tile order = index order: warp w holds keys [w*512, w*512+512) of the tile, in 16 rows of 32.
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const uint64_t *__restrict__ in, uint64_t n, int shift,
                                                              const uint64_t *__restrict__ base, uint64_t tiles,
                                                              uint64_t *__restrict__ out) {
    __shared__ uint64_t keys[RS_TILE];
    __shared__ uint32_t wcnt[8][256];   // per warp: keys of each digit seen so far / exclusive over warps afterwards
    __shared__ uint32_t binstart[256];  // first slot of each digit inside the tile
    __shared__ uint64_t gbase[256];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        for (int w = 0; w < 8; w++) wcnt[w][tid] = 0;
        __syncthreads();
        const uint64_t b = t * RS_TILE + (uint64_t)wid * (RS_TILE / 8);
        uint64_t k[RS_ITEMS];
        uint32_t rank[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const uint64_t idx = b + (uint64_t)i * 32 + lane;
            const bool valid = idx < n;
            k[i] = valid ? in[idx] : ~0ULL;
            // keys past the end sort as digit 255 behind every real key of that digit (they come last in the tile)
            const uint32_t d = valid ? (uint32_t)(k[i] >> shift) & 255u : 255u;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t before = __popc(peers & ((1u << lane) - 1u));
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if ((int)lane == leader) { old = wcnt[wid][d]; wcnt[wid][d] = old + __popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[i] = old + before;
            __syncwarp();
        }
        __syncthreads();
        { // digit `tid`: exclusive over the warps, tile histogram, then the exclusive scan over digits
            uint32_t run = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = wcnt[w][tid]; wcnt[w][tid] = run; run += c; }
            __shared__ uint32_t warp_sums[34];
            uint32_t total;
            const uint32_t ex = block_excl_scan(run, warp_sums, &total);
            binstart[tid] = ex;
            gbase[tid] = base[(uint64_t)tid * tiles + t];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const uint64_t idx = b + (uint64_t)i * 32 + lane;
            const uint32_t d = idx < n ? (uint32_t)(k[i] >> shift) & 255u : 255u;
            keys[binstart[d] + wcnt[wid][d] + rank[i]] = k[i];
        }
        __syncthreads();
        const uint32_t valid_n = (uint32_t)min((uint64_t)RS_TILE, n - t * RS_TILE);
        for (uint32_t s = tid; s < valid_n; s += RS_THREADS) {
            const uint64_t v = keys[s];
            const uint32_t d = (uint32_t)(v >> shift) & 255u;
            out[gbase[d] + (s - binstart[d])] = v;
        }
        __syncthreads();
    }
}

// first element of every run of equal values, order kept
__global__ void __launch_bounds__(RS_THREADS) k_unique(const uint64_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out,
                                                       uint64_t capacity, uint64_t *state, unsigned long long *ticket,
                                                       unsigned long long *count) {
    __shared__ uint32_t warp_sums[34];
    __shared__ uint64_t s_tile, s_base;
    const uint32_t tid = threadIdx.x;
    const uint64_t tiles = (n + SC_TILE - 1) / SC_TILE;
    for (;;) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1ULL);
        __syncthreads();
        const uint64_t t = s_tile;
        if (t >= tiles) break;
        const uint64_t b = t * SC_TILE + (uint64_t)tid * SC_ITEMS;
        uint64_t v[SC_ITEMS];
        uint64_t prev = b > 0 && b - 1 < n ? in[b - 1] : 0;
        uint32_t keep = 0, mine = 0;
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) {
            v[i] = b + i < n ? in[b + i] : 0;
            const bool k = b + i < n && (b + i == 0 || v[i] != prev);
            prev = v[i];
            keep |= (k ? 1u : 0u) << i;
            mine += k;
        }
        uint32_t total;
        const uint32_t excl = block_excl_scan(mine, warp_sums, &total);
        if (tid < 32) {
            const uint64_t p = lookback_exclusive(state, t, total);
            if (tid == 0) {
                s_base = p;
                if (t + 1 == tiles) *count = p + total;
            }
        }
        __syncthreads();
        uint64_t o = s_base + excl;
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++)
            if ((keep >> i) & 1u) {
                if (o < capacity) out[o] = v[i];
                o++;
            }
        __syncthreads();
    }
}

int rfail(b200sk_ctx *ctx, cudaError_t e, const char *what) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    ctx_set_error(ctx, buf);
    return B200SK_ERR_CUDA;
}
#define RCK(call)                                           \
    do {                                                    \
        cudaError_t _e = (call);                            \
        if (_e != cudaSuccess) return rfail(ctx, _e, #call); \
    } while (0)

ReduceState *state_of(b200sk_ctx *ctx) {
    void **slot = ctx_reduce_slot(ctx);
    if (!*slot) *slot = new ReduceState();
    return static_cast<ReduceState *>(*slot);
}

} // namespace

// append the values <= max_hash to out[*count ...] (count keeps running across calls)
cudaError_t launch_filter_scale(const uint64_t *in, uint64_t n, uint64_t max_hash, uint64_t *out, uint64_t capacity,
                                unsigned long long *count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<uint64_t>((n + 2047) / 2048, 148ull * 8);
    k_filter_scale<<<blocks, 256, 0, st>>>(in, n, max_hash, out, capacity, count);
    return cudaGetLastError();
}

void reduce_free(void *p) {
    ReduceState *s = static_cast<ReduceState *>(p);
    if (!s) return;
    for (void *q : {s->tmp, s->hist, s->base, s->state, (void *)s->meta})
        if (q) cudaFree(q);
    if (s->h_meta) cudaFreeHost(s->h_meta);
    delete s;
}

} // namespace b200sk

using namespace b200sk;

extern "C" {

uint64_t b200sk_scale_max_hash(uint32_t scale) { return scale <= 1 ? ~0ULL : ~0ULL / scale; }

// d_val[0..n) -> d_out: the values <= MaxUint64/scale (scale <= 1: all), sorted ascending, duplicates dropped
// when `unique`.  d_val is used as scratch (its content is lost); d_out may not alias d_val.  *n_out = elements
// produced (or needed, with B200SK_ERR_CAPACITY).  Synchronises `stream` once.
int b200sk_reduce_device(b200sk_ctx *ctx, uint64_t *d_val, uint64_t n, uint32_t scale, int unique, uint64_t *d_out,
                         uint64_t capacity, uint64_t *n_out, void *stream) {
    if (!ctx || (n && (!d_val || !d_out)) || d_val == d_out) return B200SK_ERR_BAD_ARG;
    RCK(cudaSetDevice(ctx_device(ctx)));
    cudaStream_t st = (cudaStream_t)stream;
    ReduceState *rs = state_of(ctx);
    if (!rs->meta) {
        RCK(cudaMalloc((void **)&rs->meta, 64));
        RCK(cudaHostAlloc((void **)&rs->h_meta, 64, cudaHostAllocDefault));
    }
    if (n_out) *n_out = 0;
    if (n == 0) return 0;
    uint64_t launches = 0;
    uint64_t m = n;            // live elements
    uint64_t *cur = d_val;     // where they are
    const uint64_t max_hash = b200sk_scale_max_hash(scale);
    RCK(reserve(rs->tmp, rs->tmp_cap, 8)); // placeholder so that tmp is never null
    if (scale > 1) {
        // the filter compacts d_val into d_out (capacity permitting), the sort then ping-pongs between the two
        RCK(cudaMemsetAsync(rs->meta, 0, 64, st));
        const unsigned blocks = (unsigned)std::min<uint64_t>((n + 2047) / 2048, 148ull * 8);
        k_filter_scale<<<blocks, 256, 0, st>>>(d_val, n, max_hash, d_out, capacity, rs->meta + 1);
        launches++;
        RCK(cudaMemcpyAsync(rs->h_meta, rs->meta, 64, cudaMemcpyDeviceToHost, st));
        RCK(cudaStreamSynchronize(st));
        m = rs->h_meta[1];
        if (m > capacity) {
            if (n_out) *n_out = m;
            return B200SK_ERR_CAPACITY;
        }
        cur = d_out;
    }
    uint64_t *other = cur == d_out ? d_val : d_out;
    if (cur == d_val && capacity < m) { // sorting d_val in place of d_out needs room for every element
        if (n_out) *n_out = m;
        return B200SK_ERR_CAPACITY;
    }
    if (m > 1) {
        const uint64_t tiles = (m + RS_TILE - 1) / RS_TILE;
        const uint64_t cells = tiles * 256;
        RCK(reserve(rs->hist, rs->hist_cap, cells * 4));
        RCK(reserve(rs->base, rs->base_cap, cells * 8));
        const uint64_t sc_tiles = (cells + SC_TILE - 1) / SC_TILE;
        RCK(reserve(rs->state, rs->state_cap, (std::max<uint64_t>(sc_tiles, (m + SC_TILE - 1) / SC_TILE) + 1) * 8));
        const unsigned blocks = (unsigned)std::min<uint64_t>(tiles, 148ull * 8);
        // digits that are all zero under the filter need no pass
        int top_bits_zero = 0;
        for (uint64_t x = max_hash; !(x >> 63) && top_bits_zero < 64; x <<= 1) top_bits_zero++;
        const int passes = 8 - top_bits_zero / 8;
        for (int ps = 0; ps < passes; ps++) {
            const int shift = ps * 8;
            k_radix_hist<<<blocks, RS_THREADS, 0, st>>>(cur, m, shift, (uint32_t *)rs->hist, tiles);
            RCK(cudaMemsetAsync(rs->state, 0, (sc_tiles + 1) * 8, st));
            RCK(cudaMemsetAsync(rs->meta, 0, 8, st));
            k_scan_hist<<<(unsigned)std::min<uint64_t>(sc_tiles, 148ull * 4), RS_THREADS, 0, st>>>(
                (const uint32_t *)rs->hist, cells, (uint64_t *)rs->base, (uint64_t *)rs->state, rs->meta);
            k_radix_scatter<<<blocks, RS_THREADS, 0, st>>>(cur, m, shift, (const uint64_t *)rs->base, tiles, other);
            launches += 3;
            std::swap(cur, other);
        }
        RCK(cudaGetLastError());
    }
    uint64_t produced = m;
    if (unique && m > 1) {
        // the sorted run sits in `cur`; the distinct values go to the other buffer
        const uint64_t u_tiles = (m + SC_TILE - 1) / SC_TILE;
        RCK(reserve(rs->state, rs->state_cap, (u_tiles + 1) * 8));
        RCK(cudaMemsetAsync(rs->state, 0, (u_tiles + 1) * 8, st));
        RCK(cudaMemsetAsync(rs->meta, 0, 64, st));
        const uint64_t cap_other = other == d_out ? capacity : n;
        k_unique<<<(unsigned)std::min<uint64_t>(u_tiles, 148ull * 4), RS_THREADS, 0, st>>>(
            cur, m, other, cap_other, (uint64_t *)rs->state, rs->meta, rs->meta + 1);
        launches++;
        RCK(cudaMemcpyAsync(rs->h_meta, rs->meta, 64, cudaMemcpyDeviceToHost, st));
        RCK(cudaStreamSynchronize(st));
        produced = rs->h_meta[1];
        std::swap(cur, other);
        if (produced > cap_other) {
            if (n_out) *n_out = produced;
            return B200SK_ERR_CAPACITY;
        }
    }
    if (cur != d_out) { // an odd number of hops left the result in d_val
        if (produced > capacity) {
            if (n_out) *n_out = produced;
            return B200SK_ERR_CAPACITY;
        }
        RCK(cudaMemcpyAsync(d_out, cur, produced * 8, cudaMemcpyDeviceToDevice, st));
    }
    RCK(cudaStreamSynchronize(st));
    ctx_add_launches(ctx, launches);
    if (n_out) *n_out = produced;
    return 0;
}

} // extern "C"
