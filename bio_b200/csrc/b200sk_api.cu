// Host side of libb200sketch.so: context, planning, the device entry points and
// the pipelined host entry point.  C ABI declared in include/b200sketch.h.
// No CPU fallback exists anywhere in this file: without a device every entry
// point fails with B200SK_ERR_NO_DEVICE / B200SK_ERR_CUDA.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "b200sk_kernels.cuh"

namespace b200sk {
cudaError_t launch_prepass(const KArgs &a, unsigned long long *meta, cudaStream_t st);
cudaError_t launch_scan_items(const KArgs &a, uint64_t *item_first, uint64_t *tile_state,
                              unsigned long long *ticket, cudaStream_t st, uint64_t *tile_read);
cudaError_t launch_main(const KArgs &a, int threads, int blocks, cudaStream_t st);
int main_kernel_occupancy(const KArgs &a, int threads);
cudaError_t launch_sparse_reg(const KArgs &a, int threads, int blocks, cudaStream_t st, int *occ);
bool sparse_reg_supported(int mode, int k, int w, int s);
int sparse_reg_max_warps(int mode, int k, int w, int s);
cudaError_t launch_scan_counts(const KArgs &a, uint64_t *tile_state, unsigned long long *ticket, cudaStream_t st);
cudaError_t launch_first_illegal(const uint8_t *bases, const uint64_t *off, uint64_t n_reads, uint32_t *ill,
                                 cudaStream_t st, uint64_t n_bases);
cudaError_t launch_dense(const KArgs &a, int threads, int blocks, cudaStream_t st, int *occ);
bool build_codon_aux(int id, uint8_t *aux);
cudaError_t launch_protein6_warp(const KArgs &a, cudaStream_t st); // b200sk_nthash.cu
bool nthash_warp_fits(uint32_t span_max);
cudaError_t launch_filter_scale(const uint64_t *in, uint64_t n, uint64_t max_hash, uint64_t *out, uint64_t capacity,
                                unsigned long long *count, cudaStream_t st);
cudaError_t launch_scan_geom(const uint64_t *off, uint64_t n_reads, const ReadGeom &g, uint64_t *out,
                             uint64_t *tile_state, unsigned long long *ticket, cudaStream_t st);
cudaError_t launch_translate(const uint8_t *bases, const uint64_t *off, const uint64_t *aa_off, uint64_t n_reads,
                             int frame, const uint8_t *aux, uint8_t *aa, cudaStream_t st);
cudaError_t launch_circularize(const uint8_t *bases, const uint64_t *off, uint64_t n_reads, int k,
                               uint8_t *bases2, uint64_t *off2, uint64_t *tile_state,
                               unsigned long long *ticket, cudaStream_t st);
} // namespace b200sk

using namespace b200sk;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return e;
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostBuf { // pinned
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes, bool keep = false) {
        if (bytes <= cap) return cudaSuccess;
        size_t want = bytes + bytes / 8 + 256;
        void *q = nullptr;
        cudaError_t e = cudaHostAlloc(&q, want, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        if (keep && p && cap) memcpy(q, p, cap);
        if (p) cudaFreeHost(p);
        p = q;
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct Plan {
    int T = 128;
    bool chunked = false;
    bool reg = false;   // window state in registers (b200sk_sparse_reg.cu)
    bool keyed = false; // ... and the window minimum on 32-bit keys (minimizers, w <= 16)
    bool dense = false; // b200sk_dense.cu
    int ctas_per_sm = 1;
    uint32_t C = 0, span_max = 0, lcap = 0;
    uint32_t sm_tile = 0, sm_tile_bytes = 0, sm_ring = 0, sm_ring_bytes = 0, sm_listv = 0, sm_listp = 0,
             sm_total = 0;
};

const uint32_t kChunk = 252;        // positions per chunk for long reads: 63 words, so the 32 lanes of a warp (consecutive
                                    // chunks of one read) start on 32 different shared-memory banks
const uint32_t kChunkReg = 132;     // ... for the register-window kernels (shared memory per lane is the limit) -- the fallback: make_plan picks C
// tuning knob, off by default: from this many ranks on the sharded chain stages a WHOLE tile per bulk store (make_plan)
static const int kWholeTileRanks = [] { const char *e = getenv("B200SK_WHOLE_TILE_RANKS"); return e ? atoi(e) : 1000; }();
const uint32_t kSingleMaxLen = 384; // reads up to this length are one item each
const uint32_t kSmemCtl = 14336 + 256;
const uint32_t kSmemLimit = 227 * 1024;

inline uint32_t up16(uint32_t v) { return (v + 15u) & ~15u; }

// testing knob: B200SK_WALKER=keyed (or rewalk / norewalk, its two diagnostic variants) runs minimizer batches with
// w in {3, 5, 11, 15} through the keyed window minimum (b200sk_sparse_reg.cu) instead of the 64-bit register window.
// Both are bit-exact; the keyed walk is not faster and stays off by default.
bool keyed_walk_enabled() {
    static const bool on = [] {
        const char *e = getenv("B200SK_WALKER");
        return e && (strcmp(e, "keyed") == 0 || strcmp(e, "rewalk") == 0 || strcmp(e, "norewalk") == 0);
    }();
    return on;
}

} // namespace

struct b200sk_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;  // compute stream of the host path
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    DevBuf meta;       // [0] ticket, [1] flags, [2] n_items, [3] max_len, [4] scan ticket, [5] exact re-walks  (u64 each)
    DevBuf tile_state; // main kernel look-back words
    DevBuf scan_state; // item scan look-back words
    DevBuf item_first, tile_read;
    DevBuf circ_bases, circ_off;
    DevBuf ill, aux, aa_bases, aa_off;
    int aux_table = -1;
    std::vector<uint8_t> aux_host;
    // host path device buffers
    DevBuf d_bases, d_off, d_val, d_pos, d_ooff, d_status;       // slot 0
    DevBuf d_bases2, d_off2, d_val2, d_pos2, d_ooff2, d_status2; // slot 1
    HostBuf h_val, h_pos, h_ooff, h_status, h_meta;
    DevBuf d_acc, d_acc2, d_acc_cnt; // b200sk_run_reduced: accumulation array, sort buffer, kept-element counter
    uint64_t acc_cap_hint = 0;      // ... and the size a denser-than-planned batch asked for
    void *fx = nullptr; // record feeder state (b200sk_fastx.cu)
    void *reduce = nullptr; // sort / unique workspace (b200sk_reduce.cu)
    uint64_t launches = 0;
    // the scratch words above belong to ONE batch at a time: a batch enqueued on another stream than the batch before
    // it waits for that batch on the device (enqueue), so two streams of one context cannot race on them
    cudaEvent_t last_done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    std::string last_error;
};

// hooks for b200sk_fastx.cu
namespace b200sk {
void fx_free(void *p);
int ctx_device(b200sk_ctx *ctx) { return ctx->device; }
cudaStream_t ctx_stream(b200sk_ctx *ctx) { return ctx->own_stream; }
void **ctx_fx_slot(b200sk_ctx *ctx) { return &ctx->fx; }
void **ctx_reduce_slot(b200sk_ctx *ctx) { return &ctx->reduce; }
void reduce_free(void *p);
void ctx_set_error(b200sk_ctx *ctx, const char *msg) { ctx->last_error = msg; }
void ctx_add_launches(b200sk_ctx *ctx, uint64_t n) { ctx->launches += n; }
} // namespace b200sk

namespace {

int cuda_fail(b200sk_ctx *ctx, cudaError_t e, const char *what) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    if (ctx) ctx->last_error = buf;
    return B200SK_ERR_CUDA;
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call); \
    } while (0)

// Build the tile plan for a batch whose longest read is max_len (0 = unknown -> chunked geometry).
inline uint32_t pos_width_of(const b200sk_params &p) { return p.pos_width == 1 ? 1u : p.pos_width == 2 ? 2u : 4u; }

// whole_tile_stage: size the register-window kernels' tile buffer so that a WHOLE tile's output fits it (the sharded
// chain: one bulk store per array and tile instead of two rounds, at the price of two warps per SM).  Measured at
// N = 4 (profiles/r02_multi_gpu.md): 26.8 vs 28.0 ms with the uint8 positions, but 26.6 vs 23.4 ms without them --
// one 5.6 KB store per warp in flight does not keep the link as busy as two shorter ones.  Off by default.
int make_plan_c(const b200sk_params &p, uint64_t max_len, Plan &pl, bool whole_tile_stage, uint32_t chunk_reg);

// Long reads are cut into chunks of C positions, one lane each.  A chunk pays for its halo again (window warm-up, the
// k-1 folds) and for its share of the tile's fixed work, so C wants to be large; the lane's bases and staged lists
// want it small (warps per SM).  The register-window kernels take the C that maximises
//     warps(nw) x C / (C + 2 window + k)
// over C = 4 (mod 8) -- consecutive lanes start C bytes apart in shared memory, and only an odd number of words between
// them keeps the 32 lanes' word loads on 32 banks (C = 160: 8-way conflicts, syncmers 159 instead of 190 Gbases/s).
// Both terms are fits to measurements on ONT-like reads (profiles/r02ax_chunk.txt): the per-chunk overhead 2 window + k
// (43 steps for minimizers k=21 w=11, 61 for syncmers k=21 s=11), and the occupancy curve -- 1 at the cap (16 warps; 12
// for wide syncmer windows and protein minimizers), 0.96 at 12 of 16, 0.92 for the counts that load the four schedulers
// unevenly, 0.90 at 11 of 12.  Result: syncmers k=21 s=11 C = 156 (181 -> 190 Gbases/s against the fixed 132),
// minimizers k=21 w=11 C = 172 at 12 warps (310 -> 338).
int make_plan(const b200sk_params &p, uint64_t max_len, Plan &pl, bool whole_tile_stage = false) {
    static const uint32_t forced = [] { const char *e = getenv("B200SK_CHUNK_REG"); return e ? (uint32_t)atoi(e) : 0u; }();
    const int mode = p.mode;
    const uint64_t ext = p.circular ? (uint64_t)(p.k - 1) : 0;
    const bool chunked = !(max_len && max_len + ext <= kSingleMaxLen);
    const bool sparse = mode == B200SK_MODE_MINIMIZER || mode == B200SK_MODE_SYNCMER || mode == B200SK_MODE_PROTEIN_MINIMIZER;
    if (!chunked || !sparse || !sparse_reg_supported(mode, p.k, p.w, p.s))
        return make_plan_c(p, max_len, pl, whole_tile_stage, kChunkReg);
    if (forced) return make_plan_c(p, max_len, pl, whole_tile_stage, forced);
    const int cap = sparse_reg_max_warps(mode, p.k, p.w, p.s);
    const double window = mode == B200SK_MODE_SYNCMER ? 2.0 * (p.k - p.s) : (double)p.w;
    double best_score = -1.0;
    for (uint32_t c = 100; c <= 196; c += 8) {
        Plan cand;
        if (make_plan_c(p, max_len, cand, whole_tile_stage, c) || !cand.reg) continue;
        const int nw = cand.T / 32;
        double warps;
        if (nw >= cap) warps = 1.0;
        else if (cap == 16) warps = nw >= 12 ? (nw % 4 == 0 ? 0.96 : 0.92) : 0.96 * nw / 12.0;
        else warps = nw == cap - 1 ? 0.90 : 0.95 * nw / cap;
        const double score = warps * c / (c + 2.0 * window + p.k);
        if (score > best_score) { best_score = score; pl = cand; }
    }
    return best_score > 0 ? 0 : make_plan_c(p, max_len, pl, whole_tile_stage, kChunkReg);
}

int make_plan_c(const b200sk_params &p, uint64_t max_len, Plan &pl, bool whole_tile_stage, uint32_t chunk_reg) {
    const int mode = p.mode;
    const int k = p.k, w = p.w, s = p.s;
    const int d = mode == B200SK_MODE_SYNCMER ? k - s : 0;
    const uint32_t ww = (mode == B200SK_MODE_MINIMIZER || mode == B200SK_MODE_PROTEIN_MINIMIZER) ? (uint32_t)w
                        : mode == B200SK_MODE_SYNCMER ? 2u * d : 0u;
    uint64_t halo; // bases an item touches beyond its C positions
    double density;
    switch (mode) {
    case B200SK_MODE_MINIMIZER:
    case B200SK_MODE_PROTEIN_MINIMIZER: halo = (uint64_t)w + k - 1; density = 2.0 / (w + 1.0); break;
    case B200SK_MODE_SYNCMER: halo = 2ull * d + s - 1; density = 2.0 / (d + 1.0); break;
    case B200SK_MODE_PROTEIN: halo = (uint64_t)k - 1; density = 1.0; break;
    default: halo = (uint64_t)k - 1; density = 1.0; break;
    }
    const uint64_t ext = p.circular ? (uint64_t)(k - 1) : 0;
    if (max_len) max_len += ext;
    pl.chunked = !(max_len && max_len <= kSingleMaxLen);
    if (!pl.chunked) {
        int32_t st;
        ReadGeom g;
        g.mode = mode; g.k = k; g.w = w; g.s = s; g.frame = 1; g.canonical = p.canonical;
        g.protein_input = p.alphabet == B200SK_ALPHABET_PROTEIN; g.ill = nullptr;
        uint32_t np = read_positions(g, 0, max_len, max_len, &st);
        pl.C = np ? np : 1;
        pl.span_max = (uint32_t)max_len;
    } else {
        pl.C = (mode == B200SK_MODE_MINIMIZER || mode == B200SK_MODE_SYNCMER || mode == B200SK_MODE_PROTEIN_MINIMIZER) &&
                       sparse_reg_supported(mode, k, w, s)
                   ? chunk_reg : kChunk;
        if ((uint64_t)pl.C + halo > 20000) return B200SK_ERR_UNSUPPORTED;
        pl.span_max = (uint32_t)(pl.C + halo);
        if (mode == B200SK_MODE_PROTEIN && p.alphabet != B200SK_ALPHABET_PROTEIN) pl.span_max *= 3; // codons
    }
    if (pl.span_max < 16) pl.span_max = 16;
    const bool sparse = mode == B200SK_MODE_MINIMIZER || mode == B200SK_MODE_SYNCMER ||
                        mode == B200SK_MODE_PROTEIN_MINIMIZER;
    pl.lcap = 0;
    pl.dense = !sparse;
    if (pl.dense) {
        // tables 8 KB + control, per-warp output staging (two areas for both-strand k-mers), tile,
        // per-thread amino-acid buffers (protein)
        const uint32_t kDenseS = mode == B200SK_MODE_KMER ? 8 : 16; // must match DENSE_S in b200sk_dense.cu
        const uint32_t stage_w = (mode == B200SK_MODE_KMER ? 2u : 1u) * (32u * (kDenseS * 8u + 8u) + 512u);
        uint32_t aa_stride = 0;
        if (mode == B200SK_MODE_PROTEIN) {
            aa_stride = ((pl.C + (uint32_t)k - 1 + 3) / 4) | 1u; // odd number of words: conflict-free columns
            aa_stride *= 4;
        }
        pl.lcap = aa_stride;
        uint32_t ring_per_thread = 0; // SimHash: the window of n = k-m+1 m-mer hashes (w carries m)
        if (mode == B200SK_MODE_SIMHASH) ring_per_thread = (uint32_t)(k - w + 1) * 8u;
        for (int T : {128, 64, 32}) {
            pl.T = T;
            pl.sm_ring = 8192 + 768; // byte tables, control block, pair tables
            pl.sm_ring_bytes = (uint32_t)(T / 32) * stage_w;
            pl.sm_tile = pl.sm_ring + pl.sm_ring_bytes;
            pl.sm_tile_bytes = up16((uint32_t)T * pl.span_max + 64); // + alignment slop + word-granular look-ahead
            pl.sm_listv = pl.sm_tile + pl.sm_tile_bytes;
            pl.sm_listp = pl.sm_listv + up16((uint32_t)T * (aa_stride + ring_per_thread));
            pl.sm_total = pl.sm_listp;
            if (pl.sm_total <= 112 * 1024) return 0;
        }
        return pl.sm_total <= kSmemLimit ? 0 : B200SK_ERR_UNSUPPORTED;
    }
    const double slack = mode == B200SK_MODE_SYNCMER ? 1.15 : 1.45;
    if (sparse) {
        // staged-list capacity: the first window always emits, later ones at the density; +45% (about four
        // standard deviations on random reads) keeps the overflow path to ~1e-4 of the items
        // (bounded closed syncmers come out rarer and far more evenly spaced than 2/(d+1) suggests: 16.8 +- 1.4 per
        // 150-bp read, at most 25 in 300 k reads for k=21 s=11 -- a tighter factor buys one to three more warps per SM)
        double e = (1.0 + (pl.C - 1) * density) * slack + 2.0;
        pl.lcap = (uint32_t)std::min<double>(pl.C, e);
        if (pl.lcap < 1) pl.lcap = 1;
    }
    pl.reg = sparse && sparse_reg_supported(mode, k, w, s);
    pl.keyed = pl.reg && mode == B200SK_MODE_MINIMIZER && (w == 3 || w == 5 || w == 11 || w == 15) && keyed_walk_enabled();
    if (pl.reg) {
        // one tile per warp.  tables 4 KB, then per warp: mbarrier 16 B, tile, k-mer ring (syncmer),
        // lists of (lcap+1) slots x 32 lanes x (8 B value + 1 B position delta).
        int best_nw = 0;
        Plan best = pl;
        for (uint32_t shave = 0; shave <= 2 && shave + 8 < pl.lcap; shave++) {
            Plan c = pl;
            c.lcap = pl.lcap - shave;
            c.sm_tile = mode == B200SK_MODE_SYNCMER ? 3584u + 1024u : 4096u; // tables (+ the syncmer's fast tables)
            c.sm_tile_bytes = up16(32u * c.span_max + 32);
            // reads of n x 32 bytes are staged with one word of skew per lane (k_sparse_warp: bank conflicts)
            if (!c.chunked && c.span_max % 32u == 0 && mode != B200SK_MODE_PROTEIN_MINIMIZER) c.sm_tile_bytes += 128u;
            if (whole_tile_stage) { // expected elements of a tile + 10 %, 8-byte value + position each, + alignment slack
                const uint32_t pw = p.want_pos ? pos_width_of(p) : 0u;
                const uint32_t need = (uint32_t)(32.0 * c.lcap / slack * 1.10) * (8u + pw) + 64u;
                c.sm_tile_bytes = std::max(c.sm_tile_bytes, up16(need));
            }
            c.sm_ring = 16 + c.sm_tile_bytes;
            c.sm_listv = c.sm_ring + (uint32_t)d * 256u;
            c.sm_listp = c.sm_listv + (c.lcap + 1) * 256u;
            const uint32_t pos_bytes = (c.lcap + 1) * 32u; // one position byte per slot and lane
            c.sm_ring_bytes = up16(c.sm_listp + pos_bytes); // per-warp stride
            int nw = (int)((232448u - 1024u - c.sm_tile) / c.sm_ring_bytes);
            if (nw > sparse_reg_max_warps(mode, k, w, s)) nw = sparse_reg_max_warps(mode, k, w, s);
            {
                static const int cap_nw = [] { const char *e = getenv("B200SK_MAX_WARPS"); return e ? atoi(e) : 0; }();
                if (cap_nw > 0 && nw > cap_nw) nw = cap_nw; // testing knob: occupancy sweep
            }
            if (nw < 1) continue;
            c.T = nw * 32;
            c.ctas_per_sm = 1;
            c.sm_total = c.sm_tile + (uint32_t)nw * c.sm_ring_bytes;
            if (nw > best_nw) { best_nw = nw; best = c; }
        }
        if (best_nw > 0) { pl = best; return 0; }
        pl.reg = false;
    }
    for (int T : {128, 64, 32}) {
        pl.T = T;
        pl.sm_tile = kSmemCtl;
        pl.sm_tile_bytes = up16((uint32_t)T * pl.span_max + 32);
        pl.sm_ring = pl.sm_tile + pl.sm_tile_bytes;
        uint64_t ring = (uint64_t)ww * T * 8 + (uint64_t)d * T * 8 + (uint64_t)ww * T * 2;
        if (ring > kSmemLimit) continue;
        pl.sm_ring_bytes = up16((uint32_t)ring);
        pl.sm_listv = pl.sm_ring + pl.sm_ring_bytes;
        pl.sm_listp = pl.sm_listv + up16(pl.lcap * T * 8);
        pl.sm_total = pl.sm_listp + up16(pl.lcap * T * 2);
        if (pl.sm_total <= 112 * 1024) return 0;
    }
    if (pl.sm_total <= kSmemLimit) return 0; // T = 32, one CTA per SM
    return B200SK_ERR_UNSUPPORTED;
}

int ensure_meta(b200sk_ctx *ctx) {
    CK(ctx->meta.reserve(64));
    return 0;
}

// Everything that runs on the device for one batch (no host synchronisation unless the longest
// read must be measured).  d_flags: where the kernels OR their flags (device memory).

int enqueue(b200sk_ctx *ctx, const b200sk_params &p, const uint8_t *d_bases, const uint64_t *d_off,
            uint64_t n_reads, uint64_t n_bases, uint64_t *d_val, void *d_pos, uint64_t *d_ooff,
            int32_t *d_status, uint64_t capacity, uint64_t out_base, cudaStream_t st, uint32_t *d_flags,
            const b200sk_shard_spec *spec = nullptr) {
    int rc = b200sk_check_params(&p);
    if (rc) return rc;
    if (spec) {
        if (spec->n_ranks < 1 || spec->n_ranks > B200SK_MAX_RANKS || spec->rank < 0 || spec->rank >= spec->n_ranks ||
            spec->chunk_reads == 0 || spec->chunk_reads % 32u != 0 || (spec->epoch & 0x3fffu) == 0)
            return B200SK_ERR_BAD_ARG;
        for (int r = 0; r < spec->n_ranks; r++)
            if (!spec->state[r]) return B200SK_ERR_BAD_ARG;
        if (p.mode != B200SK_MODE_MINIMIZER && p.mode != B200SK_MODE_SYNCMER) return B200SK_ERR_UNSUPPORTED;
        if (p.circular || p.max_read_len == 0) return B200SK_ERR_UNSUPPORTED; // one item per read, known up front
    }
    if (!d_off || !d_ooff || (n_bases && !d_bases)) return B200SK_ERR_BAD_ARG;
    if (((uintptr_t)d_bases & 15u) != 0) return B200SK_ERR_BAD_ARG;
    if ((rc = ensure_meta(ctx))) return rc;
    if (ctx->last_done && st != ctx->last_stream) CK(cudaStreamWaitEvent(st, ctx->last_done, 0));
    unsigned long long *meta = (unsigned long long *)ctx->meta.p;
    CK(cudaMemsetAsync(meta, 0, 64, st));
    if (n_reads == 0) {
        uint64_t zero = out_base;
        CK(cudaMemcpyAsync(d_ooff, &zero, 8, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        return 0;
    }
    b200sk_params q = p;
    // "every k-mer" degenerations: syncmer with s == k (sketch.go:160,328-331), minimizer with w == 1
    // (sketch.go:103,218-222) -- both are the canonical hash stream with Index() = k-mer position
    if (q.mode == B200SK_MODE_SYNCMER && q.s == q.k) { q.mode = B200SK_MODE_NTHASH; q.canonical = 1; }
    if (q.mode == B200SK_MODE_MINIMIZER && q.w == 1) { q.mode = B200SK_MODE_NTHASH; q.canonical = 1; }
    if (q.mode == B200SK_MODE_MINIMIZER || q.mode == B200SK_MODE_SYNCMER) q.canonical = 1;
    if (q.mode == B200SK_MODE_PROTEIN || q.mode == B200SK_MODE_PROTEIN_MINIMIZER) q.circular = 0; // no such option

    KArgs a;
    memset(&a, 0, sizeof(a));
    a.bases = d_bases; a.off = d_off; a.off_orig = nullptr; a.n_reads = n_reads;
    a.mode = q.mode; a.k = q.k; a.w = q.w; a.s = q.s; a.canonical = q.canonical; a.frame = q.frame;
    a.alphabet = q.alphabet;
    if (q.mode == B200SK_MODE_SIMHASH) { // inside the kernels w carries m and s carries scale
        q.w = p.m; q.s = p.scale;
        a.w = p.m; a.s = p.scale;
    }

    if (q.mode == B200SK_MODE_KMER) {
        // NextKmer stops at the first illegal base (iterator.go:730-748): find it once per read
        CK(ctx->ill.reserve(n_reads * 4));
        CK(launch_first_illegal(d_bases, d_off, n_reads, (uint32_t *)ctx->ill.p, st, n_bases));
        ctx->launches++;
        a.ill = (const uint32_t *)ctx->ill.p;
    }
    const bool translate = (q.mode == B200SK_MODE_PROTEIN || q.mode == B200SK_MODE_PROTEIN_MINIMIZER) &&
                           q.alphabet != B200SK_ALPHABET_PROTEIN;
    if (translate) {
        if (ctx->aux_table != q.codon_table) {
            ctx->aux_host.resize(4608);
            if (!build_codon_aux(q.codon_table, ctx->aux_host.data())) return B200SK_ERR_CODON_TABLE;
            CK(ctx->aux.reserve(4608));
            CK(cudaMemcpyAsync(ctx->aux.p, ctx->aux_host.data(), 4608, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st)); // aux_host may be rebuilt by a later call
            ctx->aux_table = q.codon_table;
        }
        a.aux = (const uint8_t *)ctx->aux.p;
    } else if (q.mode == B200SK_MODE_PROTEIN) {
        CK(ctx->aux.reserve(4608));
        a.aux = (const uint8_t *)ctx->aux.p; // unused for amino-acid input, but the kernel copies the area
    }
    if (q.mode == B200SK_MODE_PROTEIN_MINIMIZER) {
        // sketch-protein.go:84: translate the frame once (k_translate), then every window of w consecutive
        // wyhash values over the amino acids.  Length checks keep looking at the record's own length.
        a.off_orig = d_off;
        if (translate) {
            ReadGeom g;
            memset(&g, 0, sizeof(g));
            g.mode = 100; g.frame = q.frame; g.k = q.k;
            const size_t sb2 = ((n_reads + 1023) / 1024 + 1) * 8;
            CK(ctx->aa_off.reserve((n_reads + 1) * 8));
            CK(ctx->aa_bases.reserve(n_bases / 3 + n_reads + 64));
            CK(ctx->scan_state.reserve(sb2));
            CK(cudaMemsetAsync(ctx->scan_state.p, 0, sb2, st));
            CK(cudaMemsetAsync(meta + 4, 0, 8, st));
            CK(launch_scan_geom(d_off, n_reads, g, (uint64_t *)ctx->aa_off.p, (uint64_t *)ctx->scan_state.p, meta + 4, st));
            CK(launch_translate(d_bases, d_off, (const uint64_t *)ctx->aa_off.p, n_reads, q.frame,
                                (const uint8_t *)ctx->aux.p, (uint8_t *)ctx->aa_bases.p, st));
            ctx->launches += 2;
            a.bases = (const uint8_t *)ctx->aa_bases.p;
            a.off = (const uint64_t *)ctx->aa_off.p;
            n_bases = n_bases / 3 + n_reads;
        }
        if (q.w == 1) { // every amino-acid k-mer (sketch-protein.go:95,120-124): the dense protein kernel
            q.mode = B200SK_MODE_PROTEIN;
            a.mode = B200SK_MODE_PROTEIN;
            q.alphabet = B200SK_ALPHABET_PROTEIN;
            a.alphabet = B200SK_ALPHABET_PROTEIN;
        }
    }

    if (q.circular) {
        // seq2 = S + S[0:k-1] (iterator.go:642-646, sketch.go:106-110): materialise the extended
        // batch once on the device; the kernels then see ordinary linear reads.
        CK(ctx->circ_bases.reserve(n_bases + n_reads * (uint64_t)(q.k - 1) + 64));
        CK(ctx->circ_off.reserve((n_reads + 1) * 8));
        CK(ctx->scan_state.reserve(((n_reads + 1023) / 1024 + 1) * 8));
        CK(cudaMemsetAsync(ctx->scan_state.p, 0, ((n_reads + 1023) / 1024 + 1) * 8, st));
        CK(launch_circularize(d_bases, d_off, n_reads, q.k, (uint8_t *)ctx->circ_bases.p,
                              (uint64_t *)ctx->circ_off.p, (uint64_t *)ctx->scan_state.p, meta + 4, st));
        ctx->launches += 2;
        a.off_orig = d_off;
        a.bases = (const uint8_t *)ctx->circ_bases.p;
        a.off = (const uint64_t *)ctx->circ_off.p;
        n_bases += n_reads * (uint64_t)(q.k - 1);
    }

    uint64_t max_len = p.max_read_len;
    if (p.mode == B200SK_MODE_PROTEIN_MINIMIZER && translate && max_len) max_len = max_len / 3 + 1;
    Plan pl;
    uint64_t n_items_host = n_reads;
    bool have_items_host = true;
    if (max_len == 0) {
        // measure the longest read (and, with the chunked geometry, the item count)
        { Plan tmp; if ((rc = make_plan(q, 0, tmp))) return rc; a.C = tmp.C; }
        CK(launch_prepass(a, meta + 2, st));
        ctx->launches++;
        unsigned long long hm[2];
        CK(cudaMemcpyAsync(hm, meta + 2, 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        max_len = hm[1] ? hm[1] : 1;
        if (max_len > 0xfffffff0ull) return B200SK_ERR_UNSUPPORTED; // positions of one record are 32-bit (read_positions)
        if (q.circular) max_len = max_len > (uint64_t)(q.k - 1) ? max_len - (q.k - 1) : 1; // plan re-adds it
        if ((rc = make_plan(q, max_len, pl, spec && spec->n_ranks >= kWholeTileRanks))) return rc;
        if (pl.chunked) n_items_host = hm[0];
    } else {
        if ((rc = make_plan(q, max_len, pl, spec && spec->n_ranks >= kWholeTileRanks))) return rc;
        if (pl.chunked) {
            a.C = pl.C;
            CK(launch_prepass(a, meta + 2, st));
            ctx->launches++;
            have_items_host = false;
        }
    }
    if (spec) {
        // the global tile chain needs one item per read and the warp-tile kernel (32 reads per tile)
        if (pl.chunked || !pl.reg) return B200SK_ERR_UNSUPPORTED;
        a.shard.n = (uint32_t)spec->n_ranks; a.shard.rank = (uint32_t)spec->rank; a.shard.epoch = spec->epoch;
        for (int r = 0; r < spec->n_ranks; r++) a.shard.copy[r] = spec->state[r];
        static const uint32_t poll_ns = [] { const char *e = getenv("B200SK_CHAIN_POLL_NS"); return e ? (uint32_t)atoi(e) : 1000u; }();
        a.shard.poll_ns = poll_ns; // tuning knob; the default is the measured optimum
        static const uint32_t poll_free = [] { const char *e = getenv("B200SK_CHAIN_POLL_FREE"); return e ? (uint32_t)atoi(e) : 0u; }();
        a.shard.poll_free = poll_free;
        a.shard_chunk_tiles = spec->chunk_reads / 32u;
        a.shard_n_reads = spec->n_reads_global;
    }
    a.C = pl.C; a.span_max = pl.span_max; a.lcap = pl.lcap;
    a.keyed = pl.keyed ? 1u : 0u;
    a.skew = (pl.reg && !pl.chunked && !spec && pl.span_max % 32u == 0 && q.mode != B200SK_MODE_PROTEIN_MINIMIZER) ? 1u : 0u; // lanes a read apart would share a bank
    if (pl.keyed) {
        static const bool force_rewalk = [] { const char *e = getenv("B200SK_WALKER"); return e && strcmp(e, "rewalk") == 0; }();
        if (force_rewalk) a.keyed |= 2u; // testing knob: every item also takes the exact re-walk
        static const bool no_rewalk = [] { const char *e = getenv("B200SK_WALKER"); return e && strcmp(e, "norewalk") == 0; }();
        if (no_rewalk) a.keyed |= 4u; // timing experiment only: WRONG output for the items the keyed walk hands back
    }
#ifdef B200SK_EXPERIMENTS
    {
        static const uint32_t spin = [] { const char *e = getenv("B200SK_SPIN_NS"); return e ? (uint32_t)atoi(e) : 0u; }();
        a.spin_ns = spin;
        static const bool unordered = [] { const char *e = getenv("B200SK_UNORDERED"); return e && atoi(e) != 0; }();
        a.unordered = unordered ? meta + 6 : nullptr;
    }
#endif
    a.rewalks = meta + 5;
    a.sm_tile = pl.sm_tile; a.sm_tile_bytes = pl.sm_tile_bytes; a.sm_ring = pl.sm_ring;
    a.sm_ring_bytes = pl.sm_ring_bytes; a.sm_listv = pl.sm_listv; a.sm_listp = pl.sm_listp;
    a.sm_total = pl.sm_total;
    a.out_val = d_val; a.out_pos = p.want_pos ? d_pos : nullptr; a.out_off = d_ooff; a.status = d_status;
    a.pos_width = pos_width_of(p);
    a.capacity = d_val ? capacity : 0; a.out_base = out_base;
    a.flags = d_flags ? d_flags : (uint32_t *)(meta + 1);

    const size_t sbytes = ((n_reads + 1023) / 1024 + 1) * 8;
    if (pl.dense) {
        // dense modes: output offsets and statuses follow from the lengths
        CK(ctx->scan_state.reserve(sbytes));
        CK(cudaMemsetAsync(ctx->scan_state.p, 0, sbytes, st));
        CK(cudaMemsetAsync(meta + 4, 0, 8, st));
        CK(launch_scan_counts(a, (uint64_t *)ctx->scan_state.p, meta + 4, st));
        ctx->launches++;
    }
    uint64_t items_bound = n_items_host;
    if (pl.chunked) {
        if (!have_items_host) {
            uint64_t unit = pl.C;
            if (q.mode == B200SK_MODE_PROTEIN && q.alphabet != B200SK_ALPHABET_PROTEIN) unit *= 3;
            items_bound = n_reads + n_bases / unit + 1;
        }
        CK(ctx->item_first.reserve((n_reads + 1) * 8));
        CK(ctx->scan_state.reserve(sbytes));
        CK(cudaMemsetAsync(ctx->scan_state.p, 0, sbytes, st));
        CK(cudaMemsetAsync(meta + 4, 0, 8, st));
        // the read that holds the first item of every group of 32 items: item -> (read, chunk) is then a search
        // over 32 reads instead of all of them
        CK(ctx->tile_read.reserve((items_bound / 32 + 2) * 8));
        CK(launch_scan_items(a, (uint64_t *)ctx->item_first.p, (uint64_t *)ctx->scan_state.p, meta + 4, st,
                             (uint64_t *)ctx->tile_read.p));
        a.tile_read = (const uint64_t *)ctx->tile_read.p;
        ctx->launches++;
        a.item_first = (const uint64_t *)ctx->item_first.p;
        a.n_items_dev = (const uint64_t *)ctx->item_first.p + n_reads;
        a.n_items = 0;
    } else {
        a.item_first = nullptr;
        a.n_items_dev = nullptr;
        a.n_items = n_reads;
    }
    const uint32_t tile_items = pl.reg ? 32u : (uint32_t)pl.T; // the register-window kernels tile per warp
    const uint64_t n_tiles = (items_bound + tile_items - 1) / tile_items + 1;
    if (!pl.dense && !spec) { // (a sharded batch chains through the ranks' epoch-tagged words instead: never reset)
        CK(ctx->tile_state.reserve(n_tiles * 8));
        CK(cudaMemsetAsync(ctx->tile_state.p, 0, n_tiles * 8, st));
    }
    a.tile_state = (uint64_t *)ctx->tile_state.p;
    a.ticket = meta + 0;
    int occ = 1;
    if (pl.dense) CK(launch_dense(a, pl.T, 0, st, &occ));
    else if (pl.reg) CK(launch_sparse_reg(a, pl.T, 0, st, &occ));
    else occ = main_kernel_occupancy(a, pl.T);
    uint64_t blocks = (uint64_t)occ * ctx->sm_count;
    const uint64_t tiles_per_block = pl.reg ? (uint64_t)pl.T / 32 : 1;
    if (blocks > (n_tiles + tiles_per_block - 1) / tiles_per_block) blocks = (n_tiles + tiles_per_block - 1) / tiles_per_block;
    if (blocks < 1) blocks = 1;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->timing) {
        CK(cudaEventCreate(&ev0));
        CK(cudaEventCreate(&ev1));
        CK(cudaEventRecord(ev0, st));
    }
    if (pl.dense) CK(launch_dense(a, pl.T, (int)blocks, st, nullptr));
    else if (pl.reg) CK(launch_sparse_reg(a, pl.T, (int)blocks, st, nullptr));
    else CK(launch_main(a, pl.T, (int)blocks, st));
    ctx->launches++;
    if (ctx->timing) {
        CK(cudaEventRecord(ev1, st));
        ctx->timing_events.emplace_back(ev0, ev1);
    }
    if (!ctx->last_done) CK(cudaEventCreateWithFlags(&ctx->last_done, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->last_done, st));
    ctx->last_stream = st;
    return 0;
}

} // namespace

// ------------------------------------------------------------------ C ABI
extern "C" {

int b200sk_version(void) { return 100; }

int b200sk_check_params(const b200sk_params *p) {
    if (!p) return B200SK_ERR_BAD_ARG;
    if (p->pos_width != 0 && p->pos_width != 1 && p->pos_width != 2 && p->pos_width != 4) return B200SK_ERR_BAD_ARG;
    if (p->want_pos && (p->pos_width == 1 || p->pos_width == 2)) {
        // every Index() must fit: needs the hint, counted on the extended length when circular
        const uint64_t lim = p->pos_width == 1 ? 256 : 65536;
        const uint64_t ext = p->circular && p->k > 0 ? (uint64_t)p->k - 1 : 0;
        if (p->max_read_len == 0 || (uint64_t)p->max_read_len + ext > lim) return B200SK_ERR_BAD_ARG;
    }
    switch (p->mode) {
    case B200SK_MODE_KMER:
        if (p->k < 1) return B200SK_ERR_INVALID_K;       // iterator.go:669
        if (p->k > 32) return B200SK_ERR_K_OVERFLOW;     // kmers.Encode, iterator.go:742
        return 0;
    case B200SK_MODE_NTHASH:
        if (p->k < 1) return B200SK_ERR_INVALID_K;       // iterator.go:616
        return 0;
    case B200SK_MODE_MINIMIZER:
        if (p->k < 1) return B200SK_ERR_INVALID_K;       // sketch.go:86
        if (p->w < 1) return B200SK_ERR_INVALID_W;       // sketch.go:89 (w > 2^31-1 cannot be expressed in int32)
        return 0;
    case B200SK_MODE_SYNCMER:
        if (p->k < 1) return B200SK_ERR_INVALID_K;       // sketch.go:143
        if (p->s > p->k || p->s <= 0) return B200SK_ERR_INVALID_S; // sketch.go:146 (s<0: uint(s) overflows NewHasher)
        return 0;
    case B200SK_MODE_PROTEIN: {
        if (p->k < 1) return B200SK_ERR_INVALID_K;       // iterator-protein.go:47
        if (p->alphabet == B200SK_ALPHABET_PROTEIN) return 0; // already amino acids (iterator-protein.go:68)
        if (p->alphabet == B200SK_ALPHABET_UNLIMIT) return B200SK_ERR_BAD_ARG; // seq.go:686: only DNA/RNA translate
        uint8_t tmp[4608];
        if (!build_codon_aux(p->codon_table, tmp)) return B200SK_ERR_CODON_TABLE;      // seq.go:691
        if (p->frame < -3 || p->frame > 3 || p->frame == 0) return B200SK_ERR_INVALID_FRAME; // seq.go:694
        return 0;
    }
    case B200SK_MODE_SIMHASH:
        if (p->k < 1) return B200SK_ERR_INVALID_K;                       // iterator.go:114
        if (p->k >= 65535) return B200SK_ERR_K_TOO_LARGE;                // iterator.go:117
        if (p->m < 4 || p->m > p->k) return B200SK_ERR_INVALID_M;        // iterator.go:121
        if (p->scale < 1 || p->scale > p->k - p->m + 1) return B200SK_ERR_INVALID_SCALE; // iterator.go:124
        if (p->k - p->m + 1 > 255) return B200SK_ERR_UNSUPPORTED;        // 8 counter planes
        return 0;
    case B200SK_MODE_PROTEIN_MINIMIZER: {
        if (p->k < 1) return B200SK_ERR_INVALID_K;       // sketch-protein.go:63
        if (p->w < 1) return B200SK_ERR_INVALID_W;       // sketch-protein.go:70
        if (p->alphabet == B200SK_ALPHABET_PROTEIN) return 0;
        if (p->alphabet == B200SK_ALPHABET_UNLIMIT) return B200SK_ERR_BAD_ARG;
        uint8_t tmp[4608];
        if (!build_codon_aux(p->codon_table, tmp)) return B200SK_ERR_CODON_TABLE;
        if (p->frame < -3 || p->frame > 3 || p->frame == 0) return B200SK_ERR_INVALID_FRAME;
        return 0;
    }
    default:
        return B200SK_ERR_BAD_ARG;
    }
}

const char *b200sk_strerror(int code) {
    switch (code) {
    case B200SK_OK: return "ok";
    case B200SK_ERR_INVALID_K: return "sketches: invalid k-mer size";
    case B200SK_ERR_SHORT_SEQ: return "sketches: sequence too short";
    case B200SK_ERR_INVALID_W: return "kmers: invalid minimimzer window";
    case B200SK_ERR_INVALID_S: return "kmers: invalid s-mer size";
    case B200SK_ERR_ILLEGAL_BASE: return "sketches: illegal base";
    case B200SK_ERR_K_OVERFLOW: return "unikmer: k-mer size (1-32) overflow";
    case B200SK_ERR_INVALID_FRAME: return "seq: invalid frame. available: 1, 2, 3, -1, -2, -3";
    case B200SK_ERR_CODON_TABLE: return "seq: invalid codon table";
    case B200SK_ERR_TRANSLATE_SHORT: return "seq: sequence too short to translate";
    case B200SK_ERR_INVALID_CODON: return "seq: invalid DNA base";
    case B200SK_ERR_INVALID_M: return "sketches: invalid m-mer size, should be in range of [4, k]";
    case B200SK_ERR_INVALID_SCALE: return "sketches: invalid scale, should be in range of [1, k-m+1]";
    case B200SK_ERR_K_TOO_LARGE: return "sketches: k-mer size is too large";
    case B200SK_ERR_NOT_FASTX: return "fastx: invalid FASTA/Q format";
    case B200SK_ERR_BAD_FASTQ: return "fastx: bad fastq format";
    case B200SK_ERR_CUDA: return "b200sketch: CUDA error";
    case B200SK_ERR_NO_DEVICE: return "b200sketch: no CUDA device (there is no CPU fallback)";
    case B200SK_ERR_UNSUPPORTED: return "b200sketch: parameters outside the implemented range";
    case B200SK_ERR_CAPACITY: return "b200sketch: output capacity too small";
    case B200SK_ERR_BAD_ARG: return "b200sketch: bad argument";
    case B200SK_ERR_NOMEM: return "b200sketch: out of memory";
    default: return "b200sketch: unknown error";
    }
}

int b200sk_create(b200sk_ctx **out, int device) {
    if (!out) return B200SK_ERR_BAD_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return B200SK_ERR_NO_DEVICE;
    }
    b200sk_ctx *ctx = new b200sk_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return B200SK_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return B200SK_ERR_CUDA;
    }
    *out = ctx;
    return 0;
}

void b200sk_destroy(b200sk_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (DevBuf *b : {&ctx->meta, &ctx->tile_state, &ctx->scan_state, &ctx->item_first, &ctx->tile_read, &ctx->circ_bases,
                      &ctx->circ_off, &ctx->ill, &ctx->aux, &ctx->aa_bases, &ctx->aa_off, &ctx->d_bases, &ctx->d_off, &ctx->d_val, &ctx->d_pos, &ctx->d_ooff,
                      &ctx->d_status, &ctx->d_bases2, &ctx->d_off2, &ctx->d_val2, &ctx->d_pos2, &ctx->d_ooff2,
                      &ctx->d_status2, &ctx->d_acc, &ctx->d_acc2, &ctx->d_acc_cnt})
        b->release();
    for (HostBuf *b : {&ctx->h_val, &ctx->h_pos, &ctx->h_ooff, &ctx->h_status, &ctx->h_meta}) b->release();
    b200sk::fx_free(ctx->fx);
    b200sk::reduce_free(ctx->reduce);
    if (ctx->last_done) cudaEventDestroy(ctx->last_done);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    delete ctx;
}

void *b200sk_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void b200sk_free_pinned(void *p) {
    if (p) cudaFreeHost(p);
}

void b200sk_timing_enable(b200sk_ctx *ctx, int on) {
    if (ctx) ctx->timing = on != 0;
}
int b200sk_timing_collect(b200sk_ctx *ctx, double *sum_ms, uint64_t *n) {
    if (!ctx) return B200SK_ERR_BAD_ARG;
    double sum = 0;
    uint64_t cnt = 0;
    for (auto &pr : ctx->timing_events) {
        float ms = 0;
        CK(cudaEventSynchronize(pr.second));
        CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
        sum += ms;
        cnt++;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    ctx->timing_events.clear();
    if (sum_ms) *sum_ms = sum;
    if (n) *n = cnt;
    return 0;
}

const char *b200sk_last_error(const b200sk_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t b200sk_kernel_launches(const b200sk_ctx *ctx) { return ctx ? ctx->launches : 0; }

uint64_t b200sk_output_bound(const b200sk_params *p, uint64_t n_bases, uint64_t n_reads, int exact) {
    if (!p) return 0;
    const uint64_t ext = p->circular ? n_reads * (uint64_t)(p->k > 0 ? p->k - 1 : 0) : 0;
    const uint64_t nb = n_bases + ext;
    switch (p->mode) {
    case B200SK_MODE_KMER: return (p->canonical ? 1 : 2) * nb;
    case B200SK_MODE_NTHASH: return nb;
    case B200SK_MODE_SIMHASH: return nb;
    case B200SK_MODE_PROTEIN: return nb / 3 + n_reads;
    case B200SK_MODE_PROTEIN_MINIMIZER:
        if (exact || p->w <= 1) return nb / 3 + n_reads;
        return (uint64_t)((nb / 3 + n_reads) * (2.0 / (p->w + 1.0)) * 1.25) + n_reads + 1024;
    case B200SK_MODE_MINIMIZER:
        if (exact || p->w <= 1) return nb;
        return (uint64_t)(nb * (2.0 / (p->w + 1.0)) * 1.25) + n_reads + 1024;
    case B200SK_MODE_SYNCMER: {
        const int d = p->k - p->s;
        if (exact || d <= 0) return nb;
        return (uint64_t)(nb * (2.0 / (d + 1.0)) * 1.25) + n_reads + 1024;
    }
    default: return 0;
    }
}

int b200sk_enqueue_device(b200sk_ctx *ctx, const b200sk_params *p, const uint8_t *d_bases,
                          const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases, uint64_t *d_out_val,
                          uint32_t *d_out_pos, uint64_t *d_out_off, int32_t *d_read_status, uint64_t capacity,
                          void *stream, uint32_t *d_flags) {
    if (!ctx || !p) return B200SK_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    if (d_flags) CK(cudaMemsetAsync(d_flags, 0, 4, (cudaStream_t)stream));
    return enqueue(ctx, *p, d_bases, d_read_off, n_reads, n_bases, d_out_val, d_out_pos, d_out_off,
                   d_read_status, capacity, 0, (cudaStream_t)stream, d_flags);
}

int b200sk_enqueue_device_sharded(b200sk_ctx *ctx, const b200sk_params *p, const b200sk_shard_spec *spec,
                                  const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases,
                                  uint64_t *d_out_val, uint32_t *d_out_pos, uint64_t *d_out_off, int32_t *d_read_status,
                                  uint64_t capacity, void *stream, uint32_t *d_flags) {
    if (!ctx || !p || !spec) return B200SK_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    if (d_flags) CK(cudaMemsetAsync(d_flags, 0, 4, (cudaStream_t)stream));
    if (n_reads == 0) return 0; // a rank without reads publishes nothing: no tile of the chain is its own
    return enqueue(ctx, *p, d_bases, d_read_off, n_reads, n_bases, d_out_val, d_out_pos, d_out_off, d_read_status,
                   capacity, 0, (cudaStream_t)stream, d_flags, spec);
}

// All six frames of ProteinIterator over one batch.  Reads of one item each (max_read_len hint), k <= 16, nucleotide
// input: six offset scans (one per frame: counts follow from the lengths) and ONE sketching launch that fetches and
// rewrites every tile of reads once and walks it six times (k_protein6_warp).  Anything else: six ordinary batches.
int b200sk_enqueue_device_frames(b200sk_ctx *ctx, const b200sk_params *p, const uint8_t *d_bases,
                                 const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases,
                                 uint64_t *const *d_out_val, uint64_t *const *d_out_off, int32_t *const *d_read_status,
                                 uint64_t capacity, void *stream, uint32_t *d_flags) {
    if (!ctx || !p || !d_out_val || !d_out_off || !d_read_off) return B200SK_ERR_BAD_ARG;
    if (p->mode != B200SK_MODE_PROTEIN) return B200SK_ERR_BAD_ARG;
    static const int kFrames[6] = {1, 2, 3, -1, -2, -3};
    for (int fi = 0; fi < 6; fi++)
        if (!d_out_off[fi] || (capacity && !d_out_val[fi])) return B200SK_ERR_BAD_ARG;
    b200sk_params q = *p;
    q.want_pos = 0; // Index() of a dense mode is the running position
    q.circular = 0;
    q.frame = 1;
    int rc = b200sk_check_params(&q);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (d_flags) CK(cudaMemsetAsync(d_flags, 0, 4, st));
    const bool fused = q.alphabet != B200SK_ALPHABET_PROTEIN && q.k <= 16 && q.max_read_len != 0 &&
                       q.max_read_len <= kSingleMaxLen && nthash_warp_fits(q.max_read_len) && n_reads != 0 &&
                       ((uintptr_t)d_bases & 15u) == 0;
    if (!fused) {
        for (int fi = 0; fi < 6; fi++) {
            q.frame = kFrames[fi];
            rc = enqueue(ctx, q, d_bases, d_read_off, n_reads, n_bases, d_out_val[fi], nullptr, d_out_off[fi],
                         d_read_status ? d_read_status[fi] : nullptr, capacity, 0, st, d_flags);
            if (rc) return rc;
        }
        return 0;
    }
    if ((rc = ensure_meta(ctx))) return rc;
    if (ctx->last_done && st != ctx->last_stream) CK(cudaStreamWaitEvent(st, ctx->last_done, 0));
    unsigned long long *meta = (unsigned long long *)ctx->meta.p;
    CK(cudaMemsetAsync(meta, 0, 64, st));
    if (ctx->aux_table != q.codon_table) {
        ctx->aux_host.resize(4608);
        if (!build_codon_aux(q.codon_table, ctx->aux_host.data())) return B200SK_ERR_CODON_TABLE;
        CK(ctx->aux.reserve(4608));
        CK(cudaMemcpyAsync(ctx->aux.p, ctx->aux_host.data(), 4608, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st)); // aux_host may be rebuilt by a later call
        ctx->aux_table = q.codon_table;
    }
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.bases = d_bases; a.off = d_read_off; a.n_reads = n_reads; a.n_items = n_reads;
    a.mode = q.mode; a.k = q.k; a.canonical = q.canonical; a.alphabet = q.alphabet;
    a.aux = (const uint8_t *)ctx->aux.p;
    a.capacity = capacity; a.out_base = 0; a.pos_width = 4;
    a.flags = d_flags ? d_flags : (uint32_t *)(meta + 1);
    a.C = 1; a.span_max = q.max_read_len;
    const size_t sbytes = ((n_reads + 1023) / 1024 + 1) * 8;
    CK(ctx->scan_state.reserve(sbytes));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->timing) {
        CK(cudaEventCreate(&ev0));
        CK(cudaEventCreate(&ev1));
        CK(cudaEventRecord(ev0, st));
    }
    for (int fi = 0; fi < 6; fi++) { // offsets and statuses of every frame follow from the read lengths
        a.frame = kFrames[fi];
        a.out_off = d_out_off[fi];
        a.status = d_read_status ? d_read_status[fi] : nullptr;
        CK(cudaMemsetAsync(ctx->scan_state.p, 0, sbytes, st));
        CK(cudaMemsetAsync(meta + 4, 0, 8, st));
        CK(launch_scan_counts(a, (uint64_t *)ctx->scan_state.p, meta + 4, st));
        a.fr_val[fi] = d_out_val[fi];
        a.fr_off[fi] = d_out_off[fi];
    }
    a.frame = 1; a.out_off = nullptr; a.status = nullptr;
    a.ticket = meta + 0;
    CK(launch_protein6_warp(a, st));
    ctx->launches += 7;
    if (ctx->timing) {
        CK(cudaEventRecord(ev1, st));
        ctx->timing_events.emplace_back(ev0, ev1);
    }
    if (!ctx->last_done) CK(cudaEventCreateWithFlags(&ctx->last_done, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->last_done, st));
    ctx->last_stream = st;
    return 0;
}

// Host entry point of the six-frame call: the batch goes to the device ONCE, the six frames' sketches come back in
// library-owned pinned arrays (valid until the next b200sk_run* on this context).  One copy in, one launch group, six
// copies out on the context's stream; the values dominate the PCIe traffic (8 B per amino-acid k-mer and frame).
int b200sk_run_frames(b200sk_ctx *ctx, const b200sk_params *p, const uint8_t *bases, const uint64_t *read_off,
                      uint64_t n_reads, uint64_t **out_val, uint64_t **out_off, int32_t **read_status, uint64_t *n_out) {
    if (!ctx || !p || !read_off || !out_val || !out_off || !n_out) return B200SK_ERR_BAD_ARG;
    if (p->mode != B200SK_MODE_PROTEIN) return B200SK_ERR_BAD_ARG;
    int rc = b200sk_check_params(p);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    const uint64_t b0 = read_off[0], n_bases = read_off[n_reads] - b0;
    if (n_bases && !bases) return B200SK_ERR_BAD_ARG;
    const uint64_t cap = b200sk_output_bound(p, n_bases, n_reads, 1) + 8; // per frame
    CK(ctx->d_bases.reserve(n_bases + 64));
    CK(ctx->d_off.reserve((n_reads + 1) * 8));
    CK(ctx->d_val.reserve(6 * cap * 8));
    CK(ctx->d_ooff.reserve(6 * (n_reads + 1) * 8));
    CK(ctx->d_status.reserve(6 * (n_reads + 1) * 4));
    CK(ctx->h_ooff.reserve(6 * (n_reads + 1) * 8));
    CK(ctx->h_status.reserve(6 * (n_reads + 1) * 4));
    CK(ctx->h_meta.reserve(64));
    const uint64_t *offs = read_off;
    std::vector<uint64_t> rebased;
    if (b0) { // offsets relative to the first base of the batch
        rebased.resize(n_reads + 1);
        for (uint64_t i = 0; i <= n_reads; i++) rebased[i] = read_off[i] - b0;
        offs = rebased.data();
    }
    CK(cudaMemsetAsync((uint8_t *)ctx->d_bases.p + n_bases, 0, 64, st)); // the TMA tile may read up to 16 bytes past the end
    if (n_bases) CK(cudaMemcpyAsync(ctx->d_bases.p, bases + b0, n_bases, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_off.p, offs, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    uint64_t *dv[6], *doff[6];
    int32_t *dst[6];
    for (int fi = 0; fi < 6; fi++) {
        dv[fi] = (uint64_t *)ctx->d_val.p + fi * cap;
        doff[fi] = (uint64_t *)ctx->d_ooff.p + fi * (n_reads + 1);
        dst[fi] = (int32_t *)ctx->d_status.p + fi * (n_reads + 1);
    }
    uint32_t *d_flags = (uint32_t *)ctx->d_status.p + 6 * (n_reads + 1) - 1; // the last (unused) status slot
    rc = b200sk_enqueue_device_frames(ctx, p, (const uint8_t *)ctx->d_bases.p, (const uint64_t *)ctx->d_off.p, n_reads,
                                      n_bases, dv, doff, dst, cap, st, d_flags);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_ooff.p, ctx->d_ooff.p, 6 * (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_status.p, ctx->d_status.p, 6 * (n_reads + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st)); // the offset tables hold the totals
    const uint32_t flags = ((const uint32_t *)ctx->h_status.p)[6 * (n_reads + 1) - 1];
    if (flags & B200SK_FLAG_SPAN) return B200SK_ERR_BAD_ARG; // max_read_len hint smaller than a read
    if (flags & B200SK_FLAG_CAPACITY) return B200SK_ERR_CAPACITY;
    uint64_t total = 0, start[6];
    for (int fi = 0; fi < 6; fi++) {
        n_out[fi] = ((const uint64_t *)ctx->h_ooff.p)[fi * (n_reads + 1) + n_reads];
        start[fi] = total;
        total += n_out[fi];
    }
    CK(ctx->h_val.reserve((total + 1) * 8));
    for (int fi = 0; fi < 6; fi++) {
        if (n_out[fi]) CK(cudaMemcpyAsync((uint64_t *)ctx->h_val.p + start[fi], dv[fi], n_out[fi] * 8, cudaMemcpyDeviceToHost, st));
        out_val[fi] = (uint64_t *)ctx->h_val.p + start[fi];
        out_off[fi] = (uint64_t *)ctx->h_ooff.p + fi * (n_reads + 1);
        if (read_status) read_status[fi] = (int32_t *)ctx->h_status.p + fi * (n_reads + 1);
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

int b200sk_run_device(b200sk_ctx *ctx, const b200sk_params *p, const uint8_t *d_bases,
                      const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases, uint64_t *d_out_val,
                      uint32_t *d_out_pos, uint64_t *d_out_off, int32_t *d_read_status, uint64_t capacity,
                      void *stream, uint64_t *n_out) {
    if (!ctx || !p) return B200SK_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = enqueue(ctx, *p, d_bases, d_read_off, n_reads, n_bases, d_out_val, d_out_pos, d_out_off,
                     d_read_status, capacity, 0, st, nullptr);
    if (rc) return rc;
    uint64_t total = 0;
    unsigned long long flags = 0;
    CK(cudaMemcpyAsync(&total, d_out_off + n_reads, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&flags, (unsigned long long *)ctx->meta.p + 1, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_out) *n_out = total;
    if (flags & B200SK_FLAG_SPAN) return B200SK_ERR_BAD_ARG; // max_read_len hint smaller than a read
    if (flags & B200SK_FLAG_CAPACITY) return B200SK_ERR_CAPACITY;
    return 0;
}

// Host entry point.  Sub-batches of reads flow through three streams (H2D copy, kernels, D2H copy) and
// two device slots, so the PCIe transfers of neighbouring sub-batches overlap each other (full duplex)
// and the kernels.  Output offsets are made global on the device (out_base = elements emitted so far).
// reduce != nullptr: the sketches of every sub-batch stay on the device, their FracMinHash fraction is appended to
// one accumulation array there, and only sort | unique of that array comes back (b200sk_run_reduced)
struct ReduceOpt { uint32_t scale; int unique; };
static int run_host(b200sk_ctx *ctx, const b200sk_params *p_in, const uint8_t *bases, const uint64_t *read_off,
                    uint64_t n_reads, uint64_t **out_val, uint32_t **out_pos, uint64_t **out_off,
                    int32_t **read_status, uint64_t *n_out, const ReduceOpt *reduce) {
    if (!ctx || !p_in || !read_off) return B200SK_ERR_BAD_ARG;
    b200sk_params pcopy = *p_in;
    if (reduce) pcopy.want_pos = 0; // an Index() means nothing once the values are sorted
    const b200sk_params *p = &pcopy;
    int rc = b200sk_check_params(p);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s_in = ctx->copy_in, s_k = ctx->own_stream, s_out = ctx->copy_out;
    const bool want_pos = p->want_pos != 0;
    const uint64_t n_bases = read_off[n_reads] - read_off[0];
    CK(ctx->h_ooff.reserve((n_reads + 1) * 8));
    CK(ctx->h_status.reserve((n_reads + 1) * 4));
    CK(ctx->h_meta.reserve(64));
    uint64_t host_cap = b200sk_output_bound(p, n_bases, n_reads, 0);
    uint64_t acc_cap = 0;
    const uint64_t max_hash = reduce ? b200sk_scale_max_hash(reduce->scale) : ~0ULL;
    if (reduce) { // the accumulation array: the expected fraction with a wide margin (checked, not trusted)
        // (a window minimum of `win` hashes is <= x with probability ~ win * x: minimizers pass the filter win times
        // as often as plain k-mer hashes do)
        const uint64_t win = p->mode == B200SK_MODE_MINIMIZER || p->mode == B200SK_MODE_PROTEIN_MINIMIZER ? (uint64_t)p->w + 1
                             : p->mode == B200SK_MODE_SYNCMER ? 2ull * (uint64_t)(p->k - p->s) + 1 : 1ull;
        acc_cap = reduce->scale > 1 ? std::min<uint64_t>(host_cap, host_cap / reduce->scale * 2 * win + (1u << 20)) : host_cap;
        if (ctx->acc_cap_hint > acc_cap) acc_cap = std::min<uint64_t>(host_cap, ctx->acc_cap_hint);
        CK(ctx->d_acc.reserve(acc_cap * 8 + 64));
        CK(ctx->d_acc2.reserve(acc_cap * 8 + 64));
        CK(ctx->d_acc_cnt.reserve(64));
        CK(cudaMemsetAsync(ctx->d_acc_cnt.p, 0, 64, s_k));
        host_cap = 0;
    }
    CK(ctx->h_val.reserve(host_cap * 8 + 8));
    const uint32_t pw = pos_width_of(*p);
    if (want_pos) CK(ctx->h_pos.reserve(host_cap * pw + 4));
    uint64_t *h_ooff = (uint64_t *)ctx->h_ooff.p;
    volatile uint64_t *h_meta = (volatile uint64_t *)ctx->h_meta.p;
    if (n_reads == 0) {
        h_ooff[0] = 0;
        if (out_val) *out_val = (uint64_t *)ctx->h_val.p;
        if (out_pos) *out_pos = want_pos ? (uint32_t *)ctx->h_pos.p : nullptr;
        if (out_off) *out_off = h_ooff;
        if (read_status) *read_status = (int32_t *)ctx->h_status.p;
        if (n_out) *n_out = 0;
        return 0;
    }
    // sub-batch boundaries: about kSubBytes of bases each
    uint64_t kSubBytes = 384ull << 20;
    if (const char *e = getenv("B200SK_SUB_BYTES")) { // testing knob: force many small sub-batches
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= 16) kSubBytes = v;
    }
    std::vector<uint64_t> cut;
    cut.push_back(0);
    while (cut.back() < n_reads) {
        const uint64_t r0 = cut.back();
        const uint64_t target = read_off[r0] + kSubBytes;
        uint64_t r1 = (uint64_t)(std::upper_bound(read_off + r0, read_off + n_reads + 1, target) - read_off);
        if (r1 > r0 + 1) r1--; // last read whose end is <= target ... but always advance by at least one read
        if (r1 <= r0) r1 = r0 + 1;
        if (r1 > n_reads) r1 = n_reads;
        // very short reads: bound the read count too so offsets stay a small share of the slot
        if (r1 - r0 > (64ull << 20)) r1 = r0 + (64ull << 20);
        cut.push_back(r1);
    }
    const size_t n_sub = cut.size() - 1;
    uint64_t max_sub_bases = 0, max_sub_reads = 0;
    for (size_t j = 0; j < n_sub; j++) {
        max_sub_bases = std::max<uint64_t>(max_sub_bases, read_off[cut[j + 1]] - (read_off[cut[j]] & ~15ull));
        max_sub_reads = std::max<uint64_t>(max_sub_reads, cut[j + 1] - cut[j]);
    }
    struct Slot {
        DevBuf bases, off, val, pos, ooff, status;
        cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
        uint64_t cap = 0;
        bool out_pending = false;
    };
    Slot slot[2];
    auto free_slots = [&]() {
        for (auto &sl : slot) {
            sl.bases.release(); sl.off.release(); sl.val.release(); sl.pos.release(); sl.ooff.release();
            sl.status.release();
            if (sl.in_done) cudaEventDestroy(sl.in_done);
            if (sl.k_done) cudaEventDestroy(sl.k_done);
            if (sl.out_done) cudaEventDestroy(sl.out_done);
        }
    };
    // the slot buffers persist in the ctx between calls (allocation is not free): move them in and out
    DevBuf *persist[2][6] = {{&ctx->d_bases, &ctx->d_off, &ctx->d_val, &ctx->d_pos, &ctx->d_ooff, &ctx->d_status},
                             {&ctx->d_bases2, &ctx->d_off2, &ctx->d_val2, &ctx->d_pos2, &ctx->d_ooff2, &ctx->d_status2}};
    for (int i = 0; i < 2; i++) {
        slot[i].bases = *persist[i][0]; slot[i].off = *persist[i][1]; slot[i].val = *persist[i][2];
        slot[i].pos = *persist[i][3]; slot[i].ooff = *persist[i][4]; slot[i].status = *persist[i][5];
        for (int q = 0; q < 6; q++) { persist[i][q]->p = nullptr; persist[i][q]->cap = 0; }
        const uint64_t cv = slot[i].val.cap >= 64 ? (slot[i].val.cap - 64) / 8 : 0;
        const uint64_t cp = slot[i].pos.cap >= 64 ? (slot[i].pos.cap - 64) / 4 : 0;
        slot[i].cap = want_pos ? std::min(cv, cp) : cv;
    }
    auto save_slots = [&]() {
        for (int i = 0; i < 2; i++) {
            *persist[i][0] = slot[i].bases; *persist[i][1] = slot[i].off; *persist[i][2] = slot[i].val;
            *persist[i][3] = slot[i].pos; *persist[i][4] = slot[i].ooff; *persist[i][5] = slot[i].status;
            slot[i].bases = DevBuf(); slot[i].off = DevBuf(); slot[i].val = DevBuf(); slot[i].pos = DevBuf();
            slot[i].ooff = DevBuf(); slot[i].status = DevBuf();
        }
        free_slots();
    };
#define CKS(call)                                                 \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) {                                  \
            cudaDeviceSynchronize();                              \
            save_slots();                                         \
            return cuda_fail(ctx, _e, #call);                     \
        }                                                         \
    } while (0)
    for (auto &sl : slot) {
        CKS(cudaEventCreateWithFlags(&sl.in_done, cudaEventDisableTiming));
        CKS(cudaEventCreateWithFlags(&sl.k_done, cudaEventDisableTiming));
        CKS(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
        CKS(sl.bases.reserve(max_sub_bases + 64));
        CKS(sl.off.reserve((max_sub_reads + 1) * 8));
        CKS(sl.ooff.reserve((max_sub_reads + 1) * 8));
        CKS(sl.status.reserve((max_sub_reads + 1) * 4));
    }
    auto issue_h2d = [&](size_t j) -> cudaError_t {
        Slot &sl = slot[j & 1];
        const uint64_t r0 = cut[j], r1 = cut[j + 1];
        const uint64_t b0 = read_off[r0] & ~15ull, b1 = read_off[r1];
        cudaError_t e;
        if (b1 > b0 && (e = cudaMemcpyAsync(sl.bases.p, bases + b0, b1 - b0, cudaMemcpyHostToDevice, s_in)) != cudaSuccess)
            return e;
        if ((e = cudaMemcpyAsync(sl.off.p, read_off + r0, (r1 - r0 + 1) * 8, cudaMemcpyHostToDevice, s_in)) != cudaSuccess)
            return e;
        return cudaEventRecord(sl.in_done, s_in);
    };
    uint64_t running = 0;
    CKS(issue_h2d(0));
    for (size_t j = 0; j < n_sub; j++) {
        Slot &sl = slot[j & 1];
        const uint64_t r0 = cut[j], r1 = cut[j + 1], nr = r1 - r0;
        const uint64_t b0 = read_off[r0] & ~15ull, nb = read_off[r1] - read_off[r0];
        // the other slot's kernel (sub-batch j-1) has been waited for below, so its inputs are free
        if (j + 1 < n_sub) CKS(issue_h2d(j + 1));
        CKS(cudaStreamWaitEvent(s_k, sl.in_done, 0));
        if (sl.out_pending) CKS(cudaStreamWaitEvent(s_k, sl.out_done, 0)); // slot outputs still draining
        uint64_t cap = std::max<uint64_t>(sl.cap, b200sk_output_bound(p, nb, nr, 0));
        uint64_t total = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            if (cap > sl.cap) {
                if (sl.out_pending) CKS(cudaEventSynchronize(sl.out_done));
                CKS(sl.val.reserve(cap * 8 + 64));
                if (want_pos) CKS(sl.pos.reserve(cap * 4 + 64));
                sl.cap = cap;
            }
            // offsets stay absolute: the kernels add read_off[r] >= b0 to a base address b0 bytes below the slot's
            // buffer (formed in integer arithmetic: it is not a pointer into any object until an offset is added)
            const uint8_t *dbase = (const uint8_t *)((uintptr_t)sl.bases.p - (uintptr_t)b0);
            rc = enqueue(ctx, *p, dbase, (const uint64_t *)sl.off.p, nr, nb, (uint64_t *)sl.val.p,
                         want_pos ? sl.pos.p : nullptr, (uint64_t *)sl.ooff.p, (int32_t *)sl.status.p,
                         sl.cap, running, s_k, nullptr);
            if (rc) { cudaDeviceSynchronize(); save_slots(); return rc; }
            CKS(cudaMemcpyAsync((void *)h_meta, (uint64_t *)sl.ooff.p + nr, 8, cudaMemcpyDeviceToHost, s_k));
            CKS(cudaMemcpyAsync((void *)(h_meta + 1), (unsigned long long *)ctx->meta.p + 1, 8, cudaMemcpyDeviceToHost, s_k));
            CKS(cudaEventRecord(sl.k_done, s_k));
            CKS(cudaEventSynchronize(sl.k_done));
            total = h_meta[0] - running;
            const uint64_t flags = h_meta[1];
            if (flags & B200SK_FLAG_SPAN) { cudaDeviceSynchronize(); save_slots(); return B200SK_ERR_BAD_ARG; }
            if (!(flags & B200SK_FLAG_CAPACITY)) break;
            cap = total; // exact requirement reported by the first pass
            if (attempt == 1) { cudaDeviceSynchronize(); save_slots(); return B200SK_ERR_CAPACITY; }
        }
        if (reduce) {
            // nothing of this sub-batch goes to the host: append the kept fraction of its values to the accumulation array
            if (total) {
                CKS(launch_filter_scale((const uint64_t *)sl.val.p, total, max_hash, (uint64_t *)ctx->d_acc.p, acc_cap,
                                        (unsigned long long *)ctx->d_acc_cnt.p, s_k));
                ctx->launches++;
            }
            CKS(cudaEventRecord(sl.out_done, s_k)); // the slot's outputs are free once the filter has read them
            sl.out_pending = true;
            running += total;
            continue;
        }
        if (running + total > host_cap) { // estimate too small: grow the pinned result arrays, keeping their content
            CKS(cudaStreamSynchronize(s_out));
            host_cap = (running + total) + (running + total) / 4 + 1024;
            CKS(ctx->h_val.reserve(host_cap * 8 + 8, true));
            if (want_pos) CKS(ctx->h_pos.reserve(host_cap * pw + 4, true));
        }
        CKS(cudaStreamWaitEvent(s_out, sl.k_done, 0));
        if (total) {
            CKS(cudaMemcpyAsync((uint64_t *)ctx->h_val.p + running, sl.val.p, total * 8, cudaMemcpyDeviceToHost, s_out));
            if (want_pos)
                CKS(cudaMemcpyAsync((uint8_t *)ctx->h_pos.p + running * pw, sl.pos.p, total * pw, cudaMemcpyDeviceToHost, s_out));
        }
        CKS(cudaMemcpyAsync(h_ooff + r0, sl.ooff.p, (nr + 1) * 8, cudaMemcpyDeviceToHost, s_out));
        CKS(cudaMemcpyAsync((int32_t *)ctx->h_status.p + r0, sl.status.p, nr * 4, cudaMemcpyDeviceToHost, s_out));
        CKS(cudaEventRecord(sl.out_done, s_out));
        sl.out_pending = true;
        running += total;
    }
    CKS(cudaStreamSynchronize(s_out));
    if (reduce) {
        unsigned long long kept = 0;
        CKS(cudaMemcpyAsync(&kept, ctx->d_acc_cnt.p, 8, cudaMemcpyDeviceToHost, s_k));
        CKS(cudaStreamSynchronize(s_k));
        save_slots();
        if (kept > acc_cap) { // the fraction was larger than planned: once more with the size just measured
            if (ctx->acc_cap_hint >= kept) return B200SK_ERR_CAPACITY;
            ctx->acc_cap_hint = kept + kept / 8 + 1024;
            return run_host(ctx, p_in, bases, read_off, n_reads, out_val, out_pos, out_off, read_status, n_out, reduce);
        }
        uint64_t m = 0;
        rc = b200sk_reduce_device(ctx, (uint64_t *)ctx->d_acc.p, kept, reduce->scale, reduce->unique, (uint64_t *)ctx->d_acc2.p,
                                  acc_cap, &m, s_k);
        if (rc) return rc;
        CK(ctx->h_val.reserve(m * 8 + 8));
        if (m) CK(cudaMemcpyAsync(ctx->h_val.p, ctx->d_acc2.p, m * 8, cudaMemcpyDeviceToHost, s_k));
        CK(cudaStreamSynchronize(s_k));
        if (out_val) *out_val = (uint64_t *)ctx->h_val.p;
        if (n_out) *n_out = m;
        return 0;
    }
    save_slots();
#undef CKS
    if (out_val) *out_val = (uint64_t *)ctx->h_val.p;
    if (out_pos) *out_pos = want_pos ? (uint32_t *)ctx->h_pos.p : nullptr;
    if (out_off) *out_off = h_ooff;
    if (read_status) *read_status = (int32_t *)ctx->h_status.p;
    if (n_out) *n_out = running;
    return 0;
}

int b200sk_run(b200sk_ctx *ctx, const b200sk_params *p, const uint8_t *bases, const uint64_t *read_off,
               uint64_t n_reads, uint64_t **out_val, uint32_t **out_pos, uint64_t **out_off,
               int32_t **read_status, uint64_t *n_out) {
    return run_host(ctx, p, bases, read_off, n_reads, out_val, out_pos, out_off, read_status, n_out, nullptr);
}

int b200sk_run_reduced(b200sk_ctx *ctx, const b200sk_params *p, uint32_t scale, int unique, const uint8_t *bases,
                       const uint64_t *read_off, uint64_t n_reads, uint64_t **out_val, uint64_t *n_out) {
    const ReduceOpt ro = {scale, unique};
    return run_host(ctx, p, bases, read_off, n_reads, out_val, nullptr, nullptr, nullptr, n_out, &ro);
}

} // extern "C"
