// Multi-GPU side of libb200sketch.so: records shard over devices with no cross-record state
// (sketches/iterator.go:615-655, sketches/sketch.go:85-202), so the only exchange is the gather of the per-GPU
// uint64 hash arrays.  Two shapes, both declared in include/b200sketch.h:
//
//  * one process per GPU (the torchrun shape): the root exports ONE gather buffer over CUDA IPC, every other
//    rank maps it and hands a pointer INTO it to b200sk_enqueue_device as its out_val -- the sketching
//    kernel's own coalesced flush then stores the minimizers straight into the root's HBM over NVLink (peer
//    st.global from the same kernel that hashes), tile by tile while the rest of the shard is still being
//    walked.  No staging copy, no collective on the data path; only the per-rank element counts (8 bytes each)
//    travel through whatever the host side uses (NCCL / gloo in bench.py and the tests).  Rank segments sit at
//    bases sized by b200sk_output_bound; b200sk_compact_segments closes the gaps on the root when a contiguous
//    array is wanted.
//
//  * one process, several devices (what a Go caller would use: b200sk_group_*): one context and one worker
//    thread per device, reads sharded by cumulative bases, results assembled in read order in one pinned array.
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200sketch.h"

namespace b200sk {
int ctx_device(b200sk_ctx *ctx);
void ctx_set_error(b200sk_ctx *ctx, const char *msg);
void ctx_add_launches(b200sk_ctx *ctx, uint64_t n);
} // namespace b200sk

namespace {

int mfail(b200sk_ctx *ctx, cudaError_t e, const char *what) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    if (ctx) b200sk::ctx_set_error(ctx, buf);
    return B200SK_ERR_CUDA;
}
#define MCK(call)                                           \
    do {                                                    \
        cudaError_t _e = (call);                            \
        if (_e != cudaSuccess) return mfail(ctx, _e, #call); \
    } while (0)

// dst[i] = src[i] for i in [0, n), 16 bytes per thread and trip; the ranges must not overlap
__global__ void __launch_bounds__(256) k_move_u64(uint64_t *__restrict__ dst, const uint64_t *__restrict__ src, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // heads and tails that are not 16-byte aligned relative to each other go element by element
    if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
        const uint64_t n2 = n / 2;
        const ulonglong2 *s2 = reinterpret_cast<const ulonglong2 *>(src);
        ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(dst);
        for (uint64_t j = i; j < n2; j += stride) d2[j] = s2[j];
        if (i == 0 && (n & 1)) dst[n - 1] = src[n - 1];
    } else {
        for (uint64_t j = i; j < n; j += stride) dst[j] = src[j];
    }
}
__global__ void __launch_bounds__(256) k_add_u64(uint64_t *a, uint64_t n, uint64_t add) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] += add;
}

} // namespace

extern "C" {

// ------------------------------------------------------------------ gather buffer over CUDA IPC
int b200sk_gather_create(b200sk_ctx *ctx, uint64_t capacity_elems, uint8_t *handle, uint64_t **d_buf) {
    if (!ctx || !handle || !d_buf) return B200SK_ERR_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == B200SK_IPC_HANDLE_BYTES, "handle size");
    MCK(cudaSetDevice(b200sk::ctx_device(ctx)));
    void *p = nullptr;
    MCK(cudaMalloc(&p, (capacity_elems ? capacity_elems : 1) * 8 + 64));
    MCK(cudaMemset(p, 0, (capacity_elems ? capacity_elems : 1) * 8 + 64)); // (status words start out empty)
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return mfail(ctx, e, "cudaIpcGetMemHandle");
    }
    memcpy(handle, &h, sizeof(h));
    *d_buf = (uint64_t *)p;
    return 0;
}

int b200sk_gather_open(b200sk_ctx *ctx, const uint8_t *handle, uint64_t **d_buf) {
    if (!ctx || !handle || !d_buf) return B200SK_ERR_BAD_ARG;
    MCK(cudaSetDevice(b200sk::ctx_device(ctx)));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    MCK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess)); // peer mapping of the root's allocation
    *d_buf = (uint64_t *)p;
    return 0;
}

int b200sk_gather_close(b200sk_ctx *ctx, uint64_t *d_buf, int is_owner) {
    if (!ctx) return B200SK_ERR_BAD_ARG;
    if (!d_buf) return 0;
    MCK(cudaSetDevice(b200sk::ctx_device(ctx)));
    if (is_owner) MCK(cudaFree(d_buf));
    else MCK(cudaIpcCloseMemHandle(d_buf));
    return 0;
}

// Close the gaps between rank segments in place: segment r = d_buf[seg_base[r] .. +seg_count[r]) moves down to
// the end of segment r-1.  A segment moves in waves no longer than its distance to the destination, so source
// and destination of one launch never overlap.
int b200sk_compact_segments(b200sk_ctx *ctx, uint64_t *d_buf, const uint64_t *seg_base, const uint64_t *seg_count,
                            int n_seg, void *stream) {
    if (!ctx || !d_buf || !seg_base || !seg_count || n_seg < 1) return B200SK_ERR_BAD_ARG;
    MCK(cudaSetDevice(b200sk::ctx_device(ctx)));
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t end = seg_base[0] + seg_count[0];
    if (seg_base[0] != 0) return B200SK_ERR_BAD_ARG;
    uint64_t launches = 0;
    for (int r = 1; r < n_seg; r++) {
        if (seg_base[r] < end) return B200SK_ERR_BAD_ARG; // segments must be ordered and disjoint
        const uint64_t gap = seg_base[r] - end;
        if (gap) {
            uint64_t done = 0;
            while (done < seg_count[r]) {
                const uint64_t n = std::min<uint64_t>(gap, seg_count[r] - done);
                const unsigned blocks = (unsigned)std::min<uint64_t>((n / 2 + 255) / 256 + 1, 148ull * 16);
                k_move_u64<<<blocks, 256, 0, st>>>(d_buf + end + done, d_buf + seg_base[r] + done, n);
                launches++;
                done += n;
            }
            MCK(cudaGetLastError());
        }
        end += seg_count[r];
    }
    b200sk::ctx_add_launches(ctx, launches);
    return 0;
}

} // extern "C"

// ------------------------------------------------------------------ one process, several devices
struct b200sk_group {
    std::vector<int> devices;
    std::vector<b200sk_ctx *> ctx;
    struct Dev {
        uint8_t *d_bases = nullptr; size_t cap_bases = 0;
        uint64_t *d_off = nullptr, *d_ooff = nullptr; size_t cap_reads = 0;
        int32_t *d_status = nullptr;
        uint64_t *d_val = nullptr; void *d_pos = nullptr; size_t cap_out = 0; bool has_pos = false;
        cudaStream_t st = nullptr;
    };
    std::vector<Dev> dev;
    // pinned result arrays (read order)
    uint64_t *h_val = nullptr; void *h_pos = nullptr; uint64_t *h_ooff = nullptr; int32_t *h_status = nullptr;
    size_t cap_val = 0, cap_pos = 0, cap_reads = 0;
    std::string last_error;
};

namespace {

template <class T> cudaError_t grow(T *&p, size_t &cap, size_t want_elems, size_t elem) {
    if (want_elems <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t w = want_elems + want_elems / 8 + 64;
    cudaError_t e = cudaMalloc((void **)&p, w * elem);
    if (e == cudaSuccess) cap = w;
    return e;
}
template <class T> cudaError_t grow_pinned(T *&p, size_t &cap, size_t want_bytes) {
    if (want_bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t w = want_bytes + want_bytes / 8 + 256;
    cudaError_t e = cudaHostAlloc((void **)&p, w, cudaHostAllocPortable);
    if (e == cudaSuccess) cap = w;
    return e;
}

} // namespace

extern "C" {

int b200sk_group_create(b200sk_group **out, const int *devices, int n_devices) {
    if (!out || !devices || n_devices < 1) return B200SK_ERR_BAD_ARG;
    *out = nullptr;
    b200sk_group *g = new b200sk_group();
    for (int i = 0; i < n_devices; i++) {
        b200sk_ctx *c = nullptr;
        int rc = b200sk_create(&c, devices[i]);
        if (rc) {
            for (b200sk_ctx *q : g->ctx) b200sk_destroy(q);
            delete g;
            return rc;
        }
        g->devices.push_back(devices[i]);
        g->ctx.push_back(c);
    }
    g->dev.resize(n_devices);
    for (int i = 0; i < n_devices; i++) {
        cudaSetDevice(devices[i]);
        if (cudaStreamCreateWithFlags(&g->dev[i].st, cudaStreamNonBlocking) != cudaSuccess) {
            b200sk_group_destroy(g);
            return B200SK_ERR_CUDA;
        }
    }
    *out = g;
    return 0;
}

void b200sk_group_destroy(b200sk_group *g) {
    if (!g) return;
    for (size_t i = 0; i < g->dev.size(); i++) {
        cudaSetDevice(g->devices[i]);
        b200sk_group::Dev &d = g->dev[i];
        if (d.st) { cudaStreamSynchronize(d.st); cudaStreamDestroy(d.st); }
        for (void *p : {(void *)d.d_bases, (void *)d.d_off, (void *)d.d_ooff, (void *)d.d_status, (void *)d.d_val, d.d_pos})
            if (p) cudaFree(p);
    }
    for (b200sk_ctx *c : g->ctx) b200sk_destroy(c);
    for (void *p : {(void *)g->h_val, g->h_pos, (void *)g->h_ooff, (void *)g->h_status})
        if (p) cudaFreeHost(p);
    delete g;
}

int b200sk_group_size(const b200sk_group *g) { return g ? (int)g->ctx.size() : 0; }
const char *b200sk_group_last_error(const b200sk_group *g) { return g ? g->last_error.c_str() : ""; }
uint64_t b200sk_group_kernel_launches(const b200sk_group *g) {
    uint64_t n = 0;
    if (g)
        for (b200sk_ctx *c : g->ctx) n += b200sk_kernel_launches(c);
    return n;
}

// Shard boundaries balanced by cumulative bases: device d owns reads [cut[d], cut[d+1]).
void b200sk_shard_by_bases(const uint64_t *read_off, uint64_t n_reads, int n_shards, uint64_t *cut) {
    cut[0] = 0;
    const uint64_t b0 = read_off[0], total = read_off[n_reads] - b0;
    for (int d = 1; d < n_shards; d++) {
        const uint64_t target = b0 + (uint64_t)((unsigned __int128)total * (unsigned)d / (unsigned)n_shards);
        uint64_t r = (uint64_t)(std::lower_bound(read_off, read_off + n_reads + 1, target) - read_off);
        if (r > n_reads) r = n_reads;
        cut[d] = std::max(r, cut[d - 1]);
    }
    cut[n_shards] = n_reads;
}

// Replaces the per-record loop over ALL records of a batch with every device of the group at work: phase 1 copies
// each shard to its device and sketches it there (one worker thread per device); the element counts then fix
// where every shard lands in the result, and phase 2 copies each device's arrays straight to that place.
int b200sk_group_run(b200sk_group *g, const b200sk_params *p, const uint8_t *bases, const uint64_t *read_off,
                     uint64_t n_reads, uint64_t **out_val, uint32_t **out_pos, uint64_t **out_off,
                     int32_t **read_status, uint64_t *n_out) {
    if (!g || !p || !read_off) return B200SK_ERR_BAD_ARG;
    int rc = b200sk_check_params(p);
    if (rc) return rc;
    const int nd = (int)g->ctx.size();
    const bool want_pos = p->want_pos != 0;
    const uint32_t pw = p->pos_width == 1 ? 1u : p->pos_width == 2 ? 2u : 4u;
    std::vector<uint64_t> cut(nd + 1);
    b200sk_shard_by_bases(read_off, n_reads, nd, cut.data());
    std::vector<uint64_t> count(nd, 0);
    std::vector<int> status(nd, 0);
    std::vector<std::string> err(nd);
    auto phase1 = [&](int d) {
        b200sk_group::Dev &D = g->dev[d];
        b200sk_ctx *ctx = g->ctx[d];
        const uint64_t r0 = cut[d], r1 = cut[d + 1], nr = r1 - r0;
        auto fail = [&](cudaError_t e, const char *what) { err[d] = std::string(what) + ": " + cudaGetErrorString(e); status[d] = B200SK_ERR_CUDA; };
#define GCK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { fail(_e, #call); return; } } while (0)
        GCK(cudaSetDevice(g->devices[d]));
        if (nr == 0) return;
        const uint64_t b0 = read_off[r0] & ~15ull, b1 = read_off[r1], nb = read_off[r1] - read_off[r0];
        GCK(grow(D.d_bases, D.cap_bases, (size_t)(b1 - b0) + 64, 1));
        if (nr + 1 > D.cap_reads) {
            for (void *q : {(void *)D.d_off, (void *)D.d_ooff, (void *)D.d_status}) if (q) cudaFree(q);
            D.d_off = D.d_ooff = nullptr; D.d_status = nullptr; D.cap_reads = 0;
            const size_t w = (size_t)(nr + 1) + (size_t)(nr + 1) / 8 + 64;
            GCK(cudaMalloc((void **)&D.d_off, w * 8));
            GCK(cudaMalloc((void **)&D.d_ooff, w * 8));
            GCK(cudaMalloc((void **)&D.d_status, w * 4));
            D.cap_reads = w;
        }
        GCK(cudaMemcpyAsync(D.d_bases, bases + b0, b1 - b0, cudaMemcpyHostToDevice, D.st));
        GCK(cudaMemcpyAsync(D.d_off, read_off + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, D.st));
        uint64_t cap = std::max<uint64_t>(D.cap_out, b200sk_output_bound(p, nb, nr, 0));
        for (int attempt = 0; attempt < 2; attempt++) {
            if (cap > D.cap_out || (want_pos && !D.has_pos)) {
                if (D.d_val) cudaFree(D.d_val);
                if (D.d_pos) cudaFree(D.d_pos);
                D.d_val = nullptr; D.d_pos = nullptr; D.cap_out = 0; D.has_pos = false;
                GCK(cudaMalloc((void **)&D.d_val, cap * 8 + 64));
                if (want_pos) GCK(cudaMalloc(&D.d_pos, cap * 4 + 64));
                D.cap_out = cap; D.has_pos = want_pos;
            }
            uint64_t total = 0;
            // offsets stay absolute: the kernels index bases with read_off, so hand them the shifted base pointer
            const int r = b200sk_run_device(ctx, p, D.d_bases - b0, D.d_off, nr, nb, D.d_val, (uint32_t *)(want_pos ? D.d_pos : nullptr),
                                            D.d_ooff, D.d_status, D.cap_out, D.st, &total);
            count[d] = total;
            if (r == B200SK_ERR_CAPACITY && attempt == 0) { cap = total; continue; }
            if (r) { status[d] = r; if (r == B200SK_ERR_CUDA) err[d] = b200sk_last_error(ctx); }
            break;
        }
#undef GCK
    };
    {
        std::vector<std::thread> th;
        for (int d = 1; d < nd; d++) th.emplace_back(phase1, d);
        phase1(0);
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < nd; d++)
        if (status[d]) { g->last_error = err[d]; return status[d]; }
    uint64_t total = 0;
    std::vector<uint64_t> base(nd + 1, 0);
    for (int d = 0; d < nd; d++) { base[d] = total; total += count[d]; }
    base[nd] = total;
    b200sk_ctx *ctx = g->ctx[0];
    MCK(cudaSetDevice(g->devices[0]));
    MCK(grow_pinned(g->h_val, g->cap_val, total * 8 + 8));
    if (want_pos) MCK(grow_pinned(g->h_pos, g->cap_pos, total * pw + 8));
    if ((n_reads + 1) > g->cap_reads) {
        if (g->h_ooff) cudaFreeHost(g->h_ooff);
        if (g->h_status) cudaFreeHost(g->h_status);
        g->h_ooff = nullptr; g->h_status = nullptr; g->cap_reads = 0;
        const size_t w = (size_t)(n_reads + 1) + (size_t)(n_reads + 1) / 8 + 64;
        MCK(cudaHostAlloc((void **)&g->h_ooff, w * 8, cudaHostAllocPortable));
        MCK(cudaHostAlloc((void **)&g->h_status, w * 4, cudaHostAllocPortable));
        g->cap_reads = w;
    }
    auto phase2 = [&](int d) {
        b200sk_group::Dev &D = g->dev[d];
        const uint64_t r0 = cut[d], nr = cut[d + 1] - r0;
        if (cudaSetDevice(g->devices[d]) != cudaSuccess) { status[d] = B200SK_ERR_CUDA; return; }
        if (nr == 0) return;
        cudaError_t e = cudaSuccess;
        if (base[d]) {
            k_add_u64<<<(unsigned)std::min<uint64_t>((nr + 256) / 256, 148ull * 8), 256, 0, D.st>>>(D.d_ooff, nr + 1, base[d]);
            b200sk::ctx_add_launches(g->ctx[d], 1);
        }
        if (count[d]) {
            e = cudaMemcpyAsync(g->h_val + base[d], D.d_val, count[d] * 8, cudaMemcpyDeviceToHost, D.st);
            if (e == cudaSuccess && want_pos)
                e = cudaMemcpyAsync((uint8_t *)g->h_pos + base[d] * pw, D.d_pos, count[d] * pw, cudaMemcpyDeviceToHost, D.st);
        }
        // shard d's last offset entry equals shard d+1's first: the later copy writes the same value again
        if (e == cudaSuccess) e = cudaMemcpyAsync(g->h_ooff + r0, D.d_ooff, (nr + 1) * 8, cudaMemcpyDeviceToHost, D.st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(g->h_status + r0, D.d_status, nr * 4, cudaMemcpyDeviceToHost, D.st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(D.st);
        if (e != cudaSuccess) { status[d] = B200SK_ERR_CUDA; err[d] = cudaGetErrorString(e); }
    };
    {
        std::vector<std::thread> th;
        for (int d = 1; d < nd; d++) th.emplace_back(phase2, d);
        phase2(0);
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < nd; d++)
        if (status[d]) { g->last_error = err[d]; return status[d]; }
    if (n_reads == 0 || cut[nd] == 0) g->h_ooff[0] = 0;
    // empty shards leave holes in the offset table: fill them with the running total
    for (int d = 0; d < nd; d++)
        if (cut[d + 1] == cut[d]) g->h_ooff[cut[d]] = base[d];
    g->h_ooff[n_reads] = total;
    if (out_val) *out_val = g->h_val;
    if (out_pos) *out_pos = want_pos ? (uint32_t *)g->h_pos : nullptr;
    if (out_off) *out_off = g->h_ooff;
    if (read_status) *read_status = g->h_status;
    if (n_out) *n_out = total;
    return 0;
}

} // extern "C"
