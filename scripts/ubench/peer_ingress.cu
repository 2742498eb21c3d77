// Micro-benchmark (dev helper): how fast can N-1 GPUs write into ONE GPU's HBM over NVLink, by store flavour?
//   V0 cudaMemcpyPeerAsync (copy engines)      V1 kernel, 8-byte st.global per lane (256 B per warp instruction)
//   V2 kernel, 16-byte st.global per lane       V3 kernel, TMA bulk stores shared -> peer global, 4 KB per copy
//   V4 like V1 but every warp store shifted by 8 bytes (no store starts on a 32-byte sector)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o peer_ingress peer_ingress.cu && ./peer_ingress [n_gpus]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <chrono>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_st8(uint64_t *dst, const uint64_t *src, size_t n, int shift) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + shift;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}
__global__ void k_st16(ulonglong2 *dst, const ulonglong2 *src, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// every warp: 4 KB tile local global -> shared (plain loads), then ONE bulk store shared -> peer global
__global__ void __launch_bounds__(256) k_tma(uint8_t *dst, const uint8_t *src, size_t bytes) {
    extern __shared__ __align__(128) uint8_t sm[];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint8_t *buf = sm + wid * 4096;
    const size_t nw = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + wid; t * 4096 < bytes; t += nw) {
        const uint4 *s = reinterpret_cast<const uint4 *>(src + t * 4096);
        uint4 *b = reinterpret_cast<uint4 *>(buf);
        // the previous bulk store must have read the buffer before it is overwritten
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) b[lane + 32 * i] = s[lane + 32 * i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * 4096), "r"(smem_u32(buf)), "r"(4096u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the sketching kernel's pattern: every warp owns a run of RUN elements (5.7 KB), writes it with consecutive 256-byte
// warp stores, then jumps to its next run far away (runs handed out round-robin over the warps)
#define RUN 712
__global__ void __launch_bounds__(512) k_runs(uint64_t *dst, const uint64_t *src, size_t n, int shift) {
    const uint32_t lane = threadIdx.x & 31u;
    const size_t nw = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); (t + 1) * RUN + 16 <= n; t += nw) {
        const size_t b = t * RUN + shift;
        for (uint32_t i = lane; i < RUN; i += 32) dst[b + i] = src[b + i];
    }
}
// the same runs, each as ONE bulk store from shared memory (5 696 bytes, 16-byte aligned)
__global__ void __launch_bounds__(512) k_runs_tma(uint64_t *dst, const uint64_t *src, size_t n) {
    extern __shared__ __align__(128) uint8_t sm[];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint64_t *buf = reinterpret_cast<uint64_t *>(sm + wid * 5760);
    const size_t nw = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + wid; (t + 1) * RUN + 16 <= n; t += nw) {
        const size_t b = t * RUN;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (uint32_t i = lane; i < RUN; i += 32) buf[i] = src[b + i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + b), "r"(smem_u32(buf)), "r"((uint32_t)(RUN * 8)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    int n = argc > 1 ? atoi(argv[1]) : ndev;
    if (n > ndev) n = ndev;
    if (n < 2) { printf("needs >= 2 GPUs\n"); return 0; }
    const size_t per = (size_t)2 << 30; // bytes per sender
    uint8_t *dst;
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&dst, per * (n - 1)));
    std::vector<uint8_t *> src(n);
    std::vector<cudaStream_t> st(n);
    for (int d = 1; d < n; d++) {
        CK(cudaSetDevice(d));
        CK(cudaMalloc(&src[d], per));
        CK(cudaMemset(src[d], d, per));
        CK(cudaDeviceEnablePeerAccess(0, 0));
        CK(cudaStreamCreate(&st[d]));
        CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096));
        CK(cudaFuncSetAttribute(k_runs_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 5760));
    }
    const char *names[] = {"cudaMemcpyPeerAsync", "kernel 8 B/lane", "kernel 16 B/lane", "kernel TMA bulk 4 KB", "kernel 8 B/lane, unaligned",
                           "kernel, 5.7 KB runs per warp", "kernel, runs, 8 B off a line", "kernel, runs as TMA bulk stores"};
    for (int senders = 1; senders < n; senders = senders == n - 1 ? n : (senders * 2 > n - 1 ? n - 1 : senders * 2)) {
        for (int v = 0; v < 8; v++) {
            double best = 0;
            for (int rep = 0; rep < 3; rep++) {
                for (int d = 1; d <= senders; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
                auto t0 = std::chrono::steady_clock::now();
                for (int d = 1; d <= senders; d++) {
                    CK(cudaSetDevice(d));
                    uint8_t *to = dst + per * (d - 1);
                    if (v == 0) CK(cudaMemcpyPeerAsync(to, 0, src[d], d, per, st[d]));
                    else if (v == 1) k_st8<<<148 * 8, 256, 0, st[d]>>>((uint64_t *)to, (const uint64_t *)src[d], per / 8, 0);
                    else if (v == 2) k_st16<<<148 * 8, 256, 0, st[d]>>>((ulonglong2 *)to, (const ulonglong2 *)src[d], per / 16);
                    else if (v == 3) k_tma<<<148 * 4, 256, 8 * 4096, st[d]>>>(to, src[d], per);
                    else if (v == 4) k_st8<<<148 * 8, 256, 0, st[d]>>>((uint64_t *)to, (const uint64_t *)src[d], per / 8 - 1, 1);
                    else if (v == 5) k_runs<<<148, 512, 0, st[d]>>>((uint64_t *)to, (const uint64_t *)src[d], per / 8, 0);
                    else if (v == 6) k_runs<<<148, 512, 0, st[d]>>>((uint64_t *)to, (const uint64_t *)src[d], per / 8, 1);
                    else k_runs_tma<<<148, 512, 16 * 5760, st[d]>>>((uint64_t *)to, (const uint64_t *)src[d], per / 8);
                }
                for (int d = 1; d <= senders; d++) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
                const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                const double gbs = per * (double)senders / s / 1e9;
                if (gbs > best) best = gbs;
            }
            printf("senders=%d  %-28s  %.0f GB/s into GPU 0\n", senders, names[v], best);
        }
    }
    return 0;
}
