// Micro-benchmark: per-lane 128-byte cp.async.bulk shared->global stores (dev helper): is the TMA engine a
// usable exit for many small contiguous rows?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulkstore bulkstore.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int BYTES, int DEPTH>
__global__ void __launch_bounds__(512, 1) k(uint8_t *out, int iters) {
    extern __shared__ __align__(128) uint8_t sm[];
    const uint32_t tid = threadIdx.x;
    uint8_t *row = sm + tid * (BYTES + 16);
    for (int i = 0; i < BYTES / 8; i++) reinterpret_cast<uint64_t *>(row)[i] = tid * 1000 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t srow = (uint32_t)__cvta_generic_to_shared(row);
    uint8_t *dst = out + ((size_t)blockIdx.x * 512 + tid) * BYTES;
    const size_t step = (size_t)gridDim.x * 512 * BYTES;
    for (int it = 0; it < iters; it++) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(srow), "n"(BYTES) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if ((it % DEPTH) == DEPTH - 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        dst += step;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int BYTES>
__global__ void __launch_bounds__(512, 1) k_stg(uint8_t *out, int iters) { // same traffic with plain coalesced stores
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint64_t *dst = reinterpret_cast<uint64_t *>(out + ((size_t)blockIdx.x * 512 + w * 32) * BYTES);
    const size_t step = (size_t)gridDim.x * 512 * BYTES / 8;
    for (int it = 0; it < iters; it++) {
        for (int j = lane; j < 32 * BYTES / 8; j += 32) dst[j] = tid + j;
        dst += step;
    }
}
// the dense ntHash kernel's pattern: a warp owns 32 consecutive reads of 130 values (1040 B apart); per block of
// 16 values it writes 32 rows of 128 B, two rows per store instruction; ROWLEN values per row variant
template <int ROWLEN, bool ALIGN, int AL = 4>
__global__ void __launch_bounds__(736, 1) k_rows(uint8_t *out, long long ntiles, unsigned long long *ticket) {
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if ((long long)t >= ntiles) break;
        uint64_t *base = reinterpret_cast<uint64_t *>(out) + t * 32ull * 130ull;
        constexpr int LPR = ROWLEN;            // lanes per row (one value each)
        constexpr int RPI = 32 / LPR;          // rows per instruction
        for (int b = 0; b < 130 + ROWLEN; b += ROWLEN) {
            for (int i = 0; i < 32; i += RPI) {
                const int row = i + (int)lane / LPR, e = (int)lane % LPR;
                // ALIGN: virtual index v = u + shift with shift = element index of the read's start mod 4
                const int shift = ALIGN ? (int)((t * 32ull * 130ull + (unsigned long long)row * 130ull) & (unsigned long long)(AL - 1)) : 0;
                const int u = b + e - shift;
                if (u >= 0 && u < 130) base[row * 130 + u] = t + row + e;
            }
        }
    }
}
template <int BYTES, int DEPTH> void run(const char *name, uint8_t *out, size_t cap) {
    const int iters = (int)(cap / ((size_t)148 * 512 * BYTES));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k<BYTES, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * (BYTES + 16));
    k<BYTES, DEPTH><<<148, 512, 512 * (BYTES + 16)>>>(out, iters);
    cudaEventRecord(e0);
    k<BYTES, DEPTH><<<148, 512, 512 * (BYTES + 16)>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)iters * 148 * 512 * BYTES;
    printf("%-30s %6.2f ms  %7.1f GB/s  %6.1f M rows/s/SM  err=%s\n", name, ms, bytes / ms / 1e6, (double)iters * 512 / ms / 1e3, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    size_t cap = 8ull << 30;
    uint8_t *out; cudaMalloc(&out, cap);
    run<128, 1>("bulk 128 B, wait every op", out, cap);
    run<128, 4>("bulk 128 B, wait every 4", out, cap);
    run<256, 4>("bulk 256 B, wait every 4", out, cap);
    run<64, 4>("bulk 64 B, wait every 4", out, cap);
    {
        const int iters = (int)(cap / ((size_t)148 * 512 * 128));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_stg<128><<<148, 512>>>(out, iters);
        cudaEventRecord(e0); k_stg<128><<<148, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-30s %6.2f ms  %7.1f GB/s\n", "plain coalesced STG.64", ms, (double)iters * 148 * 512 * 128 / ms / 1e6);
    }
    {
        unsigned long long *ticket; cudaMalloc(&ticket, 8);
        const long long ntiles = 7000000 / 32;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int v = 0; v < 6; v++) {
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
                cudaMemset(ticket, 0, 8);
                cudaEventRecord(e0);
                if (v == 0) k_rows<16, false><<<148, 736>>>(out, ntiles, ticket); else if (v == 1) k_rows<32, false><<<148, 736>>>(out, ntiles, ticket); else if (v == 2) k_rows<16, true><<<148, 736>>>(out, ntiles, ticket); else if (v == 3) k_rows<32, true><<<148, 736>>>(out, ntiles, ticket); else if (v == 4) k_rows<16, true, 16><<<148, 736>>>(out, ntiles, ticket); else k_rows<32, true, 32><<<148, 736>>>(out, ntiles, ticket);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            printf("rows of %d values, stride 1040 B, %s   %6.2f ms  %7.1f GB/s\n", (v & 1) ? 32 : 16, v >= 4 ? "row = whole lines " : v >= 2 ? "32-B aligned rows" : "8-B aligned rows ", ms, 7000000.0 * 130 * 8 / ms / 1e6);
        }
    }
    return 0;
}
