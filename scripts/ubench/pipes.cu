// Micro-benchmark: issue rates of ALU-pipe vs FMA-pipe integer instructions on sm_100a (dev helper).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
template <int KIND>
__global__ void __launch_bounds__(512, 1) k(uint32_t *out, uint32_t one, uint32_t two, long long *cyc) {
    uint32_t a[8], b[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 7 + i; b[i] = threadIdx.x * 13 + i * 3; }
    uint64_t w[4];
    for (int i = 0; i < 4; i++) w[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0) { // LOP3 only
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(one));
            } else if (KIND == 1) { // IMAD only
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(two), "r"(b[i]));
            } else if (KIND == 2) { // IMAD.WIDE only
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i & 3]) : "r"(a[i]), "r"(two));
            } else if (KIND == 3) { // LOP3 + IMAD 1:1
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(one));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(two), "r"(one));
            } else if (KIND == 4) { // LOP3 + IMAD.WIDE 1:1
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(one));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i & 3]) : "r"(b[i]), "r"(two));
            } else if (KIND == 5) { // predicated IMAD + LOP3
                asm volatile("{.reg .pred p; setp.lt.u32 p, %1, %2; @p mad.lo.u32 %0, %1, %3, 0;}" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) & 7]), "r"(one));
            } else if (KIND == 6) { // setp + selp (ALU only)
                asm volatile("{.reg .pred p; setp.lt.u32 p, %1, %2; selp.u32 %0, %1, %0, p;}" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) & 7]));
            } else if (KIND == 7) { // 2 LOP3 + 1 IMAD
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(one));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(two));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[(i + 3) & 7]) : "r"(two), "r"(one));
            } else if (KIND == 8) { // SHF only
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (KIND == 9) { // 1 LOP3 + 2 IMAD
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(one));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(two), "r"(one));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[(i + 3) & 7]) : "r"(two), "r"(one));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + b[i];
    for (int i = 0; i < 4; i++) s += (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND> void run(const char *name, int per_iter, int threads) {
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<KIND><<<148, threads>>>(out, 1, 2, cyc);
    k<KIND><<<148, threads>>>(out, 1, 2, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double warps_per_smsp = threads / 32.0 / 4.0;
    double inst = (double)ITER * per_iter * warps_per_smsp;
    printf("%-28s threads=%4d cycles=%8lld  warp-inst/clk/SMSP=%.3f\n", name, threads, h, inst / h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int threads : {512, 1024}) {
        run<0>("LOP3", 8, threads);
        run<8>("SHF", 8, threads);
        run<1>("IMAD", 8, threads);
        run<2>("IMAD.WIDE", 8, threads);
        run<3>("LOP3+IMAD 1:1", 16, threads);
        run<4>("LOP3+IMAD.WIDE 1:1", 16, threads);
        run<5>("ISETP+@p IMAD", 16, threads);
        run<6>("ISETP+SEL", 16, threads);
        run<7>("2 LOP3 + 1 IMAD", 24, threads);
        run<9>("1 LOP3 + 2 IMAD", 24, threads);
    }
    return 0;
}
