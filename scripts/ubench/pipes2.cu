// Micro-benchmark (dev helper): issue rates of the integer min / add-min / multiply-high instructions the keyed
// window minimum uses, alone and mixed with LOP3 / IMAD, on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes2 pipes2.cu && ./pipes2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
template <int KIND>
__global__ void __launch_bounds__(512, 1) k(uint32_t *out, uint32_t one, uint32_t two, long long *cyc) {
    uint32_t a[8], b[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 7 + i; b[i] = threadIdx.x * 13 + i * 3; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0) { // VIMNMX
                asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (KIND == 1) { // VIMNMX3
                asm volatile("{.reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b[i]), "r"(one));
            } else if (KIND == 2) { // VIADDMNMX (imm)
                asm volatile("{.reg .u32 t; add.u32 t, %0, 5; min.u32 %0, t, %1;}" : "+r"(a[i]) : "r"(b[i]));
            } else if (KIND == 3) { // VIADDMNMX (-reg)
                asm volatile("{.reg .u32 t; sub.u32 t, %1, %2; min.u32 %0, t, %0;}" : "+r"(a[i]) : "r"(b[i]), "r"(b[(i + 1) & 7]));
            } else if (KIND == 4) { // IMAD.HI
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(two));
            } else if (KIND == 5) { // IMAD.HI + LOP3 1:1
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(two));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[(i + 1) & 7]), "r"(one));
            } else if (KIND == 6) { // VIMNMX + IMAD 1:1
                asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(two), "r"(one));
            } else if (KIND == 7) { // SHF.L.W funnel
                asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(one));
            } else if (KIND == 8) { // LOP3 with predicate output + SEL
                asm volatile("{.reg .pred p; .reg .u32 t; and.b32 t, %1, 32; setp.ne.u32 p, t, 0; selp.u32 %0, %0, %1, p;}" : "+r"(a[i]) : "r"(b[i]));
            } else if (KIND == 9) { // predicated STS-free add: @p add
                asm volatile("{.reg .pred p; .reg .u32 t; and.b32 t, %1, 32; setp.ne.u32 p, t, 0; @p add.u32 %0, %0, 256;}" : "+r"(a[i]) : "r"(b[i]));
            } else if (KIND == 10) { // 2 ALU + 1 IMAD (mix like the new loop)
                asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(two));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[(i + 3) & 7]) : "r"(two), "r"(one));
            } else if (KIND == 11) { // IMAD.HI + IMAD 1:1 (both fma pipe?)
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(two));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(two), "r"(one));
            } else if (KIND == 12) { // IADD3
                asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b[i]), "r"(one));
            } else if (KIND == 13) { // 64-bit compare-select (the old window step): setp.lt.u64 + 2 selp
                asm volatile("{.reg .pred p; .reg .u64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; setp.lt.u64 p, x, y; selp.u32 %0, %0, %2, p; selp.u32 %1, %1, %3, p;}"
                             : "+r"(a[i]), "+r"(b[i]) : "r"(a[(i + 1) & 7]), "r"(b[(i + 1) & 7]));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + b[i];
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
    printf("%-34s threads=%4d cycles=%8lld  warp-inst/clk/SMSP=%.3f\n", name, threads, h, inst / h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int threads : {512, 1024}) {
        run<0>("VIMNMX", 8, threads);
        run<1>("VIMNMX3", 8, threads);
        run<2>("VIADDMNMX imm", 8, threads);
        run<3>("VIADDMNMX -reg", 8, threads);
        run<4>("IMAD.HI", 8, threads);
        run<5>("IMAD.HI + LOP3 1:1", 16, threads);
        run<6>("VIMNMX + IMAD 1:1", 16, threads);
        run<7>("SHF.L.W", 8, threads);
        run<8>("LOP3->P + SEL", 16, threads);
        run<9>("LOP3->P + @P IADD", 16, threads);
        run<10>("VIMNMX + LOP3 + IMAD", 24, threads);
        run<11>("IMAD.HI + IMAD 1:1", 16, threads);
        run<12>("IADD3", 8, threads);
        run<13>("ISETP.64 + 2 SEL (3-4 inst)", 32, threads);
    }
    return 0;
}
