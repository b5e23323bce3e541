// Micro-benchmark: throughput of the packed min/max flavours used (or considered) by the FAST kernel on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cuda_fp16.h>
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE> __global__ void k(unsigned* out, unsigned seed) {
    unsigned a0 = threadIdx.x + seed, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    __half2 h0 = *(__half2*)&a0, h1 = *(__half2*)&a1, h2 = *(__half2*)&a2, h3 = *(__half2*)&a3;
    __half2 h4 = *(__half2*)&a4, h5 = *(__half2*)&a5, h6 = *(__half2*)&a6, h7 = *(__half2*)&a7;
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
        if (MODE == 0) {          // 8 independent VIMNMX3.S16x2
            a0 = __vimin3_s16x2(a0, a1, a2); a1 = __vimax3_s16x2(a1, a2, a3); a2 = __vimin3_s16x2(a2, a3, a4); a3 = __vimax3_s16x2(a3, a4, a5);
            a4 = __vimin3_s16x2(a4, a5, a6); a5 = __vimax3_s16x2(a5, a6, a7); a6 = __vimin3_s16x2(a6, a7, a0); a7 = __vimax3_s16x2(a7, a0, a1);
        } else if (MODE == 1) {   // 8 independent VIMNMX.S16x2 (2-input)
            a0 = __vmins2(a0, a1); a1 = __vmaxs2(a1, a2); a2 = __vmins2(a2, a3); a3 = __vmaxs2(a3, a4);
            a4 = __vmins2(a4, a5); a5 = __vmaxs2(a5, a6); a6 = __vmins2(a6, a7); a7 = __vmaxs2(a7, a0);
        } else if (MODE == 2) {   // 8 independent HMNMX2
            h0 = __hmin2(h0, h1); h1 = __hmax2(h1, h2); h2 = __hmin2(h2, h3); h3 = __hmax2(h3, h4);
            h4 = __hmin2(h4, h5); h5 = __hmax2(h5, h6); h6 = __hmin2(h6, h7); h7 = __hmax2(h7, h0);
        } else if (MODE == 3) {   // 4 VIMNMX3 + 4 HMNMX2 interleaved
            a0 = __vimin3_s16x2(a0, a1, a2); h0 = __hmin2(h0, h1); a1 = __vimax3_s16x2(a1, a2, a3); h1 = __hmax2(h1, h2);
            a2 = __vimin3_s16x2(a2, a3, a0); h2 = __hmin2(h2, h3); a3 = __vimax3_s16x2(a3, a0, a1); h3 = __hmax2(h3, h0);
        } else if (MODE == 4) {   // 8 PRMT
            a0 = __byte_perm(a0, a1, 0x4321); a1 = __byte_perm(a1, a2, 0x5432); a2 = __byte_perm(a2, a3, 0x6543); a3 = __byte_perm(a3, a4, 0x4240);
            a4 = __byte_perm(a4, a5, 0x4341); a5 = __byte_perm(a5, a6, 0x4321); a6 = __byte_perm(a6, a7, 0x5432); a7 = __byte_perm(a7, a0, 0x6543);
        } else if (MODE == 5) {   // 8 IMAD (fma pipe)
            a0 = a0 * a1 + a2; a1 = a1 * a2 + a3; a2 = a2 * a3 + a4; a3 = a3 * a4 + a5; a4 = a4 * a5 + a6; a5 = a5 * a6 + a7; a6 = a6 * a7 + a0; a7 = a7 * a0 + a1;
        } else if (MODE == 6) {   // 4 VIMNMX3 + 4 IMAD interleaved
            a0 = __vimin3_s16x2(a0, a1, a2); a4 = a4 * a5 + a6; a1 = __vimax3_s16x2(a1, a2, a3); a5 = a5 * a6 + a7;
            a2 = __vimin3_s16x2(a2, a3, a0); a6 = a6 * a7 + a4; a3 = __vimax3_s16x2(a3, a0, a1); a7 = a7 * a4 + a5;
        } else if (MODE == 7) {   // 8 POPC
            a0 += __popc(a1); a1 += __popc(a2); a2 += __popc(a3); a3 += __popc(a4); a4 += __popc(a5); a5 += __popc(a6); a6 += __popc(a7); a7 += __popc(a0);
        } else if (MODE == 8) {   // 8 IDP.4A
            a0 = __dp4a(a1, a2, a0); a1 = __dp4a(a2, a3, a1); a2 = __dp4a(a3, a4, a2); a3 = __dp4a(a4, a5, a3);
            a4 = __dp4a(a5, a6, a4); a5 = __dp4a(a6, a7, a5); a6 = __dp4a(a7, a0, a6); a7 = __dp4a(a0, a1, a7);
        } else if (MODE == 9) {   // 8 VABSDIFF4
            a0 = __vabsdiffu4(a0, a1); a1 = __vabsdiffu4(a1, a2); a2 = __vabsdiffu4(a2, a3); a3 = __vabsdiffu4(a3, a4);
            a4 = __vabsdiffu4(a4, a5); a5 = __vabsdiffu4(a5, a6); a6 = __vabsdiffu4(a6, a7); a7 = __vabsdiffu4(a7, a0);
        }
    }
    unsigned r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ *(unsigned*)&h0 ^ *(unsigned*)&h1 ^ *(unsigned*)&h2 ^ *(unsigned*)&h3 ^
                 *(unsigned*)&h4 ^ *(unsigned*)&h5 ^ *(unsigned*)&h6 ^ *(unsigned*)&h7;
    if (r == 0x12345678) out[0] = r;
}
template <int MODE> void run(const char* name, int ops) {
    unsigned* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(d, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warpInstr = (double)blocks * threads / 32 * ITER * ops;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double perClkSM = warpInstr * 32 / (ms * 1e-3) / (clk * 1e3) / 148;   // thread-ops per clock per SM (at max clock)
    printf("%-34s %8.3f ms  %7.1f thread-ops/clk/SM (at %d MHz)\n", name, ms, perClkSM, clk / 1000);
    cudaFree(d);
}
int main() {
    run<0>("VIMNMX3.S16x2 x8", 8); run<1>("VIMNMX.S16x2 x8", 8); run<2>("HMNMX2 x8", 8); run<3>("4 VIMNMX3 + 4 HMNMX2", 8);
    run<4>("PRMT x8", 8); run<5>("IMAD x8", 8); run<6>("4 VIMNMX3 + 4 IMAD", 8); run<7>("POPC x8 (+IADD)", 8); run<8>("IDP.4A x8", 8);
    run<9>("VABSDIFF4 x8", 8);
    return 0;
}
