// How fast can ONE warp alone on its scheduler issue?  (The rounds half of the two-warp SHA-256 is one warp per
// SM sub-partition; ncu shows it issuing every ~2 cycles.)  Cycles per instruction, clock64 around an unrolled
// loop, one CTA of 32 threads:
//   mode 0  8 independent IADD3 chains            (issue rate, ALU pipe)
//   mode 1  1 dependent IADD3 chain               (ALU latency)
//   mode 2  4 independent IADD3 + 4 independent IMAD chains, interleaved   (two pipes)
//   mode 3  1 dependent chain SHF -> LOP3 -> IADD3 (the SHA-256 e chain)
//   mode 4  8 independent IMAD chains             (issue rate, multiply pipe)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(32) k(uint32_t* out, long long* cyc, uint32_t seed, int iters) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (i + 1) + threadIdx.x;
    const uint32_t b = seed | 1u, c = seed ^ 0x55u;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[0]) : "r"(b));
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[4 + i]) : "r"(b), "r"(c));
                }
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    asm volatile("shf.r.wrap.b32 %0, %0, %0, 6;" : "+r"(a[0]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[0]) : "r"(b), "r"(c));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[0]) : "r"(b));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[0]) : "r"(c));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i];
    out[threadIdx.x] = r;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 128); cudaMallocManaged(&cyc, 8);
    const int iters = 4096; const char* names[5] = {"8 independent IADD3", "1 dependent IADD3 chain", "4 IADD3 + 4 IMAD independent", "SHF->LOP3->IADD3->IADD3 chain", "8 independent IMAD"};
#define RUN(M) { k<M><<<1, 32>>>(out, cyc, 12345, iters); cudaDeviceSynchronize(); k<M><<<1, 32>>>(out, cyc, 12345, iters); cudaDeviceSynchronize(); \
    printf("%-34s %.2f cycles per instruction\n", names[M], (double)*cyc / ((double)iters * 16 * 8)); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
    return 0;
}
