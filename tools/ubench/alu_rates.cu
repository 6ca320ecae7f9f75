// ALU-pipe instruction rates on sm_100a and their co-issue with IMAD.WIDE.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NCH 8
__device__ __forceinline__ uint32_t launder(uint32_t x) { asm("" : "+r"(x)); return x; }
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, uint32_t seed, int iters) {
    uint32_t lo[NCH], hi[NCH], y[NCH], z[NCH]; uint64_t w[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { lo[i] = threadIdx.x + i; hi[i] = seed + i; y[i] = seed * (i + 3) + threadIdx.x; z[i] = seed * (i + 11) ^ threadIdx.x; w[i] = lo[i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                if (MODE == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(y[i]));                                             // IADD3
                if (MODE == 1) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(y[i]), "r"(z[i])); // 64-bit add: IADD3 + IADD3.X
                if (MODE == 2) asm volatile("shf.r.wrap.b32 %0, %0, %1, 30;" : "+r"(lo[i]) : "r"(y[i]));                                  // SHF
                if (MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[i]) : "r"(y[i]), "r"(z[i]));                       // LOP3
                if (MODE == 4) { asm volatile("mad.wide.u32 %0, %1, 0x2affffac, %0;" : "+l"(w[i]) : "r"(y[i]));                          // IMAD.WIDE imm-acc
                                 asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(z[i])); }                                         // + 1 IADD3
                if (MODE == 5) { asm volatile("mad.wide.u32 %0, %1, 0x2affffac, %0;" : "+l"(w[i]) : "r"(y[i]));
                                 asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(y[i]), "r"(z[i])); } // + 64-bit add
                if (MODE == 6) { asm volatile("mad.wide.u32 %0, %1, 0x2affffac, %0;" : "+l"(w[i]) : "r"(y[i]));
                                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[i]) : "r"(y[i]), "r"(z[i])); }                   // + LOP3
                if (MODE == 7) { asm volatile("mad.wide.u32 %0, %1, 0x2affffac, %0;" : "+l"(w[i]) : "r"(y[i]));
                                 asm volatile("shf.r.wrap.b32 %0, %0, %1, 30;" : "+r"(lo[i]) : "r"(z[i])); }                              // + SHF
                if (MODE == 8) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(y[i]));                                          // IADD3 with carry out only
            }
        }
#pragma unroll
        for (int i = 0; i < NCH; i++) { y[i] = launder(y[i] ^ lo[i] ^ (uint32_t)w[i]); }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) r ^= w[i] ^ lo[i] ^ ((uint64_t)hi[i] << 32);
    out[blockIdx.x * 256 + threadIdx.x] = r;
}
template <typename F> static float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    uint64_t* out; cudaMalloc(&out, 8 * 256 * sms * 8);
    const int iters = 1024;
    const char* names[9] = {"IADD3", "IADD3+IADD3.X (64b add)", "SHF", "LOP3", "WIDE.imm + IADD3", "WIDE.imm + 64b add", "WIDE.imm + LOP3", "WIDE.imm + SHF", "IADD3 carry-out"};
    const int per[9] = {1, 2, 1, 1, 2, 3, 2, 2, 1};
#define RUN(M) { float ms = time_ms([&] { k<M><<<sms * 4, 256>>>(out, 12345, iters); }); double grp = (double)sms * 4 * 256 * iters * 16 * NCH; \
    printf("%-26s %8.3f ms  %6.2f T groups/s = %5.1f groups/clk/SM  (%d instr per group -> %5.1f instr/clk/SM)\n", names[M], ms, grp / ms / 1e9, grp / (ms * 1e-3) / sms / 1.965e9, per[M], per[M] * grp / (ms * 1e-3) / sms / 1.965e9); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    return 0;
}
