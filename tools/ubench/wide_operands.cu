// IMAD.WIDE.U32 issue rate as a function of where its operands come from.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NCH 12
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, uint32_t seed, int iters) {
    uint64_t w[NCH]; uint32_t a[NCH], b[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { w[i] = threadIdx.x + i; a[i] = seed * (i + 3) + threadIdx.x; b[i] = seed * (i + 7) ^ threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[0]), "r"(b[0]));          // both operands shared
                if (MODE == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[0]));          // one shared
                if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[(i + u) % NCH])); // all distinct
                if (MODE == 3) asm volatile("mad.wide.u32 %0, %1, 0x2affffac, %0;" : "+l"(w[i]) : "r"(a[i]));              // immediate
                if (MODE == 4) asm volatile("mad.wide.u32 %0, %1, 0x2affffac, %0;" : "+l"(w[i]) : "r"(a[0]));              // immediate + shared
                if (MODE == 5) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(b[(i + u) % NCH]));    // no addend
                if (MODE == 6) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(b[(i + u) % NCH]));   // 32-bit, 3 regs
            }
        }
        // operands must change, or ptxas hoists the loop-invariant products out of the loop
#pragma unroll
        for (int i = 0; i < NCH; i++) { a[i] ^= (uint32_t)w[i]; b[i] += (uint32_t)(w[(i + 1) % NCH] >> 32); }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) r ^= w[i] ^ a[i];
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
    const int iters = 2048; const char* names[7] = {"both shared", "one shared", "all distinct", "imm, a distinct", "imm, a shared", "mul.wide distinct", "mad.lo 3 regs"};
#define RUN(M) { float ms = time_ms([&] { k<M><<<sms * 4, 256>>>(out, 12345, iters); }); double ops = (double)sms * 4 * 256 * iters * 8 * NCH; \
    printf("%-20s %8.3f ms  %6.2f T/s  (%.1f per clk per SM at 1.965 GHz)\n", names[M], ms, ops / ms / 1e9, ops / (ms * 1e-3) / sms / 1.965e9); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    return 0;
}
