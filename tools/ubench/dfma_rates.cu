// FP64 FMA issue rate on sm_100a and whether it overlaps the integer pipes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NCH 8
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double seed, int iters) {
    double x[NCH], b = seed, c = 1.0 / 3; uint64_t w[NCH]; uint32_t lo[NCH], y[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { x[i] = threadIdx.x + i * seed; w[i] = threadIdx.x + i; lo[i] = i; y[i] = threadIdx.x * 3 + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(b), "d"(c));
                if (MODE == 1) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(b), "d"(c));
                                 asm volatile("mad.wide.u32 %0, %1, 0x2affffac, %0;" : "+l"(w[i]) : "r"(y[i])); }
                if (MODE == 2) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(b), "d"(c));
                                 asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(y[i])); }
                if (MODE == 3) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(b), "d"(c));
                                 asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(lo[i]) : "r"(y[i])); }
                if (MODE == 4) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(b), "d"(c));
                if (MODE == 5) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(b));
            }
        }
#pragma unroll
        for (int i = 0; i < NCH; i++) y[i] ^= lo[i] ^ (uint32_t)w[i];
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) r += x[i] + (double)w[i] + lo[i];
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
    double* out; cudaMalloc(&out, 8 * 256 * sms * 8);
    const int iters = 512;
    const char* names[6] = {"DFMA", "DFMA + WIDE.imm", "DFMA + IADD3", "DFMA + LOP3", "DFMA.rz", "DADD"};
#define RUN(M) { float ms = time_ms([&] { k<M><<<sms * 4, 256>>>(out, 1.0000001, iters); }); double grp = (double)sms * 4 * 256 * iters * 16 * NCH; \
    printf("%-20s %8.3f ms  %6.2f T groups/s = %5.1f groups/clk/SM\n", names[M], ms, grp / ms / 1e9, grp / (ms * 1e-3) / sms / 1.965e9); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    return 0;
}
