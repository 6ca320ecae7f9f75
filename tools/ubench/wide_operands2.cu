// IMAD.WIDE.U32 rate by operand source, written in C so that ptxas keeps the fused form.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NCH 12
__device__ __forceinline__ uint32_t launder(uint32_t x) { asm("" : "+r"(x)); return x; }
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
                if (MODE == 0) w[i] += (uint64_t)a[0] * b[0];                    // both shared (hoistable -> adds)
                if (MODE == 1) w[i] += (uint64_t)a[i] * b[u];                    // column-stationary b
                if (MODE == 2) w[i] += (uint64_t)a[i] * b[(i + u) % NCH];        // all distinct
                if (MODE == 3) w[i] += (uint64_t)a[i] * 0x2affffacu;             // immediate
                if (MODE == 4) w[(i + u) % NCH] += (uint64_t)a[i] * b[u];        // product-scanning like: a[i]*b[u] -> column i+u
            }
        }
#pragma unroll
        for (int i = 0; i < NCH; i++) { a[i] = launder(a[i] ^ (uint32_t)w[i]); b[i] = launder(b[i] + (uint32_t)(w[(i + 1) % NCH] >> 32)); }
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
    const int iters = 2048; const char* names[5] = {"both shared", "b shared per 12", "all distinct", "immediate", "product-scan"};
#define RUN(M) { float ms = time_ms([&] { k<M><<<sms * 4, 256>>>(out, 12345, iters); }); double ops = (double)sms * 4 * 256 * iters * 8 * NCH; \
    printf("%-20s %8.3f ms  %6.2f T/s  (%.1f per clk per SM at 1.965 GHz)\n", names[M], ms, ops / ms / 1e9, ops / (ms * 1e-3) / sms / 1.965e9); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
    return 0;
}
