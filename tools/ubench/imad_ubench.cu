// Integer-pipe microbenchmarks for sm_100a: which multiply form is the cheapest
// way to get a 32x32->64 MAC, and what the achievable peak is.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o imad_ubench imad_ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 16

template <int MODE>
__global__ void __launch_bounds__(256) k_imad(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a[CHAINS], b = seed | 1u, c = threadIdx.x + 12345u;
    uint64_t w[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { a[i] = threadIdx.x * 7 + i + seed; w[i] = a[i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int i = 0; i < CHAINS; i++) {
                if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
                if (MODE == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));   // IADD3 (alu pipe)
                if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) r ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// carry chains as used by a 12-limb row of a Montgomery product:
// acc[0..12] += a[0..11] * b   (even/odd split -> mad.lo.cc / madc.hi.cc pairs)
template <int MODE>
__global__ void __launch_bounds__(256) k_row(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a[12], acc[14], b = seed | 1u;
#pragma unroll
    for (int i = 0; i < 12; i++) a[i] = threadIdx.x * 3 + i + seed;
#pragma unroll
    for (int i = 0; i < 14; i++) acc[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (MODE == 0) {
                // even limbs: lo/hi pairs chained with carry  (12 IMAD-class instrs)
                asm volatile(
                    "mad.lo.cc.u32 %0, %12, %24, %0;\n\t"  "madc.hi.cc.u32 %1, %12, %24, %1;\n\t"
                    "madc.lo.cc.u32 %2, %14, %24, %2;\n\t" "madc.hi.cc.u32 %3, %14, %24, %3;\n\t"
                    "madc.lo.cc.u32 %4, %16, %24, %4;\n\t" "madc.hi.cc.u32 %5, %16, %24, %5;\n\t"
                    "madc.lo.cc.u32 %6, %18, %24, %6;\n\t" "madc.hi.cc.u32 %7, %18, %24, %7;\n\t"
                    "madc.lo.cc.u32 %8, %20, %24, %8;\n\t" "madc.hi.cc.u32 %9, %20, %24, %9;\n\t"
                    "madc.lo.cc.u32 %10, %22, %24, %10;\n\t" "madc.hi.u32 %11, %22, %24, %11;\n\t"
                    : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
                      "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11])
                    : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
                      "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(b));
            } else {
                // same work written as 6 mad.wide + 64-bit carry adds
                uint64_t t[6];
#pragma unroll
                for (int i = 0; i < 6; i++) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t[i]) : "r"(a[2 * i]), "r"(b));
                uint32_t lo[6], hi[6];
#pragma unroll
                for (int i = 0; i < 6; i++) { lo[i] = (uint32_t)t[i]; hi[i] = (uint32_t)(t[i] >> 32); }
                asm volatile(
                    "add.cc.u32 %0, %0, %12;\n\t"  "addc.cc.u32 %1, %1, %13;\n\t"
                    "addc.cc.u32 %2, %2, %14;\n\t" "addc.cc.u32 %3, %3, %15;\n\t"
                    "addc.cc.u32 %4, %4, %16;\n\t" "addc.cc.u32 %5, %5, %17;\n\t"
                    "addc.cc.u32 %6, %6, %18;\n\t" "addc.cc.u32 %7, %7, %19;\n\t"
                    "addc.cc.u32 %8, %8, %20;\n\t" "addc.cc.u32 %9, %9, %21;\n\t"
                    "addc.cc.u32 %10, %10, %22;\n\t" "addc.u32 %11, %11, %23;\n\t"
                    : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
                      "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11])
                    : "r"(lo[0]), "r"(hi[0]), "r"(lo[1]), "r"(hi[1]), "r"(lo[2]), "r"(hi[2]),
                      "r"(lo[3]), "r"(hi[3]), "r"(lo[4]), "r"(hi[4]), "r"(lo[5]), "r"(hi[5]));
            }
            b += acc[11];
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) r ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, clk_khz);
    uint32_t* out; cudaMalloc(&out, sizeof(uint32_t) * 256 * sms * 8);
    const int iters = 4096;
    const char* names[5] = {"mad.lo.u32", "mad.hi.u32", "mad.wide.u32", "add.u32", "lop3"};
    for (int wpb = 1; wpb <= 8; wpb *= 2) {       // resident CTAs of 256 threads per SM
        dim3 grid(sms * wpb), block(256);
        double ops = (double)grid.x * 256 * iters * UNROLL * CHAINS;
        double t;
        t = time_ms([&] { k_imad<0><<<grid, block>>>(out, 3, iters); }); printf("ctas/sm %d %-14s %8.3f ms  %7.2f Tops/s  %6.2f ops/clk/SM@%dMHz\n", wpb, names[0], t, ops / t / 1e9, ops / t / 1e3 / sms / clk_khz, clk_khz/1000);
        t = time_ms([&] { k_imad<1><<<grid, block>>>(out, 3, iters); }); printf("ctas/sm %d %-14s %8.3f ms  %7.2f Tops/s  %6.2f ops/clk/SM\n", wpb, names[1], t, ops / t / 1e9, ops / t / 1e3 / sms / clk_khz);
        t = time_ms([&] { k_imad<2><<<grid, block>>>(out, 3, iters); }); printf("ctas/sm %d %-14s %8.3f ms  %7.2f Tops/s  %6.2f ops/clk/SM\n", wpb, names[2], t, ops / t / 1e9, ops / t / 1e3 / sms / clk_khz);
        t = time_ms([&] { k_imad<3><<<grid, block>>>(out, 3, iters); }); printf("ctas/sm %d %-14s %8.3f ms  %7.2f Tops/s  %6.2f ops/clk/SM\n", wpb, names[3], t, ops / t / 1e9, ops / t / 1e3 / sms / clk_khz);
        t = time_ms([&] { k_imad<4><<<grid, block>>>(out, 3, iters); }); printf("ctas/sm %d %-14s %8.3f ms  %7.2f Tops/s  %6.2f ops/clk/SM\n", wpb, names[4], t, ops / t / 1e9, ops / t / 1e3 / sms / clk_khz);
        double rops = (double)grid.x * 256 * iters * 4 * 6;   // 32x32->64 MACs
        t = time_ms([&] { k_row<0><<<grid, block>>>(out, 3, iters); }); printf("ctas/sm %d %-14s %8.3f ms  %7.2f TMAC/s %6.2f MAC/clk/SM\n", wpb, "row lo.cc/hi.cc", t, rops / t / 1e9, rops / t / 1e3 / sms / clk_khz);
        t = time_ms([&] { k_row<1><<<grid, block>>>(out, 3, iters); }); printf("ctas/sm %d %-14s %8.3f ms  %7.2f TMAC/s %6.2f MAC/clk/SM\n", wpb, "row wide+addc", t, rops / t / 1e9, rops / t / 1e3 / sms / clk_khz);
    }
    return 0;
}
