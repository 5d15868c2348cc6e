#include <cstdio>
#include <cuda_runtime.h>
#include "u29.cuh"
using namespace u29;
template <int MODE>
__global__ void __launch_bounds__(256) k_chain(uint32_t* out, int iters, int seed) {
    Fq a, b;
#pragma unroll
    for (int i = 0; i < 9; i++) { a.l[i] = (0x1234567u * (i + 1) + threadIdx.x) & U_MASK; b.l[i] = (0x7654321u * (i + 3) + blockIdx.x + seed + threadIdx.x * 7) & U_MASK; }
    Fq c = b, d = a; c.l[1] ^= threadIdx.x * 3; d.l[2] ^= threadIdx.x * 5;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { a = mul(a, b); c = mul(c, d); }
        else if (MODE == 2) { Fq x[6] = {a, b, c, d, a, c}, y[6] = {b, c, d, a, d, b}; a = dot<6>(x, y); Fq x2[6] = {c, a, d, b, c, a}; c = dot<6>(x2, y); }
        else if (MODE == 3) { a = mul(a, b); }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) r ^= a.l[i] ^ c.l[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// raw instruction rates
template <int S>
__global__ void __launch_bounds__(256) k_wide(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint64_t r0 = a, r1 = a + 1, r2 = a + 2, r3 = a + 3, r4 = a + 4, r5 = a + 5, r6 = a + 6, r7 = a + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (S) asm volatile("mad.wide.s32 %0, %8, %9, %0;\n\tmad.wide.s32 %1, %8, %9, %1;\n\tmad.wide.s32 %2, %8, %9, %2;\n\tmad.wide.s32 %3, %8, %9, %3;\n\t"
                         "mad.wide.s32 %4, %8, %9, %4;\n\tmad.wide.s32 %5, %8, %9, %5;\n\tmad.wide.s32 %6, %8, %9, %6;\n\tmad.wide.s32 %7, %8, %9, %7;"
                         : "+l"(r0), "+l"(r1), "+l"(r2), "+l"(r3), "+l"(r4), "+l"(r5), "+l"(r6), "+l"(r7) : "r"(b), "r"(a));
            else asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\tmad.wide.u32 %1, %8, %9, %1;\n\tmad.wide.u32 %2, %8, %9, %2;\n\tmad.wide.u32 %3, %8, %9, %3;\n\t"
                         "mad.wide.u32 %4, %8, %9, %4;\n\tmad.wide.u32 %5, %8, %9, %5;\n\tmad.wide.u32 %6, %8, %9, %6;\n\tmad.wide.u32 %7, %8, %9, %7;"
                         : "+l"(r0), "+l"(r1), "+l"(r2), "+l"(r3), "+l"(r4), "+l"(r5), "+l"(r6), "+l"(r7) : "r"(b), "r"(a));
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
}
template <class K>
void timeit(K launch, double ops, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(0); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); launch(rep + 1); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("%-30s %8.3f ms  %9.2f Gop/s\n", name, best, ops / best / 1e6);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sm = p.multiProcessorCount;
    uint32_t* out; cudaMalloc(&out, (size_t)sm * 8 * 256 * 8);
    const int it = 2000;
    timeit([&](int s) { k_wide<0><<<sm * 8, 256>>>((uint64_t*)out, it, s); }, 64.0 * it * sm * 8 * 256, "mad.wide.u32");
    timeit([&](int s) { k_wide<1><<<sm * 8, 256>>>((uint64_t*)out, it, s); }, 64.0 * it * sm * 8 * 256, "mad.wide.s32");
    timeit([&](int s) { k_chain<0><<<sm * 8, 256>>>(out, it, s); }, 2.0 * it * sm * 8 * 256, "u29 mul x2 (sat)");
    timeit([&](int s) { k_chain<2><<<sm * 8, 256>>>(out, it / 4, s); }, 12.0 * it / 4 * sm * 8 * 256, "u29 dot<6> products (sat)");
    timeit([&](int s) { k_chain<0><<<sm * 4, 128>>>(out, it, s); }, 2.0 * it * sm * 4 * 128, "u29 mul x2, 1 warp/sched");
    timeit([&](int s) { k_chain<3><<<sm * 4, 128>>>(out, it, s); }, 1.0 * it * sm * 4 * 128, "u29 mul x1, 1 warp/sched");
    timeit([&](int s) { k_chain<3><<<1, 32>>>(out, it, s); }, 1.0 * it * 32, "u29 mul x1, single warp");
    timeit([&](int s) { k_chain<2><<<1, 32>>>(out, it / 4, s); }, 12.0 * it / 4 * 32, "u29 dot<6>, single warp");
    return 0;
}
