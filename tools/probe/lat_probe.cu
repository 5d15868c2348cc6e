// lat_probe.cu -- single-warp latency (cycles per dependent op) of the field primitives the lane engines execute
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sipp_b200/csrc/engine12.cuh"
using namespace sipp;
template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, uint32_t seed, long long* cyc) {
    Fq a = fq_one(), b = fq_r2();
    a.l[0] ^= threadIdx.x; b.l[0] ^= seed; a.l[7] &= 0x0fffffffu; b.l[7] &= 0x0fffffffu;
    Fq v[6] = {a, b, a, b, a, b};
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) a = fq_mul(a, b);
        if (MODE == 1) { Fq x[2] = {a, v[1]}; a = fq_dot<2>(x, v); }
        if (MODE == 2) { Fq x[4] = {a, v[1], v[2], v[3]}; a = fq_dot<4>(x, v); }
        if (MODE == 3) { Fq x[6] = {a, v[1], v[2], v[3], v[4], v[5]}; a = fq_dot<6>(x, v); }
        if (MODE == 4) a = fq_lincomb4(a, v[1], v[2], v[3], 9, -1, 3, -2);
        if (MODE == 5) a = fq_add(a, b);
        if (MODE == 6) a = fq_inv(a);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r ^= a.l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, int threads, int iters) {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    k<MODE><<<1, threads>>>(out, iters, 3, cyc); cudaDeviceSynchronize();
    k<MODE><<<1, threads>>>(out, iters, 5, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-14s threads %3d: %8.1f cycles per op\n", name, threads, (double)h / iters);
}
int main() {
    run<0>("fq_mul", 32, 2000); run<0>("fq_mul", 128, 2000);
    run<1>("fq_dot<2>", 32, 1000); run<2>("fq_dot<4>", 32, 1000); run<3>("fq_dot<6>", 32, 1000); run<3>("fq_dot<6>", 128, 1000);
    run<0>("fq_mul", 512, 2000); run<3>("fq_dot<6>", 256, 1000);
    run<4>("fq_lincomb4", 32, 2000); run<5>("fq_add", 32, 4000); run<6>("fq_inv", 32, 20);
    return 0;
}
