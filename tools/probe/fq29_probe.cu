// fq29_probe.cu -- microbenchmark of the radix-2^29 field (new fq.cuh) : throughput and single-warp latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o fq29_probe fq29_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include FQ_HEADER
using namespace sipp;

template <int MODE>
__global__ void __launch_bounds__(256) k_chain(int32_t* out, int iters, int seed) {
    Fq a = fq_one(), b = fq_r2();
    a.l[0] ^= threadIdx.x; b.l[0] ^= (blockIdx.x + seed) & 0xffff;
    Fq c = b, d = a; c.l[1] ^= threadIdx.x * 3; d.l[2] ^= threadIdx.x * 5;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { a = fq_mul(a, b); c = fq_mul(c, d); }
        else if (MODE == 1) { a = fq_sqr(a); c = fq_sqr(c); }
        else if (MODE == 2) { Fq x[6] = {a, b, c, d, a, c}, y[6] = {b, c, d, a, d, b}; a = fq_dot<6>(x, y); Fq x2[6] = {c, a, d, b, c, a}; c = fq_dot<6>(x2, y); }
        else if (MODE == 3) { a = fq_mul(a, b); }   // single dependent chain (latency)
        else if (MODE == 4) { a = fq_mul(fq_add(a, c), fq_sub(b, d)); c = fq_mul(fq_sub(c, a), fq_add(d, b)); }
    }
    int32_t r = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) r ^= a.l[i] ^ c.l[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
double run(int blocks, int threads, int iters, double ops_per_iter, const char* name) {
    int32_t* out;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_chain<MODE><<<blocks, threads>>>(out, iters / 10, 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k_chain<MODE><<<blocks, threads>>>(out, iters, rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = ops_per_iter * iters * (double)blocks * threads;
    printf("%-28s blocks %5d x %3d  %8.3f ms  %8.2f Gop/s  %8.1f ns/op/thread\n", name, blocks, threads, best, ops / best / 1e6, best * 1e6 / (ops_per_iter * iters));
    cudaFree(out);
    return ops / best / 1e6;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sm = p.multiProcessorCount;
    printf("%s, %d SMs, %d MHz\n", p.name, sm, p.clockRate / 1000);
    run<0>(sm * 8, 256, 2000, 2, "fq_mul x2 chains (sat)");
    run<1>(sm * 8, 256, 2000, 2, "fq_sqr x2 chains (sat)");
    run<2>(sm * 8, 256, 500, 12, "fq_dot<6> (Fq products, sat)");
    run<4>(sm * 8, 256, 2000, 2, "fq_mul(add,sub) x2 (sat)");
    run<0>(sm * 4, 128, 2000, 2, "fq_mul x2, 1 warp/sched");
    run<3>(sm * 4, 128, 2000, 1, "fq_mul x1, 1 warp/sched");
    run<3>(1, 32, 2000, 1, "fq_mul x1, single warp");
    run<1>(1, 32, 2000, 2, "fq_sqr x2, single warp");
    run<2>(1, 32, 500, 12, "fq_dot<6>, single warp");
    return 0;
}
