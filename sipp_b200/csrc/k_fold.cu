// k_fold.cu -- see kernels overview in device_common.cuh
#include "device_common.cuh"

namespace sipp {

// ------------------------------------------------------------------------------------------------ K4
__global__ void __launch_bounds__(64) k_fold_g1(uint32_t* __restrict__ A, size_t h, Scalar256 k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h) return;
    G1A p1 = load_g1(A, i), p2 = load_g1(A, i + h);
    G1A r = jac_to_affine(fold_point_jac(p1, p2, k.w));
    store_g1(A, i, r);
}
__global__ void __launch_bounds__(64) k_fold_g2(uint32_t* __restrict__ B, size_t h, Scalar256 k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h) return;
    G2A p1 = load_g2(B, i), p2 = load_g2(B, i + h);
    G2A r = jac_to_affine(fold_point_jac(p1, p2, k.w));
    store_g2(B, i, r);
}

// ------------------------------------------------------------------------------------------------ inputs
__device__ __forceinline__ uint64_t splitmix64_at(uint64_t seed, uint64_t step) {  // value of the step-th output (1-based)
    uint64_t z = seed + step * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ void seeded_scalar(uint32_t* k, uint64_t seed, uint64_t index) {
    const uint32_t RL[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
#pragma unroll
    for (int w = 0; w < 4; w++) {
        uint64_t v = splitmix64_at(seed, 4 * index + w + 1);
        k[2 * w] = (uint32_t)v; k[2 * w + 1] = (uint32_t)(v >> 32);
    }
    for (int it = 0; it < 6; it++) {  // v < 2^256 < 6r
        uint32_t t[8];
        if (sub8(t, k, RL) == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) k[i] = t[i];
        }
    }
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) any |= k[i];
    if (!any) k[0] = 1;
}
// thread i < n makes A_i, thread n + i makes B_i; outputs in boundary format
__global__ void __launch_bounds__(64) k_seeded_inputs(uint64_t seed, size_t n, uint32_t* __restrict__ dA, uint32_t* __restrict__ dB) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    uint32_t k[8];
    if (t < n) {
        seeded_scalar(k, seed, 2 * t);
        G1A g = G1A SIPP_G1_GEN_INIT;
        G1A r = jac_to_affine(jac_scalar_mul(g, k));
        g1_encode(dA + 16 * t, r);
    } else {
        size_t i = t - n;
        seeded_scalar(k, seed, 2 * i + 1);
        G2A g = G2A SIPP_G2_GEN_INIT;
        G2A r = jac_to_affine(jac_scalar_mul(g, k));
        g2_encode(dB + 32 * i, r);
    }
}


int launch_fold(uint32_t* A, uint32_t* B, size_t h, const Scalar256& x, const Scalar256& xinv, cudaStream_t s) {
    k_fold_g1<<<(unsigned)((h + 63) / 64), 64, 0, s>>>(A, h, x);
    k_fold_g2<<<(unsigned)((h + 63) / 64), 64, 0, s>>>(B, h, xinv);
    return (int)cudaGetLastError();
}
int launch_seeded_inputs(uint64_t seed, size_t n, uint32_t* dA, uint32_t* dB, cudaStream_t s) {
    k_seeded_inputs<<<(unsigned)((2 * n + 63) / 64), 64, 0, s>>>(seed, n, dA, dB);
    return (int)cudaGetLastError();
}

}  // namespace sipp
