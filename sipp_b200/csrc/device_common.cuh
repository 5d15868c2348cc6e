// kernels.cuh -- CUDA kernels of the SIPP prover hot path (sm_100a).
//
//   k_codec_*        boundary bytes <-> Montgomery limbs in HBM
//   k_miller_block   K1+K2a: one optimal-ate Miller loop per thread, block-level Fq12 tree product in shared memory
//                    -> one 384 B partial per block   (replaces the per-pair `pairing` + serial product of
//                    /root/reference/src/prover_native.rs:17-22; blockIdx.y selects Z_L / Z_R of :48-49)
//   k_reduce_fe      K2b+K3: product of partials (from this GPU's blocks or gathered from all ranks) and ONE final
//                    exponentiation per product, encoded to boundary bytes
//   k_fold_g1/g2     K4: A_i <- A_i + x A_{i+h}, B_i <- B_i + x^-1 B_{i+h} in place (prover_native.rs:60-69)
//   k_gt_fold        verifier_native.rs:59-61
//   k_seeded_inputs  synthetic input generator (SplitMix64 scalars times the generators)
//
// HBM layout: A = n x 16 words, B = n x 32 words (Montgomery limbs, array of structures: one point is one or two
// 64-byte runs read with 128-bit loads); partial products = 96 words (w-basis Fq12, Montgomery).
#pragma once
#include <cuda_runtime.h>

#include "codec.cuh"
#include "pairing.cuh"
#include "launch.h"

namespace sipp {


// ------------------------------------------------------------------------------------------------ loads
__device__ __forceinline__ void load_words16(uint32_t* dst, const uint32_t* src) {  // 64 B, 16-byte aligned
    const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint4 v = __ldg(s + i);
        dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
}
__device__ __forceinline__ void store_words16(uint32_t* dst, const uint32_t* src) {
    uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = make_uint4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
}
__device__ __forceinline__ G1A load_g1(const uint32_t* A, size_t i) {
    uint32_t w[16];
    load_words16(w, A + 16 * i);
    G1A p;
#pragma unroll
    for (int k = 0; k < 8; k++) { p.x.l[k] = w[k]; p.y.l[k] = w[8 + k]; }
    return p;
}
__device__ __forceinline__ void store_g1(uint32_t* A, size_t i, const G1A& p) {
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 8; k++) { w[k] = p.x.l[k]; w[8 + k] = p.y.l[k]; }
    store_words16(A + 16 * i, w);
}
__device__ __forceinline__ G2A load_g2(const uint32_t* B, size_t i) {
    uint32_t w[32];
    load_words16(w, B + 32 * i);
    load_words16(w + 16, B + 32 * i + 16);
    G2A q;
#pragma unroll
    for (int k = 0; k < 8; k++) { q.x.c0.l[k] = w[k]; q.x.c1.l[k] = w[8 + k]; q.y.c0.l[k] = w[16 + k]; q.y.c1.l[k] = w[24 + k]; }
    return q;
}
__device__ __forceinline__ void store_g2(uint32_t* B, size_t i, const G2A& q) {
    uint32_t w[32];
#pragma unroll
    for (int k = 0; k < 8; k++) { w[k] = q.x.c0.l[k]; w[8 + k] = q.x.c1.l[k]; w[16 + k] = q.y.c0.l[k]; w[24 + k] = q.y.c1.l[k]; }
    store_words16(B + 32 * i, w);
    store_words16(B + 32 * i + 16, w + 16);
}
__device__ __forceinline__ void store_fq12(uint32_t* dst, const Fq12& f) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) { dst[16 * i + k] = f.g[i].c0.l[k]; dst[16 * i + 8 + k] = f.g[i].c1.l[k]; }
    }
}
__device__ __forceinline__ Fq12 load_fq12(const uint32_t* src) {
    Fq12 f;
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) { f.g[i].c0.l[k] = src[16 * i + k]; f.g[i].c1.l[k] = src[16 * i + 8 + k]; }
    }
    return f;
}

static __device__ __noinline__ void block_product_fq12(Fq12* sh, int tid, int nthreads) {
    // tree product over sh[0..nthreads), nthreads a power of two; result in sh[0]
    for (int s = nthreads >> 1; s > 0; s >>= 1) {
        __syncthreads();
        if (tid < s) sh[tid] = fq12_mul(sh[tid], sh[tid + s]);
    }
    __syncthreads();
}

}  // namespace sipp
