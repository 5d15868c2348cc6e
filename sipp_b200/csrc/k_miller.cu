// k_miller.cu -- see kernels overview in device_common.cuh
#include "device_common.cuh"

namespace sipp {

// ------------------------------------------------------------------------------------------------ K1 + K2a
// grid = (ceil(m / 64), nprod).  Product `blockIdx.y` pairs A[a_off[y] + j] with B[b_off[y] + j], j < m.


__global__ void __launch_bounds__(SIPP_MILLER_BLOCK) k_miller_block(const uint32_t* __restrict__ A, const uint32_t* __restrict__ B, MillerJob job,
                                                                  uint32_t* __restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Fq12* sh = reinterpret_cast<Fq12*>(smem_raw);
    const int tid = threadIdx.x;
    const int prod = blockIdx.y;
    size_t j = (size_t)blockIdx.x * SIPP_MILLER_BLOCK + tid;
    Fq12 f;
    if (j < job.m) {
        G1A p = load_g1(A, job.a_off[prod] + j);
        G2A q = load_g2(B, job.b_off[prod] + j);
        f = miller_loop(p, q);
    } else {
        f = fq12_one();
    }
    sh[tid] = f;
    block_product_fq12(sh, tid, SIPP_MILLER_BLOCK);
    if (tid == 0) store_fq12(partials + ((size_t)blockIdx.x * gridDim.y + prod) * 96, sh[0]);
}


int launch_miller_block(const uint32_t* A, const uint32_t* B, const MillerJob& job, int nprod, uint32_t* partials, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_miller_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SIPP_MILLER_BLOCK * sizeof(Fq12)));
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    size_t blocks = (job.m + SIPP_MILLER_BLOCK - 1) / SIPP_MILLER_BLOCK;
    dim3 grid((unsigned)blocks, (unsigned)nprod);
    k_miller_block<<<grid, SIPP_MILLER_BLOCK, SIPP_MILLER_BLOCK * sizeof(Fq12), s>>>(A, B, job, partials);
    return (int)cudaGetLastError();
}

}  // namespace sipp
