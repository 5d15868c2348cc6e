// k_reduce.cu -- see kernels overview in device_common.cuh
#include "device_common.cuh"

namespace sipp {

// ------------------------------------------------------------------------------------------------ K2b + K3
// grid = nprod blocks of 32 threads.  partials layout [count][nprod][96 words].  out: nprod x 96 words, boundary
// format (canonical, arkworks nested order).  final_exp = 0 writes the raw product in device format instead.
__global__ void __launch_bounds__(32) k_reduce_fe(const uint32_t* __restrict__ partials, int count, int nprod, uint32_t* __restrict__ out,
                                                  int final_exp, int ark_norm) {
    __shared__ Fq12 sh[32];
    const int tid = threadIdx.x, prod = blockIdx.x;
    Fq12 acc = fq12_one();
    bool first = true;
    for (int i = tid; i < count; i += 32) {
        Fq12 v = load_fq12(partials + ((size_t)i * nprod + prod) * 96);
        if (first) { acc = v; first = false; }
        else acc = fq12_mul(acc, v);
    }
    sh[tid] = acc;
    block_product_fq12(sh, tid, 32);
    if (tid == 0) {
        if (final_exp) {
            Fq12 r = final_exponentiation(sh[0], ark_norm != 0);
            fq12_encode(out + prod * 96, r);
        } else {
            store_fq12(out + prod * 96, sh[0]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ verifier GT fold
// out = zl^x * z * zr^xinv (generic square-and-multiply: inputs need not be in the cyclotomic subgroup).
// in: 3 x 96 words boundary format (zl, z, zr); 2 threads, one power each.
static __device__ __noinline__ Fq12 fq12_pow256(const Fq12& a, const uint32_t* k) {
    Fq12 acc = fq12_one();
    int top = 255;
    while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1u)) top--;
    for (int i = top; i >= 0; i--) {
        acc = fq12_sqr(acc);
        if ((k[i >> 5] >> (i & 31)) & 1u) acc = fq12_mul(acc, a);
    }
    return acc;
}
__global__ void __launch_bounds__(32) k_gt_fold(const uint32_t* __restrict__ in, Scalar256 x, Scalar256 xinv, uint32_t* __restrict__ out) {
    __shared__ Fq12 sh[2];
    const int tid = threadIdx.x;
    if (tid < 2) {
        Fq12 base = fq12_decode(in + (tid == 0 ? 0 : 192));
        sh[tid] = fq12_pow256(base, tid == 0 ? x.w : xinv.w);
    }
    __syncthreads();
    if (tid == 0) {
        Fq12 z = fq12_decode(in + 96);
        Fq12 r = fq12_mul(fq12_mul(sh[0], z), sh[1]);
        fq12_encode(out, r);
    }
}


int launch_reduce_fe(const uint32_t* partials, int count, int nprod, uint32_t* out, int final_exp, int ark_norm, cudaStream_t s) {
    k_reduce_fe<<<nprod, 32, 0, s>>>(partials, count, nprod, out, final_exp, ark_norm);
    return (int)cudaGetLastError();
}
int launch_gt_fold(const uint32_t* in, const Scalar256& x, const Scalar256& xinv, uint32_t* out, cudaStream_t s) {
    k_gt_fold<<<1, 32, 0, s>>>(in, x, xinv, out);
    return (int)cudaGetLastError();
}

}  // namespace sipp
