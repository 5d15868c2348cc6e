// k_bls.cu -- the input producers of the reference's BLS-aggregation demo as kernels (SURVEY 8f row 3;
// /root/reference/src/bin/bls_aggregation.rs:95-117), also what generates the large synthetic inputs of the benchmarks:
//
//   public_keys[i] = (G1Affine::generator() * sk_i).into()              :96-99    k_fixed_base_mul<Fq>  (keygen)
//   signatures[i]  = (m_i * sk_i).into()                                :105-109  k_g2_mul_var          (signing)
//   aggregated     = fold(G2Projective::zero(), |acc, s| acc + s).into() :110-113 k_g2_sum_blocks + k_g2_sum_final
//
// (`map_to_g2_without_cofactor_mul(u).mul_by_cofactor()`, :100-104, lives in the un-vendored starky-bn254 crate: no spec
// offline, not built.)  Results are canonical affine points, so any correct algorithm is bit-identical to arkworks'.
//
// Fixed base: a window table T[w][d] = [d 2^(8w)] P (32 windows x 255 affine points, built once per base by k_window_table)
// turns a 254-bit scalar multiplication into at most 32 mixed additions and no doubling -- one thread per scalar, table reads
// served from L2 (512 KB for G1, 1 MB for G2).  Variable base: one thread per (point, scalar), MSB-first double-and-add.
// Sum: every thread adds a strided slice into a Jacobian accumulator, shared-memory tree per block, one 192-byte partial per
// block, a last block adds the partials and normalises.
#define SIPP_CURVE_FQ2_CALLS 1
#define SIPP_FQ_CALLS 1
#include "coop.cuh"
#include "device_common.cuh"

namespace sipp {

template <class F> struct PointIO;
template <> struct PointIO<Fq> {
    static constexpr int WORDS = 16;
    static __device__ __forceinline__ G1A load(const uint32_t* p, size_t i) { return load_g1(p, i); }
    static __device__ __forceinline__ void store(uint32_t* p, size_t i, const G1A& v) { store_g1(p, i, v); }
};
template <> struct PointIO<Fq2> {
    static constexpr int WORDS = 32;
    static __device__ __forceinline__ G2A load(const uint32_t* p, size_t i) { return load_g2(p, i); }
    static __device__ __forceinline__ void store(uint32_t* p, size_t i, const G2A& v) { store_g2(p, i, v); }
};

// T[w][d - 1] = [d 2^(8w)] base, d = 1..255, w = 0..31 (Montgomery affine).  One thread per entry: 8w doublings of the base,
// then an 8-bit double-and-add, then one inversion -- 8,160 short chains, a one-off per base.
template <class F>
__global__ void __launch_bounds__(64) k_window_table(const uint32_t* __restrict__ base, uint32_t* __restrict__ table) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 32 * 255) return;
    const int w = t / 255, d = t % 255 + 1;
    const Affine<F> b = PointIO<F>::load(base, 0);
    Jac<F> bw;
    bw.x = b.x; bw.y = b.y; f_set_one(bw.z);
    if (affine_is_identity(b)) bw = jac_identity<F>();
    for (int i = 0; i < 8 * w; i++) bw = jac_dbl(bw);
    Jac<F> acc = jac_identity<F>();
    for (int i = 7; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((d >> i) & 1) acc = jac_add(acc, bw);
    }
    PointIO<F>::store(table, (size_t)t, jac_to_affine(acc));
}

// out[i] = [k_i] base through the window table; scalars: count x 8 words (canonical little-endian integers < 2^256)
template <class F>
__global__ void __launch_bounds__(128) k_fixed_base_mul(const uint32_t* __restrict__ table, const uint32_t* __restrict__ scalars, size_t count,
                                                        uint32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t k[8];
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = scalars[8 * i + j];
    Jac<F> acc = jac_identity<F>();
#pragma unroll 1
    for (int w = 0; w < 32; w++) {
        const uint32_t d = (k[w >> 2] >> (8 * (w & 3))) & 255u;
        if (d) acc = jac_add_affine(acc, PointIO<F>::load(table, (size_t)(w * 255 + (int)d - 1)));
    }
    PointIO<F>::store(out, i, jac_to_affine(acc));
}

// out[i] = [k_i] P_i (variable base): MSB-first double-and-add (what ark's `Affine * Fr` does), one thread per element
__global__ void __launch_bounds__(64) k_g2_mul_var(const uint32_t* __restrict__ points, const uint32_t* __restrict__ scalars, size_t count,
                                                   uint32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t k[8];
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = scalars[8 * i + j];
    store_g2(out, i, jac_to_affine(jac_scalar_mul(load_g2(points, i), k)));
}

// ---- sum of G2 points ----
#define SIPP_SUM_THREADS 128
__device__ __forceinline__ void sum_store(uint32_t* p, const Jac<Fq2>& v) {
    store_fq2_words(p, v.x); store_fq2_words(p + 16, v.y); store_fq2_words(p + 32, v.z);
}
__device__ __forceinline__ Jac<Fq2> sum_load(const uint32_t* p) {
    Jac<Fq2> v;
    v.x = load_fq2_words(p); v.y = load_fq2_words(p + 16); v.z = load_fq2_words(p + 32);
    return v;
}
__device__ __noinline__ void block_sum_jac(uint32_t* sh, Jac<Fq2>& acc, int tid) {
    for (int s = SIPP_SUM_THREADS >> 1; s > 0; s >>= 1) {
        __syncthreads();
        if (tid >= s && tid < 2 * s) sum_store(sh + 48 * tid, acc);
        __syncthreads();
        if (tid < s) acc = jac_add(acc, sum_load(sh + 48 * (tid + s)));
    }
}
// partials[b] (48 words, Jacobian) = sum of the points b * T + t, stride gridDim * T
__global__ void __launch_bounds__(SIPP_SUM_THREADS) k_g2_sum_blocks(const uint32_t* __restrict__ points, size_t count, uint32_t* __restrict__ partials) {
    __shared__ __align__(16) uint32_t sh[SIPP_SUM_THREADS * 48];
    const int tid = threadIdx.x;
    Jac<Fq2> acc = jac_identity<Fq2>();
    for (size_t i = (size_t)blockIdx.x * SIPP_SUM_THREADS + tid; i < count; i += (size_t)gridDim.x * SIPP_SUM_THREADS)
        acc = jac_add_affine(acc, load_g2(points, i));
    block_sum_jac(sh, acc, tid);
    if (tid == 0) sum_store(partials + 48 * blockIdx.x, acc);
}
// out (32 words, Montgomery affine) = sum of `nparts` Jacobian partials; one block
__global__ void __launch_bounds__(SIPP_SUM_THREADS) k_g2_sum_final(const uint32_t* __restrict__ partials, int nparts, uint32_t* __restrict__ out) {
    __shared__ __align__(16) uint32_t sh[SIPP_SUM_THREADS * 48];
    const int tid = threadIdx.x;
    Jac<Fq2> acc = jac_identity<Fq2>();
    for (int i = tid; i < nparts; i += SIPP_SUM_THREADS) acc = jac_add(acc, sum_load(partials + 48 * i));
    block_sum_jac(sh, acc, tid);
    if (tid == 0) store_g2(out, 0, jac_to_affine(acc));
}

// ---- seeded scalars (the documented SplitMix64 stream of oracle_seeded_scalars: a_0, b_0, a_1, b_1, ...) ----
__device__ __forceinline__ uint64_t bls_splitmix64_at(uint64_t seed, uint64_t step) {
    uint64_t z = seed + step * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// sa[i] = scalar 2i, sb[i] = scalar 2i + 1 of the stream: 4 words -> 256-bit little-endian -> mod r (0 -> 1)
__global__ void k_seeded_scalars(uint64_t seed, size_t n, uint32_t* __restrict__ sa, uint32_t* __restrict__ sb) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    const uint32_t RL[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    uint32_t k[8];
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint64_t v = bls_splitmix64_at(seed, 4 * t + w + 1);
        k[2 * w] = (uint32_t)v; k[2 * w + 1] = (uint32_t)(v >> 32);
    }
    for (int it = 0; it < 6; it++) {  // v < 2^256 < 6r
        uint32_t d[8];
        if (sub8(d, k, RL) == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) k[i] = d[i];
        }
    }
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) any |= k[i];
    if (!any) k[0] = 1;
    uint32_t* dst = (t & 1) ? sb + 8 * (t >> 1) : sa + 8 * (t >> 1);
#pragma unroll
    for (int i = 0; i < 8; i++) dst[i] = k[i];
}
// the two generators in Montgomery form: gens[0..16) = G1 generator, gens[16..48) = G2 generator
__global__ void k_generators(uint32_t* __restrict__ gens) {
    if (blockIdx.x || threadIdx.x) return;
    store_g1(gens, 0, G1A SIPP_G1_GEN_INIT);
    store_g2(gens + 16, 0, G2A SIPP_G2_GEN_INIT);
}

size_t window_table_bytes(int group) { return (size_t)32 * 255 * (group == 1 ? 64 : 128); }
int launch_window_table(int group, const uint32_t* base, uint32_t* table, cudaStream_t s) {
    const unsigned blocks = (32 * 255 + 63) / 64;
    if (group == 1) k_window_table<Fq><<<blocks, 64, 0, s>>>(base, table);
    else k_window_table<Fq2><<<blocks, 64, 0, s>>>(base, table);
    return (int)cudaGetLastError();
}
int launch_fixed_base_mul(int group, const uint32_t* table, const uint32_t* scalars, size_t count, uint32_t* out, cudaStream_t s) {
    const unsigned blocks = (unsigned)((count + 127) / 128);
    if (group == 1) k_fixed_base_mul<Fq><<<blocks, 128, 0, s>>>(table, scalars, count, out);
    else k_fixed_base_mul<Fq2><<<blocks, 128, 0, s>>>(table, scalars, count, out);
    return (int)cudaGetLastError();
}
int launch_g2_mul_var(const uint32_t* points, const uint32_t* scalars, size_t count, uint32_t* out, cudaStream_t s) {
    k_g2_mul_var<<<(unsigned)((count + 63) / 64), 64, 0, s>>>(points, scalars, count, out);
    return (int)cudaGetLastError();
}
int g2_sum_blocks(size_t count, int sm_count) {
    size_t b = (count + SIPP_SUM_THREADS * 4 - 1) / (SIPP_SUM_THREADS * 4);  // >= 4 points per thread before the tree pays
    if (b < 1) b = 1;
    if (b > (size_t)sm_count * 4) b = (size_t)sm_count * 4;
    return (int)b;
}
int launch_g2_sum(const uint32_t* points, size_t count, uint32_t* partials, int blocks, uint32_t* out, cudaStream_t s) {
    k_g2_sum_blocks<<<blocks, SIPP_SUM_THREADS, 0, s>>>(points, count, partials);
    k_g2_sum_final<<<1, SIPP_SUM_THREADS, 0, s>>>(partials, blocks, out);
    return (int)cudaGetLastError();
}
int launch_seeded_scalars(uint64_t seed, size_t n, uint32_t* sa, uint32_t* sb, cudaStream_t s) {
    k_seeded_scalars<<<(unsigned)((2 * n + 255) / 256), 256, 0, s>>>(seed, n, sa, sb);
    return (int)cudaGetLastError();
}
int launch_generators(uint32_t* gens, cudaStream_t s) {
    k_generators<<<1, 32, 0, s>>>(gens);
    return (int)cudaGetLastError();
}

}  // namespace sipp
