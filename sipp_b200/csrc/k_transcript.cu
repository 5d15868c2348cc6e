// k_transcript.cu -- the Fiat-Shamir transcript on the device, one hash chain per SIPP instance (batched proving,
// BASELINE config "4096 independent n=128 instances").
//
// A single proof keeps its transcript on the host (transcript.cc): the chain is strictly sequential and a CPU core runs
// it faster than a GPU thread.  A batch of independent instances has one independent chain per instance, so the chains
// run side by side here and whole instances stay resident in HBM: no challenge ever crosses PCIe.
//
// Mirrors /root/reference/src/transcript_native.rs:14-77 (`Transcript<GoldilocksField>`):
//   append      state <- hash_n_to_hash_no_pad(state || msg)   :25-30  (overwrite-mode sponge, rate 8, width 12)
//   append_g1 / append_g2 / append_fq12                        :32-54  (8 little-endian u32 limbs per Fq, MyFq12 order)
//   get_challenge                                              :56-65  (zero-stripped u32 digits of the digest, mod r)
// followed by what the prover does with the challenge: x^-1 (prover_native.rs:58) and the GLV / GLS recoding the fold
// kernel consumes (glv_core.h).  The permutation is the textbook Poseidon (30 rounds of constants, S-box, MDS): the same
// function as the host's sparse-matrix formulation (tests compare the two).
#include <cuda_runtime.h>
#include <stdint.h>

#define SIPP_GLV_CONST static __device__ const
#include "glv_consts.h"
#include "glv_core.h"
#include "launch.h"
#include "poseidon_rc.h"

namespace sipp {

namespace {

__constant__ uint64_t c_rc[360];
bool g_rc_loaded[64] = {false};

#define GL_EPS 0xFFFFFFFFull
#define GL_P 0xFFFFFFFF00000001ull

__device__ __forceinline__ uint64_t gl_add(uint64_t a, uint64_t b) {  // any u64 representatives
    uint64_t r = a + b;
    if (r < a) {  // wrapped: 2^64 = EPS (mod p)
        uint64_t t = r + GL_EPS;
        r = t < r ? t + GL_EPS : t;
    }
    return r;
}
// x = hi * 2^64 + lo  ->  representative in [0, 2^64)
__device__ __forceinline__ uint64_t gl_reduce128(uint64_t lo, uint64_t hi) {
    uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t = lo - hh;  // 2^96 = -1
    if (lo < hh) t -= GL_EPS;
    uint64_t m = hl * GL_EPS;  // 2^64 = 2^32 - 1
    uint64_t r = t + m;
    if (r < m) r += GL_EPS;
    return r;
}
__device__ __forceinline__ uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128(a * b, __umul64hi(a, b)); }
__device__ __forceinline__ uint64_t gl_pow7(uint64_t x) {
    uint64_t x2 = gl_mul(x, x), x3 = gl_mul(x2, x), x4 = gl_mul(x2, x2);
    return gl_mul(x3, x4);
}
__device__ __forceinline__ uint64_t gl_canon(uint64_t a) { return a >= GL_P ? a - GL_P : a; }

// out[r] = sum_i s[(i + r) mod 12] * CIRC[i] + 8 s[0] [r == 0]; 32-bit halves keep the dot products inside u64
__device__ __forceinline__ void mds_layer(uint64_t* s) {
    const uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { lo[i] = (uint32_t)s[i]; hi[i] = (uint32_t)(s[i] >> 32); }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        uint64_t al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (uint64_t)lo[(i + r) % 12] * C[i];
            ah += (uint64_t)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) { al += (uint64_t)lo[0] * 8; ah += (uint64_t)hi[0] * 8; }
        // value = al + ah 2^32 < 2^75
        uint64_t l = al + (ah << 32);
        uint64_t h = (ah >> 32) + (l < al ? 1 : 0);  // < 2^11
        uint64_t m = h * GL_EPS;
        uint64_t x = l + m;
        if (x < m) x += GL_EPS;
        s[r] = x;
    }
}

__device__ __noinline__ void poseidon_permute(uint64_t* s) {
#pragma unroll 1
    for (int rnd = 0; rnd < 30; rnd++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], c_rc[12 * rnd + i]);
        if (rnd < 4 || rnd >= 26) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
        } else {
            s[0] = gl_pow7(s[0]);
        }
        mds_layer(s);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}

// state <- hash_n_to_hash_no_pad(state || msg); msg = n u32-valued field elements read through `get(i)`
template <class Get>
__device__ __forceinline__ void tr_append(uint64_t st[4], int n, Get get) {
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] = st[i];
#pragma unroll
    for (int i = 4; i < 12; i++) s[i] = 0;
    int first = n < 4 ? n : 4;
    for (int i = 0; i < first; i++) s[4 + i] = get(i);
    poseidon_permute(s);
    for (int off = first; off < n; off += 8) {
        int len = n - off < 8 ? n - off : 8;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < len) s[i] = get(off + i);
        poseidon_permute(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) st[i] = s[i];
}

// f: 96 canonical u32 words in the boundary (ark nested) order; order 0 = MyFq12 w-basis, 1 = nested
__device__ __forceinline__ void tr_append_fq12(uint64_t st[4], const uint32_t* f, int order) {
    tr_append(st, 96, [&](int i) -> uint64_t {
        if (order == 1) return f[i];
        int c = i >> 3, w = i & 7;          // coefficient c of MyFq12.coeffs, limb w
        int g = c < 6 ? c : c - 6;          // Fq2 coefficient of w^g; c >= 6 selects its c1
        int slot = (g & 1) * 3 + (g >> 1);
        return f[16 * slot + (c < 6 ? 0 : 8) + w];
    });
}

}  // namespace

// one thread per instance: register A and B (prover_native.rs:36-39), 8 permutations per pair
__global__ void __launch_bounds__(32) k_tr_absorb_pairs(const uint32_t* __restrict__ bytesA, const uint32_t* __restrict__ bytesB, size_t n, size_t count,
                                                         uint64_t* __restrict__ states) {
    size_t inst = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= count) return;
    uint64_t st[4] = {0, 0, 0, 0};  // Transcript::new  transcript_native.rs:19-23
    const uint32_t* a = bytesA + inst * n * 16;
    const uint32_t* b = bytesB + inst * n * 32;
    for (size_t i = 0; i < n; i++) {
        const uint32_t* pa = a + 16 * i;
        const uint32_t* pb = b + 32 * i;
        tr_append(st, 16, [&](int k) -> uint64_t { return pa[k]; });  // append_g1  :42-46
        tr_append(st, 32, [&](int k) -> uint64_t { return pb[k]; });  // append_g2  :48-54
    }
#pragma unroll
    for (int i = 0; i < 4; i++) states[4 * inst + i] = st[i];
}

// one thread per instance: absorb Z (first round only), Z_L, Z_R, derive the challenge, invert it, recode it.
// proofs: [count][np][96 words] boundary bytes; slots index the Fq12 inside an instance's proof (slot_z < 0: skip).
// flags[0] |= 1 on a zero challenge (x.inverse().unwrap() panics in the reference), |= 2 on a recoding failure.
__global__ void __launch_bounds__(32) k_tr_round(uint64_t* __restrict__ states, const uint32_t* __restrict__ proofs, size_t np, int slot_z, int slot_l, int slot_r,
                                                  int order, size_t count, FoldPlan* __restrict__ plans, uint64_t* __restrict__ challenges, int* __restrict__ flags) {
    size_t inst = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= count) return;
    uint64_t st[4];
#pragma unroll
    for (int i = 0; i < 4; i++) st[i] = states[4 * inst + i];
    const uint32_t* pf = proofs + inst * np * 96;
    if (slot_z >= 0) tr_append_fq12(st, pf + (size_t)slot_z * 96, order);  // prover_native.rs:42-43
    tr_append_fq12(st, pf + (size_t)slot_l * 96, order);                    // :52-53
    tr_append_fq12(st, pf + (size_t)slot_r * 96, order);                    // :54-55
#pragma unroll
    for (int i = 0; i < 4; i++) states[4 * inst + i] = st[i];
    // get_challenge (&self: the state is not advanced)  transcript_native.rs:56-65
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] = st[i];
#pragma unroll
    for (int i = 4; i < 12; i++) s[i] = 0;
    poseidon_permute(s);
    uint32_t digits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int nd = 0;
    for (int k = 0; k < 4; k++) {  // to_u32_digits strips high zero digits (0 -> no digit at all)
        uint64_t d = s[k];
        while (d) {
            digits[nd++] = (uint32_t)d;
            d >>= 32;
        }
    }
    uint64_t v[4];
    for (int j = 0; j < 4; j++) v[j] = (uint64_t)digits[2 * j] | ((uint64_t)digits[2 * j + 1] << 32);
    const uint64_t RM[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    for (int it = 0; it < 6 && glv::fr_geq(v, RM); it++) {  // v < 2^256 < 6r
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            unsigned __int128 d = (unsigned __int128)v[i] - RM[i] - borrow;
            v[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    uint64_t vi[4] = {0, 0, 0, 0};
    int rc = glv::fr_inverse(v, vi);                                        // prover_native.rs:58
    if (challenges) {
        for (int i = 0; i < 4; i++) { challenges[8 * inst + i] = v[i]; challenges[8 * inst + 4 + i] = vi[i]; }
    }
    FoldPlan plan;
    for (int i = 0; i < (int)(sizeof(FoldPlan) / 4); i++) ((uint32_t*)&plan)[i] = 0;
    if (rc) {
        atomicOr(flags, 1);
    } else {
        const glv::Tables t = {&SIPP_GLV_G1_BASIS[0][0][0], &SIPP_GLV_G1_RECIP[0][0], SIPP_GLV_G1_RECIP_SIGN,
                               &SIPP_GLS_G2_BASIS[0][0][0], &SIPP_GLS_G2_RECIP[0][0], SIPP_GLS_G2_RECIP_SIGN};
        if (glv::plan_build(v, vi, t, &plan)) atomicOr(flags, 2);
    }
    plans[inst] = plan;
}

// test hook: `count` independent permutations
__global__ void __launch_bounds__(32) k_test_poseidon(uint64_t* __restrict__ states, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t s[12];
    for (int k = 0; k < 12; k++) s[k] = states[12 * i + k];
    poseidon_permute(s);
    for (int k = 0; k < 12; k++) states[12 * i + k] = s[k];
}

static int load_rc() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_rc_loaded[dev]) return 0;
    cudaError_t e = cudaMemcpyToSymbol(c_rc, SIPP_POSEIDON_RC, sizeof(uint64_t) * 360);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // once per device: the kernels run on a non-blocking stream
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) g_rc_loaded[dev] = true;
    return 0;
}

int launch_tr_absorb_pairs(const uint32_t* bytesA, const uint32_t* bytesB, size_t n, size_t count, uint64_t* states, cudaStream_t s) {
    int e = load_rc();
    if (e) return e;
    k_tr_absorb_pairs<<<(unsigned)((count + 31) / 32), 32, 0, s>>>(bytesA, bytesB, n, count, states);
    return (int)cudaGetLastError();
}
int launch_tr_round(uint64_t* states, const uint32_t* proofs, size_t np, int slot_z, int slot_l, int slot_r, int order, size_t count, FoldPlan* plans,
                    uint64_t* challenges, int* flags, cudaStream_t s) {
    int e = load_rc();
    if (e) return e;
    k_tr_round<<<(unsigned)((count + 31) / 32), 32, 0, s>>>(states, proofs, np, slot_z, slot_l, slot_r, order, count, plans, challenges, flags);
    return (int)cudaGetLastError();
}
int launch_test_poseidon(uint64_t* states, size_t count, cudaStream_t s) {
    int e = load_rc();
    if (e) return e;
    k_test_poseidon<<<(unsigned)((count + 31) / 32), 32, 0, s>>>(states, count);
    return (int)cudaGetLastError();
}

}  // namespace sipp
