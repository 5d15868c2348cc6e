// glv_core.h -- recoding of a round challenge for the fixed-scalar fold kernels, shared by the host (glv.cc: one plan per
// round of a single proof) and the device (k_transcript.cu: one plan per instance per round of a batched proof).
//
// Every element of a round is multiplied by the same scalar (x for G1, x^-1 for G2: /root/reference/src/prover_native.rs:60-69):
//   G2:  k = k0 + k1 L + k2 L^2 + k3 L^3 (mod r), L = 6x^2, |k_j| < 2^66   (psi = twist o Frobenius o untwist acts as [L])
//   G1:  k = k0 + k1 L1 (mod r), |k_j| < 2^128                              (phi(x, y) = (beta x, y) acts as [L1])
// Babai rounding against the LLL-reduced bases of tools/gen_glv.py: c_i = sign_i * floor(k * g_i / 2^320),
// k_j = [j == 0] k - sum_i c_i B[i][j], all in 320-bit two's complement.  Each sub-scalar is then written in
// non-adjacent form as a pair of bit masks (digit +1 / digit -1).
//
// The lattice tables are passed in as pointers so that the same code reads host tables (glv.cc) or __device__ tables.
#pragma once
#include <stdint.h>

#include "fold_plan.h"

#if defined(__CUDACC__)
#define SIPP_GLV_FN __host__ __device__ inline
#else
#define SIPP_GLV_FN inline
#endif

namespace sipp {
namespace glv {

typedef unsigned __int128 u128;
struct I320 {
    uint64_t l[5];
};

SIPP_GLV_FN I320 from_limbs(const uint64_t* p, int n) {
    I320 r;
    for (int i = 0; i < 5; i++) r.l[i] = i < n ? p[i] : 0;
    return r;
}
SIPP_GLV_FN I320 neg(const I320& a) {
    I320 r;
    u128 c = 1;
    for (int i = 0; i < 5; i++) {
        c += (u128)(~a.l[i]);
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    return r;
}
SIPP_GLV_FN I320 sub(const I320& a, const I320& b) {
    I320 r;
    uint64_t borrow = 0;
    for (int i = 0; i < 5; i++) {
        u128 d = (u128)a.l[i] - b.l[i] - borrow;
        r.l[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return r;
}
// low 320 bits of a * b (two's complement product of sign-extended operands)
SIPP_GLV_FN I320 mul_lo(const I320& a, const I320& b) {
    I320 r = {{0, 0, 0, 0, 0}};
    for (int i = 0; i < 5; i++) {
        u128 c = 0;
        for (int j = 0; i + j < 5; j++) {
            c += (u128)a.l[i] * b.l[j] + r.l[i + j];
            r.l[i + j] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
// floor(k * g / 2^320) for unsigned k (4 limbs) and g (5 limbs)
SIPP_GLV_FN I320 mul_shift320(const uint64_t k[4], const uint64_t g[5]) {
    uint64_t t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 5; j++) {
            c += (u128)k[i] * g[j] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        t[i + 5] = (uint64_t)c;
    }
    I320 r;
    for (int i = 0; i < 5; i++) r.l[i] = i + 5 < 9 ? t[i + 5] : 0;
    return r;
}

// basis: D x D x 5 limbs, recip: D x 5 limbs
template <int D>
SIPP_GLV_FN int decompose(const uint64_t k[4], const uint64_t* basis, const uint64_t* recip, const int* sign, FoldSubScalar out[D]) {
    I320 c[D];
    for (int i = 0; i < D; i++) {
        c[i] = mul_shift320(k, recip + 5 * i);
        if (sign[i] < 0) c[i] = neg(c[i]);
    }
    for (int j = 0; j < D; j++) {
        I320 v = j == 0 ? from_limbs(k, 4) : from_limbs(nullptr, 0);
        for (int i = 0; i < D; i++) v = sub(v, mul_lo(c[i], from_limbs(basis + (i * D + j) * 5, 5)));
        out[j].neg = (int)(v.l[4] >> 63);
        if (out[j].neg) v = neg(v);
        if (v.l[3] | v.l[4]) return -1;  // sub-scalar does not fit 192 bits: cannot happen for k < r
        for (int w = 0; w < 3; w++) out[j].mag[w] = v.l[w];
    }
    return 0;
}

// non-adjacent form of a 192-bit magnitude: digit i in {-1, 0, +1}; returns the number of digits
SIPP_GLV_FN int naf(const uint64_t mag[3], uint32_t plus[SIPP_FOLD_MASK_WORDS], uint32_t minus[SIPP_FOLD_MASK_WORDS]) {
    uint64_t m[4] = {mag[0], mag[1], mag[2], 0};
    for (int w = 0; w < SIPP_FOLD_MASK_WORDS; w++) plus[w] = minus[w] = 0;
    int len = 0;
    for (int i = 0; (m[0] | m[1] | m[2] | m[3]) != 0; i++) {
        if (i >= 32 * SIPP_FOLD_MASK_WORDS) return 32 * SIPP_FOLD_MASK_WORDS + 1;  // does not fit the masks (caller rejects)
        if (m[0] & 1) {
            if ((m[0] & 3) == 1) {
                plus[i >> 5] |= 1u << (i & 31);
                m[0] &= ~1ull;  // m -= 1
            } else {
                minus[i >> 5] |= 1u << (i & 31);
                for (int w = 0; w < 4; w++) {  // m += 1
                    if (++m[w] != 0) break;
                }
            }
            len = i + 1;
        }
        for (int w = 0; w < 3; w++) m[w] = (m[w] >> 1) | (m[w + 1] << 63);
        m[3] >>= 1;
    }
    return len;
}

struct Tables {
    const uint64_t* g1_basis;  // [2][2][5]
    const uint64_t* g1_recip;  // [2][5]
    const int* g1_sign;
    const uint64_t* g2_basis;  // [4][4][5]
    const uint64_t* g2_recip;  // [4][5]
    const int* g2_sign;
};

// kx = x (G1 scalar), ki = x^-1 (G2 scalar), both canonical and < r
SIPP_GLV_FN int plan_build(const uint64_t kx[4], const uint64_t ki[4], const Tables& t, FoldPlan* plan) {
    FoldSubScalar s1[2], s2[4];
    if (decompose<2>(kx, t.g1_basis, t.g1_recip, t.g1_sign, s1) || decompose<4>(ki, t.g2_basis, t.g2_recip, t.g2_sign, s2)) return -1;
    plan->g1_bits = plan->g2_bits = 0;
    for (int j = 0; j < 2; j++) {
        int len = naf(s1[j].mag, plan->g1[j].plus, plan->g1[j].minus);
        plan->g1[j].neg = s1[j].neg;
        if (len > plan->g1_bits) plan->g1_bits = len;
    }
    for (int j = 0; j < 4; j++) {
        int len = naf(s2[j].mag, plan->g2[j].plus, plan->g2[j].minus);
        plan->g2[j].neg = s2[j].neg;
        if (len > plan->g2_bits) plan->g2_bits = len;
    }
    if (plan->g1_bits > 32 * SIPP_FOLD_MASK_WORDS || plan->g2_bits > 32 * SIPP_FOLD_MASK_WORDS) return -1;
    return 0;
}

// ---- Fr (the scalar field, r = group order): x^-1 = x^(r-2) with 64-bit Montgomery arithmetic (prover_native.rs:58)
SIPP_GLV_FN bool fr_geq(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; i--)
        if (a[i] != b[i]) return a[i] > b[i];
    return true;
}
SIPP_GLV_FN void fr_mont_mul(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    const uint64_t M[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    const uint64_t INV = 0xc2e1f593efffffffull;
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * INV;
        c = (u128)m * M[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * M[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || fr_geq(t, M)) {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - M[i] - borrow; t[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1; }
    }
    for (int i = 0; i < 4; i++) r[i] = t[i];
}
// returns 0, or -1 when v >= r, -2 when v == 0
SIPP_GLV_FN int fr_inverse(const uint64_t v[4], uint64_t out[4]) {
    const uint64_t M[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    const uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
    if (fr_geq(v, M)) return -1;
    if (!(v[0] | v[1] | v[2] | v[3])) return -2;
    uint64_t base[4], acc[4], one[4] = {1, 0, 0, 0};
    fr_mont_mul(base, v, R2);
    fr_mont_mul(acc, one, R2);
    const uint64_t e[4] = {M[0] - 2, M[1], M[2], M[3]};
    for (int i = 255; i >= 0; i--) {
        fr_mont_mul(acc, acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) fr_mont_mul(acc, acc, base);
    }
    fr_mont_mul(acc, acc, one);
    for (int i = 0; i < 4; i++) out[i] = acc[i];
    return 0;
}

// x^-1 by the binary extended Euclid of Kaliski ("almost inverse") followed by k modular halvings: ~400 + ~400 iterations of a
// few 256-bit shifts / additions instead of the 380 Montgomery products of x^(r-2) -- this is what one lane of the device
// transcript spends its time on every round.  Returns 0, or -1 when v >= r, -2 when v == 0.  Same result as fr_inverse.
SIPP_GLV_FN void u256_shr1(uint64_t* a) {
    for (int i = 0; i < 3; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
    a[3] >>= 1;
}
SIPP_GLV_FN void u256_shl1(uint64_t* a) {
    for (int i = 3; i > 0; i--) a[i] = (a[i] << 1) | (a[i - 1] >> 63);
    a[0] <<= 1;
}
SIPP_GLV_FN void u256_add(uint64_t* a, const uint64_t* b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; a[i] = (uint64_t)c; c >>= 64; }
}
SIPP_GLV_FN void u256_sub(uint64_t* a, const uint64_t* b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - borrow; a[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1; }
}
SIPP_GLV_FN int fr_inverse_binary(const uint64_t v_in[4], uint64_t out[4]) {
    const uint64_t M[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    if (fr_geq(v_in, M)) return -1;
    if (!(v_in[0] | v_in[1] | v_in[2] | v_in[3])) return -2;
    uint64_t u[4], v[4], r[4] = {0, 0, 0, 0}, s[4] = {1, 0, 0, 0};
    for (int i = 0; i < 4; i++) { u[i] = M[i]; v[i] = v_in[i]; }
    int k = 0;
    // invariants: a r = -u 2^k, a s = v 2^k (mod M); r, s < 2 M < 2^255
    while (v[0] | v[1] | v[2] | v[3]) {
        if (!(u[0] & 1)) { u256_shr1(u); u256_shl1(s); }
        else if (!(v[0] & 1)) { u256_shr1(v); u256_shl1(r); }
        else if (!fr_geq(v, u)) { u256_sub(u, v); u256_shr1(u); u256_add(r, s); u256_shl1(s); }   // v < u
        else { u256_sub(v, u); u256_shr1(v); u256_add(s, r); u256_shl1(r); }
        k++;
    }
    if (fr_geq(r, M)) u256_sub(r, M);
    uint64_t x[4] = {M[0], M[1], M[2], M[3]};
    u256_sub(x, r);  // x = a^-1 2^k (mod M), 0 < x < M
    for (int i = 0; i < k; i++) {  // x <- x / 2 (mod M)
        if (x[0] & 1) u256_add(x, M);  // < 2^255: no carry out
        u256_shr1(x);
    }
    for (int i = 0; i < 4; i++) out[i] = x[i];
    return 0;
}

}  // namespace glv
}  // namespace sipp
