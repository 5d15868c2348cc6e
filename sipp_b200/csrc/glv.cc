// glv.cc -- host-side recoding of the round challenge for the fixed-scalar fold kernels.
//
// Every element of a round is multiplied by the same scalar (x for G1, x^-1 for G2: /root/reference/src/prover_native.rs:60-69),
// so the scalar is decomposed and recoded ONCE per round on the host and the kernels run a warp-uniform schedule:
//   G2:  k = k0 + k1 L + k2 L^2 + k3 L^3 (mod r), L = 6x^2, |k_j| < 2^66   (psi = twist o Frobenius o untwist acts as [L])
//   G1:  k = k0 + k1 L1 (mod r), |k_j| < 2^128                              (phi(x, y) = (beta x, y) acts as [L1])
// Babai rounding against the LLL-reduced bases of tools/gen_glv.py: c_i = sign_i * floor(k * g_i / 2^320),
// k_j = [j == 0] k - sum_i c_i B[i][j], all in 320-bit two's complement.  Each sub-scalar is then written in
// non-adjacent form as a pair of bit masks (digit +1 / digit -1).
#include <string.h>

#include "fold_plan.h"
#include "glv_consts.h"

namespace sipp {
namespace {

typedef unsigned __int128 u128;
struct I320 {
    uint64_t l[5];
};

I320 from_limbs(const uint64_t* p, int n) {
    I320 r;
    for (int i = 0; i < 5; i++) r.l[i] = i < n ? p[i] : 0;
    return r;
}
I320 neg(const I320& a) {
    I320 r;
    u128 c = 1;
    for (int i = 0; i < 5; i++) {
        c += (u128)(~a.l[i]);
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    return r;
}
I320 sub(const I320& a, const I320& b) {
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
I320 mul_lo(const I320& a, const I320& b) {
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
I320 mul_shift320(const uint64_t k[4], const uint64_t g[5]) {
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

template <int D>
int decompose(const uint64_t k[4], const uint64_t basis[D][D][5], const uint64_t recip[D][5], const int* sign, FoldSubScalar out[D]) {
    I320 c[D];
    for (int i = 0; i < D; i++) {
        c[i] = mul_shift320(k, recip[i]);
        if (sign[i] < 0) c[i] = neg(c[i]);
    }
    for (int j = 0; j < D; j++) {
        I320 v = j == 0 ? from_limbs(k, 4) : from_limbs(nullptr, 0);
        for (int i = 0; i < D; i++) v = sub(v, mul_lo(c[i], from_limbs(basis[i][j], 5)));
        out[j].neg = (int)(v.l[4] >> 63);
        if (out[j].neg) v = neg(v);
        if (v.l[3] | v.l[4]) return -1;  // sub-scalar does not fit 192 bits: cannot happen for k < r
        for (int w = 0; w < 3; w++) out[j].mag[w] = v.l[w];
    }
    return 0;
}

// non-adjacent form of a 192-bit magnitude: digit i in {-1, 0, +1}; returns the number of digits
int naf(const uint64_t mag[3], uint32_t plus[SIPP_FOLD_MASK_WORDS], uint32_t minus[SIPP_FOLD_MASK_WORDS]) {
    uint64_t m[4] = {mag[0], mag[1], mag[2], 0};
    memset(plus, 0, sizeof(uint32_t) * SIPP_FOLD_MASK_WORDS);
    memset(minus, 0, sizeof(uint32_t) * SIPP_FOLD_MASK_WORDS);
    int len = 0;
    for (int i = 0; (m[0] | m[1] | m[2] | m[3]) != 0; i++) {
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

}  // namespace

int fold_decompose_g2(const uint64_t k[4], FoldSubScalar out[4]) { return decompose<4>(k, SIPP_GLS_G2_BASIS, SIPP_GLS_G2_RECIP, SIPP_GLS_G2_RECIP_SIGN, out); }
int fold_decompose_g1(const uint64_t k[4], FoldSubScalar out[2]) { return decompose<2>(k, SIPP_GLV_G1_BASIS, SIPP_GLV_G1_RECIP, SIPP_GLV_G1_RECIP_SIGN, out); }

int fold_plan_build(const uint8_t x[32], const uint8_t x_inv[32], FoldPlan* plan) {
    uint64_t kx[4], ki[4];
    memcpy(kx, x, 32);
    memcpy(ki, x_inv, 32);
    memset(plan, 0, sizeof *plan);
    FoldSubScalar s1[2], s2[4];
    if (fold_decompose_g1(kx, s1) || fold_decompose_g2(ki, s2)) return -1;
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

}  // namespace sipp
