// fold_plan.h -- the per-round recoding of the challenge that the host hands to the fold kernels (glv.cc builds it,
// k_fold.cu consumes it as a kernel parameter).
#pragma once
#include <stdint.h>

namespace sipp {

#define SIPP_FOLD_MASK_WORDS 5  // 160 NAF digits: G1 sub-scalars are < 2^128, G2 sub-scalars < 2^66

struct FoldSubScalar {
    uint64_t mag[3];
    int neg;
};
// one endomorphism component: the scalar is sum_i (plus_i - minus_i) 2^i, applied to (-1)^neg * endo^j(P)
struct FoldComp {
    uint32_t plus[SIPP_FOLD_MASK_WORDS];
    uint32_t minus[SIPP_FOLD_MASK_WORDS];
    int neg;
};
struct FoldPlan {
    FoldComp g1[2];  // x      = k0 + k1 L1           (phi)
    FoldComp g2[4];  // x^-1   = k0 + k1 L + k2 L^2 + k3 L^3   (psi)
    int g1_bits, g2_bits;  // NAF length (max over the components)
};

// Pairing-matrix tail (k_mat.cu): the same four-dimensional recoding applied to BOTH x and x^-1, used as exponents in GT, where the
// p-power Frobenius acts as [L] exactly as psi does on G2 (p = 6x^2 mod r).  c[0..3]: x, c[4..7]: x^-1.
struct GtPlan {
    FoldComp c[8];
    int bits;
};

int fold_decompose_g1(const uint64_t k[4], FoldSubScalar out[2]);
int fold_decompose_g2(const uint64_t k[4], FoldSubScalar out[4]);
int fold_plan_build(const uint8_t x[32], const uint8_t x_inv[32], FoldPlan* plan);
int gt_plan_build(const uint8_t x[32], const uint8_t x_inv[32], GtPlan* plan);

}  // namespace sipp
