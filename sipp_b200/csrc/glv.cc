// glv.cc -- host-side recoding of the round challenge for the fixed-scalar fold kernels (algorithm: glv_core.h).
//
// Every element of a round is multiplied by the same scalar (x for G1, x^-1 for G2: /root/reference/src/prover_native.rs:60-69),
// so the scalar is decomposed and recoded ONCE per round on the host and the kernels run a warp-uniform schedule.
#include <string.h>

#include "glv_consts.h"
#include "glv_core.h"

namespace sipp {

static const glv::Tables HOST_TABLES = {&SIPP_GLV_G1_BASIS[0][0][0], &SIPP_GLV_G1_RECIP[0][0], SIPP_GLV_G1_RECIP_SIGN,
                                        &SIPP_GLS_G2_BASIS[0][0][0], &SIPP_GLS_G2_RECIP[0][0], SIPP_GLS_G2_RECIP_SIGN};

int fold_decompose_g2(const uint64_t k[4], FoldSubScalar out[4]) {
    return glv::decompose<4>(k, HOST_TABLES.g2_basis, HOST_TABLES.g2_recip, HOST_TABLES.g2_sign, out);
}
int fold_decompose_g1(const uint64_t k[4], FoldSubScalar out[2]) {
    return glv::decompose<2>(k, HOST_TABLES.g1_basis, HOST_TABLES.g1_recip, HOST_TABLES.g1_sign, out);
}

int fold_plan_build(const uint8_t x[32], const uint8_t x_inv[32], FoldPlan* plan) {
    uint64_t kx[4], ki[4];
    memcpy(kx, x, 32);
    memcpy(ki, x_inv, 32);
    memset(plan, 0, sizeof *plan);
    return glv::plan_build(kx, ki, HOST_TABLES, plan);
}

int gt_plan_build(const uint8_t x[32], const uint8_t x_inv[32], GtPlan* plan) {
    memset(plan, 0, sizeof *plan);
    for (int which = 0; which < 2; which++) {
        uint64_t k[4];
        memcpy(k, which ? x_inv : x, 32);
        FoldSubScalar s[4];
        if (fold_decompose_g2(k, s)) return -1;
        for (int j = 0; j < 4; j++) {
            FoldComp& c = plan->c[4 * which + j];
            const int len = glv::naf(s[j].mag, c.plus, c.minus);
            c.neg = s[j].neg;
            if (len > plan->bits) plan->bits = len;
        }
    }
    return plan->bits > 32 * SIPP_FOLD_MASK_WORDS ? -1 : 0;
}

}  // namespace sipp

// host copy of the inversion the device transcript uses (glv_core.h), for the CPU tests: 0 ok, -1 v >= r, -2 v == 0
extern "C" int sipp_test_fr_inverse_binary(const uint8_t x[32], uint8_t out[32]) {
    uint64_t v[4], o[4] = {0, 0, 0, 0};
    memcpy(v, x, 32);
    int rc = sipp::glv::fr_inverse_binary(v, o);
    memcpy(out, o, 32);
    return rc;
}
