// hostcheck.cpp -- TEST HARNESS ONLY.  Compiles the *device* arithmetic headers (sipp_b200/csrc/*.cuh) for the
// host with g++ so that the kernels' algorithms (tower, Miller loop, final exponentiation, folds, codecs) can be
// checked against the oracle on a machine without a GPU.  The PTX carry chains are replaced by the plain-C++
// emulation in fq.cuh; everything above them is the exact code the GPU runs.  Never linked into the product.
#include <cstring>
#include <vector>

#include "../../sipp_b200/csrc/codec.cuh"
#include "../../sipp_b200/csrc/pairing.cuh"
#include "../../sipp_b200/csrc/fqdot.cuh"
#include "../../sipp_b200/csrc/coop.cuh"

using namespace sipp;

static const uint32_t* W(const uint8_t* p) { return reinterpret_cast<const uint32_t*>(p); }
static uint32_t* W(uint8_t* p) { return reinterpret_cast<uint32_t*>(p); }

extern "C" {

// op: 0 mul (carry-chain algorithm), 1 mul (portable), 2 add, 3 sub, 4 inv, 5 neg
int hc_fq_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t count) {
    for (size_t i = 0; i < count; i++) {
        Fq x = fq_decode(W(a + 32 * i)), y = b ? fq_decode(W(b + 32 * i)) : fq_zero(), r;
        switch (op) {
            case 0: r = fq_mul(x, y); break;
            case 1: r = fq_mul_portable(x, y); break;
            case 2: r = fq_add(x, y); break;
            case 3: r = fq_sub(x, y); break;
            case 4: r = fq_inv(x); break;
            case 5: r = fq_neg(x); break;
            default: return -1;
        }
        fq_encode(W(out + 32 * i), r);
    }
    return 0;
}

// op: 0 mul, 1 sqr, 2 inv, 3..5 frob1..3, 6 conj, 7 cyc_sqr, 8 cyc_exp_x, 9 sparse (b holds l0,l1,l3 in its first 192 bytes)
int hc_fq12_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    Fq12 x = fq12_decode(W(a)), r;
    Fq12 y = b ? fq12_decode(W(b)) : fq12_one();
    switch (op) {
        case 0: r = fq12_mul(x, y); break;
        case 1: r = fq12_sqr(x); break;
        case 2: r = fq12_inv(x); break;
        case 3: case 4: case 5: r = fq12_frob(x, op - 2); break;
        case 6: r = fq12_conj(x); break;
        case 7: r = fq12_cyc_sqr(x); break;
        case 8: r = fq12_cyc_exp_x(x); break;
        case 9: r = fq12_mul_sparse(x, fq2_decode(W(b)), fq2_decode(W(b + 64)), fq2_decode(W(b + 128))); break;
        default: return -1;
    }
    fq12_encode(W(out), r);
    return 0;
}

int hc_pairing(const uint8_t* a, const uint8_t* b, uint8_t* out, int ark_norm) {
    Fq12 f = miller_loop(g1_decode(W(a)), g2_decode(W(b)));
    fq12_encode(W(out), final_exponentiation(f, ark_norm != 0));
    return 0;
}

// product of Miller loops then one final exponentiation, exactly as the GPU pipeline composes it
int hc_inner_product(const uint8_t* A, const uint8_t* B, size_t n, uint8_t* out) {
    Fq12 acc = fq12_one();
    for (size_t i = 0; i < n; i++) acc = fq12_mul(acc, miller_loop(g1_decode(W(A + 64 * i)), g2_decode(W(B + 128 * i))));
    fq12_encode(W(out), final_exponentiation(acc, false));
    return 0;
}

int hc_fold_g1(const uint8_t* p1, const uint8_t* p2, const uint8_t* k, uint8_t* out) {
    G1A r = jac_to_affine(fold_point_jac(g1_decode(W(p1)), g1_decode(W(p2)), W(k)));
    g1_encode(W(out), r);
    return 0;
}
int hc_fold_g2(const uint8_t* p1, const uint8_t* p2, const uint8_t* k, uint8_t* out) {
    G2A r = jac_to_affine(fold_point_jac(g2_decode(W(p1)), g2_decode(W(p2)), W(k)));
    g2_encode(W(out), r);
    return 0;
}

}  // extern "C"

#include "hostcheck_coop.inc"

// lane-split fold (k_fold_split): the per-warp component products and the combine step, emulated sequentially with
// the device code's own functions; the plan comes from the product's host recoder (glv.cc, linked into this harness)
#include "../../sipp_b200/csrc/fold_plan.h"
extern "C" int hc_fold_split_g1(const uint8_t* p1, const uint8_t* p2, const uint8_t* x, const uint8_t* xinv, uint8_t* out) {
    FoldPlan plan;
    if (fold_plan_build(x, xinv, &plan)) return -1;
    G1A a1 = g1_decode(W(p1)), a2 = g1_decode(W(p2));
    Jac<Fq> acc = jac_identity<Fq>();
    for (int j = 0; j < 2; j++) {
        G1A q = endo_apply(a2, j);
        if (plan.g1[j].neg) q.y = f_neg(q.y);
        acc = jac_add(acc, jac_scalar_mul_naf(q, plan.g1[j].plus, plan.g1[j].minus, plan.g1_bits));
    }
    g1_encode(W(out), jac_to_affine(jac_add_affine(acc, a1)));
    return 0;
}
extern "C" int hc_fold_split_g2(const uint8_t* p1, const uint8_t* p2, const uint8_t* x, const uint8_t* xinv, uint8_t* out) {
    FoldPlan plan;
    if (fold_plan_build(x, xinv, &plan)) return -1;
    G2A b1 = g2_decode(W(p1)), b2 = g2_decode(W(p2));
    Jac<Fq2> acc = jac_identity<Fq2>();
    for (int j = 0; j < 4; j++) {
        G2A q = endo_apply(b2, j);
        if (plan.g2[j].neg) q.y = f_neg(q.y);
        acc = jac_add(acc, jac_scalar_mul_naf(q, plan.g2[j].plus, plan.g2[j].minus, plan.g2_bits));
    }
    g2_encode(W(out), jac_to_affine(jac_add_affine(acc, b1)));
    return 0;
}

// shared-doubling fold (k_fold_straus): one accumulator for all components of an element, same plan
extern "C" int hc_fold_straus_g1(const uint8_t* p1, const uint8_t* p2, const uint8_t* x, const uint8_t* xinv, uint8_t* out) {
    FoldPlan plan;
    if (fold_plan_build(x, xinv, &plan)) return -1;
    G1A a1 = g1_decode(W(p1)), a2 = g1_decode(W(p2)), tbl[2];
    for (int c = 0; c < 2; c++) {
        tbl[c] = endo_apply(a2, c);
        if (plan.g1[c].neg) tbl[c].y = f_neg(tbl[c].y);
    }
    Jac<Fq> acc = straus_naf<Fq, 2>(tbl, plan.g1, plan.g1_bits);
    g1_encode(W(out), jac_to_affine(jac_add_affine(acc, a1)));
    return 0;
}
extern "C" int hc_fold_straus_g2(const uint8_t* p1, const uint8_t* p2, const uint8_t* x, const uint8_t* xinv, uint8_t* out) {
    FoldPlan plan;
    if (fold_plan_build(x, xinv, &plan)) return -1;
    G2A b1 = g2_decode(W(p1)), b2 = g2_decode(W(p2)), tbl[4];
    for (int c = 0; c < 4; c++) {
        tbl[c] = endo_apply(b2, c);
        if (plan.g2[c].neg) tbl[c].y = f_neg(tbl[c].y);
    }
    Jac<Fq2> acc = straus_naf<Fq2, 4>(tbl, plan.g2, plan.g2_bits);
    g2_encode(W(out), jac_to_affine(jac_add_affine(acc, b1)));
    return 0;
}
