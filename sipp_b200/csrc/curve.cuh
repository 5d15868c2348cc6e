// curve.cuh -- BN254 G1 (over Fq) and G2 (twist, over Fq2): Jacobian arithmetic and the fold step.
//
// Replaces the ark-ec 0.4 short-Weierstrass group law behind
//   new_A[i] = (a1 + a2.mul(x)).into()       /root/reference/src/prover_native.rs:60-64
//   new_B[i] = (b1 + b2.mul(inv_x)).into()   /root/reference/src/prover_native.rs:65-69
// Results are canonical affine points, so any correct algorithm is bit-identical to the reference's
// MSB-first double-and-add.  The identity is encoded as x = y = 0 (ark's `infinity` has x=y=0).
#pragma once
#include "fold_plan.h"
#include "tower.cuh"

namespace sipp {

// field-generic helpers (overloads pick Fq or Fq2)
SIPP_HD Fq f_add(const Fq& a, const Fq& b) { return fq_add(a, b); }
SIPP_HD Fq f_sub(const Fq& a, const Fq& b) { return fq_sub(a, b); }
#if defined(SIPP_FQ_CALLS)
SIPP_HD Fq f_mul(const Fq& a, const Fq& b) { return fq_mul_call(a, b); }
SIPP_HD Fq f_sqr(const Fq& a) { return fq_mul_call(a, a); }
#else
SIPP_HD Fq f_mul(const Fq& a, const Fq& b) { return fq_mul(a, b); }
SIPP_HD Fq f_sqr(const Fq& a) { return fq_sqr(a); }
#endif
SIPP_HD Fq f_dbl(const Fq& a) { return fq_dbl(a); }
SIPP_HD Fq f_neg(const Fq& a) { return fq_neg(a); }
SIPP_HD Fq f_inv(const Fq& a) { return fq_inv(a); }
SIPP_HD bool f_is_zero(const Fq& a) { return fq_is_zero(a); }
SIPP_HD void f_set_one(Fq& a) { a = fq_one(); }
SIPP_HD void f_set_zero(Fq& a) { a = fq_zero(); }
SIPP_HD Fq2 f_add(const Fq2& a, const Fq2& b) { return fq2_add(a, b); }
SIPP_HD Fq2 f_sub(const Fq2& a, const Fq2& b) { return fq2_sub(a, b); }
// Fq2 products of the group law: inlined by default; a translation unit whose kernels would otherwise outgrow the
// instruction cache (k_lines: 47k SASS instructions when inlined) defines SIPP_CURVE_FQ2_CALLS to make them real calls
#if defined(SIPP_CURVE_FQ2_CALLS)
SIPP_HD Fq2 f_mul(const Fq2& a, const Fq2& b) { return fq2_mul(a, b); }
SIPP_HD Fq2 f_sqr(const Fq2& a) { return fq2_sqr(a); }
#else
SIPP_HD Fq2 f_mul(const Fq2& a, const Fq2& b) { return fq2_mul_inl(a, b); }
SIPP_HD Fq2 f_sqr(const Fq2& a) { return fq2_sqr_inl(a); }
#endif
SIPP_HD Fq2 f_dbl(const Fq2& a) { return fq2_dbl(a); }
SIPP_HD Fq2 f_neg(const Fq2& a) { return fq2_neg(a); }
SIPP_HD Fq2 f_inv(const Fq2& a) { return fq2_inv(a); }
SIPP_HD bool f_is_zero(const Fq2& a) { return fq2_is_zero(a); }
SIPP_HD void f_set_one(Fq2& a) { a = fq2_one(); }
SIPP_HD void f_set_zero(Fq2& a) { a = fq2_zero(); }

template <class F>
struct Affine {
    F x, y;  // identity: x = y = 0
};
template <class F>
struct Jac {
    F x, y, z;  // identity: z = 0
};
typedef Affine<Fq> G1A;
typedef Affine<Fq2> G2A;

template <class F>
SIPP_HD bool affine_is_identity(const Affine<F>& p) { return f_is_zero(p.x) && f_is_zero(p.y); }

template <class F>
SIPP_HD Jac<F> jac_identity() {
    Jac<F> r;
    f_set_one(r.x); f_set_one(r.y); f_set_zero(r.z);
    return r;
}

// 2P, a = 0 (dbl-2009-l): 2M + 5S.  Correct for the identity (z = 0 stays 0).
template <class F>
SIPP_HD Jac<F> jac_dbl(const Jac<F>& p) {
    F A = f_sqr(p.x), B = f_sqr(p.y), C = f_sqr(B);
    F t = f_sub(f_sub(f_sqr(f_add(p.x, B)), A), C);
    F D = f_dbl(t);
    F E = f_add(f_dbl(A), A);
    F Fv = f_sqr(E);
    Jac<F> r;
    r.x = f_sub(Fv, f_dbl(D));
    F C8 = f_dbl(f_dbl(f_dbl(C)));
    r.y = f_sub(f_mul(E, f_sub(D, r.x)), C8);
    r.z = f_dbl(f_mul(p.y, p.z));
    return r;
}

// P + Q with Q affine; handles P = identity, Q = identity, P = Q and P = -Q.
template <class F>
SIPP_HD Jac<F> jac_add_affine(const Jac<F>& p, const Affine<F>& q) {
    if (affine_is_identity(q)) return p;
    if (f_is_zero(p.z)) {
        Jac<F> r;
        r.x = q.x; r.y = q.y; f_set_one(r.z);
        return r;
    }
    F zz = f_sqr(p.z);
    F u2 = f_mul(q.x, zz);
    F s2 = f_mul(f_mul(q.y, p.z), zz);
    F h = f_sub(u2, p.x);
    F rr = f_sub(s2, p.y);
    if (f_is_zero(h)) {
        if (f_is_zero(rr)) return jac_dbl(p);
        return jac_identity<F>();
    }
    F hh = f_sqr(h);
    F hhh = f_mul(hh, h);
    F v = f_mul(p.x, hh);
    Jac<F> r;
    r.x = f_sub(f_sub(f_sub(f_sqr(rr), hhh), v), v);
    r.y = f_sub(f_mul(rr, f_sub(v, r.x)), f_mul(p.y, hhh));
    r.z = f_mul(p.z, h);
    return r;
}

// P + Q, both Jacobian (add-2007-bl without the Z-sum trick: 12M + 4S); handles identities, P = Q and P = -Q.
template <class F>
SIPP_HD Jac<F> jac_add(const Jac<F>& p, const Jac<F>& q) {
    if (f_is_zero(p.z)) return q;
    if (f_is_zero(q.z)) return p;
    F z1z1 = f_sqr(p.z), z2z2 = f_sqr(q.z);
    F u1 = f_mul(p.x, z2z2), u2 = f_mul(q.x, z1z1);
    F s1 = f_mul(f_mul(p.y, q.z), z2z2), s2 = f_mul(f_mul(q.y, p.z), z1z1);
    F h = f_sub(u2, u1);
    F rr = f_sub(s2, s1);
    if (f_is_zero(h)) {
        if (f_is_zero(rr)) return jac_dbl(p);
        return jac_identity<F>();
    }
    F hh = f_sqr(h);
    F hhh = f_mul(hh, h);
    F v = f_mul(u1, hh);
    Jac<F> r;
    r.x = f_sub(f_sub(f_sub(f_sqr(rr), hhh), v), v);
    r.y = f_sub(f_mul(rr, f_sub(v, r.x)), f_mul(s1, hhh));
    r.z = f_mul(f_mul(p.z, q.z), h);
    return r;
}

template <class F>
SIPP_HD Affine<F> jac_to_affine(const Jac<F>& p) {
    Affine<F> r;
    if (f_is_zero(p.z)) {
        f_set_zero(r.x); f_set_zero(r.y);
        return r;
    }
    F zi = f_inv(p.z);
    F zi2 = f_sqr(zi);
    r.x = f_mul(p.x, zi2);
    r.y = f_mul(p.y, f_mul(zi2, zi));
    return r;
}

// [k]P, MSB-first double-and-add over a 256-bit scalar (8 x u32, little endian).  The scalar is shared by the
// whole launch in the fold kernels, so the bit tests are warp-uniform branches.
template <class F>
SIPP_HD Jac<F> jac_scalar_mul(const Affine<F>& p, const uint32_t* k) {
    Jac<F> acc = jac_identity<F>();
    int top = 255;
    while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1u)) top--;
    for (int i = top; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((k[i >> 5] >> (i & 31)) & 1u) acc = jac_add_affine(acc, p);
    }
    return acc;
}

// out = p1 + [k] p2 as a Jacobian point (affine conversion is done by the caller, possibly batched)
template <class F>
SIPP_HD Jac<F> fold_point_jac(const Affine<F>& p1, const Affine<F>& p2, const uint32_t* k) {
    Jac<F> t = jac_scalar_mul(p2, k);
    return jac_add_affine(t, p1);
}

// ---------------------------------------------------------------------------------------------------------
// endomorphisms used by the lane-split fold (fold_plan.h): component j of an element works on endo^j(P)
//   G1: phi(x, y) = (beta x, y) = [L1] P            G2: psi(x, y) = (conj(x) g_{1,2}, conj(y) g_{1,3}) = [6x^2] P
// psi^j(x, y) = (conj^j(x) gamma[j][2], conj^j(y) gamma[j][3])  (gamma[j][i] = xi^(i (p^j - 1) / 6), tower.cuh)
SIPP_HD G1A endo_apply(const G1A& p, int j) {
    if (j == 0) return p;
    G1A r;
    r.x = fq_mul(p.x, Fq SIPP_GLV_BETA_INIT);
    r.y = p.y;
    return r;
}
SIPP_HD G2A endo_apply(const G2A& p, int j) {
    if (j == 0) return p;
    G2A r;
    r.x = f_mul((j & 1) ? fq2_conj(p.x) : p.x, frob_gamma(j, 2));
    r.y = f_mul((j & 1) ? fq2_conj(p.y) : p.y, frob_gamma(j, 3));
    return r;
}

// sum_i (plus_i - minus_i) 2^i * Q over `bits` NAF digits (masks shared by the whole warp: uniform branches)
template <class F>
SIPP_HD Jac<F> jac_scalar_mul_naf(const Affine<F>& q, const uint32_t* plus, const uint32_t* minus, int bits) {
    Jac<F> acc = jac_identity<F>();
    for (int i = bits - 1; i >= 0; i--) {
        acc = jac_dbl(acc);
        const uint32_t m = 1u << (i & 31);
        const bool dp = (plus[i >> 5] & m) != 0, dm = (minus[i >> 5] & m) != 0;
        if (dp || dm) {
            Affine<F> t = q;
            if (dm) t.y = f_neg(q.y);
            acc = jac_add_affine(acc, t);
        }
    }
    return acc;
}

// joint (Straus / Shamir) form of the same sum for the throughput fold (k_fold_straus): the NC sub-scalars of an element share ONE
// accumulator and its doublings; tbl[c] = (+-) endo^c(P), comps[c] the NAF masks of sub-scalar c (fold_plan.h)
template <class F, int NC>
SIPP_HD Jac<F> straus_naf(const Affine<F>* tbl, const FoldComp* comps, int bits) {
    Jac<F> acc = jac_identity<F>();
    for (int i = bits - 1; i >= 0; i--) {
        acc = jac_dbl(acc);
        const uint32_t m = 1u << (i & 31);
        const int w = i >> 5;
        for (int c = 0; c < NC; c++) {
            const bool dp = (comps[c].plus[w] & m) != 0, dm = (comps[c].minus[w] & m) != 0;
            if (dp || dm) {
                Affine<F> t = tbl[c];
                if (dm) t.y = f_neg(t.y);
                acc = jac_add_affine(acc, t);
            }
        }
    }
    return acc;
}

}  // namespace sipp
