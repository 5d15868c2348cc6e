// curve.cuh -- BN254 G1 (over Fq) and G2 (twist, over Fq2): Jacobian arithmetic and the fold step.
//
// Replaces the ark-ec 0.4 short-Weierstrass group law behind
//   new_A[i] = (a1 + a2.mul(x)).into()       /root/reference/src/prover_native.rs:60-64
//   new_B[i] = (b1 + b2.mul(inv_x)).into()   /root/reference/src/prover_native.rs:65-69
// Results are canonical affine points, so any correct algorithm is bit-identical to the reference's
// MSB-first double-and-add.  The identity is encoded as x = y = 0 (ark's `infinity` has x=y=0).
#pragma once
#include "tower.cuh"

namespace sipp {

// field-generic helpers (overloads pick Fq or Fq2)
SIPP_HD Fq f_add(const Fq& a, const Fq& b) { return fq_add(a, b); }
SIPP_HD Fq f_sub(const Fq& a, const Fq& b) { return fq_sub(a, b); }
SIPP_HD Fq f_mul(const Fq& a, const Fq& b) { return fq_mul(a, b); }
SIPP_HD Fq f_sqr(const Fq& a) { return fq_sqr(a); }
SIPP_HD Fq f_dbl(const Fq& a) { return fq_dbl(a); }
SIPP_HD Fq f_neg(const Fq& a) { return fq_neg(a); }
SIPP_HD Fq f_inv(const Fq& a) { return fq_inv(a); }
SIPP_HD bool f_is_zero(const Fq& a) { return fq_is_zero(a); }
SIPP_HD void f_set_one(Fq& a) { a = fq_one(); }
SIPP_HD void f_set_zero(Fq& a) { a = fq_zero(); }
SIPP_HD Fq2 f_add(const Fq2& a, const Fq2& b) { return fq2_add(a, b); }
SIPP_HD Fq2 f_sub(const Fq2& a, const Fq2& b) { return fq2_sub(a, b); }
SIPP_HD Fq2 f_mul(const Fq2& a, const Fq2& b) { return fq2_mul_inl(a, b); }
SIPP_HD Fq2 f_sqr(const Fq2& a) { return fq2_sqr_inl(a); }
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

}  // namespace sipp
