// pairing.cuh -- optimal-ate pairing on BN254: line functions, Miller loop, final exponentiation.
//
// Replaces `plonky2_bn254_pairing::pairing::pairing(a, b)` as called at
//   /root/reference/src/prover_native.rs:20   (once per pair inside inner_product)
//   /root/reference/src/verifier_native.rs:80 (the final check)
// (dependency rev fe5c3a8, not vendored; algorithm per SURVEY.md A.1: Miller loop over the 65-digit signed
// form of 6x+2, two Frobenius additions, easy part, Devegili-Scott-Dahab hard part = exponent (p^12-1)/r.)
//
// The GPU computes  FE( prod_i Miller(A_i, B_i) )  instead of  prod_i FE(Miller(A_i, B_i)) -- the same value
// (FE is a homomorphism) with one final exponentiation per product instead of one per pair.
//
// Lines are  l(P) = l0 * yP + l1 * xP * w + l3 * w^3  with T in homogeneous projective coordinates on the
// twist; any Fq2 scaling of a line is killed by the easy part of the final exponentiation.
#pragma once
#include "curve.cuh"

namespace sipp {

struct G2H {
    Fq2 x, y, z;
};

#if defined(SIPP_CURVE_FQ2_CALLS)
#define SIPP_LINE_FN SIPP_HD_NOINLINE  // one copy of each step in instruction-cache-bound kernels
#else
#define SIPP_LINE_FN SIPP_HD
#endif
SIPP_LINE_FN Fq2 fq2_scale_step(const Fq2& a, const Fq& k) { return fq2_scale(a, k); }

// tangent at T, then T <- 2T.  l0 = -2YZ, l1 = 3X^2, l3 = 3b'Z^2 - Y^2.   (6 squarings + 3 products + 2 halvings)
SIPP_LINE_FN void line_double(G2H& t, Fq2& l0, Fq2& l1, Fq2& l3) {
    const Fq half = fq_two_inv();
    const Fq2 bt = fq2_b_twist();
    Fq2 a = fq2_scale(f_mul(t.x, t.y), half);
    Fq2 b = f_sqr(t.y);
    Fq2 c = f_sqr(t.z);
    Fq2 e = f_mul(fq2_add(fq2_dbl(c), c), bt);
    Fq2 f = fq2_add(fq2_dbl(e), e);
    Fq2 g = fq2_scale(fq2_add(b, f), half);
    Fq2 h = fq2_sub(f_sqr(fq2_add(t.y, t.z)), fq2_add(b, c));
    Fq2 j = f_sqr(t.x);
    Fq2 e2 = f_sqr(e);
    l3 = fq2_sub(e, b);
    t.x = f_mul(a, fq2_sub(b, f));
    t.y = fq2_sub(f_sqr(g), fq2_add(fq2_dbl(e2), e2));
    t.z = f_mul(b, h);
    l0 = fq2_neg(h);
    l1 = fq2_add(fq2_dbl(j), j);
}

// chord through T and Q (affine), then T <- T + Q.  l0 = lambda, l1 = -theta, l3 = theta xQ - lambda yQ
SIPP_LINE_FN void line_add(G2H& t, const G2A& q, Fq2& l0, Fq2& l1, Fq2& l3) {
    Fq2 theta = fq2_sub(t.y, f_mul(q.y, t.z));
    Fq2 lambda = fq2_sub(t.x, f_mul(q.x, t.z));
    Fq2 c = f_sqr(theta);
    Fq2 d = f_sqr(lambda);
    Fq2 e = f_mul(lambda, d);
    Fq2 f = f_mul(t.z, c);
    Fq2 g = f_mul(t.x, d);
    Fq2 h = fq2_sub(fq2_add(e, f), fq2_dbl(g));
    t.x = f_mul(lambda, h);
    t.y = fq2_sub(f_mul(theta, fq2_sub(g, h)), f_mul(e, t.y));
    t.z = f_mul(t.z, e);
    l3 = fq2_sub(f_mul(theta, q.x), f_mul(lambda, q.y));
    l0 = lambda;
    l1 = fq2_neg(theta);
}

// pi(Q) = (conj(x) gamma_{1,2}, conj(y) gamma_{1,3})
SIPP_HD G2A g2_frobenius(const G2A& q) {
    G2A r;
    r.x = f_mul(fq2_conj(q.x), frob_gamma(1, 2));
    r.y = f_mul(fq2_conj(q.y), frob_gamma(1, 3));
    return r;
}

// Number of line evaluations per pair: 64 tangents + 25 in-loop chords + 2 Frobenius chords
#define SIPP_LINES_PER_PAIR 91

// Generic line generator: calls emit(step, l0*yP, l1*xP, l3) for every line of the Miller loop of (P, Q), in
// loop order.  `sq(step)` is called before each tangent except the first (where f is still 1).
template <class Sq, class Emit>
SIPP_HD void miller_lines(const G1A& p, const G2A& q, Sq sq, Emit emit) {
    G2H t;
    t.x = q.x; t.y = q.y; t.z = fq2_one();
    G2A nq;
    nq.x = q.x; nq.y = fq2_neg(q.y);
    const unsigned long long plus = SIPP_ATE_PLUS_MASK, minus = SIPP_ATE_MINUS_MASK;
    Fq2 l0, l1, l3;
    int step = 0;
    for (int i = 63; i >= 0; i--) {
        if (i != 63) sq(step);
        line_double(t, l0, l1, l3);
        emit(step++, fq2_scale_step(l0, p.y), fq2_scale_step(l1, p.x), l3);
        const bool dp = (plus >> i) & 1ull, dm = (minus >> i) & 1ull;
        if (dp || dm) {  // one chord site for both signs: the negated point differs only in y
            G2A qs = q;
            if (dm) qs.y = nq.y;
            line_add(t, qs, l0, l1, l3);
            emit(step++, fq2_scale_step(l0, p.y), fq2_scale_step(l1, p.x), l3);
        }
    }
    G2A q1 = g2_frobenius(q);
    G2A q2 = g2_frobenius(q1);
    q2.y = fq2_neg(q2.y);
    line_add(t, q1, l0, l1, l3);
    emit(step++, fq2_scale_step(l0, p.y), fq2_scale_step(l1, p.x), l3);
    line_add(t, q2, l0, l1, l3);
    emit(step++, fq2_scale_step(l0, p.y), fq2_scale_step(l1, p.x), l3);
}

// One full Miller loop in one thread (baseline / small-n path).  Identity inputs contribute 1 (ark convention).
SIPP_HD_NOINLINE Fq12 miller_loop(const G1A& p, const G2A& q) {
    Fq12 f = fq12_one();
    if (affine_is_identity(p) || affine_is_identity(q)) return f;
    miller_lines(
        p, q, [&](int) { f = fq12_sqr(f); },
        [&](int, const Fq2& a, const Fq2& b, const Fq2& c) { f = fq12_mul_sparse(f, a, b, c); });
    return f;
}

// f^((p^12-1)/r): easy part, then Devegili-Scott-Dahab (eprint 2008/490 sec. 5) with cyclotomic squarings.
// `ark_norm` additionally raises to 2x(6x^2+3x+1) (arkworks' Bn254 final exponentiation value; SURVEY A.1 switch).
SIPP_HD_NOINLINE Fq12 final_exponentiation(const Fq12& f, bool ark_norm) {
    Fq12 t = fq12_mul(fq12_conj(f), fq12_inv(f));
    Fq12 m = fq12_mul(fq12_frob(t, 2), t);
    Fq12 mx = fq12_cyc_exp_x(m), mx2 = fq12_cyc_exp_x(mx), mx3 = fq12_cyc_exp_x(mx2);
    Fq12 y0 = fq12_mul(fq12_mul(fq12_frob(m, 1), fq12_frob(m, 2)), fq12_frob(m, 3));
    Fq12 y1 = fq12_conj(m);
    Fq12 y2 = fq12_frob(mx2, 2);
    Fq12 y3 = fq12_conj(fq12_frob(mx, 1));
    Fq12 y4 = fq12_conj(fq12_mul(mx, fq12_frob(mx2, 1)));
    Fq12 y5 = fq12_conj(mx2);
    Fq12 y6 = fq12_conj(fq12_mul(mx3, fq12_frob(mx3, 1)));
    Fq12 t0 = fq12_mul(fq12_mul(fq12_cyc_sqr(y6), y4), y5);
    Fq12 t1 = fq12_mul(fq12_mul(y3, y5), t0);
    t0 = fq12_mul(t0, y2);
    t1 = fq12_cyc_sqr(fq12_mul(fq12_cyc_sqr(t1), t0));
    t0 = fq12_mul(t1, y1);
    t1 = fq12_mul(t1, y0);
    Fq12 out = fq12_mul(fq12_cyc_sqr(t0), t1);
    if (ark_norm) {
        Fq12 a = fq12_cyc_exp_x(out), b = fq12_cyc_exp_x(a), c = fq12_cyc_exp_x(b);
        Fq12 b3 = fq12_mul(fq12_cyc_sqr(b), b);
        Fq12 c6 = fq12_cyc_sqr(fq12_mul(fq12_cyc_sqr(c), c));
        out = fq12_cyc_sqr(fq12_mul(fq12_mul(a, b3), c6));
    }
    return out;
}

}  // namespace sipp
