// tower.cuh -- Fq2 / Fq6 / Fq12 over fq.cuh.
//
// Fq2 = Fq[u]/(u^2+1); Fq12 is held in the *w-power basis*  f = sum_i g[i] w^i, g[i] in Fq2, w^6 = xi = 9+u.
// arkworks' nested tower (Fq6 = Fq2[v]/(v^3-xi), Fq12 = Fq6[w]/(w^2-v)) is the same field with
//   c0 = (g0, g2, g4), c1 = (g1, g3, g5)
// so (de)serialisation to ark order is the index map [0,2,4,1,3,5], and plonky2-bn254's `MyFq12.coeffs`
// (transcript_native.rs:33) is [g_i.c0 for i] ++ [g_i.c1 for i]  (SURVEY A.2).
// The w-basis makes the sparse line product (w^0, w^1, w^3), the Frobenius map, Granger-Scott squaring
// (pairs (g0,g3), (g1,g4), (g2,g5)) and the lane-cooperative kernels index-uniform.
#pragma once
#include "constants.cuh"
#include "fq.cuh"

#ifndef SIPP_FQ2_CALL
// Fq2 multiply / square are real calls in the generic (struct) tower to keep code size and ptxas time sane;
// the throughput kernels use their own inlined paths.
#define SIPP_FQ2_CALL SIPP_HD_NOINLINE
#endif

namespace sipp {

struct Fq2 {
    Fq c0, c1;
};
struct Fq6 {
    Fq2 c[3];
};
struct Fq12 {
    Fq2 g[6];
};

// ------------------------------------------------------------------ constants
struct Fq2Raw { uint32_t c0[8]; uint32_t c1[8]; };
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t c_gamma[3][6][2][8] = SIPP_GAMMA_INIT;
#endif
static const uint32_t h_gamma[3][6][2][8] = SIPP_GAMMA_INIT;

SIPP_HD Fq fq_load_const(const uint32_t* p) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = p[i];
    return r;
}
SIPP_HD Fq2 frob_gamma(int k, int i) {  // k in 1..3
#if defined(__CUDA_ARCH__)
    return Fq2{fq_load_const(c_gamma[k - 1][i][0]), fq_load_const(c_gamma[k - 1][i][1])};
#else
    return Fq2{fq_load_const(h_gamma[k - 1][i][0]), fq_load_const(h_gamma[k - 1][i][1])};
#endif
}
SIPP_HD Fq2 fq2_b_twist() { return Fq2 SIPP_B_TWIST_INIT; }
SIPP_HD Fq fq_two_inv() { return Fq SIPP_TWO_INV_INIT; }

// ------------------------------------------------------------------ Fq2
SIPP_HD Fq2 fq2_zero() { return Fq2{fq_zero(), fq_zero()}; }
SIPP_HD Fq2 fq2_one() { return Fq2{fq_one(), fq_zero()}; }
SIPP_HD Fq2 fq2_add(const Fq2& a, const Fq2& b) { return Fq2{fq_add(a.c0, b.c0), fq_add(a.c1, b.c1)}; }
SIPP_HD Fq2 fq2_sub(const Fq2& a, const Fq2& b) { return Fq2{fq_sub(a.c0, b.c0), fq_sub(a.c1, b.c1)}; }
SIPP_HD Fq2 fq2_dbl(const Fq2& a) { return Fq2{fq_dbl(a.c0), fq_dbl(a.c1)}; }
SIPP_HD Fq2 fq2_neg(const Fq2& a) { return Fq2{fq_neg(a.c0), fq_neg(a.c1)}; }
SIPP_HD Fq2 fq2_conj(const Fq2& a) { return Fq2{a.c0, fq_neg(a.c1)}; }
SIPP_HD bool fq2_is_zero(const Fq2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }
SIPP_HD bool fq2_eq(const Fq2& a, const Fq2& b) { return fq_eq(a.c0, b.c0) && fq_eq(a.c1, b.c1); }
SIPP_HD Fq2 fq2_scale(const Fq2& a, const Fq& k) { return Fq2{fq_mul(a.c0, k), fq_mul(a.c1, k)}; }

SIPP_HD Fq2 fq2_mul_inl(const Fq2& a, const Fq2& b) {  // Karatsuba: 3 Fq products
    Fq v0 = fq_mul(a.c0, b.c0);
    Fq v1 = fq_mul(a.c1, b.c1);
    Fq s = fq_mul(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1));
    return Fq2{fq_sub(v0, v1), fq_sub(fq_sub(s, v0), v1)};
}
SIPP_HD Fq2 fq2_sqr_inl(const Fq2& a) {  // (a0+a1)(a0-a1), 2 a0 a1
    Fq m = fq_mul(a.c0, a.c1);
    Fq r0 = fq_mul(fq_add(a.c0, a.c1), fq_sub(a.c0, a.c1));
    return Fq2{r0, fq_dbl(m)};
}
#if defined(SIPP_FQ_CALLS)
// smallest code: one out-of-line Fq multiplier shared by every product of the translation unit (the scalar-multiplication loops of
// k_fold.cu stalled on instruction fetch even with Fq2 products as calls: ncu no_instruction 3.9 per issued instruction)
SIPP_HD_NOINLINE Fq fq_mul_call(const Fq& a, const Fq& b) { return fq_mul(a, b); }
SIPP_FQ2_CALL Fq2 fq2_mul(const Fq2& a, const Fq2& b) {
    Fq v0 = fq_mul_call(a.c0, b.c0);
    Fq v1 = fq_mul_call(a.c1, b.c1);
    Fq s = fq_mul_call(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1));
    return Fq2{fq_sub(v0, v1), fq_sub(fq_sub(s, v0), v1)};
}
SIPP_FQ2_CALL Fq2 fq2_sqr(const Fq2& a) {
    Fq m = fq_mul_call(a.c0, a.c1);
    Fq r0 = fq_mul_call(fq_add(a.c0, a.c1), fq_sub(a.c0, a.c1));
    return Fq2{r0, fq_dbl(m)};
}
#else
SIPP_FQ2_CALL Fq2 fq2_mul(const Fq2& a, const Fq2& b) { return fq2_mul_inl(a, b); }
SIPP_FQ2_CALL Fq2 fq2_sqr(const Fq2& a) { return fq2_sqr_inl(a); }
#endif

SIPP_HD Fq2 fq2_mul_xi(const Fq2& a) {  // (9+u)(a0 + a1 u) = (9 a0 - a1) + (9 a1 + a0) u
    Fq2 t = fq2_dbl(fq2_dbl(fq2_dbl(a)));
    t = fq2_add(t, a);
    return Fq2{fq_sub(t.c0, a.c1), fq_add(t.c1, a.c0)};
}
SIPP_HD_NOINLINE Fq2 fq2_inv(const Fq2& a) {
    Fq n = fq_inv(fq_add(fq_sqr(a.c0), fq_sqr(a.c1)));
    return Fq2{fq_mul(a.c0, n), fq_neg(fq_mul(a.c1, n))};
}

// ------------------------------------------------------------------ Fq6 = Fq2[v]/(v^3 - xi)
SIPP_HD Fq6 fq6_add(const Fq6& a, const Fq6& b) { return Fq6{{fq2_add(a.c[0], b.c[0]), fq2_add(a.c[1], b.c[1]), fq2_add(a.c[2], b.c[2])}}; }
SIPP_HD Fq6 fq6_sub(const Fq6& a, const Fq6& b) { return Fq6{{fq2_sub(a.c[0], b.c[0]), fq2_sub(a.c[1], b.c[1]), fq2_sub(a.c[2], b.c[2])}}; }
SIPP_HD Fq6 fq6_neg(const Fq6& a) { return Fq6{{fq2_neg(a.c[0]), fq2_neg(a.c[1]), fq2_neg(a.c[2])}}; }
SIPP_HD Fq6 fq6_mul_v(const Fq6& a) { return Fq6{{fq2_mul_xi(a.c[2]), a.c[0], a.c[1]}}; }
SIPP_HD_NOINLINE Fq6 fq6_mul(const Fq6& a, const Fq6& b) {  // 6 Fq2 products
    Fq2 v0 = fq2_mul(a.c[0], b.c[0]), v1 = fq2_mul(a.c[1], b.c[1]), v2 = fq2_mul(a.c[2], b.c[2]);
    Fq2 x = fq2_mul(fq2_add(a.c[1], a.c[2]), fq2_add(b.c[1], b.c[2]));
    Fq2 r0 = fq2_add(fq2_mul_xi(fq2_sub(fq2_sub(x, v1), v2)), v0);
    x = fq2_mul(fq2_add(a.c[0], a.c[1]), fq2_add(b.c[0], b.c[1]));
    Fq2 r1 = fq2_add(fq2_sub(fq2_sub(x, v0), v1), fq2_mul_xi(v2));
    x = fq2_mul(fq2_add(a.c[0], a.c[2]), fq2_add(b.c[0], b.c[2]));
    Fq2 r2 = fq2_add(fq2_sub(fq2_sub(x, v0), v2), v1);
    return Fq6{{r0, r1, r2}};
}
SIPP_HD_NOINLINE Fq6 fq6_inv(const Fq6& a) {
    Fq2 t0 = fq2_sub(fq2_sqr(a.c[0]), fq2_mul_xi(fq2_mul(a.c[1], a.c[2])));
    Fq2 t1 = fq2_sub(fq2_mul_xi(fq2_sqr(a.c[2])), fq2_mul(a.c[0], a.c[1]));
    Fq2 t2 = fq2_sub(fq2_sqr(a.c[1]), fq2_mul(a.c[0], a.c[2]));
    Fq2 d = fq2_add(fq2_mul(a.c[0], t0), fq2_mul_xi(fq2_add(fq2_mul(a.c[2], t1), fq2_mul(a.c[1], t2))));
    d = fq2_inv(d);
    return Fq6{{fq2_mul(t0, d), fq2_mul(t1, d), fq2_mul(t2, d)}};
}

// ------------------------------------------------------------------ Fq12 (w-basis)
SIPP_HD Fq12 fq12_one() {
    Fq12 r;
    r.g[0] = fq2_one();
#pragma unroll
    for (int i = 1; i < 6; i++) r.g[i] = fq2_zero();
    return r;
}
SIPP_HD Fq6 fq12_even(const Fq12& a) { return Fq6{{a.g[0], a.g[2], a.g[4]}}; }
SIPP_HD Fq6 fq12_odd(const Fq12& a) { return Fq6{{a.g[1], a.g[3], a.g[5]}}; }
SIPP_HD Fq12 fq12_join(const Fq6& e, const Fq6& o) { return Fq12{{e.c[0], o.c[0], e.c[1], o.c[1], e.c[2], o.c[2]}}; }

SIPP_HD_NOINLINE Fq12 fq12_mul(const Fq12& a, const Fq12& b) {  // 18 Fq2 products
    Fq6 a0 = fq12_even(a), a1 = fq12_odd(a), b0 = fq12_even(b), b1 = fq12_odd(b);
    Fq6 v0 = fq6_mul(a0, b0), v1 = fq6_mul(a1, b1);
    Fq6 x = fq6_mul(fq6_add(a0, a1), fq6_add(b0, b1));
    return fq12_join(fq6_add(v0, fq6_mul_v(v1)), fq6_sub(fq6_sub(x, v0), v1));
}
SIPP_HD_NOINLINE Fq12 fq12_sqr(const Fq12& a) {  // complex squaring: 12 Fq2 products
    Fq6 a0 = fq12_even(a), a1 = fq12_odd(a);
    Fq6 ab = fq6_mul(a0, a1);
    Fq6 u = fq6_mul(fq6_add(a0, a1), fq6_add(a0, fq6_mul_v(a1)));
    return fq12_join(fq6_sub(fq6_sub(u, ab), fq6_mul_v(ab)), fq6_add(ab, ab));
}
SIPP_HD Fq12 fq12_conj(const Fq12& a) {  // w -> -w  (the p^6 Frobenius)
    return Fq12{{a.g[0], fq2_neg(a.g[1]), a.g[2], fq2_neg(a.g[3]), a.g[4], fq2_neg(a.g[5])}};
}
SIPP_HD_NOINLINE Fq12 fq12_inv(const Fq12& a) {
    Fq6 a0 = fq12_even(a), a1 = fq12_odd(a);
    Fq6 t = fq6_inv(fq6_sub(fq6_mul(a0, a0), fq6_mul_v(fq6_mul(a1, a1))));
    return fq12_join(fq6_mul(a0, t), fq6_neg(fq6_mul(a1, t)));
}
SIPP_HD_NOINLINE Fq12 fq12_frob(const Fq12& a, int k) {  // a^(p^k), k in 1..3
    Fq12 r;
    r.g[0] = (k & 1) ? fq2_conj(a.g[0]) : a.g[0];
    for (int i = 1; i < 6; i++) {
        Fq2 g = (k & 1) ? fq2_conj(a.g[i]) : a.g[i];
        r.g[i] = fq2_mul(g, frob_gamma(k, i));
    }
    return r;
}
SIPP_HD bool fq12_eq(const Fq12& a, const Fq12& b) {
    bool e = true;
#pragma unroll
    for (int i = 0; i < 6; i++) e = e && fq2_eq(a.g[i], b.g[i]);
    return e;
}

// f * (l0 + l1 w + l3 w^3): 13 Fq2 products (ark "mul_by_034")
SIPP_HD_NOINLINE Fq12 fq12_mul_sparse(const Fq12& f, const Fq2& l0, const Fq2& l1, const Fq2& l3) {
    Fq6 f0 = fq12_even(f), f1 = fq12_odd(f);
    // a = f0 * l0
    Fq6 a = Fq6{{fq2_mul(f0.c[0], l0), fq2_mul(f0.c[1], l0), fq2_mul(f0.c[2], l0)}};
    // b = f1 * (l1 + l3 v): 5 products
    Fq2 v0 = fq2_mul(f1.c[0], l1), v1 = fq2_mul(f1.c[1], l3);
    Fq2 b0 = fq2_add(v0, fq2_mul_xi(fq2_mul(f1.c[2], l3)));
    Fq2 b1 = fq2_sub(fq2_sub(fq2_mul(fq2_add(f1.c[0], f1.c[1]), fq2_add(l1, l3)), v0), v1);
    Fq2 b2 = fq2_add(v1, fq2_mul(f1.c[2], l1));
    Fq6 b = Fq6{{b0, b1, b2}};
    // e = (f0 + f1) * ((l0 + l1) + l3 v): 5 products
    Fq6 s = fq6_add(f0, f1);
    Fq2 m0 = fq2_add(l0, l1);
    Fq2 w0 = fq2_mul(s.c[0], m0), w1 = fq2_mul(s.c[1], l3);
    Fq2 e0 = fq2_add(w0, fq2_mul_xi(fq2_mul(s.c[2], l3)));
    Fq2 e1 = fq2_sub(fq2_sub(fq2_mul(fq2_add(s.c[0], s.c[1]), fq2_add(m0, l3)), w0), w1);
    Fq2 e2 = fq2_add(w1, fq2_mul(s.c[2], m0));
    Fq6 e = Fq6{{e0, e1, e2}};
    return fq12_join(fq6_add(a, fq6_mul_v(b)), fq6_sub(fq6_sub(e, a), b));
}

// Granger-Scott squaring in the cyclotomic subgroup.  Fq12 = Fq4[w]/(w^3 - s), Fq4 = Fq2[s]/(s^2 - xi), s = w^3:
//   f = A + B w + C w^2 with A = g0 + g3 s, B = g1 + g4 s, C = g2 + g5 s
//   f^2 = (3A^2 - 2 conj(A)) + (3 s C^2 + 2 conj(B)) w + (3B^2 - 2 conj(C)) w^2
SIPP_HD void fq4_sqr(Fq2& t0, Fq2& t1, const Fq2& a, const Fq2& b) {  // (a + b s)^2 = (a^2 + xi b^2) + 2ab s
    Fq2 ab = fq2_mul(a, b);
    Fq2 s = fq2_mul(fq2_add(a, b), fq2_add(a, fq2_mul_xi(b)));
    t0 = fq2_sub(fq2_sub(s, ab), fq2_mul_xi(ab));
    t1 = fq2_dbl(ab);
}
SIPP_HD Fq2 fq2_3t_minus_2z(const Fq2& t, const Fq2& z) { return fq2_add(fq2_dbl(fq2_sub(t, z)), t); }
SIPP_HD Fq2 fq2_3t_plus_2z(const Fq2& t, const Fq2& z) { return fq2_add(fq2_dbl(fq2_add(t, z)), t); }
SIPP_HD_NOINLINE Fq12 fq12_cyc_sqr(const Fq12& f) {
    Fq2 a0, a1, b0, b1, c0, c1;
    fq4_sqr(a0, a1, f.g[0], f.g[3]);
    fq4_sqr(b0, b1, f.g[1], f.g[4]);
    fq4_sqr(c0, c1, f.g[2], f.g[5]);
    Fq12 r;
    r.g[0] = fq2_3t_minus_2z(a0, f.g[0]);
    r.g[3] = fq2_3t_plus_2z(a1, f.g[3]);
    r.g[1] = fq2_3t_plus_2z(fq2_mul_xi(c1), f.g[1]);
    r.g[4] = fq2_3t_minus_2z(c0, f.g[4]);
    r.g[2] = fq2_3t_minus_2z(b0, f.g[2]);
    r.g[5] = fq2_3t_plus_2z(b1, f.g[5]);
    return r;
}
SIPP_HD_NOINLINE Fq12 fq12_cyc_exp_x(const Fq12& a) {  // a^x, x = BN parameter (bit 62 is its top bit)
    Fq12 acc = a;
    const unsigned long long x = SIPP_BN_X;
    for (int b = 61; b >= 0; b--) {
        acc = fq12_cyc_sqr(acc);
        if ((x >> b) & 1ull) acc = fq12_mul(acc, a);
    }
    return acc;
}

}  // namespace sipp
