/*
 * sipp_oracle.c -- CPU oracle (plain C11 + unsigned __int128) for the SIPP native prover hot path.
 * TEST INFRASTRUCTURE ONLY -- see sipp_oracle.h for who may use it and for the "parity unpinned" notice.
 *
 * Restates, function by function:
 *   /root/reference/src/prover_native.rs:15-23   inner_product           -> oracle_inner_product
 *   /root/reference/src/prover_native.rs:26-80   sipp_prove_native       -> oracle_sipp_prove
 *   /root/reference/src/verifier_native.rs:14-85 sipp_verify_native      -> oracle_sipp_verify
 *   /root/reference/src/transcript_native.rs:14-77 Transcript, from_fq_to_f -> oracle_transcript_*
 * and the published algorithms of the un-vendored crates those lines call:
 *   plonky2-bn254-pairing @ fe5c3a8  pairing = final_exp(miller_loop)   (optimal ate, DSD hard part)
 *   plonky2-bn254 @ d616d57          MyFq12 coefficient order (w-power basis)
 *   plonky2 @ 541e127                Poseidon over Goldilocks, hash_n_to_hash_no_pad
 *   ark-bn254 / ark-ec / ark-ff 0.4  tower Fq2/Fq6/Fq12, short-Weierstrass group law, Fr inverse
 *   num-bigint 0.4                   to_u32_digits zero stripping
 *
 * Representation: Fq/Fr = 4 x 64-bit limbs, Montgomery form (R = 2^256), CIOS multiplication.
 * Every derived constant (Montgomery R, R^2, Frobenius coefficients, twist b', Poseidon round
 * constants) is computed at start-up from p, r, x and the generators -- nothing else is hard-coded.
 */
#include "sipp_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fp;           /* Montgomery residue, either modulus */
typedef struct { fp c0, c1; } fq2;
typedef struct { fq2 c0, c1, c2; } fq6;
typedef struct { fq6 c0, c1; } fq12;

typedef struct { uint64_t m[4]; uint64_t inv; fp one; fp r2; } modulus;

static modulus MQ = {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}, 0, {{0}}, {{0}}};
static modulus MR = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}, 0, {{0}}, {{0}}};
static const uint64_t BN_X = 4965661367192848881ULL;
static const int8_t ATE_DIGITS[65] = {0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0, 1, 1,
                                      1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, 1, 1};

/* ------------------------------------------------------------------------------------------- */
/* multi-precision helpers                                                                     */
/* ------------------------------------------------------------------------------------------- */
static int geq4(const uint64_t *a, const uint64_t *b) {
    for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
    return 1;
}
static uint64_t add4(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static uint64_t sub4(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
static void fp_add(fp *r, const fp *a, const fp *b, const modulus *M) {
    uint64_t c = add4(r->l, a->l, b->l);
    if (c || geq4(r->l, M->m)) sub4(r->l, r->l, M->m);
}
static void fp_sub(fp *r, const fp *a, const fp *b, const modulus *M) {
    if (sub4(r->l, a->l, b->l)) add4(r->l, r->l, M->m);
}
static int fp_is_zero(const fp *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static int fp_eq(const fp *a, const fp *b) { return memcmp(a, b, sizeof(fp)) == 0; }
static void fp_neg(fp *r, const fp *a, const modulus *M) {
    if (fp_is_zero(a)) { *r = *a; return; }
    sub4(r->l, M->m, a->l);
}
/* CIOS Montgomery product */
static void fp_mul(fp *r, const fp *a, const fp *b, const modulus *M) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * M->inv;
        c = (u128)m * M->m[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * M->m[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    memcpy(r->l, t, 32);
    if (t[4] || geq4(r->l, M->m)) sub4(r->l, r->l, M->m);
}
static void fp_sqr(fp *r, const fp *a, const modulus *M) { fp_mul(r, a, a, M); }
static void fp_pow(fp *r, const fp *a, const uint64_t *e, int nlimbs, const modulus *M) {
    fp acc = M->one, base = *a;
    for (int i = 0; i < nlimbs; i++)
        for (int b = 0; b < 64; b++) {
            if ((e[i] >> b) & 1) fp_mul(&acc, &acc, &base, M);
            fp_sqr(&base, &base, M);
        }
    *r = acc;
}
static void fp_inv(fp *r, const fp *a, const modulus *M) { /* Fermat; 0 -> 0 */
    uint64_t e[4]; const uint64_t two[4] = {2, 0, 0, 0};
    sub4(e, M->m, two);
    fp_pow(r, a, e, 4, M);
}
static void fp_from_u64x4(fp *r, const uint64_t *v, const modulus *M) { /* v < m required */
    fp t; memcpy(t.l, v, 32);
    fp_mul(r, &t, &M->r2, M);
}
static void fp_to_u64x4(uint64_t *v, const fp *a, const modulus *M) {
    fp one = {{1, 0, 0, 0}}, t;
    fp_mul(&t, a, &one, M);
    memcpy(v, t.l, 32);
}
static void fp_from_small(fp *r, uint64_t v, const modulus *M) { uint64_t t[4] = {v, 0, 0, 0}; fp_from_u64x4(r, t, M); }
static int fp_from_bytes(fp *r, const uint8_t *b, const modulus *M) { /* little endian; returns -1 if >= m */
    uint64_t v[4];
    for (int i = 0; i < 4; i++) { v[i] = 0; for (int j = 7; j >= 0; j--) v[i] = (v[i] << 8) | b[8 * i + j]; }
    if (geq4(v, M->m)) return -1;
    fp_from_u64x4(r, v, M);
    return 0;
}
static void fp_to_bytes(uint8_t *b, const fp *a, const modulus *M) {
    uint64_t v[4]; fp_to_u64x4(v, a, M);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[8 * i + j] = (uint8_t)(v[i] >> (8 * j));
}
static void modulus_init(modulus *M) {
    uint64_t inv = 1; /* Newton: inv = m^-1 mod 2^64 */
    for (int i = 0; i < 6; i++) inv *= 2 - M->m[0] * inv;
    M->inv = (uint64_t)(0 - inv);
    /* R mod m by 256 doublings of 1, R^2 by 256 more */
    fp t = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) {
        uint64_t c = add4(t.l, t.l, t.l);
        if (c || geq4(t.l, M->m)) sub4(t.l, t.l, M->m);
        if (i == 255) M->one = t;
    }
    M->r2 = t;
}

/* ------------------------------------------------------------------------------------------- */
/* Fq wrappers and the tower  Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v) */
/* ------------------------------------------------------------------------------------------- */
#define Q (&MQ)
static void q_add(fp *r, const fp *a, const fp *b) { fp_add(r, a, b, Q); }
static void q_sub(fp *r, const fp *a, const fp *b) { fp_sub(r, a, b, Q); }
static void q_mul(fp *r, const fp *a, const fp *b) { fp_mul(r, a, b, Q); }
static void q_neg(fp *r, const fp *a) { fp_neg(r, a, Q); }

static fq2 XI, B_TWIST, GAMMA[4][6]; /* GAMMA[k][i] = xi^(i (p^k-1)/6), k = 1..3 */
static fp TWO_INV;
static fq12 F12_ONE;
static fp G1GEN[2]; static fq2 G2GEN[2];
static uint64_t POSEIDON_RC[360];

static void f2_add(fq2 *r, const fq2 *a, const fq2 *b) { q_add(&r->c0, &a->c0, &b->c0); q_add(&r->c1, &a->c1, &b->c1); }
static void f2_sub(fq2 *r, const fq2 *a, const fq2 *b) { q_sub(&r->c0, &a->c0, &b->c0); q_sub(&r->c1, &a->c1, &b->c1); }
static void f2_dbl(fq2 *r, const fq2 *a) { f2_add(r, a, a); }
static void f2_neg(fq2 *r, const fq2 *a) { q_neg(&r->c0, &a->c0); q_neg(&r->c1, &a->c1); }
static void f2_conj(fq2 *r, const fq2 *a) { r->c0 = a->c0; q_neg(&r->c1, &a->c1); }
static int f2_is_zero(const fq2 *a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static int f2_eq(const fq2 *a, const fq2 *b) { return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1); }
static void f2_mul(fq2 *r, const fq2 *a, const fq2 *b) { /* Karatsuba, 3 base multiplications */
    fp v0, v1, s, t;
    q_mul(&v0, &a->c0, &b->c0); q_mul(&v1, &a->c1, &b->c1);
    q_add(&s, &a->c0, &a->c1); q_add(&t, &b->c0, &b->c1);
    q_mul(&s, &s, &t);
    q_sub(&s, &s, &v0); q_sub(&r->c1, &s, &v1);
    q_sub(&r->c0, &v0, &v1);
}
static void f2_sqr(fq2 *r, const fq2 *a) { /* (a0+a1)(a0-a1), 2 a0 a1 */
    fp s, d, m;
    q_add(&s, &a->c0, &a->c1); q_sub(&d, &a->c0, &a->c1); q_mul(&m, &a->c0, &a->c1);
    q_mul(&r->c0, &s, &d); q_add(&r->c1, &m, &m);
}
static void f2_scale(fq2 *r, const fq2 *a, const fp *k) { q_mul(&r->c0, &a->c0, k); q_mul(&r->c1, &a->c1, k); }
static void f2_mul_xi(fq2 *r, const fq2 *a) { /* (9+u)(a0 + a1 u) = (9a0 - a1) + (9a1 + a0) u */
    fq2 t8, t9; f2_dbl(&t8, a); f2_dbl(&t8, &t8); f2_dbl(&t8, &t8); f2_add(&t9, &t8, a);
    fp c0, c1; q_sub(&c0, &t9.c0, &a->c1); q_add(&c1, &t9.c1, &a->c0);
    r->c0 = c0; r->c1 = c1;
}
static void f2_inv(fq2 *r, const fq2 *a) {
    fp n, t; q_mul(&n, &a->c0, &a->c0); q_mul(&t, &a->c1, &a->c1); q_add(&n, &n, &t);
    fp_inv(&n, &n, Q);
    q_mul(&r->c0, &a->c0, &n); q_mul(&t, &a->c1, &n); q_neg(&r->c1, &t);
}
static void f2_pow(fq2 *r, const fq2 *a, const uint64_t *e, int nlimbs) {
    fq2 acc; acc.c0 = MQ.one; memset(&acc.c1, 0, sizeof(fp));
    fq2 base = *a;
    for (int i = 0; i < nlimbs; i++)
        for (int b = 0; b < 64; b++) { if ((e[i] >> b) & 1) f2_mul(&acc, &acc, &base); f2_sqr(&base, &base); }
    *r = acc;
}

static void f6_add(fq6 *r, const fq6 *a, const fq6 *b) { f2_add(&r->c0, &a->c0, &b->c0); f2_add(&r->c1, &a->c1, &b->c1); f2_add(&r->c2, &a->c2, &b->c2); }
static void f6_sub(fq6 *r, const fq6 *a, const fq6 *b) { f2_sub(&r->c0, &a->c0, &b->c0); f2_sub(&r->c1, &a->c1, &b->c1); f2_sub(&r->c2, &a->c2, &b->c2); }
static void f6_neg(fq6 *r, const fq6 *a) { f2_neg(&r->c0, &a->c0); f2_neg(&r->c1, &a->c1); f2_neg(&r->c2, &a->c2); }
static void f6_mul_v(fq6 *r, const fq6 *a) { fq2 t; f2_mul_xi(&t, &a->c2); r->c2 = a->c1; r->c1 = a->c0; r->c0 = t; }
static void f6_mul(fq6 *r, const fq6 *a, const fq6 *b) { /* Karatsuba over Fq2: 6 multiplications */
    fq2 v0, v1, v2, s, t, x, c0, c1, c2;
    f2_mul(&v0, &a->c0, &b->c0); f2_mul(&v1, &a->c1, &b->c1); f2_mul(&v2, &a->c2, &b->c2);
    f2_add(&s, &a->c1, &a->c2); f2_add(&t, &b->c1, &b->c2); f2_mul(&x, &s, &t);
    f2_sub(&x, &x, &v1); f2_sub(&x, &x, &v2); f2_mul_xi(&x, &x); f2_add(&c0, &x, &v0);
    f2_add(&s, &a->c0, &a->c1); f2_add(&t, &b->c0, &b->c1); f2_mul(&x, &s, &t);
    f2_sub(&x, &x, &v0); f2_sub(&x, &x, &v1); f2_mul_xi(&s, &v2); f2_add(&c1, &x, &s);
    f2_add(&s, &a->c0, &a->c2); f2_add(&t, &b->c0, &b->c2); f2_mul(&x, &s, &t);
    f2_sub(&x, &x, &v0); f2_sub(&x, &x, &v2); f2_add(&c2, &x, &v1);
    r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static void f6_mul_by_01(fq6 *r, const fq6 *a, const fq2 *b0, const fq2 *b1) { /* a * (b0 + b1 v), 5 multiplications */
    fq2 v0, v1, a2b1, a2b0, s, t, x, c0, c1, c2;
    f2_mul(&v0, &a->c0, b0); f2_mul(&v1, &a->c1, b1); f2_mul(&a2b1, &a->c2, b1); f2_mul(&a2b0, &a->c2, b0);
    f2_add(&s, &a->c0, &a->c1); f2_add(&t, b0, b1); f2_mul(&x, &s, &t);
    f2_mul_xi(&c0, &a2b1); f2_add(&c0, &c0, &v0);
    f2_sub(&c1, &x, &v0); f2_sub(&c1, &c1, &v1);
    f2_add(&c2, &v1, &a2b0);
    r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static void f6_inv(fq6 *r, const fq6 *a) {
    fq2 t0, t1, t2, x, d;
    f2_sqr(&t0, &a->c0); f2_mul(&x, &a->c1, &a->c2); f2_mul_xi(&x, &x); f2_sub(&t0, &t0, &x);
    f2_sqr(&t1, &a->c2); f2_mul_xi(&t1, &t1); f2_mul(&x, &a->c0, &a->c1); f2_sub(&t1, &t1, &x);
    f2_sqr(&t2, &a->c1); f2_mul(&x, &a->c0, &a->c2); f2_sub(&t2, &t2, &x);
    f2_mul(&d, &a->c2, &t1); f2_mul(&x, &a->c1, &t2); f2_add(&d, &d, &x); f2_mul_xi(&d, &d);
    f2_mul(&x, &a->c0, &t0); f2_add(&d, &d, &x);
    f2_inv(&d, &d);
    f2_mul(&r->c0, &t0, &d); f2_mul(&r->c1, &t1, &d); f2_mul(&r->c2, &t2, &d);
}

static void f12_mul(fq12 *r, const fq12 *a, const fq12 *b) {
    fq6 v0, v1, s, t, x;
    f6_mul(&v0, &a->c0, &b->c0); f6_mul(&v1, &a->c1, &b->c1);
    f6_add(&s, &a->c0, &a->c1); f6_add(&t, &b->c0, &b->c1); f6_mul(&x, &s, &t);
    f6_sub(&x, &x, &v0); f6_sub(&r->c1, &x, &v1);
    f6_mul_v(&s, &v1); f6_add(&r->c0, &v0, &s);
}
static void f12_sqr(fq12 *r, const fq12 *a) { /* complex squaring */
    fq6 ab, s, t, u;
    f6_mul(&ab, &a->c0, &a->c1);
    f6_add(&s, &a->c0, &a->c1); f6_mul_v(&t, &a->c1); f6_add(&t, &t, &a->c0);
    f6_mul(&u, &s, &t);
    f6_sub(&u, &u, &ab); f6_mul_v(&s, &ab); f6_sub(&r->c0, &u, &s);
    f6_add(&r->c1, &ab, &ab);
}
static void f12_conj(fq12 *r, const fq12 *a) { r->c0 = a->c0; f6_neg(&r->c1, &a->c1); }
static void f12_inv(fq12 *r, const fq12 *a) {
    fq6 t0, t1;
    f6_mul(&t0, &a->c0, &a->c0); f6_mul(&t1, &a->c1, &a->c1); f6_mul_v(&t1, &t1); f6_sub(&t0, &t0, &t1);
    f6_inv(&t0, &t0);
    f6_mul(&r->c0, &a->c0, &t0); f6_mul(&t1, &a->c1, &t0); f6_neg(&r->c1, &t1);
}
static int f12_eq(const fq12 *a, const fq12 *b) { return memcmp(a, b, sizeof(fq12)) == 0; }
/* w-power view: g[0..5] with f = sum g_i w^i;  c0 = (g0,g2,g4), c1 = (g1,g3,g5) */
static fq2 *f12_g(fq12 *a, int i) { fq6 *h = (i & 1) ? &a->c1 : &a->c0; return (i >> 1) == 0 ? &h->c0 : (i >> 1) == 1 ? &h->c1 : &h->c2; }
static void f12_frob(fq12 *r, const fq12 *a, int k) {
    fq12 t = *a;
    for (int i = 0; i < 6; i++) {
        fq2 *g = f12_g(&t, i);
        if (k & 1) f2_conj(g, g);
        f2_mul(g, g, &GAMMA[k][i]);
    }
    *r = t;
}
/* f * (l0 + l1 w + l3 w^3): the sparse "034" product, 13 Fq2 multiplications */
static void f12_mul_sparse(fq12 *f, const fq2 *l0, const fq2 *l1, const fq2 *l3) {
    fq6 a, b, e, s; fq2 t;
    f2_mul(&a.c0, &f->c0.c0, l0); f2_mul(&a.c1, &f->c0.c1, l0); f2_mul(&a.c2, &f->c0.c2, l0);
    f6_mul_by_01(&b, &f->c1, l1, l3);
    f2_add(&t, l0, l1); f6_add(&s, &f->c0, &f->c1); f6_mul_by_01(&e, &s, &t, l3);
    f6_sub(&e, &e, &a); f6_sub(&f->c1, &e, &b);
    f6_mul_v(&b, &b); f6_add(&f->c0, &a, &b);
}
/* Granger-Scott squaring in the cyclotomic subgroup; Fq12 = Fq4[w]/(w^3 - s), Fq4 = Fq2[s]/(s^2 - xi), s = w^3 */
static void fq4_sqr(fq2 *t0, fq2 *t1, const fq2 *a, const fq2 *b) { /* (a + b s)^2 = (a^2 + xi b^2) + 2ab s */
    fq2 ab, s, x;
    f2_mul(&ab, a, b); f2_add(&s, a, b); f2_mul_xi(&x, b); f2_add(&x, &x, a); f2_mul(&s, &s, &x);
    f2_sub(&s, &s, &ab); f2_mul_xi(&x, &ab); f2_sub(t0, &s, &x);
    f2_dbl(t1, &ab);
}
static void f12_cyc_sqr(fq12 *r, const fq12 *a) {
    fq12 in = *a;
    fq2 *g0 = f12_g(&in, 0), *g1 = f12_g(&in, 1), *g2 = f12_g(&in, 2), *g3 = f12_g(&in, 3), *g4 = f12_g(&in, 4), *g5 = f12_g(&in, 5);
    fq2 a0, a1, b0, b1, c0, c1, t;
    fq4_sqr(&a0, &a1, g0, g3); fq4_sqr(&b0, &b1, g1, g4); fq4_sqr(&c0, &c1, g2, g5);
#define THREE_MINUS_TWO(out, tt, z) do { f2_sub(&t, tt, z); f2_dbl(&t, &t); f2_add(out, &t, tt); } while (0)
#define THREE_PLUS_TWO(out, tt, z)  do { f2_add(&t, tt, z); f2_dbl(&t, &t); f2_add(out, &t, tt); } while (0)
    fq2 n0, n1, n2, n3, n4, n5, xc1;
    THREE_MINUS_TWO(&n0, &a0, g0); THREE_PLUS_TWO(&n3, &a1, g3);
    f2_mul_xi(&xc1, &c1);
    THREE_PLUS_TWO(&n1, &xc1, g1); THREE_MINUS_TWO(&n4, &c0, g4);
    THREE_MINUS_TWO(&n2, &b0, g2); THREE_PLUS_TWO(&n5, &b1, g5);
    *f12_g(r, 0) = n0; *f12_g(r, 1) = n1; *f12_g(r, 2) = n2; *f12_g(r, 3) = n3; *f12_g(r, 4) = n4; *f12_g(r, 5) = n5;
}
static void f12_cyc_exp_x(fq12 *r, const fq12 *a) { /* a^x for the BN parameter x, a in the cyclotomic subgroup */
    fq12 acc = *a;
    for (int b = 61; b >= 0; b--) { /* x has bit 62 set */
        f12_cyc_sqr(&acc, &acc);
        if ((BN_X >> b) & 1) f12_mul(&acc, &acc, a);
    }
    *r = acc;
}
static void f12_pow(fq12 *r, const fq12 *a, const uint64_t *e, int nlimbs) { /* generic square-and-multiply, LSB first */
    fq12 acc = F12_ONE, base = *a;
    for (int i = 0; i < nlimbs; i++)
        for (int b = 0; b < 64; b++) { if ((e[i] >> b) & 1) f12_mul(&acc, &acc, &base); f12_sqr(&base, &base); }
    *r = acc;
}

/* bytes <-> tower */
static int f2_from_bytes(fq2 *r, const uint8_t *b) { return fp_from_bytes(&r->c0, b, Q) | fp_from_bytes(&r->c1, b + 32, Q); }
static void f2_to_bytes(uint8_t *b, const fq2 *a) { fp_to_bytes(b, &a->c0, Q); fp_to_bytes(b + 32, &a->c1, Q); }
static int f12_from_bytes(fq12 *r, const uint8_t *b) {
    int rc = 0; fq2 *p = &r->c0.c0;
    for (int i = 0; i < 6; i++) rc |= f2_from_bytes(p + i, b + 64 * i);
    return rc;
}
static void f12_to_bytes(uint8_t *b, const fq12 *a) {
    const fq2 *p = &a->c0.c0;
    for (int i = 0; i < 6; i++) f2_to_bytes(b + 64 * i, p + i);
}

/* ------------------------------------------------------------------------------------------- */
/* Curves: affine (x, y, inf) at the boundary, Jacobian inside scalar multiplication            */
/* ------------------------------------------------------------------------------------------- */
typedef struct { fp x, y; int inf; } g1a;
typedef struct { fq2 x, y; int inf; } g2a;
typedef struct { fp x, y, z; } g1j;   /* z == 0 <=> identity */
typedef struct { fq2 x, y, z; } g2j;

static int is_zero_bytes(const uint8_t *b, size_t n) { for (size_t i = 0; i < n; i++) if (b[i]) return 0; return 1; }
static int g1_from_bytes(g1a *p, const uint8_t *b) {
    memset(p, 0, sizeof *p);
    if (is_zero_bytes(b, 64)) { p->inf = 1; return 0; }
    return fp_from_bytes(&p->x, b, Q) | fp_from_bytes(&p->y, b + 32, Q);
}
static void g1_to_bytes(uint8_t *b, const g1a *p) {
    if (p->inf) { memset(b, 0, 64); return; }
    fp_to_bytes(b, &p->x, Q); fp_to_bytes(b + 32, &p->y, Q);
}
static int g2_from_bytes(g2a *p, const uint8_t *b) {
    memset(p, 0, sizeof *p);
    if (is_zero_bytes(b, 128)) { p->inf = 1; return 0; }
    return f2_from_bytes(&p->x, b) | f2_from_bytes(&p->y, b + 64);
}
static void g2_to_bytes(uint8_t *b, const g2a *p) {
    if (p->inf) { memset(b, 0, 128); return; }
    f2_to_bytes(b, &p->x); f2_to_bytes(b + 64, &p->y);
}

static void g1j_dbl(g1j *r, const g1j *p) {
    if (fp_is_zero(&p->z)) { *r = *p; return; }
    fp A, B, C, D, E, F, t, X3, Y3, Z3;
    q_mul(&A, &p->x, &p->x); q_mul(&B, &p->y, &p->y); q_mul(&C, &B, &B);
    q_add(&t, &p->x, &B); q_mul(&t, &t, &t); q_sub(&t, &t, &A); q_sub(&t, &t, &C); q_add(&D, &t, &t);
    q_add(&E, &A, &A); q_add(&E, &E, &A); q_mul(&F, &E, &E);
    q_sub(&X3, &F, &D); q_sub(&X3, &X3, &D);
    q_sub(&t, &D, &X3); q_mul(&Y3, &E, &t);
    q_add(&C, &C, &C); q_add(&C, &C, &C); q_add(&C, &C, &C); q_sub(&Y3, &Y3, &C);
    q_mul(&Z3, &p->y, &p->z); q_add(&Z3, &Z3, &Z3);
    r->x = X3; r->y = Y3; r->z = Z3;
}
static void g1j_add_affine(g1j *r, const g1j *p, const g1a *q) { /* complete w.r.t. identity / doubling / inverse */
    if (q->inf) { *r = *p; return; }
    if (fp_is_zero(&p->z)) { r->x = q->x; r->y = q->y; r->z = MQ.one; return; }
    fp zz, u2, s2, h, rr, hh, hhh, v, t, X3, Y3, Z3;
    q_mul(&zz, &p->z, &p->z); q_mul(&u2, &q->x, &zz); q_mul(&s2, &q->y, &p->z); q_mul(&s2, &s2, &zz);
    q_sub(&h, &u2, &p->x); q_sub(&rr, &s2, &p->y);
    if (fp_is_zero(&h)) {
        if (fp_is_zero(&rr)) { g1j_dbl(r, p); return; }
        memset(r, 0, sizeof *r); r->x = MQ.one; r->y = MQ.one; return;
    }
    q_mul(&hh, &h, &h); q_mul(&hhh, &hh, &h); q_mul(&v, &p->x, &hh);
    q_mul(&X3, &rr, &rr); q_sub(&X3, &X3, &hhh); q_sub(&X3, &X3, &v); q_sub(&X3, &X3, &v);
    q_sub(&t, &v, &X3); q_mul(&Y3, &rr, &t); q_mul(&t, &p->y, &hhh); q_sub(&Y3, &Y3, &t);
    q_mul(&Z3, &p->z, &h);
    r->x = X3; r->y = Y3; r->z = Z3;
}
static void g1j_to_affine(g1a *r, const g1j *p) {
    memset(r, 0, sizeof *r);
    if (fp_is_zero(&p->z)) { r->inf = 1; return; }
    fp zi, zi2; fp_inv(&zi, &p->z, Q); q_mul(&zi2, &zi, &zi);
    q_mul(&r->x, &p->x, &zi2); q_mul(&zi2, &zi2, &zi); q_mul(&r->y, &p->y, &zi2);
}
/* ark `Affine * Fr`: MSB-first double-and-add over the canonical scalar */
static void g1_scalar_mul(g1j *r, const g1a *p, const uint64_t k[4]) {
    g1j acc; memset(&acc, 0, sizeof acc); acc.x = MQ.one; acc.y = MQ.one;
    for (int i = 255; i >= 0; i--) {
        g1j_dbl(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) g1j_add_affine(&acc, &acc, p);
    }
    *r = acc;
}

static void g2j_dbl(g2j *r, const g2j *p) {
    if (f2_is_zero(&p->z)) { *r = *p; return; }
    fq2 A, B, C, D, E, F, t, X3, Y3, Z3;
    f2_sqr(&A, &p->x); f2_sqr(&B, &p->y); f2_sqr(&C, &B);
    f2_add(&t, &p->x, &B); f2_sqr(&t, &t); f2_sub(&t, &t, &A); f2_sub(&t, &t, &C); f2_dbl(&D, &t);
    f2_dbl(&E, &A); f2_add(&E, &E, &A); f2_sqr(&F, &E);
    f2_sub(&X3, &F, &D); f2_sub(&X3, &X3, &D);
    f2_sub(&t, &D, &X3); f2_mul(&Y3, &E, &t);
    f2_dbl(&C, &C); f2_dbl(&C, &C); f2_dbl(&C, &C); f2_sub(&Y3, &Y3, &C);
    f2_mul(&Z3, &p->y, &p->z); f2_dbl(&Z3, &Z3);
    r->x = X3; r->y = Y3; r->z = Z3;
}
static void g2j_add_affine(g2j *r, const g2j *p, const g2a *q) {
    if (q->inf) { *r = *p; return; }
    if (f2_is_zero(&p->z)) { r->x = q->x; r->y = q->y; memset(&r->z, 0, sizeof(fq2)); r->z.c0 = MQ.one; return; }
    fq2 zz, u2, s2, h, rr, hh, hhh, v, t, X3, Y3, Z3;
    f2_sqr(&zz, &p->z); f2_mul(&u2, &q->x, &zz); f2_mul(&s2, &q->y, &p->z); f2_mul(&s2, &s2, &zz);
    f2_sub(&h, &u2, &p->x); f2_sub(&rr, &s2, &p->y);
    if (f2_is_zero(&h)) {
        if (f2_is_zero(&rr)) { g2j_dbl(r, p); return; }
        memset(r, 0, sizeof *r); r->x.c0 = MQ.one; r->y.c0 = MQ.one; return;
    }
    f2_sqr(&hh, &h); f2_mul(&hhh, &hh, &h); f2_mul(&v, &p->x, &hh);
    f2_sqr(&X3, &rr); f2_sub(&X3, &X3, &hhh); f2_sub(&X3, &X3, &v); f2_sub(&X3, &X3, &v);
    f2_sub(&t, &v, &X3); f2_mul(&Y3, &rr, &t); f2_mul(&t, &p->y, &hhh); f2_sub(&Y3, &Y3, &t);
    f2_mul(&Z3, &p->z, &h);
    r->x = X3; r->y = Y3; r->z = Z3;
}
static void g2j_to_affine(g2a *r, const g2j *p) {
    memset(r, 0, sizeof *r);
    if (f2_is_zero(&p->z)) { r->inf = 1; return; }
    fq2 zi, zi2; f2_inv(&zi, &p->z); f2_sqr(&zi2, &zi);
    f2_mul(&r->x, &p->x, &zi2); f2_mul(&zi2, &zi2, &zi); f2_mul(&r->y, &p->y, &zi2);
}
static void g2_scalar_mul(g2j *r, const g2a *p, const uint64_t k[4]) {
    g2j acc; memset(&acc, 0, sizeof acc); acc.x.c0 = MQ.one; acc.y.c0 = MQ.one;
    for (int i = 255; i >= 0; i--) {
        g2j_dbl(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) g2j_add_affine(&acc, &acc, p);
    }
    *r = acc;
}
static int g1_on_curve(const g1a *p) {
    if (p->inf) return 1;
    fp l, r, three; q_mul(&l, &p->y, &p->y); q_mul(&r, &p->x, &p->x); q_mul(&r, &r, &p->x);
    fp_from_small(&three, 3, Q); q_add(&r, &r, &three);
    return fp_eq(&l, &r);
}
static int g2_on_curve(const g2a *p) {
    if (p->inf) return 1;
    fq2 l, r; f2_sqr(&l, &p->y); f2_sqr(&r, &p->x); f2_mul(&r, &r, &p->x); f2_add(&r, &r, &B_TWIST);
    return f2_eq(&l, &r);
}

/* ------------------------------------------------------------------------------------------- */
/* Optimal-ate pairing.  Miller loop: homogeneous projective T on the twist, lines               */
/*   l(P) = l0 * yP + l1 * xP * w + l3 * w^3   (sparse in w^0, w^1, w^3)                         */
/* ------------------------------------------------------------------------------------------- */
typedef struct { fq2 x, y, z; } g2h;
static void line_double(g2h *t, fq2 *l0, fq2 *l1, fq2 *l3) {
    /* tangent at T, then T <- 2T.  l0 = -2YZ, l1 = 3X^2, l3 = 3b'Z^2 - Y^2 */
    fq2 a, b, c, e, f, g, h, i, j, e2, s;
    f2_mul(&a, &t->x, &t->y); f2_scale(&a, &a, &TWO_INV);
    f2_sqr(&b, &t->y); f2_sqr(&c, &t->z);
    f2_dbl(&e, &c); f2_add(&e, &e, &c); f2_mul(&e, &e, &B_TWIST);
    f2_dbl(&f, &e); f2_add(&f, &f, &e);
    f2_add(&g, &b, &f); f2_scale(&g, &g, &TWO_INV);
    f2_add(&h, &t->y, &t->z); f2_sqr(&h, &h); f2_add(&s, &b, &c); f2_sub(&h, &h, &s);
    f2_sub(&i, &e, &b);
    f2_sqr(&j, &t->x);
    f2_sqr(&e2, &e);
    f2_sub(&s, &b, &f); f2_mul(&t->x, &a, &s);
    f2_sqr(&g, &g); f2_dbl(&s, &e2); f2_add(&s, &s, &e2); f2_sub(&t->y, &g, &s);
    f2_mul(&t->z, &b, &h);
    f2_neg(l0, &h); f2_dbl(l1, &j); f2_add(l1, l1, &j); *l3 = i;
}
static void line_add(g2h *t, const g2a *q, fq2 *l0, fq2 *l1, fq2 *l3) {
    /* chord through T and Q, then T <- T + Q.  l0 = lambda, l1 = -theta, l3 = theta xQ - lambda yQ */
    fq2 theta, lambda, c, d, e, f, g, h, s, u;
    f2_mul(&s, &q->y, &t->z); f2_sub(&theta, &t->y, &s);
    f2_mul(&s, &q->x, &t->z); f2_sub(&lambda, &t->x, &s);
    f2_sqr(&c, &theta); f2_sqr(&d, &lambda); f2_mul(&e, &lambda, &d); f2_mul(&f, &t->z, &c); f2_mul(&g, &t->x, &d);
    f2_add(&h, &e, &f); f2_sub(&h, &h, &g); f2_sub(&h, &h, &g);
    f2_mul(&t->x, &lambda, &h);
    f2_sub(&s, &g, &h); f2_mul(&s, &theta, &s); f2_mul(&u, &e, &t->y); f2_sub(&t->y, &s, &u);
    f2_mul(&t->z, &t->z, &e);
    f2_mul(&s, &theta, &q->x); f2_mul(&u, &lambda, &q->y); f2_sub(l3, &s, &u);
    *l0 = lambda; f2_neg(l1, &theta);
}
static void eval_and_mul(fq12 *f, const g1a *p, const fq2 *l0, const fq2 *l1, const fq2 *l3) {
    fq2 a, b; f2_scale(&a, l0, &p->y); f2_scale(&b, l1, &p->x);
    f12_mul_sparse(f, &a, &b, l3);
}
static void g2_frob(g2a *r, const g2a *q) {
    f2_conj(&r->x, &q->x); f2_mul(&r->x, &r->x, &GAMMA[1][2]);
    f2_conj(&r->y, &q->y); f2_mul(&r->y, &r->y, &GAMMA[1][3]);
    r->inf = q->inf;
}
static void miller_loop(fq12 *out, const g1a *p, const g2a *q) {
    fq12 f = F12_ONE;
    if (p->inf || q->inf) { *out = f; return; } /* ark convention: identity pairs contribute 1 */
    g2h t; t.x = q->x; t.y = q->y; memset(&t.z, 0, sizeof(fq2)); t.z.c0 = MQ.one;
    g2a nq = *q; f2_neg(&nq.y, &q->y);
    fq2 l0, l1, l3;
    for (int i = 63; i >= 0; i--) {
        if (i != 63) f12_sqr(&f, &f);
        line_double(&t, &l0, &l1, &l3); eval_and_mul(&f, p, &l0, &l1, &l3);
        if (ATE_DIGITS[i] == 1) { line_add(&t, q, &l0, &l1, &l3); eval_and_mul(&f, p, &l0, &l1, &l3); }
        else if (ATE_DIGITS[i] == -1) { line_add(&t, &nq, &l0, &l1, &l3); eval_and_mul(&f, p, &l0, &l1, &l3); }
    }
    g2a q1, q2; g2_frob(&q1, q); g2_frob(&q2, &q1); f2_neg(&q2.y, &q2.y);
    line_add(&t, &q1, &l0, &l1, &l3); eval_and_mul(&f, p, &l0, &l1, &l3);
    line_add(&t, &q2, &l0, &l1, &l3); eval_and_mul(&f, p, &l0, &l1, &l3);
    *out = f;
}
static void final_exp(fq12 *out, const fq12 *f, unsigned opts) {
    /* easy part f^((p^6-1)(p^2+1)), then the Devegili-Scott-Dahab chain (eprint 2008/490 sec. 5):
       exponent exactly (p^12-1)/r  [H1, SURVEY A.1] */
    fq12 t, m, mp, mp2, mp3, mx, mx2, mx3, y0, y1, y2, y3, y4, y5, y6, t0, t1;
    f12_inv(&t, f); f12_conj(&m, f); f12_mul(&t, &m, &t);
    f12_frob(&m, &t, 2); f12_mul(&m, &m, &t);
    f12_frob(&mp, &m, 1); f12_frob(&mp2, &m, 2); f12_frob(&mp3, &m, 3);
    f12_cyc_exp_x(&mx, &m); f12_cyc_exp_x(&mx2, &mx); f12_cyc_exp_x(&mx3, &mx2);
    f12_mul(&y0, &mp, &mp2); f12_mul(&y0, &y0, &mp3);
    f12_conj(&y1, &m);
    f12_frob(&y2, &mx2, 2);
    f12_frob(&y3, &mx, 1); f12_conj(&y3, &y3);
    f12_frob(&y4, &mx2, 1); f12_mul(&y4, &y4, &mx); f12_conj(&y4, &y4);
    f12_conj(&y5, &mx2);
    f12_frob(&y6, &mx3, 1); f12_mul(&y6, &y6, &mx3); f12_conj(&y6, &y6);
    f12_cyc_sqr(&t0, &y6); f12_mul(&t0, &t0, &y4); f12_mul(&t0, &t0, &y5);
    f12_mul(&t1, &y3, &y5); f12_mul(&t1, &t1, &t0);
    f12_mul(&t0, &t0, &y2);
    f12_cyc_sqr(&t1, &t1); f12_mul(&t1, &t1, &t0); f12_cyc_sqr(&t1, &t1);
    f12_mul(&t0, &t1, &y1);
    f12_mul(&t1, &t1, &y0);
    f12_cyc_sqr(&t0, &t0); f12_mul(&t0, &t0, &t1);
    if (opts & SIPP_ORACLE_FE_ARK) { /* arkworks value = exact ^ (2x(6x^2+3x+1)) */
        fq12 a, b, c;
        f12_cyc_exp_x(&a, &t0);              /* ^x */
        f12_cyc_exp_x(&b, &a);               /* ^x^2 */
        f12_cyc_exp_x(&c, &b);               /* ^x^3 */
        /* 6x^3 + 3x^2 + x = x + 3x^2 + 6x^3 ; then square for the factor 2 */
        fq12 b3, c6;
        f12_cyc_sqr(&b3, &b); f12_mul(&b3, &b3, &b);
        f12_cyc_sqr(&c6, &c); f12_mul(&c6, &c6, &c); f12_cyc_sqr(&c6, &c6);
        f12_mul(&a, &a, &b3); f12_mul(&a, &a, &c6); f12_cyc_sqr(&t0, &a);
    }
    *out = t0;
}

/* ------------------------------------------------------------------------------------------- */
/* Poseidon over Goldilocks (plonky2), constants regenerated from ChaCha8(seed 0)  */
/* ------------------------------------------------------------------------------------------- */
#define GL_P 0xFFFFFFFF00000001ULL
static const uint64_t MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static uint64_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    /* 2^64 = 2^32 - 1, 2^96 = -1 (mod p) */
    uint64_t t = lo - hh; if (lo < hh) t -= 0xFFFFFFFFULL; /* borrow: + p = subtract 2^32-1 in wrapped arithmetic */
    uint64_t m = hl * 0xFFFFFFFFULL;
    uint64_t r = t + m; if (r < m) r += 0xFFFFFFFFULL;
    if (r >= GL_P) r -= GL_P;
    return r;
}
static uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((u128)a * b); }
static uint64_t gl_add(uint64_t a, uint64_t b) { uint64_t r = a + b; if (r < a || r >= GL_P) r -= GL_P; return r; }
static uint64_t gl_pow7(uint64_t x) { uint64_t x2 = gl_mul(x, x), x3 = gl_mul(x2, x), x4 = gl_mul(x2, x2); return gl_mul(x3, x4); }

static uint32_t rotl32(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
#define CHACHA_QR(a, b, c, d) do { s[a] += s[b]; s[d] = rotl32(s[d] ^ s[a], 16); s[c] += s[d]; s[b] = rotl32(s[b] ^ s[c], 12); \
                                   s[a] += s[b]; s[d] = rotl32(s[d] ^ s[a], 8);  s[c] += s[d]; s[b] = rotl32(s[b] ^ s[c], 7); } while (0)
static void poseidon_gen_constants(void) {
    /* rand_chacha ChaCha8Rng::seed_from_u64(0): key from PCG32 stream, then rand 0.8 gen_range(0..p) by widening multiply */
    uint32_t key[8]; uint64_t st = 0;
    for (int i = 0; i < 8; i++) {
        st = st * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27); int rot = (int)(st >> 59);
        key[i] = rot ? ((xs >> rot) | (xs << (32 - rot))) : xs;
    }
    uint32_t buf[16]; int pos = 16; uint64_t ctr = 0; int n = 0;
    while (n < 360) {
        uint32_t w[2];
        for (int k = 0; k < 2; k++) {
            if (pos == 16) {
                uint32_t init[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                                     (uint32_t)ctr, (uint32_t)(ctr >> 32), 0, 0};
                uint32_t s[16]; memcpy(s, init, sizeof s);
                for (int r = 0; r < 4; r++) {
                    CHACHA_QR(0, 4, 8, 12); CHACHA_QR(1, 5, 9, 13); CHACHA_QR(2, 6, 10, 14); CHACHA_QR(3, 7, 11, 15);
                    CHACHA_QR(0, 5, 10, 15); CHACHA_QR(1, 6, 11, 12); CHACHA_QR(2, 7, 8, 13); CHACHA_QR(3, 4, 9, 14);
                }
                for (int i = 0; i < 16; i++) buf[i] = s[i] + init[i];
                pos = 0; ctr++;
            }
            w[k] = buf[pos++];
        }
        uint64_t v = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
        u128 m = (u128)v * GL_P;
        if ((uint64_t)m <= GL_P - 1) POSEIDON_RC[n++] = (uint64_t)(m >> 64);
    }
}
void oracle_poseidon_perm(uint64_t s[12]) {
    int rnd = 0;
    for (int phase = 0; phase < 3; phase++) {
        int nr = phase == 1 ? 22 : 4;
        for (int k = 0; k < nr; k++, rnd++) {
            for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], POSEIDON_RC[12 * rnd + i]);
            if (phase == 1) s[0] = gl_pow7(s[0]);
            else for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
            uint64_t o[12];
            for (int r = 0; r < 12; r++) {
                u128 acc = 0; /* 12 * 41 * 2^64 < 2^74: no overflow */
                for (int i = 0; i < 12; i++) acc += (u128)s[(i + r) % 12] * MDS_CIRC[i];
                if (r == 0) acc += (u128)s[0] * 8; /* MDS diagonal = [8, 0, ..., 0] */
                o[r] = gl_reduce128(acc);
            }
            memcpy(s, o, sizeof o);
        }
    }
}
void oracle_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]) {
    uint64_t s[12] = {0};
    for (size_t off = 0; off < n; off += 8) {
        size_t len = n - off < 8 ? n - off : 8;
        memcpy(s, in + off, len * 8); /* overwrite mode */
        oracle_poseidon_perm(s);
    }
    memcpy(out, s, 32);
}

/* ------------------------------------------------------------------------------------------- */
/* start-up                                                                                     */
/* ------------------------------------------------------------------------------------------- */
static pthread_once_t INIT_ONCE = PTHREAD_ONCE_INIT;
static void div_small(uint64_t *q, const uint64_t *a, int n, uint64_t d) { /* q = a / d */
    u128 rem = 0;
    for (int i = n - 1; i >= 0; i--) { u128 cur = (rem << 64) | a[i]; q[i] = (uint64_t)(cur / d); rem = cur % d; }
}
static void dec_to_fp(fp *r, const char *dec) {
    fp acc, ten, d; memset(&acc, 0, sizeof acc); fp_from_small(&ten, 10, Q);
    for (; *dec; dec++) { q_mul(&acc, &acc, &ten); fp_from_small(&d, (uint64_t)(*dec - '0'), Q); q_add(&acc, &acc, &d); }
    *r = acc;
}
static void init_all(void) {
    modulus_init(&MQ); modulus_init(&MR);
    fp two; fp_from_small(&two, 2, Q); fp_inv(&TWO_INV, &two, Q);
    fp_from_small(&XI.c0, 9, Q); XI.c1 = MQ.one;
    fq2 xi_inv; f2_inv(&xi_inv, &XI);
    fp three; fp_from_small(&three, 3, Q); f2_scale(&B_TWIST, &xi_inv, &three);
    memset(&F12_ONE, 0, sizeof F12_ONE); F12_ONE.c0.c0.c0 = MQ.one;
    /* gamma_{1,1} = xi^((p-1)/6); gamma_{1,i} = gamma_{1,1}^i; gamma_2 = gamma_1 * conj(gamma_1); gamma_3 = gamma_1 * gamma_2 */
    uint64_t e[4], pm1[4]; const uint64_t one[4] = {1, 0, 0, 0};
    sub4(pm1, MQ.m, one); div_small(e, pm1, 4, 6);
    fq2 g11; f2_pow(&g11, &XI, e, 4);
    memset(GAMMA, 0, sizeof GAMMA);
    GAMMA[1][0].c0 = MQ.one;
    for (int i = 1; i < 6; i++) f2_mul(&GAMMA[1][i], &GAMMA[1][i - 1], &g11);
    for (int i = 0; i < 6; i++) {
        fq2 c; f2_conj(&c, &GAMMA[1][i]);
        f2_mul(&GAMMA[2][i], &GAMMA[1][i], &c);
        f2_mul(&GAMMA[3][i], &GAMMA[1][i], &GAMMA[2][i]);
    }
    fp_from_small(&G1GEN[0], 1, Q); fp_from_small(&G1GEN[1], 2, Q);
    dec_to_fp(&G2GEN[0].c0, "10857046999023057135944570762232829481370756359578518086990519993285655852781");
    dec_to_fp(&G2GEN[0].c1, "11559732032986387107991004021392285783925812861821192530917403151452391805634");
    dec_to_fp(&G2GEN[1].c0, "8495653923123431417604973247489272438418190587263600148770280649306958101930");
    dec_to_fp(&G2GEN[1].c1, "4082367875863433681332203403145435568316851327593401208105741076214120093531");
    poseidon_gen_constants();
}
static void ensure_init(void) { pthread_once(&INIT_ONCE, init_all); }
void oracle_poseidon_round_constants(uint64_t out[360]) { ensure_init(); memcpy(out, POSEIDON_RC, sizeof POSEIDON_RC); }

/* ------------------------------------------------------------------------------------------- */
/* tiny pthread parallel-for                                                                    */
/* ------------------------------------------------------------------------------------------- */
typedef void (*range_fn)(void *ctx, size_t lo, size_t hi, int tid);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; int tid; } par_job;
static void *par_thunk(void *a) { par_job *j = a; j->fn(j->ctx, j->lo, j->hi, j->tid); return NULL; }
static void parallel_for(size_t n, int threads, range_fn fn, void *ctx) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    if (threads == 1) { fn(ctx, 0, n, 0); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads); par_job *jobs = malloc(sizeof(par_job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (par_job){fn, ctx, n * (size_t)t / (size_t)threads, n * (size_t)(t + 1) / (size_t)threads, t};
        pthread_create(&th[t], NULL, par_thunk, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* ------------------------------------------------------------------------------------------- */
/* exported: field ops                                                                          */
/* ------------------------------------------------------------------------------------------- */
int oracle_field_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t count) {
    ensure_init();
    for (size_t i = 0; i < count; i++) {
        if (op < 10) {
            fp x, y, r; memset(&y, 0, sizeof y);
            if (fp_from_bytes(&x, a + 32 * i, Q)) return -1;
            if (b && fp_from_bytes(&y, b + 32 * i, Q)) return -1;
            switch (op) {
                case ORC_FQ_MUL: q_mul(&r, &x, &y); break;
                case ORC_FQ_ADD: q_add(&r, &x, &y); break;
                case ORC_FQ_SUB: q_sub(&r, &x, &y); break;
                case ORC_FQ_INV: fp_inv(&r, &x, Q); break;
                case ORC_FQ_SQR: q_mul(&r, &x, &x); break;
                default: return -2;
            }
            fp_to_bytes(out + 32 * i, &r, Q);
        } else if (op < 20) {
            fq2 x, y, r; memset(&y, 0, sizeof y);
            if (f2_from_bytes(&x, a + 64 * i)) return -1;
            if (b && f2_from_bytes(&y, b + 64 * i)) return -1;
            switch (op) {
                case ORC_FQ2_MUL: f2_mul(&r, &x, &y); break;
                case ORC_FQ2_SQR: f2_sqr(&r, &x); break;
                case ORC_FQ2_INV: f2_inv(&r, &x); break;
                case ORC_FQ2_MUL_XI: f2_mul_xi(&r, &x); break;
                default: return -2;
            }
            f2_to_bytes(out + 64 * i, &r);
        } else {
            fq12 x, y, r; memset(&y, 0, sizeof y);
            if (f12_from_bytes(&x, a + 384 * i)) return -1;
            if (b && f12_from_bytes(&y, b + 384 * i)) return -1;
            switch (op) {
                case ORC_FQ12_MUL: f12_mul(&r, &x, &y); break;
                case ORC_FQ12_SQR: f12_sqr(&r, &x); break;
                case ORC_FQ12_INV: f12_inv(&r, &x); break;
                case ORC_FQ12_FROB1: f12_frob(&r, &x, 1); break;
                case ORC_FQ12_FROB2: f12_frob(&r, &x, 2); break;
                case ORC_FQ12_FROB3: f12_frob(&r, &x, 3); break;
                case ORC_FQ12_CONJ: f12_conj(&r, &x); break;
                case ORC_FQ12_CYC_SQR: f12_cyc_sqr(&r, &x); break;
                default: return -2;
            }
            f12_to_bytes(out + 384 * i, &r);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* exported: curve ops and folds                                                                */
/* ------------------------------------------------------------------------------------------- */
static int scalar_from_bytes(uint64_t k[4], const uint8_t *b) {
    for (int i = 0; i < 4; i++) { k[i] = 0; for (int j = 7; j >= 0; j--) k[i] = (k[i] << 8) | b[8 * i + j]; }
    return geq4(k, MR.m) ? -1 : 0;
}
int oracle_g1_mul(const uint8_t a[64], const uint8_t k[32], uint8_t out[64]) {
    ensure_init(); g1a p, r; g1j j; uint64_t s[4];
    if (g1_from_bytes(&p, a) || scalar_from_bytes(s, k)) return -1;
    g1_scalar_mul(&j, &p, s); g1j_to_affine(&r, &j); g1_to_bytes(out, &r); return 0;
}
int oracle_g2_mul(const uint8_t b[128], const uint8_t k[32], uint8_t out[128]) {
    ensure_init(); g2a p, r; g2j j; uint64_t s[4];
    if (g2_from_bytes(&p, b) || scalar_from_bytes(s, k)) return -1;
    g2_scalar_mul(&j, &p, s); g2j_to_affine(&r, &j); g2_to_bytes(out, &r); return 0;
}
int oracle_g1_on_curve(const uint8_t a[64]) { ensure_init(); g1a p; if (g1_from_bytes(&p, a)) return 0; return g1_on_curve(&p); }
int oracle_g2_on_curve(const uint8_t b[128]) { ensure_init(); g2a p; if (g2_from_bytes(&p, b)) return 0; return g2_on_curve(&p); }

/* new_A[i] = (a1 + a2.mul(x)).into()   prover_native.rs:60-64 */
static void fold_g1_one(g1a *out, const g1a *a1, const g1a *a2, const uint64_t x[4]) {
    g1j t; g1_scalar_mul(&t, a2, x); g1j_add_affine(&t, &t, a1); g1j_to_affine(out, &t);
}
/* new_B[i] = (b1 + b2.mul(inv_x)).into()   prover_native.rs:65-69 */
static void fold_g2_one(g2a *out, const g2a *b1, const g2a *b2, const uint64_t x[4]) {
    g2j t; g2_scalar_mul(&t, b2, x); g2j_add_affine(&t, &t, b1); g2j_to_affine(out, &t);
}
int oracle_fold_g1(const uint8_t *A, size_t n, const uint8_t x[32], uint8_t *out) {
    ensure_init(); uint64_t s[4]; if (scalar_from_bytes(s, x)) return -1;
    size_t h = n / 2;
    for (size_t i = 0; i < h; i++) {
        g1a a1, a2, r; if (g1_from_bytes(&a1, A + 64 * i) || g1_from_bytes(&a2, A + 64 * (i + h))) return -1;
        fold_g1_one(&r, &a1, &a2, s); g1_to_bytes(out + 64 * i, &r);
    }
    return 0;
}
int oracle_fold_g2(const uint8_t *B, size_t n, const uint8_t xinv[32], uint8_t *out) {
    ensure_init(); uint64_t s[4]; if (scalar_from_bytes(s, xinv)) return -1;
    size_t h = n / 2;
    for (size_t i = 0; i < h; i++) {
        g2a b1, b2, r; if (g2_from_bytes(&b1, B + 128 * i) || g2_from_bytes(&b2, B + 128 * (i + h))) return -1;
        fold_g2_one(&r, &b1, &b2, s); g2_to_bytes(out + 128 * i, &r);
    }
    return 0;
}
int oracle_fr_inverse(const uint8_t x[32], uint8_t out[32]) {
    ensure_init(); fp v; if (fp_from_bytes(&v, x, &MR)) return -1;
    if (fp_is_zero(&v)) return -3; /* x.inverse().unwrap() panics, prover_native.rs:58 */
    fp_inv(&v, &v, &MR); fp_to_bytes(out, &v, &MR); return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* exported: pairing and inner product                                                          */
/* ------------------------------------------------------------------------------------------- */
int oracle_miller_loop(const uint8_t a[64], const uint8_t b[128], uint8_t out[384]) {
    ensure_init(); g1a p; g2a q; fq12 f;
    if (g1_from_bytes(&p, a) || g2_from_bytes(&q, b)) return -1;
    miller_loop(&f, &p, &q); f12_to_bytes(out, &f); return 0;
}
int oracle_final_exp(const uint8_t f[384], uint8_t out[384], unsigned opts) {
    ensure_init(); fq12 x, r; if (f12_from_bytes(&x, f)) return -1;
    final_exp(&r, &x, opts); f12_to_bytes(out, &r); return 0;
}
static void pairing_pt(fq12 *out, const g1a *p, const g2a *q, unsigned opts) { fq12 f; miller_loop(&f, p, q); final_exp(out, &f, opts); }
int oracle_pairing(const uint8_t a[64], const uint8_t b[128], uint8_t out[384], unsigned opts) {
    ensure_init(); g1a p; g2a q; fq12 f;
    if (g1_from_bytes(&p, a) || g2_from_bytes(&q, b)) return -1;
    pairing_pt(&f, &p, &q, opts); f12_to_bytes(out, &f); return 0;
}

typedef struct { const g1a *A; const g2a *B; unsigned opts; fq12 *partial; } ip_ctx;
static void ip_range(void *c, size_t lo, size_t hi, int tid) {
    ip_ctx *x = c; fq12 acc = F12_ONE, t;
    for (size_t i = lo; i < hi; i++) {
        if (x->opts & SIPP_ORACLE_FAITHFUL) pairing_pt(&t, &x->A[i], &x->B[i], x->opts); /* pairing(*a, *b), prover_native.rs:20 */
        else miller_loop(&t, &x->A[i], &x->B[i]);
        f12_mul(&acc, &acc, &t);                                                          /* fold(Fq12::one(), acc * x), :22 */
    }
    x->partial[tid] = acc;
}
static void inner_product_pts(fq12 *out, const g1a *A, const g2a *B, size_t n, unsigned opts, int threads) {
    if (threads < 1) threads = 1;
    fq12 *partial = malloc(sizeof(fq12) * (size_t)threads);
    for (int t = 0; t < threads; t++) partial[t] = F12_ONE;
    ip_ctx c = {A, B, opts, partial};
    parallel_for(n, threads, ip_range, &c);
    fq12 acc = F12_ONE;
    for (int t = 0; t < threads; t++) f12_mul(&acc, &acc, &partial[t]);
    free(partial);
    if (!(opts & SIPP_ORACLE_FAITHFUL)) final_exp(&acc, &acc, opts);
    *out = acc;
}
static int load_points(const uint8_t *A, const uint8_t *B, size_t n, g1a **pa, g2a **pb) {
    *pa = malloc(sizeof(g1a) * (n ? n : 1)); *pb = malloc(sizeof(g2a) * (n ? n : 1));
    for (size_t i = 0; i < n; i++)
        if (g1_from_bytes(&(*pa)[i], A + 64 * i) || g2_from_bytes(&(*pb)[i], B + 128 * i)) { free(*pa); free(*pb); return -1; }
    return 0;
}
int oracle_inner_product_mt(const uint8_t *A, const uint8_t *B, size_t n, uint8_t out[384], unsigned opts, int threads) {
    ensure_init(); g1a *pa; g2a *pb; fq12 z;
    if (load_points(A, B, n, &pa, &pb)) return -1;
    inner_product_pts(&z, pa, pb, n, opts, threads);
    f12_to_bytes(out, &z); free(pa); free(pb); return 0;
}
int oracle_inner_product(const uint8_t *A, const uint8_t *B, size_t n, uint8_t out[384], unsigned opts) {
    return oracle_inner_product_mt(A, B, n, out, opts, 1);
}

/* ------------------------------------------------------------------------------------------- */
/* exported: transcript                                                                         */
/* ------------------------------------------------------------------------------------------- */
void oracle_transcript_new(oracle_transcript *t) { ensure_init(); memset(t, 0, sizeof *t); }
void oracle_transcript_append(oracle_transcript *t, const uint64_t *msg, size_t n) { /* transcript_native.rs:25-30 */
    uint64_t buf[4 + 96];
    if (n > 96) return;
    memcpy(buf, t->state, 32); memcpy(buf + 4, msg, n * 8);
    oracle_hash_no_pad(buf, 4 + n, t->state);
    t->perms += (4 + n + 7) / 8;
}
static void fq_bytes_to_limbs(uint64_t *out, const uint8_t *b) { /* from_fq_to_f, transcript_native.rs:68-77 */
    for (int i = 0; i < 8; i++) out[i] = (uint64_t)b[4 * i] | ((uint64_t)b[4 * i + 1] << 8) | ((uint64_t)b[4 * i + 2] << 16) | ((uint64_t)b[4 * i + 3] << 24);
}
void oracle_transcript_append_g1(oracle_transcript *t, const uint8_t a[64]) { /* :42-46 */
    uint64_t m[16]; fq_bytes_to_limbs(m, a); fq_bytes_to_limbs(m + 8, a + 32); oracle_transcript_append(t, m, 16);
}
void oracle_transcript_append_g2(oracle_transcript *t, const uint8_t b[128]) { /* :48-54 */
    uint64_t m[32]; for (int i = 0; i < 4; i++) fq_bytes_to_limbs(m + 8 * i, b + 32 * i); oracle_transcript_append(t, m, 32);
}
void oracle_transcript_append_fq12(oracle_transcript *t, const uint8_t f[384], unsigned opts) { /* :32-40 */
    uint64_t m[96];
    if (opts & SIPP_ORACLE_FQ12_NESTED) { for (int i = 0; i < 12; i++) fq_bytes_to_limbs(m + 8 * i, f + 32 * i); }
    else {
        /* MyFq12.coeffs (H2): coeffs[i] = g_i.c0, coeffs[i+6] = g_i.c1, with g_i the Fq2 coefficient of w^i.
           nested slot of g_i: c0=(g0,g2,g4), c1=(g1,g3,g5) -> index (i&1)*3 + (i>>1) */
        for (int i = 0; i < 6; i++) {
            int slot = (i & 1) * 3 + (i >> 1);
            fq_bytes_to_limbs(m + 8 * i, f + 64 * slot);
            fq_bytes_to_limbs(m + 8 * (i + 6), f + 64 * slot + 32);
        }
    }
    oracle_transcript_append(t, m, 96);
}
void oracle_challenge_from_digest(const uint64_t digest[4], uint8_t x[32]) { /* :58-64 and SURVEY A.4 */
    ensure_init();
    uint32_t digits[8]; int nd = 0;
    for (int k = 0; k < 4; k++) { /* BigUint::to_u32_digits: little endian, high zero digits stripped, 0 -> none */
        uint64_t d = digest[k];
        while (d) { digits[nd++] = (uint32_t)d; d >>= 32; }
    }
    uint64_t v[4] = {0, 0, 0, 0};
    for (int j = 0; j < nd; j++) v[j >> 1] |= (uint64_t)digits[j] << (32 * (j & 1));
    /* reduce mod r: v < 2^256 < 6r */
    while (geq4(v, MR.m)) sub4(v, v, MR.m);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) x[8 * i + j] = (uint8_t)(v[i] >> (8 * j));
}
void oracle_transcript_get_challenge(const oracle_transcript *t, uint8_t x[32]) { /* :56-65, does not mutate state */
    uint64_t digest[4]; oracle_hash_no_pad(t->state, 4, digest);
    oracle_challenge_from_digest(digest, x);
}

/* ------------------------------------------------------------------------------------------- */
/* exported: prover and verifier                                                                */
/* ------------------------------------------------------------------------------------------- */
typedef struct { const g1a *A; const g2a *B; g1a *nA; g2a *nB; size_t h; const uint64_t *x; const uint64_t *xinv; } fold_ctx;
static void fold_range(void *c, size_t lo, size_t hi, int tid) {
    (void)tid; fold_ctx *f = c;
    for (size_t i = lo; i < hi; i++) {
        fold_g1_one(&f->nA[i], &f->A[i], &f->A[i + f->h], f->x);
        fold_g2_one(&f->nB[i], &f->B[i], &f->B[i + f->h], f->xinv);
    }
}
static int challenge_and_inverse(const oracle_transcript *tr, uint64_t x[4], uint64_t xinv[4]) {
    uint8_t xb[32], ib[32];
    oracle_transcript_get_challenge(tr, xb);            /* let x = transcript.get_challenge(); */
    if (oracle_fr_inverse(xb, ib)) return -3;           /* let inv_x = x.inverse().unwrap();   */
    scalar_from_bytes(x, xb); scalar_from_bytes(xinv, ib);
    return 0;
}
static void scalar_to_bytes(uint8_t *b, const uint64_t k[4]) { for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[8 * i + j] = (uint8_t)(k[i] >> (8 * j)); }

int oracle_sipp_prove(const uint8_t *Ab, const uint8_t *Bb, size_t n, uint8_t *proof_out, unsigned opts, int threads,
                      uint8_t *challenges, uint8_t *foldedA, uint8_t *foldedB) {
    ensure_init();
    if (n == 0 || (n & (n - 1))) return -2;
    g1a *A; g2a *B;
    if (load_points(Ab, Bb, n, &A, &B)) return -1;
    size_t rounds = 0; for (size_t m = n; m > 1; m >>= 1) rounds++;
    size_t np = 2 * rounds + 1, k = 0;
    fq12 *proof = malloc(sizeof(fq12) * np);
    uint8_t zb[384];
    fq12 Z; inner_product_pts(&Z, A, B, n, opts, threads);                 /* let Z = inner_product(A, B);  :29 */
    oracle_transcript tr; oracle_transcript_new(&tr);                      /* :32 */
    for (size_t i = 0; i < n; i++) {                                       /* register A and B  :36-39 */
        uint8_t pb[128];
        g1_to_bytes(pb, &A[i]); oracle_transcript_append_g1(&tr, pb);
        g2_to_bytes(pb, &B[i]); oracle_transcript_append_g2(&tr, pb);
    }
    proof[k++] = Z; f12_to_bytes(zb, &Z); oracle_transcript_append_fq12(&tr, zb, opts);   /* :42-43 */
    size_t foff = 0, round = 0; int rc = 0;
    while (n > 1) {                                                        /* :45 */
        size_t h = n / 2;
        fq12 ZL, ZR;
        inner_product_pts(&ZL, A + h, B, h, opts, threads);                /* Z_L = inner_product(A2, B1)  :48 */
        inner_product_pts(&ZR, A, B + h, h, opts, threads);                /* Z_R = inner_product(A1, B2)  :49 */
        proof[k++] = ZL; f12_to_bytes(zb, &ZL); oracle_transcript_append_fq12(&tr, zb, opts);  /* :52-53 */
        proof[k++] = ZR; f12_to_bytes(zb, &ZR); oracle_transcript_append_fq12(&tr, zb, opts);  /* :54-55 */
        uint64_t x[4], xinv[4];
        if ((rc = challenge_and_inverse(&tr, x, xinv))) break;             /* :57-58 */
        if (challenges) scalar_to_bytes(challenges + 32 * round, x);
        g1a *nA = malloc(sizeof(g1a) * h); g2a *nB = malloc(sizeof(g2a) * h);
        fold_ctx fc = {A, B, nA, nB, h, x, xinv};
        parallel_for(h, threads, fold_range, &fc);                         /* :60-69 */
        for (size_t i = 0; i < h; i++) {
            if (foldedA) g1_to_bytes(foldedA + 64 * (foff + i), &nA[i]);
            if (foldedB) g2_to_bytes(foldedB + 128 * (foff + i), &nB[i]);
        }
        foff += h; round++;
        free(A); free(B); A = nA; B = nB; n = h;                           /* :72-74 */
    }
    if (!rc) for (size_t i = 0; i < np; i++) f12_to_bytes(proof_out + 384 * i, &proof[np - 1 - i]);  /* proof.reverse()  :78 */
    free(A); free(B); free(proof);
    return rc;
}

int oracle_sipp_verify(const uint8_t *Ab, const uint8_t *Bb, size_t n, const uint8_t *proofb, size_t proof_len,
                       unsigned opts, int threads, uint8_t *final_A, uint8_t *final_B, uint8_t *final_Z) {
    ensure_init();
    if (n == 0 || (n & (n - 1))) return -2;
    size_t rounds = 0; for (size_t m = n; m > 1; m >>= 1) rounds++;
    if (proof_len < 2 * rounds + 1) return -4;                             /* proof.pop().unwrap() would panic, :31,:40,:42 */
    g1a *A; g2a *B;
    if (load_points(Ab, Bb, n, &A, &B)) return -1;
    oracle_transcript tr; oracle_transcript_new(&tr);
    for (size_t i = 0; i < n; i++) {                                       /* :25-28 */
        uint8_t pb[128];
        g1_to_bytes(pb, &A[i]); oracle_transcript_append_g1(&tr, pb);
        g2_to_bytes(pb, &B[i]); oracle_transcript_append_g2(&tr, pb);
    }
    size_t top = proof_len;                                                /* pop from the back */
    fq12 Z, ZL, ZR;
    if (f12_from_bytes(&Z, proofb + 384 * --top)) { free(A); free(B); return -1; }          /* :31-33 */
    oracle_transcript_append_fq12(&tr, proofb + 384 * top, opts);
    int rc = 0;
    while (n > 1) {
        size_t h = n / 2;
        if (f12_from_bytes(&ZL, proofb + 384 * --top)) { rc = -1; break; }                  /* :40-41 */
        oracle_transcript_append_fq12(&tr, proofb + 384 * top, opts);
        if (f12_from_bytes(&ZR, proofb + 384 * --top)) { rc = -1; break; }                  /* :42-43 */
        oracle_transcript_append_fq12(&tr, proofb + 384 * top, opts);
        uint64_t x[4], xinv[4];
        if ((rc = challenge_and_inverse(&tr, x, xinv))) break;                              /* :45-46 */
        g1a *nA = malloc(sizeof(g1a) * h); g2a *nB = malloc(sizeof(g2a) * h);
        fold_ctx fc = {A, B, nA, nB, h, x, xinv};
        parallel_for(h, threads, fold_range, &fc);                                          /* :48-57 */
        fq12 l, r;
        f12_pow(&l, &ZL, x, 4); f12_pow(&r, &ZR, xinv, 4);                                  /* :59-61 */
        f12_mul(&l, &l, &Z); f12_mul(&Z, &l, &r);
        free(A); free(B); A = nA; B = nB; n = h;
    }
    if (rc) { free(A); free(B); return rc; }
    if (final_A) g1_to_bytes(final_A, &A[0]);
    if (final_B) g2_to_bytes(final_B, &B[0]);
    if (final_Z) f12_to_bytes(final_Z, &Z);
    fq12 e; pairing_pt(&e, &A[0], &B[0], opts);                                             /* :80 */
    int ok = f12_eq(&e, &Z);
    free(A); free(B);
    return ok;
}

/* ------------------------------------------------------------------------------------------- */
/* exported: seeded inputs                                                                      */
/* ------------------------------------------------------------------------------------------- */
void oracle_seeded_scalars(uint64_t seed, size_t n, uint8_t *scalars) {
    ensure_init();
    uint64_t s = seed;
    for (size_t i = 0; i < 2 * n; i++) {
        uint64_t v[4];
        for (int j = 0; j < 4; j++) {
            s += 0x9E3779B97F4A7C15ULL;
            uint64_t z = s;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            v[j] = z ^ (z >> 31);
        }
        while (geq4(v, MR.m)) sub4(v, v, MR.m);
        if (!(v[0] | v[1] | v[2] | v[3])) v[0] = 1;
        scalar_to_bytes(scalars + 32 * i, v);
    }
}
typedef struct { const uint8_t *sc; uint8_t *A, *B; } seed_ctx;
static void seed_range(void *c, size_t lo, size_t hi, int tid) {
    (void)tid; seed_ctx *s = c;
    g1a g1; g2a g2; memset(&g1, 0, sizeof g1); memset(&g2, 0, sizeof g2);
    g1.x = G1GEN[0]; g1.y = G1GEN[1]; g2.x = G2GEN[0]; g2.y = G2GEN[1];
    for (size_t i = lo; i < hi; i++) {
        uint64_t k[4]; g1j a; g2j b; g1a aa; g2a ba;
        scalar_from_bytes(k, s->sc + 64 * i); g1_scalar_mul(&a, &g1, k); g1j_to_affine(&aa, &a); g1_to_bytes(s->A + 64 * i, &aa);
        scalar_from_bytes(k, s->sc + 64 * i + 32); g2_scalar_mul(&b, &g2, k); g2j_to_affine(&ba, &b); g2_to_bytes(s->B + 128 * i, &ba);
    }
}
int oracle_seeded_inputs(uint64_t seed, size_t n, uint8_t *A, uint8_t *B, int threads) {
    ensure_init();
    uint8_t *sc = malloc(64 * (n ? n : 1));
    oracle_seeded_scalars(seed, n, sc);
    seed_ctx c = {sc, A, B};
    parallel_for(n, threads, seed_range, &c);
    free(sc);
    return 0;
}
