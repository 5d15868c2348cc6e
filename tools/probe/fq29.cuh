// fq.cuh -- BN254 base field Fq on NINE SIGNED 29-BIT LIMBS (radix 2^29, Montgomery R = 2^261), for sm_100a.
//
// Replaces ark-ff `Fp256<MontBackend<FqConfig,4>>` arithmetic that every call on the reference hot path
// bottoms out in (/root/reference/src/prover_native.rs:20,63,68 via ark-bn254 0.4).
//
// Why not 8 x 32-bit limbs with carry chains: on B200 the carry-in/carry-out form of the wide multiply-add
// (IMAD.WIDE.U32.X) issues at HALF the rate of the plain IMAD.WIDE (measured: 8.55 vs 18.4 T/s) and every
// row of a saturated-limb product is one long dependent chain.  With 29-bit limbs a 29x29 product has 58 bits, so a
// 64-bit column accumulator takes dozens of products with plain full-rate IMAD.WIDE and NO carries between columns;
// all rows are independent (ILP), additions and subtractions are nine independent 32-bit adds with no conditional
// subtraction, and sums of products need one Montgomery reduction for the whole sum (fq_dot).
//
// Representation.  value = sum_i l[i] 2^(29 i), limbs SIGNED.  A value is "normalised" when |l[i]| <= 2^28 (+ a few
// units) for i < 8; l[8] is whatever is left (|value| < 2^k p  =>  |l[8]| < 2^(k+22)).  Values are residues mod p but
// NOT canonical: any integer congruent to the element, of either sign, is valid.  Bounds (checked exhaustively over
// the executed paths by tests/hostcheck with SIPP_FQ_BOUNDS, which tracks worst-case limb and value bounds):
//   * columns are int64: a sum of products  sum_n a_n b_n  is allowed when  sum_n LB(a_n) LB(b_n) <= 13, where LB is the
//     limb bound in units of 2^28 (9 terms per column, plus 9 reduction terms of 2^56: 9 * 13 * 2^56 + 9 * 2^56 < 2^63);
//   * Montgomery reduction with a balanced quotient digit returns |r| <= |T| / R + p / 2, and p / R = 2^-7.4, so products
//     of values below 8p come back below 0.9p: magnitudes never grow through multiplications;
//   * add / sub / neg / small multiples are limb-wise and only grow LB; fq_norm (one parallel carry pass) restores LB ~ 1.
// Canonical form (for the boundary bytes and for equality tests) is produced only by fq_to_canonical / fq_is_zero.
//
// The same code compiles for the host (tests/hostcheck runs the device algorithms on the CPU against the oracle);
// the product never computes on the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
// device-only under nvcc: the host never computes field arithmetic in the product (no CPU fallback)
#define SIPP_HD __device__ __forceinline__
#define SIPP_HD_NOINLINE static __device__ __noinline__
#else
#define SIPP_HD inline
#define SIPP_HD_NOINLINE inline
#endif

#if defined(SIPP_FQ_BOUNDS) && !defined(__CUDACC__)
#include <cstdio>
#include <cstdlib>
#define SIPP_BOUNDS 1
#else
#define SIPP_BOUNDS 0
#endif

namespace sipp {

#define SIPP_FQ_LIMBS 9
#define SIPP_FQ_BITS 29
#define SIPP_FQ_MASK 0x1fffffffu
#define SIPP_FQ_SUM_BUDGET 13.0  // sum_n LB(a_n) LB(b_n) allowed in one column accumulation

struct Fq {
    int32_t l[SIPP_FQ_LIMBS];
#if SIPP_BOUNDS
    double lb = 1.0;  // worst-case |limb| / 2^28 over limbs 0..7
    double vb = 1.0;  // worst-case |value| / p
#endif
};

#if SIPP_BOUNDS
struct FqBoundsStats {
    double max_sum = 0, max_lb = 0, max_vb = 0;
};
inline FqBoundsStats& fq_bounds_stats() { static FqBoundsStats s; return s; }
inline void fq_bounds_fail(const char* what, double v) {
    fprintf(stderr, "SIPP_FQ_BOUNDS violation: %s = %g\n", what, v);
    abort();
}
#define SIPP_SET_BOUNDS(r, L, V) do { (r).lb = (L); (r).vb = (V); if ((r).lb > fq_bounds_stats().max_lb) fq_bounds_stats().max_lb = (r).lb; \
                                      if ((r).vb > fq_bounds_stats().max_vb) fq_bounds_stats().max_vb = (r).vb; \
                                      if ((r).lb > 7.0) fq_bounds_fail("limb bound (int32 headroom)", (r).lb); } while (0)
#else
#define SIPP_SET_BOUNDS(r, L, V) do { } while (0)
#endif

// p = sum PB[i] 2^(29 i), balanced digits
#define SIPP_PB0 (-126026425)
#define SIPP_PB1 17064119
#define SIPP_PB2 (-59595953)
#define SIPP_PB3 47522513
#define SIPP_PB4 (-175777416)
#define SIPP_PB5 47923393
#define SIPP_PB6 10936641
#define SIPP_PB7 240920116
#define SIPP_PB8 3171406
#define SIPP_PINV29 0x04866389u  // -p^-1 mod 2^29
// p, little-endian 32-bit words (range checks at the boundary)
#define SIPP_P0 0xd87cfd47u
#define SIPP_P1 0x3c208c16u
#define SIPP_P2 0x6871ca8du
#define SIPP_P3 0x97816a91u
#define SIPP_P4 0x8181585du
#define SIPP_P5 0xb85045b6u
#define SIPP_P6 0xe131a029u
#define SIPP_P7 0x30644e72u

SIPP_HD int32_t fq_pb(int i) {
    switch (i) {
        case 0: return SIPP_PB0; case 1: return SIPP_PB1; case 2: return SIPP_PB2; case 3: return SIPP_PB3; case 4: return SIPP_PB4;
        case 5: return SIPP_PB5; case 6: return SIPP_PB6; case 7: return SIPP_PB7; default: return SIPP_PB8;
    }
}
SIPP_HD uint32_t fq_p_word(int i) {
    switch (i) {
        case 0: return SIPP_P0; case 1: return SIPP_P1; case 2: return SIPP_P2; case 3: return SIPP_P3;
        case 4: return SIPP_P4; case 5: return SIPP_P5; case 6: return SIPP_P6; default: return SIPP_P7;
    }
}

SIPP_HD Fq fq_make(int32_t a0, int32_t a1, int32_t a2, int32_t a3, int32_t a4, int32_t a5, int32_t a6, int32_t a7, int32_t a8) {
    Fq r;
    r.l[0] = a0; r.l[1] = a1; r.l[2] = a2; r.l[3] = a3; r.l[4] = a4; r.l[5] = a5; r.l[6] = a6; r.l[7] = a7; r.l[8] = a8;
    SIPP_SET_BOUNDS(r, 1.0, 1.0);
    return r;
}
// R mod p (Montgomery one) and R^2 mod p, balanced digits
SIPP_HD Fq fq_one() { return fq_make(-176370655, -199481511, -128831276, 21759002, 178483129, -45989682, -237679608, 86689705, 903222); }
SIPP_HD Fq fq_r2() { return fq_make(94088208, 219480995, 25171640, -257225560, 40052282, 46143135, -157549229, -242836147, 2757031); }
SIPP_HD Fq fq_zero() {
    Fq r = fq_make(0, 0, 0, 0, 0, 0, 0, 0, 0);
    SIPP_SET_BOUNDS(r, 0.0, 0.0);
    return r;
}

SIPP_HD int32_t fq_sext29(uint32_t x) { return (int32_t)(x << 3) >> 3; }

// ---------------------------------------------------------------------------------------------------------
// limb-wise linear operations (no carries, no reduction)
// ---------------------------------------------------------------------------------------------------------
SIPP_HD Fq fq_add(const Fq& a, const Fq& b) {
    Fq r;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) r.l[i] = a.l[i] + b.l[i];
    SIPP_SET_BOUNDS(r, a.lb + b.lb, a.vb + b.vb);
    return r;
}
SIPP_HD Fq fq_sub(const Fq& a, const Fq& b) {
    Fq r;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) r.l[i] = a.l[i] - b.l[i];
    SIPP_SET_BOUNDS(r, a.lb + b.lb, a.vb + b.vb);
    return r;
}
SIPP_HD Fq fq_neg(const Fq& a) {
    Fq r;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) r.l[i] = -a.l[i];
    SIPP_SET_BOUNDS(r, a.lb, a.vb);
    return r;
}
SIPP_HD Fq fq_dbl(const Fq& a) {
    Fq r;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) r.l[i] = a.l[i] * 2;
    SIPP_SET_BOUNDS(r, 2 * a.lb, 2 * a.vb);
    return r;
}
// a * k for a small non-negative compile-time-ish constant
SIPP_HD Fq fq_mul_small(const Fq& a, int k) {
    Fq r;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) r.l[i] = a.l[i] * k;
    SIPP_SET_BOUNDS(r, k * a.lb, k * a.vb);
    return r;
}
SIPP_HD Fq fq_select(bool c, const Fq& a, const Fq& b) {
    Fq r;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) r.l[i] = c ? a.l[i] : b.l[i];
    SIPP_SET_BOUNDS(r, a.lb > b.lb ? a.lb : b.lb, a.vb > b.vb ? a.vb : b.vb);
    return r;
}

// one parallel carry pass: limbs 0..7 back to [-2^28, 2^28) plus an incoming carry of a few units
SIPP_HD Fq fq_norm(const Fq& a) {
    Fq r;
    int32_t c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c[i] = (a.l[i] + (1 << 28)) >> SIPP_FQ_BITS;
        r.l[i] = a.l[i] - (c[i] << SIPP_FQ_BITS);
    }
#pragma unroll
    for (int i = 1; i < 8; i++) r.l[i] += c[i - 1];
    r.l[8] = a.l[8] + c[7];
    SIPP_SET_BOUNDS(r, 1.0 + (a.lb + 1.0) / 2.0 / (1 << 28) * 2.0, a.vb);
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// column accumulator: 17 int64 columns; row i of a product adds a.l[0..8] * b_i into columns i..i+8
// ---------------------------------------------------------------------------------------------------------
struct FqCols {
    int64_t t[18];
#if SIPP_BOUNDS
    double sum = 0;  // sum_n LB(a_n) LB(b_n) accumulated so far
    double val = 0;  // sum_n VB(a_n) VB(b_n)
#endif
};
SIPP_HD void cols_clear(FqCols& c) {
#pragma unroll
    for (int i = 0; i < 18; i++) c.t[i] = 0;
}
template <bool SUB>
SIPP_HD void cols_row(FqCols& c, int i, const Fq& a, int32_t bi) {
    const int32_t k = SUB ? -bi : bi;
#pragma unroll
    for (int j = 0; j < SIPP_FQ_LIMBS; j++) c.t[i + j] += (int64_t)a.l[j] * k;
}
// Montgomery step i: make column i divisible by 2^29 with a balanced quotient digit and push its carry up
SIPP_HD void cols_reduce_step(FqCols& c, int i) {
    const int32_t m = fq_sext29((uint32_t)c.t[i] * SIPP_PINV29);
    c.t[i + 0] += (int64_t)m * SIPP_PB0;
    c.t[i + 1] += (int64_t)m * SIPP_PB1;
    c.t[i + 2] += (int64_t)m * SIPP_PB2;
    c.t[i + 3] += (int64_t)m * SIPP_PB3;
    c.t[i + 4] += (int64_t)m * SIPP_PB4;
    c.t[i + 5] += (int64_t)m * SIPP_PB5;
    c.t[i + 6] += (int64_t)m * SIPP_PB6;
    c.t[i + 7] += (int64_t)m * SIPP_PB7;
    c.t[i + 8] += (int64_t)m * SIPP_PB8;
    c.t[i + 1] += c.t[i] >> SIPP_FQ_BITS;
}
// columns 9..16 (+ carry) -> normalised limbs
SIPP_HD Fq cols_finish(const FqCols& c) {
    Fq r;
    int64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int64_t v = c.t[9 + k] + carry;
        r.l[k] = fq_sext29((uint32_t)v);
        carry = (v + (1 << 28)) >> SIPP_FQ_BITS;
    }
    r.l[8] = (int32_t)carry;
#if SIPP_BOUNDS
    if (c.sum > SIPP_FQ_SUM_BUDGET) fq_bounds_fail("column budget sum LB(a) LB(b)", c.sum);
    if (c.sum > fq_bounds_stats().max_sum) fq_bounds_stats().max_sum = c.sum;
    SIPP_SET_BOUNDS(r, 1.0, c.val * 0.0059073 + 0.5 + 1e-6);
#endif
    return r;
}
#if SIPP_BOUNDS
#define SIPP_COLS_ACCOUNT(c, a, b) do { (c).sum += (a).lb * (b).lb; (c).val += (a).vb * (b).vb; } while (0)
#else
#define SIPP_COLS_ACCOUNT(c, a, b) do { } while (0)
#endif

// sum_n (+-) a[n] * b[n]  with ONE Montgomery reduction; bit n of NEGMASK set = subtract product n
template <int N, unsigned NEGMASK>
SIPP_HD Fq fq_dot_signed(const Fq* a, const Fq* b) {
    FqCols c;
    cols_clear(c);
#pragma unroll
    for (int n = 0; n < N; n++) SIPP_COLS_ACCOUNT(c, a[n], b[n]);
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) {
#pragma unroll
        for (int n = 0; n < N; n++) {
            if ((NEGMASK >> n) & 1u) cols_row<true>(c, i, a[n], b[n].l[i]);
            else cols_row<false>(c, i, a[n], b[n].l[i]);
        }
        cols_reduce_step(c, i);
    }
    return cols_finish(c);
}
template <int N>
SIPP_HD Fq fq_dot(const Fq* a, const Fq* b) { return fq_dot_signed<N, 0u>(a, b); }

SIPP_HD Fq fq_mul(const Fq& a, const Fq& b) { return fq_dot_signed<1, 0u>(&a, &b); }

// a^2: the 36 off-diagonal products are computed once against the doubled operand
SIPP_HD Fq fq_sqr(const Fq& a) {
    FqCols c;
    cols_clear(c);
    SIPP_COLS_ACCOUNT(c, a, a);
    int32_t d[SIPP_FQ_LIMBS];
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) d[i] = a.l[i] * 2;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) {
        // row i: a_i^2 into column 2i, 2 a_i a_j (j > i) into column i + j.  Column i is complete after rows <= i/2 ... i
        // have run only if every pair (j, i - j) has been visited: pairs are visited at row min(j, i - j) <= i, so it is.
        c.t[2 * i] += (int64_t)a.l[i] * a.l[i];
#pragma unroll
        for (int j = i + 1; j < SIPP_FQ_LIMBS; j++) c.t[i + j] += (int64_t)d[i] * a.l[j];
        cols_reduce_step(c, i);
    }
    return cols_finish(c);
}

// Montgomery reduction of a plain value: a / R mod p, |result| <= p/2 + |a|/R
SIPP_HD Fq fq_redc(const Fq& a) {
    FqCols c;
    cols_clear(c);
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) c.t[i] = a.l[i];
#if SIPP_BOUNDS
    c.sum = 0; c.val = a.vb / 0.0059073 * 1e-9;  // |a| / R is negligible
#endif
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) cols_reduce_step(c, i);
    return cols_finish(c);
}

// ---------------------------------------------------------------------------------------------------------
// predicates and canonical form
// ---------------------------------------------------------------------------------------------------------
// a == 0 (mod p).  a / R lies in (-p, p) after one reduction, where the only multiple of p is 0, and the balanced
// digit representation of 0 is all-zero limbs.
SIPP_HD bool fq_is_zero(const Fq& a) {
    const Fq r = fq_redc(a);
    int32_t o = 0;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) o |= r.l[i];
    return o == 0;
}
SIPP_HD bool fq_eq(const Fq& a, const Fq& b) { return fq_is_zero(fq_sub(a, b)); }

// full sequential carry to digits in [0, 2^29) with a signed top digit; returns the top digit
SIPP_HD int32_t fq_carry_unsigned(int32_t* d, const Fq& a) {
    int32_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int32_t v = a.l[i] + carry;
        d[i] = v & (int32_t)SIPP_FQ_MASK;
        carry = v >> SIPP_FQ_BITS;
    }
    d[8] = a.l[8] + carry;
    return d[8];
}
// Montgomery form -> canonical integer in [0, p) as 8 little-endian 32-bit words
SIPP_HD void fq_to_canonical(uint32_t* w, const Fq& a) {
    const Fq r = fq_redc(a);  // a / R, in (-p, p)
    Fq rp;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) rp.l[i] = r.l[i] + fq_pb(i);
    int32_t d0[SIPP_FQ_LIMBS], d1[SIPP_FQ_LIMBS];
    const int32_t top = fq_carry_unsigned(d0, r);
    fq_carry_unsigned(d1, rp);
    uint32_t d[SIPP_FQ_LIMBS];
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) d[i] = (uint32_t)(top < 0 ? d1[i] : d0[i]);
    // 9 x 29 bits -> 8 x 32 bits
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int bit = 32 * k, i = bit / SIPP_FQ_BITS, s = bit - i * SIPP_FQ_BITS;
        uint32_t v = d[i] >> s;
        if (i + 1 < SIPP_FQ_LIMBS) v |= d[i + 1] << (SIPP_FQ_BITS - s);
        if (s > 26 && i + 2 < SIPP_FQ_LIMBS) v |= d[i + 2] << (2 * SIPP_FQ_BITS - s);
        w[k] = v;
    }
}
// canonical integer (8 little-endian 32-bit words, < 2^256) -> Montgomery form
SIPP_HD Fq fq_from_words_plain(const uint32_t* w) {  // the integer itself as digits in [0, 2^29), no Montgomery factor
    Fq r;
#pragma unroll
    for (int i = 0; i < SIPP_FQ_LIMBS; i++) {
        const int bit = SIPP_FQ_BITS * i, k = bit >> 5, s = bit & 31;
        uint32_t v = w[k] >> s;
        if (s + SIPP_FQ_BITS > 32 && k + 1 < 8) v |= w[k + 1] << (32 - s);
        r.l[i] = (int32_t)(i < 8 ? (v & SIPP_FQ_MASK) : v);
    }
    SIPP_SET_BOUNDS(r, 2.0, 6.0);  // digits < 2^29, value < 2^256 < 5.3 p
    return r;
}
SIPP_HD Fq fq_from_canonical(const uint32_t* w) { return fq_mul(fq_from_words_plain(w), fq_r2()); }
SIPP_HD Fq fq_to_mont(const Fq& plain) { return fq_mul(plain, fq_r2()); }

// a^(p-2) by square-and-multiply over the bits of p-2 (uniform control flow: the exponent is a constant)
SIPP_HD_NOINLINE Fq fq_inv(const Fq& a_in) {
    const Fq a = fq_norm(a_in);
    Fq acc = fq_one();
    for (int i = 7; i >= 0; i--) {
        uint32_t w = fq_p_word(i) - (i == 0 ? 2u : 0u);  // p-2: low word 0xd87cfd47 - 2, no borrow
        for (int b = 31; b >= 0; b--) {
            acc = fq_sqr(acc);
            if ((w >> b) & 1u) acc = fq_mul(acc, a);
        }
    }
    return acc;
}

}  // namespace sipp
