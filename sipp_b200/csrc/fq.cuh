// fq.cuh -- BN254 base field Fq on 8 x 32-bit limbs, Montgomery form (R = 2^256), for sm_100a.
//
// Replaces ark-ff `Fp256<MontBackend<FqConfig,4>>` arithmetic that every call on the reference hot path
// bottoms out in (/root/reference/src/prover_native.rs:20,63,68 via ark-bn254 0.4).
//
// Multiplication is an interleaved (CIOS) Montgomery product kept in TWO accumulators so that every
// 32x32->64 partial product lands on a 64-bit aligned register pair: even-indexed limbs of `a` (and of p)
// feed the "aligned" accumulator, odd-indexed limbs feed the accumulator that sits 32 bits higher.  Each row
// is one `mad.lo.cc / madc.hi.cc` carry chain (ptxas fuses each lo/hi pair into one IMAD.WIDE.U32 with
// carry), the two rows of a step are independent chains (ILP 2), and the per-step shift by one limb is a
// role swap of the two accumulators plus a one-limb fix-up -- no data movement.
//
// The same row primitives have a plain-C++ host implementation (only used by tests/hostcheck, which runs the
// *device algorithms* on the CPU against the oracle; the product never computes on the host).
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

namespace sipp {

struct Fq {
    uint32_t l[8];
};

// p, little-endian 32-bit limbs
#define SIPP_P0 0xd87cfd47u
#define SIPP_P1 0x3c208c16u
#define SIPP_P2 0x6871ca8du
#define SIPP_P3 0x97816a91u
#define SIPP_P4 0x8181585du
#define SIPP_P5 0xb85045b6u
#define SIPP_P6 0xe131a029u
#define SIPP_P7 0x30644e72u
#define SIPP_PINV 0xe4866389u  // -p^-1 mod 2^32

SIPP_HD uint32_t fq_p_limb(int i) {
    switch (i) {
        case 0: return SIPP_P0; case 1: return SIPP_P1; case 2: return SIPP_P2; case 3: return SIPP_P3;
        case 4: return SIPP_P4; case 5: return SIPP_P5; case 6: return SIPP_P6; default: return SIPP_P7;
    }
}

// R mod p (Montgomery one) and R^2 mod p
SIPP_HD Fq fq_one() { return Fq{{0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}}; }
SIPP_HD Fq fq_r2() { return Fq{{0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u}}; }
SIPP_HD Fq fq_zero() { return Fq{{0, 0, 0, 0, 0, 0, 0, 0}}; }

// ---------------------------------------------------------------------------------------------------------
// row primitives
// ---------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)

// acc[0..7] = (a0, a2, a4, a6) * b laid out as four aligned 64-bit products (no accumulate)
__device__ __forceinline__ void row_mul(uint32_t* acc, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=&r"(acc[0]), "=&r"(acc[1]), "=&r"(acc[2]), "=&r"(acc[3]), "=&r"(acc[4]), "=&r"(acc[5]), "=&r"(acc[6]), "=&r"(acc[7])
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}

// acc[0..7] += (a0, a2, a4, a6) * b ; carry out of limb 7 is added into `top`
__device__ __forceinline__ void row_mad_carry(uint32_t* acc, uint32_t& top, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}

// acc[0..7] += (a0, a2, a4, a6) * b ; no carry out (caller guarantees the bound)
__device__ __forceinline__ void row_mad(uint32_t* acc, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}

// The role swap of one CIOS step:  x0 += y[1] (carry c);  y <- (y >> 64) + (a1, a3, a5, a7) * b + c
__device__ __forceinline__ void row_shift_mad(uint32_t& x0, uint32_t* y, uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7, uint32_t b) {
    asm("add.cc.u32 %8, %8, %1;\n\t"
        "madc.lo.cc.u32 %0, %9, %13, %2;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %4;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %6;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, 0;\n\t"
        "madc.hi.u32 %7, %12, %13, 0;"
        : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7]), "+r"(x0)
        : "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
}

// r = (x >> 32) + y  over 8 limbs (x[0] is known to be zero)
__device__ __forceinline__ void merge_acc(uint32_t* r, const uint32_t* x, const uint32_t* y) {
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]),
          "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
}

// r = a + b over 8 limbs, returns carry out
__device__ __forceinline__ uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
}

// r = a - b over 8 limbs, returns borrow (1 if a < b)
__device__ __forceinline__ uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t c;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c & 1u;
}

#else  // host emulation of the same primitives (tests/hostcheck only)

inline void row_mul(uint32_t* acc, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    const uint32_t a[4] = {a0, a2, a4, a6};
    for (int j = 0; j < 4; j++) {
        uint64_t p = (uint64_t)a[j] * b;
        acc[2 * j] = (uint32_t)p; acc[2 * j + 1] = (uint32_t)(p >> 32);
    }
}
inline uint32_t row_mad_host(uint32_t* acc, const uint32_t* base, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b, uint32_t cin) {
    const uint32_t a[4] = {a0, a2, a4, a6};
    uint64_t carry = cin;
    for (int j = 0; j < 4; j++) {
        uint64_t p = (uint64_t)a[j] * b;
        uint64_t lo = (uint64_t)(uint32_t)p + base[2 * j] + carry;
        uint64_t hi = (p >> 32) + base[2 * j + 1] + (lo >> 32);
        acc[2 * j] = (uint32_t)lo; acc[2 * j + 1] = (uint32_t)hi; carry = hi >> 32;
    }
    return (uint32_t)carry;
}
inline void row_mad_carry(uint32_t* acc, uint32_t& top, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    uint32_t base[8]; for (int i = 0; i < 8; i++) base[i] = acc[i];
    top += row_mad_host(acc, base, a0, a2, a4, a6, b, 0);
}
inline void row_mad(uint32_t* acc, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    uint32_t base[8]; for (int i = 0; i < 8; i++) base[i] = acc[i];
    row_mad_host(acc, base, a0, a2, a4, a6, b, 0);
}
inline void row_shift_mad(uint32_t& x0, uint32_t* y, uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7, uint32_t b) {
    uint64_t s = (uint64_t)x0 + y[1];
    x0 = (uint32_t)s;
    uint32_t base[8] = {y[2], y[3], y[4], y[5], y[6], y[7], 0, 0};
    row_mad_host(y, base, a1, a3, a5, a7, b, (uint32_t)(s >> 32));
}
inline void merge_acc(uint32_t* r, const uint32_t* x, const uint32_t* y) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)y[i] + (i < 7 ? x[i + 1] : 0); r[i] = (uint32_t)c; c >>= 32; }
}
inline uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
inline uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a[i] - b[i] - br; r[i] = (uint32_t)d; br = (d >> 32) & 1; }
    return (uint32_t)br;
}
#endif

// ---------------------------------------------------------------------------------------------------------
// field operations (all inputs and outputs canonical: in [0, p))
// ---------------------------------------------------------------------------------------------------------
SIPP_HD void fq_cond_sub_p(uint32_t* r) {  // r in [0, 2p) -> [0, p)
    const uint32_t P[8] = {SIPP_P0, SIPP_P1, SIPP_P2, SIPP_P3, SIPP_P4, SIPP_P5, SIPP_P6, SIPP_P7};
    uint32_t t[8];
    uint32_t borrow = sub8(t, r, P);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : t[i];
}

SIPP_HD Fq fq_add(const Fq& a, const Fq& b) {
    Fq r;
    add8(r.l, a.l, b.l);  // < 2p < 2^255: no carry out
    fq_cond_sub_p(r.l);
    return r;
}

SIPP_HD Fq fq_sub(const Fq& a, const Fq& b) {
    const uint32_t P[8] = {SIPP_P0, SIPP_P1, SIPP_P2, SIPP_P3, SIPP_P4, SIPP_P5, SIPP_P6, SIPP_P7};
    Fq r;
    uint32_t t[8];
    uint32_t borrow = sub8(r.l, a.l, b.l);
    add8(t, r.l, P);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = borrow ? t[i] : r.l[i];
    return r;
}

SIPP_HD bool fq_is_zero(const Fq& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7]) == 0; }
SIPP_HD bool fq_eq(const Fq& a, const Fq& b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a.l[i] ^ b.l[i];
    return d == 0;
}

SIPP_HD Fq fq_neg(const Fq& a) {
    const uint32_t P[8] = {SIPP_P0, SIPP_P1, SIPP_P2, SIPP_P3, SIPP_P4, SIPP_P5, SIPP_P6, SIPP_P7};
    Fq r;
    sub8(r.l, P, a.l);
    bool z = fq_is_zero(a);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = z ? 0u : r.l[i];
    return r;
}

SIPP_HD Fq fq_dbl(const Fq& a) { return fq_add(a, a); }

// one CIOS step on the accumulator pair (x = aligned, y = 32 bits higher); roles swap between steps
SIPP_HD void fq_mont_reduce_step(uint32_t* x, uint32_t* y) {
    uint32_t m = x[0] * SIPP_PINV;
    row_mad(y, SIPP_P1, SIPP_P3, SIPP_P5, SIPP_P7, m);
    row_mad_carry(x, y[7], SIPP_P0, SIPP_P2, SIPP_P4, SIPP_P6, m);
}
SIPP_HD void fq_mont_first_step(uint32_t* x, uint32_t* y, const Fq& a, uint32_t b) {
    row_mul(y, a.l[1], a.l[3], a.l[5], a.l[7], b);
    row_mul(x, a.l[0], a.l[2], a.l[4], a.l[6], b);
    fq_mont_reduce_step(x, y);
}
SIPP_HD void fq_mont_step(uint32_t* x, uint32_t* y, const Fq& a, uint32_t b) {
    // incoming: value = y_arr + x_arr[1] + 2^32 (x_arr >> 64) with x_arr[0] == 0.  `x` here is the NEW aligned
    // accumulator (the old y), `y` the new high accumulator (the old x).
    row_shift_mad(x[0], y, a.l[1], a.l[3], a.l[5], a.l[7], b);
    row_mad_carry(x, y[7], a.l[0], a.l[2], a.l[4], a.l[6], b);
    fq_mont_reduce_step(x, y);
}

SIPP_HD Fq fq_mul(const Fq& a, const Fq& b) {
    uint32_t e[8], o[8];
    fq_mont_first_step(e, o, a, b.l[0]);
    fq_mont_step(o, e, a, b.l[1]);
    fq_mont_step(e, o, a, b.l[2]);
    fq_mont_step(o, e, a, b.l[3]);
    fq_mont_step(e, o, a, b.l[4]);
    fq_mont_step(o, e, a, b.l[5]);
    fq_mont_step(e, o, a, b.l[6]);
    fq_mont_step(o, e, a, b.l[7]);
    // after the last step the aligned accumulator is `o` (o[0] == 0), the high one is `e`
    Fq r;
    merge_acc(r.l, o, e);
    fq_cond_sub_p(r.l);
    return r;
}

SIPP_HD Fq fq_sqr(const Fq& a) { return fq_mul(a, a); }

// portable reference multiplication (64-bit C arithmetic, no inline PTX); used by the K0 parity test to
// cross-check the carry-chain version on the device and as the "naive" arm of the microbenchmark
SIPP_HD Fq fq_mul_portable(const Fq& a, const Fq& b) {
    uint32_t t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { c += (uint64_t)a.l[j] * b.l[i] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
        c += t[8]; t[8] = (uint32_t)c; t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * SIPP_PINV;
        c = (uint64_t)m * fq_p_limb(0) + t[0]; c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; j++) { c += (uint64_t)m * fq_p_limb(j) + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
        c += t[8]; t[7] = (uint32_t)c; t[8] = t[9] + (uint32_t)(c >> 32);
    }
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = t[i];
    fq_cond_sub_p(r.l);
    return r;
}

SIPP_HD Fq fq_to_mont(const Fq& canonical) { return fq_mul(canonical, fq_r2()); }
SIPP_HD Fq fq_from_mont(const Fq& a) { return fq_mul(a, Fq{{1, 0, 0, 0, 0, 0, 0, 0}}); }

// a^(p-2) by square-and-multiply over the bits of p-2 (uniform control flow: the exponent is a constant).  Kept as the
// reference implementation for the parity tests; the kernels use fq_inv (binary, ~10x fewer instructions).
SIPP_HD_NOINLINE Fq fq_inv_fermat(const Fq& a) {
    Fq acc = fq_one();
    for (int i = 7; i >= 0; i--) {
        uint32_t w = fq_p_limb(i) - (i == 0 ? 2u : 0u);  // p-2: low limb 0xd87cfd47 - 2, no borrow
        for (int b = 31; b >= 0; b--) {
            acc = fq_sqr(acc);
            if ((w >> b) & 1u) acc = fq_mul(acc, a);
        }
    }
    return acc;
}

// ---- binary (Kaliski) Montgomery inversion --------------------------------------------------------------------------
// Phase 1 ("almost inverse"): u = p, v = a, r = 0, s = 1; every iteration halves u or v (or their difference) and
// doubles r or s, keeping  a r = -u 2^k,  a s = v 2^k  (mod p).  It ends with v = 0, u = 1 after k in [254, 508] iterations:
// x = p - r = a^-1 2^k.  Phase 2 multiplies by 2^(512 - k) with Montgomery products: for the Montgomery representative
// a = A R this gives (A R)^-1 2^k 2^(512-k) = A^-1 R.  ~30 cheap instructions per iteration instead of 380 dependent Fq
// multiplications (Fermat): what the fold's affine conversion and the final exponentiation spend their latency on.
// Data-dependent control flow: lanes of a warp diverge over four short paths, which is still several times cheaper.
SIPP_HD void limbs_shr1(uint32_t* x) {
#pragma unroll
    for (int i = 0; i < 7; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[7] >>= 1;
}
SIPP_HD void limbs_shl1(uint32_t* x) {
#pragma unroll
    for (int i = 7; i > 0; i--) x[i] = (x[i] << 1) | (x[i - 1] >> 31);
    x[0] <<= 1;
}
SIPP_HD_NOINLINE Fq fq_inv_kaliski(const Fq& a) {
    if (fq_is_zero(a)) return fq_zero();  // same convention as a^(p-2)
    const uint32_t P[8] = {SIPP_P0, SIPP_P1, SIPP_P2, SIPP_P3, SIPP_P4, SIPP_P5, SIPP_P6, SIPP_P7};
    uint32_t u[8], v[8], r[8], s[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = P[i]; v[i] = a.l[i]; r[i] = 0; s[i] = 0; }
    s[0] = 1;
    int k = 0;
    for (;;) {
        uint32_t vz = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) vz |= v[i];
        if (vz == 0) break;
        if (!(u[0] & 1u)) {
            limbs_shr1(u); limbs_shl1(s);
        } else if (!(v[0] & 1u)) {
            limbs_shr1(v); limbs_shl1(r);
        } else if (sub8(t, v, u) != 0) {  // v < u:  u = (u - v) / 2, r += s, s *= 2
            sub8(u, u, v); limbs_shr1(u);
            add8(r, r, s); limbs_shl1(s);
        } else {                          // v >= u: v = (v - u) / 2, s += r, r *= 2
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = t[i];
            limbs_shr1(v);
            add8(s, s, r); limbs_shl1(r);
        }
        k++;
    }
    // r < 2p: bring to [0, p), then x = p - r
    fq_cond_sub_p(r);
    Fq x;
    sub8(x.l, P, r);
    // x * 2^(512 - k), 512 - k in [4, 258]: two factors 2^e1 2^e2 with e1, e2 <= 253 (plain powers of two are < p)
    const int e = 512 - k;
    const int e1 = e > 253 ? 253 : e, e2 = e - e1;
    Fq p1 = fq_zero(), p2 = fq_zero();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        p1.l[i] = (e1 >> 5) == i ? (1u << (e1 & 31)) : 0u;
        p2.l[i] = (e2 >> 5) == i ? (1u << (e2 & 31)) : 0u;
    }
    x = fq_mul(fq_mul(x, fq_r2()), p1);   // x 2^e1
    return fq_mul(fq_mul(x, fq_r2()), p2);  // x 2^e1 2^e2
}

// ---- inversion by divsteps (Bernstein-Yang "safegcd", the variable-time form with 30-bit signed limbs) -------------------
// f = p, g = a, d = 0, e = 1 and the invariants  d a = f,  e a = g  (mod p).  An outer iteration runs 30 divsteps on the low
// words of f and g alone, collecting a 2x2 transition matrix t with  t (f, g) = 2^30 (f', g'),  and applies t / 2^30 to (f, g)
// exactly and to (d, e) modulo p.  g reaches 0 after at most 20 outer iterations (19 observed on 10^6 values; 590 divsteps is
// the proven bound for 256 bits), f = +-1, and a^-1 = +-d.  The binary algorithm above does one 256-bit shift / compare / add per
// BIT; here the per-bit work is on single words and the 256-bit updates happen once per 30 bits.  Measured in place: fold of an
// n = 2^12 prove 8.2 -> 7.8 ms (12 launches, one inversion on the chain of each), final exponentiation launch 0.912 -> 0.90 ms.
// Data-dependent trip counts only (the inner loop strips trailing zeros of g).
SIPP_HD int fq_ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
SIPP_HD_NOINLINE Fq fq_inv(const Fq& a) {
    if (fq_is_zero(a)) return fq_zero();  // same convention as a^(p-2)
    const int32_t M30 = 0x3fffffff;
    const int32_t PM[9] = {0x187cfd47, 0x3082305b, 0x071ca8d3, 0x205aa45a, 0x01585d97, 0x0116da06, 0x1a029b85, 0x139cb84c, 0x00003064};
    const uint32_t PINV30 = 0x1b799c77u;  // p^-1 mod 2^30
    int32_t f[9], g[9], d[9], e[9];
    {   // 8 x 32 bits -> 9 x 30 bits
        uint64_t acc = 0;
        int bits = 0, k = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            acc |= (uint64_t)a.l[i] << bits;
            bits += 32;
            while (bits >= 30) { g[k++] = (int32_t)(acc & (uint64_t)M30); acc >>= 30; bits -= 30; }
        }
        g[k] = (int32_t)acc;  // k == 8
    }
#pragma unroll
    for (int i = 0; i < 9; i++) { f[i] = PM[i]; d[i] = 0; e[i] = 0; }
    e[0] = 1;
    int32_t eta = -1;
    for (int iter = 0; iter < 24; iter++) {
        int32_t gz = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) gz |= g[i];
        if (gz == 0) break;
        // ---- 30 divsteps on the low words -> t = (u v; q r)
        uint32_t u = 1, v = 0, q = 0, r = 1, fl = (uint32_t)f[0] | ((uint32_t)f[1] << 30), gl = (uint32_t)g[0] | ((uint32_t)g[1] << 30);
        int i = 30;
        for (;;) {
            const int zeros = fq_ctz32(gl | (0xffffffffu << i));
            gl >>= zeros; u <<= zeros; v <<= zeros; eta -= zeros; i -= zeros;
            if (i == 0) break;
            if (eta < 0) {
                eta = -eta;
                uint32_t t;
                t = fl; fl = gl; gl = 0u - t;
                t = u; u = q; q = 0u - t;
                t = v; v = r; r = 0u - t;
            }
            gl += fl; q += u; r += v;  // both odd: g even again
        }
        // 32-bit factors: every product below is one widening multiply-add (mul.wide.s32), not a 64 x 64 product
        const int32_t tu = (int32_t)u, tv = (int32_t)v, tq = (int32_t)q, tr = (int32_t)r;
#define SIPP_W(a, b) ((int64_t)(a) * (int64_t)(b))
        {   // (d, e) <- t (d, e) / 2^30 mod p, kept in (-2p, p)
            const int32_t sd = d[8] >> 31, se = e[8] >> 31;
            int32_t md = ((int32_t)u & sd) + ((int32_t)v & se), me = ((int32_t)q & sd) + ((int32_t)r & se);
            int64_t cd = SIPP_W(tu, d[0]) + SIPP_W(tv, e[0]), ce = SIPP_W(tq, d[0]) + SIPP_W(tr, e[0]);
            md -= (int32_t)((PINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
            me -= (int32_t)((PINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
            cd += SIPP_W(PM[0], md); ce += SIPP_W(PM[0], me);
            cd >>= 30; ce >>= 30;
#pragma unroll
            for (int k = 1; k < 9; k++) {
                cd += SIPP_W(tu, d[k]) + SIPP_W(tv, e[k]) + SIPP_W(PM[k], md);
                ce += SIPP_W(tq, d[k]) + SIPP_W(tr, e[k]) + SIPP_W(PM[k], me);
                d[k - 1] = (int32_t)cd & M30; cd >>= 30;
                e[k - 1] = (int32_t)ce & M30; ce >>= 30;
            }
            d[8] = (int32_t)cd; e[8] = (int32_t)ce;
        }
        {   // (f, g) <- t (f, g) / 2^30 (exact)
            int64_t cf = SIPP_W(tu, f[0]) + SIPP_W(tv, g[0]), cg = SIPP_W(tq, f[0]) + SIPP_W(tr, g[0]);
            cf >>= 30; cg >>= 30;
#pragma unroll
            for (int k = 1; k < 9; k++) {
                cf += SIPP_W(tu, f[k]) + SIPP_W(tv, g[k]);
                cg += SIPP_W(tq, f[k]) + SIPP_W(tr, g[k]);
                f[k - 1] = (int32_t)cf & M30; cf >>= 30;
                g[k - 1] = (int32_t)cg & M30; cg >>= 30;
            }
            f[8] = (int32_t)cf; g[8] = (int32_t)cg;
        }
#undef SIPP_W
    }
    // a^-1 = sign(f) d, brought from (-2p, p) to [0, p)
    auto add_p = [&]() {
        int32_t c = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { const int32_t t = d[k] + PM[k] + c; d[k] = t & M30; c = t >> 30; }
        d[8] += PM[8] + c;
    };
    if (d[8] < 0) add_p();
    if (f[8] < 0) {
        int32_t c = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { const int32_t t = c - d[k]; d[k] = t & M30; c = t >> 30; }
        d[8] = c - d[8];
    }
    if (d[8] < 0) add_p();
    Fq x;
    {   // 9 x 30 bits -> 8 x 32 bits
        uint64_t acc = 0;
        int bits = 0, k = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            acc |= (uint64_t)(uint32_t)d[i] << bits;
            bits += 30;
            if (bits >= 32 && k < 8) { x.l[k++] = (uint32_t)acc; acc >>= 32; bits -= 32; }
        }
        if (k < 8) x.l[k] = (uint32_t)acc;
    }
    // the input is a Montgomery representative a = A R: a^-1 = A^-1 R^-1, and two products by R^2 give A^-1 R
    return fq_mul(fq_mul(x, fq_r2()), fq_r2());
}

}  // namespace sipp
