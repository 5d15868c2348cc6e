// fqdot.cuh -- lazy-reduction inner products in Fq:  fq_dot<N>(a, b) = sum_i a[i] * b[i]  (Montgomery form).
//
// A multi-product CIOS: for every 32-bit limb t of the b operands the rows of ALL N products are accumulated into
// the same two-accumulator pair as fq_mul (fq.cuh), then ONE Montgomery reduction step is applied.  N products cost
// 64 N + 72 multiply-adds instead of 136 N, which is what makes the Fq2 / Fq12 sums of products (sparse line
// product, cooperative Fq12 multiplication) cheap.  The running value stays below (N + 1) p, so the high
// accumulator carries a ninth limb; the final value is below (N/4 + 1) p and is brought to [0, p) with at most two
// conditional subtractions for N <= 12.
#pragma once
#include "fq.cuh"

namespace sipp {

#if defined(__CUDA_ARCH__)
// acc[0..7] += (a0, a2, a4, a6) * b ; carry into acc8
__device__ __forceinline__ void row_mad9(uint32_t* acc, uint32_t& acc8, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    row_mad_carry(acc, acc8, a0, a2, a4, a6, b);
}
// x[0..7] += (a0, a2, a4, a6) * b ; carry into (y7, y8)
__device__ __forceinline__ void row_mad_carry2(uint32_t* x, uint32_t& y7, uint32_t& y8, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
        "addc.cc.u32 %8, %8, 0;\n\t"
        "addc.u32 %9, %9, 0;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(y7), "+r"(y8)
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}
// role swap with a ninth limb:  x0 += y[1] (carry c);  y[0..7] <- (y >> 64) + x8 * 2^224 + (a1,a3,a5,a7) * b + c ; y8 <- carry
__device__ __forceinline__ void row_shift_mad9(uint32_t& x0, uint32_t x8, uint32_t* y, uint32_t& y8, uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7,
                                               uint32_t b) {
    asm("add.cc.u32 %9, %9, %1;\n\t"
        "madc.lo.cc.u32 %0, %11, %15, %2;\n\t"
        "madc.hi.cc.u32 %1, %11, %15, %3;\n\t"
        "madc.lo.cc.u32 %2, %12, %15, %4;\n\t"
        "madc.hi.cc.u32 %3, %12, %15, %5;\n\t"
        "madc.lo.cc.u32 %4, %13, %15, %6;\n\t"
        "madc.hi.cc.u32 %5, %13, %15, %7;\n\t"
        "madc.lo.cc.u32 %6, %14, %15, 0;\n\t"
        "madc.hi.cc.u32 %7, %14, %15, %10;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7]), "=r"(y8), "+r"(x0)
        : "r"(x8), "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
}
// r[0..7] = (x >> 32) + y[0..7]   (x[0] == 0; the ninth limb of y is zero by the bound)
#else
inline void row_mad9(uint32_t* acc, uint32_t& acc8, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    row_mad_carry(acc, acc8, a0, a2, a4, a6, b);
}
inline void row_mad_carry2(uint32_t* x, uint32_t& y7, uint32_t& y8, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    uint32_t c = 0;
    row_mad_carry(x, c, a0, a2, a4, a6, b);
    uint64_t s = (uint64_t)y7 + c;
    y7 = (uint32_t)s;
    y8 += (uint32_t)(s >> 32);
}
inline void row_shift_mad9(uint32_t& x0, uint32_t x8, uint32_t* y, uint32_t& y8, uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7, uint32_t b) {
    uint64_t s = (uint64_t)x0 + y[1];
    x0 = (uint32_t)s;
    uint32_t base[8] = {y[2], y[3], y[4], y[5], y[6], y[7], 0, x8};
    y8 = row_mad_host(y, base, a1, a3, a5, a7, b, (uint32_t)(s >> 32));
}
#endif

// accumulator pair; `e8` / `o8` are the ninth limbs (only one of them is live at a time)
struct DotAcc {
    uint32_t e[8], o[8];
    uint32_t e8, o8;
};

template <int N>
SIPP_HD void dot_reduce_step(uint32_t* x, uint32_t* y, uint32_t& y8) {
    uint32_t m = x[0] * SIPP_PINV;
    row_mad9(y, y8, SIPP_P1, SIPP_P3, SIPP_P5, SIPP_P7, m);
    row_mad_carry2(x, y[7], y8, SIPP_P0, SIPP_P2, SIPP_P4, SIPP_P6, m);
}

// first limb: x = aligned accumulator, y = high accumulator, both fresh
template <int N>
SIPP_HD void dot_first_step(uint32_t* x, uint32_t* y, uint32_t& y8, const Fq* a, const Fq* b) {
    row_mul(y, a[0].l[1], a[0].l[3], a[0].l[5], a[0].l[7], b[0].l[0]);
    row_mul(x, a[0].l[0], a[0].l[2], a[0].l[4], a[0].l[6], b[0].l[0]);
    y8 = 0;
#pragma unroll
    for (int i = 1; i < N; i++) {
        row_mad9(y, y8, a[i].l[1], a[i].l[3], a[i].l[5], a[i].l[7], b[i].l[0]);
        row_mad_carry2(x, y[7], y8, a[i].l[0], a[i].l[2], a[i].l[4], a[i].l[6], b[i].l[0]);
    }
    dot_reduce_step<N>(x, y, y8);
}
// limb t >= 1: x = NEW aligned accumulator (old high, ninth limb x8), y = NEW high accumulator (old aligned)
template <int N>
SIPP_HD void dot_step(uint32_t* x, uint32_t x8, uint32_t* y, uint32_t& y8, const Fq* a, const Fq* b, int t) {
    row_shift_mad9(x[0], x8, y, y8, a[0].l[1], a[0].l[3], a[0].l[5], a[0].l[7], b[0].l[t]);
    row_mad_carry2(x, y[7], y8, a[0].l[0], a[0].l[2], a[0].l[4], a[0].l[6], b[0].l[t]);
#pragma unroll
    for (int i = 1; i < N; i++) {
        row_mad9(y, y8, a[i].l[1], a[i].l[3], a[i].l[5], a[i].l[7], b[i].l[t]);
        row_mad_carry2(x, y[7], y8, a[i].l[0], a[i].l[2], a[i].l[4], a[i].l[6], b[i].l[t]);
    }
    dot_reduce_step<N>(x, y, y8);
}

SIPP_HD void fq_cond_sub_2p(uint32_t* r) {  // r in [0, 4p) -> [0, 2p)
    const uint32_t P2[8] = {0xb0f9fa8eu, 0x7841182du, 0xd0e3951au, 0x2f02d522u, 0x0302b0bbu, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u};  // 2p
    uint32_t t[8];
    uint32_t borrow = sub8(t, r, P2);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : t[i];
}

// sum_i a[i] * b[i], all operands canonical Montgomery residues, N <= 12; result canonical
template <int N>
SIPP_HD Fq fq_dot(const Fq* a, const Fq* b) {
    static_assert(N >= 1 && N <= 12, "fq_dot supports 1..12 products");
    DotAcc s;
    dot_first_step<N>(s.e, s.o, s.o8, a, b);
    dot_step<N>(s.o, s.o8, s.e, s.e8, a, b, 1);
    dot_step<N>(s.e, s.e8, s.o, s.o8, a, b, 2);
    dot_step<N>(s.o, s.o8, s.e, s.e8, a, b, 3);
    dot_step<N>(s.e, s.e8, s.o, s.o8, a, b, 4);
    dot_step<N>(s.o, s.o8, s.e, s.e8, a, b, 5);
    dot_step<N>(s.e, s.e8, s.o, s.o8, a, b, 6);
    dot_step<N>(s.o, s.o8, s.e, s.e8, a, b, 7);
    // aligned accumulator = o (o[0] == 0), high = e (e8 == 0 because the value is < 4p < 2^256)
    Fq r;
    merge_acc(r.l, s.o, s.e);
    if (N > 4) fq_cond_sub_2p(r.l);  // < (N/4 + 1) p <= 4p
    if (N > 1) fq_cond_sub_p(r.l);
    fq_cond_sub_p(r.l);
    return r;
}

}  // namespace sipp
