// fq2h.cuh -- an Fq2 value spread over TWO adjacent lanes: the even lane holds c0, the odd lane c1 (device only).
//
// The lane-split fold kernels walk a G2 scalar multiplication as one dependent chain of Fq2 products per thread; an Fq2
// product is three dependent Fq products on one thread (~2,600 cycles on a lone warp).  With the two halves on two lanes every
// lane does ONE lazy inner product (c0 = a0 b0 - a1 b1 or c1 = a0 b1 + a1 b0: fq_dot<2>, ~1,300 cycles) after fetching the
// partner's halves with shuffles, and a squaring is ONE Fq product per lane ((a0 + a1)(a0 - a1) | 2 a0 a1).  Additions,
// subtractions and negations are component-wise and need no exchange.  The type plugs into the curve templates of curve.cuh
// (f_add, f_mul, ...), so the Jacobian formulas and the NAF loops are the same code as for one thread per value.
// Both lanes of a pair must execute every call together (all branches on values are pair-uniform: f_is_zero combines both halves).
#pragma once
#include "fqdot.cuh"
#include "tower.cuh"

#if defined(__CUDACC__)
namespace sipp {

struct Fq2H {
    Fq v;
};
__device__ __forceinline__ bool h_odd() { return (threadIdx.x & 1) != 0; }
// shuffles name only the two lanes of the pair: pairs of one warp may sit in different branches (an identity point, a doubling)
__device__ __forceinline__ unsigned h_mask() { return 3u << (threadIdx.x & 30u); }
__device__ __forceinline__ Fq h_partner(const Fq& a) {
    const unsigned m = h_mask();
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(m, a.l[i], 1);
    return r;
}
__device__ __forceinline__ Fq h_select(bool c, const Fq& a, const Fq& b) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
__device__ __forceinline__ Fq2H h_split(const Fq2& a) { return Fq2H{h_odd() ? a.c1 : a.c0}; }
// the full value on both lanes
__device__ __forceinline__ Fq2 h_join(const Fq2H& a) {
    const Fq p = h_partner(a.v);
    return h_odd() ? Fq2{p, a.v} : Fq2{a.v, p};
}

__device__ __forceinline__ Fq2H f_add(const Fq2H& a, const Fq2H& b) { return Fq2H{fq_add(a.v, b.v)}; }
__device__ __forceinline__ Fq2H f_sub(const Fq2H& a, const Fq2H& b) { return Fq2H{fq_sub(a.v, b.v)}; }
__device__ __forceinline__ Fq2H f_dbl(const Fq2H& a) { return Fq2H{fq_dbl(a.v)}; }
__device__ __forceinline__ Fq2H f_neg(const Fq2H& a) { return Fq2H{fq_neg(a.v)}; }
__device__ __noinline__ Fq2H f_mul(const Fq2H& a, const Fq2H& b) {
    const bool odd = h_odd();
    const Fq pa = h_partner(a.v), pb = h_partner(b.v);
    // even: a0 b0 + a1 (-b1);   odd (own = a1, b1; partner = a0, b0): a1 b0 + a0 b1
    const Fq x[2] = {a.v, pa};
    const Fq y[2] = {h_select(odd, pb, b.v), h_select(odd, b.v, fq_neg(pb))};
    return Fq2H{fq_dot<2>(x, y)};
}
__device__ __noinline__ Fq2H f_sqr(const Fq2H& a) {
    const bool odd = h_odd();
    const Fq pa = h_partner(a.v);
    // even: (a0 + a1)(a0 - a1);   odd: (2 a1) a0
    const Fq u = h_select(odd, fq_dbl(a.v), fq_add(a.v, pa));
    const Fq w = h_select(odd, pa, fq_sub(a.v, pa));
    return Fq2H{fq_mul(u, w)};
}
__device__ __forceinline__ bool f_is_zero(const Fq2H& a) {
    const bool z = fq_is_zero(a.v);
    return __shfl_xor_sync(h_mask(), z ? 1 : 0, 1) != 0 && z;
}
__device__ __forceinline__ void f_set_one(Fq2H& a) { a.v = h_odd() ? fq_zero() : fq_one(); }
__device__ __forceinline__ void f_set_zero(Fq2H& a) { a.v = fq_zero(); }
// 1 / (a0 + a1 u) = (a0 - a1 u) / (a0^2 + a1^2): both lanes invert the same norm
__device__ __noinline__ Fq2H f_inv(const Fq2H& a) {
    const Fq s = fq_sqr(a.v);
    const Fq n = fq_inv(fq_add(s, h_partner(s)));
    const Fq r = fq_mul(a.v, n);
    return Fq2H{h_odd() ? fq_neg(r) : r};
}

}  // namespace sipp
#endif
