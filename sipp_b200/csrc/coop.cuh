// coop.cuh -- lane-cooperative Fq12 arithmetic: one Fq12 is spread over a group of 6 lanes, lane k holding the
// Fq2 coefficient g_k of w^k (w^6 = xi).  All sums of products use the lazy-reduction inner products of fqdot.cuh.
//
// The per-lane formulas are pure functions (`lane_*`) of values the caller has already fetched from the other lanes,
// so tests/hostcheck can run them on the CPU against the oracle; the `coop_*` wrappers (device only) add the warp
// shuffles.  5 groups fit in a warp (lanes 30, 31 idle).
//
//   product      c_k = sum_{i<=k} a_i b_{k-i} + xi * sum_{i>k} a_i b_{k-i+6}          6 Fq2 products per lane
//   sparse line  c_k = g_k l0 + g_{k-1} l1 [xi if k<1] + g_{k-3} l3 [xi if k<3]        3 Fq2 products per lane
//   cyclotomic   Granger-Scott on the pairs (g_k, g_{k+3}): one Fq2 "dot2" per lane
#pragma once
#include "fqdot.cuh"
#include "tower.cuh"

namespace sipp {

// sum_{i<M} a[i] * b[i] in Fq2 with two lazy inner products over Fq (M <= 6)
template <int M>
SIPP_HD Fq2 fq2_dot(const Fq2* a, const Fq2* b) {
    Fq x[2 * M], y[2 * M];
#pragma unroll
    for (int i = 0; i < M; i++) {
        x[2 * i] = a[i].c0;
        x[2 * i + 1] = a[i].c1;
        y[2 * i] = b[i].c0;
        y[2 * i + 1] = fq_neg(b[i].c1);
    }
    Fq2 r;
    r.c0 = fq_dot<2 * M>(x, y);  // a0 b0 - a1 b1
#pragma unroll
    for (int i = 0; i < M; i++) {
        y[2 * i] = b[i].c1;
        y[2 * i + 1] = b[i].c0;
    }
    r.c1 = fq_dot<2 * M>(x, y);  // a0 b1 + a1 b0
    return r;
}

// ---- per-lane formulas ---------------------------------------------------------------------------------------
// sparse line product for lane k: inputs g_k, g_{k-1 mod 6}, g_{k-3 mod 6} and the line coefficients already
// selected for this lane (l1s = k < 1 ? xi l1 : l1, l3s = k < 3 ? xi l3 : l3)
SIPP_HD Fq2 lane_sparse(const Fq2& gk, const Fq2& gkm1, const Fq2& gkm3, const Fq2& l0, const Fq2& l1s, const Fq2& l3s) {
    Fq2 a[3] = {gk, gkm1, gkm3};
    Fq2 b[3] = {l0, l1s, l3s};
    return fq2_dot<3>(a, b);
}

// half of a dense product for lane k: as[t] = a_{3h+t}, bs[t] = (3h+t <= k) ? b_{k-3h-t} : xi b_{k-3h-t+6}
SIPP_HD Fq2 lane_mul_half(const Fq2* as, const Fq2* bs) { return fq2_dot<3>(as, bs); }

// Granger-Scott: lane k < 3 computes a^2 + xi b^2, lane k >= 3 computes 2ab for the pair (a, b) = (g_{k mod 3}, g_{k mod 3 + 3})
SIPP_HD Fq2 lane_cyc_part(int k, const Fq2& a, const Fq2& b) {
    Fq2 xb = fq2_mul_xi(b);
    Fq2 u[2], v[2];
    u[0] = a;
    u[1] = (k < 3) ? xb : a;
    v[0] = (k < 3) ? a : b;
    v[1] = b;
    return fq2_dot<2>(u, v);
}
// t = the part this lane needs (already fetched: see coop_cyc_sqr), g = own coefficient
SIPP_HD Fq2 lane_cyc_finish(int k, const Fq2& t_in, const Fq2& g) {
    Fq2 t = (k == 1) ? fq2_mul_xi(t_in) : t_in;
    Fq2 sg = (k & 1) ? g : fq2_neg(g);
    return fq2_add(fq2_dbl(fq2_add(t, sg)), t);  // 3t + 2 sg
}
// index plumbing shared by the device wrappers and the host check
SIPP_HD void coop_mul_operand(int k, int i, int& j, bool& wrapped) {  // term a_i * b_j (times xi when wrapped)
    j = k - i;
    wrapped = j < 0;
    if (wrapped) j += 6;
}
SIPP_HD int coop_cyc_src(int k) { return (0x423150 >> (4 * k)) & 0xF; }  // 0<-0, 1<-5, 2<-1, 3<-3, 4<-2, 5<-4
SIPP_HD Fq2 lane_conj(int k, const Fq2& g) { return (k & 1) ? fq2_neg(g) : g; }
SIPP_HD Fq2 lane_frob(int k, const Fq2& g, int power) {
    Fq2 c = (power & 1) ? fq2_conj(g) : g;
    return fq2_mul_inl(c, frob_gamma(power, k));
}

#if defined(__CUDACC__)
// ---- warp plumbing ---------------------------------------------------------------------------------------------
struct Lane6 {
    int k;     // coefficient index 0..5 (6, 7 on the two idle lanes of a warp)
    int base;  // first lane of the group inside the warp
};
__device__ __forceinline__ Lane6 lane6_of_thread() {
    int lane = threadIdx.x & 31;
    Lane6 L;
    L.base = (lane / 6) * 6;
    L.k = lane - L.base + (lane >= 30 ? 6 : 0);
    return L;
}
__device__ __forceinline__ Fq2 shfl_fq2(const Fq2& v, int src) {
    Fq2 r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.c0.l[i] = __shfl_sync(0xffffffffu, v.c0.l[i], src);
        r.c1.l[i] = __shfl_sync(0xffffffffu, v.c1.l[i], src);
    }
    return r;
}
__device__ __forceinline__ Fq2 select_fq2(bool c, const Fq2& a, const Fq2& b) {
    Fq2 r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.c0.l[i] = c ? a.c0.l[i] : b.c0.l[i];
        r.c1.l[i] = c ? a.c1.l[i] : b.c1.l[i];
    }
    return r;
}

static __device__ __noinline__ Fq2 coop_mul(const Lane6& L, const Fq2& a, const Fq2& b) {
    const int k = L.k < 6 ? L.k : 0;
    Fq2 xb = fq2_mul_xi(b);
    Fq2 acc;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        Fq2 as[3], bs[3];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            int i = 3 * h + t;
            int j;
            bool wrapped;
            coop_mul_operand(k, i, j, wrapped);
            as[t] = shfl_fq2(a, L.base + i);
            Fq2 bj = shfl_fq2(b, L.base + j);
            Fq2 xbj = shfl_fq2(xb, L.base + j);
            bs[t] = select_fq2(wrapped, xbj, bj);
        }
        Fq2 part = lane_mul_half(as, bs);
        acc = h ? fq2_add(acc, part) : part;
    }
    return acc;
}
__device__ __forceinline__ Fq2 coop_sqr(const Lane6& L, const Fq2& a) { return coop_mul(L, a, a); }

// f * (l0 + l1 w + l3 w^3); `line` points at 5 Fq2 = {l0, l1, xi l1, l3, xi l3} in Montgomery limbs (80 words)
__device__ __forceinline__ Fq2 load_fq2_words(const uint32_t* p) {
    Fq2 r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3];
    r.c0.l[0] = v0.x; r.c0.l[1] = v0.y; r.c0.l[2] = v0.z; r.c0.l[3] = v0.w;
    r.c0.l[4] = v1.x; r.c0.l[5] = v1.y; r.c0.l[6] = v1.z; r.c0.l[7] = v1.w;
    r.c1.l[0] = v2.x; r.c1.l[1] = v2.y; r.c1.l[2] = v2.z; r.c1.l[3] = v2.w;
    r.c1.l[4] = v3.x; r.c1.l[5] = v3.y; r.c1.l[6] = v3.z; r.c1.l[7] = v3.w;
    return r;
}
__device__ __forceinline__ void store_fq2_words(uint32_t* p, const Fq2& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.c0.l[0], v.c0.l[1], v.c0.l[2], v.c0.l[3]);
    q[1] = make_uint4(v.c0.l[4], v.c0.l[5], v.c0.l[6], v.c0.l[7]);
    q[2] = make_uint4(v.c1.l[0], v.c1.l[1], v.c1.l[2], v.c1.l[3]);
    q[3] = make_uint4(v.c1.l[4], v.c1.l[5], v.c1.l[6], v.c1.l[7]);
}
// line-table writes are a stream (3.8 GB per 2^17 pairs, read back once by the accumulation kernel): evict-first stores, so that
// they do not push the line kernel's own stack frames out of L2 (ncu: 7.2 GB written per launch for a 3.8 GB table before)
__device__ __forceinline__ void stream_fq2_words(uint32_t* p, const Fq2& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    __stcs(q, make_uint4(v.c0.l[0], v.c0.l[1], v.c0.l[2], v.c0.l[3]));
    __stcs(q + 1, make_uint4(v.c0.l[4], v.c0.l[5], v.c0.l[6], v.c0.l[7]));
    __stcs(q + 2, make_uint4(v.c1.l[0], v.c1.l[1], v.c1.l[2], v.c1.l[3]));
    __stcs(q + 3, make_uint4(v.c1.l[4], v.c1.l[5], v.c1.l[6], v.c1.l[7]));
}
static __device__ __noinline__ Fq2 coop_sparse(const Lane6& L, const Fq2& g, const uint32_t* line) {
    const int k = L.k < 6 ? L.k : 0;
    Fq2 gm1 = shfl_fq2(g, L.base + (k + 5) % 6);
    Fq2 gm3 = shfl_fq2(g, L.base + (k + 3) % 6);
    Fq2 l0 = load_fq2_words(line);
    Fq2 l1s = load_fq2_words(line + (k < 1 ? 32 : 16));
    Fq2 l3s = load_fq2_words(line + (k < 3 ? 64 : 48));
    return lane_sparse(g, gm1, gm3, l0, l1s, l3s);
}

static __device__ __noinline__ Fq2 coop_cyc_sqr(const Lane6& L, const Fq2& g) {
    const int k = L.k < 6 ? L.k : 0;
    Fq2 partner = shfl_fq2(g, L.base + (k + 3) % 6);
    Fq2 a = (k < 3) ? g : partner;
    Fq2 b = (k < 3) ? partner : g;
    Fq2 part = lane_cyc_part(k, a, b);
    // lane k < 3 now holds t0 of pair k, lane k + 3 holds t1 of pair k.  Needed: 0<-t0(A)=0, 3<-t1(A)=3, 1<-t1(C)=5, 4<-t0(C)=2,
    // 2<-t0(B)=1, 5<-t1(B)=4
    Fq2 t = shfl_fq2(part, L.base + coop_cyc_src(k));
    return lane_cyc_finish(k, t, g);
}
__device__ __forceinline__ Fq2 coop_conj(const Lane6& L, const Fq2& g) { return lane_conj(L.k, g); }
static __device__ __noinline__ Fq2 coop_frob(const Lane6& L, const Fq2& g, int power) { return lane_frob(L.k < 6 ? L.k : 0, g, power); }

// a^x for the BN parameter x (a in the cyclotomic subgroup)
static __device__ __noinline__ Fq2 coop_cyc_exp_x(const Lane6& L, const Fq2& a) {
    Fq2 acc = a;
    const unsigned long long x = SIPP_BN_X;
#pragma unroll 1
    for (int b = 61; b >= 0; b--) {
        acc = coop_cyc_sqr(L, acc);
        if ((x >> b) & 1ull) acc = coop_mul(L, acc, a);
    }
    return acc;
}

// f^-1: n = f * conj(f) lies in Fq6 = span{w^0, w^2, w^4}; invert it with the cubic-extension adjugate (computed
// redundantly on every lane) and multiply back
static __device__ __noinline__ Fq2 coop_inv(const Lane6& L, const Fq2& f) {
    const int k = L.k < 6 ? L.k : 0;
    Fq2 cf = coop_conj(L, f);
    Fq2 n = coop_mul(L, f, cf);
    Fq2 n0 = shfl_fq2(n, L.base + 0), n1 = shfl_fq2(n, L.base + 2), n2 = shfl_fq2(n, L.base + 4);
    Fq2 t0 = fq2_sub(fq2_sqr(n0), fq2_mul_xi(fq2_mul(n1, n2)));
    Fq2 t1 = fq2_sub(fq2_mul_xi(fq2_sqr(n2)), fq2_mul(n0, n1));
    Fq2 t2 = fq2_sub(fq2_sqr(n1), fq2_mul(n0, n2));
    Fq2 d = fq2_add(fq2_mul(n0, t0), fq2_mul_xi(fq2_add(fq2_mul(n2, t1), fq2_mul(n1, t2))));
    d = fq2_inv(d);
    Fq2 mine = (k == 0) ? t0 : (k == 2) ? t1 : t2;
    Fq2 ninv = fq2_mul(mine, d);
    if (k & 1) ninv = fq2_zero();
    return coop_mul(L, cf, ninv);
}

// final exponentiation f^((p^12-1)/r) (+ the arkworks multiple if ark_norm), all six lanes of a group cooperating
static __device__ __noinline__ Fq2 coop_final_exp(const Lane6& L, const Fq2& f, bool ark_norm) {
    Fq2 t = coop_mul(L, coop_conj(L, f), coop_inv(L, f));
    Fq2 m = coop_mul(L, coop_frob(L, t, 2), t);
    Fq2 mx = coop_cyc_exp_x(L, m);
    Fq2 mx2 = coop_cyc_exp_x(L, mx);
    Fq2 mx3 = coop_cyc_exp_x(L, mx2);
    Fq2 y0 = coop_mul(L, coop_mul(L, coop_frob(L, m, 1), coop_frob(L, m, 2)), coop_frob(L, m, 3));
    Fq2 y1 = coop_conj(L, m);
    Fq2 y2 = coop_frob(L, mx2, 2);
    Fq2 y3 = coop_conj(L, coop_frob(L, mx, 1));
    Fq2 y4 = coop_conj(L, coop_mul(L, mx, coop_frob(L, mx2, 1)));
    Fq2 y5 = coop_conj(L, mx2);
    Fq2 y6 = coop_conj(L, coop_mul(L, mx3, coop_frob(L, mx3, 1)));
    Fq2 t0 = coop_mul(L, coop_mul(L, coop_cyc_sqr(L, y6), y4), y5);
    Fq2 t1 = coop_mul(L, coop_mul(L, y3, y5), t0);
    t0 = coop_mul(L, t0, y2);
    t1 = coop_cyc_sqr(L, coop_mul(L, coop_cyc_sqr(L, t1), t0));
    t0 = coop_mul(L, t1, y1);
    t1 = coop_mul(L, t1, y0);
    Fq2 out = coop_mul(L, coop_cyc_sqr(L, t0), t1);
    if (ark_norm) {
        Fq2 a = coop_cyc_exp_x(L, out), b = coop_cyc_exp_x(L, a), c = coop_cyc_exp_x(L, b);
        Fq2 b3 = coop_mul(L, coop_cyc_sqr(L, b), b);
        Fq2 c6 = coop_cyc_sqr(L, coop_mul(L, coop_cyc_sqr(L, c), c));
        out = coop_cyc_sqr(L, coop_mul(L, coop_mul(L, a, b3), c6));
    }
    return out;
}
#endif  // __CUDACC__

}  // namespace sipp
