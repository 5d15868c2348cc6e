// engine.cuh -- lane-parallel executor for the straight-line programs of line_programs.h (tools/gen_line_programs.py).
//
// A GROUP of 16 lanes owns one (P, Q) pair and a slot file of Fq values in shared memory.  A program is a list of
// levels; in a level every lane executes ONE instruction of the same kind (so the warp never diverges):
//     MUL  dst = s0 * s1 (+/-) s2 * s3      one lazy-reduction inner product, fq_dot<2>
//     LIN  dst = c0 s0 + c1 s1 + c2 s2 + c3 s3   small signed coefficients, one reduction (fq_lincomb4)
// All reads of a level happen before its writes.  This shortens the dependent chain of the G2 side of the Miller loop
// (the reference's per-pair `pairing`, /root/reference/src/prover_native.rs:20) from ~28 sequential Fq2 products per
// doubling step to two MUL levels, which is what the latency-bound rounds of the prover need.
//
// The per-lane evaluation (`lp_eval`) and the Miller schedule (`lp_miller`) are plain functions over a "machine"
// interface so tests/hostcheck can run the very same tables on the CPU against the oracle.
#pragma once
#include "fqdot.cuh"
#include "tower.cuh"
#include "line_programs.h"
#include "fold_plan.h"

namespace sipp {

struct LpIns {
    uint32_t w0, w1, w2, w3;
};

// sum_k c_k s_k mod p for canonical s_k and |c_0| + ... + |c_3| <= 2047; result canonical
SIPP_HD Fq fq_lincomb4(const Fq& s0, const Fq& s1, const Fq& s2, const Fq& s3, int c0, int c1, int c2, int c3) {
    // R = sum c_k s_k + 2048 p lies in (p, 4095 p): nine 32-bit limbs, non-negative
    const uint32_t KP[9] = {0xe7ea3800u, 0x0460b6c3u, 0x8e5469e1u, 0x0b548b43u, 0x0ac2ecbcu, 0x822db40cu, 0x8d014dc2u, 0x22739709u, 0x00000183u};
    uint32_t r[9];
    int64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int64_t v = carry + (int64_t)KP[i];
        v += (int64_t)c0 * (int64_t)(uint64_t)s0.l[i];
        v += (int64_t)c1 * (int64_t)(uint64_t)s1.l[i];
        v += (int64_t)c2 * (int64_t)(uint64_t)s2.l[i];
        v += (int64_t)c3 * (int64_t)(uint64_t)s3.l[i];
        r[i] = (uint32_t)v;
        carry = v >> 32;
    }
    r[8] = (uint32_t)(carry + (int64_t)KP[8]);
    // quotient estimate from the top 64 bits: q <= floor(R / p) <= q + 2  (divisor rounded up to (p >> 224) + 1)
    const uint64_t top = ((uint64_t)r[8] << 32) | r[7];
#if defined(__CUDA_ARCH__)
    const uint32_t q = (uint32_t)__umul64hi(top, 0x000000054a474622ull);
#else
    const uint32_t q = (uint32_t)(((unsigned __int128)top * 0x000000054a474622ull) >> 64);
#endif
    // R -= q p  (fits eight limbs afterwards: < 3p)
    Fq out;
    int64_t bc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t v = bc + (int64_t)r[i] - (int64_t)((uint64_t)q * fq_p_limb(i));
        out.l[i] = (uint32_t)v;
        bc = v >> 32;
    }
    fq_cond_sub_p(out.l);
    fq_cond_sub_p(out.l);
    return out;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ Fq lp_load(const uint32_t* slots, int s) {
    const uint4* q = reinterpret_cast<const uint4*>(slots + 8 * s);
    const uint4 a = q[0], b = q[1];
    return Fq{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ void lp_store(uint32_t* slots, int s, const Fq& v) {
    uint4* q = reinterpret_cast<uint4*>(slots + 8 * s);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
#else
inline Fq lp_load(const uint32_t* slots, int s) {
    Fq r;
    for (int i = 0; i < 8; i++) r.l[i] = slots[8 * s + i];
    return r;
}
inline void lp_store(uint32_t* slots, int s, const Fq& v) {
    for (int i = 0; i < 8; i++) slots[8 * s + i] = v.l[i];
}
#endif

SIPP_HD int lp_dst(const LpIns& ins) { return (int)(ins.w0 & 255u); }

// one instruction on one lane; `type` is uniform over the level (0 = MUL, 1 = LIN, 2 = MUL1)
SIPP_HD Fq lp_eval(int type, const LpIns& ins, const uint32_t* slots) {
    const Fq a = lp_load(slots, (ins.w0 >> 8) & 255u), b = lp_load(slots, (ins.w0 >> 16) & 255u);
    const Fq c = lp_load(slots, ins.w0 >> 24);
    Fq d = lp_load(slots, ins.w1 & 255u);
    if (type == 0) {
        const bool neg = (ins.w1 >> 8) & 1u;
        const Fq nd = fq_neg(d);
#pragma unroll
        for (int i = 0; i < 8; i++) d.l[i] = neg ? nd.l[i] : d.l[i];
        const Fq x[2] = {a, c}, y[2] = {b, d};
        return fq_dot<2>(x, y);
    }
    if (type == 2) return fq_mul(a, b);
    return fq_lincomb4(a, b, c, d, (int)(int16_t)(ins.w2 & 0xffffu), (int)(int16_t)(ins.w2 >> 16), (int)(int16_t)(ins.w3 & 0xffffu),
                       (int)(int16_t)(ins.w3 >> 16));
}

// The Miller-loop schedule over a machine M:  M.run(first_level, n_levels) executes a program, M.emit(step) consumes the
// ten OUT slots (one line).  64 tangents, the chords of the signed digits of 6x+2, two Frobenius chords: 91 lines.
template <class M>
SIPP_HD void lp_miller(M& mach) {
    const unsigned long long plus = SIPP_ATE_PLUS_MASK, minus = SIPP_ATE_MINUS_MASK;
    mach.run(SIPP_LP_SETUP_FIRST, SIPP_LP_SETUP_LEVELS);
    int step = 0;
    for (int i = 63; i >= 0; i--) {
        mach.run(SIPP_LP_DBL_FIRST, SIPP_LP_DBL_LEVELS);
        mach.emit(step++);
        if ((plus >> i) & 1ull) {
            mach.run(SIPP_LP_ADD_P_FIRST, SIPP_LP_ADD_P_LEVELS);
            mach.emit(step++);
        } else if ((minus >> i) & 1ull) {
            mach.run(SIPP_LP_ADD_M_FIRST, SIPP_LP_ADD_M_LEVELS);
            mach.emit(step++);
        }
    }
    mach.run(SIPP_LP_ADD_Q1_FIRST, SIPP_LP_ADD_Q1_LEVELS);
    mach.emit(step++);
    mach.run(SIPP_LP_ADD_Q2_FIRST, SIPP_LP_ADD_Q2_LEVELS);
    mach.emit(step++);
}

// [k]Q for the fold kernels: k given as signed binary digits (bit i of plus / minus set = digit +1 / -1), most significant
// digit at `top` (must be +1: the caller negates Q and swaps the masks otherwise).  T starts as Q; G2 uses the line
// programs' point arithmetic (their line outputs are simply not consumed), G1 the Fq programs.
template <class M>
SIPP_HD void lp_scalar_mul(M& mach, bool g2, const uint32_t* plus, const uint32_t* minus, int top) {
    for (int i = top - 1; i >= 0; i--) {
        const uint32_t bit = 1u << (i & 31);
        const bool dp = (plus[i >> 5] & bit) != 0, dm = (minus[i >> 5] & bit) != 0;
        if (g2) {
            mach.run(SIPP_LP_DBL_FIRST, SIPP_LP_DBL_LEVELS);
            if (dp) mach.run(SIPP_LP_ADD_P_FIRST, SIPP_LP_ADD_P_LEVELS);
            else if (dm) mach.run(SIPP_LP_ADD_M_FIRST, SIPP_LP_ADD_M_LEVELS);
        } else {
            mach.run(SIPP_LP_DBL1_FIRST, SIPP_LP_DBL1_LEVELS);
            if (dp) mach.run(SIPP_LP_ADD1_P_FIRST, SIPP_LP_ADD1_P_LEVELS);
            else if (dm) mach.run(SIPP_LP_ADD1_M_FIRST, SIPP_LP_ADD1_M_LEVELS);
        }
    }
}

// digits of one fold component as lp_scalar_mul wants them
struct FoldDigits {
    uint32_t plus[SIPP_FOLD_MASK_WORDS], minus[SIPP_FOLD_MASK_WORDS];
    int top;    // index of the most significant non-zero digit, -1 if the sub-scalar is zero
    bool flip;  // that digit is -1: work with -Q and negated digits
};
SIPP_HD FoldDigits fold_digits(const FoldComp& c, int bits) {
    FoldDigits d;
    d.top = -1;
    for (int i = bits - 1; i >= 0; i--) {
        const uint32_t bit = 1u << (i & 31);
        if ((c.plus[i >> 5] | c.minus[i >> 5]) & bit) { d.top = i; break; }
    }
    d.flip = d.top >= 0 && (c.minus[d.top >> 5] >> (d.top & 31)) & 1u;
    for (int w = 0; w < SIPP_FOLD_MASK_WORDS; w++) {
        d.plus[w] = d.flip ? c.minus[w] : c.plus[w];
        d.minus[w] = d.flip ? c.plus[w] : c.minus[w];
    }
    return d;
}
// the fixed slots every group starts from (P, Q in Montgomery limbs)
SIPP_HD void lp_fill_fixed(uint32_t* slots, const Fq& xp, const Fq& yp, const Fq2& qx, const Fq2& qy) {
    const Fq2 xi_inv = Fq2 SIPP_XI_INV_INIT;
    const Fq2 g12 = frob_gamma(1, 2), g13 = frob_gamma(1, 3);
    lp_store(slots, SIPP_LP_SLOT_ZERO, fq_zero());
    lp_store(slots, SIPP_LP_SLOT_XP, xp);
    lp_store(slots, SIPP_LP_SLOT_YP, yp);
    lp_store(slots, SIPP_LP_SLOT_QX0, qx.c0);
    lp_store(slots, SIPP_LP_SLOT_QX1, qx.c1);
    lp_store(slots, SIPP_LP_SLOT_QY0, qy.c0);
    lp_store(slots, SIPP_LP_SLOT_QY1, qy.c1);
    lp_store(slots, SIPP_LP_SLOT_XIINV0, xi_inv.c0);
    lp_store(slots, SIPP_LP_SLOT_XIINV1, xi_inv.c1);
    lp_store(slots, SIPP_LP_SLOT_G12_0, g12.c0);
    lp_store(slots, SIPP_LP_SLOT_G12_1, g12.c1);
    lp_store(slots, SIPP_LP_SLOT_G13_0, g13.c0);
    lp_store(slots, SIPP_LP_SLOT_G13_1, g13.c1);
}

}  // namespace sipp
