// engine12.cuh -- 32-lane executor for the Fq12 programs of fq12_programs.h (tools/gen_fq12_programs.py) and the final
// exponentiation written on top of it.
//
// One warp = one machine.  A level is DOTn (n = 2, 4, 6: one fq_dot<n> per lane), LIN (fq_lincomb4) or INV; operands
// are registers of 24 Fq slots in shared memory named at run time (banks D, A, B) plus a bank of globals (Frobenius
// constants, scratch).  A dense Fq12 product is one DOT6 level on 24 lanes where the 6-lane version (coop.cuh) runs four
// dependent fq_dot<6> per lane; the final exponentiation -- one per product, a strictly serial chain -- gets ~3x shorter.
//
// Replaces the final-exponentiation half of `pairing` (/root/reference/src/prover_native.rs:20, verifier_native.rs:80).
// Everything here compiles for the host too (tests/hostcheck emulates the 32 lanes sequentially).
#pragma once
#include "engine.cuh"
#include "fq12_programs.h"

namespace sipp {

struct F12Ins {
    uint32_t w[4];
};
SIPP_HD uint32_t f12_byte(const F12Ins& ins, int i) { return (ins.w[i >> 2] >> (8 * (i & 3))) & 255u; }
SIPP_HD int f12_slot(uint32_t byte, const int* base) {  // selects, not an indexed load: `base` stays in registers
    const uint32_t bank = byte >> 6;
    const int bs = bank == 0 ? base[0] : (bank == 1 ? base[1] : (bank == 2 ? base[2] : base[3]));
    return bs + (int)(byte & 63u);
}
SIPP_HD int f12_reg_base(int r) { return SIPP_F12_GLOBAL_SLOTS + SIPP_F12_REG_SLOTS * r; }

template <int N>
SIPP_HD Fq f12_dot(const F12Ins& ins, const uint32_t* slots, const int* base) {
    Fq a[N], b[N];
    const uint32_t mask = f12_byte(ins, 13);
#pragma unroll
    for (int t = 0; t < N; t++) {
        a[t] = lp_load(slots, f12_slot(f12_byte(ins, 1 + 2 * t), base));
        const Fq v = lp_load(slots, f12_slot(f12_byte(ins, 2 + 2 * t), base));
        const Fq nv = fq_neg(v);
        const bool neg = (mask >> t) & 1u;
#pragma unroll
        for (int i = 0; i < 8; i++) b[t].l[i] = neg ? nv.l[i] : v.l[i];
    }
    return fq_dot<N>(a, b);
}

// one instruction on one lane; `type` is uniform over the level.  Returns false when the lane has nothing to write.
SIPP_HD bool f12_eval(int type, const F12Ins& ins, const uint32_t* slots, const int* base, Fq& r) {
    switch (type) {
        case 0: r = f12_dot<2>(ins, slots, base); return true;
        case 2: r = f12_dot<4>(ins, slots, base); return true;
        case 3: r = f12_dot<6>(ins, slots, base); return true;
        case 1: {
            const Fq s0 = lp_load(slots, f12_slot(f12_byte(ins, 1), base)), s1 = lp_load(slots, f12_slot(f12_byte(ins, 2), base));
            const Fq s2 = lp_load(slots, f12_slot(f12_byte(ins, 3), base)), s3 = lp_load(slots, f12_slot(f12_byte(ins, 4), base));
            r = fq_lincomb4(s0, s1, s2, s3, (int)(int16_t)(ins.w[2] & 0xffffu), (int)(int16_t)(ins.w[2] >> 16), (int)(int16_t)(ins.w[3] & 0xffffu),
                            (int)(int16_t)(ins.w[3] >> 16));
            return true;
        }
        default:
            if (f12_byte(ins, 14)) return false;
            r = fq_inv(lp_load(slots, f12_slot(f12_byte(ins, 1), base)));
            return true;
    }
}

// Machine interface used below:  mc.run(first_level, n_levels, regD, regA, regB)
#define F12_OP3(mc, NAME, d, a, b) (mc).run(SIPP_F12_##NAME##_FIRST, SIPP_F12_##NAME##_LEVELS, d, a, b)
#define F12_OP2(mc, NAME, d, a) (mc).run(SIPP_F12_##NAME##_FIRST, SIPP_F12_##NAME##_LEVELS, d, a, a)

// d = a^x for the BN parameter x (a in the cyclotomic subgroup); d != a.  The 62 squarings and 27 products are chain links
// (CSQRX / MUL12X): the xi-multiples an operation reads are left behind by the closing LIN level of the previous one (and
// once for the base by XI6), so every link is two levels -- products, closing LIN -- instead of three.  The three
// exponentiations are 60 % of the levels of a final exponentiation, and a LIN level costs a third of a squaring's products.
template <class M>
SIPP_HD void f12_exp_x(M& mc, int d, int a) {
    F12_OP2(mc, XI6, a, a);
    F12_OP2(mc, COPYX, d, a);
    const unsigned long long x = SIPP_BN_X;
    for (int b = 61; b >= 0; b--) {
        F12_OP2(mc, CSQRX, d, d);
        if ((x >> b) & 1ull) F12_OP3(mc, MUL12X, d, d, a);
    }
}

// register 0 <- a^k for a 256-bit exponent k (8 x u32, little endian), a = register 1 on entry; registers 2, 3 become a^2, a^3.
// GENERIC arithmetic (no cyclotomic shortcut: this is the verifier's Z_L.pow(x), /root/reference/src/verifier_native.rs:59-61,
// on proof elements that need not lie in GT): fixed 2-bit windows, every squaring and product a MUL12Y chain link (two levels).
// Returns false when k == 0 (register 0 untouched: the caller writes 1).  On exit register 0's xi slots are valid.
template <class M>
SIPP_HD bool f12_pow_w2(M& mc, const uint32_t* k) {
    F12_OP2(mc, XI6, 1, 1);
    F12_OP3(mc, MUL12Y, 2, 1, 1);
    F12_OP3(mc, MUL12Y, 3, 2, 1);
    bool started = false;
    for (int i = 127; i >= 0; i--) {
        const int d = (int)((k[i >> 4] >> (2 * (i & 15))) & 3u);
        if (started) {
            F12_OP3(mc, MUL12Y, 0, 0, 0);
            F12_OP3(mc, MUL12Y, 0, 0, 0);
            if (d) F12_OP3(mc, MUL12Y, 0, 0, d);
        } else if (d) {
            F12_OP2(mc, COPYX, 0, d);
            started = true;
        }
    }
    return started;
}

// register 0 <- (+-frob^comp(register 1))^k for one recoded sub-scalar k (fold_plan.h GtPlan: NAF digits, sign), register 1 a
// CYCLOTOMIC element (a pairing value); registers 1, 2 are overwritten.  The pairing-matrix tail's GT exponentiation
// e(A_i + x A_j, .) = e(A_i, .) e(A_j, .)^x  (/root/reference/src/prover_native.rs:60-69 carried over to GT): Frobenius images are
// one DOT2 level, inverses are conjugates, squarings / products are CSQRX / MUL12X chain links.  Returns false when k == 0
// (register 0 untouched: the caller writes 1).
template <class M>
SIPP_HD bool f12_gt_pow_comp(M& mc, int comp, const FoldComp& c, int bits) {
    const FoldDigits d = fold_digits(c, bits);
    if (d.top < 0) return false;
    if (comp == 1) F12_OP2(mc, FROB1, 1, 1);
    else if (comp == 2) F12_OP2(mc, FROB2, 1, 1);
    else if (comp == 3) F12_OP2(mc, FROB3, 1, 1);
    if ((c.neg != 0) != d.flip) F12_OP2(mc, CONJ, 1, 1);
    F12_OP2(mc, XI6, 1, 1);
    F12_OP2(mc, CONJ, 2, 1);
    F12_OP2(mc, XI6, 2, 2);
    F12_OP2(mc, COPYX, 0, 1);
    for (int i = d.top - 1; i >= 0; i--) {
        F12_OP2(mc, CSQRX, 0, 0);
        const uint32_t bit = 1u << (i & 31);
        if (d.plus[i >> 5] & bit) F12_OP3(mc, MUL12X, 0, 0, 1);
        else if (d.minus[i >> 5] & bit) F12_OP3(mc, MUL12X, 0, 0, 2);
    }
    return true;
}

#define SIPP_F12_FE_REGS 11
// register 0 holds f on entry; returns the register that holds f^((p^12-1)/r) (+ the arkworks multiple if ark_norm).
// Same formulas, in the same order, as coop_final_exp (coop.cuh) / final_exponentiation (pairing.cuh).
template <class M>
SIPP_HD int f12_final_exp(M& mc, bool ark_norm) {
    enum { F = 0, T = 1, MM = 2, MX = 3, MX2 = 4, MX3 = 5, Y0 = 6, A = 7, B = 8, C = 9, S = 10 };
    F12_OP3(mc, INV12, A, F, S);        // f^-1
    F12_OP2(mc, CONJ, B, F);
    F12_OP3(mc, MUL12, T, B, A);        // t = conj(f) f^-1
    F12_OP2(mc, FROB2, A, T);
    F12_OP3(mc, MUL12, MM, A, T);       // m = frob2(t) t
    f12_exp_x(mc, MX, MM);
    f12_exp_x(mc, MX2, MX);
    f12_exp_x(mc, MX3, MX2);
    F12_OP2(mc, FROB1, A, MM);
    F12_OP2(mc, FROB2, B, MM);
    F12_OP3(mc, MUL12, A, A, B);
    F12_OP2(mc, FROB3, B, MM);
    F12_OP3(mc, MUL12, Y0, A, B);       // y0 = frob1(m) frob2(m) frob3(m)
    F12_OP2(mc, FROB1, A, MX3);
    F12_OP3(mc, MUL12, A, MX3, A);
    F12_OP2(mc, CONJ, A, A);            // y6 = conj(mx3 frob1(mx3))
    F12_OP2(mc, CSQR, A, A);            // y6^2
    F12_OP2(mc, FROB1, B, MX2);
    F12_OP3(mc, MUL12, B, MX, B);
    F12_OP2(mc, CONJ, B, B);            // y4 = conj(mx frob1(mx2))
    F12_OP3(mc, MUL12, A, A, B);
    F12_OP2(mc, CONJ, B, MX2);          // y5 = conj(mx2)
    F12_OP3(mc, MUL12, A, A, B);        // t0 = y6^2 y4 y5
    F12_OP2(mc, FROB1, C, MX);
    F12_OP2(mc, CONJ, C, C);            // y3 = conj(frob1(mx))
    F12_OP3(mc, MUL12, C, C, B);
    F12_OP3(mc, MUL12, C, C, A);        // t1 = y3 y5 t0
    F12_OP2(mc, FROB2, B, MX2);         // y2 = frob2(mx2)
    F12_OP3(mc, MUL12, A, A, B);        // t0 = t0 y2
    F12_OP2(mc, CSQR, C, C);
    F12_OP3(mc, MUL12, C, C, A);
    F12_OP2(mc, CSQR, C, C);            // t1 = (t1^2 t0)^2
    F12_OP2(mc, CONJ, B, MM);           // y1 = conj(m)
    F12_OP3(mc, MUL12, A, C, B);        // t0 = t1 y1
    F12_OP3(mc, MUL12, C, C, Y0);       // t1 = t1 y0
    F12_OP2(mc, CSQR, A, A);
    F12_OP3(mc, MUL12, A, A, C);        // out = t0^2 t1
    if (!ark_norm) return A;
    // arkworks' value: out^(2x(6x^2+3x+1)) = (a b^3 c^6)^2 with a = out^x, b = a^x, c = b^x
    f12_exp_x(mc, MX, A);
    f12_exp_x(mc, MX2, MX);
    f12_exp_x(mc, MX3, MX2);
    F12_OP2(mc, CSQR, B, MX2);
    F12_OP3(mc, MUL12, B, B, MX2);      // b^3
    F12_OP2(mc, CSQR, C, MX3);
    F12_OP3(mc, MUL12, C, C, MX3);
    F12_OP2(mc, CSQR, C, C);            // c^6
    F12_OP3(mc, MUL12, A, MX, B);
    F12_OP3(mc, MUL12, A, A, C);
    F12_OP2(mc, CSQR, A, A);
    return A;
}

// globals every machine starts from: slot 0 = 0, gamma[k][i] at SIPP_F12_GAMMA_SLOT0 + ((k-1) 6 + i) 2 + c
SIPP_HD void f12_fill_global(uint32_t* slots, int j) {  // j in [0, 37): one slot per call (lanes share the work)
    if (j == 0) { lp_store(slots, 0, fq_zero()); return; }
    const int idx = j - SIPP_F12_GAMMA_SLOT0, c = idx & 1, i = (idx >> 1) % 6, k = (idx >> 1) / 6 + 1;
    const Fq2 g = frob_gamma(k, i);
    lp_store(slots, j, c ? g.c1 : g.c0);
}

}  // namespace sipp
