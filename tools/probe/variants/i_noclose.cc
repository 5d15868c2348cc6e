// poseidon_avx512.cc -- AVX-512 implementation of the plonky2 Poseidon permutation over Goldilocks (width 12).
//
// The Fiat-Shamir transcript of the SIPP native protocol (/root/reference/src/transcript_native.rs:25-30) is a strictly
// sequential chain of 8n + 13 + 27 log2(n) permutations (prover_native.rs:36-39 absorbs every A_i, B_i) that must stay on
// the host; at the sizes the GPU finishes in milliseconds this chain IS the prove time, so one permutation has to be as
// short as the machine allows.  This file computes exactly the same function as the portable code in transcript.cc
// (selected at run time when the CPU has AVX-512 F/DQ/VL + BMI2; with AVX-512 IFMA as well, the faster path described below):
//   full rounds    state in two zmm registers (lanes 0..7, 8..11); x^7 with 4 x vpmuludq 64x64->128 products (high halves by
//                  movehdup, the low word joined by moveldup + blend: port 5 instead of more shifts on port 0) and the
//                  2^64 = 2^32 - 1, 2^96 = -1 reduction; the circulant MDS layer as 36 FP64 FMAs on the 32-bit halves
//                  (sums < 2^43 are exact in double) in COLUMN form: the halves are stored once as doubles and every s[j]
//                  comes back as a broadcast load times a constant column -- no permute network on port 5
//   partial rounds sparse form (tables derived in transcript.cc): the lane-0 S-box and the 11-term dot product run on the
//                  scalar ports (mulx / adc, 192-bit lazy accumulation in two carry chains) while the rank-1 update of
//                  lanes 1..11 runs on the vector ports.  The 11-term sum of round r reads the state of round r - 1 (one
//                  extra product restores the missing rank-1 term) and the constant of the S-box output is folded in, so
//                  the dependent chain of a round is the S-box, one multiply-add and one reduction.
// Measured on the GPU box's Xeon (3.79 GHz, tools/probe/run_poseidon_lab.sh): 1.506 -> 1.385 (dot order) -> 1.223 (register MDS)
// -> 1.18 -> 1.153 (reduction on the carry flag) -> 1.116 (192-bit accumulators pinned in registers by asm blocks) -> 1.065 (the
// scalar 128 -> 64-bit reduction as one 11-instruction asm block: latency 12.2 -> 10.3 cycles) -> 1.027 (column-form MDS,
// port-balanced vector product) -> 1.012 (a partial round as three hand-allocated asm blocks) -> 1.004 us per permutation (the
// two carries of the MDS recombination decided in parallel).  Layer times: vector product 36 cycles latency / 12.8 throughput, S-box layer
// 135 cycles (latency-bound: 3 dependent products), MDS layer ~95, full round 230, partial round ~92 (scalar x^7 chain 33).
//
// A second path (poseidon_permute_ifma, used when the CPU also has AVX-512 IFMA) computes the same function in 0.611 us:
//   partial rounds every rank-1 update unrolled algebraically (PoseidonIfmaTables, poseidon_fast.h): the state after round j is a
//                  linear function of the entering state y and the S-box outputs x_0..x_j, so the 22 x (11-term sum + 11 updates)
//                  become 32 row sums fed by vpmadd52luq / vpmadd52huq -- seven instructions per eight 64 x 64-bit products, one
//                  reduction per ROW instead of one per product -- and the chain variable is rescaled round by round so that a
//                  round's dependent chain is the S-box and one modular addition (36 cycles; 22 rounds: ~950 instead of 2,020)
//   full rounds    lanes 0..7 in ONE vector with a latency-optimised product, lanes 8..11 on the scalar ports in its shadow; MDS
//                  layer on vpmadd52luq (32-bit halves x small entries are exact, no int <-> double conversions) with the round
//                  constants of the next layer as initial values; 170 instead of 230 cycles
//   lessons kept in the code: SIPP_THROUGH_MEMORY (GCC forwards stored vectors through port-5 shuffles unless told not to: 93 -> 74
//                  cycles per MDS layer), accumulators as named variables selected by switch (an array or a pointer select keeps
//                  them in memory), every vector compare that can be made rare moved behind a cold branch.
// Steps on the box: 1.004 -> 0.755 (IFMA partial rounds) -> 0.742 (IFMA MDS) -> 0.726 (scalar lanes 8..11, fast product) -> 0.669
// (real memory broadcasts) -> 0.644 (constants folded) -> 0.613 (borrow on a cold branch) -> 0.611 us (MDS carry on a cold branch).
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include "poseidon_fast.h"

#if defined(__x86_64__)
#define SIPP_AVX512 __attribute__((target("avx512f,avx512dq,avx512vl,bmi2,adx")))

namespace sipp {
namespace {

typedef unsigned __int128 u128;
const uint64_t EPS = 0xFFFFFFFFull;
const uint64_t GL_P = 0xFFFFFFFF00000001ull;

// ------------------------------------------------------------------------------------------------ scalar helpers
// (lo + 2^64 hi) mod p, result < 2^64: 2^64 = 2^32 - 1 and 2^96 = -1 (mod p).  One asm block, everything in registers: with the
// _subborrow_u64 / _addcarry_u64 intrinsics GCC spilled the intermediate through the stack (their result is a pointer argument)
// and every reduction on the S-box chain paid a store-to-load forward.  The borrow of lo - (hi >> 32) has probability 2^-32:
// a forward branch that is never taken; the carry of the final addition is a coin flip: branch-free (sbb mask).
// (lo + 2^64 hi) mod p as ONE asm block of 10 instructions: 2^64 = 2^32 - 1, 2^96 = -1 (mod p).  The borrow of lo - (hi >> 32) has
// probability 2^-32 per call: a forward branch to a cold fix-up (sub + jc fuse) instead of the four-instruction cmov sequence GCC
// makes of it; the carry of the final addition is a coin flip: lea + cmovc on the flags of the addition itself.
SIPP_AVX512 inline uint64_t s_red128(uint64_t lo, uint64_t hi) {
    uint64_t t, m = hi;
    const uint64_t eps = EPS;
    asm("mov %[m], %[t]\n\t"
        "shr $32, %[t]\n\t"            // hh
        "mov %k[m], %k[m]\n\t"          // hl
        "sub %[t], %[lo]\n\t"
        "jc 2f\n"
        "1:\n\t"
        "mov %[m], %[t]\n\t"
        "shl $32, %[t]\n\t"
        "sub %[m], %[t]\n\t"            // hl (2^32 - 1)
        "add %[t], %[lo]\n\t"
        "lea (%[lo],%[eps]), %[t]\n\t"
        "cmovc %[t], %[lo]\n\t"
        ".subsection 1\n"
        "2:\n\t"
        "sub %[eps], %[lo]\n\t"
        "jmp 1b\n\t"
        ".previous"
        : [lo] "+r"(lo), [m] "+r"(m), [t] "=&r"(t)
        : [eps] "r"(eps)
        : "cc");
    return lo;
}
SIPP_AVX512 inline uint64_t s_mul(uint64_t a, uint64_t b) {
    unsigned long long hi;
    uint64_t lo = _mulx_u64(a, b, &hi);
    return s_red128(lo, hi);
}
SIPP_AVX512 inline uint64_t s_add(uint64_t a, uint64_t b) {  // any a, b
    uint64_t r = a + b;
    uint64_t t = r + ((0 - (uint64_t)(r < a)) & EPS);
    return t + ((0 - (uint64_t)(t < r)) & EPS);
}
SIPP_AVX512 inline uint64_t s_pow7(uint64_t x) {
    uint64_t x2 = s_mul(x, x), x3 = s_mul(x2, x), x4 = s_mul(x2, x2);
    return s_mul(x3, x4);
}
// (lo, hi, top) += a * b as ONE asm block: mulx + add / adc / adc with the three limbs pinned in registers.  Written with the
// _addcarry_u64 intrinsics (which take the address of their result) GCC kept the accumulator in memory and every term went
// through a store-to-load forward -- 109 cycles per partial round on the GPU box's Xeon for a 55-cycle dependent chain.
SIPP_AVX512 inline void mac192(unsigned long long& lo, unsigned long long& hi, unsigned long long& top, uint64_t a, uint64_t b) {
    unsigned long long pl, ph;
    asm("mulx %3, %0, %1\n\t"
        "add %0, %4\n\t"
        "adc %1, %5\n\t"
        "adc $0, %6"
        : "=&r"(pl), "=&r"(ph), "+d"(a), "+rm"(b), "+r"(lo), "+r"(hi), "+r"(top)
        :
        : "cc");
}
// sum_{i<11} a[i] * b[i] + extra_a * extra_b, reduced once (three-limb lazy accumulation; 2^128 = -2^32 mod p)
SIPP_AVX512 inline uint64_t s_dot11p(const uint64_t* a, const uint64_t* b, uint64_t ea, uint64_t eb) {
    unsigned long long lo = 0, hi = 0, top = 0;
#pragma GCC unroll 11
    for (int i = 0; i < 11; i++) mac192(lo, hi, top, a[i], b[i]);
    // the term that depends on the S-box output of this round enters last: it is the only one on the dependent chain
    mac192(lo, hi, top, ea, eb);
    uint64_t r = s_red128(lo, hi);
    uint64_t t = (uint64_t)top << 32;  // top <= 12
    uint64_t d = r - t;
    if (__builtin_expect(r < t, 0)) d -= EPS;  // borrowed 2^64 = EPS
    return d;
}


// 192-bit lazy sum of 11 products in two independent chains (even / odd terms)
struct Acc192 { unsigned long long lo, hi, top; };
SIPP_AVX512 inline Acc192 s_dot11_raw(const uint64_t* a, const uint64_t* b) {
    unsigned long long lo0 = 0, hi0 = 0, top0 = 0, lo1 = 0, hi1 = 0, top1 = 0;
#pragma GCC unroll 6
    for (int i = 0; i < 11; i += 2) {
        mac192(lo0, hi0, top0, a[i], b[i]);
        if (i + 1 < 11) mac192(lo1, hi1, top1, a[i + 1], b[i + 1]);
    }
    asm("add %3, %0\n\tadc %4, %1\n\tadc %5, %2" : "+r"(lo0), "+r"(hi0), "+r"(top0) : "r"(lo1), "r"(hi1), "r"(top1) : "cc");
    return Acc192{lo0, hi0, top0};
}
SIPP_AVX512 inline void acc_mul(Acc192& s, uint64_t a, uint64_t b) {
    unsigned long long lo = s.lo, hi = s.hi, top = s.top;
    mac192(lo, hi, top, a, b);
    s.lo = lo; s.hi = hi; s.top = top;
}
SIPP_AVX512 inline void acc_add(Acc192& s, uint64_t v) {
    unsigned long long lo = s.lo, hi = s.hi, top = s.top;
    asm("add %3, %0\n\tadc $0, %1\n\tadc $0, %2" : "+r"(lo), "+r"(hi), "+r"(top) : "r"(v) : "cc");
    s.lo = lo; s.hi = hi; s.top = top;
}
SIPP_AVX512 inline uint64_t acc_reduce(const Acc192& s) {
    uint64_t r = s_red128(s.lo, s.hi);
    uint64_t t = (uint64_t)s.top << 32;  // top <= 13; 2^128 = -2^32
    uint64_t d = r - t;
    if (__builtin_expect(r < t, 0)) d -= EPS;
    return d;
}

// The MDS layers below store the state halves and read every one back as a broadcast LOAD (load ports).  Left alone, GCC forwards
// the stored vectors through registers instead -- vextracti64x2 / valignq / vpbroadcastq chains, all on port 5, ~40 shuffles per
// layer; this barrier makes the round trip through memory real.
// (the operand names the one array: a "memory" clobber would also spill and reload every accumulator that lives in an array)
#define SIPP_THROUGH_MEMORY(arr) asm volatile("" : "+m"(arr))

// ------------------------------------------------------------------------------------------------ vector helpers
SIPP_AVX512 inline __m512i v_reduce(__m512i lo, __m512i hi) {
    const __m512i eps = _mm512_set1_epi64((long long)EPS);
    __m512i hh = _mm512_srli_epi64(hi, 32);
    __m512i t = _mm512_sub_epi64(lo, hh);
    __mmask8 b = _mm512_cmplt_epu64_mask(lo, hh);
    t = _mm512_mask_sub_epi64(t, b, t, eps);
    __m512i m = _mm512_mul_epu32(hi, eps);  // (hi & 0xffffffff) * (2^32 - 1)
    __m512i r = _mm512_add_epi64(t, m);
    __mmask8 c = _mm512_cmplt_epu64_mask(r, m);
    return _mm512_mask_add_epi64(r, c, r, eps);
}
// the high 32-bit halves as multiplier operands come from movehdup (port 5) instead of a shift (port 0, where the four
// multiplies already queue): vpmuludq reads only the low half of each lane; likewise the low word is assembled by moveldup + blend
SIPP_AVX512 inline __m512i v_hi(__m512i x) { return _mm512_castps_si512(_mm512_movehdup_ps(_mm512_castsi512_ps(x))); }
SIPP_AVX512 inline __m512i v_join(__m512i ll, __m512i t1) {  // (ll & 0xffffffff) | (t1 << 32)
    return _mm512_mask_blend_epi32(0xAAAA, ll, _mm512_castps_si512(_mm512_moveldup_ps(_mm512_castsi512_ps(t1))));
}
SIPP_AVX512 inline __m512i v_mul(__m512i x, __m512i y) {
    const __m512i lo32 = _mm512_set1_epi64((long long)EPS);
    __m512i xh = v_hi(x), yh = v_hi(y);
    __m512i ll = _mm512_mul_epu32(x, y), lh = _mm512_mul_epu32(x, yh), hl = _mm512_mul_epu32(xh, y), hh = _mm512_mul_epu32(xh, yh);
    __m512i t0 = _mm512_add_epi64(hl, _mm512_srli_epi64(ll, 32));
    __m512i t1 = _mm512_add_epi64(lh, _mm512_and_si512(t0, lo32));
    __m512i hi = _mm512_add_epi64(hh, _mm512_add_epi64(_mm512_srli_epi64(t0, 32), _mm512_srli_epi64(t1, 32)));
    __m512i lo = v_join(ll, t1);
    return v_reduce(lo, hi);
}
SIPP_AVX512 inline __m512i v_sqr(__m512i x) {
    const __m512i lo32 = _mm512_set1_epi64((long long)EPS);
    __m512i xh = v_hi(x);
    __m512i ll = _mm512_mul_epu32(x, x), lh = _mm512_mul_epu32(x, xh), hh = _mm512_mul_epu32(xh, xh);
    __m512i t0 = _mm512_add_epi64(lh, _mm512_srli_epi64(ll, 32));
    __m512i t1 = _mm512_add_epi64(lh, _mm512_and_si512(t0, lo32));
    __m512i hi = _mm512_add_epi64(hh, _mm512_add_epi64(_mm512_srli_epi64(t0, 32), _mm512_srli_epi64(t1, 32)));
    __m512i lo = v_join(ll, t1);
    return v_reduce(lo, hi);
}
SIPP_AVX512 inline __m512i v_pow7(__m512i x) {
    __m512i x2 = v_sqr(x), x4 = v_sqr(x2), x3 = v_mul(x2, x);
    return v_mul(x3, x4);
}
// a + b with b canonical (< p): a single wrap correction suffices
SIPP_AVX512 inline __m512i v_add_canon(__m512i a, __m512i b) {
    const __m512i eps = _mm512_set1_epi64((long long)EPS);
    __m512i r = _mm512_add_epi64(a, b);
    __mmask8 c = _mm512_cmplt_epu64_mask(r, a);
    return _mm512_mask_add_epi64(r, c, r, eps);
}
SIPP_AVX512 inline __m512i v_canon(__m512i a) {
    const __m512i p = _mm512_set1_epi64((long long)GL_P);
    return _mm512_min_epu64(a, _mm512_sub_epi64(a, p));
}

// out[r] = sum_i s[(i + r) mod 12] * CIRC[i] + 8 s[0] [r == 0] on the 32-bit halves, in FP64 (sums < 2^43 are exact).
// The twelve rotations of the state are built in registers: with E0 = s[0..7], E1 = s[8..11, 0..3], E2 = s[4..11] the window
// s[i..i+7] is one valignq of two neighbours.  Rows 8..11 only fill half a vector, so their low and high halves share one:
// F_k = lo[4k..4k+3] | hi[4k..4k+3], and the window s[8+i..11+i] is a two-source permute of two neighbouring F's.
// Column form: out = sum_j s[j] * column_j.  The state halves are stored once as doubles and every s[j] comes back as a
// broadcast LOAD (load ports; forwarded from the 64-byte stores), multiplied by a constant column vector -- no valignq / permute
// network on port 5.  Rows 0..7: two accumulator sets (low / high halves); rows 8..11: one register with the low sums in lanes
// 0..3 and the high sums in lanes 4..7 (the broadcast of the high half is merged into the upper lanes by the load itself).
SIPP_AVX512 inline void v_mds(__m512i& s0, __m512i& s1, const PoseidonFastTables& T) {
    const __m512i lo32 = _mm512_set1_epi64((long long)EPS);
    alignas(64) double L[16], H[16];
    _mm512_store_pd(L, _mm512_cvtepu64_pd(_mm512_and_si512(s0, lo32)));
    _mm512_store_pd(L + 8, _mm512_cvtepu64_pd(_mm512_and_si512(s1, lo32)));
    _mm512_store_pd(H, _mm512_cvtepu64_pd(_mm512_srli_epi64(s0, 32)));
    _mm512_store_pd(H + 8, _mm512_cvtepu64_pd(_mm512_srli_epi64(s1, 32)));
    SIPP_THROUGH_MEMORY(L);
    SIPP_THROUGH_MEMORY(H);
    const double *Lm = L, *Hm = H;
    __m512d al[4], ah[4], ab[4];
#pragma GCC unroll 12
    for (int j = 0; j < 12; j++) {
        const __m512d bl = _mm512_set1_pd(Lm[j]), bh = _mm512_set1_pd(Hm[j]);
        const __m512d bb = _mm512_mask_broadcastsd_pd(bl, 0xF0, _mm_load_sd(&Hm[j]));
        const __m512d ca = _mm512_load_pd(T.mds_col_a[j]), cb = _mm512_load_pd(T.mds_col_b[j]);
        if (j < 4) {
            al[j] = _mm512_mul_pd(bl, ca); ah[j] = _mm512_mul_pd(bh, ca); ab[j] = _mm512_mul_pd(bb, cb);
        } else {
            al[j & 3] = _mm512_fmadd_pd(bl, ca, al[j & 3]); ah[j & 3] = _mm512_fmadd_pd(bh, ca, ah[j & 3]); ab[j & 3] = _mm512_fmadd_pd(bb, cb, ab[j & 3]);
        }
    }
    const __m512i eps = lo32;
    auto combine = [&](__m512i alo, __m512i ahi) SIPP_AVX512 {  // < 2^43 each; value = alo + 2^32 ahi
        // the carry of alo + (ahi << 32) and the carry of adding (ahi >> 32) (2^32 - 1) are decided in parallel (they exclude each
        // other: a wrapped low word is < 2^43), each worth one + EPS
        __m512i lo = _mm512_add_epi64(alo, _mm512_slli_epi64(ahi, 32));
        __m512i hi = _mm512_srli_epi64(ahi, 32);
        __m512i m = _mm512_sub_epi64(_mm512_slli_epi64(hi, 32), hi);  // hi * (2^32 - 1), hi < 2^11
        __m512i r = _mm512_add_epi64(lo, m);
        __mmask8 c1 = _mm512_cmplt_epu64_mask(lo, alo);
        __mmask8 c2 = _mm512_cmplt_epu64_mask(r, m);
        return _mm512_mask_add_epi64(r, (__mmask8)(c1 | c2), r, eps);
    };
    s0 = combine(_mm512_cvtpd_epu64(_mm512_add_pd(_mm512_add_pd(al[0], al[1]), _mm512_add_pd(al[2], al[3]))),
                 _mm512_cvtpd_epu64(_mm512_add_pd(_mm512_add_pd(ah[0], ah[1]), _mm512_add_pd(ah[2], ah[3]))));
    const __m512i bi = _mm512_cvtpd_epu64(_mm512_add_pd(_mm512_add_pd(ab[0], ab[1]), _mm512_add_pd(ab[2], ab[3])));
    s1 = combine(bi, _mm512_alignr_epi64(bi, bi, 4));              // lanes 4..7 of s1 are don't-care
}

// ------------------------------------------------------------------------------------------------ partial rounds, hand-allocated
// One partial round as three asm blocks with every value pinned in a register.  Written with intrinsics the loop body was 285
// instructions of which ~80 were register moves and spills (GCC ran out of registers around the 192-bit accumulators) at ~3.1
// instructions per cycle: 92 cycles per round for a 51-cycle dependent chain.  Everything a round reads besides the state sits in
// one 256-byte record (PoseidonFastTables::pr), addressed off a single pointer.
//   pr_dot:   acc = sum_i vhat[i] U[i] (two chains) + x_prev kprev + m00 post        (off the chain: U is the state of round r - 1)
//   pr_sbox:  p7 = u0^7 (three dependent products, 10-instruction reductions), x = p7 + post
//   pr_finish: u0' = (acc + p7 m00) mod p
SIPP_AVX512 inline void pr_dot(const PartialRound* p, const uint64_t* ub, uint64_t xp, unsigned long long& lo, unsigned long long& hi, unsigned long long& top) {
    unsigned long long lo0, hi0, t0, lo1, hi1, t1, pl, ph;
    asm("mov 0(%[ub]), %%rdx\n\t" "mulx 128(%[p]), %[lo0], %[hi0]\n\t"
        "mov 8(%[ub]), %%rdx\n\t" "mulx 136(%[p]), %[lo1], %[hi1]\n\t"
        "xor %k[t0], %k[t0]\n\t" "xor %k[t1], %k[t1]\n\t"
        "mov 16(%[ub]), %%rdx\n\t" "mulx 144(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo0]\n\t" "adc %[ph], %[hi0]\n\t" "adc $0, %[t0]\n\t"
        "mov 24(%[ub]), %%rdx\n\t" "mulx 152(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo1]\n\t" "adc %[ph], %[hi1]\n\t" "adc $0, %[t1]\n\t"
        "mov 32(%[ub]), %%rdx\n\t" "mulx 160(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo0]\n\t" "adc %[ph], %[hi0]\n\t" "adc $0, %[t0]\n\t"
        "mov 40(%[ub]), %%rdx\n\t" "mulx 168(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo1]\n\t" "adc %[ph], %[hi1]\n\t" "adc $0, %[t1]\n\t"
        "mov 48(%[ub]), %%rdx\n\t" "mulx 176(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo0]\n\t" "adc %[ph], %[hi0]\n\t" "adc $0, %[t0]\n\t"
        "mov 56(%[ub]), %%rdx\n\t" "mulx 184(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo1]\n\t" "adc %[ph], %[hi1]\n\t" "adc $0, %[t1]\n\t"
        "mov 64(%[ub]), %%rdx\n\t" "mulx 192(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo0]\n\t" "adc %[ph], %[hi0]\n\t" "adc $0, %[t0]\n\t"
        "mov 72(%[ub]), %%rdx\n\t" "mulx 200(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo1]\n\t" "adc %[ph], %[hi1]\n\t" "adc $0, %[t1]\n\t"
        "mov 80(%[ub]), %%rdx\n\t" "mulx 208(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo0]\n\t" "adc %[ph], %[hi0]\n\t" "adc $0, %[t0]\n\t"
        "mov %[xp], %%rdx\n\t" "mulx 216(%[p]), %[pl], %[ph]\n\t" "add %[pl], %[lo1]\n\t" "adc %[ph], %[hi1]\n\t" "adc $0, %[t1]\n\t"
        "add 224(%[p]), %[lo1]\n\t" "adc $0, %[hi1]\n\t" "adc $0, %[t1]\n\t"
        "add %[lo1], %[lo0]\n\t" "adc %[hi1], %[hi0]\n\t" "adc %[t1], %[t0]"
        : [lo0] "=&r"(lo0), [hi0] "=&r"(hi0), [t0] "=&r"(t0), [lo1] "=&r"(lo1), [hi1] "=&r"(hi1), [t1] "=&r"(t1), [pl] "=&r"(pl), [ph] "=&r"(ph)
        : [p] "r"(p), [ub] "r"(ub), [xp] "r"(xp)
        : "rdx", "cc", "memory");
    lo = lo0; hi = hi0; top = t0;
}
SIPP_AVX512 inline void pr_sbox(const PartialRound* p, uint64_t u, uint64_t& p7, uint64_t& x) {
    unsigned long long a, b, h, t;
    const uint64_t eps = EPS;
    asm("mov %[u], %%rdx\n\t" "mulx %[u], %[a], %[h]\n\t"
        "mov %[h], %[t]\n\t" "shr $32, %[t]\n\t" "mov %k[h], %k[h]\n\t" "sub %[t], %[a]\n\t" "jc 11f\n" "10:\n\t"
        "mov %[h], %[t]\n\t" "shl $32, %[t]\n\t" "sub %[h], %[t]\n\t" "add %[t], %[a]\n\t" "lea (%[a],%[eps]), %[t]\n\t" "cmovc %[t], %[a]\n\t"
        "mulx %[a], %[b], %[h]\n\t"
        "mov %[h], %[t]\n\t" "shr $32, %[t]\n\t" "mov %k[h], %k[h]\n\t" "sub %[t], %[b]\n\t" "jc 13f\n" "12:\n\t"
        "mov %[h], %[t]\n\t" "shl $32, %[t]\n\t" "sub %[h], %[t]\n\t" "add %[t], %[b]\n\t" "lea (%[b],%[eps]), %[t]\n\t" "cmovc %[t], %[b]\n\t"
        "mov %[a], %%rdx\n\t" "mulx %[a], %[a], %[h]\n\t"
        "mov %[h], %[t]\n\t" "shr $32, %[t]\n\t" "mov %k[h], %k[h]\n\t" "sub %[t], %[a]\n\t" "jc 15f\n" "14:\n\t"
        "mov %[h], %[t]\n\t" "shl $32, %[t]\n\t" "sub %[h], %[t]\n\t" "add %[t], %[a]\n\t" "lea (%[a],%[eps]), %[t]\n\t" "cmovc %[t], %[a]\n\t"
        "mov %[b], %%rdx\n\t" "mulx %[a], %[a], %[h]\n\t"
        "mov %[h], %[t]\n\t" "shr $32, %[t]\n\t" "mov %k[h], %k[h]\n\t" "sub %[t], %[a]\n\t" "jc 17f\n" "16:\n\t"
        "mov %[h], %[t]\n\t" "shl $32, %[t]\n\t" "sub %[h], %[t]\n\t" "add %[t], %[a]\n\t" "lea (%[a],%[eps]), %[t]\n\t" "cmovc %[t], %[a]\n\t"
        "mov %[a], %[b]\n\t" "add 232(%[p]), %[b]\n\t" "lea (%[b],%[eps]), %[t]\n\t" "cmovc %[t], %[b]\n\t"
        ".subsection 1\n"
        "11:\n\t" "sub %[eps], %[a]\n\t" "jmp 10b\n"
        "13:\n\t" "sub %[eps], %[b]\n\t" "jmp 12b\n"
        "15:\n\t" "sub %[eps], %[a]\n\t" "jmp 14b\n"
        "17:\n\t" "sub %[eps], %[a]\n\t" "jmp 16b\n"
        ".previous"
        : [a] "=&r"(a), [b] "=&r"(b), [h] "=&r"(h), [t] "=&r"(t)
        : [u] "r"(u), [p] "r"(p), [eps] "r"(eps)
        : "rdx", "cc", "memory");
    p7 = a; x = b;
}
SIPP_AVX512 inline uint64_t pr_finish(unsigned long long lo, unsigned long long hi, unsigned long long top, uint64_t p7, uint64_t m00) {
    unsigned long long pl, ph;
    const uint64_t eps = EPS;
    asm("mov %[p7], %%rdx\n\t" "mulx %[m00], %[pl], %[ph]\n\t" "add %[pl], %[lo]\n\t" "adc %[ph], %[hi]\n\t" "adc $0, %[top]\n\t"
        "mov %[hi], %[pl]\n\t" "shr $32, %[pl]\n\t" "mov %k[hi], %k[hi]\n\t" "sub %[pl], %[lo]\n\t" "jc 21f\n" "20:\n\t"
        "mov %[hi], %[pl]\n\t" "shl $32, %[pl]\n\t" "sub %[hi], %[pl]\n\t" "add %[pl], %[lo]\n\t" "lea (%[lo],%[eps]), %[pl]\n\t" "cmovc %[pl], %[lo]\n\t"
        "shl $32, %[top]\n\t" "sub %[top], %[lo]\n\t" "jc 23f\n" "22:\n\t"
        ".subsection 1\n"
        "21:\n\t" "sub %[eps], %[lo]\n\t" "jmp 20b\n"
        "23:\n\t" "sub %[eps], %[lo]\n\t" "jmp 22b\n"
        ".previous"
        : [lo] "+r"(lo), [hi] "+r"(hi), [top] "+r"(top), [pl] "=&r"(pl), [ph] "=&r"(ph)
        : [p7] "r"(p7), [m00] "r"(m00), [eps] "r"(eps)
        : "rdx", "cc");
    return lo;
}

SIPP_AVX512 inline void v_full_round(__m512i& s0, __m512i& s1, const uint64_t* rc16, const PoseidonFastTables& T) {
    s0 = v_pow7(v_add_canon(s0, _mm512_load_si512(rc16)));
    s1 = v_pow7(v_add_canon(s1, _mm512_load_si512(rc16 + 8)));
    v_mds(s0, s1, T);
}


// ------------------------------------------------------------------------------------------------ IFMA path
// Partial rounds with every rank-1 update unrolled algebraically (tables: poseidon_fast.h, PoseidonIfmaTables).  The scalar ports
// run nothing but the dependent chain -- S-box, one multiply-add by m00, a 5-instruction reduction -- plus, off the chain, one
// row close per round (recombine the three accumulator lanes of C-row j + 1, add its newest term, reduce); every other product of
// the 22 rounds (330 for the initial matrix, 450 for the x_k terms) is a vpmadd52 lane on the vector ports, seven instructions per
// eight 64 x 64-bit products, with ONE reduction per row at the very end instead of one per product.
#define SIPP_IFMA __attribute__((target("avx512f,avx512dq,avx512vl,avx512ifma,bmi2,adx")))
struct IfmaBlock { __m512i a0, a1, a2; };
// acc += x * c over eight lanes: x = xl + 2^52 xh, c = cl + 2^52 ch (vpmadd52 reads the low 52 bits of its operands)
SIPP_IFMA inline void ifma_unit(IfmaBlock& A, __m512i xb, __m512i xh, const uint64_t* c) {
    const __m512i cl = _mm512_load_si512(c), ch = _mm512_load_si512(c + 8);
    A.a0 = _mm512_madd52lo_epu64(A.a0, xb, cl);
    A.a1 = _mm512_madd52hi_epu64(A.a1, xb, cl);
    A.a2 = _mm512_madd52hi_epu64(A.a2, xb, ch);
    A.a1 = _mm512_madd52lo_epu64(A.a1, xb, ch);
    A.a2 = _mm512_madd52hi_epu64(A.a2, xh, cl);
    A.a1 = _mm512_madd52lo_epu64(A.a1, xh, cl);
    A.a2 = _mm512_madd52lo_epu64(A.a2, xh, ch);
}
SIPP_IFMA inline void ifma_store(uint64_t* sc, const IfmaBlock& A) {
    _mm512_store_si512(sc, A.a0);
    _mm512_store_si512(sc + 8, A.a1);
    _mm512_store_si512(sc + 16, A.a2);
}
// (a0 + 2^52 a1 - 2^8 a2 + c x) mod p: one lane of an accumulator block (2^104 = -2^8; the row constant carries a + p, so the
// subtraction cannot borrow) plus the newest term of the row
SIPP_IFMA inline uint64_t row_close(uint64_t a0, uint64_t a1, uint64_t a2, uint64_t c, uint64_t x) {
    unsigned long long t, pl, ph;  // (a2 doubles as the third limb once it has been subtracted: the block is short of registers around it)
    const uint64_t eps = EPS;
    asm("mov %[a1], %[t]\n\t" "shl $52, %[t]\n\t" "shr $12, %[a1]\n\t" "add %[t], %[a0]\n\t" "adc $0, %[a1]\n\t"
        "shl $8, %[a2]\n\t" "sub %[a2], %[a0]\n\t" "sbb $0, %[a1]\n\t"
        "mulx %[c], %[pl], %[ph]\n\t" "xor %k[a2], %k[a2]\n\t" "add %[pl], %[a0]\n\t" "adc %[ph], %[a1]\n\t" "adc $0, %[a2]\n\t"
        "mov %[a1], %[t]\n\t" "shr $32, %[t]\n\t" "mov %k[a1], %k[a1]\n\t" "sub %[t], %[a0]\n\t" "jc 31f\n" "30:\n\t"
        "mov %[a1], %[t]\n\t" "shl $32, %[t]\n\t" "sub %[a1], %[t]\n\t" "add %[t], %[a0]\n\t" "lea (%[a0],%[eps]), %[t]\n\t" "cmovc %[t], %[a0]\n\t"
        "shl $32, %[a2]\n\t" "sub %[a2], %[a0]\n\t" "jc 33f\n" "32:\n\t"
        ".subsection 1\n"
        "31:\n\t" "sub %[eps], %[a0]\n\t" "jmp 30b\n"
        "33:\n\t" "sub %[eps], %[a0]\n\t" "jmp 32b\n"
        ".previous"
        : [a0] "+r"(a0), [a1] "+r"(a1), [a2] "+r"(a2), [t] "=&r"(t), [pl] "=&r"(pl), [ph] "=&r"(ph)
        : [c] "rm"(c), "d"(x), [eps] "r"(eps)
        : "cc");
    return a0;
}
// (e + z7) mod 2^64-representative: one wrap correction
SIPP_IFMA inline uint64_t chain_close(uint64_t e, uint64_t z7) {
    unsigned long long t;
    const uint64_t eps = EPS;
    asm("add %[z7], %[e]\n\t" "lea (%[e],%[eps]), %[t]\n\t" "cmovc %[t], %[e]" : [e] "+r"(e), [t] "=&r"(t) : [z7] "r"(z7), [eps] "r"(eps) : "cc");
    return e;
}
// Latency-optimised vector product for the path below, where ONE vector (lanes 0..7) is left on the vector ports and its x^7 chain
// is the critical path of a full round: the partial products are summed as a shallow tree (three 32-bit middle terms, no carry
// between them) and hl (2^32 - 1) is a shift and a subtraction instead of a multiply -- 28 instead of 35 cycles, four more micro-ops.
// the borrow of lo - (hi >> 32) needs lo < 2^32: 2^-32 per lane.  Its fix is a cold out-of-line call behind a branch on the mask
// (kortest + jne, predicted), so the compare is off the dependent chain of the product: 3 cycles less per product.
SIPP_IFMA __attribute__((noinline, cold)) __m512i v_borrow_fix(__m512i t, __mmask8 b) {
    return _mm512_mask_sub_epi64(t, b, t, _mm512_set1_epi64((long long)EPS));
}
SIPP_IFMA __attribute__((noinline, cold)) __m512i v_carry_fix(__m512i r, __mmask8 c) {
    return _mm512_mask_add_epi64(r, c, r, _mm512_set1_epi64((long long)EPS));
}
SIPP_IFMA inline __m512i v_reduce_fast(__m512i lo, __m512i hi) {
    const __m512i eps = _mm512_set1_epi64((long long)EPS);
    __m512i hh = _mm512_srli_epi64(hi, 32);
    __m512i t = _mm512_sub_epi64(lo, hh);
    __mmask8 b = _mm512_cmplt_epu64_mask(lo, hh);
    if (__builtin_expect(b != 0, 0)) t = v_borrow_fix(t, b);
    __m512i m = _mm512_sub_epi64(_mm512_slli_epi64(hi, 32), _mm512_and_si512(hi, eps));
    __m512i r = _mm512_add_epi64(t, m);
    __mmask8 c = _mm512_cmplt_epu64_mask(r, m);
    return _mm512_mask_add_epi64(r, c, r, eps);
}
SIPP_IFMA inline __m512i v_mul_fast(__m512i x, __m512i y) {
    const __m512i lo32 = _mm512_set1_epi64((long long)EPS);
    __m512i xh = v_hi(x), yh = v_hi(y);
    __m512i ll = _mm512_mul_epu32(x, y), lh = _mm512_mul_epu32(x, yh), hl = _mm512_mul_epu32(xh, y), hh = _mm512_mul_epu32(xh, yh);
    __m512i mid = _mm512_add_epi64(_mm512_add_epi64(_mm512_srli_epi64(ll, 32), _mm512_and_si512(lh, lo32)), _mm512_and_si512(hl, lo32));
    __m512i hs = _mm512_add_epi64(_mm512_add_epi64(_mm512_srli_epi64(lh, 32), _mm512_srli_epi64(hl, 32)), hh);
    return v_reduce_fast(v_join(ll, mid), _mm512_add_epi64(hs, _mm512_srli_epi64(mid, 32)));
}
SIPP_IFMA inline __m512i v_sqr_fast(__m512i x) {
    const __m512i lo32 = _mm512_set1_epi64((long long)EPS);
    __m512i xh = v_hi(x);
    __m512i ll = _mm512_mul_epu32(x, x), lh = _mm512_mul_epu32(x, xh), hh = _mm512_mul_epu32(xh, xh);
    __m512i lhl = _mm512_and_si512(lh, lo32), lhh = _mm512_srli_epi64(lh, 32);
    __m512i mid = _mm512_add_epi64(_mm512_add_epi64(_mm512_srli_epi64(ll, 32), lhl), lhl);
    __m512i hs = _mm512_add_epi64(_mm512_add_epi64(lhh, lhh), hh);
    return v_reduce_fast(v_join(ll, mid), _mm512_add_epi64(hs, _mm512_srli_epi64(mid, 32)));
}
SIPP_IFMA inline __m512i v_pow7_fast(__m512i x) {
    __m512i x2 = v_sqr_fast(x), x4 = v_sqr_fast(x2), x3 = v_mul_fast(x2, x);
    return v_mul_fast(x3, x4);
}
// a + c, one wrap correction (c canonical, or a + c < 2^65 - 2^32)
SIPP_IFMA inline uint64_t s_add1(uint64_t a, uint64_t c) {
    unsigned long long t;
    const uint64_t eps = EPS;
    asm("add %[c], %[a]\n\t" "lea (%[a],%[eps]), %[t]\n\t" "cmovc %[t], %[a]" : [a] "+r"(a), [t] "=&r"(t) : [c] "rm"(c), [eps] "r"(eps) : "cc");
    return a;
}
// u^7, three dependent products (the S-box block of pr_sbox without the constant)
SIPP_IFMA inline uint64_t sbox7(uint64_t u) {
    unsigned long long a, b, h, t;
    const uint64_t eps = EPS;
    asm("mov %[u], %%rdx\n\t" "mulx %[u], %[a], %[h]\n\t"
        "mov %[h], %[t]\n\t" "shr $32, %[t]\n\t" "mov %k[h], %k[h]\n\t" "sub %[t], %[a]\n\t" "jc 41f\n" "40:\n\t"
        "mov %[h], %[t]\n\t" "shl $32, %[t]\n\t" "sub %[h], %[t]\n\t" "add %[t], %[a]\n\t" "lea (%[a],%[eps]), %[t]\n\t" "cmovc %[t], %[a]\n\t"
        "mulx %[a], %[b], %[h]\n\t"
        "mov %[h], %[t]\n\t" "shr $32, %[t]\n\t" "mov %k[h], %k[h]\n\t" "sub %[t], %[b]\n\t" "jc 43f\n" "42:\n\t"
        "mov %[h], %[t]\n\t" "shl $32, %[t]\n\t" "sub %[h], %[t]\n\t" "add %[t], %[b]\n\t" "lea (%[b],%[eps]), %[t]\n\t" "cmovc %[t], %[b]\n\t"
        "mov %[a], %%rdx\n\t" "mulx %[a], %[a], %[h]\n\t"
        "mov %[h], %[t]\n\t" "shr $32, %[t]\n\t" "mov %k[h], %k[h]\n\t" "sub %[t], %[a]\n\t" "jc 45f\n" "44:\n\t"
        "mov %[h], %[t]\n\t" "shl $32, %[t]\n\t" "sub %[h], %[t]\n\t" "add %[t], %[a]\n\t" "lea (%[a],%[eps]), %[t]\n\t" "cmovc %[t], %[a]\n\t"
        "mov %[b], %%rdx\n\t" "mulx %[a], %[a], %[h]\n\t"
        "mov %[h], %[t]\n\t" "shr $32, %[t]\n\t" "mov %k[h], %k[h]\n\t" "sub %[t], %[a]\n\t" "jc 47f\n" "46:\n\t"
        "mov %[h], %[t]\n\t" "shl $32, %[t]\n\t" "sub %[h], %[t]\n\t" "add %[t], %[a]\n\t" "lea (%[a],%[eps]), %[t]\n\t" "cmovc %[t], %[a]\n\t"
        ".subsection 1\n"
        "41:\n\t" "sub %[eps], %[a]\n\t" "jmp 40b\n"
        "43:\n\t" "sub %[eps], %[b]\n\t" "jmp 42b\n"
        "45:\n\t" "sub %[eps], %[a]\n\t" "jmp 44b\n"
        "47:\n\t" "sub %[eps], %[a]\n\t" "jmp 46b\n"
        ".previous"
        : [a] "=&r"(a), [b] "=&r"(b), [h] "=&r"(h), [t] "=&r"(t)
        : [u] "r"(u), [eps] "r"(eps)
        : "rdx", "cc");
    return a;
}
// A full round with lanes 0..7 in one vector and lanes 8..11 on the scalar ports.  Two zmm x^7 share the two 512-bit ports and the
// second is half empty (134 cycles for the S-box layer against 115 for one vector alone); the four scalar x^7 (33 cycles each,
// independent) run in the shadow of the vector chain, which comes first in program order so that it is served first.  The
// scalar lanes reach the MDS layer as plain stores of their 32-bit halves, and rows 8..11 come back through one 64-byte store
// (every scalar instruction saved here is worth ~0.3 cycles per round: the round is bound by the total micro-op flow, see the
// ablations in tools/probe/README.md -- so rows 8..11 are recombined on the vector side and carry their constants too).
// All lanes arrive with their round constant already added: the previous MDS layer starts its sums from the halves of the NEXT
// constants (rc_next) -- an addition less on the vector chain, four fewer scalar additions.
SIPP_IFMA inline __attribute__((always_inline)) void full_round_mixed(__m512i& s0, uint64_t* t, const uint64_t* rc_next, const PoseidonIfmaTables& I) {
    const __m512i lo32 = _mm512_set1_epi64((long long)EPS);
    s0 = v_pow7_fast(s0);
    // MDS layer.  The 32-bit halves of every lane are stored as (low, high) pairs: a 64-bit broadcast of either feeds rows 0..7,
    // ONE 128-bit broadcast of the pair feeds rows 8..11 (lane 2i: low sums of row 8 + i, lane 2i + 1: high sums) -- no merge of
    // two broadcasts on the vector ports; rows 8..11 are recombined on the scalar ports, where their lanes live.
    alignas(64) uint64_t pr[24];
#pragma GCC unroll 4
    for (int i = 0; i < 4; i++) {
        const uint64_t q = sbox7(t[i]);
        pr[16 + 2 * i] = (uint32_t)q;
        pr[17 + 2 * i] = q >> 32;
    }
    {
        const __m512i L = _mm512_and_si512(s0, lo32), H = _mm512_srli_epi64(s0, 32);
        _mm512_store_si512(pr, _mm512_unpacklo_epi64(L, H));      // pairs of lanes 0, 2, 4, 6
        _mm512_store_si512(pr + 8, _mm512_unpackhi_epi64(L, H));  // pairs of lanes 1, 3, 5, 7
    }
    SIPP_THROUGH_MEMORY(pr);
    const uint64_t* prm = pr;
    __m512i al[4], ah[4], ab[4];
    const __m512i zero = _mm512_setzero_si512();
#pragma GCC unroll 12
    for (int j = 0; j < 12; j++) {
        const uint64_t* q = prm + (j < 8 ? ((j & 1) ? 7 + j : j) : 2 * j);
        const __m512i bl = _mm512_set1_epi64((long long)q[0]), bh = _mm512_set1_epi64((long long)q[1]);
        // (the pairs of the scalar lanes are written by two 8-byte stores: a 16-byte load across them would not be forwarded)
        const __m512i bb = j < 8 ? _mm512_broadcast_i64x2(_mm_load_si128((const __m128i*)q)) : _mm512_mask_blend_epi64(0xAA, bl, bh);
        const __m512i ca = _mm512_load_si512(I.mds_icol_a[j]), cp = _mm512_load_si512(I.mds_icol_p[j]);
        al[j & 3] = _mm512_madd52lo_epu64(j == 0 ? _mm512_load_si512(rc_next) : j < 4 ? zero : al[j & 3], bl, ca);
        ah[j & 3] = _mm512_madd52lo_epu64(j == 0 ? _mm512_load_si512(rc_next + 8) : j < 4 ? zero : ah[j & 3], bh, ca);
        ab[j & 3] = _mm512_madd52lo_epu64(j == 0 ? _mm512_load_si512(rc_next + 16) : j < 4 ? zero : ab[j & 3], bb, cp);
    }
    // alo + 2^32 ahi = (alo + (ahi >> 32) (2^32 - 1)) + ((ahi mod 2^32) << 32): the first bracket is below 2^45, so the sum wraps only
    // when the low word of ahi is within 2^13 of 2^32 -- 2^-19 per lane: a cold branch, no compare on the chain
    auto combine = [&](__m512i alo, __m512i ahi) SIPP_IFMA {
        __m512i hi = _mm512_srli_epi64(ahi, 32);
        __m512i small = _mm512_add_epi64(alo, _mm512_sub_epi64(_mm512_slli_epi64(hi, 32), hi));
        __m512i r = _mm512_add_epi64(_mm512_slli_epi64(ahi, 32), small);
        __mmask8 c = _mm512_cmplt_epu64_mask(r, small);
        if (__builtin_expect(c != 0, 0)) r = v_carry_fix(r, c);
        return r;
    };
    s0 = combine(_mm512_add_epi64(_mm512_add_epi64(al[0], al[1]), _mm512_add_epi64(al[2], al[3])),
                 _mm512_add_epi64(_mm512_add_epi64(ah[0], ah[1]), _mm512_add_epi64(ah[2], ah[3])));
    // rows 8..11: (low, high) sums side by side; the high ones move one lane down and the same recombination leaves the rows in
    // the even lanes -- 7 vector instructions instead of 4 x 11 scalar ones
    const __m512i bi = _mm512_add_epi64(_mm512_add_epi64(ab[0], ab[1]), _mm512_add_epi64(ab[2], ab[3]));
    alignas(64) uint64_t o[8];
    _mm512_store_si512(o, combine(bi, _mm512_alignr_epi64(bi, bi, 1)));
    SIPP_THROUGH_MEMORY(o);
    const uint64_t* om = o;
#pragma GCC unroll 4
    for (int i = 0; i < 4; i++) t[i] = om[2 * i];
}
// all eight lanes of a block closed at once: (a0 + 2^52 a1 - 2^8 a2) mod p
SIPP_IFMA inline __m512i v_close(const IfmaBlock& A) {
    const __m512i one = _mm512_set1_epi64(1);
    __m512i t = _mm512_slli_epi64(A.a1, 52);
    __m512i lo = _mm512_add_epi64(A.a0, t);
    __mmask8 c = _mm512_cmplt_epu64_mask(lo, t);
    __m512i hi = _mm512_srli_epi64(A.a1, 12);
    hi = _mm512_mask_add_epi64(hi, c, hi, one);
    __m512i s = _mm512_slli_epi64(A.a2, 8);
    __mmask8 b = _mm512_cmplt_epu64_mask(lo, s);
    lo = _mm512_sub_epi64(lo, s);
    hi = _mm512_mask_sub_epi64(hi, b, hi, one);
    return v_reduce(lo, hi);
}

}  // namespace

SIPP_IFMA void poseidon_permute_ifma(uint64_t s[12], const PoseidonFastTables& T, const PoseidonIfmaTables& I) {
    __m512i s0 = _mm512_loadu_si512(s);
    uint64_t t[4];
    for (int i = 0; i < 4; i++) t[i] = s_add1(s[8 + i], T.rc_full[0][8 + i]);
    s0 = v_add_canon(s0, _mm512_load_si512(T.rc_full[0]));
    for (int k = 0; k < 4; k++) full_round_mixed(s0, t, I.rc_next[k][0], I);

    alignas(64) uint64_t y[32];  // (lanes 0..7 already carry `first`)  // y[0..11], y[16 + i] = y[i] >> 52
    _mm512_store_si512(y, s0);
    _mm512_store_si512(y + 16, _mm512_srli_epi64(s0, 52));
#pragma GCC unroll 4
    for (int i = 0; i < 4; i++) {
        y[8 + i] = t[i];  // (`first` rode along in the MDS layer of the fourth full round)
        y[24 + i] = t[i] >> 52;
    }
    // four named blocks selected by switch statements, never through a pointer or an array index: anything else keeps the twelve
    // accumulators in memory (a load and a store around every vpmadd52)
    IfmaBlock A0, A1, A2, A3;
#define BLK_DO(b, stmt) switch (b) { case 0: { IfmaBlock& B = A0; stmt; } break; case 1: { IfmaBlock& B = A1; stmt; } break; \
                                     case 2: { IfmaBlock& B = A2; stmt; } break; default: { IfmaBlock& B = A3; stmt; } break; }
#pragma GCC unroll 4
    for (int b = 0; b < 4; b++)
        BLK_DO(b, (B.a0 = _mm512_load_si512(I.acc_init[b][0]), B.a1 = _mm512_load_si512(I.acc_init[b][1]), B.a2 = _mm512_setzero_si512()))
    // the y part of a row block: 11 units; block 0 now, the others spread over the rounds that precede their first use
    SIPP_THROUGH_MEMORY(y);
    const uint64_t *ym = y, *ys = y;  // ys: the scalar readers (C-row 0) -- a pointer of their own, or GCC loads every y once into a
    asm("" : "+r"(ys));                // general register and broadcasts from there (port 5)
#define init_unit(i, b) BLK_DO(b, ifma_unit(B, _mm512_set1_epi64((long long)ym[1 + (i)]), _mm512_set1_epi64((long long)ym[17 + (i)]), I.init_c[i][b][0]))
#pragma GCC unroll 11
    for (int i = 0; i < 11; i++) init_unit(i, 0)
    uint64_t u0 = ys[0];
    uint64_t e;
    {
        Acc192 a = s_dot11_raw(I.row0, ys + 1);
        acc_add(a, I.k0);
        e = acc_reduce(a);
    }
    __m512i xb = _mm512_setzero_si512(), xh = xb;  // broadcasts of the previous round's S-box output
    alignas(64) uint64_t lanes[24];
#pragma GCC unroll 22
    for (int j = 0; j < 22; j++) {
        const uint64_t p7 = sbox7(u0);
        u0 = chain_close(e, p7);  // z_{j+1}
        if (j >= 1) {  // the vector terms of x_{j-1}: behind the chain of this round in program order, so the chain is served first
            const int k = j - 1;
            if (k <= 6) ifma_unit(A0, xb, xh, I.upd_c[k][0][0]);
            if (k <= 14) ifma_unit(A1, xb, xh, I.upd_c[k][1][0]);
            ifma_unit(A2, xb, xh, I.upd_c[k][2][0]);
            ifma_unit(A3, xb, xh, I.upd_c[k][3][0]);
        }
        {
            const int blk = j <= 5 ? 1 : (j >= 7 && j <= 12) ? 2 : (j >= 13 && j <= 18) ? 3 : 0;
            const int base = j <= 5 ? 0 : j <= 12 ? 7 : 13;
            if (blk) {
                const int i0 = 2 * (j - base);
                init_unit(i0, blk)
                if (i0 + 1 < 11) init_unit(i0 + 1, blk)
            }
        }
        if (j + 1 <= 21) {  // close C-row j + 1: vector terms k <= j - 1, newest term x_j
            const int row = j + 1, b = row <= 8 ? 0 : row <= 16 ? 1 : row <= 20 ? 2 : 3;
            const int lane = row <= 8 ? row - 1 : row <= 16 ? row - 9 : row <= 20 ? row - 13 : 0;
            uint64_t a0, a1, a2;
            BLK_DO(b, ifma_store(lanes, B))
            // (no barrier here: GCC turns this store + three loads into lane extracts, measured 9 ns per permutation faster)
            a0 = lanes[lane]; a1 = lanes[8 + lane]; a2 = lanes[16 + lane];
        }
        xb = _mm512_set1_epi64((long long)p7);
        xh = _mm512_srli_epi64(xb, 52);
    }
    {
        ifma_unit(A2, xb, xh, I.upd_c[21][2][0]);
        ifma_unit(A3, xb, xh, I.upd_c[21][3][0]);
    }
    s0 = _mm512_mask_set1_epi64(v_close(A3), 1, (long long)s_add1(s_mul(u0, I.lam22), T.rc_full[4][0]));  // (lanes 1..7: the constant is in the rows)
    {
        alignas(64) uint64_t o[8];
        _mm512_store_si512(o, v_close(A2));
        SIPP_THROUGH_MEMORY(o);
        const uint64_t* om = o;
        t[0] = om[0]; t[1] = om[1]; t[2] = om[2]; t[3] = om[3];
    }
    for (int k = 0; k < 4; k++) full_round_mixed(s0, t, I.rc_next[4 + k][0], I);
    _mm512_storeu_si512(s, v_canon(s0));
    for (int i = 0; i < 4; i++) s[8 + i] = t[i] - (t[i] >= GL_P ? GL_P : 0);
}
#undef BLK_DO
#undef init_unit
bool poseidon_ifma_supported() { return poseidon_avx512_supported() && __builtin_cpu_supports("avx512ifma"); }
SIPP_IFMA uint64_t poseidon_test_vmul_fast(uint64_t x, uint64_t y, int square) {
    alignas(64) uint64_t t[8];
    const __m512i vx = _mm512_set1_epi64((long long)x), vy = _mm512_set1_epi64((long long)y);
    _mm512_store_si512(t, square ? v_sqr_fast(vx) : v_mul_fast(vx, vy));
    return t[5];
}
SIPP_IFMA void poseidon_test_ifma_close(const uint64_t in[5], uint64_t out[2]) {
    out[0] = row_close(in[0], in[1], in[2], in[3], in[4]);
    IfmaBlock B;
    B.a0 = _mm512_set1_epi64((long long)in[0]);
    B.a1 = _mm512_set1_epi64((long long)in[1]);
    B.a2 = _mm512_set1_epi64((long long)in[2]);
    alignas(64) uint64_t t[8];
    _mm512_store_si512(t, v_close(B));
    out[1] = t[3];
}
namespace {
}  // namespace

SIPP_AVX512 void poseidon_permute_avx512(uint64_t s[12], const PoseidonFastTables& T) {
    alignas(64) uint64_t buf[16];
    memcpy(buf, s, 96);
    buf[12] = buf[13] = buf[14] = buf[15] = 0;
    __m512i s0 = _mm512_load_si512(buf), s1 = _mm512_load_si512(buf + 8);
    for (int k = 0; k < 4; k++) v_full_round(s0, s1, T.rc_full[k], T);

    // ---- 22 partial rounds, sparse form ----
    s0 = v_add_canon(s0, _mm512_load_si512(T.first));
    s1 = v_add_canon(s1, _mm512_load_si512(T.first + 8));
    _mm512_store_si512(buf, s0);
    _mm512_store_si512(buf + 8, s1);
    uint64_t u0 = buf[0];
    alignas(64) uint64_t ub[16];  // ub[i] = lane i (1..11); ub[0] unused
    {
        uint64_t zero = 0;
        for (int i = 0; i < 11; i++) ub[i + 1] = s_dot11p(T.init[i], buf + 1, zero, zero);
        ub[0] = 0; ub[12] = ub[13] = ub[14] = ub[15] = 0;
    }
    // d_r = vhat_r . U_r + m00 x_r with U_r = U_{r-1} + x_{r-1} w_{r-1}, so d_r = vhat_r . U_{r-1} + x_{r-1} (vhat_r . w_{r-1}) + m00 x_r:
    // the 11-term sum reads the state of ONE ROUND EARLIER (in memory long before it is needed), and the dependent chain of a
    // round is the S-box plus two multiply-adds.  Two buffers alternate: round r reads U_{r-1}, writes U_{r+1}.
    alignas(64) uint64_t um[2][16];
    memcpy(um[0], ub, sizeof ub);
    memcpy(um[1], ub, sizeof ub);  // round 0 reads "U_{-1}" = U_0 (its correction term has x_prev = 0)
    __m512i v0 = _mm512_load_si512(ub), v1 = _mm512_load_si512(ub + 8);
    uint64_t x_prev = 0;
    const uint64_t m00 = T.m00;
#pragma GCC unroll 2
    for (int r = 0; r < 22; r++) {
        const PartialRound* p = &T.pr[r];
        uint64_t* cur = um[(r + 1) & 1];  // holds U_{r-1}; receives U_{r+1} at the end of the round
        unsigned long long lo, hi, top;
        pr_dot(p, cur + 1, x_prev, lo, hi, top);
        uint64_t p7, x;
        pr_sbox(p, u0, p7, x);
        u0 = pr_finish(lo, hi, top, p7, m00);
        __m512i xb = _mm512_set1_epi64((long long)x);
        v0 = v_add_canon(v0, v_canon(v_mul(xb, _mm512_load_si512(p->w16))));
        v1 = v_add_canon(v1, v_canon(v_mul(xb, _mm512_load_si512(p->w16 + 8))));
        _mm512_store_si512(cur, v0);
        _mm512_store_si512(cur + 8, v1);
        x_prev = x;
    }
    _mm512_store_si512(ub, v0);
    _mm512_store_si512(ub + 8, v1);
    ub[0] = u0;
    s0 = _mm512_load_si512(ub);
    s1 = _mm512_load_si512(ub + 8);
    for (int k = 0; k < 4; k++) v_full_round(s0, s1, T.rc_full[4 + k], T);
    s0 = v_canon(s0);
    s1 = v_canon(s1);
    _mm512_store_si512(buf, s0);
    _mm512_store_si512(buf + 8, s1);
    memcpy(s, buf, 96);
}

// test hooks for the scalar helpers (their cold fix-up paths -- a borrow of probability 2^-32 -- are not reachable by random
// permutations; tests/test_host_cpu.py feeds them crafted operands)
SIPP_AVX512 uint64_t poseidon_test_red128(uint64_t lo, uint64_t hi) { return s_red128(lo, hi); }
SIPP_AVX512 uint64_t poseidon_test_finish(uint64_t lo, uint64_t hi, uint64_t top, uint64_t p7, uint64_t m00) { return pr_finish(lo, hi, top, p7, m00); }
SIPP_AVX512 uint64_t poseidon_test_sbox(uint64_t u, uint64_t post, uint64_t* x_out) {
    PartialRound pr;
    memset(&pr, 0, sizeof pr);
    pr.post = post;
    uint64_t p7, x;
    pr_sbox(&pr, u, p7, x);
    *x_out = x;
    return p7;
}

bool poseidon_avx512_supported() {
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
           __builtin_cpu_supports("bmi2");
}

}  // namespace sipp
#else
namespace sipp {
void poseidon_permute_avx512(uint64_t*, const PoseidonFastTables&) {}
void poseidon_permute_ifma(uint64_t*, const PoseidonFastTables&, const PoseidonIfmaTables&) {}
bool poseidon_ifma_supported() { return false; }
void poseidon_test_ifma_close(const uint64_t*, uint64_t*) {}
uint64_t poseidon_test_vmul_fast(uint64_t, uint64_t, int) { return 0; }
uint64_t poseidon_test_red128(uint64_t, uint64_t) { return 0; }
uint64_t poseidon_test_finish(uint64_t, uint64_t, uint64_t, uint64_t, uint64_t) { return 0; }
uint64_t poseidon_test_sbox(uint64_t, uint64_t, uint64_t*) { return 0; }
bool poseidon_avx512_supported() { return false; }
}  // namespace sipp
#endif
