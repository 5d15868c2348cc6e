// poseidon_avx512.cc -- AVX-512 implementation of the plonky2 Poseidon permutation over Goldilocks (width 12).
//
// The Fiat-Shamir transcript of the SIPP native protocol (/root/reference/src/transcript_native.rs:25-30) is a strictly
// sequential chain of 8n + 13 + 27 log2(n) permutations (prover_native.rs:36-39 absorbs every A_i, B_i) that must stay on
// the host; at the sizes the GPU finishes in milliseconds this chain IS the prove time, so one permutation has to be as
// short as the machine allows.  This file computes exactly the same function as the portable code in transcript.cc
// (selected at run time when the CPU has AVX-512 F/DQ/VL + BMI2):
//   full rounds    state in two zmm registers (lanes 0..7, 8..11); x^7 with 4 x vpmuludq 64x64->128 products and the
//                  2^64 = 2^32 - 1, 2^96 = -1 reduction; the circulant MDS layer as 48 FP64 FMAs on the 32-bit halves
//                  (sums < 2^43 are exact in double), rotations done as unaligned loads of a twice-stored copy
//   partial rounds sparse form (tables derived in transcript.cc): the lane-0 S-box and the 11-term dot product run on the
//                  scalar ports (mulx / adc, 192-bit lazy accumulation) while the rank-1 update of lanes 1..11 runs on
//                  the vector ports
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
SIPP_AVX512 inline uint64_t s_red128(uint64_t lo, uint64_t hi) {
    uint64_t hh = hi >> 32, hl = hi & EPS;
    uint64_t t = lo - hh;
    if (__builtin_expect(lo < hh, 0)) t -= EPS;
    uint64_t m = (hl << 32) - hl;
    uint64_t r = t + m;
    r += (0 - (uint64_t)(r < m)) & EPS;
    return r;
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
// sum_{i<11} a[i] * b[i] + extra_a * extra_b, reduced once (three-limb lazy accumulation; 2^128 = -2^32 mod p)
SIPP_AVX512 inline uint64_t s_dot11p(const uint64_t* a, const uint64_t* b, uint64_t ea, uint64_t eb) {
    unsigned long long lo, hi, top = 0, pl, ph;
    lo = _mulx_u64(ea, eb, &hi);
#pragma GCC unroll 11
    for (int i = 0; i < 11; i++) {
        pl = _mulx_u64(a[i], b[i], &ph);
        unsigned char c = _addcarry_u64(0, lo, pl, &lo);
        c = _addcarry_u64(c, hi, ph, &hi);
        top += c;
    }
    uint64_t r = s_red128(lo, hi);
    uint64_t t = (uint64_t)top << 32;  // top <= 11
    uint64_t d = r - t;
    if (__builtin_expect(r < t, 0)) d -= EPS;  // borrowed 2^64 = EPS
    return d;
}

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
SIPP_AVX512 inline __m512i v_mul(__m512i x, __m512i y) {
    const __m512i lo32 = _mm512_set1_epi64((long long)EPS);
    __m512i xh = _mm512_srli_epi64(x, 32), yh = _mm512_srli_epi64(y, 32);
    __m512i ll = _mm512_mul_epu32(x, y), lh = _mm512_mul_epu32(x, yh), hl = _mm512_mul_epu32(xh, y), hh = _mm512_mul_epu32(xh, yh);
    __m512i t0 = _mm512_add_epi64(hl, _mm512_srli_epi64(ll, 32));
    __m512i t1 = _mm512_add_epi64(lh, _mm512_and_si512(t0, lo32));
    __m512i hi = _mm512_add_epi64(hh, _mm512_add_epi64(_mm512_srli_epi64(t0, 32), _mm512_srli_epi64(t1, 32)));
    __m512i lo = _mm512_or_si512(_mm512_and_si512(ll, lo32), _mm512_slli_epi64(t1, 32));
    return v_reduce(lo, hi);
}
SIPP_AVX512 inline __m512i v_sqr(__m512i x) {
    const __m512i lo32 = _mm512_set1_epi64((long long)EPS);
    __m512i xh = _mm512_srli_epi64(x, 32);
    __m512i ll = _mm512_mul_epu32(x, x), lh = _mm512_mul_epu32(x, xh), hh = _mm512_mul_epu32(xh, xh);
    __m512i t0 = _mm512_add_epi64(lh, _mm512_srli_epi64(ll, 32));
    __m512i t1 = _mm512_add_epi64(lh, _mm512_and_si512(t0, lo32));
    __m512i hi = _mm512_add_epi64(hh, _mm512_add_epi64(_mm512_srli_epi64(t0, 32), _mm512_srli_epi64(t1, 32)));
    __m512i lo = _mm512_or_si512(_mm512_and_si512(ll, lo32), _mm512_slli_epi64(t1, 32));
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

// out[r] = sum_i s[(i + r) mod 12] * CIRC[i] + 8 s[0] [r == 0] on the 32-bit halves, in FP64
SIPP_AVX512 inline void v_mds(__m512i& s0, __m512i& s1, const PoseidonFastTables& T) {
    const __m512i lo32 = _mm512_set1_epi64((long long)EPS);
    alignas(64) double dl[32], dh[32];
    __m512d l0 = _mm512_cvtepu64_pd(_mm512_and_si512(s0, lo32)), l1 = _mm512_cvtepu64_pd(_mm512_and_si512(s1, lo32));
    __m512d h0 = _mm512_cvtepu64_pd(_mm512_srli_epi64(s0, 32)), h1 = _mm512_cvtepu64_pd(_mm512_srli_epi64(s1, 32));
    // s0 | s1 (4 live lanes) | s0 | s1: the second copy of s0 overwrites the dead lanes of the first s1
    _mm512_store_pd(dl, l0); _mm512_store_pd(dl + 8, l1); _mm512_storeu_pd(dl + 12, l0); _mm512_storeu_pd(dl + 20, l1);
    _mm512_store_pd(dh, h0); _mm512_store_pd(dh + 8, h1); _mm512_storeu_pd(dh + 12, h0); _mm512_storeu_pd(dh + 20, h1);
    __m512d al0 = _mm512_mul_pd(l0, _mm512_load_pd(T.mds_c0a)), al1 = _mm512_setzero_pd();
    __m512d ah0 = _mm512_mul_pd(h0, _mm512_load_pd(T.mds_c0a)), ah1 = _mm512_setzero_pd();
    __m512d bl0 = _mm512_setzero_pd(), bl1 = _mm512_setzero_pd(), bh0 = _mm512_setzero_pd(), bh1 = _mm512_setzero_pd();
#pragma GCC unroll 12
    for (int i = 0; i < 12; i++) {
        const __m512d c = _mm512_set1_pd(T.mds_circ[i]);
        if (i > 0) {
            if (i & 1) {
                al1 = _mm512_fmadd_pd(_mm512_loadu_pd(dl + i), c, al1);
                ah1 = _mm512_fmadd_pd(_mm512_loadu_pd(dh + i), c, ah1);
            } else {
                al0 = _mm512_fmadd_pd(_mm512_loadu_pd(dl + i), c, al0);
                ah0 = _mm512_fmadd_pd(_mm512_loadu_pd(dh + i), c, ah0);
            }
        }
        if (i & 1) {
            bl1 = _mm512_fmadd_pd(_mm512_loadu_pd(dl + i + 8), c, bl1);
            bh1 = _mm512_fmadd_pd(_mm512_loadu_pd(dh + i + 8), c, bh1);
        } else {
            bl0 = _mm512_fmadd_pd(_mm512_loadu_pd(dl + i + 8), c, bl0);
            bh0 = _mm512_fmadd_pd(_mm512_loadu_pd(dh + i + 8), c, bh0);
        }
    }
    const __m512i eps = lo32;
    auto combine = [&](__m512d lo_d, __m512d hi_d) SIPP_AVX512 {
        __m512i alo = _mm512_cvtpd_epu64(lo_d), ahi = _mm512_cvtpd_epu64(hi_d);  // < 2^43 each; value = alo + 2^32 ahi
        __m512i lo = _mm512_add_epi64(alo, _mm512_slli_epi64(ahi, 32));
        __mmask8 c = _mm512_cmplt_epu64_mask(lo, alo);
        __m512i hi = _mm512_srli_epi64(ahi, 32);
        hi = _mm512_mask_add_epi64(hi, c, hi, _mm512_set1_epi64(1));
        __m512i m = _mm512_sub_epi64(_mm512_slli_epi64(hi, 32), hi);  // hi * (2^32 - 1), hi < 2^12
        __m512i r = _mm512_add_epi64(lo, m);
        __mmask8 c2 = _mm512_cmplt_epu64_mask(r, m);
        return _mm512_mask_add_epi64(r, c2, r, eps);
    };
    s0 = combine(_mm512_add_pd(al0, al1), _mm512_add_pd(ah0, ah1));
    s1 = combine(_mm512_add_pd(bl0, bl1), _mm512_add_pd(bh0, bh1));
}

SIPP_AVX512 inline void v_full_round(__m512i& s0, __m512i& s1, const uint64_t* rc16, const PoseidonFastTables& T) {
    s0 = v_pow7(v_add_canon(s0, _mm512_load_si512(rc16)));
    s1 = v_pow7(v_add_canon(s1, _mm512_load_si512(rc16 + 8)));
    v_mds(s0, s1, T);
}

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
    __m512i v0 = _mm512_load_si512(ub), v1 = _mm512_load_si512(ub + 8);
    for (int r = 0; r < 22; r++) {
        uint64_t x = s_add(s_pow7(u0), T.post[r]);
        // d = m00 x + vhat . u ; u <- u + x w
        uint64_t d = s_dot11p(T.vhat[r], ub + 1, x, T.m00);
        __m512i xb = _mm512_set1_epi64((long long)x);
        v0 = v_add_canon(v0, v_canon(v_mul(xb, _mm512_load_si512(T.w16[r]))));
        v1 = v_add_canon(v1, v_canon(v_mul(xb, _mm512_load_si512(T.w16[r] + 8))));
        _mm512_store_si512(ub, v0);
        _mm512_store_si512(ub + 8, v1);
        u0 = d;
    }
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

bool poseidon_avx512_supported() {
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
           __builtin_cpu_supports("bmi2");
}

}  // namespace sipp
#else
namespace sipp {
void poseidon_permute_avx512(uint64_t*, const PoseidonFastTables&) {}
bool poseidon_avx512_supported() { return false; }
}  // namespace sipp
#endif
