// transcript.cc -- host-side Fiat-Shamir transcript of the SIPP native protocol.
//
// Mirrors /root/reference/src/transcript_native.rs:14-77 (`Transcript<GoldilocksField>`): a Poseidon hash chain
// (plonky2 @ 541e127 `hash_n_to_hash_no_pad::<F, PoseidonPermutation<F>>`) over the 32-bit limbs of the
// points / Fq12 elements, and `get_challenge` with num-bigint's zero-stripping `to_u32_digits`.
// The transcript is a strictly sequential chain and stays on the host (north star: "transcript_native ... stay as
// they are"); the prover overlaps the 8n-permutation absorb of A, B with the GPU's first products.
#include <string.h>

#include "../../include/sipp_b200.h"
#include "poseidon_rc.h"

namespace {

typedef unsigned __int128 u128;
const uint64_t GL_P = 0xFFFFFFFF00000001ull;
const uint64_t EPS = 0xFFFFFFFFull;  // 2^64 mod p

// values are kept as arbitrary u64 representatives mod p and canonicalised only on output
inline uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t r = a + b;
    if (r < a) {  // wrapped: 2^64 = EPS
        r += EPS;
        if (r < EPS) r += EPS;
    }
    return r;
}
inline uint64_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & EPS;
    uint64_t t = lo - hh;  // 2^96 = -1
    if (lo < hh) t -= EPS;
    uint64_t m = hl * EPS;  // 2^64 = 2^32 - 1
    uint64_t r = t + m;
    if (r < m) r += EPS;
    return r;
}
inline uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((u128)a * b); }
inline uint64_t gl_pow7(uint64_t x) {
    uint64_t x2 = gl_mul(x, x), x3 = gl_mul(x2, x), x4 = gl_mul(x2, x2);
    return gl_mul(x3, x4);
}
inline uint64_t gl_canon(uint64_t a) { return a >= GL_P ? a - GL_P : a; }

const uint64_t MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

// out[r] = sum_i s[(i + r) mod 12] * CIRC[i] + 8 s[0] [r == 0].  Split into 32-bit halves so that the twelve
// 12-term dot products stay inside u64 (12 * 41 * 2^32 < 2^42) and vectorise.
inline void mds_layer(uint64_t* s) {
    uint64_t lo[24], hi[24];
    for (int i = 0; i < 12; i++) {
        lo[i] = lo[i + 12] = s[i] & EPS;
        hi[i] = hi[i + 12] = s[i] >> 32;
    }
    uint64_t alo[12], ahi[12];
    for (int r = 0; r < 12; r++) {
        uint64_t a = 0, b = 0;
        for (int i = 0; i < 12; i++) {
            a += lo[i + r] * MDS_CIRC[i];
            b += hi[i + r] * MDS_CIRC[i];
        }
        alo[r] = a;
        ahi[r] = b;
    }
    alo[0] += lo[0] * 8;
    ahi[0] += hi[0] * 8;
    for (int r = 0; r < 12; r++) {
        // value = alo + ahi * 2^32 < 2^75
        u128 v = (u128)alo[r] + ((u128)ahi[r] << 32);
        uint64_t l = (uint64_t)v, h = (uint64_t)(v >> 64);  // h < 2^11
        uint64_t m = h * EPS;
        uint64_t x = l + m;
        if (x < m) x += EPS;
        s[r] = x;
    }
}

}  // namespace

extern "C" {

void sipp_poseidon_permute(uint64_t s[12]) {
    int rnd = 0;
    for (int k = 0; k < 4; k++, rnd++) {
        for (int i = 0; i < 12; i++) s[i] = gl_pow7(gl_add(s[i], SIPP_POSEIDON_RC[12 * rnd + i]));
        mds_layer(s);
    }
    for (int k = 0; k < 22; k++, rnd++) {
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], SIPP_POSEIDON_RC[12 * rnd + i]);
        s[0] = gl_pow7(s[0]);
        mds_layer(s);
    }
    for (int k = 0; k < 4; k++, rnd++) {
        for (int i = 0; i < 12; i++) s[i] = gl_pow7(gl_add(s[i], SIPP_POSEIDON_RC[12 * rnd + i]));
        mds_layer(s);
    }
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}

void sipp_transcript_new(sipp_transcript* t) { memset(t, 0, sizeof *t); }

// state <- hash_n_to_hash_no_pad(state || msg): overwrite-mode sponge, rate 8, no padding
void sipp_transcript_append(sipp_transcript* t, const uint64_t* msg, size_t n) {
    uint64_t s[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    // first chunk: 4 state elements + up to 4 message elements
    size_t first = n < 4 ? n : 4;
    memcpy(s, t->state, 32);
    memcpy(s + 4, msg, first * 8);
    sipp_poseidon_permute(s);
    for (size_t off = first; off < n; off += 8) {
        size_t len = n - off < 8 ? n - off : 8;
        memcpy(s, msg + off, len * 8);
        sipp_poseidon_permute(s);
    }
    memcpy(t->state, s, 32);
}

static inline void fq_limbs(uint64_t* out, const uint8_t* b) {  // from_fq_to_f: 8 little-endian u32 digits
    uint32_t w[8];
    memcpy(w, b, 32);
    for (int i = 0; i < 8; i++) out[i] = w[i];
}

void sipp_transcript_append_g1(sipp_transcript* t, const uint8_t a[64]) {
    uint64_t m[16];
    fq_limbs(m, a);
    fq_limbs(m + 8, a + 32);
    sipp_transcript_append(t, m, 16);
}

void sipp_transcript_append_g2(sipp_transcript* t, const uint8_t b[128]) {
    uint64_t m[32];
    for (int i = 0; i < 4; i++) fq_limbs(m + 8 * i, b + 32 * i);
    sipp_transcript_append(t, m, 32);
}

void sipp_transcript_append_fq12(sipp_transcript* t, const uint8_t f[384]) {
    uint64_t m[96];
    if (sipp_get_option(SIPP_OPT_FQ12_ORDER) == 1) {
        for (int i = 0; i < 12; i++) fq_limbs(m + 8 * i, f + 32 * i);
    } else {
        // MyFq12.coeffs[i] = g_i.c0, coeffs[i + 6] = g_i.c1 where g_i is the Fq2 coefficient of w^i; in the nested
        // byte order g_i sits in slot (i & 1) * 3 + (i >> 1)
        for (int i = 0; i < 6; i++) {
            int slot = (i & 1) * 3 + (i >> 1);
            fq_limbs(m + 8 * i, f + 64 * slot);
            fq_limbs(m + 8 * (i + 6), f + 64 * slot + 32);
        }
    }
    sipp_transcript_append(t, m, 96);
}

void sipp_transcript_append_pairs(sipp_transcript* t, const uint8_t* A, const uint8_t* B, size_t n) {
    for (size_t i = 0; i < n; i++) {
        sipp_transcript_append_g1(t, A + 64 * i);
        sipp_transcript_append_g2(t, B + 128 * i);
    }
}

void sipp_transcript_get_challenge(const sipp_transcript* t, uint8_t x[32]) {
    uint64_t s[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    memcpy(s, t->state, 32);
    sipp_poseidon_permute(s);
    // concatenate the base-2^32 digits of the four digest elements, high zero digits stripped (0 -> none)
    uint32_t digits[8];
    int nd = 0;
    for (int k = 0; k < 4; k++) {
        uint64_t d = s[k];
        while (d) {
            digits[nd++] = (uint32_t)d;
            d >>= 32;
        }
    }
    uint64_t v[4] = {0, 0, 0, 0};
    for (int j = 0; j < nd; j++) v[j >> 1] |= (uint64_t)digits[j] << (32 * (j & 1));
    // reduce mod r (v < 2^256 < 6r)
    static const uint64_t RM[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    for (;;) {
        bool ge = true;
        for (int i = 3; i >= 0; i--) {
            if (v[i] != RM[i]) {
                ge = v[i] > RM[i];
                break;
            }
        }
        if (!ge) break;
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)v[i] - RM[i] - borrow;
            v[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    memcpy(x, v, 32);
}

}  // extern "C"
