// statement.cc -- hand-off of the native verifier's result to the (untouched) circuit side: the vector of 32-bit limbs that
// `SIPPStatement::from_vec` parses (/root/reference/src/statements.rs:133-170) and `SIPPStatementTarget::to_vec` emits as the
// plonky2 public inputs (statements.rs:24-39; compared in verifier_circuit.rs:255-268).  Host code, no GPU involved.
//
// Layout (8 little-endian u32 limbs per Fq, statements.rs:90-131):
//   A (n x 16) | B (n x 32) | Z (96) | final_A (16) | final_B (32) | final_Z (96)
// G1 = x | y and G2 = x.c0 | x.c1 | y.c0 | y.c1 are byte-identical to the canonical little-endian boundary format.  An Fq12 is
// `MyFq12.coeffs` (statements.rs:120-131): coeffs[i] = g_i.c0, coeffs[i + 6] = g_i.c1 where g_i is the Fq2 coefficient of w^i
// (hypothesis H2; SIPP_OPT_FQ12_ORDER = 1 switches to the plain nested order, as for the transcript).
#include <string.h>

#include "../../include/sipp_b200.h"

namespace sipp_host {
int fail(int code, const char* what);
}
using sipp_host::fail;

namespace {

const uint32_t P_LIMBS[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};

bool fq_canonical(const uint32_t* w) {
    for (int i = 7; i >= 0; i--) {
        if (w[i] != P_LIMBS[i]) return w[i] < P_LIMBS[i];
    }
    return false;  // == p
}

// nested slot of the Fq2 coefficient of w^i: g_0..g_5 = c0.c0, c1.c0, c0.c1, c1.c1, c0.c2, c1.c2
inline int w_slot(int i) { return (i & 1) * 3 + (i >> 1); }

void fq12_to_limbs(const uint8_t f[384], uint32_t* out) {
    if (sipp_get_option(SIPP_OPT_FQ12_ORDER) == 1) {
        memcpy(out, f, 384);
        return;
    }
    for (int i = 0; i < 6; i++) {
        memcpy(out + 8 * i, f + 64 * w_slot(i), 32);
        memcpy(out + 8 * (i + 6), f + 64 * w_slot(i) + 32, 32);
    }
}

void limbs_to_fq12(const uint32_t* in, uint8_t f[384]) {
    if (sipp_get_option(SIPP_OPT_FQ12_ORDER) == 1) {
        memcpy(f, in, 384);
        return;
    }
    for (int i = 0; i < 6; i++) {
        memcpy(f + 64 * w_slot(i), in + 8 * i, 32);
        memcpy(f + 64 * w_slot(i) + 32, in + 8 * (i + 6), 32);
    }
}

}  // namespace

extern "C" {

size_t sipp_statement_u32_len(size_t n) { return 16 * n + 32 * n + 96 + 16 + 32 + 96; }

int sipp_statement_to_u32(const uint8_t* A, const uint8_t* B, size_t n, const uint8_t Z[384], const uint8_t final_A[64], const uint8_t final_B[128],
                          const uint8_t final_Z[384], uint32_t* out, size_t out_len) {
    if (!A || !B || !Z || !final_A || !final_B || !final_Z || !out) return fail(SIPP_ERR_ARG, "null argument");
    if (out_len != sipp_statement_u32_len(n)) return fail(SIPP_ERR_LENGTH, "assert!(input.len() == total_len)");  // statements.rs:139
    uint32_t* o = out;
    memcpy(o, A, 64 * n); o += 16 * n;
    memcpy(o, B, 128 * n); o += 32 * n;
    fq12_to_limbs(Z, o); o += 96;
    memcpy(o, final_A, 64); o += 16;
    memcpy(o, final_B, 128); o += 32;
    fq12_to_limbs(final_Z, o);
    return SIPP_OK;
}

int sipp_statement_from_u32(size_t n, const uint32_t* in, size_t in_len, uint8_t* A, uint8_t* B, uint8_t Z[384], uint8_t final_A[64],
                            uint8_t final_B[128], uint8_t final_Z[384]) {
    if (!in || !A || !B || !Z || !final_A || !final_B || !final_Z) return fail(SIPP_ERR_ARG, "null argument");
    if (in_len != sipp_statement_u32_len(n)) return fail(SIPP_ERR_LENGTH, "assert!(input.len() == total_len)");  // statements.rs:139
    // every group of 8 limbs is one Fq: the boundary format is canonical, so a value >= p is refused (the reference's
    // `BigUint -> Fq` would reduce it; the circuit never emits one)
    for (size_t i = 0; i < in_len; i += 8)
        if (!fq_canonical(in + i)) return fail(SIPP_ERR_ENCODING, "statement limb group >= p");
    const uint32_t* p = in;
    memcpy(A, p, 64 * n); p += 16 * n;
    memcpy(B, p, 128 * n); p += 32 * n;
    limbs_to_fq12(p, Z); p += 96;
    memcpy(final_A, p, 64); p += 16;
    memcpy(final_B, p, 128); p += 32;
    limbs_to_fq12(p, final_Z);
    return SIPP_OK;
}

}  // extern "C"
