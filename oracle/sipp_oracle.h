/*
 * sipp_oracle.h -- CPU oracle for the SIPP native prover hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (sipp_b200/) never links, imports or falls back to it.
 *
 * PARITY UNPINNED against the real reference: qope/SIPP holds no golden vectors for this path and
 * cannot be built here (no Rust toolchain; plonky2-bn254-pairing @ fe5c3a8, plonky2-bn254 @ d616d57,
 * plonky2 @ 541e127 and ark-* 0.4 are not vendored).  The oracle restates their published
 * algorithms and is pinned to: upstream plonky2 Poseidon KATs, the
 * self-derived digests of SURVEY.md Appendix D, and the independent pure-Python model under
 * tests/golden/ (fixtures committed there).
 *
 * Byte formats (shared with include/sipp_b200.h):
 *   Fq   : 32 bytes, little-endian canonical integer in [0,p)            (ark-serialize Fq)
 *   G1   : x || y                      = 64 bytes;  identity = all zero  (ark "infinity" has x=y=0)
 *   G2   : x.c0 || x.c1 || y.c0 || y.c1 = 128 bytes; identity = all zero
 *   Fq12 : 12 Fq in arkworks nested order c0.c0.c0, c0.c0.c1, c0.c1.c0 ... c1.c2.c1 = 384 bytes
 *   Fr   : 32 bytes little-endian canonical
 */
#ifndef SIPP_ORACLE_H
#define SIPP_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* option bits */
#define SIPP_ORACLE_FE_ARK      1u /* final exponentiation normalised as arkworks (exact ^ 2x(6x^2+3x+1)); default exact (H1) */
#define SIPP_ORACLE_FQ12_NESTED 2u /* transcript absorbs Fq12 in nested order instead of MyFq12 w-basis (H2) */
#define SIPP_ORACLE_FAITHFUL    4u /* inner_product = one full pairing per pair (prover_native.rs:17-22); default: product of
                                      Miller loops + one final exponentiation (same value) */

/* field / tower ops on canonical bytes, for per-op parity tests.  op codes: */
enum { ORC_FQ_MUL = 0, ORC_FQ_ADD, ORC_FQ_SUB, ORC_FQ_INV, ORC_FQ_SQR,
       ORC_FQ2_MUL = 10, ORC_FQ2_SQR, ORC_FQ2_INV, ORC_FQ2_MUL_XI,
       ORC_FQ12_MUL = 20, ORC_FQ12_SQR, ORC_FQ12_INV, ORC_FQ12_FROB1, ORC_FQ12_FROB2, ORC_FQ12_FROB3,
       ORC_FQ12_CONJ, ORC_FQ12_CYC_SQR /* input must be in the cyclotomic subgroup */ };
/* out[i] = op(a[i], b[i]); element size is implied by op (32 / 64 / 384 bytes); b may be NULL for unary ops */
int oracle_field_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t count);

/* curve ops */
int oracle_g1_mul(const uint8_t a[64], const uint8_t k[32], uint8_t out[64]);
int oracle_g2_mul(const uint8_t b[128], const uint8_t k[32], uint8_t out[128]);
int oracle_g1_on_curve(const uint8_t a[64]);
int oracle_g2_on_curve(const uint8_t b[128]);
/* A'_i = A_i + x * A_{i+h}  (prover_native.rs:60-64) and B'_i = B_i + xinv * B_{i+h} (:65-69), h = n/2 */
int oracle_fold_g1(const uint8_t *A, size_t n, const uint8_t x[32], uint8_t *out);
int oracle_fold_g2(const uint8_t *B, size_t n, const uint8_t xinv[32], uint8_t *out);
int oracle_fr_inverse(const uint8_t x[32], uint8_t out[32]);

/* pairing (plonky2_bn254_pairing::pairing) */
int oracle_pairing(const uint8_t a[64], const uint8_t b[128], uint8_t out[384], unsigned opts);
int oracle_miller_loop(const uint8_t a[64], const uint8_t b[128], uint8_t out[384]);
int oracle_final_exp(const uint8_t f[384], uint8_t out[384], unsigned opts);
/* prover_native.rs:15-23 */
int oracle_inner_product(const uint8_t *A, const uint8_t *B, size_t n, uint8_t out[384], unsigned opts);
/* multithreaded variant for the "generous" CPU baseline (pthreads) */
int oracle_inner_product_mt(const uint8_t *A, const uint8_t *B, size_t n, uint8_t out[384], unsigned opts, int threads);

/* Poseidon / transcript (transcript_native.rs:14-77) */
void oracle_poseidon_perm(uint64_t state[12]);
void oracle_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]);
void oracle_poseidon_round_constants(uint64_t out[360]);
typedef struct { uint64_t state[4]; uint64_t perms; } oracle_transcript;
void oracle_transcript_new(oracle_transcript *t);
void oracle_transcript_append(oracle_transcript *t, const uint64_t *msg, size_t n);
void oracle_transcript_append_g1(oracle_transcript *t, const uint8_t a[64]);
void oracle_transcript_append_g2(oracle_transcript *t, const uint8_t b[128]);
void oracle_transcript_append_fq12(oracle_transcript *t, const uint8_t f[384], unsigned opts);
void oracle_transcript_get_challenge(const oracle_transcript *t, uint8_t x[32]);
/* challenge from an explicit 4-element digest (exercises the zero-limb-stripping quirk) */
void oracle_challenge_from_digest(const uint64_t digest[4], uint8_t x[32]);

/* prover_native.rs:26-80.  proof_out: (2 log2 n + 1) * 384 bytes, in the returned (reversed) order.
 * Optional trace outputs (may be NULL): challenges[log2 n][32]; foldedA / foldedB receive the folded
 * vectors of every round concatenated (n/2 + n/4 + ... + 1 = n-1 points).  threads>1 parallelises
 * pairs/folds with pthreads (generous baseline); threads<=1 is the single-thread structure of the reference. */
int oracle_sipp_prove(const uint8_t *A, const uint8_t *B, size_t n, uint8_t *proof_out, unsigned opts, int threads,
                      uint8_t *challenges, uint8_t *foldedA, uint8_t *foldedB);
/* verifier_native.rs:14-85.  returns 1 = Ok(statement), 0 = Err("Verification failed"), <0 = malformed.
 * statement outputs (may be NULL): final_A[64], final_B[128], final_Z[384]. */
int oracle_sipp_verify(const uint8_t *A, const uint8_t *B, size_t n, const uint8_t *proof, size_t proof_len,
                       unsigned opts, int threads, uint8_t *final_A, uint8_t *final_B, uint8_t *final_Z);

/* documented seeded input generator: SplitMix64(seed) -> 4 words -> 256-bit LE -> mod r (0 -> 1);
 * scalars interleaved a_0, b_0, a_1, b_1, ...;  A_i = [a_i]G1gen, B_i = [b_i]G2gen */
void oracle_seeded_scalars(uint64_t seed, size_t n, uint8_t *scalars /* 2n x 32 */);
int oracle_seeded_inputs(uint64_t seed, size_t n, uint8_t *A, uint8_t *B, int threads);

#ifdef __cplusplus
}
#endif
#endif
