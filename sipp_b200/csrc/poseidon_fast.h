// poseidon_fast.h -- tables shared by the portable (transcript.cc) and AVX-512 (poseidon_avx512.cc) Poseidon code.
#pragma once
#include <stdint.h>

namespace sipp {

// everything one sparse partial round reads besides the state, in one record addressed off a single pointer (the asm blocks of
// poseidon_avx512.cc use these byte offsets: w16 0, vhat 128, kprev 216, mpost 224, post 232)
struct alignas(64) PartialRound {
    uint64_t w16[16];    // w16[i] = w_r[i - 1] for i = 1..11, 0 elsewhere
    uint64_t vhat[11];
    uint64_t kprev;      // vhat[r] . w[r-1] (0 for r = 0)
    uint64_t mpost;      // m00 post[r]
    uint64_t post;
    uint64_t pad[2];
};
static_assert(sizeof(PartialRound) == 256, "PartialRound layout");

// Sparse-form partial rounds (derivation in transcript.cc) plus vector-friendly copies of the round constants.
struct PoseidonFastTables {
    alignas(64) uint64_t rc_full[8][16];   // round constants of the 4 + 4 full rounds, lanes 12..15 = 0
    alignas(64) uint64_t first[16];        // constants added before the first partial round
    alignas(64) uint64_t w16[22][16];      // w16[r][i] = w_r[i - 1] for i = 1..11, 0 elsewhere
    alignas(64) double mds_col_a[12][8];   // column form of the MDS layer: coefficient of s[j] in row r (r = 0..7) ...
    alignas(64) double mds_col_b[12][8];   // ... and in row 8 + (r & 3) (rows 8..11 twice: low sums in lanes 0..3, high sums in 4..7)
    uint64_t post[22];
    uint64_t mpost[22];
    uint64_t kprev[22];                    // kprev[r] = vhat[r] . w[r-1] (0 for r = 0)
    uint64_t init[11][11];
    uint64_t vhat[22][11];
    uint64_t m00;
    PartialRound pr[22];
};

// Partial rounds with the rank-1 updates UNROLLED algebraically (poseidon_avx512.cc, IFMA path).  With x_k the S-box output of round k
// and y the state entering the partial rounds, every later quantity is linear in (y, x_0, x_1, ...):
//     u0_{j+1} = m00 x_j + R_j,     R_j = (vhat_j init) . y + sum_{k<j} (vhat_j . w_k) x_k        ("C-row" j, j = 0..21)
//     U_22[i]  = init[i] . y + sum_k w_k[i] x_k                                                    ("U-row" i, i = 0..10)
// (post constants folded into per-row constants).  The chain variable is rescaled round by round, z_j = u0_j / lam_j with
// lam_{j+1} = m00 lam_j^7, so that z_{j+1} = z_j^7 + R_j / lam_{j+1}: the scalings sit in the coefficients, the dependent chain of a
// round is the S-box and ONE modular addition.  Rows live in the 64-bit lanes of four accumulator blocks, three registers each
// (weights 2^0, 2^52, 2^104, fed by vpmadd52luq / vpmadd52huq):
//     block 0: C-rows 1..8     block 1: C-rows 9..16     block 2: U-rows 7..10 | C-rows 17..20     block 3: C-row 21 | U-rows 0..6
// C-row 0 and the newest term (k = j - 1) of every C-row are scalar multiply-adds.
struct alignas(64) PoseidonIfmaTables {
    uint64_t acc_init[4][2][8];    // [block][weight 2^0 | 2^52][lane]: the row constant + p (so that the - 2^8 a2 correction cannot go negative)
    uint64_t init_c[11][4][2][8];  // [i][block][c | c >> 52][lane]: coefficient of y[1 + i]
    uint64_t upd_c[22][4][2][8];   // [k][block][c | c >> 52][lane]: coefficient of x_k (0 in C-rows j <= k + 1)
    uint64_t row0[11];             // C-row 0: vhat_0 init
    uint64_t k0;                   // its constant, m00 post_0
    uint64_t cdiag[22];            // cdiag[j]: coefficient of x_{j-1} in C-row j, j = 1..21
    uint64_t lam22;                // u0_22 = lam22 z_22
    alignas(64) uint64_t mds_icol_a[12][8];  // PoseidonFastTables::mds_col_a / _b as integers (the MDS layer on vpmadd52luq)
    alignas(64) uint64_t mds_icol_b[12][8];
    alignas(64) uint64_t mds_icol_p[12][8];
    alignas(64) uint64_t rc_next[8][3][8];   // the constants the NEXT layer adds, folded into the MDS accumulators of full round k (k = 3: `first`; k = 7:
                                              // nothing follows): [k][0 | 1][lane 0..7] low | high 32-bit halves for lanes 0..7; [k][2][2 i | 2 i + 1] the
                                              // (low, high) halves for lane 8 + i  // rows 8..11 for a (low half, high half) pair broadcast: lane l holds the entry of row 8 + l / 2
};

void poseidon_permute_avx512(uint64_t s[12], const PoseidonFastTables& T);
void poseidon_permute_ifma(uint64_t s[12], const PoseidonFastTables& T, const PoseidonIfmaTables& I);
bool poseidon_ifma_supported();
bool poseidon_avx512_supported();
// test hooks (poseidon_avx512.cc)
uint64_t poseidon_test_red128(uint64_t lo, uint64_t hi);
uint64_t poseidon_test_finish(uint64_t lo, uint64_t hi, uint64_t top, uint64_t p7, uint64_t m00);
uint64_t poseidon_test_sbox(uint64_t u, uint64_t post, uint64_t* x_out);
uint64_t poseidon_test_vmul_fast(uint64_t x, uint64_t y, int square);  // one lane of the IFMA path's vector product / square
void poseidon_test_ifma_close(const uint64_t in[5], uint64_t out[2]);  // out[0] = row_close(in...), out[1] = v_close lane of (in[0..2])

}  // namespace sipp
