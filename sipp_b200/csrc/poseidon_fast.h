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

void poseidon_permute_avx512(uint64_t s[12], const PoseidonFastTables& T);
bool poseidon_avx512_supported();
// test hooks (poseidon_avx512.cc)
uint64_t poseidon_test_red128(uint64_t lo, uint64_t hi);
uint64_t poseidon_test_finish(uint64_t lo, uint64_t hi, uint64_t top, uint64_t p7, uint64_t m00);
uint64_t poseidon_test_sbox(uint64_t u, uint64_t post, uint64_t* x_out);

}  // namespace sipp
