// transcript.cc -- host-side Fiat-Shamir transcript of the SIPP native protocol.
//
// Mirrors /root/reference/src/transcript_native.rs:14-77 (`Transcript<GoldilocksField>`): a Poseidon hash chain
// (plonky2 @ 541e127 `hash_n_to_hash_no_pad::<F, PoseidonPermutation<F>>`) over the 32-bit limbs of the
// points / Fq12 elements, and `get_challenge` with num-bigint's zero-stripping `to_u32_digits`.
// The transcript is a strictly sequential chain and stays on the host (north star: "transcript_native ... stay as
// they are"); the prover overlaps the 8n-permutation absorb of A, B with the GPU's first products.
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include "../../include/sipp_b200.h"
#include "poseidon_fast.h"
#include "poseidon_rc.h"

namespace {

typedef unsigned __int128 u128;
const uint64_t GL_P = 0xFFFFFFFF00000001ull;
const uint64_t EPS = 0xFFFFFFFFull;  // 2^64 mod p

// values are kept as arbitrary u64 representatives mod p and canonicalised only on output.  All helpers are
// branch-free (data-dependent branches on field values mispredict half the time).
inline uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t r = a + b;
    uint64_t c = r < a;             // wrapped: 2^64 = EPS (mod p)
    uint64_t t = r + ((0 - c) & EPS);
    uint64_t c2 = t < r;
    return t + ((0 - c2) & EPS);
}
inline uint64_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & EPS;
    uint64_t t = lo - hh;           // 2^96 = -1
    t -= (0 - (uint64_t)(lo < hh)) & EPS;
    uint64_t m = hl * EPS;          // 2^64 = 2^32 - 1
    uint64_t r = t + m;
    r += (0 - (uint64_t)(r < m)) & EPS;
    return r;
}
inline uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((u128)a * b); }
inline uint64_t gl_pow7(uint64_t x) {
    uint64_t x2 = gl_mul(x, x), x3 = gl_mul(x2, x), x4 = gl_mul(x2, x2);
    return gl_mul(x3, x4);
}
inline uint64_t gl_canon(uint64_t a) { return a - ((0 - (uint64_t)(a >= GL_P)) & GL_P); }

const uint64_t MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

// out[r] = sum_i s[(i + r) mod 12] * CIRC[i] + 8 s[0] [r == 0].  Split into 32-bit halves so that the twelve
// 12-term dot products stay inside u64 (12 * 41 * 2^32 < 2^42); the inner loops run over r with unit stride so the
// compiler vectorises them (AVX2: 3 vectors per 12 lanes).
inline void mds_layer(uint64_t* __restrict__ s) {
    alignas(32) uint64_t lo[24], hi[24], alo[12], ahi[12];
    for (int i = 0; i < 12; i++) {
        lo[i] = lo[i + 12] = s[i] & EPS;
        hi[i] = hi[i + 12] = s[i] >> 32;
        alo[i] = 0;
        ahi[i] = 0;
    }
    for (int i = 0; i < 12; i++) {
        const uint64_t c = MDS_CIRC[i];
        for (int r = 0; r < 12; r++) {
            alo[r] += lo[i + r] * c;
            ahi[r] += hi[i + r] * c;
        }
    }
    alo[0] += lo[0] * 8;
    ahi[0] += hi[0] * 8;
    for (int r = 0; r < 12; r++) {
        // value = alo + ahi * 2^32 < 2^75
        u128 v = (u128)alo[r] + ((u128)ahi[r] << 32);
        uint64_t l = (uint64_t)v, h = (uint64_t)(v >> 64);  // h < 2^11
        uint64_t m = h * EPS;
        uint64_t x = l + m;
        x += (0 - (uint64_t)(x < m)) & EPS;
        s[r] = x;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Partial rounds in "sparse matrix" form (Poseidon paper, appendix on equivalent matrices; plonky2 uses the same
// trick).  With M = [[m00, v],[w, Mh]] and the S-box acting on lane 0 only, the 22 rounds
//      x <- M * Sbox0(x + c_t)
// are rewritten, exactly, as
//      u  = diag(1, Minit) * (x + first)
//      u <- Msparse_t * (Sbox0(u) + post_t e0),   Msparse_t = [[m00, vhat_t], [w_t, I]]
// where `first`, `post`, `Minit`, `vhat_t`, `w_t` are derived once at start-up from the MDS matrix and the round
// constants by 11x11 / 12x12 inversions over Goldilocks.  Each sparse round costs 22 field multiplications instead
// of a 12x12 matrix-vector product.
// ---------------------------------------------------------------------------------------------------------
struct FastPartial {
    uint64_t first[12];
    uint64_t post[22];
    uint64_t init[11][11];  // u_i = sum_j init[i][j] y_j over lanes 1..11
    uint64_t vhat[22][11];
    uint64_t w[22][11];
    uint64_t m00;
};
FastPartial g_fp;

uint64_t gl_sub(uint64_t a, uint64_t b) { return gl_add(gl_canon(a), GL_P - gl_canon(b)); }
uint64_t gl_inv(uint64_t a) {  // a^(p-2)
    uint64_t e = GL_P - 2, r = 1, b = gl_canon(a);
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_mul(b, b);
        e >>= 1;
    }
    return gl_canon(r);
}
// inverse of an n x n matrix (row-major, n <= 12) by Gauss-Jordan elimination
void mat_inverse(const uint64_t* a, uint64_t* inv, int n) {
    uint64_t m[12][24];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) { m[i][j] = gl_canon(a[i * n + j]); m[i][n + j] = i == j; }
    for (int col = 0; col < n; col++) {
        int piv = col;
        while (m[piv][col] == 0) piv++;
        if (piv != col)
            for (int j = 0; j < 2 * n; j++) { uint64_t t = m[piv][j]; m[piv][j] = m[col][j]; m[col][j] = t; }
        uint64_t pi = gl_inv(m[col][col]);
        for (int j = 0; j < 2 * n; j++) m[col][j] = gl_canon(gl_mul(m[col][j], pi));
        for (int r = 0; r < n; r++) {
            if (r == col || m[r][col] == 0) continue;
            uint64_t f = m[r][col];
            for (int j = 0; j < 2 * n; j++) m[r][j] = gl_canon(gl_sub(m[r][j], gl_mul(f, m[col][j])));
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) inv[i * n + j] = m[i][n + j];
}

void init_fast_partial() {
    uint64_t M[12][12], Minv[144];
    for (int r = 0; r < 12; r++)
        for (int c = 0; c < 12; c++) M[r][c] = MDS_CIRC[(c - r + 12) % 12] + ((r == 0 && c == 0) ? 8 : 0);
    mat_inverse(&M[0][0], Minv, 12);
    // constants: push every round-constant vector back through M^-1; lane 0 of the pulled-back vector is added after
    // the previous round's S-box, the other lanes merge into the previous round's constants
    uint64_t c[22][12];
    for (int t = 0; t < 22; t++)
        for (int i = 0; i < 12; i++) c[t][i] = SIPP_POSEIDON_RC[12 * (4 + t) + i];
    for (int t = 21; t >= 1; t--) {
        uint64_t d[12];
        for (int r = 0; r < 12; r++) {
            uint64_t acc = 0;
            for (int k = 0; k < 12; k++) acc = gl_add(acc, gl_mul(Minv[r * 12 + k], c[t][k]));
            d[r] = gl_canon(acc);
        }
        g_fp.post[t - 1] = d[0];
        for (int i = 1; i < 12; i++) c[t - 1][i] = gl_canon(gl_add(c[t - 1][i], d[i]));
    }
    g_fp.post[21] = 0;
    for (int i = 0; i < 12; i++) g_fp.first[i] = c[0][i];
    // matrices: Mcur = Msparse_t * diag(1, Mh_t), then Mcur <- diag(1, Mh_t) * M for the round before
    uint64_t cur[12][12];
    for (int r = 0; r < 12; r++)
        for (int k = 0; k < 12; k++) cur[r][k] = M[r][k];
    g_fp.m00 = M[0][0];
    for (int t = 21; t >= 0; t--) {
        uint64_t Mh[121], Mhi[121];
        for (int i = 0; i < 11; i++)
            for (int j = 0; j < 11; j++) Mh[i * 11 + j] = cur[i + 1][j + 1];
        mat_inverse(Mh, Mhi, 11);
        for (int j = 0; j < 11; j++) {  // vhat = v * Mh^-1 (row vector)
            uint64_t acc = 0;
            for (int k = 0; k < 11; k++) acc = gl_add(acc, gl_mul(cur[0][k + 1], Mhi[k * 11 + j]));
            g_fp.vhat[t][j] = gl_canon(acc);
            g_fp.w[t][j] = gl_canon(cur[j + 1][0]);
        }
        uint64_t nxt[12][12];  // diag(1, Mh) * M
        for (int k = 0; k < 12; k++) nxt[0][k] = M[0][k];
        for (int i = 0; i < 11; i++)
            for (int k = 0; k < 12; k++) {
                uint64_t acc = 0;
                for (int j = 0; j < 11; j++) acc = gl_add(acc, gl_mul(Mh[i * 11 + j], M[j + 1][k]));
                nxt[i + 1][k] = gl_canon(acc);
            }
        if (t == 0) {
            for (int i = 0; i < 11; i++)
                for (int j = 0; j < 11; j++) g_fp.init[i][j] = Mh[i * 11 + j];
        }
        for (int r = 0; r < 12; r++)
            for (int k = 0; k < 12; k++) cur[r][k] = nxt[r][k];
    }
}
sipp::PoseidonFastTables g_tab;
sipp::PoseidonIfmaTables g_ifma;
bool g_use_avx512 = false;
bool g_use_ifma = false;

uint64_t gl_dot11(const uint64_t* a, const uint64_t* b) {
    uint64_t acc = 0;
    for (int i = 0; i < 11; i++) acc = gl_add(acc, gl_mul(a[i], b[i]));
    return gl_canon(acc);
}
// tables of the algebraically unrolled partial rounds (layout and derivation: poseidon_fast.h)
void init_ifma_tables() {
    memset(&g_ifma, 0, sizeof g_ifma);
    // lane -> row: C-row j as j (0..21), U-row i as 32 + i
    auto row_of = [](int b, int l) { return b == 0 ? 1 + l : b == 1 ? 9 + l : b == 2 ? (l < 4 ? 32 + 7 + l : 13 + l) : (l == 0 ? 21 : 32 + l - 1); };
    // the chain runs on z_j = u0_j / lam_j with lam_0 = 1, lam_{j+1} = m00 lam_j^7: then z_{j+1} = z_j^7 + R_j / lam_{j+1}, the product by
    // m00 has left the dependent chain; the S-box outputs the vector side sees are z_k^7 = p7_k / lam_k^7
    uint64_t lam[23], lam7[22], lam_inv[23];
    lam[0] = 1;
    for (int j = 0; j < 22; j++) {
        lam7[j] = gl_canon(gl_pow7(lam[j]));
        lam[j + 1] = gl_canon(gl_mul(g_fp.m00, lam7[j]));
    }
    for (int j = 0; j < 23; j++) lam_inv[j] = gl_inv(lam[j]);
    auto row_scale = [&](int row) -> uint64_t { return row < 32 ? lam_inv[row + 1] : 1; };
    auto raw_y = [&](int row, int i) -> uint64_t {  // coefficient of y[1 + i]
        if (row >= 32) return g_fp.init[row - 32][i];
        uint64_t acc = 0;
        for (int m = 0; m < 11; m++) acc = gl_add(acc, gl_mul(g_fp.vhat[row][m], g_fp.init[m][i]));
        return gl_canon(acc);
    };
    auto raw_x = [&](int row, int k) -> uint64_t {  // coefficient of x_k = p7_k + post_k (every k < row for a C-row)
        if (row >= 32) return g_fp.w[k][row - 32];
        return k < row ? gl_dot11(g_fp.vhat[row], g_fp.w[k]) : 0;
    };
    auto coef_y = [&](int row, int i) -> uint64_t { return gl_canon(gl_mul(raw_y(row, i), row_scale(row))); };
    auto coef_x = [&](int row, int k) -> uint64_t { return gl_canon(gl_mul(gl_mul(raw_x(row, k), lam7[k]), row_scale(row))); };
    auto konst = [&](int row) -> uint64_t {  // the post part of every term
        uint64_t acc = row < 32 ? gl_mul(g_fp.m00, g_fp.post[row]) : 0;
        for (int k = 0; k < 22; k++) acc = gl_add(acc, gl_mul(raw_x(row, k), g_fp.post[k]));
        return gl_canon(gl_mul(acc, row_scale(row)));
    };
    const uint64_t M52 = (1ull << 52) - 1;
    for (int b = 0; b < 4; b++)
        for (int l = 0; l < 8; l++) {
            const int row = row_of(b, l);
            // the U-rows become lanes 1..11 of the state the second half of the full rounds starts from: their round constant rides along
            const uint64_t K = row >= 32 ? gl_canon(gl_add(konst(row), SIPP_POSEIDON_RC[12 * 26 + 1 + (row - 32)])) : konst(row);
            g_ifma.acc_init[b][0][l] = (K & M52) + (GL_P & M52);
            g_ifma.acc_init[b][1][l] = (K >> 52) + (GL_P >> 52);
            for (int i = 0; i < 11; i++) {
                const uint64_t c = coef_y(row, i);
                g_ifma.init_c[i][b][0][l] = c;
                g_ifma.init_c[i][b][1][l] = c >> 52;
            }
            for (int k = 0; k < 22; k++) {
                const uint64_t c = (row < 32 && k > row - 2) ? 0 : coef_x(row, k);  // the newest term of a C-row is a scalar multiply-add
                g_ifma.upd_c[k][b][0][l] = c;
                g_ifma.upd_c[k][b][1][l] = c >> 52;
            }
        }
    for (int i = 0; i < 11; i++) g_ifma.row0[i] = coef_y(0, i);
    g_ifma.k0 = konst(0);
    for (int j = 1; j < 22; j++) g_ifma.cdiag[j] = coef_x(j, j - 1);
    g_ifma.lam22 = lam[22];
    for (int k = 0; k < 8; k++)
        for (int l = 0; l < 8; l++) {
            const uint64_t c = k == 7 ? 0 : k == 3 ? g_fp.first[l] : SIPP_POSEIDON_RC[12 * (k < 3 ? k + 1 : 22 + k + 1) + l];
            g_ifma.rc_next[k][0][l] = c & 0xFFFFFFFFull;
            g_ifma.rc_next[k][1][l] = c >> 32;
            const int hl = 8 + (l >> 1);  // lanes 8..11 as (low, high) pairs
            const uint64_t d = k == 7 ? 0 : k == 3 ? g_fp.first[hl] : SIPP_POSEIDON_RC[12 * (k < 3 ? k + 1 : 22 + k + 1) + hl];
            g_ifma.rc_next[k][2][l] = (l & 1) ? d >> 32 : d & 0xFFFFFFFFull;
        }
}
struct FastPartialInit {
    FastPartialInit() {
        init_fast_partial();
        init_ifma_tables();
        memset(&g_tab, 0, sizeof g_tab);
        for (int k = 0; k < 8; k++)
            for (int i = 0; i < 12; i++) g_tab.rc_full[k][i] = SIPP_POSEIDON_RC[12 * (k < 4 ? k : 22 + k) + i];
        for (int i = 0; i < 12; i++) g_tab.first[i] = g_fp.first[i];
        for (int r = 0; r < 22; r++) {
            g_tab.post[r] = g_fp.post[r];
            g_tab.mpost[r] = gl_canon(gl_mul(g_fp.m00, g_fp.post[r]));
            g_tab.kprev[r] = 0;
            for (int i = 0; r > 0 && i < 11; i++) g_tab.kprev[r] = gl_canon(gl_add(g_tab.kprev[r], gl_mul(g_fp.vhat[r][i], g_fp.w[r - 1][i])));
            for (int i = 0; i < 11; i++) { g_tab.w16[r][i + 1] = g_fp.w[r][i]; g_tab.vhat[r][i] = g_fp.vhat[r][i]; }
        }
        for (int i = 0; i < 11; i++)
            for (int j = 0; j < 11; j++) g_tab.init[i][j] = g_fp.init[i][j];
        // out[r] = sum_i s[(i + r) mod 12] CIRC[i] + 8 s[0] [r == 0]  =>  the coefficient of s[j] in row r is CIRC[(j - r) mod 12]
        for (int j = 0; j < 12; j++)
            for (int r = 0; r < 8; r++) {
                g_tab.mds_col_a[j][r] = (double)(MDS_CIRC[(j - r + 12) % 12] + ((j == 0 && r == 0) ? 8 : 0));
                g_tab.mds_col_b[j][r] = (double)MDS_CIRC[(j - 8 - (r & 3) + 24) % 12];
            }
        for (int j = 0; j < 12; j++)
            for (int r = 0; r < 8; r++) {
                g_ifma.mds_icol_a[j][r] = (uint64_t)g_tab.mds_col_a[j][r];
                g_ifma.mds_icol_b[j][r] = (uint64_t)g_tab.mds_col_b[j][r];
                g_ifma.mds_icol_p[j][r] = MDS_CIRC[(j - 8 - (r >> 1) + 24) % 12];
            }
        g_tab.m00 = g_fp.m00;
        for (int r = 0; r < 22; r++) {
            sipp::PartialRound& pr = g_tab.pr[r];
            memset(&pr, 0, sizeof pr);
            for (int i = 0; i < 16; i++) pr.w16[i] = g_tab.w16[r][i];
            for (int i = 0; i < 11; i++) pr.vhat[i] = g_tab.vhat[r][i];
            pr.kprev = g_tab.kprev[r];
            pr.mpost = g_tab.mpost[r];
            pr.post = g_tab.post[r];
        }
        const char* force = getenv("SIPP_POSEIDON");  // "portable" forces the scalar path (tests compare both)
        g_use_avx512 = sipp::poseidon_avx512_supported() && !(force && !strcmp(force, "portable"));
        g_use_ifma = g_use_avx512 && sipp::poseidon_ifma_supported() && !(force && !strcmp(force, "avx512"));  // "avx512": the path without IFMA
    }
} g_fp_init;

// sum of 11 products of canonical-or-not u64 values, reduced once: accumulate the 128-bit products in three limbs
inline uint64_t dot11(const uint64_t* a, const uint64_t* b) {
    u128 lo = 0;
    uint64_t hi = 0;  // counts overflows of the 128-bit accumulator
    for (int i = 0; i < 11; i++) {
        u128 p = (u128)a[i] * b[i];
        lo += p;
        hi += lo < p;
    }
    // value = lo + hi * 2^128, and 2^128 = 2^64 * 2^64 = EPS^2 = 2^64 - 2^33 + 1 ... reduce hi separately:
    // 2^128 mod p = (2^32 - 1)^2 mod p = 2^64 - 2^33 + 1 - p = -2^32 mod p  => hi * 2^128 = -(hi << 32)
    uint64_t r = gl_reduce128(lo);
    uint64_t t = hi << 32;  // hi <= 11
    return gl_sub(r, t);
}

inline void partial_rounds(uint64_t* s) {
    uint64_t u[12];
    for (int i = 0; i < 12; i++) u[i] = gl_canon(gl_add(s[i], g_fp.first[i]));
    {
        uint64_t t[11];
        for (int i = 0; i < 11; i++) t[i] = dot11(g_fp.init[i], u + 1);
        for (int i = 0; i < 11; i++) u[i + 1] = t[i];
    }
    for (int r = 0; r < 22; r++) {
        uint64_t x = gl_add(gl_pow7(u[0]), g_fp.post[r]);
        uint64_t d = gl_add(gl_mul(x, g_fp.m00), dot11(g_fp.vhat[r], u + 1));
        for (int i = 0; i < 11; i++) u[i + 1] = gl_add(u[i + 1], gl_mul(x, g_fp.w[r][i]));
        u[0] = d;
    }
    for (int i = 0; i < 12; i++) s[i] = u[i];
}

}  // namespace

extern "C" {

// portable implementation (also the reference the AVX-512 path is tested against)
void sipp_poseidon_permute_portable(uint64_t s[12]) {
    int rnd = 0;
    for (int k = 0; k < 4; k++, rnd++) {
        for (int i = 0; i < 12; i++) s[i] = gl_pow7(gl_add(s[i], SIPP_POSEIDON_RC[12 * rnd + i]));
        mds_layer(s);
    }
    partial_rounds(s);
    rnd += 22;
    for (int k = 0; k < 4; k++, rnd++) {
        for (int i = 0; i < 12; i++) s[i] = gl_pow7(gl_add(s[i], SIPP_POSEIDON_RC[12 * rnd + i]));
        mds_layer(s);
    }
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}

void sipp_poseidon_permute(uint64_t s[12]) {
    if (g_use_ifma) sipp::poseidon_permute_ifma(s, g_tab, g_ifma);
    else if (g_use_avx512) sipp::poseidon_permute_avx512(s, g_tab);
    else sipp_poseidon_permute_portable(s);
}
int sipp_poseidon_backend(void) { return g_use_ifma ? 2 : g_use_avx512 ? 1 : 0; }
const void* sipp_test_poseidon_ifma_tables(void) { return &g_ifma; }
// nanoseconds per permutation of the selected back end on this CPU: a dependent chain of `count` permutations (what the absorb of
// A, B is), best of three -- bench.py reports it next to the prove time it explains
double sipp_poseidon_ns_per_permutation(long count) {
    uint64_t s[12] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12};
    double best = 1e300;
    for (int rep = 0; rep < 3 && count > 0; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        for (long i = 0; i < count; i++) sipp_poseidon_permute(s);
        const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / (double)count;
        if (ns < best) best = ns;
    }
    return s[0] == 0 ? -best : best;  // (keeps the chain alive)
}
const void* sipp_test_poseidon_tables(void) { return &g_tab; }
// a chain of `count` permutations run by the AVX-512 and the portable code side by side; returns the index of the first
// permutation whose outputs differ, -1 if none (or if the CPU has no AVX-512).  The rare carry paths of the vector code need
// ~10^5 permutations to show up, which is too slow through ctypes one call at a time.
long sipp_test_poseidon_chain(uint64_t seed, long count) {
    if (!sipp::poseidon_avx512_supported()) return -1;
    const bool ifma = sipp::poseidon_ifma_supported();  // every path this CPU can run, whichever one is selected
    uint64_t a[12], b[12], c[12];
    uint64_t z = seed;
    for (int i = 0; i < 12; i++) { z = z * 6364136223846793005ull + 1442695040888963407ull; a[i] = b[i] = c[i] = z; }
    for (long k = 0; k < count; k++) {
        sipp::poseidon_permute_avx512(a, g_tab);
        if (ifma) sipp::poseidon_permute_ifma(c, g_tab, g_ifma);
        sipp_poseidon_permute_portable(b);
        for (int i = 0; i < 12; i++)
            if (a[i] != b[i] || (ifma && c[i] != b[i])) return k;
        if ((k & 1023) == 1023) { a[k % 12] = b[k % 12] = c[k % 12] = ~a[(k + 5) % 12]; }  // also non-canonical lanes now and then
    }
    return -1;
}
// scalar helpers of the AVX-512 file with crafted operands: which = 0 (lo + 2^64 hi) mod p; 1 the closing multiply-add + reduction of a
// partial round ((lo + 2^64 hi + 2^128 top) + p7 m00) mod p; 2 u^7 (out[0]) and u^7 + post (out[1]); 3 (IFMA path) the closing of an
// accumulator lane, in = a0, a1, a2, c, x: out[0] = scalar, out[1] = vector form of (a0 + 2^52 a1 - 2^8 a2 [+ c x]) mod p; 4 (IFMA path) the
// vector product of the full rounds: out[0] = in[0] in[1], out[1] = in[0]^2
int sipp_test_poseidon_scalar(int which, const uint64_t* in, uint64_t* out) {
    if (!sipp::poseidon_avx512_supported()) return -1;
    if (which == 0) out[0] = sipp::poseidon_test_red128(in[0], in[1]);
    else if (which == 1) out[0] = sipp::poseidon_test_finish(in[0], in[1], in[2], in[3], in[4]);
    else if (which == 2) out[0] = sipp::poseidon_test_sbox(in[0], in[1], &out[1]);
    else if (which == 3) {
        if (!sipp::poseidon_ifma_supported()) return -1;
        sipp::poseidon_test_ifma_close(in, out);
    } else {
        if (!sipp::poseidon_ifma_supported()) return -1;
        out[0] = sipp::poseidon_test_vmul_fast(in[0], in[1], 0);
        out[1] = sipp::poseidon_test_vmul_fast(in[0], in[0], 1);
    }
    return 0;
}  // benchmark hook: tools/probe/poseidon_lab.cc times the layers

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
