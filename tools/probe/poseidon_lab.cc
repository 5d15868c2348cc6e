// poseidon_lab.cc -- measurements on the GPU box's host CPU behind the Poseidon design (tools/probe/README.md).
// Build: g++ -O3 -march=x86-64-v3 -std=c++17 -o /tmp/poseidon_lab tools/probe/poseidon_lab.cc sipp_b200/csrc/transcript.o \
//        (transcript.o from the library build; the AVX-512 file is included as source)
// Prints: core clock estimate, latency / throughput of the scalar and vector modular products, of the MDS layer, per-permutation time.
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <x86intrin.h>

#include <chrono>

extern "C" void sipp_poseidon_permute(uint64_t s[12]);
extern "C" void sipp_poseidon_permute_portable(uint64_t s[12]);
extern "C" int sipp_poseidon_backend(void);
extern "C" int sipp_get_option(int) { return 0; }  // transcript.o asks for the Fq12 order switch (lives in the CUDA TU)

#define T512 __attribute__((target("avx512f,avx512dq,avx512vl,bmi2,adx")))

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// the implementation file itself: its internal helpers (anonymous namespace) are visible in this translation unit
#ifndef POS_IMPL
#define POS_IMPL "../../sipp_b200/csrc/poseidon_avx512.cc"
#endif
#include POS_IMPL
using namespace sipp;
extern "C" const void* sipp_test_poseidon_tables(void);
extern "C" const void* sipp_test_poseidon_ifma_tables(void);
#define T256 __attribute__((target("avx512f,avx512dq,avx512vl,bmi2,adx")))
T256 static inline __m256i w_reduce(__m256i lo, __m256i hi) {
    const __m256i eps = _mm256_set1_epi64x((long long)EPS);
    __m256i hh = _mm256_srli_epi64(hi, 32);
    __m256i t = _mm256_sub_epi64(lo, hh);
    __mmask8 b = _mm256_cmplt_epu64_mask(lo, hh);
    t = _mm256_mask_sub_epi64(t, b, t, eps);
    __m256i m = _mm256_mul_epu32(hi, eps);
    __m256i r = _mm256_add_epi64(t, m);
    __mmask8 c = _mm256_cmplt_epu64_mask(r, m);
    return _mm256_mask_add_epi64(r, c, r, eps);
}
T256 static inline __m256i w_mul(__m256i x, __m256i y) {
    const __m256i lo32 = _mm256_set1_epi64x((long long)EPS);
    __m256i xh = _mm256_srli_epi64(x, 32), yh = _mm256_srli_epi64(y, 32);
    __m256i ll = _mm256_mul_epu32(x, y), lh = _mm256_mul_epu32(x, yh), hl = _mm256_mul_epu32(xh, y), hh = _mm256_mul_epu32(xh, yh);
    __m256i t0 = _mm256_add_epi64(hl, _mm256_srli_epi64(ll, 32));
    __m256i t1 = _mm256_add_epi64(lh, _mm256_and_si256(t0, lo32));
    __m256i hi = _mm256_add_epi64(hh, _mm256_add_epi64(_mm256_srli_epi64(t0, 32), _mm256_srli_epi64(t1, 32)));
    __m256i lo = _mm256_or_si256(_mm256_and_si256(ll, lo32), _mm256_slli_epi64(t1, 32));
    return w_reduce(lo, hi);
}

#if !defined(LAB_NO_IFMA) && !defined(LAB_NO_IFMA_LAYERS)
template <class Rep> SIPP_IFMA static void lab_ifma_layers(__m512i& s0, const PoseidonFastTables& T, const PoseidonIfmaTables& I, int R, Rep report) {
    uint64_t t[4] = {9, 10, 11, 12};
    double t0 = now();
    for (int i = 0; i < R; i++) s0 = v_pow7_fast(s0);
    report("IFMA path: one zmm x^7 (fast product), chained", now() - t0, R);
    t0 = now();
    for (int i = 0; i < R; i++) s0 = v_mul_fast(s0, s0);
    report("IFMA path: fast zmm modmul, chained", now() - t0, R);
    t0 = now();
    for (int i = 0; i < R; i++) full_round_mixed(s0, t, I.rc_next[i & 7][0], I);
    report("IFMA path: mixed full round, chained", now() - t0, R);
    printf("  [%llu]\n", (unsigned long long)t[0]);
}
#endif
T512 int main() {
    const int N = 20000000;
    // core clock: dependent 64-bit adds retire one per cycle
    double ghz;
    {
        uint64_t a = 1, b = 3;
        double t0 = now();
        for (int i = 0; i < N; i++) {
            asm volatile("add %1, %0\n\tadd %1, %0\n\tadd %1, %0\n\tadd %1, %0\n\tadd %1, %0\n\tadd %1, %0\n\tadd %1, %0\n\tadd %1, %0" : "+r"(a) : "r"(b));
        }
        double dt = now() - t0;
        ghz = 8.0 * N / dt / 1e9;
        printf("core clock estimate: %.2f GHz (dependent add chain)  [%llu]\n", ghz, (unsigned long long)a);
    }
    const bool quick = getenv("LAB_QUICK") != nullptr;  // only the whole permutations
    auto report = [&](const char* name, double dt, double ops) { printf("%-46s %7.2f ns = %6.1f cycles\n", name, dt / ops * 1e9, dt / ops * 1e9 * ghz); };
    if (!quick) {   // scalar modular product: latency
        uint64_t x = 0x123456789abcdef1ull, y = 0xfedcba9876543211ull;
        double t0 = now();
        for (int i = 0; i < N; i++) x = s_mul(x, y);
        report("scalar modmul, dependent chain", now() - t0, N);
        uint64_t a[8] = {1, 2, 3, 4, 5, 6, 7, 8};
        t0 = now();
        for (int i = 0; i < N / 8; i++)
            for (int k = 0; k < 8; k++) a[k] = s_mul(a[k], y);
        report("scalar modmul, 8 independent chains (per op)", now() - t0, N / 8 * 8);
        printf("  [%llu %llu]\n", (unsigned long long)x, (unsigned long long)a[3]);
    }
    if (!quick) {   // vector modular product
        __m512i x = _mm512_set1_epi64(0x123456789abcdef1ll), y = _mm512_set1_epi64(0x7edcba9876543211ll);
        double t0 = now();
        for (int i = 0; i < N / 4; i++) x = v_mul(x, y);
        report("zmm modmul, dependent chain", now() - t0, N / 4);
        __m512i a = x, b = y, c = _mm512_add_epi64(x, y), d = _mm512_sub_epi64(x, y);
        t0 = now();
        for (int i = 0; i < N / 4; i++) { a = v_mul(a, y); b = v_mul(b, y); c = v_mul(c, y); d = v_mul(d, y); }
        report("zmm modmul, 4 independent chains (per op)", now() - t0, N / 4 * 4);
        __m256i p = _mm256_set1_epi64x(0x123456789abcdef1ll), q = _mm256_set1_epi64x(0x7edcba9876543211ll);
        t0 = now();
        for (int i = 0; i < N / 4; i++) p = w_mul(p, q);
        report("ymm modmul, dependent chain", now() - t0, N / 4);
        __m256i e = p, f = q, g = _mm256_add_epi64(p, q), h = _mm256_sub_epi64(p, q), e2 = _mm256_add_epi64(g, q), f2 = _mm256_sub_epi64(h, q);
        t0 = now();
        for (int i = 0; i < N / 4; i++) { e = w_mul(e, q); f = w_mul(f, q); g = w_mul(g, q); h = w_mul(h, q); e2 = w_mul(e2, q); f2 = w_mul(f2, q); }
        report("ymm modmul, 6 independent chains (per op)", now() - t0, N / 4 * 6);
        alignas(64) uint64_t out[8], out2[4];
        _mm512_store_si512(out, _mm512_add_epi64(_mm512_add_epi64(a, b), _mm512_add_epi64(c, _mm512_add_epi64(d, x))));
        _mm256_store_si256((__m256i*)out2, _mm256_add_epi64(_mm256_add_epi64(e, f), _mm256_add_epi64(_mm256_add_epi64(g, h), _mm256_add_epi64(e2, f2))));
        printf("  [%llu %llu %llu]\n", (unsigned long long)out[0], (unsigned long long)out2[1], (unsigned long long)_mm256_extract_epi64(p, 0));
    }
    if (!quick) {   // mixed: one zmm chain + 4 scalar chains side by side (do the scalar ports run under the 512-bit work?)
        __m512i x = _mm512_set1_epi64(0x123456789abcdef1ll), y = _mm512_set1_epi64(0x7edcba9876543211ll);
        uint64_t a[4] = {1, 2, 3, 4}, ys = 0xfedcba9876543211ull;
        double t0 = now();
        for (int i = 0; i < N / 4; i++) {
            x = v_mul(x, y);
            for (int k = 0; k < 4; k++) a[k] = s_mul(a[k], ys);
        }
        report("zmm modmul chain + 4 scalar chains (per step)", now() - t0, N / 4);
        alignas(64) uint64_t out[8];
        _mm512_store_si512(out, x);
        printf("  [%llu %llu]\n", (unsigned long long)out[0], (unsigned long long)a[2]);
    }
    if (!quick) {   // the layers of a full round, each as a dependent chain over (s0, s1)
        const PoseidonFastTables& T = *(const PoseidonFastTables*)sipp_test_poseidon_tables();
        __m512i s0 = _mm512_set_epi64(8, 7, 6, 5, 4, 3, 2, 1), s1 = _mm512_set_epi64(0, 0, 0, 0, 12, 11, 10, 9);
        const int R = 2000000;
        double t0 = now();
        for (int i = 0; i < R; i++) { s0 = v_pow7(s0); s1 = v_pow7(s1); }
        report("S-box layer (2 zmm x^7), chained", now() - t0, R);
        t0 = now();
        for (int i = 0; i < R; i++) s0 = v_pow7(s0);
        report("S-box, one zmm x^7, chained", now() - t0, R);
        t0 = now();
        for (int i = 0; i < R; i++) v_mds(s0, s1, T);
        report("MDS layer, chained", now() - t0, R);
        t0 = now();
        for (int i = 0; i < R; i++) v_full_round(s0, s1, T.rc_full[i & 7], T);
        report("full round, chained", now() - t0, R);
#if !defined(LAB_NO_IFMA) && !defined(LAB_NO_IFMA_LAYERS)
        if (poseidon_ifma_supported()) {
            const PoseidonIfmaTables& I = *(const PoseidonIfmaTables*)sipp_test_poseidon_ifma_tables();
            lab_ifma_layers(s0, T, I, R, report);
        }
#endif
        uint64_t u = 12345;
        t0 = now();
        for (int i = 0; i < R; i++) u = s_pow7(u);
        report("scalar x^7, chained", now() - t0, R);
        uint64_t q[4] = {1, 2, 3, 4};
        t0 = now();
        for (int i = 0; i < R; i++) { s0 = v_pow7(s0); for (int k = 0; k < 4; k++) q[k] = s_pow7(q[k]); }
        report("one zmm x^7 + 4 scalar x^7 side by side", now() - t0, R);
        alignas(64) uint64_t out[8];
        _mm512_store_si512(out, _mm512_add_epi64(s0, s1));
        printf("  [%llu %llu %llu]\n", (unsigned long long)out[0], (unsigned long long)u, (unsigned long long)q[1]);
    }
    {   // whole permutation, chained
        uint64_t s[12];
        for (int i = 0; i < 12; i++) s[i] = i;
        const int P = 400000;
        double best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            double t0 = now();
            for (int i = 0; i < P; i++) poseidon_permute_avx512(s, *(const PoseidonFastTables*)sipp_test_poseidon_tables());
            double dt = now() - t0;
            if (dt < best) best = dt;
        }
        printf("backend %d\n", sipp_poseidon_backend());
#ifndef LAB_NO_IFMA
        if (poseidon_ifma_supported()) {
            double b2 = 1e9;
            for (int rep = 0; rep < 5; rep++) {
                double t0 = now();
                for (int i = 0; i < P; i++) poseidon_permute_ifma(s, *(const PoseidonFastTables*)sipp_test_poseidon_tables(), *(const PoseidonIfmaTables*)sipp_test_poseidon_ifma_tables());
                double dt = now() - t0;
                if (dt < b2) b2 = dt;
            }
            report("Poseidon permutation (IFMA partial rounds), chained", b2, P);
        }
#endif
        report("Poseidon permutation (this variant), chained", best, P);
        { uint64_t a[12], b[12]; for (int i = 0; i < 12; i++) a[i] = b[i] = 0x123456789abcdefull * (i + 3); poseidon_permute_avx512(a, *(const PoseidonFastTables*)sipp_test_poseidon_tables()); sipp_poseidon_permute_portable(b); printf("bit-exact vs portable: %s\n", memcmp(a, b, 96) == 0 ? "yes" : "NO"); }
        best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            double t0 = now();
            for (int i = 0; i < P / 4; i++) sipp_poseidon_permute_portable(s);
            double dt = now() - t0;
            if (dt < best) best = dt;
        }
        report("Poseidon permutation (portable), chained", best, P / 4);
        printf("  [%llu]\n", (unsigned long long)s[0]);
    }
    return 0;
}
