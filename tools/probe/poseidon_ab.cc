// poseidon_ab.cc -- interleaved timing of the host Poseidon back ends (min over many short runs; the ratio is what counts on a noisy machine).
// Build: g++ -O2 -std=c++17 -o /tmp/poseidon_ab tools/probe/poseidon_ab.cc sipp_b200/csrc/transcript.o sipp_b200/csrc/poseidon_avx512.o
#include <stdint.h>
#include <stdio.h>

#include <chrono>

#include "../../sipp_b200/csrc/poseidon_fast.h"
extern "C" const void* sipp_test_poseidon_tables(void);
extern "C" const void* sipp_test_poseidon_ifma_tables(void);
extern "C" void sipp_poseidon_permute_portable(uint64_t s[12]);
extern "C" int sipp_get_option(int) { return 0; }
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const auto& T = *(const sipp::PoseidonFastTables*)sipp_test_poseidon_tables();
    const auto& I = *(const sipp::PoseidonIfmaTables*)sipp_test_poseidon_ifma_tables();
    uint64_t s[12];
    for (int i = 0; i < 12; i++) s[i] = i;
    const int P = 20000, REPS = 200;
    double best[2] = {1e9, 1e9};
    const bool ifma = sipp::poseidon_ifma_supported();
    for (int rep = 0; rep < REPS; rep++)
        for (int v = 0; v < (ifma ? 2 : 1); v++) {
            double t0 = now();
            if (v == 0) for (int i = 0; i < P; i++) sipp::poseidon_permute_avx512(s, T);
            else for (int i = 0; i < P; i++) sipp::poseidon_permute_ifma(s, T, I);
            double dt = now() - t0;
            if (dt < best[v]) best[v] = dt;
        }
    printf("avx512 %.1f ns   ifma %.1f ns   ratio %.3f   [%llu]\n", best[0] / P * 1e9, best[1] / P * 1e9, best[1] / best[0], (unsigned long long)s[0]);
    return 0;
}
