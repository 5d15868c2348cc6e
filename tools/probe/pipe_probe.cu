// pipe_probe.cu -- do IMAD.WIDE and the integer ALU ops (IADD3 / LOP3 / SHF) co-issue on B200?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint64_t r[8]; uint32_t s[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[i] = a + i; s[i] = b + i * 7; t[i] = a * 5 + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#define WIDE(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"((uint32_t)r[(i + 1) & 7]), "r"(b));
#define WIDE2(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"((uint32_t)r[(i + 1) & 7]), "r"((uint32_t)(r[(i + 5) & 7] >> 32)));
#define ADD(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(a));
#define LOP(i) asm volatile("xor.b32 %0, %0, %1;" : "+r"(s[i]) : "r"(a));
#define SHF(i) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(s[i]) : "r"(a));
#define MADLO(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(a));
#define MADHI(i) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(a));
#define WIDEX(i) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(s[i]), "+r"(t[i]) : "r"(s[(i + 1) & 7]), "r"(b));
#define CARRY8 asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %10, %9, %2;\n\tmadc.hi.cc.u32 %3, %10, %9, %3;\n\tmadc.lo.cc.u32 %4, %11, %9, %4;\n\tmadc.hi.cc.u32 %5, %11, %9, %5;\n\tmadc.lo.cc.u32 %6, %12, %9, %6;\n\tmadc.hi.u32 %7, %12, %9, %7;" : "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]), "+r"(s[4]), "+r"(s[5]), "+r"(s[6]), "+r"(s[7]) : "r"(t[0]), "r"(b), "r"(t[1]), "r"(t[2]), "r"(t[3]));
#define ADD64(i) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(s[i]), "+r"(s[(i + 4) & 7]) : "r"(a), "r"(b));
            if (MODE == 0) { REP8(WIDE) }
            if (MODE == 1) { REP8(ADD) }
            if (MODE == 2) { REP8(LOP) }
            if (MODE == 3) { REP8(SHF) }
            if (MODE == 4) { WIDE(0) ADD(0) WIDE(1) ADD(1) WIDE(2) ADD(2) WIDE(3) ADD(3) WIDE(4) ADD(4) WIDE(5) ADD(5) WIDE(6) ADD(6) WIDE(7) ADD(7) }
            if (MODE == 5) { WIDE(0) LOP(0) WIDE(1) LOP(1) WIDE(2) LOP(2) WIDE(3) LOP(3) WIDE(4) LOP(4) WIDE(5) LOP(5) WIDE(6) LOP(6) WIDE(7) LOP(7) }
            if (MODE == 6) { WIDE(0) SHF(0) WIDE(1) SHF(1) WIDE(2) SHF(2) WIDE(3) SHF(3) WIDE(4) SHF(4) WIDE(5) SHF(5) WIDE(6) SHF(6) WIDE(7) SHF(7) }
            if (MODE == 7) { REP8(MADLO) }
            if (MODE == 8) { ADD64(0) ADD64(1) ADD64(2) ADD64(3) }
            if (MODE == 9) { REP8(WIDE2) }
            if (MODE == 11) { REP8(MADHI) }
            if (MODE == 12) { CARRY8 t[0] ^= s[7]; }
            if (MODE == 13) { WIDE(0) ADD(0) LOP(1) SHF(2) WIDE(1) ADD(3) LOP(4) SHF(5) WIDE(2) ADD(6) LOP(7) SHF(0) WIDE(3) ADD(1) LOP(2) SHF(3) }
            if (MODE == 10) { WIDE(0) WIDE(1) ADD(0) WIDE(2) WIDE(3) ADD(1) WIDE(4) WIDE(5) ADD(2) WIDE(6) WIDE(7) ADD(3) }
        }
    }
    uint64_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= r[i] ^ s[i] ^ t[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int MODE>
void run(const char* name, double ops_per_group, int sm, uint64_t* out) {
    const int it = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sm * 8, 256>>>(out, 100, 1); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); k<MODE><<<sm * 8, 256>>>(out, it, rep); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double instr = ops_per_group * 8 * it;                     // per thread
    double warp_instr_per_smsp = instr * 16;                    // 16 warps per SMSP
    double cycles = best * 1e-3 * 1.965e9;
    printf("%-34s %8.3f ms  %6.2f cycles per warp-instruction per SMSP\n", name, best, cycles / warp_instr_per_smsp);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sm = p.multiProcessorCount;
    uint64_t* out; cudaMalloc(&out, (size_t)sm * 8 * 256 * 8);
    run<0>("IMAD.WIDE.U32 (shared a,b)", 8, sm, out);
    run<9>("IMAD.WIDE.U32 (distinct regs)", 8, sm, out);
    run<7>("IMAD lo", 8, sm, out);
    run<1>("IADD", 8, sm, out);
    run<2>("LOP3", 8, sm, out);
    run<3>("SHF", 8, sm, out);
    run<8>("IADD3 + IADD3.X pairs", 8, sm, out);
    run<4>("WIDE + IADD 1:1", 16, sm, out);
    run<5>("WIDE + LOP 1:1", 16, sm, out);
    run<6>("WIDE + SHF 1:1", 16, sm, out);
    run<10>("WIDE + IADD 2:1", 12, sm, out);
    run<11>("IMAD.HI", 8, sm, out);
    run<12>("carry chain of 4 WIDE.X (per wide)", 4, sm, out);
    run<13>("WIDE + 3 ALU (per instr)", 16, sm, out);
    return 0;
}
