// unsigned radix-2^29 probe variant (uint64 columns, IMAD.WIDE.U32 only)
#pragma once
#include <stdint.h>
namespace u29 {
struct Fq { uint32_t l[9]; };
#define U_P0 0x187cfd47u
#define U_P1 0x010460b6u
#define U_P2 0x1c72a34fu
#define U_P3 0x02d522d0u
#define U_P4 0x1585d978u
#define U_P5 0x02db40c0u
#define U_P6 0x00a6e141u
#define U_P7 0x0e5c2634u
#define U_P8 0x0030644eu
#define U_PINV 0x04866389u
#define U_MASK 0x1fffffffu
__device__ __forceinline__ void red_step(uint64_t* t, int i) {
    const uint32_t m = ((uint32_t)t[i] * U_PINV) & U_MASK;
    t[i + 0] += (uint64_t)m * U_P0; t[i + 1] += (uint64_t)m * U_P1; t[i + 2] += (uint64_t)m * U_P2;
    t[i + 3] += (uint64_t)m * U_P3; t[i + 4] += (uint64_t)m * U_P4; t[i + 5] += (uint64_t)m * U_P5;
    t[i + 6] += (uint64_t)m * U_P6; t[i + 7] += (uint64_t)m * U_P7; t[i + 8] += (uint64_t)m * U_P8;
    t[i + 1] += t[i] >> 29;
}
__device__ __forceinline__ Fq finish(const uint64_t* t) {
    Fq r; uint64_t c = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { uint64_t v = t[9 + k] + c; r.l[k] = (uint32_t)v & U_MASK; c = v >> 29; }
    r.l[8] = (uint32_t)c;
    return r;
}
template <int N>
__device__ __forceinline__ Fq dot(const Fq* a, const Fq* b) {
    uint64_t t[18];
#pragma unroll
    for (int i = 0; i < 18; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
#pragma unroll
        for (int n = 0; n < N; n++) {
#pragma unroll
            for (int j = 0; j < 9; j++) t[i + j] += (uint64_t)a[n].l[j] * b[n].l[i];
        }
        red_step(t, i);
    }
    return finish(t);
}
__device__ __forceinline__ Fq mul(const Fq& a, const Fq& b) { return dot<1>(&a, &b); }
}
