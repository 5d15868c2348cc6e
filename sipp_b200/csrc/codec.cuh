// codec.cuh -- boundary byte formats <-> device Montgomery limbs.
//
// The C ABI (include/sipp_b200.h) speaks ark-serialize's canonical little-endian integers:
//   Fq 32 B | G1 = x||y 64 B | G2 = x.c0||x.c1||y.c0||y.c1 128 B | Fq12 = 12 Fq in nested order 384 B
// Identity points are all-zero (ark's `infinity` carries x = y = 0).  On a little-endian host the 32 bytes of
// a canonical Fq are exactly 8 u32 limbs, so "decoding" is one Montgomery multiplication by R^2.
#pragma once
#include "curve.cuh"

namespace sipp {

SIPP_HD Fq fq_decode(const uint32_t* w) {  // canonical limbs -> Montgomery
    Fq c;
#pragma unroll
    for (int i = 0; i < 8; i++) c.l[i] = w[i];
    return fq_to_mont(c);
}
SIPP_HD void fq_encode(uint32_t* w, const Fq& a) {  // Montgomery -> canonical limbs
    Fq c = fq_from_mont(a);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = c.l[i];
}
SIPP_HD Fq2 fq2_decode(const uint32_t* w) { return Fq2{fq_decode(w), fq_decode(w + 8)}; }
SIPP_HD void fq2_encode(uint32_t* w, const Fq2& a) { fq_encode(w, a.c0); fq_encode(w + 8, a.c1); }
SIPP_HD G1A g1_decode(const uint32_t* w) { return G1A{fq_decode(w), fq_decode(w + 8)}; }
SIPP_HD void g1_encode(uint32_t* w, const G1A& p) { fq_encode(w, p.x); fq_encode(w + 8, p.y); }
SIPP_HD G2A g2_decode(const uint32_t* w) { return G2A{fq2_decode(w), fq2_decode(w + 16)}; }
SIPP_HD void g2_encode(uint32_t* w, const G2A& p) { fq2_encode(w, p.x); fq2_encode(w + 16, p.y); }

// ark nested slot k (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2) holds the w-power coefficient g[SLOT2W[k]]
SIPP_HD int fq12_slot_to_w(int k) { return k < 3 ? 2 * k : 2 * (k - 3) + 1; }
SIPP_HD Fq12 fq12_decode(const uint32_t* w) {
    Fq12 f;
#pragma unroll
    for (int k = 0; k < 6; k++) f.g[fq12_slot_to_w(k)] = fq2_decode(w + 16 * k);
    return f;
}
SIPP_HD void fq12_encode(uint32_t* w, const Fq12& f) {
#pragma unroll
    for (int k = 0; k < 6; k++) fq2_encode(w + 16 * k, f.g[fq12_slot_to_w(k)]);
}

}  // namespace sipp
