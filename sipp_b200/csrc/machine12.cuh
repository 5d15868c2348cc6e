// machine12.cuh -- the device-side executor of the 32-lane Fq12 machine (engine12.cuh): code table + DevMachine12.  Included by every
// translation unit that runs Fq12 programs (k_fe.cu, k_mat.cu); the table is `static`, one copy per unit.
#pragma once
#include "engine12.cuh"

namespace sipp {

static __device__ const F12Ins d_f12_code[SIPP_F12_LEVELS * SIPP_F12_LANES] = SIPP_F12_CODE_INIT;
static __constant__ unsigned char c_f12_types[SIPP_F12_LEVELS] = SIPP_F12_TYPES_INIT;

struct DevMachine12 {
    uint32_t* slots;
    int lane;
    // a real call: the final exponentiation invokes ~500 programs, inlining the executor into each would explode
    __device__ __noinline__ void run(int first, int n, int d, int a, int b) {
        const int base[4] = {0, f12_reg_base(d), f12_reg_base(a), f12_reg_base(b)};
#pragma unroll 1
        for (int L = first; L < first + n; L++) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(d_f12_code) + L * SIPP_F12_LANES + lane);
            const F12Ins ins{{w.x, w.y, w.z, w.w}};
            Fq r;
            const bool wr = f12_eval(c_f12_types[L], ins, slots, base, r);
            __syncwarp();
            if (wr) lp_store(slots, f12_slot(f12_byte(ins, 0), base), r);
            __syncwarp();
        }
    }
};

}  // namespace sipp
