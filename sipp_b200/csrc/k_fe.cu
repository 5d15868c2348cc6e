// k_fe.cu -- reduction of the per-block partial products and the final exponentiation on the 32-lane Fq12 machine.
#include "device_common.cuh"

// ------------------------------------------------------------------------------------------------ reduction + final exponentiation
// k_reduce_fe_eng: product of the per-block partials (6-lane cooperative tree, as k_reduce_fe_coop) and then ONE final
// exponentiation per product on the 32-lane Fq12 machine (engine12.cuh).  grid = nprod blocks of 128 threads.
#include "coop.cuh"
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

#define SIPP_RFE_THREADS 128
#define SIPP_RFE_GROUPS 20
__device__ __forceinline__ Fq2 rfe_lane_one(int k) { return k == 0 ? fq2_one() : fq2_zero(); }

__global__ void __launch_bounds__(SIPP_RFE_THREADS) k_reduce_fe_eng(const uint32_t* __restrict__ partials, int count, int nprod, uint32_t* __restrict__ out,
                                                                   int final_exp, int ark_norm) {
    __shared__ __align__(16) uint32_t red[SIPP_RFE_GROUPS * 96];
    __shared__ __align__(16) uint32_t mslots[(SIPP_F12_GLOBAL_SLOTS + SIPP_F12_REG_SLOTS * SIPP_F12_FE_REGS) * 8];
    const Lane6 L = lane6_of_thread();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool active_lane = L.k < 6;
    const int group = active_lane ? warp * 5 + L.base / 6 : SIPP_RFE_GROUPS;
    const int k = active_lane ? L.k : 0;
    const int prod = blockIdx.x;
    Fq2 f = rfe_lane_one(k);
    const int rounds = (count + SIPP_RFE_GROUPS - 1) / SIPP_RFE_GROUPS;
    for (int r = 0; r < rounds; r++) {
        const int i = r * SIPP_RFE_GROUPS + group;
        const bool have = active_lane && i < count;
        const Fq2 v = have ? load_fq2_words(partials + ((size_t)i * nprod + prod) * 96 + k * 16) : rfe_lane_one(k);
        f = (r == 0) ? v : coop_mul(L, f, v);
    }
    // tree product over the groups of the block (result in group 0)
    int n = count < SIPP_RFE_GROUPS ? count : SIPP_RFE_GROUPS;
    while (n > 1) {
        const int half = (n + 1) >> 1;
        __syncthreads();
        if (active_lane && group >= half && group < n) store_fq2_words(red + group * 96 + k * 16, f);
        __syncthreads();
        const bool take = active_lane && group + half < n;
        const Fq2 other = take ? load_fq2_words(red + (group + half) * 96 + k * 16) : rfe_lane_one(k);
        f = coop_mul(L, f, other);
        n = half;
    }
    if (warp != 0) return;
    if (!final_exp) {
        if (active_lane && group == 0) store_fq2_words(out + prod * 96 + k * 16, f);
        return;
    }
    // hand f to the machine: register 0, slot 2k + c
    if (active_lane && group == 0) {
        lp_store(mslots, f12_reg_base(0) + 2 * k, f.c0);
        lp_store(mslots, f12_reg_base(0) + 2 * k + 1, f.c1);
    }
    for (int j = lane; j < 37; j += 32) f12_fill_global(mslots, j);
    __syncwarp();
    DevMachine12 mc;
    mc.slots = mslots;
    mc.lane = lane;
    const int res = f12_final_exp(mc, ark_norm != 0);
    if (lane < 6) {
        const Fq2 g = Fq2{lp_load(mslots, f12_reg_base(res) + 2 * lane), lp_load(mslots, f12_reg_base(res) + 2 * lane + 1)};
        const int slot = (lane & 1) * 3 + (lane >> 1);
        fq2_encode(out + prod * 96 + slot * 16, g);
    }
}

int launch_reduce_fe_eng(const uint32_t* partials, int count, int nprod, uint32_t* out, int final_exp, int ark_norm, cudaStream_t s) {
    k_reduce_fe_eng<<<nprod, SIPP_RFE_THREADS, 0, s>>>(partials, count, nprod, out, final_exp, ark_norm);
    return (int)cudaGetLastError();
}

}  // namespace sipp
