// k_wide.cu -- k_lines_wide: the G2 side of the Miller loop with 16 lanes per pair (engine.cuh), for the latency-bound
// rounds.  Same inputs and the same line table as k_lines (k_coop.cu): lines[prod][pair][91][80 words], so k_accum
// consumes either.  Lines differ from k_lines' by Fq2 factors only (projective scaling), which the final
// exponentiation removes; Z, Z_L, Z_R are bit-identical (tests/test_gpu_parity.py::test_wide_lines_*).
//
// Replaces the per-pair `pairing` calls of /root/reference/src/prover_native.rs:17-22, :48-49 (G2 arithmetic part).
#include "engine.cuh"
#include "device_common.cuh"

namespace sipp {

#define SIPP_WIDE_THREADS 128
#define SIPP_WIDE_GROUPS (SIPP_WIDE_THREADS / SIPP_LP_LANES)
#define SIPP_LINE_WORDS 80

static __device__ const LpIns d_lp_code[SIPP_LP_LEVELS * SIPP_LP_LANES] = SIPP_LP_CODE_INIT;
static __constant__ unsigned char c_lp_types[SIPP_LP_LEVELS] = SIPP_LP_TYPES_INIT;

struct DevMachine {
    uint32_t* slots;
    uint32_t* out;  // line table of this pair
    int lane;
    bool store, ident;
    __device__ __forceinline__ void run(int first, int n) {
#pragma unroll 1
        for (int L = first; L < first + n; L++) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(d_lp_code) + L * SIPP_LP_LANES + lane);
            const LpIns ins{w.x, w.y, w.z, w.w};
            const Fq r = lp_eval(c_lp_types[L], ins, slots);
            __syncwarp();  // every lane of the level has read its operands
            lp_store(slots, lp_dst(ins), r);
            __syncwarp();
        }
    }
    __device__ __forceinline__ void emit(int step) {
        if (lane < 10 && store) {
            Fq v = lp_load(slots, SIPP_LP_SLOT_OUT0 + lane);
            if (ident) v = (lane == 0) ? fq_one() : fq_zero();  // identity input: the pair contributes the factor 1
            uint4* o = reinterpret_cast<uint4*>(out + (size_t)step * SIPP_LINE_WORDS + lane * 8);
            o[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
            o[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
        }
    }
};

__global__ void __launch_bounds__(SIPP_WIDE_THREADS) k_lines_wide(const uint32_t* __restrict__ A, const uint32_t* __restrict__ B, MillerJob job, int nprod,
                                                                 size_t c0, size_t mc, uint32_t* __restrict__ lines) {
    __shared__ __align__(16) uint32_t smem[SIPP_WIDE_GROUPS * SIPP_LP_SLOTS * 8];
    __shared__ int ident_flag[SIPP_WIDE_GROUPS];
    const int group = threadIdx.x / SIPP_LP_LANES, lane = threadIdx.x % SIPP_LP_LANES;
    const size_t total = mc * (size_t)nprod;
    size_t t = (size_t)blockIdx.x * SIPP_WIDE_GROUPS + group;
    const bool valid = t < total;
    if (!valid) t = total - 1;  // surplus groups shadow the last pair (uniform control flow), nothing is stored
    const int prod = (int)(t / mc);
    const size_t j = c0 + (t - (size_t)prod * mc);
    uint32_t* slots = smem + group * (SIPP_LP_SLOTS * 8);
    if (lane == 0) {
        const G1A p = load_g1(A, job.a_off[prod] + j);
        const G2A q = load_g2(B, job.b_off[prod] + j);
        ident_flag[group] = (affine_is_identity(p) || affine_is_identity(q)) ? 1 : 0;
        lp_fill_fixed(slots, p.x, p.y, q.x, q.y);
    }
    __syncwarp();
    DevMachine mach;
    mach.slots = slots;
    mach.out = lines + t * (size_t)(SIPP_LINES_PER_PAIR * SIPP_LINE_WORDS);
    mach.lane = lane;
    mach.store = valid;
    mach.ident = ident_flag[group] != 0;
    lp_miller(mach);
}

int launch_lines_wide(const uint32_t* A, const uint32_t* B, const MillerJob& job, int nprod, size_t c0, size_t mc, uint32_t* lines, cudaStream_t s) {
    const size_t groups = mc * (size_t)nprod;
    k_lines_wide<<<(unsigned)((groups + SIPP_WIDE_GROUPS - 1) / SIPP_WIDE_GROUPS), SIPP_WIDE_THREADS, 0, s>>>(A, B, job, nprod, c0, mc, lines);
    return (int)cudaGetLastError();
}

}  // namespace sipp
