// k_fe.cu -- reduction of the per-block partial products and the final exponentiation on the 32-lane Fq12 machine.
#include "device_common.cuh"

// ------------------------------------------------------------------------------------------------ reduction + final exponentiation
// k_reduce_fe_eng: product of the per-block partials (6-lane cooperative tree, as k_reduce_fe_coop) and then ONE final
// exponentiation per product on the 32-lane Fq12 machine (engine12.cuh).  grid = nprod blocks of 128 threads.
#include "coop.cuh"
#include "machine12.cuh"
namespace sipp {

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
    if (final_exp != 1 && final_exp != 3) {  // 0: the raw product (a partial for another reduction); 2: the product of values that are already
                           // exponentiated (pairing-matrix tail), encoded like a final-exponentiation result
        if (warp == 0 && active_lane && group == 0) {
            if (final_exp == 2) fq2_encode(out + prod * 96 + ((k & 1) * 3 + (k >> 1)) * 16, f);
            else store_fq2_words(out + prod * 96 + k * 16, f);
        }
        return;
    }
    // hand f to the machine: register 0, slot 2k + c.  The exponentiation runs on warp (block index mod 4): with many products per
    // launch (a pairing matrix: up to 1,024 blocks) the machines of the blocks sharing an SM then sit on different schedulers --
    // always on warp 0 they all queued on ONE sub-partition (1,024 exponentiations: 3.45 ms against 1.07 ms for k_mat_fe).
    if (warp == 0 && active_lane && group == 0) {
        lp_store(mslots, f12_reg_base(0) + 2 * k, f.c0);
        lp_store(mslots, f12_reg_base(0) + 2 * k + 1, f.c1);
    }
    __syncthreads();
    if (warp != (int)(blockIdx.x & 3)) return;
    for (int j = lane; j < 37; j += 32) f12_fill_global(mslots, j);
    __syncwarp();
    DevMachine12 mc;
    mc.slots = mslots;
    mc.lane = lane;
    const int res = f12_final_exp(mc, ark_norm != 0);
    if (final_exp == 3) {  // an entry of the pairing matrix (k_mat.cu): stays on the device, register-shaped
        for (int w = lane; w < 96; w += 32) out[prod * 96 + w] = mslots[f12_reg_base(res) * 8 + w];
        return;
    }
    if (lane < 6) {
        const Fq2 g = Fq2{lp_load(mslots, f12_reg_base(res) + 2 * lane), lp_load(mslots, f12_reg_base(res) + 2 * lane + 1)};
        const int slot = (lane & 1) * 3 + (lane >> 1);
        fq2_encode(out + prod * 96 + slot * 16, g);
    }
}

// ------------------------------------------------------------------------------------------------ accumulation on the machine
// k_accum_eng: the Fq12 side of the Miller loop for the latency-bound rounds.  One warp = one machine holds the
// accumulator f of a group of `kpg` pairs (register 0) and folds their lines into it: per tangent step ONE shared
// squaring (MUL12: a DOT6 level on 24 lanes), then per pair one sparse product (SPARSE: a DOT6 level on 12 lanes).
// Lines are staged global -> shared with cp.async one line ahead into two line registers.  4 machines per block, tree
// product through shared memory -> one 384-byte partial per block, same layout as k_accum's.
// grid = (blocks, nprod); group g of product `prod` folds pairs [g kpg, (g + 1) kpg).
#define SIPP_ACC_MACHINES 4
#define SIPP_ACC_REGS 4  // 0 accumulator, 1 / 2 line buffers, 3 exchange
#define SIPP_ACC_SLOTS (SIPP_F12_GLOBAL_SLOTS + SIPP_F12_REG_SLOTS * SIPP_ACC_REGS)
__device__ __forceinline__ void acc_cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__global__ void __launch_bounds__(SIPP_ACC_MACHINES * 32) k_accum_eng(const uint32_t* __restrict__ lines, size_t m_chunk, int kpg, uint32_t* __restrict__ partials,
                                                                     int partial_stride_prod, int block_offset, int tree) {
    __shared__ __align__(16) uint32_t smem[SIPP_ACC_MACHINES * SIPP_ACC_SLOTS * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prod = blockIdx.y;
    uint32_t* slots = smem + warp * (SIPP_ACC_SLOTS * 8);
    const size_t gid = (size_t)blockIdx.x * SIPP_ACC_MACHINES + warp;
    const size_t j0 = gid * (size_t)kpg;
    size_t j1 = j0 + (size_t)kpg;
    if (j1 > m_chunk) j1 = m_chunk;
    const int npairs = j0 < m_chunk ? (int)(j1 - j0) : 0;
    const uint32_t* base = lines + ((size_t)prod * m_chunk + j0) * (size_t)(SIPP_LINES_PER_PAIR * 80);
    DevMachine12 mc;
    mc.slots = slots;
    mc.lane = lane;
    // f = 1, slot 0 of the globals = 0
    if (lane < 12) lp_store(slots, f12_reg_base(0) + lane, lane == 0 ? fq_one() : fq_zero());
    if (lane == 12) lp_store(slots, 0, fq_zero());
    const int total = SIPP_LINES_PER_PAIR * npairs;  // fetch order: step-major (step s, pair q)
    auto prefetch = [&](int idx, int buf) {
        if (idx < total && lane < 20) {
            const int s = idx / npairs, q = idx - s * npairs;
            const uint32_t* src = base + (size_t)q * (SIPP_LINES_PER_PAIR * 80) + s * 80;
            acc_cp_async16(slots + (size_t)f12_reg_base(1 + buf) * 8 + lane * 4, src + lane * 4);
        }
        asm volatile("cp.async.commit_group;");
    };
    int fetch = 0;
    prefetch(0, 0);
    auto fold_lines = [&]() {
        for (int q = 0; q < npairs; q++) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const int buf = fetch & 1;
            prefetch(fetch + 1, buf ^ 1);
            F12_OP3(mc, SPARSE, 0, 0, 1 + buf);
            fetch++;
        }
    };
    const unsigned long long plus = SIPP_ATE_PLUS_MASK, minus = SIPP_ATE_MINUS_MASK;
    if (npairs > 0) {
        for (int i = 63; i >= 0; i--) {
            if (i != 63) F12_OP3(mc, MUL12, 0, 0, 0);
            fold_lines();
            if (((plus | minus) >> i) & 1ull) fold_lines();
        }
        fold_lines();
        fold_lines();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!tree) {  // pairing-matrix tail: every group's Miller value on its own (group index = output index)
        __syncwarp();
        if (npairs > 0) {
            uint32_t* o = partials + ((gid + (size_t)block_offset) * partial_stride_prod + prod) * 96;
            for (int w = lane; w < 96; w += 32) o[w] = slots[f12_reg_base(0) * 8 + w];
        }
        return;
    }
    // tree product over the machines of the block: 2 <- 3 and 0 <- 1 are independent, then 0 <- 2
#pragma unroll 1
    for (int half = SIPP_ACC_MACHINES / 2; half >= 1; half >>= 1) {
        __syncthreads();
        if (warp < half) {
            const uint32_t* other = smem + (warp + half) * (SIPP_ACC_SLOTS * 8) + f12_reg_base(0) * 8;
            for (int w = lane; w < 96; w += 32) slots[f12_reg_base(3) * 8 + w] = other[w];
            __syncwarp();
            F12_OP3(mc, MUL12, 0, 0, 3);
        }
    }
    if (warp == 0) {
        uint32_t* o = partials + ((size_t)(blockIdx.x + block_offset) * partial_stride_prod + prod) * 96;
        for (int w = lane; w < 96; w += 32) o[w] = slots[f12_reg_base(0) * 8 + w];
    }
}
int accum_eng_blocks(size_t m_chunk, int kpg) {
    const size_t groups = (m_chunk + (size_t)kpg - 1) / (size_t)kpg;
    return (int)((groups + SIPP_ACC_MACHINES - 1) / SIPP_ACC_MACHINES);
}
int launch_accum_eng(const uint32_t* lines, size_t m_chunk, int nprod, int kpg, uint32_t* partials, int block_offset, cudaStream_t s) {
    dim3 grid((unsigned)accum_eng_blocks(m_chunk, kpg), (unsigned)nprod);
    k_accum_eng<<<grid, SIPP_ACC_MACHINES * 32, 0, s>>>(lines, m_chunk, kpg, partials, nprod, block_offset, 1);
    return (int)cudaGetLastError();
}
// one Miller value per pair (no product): out[j] = 96 words, register-shaped (slot 2k + c), j < m
int launch_accum_eng_each(const uint32_t* lines, size_t m, uint32_t* out, cudaStream_t s) {
    dim3 grid((unsigned)accum_eng_blocks(m, 1), 1u);
    k_accum_eng<<<grid, SIPP_ACC_MACHINES * 32, 0, s>>>(lines, m, 1, out, 1, 0, 0);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ batched instances
// k_fe_batch_eng: one machine (warp) per product -- multiply the product's `gpp` accumulator-group results, ONE final
// exponentiation, encode into the instance's proof vector (same contract as k_fe_batch, k_coop.cu).  For launches that do not
// fill the GPU the 32-lane machine finishes a final exponentiation in 1.1 ms where the 6-lane groups of k_fe_batch need 1.6 ms.
#define SIPP_FEB_MACHINES 4
#define SIPP_FEB_SLOTS (SIPP_F12_GLOBAL_SLOTS + SIPP_F12_REG_SLOTS * SIPP_F12_FE_REGS)
__global__ void __launch_bounds__(SIPP_FEB_MACHINES * 32) k_fe_batch_eng(const uint32_t* __restrict__ partials, size_t nproducts, int gpp, int nprod,
                                                                        uint32_t* __restrict__ out, size_t out_stride, int slot0, int slot1, int ark_norm) {
    __shared__ __align__(16) uint32_t smem[SIPP_FEB_MACHINES * SIPP_FEB_SLOTS * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* slots = smem + warp * (SIPP_FEB_SLOTS * 8);
    size_t P = (size_t)blockIdx.x * SIPP_FEB_MACHINES + warp;
    const bool have = P < nproducts;
    if (!have) P = nproducts - 1;  // surplus machines shadow the last product, nothing is stored
    for (int j = lane; j < 37; j += 32) f12_fill_global(slots, j);
    const uint32_t* src = partials + P * (size_t)gpp * 96;
    for (int w = lane; w < 96; w += 32) slots[f12_reg_base(0) * 8 + w] = src[w];  // a partial is register-shaped: slot 2k + c
    __syncwarp();
    DevMachine12 mc;
    mc.slots = slots;
    mc.lane = lane;
    for (int g = 1; g < gpp; g++) {
        for (int w = lane; w < 96; w += 32) slots[f12_reg_base(1) * 8 + w] = src[(size_t)g * 96 + w];
        __syncwarp();
        F12_OP3(mc, MUL12, 0, 0, 1);
    }
    const int res = f12_final_exp(mc, ark_norm != 0);
    if (have && lane < 6) {
        const Fq2 g = Fq2{lp_load(slots, f12_reg_base(res) + 2 * lane), lp_load(slots, f12_reg_base(res) + 2 * lane + 1)};
        const int slot = (lane & 1) * 3 + (lane >> 1);
        const size_t inst = P / (size_t)nprod;
        const int y = (int)(P % (size_t)nprod);
        fq2_encode(out + inst * out_stride + (size_t)(y ? slot1 : slot0) * 96 + slot * 16, g);
    }
}
int launch_fe_batch_eng(const uint32_t* partials, size_t nproducts, int gpp, int nprod, uint32_t* out, size_t out_stride, int slot0, int slot1, int ark_norm,
                        cudaStream_t s) {
    k_fe_batch_eng<<<(unsigned)((nproducts + SIPP_FEB_MACHINES - 1) / SIPP_FEB_MACHINES), SIPP_FEB_MACHINES * 32, 0, s>>>(partials, nproducts, gpp, nprod, out,
                                                                                                                        out_stride, slot0, slot1, ark_norm);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ verifier GT update on the machine
// k_gt_fold_eng: out = zl^x * z * zr^xinv (verifier_native.rs:59-61).  Two machines (warps), one power each, GENERIC arithmetic
// (a proof element need not lie in the cyclotomic subgroup and the reference's `pow` does not assume it): fixed 2-bit windows
// over the table {a, a^2, a^3}, every squaring and product a MUL12Y chain link of two levels -- 254 squarings + ~95 products
// per power, against the 254 x (squaring + product) of one thread per power in k_gt_fold (14.5 ms per call).
// in: 3 x 96 words boundary format (zl, z, zr); out: 96 words boundary format.
#define SIPP_GTF_REGS 6  // 0 accumulator, 1..3 table, 4 the other power, 5 z
#define SIPP_GTF_SLOTS (SIPP_F12_GLOBAL_SLOTS + SIPP_F12_REG_SLOTS * SIPP_GTF_REGS)
__global__ void __launch_bounds__(64) k_gt_fold_eng(const uint32_t* __restrict__ in, Scalar256 x, Scalar256 xinv, uint32_t* __restrict__ out) {
    __shared__ __align__(16) uint32_t smem[2 * SIPP_GTF_SLOTS * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* slots = smem + warp * (SIPP_GTF_SLOTS * 8);
    for (int j = lane; j < 37; j += 32) f12_fill_global(slots, j);
    auto load12 = [&](int reg, const uint32_t* src) {  // boundary bytes -> register `reg` (slot 2k + c of the w-basis coefficient k)
        if (lane < 6) {
            const Fq2 g = fq2_decode(src + 16 * ((lane & 1) * 3 + (lane >> 1)));
            lp_store(slots, f12_reg_base(reg) + 2 * lane, g.c0);
            lp_store(slots, f12_reg_base(reg) + 2 * lane + 1, g.c1);
        }
        __syncwarp();
    };
    DevMachine12 mc;
    mc.slots = slots;
    mc.lane = lane;
    load12(1, in + (warp ? 192 : 0));
    const bool started = f12_pow_w2(mc, warp ? xinv.w : x.w);
    if (!started) {  // exponent 0 (the host refuses a zero challenge before it gets here): a^0 = 1
        if (lane < 12) lp_store(slots, f12_reg_base(0) + lane, lane == 0 ? fq_one() : fq_zero());
        __syncwarp();
        F12_OP2(mc, XI6, 0, 0);
    }
    __syncthreads();
    if (warp != 0) return;
    const uint32_t* other = smem + SIPP_GTF_SLOTS * 8 + f12_reg_base(0) * 8;
    for (int w = lane; w < SIPP_F12_REG_SLOTS * 8; w += 32) slots[f12_reg_base(4) * 8 + w] = other[w];
    load12(5, in + 96);
    F12_OP2(mc, XI6, 5, 5);
    F12_OP3(mc, MUL12Y, 0, 0, 5);
    F12_OP3(mc, MUL12Y, 0, 0, 4);
    if (lane < 6) {
        const Fq2 g = Fq2{lp_load(slots, f12_reg_base(0) + 2 * lane), lp_load(slots, f12_reg_base(0) + 2 * lane + 1)};
        fq2_encode(out + ((lane & 1) * 3 + (lane >> 1)) * 16, g);
    }
}
int launch_gt_fold_eng(const uint32_t* in, const Scalar256& x, const Scalar256& xinv, uint32_t* out, cudaStream_t s) {
    k_gt_fold_eng<<<1, 64, 0, s>>>(in, x, xinv, out);
    return (int)cudaGetLastError();
}

// k_gt_fold_rounds: ALL GT updates of one verification at once.  The verifier's challenges depend only on A, B and the proof
// (verifier_native.rs:33-45), so once the transcript has been replayed every Z_L^x, Z_R^(1/x) of every round is an independent power:
// final_Z = Z * prod_rounds Z_L^x_k Z_R^(1/x_k)  (:59-61 unrolled).  Block k < rounds: two machines as in k_gt_fold_eng, their product
// left as a raw partial; block `rounds`: Z itself.  A reduction (k_reduce_fe_eng, product + encode) finishes.
// elems: (2 rounds + 1) x 96 words boundary format in the verifier's read order  Z, Z_L(1), Z_R(1), Z_L(2), ...;
// scalars: rounds x 16 words (x_k, then 1/x_k); partials: (rounds + 1) x 96 words.
__global__ void __launch_bounds__(64) k_gt_fold_rounds(const uint32_t* __restrict__ elems, const uint32_t* __restrict__ scalars, int rounds,
                                                      uint32_t* __restrict__ partials) {
    __shared__ __align__(16) uint32_t smem[2 * SIPP_GTF_SLOTS * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x;
    uint32_t* slots = smem + warp * (SIPP_GTF_SLOTS * 8);
    for (int j = lane; j < 37; j += 32) f12_fill_global(slots, j);
    auto load12 = [&](int reg, const uint32_t* src) {
        if (lane < 6) {
            const Fq2 g = fq2_decode(src + 16 * ((lane & 1) * 3 + (lane >> 1)));
            lp_store(slots, f12_reg_base(reg) + 2 * lane, g.c0);
            lp_store(slots, f12_reg_base(reg) + 2 * lane + 1, g.c1);
        }
        __syncwarp();
    };
    uint32_t* o = partials + (size_t)k * 96;
    if (k == rounds) {  // the factor Z
        if (warp == 0) {
            load12(0, elems);
            for (int w = lane; w < 96; w += 32) o[w] = slots[f12_reg_base(0) * 8 + w];
        }
        return;
    }
    DevMachine12 mc;
    mc.slots = slots;
    mc.lane = lane;
    load12(1, elems + (size_t)(1 + 2 * k + warp) * 96);
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = scalars[(size_t)k * 16 + warp * 8 + i];
    const bool started = f12_pow_w2(mc, e);
    if (!started) {
        if (lane < 12) lp_store(slots, f12_reg_base(0) + lane, lane == 0 ? fq_one() : fq_zero());
        __syncwarp();
        F12_OP2(mc, XI6, 0, 0);
    }
    __syncthreads();
    if (warp != 0) return;
    const uint32_t* other = smem + SIPP_GTF_SLOTS * 8 + f12_reg_base(0) * 8;
    for (int w = lane; w < SIPP_F12_REG_SLOTS * 8; w += 32) slots[f12_reg_base(4) * 8 + w] = other[w];
    __syncwarp();
    F12_OP3(mc, MUL12Y, 0, 0, 4);
    for (int w = lane; w < 96; w += 32) o[w] = slots[f12_reg_base(0) * 8 + w];
}
int launch_gt_fold_rounds(const uint32_t* elems, const uint32_t* scalars, int rounds, uint32_t* partials, cudaStream_t s) {
    k_gt_fold_rounds<<<rounds + 1, 64, 0, s>>>(elems, scalars, rounds, partials);
    return (int)cudaGetLastError();
}

int launch_reduce_fe_eng(const uint32_t* partials, int count, int nprod, uint32_t* out, int final_exp, int ark_norm, cudaStream_t s) {
    k_reduce_fe_eng<<<nprod, SIPP_RFE_THREADS, 0, s>>>(partials, count, nprod, out, final_exp, ark_norm);
    return (int)cudaGetLastError();
}

}  // namespace sipp
