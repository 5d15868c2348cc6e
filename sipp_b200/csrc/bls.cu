// bls.cu -- C ABI of the BLS-aggregation input producers (include/sipp_b200.h: sipp_g1_generator_mul_batch, sipp_g2_mul_batch,
// sipp_g2_sum and their device-pointer forms; /root/reference/src/bin/bls_aggregation.rs:95-117), and the seeded synthetic inputs
// built from them (sipp_seeded_inputs[_device]: A_i = [a_i] G1 = keygen, B_i = [b_i] G2 = "signing" the generator).
#include <string.h>

#include "host_state.h"

using namespace sipp;
using namespace sipp_host;

namespace {

// per-device cache: the generators and their window tables (built on first use)
struct GenTables {
    int device = -1;
    uint32_t* gens = nullptr;    // 16 + 32 words, Montgomery
    uint32_t* table1 = nullptr;  // G1 generator
    uint32_t* table2 = nullptr;  // G2 generator
} g_gen;

int gen_tables(cudaStream_t s) {
    if (g_gen.device == g_device) return SIPP_OK;
    if (g_gen.gens) { cudaFree(g_gen.gens); cudaFree(g_gen.table1); cudaFree(g_gen.table2); g_gen = GenTables(); }
    CK(cudaMalloc(&g_gen.gens, 48 * sizeof(uint32_t)));
    CK(cudaMalloc(&g_gen.table1, window_table_bytes(1)));
    CK(cudaMalloc(&g_gen.table2, window_table_bytes(2)));
    int e = launch_generators(g_gen.gens, s);
    if (!e) e = launch_window_table(1, g_gen.gens, g_gen.table1, s);
    if (!e) e = launch_window_table(2, g_gen.gens + 16, g_gen.table2, s);
    if (e) return cuda_fail((cudaError_t)e, "k_window_table");
    g_stats.launches += 3;
    g_gen.device = g_device;
    return SIPP_OK;
}

struct Tmp {  // pool blocks released on scope exit (recycled by later work on the same stream only)
    void* p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    cudaError_t get(void** out, size_t bytes) {
        cudaError_t e = pool_alloc(out, bytes);
        if (e == cudaSuccess) p[n++] = *out;
        return e;
    }
    ~Tmp() { for (int i = 0; i < n; i++) pool_free(p[i]); }
};

bool scalars_below_r(const uint8_t* k, size_t count) {
    static const uint64_t RM[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    for (size_t i = 0; i < count; i++) {
        uint64_t w[4];
        memcpy(w, k + 32 * i, 32);
        bool less = false;
        for (int j = 3; j >= 0; j--) {
            if (w[j] != RM[j]) { less = w[j] < RM[j]; break; }
        }
        if (!less) return false;
    }
    return true;
}

// d_scalars: count x 32 B canonical; d_out: count x (64 | 128) B boundary format; base: NULL = the generator, else one device point (boundary format)
int fixed_base_device(int group, const void* d_base, const void* d_scalars, size_t count, void* d_out, cudaStream_t s) {
    int rc = gen_tables(s);
    if (rc) return rc;
    const int words = group == 1 ? 16 : 32;
    Tmp tmp;
    uint32_t* table = group == 1 ? g_gen.table1 : g_gen.table2;
    if (d_base) {
        uint32_t* mont;
        CK(tmp.get((void**)&mont, words * 4));
        CK(tmp.get((void**)&table, window_table_bytes(group)));
        int e = launch_codec_decode((const uint32_t*)d_base, mont, words / 8, nullptr, s);
        if (!e) e = launch_window_table(group, mont, table, s);
        if (e) return cuda_fail((cudaError_t)e, "k_window_table");
        g_stats.launches += 2;
    }
    uint32_t* mout;
    CK(tmp.get((void**)&mout, count * words * 4));
    Span sp(3, s);
    int e = launch_fixed_base_mul(group, table, (const uint32_t*)d_scalars, count, mout, s);
    if (!e) e = launch_codec_encode(mout, (uint32_t*)d_out, count * (words / 8), s);
    if (e) return cuda_fail((cudaError_t)e, "k_fixed_base_mul");
    g_stats.launches += 2;
    return SIPP_OK;
}

}  // namespace

extern "C" {

int sipp_g1_generator_mul_batch_device(const void* d_scalars, size_t count, void* d_out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!d_scalars || !d_out || count == 0) return fail(SIPP_ERR_ARG, "bad argument");
    rc = fixed_base_device(1, nullptr, d_scalars, count, d_out, g_stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_stream));
    return SIPP_OK;
}

int sipp_g2_mul_batch_device(const void* d_points, size_t point_count, const void* d_scalars, size_t count, void* d_out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!d_points || !d_scalars || !d_out || count == 0) return fail(SIPP_ERR_ARG, "bad argument");
    if (point_count != 1 && point_count != count) return fail(SIPP_ERR_ARG, "point_count must be 1 (one base for every scalar) or count");
    if (point_count == 1 && count >= 64) {  // one base, many scalars: a window table pays for itself
        rc = fixed_base_device(2, d_points, d_scalars, count, d_out, g_stream);
    } else {
        Tmp tmp;
        uint32_t *mp, *mo;
        CK(tmp.get((void**)&mp, count * 128));
        CK(tmp.get((void**)&mo, count * 128));
        int e = 0;
        if (point_count == 1) {
            for (size_t i = 0; i < count && !e; i++) e = launch_codec_decode((const uint32_t*)d_points, mp + 32 * i, 4, nullptr, g_stream);
        } else {
            e = launch_codec_decode((const uint32_t*)d_points, mp, count * 4, nullptr, g_stream);
        }
        Span sp(3, g_stream);
        if (!e) e = launch_g2_mul_var(mp, (const uint32_t*)d_scalars, count, mo, g_stream);
        if (!e) e = launch_codec_encode(mo, (uint32_t*)d_out, count * 4, g_stream);
        if (e) return cuda_fail((cudaError_t)e, "k_g2_mul_var");
        g_stats.launches += 3;
        rc = SIPP_OK;
    }
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_stream));
    return SIPP_OK;
}

int sipp_g2_sum_device(const void* d_points, size_t count, void* d_out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!d_out || (count && !d_points)) return fail(SIPP_ERR_ARG, "bad argument");
    if (count == 0) {  // fold over an empty iterator: G2Projective::zero().into() = the identity (all-zero bytes)
        CK(cudaMemsetAsync(d_out, 0, 128, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        return SIPP_OK;
    }
    Tmp tmp;
    const int blocks = g2_sum_blocks(count, g_sm_count);
    uint32_t *mp, *parts, *mo;
    CK(tmp.get((void**)&mp, count * 128));
    CK(tmp.get((void**)&parts, (size_t)blocks * 48 * 4));
    CK(tmp.get((void**)&mo, 128));
    int e = launch_codec_decode((const uint32_t*)d_points, mp, count * 4, nullptr, g_stream);
    {
        Span sp(3, g_stream);
        if (!e) e = launch_g2_sum(mp, count, parts, blocks, mo, g_stream);
    }
    if (!e) e = launch_codec_encode(mo, (uint32_t*)d_out, 4, g_stream);
    if (e) return cuda_fail((cudaError_t)e, "k_g2_sum");
    g_stats.launches += 4;
    CK(cudaStreamSynchronize(g_stream));
    return SIPP_OK;
}

// ---- host-buffer forms ----
int sipp_g1_generator_mul_batch(const uint8_t* scalars, size_t count, uint8_t* out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!scalars || !out || count == 0) return fail(SIPP_ERR_ARG, "bad argument");
    if (!scalars_below_r(scalars, count)) return fail(SIPP_ERR_ENCODING, "scalar >= r");
    Tmp tmp;
    uint8_t *dk, *dout;
    CK(tmp.get((void**)&dk, count * 32));
    CK(tmp.get((void**)&dout, count * 64));
    CK(cudaMemcpyAsync(dk, scalars, count * 32, cudaMemcpyHostToDevice, g_stream));
    rc = sipp_g1_generator_mul_batch_device(dk, count, dout);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, count * 64, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return SIPP_OK;
}

int sipp_g2_mul_batch(const uint8_t* points, size_t point_count, const uint8_t* scalars, size_t count, uint8_t* out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!points || !scalars || !out || count == 0) return fail(SIPP_ERR_ARG, "bad argument");
    if (point_count != 1 && point_count != count) return fail(SIPP_ERR_ARG, "point_count must be 1 (one base for every scalar) or count");
    if (!scalars_below_r(scalars, count)) return fail(SIPP_ERR_ENCODING, "scalar >= r");
    if (!fq_bytes_canonical(points, 4 * point_count)) return fail(SIPP_ERR_ENCODING, "input coordinate >= p");
    Tmp tmp;
    uint8_t *dp, *dk, *dout;
    CK(tmp.get((void**)&dp, point_count * 128));
    CK(tmp.get((void**)&dk, count * 32));
    CK(tmp.get((void**)&dout, count * 128));
    CK(cudaMemcpyAsync(dp, points, point_count * 128, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(dk, scalars, count * 32, cudaMemcpyHostToDevice, g_stream));
    rc = sipp_g2_mul_batch_device(dp, point_count, dk, count, dout);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, count * 128, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return SIPP_OK;
}

int sipp_g2_sum(const uint8_t* points, size_t count, uint8_t out[128]) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!out || (count && !points)) return fail(SIPP_ERR_ARG, "bad argument");
    if (count && !fq_bytes_canonical(points, 4 * count)) return fail(SIPP_ERR_ENCODING, "input coordinate >= p");
    Tmp tmp;
    uint8_t *dp, *dout;
    CK(tmp.get((void**)&dp, count ? count * 128 : 128));
    CK(tmp.get((void**)&dout, 128));
    if (count) CK(cudaMemcpyAsync(dp, points, count * 128, cudaMemcpyHostToDevice, g_stream));
    rc = sipp_g2_sum_device(dp, count, dout);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, 128, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return SIPP_OK;
}

// -G1Affine::generator()   bls_aggregation.rs:116   (host: (1, p - 2))
int sipp_g1_neg_generator(uint8_t out[64]) {
    if (!out) return fail(SIPP_ERR_ARG, "null argument");
    static const uint64_t PM2[4] = {0x3c208c16d87cfd45ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    memset(out, 0, 64);
    out[0] = 1;
    memcpy(out + 32, PM2, 32);
    return SIPP_OK;
}

// ------------------------------------------------------------------------------------------------ seeded synthetic inputs
// A_i = [a_i] G1 (keygen with the seeded scalars), B_i = [b_i] G2 (the same producer over the G2 generator's table)
int sipp_seeded_inputs_device(uint64_t seed, size_t n, void* dA, void* dB) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!dA || !dB || n == 0) return fail(SIPP_ERR_ARG, "bad argument");
    Tmp tmp;
    uint32_t *sa, *sb;
    CK(tmp.get((void**)&sa, n * 32));
    CK(tmp.get((void**)&sb, n * 32));
    int e = launch_seeded_scalars(seed, n, sa, sb, g_stream);
    if (e) return cuda_fail((cudaError_t)e, "k_seeded_scalars");
    g_stats.launches++;
    rc = fixed_base_device(1, nullptr, sa, n, dA, g_stream);
    if (!rc) rc = fixed_base_device(2, nullptr, sb, n, dB, g_stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_stream));
    return SIPP_OK;
}

int sipp_seeded_inputs(uint64_t seed, size_t n, uint8_t* A, uint8_t* B) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!A || !B || n == 0) return fail(SIPP_ERR_ARG, "bad argument");
    Tmp tmp;
    uint8_t *dA, *dB;
    CK(tmp.get((void**)&dA, n * 64));
    CK(tmp.get((void**)&dB, n * 128));
    rc = sipp_seeded_inputs_device(seed, n, dA, dB);
    if (rc) return rc;
    CK(cudaMemcpyAsync(A, dA, n * 64, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaMemcpyAsync(B, dB, n * 128, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return SIPP_OK;
}

}  // extern "C"
