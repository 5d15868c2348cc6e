// host_state.h -- process-wide host state of libsipp_b200 shared by its host translation units (sipp_b200.cu: the C ABI of the
// single proof; batch.cu: batched instances).  One host thread per process by contract, so none of this is locked.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/sipp_b200.h"
#include "launch.h"

// device-resident A, B of the current round of one prover (or one rank's strided shard of it)
struct sipp_ctx {
    uint32_t* dA = nullptr;  // n x 16 words, Montgomery
    uint32_t* dB = nullptr;  // n x 32 words
    size_t n = 0, cap = 0;
};

namespace sipp_host {

extern int g_device;             // CUDA device of this process, -1 before sipp_init
extern int g_sm_count;
extern cudaStream_t g_stream;    // the library's non-blocking stream
extern int g_opt_fe_norm, g_opt_fq12_order, g_opt_profile, g_opt_fold_straus, g_opt_batch_kpg_max, g_opt_batch_streams, g_opt_batch_qlines, g_opt_wide_max, g_opt_wide_fold_max, g_opt_fe_engine;
extern sipp_stats g_stats;

int fail(int code, const char* what);             // records the message for sipp_last_error, returns `code`
int cuda_fail(cudaError_t e, const char* what);   // same for a CUDA error, returns SIPP_ERR_CUDA
int ensure_init();
bool is_pow2(size_t n);

// context with uninitialised device arrays for n points (pool blocks)
int ctx_alloc(size_t n, sipp_ctx** out);
// grow-only device memory pool (cudaFree synchronises the device; blocks are recycled, released in sipp_shutdown)
cudaError_t pool_alloc(void** out, size_t bytes);
void pool_free(void* p);
// the line table shared by every Miller launch of the process (grow-only)
int lines_reserve(size_t bytes);
uint32_t* lines_buffer();
// make `later` wait for everything already enqueued on `earlier`
cudaError_t order_after(cudaStream_t later, cudaStream_t earlier);

// CUDA-event span around a kernel class (SIPP_OPT_PROFILE): kind 0 miller, 1 reduce / final exponentiation, 2 fold, 3 other
int span_begin(int kind, cudaStream_t s);
void span_end(int idx, cudaStream_t s);
struct Span {
    int idx;
    cudaStream_t stream;
    Span(int kind, cudaStream_t s) : idx(span_begin(kind, s)), stream(s) {}
    ~Span() { span_end(idx, stream); }
};

}  // namespace sipp_host

#define CK(call)                                                    \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return sipp_host::cuda_fail(e_, #call); \
    } while (0)
