// sharded.cu -- sipp_prove_native over `world` GPUs with the collective INSIDE the library (C ABI: sipp_comm_*,
// sipp_prove_native_sharded[_device], sipp_prove_native_sharded_backend).
//
// One process per GPU.  Rank g owns the pairs i = g (mod world) ("strided ownership"): index i and its fold partner
// i + n/2 are congruent mod world while n >= 2 world, so each rank folds its local array of n / world elements exactly
// like a single-GPU instance and no point ever moves -- until a rank is down to one pair, when the `world` remaining
// pairs are gathered to rank 0, which finishes the tail rounds alone.  What crosses the fabric per round is each rank's
// un-exponentiated partial Miller products (2 x 384 B, written by the reduce kernel straight into this rank's slot of
// the gather buffer, then ONE in-place ncclAllGather on the library stream) and the 64-byte challenge (x, x^-1) broadcast
// by rank 0, which owns the Fiat-Shamir transcript (a strictly serial hash chain: transcript_native.rs:25-30).
//
// The protocol loop (the order of transcript appends, when the tail collapses, the reversal of the proof) is
// `sharded_protocol` below and follows /root/reference/src/prover_native.rs:26-80 line by line; it talks to the
// compute / exchange side through the six callbacks of sipp_shard_backend.  The product backend is CudaBackend
// (this file: the kernels of this library + NCCL, or + host-memory collectives supplied by the caller's own fabric);
// the CPU tests plug a backend of their own into the same loop (tests/test_sharded_gloo.py).
#include <dlfcn.h>
#include <string.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "host_state.h"

using namespace sipp;
using namespace sipp_host;

// ------------------------------------------------------------------------------------------------ NCCL, bound at run time
// libnccl.so.2 is resolved with dlopen when a communicator is first asked for: a single-GPU user of this library never
// needs NCCL, and inside a process that already loaded a libnccl.so.2 (e.g. through torch) the same copy is used.
namespace {

typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm;
enum { NCCL_UINT8 = 1 };  // ncclUint8 (ncclInt8 = ncclChar = 0), stable across NCCL 2.x

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm*, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*GetVersion)(int*) = nullptr;
} g_nccl;

int nccl_load() {
    if (g_nccl.handle) return SIPP_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(SIPP_ERR_COMM, "libnccl.so.2 not found (dlopen)");
#define BIND(field, sym)                                                   \
    *(void**)(&g_nccl.field) = dlsym(h, sym);                              \
    if (!g_nccl.field) { dlclose(h); return fail(SIPP_ERR_COMM, "libnccl lacks " sym); }
    BIND(GetUniqueId, "ncclGetUniqueId")
    BIND(CommInitRank, "ncclCommInitRank")
    BIND(CommDestroy, "ncclCommDestroy")
    BIND(GetErrorString, "ncclGetErrorString")
    BIND(AllGather, "ncclAllGather")
    BIND(Broadcast, "ncclBroadcast")
    BIND(GetVersion, "ncclGetVersion")
#undef BIND
    g_nccl.handle = h;
    return SIPP_OK;
}

int nccl_fail(int rc, const char* what) {
    std::string m = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
    return fail(SIPP_ERR_COMM, m.c_str());
}
#define NK(call)                                      \
    do {                                              \
        int r_ = (call);                              \
        if (r_ != 0) return nccl_fail(r_, #call);     \
    } while (0)

// ------------------------------------------------------------------------------------------------ communicator of this process
enum { COMM_NONE = 0, COMM_NCCL = 1, COMM_HOST = 2 };
const size_t SLOT = 2 * SIPP_PARTIAL_BYTES;  // bytes per rank in the gather buffer per round: two partials
const size_t STAGE_ENTRIES_MAX = 1024;        // entries of a distributed first stage (32 x 32 blocks): each rank ships 384 B per entry
const size_t COLLAPSE_MAX = 4096;            // the collapse gathers at most this many points (192 B each) through the same buffer
const size_t XS = 72;                        // x || x^-1 || status word (+ padding)

struct Comm {
    int kind = COMM_NONE, rank = 0, world = 1;
    nccl_comm nccl = nullptr;
    sipp_allgather_fn h_allgather = nullptr;
    sipp_broadcast_fn h_broadcast = nullptr;
    void* user = nullptr;
    uint8_t* d_gather = nullptr;  // [world][SLOT], or the points of the collapse
    uint8_t* d_xs = nullptr;      // XS bytes
    uint8_t* h_stage = nullptr;   // pinned: gather bytes + XS
    size_t gather_bytes = 0;
} g_comm;

int comm_buffers() {
    g_comm.gather_bytes = (size_t)g_comm.world * SLOT;
    if (g_comm.gather_bytes < COLLAPSE_MAX * 192) g_comm.gather_bytes = COLLAPSE_MAX * 192;
    if (g_comm.gather_bytes < (size_t)g_comm.world * STAGE_ENTRIES_MAX * SIPP_PARTIAL_BYTES) g_comm.gather_bytes = (size_t)g_comm.world * STAGE_ENTRIES_MAX * SIPP_PARTIAL_BYTES;
    CK(cudaMalloc(&g_comm.d_gather, g_comm.gather_bytes));
    CK(cudaMalloc(&g_comm.d_xs, XS));
    CK(cudaMallocHost(&g_comm.h_stage, g_comm.gather_bytes + XS));
    return SIPP_OK;
}

// all-gather `bytes` per rank, in place in d_gather ([rank][bytes], this rank's part already written), on the library stream
int comm_allgather(size_t bytes) {
    Comm& c = g_comm;
    if (c.world == 1) return SIPP_OK;
    if (c.kind == COMM_NCCL) {
        NK(g_nccl.AllGather(c.d_gather + (size_t)c.rank * bytes, c.d_gather, bytes, NCCL_UINT8, c.nccl, g_stream));
        g_stats.launches++;  // NCCL's kernel, counted as a launch on our stream
        return SIPP_OK;
    }
    // host collectives: stage this rank's part through pinned memory
    uint8_t* mine = c.h_stage + (size_t)c.rank * bytes;
    CK(cudaMemcpyAsync(mine, c.d_gather + (size_t)c.rank * bytes, bytes, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    std::vector<uint8_t> send(mine, mine + bytes);
    if (c.h_allgather(c.user, send.data(), c.h_stage, bytes)) return fail(SIPP_ERR_COMM, "host all-gather callback failed");
    CK(cudaMemcpyAsync(c.d_gather, c.h_stage, (size_t)c.world * bytes, cudaMemcpyHostToDevice, g_stream));
    return SIPP_OK;
}

// broadcast XS bytes from rank 0's host buffer to every rank's host buffer
int comm_broadcast_xs(uint8_t* xs) {
    Comm& c = g_comm;
    if (c.world == 1) return SIPP_OK;
    if (c.kind == COMM_NCCL) {
        uint8_t* h = c.h_stage + c.gather_bytes;
        if (c.rank == 0) {
            memcpy(h, xs, XS);
            CK(cudaMemcpyAsync(c.d_xs, h, XS, cudaMemcpyHostToDevice, g_stream));
        }
        NK(g_nccl.Broadcast(c.d_xs, c.d_xs, XS, NCCL_UINT8, 0, c.nccl, g_stream));
        g_stats.launches++;
        if (c.rank != 0) {
            CK(cudaMemcpyAsync(h, c.d_xs, XS, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            memcpy(xs, h, XS);
        }
        return SIPP_OK;
    }
    if (c.h_broadcast(c.user, xs, XS, 0)) return fail(SIPP_ERR_COMM, "host broadcast callback failed");
    return SIPP_OK;
}

// ------------------------------------------------------------------------------------------------ the product backend
struct CudaBackend {
    sipp_ctx* ctx = nullptr;
    bool collapsed = false;  // the tail lives on rank 0 alone
    size_t collapse_at = 0;  // points left (in all) at which the tail moves to rank 0
    int last_nprod = 0;
    uint8_t tail_out[768];   // products that rank 0 computes alone: the collapsed tail, and the rounds of a distributed first stage
    bool tail_round = false;
    int stage_left = 0;      // rounds the distributed first stage still serves (the same on every rank)
};
// The FIRST stage over the ranks (sipp_b200.cu mat_stage_first, k_mat.cu): with strided ownership a block of m consecutive global
// indices holds m / world points of every rank, so each rank computes ITS factor of every block product <A_block_i, B_block_j>
// (un-exponentiated), the factors are all-gathered (384 B per entry and rank: 3 MB at 32 x 32 blocks on 8 GPUs -- the one real
// exchange of the protocol) and rank 0 multiplies and exponentiates them.  The first log2(nr) rounds then need no collective but
// the challenge broadcast: rank 0 folds the matrix, every rank folds its own points.  Blocks: as many as keep a rank inside the
// Miller-loop budget of mat_first_budget, at most 32, and the stage must end before the tail moves to rank 0.
size_t cb_first_stage_blocks(size_t local_n, size_t collapse_at) {
    const size_t world = (size_t)g_comm.world, n = local_n * world;
    if (world == 1 || !g_opt_matrix_first || !g_opt_pipeline || !g_opt_fe_engine || g_opt_matrix_n < 2) return 0;
    size_t end_n = collapse_at > 2 * world ? collapse_at : 2 * world;
    size_t nr = 32;
    while (nr >= 4 && (local_n * nr > mat_first_budget(n) || nr > local_n || n / nr < end_n)) nr >>= 1;
    return nr >= 4 ? nr : 0;
}
bool cb_alone(const CudaBackend* b) { return b->collapsed || g_comm.world == 1; }

size_t cb_local_len(void* u) { return ((CudaBackend*)u)->ctx ? ((CudaBackend*)u)->ctx->n : 0; }

int cb_products(void* u, int which) {
    CudaBackend* b = (CudaBackend*)u;
    const int nprod = which == 0 ? 1 : 2;
    b->last_nprod = nprod;
    if (which == 0 && !b->collapsed) {
        const size_t nr = cb_first_stage_blocks(b->ctx->n, b->collapse_at);
        if (nr) {
            // every rank: its factors of the nr^2 block products, straight into its slot of the gather buffer
            const size_t bytes = nr * nr * SIPP_PARTIAL_BYTES;
            MatTail none;
            int rc = mat_build_ex(b->ctx, none, nr, (uint32_t*)(g_comm.d_gather + (size_t)g_comm.rank * bytes));
            if (!rc) rc = comm_allgather(bytes);
            if (rc) return rc;
            int log2nr = 0;
            while (((size_t)1 << log2nr) < nr) log2nr++;
            b->stage_left = log2nr;
            b->tail_round = true;
            if (g_comm.rank != 0) return SIPP_OK;
            rc = mat_adopt(b->ctx->mt, (const uint32_t*)g_comm.d_gather, g_comm.world, nr);
            return rc ? rc : mat_diag_product(b->ctx->mt, b->tail_out);      // Z = prod E[i][i]
        }
    }
    if (which == 1 && b->stage_left > 0 && !b->collapsed) {                  // a round of the distributed stage: rank 0's matrix
        b->tail_round = true;
        return g_comm.rank == 0 ? mat_products(b->ctx->mt, b->tail_out, b->tail_out + 384) : SIPP_OK;
    }
    b->tail_round = which == 1 && cb_alone(b);
    if (b->tail_round) {
        b->ctx->stages = true;
        return sipp_ctx_cross_products(b->ctx, b->tail_out, b->tail_out + 384);
    }
    const size_t bytes = (size_t)nprod * SIPP_PARTIAL_BYTES;
    const int slot = b->collapsed ? 0 : g_comm.rank;
    // the reduce kernel's epilogue writes this rank's partial(s) straight into its slot of the gather buffer
    int rc = sipp_ctx_partial_products(b->ctx, which, g_comm.d_gather + (size_t)slot * bytes, g_stream);
    if (rc || b->collapsed) return rc;
    return comm_allgather(bytes);
}

int cb_combine(void* u, int nprod, uint8_t* out) {
    CudaBackend* b = (CudaBackend*)u;
    if (b->tail_round) {
        memcpy(out, b->tail_out, (size_t)nprod * 384);
        return SIPP_OK;
    }
    return sipp_combine_partials(g_comm.d_gather, b->collapsed ? 1 : g_comm.world, nprod, out, g_stream);
}

int cb_broadcast(void*, uint8_t* xs) { return comm_broadcast_xs(xs); }

int cb_fold(void* u, const uint8_t* x, const uint8_t* xinv) {
    CudaBackend* b = (CudaBackend*)u;
    if (b->stage_left > 0) b->stage_left--;
    return sipp_ctx_fold(b->ctx, x, xinv);  // rank 0 during a stage: the matrix AND (side stream) its points; otherwise the points
}

// The tail moves to rank 0: every rank holds L = (points left) / world of them (strided: local element l is global l world + rank).
// Each rank ships [A_local (64 L) | B_local (128 L)], rank 0 interleaves them back into global order.
int cb_collapse(void* u) {
    CudaBackend* b = (CudaBackend*)u;
    Comm& c = g_comm;
    const size_t L = b->ctx->n, rec = 192 * L;
    if ((size_t)c.world * rec > c.gather_bytes) return fail(SIPP_ERR_ARG, "collapse: more points than the gather buffer holds");
    uint8_t* mine = c.d_gather + (size_t)c.rank * rec;
    CK(cudaMemcpyAsync(mine, b->ctx->dA, 64 * L, cudaMemcpyDeviceToDevice, g_stream));
    CK(cudaMemcpyAsync(mine + 64 * L, b->ctx->dB, 128 * L, cudaMemcpyDeviceToDevice, g_stream));
    int rc = comm_allgather(rec);
    if (rc) return rc;
    b->collapsed = true;
    if (c.rank != 0) return SIPP_OK;
    sipp_ctx* tail;
    rc = ctx_alloc((size_t)c.world * L, &tail);
    if (rc) return rc;
    cudaError_t e = cudaSuccess;
    for (int r = 0; r < c.world && e == cudaSuccess; r++) {
        const uint8_t* src = c.d_gather + (size_t)r * rec;
        e = cudaMemcpy2DAsync((uint8_t*)tail->dA + 64 * (size_t)r, 64 * (size_t)c.world, src, 64, 64, L, cudaMemcpyDeviceToDevice, g_stream);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync((uint8_t*)tail->dB + 128 * (size_t)r, 128 * (size_t)c.world, src + 64 * L, 128, 128, L, cudaMemcpyDeviceToDevice, g_stream);
    }
    if (e != cudaSuccess) { sipp_ctx_destroy(tail); return cuda_fail(e, "tail gather"); }
    sipp_ctx_destroy(b->ctx);  // pool block: recycled only by later work on the same stream
    b->ctx = tail;
    return SIPP_OK;
}

// first non-zero status over the ranks (a rank that failed to build its shard must not leave the others in a collective)
int cb_agree(void*, int rc) {
    Comm& c = g_comm;
    if (c.world == 1) return rc;
    int32_t v = rc;
    cudaError_t e = cudaMemcpyAsync(c.d_gather + (size_t)c.rank * 4, &v, 4, cudaMemcpyHostToDevice, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);  // `v` is pageable stack memory
    if (e != cudaSuccess) return cuda_fail(e, "status exchange");
    int r2 = comm_allgather(4);
    if (r2) return r2;
    std::vector<int32_t> all(c.world);
    e = cudaMemcpyAsync(all.data(), c.d_gather, 4 * (size_t)c.world, cudaMemcpyDeviceToHost, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
    if (e != cudaSuccess) return cuda_fail(e, "status exchange");
    for (int r = 0; r < c.world; r++)
        if (all[r]) return rc ? rc : fail(all[r], "another rank failed to build its shard");
    return SIPP_OK;
}

// ------------------------------------------------------------------------------------------------ the protocol loop
// `started`: rank 0's absorb job when the caller already started it (before the upload of its shard), else NULL
// `collapse_at`: the tail moves to rank 0 as soon as at most this many points are left in all (never later than one pair per rank)
int sharded_protocol(const sipp_shard_backend* be, size_t n, const uint8_t* A_full, const uint8_t* B_full, uint8_t* proof, AbsorbJob* started,
                     size_t collapse_at) {
    const int rank = be->rank, world = be->world;
    const size_t np = sipp_proof_len(n);
    std::vector<uint8_t> fwd(rank == 0 ? np * 384 : 0);  // proof in push order; reversed at the end (prover_native.rs:78)
    size_t k = 0;
    AbsorbJob own;
    AbsorbJob& job = started ? *started : own;
    // register A and B (prover_native.rs:36-39): 8n strictly serial permutations that need nothing from the GPUs, so the
    // chain runs on a host thread while every rank computes Z and the first Z_L, Z_R
    if (rank == 0 && !started) own.start(A_full, B_full, n);
    sipp_transcript& tr = job.tr;

    int rc = be->products(be->user, 0);                                        // let Z = inner_product(A, B);   :29
    if (!rc && rank == 0) rc = be->combine(be->user, 1, &fwd[0]);
    k = 1;
    size_t cur = n;
    bool first = true;
    auto absorb_z = [&]() {
        g_stats.transcript_ms += job.join();
        sipp_transcript_append_fq12(&tr, &fwd[0]);                             // proof.push(Z); transcript.append_fq12(Z)  :42-43
        first = false;
    };
    // one round on whichever ranks take part: products, (rank 0) transcript and challenge, fold
    auto round = [&](bool collective) -> int {
        uint8_t xs[XS];
        memset(xs, 0, sizeof xs);
        int r = be->products(be->user, 1);                                     // Z_L, Z_R   :46-49
        if (rank == 0) {
            uint8_t* zl = &fwd[384 * k];
            uint8_t* zr = zl + 384;
            if (!r) r = be->combine(be->user, 2, zl);
            if (!r) {
                if (first) absorb_z();
                auto h0 = std::chrono::steady_clock::now();
                sipp_transcript_append_fq12(&tr, zl);                          // :52-53
                sipp_transcript_append_fq12(&tr, zr);                          // :54-55
                k += 2;
                sipp_transcript_get_challenge(&tr, xs);                        // :57
                r = sipp_fr_inverse(xs, xs + 32);                              // :58
                g_stats.transcript_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
            }
            int32_t st = r;
            memcpy(xs + 64, &st, 4);
        }
        if (collective) {
            // the challenge is broadcast together with rank 0's status, so that a failure there ends every rank
            int rb = be->broadcast(be->user, xs);
            if (rb) return rb;
            int32_t st;
            memcpy(&st, xs + 64, 4);
            if (st) return rank == 0 ? r : fail(st, "rank 0 failed in this round");
        }
        if (r) return r;
        return be->fold(be->user, xs, xs + 32);                                // :60-74 on the local shard
    };
    while (!rc && cur > 1 && cur / (size_t)world >= 2 && cur > collapse_at) {  // folds stay local while n >= 2 world
        rc = round(true);
        cur /= 2;
    }
    if (rc) return rc;
    if (rank == 0 && first) absorb_z();                                        // no local round ran (n == world, or n == 1)
    if (cur > 1) {
        rc = be->collapse(be->user);                                           // the tail moves to rank 0
        if (rc) return rc;
        while (rank == 0 && cur > 1) {
            rc = round(false);
            if (rc) return rc;
            cur /= 2;
        }
    }
    if (rank == 0)
        for (size_t i = 0; i < np; i++) memcpy(proof + 384 * i, &fwd[384 * (np - 1 - i)], 384);  // proof.reverse()  :78
    return SIPP_OK;
}

int check_shape(size_t n, int world) {
    if (!is_pow2(n)) return fail(SIPP_ERR_ARG, "n must be a non-zero power of two");
    if (world < 1 || (world & (world - 1))) return fail(SIPP_ERR_ARG, "the number of ranks must be a power of two");
    if (n < (size_t)world) return fail(SIPP_ERR_ARG, "n must be at least the number of ranks");
    return SIPP_OK;
}

int prove_sharded(sipp_ctx* c, int create_rc, size_t n, const uint8_t* A_full, const uint8_t* B_full, uint8_t* proof, AbsorbJob* job) {
    if (g_comm.world == 1 && job) {  // a world of one is the single-GPU prover
        int rc1 = create_rc ? create_rc : prove_core(c, *job, proof);
        if (c) {
            cudaStreamSynchronize(g_stream);
            sipp_ctx_destroy(c);
        }
        return rc1;
    }
    CudaBackend b;
    b.ctx = c;
    int rc = cb_agree(&b, create_rc);
    if (!rc) {
        sipp_shard_backend be;
        be.user = &b; be.rank = g_comm.rank; be.world = g_comm.world;
        be.local_len = cb_local_len; be.products = cb_products; be.combine = cb_combine; be.broadcast = cb_broadcast;
        be.fold = cb_fold; be.collapse = cb_collapse;
        // rank 0 runs the look-ahead stages and the pairing-matrix tail (k_mat.cu) alone: collapse where they begin
        size_t at = g_opt_pipeline && g_opt_fe_engine ? (size_t)(g_opt_matrix_block_n > g_opt_matrix_n ? g_opt_matrix_block_n : g_opt_matrix_n) : 0;
        if (at > COLLAPSE_MAX) at = COLLAPSE_MAX;
        b.collapse_at = at;
        rc = sharded_protocol(&be, n, A_full, B_full, proof, job, at);
    }
    if (b.ctx) {
        cudaStreamSynchronize(g_stream);
        sipp_ctx_destroy(b.ctx);
    }
    return rc;
}

int sharded_args(size_t n, const uint8_t* A_full, const uint8_t* B_full, uint8_t* proof) {
    int rc = ensure_init();
    if (rc) return rc;
    if (g_comm.kind == COMM_NONE) {  // no communicator: a world of one
        g_comm.rank = 0;
        g_comm.world = 1;
        if (!g_comm.d_gather && (rc = comm_buffers())) return rc;
    }
    rc = check_shape(n, g_comm.world);
    if (rc) return rc;
    if (g_comm.rank == 0 && (!A_full || !B_full || !proof)) return fail(SIPP_ERR_ARG, "rank 0 needs the full A, B (transcript) and the proof buffer");
    return SIPP_OK;
}

}  // namespace

extern "C" {

int sipp_comm_get_unique_id(uint8_t id[SIPP_COMM_ID_BYTES]) {
    if (!id) return fail(SIPP_ERR_ARG, "null argument");
    int rc = nccl_load();
    if (rc) return rc;
    nccl_unique_id u;
    NK(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, SIPP_COMM_ID_BYTES);
    return SIPP_OK;
}

int sipp_comm_destroy(void) {
    if (g_comm.kind == COMM_NCCL && g_comm.nccl) g_nccl.CommDestroy(g_comm.nccl);
    if (g_comm.d_gather) cudaFree(g_comm.d_gather);
    if (g_comm.d_xs) cudaFree(g_comm.d_xs);
    if (g_comm.h_stage) cudaFreeHost(g_comm.h_stage);
    g_comm = Comm();
    return SIPP_OK;
}

int sipp_comm_init(const uint8_t id[SIPP_COMM_ID_BYTES], int rank, int world) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!id || world < 1 || rank < 0 || rank >= world || (world & (world - 1))) return fail(SIPP_ERR_ARG, "bad rank / world (a power of two)");
    rc = nccl_load();
    if (rc) return rc;
    sipp_comm_destroy();
    nccl_unique_id u;
    memcpy(u.internal, id, SIPP_COMM_ID_BYTES);
    nccl_comm comm = nullptr;
    NK(g_nccl.CommInitRank(&comm, world, u, rank));  // binds to the CUDA device selected by sipp_init
    g_comm.kind = COMM_NCCL; g_comm.rank = rank; g_comm.world = world; g_comm.nccl = comm;
    rc = comm_buffers();
    if (rc) sipp_comm_destroy();
    return rc;
}

int sipp_comm_init_host(int rank, int world, sipp_allgather_fn allgather, sipp_broadcast_fn broadcast, void* user) {
    int rc = ensure_init();
    if (rc) return rc;
    if (world < 1 || rank < 0 || rank >= world || (world & (world - 1)) || (world > 1 && (!allgather || !broadcast)))
        return fail(SIPP_ERR_ARG, "bad rank / world (a power of two) or missing callback");
    sipp_comm_destroy();
    g_comm.kind = COMM_HOST; g_comm.rank = rank; g_comm.world = world;
    g_comm.h_allgather = allgather; g_comm.h_broadcast = broadcast; g_comm.user = user;
    rc = comm_buffers();
    if (rc) sipp_comm_destroy();
    return rc;
}

int sipp_comm_rank(void) { return g_comm.kind == COMM_NONE ? 0 : g_comm.rank; }
int sipp_comm_world(void) { return g_comm.kind == COMM_NONE ? 1 : g_comm.world; }
int sipp_comm_nccl_version(void) {
    int v = 0;
    if (nccl_load() || g_nccl.GetVersion(&v)) return 0;
    return v;
}

int sipp_prove_native_sharded(const uint8_t* A_local, const uint8_t* B_local, size_t n, const uint8_t* A_full, const uint8_t* B_full, uint8_t* proof) {
    int rc = sharded_args(n, A_full, B_full, proof);
    if (rc) return rc;
    AbsorbJob job;  // rank 0: the hash chain starts before the upload of the shard
    if (g_comm.rank == 0) job.start(A_full, B_full, n);
    sipp_ctx* c = nullptr;
    rc = (A_local && B_local) ? sipp_ctx_create(A_local, B_local, n / (size_t)g_comm.world, &c) : fail(SIPP_ERR_ARG, "null shard");
    return prove_sharded(c, rc, n, A_full, B_full, proof, g_comm.rank == 0 ? &job : nullptr);
}

int sipp_prove_native_sharded_device(const void* dA_local, const void* dB_local, size_t n, const uint8_t* A_full, const uint8_t* B_full, uint8_t* proof) {
    int rc = sharded_args(n, A_full, B_full, proof);
    if (rc) return rc;
    AbsorbJob job;
    if (g_comm.rank == 0) job.start(A_full, B_full, n);
    sipp_ctx* c = nullptr;
    rc = (dA_local && dB_local) ? sipp_ctx_create_from_device(dA_local, dB_local, n / (size_t)g_comm.world, &c) : fail(SIPP_ERR_ARG, "null shard");
    return prove_sharded(c, rc, n, A_full, B_full, proof, g_comm.rank == 0 ? &job : nullptr);
}

int sipp_prove_native_sharded_backend(const sipp_shard_backend* be, size_t n, const uint8_t* A_full, const uint8_t* B_full, uint8_t* proof) {
    if (!be || !be->products || !be->combine || !be->broadcast || !be->fold || !be->collapse) return fail(SIPP_ERR_ARG, "incomplete backend");
    if (be->rank < 0 || be->rank >= be->world) return fail(SIPP_ERR_ARG, "bad rank");
    int rc = check_shape(n, be->world);
    if (rc) return rc;
    if (be->rank == 0 && (!A_full || !B_full || !proof)) return fail(SIPP_ERR_ARG, "rank 0 needs the full A, B (transcript) and the proof buffer");
    return sharded_protocol(be, n, A_full, B_full, proof, nullptr, 0);  // a caller's backend: one pair per rank, as documented
}

}  // extern "C"
