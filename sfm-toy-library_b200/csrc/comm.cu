// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL communicator owned by the context.
//
// The reference has no distributed code at all; the only exchange step this library adds is the
// sum over ranks of the reduced camera system of bundle adjustment (ba.cu).  NCCL is resolved with dlopen at
// sfmb200_comm_init() time: inside a Python process that already imported torch this binds to the NCCL torch loaded
// (same soname, libnccl.so.2); in a plain C++ host it binds to the system libnccl.  No link-time dependency, so the
// library loads (and every single-GPU entry point works) on machines without NCCL.
#include "common.cuh"
#include <dlfcn.h>

namespace {
// the slice of the NCCL ABI we use (stable since NCCL 2.0)
typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { NCCL_SUCCESS = 0, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(nccl_comm_t*, int, nccl_unique_id, int);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
    void* lib = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_get_error_string get_error_string = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
        get_unique_id = (fn_get_unique_id)dlsym(lib, "ncclGetUniqueId");
        comm_init_rank = (fn_comm_init_rank)dlsym(lib, "ncclCommInitRank");
        comm_destroy = (fn_comm_destroy)dlsym(lib, "ncclCommDestroy");
        all_reduce = (fn_all_reduce)dlsym(lib, "ncclAllReduce");
        all_gather = (fn_all_gather)dlsym(lib, "ncclAllGather");
        get_error_string = (fn_get_error_string)dlsym(lib, "ncclGetErrorString");
        if (!get_unique_id || !comm_init_rank || !comm_destroy || !all_reduce) { err = "libnccl lacks a required symbol"; lib = nullptr; return false; }
        return true;
    }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
}  // namespace

struct CommState { nccl_comm_t comm = nullptr; };

static int allreduce_f64(sfmb200_ctx* ctx, double* dbuf, size_t n, int op) {
    if (ctx->nranks <= 1 || n == 0) return SFMB200_OK;
    if (!ctx->comm || !ctx->comm->comm) return sfmb200_fail(ctx, SFMB200_ERR_COMM, "communicator not initialised");
    const int rc = g_nccl.all_reduce(dbuf, dbuf, n, NCCL_FLOAT64, op, ctx->comm->comm, ctx->stream);
    if (rc != NCCL_SUCCESS)
        return sfmb200_fail(ctx, SFMB200_ERR_COMM, "ncclAllReduce: %s", g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "error");
    return SFMB200_OK;
}
int sfmb200_allreduce_sum_f64(sfmb200_ctx* ctx, double* dbuf, size_t n) { return allreduce_f64(ctx, dbuf, n, NCCL_SUM); }
// all-gather of `bytes` bytes per rank (device buffers, rank order) on the ctx stream: used once per problem to hand the CUDA-IPC
// handles of the exchange buffers around when the one-shot solve attaches its peers itself
int sfmb200_allgather_bytes(sfmb200_ctx* ctx, const void* d_send, void* d_recv, size_t bytes) {
    if (ctx->nranks <= 1) return SFMB200_OK;
    if (!ctx->comm || !ctx->comm->comm || !g_nccl.all_gather) return sfmb200_fail(ctx, SFMB200_ERR_COMM, "communicator not initialised");
    const int rc = g_nccl.all_gather(d_send, d_recv, bytes, /* ncclChar */ 0, ctx->comm->comm, ctx->stream);
    if (rc != NCCL_SUCCESS)
        return sfmb200_fail(ctx, SFMB200_ERR_COMM, "ncclAllGather: %s", g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "error");
    return SFMB200_OK;
}
int sfmb200_allreduce_max_f64(sfmb200_ctx* ctx, double* dbuf, size_t n) { return allreduce_f64(ctx, dbuf, n, NCCL_MAX); }

void sfmb200_comm_destroy(sfmb200_ctx* ctx) {
    if (ctx->comm) {
        if (ctx->comm->comm && g_nccl.comm_destroy) g_nccl.comm_destroy(ctx->comm->comm);
        delete ctx->comm; ctx->comm = nullptr;
    }
    ctx->rank = 0; ctx->nranks = 1;
}

extern "C" {

int sfmb200_comm_unique_id(uint8_t* id) {
    if (!id) return SFMB200_ERR_INVALID;
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!g_nccl.load()) return sfmb200_fail(nullptr, SFMB200_ERR_COMM, "%s", g_nccl.err.c_str());
    nccl_unique_id u;
    const int rc = g_nccl.get_unique_id(&u);
    if (rc != NCCL_SUCCESS) return sfmb200_fail(nullptr, SFMB200_ERR_COMM, "ncclGetUniqueId failed (%d)", rc);
    static_assert(sizeof(u) == SFMB200_UNIQUE_ID_BYTES, "unique id size");
    memcpy(id, &u, sizeof u);
    return SFMB200_OK;
}

int sfmb200_comm_init(sfmb200_ctx* ctx, const uint8_t* id, int rank, int nranks) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return SFMB200_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    sfmb200_comm_destroy(ctx);
    if (nranks == 1) return SFMB200_OK;
    if (!id) return sfmb200_fail(ctx, SFMB200_ERR_INVALID, "unique id required for nranks > 1");
    {
        std::lock_guard<std::mutex> lk2(g_nccl_mu);
        if (!g_nccl.load()) return sfmb200_fail(ctx, SFMB200_ERR_COMM, "%s", g_nccl.err.c_str());
    }
    SFM_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_unique_id u; memcpy(&u, id, sizeof u);
    CommState* cs = new CommState();
    const int rc = g_nccl.comm_init_rank(&cs->comm, nranks, u, rank);
    if (rc != NCCL_SUCCESS) {
        delete cs;
        return sfmb200_fail(ctx, SFMB200_ERR_COMM, "ncclCommInitRank: %s", g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "error");
    }
    ctx->comm = cs; ctx->rank = rank; ctx->nranks = nranks;
    return SFMB200_OK;
}

int sfmb200_comm_rank(const sfmb200_ctx* ctx) { return ctx ? ctx->rank : 0; }
int sfmb200_comm_size(const sfmb200_ctx* ctx) { return ctx ? ctx->nranks : 1; }

}  // extern "C"
