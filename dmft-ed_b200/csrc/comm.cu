// comm.cu -- the one collective of the sharded Lanczos path: a sum of a few doubles over the ranks (one process per GPU).
//
// Sharding by conserved occupation pairs (hxv_fiber.cu) leaves H*v without any exchange; what remains of
// directMatVec_MPI_cc's communication (ED_HAMILTONIAN_DIRECT_HxV.f90:163-166, and the MPI_Allreduce of the dot products in
// SciFortran's MPI Lanczos) is the reduction of the Lanczos scalars.  NCCL is loaded at run time (dlopen of libnccl.so.2:
// the copy torch already mapped in a Python host, the system one under a Fortran/MPI host), so single-GPU users carry no
// NCCL dependency.  The unique id is created on rank 0 and broadcast by the HOST (MPI_Bcast / torch.distributed).
#include "edgpu_internal.h"
#include <dlfcn.h>
#include <cstring>

namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(nccl_comm *, int, nccl_uid, int);
typedef int (*fn_destroy)(nccl_comm);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t);
typedef const char *(*fn_errstr)(int);
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
    void *lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_errstr errstr = nullptr;
};

NcclApi *nccl_api(edgpu_ctx *ctx)
{
    static NcclApi api;
    if (api.lib) return &api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) { edgpu_fail(ctx, "sharded path: cannot load libnccl.so.2 (%s)", dlerror()); return nullptr; }
    api.get_uid = (fn_get_uid)dlsym(api.lib, "ncclGetUniqueId");
    api.init_rank = (fn_init_rank)dlsym(api.lib, "ncclCommInitRank");
    api.destroy = (fn_destroy)dlsym(api.lib, "ncclCommDestroy");
    api.allreduce = (fn_allreduce)dlsym(api.lib, "ncclAllReduce");
    api.errstr = (fn_errstr)dlsym(api.lib, "ncclGetErrorString");
    if (!api.get_uid || !api.init_rank || !api.destroy || !api.allreduce) {
        edgpu_fail(ctx, "sharded path: libnccl lacks an expected symbol");
        api.lib = nullptr;
        return nullptr;
    }
    return &api;
}
}   // namespace

struct EdComm {
    nccl_comm comm = nullptr;
    int rank = 0, nranks = 1;
};

extern "C" int edgpu_comm_unique_id(edgpu_ctx *ctx, unsigned char id[128])
{
    if (!ctx || !id) return 1;
    NcclApi *api = nccl_api(ctx);
    if (!api) return 1;
    nccl_uid u;
    const int r = api->get_uid(&u);
    if (r != 0) return edgpu_fail(ctx, "ncclGetUniqueId failed: %s", api->errstr ? api->errstr(r) : "?");
    memcpy(id, &u, 128);
    return 0;
}

extern "C" int edgpu_comm_init(edgpu_ctx *ctx, const unsigned char id[128], int32_t rank, int32_t nranks)
{
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return ctx ? edgpu_fail(ctx, "edgpu_comm_init: bad arguments") : 1;
    NcclApi *api = nccl_api(ctx);
    if (!api) return 1;
    if (ctx->comm) return edgpu_fail(ctx, "edgpu_comm_init: communicator already initialised");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    nccl_uid u;
    memcpy(&u, id, 128);
    auto *c = new EdComm();
    c->rank = rank; c->nranks = nranks;
    const int r = api->init_rank(&c->comm, nranks, u, rank);
    if (r != 0) { delete c; return edgpu_fail(ctx, "ncclCommInitRank failed: %s", api->errstr ? api->errstr(r) : "?"); }
    ctx->comm = c;
    return 0;
}

extern "C" int edgpu_comm_finalize(edgpu_ctx *ctx)
{
    if (!ctx || !ctx->comm) return 0;
    NcclApi *api = nccl_api(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (api && ctx->comm->comm) api->destroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
    return 0;
}

int comm_nranks(const edgpu_ctx *ctx) { return ctx->comm ? ctx->comm->nranks : 1; }
int comm_rank(const edgpu_ctx *ctx) { return ctx->comm ? ctx->comm->rank : 0; }

extern "C" int edgpu_comm_info(const edgpu_ctx *ctx, int32_t *rank, int32_t *nranks)
{
    if (!ctx) return 1;
    if (rank) *rank = comm_rank(ctx);
    if (nranks) *nranks = comm_nranks(ctx);
    return 0;
}

/* host[0..n) <- reduction over the ranks (op 0: sum, 1: min, 2: max); staged through device memory on the context stream */
extern "C" int edgpu_comm_allreduce_host(edgpu_ctx *ctx, double *host, int64_t n, int32_t op)
{
    if (!ctx || !host || n < 0) return 1;
    if (!ctx->comm || ctx->comm->nranks == 1 || n == 0) return 0;
    NcclApi *api = nccl_api(ctx);
    if (!api) return 1;
    double *d = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d, sizeof(double) * (size_t)n));
    cudaError_t e = cudaMemcpyAsync(d, host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
    int r = 0;
    if (e == cudaSuccess) r = api->allreduce(d, d, (size_t)n, kNcclFloat64, op == 1 ? 3 /* ncclMin */ : op == 2 ? 2 /* ncclMax */ : kNcclSum, ctx->comm->comm, ctx->stream);
    if (e == cudaSuccess && r == 0) e = cudaMemcpyAsync(host, d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (r != 0) return edgpu_fail(ctx, "ncclAllReduce failed: %s", api->errstr ? api->errstr(r) : "?");
    if (e != cudaSuccess || e2 != cudaSuccess) return edgpu_fail(ctx, "edgpu_comm_allreduce_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return 0;
}

// in-place sum of n doubles at d_buf over all ranks, ordered on the context stream
int comm_allreduce_sum(edgpu_ctx *ctx, double *d_buf, int n)
{
    if (!ctx->comm || ctx->comm->nranks == 1) return 0;
    NcclApi *api = nccl_api(ctx);
    if (!api) return 1;
    const int r = api->allreduce(d_buf, d_buf, (size_t)n, kNcclFloat64, kNcclSum, ctx->comm->comm, ctx->stream);
    if (r != 0) return edgpu_fail(ctx, "ncclAllReduce failed: %s", api->errstr ? api->errstr(r) : "?");
    return 0;
}
