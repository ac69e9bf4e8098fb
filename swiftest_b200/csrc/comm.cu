// comm.cu -- multi-GPU exchange for the i-sliced pl-pl path: one process (context) per GPU, an NCCL allgather
// of the drifted position (and velocity) slices over NVLink each step.  The reference has no counterpart
// (its only multi-process strategy is Coarray test-particle sharding, swiftest_coarray.f90:672-733, which
// needs no per-step traffic and maps to tp block partitioning here).
//
// NCCL is loaded with dlopen at swcu_comm_init so a single-GPU run has no NCCL dependency; in a torchrun
// job the already-loaded libnccl.so.2 of PyTorch is picked up.
#include "swcu_internal.cuh"

#include <dlfcn.h>
#include <string.h>

namespace swcu {

typedef struct { char internal[SWCU_NCCL_ID_BYTES]; } NcclUniqueId;
typedef int NcclResult;
typedef void *NcclComm;
constexpr int NCCL_FLOAT64 = 8;  // ncclFloat64 / ncclDouble in every NCCL 2.x

struct NcclApi {
    void *handle = nullptr;
    NcclResult (*GetUniqueId)(NcclUniqueId *) = nullptr;
    NcclResult (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    NcclResult (*CommDestroy)(NcclComm) = nullptr;
    NcclResult (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    NcclResult (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(NcclResult) = nullptr;
};

namespace {

int load_nccl(swcu_context *ctx)
{
    if (ctx->nccl) return SWCU_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(ctx, SWCU_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    NcclApi *api = new NcclApi;
    api->handle = h;
    api->GetUniqueId = (NcclResult(*)(NcclUniqueId *))dlsym(h, "ncclGetUniqueId");
    api->CommInitRank = (NcclResult(*)(NcclComm *, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
    api->CommDestroy = (NcclResult(*)(NcclComm))dlsym(h, "ncclCommDestroy");
    api->AllGather =
        (NcclResult(*)(const void *, void *, size_t, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
    api->AllReduce =
        (NcclResult(*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllReduce");
    api->GetErrorString = (const char *(*)(NcclResult))dlsym(h, "ncclGetErrorString");
    if (!api->GetUniqueId || !api->CommInitRank || !api->CommDestroy || !api->AllGather || !api->AllReduce || !api->GetErrorString) {
        delete api;
        return fail(ctx, SWCU_ERR_NCCL, "libnccl is missing a required symbol");
    }
    ctx->nccl = api;
    return SWCU_OK;
}

#define SWCU_NCCL(ctx, call)                                                                              \
    do {                                                                                                  \
        swcu::NcclResult r__ = (call);                                                                    \
        if (r__ != 0)                                                                                     \
            return swcu::fail((ctx), SWCU_ERR_NCCL, "%s failed: %s", #call, (ctx)->nccl->GetErrorString(r__)); \
    } while (0)

// pack this rank's slice of up to six SoA arrays into one contiguous send buffer [narr][maxcount]
__global__ void pack_slice_kernel(int narr, int i0, int cnt, int maxcount, const double *a0, const double *a1,
                                  const double *a2, const double *a3, const double *a4, const double *a5, double *send)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= maxcount) return;
    const double *arr[6] = {a0, a1, a2, a3, a4, a5};
    for (int c = 0; c < narr; ++c) send[(size_t)c * maxcount + t] = (t < cnt) ? arr[c][i0 + t] : 0.0;
}

// scatter the gathered [rank][narr][maxcount] buffer back into the SoA arrays (other ranks' slices only)
__global__ void unpack_slices_kernel(int narr, int n, int nranks, int myrank, int maxcount, const double *recv, double *a0,
                                     double *a1, double *a2, double *a3, double *a4, double *a5)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // balanced partition: the first (n % nranks) ranks own one more row
    const int q = n / nranks, r = n % nranks;
    int rank, off;
    if (i < (q + 1) * r) {
        rank = i / (q + 1);
        off = i - rank * (q + 1);
    } else {
        rank = r + (i - (q + 1) * r) / (q > 0 ? q : 1);
        off = (i - (q + 1) * r) - (rank - r) * q;
    }
    if (rank == myrank) return;
    double *arr[6] = {a0, a1, a2, a3, a4, a5};
    for (int c = 0; c < narr; ++c) arr[c][i] = recv[((size_t)rank * narr + c) * maxcount + off];
}

}  // namespace

int comm_allgather_pl(swcu_context *ctx, int with_v)
{
    if (ctx->nranks <= 1) return SWCU_OK;
    if (!ctx->comm) return fail(ctx, SWCU_ERR_STATE, "swcu_pl_allgather: communicator not initialised");
    Body &pl = ctx->pl;
    if (!pl.valid) return fail(ctx, SWCU_ERR_STATE, "swcu_pl_allgather: pl population not resident");
    int e0, e1;
    swcu_partition(pl.n, ctx->nranks, ctx->rank, &e0, &e1);
    if (e0 != pl.slice0 || e1 != pl.slice1)
        return fail(ctx, SWCU_ERR_STATE, "swcu_pl_allgather: slice [%d,%d) is not the balanced partition [%d,%d)",
                    pl.slice0, pl.slice1, e0, e1);
    FamTimer ft(ctx, FAM_ALLGATHER);
    const int narr = with_v ? 6 : 3;
    const int maxcount = (pl.n + ctx->nranks - 1) / ctx->nranks;
    SWCU_CUDA(ctx, ctx->sendbuf.ensure(sizeof(double) * (size_t)narr * maxcount));
    SWCU_CUDA(ctx, ctx->recvbuf.ensure(sizeof(double) * (size_t)narr * maxcount * ctx->nranks));
    pack_slice_kernel<<<cdiv(maxcount, 256), 256, 0, ctx->stream>>>(
        narr, pl.slice0, pl.slice1 - pl.slice0, maxcount, pl.rx.as<double>(), pl.ry.as<double>(), pl.rz.as<double>(),
        pl.vx.as<double>(), pl.vy.as<double>(), pl.vz.as<double>(), ctx->sendbuf.as<double>());
    SWCU_KERNEL_CHECK(ctx);
    SWCU_NCCL(ctx, ctx->nccl->AllGather(ctx->sendbuf.p, ctx->recvbuf.p, (size_t)narr * maxcount, NCCL_FLOAT64,
                                        (NcclComm)ctx->comm, ctx->stream));
    unpack_slices_kernel<<<cdiv(pl.n, 256), 256, 0, ctx->stream>>>(
        narr, pl.n, ctx->nranks, ctx->rank, maxcount, ctx->recvbuf.as<double>(), pl.rx.as<double>(), pl.ry.as<double>(),
        pl.rz.as<double>(), pl.vx.as<double>(), pl.vy.as<double>(), pl.vz.as<double>());
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

// in-place sum over ranks (ncclSum == 0) of a device buffer; used by the third-law kernel's pair slices
int comm_allreduce_sum(swcu_context *ctx, double *buf, size_t count)
{
    if (ctx->nranks <= 1) return SWCU_OK;
    if (!ctx->comm) return fail(ctx, SWCU_ERR_STATE, "allreduce: communicator not initialised");
    FamTimer ft(ctx, FAM_ALLGATHER);
    SWCU_NCCL(ctx, ctx->nccl->AllReduce(buf, buf, count, NCCL_FLOAT64, 0, (NcclComm)ctx->comm, ctx->stream));
    return SWCU_OK;
}

void comm_release(swcu_context *ctx)
{
    if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy((NcclComm)ctx->comm);
    ctx->comm = nullptr;
    if (ctx->nccl) {
        delete ctx->nccl;  // the library handle stays loaded for the life of the process
        ctx->nccl = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
}

}  // namespace swcu

using namespace swcu;

extern "C" int swcu_partition(int32_t n, int32_t nranks, int32_t rank, int32_t *i0, int32_t *i1)
{
    if (nranks <= 0 || rank < 0 || rank >= nranks || n < 0 || !i0 || !i1) return SWCU_ERR_ARG;
    const int q = n / nranks, r = n % nranks;
    *i0 = rank * q + (rank < r ? rank : r);
    *i1 = *i0 + q + (rank < r ? 1 : 0);
    return SWCU_OK;
}

extern "C" int swcu_comm_unique_id(swcu_context *ctx, void *id128)
{
    if (!ctx || !id128) return SWCU_ERR_ARG;
    SWCU_TRY(load_nccl(ctx));
    NcclUniqueId id;
    SWCU_NCCL(ctx, ctx->nccl->GetUniqueId(&id));
    memcpy(id128, &id, SWCU_NCCL_ID_BYTES);
    return SWCU_OK;
}

extern "C" int swcu_comm_init(swcu_context *ctx, int32_t nranks, int32_t rank, const void *id128)
{
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return SWCU_ERR_ARG;
    if (nranks == 1) {
        ctx->nranks = 1;
        ctx->rank = 0;
        return SWCU_OK;
    }
    if (!id128) return SWCU_ERR_ARG;
    SWCU_TRY(load_nccl(ctx));
    SWCU_CUDA(ctx, cudaSetDevice(ctx->device));
    NcclUniqueId id;
    memcpy(&id, id128, SWCU_NCCL_ID_BYTES);
    NcclComm comm = nullptr;
    SWCU_NCCL(ctx, ctx->nccl->CommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return SWCU_OK;
}

extern "C" int swcu_comm_finalize(swcu_context *ctx)
{
    if (!ctx) return SWCU_ERR_ARG;
    comm_release(ctx);
    return SWCU_OK;
}

extern "C" int swcu_pl_set_slice(swcu_context *ctx, int32_t i0, int32_t i1)
{
    if (!ctx) return SWCU_ERR_ARG;
    if (!ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "swcu_pl_set_slice: pl population not resident");
    if (i0 < 0 || i1 < i0 || i1 > ctx->pl.n) return fail(ctx, SWCU_ERR_ARG, "swcu_pl_set_slice: bad slice [%d,%d)", i0, i1);
    ctx->pl.slice0 = i0;
    ctx->pl.slice1 = i1;
    return SWCU_OK;
}

extern "C" int swcu_pl_allgather(swcu_context *ctx, int32_t with_v)
{
    if (!ctx) return SWCU_ERR_ARG;
    return comm_allgather_pl(ctx, with_v);
}

// ======================================================================================================
// Peer-memory exchange: CUDA-IPC export/import of the buffers the fused step touches on every rank
//   buffer 0: F (partial accelerations, 3*stride doubles)   buffers 1..6: pl rx,ry,rz,vx,vy,vz   buffer 7: flags
// ======================================================================================================
extern "C" int swcu_p2p_export(swcu_context *ctx, void *handles)
{
    if (!ctx || !handles) return SWCU_ERR_ARG;
    SWCU_CUDA(ctx, cudaSetDevice(ctx->device));
    Body &pl = ctx->pl;
    if (!pl.valid || pl.n <= 0) return fail(ctx, SWCU_ERR_STATE, "swcu_p2p_export: pl population not resident");
    auto &P = ctx->p2p;
    P.stride = ((size_t)pl.n + 31) & ~size_t(31);
    SWCU_CUDA(ctx, P.F.ensure(sizeof(double) * 3 * P.stride));
    SWCU_CUDA(ctx, P.flags.ensure(sizeof(unsigned long long) * 64));
    SWCU_CUDA(ctx, cudaMemsetAsync(P.flags.p, 0, sizeof(unsigned long long) * 64, ctx->stream));
    SWCU_CUDA(ctx, cudaMemsetAsync(P.F.p, 0, sizeof(double) * 3 * P.stride, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    void *bufs[SWCU_P2P_NBUF] = {P.F.p, pl.rx.p, pl.ry.p, pl.rz.p, pl.vx.p, pl.vy.p, pl.vz.p, P.flags.p};
    static_assert(sizeof(cudaIpcMemHandle_t) == SWCU_IPC_HANDLE_BYTES, "IPC handle size");
    for (int b = 0; b < SWCU_P2P_NBUF; ++b) {
        cudaIpcMemHandle_t h;
        SWCU_CUDA(ctx, cudaIpcGetMemHandle(&h, bufs[b]));
        memcpy((char *)handles + (size_t)b * SWCU_IPC_HANDLE_BYTES, &h, SWCU_IPC_HANDLE_BYTES);
    }
    return SWCU_OK;
}

extern "C" int swcu_p2p_import(swcu_context *ctx, int32_t nranks, int32_t rank, const void *all_handles)
{
    if (!ctx || !all_handles || nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) return SWCU_ERR_ARG;
    SWCU_CUDA(ctx, cudaSetDevice(ctx->device));
    auto &P = ctx->p2p;
    Body &pl = ctx->pl;
    if (!pl.valid || P.F.p == nullptr) return fail(ctx, SWCU_ERR_STATE, "swcu_p2p_import: call swcu_p2p_export first");
    void *own[SWCU_P2P_NBUF] = {P.F.p, pl.rx.p, pl.ry.p, pl.rz.p, pl.vx.p, pl.vy.p, pl.vz.p, P.flags.p};
    for (int r = 0; r < nranks; ++r) {
        for (int b = 0; b < SWCU_P2P_NBUF; ++b) {
            if (r == rank) {
                P.peer[r][b] = own[b];
                continue;
            }
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char *)all_handles + ((size_t)r * SWCU_P2P_NBUF + b) * SWCU_IPC_HANDLE_BYTES,
                   SWCU_IPC_HANDLE_BYTES);
            void *p = nullptr;
            SWCU_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            P.peer[r][b] = p;
        }
    }
    P.nranks = nranks;
    P.rank = rank;
    P.epoch = 0;
    if (!P.h_err) SWCU_CUDA(ctx, cudaHostAlloc((void **)&P.h_err, sizeof(unsigned long long), cudaHostAllocDefault));
    *P.h_err = 0ull;
    P.ready = true;
    return SWCU_OK;
}

namespace swcu {
// A bounded spin of the peer-memory exchange ran out on some rank (a peer died, hung or was never launched): the
// kernels raised the error word on every rank instead of hanging; turn it into a status and clear it.
int p2p_check_error(swcu_context *ctx)
{
    auto &P = ctx->p2p;
    if (!P.ready || !P.h_err || *P.h_err == 0ull) return SWCU_OK;
    *P.h_err = 0ull;
    cudaMemsetAsync((unsigned long long *)P.peer[P.rank][7] + 32, 0, sizeof(unsigned long long), ctx->stream);
    return fail(ctx, SWCU_ERR_STATE, "peer-memory exchange timed out (rank %d of %d, epoch %llu): a peer never signalled; "
                                     "the resident pl state is not valid", P.rank, P.nranks, P.epoch);
}
}  // namespace swcu

extern "C" int swcu_p2p_close(swcu_context *ctx)
{
    if (!ctx) return SWCU_ERR_ARG;
    auto &P = ctx->p2p;
    if (P.ready) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        for (int r = 0; r < P.nranks; ++r)
            if (r != P.rank)
                for (int b = 0; b < SWCU_P2P_NBUF; ++b)
                    if (P.peer[r][b]) cudaIpcCloseMemHandle(P.peer[r][b]);
    }
    const int rc = P.ready ? p2p_check_error(ctx) : SWCU_OK;
    P.ready = false;
    P.nranks = 1;
    P.rank = 0;
    if (P.h_err) cudaFreeHost(P.h_err);
    P.h_err = nullptr;
    return rc;
}

extern "C" int swcu_pl_kick_drift_p2p(swcu_context *ctx, int32_t lclose, double dt, int32_t *nfail)
{
    if (!ctx) return SWCU_ERR_ARG;
    SWCU_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->p2p.ready) return fail(ctx, SWCU_ERR_STATE, "swcu_pl_kick_drift_p2p: peer buffers not imported");
    Body &pl = ctx->pl;
    if (!pl.valid) return fail(ctx, SWCU_ERR_STATE, "swcu_pl_kick_drift_p2p: pl population not resident");
    SWCU_TRY(p2p_check_error(ctx));  // a timeout reported by an earlier step
    SWCU_TRY(kick_pl_flat(ctx, pl, lclose != 0, pl.nplm, /*reduce=*/false));
    return p2p_step_after_kick(ctx, dt, nfail);
}
