// comm.cu -- inter-rank plumbing (NCCL over NVLink).  Replaces the reference's host MPI +
// pinned staging (UM/comm_meso.cu:135-147,385-402; src/comm.cpp:686-753; MPI_Allreduce in
// UM/compute_temp_meso.cu:97).
#include "internal.h"
#include <nccl.h>
#include <cstring>

namespace meso {

#define MESO_NCCL(call)                                                                   \
    do {                                                                                  \
        ncclResult_t r_ = (call);                                                         \
        if (r_ != ncclSuccess) {                                                          \
            ctx->err = std::string(#call) + ": " + ncclGetErrorString(r_);                \
            return MESO_ENCCL;                                                            \
        }                                                                                 \
    } while (0)

int comm_init(meso_ctx *ctx, const void *nccl_id)
{
    if (ctx->nccl) return MESO_OK;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    memcpy(&id, nccl_id, sizeof id);
    MESO_CUDA(cudaSetDevice(ctx->device));
    ncclComm_t comm;
    MESO_NCCL(ncclCommInitRank(&comm, ctx->nranks, id, ctx->rank));
    ctx->nccl = comm;
    return MESO_OK;
}

void comm_destroy(meso_ctx *ctx)
{
    if (ctx->nccl) { ncclCommDestroy((ncclComm_t)ctx->nccl); ctx->nccl = nullptr; }
}

// small host-visible reductions (thermo scalars, global atom count)
int comm_allreduce_sum(meso_ctx *ctx, double *host_vals, int n)
{
    if (ctx->nranks == 1) return MESO_OK;
    if (!ctx->nccl) { ctx->err = "communicator not initialised"; return MESO_ENCCL; }
    if (!ctx->partial.reserve(64)) { ctx->err = "out of device memory"; return MESO_ECUDA; }
    MESO_CUDA(cudaMemcpyAsync(ctx->partial.p, host_vals, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    MESO_NCCL(ncclAllReduce(ctx->partial.p, ctx->partial.p, n, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
    MESO_CUDA(cudaMemcpyAsync(host_vals, ctx->partial.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return MESO_OK;
}

}  // namespace meso

extern "C" int meso_comm_unique_id(void *id128)
{
    if (!id128) return MESO_EINVAL;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return MESO_ENCCL;
    memcpy(id128, &id, sizeof id);
    return MESO_OK;
}
