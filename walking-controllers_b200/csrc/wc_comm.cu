// wc_comm.cu — result collection across the GPUs of one box (SURVEY.md §8(b) `wc_gather`, §8(e)).
//
// Controller instances are independent: rank g owns a contiguous instance range and no collective sits on the tick path.
// What a Monte-Carlo driver needs is, at its collection points, (1) the per-instance results of every rank
// {u_0 (2 doubles), qdot (nJ doubles), OSQP status, iterations, QP status, working-set changes} and (2) a small statistics
// vector (status histograms, sums / maxima of iterations, violation counts) reduced over the ranks.  Both run over NCCL
// (NVLink 5 / NVSwitch) on the communicator's OWN side stream, after an event on the stream that produced the results, so the
// next tick's kernels overlap the collection.
//
// NCCL is bound at run time (dlopen("libnccl.so.2")): the library has no link-time dependency on it, a process that never
// creates a wc_comm never loads it, and a process that already loaded an NCCL (e.g. the one torch bundles) shares it.
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <new>

#include <cuda_runtime.h>

#include "../../include/wc_capi.h"

namespace {

// the handful of NCCL entry points used, with NCCL's own signatures (nccl.h, 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2, ncclFloat64 = 8 };  // ncclDataType_t
enum { ncclSum = 0, ncclMax = 2 };        // ncclRedOp_t

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

Nccl& nccl()
{
    static Nccl n;
    if (n.lib) return n;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (n.lib) break;
    }
    if (!n.lib) { std::fprintf(stderr, "[wc_comm] libnccl.so.2 not found: %s\n", dlerror()); return n; }
#define BIND(field, sym) *(void**)(&n.field) = dlsym(n.lib, sym)
    BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllGather, "ncclAllGather"); BIND(AllReduce, "ncclAllReduce"); BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather && n.AllReduce && n.GroupStart && n.GroupEnd;
    return n;
}

} // namespace

struct wc_comm {
    ncclComm_t comm;
    int world, rank, device;
    cudaStream_t side;
    cudaEvent_t ready, done;
    int64_t collectives;
};

#define NCHECK(x) do { const ncclResult_t r_ = (x); if (r_ != 0) { \
    std::fprintf(stderr, "[wc_comm] NCCL error %d (%s) at %s\n", r_, N.GetErrorString ? N.GetErrorString(r_) : "?", #x); return WC_ERR_CUDA; } } while (0)

extern "C" {

wc_status wc_comm_unique_id(uint8_t* id)
{
    if (!id) return WC_ERR_INVALID_ARG;
    Nccl& N = nccl();
    if (!N.ok) return WC_ERR_UNSUPPORTED;
    ncclUniqueId u;
    NCHECK(N.GetUniqueId(&u));
    static_assert(sizeof(u) == WC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    std::memcpy(id, &u, sizeof(u));
    return WC_OK;
}

wc_status wc_comm_create(const uint8_t* id, int32_t world, int32_t rank, int32_t device, wc_comm** out)
{
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return WC_ERR_INVALID_ARG;
    *out = nullptr;
    Nccl& N = nccl();
    if (!N.ok) return WC_ERR_UNSUPPORTED;
    if (cudaSetDevice(device) != cudaSuccess) return WC_ERR_CUDA;
    wc_comm* c = new (std::nothrow) wc_comm();
    if (!c) return WC_ERR_ALLOC;
    c->world = world; c->rank = rank; c->device = device; c->collectives = 0; c->comm = nullptr;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming) != cudaSuccess) { delete c; return WC_ERR_CUDA; }
    const ncclResult_t r = N.CommInitRank(&c->comm, world, u, rank);
    if (r != 0) {
        std::fprintf(stderr, "[wc_comm_create] ncclCommInitRank failed: %s\n", N.GetErrorString ? N.GetErrorString(r) : "?");
        cudaStreamDestroy(c->side); cudaEventDestroy(c->ready); cudaEventDestroy(c->done); delete c;
        return WC_ERR_CUDA;
    }
    *out = c;
    return WC_OK;
}

void wc_comm_destroy(wc_comm* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->side);
    Nccl& N = nccl();
    if (c->comm && N.CommDestroy) N.CommDestroy(c->comm);
    cudaStreamDestroy(c->side); cudaEventDestroy(c->ready); cudaEventDestroy(c->done);
    delete c;
}

int32_t wc_comm_world(const wc_comm* c) { return c ? c->world : 0; }
int32_t wc_comm_rank(const wc_comm* c) { return c ? c->rank : -1; }

// All-gather of the per-instance results of this rank's B instances and all-reduce of the statistics vectors, as ONE NCCL
// group on the communicator's side stream, ordered after everything already enqueued on `producer_stream`.
wc_status wc_gather_async(wc_comm* c, const wc_gather_io* io, int32_t B, int32_t n_dofs, void* producer_stream)
{
    if (!c || !io || B < 1 || n_dofs < 0) return WC_ERR_INVALID_ARG;
    Nccl& N = nccl();
    if (!N.ok) return WC_ERR_UNSUPPORTED;
    if (cudaSetDevice(c->device) != cudaSuccess) return WC_ERR_CUDA;
    cudaStream_t ps = static_cast<cudaStream_t>(producer_stream);
    if (cudaEventRecord(c->ready, ps) != cudaSuccess || cudaStreamWaitEvent(c->side, c->ready, 0) != cudaSuccess) return WC_ERR_CUDA;
    const size_t Bz = (size_t)B;
    NCHECK(N.GroupStart());
    // inside the group an error must not return before ncclGroupEnd (the group would stay open on this thread)
    ncclResult_t bad = 0;
    auto ok = [&](ncclResult_t r) { if (r != 0 && bad == 0) bad = r; };
    if (io->zmp && io->zmp_all) ok(N.AllGather(io->zmp, io->zmp_all, 2 * Bz, ncclFloat64, c->comm, c->side));
    if (io->dq && io->dq_all) ok(N.AllGather(io->dq, io->dq_all, (size_t)n_dofs * Bz, ncclFloat64, c->comm, c->side));
    if (io->osqp_status && io->osqp_status_all) ok(N.AllGather(io->osqp_status, io->osqp_status_all, Bz, ncclInt32, c->comm, c->side));
    if (io->iters && io->iters_all) ok(N.AllGather(io->iters, io->iters_all, Bz, ncclInt32, c->comm, c->side));
    if (io->qp_status && io->qp_status_all) ok(N.AllGather(io->qp_status, io->qp_status_all, Bz, ncclInt32, c->comm, c->side));
    if (io->nwsr && io->nwsr_all) ok(N.AllGather(io->nwsr, io->nwsr_all, Bz, ncclInt32, c->comm, c->side));
    if (io->stats_sum && io->n_stats_sum > 0)
        ok(N.AllReduce(io->stats_sum, io->stats_sum, (size_t)io->n_stats_sum, ncclFloat64, ncclSum, c->comm, c->side));
    if (io->stats_max && io->n_stats_max > 0)
        ok(N.AllReduce(io->stats_max, io->stats_max, (size_t)io->n_stats_max, ncclFloat64, ncclMax, c->comm, c->side));
    ok(N.GroupEnd());
    if (bad != 0) {
        std::fprintf(stderr, "[wc_gather_async] NCCL error %d (%s)\n", bad, N.GetErrorString ? N.GetErrorString(bad) : "?");
        return WC_ERR_CUDA;
    }
    if (cudaEventRecord(c->done, c->side) != cudaSuccess) return WC_ERR_CUDA;
    c->collectives += 1;
    return WC_OK;
}

// make `consumer_stream` wait for the last wc_gather_async (consumer_stream == NULL: block the host instead)
wc_status wc_gather_wait(wc_comm* c, void* consumer_stream, int32_t block_host)
{
    if (!c) return WC_ERR_INVALID_ARG;
    if (cudaSetDevice(c->device) != cudaSuccess) return WC_ERR_CUDA;
    if (block_host) return cudaEventSynchronize(c->done) == cudaSuccess ? WC_OK : WC_ERR_CUDA;
    return cudaStreamWaitEvent(static_cast<cudaStream_t>(consumer_stream), c->done, 0) == cudaSuccess ? WC_OK : WC_ERR_CUDA;
}

} // extern "C"
