// c2d_comm.cu — the NCCL plumbing of a multi-GPU group (include/c2d_gpu.h "multi-GPU").
//
// One context per GPU, one communicator per group.  Two data-path collectives exist, both device to device over
// NVLink / NVSwitch and both enqueued on the context's own stream (no host staging, no torch):
//   * the move exchange of a frame: ncclAllGather of a 16-byte header per rank, then one grouped ncclBroadcast per rank
//     of exactly its 12-byte translation records, straight out of the rank's device log (c2d_gpu.cu: flush);
//   * the result gather: ncclAllGather of the pair counts, then one ncclBroadcast per rank (C2D_RESULTS_ALL) or
//     ncclSend / ncclRecv to rank 0 (C2D_RESULTS_ROOT) of the 84-byte records.
// Everything else of the decomposition (slab ownership, halo, owner rule) needs no communication: shape state is
// replicated, so every rank computes the same keys and splitters (c2d_slab.cuh).
//
// libnccl is resolved with dlopen at first use: a single-GPU application never needs it, and inside a torch process the
// already loaded libnccl.so.2 (same soname) is the one that answers.
#include "c2d_ctx.hpp"

#include <dlfcn.h>
#include <nccl.h> // types and prototypes only; no symbol of libnccl is linked

#include <mutex>

namespace
{

struct NcclApi
{
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    void* handle = nullptr;
    std::string error;
};

NcclApi& nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {std::getenv("C2D_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* name : names)
        {
            if (!name || !*name)
                continue;
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle)
                break;
            api.error = dlerror();
        }
        if (!api.handle)
        {
            api.error = "cannot load libnccl.so.2 (" + api.error + "); set C2D_NCCL_LIB to its path";
            return;
        }
        bool ok = true;
        auto sym = [&](const char* n) {
            void* p = dlsym(api.handle, n);
            if (!p)
            {
                ok = false;
                api.error = std::string("libnccl lacks ") + n;
            }
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        if (!ok)
        {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return api;
}

#define C2D_NCCL(ctx, call)                                                                                            \
    do                                                                                                                 \
    {                                                                                                                  \
        const ncclResult_t _r = (call);                                                                                \
        if (_r != ncclSuccess)                                                                                         \
            return ::c2d_host::fail(ctx, std::string(#call) + ": " + nccl_api().GetErrorString(_r));                   \
    } while (0)

inline ncclComm_t comm_of(c2d_gpu_ctx* ctx)
{
    return static_cast<ncclComm_t>(ctx->comm.comm);
}

} // namespace

int c2d_host::comm_allgather(c2d_gpu_ctx* ctx, const void* send, void* recv, size_t bytesPerRank)
{
    if (!ctx->comm.comm)
        return fail(ctx, "no communicator: call c2d_gpu_comm_init first");
    C2D_NCCL(ctx, nccl_api().AllGather(send, recv, bytesPerRank, ncclChar, comm_of(ctx), ctx->stream));
    return 0;
}

int c2d_host::comm_allgather2(c2d_gpu_ctx* ctx, const void* sendA, void* recvA, size_t bytesA, const void* sendB, void* recvB,
                              size_t bytesB)
{
    if (!ctx->comm.comm)
        return fail(ctx, "no communicator: call c2d_gpu_comm_init first");
    NcclApi& api = nccl_api();
    C2D_NCCL(ctx, api.GroupStart());
    C2D_NCCL(ctx, api.AllGather(sendA, recvA, bytesA, ncclChar, comm_of(ctx), ctx->stream));
    if (bytesB)
        C2D_NCCL(ctx, api.AllGather(sendB, recvB, bytesB, ncclChar, comm_of(ctx), ctx->stream));
    C2D_NCCL(ctx, api.GroupEnd());
    return 0;
}

int c2d_host::comm_gatherv(c2d_gpu_ctx* ctx, const void* send, void* recvBase, const size_t* bytes, size_t strideBytes)
{
    if (!ctx->comm.comm)
        return fail(ctx, "no communicator: call c2d_gpu_comm_init first");
    NcclApi& api = nccl_api();
    C2D_NCCL(ctx, api.GroupStart());
    for (uint32_t r = 0; r < ctx->shardWorld; ++r)
    {
        if (!bytes[r])
            continue;
        char* dst = static_cast<char*>(recvBase) + static_cast<size_t>(r) * strideBytes;
        C2D_NCCL(ctx, api.Broadcast(r == ctx->shardRank ? send : dst, dst, bytes[r], ncclChar, static_cast<int>(r), comm_of(ctx),
                                    ctx->stream));
    }
    C2D_NCCL(ctx, api.GroupEnd());
    return 0;
}

// The records of all ranks, in rank order, into comm.outAll.  Runs after pipeline_end's loop: every rank reaches it
// exactly once per frame whatever its local overflow retries were.
int c2d_host::comm_gather_results(c2d_gpu_ctx* ctx)
{
    NvtxRange nvtx("c2d gather results (NCCL)");
    CommState& cm = ctx->comm;
    cm.resultsGathered = false;
    cm.gatherBytes = 0;
    cm.gatherTimed = false;
    if (!cm.comm || cm.resultMode == C2D_RESULTS_LOCAL || ctx->shardWorld < 2)
        return 0;
    NcclApi& api = nccl_api();
    cudaStream_t s = ctx->stream;
    const uint32_t world = ctx->shardWorld, rank = ctx->shardRank;
    C2D_CUDA(ctx, cudaEventRecord(cm.evX[2], s));
    C2D_CUDA(ctx, cm.dPairCounts.reserve(world + 1));
    cm.hPairCounts[0] = ctx->lastPairs;
    C2D_CUDA(ctx, cudaMemcpyAsync(cm.dPairCounts.ptr, cm.hPairCounts, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    C2D_NCCL(ctx, api.AllGather(cm.dPairCounts.ptr, cm.dPairCounts.ptr + 1, sizeof(uint64_t), ncclChar, comm_of(ctx), s));
    C2D_CUDA(ctx, cudaMemcpyAsync(cm.hPairCounts + 1, cm.dPairCounts.ptr + 1, world * sizeof(uint64_t),
                                  cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    uint64_t total = 0;
    for (uint32_t r = 0; r < world; ++r)
        total += cm.hPairCounts[1 + r];
    const bool receiver = cm.resultMode == C2D_RESULTS_ALL || rank == 0;
    const PairOut* local = ctx->resultsOrdered ? ctx->order.sorted.ptr : ctx->out.ptr; // this rank's records
    if (receiver)
        C2D_CUDA(ctx, cm.outAll.reserve(total ? total : 1));
    C2D_NCCL(ctx, api.GroupStart());
    uint64_t offset = 0;
    for (uint32_t r = 0; r < world; ++r)
    {
        const uint64_t n = cm.hPairCounts[1 + r];
        const size_t bytes = static_cast<size_t>(n) * sizeof(PairOut);
        if (n)
        {
            if (cm.resultMode == C2D_RESULTS_ALL)
                C2D_NCCL(ctx, api.Broadcast(local, cm.outAll.ptr + offset, bytes, ncclChar, static_cast<int>(r),
                                            comm_of(ctx), s));
            else if (rank == 0 && r != 0)
                C2D_NCCL(ctx, api.Recv(cm.outAll.ptr + offset, bytes, ncclChar, static_cast<int>(r), comm_of(ctx), s));
            else if (rank == r && r != 0)
                C2D_NCCL(ctx, api.Send(local, bytes, ncclChar, 0, comm_of(ctx), s));
        }
        offset += n;
    }
    C2D_NCCL(ctx, api.GroupEnd());
    if (cm.resultMode == C2D_RESULTS_ROOT && rank == 0 && cm.hPairCounts[1])
        C2D_CUDA(ctx, cudaMemcpyAsync(cm.outAll.ptr, local, static_cast<size_t>(cm.hPairCounts[1]) * sizeof(PairOut),
                                      cudaMemcpyDeviceToDevice, s));
    C2D_CUDA(ctx, cudaEventRecord(cm.evX[3], s));
    cm.gatherTimed = true;
    if (receiver)
    {
        cm.resultsGathered = true;
        cm.gatheredPairs = total;
        cm.gatherBytes = (total - cm.hPairCounts[1 + rank]) * sizeof(PairOut);
    }
    return 0;
}

void c2d_host::comm_release(c2d_gpu_ctx* ctx)
{
    CommState& cm = ctx->comm;
    if (cm.comm)
    {
        nccl_api().CommDestroy(comm_of(ctx));
        cm.comm = nullptr;
    }
    if (cm.hHeader)
        cudaFreeHost(cm.hHeader);
    if (cm.hPairCounts)
        cudaFreeHost(cm.hPairCounts);
    cm.hHeader = nullptr;
    cm.hPairCounts = nullptr;
    for (auto& ev : cm.evX)
    {
        if (ev)
            cudaEventDestroy(ev);
        ev = nullptr;
    }
}

extern "C" {

int c2d_gpu_comm_unique_id(void* out_id128)
{
    NcclApi& api = nccl_api();
    if (!api.handle)
        return fail(nullptr, "c2d_gpu_comm_unique_id: " + api.error);
    if (!out_id128)
        return fail(nullptr, "c2d_gpu_comm_unique_id: out_id128 is NULL");
    ncclUniqueId id;
    C2D_NCCL(nullptr, api.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out_id128, &id, sizeof(id));
    return 0;
}

int c2d_gpu_comm_init(c2d_gpu_ctx* ctx, uint32_t rank, uint32_t world, const void* id128)
{
    NcclApi& api = nccl_api();
    if (!api.handle)
        return fail(ctx, "c2d_gpu_comm_init: " + api.error);
    if (!id128 || world == 0 || rank >= world || world > 32)
        return fail(ctx, "c2d_gpu_comm_init: bad argument (rank < world <= 32, a 128-byte id)");
    if (ctx->comm.comm)
        return fail(ctx, "c2d_gpu_comm_init: this context already has a communicator");
    C2D_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    C2D_NCCL(ctx, api.CommInitRank(&comm, static_cast<int>(world), id, static_cast<int>(rank)));
    CommState& cm = ctx->comm;
    cm.comm = comm;
    for (auto& ev : cm.evX)
        C2D_CUDA(ctx, cudaEventCreate(&ev));
    C2D_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&cm.hHeader), ((world + 1) * 4 + 1) * sizeof(uint32_t), cudaHostAllocDefault));
    C2D_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&cm.hPairCounts), (world + 1) * sizeof(uint64_t), cudaHostAllocDefault));
    return c2d_gpu_set_shard(ctx, rank, world);
}

int c2d_gpu_comm_destroy(c2d_gpu_ctx* ctx)
{
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    comm_release(ctx);
    return 0;
}

int c2d_gpu_comm_set_move_exchange(c2d_gpu_ctx* ctx, int on)
{
    if (on && !ctx->comm.comm)
        return fail(ctx, "c2d_gpu_comm_set_move_exchange: no communicator (c2d_gpu_comm_init)");
    if (ctx->transLog->size)
        return fail(ctx, "c2d_gpu_comm_set_move_exchange: translations are pending; switch right after a query");
    ctx->comm.exchangeMoves = on != 0;
    ctx->comm.stride = 0; // the first exchange after a switch uses the exact protocol and learns the counts
    return 0;
}

int c2d_gpu_comm_set_result_mode(c2d_gpu_ctx* ctx, int mode)
{
    if (mode != C2D_RESULTS_LOCAL && mode != C2D_RESULTS_ALL && mode != C2D_RESULTS_ROOT)
        return fail(ctx, "c2d_gpu_comm_set_result_mode: unknown mode");
    if (mode != C2D_RESULTS_LOCAL && !ctx->comm.comm)
        return fail(ctx, "c2d_gpu_comm_set_result_mode: no communicator (c2d_gpu_comm_init)");
    ctx->comm.resultMode = mode;
    return 0;
}

int c2d_gpu_comm_get_stats(c2d_gpu_ctx* ctx, c2d_gpu_comm_stats* out)
{
    if (!out)
        return fail(ctx, "c2d_gpu_comm_get_stats: out is NULL");
    std::memset(out, 0, sizeof(*out));
    const CommState& cm = ctx->comm;
    out->exchange_bytes = cm.exchangeBytes;
    out->gather_bytes = cm.gatherBytes;
    out->local_pairs = ctx->lastPairs;
    out->slab_owned = ctx->slabInUse ? ctx->slab.owned : 0;
    // owned leaves + the halo shapes that queried them (the halo count arrives with the frame's asynchronous copies)
    out->slab_leaves = ctx->slabInUse ? ctx->slab.owned + (ctx->slab.hCounters ? ctx->slab.hCounters[0] : 0) : 0;
    cudaSetDevice(ctx->device);
    if (cm.exchangeTimed)
        cudaEventElapsedTime(&out->ms_exchange, cm.evX[0], cm.evX[1]);
    if (cm.gatherTimed && cudaEventSynchronize(cm.evX[3]) == cudaSuccess)
        cudaEventElapsedTime(&out->ms_gather, cm.evX[2], cm.evX[3]);
    if (ctx->slabInUse && ctx->slab.timed)
        cudaEventElapsedTime(&out->ms_slab, ctx->slab.ev[0], ctx->slab.ev[1]);
    return 0;
}

} // extern "C"
