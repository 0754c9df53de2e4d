// c2d_gpu.cu — context, mutation log, per-frame pipeline and the C ABI of libc2d_gpu.so (include/c2d_gpu.h).
//
// Host-side structure (one context per c2d::Registry):
//   * shape state lives in HBM as SoA (ShapeArrays); the host keeps only what it needs to schedule work:
//     body type / alive bit per shape, the member list of each of the three trees (Static, Dynamic, Bullet —
//     the reference's mTrees[3], Registry.hpp:157-165) and the hazard stamps of the mutation log;
//   * mutations are appended to pinned logs and replayed by the K1 kernels in "rounds": inside a round the
//     phases run flags -> adds -> translates -> moves and a shape appears at most once per phase with
//     strictly increasing phase, so per-shape program order is preserved without one launch per call;
//   * a query = flush + (re)build of the dirty trees + 5 traversals + binning + narrowphase on one stream.
#include "c2d_ctx.hpp"
#include "c2d_kernels.cuh"
#include "c2d_slab.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#if defined(__linux__)
#include <sys/mman.h>
#endif

using namespace c2d_dev;

namespace
{

thread_local std::string g_lastError;

} // namespace

namespace c2d_host
{
int fail(c2d_gpu_ctx* ctx, const std::string& msg)
{
    if (ctx)
        ctx->error = msg;
    g_lastError = msg;
    return 1;
}
} // namespace c2d_host

namespace
{

// Optional per-kernel CUDA-event timing (c2d_gpu_set_profiling): one event pair around each launch, read back
// after the frame's final synchronize.  Events on the launching stream add no synchronisation.
struct ProfScope
{
    c2d_gpu_ctx* ctx;
    cudaStream_t stream;
    cudaEvent_t stop = nullptr;
    ProfScope(c2d_gpu_ctx* c, const char* name, cudaStream_t st = nullptr) : ctx(c), stream(st ? st : c->stream)
    {
        if (!ctx->profiling)
            return;
        if (!ctx->profileOnly.empty() && ctx->profileOnly != name)
            return;
        cudaEvent_t a = nullptr;
        if (ctx->eventPool.size() >= 2)
        {
            a = ctx->eventPool.back();
            ctx->eventPool.pop_back();
            stop = ctx->eventPool.back();
            ctx->eventPool.pop_back();
        }
        else
        {
            cudaEventCreate(&a);
            cudaEventCreate(&stop);
        }
        cudaEventRecord(a, stream);
        ctx->spans.push_back(ProfSpan{name, a, stop});
    }
    ~ProfScope()
    {
        if (stop)
            cudaEventRecord(stop, stream);
    }
};
#define C2D_TIMED(ctx, name, ...)                                                                                      \
    do                                                                                                                 \
    {                                                                                                                  \
        ProfScope _ps(ctx, name);                                                                                      \
        __VA_ARGS__;                                                                                                   \
    } while (0)
#define C2D_TIMED_ON(ctx, stream, name, ...)                                                                           \
    do                                                                                                                 \
    {                                                                                                                  \
        ProfScope _ps(ctx, name, stream);                                                                              \
        __VA_ARGS__;                                                                                                   \
    } while (0)

// call after a stream synchronize
void collect_profile(c2d_gpu_ctx* ctx)
{
    for (const ProfSpan& sp : ctx->spans)
    {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, sp.start, sp.stop) == cudaSuccess)
        {
            auto& slot = ctx->kernelTotals[sp.name];
            slot.first += ms;
            slot.second += 1;
        }
        ctx->eventPool.push_back(sp.start);
        ctx->eventPool.push_back(sp.stop);
    }
    ctx->spans.clear();
}

template <typename T>
cudaError_t grow_array(T*& ptr, uint32_t oldCap, uint32_t newCap, size_t perShape, cudaStream_t stream)
{
    T* fresh = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&fresh), static_cast<size_t>(newCap) * perShape * sizeof(T));
    if (e != cudaSuccess)
        return e;
    e = cudaMemsetAsync(fresh, 0, static_cast<size_t>(newCap) * perShape * sizeof(T), stream);
    if (e != cudaSuccess)
        return e;
    if (ptr && oldCap)
    {
        e = cudaMemcpyAsync(fresh, ptr, static_cast<size_t>(oldCap) * perShape * sizeof(T), cudaMemcpyDeviceToDevice,
                            stream);
        if (e != cudaSuccess)
            return e;
    }
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess)
        return e;
    if (ptr)
        cudaFree(ptr);
    ptr = fresh;
    return cudaSuccess;
}

int ensure_shape_capacity(c2d_gpu_ctx* ctx, uint32_t want)
{
    if (want <= ctx->capacity)
        return 0;
    uint32_t cap = ctx->capacity ? ctx->capacity : 4096;
    while (cap < want)
        cap *= 2;
    const uint32_t old = ctx->capacity;
    cudaStream_t s = ctx->stream;
    C2D_CUDA(ctx, grow_array(ctx->d.xf, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.xa, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.pxf, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.pxa, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.tight, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.fat, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.meta, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.entity, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.category, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.mask, old, cap, 1, s));
    C2D_CUDA(ctx, grow_array(ctx->d.geom, old, cap, 32, s));
    ctx->capacity = cap;
    return 0;
}

// ---- mutation log ----------------------------------------------------------------------------------
void begin_round(c2d_gpu_ctx* ctx)
{
    ++ctx->round;
    ctx->rounds.push_back(Round{ctx->flagLog->size, ctx->addLog->size, ctx->transLog->size, ctx->moveLog->size, -1});
}

// Makes sure an op of `phase` on `shape` can join the current round (strictly increasing phase per shape).
inline void claim(c2d_gpu_ctx* ctx, uint32_t shape, Phase phase)
{
    if (shape >= ctx->stamp.size())
    {
        ctx->stamp.resize(std::max<size_t>(shape + 1, ctx->stamp.size() * 2), 0);
        ctx->hostMeta.resize(ctx->stamp.size(), 0);
        ctx->memberPos.resize(ctx->stamp.size(), 0);
    }
    if (ctx->rounds.empty())
        begin_round(ctx);
    const uint32_t st = ctx->stamp[shape];
    if ((st >> 2) == ctx->round && (st & 3u) >= static_cast<uint32_t>(phase))
        begin_round(ctx);
    ctx->stamp[shape] = (ctx->round << 2) | static_cast<uint32_t>(phase);
}

// the boxes or the membership of a body's tree changed (the slab tree of a multi-GPU group shadows the Dynamic tree)
inline void mark_dirty(c2d_gpu_ctx* ctx, uint32_t body)
{
    ctx->treeDirty[body] = true;
    if (body == C2D_DYNAMIC)
        ctx->slabDirty = true;
}

void member_add(c2d_gpu_ctx* ctx, uint32_t shape, uint32_t body)
{
    ctx->memberPos[shape] = static_cast<uint32_t>(ctx->members[body].size());
    ctx->members[body].push_back(shape);
    ctx->membersDirty[body] = true;
    if (body == C2D_DYNAMIC)
        ctx->slab.membershipChanged = true;
    mark_dirty(ctx, body);
}

void member_remove(c2d_gpu_ctx* ctx, uint32_t shape, uint32_t body)
{
    auto& list = ctx->members[body];
    const uint32_t pos = ctx->memberPos[shape];
    const uint32_t last = list.back();
    list[pos] = last;
    ctx->memberPos[last] = pos;
    list.pop_back();
    ctx->membersDirty[body] = true;
    if (body == C2D_DYNAMIC)
        ctx->slab.membershipChanged = true;
    mark_dirty(ctx, body);
}

} // namespace

namespace
{
// block size of the next one-step exchange: the largest count of this frame + 25 %, in steps of 4096 records
inline uint32_t next_exchange_stride(uint32_t largestCount)
{
    const uint64_t want = static_cast<uint64_t>(largestCount) + largestCount / 4 + 1;
    return static_cast<uint32_t>(std::min<uint64_t>((want + 4095) / 4096 * 4096, 0xfffff000ull));
}

// Move exchange of a multi-GPU group (c2d_gpu_comm_set_move_exchange): the frame's translation records of every rank are
// all-gathered device to device and applied on every rank.  Collective: called by every rank's c2d_gpu_collide* flush,
// also by ranks that logged nothing.  `local` = this rank's records in the device log, bodies = the trees they touch.
// Two protocols.  Exact (first exchange, or after an overflow): all-gather of the headers, a host round trip for the counts,
// then one broadcast per rank of exactly its records.  One-step (every later frame): the block size `stride` was agreed
// from the previous frame's counts (+25 %), so headers and padded record blocks go out in ONE grouped NCCL launch with no
// host round trip; the gathered headers come back to the host with the frame's other counters and size the next frame.
int exchange_translations(c2d_gpu_ctx* ctx, uint32_t count, uint32_t bodies, bool exact)
{
    CommState& cm = ctx->comm;
    cudaStream_t s = ctx->stream;
    const uint32_t world = ctx->shardWorld;
    const size_t headerWords = 4 * (static_cast<size_t>(world) + 1) + 1;
    C2D_CUDA(ctx, cudaEventRecord(cm.evX[0], s));
    C2D_CUDA(ctx, cm.dHeader.reserve(headerWords));
    cm.hHeader[0] = count;
    cm.hHeader[1] = bodies;
    cm.hHeader[2] = cm.hHeader[3] = 0;
    C2D_CUDA(ctx, cudaMemcpyAsync(cm.dHeader.ptr, cm.hHeader, 16, cudaMemcpyHostToDevice, s));
    uint32_t* dOverflow = cm.dHeader.ptr + 4 * (world + 1);
    cm.fastExchanged = false;
    if (!exact && cm.stride > 0)
    {
        const uint32_t stride = cm.stride;
        C2D_CUDA(ctx, cudaMemsetAsync(dOverflow, 0, sizeof(uint32_t), s));
        C2D_CUDA(ctx, cm.gathered.reserve(static_cast<size_t>(stride) * world));
        if (int rc = comm_allgather2(ctx, cm.dHeader.ptr, cm.dHeader.ptr + 4, 16, ctx->dTransLog.ptr, cm.gathered.ptr,
                                     static_cast<size_t>(stride) * sizeof(TransRec)))
            return rc;
        const size_t threads = static_cast<size_t>(stride) * world;
        C2D_TIMED(ctx, "k_apply_translates", k_apply_translates_gathered<<<blocks_for(threads, 256), 256, 0, s>>>(
                      cm.gathered.ptr, stride, cm.dHeader.ptr + 4, world, ctx->d, dOverflow));
        ++ctx->launches;
        // headers + flag follow the frame home (read in pipeline_end after its synchronise)
        C2D_CUDA(ctx, cudaMemcpyAsync(cm.hHeader + 4, cm.dHeader.ptr + 4, (4 * world + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        cm.fastExchanged = true;
        cm.exchangeBytes = static_cast<uint64_t>(stride) * (world - 1) * sizeof(TransRec);
        mark_dirty(ctx, C2D_DYNAMIC); // which bodies the other ranks moved is not known yet: Static is excluded by contract
        mark_dirty(ctx, C2D_BULLET);
        C2D_CUDA(ctx, cudaEventRecord(cm.evX[1], s));
        cm.exchangeTimed = true;
        return 0;
    }
    if (int rc = comm_allgather(ctx, cm.dHeader.ptr, cm.dHeader.ptr + 4, 16))
        return rc;
    C2D_CUDA(ctx, cudaMemcpyAsync(cm.hHeader + 4, cm.dHeader.ptr + 4, 16 * world, cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    uint32_t stride = 0, touched = 0;
    uint64_t total = 0;
    for (uint32_t r = 0; r < world; ++r)
    {
        stride = std::max(stride, cm.hHeader[4 + 4 * r]);
        total += cm.hHeader[4 + 4 * r];
        touched |= cm.hHeader[4 + 4 * r + 1];
    }
    cm.exchangeBytes = (total - count) * sizeof(TransRec);
    cm.stride = next_exchange_stride(stride);
    if (stride)
    {
        // this rank's records are in dTransLog already (flush: early chunks + tail); every rank broadcasts exactly its
        // count into block r of the gathered buffer
        C2D_CUDA(ctx, cm.gathered.reserve(static_cast<size_t>(stride) * world));
        size_t bytes[kSlabMaxWorld];
        for (uint32_t r = 0; r < world; ++r)
            bytes[r] = static_cast<size_t>(cm.hHeader[4 + 4 * r]) * sizeof(TransRec);
        if (int rc = comm_gatherv(ctx, ctx->dTransLog.ptr, cm.gathered.ptr, bytes, static_cast<size_t>(stride) * sizeof(TransRec)))
            return rc;
        const size_t threads = static_cast<size_t>(stride) * world;
        C2D_TIMED(ctx, "k_apply_translates", k_apply_translates_gathered<<<blocks_for(threads, 256), 256, 0, s>>>(
                      cm.gathered.ptr, stride, cm.dHeader.ptr + 4, world, ctx->d, nullptr));
        ++ctx->launches;
        for (uint32_t b = 0; b < 3; ++b)
            if (touched & (1u << b))
                mark_dirty(ctx, b);
    }
    C2D_CUDA(ctx, cudaEventRecord(cm.evX[1], s));
    cm.exchangeTimed = true;
    return 0;
}
} // namespace

namespace
{
constexpr size_t kEarlyUploadRecords = (1u << 20) / sizeof(TransRec); // ~1 MB of 12-byte records per chunk

// Records [transUploaded, size) of the translation log -> the same offsets of dTransLog, on the copy stream.  The copies
// overlap the caller's own moveEntity loop; flush() makes the main stream wait for them and uploads only the tail.
int early_upload_translations(c2d_gpu_ctx* ctx)
{
    PinnedVec<TransRec>& log = *ctx->transLog;
    if (ctx->dTransLog.capacity < log.capacity)
    {
        // growing the device log drops its contents: start over (warm-up only; both logs settle at the frame's size)
        C2D_CUDA(ctx, cudaStreamSynchronize(ctx->copyStream));
        C2D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        C2D_CUDA(ctx, ctx->dTransLog.reserve(log.capacity));
        ctx->transUploaded = 0;
    }
    if (ctx->transUploaded == 0)
        C2D_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->transApplied, 0)); // last frame's kernels still read dTransLog
    const size_t first = ctx->transUploaded, count = log.size - first;
    C2D_CUDA(ctx, cudaMemcpyAsync(ctx->dTransLog.ptr + first, log.data + first, count * sizeof(TransRec), cudaMemcpyHostToDevice,
                                  ctx->copyStream));
    ctx->transUploaded = log.size;
    return 0;
}
} // namespace

// collective: this flush is the one inside c2d_gpu_collide*, which every rank of a group calls in step
int c2d_host::flush(c2d_gpu_ctx* ctx, bool collective)
{
    NvtxRange nvtx("c2d flush (mutation log -> device, refit)");
    const bool exchange = ctx->comm.exchangeMoves && ctx->comm.comm && ctx->shardWorld > 1;
    if (exchange && !collective && ctx->transLog->size)
        return fail(ctx, "translations are pending while the move exchange is on: only c2d_gpu_collide* (the collective) may "
                         "flush them");
    ctx->comm.exchangeTimed = false;
    ctx->comm.exchangeBytes = 0;
    if (exchange && collective)
    {
        // the frame's translations: one round at most (an entity is translated once per frame), applied after everything else
        uint32_t roundsWithTranslations = 0;
        for (size_t r = 0; r < ctx->rounds.size(); ++r)
        {
            const size_t t1 = r + 1 == ctx->rounds.size() ? ctx->transLog->size : ctx->rounds[r + 1].t0;
            if (ctx->rounds[r].staged >= 0)
                return fail(ctx, "staged move batches cannot be combined with the move exchange");
            roundsWithTranslations += t1 > ctx->rounds[r].t0 ? 1 : 0;
        }
        if (roundsWithTranslations > 1)
            return fail(ctx, "move exchange: an entity may be translated only once per frame");
    }
    if (ctx->rounds.empty() && !(exchange && collective))
        return 0;
    cudaStream_t s = ctx->stream;
    // every id the host tables accept (c2d_gpu_shape_set_active on a stale slot, Registry.hpp:353-358) must be addressable
    if (int rc = ensure_shape_capacity(ctx, std::max<uint32_t>(ctx->shapeSlots, static_cast<uint32_t>(ctx->hostMeta.size()))))
        return rc;
    const size_t nf = ctx->flagLog->size, na = ctx->addLog->size, nm = ctx->moveLog->size;
    const size_t ntLogged = ctx->transLog->size;
    const size_t nt = exchange && collective ? 0 : ntLogged; // exchanged translations take their own path below
    if (nf)
    {
        C2D_CUDA(ctx, ctx->dFlagLog.reserve(nf));
        C2D_CUDA(ctx, cudaMemcpyAsync(ctx->dFlagLog.ptr, ctx->flagLog->data, nf * sizeof(FlagRec), cudaMemcpyHostToDevice, s));
    }
    if (na)
    {
        C2D_CUDA(ctx, ctx->dAddLog.reserve(na));
        C2D_CUDA(ctx, cudaMemcpyAsync(ctx->dAddLog.ptr, ctx->addLog->data, na * sizeof(AddRec), cudaMemcpyHostToDevice, s));
    }
    if (ntLogged || (exchange && collective))
    {
        // the chunks that went up during the caller's move loop, then the tail (in a group with the move exchange the
        // device log is this rank's send buffer)
        size_t done = std::min(ctx->transUploaded, ntLogged);
        // (the one-step move exchange sends `stride` records out of this buffer whatever the count)
        const size_t needed = std::max<size_t>(ntLogged, exchange && collective ? ctx->comm.stride : 0);
        if (ctx->dTransLog.capacity < needed)
        {
            C2D_CUDA(ctx, cudaStreamSynchronize(ctx->copyStream));
            C2D_CUDA(ctx, ctx->dTransLog.reserve(needed));
            done = 0;
        }
        if (done)
        {
            C2D_CUDA(ctx, cudaEventRecord(ctx->transUploadedEv, ctx->copyStream));
            C2D_CUDA(ctx, cudaStreamWaitEvent(s, ctx->transUploadedEv, 0));
        }
        if (ntLogged > done)
            C2D_CUDA(ctx, cudaMemcpyAsync(ctx->dTransLog.ptr + done, ctx->transLog->data + done, (ntLogged - done) * sizeof(TransRec),
                                          cudaMemcpyHostToDevice, s));
    }
    ctx->transUploaded = 0;
    if (nm)
    {
        C2D_CUDA(ctx, ctx->dMoveLog.reserve(nm));
        C2D_CUDA(ctx, cudaMemcpyAsync(ctx->dMoveLog.ptr, ctx->moveLog->data, nm * sizeof(MoveRec), cudaMemcpyHostToDevice, s));
    }
    ctx->h2dBytes += nf * sizeof(FlagRec) + na * sizeof(AddRec) + ntLogged * sizeof(TransRec) + nm * sizeof(MoveRec);

    for (size_t r = 0; r < ctx->rounds.size(); ++r)
    {
        const Round& cur = ctx->rounds[r];
        if (cur.staged >= 0)
        {
            const StagedBatch* sb = ctx->staged[cur.staged];
            if (sb->n)
            {
                C2D_TIMED(ctx, "k_apply_translates", k_apply_translates<<<blocks_for(sb->n, 256), 256, 0, s>>>(sb->records.ptr, sb->n, ctx->d));
                ++ctx->launches;
            }
            continue;
        }
        const bool lastRound = r + 1 == ctx->rounds.size();
        const size_t f1 = lastRound ? nf : ctx->rounds[r + 1].f0;
        const size_t a1 = lastRound ? na : ctx->rounds[r + 1].a0;
        const size_t t1 = nt == 0 ? cur.t0 : (lastRound ? nt : ctx->rounds[r + 1].t0);
        const size_t m1 = lastRound ? nm : ctx->rounds[r + 1].m0;
        if (f1 > cur.f0)
        {
            const uint32_t n = static_cast<uint32_t>(f1 - cur.f0);
            C2D_TIMED(ctx, "k_apply_flags", k_apply_flags<<<blocks_for(n, 256), 256, 0, s>>>(ctx->dFlagLog.ptr + cur.f0, n, ctx->d));
            ++ctx->launches;
        }
        if (a1 > cur.a0)
        {
            const uint32_t n = static_cast<uint32_t>(a1 - cur.a0);
            C2D_TIMED(ctx, "k_apply_adds", k_apply_adds<<<blocks_for(static_cast<size_t>(n) * 32, 128), 128, 0, s>>>(ctx->dAddLog.ptr + cur.a0, n, ctx->d));
            ++ctx->launches;
        }
        if (t1 > cur.t0)
        {
            const uint32_t n = static_cast<uint32_t>(t1 - cur.t0);
            C2D_TIMED(ctx, "k_apply_translates", k_apply_translates<<<blocks_for(n, 256), 256, 0, s>>>(ctx->dTransLog.ptr + cur.t0, n, ctx->d));
            ++ctx->launches;
        }
        if (m1 > cur.m0)
        {
            const uint32_t n = static_cast<uint32_t>(m1 - cur.m0);
            C2D_TIMED(ctx, "k_apply_moves", k_apply_moves<<<blocks_for(n, 256), 256, 0, s>>>(ctx->dMoveLog.ptr + cur.m0, n, ctx->d));
            ++ctx->launches;
        }
    }
    if (exchange && collective)
    {
        if (int rc = exchange_translations(ctx, static_cast<uint32_t>(ntLogged), ctx->transBodies, false))
            return rc;
    }
    ctx->transBodies = 0;
    C2D_CUDA(ctx, cudaGetLastError());
    C2D_CUDA(ctx, cudaEventRecord(ctx->transApplied, s));
    // The pinned logs are double buffered: the host keeps logging into the other set while these copies drain, and
    // only waits (normally not at all) if that set's previous upload has not completed yet.
    C2D_CUDA(ctx, cudaEventRecord(ctx->logConsumed[ctx->logIndex], s));
    ctx->logIndex ^= 1;
    ctx->flagLog = &ctx->flagLogs[ctx->logIndex];
    ctx->addLog = &ctx->addLogs[ctx->logIndex];
    ctx->transLog = &ctx->transLogs[ctx->logIndex];
    ctx->moveLog = &ctx->moveLogs[ctx->logIndex];
    C2D_CUDA(ctx, cudaEventSynchronize(ctx->logConsumed[ctx->logIndex]));
    ctx->flagLog->size = ctx->addLog->size = ctx->transLog->size = ctx->moveLog->size = 0;
    ctx->rounds.clear();
    return 0;
}

int c2d_host::upload_members(c2d_gpu_ctx* ctx, int body)
{
    if (!ctx->membersDirty[body])
        return 0;
    TreeBuf& t = ctx->tree[body];
    t.membershipChanged = true;
    const uint32_t n = static_cast<uint32_t>(ctx->members[body].size());
    if (n)
    {
        C2D_CUDA(ctx, t.members.reserve(n));
        C2D_CUDA(ctx, cudaMemcpyAsync(t.members.ptr, ctx->members[body].data(), n * sizeof(uint32_t),
                                      cudaMemcpyHostToDevice, ctx->stream));
        // the host vector may be modified right after this call returns
        C2D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->h2dBytes += n * sizeof(uint32_t);
    }
    ctx->membersDirty[body] = false;
    return 0;
}

namespace
{

// ---- tree build --------------------------------------------------------------------------------------
int build_tree(c2d_gpu_ctx* ctx, int body)
{
    TreeBuf& t = ctx->tree[body];
    cudaStream_t s = ctx->stream;
    const uint32_t n = static_cast<uint32_t>(ctx->members[body].size());
    if (int rc = upload_members(ctx, body))
        return rc;
    // Topology (Morton order + Karras links) is reused for up to rebuildPeriod - 1 frames when only boxes moved:
    // a bottom-up refit keeps every node box exact, so the candidate set is unchanged; only traversal efficiency
    // decays slowly as shapes drift away from their Morton slots (the reference's DynamicBVH never rebuilds at all).
    const bool topologyValid = t.n == n && n > 1 && !t.membershipChanged && t.age + 1 < ctx->rebuildPeriod;
    t.membershipChanged = false;
    t.n = n;
    ctx->treeDirty[body] = false;
    if (n == 0)
        return 0;
    if (topologyValid)
    {
        ++t.age;
        if (body == C2D_DYNAMIC)
        {
            C2D_CUDA(ctx, cudaEventRecord(ctx->evPreFit, s));
            ctx->preFitRecorded = true;
        }
        if (ctx->incrementalRefit)
            // leaf boxes change only through the mutation kernels, which flag them: grow the ancestors of the flagged leaves
            C2D_TIMED(ctx, "k_refit_dirty", k_refit_dirty<<<blocks_for(n, 256), 256, 0, s>>>(static_cast<int>(n), t.vals.ptr, ctx->d.fat, ctx->d.meta,
                                                             t.nodes.ptr, t.nodeParent.ptr, t.leafParent.ptr));
        else
        {
            C2D_CUDA(ctx, cudaMemsetAsync(t.arrival.ptr, 0, n * sizeof(uint32_t), s));
            C2D_TIMED(ctx, "k_fit", k_fit<<<blocks_for(n, 256), 256, 0, s>>>(static_cast<int>(n), t.vals.ptr, ctx->d.fat, ctx->d.category,
                                                     t.nodes.ptr, t.nodeParent.ptr, t.leafParent.ptr, t.arrival.ptr));
        }
        ++ctx->launches;
        return 0;
    }
    t.age = 0;
    C2D_CUDA(ctx, t.keys.reserve(n));
    C2D_CUDA(ctx, t.vals.reserve(n));
    C2D_CUDA(ctx, t.keysTmp.reserve(n));
    C2D_CUDA(ctx, t.valsTmp.reserve(n));
    C2D_CUDA(ctx, t.nodes.reserve(n));
    C2D_CUDA(ctx, t.nodeParent.reserve(n));
    C2D_CUDA(ctx, t.leafParent.reserve(n));
    C2D_CUDA(ctx, t.arrival.reserve(n));
    C2D_CUDA(ctx, t.sortWs.reserve(sort_workspace_bytes(n, 4)));

    uint32_t* bounds = &ctx->dScratch->bounds[body][0];
    uint32_t rb = blocks_for(n, 256);
    if (rb > 148 * 8)
        rb = 148 * 8;
    C2D_TIMED(ctx, "k_bounds", k_bounds<<<rb, 256, 0, s>>>(t.members.ptr, n, ctx->d.fat, bounds));
    C2D_TIMED(ctx, "k_morton", k_morton<<<blocks_for(n, 256), 256, 0, s>>>(t.members.ptr, n, ctx->d.fat, bounds, t.keys.ptr, t.vals.ptr));
    ctx->launches += 2;
    {
        ProfScope ps(ctx, "radix_sort(1 memset + 5 kernels)");
        ctx->launches += radix_sort_pairs(s, t.keys.ptr, t.vals.ptr, t.keysTmp.ptr, t.valsTmp.ptr, n, 4, t.sortWs.ptr);
    }
    if (n > 1)
    {
        C2D_CUDA(ctx, cudaMemsetAsync(t.arrival.ptr, 0, n * sizeof(uint32_t), s));
        C2D_TIMED(ctx, "k_karras_build", k_karras_build<<<blocks_for(n - 1, 256), 256, 0, s>>>(t.keys.ptr, t.vals.ptr, static_cast<int>(n), t.nodes.ptr,
                                                             t.nodeParent.ptr, t.leafParent.ptr));
        if (body == C2D_DYNAMIC)
        {
            C2D_CUDA(ctx, cudaEventRecord(ctx->evPreFit, s));
            ctx->preFitRecorded = true;
        }
        C2D_TIMED(ctx, "k_fit", k_fit<<<blocks_for(n, 256), 256, 0, s>>>(static_cast<int>(n), t.vals.ptr, ctx->d.fat, ctx->d.category,
                                                 t.nodes.ptr, t.nodeParent.ptr, t.leafParent.ptr, t.arrival.ptr));
        ctx->launches += 2;
    }
    C2D_CUDA(ctx, cudaGetLastError());
    return 0;
}

// The Dynamic tree of this rank's Morton slab (multi-GPU, c2d_slab.cuh).  Rebalance (membership changed, or every
// rebuildPeriod-th frame): keys of ALL Dynamic shapes, one sort (replicated, identical on every rank), the rank's range of
// the order becomes its tree.  Every frame: refit of that tree, the cells its leaf boxes touch, the halo selection.
// Nothing here waits for the device.
int update_slab(c2d_gpu_ctx* ctx)
{
    NvtxRange nvtx("c2d slab (order / refit / halo)");
    SlabState& sl = ctx->slab;
    TreeBuf& t = ctx->slabTree;
    cudaStream_t s = ctx->stream;
    const uint32_t n = static_cast<uint32_t>(ctx->members[C2D_DYNAMIC].size());
    const uint32_t world = ctx->shardWorld, rank = ctx->shardRank;
    if (!sl.hCounters)
    {
        C2D_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&sl.hCounters), kSlabCounterWords * sizeof(uint32_t), cudaHostAllocDefault));
        std::memset(sl.hCounters, 0, kSlabCounterWords * sizeof(uint32_t));
        C2D_CUDA(ctx, cudaEventCreate(&sl.ev[0]));
        C2D_CUDA(ctx, cudaEventCreate(&sl.ev[1]));
        C2D_CUDA(ctx, sl.bounds.reserve(4));
        C2D_CUDA(ctx, sl.cellBits.reserve(kSlabMaskWords));
        C2D_CUDA(ctx, sl.counters.reserve(kSlabCounterWords));
    }
    C2D_CUDA(ctx, cudaEventRecord(sl.ev[0], s));
    const bool rebalance = !sl.valid || sl.membershipChanged || sl.n != n || sl.age + 1 >= ctx->rebuildPeriod;
    if (rebalance)
    {
        if (int rc = upload_members(ctx, C2D_DYNAMIC))
            return rc;
        const uint32_t* members = ctx->tree[C2D_DYNAMIC].members.ptr;
        C2D_CUDA(ctx, sl.allKeys.reserve(n));
        C2D_CUDA(ctx, sl.allVals.reserve(n));
        C2D_CUDA(ctx, sl.keysTmp.reserve(n));
        C2D_CUDA(ctx, sl.valsTmp.reserve(n));
        C2D_CUDA(ctx, sl.halo.reserve(n));
        C2D_CUDA(ctx, sl.sortWs.reserve(sort_workspace_bytes(n, 4)));
        const uint32_t grid = std::min<uint32_t>(blocks_for(n, 256), 148 * 8);
        k_slab_reset_bounds<<<1, 32, 0, s>>>(sl.bounds.ptr);
        C2D_TIMED(ctx, "k_bounds", k_bounds<<<grid, 256, 0, s>>>(members, n, ctx->d.fat, sl.bounds.ptr));
        C2D_TIMED(ctx, "k_morton", k_morton<<<blocks_for(n, 256), 256, 0, s>>>(members, n, ctx->d.fat, sl.bounds.ptr, sl.allKeys.ptr, sl.allVals.ptr));
        ctx->launches += 3;
        {
            ProfScope ps(ctx, "radix_sort(1 memset + 5 kernels)");
            ctx->launches += radix_sort_pairs(s, sl.allKeys.ptr, sl.allVals.ptr, sl.keysTmp.ptr, sl.valsTmp.ptr, n, 4, sl.sortWs.ptr);
        }
        sl.lo = static_cast<uint32_t>(static_cast<uint64_t>(n) * rank / world);
        sl.hi = static_cast<uint32_t>(static_cast<uint64_t>(n) * (rank + 1) / world);
        sl.owned = sl.hi - sl.lo;
        sl.laterCount = n - sl.hi;
        sl.n = n;
        sl.age = 0;
        sl.valid = true;
        sl.membershipChanged = false;
        const uint32_t m = sl.owned;
        t.n = m;
        if (m)
        {
            C2D_CUDA(ctx, t.keys.reserve(m));
            C2D_CUDA(ctx, t.vals.reserve(m));
            C2D_CUDA(ctx, t.nodes.reserve(m));
            C2D_CUDA(ctx, t.nodeParent.reserve(m));
            C2D_CUDA(ctx, t.leafParent.reserve(m));
            C2D_CUDA(ctx, t.arrival.reserve(m));
            C2D_CUDA(ctx, cudaMemcpyAsync(t.keys.ptr, sl.allKeys.ptr + sl.lo, m * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
            C2D_CUDA(ctx, cudaMemcpyAsync(t.vals.ptr, sl.allVals.ptr + sl.lo, m * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        }
        if (m > 1)
        {
            C2D_TIMED(ctx, "k_karras_build", k_karras_build<<<blocks_for(m - 1, 256), 256, 0, s>>>(t.keys.ptr, t.vals.ptr, static_cast<int>(m), t.nodes.ptr,
                                                                 t.nodeParent.ptr, t.leafParent.ptr));
            ++ctx->launches;
        }
    }
    else
        ++sl.age;
    const uint32_t m = sl.owned;
    C2D_CUDA(ctx, cudaEventRecord(ctx->evPreFit, s));
    ctx->preFitRecorded = true;
    if (m > 1)
    {
        if (!rebalance && ctx->incrementalRefit)
            C2D_TIMED(ctx, "k_refit_dirty", k_refit_dirty<<<blocks_for(m, 256), 256, 0, s>>>(static_cast<int>(m), t.vals.ptr, ctx->d.fat, ctx->d.meta,
                                                             t.nodes.ptr, t.nodeParent.ptr, t.leafParent.ptr));
        else
        {
            C2D_CUDA(ctx, cudaMemsetAsync(t.arrival.ptr, 0, m * sizeof(uint32_t), s));
            C2D_TIMED(ctx, "k_fit", k_fit<<<blocks_for(m, 256), 256, 0, s>>>(static_cast<int>(m), t.vals.ptr, ctx->d.fat, ctx->d.category,
                                                     t.nodes.ptr, t.nodeParent.ptr, t.leafParent.ptr, t.arrival.ptr));
        }
        ++ctx->launches;
    }
    // the halo of this frame: exact for the CURRENT boxes
    C2D_CUDA(ctx, cudaMemsetAsync(sl.cellBits.ptr, 0, kSlabMaskWords * sizeof(uint32_t), s));
    C2D_CUDA(ctx, cudaMemsetAsync(sl.counters.ptr, 0, kSlabCounterWords * sizeof(uint32_t), s));
    if (m && sl.laterCount)
    {
        C2D_TIMED(ctx, "k_slab_mark", k_slab_mark<<<std::min<uint32_t>(blocks_for(m, 256), 148 * 2), 256, 0, s>>>(
                      t.vals.ptr, m, ctx->d.fat, sl.bounds.ptr, sl.cellBits.ptr, sl.counters.ptr));
        C2D_TIMED(ctx, "k_slab_select", k_slab_select<<<std::min<uint32_t>(blocks_for(sl.laterCount, 256), 148 * 4), 256, 0, s>>>(
                      sl.allVals.ptr + sl.hi, sl.laterCount, ctx->d.fat, sl.bounds.ptr, sl.cellBits.ptr, sl.halo.ptr, sl.counters.ptr));
        ctx->launches += 2;
    }
    C2D_CUDA(ctx, cudaEventRecord(sl.ev[1], s));
    sl.timed = true;
    // statistics only: read by c2d_gpu_comm_get_stats after the frame
    C2D_CUDA(ctx, cudaMemcpyAsync(sl.hCounters, sl.counters.ptr, kSlabCounterWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    ctx->slabDirty = false;
    C2D_CUDA(ctx, cudaGetLastError());
    return 0;
}

// Halo shapes of the slab query the rank's own Dynamic leaves: the Dynamic x Dynamic pairs that straddle two slabs.  The
// launch is sized by the number of later shapes; the kernel reads the real count on the device.
void launch_traverse_halo(c2d_gpu_ctx* ctx)
{
    const SlabState& sl = ctx->slab;
    const TreeBuf& t = ctx->slabTree;
    if (t.n < 2 || sl.laterCount == 0)
        return;
    // most launches have a few thousand halo shapes: a grid for laterCount would be mostly empty blocks, so the grid covers
    // a quarter of the later shapes per pass (every pass reads the count; passes beyond it exit at once)
    // most frames have a few thousand halo shapes: the first pass covers twice the halo of the LAST frame (known on the host
    // since that frame's counters came back), a second pass covers whatever could lie beyond it - its blocks read the real
    // count on the device and exit at once
    const uint32_t lastHalo = sl.hCounters ? sl.hCounters[kSlabHaloCount] : 0u;
    const uint32_t span = std::min(sl.laterCount, std::max<uint32_t>(4096u, 2u * lastHalo));
    for (uint32_t lo = 0; lo < sl.laterCount; lo = (lo == 0 ? span : sl.laterCount))
    {
        const uint32_t hi = lo == 0 ? span : sl.laterCount;
        C2D_TIMED(ctx, "k_traverse<halo>", (k_traverse_sync<false, true><<<blocks_for(hi - lo, 128), 128, 0, ctx->stream>>>(
                      sl.halo.ptr, sl.laterCount, lo, hi, sl.counters.ptr + kSlabHaloCount, t.n, t.nodes.ptr, ctx->d.fat, ctx->d.category,
                      0u, ctx->cand.ptr, static_cast<uint32_t>(ctx->cand.capacity), &ctx->dScratch->candCount, &ctx->dScratch->stackOverflow)));
        ++ctx->launches;
    }
}

} // namespace

int c2d_host::ensure_trees(c2d_gpu_ctx* ctx)
{
    if (int rc = flush(ctx))
        return rc;
    bool any = false;
    for (int body = 0; body < 3; ++body)
        any = any || ctx->treeDirty[body];
    if (!any)
        return 0;
    C2D_TIMED(ctx, "k_frame_init", k_frame_init<<<1, 32, 0, ctx->stream>>>(ctx->dScratch));
    ++ctx->launches;
    for (int body = 0; body < 3; ++body)
        if (ctx->treeDirty[body])
            if (int rc = build_tree(ctx, body))
                return rc;
    return 0;
}

namespace
{

// Traversal of the query leaves [lo, hi) of tree `q` against tree `t` on stream `st` (default: the context's stream).
// rejectLeaf: leaf links carrying any of these bits are skipped (no caller uses it at present).
template <bool SELF>
void launch_traverse_trees(c2d_gpu_ctx* ctx, const TreeBuf& q, const TreeBuf& t, uint32_t lo, uint32_t hi, uint32_t rejectLeaf,
                           cudaStream_t st = nullptr)
{
    if (q.n == 0 || t.n == 0 || (SELF && q.n < 2) || hi <= lo)
        return;
    if (!st)
        st = ctx->stream;
    uint32_t* overflow = &ctx->dScratch->stackOverflow;
    // few queries (bullets, whose swept leaves overlap hundreds of boxes): one WARP per query leaf
    if (static_cast<uint64_t>(hi - lo) * 32 <= 148ull * 2048 * 4)
        C2D_TIMED_ON(ctx, st, SELF ? "k_traverse_warp<self>" : "k_traverse_warp<cross>",
                  k_traverse_warp<SELF><<<blocks_for(static_cast<size_t>(hi - lo) * 32, 128), 128, 0, st>>>(
                      q.vals.ptr, q.n, lo, hi, t.vals.ptr, t.n, t.nodes.ptr, ctx->d.fat, ctx->d.category, rejectLeaf,
                      ctx->cand.ptr, static_cast<uint32_t>(ctx->cand.capacity), &ctx->dScratch->candCount, overflow));
    else if (ctx->traverseVariant == 0 || t.n < 2)
        C2D_TIMED_ON(ctx, st, SELF ? "k_traverse<self>" : "k_traverse<cross>", k_traverse<SELF><<<blocks_for(hi - lo, 128), 128, 0, st>>>(
        q.vals.ptr, q.n, lo, hi, t.vals.ptr, t.n, t.nodes.ptr, ctx->d.fat, ctx->d.category, ctx->cand.ptr,
        static_cast<uint32_t>(ctx->cand.capacity), &ctx->dScratch->candCount, overflow));
    else
        C2D_TIMED_ON(ctx, st, SELF ? "k_traverse<self>" : "k_traverse<cross>", k_traverse_sync<SELF><<<blocks_for(hi - lo, 128), 128, 0, st>>>(
        q.vals.ptr, q.n, lo, hi, nullptr, t.n, t.nodes.ptr, ctx->d.fat, ctx->d.category, rejectLeaf, ctx->cand.ptr,
        static_cast<uint32_t>(ctx->cand.capacity), &ctx->dScratch->candCount, overflow));
    ++ctx->launches;
}

template <bool SELF>
void launch_traverse(c2d_gpu_ctx* ctx, int qBody, int tBody, cudaStream_t st = nullptr)
{
    const TreeBuf& q = ctx->tree[qBody];
    const TreeBuf& t = ctx->tree[tBody];
    // multi-GPU: this context owns the query leaves of Morton-rank slab [lo, hi)
    const uint32_t lo = static_cast<uint32_t>(static_cast<uint64_t>(q.n) * ctx->shardRank / ctx->shardWorld);
    const uint32_t hi = static_cast<uint32_t>(static_cast<uint64_t>(q.n) * (ctx->shardRank + 1) / ctx->shardWorld);
    launch_traverse_trees<SELF>(ctx, q, t, lo, hi, 0u, st);
}

void fill_stats(c2d_gpu_ctx* ctx, c2d_gpu_stats* stats);

inline uint32_t env_u32(const char* name, uint32_t fallback)
{
    const char* v = std::getenv(name);
    const unsigned long n = v ? std::strtoul(v, nullptr, 10) : 0ul;
    return n >= 1 && n <= 64 ? static_cast<uint32_t>(n) : fallback;
}

// the records a query returns: this rank's, or the group's when the result mode gathered them (c2d_comm.cu)
inline const PairOut* result_ptr(const c2d_gpu_ctx* ctx)
{
    return ctx->comm.resultsGathered ? ctx->comm.outAll.ptr
                                     : (ctx->resultsOrdered ? ctx->order.sorted.ptr : (ctx->resultsOnHost ? ctx->smallOut.data : ctx->out.ptr));
}
// the records are already in host memory (a small scene's mapped output, neither reordered nor gathered)
inline bool result_on_host(const c2d_gpu_ctx* ctx)
{
    return ctx->resultsOnHost && !ctx->resultsOrdered && !ctx->comm.resultsGathered;
}
inline uint64_t result_count(const c2d_gpu_ctx* ctx)
{
    return ctx->comm.resultsGathered ? ctx->comm.gatheredPairs : ctx->lastPairs;
}

// One attempt at traversal + narrowphase with the current pair-list capacities; everything is enqueued on the
// stream, the frame counters follow in an asynchronous copy to the pinned FrameScratch.
int pipeline_attempt(c2d_gpu_ctx* ctx, bool retry = false)
{
    NvtxRange nvtx("c2d traverse + narrowphase (enqueue)");
    cudaStream_t s = ctx->stream;
    if (ctx->frame.smallScene)
    {
        const uint32_t nS = static_cast<uint32_t>(ctx->members[C2D_STATIC].size());
        const uint32_t nD = static_cast<uint32_t>(ctx->members[C2D_DYNAMIC].size());
        const uint32_t nB = static_cast<uint32_t>(ctx->members[C2D_BULLET].size());
        if (nD + nB)
            C2D_TIMED(ctx, "k_small_pairs", k_small_pairs<<<dim3(blocks_for(nD + nB, kSmallQueries), blocks_for(nS + nD + nB, kSmallTile)), kSmallQueries, 0, s>>>(
                ctx->tree[C2D_STATIC].members.ptr, nS, ctx->tree[C2D_DYNAMIC].members.ptr, nD, ctx->tree[C2D_BULLET].members.ptr, nB,
                ctx->d.fat, ctx->d.category, ctx->cand.ptr, static_cast<uint32_t>(ctx->cand.capacity), &ctx->dScratch->candCount));
        C2D_CUDA(ctx, cudaEventRecord(ctx->ev[3], s));
        if (ctx->smallOut.capacity == 0 && !ctx->smallOut.reserve(1u << 14))
            return fail(ctx, "pinned host allocation failed");
        C2D_TIMED(ctx, "k_small_narrow", k_small_narrow<<<148 * 2, 128, 0, s>>>(ctx->cand.ptr, static_cast<uint32_t>(ctx->cand.capacity), ctx->d,
                                                                             ctx->dScratch, ctx->smallOut.data,
                                                                             static_cast<uint32_t>(ctx->smallOut.capacity), ctx->hScratch));
        ctx->launches += 2;
        C2D_CUDA(ctx, cudaEventRecord(ctx->ev[4], s));
        ctx->frame.d2h += 16;
        return 0;
    }
    if (ctx->broadphase == C2D_BROADPHASE_GRID)
    {
        ProfScope ps(ctx, "grid_broadphase(extent+count+scan+fill+sort+gather+pairs)");
        if (int rc = grid_broadphase(ctx))
            return rc;
    }
    else
    {
        // The five groups of the reference (Registry.hpp:487-545).  Dynamic x Static needs the Static tree and the Dynamic
        // ORDER only, not the Dynamic tree's node boxes: it runs on the side stream next to that tree's bottom-up fit
        // (k_fit) and to the Dynamic x Dynamic traversal - all three are latency bound.
        cudaStream_t side = ctx->overlap ? ctx->sideStream : s;
        if (side != s)
        {
            if (ctx->preFitRecorded && !retry)
                C2D_CUDA(ctx, cudaStreamWaitEvent(side, ctx->evPreFit, 0));
            else
            {
                C2D_CUDA(ctx, cudaEventRecord(ctx->evFork, s));
                C2D_CUDA(ctx, cudaStreamWaitEvent(side, ctx->evFork, 0));
            }
        }
        if (ctx->slabInUse)
        {
            // multi-GPU slab: the rank's tree holds its owned Dynamic leaves; the halo shapes (higher ranks) query it;
            // every bullet meets the owned Dynamic leaves of each rank exactly once
            const TreeBuf& slab = ctx->slabTree;
            const TreeBuf& bullets = ctx->tree[C2D_BULLET];
            launch_traverse_trees<false>(ctx, slab, ctx->tree[C2D_STATIC], 0, slab.n, 0u, side);
            launch_traverse_trees<true>(ctx, slab, slab, 0, slab.n, 0u);
            launch_traverse_halo(ctx);
            launch_traverse<false>(ctx, C2D_BULLET, C2D_STATIC);
            launch_traverse_trees<false>(ctx, bullets, slab, 0, bullets.n, 0u);
            launch_traverse<true>(ctx, C2D_BULLET, C2D_BULLET);
        }
        else
        {
            launch_traverse<false>(ctx, C2D_DYNAMIC, C2D_STATIC, side);
            launch_traverse<true>(ctx, C2D_DYNAMIC, C2D_DYNAMIC);
            launch_traverse<false>(ctx, C2D_BULLET, C2D_STATIC);
            launch_traverse<false>(ctx, C2D_BULLET, C2D_DYNAMIC);
            launch_traverse<true>(ctx, C2D_BULLET, C2D_BULLET);
        }
        if (side != s)
        {
            C2D_CUDA(ctx, cudaEventRecord(ctx->evJoin, side));
            C2D_CUDA(ctx, cudaStreamWaitEvent(s, ctx->evJoin, 0));
        }
    }
    C2D_CUDA(ctx, cudaEventRecord(ctx->ev[3], s));

    const uint32_t cap = static_cast<uint32_t>(ctx->cand.capacity);
    const uint32_t grid = 148 * 4;
    C2D_TIMED(ctx, "k_bin_count", k_bin_count<<<grid, 256, 0, s>>>(ctx->cand.ptr, cap, ctx->d.meta, ctx->dScratch));
    C2D_TIMED(ctx, "k_bin_scatter", k_bin_scatter<<<grid, 256, 0, s>>>(ctx->cand.ptr, cap, ctx->d.meta, ctx->dScratch, ctx->binned.ptr));
    // the polygon bin (8 lanes per pair) and the other bins are disjoint slices of the binned list with their own hit
    // counters and output ranges: the two filter kernels, and later the two manifold kernels, run side by side
    cudaStream_t side = ctx->overlap ? ctx->sideStream : s;
    const auto fork = [&]() -> int {
        if (side != s)
        {
            C2D_CUDA(ctx, cudaEventRecord(ctx->evFork, s));
            C2D_CUDA(ctx, cudaStreamWaitEvent(side, ctx->evFork, 0));
        }
        return 0;
    };
    const auto join = [&]() -> int {
        if (side != s)
        {
            C2D_CUDA(ctx, cudaEventRecord(ctx->evJoin, side));
            C2D_CUDA(ctx, cudaStreamWaitEvent(s, ctx->evJoin, 0));
        }
        return 0;
    };
    if (int rc = fork())
        return rc;
    const bool bullets = !ctx->members[C2D_BULLET].empty();
    // blocks per SM of the four narrowphase kernels (grid-stride loops: any grid is correct); C2D_GRID_* override them
    static const uint32_t gFilter = 148 * env_u32("C2D_GRID_FILTER", 16), gFilterPP = 148 * env_u32("C2D_GRID_FILTER_PP", 16),
                          gManifold = 148 * env_u32("C2D_GRID_MANIFOLD", 16), gManifoldPP = 148 * env_u32("C2D_GRID_MANIFOLD_PP", 16);
    const uint32_t scanBlocks = gFilter + gFilterPP; // the last of them scans the hit counts
    if (bullets)
        C2D_TIMED_ON(ctx, side, "k_narrow_filter", k_narrow_filter<true><<<gFilter, 128, 0, side>>>(ctx->binned.ptr, cap, ctx->d, ctx->dScratch, ctx->hits.ptr, scanBlocks));
    else
        C2D_TIMED_ON(ctx, side, "k_narrow_filter", k_narrow_filter<false><<<gFilter, 128, 0, side>>>(ctx->binned.ptr, cap, ctx->d, ctx->dScratch, ctx->hits.ptr, scanBlocks));
    C2D_TIMED(ctx, "k_narrow_filter_pp", k_narrow_filter_pp<<<gFilterPP, 128, 0, s>>>(ctx->binned.ptr, ctx->d, ctx->dScratch, ctx->hits.ptr, scanBlocks));
    if (int rc = join())
        return rc;
    if (int rc = fork())
        return rc;
    if (bullets)
        C2D_TIMED_ON(ctx, side, "k_narrow_manifold", k_narrow_manifold<true><<<gManifold, 128, 0, side>>>(ctx->hits.ptr, ctx->d, ctx->dScratch, ctx->out.ptr,
                                         static_cast<uint32_t>(ctx->out.capacity)));
    else
        C2D_TIMED_ON(ctx, side, "k_narrow_manifold", k_narrow_manifold<false><<<gManifold, 128, 0, side>>>(ctx->hits.ptr, ctx->d, ctx->dScratch, ctx->out.ptr,
                                         static_cast<uint32_t>(ctx->out.capacity)));
    C2D_TIMED(ctx, "k_narrow_manifold_pp", k_narrow_manifold_pp<<<gManifoldPP, 128, 0, s>>>(ctx->hits.ptr, ctx->d, ctx->dScratch, ctx->out.ptr,
                                     static_cast<uint32_t>(ctx->out.capacity)));
    if (int rc = join())
        return rc;
    ctx->launches += 6;
    C2D_CUDA(ctx, cudaEventRecord(ctx->ev[4], s));
    C2D_CUDA(ctx, cudaMemcpyAsync(ctx->hScratch, ctx->dScratch, sizeof(FrameScratch), cudaMemcpyDeviceToHost, s));
    ctx->frame.d2h += sizeof(FrameScratch);
    return 0;
}

int pipeline_build(c2d_gpu_ctx* ctx);

// Replays the mutation log, refreshes the trees and enqueues the first attempt; returns without waiting for the GPU.
int pipeline_begin(c2d_gpu_ctx* ctx)
{
    if (ctx->frame.open)
        return fail(ctx, "c2d_gpu_collide_begin: the previous frame was not closed with c2d_gpu_collide_end");
    ctx->frame = c2d_gpu_ctx::FrameState{};
    ctx->frame.wall0 = std::chrono::steady_clock::now();
    cudaStream_t s = ctx->stream;
    ctx->launches = 0;
    ctx->h2dBytes = 0;
    ctx->preFitRecorded = false;

    C2D_CUDA(ctx, cudaEventRecord(ctx->ev[0], s));
    if (int rc = flush(ctx, true))
        return rc;
    C2D_CUDA(ctx, cudaEventRecord(ctx->ev[1], s));
    if (int rc = pipeline_build(ctx))
        return rc;
    ctx->frame.open = true;
    ctx->frame.msEnqueued = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - ctx->frame.wall0).count();
    return 0;
}

// Frame counters, the trees (or slab / grid / small-scene inputs) of the current boxes, then the first attempt.
int pipeline_build(c2d_gpu_ctx* ctx)
{
    NvtxRange nvtx("c2d build (trees / slab)");
    cudaStream_t s = ctx->stream;
    ctx->preFitRecorded = false;
    // small scenes: all-pairs box test + one narrowphase kernel instead of the tree pipeline (launch-latency bound);
    // like the grid broadphase this needs no tree, rays build them on demand (ensure_trees)
    const size_t alive = ctx->members[0].size() + ctx->members[1].size() + ctx->members[2].size();
    ctx->frame.smallScene = alive <= ctx->smallSceneMax && ctx->shardWorld == 1;
    if (!(ctx->frame.smallScene && ctx->scratchClean)) // a small frame leaves the counters zeroed for the next one
    {
        C2D_TIMED(ctx, "k_frame_init", k_frame_init<<<1, 32, 0, s>>>(ctx->dScratch));
        ++ctx->launches;
    }
    ctx->scratchClean = false;
    if (ctx->frame.smallScene)
    {
        for (int body = 0; body < 3; ++body)
            if (int rc = upload_members(ctx, body))
                return rc;
    }
    else if (ctx->broadphase != C2D_BROADPHASE_GRID)
    {
        ctx->slabInUse = ctx->shardWorld > 1 && ctx->slabMinShapes > 0 && ctx->members[C2D_DYNAMIC].size() >= ctx->slabMinShapes &&
                         ctx->members[C2D_DYNAMIC].size() >= 4ull * ctx->shardWorld;
        for (int body = 0; body < 3; ++body)
        {
            if (body == C2D_DYNAMIC && ctx->slabInUse)
            {
                // the full Dynamic tree stays stale (rays and per-entity queries rebuild it on demand, ensure_trees)
                if (ctx->slabDirty)
                    if (int rc = update_slab(ctx))
                        return rc;
            }
            else if (ctx->treeDirty[body])
                if (int rc = build_tree(ctx, body))
                    return rc;
        }
    }
    C2D_CUDA(ctx, cudaEventRecord(ctx->ev[2], s));

    if (ctx->cand.capacity == 0)
    {
        C2D_CUDA(ctx, ctx->cand.reserve(1u << 16));
        C2D_CUDA(ctx, ctx->binned.reserve(1u << 16));
        C2D_CUDA(ctx, ctx->hits.reserve(1u << 16));
    }
    if (ctx->out.capacity == 0)
        C2D_CUDA(ctx, ctx->out.reserve(1u << 15));
    return pipeline_attempt(ctx);
}

// The frame's records in the reference's order (k_order_* in c2d_kernels.cuh): out -> order.sorted.
int order_results(c2d_gpu_ctx* ctx, uint32_t n)
{
    ctx->resultsOrdered = false;
    if (!ctx->referenceOrder || n == 0)
        return 0;
    if (ctx->shapeSlots > 0x1fffffffu)
        return fail(ctx, "c2d_gpu_set_result_order: shape ids must stay below 2^29");
    auto& ob = ctx->order;
    const PairOut* produced = ctx->resultsOnHost ? ctx->smallOut.data : ctx->out.ptr; // mapped host memory reads fine on the device
    cudaStream_t s = ctx->stream;
    C2D_CUDA(ctx, ob.keys.reserve(n));
    C2D_CUDA(ctx, ob.vals.reserve(n));
    C2D_CUDA(ctx, ob.keysTmp.reserve(n));
    C2D_CUDA(ctx, ob.valsTmp.reserve(n));
    C2D_CUDA(ctx, ob.sortWs.reserve(sort_workspace_bytes(n, 4)));
    C2D_CUDA(ctx, ob.sorted.reserve(n));
    ProfScope ps(ctx, "result_order(2 sorts + gather)");
    k_order_key_b<<<blocks_for(n, 256), 256, 0, s>>>(produced, n, ob.keys.ptr, ob.vals.ptr);
    ctx->launches += 1 + radix_sort_pairs(s, ob.keys.ptr, ob.vals.ptr, ob.keysTmp.ptr, ob.valsTmp.ptr, n, 4, ob.sortWs.ptr);
    k_order_key_a<<<blocks_for(n, 256), 256, 0, s>>>(produced, n, ob.vals.ptr, ctx->d.meta, ob.keys.ptr);
    ctx->launches += 1 + radix_sort_pairs(s, ob.keys.ptr, ob.vals.ptr, ob.keysTmp.ptr, ob.valsTmp.ptr, n, 4, ob.sortWs.ptr);
    k_order_gather<<<blocks_for(static_cast<size_t>(n) * 32, 256), 256, 0, s>>>(produced, n, ob.vals.ptr, ob.sorted.ptr);
    ++ctx->launches;
    C2D_CUDA(ctx, cudaGetLastError());
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->resultsOrdered = true;
    return 0;
}

// Waits for the attempt in flight; on a pair-list overflow grows the lists and re-runs traversal + narrowphase
// (boxes and trees are unchanged).
int pipeline_end(c2d_gpu_ctx* ctx, c2d_gpu_stats* stats)
{
    NvtxRange nvtx("c2d wait for the frame");
    if (!ctx->frame.open)
        return fail(ctx, "c2d_gpu_collide_end without c2d_gpu_collide_begin");
    ctx->frame.open = false;
    cudaStream_t s = ctx->stream;
    uint32_t candCount = 0, pairCount = 0;
    while (true)
    {
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
        C2D_CUDA(ctx, cudaGetLastError());
        if (ctx->comm.fastExchanged)
        {
            // the one-step move exchange of this frame: its headers size the next frame's blocks; if a rank logged more
            // than its block held, nothing was applied (on any rank: all see the same headers) and the frame is redone
            // after an exchange with exact sizes
            CommState& cm = ctx->comm;
            cm.fastExchanged = false;
            const uint32_t world = ctx->shardWorld;
            uint32_t largest = 0;
            uint64_t total = 0;
            for (uint32_t r = 0; r < world; ++r)
            {
                largest = std::max(largest, cm.hHeader[4 + 4 * r]);
                total += cm.hHeader[4 + 4 * r];
            }
            const bool overflow = cm.hHeader[4 + 4 * world] != 0;
            cm.stride = next_exchange_stride(largest);
            if (overflow)
            {
                ++ctx->frame.retries;
                if (int rc = exchange_translations(ctx, cm.hHeader[0], cm.hHeader[1], true))
                    return rc;
                if (int rc = pipeline_build(ctx))
                    return rc;
                continue;
            }
            cm.exchangeBytes = (total - cm.hHeader[0]) * sizeof(TransRec);
        }
        candCount = ctx->hScratch->candCount;
        pairCount = ctx->hScratch->outCount;
        if (ctx->hScratch->clipOverflow)
            return fail(ctx, "polygon clip scratch overflow (non-convex polygon input?)");
        if (ctx->hScratch->stackOverflow)
            return fail(ctx, "LBVH traversal stack overflow (tree deeper than 64 levels)");
        const bool small = ctx->frame.smallScene;
        const bool candOverflow = candCount > ctx->cand.capacity;
        const bool outOverflow = pairCount > (small ? ctx->smallOut.capacity : ctx->out.capacity);
        if (!candOverflow && !outOverflow)
        {
            ctx->scratchClean = small;
            ctx->resultsOnHost = small;
            if (small)
                ctx->frame.d2h += static_cast<uint64_t>(pairCount) * sizeof(PairOut); // written across PCIe by the kernel
            break;
        }
        ++ctx->frame.retries;
        if (candOverflow)
        {
            C2D_CUDA(ctx, ctx->cand.reserve(static_cast<size_t>(candCount) + candCount / 4));
            C2D_CUDA(ctx, ctx->binned.reserve(ctx->cand.capacity));
            C2D_CUDA(ctx, ctx->hits.reserve(ctx->cand.capacity));
        }
        if (outOverflow && small)
        {
            ctx->smallOut.size = 0;
            if (!ctx->smallOut.reserve(static_cast<size_t>(pairCount) + pairCount / 4))
                return fail(ctx, "pinned host allocation failed");
        }
        else if (outOverflow)
            C2D_CUDA(ctx, ctx->out.reserve(static_cast<size_t>(pairCount) + pairCount / 4));
        C2D_TIMED(ctx, "k_frame_init", k_frame_init<<<1, 32, 0, s>>>(ctx->dScratch));
        ++ctx->launches;
        if (int rc = pipeline_attempt(ctx, true))
            return rc;
    }
    ctx->lastCandidates = candCount;
    ctx->lastPairs = pairCount;
    ctx->frame.msWaited = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - ctx->frame.wall0).count();
    if (int rc = order_results(ctx, pairCount))
        return rc;
    if (int rc = comm_gather_results(ctx)) // multi-GPU group with a result mode: collective, once per frame
        return rc;
    if (ctx->profiling)
        collect_profile(ctx);
    if (stats)
        fill_stats(ctx, stats);
    return 0;
}

void fill_stats(c2d_gpu_ctx* ctx, c2d_gpu_stats* stats)
{
    std::memset(stats, 0, sizeof(*stats));
    stats->shapes_alive = ctx->members[0].size() + ctx->members[1].size() + ctx->members[2].size();
    stats->candidates = ctx->lastCandidates;
    stats->pairs = result_count(ctx);
    stats->h2d_bytes = ctx->h2dBytes;
    stats->d2h_bytes = ctx->frame.d2h;
    stats->kernel_launches = ctx->launches;
    stats->retries = ctx->frame.retries;
    cudaEventElapsedTime(&stats->ms_refit, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&stats->ms_build, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&stats->ms_broad, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&stats->ms_narrow, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&stats->ms_device, ctx->ev[0], ctx->ev[4]);
    stats->ms_total =
        std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - ctx->frame.wall0).count();
    stats->ms_host_enqueue = ctx->frame.msEnqueued;
    stats->ms_host_wait = ctx->frame.msWaited;
}

// Records of the last frame -> the library's pinned buffer (one copy), for the pointer-returning entry points.
int download_records(c2d_gpu_ctx* ctx, c2d_gpu_stats* stats)
{
    const uint64_t pairCount = result_count(ctx);
    if (pairCount && !result_on_host(ctx))
    {
        if (!ctx->hostOut.reserve(pairCount))
            return fail(ctx, "pinned host allocation failed");
        C2D_CUDA(ctx, cudaMemcpyAsync(ctx->hostOut.data, result_ptr(ctx), static_cast<size_t>(pairCount) * sizeof(PairOut),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        C2D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->frame.d2h += pairCount * sizeof(PairOut);
    }
    if (stats)
        fill_stats(ctx, stats);
    return 0;
}

int run_pipeline(c2d_gpu_ctx* ctx, bool download, const c2d_pair_record** outRecords, uint64_t* outCount,
                 c2d_gpu_stats* stats)
{
    if (int rc = pipeline_begin(ctx))
        return rc;
    if (int rc = pipeline_end(ctx, download ? nullptr : stats))
        return rc;
    if (download)
        if (int rc = download_records(ctx, stats))
            return rc;
    if (outRecords)
        *outRecords = reinterpret_cast<const c2d_pair_record*>(result_on_host(ctx) ? ctx->smallOut.data : ctx->hostOut.data);
    if (outCount)
        *outCount = result_count(ctx);
    return 0;
}

} // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

// shared with c2d_batch.cu: the message returned by c2d_gpu_last_error(NULL)
void c2d_gpu_internal_set_error(const char* msg)
{
    g_lastError = msg ? msg : "";
}

int c2d_gpu_create(int device, c2d_gpu_ctx** out_ctx)
{
    if (!out_ctx)
        return fail(nullptr, "c2d_gpu_create: out_ctx is NULL");
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, std::string("c2d_gpu_create: no CUDA device (") + cudaGetErrorString(e) +
                                 "); libc2d_gpu has no CPU fallback");
    if (device < 0)
        cudaGetDevice(&device);
    C2D_CUDA(nullptr, cudaSetDevice(device));
    c2d_gpu_ctx* ctx = new c2d_gpu_ctx();
    ctx->device = device;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i)
        e = cudaEventCreate(&ctx->ev[i]);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i)
        e = cudaEventCreateWithFlags(&ctx->logConsumed[i], cudaEventDisableTiming);
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&ctx->sideStream, cudaStreamNonBlocking);
    for (cudaEvent_t* ev : {&ctx->evPreFit, &ctx->evFork, &ctx->evJoin})
        if (e == cudaSuccess)
            e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (e == cudaSuccess)
        e = cudaEventCreateWithFlags(&ctx->transApplied, cudaEventDisableTiming);
    if (e == cudaSuccess)
        e = cudaEventCreateWithFlags(&ctx->transUploadedEv, cudaEventDisableTiming);
    if (e == cudaSuccess)
        e = cudaMalloc(reinterpret_cast<void**>(&ctx->dScratch), sizeof(FrameScratch));
    if (e == cudaSuccess)
        e = cudaHostAlloc(reinterpret_cast<void**>(&ctx->hScratch), sizeof(FrameScratch), cudaHostAllocDefault);
    if (e != cudaSuccess)
    {
        fail(nullptr, std::string("c2d_gpu_create: ") + cudaGetErrorString(e));
        delete ctx;
        return 1;
    }
    if (const char* env = std::getenv("C2D_GPU_TRAVERSE")) // A/B switch for measurements (see c2d_ctx.hpp)
        ctx->traverseVariant = std::atoi(env);
    *out_ctx = ctx;
    return 0;
}

void c2d_gpu_destroy(c2d_gpu_ctx* ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d.xf);
    cudaFree(ctx->d.xa);
    cudaFree(ctx->d.pxf);
    cudaFree(ctx->d.pxa);
    cudaFree(ctx->d.tight);
    cudaFree(ctx->d.fat);
    cudaFree(ctx->d.meta);
    cudaFree(ctx->d.entity);
    cudaFree(ctx->d.category);
    cudaFree(ctx->d.mask);
    cudaFree(ctx->d.geom);
    cudaFree(ctx->dScratch);
    cudaFreeHost(ctx->hScratch);
    comm_release(ctx);
    if (ctx->slab.hCounters)
        cudaFreeHost(ctx->slab.hCounters);
    for (cudaEvent_t ev : ctx->slab.ev)
        if (ev)
            cudaEventDestroy(ev);
    for (StagedBatch* sb : ctx->staged)
        delete sb;
    for (auto& ev : ctx->ev)
        if (ev)
            cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->eventPool)
        cudaEventDestroy(ev);
    ctx->copyPool.reset();
    for (cudaEvent_t ev : ctx->copyEvents)
        cudaEventDestroy(ev);
    if (ctx->copyStream)
    {
        cudaStreamSynchronize(ctx->copyStream);
        cudaStreamDestroy(ctx->copyStream);
    }
    if (ctx->sideStream)
    {
        cudaStreamSynchronize(ctx->sideStream);
        cudaStreamDestroy(ctx->sideStream);
    }
    for (cudaEvent_t ev : {ctx->evPreFit, ctx->evFork, ctx->evJoin})
        if (ev)
            cudaEventDestroy(ev);
    if (ctx->transApplied)
        cudaEventDestroy(ctx->transApplied);
    if (ctx->transUploadedEv)
        cudaEventDestroy(ctx->transUploadedEv);
    for (cudaEvent_t ev : ctx->logConsumed)
        if (ev)
            cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* c2d_gpu_last_error(c2d_gpu_ctx* ctx)
{
    return ctx ? ctx->error.c_str() : g_lastError.c_str();
}

int c2d_gpu_clear(c2d_gpu_ctx* ctx)
{
    ++ctx->mutationEpoch;
    cudaSetDevice(ctx->device);
    C2D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    C2D_CUDA(ctx, cudaStreamSynchronize(ctx->copyStream));
    ctx->transUploaded = 0;
    ctx->flagLog->size = ctx->addLog->size = ctx->transLog->size = ctx->moveLog->size = 0;
    ctx->rounds.clear();
    for (StagedBatch* sb : ctx->staged)
        delete sb;
    ctx->staged.clear();
    std::fill(ctx->hostMeta.begin(), ctx->hostMeta.end(), 0u);
    std::fill(ctx->stamp.begin(), ctx->stamp.end(), 0u);
    ctx->shapeSlots = 0;
    for (int b = 0; b < 3; ++b)
    {
        ctx->members[b].clear();
        ctx->membersDirty[b] = true;
        mark_dirty(ctx, b);
        ctx->tree[b].n = 0;
    }
    ctx->slab.valid = false;
    if (ctx->capacity)
        C2D_CUDA(ctx, cudaMemsetAsync(ctx->d.meta, 0, ctx->capacity * sizeof(uint32_t), ctx->stream));
    return 0;
}

int c2d_gpu_set_broadphase(c2d_gpu_ctx* ctx, int mode, uint32_t cell_size)
{
    if (mode != C2D_BROADPHASE_LBVH && mode != C2D_BROADPHASE_GRID)
        return fail(ctx, "c2d_gpu_set_broadphase: unknown mode");
    ctx->broadphase = mode;
    ctx->cellSize = cell_size ? cell_size : 120;
    return 0;
}

int c2d_gpu_set_profiling(c2d_gpu_ctx* ctx, int on)
{
    ctx->profiling = on != 0;
    return 0;
}

int c2d_gpu_set_profile_filter(c2d_gpu_ctx* ctx, const char* kernel_name)
{
    ctx->profileOnly = kernel_name ? kernel_name : "";
    return 0;
}

int c2d_gpu_kernel_profile(c2d_gpu_ctx* ctx, char* buffer, uint64_t capacity, int reset)
{
    std::string text;
    for (const auto& kv : ctx->kernelTotals)
        text += kv.first + "\t" + std::to_string(kv.second.second) + "\t" + std::to_string(kv.second.first) + "\n";
    if (buffer && capacity)
    {
        const size_t n = std::min<size_t>(text.size(), capacity - 1);
        std::memcpy(buffer, text.data(), n);
        buffer[n] = 0;
    }
    if (reset)
        ctx->kernelTotals.clear();
    return 0;
}

int c2d_gpu_set_rebuild_period(c2d_gpu_ctx* ctx, uint32_t frames)
{
    ctx->rebuildPeriod = frames ? frames : 1;
    return 0;
}

int c2d_gpu_mutation_epoch(c2d_gpu_ctx* ctx, uint64_t* out_epoch)
{
    if (out_epoch)
        *out_epoch = ctx->mutationEpoch;
    return 0;
}

int c2d_gpu_set_small_scene_limit(c2d_gpu_ctx* ctx, uint32_t shapes)
{
    ctx->smallSceneMax = shapes;
    return 0;
}

int c2d_gpu_set_shard(c2d_gpu_ctx* ctx, uint32_t rank, uint32_t world)
{
    if (world == 0 || rank >= world)
        return fail(ctx, "c2d_gpu_set_shard: rank must be < world");
    if (world > static_cast<uint32_t>(kSlabMaxWorld))
        return fail(ctx, "c2d_gpu_set_shard: at most 32 ranks");
    ctx->shardRank = rank;
    ctx->shardWorld = world;
    ctx->slabDirty = true;
    ctx->slab.valid = false;
    return 0;
}

int c2d_gpu_set_result_order(c2d_gpu_ctx* ctx, int reference_order)
{
    ctx->referenceOrder = reference_order != 0;
    ctx->resultsOrdered = false;
    return 0;
}

int c2d_gpu_set_incremental_refit(c2d_gpu_ctx* ctx, int on)
{
    ctx->incrementalRefit = on != 0;
    return 0;
}

int c2d_gpu_set_overlap(c2d_gpu_ctx* ctx, int on)
{
    ctx->overlap = on != 0;
    return 0;
}

int c2d_gpu_set_ray_sharding(c2d_gpu_ctx* ctx, int on)
{
    ctx->raySharding = on != 0;
    return 0;
}

int c2d_gpu_set_slab_mode(c2d_gpu_ctx* ctx, uint32_t min_shapes)
{
    ctx->slabMinShapes = min_shapes;
    ctx->slabDirty = true;
    ctx->slab.valid = false;
    return 0;
}

int c2d_gpu_shape_add(c2d_gpu_ctx* ctx, uint32_t shape, uint32_t entity_tag, uint32_t type, uint32_t body,
                      uint32_t vertex_count, uint64_t category, uint64_t mask, const float* geom32, const c2d_xf* xf,
                      const c2d_xf* prev)
{
    ++ctx->mutationEpoch;
    if (type > 3 || body > 2 || vertex_count > 8 || !geom32 || !xf)
        return fail(ctx, "c2d_gpu_shape_add: bad argument");
    claim(ctx, shape, PH_ADD);
    if (ctx->hostMeta[shape] & META_ALIVE)
        return fail(ctx, "c2d_gpu_shape_add: shape id is already in use");
    AddRec* r = ctx->addLog->push();
    if (!r)
        return fail(ctx, "pinned host allocation failed");
    const uint32_t meta = type | (body << META_BODY_SHIFT) | META_ACTIVE | META_ALIVE | (vertex_count << META_COUNT_SHIFT);
    r->shape = shape;
    r->entity = entity_tag;
    r->meta = meta;
    r->_pad = 0;
    r->category = category;
    r->mask = mask;
    std::memcpy(r->xf, xf, sizeof(c2d_xf));
    std::memcpy(r->prev, prev ? prev : xf, sizeof(c2d_xf));
    std::memcpy(r->geom, geom32, sizeof(r->geom));
    ctx->hostMeta[shape] = meta;
    if (shape + 1 > ctx->shapeSlots)
        ctx->shapeSlots = shape + 1;
    member_add(ctx, shape, body);
    return 0;
}

int c2d_gpu_shape_remove(c2d_gpu_ctx* ctx, uint32_t shape)
{
    ++ctx->mutationEpoch;
    if (shape >= ctx->hostMeta.size() || !(ctx->hostMeta[shape] & META_ALIVE))
        return fail(ctx, "c2d_gpu_shape_remove: unknown shape");
    claim(ctx, shape, PH_FLAG);
    FlagRec* r = ctx->flagLog->push();
    if (!r)
        return fail(ctx, "pinned host allocation failed");
    r->shape = shape;
    r->op = 0;
    member_remove(ctx, shape, meta_body(ctx->hostMeta[shape]));
    ctx->hostMeta[shape] &= ~META_ALIVE;
    return 0;
}

int c2d_gpu_shape_set_active(c2d_gpu_ctx* ctx, uint32_t shape, int active)
{
    ++ctx->mutationEpoch;
    // like the reference (Registry.hpp:353-358) this also works on a removed shape's stale slot
    if (shape >= ctx->hostMeta.size())
        return fail(ctx, "c2d_gpu_shape_set_active: unknown shape");
    claim(ctx, shape, PH_FLAG);
    FlagRec* r = ctx->flagLog->push();
    if (!r)
        return fail(ctx, "pinned host allocation failed");
    r->shape = shape;
    r->op = active ? 1u : 2u;
    return 0;
}

int c2d_gpu_shape_set_ghost(c2d_gpu_ctx* ctx, uint32_t shape, int ghost)
{
    ++ctx->mutationEpoch;
    if (shape >= ctx->hostMeta.size() || !(ctx->hostMeta[shape] & META_ALIVE))
        return fail(ctx, "c2d_gpu_shape_set_ghost: unknown shape");
    claim(ctx, shape, PH_FLAG);
    FlagRec* r = ctx->flagLog->push();
    if (!r)
        return fail(ctx, "pinned host allocation failed");
    r->shape = shape;
    r->op = ghost ? 3u : 4u;
    return 0;
}

int c2d_gpu_shape_translate(c2d_gpu_ctx* ctx, uint32_t shape, float dx, float dy)
{
    ++ctx->mutationEpoch;
    if (shape >= ctx->hostMeta.size() || !(ctx->hostMeta[shape] & META_ALIVE))
        return fail(ctx, "c2d_gpu_shape_translate: unknown shape");
    claim(ctx, shape, PH_TRANS);
    if (ctx->transUploaded && ctx->transLog->size == ctx->transLog->capacity)
        cudaStreamSynchronize(ctx->copyStream); // the pinned log is about to move: its chunks in flight must land first
    TransRec* r = ctx->transLog->push();
    if (!r)
        return fail(ctx, "pinned host allocation failed");
    r->shape = shape;
    r->dx = dx;
    r->dy = dy;
    const uint32_t body = meta_body(ctx->hostMeta[shape]);
    if (body == C2D_STATIC && ctx->comm.exchangeMoves)
        return fail(ctx, "move exchange: translations of Static shapes are not exchanged (move them on every rank with the exchange off)");
    ctx->transBodies |= 1u << body;
    mark_dirty(ctx, body);
    if (ctx->earlyUpload && ctx->transLog->size - ctx->transUploaded >= kEarlyUploadRecords)
        return early_upload_translations(ctx);
    return 0;
}

int c2d_gpu_shapes_translate(c2d_gpu_ctx* ctx, uint32_t n, const uint32_t* shapes, const float* dxy)
{
    ++ctx->mutationEpoch;
    if (n == 0)
        return 0;
    if (!shapes || !dxy)
        return fail(ctx, "c2d_gpu_shapes_translate: NULL argument");
    if (ctx->transUploaded && ctx->transLog->size + n > ctx->transLog->capacity)
        cudaStreamSynchronize(ctx->copyStream);
    if (!ctx->transLog->reserve(ctx->transLog->size + n))
        return fail(ctx, "pinned host allocation failed");
    // the loop of c2d_gpu_shape_translate without its per-record function call, capacity check and upload check
    const size_t known = ctx->hostMeta.size();
    uint32_t bodies = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        const uint32_t shape = shapes[i];
        if (shape >= known || !(ctx->hostMeta[shape] & META_ALIVE))
            return fail(ctx, "c2d_gpu_shapes_translate: unknown shape");
        claim(ctx, shape, PH_TRANS);
        TransRec* r = &ctx->transLog->data[ctx->transLog->size++];
        r->shape = shape;
        r->dx = dxy[2 * i];
        r->dy = dxy[2 * i + 1];
        bodies |= 1u << meta_body(ctx->hostMeta[shape]);
        if ((bodies & (1u << C2D_STATIC)) && ctx->comm.exchangeMoves)
            return fail(ctx, "move exchange: translations of Static shapes are not exchanged (move them on every rank with the exchange off)");
        if ((i & 0xfffu) == 0xfffu && ctx->earlyUpload && ctx->transLog->size - ctx->transUploaded >= kEarlyUploadRecords)
            if (int rc = early_upload_translations(ctx))
                return rc;
    }
    ctx->transBodies |= bodies;
    for (uint32_t b = 0; b < 3; ++b)
        if (bodies & (1u << b))
            mark_dirty(ctx, b);
    return 0;
}

int c2d_gpu_shape_move(c2d_gpu_ctx* ctx, uint32_t shape, const c2d_xf* xf)
{
    ++ctx->mutationEpoch;
    if (shape >= ctx->hostMeta.size() || !(ctx->hostMeta[shape] & META_ALIVE) || !xf)
        return fail(ctx, "c2d_gpu_shape_move: unknown shape");
    claim(ctx, shape, PH_MOVE);
    MoveRec* r = ctx->moveLog->push();
    if (!r)
        return fail(ctx, "pinned host allocation failed");
    r->shape = shape;
    r->tx = xf->tx;
    r->ty = xf->ty;
    r->angle = xf->angle;
    r->sin = xf->sin;
    r->cos = xf->cos;
    r->csin = xf->csin;
    r->ccos = xf->ccos;
    mark_dirty(ctx, meta_body(ctx->hostMeta[shape]));
    return 0;
}

int c2d_gpu_stage_translations(c2d_gpu_ctx* ctx, uint32_t n, const uint32_t* shapes, const float* dxy,
                               uint32_t* out_handle)
{
    cudaSetDevice(ctx->device);
    if (!out_handle || (n && (!shapes || !dxy)))
        return fail(ctx, "c2d_gpu_stage_translations: bad argument");
    StagedBatch* sb = new StagedBatch();
    std::vector<TransRec> host(n);
    std::vector<uint8_t> seen(ctx->hostMeta.size(), 0);
    for (uint32_t i = 0; i < n; ++i)
    {
        const uint32_t s = shapes[i];
        if (s >= ctx->hostMeta.size() || !(ctx->hostMeta[s] & META_ALIVE) || seen[s])
        {
            delete sb;
            return fail(ctx, "c2d_gpu_stage_translations: unknown or repeated shape in batch");
        }
        seen[s] = 1;
        sb->touches[meta_body(ctx->hostMeta[s])] = true;
        host[i] = TransRec{s, dxy[2 * i], dxy[2 * i + 1]};
    }
    sb->n = n;
    sb->live = true;
    if (n)
    {
        if (sb->records.reserve(n) != cudaSuccess ||
            cudaMemcpy(sb->records.ptr, host.data(), n * sizeof(TransRec), cudaMemcpyHostToDevice) != cudaSuccess)
        {
            delete sb;
            return fail(ctx, "c2d_gpu_stage_translations: device allocation / copy failed");
        }
    }
    size_t slot = 0;
    while (slot < ctx->staged.size() && ctx->staged[slot])
        ++slot;
    if (slot == ctx->staged.size())
        ctx->staged.push_back(nullptr);
    ctx->staged[slot] = sb;
    *out_handle = static_cast<uint32_t>(slot);
    return 0;
}

int c2d_gpu_apply_staged(c2d_gpu_ctx* ctx, uint32_t handle)
{
    ++ctx->mutationEpoch;
    if (handle >= ctx->staged.size() || !ctx->staged[handle])
        return fail(ctx, "c2d_gpu_apply_staged: unknown handle");
    const StagedBatch* sb = ctx->staged[handle];
    // the batch is its own round (a barrier between everything logged before and after it), so no per-shape
    // hazard stamps are needed: O(1) host work per frame.  Shapes removed since staging must not be in it.
    begin_round(ctx);
    ctx->rounds.back().staged = static_cast<int>(handle);
    begin_round(ctx);
    for (int b = 0; b < 3; ++b)
        if (sb->touches[b])
            mark_dirty(ctx, b);
    return 0;
}

int c2d_gpu_release_staged(c2d_gpu_ctx* ctx, uint32_t handle)
{
    cudaSetDevice(ctx->device);
    if (handle >= ctx->staged.size() || !ctx->staged[handle])
        return fail(ctx, "c2d_gpu_release_staged: unknown handle");
    if (int rc = flush(ctx)) // a logged round may still reference the batch
        return rc;
    delete ctx->staged[handle];
    ctx->staged[handle] = nullptr;
    return 0;
}

int c2d_gpu_flush(c2d_gpu_ctx* ctx)
{
    cudaSetDevice(ctx->device);
    return flush(ctx);
}

int c2d_gpu_collide(c2d_gpu_ctx* ctx, const c2d_pair_record** out_records, uint64_t* out_count, c2d_gpu_stats* stats)
{
    cudaSetDevice(ctx->device);
    return run_pipeline(ctx, true, out_records, out_count, stats);
}

int c2d_gpu_collide_device(c2d_gpu_ctx* ctx, uint64_t* out_count, c2d_gpu_stats* stats)
{
    cudaSetDevice(ctx->device);
    return run_pipeline(ctx, false, nullptr, out_count, stats);
}

} // extern "C"

// ---- copy-out straight into caller storage -------------------------------------------------------------------
namespace
{
constexpr uint64_t kPoolThresholdBytes = 1ull << 20; // smaller results: one copy on the calling thread (a thread moves pinned ->
                                                     // pageable memory at 5-9 GB/s here: at 4 MB the pipelined, two-thread path is already 0.4 ms faster)
constexpr uint64_t kFetchWaveBytes = 64ull << 20; // 64 MB per wave; two staging halves alternate
constexpr uint64_t kFetchChunkBytes = 1ull << 20; // copy_host: pinned -> pageable split between the helper threads

// Destination buffers are usually freshly allocated (a new std::vector per call): ask for transparent huge pages so
// that filling a GB-sized result costs thousands of page faults instead of hundreds of thousands.  A hint only.
void advise_huge_pages(void* dst, uint64_t bytes)
{
#if defined(__linux__) && defined(MADV_HUGEPAGE)
    constexpr uintptr_t kHuge = 2u << 20;
    const uintptr_t lo = (reinterpret_cast<uintptr_t>(dst) + kHuge - 1) & ~(kHuge - 1);
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(dst) + bytes) & ~(kHuge - 1);
    if (hi > lo)
        madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
#else
    (void)dst, (void)bytes;
#endif
}

uint64_t env_kb(const char* name, uint64_t fallbackKb)
{
    const char* env = std::getenv(name);
    uint64_t kb = env ? std::strtoull(env, nullptr, 10) : fallbackKb;
    kb = kb < 16 ? 16 : (kb > 65536 ? 65536 : kb);
    return kb << 10;
}

c2d_host::CopyPool* copy_pool(c2d_gpu_ctx* ctx)
{
    if (!ctx->copyPool)
    {
        unsigned helpers = std::thread::hardware_concurrency() / 2;
        // one process per GPU (torchrun exports LOCAL_WORLD_SIZE): the ranks of a node share its cores
        if (const char* lws = std::getenv("LOCAL_WORLD_SIZE"))
        {
            const unsigned ranks = static_cast<unsigned>(std::strtoul(lws, nullptr, 10));
            if (ranks > 1) // all cores, split between the ranks (their main threads are blocked in this copy anyway)
                helpers = std::thread::hardware_concurrency() / ranks > 0 ? std::thread::hardware_concurrency() / ranks : 1;
        }
        helpers = helpers > 8 ? 8 : helpers;
        if (const char* env = std::getenv("C2D_GPU_COPY_THREADS"))
            helpers = static_cast<unsigned>(std::strtoul(env, nullptr, 10));
        helpers = helpers > 64 ? 64 : helpers;
        ctx->copyPool = std::make_unique<c2d_host::CopyPool>(helpers > 0 ? helpers - 1 : 0); // the caller is a worker too
    }
    return ctx->copyPool.get();
}
} // namespace

// Device memory -> caller (pageable) memory: device-to-host copies into pinned staging, moved on to `dst` by the helper
// threads (and this thread) as the data lands.  Measured on B200 / PCIe 5 (scripts/probe_copy.py, 35 MB): one copy into
// pinned memory takes 0.63 ms (55.7 GB/s); one thread moves pinned -> pageable memory at ~9 GB/s; a cudaMemcpyAsync +
// cudaEventRecord pair costs ~8 us of host time.  So the DMA goes in 4 MB chunks (few API calls, few bubbles on the
// link), each chunk is memcpy'd as eight 512 KB units by whichever workers are free, and the last two chunks are cut
// into 256 KB DMA pieces because the memcpy of the very last piece is the tail nobody can overlap: 0.77 ms
// (1 MB chunks = units: 0.85-0.93 ms; 8 MB: 0.83 ms; 256 KB: 1.2 ms, issue bound).
int c2d_host::copy_out(c2d_gpu_ctx* ctx, const void* srcDevice, void* dstHost, uint64_t total)
{
    NvtxRange nvtx("c2d copy-out (device -> caller memory)");
    if (!total)
        return 0;
    cudaStream_t s = ctx->stream;
    const char* src = static_cast<const char*>(srcDevice);
    char* out = static_cast<char*>(dstHost);
    if (total < kPoolThresholdBytes)
    {
        if (!ctx->copyStage.reserve(total))
            return fail(ctx, "pinned host allocation failed");
        C2D_CUDA(ctx, cudaMemcpyAsync(ctx->copyStage.data, src, total, cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
        std::memcpy(out, ctx->copyStage.data, total);
        return 0;
    }
    // DMA chunk (one cudaMemcpyAsync + one event), memcpy unit of the workers, DMA piece size inside the tail region
    static const uint64_t chunkBytes = env_kb("C2D_FETCH_CHUNK_KB", 4096), unitBytes = env_kb("C2D_FETCH_UNIT_KB", 512),
                          tailPiece = env_kb("C2D_FETCH_TAIL_KB", 256);
    c2d_host::CopyPool* pool = copy_pool(ctx);
    pool->wait();
    advise_huge_pages(dstHost, total);
    // two staging halves: while the workers drain wave w, the device-to-host copies of wave w + 1 are already in flight
    const uint64_t waveBytes = kFetchWaveBytes;
    const uint64_t waves = (total + waveBytes - 1) / waveBytes;
    if (!ctx->copyStage.reserve(waves > 1 ? 2 * waveBytes : waveBytes))
        return fail(ctx, "pinned host allocation failed");
    struct Unit
    {
        uint64_t off, len; // within the wave
        uint32_t event;    // index of the DMA chunk that carries it
    };
    struct Chunk
    {
        uint64_t off, len;
    };
    const auto wave_bytes = [&](uint64_t w) { return (w + 1) * waveBytes <= total ? waveBytes : total - w * waveBytes; };
    // a mid-size result (a rank's share of a multi-GPU frame, 4-16 MB) is cut into 6 DMA pieces instead: the 256 KB tail
    // pieces would cover all of it, ~8 us of API calls each
    const bool midSize = total < 4 * chunkBytes;
    const uint64_t midChunk = ((total / 6 + (64u << 10)) >> 16) << 16;
    const auto plan = [&](uint64_t w, std::vector<Chunk>& chunks, std::vector<Unit>& units) {
        chunks.clear();
        units.clear();
        const uint64_t bytes = wave_bytes(w);
        for (uint64_t off = 0; off < bytes;)
        {
            uint64_t len = std::min(midSize ? midChunk : chunkBytes, bytes - off);
            if (!midSize && w + 1 == waves && off + len + chunkBytes >= bytes) // tail region of the transfer: small DMA pieces
                len = std::min(tailPiece, bytes - off);
            const uint64_t unit = midSize ? std::min<uint64_t>(unitBytes, 256u << 10) : unitBytes;
            for (uint64_t p = 0; p < len; p += unit)
                units.push_back(Unit{off + p, std::min(unit, len - p), static_cast<uint32_t>(chunks.size())});
            chunks.push_back(Chunk{off, len});
            off += len;
        }
    };
    const size_t maxChunks = static_cast<size_t>(waveBytes / chunkBytes + 2 * chunkBytes / tailPiece + 4);
    while (ctx->copyEvents.size() < 2 * maxChunks)
    {
        cudaEvent_t ev = nullptr;
        C2D_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->copyEvents.push_back(ev);
    }
    char* stageBase = reinterpret_cast<char*>(ctx->copyStage.data);
    const int device = ctx->device;
    std::vector<Chunk> chunkPlan[2];
    std::vector<Unit> unitPlan[2];
    const auto issue = [&](uint64_t w) -> int {
        std::vector<Chunk>& chunks = chunkPlan[w & 1];
        plan(w, chunks, unitPlan[w & 1]);
        char* stage = stageBase + (w & 1) * waveBytes;
        for (size_t k = 0; k < chunks.size(); ++k)
        {
            C2D_CUDA(ctx, cudaMemcpyAsync(stage + chunks[k].off, src + w * waveBytes + chunks[k].off, chunks[k].len,
                                          cudaMemcpyDeviceToHost, s));
            C2D_CUDA(ctx, cudaEventRecord(ctx->copyEvents[(w & 1) * maxChunks + k], s));
        }
        return 0;
    };
    if (int rc = issue(0))
        return rc;
    for (uint64_t w = 0; w < waves; ++w)
    {
        if (w + 1 < waves)
            if (int rc = issue(w + 1))
                return rc;
        // every worker (helpers + this thread) takes the next unit as soon as the chunk that carries it has landed
        std::atomic<uint32_t> next{0};
        std::atomic<int> cudaErr{0};
        const char* stage = stageBase + (w & 1) * waveBytes;
        char* dst = out + w * waveBytes;
        const cudaEvent_t* events = ctx->copyEvents.data() + (w & 1) * maxChunks;
        const Unit* units = unitPlan[w & 1].data();
        const uint32_t count = static_cast<uint32_t>(unitPlan[w & 1].size());
        auto work = [&next, &cudaErr, device, events, units, count, stage, dst](unsigned) {
            cudaSetDevice(device);
            uint32_t landed = 0xffffffffu; // last chunk this thread saw complete (chunks complete in order)
            for (uint32_t k = next.fetch_add(1); k < count; k = next.fetch_add(1))
            {
                if (landed == 0xffffffffu || units[k].event > landed)
                {
                    const cudaError_t e = cudaEventSynchronize(events[units[k].event]);
                    if (e != cudaSuccess)
                    {
                        cudaErr.store(static_cast<int>(e));
                        continue;
                    }
                    landed = units[k].event;
                }
                std::memcpy(dst + units[k].off, stage + units[k].off, units[k].len);
            }
        };
        pool->launch(work);
        work(pool->helpers());
        pool->wait();
        if (cudaErr.load())
            return fail(ctx, std::string("copy_out: ") + cudaGetErrorString(static_cast<cudaError_t>(cudaErr.load())));
    }
    return 0;
}

// Host memory (the library's pinned result buffer) -> caller memory, split over the helper threads.
int c2d_host::copy_host(c2d_gpu_ctx* ctx, const void* srcHost, void* dstHost, uint64_t total)
{
    if (total < kPoolThresholdBytes)
    {
        if (total)
            std::memcpy(dstHost, srcHost, total);
        return 0;
    }
    c2d_host::CopyPool* pool = copy_pool(ctx);
    pool->wait();
    advise_huge_pages(dstHost, total);
    const uint32_t chunks = static_cast<uint32_t>((total + kFetchChunkBytes - 1) / kFetchChunkBytes);
    std::atomic<uint32_t> next{0};
    const char* src = static_cast<const char*>(srcHost);
    char* dst = static_cast<char*>(dstHost);
    auto work = [&next, chunks, total, src, dst](unsigned) {
        for (uint32_t k = next.fetch_add(1); k < chunks; k = next.fetch_add(1))
        {
            const uint64_t off = k * kFetchChunkBytes, len = off + kFetchChunkBytes < total ? kFetchChunkBytes : total - off;
            std::memcpy(dst + off, src + off, len);
        }
    };
    pool->launch(work);
    work(pool->helpers());
    pool->wait();
    return 0;
}

extern "C" {

int c2d_gpu_collide_begin(c2d_gpu_ctx* ctx)
{
    cudaSetDevice(ctx->device);
    if (ctx->frame.open) // abandoned bracket (the caller unwound between begin and end): drain it
    {
        if (ctx->copyPool)
            ctx->copyPool->wait();
        cudaStreamSynchronize(ctx->stream);
        ctx->frame.open = false;
    }
    return pipeline_begin(ctx);
}

int c2d_gpu_prefault(c2d_gpu_ctx* ctx, void* dst, uint64_t bytes)
{
    const uint64_t page = 4096;
    char* base = static_cast<char*>(dst);
    if (!dst || bytes < (64u << 10))
        return 0;
    if (bytes < kPoolThresholdBytes)
    {
        // a mid-size result (a rank's share of a multi-GPU frame): the calling thread has nothing to do until the device
        // finishes, so it takes the ~1000 first-touch page faults itself now instead of inside the copy afterwards
        // (measured at N = 8: 3.9 MB results took 0.4 ms longer than 4.4 MB ones, which the helpers pre-fault)
        for (uint64_t off = 0; off < bytes; off += page)
            *reinterpret_cast<volatile char*>(base + off) = 0;
        return 0;
    }
    c2d_host::CopyPool* pool = copy_pool(ctx);
    advise_huge_pages(dst, bytes);
    const unsigned workers = pool->helpers();
    if (workers == 0)
        return 0;
    pool->launch([=](unsigned index) {
        // interleaved 2 MB stripes: neighbouring threads fault neighbouring regions (and whole huge pages with THP)
        const uint64_t stripe = 2ull << 20;
        for (uint64_t lo = static_cast<uint64_t>(index) * stripe; lo < bytes; lo += static_cast<uint64_t>(workers) * stripe)
        {
            const uint64_t hi = lo + stripe < bytes ? lo + stripe : bytes;
            for (uint64_t off = lo; off < hi; off += page)
                *reinterpret_cast<volatile char*>(base + off) = 0;
        }
    });
    return 0;
}

int c2d_gpu_collide_end(c2d_gpu_ctx* ctx, uint64_t* out_count, c2d_gpu_stats* stats)
{
    cudaSetDevice(ctx->device);
    const int rc = pipeline_end(ctx, stats);
    if (ctx->copyPool)
        ctx->copyPool->wait(); // the pre-fault job must not outlive the caller's storage
    if (out_count)
        *out_count = rc ? 0 : result_count(ctx);
    return rc;
}

int c2d_gpu_collide_fetch(c2d_gpu_ctx* ctx, c2d_pair_record* dst, uint64_t count, c2d_gpu_stats* stats)
{
    cudaSetDevice(ctx->device);
    if (count > result_count(ctx))
        return fail(ctx, "c2d_gpu_collide_fetch: more records requested than the last frame produced");
    if (count && !dst)
        return fail(ctx, "c2d_gpu_collide_fetch: dst is NULL");
    const uint64_t total = count * sizeof(PairOut);
    if (result_on_host(ctx))
    {
        if (int rc = c2d_host::copy_host(ctx, ctx->smallOut.data, dst, total))
            return rc;
    }
    else
    {
        if (int rc = c2d_host::copy_out(ctx, result_ptr(ctx), dst, total))
            return rc;
        ctx->frame.d2h += total;
    }
    if (stats)
        fill_stats(ctx, stats);
    return 0;
}

#ifndef C2D_HAVE_RAYS
int c2d_gpu_raycast(c2d_gpu_ctx* ctx, uint32_t, const c2d_ray*, const c2d_hit_record**, const uint64_t**, c2d_gpu_stats*)
{
    return fail(ctx, "c2d_gpu_raycast: not built");
}
#endif

int c2d_gpu_device_records(c2d_gpu_ctx* ctx, const void** out_device_ptr, uint64_t* out_count)
{
    if (out_device_ptr)
        *out_device_ptr = result_ptr(ctx);
    if (out_count)
        *out_count = result_count(ctx);
    return 0;
}

int c2d_gpu_read_boxes(c2d_gpu_ctx* ctx, uint32_t first, uint32_t count, float* out_tight4, float* out_fat4)
{
    cudaSetDevice(ctx->device);
    if (int rc = flush(ctx))
        return rc;
    // a snapshot's shape table may end in freed ids the device never saw (all shapes removed, or trailing free ids
    // past a capacity step): the arrays grow to the requested range (zero-filled = dead shapes)
    if (static_cast<uint64_t>(first) + count > 0xffffffffull)
        return fail(ctx, "c2d_gpu_read_boxes: range out of bounds");
    if (int rc = ensure_shape_capacity(ctx, first + count))
        return rc;
    if (out_tight4)
        C2D_CUDA(ctx, cudaMemcpyAsync(out_tight4, ctx->d.tight + first, count * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_fat4)
        C2D_CUDA(ctx, cudaMemcpyAsync(out_fat4, ctx->d.fat + first, count * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    C2D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int c2d_gpu_read_shapes(c2d_gpu_ctx* ctx, uint32_t first, uint32_t count, uint8_t* out_type, uint8_t* out_vertex_count,
                        uint8_t* out_flags, uint64_t* out_category, uint64_t* out_mask, float* out_geom32)
{
    cudaSetDevice(ctx->device);
    if (int rc = flush(ctx))
        return rc;
    // a snapshot's shape table may end in freed ids the device never saw (all shapes removed, or trailing free ids
    // past a capacity step): the arrays grow to the requested range (zero-filled = dead shapes)
    if (static_cast<uint64_t>(first) + count > 0xffffffffull)
        return fail(ctx, "c2d_gpu_read_shapes: range out of bounds");
    if (int rc = ensure_shape_capacity(ctx, first + count))
        return rc;
    cudaStream_t s = ctx->stream;
    std::vector<uint32_t> meta(count);
    if (count)
        C2D_CUDA(ctx, cudaMemcpyAsync(meta.data(), ctx->d.meta + first, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (out_category && count)
        C2D_CUDA(ctx, cudaMemcpyAsync(out_category, ctx->d.category + first, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (out_mask && count)
        C2D_CUDA(ctx, cudaMemcpyAsync(out_mask, ctx->d.mask + first, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (out_geom32 && count)
        C2D_CUDA(ctx, cudaMemcpyAsync(out_geom32, ctx->d.geom + static_cast<size_t>(first) * 32,
                                      static_cast<size_t>(count) * 32 * sizeof(float), cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    for (uint32_t i = 0; i < count; ++i)
    {
        if (out_type)
            out_type[i] = static_cast<uint8_t>(meta_type(meta[i]));
        if (out_vertex_count)
            out_vertex_count[i] = static_cast<uint8_t>(meta_count(meta[i]));
        if (out_flags)
            out_flags[i] = static_cast<uint8_t>(((meta[i] & META_ALIVE) ? 1u : 0u) | ((meta[i] & META_ACTIVE) ? 2u : 0u) |
                                                ((meta[i] & META_GHOST) ? 4u : 0u));
    }
    return 0;
}

int c2d_gpu_read_transforms(c2d_gpu_ctx* ctx, uint32_t first, uint32_t count, float* out_xf4, float* out_xa4, float* out_pxf4,
                            float* out_pxa4)
{
    cudaSetDevice(ctx->device);
    if (int rc = flush(ctx))
        return rc;
    if (static_cast<uint64_t>(first) + count > ctx->capacity)
        return fail(ctx, "c2d_gpu_read_transforms: range out of bounds");
    cudaStream_t s = ctx->stream;
    const size_t bytes = static_cast<size_t>(count) * sizeof(float4);
    const struct
    {
        float* dst;
        const float4* src;
    } parts[] = {{out_xf4, ctx->d.xf}, {out_xa4, ctx->d.xa}, {out_pxf4, ctx->d.pxf}, {out_pxa4, ctx->d.pxa}};
    for (const auto& p : parts)
        if (p.dst && count)
            C2D_CUDA(ctx, cudaMemcpyAsync(p.dst, p.src + first, bytes, cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    return 0;
}

int c2d_gpu_write_boxes(c2d_gpu_ctx* ctx, uint32_t first, uint32_t count, const float* tight4, const float* fat4)
{
    ++ctx->mutationEpoch;
    cudaSetDevice(ctx->device);
    if (int rc = flush(ctx))
        return rc;
    // a snapshot's shape table may end in freed ids the device never saw (all shapes removed, or trailing free ids
    // past a capacity step): the arrays grow to the requested range (zero-filled = dead shapes)
    if (static_cast<uint64_t>(first) + count > 0xffffffffull)
        return fail(ctx, "c2d_gpu_write_boxes: range out of bounds");
    if (int rc = ensure_shape_capacity(ctx, first + count))
        return rc;
    cudaStream_t s = ctx->stream;
    if (tight4 && count)
        C2D_CUDA(ctx, cudaMemcpyAsync(ctx->d.tight + first, tight4, count * sizeof(float4), cudaMemcpyHostToDevice, s));
    if (fat4 && count)
        C2D_CUDA(ctx, cudaMemcpyAsync(ctx->d.fat + first, fat4, count * sizeof(float4), cudaMemcpyHostToDevice, s));
    if (tight4 && count)
        k_clear_meta_bits<<<blocks_for(count, 256), 256, 0, s>>>(ctx->d.meta + first, count, META_TIGHT_EXACT);
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    for (int body = 0; body < 3; ++body)
    {
        mark_dirty(ctx, body); // every leaf box may have changed: rebuild, do not just refit
        ctx->tree[body].membershipChanged = true;
        ctx->slab.membershipChanged = true;
    }
    return 0;
}

int c2d_gpu_shape_slots(c2d_gpu_ctx* ctx, uint32_t* out_slots)
{
    if (out_slots)
        *out_slots = ctx->shapeSlots;
    return 0;
}

int c2d_gpu_read_candidates(c2d_gpu_ctx* ctx, uint64_t capacity, uint32_t* out_pairs2, uint64_t* out_count)
{
    cudaSetDevice(ctx->device);
    if (out_count)
        *out_count = ctx->lastCandidates;
    const uint64_t n = std::min<uint64_t>(capacity, ctx->lastCandidates);
    if (n && out_pairs2)
    {
        C2D_CUDA(ctx, cudaMemcpyAsync(out_pairs2, ctx->cand.ptr, n * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
        C2D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

} // extern "C"
