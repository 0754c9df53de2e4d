// c2d_ctx.hpp — the context object behind the C ABI (shared by c2d_gpu.cu and c2d_rays.cu).
#pragma once

#include "c2d_internal.hpp"

#include "../../include/c2d_gpu.h"
#include "c2d_copy_pool.hpp"

#include <nvtx3/nvToolsExt.h> // header-only: ranges cost nothing unless a profiler is attached

#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace c2d_host
{
using namespace c2d_dev;

// NVTX range of one pipeline stage (the reference has no tracing; timelines of this library show
// flush / build / traverse / narrow / copy-out per frame in Nsight Systems)
struct NvtxRange
{
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// ---- small RAII helpers ---------------------------------------------------------------------------
template <typename T>
struct PinnedVec
{
    T* data = nullptr;
    size_t size = 0;
    size_t capacity = 0;

    ~PinnedVec()
    {
        if (data)
            cudaFreeHost(data);
    }
    bool reserve(size_t want)
    {
        if (want <= capacity)
            return true;
        size_t cap = capacity ? capacity : 1024;
        while (cap < want)
            cap *= 2;
        T* fresh = nullptr;
        if (cudaHostAlloc(reinterpret_cast<void**>(&fresh), cap * sizeof(T), cudaHostAllocDefault) != cudaSuccess)
            return false;
        if (size)
            std::memcpy(fresh, data, size * sizeof(T));
        if (data)
            cudaFreeHost(data);
        data = fresh;
        capacity = cap;
        return true;
    }
    T* push()
    {
        if (size == capacity && !reserve(size + 1))
            return nullptr;
        return &data[size++];
    }
};

template <typename T>
struct DevBuf
{
    T* ptr = nullptr;
    size_t capacity = 0;
    ~DevBuf()
    {
        if (ptr)
            cudaFree(ptr);
    }
    // contents are NOT preserved
    cudaError_t reserve(size_t want)
    {
        if (want <= capacity)
            return cudaSuccess;
        size_t cap = capacity ? capacity : 1024;
        while (cap < want)
            cap *= 2;
        if (ptr)
            cudaFree(ptr);
        ptr = nullptr;
        capacity = 0;
        const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ptr), cap * sizeof(T));
        if (e == cudaSuccess)
            capacity = cap;
        return e;
    }
};

enum Phase : uint32_t
{
    PH_FLAG = 0,
    PH_ADD = 1,
    PH_TRANS = 2,
    PH_MOVE = 3
};

struct Round
{
    size_t f0, a0, t0, m0; // begin offsets in the four logs; the end is the next round's begin (or the log size)
    int staged = -1;       // >= 0: this round replays a device-resident translation batch instead of log ranges
};

// buffers of the batched rayCast (c2d_rays.cu)
struct RayBuffers
{
    DevBuf<c2d_ray> dRays;
    DevBuf<uint32_t> counts, sums, totals;   // per-ray hit counts -> exclusive offsets (scan)
    DevBuf<unsigned char> smallHeaders;      // 16-byte control block + one SmallHeader per ray
    DevBuf<c2d_hit_record> packed;           // hits in block-completion order, before the gather into ray order
    DevBuf<unsigned char> segs;              // global scratch of the heavy tier (2 x leaf-hit entries)
    PinnedVec<c2d_hit_record> hostHits;      // results of the pointer-returning entry points / the latency path
    PinnedVec<uint64_t> hostOffsets;
    PinnedVec<uint32_t> hostOffsets32;
    PinnedVec<unsigned char> hostSmall;      // latency path: headers downloaded in one copy
    PinnedVec<c2d_hit_record> hostReorder;   // latency path: completion order -> ray order on the host
    // rays with more leaf-box hits than the shared-memory kernel holds (global-scratch path)
    DevBuf<unsigned char> heavy;
    DevBuf<uint32_t> heavyIndex;
    PinnedVec<unsigned char> hostHeavy;
    // result of the last run: on the host (latency path) or ray-ordered per chunk on the device
    bool resultOnHost = true;
    uint64_t totalHits = 0;
    std::vector<std::unique_ptr<DevBuf<c2d_hit_record>>> chunkOut;
    std::vector<uint64_t> chunkCount;
};

// buffers of the uniform-grid broadphase (c2d_grid.cu)
struct GridBuffers
{
    DevBuf<uint32_t> ext, counts, sums, totals, keys, vals, keysTmp, valsTmp, sbody;
    DevBuf<unsigned char> params, sortWs;
    DevBuf<float4> sbox;
    DevBuf<uint64_t> scat;
};

// buffers of the per-entity queries (c2d_query.cu)
struct QueryBuffers
{
    DevBuf<uint32_t> shapes, counter;
    DevBuf<uint8_t> flags;
    DevBuf<c2d_dev::PairOut> out;
    PinnedVec<c2d_dev::PairOut> hostOut;
};

struct ProfSpan
{
    const char* name;
    cudaEvent_t start, stop;
};

// A batch of per-shape translations already resident in HBM (c2d_gpu_stage_translations).
struct StagedBatch
{
    DevBuf<TransRec> records;
    uint32_t n = 0;
    bool touches[3] = {false, false, false};
    bool live = false;
};

// Multi-GPU (c2d_comm.cu): one context per GPU / process, joined by an NCCL communicator.  Shape state is REPLICATED
// on every rank (each rank replays every mutation, or receives the others' translations); what is sharded is the
// per-frame work.
struct CommState
{
    void* comm = nullptr;       // ncclComm_t (NCCL is loaded with dlopen: libc2d_gpu.so has no link-time NCCL dependency)
    bool exchangeMoves = false; // translation records logged on one rank are all-gathered to every rank at the frame's flush
    int resultMode = 0;         // C2D_RESULTS_*
    DevBuf<uint32_t> dHeader;   // [4] mine + [world x 4] per-rank header of the move exchange (count, body mask, 0, 0) + [1] overflow flag
    uint32_t stride = 0;        // records per rank of the one-step exchange, agreed from the previous frame's headers (0: unknown)
    bool fastExchanged = false; // this frame used the one-step exchange: its headers / overflow flag come back with the frame
    uint32_t* hHeader = nullptr; // pinned: [0..3] mine, [4 ..] everybody's
    DevBuf<TransRec> gathered;  // world x stride translation records
    DevBuf<PairOut> outAll;     // gathered results
    DevBuf<uint64_t> dPairCounts;
    uint64_t* hPairCounts = nullptr; // pinned: [0] mine, [1 ..] everybody's
    bool resultsGathered = false;    // the last frame's records are in outAll (gatheredPairs of them)
    uint64_t gatheredPairs = 0;
    cudaEvent_t evX[4] = {};    // exchange start / stop, gather start / stop
    // statistics of the last frame
    uint64_t exchangeBytes = 0, gatherBytes = 0;
    bool exchangeTimed = false, gatherTimed = false;
};

// Morton slab of the Dynamic tree (update_slab in c2d_gpu.cu, kernels in c2d_slab.cuh): this rank's tree holds the leaves
// whose position in the (replicated) Morton order falls into its range; the shapes of higher ranks that may overlap
// them ("halo") are selected every frame and query that tree.
struct SlabState
{
    DevBuf<uint32_t> allKeys, allVals, keysTmp, valsTmp; // Morton order of ALL Dynamic members as of the last rebalance
    DevBuf<unsigned char> sortWs;
    DevBuf<uint32_t> bounds;           // centre bounds of that rebalance (the cell frame of mark / select)
    DevBuf<uint32_t> cellBits;         // 256 x 256 cells touched by the owned shapes' leaf boxes, one bit each
    DevBuf<uint32_t> halo;             // halo shape ids of this frame
    DevBuf<uint32_t> counters;         // [0] halo count, [1] "an owned box spans too many cells"
    uint32_t* hCounters = nullptr;     // pinned: the counters of the last frame (statistics only, copied asynchronously)
    cudaEvent_t ev[2] = {};
    bool timed = false;
    bool valid = false;                // allKeys / allVals / the slab tree topology describe the current membership
    bool membershipChanged = true;     // a Dynamic shape was added or removed since the last rebalance
    uint32_t age = 0;                  // frames since the last rebalance
    uint32_t n = 0;                    // Dynamic members at the last rebalance
    uint32_t lo = 0, hi = 0;           // this rank's positions in the order
    uint32_t owned = 0, laterCount = 0;
};

struct TreeBuf
{
    DevBuf<uint32_t> members; // shape ids, host order
    DevBuf<uint32_t> keys, vals, keysTmp, valsTmp;
    DevBuf<BvhNode> nodes;
    DevBuf<int> nodeParent, leafParent;
    DevBuf<uint32_t> arrival;
    DevBuf<unsigned char> sortWs;
    uint32_t n = 0; // leaves in the built tree
    uint32_t age = 0; // frames since the topology (sort + Karras) was last rebuilt
    bool membershipChanged = true;
};

} // namespace c2d_host

using namespace c2d_dev;  // internal header: only included by the library's own translation units
using namespace c2d_host;

struct c2d_gpu_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};
    std::string error;

    // shapes
    uint32_t capacity = 0;
    uint32_t shapeSlots = 0; // 1 + highest shape id ever added
    ShapeArrays d{};
    std::vector<uint32_t> hostMeta; // body / alive per shape
    std::vector<uint32_t> stamp;    // (round << 2) | phase of the last logged op
    std::vector<uint32_t> memberPos;
    std::vector<uint32_t> members[3];
    bool membersDirty[3] = {false, false, false};
    bool treeDirty[3] = {false, false, false};

    // mutation log
    uint32_t round = 1;
    PinnedVec<FlagRec> flagLogs[2];
    PinnedVec<AddRec> addLogs[2];
    PinnedVec<TransRec> transLogs[2];
    PinnedVec<MoveRec> moveLogs[2];
    PinnedVec<FlagRec>* flagLog = &flagLogs[0]; // the set the host is logging into
    PinnedVec<AddRec>* addLog = &addLogs[0];
    PinnedVec<TransRec>* transLog = &transLogs[0];
    PinnedVec<MoveRec>* moveLog = &moveLogs[0];
    int logIndex = 0;
    cudaEvent_t logConsumed[2] = {}; // recorded after the H2D copies of a set were enqueued
    // Early upload of the translation log: while the host is still inside its moveEntity loop, every ~1 MB of new records
    // goes up on a second stream, so the query's flush only uploads the tail (c2d_gpu.cu: early_upload_translations)
    cudaStream_t copyStream = nullptr;
    cudaEvent_t transApplied = nullptr;  // main stream: the last kernels that read dTransLog
    cudaEvent_t transUploadedEv = nullptr; // copy stream: the last early chunk
    size_t transUploaded = 0;            // records of the current transLog already enqueued into dTransLog
    bool earlyUpload = true;
    // Side stream of a frame: independent kernels of the pipeline run next to each other (the Dynamic x Static traversal
    // next to the Dynamic tree's bottom-up fit, the two filter kernels, the two manifold kernels) - each of them is latency
    // bound and leaves most of the SMs idle
    cudaStream_t sideStream = nullptr;
    cudaEvent_t evPreFit = nullptr;   // main: the Dynamic order / links of this frame are ready, its fit not yet launched
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    bool preFitRecorded = false;      // evPreFit belongs to this frame
    bool overlap = true;              // c2d_gpu_set_overlap
    bool incrementalRefit = true;     // frames that keep a tree's topology refit only the flagged leaves (k_refit_dirty)
    uint32_t transBodies = 0;            // bodies (bit per BodyType) touched by the translations logged since the last flush
    std::vector<Round> rounds;
    DevBuf<FlagRec> dFlagLog;
    DevBuf<AddRec> dAddLog;
    DevBuf<TransRec> dTransLog;
    DevBuf<MoveRec> dMoveLog;
    std::vector<StagedBatch*> staged;

    // trees
    TreeBuf tree[3];
    int broadphase = C2D_BROADPHASE_LBVH;
    uint32_t cellSize = 120;
    uint32_t shardRank = 0, shardWorld = 1;
    // Morton-slab build of the Dynamic tree when world > 1 and the tree has at least this many leaves (0 = never)
    uint32_t slabMinShapes = 65536;
    TreeBuf slabTree;       // the rank-local tree (owned leaves) of the Dynamic body
    bool slabDirty = true;  // the Dynamic shapes changed since slabTree was built
    bool slabInUse = false; // this frame's Dynamic traversals run on slabTree
    bool raySharding = false; // batched rayCast casts only this rank's share of the batch (c2d_gpu_set_ray_sharding)
    c2d_host::SlabState slab;
    c2d_host::CommState comm;
    int traverseVariant = 1;    // 0: divergent emit (k_traverse), 1: convergent loop + staged appends (k_traverse_sync)
    uint32_t rebuildPeriod = 8; // LBVH topology rebuilt every N-th dirty frame, refit in between
    uint64_t mutationEpoch = 0;    // bumped by every entry point that changes what a query would answer
    uint32_t smallSceneMax = 8192; // <= this many live shapes: all-pairs broadphase (c2d_gpu_set_small_scene_limit)

    // frame
    c2d_dev::FrameScratch* dScratch = nullptr;
    c2d_dev::FrameScratch* hScratch = nullptr; // pinned
    c2d_host::DevBuf<uint2> cand, binned;
    c2d_host::DevBuf<uint4> hits;
    c2d_host::DevBuf<c2d_dev::PairOut> out;
    c2d_host::PinnedVec<c2d_dev::PairOut> hostOut;
    // small scenes: the narrowphase writes its records straight into this pinned, device-mapped buffer
    c2d_host::PinnedVec<c2d_dev::PairOut> smallOut;
    bool resultsOnHost = false; // the last frame's records are in smallOut
    bool scratchClean = false;  // the device frame counters are zero (left so by the last small frame)
    uint64_t lastCandidates = 0;
    uint64_t lastPairs = 0;
    struct FrameState // one c2d_gpu_collide_begin .. c2d_gpu_collide_end bracket
    {
        bool open = false;
        bool smallScene = false; // this frame runs k_small_pairs / k_small_narrow instead of the tree pipeline
        std::chrono::steady_clock::time_point wall0{};
        float msEnqueued = 0.0f, msWaited = 0.0f; // host clock since wall0: frame enqueued / device finished
        uint32_t retries = 0;
        uint64_t d2h = 0;
    } frame;
    std::unique_ptr<c2d_host::CopyPool> copyPool; // created by the first large copy-out
    std::vector<cudaEvent_t> copyEvents;          // one per in-flight chunk of copy_out
    c2d_host::PinnedVec<unsigned char> copyStage; // pinned staging of copy_out (one 64 MB wave)
    // reference result order (c2d_gpu_set_result_order)
    struct OrderBuffers
    {
        c2d_host::DevBuf<uint32_t> keys, vals, keysTmp, valsTmp;
        c2d_host::DevBuf<unsigned char> sortWs;
        c2d_host::DevBuf<c2d_dev::PairOut> sorted;
    } order;
    bool referenceOrder = false; // the records of a frame are returned in the reference's order
    bool resultsOrdered = false; // the last frame's records are in order.sorted
    c2d_host::RayBuffers rays;
    c2d_host::GridBuffers grid;
    c2d_host::QueryBuffers query;

    // optional per-kernel timing
    bool profiling = false;
    std::string profileOnly; // non-empty: only the spans of this name are timed (c2d_gpu_set_profile_filter)
    std::vector<c2d_host::ProfSpan> spans;
    std::vector<cudaEvent_t> eventPool;
    std::map<std::string, std::pair<double, uint32_t>> kernelTotals;
    uint32_t launches = 0;
    uint64_t h2dBytes = 0;
};


namespace c2d_host
{
// c2d_gpu.cu
int fail(c2d_gpu_ctx* ctx, const std::string& msg);
int flush(c2d_gpu_ctx* ctx, bool collective = false); // replays the mutation log (collective: from c2d_gpu_collide*)
int ensure_trees(c2d_gpu_ctx* ctx); // flush + rebuild of the dirty trees (launches k_frame_init first)
int upload_members(c2d_gpu_ctx* ctx, int body); // host member list of one tree -> device, if it changed
int copy_out(c2d_gpu_ctx* ctx, const void* srcDevice, void* dstHost, uint64_t bytes); // device -> pageable, threaded
int copy_host(c2d_gpu_ctx* ctx, const void* srcHost, void* dstHost, uint64_t bytes);  // pinned -> pageable, threaded
// c2d_comm.cu (all collective over the group, enqueued on the context's stream)
int comm_allgather(c2d_gpu_ctx* ctx, const void* send, void* recv, size_t bytesPerRank);
// two all-gathers in one NCCL group (one launch): a small header and a payload
int comm_allgather2(c2d_gpu_ctx* ctx, const void* sendA, void* recvA, size_t bytesA, const void* sendB, void* recvB, size_t bytesB);
// rank r's bytes[r] bytes at `send` arrive at recvBase + r * strideBytes on every rank (one grouped ncclBroadcast per rank)
int comm_gatherv(c2d_gpu_ctx* ctx, const void* send, void* recvBase, const size_t* bytes, size_t strideBytes);
int comm_gather_results(c2d_gpu_ctx* ctx); // after a frame: the records of all ranks -> comm.outAll (per the result mode)
void comm_release(c2d_gpu_ctx* ctx);
// c2d_rays.cu
int exclusive_scan_u32(c2d_gpu_ctx* ctx, uint32_t* data, uint32_t n, DevBuf<uint32_t>& sums, uint32_t* dTotal);
// c2d_grid.cu
int grid_broadphase(c2d_gpu_ctx* ctx);

#define C2D_CUDA(ctx, call)                                                                                            \
    do                                                                                                                 \
    {                                                                                                                  \
        const cudaError_t _e = (call);                                                                                 \
        if (_e != cudaSuccess)                                                                                         \
            return ::c2d_host::fail(ctx, std::string(#call) + ": " + cudaGetErrorString(_e));                          \
    } while (0)

inline uint32_t blocks_for(size_t n, uint32_t threads)
{
    return static_cast<uint32_t>((n + threads - 1) / threads);
}
} // namespace c2d_host
