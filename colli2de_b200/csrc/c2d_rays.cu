// c2d_rays.cu — batched Registry::rayCast on the LBVHs (K8).
//
// Replaces, per ray: DynamicBVH::piercingRaycast over the Static, Dynamic and Bullet trees (reference
// DynamicBVH.hpp:775-806 finite / 853-884 infinite; AABB::intersects(ray) AABB.cpp:39-164), the per-hit
// narrowphase raycast(shape, transform, ray) (RaycastShapes.cpp) or, for bullets, sweep(shape, prev, cur, ray)
// (Collision.hpp:323-344), and the ordered std::set of hits (Registry.hpp:1050-1182).
//
// std::set semantics that are reproduced:
//   * per tree, leaf-box hits are keyed on (entry, exit): of several leaves with identical box times only ONE
//     survives, and that happens BEFORE the active check and the narrowphase (DynamicBVH.hpp:798,
//     Registry.hpp:1061-1070).  The reference keeps whichever its traversal meets first; here the LOWEST shape
//     id wins (deterministic; the oracle uses the same rule);
//   * the final set is keyed on (entryTime, exitTime) of the narrowphase hit; on equal keys the first inserted
//     wins, insertion order being Static, Dynamic, Bullet and, inside a tree, ascending leaf-box key.
//
// One thread BLOCK per ray (k_ray_block): the block pops up to blockDim nodes of a shared traversal stack per step,
// collects the leaf-box hits in shared memory, bitonic-sorts them, dedupes, runs the narrowphase one hit per thread,
// bitonic-sorts the survivors, dedupes again and appends 32-byte RaycastHit records to a packed buffer.
//   * <= 64 rays (the reference's one-ray-per-call API): ONE launch of the 256-thread / 2048-hit instantiation, one
//     host synchronisation, results assembled in pinned host memory;
//   * larger batches, per chunk of 2^20 rays: tier 1 = the 64-thread / 128-hit instantiation for every ray; tier 2 =
//     the 256-thread / 2048-hit instantiation for the rays that overflowed tier 1; tier 3 (k_heavy_*) = global scratch
//     for the rays that overflowed tier 2; then a scan of the per-ray counts and a gather into ray order.  The records
//     stay on the device until c2d_gpu_raycast_fetch / c2d_gpu_raycast copies them out.
#include "c2d_ctx.hpp"
#include "c2d_ray.cuh"

#include <algorithm>
#include <chrono>

using namespace c2d_dev;
using namespace c2d_host;

namespace
{

struct RayIn // = c2d_ray
{
    float sx, sy, ex, ey;
    unsigned long long mask;
    uint32_t infinite, _pad;
};
static_assert(sizeof(RayIn) == sizeof(c2d_ray), "ray layout");

struct LeafHit // 16 bytes
{
    float t0, t1;
    uint32_t shape;
    uint32_t order; // phase 1: tree index (0 Static, 1 Dynamic, 2 Bullet); phase 2: insertion index
};

struct TreeView
{
    const uint32_t* sorted;
    const BvhNode* nodes;
    uint32_t n;
};

__device__ __forceinline__ RayQ make_ray(const RayIn& r, V2& second)
{
    RayQ q;
    q.start = mk(r.sx, r.sy);
    second = mk(r.ex, r.ey);
    q.infinite = r.infinite != 0;
    q.d = q.infinite ? second : (second - q.start);
    return q;
}

// PartitioningMethod::Grid casts an infinite ray as the FINITE ray
// Ray{start, start + direction.normalize() * cellSize * 50} through its cells (reference BroadPhaseTree.hpp:22,
// 531-537): only leaves whose box that segment meets are found, and the leaf-box times are fractions of it.  The
// narrowphase still receives the original infinite ray (Registry.hpp:1150-1153).  gridCell == 0: no truncation.
__device__ __forceinline__ RayQ box_ray_of(const RayQ& q, float gridCell)
{
    if (!q.infinite || gridCell <= 0.0f)
        return q;
    const V2 unit = normalize(q.d);
    const V2 span = (unit * gridCell) * 50.0f;
    const V2 end = q.start + span;
    RayQ b;
    b.start = q.start;
    b.d = end - q.start;
    b.infinite = false;
    return b;
}

__global__ void __launch_bounds__(1024) k_scan_local(uint32_t* data, uint32_t n, uint32_t* blockSums)
{
    __shared__ uint32_t warpSums[32];
    const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
    const uint32_t v = i < n ? data[i] : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = warpSums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o)
                w += t;
        }
        warpSums[lane] = w; // inclusive over warps
    }
    __syncthreads();
    const uint32_t warpOff = warp ? warpSums[warp - 1] : 0;
    if (i < n)
        data[i] = warpOff + incl - v;
    if (threadIdx.x == 1023)
        blockSums[blockIdx.x] = warpOff + incl;
}

__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* blockSums, uint32_t nblocks, uint32_t* total)
{
    // single block: sequential chunks of 1024 with a running carry
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nblocks; base += 1024)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? blockSums[i] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (lane == 31)
            warpSums[warp] = incl;
        __syncthreads();
        if (warp == 0)
        {
            uint32_t w = warpSums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o)
                    w += t;
            }
            warpSums[lane] = w;
        }
        __syncthreads();
        const uint32_t off = carry + (warp ? warpSums[warp - 1] : 0);
        if (i < nblocks)
            blockSums[i] = off + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023)
            carry = off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *total = carry;
}

__global__ void __launch_bounds__(1024) k_scan_add(uint32_t* data, uint32_t n, const uint32_t* __restrict__ blockSums)
{
    const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
    if (i < n)
        data[i] += blockSums[blockIdx.x];
}

// ---- orderings of the leaf-box hits and of the narrow hits ------------------------------------------------------
struct ByTreeTimeShape
{
    __device__ bool operator()(const LeafHit& a, const LeafHit& b) const
    {
        if (a.order != b.order)
            return a.order < b.order;
        if (a.t0 != b.t0)
            return a.t0 < b.t0;
        if (a.t1 != b.t1)
            return a.t1 < b.t1;
        return a.shape < b.shape;
    }
};
struct ByTimeInsertion
{
    __device__ bool operator()(const LeafHit& a, const LeafHit& b) const
    {
        if (a.t0 != b.t0)
            return a.t0 < b.t0;
        if (a.t1 != b.t1)
            return a.t1 < b.t1;
        return a.order < b.order;
    }
};

TreeView view_of(const c2d_gpu_ctx* ctx, int body)
{
    TreeView v;
    v.sorted = ctx->tree[body].vals.ptr;
    v.nodes = ctx->tree[body].nodes.ptr;
    v.n = ctx->tree[body].n;
    return v;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------
// k_ray_block<THREADS, CAP, STACK>: one block per ray.  A ray with more than CAP leaf-box hits (or a traversal
// frontier above STACK) sets the overflow flag of its header and is redone by the next tier.
// ---------------------------------------------------------------------------------------------------------
// two instantiations of k_ray_block<threads, leaf-box hits held in shared memory, traversal frontier>
constexpr int kSmallThreads = 256, kSmallCap = 2048, kSmallStack = 2048; // 40 KB: the latency path and the second tier
constexpr int kLightThreads = 64, kLightCap = 128, kLightStack = 128;    // 2.5 KB: first tier of the throughput path

struct SmallHeader // per ray, downloaded in one copy
{
    uint32_t count;    // final hits
    uint32_t base;     // first record in the packed output
    uint32_t leafHits; // leaf-box hits (statistics)
    uint32_t overflow; // 1: fall back to the general pipeline
};

template <typename Less>
__device__ void bitonic_sort_shared(LeafHit* a, int n, Less less)
{
    // n is a power of two; all threads of the block participate
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            for (int i = threadIdx.x; i < n; i += blockDim.x)
            {
                const int ixj = i ^ j;
                if (ixj > i)
                {
                    const bool up = (i & k) == 0;
                    const LeafHit x = a[i], y = a[ixj];
                    if (up ? less(y, x) : less(x, y))
                    {
                        a[i] = y;
                        a[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
}

__device__ __forceinline__ bool narrow_ray_hit(const ShapeArrays& s, const LeafHit& h, uint32_t tree, const RayQ& q,
                                               V2 second, float& a, float& b)
{
    const uint32_t meta = s.meta[h.shape];
    ShapeRef S;
    S.type = static_cast<int>(meta_type(meta));
    S.count = static_cast<int>(meta_count(meta));
    S.g = s.geom + static_cast<size_t>(h.shape) * 32;
    const float4 x = s.xf[h.shape];
    XF cur;
    cur.tx = x.x;
    cur.ty = x.y;
    cur.s = x.z;
    cur.c = x.w;
    if (tree == 2u)
    {
        const float4 p = s.pxf[h.shape], pa = s.pxa[h.shape], ca = s.xa[h.shape];
        XF prev;
        prev.tx = p.x;
        prev.ty = p.y;
        prev.s = p.z;
        prev.c = p.w;
        return ray_sweep(S, motion_between(prev, pa.x, pa.y, pa.z, cur, ca.x), q, second, a, b);
    }
    return ray_shape(S, cur, q, second, a, b);
}

template <int THREADS, int CAP, int STACK>
__global__ void __launch_bounds__(THREADS)
    k_ray_block(const RayIn* __restrict__ rays, const uint32_t* __restrict__ rayIndex, uint32_t n, TreeView t0, TreeView t1,
                TreeView t2, ShapeArrays s, float gridCell, bool firstHitOnly, SmallHeader* __restrict__ headers,
                uint32_t* __restrict__ cursor, c2d_hit_record* __restrict__ out, uint32_t outCapacity)
{
    __shared__ LeafHit hits[CAP];
    __shared__ int stack[STACK];
    __shared__ int stackSize, hitCount, keptCount, overflow, outBase;
    if (blockIdx.x >= n)
        return;
    const uint32_t rayIdx = rayIndex ? rayIndex[blockIdx.x] : blockIdx.x;
    V2 second;
    const RayIn r = rays[rayIdx];
    const RayQ q = make_ray(r, second);
    const RayQ bq = box_ray_of(q, gridCell);
    const TreeView trees[3] = {t0, t1, t2};
    if (threadIdx.x == 0)
    {
        stackSize = hitCount = keptCount = overflow = 0;
    }
    __syncthreads();

    // ---- traversal: (tree << 28 | node) entries, up to 256 popped per step
    if (threadIdx.x < 3)
    {
        const TreeView& t = trees[threadIdx.x];
        if (t.n == 1)
        {
            const uint32_t shape = t.sorted[0];
            const float4 b = s.fat[shape];
            float a0, a1;
            if ((s.category[shape] & r.mask) != 0 && ray_box(bq, b.x, b.y, b.z, b.w, a0, a1))
                hits[atomicAdd(&hitCount, 1)] = LeafHit{a0, a1, shape, threadIdx.x};
        }
        else if (t.n > 1)
            stack[atomicAdd(&stackSize, 1)] = static_cast<int>(threadIdx.x << 28);
    }
    __syncthreads();
    while (true)
    {
        const int size = stackSize;
        if (size == 0 || overflow)
            break;
        const int take = size < THREADS ? size : THREADS;
        int entry = -1;
        if (static_cast<int>(threadIdx.x) < take)
            entry = stack[size - 1 - threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0)
            stackSize = size - take;
        __syncthreads();
        if (entry >= 0)
        {
            const uint32_t tree = static_cast<uint32_t>(entry) >> 28;
            const BvhNode N = trees[tree].nodes[entry & 0x0fffffff];
            float a0, a1;
            if ((N.catL & r.mask) != 0 && ray_box(bq, N.boxL.x, N.boxL.y, N.boxL.z, N.boxL.w, a0, a1))
            {
                if (N.left < 0)
                {
                    const int at = atomicAdd(&hitCount, 1);
                    if (at < CAP)
                        hits[at] = LeafHit{a0, a1, leaf_shape(N.left), tree};
                    else
                        overflow = 1;
                }
                else
                {
                    const int at = atomicAdd(&stackSize, 1);
                    if (at < STACK)
                        stack[at] = static_cast<int>(tree << 28) | N.left;
                    else
                        overflow = 1;
                }
            }
            if ((N.catR & r.mask) != 0 && ray_box(bq, N.boxR.x, N.boxR.y, N.boxR.z, N.boxR.w, a0, a1))
            {
                if (N.right < 0)
                {
                    const int at = atomicAdd(&hitCount, 1);
                    if (at < CAP)
                        hits[at] = LeafHit{a0, a1, leaf_shape(N.right), tree};
                    else
                        overflow = 1;
                }
                else
                {
                    const int at = atomicAdd(&stackSize, 1);
                    if (at < STACK)
                        stack[at] = static_cast<int>(tree << 28) | N.right;
                    else
                        overflow = 1;
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (overflow)
    {
        if (threadIdx.x == 0)
            headers[rayIdx] = SmallHeader{0u, 0u, static_cast<uint32_t>(hitCount), 1u};
        return;
    }
    const int c = hitCount;
    // ---- sort the leaf-box hits by (tree, entry, exit, shape); pad to a power of two with +inf keys
    int padded = 1;
    while (padded < c)
        padded <<= 1;
    for (int i = c + threadIdx.x; i < padded; i += blockDim.x)
        hits[i] = LeafHit{FLT_MAX, FLT_MAX, 0xffffffffu, 0xffffffffu};
    __syncthreads();
    bitonic_sort_shared(hits, padded, ByTreeTimeShape());

    if (firstHitOnly)
    {
        // the reference's countdown is inherently sequential (Registry.hpp:866-1048); one thread walks the sorted sets
        if (threadIdx.x == 0)
        {
            bool have = false;
            float bestEntry = FLT_MAX;
            LeafHit best = hits[0];
            int k = 0;
            for (uint32_t tree = 0; tree < 3; ++tree)
            {
                const int b = k;
                while (k < c && hits[k].order == tree)
                    ++k;
                uint32_t toCheck = 0;
                for (int j = b; j < k; ++j)
                    if (j == b || hits[j].t0 != hits[j - 1].t0 || hits[j].t1 != hits[j - 1].t1)
                        ++toCheck;
                for (int j = b; j < k; ++j)
                {
                    if (j != b && hits[j].t0 == hits[j - 1].t0 && hits[j].t1 == hits[j - 1].t1)
                        continue;
                    if (!(s.meta[hits[j].shape] & META_ACTIVE))
                        continue;
                    float a0 = 0.0f, a1 = 0.0f;
                    if (narrow_ray_hit(s, hits[j], tree, q, second, a0, a1) && a0 < bestEntry)
                    {
                        bestEntry = a0;
                        best = LeafHit{a0, a1, hits[j].shape, 0u};
                        have = true;
                        toCheck = 5;
                    }
                    if (--toCheck == 0)
                        break;
                }
            }
            uint32_t base = 0;
            if (have)
            {
                base = atomicAdd(cursor, 1u);
                if (base < outCapacity)
                {
                    const V2 en = q.start + q.d * best.t0, ex = q.start + q.d * best.t1;
                    out[base] = c2d_hit_record{s.entity[best.shape], best.shape, {en.x, en.y}, {ex.x, ex.y}, best.t0, best.t1};
                }
            }
            headers[rayIdx] = SmallHeader{have ? 1u : 0u, base, static_cast<uint32_t>(c), 0u};
        }
        return;
    }

    // ---- leaf-box set dedupe, active check, narrowphase: one sorted entry per thread; survivors keep their sorted
    // position as insertion index and are compacted in any order (the final sort orders them)
    LeafHit mine[CAP / THREADS];
    bool keep[CAP / THREADS];
#pragma unroll
    for (int u = 0; u < CAP / THREADS; ++u)
    {
        const int i = threadIdx.x + u * THREADS;
        keep[u] = false;
        if (i < c)
        {
            const LeafHit h = hits[i];
            const bool dup = i > 0 && hits[i - 1].order == h.order && hits[i - 1].t0 == h.t0 && hits[i - 1].t1 == h.t1;
            if (!dup && (s.meta[h.shape] & META_ACTIVE))
            {
                float a0 = 0.0f, a1 = 0.0f;
                if (narrow_ray_hit(s, h, h.order, q, second, a0, a1))
                {
                    keep[u] = true;
                    mine[u] = LeafHit{a0, a1, h.shape, static_cast<uint32_t>(i)};
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < CAP / THREADS; ++u)
        if (keep[u])
            hits[atomicAdd(&keptCount, 1)] = mine[u];
    __syncthreads();
    const int kept = keptCount;
    padded = 1;
    while (padded < kept)
        padded <<= 1;
    for (int i = kept + threadIdx.x; i < padded; i += blockDim.x)
        hits[i] = LeafHit{FLT_MAX, FLT_MAX, 0xffffffffu, 0xffffffffu};
    __syncthreads();
    bitonic_sort_shared(hits, padded, ByTimeInsertion());
    // ---- final std::set: drop entries whose (entry, exit) equals the previous one; ordered compaction by a block scan
    __shared__ int scan[THREADS];
    int local = 0;
    constexpr int kPer = CAP / THREADS;
    bool flag[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u)
    {
        const int i = threadIdx.x * kPer + u;
        flag[u] = i < kept && (i == 0 || hits[i - 1].t0 != hits[i].t0 || hits[i - 1].t1 != hits[i].t1);
        local += flag[u] ? 1 : 0;
    }
    scan[threadIdx.x] = local;
    __syncthreads();
    for (int o = 1; o < THREADS; o <<= 1)
    {
        const int v = threadIdx.x >= o ? scan[threadIdx.x - o] : 0;
        __syncthreads();
        scan[threadIdx.x] += v;
        __syncthreads();
    }
    const int total = scan[THREADS - 1];
    if (threadIdx.x == 0)
        outBase = static_cast<int>(atomicAdd(cursor, static_cast<uint32_t>(total)));
    __syncthreads();
    int pos = outBase + scan[threadIdx.x] - local;
#pragma unroll
    for (int u = 0; u < kPer; ++u)
    {
        const int i = threadIdx.x * kPer + u;
        if (flag[u])
        {
            if (static_cast<uint32_t>(pos) < outCapacity)
            {
                const LeafHit h = hits[i];
                const V2 en = q.start + q.d * h.t0, ex = q.start + q.d * h.t1;
                out[pos] = c2d_hit_record{s.entity[h.shape], h.shape, {en.x, en.y}, {ex.x, ex.y}, h.t0, h.t1};
            }
            ++pos;
        }
    }
    if (threadIdx.x == 0)
        headers[rayIdx] = SmallHeader{static_cast<uint32_t>(total), static_cast<uint32_t>(outBase), static_cast<uint32_t>(c), 0u};
}

// ---------------------------------------------------------------------------------------------------------
// Heavy rays of a small batch: a ray whose leaf-box hits do not fit k_ray_small's shared memory (e.g. one ray through
// 100 k stacked bullets, the reference's "Bullet raycast among 100k entities" benchmark).  Same semantics, global
// scratch instead of shared memory, and the leaf scan spread over the whole GPU:
//   k_heavy_scan<false>  counts the leaves of the three trees whose box the ray hits (brute force over the leaf
//                        arrays — a ray with thousands of box hits touches a good part of the tree anyway);
//   k_heavy_scan<true>   stores them (warp-aggregated append into the ray's scratch segment);
//   k_heavy_resolve      one 1024-thread block per ray: sort, set-dedupe, narrowphase, sort, set-dedupe, emit.
// ---------------------------------------------------------------------------------------------------------
constexpr int kHeavyThreads = 1024;

struct HeavyRay
{
    uint32_t count;  // leaf-box hits
    uint32_t cursor; // append position during the fill
    uint32_t offset; // first scratch entry of this ray
    uint32_t cap;    // power of two >= count
};

template <bool FILL>
__global__ void __launch_bounds__(256)
    k_heavy_scan(const RayIn* __restrict__ rays, const uint32_t* __restrict__ rayIndex, TreeView t0, TreeView t1, TreeView t2,
                 const float4* __restrict__ fat, const uint64_t* __restrict__ category, float gridCell,
                 HeavyRay* __restrict__ heavy, LeafHit* __restrict__ scratch)
{
    const uint32_t rayIdx = blockIdx.y; // position in this launch's list of heavy rays
    V2 second;
    const RayIn r = rays[rayIndex ? rayIndex[rayIdx] : rayIdx];
    const RayQ q = make_ray(r, second);
    const RayQ bq = box_ray_of(q, gridCell);
    const TreeView trees[3] = {t0, t1, t2};
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t base = FILL ? heavy[rayIdx].offset : 0u;
    uint32_t local = 0;
    for (uint32_t tree = 0; tree < 3; ++tree)
    {
        const TreeView& t = trees[tree];
        const uint32_t stride = gridDim.x * blockDim.x;
        const uint32_t rounds = (t.n + stride - 1) / stride; // warp-uniform trip count: the ballots below need all lanes
        for (uint32_t k = 0; k < rounds; ++k)
        {
            const uint32_t i = k * stride + blockIdx.x * blockDim.x + threadIdx.x;
            bool hit = false;
            float a0 = 0.0f, a1 = 0.0f;
            uint32_t shape = 0;
            if (i < t.n)
            {
                shape = t.sorted[i];
                const float4 b = fat[shape];
                hit = (category[shape] & r.mask) != 0 && ray_box(bq, b.x, b.y, b.z, b.w, a0, a1);
            }
            if (FILL)
            {
                const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
                if (ballot)
                {
                    uint32_t first = 0;
                    if (lane == static_cast<uint32_t>(__ffs(ballot) - 1))
                        first = atomicAdd(&heavy[rayIdx].cursor, static_cast<uint32_t>(__popc(ballot)));
                    first = __shfl_sync(0xffffffffu, first, __ffs(ballot) - 1);
                    if (hit)
                        scratch[base + first + __popc(ballot & ((1u << lane) - 1u))] = LeafHit{a0, a1, shape, tree};
                }
            }
            else
                local += hit ? 1u : 0u;
        }
    }
    if (!FILL)
    {
        local = __reduce_add_sync(0xffffffffu, local);
        if (lane == 0 && local)
            atomicAdd(&heavy[rayIdx].count, local);
    }
}

__global__ void __launch_bounds__(kHeavyThreads)
    k_heavy_resolve(const RayIn* __restrict__ rays, const uint32_t* __restrict__ rayIndex, ShapeArrays s, bool firstHitOnly,
                    const HeavyRay* __restrict__ heavy, LeafHit* __restrict__ scratchA, LeafHit* __restrict__ scratchB,
                    SmallHeader* __restrict__ headers, uint32_t* __restrict__ cursor, c2d_hit_record* __restrict__ out,
                    uint32_t outCapacity)
{
    __shared__ int keptCount, outBase;
    __shared__ int scan[kHeavyThreads];
    const HeavyRay hr = heavy[blockIdx.x];
    const uint32_t rayIdx = rayIndex ? rayIndex[blockIdx.x] : blockIdx.x;
    V2 second;
    const RayIn r = rays[rayIdx];
    const RayQ q = make_ray(r, second);
    const int c = static_cast<int>(hr.count);
    LeafHit* hits = scratchA + hr.offset;
    LeafHit* keptHits = scratchB + hr.offset;
    if (threadIdx.x == 0)
        keptCount = 0;
    for (int i = c + threadIdx.x; i < static_cast<int>(hr.cap); i += blockDim.x)
        hits[i] = LeafHit{FLT_MAX, FLT_MAX, 0xffffffffu, 0xffffffffu};
    __syncthreads();
    bitonic_sort_shared(hits, static_cast<int>(hr.cap), ByTreeTimeShape());

    if (firstHitOnly)
    {
        // same sequential countdown as k_ray_small / k_ray_fill (Registry.hpp:866-1048)
        if (threadIdx.x == 0)
        {
            bool have = false;
            float bestEntry = FLT_MAX;
            LeafHit best = hits[0];
            int k = 0;
            for (uint32_t tree = 0; tree < 3; ++tree)
            {
                const int b = k;
                while (k < c && hits[k].order == tree)
                    ++k;
                uint32_t toCheck = 0;
                for (int j = b; j < k; ++j)
                    if (j == b || hits[j].t0 != hits[j - 1].t0 || hits[j].t1 != hits[j - 1].t1)
                        ++toCheck;
                for (int j = b; j < k; ++j)
                {
                    if (j != b && hits[j].t0 == hits[j - 1].t0 && hits[j].t1 == hits[j - 1].t1)
                        continue;
                    if (!(s.meta[hits[j].shape] & META_ACTIVE))
                        continue;
                    float a0 = 0.0f, a1 = 0.0f;
                    if (narrow_ray_hit(s, hits[j], tree, q, second, a0, a1) && a0 < bestEntry)
                    {
                        bestEntry = a0;
                        best = LeafHit{a0, a1, hits[j].shape, 0u};
                        have = true;
                        toCheck = 5;
                    }
                    if (--toCheck == 0)
                        break;
                }
            }
            uint32_t base = 0;
            if (have)
            {
                base = atomicAdd(cursor, 1u);
                if (base < outCapacity)
                {
                    const V2 en = q.start + q.d * best.t0, ex = q.start + q.d * best.t1;
                    out[base] = c2d_hit_record{s.entity[best.shape], best.shape, {en.x, en.y}, {ex.x, ex.y}, best.t0, best.t1};
                }
            }
            headers[rayIdx] = SmallHeader{have ? 1u : 0u, base, static_cast<uint32_t>(c), 0u};
        }
        return;
    }

    // leaf-box set dedupe, active check, narrowphase; survivors carry their sorted position as insertion index
    for (int i = threadIdx.x; i < c; i += blockDim.x)
    {
        const LeafHit h = hits[i];
        const bool dup = i > 0 && hits[i - 1].order == h.order && hits[i - 1].t0 == h.t0 && hits[i - 1].t1 == h.t1;
        if (!dup && (s.meta[h.shape] & META_ACTIVE))
        {
            float a0 = 0.0f, a1 = 0.0f;
            if (narrow_ray_hit(s, h, h.order, q, second, a0, a1))
                keptHits[atomicAdd(&keptCount, 1)] = LeafHit{a0, a1, h.shape, static_cast<uint32_t>(i)};
        }
    }
    __syncthreads();
    const int kept = keptCount;
    int padded = 1;
    while (padded < kept)
        padded <<= 1;
    for (int i = kept + threadIdx.x; i < padded; i += blockDim.x)
        keptHits[i] = LeafHit{FLT_MAX, FLT_MAX, 0xffffffffu, 0xffffffffu};
    __syncthreads();
    bitonic_sort_shared(keptHits, padded, ByTimeInsertion());

    // final std::set: an entry whose (entry, exit) equals its predecessor's is dropped; ordered compaction
    const int per = (kept + kHeavyThreads - 1) / kHeavyThreads;
    const int lo = static_cast<int>(threadIdx.x) * per, hi = lo + per < kept ? lo + per : kept;
    int local = 0;
    for (int i = lo; i < hi; ++i)
        local += (i == 0 || keptHits[i - 1].t0 != keptHits[i].t0 || keptHits[i - 1].t1 != keptHits[i].t1) ? 1 : 0;
    scan[threadIdx.x] = local;
    __syncthreads();
    for (int o = 1; o < kHeavyThreads; o <<= 1)
    {
        const int v = static_cast<int>(threadIdx.x) >= o ? scan[threadIdx.x - o] : 0;
        __syncthreads();
        scan[threadIdx.x] += v;
        __syncthreads();
    }
    const int total = scan[kHeavyThreads - 1];
    if (threadIdx.x == 0)
        outBase = static_cast<int>(atomicAdd(cursor, static_cast<uint32_t>(total)));
    __syncthreads();
    int pos = outBase + scan[threadIdx.x] - local;
    for (int i = lo; i < hi; ++i)
        if (i == 0 || keptHits[i - 1].t0 != keptHits[i].t0 || keptHits[i - 1].t1 != keptHits[i].t1)
        {
            if (static_cast<uint32_t>(pos) < outCapacity)
            {
                const LeafHit h = keptHits[i];
                const V2 en = q.start + q.d * h.t0, ex = q.start + q.d * h.t1;
                out[pos] = c2d_hit_record{s.entity[h.shape], h.shape, {en.x, en.y}, {ex.x, ex.y}, h.t0, h.t1};
            }
            ++pos;
        }
    if (threadIdx.x == 0)
        headers[rayIdx] = SmallHeader{static_cast<uint32_t>(total), static_cast<uint32_t>(outBase), static_cast<uint32_t>(c), 0u};
}

// The blocks of k_ray_small / k_heavy_resolve pack their hits in completion order; these two bring them into ray order.
__global__ void __launch_bounds__(256)
    k_header_counts(const SmallHeader* __restrict__ headers, uint32_t n, uint32_t* __restrict__ counts,
                    unsigned long long* __restrict__ leafHitSum)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t leaf = 0;
    if (i < n)
    {
        const SmallHeader h = headers[i];
        counts[i] = h.count;
        leaf = h.leafHits;
    }
    leaf = __reduce_add_sync(0xffffffffu, leaf);
    if ((threadIdx.x & 31u) == 0 && leaf)
        atomicAdd(leafHitSum, static_cast<unsigned long long>(leaf));
}

// Rays of `rayIndex[0..n)` (nullptr: 0..n-1) whose header carries the overflow flag -> outIndex (any order), *outCount.
__global__ void __launch_bounds__(256)
    k_collect_overflow(const SmallHeader* __restrict__ headers, const uint32_t* __restrict__ rayIndex, uint32_t n,
                       uint32_t* __restrict__ outIndex, uint32_t* __restrict__ outCount)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ray = 0;
    bool over = false;
    if (i < n)
    {
        ray = rayIndex ? rayIndex[i] : i;
        over = headers[ray].overflow != 0;
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, over);
    if (!ballot)
        return;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == static_cast<uint32_t>(__ffs(ballot) - 1))
        base = atomicAdd(outCount, static_cast<uint32_t>(__popc(ballot)));
    base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
    if (over)
        outIndex[base + __popc(ballot & ((1u << lane) - 1u))] = ray;
}

__global__ void __launch_bounds__(256)
    k_header_gather(const SmallHeader* __restrict__ headers, const uint32_t* __restrict__ offsets, uint32_t n,
                    const c2d_hit_record* __restrict__ packed, c2d_hit_record* __restrict__ ordered)
{
    const uint32_t ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= n)
        return;
    const SmallHeader h = headers[ray];
    const uint4* src = reinterpret_cast<const uint4*>(packed + h.base);
    uint4* dst = reinterpret_cast<uint4*>(ordered + offsets[ray]);
    for (uint32_t k = threadIdx.x & 31u; k < 2u * h.count; k += 32u) // a record is two 16-byte words
        dst[k] = src[k];
}

int c2d_host::exclusive_scan_u32(c2d_gpu_ctx* ctx, uint32_t* data, uint32_t n, DevBuf<uint32_t>& sums, uint32_t* dTotal)
{
    const uint32_t nblocks = (n + 1023) / 1024;
    C2D_CUDA(ctx, sums.reserve(nblocks + 1));
    k_scan_local<<<nblocks, 1024, 0, ctx->stream>>>(data, n, sums.ptr);
    k_scan_sums<<<1, 1024, 0, ctx->stream>>>(sums.ptr, nblocks, dTotal);
    k_scan_add<<<nblocks, 1024, 0, ctx->stream>>>(data, n, sums.ptr);
    ctx->launches += 3;
    return 0;
}

namespace
{

constexpr uint32_t kLatencyBatch = 64;     // <= this many rays: the one-kernel, one-sync latency path
constexpr uint32_t kRayChunk = 1u << 20;   // rays per pass of the throughput path
constexpr uint32_t kHeavyGroup = 4096;     // heavy rays per launch (grid.y limit, scratch size)
constexpr uint64_t kHeavyScratchMax = 1ull << 27; // leaf-hit entries per group (2 x 2 GB of scratch)

struct RayRun
{
    c2d_gpu_ctx* ctx;
    TreeView t0, t1, t2;
    float gridCell;
    bool firstHitOnly;
};

// Heavy rays `idx[0..count)` (positions in `dr`; nullptr = 0..count-1): count pass.  Fills hostHeavy[i].count.
int heavy_count(const RayRun& run, const RayIn* dr, const uint32_t* dIdx, uint32_t count, HeavyRay* dHeavy, HeavyRay* hHeavy)
{
    c2d_gpu_ctx* ctx = run.ctx;
    cudaStream_t s = ctx->stream;
    C2D_CUDA(ctx, cudaMemsetAsync(dHeavy, 0, static_cast<size_t>(count) * sizeof(HeavyRay), s));
    for (uint32_t g = 0; g < count; g += kHeavyGroup)
    {
        const uint32_t nb = std::min(kHeavyGroup, count - g);
        const dim3 grid(std::max(1u, 296u / nb), nb);
        k_heavy_scan<false><<<grid, 256, 0, s>>>(dr, dIdx ? dIdx + g : nullptr, run.t0, run.t1, run.t2, ctx->d.fat, ctx->d.category,
                                                 run.gridCell, dHeavy + g, nullptr);
        ++ctx->launches;
    }
    C2D_CUDA(ctx, cudaMemcpyAsync(hHeavy, dHeavy, static_cast<size_t>(count) * sizeof(HeavyRay), cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    return 0;
}

// Fill + resolve of the counted heavy rays, appended to `packed` at *dCursor; headers[ray] is rewritten.
int heavy_resolve(const RayRun& run, const RayIn* dr, const uint32_t* dIdx, uint32_t count, HeavyRay* dHeavy, HeavyRay* hHeavy,
                  SmallHeader* dHeaders, uint32_t* dCursor, c2d_hit_record* packed, uint32_t packedCap)
{
    c2d_gpu_ctx* ctx = run.ctx;
    RayBuffers& rb = ctx->rays;
    cudaStream_t s = ctx->stream;
    uint32_t g = 0;
    while (g < count)
    {
        // as many rays as fit the scratch budget (at least one)
        uint64_t scratch = 0;
        uint32_t nb = 0;
        while (g + nb < count && nb < kHeavyGroup)
        {
            HeavyRay& h = hHeavy[g + nb];
            uint32_t cap = 1;
            while (cap < h.count)
                cap <<= 1;
            if (nb > 0 && scratch + cap > kHeavyScratchMax)
                break;
            h.cursor = 0;
            h.offset = static_cast<uint32_t>(scratch);
            h.cap = cap;
            scratch += cap;
            ++nb;
        }
        if (scratch > 0x7fffffffull)
            return fail(ctx, "c2d_gpu_raycast: a single ray hits more than 2^31 leaf boxes");
        C2D_CUDA(ctx, rb.segs.reserve(2 * scratch * sizeof(LeafHit)));
        LeafHit* scratchA = reinterpret_cast<LeafHit*>(rb.segs.ptr);
        LeafHit* scratchB = scratchA + scratch;
        C2D_CUDA(ctx, cudaMemcpyAsync(dHeavy + g, hHeavy + g, static_cast<size_t>(nb) * sizeof(HeavyRay), cudaMemcpyHostToDevice, s));
        const dim3 grid(std::max(1u, 296u / nb), nb);
        k_heavy_scan<true><<<grid, 256, 0, s>>>(dr, dIdx ? dIdx + g : nullptr, run.t0, run.t1, run.t2, ctx->d.fat, ctx->d.category,
                                                run.gridCell, dHeavy + g, scratchA);
        k_heavy_resolve<<<nb, kHeavyThreads, 0, s>>>(dr, dIdx ? dIdx + g : nullptr, ctx->d, run.firstHitOnly, dHeavy + g, scratchA,
                                                     scratchB, dHeaders, dCursor, packed, packedCap);
        ctx->launches += 2;
        // the scratch is reused by the next group: the stream orders the launches, the host copy above is consumed
        // before hHeavy[g..] could change (it never does again)
        g += nb;
    }
    return 0;
}

// ---- latency path: n <= kLatencyBatch rays, results assembled in the pinned host buffer ---------------------------
int raycast_latency(const RayRun& run, uint32_t n, const c2d_ray* rays, uint64_t& totalHits, uint64_t& leafHits, uint64_t& d2h,
                    uint32_t& retries)
{
    c2d_gpu_ctx* ctx = run.ctx;
    RayBuffers& rb = ctx->rays;
    cudaStream_t s = ctx->stream;
    uint32_t outCap = kLatencyBatch * kSmallCap;
    C2D_CUDA(ctx, rb.dRays.reserve(kLatencyBatch));
    C2D_CUDA(ctx, rb.smallHeaders.reserve(kLatencyBatch * sizeof(SmallHeader) + 16));
    C2D_CUDA(ctx, rb.packed.reserve(outCap));
    if (!rb.hostSmall.reserve(kLatencyBatch * sizeof(SmallHeader) + 16) || !rb.hostHits.reserve(outCap))
        return fail(ctx, "pinned host allocation failed");
    SmallHeader* dHeaders = reinterpret_cast<SmallHeader*>(rb.smallHeaders.ptr + 16);
    uint32_t* dCursor = reinterpret_cast<uint32_t*>(rb.smallHeaders.ptr);
    C2D_CUDA(ctx, cudaMemcpyAsync(rb.dRays.ptr, rays, static_cast<size_t>(n) * sizeof(c2d_ray), cudaMemcpyHostToDevice, s));
    const RayIn* dr = reinterpret_cast<const RayIn*>(rb.dRays.ptr);
    for (int attempt = 0; attempt < 2; ++attempt)
    {
        C2D_CUDA(ctx, cudaMemsetAsync(dCursor, 0, 16, s));
        if (attempt == 0)
        {
            k_ray_block<kSmallThreads, kSmallCap, kSmallStack><<<n, kSmallThreads, 0, s>>>(
                dr, nullptr, n, run.t0, run.t1, run.t2, ctx->d, run.gridCell, run.firstHitOnly, dHeaders, dCursor, rb.packed.ptr, outCap);
            ++ctx->launches;
        }
        else
        {
            // some ray overflowed the shared-memory kernel: redo the (small) batch with the global-scratch kernels
            C2D_CUDA(ctx, rb.heavy.reserve(kLatencyBatch * sizeof(HeavyRay)));
            if (!rb.hostHeavy.reserve(kLatencyBatch * sizeof(HeavyRay)))
                return fail(ctx, "pinned host allocation failed");
            HeavyRay* dHeavy = reinterpret_cast<HeavyRay*>(rb.heavy.ptr);
            HeavyRay* hHeavy = reinterpret_cast<HeavyRay*>(rb.hostHeavy.data);
            if (int rc = heavy_count(run, dr, nullptr, n, dHeavy, hHeavy))
                return rc;
            uint64_t hitTotal = 0;
            for (uint32_t i = 0; i < n; ++i)
                hitTotal += hHeavy[i].count;
            if (hitTotal > 0x7fffffffull)
                return fail(ctx, "c2d_gpu_raycast: too many leaf-box hits in one small batch");
            outCap = static_cast<uint32_t>(std::max<uint64_t>(hitTotal, 64));
            C2D_CUDA(ctx, rb.packed.reserve(outCap));
            if (!rb.hostHits.reserve(outCap))
                return fail(ctx, "pinned host allocation failed");
            if (int rc = heavy_resolve(run, dr, nullptr, n, dHeavy, hHeavy, dHeaders, dCursor, rb.packed.ptr, outCap))
                return rc;
            ++retries;
        }
        C2D_CUDA(ctx, cudaEventRecord(ctx->ev[4], s));
        // header block + the first records in ONE copy; a second copy only if there are more
        constexpr uint32_t kEager = 64;
        C2D_CUDA(ctx, cudaMemcpyAsync(rb.hostSmall.data, rb.smallHeaders.ptr, 16 + static_cast<size_t>(n) * sizeof(SmallHeader),
                                      cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaMemcpyAsync(rb.hostHits.data, rb.packed.ptr, kEager * sizeof(c2d_hit_record), cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
        C2D_CUDA(ctx, cudaGetLastError());
        const uint32_t produced = *reinterpret_cast<const uint32_t*>(rb.hostSmall.data);
        const SmallHeader* hh = reinterpret_cast<const SmallHeader*>(rb.hostSmall.data + 16);
        bool overflowed = produced > outCap;
        for (uint32_t i = 0; i < n; ++i)
            overflowed = overflowed || hh[i].overflow != 0;
        if (overflowed)
        {
            if (attempt == 1)
                return fail(ctx, "c2d_gpu_raycast: heavy-ray path overflowed its own counts");
            continue;
        }
        if (produced > kEager)
        {
            C2D_CUDA(ctx, cudaMemcpyAsync(rb.hostHits.data + kEager, rb.packed.ptr + kEager,
                                          static_cast<size_t>(produced - kEager) * sizeof(c2d_hit_record), cudaMemcpyDeviceToHost, s));
            C2D_CUDA(ctx, cudaStreamSynchronize(s));
        }
        // the blocks packed their hits in completion order: bring them into ray order
        if (n == 1)
        {
            rb.hostOffsets.data[0] = 0;
            rb.hostOffsets.data[1] = hh[0].count;
            leafHits = hh[0].leafHits;
            totalHits = hh[0].count;
        }
        else
        {
            if (!rb.hostReorder.reserve(std::max<uint32_t>(produced, 1)))
                return fail(ctx, "pinned host allocation failed");
            for (uint32_t i = 0; i < n; ++i)
            {
                rb.hostOffsets.data[i] = totalHits;
                std::memcpy(rb.hostReorder.data + totalHits, rb.hostHits.data + hh[i].base,
                            static_cast<size_t>(hh[i].count) * sizeof(c2d_hit_record));
                totalHits += hh[i].count;
                leafHits += hh[i].leafHits;
            }
            rb.hostOffsets.data[n] = totalHits;
            std::memcpy(rb.hostHits.data, rb.hostReorder.data, static_cast<size_t>(totalHits) * sizeof(c2d_hit_record));
        }
        rb.hostHits.size = totalHits;
        d2h = 16 + static_cast<uint64_t>(n) * sizeof(SmallHeader) + std::max<uint64_t>(produced, kEager) * sizeof(c2d_hit_record);
        return 0;
    }
    return fail(ctx, "c2d_gpu_raycast: unreachable");
}

// ---- throughput path: one chunk of m rays, block per ray; the ray-ordered records stay on the device in `ordered`,
// hostCounts receives the exclusive per-ray offsets within the chunk ---------------------------------------------
int raycast_chunk(const RayRun& run, const RayIn* dr, uint32_t m, c2d_host::DevBuf<c2d_hit_record>& ordered, uint32_t* hostOffsets32,
                  uint64_t& outCount, uint64_t& leafHits, uint64_t& d2h, uint32_t& retries)
{
    c2d_gpu_ctx* ctx = run.ctx;
    RayBuffers& rb = ctx->rays;
    cudaStream_t s = ctx->stream;
    const size_t headerBytes = 16 + static_cast<size_t>(m) * sizeof(SmallHeader);
    C2D_CUDA(ctx, rb.smallHeaders.reserve(headerBytes));
    C2D_CUDA(ctx, rb.counts.reserve(m));
    C2D_CUDA(ctx, rb.totals.reserve(2));
    SmallHeader* dHeaders = reinterpret_cast<SmallHeader*>(rb.smallHeaders.ptr + 16);
    uint32_t* dCursor = reinterpret_cast<uint32_t*>(rb.smallHeaders.ptr);
    if (rb.packed.capacity == 0)
        C2D_CUDA(ctx, rb.packed.reserve(std::max<size_t>(static_cast<size_t>(m) * 8, 1u << 16)));

    // control block at the head of smallHeaders: [0] packed cursor, [1] rays overflowing tier 1, [2] rays overflowing tier 2
    C2D_CUDA(ctx, rb.heavyIndex.reserve(2 * static_cast<size_t>(m)));
    uint32_t* dIdxB = rb.heavyIndex.ptr;     // rays for the second tier
    uint32_t* dIdxC = rb.heavyIndex.ptr + m; // rays for the global-scratch kernels
    uint32_t control[4] = {0, 0, 0, 0};
    for (int attempt = 0;; ++attempt)
    {
        if (attempt == 4)
            return fail(ctx, "c2d_gpu_raycast: the packed hit buffer kept overflowing");
        const uint32_t cap = static_cast<uint32_t>(std::min<size_t>(rb.packed.capacity, 0xffffffffu));
        C2D_CUDA(ctx, cudaMemsetAsync(dCursor, 0, 16, s));
        // tier 1: 64-thread blocks with a small shared-memory budget - most rays meet a handful of leaves
        k_ray_block<kLightThreads, kLightCap, kLightStack><<<m, kLightThreads, 0, s>>>(
            dr, nullptr, m, run.t0, run.t1, run.t2, ctx->d, run.gridCell, run.firstHitOnly, dHeaders, dCursor, rb.packed.ptr, cap);
        k_collect_overflow<<<blocks_for(m, 256), 256, 0, s>>>(dHeaders, nullptr, m, dIdxB, dCursor + 1);
        ctx->launches += 2;
        C2D_CUDA(ctx, cudaMemcpyAsync(control, dCursor, sizeof(control), cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
        const uint32_t countB = control[1];
        uint32_t countC = 0;
        if (countB)
        {
            // tier 2: the rays tier 1 could not hold, 256-thread blocks with 2048 leaf-box hits of shared memory
            k_ray_block<kSmallThreads, kSmallCap, kSmallStack><<<countB, kSmallThreads, 0, s>>>(
                dr, dIdxB, countB, run.t0, run.t1, run.t2, ctx->d, run.gridCell, run.firstHitOnly, dHeaders, dCursor, rb.packed.ptr, cap);
            k_collect_overflow<<<blocks_for(countB, 256), 256, 0, s>>>(dHeaders, dIdxB, countB, dIdxC, dCursor + 2);
            ctx->launches += 2;
            C2D_CUDA(ctx, cudaMemcpyAsync(control, dCursor, sizeof(control), cudaMemcpyDeviceToHost, s));
            C2D_CUDA(ctx, cudaStreamSynchronize(s));
            countC = control[2];
        }
        C2D_CUDA(ctx, cudaGetLastError());
        d2h += 2 * sizeof(control);
        const uint64_t produced = control[0];
        uint64_t heavyHits = 0;
        HeavyRay* dHeavy = nullptr;
        HeavyRay* hHeavy = nullptr;
        if (countC)
        {
            // tier 3: global scratch, sized from an exact count of the leaf-box hits
            C2D_CUDA(ctx, rb.heavy.reserve(static_cast<size_t>(countC) * sizeof(HeavyRay)));
            if (!rb.hostHeavy.reserve(static_cast<size_t>(countC) * sizeof(HeavyRay)))
                return fail(ctx, "pinned host allocation failed");
            dHeavy = reinterpret_cast<HeavyRay*>(rb.heavy.ptr);
            hHeavy = reinterpret_cast<HeavyRay*>(rb.hostHeavy.data);
            if (int rc = heavy_count(run, dr, dIdxC, countC, dHeavy, hHeavy))
                return rc;
            for (uint32_t i = 0; i < countC; ++i)
                heavyHits += hHeavy[i].count; // upper bound of what the ray can emit
        }
        const uint64_t need = produced + heavyHits;
        if (need > 0xffffffffull)
            return fail(ctx, "c2d_gpu_raycast: more than 2^32 hits in one chunk of rays; split the batch");
        if (need > cap)
        {
            // the packed buffer was too small (first call, or a denser batch than the last one): grow and redo the chunk
            C2D_CUDA(ctx, rb.packed.reserve(need + need / 8));
            ++retries;
            continue;
        }
        if (countC)
            if (int rc = heavy_resolve(run, dr, dIdxC, countC, dHeavy, hHeavy, dHeaders, dCursor, rb.packed.ptr, cap))
                return rc;
        break;
    }

    // completion order -> ray order
    unsigned long long* dLeafSum = reinterpret_cast<unsigned long long*>(dCursor + 2); // control words 2-3 are free again
    C2D_CUDA(ctx, cudaMemsetAsync(dLeafSum, 0, sizeof(unsigned long long), s));
    k_header_counts<<<blocks_for(m, 256), 256, 0, s>>>(dHeaders, m, rb.counts.ptr, dLeafSum);
    ++ctx->launches;
    if (int rc = exclusive_scan_u32(ctx, rb.counts.ptr, m, rb.sums, rb.totals.ptr))
        return rc;
    uint32_t total = 0;
    unsigned long long chunkLeafHits = 0;
    C2D_CUDA(ctx, cudaMemcpyAsync(&chunkLeafHits, dLeafSum, sizeof(chunkLeafHits), cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaMemcpyAsync(&total, rb.totals.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaMemcpyAsync(hostOffsets32, rb.counts.ptr, static_cast<size_t>(m) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    d2h += static_cast<uint64_t>(m) * sizeof(uint32_t) + sizeof(uint32_t);
    leafHits += chunkLeafHits;
    C2D_CUDA(ctx, ordered.reserve(std::max<uint32_t>(total, 1)));
    k_header_gather<<<blocks_for(static_cast<size_t>(m) * 32, 256), 256, 0, s>>>(dHeaders, rb.counts.ptr, m, rb.packed.ptr, ordered.ptr);
    ++ctx->launches;
    C2D_CUDA(ctx, cudaGetLastError());
    outCount = total;
    return 0;
}

// Runs the whole batch.  Afterwards: offsets in rb.hostOffsets; the records either in rb.hostHits (latency path,
// rb.resultOnHost) or, ray-ordered, in the per-chunk device buffers rb.chunkOut / rb.chunkCount.
int raycast_run(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, bool firstHitOnly, c2d_gpu_stats* stats)
{
    NvtxRange nvtx("c2d rayCast batch");
    const auto wall0 = std::chrono::steady_clock::now();
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ctx->launches = 0;
    ctx->h2dBytes = 0;
    if (n && !rays)
        return fail(ctx, "c2d_gpu_raycast: rays is NULL");
    C2D_CUDA(ctx, cudaEventRecord(ctx->ev[0], s));
    if (int rc = ensure_trees(ctx))
        return rc;

    RayBuffers& rb = ctx->rays;
    if (!rb.hostOffsets.reserve(static_cast<size_t>(n) + 1))
        return fail(ctx, "pinned host allocation failed");
    rb.hostOffsets.size = static_cast<size_t>(n) + 1;
    rb.hostOffsets.data[0] = 0;
    rb.hostHits.size = 0;
    rb.resultOnHost = true;
    rb.chunkCount.clear();
    rb.totalHits = 0;
    uint64_t totalHits = 0, d2h = 0, leafHits = 0;
    uint32_t retries = 0;

    RayRun run{ctx, view_of(ctx, 0), view_of(ctx, 1), view_of(ctx, 2),
               ctx->broadphase == C2D_BROADPHASE_GRID ? static_cast<float>(static_cast<int32_t>(ctx->cellSize)) : 0.0f, firstHitOnly};
    // multi-GPU ray sharding (c2d_gpu_set_ray_sharding): every rank is handed the same batch over the same (mirrored) scene
    // and casts the contiguous share [lo, hi) of it; the other rays come back with empty hit lists on this rank
    uint32_t lo = 0, hi = n;
    if (ctx->raySharding && ctx->shardWorld > 1)
    {
        lo = static_cast<uint32_t>(static_cast<uint64_t>(n) * ctx->shardRank / ctx->shardWorld);
        hi = static_cast<uint32_t>(static_cast<uint64_t>(n) * (ctx->shardRank + 1) / ctx->shardWorld);
    }
    const bool sharded = lo != 0 || hi != n;
    if (n > 0 && n <= kLatencyBatch && !sharded)
    {
        if (int rc = raycast_latency(run, n, rays, totalHits, leafHits, d2h, retries))
            return rc;
        ctx->h2dBytes += static_cast<uint64_t>(n) * sizeof(c2d_ray);
    }
    else if (n > 0)
    {
        rb.resultOnHost = false;
        size_t ci = 0;
        for (uint32_t i = 0; i < lo; ++i)
            rb.hostOffsets.data[i] = 0;
        for (uint32_t first = lo; first < hi; first += kRayChunk, ++ci)
        {
            const uint32_t m = std::min(kRayChunk, hi - first);
            C2D_CUDA(ctx, rb.dRays.reserve(m));
            if (!rb.hostOffsets32.reserve(m))
                return fail(ctx, "pinned host allocation failed");
            C2D_CUDA(ctx, cudaMemcpyAsync(rb.dRays.ptr, rays + first, static_cast<size_t>(m) * sizeof(c2d_ray), cudaMemcpyHostToDevice, s));
            ctx->h2dBytes += static_cast<uint64_t>(m) * sizeof(c2d_ray);
            if (rb.chunkOut.size() <= ci)
                rb.chunkOut.emplace_back(std::make_unique<c2d_host::DevBuf<c2d_hit_record>>());
            uint64_t outCount = 0;
            if (int rc = raycast_chunk(run, reinterpret_cast<const RayIn*>(rb.dRays.ptr), m, *rb.chunkOut[ci], rb.hostOffsets32.data,
                                       outCount, leafHits, d2h, retries))
                return rc;
            for (uint32_t i = 0; i < m; ++i)
                rb.hostOffsets.data[first + i] = totalHits + rb.hostOffsets32.data[i];
            totalHits += outCount;
            rb.chunkCount.push_back(outCount);
        }
        for (uint32_t i = hi; i <= n; ++i)
            rb.hostOffsets.data[i] = totalHits;
        C2D_CUDA(ctx, cudaEventRecord(ctx->ev[4], s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
    }
    rb.totalHits = totalHits;
    if (stats)
    {
        std::memset(stats, 0, sizeof(*stats));
        stats->shapes_alive = ctx->members[0].size() + ctx->members[1].size() + ctx->members[2].size();
        stats->candidates = leafHits;
        stats->pairs = totalHits;
        stats->h2d_bytes = ctx->h2dBytes;
        stats->d2h_bytes = d2h;
        stats->kernel_launches = ctx->launches;
        stats->retries = retries;
        if (n > 0)
            cudaEventElapsedTime(&stats->ms_device, ctx->ev[0], ctx->ev[4]);
        stats->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    }
    return 0;
}

// Records of the last raycast_run -> dst (pageable caller memory or the library's own pinned buffer).
int raycast_deliver(c2d_gpu_ctx* ctx, c2d_hit_record* dst, bool dstPinned, c2d_gpu_stats* stats)
{
    RayBuffers& rb = ctx->rays;
    if (rb.resultOnHost)
    {
        if (dst != rb.hostHits.data)
            return c2d_host::copy_host(ctx, rb.hostHits.data, dst, rb.totalHits * sizeof(c2d_hit_record));
        return 0;
    }
    uint64_t at = 0;
    for (size_t ci = 0; ci < rb.chunkCount.size(); ++ci)
    {
        const uint64_t bytes = rb.chunkCount[ci] * sizeof(c2d_hit_record);
        if (dstPinned)
        {
            if (bytes)
                C2D_CUDA(ctx, cudaMemcpyAsync(dst + at, rb.chunkOut[ci]->ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        }
        else if (int rc = c2d_host::copy_out(ctx, rb.chunkOut[ci]->ptr, dst + at, bytes))
            return rc;
        at += rb.chunkCount[ci];
    }
    if (dstPinned)
        C2D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (stats)
        stats->d2h_bytes += rb.totalHits * sizeof(c2d_hit_record);
    return 0;
}

int raycast_impl(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, bool firstHitOnly, const c2d_hit_record** out_hits,
                 const uint64_t** out_offsets, c2d_gpu_stats* stats)
{
    const auto wall0 = std::chrono::steady_clock::now();
    if (int rc = raycast_run(ctx, n, rays, firstHitOnly, stats))
        return rc;
    RayBuffers& rb = ctx->rays;
    if (!rb.resultOnHost)
    {
        if (!rb.hostHits.reserve(std::max<uint64_t>(rb.totalHits, 1)))
            return fail(ctx, "pinned host allocation failed");
        if (int rc = raycast_deliver(ctx, rb.hostHits.data, true, stats))
            return rc;
        rb.hostHits.size = rb.totalHits;
    }
    if (stats)
        stats->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    if (out_hits)
        *out_hits = rb.hostHits.data;
    if (out_offsets)
        *out_offsets = rb.hostOffsets.data;
    return 0;
}

} // namespace

extern "C" int c2d_gpu_raycast(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, const c2d_hit_record** out_hits,
                               const uint64_t** out_offsets, c2d_gpu_stats* stats)
{
    return raycast_impl(ctx, n, rays, false, out_hits, out_offsets, stats);
}

extern "C" int c2d_gpu_raycast_first(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, const c2d_hit_record** out_hits,
                                     const uint64_t** out_offsets, c2d_gpu_stats* stats)
{
    return raycast_impl(ctx, n, rays, true, out_hits, out_offsets, stats);
}

extern "C" int c2d_gpu_raycast_begin(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, int first_hit_only, uint64_t* out_total,
                                     const uint64_t** out_offsets, c2d_gpu_stats* stats)
{
    if (int rc = raycast_run(ctx, n, rays, first_hit_only != 0, stats))
        return rc;
    if (out_total)
        *out_total = ctx->rays.totalHits;
    if (out_offsets)
        *out_offsets = ctx->rays.hostOffsets.data;
    return 0;
}

extern "C" int c2d_gpu_raycast_fetch(c2d_gpu_ctx* ctx, c2d_hit_record* dst, uint64_t count, c2d_gpu_stats* stats)
{
    cudaSetDevice(ctx->device);
    if (count != ctx->rays.totalHits)
        return fail(ctx, "c2d_gpu_raycast_fetch: count must be the total reported by c2d_gpu_raycast_begin");
    if (count && !dst)
        return fail(ctx, "c2d_gpu_raycast_fetch: dst is NULL");
    return raycast_deliver(ctx, dst, false, stats);
}
