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
// Pipeline for a chunk of rays (one ray per thread): count leaf-box hits -> exclusive scan -> fill the
// per-ray segments, heapsort + dedupe + narrowphase + heapsort + dedupe in place -> scan of the final counts
// -> emit 32-byte RaycastHit records in ray order.
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

// Visits every leaf of `tree` whose (fat) box the ray hits and whose category matches the mask
// (node.matchesMask(maskBits), DynamicBVH.hpp:789).  FILL = false only counts.
template <bool FILL>
__device__ uint32_t ray_tree(const RayQ& ray, unsigned long long mask, const TreeView& tree,
                             const float4* __restrict__ fat, const uint64_t* __restrict__ category, uint32_t treeIndex,
                             LeafHit* __restrict__ out)
{
    uint32_t count = 0;
    if (tree.n == 0)
        return 0;
    float t0, t1;
    if (tree.n == 1)
    {
        const uint32_t s = tree.sorted[0];
        const float4 b = fat[s];
        if ((category[s] & mask) != 0 && ray_box(ray, b.x, b.y, b.z, b.w, t0, t1))
        {
            if (FILL)
                out[0] = LeafHit{t0, t1, s, treeIndex};
            count = 1;
        }
        return count;
    }
    int stack[64];
    int sp = 0;
    int node = 0;
    while (true)
    {
        const float4* np = reinterpret_cast<const float4*>(tree.nodes + node);
        const float4 boxL = __ldg(np);
        const float4 boxR = __ldg(np + 1);
        const int4 links = __ldg(reinterpret_cast<const int4*>(np + 2));
        const ulonglong2 cats = __ldg(reinterpret_cast<const ulonglong2*>(np + 3));
        bool hitL = (cats.x & mask) != 0 && ray_box(ray, boxL.x, boxL.y, boxL.z, boxL.w, t0, t1);
        if (hitL && links.x < 0)
        {
            if (FILL)
                out[count] = LeafHit{t0, t1, tree.sorted[~links.x], treeIndex};
            ++count;
            hitL = false;
        }
        bool hitR = (cats.y & mask) != 0 && ray_box(ray, boxR.x, boxR.y, boxR.z, boxR.w, t0, t1);
        if (hitR && links.y < 0)
        {
            if (FILL)
                out[count] = LeafHit{t0, t1, tree.sorted[~links.y], treeIndex};
            ++count;
            hitR = false;
        }
        if (hitL)
        {
            if (hitR && sp < 64)
                stack[sp++] = links.y;
            node = links.x;
        }
        else if (hitR)
            node = links.y;
        else
        {
            if (sp == 0)
                break;
            node = stack[--sp];
        }
    }
    return count;
}

__global__ void __launch_bounds__(128) k_ray_count(const RayIn* __restrict__ rays, uint32_t n, TreeView t0, TreeView t1,
                                                   TreeView t2, const float4* __restrict__ fat,
                                                   const uint64_t* __restrict__ category, uint32_t* __restrict__ counts,
                                                   float gridCell)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    V2 second;
    const RayIn r = rays[i];
    const RayQ q = box_ray_of(make_ray(r, second), gridCell);
    uint32_t c = ray_tree<false>(q, r.mask, t0, fat, category, 0, nullptr);
    c += ray_tree<false>(q, r.mask, t1, fat, category, 1, nullptr);
    c += ray_tree<false>(q, r.mask, t2, fat, category, 2, nullptr);
    counts[i] = c;
}

// ---- exclusive scan of uint32 (n up to ~16 M): per-block scan + block sums + add -----------------------
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

// ---- in-place heapsort of a per-thread segment ---------------------------------------------------------------
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

template <typename Less>
__device__ void heapsort(LeafHit* a, int n, Less less)
{
    auto sift = [&](int start, int end)
    {
        int root = start;
        const LeafHit v = a[root];
        while (true)
        {
            int child = 2 * root + 1;
            if (child >= end)
                break;
            if (child + 1 < end && less(a[child], a[child + 1]))
                ++child;
            if (!less(v, a[child]))
                break;
            a[root] = a[child];
            root = child;
        }
        a[root] = v;
    };
    for (int s = n / 2 - 1; s >= 0; --s)
        sift(s, n);
    for (int end = n - 1; end > 0; --end)
    {
        const LeafHit t = a[0];
        a[0] = a[end];
        a[end] = t;
        sift(0, end);
    }
}

__global__ void __launch_bounds__(128)
    k_ray_fill(const RayIn* __restrict__ rays, uint32_t n, TreeView t0, TreeView t1, TreeView t2, ShapeArrays s,
               const uint32_t* __restrict__ offsets, LeafHit* __restrict__ segs, uint32_t* __restrict__ finalCounts,
               float gridCell, bool firstHitOnly)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    V2 second;
    const RayIn r = rays[i];
    const RayQ q = make_ray(r, second);
    const RayQ bq = box_ray_of(q, gridCell);
    LeafHit* seg = segs + offsets[i];
    uint32_t c = ray_tree<true>(bq, r.mask, t0, s.fat, s.category, 0, seg);
    c += ray_tree<true>(bq, r.mask, t1, s.fat, s.category, 1, seg + c);
    c += ray_tree<true>(bq, r.mask, t2, s.fat, s.category, 2, seg + c);
    if (c == 0)
    {
        finalCounts[i] = 0;
        return;
    }
    heapsort(seg, static_cast<int>(c), ByTreeTimeShape());

    if (firstHitOnly)
    {
        // Registry::firstHitRayCast (Registry.hpp:866-1048): per tree, walk the leaf-box set in (entry, exit) order;
        // a narrow hit that improves the best entry time re-arms the countdown to 5 more set entries; inactive
        // shapes are skipped WITHOUT counting (the `continue` precedes the decrement).
        bool have = false;
        float bestEntry = FLT_MAX;
        LeafHit best = seg[0];
        uint32_t k = 0;
        for (uint32_t tree = 0; tree < 3; ++tree)
        {
            uint32_t b = k;
            while (k < c && seg[k].order == tree)
                ++k;
            // size of this tree's std::set = distinct (entry, exit) keys
            uint32_t toCheck = 0;
            for (uint32_t j = b; j < k; ++j)
                if (j == b || seg[j].t0 != seg[j - 1].t0 || seg[j].t1 != seg[j - 1].t1)
                    ++toCheck;
            for (uint32_t j = b; j < k; ++j)
            {
                if (j != b && seg[j].t0 == seg[j - 1].t0 && seg[j].t1 == seg[j - 1].t1)
                    continue;
                const LeafHit h = seg[j];
                const uint32_t meta = s.meta[h.shape];
                if (!(meta & META_ACTIVE))
                    continue;
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
                float t0n = 0.0f, t1n = 0.0f;
                bool hit;
                if (tree == 2u)
                {
                    const float4 p = s.pxf[h.shape], pa = s.pxa[h.shape], ca = s.xa[h.shape];
                    XF prev;
                    prev.tx = p.x;
                    prev.ty = p.y;
                    prev.s = p.z;
                    prev.c = p.w;
                    hit = ray_sweep(S, motion_between(prev, pa.x, pa.y, pa.z, cur, ca.x), q, second, t0n, t1n);
                }
                else
                    hit = ray_shape(S, cur, q, second, t0n, t1n);
                if (hit && t0n < bestEntry)
                {
                    bestEntry = t0n;
                    best = LeafHit{t0n, t1n, h.shape, 0u};
                    have = true;
                    toCheck = 5;
                }
                if (--toCheck == 0)
                    break;
            }
        }
        if (have)
            seg[0] = best;
        finalCounts[i] = have ? 1u : 0u;
        return;
    }

    // leaf-box set dedupe, active check, narrowphase — survivors are compacted to the front in insertion order
    uint32_t kept = 0;
    uint32_t prevTree = 0xffffffffu;
    float prevT0 = 0.0f, prevT1 = 0.0f;
    for (uint32_t k = 0; k < c; ++k)
    {
        const LeafHit h = seg[k];
        if (h.order == prevTree && h.t0 == prevT0 && h.t1 == prevT1)
            continue; // equal (entry, exit) key in the same tree's std::set: not inserted
        prevTree = h.order;
        prevT0 = h.t0;
        prevT1 = h.t1;
        const uint32_t meta = s.meta[h.shape];
        if (!(meta & META_ACTIVE))
            continue; // Registry.hpp:1068-1069 — after the set was built
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
        float a = 0.0f, b = 0.0f;
        bool hit;
        if (h.order == 2u)
        {
            const float4 p = s.pxf[h.shape], pa = s.pxa[h.shape], ca = s.xa[h.shape];
            XF prev;
            prev.tx = p.x;
            prev.ty = p.y;
            prev.s = p.z;
            prev.c = p.w;
            hit = ray_sweep(S, motion_between(prev, pa.x, pa.y, pa.z, cur, ca.x), q, second, a, b);
        }
        else
            hit = ray_shape(S, cur, q, second, a, b);
        if (hit)
        {
            seg[kept] = LeafHit{a, b, h.shape, kept};
            ++kept;
        }
    }
    // final std::set<RaycastHit>: ordered by (entryTime, exitTime); an equal key keeps the first inserted
    heapsort(seg, static_cast<int>(kept), ByTimeInsertion());
    uint32_t out = 0;
    for (uint32_t k = 0; k < kept; ++k)
    {
        const LeafHit h = seg[k];
        if (out && seg[out - 1].t0 == h.t0 && seg[out - 1].t1 == h.t1)
            continue;
        seg[out++] = h;
    }
    finalCounts[i] = out;
}

__global__ void __launch_bounds__(128)
    k_ray_emit(const RayIn* __restrict__ rays, uint32_t n, const uint32_t* __restrict__ segOffsets,
               const LeafHit* __restrict__ segs, const uint32_t* __restrict__ outOffsets, uint32_t outCount,
               const uint32_t* __restrict__ entity, c2d_hit_record* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t begin = outOffsets[i];
    const uint32_t end = (i + 1 < n) ? outOffsets[i + 1] : outCount;
    V2 second;
    const RayQ q = make_ray(rays[i], second);
    const LeafHit* seg = segs + segOffsets[i];
    for (uint32_t k = begin; k < end; ++k)
    {
        const LeafHit h = seg[k - begin];
        // RaycastHit::fromRay (Ray.hpp:34-65): start + (end - start) * t  |  start + direction * t
        const V2 en = q.start + q.d * h.t0;
        const V2 ex = q.start + q.d * h.t1;
        c2d_hit_record rec;
        rec.entity = entity[h.shape];
        rec.shape = h.shape;
        rec.entry[0] = en.x;
        rec.entry[1] = en.y;
        rec.exit[0] = ex.x;
        rec.exit[1] = ex.y;
        rec.entryTime = h.t0;
        rec.exitTime = h.t1;
        out[k] = rec;
    }
}

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
// Small batches (the reference's one-ray-per-call rayCast / firstHitRayCast): ONE fused kernel, one 256-thread
// block per ray.  The thread-per-ray pipeline above is a throughput design — a single long ray would be walked,
// sorted and narrow-phased by one thread (milliseconds at 1 M shapes).  Here the block pops up to 256 nodes of a
// shared stack per step, collects the leaf-box hits in shared memory, bitonic-sorts them, runs the narrowphase one
// hit per thread and bitonic-sorts the survivors; the std::set semantics are the ones described at the top of the
// file.  A ray with more than kSmallCap leaf-box hits (or a frontier above kSmallStack) sets an overflow flag and the
// caller falls back to the general pipeline.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSmallThreads = 256;
constexpr int kSmallCap = 2048;   // leaf-box hits per ray held in shared memory (32 KB)
constexpr int kSmallStack = 2048; // traversal frontier (8 KB)

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

__global__ void __launch_bounds__(kSmallThreads)
    k_ray_small(const RayIn* __restrict__ rays, uint32_t n, TreeView t0, TreeView t1, TreeView t2, ShapeArrays s,
                float gridCell, bool firstHitOnly, SmallHeader* __restrict__ headers, uint32_t* __restrict__ cursor,
                c2d_hit_record* __restrict__ out, uint32_t outCapacity)
{
    __shared__ LeafHit hits[kSmallCap];
    __shared__ int stack[kSmallStack];
    __shared__ int stackSize, hitCount, keptCount, overflow, outBase;
    const uint32_t rayIdx = blockIdx.x;
    if (rayIdx >= n)
        return;
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
        const int take = size < kSmallThreads ? size : kSmallThreads;
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
                    if (at < kSmallCap)
                        hits[at] = LeafHit{a0, a1, trees[tree].sorted[~N.left], tree};
                    else
                        overflow = 1;
                }
                else
                {
                    const int at = atomicAdd(&stackSize, 1);
                    if (at < kSmallStack)
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
                    if (at < kSmallCap)
                        hits[at] = LeafHit{a0, a1, trees[tree].sorted[~N.right], tree};
                    else
                        overflow = 1;
                }
                else
                {
                    const int at = atomicAdd(&stackSize, 1);
                    if (at < kSmallStack)
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
    LeafHit mine[kSmallCap / kSmallThreads];
    bool keep[kSmallCap / kSmallThreads];
#pragma unroll
    for (int u = 0; u < kSmallCap / kSmallThreads; ++u)
    {
        const int i = threadIdx.x + u * kSmallThreads;
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
    for (int u = 0; u < kSmallCap / kSmallThreads; ++u)
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
    __shared__ int scan[kSmallThreads];
    int local = 0;
    constexpr int kPer = kSmallCap / kSmallThreads;
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
    for (int o = 1; o < kSmallThreads; o <<= 1)
    {
        const int v = threadIdx.x >= o ? scan[threadIdx.x - o] : 0;
        __syncthreads();
        scan[threadIdx.x] += v;
        __syncthreads();
    }
    const int total = scan[kSmallThreads - 1];
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

static int raycast_impl(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, bool firstHitOnly,
                        const c2d_hit_record** out_hits, const uint64_t** out_offsets, c2d_gpu_stats* stats);

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

static int raycast_impl(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, bool firstHitOnly,
                        const c2d_hit_record** out_hits, const uint64_t** out_offsets, c2d_gpu_stats* stats)
{
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
    rb.hostHits.size = 0;
    uint64_t totalHits = 0, d2h = 0, leafHits = 0;

    const TreeView t0 = view_of(ctx, 0), t1 = view_of(ctx, 1), t2 = view_of(ctx, 2);
    const float gridCell = ctx->broadphase == C2D_BROADPHASE_GRID ? static_cast<float>(static_cast<int32_t>(ctx->cellSize)) : 0.0f;

    // ---- small batches: one fused kernel, one block per ray (falls through to the general path on overflow)
    constexpr uint32_t kSmallBatch = 64;
    if (n > 0 && n <= kSmallBatch)
    {
        const uint32_t outCap = kSmallBatch * kSmallCap;
        C2D_CUDA(ctx, rb.dRays.reserve(kSmallBatch));
        C2D_CUDA(ctx, rb.smallHeaders.reserve(kSmallBatch * sizeof(SmallHeader) + 16));
        C2D_CUDA(ctx, rb.out.reserve(outCap));
        if (!rb.hostSmall.reserve(kSmallBatch * sizeof(SmallHeader) + 16) || !rb.hostHits.reserve(outCap))
            return fail(ctx, "pinned host allocation failed");
        SmallHeader* dHeaders = reinterpret_cast<SmallHeader*>(rb.smallHeaders.ptr + 16);
        uint32_t* dCursor = reinterpret_cast<uint32_t*>(rb.smallHeaders.ptr);
        C2D_CUDA(ctx, cudaMemsetAsync(dCursor, 0, 16, s));
        C2D_CUDA(ctx, cudaMemcpyAsync(rb.dRays.ptr, rays, static_cast<size_t>(n) * sizeof(c2d_ray), cudaMemcpyHostToDevice, s));
        k_ray_small<<<n, kSmallThreads, 0, s>>>(reinterpret_cast<const RayIn*>(rb.dRays.ptr), n, t0, t1, t2, ctx->d, gridCell,
                                                firstHitOnly, dHeaders, dCursor, rb.out.ptr, outCap);
        ++ctx->launches;
        // header block + the first records in ONE copy; a second copy only if there are more
        constexpr uint32_t kEager = 64;
        C2D_CUDA(ctx, cudaMemcpyAsync(rb.hostSmall.data, rb.smallHeaders.ptr, 16 + static_cast<size_t>(n) * sizeof(SmallHeader),
                                      cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaMemcpyAsync(rb.hostHits.data, rb.out.ptr, kEager * sizeof(c2d_hit_record), cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
        C2D_CUDA(ctx, cudaGetLastError());
        const uint32_t produced = *reinterpret_cast<const uint32_t*>(rb.hostSmall.data);
        const SmallHeader* hh = reinterpret_cast<const SmallHeader*>(rb.hostSmall.data + 16);
        bool overflowed = produced > outCap;
        for (uint32_t i = 0; i < n; ++i)
            overflowed = overflowed || hh[i].overflow != 0;
        if (!overflowed)
        {
            if (produced > kEager)
            {
                C2D_CUDA(ctx, cudaMemcpyAsync(rb.hostHits.data + kEager, rb.out.ptr + kEager,
                                              static_cast<size_t>(produced - kEager) * sizeof(c2d_hit_record),
                                              cudaMemcpyDeviceToHost, s));
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
            if (out_hits)
                *out_hits = rb.hostHits.data;
            if (out_offsets)
                *out_offsets = rb.hostOffsets.data;
            if (stats)
            {
                std::memset(stats, 0, sizeof(*stats));
                stats->shapes_alive = ctx->members[0].size() + ctx->members[1].size() + ctx->members[2].size();
                stats->candidates = leafHits;
                stats->pairs = totalHits;
                stats->h2d_bytes = static_cast<uint64_t>(n) * sizeof(c2d_ray);
                stats->d2h_bytes = 16 + static_cast<uint64_t>(n) * sizeof(SmallHeader) + std::max<uint64_t>(produced, kEager) * sizeof(c2d_hit_record);
                stats->kernel_launches = ctx->launches;
                stats->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
            }
            return 0;
        }
        totalHits = 0;
        leafHits = 0;
    }
    const uint32_t kChunk = 1u << 20;
    for (uint32_t first = 0; first < n; first += kChunk)
    {
        const uint32_t m = std::min(kChunk, n - first);
        C2D_CUDA(ctx, rb.dRays.reserve(m));
        C2D_CUDA(ctx, rb.counts.reserve(m));
        C2D_CUDA(ctx, rb.finalCounts.reserve(m));
        C2D_CUDA(ctx, rb.totals.reserve(2));
        C2D_CUDA(ctx, cudaMemcpyAsync(rb.dRays.ptr, rays + first, static_cast<size_t>(m) * sizeof(c2d_ray),
                                      cudaMemcpyHostToDevice, s));
        ctx->h2dBytes += static_cast<uint64_t>(m) * sizeof(c2d_ray);
        const RayIn* dr = reinterpret_cast<const RayIn*>(rb.dRays.ptr);
        k_ray_count<<<blocks_for(m, 128), 128, 0, s>>>(dr, m, t0, t1, t2, ctx->d.fat, ctx->d.category, rb.counts.ptr, gridCell);
        ++ctx->launches;
        if (int rc = exclusive_scan_u32(ctx, rb.counts.ptr, m, rb.sums, rb.totals.ptr))
            return rc;
        uint32_t segTotal = 0;
        C2D_CUDA(ctx, cudaMemcpyAsync(&segTotal, rb.totals.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
        leafHits += segTotal;
        C2D_CUDA(ctx, rb.segs.reserve(std::max<size_t>(segTotal, 1) * sizeof(LeafHit)));
        LeafHit* segs = reinterpret_cast<LeafHit*>(rb.segs.ptr);
        k_ray_fill<<<blocks_for(m, 128), 128, 0, s>>>(dr, m, t0, t1, t2, ctx->d, rb.counts.ptr, segs, rb.finalCounts.ptr, gridCell, firstHitOnly);
        ++ctx->launches;
        if (int rc = exclusive_scan_u32(ctx, rb.finalCounts.ptr, m, rb.sums, rb.totals.ptr + 1))
            return rc;
        uint32_t outTotal = 0;
        C2D_CUDA(ctx, cudaMemcpyAsync(&outTotal, rb.totals.ptr + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
        C2D_CUDA(ctx, rb.out.reserve(std::max<uint32_t>(outTotal, 1)));
        k_ray_emit<<<blocks_for(m, 128), 128, 0, s>>>(dr, m, rb.counts.ptr, segs, rb.finalCounts.ptr, outTotal,
                                                     ctx->d.entity, rb.out.ptr);
        ++ctx->launches;
        C2D_CUDA(ctx, cudaGetLastError());
        // download this chunk: offsets (as uint32, widened below) and records
        if (!rb.hostHits.reserve(totalHits + outTotal) || !rb.hostOffsets32.reserve(m))
            return fail(ctx, "pinned host allocation failed");
        C2D_CUDA(ctx, cudaMemcpyAsync(rb.hostOffsets32.data, rb.finalCounts.ptr, static_cast<size_t>(m) * sizeof(uint32_t),
                                      cudaMemcpyDeviceToHost, s));
        if (outTotal)
            C2D_CUDA(ctx, cudaMemcpyAsync(rb.hostHits.data + totalHits, rb.out.ptr,
                                          static_cast<size_t>(outTotal) * sizeof(c2d_hit_record), cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
        d2h += static_cast<uint64_t>(m) * sizeof(uint32_t) + static_cast<uint64_t>(outTotal) * sizeof(c2d_hit_record);
        for (uint32_t i = 0; i < m; ++i)
            rb.hostOffsets.data[first + i] = totalHits + rb.hostOffsets32.data[i];
        totalHits += outTotal;
        rb.hostHits.size = totalHits;
    }
    rb.hostOffsets.data[n] = totalHits;
    C2D_CUDA(ctx, cudaEventRecord(ctx->ev[4], s));
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    if (out_hits)
        *out_hits = rb.hostHits.data;
    if (out_offsets)
        *out_offsets = rb.hostOffsets.data;
    if (stats)
    {
        std::memset(stats, 0, sizeof(*stats));
        stats->shapes_alive = ctx->members[0].size() + ctx->members[1].size() + ctx->members[2].size();
        stats->candidates = leafHits;
        stats->pairs = totalHits;
        stats->h2d_bytes = ctx->h2dBytes;
        stats->d2h_bytes = d2h;
        stats->kernel_launches = ctx->launches;
        cudaEventElapsedTime(&stats->ms_device, ctx->ev[0], ctx->ev[4]);
        stats->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    }
    return 0;
}
