// c2d_kernels.cuh — the sm_100a kernels of the getCollidingPairs pipeline.
//
//   K1  k_apply_flags / k_apply_adds / k_apply_translates / k_apply_moves   (mutation replay = AABB refit)
//   K2  k_bounds / k_morton                                                (scene bounds, Morton keys)
//   K3  radix sort                                                         (c2d_sort.cu)
//   K4  k_karras_build / k_fit                                             (LBVH topology, bottom-up boxes)
//   K5  k_traverse<SELF> / k_traverse_warp<SELF>                           (candidate pairs; warp-per-query for bullets)
//   K6  k_bin_count / k_bin_scatter / k_narrow_filter(_pp) / k_narrow_manifold(_pp) (binning by code path, boolean tests and
//       sweep searches, manifolds; the polygon bin runs 8 lanes per pair, c2d_narrow_pp.cuh)
// The uniform-grid broadphase is in c2d_grid.cu, rays in c2d_rays.cu, per-entity queries in c2d_query.cu.
//
// All floating point on the parity path is binary32 without FMA contraction (-fmad=false).
#pragma once

#include "c2d_internal.hpp"
#include "c2d_narrow.cuh"
#include "c2d_narrow_pp.cuh"

#include <cooperative_groups.h>

namespace c2d_dev
{

namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------------------------------
// K1 — mutation replay.  One thread per log record; records of one launch touch distinct shapes.
// ---------------------------------------------------------------------------------------------------

// DynamicBVH::moveProxy leaf rule with margin 0 (reference DynamicBVH.hpp:386-400, AABB.cpp:212-228):
// keep the fat box if it still contains the target, else target grown by 2.2 x (target.min - fat.min)
// in the direction of motion.
__device__ __forceinline__ bool fat_leaf_rule(float4* fatSlot, const Box& target) // true: the leaf box changed
{
    const Box fat = box_from(*fatSlot);
    if (box_contains(fat, target))
        return false;
    const float dispx = (target.minx - fat.minx) * 2.2f;
    const float dispy = (target.miny - fat.miny) * 2.2f;
    Box nf = target; // fattened(margin = 0)
    if (dispx < 0.0f)
        nf.minx += dispx;
    else
        nf.maxx += dispx;
    if (dispy < 0.0f)
        nf.miny += dispy;
    else
        nf.maxy += dispy;
    *fatSlot = box_to(nf);
    return true;
}

__global__ void __launch_bounds__(256) k_apply_flags(const FlagRec* __restrict__ log, uint32_t n, ShapeArrays s)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const FlagRec r = log[i];
    uint32_t m = s.meta[r.shape];
    if (r.op == 0)
        m &= ~META_ALIVE;
    else if (r.op == 1)
        m |= META_ACTIVE;
    else if (r.op == 2)
        m &= ~META_ACTIVE;
    else if (r.op == 3)
        m |= META_GHOST;
    else
        m &= ~META_GHOST;
    s.meta[r.shape] = m;
}

// Registry::addShape (reference Registry.hpp:272-330) + DynamicBVH::addProxy (DynamicBVH.hpp:352-365)
__global__ void __launch_bounds__(128) k_apply_adds(const AddRec* __restrict__ log, uint32_t n, ShapeArrays s)
{
    // one warp per record: the 224-byte record is copied with coalesced loads, lane 0 finishes the boxes
    const uint32_t warpId = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warpId >= n)
        return;
    const AddRec* r = log + warpId;
    const uint32_t shape = r->shape;
    s.geom[static_cast<size_t>(shape) * 32 + lane] = r->geom[lane];
    if (lane == 0)
    {
        const uint32_t meta = r->meta | META_TIGHT_EXACT;
        s.meta[shape] = meta;
        s.entity[shape] = r->entity;
        s.category[shape] = r->category;
        s.mask[shape] = r->mask;
        XF t;
        t.tx = r->xf[0];
        t.ty = r->xf[1];
        t.s = r->xf[3];
        t.c = r->xf[4];
        s.xf[shape] = make_float4(t.tx, t.ty, t.s, t.c);
        s.xa[shape] = make_float4(r->xf[2], r->xf[5], r->xf[6], 0.0f);
        s.pxf[shape] = make_float4(r->prev[0], r->prev[1], r->prev[3], r->prev[4]);
        s.pxa[shape] = make_float4(r->prev[2], r->prev[5], r->prev[6], 0.0f);
        const Box b = compute_aabb(meta_type(meta), meta_count(meta), r->geom, t);
        s.tight[shape] = box_to(b);
        s.fat[shape] = box_to(b); // boundingBox.fattened(mFatAABBMargin = 0)
    }
}

// Registry::moveEntity(id, Translation) (reference Registry.hpp:434-453, 214-231)
__device__ __forceinline__ void apply_translate(const TransRec rec, const ShapeArrays& s)
{
    const uint32_t shape = rec.shape;
    const float dx = rec.dx, dy = rec.dy;
    const uint32_t meta = s.meta[shape];
    const bool bullet = meta_body(meta) == 2u;

    const Box old = box_from(s.tight[shape]);
    Box nt; // AABB::translate: min += d; max += d (AABB.cpp:248-252)
    nt.minx = old.minx + dx;
    nt.miny = old.miny + dy;
    nt.maxx = old.maxx + dx;
    nt.maxy = old.maxy + dy;
    s.tight[shape] = box_to(nt);
    const bool leafChanged = fat_leaf_rule(&s.fat[shape], bullet ? box_combine(old, nt) : nt);
    // the box is shifted from now on, not recomputed (rounding drift); a changed leaf box is flagged for the tree's refit
    const uint32_t after = (meta & ~META_TIGHT_EXACT) | (leafChanged ? META_FAT_DIRTY : 0u);
    if (after != meta)
        s.meta[shape] = after;

    float4 t = s.xf[shape];
    if (bullet)
    {
        s.pxf[shape] = t; // mBulletPreviousTransforms[id] = entity.mTransform
        s.pxa[shape] = s.xa[shape];
    }
    t.x += dx;
    t.y += dy;
    s.xf[shape] = t;
}

__global__ void __launch_bounds__(256) k_apply_translates(const TransRec* __restrict__ log, uint32_t n, ShapeArrays s)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    apply_translate(log[i], s); // 12-byte records: three coalesced 4-byte loads per thread
}

// The same over the all-gathered records of a multi-GPU group (c2d_gpu.cu: flush with the move exchange on):
// `world` blocks of `stride` records, block r holding header[4 * r] valid ones.
// overflow (may be null): the one-step exchange sized every rank's block from the PREVIOUS frame's counts; if a rank logged
// more records than fit, nothing is applied anywhere, the flag goes up and the host repeats the exchange with exact sizes.
__global__ void __launch_bounds__(256) k_apply_translates_gathered(const TransRec* __restrict__ gathered, uint32_t stride,
                                                                   const uint32_t* __restrict__ header, uint32_t world,
                                                                   ShapeArrays s, uint32_t* overflow)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (overflow)
    {
        bool any = false;
        for (uint32_t q = 0; q < world; ++q)
            any = any || header[4 * q] > stride;
        if (any)
        {
            if (i == 0)
                *overflow = 1;
            return;
        }
    }
    const uint32_t r = i / stride, k = i - r * stride;
    if (r >= world || k >= header[4 * r])
        return;
    apply_translate(gathered[i], s);
}

// Registry::moveEntity(id, Transform) / teleportEntity(id, Transform) (reference Registry.hpp:382-432, 191-212)
__global__ void __launch_bounds__(256) k_apply_moves(const MoveRec* __restrict__ log, uint32_t n, ShapeArrays s)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float4 r0 = reinterpret_cast<const float4*>(log)[2 * i];
    const float4 r1 = reinterpret_cast<const float4*>(log)[2 * i + 1];
    const uint32_t shape = __float_as_uint(r0.x);
    XF t;
    t.tx = r0.y;
    t.ty = r0.z;
    t.s = r1.x;
    t.c = r1.y;
    const uint32_t meta = s.meta[shape];
    const bool bullet = meta_body(meta) == 2u;

    const Box old = box_from(s.tight[shape]);
    const Box nt = compute_aabb(meta_type(meta), meta_count(meta), s.geom + static_cast<size_t>(shape) * 32, t);
    s.tight[shape] = box_to(nt);
    const bool leafChanged = fat_leaf_rule(&s.fat[shape], bullet ? box_combine(old, nt) : nt);
    const uint32_t after = meta | META_TIGHT_EXACT | (leafChanged ? META_FAT_DIRTY : 0u);
    if (after != meta)
        s.meta[shape] = after;
    if (bullet)
    {
        s.pxf[shape] = s.xf[shape];
        s.pxa[shape] = s.xa[shape];
    }
    s.xf[shape] = make_float4(t.tx, t.ty, t.s, t.c);
    s.xa[shape] = make_float4(r0.w, r1.z, r1.w, 0.0f);
}

// c2d_gpu_write_boxes: restored boxes are history, not computeAABB of the current transform
__global__ void __launch_bounds__(256) k_clear_meta_bits(uint32_t* __restrict__ meta, uint32_t n, uint32_t bits)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        meta[i] &= ~bits;
}

// ---------------------------------------------------------------------------------------------------
// K2 — bounds of the leaf-box centres and Morton keys
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_to_ordered(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_frame_init(FrameScratch* fs)
{
    const int t = threadIdx.x;
    if (t < 12)
        fs->bounds[t / 4][t % 4] = (t % 4) < 2 ? 0xffffffffu : 0u;
    if (t == 12)
        fs->candCount = 0;
    if (t == 13)
        fs->outCount = 0;
    if (t == 14)
    {
        fs->clipOverflow = 0;
        fs->stackOverflow = 0;
    }
    if (t >= 16 && t < 32)
    {
        fs->binCount[t - 16] = 0;
        fs->binOffset[t - 16] = 0;
        fs->binCursor[t - 16] = 0;
        fs->hitCount[t - 16] = 0;
        fs->hitBase[t - 16] = 0;
    }
    if (t == 15)
    {
        fs->hitBase[16] = 0;
        fs->ticket = 0;
    }
}

__global__ void __launch_bounds__(256) k_bounds(const uint32_t* __restrict__ members, uint32_t n,
                                                const float4* __restrict__ fat, uint32_t* bounds4)
{
    float minx = FLT_MAX, miny = FLT_MAX, maxx = -FLT_MAX, maxy = -FLT_MAX;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 b = fat[members[i]];
        const float cx = 0.5f * (b.x + b.z), cy = 0.5f * (b.y + b.w);
        minx = fminf(minx, cx);
        miny = fminf(miny, cy);
        maxx = fmaxf(maxx, cx);
        maxy = fmaxf(maxy, cy);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        minx = fminf(minx, __shfl_xor_sync(0xffffffffu, minx, o));
        miny = fminf(miny, __shfl_xor_sync(0xffffffffu, miny, o));
        maxx = fmaxf(maxx, __shfl_xor_sync(0xffffffffu, maxx, o));
        maxy = fmaxf(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
    }
    if ((threadIdx.x & 31) == 0 && minx <= maxx)
    {
        atomicMin(&bounds4[0], float_to_ordered(minx));
        atomicMin(&bounds4[1], float_to_ordered(miny));
        atomicMax(&bounds4[2], float_to_ordered(maxx));
        atomicMax(&bounds4[3], float_to_ordered(maxy));
    }
}

__device__ __forceinline__ uint32_t spread16(uint32_t v)
{
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

__global__ void __launch_bounds__(256) k_morton(const uint32_t* __restrict__ members, uint32_t n,
                                                const float4* __restrict__ fat, const uint32_t* __restrict__ bounds4,
                                                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float minx = ordered_to_float(bounds4[0]), miny = ordered_to_float(bounds4[1]);
    const float maxx = ordered_to_float(bounds4[2]), maxy = ordered_to_float(bounds4[3]);
    const float ex = maxx - minx, ey = maxy - miny;
    const float sx = ex > 0.0f ? 65535.0f / ex : 0.0f;
    const float sy = ey > 0.0f ? 65535.0f / ey : 0.0f;
    const uint32_t shape = members[i];
    const float4 b = fat[shape];
    const float cx = 0.5f * (b.x + b.z), cy = 0.5f * (b.y + b.w);
    const uint32_t qx = static_cast<uint32_t>(fminf(fmaxf((cx - minx) * sx, 0.0f), 65535.0f));
    const uint32_t qy = static_cast<uint32_t>(fminf(fmaxf((cy - miny) * sy, 0.0f), 65535.0f));
    keys[i] = spread16(qx) | (spread16(qy) << 1);
    vals[i] = shape;
}

// ---------------------------------------------------------------------------------------------------
// K4 — LBVH topology (Karras 2012) and bottom-up fit
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int karras_delta(const uint32_t* __restrict__ keys, int n, int i, uint32_t ki, int j)
{
    if (j < 0 || j >= n)
        return -1;
    const uint32_t kj = keys[j];
    if (ki == kj)
        return 32 + __clz(static_cast<uint32_t>(i) ^ static_cast<uint32_t>(j));
    return __clz(ki ^ kj);
}

// parent code: (parent index << 1) | (1 if right child)
__global__ void __launch_bounds__(256) k_karras_build(const uint32_t* __restrict__ keys,
                                                      const uint32_t* __restrict__ sorted, int n, BvhNode* nodes,
                                                      int* __restrict__ nodeParent, int* __restrict__ leafParent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1)
        return;
    const uint32_t ki = keys[i];
    const int d = (karras_delta(keys, n, i, ki, i + 1) - karras_delta(keys, n, i, ki, i - 1)) >= 0 ? 1 : -1;
    const int dmin = karras_delta(keys, n, i, ki, i - d);
    int lmax = 2;
    while (karras_delta(keys, n, i, ki, i + lmax * d) > dmin)
        lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (karras_delta(keys, n, i, ki, i + (l + t) * d) > dmin)
            l += t;
    const int j = i + l * d;
    const int dnode = karras_delta(keys, n, i, ki, j);
    int s = 0;
    int t = l;
    do
    {
        t = (t + 1) >> 1;
        if (karras_delta(keys, n, i, ki, i + (s + t) * d) > dnode)
            s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);

    int left, right;
    if (lo == gamma)
    {
        left = static_cast<int>(kLeafBit | sorted[gamma]);
        leafParent[gamma] = (i << 1);
    }
    else
    {
        left = gamma;
        nodeParent[gamma] = (i << 1);
    }
    if (hi == gamma + 1)
    {
        right = static_cast<int>(kLeafBit | sorted[gamma + 1]);
        leafParent[gamma + 1] = (i << 1) | 1;
    }
    else
    {
        right = gamma + 1;
        nodeParent[gamma + 1] = (i << 1) | 1;
    }
    nodes[i].left = left;
    nodes[i].right = right;
    nodes[i].lastL = gamma;
    nodes[i].lastR = hi;
    if (i == 0)
        nodeParent[0] = -1;
}

__global__ void __launch_bounds__(256) k_fit(int n, const uint32_t* __restrict__ sorted, const float4* __restrict__ fat,
                                             const uint64_t* __restrict__ category, BvhNode* nodes,
                                             const int* __restrict__ nodeParent, const int* __restrict__ leafParent,
                                             uint32_t* arrival)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    const uint32_t shape = sorted[p] & kLeafIdMask;
    float4 box = fat[shape];
    uint64_t cat = category[shape];
    int code = leafParent[p];
    while (true)
    {
        const int q = code >> 1;
        const bool right = code & 1;
        BvhNode* node = nodes + q;
        if (right)
        {
            node->boxR = box;
            node->catR = cat;
        }
        else
        {
            node->boxL = box;
            node->catL = cat;
        }
        __threadfence();
        const uint32_t old = atomicAdd(&arrival[q], 1u);
        if (old == 0)
            return; // the sibling subtree is not finished: its last thread continues upward
        // the sibling's box was published before its own fence + atomic; ld.cg reads it from L2 (coherence point)
        const float4 sib = __ldcg(right ? &node->boxL : &node->boxR);
        const uint64_t sibCat = __ldcg(reinterpret_cast<const unsigned long long*>(right ? &node->catL : &node->catR));
        box.x = fminf(box.x, sib.x);
        box.y = fminf(box.y, sib.y);
        box.z = fmaxf(box.z, sib.z);
        box.w = fmaxf(box.w, sib.w);
        cat |= sibCat;
        if (q == 0)
            return;
        code = nodeParent[q];
    }
}

// Refit of a tree whose topology is kept (frames between two rebuilds) when only translations moved its leaves: most leaf
// boxes did not change at all (the fat-leaf rule keeps a box while it still contains the shape), so instead of the
// full bottom-up pass - a chain of ~log n dependent fence + atomic steps per leaf - only the leaves flagged
// META_FAT_DIRTY act: the leaf's slot in its parent gets the exact new box, and every ancestor slot on the way up is
// GROWN to contain it (component-wise atomic min / max), stopping at the first ancestor that already does.  Leaf slots
// stay exact, so the candidate set is unchanged; internal boxes are conservative until the next rebuild (at most
// rebuildPeriod - 1 frames later) recomputes them.  Why stopping early is safe: every bound a thread relies on was put
// there by the last full fit (where parent slots contain child slots) or by a thread that lowered / raised that very
// component and carries the same value further up.
__device__ __forceinline__ void atomic_min_float(float* addr, float v)
{
    if (v >= 0.0f)
        atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v)
{
    if (v >= 0.0f)
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256) k_refit_dirty(int n, const uint32_t* __restrict__ sorted, const float4* __restrict__ fat,
                                                     uint32_t* __restrict__ meta, BvhNode* nodes,
                                                     const int* __restrict__ nodeParent, const int* __restrict__ leafParent)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    const uint32_t shape = sorted[p] & kLeafIdMask;
    const uint32_t m = meta[shape];
    if (!(m & META_FAT_DIRTY))
        return;
    meta[shape] = m & ~META_FAT_DIRTY;
    float4 box = fat[shape];
    // -0.0f would order below +0.0f in the integer comparisons of the atomics: canonicalise
    box.x += 0.0f, box.y += 0.0f, box.z += 0.0f, box.w += 0.0f;
    int code = leafParent[p];
    int q = code >> 1;
    float4* slot = (code & 1) ? &nodes[q].boxR : &nodes[q].boxL;
    *slot = fat[shape]; // the leaf's own slot: exact, single writer
    while (q != 0)
    {
        code = nodeParent[q];
        q = code >> 1;
        slot = (code & 1) ? &nodes[q].boxR : &nodes[q].boxL;
        const float4 cur = __ldcg(slot);
        const bool gx = box.x < cur.x, gy = box.y < cur.y, gz = box.z > cur.z, gw = box.w > cur.w;
        if (!(gx || gy || gz || gw))
            return;
        if (gx)
            atomic_min_float(&slot->x, box.x);
        if (gy)
            atomic_min_float(&slot->y, box.y);
        if (gz)
            atomic_max_float(&slot->z, box.z);
        if (gw)
            atomic_max_float(&slot->w, box.w);
    }
}

// ---------------------------------------------------------------------------------------------------
// K5 — traversal: one query leaf per thread, candidate pairs appended with warp-aggregated atomics.
// Replaces DynamicBVH::findAllCollisions self / cross (reference DynamicBVH.hpp:931-1031): a pair is a
// candidate iff (categoryA & categoryB) != 0 (matchesMask, DynamicBVH.hpp:50-53,955) and the two LEAF
// (fat) boxes intersect, all four comparisons non-strict (AABB.cpp:34-37).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_pair(uint32_t a, uint32_t b, uint2* __restrict__ out, uint32_t capacity,
                                          uint32_t* counter)
{
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0)
        base = atomicAdd(counter, g.size());
    base = g.shfl(base, 0);
    const uint32_t idx = base + g.thread_rank();
    if (idx < capacity)
        out[idx] = make_uint2(a, b);
}

template <bool SELF>
__global__ void __launch_bounds__(128)
    k_traverse(const uint32_t* __restrict__ qSorted, uint32_t qn, uint32_t qLo, uint32_t qHi,
               const uint32_t* __restrict__ tSorted, uint32_t tn, const BvhNode* __restrict__ nodes,
               const float4* __restrict__ fat, const uint64_t* __restrict__ category, uint2* __restrict__ out,
               uint32_t capacity, uint32_t* counter, uint32_t* stackOverflow)
{
    const uint32_t i = qLo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= qHi || i >= qn || tn == 0)
        return;
    const uint32_t qs = qSorted[i] & kLeafIdMask;
    const float4 qb = fat[qs];
    const uint64_t qc = category[qs];

    if (tn == 1)
    {
        if (!SELF)
        {
            const uint32_t ts = tSorted[0] & kLeafIdMask;
            if ((qc & category[ts]) != 0 && box_overlap(qb, fat[ts]))
                emit_pair(qs, ts, out, capacity, counter);
        }
        return;
    }

    int stack[64];
    int sp = 0;
    int node = 0;
    while (true)
    {
        const float4* np = reinterpret_cast<const float4*>(nodes + node);
        const float4 boxL = __ldg(np);
        const float4 boxR = __ldg(np + 1);
        const int4 links = __ldg(reinterpret_cast<const int4*>(np + 2));
        const ulonglong2 cats = __ldg(reinterpret_cast<const ulonglong2*>(np + 3));
        bool hitL = box_overlap(qb, boxL) && (qc & cats.x) != 0 && (!SELF || links.z > static_cast<int>(i));
        bool hitR = box_overlap(qb, boxR) && (qc & cats.y) != 0 && (!SELF || links.w > static_cast<int>(i));
        if (hitL && links.x < 0)
        {
            const uint32_t ts = leaf_shape(links.x);
            if (SELF)
                emit_pair(min(qs, ts), max(qs, ts), out, capacity, counter);
            else
                emit_pair(qs, ts, out, capacity, counter);
            hitL = false;
        }
        if (hitR && links.y < 0)
        {
            const uint32_t ts = leaf_shape(links.y);
            if (SELF)
                emit_pair(min(qs, ts), max(qs, ts), out, capacity, counter);
            else
                emit_pair(qs, ts, out, capacity, counter);
            hitR = false;
        }
        if (hitL)
        {
            if (hitR)
            {
                if (sp < 64)
                    stack[sp++] = links.y;
                else
                    *stackOverflow = 1; // cannot happen (Karras depth <= 64 with the index tie-break); never silent
            }
            node = links.x;
        }
        else if (hitR)
        {
            node = links.y;
        }
        else
        {
            if (sp == 0)
                break;
            node = stack[--sp];
        }
    }
}

// k_traverse_sync: the same traversal with a warp-CONVERGENT loop.  Every iteration each live lane visits one node; the
// warp reconverges before the candidates of the iteration are appended, so
//   * the two emit sites of k_traverse (each a divergent region with its own coalesced-group election and global atomic)
//     become two full-mask ballots,
//   * candidates are staged in a per-warp shared-memory buffer and flushed 32 at a time: one global atomic and one
//     coalesced 256-byte store per 32 candidates instead of one contended same-address atomic per (partial) warp emit.
// rejectLeaf: leaf links with any of these bits are skipped.  No tree of the current decomposition flags its leaves (the
// first slab trees of round 2 carried their halo leaves, bit 30); every launch passes 0.
constexpr int kStageCap = 96; // < 32 staged + <= 64 new entries per iteration

__device__ __forceinline__ void stage_flush32(uint2* stage, int& staged, uint2* __restrict__ out, uint32_t capacity,
                                              uint32_t* counter, int lane)
{
    // warp-convergent; staged >= 32: the OLDEST 32 entries go out, the rest slides down
    uint32_t base = 0;
    if (lane == 0)
        base = atomicAdd(counter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    const uint2 mine = stage[lane];
    if (base + lane < capacity)
        out[base + lane] = mine;
    const int rest = staged - 32;
    // slide: rest < 64 entries, two per lane
    const uint2 a = lane < rest ? stage[32 + lane] : make_uint2(0, 0);
    const uint2 b = 32 + lane < rest ? stage[64 + lane] : make_uint2(0, 0);
    __syncwarp();
    if (lane < rest)
        stage[lane] = a;
    if (32 + lane < rest)
        stage[32 + lane] = b;
    __syncwarp();
    staged = rest;
}

// ORD: emit (min, max) shape ids although the query is not a leaf of the tree (halo shapes of a multi-GPU slab querying the
// rank's own Dynamic leaves: a same-tree pair of the reference).  qCount (may be null): device-side length of the query
// list, so a launch sized by an upper bound needs no host round trip.
template <bool SELF, bool ORD = SELF>
__global__ void __launch_bounds__(128)
    k_traverse_sync(const uint32_t* __restrict__ qSorted, uint32_t qn, uint32_t qLo, uint32_t qHi,
                    const uint32_t* __restrict__ qCount, uint32_t tn, const BvhNode* __restrict__ nodes,
                    const float4* __restrict__ fat, const uint64_t* __restrict__ category, uint32_t rejectLeaf,
                    uint2* __restrict__ out, uint32_t capacity, uint32_t* counter, uint32_t* stackOverflow)
{
    __shared__ uint2 stages[4][kStageCap];
    uint2* stage = stages[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t i = qLo + blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t qEnd = qCount ? min(qHi, *qCount) : qHi;
    const bool live = i < qEnd && i < qn && tn >= 2;
    uint32_t qs = 0;
    float4 qb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    uint64_t qc = 0;
    if (live)
    {
        qs = qSorted[i] & kLeafIdMask;
        qb = fat[qs];
        qc = category[qs];
    }
    int stack[64];
    int sp = 0;
    int node = live ? 0 : -1;
    int staged = 0; // warp-uniform
    while (__any_sync(0xffffffffu, node >= 0))
    {
        bool emitL = false, emitR = false;
        int4 links = make_int4(0, 0, 0, 0);
        if (node >= 0)
        {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            const float4 boxL = __ldg(np);
            const float4 boxR = __ldg(np + 1);
            links = __ldg(reinterpret_cast<const int4*>(np + 2));
            const ulonglong2 cats = __ldg(reinterpret_cast<const ulonglong2*>(np + 3));
            const bool hitL = box_overlap(qb, boxL) && (qc & cats.x) != 0 && (!SELF || links.z > static_cast<int>(i));
            const bool hitR = box_overlap(qb, boxR) && (qc & cats.y) != 0 && (!SELF || links.w > static_cast<int>(i));
            emitL = hitL && links.x < 0 && (static_cast<uint32_t>(links.x) & rejectLeaf) == 0;
            emitR = hitR && links.y < 0 && (static_cast<uint32_t>(links.y) & rejectLeaf) == 0;
            const bool goL = hitL && links.x >= 0, goR = hitR && links.y >= 0;
            if (goL)
            {
                if (goR)
                {
                    if (sp < 64)
                        stack[sp++] = links.y;
                    else
                        *stackOverflow = 1;
                }
                node = links.x;
            }
            else if (goR)
                node = links.y;
            else
                node = sp > 0 ? stack[--sp] : -1;
        }
        const uint32_t eL = __ballot_sync(0xffffffffu, emitL), eR = __ballot_sync(0xffffffffu, emitR);
        if (eL | eR)
        {
            if (emitL)
            {
                const uint32_t ts = leaf_shape(links.x);
                stage[staged + __popc(eL & ltMask)] = make_uint2(ORD ? min(qs, ts) : qs, ORD ? max(qs, ts) : ts);
            }
            if (emitR)
            {
                const uint32_t ts = leaf_shape(links.y);
                stage[staged + __popc(eL) + __popc(eR & ltMask)] =
                    make_uint2(ORD ? min(qs, ts) : qs, ORD ? max(qs, ts) : ts);
            }
            staged += __popc(eL) + __popc(eR);
            __syncwarp();
            while (staged >= 32)
                stage_flush32(stage, staged, out, capacity, counter, lane);
        }
    }
    if (staged > 0)
    {
        uint32_t base = 0;
        if (lane == 0)
            base = atomicAdd(counter, static_cast<uint32_t>(staged));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < staged && base + lane < capacity)
            out[base + lane] = stage[lane];
    }
}

// Warp-per-query variant for SMALL query sets with large boxes (bullets: the leaf is a swept, displacement-fattened
// box that overlaps hundreds of leaves).  One thread per query would leave the GPU idle (10 k threads) and serialise
// each long traversal; here the 32 lanes of a warp pop up to 32 nodes of a shared stack per step and push / emit
// through ballots.  Same candidate semantics as k_traverse.
constexpr int kWarpStack = 1024;

template <bool SELF>
__device__ void lane_dfs(int root, uint32_t i, uint32_t qs, const float4& qb, uint64_t qc, uint32_t rejectLeaf,
                         const BvhNode* __restrict__ nodes, uint2* __restrict__ out, uint32_t capacity,
                         uint32_t* counter, uint32_t* stackOverflow)
{
    // overflow fallback of k_traverse_warp: a plain per-lane stack traversal of one subtree
    int stack[64];
    int sp = 0;
    int node = root;
    while (true)
    {
        const BvhNode N = nodes[node];
        bool hitL = box_overlap(qb, N.boxL) && (qc & N.catL) != 0 && (!SELF || N.lastL > static_cast<int>(i));
        bool hitR = box_overlap(qb, N.boxR) && (qc & N.catR) != 0 && (!SELF || N.lastR > static_cast<int>(i));
        if (hitL && N.left < 0)
        {
            const uint32_t ts = leaf_shape(N.left);
            if ((static_cast<uint32_t>(N.left) & rejectLeaf) == 0)
                emit_pair(SELF ? min(qs, ts) : qs, SELF ? max(qs, ts) : ts, out, capacity, counter);
            hitL = false;
        }
        if (hitR && N.right < 0)
        {
            const uint32_t ts = leaf_shape(N.right);
            if ((static_cast<uint32_t>(N.right) & rejectLeaf) == 0)
                emit_pair(SELF ? min(qs, ts) : qs, SELF ? max(qs, ts) : ts, out, capacity, counter);
            hitR = false;
        }
        if (hitL)
        {
            if (hitR)
            {
                if (sp < 64)
                    stack[sp++] = N.right;
                else
                    *stackOverflow = 1;
            }
            node = N.left;
        }
        else if (hitR)
            node = N.right;
        else
        {
            if (sp == 0)
                break;
            node = stack[--sp];
        }
    }
}

template <bool SELF>
__global__ void __launch_bounds__(128)
    k_traverse_warp(const uint32_t* __restrict__ qSorted, uint32_t qn, uint32_t qLo, uint32_t qHi,
                    const uint32_t* __restrict__ tSorted, uint32_t tn, const BvhNode* __restrict__ nodes,
                    const float4* __restrict__ fat, const uint64_t* __restrict__ category, uint32_t rejectLeaf,
                    uint2* __restrict__ out, uint32_t capacity, uint32_t* counter, uint32_t* stackOverflow)
{
    __shared__ int stacks[4][kWarpStack];
    const int lane = threadIdx.x & 31;
    int* stack = stacks[threadIdx.x >> 5];
    const uint32_t i = qLo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= qHi || i >= qn || tn == 0)
        return; // warp-uniform
    const uint32_t qs = qSorted[i] & kLeafIdMask;
    const float4 qb = fat[qs];
    const uint64_t qc = category[qs];
    if (tn == 1)
    {
        if (!SELF && lane == 0)
        {
            const uint32_t ts = tSorted[0] & kLeafIdMask;
            if ((tSorted[0] & rejectLeaf) == 0 && (qc & category[ts]) != 0 && box_overlap(qb, fat[ts]))
                emit_pair(qs, ts, out, capacity, counter);
        }
        return;
    }
    const uint32_t ltMask = (1u << lane) - 1u;
    int size = 1;
    if (lane == 0)
        stack[0] = 0;
    __syncwarp();
    while (size > 0)
    {
        const int take = size < 32 ? size : 32;
        const int node = lane < take ? stack[size - 1 - lane] : -1;
        __syncwarp();
        size -= take;
        bool pushL = false, pushR = false, emitL = false, emitR = false;
        int4 links = make_int4(0, 0, 0, 0);
        if (node >= 0)
        {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            const float4 boxL = __ldg(np);
            const float4 boxR = __ldg(np + 1);
            links = __ldg(reinterpret_cast<const int4*>(np + 2));
            const ulonglong2 cats = __ldg(reinterpret_cast<const ulonglong2*>(np + 3));
            const bool hitL = box_overlap(qb, boxL) && (qc & cats.x) != 0 && (!SELF || links.z > static_cast<int>(i));
            const bool hitR = box_overlap(qb, boxR) && (qc & cats.y) != 0 && (!SELF || links.w > static_cast<int>(i));
            emitL = hitL && links.x < 0 && (static_cast<uint32_t>(links.x) & rejectLeaf) == 0;
            emitR = hitR && links.y < 0 && (static_cast<uint32_t>(links.y) & rejectLeaf) == 0;
            pushL = hitL && links.x >= 0;
            pushR = hitR && links.y >= 0;
        }
        // pushes
        const uint32_t bL = __ballot_sync(0xffffffffu, pushL), bR = __ballot_sync(0xffffffffu, pushR);
        const int posL = size + __popc(bL & ltMask);
        const int posR = size + __popc(bL) + __popc(bR & ltMask);
        if (pushL)
        {
            if (posL < kWarpStack)
                stack[posL] = links.x;
            else
                lane_dfs<SELF>(links.x, i, qs, qb, qc, rejectLeaf, nodes, out, capacity, counter, stackOverflow);
        }
        if (pushR)
        {
            if (posR < kWarpStack)
                stack[posR] = links.y;
            else
                lane_dfs<SELF>(links.y, i, qs, qb, qc, rejectLeaf, nodes, out, capacity, counter, stackOverflow);
        }
        size = min(size + __popc(bL) + __popc(bR), kWarpStack);
        // emits: one atomic per step
        const uint32_t eL = __ballot_sync(0xffffffffu, emitL), eR = __ballot_sync(0xffffffffu, emitR);
        const int total = __popc(eL) + __popc(eR);
        if (total)
        {
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(counter, static_cast<uint32_t>(total));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (emitL)
            {
                const uint32_t idx = base + __popc(eL & ltMask);
                const uint32_t ts = leaf_shape(links.x);
                if (idx < capacity)
                    out[idx] = make_uint2(SELF ? min(qs, ts) : qs, SELF ? max(qs, ts) : ts);
            }
            if (emitR)
            {
                const uint32_t idx = base + __popc(eL) + __popc(eR & ltMask);
                const uint32_t ts = leaf_shape(links.y);
                if (idx < capacity)
                    out[idx] = make_uint2(SELF ? min(qs, ts) : qs, SELF ? max(qs, ts) : ts);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// K6 — narrowphase.  Candidates are binned by (mode, core) so a warp runs one code path:
//   mode 0 collide (no bullet), 1 sweep with A = bullet moving, 2 sweep with both moving
//   (narrowPhaseCollision, reference Registry.hpp:608-710).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pair_bin(uint32_t ma, uint32_t mb)
{
    const bool ba = meta_body(ma) == 2u, bb = meta_body(mb) == 2u;
    const int mode = (ba && bb) ? 2 : ((ba || bb) ? 1 : 0);
    return mode * CORE_COUNT + core_of(static_cast<int>(meta_type(ma)), static_cast<int>(meta_type(mb)));
}

// "Last block done" (the threadFenceReduction pattern): every thread that published results with global atomics fences,
// the block synchronises, one thread takes a ticket; the block that draws the last of `totalBlocks` tickets sees every
// other block's results and runs the tiny serial step that used to be a kernel of its own (k_bin_scan, k_hit_scan: one
// thread each, ~5 us of launch latency on the critical path of a 0.65 ms frame).  Call with the whole block converged.
__device__ __forceinline__ bool last_block_done(FrameScratch* fs, uint32_t totalBlocks)
{
    __shared__ bool isLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        isLast = atomicAdd(&fs->ticket, 1u) == totalBlocks - 1;
    __syncthreads();
    return isLast;
}

__device__ __forceinline__ void scan_bins(FrameScratch* fs)
{
    uint32_t run = 0;
    for (int b = 0; b < 16; ++b)
    {
        fs->binOffset[b] = run;
        run += atomicAdd(&fs->binCount[b], 0u); // through L2: the other blocks' atomics are there
    }
    fs->ticket = 0;
}

__device__ __forceinline__ void scan_hits(FrameScratch* fs)
{
    uint32_t run = 0;
    for (int b = 0; b < 16; ++b)
    {
        fs->hitBase[b] = run;
        run += atomicAdd(&fs->hitCount[b], 0u);
    }
    fs->hitBase[16] = run;
    fs->outCount = run;
    fs->ticket = 0;
}

__global__ void __launch_bounds__(256) k_bin_count(const uint2* __restrict__ cand, uint32_t capacity,
                                                   const uint32_t* __restrict__ meta, FrameScratch* fs)
{
    __shared__ uint32_t sh[16];
    if (threadIdx.x < 16)
        sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = min(fs->candCount, capacity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint2 p = cand[i];
        atomicAdd(&sh[pair_bin(meta[p.x], meta[p.y])], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 16 && sh[threadIdx.x])
        atomicAdd(&fs->binCount[threadIdx.x], sh[threadIdx.x]);
    // the last block turns the counts into offsets (was: k_bin_scan)
    if (last_block_done(fs, gridDim.x) && threadIdx.x == 0)
        scan_bins(fs);
}

__global__ void k_bin_scan(FrameScratch* fs) // (kept for reference; the pipeline no longer launches it)
{
    if (threadIdx.x == 0)
    {
        uint32_t run = 0;
        for (int b = 0; b < 16; ++b)
        {
            fs->binOffset[b] = run;
            run += fs->binCount[b];
        }
    }
}

// Tiles of 1024 candidates per block pass: ranks inside the tile come from shared-memory counters, the block reserves its
// slice of every bin with ONE global atomic per bin (16 per tile instead of several per warp on the same 16 addresses).
__global__ void __launch_bounds__(256) k_bin_scatter(const uint2* __restrict__ cand, uint32_t capacity,
                                                     const uint32_t* __restrict__ meta, FrameScratch* fs,
                                                     uint2* __restrict__ binned)
{
    constexpr int kPerThread = 4;
    __shared__ uint32_t count[16];
    __shared__ uint32_t base[16];
    const uint32_t n = min(fs->candCount, capacity);
    const uint32_t tile = blockDim.x * kPerThread;
    for (uint32_t t0 = blockIdx.x * tile; t0 < n; t0 += gridDim.x * tile)
    {
        if (threadIdx.x < 16)
            count[threadIdx.x] = 0;
        __syncthreads();
        uint2 p[kPerThread];
        int bin[kPerThread];
        uint32_t rank[kPerThread];
#pragma unroll
        for (int k = 0; k < kPerThread; ++k)
        {
            const uint32_t i = t0 + k * blockDim.x + threadIdx.x;
            bin[k] = -1;
            if (i < n)
            {
                p[k] = cand[i];
                bin[k] = pair_bin(meta[p[k].x], meta[p[k].y]);
                rank[k] = atomicAdd(&count[bin[k]], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 16)
            base[threadIdx.x] = count[threadIdx.x] ? fs->binOffset[threadIdx.x] + atomicAdd(&fs->binCursor[threadIdx.x], count[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kPerThread; ++k)
            if (bin[k] >= 0)
                binned[base[bin[k]] + rank[k]] = p[k];
        __syncthreads();
    }
}

__device__ __forceinline__ XF xf_from(float4 f)
{
    XF t;
    t.tx = f.x;
    t.ty = f.y;
    t.s = f.z;
    t.c = f.w;
    return t;
}

// Per-pair inputs shared by the two narrowphase phases.
struct PairCtx
{
    ShapeRef A, B;
    XF xa, xb;
    uint32_t ea, eb;
    bool bulletA, bulletB;
    bool exactTightB; // s.tight[shapeB] is the exact world AABB of B (B never translated since its box was computed)
};

__device__ __forceinline__ bool load_pair(const ShapeArrays& s, uint2 p, PairCtx& c)
{
    const uint32_t ma = s.meta[p.x], mb = s.meta[p.y];
    c.ea = s.entity[p.x];
    c.eb = s.entity[p.y];
    // self-collision and inactive shapes are rejected here, not in the broadphase (Registry.hpp:619-624)
    if (c.ea == c.eb || !(ma & META_ACTIVE) || !(mb & META_ACTIVE))
        return false;
    // Multi-GPU owner rule: every rank holds its slab plus halo ("ghost") copies of the neighbours'
    // boundary shapes.  A pair belongs to exactly one rank: the one that OWNS the pair's owner shape — shape A for
    // cross-tree pairs (the Dynamic / Bullet side), the shape of the entity with the smaller tag for same-tree pairs
    // (entity tags are global, local shape ids are not).
    if ((ma | mb) & META_GHOST)
    {
        const bool sameTree = meta_body(ma) == meta_body(mb);
        const uint32_t ownerMeta = (!sameTree || c.ea < c.eb) ? ma : mb;
        if (ownerMeta & META_GHOST)
            return false;
    }
    c.A.type = static_cast<int>(meta_type(ma));
    c.A.count = static_cast<int>(meta_count(ma));
    c.A.g = s.geom + static_cast<size_t>(p.x) * 32;
    c.B.type = static_cast<int>(meta_type(mb));
    c.B.count = static_cast<int>(meta_count(mb));
    c.B.g = s.geom + static_cast<size_t>(p.y) * 32;
    c.xa = xf_from(s.xf[p.x]);
    c.xb = xf_from(s.xf[p.y]);
    c.bulletA = meta_body(ma) == 2u;
    c.bulletB = meta_body(mb) == 2u;
    c.exactTightB = (mb & META_TIGHT_EXACT) != 0;
    return true;
}

// A is the bullet whenever exactly one side is (cross-tree pairs are emitted bullet-first)
__device__ __forceinline__ void load_motions(const ShapeArrays& s, uint2 p, const PairCtx& c, Motion& moA, Motion& moB)
{
    const float4 pa = s.pxf[p.x], paa = s.pxa[p.x], ca = s.xa[p.x];
    moA = motion_between(xf_from(pa), paa.x, paa.y, paa.z, c.xa, ca.x);
    if (c.bulletB)
    {
        const float4 pb = s.pxf[p.y], pba = s.pxa[p.y], cb = s.xa[p.y];
        moB = motion_between(xf_from(pb), pba.x, pba.y, pba.z, c.xb, cb.x);
    }
    else
        moB = motion_fixed(c.xb);
}

// Conservative early-out for translation-only sweeps.  The broadphase candidate comes from the bullet's LEAF box
// (swept and displacement-fattened, DynamicBVH.hpp:386-400) which is far larger than the region the shape really
// sweeps.  For delta-angle == 0 every sample transform of sweepHelper is a translation between start and end with the
// same (libm) sin/cos, so the shape stays inside the union of its AABBs at the two ends (plus its AABB at the STORED
// start transform, used by the t = 0 test).  If that union and the other body's are further apart than a margin that
// dwarfs the narrowphase slack (1e-6) and float rounding, none of the <= 19 boolean tests can succeed: skipping them
// cannot change the result.
struct SweptBounds
{
    Box start; // the shape at the start of the sweep: stored transform (t = 0 test) and libm sin/cos (later samples)
    Box all;   // ... united with the shape at the end of the sweep
};

__device__ __forceinline__ SweptBounds swept_extent(const ShapeRef& S, const Motion& mo)
{
    SweptBounds r;
    r.start = compute_aabb(static_cast<uint32_t>(S.type), static_cast<uint32_t>(S.count), S.g, mo.start);
    r.all = r.start;
    if (mo.moving)
    {
        XF t = mo.start;
        t.s = mo.csin;
        t.c = mo.ccos;
        r.start = box_combine(r.start, compute_aabb(static_cast<uint32_t>(S.type), static_cast<uint32_t>(S.count), S.g, t));
        t.tx = mo.start.tx + mo.dx;
        t.ty = mo.start.ty + mo.dy;
        r.all = box_combine(r.start, compute_aabb(static_cast<uint32_t>(S.type), static_cast<uint32_t>(S.count), S.g, t));
    }
    return r;
}

// exactB: the stored tight box of B when B does not move and that box is known to be computeAABB(B, its transform) —
// then B's 128-byte geometry block need not be touched for the (vast majority of) candidates rejected here.
__device__ __forceinline__ bool sweep_cannot_hit(const ShapeRef& A, const Motion& ma, const ShapeRef& B, const Motion& mb,
                                                 const float4* exactB = nullptr)
{
    if ((ma.moving && ma.da != 0.0f) || (mb.moving && mb.da != 0.0f))
        return false; // rotating sweep: no cheap bound
    const SweptBounds sa = swept_extent(A, ma);
    SweptBounds sb;
    if (exactB != nullptr && !mb.moving)
        sb.start = sb.all = box_from(*exactB);
    else
        sb = swept_extent(B, mb);
    const Box &a = sa.all, &b = sb.all;
    const float scale = fmaxf(fmaxf(fmaxf(fabsf(a.minx), fabsf(a.maxx)), fmaxf(fabsf(a.miny), fabsf(a.maxy))),
                              fmaxf(fmaxf(fabsf(b.minx), fabsf(b.maxx)), fmaxf(fabsf(b.miny), fabsf(b.maxy))));
    const float margin = 1e-3f + 1e-5f * scale;
    if (a.minx > b.maxx + margin || b.minx > a.maxx + margin || a.miny > b.maxy + margin || b.miny > a.maxy + margin)
        return true;
    // The union boxes of a long diagonal move are mostly empty.  Tighter: during a translation-only sweep each shape
    // stays inside the disc that circumscribes its start box, carried along its motion; the discs' centres move
    // relative to each other along one segment.  If the smallest centre distance over t in [0, 1] exceeds the two
    // radii (plus the same generous margin), no sample of sweepHelper can find the shapes in contact.
    const float ax = 0.5f * (sa.start.minx + sa.start.maxx), ay = 0.5f * (sa.start.miny + sa.start.maxy);
    const float bx = 0.5f * (sb.start.minx + sb.start.maxx), by = 0.5f * (sb.start.miny + sb.start.maxy);
    const float ra = 0.5f * sqrtf((sa.start.maxx - sa.start.minx) * (sa.start.maxx - sa.start.minx) +
                                  (sa.start.maxy - sa.start.miny) * (sa.start.maxy - sa.start.miny));
    const float rb = 0.5f * sqrtf((sb.start.maxx - sb.start.minx) * (sb.start.maxx - sb.start.minx) +
                                  (sb.start.maxy - sb.start.miny) * (sb.start.maxy - sb.start.miny));
    const float rx = ax - bx, ry = ay - by;
    const float dx = (ma.moving ? ma.dx : 0.0f) - (mb.moving ? mb.dx : 0.0f);
    const float dy = (ma.moving ? ma.dy : 0.0f) - (mb.moving ? mb.dy : 0.0f);
    const float dd = dx * dx + dy * dy;
    float t = dd > 0.0f ? -(rx * dx + ry * dy) / dd : 0.0f;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    const float cx = rx + t * dx, cy = ry + t * dy;
    const float reach = (ra + rb) * 1.0001f + 2.0f * margin;
    return cx * cx + cy * cy > reach * reach; // NaN compares false: keep the pair
}

// Phase 1: boolean test (areColliding / sweepHelper) of every binned candidate; hits are compacted per bin
// into hits[binOffset[bin] + k] = (shapeA, shapeB, fraction bits, 0).  Most candidates are rejected here by
// the cheap early-out paths, so the expensive manifold code of phase 2 runs on full warps of real hits.
// All lanes of the warp call this; lanes with `hit` append (shapeA, shapeB, fraction) to the hit list of entry i's bin.
__device__ __forceinline__ void append_hits(bool hit, uint32_t i, uint2 p, float fraction, FrameScratch* fs,
                                            uint4* __restrict__ hits, int lane)
{
    // the bin of entry i: the binned array is the concatenation of the bins
    int bin = 31;
    if (hit)
    {
        bin = 0;
#pragma unroll
        for (int b = 1; b < 15; ++b)
            bin += (i >= fs->binOffset[b]) ? 1 : 0;
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, bin);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (hit && lane == leader)
        base = atomicAdd(&fs->hitCount[bin], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (hit)
        hits[fs->binOffset[bin] + base + __popc(peers & ((1u << lane) - 1u))] =
            make_uint4(p.x, p.y, __float_as_uint(fraction), 0u);
}

struct SweepTodo // a bullet candidate that survived the conservative bound and needs the real sweep search
{
    uint2 p;
    uint32_t i;
};

// The sweep searches (<= 19 boolean tests each) of `count` queued candidates, one per lane.
__device__ __forceinline__ void run_sweeps(const SweepTodo* queue, int count, const ShapeArrays& s, FrameScratch* fs,
                                           uint4* __restrict__ hits, int lane)
{
    uint2 p = make_uint2(0, 0);
    uint32_t i = 0;
    float fraction = 0.0f;
    bool hit = false;
    if (lane < count)
    {
        p = queue[lane].p;
        i = queue[lane].i;
        PairCtx c;
        load_pair(s, p, c); // accepted once already
        Motion moA, moB;
        load_motions(s, p, c, moA, moB);
        hit = sweep_search(c.A, moA, c.B, moB, fraction);
    }
    __syncwarp();
    append_hits(hit, i, p, fraction, fs, hits, lane);
}

// BULLETS = false: the scene holds no Bullet shape, the sweep code (and its registers: 96 -> occupancy 28 %) is compiled out.
// scanBlocks: the number of blocks of k_narrow_filter AND k_narrow_filter_pp of this frame (they run side by side); the
// block that finishes last among all of them turns the hit counts into output offsets (was: k_hit_scan).
template <bool BULLETS>
__global__ void __launch_bounds__(128) k_narrow_filter(const uint2* __restrict__ binned, uint32_t capacity,
                                                       ShapeArrays s, FrameScratch* fs, uint4* __restrict__ hits,
                                                       uint32_t scanBlocks)
{
    // Bullet candidates: the conservative bound rejects nearly all of them (98 % among 1 M static polygons), and the few
    // survivors would each run their 19-test sweep search alone in a warp.  They are queued per warp instead and searched
    // 32 at a time.
    __shared__ SweepTodo queues[4][64];
    SweepTodo* queue = queues[threadIdx.x >> 5];
    int queued = 0; // warp-uniform
    const uint32_t n = min(fs->candCount, capacity);
    // the polygon bin of mode 0 is handled by k_narrow_filter_pp (8 lanes per pair)
    const uint32_t ppOff = fs->binOffset[CORE_PP], ppCnt = fs->binCount[CORE_PP];
    const uint32_t total = n - ppCnt;
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count and an explicit reconvergence per iteration: without it the lanes that reject their
    // pair early run ahead into the next iteration and the warp never executes the test code with a full mask
    for (uint32_t base0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base0 < total; base0 += gridDim.x * blockDim.x)
    {
        const uint32_t i0 = base0 + lane;
        const bool valid = i0 < total;
        const uint32_t i = i0 < ppOff ? i0 : i0 + ppCnt;
        uint2 p = make_uint2(0, 0);
        bool hit = false, sweep = false;
        if (valid)
        {
            p = binned[i];
            PairCtx c;
            if (load_pair(s, p, c))
            {
                if (!BULLETS || (!c.bulletA && !c.bulletB))
                {
                    Manifold unused;
                    hit = narrow_test<false>(c.A, c.xa, c.B, c.xb, unused);
                }
                else
                {
                    Motion moA, moB;
                    load_motions(s, p, c, moA, moB);
                    sweep = !sweep_cannot_hit(c.A, moA, c.B, moB, c.exactTightB ? &s.tight[p.y] : nullptr);
                }
            }
        }
        __syncwarp();
        append_hits(hit, i, p, 0.0f, fs, hits, lane);
        const uint32_t todo = BULLETS ? __ballot_sync(0xffffffffu, sweep) : 0u;
        if (BULLETS && todo)
        {
            if (sweep)
                queue[queued + __popc(todo & ((1u << lane) - 1u))] = SweepTodo{p, i};
            queued += __popc(todo);
            __syncwarp();
            if (queued >= 32)
            {
                queued -= 32;
                run_sweeps(queue + queued, 32, s, fs, hits, lane);
                __syncwarp();
            }
        }
    }
    if (BULLETS && queued > 0)
        run_sweeps(queue, queued, s, fs, hits, lane);
    if (last_block_done(fs, scanBlocks) && threadIdx.x == 0)
        scan_hits(fs);
}

__global__ void k_hit_scan(FrameScratch* fs) // (kept for reference; the pipeline no longer launches it)
{
    if (threadIdx.x == 0)
    {
        uint32_t run = 0;
        for (int b = 0; b < 16; ++b)
        {
            fs->hitBase[b] = run;
            run += fs->hitCount[b];
        }
        fs->hitBase[16] = run;
        fs->outCount = run;
    }
}

// Phase 2: the manifold of every hit (collide(), or collide() at the sweep fraction), written as an 84-byte
// EntityCollision record at out[hitBase[bin] + k]: no atomics, output grouped by bin.
template <bool BULLETS>
__global__ void __launch_bounds__(128) k_narrow_manifold(const uint4* __restrict__ hits, ShapeArrays s,
                                                         const FrameScratch* __restrict__ fs,
                                                         PairOut* __restrict__ out, uint32_t outCapacity)
{
    const uint32_t total = fs->hitBase[16];
    // the polygon bin of mode 0 is handled by k_narrow_manifold_pp
    const uint32_t ppBase = fs->hitBase[CORE_PP], ppHits = fs->hitCount[CORE_PP];
    for (uint32_t j0 = blockIdx.x * blockDim.x + threadIdx.x; j0 < total - ppHits; j0 += gridDim.x * blockDim.x)
    {
        const uint32_t j = j0 < ppBase ? j0 : j0 + ppHits;
        int bin = 0;
#pragma unroll
        for (int b = 1; b < 15; ++b)
            bin += (j >= fs->hitBase[b]) ? 1 : 0;
        const uint4 h = hits[fs->binOffset[bin] + (j - fs->hitBase[bin])];
        const uint2 p = make_uint2(h.x, h.y);
        PairCtx c;
        load_pair(s, p, c);
        Manifold m;
        if (!BULLETS || (!c.bulletA && !c.bulletB))
        {
            narrow_test<true>(c.A, c.xa, c.B, c.xb, m);
        }
        else
        {
            // a sweep hit is reported even if the final collide() yields no point (Registry.hpp:644,709:
            // the optional holds the manifold as soon as sweep() returned a value)
            Motion moA, moB;
            load_motions(s, p, c, moA, moB);
            sweep_manifold(c.A, moA, c.B, moB, __uint_as_float(h.z), m);
        }
        if (j < outCapacity)
        {
            PairOut* o = out + j;
            o->entityA = c.ea;
            o->entityB = c.eb;
            o->shapeA = p.x;
            o->shapeB = p.y;
            manifold_store(m, o->manifold);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Reference result order (c2d_gpu_set_result_order): the reference returns its five groups one after the other - Dynamic x
// Static, Dynamic x Dynamic, Bullet x Static, Bullet x Dynamic, Bullet x Bullet - each sorted by (shapeA, shapeB)
// (Registry.hpp:487-545: std::sort of the candidate list, narrowphase in that order).  The records come out of the
// manifold kernels grouped by narrowphase bin; two stable radix sorts of an index permutation (by shapeB, then by
// group | shapeA) and one gather put them in that order.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_order_key_b(const PairOut* __restrict__ out, uint32_t n, uint32_t* __restrict__ keys,
                                                     uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    keys[i] = out[i].shapeB;
    vals[i] = i;
}

__global__ void __launch_bounds__(256) k_order_key_a(const PairOut* __restrict__ out, uint32_t n, const uint32_t* __restrict__ perm,
                                                     const uint32_t* __restrict__ meta, uint32_t* __restrict__ keys)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const PairOut* o = out + perm[i];
    const uint32_t ba = meta_body(meta[o->shapeA]), bb = meta_body(meta[o->shapeB]);
    // (Dynamic, Static) 0, (Dynamic, Dynamic) 1, (Bullet, Static) 2, (Bullet, Dynamic) 3, (Bullet, Bullet) 4
    const uint32_t group = ba == 1u ? (bb == 0u ? 0u : 1u) : (bb == 0u ? 2u : (bb == 1u ? 3u : 4u));
    keys[i] = (group << 29) | (o->shapeA & 0x1fffffffu);
}

// one warp per record: 21 coalesced 4-byte words
__global__ void __launch_bounds__(256) k_order_gather(const PairOut* __restrict__ out, uint32_t n, const uint32_t* __restrict__ perm,
                                                      PairOut* __restrict__ sorted)
{
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n)
        return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(out + perm[w]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sorted + w);
    if (lane < 21)
        dst[lane] = src[lane];
}

// ---------------------------------------------------------------------------------------------------
// Small scenes (a few thousand shapes — the reference's own 1k / 10k benchmarks): the frame is bound by the ~30
// dependent launches of the tree pipeline, not by work.  Two kernels replace sort + build + fit + 5 traversals and
// the 8 narrowphase launches: an all-pairs box test over the member lists (tiles of targets staged in shared
// memory; same predicate and same pair orientation as k_traverse) and a one-thread-per-candidate narrowphase.
// ---------------------------------------------------------------------------------------------------
constexpr int kSmallQueries = 128; // queries per block (one per thread)
constexpr int kSmallTile = 64;     // targets per block: grid = (query blocks, target tiles), 64 tests per thread

__global__ void __launch_bounds__(kSmallQueries)
    k_small_pairs(const uint32_t* __restrict__ membersS, uint32_t nS, const uint32_t* __restrict__ membersD, uint32_t nD,
                  const uint32_t* __restrict__ membersB, uint32_t nB, const float4* __restrict__ fat,
                  const uint64_t* __restrict__ category, uint2* __restrict__ out, uint32_t capacity, uint32_t* counter)
{
    __shared__ float4 tileBox[kSmallTile];
    __shared__ uint64_t tileCat[kSmallTile];
    __shared__ uint32_t tileShape[kSmallTile];
    // targets: Static, Dynamic, Bullet lists back to back; this block sees the tile blockIdx.y of them
    const uint32_t nT = nS + nD + nB;
    const uint32_t t0 = blockIdx.y * kSmallTile;
    if (threadIdx.x < kSmallTile && t0 + threadIdx.x < nT)
    {
        const uint32_t t = t0 + threadIdx.x;
        const uint32_t ts = t < nS ? membersS[t] : (t < nS + nD ? membersD[t - nS] : membersB[t - nS - nD]);
        tileShape[threadIdx.x] = ts;
        tileBox[threadIdx.x] = fat[ts];
        tileCat[threadIdx.x] = category[ts];
    }
    __syncthreads();
    // queries: the Dynamic members, then the Bullet members (Static shapes never query, Registry.hpp:487-545)
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nD + nB)
        return;
    const bool bullet = q >= nD;
    const uint32_t qs = bullet ? membersB[q - nD] : membersD[q];
    const float4 qb = fat[qs];
    const uint64_t qc = category[qs];
    const uint32_t count = min(static_cast<uint32_t>(kSmallTile), nT - t0);
    for (uint32_t k = 0; k < count; ++k)
    {
        const uint32_t ti = t0 + k; // position in the concatenated target list; nS + q is the query's own position
        // Dynamic queries: all Static, Dynamic after itself.  Bullet queries: all Static, all Dynamic, Bullets after itself.
        const bool wanted = ti < nS || (bullet ? (ti < nS + nD || ti > nS + q) : (ti > nS + q && ti < nS + nD));
        if (wanted && (qc & tileCat[k]) != 0 && box_overlap(qb, tileBox[k]))
        {
            const uint32_t ts = tileShape[k];
            const bool sameTree = bullet ? ti >= nS + nD : ti >= nS;
            emit_pair(sameTree ? min(qs, ts) : qs, sameTree ? max(qs, ts) : ts, out, capacity, counter);
        }
    }
}

// `out` is pinned host memory mapped into the device's address space: a small scene's few records cross PCIe as they are
// written, so the host needs ONE synchronisation per frame (no count read-back followed by a second copy).  The last
// block to finish publishes the frame counters to `hostFs` (mapped as well) and resets the device counters, which leaves
// the scratch ready for the next small frame without a k_frame_init launch.
__global__ void __launch_bounds__(128) k_small_narrow(const uint2* __restrict__ cand, uint32_t capacity, ShapeArrays s,
                                                      FrameScratch* fs, PairOut* __restrict__ out, uint32_t outCapacity,
                                                      FrameScratch* hostFs)
{
    const uint32_t n = min(fs->candCount, capacity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint2 p = cand[i];
        PairCtx c;
        if (!load_pair(s, p, c))
            continue;
        Manifold m;
        bool hit;
        if (!c.bulletA && !c.bulletB)
            hit = narrow_test<true>(c.A, c.xa, c.B, c.xb, m);
        else
        {
            Motion moA, moB;
            load_motions(s, p, c, moA, moB);
            float fraction = 0.0f;
            hit = !sweep_cannot_hit(c.A, moA, c.B, moB) && sweep_search(c.A, moA, c.B, moB, fraction);
            if (hit)
                sweep_manifold(c.A, moA, c.B, moB, fraction, m);
        }
        if (!hit)
            continue;
        const uint32_t j = atomicAdd(&fs->outCount, 1u);
        if (j < outCapacity)
        {
            PairOut* o = out + j;
            o->entityA = c.ea;
            o->entityB = c.eb;
            o->shapeA = p.x;
            o->shapeB = p.y;
            manifold_store(m, o->manifold);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        if (atomicAdd(&fs->ticket, 1u) == gridDim.x - 1)
        {
            hostFs->candCount = fs->candCount;
            hostFs->outCount = atomicAdd(&fs->outCount, 0u);
            hostFs->clipOverflow = atomicAdd(&fs->clipOverflow, 0u);
            hostFs->stackOverflow = 0;
            fs->candCount = 0;
            fs->outCount = 0;
            fs->clipOverflow = 0;
            fs->stackOverflow = 0;
            fs->ticket = 0;
            __threadfence_system();
        }
    }
}

// ---- polygon bin of mode 0 (polygon vs polygon | capsule | segment, no bullet): 8 lanes per pair --------------
__device__ __forceinline__ bool pp_setup(const ShapeArrays& s, uint2 p, bool valid, PairCtx& c, bool& rev)
{
    rev = false;
    if (!valid)
        return false;
    const bool ok = load_pair(s, p, c);
    // collide(polygon, capsule|segment) = collide(capsule, tB, polygon, tA).reverse() (Collision.cpp:498-510)
    rev = c.A.type == T_POLYGON && c.B.type != T_POLYGON;
    return ok;
}

__global__ void __launch_bounds__(128) k_narrow_filter_pp(const uint2* __restrict__ binned, ShapeArrays s,
                                                          FrameScratch* fs, uint4* __restrict__ hits, uint32_t scanBlocks)
{
    __shared__ OctetShared sh[16];
    const uint32_t off = fs->binOffset[CORE_PP], cnt = fs->binCount[CORE_PP];
    const int sub = threadIdx.x & 7, octetInWarp = (threadIdx.x & 31) >> 3, octetInBlock = threadIdx.x >> 3;
    const uint32_t octetsPerGrid = gridDim.x * 16;
    bool overflow = false;
    for (uint32_t first = blockIdx.x * 16 + (octetInBlock & ~3u); first < cnt; first += octetsPerGrid)
    {
        const uint32_t idx = first + octetInWarp;
        const bool valid = idx < cnt;
        const uint2 p = valid ? binned[off + idx] : make_uint2(0, 0);
        PairCtx c;
        bool rev;
        const bool alive = pp_setup(s, p, valid, c, rev);
        Manifold unused;
        const bool hit = octet_poly_poly<false>(alive, sub, octetInWarp, rev ? c.B : c.A, rev ? c.xb : c.xa,
                                                rev ? c.A : c.B, rev ? c.xa : c.xb, sh[octetInBlock], unused, overflow);
        __syncwarp();
        if (hit && sub == 0)
        {
            cg::coalesced_group g = cg::coalesced_threads();
            uint32_t base = 0;
            if (g.thread_rank() == 0)
                base = atomicAdd(&fs->hitCount[CORE_PP], g.size());
            base = g.shfl(base, 0);
            hits[off + base + g.thread_rank()] = make_uint4(p.x, p.y, 0u, 0u);
        }
        __syncwarp();
    }
    if (last_block_done(fs, scanBlocks) && threadIdx.x == 0)
        scan_hits(fs);
}

__global__ void __launch_bounds__(128, 8) k_narrow_manifold_pp(const uint4* __restrict__ hits, ShapeArrays s,
                                                            FrameScratch* fs, PairOut* __restrict__ out,
                                                            uint32_t outCapacity)
{
    __shared__ OctetShared sh[16];
    const uint32_t off = fs->binOffset[CORE_PP], cnt = fs->hitCount[CORE_PP], outBase = fs->hitBase[CORE_PP];
    const int sub = threadIdx.x & 7, octetInWarp = (threadIdx.x & 31) >> 3, octetInBlock = threadIdx.x >> 3;
    const uint32_t octetsPerGrid = gridDim.x * 16;
    bool overflow = false;
    for (uint32_t first = blockIdx.x * 16 + (octetInBlock & ~3u); first < cnt; first += octetsPerGrid)
    {
        const uint32_t idx = first + octetInWarp;
        const bool valid = idx < cnt;
        const uint4 h = valid ? hits[off + idx] : make_uint4(0, 0, 0, 0);
        const uint2 p = make_uint2(h.x, h.y);
        PairCtx c;
        bool rev;
        const bool alive = pp_setup(s, p, valid, c, rev);
        Manifold m;
        const bool hit = octet_poly_poly<true>(alive, sub, octetInWarp, rev ? c.B : c.A, rev ? c.xb : c.xa,
                                               rev ? c.A : c.B, rev ? c.xa : c.xb, sh[octetInBlock], m, overflow);
        __syncwarp();
        if (valid && sub == 0 && outBase + idx < outCapacity)
        {
            if (!hit)
                manifold_clear(m); // cannot happen: phase 1 accepted the pair with the same arithmetic
            else if (rev)
                manifold_reverse(m);
            PairOut* o = out + outBase + idx;
            o->entityA = c.ea;
            o->entityB = c.eb;
            o->shapeA = p.x;
            o->shapeB = p.y;
            manifold_store(m, o->manifold);
        }
        __syncwarp();
    }
    if (overflow)
        fs->clipOverflow = 1;
}

} // namespace c2d_dev
