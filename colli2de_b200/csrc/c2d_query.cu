// c2d_query.cu — the per-entity queries of the Registry as batched GPU queries:
//   Registry::getCollisions(id)     reference include/colli2de/Registry.hpp:551-605
//   Registry::areColliding(a, b)    reference Registry.hpp:748-864  (boolSweepHelper: Collision.hpp:126-149, 215-244)
// Same trees, same narrowphase device functions as getCollidingPairs; only the front end differs:
//   * getCollisions queries with the TIGHT box of every active shape of the entity (bullets: combine(computeAABB at
//     the previous transform, tight box)), filters leaves by (leaf.category & shape.maskBits) != 0 — the MASK is
//     used here, unlike getCollidingPairs — skips the Static tree for a Static entity, and reports
//     (queried shape, other shape) with the manifold of narrowPhaseCollision: when exactly one side is a bullet the
//     manifold is the bullet's sweep against the other shape whichever side was queried (Registry.hpp:669-691).
//   * areColliding evaluates every (shape of a, shape of b) pair without a broadphase; bullets use the boolean
//     sweep (start + 8 coarse samples, no bisection).
#include "c2d_ctx.hpp"
#include "c2d_narrow.cuh"

#include <cooperative_groups.h>

using namespace c2d_dev;
using namespace c2d_host;
namespace cg = cooperative_groups;

namespace
{

struct TreeView
{
    const uint32_t* sorted;
    const BvhNode* nodes;
    uint32_t n;
};

__device__ __forceinline__ XF xf4(float4 f)
{
    XF t;
    t.tx = f.x;
    t.ty = f.y;
    t.s = f.z;
    t.c = f.w;
    return t;
}

__device__ __forceinline__ ShapeRef shape_ref(const ShapeArrays& s, uint32_t shape, uint32_t meta)
{
    ShapeRef r;
    r.type = static_cast<int>(meta_type(meta));
    r.count = static_cast<int>(meta_count(meta));
    r.g = s.geom + static_cast<size_t>(shape) * 32;
    return r;
}

__device__ __forceinline__ Motion motion_of(const ShapeArrays& s, uint32_t shape)
{
    const float4 p = s.pxf[shape], pa = s.pxa[shape], ca = s.xa[shape];
    return motion_between(xf4(p), pa.x, pa.y, pa.z, xf4(s.xf[shape]), ca.x);
}

// narrowPhaseCollision(shapeQueried = a, other = b) (Registry.hpp:608-710)
__device__ void query_narrow(const ShapeArrays& s, uint32_t a, uint32_t ma, uint32_t b, PairOut* __restrict__ out,
                             uint32_t capacity, uint32_t* counter)
{
    const uint32_t mb = s.meta[b];
    const uint32_t ea = s.entity[a], eb = s.entity[b];
    if (ea == eb || !(ma & META_ACTIVE) || !(mb & META_ACTIVE))
        return;
    const ShapeRef A = shape_ref(s, a, ma), B = shape_ref(s, b, mb);
    const bool ba = meta_body(ma) == 2u, bb = meta_body(mb) == 2u;
    Manifold m;
    bool hit;
    float fraction;
    if (ba && bb)
        hit = narrow_sweep(A, motion_of(s, a), B, motion_of(s, b), fraction, m);
    else if (ba)
        hit = narrow_sweep(A, motion_of(s, a), B, motion_fixed(xf4(s.xf[b])), fraction, m);
    else if (bb) // the bullet is the moving shape and comes first in sweep(), the record keeps (queried, other)
        hit = narrow_sweep(B, motion_of(s, b), A, motion_fixed(xf4(s.xf[a])), fraction, m);
    else
        hit = narrow_test<true>(A, xf4(s.xf[a]), B, xf4(s.xf[b]), m);
    if (!hit)
        return;
    const uint32_t idx = atomicAdd(counter, 1u);
    if (idx < capacity)
    {
        PairOut* o = out + idx;
        o->entityA = ea;
        o->entityB = eb;
        o->shapeA = a;
        o->shapeB = b;
        manifold_store(m, o->manifold);
    }
}

__device__ void query_tree(const ShapeArrays& s, const TreeView& tree, uint32_t a, uint32_t ma, const float4& qb,
                           uint64_t mask, PairOut* __restrict__ out, uint32_t capacity, uint32_t* counter)
{
    if (tree.n == 0)
        return;
    if (tree.n == 1)
    {
        const uint32_t b = tree.sorted[0];
        if ((s.category[b] & mask) != 0 && box_overlap(s.fat[b], qb))
            query_narrow(s, a, ma, b, out, capacity, counter);
        return;
    }
    int stack[64];
    int sp = 0;
    int node = 0;
    while (true)
    {
        const BvhNode N = tree.nodes[node];
        bool hitL = (N.catL & mask) != 0 && box_overlap(N.boxL, qb);
        bool hitR = (N.catR & mask) != 0 && box_overlap(N.boxR, qb);
        if (hitL && N.left < 0)
        {
            query_narrow(s, a, ma, leaf_shape(N.left), out, capacity, counter);
            hitL = false;
        }
        if (hitR && N.right < 0)
        {
            query_narrow(s, a, ma, leaf_shape(N.right), out, capacity, counter);
            hitR = false;
        }
        if (hitL)
        {
            if (hitR && sp < 64)
                stack[sp++] = N.right;
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

__global__ void __launch_bounds__(64) k_query_shapes(const uint32_t* __restrict__ shapes, uint32_t n, TreeView t0,
                                                     TreeView t1, TreeView t2, ShapeArrays s, PairOut* __restrict__ out,
                                                     uint32_t capacity, uint32_t* counter)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n)
        return;
    const uint32_t a = shapes[q];
    const uint32_t ma = s.meta[a];
    if (!(ma & META_ALIVE) || !(ma & META_ACTIVE))
        return; // inactive shapes of the entity are not queried (Registry.hpp:568-570, 583-585)
    const uint32_t body = meta_body(ma);
    float4 qb = s.tight[a];
    if (body == 2u)
    {
        // combine(computeAABB(shape, previousTransform), shape.mBoundingBox) (Registry.hpp:572-576)
        const Box old = compute_aabb(meta_type(ma), meta_count(ma), s.geom + static_cast<size_t>(a) * 32, xf4(s.pxf[a]));
        qb = box_to(box_combine(old, box_from(qb)));
    }
    const uint64_t mask = s.mask[a];
    if (body != 0u)
        query_tree(s, t0, a, ma, qb, mask, out, capacity, counter);
    query_tree(s, t1, a, ma, qb, mask, out, capacity, counter);
    query_tree(s, t2, a, ma, qb, mask, out, capacity, counter);
}

// boolSweepHelper (Collision.hpp:126-149 / 215-244): stored start transforms, then 8 coarse samples
__device__ bool sweep_bool(const ShapeRef& A, const Motion& ma, const ShapeRef& B, const Motion& mb)
{
    Manifold unused;
    if (narrow_test<false>(A, ma.start, B, mb.start, unused))
        return true;
    for (int step = 1; step <= 8; ++step)
    {
        const float f = static_cast<float>(step) / 8.0f;
        if (narrow_test<false>(A, motion_at(ma, f), B, motion_at(mb, f), unused))
            return true;
    }
    return false;
}

__global__ void __launch_bounds__(64) k_pairs_colliding(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ sb,
                                                        uint32_t n, ShapeArrays s, uint8_t* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t a = sa[i], b = sb[i];
    const uint32_t ma = s.meta[a], mb = s.meta[b];
    bool hit = false;
    if ((ma & META_ACTIVE) && (mb & META_ACTIVE) && (ma & META_ALIVE) && (mb & META_ALIVE))
    {
        const ShapeRef A = shape_ref(s, a, ma), B = shape_ref(s, b, mb);
        const bool ba = meta_body(ma) == 2u, bb = meta_body(mb) == 2u;
        Manifold unused;
        if (ba && bb)
            hit = sweep_bool(A, motion_of(s, a), B, motion_of(s, b));
        else if (ba)
            hit = sweep_bool(A, motion_of(s, a), B, motion_fixed(xf4(s.xf[b])));
        else if (bb)
            hit = sweep_bool(B, motion_of(s, b), A, motion_fixed(xf4(s.xf[a])));
        else
            hit = narrow_test<false>(A, xf4(s.xf[a]), B, xf4(s.xf[b]), unused);
    }
    out[i] = hit ? 1 : 0;
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

extern "C" {

int c2d_gpu_query_shapes(c2d_gpu_ctx* ctx, uint32_t n, const uint32_t* shapes, const c2d_pair_record** out_records,
                         uint64_t* out_count)
{
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    if (out_count)
        *out_count = 0;
    if (n == 0)
        return 0;
    if (!shapes)
        return fail(ctx, "c2d_gpu_query_shapes: shapes is NULL");
    for (uint32_t i = 0; i < n; ++i)
        if (shapes[i] >= ctx->hostMeta.size())
            return fail(ctx, "c2d_gpu_query_shapes: unknown shape");
    if (int rc = ensure_trees(ctx))
        return rc;
    QueryBuffers& qb = ctx->query;
    C2D_CUDA(ctx, qb.shapes.reserve(n));
    C2D_CUDA(ctx, qb.counter.reserve(1));
    if (qb.out.capacity == 0)
        C2D_CUDA(ctx, qb.out.reserve(4096));
    C2D_CUDA(ctx, cudaMemcpyAsync(qb.shapes.ptr, shapes, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    uint32_t count = 0;
    while (true)
    {
        C2D_CUDA(ctx, cudaMemsetAsync(qb.counter.ptr, 0, sizeof(uint32_t), s));
        k_query_shapes<<<blocks_for(n, 64), 64, 0, s>>>(qb.shapes.ptr, n, view_of(ctx, 0), view_of(ctx, 1), view_of(ctx, 2),
                                                        ctx->d, qb.out.ptr, static_cast<uint32_t>(qb.out.capacity),
                                                        qb.counter.ptr);
        C2D_CUDA(ctx, cudaMemcpyAsync(&count, qb.counter.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
        C2D_CUDA(ctx, cudaGetLastError());
        if (count <= qb.out.capacity)
            break;
        C2D_CUDA(ctx, qb.out.reserve(count));
    }
    if (count)
    {
        if (!qb.hostOut.reserve(count))
            return fail(ctx, "pinned host allocation failed");
        C2D_CUDA(ctx, cudaMemcpyAsync(qb.hostOut.data, qb.out.ptr, static_cast<size_t>(count) * sizeof(PairOut),
                                      cudaMemcpyDeviceToHost, s));
        C2D_CUDA(ctx, cudaStreamSynchronize(s));
    }
    if (out_records)
        *out_records = reinterpret_cast<const c2d_pair_record*>(qb.hostOut.data);
    if (out_count)
        *out_count = count;
    return 0;
}

int c2d_gpu_pairs_colliding(c2d_gpu_ctx* ctx, uint32_t n, const uint32_t* shapeA, const uint32_t* shapeB,
                            uint8_t* out_colliding)
{
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    if (n == 0)
        return 0;
    if (!shapeA || !shapeB || !out_colliding)
        return fail(ctx, "c2d_gpu_pairs_colliding: bad argument");
    for (uint32_t i = 0; i < n; ++i)
        if (shapeA[i] >= ctx->hostMeta.size() || shapeB[i] >= ctx->hostMeta.size())
            return fail(ctx, "c2d_gpu_pairs_colliding: unknown shape");
    if (int rc = flush(ctx))
        return rc;
    QueryBuffers& qb = ctx->query;
    C2D_CUDA(ctx, qb.shapes.reserve(static_cast<size_t>(n) * 2));
    C2D_CUDA(ctx, qb.flags.reserve(n));
    C2D_CUDA(ctx, cudaMemcpyAsync(qb.shapes.ptr, shapeA, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    C2D_CUDA(ctx, cudaMemcpyAsync(qb.shapes.ptr + n, shapeB, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    k_pairs_colliding<<<blocks_for(n, 64), 64, 0, s>>>(qb.shapes.ptr, qb.shapes.ptr + n, n, ctx->d, qb.flags.ptr);
    C2D_CUDA(ctx, cudaMemcpyAsync(out_colliding, qb.flags.ptr, n, cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    C2D_CUDA(ctx, cudaGetLastError());
    return 0;
}

} // extern "C"
