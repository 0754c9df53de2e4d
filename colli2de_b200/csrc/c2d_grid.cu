// c2d_grid.cu — uniform-grid sort-and-sweep broadphase: the B200 replacement of PartitioningMethod::Grid.
//
// The reference's BroadPhaseTree (reference internal/data_structures/BroadPhaseTree.hpp:71-88, 188-224, 374-400)
// keeps one DynamicBVH per grid cell, registers a proxy in every cell its box touches and runs
// findAllCollisions per cell — reporting a pair once per shared cell, and as (a,b) AND (b,a) when two cells see
// it in different orders.  Here the same candidate SET is produced without any tree:
//   1. extent of all leaf (fat) boxes -> grid origin; cell = max(cellSize, extent / 2046) (<= 2047 cells per axis);
//   2. every shape emits one entry per cell its leaf box touches: key = cy:11 | cx:11 | qminx:10 (the box's
//      min.x quantised inside the cell), value = shape id  (count -> scan -> fill);
//   3. one radix sort of the entries (c2d_sort.cu): cells become contiguous and, inside a cell, entries are
//      ordered by min.x;
//   4. sweep: entry i scans forward while the cell matches and qminx_j <= qmaxx_i, testing the exact boxes;
//      a pair is emitted only by the cell that contains (max(min.x), max(min.y)) of the two boxes — that point
//      lies in both boxes, so both shapes are registered there and every pair comes out exactly once.
// The candidate test is the reference's: leaf boxes intersect (non-strict, AABB.cpp:34-37) and
// (categoryA & categoryB) != 0 (DynamicBVH.hpp:955); Static x Static is never tested (Registry.hpp:487-534).
// The pairs go to the same binning + narrowphase kernels as the LBVH path.
#include "c2d_ctx.hpp"
#include "c2d_math.cuh"

#include <float.h>

using namespace c2d_dev;
using namespace c2d_host;

namespace
{

struct GridParams // device-resident, written by k_grid_setup
{
    float minx, miny, cell, invCell;
};

__device__ __forceinline__ uint32_t f2o(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_grid_reset(uint32_t* ext)
{
    if (threadIdx.x < 2)
        ext[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 4)
        ext[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) k_grid_extent(const uint32_t* __restrict__ members, uint32_t n,
                                                     const float4* __restrict__ fat, uint32_t* ext)
{
    float minx = FLT_MAX, miny = FLT_MAX, maxx = -FLT_MAX, maxy = -FLT_MAX;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 b = fat[members[i]];
        minx = fminf(minx, b.x);
        miny = fminf(miny, b.y);
        maxx = fmaxf(maxx, b.z);
        maxy = fmaxf(maxy, b.w);
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
        atomicMin(&ext[0], f2o(minx));
        atomicMin(&ext[1], f2o(miny));
        atomicMax(&ext[2], f2o(maxx));
        atomicMax(&ext[3], f2o(maxy));
    }
}

__global__ void k_grid_setup(const uint32_t* ext, float cellSize, GridParams* gp)
{
    const float minx = o2f(ext[0]), miny = o2f(ext[1]), maxx = o2f(ext[2]), maxy = o2f(ext[3]);
    const float extent = fmaxf(maxx - minx, maxy - miny);
    float cell = cellSize;
    if (!(extent / cell < 2046.0f))
        cell = extent / 2046.0f;
    gp->minx = minx;
    gp->miny = miny;
    gp->cell = cell;
    gp->invCell = 1.0f / cell;
}

__device__ __forceinline__ int cell_of(float x, float origin, float invCell)
{
    const int c = static_cast<int>(floorf((x - origin) * invCell));
    return c < 0 ? 0 : (c > 2047 ? 2047 : c);
}

__global__ void __launch_bounds__(256) k_grid_count(const uint32_t* __restrict__ members, uint32_t n,
                                                    const float4* __restrict__ fat, const GridParams* gp,
                                                    uint32_t* __restrict__ counts)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const GridParams g = *gp;
    const float4 b = fat[members[i]];
    const int cx0 = cell_of(b.x, g.minx, g.invCell), cx1 = cell_of(b.z, g.minx, g.invCell);
    const int cy0 = cell_of(b.y, g.miny, g.invCell), cy1 = cell_of(b.w, g.miny, g.invCell);
    counts[i] = static_cast<uint32_t>((cx1 - cx0 + 1) * (cy1 - cy0 + 1));
}

__device__ __forceinline__ uint32_t quant_in_cell(float x, float origin, int c, const GridParams& g)
{
    // position of x inside cell c in 1/1024ths, clamped: boxes starting left of the cell sort first,
    // boxes ending right of it keep the sweep open to the end of the cell
    const float rel = ((x - origin) * g.invCell - static_cast<float>(c)) * 1024.0f;
    const int q = static_cast<int>(floorf(rel));
    return static_cast<uint32_t>(q < 0 ? 0 : (q > 1023 ? 1023 : q));
}

__global__ void __launch_bounds__(256) k_grid_fill(const uint32_t* __restrict__ members, uint32_t n,
                                                   const float4* __restrict__ fat, const GridParams* gp,
                                                   const uint32_t* __restrict__ offsets, uint32_t* __restrict__ keys,
                                                   uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const GridParams g = *gp;
    const uint32_t shape = members[i];
    const float4 b = fat[shape];
    const int cx0 = cell_of(b.x, g.minx, g.invCell), cx1 = cell_of(b.z, g.minx, g.invCell);
    const int cy0 = cell_of(b.y, g.miny, g.invCell), cy1 = cell_of(b.w, g.miny, g.invCell);
    uint32_t e = offsets[i];
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx)
        {
            keys[e] = (static_cast<uint32_t>(cy) << 21) | (static_cast<uint32_t>(cx) << 10) |
                      quant_in_cell(b.x, g.minx, cx, g);
            vals[e] = shape;
            ++e;
        }
}

__global__ void __launch_bounds__(256) k_grid_gather(const uint32_t* __restrict__ vals, uint32_t n,
                                                     const float4* __restrict__ fat, const uint64_t* __restrict__ category,
                                                     const uint32_t* __restrict__ meta, float4* __restrict__ sbox,
                                                     uint64_t* __restrict__ scat, uint32_t* __restrict__ sbody)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t s = vals[i];
    sbox[i] = fat[s];
    scat[i] = category[s];
    sbody[i] = meta_body(meta[s]);
}

// One entry per thread; the forward scan of a warp's 32 entries reads a sliding window of the sorted arrays.
__global__ void __launch_bounds__(128)
    k_grid_pairs(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n,
                 const float4* __restrict__ sbox, const uint64_t* __restrict__ scat, const uint32_t* __restrict__ sbody,
                 const GridParams* gp, uint32_t lo, uint32_t hi, uint2* __restrict__ out, uint32_t capacity,
                 uint32_t* counter)
{
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi || i >= n)
        return;
    const GridParams g = *gp;
    const uint32_t keyI = keys[i];
    const uint32_t cellI = keyI >> 10;
    const int cx = static_cast<int>((keyI >> 10) & 0x7ffu), cy = static_cast<int>(keyI >> 21);
    const float4 bi = sbox[i];
    const uint64_t ci = scat[i];
    const uint32_t bodyI = sbody[i];
    const uint32_t si = vals[i];
    const uint32_t qmax = quant_in_cell(bi.z, g.minx, cx, g);
    for (uint32_t j = i + 1; j < n; ++j)
    {
        const uint32_t keyJ = keys[j];
        if ((keyJ >> 10) != cellI || (keyJ & 0x3ffu) > qmax)
            break;
        const uint32_t bodyJ = sbody[j];
        if (bodyI == 0u && bodyJ == 0u)
            continue; // Static x Static is never queried
        if ((ci & scat[j]) == 0)
            continue;
        const float4 bj = sbox[j];
        if (!box_overlap(bi, bj))
            continue;
        // emitted only by the cell holding the corner (max of the min corners) of the intersection
        if (cell_of(fmaxf(bi.x, bj.x), g.minx, g.invCell) != cx || cell_of(fmaxf(bi.y, bj.y), g.miny, g.invCell) != cy)
            continue;
        const uint32_t sj = vals[j];
        // orientation: cross-tree pairs have A on the Dynamic (vs Static) / Bullet side, same-tree pairs shapeA < shapeB
        uint32_t a = si, b = sj;
        if (bodyI == bodyJ ? (si > sj) : (bodyJ > bodyI))
        {
            a = sj;
            b = si;
        }
        const uint32_t idx = atomicAdd(counter, 1u);
        if (idx < capacity)
            out[idx] = make_uint2(a, b);
    }
}

} // namespace

namespace c2d_host
{

// Fills ctx->cand / dScratch->candCount with the candidate pairs of all five groups.  Returns 0 or an error code.
int grid_broadphase(c2d_gpu_ctx* ctx)
{
    cudaStream_t s = ctx->stream;
    GridBuffers& gb = ctx->grid;
    uint32_t total = 0;
    for (int b = 0; b < 3; ++b)
        total += static_cast<uint32_t>(ctx->members[b].size());
    if (total == 0)
        return 0;
    for (int b = 0; b < 3; ++b)
        if (int rc = upload_members(ctx, b))
            return rc;
    C2D_CUDA(ctx, gb.ext.reserve(4));
    C2D_CUDA(ctx, gb.params.reserve(sizeof(GridParams)));
    C2D_CUDA(ctx, gb.counts.reserve(total));
    C2D_CUDA(ctx, gb.totals.reserve(1));
    GridParams* gp = reinterpret_cast<GridParams*>(gb.params.ptr);

    k_grid_reset<<<1, 32, 0, s>>>(gb.ext.ptr);
    ++ctx->launches;
    for (int b = 0; b < 3; ++b)
    {
        const uint32_t n = static_cast<uint32_t>(ctx->members[b].size());
        if (!n)
            continue;
        uint32_t blocks = blocks_for(n, 256);
        if (blocks > 148 * 8)
            blocks = 148 * 8;
        k_grid_extent<<<blocks, 256, 0, s>>>(ctx->tree[b].members.ptr, n, ctx->d.fat, gb.ext.ptr);
        ++ctx->launches;
    }
    k_grid_setup<<<1, 1, 0, s>>>(gb.ext.ptr, static_cast<float>(ctx->cellSize), gp);
    ++ctx->launches;
    uint32_t first = 0;
    for (int b = 0; b < 3; ++b)
    {
        const uint32_t n = static_cast<uint32_t>(ctx->members[b].size());
        if (!n)
            continue;
        k_grid_count<<<blocks_for(n, 256), 256, 0, s>>>(ctx->tree[b].members.ptr, n, ctx->d.fat, gp, gb.counts.ptr + first);
        ++ctx->launches;
        first += n;
    }
    if (int rc = exclusive_scan_u32(ctx, gb.counts.ptr, total, gb.sums, gb.totals.ptr))
        return rc;
    uint32_t entries = 0;
    C2D_CUDA(ctx, cudaMemcpyAsync(&entries, gb.totals.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    C2D_CUDA(ctx, cudaStreamSynchronize(s));
    if (entries == 0)
        return 0;
    C2D_CUDA(ctx, gb.keys.reserve(entries));
    C2D_CUDA(ctx, gb.vals.reserve(entries));
    C2D_CUDA(ctx, gb.keysTmp.reserve(entries));
    C2D_CUDA(ctx, gb.valsTmp.reserve(entries));
    C2D_CUDA(ctx, gb.sbox.reserve(entries));
    C2D_CUDA(ctx, gb.scat.reserve(entries));
    C2D_CUDA(ctx, gb.sbody.reserve(entries));
    C2D_CUDA(ctx, gb.sortWs.reserve(sort_workspace_bytes(entries, 4)));
    first = 0;
    for (int b = 0; b < 3; ++b)
    {
        const uint32_t n = static_cast<uint32_t>(ctx->members[b].size());
        if (!n)
            continue;
        k_grid_fill<<<blocks_for(n, 256), 256, 0, s>>>(ctx->tree[b].members.ptr, n, ctx->d.fat, gp, gb.counts.ptr + first,
                                                       gb.keys.ptr, gb.vals.ptr);
        ++ctx->launches;
        first += n;
    }
    ctx->launches += radix_sort_pairs(s, gb.keys.ptr, gb.vals.ptr, gb.keysTmp.ptr, gb.valsTmp.ptr, entries, 4, gb.sortWs.ptr);
    k_grid_gather<<<blocks_for(entries, 256), 256, 0, s>>>(gb.vals.ptr, entries, ctx->d.fat, ctx->d.category, ctx->d.meta,
                                                          gb.sbox.ptr, gb.scat.ptr, gb.sbody.ptr);
    // multi-GPU: this context owns the entries of slab [lo, hi) of the cell-sorted list (a Morton-like range of cells)
    const uint32_t lo = static_cast<uint32_t>(static_cast<uint64_t>(entries) * ctx->shardRank / ctx->shardWorld);
    const uint32_t hi = static_cast<uint32_t>(static_cast<uint64_t>(entries) * (ctx->shardRank + 1) / ctx->shardWorld);
    if (hi > lo)
        k_grid_pairs<<<blocks_for(hi - lo, 128), 128, 0, s>>>(gb.keys.ptr, gb.vals.ptr, entries, gb.sbox.ptr, gb.scat.ptr,
                                                            gb.sbody.ptr, gp, lo, hi, ctx->cand.ptr,
                                                            static_cast<uint32_t>(ctx->cand.capacity),
                                                            &ctx->dScratch->candCount);
    ctx->launches += 2;
    C2D_CUDA(ctx, cudaGetLastError());
    return 0;
}

} // namespace c2d_host
