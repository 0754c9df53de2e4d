// c2d_slab.cuh — Morton-range slabs of the Dynamic tree for the multi-GPU path.
//
// Shape state is replicated on every rank, so ownership needs no communication: every rank computes the same Morton
// keys from the same leaf boxes and the same splitters from the same key histogram.
//
//   k_morton_hist    keys of ALL members + coarse histogram (top 12 key bits = 64 x 64 world cells)
//   k_slab_split1/2  balanced splitters: coarse bin, then a 12-bit refinement histogram (k_slab_hist2) inside each
//                    splitter bin -> splitter keys with 24 significant bits; slab r = keys in [split[r], split[r+1])
//   k_slab_mark      the 256 x 256 world cells touched by the leaf boxes of the shapes this rank OWNS
//   k_slab_select    this rank's leaves: owned shapes + halo = shapes of HIGHER ranks whose leaf box touches a marked
//                    cell (two overlapping boxes share a point, hence a cell).  Owner rule: a pair is emitted by the rank
//                    that owns its first shape in global (key, slab) order; shapes of lower ranks come earlier in that
//                    order, so they never have to be in the halo.  Halo leaves carry kLeafGhost (bit 30) in the sorted value array and in
//                    their leaf link.
#pragma once

#include "c2d_internal.hpp"

namespace c2d_dev
{

constexpr uint32_t kSlabCoarseBins = 4096;
constexpr int kSlabMaxWorld = 32;
constexpr uint32_t kSlabCells = 256; // per axis
constexpr int kSlabMaxCellSpan = 16; // boxes wider than this many cells are treated as touching everything

// layout of the uint32 split buffer (split[world] is implied: 2^32): [0..31] splitter keys (split[0] = 0), [32..63] coarse bin of splitter j,
// [64..95] residual rank inside that bin, [96] "some owned box spans too many cells" flag.
constexpr int kSplitKeys = 0, kSplitBin = 32, kSplitRest = 64, kSplitHuge = 96, kSplitWords = 128;

__device__ __forceinline__ uint32_t morton_of_box(const float4 b, float minx, float miny, float sx, float sy)
{
    const float cx = 0.5f * (b.x + b.z), cy = 0.5f * (b.y + b.w);
    const uint32_t qx = static_cast<uint32_t>(fminf(fmaxf((cx - minx) * sx, 0.0f), 65535.0f));
    const uint32_t qy = static_cast<uint32_t>(fminf(fmaxf((cy - miny) * sy, 0.0f), 65535.0f));
    return spread16(qx) | (spread16(qy) << 1);
}

__global__ void __launch_bounds__(256) k_morton_hist(const uint32_t* __restrict__ members, uint32_t n,
                                                     const float4* __restrict__ fat, const uint32_t* __restrict__ bounds4,
                                                     uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                     uint32_t* __restrict__ hist /* null: keys only */)
{
    __shared__ uint32_t sh[kSlabCoarseBins];
    if (hist)
    {
        for (uint32_t i = threadIdx.x; i < kSlabCoarseBins; i += blockDim.x)
            sh[i] = 0;
        __syncthreads();
    }
    const float minx = ordered_to_float(bounds4[0]), miny = ordered_to_float(bounds4[1]);
    const float ex = ordered_to_float(bounds4[2]) - minx, ey = ordered_to_float(bounds4[3]) - miny;
    const float sx = ex > 0.0f ? 65535.0f / ex : 0.0f;
    const float sy = ey > 0.0f ? 65535.0f / ey : 0.0f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t shape = members[i];
        const uint32_t key = morton_of_box(fat[shape], minx, miny, sx, sy);
        keys[i] = key;
        vals[i] = shape;
        if (hist)
            atomicAdd(&sh[key >> 20], 1u);
    }
    if (hist)
    {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < kSlabCoarseBins; i += blockDim.x)
            if (sh[i])
                atomicAdd(&hist[i], sh[i]);
    }
}

// Block-wide inclusive scan of 4096 counters held 4 per thread (1024 threads); returns the exclusive prefix of this
// thread's first counter.
__device__ __forceinline__ uint32_t scan4096(const uint32_t v[4], uint32_t* warpSums)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
    if (lane == 31)
        warpSums[warp] = inc;
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
    const uint32_t before = (warp ? warpSums[warp - 1] : 0u) + inc - mine;
    __syncthreads();
    return before;
}

// splitter j (1 <= j < world): the coarse bin in which the cumulative count crosses n * j / world
__global__ void __launch_bounds__(1024) k_slab_split1(const uint32_t* __restrict__ hist, uint32_t n, uint32_t world,
                                                      uint32_t* __restrict__ split)
{
    __shared__ uint32_t warpSums[32];
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        v[k] = hist[threadIdx.x * 4 + k];
    uint32_t run = scan4096(v, warpSums);
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const uint32_t lo = run, hi = run + v[k]; // this bin holds global ranks [lo, hi)
        for (uint32_t j = 1; j < world; ++j)
        {
            const uint32_t target = static_cast<uint32_t>(static_cast<uint64_t>(n) * j / world);
            if (target >= lo && target < hi)
            {
                split[kSplitBin + j] = threadIdx.x * 4 + k;
                split[kSplitRest + j] = target - lo;
            }
        }
        run = hi;
    }
    if (threadIdx.x == 0)
    {
        split[kSplitKeys] = 0;
        split[kSplitHuge] = 0;
    }
}

__global__ void __launch_bounds__(256) k_slab_hist2(const uint32_t* __restrict__ keys, uint32_t n, uint32_t world,
                                                    const uint32_t* __restrict__ split, uint32_t* __restrict__ hist2)
{
    __shared__ uint32_t bins[kSlabMaxWorld];
    if (threadIdx.x < kSlabMaxWorld)
        bins[threadIdx.x] = threadIdx.x >= 1 && threadIdx.x < world ? split[kSplitBin + threadIdx.x] : 0xffffffffu;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t key = keys[i];
        const uint32_t coarse = key >> 20;
        for (uint32_t j = 1; j < world; ++j)
            if (bins[j] == coarse)
                atomicAdd(&hist2[static_cast<size_t>(j) * kSlabCoarseBins + ((key >> 8) & 0xfffu)], 1u);
    }
}

// one block per splitter: the sub-bin in which the residual rank falls -> a splitter key with 24 significant bits
__global__ void __launch_bounds__(1024) k_slab_split2(const uint32_t* __restrict__ hist2, uint32_t world,
                                                      uint32_t* __restrict__ split)
{
    __shared__ uint32_t warpSums[32];
    const uint32_t j = blockIdx.x + 1;
    if (j >= world)
        return;
    const uint32_t* h = hist2 + static_cast<size_t>(j) * kSlabCoarseBins;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        v[k] = h[threadIdx.x * 4 + k];
    uint32_t run = scan4096(v, warpSums);
    const uint32_t rest = split[kSplitRest + j], bin = split[kSplitBin + j];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        if (rest >= run && rest < run + v[k])
            split[kSplitKeys + j] = (bin << 20) | ((threadIdx.x * 4 + k) << 8);
        run += v[k];
    }
}

__device__ __forceinline__ uint32_t slab_owner(uint32_t key, const uint32_t* splitKeys, uint32_t world)
{
    uint32_t o = 0;
    for (uint32_t j = 1; j < world; ++j)
        o += key >= splitKeys[j] ? 1u : 0u;
    return o;
}

struct CellRange
{
    int x0, y0, x1, y1;
};

__device__ __forceinline__ CellRange cells_of_box(const float4 b, float minx, float miny, float sx, float sy)
{
    // same quantisation as the Morton key (clamping is monotone: two boxes that share a point share a cell)
    CellRange r;
    r.x0 = static_cast<int>(fminf(fmaxf((b.x - minx) * sx, 0.0f), 65535.0f)) >> 8;
    r.y0 = static_cast<int>(fminf(fmaxf((b.y - miny) * sy, 0.0f), 65535.0f)) >> 8;
    r.x1 = static_cast<int>(fminf(fmaxf((b.z - minx) * sx, 0.0f), 65535.0f)) >> 8;
    r.y1 = static_cast<int>(fminf(fmaxf((b.w - miny) * sy, 0.0f), 65535.0f)) >> 8;
    return r;
}

__global__ void __launch_bounds__(256) k_slab_mark(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                   uint32_t n, const float4* __restrict__ fat,
                                                   const uint32_t* __restrict__ bounds4, uint32_t* split, uint32_t rank,
                                                   uint32_t world, uint8_t* __restrict__ cellMask)
{
    __shared__ uint32_t sk[kSlabMaxWorld];
    if (threadIdx.x < kSlabMaxWorld)
        sk[threadIdx.x] = threadIdx.x < world ? split[kSplitKeys + threadIdx.x] : 0xffffffffu;
    __syncthreads();
    const float minx = ordered_to_float(bounds4[0]), miny = ordered_to_float(bounds4[1]);
    const float ex = ordered_to_float(bounds4[2]) - minx, ey = ordered_to_float(bounds4[3]) - miny;
    const float sx = ex > 0.0f ? 65535.0f / ex : 0.0f;
    const float sy = ey > 0.0f ? 65535.0f / ey : 0.0f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        if (slab_owner(keys[i], sk, world) != rank)
            continue;
        const CellRange c = cells_of_box(fat[vals[i]], minx, miny, sx, sy);
        if (c.x1 - c.x0 >= kSlabMaxCellSpan || c.y1 - c.y0 >= kSlabMaxCellSpan)
        {
            split[kSplitHuge] = 1; // an owned leaf box spans the world: every later shape joins the halo
            continue;
        }
        for (int y = c.y0; y <= c.y1; ++y)
            for (int x = c.x0; x <= c.x1; ++x)
                if (!cellMask[y * kSlabCells + x])
                    cellMask[y * kSlabCells + x] = 1;
    }
}

// counters: [0] leaves selected, [1] of them owned
__global__ void __launch_bounds__(256) k_slab_select(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                     uint32_t n, const float4* __restrict__ fat,
                                                     const uint32_t* __restrict__ bounds4, const uint32_t* __restrict__ split,
                                                     uint32_t rank, uint32_t world, const uint8_t* __restrict__ cellMask,
                                                     uint32_t* __restrict__ outKeys, uint32_t* __restrict__ outVals,
                                                     uint32_t* counters)
{
    __shared__ uint32_t sk[kSlabMaxWorld];
    if (threadIdx.x < kSlabMaxWorld)
        sk[threadIdx.x] = threadIdx.x < world ? split[kSplitKeys + threadIdx.x] : 0xffffffffu;
    __syncthreads();
    const float minx = ordered_to_float(bounds4[0]), miny = ordered_to_float(bounds4[1]);
    const float ex = ordered_to_float(bounds4[2]) - minx, ey = ordered_to_float(bounds4[3]) - miny;
    const float sx = ex > 0.0f ? 65535.0f / ex : 0.0f;
    const float sy = ey > 0.0f ? 65535.0f / ey : 0.0f;
    const bool huge = split[kSplitHuge] != 0;
    const int lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x)
    {
        const uint32_t i = base + lane;
        bool take = false, owned = false;
        uint32_t key = 0, shape = 0;
        if (i < n)
        {
            key = keys[i];
            shape = vals[i];
            const uint32_t owner = slab_owner(key, sk, world);
            owned = owner == rank;
            take = owned;
            if (owner > rank)
            {
                const CellRange c = cells_of_box(fat[shape], minx, miny, sx, sy);
                take = huge || c.x1 - c.x0 >= kSlabMaxCellSpan || c.y1 - c.y0 >= kSlabMaxCellSpan;
                for (int y = c.y0; y <= c.y1 && !take; ++y)
                    for (int x = c.x0; x <= c.x1 && !take; ++x)
                        take = cellMask[y * kSlabCells + x] != 0;
            }
        }
        const uint32_t tb = __ballot_sync(0xffffffffu, take), ob = __ballot_sync(0xffffffffu, owned);
        if (tb == 0)
            continue;
        uint32_t slot = 0;
        if (lane == 0)
        {
            slot = atomicAdd(&counters[0], static_cast<uint32_t>(__popc(tb)));
            if (ob)
                atomicAdd(&counters[1], static_cast<uint32_t>(__popc(ob)));
        }
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (take)
        {
            const uint32_t pos = slot + __popc(tb & ((1u << lane) - 1u));
            outKeys[pos] = key;
            outVals[pos] = owned ? shape : (shape | kLeafGhost);
        }
    }
}

} // namespace c2d_dev
