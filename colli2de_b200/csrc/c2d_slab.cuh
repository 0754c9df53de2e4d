// c2d_slab.cuh — Morton-range slabs of the Dynamic tree for the multi-GPU path.
//
// Shape state is replicated on every rank, so ownership needs no communication: every rank computes the same Morton
// keys from the same leaf boxes and the same (stable, deterministic) sort, and rank r owns the sorted positions
// [n r / world, n (r+1) / world) - balanced by construction whatever the density.  The order is recomputed when the
// membership changes or every `rebuild period` frames (like the LBVH topology of the single-GPU path); in between the
// OWNERSHIP is frozen, which is all the owner rule needs, while the halo is exact every frame:
//
//   k_slab_mark      the 256 x 256 world cells touched by the CURRENT leaf boxes of the shapes this rank owns
//                    (per-block bit mask in shared memory, merged into 2048 global words)
//   k_slab_select    halo = shapes of HIGHER ranks whose current leaf box touches a marked cell (two overlapping boxes share
//                    a point, hence a cell), compacted into a list whose length stays on the device
//
// The rank's tree holds its owned leaves only (refit every frame, relinked at a rebalance); the halo shapes are QUERIES
// against it (k_traverse_sync<cross, ordered> sized by the device-side count), so nothing is sorted or built for them and
// the host never waits for the selection.  Owner rule: a Dynamic x Dynamic pair is emitted by the rank that owns its
// earlier shape in the global order; shapes of lower ranks come earlier, so they never have to be in a halo.
#pragma once

#include "c2d_internal.hpp"

namespace c2d_dev
{

constexpr int kSlabMaxWorld = 32;
constexpr uint32_t kSlabCells = 256; // per axis
constexpr uint32_t kSlabMaskWords = kSlabCells * kSlabCells / 32;
constexpr int kSlabMaxCellSpan = 16; // boxes wider than this many cells are treated as touching everything

// counters of the selection: [0] halo shapes selected, [1] "an owned box spans too many cells" flag
constexpr int kSlabHaloCount = 0, kSlabHuge = 1, kSlabCounterWords = 4;

__global__ void k_slab_reset_bounds(uint32_t* bounds4)
{
    if (threadIdx.x < 4)
        bounds4[threadIdx.x] = threadIdx.x < 2 ? 0xffffffffu : 0u;
}

struct CellRange
{
    int x0, y0, x1, y1;
};

struct CellFrame
{
    float minx, miny, sx, sy;
};

__device__ __forceinline__ CellFrame cell_frame(const uint32_t* __restrict__ bounds4)
{
    CellFrame f;
    f.minx = ordered_to_float(bounds4[0]);
    f.miny = ordered_to_float(bounds4[1]);
    const float ex = ordered_to_float(bounds4[2]) - f.minx, ey = ordered_to_float(bounds4[3]) - f.miny;
    f.sx = ex > 0.0f ? 65535.0f / ex : 0.0f;
    f.sy = ey > 0.0f ? 65535.0f / ey : 0.0f;
    return f;
}

__device__ __forceinline__ CellRange cells_of_box(const float4 b, const CellFrame f)
{
    // the quantisation of the Morton key; the clamp is monotone, so two boxes that share a point share a cell even when
    // they have left the bounds the frame was computed from
    CellRange r;
    r.x0 = static_cast<int>(fminf(fmaxf((b.x - f.minx) * f.sx, 0.0f), 65535.0f)) >> 8;
    r.y0 = static_cast<int>(fminf(fmaxf((b.y - f.miny) * f.sy, 0.0f), 65535.0f)) >> 8;
    r.x1 = static_cast<int>(fminf(fmaxf((b.z - f.minx) * f.sx, 0.0f), 65535.0f)) >> 8;
    r.y1 = static_cast<int>(fminf(fmaxf((b.w - f.miny) * f.sy, 0.0f), 65535.0f)) >> 8;
    return r;
}

// owned = sorted shape ids of this rank's slab (spatially coherent: a block touches few mask words)
__global__ void __launch_bounds__(256) k_slab_mark(const uint32_t* __restrict__ owned, uint32_t m,
                                                   const float4* __restrict__ fat, const uint32_t* __restrict__ bounds4,
                                                   uint32_t* __restrict__ cellBits, uint32_t* __restrict__ counters)
{
    __shared__ uint32_t bits[kSlabMaskWords];
    for (uint32_t w = threadIdx.x; w < kSlabMaskWords; w += blockDim.x)
        bits[w] = 0;
    __syncthreads();
    const CellFrame f = cell_frame(bounds4);
    const uint32_t per = (m + gridDim.x - 1) / gridDim.x; // a contiguous run of the sorted order per block
    const uint32_t lo = blockIdx.x * per, hi = min(m, lo + per);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
    {
        const CellRange c = cells_of_box(fat[owned[i] & kLeafIdMask], f);
        if (c.x1 - c.x0 >= kSlabMaxCellSpan || c.y1 - c.y0 >= kSlabMaxCellSpan)
        {
            counters[kSlabHuge] = 1; // an owned leaf box spans the world: every later shape joins the halo
            continue;
        }
        for (int y = c.y0; y <= c.y1; ++y)
            for (int x = c.x0; x <= c.x1; ++x)
            {
                const uint32_t cell = static_cast<uint32_t>(y) * kSlabCells + static_cast<uint32_t>(x);
                const uint32_t bit = 1u << (cell & 31u);
                if (!(bits[cell >> 5] & bit))
                    atomicOr(&bits[cell >> 5], bit);
            }
    }
    __syncthreads();
    for (uint32_t w = threadIdx.x; w < kSlabMaskWords; w += blockDim.x)
    {
        const uint32_t b = bits[w];
        if (b && (cellBits[w] & b) != b)
            atomicOr(&cellBits[w], b);
    }
}

// later = the sorted shape ids that follow this rank's slab (the shapes of higher ranks)
__global__ void __launch_bounds__(256) k_slab_select(const uint32_t* __restrict__ later, uint32_t count,
                                                     const float4* __restrict__ fat, const uint32_t* __restrict__ bounds4,
                                                     const uint32_t* __restrict__ cellBits, uint32_t* __restrict__ halo,
                                                     uint32_t* counters)
{
    __shared__ uint32_t bits[kSlabMaskWords];
    for (uint32_t w = threadIdx.x; w < kSlabMaskWords; w += blockDim.x)
        bits[w] = cellBits[w];
    __syncthreads();
    const CellFrame f = cell_frame(bounds4);
    const bool huge = counters[kSlabHuge] != 0;
    const int lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < count; base += gridDim.x * blockDim.x)
    {
        const uint32_t i = base + lane;
        bool take = false;
        uint32_t shape = 0;
        if (i < count)
        {
            shape = later[i] & kLeafIdMask;
            const CellRange c = cells_of_box(fat[shape], f);
            take = huge || c.x1 - c.x0 >= kSlabMaxCellSpan || c.y1 - c.y0 >= kSlabMaxCellSpan;
            for (int y = c.y0; y <= c.y1 && !take; ++y)
                for (int x = c.x0; x <= c.x1 && !take; ++x)
                {
                    const uint32_t cell = static_cast<uint32_t>(y) * kSlabCells + static_cast<uint32_t>(x);
                    take = (bits[cell >> 5] >> (cell & 31u)) & 1u;
                }
        }
        const uint32_t tb = __ballot_sync(0xffffffffu, take);
        if (tb == 0)
            continue;
        uint32_t slot = 0;
        if (lane == 0)
            slot = atomicAdd(&counters[kSlabHaloCount], static_cast<uint32_t>(__popc(tb)));
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (take)
            halo[slot + __popc(tb & ((1u << lane) - 1u))] = shape;
    }
}

} // namespace c2d_dev
