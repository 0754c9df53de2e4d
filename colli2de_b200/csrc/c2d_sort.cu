// c2d_sort.cu — CUB-free LSD radix sort of (uint32 key, uint32 value) pairs for sm_100a.
//
// Used for (a) Morton ordering of the leaf boxes before the LBVH build and (b) ordering candidate /
// result pairs; it replaces the std::sort calls of the reference's getCollidingPairs
// (reference include/colli2de/Registry.hpp:495,504,514,524,533) and has no other analogue there.
//
// Design ("onesweep"): one upfront kernel builds the global digit histograms of all passes; each pass
// is then ONE kernel in which a tile (4096 keys) ranks its keys locally (warp match-any ranking, stable),
// publishes its per-digit counts and obtains its global offsets by decoupled look-back over the previous
// tiles' published counts — keys and values are read once and written once per pass.  Tiles take their
// index from an atomic ticket so that every tile a block waits on is already resident (forward progress).
#include "c2d_internal.hpp"

namespace c2d_dev
{

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems; // 4096
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRadix = 256;

constexpr uint32_t kFlagMask = 0xC0000000u;
constexpr uint32_t kFlagAggregate = 0x40000000u;
constexpr uint32_t kFlagPrefix = 0x80000000u;
constexpr uint32_t kValueMask = 0x3FFFFFFFu;

// Global digit histograms for up to 4 passes (8-bit digits of a 32-bit key).
__global__ void __launch_bounds__(256) k_sort_histogram(const uint32_t* __restrict__ keys, uint32_t n,
                                                        uint32_t* __restrict__ ghist, int passes)
{
    __shared__ uint32_t sh[4 * kRadix];
    for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x)
        sh[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t k = keys[i];
        for (int p = 0; p < passes; ++p)
            atomicAdd(&sh[p * kRadix + ((k >> (8 * p)) & 0xffu)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x)
        if (sh[i])
            atomicAdd(&ghist[i], sh[i]);
}

// One radix pass.  state: numTiles x 256 status words (zeroed), ticket: zeroed counter.
__global__ void __launch_bounds__(kSortThreads)
    k_sort_pass(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, uint32_t n, int shift,
                const uint32_t* __restrict__ ghist, volatile uint32_t* state, uint32_t* ticket)
{
    __shared__ uint32_t wcount[kSortWarps][kRadix];
    __shared__ uint32_t base[kRadix];
    __shared__ uint32_t scanTmp[kSortWarps];
    __shared__ uint32_t tileShared;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (tid == 0)
        tileShared = atomicAdd(ticket, 1u);
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads)
        (&wcount[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = tileShared;
    const uint32_t tileBase = tile * kSortTile;

    // exclusive scan of the global digit histogram -> first output index of each digit (thread d = digit d)
    uint32_t gbase;
    {
        const uint32_t v = ghist[tid];
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (lane == 31)
            scanTmp[warp] = incl;
        __syncthreads();
        uint32_t warpOff = 0;
        for (int w = 0; w < warp; ++w)
            warpOff += scanTmp[w];
        gbase = warpOff + incl - v;
    }

    // load + rank
    uint32_t key[kSortItems];
    uint32_t rank[kSortItems];
    const uint32_t warpBase = tileBase + warp * (32 * kSortItems);
    const uint32_t ltMask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k)
    {
        const uint32_t idx = warpBase + k * 32 + lane;
        key[k] = idx < n ? keysIn[idx] : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < kSortItems; ++k)
    {
        const uint32_t d = (key[k] >> shift) & 0xffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader)
        {
            old = wcount[warp][d];
            wcount[warp][d] = old + __popc(peers);
        }
        __syncwarp();
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = old + __popc(peers & ltMask);
    }
    __syncthreads();

    // per-digit: prefix over warps, tile total, decoupled look-back
    {
        const int d = tid;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w)
        {
            const uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
        const uint32_t total = run;
        volatile uint32_t* myState = state + static_cast<size_t>(tile) * kRadix + d;
        uint32_t exclusive = 0;
        if (tile == 0)
        {
            *myState = kFlagPrefix | total;
        }
        else
        {
            *myState = kFlagAggregate | total;
            int j = static_cast<int>(tile) - 1;
            while (true)
            {
                uint32_t s;
                do
                {
                    s = state[static_cast<size_t>(j) * kRadix + d];
                } while ((s & kFlagMask) == 0);
                exclusive += s & kValueMask;
                if (s & kFlagPrefix)
                    break;
                --j;
            }
            *myState = kFlagPrefix | (exclusive + total);
        }
        base[d] = gbase + exclusive;
    }
    __syncthreads();

    // scatter
#pragma unroll
    for (int k = 0; k < kSortItems; ++k)
    {
        const uint32_t idx = warpBase + k * 32 + lane;
        if (idx < n)
        {
            const uint32_t d = (key[k] >> shift) & 0xffu;
            const uint32_t pos = base[d] + wcount[warp][d] + rank[k];
            keysOut[pos] = key[k];
            valsOut[pos] = valsIn[idx];
        }
    }
}

size_t sort_workspace_bytes(uint32_t n, int passes)
{
    const uint32_t tiles = (n + kSortTile - 1) / kSortTile;
    // [ghist 4*256][tickets 4][state passes*tiles*256]
    return sizeof(uint32_t) * (4 * kRadix + 4 + static_cast<size_t>(passes) * (tiles ? tiles : 1) * kRadix);
}

// Sorts (keys, vals) by the low 8*passes bits of the key.  keys/vals are the input AND (for an even number
// of passes) the output; keysTmp/valsTmp are scratch of the same size.  Returns the number of launches; the
// sorted data ends in keys/vals when `passes` is even, otherwise in keysTmp/valsTmp.
int radix_sort_pairs(cudaStream_t stream, uint32_t* keys, uint32_t* vals, uint32_t* keysTmp, uint32_t* valsTmp,
                     uint32_t n, int passes, void* workspace)
{
    if (n == 0)
        return 0;
    const uint32_t tiles = (n + kSortTile - 1) / kSortTile;
    uint32_t* ghist = static_cast<uint32_t*>(workspace);
    uint32_t* tickets = ghist + 4 * kRadix;
    uint32_t* state = tickets + 4;
    cudaMemsetAsync(workspace, 0, sort_workspace_bytes(n, passes), stream);
    int blocks = static_cast<int>((n + 256 * 8 - 1) / (256 * 8));
    if (blocks > 148 * 4)
        blocks = 148 * 4;
    k_sort_histogram<<<blocks, 256, 0, stream>>>(keys, n, ghist, passes);
    uint32_t *ki = keys, *vi = vals, *ko = keysTmp, *vo = valsTmp;
    for (int p = 0; p < passes; ++p)
    {
        k_sort_pass<<<tiles, kSortThreads, 0, stream>>>(ki, vi, ko, vo, n, 8 * p, ghist + p * kRadix,
                                                       state + static_cast<size_t>(p) * tiles * kRadix, tickets + p);
        uint32_t* t = ki;
        ki = ko;
        ko = t;
        t = vi;
        vi = vo;
        vo = t;
    }
    return passes + 1;
}

} // namespace c2d_dev
