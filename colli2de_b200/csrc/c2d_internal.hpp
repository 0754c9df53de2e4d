// c2d_internal.hpp — declarations shared by the translation units of libc2d_gpu.so.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace c2d_dev
{

// ---- per-shape meta word -------------------------------------------------------------------------
// bits 0-1 ShapeType, 2-3 BodyType, 4 active (ShapeInstance::mIsActive), 5 alive (has a tree proxy),
// bits 8-11 polygon vertex count.
constexpr uint32_t META_TYPE_MASK = 0x3u;
constexpr uint32_t META_BODY_SHIFT = 2;
constexpr uint32_t META_ACTIVE = 1u << 4;
constexpr uint32_t META_ALIVE = 1u << 5;
constexpr uint32_t META_GHOST = 1u << 6; // halo copy of a shape owned by another rank (multi-GPU slabs)
constexpr uint32_t META_TIGHT_EXACT = 1u << 7; // tight box == computeAABB(shape, current transform): set by add / Transform
                                               // moves, cleared by translations (which shift the box) and restores
constexpr uint32_t META_COUNT_SHIFT = 8;
constexpr uint32_t META_FAT_DIRTY = 1u << 12; // the leaf (fat) box changed since the shape's tree last saw it (k_refit_dirty)

__host__ __device__ inline uint32_t meta_type(uint32_t m) { return m & META_TYPE_MASK; }
__host__ __device__ inline uint32_t meta_body(uint32_t m) { return (m >> META_BODY_SHIFT) & 0x3u; }
__host__ __device__ inline uint32_t meta_count(uint32_t m) { return (m >> META_COUNT_SHIFT) & 0xfu; }

// ---- mutation log records (pinned host -> device) ---------------------------------------------------
struct FlagRec // 8 B: remove / activate / deactivate
{
    uint32_t shape;
    uint32_t op; // 0 remove, 1 activate, 2 deactivate, 3 mark ghost, 4 clear ghost
};
struct TransRec // 12 B: moveEntity(id, Translation) for one shape (the bulk of a frame's upload: no padding)
{
    uint32_t shape;
    float dx, dy;
};
struct MoveRec // 32 B: moveEntity(id, Transform) / teleportEntity(id, Transform) for one shape
{
    uint32_t shape;
    float tx, ty, angle, sin, cos, csin, ccos;
};
struct AddRec // 224 B: addShape
{
    uint32_t shape, entity, meta, _pad;
    uint64_t category, mask;
    float xf[8];   // tx, ty, angle, sin, cos, csin, ccos, -
    float prev[8]; // previous transform (bullets)
    float geom[32];
};
static_assert(sizeof(FlagRec) == 8 && sizeof(TransRec) == 12 && sizeof(MoveRec) == 32 && sizeof(AddRec) == 224,
              "log record layout");

// ---- per-shape state in HBM (SoA) ---------------------------------------------------------------------
struct ShapeArrays
{
    float4* xf;    // tx, ty, sin, cos          (entity transform, as stored by the reference)
    float4* xa;    // angle, csin, ccos, 0
    float4* pxf;   // previous transform (bullets): tx, ty, sin, cos
    float4* pxa;   // previous: angle, csin, ccos, 0
    float4* tight; // ShapeInstance::mBoundingBox   (minx, miny, maxx, maxy)
    float4* fat;   // leaf box of the reference's DynamicBVH
    uint32_t* meta;
    uint32_t* entity;
    uint64_t* category;
    uint64_t* mask;
    float* geom; // 32 floats per shape
};

// ---- LBVH node: both children's boxes live in the parent so one 64 B read serves both tests -------------
// child link of a node: >= 0 internal node index; < 0 (bit 31) a leaf that carries its ShapeId in the low 30 bits, so
// emitting a candidate needs no dependent load of the sorted id array; bit 30 is a spare leaf flag (unused: the
// traversal kernels can skip leaves by flag, see rejectLeaf)
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kLeafGhost = 0x40000000u; // spare leaf flag
constexpr uint32_t kLeafIdMask = 0x3fffffffu;
__host__ __device__ inline uint32_t leaf_shape(int link) { return static_cast<uint32_t>(link) & kLeafIdMask; }

struct __align__(16) BvhNode
{
    float4 boxL;
    float4 boxR;
    int left;  // >= 0: internal node index; < 0: leaf (kLeafBit | ghost bit | ShapeId)
    int right;
    int lastL; // last sorted leaf position covered by the left / right subtree
    int lastR;
    uint64_t catL; // OR of the category bits below (BVHNode::mCategoryBits, DynamicBVH.hpp:440,478)
    uint64_t catR;
};
static_assert(sizeof(BvhNode) == 64, "node is one 64-byte record");

// frame scratch: [0..3] bounds tree 0 (minx, miny, maxx, maxy as ordered uints), [4..7] tree 1, [8..11] tree 2,
// then counters.
struct FrameScratch
{
    uint32_t bounds[3][4];
    uint32_t candCount;
    uint32_t outCount;
    uint32_t clipOverflow;
    uint32_t stackOverflow; // a traversal stack ran out (cannot happen with the depth bound of the Karras tree; never silent)
    uint32_t binCount[16];
    uint32_t binOffset[16];
    uint32_t binCursor[16];
    uint32_t hitCount[16]; // narrowphase hits per bin (phase 1)
    uint32_t hitBase[17];  // exclusive scan of hitCount; [16] = total
    uint32_t ticket;       // blocks of k_small_narrow that have finished (the last one publishes and resets the counters)
    uint32_t _pad2[2];
};

struct PairOut // 84 bytes, = Registry<uint32_t>::EntityCollision
{
    uint32_t entityA, entityB, shapeA, shapeB;
    float manifold[17];
};

// radix sort (c2d_sort.cu)
size_t sort_workspace_bytes(uint32_t n, int passes);
int radix_sort_pairs(cudaStream_t stream, uint32_t* keys, uint32_t* vals, uint32_t* keysTmp, uint32_t* valsTmp,
                     uint32_t n, int passes, void* workspace);

} // namespace c2d_dev
