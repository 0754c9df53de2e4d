/*
 * c2d_gpu.h — the C ABI of libc2d_gpu.so: the B200 (sm_100a) implementation of
 * Colli2De's per-frame collision query path.
 *
 * The reference has no FFI/plugin interface: its boundary is the C++ class
 * c2d::Registry<EntityId, Method> (reference include/colli2de/Registry.hpp:40-106).
 * In this repo that class (include/colli2de/Registry.hpp) is a host-side mirror
 * that owns ids and transform arithmetic and forwards everything else through
 * the entry points below.  Every entry point cites the reference code it
 * replaces.  All functions return 0 on success and a non-zero code on failure
 * (c2d_gpu_last_error gives the message); no exceptions cross this boundary.
 * Plain pointers and sizes only.  One context per Registry; calls on one
 * context must be serialised by the caller (the reference Registry is not
 * thread safe for mutation either, README.md:142).
 *
 * Shapes are addressed by the Registry's ShapeId (dense, recycled LIFO by the
 * host, Registry.hpp:283-289).  The device keeps, per shape: type, body type,
 * active/alive bits, entity tag, category bits, the entity transform
 * (tx, ty, angle, sin, cos) and - for bullets - the previous transform, the
 * tight AABB (ShapeInstance::mBoundingBox, Registry.hpp:117) and the fat leaf
 * AABB (BVHNode::mBoundingBox of the leaf, DynamicBVH.hpp:31), plus geometry.
 *
 * Mutations are LOGGED on the host in pinned memory and replayed on the device
 * by the refit kernels at the next query (or c2d_gpu_flush), in an order that
 * preserves per-shape program order.
 */
#ifndef C2D_GPU_H
#define C2D_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct c2d_gpu_ctx c2d_gpu_ctx;

/* Shape / body codes: c2d::ShapeType (Shapes.hpp:15-21) and c2d::BodyType (Registry.hpp:27-32). */
enum { C2D_CIRCLE = 0, C2D_CAPSULE = 1, C2D_SEGMENT = 2, C2D_POLYGON = 3 };
enum { C2D_STATIC = 0, C2D_DYNAMIC = 1, C2D_BULLET = 2 };
/* Broadphase: LBVH (PartitioningMethod::None, the parity target) or uniform grid (PartitioningMethod::Grid). */
enum { C2D_BROADPHASE_LBVH = 0, C2D_BROADPHASE_GRID = 1 };

/* Transform as uploaded: the raw fields of c2d::Transform (Transform.hpp:122-126) plus the libm
 * sinf/cosf of `angle` (`csin`, `ccos`) that Transform(Vec2, float) would produce; sweeps re-derive
 * sin/cos from the angle (Collision.hpp:89-91) while collide() uses the stored, possibly drifted,
 * sin/cos.  32 bytes. */
typedef struct c2d_xf
{
    float tx, ty, angle, sin, cos, csin, ccos, _pad;
} c2d_xf;

/* Geometry slot, 32 floats:
 *   circle  : cx, cy, r                      (Shapes.hpp:39-43)
 *   capsule : c1x, c1y, c2x, c2y, r          (Shapes.hpp:64-69)
 *   segment : sx, sy, ex, ey                 (Shapes.hpp:93-96)
 *   polygon : v[0..7] (x,y) in [0,16), normals[0..7] (x,y) in [16,32)   (Shapes.hpp:118-122) */
#define C2D_GEOM_FLOATS 32

/* Output record of getCollidingPairs: byte-identical to
 * Registry<uint32_t>::EntityCollision (Registry.hpp:44-51) = 84 bytes. */
typedef struct c2d_manifold_point
{
    float point[2], anchorA[2], anchorB[2], separation;
} c2d_manifold_point;
typedef struct c2d_manifold
{
    float normal[2];
    c2d_manifold_point points[2];
    uint8_t pointCount;
    uint8_t _pad[3];
} c2d_manifold;
typedef struct c2d_pair_record
{
    uint32_t entityA, entityB; /* the entity tags given to c2d_gpu_shape_add */
    uint32_t shapeA, shapeB;
    c2d_manifold manifold;
} c2d_pair_record;

/* Output record of rayCast: RaycastHit<pair<uint32_t, ShapeId>> (Ray.hpp:22-30) = 32 bytes. */
typedef struct c2d_hit_record
{
    uint32_t entity, shape;
    float entry[2], exit[2], entryTime, exitTime;
} c2d_hit_record;

/* Ray as uploaded: start, then end (finite, Ray.hpp:10-14) or direction (infinite, Ray.hpp:16-20). */
typedef struct c2d_ray
{
    float sx, sy, ex, ey;
    uint64_t mask; /* maskBits of rayCast (Registry.hpp:1052) */
    uint32_t infinite;
    uint32_t _pad;
} c2d_ray;

/* Per-query statistics.  Times are CUDA-event milliseconds on the context's stream. */
typedef struct c2d_gpu_stats
{
    uint64_t shapes_alive;    /* proxies in the three trees */
    uint64_t candidates;      /* broadphase pairs (fat AABB overlap & category match) */
    uint64_t pairs;           /* colliding pairs returned */
    uint64_t h2d_bytes;       /* mutation log bytes uploaded by this call */
    uint64_t d2h_bytes;       /* result bytes downloaded by this call */
    uint32_t kernel_launches; /* kernels launched by this call */
    uint32_t retries;         /* pair-buffer overflow re-runs */
    float ms_refit;           /* mutation replay (K1) */
    float ms_build;           /* bounds + morton + sort + tree build + fit (K2-K4) */
    float ms_broad;           /* traversal (K5) */
    float ms_narrow;          /* binning + narrowphase/sweeps (K6/K7) */
    float ms_device;          /* first kernel -> last kernel */
    float ms_total;           /* host wall clock of the call incl. copies */
    float ms_host_enqueue;    /* host wall clock from the start of the call until the whole frame was enqueued */
    float ms_host_wait;       /* ... until the device had finished it (counters read, retries done) */
    float _reserved;
} c2d_gpu_stats;

/* ---- lifetime -------------------------------------------------------------------------------- */
/* Registry() / Registry(cellSize) (Registry.hpp:53-59).  `device` < 0 selects the current device. */
int c2d_gpu_create(int device, c2d_gpu_ctx** out_ctx);
void c2d_gpu_destroy(c2d_gpu_ctx* ctx);
/* Message of the last failure on `ctx` (or of the last stateless/creation failure when ctx is NULL). */
const char* c2d_gpu_last_error(c2d_gpu_ctx* ctx);
/* Registry::clear (Registry.hpp:1190-1203). */
int c2d_gpu_clear(c2d_gpu_ctx* ctx);
/* PartitioningMethod::None -> LBVH; PartitioningMethod::Grid(cellSize) -> uniform grid (BroadPhaseTree.hpp:96). */
int c2d_gpu_set_broadphase(c2d_gpu_ctx* ctx, int mode, uint32_t cell_size);
/* LBVH maintenance policy: the topology (Morton sort + Karras links) of a tree whose membership did not change is
 * rebuilt every `frames`-th dirty frame and only refit in between (leaf boxes exact, see c2d_gpu_set_incremental_refit).
 * 1 = rebuild every frame.  Default 8.  Results do not depend on it. */
int c2d_gpu_set_rebuild_period(c2d_gpu_ctx* ctx, uint32_t frames);
/* A counter bumped by every call that changes what a query would answer (add / remove / activate / ghost / translate /
 * move / apply_staged / clear / write_boxes).  Equal epochs = identical answers: lets a host layer memoise query results
 * (c2d::Registry caches the per-entity results of getCollisions() between mutations). */
int c2d_gpu_mutation_epoch(c2d_gpu_ctx* ctx, uint64_t* out_epoch);
/* Scenes with at most `shapes` live shapes skip the tree pipeline: one all-pairs box-test kernel over the member
 * lists and one narrowphase kernel (same candidate predicate, same results) - the ~30 dependent launches of the
 * tree pipeline cost more than the work at that size.  Default 8192; 0 disables the shortcut. */
int c2d_gpu_set_small_scene_limit(c2d_gpu_ctx* ctx, uint32_t shapes);
/* ---- multi-GPU --------------------------------------------------------------------
 * One context per GPU (one process per GPU under torchrun, or several contexts in one process).  Shape state is
 * REPLICATED: every rank of a group replays every mutation (or receives the other ranks' translations, below), so
 * ownership needs no communication - every rank derives the same Morton keys and the same slab boundaries from the same
 * boxes.  What is sharded is the per-frame work of Registry::getCollidingPairs (Registry.hpp:464-549): each rank links,
 * refits and traverses only ITS slab and returns its share of the pairs (the Morton sort that defines the slabs is
 * replicated, once per rebuild period); the union over the ranks is the
 * reference's pair set, each pair exactly once (owner rule).  (The grid broadphase splits its sorted cell entries instead.)
 *
 * c2d_gpu_set_shard: this context is `rank` of `world`.  Default 0 of 1.
 *   - the Dynamic tree (c2d_gpu_set_slab_mode, below) is cut into Morton slabs: all Dynamic shapes are ordered by the
 *     Morton key of their leaf box (the same stable sort on every rank) and rank r owns positions [n r / world, n (r+1) / world)
 *     - balanced by construction.  The order is renewed when shapes are added / removed and every `rebuild period` frames
 *     (c2d_gpu_set_rebuild_period); ownership is frozen in between.  A rank's tree holds its owned leaves (refit every
 *     frame); the HALO - shapes of higher ranks whose current leaf box may overlap an owned one - is selected every frame on
 *     the device and queries that tree.  Only owned leaves and halo shapes query; a pair of two slabs is emitted by the
 *     lower rank.
 *   - Bullet queries are split by index range; Bullet x Dynamic pairs are emitted by the rank that owns the Dynamic shape. */
int c2d_gpu_set_shard(c2d_gpu_ctx* ctx, uint32_t rank, uint32_t world);
/* Morton-slab build of the Dynamic tree when world > 1 and the tree has at least `min_shapes` leaves (smaller trees are
 * mirrored on every rank and only the query leaves are split).  0 = never.  Default 65536. */
int c2d_gpu_set_slab_mode(c2d_gpu_ctx* ctx, uint32_t min_shapes);
/* Batched rayCast in a group (SURVEY.md section 8(e) "Rays", cfg 5): rays are independent units, so with ray sharding on every
 * rank is handed the SAME batch (c2d_gpu_raycast / _begin / _first) over its mirror of the scene and casts the contiguous
 * share [n rank / world, n (rank+1) / world) of it; the other rays come back with empty hit lists on this rank.  The
 * per-ray lists of the ranks concatenate to the single-GPU result; no exchange, every rank downloads over its own PCIe
 * link.  Needs c2d_gpu_set_shard (or c2d_gpu_comm_init) only.  Default off. */
int c2d_gpu_set_ray_sharding(c2d_gpu_ctx* ctx, int on);
/* The NCCL communicator of the group (libnccl.so.2 is loaded with dlopen on first use; C2D_NCCL_LIB overrides the path).
 * c2d_gpu_comm_unique_id writes the 128-byte ncclUniqueId that rank 0 hands to the other ranks (by any out-of-band
 * channel: torch.distributed, MPI, a file); c2d_gpu_comm_init is collective over the group and implies
 * c2d_gpu_set_shard(rank, world). */
int c2d_gpu_comm_unique_id(void* out_id128);
int c2d_gpu_comm_init(c2d_gpu_ctx* ctx, uint32_t rank, uint32_t world, const void* id128);
int c2d_gpu_comm_destroy(c2d_gpu_ctx* ctx);
/* Move exchange (off by default = every rank logs every move itself).  On: each rank logs only the translations
 * (c2d_gpu_shape_translate / c2d_gpu_shapes_translate) of the entities IT drives; the next c2d_gpu_collide* exchanges the
 * 12-byte records of all ranks over NVLink, device to device, straight out of each rank's device log, and applies them
 * everywhere, so the host-to-device upload of a frame shrinks to 1 / world per rank.  The first exchange all-gathers the
 * counts, waits for them on the host and broadcasts exactly each rank's records; every later frame sends fixed-size
 * blocks (the previous frame's largest count + 25 %) and their headers in ONE grouped NCCL launch without a host round
 * trip - if a rank logged more than its block holds, no rank applies anything and the frame is redone with exact sizes.
 * Contract: an entity is translated by one rank and at most once per frame; Static shapes are not translated while the
 * exchange is on; every other mutation stays replicated and is applied BEFORE the frame's translations; while
 * translations are pending only c2d_gpu_collide* may flush them (it is the collective). */
int c2d_gpu_comm_set_move_exchange(c2d_gpu_ctx* ctx, int on);
/* Where the 84-byte records of a frame end up: C2D_RESULTS_LOCAL every rank keeps its share; C2D_RESULTS_ALL every rank
 * receives all records (ncclAllGather of the counts + one ncclBroadcast per rank, device to device, in rank order);
 * C2D_RESULTS_ROOT rank 0 receives all records (ncclSend / ncclRecv), the others keep their share. */
enum { C2D_RESULTS_LOCAL = 0, C2D_RESULTS_ALL = 1, C2D_RESULTS_ROOT = 2 };
int c2d_gpu_comm_set_result_mode(c2d_gpu_ctx* ctx, int mode);
/* Collectives of the last c2d_gpu_collide* on this rank: payload bytes received and device milliseconds (CUDA events
 * around the NCCL calls on the context's stream). */
typedef struct c2d_gpu_comm_stats
{
    uint64_t exchange_bytes;  /* translation records received from all ranks */
    uint64_t gather_bytes;    /* pair records received */
    uint64_t local_pairs;     /* pairs this rank computed */
    uint64_t slab_owned;      /* Dynamic leaves this rank owns */
    uint64_t slab_leaves;     /* ... plus the halo shapes that queried them this frame */
    float ms_exchange;
    float ms_gather;
    float ms_slab;            /* the slab step of the frame: (rebalance: keys + sort + links) + refit + halo selection */
} c2d_gpu_comm_stats;
int c2d_gpu_comm_get_stats(c2d_gpu_ctx* ctx, c2d_gpu_comm_stats* out);

/* Order of the records a query returns.  0 (default): grouped by narrowphase bin, order inside a bin unspecified and not
 * reproducible from run to run (warp-aggregated appends).  1: the reference's order (Registry.hpp:487-545) - its five groups
 * one after the other (Dynamic x Static, Dynamic x Dynamic, Bullet x Static, Bullet x Dynamic, Bullet x Bullet), each
 * sorted by (shapeA, shapeB); byte-identical from run to run.  Same-tree groups use this library's canonical
 * orientation shapeA < shapeB (the reference's orientation there is an accident of its tree layout).  Costs two radix sorts
 * of an index permutation and one gather of the 84-byte records per frame.  In a multi-GPU group every rank orders its
 * own share. */
int c2d_gpu_set_result_order(c2d_gpu_ctx* ctx, int reference_order);
/* Refit of the frames that keep a tree's topology (c2d_gpu_set_rebuild_period): 1 (default) only the leaves whose leaf box
 * changed act - the exact box goes into the leaf's slot and the ancestors are grown to contain it; internal boxes are
 * conservative until the next rebuild, leaf boxes (hence the candidate set) stay exact.  0: full bottom-up fit every frame. */
int c2d_gpu_set_incremental_refit(c2d_gpu_ctx* ctx, int on);
/* Independent kernels of a frame run side by side on a second stream (the Dynamic x Static traversal next to the Dynamic
 * tree's refit and self traversal, the two filter kernels, the two manifold kernels).  Results are identical; 0 puts
 * everything back on one stream (for per-kernel timing that should not see the neighbours).  Default on. */
int c2d_gpu_set_overlap(c2d_gpu_ctx* ctx, int on);
/* Per-kernel CUDA-event timing of the queries (off by default).  c2d_gpu_kernel_profile writes one line per
 * kernel name, "name<TAB>launches<TAB>total_ms", accumulated since the last reset. */
int c2d_gpu_set_profiling(c2d_gpu_ctx* ctx, int on);
/* With profiling on, time only the launches of this kernel name (NULL or "" = all).  One event pair per kernel costs
 * ~3 us; timing all ~25 launches of a frame costs ~0.12 ms per frame (measured), timing one is free. */
int c2d_gpu_set_profile_filter(c2d_gpu_ctx* ctx, const char* kernel_name);
int c2d_gpu_kernel_profile(c2d_gpu_ctx* ctx, char* buffer, uint64_t capacity, int reset);

/* ---- mutation log ----------------------------------------------------------------------------- */
/* Registry::addShape (Registry.hpp:272-330): tight AABB = computeAABB(shape, xf) (ShapesUtils.hpp:13-60),
 * leaf AABB = tight.fattened(0) (DynamicBVH.hpp:352-365).  `prev` is the bullet's previous transform
 * (Registry.hpp:167); pass NULL for `prev` to use `xf`. */
int c2d_gpu_shape_add(c2d_gpu_ctx* ctx, uint32_t shape, uint32_t entity_tag, uint32_t type, uint32_t body,
                      uint32_t vertex_count, uint64_t category, uint64_t mask, const float* geom32,
                      const c2d_xf* xf, const c2d_xf* prev);
/* Registry::removeShape / removeEntity (Registry.hpp:332-380): the proxy leaves its tree. */
int c2d_gpu_shape_remove(c2d_gpu_ctx* ctx, uint32_t shape);
/* Registry::setShapeActive (Registry.hpp:353-358). */
int c2d_gpu_shape_set_active(c2d_gpu_ctx* ctx, uint32_t shape, int active);
/* Multi-GPU slabs (no reference counterpart): marks a shape as a halo ("ghost") copy of a shape another rank owns.
 * Ghosts take part in the broadphase and narrowphase like any shape, but a pair whose OWNER shape (shape A of a
 * cross-tree pair; the shape of the entity with the smaller tag of a same-tree pair) is a ghost is left to the rank
 * that owns it, so the union of all ranks' results has every pair exactly once. */
int c2d_gpu_shape_set_ghost(c2d_gpu_ctx* ctx, uint32_t shape, int ghost);
/* Registry::moveEntity(id, Translation) per shape (Registry.hpp:434-453 -> translateDynamicShape /
 * translateBulletShape 214-231 -> DynamicBVH::moveProxy 386-417): tight box shifted, bullets sweep
 * combine(old, new), fat-leaf rule, previous transform captured for bullets, translation += d. */
int c2d_gpu_shape_translate(c2d_gpu_ctx* ctx, uint32_t shape, float dx, float dy);
int c2d_gpu_shapes_translate(c2d_gpu_ctx* ctx, uint32_t n, const uint32_t* shapes, const float* dxy);
/* Registry::moveEntity(id, Transform) / teleportEntity(id, Transform) per shape (Registry.hpp:382-432 ->
 * moveDynamicShape / moveBulletShape 191-212): tight = computeAABB at the new transform `xf`
 * (the host has already applied Transform += delta), same leaf rule. */
int c2d_gpu_shape_move(c2d_gpu_ctx* ctx, uint32_t shape, const c2d_xf* xf);
/* Device-resident move batches (extension; the reference has only the per-entity call): the per-shape
 * translations are uploaded ONCE, then each c2d_gpu_apply_staged logs "translate all of them" in O(1) host
 * work and no PCIe traffic - the moveEntity(Translation) semantics above, n shapes at a time.  A batch must
 * name each shape at most once and only live shapes (also at apply time). */
int c2d_gpu_stage_translations(c2d_gpu_ctx* ctx, uint32_t n, const uint32_t* shapes, const float* dxy,
                               uint32_t* out_handle);
int c2d_gpu_apply_staged(c2d_gpu_ctx* ctx, uint32_t handle);
int c2d_gpu_release_staged(c2d_gpu_ctx* ctx, uint32_t handle);
/* Replays the log on the device now (otherwise done by the next query). */
int c2d_gpu_flush(c2d_gpu_ctx* ctx);

/* ---- queries ---------------------------------------------------------------------------------- */
/* Registry::getCollidingPairs (Registry.hpp:464-549): broadphase (DynamicBVH::findAllCollisions x5,
 * DynamicBVH.hpp:931-1031) + narrowphase (narrowPhaseCollision, Registry.hpp:608-710; collide / sweep).
 * *out_records points at context-owned pinned host memory valid until the next call on ctx.
 * Same-tree pairs are emitted with shapeA < shapeB; cross-tree pairs with A on the Dynamic/Bullet side.
 * The ORDER of the records is unspecified (they are grouped by narrowphase code path); the reference's order
 * (five groups, each sorted by (shapeA, shapeB), Registry.hpp:487-545) is obtained by sorting on
 * (group, shapeA, shapeB) — parity is defined on the sorted tuples.  `stats` may be NULL. */
int c2d_gpu_collide(c2d_gpu_ctx* ctx, const c2d_pair_record** out_records, uint64_t* out_count,
                    c2d_gpu_stats* stats);
/* Same pipeline, results left on the device (no D2H of records): used to time the device-resident path.
 * out_count receives the number of colliding pairs. */
int c2d_gpu_collide_device(c2d_gpu_ctx* ctx, uint64_t* out_count, c2d_gpu_stats* stats);
/* The copy-out of Registry::getCollidingPairs (Registry.hpp:464-549 returns a fresh std::vector, reserved from the
 * previous frame's count, 467 and 547) without the pinned-buffer detour of c2d_gpu_collide:
 *   c2d_gpu_collide_begin   replays the log and enqueues the whole pipeline, returns while the GPU works;
 *   c2d_gpu_prefault        (optional) helper threads touch the pages of the storage the caller reserved meanwhile;
 *   c2d_gpu_collide_end     waits (re-running on a pair-list overflow), joins the pre-fault, reports the count;
 *   c2d_gpu_collide_fetch   copies `count` (<= the frame's count) records into `dst`: chunked device-to-host copies
 *                           into pinned staging, moved on to `dst` by the helper threads as they land.
 * Helper threads: min(8, hardware threads / 2) including the caller (with torchrun's LOCAL_WORLD_SIZE = R > 1:
 * min(8, hardware threads / R)), C2D_GPU_COPY_THREADS overrides
 * (1 = none);
 * results below 1 MB are copied by the calling thread alone.  Records are the same as c2d_gpu_collide's. */
int c2d_gpu_collide_begin(c2d_gpu_ctx* ctx);
int c2d_gpu_prefault(c2d_gpu_ctx* ctx, void* dst, uint64_t bytes);
int c2d_gpu_collide_end(c2d_gpu_ctx* ctx, uint64_t* out_count, c2d_gpu_stats* stats);
int c2d_gpu_collide_fetch(c2d_gpu_ctx* ctx, c2d_pair_record* dst, uint64_t count, c2d_gpu_stats* stats);
/* Device address and count of the 84-byte records produced by the last c2d_gpu_collide / c2d_gpu_collide_device
 * (valid until the next query on ctx): lets a multi-GPU driver gather the per-slab results GPU-to-GPU (NCCL)
 * without a host round trip. */
int c2d_gpu_device_records(c2d_gpu_ctx* ctx, const void** out_device_ptr, uint64_t* out_count);
/* Registry::rayCast x n (Registry.hpp:1050-1182; DynamicBVH::piercingRaycast 775-806/853-884;
 * raycast(shape) RaycastShapes.cpp; sweep(shape, ray) Collision.hpp:323-344).  Hits of ray i are
 * out_hits[out_offsets[i] .. out_offsets[i+1]) ordered by (entryTime, exitTime) with exact-key
 * duplicates dropped like std::set does.  Buffers are context-owned pinned memory. */
int c2d_gpu_raycast(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, const c2d_hit_record** out_hits,
                    const uint64_t** out_offsets, c2d_gpu_stats* stats);

/* c2d_gpu_raycast / c2d_gpu_raycast_first with the copy-out of c2d_gpu_collide_fetch: _begin runs the batch and
 * reports the total number of hits and the per-ray offsets (context-owned, n + 1 entries); _fetch copies all `count`
 * (== that total) records into caller storage with the helper threads, straight from the device. */
int c2d_gpu_raycast_begin(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, int first_hit_only, uint64_t* out_total,
                          const uint64_t** out_offsets, c2d_gpu_stats* stats);
int c2d_gpu_raycast_fetch(c2d_gpu_ctx* ctx, c2d_hit_record* dst, uint64_t count, c2d_gpu_stats* stats);

/* Registry::firstHitRayCast x n (Registry.hpp:866-1048): at most one hit per ray (out_offsets[i+1] - out_offsets[i]
 * is 0 or 1) — the narrow hit with the smallest entry time found by the reference's search, which walks each tree's
 * leaf-box hits in (entry, exit) order and stops 5 candidates after its last improvement. */
int c2d_gpu_raycast_first(c2d_gpu_ctx* ctx, uint32_t n, const c2d_ray* rays, const c2d_hit_record** out_hits,
                          const uint64_t** out_offsets, c2d_gpu_stats* stats);
/* Registry::getCollisions (Registry.hpp:551-605) for the given shapes of one or more entities: every active listed
 * shape queries the trees with its tight AABB (bullets: swept) and ITS MASK BITS, candidates go through
 * narrowPhaseCollision.  Records are (queried shape, other shape); order unspecified. */
int c2d_gpu_query_shapes(c2d_gpu_ctx* ctx, uint32_t n, const uint32_t* shapes, const c2d_pair_record** out_records,
                         uint64_t* out_count);
/* Registry::areColliding (Registry.hpp:748-864) per shape pair: boolean narrowphase (bullets: boolean sweep),
 * no broadphase; inactive shapes never collide. */
int c2d_gpu_pairs_colliding(c2d_gpu_ctx* ctx, uint32_t n, const uint32_t* shapeA, const uint32_t* shapeB,
                            uint8_t* out_colliding);

/* ---- introspection (secondary oracles) --------------------------- */
/* Tight (ShapeInstance::mBoundingBox) and fat (leaf) AABBs as (minx, miny, maxx, maxy); flushes first. */
int c2d_gpu_read_boxes(c2d_gpu_ctx* ctx, uint32_t first, uint32_t count, float* out_tight4, float* out_fat4);
/* Per-shape state for snapshots (Registry::serialize, Registry.hpp:1205-1264 / ShapeInstance::serialize 1493-1570):
 * shape type (0 circle, 1 capsule, 2 segment, 3 polygon), polygon vertex count, flags (bit 0 alive, bit 1 active,
 * bit 2 ghost), category / mask bits and the 32-float geometry block; any output may be NULL.  Flushes first. */
int c2d_gpu_read_shapes(c2d_gpu_ctx* ctx, uint32_t first, uint32_t count, uint8_t* out_type, uint8_t* out_vertex_count,
                        uint8_t* out_flags, uint64_t* out_category, uint64_t* out_mask, float* out_geom32);
/* Per-shape entity transforms as the device holds them: xf4 = (tx, ty, sin, cos), xa4 = (angle, sin(angle), cos(angle), 0)
 * and the same for the previous transform of bullets (Registry.hpp:191-260 mBulletPreviousTransforms).  Used by a
 * Registry of a multi-GPU group to bring its host mirror up to date after other ranks' translations arrived through
 * the move exchange.  Any output may be NULL.  Flushes first. */
int c2d_gpu_read_transforms(c2d_gpu_ctx* ctx, uint32_t first, uint32_t count, float* out_xf4, float* out_xa4, float* out_pxf4,
                            float* out_pxa4);
/* Restores the tight (ShapeInstance::mBoundingBox) and leaf (BVHNode::mBoundingBox) boxes of a shape range, e.g. after
 * Registry::deserialize (Registry.hpp:1266-1344) re-added the shapes of a snapshot: both boxes are HISTORY (the tight
 * box is shifted by translations, the leaf box follows the fat-leaf rule), so they cannot be recomputed.  Either
 * pointer may be NULL.  Flushes first; the next query rebuilds the trees. */
int c2d_gpu_write_boxes(c2d_gpu_ctx* ctx, uint32_t first, uint32_t count, const float* tight4, const float* fat4);
/* 1 + the highest ShapeId ever added (the size of the reference's mShapes table, Registry.hpp:159). */
int c2d_gpu_shape_slots(c2d_gpu_ctx* ctx, uint32_t* out_slots);
/* Broadphase candidate pairs of the last c2d_gpu_collide call (shapeA, shapeB), unsorted. */
int c2d_gpu_read_candidates(c2d_gpu_ctx* ctx, uint64_t capacity, uint32_t* out_pairs2, uint64_t* out_count);

/* ---- stateless batches of the narrowphase free functions ----------------------------------------- */
/* geom: n x 32 floats (layout above); xf5: n x (tx, ty, angle, sin, cos); out arrays may be NULL. */
/* c2d::collide / c2d::areColliding for all 16 type pairs (Collision.cpp:15-748). */
int c2d_gpu_collide_batch(uint32_t n, const uint8_t* typeA, const float* geomA, const uint8_t* cntA,
                          const float* xfA5, const uint8_t* typeB, const float* geomB, const uint8_t* cntB,
                          const float* xfB5, void* out_manifold68, uint8_t* out_are_colliding);
/* c2d::sweep one-moving (mode 1, Collision.hpp:246-265) / two-moving (mode 2, Collision.hpp:280-307). */
int c2d_gpu_sweep_batch(uint32_t n, int mode, const uint8_t* typeA, const float* geomA, const uint8_t* cntA,
                        const float* xfA5, const float* xfA5_end, const uint8_t* typeB, const float* geomB,
                        const uint8_t* cntB, const float* xfB5, const float* xfB5_end, float* out_fraction,
                        void* out_manifold68, uint8_t* out_hit);
/* c2d::raycast(shape, T, ray) (RaycastShapes.cpp:11-320); with xf5_end != NULL: c2d::sweep(shape, T0, T1, ray)
 * (Collision.hpp:323-344).  ray4 = start.xy + end.xy | direction.xy. */
int c2d_gpu_raycast_shape_batch(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt,
                                const float* xf5, const float* xf5_end, const float* ray4,
                                const uint8_t* infinite, float* out_times2, uint8_t* out_hit);
/* c2d::computeAABB (ShapesUtils.hpp:13-60). */
int c2d_gpu_compute_aabb_batch(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt,
                               const float* xf5, float* out_aabb4);

#ifdef __cplusplus
}
#endif
#endif /* C2D_GPU_H */
