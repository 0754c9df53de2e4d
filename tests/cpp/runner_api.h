/*
 * runner_api.h — one flat C interface ("rn_*") that drives a c2d::Registry
 * through its PUBLIC API only.  TEST INFRASTRUCTURE, not product code.
 *
 * The same runner.cpp is compiled twice:
 *   - against /root/reference/include + the reference's own four .cpp files
 *     -> oracle/_ref/libc2d_ref_runner.so      (the real reference, CPU)
 *   - against this repo's include/colli2de + libc2d_gpu.so
 *     -> colli2de_b200/libc2d_b200_runner.so   (the B200 drop-in)
 * which is itself the proof that the public API is source compatible.
 * oracle/c2d_oracle.c exports the same symbols natively (plain C restatement).
 *
 * Layouts (all little-endian, packed as numpy structured arrays in tests):
 *   xf3   = tx, ty, angleRadians                     (Transform(Vec2, float) ctor)
 *   xf5   = tx, ty, angleRadians, sin, cos           (raw Transform fields)
 *   geom16= circle  : cx, cy, r
 *           capsule : c1x, c1y, c2x, c2y, r
 *           segment : sx, sy, ex, ey
 *           polygon : v0x, v0y, ... v7x, v7y   (count in vcount[]; normals are
 *                     derived by Polygon::computeNormals, Shapes.hpp:142-154)
 *   pair record (84 B) = EntityCollision<uint32_t>  (Registry.hpp:44-51):
 *           entityA, entityB, shapeA, shapeB, Manifold{normal, points[2], pointCount}
 *   hit record (32 B)  = RaycastHit<pair<uint32_t, ShapeId>> (Ray.hpp:22-30):
 *           entity, shape, entry.xy, exit.xy, entryTime, exitTime
 *   ray4  = start.xy, end.xy (finite)  |  start.xy, direction.xy (infinite)
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rn_ctx rn_ctx;

const char* rn_backend_name(void);
rn_ctx* rn_create(int method /*0 None, 1 Grid*/, uint32_t cell_size);
void rn_destroy(rn_ctx*);
const char* rn_last_error(rn_ctx*);

void rn_create_entities(rn_ctx*, uint32_t n, const uint32_t* ids, const uint8_t* body, const float* xf3);
void rn_create_entities_raw(rn_ctx*, uint32_t n, const uint32_t* ids, const uint8_t* body, const float* xf5);
void rn_add_shapes(rn_ctx*, uint32_t n, const uint32_t* entity, const uint8_t* type, const float* geom16,
                   const uint8_t* vcount, const uint64_t* cat, const uint64_t* mask, uint32_t* out_ids);
void rn_remove_shapes(rn_ctx*, uint32_t n, const uint32_t* shape_ids);
void rn_remove_entities(rn_ctx*, uint32_t n, const uint32_t* ids);
void rn_set_active(rn_ctx*, uint32_t n, const uint32_t* shape_ids, const uint8_t* active);
void rn_move_translate(rn_ctx*, uint32_t n, const uint32_t* ids, const float* dxy);
void rn_move_transform(rn_ctx*, uint32_t n, const uint32_t* ids, const float* dxf3);
/* B200: Registry::moveEntities (teleport = 0) / teleportEntities(positions) (teleport = 1); reference: the loop of single
 * calls.  Bullets' previous transforms are not tracked by the runner here.  Not in the C oracle. */
void rn_move_entities_bulk(rn_ctx*, uint32_t n, const uint32_t* ids, const float* dxy, int teleport, double* seconds);
void rn_teleport_transform(rn_ctx*, uint32_t n, const uint32_t* ids, const float* xf3);
void rn_teleport_translation(rn_ctx*, uint32_t n, const uint32_t* ids, const float* xy);
void rn_get_transforms(rn_ctx*, uint32_t n, const uint32_t* ids, float* out_xf5);
/* previous transforms of bullets, tracked by the runner from getEntityTransform() before each move */
void rn_get_prev_transforms(rn_ctx*, uint32_t n, const uint32_t* ids, float* out_xf5);
uint64_t rn_size(rn_ctx*);
void rn_clear(rn_ctx*);

/* getCollidingPairs(): runs the query, keeps the result; returns the pair count.
 * seconds (may be NULL) = steady_clock around the getCollidingPairs() call only. */
uint64_t rn_get_colliding_pairs(rn_ctx*, double* seconds);
void rn_copy_pairs(rn_ctx*, void* out84);

/* getCollidingPairs() whose records stay on the device (B200: Registry::countCollidingPairsOnDevice; the
 * reference simply runs getCollidingPairs()).  Used for the device-resident timing of bench.py. */
uint64_t rn_count_colliding_pairs_device(rn_ctx*, double* seconds);
/* c2d_gpu_stats of the last query as 16 doubles (B200 only, zeros elsewhere):
 * shapes_alive, candidates, pairs, h2d_bytes, d2h_bytes, kernel_launches, retries,
 * ms_refit, ms_build, ms_broad, ms_narrow, ms_device, ms_total, 0, 0, 0 */
void rn_query_stats(rn_ctx*, double* out16);
/* Move batches: stage n (entity, translation) moves, then apply them any number of times.  B200: the batch
 * lives in HBM (Registry::stageMoves / applyStagedMoves); reference: a loop of moveEntity(id, Translation). */
int rn_stage_translations(rn_ctx*, uint32_t n, const uint32_t* ids, const float* dxy);
void rn_apply_staged(rn_ctx*, int handle);
void rn_release_staged(rn_ctx*, int handle);
/* `frames` x (rn_apply_staged(handles[f]); rn_count_colliding_pairs_device) in native code; sums16 = the frames' rn_query_stats
 * summed; returns the summed pair counts (not in the C oracle) */
uint64_t rn_run_staged_frames(rn_ctx*, uint32_t frames, const int* handles, double* sums16, double* seconds);

/* getCollisions(id): keeps the result like rn_get_colliding_pairs; returns the count (copy with rn_copy_pairs). */
uint64_t rn_get_collisions(rn_ctx*, uint32_t id);
/* getCollisions(ids[i]) looped in native code; returns the summed result sizes, *seconds = time of the loop */
uint64_t rn_get_collisions_many(rn_ctx*, uint32_t n, const uint32_t* ids, double* seconds);
/* areColliding(idsA[i], idsB[i]) for i < n */
void rn_are_colliding(rn_ctx*, uint32_t n, const uint32_t* idsA, const uint32_t* idsB, uint8_t* out);
/* firstHitRayCast() looped over n rays: out_found[i] = has_value, out_hits32[i] = the hit (zeroed when none) */
void rn_first_hit_raycast(rn_ctx*, uint32_t n, const float* ray4, const uint8_t* infinite, const uint64_t* masks,
                          uint8_t* out_found, void* out_hits32);

/* getCollidingPairs() through the zero-copy extension (B200: Registry::getCollidingPairsView, a view of the pinned
 * result buffer; reference: getCollidingPairs() into a kept vector).  *records stays valid until the next query. */
uint64_t rn_get_colliding_pairs_view(rn_ctx*, double* seconds, const void** records);

/* Multi-GPU halo copies (extension; B200 and the C oracle): see c2d_gpu_shape_set_ghost in include/c2d_gpu.h. */
/* Registry::serialize into a context-owned blob (returns its size; rn_copy_blob copies it out), Registry::deserialize
 * into a NEW context, Registry::operator== (1 equal, 0 different, -1 error). */
uint64_t rn_serialize(rn_ctx*);
void rn_copy_blob(rn_ctx*, void* out);
rn_ctx* rn_deserialize(int method, const void* bytes, uint64_t size);
void rn_seed_entities(rn_ctx*, uint32_t n, const uint32_t* ids, const uint8_t* body, const float* prev_xf5);
int rn_equals(rn_ctx* a, rn_ctx* b);
void rn_set_ghost(rn_ctx*, uint32_t n, const uint32_t* shape_ids, const uint8_t* ghost);

/* Multi-GPU group of Registries (extension, B200 only; not in the C oracle): Registry::createGroupId / joinGroup /
 * setGroupResults (0 local, 1 all, 2 root) + setGroupMoveExchange / leaveGroup. */
void rn_group_id(rn_ctx*, void* out128);
void rn_group_join(rn_ctx*, uint32_t rank, uint32_t world, const void* id128);
void rn_group_options(rn_ctx*, int results, int exchange);
void rn_group_leave(rn_ctx*);
void rn_group_ray_sharding(rn_ctx*, int on); /* Registry::setGroupRaySharding */

/* Secondary oracles.  Reference: read from the Registry's private members
 * (mShapes[i].mBoundingBox; the leaf box mTrees[body].getNode(mShapeTreeHandles[i]).mBoundingBox; the five
 * mTrees[..].findAllCollisions groups of Registry.hpp:487-545) - PartitioningMethod::None only.  B200: c2d_gpu_read_boxes /
 * c2d_gpu_read_shapes / c2d_gpu_read_candidates after a device-resident query.  Same-tree candidates are (min, max). */
uint32_t rn_shape_slots(rn_ctx*);
void rn_read_boxes(rn_ctx*, uint32_t first, uint32_t count, float* out_tight4, float* out_fat4, uint8_t* out_alive);
uint64_t rn_candidates(rn_ctx*); /* runs the broadphase, keeps the pairs; copy with rn_copy_candidates */
void rn_copy_candidates(rn_ctx*, uint32_t* out_pairs2);

/* B200 only: the c2d_gpu_ctx* behind the Registry (Registry::nativeContext); NULL elsewhere. */
void* rn_native_context(rn_ctx*);

/* rayCast() looped over n rays; offsets has n+1 entries; hits are concatenated in std::set order. */
uint64_t rn_raycast(rn_ctx*, uint32_t n, const float* ray4, const uint8_t* infinite, const uint64_t* masks,
                    uint64_t* offsets, double* seconds);
void rn_copy_hits(rn_ctx*, void* out32);

/* free functions of the narrowphase (Collision.hpp / RaycastShapes.hpp / ShapesUtils.hpp), batched over n */
void rn_collide(uint32_t n, const uint8_t* typeA, const float* geomA, const uint8_t* cntA, const float* xfA5,
                const uint8_t* typeB, const float* geomB, const uint8_t* cntB, const float* xfB5,
                void* out_manifold68, uint8_t* out_are_colliding);
/* mode 1: A moves xfA5 -> xfA5_end, B fixed at xfB5.  mode 2: both move. out_hit=0 => nullopt */
void rn_sweep(uint32_t n, int mode, const uint8_t* typeA, const float* geomA, const uint8_t* cntA,
              const float* xfA5, const float* xfA5_end, const uint8_t* typeB, const float* geomB,
              const uint8_t* cntB, const float* xfB5, const float* xfB5_end, float* out_fraction,
              void* out_manifold68, uint8_t* out_hit);
void rn_raycast_shape(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                      const float* ray4, const uint8_t* infinite, float* out_times2, uint8_t* out_hit);
void rn_sweep_ray(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                  const float* xf5_end, const float* ray4, const uint8_t* infinite, float* out_times2,
                  uint8_t* out_hit);
void rn_compute_aabb(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                     float* out_aabb4);
/* polygon normals as stored by Polygon::computeNormals: out = n x 16 floats */
void rn_polygon_normals(uint32_t n, const float* geom, const uint8_t* cnt, float* out_normals16);

#ifdef __cplusplus
}
#endif
