// colli2de-b200 — c2d::Registry<EntityId, Method>: the drop-in front end of the B200 collision query path.
//
// Same public signatures as the reference's include/colli2de/Registry.hpp:40-106, but the class is a
// HOST-SIDE MIRROR: it owns what must stay on the host —
//   * the EntityId -> dense entity index map and the entity transforms, with the reference's exact
//     transform arithmetic (Transform += delta incl. the Rotation::operator+= quirk, SURVEY.md D5;
//     bullet previous-transform capture, Registry.hpp:167,396-397,419-420,449-450),
//   * ShapeId allocation with LIFO recycling (Registry.hpp:283-289,345,372),
// and forwards everything else through the C ABI of libc2d_gpu.so (include/c2d_gpu.h): per-shape mutation
// records are logged and replayed by the sm_100a refit kernels, getCollidingPairs()/rayCast() run the
// LBVH broadphase + narrowphase kernels and return the reference's value types.
//
// There is NO CPU fallback: constructing a Registry without a CUDA device throws std::runtime_error.
//
// Differences a caller can observe (all documented in DESIGN.md):
//   * the Registry is movable but not copyable (it owns device state);
//   * getCollidingPairs() returns the reference's SET of collisions in an unspecified order (the reference
//     lists five body-type groups, each sorted by (shapeA, shapeB)); same-tree pairs always come out with
//     shapeA < shapeB (the reference's A/B for those depends on its tree topology);
//   * serialize / deserialize speak the reference's binary snapshot format (the trees in a written snapshot are
//     rebuilt, see internal/utils/Snapshot.hpp); operator== compares the logical state (entities, shapes, boxes), not
//     tree topology;
//   * extensions: moveEntities(), stageMoves() / applyStagedMoves(), rayCastBatch(), lastQueryStats().
#pragma once

#include <colli2de/Manifold.hpp>
#include <colli2de/Ray.hpp>
#include <colli2de/Shapes.hpp>
#include <colli2de/Transform.hpp>
#include <colli2de/internal/utils/Debug.hpp>
#include <colli2de/internal/utils/Snapshot.hpp>

#include <c2d_gpu.h>

#include <algorithm>
#include <array>
#include <bit>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <istream>
#include <optional>
#include <ostream>
#include <set>
#include <span>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace c2d
{

namespace detail
{
// Sets the size of a vector of trivially copyable records WITHOUT value-initialising the new elements, so that the
// library's copy threads can fill the storage in parallel (std::vector::resize would zero 35 MB on one thread first).
// Only libstdc++'s layout is known here; elsewhere the caller falls back to the single-threaded range constructor.
#if defined(__GLIBCXX__) && !defined(_GLIBCXX_DEBUG) && !defined(_GLIBCXX_SANITIZE_VECTOR)
inline constexpr bool kHasUninitializedResize = true;
template <typename T>
struct VectorSizeAccess : std::vector<T>
{
    static void set(std::vector<T>& v, std::size_t n)
    {
        // bit-copyable in practice: std::pair members make RaycastHit formally non-trivially-copyable (assignment)
        static_assert(std::is_trivially_copy_constructible_v<T> && std::is_trivially_destructible_v<T>);
        static_assert(sizeof(VectorSizeAccess) == sizeof(std::vector<T>));
        auto& self = static_cast<VectorSizeAccess&>(v);
        self._M_impl._M_finish = self._M_impl._M_start + n;
    }
};
template <typename T>
void resizeUninitialized(std::vector<T>& v, std::size_t n)
{
    if (v.capacity() < n)
        v.reserve(n);
    VectorSizeAccess<T>::set(v, n);
}
#else
inline constexpr bool kHasUninitializedResize = false;
template <typename T>
void resizeUninitialized(std::vector<T>& v, std::size_t n)
{
    v.resize(n);
}
#endif
} // namespace detail

using ShapeId = uint32_t;
using BitMaskType = uint64_t;

inline constexpr uint8_t bodyTypesCount = 3;
enum class BodyType : uint8_t
{
    Static = 0,
    Dynamic,
    Bullet,
};

enum class PartitioningMethod
{
    None = 0, // LBVH broadphase (the parity target)
    Grid,     // uniform-grid broadphase
};

namespace detail
{

// EntityId -> dense index.  Open addressing over 32-bit slots; the keys themselves live in the entity array,
// so a lookup touches one slot run plus one entity record.
template <typename EntityId, typename KeyOf>
class IndexMap
{
  public:
    static constexpr uint32_t npos = 0xffffffffu;

    uint32_t find(const EntityId& id, const KeyOf& keyOf) const
    {
        if (mSlots.empty())
            return npos;
        for (size_t pos = home(id);; pos = (pos + 1) & mMask)
        {
            const uint32_t slot = mSlots[pos];
            if (slot == kEmpty)
                return npos;
            if (slot != kTomb && keyOf(slot - 2) == id)
                return slot - 2;
        }
    }

    // bulk lookups (Registry::moveEntities): the slot of a key a few iterations ahead is fetched early, then peeked to
    // fetch the entity record it names - a lookup in a table of a million entities is two dependent cache misses otherwise
    void prefetch(const EntityId& id) const
    {
        if (!mSlots.empty())
            __builtin_prefetch(&mSlots[home(id)]);
    }
    uint32_t peek(const EntityId& id) const // the home slot's entry (npos if empty / tombstone): a hint, not a lookup
    {
        if (mSlots.empty())
            return npos;
        const uint32_t slot = mSlots[home(id)];
        return slot >= 2 ? slot - 2 : npos;
    }

    // precondition: !needsRehash()
    void insert(const EntityId& id, uint32_t index)
    {
        size_t pos = home(id);
        while (mSlots[pos] != kEmpty && mSlots[pos] != kTomb)
            pos = (pos + 1) & mMask;
        if (mSlots[pos] == kEmpty)
            ++mUsed;
        mSlots[pos] = index + 2;
    }

    bool needsRehash() const { return mSlots.empty() || (mUsed + 1) * 10 >= mSlots.size() * 7; }

    // rebuilds the table (dropping tombstones) with room for twice the live entries
    void rehash(const KeyOf& keyOf, const std::vector<uint32_t>& liveIndices)
    {
        size_t cap = 16;
        while (cap * 7 < (liveIndices.size() + 1) * 2 * 10)
            cap *= 2;
        mSlots.assign(cap, kEmpty);
        mMask = cap - 1;
        mShift = 64 - std::countr_zero(cap);
        mUsed = 0;
        for (uint32_t idx : liveIndices)
            insert(keyOf(idx), idx);
    }

    void erase(const EntityId& id, const KeyOf& keyOf)
    {
        for (size_t pos = home(id);; pos = (pos + 1) & mMask)
        {
            const uint32_t slot = mSlots[pos];
            if (slot == kEmpty)
                return;
            if (slot != kTomb && keyOf(slot - 2) == id)
            {
                mSlots[pos] = kTomb;
                return;
            }
        }
    }

    void clear()
    {
        mSlots.clear();
        mMask = mUsed = 0;
    }

  private:
    static constexpr uint32_t kEmpty = 0, kTomb = 1;
    size_t home(const EntityId& id) const
    {
        return static_cast<size_t>((static_cast<uint64_t>(std::hash<EntityId>{}(id)) * 0x9E3779B97F4A7C15ull) >> mShift);
    }
    std::vector<uint32_t> mSlots;
    size_t mMask = 0, mUsed = 0; // mUsed counts live + tombstone slots
    int mShift = 60;
};

} // namespace detail

template <typename EntityId, PartitioningMethod Method = PartitioningMethod::None>
class Registry
{
  public:
    struct EntityCollision
    {
        EntityId entityA;
        EntityId entityB;
        ShapeId shapeA;
        ShapeId shapeB;
        Manifold manifold;
    };

    using RayHit = RaycastHit<std::pair<EntityId, ShapeId>>;

    Registry() { init(0); }
    Registry(uint32_t cellSize) requires(Method == PartitioningMethod::Grid) { init(cellSize); }
    ~Registry() { c2d_gpu_destroy(mCtx); }

    Registry(const Registry&) = delete;
    Registry& operator=(const Registry&) = delete;
    Registry(Registry&& o) noexcept { *this = std::move(o); }
    Registry& operator=(Registry&& o) noexcept
    {
        if (this != &o)
        {
            c2d_gpu_destroy(mCtx);
            mCtx = std::exchange(o.mCtx, nullptr);
            mIndex = std::move(o.mIndex);
            mEntities = std::move(o.mEntities);
            mFreeEntitySlots = std::move(o.mFreeEntitySlots);
            mLiveEntities = std::exchange(o.mLiveEntities, 0);
            mShapes = std::move(o.mShapes);
            mFreeShapeIds = std::move(o.mFreeShapeIds);
            mStaged = std::move(o.mStaged);
            mPendingApplies = std::move(o.mPendingApplies);
            mStructureEpoch = o.mStructureEpoch;
            mStats = o.mStats;
            mPreviousAllCollisionsCount = o.mPreviousAllCollisionsCount;
            mCellSize = o.mCellSize;
            mCollisionCache = CollisionCache{};
        }
        return *this;
    }

    void createEntity(EntityId id, BodyType type, const Transform& transform = Transform{});
    void removeEntity(EntityId id);

    template <IsShape Shape>
    ShapeId addShape(EntityId entityId, const Shape& shape, uint64_t categoryBits = 1, uint64_t maskBits = ~0ull);
    void removeShape(ShapeId shapeId);
    void setShapeActive(ShapeId shapeId, bool isActive = true);

    void teleportEntity(EntityId id, const Transform& transform);
    void teleportEntity(EntityId id, Translation translation);
    void moveEntity(EntityId id, const Transform& delta);
    void moveEntity(EntityId id, Translation deltaTranslation);
    Transform getEntityTransform(EntityId id) const;

    std::vector<EntityCollision> getCollidingPairs();
    std::vector<EntityCollision> getCollisions(EntityId id) const;
    bool areColliding(EntityId a, EntityId b) const;

    std::optional<RayHit> firstHitRayCast(Ray ray, BitMaskType maskBits = ~0ull) const;
    std::optional<RayHit> firstHitRayCast(InfiniteRay ray, BitMaskType maskBits = ~0ull) const;
    std::set<RayHit> rayCast(Ray ray, BitMaskType maskBits = ~0ull) const;
    std::set<RayHit> rayCast(InfiniteRay ray, BitMaskType maskBits = ~0ull) const;

    // Binary snapshots in the reference's wire format (reference Registry.hpp:1205-1344; Grid: BroadPhaseTree.hpp:539-733).
    void serialize(std::ostream& out) const;
    static Registry deserialize(std::istream& in);
    // Same logical state: entities, transforms, shapes, category / mask bits, activity, tight and leaf boxes, free ids.
    bool operator==(const Registry& other) const;
    bool operator!=(const Registry& other) const { return !(*this == other); }

    size_t size() const { return mLiveEntities; }
    void clear();

    // ---- extensions (not in the reference) ------------------------------------------------------------
    // moveEntity(ids[i], deltas[i]) for every i, in order.
    void moveEntities(std::span<const EntityId> ids, std::span<const Translation> deltas);
    // teleportEntity(ids[i], positions[i]) / teleportEntity(ids[i], transforms[i]) for every i, in order.  An entity may
    // appear once per call (the translation variant resolves all deltas against the transforms at entry).
    void teleportEntities(std::span<const EntityId> ids, std::span<const Translation> positions);
    void teleportEntities(std::span<const EntityId> ids, std::span<const Transform> transforms);
    // getCollisions(id) for many entities in ONE device pass (a single call costs a ~30 us round trip whatever the
    // scene size): the concatenated results; entityA of a record names the queried entity.
    std::vector<EntityCollision> getCollisions(std::span<const EntityId> ids) const;
    // rayCast for a batch of rays in ONE device pass (the single-ray calls above are latency bound: one thread
    // walks the trees); hits of ray i are result[i] in std::set order.
    std::vector<std::vector<RayHit>> rayCastBatch(std::span<const Ray> rays, std::span<const InfiniteRay> infiniteRays,
                                                  BitMaskType maskBits = ~0ull) const;
    // Mixed batch with one mask per ray: `second` is the end point (finite) or the direction (infinite).
    struct RayQuery
    {
        Vec2 start;
        Vec2 second;
        BitMaskType maskBits = ~0ull;
        bool infinite = false;
    };
    // Flat result: hits of query i are hits[offsets[i] .. offsets[i+1]).
    struct RayBatchResult
    {
        std::vector<uint64_t> offsets;
        std::vector<RayHit> hits;
    };
    RayBatchResult rayCastBatch(std::span<const RayQuery> queries) const;
    // Device-resident move batches: stageMoves() uploads moveEntity(ids[i], deltas[i]) for all i ONCE and returns
    // a handle; every applyStagedMoves(handle) then performs those moves with O(1) host work and no PCIe
    // traffic (the host-side transforms are brought up to date lazily, on the next call that reads them).
    // A batch names each entity at most once and is invalidated by removeShape / removeEntity / clear.
    using MoveBatch = uint32_t;
    MoveBatch stageMoves(std::span<const EntityId> ids, std::span<const Translation> deltas);
    void applyStagedMoves(MoveBatch batch);
    void releaseStagedMoves(MoveBatch batch);
    // getCollidingPairs() without the copy into a fresh std::vector: a view of the records in the library's pinned
    // result buffer (or, for EntityIds that are not 4 trivially-copyable bytes, in an internal vector), valid until
    // the next query on this Registry.
    std::span<const EntityCollision> getCollidingPairsView();
    // Multi-GPU slabs: marks a shape as the halo copy of a shape owned by another rank's Registry (see
    // c2d_gpu_shape_set_ghost in c2d_gpu.h for the owner rule that keeps the union of the ranks' results duplicate free).
    // The owner rule compares entity tags across ranks, so the tags must be global: only EntityIds that travel as their
    // own 32-bit pattern qualify (any other id type is a rank-local dense index on the device).
    void setShapeGhost(ShapeId shapeId, bool isGhost = true)
    {
        static_assert(kDirectTag, "setShapeGhost needs a 4-byte trivially copyable EntityId (global entity tags)");
        check(c2d_gpu_shape_set_ghost(mCtx, shapeId, isGhost ? 1 : 0));
    }
    // ---- multi-GPU group (include/c2d_gpu.h "multi-GPU") --------------------------------
    // One Registry per GPU (one process per GPU, or several Registries in one process), all holding the SAME scene
    // (every rank replays every mutation).  getCollidingPairs() is then collective: each rank sorts, builds and
    // traverses only its Morton-key slab of the Dynamic tree and returns ITS share of the pairs; the union over the
    // ranks is the reference's pair set, each pair once.  The id comes from createGroupId() on one rank and travels to
    // the others by any channel (MPI, torch.distributed, a file).
    using GroupId = std::array<unsigned char, 128>;
    static GroupId createGroupId()
    {
        GroupId id{};
        if (c2d_gpu_comm_unique_id(id.data()) != 0)
            throw std::runtime_error(std::string("colli2de-b200: ") + c2d_gpu_last_error(nullptr));
        return id;
    }
    void joinGroup(uint32_t rank, uint32_t world, const GroupId& id)
    {
        check(c2d_gpu_comm_init(mCtx, rank, world, id.data()));
        mGroupWorld = world;
    }
    void leaveGroup()
    {
        syncHost();
        check(c2d_gpu_comm_set_result_mode(mCtx, C2D_RESULTS_LOCAL));
        check(c2d_gpu_comm_destroy(mCtx));
        check(c2d_gpu_set_shard(mCtx, 0, 1));
        check(c2d_gpu_set_ray_sharding(mCtx, 0));
        mGroupWorld = 1;
        mGroupMoveExchange = false;
    }
    // Which records getCollidingPairs() returns on this rank: its own share (default), everybody's on every rank, or
    // everybody's on rank 0 (the other ranks keep their share).  Gathered over NCCL device to device.
    enum class GroupResults
    {
        Local = C2D_RESULTS_LOCAL,
        All = C2D_RESULTS_ALL,
        Root = C2D_RESULTS_ROOT
    };
    void setGroupResults(GroupResults mode) { check(c2d_gpu_comm_set_result_mode(mCtx, static_cast<int>(mode))); }
    // Move exchange: moveEntity(id, Translation) is called for an entity on ONE rank only (the rank that simulates it),
    // at most once per frame; the next getCollidingPairs() all-gathers the 12-byte records over NVLink and applies them
    // on every rank.  All other mutators stay replicated.  The host mirror of the entities other ranks moved is brought
    // up to date lazily (one download of the device transforms) by the next call that reads a transform.
    void setGroupMoveExchange(bool on)
    {
        syncHost();
        check(c2d_gpu_comm_set_move_exchange(mCtx, on ? 1 : 0));
        mGroupMoveExchange = on;
    }
    // Ray sharding: every rank calls rayCastBatch with the SAME batch; this rank casts its contiguous share of it and
    // returns empty hit lists for the other rays (the per-ray lists of the ranks concatenate to the full result).
    void setGroupRaySharding(bool on) { check(c2d_gpu_set_ray_sharding(mCtx, on ? 1 : 0)); }
    c2d_gpu_comm_stats lastGroupStats() const
    {
        c2d_gpu_comm_stats st{};
        check(c2d_gpu_comm_get_stats(mCtx, &st));
        return st;
    }
    // getCollidingPairs() returns its records in the reference's order (its five body-type groups one after the other,
    // each sorted by (shapeA, shapeB); Registry.hpp:487-545), byte-identical from run to run.  Off by default: the
    // unordered result saves two radix sorts and a gather per frame.
    void setDeterministicOrder(bool on) { check(c2d_gpu_set_result_order(mCtx, on ? 1 : 0)); }
    // Counters and CUDA-event timings of the last getCollidingPairs() / rayCast().
    const c2d_gpu_stats& lastQueryStats() const { return mStats; }
    // Runs the pipeline but leaves the records on the device; returns the number of colliding pairs.
    uint64_t countCollidingPairsOnDevice();
    c2d_gpu_ctx* nativeContext() const { return mCtx; }

  private:
    static constexpr uint32_t kNone = 0xffffffffu;
    // EntityIds that are 4 trivially-copyable bytes travel through the GPU as their own bit pattern, so the
    // 84-byte records come back already in EntityCollision layout; any other id type uses the dense index.
    static constexpr bool kDirectTag = sizeof(EntityId) == 4 && std::is_trivially_copyable_v<EntityId> &&
                                       std::has_unique_object_representations_v<EntityId>;

    struct EntityInfo
    {
        EntityId id{};
        Transform transform{};
        Transform previous{}; // bullets only (reference mBulletPreviousTransforms)
        uint32_t firstShape = kNone;
        uint32_t lastShape = kNone;
        BodyType type = BodyType::Static;
        bool alive = false;
    };
    struct ShapeLink
    {
        uint32_t entity = kNone; // dense index
        uint32_t prev = kNone, next = kNone;
    };
    struct KeyOf
    {
        const std::vector<EntityInfo>* entities;
        const EntityId& operator()(uint32_t index) const { return (*entities)[index].id; }
    };

    void init(uint32_t cellSize)
    {
        if (c2d_gpu_create(-1, &mCtx) != 0)
            throw std::runtime_error(std::string("colli2de-b200: ") + c2d_gpu_last_error(nullptr));
        if constexpr (Method == PartitioningMethod::Grid)
        {
            mCellSize = cellSize ? cellSize : 120; // BroadPhaseTree's default (reference BroadPhaseTree.hpp:96)
            check(c2d_gpu_set_broadphase(mCtx, C2D_BROADPHASE_GRID, mCellSize));
        }
    }
    snapshot::RegistryState<EntityId> captureState() const;
    void check(int rc) const
    {
        if (rc != 0)
            throw std::runtime_error(std::string("colli2de-b200: ") + c2d_gpu_last_error(mCtx));
    }
    uint32_t indexOf(const EntityId& id) const { return mIndex.find(id, KeyOf{&mEntities}); }
    EntityInfo& entityAt(const EntityId& id)
    {
        const uint32_t idx = indexOf(id);
        C2D_DEBUG_ASSERT(idx != kNone, "Entity with this ID does not exist");
        return mEntities[idx];
    }
    uint32_t tagOf(uint32_t denseIndex) const
    {
        if constexpr (kDirectTag)
            return std::bit_cast<uint32_t>(mEntities[denseIndex].id);
        else
            return denseIndex;
    }
    static c2d_xf pack(const Transform& t, bool withLibmTrig)
    {
        c2d_xf x;
        x.tx = t.translation.x;
        x.ty = t.translation.y;
        x.angle = t.rotation.angleRadians;
        x.sin = t.rotation.sin;
        x.cos = t.rotation.cos;
        // what Transform(translation, angle) would hold: used by sweeps, which rebuild sin/cos from the angle
        x.csin = withLibmTrig ? std::sin(t.rotation.angleRadians) : t.rotation.sin;
        x.ccos = withLibmTrig ? std::cos(t.rotation.angleRadians) : t.rotation.cos;
        x._pad = 0.0f;
        return x;
    }
    static void packGeometry(const Circle& s, float* g, uint32_t& count)
    {
        g[0] = s.center.x, g[1] = s.center.y, g[2] = s.radius;
        count = 0;
    }
    static void packGeometry(const Capsule& s, float* g, uint32_t& count)
    {
        g[0] = s.center1.x, g[1] = s.center1.y, g[2] = s.center2.x, g[3] = s.center2.y, g[4] = s.radius;
        count = 0;
    }
    static void packGeometry(const Segment& s, float* g, uint32_t& count)
    {
        g[0] = s.start.x, g[1] = s.start.y, g[2] = s.end.x, g[3] = s.end.y;
        count = 0;
    }
    static void packGeometry(const Polygon& s, float* g, uint32_t& count)
    {
        count = s.count;
        for (uint32_t i = 0; i < s.count && i < MAX_POLYGON_VERTICES; ++i)
        {
            g[2 * i] = s.vertices[i].x, g[2 * i + 1] = s.vertices[i].y;
            g[16 + 2 * i] = s.normals[i].x, g[16 + 2 * i + 1] = s.normals[i].y;
        }
    }
    void pushShapeTransforms(EntityInfo& entity)
    {
        const c2d_xf x = pack(entity.transform, entity.type == BodyType::Bullet);
        for (uint32_t s = entity.firstShape; s != kNone; s = mShapes[s].next)
            check(c2d_gpu_shape_move(mCtx, s, &x));
    }
    template <typename RayT>
    std::set<RayHit> rayCastOne(RayT ray, BitMaskType maskBits, bool infinite) const;
    template <typename RayT>
    std::optional<RayHit> firstHitOne(RayT ray, BitMaskType maskBits, bool infinite) const;
    template <typename RayT>
    static c2d_ray packRay(RayT ray, BitMaskType maskBits, bool infinite)
    {
        c2d_ray r{};
        r.sx = ray.start.x;
        r.sy = ray.start.y;
        if constexpr (std::is_same_v<RayT, Ray>)
            r.ex = ray.end.x, r.ey = ray.end.y;
        else
            r.ex = ray.direction.x, r.ey = ray.direction.y;
        r.mask = maskBits;
        r.infinite = infinite ? 1u : 0u;
        return r;
    }
    RayHit toRayHit(const c2d_hit_record& h) const
    {
        EntityId id;
        if constexpr (kDirectTag)
            id = std::bit_cast<EntityId>(h.entity);
        else
            id = mEntities[h.entity].id;
        return RayHit{{id, h.shape}, Vec2{h.entry[0], h.entry[1]}, Vec2{h.exit[0], h.exit[1]}, h.entryTime, h.exitTime};
    }
    EntityCollision toCollision(const c2d_pair_record& r) const
    {
        EntityId a, b;
        if constexpr (kDirectTag)
            a = std::bit_cast<EntityId>(r.entityA), b = std::bit_cast<EntityId>(r.entityB);
        else
            a = mEntities[r.entityA].id, b = mEntities[r.entityB].id;
        EntityCollision c{a, b, r.shapeA, r.shapeB, Manifold{}};
        static_assert(sizeof(Manifold) == sizeof(c2d_manifold));
        std::memcpy(static_cast<void*>(&c.manifold), &r.manifold, sizeof(Manifold));
        return c;
    }
    std::vector<EntityCollision> toCollisions(const c2d_pair_record* records, uint64_t count) const
    {
        if constexpr (kDirectTag && sizeof(EntityCollision) == sizeof(c2d_pair_record) &&
                      std::is_trivially_copyable_v<EntityCollision>)
        {
            // the device wrote EntityCollision records: one block copy out of the pinned buffer
            const auto* first = reinterpret_cast<const EntityCollision*>(records);
            return std::vector<EntityCollision>(first, first + count);
        }
        else
        {
            std::vector<EntityCollision> out;
            out.reserve(count);
            for (uint64_t i = 0; i < count; ++i)
            {
                const c2d_pair_record& r = records[i];
                EntityCollision c{mEntities[r.entityA].id, mEntities[r.entityB].id, r.shapeA, r.shapeB, Manifold{}};
                static_assert(sizeof(Manifold) == sizeof(c2d_manifold));
                std::memcpy(static_cast<void*>(&c.manifold), &r.manifold, sizeof(Manifold));
                out.push_back(std::move(c));
            }
            return out;
        }
    }

    struct StagedMoves
    {
        uint32_t handle = 0;
        uint64_t epoch = 0;
        bool live = false;
        std::vector<uint32_t> entityIndex;
        std::vector<Translation> deltas;
    };
    // applies the staged batches that were replayed on the device but not yet on the host-side transforms
    void syncHost() const
    {
        for (const MoveBatch b : mPendingApplies)
        {
            const StagedMoves& batch = mStaged[b];
            for (size_t i = 0; i < batch.entityIndex.size(); ++i)
            {
                EntityInfo& e = mEntities[batch.entityIndex[i]];
                if (e.type == BodyType::Bullet)
                    e.previous = e.transform;
                e.transform.translation += batch.deltas[i];
            }
        }
        mPendingApplies.clear();
        if (mRemoteMovesPending)
            pullTransforms();
    }
    // move exchange: other ranks' translations reached the device, not this host mirror - read the transforms back
    // (the first shape of an entity carries the entity's transform; an entity without shapes is never exchanged)
    void pullTransforms() const
    {
        mRemoteMovesPending = false;
        const uint32_t n = static_cast<uint32_t>(mShapes.size());
        if (n == 0)
            return;
        std::vector<float> xf(4 * static_cast<size_t>(n)), pxf(xf.size());
        check(c2d_gpu_read_transforms(mCtx, 0, n, xf.data(), nullptr, pxf.data(), nullptr));
        for (EntityInfo& e : mEntities)
        {
            if (!e.alive || e.firstShape == kNone || e.type == BodyType::Static)
                continue;
            const size_t s = 4 * static_cast<size_t>(e.firstShape);
            e.transform.translation = Vec2{xf[s], xf[s + 1]}; // translations only: the rotation never travels
            if (e.type == BodyType::Bullet)
                e.previous.translation = Vec2{pxf[s], pxf[s + 1]};
        }
    }

    c2d_gpu_ctx* mCtx = nullptr;
    uint32_t mGroupWorld = 1;
    bool mGroupMoveExchange = false;
    mutable bool mRemoteMovesPending = false;
    detail::IndexMap<EntityId, KeyOf> mIndex;
    mutable std::vector<EntityInfo> mEntities;
    std::vector<StagedMoves> mStaged;
    mutable std::vector<MoveBatch> mPendingApplies;
    uint64_t mStructureEpoch = 0;
    std::vector<uint32_t> mFreeEntitySlots;
    size_t mLiveEntities = 0;
    std::vector<ShapeLink> mShapes;
    std::vector<ShapeId> mFreeShapeIds;
    mutable c2d_gpu_stats mStats{};
    std::vector<EntityCollision> mViewStorage;
    std::vector<uint32_t> mBulkShapes; // scratch of moveEntities / teleportEntities
    std::vector<float> mBulkDeltas;
    std::vector<Translation> mBulkTargets;
    uint32_t mPreviousAllCollisionsCount = 1024; // reference Registry.hpp:168
    // memoised getCollisions(id) results, valid while the library's mutation epoch does not change
    static constexpr uint32_t kCollisionCacheWarmup = 4;
    struct CollisionCache
    {
        uint64_t epoch = ~0ull;
        uint32_t singleCalls = 0;
        bool built = false;
        std::vector<uint64_t> offsets; // per dense entity index
        std::vector<EntityCollision> records;
    };
    mutable CollisionCache mCollisionCache;
    uint32_t mCellSize = 0;                      // PartitioningMethod::Grid only
};

// ---------------------------------------------------------------------------------------------------------
// mutators
// ---------------------------------------------------------------------------------------------------------
template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::createEntity(EntityId id, BodyType type, const Transform& transform)
{
    C2D_DEBUG_ASSERT(indexOf(id) == kNone, "Entity with this ID already exists");
    uint32_t idx;
    if (!mFreeEntitySlots.empty())
    {
        idx = mFreeEntitySlots.back();
        mFreeEntitySlots.pop_back();
    }
    else
    {
        idx = static_cast<uint32_t>(mEntities.size());
        mEntities.emplace_back();
    }
    EntityInfo& e = mEntities[idx];
    e = EntityInfo{};
    e.id = id;
    e.transform = transform;
    e.previous = transform; // mBulletPreviousTransforms.emplace(id, transform) (reference Registry.hpp:268-269)
    e.type = type;
    e.alive = true;
    ++mLiveEntities;
    if (mIndex.needsRehash())
    {
        std::vector<uint32_t> live;
        live.reserve(mLiveEntities);
        for (uint32_t i = 0; i < mEntities.size(); ++i)
            if (mEntities[i].alive && i != idx)
                live.push_back(i);
        mIndex.rehash(KeyOf{&mEntities}, live);
    }
    mIndex.insert(id, idx);
}

template <typename EntityId, PartitioningMethod Method>
template <IsShape Shape>
ShapeId Registry<EntityId, Method>::addShape(EntityId entityId, const Shape& shape, uint64_t categoryBits,
                                             uint64_t maskBits)
{
    syncHost();
    // a live staged batch captured the entity's shape list: a new shape would be left behind by applyStagedMoves()
    ++mStructureEpoch;
    const uint32_t idx = indexOf(entityId);
    C2D_DEBUG_ASSERT(idx != kNone, "Entity with this ID does not exist");
    EntityInfo& entity = mEntities[idx];

    ShapeId shapeId;
    if (!mFreeShapeIds.empty())
    {
        shapeId = mFreeShapeIds.back(); // LIFO reuse (reference Registry.hpp:283-289)
        mFreeShapeIds.pop_back();
    }
    else
    {
        shapeId = static_cast<ShapeId>(mShapes.size());
        mShapes.emplace_back();
    }

    float geom[C2D_GEOM_FLOATS] = {};
    uint32_t count = 0;
    packGeometry(shape, geom, count);
    const bool bullet = entity.type == BodyType::Bullet;
    const c2d_xf cur = pack(entity.transform, bullet);
    const c2d_xf prev = pack(entity.previous, bullet);
    check(c2d_gpu_shape_add(mCtx, shapeId, tagOf(idx), static_cast<uint32_t>(shape.getType()),
                            static_cast<uint32_t>(entity.type), count, categoryBits, maskBits, geom, &cur, &prev));

    ShapeLink& link = mShapes[shapeId];
    link.entity = idx;
    link.prev = entity.lastShape;
    link.next = kNone;
    if (entity.lastShape != kNone)
        mShapes[entity.lastShape].next = shapeId;
    else
        entity.firstShape = shapeId;
    entity.lastShape = shapeId;
    return shapeId;
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::removeShape(ShapeId shapeId)
{
    if (shapeId >= mShapes.size()) // the reference indexes with .at() (Registry.hpp:334)
        throw std::out_of_range("colli2de-b200: removeShape: shape id out of range");
    syncHost();
    ++mStructureEpoch;
    ShapeLink& link = mShapes[shapeId];
    C2D_DEBUG_ASSERT(link.entity != kNone, "Shape with this ID does not exist in the entity");
    EntityInfo& entity = mEntities[link.entity];
    check(c2d_gpu_shape_remove(mCtx, shapeId));
    if (link.prev != kNone)
        mShapes[link.prev].next = link.next;
    else
        entity.firstShape = link.next;
    if (link.next != kNone)
        mShapes[link.next].prev = link.prev;
    else
        entity.lastShape = link.prev;
    link = ShapeLink{};
    mFreeShapeIds.push_back(shapeId);
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::setShapeActive(ShapeId shapeId, bool isActive)
{
    if (shapeId >= mShapes.size())
        throw std::out_of_range("colli2de-b200: setShapeActive: shape id out of range");
    check(c2d_gpu_shape_set_active(mCtx, shapeId, isActive ? 1 : 0));
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::removeEntity(EntityId id)
{
    syncHost();
    ++mStructureEpoch;
    const uint32_t idx = indexOf(id);
    C2D_DEBUG_ASSERT(idx != kNone, "Entity with this ID does not exist");
    EntityInfo& entity = mEntities[idx];
    for (uint32_t s = entity.firstShape; s != kNone;)
    {
        const uint32_t next = mShapes[s].next;
        check(c2d_gpu_shape_remove(mCtx, s));
        mShapes[s] = ShapeLink{};
        mFreeShapeIds.push_back(s); // in insertion order, like the reference (Registry.hpp:367-373)
        s = next;
    }
    mIndex.erase(id, KeyOf{&mEntities});
    entity = EntityInfo{};
    mFreeEntitySlots.push_back(idx);
    --mLiveEntities;
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::teleportEntity(EntityId id, const Transform& transform)
{
    syncHost();
    EntityInfo& entity = entityAt(id);
    // reference Registry.hpp:382-401: shapes are refit at `transform`, then previous <- current <- transform
    if (entity.type == BodyType::Bullet)
        entity.previous = entity.transform;
    entity.transform = transform;
    pushShapeTransforms(entity);
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::teleportEntity(EntityId id, Translation translation)
{
    syncHost();
    const EntityInfo& entity = entityAt(id);
    moveEntity(id, translation - entity.transform.translation); // reference Registry.hpp:403-411
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::moveEntity(EntityId id, const Transform& delta)
{
    syncHost();
    EntityInfo& entity = entityAt(id);
    if (entity.type == BodyType::Bullet)
        entity.previous = entity.transform;
    entity.transform += delta; // Rotation::operator+= quirk included
    pushShapeTransforms(entity);
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::moveEntity(EntityId id, Translation d)
{
    if (!mPendingApplies.empty())
        syncHost();
    EntityInfo& entity = entityAt(id);
    for (uint32_t s = entity.firstShape; s != kNone; s = mShapes[s].next)
        check(c2d_gpu_shape_translate(mCtx, s, d.x, d.y));
    if (entity.type == BodyType::Bullet)
        entity.previous = entity.transform;
    entity.transform.translation += d;
}

// moveEntity(ids[i], deltas[i]) for every i, in order (reference Registry.hpp:434-453), as ONE pass: the index lookups
// are software-pipelined (a million scattered lookups are cache-miss bound) and the per-shape records reach the library
// in one c2d_gpu_shapes_translate call instead of one C-ABI call per shape.
template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::moveEntities(std::span<const EntityId> ids, std::span<const Translation> deltas)
{
    C2D_DEBUG_ASSERT(ids.size() == deltas.size());
    if (!mPendingApplies.empty())
        syncHost();
    const size_t n = ids.size();
    mBulkShapes.clear();
    mBulkDeltas.clear();
    mBulkShapes.reserve(n);
    mBulkDeltas.reserve(2 * n);
    constexpr size_t kAhead = 16;
    for (size_t i = 0; i < n; ++i)
    {
        if (i + 2 * kAhead < n)
            mIndex.prefetch(ids[i + 2 * kAhead]);
        if (i + kAhead < n)
        {
            const uint32_t hint = mIndex.peek(ids[i + kAhead]);
            if (hint < mEntities.size())
                __builtin_prefetch(&mEntities[hint]);
        }
        const uint32_t idx = indexOf(ids[i]);
        C2D_DEBUG_ASSERT(idx != kNone, "Entity with this ID does not exist");
        EntityInfo& entity = mEntities[idx];
        const Translation d = deltas[i];
        for (uint32_t s = entity.firstShape; s != kNone; s = mShapes[s].next)
        {
            mBulkShapes.push_back(s);
            mBulkDeltas.push_back(d.x);
            mBulkDeltas.push_back(d.y);
        }
        if (entity.type == BodyType::Bullet)
            entity.previous = entity.transform;
        entity.transform.translation += d;
    }
    check(c2d_gpu_shapes_translate(mCtx, static_cast<uint32_t>(mBulkShapes.size()), mBulkShapes.data(), mBulkDeltas.data()));
}

// teleportEntity(ids[i], positions[i]) for every i (reference Registry.hpp:403-411: a move by position - current).
template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::teleportEntities(std::span<const EntityId> ids, std::span<const Translation> positions)
{
    C2D_DEBUG_ASSERT(ids.size() == positions.size());
    syncHost();
    mBulkTargets.resize(ids.size());
    for (size_t i = 0; i < ids.size(); ++i)
        mBulkTargets[i] = positions[i] - entityAt(ids[i]).transform.translation;
    moveEntities(ids, std::span<const Translation>(mBulkTargets));
}

// teleportEntity(ids[i], transforms[i]) for every i (reference Registry.hpp:382-401).
template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::teleportEntities(std::span<const EntityId> ids, std::span<const Transform> transforms)
{
    C2D_DEBUG_ASSERT(ids.size() == transforms.size());
    syncHost();
    for (size_t i = 0; i < ids.size(); ++i)
    {
        EntityInfo& entity = entityAt(ids[i]);
        if (entity.type == BodyType::Bullet)
            entity.previous = entity.transform;
        entity.transform = transforms[i];
        pushShapeTransforms(entity);
    }
}

template <typename EntityId, PartitioningMethod Method>
typename Registry<EntityId, Method>::MoveBatch Registry<EntityId, Method>::stageMoves(
    std::span<const EntityId> ids, std::span<const Translation> deltas)
{
    C2D_DEBUG_ASSERT(ids.size() == deltas.size());
    syncHost();
    StagedMoves batch;
    batch.epoch = mStructureEpoch;
    batch.live = true;
    batch.entityIndex.reserve(ids.size());
    batch.deltas.assign(deltas.begin(), deltas.end());
    std::vector<uint32_t> shapes;
    std::vector<float> dxy;
    shapes.reserve(ids.size());
    dxy.reserve(ids.size() * 2);
    for (size_t i = 0; i < ids.size(); ++i)
    {
        const uint32_t idx = indexOf(ids[i]);
        C2D_DEBUG_ASSERT(idx != kNone, "Entity with this ID does not exist");
        batch.entityIndex.push_back(idx);
        for (uint32_t s = mEntities[idx].firstShape; s != kNone; s = mShapes[s].next)
        {
            shapes.push_back(s);
            dxy.push_back(deltas[i].x);
            dxy.push_back(deltas[i].y);
        }
    }
    check(c2d_gpu_stage_translations(mCtx, static_cast<uint32_t>(shapes.size()), shapes.data(), dxy.data(),
                                     &batch.handle));
    MoveBatch slot = 0;
    while (slot < mStaged.size() && mStaged[slot].live)
        ++slot;
    if (slot == mStaged.size())
        mStaged.emplace_back();
    mStaged[slot] = std::move(batch);
    return slot;
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::applyStagedMoves(MoveBatch batch)
{
    if (batch >= mStaged.size() || !mStaged[batch].live)
        throw std::runtime_error("colli2de-b200: applyStagedMoves: unknown batch");
    if (mStaged[batch].epoch != mStructureEpoch)
        throw std::runtime_error("colli2de-b200: applyStagedMoves: shapes were added or removed since the batch was staged");
    check(c2d_gpu_apply_staged(mCtx, mStaged[batch].handle));
    mPendingApplies.push_back(batch);
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::releaseStagedMoves(MoveBatch batch)
{
    if (batch >= mStaged.size() || !mStaged[batch].live)
        return;
    syncHost();
    const uint32_t handle = mStaged[batch].handle;
    mStaged[batch] = StagedMoves{};
    check(c2d_gpu_release_staged(mCtx, handle));
}

template <typename EntityId, PartitioningMethod Method>
Transform Registry<EntityId, Method>::getEntityTransform(EntityId id) const
{
    syncHost();
    const uint32_t idx = indexOf(id);
    C2D_DEBUG_ASSERT(idx != kNone);
    return mEntities[idx].transform;
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::clear()
{
    check(c2d_gpu_clear(mCtx));
    mPreviousAllCollisionsCount = 1000; // reference Registry.hpp:1202
    ++mStructureEpoch;
    mStaged.clear();
    mPendingApplies.clear();
    mIndex.clear();
    mEntities.clear();
    mFreeEntitySlots.clear();
    mLiveEntities = 0;
    mShapes.clear();
    mFreeShapeIds.clear();
}

// ---------------------------------------------------------------------------------------------------------
// queries
// ---------------------------------------------------------------------------------------------------------
template <typename EntityId, PartitioningMethod Method>
std::vector<typename Registry<EntityId, Method>::EntityCollision> Registry<EntityId, Method>::getCollidingPairs()
{
    if constexpr (kDirectTag && sizeof(EntityCollision) == sizeof(c2d_pair_record) &&
                  std::is_trivially_copyable_v<EntityCollision> && detail::kHasUninitializedResize)
    {
        // the device writes EntityCollision records: reserve the result from the previous frame's count like the
        // reference does (Registry.hpp:467, 547), let the library touch its pages while the GPU works, then have
        // the records copied straight into it
        check(c2d_gpu_collide_begin(mCtx));
        mRemoteMovesPending |= mGroupMoveExchange;
        std::vector<EntityCollision> out;
        out.reserve(static_cast<std::size_t>(mPreviousAllCollisionsCount) + mPreviousAllCollisionsCount / 8 + 64);
        check(c2d_gpu_prefault(mCtx, static_cast<void*>(out.data()), out.capacity() * sizeof(EntityCollision)));
        uint64_t count = 0;
        check(c2d_gpu_collide_end(mCtx, &count, &mStats));
        detail::resizeUninitialized(out, count);
        check(c2d_gpu_collide_fetch(mCtx, reinterpret_cast<c2d_pair_record*>(out.data()), count, &mStats));
        mPreviousAllCollisionsCount = static_cast<uint32_t>(count);
        return out;
    }
    else
    {
        const c2d_pair_record* records = nullptr;
        uint64_t count = 0;
        check(c2d_gpu_collide(mCtx, &records, &count, &mStats));
    mRemoteMovesPending |= mGroupMoveExchange;
        mPreviousAllCollisionsCount = static_cast<uint32_t>(count);
        return toCollisions(records, count);
    }
}

template <typename EntityId, PartitioningMethod Method>
std::span<const typename Registry<EntityId, Method>::EntityCollision> Registry<EntityId, Method>::getCollidingPairsView()
{
    const c2d_pair_record* records = nullptr;
    uint64_t count = 0;
    check(c2d_gpu_collide(mCtx, &records, &count, &mStats));
    mRemoteMovesPending |= mGroupMoveExchange;
    mPreviousAllCollisionsCount = static_cast<uint32_t>(count);
    if constexpr (kDirectTag && sizeof(EntityCollision) == sizeof(c2d_pair_record) &&
                  std::is_trivially_copyable_v<EntityCollision>)
        return std::span<const EntityCollision>(reinterpret_cast<const EntityCollision*>(records), count);
    else
    {
        mViewStorage = toCollisions(records, count);
        return std::span<const EntityCollision>(mViewStorage);
    }
}

// Registry::getCollisions (reference Registry.hpp:551-605): every active shape of the entity queries the trees
// with its tight box and its mask bits; order of the result is unspecified.
//
// A single call is a device round trip (~30 us whatever the scene size), three orders of magnitude above the
// reference's pointer walk on a small scene.  Applications (and the reference's own benchmark) call it in loops over
// many entities between mutations, so the results are memoised: once the single calls of an epoch (no intervening
// mutation, c2d_gpu_mutation_epoch) have cost about what one pass over ALL entities costs, the next call asks the
// device for everybody's collisions, keeps them grouped by entity on the host, and every further call until the next
// mutation is a copy out of that table.
template <typename EntityId, PartitioningMethod Method>
std::vector<typename Registry<EntityId, Method>::EntityCollision> Registry<EntityId, Method>::getCollisions(
    EntityId id) const
{
    const uint32_t idx = indexOf(id);
    C2D_DEBUG_ASSERT(idx != kNone);
    uint64_t epoch = 0;
    check(c2d_gpu_mutation_epoch(mCtx, &epoch));
    CollisionCache& cache = mCollisionCache;
    if (cache.epoch != epoch)
    {
        cache.epoch = epoch;
        cache.singleCalls = 0;
        cache.built = false;
    }
    // ski rental: a single call costs one device round trip, the table roughly one per 128 live shapes — switch once the
    // calls of this epoch have cost about as much as the table would (never more than twice the best strategy)
    if (!cache.built && ++cache.singleCalls > std::max<uint64_t>(kCollisionCacheWarmup, mShapes.size() / 128))
    {
        // all live shapes in one device pass; records are (queried shape, other shape): group by the queried entity
        std::vector<uint32_t> shapes;
        shapes.reserve(mShapes.size());
        for (uint32_t s = 0; s < mShapes.size(); ++s)
            if (mShapes[s].entity != kNone)
                shapes.push_back(s);
        const c2d_pair_record* records = nullptr;
        uint64_t count = 0;
        check(c2d_gpu_query_shapes(mCtx, static_cast<uint32_t>(shapes.size()), shapes.data(), &records, &count));
        cache.offsets.assign(mEntities.size() + 1, 0);
        for (uint64_t i = 0; i < count; ++i)
            ++cache.offsets[mShapes[records[i].shapeA].entity + 1];
        for (size_t e = 0; e < mEntities.size(); ++e)
            cache.offsets[e + 1] += cache.offsets[e];
        std::vector<uint64_t> cursor(cache.offsets.begin(), cache.offsets.end() - 1);
        cache.records.resize(count);
        for (uint64_t i = 0; i < count; ++i)
            cache.records[cursor[mShapes[records[i].shapeA].entity]++] = toCollision(records[i]);
        cache.built = true;
    }
    if (cache.built && idx + 1 >= cache.offsets.size())
        return {}; // created after the table was built and still without shapes (addShape would have bumped the epoch)
    if (cache.built)
        return std::vector<EntityCollision>(cache.records.begin() + static_cast<std::ptrdiff_t>(cache.offsets[idx]),
                                            cache.records.begin() + static_cast<std::ptrdiff_t>(cache.offsets[idx + 1]));
    std::vector<uint32_t> shapes;
    for (uint32_t s = mEntities[idx].firstShape; s != kNone; s = mShapes[s].next)
        shapes.push_back(s);
    const c2d_pair_record* records = nullptr;
    uint64_t count = 0;
    check(c2d_gpu_query_shapes(mCtx, static_cast<uint32_t>(shapes.size()), shapes.data(), &records, &count));
    return toCollisions(records, count);
}

template <typename EntityId, PartitioningMethod Method>
std::vector<typename Registry<EntityId, Method>::EntityCollision> Registry<EntityId, Method>::getCollisions(
    std::span<const EntityId> ids) const
{
    std::vector<uint32_t> shapes;
    for (const EntityId& id : ids)
    {
        const uint32_t idx = indexOf(id);
        C2D_DEBUG_ASSERT(idx != kNone);
        for (uint32_t s = mEntities[idx].firstShape; s != kNone; s = mShapes[s].next)
            shapes.push_back(s);
    }
    const c2d_pair_record* records = nullptr;
    uint64_t count = 0;
    check(c2d_gpu_query_shapes(mCtx, static_cast<uint32_t>(shapes.size()), shapes.data(), &records, &count));
    return toCollisions(records, count);
}

// Registry::areColliding (reference Registry.hpp:748-864): all active shape pairs, no broadphase.
template <typename EntityId, PartitioningMethod Method>
bool Registry<EntityId, Method>::areColliding(EntityId a, EntityId b) const
{
    const uint32_t ia = indexOf(a), ib = indexOf(b);
    C2D_DEBUG_ASSERT(ia != kNone, "Entity A does not exist");
    C2D_DEBUG_ASSERT(ib != kNone, "Entity B does not exist");
    const EntityInfo& ea = mEntities[ia];
    const EntityInfo& eb = mEntities[ib];
    if (ea.type == BodyType::Static && eb.type == BodyType::Static)
        return false;
    std::vector<uint32_t> sa, sb;
    for (uint32_t x = ea.firstShape; x != kNone; x = mShapes[x].next)
        for (uint32_t y = eb.firstShape; y != kNone; y = mShapes[y].next)
        {
            sa.push_back(x);
            sb.push_back(y);
        }
    if (sa.empty())
        return false;
    std::vector<uint8_t> flags(sa.size(), 0);
    check(c2d_gpu_pairs_colliding(mCtx, static_cast<uint32_t>(sa.size()), sa.data(), sb.data(), flags.data()));
    for (uint8_t f : flags)
        if (f)
            return true;
    return false;
}

template <typename EntityId, PartitioningMethod Method>
uint64_t Registry<EntityId, Method>::countCollidingPairsOnDevice()
{
    uint64_t count = 0;
    check(c2d_gpu_collide_device(mCtx, &count, &mStats));
    mRemoteMovesPending |= mGroupMoveExchange;
    return count;
}

template <typename EntityId, PartitioningMethod Method>
template <typename RayT>
std::set<typename Registry<EntityId, Method>::RayHit> Registry<EntityId, Method>::rayCastOne(RayT ray,
                                                                                             BitMaskType maskBits,
                                                                                             bool infinite) const
{
    const c2d_ray r = packRay(ray, maskBits, infinite);
    const c2d_hit_record* hits = nullptr;
    const uint64_t* offsets = nullptr;
    check(c2d_gpu_raycast(mCtx, 1, &r, &hits, &offsets, &mStats));
    std::set<RayHit> out;
    for (uint64_t i = offsets[0]; i < offsets[1]; ++i)
        out.insert(out.end(), toRayHit(hits[i]));
    return out;
}

template <typename EntityId, PartitioningMethod Method>
template <typename RayT>
std::optional<typename Registry<EntityId, Method>::RayHit> Registry<EntityId, Method>::firstHitOne(RayT ray,
                                                                                                   BitMaskType maskBits,
                                                                                                   bool infinite) const
{
    const c2d_ray r = packRay(ray, maskBits, infinite);
    const c2d_hit_record* hits = nullptr;
    const uint64_t* offsets = nullptr;
    check(c2d_gpu_raycast_first(mCtx, 1, &r, &hits, &offsets, &mStats));
    if (offsets[1] == offsets[0])
        return std::nullopt;
    return toRayHit(hits[offsets[0]]);
}

template <typename EntityId, PartitioningMethod Method>
std::optional<typename Registry<EntityId, Method>::RayHit> Registry<EntityId, Method>::firstHitRayCast(
    Ray ray, BitMaskType maskBits) const
{
    return firstHitOne(ray, maskBits, false);
}

template <typename EntityId, PartitioningMethod Method>
std::optional<typename Registry<EntityId, Method>::RayHit> Registry<EntityId, Method>::firstHitRayCast(
    InfiniteRay ray, BitMaskType maskBits) const
{
    return firstHitOne(ray, maskBits, true);
}

template <typename EntityId, PartitioningMethod Method>
std::set<typename Registry<EntityId, Method>::RayHit> Registry<EntityId, Method>::rayCast(Ray ray,
                                                                                          BitMaskType maskBits) const
{
    return rayCastOne(ray, maskBits, false);
}

template <typename EntityId, PartitioningMethod Method>
std::set<typename Registry<EntityId, Method>::RayHit> Registry<EntityId, Method>::rayCast(InfiniteRay ray,
                                                                                          BitMaskType maskBits) const
{
    return rayCastOne(ray, maskBits, true);
}

template <typename EntityId, PartitioningMethod Method>
std::vector<std::vector<typename Registry<EntityId, Method>::RayHit>> Registry<EntityId, Method>::rayCastBatch(
    std::span<const Ray> rays, std::span<const InfiniteRay> infiniteRays, BitMaskType maskBits) const
{
    std::vector<c2d_ray> packed;
    packed.reserve(rays.size() + infiniteRays.size());
    for (const Ray& ray : rays)
        packed.push_back(c2d_ray{ray.start.x, ray.start.y, ray.end.x, ray.end.y, maskBits, 0u, 0u});
    for (const InfiniteRay& ray : infiniteRays)
        packed.push_back(c2d_ray{ray.start.x, ray.start.y, ray.direction.x, ray.direction.y, maskBits, 1u, 0u});
    const c2d_hit_record* hits = nullptr;
    const uint64_t* offsets = nullptr;
    check(c2d_gpu_raycast(mCtx, static_cast<uint32_t>(packed.size()), packed.data(), &hits, &offsets, &mStats));
    std::vector<std::vector<RayHit>> out(packed.size());
    for (size_t i = 0; i < packed.size(); ++i)
    {
        out[i].reserve(offsets[i + 1] - offsets[i]);
        for (uint64_t k = offsets[i]; k < offsets[i + 1]; ++k)
            out[i].push_back(toRayHit(hits[k]));
    }
    return out;
}

template <typename EntityId, PartitioningMethod Method>
typename Registry<EntityId, Method>::RayBatchResult Registry<EntityId, Method>::rayCastBatch(
    std::span<const RayQuery> queries) const
{
    std::vector<c2d_ray> packed;
    packed.reserve(queries.size());
    for (const RayQuery& q : queries)
        packed.push_back(c2d_ray{q.start.x, q.start.y, q.second.x, q.second.y, q.maskBits, q.infinite ? 1u : 0u, 0u});
    RayBatchResult out;
    if constexpr (kDirectTag && sizeof(RayHit) == sizeof(c2d_hit_record) && std::is_standard_layout_v<RayHit> &&
                  std::is_trivially_copy_constructible_v<RayHit> && std::is_trivially_destructible_v<RayHit> &&
                  detail::kHasUninitializedResize)
    {
        // the device writes RaycastHit records ({entity, shape}, entry, exit, entryTime, exitTime): they are copied
        // straight from the device into the result vector by the library's copy threads
        static_assert(offsetof(RayHit, entry) == offsetof(c2d_hit_record, entry) &&
                      offsetof(RayHit, exit) == offsetof(c2d_hit_record, exit) &&
                      offsetof(RayHit, entryTime) == offsetof(c2d_hit_record, entryTime) &&
                      offsetof(RayHit, exitTime) == offsetof(c2d_hit_record, exitTime));
        uint64_t total = 0;
        const uint64_t* offsets = nullptr;
        check(c2d_gpu_raycast_begin(mCtx, static_cast<uint32_t>(packed.size()), packed.data(), 0, &total, &offsets, &mStats));
        out.offsets.assign(offsets, offsets + packed.size() + 1);
        detail::resizeUninitialized(out.hits, total);
        check(c2d_gpu_raycast_fetch(mCtx, reinterpret_cast<c2d_hit_record*>(out.hits.data()), total, &mStats));
    }
    else
    {
        const c2d_hit_record* hits = nullptr;
        const uint64_t* offsets = nullptr;
        check(c2d_gpu_raycast(mCtx, static_cast<uint32_t>(packed.size()), packed.data(), &hits, &offsets, &mStats));
        out.offsets.assign(offsets, offsets + packed.size() + 1);
        out.hits.reserve(out.offsets.back());
        for (uint64_t k = 0; k < out.offsets.back(); ++k)
            out.hits.push_back(toRayHit(hits[k]));
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------------
// snapshots (reference Registry.hpp:1205-1344)
// ---------------------------------------------------------------------------------------------------------
template <typename EntityId, PartitioningMethod Method>
snapshot::RegistryState<EntityId> Registry<EntityId, Method>::captureState() const
{
    syncHost();
    snapshot::RegistryState<EntityId> state;
    const uint32_t shapeCount = static_cast<uint32_t>(mShapes.size());
    std::vector<uint8_t> type(shapeCount), vertexCount(shapeCount), flags(shapeCount);
    std::vector<uint64_t> category(shapeCount), mask(shapeCount);
    std::vector<float> geom(static_cast<size_t>(shapeCount) * C2D_GEOM_FLOATS), tight(static_cast<size_t>(shapeCount) * 4),
        fat(static_cast<size_t>(shapeCount) * 4);
    if (shapeCount)
    {
        check(c2d_gpu_read_shapes(mCtx, 0, shapeCount, type.data(), vertexCount.data(), flags.data(), category.data(), mask.data(),
                                  geom.data()));
        check(c2d_gpu_read_boxes(mCtx, 0, shapeCount, tight.data(), fat.data()));
    }
    state.shapes.resize(shapeCount);
    state.shapeEntity.resize(shapeCount);
    for (uint32_t s = 0; s < shapeCount; ++s)
    {
        snapshot::ShapeState& shape = state.shapes[s];
        shape.alive = mShapes[s].entity != kNone;
        if (!shape.alive)
            continue;
        const EntityInfo& owner = mEntities[mShapes[s].entity];
        state.shapeEntity[s] = owner.id;
        shape.type = type[s];
        shape.vertexCount = vertexCount[s];
        shape.body = static_cast<uint8_t>(owner.type);
        shape.active = (flags[s] & 2u) != 0;
        std::memcpy(shape.geom, geom.data() + static_cast<size_t>(s) * C2D_GEOM_FLOATS, sizeof(shape.geom));
        std::memcpy(shape.tight, tight.data() + static_cast<size_t>(s) * 4, sizeof(shape.tight));
        std::memcpy(shape.fat, fat.data() + static_cast<size_t>(s) * 4, sizeof(shape.fat));
        shape.category = category[s];
        shape.mask = mask[s];
    }
    state.entities.reserve(mLiveEntities);
    for (const EntityInfo& e : mEntities)
    {
        if (!e.alive)
            continue;
        snapshot::EntityState<EntityId> out;
        out.id = e.id;
        out.body = static_cast<uint8_t>(e.type);
        const auto store = [](const Transform& t, float* xf) {
            xf[0] = t.translation.x, xf[1] = t.translation.y;
            xf[2] = t.rotation.angleRadians, xf[3] = t.rotation.sin, xf[4] = t.rotation.cos;
        };
        store(e.transform, out.xf);
        out.hasPrev = e.type == BodyType::Bullet;
        store(e.previous, out.prev);
        for (uint32_t s = e.firstShape; s != kNone; s = mShapes[s].next)
            out.shapeIds.push_back(s);
        state.entities.push_back(std::move(out));
    }
    state.freeShapeIds = mFreeShapeIds;
    state.previousAllCollisionsCount = mPreviousAllCollisionsCount;
    state.gridCellSize = Method == PartitioningMethod::Grid ? static_cast<int32_t>(mCellSize) : 0;
    return state;
}

template <typename EntityId, PartitioningMethod Method>
void Registry<EntityId, Method>::serialize(std::ostream& out) const
{
    snapshot::write(out, captureState());
}

template <typename EntityId, PartitioningMethod Method>
Registry<EntityId, Method> Registry<EntityId, Method>::deserialize(std::istream& in)
{
    const snapshot::RegistryState<EntityId> state = snapshot::read<EntityId>(in, Method == PartitioningMethod::Grid);
    Registry registry = [&] {
        if constexpr (Method == PartitioningMethod::Grid)
            return Registry(static_cast<uint32_t>(state.gridCellSize));
        else
            return Registry();
    }();
    const auto load = [](const float* xf) {
        Transform t{};
        t.translation = Vec2{xf[0], xf[1]};
        t.rotation.angleRadians = xf[2];
        t.rotation.sin = xf[3];
        t.rotation.cos = xf[4];
        return t;
    };
    const uint32_t shapeCount = static_cast<uint32_t>(state.shapes.size());
    registry.mShapes.assign(shapeCount, ShapeLink{});
    registry.mFreeShapeIds = state.freeShapeIds;
    registry.mPreviousAllCollisionsCount = state.previousAllCollisionsCount;
    std::vector<uint32_t> inactive;
    for (const snapshot::EntityState<EntityId>& e : state.entities)
    {
        registry.createEntity(e.id, static_cast<BodyType>(e.body), load(e.xf));
        const uint32_t idx = registry.indexOf(e.id);
        EntityInfo& entity = registry.mEntities[idx];
        if (e.hasPrev)
            entity.previous = load(e.prev);
        const bool bullet = entity.type == BodyType::Bullet;
        const c2d_xf cur = pack(entity.transform, bullet);
        const c2d_xf prev = pack(entity.previous, bullet);
        for (const uint32_t shapeId : e.shapeIds)
        {
            const snapshot::ShapeState& shape = state.shapes[shapeId];
            registry.check(c2d_gpu_shape_add(registry.mCtx, shapeId, registry.tagOf(idx), shape.type, e.body, shape.vertexCount,
                                             shape.category, shape.mask, shape.geom, &cur, &prev));
            ShapeLink& link = registry.mShapes[shapeId];
            link.entity = idx;
            link.prev = entity.lastShape;
            link.next = kNone;
            if (entity.lastShape != kNone)
                registry.mShapes[entity.lastShape].next = shapeId;
            else
                entity.firstShape = shapeId;
            entity.lastShape = shapeId;
            if (!shape.active)
                inactive.push_back(shapeId);
        }
    }
    // the boxes are history (translations shift the tight box, the leaf box follows the fat-leaf rule): restore them
    if (shapeCount)
    {
        std::vector<float> tight(static_cast<size_t>(shapeCount) * 4), fat(static_cast<size_t>(shapeCount) * 4);
        for (uint32_t s = 0; s < shapeCount; ++s)
        {
            std::memcpy(tight.data() + static_cast<size_t>(s) * 4, state.shapes[s].tight, 4 * sizeof(float));
            std::memcpy(fat.data() + static_cast<size_t>(s) * 4, state.shapes[s].fat, 4 * sizeof(float));
        }
        registry.check(c2d_gpu_write_boxes(registry.mCtx, 0, shapeCount, tight.data(), fat.data()));
    }
    for (const uint32_t shapeId : inactive)
        registry.check(c2d_gpu_shape_set_active(registry.mCtx, shapeId, 0));
    return registry;
}

template <typename EntityId, PartitioningMethod Method>
bool Registry<EntityId, Method>::operator==(const Registry& other) const
{
    const snapshot::RegistryState<EntityId> a = captureState(), b = other.captureState();
    if (a.entities.size() != b.entities.size() || a.shapes.size() != b.shapes.size() || a.freeShapeIds != b.freeShapeIds)
        return false;
    for (uint32_t s = 0; s < a.shapes.size(); ++s)
    {
        const snapshot::ShapeState &x = a.shapes[s], &y = b.shapes[s];
        if (x.alive != y.alive)
            return false;
        if (!x.alive)
            continue;
        if (!(a.shapeEntity[s] == b.shapeEntity[s]) || x.type != y.type || x.vertexCount != y.vertexCount || x.body != y.body ||
            x.active != y.active || x.category != y.category || x.mask != y.mask ||
            std::memcmp(x.geom, y.geom, sizeof(x.geom)) != 0 || std::memcmp(x.tight, y.tight, sizeof(x.tight)) != 0 ||
            std::memcmp(x.fat, y.fat, sizeof(x.fat)) != 0)
            return false;
    }
    // entities may sit in different slots: match them through the other registry's index
    for (const snapshot::EntityState<EntityId>& e : a.entities)
    {
        const uint32_t idx = other.indexOf(e.id);
        if (idx == kNone)
            return false;
        const EntityInfo& o = other.mEntities[idx];
        const float oxf[5] = {o.transform.translation.x, o.transform.translation.y, o.transform.rotation.angleRadians,
                              o.transform.rotation.sin, o.transform.rotation.cos};
        const float opv[5] = {o.previous.translation.x, o.previous.translation.y, o.previous.rotation.angleRadians,
                              o.previous.rotation.sin, o.previous.rotation.cos};
        if (e.body != static_cast<uint8_t>(o.type) || std::memcmp(e.xf, oxf, sizeof(oxf)) != 0)
            return false;
        if (e.hasPrev && std::memcmp(e.prev, opv, sizeof(opv)) != 0)
            return false;
        uint32_t s = o.firstShape;
        for (const uint32_t shapeId : e.shapeIds)
        {
            if (s != shapeId)
                return false;
            s = other.mShapes[s].next;
        }
        if (s != kNone)
            return false;
    }
    return true;
}

} // namespace c2d
