// runner.cpp — drives c2d::Registry<uint32_t, Method> through its public API.
// TEST INFRASTRUCTURE (see runner_api.h).  Compiled unchanged against
//   (a) the reference headers  (-DRN_BACKEND_REF)   -> oracle/_ref/
//   (b) this repo's headers    (-DRN_BACKEND_B200)  -> colli2de_b200/
// Everything below the "free functions" marker is backend specific only in
// how a batch of (shape, transform) pairs is evaluated: the reference calls
// c2d::collide & co. directly, the B200 build forwards the batch to the
// C-ABI (c2d_gpu_*_batch), i.e. to the CUDA narrowphase.
#include "runner_api.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <malloc.h>
#include <memory>
#include <span>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#if defined(RN_BACKEND_REF)
// The secondary oracles of SURVEY.md section 8(c) item 5 (tight boxes, leaf boxes, broadphase candidates) live in
// private members of the reference's Registry (mShapes / mShapeTreeHandles / mTrees).  The reference sources stay
// untouched: its headers are read with `private` spelled `public` (access does not change layout or mangling), after
// every standard header they use has been included normally.
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <execution>
#include <format>
#include <functional>
#include <iostream>
#include <istream>
#include <limits>
#include <map>
#include <memory_resource>
#include <optional>
#include <ostream>
#include <print>
#include <ranges>
#include <set>
#include <thread>
#include <tsl/robin_map.h>
#include <utility>
#include <variant>
#define private public
#include <colli2de/Registry.hpp>
#undef private
#else
#include <colli2de/Registry.hpp>
#endif

#if defined(RN_BACKEND_B200)
#include <c2d_gpu.h>
#endif

using namespace c2d;

namespace
{

struct PairRec
{
    uint32_t entityA, entityB, shapeA, shapeB;
    float manifold[17]; // normal(2) + 2 x (point, anchorA, anchorB, separation) + pointCount byte
};
static_assert(sizeof(PairRec) == 84);

struct HitRec
{
    uint32_t entity, shape;
    float entry[2], exit[2], entryTime, exitTime;
};
static_assert(sizeof(HitRec) == 32);

Transform makeRawTransform(const float* xf5)
{
    Transform t{};
    t.translation = Vec2{xf5[0], xf5[1]};
    t.rotation.angleRadians = xf5[2];
    t.rotation.sin = xf5[3];
    t.rotation.cos = xf5[4];
    return t;
}

void storeRawTransform(const Transform& t, float* xf5)
{
    xf5[0] = t.translation.x;
    xf5[1] = t.translation.y;
    xf5[2] = t.rotation.angleRadians;
    xf5[3] = t.rotation.sin;
    xf5[4] = t.rotation.cos;
}

Circle makeCircle(const float* g)
{
    return Circle{Vec2{g[0], g[1]}, g[2]};
}
Capsule makeCapsule(const float* g)
{
    return Capsule{Vec2{g[0], g[1]}, Vec2{g[2], g[3]}, g[4]};
}
Segment makeSegment(const float* g)
{
    return Segment{Vec2{g[0], g[1]}, Vec2{g[2], g[3]}};
}
Polygon makePoly(const float* g, uint8_t count)
{
    std::array<Vec2, MAX_POLYGON_VERTICES> verts{};
    for (uint8_t i = 0; i < count; ++i)
        verts[i] = Vec2{g[2 * i], g[2 * i + 1]};
    return Polygon(verts, count);
}

ShapeVariant makeShape(uint8_t type, const float* g, uint8_t count)
{
    switch (type)
    {
    case 0:
        return makeCircle(g);
    case 1:
        return makeCapsule(g);
    case 2:
        return makeSegment(g);
    default:
        return makePoly(g, count);
    }
}

void storeManifold(const Manifold& m, void* out68)
{
    static_assert(sizeof(Manifold) == 68);
    std::memcpy(out68, &m, 68);
}

struct IRegistry
{
    virtual ~IRegistry() = default;
    virtual void createEntity(uint32_t id, BodyType type, const Transform& t) = 0;
    virtual void removeEntity(uint32_t id) = 0;
    virtual ShapeId addShape(uint32_t e, const ShapeVariant& s, uint64_t cat, uint64_t mask) = 0;
    virtual void removeShape(ShapeId s) = 0;
    virtual void setShapeActive(ShapeId s, bool a) = 0;
    virtual void teleportEntity(uint32_t id, const Transform& t) = 0;
    virtual void teleportEntity(uint32_t id, Translation t) = 0;
    virtual void moveEntity(uint32_t id, const Transform& d) = 0;
    virtual void moveEntity(uint32_t id, Translation d) = 0;
    virtual void moveEntitiesBulk(uint32_t n, const uint32_t* ids, const float* dxy) = 0;
    virtual void teleportEntitiesBulk(uint32_t n, const uint32_t* ids, const float* xy) = 0;
    virtual Transform getEntityTransform(uint32_t id) const = 0;
    virtual void getCollidingPairs(std::vector<PairRec>& out, double* seconds) = 0;
    virtual void getCollisions(uint32_t id, std::vector<PairRec>& out) const = 0;
    // all ids in one call where the backend has a batched entry point; false -> the caller loops getCollisions()
    virtual bool getCollisionsBatch(uint32_t n, const uint32_t* ids, std::vector<PairRec>& out) const = 0;
    virtual bool areColliding(uint32_t a, uint32_t b) const = 0;
    virtual bool firstHit(Ray r, uint64_t mask, HitRec& out) const = 0;
    virtual bool firstHit(InfiniteRay r, uint64_t mask, HitRec& out) const = 0;
    virtual uint64_t countPairsDevice(double* seconds) = 0;
    virtual uint64_t pairsView(double* seconds, const void** records) = 0;
    virtual void* native() const = 0;
    virtual void setGhost(ShapeId s, bool g) = 0;
    virtual void groupId(unsigned char* out128) = 0;
    virtual void groupJoin(uint32_t rank, uint32_t world, const unsigned char* id128) = 0;
    virtual void groupOptions(int results, int exchange) = 0;
    virtual void groupLeave() = 0;
    virtual void groupRaySharding(bool on) = 0;
    virtual void stats(double* out16) const = 0;
    virtual int stage(uint32_t n, const uint32_t* ids, const float* dxy) = 0;
    virtual void applyStaged(int handle) = 0;
    virtual void releaseStaged(int handle) = 0;
    virtual void rayCast(Ray r, uint64_t mask, std::vector<HitRec>& out) const = 0;
    virtual void rayCast(InfiniteRay r, uint64_t mask, std::vector<HitRec>& out) const = 0;
    // all rays in one call where the backend has a batched entry point; false -> the caller loops rayCast()
    // *apiSeconds = the time of the Registry call alone (packing the queries and unpacking the result are test plumbing)
    virtual bool rayCastBatch(uint32_t n, const float* ray4, const uint8_t* infinite, const uint64_t* masks,
                              uint64_t* offsets, std::vector<HitRec>& out, double* apiSeconds) const = 0;
    virtual size_t size() const = 0;
    virtual void clear() = 0;
    // secondary oracles: shape table size, per-shape tight / leaf boxes + alive flag, broadphase candidate pairs
    virtual uint32_t shapeSlots() = 0;
    virtual void readBoxes(uint32_t first, uint32_t count, float* tight4, float* fat4, uint8_t* alive) = 0;
    virtual void candidates(std::vector<std::pair<uint32_t, uint32_t>>& out) = 0;
    virtual void serialize(std::string& out) const = 0;
    virtual bool equals(const IRegistry& other) const = 0;
};

template <PartitioningMethod Method>
struct RegistryHolder final : IRegistry
{
    using Reg = Registry<uint32_t, Method>;
    Reg reg;

    RegistryHolder() requires(Method == PartitioningMethod::None) = default;
    explicit RegistryHolder(uint32_t cell) requires(Method == PartitioningMethod::Grid) : reg(cell) {}
    explicit RegistryHolder(Reg&& loaded) : reg(std::move(loaded)) {}

    void serialize(std::string& out) const override
    {
        std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
        reg.serialize(ss);
        out = ss.str();
    }
    bool equals(const IRegistry& other) const override
    {
        const auto* o = dynamic_cast<const RegistryHolder*>(&other);
        return o != nullptr && reg == o->reg;
    }

    void createEntity(uint32_t id, BodyType type, const Transform& t) override
    {
        reg.createEntity(id, type, t);
    }
    void removeEntity(uint32_t id) override
    {
        reg.removeEntity(id);
    }
    ShapeId addShape(uint32_t e, const ShapeVariant& s, uint64_t cat, uint64_t mask) override
    {
        return std::visit([&](const auto& concrete) { return reg.addShape(e, concrete, cat, mask); }, s);
    }
    void removeShape(ShapeId s) override
    {
        reg.removeShape(s);
    }
    void setShapeActive(ShapeId s, bool a) override
    {
        reg.setShapeActive(s, a);
    }
    void teleportEntity(uint32_t id, const Transform& t) override
    {
        reg.teleportEntity(id, t);
    }
    void teleportEntity(uint32_t id, Translation t) override
    {
        reg.teleportEntity(id, t);
    }
    void moveEntity(uint32_t id, const Transform& d) override
    {
        reg.moveEntity(id, d);
    }
    void moveEntity(uint32_t id, Translation d) override
    {
        reg.moveEntity(id, d);
    }
    Transform getEntityTransform(uint32_t id) const override
    {
        return reg.getEntityTransform(id);
    }
    void getCollidingPairs(std::vector<PairRec>& out, double* seconds) override
    {
        const auto t0 = std::chrono::steady_clock::now();
        auto pairs = reg.getCollidingPairs();
        const auto t1 = std::chrono::steady_clock::now();
        if (seconds)
            *seconds = std::chrono::duration<double>(t1 - t0).count();
        static_assert(sizeof(typename Reg::EntityCollision) == sizeof(PairRec));
        out.resize(pairs.size());
        if (!pairs.empty())
            std::memcpy(out.data(), pairs.data(), pairs.size() * sizeof(PairRec));
    }
    bool getCollisionsBatch(uint32_t n, const uint32_t* ids, std::vector<PairRec>& out) const override
    {
#if defined(RN_BACKEND_B200)
        const auto pairs = reg.getCollisions(std::span<const uint32_t>(ids, n));
        out.resize(pairs.size());
        if (!pairs.empty())
            std::memcpy(out.data(), pairs.data(), pairs.size() * sizeof(PairRec));
        return true;
#else
        (void)n, (void)ids, (void)out;
        return false;
#endif
    }
    void getCollisions(uint32_t id, std::vector<PairRec>& out) const override
    {
        const auto pairs = reg.getCollisions(id);
        out.resize(pairs.size());
        if (!pairs.empty())
            std::memcpy(out.data(), pairs.data(), pairs.size() * sizeof(PairRec));
    }
    bool areColliding(uint32_t a, uint32_t b) const override
    {
        return reg.areColliding(a, b);
    }
    template <typename RayT>
    bool firstHitImpl(RayT r, uint64_t mask, HitRec& out) const
    {
        const auto h = reg.firstHitRayCast(r, mask);
        if (!h)
            return false;
        out = HitRec{h->id.first, h->id.second, {h->entry.x, h->entry.y}, {h->exit.x, h->exit.y}, h->entryTime, h->exitTime};
        return true;
    }
    bool firstHit(Ray r, uint64_t mask, HitRec& out) const override
    {
        return firstHitImpl(r, mask, out);
    }
    bool firstHit(InfiniteRay r, uint64_t mask, HitRec& out) const override
    {
        return firstHitImpl(r, mask, out);
    }
#if defined(RN_BACKEND_B200)
    void moveEntitiesBulk(uint32_t n, const uint32_t* ids, const float* dxy) override
    {
        static_assert(sizeof(Translation) == 2 * sizeof(float));
        reg.moveEntities(std::span<const uint32_t>(ids, n), std::span<const Translation>(reinterpret_cast<const Translation*>(dxy), n));
    }
    void teleportEntitiesBulk(uint32_t n, const uint32_t* ids, const float* xy) override
    {
        reg.teleportEntities(std::span<const uint32_t>(ids, n), std::span<const Translation>(reinterpret_cast<const Translation*>(xy), n));
    }
    void setGhost(ShapeId s, bool g) override
    {
        reg.setShapeGhost(s, g);
    }
    void groupId(unsigned char* out128) override
    {
        const auto id = Reg::createGroupId();
        std::memcpy(out128, id.data(), id.size());
    }
    void groupJoin(uint32_t rank, uint32_t world, const unsigned char* id128) override
    {
        typename Reg::GroupId id;
        std::memcpy(id.data(), id128, id.size());
        reg.joinGroup(rank, world, id);
    }
    void groupOptions(int results, int exchange) override
    {
        reg.setGroupResults(static_cast<typename Reg::GroupResults>(results));
        reg.setGroupMoveExchange(exchange != 0);
    }
    void groupLeave() override
    {
        reg.leaveGroup();
    }
    void groupRaySharding(bool on) override
    {
        reg.setGroupRaySharding(on);
    }
    uint64_t pairsView(double* seconds, const void** records) override
    {
        const auto t0 = std::chrono::steady_clock::now();
        const auto view = reg.getCollidingPairsView();
        if (seconds)
            *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *records = view.data();
        return view.size();
    }
    bool rayCastBatch(uint32_t n, const float* ray4, const uint8_t* infinite, const uint64_t* masks, uint64_t* offsets,
                      std::vector<HitRec>& out, double* apiSeconds) const override
    {
        std::vector<typename Reg::RayQuery> q(n);
        for (uint32_t i = 0; i < n; ++i)
            q[i] = typename Reg::RayQuery{Vec2{ray4[4 * i], ray4[4 * i + 1]}, Vec2{ray4[4 * i + 2], ray4[4 * i + 3]},
                                          masks ? masks[i] : ~0ull, infinite && infinite[i]};
        const auto t0 = std::chrono::steady_clock::now();
        const auto res = reg.rayCastBatch(std::span<const typename Reg::RayQuery>(q));
        if (apiSeconds)
            *apiSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::copy(res.offsets.begin(), res.offsets.end(), offsets);
        static_assert(sizeof(typename Reg::RayHit) == sizeof(HitRec));
        out.resize(res.hits.size());
        if (!res.hits.empty())
            std::memcpy(static_cast<void*>(out.data()), static_cast<const void*>(res.hits.data()), res.hits.size() * sizeof(HitRec));
        return true;
    }
    void* native() const override
    {
        return reg.nativeContext();
    }
    uint64_t countPairsDevice(double* seconds) override
    {
        const auto t0 = std::chrono::steady_clock::now();
        const uint64_t n = reg.countCollidingPairsOnDevice();
        if (seconds)
            *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return n;
    }
    void stats(double* o) const override
    {
        const c2d_gpu_stats& s = reg.lastQueryStats();
        const double v[16] = {double(s.shapes_alive), double(s.candidates), double(s.pairs), double(s.h2d_bytes),
                              double(s.d2h_bytes), double(s.kernel_launches), double(s.retries), s.ms_refit,
                              s.ms_build, s.ms_broad, s.ms_narrow, s.ms_device, s.ms_total, s.ms_host_enqueue, s.ms_host_wait, 0};
        std::memcpy(o, v, sizeof(v));
    }
    int stage(uint32_t n, const uint32_t* ids, const float* dxy) override
    {
        std::vector<Translation> d(n);
        for (uint32_t i = 0; i < n; ++i)
            d[i] = Translation{dxy[2 * i], dxy[2 * i + 1]};
        return static_cast<int>(reg.stageMoves(std::span<const uint32_t>(ids, n), d));
    }
    void applyStaged(int handle) override
    {
        reg.applyStagedMoves(static_cast<uint32_t>(handle));
    }
    void releaseStaged(int handle) override
    {
        reg.releaseStagedMoves(static_cast<uint32_t>(handle));
    }
    static void gpuCheck(c2d_gpu_ctx* g, int rc)
    {
        if (rc != 0)
            throw std::runtime_error(c2d_gpu_last_error(g));
    }
    uint32_t shapeSlots() override
    {
        uint32_t n = 0;
        gpuCheck(reg.nativeContext(), c2d_gpu_shape_slots(reg.nativeContext(), &n));
        return n;
    }
    void readBoxes(uint32_t first, uint32_t count, float* tight4, float* fat4, uint8_t* alive) override
    {
        c2d_gpu_ctx* g = reg.nativeContext();
        gpuCheck(g, c2d_gpu_read_boxes(g, first, count, tight4, fat4));
        if (alive)
        {
            gpuCheck(g, c2d_gpu_read_shapes(g, first, count, nullptr, nullptr, alive, nullptr, nullptr, nullptr));
            for (uint32_t i = 0; i < count; ++i)
                alive[i] &= 1u;
        }
    }
    void candidates(std::vector<std::pair<uint32_t, uint32_t>>& out) override
    {
        c2d_gpu_ctx* g = reg.nativeContext();
        reg.countCollidingPairsOnDevice();
        uint64_t n = 0;
        gpuCheck(g, c2d_gpu_read_candidates(g, 0, nullptr, &n));
        out.resize(n);
        static_assert(sizeof(std::pair<uint32_t, uint32_t>) == 8);
        if (n)
            gpuCheck(g, c2d_gpu_read_candidates(g, n, reinterpret_cast<uint32_t*>(out.data()), &n));
    }
#else
    uint32_t shapeSlots() override
    {
        return static_cast<uint32_t>(reg.mShapes.size());
    }
    void readBoxes(uint32_t first, uint32_t count, float* tight4, float* fat4, uint8_t* alive) override
    {
        if constexpr (Method == PartitioningMethod::None)
        {
            std::vector<uint8_t> isFree(reg.mShapes.size(), 0);
            for (const ShapeId f : reg.mFreeShapeIds)
                isFree[f] = 1;
            for (uint32_t i = 0; i < count; ++i)
            {
                const uint32_t s = first + i;
                if (s >= reg.mShapes.size())
                    throw std::runtime_error("rn_read_boxes: shape id out of range");
                const AABB tight = reg.mShapes[s].mBoundingBox; // ShapeInstance::mBoundingBox (Registry.hpp:117)
                AABB fat = tight;
                if (!isFree[s])
                {
                    // the leaf of the shape's proxy (DynamicBVH.hpp:119-122, 352-365, 386-417)
                    const BodyType body = reg.mEntities.at(reg.mShapeEntity[s]).mType;
                    fat = reg.treeFor(body).getNode(static_cast<NodeIndex>(reg.mShapeTreeHandles[s])).mBoundingBox;
                }
                if (tight4)
                    tight4[4 * i] = tight.min.x, tight4[4 * i + 1] = tight.min.y, tight4[4 * i + 2] = tight.max.x,
                              tight4[4 * i + 3] = tight.max.y;
                if (fat4)
                    fat4[4 * i] = fat.min.x, fat4[4 * i + 1] = fat.min.y, fat4[4 * i + 2] = fat.max.x, fat4[4 * i + 3] = fat.max.y;
                if (alive)
                    alive[i] = isFree[s] ? 0 : 1;
            }
        }
        else
            throw std::runtime_error("rn_read_boxes: the reference's Grid registry keeps one leaf per touched cell");
    }
    void candidates(std::vector<std::pair<uint32_t, uint32_t>>& out) override
    {
        out.clear();
        if constexpr (Method == PartitioningMethod::None)
        {
            // the five broadphase groups of getCollidingPairs (Registry.hpp:487-545), each sorted like there
            const auto cross = [&](BodyType a, BodyType b)
            {
                const size_t first = out.size();
                reg.treeFor(a).findAllCollisions(reg.treeFor(b), [&](ShapeId x, ShapeId y) { out.emplace_back(x, y); });
                std::sort(out.begin() + first, out.end());
            };
            const auto self = [&](BodyType a)
            {
                const size_t first = out.size();
                reg.treeFor(a).findAllCollisions([&](ShapeId x, ShapeId y) { out.emplace_back(std::min(x, y), std::max(x, y)); });
                std::sort(out.begin() + first, out.end());
            };
            cross(BodyType::Dynamic, BodyType::Static);
            self(BodyType::Dynamic);
            cross(BodyType::Bullet, BodyType::Static);
            cross(BodyType::Bullet, BodyType::Dynamic);
            self(BodyType::Bullet);
        }
        else
            throw std::runtime_error("rn_candidates: PartitioningMethod::None only");
    }
    std::vector<std::pair<std::vector<uint32_t>, std::vector<Translation>>> staged;
    std::vector<typename Reg::EntityCollision> viewStorage;
    void moveEntitiesBulk(uint32_t n, const uint32_t* ids, const float* dxy) override
    {
        for (uint32_t i = 0; i < n; ++i)
            reg.moveEntity(ids[i], Translation{dxy[2 * i], dxy[2 * i + 1]});
    }
    void teleportEntitiesBulk(uint32_t n, const uint32_t* ids, const float* xy) override
    {
        for (uint32_t i = 0; i < n; ++i)
            reg.teleportEntity(ids[i], Translation{xy[2 * i], xy[2 * i + 1]});
    }
    void setGhost(ShapeId, bool) override
    {
        throw std::runtime_error("ghost shapes are an extension of the B200 build");
    }
    void groupId(unsigned char*) override
    {
        throw std::runtime_error("multi-GPU groups are an extension of the B200 build");
    }
    void groupJoin(uint32_t, uint32_t, const unsigned char*) override
    {
        throw std::runtime_error("multi-GPU groups are an extension of the B200 build");
    }
    void groupOptions(int, int) override
    {
        throw std::runtime_error("multi-GPU groups are an extension of the B200 build");
    }
    void groupLeave() override
    {
        throw std::runtime_error("multi-GPU groups are an extension of the B200 build");
    }
    void groupRaySharding(bool) override
    {
        throw std::runtime_error("multi-GPU groups are an extension of the B200 build");
    }
    uint64_t pairsView(double* seconds, const void** records) override
    {
        const auto t0 = std::chrono::steady_clock::now();
        viewStorage = reg.getCollidingPairs();
        if (seconds)
            *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *records = viewStorage.data();
        return viewStorage.size();
    }
    bool rayCastBatch(uint32_t, const float*, const uint8_t*, const uint64_t*, uint64_t*, std::vector<HitRec>&, double*) const override
    {
        return false;
    }
    void* native() const override
    {
        return nullptr;
    }
    uint64_t countPairsDevice(double* seconds) override
    {
        const auto t0 = std::chrono::steady_clock::now();
        const uint64_t n = reg.getCollidingPairs().size();
        if (seconds)
            *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return n;
    }
    void stats(double* o) const override
    {
        std::memset(o, 0, 16 * sizeof(double));
    }
    int stage(uint32_t n, const uint32_t* ids, const float* dxy) override
    {
        std::vector<Translation> d(n);
        for (uint32_t i = 0; i < n; ++i)
            d[i] = Translation{dxy[2 * i], dxy[2 * i + 1]};
        staged.emplace_back(std::vector<uint32_t>(ids, ids + n), std::move(d));
        return static_cast<int>(staged.size() - 1);
    }
    void applyStaged(int handle) override
    {
        const auto& b = staged.at(handle);
        for (size_t i = 0; i < b.first.size(); ++i)
            reg.moveEntity(b.first[i], b.second[i]);
    }
    void releaseStaged(int handle) override
    {
        staged.at(handle) = {};
    }
#endif
    template <typename RayT>
    void rayCastImpl(RayT r, uint64_t mask, std::vector<HitRec>& out) const
    {
        const auto hits = reg.rayCast(r, mask);
        for (const auto& h : hits)
            out.push_back(HitRec{
                h.id.first, h.id.second, {h.entry.x, h.entry.y}, {h.exit.x, h.exit.y}, h.entryTime, h.exitTime});
    }
    void rayCast(Ray r, uint64_t mask, std::vector<HitRec>& out) const override
    {
        rayCastImpl(r, mask, out);
    }
    void rayCast(InfiniteRay r, uint64_t mask, std::vector<HitRec>& out) const override
    {
        rayCastImpl(r, mask, out);
    }
    size_t size() const override
    {
        return reg.size();
    }
    void clear() override
    {
        reg.clear();
    }
};

} // namespace

struct rn_ctx
{
    std::unique_ptr<IRegistry> reg;
    std::unordered_map<uint32_t, uint8_t> body;
    std::unordered_map<uint32_t, Transform> prev; // bullets only (mirrors Registry.hpp:167 by observation)
    std::vector<PairRec> pairs;
    std::vector<HitRec> hits;
    std::vector<std::pair<uint32_t, uint32_t>> cands; // last rn_candidates
    std::string blob; // last rn_serialize
    std::string error;

    void notePrev(uint32_t id)
    {
        auto it = body.find(id);
        if (it != body.end() && it->second == 2)
            prev[id] = reg->getEntityTransform(id);
    }
};

#define RN_TRY(ctx) (ctx)->error.clear(); try {
#define RN_CATCH(ctx)                                                                                                  \
    }                                                                                                                  \
    catch (const std::exception& e)                                                                                    \
    {                                                                                                                  \
        (ctx)->error = e.what();                                                                                       \
    }

extern "C" {

const char* rn_backend_name(void)
{
#if defined(RN_BACKEND_REF)
    return "reference";
#elif defined(RN_BACKEND_B200)
    return "b200";
#else
    return "unknown";
#endif
}

rn_ctx* rn_create(int method, uint32_t cell_size)
{
    // Application-level allocator policy, applied to BOTH backends: getCollidingPairs() returns a fresh
    // std::vector every frame (tens of MB at 1 M shapes); keeping freed blocks in the heap instead of
    // mmap/munmap-ing them avoids a page-fault storm per frame.
    static const bool mallocTuned = []
    {
        mallopt(M_MMAP_THRESHOLD, 1 << 30);
        mallopt(M_TRIM_THRESHOLD, 1 << 30);
        return true;
    }();
    (void)mallocTuned;
    auto* ctx = new rn_ctx();
    try
    {
        if (method == 1)
            ctx->reg = std::make_unique<RegistryHolder<PartitioningMethod::Grid>>(cell_size ? cell_size : 120u);
        else
            ctx->reg = std::make_unique<RegistryHolder<PartitioningMethod::None>>();
    }
    catch (const std::exception& e)
    {
        ctx->error = e.what();
    }
    return ctx;
}

void rn_destroy(rn_ctx* ctx)
{
    delete ctx;
}

const char* rn_last_error(rn_ctx* ctx)
{
    return ctx->error.c_str();
}

void rn_create_entities(rn_ctx* ctx, uint32_t n, const uint32_t* ids, const uint8_t* body, const float* xf3)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        const Transform t(Vec2{xf3[3 * i], xf3[3 * i + 1]}, xf3[3 * i + 2]);
        ctx->reg->createEntity(ids[i], static_cast<BodyType>(body[i]), t);
        ctx->body[ids[i]] = body[i];
        if (body[i] == 2)
            ctx->prev[ids[i]] = t;
    }
    RN_CATCH(ctx)
}

void rn_create_entities_raw(rn_ctx* ctx, uint32_t n, const uint32_t* ids, const uint8_t* body, const float* xf5)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        const Transform t = makeRawTransform(xf5 + 5 * i);
        ctx->reg->createEntity(ids[i], static_cast<BodyType>(body[i]), t);
        ctx->body[ids[i]] = body[i];
        if (body[i] == 2)
            ctx->prev[ids[i]] = t;
    }
    RN_CATCH(ctx)
}

void rn_add_shapes(rn_ctx* ctx, uint32_t n, const uint32_t* entity, const uint8_t* type, const float* geom16,
                   const uint8_t* vcount, const uint64_t* cat, const uint64_t* mask, uint32_t* out_ids)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        const ShapeVariant s = makeShape(type[i], geom16 + 16 * i, vcount[i]);
        const ShapeId id = ctx->reg->addShape(entity[i], s, cat ? cat[i] : 1ull, mask ? mask[i] : ~0ull);
        if (out_ids)
            out_ids[i] = id;
    }
    RN_CATCH(ctx)
}

void rn_remove_shapes(rn_ctx* ctx, uint32_t n, const uint32_t* shape_ids)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
        ctx->reg->removeShape(shape_ids[i]);
    RN_CATCH(ctx)
}

void rn_remove_entities(rn_ctx* ctx, uint32_t n, const uint32_t* ids)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        ctx->reg->removeEntity(ids[i]);
        ctx->body.erase(ids[i]);
        ctx->prev.erase(ids[i]);
    }
    RN_CATCH(ctx)
}

void rn_set_active(rn_ctx* ctx, uint32_t n, const uint32_t* shape_ids, const uint8_t* active)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
        ctx->reg->setShapeActive(shape_ids[i], active[i] != 0);
    RN_CATCH(ctx)
}

void rn_move_translate(rn_ctx* ctx, uint32_t n, const uint32_t* ids, const float* dxy)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        ctx->notePrev(ids[i]);
        ctx->reg->moveEntity(ids[i], Translation{dxy[2 * i], dxy[2 * i + 1]});
    }
    RN_CATCH(ctx)
}

// Registry::moveEntities / teleportEntities (B200 extension; the reference loops over moveEntity / teleportEntity).  The
// runner's own bookkeeping of bullets' previous transforms is skipped: seconds = the library's cost of the batch alone.
void rn_move_entities_bulk(rn_ctx* ctx, uint32_t n, const uint32_t* ids, const float* dxy, int teleport, double* seconds)
{
    RN_TRY(ctx)
    const auto t0 = std::chrono::steady_clock::now();
    if (teleport)
        ctx->reg->teleportEntitiesBulk(n, ids, dxy);
    else
        ctx->reg->moveEntitiesBulk(n, ids, dxy);
    if (seconds)
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    RN_CATCH(ctx)
}

void rn_move_transform(rn_ctx* ctx, uint32_t n, const uint32_t* ids, const float* d)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        ctx->notePrev(ids[i]);
        ctx->reg->moveEntity(ids[i], Transform(Vec2{d[3 * i], d[3 * i + 1]}, d[3 * i + 2]));
    }
    RN_CATCH(ctx)
}

void rn_teleport_transform(rn_ctx* ctx, uint32_t n, const uint32_t* ids, const float* xf3)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        ctx->notePrev(ids[i]);
        ctx->reg->teleportEntity(ids[i], Transform(Vec2{xf3[3 * i], xf3[3 * i + 1]}, xf3[3 * i + 2]));
    }
    RN_CATCH(ctx)
}

void rn_teleport_translation(rn_ctx* ctx, uint32_t n, const uint32_t* ids, const float* xy)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        ctx->notePrev(ids[i]);
        ctx->reg->teleportEntity(ids[i], Translation{xy[2 * i], xy[2 * i + 1]});
    }
    RN_CATCH(ctx)
}

void rn_get_transforms(rn_ctx* ctx, uint32_t n, const uint32_t* ids, float* out)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
        storeRawTransform(ctx->reg->getEntityTransform(ids[i]), out + 5 * i);
    RN_CATCH(ctx)
}

void rn_get_prev_transforms(rn_ctx* ctx, uint32_t n, const uint32_t* ids, float* out)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        auto it = ctx->prev.find(ids[i]);
        if (it != ctx->prev.end())
            storeRawTransform(it->second, out + 5 * i);
        else
            storeRawTransform(ctx->reg->getEntityTransform(ids[i]), out + 5 * i);
    }
    RN_CATCH(ctx)
}

uint64_t rn_size(rn_ctx* ctx)
{
    return ctx->reg->size();
}

void rn_clear(rn_ctx* ctx)
{
    RN_TRY(ctx)
    ctx->reg->clear();
    ctx->body.clear();
    ctx->prev.clear();
    RN_CATCH(ctx)
}

uint64_t rn_get_colliding_pairs(rn_ctx* ctx, double* seconds)
{
    RN_TRY(ctx)
    ctx->reg->getCollidingPairs(ctx->pairs, seconds);
    return ctx->pairs.size();
    RN_CATCH(ctx)
    return 0;
}

uint64_t rn_count_colliding_pairs_device(rn_ctx* ctx, double* seconds)
{
    RN_TRY(ctx)
    return ctx->reg->countPairsDevice(seconds);
    RN_CATCH(ctx)
    return 0;
}

uint64_t rn_get_collisions_many(rn_ctx* ctx, uint32_t n, const uint32_t* ids, double* seconds)
{
    RN_TRY(ctx)
    uint64_t total = 0;
    const auto t0 = std::chrono::steady_clock::now();
    if (std::getenv("RN_COLLISIONS_BATCH") != nullptr && ctx->reg->getCollisionsBatch(n, ids, ctx->pairs))
        total = ctx->pairs.size(); // the batched extension; the records stay available to rn_copy_pairs
    else
        for (uint32_t i = 0; i < n; ++i)
        {
            ctx->reg->getCollisions(ids[i], ctx->pairs);
            total += ctx->pairs.size();
        }
    if (seconds)
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return total;
    RN_CATCH(ctx)
    return 0;
}

uint64_t rn_get_collisions(rn_ctx* ctx, uint32_t id)
{
    RN_TRY(ctx)
    ctx->reg->getCollisions(id, ctx->pairs);
    return ctx->pairs.size();
    RN_CATCH(ctx)
    return 0;
}

void rn_are_colliding(rn_ctx* ctx, uint32_t n, const uint32_t* idsA, const uint32_t* idsB, uint8_t* out)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
        out[i] = ctx->reg->areColliding(idsA[i], idsB[i]) ? 1 : 0;
    RN_CATCH(ctx)
}

void rn_first_hit_raycast(rn_ctx* ctx, uint32_t n, const float* ray4, const uint8_t* infinite, const uint64_t* masks,
                          uint8_t* out_found, void* out_hits32)
{
    RN_TRY(ctx)
    auto* hits = static_cast<HitRec*>(out_hits32);
    for (uint32_t i = 0; i < n; ++i)
    {
        const float* r = ray4 + 4 * i;
        const uint64_t mask = masks ? masks[i] : ~0ull;
        hits[i] = HitRec{};
        if (infinite && infinite[i])
            out_found[i] = ctx->reg->firstHit(InfiniteRay{Vec2{r[0], r[1]}, Vec2{r[2], r[3]}}, mask, hits[i]) ? 1 : 0;
        else
            out_found[i] = ctx->reg->firstHit(Ray{Vec2{r[0], r[1]}, Vec2{r[2], r[3]}}, mask, hits[i]) ? 1 : 0;
    }
    RN_CATCH(ctx)
}

uint64_t rn_get_colliding_pairs_view(rn_ctx* ctx, double* seconds, const void** records)
{
    RN_TRY(ctx)
    return ctx->reg->pairsView(seconds, records);
    RN_CATCH(ctx)
    return 0;
}

uint64_t rn_serialize(rn_ctx* ctx)
{
    RN_TRY(ctx)
    ctx->reg->serialize(ctx->blob);
    return ctx->blob.size();
    RN_CATCH(ctx)
    return 0;
}

void rn_copy_blob(rn_ctx* ctx, void* out)
{
    if (!ctx->blob.empty())
        std::memcpy(out, ctx->blob.data(), ctx->blob.size());
}

rn_ctx* rn_deserialize(int method, const void* bytes, uint64_t size)
{
    auto* ctx = new rn_ctx();
    try
    {
        std::stringstream ss(std::string(static_cast<const char*>(bytes), size), std::ios::in | std::ios::out | std::ios::binary);
        if (method == 1)
            ctx->reg = std::make_unique<RegistryHolder<PartitioningMethod::Grid>>(
                Registry<uint32_t, PartitioningMethod::Grid>::deserialize(ss));
        else
            ctx->reg = std::make_unique<RegistryHolder<PartitioningMethod::None>>(
                Registry<uint32_t, PartitioningMethod::None>::deserialize(ss));
    }
    catch (const std::exception& e)
    {
        ctx->error = e.what();
    }
    return ctx;
}

/* the harness mirrors body types and bullets' previous transforms by observation (rn_ctx::body / prev): after
 * rn_deserialize they are re-seeded from the caller's knowledge of the scene */
void rn_seed_entities(rn_ctx* ctx, uint32_t n, const uint32_t* ids, const uint8_t* body, const float* prev_xf5)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
    {
        ctx->body[ids[i]] = body[i];
        if (body[i] == 2)
            ctx->prev[ids[i]] = makeRawTransform(prev_xf5 + 5 * i);
    }
    RN_CATCH(ctx)
}

int rn_equals(rn_ctx* a, rn_ctx* b)
{
    RN_TRY(a)
    return a->reg->equals(*b->reg) ? 1 : 0;
    RN_CATCH(a)
    return -1;
}

void rn_set_ghost(rn_ctx* ctx, uint32_t n, const uint32_t* shape_ids, const uint8_t* ghost)
{
    RN_TRY(ctx)
    for (uint32_t i = 0; i < n; ++i)
        ctx->reg->setGhost(shape_ids[i], ghost[i] != 0);
    RN_CATCH(ctx)
}

void rn_group_id(rn_ctx* ctx, void* out128)
{
    RN_TRY(ctx)
    ctx->reg->groupId(static_cast<unsigned char*>(out128));
    RN_CATCH(ctx)
}

void rn_group_join(rn_ctx* ctx, uint32_t rank, uint32_t world, const void* id128)
{
    RN_TRY(ctx)
    ctx->reg->groupJoin(rank, world, static_cast<const unsigned char*>(id128));
    RN_CATCH(ctx)
}

void rn_group_options(rn_ctx* ctx, int results, int exchange)
{
    RN_TRY(ctx)
    ctx->reg->groupOptions(results, exchange);
    RN_CATCH(ctx)
}

void rn_group_ray_sharding(rn_ctx* ctx, int on)
{
    RN_TRY(ctx)
    ctx->reg->groupRaySharding(on != 0);
    RN_CATCH(ctx)
}

void rn_group_leave(rn_ctx* ctx)
{
    RN_TRY(ctx)
    ctx->reg->groupLeave();
    RN_CATCH(ctx)
}

void* rn_native_context(rn_ctx* ctx)
{
    return ctx->reg->native();
}

uint32_t rn_shape_slots(rn_ctx* ctx)
{
    RN_TRY(ctx)
    return ctx->reg->shapeSlots();
    RN_CATCH(ctx)
    return 0;
}

void rn_read_boxes(rn_ctx* ctx, uint32_t first, uint32_t count, float* out_tight4, float* out_fat4, uint8_t* out_alive)
{
    RN_TRY(ctx)
    ctx->reg->readBoxes(first, count, out_tight4, out_fat4, out_alive);
    RN_CATCH(ctx)
}

uint64_t rn_candidates(rn_ctx* ctx)
{
    RN_TRY(ctx)
    ctx->reg->candidates(ctx->cands);
    return ctx->cands.size();
    RN_CATCH(ctx)
    return 0;
}

void rn_copy_candidates(rn_ctx* ctx, uint32_t* out_pairs2)
{
    if (!ctx->cands.empty())
        std::memcpy(out_pairs2, ctx->cands.data(), ctx->cands.size() * sizeof(ctx->cands[0]));
}

void rn_query_stats(rn_ctx* ctx, double* out16)
{
    ctx->reg->stats(out16);
}

int rn_stage_translations(rn_ctx* ctx, uint32_t n, const uint32_t* ids, const float* dxy)
{
    RN_TRY(ctx)
    return ctx->reg->stage(n, ids, dxy);
    RN_CATCH(ctx)
    return -1;
}

void rn_apply_staged(rn_ctx* ctx, int handle)
{
    RN_TRY(ctx)
    // previous transforms of bullets are tracked by observation (see notePrev): not needed for staged batches
    // in the tests that use them (no bullets are staged there)
    ctx->reg->applyStaged(handle);
    RN_CATCH(ctx)
}

// K frames of "apply a staged move batch, count the colliding pairs on the device" without leaving native code (the
// device-resident arm of bench.py: a Python call per frame costs tens of microseconds, a visible share of a 0.5-1 ms
// frame).  sums16 accumulates rn_query_stats over the frames; returns the summed pair counts.
uint64_t rn_run_staged_frames(rn_ctx* ctx, uint32_t frames, const int* handles, double* sums16, double* seconds)
{
    RN_TRY(ctx)
    uint64_t pairs = 0;
    double one[16];
    for (int k = 0; k < 16; ++k)
        sums16[k] = 0.0;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t f = 0; f < frames; ++f)
    {
        ctx->reg->applyStaged(handles[f]);
        pairs += ctx->reg->countPairsDevice(nullptr);
        ctx->reg->stats(one);
        for (int k = 0; k < 16; ++k)
            sums16[k] += one[k];
    }
    if (seconds)
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return pairs;
    RN_CATCH(ctx)
    return 0;
}

void rn_release_staged(rn_ctx* ctx, int handle)
{
    RN_TRY(ctx)
    ctx->reg->releaseStaged(handle);
    RN_CATCH(ctx)
}

void rn_copy_pairs(rn_ctx* ctx, void* out84)
{
    if (!ctx->pairs.empty())
        std::memcpy(out84, ctx->pairs.data(), ctx->pairs.size() * sizeof(PairRec));
}

uint64_t rn_raycast(rn_ctx* ctx, uint32_t n, const float* ray4, const uint8_t* infinite, const uint64_t* masks,
                    uint64_t* offsets, double* seconds)
{
    RN_TRY(ctx)
    ctx->hits.clear();
    const auto t0 = std::chrono::steady_clock::now();
    const bool single = std::getenv("RN_RAYCAST_SINGLE") != nullptr; // force the one-ray-per-call API
    if (!single && ctx->reg->rayCastBatch(n, ray4, infinite, masks, offsets, ctx->hits, seconds))
        return ctx->hits.size();
    for (uint32_t i = 0; i < n; ++i)
    {
        offsets[i] = ctx->hits.size();
        const float* r = ray4 + 4 * i;
        const uint64_t mask = masks ? masks[i] : ~0ull;
        if (infinite && infinite[i])
            ctx->reg->rayCast(InfiniteRay{Vec2{r[0], r[1]}, Vec2{r[2], r[3]}}, mask, ctx->hits);
        else
            ctx->reg->rayCast(Ray{Vec2{r[0], r[1]}, Vec2{r[2], r[3]}}, mask, ctx->hits);
    }
    offsets[n] = ctx->hits.size();
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds)
        *seconds = std::chrono::duration<double>(t1 - t0).count();
    return ctx->hits.size();
    RN_CATCH(ctx)
    return 0;
}

void rn_copy_hits(rn_ctx* ctx, void* out32)
{
    if (!ctx->hits.empty())
        std::memcpy(out32, ctx->hits.data(), ctx->hits.size() * sizeof(HitRec));
}

// ---------------------------------------------------------------------------------------------------
// free functions
// ---------------------------------------------------------------------------------------------------
#if defined(RN_BACKEND_REF)

void rn_collide(uint32_t n, const uint8_t* typeA, const float* geomA, const uint8_t* cntA, const float* xfA5,
                const uint8_t* typeB, const float* geomB, const uint8_t* cntB, const float* xfB5,
                void* out_manifold68, uint8_t* out_are_colliding)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const ShapeVariant a = makeShape(typeA[i], geomA + 16 * i, cntA[i]);
        const ShapeVariant b = makeShape(typeB[i], geomB + 16 * i, cntB[i]);
        const Transform ta = makeRawTransform(xfA5 + 5 * i);
        const Transform tb = makeRawTransform(xfB5 + 5 * i);
        std::visit(
            [&](const auto& sa, const auto& sb)
            {
                if (out_manifold68)
                    storeManifold(c2d::collide(sa, ta, sb, tb), static_cast<char*>(out_manifold68) + 68 * i);
                if (out_are_colliding)
                    out_are_colliding[i] = c2d::areColliding(sa, ta, sb, tb) ? 1 : 0;
            },
            a,
            b);
    }
}

void rn_sweep(uint32_t n, int mode, const uint8_t* typeA, const float* geomA, const uint8_t* cntA,
              const float* xfA5, const float* xfA5_end, const uint8_t* typeB, const float* geomB,
              const uint8_t* cntB, const float* xfB5, const float* xfB5_end, float* out_fraction,
              void* out_manifold68, uint8_t* out_hit)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const ShapeVariant a = makeShape(typeA[i], geomA + 16 * i, cntA[i]);
        const ShapeVariant b = makeShape(typeB[i], geomB + 16 * i, cntB[i]);
        const Transform ta0 = makeRawTransform(xfA5 + 5 * i);
        const Transform ta1 = makeRawTransform(xfA5_end + 5 * i);
        const Transform tb0 = makeRawTransform(xfB5 + 5 * i);
        std::visit(
            [&](const auto& sa, const auto& sb)
            {
                std::optional<SweepManifold> r;
                if (mode == 2)
                    r = c2d::sweep(sa, ta0, ta1, sb, tb0, makeRawTransform(xfB5_end + 5 * i));
                else
                    r = c2d::sweep(sa, ta0, ta1, sb, tb0);
                out_hit[i] = r ? 1 : 0;
                Manifold m{};
                if (r)
                    m = r->manifold;
                if (out_fraction)
                    out_fraction[i] = r ? r->fraction : -1.0f;
                if (out_manifold68)
                    storeManifold(m, static_cast<char*>(out_manifold68) + 68 * i);
            },
            a,
            b);
    }
}

void rn_raycast_shape(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                      const float* ray4, const uint8_t* infinite, float* out_times2, uint8_t* out_hit)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const ShapeVariant s = makeShape(type[i], geom + 16 * i, cnt[i]);
        const Transform t = makeRawTransform(xf5 + 5 * i);
        const float* r = ray4 + 4 * i;
        std::optional<std::pair<float, float>> res;
        std::visit(
            [&](const auto& sc)
            {
                if (infinite && infinite[i])
                    res = c2d::raycast(sc, t, InfiniteRay{Vec2{r[0], r[1]}, Vec2{r[2], r[3]}});
                else
                    res = c2d::raycast(sc, t, Ray{Vec2{r[0], r[1]}, Vec2{r[2], r[3]}});
            },
            s);
        out_hit[i] = res ? 1 : 0;
        out_times2[2 * i] = res ? res->first : 0.0f;
        out_times2[2 * i + 1] = res ? res->second : 0.0f;
    }
}

void rn_sweep_ray(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                  const float* xf5_end, const float* ray4, const uint8_t* infinite, float* out_times2,
                  uint8_t* out_hit)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const ShapeVariant s = makeShape(type[i], geom + 16 * i, cnt[i]);
        const Transform t0 = makeRawTransform(xf5 + 5 * i);
        const Transform t1 = makeRawTransform(xf5_end + 5 * i);
        const float* r = ray4 + 4 * i;
        std::optional<std::pair<float, float>> res;
        std::visit(
            [&](const auto& sc)
            {
                if (infinite && infinite[i])
                    res = c2d::sweep(sc, t0, t1, InfiniteRay{Vec2{r[0], r[1]}, Vec2{r[2], r[3]}});
                else
                    res = c2d::sweep(sc, t0, t1, Ray{Vec2{r[0], r[1]}, Vec2{r[2], r[3]}});
            },
            s);
        out_hit[i] = res ? 1 : 0;
        out_times2[2 * i] = res ? res->first : 0.0f;
        out_times2[2 * i + 1] = res ? res->second : 0.0f;
    }
}

void rn_compute_aabb(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                     float* out_aabb4)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const ShapeVariant s = makeShape(type[i], geom + 16 * i, cnt[i]);
        const Transform t = makeRawTransform(xf5 + 5 * i);
        const AABB box = std::visit([&](const auto& sc) { return c2d::computeAABB(sc, t); }, s);
        out_aabb4[4 * i + 0] = box.min.x;
        out_aabb4[4 * i + 1] = box.min.y;
        out_aabb4[4 * i + 2] = box.max.x;
        out_aabb4[4 * i + 3] = box.max.y;
    }
}

#elif defined(RN_BACKEND_B200)

// The B200 build evaluates the same batches on the GPU through the C-ABI; the
// host side only packs shapes exactly as Registry::addShape does.
namespace
{
void packShapes(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, std::vector<float>& out32)
{
    out32.assign(size_t(n) * 32, 0.0f);
    for (uint32_t i = 0; i < n; ++i)
    {
        float* o = out32.data() + size_t(i) * 32;
        if (type[i] == 3)
        {
            const Polygon p = makePoly(geom + 16 * i, cnt[i]);
            for (uint8_t k = 0; k < p.count; ++k)
            {
                o[2 * k] = p.vertices[k].x;
                o[2 * k + 1] = p.vertices[k].y;
                o[16 + 2 * k] = p.normals[k].x;
                o[16 + 2 * k + 1] = p.normals[k].y;
            }
        }
        else
            std::memcpy(o, geom + 16 * i, 5 * sizeof(float));
    }
}
void throwOnError(int rc)
{
    if (rc != 0)
        throw std::runtime_error(std::string("c2d_gpu batch call failed: ") + c2d_gpu_last_error(nullptr));
}
} // namespace

void rn_collide(uint32_t n, const uint8_t* typeA, const float* geomA, const uint8_t* cntA, const float* xfA5,
                const uint8_t* typeB, const float* geomB, const uint8_t* cntB, const float* xfB5,
                void* out_manifold68, uint8_t* out_are_colliding)
{
    std::vector<float> a, b;
    packShapes(n, typeA, geomA, cntA, a);
    packShapes(n, typeB, geomB, cntB, b);
    throwOnError(c2d_gpu_collide_batch(n, typeA, a.data(), cntA, xfA5, typeB, b.data(), cntB, xfB5, out_manifold68,
                                       out_are_colliding));
}

void rn_sweep(uint32_t n, int mode, const uint8_t* typeA, const float* geomA, const uint8_t* cntA,
              const float* xfA5, const float* xfA5_end, const uint8_t* typeB, const float* geomB,
              const uint8_t* cntB, const float* xfB5, const float* xfB5_end, float* out_fraction,
              void* out_manifold68, uint8_t* out_hit)
{
    std::vector<float> a, b;
    packShapes(n, typeA, geomA, cntA, a);
    packShapes(n, typeB, geomB, cntB, b);
    throwOnError(c2d_gpu_sweep_batch(n, mode, typeA, a.data(), cntA, xfA5, xfA5_end, typeB, b.data(), cntB, xfB5,
                                     xfB5_end, out_fraction, out_manifold68, out_hit));
}

void rn_raycast_shape(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                      const float* ray4, const uint8_t* infinite, float* out_times2, uint8_t* out_hit)
{
    std::vector<float> s;
    packShapes(n, type, geom, cnt, s);
    throwOnError(c2d_gpu_raycast_shape_batch(n, type, s.data(), cnt, xf5, nullptr, ray4, infinite, out_times2,
                                             out_hit));
}

void rn_sweep_ray(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                  const float* xf5_end, const float* ray4, const uint8_t* infinite, float* out_times2,
                  uint8_t* out_hit)
{
    std::vector<float> s;
    packShapes(n, type, geom, cnt, s);
    throwOnError(c2d_gpu_raycast_shape_batch(n, type, s.data(), cnt, xf5, xf5_end, ray4, infinite, out_times2,
                                             out_hit));
}

void rn_compute_aabb(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                     float* out_aabb4)
{
    std::vector<float> s;
    packShapes(n, type, geom, cnt, s);
    throwOnError(c2d_gpu_compute_aabb_batch(n, type, s.data(), cnt, xf5, out_aabb4));
}

#endif

void rn_polygon_normals(uint32_t n, const float* geom, const uint8_t* cnt, float* out_normals16)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const Polygon p = makePoly(geom + 16 * i, cnt[i]);
        for (uint8_t k = 0; k < 8; ++k)
        {
            out_normals16[16 * i + 2 * k] = k < p.count ? p.normals[k].x : 0.0f;
            out_normals16[16 * i + 2 * k + 1] = k < p.count ? p.normals[k].y : 0.0f;
        }
    }
}

} // extern "C"
