// api_check.cpp — compiles against include/colli2de/Registry.hpp exactly like an application would and exercises the
// public API with EntityId types other than uint32_t (std::string, int64_t, an enum): the ids then travel through
// the GPU as dense indices and are translated back on the host.  Scenarios follow the reference's
// tests/registry/test_registry_basic.cpp and test_registry_raycast.cpp.  TEST INFRASTRUCTURE; prints "ok" on success.
#include <colli2de/Registry.hpp>

#include <algorithm>
#include <cstdio>
#include <span>
#include <vector>
#include <cstdlib>
#include <sstream>
#include <string>
#include <type_traits>

using namespace c2d;

#define CHECK(cond)                                                                                                    \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            std::fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond);                                    \
            std::exit(1);                                                                                              \
        }                                                                                                              \
    } while (0)

template <typename Id, typename MakeId>
void scenario(MakeId id)
{
    Registry<Id> reg;
    std::vector<ShapeId> shapes;
    reg.createEntity(id(1), BodyType::Static, Transform({0.0f, 0.0f}));
    shapes.push_back(reg.addShape(id(1), Circle{{0.0f, 0.0f}, 1.0f}));
    reg.createEntity(id(2), BodyType::Dynamic, Transform({3.0f, 0.0f}));
    shapes.push_back(reg.addShape(id(2), Circle{{0.0f, 0.0f}, 1.0f}));
    CHECK(!reg.areColliding(id(1), id(2)));
    CHECK(reg.getCollidingPairs().empty());
    reg.teleportEntity(id(2), Transform({1.0f, 0.0f}));
    CHECK(reg.areColliding(id(1), id(2)));
    reg.createEntity(id(3), BodyType::Dynamic, Transform({2.0f, 0.0f}));
    reg.addShape(id(3), makeRectangle({0.0f, 0.0f}, 0.6f, 0.6f));
    reg.createEntity(id(4), BodyType::Bullet, Transform({-9.0f, 0.2f}));
    reg.addShape(id(4), Capsule{{-0.3f, 0.0f}, {0.3f, 0.0f}, 0.2f});
    reg.moveEntity(id(4), Translation{18.0f, 0.0f}); // the bullet sweeps through everything
    auto pairs = reg.getCollidingPairs();
    auto has = [&](Id a, Id b)
    {
        return std::any_of(pairs.begin(), pairs.end(), [&](const auto& c)
                           { return (c.entityA == a && c.entityB == b) || (c.entityA == b && c.entityB == a); });
    };
    // expectations computed with the reference itself (oracle/_ref) on the same scenario
    CHECK(pairs.size() == 5);
    CHECK(has(id(1), id(2)) && !has(id(1), id(3)) && has(id(2), id(3)));
    CHECK(has(id(4), id(1)) && has(id(4), id(2)) && has(id(4), id(3)));
    for (const auto& c : pairs)
        CHECK(c.manifold.isColliding() && c.manifold.pointCount <= 2);
    const auto view = reg.getCollidingPairsView();
    CHECK(view.size() == pairs.size());
    // per-entity queries
    CHECK(reg.getCollisions(id(2)).size() == 3);
    for (const auto& c : reg.getCollisions(id(2)))
        CHECK(c.entityA == id(2));
    // rays
    const auto hits = reg.rayCast(Ray{{-5.0f, 0.0f}, {5.0f, 0.0f}});
    CHECK(hits.size() == 4);
    CHECK(hits.begin()->id.first == id(4)); // the bullet's ray sweep reports entry time 0
    float last = -1.0f;
    for (const auto& h : hits)
    {
        CHECK(h.entryTime >= last);
        last = h.entryTime;
    }
    const auto first = reg.firstHitRayCast(InfiniteRay{{-5.0f, 0.0f}, {1.0f, 0.0f}});
    CHECK(first && first->id.first == id(4));
    CHECK(!reg.firstHitRayCast(Ray{{-5.0f, 50.0f}, {5.0f, 50.0f}}));
    CHECK(reg.rayCast(Ray{{-5.0f, 0.0f}, {5.0f, 0.0f}}, 2).empty()); // nothing has category bit 2
    // activation, removal, transforms
    reg.setShapeActive(shapes[1], false);
    CHECK(!reg.areColliding(id(1), id(2)));
    reg.setShapeActive(shapes[1], true);
    reg.moveEntity(id(2), Transform({0.5f, 0.0f}, 0.4f));
    const Transform t = reg.getEntityTransform(id(2));
    CHECK(t.translation == Vec2(1.5f, 0.0f) && float_equals(t.rotation.angleRadians, 0.4f));
    reg.removeEntity(id(2));
    CHECK(reg.size() == 3);
    pairs = reg.getCollidingPairs();
    CHECK(!has(id(1), id(2)) && has(id(4), id(3)) && pairs.size() == 2);
    bool threw = false;
    try
    {
        reg.removeShape(shapes[1]); // already gone with its entity: the C ABI reports it, the Registry throws
    }
    catch (const std::runtime_error&)
    {
        threw = true;
    }
    CHECK(threw);
    // snapshots (the reference's wire format stores the raw bytes of the id: trivially copyable ids only)
    if constexpr (std::is_trivially_copyable_v<Id>)
    {
        std::stringstream blob(std::ios::in | std::ios::out | std::ios::binary);
        reg.serialize(blob);
        Registry<Id> loaded = Registry<Id>::deserialize(blob);
        CHECK(loaded == reg && loaded.size() == reg.size());
        CHECK(loaded.getCollidingPairs().size() == pairs.size());
        loaded.moveEntity(id(3), Translation{0.25f, 0.0f});
        CHECK(loaded != reg);
        CHECK(loaded.getEntityTransform(id(4)).translation == reg.getEntityTransform(id(4)).translation);
        // the bullet's previous transform travelled too: the sweep still hits entity 3
        auto again = loaded.getCollidingPairs();
        CHECK(std::any_of(again.begin(), again.end(), [&](const auto& c) { return c.entityA == id(4) && c.entityB == id(3); }));
    }
    // extensions with every id type: batched moves / teleports (one pass), the reference's result order
    {
        const std::vector<Id> ids{id(3), id(1)}; // (not the bullet: a move would end the sweep the expected pairs come from)
        const std::vector<Translation> back{Translation{-0.25f, 0.0f}, Translation{0.0f, 0.5f}};
        const Transform before3 = reg.getEntityTransform(id(3)), before1 = reg.getEntityTransform(id(1));
        reg.moveEntities(ids, back);
        CHECK(reg.getEntityTransform(id(3)).translation == before3.translation + Vec2(-0.25f, 0.0f));
        CHECK(reg.getEntityTransform(id(1)).translation == before1.translation + Vec2(0.0f, 0.5f));
        const std::vector<Translation> home{before3.translation, before1.translation};
        reg.teleportEntities(std::span<const Id>(ids), std::span<const Translation>(home));
        CHECK(reg.getEntityTransform(id(3)).translation == before3.translation);
        reg.setDeterministicOrder(true);
        const auto ordered = reg.getCollidingPairs();
        CHECK(ordered.size() == pairs.size());
        for (size_t k = 1; k < ordered.size(); ++k) // same group here (Bullet x ...): ascending (shapeA, shapeB) inside a group
            CHECK(ordered[k - 1].shapeA != ordered[k].shapeA || ordered[k - 1].shapeB < ordered[k].shapeB);
        reg.setDeterministicOrder(false);
    }
    Registry<Id> moved = std::move(reg);
    CHECK(moved.size() == 3 && moved.getCollidingPairs().size() == pairs.size());
    moved.clear();
    CHECK(moved.size() == 0 && moved.getCollidingPairs().empty());
}

enum class Handle : uint16_t
{
};

int main()
{
    scenario<std::string>([](int i) { return "entity-" + std::to_string(i); });
    scenario<int64_t>([](int i) { return int64_t(i) * 1000000007ll; });
    scenario<Handle>([](int i) { return static_cast<Handle>(i); });
    scenario<uint32_t>([](int i) { return uint32_t(i); });
    Registry<int, PartitioningMethod::Grid> grid(32);
    grid.createEntity(1, BodyType::Dynamic);
    grid.createEntity(2, BodyType::Dynamic, Transform({0.5f, 0.0f}));
    grid.addShape(1, Circle{{0, 0}, 1});
    grid.addShape(2, Segment{{-1, 0}, {1, 0}});
    CHECK(grid.getCollidingPairs().size() == 1);
    std::puts("ok");
    return 0;
}
