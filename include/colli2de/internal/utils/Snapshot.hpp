// Snapshot.hpp — the reference's binary Registry snapshot (wire format), host-only code.
//
// Reads what c2d::Registry<EntityId, PartitioningMethod::None>::serialize of the reference writes and writes what its
// deserialize reads (reference Registry.hpp:1205-1344 registry, 1445-1494 EntityInfo, 1493-1649 ShapeInstance;
// DynamicBVH.hpp:1033-1079 tree header, 1113-1147 BVHNode; Serialization.hpp:19-77 = raw native-endian bytes of every
// field, no padding, size_t counts).  Layout of a snapshot:
//
//   size_t E;  E x { EntityId id; u8 bodyType; f32 tx, ty, angle, sin, cos; size_t S; ShapeId[S] }
//   size_t N;  N x EntityId                                   (shape -> entity, stale for freed ids)
//   size_t N;  N x { u8 type; geometry; f32 aabb[4]; u64 category, mask; u8 moveFnc[2]; u8 translateFnc; u8 active }
//   size_t N;  N x size_t                                     (shape -> leaf index in its tree)
//   size_t F;  F x ShapeId                                    (free list, popped from the back)
//   3 x tree (Static, Dynamic, Bullet): f32 margin; i32 root; u32 nodeCount; u32 proxyCount; i32 nextFree; size_t C;
//              C x { f32 box[4]; u64 category, hitting; i32 parent, child1, child2, height; ShapeId id }
//   size_t B;  B x { EntityId id; f32 tx, ty, angle, sin, cos }   (bullets' previous transforms)
//   u32 previousAllCollisionsCount
//
// The GPU library keeps no pointer tree, so a snapshot written here carries freshly built, VALID DynamicBVH images
// (balanced median splits over the Morton-sorted leaves, exact heights / parents / unioned boxes and category bits, a
// power-of-two node array whose tail is the free list) — the reference loads them and keeps inserting, moving and
// removing proxies on top.  The per-shape state that IS history — the tight box and the leaf (fat) box — is carried
// bit for bit in both directions.
//
// PartitioningMethod::Grid (reference BroadPhaseTree.hpp:539-733): each of the three trees is
//   i32 cellSize; size_t P; P x { ShapeId id; f32 box[4]; u64 category, mask; size_t H; H x { i32 cellX, cellY; i32 node } }
//   size_t R; R x { i32 cellX, cellY; size_t bvhIndex };   size_t B; B x { i32 cellX, cellY; tree (as above) }
//   size_t freeProxies; ... ; size_t freeBvhs; ...
// A proxy is registered in every cell its box touches (getCellFor, BroadPhaseTree.hpp:71-79, quirks included) and owns
// one leaf per cell.  In the reference those per-cell leaves have per-cell histories; the GPU library keeps ONE leaf box
// per shape, so a snapshot written here gives every cell leaf that box (it contains the tight box, which is all the
// reference's queries need), and a snapshot read here takes the union of the shape's cell leaves.  Query RESULTS are
// unaffected either way — leaf boxes only decide which pairs reach the narrowphase.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace c2d::snapshot
{

inline constexpr int32_t kInvalidNode = -99; // DynamicBVH.hpp:28
inline constexpr uint32_t kInitialTreeCapacity = 64; // DynamicBVH.hpp:77

struct ShapeState
{
    uint8_t type = 0;        // ShapeType / variant index: 0 circle, 1 capsule, 2 segment, 3 polygon
    uint8_t vertexCount = 0; // polygons
    uint8_t body = 0;        // mMoveFncIndices.first == mTranslateFncIndex == BodyType of the owning entity
    bool active = true;
    bool alive = false;      // false: the id is in the free list, everything else in this record is stale
    float geom[32] = {};     // circle cx,cy,r | capsule c1,c2,r | segment start,end | polygon verts[0..16) normals[16..32)
    float tight[4] = {};     // ShapeInstance::mBoundingBox
    float fat[4] = {};       // BVHNode::mBoundingBox of the shape's leaf
    uint64_t category = 1, mask = ~0ull;
};

template <typename EntityId>
struct EntityState
{
    EntityId id{};
    uint8_t body = 0;
    float xf[5] = {};   // tx, ty, angle, sin, cos
    float prev[5] = {}; // bullets
    bool hasPrev = false;
    std::vector<uint32_t> shapeIds;
};

template <typename EntityId>
struct RegistryState
{
    std::vector<EntityState<EntityId>> entities;
    std::vector<EntityId> shapeEntity;
    std::vector<ShapeState> shapes;
    std::vector<uint32_t> freeShapeIds;
    uint32_t previousAllCollisionsCount = 1024;
    int32_t gridCellSize = 0; // > 0: a PartitioningMethod::Grid registry (BroadPhaseTree images instead of DynamicBVHs)
};

namespace detail
{

template <typename T>
void put(std::ostream& out, const T& v)
{
    static_assert(std::is_trivially_copyable_v<T>);
    out.write(reinterpret_cast<const char*>(&v), sizeof(T));
}

template <typename T>
void get(std::istream& in, T& v)
{
    static_assert(std::is_trivially_copyable_v<T>);
    in.read(reinterpret_cast<char*>(&v), sizeof(T));
    if (!in)
        throw std::runtime_error("colli2de snapshot: unexpected end of stream");
}

inline size_t getCount(std::istream& in, size_t bytesPerItem)
{
    size_t n = 0;
    get(in, n);
    // a corrupt count must not turn into a multi-terabyte allocation
    if (bytesPerItem != 0 && n > (size_t(1) << 40) / bytesPerItem)
        throw std::runtime_error("colli2de snapshot: implausible element count");
    return n;
}

struct Node // BVHNode<ShapeId>
{
    float box[4] = {0, 0, 0, 0};
    uint64_t category = 1, hitting = ~0ull;
    int32_t parent = kInvalidNode, child1 = kInvalidNode, child2 = kInvalidNode, height = -1;
    uint32_t id = 0;
};

struct TreeImage // DynamicBVH<ShapeId>
{
    float margin = 0.0f;
    int32_t root = kInvalidNode;
    uint32_t nodeCount = 0, proxyCount = 0;
    int32_t nextFree = kInvalidNode;
    std::vector<Node> nodes;
};

inline void putTree(std::ostream& out, const TreeImage& t)
{
    put(out, t.margin);
    put(out, t.root);
    put(out, t.nodeCount);
    put(out, t.proxyCount);
    put(out, t.nextFree);
    put(out, static_cast<size_t>(t.nodes.size()));
    for (const Node& n : t.nodes)
    {
        for (float f : n.box)
            put(out, f);
        put(out, n.category);
        put(out, n.hitting);
        put(out, n.parent);
        put(out, n.child1);
        put(out, n.child2);
        put(out, n.height);
        put(out, n.id);
    }
}

inline TreeImage getTree(std::istream& in)
{
    TreeImage t;
    get(in, t.margin);
    get(in, t.root);
    get(in, t.nodeCount);
    get(in, t.proxyCount);
    get(in, t.nextFree);
    t.nodes.resize(getCount(in, 56));
    for (Node& n : t.nodes)
    {
        for (float& f : n.box)
            get(in, f);
        get(in, n.category);
        get(in, n.hitting);
        get(in, n.parent);
        get(in, n.child1);
        get(in, n.child2);
        get(in, n.height);
        get(in, n.id);
    }
    return t;
}

inline uint32_t spreadBits(uint32_t v) // 16 bits -> every other bit
{
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

struct Leaf
{
    uint32_t shape;
    const ShapeState* state;
};

// A valid DynamicBVH over `leaves`: leaf i of the list is node i (so the shape's tree handle is its position in the
// list), internal nodes follow in post-order, the rest of the power-of-two array is the free list.
inline TreeImage buildTree(const std::vector<Leaf>& leaves)
{
    TreeImage t;
    const uint32_t n = static_cast<uint32_t>(leaves.size());
    const uint32_t used = n == 0 ? 0u : 2u * n - 1u;
    uint32_t capacity = kInitialTreeCapacity;
    while (capacity < used)
        capacity <<= 1;
    t.nodes.resize(capacity);
    t.nodeCount = used;
    t.proxyCount = n;
    for (uint32_t i = 0; i < n; ++i)
    {
        Node& leaf = t.nodes[i];
        std::memcpy(leaf.box, leaves[i].state->fat, sizeof(leaf.box));
        leaf.category = leaves[i].state->category;
        leaf.hitting = leaves[i].state->mask;
        leaf.child1 = leaf.child2 = -1; // isLeaf(): mChild1Index == -1
        leaf.height = 0;
        leaf.id = leaves[i].shape;
    }
    if (n > 0)
    {
        // spatial order: Morton code of the leaf-box centres on a 16-bit grid over the scene bounds
        float lo[2] = {leaves[0].state->fat[0], leaves[0].state->fat[1]}, hi[2] = {lo[0], lo[1]};
        std::vector<float> cx(n), cy(n);
        for (uint32_t i = 0; i < n; ++i)
        {
            const float* b = leaves[i].state->fat;
            cx[i] = 0.5f * (b[0] + b[2]);
            cy[i] = 0.5f * (b[1] + b[3]);
            lo[0] = std::min(lo[0], cx[i]), hi[0] = std::max(hi[0], cx[i]);
            lo[1] = std::min(lo[1], cy[i]), hi[1] = std::max(hi[1], cy[i]);
        }
        const float sx = hi[0] > lo[0] ? 65535.0f / (hi[0] - lo[0]) : 0.0f;
        const float sy = hi[1] > lo[1] ? 65535.0f / (hi[1] - lo[1]) : 0.0f;
        std::vector<std::pair<uint32_t, uint32_t>> order(n); // (code, leaf index)
        for (uint32_t i = 0; i < n; ++i)
        {
            const auto qx = static_cast<uint32_t>(std::clamp((cx[i] - lo[0]) * sx, 0.0f, 65535.0f));
            const auto qy = static_cast<uint32_t>(std::clamp((cy[i] - lo[1]) * sy, 0.0f, 65535.0f));
            order[i] = {spreadBits(qx) | (spreadBits(qy) << 1), i};
        }
        std::sort(order.begin(), order.end());

        uint32_t nextInternal = n;
        // explicit stack of [first, last) ranges; a range is finished (emitted) once both halves are
        struct Frame
        {
            uint32_t first, last;
            int32_t left, right; // node indices of the finished halves, kInvalidNode while pending
        };
        std::vector<Frame> stack;
        stack.push_back({0, n, kInvalidNode, kInvalidNode});
        int32_t result = kInvalidNode;
        while (!stack.empty())
        {
            Frame& f = stack.back();
            if (f.last - f.first == 1)
            {
                result = static_cast<int32_t>(order[f.first].second);
                stack.pop_back();
            }
            else if (f.left == kInvalidNode)
            {
                const uint32_t mid = f.first + (f.last - f.first) / 2;
                stack.push_back({f.first, mid, kInvalidNode, kInvalidNode});
                continue;
            }
            else if (f.right == kInvalidNode)
            {
                const uint32_t mid = f.first + (f.last - f.first) / 2;
                stack.push_back({mid, f.last, kInvalidNode, kInvalidNode});
                continue;
            }
            else
            {
                const int32_t self = static_cast<int32_t>(nextInternal++);
                Node& node = t.nodes[self];
                Node& a = t.nodes[f.left];
                Node& b = t.nodes[f.right];
                node.child1 = f.left;
                node.child2 = f.right;
                a.parent = b.parent = self;
                node.box[0] = std::min(a.box[0], b.box[0]);
                node.box[1] = std::min(a.box[1], b.box[1]);
                node.box[2] = std::max(a.box[2], b.box[2]);
                node.box[3] = std::max(a.box[3], b.box[3]);
                node.category = a.category | b.category;
                node.hitting = ~0ull;
                node.height = 1 + std::max(a.height, b.height);
                node.id = 0;
                result = self;
                stack.pop_back();
            }
            // hand the finished subtree to the parent frame
            if (!stack.empty())
            {
                Frame& parent = stack.back();
                if (parent.left == kInvalidNode)
                    parent.left = result;
                else
                    parent.right = result;
            }
        }
        t.root = result;
        t.nodes[t.root].parent = kInvalidNode;
    }
    // free list through mParentIndex (DynamicBVH.hpp:283-295)
    for (uint32_t i = used; i < capacity; ++i)
        t.nodes[i].parent = i + 1 < capacity ? static_cast<int32_t>(i + 1) : kInvalidNode;
    t.nextFree = used < capacity ? static_cast<int32_t>(used) : kInvalidNode;
    return t;
}


// ---- PartitioningMethod::Grid ------------------------------------------------------------------------------------
struct Cell
{
    int32_t x, y;
    bool operator<(const Cell& o) const { return x < o.x || (x == o.x && y < o.y); } // GridCell::operator<
    bool operator==(const Cell& o) const { return x == o.x && y == o.y; }
};

// getCellFor (BroadPhaseTree.hpp:71-79), including its treatment of negative positions
inline Cell cellFor(float px, float py, int32_t cellSize)
{
    const int32_t ix = static_cast<int32_t>(px), iy = static_cast<int32_t>(py);
    return Cell{ix - (px >= 0 ? ix % cellSize : (ix % cellSize) + cellSize),
                iy - (py >= 0 ? iy % cellSize : (iy % cellSize) + cellSize)};
}

struct GridEntry // one (cell, shape) registration
{
    Cell cell;
    uint32_t leaf; // index into the tree's leaf list
};

// BroadPhaseTree<ShapeId>::serialize for the shapes in `leaves` (proxy handle = position in the list)
inline void putGridTree(std::ostream& out, const std::vector<Leaf>& leaves, int32_t cellSize)
{
    std::vector<GridEntry> entries;
    for (uint32_t i = 0; i < leaves.size(); ++i)
    {
        const float* b = leaves[i].state->tight; // the proxy's box is the last box handed to moveProxy: the tight box
        const Cell lo = cellFor(b[0], b[1], cellSize), hi = cellFor(b[2], b[3], cellSize);
        for (int32_t y = lo.y; y <= hi.y; y += cellSize) // forEachCell (BroadPhaseTree.hpp:81-90)
            for (int32_t x = lo.x; x <= hi.x; x += cellSize)
                entries.push_back(GridEntry{Cell{x, y}, i});
    }
    std::stable_sort(entries.begin(), entries.end(), [](const GridEntry& a, const GridEntry& b) { return a.cell < b.cell; });
    // one DynamicBVH per occupied cell, in cell order; node index of (cell, leaf) = position inside the cell's list
    std::vector<std::pair<Cell, TreeImage>> bvhs;
    std::vector<std::vector<std::pair<Cell, int32_t>>> handles(leaves.size());
    for (size_t first = 0; first < entries.size();)
    {
        size_t last = first;
        std::vector<Leaf> cellLeaves;
        while (last < entries.size() && entries[last].cell == entries[first].cell)
        {
            handles[entries[last].leaf].push_back({entries[first].cell, static_cast<int32_t>(cellLeaves.size())});
            cellLeaves.push_back(leaves[entries[last].leaf]);
            ++last;
        }
        bvhs.emplace_back(entries[first].cell, buildTree(cellLeaves));
        first = last;
    }
    put(out, cellSize);
    put(out, static_cast<size_t>(leaves.size()));
    for (uint32_t i = 0; i < leaves.size(); ++i)
    {
        const ShapeState& s = *leaves[i].state;
        put(out, leaves[i].shape);
        for (float f : s.tight)
            put(out, f);
        put(out, s.category);
        put(out, s.mask);
        put(out, static_cast<size_t>(handles[i].size()));
        for (const auto& [cell, node] : handles[i]) // already in std::map order: entries were sorted by cell
        {
            put(out, cell.x);
            put(out, cell.y);
            put(out, node);
        }
    }
    put(out, static_cast<size_t>(bvhs.size()));
    for (size_t i = 0; i < bvhs.size(); ++i)
    {
        put(out, bvhs[i].first.x);
        put(out, bvhs[i].first.y);
        put(out, i);
    }
    put(out, static_cast<size_t>(bvhs.size()));
    for (const auto& [cell, tree] : bvhs)
    {
        put(out, cell.x);
        put(out, cell.y);
        putTree(out, tree);
    }
    put(out, static_cast<size_t>(0)); // proxies free list
    put(out, static_cast<size_t>(0)); // BVH free list
}

struct GridProxy
{
    uint32_t id = 0;
    float box[4] = {};
    uint64_t category = 1, mask = ~0ull;
    std::vector<std::pair<Cell, int32_t>> handles;
};

struct GridTreeImage
{
    int32_t cellSize = 0;
    std::vector<GridProxy> proxies;
    std::vector<std::pair<Cell, size_t>> regions; // sorted by cell
    std::vector<std::pair<Cell, TreeImage>> bvhs;
};

inline GridTreeImage getGridTree(std::istream& in)
{
    GridTreeImage t;
    get(in, t.cellSize);
    if (t.cellSize <= 0)
        throw std::runtime_error("colli2de snapshot: invalid grid cell size");
    t.proxies.resize(getCount(in, 44));
    for (GridProxy& p : t.proxies)
    {
        get(in, p.id);
        for (float& f : p.box)
            get(in, f);
        get(in, p.category);
        get(in, p.mask);
        p.handles.resize(getCount(in, 12));
        for (auto& [cell, node] : p.handles)
        {
            get(in, cell.x);
            get(in, cell.y);
            get(in, node);
        }
    }
    t.regions.resize(getCount(in, 16));
    for (auto& [cell, index] : t.regions)
    {
        get(in, cell.x);
        get(in, cell.y);
        get(in, index);
    }
    std::sort(t.regions.begin(), t.regions.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    t.bvhs.resize(getCount(in, 36));
    for (auto& [cell, tree] : t.bvhs)
    {
        get(in, cell.x);
        get(in, cell.y);
        tree = getTree(in);
    }
    size_t skip = getCount(in, sizeof(size_t));
    for (size_t i = 0, v = 0; i < skip; ++i)
        get(in, v);
    skip = getCount(in, sizeof(size_t));
    for (size_t i = 0, v = 0; i < skip; ++i)
        get(in, v);
    return t;
}

// Union of the leaves the proxy `handle` owns in its cells; false if the image is inconsistent.
inline bool gridLeafBox(const GridTreeImage& t, size_t handle, uint32_t shape, float* box)
{
    if (handle >= t.proxies.size() || t.proxies[handle].id != shape || t.proxies[handle].handles.empty())
        return false;
    bool first = true;
    for (const auto& [cell, node] : t.proxies[handle].handles)
    {
        const auto it = std::lower_bound(t.regions.begin(), t.regions.end(), cell,
                                         [](const auto& r, const Cell& c) { return r.first < c; });
        if (it == t.regions.end() || !(it->first == cell) || it->second >= t.bvhs.size())
            return false;
        const TreeImage& bvh = t.bvhs[it->second].second;
        if (node < 0 || static_cast<size_t>(node) >= bvh.nodes.size() || bvh.nodes[node].child1 != -1 || bvh.nodes[node].id != shape)
            return false;
        const float* b = bvh.nodes[node].box;
        if (first)
            std::memcpy(box, b, 4 * sizeof(float));
        else
        {
            box[0] = std::min(box[0], b[0]), box[1] = std::min(box[1], b[1]);
            box[2] = std::max(box[2], b[2]), box[3] = std::max(box[3], b[3]);
        }
        first = false;
    }
    return true;
}

} // namespace detail

// Registry<EntityId, Method>::serialize (reference Registry.hpp:1205-1264); state.gridCellSize selects the tree format.
template <typename EntityId>
void write(std::ostream& out, const RegistryState<EntityId>& state)
{
    using namespace detail;
    static_assert(std::is_trivially_copyable_v<EntityId>, "snapshots store the raw bytes of the entity id");
    put(out, static_cast<size_t>(state.entities.size()));
    for (const EntityState<EntityId>& e : state.entities)
    {
        put(out, e.id);
        put(out, e.body);
        for (float f : e.xf)
            put(out, f);
        put(out, static_cast<size_t>(e.shapeIds.size()));
        for (uint32_t s : e.shapeIds)
            put(out, s);
    }
    const size_t shapeCount = state.shapes.size();
    put(out, shapeCount);
    for (size_t i = 0; i < shapeCount; ++i)
        put(out, i < state.shapeEntity.size() ? state.shapeEntity[i] : EntityId{});

    // leaves of the three trees in shape order; a shape's handle is its position among its tree's leaves
    std::vector<Leaf> leaves[3];
    std::vector<size_t> handles(shapeCount, 0);
    for (size_t i = 0; i < shapeCount; ++i)
    {
        const ShapeState& s = state.shapes[i];
        if (!s.alive)
            continue;
        if (s.body > 2)
            throw std::runtime_error("colli2de snapshot: invalid body type");
        handles[i] = leaves[s.body].size();
        leaves[s.body].push_back(Leaf{static_cast<uint32_t>(i), &s});
    }

    put(out, shapeCount);
    for (const ShapeState& s : state.shapes)
    {
        put(out, s.type);
        switch (s.type)
        {
        case 0:
            for (int k = 0; k < 3; ++k)
                put(out, s.geom[k]);
            break;
        case 1:
            for (int k = 0; k < 5; ++k)
                put(out, s.geom[k]);
            break;
        case 2:
            for (int k = 0; k < 4; ++k)
                put(out, s.geom[k]);
            break;
        case 3:
            put(out, s.vertexCount);
            for (int k = 0; k < 2 * s.vertexCount; ++k)
                put(out, s.geom[k]);
            for (int k = 0; k < 2 * s.vertexCount; ++k)
                put(out, s.geom[16 + k]);
            break;
        default:
            throw std::runtime_error("colli2de snapshot: invalid shape type");
        }
        for (float f : s.tight)
            put(out, f);
        put(out, s.category);
        put(out, s.mask);
        put(out, s.body);  // mMoveFncIndices.first  (Registry.hpp:309)
        put(out, s.type);  // mMoveFncIndices.second
        put(out, s.body);  // mTranslateFncIndex      (Registry.hpp:310)
        put(out, static_cast<uint8_t>(s.active ? 1 : 0));
    }
    put(out, shapeCount);
    for (size_t h : handles)
        put(out, h);
    put(out, static_cast<size_t>(state.freeShapeIds.size()));
    for (uint32_t s : state.freeShapeIds)
        put(out, s);

    for (int body = 0; body < 3; ++body)
        if (state.gridCellSize > 0)
            putGridTree(out, leaves[body], state.gridCellSize);
        else
            putTree(out, buildTree(leaves[body]));

    size_t bullets = 0;
    for (const EntityState<EntityId>& e : state.entities)
        bullets += e.hasPrev ? 1 : 0;
    put(out, bullets);
    for (const EntityState<EntityId>& e : state.entities)
        if (e.hasPrev)
        {
            put(out, e.id);
            for (float f : e.prev)
                put(out, f);
        }
    put(out, state.previousAllCollisionsCount);
    if (!out)
        throw std::runtime_error("colli2de snapshot: write failed");
}

// Registry<EntityId, Method>::deserialize (reference Registry.hpp:1266-1344); grid = PartitioningMethod::Grid.
template <typename EntityId>
RegistryState<EntityId> read(std::istream& in, bool grid = false)
{
    using namespace detail;
    static_assert(std::is_trivially_copyable_v<EntityId>, "snapshots store the raw bytes of the entity id");
    RegistryState<EntityId> state;
    state.entities.resize(getCount(in, sizeof(EntityId) + 29));
    for (EntityState<EntityId>& e : state.entities)
    {
        get(in, e.id);
        get(in, e.body);
        for (float& f : e.xf)
            get(in, f);
        e.shapeIds.resize(getCount(in, sizeof(uint32_t)));
        for (uint32_t& s : e.shapeIds)
            get(in, s);
        if (e.body > 2)
            throw std::runtime_error("colli2de snapshot: invalid body type");
    }
    state.shapeEntity.resize(getCount(in, sizeof(EntityId)));
    for (EntityId& id : state.shapeEntity)
        get(in, id);
    state.shapes.resize(getCount(in, 36));
    for (ShapeState& s : state.shapes)
    {
        get(in, s.type);
        switch (s.type)
        {
        case 0:
            for (int k = 0; k < 3; ++k)
                get(in, s.geom[k]);
            break;
        case 1:
            for (int k = 0; k < 5; ++k)
                get(in, s.geom[k]);
            break;
        case 2:
            for (int k = 0; k < 4; ++k)
                get(in, s.geom[k]);
            break;
        case 3:
            get(in, s.vertexCount);
            if (s.vertexCount > 8)
                throw std::runtime_error("colli2de snapshot: polygon with more than 8 vertices");
            for (int k = 0; k < 2 * s.vertexCount; ++k)
                get(in, s.geom[k]);
            for (int k = 0; k < 2 * s.vertexCount; ++k)
                get(in, s.geom[16 + k]);
            break;
        default:
            throw std::runtime_error("colli2de snapshot: invalid shape type");
        }
        for (float& f : s.tight)
            get(in, f);
        get(in, s.category);
        get(in, s.mask);
        uint8_t moveFirst = 0, moveSecond = 0, translate = 0, active = 0;
        get(in, moveFirst);
        get(in, moveSecond);
        get(in, translate);
        get(in, active);
        s.body = moveFirst;
        s.active = active != 0;
    }
    std::vector<size_t> handles(getCount(in, sizeof(size_t)));
    for (size_t& h : handles)
        get(in, h);
    state.freeShapeIds.resize(getCount(in, sizeof(uint32_t)));
    for (uint32_t& s : state.freeShapeIds)
        get(in, s);
    TreeImage trees[3];
    GridTreeImage gridTrees[3];
    for (int body = 0; body < 3; ++body)
        if (grid)
            gridTrees[body] = getGridTree(in);
        else
            trees[body] = getTree(in);
    if (grid)
        state.gridCellSize = gridTrees[0].cellSize;
    const size_t bullets = getCount(in, sizeof(EntityId) + 20);
    std::vector<std::pair<EntityId, std::vector<float>>> previous(bullets);
    for (auto& [id, xf] : previous)
    {
        get(in, id);
        xf.resize(5);
        for (float& f : xf)
            get(in, f);
    }
    get(in, state.previousAllCollisionsCount);

    if (handles.size() != state.shapes.size() || state.shapeEntity.size() != state.shapes.size())
        throw std::runtime_error("colli2de snapshot: inconsistent shape tables");
    // a shape is alive iff an entity lists it; its body type is its entity's, its leaf box comes from that tree
    for (const EntityState<EntityId>& e : state.entities)
        for (uint32_t shape : e.shapeIds)
        {
            if (shape >= state.shapes.size() || state.shapes[shape].alive)
                throw std::runtime_error("colli2de snapshot: invalid shape id in an entity");
            ShapeState& s = state.shapes[shape];
            s.alive = true;
            s.body = e.body;
            const size_t h = handles[shape];
            if (grid)
            {
                if (!gridLeafBox(gridTrees[e.body], h, shape, s.fat))
                    throw std::runtime_error("colli2de snapshot: shape handle does not name its grid proxy");
                continue;
            }
            const TreeImage& t = trees[e.body];
            if (h >= t.nodes.size() || t.nodes[h].child1 != -1 || t.nodes[h].id != shape)
                throw std::runtime_error("colli2de snapshot: shape handle does not name its leaf");
            std::memcpy(s.fat, t.nodes[h].box, sizeof(s.fat));
        }
    for (uint32_t shape : state.freeShapeIds)
        if (shape >= state.shapes.size() || state.shapes[shape].alive)
            throw std::runtime_error("colli2de snapshot: free list names a live shape");
    if (!previous.empty())
    {
        // byte-wise ordering of the ids (they need not be comparable types): sort once, binary-search each bullet
        const auto bytesLess = [](const EntityId& a, const EntityId& b) { return std::memcmp(&a, &b, sizeof(EntityId)) < 0; };
        std::vector<uint32_t> byId(state.entities.size());
        for (uint32_t i = 0; i < byId.size(); ++i)
            byId[i] = i;
        std::sort(byId.begin(), byId.end(),
                  [&](uint32_t a, uint32_t b) { return bytesLess(state.entities[a].id, state.entities[b].id); });
        for (auto& [id, xf] : previous)
        {
            const auto it = std::lower_bound(byId.begin(), byId.end(), id,
                                             [&](uint32_t a, const EntityId& key) { return bytesLess(state.entities[a].id, key); });
            if (it == byId.end() || std::memcmp(&state.entities[*it].id, &id, sizeof(EntityId)) != 0)
                continue; // previous transform of an entity that no longer exists: the reference would keep it, unused
            EntityState<EntityId>& e = state.entities[*it];
            e.hasPrev = true;
            std::copy(xf.begin(), xf.end(), e.prev);
        }
    }
    return state;
}

} // namespace c2d::snapshot
