// colli2de-b200 — shape value types of the public API: Circle, Capsule,
// Segment, Polygon (<= MAX_POLYGON_VERTICES vertices + outward edge normals),
// ShapeVariant and the polygon factories.
//
// Source compatible with the reference's include/colli2de/Shapes.hpp and
// src/geometry/Shapes.cpp (the factories are inline here: this library has no
// host-side compiled geometry code, only libc2d_gpu.so).  Semantics kept:
//   * Polygon::computeNormals: normal[i] = normalized(e.y, -e.x) with
//     e = v[i+1] - v[i], via reciprocal-sqrt-then-multiply (Shapes.hpp:142-154).
//     Normals are computed ON THE HOST at construction and uploaded with the
//     vertices, exactly like the reference stores them in the shape.
//   * makeRectangle / makeTriangle / makeRegularPolygon / makePolygon vertex
//     formulas (Shapes.cpp:8-86).
#pragma once

#include <colli2de/Constants.hpp>
#include <colli2de/Vec2.hpp>

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <type_traits>
#include <variant>

namespace c2d
{

enum class ShapeType : uint8_t
{
    Circle,
    Capsule,
    Segment,
    Polygon
};

struct Circle
{
    Vec2 center{};   // local space
    float radius{1}; // >= 0

    constexpr Circle() = default;
    constexpr Circle(Vec2 center, float radius) : center(center), radius(radius) {}

    constexpr ShapeType getType() const { return ShapeType::Circle; }

    bool operator==(const Circle& o) const { return center == o.center && float_equals(radius, o.radius); }
    bool operator!=(const Circle& o) const { return !(*this == o); }
};

struct Capsule
{
    Vec2 center1{};
    Vec2 center2{};
    float radius{1};

    constexpr Capsule() = default;
    constexpr Capsule(Vec2 center1, Vec2 center2, float radius) : center1(center1), center2(center2), radius(radius) {}

    constexpr ShapeType getType() const { return ShapeType::Capsule; }

    bool operator==(const Capsule& o) const
    {
        return center1 == o.center1 && center2 == o.center2 && float_equals(radius, o.radius);
    }
    bool operator!=(const Capsule& o) const { return !(*this == o); }
};

struct Segment
{
    Vec2 start{};
    Vec2 end{};

    constexpr Segment() = default;
    constexpr Segment(Vec2 start, Vec2 end) : start(start), end(end) {}

    constexpr ShapeType getType() const { return ShapeType::Segment; }

    bool operator==(const Segment& o) const { return start == o.start && end == o.end; }
    bool operator!=(const Segment& o) const { return !(*this == o); }
};

struct Polygon
{
    std::array<Vec2, MAX_POLYGON_VERTICES> vertices{};
    std::array<Vec2, MAX_POLYGON_VERTICES> normals{}; // outward edge normals, edge i = (v[i], v[i+1])
    uint8_t count{0};

    constexpr Polygon() = default;

    template <std::size_t N>
    constexpr Polygon(const std::array<Vec2, N>& verts, uint8_t count) requires(N <= MAX_POLYGON_VERTICES)
        : count(count)
    {
        for (uint8_t i = 0; i < count; ++i)
            vertices[i] = verts[i];
        computeNormals();
    }

    constexpr ShapeType getType() const { return ShapeType::Polygon; }

    constexpr void computeNormals()
    {
        for (uint8_t i = 0; i < count; ++i)
        {
            const uint8_t next = (i + 1 == count) ? uint8_t(0) : uint8_t(i + 1);
            const Vec2 edge = vertices[next] - vertices[i];
            normals[i] = Vec2::normalized(edge.y, -edge.x);
        }
    }

    bool operator==(const Polygon& o) const
    {
        if (count != o.count)
            return false;
        for (uint8_t i = 0; i < count; ++i)
            if (vertices[i] != o.vertices[i] || normals[i] != o.normals[i])
                return false;
        return true;
    }
    bool operator!=(const Polygon& o) const { return !(*this == o); }
};

template <typename S>
concept IsShape =
    std::is_same_v<S, Circle> || std::is_same_v<S, Capsule> || std::is_same_v<S, Segment> || std::is_same_v<S, Polygon>;

inline constexpr uint8_t shapeTypesCount = 4;
using Shape0 = Circle;
using Shape1 = Capsule;
using Shape2 = Segment;
using Shape3 = Polygon;
using ShapeVariant = std::variant<Shape0, Shape1, Shape2, Shape3>;

// --- factories (reference: src/geometry/Shapes.cpp) ---------------------------------------------

// Box centred at `center`, rotated by `angle`; corners listed counter-clockwise from (-w, -h).
inline Polygon makeRectangle(Vec2 center, float halfWidth, float halfHeight, float angle = 0.0f)
{
    C2D_DEBUG_ASSERT(halfWidth > 0.0f && halfHeight > 0.0f);
    const float c = std::cos(angle);
    const float s = std::sin(angle);
    const float lx[4] = {-halfWidth, halfWidth, halfWidth, -halfWidth};
    const float ly[4] = {-halfHeight, -halfHeight, halfHeight, halfHeight};

    Polygon poly;
    poly.count = 4;
    for (uint8_t i = 0; i < 4; ++i)
        poly.vertices[i] = Vec2{lx[i] * c - ly[i] * s + center.x, lx[i] * s + ly[i] * c + center.y};
    poly.computeNormals();
    return poly;
}

inline Polygon makeTriangle(Vec2 v0, Vec2 v1, Vec2 v2)
{
    Polygon poly;
    poly.count = 3;
    poly.vertices[0] = v0;
    poly.vertices[1] = v1;
    poly.vertices[2] = v2;
    poly.computeNormals();
    return poly;
}

// Regular convex n-gon, 3 <= n <= MAX_POLYGON_VERTICES, first vertex at `rotationAngle`.
inline Polygon makeRegularPolygon(uint8_t n, Vec2 center, float radius, float rotationAngle = 0.0f)
{
    C2D_DEBUG_ASSERT(n >= 3 && n <= MAX_POLYGON_VERTICES);
    Polygon poly;
    poly.count = n;
    const float step = 2.0f * PI / float(n);
    for (uint8_t i = 0; i < n; ++i)
    {
        const float theta = rotationAngle + step * i;
        poly.vertices[i] = center + radius * Vec2{std::cos(theta), std::sin(theta)};
    }
    poly.computeNormals();
    return poly;
}

// Points are taken as-is (the reference does not build a hull either, Shapes.cpp:71-86).
inline Polygon makePolygon(const std::initializer_list<Vec2>& points)
{
    C2D_DEBUG_ASSERT(points.size() >= 3 && points.size() <= MAX_POLYGON_VERTICES);
    Polygon poly;
    poly.count = static_cast<uint8_t>(points.size());
    std::copy(points.begin(), points.end(), poly.vertices.begin());
    poly.computeNormals();
    return poly;
}

} // namespace c2d
