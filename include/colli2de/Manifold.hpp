// colli2de-b200 — contact manifold value types of the public API.
//
// Layout compatible with the reference's include/colli2de/Manifold.hpp
// (ManifoldPoint 28 B, Manifold 68 B): the CUDA narrowphase writes these
// records verbatim and the host memcpy's them into the returned vector.
//   * Manifold::reverse(): negate the normal, swap anchorA/anchorB of the
//     first pointCount points; point and separation unchanged (Manifold.hpp:43-50).
#pragma once

#include <colli2de/Vec2.hpp>

#include <cstddef>
#include <cstdint>
#include <format>
#include <string>
#include <utility>

namespace c2d
{

inline constexpr std::size_t MaxManifoldPoints = 2;

struct ManifoldPoint
{
    Vec2 point;       // world-space contact
    Vec2 anchorA;     // contact in shape A's entity frame
    Vec2 anchorB;     // contact in shape B's entity frame
    float separation; // < 0 when overlapping

    constexpr ManifoldPoint() : point(), anchorA(), anchorB(), separation(0.0f) {}

    std::string toString() const
    {
        return std::format("ManifoldPoint(point={}, anchorA={}, anchorB={}, separation={})",
                           point.toString(), anchorA.toString(), anchorB.toString(), separation);
    }
};

struct Manifold
{
    Vec2 normal;                             // from shape A towards shape B
    ManifoldPoint points[MaxManifoldPoints]; // up to two contacts
    uint8_t pointCount{0};

    constexpr Manifold() = default;

    constexpr void reverse()
    {
        normal = -normal;
        for (uint8_t i = 0; i < pointCount; ++i)
            std::swap(points[i].anchorA, points[i].anchorB);
    }

    bool isColliding() const { return pointCount > 0; }

    std::string toString() const
    {
        std::string text = "Manifold(normal=" + normal.toString() + ", points=[";
        for (uint8_t i = 0; i < pointCount; ++i)
            text += (i ? ", " : "") + points[i].toString();
        return text + "], pointCount=" + std::to_string(pointCount) + ")";
    }
};

static_assert(sizeof(ManifoldPoint) == 28 && sizeof(Manifold) == 68, "device records are written with this layout");

struct SweepManifold
{
    float fraction{};
    Manifold manifold{};
};

} // namespace c2d
