// colli2de-b200 — c2d::Rotation / c2d::Transform of the public API.
//
// Source compatible with the reference's include/colli2de/Transform.hpp.  The
// host-side Registry mirror does ALL transform arithmetic with these types and
// uploads the resulting (tx, ty, angle, sin, cos) tuples, so the following
// reference behaviours are restated exactly:
//   * Rotation(angle): sin = sinf(angle), cos = cosf(angle)   (Transform.hpp:23-28)
//   * apply   : (cos*x - sin*y, sin*x + cos*y)                 (Transform.hpp:33-36)
//   * inverse : (cos*x + sin*y, -sin*x + cos*y)                (Transform.hpp:43-46)
//   * Rotation += Rotation updates sin FIRST and then computes cos from the
//     NEW sin (Transform.hpp:75-81; same for -=, 83-89).  After a
//     moveEntity(id, Transform) the stored (sin, cos) therefore drifts away
//     from (sinf(angle), cosf(angle)) — SURVEY.md D5.  Kept on purpose: the
//     colliding-pair set must be bit-exact against the reference.
//   * Transform::apply = rotate then translate; toLocal = inverse-rotate (v - t).
#pragma once

#include <colli2de/Vec2.hpp>
#include <colli2de/internal/utils/Methods.hpp>

#include <cmath>
#include <format>
#include <string>

namespace c2d
{

struct Rotation
{
    float angleRadians = 0.0f;
    float sin = 0.0f;
    float cos = 1.0f;

    constexpr Rotation() = default;

    constexpr Rotation(float angleRadians)
        : angleRadians(angleRadians), sin(std::sin(angleRadians)), cos(std::cos(angleRadians))
    {
    }

    constexpr Rotation(float sin, float cos) : angleRadians(std::atan2(sin, cos)), sin(sin), cos(cos) {}

    constexpr Vec2 apply(Vec2 v) const { return Vec2{cos * v.x - sin * v.y, sin * v.x + cos * v.y}; }
    constexpr Vec2 inverse(Vec2 v) const { return Vec2{cos * v.x + sin * v.y, -sin * v.x + cos * v.y}; }
    constexpr Rotation inverse() const { return raw(-angleRadians, -sin, cos); }

    constexpr Rotation operator+(const Rotation& o) const
    {
        return raw(angleRadians + o.angleRadians, sin * o.cos + cos * o.sin, cos * o.cos - sin * o.sin);
    }
    constexpr Rotation operator-(const Rotation& o) const
    {
        return raw(angleRadians - o.angleRadians, sin * o.cos - cos * o.sin, cos * o.cos + sin * o.sin);
    }
    constexpr Rotation operator+(float angle) const { return Rotation{angleRadians + angle}; }
    constexpr Rotation operator-(float angle) const { return Rotation{angleRadians - angle}; }

    // NOTE (D5): `cos` below is computed from the already-updated `sin`, as in the reference.
    constexpr Rotation& operator+=(const Rotation& o)
    {
        angleRadians += o.angleRadians;
        sin = sin * o.cos + cos * o.sin;
        cos = cos * o.cos - sin * o.sin;
        return *this;
    }
    constexpr Rotation& operator-=(const Rotation& o)
    {
        angleRadians -= o.angleRadians;
        sin = sin * o.cos - cos * o.sin;
        cos = cos * o.cos + sin * o.sin;
        return *this;
    }
    constexpr Rotation& operator+=(float angle) { return *this = Rotation{angleRadians + angle}; }
    constexpr Rotation& operator-=(float angle) { return *this = Rotation{angleRadians - angle}; }

    constexpr bool operator==(const Rotation& o) const { return float_equals(angleRadians, o.angleRadians); }
    constexpr bool operator!=(const Rotation& o) const { return !(*this == o); }

    std::string toString() const
    {
        return std::format("Rotation(angleRadians={}, sin={}, cos={})", angleRadians, sin, cos);
    }

  private:
    static constexpr Rotation raw(float angle, float s, float c)
    {
        Rotation r;
        r.angleRadians = angle;
        r.sin = s;
        r.cos = c;
        return r;
    }
};

using Translation = Vec2;

struct Transform
{
    Translation translation{};
    Rotation rotation{};

    constexpr Transform() = default;
    constexpr Transform(Translation translation) : translation(translation), rotation() {}
    constexpr Transform(Translation translation, float angleRadians) : translation(translation), rotation(angleRadians)
    {
    }
    constexpr Transform(Translation translation, Rotation rotation) : translation(translation), rotation(rotation) {}

    constexpr Vec2 apply(Vec2 v) const { return rotation.apply(v) + translation; }
    constexpr Vec2 toLocal(Vec2 v) const { return rotation.inverse(v - translation); }

    constexpr Transform operator+(const Transform& o) const
    {
        return Transform{translation + o.translation, rotation + o.rotation};
    }
    constexpr Transform operator-(const Transform& o) const
    {
        return Transform{translation - o.translation, rotation - o.rotation};
    }
    constexpr Transform operator+(float s) const { return Transform{translation + s, rotation + s}; }
    constexpr Transform operator-(float s) const { return Transform{translation - s, rotation - s}; }
    constexpr Transform operator*(float s) const { return Transform{translation * s, rotation.angleRadians * s}; }
    constexpr Transform operator/(float s) const { return Transform{translation / s, rotation.angleRadians / s}; }

    constexpr Transform& operator+=(const Transform& o)
    {
        translation += o.translation;
        rotation += o.rotation;
        return *this;
    }
    constexpr Transform& operator-=(const Transform& o)
    {
        translation -= o.translation;
        rotation -= o.rotation;
        return *this;
    }
    constexpr Transform& operator+=(float s)
    {
        translation += s;
        rotation += s;
        return *this;
    }
    constexpr Transform& operator-=(float s)
    {
        translation -= s;
        rotation -= s;
        return *this;
    }
    constexpr Transform& operator*=(float s)
    {
        translation *= s;
        rotation = Rotation{rotation.angleRadians * s};
        return *this;
    }
    constexpr Transform& operator/=(float s)
    {
        translation /= s;
        rotation = Rotation{rotation.angleRadians / s};
        return *this;
    }

    constexpr bool operator==(const Transform& o) const
    {
        return translation == o.translation && rotation == o.rotation;
    }
    constexpr bool operator!=(const Transform& o) const { return !(*this == o); }

    std::string toString() const
    {
        return std::format("Transform(translation={}, rotation={})", translation.toString(), rotation.toString());
    }
};

} // namespace c2d

template <>
struct std::formatter<c2d::Rotation> : std::formatter<std::string>
{
    template <typename FormatContext>
    auto format(const c2d::Rotation& r, FormatContext& ctx) const
    {
        return std::formatter<std::string>::format(r.toString(), ctx);
    }
};

template <>
struct std::formatter<c2d::Transform> : std::formatter<std::string>
{
    template <typename FormatContext>
    auto format(const c2d::Transform& t, FormatContext& ctx) const
    {
        return std::formatter<std::string>::format(t.toString(), ctx);
    }
};
