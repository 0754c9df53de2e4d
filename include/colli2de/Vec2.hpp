// colli2de-b200 — c2d::Vec2, the 2-float value type of the public API.
//
// Source compatible with the reference's include/colli2de/Vec2.hpp (same member
// names and operator set).  Semantics that the collision path relies on and
// that therefore must not drift:
//   * normalize()/normalized(): reciprocal length first, then two multiplies
//     (reference Vec2.hpp:42-46,123-127) — NOT a per-component division;
//   * operator/(float) is a true division per component (Vec2.hpp:190-195);
//   * operator< / operator> are strict lexicographic on (x, y)
//     (Vec2.hpp:288-304) — contact points are ordered with it;
//   * operator== is the epsilon comparison float_equals per component
//     (Vec2.hpp:278-281), not bitwise.
#pragma once

#include <colli2de/internal/utils/Debug.hpp>
#include <colli2de/internal/utils/Methods.hpp>

#include <cmath>
#include <format>
#include <ostream>
#include <string>
#include <type_traits>

namespace c2d
{

class Vec2
{
  public:
    float x;
    float y;

    constexpr Vec2() : x(0.0f), y(0.0f) {}
    constexpr Vec2(float x, float y) : x(x), y(y) {}

    // --- measures ---------------------------------------------------------------------------
    constexpr float lengthSqr() const { return x * x + y * y; }
    constexpr float length() const { return std::sqrt(x * x + y * y); }
    constexpr float inverseLength() const { return 1.0f / std::sqrt(x * x + y * y); }
    constexpr float dot(Vec2 o) const { return x * o.x + y * o.y; }
    constexpr float cross(Vec2 o) const { return x * o.y - y * o.x; }
    constexpr float getBigger() const { return x > y ? x : y; }
    constexpr Vec2 abs() const { return Vec2(std::abs(x), std::abs(y)); }

    constexpr Vec2 interpolate(Vec2 o, float percentage) const
    {
        return Vec2(x + (o.x - x) * percentage, y + (o.y - y) * percentage);
    }

    constexpr Vec2 normalize() const
    {
        const float k = inverseLength();
        return Vec2(x * k, y * k);
    }

    static constexpr Vec2 normalized(float x, float y)
    {
        const float k = 1.0f / std::sqrt(x * x + y * y);
        return Vec2(x * k, y * k);
    }

    // --- compound assignment ------------------------------------------------------------------
    constexpr Vec2& operator+=(Vec2 o) { x += o.x; y += o.y; return *this; }
    constexpr Vec2& operator-=(Vec2 o) { x -= o.x; y -= o.y; return *this; }
    constexpr Vec2& operator*=(Vec2 o) { x *= o.x; y *= o.y; return *this; }
    constexpr Vec2& operator/=(Vec2 o) { x /= o.x; y /= o.y; return *this; }
    constexpr Vec2& operator+=(float s) { x += s; y += s; return *this; }
    constexpr Vec2& operator-=(float s) { x -= s; y -= s; return *this; }
    constexpr Vec2& operator*=(float s) { x *= s; y *= s; return *this; }
    constexpr Vec2& operator/=(float s) { x /= s; y /= s; return *this; }

    // --- binary arithmetic ----------------------------------------------------------------------
    constexpr Vec2 operator+(Vec2 o) const { return Vec2(x + o.x, y + o.y); }
    constexpr Vec2 operator-(Vec2 o) const { return Vec2(x - o.x, y - o.y); }
    constexpr Vec2 operator*(Vec2 o) const { return Vec2(x * o.x, y * o.y); }
    constexpr Vec2 operator/(Vec2 o) const { return Vec2(x / o.x, y / o.y); }
    constexpr Vec2 operator+(float s) const { return Vec2(x + s, y + s); }
    constexpr Vec2 operator-(float s) const { return Vec2(x - s, y - s); }
    constexpr Vec2 operator*(float s) const { return Vec2(x * s, y * s); }
    constexpr Vec2 operator/(float s) const { return Vec2(x / s, y / s); }
    constexpr Vec2 operator-() const { return Vec2(-x, -y); }

    friend constexpr Vec2 operator+(float s, Vec2 v) { return Vec2(v.x + s, v.y + s); }
    friend constexpr Vec2 operator-(float s, Vec2 v) { return Vec2(s - v.x, s - v.y); }
    friend constexpr Vec2 operator*(float s, Vec2 v) { return Vec2(v.x * s, v.y * s); }
    friend constexpr Vec2 operator/(float s, Vec2 v) { return Vec2(s / v.x, s / v.y); }

    constexpr float operator[](int index) const
    {
        C2D_DEBUG_ASSERT(index == 0 || index == 1);
        return index == 0 ? x : y;
    }

    // --- comparison -----------------------------------------------------------------------------
    constexpr bool operator==(Vec2 o) const { return float_equals(x, o.x) && float_equals(y, o.y); }
    constexpr bool operator!=(Vec2 o) const { return !(*this == o); }
    constexpr bool operator<(Vec2 o) const { return x < o.x || (!(x > o.x) && y < o.y); }
    constexpr bool operator>(Vec2 o) const { return x > o.x || (!(x < o.x) && y > o.y); }

    std::string toString() const { return std::format("Vec2({}, {})", x, y); }

    friend std::ostream& operator<<(std::ostream& os, const Vec2& v) { return os << v.toString(); }
};

static_assert(std::is_trivially_copyable_v<Vec2> && sizeof(Vec2) == 8, "Vec2 is two packed floats");

} // namespace c2d

template <>
struct std::formatter<c2d::Vec2> : std::formatter<std::string>
{
    template <typename FormatContext>
    auto format(c2d::Vec2 v, FormatContext& ctx) const
    {
        return std::formatter<std::string>::format(v.toString(), ctx);
    }
};
