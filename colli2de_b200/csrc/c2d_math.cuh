// c2d_math.cuh — binary32 vector / transform primitives for the sm_100a kernels.
//
// Every expression here restates, operation for operation, the reference's
// Vec2 / Rotation / Transform arithmetic (reference include/colli2de/Vec2.hpp,
// Transform.hpp:33-46,136-144) so that device results are bit-identical to the
// reference's Release build (-ffp-contract=off, CMakeLists.txt:17-20).  The
// translation unit MUST be compiled with -fmad=false (no FMA contraction),
// default -prec-div=true -prec-sqrt=true, and never with -use_fast_math.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace c2d_dev
{

struct V2
{
    float x, y;
};

// (tx, ty, sin, cos): the part of c2d::Transform that apply()/toLocal() read.
struct XF
{
    float tx, ty, s, c;
};

__device__ __forceinline__ V2 mk(float x, float y)
{
    V2 v;
    v.x = x;
    v.y = y;
    return v;
}
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator*(V2 a, float s) { return mk(a.x * s, a.y * s); }
__device__ __forceinline__ V2 operator/(V2 a, float s) { return mk(a.x / s, a.y / s); }
__device__ __forceinline__ V2 operator-(V2 a) { return mk(-a.x, -a.y); }
__device__ __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float lensq(V2 a) { return a.x * a.x + a.y * a.y; }
// Vec2::normalize (Vec2.hpp:103-106,123-127): reciprocal of the length, then two multiplies.
__device__ __forceinline__ V2 normalize(V2 a)
{
    const float k = 1.0f / sqrtf(a.x * a.x + a.y * a.y);
    return mk(a.x * k, a.y * k);
}
// Vec2::operator< (Vec2.hpp:288-295)
__device__ __forceinline__ bool lex_less(V2 a, V2 b) { return a.x < b.x || (!(a.x > b.x) && a.y < b.y); }

// std::min / std::max / std::clamp semantics (NOT fminf/fmaxf): the first argument wins ties.
__device__ __forceinline__ float cmin(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float cmax(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float cclamp(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

// Rotation::apply / Rotation::inverse(Vec2) (Transform.hpp:33-36,43-46)
__device__ __forceinline__ V2 rot(const XF& t, V2 v) { return mk(t.c * v.x - t.s * v.y, t.s * v.x + t.c * v.y); }
__device__ __forceinline__ V2 rot_inv(const XF& t, V2 v) { return mk(t.c * v.x + t.s * v.y, -t.s * v.x + t.c * v.y); }
// Transform::apply / Transform::toLocal (Transform.hpp:136-144)
__device__ __forceinline__ V2 apply(const XF& t, V2 v)
{
    const V2 r = rot(t, v);
    return mk(r.x + t.tx, r.y + t.ty);
}
__device__ __forceinline__ V2 to_local(const XF& t, V2 v) { return rot_inv(t, mk(v.x - t.tx, v.y - t.ty)); }

__device__ __forceinline__ XF xf_identity()
{
    XF t;
    t.tx = 0.0f;
    t.ty = 0.0f;
    t.s = 0.0f;
    t.c = 1.0f;
    return t;
}

// AABB as (minx, miny, maxx, maxy)
struct Box
{
    float minx, miny, maxx, maxy;
};
__device__ __forceinline__ Box box_from(float4 f)
{
    Box b;
    b.minx = f.x;
    b.miny = f.y;
    b.maxx = f.z;
    b.maxy = f.w;
    return b;
}
__device__ __forceinline__ float4 box_to(const Box& b) { return make_float4(b.minx, b.miny, b.maxx, b.maxy); }
// AABB::intersects(AABB) — all comparisons non-strict (AABB.cpp:34-37)
__device__ __forceinline__ bool box_overlap(const float4& a, const float4& b)
{
    return a.x <= b.z && a.z >= b.x && a.y <= b.w && a.w >= b.y;
}
// AABB::contains(AABB) (AABB.cpp:29-32)
__device__ __forceinline__ bool box_contains(const Box& a, const Box& o)
{
    return a.minx <= o.minx && a.miny <= o.miny && a.maxx >= o.maxx && a.maxy >= o.maxy;
}
// AABB::combine (AABB.cpp:273-281)
__device__ __forceinline__ Box box_combine(const Box& a, const Box& b)
{
    Box r;
    r.minx = cmin(a.minx, b.minx);
    r.miny = cmin(a.miny, b.miny);
    r.maxx = cmax(a.maxx, b.maxx);
    r.maxy = cmax(a.maxy, b.maxy);
    return r;
}

} // namespace c2d_dev
