// c2d_ray.cuh — device ray tests: ray vs AABB (slab), ray vs circle / segment /
// polygon / capsule, finite and infinite, and the bullet ray-sweep.
//
// Restates, operation for operation:
//   AABB::intersects(Ray / InfiniteRay)       reference src/geometry/AABB.cpp:39-164
//   raycast(shape, transform, ray) x 8        reference src/geometry/RaycastShapes.cpp:11-320
//   sweep(shape, T0, T1, ray)                 reference internal/collision/Collision.hpp:323-344
// A ray is (start, d, tmax): finite rays have d = end - start and tmax = 1, infinite rays have
// d = direction and tmax = FLT_MAX ("infinite" also switches the few places where the two overloads
// differ in which bounds they test).
#pragma once

#include "c2d_narrow.cuh"

namespace c2d_dev
{

struct RayQ
{
    V2 start;
    V2 d;          // end - start (finite) or direction (infinite)
    bool infinite;
};

// AABB::intersects(Ray) / AABB::intersects(InfiniteRay): slab method, parallel threshold |d| < 1e-8.
__device__ __forceinline__ bool ray_box(const RayQ& r, float minx, float miny, float maxx, float maxy, float& t0,
                                        float& t1)
{
    float tmin = 0.0f;
    float tmax = r.infinite ? FLT_MAX : 1.0f;
    {
        const float dir = r.d.x;
        const float origin = r.start.x;
        if (fabsf(dir) < 1e-8f)
        {
            if (origin < minx || origin > maxx)
                return false;
        }
        else
        {
            const float invD = 1.0f / dir;
            float entry = (minx - origin) * invD;
            float exit = (maxx - origin) * invD;
            if (invD < 0.0f)
            {
                const float t = entry;
                entry = exit;
                exit = t;
            }
            tmin = cmax(tmin, entry);
            tmax = cmin(tmax, exit);
            if (tmax < tmin)
                return false;
        }
    }
    {
        const float dir = r.d.y;
        const float origin = r.start.y;
        if (fabsf(dir) < 1e-8f)
        {
            if (origin < miny || origin > maxy)
                return false;
        }
        else
        {
            const float invD = 1.0f / dir;
            float entry = (miny - origin) * invD;
            float exit = (maxy - origin) * invD;
            if (invD < 0.0f)
            {
                const float t = entry;
                entry = exit;
                exit = t;
            }
            tmin = cmax(tmin, entry);
            tmax = cmin(tmax, exit);
            if (tmax < tmin)
                return false;
        }
    }
    t0 = tmin;
    t1 = tmax;
    return true;
}

// raycast(Circle, transform, Ray|InfiniteRay) (RaycastShapes.cpp:11-88)
__device__ __forceinline__ bool ray_circle(V2 localCenter, float radius, const XF& t, const RayQ& r, float& t0,
                                           float& t1)
{
    const V2 center = apply(t, localCenter);
    const V2 fromStart = r.start - center;
    const float dirLenSq = dot(r.d, r.d);
    const float startProj = dot(fromStart, r.d);
    const float startDistSq = dot(fromStart, fromStart) - radius * radius;
    if (dirLenSq <= 0.0f)
    {
        if (startDistSq <= 0.0f)
        {
            t0 = 0.0f;
            t1 = 0.0f;
            return true;
        }
        return false;
    }
    if (startDistSq > 0.0f && startProj > 0.0f)
        return false;
    const float disc = startProj * startProj - dirLenSq * startDistSq;
    if (disc < 0.0f)
        return false;
    const float root = sqrtf(disc);
    float entry = (-startProj - root) / dirLenSq;
    float exit = (-startProj + root) / dirLenSq;
    if (entry > exit)
    {
        const float tt = entry;
        entry = exit;
        exit = tt;
    }
    if (r.infinite)
    {
        if (exit < 0.0f)
            return false;
        entry = cmax(entry, 0.0f);
    }
    else
    {
        if (entry > 1.0f || exit < 0.0f)
            return false;
        entry = cmax(entry, 0.0f);
        exit = cmin(exit, 1.0f);
        if (entry > exit)
            return false;
    }
    t0 = entry;
    t1 = exit;
    return true;
}

// raycast(Segment, ...) (RaycastShapes.cpp:90-134): returns (tRay, tRay)
__device__ __forceinline__ bool ray_segment(V2 s, V2 e, const XF& t, const RayQ& r, float& t0, float& t1)
{
    const V2 p1 = apply(t, s);
    const V2 p2 = apply(t, e);
    const V2 seg = p2 - p1;
    const V2 delta = r.start - p1;
    const float crossSegRay = cross(seg, r.d);
    const float crossDeltaSeg = cross(delta, seg);
    if (fabsf(crossSegRay) < 1e-8f)
        return false;
    const float tSegment = cross(delta, r.d) / crossSegRay;
    const float tRay = crossDeltaSeg / crossSegRay;
    const bool ok = tSegment >= 0.0f && tSegment <= 1.0f && tRay >= 0.0f && (r.infinite || tRay <= 1.0f);
    if (!ok)
        return false;
    t0 = tRay;
    t1 = tRay;
    return true;
}

// raycast(Polygon, ...) (RaycastShapes.cpp:136-206): half-plane clipping of the ray parameter interval
__device__ __forceinline__ bool ray_polygon(const float* __restrict__ g, int count, const XF& t, const RayQ& r,
                                            float& t0, float& t1)
{
    float lower = 0.0f;
    float upper = r.infinite ? FLT_MAX : 1.0f;
    for (int i = 0; i < count; ++i)
    {
        const V2 vertex = apply(t, mk(g[2 * i], g[2 * i + 1]));
        const V2 normal = rot(t, mk(g[16 + 2 * i], g[16 + 2 * i + 1]));
        const float numerator = dot(normal, vertex - r.start);
        const float denominator = dot(normal, r.d);
        if (fabsf(denominator) < 1e-8f)
        {
            if (numerator < 0.0f)
                return false;
            continue;
        }
        const float tt = numerator / denominator;
        if (denominator < 0.0f)
            lower = cmax(lower, tt);
        else
            upper = cmin(upper, tt);
        if (upper < lower)
            return false;
    }
    t0 = lower;
    t1 = upper;
    return true;
}

// raycast(Capsule, ...) (RaycastShapes.cpp:208-320): ray moved into the capsule frame, box core + two caps.
// exitTime starts at FLT_MIN - the smallest POSITIVE float, as in the reference (numeric_limits<float>::min()).
// The finite overload needs the ray's END POINT (toLocal(ray.end)), not start + d, to stay bit-exact,
// so the capsule test takes the original end point explicitly.
__device__ __forceinline__ bool ray_capsule(V2 c1, V2 c2, float radius, const XF& t, const RayQ& r, V2 rayEnd,
                                                 float& t0, float& t1)
{
    const V2 worldStart = apply(t, c1);
    const V2 worldEnd = apply(t, c2);
    const V2 axis = worldEnd - worldStart;
    const float length = sqrtf(axis.x * axis.x + axis.y * axis.y);
    const XF ident = xf_identity();
    if (length <= 1e-8f)
        return ray_circle(worldStart, radius, ident, r, t0, t1);

    const V2 axisUnit = axis * (1.0f / length);
    const V2 side = mk(-axisUnit.y, axisUnit.x);
    RayQ local;
    local.infinite = r.infinite;
    {
        const V2 diff = r.start - worldStart;
        local.start = mk(dot(diff, axisUnit), dot(diff, side));
    }
    if (r.infinite)
    {
        local.d = mk(dot(r.d, axisUnit), dot(r.d, side));
    }
    else
    {
        const V2 diff = rayEnd - worldStart;
        const V2 localEnd = mk(dot(diff, axisUnit), dot(diff, side));
        local.d = localEnd - local.start;
    }

    float entryTime = FLT_MAX;
    float exitTime = FLT_MIN;
    bool hit = false;
    float a, b;
    if (ray_box(local, 0.0f, -radius, length, radius, a, b))
    {
        hit = true;
        entryTime = cmin(entryTime, a);
        exitTime = cmax(exitTime, b);
    }
    if (ray_circle(mk(0.0f, 0.0f), radius, ident, local, a, b))
    {
        hit = true;
        entryTime = cmin(entryTime, a);
        exitTime = cmax(exitTime, b);
    }
    if (ray_circle(mk(length, 0.0f), radius, ident, local, a, b))
    {
        hit = true;
        entryTime = cmin(entryTime, a);
        exitTime = cmax(exitTime, b);
    }
    if (!hit)
        return false;
    if (r.infinite)
    {
        if (exitTime < 0.0f || entryTime > exitTime)
            return false;
        entryTime = cmax(entryTime, 0.0f);
    }
    else
    {
        if (exitTime < 0.0f || entryTime > 1.0f || exitTime < entryTime)
            return false;
        entryTime = cmax(entryTime, 0.0f);
        exitTime = cmin(exitTime, 1.0f);
        if (entryTime > exitTime)
            return false;
    }
    t0 = entryTime;
    t1 = exitTime;
    return true;
}

// raycast(shape, transform, ray) dispatch
__device__ __forceinline__ bool ray_shape(const ShapeRef& s, const XF& t, const RayQ& r, V2 rayEnd, float& t0,
                                          float& t1)
{
    switch (s.type)
    {
    case T_CIRCLE:
        return ray_circle(g2(s.g, 0), s.g[2], t, r, t0, t1);
    case T_CAPSULE:
        return ray_capsule(g2(s.g, 0), g2(s.g, 1), s.g[4], t, r, rayEnd, t0, t1);
    case T_SEGMENT:
        return ray_segment(g2(s.g, 0), g2(s.g, 1), t, r, t0, t1);
    default:
        return ray_polygon(s.g, s.count, t, r, t0, t1);
    }
}

// sweep(shape, T0, T1, ray) (Collision.hpp:323-344): runs sweepHelper with raycast as the predicate and
// returns the LAST successful raycast seen during the search (not the one at the returned fraction).
__device__ __forceinline__ bool ray_sweep(const ShapeRef& s, const Motion& mo, const RayQ& r, V2 rayEnd, float& t0,
                                          float& t1)
{
    bool any = false;
    float a, b;
    if (ray_shape(s, mo.start, r, rayEnd, a, b))
    {
        t0 = a;
        t1 = b;
        return true; // sweepHelper returns 0 immediately (Collision.hpp:75-76)
    }
    float lower = 0.0f, upper = 1.0f;
    for (int step = 1; step <= 8; ++step)
    {
        const float f = static_cast<float>(step) / 8.0f;
        if (ray_shape(s, motion_at(mo, f), r, rayEnd, a, b))
        {
            t0 = a;
            t1 = b;
            any = true;
            lower = static_cast<float>(step - 1) / 8.0f;
            upper = f;
            break;
        }
    }
    if (lower == 0.0f && upper == 1.0f)
        return any;
    for (int it = 0; it < 10; ++it)
    {
        const float mid = 0.5f * (lower + upper);
        if (ray_shape(s, motion_at(mo, mid), r, rayEnd, a, b))
        {
            t0 = a;
            t1 = b;
            any = true;
            upper = mid;
        }
        else
            lower = mid;
        if (upper - lower < 1e-6f)
            break;
    }
    return any;
}

} // namespace c2d_dev
