// c2d_narrow.cuh — device narrowphase: manifolds, boolean overlap tests and
// bullet sweeps for circle / capsule / segment / convex polygon (<= 8 verts).
//
// Restates the reference narrowphase with the same operation order so results
// are bit-identical under -fmad=false:
//   collide()       reference src/collision/Collision.cpp:15-517
//   areColliding()  reference src/collision/Collision.cpp:519-748
//   sweep()         reference include/colli2de/internal/collision/Collision.hpp:71-124,151-213,246-307
// There is NO GJK/EPA in the reference: circle/capsule/segment
// pairs are closed-form closest-point tests, polygon pairs are SAT over face
// normals followed by Sutherland-Hodgman clipping, polygon-circle is
// closest-edge.  Quirks kept on purpose: capsule/segment vs polygon ignores the
// capsule radius (D4); slack is per pair type (D14).
#pragma once

#include "c2d_math.cuh"

#include <float.h>

namespace c2d_dev
{

constexpr float kEps = 1e-6f; // kEpsilon, Collision.cpp:12
constexpr int kMaxVerts = 8;  // MAX_POLYGON_VERTICES, Constants.hpp:9
constexpr int kClipCap = 32;  // scratch slots for the clipped polygon (8 verts + <= 1 per plane for convex input)

// c2d::Manifold (Manifold.hpp:35-69), 68 bytes when stored.
struct Manifold
{
    V2 normal;
    V2 point[2];
    V2 anchorA[2];
    V2 anchorB[2];
    float sep[2];
    int count;
};

__device__ __forceinline__ void manifold_clear(Manifold& m)
{
    m.normal = mk(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < 2; ++i)
    {
        m.point[i] = mk(0.0f, 0.0f);
        m.anchorA[i] = mk(0.0f, 0.0f);
        m.anchorB[i] = mk(0.0f, 0.0f);
        m.sep[i] = 0.0f;
    }
    m.count = 0;
}

// Manifold::reverse (Manifold.hpp:43-50)
__device__ __forceinline__ void manifold_reverse(Manifold& m)
{
    m.normal = -m.normal;
    for (int i = 0; i < m.count; ++i)
    {
        const V2 t = m.anchorA[i];
        m.anchorA[i] = m.anchorB[i];
        m.anchorB[i] = t;
    }
}

// 17 words: normal, 2 x (point, anchorA, anchorB, separation), pointCount (low byte; padding zeroed)
__device__ __forceinline__ void manifold_store(const Manifold& m, float* out17)
{
    out17[0] = m.normal.x;
    out17[1] = m.normal.y;
#pragma unroll
    for (int i = 0; i < 2; ++i)
    {
        float* p = out17 + 2 + 7 * i;
        p[0] = m.point[i].x;
        p[1] = m.point[i].y;
        p[2] = m.anchorA[i].x;
        p[3] = m.anchorA[i].y;
        p[4] = m.anchorB[i].x;
        p[5] = m.anchorB[i].y;
        p[6] = m.sep[i];
    }
    reinterpret_cast<uint32_t*>(out17)[16] = static_cast<uint32_t>(m.count) & 0xffu;
}

__device__ __forceinline__ void set_single_contact(Manifold& m, V2 normal, V2 world, const XF& xa, const XF& xb,
                                                   float separation)
{
    m.normal = normal;
    m.point[0] = world;
    m.anchorA[0] = to_local(xa, world);
    m.anchorB[0] = to_local(xb, world);
    m.sep[0] = separation;
    m.count = 1;
}

// ---------------------------------------------------------------------------------------------
// circle - circle (Collision.cpp:15-49 / 519-526)
// ---------------------------------------------------------------------------------------------
template <bool kManifold>
__device__ __forceinline__ bool circle_circle(V2 ca, float ra, const XF& xa, V2 cb, float rb, const XF& xb,
                                              Manifold& m)
{
    const V2 centerA = apply(xa, ca);
    const V2 centerB = apply(xb, cb);
    const V2 delta = centerB - centerA;
    const float distSq = lensq(delta);
    const float radius = ra + rb;
    if (distSq > (radius + kEps) * (radius + kEps))
        return false;
    if (!kManifold)
        return true;
    const float distance = sqrtf(distSq);
    const V2 normal = (distance > kEps) ? (delta * (1.0f / distance)) : mk(1.0f, 0.0f);
    const V2 contactA = centerA + normal * ra;
    const V2 contactB = centerB - normal * rb;
    const V2 world = (contactA + contactB) * 0.5f;
    set_single_contact(m, normal, world, xa, xb, distance - radius);
    return true;
}

// ---------------------------------------------------------------------------------------------
// capsule (A) - circle (B) (Collision.cpp:51-100 / 528-550); segments are capsules of radius 0
// ---------------------------------------------------------------------------------------------
template <bool kManifold>
__device__ __forceinline__ bool capsule_circle(V2 c1, V2 c2, float rcap, const XF& xa, V2 cc, float rc,
                                               const XF& xb, Manifold& m)
{
    const V2 center1 = apply(xa, c1);
    const V2 center2 = apply(xa, c2);
    const V2 centerCircle = apply(xb, cc);

    const V2 segment = center2 - center1;
    const float segLenSq = lensq(segment);
    float projection = 0.0f;
    if (segLenSq > 0.0f)
    {
        projection = dot(centerCircle - center1, segment) / segLenSq;
        projection = cclamp(projection, 0.0f, 1.0f);
    }
    const V2 closest = center1 + segment * projection;
    const V2 delta = centerCircle - closest;
    const float distSq = lensq(delta);
    const float sumRadius = rcap + rc;
    if (distSq > (sumRadius + kEps) * (sumRadius + kEps))
        return false;
    if (!kManifold)
        return true;

    float distance = 0.0f;
    V2 normal = mk(1.0f, 0.0f);
    if (distSq > kEps * kEps)
    {
        distance = sqrtf(distSq);
        normal = delta * (1.0f / distance);
    }
    const V2 contactA = closest + normal * rcap;
    const V2 contactB = centerCircle - normal * rc;
    const V2 world = (contactA + contactB) * 0.5f;
    set_single_contact(m, normal, world, xa, xb, distance - sumRadius);
    return true;
}

// ---------------------------------------------------------------------------------------------
// capsule - capsule (Collision.cpp:110-208 / 557-635): closest points of two segments
// ---------------------------------------------------------------------------------------------
template <bool kManifold>
__device__ __forceinline__ bool capsule_capsule(V2 a1, V2 a2, float ra, const XF& xa, V2 b1, V2 b2, float rb,
                                                const XF& xb, Manifold& m)
{
    const V2 center1A = apply(xa, a1);
    const V2 center2A = apply(xa, a2);
    const V2 center1B = apply(xb, b1);
    const V2 center2B = apply(xb, b2);

    const V2 segA = center2A - center1A;
    const V2 segB = center2B - center1B;
    const V2 startDelta = center1A - center1B;

    const float lenSqA = dot(segA, segA);
    const float lenSqB = dot(segB, segB);
    const float deltaDotB = dot(segB, startDelta);

    float paramA = 0.0f;
    float paramB = 0.0f;

    if (lenSqA <= kEps && lenSqB <= kEps)
    {
        paramA = paramB = 0.0f;
    }
    else if (lenSqA <= kEps)
    {
        paramA = 0.0f;
        paramB = cclamp(deltaDotB / lenSqB, 0.0f, 1.0f);
    }
    else
    {
        const float deltaDotA = dot(segA, startDelta);
        if (lenSqB <= kEps)
        {
            paramB = 0.0f;
            paramA = cclamp(-deltaDotA / lenSqA, 0.0f, 1.0f);
        }
        else
        {
            const float dotAB = dot(segA, segB);
            const float denominator = lenSqA * lenSqB - dotAB * dotAB;
            if (denominator != 0.0f)
                paramA = cclamp((dotAB * deltaDotB - deltaDotA * lenSqB) / denominator, 0.0f, 1.0f);
            else
                paramA = 0.0f;
            const float tNumerator = dotAB * paramA + deltaDotB;
            if (tNumerator < 0.0f)
            {
                paramB = 0.0f;
                paramA = cclamp(-deltaDotA / lenSqA, 0.0f, 1.0f);
            }
            else if (tNumerator > lenSqB)
            {
                paramB = 1.0f;
                paramA = cclamp((dotAB - deltaDotA) / lenSqA, 0.0f, 1.0f);
            }
            else
            {
                paramB = tNumerator / lenSqB;
            }
        }
    }

    const V2 pointA = center1A + segA * paramA;
    const V2 pointB = center1B + segB * paramB;
    const V2 delta = pointB - pointA;
    const float distSq = lensq(delta);
    const float radius = ra + rb;
    if (distSq > (radius + kEps) * (radius + kEps))
        return false;
    if (!kManifold)
        return true;

    float distance = 0.0f;
    V2 normal = mk(1.0f, 0.0f);
    if (distSq > kEps * kEps)
    {
        distance = sqrtf(distSq);
        normal = delta * (1.0f / distance);
    }
    const V2 contactA = pointA + normal * ra;
    const V2 contactB = pointB - normal * rb;
    const V2 world = (contactA + contactB) * 0.5f;
    set_single_contact(m, normal, world, xa, xb, distance - radius);
    return true;
}

// ---------------------------------------------------------------------------------------------
// polygons in world space
// ---------------------------------------------------------------------------------------------
struct PolyW
{
    int n;
    V2 v[kMaxVerts];   // transform.apply(vertices[i])
    V2 nrm[kMaxVerts]; // transform.rotation.apply(normals[i])
};

// g: polygon geometry slot (verts in [0,16), normals in [16,32))
__device__ __forceinline__ void poly_world(const float* __restrict__ g, int count, const XF& t, PolyW& p)
{
    p.n = count;
    for (int i = 0; i < count; ++i)
    {
        p.v[i] = apply(t, mk(g[2 * i], g[2 * i + 1]));
        p.nrm[i] = rot(t, mk(g[16 + 2 * i], g[16 + 2 * i + 1]));
    }
}

// Capsule / segment as the 2-gon of Collision.cpp:483-496: vertices {c1, c2}, local normals {n, -n} with
// axis = (c2 - c1).normalize(), n = (-axis.y, axis.x); the radius is dropped.
__device__ __forceinline__ void capsule_poly_world(V2 c1, V2 c2, const XF& t, PolyW& p)
{
    const V2 axis = normalize(c2 - c1);
    const V2 n = mk(-axis.y, axis.x);
    p.n = 2;
    p.v[0] = apply(t, c1);
    p.v[1] = apply(t, c2);
    p.nrm[0] = rot(t, n);
    p.nrm[1] = rot(t, -n);
}

// projectPolygon (Collision.cpp:249-261)
__device__ __forceinline__ void project(const PolyW& p, V2 axis, float& outMin, float& outMax)
{
    outMin = dot(axis, p.v[0]);
    outMax = outMin;
    for (int i = 1; i < p.n; ++i)
    {
        const float proj = dot(axis, p.v[i]);
        outMin = cmin(outMin, proj);
        outMax = cmax(outMax, proj);
    }
}

// One half of the SAT loop (areAxesOverlapping, Collision.cpp:327-345): axes of `owner`.
__device__ __forceinline__ bool sat_axes(const PolyW& owner, const PolyW& a, const PolyW& b, float& overlap,
                                         V2& smallestAxis)
{
    for (int i = 0; i < owner.n; ++i)
    {
        const V2 axis = owner.nrm[i];
        float minA, maxA, minB, maxB;
        project(a, axis, minA, maxA);
        project(b, axis, minB, maxB);
        const float axisOverlap = cmin(maxA, maxB) - cmax(minA, minB);
        if (axisOverlap < 0.0f)
            return false;
        if (axisOverlap < overlap)
        {
            overlap = axisOverlap;
            smallestAxis = axis;
        }
    }
    return true;
}

// clipWithPlane (Collision.cpp:274-306): Sutherland-Hodgman keeping dist <= 0.  Returns the new count.
__device__ __forceinline__ int clip_plane(const V2* in, int nin, V2* out, V2 point, V2 normal, bool& overflow)
{
    if (nin == 0)
        return 0;
    int nout = 0;
    V2 prev = in[nin - 1];
    float previousDistance = dot(normal, prev - point);
    for (int k = 0; k < nin; ++k)
    {
        const V2 curr = in[k];
        const float currentDistance = dot(normal, curr - point);
        if (currentDistance <= 0.0f)
        {
            if (previousDistance > 0.0f)
            {
                const float f = previousDistance / (previousDistance - currentDistance);
                if (nout < kClipCap)
                    out[nout++] = prev + (curr - prev) * f;
                else
                    overflow = true;
            }
            if (nout < kClipCap)
                out[nout++] = curr;
            else
                overflow = true;
        }
        else if (previousDistance <= 0.0f)
        {
            const float f = previousDistance / (previousDistance - currentDistance);
            if (nout < kClipCap)
                out[nout++] = prev + (curr - prev) * f;
            else
                overflow = true;
        }
        prev = curr;
        previousDistance = currentDistance;
    }
    return nout;
}

// polygon - polygon (Collision.cpp:308-399 / 668-690)
template <bool kManifold>
__device__ __forceinline__ bool poly_poly(const PolyW& a, const XF& xa, const PolyW& b, const XF& xb, Manifold& m)
{
    float overlap = FLT_MAX;
    V2 smallestAxis = mk(1.0f, 0.0f);
    if (!sat_axes(a, a, b, overlap, smallestAxis))
        return false;
    if (!sat_axes(b, a, b, overlap, smallestAxis))
        return false;
    if (!kManifold)
        return true;

    V2 avgA = a.v[0];
    for (int i = 1; i < a.n; ++i)
        avgA = avgA + a.v[i];
    avgA = avgA / static_cast<float>(a.n);
    V2 avgB = b.v[0];
    for (int i = 1; i < b.n; ++i)
        avgB = avgB + b.v[i];
    avgB = avgB / static_cast<float>(b.n);

    if (dot(avgB - avgA, smallestAxis) < 0.0f)
        smallestAxis = -smallestAxis;

    // polygonIntersection (Collision.cpp:308-322): A's world vertices clipped by every plane of B
    V2 bufA[kClipCap], bufB[kClipCap];
    V2* cur = bufA;
    V2* nxt = bufB;
    int n = a.n;
    for (int i = 0; i < a.n; ++i)
        cur[i] = a.v[i];
    bool overflow = false;
    for (int i = 0; i < b.n && n > 0; ++i)
    {
        n = clip_plane(cur, n, nxt, b.v[i], b.nrm[i], overflow);
        V2* t = cur;
        cur = nxt;
        nxt = t;
    }

    m.normal = smallestAxis;
    if (n == 0)
    {
        const V2 world = (avgA + avgB) * 0.5f;
        set_single_contact(m, smallestAxis, world, xa, xb, -overlap);
        return true;
    }

    // std::sort by Vec2::operator< then the first min(n, 2) points (Collision.cpp:385-396):
    // only the two lexicographically smallest points are needed.
    int i0 = 0;
    for (int k = 1; k < n; ++k)
        if (lex_less(cur[k], cur[i0]))
            i0 = k;
    const int count = n < 2 ? n : 2;
    int i1 = -1;
    if (count == 2)
    {
        for (int k = 0; k < n; ++k)
        {
            if (k == i0)
                continue;
            if (i1 < 0 || lex_less(cur[k], cur[i1]))
                i1 = k;
        }
    }
    const int idx[2] = {i0, i1};
    for (int k = 0; k < count; ++k)
    {
        const V2 c = cur[idx[k]];
        m.point[k] = c;
        m.anchorA[k] = to_local(xa, c);
        m.anchorB[k] = to_local(xb, c);
        m.sep[k] = -overlap;
    }
    m.count = count;
    return true;
}

// polygon (A) - circle (B) (Collision.cpp:402-473 / 692-721).
// Single pass over the edges straight from the geometry slot (no world-vertex array in local memory): edge i uses
// v1 = apply(v[i]) and v2 = apply(v[(i + 1) % n]); v2 of one edge is v1 of the next, apply() is deterministic, so the
// values are those of the reference.  isPointInsidePolygon (Collision.cpp:402-412) is folded into the same loop.
template <bool kManifold>
__device__ __forceinline__ bool poly_circle(const float* __restrict__ g, int n, const XF& xa, V2 cc, float radius,
                                            const XF& xb, Manifold& m)
{
    const V2 center = apply(xb, cc);
    bool isInside = true;
    bool touching = false;
    float minDistSq = FLT_MAX;
    V2 closest = mk(0.0f, 0.0f);
    V2 normal = mk(1.0f, 0.0f);
    const V2 first = n > 0 ? apply(xa, mk(g[0], g[1])) : mk(0.0f, 0.0f);
    V2 v1 = first;
    for (int i = 0; i < n; ++i)
    {
        const V2 v2 = (i + 1 < n) ? apply(xa, mk(g[2 * i + 2], g[2 * i + 3])) : first;
        const V2 edgeNormal = rot(xa, mk(g[16 + 2 * i], g[16 + 2 * i + 1]));
        if (dot(edgeNormal, center - v1) > 0.0f)
            isInside = false;
        const V2 edge = v2 - v1;
        const float edgeLenSq = lensq(edge);
        float edgeParam = 0.0f;
        if (edgeLenSq > 0.0f)
            edgeParam = cclamp(dot(center - v1, edge) / edgeLenSq, 0.0f, 1.0f);
        const V2 pointOnEdge = v1 + edge * edgeParam;
        const V2 delta = center - pointOnEdge;
        const float distSq = lensq(delta);
        if (!kManifold)
        {
            if (distSq <= radius * radius)
                touching = true;
        }
        else if (distSq < minDistSq)
        {
            minDistSq = distSq;
            closest = pointOnEdge;
            if (edgeParam > 0.0f && edgeParam < 1.0f)
                normal = edgeNormal;
            else
                normal = lensq(delta) > kEps * kEps ? normalize(delta) : edgeNormal;
        }
        v1 = v2;
    }
    if (!kManifold)
        return isInside || touching;

    if (!isInside && minDistSq > radius * radius)
        return false;

    const float distance = sqrtf(minDistSq);
    if (!isInside && distance > kEps)
        normal = (center - closest) * (1.0f / distance);
    else if (isInside)
        normal = -normal;

    const V2 contactA = closest;
    const V2 contactB = center - normal * radius;
    const V2 world = (contactA + contactB) * 0.5f;
    set_single_contact(m, normal, world, xa, xb, isInside ? -(radius + distance) : (distance - radius));
    return true;
}

// ---------------------------------------------------------------------------------------------
// type dispatch: the 16 overloads of collide()/areColliding() reduced to the five cores above
// ---------------------------------------------------------------------------------------------
enum : int
{
    T_CIRCLE = 0,
    T_CAPSULE = 1,
    T_SEGMENT = 2,
    T_POLYGON = 3
};

// Narrowphase "core" of a type pair (bin key of the candidate sort)
enum : int
{
    CORE_CC = 0,     // circle-circle
    CORE_CAPC = 1,   // capsule|segment vs circle (either order)
    CORE_CAPCAP = 2, // capsule|segment vs capsule|segment
    CORE_PC = 3,     // polygon vs circle (either order)
    CORE_PP = 4,     // polygon vs polygon|capsule|segment (either order)
    CORE_COUNT = 5
};

__device__ __forceinline__ int core_of(int ta, int tb)
{
    const bool ca = ta == T_CIRCLE, cb = tb == T_CIRCLE;
    const bool pa = ta == T_POLYGON, pb = tb == T_POLYGON;
    if (ca && cb)
        return CORE_CC;
    if (pa || pb)
        return (ca || cb) ? CORE_PC : CORE_PP;
    if (ca || cb)
        return CORE_CAPC;
    return CORE_CAPCAP;
}

struct ShapeRef
{
    int type;
    int count;                 // polygon vertex count
    const float* __restrict__ g; // geometry slot (32 floats)
};

__device__ __forceinline__ V2 g2(const float* g, int i) { return mk(g[2 * i], g[2 * i + 1]); }
// capsule radius; a segment is Capsule{start, end, 0} (Collision.cpp:211-246)
__device__ __forceinline__ float cap_radius(const ShapeRef& s) { return s.type == T_CAPSULE ? s.g[4] : 0.0f; }

__device__ __forceinline__ void shape_poly(const ShapeRef& s, const XF& t, PolyW& p)
{
    if (s.type == T_POLYGON)
        poly_world(s.g, s.count, t, p);
    else
        capsule_poly_world(g2(s.g, 0), g2(s.g, 1), t, p);
}

// collide(A, xa, B, xb) / areColliding(A, xa, B, xb).  kManifold=false never touches `m`.
//
// The 16 overloads reduce to five cores.  Overloads that the reference implements as "swap the arguments and
// Manifold::reverse()" are brought into the core's argument order FIRST and reversed at the end, so each core
// is instantiated once and a warp of one bin never splits over two copies of the same code:
//   circle-X            -> core(X, xb, circle, xa).reverse()              (Collision.cpp:102-107,218-224,475-480)
//   polygon-capsule|seg -> core(capsule, xb, polygon, xa).reverse()       (Collision.cpp:498-510)
//   segment-capsule     -> core(capsule, xb, segCap, xa).reverse()        (Collision.cpp:240-246)
//   segment-polygon     -> reverse(reverse(core(segCap, xa, polygon, xb))) = core(...)  (Collision.cpp:512-517)
template <bool kManifold>
__device__ inline bool narrow_test(const ShapeRef& A, const XF& xa, const ShapeRef& B, const XF& xb, Manifold& m)
{
    if (kManifold)
        manifold_clear(m);
    const bool rev = (A.type == T_CIRCLE && B.type != T_CIRCLE) ||
                     (A.type == T_POLYGON && (B.type == T_CAPSULE || B.type == T_SEGMENT)) ||
                     (A.type == T_SEGMENT && B.type == T_CAPSULE);
    const ShapeRef& S1 = rev ? B : A;
    const ShapeRef& S2 = rev ? A : B;
    const XF& x1 = rev ? xb : xa;
    const XF& x2 = rev ? xa : xb;

    bool hit;
    if (S1.type == T_CIRCLE)
    {
        hit = circle_circle<kManifold>(g2(S1.g, 0), S1.g[2], x1, g2(S2.g, 0), S2.g[2], x2, m);
    }
    else if (S2.type == T_CIRCLE)
    {
        if (S1.type == T_POLYGON)
            hit = poly_circle<kManifold>(S1.g, S1.count, x1, g2(S2.g, 0), S2.g[2], x2, m);
        else
            hit = capsule_circle<kManifold>(g2(S1.g, 0), g2(S1.g, 1), cap_radius(S1), x1, g2(S2.g, 0), S2.g[2], x2, m);
    }
    else if (S2.type == T_POLYGON)
    {
        PolyW p1, p2;
        shape_poly(S1, x1, p1);
        poly_world(S2.g, S2.count, x2, p2);
        hit = poly_poly<kManifold>(p1, x1, p2, x2, m);
    }
    else
    {
        hit = capsule_capsule<kManifold>(g2(S1.g, 0), g2(S1.g, 1), cap_radius(S1), x1, g2(S2.g, 0), g2(S2.g, 1),
                                         cap_radius(S2), x2, m);
    }
    if (kManifold && rev && hit)
        manifold_reverse(m);
    return hit;
}

// computeAABB(shape, transform) (reference internal/geometry/ShapesUtils.hpp:13-60)
__device__ __forceinline__ Box compute_aabb(uint32_t type, uint32_t count, const float* __restrict__ g, const XF& t)
{
    Box b;
    if (type == T_CIRCLE)
    {
        const V2 c = apply(t, mk(g[0], g[1]));
        const float r = g[2];
        b.minx = c.x - r;
        b.miny = c.y - r;
        b.maxx = c.x + r;
        b.maxy = c.y + r;
    }
    else if (type == T_CAPSULE)
    {
        const V2 p1 = apply(t, mk(g[0], g[1]));
        const V2 p2 = apply(t, mk(g[2], g[3]));
        const float r = g[4];
        b.minx = cmin(p1.x, p2.x) - r;
        b.miny = cmin(p1.y, p2.y) - r;
        b.maxx = cmax(p1.x, p2.x) + r;
        b.maxy = cmax(p1.y, p2.y) + r;
    }
    else if (type == T_SEGMENT)
    {
        const V2 p1 = apply(t, mk(g[0], g[1]));
        const V2 p2 = apply(t, mk(g[2], g[3]));
        b.minx = cmin(p1.x, p2.x);
        b.miny = cmin(p1.y, p2.y);
        b.maxx = cmax(p1.x, p2.x);
        b.maxy = cmax(p1.y, p2.y);
    }
    else
    {
        const V2 v0 = apply(t, mk(g[0], g[1]));
        b.minx = b.maxx = v0.x;
        b.miny = b.maxy = v0.y;
        for (uint32_t i = 1; i < count; ++i)
        {
            const V2 v = apply(t, mk(g[2 * i], g[2 * i + 1]));
            b.minx = cmin(b.minx, v.x);
            b.miny = cmin(b.miny, v.y);
            b.maxx = cmax(b.maxx, v.x);
            b.maxy = cmax(b.maxy, v.y);
        }
    }
    return b;
}


// ---------------------------------------------------------------------------------------------
// sweeps (Collision.hpp:71-124 one moving, 151-213 two moving, 246-307 sweep())
// ---------------------------------------------------------------------------------------------
// A moving body's motion: start/end translation + angle.  The sampled transform is
// Transform(start.t + delta.t * f, start.angle + delta.angle * f) whose sin/cos come from sinf/cosf of the
// interpolated angle.  When delta.angle == 0 the interpolated angle IS start.angle, so the host-supplied
// libm values (csin, ccos) are used and the result is bit-exact; otherwise sin/cos are evaluated in
// double precision and rounded (agrees with glibc's < 1 ulp sinf/cosf except on rare rounding ties).
struct Motion
{
    XF start;         // the start transform as stored (tx, ty, sin, cos); for a fixed body: its transform
    float sa;         // start angle
    float dx, dy, da; // end - start (translation, angle)
    float csin, ccos; // sinf(sa), cosf(sa) from the host's libm
    bool moving;
};

__device__ __forceinline__ Motion motion_fixed(const XF& t)
{
    Motion mo;
    mo.start = t;
    mo.sa = 0.0f;
    mo.dx = mo.dy = mo.da = 0.0f;
    mo.csin = t.s;
    mo.ccos = t.c;
    mo.moving = false;
    return mo;
}

// start = (tx, ty, sin, cos) + angle as stored for the previous transform, end likewise for the current one
__device__ __forceinline__ Motion motion_between(const XF& start, float startAngle, float csin, float ccos,
                                                 const XF& end, float endAngle)
{
    Motion mo;
    mo.start = start;
    mo.sa = startAngle;
    mo.dx = end.tx - start.tx;
    mo.dy = end.ty - start.ty;
    mo.da = endAngle - startAngle;
    mo.csin = csin;
    mo.ccos = ccos;
    mo.moving = true;
    return mo;
}

__device__ __forceinline__ XF motion_at(const Motion& mo, float f)
{
    if (!mo.moving)
        return mo.start;
    XF t;
    t.tx = mo.start.tx + mo.dx * f;
    t.ty = mo.start.ty + mo.dy * f;
    if (mo.da == 0.0f)
    {
        t.s = mo.csin;
        t.c = mo.ccos;
    }
    else
    {
        const float ang = mo.sa + mo.da * f;
        t.s = static_cast<float>(sin(static_cast<double>(ang)));
        t.c = static_cast<float>(cos(static_cast<double>(ang)));
    }
    return t;
}

// sweepHelper (Collision.hpp:71-124 / 151-213).  Returns true on hit with `fraction` = fractionUpper.
// The t=0 test uses the start transforms AS STORED (Collision.hpp:75-76); every later sample rebuilds sin/cos
// from the interpolated angle (Collision.hpp:85-99).
__device__ inline bool sweep_search(const ShapeRef& A, const Motion& ma, const ShapeRef& B, const Motion& mb,
                                    float& fraction)
{
    Manifold dummy;
    if (narrow_test<false>(A, ma.start, B, mb.start, dummy))
    {
        fraction = 0.0f;
        return true;
    }
    float lower = 0.0f, upper = 1.0f;
    for (int step = 1; step <= 8; ++step)
    {
        const float f = static_cast<float>(step) / 8.0f;
        if (narrow_test<false>(A, motion_at(ma, f), B, motion_at(mb, f), dummy))
        {
            lower = static_cast<float>(step - 1) / 8.0f;
            upper = f;
            break;
        }
    }
    if (lower == 0.0f && upper == 1.0f)
        return false;
    for (int it = 0; it < 10; ++it)
    {
        const float mid = 0.5f * (lower + upper);
        if (narrow_test<false>(A, motion_at(ma, mid), B, motion_at(mb, mid), dummy))
            upper = mid;
        else
            lower = mid;
        if (upper - lower < 1e-6f)
            break;
    }
    fraction = upper;
    return true;
}

// The manifold of a sweep hit: collide() at the transforms rebuilt from the hit fraction — also for fraction 0
// (Collision.hpp:260-264, 296-306).
__device__ inline void sweep_manifold(const ShapeRef& A, const Motion& ma, const ShapeRef& B, const Motion& mb,
                                      float fraction, Manifold& m)
{
    narrow_test<true>(A, motion_at(ma, fraction), B, motion_at(mb, fraction), m);
}

__device__ inline bool narrow_sweep(const ShapeRef& A, const Motion& ma, const ShapeRef& B, const Motion& mb,
                                    float& fraction, Manifold& m)
{
    if (!sweep_search(A, ma, B, mb, fraction))
        return false;
    sweep_manifold(A, ma, B, mb, fraction, m);
    return true;
}

} // namespace c2d_dev
