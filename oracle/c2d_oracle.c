/*
 * c2d_oracle.c — CPU restatement, in plain C, of Colli2De's per-frame collision query path.
 *
 * TEST INFRASTRUCTURE.  This file is the ORACLE the CUDA path is checked against; it is never shipped, linked
 * or called by the product (colli2de_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  It exports the flat rn_* interface of
 * tests/cpp/runner_api.h so the same scripted scenes replay on it, on the real reference (oracle/_ref) and on
 * the B200 build.
 *
 * PARITY IS PINNED: tests/test_oracle_golden.py checks this file against the known-answer vectors of the
 * reference's own tests (tests/collision/test_collision.cpp, tests/geometry/test_SweepShapes.cpp,
 * test_RaycastShapes.cpp, test_SweepRaycast.cpp, test_AABB.cpp, tests/registry/test_registry_basic.cpp,
 * test_registry_raycast.cpp) and tests/test_oracle_vs_reference.py checks it bit-for-bit against the
 * unmodified reference compiled into oracle/_ref on random inputs and scripted registry scenes.
 *
 * Every function cites the reference code it follows (paths are into the reference repository).  Arithmetic
 * is IEEE binary32 in the reference's operation order; build with -ffp-contract=off (reference
 * CMakeLists.txt:17-20).  std::min(a,b) = (b<a)?b:a, std::max(a,b) = (a<b)?b:a, std::clamp as in <algorithm>.
 *
 * What is restated differently on purpose (observable results are the same):
 *   - the broadphase is a sort-and-sweep over the leaf (fat) AABBs instead of the incremental DynamicBVH:
 *     the candidate set of DynamicBVH::findAllCollisions is exactly "leaf boxes intersect and
 *     (categoryA & categoryB) != 0" (DynamicBVH.hpp:944-977), and the leaf boxes follow the tree-independent
 *     rule of addProxy / moveProxy (DynamicBVH.hpp:352-365, 386-400) — SURVEY.md D12;
 *   - same-tree pairs (Dynamic-Dynamic, Bullet-Bullet) are emitted with shapeA < shapeB; the reference's
 *     orientation there depends on its tree topology;
 *   - rayCast dedupes AABB hits with equal (entry, exit) keeping the LOWEST shape id (the reference keeps
 *     whichever its traversal meets first);
 *   - PartitioningMethod::Grid yields the same unique pair set as None and is not restated
 *     separately.
 */
#define _POSIX_C_SOURCE 200809L
#include "runner_api.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------
 * math: Vec2.hpp, Transform.hpp
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float x, y; } v2;
typedef struct { float tx, ty, angle, s, c; } xf_t; /* Transform: translation, rotation{angle, sin, cos} */

static v2 V(float x, float y) { v2 r = {x, y}; return r; }
static v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static v2 vmul(v2 a, float s) { return V(a.x * s, a.y * s); }
static v2 vdiv(v2 a, float s) { return V(a.x / s, a.y / s); }
static v2 vneg(v2 a) { return V(-a.x, -a.y); }
static float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static float vcross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static float vlensq(v2 a) { return a.x * a.x + a.y * a.y; }
/* Vec2::normalize, Vec2.hpp:103-106,123-127 */
static v2 vnormalize(v2 a) { const float k = 1.0f / sqrtf(a.x * a.x + a.y * a.y); return V(a.x * k, a.y * k); }
/* Vec2::operator<, Vec2.hpp:288-295 */
static int vless(v2 a, v2 b) { if (a.x < b.x) return 1; if (a.x > b.x) return 0; return a.y < b.y; }
static float fmin_std(float a, float b) { return (b < a) ? b : a; }
static float fmax_std(float a, float b) { return (a < b) ? b : a; }
static float fclamp_std(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

/* Rotation::apply / inverse, Transform.hpp:33-36,43-46; Transform::apply / toLocal, Transform.hpp:136-144 */
static v2 xrot(const xf_t* t, v2 v) { return V(t->c * v.x - t->s * v.y, t->s * v.x + t->c * v.y); }
static v2 xrotinv(const xf_t* t, v2 v) { return V(t->c * v.x + t->s * v.y, -t->s * v.x + t->c * v.y); }
static v2 xapply(const xf_t* t, v2 v) { const v2 r = xrot(t, v); return V(r.x + t->tx, r.y + t->ty); }
static v2 xlocal(const xf_t* t, v2 v) { return xrotinv(t, V(v.x - t->tx, v.y - t->ty)); }
/* Transform(Vec2, float), Transform.hpp:23-28,129-131 */
static xf_t xf_make(float tx, float ty, float angle)
{
    xf_t t; t.tx = tx; t.ty = ty; t.angle = angle; t.s = sinf(angle); t.c = cosf(angle); return t;
}
static xf_t xf_raw(const float* f) { xf_t t; t.tx = f[0]; t.ty = f[1]; t.angle = f[2]; t.s = f[3]; t.c = f[4]; return t; }
static void xf_store(const xf_t* t, float* f) { f[0] = t->tx; f[1] = t->ty; f[2] = t->angle; f[3] = t->s; f[4] = t->c; }
static xf_t xf_identity(void) { xf_t t = {0.0f, 0.0f, 0.0f, 0.0f, 1.0f}; return t; }
/* Transform::operator+=(Transform), Transform.hpp:75-81,181-186: cos is computed from the UPDATED sin (D5) */
static void xf_add(xf_t* t, const xf_t* o)
{
    t->tx += o->tx; t->ty += o->ty;
    t->angle += o->angle;
    t->s = t->s * o->c + t->c * o->s;
    t->c = t->c * o->c - t->s * o->s;
}

/* ------------------------------------------------------------------------------------------------
 * shapes: Shapes.hpp
 * ---------------------------------------------------------------------------------------------- */
enum { CIRCLE = 0, CAPSULE = 1, SEGMENT = 2, POLYGON = 3 };
enum { STATIC = 0, DYNAMIC = 1, BULLET = 2 };
#define MAXV 8

typedef struct
{
    int type;
    int count;      /* polygon vertex count */
    v2 p[MAXV];     /* circle: p0 = center; capsule/segment: p0, p1; polygon: vertices */
    v2 n[MAXV];     /* polygon normals */
    float radius;   /* circle / capsule */
} shape_t;

/* Polygon::computeNormals, Shapes.hpp:142-154 */
static void poly_normals(shape_t* s)
{
    for (int i = 0; i < s->count; ++i)
    {
        const int j = (i + 1 == s->count) ? 0 : i + 1;
        const v2 e = vsub(s->p[j], s->p[i]);
        const float k = 1.0f / sqrtf(e.y * e.y + (-e.x) * (-e.x));
        s->n[i] = V(e.y * k, -e.x * k);
    }
}

static shape_t shape_from(int type, const float* g, int count)
{
    shape_t s;
    memset(&s, 0, sizeof(s));
    s.type = type;
    switch (type)
    {
    case CIRCLE: s.p[0] = V(g[0], g[1]); s.radius = g[2]; break;
    case CAPSULE: s.p[0] = V(g[0], g[1]); s.p[1] = V(g[2], g[3]); s.radius = g[4]; break;
    case SEGMENT: s.p[0] = V(g[0], g[1]); s.p[1] = V(g[2], g[3]); break;
    default:
        s.count = count;
        for (int i = 0; i < count && i < MAXV; ++i) s.p[i] = V(g[2 * i], g[2 * i + 1]);
        poly_normals(&s);
    }
    return s;
}

/* ------------------------------------------------------------------------------------------------
 * AABB: internal/geometry/AABB.hpp, src/geometry/AABB.cpp, ShapesUtils.hpp:13-60
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float minx, miny, maxx, maxy; } box_t;

static int box_intersects(const box_t* a, const box_t* b) /* AABB.cpp:34-37 */
{
    return a->minx <= b->maxx && a->maxx >= b->minx && a->miny <= b->maxy && a->maxy >= b->miny;
}
static int box_contains(const box_t* a, const box_t* o) /* AABB.cpp:29-32 */
{
    return a->minx <= o->minx && a->miny <= o->miny && a->maxx >= o->maxx && a->maxy >= o->maxy;
}
static box_t box_combine(box_t a, box_t b) /* AABB.cpp:273-281 */
{
    box_t r = {fmin_std(a.minx, b.minx), fmin_std(a.miny, b.miny), fmax_std(a.maxx, b.maxx), fmax_std(a.maxy, b.maxy)};
    return r;
}
static box_t compute_aabb(const shape_t* s, const xf_t* t) /* ShapesUtils.hpp:13-60 */
{
    box_t b;
    if (s->type == CIRCLE)
    {
        const v2 c = xapply(t, s->p[0]);
        b.minx = c.x - s->radius; b.miny = c.y - s->radius; b.maxx = c.x + s->radius; b.maxy = c.y + s->radius;
    }
    else if (s->type == CAPSULE || s->type == SEGMENT)
    {
        const v2 p1 = xapply(t, s->p[0]), p2 = xapply(t, s->p[1]);
        b.minx = fmin_std(p1.x, p2.x); b.miny = fmin_std(p1.y, p2.y);
        b.maxx = fmax_std(p1.x, p2.x); b.maxy = fmax_std(p1.y, p2.y);
        if (s->type == CAPSULE)
        {
            b.minx -= s->radius; b.miny -= s->radius; b.maxx += s->radius; b.maxy += s->radius;
        }
    }
    else
    {
        const v2 v0 = xapply(t, s->p[0]);
        b.minx = b.maxx = v0.x; b.miny = b.maxy = v0.y;
        for (int i = 1; i < s->count; ++i)
        {
            const v2 v = xapply(t, s->p[i]);
            b.minx = fmin_std(b.minx, v.x); b.miny = fmin_std(b.miny, v.y);
            b.maxx = fmax_std(b.maxx, v.x); b.maxy = fmax_std(b.maxy, v.y);
        }
    }
    return b;
}
/* DynamicBVH::moveProxy with mFatAABBMargin = 0, displacementFactor = 2.2 (DynamicBVH.hpp:20,386-400;
 * AABB::fattened(margin, displacement), AABB.cpp:212-228) */
static void fat_leaf_rule(box_t* fat, const box_t* target)
{
    if (box_contains(fat, target)) return;
    const float dx = (target->minx - fat->minx) * 2.2f, dy = (target->miny - fat->miny) * 2.2f;
    box_t nf = *target;
    if (dx < 0.0f) nf.minx += dx; else nf.maxx += dx;
    if (dy < 0.0f) nf.miny += dy; else nf.maxy += dy;
    *fat = nf;
}

/* ray = start + d * t; finite: d = end - start, t in [0,1]; infinite: d = direction, t in [0, FLT_MAX] */
typedef struct { v2 start, second, d; int infinite; } ray_t;
static ray_t ray_make(const float* r4, int infinite)
{
    ray_t r; r.start = V(r4[0], r4[1]); r.second = V(r4[2], r4[3]); r.infinite = infinite;
    r.d = infinite ? r.second : vsub(r.second, r.start);
    return r;
}
/* AABB::intersects(Ray) / (InfiniteRay), AABB.cpp:39-164 */
static int box_ray(const box_t* b, const ray_t* r, float* t0, float* t1)
{
    float tmin = 0.0f, tmax = r->infinite ? FLT_MAX : 1.0f;
    for (int axis = 0; axis < 2; ++axis)
    {
        const float dir = axis ? r->d.y : r->d.x, origin = axis ? r->start.y : r->start.x;
        const float lo = axis ? b->miny : b->minx, hi = axis ? b->maxy : b->maxx;
        if (fabsf(dir) < 1e-8f)
        {
            if (origin < lo || origin > hi) return 0;
        }
        else
        {
            const float inv = 1.0f / dir;
            float entry = (lo - origin) * inv, exit = (hi - origin) * inv;
            if (inv < 0.0f) { const float t = entry; entry = exit; exit = t; }
            tmin = fmax_std(tmin, entry);
            tmax = fmin_std(tmax, exit);
            if (tmax < tmin) return 0;
        }
    }
    *t0 = tmin; *t1 = tmax;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * manifolds: Manifold.hpp
 * ---------------------------------------------------------------------------------------------- */
typedef struct { v2 point, anchorA, anchorB; float separation; } mpoint_t;
typedef struct { v2 normal; mpoint_t points[2]; unsigned char pointCount; unsigned char pad[3]; } manifold_t;

static void manifold_reverse(manifold_t* m) /* Manifold.hpp:43-50 */
{
    m->normal = vneg(m->normal);
    for (int i = 0; i < m->pointCount; ++i) { const v2 t = m->points[i].anchorA; m->points[i].anchorA = m->points[i].anchorB; m->points[i].anchorB = t; }
}
static void single_contact(manifold_t* m, v2 normal, v2 world, const xf_t* ta, const xf_t* tb, float sep)
{
    m->normal = normal;
    m->points[0].point = world;
    m->points[0].anchorA = xlocal(ta, world);
    m->points[0].anchorB = xlocal(tb, world);
    m->points[0].separation = sep;
    m->pointCount = 1;
}

static const float kEpsilon = 1e-6f; /* Collision.cpp:12 */

/* collide / areColliding (Circle, Circle): Collision.cpp:15-49, 519-526.  m == NULL -> boolean only */
static int circle_circle(const shape_t* a, const xf_t* ta, const shape_t* b, const xf_t* tb, manifold_t* m)
{
    const v2 ca = xapply(ta, a->p[0]), cb = xapply(tb, b->p[0]);
    const v2 delta = vsub(cb, ca);
    const float distSq = vlensq(delta), radius = a->radius + b->radius;
    if (distSq > (radius + kEpsilon) * (radius + kEpsilon)) return 0;
    if (!m) return 1;
    const float distance = sqrtf(distSq);
    const v2 normal = (distance > kEpsilon) ? vmul(delta, 1.0f / distance) : V(1.0f, 0.0f);
    const v2 contactA = vadd(ca, vmul(normal, a->radius)), contactB = vsub(cb, vmul(normal, b->radius));
    single_contact(m, normal, vmul(vadd(contactA, contactB), 0.5f), ta, tb, distance - radius);
    return 1;
}

/* (Capsule, Circle): Collision.cpp:51-100, 528-550.  `cap` may be a segment (radius 0, Collision.cpp:218-231) */
static int capsule_circle(const shape_t* cap, const xf_t* ta, const shape_t* cir, const xf_t* tb, manifold_t* m)
{
    const v2 c1 = xapply(ta, cap->p[0]), c2 = xapply(ta, cap->p[1]), cc = xapply(tb, cir->p[0]);
    const v2 seg = vsub(c2, c1);
    const float segLenSq = vlensq(seg);
    float projection = 0.0f;
    if (segLenSq > 0.0f)
    {
        projection = vdot(vsub(cc, c1), seg) / segLenSq;
        projection = fclamp_std(projection, 0.0f, 1.0f);
    }
    const v2 closest = vadd(c1, vmul(seg, projection));
    const v2 delta = vsub(cc, closest);
    const float distSq = vlensq(delta), sumRadius = cap->radius + cir->radius;
    if (distSq > (sumRadius + kEpsilon) * (sumRadius + kEpsilon)) return 0;
    if (!m) return 1;
    float distance = 0.0f;
    v2 normal = V(1.0f, 0.0f);
    if (distSq > kEpsilon * kEpsilon)
    {
        distance = sqrtf(distSq);
        normal = vmul(delta, 1.0f / distance);
    }
    const v2 contactA = vadd(closest, vmul(normal, cap->radius)), contactB = vsub(cc, vmul(normal, cir->radius));
    single_contact(m, normal, vmul(vadd(contactA, contactB), 0.5f), ta, tb, distance - sumRadius);
    return 1;
}

/* (Capsule, Capsule): Collision.cpp:110-208, 557-635 */
static int capsule_capsule(const shape_t* a, const xf_t* ta, const shape_t* b, const xf_t* tb, manifold_t* m)
{
    const v2 a1 = xapply(ta, a->p[0]), a2 = xapply(ta, a->p[1]), b1 = xapply(tb, b->p[0]), b2 = xapply(tb, b->p[1]);
    const v2 sa = vsub(a2, a1), sb = vsub(b2, b1), sd = vsub(a1, b1);
    const float la = vdot(sa, sa), lb = vdot(sb, sb), ddb = vdot(sb, sd);
    float pa = 0.0f, pb = 0.0f;
    if (la <= kEpsilon && lb <= kEpsilon)
    {
        pa = pb = 0.0f;
    }
    else if (la <= kEpsilon)
    {
        pa = 0.0f;
        pb = fclamp_std(ddb / lb, 0.0f, 1.0f);
    }
    else
    {
        const float dda = vdot(sa, sd);
        if (lb <= kEpsilon)
        {
            pb = 0.0f;
            pa = fclamp_std(-dda / la, 0.0f, 1.0f);
        }
        else
        {
            const float dab = vdot(sa, sb);
            const float denom = la * lb - dab * dab;
            if (denom != 0.0f) pa = fclamp_std((dab * ddb - dda * lb) / denom, 0.0f, 1.0f);
            else pa = 0.0f;
            const float tnum = dab * pa + ddb;
            if (tnum < 0.0f) { pb = 0.0f; pa = fclamp_std(-dda / la, 0.0f, 1.0f); }
            else if (tnum > lb) { pb = 1.0f; pa = fclamp_std((dab - dda) / la, 0.0f, 1.0f); }
            else pb = tnum / lb;
        }
    }
    const v2 pointA = vadd(a1, vmul(sa, pa)), pointB = vadd(b1, vmul(sb, pb));
    const v2 delta = vsub(pointB, pointA);
    const float distSq = vlensq(delta), radius = a->radius + b->radius;
    if (distSq > (radius + kEpsilon) * (radius + kEpsilon)) return 0;
    if (!m) return 1;
    float distance = 0.0f;
    v2 normal = V(1.0f, 0.0f);
    if (distSq > kEpsilon * kEpsilon)
    {
        distance = sqrtf(distSq);
        normal = vmul(delta, 1.0f / distance);
    }
    const v2 contactA = vadd(pointA, vmul(normal, a->radius)), contactB = vsub(pointB, vmul(normal, b->radius));
    single_contact(m, normal, vmul(vadd(contactA, contactB), 0.5f), ta, tb, distance - radius);
    return 1;
}

/* projectPolygon, Collision.cpp:249-261 */
static void project_polygon(const shape_t* p, const xf_t* t, v2 axis, float* outMin, float* outMax)
{
    float lo = vdot(axis, xapply(t, p->p[0])), hi = lo;
    for (int i = 1; i < p->count; ++i)
    {
        const float proj = vdot(axis, xapply(t, p->p[i]));
        lo = fmin_std(lo, proj);
        hi = fmax_std(hi, proj);
    }
    *outMin = lo; *outMax = hi;
}

#define CLIPCAP 64
/* clipWithPlane, Collision.cpp:274-306 */
static int clip_with_plane(const v2* in, int nin, v2* out, v2 point, v2 normal)
{
    int nout = 0;
    if (nin == 0) return 0;
    v2 prev = in[nin - 1];
    float pd = vdot(normal, vsub(prev, point));
    for (int k = 0; k < nin; ++k)
    {
        const v2 curr = in[k];
        const float cd = vdot(normal, vsub(curr, point));
        if (cd <= 0.0f)
        {
            if (pd > 0.0f && nout < CLIPCAP) out[nout++] = vadd(prev, vmul(vsub(curr, prev), pd / (pd - cd)));
            if (nout < CLIPCAP) out[nout++] = curr;
        }
        else if (pd <= 0.0f)
        {
            if (nout < CLIPCAP) out[nout++] = vadd(prev, vmul(vsub(curr, prev), pd / (pd - cd)));
        }
        prev = curr;
        pd = cd;
    }
    return nout;
}

/* (Polygon, Polygon): Collision.cpp:308-399, 668-690 */
static int polygon_polygon(const shape_t* a, const xf_t* ta, const shape_t* b, const xf_t* tb, manifold_t* m)
{
    float overlap = FLT_MAX;
    v2 smallest = V(1.0f, 0.0f);
    for (int pass = 0; pass < 2; ++pass)
    {
        const shape_t* owner = pass ? b : a;
        const xf_t* to = pass ? tb : ta;
        for (int i = 0; i < owner->count; ++i)
        {
            const v2 axis = xrot(to, owner->n[i]);
            float minA, maxA, minB, maxB;
            project_polygon(a, ta, axis, &minA, &maxA);
            project_polygon(b, tb, axis, &minB, &maxB);
            const float ao = fmin_std(maxA, maxB) - fmax_std(minA, minB);
            if (ao < 0.0f) return 0;
            if (ao < overlap) { overlap = ao; smallest = axis; }
        }
    }
    if (!m) return 1;

    v2 avgA = xapply(ta, a->p[0]);
    for (int i = 1; i < a->count; ++i) avgA = vadd(avgA, xapply(ta, a->p[i]));
    avgA = vdiv(avgA, (float)a->count);
    v2 avgB = xapply(tb, b->p[0]);
    for (int i = 1; i < b->count; ++i) avgB = vadd(avgB, xapply(tb, b->p[i]));
    avgB = vdiv(avgB, (float)b->count);
    if (vdot(vsub(avgB, avgA), smallest) < 0.0f) smallest = vneg(smallest);

    /* polygonIntersection, Collision.cpp:308-322 */
    v2 buf0[CLIPCAP], buf1[CLIPCAP];
    v2 *cur = buf0, *nxt = buf1;
    int n = a->count;
    for (int i = 0; i < a->count; ++i) cur[i] = xapply(ta, a->p[i]);
    for (int i = 0; i < b->count && n > 0; ++i)
    {
        n = clip_with_plane(cur, n, nxt, xapply(tb, b->p[i]), xrot(tb, b->n[i]));
        v2* t = cur; cur = nxt; nxt = t;
    }
    m->normal = smallest;
    if (n == 0)
    {
        single_contact(m, smallest, vmul(vadd(avgA, avgB), 0.5f), ta, tb, -overlap);
        return 1;
    }
    /* std::sort(contactVertices) then the first min(n, 2) (Collision.cpp:385-396): insertion sort, same order */
    for (int i = 1; i < n; ++i)
    {
        const v2 key = cur[i];
        int j = i - 1;
        while (j >= 0 && vless(key, cur[j])) { cur[j + 1] = cur[j]; --j; }
        cur[j + 1] = key;
    }
    const int count = n < 2 ? n : 2;
    for (int i = 0; i < count; ++i)
    {
        m->points[i].point = cur[i];
        m->points[i].anchorA = xlocal(ta, cur[i]);
        m->points[i].anchorB = xlocal(tb, cur[i]);
        m->points[i].separation = -overlap;
    }
    m->pointCount = (unsigned char)count;
    return 1;
}

/* isPointInsidePolygon, Collision.cpp:402-412 */
static int point_inside_polygon(v2 point, const shape_t* p, const xf_t* t)
{
    for (int i = 0; i < p->count; ++i)
        if (vdot(xrot(t, p->n[i]), vsub(point, xapply(t, p->p[i]))) > 0.0f) return 0;
    return 1;
}

/* (Polygon, Circle): Collision.cpp:414-473, 692-721 */
static int polygon_circle(const shape_t* p, const xf_t* ta, const shape_t* cir, const xf_t* tb, manifold_t* m)
{
    const v2 center = xapply(tb, cir->p[0]);
    const float r = cir->radius;
    if (!m)
    {
        if (point_inside_polygon(center, p, ta)) return 1;
        for (int i = 0; i < p->count; ++i)
        {
            const v2 v1 = xapply(ta, p->p[i]), v2_ = xapply(ta, p->p[(i + 1) % p->count]);
            const v2 edge = vsub(v2_, v1);
            const float el = vlensq(edge);
            float ep = 0.0f;
            if (el > 0.0f) ep = fclamp_std(vdot(vsub(center, v1), edge) / el, 0.0f, 1.0f);
            const v2 delta = vsub(center, vadd(v1, vmul(edge, ep)));
            if (vlensq(delta) <= r * r) return 1;
        }
        return 0;
    }
    float minDistSq = FLT_MAX;
    v2 closest = V(0.0f, 0.0f), normal = V(1.0f, 0.0f);
    for (int i = 0; i < p->count; ++i)
    {
        const v2 v1 = xapply(ta, p->p[i]), v2_ = xapply(ta, p->p[(i + 1) % p->count]);
        const v2 edge = vsub(v2_, v1);
        const float el = vlensq(edge);
        float ep = 0.0f;
        if (el > 0.0f) ep = fclamp_std(vdot(vsub(center, v1), edge) / el, 0.0f, 1.0f);
        const v2 onEdge = vadd(v1, vmul(edge, ep));
        const v2 delta = vsub(center, onEdge);
        const float distSq = vlensq(delta);
        if (distSq < minDistSq)
        {
            minDistSq = distSq;
            closest = onEdge;
            if (ep > 0.0f && ep < 1.0f) normal = xrot(ta, p->n[i]);
            else normal = vlensq(delta) > kEpsilon * kEpsilon ? vnormalize(delta) : xrot(ta, p->n[i]);
        }
    }
    const int inside = point_inside_polygon(center, p, ta);
    if (!inside && minDistSq > r * r) return 0;
    const float distance = sqrtf(minDistSq);
    if (!inside && distance > kEpsilon) normal = vmul(vsub(center, closest), 1.0f / distance);
    else if (inside) normal = vneg(normal);
    const v2 contactB = vsub(center, vmul(normal, r));
    single_contact(m, normal, vmul(vadd(closest, contactB), 0.5f), ta, tb, inside ? -(r + distance) : (distance - r));
    return 1;
}

/* Capsule as the 2-gon of Collision.cpp:483-496 (radius dropped) */
static shape_t capsule_as_polygon(const shape_t* cap)
{
    shape_t p;
    memset(&p, 0, sizeof(p));
    const v2 axis = vnormalize(vsub(cap->p[1], cap->p[0]));
    const v2 normal = V(-axis.y, axis.x);
    p.type = POLYGON; p.count = 2;
    p.p[0] = cap->p[0]; p.p[1] = cap->p[1];
    p.n[0] = normal; p.n[1] = vneg(normal);
    return p;
}
/* Segment as Capsule{start, end, 0}, Collision.cpp:211-246 */
static shape_t segment_as_capsule(const shape_t* seg)
{
    shape_t c = *seg;
    c.type = CAPSULE; c.radius = 0.0f;
    return c;
}

/* The 16 overloads of collide() / areColliding(), following the reference's own delegation chains
 * (Collision.cpp:102-107, 211-246, 475-517, 552-555, 637-666, 723-748).  m == NULL: boolean only. */
static int collide_dispatch(const shape_t* a, const xf_t* ta, const shape_t* b, const xf_t* tb, manifold_t* m)
{
    int hit;
    shape_t tmp, tmp2;
    if (m) memset(m, 0, sizeof(*m));
    switch (a->type * 4 + b->type)
    {
    case CIRCLE * 4 + CIRCLE: return circle_circle(a, ta, b, tb, m);
    case CAPSULE * 4 + CIRCLE: return capsule_circle(a, ta, b, tb, m);
    case CIRCLE * 4 + CAPSULE: hit = capsule_circle(b, tb, a, ta, m); if (m && hit) manifold_reverse(m); return hit;
    case CAPSULE * 4 + CAPSULE: return capsule_capsule(a, ta, b, tb, m);
    case SEGMENT * 4 + SEGMENT: tmp = segment_as_capsule(a); tmp2 = segment_as_capsule(b); return capsule_capsule(&tmp, ta, &tmp2, tb, m);
    case CIRCLE * 4 + SEGMENT: tmp = segment_as_capsule(b); hit = capsule_circle(&tmp, tb, a, ta, m); if (m && hit) manifold_reverse(m); return hit;
    case SEGMENT * 4 + CIRCLE: tmp = segment_as_capsule(a); return capsule_circle(&tmp, ta, b, tb, m);
    case CAPSULE * 4 + SEGMENT: tmp = segment_as_capsule(b); return capsule_capsule(a, ta, &tmp, tb, m);
    case SEGMENT * 4 + CAPSULE: tmp = segment_as_capsule(a); hit = capsule_capsule(b, tb, &tmp, ta, m); if (m && hit) manifold_reverse(m); return hit;
    case POLYGON * 4 + POLYGON: return polygon_polygon(a, ta, b, tb, m);
    case POLYGON * 4 + CIRCLE: return polygon_circle(a, ta, b, tb, m);
    case CIRCLE * 4 + POLYGON: hit = polygon_circle(b, tb, a, ta, m); if (m && hit) manifold_reverse(m); return hit;
    case CAPSULE * 4 + POLYGON: tmp = capsule_as_polygon(a); return polygon_polygon(&tmp, ta, b, tb, m);
    case POLYGON * 4 + CAPSULE: tmp = capsule_as_polygon(b); hit = polygon_polygon(&tmp, tb, a, ta, m); if (m && hit) manifold_reverse(m); return hit;
    case POLYGON * 4 + SEGMENT: /* collide(polygon, tA, segCap, tB) -> the Polygon/Capsule overload above */
        tmp = capsule_as_polygon(b); hit = polygon_polygon(&tmp, tb, a, ta, m); if (m && hit) manifold_reverse(m); return hit;
    default: /* SEGMENT, POLYGON: collide(polygon, tB, segment, tA).reverse() = double reverse of (segCap-gon, tA, polygon, tB) */
        tmp = capsule_as_polygon(a); hit = polygon_polygon(&tmp, ta, b, tb, m);
        if (m && hit) { manifold_reverse(m); manifold_reverse(m); }
        return hit;
    }
}

/* ------------------------------------------------------------------------------------------------
 * sweeps: internal/collision/Collision.hpp:71-124, 151-213, 246-307
 * ---------------------------------------------------------------------------------------------- */
typedef struct { xf_t start, end; int moving; } motion_t;

static xf_t motion_at(const motion_t* mo, float f)
{
    if (!mo->moving) return mo->start;
    /* Transform(start.translation + delta.translation * f, start.angle + delta.angle * f) with
     * delta = end - start (Transform::operator-, Transform.hpp:156-159: only translation and angle are used) */
    const float dx = mo->end.tx - mo->start.tx, dy = mo->end.ty - mo->start.ty, da = mo->end.angle - mo->start.angle;
    return xf_make(mo->start.tx + dx * f, mo->start.ty + dy * f, mo->start.angle + da * f);
}

typedef int (*sweep_pred)(void* user, const xf_t* ta, const xf_t* tb);

/* sweepHelper: 0 if colliding at the stored start transforms; 8 coarse steps; <= 10 bisections; returns 1 on hit */
static int sweep_helper(const motion_t* ma, const motion_t* mb, sweep_pred pred, void* user, float* fraction)
{
    if (pred(user, &ma->start, &mb->start)) { *fraction = 0.0f; return 1; }
    float lower = 0.0f, upper = 1.0f;
    for (int step = 1; step <= 8; ++step)
    {
        const float f = (float)step / 8.0f;
        const xf_t ta = motion_at(ma, f), tb = motion_at(mb, f);
        if (pred(user, &ta, &tb)) { lower = (float)(step - 1) / 8.0f; upper = f; break; }
    }
    if (lower == 0.0f && upper == 1.0f) return 0;
    for (int it = 0; it < 10; ++it)
    {
        const float mid = 0.5f * (lower + upper);
        const xf_t ta = motion_at(ma, mid), tb = motion_at(mb, mid);
        if (pred(user, &ta, &tb)) upper = mid; else lower = mid;
        if (upper - lower < 1e-6f) break;
    }
    *fraction = upper;
    return 1;
}

typedef struct { const shape_t *a, *b; } pair_user;
static int pred_colliding(void* user, const xf_t* ta, const xf_t* tb)
{
    const pair_user* u = (const pair_user*)user;
    return collide_dispatch(u->a, ta, u->b, tb, NULL);
}
/* sweep(shape, T0, T1, shape, T) / sweep(shape, T0, T1, shape, T0', T1'): manifold = collide at fractionUpper */
static int sweep_shapes(const shape_t* a, const motion_t* ma, const shape_t* b, const motion_t* mb, float* fraction, manifold_t* m)
{
    pair_user u = {a, b};
    if (!sweep_helper(ma, mb, pred_colliding, &u, fraction)) return 0;
    const xf_t ta = motion_at(ma, *fraction), tb = motion_at(mb, *fraction);
    collide_dispatch(a, &ta, b, &tb, m);
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * shape raycasts: src/geometry/RaycastShapes.cpp
 * ---------------------------------------------------------------------------------------------- */
static int ray_circle(v2 localCenter, float radius, const xf_t* t, const ray_t* r, float* t0, float* t1) /* :11-88 */
{
    const v2 center = xapply(t, localCenter);
    const v2 from = vsub(r->start, center);
    const float dl = vdot(r->d, r->d), sp = vdot(from, r->d), sd = vdot(from, from) - radius * radius;
    if (dl <= 0.0f)
    {
        if (sd <= 0.0f) { *t0 = 0.0f; *t1 = 0.0f; return 1; }
        return 0;
    }
    if (sd > 0.0f && sp > 0.0f) return 0;
    const float disc = sp * sp - dl * sd;
    if (disc < 0.0f) return 0;
    const float root = sqrtf(disc);
    float entry = (-sp - root) / dl, exit = (-sp + root) / dl;
    if (entry > exit) { const float tt = entry; entry = exit; exit = tt; }
    if (r->infinite)
    {
        if (exit < 0.0f) return 0;
        entry = fmax_std(entry, 0.0f);
    }
    else
    {
        if (entry > 1.0f || exit < 0.0f) return 0;
        entry = fmax_std(entry, 0.0f);
        exit = fmin_std(exit, 1.0f);
        if (entry > exit) return 0;
    }
    *t0 = entry; *t1 = exit;
    return 1;
}
static int ray_segment(const shape_t* s, const xf_t* t, const ray_t* r, float* t0, float* t1) /* :90-134 */
{
    const v2 p1 = xapply(t, s->p[0]), p2 = xapply(t, s->p[1]);
    const v2 seg = vsub(p2, p1), delta = vsub(r->start, p1);
    const float csr = vcross(seg, r->d), cds = vcross(delta, seg);
    if (fabsf(csr) < 1e-8f) return 0;
    const float ts = vcross(delta, r->d) / csr, tr = cds / csr;
    if (ts >= 0.0f && ts <= 1.0f && tr >= 0.0f && (r->infinite || tr <= 1.0f)) { *t0 = tr; *t1 = tr; return 1; }
    return 0;
}
static int ray_polygon(const shape_t* p, const xf_t* t, const ray_t* r, float* t0, float* t1) /* :136-206 */
{
    float lower = 0.0f, upper = r->infinite ? FLT_MAX : 1.0f;
    for (int i = 0; i < p->count; ++i)
    {
        const v2 vertex = xapply(t, p->p[i]), normal = xrot(t, p->n[i]);
        const float num = vdot(normal, vsub(vertex, r->start)), den = vdot(normal, r->d);
        if (fabsf(den) < 1e-8f)
        {
            if (num < 0.0f) return 0;
            continue;
        }
        const float tt = num / den;
        if (den < 0.0f) lower = fmax_std(lower, tt); else upper = fmin_std(upper, tt);
        if (upper < lower) return 0;
    }
    *t0 = lower; *t1 = upper;
    return 1;
}
static int ray_capsule(const shape_t* c, const xf_t* t, const ray_t* r, float* t0, float* t1) /* :208-320 */
{
    const v2 ws = xapply(t, c->p[0]), we = xapply(t, c->p[1]);
    const float radius = c->radius;
    const v2 axis = vsub(we, ws);
    const float length = sqrtf(axis.x * axis.x + axis.y * axis.y);
    const xf_t ident = xf_identity();
    if (length <= 1e-8f) return ray_circle(ws, radius, &ident, r, t0, t1);
    const v2 au = vmul(axis, 1.0f / length), side = V(-au.y, au.x);
    ray_t lr;
    lr.infinite = r->infinite;
    { const v2 d = vsub(r->start, ws); lr.start = V(vdot(d, au), vdot(d, side)); }
    if (r->infinite) lr.d = V(vdot(r->d, au), vdot(r->d, side));
    else { const v2 d = vsub(r->second, ws); lr.second = V(vdot(d, au), vdot(d, side)); lr.d = vsub(lr.second, lr.start); }
    const box_t core = {0.0f, -radius, length, radius};
    float entry = FLT_MAX, exit = FLT_MIN; /* numeric_limits<float>::min(): smallest POSITIVE, as in the reference */
    int hit = 0;
    float a, b;
    if (box_ray(&core, &lr, &a, &b)) { hit = 1; entry = fmin_std(entry, a); exit = fmax_std(exit, b); }
    if (ray_circle(V(0.0f, 0.0f), radius, &ident, &lr, &a, &b)) { hit = 1; entry = fmin_std(entry, a); exit = fmax_std(exit, b); }
    if (ray_circle(V(length, 0.0f), radius, &ident, &lr, &a, &b)) { hit = 1; entry = fmin_std(entry, a); exit = fmax_std(exit, b); }
    if (!hit) return 0;
    if (r->infinite)
    {
        if (exit < 0.0f || entry > exit) return 0;
        entry = fmax_std(entry, 0.0f);
    }
    else
    {
        if (exit < 0.0f || entry > 1.0f || exit < entry) return 0;
        entry = fmax_std(entry, 0.0f);
        exit = fmin_std(exit, 1.0f);
        if (entry > exit) return 0;
    }
    *t0 = entry; *t1 = exit;
    return 1;
}
static int ray_shape(const shape_t* s, const xf_t* t, const ray_t* r, float* t0, float* t1)
{
    switch (s->type)
    {
    case CIRCLE: return ray_circle(s->p[0], s->radius, t, r, t0, t1);
    case CAPSULE: return ray_capsule(s, t, r, t0, t1);
    case SEGMENT: return ray_segment(s, t, r, t0, t1);
    default: return ray_polygon(s, t, r, t0, t1);
    }
}
/* sweep(shape, T0, T1, ray), Collision.hpp:323-344: the LAST successful raycast met during sweepHelper */
typedef struct { const shape_t* s; const ray_t* r; int any; float t0, t1; } raysweep_user;
static int pred_ray(void* user, const xf_t* ta, const xf_t* tb)
{
    (void)tb;
    raysweep_user* u = (raysweep_user*)user;
    float a, b;
    if (ray_shape(u->s, ta, u->r, &a, &b)) { u->any = 1; u->t0 = a; u->t1 = b; return 1; }
    return 0;
}
static int ray_sweep(const shape_t* s, const xf_t* start, const xf_t* end, const ray_t* r, float* t0, float* t1)
{
    motion_t ma = {*start, *end, 1}, mb = {xf_identity(), xf_identity(), 0};
    raysweep_user u = {s, r, 0, 0.0f, 0.0f};
    float fraction;
    sweep_helper(&ma, &mb, pred_ray, &u, &fraction);
    *t0 = u.t0; *t1 = u.t1;
    return u.any;
}

/* ------------------------------------------------------------------------------------------------
 * Registry: include/colli2de/Registry.hpp
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
    uint32_t id;
    int used, body;
    xf_t transform, previous;
    uint32_t* shapes; uint32_t nshapes, capshapes; /* mShapeIds, insertion order */
} entity_t;

typedef struct
{
    shape_t shape;
    box_t tight, fat;     /* ShapeInstance::mBoundingBox; leaf box of the tree */
    uint64_t category, mask;
    uint32_t entity;      /* entity slot */
    int active, alive;
    int ghost;            /* multi-GPU halo copy (extension, see include/c2d_gpu.h c2d_gpu_shape_set_ghost) */
} inst_t;

typedef struct { uint32_t eA, eB, sA, sB; manifold_t m; } pairrec_t;   /* 84 bytes */
typedef struct { uint32_t entity, shape; float entry[2], exit[2], t0, t1; } hitrec_t; /* 32 bytes */
typedef struct { uint32_t n; uint32_t* ids; float* dxy; } staged_t;

struct rn_ctx
{
    entity_t* ents; uint32_t nents, capents;
    uint32_t* table; uint32_t tcap; /* id -> slot + 1, open addressing */
    uint32_t live;
    inst_t* shapes; uint32_t nshapes, capshapes;
    uint32_t* freeShapes; uint32_t nfree, capfree;
    pairrec_t* pairs; uint64_t npairs, cappairs;
    hitrec_t* hits; uint64_t nhits, caphits;
    staged_t* staged; uint32_t nstaged;
    uint32_t* cands; uint64_t ncands; /* last rn_candidates */
    char error[256];
};

static void* grow(void* p, size_t elem, uint64_t need, uint64_t* cap)
{
    if (need <= *cap) return p;
    uint64_t c = *cap ? *cap : 64;
    while (c < need) c *= 2;
    *cap = c;
    return realloc(p, elem * c);
}
static uint32_t hash_u32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
static void table_insert(rn_ctx* c, uint32_t id, uint32_t slot)
{
    uint32_t pos = hash_u32(id) & (c->tcap - 1);
    while (c->table[pos] != 0 && c->table[pos] != 0xffffffffu) pos = (pos + 1) & (c->tcap - 1);
    c->table[pos] = slot + 1;
}
static void table_rebuild(rn_ctx* c, uint32_t cap)
{
    free(c->table);
    c->tcap = cap;
    c->table = (uint32_t*)calloc(cap, sizeof(uint32_t));
    for (uint32_t i = 0; i < c->nents; ++i) if (c->ents[i].used) table_insert(c, c->ents[i].id, i);
}
static int find_entity(const rn_ctx* c, uint32_t id, uint32_t* slot)
{
    if (!c->tcap) return 0;
    uint32_t pos = hash_u32(id) & (c->tcap - 1);
    while (c->table[pos] != 0)
    {
        if (c->table[pos] != 0xffffffffu && c->ents[c->table[pos] - 1].used && c->ents[c->table[pos] - 1].id == id)
        { *slot = c->table[pos] - 1; return 1; }
        pos = (pos + 1) & (c->tcap - 1);
    }
    return 0;
}
static entity_t* entity_of(rn_ctx* c, uint32_t id)
{
    uint32_t slot;
    if (!find_entity(c, id, &slot)) { strcpy(c->error, "Assertion failed: Entity with this ID does not exist"); return NULL; }
    return &c->ents[slot];
}

const char* rn_backend_name(void) { return "oracle"; }
rn_ctx* rn_create(int method, uint32_t cell) { (void)method; (void)cell; return (rn_ctx*)calloc(1, sizeof(rn_ctx)); }
void rn_clear(rn_ctx* c)
{
    for (uint32_t i = 0; i < c->nents; ++i) free(c->ents[i].shapes);
    free(c->ents); free(c->table); free(c->shapes); free(c->freeShapes);
    c->ents = NULL; c->table = NULL; c->shapes = NULL; c->freeShapes = NULL;
    c->nents = c->capents = c->tcap = c->live = c->nshapes = c->capshapes = c->nfree = c->capfree = 0;
}
void rn_destroy(rn_ctx* c)
{
    if (!c) return;
    rn_clear(c);
    for (uint32_t i = 0; i < c->nstaged; ++i) { free(c->staged[i].ids); free(c->staged[i].dxy); }
    free(c->staged); free(c->pairs); free(c->hits); free(c->cands); free(c);
}
const char* rn_last_error(rn_ctx* c) { return c->error; }
uint64_t rn_size(rn_ctx* c) { return c->live; }
void* rn_native_context(rn_ctx* c) { (void)c; return NULL; }
void rn_query_stats(rn_ctx* c, double* out16) { (void)c; memset(out16, 0, 16 * sizeof(double)); }

/* Registry::createEntity, Registry.hpp:262-270 */
static void create_entity(rn_ctx* c, uint32_t id, int body, xf_t t)
{
    if ((uint64_t)(c->live + 1) * 2 >= c->tcap) table_rebuild(c, c->tcap ? c->tcap * 2 : 64);
    uint64_t cap = c->capents;
    c->ents = (entity_t*)grow(c->ents, sizeof(entity_t), (uint64_t)c->nents + 1, &cap);
    c->capents = (uint32_t)cap;
    entity_t* e = &c->ents[c->nents];
    memset(e, 0, sizeof(*e));
    e->id = id; e->used = 1; e->body = body; e->transform = t; e->previous = t;
    table_insert(c, id, c->nents);
    ++c->nents; ++c->live;
}
void rn_create_entities(rn_ctx* c, uint32_t n, const uint32_t* ids, const uint8_t* body, const float* xf3)
{
    for (uint32_t i = 0; i < n; ++i) create_entity(c, ids[i], body[i], xf_make(xf3[3 * i], xf3[3 * i + 1], xf3[3 * i + 2]));
}
void rn_create_entities_raw(rn_ctx* c, uint32_t n, const uint32_t* ids, const uint8_t* body, const float* xf5)
{
    for (uint32_t i = 0; i < n; ++i) create_entity(c, ids[i], body[i], xf_raw(xf5 + 5 * i));
}

/* Registry::addShape, Registry.hpp:272-330 (ShapeIds recycled LIFO, 283-289) + DynamicBVH::addProxy 352-365 */
void rn_add_shapes(rn_ctx* c, uint32_t n, const uint32_t* entity, const uint8_t* type, const float* geom16,
                   const uint8_t* vcount, const uint64_t* cat, const uint64_t* mask, uint32_t* out_ids)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        uint32_t slot;
        if (!find_entity(c, entity[i], &slot)) { strcpy(c->error, "Assertion failed: Entity with this ID does not exist"); return; }
        entity_t* e = &c->ents[slot];
        uint32_t sid;
        if (c->nfree) sid = c->freeShapes[--c->nfree];
        else
        {
            uint64_t cap = c->capshapes;
            c->shapes = (inst_t*)grow(c->shapes, sizeof(inst_t), (uint64_t)c->nshapes + 1, &cap);
            c->capshapes = (uint32_t)cap;
            sid = c->nshapes++;
        }
        inst_t* s = &c->shapes[sid];
        memset(s, 0, sizeof(*s));
        s->shape = shape_from(type[i], geom16 + 16 * i, vcount[i]);
        s->tight = compute_aabb(&s->shape, &e->transform);
        s->fat = s->tight; /* fattened(0) */
        s->category = cat ? cat[i] : 1ull;
        s->mask = mask ? mask[i] : ~0ull;
        s->entity = slot; s->active = 1; s->alive = 1;
        uint64_t cap = e->capshapes;
        e->shapes = (uint32_t*)grow(e->shapes, sizeof(uint32_t), (uint64_t)e->nshapes + 1, &cap);
        e->capshapes = (uint32_t)cap;
        e->shapes[e->nshapes++] = sid;
        if (out_ids) out_ids[i] = sid;
    }
}
static void push_free(rn_ctx* c, uint32_t sid)
{
    uint64_t cap = c->capfree;
    c->freeShapes = (uint32_t*)grow(c->freeShapes, sizeof(uint32_t), (uint64_t)c->nfree + 1, &cap);
    c->capfree = (uint32_t)cap;
    c->freeShapes[c->nfree++] = sid;
}
void rn_remove_shapes(rn_ctx* c, uint32_t n, const uint32_t* ids) /* Registry.hpp:332-351 */
{
    for (uint32_t i = 0; i < n; ++i)
    {
        inst_t* s = &c->shapes[ids[i]];
        entity_t* e = &c->ents[s->entity];
        s->alive = 0;
        push_free(c, ids[i]);
        for (uint32_t k = 0; k < e->nshapes; ++k)
            if (e->shapes[k] == ids[i]) { memmove(e->shapes + k, e->shapes + k + 1, (e->nshapes - k - 1) * sizeof(uint32_t)); --e->nshapes; break; }
    }
}
void rn_remove_entities(rn_ctx* c, uint32_t n, const uint32_t* ids) /* Registry.hpp:360-380 */
{
    for (uint32_t i = 0; i < n; ++i)
    {
        uint32_t slot;
        if (!find_entity(c, ids[i], &slot)) { strcpy(c->error, "Assertion failed: Entity with this ID does not exist"); return; }
        entity_t* e = &c->ents[slot];
        for (uint32_t k = 0; k < e->nshapes; ++k) { c->shapes[e->shapes[k]].alive = 0; push_free(c, e->shapes[k]); }
        free(e->shapes);
        /* tombstone in the table */
        uint32_t pos = hash_u32(ids[i]) & (c->tcap - 1);
        while (c->table[pos] != slot + 1) pos = (pos + 1) & (c->tcap - 1);
        c->table[pos] = 0xffffffffu;
        memset(e, 0, sizeof(*e));
        --c->live;
    }
}
void rn_set_ghost(rn_ctx* c, uint32_t n, const uint32_t* ids, const uint8_t* ghost)
{
    for (uint32_t i = 0; i < n; ++i) c->shapes[ids[i]].ghost = ghost[i] != 0;
}
void rn_set_active(rn_ctx* c, uint32_t n, const uint32_t* ids, const uint8_t* active) /* Registry.hpp:353-358 */
{
    for (uint32_t i = 0; i < n; ++i) c->shapes[ids[i]].active = active[i] != 0;
}

/* moveDynamicShape / moveBulletShape (Registry.hpp:191-212) */
static void move_shapes(rn_ctx* c, entity_t* e, const xf_t* t)
{
    for (uint32_t k = 0; k < e->nshapes; ++k)
    {
        inst_t* s = &c->shapes[e->shapes[k]];
        const box_t old = s->tight;
        s->tight = compute_aabb(&s->shape, t);
        if (e->body == BULLET) { const box_t swept = box_combine(old, s->tight); fat_leaf_rule(&s->fat, &swept); }
        else fat_leaf_rule(&s->fat, &s->tight);
    }
}
/* Registry::moveEntity(id, Translation), Registry.hpp:434-453; translate*Shape 214-231; AABB::translate AABB.cpp:248-252 */
static void move_translate(rn_ctx* c, uint32_t id, float dx, float dy)
{
    entity_t* e = entity_of(c, id);
    if (!e) return;
    for (uint32_t k = 0; k < e->nshapes; ++k)
    {
        inst_t* s = &c->shapes[e->shapes[k]];
        const box_t old = s->tight;
        s->tight.minx += dx; s->tight.miny += dy; s->tight.maxx += dx; s->tight.maxy += dy;
        if (e->body == BULLET) { const box_t swept = box_combine(old, s->tight); fat_leaf_rule(&s->fat, &swept); }
        else fat_leaf_rule(&s->fat, &s->tight);
    }
    if (e->body == BULLET) e->previous = e->transform;
    e->transform.tx += dx; e->transform.ty += dy;
}
void rn_move_translate(rn_ctx* c, uint32_t n, const uint32_t* ids, const float* dxy)
{
    for (uint32_t i = 0; i < n; ++i) move_translate(c, ids[i], dxy[2 * i], dxy[2 * i + 1]);
}
void rn_move_transform(rn_ctx* c, uint32_t n, const uint32_t* ids, const float* d) /* Registry.hpp:413-432 */
{
    for (uint32_t i = 0; i < n; ++i)
    {
        entity_t* e = entity_of(c, ids[i]);
        if (!e) return;
        const xf_t delta = xf_make(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        if (e->body == BULLET) e->previous = e->transform;
        xf_add(&e->transform, &delta);
        move_shapes(c, e, &e->transform);
    }
}
void rn_teleport_transform(rn_ctx* c, uint32_t n, const uint32_t* ids, const float* xf3) /* Registry.hpp:382-401 */
{
    for (uint32_t i = 0; i < n; ++i)
    {
        entity_t* e = entity_of(c, ids[i]);
        if (!e) return;
        const xf_t t = xf_make(xf3[3 * i], xf3[3 * i + 1], xf3[3 * i + 2]);
        move_shapes(c, e, &t);
        if (e->body == BULLET) e->previous = e->transform;
        e->transform = t;
    }
}
void rn_teleport_translation(rn_ctx* c, uint32_t n, const uint32_t* ids, const float* xy) /* Registry.hpp:403-411 */
{
    for (uint32_t i = 0; i < n; ++i)
    {
        entity_t* e = entity_of(c, ids[i]);
        if (!e) return;
        move_translate(c, ids[i], xy[2 * i] - e->transform.tx, xy[2 * i + 1] - e->transform.ty);
    }
}
void rn_get_transforms(rn_ctx* c, uint32_t n, const uint32_t* ids, float* out)
{
    for (uint32_t i = 0; i < n; ++i) { entity_t* e = entity_of(c, ids[i]); if (!e) return; xf_store(&e->transform, out + 5 * i); }
}
void rn_get_prev_transforms(rn_ctx* c, uint32_t n, const uint32_t* ids, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        entity_t* e = entity_of(c, ids[i]);
        if (!e) return;
        xf_store(e->body == BULLET ? &e->previous : &e->transform, out + 5 * i);
    }
}

int rn_stage_translations(rn_ctx* c, uint32_t n, const uint32_t* ids, const float* dxy)
{
    c->staged = (staged_t*)realloc(c->staged, sizeof(staged_t) * (c->nstaged + 1));
    staged_t* s = &c->staged[c->nstaged];
    s->n = n;
    s->ids = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    s->dxy = (float*)malloc(sizeof(float) * 2 * (n ? n : 1));
    memcpy(s->ids, ids, sizeof(uint32_t) * n);
    memcpy(s->dxy, dxy, sizeof(float) * 2 * n);
    return (int)c->nstaged++;
}
void rn_apply_staged(rn_ctx* c, int h) { rn_move_translate(c, c->staged[h].n, c->staged[h].ids, c->staged[h].dxy); }
void rn_release_staged(rn_ctx* c, int h) { free(c->staged[h].ids); free(c->staged[h].dxy); c->staged[h].ids = NULL; c->staged[h].dxy = NULL; c->staged[h].n = 0; }

/* ---- getCollidingPairs, Registry.hpp:464-549 ------------------------------------------------------ */
typedef struct { float minx; uint32_t shape; } sweepitem_t;
typedef struct { uint32_t a, b; } cand_t;
static int cmp_sweep(const void* x, const void* y)
{
    const sweepitem_t *a = (const sweepitem_t*)x, *b = (const sweepitem_t*)y;
    if (a->minx < b->minx) return -1;
    if (a->minx > b->minx) return 1;
    return (a->shape > b->shape) - (a->shape < b->shape);
}
static int cmp_cand(const void* x, const void* y)
{
    const cand_t *a = (const cand_t*)x, *b = (const cand_t*)y;
    if (a->a != b->a) return a->a < b->a ? -1 : 1;
    if (a->b != b->b) return a->b < b->b ? -1 : 1;
    return 0;
}
/* group index in the reference's order: Dyn x Static, Dyn x Dyn, Bullet x Static, Bullet x Dyn, Bullet x Bullet;
 * -1 for Static x Static (never tested).  *swap = 1 when the pair must be emitted as (j, i). */
static int pair_group(int bi, int bj, uint32_t si, uint32_t sj, int* swap)
{
    *swap = 0;
    if (bi == bj)
    {
        if (bi == STATIC) return -1;
        *swap = si > sj;
        return bi == DYNAMIC ? 1 : 4;
    }
    const int hi = bi > bj ? bi : bj, lo = bi > bj ? bj : bi;
    *swap = bj > bi; /* A is the Dynamic (vs Static) or the Bullet side: findAllCollisions(other) is called on that tree */
    if (hi == DYNAMIC) return 0;
    return lo == STATIC ? 2 : 3;
}

/* narrowPhaseCollision, Registry.hpp:608-710 */
static void narrow_phase(rn_ctx* c, uint32_t sa, uint32_t sb)
{
    const inst_t *A = &c->shapes[sa], *B = &c->shapes[sb];
    const entity_t *ea = &c->ents[A->entity], *eb = &c->ents[B->entity];
    if (A->entity == B->entity) return;
    if (!A->active || !B->active) return;
    if (A->ghost || B->ghost)
    {
        /* owner rule: shape A of a cross-tree pair, the shape of the entity with the smaller id of a same-tree pair */
        const inst_t* owner = (ea->body != eb->body || ea->id < eb->id) ? A : B;
        if (owner->ghost) return;
    }
    manifold_t m;
    memset(&m, 0, sizeof(m));
    int have;
    if (ea->body == BULLET && eb->body == BULLET)
    {
        motion_t ma = {ea->previous, ea->transform, 1}, mb = {eb->previous, eb->transform, 1};
        float f;
        have = sweep_shapes(&A->shape, &ma, &B->shape, &mb, &f, &m);
    }
    else if (ea->body == BULLET)
    {
        motion_t ma = {ea->previous, ea->transform, 1}, mb = {eb->transform, eb->transform, 0};
        float f;
        have = sweep_shapes(&A->shape, &ma, &B->shape, &mb, &f, &m);
    }
    else
    {
        collide_dispatch(&A->shape, &ea->transform, &B->shape, &eb->transform, &m);
        have = m.pointCount > 0;
    }
    if (!have) return;
    c->pairs = (pairrec_t*)grow(c->pairs, sizeof(pairrec_t), c->npairs + 1, &c->cappairs);
    pairrec_t* r = &c->pairs[c->npairs++];
    memset(r, 0, sizeof(*r));
    r->eA = ea->id; r->eB = eb->id; r->sA = sa; r->sB = sb; r->m = m;
}

/* broadphase: sort-and-sweep on the leaf boxes; candidate iff boxes intersect (non-strict) and
 * (categoryA & categoryB) != 0 (DynamicBVH.hpp:944-964).  Fills the five groups of Registry.hpp:487-545, each sorted. */
static void broadphase_groups(rn_ctx* c, cand_t* cand[5], uint64_t ncand[5])
{
    uint32_t n = 0;
    sweepitem_t* items = (sweepitem_t*)malloc(sizeof(sweepitem_t) * (c->nshapes ? c->nshapes : 1));
    for (uint32_t s = 0; s < c->nshapes; ++s)
        if (c->shapes[s].alive) { items[n].minx = c->shapes[s].fat.minx; items[n].shape = s; ++n; }
    qsort(items, n, sizeof(sweepitem_t), cmp_sweep);
    uint64_t capcand[5] = {0, 0, 0, 0, 0};
    for (int g = 0; g < 5; ++g) { cand[g] = NULL; ncand[g] = 0; }
    for (uint32_t i = 0; i < n; ++i)
    {
        const uint32_t si = items[i].shape;
        const inst_t* I = &c->shapes[si];
        const int bi = c->ents[I->entity].body;
        for (uint32_t j = i + 1; j < n && items[j].minx <= I->fat.maxx; ++j)
        {
            const uint32_t sj = items[j].shape;
            const inst_t* J = &c->shapes[sj];
            int swap;
            const int g = pair_group(bi, c->ents[J->entity].body, si, sj, &swap);
            if (g < 0) continue;
            if ((I->category & J->category) == 0) continue;
            if (!box_intersects(&I->fat, &J->fat)) continue;
            cand[g] = (cand_t*)grow(cand[g], sizeof(cand_t), ncand[g] + 1, &capcand[g]);
            cand[g][ncand[g]].a = swap ? sj : si;
            cand[g][ncand[g]].b = swap ? si : sj;
            ++ncand[g];
        }
    }
    free(items);
    for (int g = 0; g < 5; ++g)
        qsort(cand[g], ncand[g], sizeof(cand_t), cmp_cand); /* std::sort of each group, Registry.hpp:495-533 */
}

uint64_t rn_get_colliding_pairs(rn_ctx* c, double* seconds)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    c->npairs = 0;
    cand_t* cand[5];
    uint64_t ncand[5];
    broadphase_groups(c, cand, ncand);
    for (int g = 0; g < 5; ++g)
    {
        for (uint64_t k = 0; k < ncand[g]; ++k) narrow_phase(c, cand[g][k].a, cand[g][k].b);
        free(cand[g]);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return c->npairs;
}

/* ---- secondary oracles ---------------------------------------------- */
uint32_t rn_shape_slots(rn_ctx* c) { return c->nshapes; }
void rn_read_boxes(rn_ctx* c, uint32_t first, uint32_t count, float* tight4, float* fat4, uint8_t* alive)
{
    for (uint32_t i = 0; i < count; ++i)
    {
        if (first + i >= c->nshapes) { snprintf(c->error, sizeof(c->error), "rn_read_boxes: shape id out of range"); return; }
        const inst_t* I = &c->shapes[first + i];
        if (tight4) { tight4[4 * i] = I->tight.minx; tight4[4 * i + 1] = I->tight.miny; tight4[4 * i + 2] = I->tight.maxx; tight4[4 * i + 3] = I->tight.maxy; }
        if (fat4) { fat4[4 * i] = I->fat.minx; fat4[4 * i + 1] = I->fat.miny; fat4[4 * i + 2] = I->fat.maxx; fat4[4 * i + 3] = I->fat.maxy; }
        if (alive) alive[i] = I->alive ? 1 : 0;
    }
}
uint64_t rn_candidates(rn_ctx* c)
{
    cand_t* cand[5];
    uint64_t ncand[5], total = 0;
    broadphase_groups(c, cand, ncand);
    for (int g = 0; g < 5; ++g) total += ncand[g];
    free(c->cands);
    c->cands = (uint32_t*)malloc(sizeof(uint32_t) * 2 * (total ? total : 1));
    c->ncands = 0;
    for (int g = 0; g < 5; ++g)
    {
        for (uint64_t k = 0; k < ncand[g]; ++k) { c->cands[2 * c->ncands] = cand[g][k].a; c->cands[2 * c->ncands + 1] = cand[g][k].b; ++c->ncands; }
        free(cand[g]);
    }
    return c->ncands;
}
void rn_copy_candidates(rn_ctx* c, uint32_t* out) { if (c->ncands) memcpy(out, c->cands, c->ncands * 2 * sizeof(uint32_t)); }
uint64_t rn_count_colliding_pairs_device(rn_ctx* c, double* seconds) { return rn_get_colliding_pairs(c, seconds); }
uint64_t rn_get_colliding_pairs_view(rn_ctx* c, double* seconds, const void** records)
{
    const uint64_t n = rn_get_colliding_pairs(c, seconds);
    *records = c->pairs;
    return n;
}
void rn_copy_pairs(rn_ctx* c, void* out) { if (c->npairs) memcpy(out, c->pairs, c->npairs * sizeof(pairrec_t)); }

/* ---- rayCast, Registry.hpp:1050-1182; DynamicBVH::piercingRaycast 775-806 / 853-884 ------------------- */
typedef struct { float t0, t1; uint32_t shape; } aabbhit_t;
static int cmp_aabbhit(const void* x, const void* y)
{
    const aabbhit_t *a = (const aabbhit_t*)x, *b = (const aabbhit_t*)y;
    if (a->t0 != b->t0) return a->t0 < b->t0 ? -1 : 1;
    if (a->t1 != b->t1) return a->t1 < b->t1 ? -1 : 1;
    return (a->shape > b->shape) - (a->shape < b->shape);
}
static void raycast_one(rn_ctx* c, const ray_t* r, uint64_t mask)
{
    const uint64_t first = c->nhits;
    aabbhit_t* ah = (aabbhit_t*)malloc(sizeof(aabbhit_t) * (c->nshapes ? c->nshapes : 1));
    for (int body = 0; body < 3; ++body) /* Static, Dynamic, Bullet trees, Registry.hpp:1054-1056 */
    {
        uint32_t n = 0;
        for (uint32_t s = 0; s < c->nshapes; ++s)
        {
            const inst_t* S = &c->shapes[s];
            if (!S->alive || c->ents[S->entity].body != body) continue;
            if ((S->category & mask) == 0) continue; /* matchesMask(maskBits), DynamicBVH.hpp:789 */
            float a, b;
            if (box_ray(&S->fat, r, &a, &b)) { ah[n].t0 = a; ah[n].t1 = b; ah[n].shape = s; ++n; }
        }
        qsort(ah, n, sizeof(aabbhit_t), cmp_aabbhit); /* std::set<RaycastInfo> ordered by (entry, exit) */
        for (uint32_t k = 0; k < n; ++k)
        {
            if (k && ah[k].t0 == ah[k - 1].t0 && ah[k].t1 == ah[k - 1].t1) continue; /* equal key: not inserted (D10) */
            const inst_t* S = &c->shapes[ah[k].shape];
            const entity_t* e = &c->ents[S->entity];
            if (!S->active) continue;
            float t0, t1;
            const int hit = body == BULLET ? ray_sweep(&S->shape, &e->previous, &e->transform, r, &t0, &t1)
                                           : ray_shape(&S->shape, &e->transform, r, &t0, &t1);
            if (!hit) continue;
            /* hits.insert(...): std::set keyed on (entryTime, exitTime) — an equal key is dropped */
            int dup = 0;
            for (uint64_t q = first; q < c->nhits; ++q) if (c->hits[q].t0 == t0 && c->hits[q].t1 == t1) { dup = 1; break; }
            if (dup) continue;
            c->hits = (hitrec_t*)grow(c->hits, sizeof(hitrec_t), c->nhits + 1, &c->caphits);
            hitrec_t* h = &c->hits[c->nhits++];
            h->entity = e->id; h->shape = ah[k].shape; h->t0 = t0; h->t1 = t1;
            /* RaycastHit::fromRay, Ray.hpp:34-65 */
            const v2 en = vadd(r->start, vmul(r->d, t0)), ex = vadd(r->start, vmul(r->d, t1));
            h->entry[0] = en.x; h->entry[1] = en.y; h->exit[0] = ex.x; h->exit[1] = ex.y;
        }
    }
    free(ah);
    /* std::set iteration order: (entryTime, exitTime) ascending; insertion sort keeps it simple */
    for (uint64_t i = first + 1; i < c->nhits; ++i)
    {
        const hitrec_t key = c->hits[i];
        uint64_t j = i;
        while (j > first && (key.t0 < c->hits[j - 1].t0 || (key.t0 == c->hits[j - 1].t0 && key.t1 < c->hits[j - 1].t1)))
        { c->hits[j] = c->hits[j - 1]; --j; }
        c->hits[j] = key;
    }
}
uint64_t rn_raycast(rn_ctx* c, uint32_t n, const float* ray4, const uint8_t* infinite, const uint64_t* masks,
                    uint64_t* offsets, double* seconds)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    c->nhits = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        offsets[i] = c->nhits;
        const ray_t r = ray_make(ray4 + 4 * i, infinite && infinite[i]);
        raycast_one(c, &r, masks ? masks[i] : ~0ull);
    }
    offsets[n] = c->nhits;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return c->nhits;
}
void rn_copy_hits(rn_ctx* c, void* out) { if (c->nhits) memcpy(out, c->hits, c->nhits * sizeof(hitrec_t)); }

/* ---- getCollisions, Registry.hpp:551-605 --------------------------------------------------------------- */
static void push_pair(rn_ctx* c, const entity_t* ea, const entity_t* eb, uint32_t sa, uint32_t sb, const manifold_t* m)
{
    c->pairs = (pairrec_t*)grow(c->pairs, sizeof(pairrec_t), c->npairs + 1, &c->cappairs);
    pairrec_t* r = &c->pairs[c->npairs++];
    memset(r, 0, sizeof(*r));
    r->eA = ea->id; r->eB = eb->id; r->sA = sa; r->sB = sb; r->m = *m;
}
/* narrowPhaseCollision(queried, other) as reached from getCollisions: when exactly one side is a bullet the sweep is
 * (bullet shape moving, other shape) whichever side was queried, and the record keeps (queried, other)
 * (Registry.hpp:669-691) */
static void narrow_phase_query(rn_ctx* c, uint32_t sa, uint32_t sb)
{
    const inst_t *A = &c->shapes[sa], *B = &c->shapes[sb];
    const entity_t *ea = &c->ents[A->entity], *eb = &c->ents[B->entity];
    if (A->entity == B->entity || !A->active || !B->active) return;
    manifold_t m;
    memset(&m, 0, sizeof(m));
    float f;
    int have;
    if (ea->body == BULLET && eb->body == BULLET)
    {
        motion_t ma = {ea->previous, ea->transform, 1}, mb = {eb->previous, eb->transform, 1};
        have = sweep_shapes(&A->shape, &ma, &B->shape, &mb, &f, &m);
    }
    else if (ea->body == BULLET)
    {
        motion_t ma = {ea->previous, ea->transform, 1}, mb = {eb->transform, eb->transform, 0};
        have = sweep_shapes(&A->shape, &ma, &B->shape, &mb, &f, &m);
    }
    else if (eb->body == BULLET)
    {
        motion_t mb = {eb->previous, eb->transform, 1}, ma = {ea->transform, ea->transform, 0};
        have = sweep_shapes(&B->shape, &mb, &A->shape, &ma, &f, &m);
    }
    else
    {
        collide_dispatch(&A->shape, &ea->transform, &B->shape, &eb->transform, &m);
        have = m.pointCount > 0;
    }
    if (have) push_pair(c, ea, eb, sa, sb, &m);
}
uint64_t rn_get_collisions(rn_ctx* c, uint32_t id)
{
    c->npairs = 0;
    entity_t* e = entity_of(c, id);
    if (!e) return 0;
    for (int body = 0; body < 3; ++body)
    {
        if (body == STATIC && e->body == STATIC) continue; /* Registry.hpp:598-601 */
        for (uint32_t k = 0; k < e->nshapes; ++k)
        {
            const uint32_t q = e->shapes[k];
            const inst_t* Q = &c->shapes[q];
            if (!Q->active) continue;
            box_t qb = Q->tight;
            if (e->body == BULLET) { const box_t old = compute_aabb(&Q->shape, &e->previous); qb = box_combine(old, Q->tight); }
            for (uint32_t s = 0; s < c->nshapes; ++s)
            {
                const inst_t* S = &c->shapes[s];
                if (!S->alive || c->ents[S->entity].body != body) continue;
                if ((S->category & Q->mask) == 0) continue; /* node.matchesMask(maskBits), DynamicBVH.hpp:725 */
                if (!box_intersects(&S->fat, &qb)) continue;
                narrow_phase_query(c, q, s);
            }
        }
    }
    return c->npairs;
}
uint64_t rn_get_collisions_many(rn_ctx* c, uint32_t n, const uint32_t* ids, double* seconds)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i) total += rn_get_collisions(c, ids[i]);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return total;
}

/* ---- areColliding, Registry.hpp:748-864; boolSweepHelper Collision.hpp:126-149, 215-244 -------------------- */
static int sweep_bool(const shape_t* a, const motion_t* ma, const shape_t* b, const motion_t* mb)
{
    if (collide_dispatch(a, &ma->start, b, &mb->start, NULL)) return 1;
    for (int step = 1; step <= 8; ++step)
    {
        const float f = (float)step / 8.0f;
        const xf_t ta = motion_at(ma, f), tb = motion_at(mb, f);
        if (collide_dispatch(a, &ta, b, &tb, NULL)) return 1;
    }
    return 0;
}
static int entities_colliding(rn_ctx* c, uint32_t ida, uint32_t idb)
{
    entity_t *ea = entity_of(c, ida), *eb = entity_of(c, idb);
    if (!ea || !eb) return 0;
    if (ea->body == STATIC && eb->body == STATIC) return 0;
    for (uint32_t i = 0; i < ea->nshapes; ++i)
    {
        const inst_t* A = &c->shapes[ea->shapes[i]];
        if (!A->active) continue;
        for (uint32_t j = 0; j < eb->nshapes; ++j)
        {
            const inst_t* B = &c->shapes[eb->shapes[j]];
            if (!B->active) continue;
            int hit;
            if (ea->body == BULLET && eb->body == BULLET)
            {
                motion_t ma = {ea->previous, ea->transform, 1}, mb = {eb->previous, eb->transform, 1};
                hit = sweep_bool(&A->shape, &ma, &B->shape, &mb);
            }
            else if (ea->body == BULLET)
            {
                motion_t ma = {ea->previous, ea->transform, 1}, mb = {eb->transform, eb->transform, 0};
                hit = sweep_bool(&A->shape, &ma, &B->shape, &mb);
            }
            else if (eb->body == BULLET)
            {
                motion_t mb = {eb->previous, eb->transform, 1}, ma = {ea->transform, ea->transform, 0};
                hit = sweep_bool(&B->shape, &mb, &A->shape, &ma);
            }
            else hit = collide_dispatch(&A->shape, &ea->transform, &B->shape, &eb->transform, NULL);
            if (hit) return 1;
        }
    }
    return 0;
}
void rn_are_colliding(rn_ctx* c, uint32_t n, const uint32_t* a, const uint32_t* b, uint8_t* out)
{
    for (uint32_t i = 0; i < n; ++i) out[i] = (uint8_t)entities_colliding(c, a[i], b[i]);
}

/* ---- firstHitRayCast, Registry.hpp:866-1048 ------------------------------------------------------------------ */
void rn_first_hit_raycast(rn_ctx* c, uint32_t n, const float* ray4, const uint8_t* infinite, const uint64_t* masks,
                          uint8_t* found, void* out_hits32)
{
    hitrec_t* out = (hitrec_t*)out_hits32;
    aabbhit_t* ah = (aabbhit_t*)malloc(sizeof(aabbhit_t) * (c->nshapes ? c->nshapes : 1));
    for (uint32_t i = 0; i < n; ++i)
    {
        const ray_t r = ray_make(ray4 + 4 * i, infinite && infinite[i]);
        const uint64_t mask = masks ? masks[i] : ~0ull;
        float best = FLT_MAX;
        found[i] = 0;
        memset(&out[i], 0, sizeof(hitrec_t));
        for (int body = 0; body < 3; ++body)
        {
            uint32_t m = 0;
            for (uint32_t s = 0; s < c->nshapes; ++s)
            {
                const inst_t* S = &c->shapes[s];
                if (!S->alive || c->ents[S->entity].body != body || (S->category & mask) == 0) continue;
                float a, b;
                if (box_ray(&S->fat, &r, &a, &b)) { ah[m].t0 = a; ah[m].t1 = b; ah[m].shape = s; ++m; }
            }
            qsort(ah, m, sizeof(aabbhit_t), cmp_aabbhit);
            uint32_t setSize = 0;
            for (uint32_t k = 0; k < m; ++k) if (!k || ah[k].t0 != ah[k - 1].t0 || ah[k].t1 != ah[k - 1].t1) ++setSize;
            uint32_t toCheck = setSize;
            for (uint32_t k = 0; k < m; ++k)
            {
                if (k && ah[k].t0 == ah[k - 1].t0 && ah[k].t1 == ah[k - 1].t1) continue;
                const inst_t* S = &c->shapes[ah[k].shape];
                const entity_t* e = &c->ents[S->entity];
                if (!S->active) continue; /* the `continue` precedes the countdown */
                float t0, t1;
                const int hit = body == BULLET ? ray_sweep(&S->shape, &e->previous, &e->transform, &r, &t0, &t1)
                                               : ray_shape(&S->shape, &e->transform, &r, &t0, &t1);
                if (hit && t0 < best)
                {
                    best = t0;
                    found[i] = 1;
                    out[i].entity = e->id; out[i].shape = ah[k].shape; out[i].t0 = t0; out[i].t1 = t1;
                    const v2 en = vadd(r.start, vmul(r.d, t0)), ex = vadd(r.start, vmul(r.d, t1));
                    out[i].entry[0] = en.x; out[i].entry[1] = en.y; out[i].exit[0] = ex.x; out[i].exit[1] = ex.y;
                    toCheck = 5; /* check 5 more after an improvement */
                }
                if (--toCheck == 0) break;
            }
        }
    }
    free(ah);
}

/* ------------------------------------------------------------------------------------------------
 * batched free functions
 * ---------------------------------------------------------------------------------------------- */
void rn_collide(uint32_t n, const uint8_t* typeA, const float* geomA, const uint8_t* cntA, const float* xfA5,
                const uint8_t* typeB, const float* geomB, const uint8_t* cntB, const float* xfB5,
                void* out_manifold68, uint8_t* out_are_colliding)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const shape_t a = shape_from(typeA[i], geomA + 16 * i, cntA[i]), b = shape_from(typeB[i], geomB + 16 * i, cntB[i]);
        const xf_t ta = xf_raw(xfA5 + 5 * i), tb = xf_raw(xfB5 + 5 * i);
        if (out_manifold68)
        {
            manifold_t m;
            collide_dispatch(&a, &ta, &b, &tb, &m);
            memcpy((char*)out_manifold68 + 68 * (size_t)i, &m, 68);
        }
        if (out_are_colliding) out_are_colliding[i] = (uint8_t)collide_dispatch(&a, &ta, &b, &tb, NULL);
    }
}
void rn_sweep(uint32_t n, int mode, const uint8_t* typeA, const float* geomA, const uint8_t* cntA,
              const float* xfA5, const float* xfA5_end, const uint8_t* typeB, const float* geomB,
              const uint8_t* cntB, const float* xfB5, const float* xfB5_end, float* out_fraction,
              void* out_manifold68, uint8_t* out_hit)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const shape_t a = shape_from(typeA[i], geomA + 16 * i, cntA[i]), b = shape_from(typeB[i], geomB + 16 * i, cntB[i]);
        motion_t ma = {xf_raw(xfA5 + 5 * i), xf_raw(xfA5_end + 5 * i), 1};
        motion_t mb = {xf_raw(xfB5 + 5 * i), mode == 2 ? xf_raw(xfB5_end + 5 * i) : xf_raw(xfB5 + 5 * i), mode == 2};
        manifold_t m;
        memset(&m, 0, sizeof(m));
        float f = -1.0f;
        const int hit = sweep_shapes(&a, &ma, &b, &mb, &f, &m);
        if (!hit) { memset(&m, 0, sizeof(m)); f = -1.0f; }
        out_hit[i] = (uint8_t)hit;
        if (out_fraction) out_fraction[i] = f;
        if (out_manifold68) memcpy((char*)out_manifold68 + 68 * (size_t)i, &m, 68);
    }
}
void rn_raycast_shape(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                      const float* ray4, const uint8_t* infinite, float* out_times2, uint8_t* out_hit)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const shape_t s = shape_from(type[i], geom + 16 * i, cnt[i]);
        const xf_t t = xf_raw(xf5 + 5 * i);
        const ray_t r = ray_make(ray4 + 4 * i, infinite && infinite[i]);
        float t0 = 0.0f, t1 = 0.0f;
        const int hit = ray_shape(&s, &t, &r, &t0, &t1);
        out_hit[i] = (uint8_t)hit;
        out_times2[2 * i] = hit ? t0 : 0.0f;
        out_times2[2 * i + 1] = hit ? t1 : 0.0f;
    }
}
void rn_sweep_ray(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                  const float* xf5_end, const float* ray4, const uint8_t* infinite, float* out_times2, uint8_t* out_hit)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const shape_t s = shape_from(type[i], geom + 16 * i, cnt[i]);
        const xf_t ta = xf_raw(xf5 + 5 * i), tb = xf_raw(xf5_end + 5 * i);
        const ray_t r = ray_make(ray4 + 4 * i, infinite && infinite[i]);
        float t0 = 0.0f, t1 = 0.0f;
        const int hit = ray_sweep(&s, &ta, &tb, &r, &t0, &t1);
        out_hit[i] = (uint8_t)hit;
        out_times2[2 * i] = hit ? t0 : 0.0f;
        out_times2[2 * i + 1] = hit ? t1 : 0.0f;
    }
}
void rn_compute_aabb(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const shape_t s = shape_from(type[i], geom + 16 * i, cnt[i]);
        const xf_t t = xf_raw(xf5 + 5 * i);
        const box_t b = compute_aabb(&s, &t);
        out[4 * i] = b.minx; out[4 * i + 1] = b.miny; out[4 * i + 2] = b.maxx; out[4 * i + 3] = b.maxy;
    }
}
void rn_polygon_normals(uint32_t n, const float* geom, const uint8_t* cnt, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const shape_t s = shape_from(POLYGON, geom + 16 * i, cnt[i]);
        for (int k = 0; k < 8; ++k)
        {
            out[16 * i + 2 * k] = k < s.count ? s.n[k].x : 0.0f;
            out[16 * i + 2 * k + 1] = k < s.count ? s.n[k].y : 0.0f;
        }
    }
}
