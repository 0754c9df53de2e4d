// c2d_narrow_pp.cuh — warp-cooperative polygon narrowphase: 8 lanes ("octet") per polygon pair, 4 pairs per warp,
// world vertices staged in shared memory.
//
// Same arithmetic as poly_poly<> in c2d_narrow.cuh (reference src/collision/Collision.cpp:249-399, 668-690), so the
// results stay bit-identical, but the work of one pair is spread over 8 lanes instead of serialising in one:
//   * lane i transforms vertex i / normal i of both polygons (coalesced 8-byte loads) and stores the world vertices
//     in shared memory;
//   * SAT: lane i owns axis i of the first and axis i of the second polygon and projects all <= 16 vertices on them;
//     "separated" is an octet ballot, the minimum-overlap axis an octet arg-min on (overlap, axis order) — the
//     sequential "first strict minimum" of the reference;
//   * Sutherland-Hodgman: for each clip plane the <= 24 input vertices are handled 8 at a time, output slots come
//     from octet ballots (each vertex emits [intersection][itself] in the reference's order);
//   * centroids, the two lexicographically smallest contacts and the anchors are computed redundantly by all
//     lanes (no divergence), lane 0 of the octet writes the 84-byte record.
// Control flow is warp-uniform (fixed trip counts + predicates) so full-mask shuffles/ballots are always legal.
#pragma once

#include "c2d_narrow.cuh"

namespace c2d_dev
{

constexpr int kClipSlots = 24;

struct OctetShared
{
    V2 va[kMaxVerts];
    V2 vb[kMaxVerts];
    V2 clip[2][kClipSlots];
    V2 _pad; // 520-byte stride: the four octets of a warp fall into different banks
};

__device__ __forceinline__ uint32_t octet_ballot(bool pred, int octetInWarp)
{
    return (__ballot_sync(0xffffffffu, pred) >> (8 * octetInWarp)) & 0xffu;
}

// S1: polygon or capsule/segment (as the 2-gon of Collision.cpp:483-496), S2: polygon.  `alive` = this octet has a
// pair to test.  Returns the boolean test; with MANIFOLD also fills m (identically in all 8 lanes).
template <bool MANIFOLD>
__device__ __forceinline__ bool octet_poly_poly(bool alive, int sub, int octetInWarp, const ShapeRef& S1, const XF& x1,
                                                const ShapeRef& S2, const XF& x2, OctetShared& sh, Manifold& m,
                                                bool& overflow)
{
    const int n1 = alive ? (S1.type == T_POLYGON ? S1.count : 2) : 0;
    const int n2 = alive ? S2.count : 0;
    const int laneBase = octetInWarp * 8;

    // --- world vertices / normals: lane `sub` owns index `sub`
    V2 a1 = mk(0.0f, 0.0f), a2 = mk(0.0f, 0.0f);
    if (sub < n1)
    {
        V2 v, nrm;
        if (S1.type == T_POLYGON)
        {
            v = mk(S1.g[2 * sub], S1.g[2 * sub + 1]);
            nrm = mk(S1.g[16 + 2 * sub], S1.g[16 + 2 * sub + 1]);
        }
        else
        {
            const V2 c1 = g2(S1.g, 0), c2 = g2(S1.g, 1);
            const V2 axis = normalize(c2 - c1);
            const V2 n = mk(-axis.y, axis.x);
            v = sub == 0 ? c1 : c2;
            nrm = sub == 0 ? n : -n;
        }
        sh.va[sub] = apply(x1, v);
        a1 = rot(x1, nrm);
    }
    if (sub < n2)
    {
        sh.vb[sub] = apply(x2, mk(S2.g[2 * sub], S2.g[2 * sub + 1]));
        a2 = rot(x2, mk(S2.g[16 + 2 * sub], S2.g[16 + 2 * sub + 1]));
    }
    __syncwarp();

    // --- SAT: project every vertex of both polygons on this lane's two axes (projectPolygon, Collision.cpp:249-261)
    float min1A = 0.0f, max1A = 0.0f, min1B = 0.0f, max1B = 0.0f; // axis a1: polygon 1 (A) / polygon 2 (B)
    float min2A = 0.0f, max2A = 0.0f, min2B = 0.0f, max2B = 0.0f; // axis a2
#pragma unroll
    for (int k = 0; k < kMaxVerts; ++k)
    {
        if (k < n1)
        {
            const V2 v = sh.va[k];
            const float p1 = dot(a1, v), p2 = dot(a2, v);
            if (k == 0)
            {
                min1A = max1A = p1;
                min2A = max2A = p2;
            }
            else
            {
                min1A = cmin(min1A, p1);
                max1A = cmax(max1A, p1);
                min2A = cmin(min2A, p2);
                max2A = cmax(max2A, p2);
            }
        }
        if (k < n2)
        {
            const V2 v = sh.vb[k];
            const float p1 = dot(a1, v), p2 = dot(a2, v);
            if (k == 0)
            {
                min1B = max1B = p1;
                min2B = max2B = p2;
            }
            else
            {
                min1B = cmin(min1B, p1);
                max1B = cmax(max1B, p1);
                min2B = cmin(min2B, p2);
                max2B = cmax(max2B, p2);
            }
        }
    }
    const float ov1 = cmin(max1A, max1B) - cmax(min1A, min1B);
    const float ov2 = cmin(max2A, max2B) - cmax(min2A, min2B);
    const bool sepLane = (sub < n1 && ov1 < 0.0f) || (sub < n2 && ov2 < 0.0f);
    const bool separated = octet_ballot(sepLane, octetInWarp) != 0;
    const bool hit = alive && !separated;
    if (!MANIFOLD)
        return hit;

    // --- minimum-overlap axis: first strict minimum in the order A.normals[0..], B.normals[0..] (Collision.cpp:327-345)
    float bestOv = FLT_MAX;
    int bestIdx = 99;
    if (sub < n1)
    {
        bestOv = ov1;
        bestIdx = sub;
    }
    if (sub < n2 && (ov2 < bestOv || bestIdx == 99))
    {
        bestOv = ov2;
        bestIdx = 8 + sub;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
    {
        const float oOv = __shfl_xor_sync(0xffffffffu, bestOv, o);
        const int oIdx = __shfl_xor_sync(0xffffffffu, bestIdx, o);
        if (oOv < bestOv || (oOv == bestOv && oIdx < bestIdx))
        {
            bestOv = oOv;
            bestIdx = oIdx;
        }
    }
    // the reference starts from overlap = FLT_MAX, axis (1, 0) and only replaces them on a strict improvement
    float overlap = FLT_MAX;
    V2 smallestAxis = mk(1.0f, 0.0f);
    {
        const int owner = laneBase + (bestIdx & 7);
        const float a1x = __shfl_sync(0xffffffffu, a1.x, owner), a1y = __shfl_sync(0xffffffffu, a1.y, owner);
        const float a2x = __shfl_sync(0xffffffffu, a2.x, owner), a2y = __shfl_sync(0xffffffffu, a2.y, owner);
        if (bestIdx != 99 && bestOv < FLT_MAX)
        {
            overlap = bestOv;
            smallestAxis = bestIdx < 8 ? mk(a1x, a1y) : mk(a2x, a2y);
        }
    }

    // --- centroids (vertex averages accumulated in index order, Collision.cpp:352-362), all lanes redundantly
    V2 avgA = sh.va[0], avgB = sh.vb[0];
#pragma unroll
    for (int k = 1; k < kMaxVerts; ++k)
    {
        if (k < n1)
            avgA = avgA + sh.va[k];
        if (k < n2)
            avgB = avgB + sh.vb[k];
    }
    avgA = avgA / static_cast<float>(n1 > 0 ? n1 : 1);
    avgB = avgB / static_cast<float>(n2 > 0 ? n2 : 1);
    if (dot(avgB - avgA, smallestAxis) < 0.0f)
        smallestAxis = -smallestAxis;

    // --- polygonIntersection (Collision.cpp:308-322): A's world vertices clipped by every plane of B
    int cur = 0;
    int n = hit ? n1 : 0;
    if (sub < n)
        sh.clip[0][sub] = sh.va[sub];
    __syncwarp();
    const uint32_t ltMask = (1u << sub) - 1u;
    // warp-uniform trip counts: as many planes as the largest clip polygon of the four octets needs
    const int planes = __reduce_max_sync(0xffffffffu, hit ? n2 : 0);
#pragma unroll 1
    for (int j = 0; j < planes; ++j)
    {
        const bool planeOn = hit && j < n2 && n > 0;
        const float nx = __shfl_sync(0xffffffffu, a2.x, laneBase + j);
        const float ny = __shfl_sync(0xffffffffu, a2.y, laneBase + j);
        const V2 normal = mk(nx, ny);
        const V2 point = sh.vb[j < n2 ? j : 0];
        const V2* in = sh.clip[cur];
        V2* out = sh.clip[cur ^ 1];
        int base = 0;
        const int rounds = (__reduce_max_sync(0xffffffffu, planeOn ? n : 0) + 7) >> 3;
#pragma unroll 1
        for (int r = 0; r < rounds; ++r)
        {
            const int k = r * 8 + sub;
            const bool valid = planeOn && k < n;
            V2 curr = mk(0.0f, 0.0f), prev = mk(0.0f, 0.0f);
            float cd = 1.0f, pd = 1.0f;
            if (valid)
            {
                curr = in[k];
                prev = in[k == 0 ? n - 1 : k - 1];
                cd = dot(normal, curr - point);
                pd = dot(normal, prev - point);
            }
            const bool emitI = valid && ((cd <= 0.0f && pd > 0.0f) || (cd > 0.0f && pd <= 0.0f));
            const bool emitC = valid && cd <= 0.0f;
            const uint32_t bI = octet_ballot(emitI, octetInWarp);
            const uint32_t bC = octet_ballot(emitC, octetInWarp);
            int pos = base + __popc(bI & ltMask) + __popc(bC & ltMask);
            if (emitI)
            {
                const float f = pd / (pd - cd);
                if (pos < kClipSlots)
                    out[pos] = prev + (curr - prev) * f;
                else
                    overflow = true;
                ++pos;
            }
            if (emitC)
            {
                if (pos < kClipSlots)
                    out[pos] = curr;
                else
                    overflow = true;
            }
            base += __popc(bI) + __popc(bC);
        }
        __syncwarp();
        if (planeOn)
        {
            n = base < kClipSlots ? base : kClipSlots;
            cur ^= 1;
        }
    }

    manifold_clear(m);
    if (!hit)
        return false;
    m.normal = smallestAxis;
    if (n == 0)
    {
        set_single_contact(m, smallestAxis, (avgA + avgB) * 0.5f, x1, x2, -overlap);
        return true;
    }
    // the two lexicographically smallest survivors (std::sort + first two, Collision.cpp:385-396)
    const V2* pts = sh.clip[cur];
    int i0 = 0;
    for (int k = 1; k < n; ++k)
        if (lex_less(pts[k], pts[i0]))
            i0 = k;
    const int count = n < 2 ? n : 2;
    int i1 = -1;
    if (count == 2)
        for (int k = 0; k < n; ++k)
            if (k != i0 && (i1 < 0 || lex_less(pts[k], pts[i1])))
                i1 = k;
    for (int k = 0; k < count; ++k)
    {
        const V2 c = pts[k == 0 ? i0 : i1];
        m.point[k] = c;
        m.anchorA[k] = to_local(x1, c);
        m.anchorB[k] = to_local(x2, c);
        m.sep[k] = -overlap;
    }
    m.count = count;
    return true;
}

} // namespace c2d_dev
