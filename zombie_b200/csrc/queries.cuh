// Device-side geometric queries over the wide quantised BVH (wide_bvh.h): closest point, closest
// ray hit / occlusion, stochastic primitive sampling inside a sphere, and the star radius for
// Neumann / Robin boundaries. One thread runs one query; the traversal stack lives in local
// memory and children are ordered by a register sorting network so the nearest child is popped
// first. These replace the five stack-based traversals of fcpw::Bvh
// (fcpw/aggregates/bvh.inl:656-771, 1141-1288, 1291-1418, 1436-1596) and
// ReflectanceBvh::computeSquaredStarRadius (zombie/utils/reflectance_boundary_bvh/bvh.h:370-506).
#pragma once

#include "vec.cuh"

// debugging hook of tools/gpu_star_debug.cu (reports the intermediate values of one primitive's star-radius test)
#ifndef ZB_PRIM_TRACE
#define ZB_PRIM_TRACE(prim, what, a, b, c, d)
#endif

namespace zb {

struct BvhView {
    const uint4 *nodes;        // 5 per node
    const float4 *prims;       // 3D: 3 per triangle (v0|index, v1, v2); 2D: 1 per segment (ax, ay, bx, by)
    const int *primIndex;      // leaf order -> caller's primitive index
    const float4 *cones;       // 2 per child slot (16 per node); reflecting boundaries only
    const float4 *adjacency;   // 3 per primitive; reflecting boundaries only
    int nPrims;
    int snchSign;              // 1: Robin handler convention; -1: FCPW SNCH (Neumann handler) convention
};

// optional traversal counters (build with -DZB_STATS): [2q] = nodes visited, [2q+1] = primitives tested,
// [10+q] = number of queries, q = 0 closest point, 1 ray, 2 occlusion ray, 3 sphere sample, 4 star radius
#if defined(ZB_HOST_STATS)
// tools/host_probe.cu: per-query records on the host pass, nothing on the device pass
#if defined(__CUDA_ARCH__)
#define ZB_COUNT_DECL
#define ZB_COUNT_NODE
#define ZB_COUNT_PRIM
#define ZB_COUNT_FLUSH(q)
#else
void zbHostRecord(int query, int nodes, int prims);
extern long long g_zbClass[24]; // star query: [0] items with silhouette tests, [1] without, [2] prims early-out by distance, [3] prims fully tested
#define ZB_CLASS(i) g_zbClass[i]++;
#define ZB_COUNT_DECL int zbNodes = 0, zbPrims = 0;
#define ZB_COUNT_NODE zbNodes++;
#define ZB_COUNT_PRIM zbPrims++;
#define ZB_COUNT_FLUSH(q) zbHostRecord((q), zbNodes, zbPrims);
#endif
#if !defined(ZB_CLASS)
#define ZB_CLASS(i)
#endif
#elif defined(ZB_STATS)
__device__ unsigned long long g_zbCounters[16];
#define ZB_COUNT_DECL int zbNodes = 0, zbPrims = 0;
#define ZB_COUNT_NODE zbNodes++;
#define ZB_COUNT_PRIM zbPrims++;
#define ZB_COUNT_FLUSH(q) { atomicAdd(&g_zbCounters[2*(q)], (unsigned long long)zbNodes); atomicAdd(&g_zbCounters[2*(q) + 1], (unsigned long long)zbPrims); atomicAdd(&g_zbCounters[10 + (q)], 1ull); }
#else
#define ZB_COUNT_DECL
#define ZB_COUNT_NODE
#define ZB_COUNT_PRIM
#define ZB_COUNT_FLUSH(q)
#endif
#if !defined(ZB_CLASS)
#define ZB_CLASS(i)
#endif

constexpr int kStackSize = 64;
constexpr uint32_t kLeafFlag = 0x80000000u;

// status bits raised by kernels (read back by the host after each launch)
enum : int { kStatusStackOverflow = 1 };

struct NodeRegs {
    uint4 q[5];
};

ZB_DEV void loadNode(const uint4 *nodes, uint32_t idx, NodeRegs& n)
{
    const uint4 *p = nodes + (size_t)idx*5;
#pragma unroll
    for (int i = 0; i < 5; i++) n.q[i] = zb_ldg(p + i);
}

ZB_DEV float byteF(uint32_t w, int i) { return (float)((w >> (8*i)) & 0xffu); }

// byte c (0..7) of the word pair (w0: children 0..3, w1: children 4..7) for a child slot only known at run time: the
// child loops of the traversals are NOT unrolled (one copy of the child test instead of eight — the walk kernel is
// bound by instruction fetch, see profiles/README.md), so the slot is a loop variable
ZB_DEV uint32_t byte8(uint32_t w0, uint32_t w1, int c)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(w0, w1, (uint32_t)c) & 0xffu;
#else
    return ((c < 4 ? w0 : w1) >> (8*(c & 3))) & 0xffu;
#endif
}

// child loops: "unroll 1" keeps one copy of the body. A/B builds: -DZB_CHILD_UNROLL=2 or 4 (some instruction-level
// parallelism between children at 2x / 4x the code), -DZB_UNROLL_CHILD_LOOPS (full unrolling, the original shape)
#if defined(ZB_UNROLL_CHILD_LOOPS)
#define ZB_CHILD_LOOP _Pragma("unroll")
#else
#ifndef ZB_CHILD_UNROLL
#define ZB_CHILD_UNROLL 1
#endif
#define ZB_PRAGMA_(x) _Pragma(#x)
#define ZB_PRAGMA(x) ZB_PRAGMA_(x)
#define ZB_CHILD_LOOP ZB_PRAGMA(unroll ZB_CHILD_UNROLL)
#endif

// word holding the quantised coordinate bytes of children 0..3 / 4..7: axis 0..2, side 0 = lo, 1 = hi
ZB_DEV uint32_t qword(const NodeRegs& n, int axis, int side, int half)
{
    int w = 8 + 6*side + 2*axis + half; // index into the 20 words of the node
    const uint32_t *words = reinterpret_cast<const uint32_t *>(&n.q[0]);
    return words[w];
}

template <int DIM>
struct NodeFrame {
    float o[DIM], s[DIM];
};

template <int DIM>
ZB_DEV NodeFrame<DIM> nodeFrame(const NodeRegs& n)
{
    NodeFrame<DIM> f;
    f.o[0] = zb_uint_as_float(n.q[0].x);
    f.o[1] = zb_uint_as_float(n.q[0].y);
    if (DIM == 3) f.o[DIM - 1] = zb_uint_as_float(n.q[0].z);
    uint32_t e = n.q[0].w;
    f.s[0] = zb_uint_as_float((e & 0xffu) << 23);
    f.s[1] = zb_uint_as_float(((e >> 8) & 0xffu) << 23);
    if (DIM == 3) f.s[DIM - 1] = zb_uint_as_float(((e >> 16) & 0xffu) << 23);
    return f;
}

template <int DIM>
ZB_DEV void childBox(const NodeRegs& n, const NodeFrame<DIM>& f, int c, float lo[DIM], float hi[DIM])
{
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        lo[k] = fmaf(f.s[k], (float)byte8(qword(n, k, 0, 0), qword(n, k, 0, 1), c), f.o[k]);
        hi[k] = fmaf(f.s[k], (float)byte8(qword(n, k, 1, 0), qword(n, k, 1, 1), c), f.o[k]);
    }
}

ZB_DEV uint32_t childMeta(const NodeRegs& n, int c) { return byte8(n.q[1].z, n.q[1].w, c); }

// stack reference of child slot c: internal -> node index; leaf -> flag | count << 28 | first primitive
ZB_DEV uint32_t childRef(const NodeRegs& n, uint32_t meta)
{
    uint32_t count = meta >> 5;
    if (count) return kLeafFlag | (count << 28) | (n.q[1].y + (meta & 31u));
    return n.q[1].x + (meta & 31u) - 1u;
}

ZB_DEV void cswap(uint32_t& a, uint32_t& b) { uint32_t lo = zb_umin(a, b); b = zb_umax(a, b); a = lo; }

// ascending sort of 8 keys (19 compare-exchanges)
ZB_DEV void sort8(uint32_t k[8])
{
    cswap(k[0], k[1]); cswap(k[2], k[3]); cswap(k[4], k[5]); cswap(k[6], k[7]);
    cswap(k[0], k[2]); cswap(k[1], k[3]); cswap(k[4], k[6]); cswap(k[5], k[7]);
    cswap(k[1], k[2]); cswap(k[5], k[6]); cswap(k[0], k[4]); cswap(k[3], k[7]);
    cswap(k[1], k[5]); cswap(k[2], k[6]);
    cswap(k[1], k[4]); cswap(k[3], k[6]);
    cswap(k[2], k[4]); cswap(k[3], k[5]);
    cswap(k[3], k[4]);
}

// inserts key t into the ascending list k[0..8) (padded with 0xffffffff; the largest entry falls off the end)
ZB_DEV void insertKey(uint32_t k[8], uint32_t t)
{
#pragma unroll
    for (int i = 0; i < 8; i++) cswap(k[i], t);
}

ZB_DEV float keyValue(uint32_t key) { return zb_uint_as_float(key & ~7u); }

// ------------------------------------------------------------------------------------------
// primitives
// ------------------------------------------------------------------------------------------

// ---- uncontracted arithmetic ---------------------------------------------------------------
// The primitive tests below use explicit round-to-nearest multiplies and adds (never fused into
// FMAs) in the evaluation order of the reference's fixed-size vector library (dot of 3 terms =
// x0*y0 + (x1*y1 + x2*y2)), so that for the same primitive and query they return bit-identical
// distances, parameters and points to FCPW's CPU build (-O3, no FMA contraction on x86-64).
ZB_DEV float xdot(const Vec<3>& a, const Vec<3>& b)
{
    return zb_fadd(zb_fmul(a[0], b[0]), zb_fadd(zb_fmul(a[1], b[1]), zb_fmul(a[2], b[2])));
}
ZB_DEV float xdot(const Vec<2>& a, const Vec<2>& b) { return zb_fadd(zb_fmul(a[0], b[0]), zb_fmul(a[1], b[1])); }
template <int D> ZB_DEV float xnorm(const Vec<D>& a) { return sqrtf(xdot(a, a)); }
template <int D> ZB_DEV Vec<D> xscale(const Vec<D>& a, float s) { Vec<D> r; for (int i = 0; i < D; i++) r[i] = zb_fmul(a[i], s); return r; }
template <int D> ZB_DEV Vec<D> xadd(const Vec<D>& a, const Vec<D>& b) { Vec<D> r; for (int i = 0; i < D; i++) r[i] = zb_fadd(a[i], b[i]); return r; }
// a*b - c*d without contraction
ZB_DEV float xdiffprod(float a, float b, float c, float d) { return zb_fsub(zb_fmul(a, b), zb_fmul(c, d)); }
ZB_DEV Vec<3> xcross(const Vec<3>& a, const Vec<3>& b)
{
    return make3(xdiffprod(a[1], b[2], a[2], b[1]), xdiffprod(a[2], b[0], a[0], b[2]), xdiffprod(a[0], b[1], a[1], b[0]));
}
ZB_DEV float xcross2(const Vec<2>& a, const Vec<2>& b) { return xdiffprod(a[0], b[1], a[1], b[0]); }

// closest point on triangle (pa, pb, pc) to x by Voronoi-region classification (Ericson, Real-Time
// Collision Detection 5.1.5), the same classification order and comparisons as
// fcpw/geometry/triangles.inl:374-458 so that region ties resolve identically.
ZB_DEV float closestPointTriangle(const Vec<3>& pa, const Vec<3>& pb, const Vec<3>& pc, const Vec<3>& x,
                                  Vec<3>& pt, float& bu, float& bv)
{
    Vec<3> ab = pb - pa, ac = pc - pa, ax = x - pa;
    float d1 = xdot(ab, ax), d2 = xdot(ac, ax);
    if (d1 <= 0.0f && d2 <= 0.0f) { bu = 1.0f; bv = 0.0f; pt = pa; return xnorm(x - pt); }

    Vec<3> bx = x - pb;
    float d3 = xdot(ab, bx), d4 = xdot(ac, bx);
    if (d3 >= 0.0f && d4 <= d3) { bu = 0.0f; bv = 1.0f; pt = pb; return xnorm(x - pt); }

    Vec<3> cx = x - pc;
    float d5 = xdot(ab, cx), d6 = xdot(ac, cx);
    if (d6 >= 0.0f && d5 <= d6) { bu = 0.0f; bv = 0.0f; pt = pc; return xnorm(x - pt); }

    float vc = xdiffprod(d1, d4, d3, d2);
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        float v = d1/(d1 - d3);
        bu = 1.0f - v; bv = v;
        pt = xadd(pa, xscale(ab, v));
        return xnorm(x - pt);
    }

    float vb = xdiffprod(d5, d2, d1, d6);
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        float w = d2/(d2 - d6);
        bu = 1.0f - w; bv = 0.0f;
        pt = xadd(pa, xscale(ac, w));
        return xnorm(x - pt);
    }

    float va = xdiffprod(d3, d6, d5, d4);
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3)/((d4 - d3) + (d5 - d6));
        bu = 0.0f; bv = 1.0f - w;
        pt = xadd(pb, xscale(pc - pb, w));
        return xnorm(x - pt);
    }

    float denom = 1.0f/(zb_fadd(zb_fadd(va, vb), vc));
    float v = zb_fmul(vb, denom), w = zb_fmul(vc, denom);
    bu = 1.0f - v - w; bv = v;
    pt = xadd(xadd(pa, xscale(ab, v)), xscale(ac, w));
    return xnorm(x - pt);
}

// fcpw/geometry/line_segments.inl:200-227
template <int DIM>
ZB_DEV float closestPointSegment(const Vec<DIM>& pa, const Vec<DIM>& pb, const Vec<DIM>& x, Vec<DIM>& pt, float& t)
{
    Vec<DIM> u = pb - pa, v = x - pa;
    float c1 = xdot(u, v);
    if (c1 <= 0.0f) { pt = pa; t = 0.0f; return xnorm(x - pt); }
    float c2 = xdot(u, u);
    if (c2 <= c1) { pt = pb; t = 1.0f; return xnorm(x - pt); }
    t = c1/c2;
    pt = xadd(pa, xscale(u, t));
    return xnorm(x - pt);
}

template <int DIM> struct Prim;

template <> struct Prim<3> {
    Vec<3> a, b, c;
    int index;
    float maxCoeff;   // reflecting boundaries: the triangle's maximum Robin coefficient and its adjacency flags, copies of
    uint32_t adjFlags; // PrimAdjacency's (wide_bvh.cpp) riding in v1.w / v2.w so that the star filter stage reads 48 bytes in all
    ZB_DEV void load(const BvhView& bvh, uint32_t i) {
        float4 v0 = zb_ldg(bvh.prims + (size_t)3*i), v1 = zb_ldg(bvh.prims + (size_t)3*i + 1), v2 = zb_ldg(bvh.prims + (size_t)3*i + 2);
        a = make3(v0.x, v0.y, v0.z); b = make3(v1.x, v1.y, v1.z); c = make3(v2.x, v2.y, v2.z);
        index = zb_float_as_int(v0.w);
        maxCoeff = v1.w;
        adjFlags = zb_float_as_uint(v2.w);
    }
    ZB_DEV float closest(const Vec<3>& x, Vec<3>& pt, float& u, float& v) const { return closestPointTriangle(a, b, c, x, pt, u, v); }
    // unnormalised normal (b-a)x(c-a): fcpw Triangle::normal()
    ZB_DEV Vec<3> rawNormal() const { return xcross(b - a, c - a); }
    ZB_DEV float area() const { return 0.5f*norm(rawNormal()); } // triangles.inl:34-37
    // Moller-Trumbore with the parallel cutoff and inclusive bounds of triangles.inl:222-260
    ZB_DEV bool intersect(const Vec<3>& o, const Vec<3>& d, float tMax, float& t, float& bv, float& bw) const {
        Vec<3> v1 = b - a, v2 = c - a;
        Vec<3> p = xcross(d, v2);
        float det = xdot(v1, p);
        if (fabsf(det) <= kEpsilon) return false;
        float invDet = 1.0f/det;
        Vec<3> s = o - a;
        float v = zb_fmul(xdot(s, p), invDet);
        if (v < 0.0f || v > 1.0f) return false;
        Vec<3> q = xcross(s, v1);
        float w = zb_fmul(xdot(d, q), invDet);
        if (w < 0.0f || zb_fadd(v, w) > 1.0f) return false;
        float tt = zb_fmul(xdot(v2, q), invDet);
        if (tt >= 0.0f && tt <= tMax) { t = tt; bv = v; bw = w; return true; }
        return false;
    }
    ZB_DEV Vec<3> pointAt(float bv, float bw) const { return xadd(xadd(a, xscale(b - a, bv)), xscale(c - a, bw)); }
    // uniform point: triangles.inl:113-131 (uses randNums[1], randNums[2]); returns pdf 2/|n|
    ZB_DEV float sample(const Vec<3>& rand, Vec<3>& p, Vec<3>& n) const {
        n = rawNormal();
        float area2 = norm(n);
        float u1 = sqrtf(rand[1]), u2 = rand[2];
        float u = 1.0f - u1, v = u2*u1, w = 1.0f - u - v;
        p = a*u + b*v + c*w;
        n = n/area2;
        return 2.0f/area2;
    }
};

template <> struct Prim<2> {
    Vec<2> a, b;
    int index;
    ZB_DEV void load(const BvhView& bvh, uint32_t i) {
        float4 v = zb_ldg(bvh.prims + i);
        a = make2(v.x, v.y); b = make2(v.z, v.w);
        index = zb_ldg(bvh.primIndex + i);
    }
    ZB_DEV float closest(const Vec<2>& x, Vec<2>& pt, float& u, float& v) const { v = 0.0f; return closestPointSegment<2>(a, b, x, pt, u); }
    ZB_DEV Vec<2> rawNormal() const { Vec<2> s = b - a; return make2(s[1], -s[0]); }
    ZB_DEV float area() const { return norm(b - a); }
    // line_segments.inl:140-175
    ZB_DEV bool intersect(const Vec<2>& o, const Vec<2>& d, float tMax, float& t, float& bs, float& unused) const {
        Vec<2> u = a - o, v = b - a;
        float dv = xcross2(d, v);
        if (fabsf(dv) <= kEpsilon) return false;
        float ud = xcross2(u, d);
        float s = ud/dv;
        if (s >= 0.0f && s <= 1.0f) {
            float uv = xcross2(u, v);
            float tt = uv/dv;
            if (tt >= 0.0f && tt <= tMax) { t = tt; bs = s; unused = 0.0f; return true; }
        }
        return false;
    }
    ZB_DEV Vec<2> pointAt(float s, float) const { return xadd(a, xscale(b - a, s)); }
    // line_segments.inl:83-98 (uses randNums[1])
    ZB_DEV float sample(const Vec<2>& rand, Vec<2>& p, Vec<2>& n) const {
        Vec<2> s = b - a;
        float len = norm(s);
        p = a + s*rand[1];
        n = make2(s[1], -s[0])/len;
        return 1.0f/len;
    }
};

ZB_DEV void setThird(Prim<3>& p, const float4& v) { p.c = make3(v.x, v.y, v.z); }
ZB_DEV void setThird(Prim<2>&, const float4&) {}

template <int DIM>
ZB_DEV Vec<DIM> unitNormal(const Prim<DIM>& p) { Vec<DIM> n = p.rawNormal(); return n/xnorm(n); }

#ifndef ZB_CLOSEST_PLANE_FILTER
#define ZB_CLOSEST_PLANE_FILTER 1 // 0: run the closest-point test on every primitive of a visited leaf (A/B builds)
#endif

// Cheap first stage of the closest-point test (3D): a triangle whose PLANE is farther than the search radius cannot
// hold a closer point, nor can one whose bounding sphere is. False = skip it. 1e-5 of slack keeps exact ties (which the reference resolves by primitive
// order, triangles.inl:479) inside the full test.
template <int DIM>
ZB_DEV bool primClosestNeedsFullTest(const float4 *prims, uint32_t i, const Vec<DIM>& x, float r2)
{
    if (DIM != 3 || !ZB_CLOSEST_PLANE_FILTER) return true;
    float4 v0 = zb_ldg(prims + (size_t)3*i), v1 = zb_ldg(prims + (size_t)3*i + 1), v2 = zb_ldg(prims + (size_t)3*i + 2);
    float e1[3] = {v1.x - v0.x, v1.y - v0.y, v1.z - v0.z}, e2[3] = {v2.x - v0.x, v2.y - v0.y, v2.z - v0.z};
    float n[3] = {e1[1]*e2[2] - e1[2]*e2[1], e1[2]*e2[0] - e1[0]*e2[2], e1[0]*e2[1] - e1[1]*e2[0]};
    float s = n[0]*(x[0] - v0.x) + n[1]*(x[1] - v0.y) + n[2]*(x[DIM - 1] - v0.z);
    float nn = n[0]*n[0] + n[1]*n[1] + n[2]*n[2];
    if (s*s > nn*(r2*1.00001f)) return false;
    // ... nor can one whose bounding sphere (about the centroid) lies beyond it: |x - c| > R + r, squared
    float c[3] = {(v0.x + v1.x + v2.x)*(1.0f/3.0f), (v0.y + v1.y + v2.y)*(1.0f/3.0f), (v0.z + v1.z + v2.z)*(1.0f/3.0f)};
    float a0 = v0.x - c[0], a1 = v0.y - c[1], a2 = v0.z - c[2], b0 = v1.x - c[0], b1 = v1.y - c[1], b2 = v1.z - c[2];
    float g0 = v2.x - c[0], g1 = v2.y - c[1], g2 = v2.z - c[2];
    float R2 = fmaxf(fmaxf(a0*a0 + a1*a1 + a2*a2, b0*b0 + b1*b1 + b2*b2), g0*g0 + g1*g1 + g2*g2)*1.00001f;
    float dx = x[0] - c[0], dy = x[1] - c[1], dz = x[DIM - 1] - c[2];
    float dc2 = dx*dx + dy*dy + dz*dz, rr = r2*1.00001f;
    // r2 may be +inf (unbounded search): inf*0 = NaN and inf compare false either way -> kept
    return !(dc2 > (R2 + rr + 2.0f*zb_sqrt_fast(R2*rr))*1.00001f);
}

// Closest point on primitive i (leaf order): ONE out-of-line copy per kernel, shared by the private depth-first phase,
// the pooled primitive rounds and the final re-evaluation of the winner (coop.cuh) and by sampleSphere. Returns
// (pt, distance) packed in a float4. Build with -DZB_INLINE_PRIM_TESTS to inline it at every call site instead.
template <int DIM>
#if defined(ZB_INLINE_PRIM_TESTS)
ZB_DEV
#else
ZB_DEV_COLD
#endif
float4 primClosestCall(const float4 *prims, uint32_t i, Vec<DIM> x)
{
    Vec<DIM> pa, pb, pt;
    float d, u, v;
    if (DIM == 3) {
        float4 v0 = zb_ldg(prims + (size_t)3*i), v1 = zb_ldg(prims + (size_t)3*i + 1), v2 = zb_ldg(prims + (size_t)3*i + 2);
        Prim<DIM> prim;
        prim.a[0] = v0.x; prim.a[1] = v0.y; prim.a[DIM - 1] = v0.z;
        prim.b[0] = v1.x; prim.b[1] = v1.y; prim.b[DIM - 1] = v1.z;
        setThird(prim, v2);
        d = prim.closest(x, pt, u, v);
    } else {
        float4 v0 = zb_ldg(prims + i);
        Prim<DIM> prim;
        prim.a[0] = v0.x; prim.a[1] = v0.y;
        prim.b[0] = v0.z; prim.b[1] = v0.w;
        d = prim.closest(x, pt, u, v);
    }
    return make_float4(pt[0], pt[1], DIM == 3 ? pt[DIM - 1] : 0.0f, d);
}

// ------------------------------------------------------------------------------------------
// closest point
// ------------------------------------------------------------------------------------------

template <int DIM>
struct ClosestHit {
    float d;        // distance (kMaxFloat when nothing found)
    int prim;       // leaf-order primitive, -1 when nothing found
    Vec<DIM> p;
};

// squared distance bounds from x to the box [lo, hi]
template <int DIM>
ZB_DEV void boxDistance2(const float lo[DIM], const float hi[DIM], const Vec<DIM>& x, float& d2min, float& d2max)
{
    d2min = 0.0f; d2max = 0.0f;
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        float u = lo[k] - x[k], v = x[k] - hi[k];
        float dn = fmaxf(fmaxf(u, v), 0.0f);
        float df = fminf(u, v);
        d2min = fmaf(dn, dn, d2min);
        d2max = fmaf(df, df, d2max);
    }
}

// Depth-first core of the closest-point query: pops up to `budget` stack entries (nearest child first) and
// returns with the unfinished part of the traversal still on the stack (sp > 0) when the budget runs out.
// hit / r2 carry the best primitive and squared search radius found so far.
template <int DIM>
ZB_DEV void closestPointDfs(const BvhView& bvh, const Vec<DIM>& x, float& r2, ClosestHit<DIM>& hit, uint2 *stack, int& sp,
                            int budget, int *status)
{
    ZB_COUNT_DECL
    while (sp > 0 && budget-- > 0) {
        uint2 e = stack[--sp];
        if (keyValue(e.x) > r2) continue;
        if (e.y & kLeafFlag) {
            uint32_t first = e.y & 0x0fffffffu, count = (e.y >> 28) & 7u;
            for (uint32_t i = 0; i < count; i++) {
                ZB_COUNT_PRIM
                if (!primClosestNeedsFullTest<DIM>(bvh.prims, first + i, x, r2)) continue;
                float4 cp = primClosestCall<DIM>(bvh.prims, first + i, x);
                float d = cp.w, d2 = d*d;
                if (d2 <= r2) { // triangles.inl:479: inclusive, later primitives replace equal distances
                    r2 = d2;
                    hit.d = d;
                    hit.prim = (int)(first + i);
                    hit.p[0] = cp.x; hit.p[1] = cp.y;
                    if (DIM == 3) hit.p[DIM - 1] = cp.z;
                }
            }
            continue;
        }

        NodeRegs n;
        loadNode(bvh.nodes, e.y, n);
        ZB_COUNT_NODE
        NodeFrame<DIM> f = nodeFrame<DIM>(n);
        uint32_t keys[8];
#pragma unroll
        for (int c = 0; c < 8; c++) keys[c] = 0xffffffffu;
        ZB_CHILD_LOOP
        for (int c = 0; c < 8; c++) {
            if (childMeta(n, c) == 0u) continue;
            float lo[DIM], hi[DIM], d2min, d2max;
            childBox<DIM>(n, f, c, lo, hi);
            boxDistance2<DIM>(lo, hi, x, d2min, d2max);
            r2 = fminf(r2, d2max);
            insertKey(keys, (zb_float_as_uint(d2min) & ~7u) | (uint32_t)c);
        }
#pragma unroll
        for (int c = 7; c >= 0; c--) {
            uint32_t key = keys[c];
            if (key != 0xffffffffu && keyValue(key) <= r2) {
                if (sp < kStackSize) stack[sp++] = make_uint2(key, childRef(n, childMeta(n, key & 7u)));
                else if (status) zb_atomic_or(status, kStatusStackOverflow);
            }
        }
    }
    ZB_COUNT_FLUSH(0)
}

// Aggregate::findClosestPoint(BoundingSphere(x, r2)) — bvh.inl:1291-1418. The search radius also
// shrinks by each child's farthest-corner distance (bvh.inl:1339-1343), which bounds the stack.
template <int DIM>
ZB_DEV bool closestPoint(const BvhView& bvh, const Vec<DIM>& x, float r2, ClosestHit<DIM>& hit, int *status)
{
    uint2 stack[kStackSize];
    int sp = 0;
    stack[sp++] = make_uint2(0u, 0u);
    hit.d = kMaxFloat;
    hit.prim = -1;
    hit.p = x;
    closestPointDfs<DIM>(bvh, x, r2, hit, stack, sp, 0x7fffffff, status);
    return hit.prim >= 0;
}

// Which primitive a closest-point query REPORTS when several are exactly equidistant (the closest point lies on a
// shared edge or vertex: a quarter of random queries on a coarse mesh). FCPW's acceptance test is d*d <= r2
// (triangles.inl:479, line_segments.inl), so a primitive visited later replaces an equidistant earlier one and the
// winner depends on the visiting order: the tree topology for the BVH (bvh.inl:1331-1335), the input order for the
// brute-force Baseline aggregate (baseline.inl:349-372), which therefore returns the LARGEST input index among the
// primitives whose squared distance d*d equals the minimum. That is the rule here: a second traversal with the search
// radius fixed at the minimum visits every primitive that can tie (boxes only prune beyond the radius) and keeps the
// largest caller index; distance and point are that primitive's. Only the batched entry point pays for it — the walk
// reads neither index nor normal.
template <int DIM>
ZB_DEV void closestTieBreak(const BvhView& bvh, const Vec<DIM>& x, ClosestHit<DIM>& hit, int *status)
{
    if (hit.prim < 0) return;
    float r2 = hit.d*hit.d;
    int bestCaller = zb_ldg(bvh.primIndex + hit.prim);
    uint2 stack[kStackSize];
    int sp = 0;
    stack[sp++] = make_uint2(0u, 0u);
    while (sp > 0) {
        uint2 e = stack[--sp];
        if (e.y & kLeafFlag) {
            uint32_t first = e.y & 0x0fffffffu, count = (e.y >> 28) & 7u;
            for (uint32_t i = 0; i < count; i++) {
                if ((int)(first + i) == hit.prim) continue;
                float4 cp = primClosestCall<DIM>(bvh.prims, first + i, x);
                float d2 = cp.w*cp.w;
                if (d2 > r2) continue;
                int caller = zb_ldg(bvh.primIndex + first + i);
                if (d2 < r2 || caller > bestCaller) { // d2 < r2 cannot happen after a complete search; kept for robustness
                    r2 = d2;
                    bestCaller = caller;
                    hit.d = cp.w;
                    hit.prim = (int)(first + i);
                    hit.p[0] = cp.x; hit.p[1] = cp.y;
                    if (DIM == 3) hit.p[DIM - 1] = cp.z;
                }
            }
            continue;
        }
        NodeRegs n;
        loadNode(bvh.nodes, e.y, n);
        NodeFrame<DIM> f = nodeFrame<DIM>(n);
        ZB_CHILD_LOOP
        for (int c = 0; c < 8; c++) {
            uint32_t meta = childMeta(n, c);
            if (meta == 0u) continue;
            float lo[DIM], hi[DIM], d2min, d2max;
            childBox<DIM>(n, f, c, lo, hi);
            boxDistance2<DIM>(lo, hi, x, d2min, d2max);
            // 8 ulps of slack as in the search itself (keys are rounded down): the box bound and the primitive's distance
            // are rounded differently
            if (keyValue(zb_float_as_uint(d2min) & ~7u) > r2) continue;
            if (sp < kStackSize) stack[sp++] = make_uint2(0u, childRef(n, meta));
            else if (status) zb_atomic_or(status, kStatusStackOverflow);
        }
    }
}

// ------------------------------------------------------------------------------------------
// rays
// ------------------------------------------------------------------------------------------

template <int DIM>
struct RayHit {
    float t;
    int prim;     // leaf order, -1 on miss
    float bv, bw; // barycentric parameters of the hit
};

// Aggregate::intersect(Ray) closest hit or occlusion — bvh.inl:656-771; slab test of
// fcpw/core/bounding_volumes.h:99-114 on the decoded child boxes.
template <int DIM, bool ANY_HIT>
ZB_DEV bool intersectRay(const BvhView& bvh, const Vec<DIM>& o, const Vec<DIM>& d, float tMax, RayHit<DIM>& hit, int *status)
{
    uint2 stack[kStackSize];
    int sp = 0;
    stack[sp++] = make_uint2(0u, 0u);
    hit.t = kMaxFloat;
    hit.prim = -1;
    hit.bv = hit.bw = 0.0f;
    float invD[DIM];
#pragma unroll
    for (int k = 0; k < DIM; k++) invD[k] = 1.0f/d[k];
    ZB_COUNT_DECL

    while (sp > 0) {
        uint2 e = stack[--sp];
        if (keyValue(e.x) > tMax) continue;
        if (e.y & kLeafFlag) {
            uint32_t first = e.y & 0x0fffffffu, count = (e.y >> 28) & 7u;
            for (uint32_t i = 0; i < count; i++) {
                Prim<DIM> prim;
                prim.load(bvh, first + i);
                float t, bv, bw;
                ZB_COUNT_PRIM
                if (prim.intersect(o, d, tMax, t, bv, bw)) {
                    if (ANY_HIT) { hit.prim = (int)(first + i); hit.t = t; ZB_COUNT_FLUSH(2) return true; }
                    tMax = fminf(tMax, t);
                    hit.t = t; hit.prim = (int)(first + i); hit.bv = bv; hit.bw = bw;
                }
            }
            continue;
        }

        NodeRegs n;
        loadNode(bvh.nodes, e.y, n);
        ZB_COUNT_NODE
        NodeFrame<DIM> f = nodeFrame<DIM>(n);
        uint32_t keys[8];
#pragma unroll
        for (int c = 0; c < 8; c++) keys[c] = 0xffffffffu;
        ZB_CHILD_LOOP
        for (int c = 0; c < 8; c++) {
            if (childMeta(n, c) == 0u) continue;
            float lo[DIM], hi[DIM];
            childBox<DIM>(n, f, c, lo, hi);
            float tNear = 0.0f, tFar = tMax;
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                float t0 = (lo[k] - o[k])*invD[k], t1 = (hi[k] - o[k])*invD[k];
                tNear = fmaxf(tNear, fminf(t0, t1));
                tFar = fminf(tFar, fmaxf(t0, t1));
            }
            if (tNear <= tFar) {
                if (ANY_HIT) {
                    // an occlusion test only asks WHETHER something is hit: the visiting order cannot change the answer,
                    // so the children go to the stack as they come, without the sorted insertion of the closest-hit search
                    if (sp < kStackSize) stack[sp++] = make_uint2(zb_float_as_uint(tNear) & ~7u, childRef(n, childMeta(n, c)));
                    else if (status) zb_atomic_or(status, kStatusStackOverflow);
                } else {
                    insertKey(keys, (zb_float_as_uint(tNear) & ~7u) | (uint32_t)c);
                }
            }
        }
        if (!ANY_HIT) {
#pragma unroll
            for (int c = 7; c >= 0; c--) {
                uint32_t key = keys[c];
                if (key != 0xffffffffu) {
                    if (sp < kStackSize) stack[sp++] = make_uint2(key, childRef(n, childMeta(n, key & 7u)));
                    else if (status) zb_atomic_or(status, kStatusStackOverflow);
                }
            }
        }
    }
    ZB_COUNT_FLUSH(ANY_HIT ? 2 : 1)
    return hit.prim >= 0;
}

// Aggregate::intersect(Ray, interactions, checkForOcclusion = false, recordAllHits = true) — bvh.inl:890-1028: every
// primitive the ray hits within tMax, sorted by distance, with adjacent entries merged when their distances differ by less
// than 1e-6 and their points by less than 1e-3 (Interaction::operator==, interaction.h:20-25: a ray through a shared edge
// or vertex counts once). Returns the number of distinct hits; GeometricQueries::insideDomain (geometric_queries.h:357-371)
// takes its parity. Hits are kept in a small sorted list; a ray with more than kMaxAllHits hits sets kStatusStackOverflow.
constexpr int kMaxAllHits = 48;

template <int DIM>
ZB_DEV int countRayHits(const BvhView& bvh, const Vec<DIM>& o, const Vec<DIM>& d, float tMax, int *status)
{
    uint2 stack[kStackSize];
    int sp = 0;
    stack[sp++] = make_uint2(0u, 0u);
    float ht[kMaxAllHits];
    Vec<DIM> hp[kMaxAllHits];
    int nHits = 0;
    float invD[DIM];
#pragma unroll
    for (int k = 0; k < DIM; k++) invD[k] = 1.0f/d[k];

    while (sp > 0) {
        uint2 e = stack[--sp];
        if (e.y & kLeafFlag) {
            uint32_t first = e.y & 0x0fffffffu, count = (e.y >> 28) & 7u;
            for (uint32_t i = 0; i < count; i++) {
                Prim<DIM> prim;
                prim.load(bvh, first + i);
                float t, bv, bw;
                if (!prim.intersect(o, d, tMax, t, bv, bw)) continue;
                if (nHits == kMaxAllHits) { if (status) zb_atomic_or(status, kStatusStackOverflow); continue; }
                Vec<DIM> p = prim.pointAt(bv, bw);
                int k = nHits++;
                while (k > 0 && ht[k - 1] > t) { ht[k] = ht[k - 1]; hp[k] = hp[k - 1]; k--; }
                ht[k] = t; hp[k] = p;
            }
            continue;
        }
        NodeRegs n;
        loadNode(bvh.nodes, e.y, n);
        NodeFrame<DIM> f = nodeFrame<DIM>(n);
        ZB_CHILD_LOOP
        for (int c = 0; c < 8; c++) {
            uint32_t meta = childMeta(n, c);
            if (meta == 0u) continue;
            float lo[DIM], hi[DIM];
            childBox<DIM>(n, f, c, lo, hi);
            float tNear = 0.0f, tFar = tMax;
            bool outside = false;
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                if (d[k] == 0.0f) {
                    // an axis-parallel ray (insideDomain casts nothing else): the slab holds for every t or for none. The
                    // product form would give 0*inf = NaN for a ray IN a face plane of the box — through a mesh vertex
                    // that is a corner of its leaf's box — and must not cull the box then (FCPW's NaN comparisons keep it)
                    outside = outside || o[k] < lo[k] || o[k] > hi[k];
                    continue;
                }
                float t0 = (lo[k] - o[k])*invD[k], t1 = (hi[k] - o[k])*invD[k];
                tNear = fmaxf(tNear, fminf(t0, t1));
                tFar = fminf(tFar, fmaxf(t0, t1));
            }
            if (outside || !(tNear <= tFar)) continue;
            if (sp < kStackSize) stack[sp++] = make_uint2(0u, childRef(n, meta));
            else if (status) zb_atomic_or(status, kStatusStackOverflow);
        }
    }
    int distinct = 0;
    for (int k = 0; k < nHits; k++) {
        bool dup = false;
        if (k > 0) {
            Vec<DIM> dp = hp[k] - hp[k - 1];
            dup = fabsf(ht[k] - ht[k - 1]) < 1e-6f && norm2(dp) < 1e-6f;
        }
        if (!dup) distinct++;
    }
    return distinct;
}

// ------------------------------------------------------------------------------------------
// stochastic sampling of one primitive inside a sphere
// ------------------------------------------------------------------------------------------

// getBranchTraversalWeightCallback (fcpw_geometric_queries.h:1391-1398): |G_harmonic3D(max(r, 1e-2))|
ZB_DEV float branchWeight(float r2)
{
    float r = fmaxf(sqrtf(r2), 1e-2f);
    return fabsf(1.0f/(4.0f*kPiF*r));
}

template <int DIM>
struct SphereSample {
    int prim;    // leaf order, -1 when no primitive intersects the sphere
    Vec<DIM> p, n;
    float pdf;
};

// Aggregate::intersect(sphere, randNums, branchTraversalWeight) — bvh.inl:1141-1288: one stochastic
// root-to-leaf descent; here every level chooses among up to 8 children in proportion to the
// traversal weight of the child box centroid, then a primitive of the chosen leaf in proportion
// to its area by reservoir selection, all driven by the single random number rand[0].
template <int DIM>
ZB_DEV bool sampleSphere(const BvhView& bvh, const Vec<DIM>& x, float r2, const Vec<DIM>& rand, SphereSample<DIM>& out)
{
    out.prim = -1;
    out.pdf = 0.0f;
    out.p = x;
    out.n = vzero<DIM>();
    float u = rand[0];
    float traversalPdf = 1.0f;
    uint32_t nodeIdx = 0;

    while (true) {
        NodeRegs n;
        loadNode(bvh.nodes, nodeIdx, n);
        NodeFrame<DIM> f = nodeFrame<DIM>(n);
        // the child loop is not unrolled (ZB_CHILD_LOOP); the weights rotate through eight registers, so child c ends
        // up in weight[7 - c] and no array is indexed with a run-time value
        float weight[8];
#pragma unroll
        for (int c = 0; c < 8; c++) weight[c] = 0.0f;
        float total = 0.0f;
        ZB_CHILD_LOOP
        for (int c = 0; c < 8; c++) {
            float w = 0.0f;
            if (childMeta(n, c) != 0u) {
                float lo[DIM], hi[DIM], d2min, d2max;
                childBox<DIM>(n, f, c, lo, hi);
                boxDistance2<DIM>(lo, hi, x, d2min, d2max);
                if (d2min <= r2) {
                    float c2 = 0.0f;
#pragma unroll
                    for (int k = 0; k < DIM; k++) { float dc = x[k] - 0.5f*(lo[k] + hi[k]); c2 = fmaf(dc, dc, c2); }
                    w = branchWeight(c2);
                }
            }
#pragma unroll
            for (int i = 7; i > 0; i--) weight[i] = weight[i - 1];
            weight[0] = w;
            total += w;
        }
        if (!(total > 0.0f)) return false;

        // select a child: u is rescaled to [0,1) after each decision, as in bvh.inl:1237-1247
        int chosen = -1;
        float prob = 0.0f;
        float remaining = total;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            float wc = weight[7 - c];
            if (chosen < 0 && wc > 0.0f) {
                float p = wc/remaining;
                if (u < p || wc >= remaining) {
                    chosen = c;
                    u = fminf(u/p, 0.99999994f);
                    prob = wc/total;
                } else {
                    u = (u - p)/(1.0f - p);
                    remaining -= wc;
                }
            }
        }
        if (chosen < 0) return false;
        traversalPdf *= prob;
        float d2NodeMax;
        {
            float lo[DIM], hi[DIM], d2min;
            childBox<DIM>(n, f, chosen, lo, hi);
            boxDistance2<DIM>(lo, hi, x, d2min, d2NodeMax);
        }
        uint32_t meta = childMeta(n, chosen);
        uint32_t count = meta >> 5;
        if (count == 0u) {
            nodeIdx = n.q[1].x + (meta & 31u) - 1u;
            continue;
        }

        // leaf: reservoir-select a primitive by area among those touching the sphere (bvh.inl:1163-1216)
        uint32_t first = n.q[1].y + (meta & 31u);
        float totalArea = 0.0f, selectedArea = 0.0f;
        int selected = -1;
        for (uint32_t i = 0; i < count; i++) {
            Prim<DIM> prim;
            prim.load(bvh, first + i);
            bool inside = d2NodeMax <= r2;
            if (!inside) {
                float d = primClosestCall<DIM>(bvh.prims, first + i, x).w;
                inside = d*d <= r2;
            }
            if (inside) {
                float area = prim.area();
                totalArea += area;
                float selectionProb = area/totalArea;
                if (u < selectionProb) {
                    u = u/selectionProb;
                    selected = (int)(first + i);
                    selectedArea = area;
                } else {
                    u = (u - selectionProb)/(1.0f - selectionProb);
                }
            }
        }
        if (selected < 0) return false;
        Prim<DIM> prim;
        prim.load(bvh, (uint32_t)selected);
        float pointPdf = prim.sample(rand, out.p, out.n);
        out.prim = selected;
        out.pdf = selectedArea*traversalPdf/totalArea*pointPdf;
        return true;
    }
}

// ------------------------------------------------------------------------------------------
// star radius (visibility silhouettes + Robin reflectance bound)
// ------------------------------------------------------------------------------------------

// silhouette_edges.inl:83-111 / silhouette_vertices.inl:59-85: is the edge (3D) or vertex (2D) shared by
// faces with unit normals n0, n1 a visibility silhouette as seen along viewDir (|viewDir| = d)?
ZB_DEV bool isSilhouetteEdge3(const Vec<3>& pa, const Vec<3>& pb, const Vec<3>& n0, const Vec<3>& n1,
                              const Vec<3>& viewDir, float d, bool flip, float precision, float convention)
{
    float sign = flip ? 1.0f : -1.0f;
    if (d <= precision) {
        Vec<3> e = pb - pa;
        e = e/norm(e);
        float angle = atan2f(dot(e, cross(n0, n1)), dot(n0, n1));
        return sign*angle > precision;
    }
    Vec<3> vd = viewDir/d;
    float dot0 = convention*dot(vd, n0), dot1 = convention*dot(vd, n1);
    if (fabsf(dot0) <= precision) return sign*dot1 > precision;
    if (fabsf(dot1) <= precision) return sign*dot0 > precision;
    return dot0*dot1 < 0.0f;
}

ZB_DEV bool isSilhouetteVertex2(const Vec<2>& n0, const Vec<2>& n1, const Vec<2>& viewDir, float d,
                                bool flip, float precision, float convention)
{
    float sign = flip ? 1.0f : -1.0f;
    if (d <= precision) return sign*cross2(n0, n1) > precision;
    Vec<2> vd = viewDir/d;
    float dot0 = convention*dot(vd, n0), dot1 = convention*dot(vd, n1);
    if (fabsf(dot0) <= precision) return sign*dot1 > precision;
    if (fabsf(dot1) <= precision) return sign*dot0 > precision;
    return dot0*dot1 < 0.0f;
}

// distance from x to silhouette candidate j of the primitive (edge j -> j+1 of a triangle, vertex j
// of a segment) and whether it is a silhouette given the adjacent normal nj
ZB_DEV bool candidateSilhouette(const Prim<3>& t, int j, const Vec<3>& x, const Vec<3>& nj, const Vec<3>& n0,
                                bool hasAdjacent, bool flip, float precision, float convention, float& sd)
{
    Vec<3> pj = j == 0 ? t.a : (j == 1 ? t.b : t.c);
    Vec<3> pk = j == 0 ? t.b : (j == 1 ? t.c : t.a);
    Vec<3> pt;
    float tt;
    sd = closestPointSegment<3>(pj, pk, x, pt, tt);
    if (!hasAdjacent) return true;
    return isSilhouetteEdge3(pj, pk, nj, n0, x - pt, sd, flip, precision, convention);
}

ZB_DEV bool candidateSilhouette(const Prim<2>& s, int j, const Vec<2>& x, const Vec<2>& nj, const Vec<2>& n0,
                                bool hasAdjacent, bool flip, float precision, float convention, float& sd)
{
    Vec<2> vd = x - (j == 0 ? s.a : s.b);
    sd = norm(vd);
    if (!hasAdjacent) return true;
    return j == 0 ? isSilhouetteVertex2(n0, nj, vd, sd, flip, precision, convention)
                  : isSilhouetteVertex2(nj, n0, vd, sd, flip, precision, convention);
}

// farthest vertex (zombie/core/geometry_helpers.h:51-63, 163-180)
ZB_DEV float farthestVertex(const Prim<3>& t, const Vec<3>& x, Vec<3>& far)
{
    Vec<3> va = x - t.a, vb = x - t.b, vc = x - t.c;
    float da = xdot(va, va), db = xdot(vb, vb), dc = xdot(vc, vc);
    if (da >= db && da >= dc) { far = t.a; return sqrtf(da); }
    if (db >= da && db >= dc) { far = t.b; return sqrtf(db); }
    far = t.c;
    return sqrtf(dc);
}

ZB_DEV float farthestVertex(const Prim<2>& s, const Vec<2>& x, Vec<2>& far)
{
    Vec<2> va = x - s.a, vb = x - s.b;
    float da = xdot(va, va), db = xdot(vb, vb);
    if (da > db) { far = s.a; return sqrtf(da); }
    far = s.b;
    return sqrtf(db);
}

ZB_DEV float clampLikeStd(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }

// radius inside which the reflectance of a primitive with Robin coefficient mu stays bounded,
// given the distance h to its supporting plane/line and the cosine range it subtends
// (RobinTriangleBound / RobinLineSegmentBound, robin_bounds.h:13-74). Returns kMaxFloat for "no bound".
template <int DIM>
ZB_DEV float robinRadiusBound(float h, float mu, float cosLower, float cosUpper)
{
    if (DIM == 3) {
        // written without FMA contraction: h mu - c^2 cancels when the query is nearly coplanar with the triangle
        float hmu = zb_fmul(h, mu);
        float maxCos = sqrtf(hmu);
        if (!(maxCos >= cosLower)) return kMaxFloat;
        float c = clampLikeStd(maxCos/1.7320508075688772935f, cosLower, cosUpper);
        return zb_fmul(zb_fmul(h, h), mu)/zb_fmul(c, zb_fsub(hmu, zb_fmul(c, c)));
    }
    float hmu = zb_fmul(h, mu);
    float c = clampLikeStd(sqrtf(hmu)/1.4142135623730950488f, cosLower, cosUpper);
    return zb_fmul(h/c, expf(zb_fmul(c, c)/hmu));
}

// Shrinks r2 by what one reflecting primitive allows: the distance to each of its silhouette
// edges/vertices, and for a non-zero Robin coefficient the radius inside which the reflectance
// stays bounded (reflectance_boundary_bvh/geometry.h:85-236).
template <int DIM>
ZB_DEV void primStarRadius2(const BvhView& bvh, uint32_t primIdx, const Vec<DIM>& x, float& r2, bool flip,
                            float precision, bool silhouetteTests)
{
    Prim<DIM> prim;
    prim.load(bvh, primIdx);
    Vec<DIM> closestPt;
    float bu, bv;
    float closestDist = prim.closest(x, closestPt, bu, bv);
    float closestDist2 = closestDist*closestDist;
    if (closestDist2 > r2) { ZB_CLASS(2) return; }
    ZB_CLASS(3)

    // PrimAdjacency: n[3][3], minCoeff, maxCoeff, flags
    float4 a0 = zb_ldg(bvh.adjacency + (size_t)3*primIdx), a1 = zb_ldg(bvh.adjacency + (size_t)3*primIdx + 1),
           a2 = zb_ldg(bvh.adjacency + (size_t)3*primIdx + 2);
    float maxCoeff = a2.z;
    uint32_t flags = zb_float_as_uint(a2.w);
    if (!(maxCoeff < kMaxFloat - kEpsilon)) { // perfectly absorbing
        r2 = fminf(r2, closestDist2);
        return;
    }

    Vec<DIM> n0 = unitNormal<DIM>(prim);
    if (silhouetteTests) {
        const float adjN[3][3] = {{a0.x, a0.y, a0.z}, {a0.w, a1.x, a1.y}, {a1.z, a1.w, a2.x}};
        float convention = (float)bvh.snchSign;
#pragma unroll
        for (int j = 0; j < DIM; j++) {
            if ((flags >> (4 + j)) & 1u) continue; // ignoreAdjacentFace
            Vec<DIM> nj;
#pragma unroll
            for (int k = 0; k < DIM; k++) nj[k] = adjN[j][k];
            float sd;
            if (candidateSilhouette(prim, j, x, nj, n0, (flags >> j) & 1u, flip, precision, convention, sd)) {
                r2 = fminf(r2, sd*sd);
            }
        }
    }

    if (maxCoeff > kEpsilon) { // reflectance case
        if (closestDist < kEpsilon) {
            r2 = closestDist; // sic (geometry.h:216)
        } else if (r2 > closestDist2) {
            Vec<DIM> far;
            float farDist = farthestVertex(prim, x, far);
            // reference-ordered, uncontracted: h and the two cosines cancel when x is nearly coplanar with the primitive
            float h = fabsf(xdot(n0, prim.a - x)/xdot(n0, n0)); // findClosestPointPlane (geometry.h:62-70)
            if (h > 0.0f) {
                float cosUpper = fabsf(xdot((x - closestPt)/closestDist, n0));
                float cosLower = fabsf(xdot((x - far)/farDist, n0));
                float radius = robinRadiusBound<DIM>(h, maxCoeff, cosLower, cosUpper);
                ZB_PRIM_TRACE(primIdx, "exact reflectance: h cosLower cosUpper radius^2", h, cosLower, cosUpper, radius*radius)
                if (radius < kMaxFloat) r2 = fminf(r2, radius*radius);
            } else {
                r2 = fminf(r2, closestDist2);
            }
        }
    }
}

// Out-of-line copy of the reference-ordered test for primStarRadius2Fast's rare redo (arguments by value: a noinline
// device function taking references would pin them in local memory at every call site).
template <int DIM>
ZB_DEV_COLD float primStarRadius2Redo(const float4 *prims, const float4 *adjacency, int snchSign, uint32_t primIdx, Vec<DIM> x,
                                      float r2, bool flip, float precision, bool silhouetteTests)
{
    BvhView bvh{};
    bvh.prims = prims;
    bvh.adjacency = adjacency;
    bvh.snchSign = snchSign;
    primStarRadius2<DIM>(bvh, primIdx, x, r2, flip, precision, silhouetteTests);
    return r2;
}

// The reflectance part of primStarRadius2 alone, in the reference's order (closest point, plane distance, clamping cosines,
// radius bound: geometry.h:196-236): what primStarRadius2Fast redoes when only its reflectance bound is ill-conditioned —
// a grazing view or a bound close to its pole, by far the most frequent reasons for a redo. r2: the radius after the
// primitive's silhouette part (or any smaller radius found since: radii combine by minimum). ~450 instructions against
// ~1,900 for the whole primitive.
template <int DIM>
ZB_DEV float primStarRobinRedo(const float4 *prims, uint32_t primIdx, Vec<DIM> x, float r2)
{
    return r2; // the 2D star test has no fast form and never asks for it
}

template <>
ZB_DEV_COLD float primStarRobinRedo<3>(const float4 *prims, uint32_t primIdx, Vec<3> x, float r2)
{
    constexpr int DIM = 3;
    BvhView bvh{};
    bvh.prims = prims;
    Prim<DIM> prim;
    prim.load(bvh, primIdx);
    Vec<DIM> closestPt;
    float bu, bv;
    float closestDist = prim.closest(x, closestPt, bu, bv);
    float closestDist2 = closestDist*closestDist;
    if (closestDist2 > r2) return r2;
    const float maxCoeff = prim.maxCoeff;
    if (!(maxCoeff > kEpsilon) || !(maxCoeff < kMaxFloat - kEpsilon)) return r2;
    if (closestDist < kEpsilon) return closestDist; // sic (geometry.h:216)
    if (!(r2 > closestDist2)) return r2;
    Vec<DIM> n0 = unitNormal<DIM>(prim);
    Vec<DIM> far;
    float farDist = farthestVertex(prim, x, far);
    float h = fabsf(xdot(n0, prim.a - x)/xdot(n0, n0)); // findClosestPointPlane (geometry.h:62-70)
    if (!(h > 0.0f)) return fminf(r2, closestDist2);
    float cosUpper = fabsf(xdot((x - closestPt)/closestDist, n0));
    float cosLower = fabsf(xdot((x - far)/farDist, n0));
    float radius = robinRadiusBound<DIM>(h, maxCoeff, cosLower, cosUpper);
    return radius < kMaxFloat ? fminf(r2, radius*radius) : r2;
}

// the on-the-edge branch of isSilhouetteEdge3 (atan2 of the dihedral angle): rare, kept out of line
ZB_DEV_COLD bool isSilhouetteEdge3Cold(Vec<3> pa, Vec<3> pb, Vec<3> n0, Vec<3> n1, Vec<3> viewDir, float d, bool flip,
                                       float precision, float convention)
{
    return isSilhouetteEdge3(pa, pb, n0, n1, viewDir, d, flip, precision, convention);
}

#ifndef ZB_NO_EDGE_FLAG
#define ZB_NO_EDGE_FLAG 1 // 0: ignore the builder's "no testable edge below" mark (A/B builds)
#endif
#ifndef ZB_SILHOUETTE_SIGN_SHORTCUT
#define ZB_SILHOUETTE_SIGN_SHORTCUT 1 // 0: run the full edge test on every edge (A/B builds)
#endif
#ifndef ZB_ROBIN_PLANE_BOUND
#define ZB_ROBIN_PLANE_BOUND 1 // 0: evaluate every reflectance bound in full (A/B builds)
#endif

// The same primitive test with fused arithmetic, reciprocal multiplies and a single normalisation (3D only; the 2D
// form forwards to primStarRadius2). Identical case analysis, so the star radius agrees with primStarRadius2 to float
// rounding (the gate on this query is FCPW's 1e-3 silhouette tolerance); roughly half the instructions.
template <int DIM>
ZB_DEV void primStarRadius2Fast(const BvhView& bvh, uint32_t primIdx, const Vec<DIM>& x, float& r2, bool flip,
                                float precision, bool silhouetteTests, int *deferRedo = nullptr)
{
    primStarRadius2<DIM>(bvh, primIdx, x, r2, flip, precision, silhouetteTests);
}

#ifndef ZB_ROBIN_REDO_GRAZING
#define ZB_ROBIN_REDO_GRAZING 0.02f // ... and when the closest point of the triangle is seen at a cosine below this
#endif
#ifndef ZB_ROBIN_REDO_CONDITIONING
#define ZB_ROBIN_REDO_CONDITIONING 0.1f // the fast reflectance bound is redone in the reference's order when h mu - c^2 < this * h mu
#endif
#ifndef ZB_ROBIN_RANGE_BOUND
#define ZB_ROBIN_RANGE_BOUND 1 // 0: only the plane-distance bound (A/B builds)
#endif

// Second, sharper early exit of the reflectance part of the star-radius test. On a triangle whose plane lies at distance h
// from x, the point at distance rho has cosine h/rho, so the reference's radius h^2 mu/(c (h mu - c^2)) (robin_bounds.h:48-74)
// is g(rho) = mu rho^3/(mu rho^2 - h) for some rho between the closest and the farthest point of the triangle — its clamping
// cosines are exactly h/closestDist and h/farDist (geometry.h:219-226). g is infinite for mu rho^2 <= h and has its only
// minimum at rho^2 = 3h/mu, so over the superset [h, farDist] of that range its minimum is g at clamp(3h/mu, h^2, farDist^2):
// a lower bound of the radius this triangle can report, from the plane distance and the three vertex distances alone. It
// rejects 97 % of the triangles the plane-distance bound lets through (tools/host_probe.cu: of 4.6 M full evaluations on a
// C3-like solve only 73 K lowered the radius). s = nRaw . (a - x), nn = |nRaw|^2. 0.2 % of slack on the squared radius; not
// applied where the fused plane distance is itself ill-conditioned (grazing views, h below 0.3 % of the vertex distance).
ZB_DEV bool robinRangeCannotShrink(const Prim<3>& prim, const Vec<3>& x, float s, float nn, float mu, float r2)
{
    if (!ZB_ROBIN_RANGE_BOUND || !(nn > 0.0f)) return false;
    Vec<3> va = x - prim.a, vb = x - prim.b, vc = x - prim.c;
    float dFar2 = fmaxf(dot(va, va), fmaxf(dot(vb, vb), dot(vc, vc)));
    float h2 = s*s/nn;
    if (!(h2 > 1e-12f && h2 > 1e-5f*dFar2)) return false;
    float h = sqrtf(h2);
    float q = fminf(fmaxf(3.0f*h/mu, h2), dFar2);
    float den = mu*q - h;
    if (den <= 0.0f) return true; // mu rho^2 <= h over the whole triangle: no reflectance bound (robin_bounds.h:57-58)
    return q*q*q*mu*mu >= r2*den*den*1.002f;
}

ZB_DEV float closestPointTriangleFast2(const Vec<3>& pa, const Vec<3>& pb, const Vec<3>& pc, const Vec<3>& x, Vec<3>& pt)
{
    Vec<3> ab = pb - pa, ac = pc - pa, ax = x - pa;
    float d1 = dot(ab, ax), d2 = dot(ac, ax);
    Vec<3> bx = x - pb, cx = x - pc;
    float d3 = dot(ab, bx), d4 = dot(ac, bx), d5 = dot(ab, cx), d6 = dot(ac, cx);
    float vc = d1*d4 - d3*d2, vb = d5*d2 - d1*d6, va = d3*d6 - d5*d4;
    float v, w; // pt = pa + ab*v + ac*w
    if (d1 <= 0.0f && d2 <= 0.0f) { v = 0.0f; w = 0.0f; }
    else if (d3 >= 0.0f && d4 <= d3) { v = 1.0f; w = 0.0f; }
    else if (d6 >= 0.0f && d5 <= d6) { v = 0.0f; w = 1.0f; }
    else if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) { v = d1/(d1 - d3); w = 0.0f; }
    else if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) { v = 0.0f; w = d2/(d2 - d6); }
    else if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) { w = (d4 - d3)/((d4 - d3) + (d5 - d6)); v = 1.0f - w; }
    else { float denom = 1.0f/(va + vb + vc); v = vb*denom; w = vc*denom; }
    pt = pa + ab*v + ac*w;
    return norm2(x - pt);
}

// deferRedo: when given, a primitive that needs a reference-ordered redo is NOT redone here: *deferRedo says which redo
// (1: the whole primitive, r2 untouched; 2: the reflectance part only, r2 = the radius after the silhouette part), so that the
// caller can collect such primitives and redo them together (coop.cuh: in a warp a redo costs the same whether one lane
// takes it or all 32 do).
template <>
ZB_DEV void primStarRadius2Fast<3>(const BvhView& bvh, uint32_t primIdx, const Vec<3>& x, float& r2, bool flip,
                                   float precision, bool silhouetteTests, int *deferRedo)
{
    Prim<3> prim;
    prim.load(bvh, primIdx);
    float maxCoeff = prim.maxCoeff;
    uint32_t flags = prim.adjFlags;
    // Can the reflectance bound of this triangle shrink the radius at all? With h the distance from x to the triangle's
    // plane, the bound h^2 mu/(c (h mu - c^2)) (robin_bounds.h:48-74) is smallest at c = sqrt(h mu/3), whatever the
    // clamping cosines are, where its square is (27/4) h/mu. So 6.7 h >= mu r^2 (0.7 % of slack for rounding) means the
    // reflectance part of the test cannot lower r2, and neither the closest nor the farthest point is needed.
    bool robinCannotShrink = false;
    if (ZB_ROBIN_PLANE_BOUND && maxCoeff < kMaxFloat - kEpsilon) {
        if (!(maxCoeff > kEpsilon)) {
            robinCannotShrink = true; // perfectly reflecting: the reference skips the reflectance part (geometry.h:215)
        } else {
            Vec<3> nRaw = cross(prim.b - prim.a, prim.c - prim.a);
            float s = dot(nRaw, prim.a - x), nn = dot(nRaw, nRaw);
            float t = fmaxf(r2*maxCoeff*(1.0f/6.7f), 1e-6f); // h > 1e-6 also keeps clear of the closestDist < epsilon case
            robinCannotShrink = nn > 0.0f && s*s >= nn*(t*t);
            if (!robinCannotShrink) robinCannotShrink = robinRangeCannotShrink(prim, x, s, nn, maxCoeff, r2);
        }
        // ... and when, on top of that, every edge is marked ignoreAdjacentFace (convex edges of a single-sided solve,
        // fcpw_geometric_queries.h:1029-1038) or the subtree holds no silhouette, the triangle has nothing to say
        if (robinCannotShrink && !(silhouetteTests && ((flags >> 4) & 7u) != 7u)) { ZB_CLASS(4) return; }
        if (robinCannotShrink) { ZB_CLASS(5) } else if (silhouetteTests) { ZB_CLASS(6) } else { ZB_CLASS(7) }
    }
    Vec<3> closestPt = x;
    float closestDist2 = 0.0f;
    if (!robinCannotShrink) {
        closestDist2 = closestPointTriangleFast2(prim.a, prim.b, prim.c, x, closestPt);
        if (closestDist2 > r2) { ZB_CLASS(2) return; }
    }
    ZB_CLASS(3)
    // Discrete decisions (is this edge a silhouette? does the reflectance bound apply?) are taken with fused
    // arithmetic; whenever one of them falls within rounding distance of its threshold, or the bound formula is
    // ill-conditioned, the primitive is redone with the reference-ordered arithmetic of primStarRadius2, so the
    // two never disagree by more than float rounding.
    const float r2in = r2;
    bool redo = false;      // a silhouette decision within rounding distance of its threshold: redo the whole primitive
    bool redoRobin = false; // the reflectance bound alone is ill-conditioned: redo that part (primStarRobinRedo)

    float4 a0 = zb_ldg(bvh.adjacency + (size_t)3*primIdx), a1 = zb_ldg(bvh.adjacency + (size_t)3*primIdx + 1),
           a2 = zb_ldg(bvh.adjacency + (size_t)3*primIdx + 2);
    if (!(maxCoeff < kMaxFloat - kEpsilon)) { // perfectly absorbing
        r2 = fminf(r2, closestDist2);
        return;
    }

    Vec<3> n0 = unitNormal<3>(prim); // reference-ordered: the plane distance below cancels and must round the same way
    if (silhouetteTests) {
        float convention = (float)bvh.snchSign;
        float sign = flip ? 1.0f : -1.0f;
        // ONE copy of the edge test: the loop is not unrolled; vertices and neighbour normals rotate through their
        // registers instead (edge j runs from the current pj to pk), so nothing is indexed with a run-time value
        Vec<3> pj = prim.a, pk = prim.b, pl = prim.c;
        Vec<3> nj = make3(a0.x, a0.y, a0.z), nk = make3(a0.w, a1.x, a1.y), nl = make3(a1.z, a1.w, a2.x);
        // Sign shortcut. The closest point of an edge lies in the planes of both faces that share it, so the two dot
        // products of the silhouette test are the signed distances s0, sj of x from those planes over the distance d
        // to the edge. If both exceed precision times an upper bound of d (the distance to a vertex of the edge) and
        // have the same sign, the edge is no silhouette (silhouette_edges.inl:97-110) whatever d is — and one of them
        // above precision in absolute terms rules out the on-the-edge branch (d <= precision) since d >= |s|. 10 % of
        // slack covers the rounding of s against the reference's normalised dot products.
        const float s0 = dot(n0, x - prim.a);
        const float p2 = 1.21f*precision*precision;
#pragma unroll 1
        for (int j = 0; j < 3; j++) {
            bool skip = ((flags >> (4 + j)) & 1u) != 0u; // ignoreAdjacentFace
            if (ZB_SILHOUETTE_SIGN_SHORTCUT && !skip && ((flags >> j) & 1u)) {
                Vec<3> v = x - pj;
                float sj = dot(nj, v), lim = p2*dot(v, v);
                skip = s0*s0 > lim && sj*sj > lim && (s0 > 0.0f) == (sj > 0.0f) && fmaxf(s0*s0, sj*sj) > p2;
                if (skip) { ZB_CLASS(8) }
            }
            if (!skip) {
                ZB_CLASS(9)
                Vec<3> e = pk - pj, v = x - pj;
                float c1 = dot(e, v), c2 = dot(e, e);
                float t = c1 <= 0.0f ? 0.0f : (c2 <= c1 ? 1.0f : c1/c2);
                Vec<3> vd = v - e*t;
                float sd2 = dot(vd, vd);
                if (sd2 < r2) { // else: cannot shrink the radius whatever the classification says
                    bool isSilhouette = true;
                    if ((flags >> j) & 1u) {
                        float d = sqrtf(sd2);
                        if (d <= precision) {
                            isSilhouette = isSilhouetteEdge3Cold(pj, pk, nj, n0, vd, d, flip, precision, convention);
                        } else {
                            float s = convention/d;
                            float dot0 = s*dot(vd, nj), dot1 = s*dot(vd, n0);
                            redo = redo || fabsf(fabsf(dot0) - precision) < 4e-6f || fabsf(fabsf(dot1) - precision) < 4e-6f;
                            if (fabsf(dot0) <= precision) isSilhouette = sign*dot1 > precision;
                            else if (fabsf(dot1) <= precision) isSilhouette = sign*dot0 > precision;
                            else isSilhouette = dot0*dot1 < 0.0f;
                        }
                    }
                    if (isSilhouette) r2 = sd2;
                }
            }
            Vec<3> tp = pj; pj = pk; pk = pl; pl = tp;
            Vec<3> tn = nj; nj = nk; nk = nl; nl = tn;
        }
    }

    const float r2AfterSilhouettes = r2;
    if (maxCoeff > kEpsilon && !robinCannotShrink) { // reflectance case
        float closestDist = sqrtf(closestDist2);
        if (closestDist < kEpsilon) {
            r2 = closestDist; // sic (geometry.h:216)
        } else if (r2 > closestDist2) {
            Vec<3> va = x - prim.a, vb = x - prim.b, vc = x - prim.c;
            float da = dot(va, va), db = dot(vb, vb), dc = dot(vc, vc);
            Vec<3> vfar = (da >= db && da >= dc) ? va : ((db >= da && db >= dc) ? vb : vc);
            float far2 = fmaxf(da, fmaxf(db, dc));
            float h = fabsf(xdot(n0, prim.a - x)/xdot(n0, n0)); // findClosestPointPlane (geometry.h:62-70), reference-ordered
            if (h > 0.0f) {
                float cosUpper = fabsf(xdot((x - closestPt)/closestDist, n0));
                float farDist = sqrtf(far2);
                float cosLower = fabsf(xdot(vfar/farDist, n0));
                float radius = robinRadiusBound<3>(h, maxCoeff, cosLower, cosUpper);
                if (radius < kMaxFloat) r2 = fminf(r2, radius*radius);
                // threshold maxCos >= cosLower, the conditioning h mu/(h mu - c^2) of the bound (robin_bounds.h:48-74), and the
                // cancellation in h and the cosines when x is nearly coplanar with the triangle (the radius is then h/c with both
                // at rounding level: only the reference's own operation order reproduces the reference's value)
                float hmu = h*maxCoeff, maxCos = sqrtf(hmu);
                float c = clampLikeStd(maxCos/1.7320508075688772935f, cosLower, cosUpper);
                if (cosUpper < ZB_ROBIN_REDO_GRAZING) { ZB_CLASS(17) } else if (fabsf(maxCos - cosLower) < 1e-5f) { ZB_CLASS(18) } else if (maxCos >= cosLower && (hmu - c*c) < ZB_ROBIN_REDO_CONDITIONING*hmu) { ZB_CLASS(19) }
                redoRobin = cosUpper < ZB_ROBIN_REDO_GRAZING || fabsf(maxCos - cosLower) < 1e-5f || (maxCos >= cosLower && (hmu - c*c) < ZB_ROBIN_REDO_CONDITIONING*hmu);
                ZB_PRIM_TRACE(primIdx, "fast reflectance: h cosLower cosUpper radius^2", h, cosLower, cosUpper, radius*radius)
            } else {
                r2 = fminf(r2, closestDist2);
            }
        }
    }
    ZB_PRIM_TRACE(primIdx, "fast result before redo: r2in r2 redo cannotShrink", r2in, r2, redo ? 1.0f : 0.0f, robinCannotShrink ? 1.0f : 0.0f)
    if (redo || redoRobin) { ZB_CLASS(15) }
    if (deferRedo) {
        // 1: redo the whole primitive from r2in; 2: the silhouette part stands, redo the reflectance part only
        if (redo) { *deferRedo = 1; r2 = r2in; return; }
        if (redoRobin) { *deferRedo = 2; r2 = r2AfterSilhouettes; return; }
    } else if (redo) {
        r2 = primStarRadius2Redo<3>(bvh.prims, bvh.adjacency, bvh.snchSign, primIdx, x, r2in, flip, precision, silhouetteTests);
    } else if (redoRobin) {
        r2 = primStarRobinRedo<3>(bvh.prims, primIdx, x, r2AfterSilhouettes);
    }
    ZB_PRIM_TRACE(primIdx, "final: r2 | query", r2, x[0], x[1], x[2])
    if (r2 < r2in) { ZB_CLASS(12) if (!silhouetteTests) { ZB_CLASS(13) } }
}

// Cheap first stage of primStarRadius2Fast for the pooled primitive rounds of coop.cuh: false when the triangle cannot
// lower r2 (its reflectance bound is cut off by the plane distance and it has no edge to test), so that only the
// survivors of a round are compacted into rounds of the full test — in a warp the full test costs the same whether
// one lane needs it or all 32 do.
template <int DIM>
ZB_DEV bool primStarNeedsFullTest(const BvhView& bvh, uint32_t primIdx, const Vec<DIM>& x, float r2, bool silhouetteTests)
{
    return true;
}

template <>
ZB_DEV bool primStarNeedsFullTest<3>(const BvhView& bvh, uint32_t primIdx, const Vec<3>& x, float r2, bool silhouetteTests)
{
    Prim<3> prim;
    prim.load(bvh, primIdx);
    float maxCoeff = prim.maxCoeff;
    uint32_t flags = prim.adjFlags;
    if (!ZB_ROBIN_PLANE_BOUND || !(maxCoeff < kMaxFloat - kEpsilon)) return true;
    if (silhouetteTests && ((flags >> 4) & 7u) != 7u) return true;
    if (!(maxCoeff > kEpsilon)) return false;
    Vec<3> nRaw = cross(prim.b - prim.a, prim.c - prim.a);
    float s = dot(nRaw, prim.a - x), nn = dot(nRaw, nRaw);
    float t = fmaxf(r2*maxCoeff*(1.0f/6.7f), 1e-6f);
    if (nn > 0.0f && s*s >= nn*(t*t)) return false;
    return !robinRangeCannotShrink(prim, x, s, nn, maxCoeff, r2);
}

// One out-of-line copy of the fast primitive test for all the traversals of a kernel (the private depth-first phase and
// the pooled primitive rounds of coop.cuh both call it): ~1300 instructions that would otherwise be resident twice.
// Build with -DZB_INLINE_PRIM_TESTS to inline it at every call site instead.
template <int DIM>
#if defined(ZB_INLINE_PRIM_TESTS)
ZB_DEV
#else
ZB_DEV_COLD
#endif
float2 primStarRadius2Call(const float4 *prims, const float4 *adjacency, int snchSign, uint32_t primIdx, Vec<DIM> x, float r2,
                           bool flip, float precision, bool silhouetteTests, bool defer = false)
{
    BvhView bvh{};
    bvh.prims = prims;
    bvh.adjacency = adjacency;
    bvh.snchSign = snchSign;
    int deferred = 0;
    primStarRadius2Fast<DIM>(bvh, primIdx, x, r2, flip, precision, silhouetteTests, defer ? &deferred : nullptr);
    // x: the radius; y (defer only): 0 done, 1 redo the whole primitive (primStarRadius2Redo), 2 redo its reflectance part
    // (primStarRobinRedo) — as a float so that the pair travels in registers
    return make_float2(r2, (float)deferred);
}

// BoundingCone::overlap (fcpw/core/bounding_volumes.h:275-321): can a normal inside the cone be
// orthogonal to a direction from x to the box? Also returns the range of angles between the two.
template <int DIM>
ZB_DEV bool coneOverlap(const float axis[3], float halfAngle, float radius, const Vec<DIM>& x,
                        const float lo[DIM], const float hi[DIM], float distToBox, float& minAngle, float& maxAngle)
{
    minAngle = 0.0f;
    maxAngle = kHalfPiF;
    if (halfAngle >= kHalfPiF || distToBox < kEpsilon) return true;

    Vec<DIM> c, view, ext;
#pragma unroll
    for (int k = 0; k < DIM; k++) { c[k] = 0.5f*(lo[k] + hi[k]); view[k] = c[k] - x[k]; ext[k] = hi[k] - c[k]; }
    float l = norm(view);
    view = view/l;
    float ad = 0.0f;
#pragma unroll
    for (int k = 0; k < DIM; k++) ad = fmaf(axis[k], view[k], ad);
    float dAxisAngle = acosf(fmaxf(-1.0f, fminf(1.0f, ad)));
    if (kHalfPiF >= dAxisAngle - halfAngle && kHalfPiF <= dAxisAngle + halfAngle) return true;

    float viewHalfAngle;
    if (l > radius) {
        viewHalfAngle = asinf(radius/l);
    } else {
        float d = 0.0f;
#pragma unroll
        for (int k = 0; k < DIM; k++) d = fmaf(ext[k], fabsf(view[k]), d);
        float s = l - d;
        if (s <= 0.0f) return true;
        float proj;
        if (DIM == 2) {
            proj = fabsf(ext[0]*fabsf(view[1]) + ext[1]*fabsf(view[0]));
        } else {
            // maximal projection of the box half-extent onto the plane orthogonal to `view`
            float sgn = copysignf(1.0f, view[DIM - 1]);
            float a = -1.0f/(sgn + view[DIM - 1]);
            float b = view[0]*view[1]*a;
            float b1x = 1.0f + sgn*view[0]*view[0]*a, b1y = sgn*b, b1z = -sgn*view[0];
            float b2x = b, b2y = sgn + view[1]*view[1]*a, b2z = -view[1];
            float r1 = ext[0]*fabsf(b1x) + ext[1]*fabsf(b1y) + ext[DIM - 1]*fabsf(b1z);
            float r2 = ext[0]*fabsf(b2x) + ext[1]*fabsf(b2y) + ext[DIM - 1]*fabsf(b2z);
            proj = sqrtf(r1*r1 + r2*r2);
        }
        viewHalfAngle = atan2f(proj, s);
    }
    float sum = halfAngle + viewHalfAngle + 1e-5f;
    minAngle = dAxisAngle - sum;
    maxAngle = dAxisAngle + sum;
    return sum >= kHalfPiF ? true : (kHalfPiF >= minAngle && kHalfPiF <= maxAngle);
}

// ReflectanceBvh::visitNode (reflectance_boundary_bvh/bvh.h:370-412) for child slot c of a wide node
// whose box [lo, hi] lies within d2min..d2max of x: bounds [lower, upper] on the squared star radius
// anything inside the child can produce, and whether a silhouette may hide inside. Returns false when
// the child cannot shrink the radius at all.
template <int DIM>
ZB_DEV bool starChildBounds(const BvhView& bvh, uint32_t nodeIdx, int c, const Vec<DIM>& x, const float lo[DIM], const float hi[DIM],
                            float d2min, float d2max, float& lower, float& upper, bool& hasSil, float precision)
{
    float4 c0 = zb_ldg(bvh.cones + ((size_t)nodeIdx*8 + c)*2), c1 = zb_ldg(bvh.cones + ((size_t)nodeIdx*8 + c)*2 + 1);
    float minCoeff = c1.y, maxCoeff = c1.z;
    lower = d2min;
    upper = d2max;
    hasSil = true;
    if (minCoeff < kMaxFloat - kEpsilon) {
        float axis[3] = {c0.x, c0.y, c0.z};
        float minAngle, maxAngle;
        // A query within `precision` of an edge / vertex classifies it by its dihedral angle alone, whatever the view
        // direction (silhouette_edges.inl:88-95, silhouette_vertices.inl:64-68), which the view-cone test cannot see:
        // such children always count as possible silhouettes, so the traversal equals a brute-force scan there too.
        if (d2min <= precision*precision || coneOverlap<DIM>(axis, c0.w, c1.x, x, lo, hi, d2min, minAngle, maxAngle)) {
            upper = kMaxFloat;
        } else {
            hasSil = false;
            if (!(maxCoeff > kEpsilon)) return false; // perfectly reflecting and no silhouette inside
            // RobinBvhNodeBound (robin_bounds.h:102-154): radius range over the whole subtree
            float rMin = sqrtf(d2min), rMax = sqrtf(d2max);
            float minCos = fminf(fabsf(cosf(minAngle)), fabsf(cosf(maxAngle)));
            if (DIM == 3) {
                if (rMax < minCos/maxCoeff) lower = kMaxFloat;
                else { float rb = rMin/(1.0f - minCos/(maxCoeff*rMax)); lower = rb*rb; }
                if (rMin < 1.0f/minCoeff) upper = kMaxFloat;
                else { float rb = rMax/(1.0f - 1.0f/(minCoeff*rMin)); upper = rb*rb; }
            } else {
                float rb = rMin*expf(minCos/(maxCoeff*rMax));
                lower = rb*rb;
                rb = rMax*expf(1.0f/(minCoeff*rMin));
                upper = rb*rb;
            }
            // keep the pruning bounds conservative under float rounding
            lower *= 0.9999f;
            upper *= 1.0001f;
        }
    }
    return true;
}

// ReflectanceBvh::visitNode (reflectance_boundary_bvh/bvh.h:370-412) for one child of a wide node, in the
// transcendental-free form: the reference's acos/asin/atan2 cone test (fcpw/core/bounding_volumes.h:275-321)
// becomes products of sines and cosines with explicit slack (node tests only prune, so any conservative form
// returns the same query results). Same contract as starChildBounds (queries.cuh):
// bounds [lower, upper] on the squared star radius anything inside the child can produce, whether a silhouette
// may hide inside; false when the child cannot shrink the radius at all. Conservative with respect to
// starChildBounds' own tests: the overlap threshold carries 2e-5 of slack on the sine, the Robin cosine bound
// 1e-5, on top of the 1e-4 rad by which the builder pads every cone.
#ifndef ZB_NODE_JOINT_BOUND
#define ZB_NODE_JOINT_BOUND 1 // 0: the reference's interval form r_min/(1 - cos/(mu r_max)) (A/B builds)
#endif

template <int DIM>
ZB_DEV bool starChildBoundsFast(const float4& c0, const float4& c1, const Vec<DIM>& x,
                                                    const float lo[DIM], const float hi[DIM], float d2min, float d2max,
                                                    float& lower, float& upper, bool& hasSil, float precision)
{
    // c0, c1: the child's ChildCone (axis | half angle, radius | coefficient range), loaded by the caller one child
    // ahead (ConePrefetch below) — the child loop is rolled, and a load issued where its value is needed would put
    // one full memory latency into every iteration
    float minCoeff = c1.y, maxCoeff = c1.z;
    lower = d2min;
    upper = d2max;
    hasSil = true;
    if (!(minCoeff < kMaxFloat - kEpsilon)) return true; // perfectly absorbing subtree: plain distance bounds

    upper = kMaxFloat;
    // The builder marks subtrees in which every edge carries the reference's ignore mark (convex edges of a single-sided
    // solve, fcpw_geometric_queries.h:1029-1038): their primitive tests skip every edge, so wherever the cone test below
    // would have to answer "may hold a silhouette" such a subtree holds none — a perfectly reflecting one (Neumann) is
    // pruned, a Robin one is kept for its reflectance bounds only (no radius bound is derived for it: lower = d2min).
    const bool noEdges = ZB_NO_EDGE_FLAG && c1.w != 0.0f;
#define ZB_MAY_HOLD_SILHOUETTE { if (!noEdges) return true; hasSil = false; return maxCoeff > kEpsilon; }
    float halfAngle = c0.w;
    if (halfAngle >= kHalfPiF || d2min < kEpsilon || d2min <= precision*precision) ZB_MAY_HOLD_SILHOUETTE // see starChildBounds
    float view[DIM], l2 = 0.0f;
#pragma unroll
    for (int k = 0; k < DIM; k++) { view[k] = 0.5f*(lo[k] + hi[k]) - x[k]; l2 = fmaf(view[k], view[k], l2); }
    float radius = c1.x;
    if (!(l2 > radius*radius)) ZB_MAY_HOLD_SILHOUETTE // inside the bounding sphere of the child: treat as overlapping
    float invL = zb_rsqrt_fast(l2);
    float cosTheta = 0.0f;
    const float axis[3] = {c0.x, c0.y, c0.z};
#pragma unroll
    for (int k = 0; k < DIM; k++) cosTheta = fmaf(axis[k], view[k], cosTheta);
    cosTheta = fminf(1.0f, fabsf(cosTheta*invL));
    float sinA, cosA;
    zb_sincos_fast(halfAngle, &sinA, &cosA);
    float sinB = fminf(1.0f, radius*invL), cosB = zb_sqrt_fast(fmaxf(0.0f, 1.0f - sinB*sinB));
    float sinG = sinA*cosB + cosA*sinB, cosG = cosA*cosB - sinA*sinB; // gamma = cone half angle + view half angle
    if (cosG <= 1e-4f || cosTheta <= sinG + 2e-5f) ZB_MAY_HOLD_SILHOUETTE // pi/2 within theta +- gamma
#undef ZB_MAY_HOLD_SILHOUETTE

    hasSil = false;
    if (!(maxCoeff > kEpsilon)) return false; // perfectly reflecting and no silhouette inside
    // RobinBvhNodeBound (robin_bounds.h:102-154): radius range over the whole subtree
    ZB_CLASS(11)
    float sinTheta = zb_sqrt_fast(fmaxf(0.0f, 1.0f - cosTheta*cosTheta));
    float minCos = fmaxf(0.0f, cosTheta*cosG - sinTheta*sinG - 1e-5f); // min |cos| over [theta - gamma, theta + gamma]
    float rMin = zb_sqrt_fast(d2min), rMax = zb_sqrt_fast(d2max);
    if (DIM == 3) {
        // r_min/(1 - cos/(mu r_max)) and r_max/(1 - 1/(mu r_min)) with approximate reciprocals: the denominators are
        // padded by 1e-6 (>> the 2-ulp error of their operands, also where they cancel) on the conservative side
#if ZB_NODE_JOINT_BOUND
        // every point y of the subtree has R(y) = rho/(1 - cos/(mu rho)) >= rho^2/(rho - k), k = minCos/maxCoeff, rho = |y - x|
        // in [rMin, rMax]; that function of rho has its minimum 4k at rho = 2k and is infinite for rho <= k
        float k = minCos*zb_rcp_fast(maxCoeff)*0.99999f;
        if (!(rMax > k*1.00001f)) lower = kMaxFloat;
        else { float rs = fminf(fmaxf(2.0f*k, rMin), rMax); float rb = rs*rs*zb_rcp_fast(rs - k + 1e-6f); lower = fmaxf(rb*rb, d2min); }
#else
        float a = minCos*zb_rcp_fast(maxCoeff*rMax);
        if (a > 1.00001f) lower = kMaxFloat; // r_max < cos/mu
        else { float rb = rMin*zb_rcp_fast(fmaxf(1.0f - a, 0.0f) + 1e-6f); lower = rb*rb; }
#endif
        if (!(rMin*minCoeff < 1.0f)) {
            float den = 1.0f - zb_rcp_fast(minCoeff*rMin) - 1e-6f;
            if (den > 0.0f) { float rb = rMax*zb_rcp_fast(den); upper = rb*rb; }
        }
    } else {
        float rb = rMin*zb_exp_fast(minCos/(maxCoeff*rMax));
        lower = rb*rb;
        rb = rMax*expf(1.0f/(minCoeff*rMin));
        upper = rb*rb;
    }
    // keep the pruning bounds conservative under float rounding
    lower *= 0.9999f;
    upper *= 1.0001f;
    return true;
}

// One-child-ahead loads of a node's eight ChildCones for the rolled child loops of the star-radius traversals.
struct ConePrefetch {
    const float4 *base;
    float4 n0, n1;
    ZB_DEV void start(const BvhView& bvh, uint32_t nodeIdx) {
        base = bvh.cones + (size_t)nodeIdx*16;
        n0 = zb_ldg(base);
        n1 = zb_ldg(base + 1);
    }
    // returns child c's cone and issues the loads of child c + 1 (every node owns eight entries, empty slots included)
    ZB_DEV void next(int c, float4& c0, float4& c1) {
        c0 = n0;
        c1 = n1;
        if (c < 7) {
            n0 = zb_ldg(base + 2*(c + 1));
            n1 = zb_ldg(base + 2*(c + 1) + 1);
        }
    }
};

// Depth-first core of the star-radius query: pops up to `budget` stack entries and returns with the unfinished
// part of the traversal on the stack (sp > 0) when the budget runs out. FAST selects the algebraic child test.
template <int DIM, bool FAST>
ZB_DEV void starRadiusDfs(const BvhView& bvh, const Vec<DIM>& x, float& r2, bool flip, float precision, uint2 *stack, int& sp,
                          int budget, int *status)
{
    ZB_COUNT_DECL
    while (sp > 0 && budget-- > 0) {
        uint2 e = stack[--sp];
        // bit 2 of the key marks "no silhouette inside this subtree" (reflectance bvh.h:474-476 sign trick)
        bool silhouetteTests = !(e.x & 4u);
        if (zb_uint_as_float(e.x & ~7u) > r2) continue;
        ZB_CLASS(silhouetteTests ? 0 : 1)
        if (e.y & kLeafFlag) {
            uint32_t first = e.y & 0x0fffffffu, count = (e.y >> 28) & 7u;
            for (uint32_t i = 0; i < count; i++) {
                ZB_COUNT_PRIM
                if (FAST) r2 = primStarRadius2Call<DIM>(bvh.prims, bvh.adjacency, bvh.snchSign, first + i, x, r2, flip, precision, silhouetteTests).x;
                else primStarRadius2<DIM>(bvh, first + i, x, r2, flip, precision, silhouetteTests);
            }
            continue;
        }

        NodeRegs n;
        loadNode(bvh.nodes, e.y, n);
        ZB_COUNT_NODE
        NodeFrame<DIM> f = nodeFrame<DIM>(n);
        uint32_t keys[8];
        uint32_t noSil = 0u;
#pragma unroll
        for (int c = 0; c < 8; c++) keys[c] = 0xffffffffu;
        ConePrefetch cones;
        if (FAST) cones.start(bvh, e.y);
        ZB_CHILD_LOOP
        for (int c = 0; c < 8; c++) {
            float4 c0, c1;
            if (FAST) cones.next(c, c0, c1);
            uint32_t meta = childMeta(n, c);
            if (meta == 0u) continue;
            float lo[DIM], hi[DIM], d2min, d2max;
            childBox<DIM>(n, f, c, lo, hi);
            boxDistance2<DIM>(lo, hi, x, d2min, d2max);
            if (d2min > r2) continue;
            ZB_CLASS(10)
            float lower, upper;
            bool hasSil;
            bool keep = FAST ? starChildBoundsFast<DIM>(c0, c1, x, lo, hi, d2min, d2max, lower, upper, hasSil, precision)
                             : starChildBounds<DIM>(bvh, e.y, c, x, lo, hi, d2min, d2max, lower, upper, hasSil, precision);
            if (!keep) continue;
            r2 = fminf(r2, upper);
            if (!hasSil || !silhouetteTests) noSil |= 1u << c;
            insertKey(keys, (zb_float_as_uint(lower) & ~7u) | (uint32_t)c);
        }
#pragma unroll
        for (int c = 7; c >= 0; c--) {
            uint32_t key = keys[c];
            if (key != 0xffffffffu && keyValue(key) <= r2) {
                uint32_t slot = key & 7u;
                uint32_t tagged = (key & ~7u) | (((noSil >> slot) & 1u) ? 4u : 0u);
                if (sp < kStackSize) stack[sp++] = make_uint2(tagged, childRef(n, childMeta(n, slot)));
                else if (status) zb_atomic_or(status, kStatusStackOverflow);
            }
        }
    }
    ZB_COUNT_FLUSH(4)
}

// GeometricQueries::computeStarRadiusForReflectingBoundary for both the Neumann and the Robin
// handler (fcpw_geometric_queries.h:1250-1291): returns the squared radius, <= r2Max. One thread per query,
// with the reference's own trigonometric cone test (the cross-check of the fast form).
template <int DIM>
ZB_DEV float starRadius2(const BvhView& bvh, const Vec<DIM>& x, float r2, bool flip, float precision, int *status)
{
    uint2 stack[kStackSize];
    int sp = 0;
    stack[sp++] = make_uint2(0u, 0u);
#if defined(ZB_SEQUENTIAL_FAST) // host tools: measure the kernels' (fast) node and primitive tests one thread per query
    starRadiusDfs<DIM, true>(bvh, x, r2, flip, precision, stack, sp, 0x7fffffff, status);
#else
    starRadiusDfs<DIM, false>(bvh, x, r2, flip, precision, stack, sp, 0x7fffffff, status);
#endif
    return r2;
}

} // namespace zb
