// On-device construction of the wide quantised hierarchy of wide_bvh.h.
//
// Replaces, for scenes built or updated on the GPU, the reference's serial host build and refit:
// fcpw Bvh::build / buildRecursive (fcpw/aggregates/bvh.inl:143-234), Bvh::refit (bvh.inl:549-600),
// ReflectanceBvh node bounds + refit (zombie/utils/reflectance_boundary_bvh/bvh.h:77-186, 271-312) and the
// adjacency tables of FcpwRobinBoundaryHandler / FcpwNeumannBoundaryHandler::buildAccelerationStructure
// (zombie/utils/fcpw_geometric_queries.h:728-795, 892-973). It produces exactly the layout the host builder
// (wide_bvh.cpp) produces — the query kernels do not know which builder ran — with a different, level-synchronous
// algorithm: primitives are sorted along a Morton curve of their centroids and the 8-ary tree is cut DIRECTLY out of
// the sorted order (no binary tree, no collapse):
//
//   1. centroid bounds (block reduction + ordered-integer atomics), 63-bit Morton keys, cub radix sort;
//   2. topology, one pass per level from the root down: a node owns a range of the sorted order and cuts it into up
//      to eight child ranges by repeatedly splitting its largest range (> 4 primitives) at the highest differing key
//      bit (binary search); ranges of <= 4 primitives are leaf children. Internal children and the primitives of
//      leaf children get their places from exclusive scans over the level, so the node array is breadth first with
//      contiguous internal children and every node's leaf primitives are contiguous — and the result is deterministic;
//   3. adjacency (reflecting boundaries): half edges sorted by their vertex pair, the opposite face of a half edge is
//      the last (largest) face in the other orientation slot of its edge — the reference's last-writer-wins tables;
//   4. primitives scattered into leaf order (de-indexed vertex records, adjacency records);
//   5. boxes bottom-up, one pass per level: exact child boxes, the node's quantisation frame, conservative 8-bit boxes;
//   6. normal cones, bounding radii, Robin-coefficient ranges and the no-testable-edge mark per child slot over the slot's
//      range of the sorted order: one 512-thread CTA per slot on the top levels, one warp per slot below.
//
// Per-primitive records are computed with the host builder's un-contracted arithmetic and are bit-identical to its
// records; the tree differs (spatial-median cuts instead of binned SAH), which changes traversal cost, not results:
// quantised boxes and cones are conservative, so every query returns what a brute-force scan returns.
//
// The per-item logic is host+device so that tools/host_build.cu can run the builder on a CPU against the host
// builder and the reference (tests/test_device_builder.py); the library only ever runs it from the kernels below.
#pragma once

#include <float.h>
#include <stdint.h>

#include "vec.cuh"
#include "wide_bvh.h"

namespace zb {
namespace lbvh {

constexpr int kLeaf = kMaxLeafSize;
#if !defined(ZB_LBVH_MEDIAN)
#define ZB_LBVH_MEDIAN 0
#endif

struct Mesh {
    int dim;
    int nPrim;
    int nVert;
    const float *pos;   // nVert*dim
    const int32_t *idx; // nPrim*dim
};

// one node of the topology pass: its range of the sorted order and the cuts between its children
struct TopoNode {
    int start, end;
    int nChildren;
    int bounds[kWidth + 1];
    int childBase; // index of its first internal child
    int primBase;  // leaf-order position of the first primitive of its leaf children
    int pad;
};

struct Box3 {
    float lo[3], hi[3];
};

ZB_HOSTDEV void loadVertex(const Mesh& m, int prim, int k, float p[3])
{
    const float *q = m.pos + (size_t)m.dim*m.idx[(size_t)m.dim*prim + k];
    p[0] = q[0]; p[1] = q[1]; p[2] = m.dim == 3 ? q[2] : 0.0f;
}

// un-contracted float arithmetic, the host builder's (wide_bvh.cpp: cross3, dot3, norm3 compiled without FMA)
ZB_HOSTDEV void ucross(const float *a, const float *b, float *c)
{
    c[0] = zb_fsub(zb_fmul(a[1], b[2]), zb_fmul(a[2], b[1]));
    c[1] = zb_fsub(zb_fmul(a[2], b[0]), zb_fmul(a[0], b[2]));
    c[2] = zb_fsub(zb_fmul(a[0], b[1]), zb_fmul(a[1], b[0]));
}
ZB_HOSTDEV float udot(const float *a, const float *b)
{
    return zb_fadd(zb_fadd(zb_fmul(a[0], b[0]), zb_fmul(a[1], b[1])), zb_fmul(a[2], b[2]));
}
ZB_HOSTDEV float unorm(const float *a) { return sqrtf(udot(a, a)); }

// unit normal: triangles (b-a)x(c-a), segments (s.y, -s.x) (wide_bvh.cpp: Geometry::normal)
ZB_HOSTDEV void primNormal(const Mesh& m, int prim, float n[3])
{
    float a[3], b[3], c[3];
    loadVertex(m, prim, 0, a); loadVertex(m, prim, 1, b);
    if (m.dim == 3) {
        loadVertex(m, prim, 2, c);
        float u[3] = {zb_fsub(b[0], a[0]), zb_fsub(b[1], a[1]), zb_fsub(b[2], a[2])};
        float v[3] = {zb_fsub(c[0], a[0]), zb_fsub(c[1], a[1]), zb_fsub(c[2], a[2])};
        ucross(u, v, n);
    } else {
        n[0] = zb_fsub(b[1], a[1]); n[1] = -zb_fsub(b[0], a[0]); n[2] = 0.0f;
    }
    float l = unorm(n);
    if (l > 0.0f) { n[0] /= l; n[1] /= l; n[2] /= l; }
}

ZB_HOSTDEV void primCentroid(const Mesh& m, int prim, float c[3])
{
    c[0] = c[1] = c[2] = 0.0f;
    for (int k = 0; k < m.dim; k++) {
        float p[3];
        loadVertex(m, prim, k, p);
        c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
    }
    for (int k = 0; k < 3; k++) c[k] = c[k]/(float)m.dim;
}

// ---- 1. Morton keys -------------------------------------------------------------------------------------------
ZB_HOSTDEV uint64_t spreadBy3(uint64_t v) // 21 bits -> every third bit
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
ZB_HOSTDEV uint64_t spreadBy2(uint64_t v) // 31 bits -> every second bit
{
    v &= 0x7fffffffull;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}
// lo / inv: lower corner of the centroid bounds and 1/extent per axis (0 for a degenerate axis)
ZB_HOSTDEV uint64_t mortonKey(int dim, const float c[3], const float lo[3], const float inv[3])
{
    if (dim == 3) {
        uint64_t q[3];
        for (int k = 0; k < 3; k++) {
            float t = (c[k] - lo[k])*inv[k];
            t = fminf(fmaxf(t, 0.0f), 1.0f);
            uint32_t u = (uint32_t)(t*2097152.0f);
            q[k] = u > 2097151u ? 2097151u : u;
        }
        return (spreadBy3(q[0]) << 2) | (spreadBy3(q[1]) << 1) | spreadBy3(q[2]);
    }
    uint64_t q[2];
    for (int k = 0; k < 2; k++) {
        double t = ((double)c[k] - (double)lo[k])*(double)inv[k];
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        uint64_t u = (uint64_t)(t*2147483648.0);
        q[k] = u > 2147483647ull ? 2147483647ull : u;
    }
    return (spreadBy2(q[0]) << 1) | spreadBy2(q[1]);
}

// ---- 2. topology ----------------------------------------------------------------------------------------------
ZB_HOSTDEV int leadingZeros64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v == 0 ? 64 : __builtin_clzll(v);
#endif
}

// cut of the sorted range [first, last), last - first >= 2: where the highest differing key bit flips; the middle when
// all keys agree. Both sides are non-empty.
ZB_HOSTDEV int findSplit(const uint64_t *keys, int first, int last)
{
    uint64_t a = keys[first], b = keys[last - 1];
    if (a == b || ZB_LBVH_MEDIAN) return first + (last - first)/2;
    int shift = 63 - leadingZeros64(a ^ b); // the differing bit
    int lo = first, hi = last - 1;          // keys[lo] agrees with a above `shift`, keys[hi] does not
    while (hi - lo > 1) {
        int mid = lo + (hi - lo)/2;
        if (((keys[mid] ^ a) >> shift) == 0) lo = mid; else hi = mid;
    }
    return hi;
}

// cuts a node's range into up to eight children; counts its internal children and the primitives of its leaf children
ZB_HOSTDEV void splitNode(const uint64_t *keys, TopoNode& n, int& nInternal, int& nLeafPrims)
{
    int b[kWidth + 1];
    int nc = 1;
    b[0] = n.start; b[1] = n.end;
    if (n.end - n.start > kLeaf) {
        while (nc < kWidth) {
            int best = -1, bestCount = kLeaf;
            for (int c = 0; c < nc; c++) {
                int cnt = b[c + 1] - b[c];
                if (cnt > bestCount) { bestCount = cnt; best = c; }
            }
            if (best < 0) break;
            int s = findSplit(keys, b[best], b[best + 1]);
            for (int c = nc; c > best; c--) b[c + 1] = b[c];
            b[best + 1] = s;
            nc++;
        }
    }
    n.nChildren = nc;
    nInternal = 0; nLeafPrims = 0;
    for (int c = 0; c <= kWidth; c++) n.bounds[c] = c <= nc ? b[c] : b[nc];
    for (int c = 0; c < nc; c++) {
        int cnt = b[c + 1] - b[c];
        if (cnt > kLeaf) nInternal++; else nLeafPrims += cnt;
    }
}

// places the children of node w: ranges of the internal children, leaf-order positions of the leaf primitives
ZB_HOSTDEV void placeChildren(TopoNode *topo, int w, int childBase, int primBase, int *leafPos)
{
    TopoNode& n = topo[w];
    n.childBase = childBase;
    n.primBase = primBase;
    int rank = 0, offset = 0;
    for (int c = 0; c < n.nChildren; c++) {
        int s = n.bounds[c], e = n.bounds[c + 1];
        if (e - s > kLeaf) {
            TopoNode& ch = topo[childBase + rank];
            ch.start = s; ch.end = e;
            rank++;
        } else {
            for (int i = s; i < e; i++) leafPos[i] = primBase + offset + (i - s);
            offset += e - s;
        }
    }
}

// ---- 3. adjacency ---------------------------------------------------------------------------------------------
// half edge j of triangle i: key = (min vertex, max vertex), value = (3i + j) << 1 | swapped
ZB_HOSTDEV void halfEdge(const Mesh& m, int he, uint64_t& key, uint32_t& val)
{
    int i = he/3, j = he - 3*i;
    int I = m.idx[3*(size_t)i + j], J = m.idx[3*(size_t)i + (j + 1)%3];
    uint32_t sw = 0;
    if (I > J) { int t = I; I = J; J = t; sw = 1; }
    key = ((uint64_t)(uint32_t)I << 32) | (uint32_t)J;
    val = ((uint32_t)he << 1) | sw;
}

// sorted position s of the half-edge array (stable sort by key: faces ascend within a key): the face in the opposite
// orientation slot of the undirected edge, last writer (= largest face) wins; -1 when the slot is empty
ZB_HOSTDEV void oppositeFace(const uint64_t *keys, const uint32_t *vals, int nHalf, int s, int *opposite)
{
    uint64_t key = keys[s];
    uint32_t mine = vals[s];
    uint32_t want = 1u - (mine & 1u);
    int last = s;
    while (last + 1 < nHalf && keys[last + 1] == key) last++;
    int found = -1;
    for (int e = last; e >= 0 && keys[e] == key; e--) {
        if ((vals[e] & 1u) == want) { found = (int)((vals[e] >> 1)/3u); break; }
    }
    opposite[mine >> 1] = found;
}

// PrimAdjacency of input primitive i (wide_bvh.cpp: computeAdjacency). 3D: opposite = 3 faces per triangle;
// 2D: startsAt / endsAt = the last segment starting / ending at each vertex.
ZB_HOSTDEV void primAdjacency(const Mesh& m, int i, const int *opposite, const int *startsAt, const int *endsAt,
                              const float *minCoeff, const float *maxCoeff, bool doubleSided, float precision, PrimAdjacency& out)
{
    for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) out.n[j][k] = 0.0f;
    out.flags = 0u;
    out.minCoeff = minCoeff ? minCoeff[i] : 0.0f;
    out.maxCoeff = maxCoeff ? maxCoeff[i] : 0.0f;
    float n0[3];
    primNormal(m, i, n0);
    if (m.dim == 2) {
        for (int j = 0; j < 2; j++) {
            int v = m.idx[2*(size_t)i + j];
            int other = j == 0 ? endsAt[v] : startsAt[v];
            if (other < 0) continue;
            float nj[3];
            primNormal(m, other, nj);
            out.flags |= 1u << j;
            for (int k = 0; k < 3; k++) out.n[j][k] = nj[k];
            float det = zb_fsub(zb_fmul(n0[0], nj[1]), zb_fmul(n0[1], nj[0]));
            float sign = j == 0 ? 1.0f : -1.0f;
            bool ignore = doubleSided ? false : zb_fmul(det, sign) < precision;
            if (ignore) out.flags |= 1u << (4 + j);
        }
        return;
    }
    for (int j = 0; j < 3; j++) {
        int other = opposite[3*(size_t)i + j];
        if (other < 0) continue;
        float nj[3];
        primNormal(m, other, nj);
        out.flags |= 1u << j;
        for (int k = 0; k < 3; k++) out.n[j][k] = nj[k];
        float pa[3], pb[3];
        loadVertex(m, i, j, pa); loadVertex(m, i, (j + 1)%3, pb);
        float ed[3] = {zb_fsub(pb[0], pa[0]), zb_fsub(pb[1], pa[1]), zb_fsub(pb[2], pa[2])};
        float l = unorm(ed);
        if (l > 0.0f) { ed[0] /= l; ed[1] /= l; ed[2] /= l; }
        float c[3];
        ucross(nj, n0, c);
        // the host builder's std::atan2 on floats; evaluated in double and rounded so that both agree to the last bit
        float dihedral = (float)atan2((double)udot(ed, c), (double)udot(n0, nj));
        bool ignore = doubleSided ? false : dihedral < precision;
        if (ignore) out.flags |= 1u << (4 + j);
    }
}

// ---- 4. leaf-order records ------------------------------------------------------------------------------------
// sorted position s -> leaf position leafPos[s]: the de-indexed vertex record and (reflecting) the adjacency record
ZB_HOSTDEV void scatterPrim(const Mesh& m, int p, int dst, const PrimAdjacency *adjInput, float *prims, int32_t *primIndex,
                            PrimAdjacency *adjacency)
{
    primIndex[dst] = p;
    if (m.dim == 3) {
        float *pf = prims + (size_t)12*dst;
        for (int v = 0; v < 3; v++) {
            float q[3];
            loadVertex(m, p, v, q);
            pf[4*v + 0] = q[0]; pf[4*v + 1] = q[1]; pf[4*v + 2] = q[2]; pf[4*v + 3] = 0.0f;
        }
        pf[3] = zb_int_as_float(p); // v0.w carries the caller's primitive index
        if (adjInput) {
            pf[7] = adjInput[p].maxCoeff;               // filter-stage data of the star query rides in v1.w / v2.w
            pf[11] = zb_uint_as_float(adjInput[p].flags);
        }
    } else {
        float *pf = prims + (size_t)4*dst;
        float q[3];
        loadVertex(m, p, 0, q); pf[0] = q[0]; pf[1] = q[1];
        loadVertex(m, p, 1, q); pf[2] = q[0]; pf[3] = q[1];
    }
    if (adjInput) adjacency[dst] = adjInput[p];
}

// ---- 5. boxes and quantisation --------------------------------------------------------------------------------
ZB_HOSTDEV float scaleFromExp(uint32_t e) { return zb_uint_as_float(e << 23); }

// smallest biased exponent e with 255 * 2^(e-127) >= extent (wide_bvh.cpp: chooseExp)
ZB_HOSTDEV uint32_t chooseExp(float extent)
{
    if (!(extent > 0.0f)) return 1;
    int ex;
    frexp((double)extent/255.0, &ex);
    int e = ex + 127;
    return (uint32_t)(e < 1 ? 1 : (e > 254 ? 254 : e));
}

ZB_HOSTDEV void growBox(Box3& b, const float p[3])
{
    for (int k = 0; k < 3; k++) { b.lo[k] = fminf(b.lo[k], p[k]); b.hi[k] = fmaxf(b.hi[k], p[k]); }
}

// writes wide node w from its topology, the leaf-order vertex records and the exact boxes of its internal children
// (which the pass over the level below wrote); stores its own exact box
ZB_HOSTDEV void writeNode(int dim, const TopoNode *topo, int w, const float *prims, Box3 *nodeBox, WideNode *nodes)
{
    const TopoNode& n = topo[w];
    Box3 cb[kWidth];
    Box3 nb;
    for (int k = 0; k < 3; k++) { nb.lo[k] = FLT_MAX; nb.hi[k] = -FLT_MAX; }
    int rank = 0, offset = 0;
    uint8_t meta[kWidth];
    for (int c = 0; c < kWidth; c++) {
        meta[c] = 0;
        if (c >= n.nChildren) continue;
        int cnt = n.bounds[c + 1] - n.bounds[c];
        Box3 bx;
        for (int k = 0; k < 3; k++) { bx.lo[k] = FLT_MAX; bx.hi[k] = -FLT_MAX; }
        if (cnt > kLeaf) {
            bx = nodeBox[n.childBase + rank];
            meta[c] = (uint8_t)(rank + 1);
            rank++;
        } else {
            for (int i = 0; i < cnt; i++) {
                int slot = n.primBase + offset + i;
                if (dim == 3) {
                    const float *pf = prims + (size_t)12*slot;
                    for (int v = 0; v < 3; v++) growBox(bx, pf + 4*v);
                } else {
                    const float *pf = prims + (size_t)4*slot;
                    float p0[3] = {pf[0], pf[1], 0.0f}, p1[3] = {pf[2], pf[3], 0.0f};
                    growBox(bx, p0); growBox(bx, p1);
                }
            }
            meta[c] = (uint8_t)((cnt << 5) | offset);
            offset += cnt;
        }
        cb[c] = bx;
        for (int k = 0; k < 3; k++) { nb.lo[k] = fminf(nb.lo[k], bx.lo[k]); nb.hi[k] = fmaxf(nb.hi[k], bx.hi[k]); }
    }
    nodeBox[w] = nb;

    WideNode wn;
    for (int k = 0; k < 20; k++) wn.w[k] = 0u;
    uint32_t exps[3];
    float scale[3];
    for (int k = 0; k < 3; k++) {
        wn.w[k] = zb_float_as_uint(nb.lo[k]);
        exps[k] = chooseExp(nb.hi[k] - nb.lo[k]);
        while (exps[k] < 254 && nb.lo[k] + scaleFromExp(exps[k])*255.0f < nb.hi[k]) exps[k]++; // far side representable
        scale[k] = scaleFromExp(exps[k]);
    }
    wn.w[3] = exps[0] | (exps[1] << 8) | (exps[2] << 16);
    wn.w[4] = (uint32_t)n.childBase;
    wn.w[5] = (uint32_t)n.primBase;
    for (int c = 0; c < kWidth; c++) {
        wn.w[6 + (c >> 2)] |= (uint32_t)meta[c] << (8*(c & 3));
        for (int k = 0; k < 3; k++) {
            uint32_t qlo = 255u, qhi = 0u; // empty slot: inverted box
            if (c < n.nChildren) {
                double rel = ((double)cb[c].lo[k] - (double)nb.lo[k])/(double)scale[k];
                int lo = (int)floor(rel);
                lo = lo < 0 ? 0 : (lo > 255 ? 255 : lo);
                while (lo > 0 && zb_fadd(nb.lo[k], zb_fmul(scale[k], (float)lo)) > cb[c].lo[k]) lo--;
                rel = ((double)cb[c].hi[k] - (double)nb.lo[k])/(double)scale[k];
                int hi = (int)ceil(rel);
                hi = hi < 0 ? 0 : (hi > 255 ? 255 : hi);
                while (hi < 255 && zb_fadd(nb.lo[k], zb_fmul(scale[k], (float)hi)) < cb[c].hi[k]) hi++;
                qlo = (uint32_t)lo; qhi = (uint32_t)hi;
            }
            wn.w[8 + 2*k + (c >> 2)] |= qlo << (8*(c & 3));
            wn.w[14 + 2*k + (c >> 2)] |= qhi << (8*(c & 3));
        }
    }
    nodes[w] = wn;
}

// decodes child slot c of a node exactly as the query kernels do (wide_bvh.cpp: decodeChildBox)
ZB_HOSTDEV void decodeChild(const WideNode& wn, int c, float lo[3], float hi[3])
{
    for (int k = 0; k < 3; k++) {
        float o = zb_uint_as_float(wn.w[k]);
        float s = scaleFromExp((wn.w[3] >> (8*k)) & 0xffu);
        uint32_t qlo = (wn.w[8 + 2*k + (c >> 2)] >> (8*(c & 3))) & 0xffu;
        uint32_t qhi = (wn.w[14 + 2*k + (c >> 2)] >> (8*(c & 3))) & 0xffu;
        lo[k] = zb_fadd(o, zb_fmul(s, (float)qlo));
        hi[k] = zb_fadd(o, zb_fmul(s, (float)qhi));
    }
}

// ---- 6. cones -------------------------------------------------------------------------------------------------
// first pass over the primitives of a child slot (reflectance_boundary_bvh/bvh.h:77-155; wide_bvh.cpp: computeCone)
struct ConeSums {
    float axis[3];
    float radius, minCoeff, maxCoeff;
    int allAdjacent, anyTestable;
};
ZB_HOSTDEV void coneInit(ConeSums& s)
{
    s.axis[0] = s.axis[1] = s.axis[2] = 0.0f;
    s.radius = 0.0f; s.minCoeff = FLT_MAX; s.maxCoeff = -FLT_MAX;
    s.allAdjacent = 1; s.anyTestable = 0;
}
ZB_HOSTDEV void coneAccumulate(const Mesh& m, const PrimAdjacency *adjInput, int p, const float centre[3], ConeSums& s)
{
    float n[3];
    primNormal(m, p, n);
    s.axis[0] += n[0]; s.axis[1] += n[1]; s.axis[2] += n[2];
    const PrimAdjacency& a = adjInput[p];
    s.minCoeff = fminf(s.minCoeff, a.minCoeff);
    s.maxCoeff = fmaxf(s.maxCoeff, a.maxCoeff);
    for (int k = 0; k < m.dim; k++) {
        float q[3];
        if (m.dim == 2) {
            loadVertex(m, p, k, q);
        } else {
            float va[3], vb[3];
            loadVertex(m, p, k, va); loadVertex(m, p, (k + 1)%3, vb);
            for (int c = 0; c < 3; c++) q[c] = zb_fmul(0.5f, zb_fadd(va[c], vb[c]));
        }
        float d[3] = {zb_fsub(q[0], centre[0]), zb_fsub(q[1], centre[1]), zb_fsub(q[2], centre[2])};
        s.radius = fmaxf(s.radius, unorm(d));
        s.allAdjacent = s.allAdjacent && ((a.flags >> k) & 1u);
        s.anyTestable = s.anyTestable || !((a.flags >> (4 + k)) & 1u);
    }
}
ZB_HOSTDEV void coneMerge(ConeSums& s, const ConeSums& o)
{
    for (int k = 0; k < 3; k++) s.axis[k] += o.axis[k];
    s.radius = fmaxf(s.radius, o.radius);
    s.minCoeff = fminf(s.minCoeff, o.minCoeff);
    s.maxCoeff = fmaxf(s.maxCoeff, o.maxCoeff);
    s.allAdjacent = s.allAdjacent && o.allAdjacent;
    s.anyTestable = s.anyTestable || o.anyTestable;
}
// second pass: largest angle between the axis and any face normal below, incl. the adjacent faces'
ZB_HOSTDEV float coneHalfAngle(const Mesh& m, const PrimAdjacency *adjInput, int p, const float axis[3])
{
    float n[3];
    primNormal(m, p, n);
    float half = acosf(fmaxf(-1.0f, fminf(1.0f, udot(axis, n))));
    const PrimAdjacency& a = adjInput[p];
    for (int k = 0; k < m.dim; k++) half = fmaxf(half, acosf(fmaxf(-1.0f, fminf(1.0f, udot(axis, a.n[k])))));
    return half;
}
ZB_HOSTDEV ChildCone emptyCone()
{
    ChildCone c;
    c.axis[0] = c.axis[1] = c.axis[2] = 0.0f;
    c.halfAngle = kPiF;
    c.radius = 0.0f; c.minCoeff = 0.0f; c.maxCoeff = 0.0f; c.noTestableEdge = 0.0f;
    return c;
}
// after the first pass: everything but the half angle; returns true when the second pass is needed (axis valid)
ZB_HOSTDEV bool coneFromSums(const ConeSums& s, ChildCone& cone, float axis[3])
{
    cone = emptyCone();
    cone.radius = s.radius;
    cone.minCoeff = s.minCoeff;
    cone.maxCoeff = s.maxCoeff;
    cone.noTestableEdge = s.anyTestable ? 0.0f : 1.0f;
    if (!s.allAdjacent) return false;
    float l = unorm(s.axis);
    if (!(l > FLT_EPSILON)) return false;
    for (int k = 0; k < 3; k++) axis[k] = s.axis[k]/l;
    return true;
}
ZB_HOSTDEV void coneFinish(ChildCone& cone, const float axis[3], float half)
{
    cone.axis[0] = axis[0]; cone.axis[1] = axis[1]; cone.axis[2] = axis[2];
    // pad: the device test must stay conservative under float rounding of acos / asin / atan2 (as the host builder)
    cone.halfAngle = fminf(kPiF, half + 1e-4f);
}

// ================================================================================================================
#if defined(__CUDACC__)

// order-preserving map float -> uint32 for atomicMin / atomicMax
__device__ __forceinline__ uint32_t orderedBits(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float fromOrderedBits(uint32_t u)
{
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    return zb_uint_as_float(v);
#endif
}

// bounds[0..2] = ordered min of the centroids, [3..5] = ordered max; flags: bit 0 index out of range, bit 1 non-finite vertex
__global__ void __launch_bounds__(256) centroidBoundsKernel(Mesh m, float *centroids, uint32_t *bounds, int *flags)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    float c[3] = {0.0f, 0.0f, 0.0f};
    bool valid = i < m.nPrim;
    if (valid) {
        int bad = 0;
        for (int k = 0; k < m.dim; k++) {
            int v = m.idx[(size_t)m.dim*i + k];
            if (v < 0 || v >= m.nVert) bad |= 1;
        }
        if (bad) { atomicOr(flags, bad); valid = false; }
    }
    if (valid) {
        primCentroid(m, i, c);
        if (!(isfinite(c[0]) && isfinite(c[1]) && isfinite(c[2]))) { atomicOr(flags, 2); valid = false; }
    }
    if (valid) { centroids[3*(size_t)i] = c[0]; centroids[3*(size_t)i + 1] = c[1]; centroids[3*(size_t)i + 2] = c[2]; }
    for (int k = 0; k < 3; k++) {
        uint32_t lo = valid ? orderedBits(c[k]) : 0xffffffffu, hi = valid ? orderedBits(c[k]) : 0u;
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[k], lo); atomicMax(&bounds[3 + k], hi); }
    }
}

__global__ void __launch_bounds__(256) mortonKernel(int dim, int nPrim, const float *centroids, const uint32_t *bounds,
                                                    uint64_t *keys, int *order)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= nPrim) return;
    float lo[3], inv[3];
    for (int k = 0; k < 3; k++) {
        lo[k] = fromOrderedBits(bounds[k]);
        float ext = fromOrderedBits(bounds[3 + k]) - lo[k];
        inv[k] = ext > 0.0f ? 1.0f/ext : 0.0f;
    }
    keys[i] = mortonKey(dim, centroids + 3*(size_t)i, lo, inv);
    order[i] = i;
}

__global__ void __launch_bounds__(128) splitLevelKernel(const uint64_t *keys, TopoNode *topo, int first, int count,
                                                        int *nInternal, int *nLeafPrims)
{
    int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= count) return;
    int ni, nl;
    splitNode(keys, topo[first + k], ni, nl);
    nInternal[k] = ni;
    nLeafPrims[k] = nl;
}

// counts[0] += internal children of the level, counts[1] += leaf primitives (read after the scans for the totals)
__global__ void __launch_bounds__(128) placeLevelKernel(TopoNode *topo, int first, int count, const int *internalOffset,
                                                        const int *leafOffset, int nextFirst, int primCursor, int *leafPos)
{
    int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= count) return;
    placeChildren(topo, first + k, nextFirst + internalOffset[k], primCursor + leafOffset[k], leafPos);
}

__global__ void __launch_bounds__(256) halfEdgeKernel(Mesh m, uint64_t *keys, uint32_t *vals)
{
    int he = blockIdx.x*blockDim.x + threadIdx.x;
    if (he >= 3*m.nPrim) return;
    halfEdge(m, he, keys[he], vals[he]);
}

__global__ void __launch_bounds__(256) oppositeFaceKernel(const uint64_t *keys, const uint32_t *vals, int nHalf, int *opposite)
{
    int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= nHalf) return;
    oppositeFace(keys, vals, nHalf, s, opposite);
}

__global__ void __launch_bounds__(256) vertexTablesKernel(Mesh m, int *startsAt, int *endsAt)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= m.nPrim) return;
    atomicMax(&startsAt[m.idx[2*(size_t)i]], i);    // last writer of the reference's loop = largest segment index
    atomicMax(&endsAt[m.idx[2*(size_t)i + 1]], i);
}

__global__ void __launch_bounds__(128) adjacencyKernel(Mesh m, const int *opposite, const int *startsAt, const int *endsAt,
                                                       const float *minCoeff, const float *maxCoeff, int doubleSided, float precision,
                                                       PrimAdjacency *adjInput)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= m.nPrim) return;
    PrimAdjacency a;
    primAdjacency(m, i, opposite, startsAt, endsAt, minCoeff, maxCoeff, doubleSided != 0, precision, a);
    adjInput[i] = a;
}

__global__ void __launch_bounds__(256) scatterPrimsKernel(Mesh m, const int *order, const int *leafPos, const PrimAdjacency *adjInput,
                                                          float *prims, int32_t *primIndex, PrimAdjacency *adjacency)
{
    int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= m.nPrim) return;
    scatterPrim(m, order[s], leafPos[s], adjInput, prims, primIndex, adjacency);
}

__global__ void __launch_bounds__(128) writeLevelKernel(int dim, const TopoNode *topo, int first, int count, const float *prims,
                                                        Box3 *nodeBox, WideNode *nodes)
{
    int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= count) return;
    writeNode(dim, topo, first + k, prims, nodeBox, nodes);
}

// butterfly over the lanes of a warp: every lane ends with the same sums (float additions commute pairwise)
__device__ __forceinline__ void coneWarpReduce(ConeSums& sums)
{
    for (int o = 16; o > 0; o >>= 1) {
        ConeSums other;
        for (int k = 0; k < 3; k++) other.axis[k] = __shfl_xor_sync(0xffffffffu, sums.axis[k], o);
        other.radius = __shfl_xor_sync(0xffffffffu, sums.radius, o);
        other.minCoeff = __shfl_xor_sync(0xffffffffu, sums.minCoeff, o);
        other.maxCoeff = __shfl_xor_sync(0xffffffffu, sums.maxCoeff, o);
        other.allAdjacent = __shfl_xor_sync(0xffffffffu, sums.allAdjacent, o);
        other.anyTestable = __shfl_xor_sync(0xffffffffu, sums.anyTestable, o);
        coneMerge(sums, other);
    }
}

// Child slots of the nodes [first, first + count). WARPS = 1: one warp per slot, four slots per CTA (the deep levels:
// many slots of few primitives). WARPS = 16: one 512-thread CTA per slot (the top levels: a handful of slots of up to
// nPrim/8 primitives each), partial sums combined through shared memory in a fixed order.
template <int WARPS>
__global__ void __launch_bounds__(WARPS == 1 ? 128 : 32*WARPS) coneLevelKernel(Mesh m, const TopoNode *topo, int first, int count,
                                                                            const WideNode *nodes, const int *order,
                                                                            const PrimAdjacency *adjInput, ChildCone *cones)
{
    __shared__ ConeSums partial[WARPS];
    __shared__ float partialHalf[WARPS];
    const int lane = threadIdx.x & 31, warpInCta = threadIdx.x >> 5;
    const long long slot = WARPS == 1 ? ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5 : (long long)blockIdx.x;
    if (slot >= (long long)count*kWidth) return;
    const int w = first + (int)(slot >> 3), c = (int)(slot & 7);
    const TopoNode& n = topo[w];
    if (c >= n.nChildren) {
        if (threadIdx.x == 0 || (WARPS == 1 && lane == 0)) cones[(size_t)w*kWidth + c] = emptyCone();
        return;
    }
    const int s = n.bounds[c], e = n.bounds[c + 1];
    const int tid = WARPS == 1 ? lane : (int)threadIdx.x, stride = 32*WARPS;
    float lo[3], hi[3], centre[3];
    decodeChild(nodes[w], c, lo, hi);
    for (int k = 0; k < 3; k++) centre[k] = 0.5f*(lo[k] + hi[k]);
    ConeSums sums;
    coneInit(sums);
    for (int i = s + tid; i < e; i += stride) coneAccumulate(m, adjInput, order[i], centre, sums);
    coneWarpReduce(sums);
    if (WARPS > 1) {
        if (lane == 0) partial[warpInCta] = sums;
        __syncthreads();
        sums = partial[0];
        for (int k = 1; k < WARPS; k++) coneMerge(sums, partial[k]);
    }
    ChildCone cone;
    float axis[3];
    if (coneFromSums(sums, cone, axis)) { // the same sums in every thread: a uniform branch
        float half = 0.0f;
        for (int i = s + tid; i < e; i += stride) half = fmaxf(half, coneHalfAngle(m, adjInput, order[i], axis));
        for (int o = 16; o > 0; o >>= 1) half = fmaxf(half, __shfl_xor_sync(0xffffffffu, half, o));
        if (WARPS > 1) {
            if (lane == 0) partialHalf[warpInCta] = half;
            __syncthreads();
            for (int k = 0; k < WARPS; k++) half = fmaxf(half, partialHalf[k]);
        }
        coneFinish(cone, axis, half);
    }
    if (tid == 0) cones[(size_t)w*kWidth + c] = cone;
}

#endif // __CUDACC__

} // namespace lbvh
} // namespace zb
