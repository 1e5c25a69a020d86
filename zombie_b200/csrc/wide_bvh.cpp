// Host-side construction of the wide quantised BVH (see wide_bvh.h). Cold path: runs once per
// boundary mesh. The reference builds a binary SAH tree serially (fcpw/aggregates/bvh.inl:143-234);
// here a binned-SAH binary tree is built with task parallelism over large subtrees and then
// collapsed into 8-ary nodes whose children are quantised against the parent box.
#include "wide_bvh.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <future>
#include <thread>

namespace zb {

namespace {

constexpr int kBins = 16;
constexpr float kPi = 3.14159265358979323846f;

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; } }
    void grow(const float *p) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    float halfArea(int dim) const {
        float e[3] = {std::max(0.0f, hi[0] - lo[0]), std::max(0.0f, hi[1] - lo[1]), std::max(0.0f, hi[2] - lo[2])};
        return dim == 3 ? e[0]*e[1] + e[1]*e[2] + e[2]*e[0] : e[0] + e[1];
    }
};

struct BinaryNode {
    Box box;
    int left = -1, right = -1; // children (internal) ...
    int start = 0, count = 0;  // ... range of `order` covered by this subtree; leaf iff left < 0
};

struct Builder {
    int dim;
    int nPrim;
    std::vector<Box> primBox;
    std::vector<float> centroid; // 3 per prim
    std::vector<int> order;
    std::vector<BinaryNode> nodes;
    std::atomic<int> nodeCount{0};
    std::atomic<int> liveTasks{0};
    int maxTasks = 1;

    int alloc() { return nodeCount.fetch_add(1); }

    void build(int nodeIdx, int start, int end, int depth) {
        BinaryNode& node = nodes[nodeIdx];
        node.start = start;
        node.count = end - start;
        Box box, cbox;
        box.reset(); cbox.reset();
        for (int i = start; i < end; i++) {
            box.grow(primBox[order[i]]);
            cbox.grow(&centroid[3*order[i]]);
        }
        node.box = box;
        int n = end - start;
        if (n <= kMaxLeafSize) return;

        // binned SAH over every axis with a non-degenerate centroid extent
        int bestAxis = -1, bestSplit = -1;
        float bestCost = FLT_MAX;
        for (int axis = 0; axis < dim; axis++) {
            float extent = cbox.hi[axis] - cbox.lo[axis];
            if (!(extent > 1e-12f)) continue;
            float scale = kBins/extent;
            Box bins[kBins];
            int counts[kBins];
            for (int b = 0; b < kBins; b++) { bins[b].reset(); counts[b] = 0; }
            for (int i = start; i < end; i++) {
                int p = order[i];
                int b = std::min(kBins - 1, std::max(0, (int)((centroid[3*p + axis] - cbox.lo[axis])*scale)));
                bins[b].grow(primBox[p]);
                counts[b]++;
            }
            float rightArea[kBins];
            int rightCount[kBins];
            Box acc; acc.reset();
            int c = 0;
            for (int b = kBins - 1; b > 0; b--) {
                if (counts[b] > 0) acc.grow(bins[b]);
                c += counts[b];
                rightArea[b] = c > 0 ? acc.halfArea(dim) : 0.0f;
                rightCount[b] = c;
            }
            acc.reset(); c = 0;
            for (int b = 0; b < kBins - 1; b++) {
                if (counts[b] > 0) acc.grow(bins[b]);
                c += counts[b];
                if (c == 0 || rightCount[b + 1] == 0) continue;
                float cost = acc.halfArea(dim)*c + rightArea[b + 1]*rightCount[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestSplit = b; }
            }
        }

        int mid;
        if (bestAxis >= 0) {
            float extent = cbox.hi[bestAxis] - cbox.lo[bestAxis];
            float scale = kBins/extent;
            float lo = cbox.lo[bestAxis];
            int axis = bestAxis, split = bestSplit;
            auto it = std::partition(order.begin() + start, order.begin() + end, [&](int p) {
                int b = std::min(kBins - 1, std::max(0, (int)((centroid[3*p + axis] - lo)*scale)));
                return b <= split;
            });
            mid = (int)(it - order.begin());
        } else {
            mid = (start + end)/2; // coincident centroids: split the index range
        }
        if (mid == start || mid == end) mid = (start + end)/2;

        int l = alloc(), r = alloc();
        nodes[nodeIdx].left = l;
        nodes[nodeIdx].right = r;
        bool spawn = n > 32768 && liveTasks.load() < maxTasks;
        if (spawn) {
            liveTasks.fetch_add(1);
            auto fut = std::async(std::launch::async, [this, l, start, mid, depth]() { build(l, start, mid, depth + 1); });
            build(r, mid, end, depth + 1);
            fut.get();
            liveTasks.fetch_sub(1);
        } else {
            build(l, start, mid, depth + 1);
            build(r, mid, end, depth + 1);
        }
    }
};

inline float scaleFromExp(uint32_t e) { uint32_t bits = e << 23; float f; std::memcpy(&f, &bits, 4); return f; }

// smallest biased exponent e such that 255 * 2^(e-127) >= extent
inline uint32_t chooseExp(float extent) {
    if (!(extent > 0.0f)) return 1;
    int ex;
    std::frexp((double)extent/255.0, &ex); // extent/255 = m * 2^ex, m in [0.5, 1)
    int e = ex + 127;                        // 2^ex >= extent/255
    return (uint32_t)std::min(254, std::max(1, e));
}

inline void cross3(const float *a, const float *b, float *c) {
    c[0] = a[1]*b[2] - a[2]*b[1];
    c[1] = a[2]*b[0] - a[0]*b[2];
    c[2] = a[0]*b[1] - a[1]*b[0];
}
inline float dot3(const float *a, const float *b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
inline float norm3(const float *a) { return std::sqrt(dot3(a, a)); }

} // namespace

void decodeChildBox(const WideNode& node, int slot, float lo[3], float hi[3])
{
    for (int k = 0; k < 3; k++) {
        float o; std::memcpy(&o, &node.w[k], 4);
        float s = scaleFromExp((node.w[3] >> (8*k)) & 0xff);
        const uint8_t *qlo = reinterpret_cast<const uint8_t *>(&node.w[8 + 2*k]);
        const uint8_t *qhi = reinterpret_cast<const uint8_t *>(&node.w[14 + 2*k]);
        lo[k] = o + s*(float)qlo[slot];
        hi[k] = o + s*(float)qhi[slot];
    }
}

namespace {

struct Geometry {
    int dim;
    const float *pos;
    const int32_t *idx;
    void vertex(int prim, int k, float *p) const {
        const float *q = pos + (size_t)dim*idx[(size_t)dim*prim + k];
        p[0] = q[0]; p[1] = q[1]; p[2] = dim == 3 ? q[2] : 0.0f;
    }
    // unit normal: triangles (b-a)x(c-a) (fcpw/geometry/triangles.inl normal()); segments (s.y,-s.x)
    void normal(int prim, float *n) const {
        float a[3], b[3], c[3];
        vertex(prim, 0, a); vertex(prim, 1, b);
        if (dim == 3) {
            vertex(prim, 2, c);
            float u[3] = {b[0]-a[0], b[1]-a[1], b[2]-a[2]}, v[3] = {c[0]-a[0], c[1]-a[1], c[2]-a[2]};
            cross3(u, v, n);
        } else {
            n[0] = b[1] - a[1]; n[1] = -(b[0] - a[0]); n[2] = 0.0f;
        }
        float l = norm3(n);
        if (l > 0.0f) { n[0] /= l; n[1] /= l; n[2] /= l; }
    }
};

// adjacency exactly as FcpwRobinBoundaryHandler<2/3>::buildAccelerationStructure derives it
// (fcpw_geometric_queries.h:728-795, 892-973): directed-edge / vertex tables, last writer wins.
void computeAdjacency(const Geometry& g, int nV, int nPrim, const ReflectingOptions& opt,
                      std::vector<PrimAdjacency>& adj /* input order */)
{
    adj.assign(nPrim, PrimAdjacency());
    std::vector<float> normals((size_t)3*nPrim);
    for (int i = 0; i < nPrim; i++) g.normal(i, &normals[3*(size_t)i]);
    for (int i = 0; i < nPrim; i++) {
        std::memset(&adj[i], 0, sizeof(PrimAdjacency));
        adj[i].minCoeff = opt.minCoeff ? opt.minCoeff[i] : 0.0f;
        adj[i].maxCoeff = opt.maxCoeff ? opt.maxCoeff[i] : 0.0f;
    }

    if (g.dim == 2) {
        std::vector<int> endsAt(nV, -1), startsAt(nV, -1); // slot[0] (j==1 writes) / slot[1] (j==0 writes)
        for (int i = 0; i < nPrim; i++) {
            startsAt[g.idx[2*i + 0]] = i; // adjacentFaceIndices[1 - 0]
            endsAt[g.idx[2*i + 1]] = i;   // adjacentFaceIndices[1 - 1]
        }
        for (int i = 0; i < nPrim; i++) {
            const float *n0 = &normals[3*(size_t)i];
            for (int j = 0; j < 2; j++) {
                int v = g.idx[2*i + j];
                int other = j == 0 ? endsAt[v] : startsAt[v]; // adjacentFaceIndices[j]
                if (other < 0) continue;
                const float *nj = &normals[3*(size_t)other];
                adj[i].flags |= 1u << j;
                for (int k = 0; k < 3; k++) adj[i].n[j][k] = nj[k];
                float det = n0[0]*nj[1] - n0[1]*nj[0];
                float sign = j == 0 ? 1.0f : -1.0f;
                bool ignore = opt.solveDoubleSided ? false : det*sign < opt.silhouettePrecision;
                if (ignore) adj[i].flags |= 1u << (4 + j);
            }
        }
        return;
    }

    struct HalfEdge { int a, b, swapped, tri; };
    std::vector<HalfEdge> edges((size_t)3*nPrim);
    for (int i = 0; i < nPrim; i++) {
        for (int j = 0; j < 3; j++) {
            int I = g.idx[3*i + j], J = g.idx[3*i + (j + 1)%3];
            int sw = 0;
            if (I > J) { std::swap(I, J); sw = 1; }
            edges[(size_t)3*i + j] = {I, J, sw, i};
        }
    }
    std::vector<int> perm(edges.size());
    for (size_t i = 0; i < perm.size(); i++) perm[i] = (int)i;
    std::sort(perm.begin(), perm.end(), [&](int x, int y) {
        const HalfEdge& ex = edges[x]; const HalfEdge& ey = edges[y];
        if (ex.a != ey.a) return ex.a < ey.a;
        if (ex.b != ey.b) return ex.b < ey.b;
        if (ex.swapped != ey.swapped) return ex.swapped < ey.swapped;
        return ex.tri < ey.tri;
    });
    // for every half edge, the face in the opposite slot of its undirected edge (last writer = max tri)
    std::vector<int> opposite(edges.size(), -1);
    size_t s = 0;
    while (s < perm.size()) {
        size_t e = s;
        int last[2] = {-1, -1};
        while (e < perm.size() && edges[perm[e]].a == edges[perm[s]].a && edges[perm[e]].b == edges[perm[s]].b) {
            last[edges[perm[e]].swapped] = edges[perm[e]].tri;
            e++;
        }
        for (size_t k = s; k < e; k++) opposite[perm[k]] = last[1 - edges[perm[k]].swapped];
        s = e;
    }
    for (int i = 0; i < nPrim; i++) {
        const float *n0 = &normals[3*(size_t)i];
        for (int j = 0; j < 3; j++) {
            int other = opposite[(size_t)3*i + j];
            if (other < 0) continue;
            const float *nj = &normals[3*(size_t)other];
            adj[i].flags |= 1u << j;
            for (int k = 0; k < 3; k++) adj[i].n[j][k] = nj[k];
            float pa[3], pb[3];
            g.vertex(i, j, pa); g.vertex(i, (j + 1)%3, pb);
            float ed[3] = {pb[0]-pa[0], pb[1]-pa[1], pb[2]-pa[2]};
            float l = norm3(ed);
            if (l > 0.0f) { ed[0] /= l; ed[1] /= l; ed[2] /= l; }
            float c[3];
            cross3(nj, n0, c);
            float dihedral = std::atan2(dot3(ed, c), dot3(n0, nj));
            bool ignore = opt.solveDoubleSided ? false : dihedral < opt.silhouettePrecision;
            if (ignore) adj[i].flags |= 1u << (4 + j);
        }
    }
}

// normal cone + coefficient range of the primitives order[start, start+count) about `centroid`
// (reflectance_boundary_bvh/bvh.h:77-155)
ChildCone computeCone(const Geometry& g, const std::vector<PrimAdjacency>& adj, const std::vector<float>& normals,
                      const int *prims, int count, const float *centroid)
{
    ChildCone cone;
    std::memset(&cone, 0, sizeof(cone));
    cone.halfAngle = kPi;
    cone.minCoeff = FLT_MAX;
    cone.maxCoeff = -FLT_MAX;
    float axis[3] = {0, 0, 0};
    bool allAdjacent = true, anyTestableEdge = false;
    int nEdges = g.dim;
    for (int i = 0; i < count; i++) {
        int p = prims[i];
        const float *n = &normals[3*(size_t)p];
        axis[0] += n[0]; axis[1] += n[1]; axis[2] += n[2];
        cone.minCoeff = std::min(cone.minCoeff, adj[p].minCoeff);
        cone.maxCoeff = std::max(cone.maxCoeff, adj[p].maxCoeff);
        for (int k = 0; k < nEdges; k++) {
            float q[3];
            if (g.dim == 2) {
                g.vertex(p, k, q);
            } else {
                float a[3], b[3];
                g.vertex(p, k, a); g.vertex(p, (k + 1)%3, b);
                for (int c = 0; c < 3; c++) q[c] = 0.5f*(a[c] + b[c]);
            }
            float d[3] = {q[0]-centroid[0], q[1]-centroid[1], q[2]-centroid[2]};
            cone.radius = std::max(cone.radius, norm3(d));
            allAdjacent = allAdjacent && ((adj[p].flags >> k) & 1u);
            anyTestableEdge = anyTestableEdge || !((adj[p].flags >> (4 + k)) & 1u);
        }
    }
    // every edge (vertex in 2D) below carries the ignore mark: no primitive test in this subtree can ever report a
    // silhouette, whatever the normal cone says
    cone.noTestableEdge = anyTestableEdge ? 0.0f : 1.0f;
    if (allAdjacent) {
        float l = norm3(axis);
        if (l > FLT_EPSILON) {
            axis[0] /= l; axis[1] /= l; axis[2] /= l;
            float half = 0.0f;
            for (int i = 0; i < count; i++) {
                int p = prims[i];
                const float *n = &normals[3*(size_t)p];
                half = std::max(half, std::acos(std::max(-1.0f, std::min(1.0f, dot3(axis, n)))));
                for (int k = 0; k < nEdges; k++) {
                    half = std::max(half, std::acos(std::max(-1.0f, std::min(1.0f, dot3(axis, adj[p].n[k])))));
                }
            }
            cone.axis[0] = axis[0]; cone.axis[1] = axis[1]; cone.axis[2] = axis[2];
            // pad: the device test must stay conservative under float rounding of acos/asin/atan2
            cone.halfAngle = std::min(kPi, half + 1e-4f);
        }
    }
    return cone;
}

} // namespace

bool buildWideBvh(int dim, const float *positions, int nV, const int32_t *indices, int nPrim,
                  const ReflectingOptions& refl, HostBvh& out, const char **err)
{
    static const char *eDim = "dim must be 2 or 3";
    static const char *eIdx = "primitive index out of range";
    static const char *eEmpty = "empty mesh";
    static const char *eNan = "non-finite vertex position";
    if (dim != 2 && dim != 3) { if (err) *err = eDim; return false; }
    if (nPrim <= 0 || nV <= 0) { if (err) *err = eEmpty; return false; }
    for (size_t i = 0; i < (size_t)nPrim*dim; i++) {
        if (indices[i] < 0 || indices[i] >= nV) { if (err) *err = eIdx; return false; }
    }
    for (size_t i = 0; i < (size_t)nV*dim; i++) {
        if (!std::isfinite(positions[i])) { if (err) *err = eNan; return false; }
    }

    Geometry g{dim, positions, indices};
    Builder b;
    b.dim = dim;
    b.nPrim = nPrim;
    b.primBox.resize(nPrim);
    b.centroid.resize((size_t)3*nPrim);
    b.order.resize(nPrim);
    for (int i = 0; i < nPrim; i++) {
        Box bx; bx.reset();
        float c[3] = {0, 0, 0};
        for (int k = 0; k < dim; k++) {
            float p[3];
            g.vertex(i, k, p);
            bx.grow(p);
            c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
        }
        b.primBox[i] = bx;
        for (int k = 0; k < 3; k++) b.centroid[3*(size_t)i + k] = c[k]/dim;
        b.order[i] = i;
    }
    b.nodes.resize((size_t)2*nPrim + 1);
    b.maxTasks = std::max(1u, std::thread::hardware_concurrency());
    int root = b.alloc();
    b.build(root, 0, nPrim, 0);

    // ---- collapse into 8-ary nodes, breadth first so that internal children are contiguous ----
    out = HostBvh();
    out.dim = dim;
    out.nPrims = nPrim;
    const int primFloats = dim == 3 ? 12 : 4;
    out.prims.assign((size_t)primFloats*nPrim, 0.0f);
    out.primIndex.assign(nPrim, -1);
    for (int k = 0; k < 3; k++) { out.boxMin[k] = b.nodes[root].box.lo[k]; out.boxMax[k] = b.nodes[root].box.hi[k]; }

    std::vector<PrimAdjacency> adjInput;
    std::vector<float> normals;
    if (refl.enabled) {
        computeAdjacency(g, nV, nPrim, refl, adjInput);
        normals.resize((size_t)3*nPrim);
        for (int i = 0; i < nPrim; i++) g.normal(i, &normals[3*(size_t)i]);
        out.adjacency.resize(nPrim);
    }

    struct Pending { int binary; int wide; int depth; };
    std::vector<Pending> queue;
    out.nodes.emplace_back();
    queue.push_back({root, 0, 1});
    int primCursor = 0;
    std::vector<int> wideToChildrenBinary; // 8 per wide node, -1 if empty (for cone pass)
    wideToChildrenBinary.reserve(1024);

    for (size_t qi = 0; qi < queue.size(); qi++) {
        Pending cur = queue[qi];
        out.maxDepth = std::max(out.maxDepth, cur.depth);
        int children[kWidth];
        int nChildren = 0;
        const BinaryNode& bn = b.nodes[cur.binary];
        if (bn.left < 0) {
            children[nChildren++] = cur.binary; // degenerate: the whole mesh is one leaf
        } else {
            children[nChildren++] = bn.left;
            children[nChildren++] = bn.right;
            while (nChildren < kWidth) {
                int best = -1;
                float bestArea = -1.0f;
                for (int c = 0; c < nChildren; c++) {
                    const BinaryNode& cn = b.nodes[children[c]];
                    if (cn.left < 0) continue;
                    float a = cn.box.halfArea(dim);
                    if (a > bestArea) { bestArea = a; best = c; }
                }
                if (best < 0) break;
                int expand = children[best];
                children[best] = b.nodes[expand].left;
                children[nChildren++] = b.nodes[expand].right;
            }
        }

        // quantisation frame of this node
        Box nb; nb.reset();
        for (int c = 0; c < nChildren; c++) nb.grow(b.nodes[children[c]].box);
        WideNode wn;
        std::memset(&wn, 0, sizeof(wn));
        uint32_t exps[3];
        float scale[3];
        for (int k = 0; k < 3; k++) {
            std::memcpy(&wn.w[k], &nb.lo[k], 4);
            exps[k] = chooseExp(nb.hi[k] - nb.lo[k]);
            // make sure the far side is representable after float rounding
            while (exps[k] < 254 && nb.lo[k] + scaleFromExp(exps[k])*255.0f < nb.hi[k]) exps[k]++;
            scale[k] = scaleFromExp(exps[k]);
        }
        wn.w[3] = exps[0] | (exps[1] << 8) | (exps[2] << 16);
        uint8_t *meta = reinterpret_cast<uint8_t *>(&wn.w[6]);
        int childBase = (int)out.nodes.size();
        int primBase = primCursor;
        int nInternal = 0, primOffset = 0;
        size_t wbase = wideToChildrenBinary.size();
        wideToChildrenBinary.resize(wbase + kWidth, -1);
        for (int c = 0; c < kWidth; c++) {
            uint8_t *qlo[3], *qhi[3];
            for (int k = 0; k < 3; k++) {
                qlo[k] = reinterpret_cast<uint8_t *>(&wn.w[8 + 2*k]);
                qhi[k] = reinterpret_cast<uint8_t *>(&wn.w[14 + 2*k]);
            }
            if (c >= nChildren) {
                for (int k = 0; k < 3; k++) { qlo[k][c] = 255; qhi[k][c] = 0; }
                meta[c] = 0;
                continue;
            }
            const BinaryNode& cn = b.nodes[children[c]];
            wideToChildrenBinary[wbase + c] = children[c];
            for (int k = 0; k < 3; k++) {
                double rel = ((double)cn.box.lo[k] - (double)nb.lo[k])/(double)scale[k];
                int lo = (int)std::floor(rel);
                lo = std::min(255, std::max(0, lo));
                while (lo > 0 && nb.lo[k] + scale[k]*(float)lo > cn.box.lo[k]) lo--;
                rel = ((double)cn.box.hi[k] - (double)nb.lo[k])/(double)scale[k];
                int hi = (int)std::ceil(rel);
                hi = std::min(255, std::max(0, hi));
                while (hi < 255 && nb.lo[k] + scale[k]*(float)hi < cn.box.hi[k]) hi++;
                qlo[k][c] = (uint8_t)lo;
                qhi[k][c] = (uint8_t)hi;
            }
            if (cn.left < 0) {
                meta[c] = (uint8_t)((cn.count << 5) | primOffset);
                for (int i = 0; i < cn.count; i++) {
                    int p = b.order[cn.start + i];
                    int dst = primBase + primOffset + i;
                    out.primIndex[dst] = p;
                    float *pf = &out.prims[(size_t)primFloats*dst];
                    if (dim == 3) {
                        for (int v = 0; v < 3; v++) {
                            float q[3];
                            g.vertex(p, v, q);
                            pf[4*v + 0] = q[0]; pf[4*v + 1] = q[1]; pf[4*v + 2] = q[2];
                        }
                        std::memcpy(&pf[3], &p, 4); // v0.w carries the caller's primitive index
                    } else {
                        float q[3];
                        g.vertex(p, 0, q); pf[0] = q[0]; pf[1] = q[1];
                        g.vertex(p, 1, q); pf[2] = q[0]; pf[3] = q[1];
                    }
                    if (refl.enabled) {
                        out.adjacency[dst] = adjInput[p];
                        if (dim == 3) { // filter-stage data rides in the free lanes of the vertex record (one 48-byte read)
                            pf[7] = adjInput[p].maxCoeff;
                            std::memcpy(&pf[11], &adjInput[p].flags, 4);
                        }
                    }
                }
                primOffset += cn.count;
                out.nLeaves++;
            } else {
                meta[c] = (uint8_t)(nInternal + 1);
                nInternal++;
            }
        }
        primCursor += primOffset;
        wn.w[4] = (uint32_t)childBase;
        wn.w[5] = (uint32_t)primBase;
        out.nodes[cur.wide] = wn;
        int rank = 0;
        for (int c = 0; c < nChildren; c++) {
            if (b.nodes[children[c]].left >= 0) {
                out.nodes.emplace_back();
                queue.push_back({children[c], childBase + rank, cur.depth + 1});
                rank++;
            }
        }
    }

    if (refl.enabled) {
        out.cones.resize(out.nodes.size()*kWidth);
        size_t nWide = out.nodes.size();
        auto work = [&](size_t lo, size_t hi) {
            for (size_t w = lo; w < hi; w++) {
                for (int c = 0; c < kWidth; c++) {
                    ChildCone cone;
                    std::memset(&cone, 0, sizeof(cone));
                    cone.halfAngle = kPi;
                    int bnIdx = wideToChildrenBinary[w*kWidth + c];
                    if (bnIdx >= 0) {
                        const BinaryNode& cn = b.nodes[bnIdx];
                        float blo[3], bhi[3], cen[3];
                        decodeChildBox(out.nodes[w], c, blo, bhi);
                        for (int k = 0; k < 3; k++) cen[k] = 0.5f*(blo[k] + bhi[k]);
                        cone = computeCone(g, adjInput, normals, &b.order[cn.start], cn.count, cen);
                    }
                    out.cones[w*kWidth + c] = cone;
                }
            }
        };
        int nThreads = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 1 + nWide/64);
        if (nThreads <= 1) {
            work(0, nWide);
        } else {
            std::vector<std::thread> threads;
            size_t chunk = (nWide + nThreads - 1)/nThreads;
            for (int t = 0; t < nThreads; t++) {
                size_t lo = t*chunk, hi = std::min(nWide, lo + chunk);
                if (lo < hi) threads.emplace_back(work, lo, hi);
            }
            for (auto& th : threads) th.join();
        }
    }
    return true;
}

long long validateWideBvh(const HostBvh& bvh, int dim, const float *positions, const int32_t *indices)
{
    long long bad = 0;
    std::vector<int> seen(bvh.nPrims, 0);
    std::vector<int> parentCount(bvh.nodes.size(), 0);
    // boxes of all ancestors' child slots on the path to a node: checked top-down with an explicit stack
    struct Item { uint32_t node; std::vector<float> boxes; }; // boxes: 6 floats per enclosing slot
    std::vector<Item> stack;
    stack.push_back({0u, {}});
    while (!stack.empty()) {
        Item it = std::move(stack.back());
        stack.pop_back();
        const WideNode& wn = bvh.nodes[it.node];
        const uint8_t *meta = reinterpret_cast<const uint8_t *>(&wn.w[6]);
        int nInternal = 0;
        for (int c = 0; c < kWidth; c++) {
            uint8_t m = meta[c];
            if (m == 0) continue;
            float lo[3], hi[3];
            decodeChildBox(wn, c, lo, hi);
            std::vector<float> boxes = it.boxes;
            boxes.insert(boxes.end(), lo, lo + 3);
            boxes.insert(boxes.end(), hi, hi + 3);
            int count = m >> 5;
            if (count > 0) {
                if (count > kMaxLeafSize) bad++;
                int first = (int)wn.w[5] + (m & 31);
                for (int i = 0; i < count; i++) {
                    int slot = first + i;
                    if (slot < 0 || slot >= bvh.nPrims) { bad++; continue; }
                    seen[slot]++;
                    int p = bvh.primIndex[slot];
                    for (int v = 0; v < dim; v++) {
                        const float *q = positions + (size_t)dim*indices[(size_t)dim*p + v];
                        for (size_t b = 0; b < boxes.size(); b += 6) {
                            for (int k = 0; k < dim; k++) {
                                if (q[k] < boxes[b + k] || q[k] > boxes[b + 3 + k]) bad++;
                            }
                        }
                        // the de-indexed copy the device reads must be the caller's vertex
                        const float *pf = &bvh.prims[(size_t)(dim == 3 ? 12 : 4)*slot + (dim == 3 ? 4*v : 2*v)];
                        for (int k = 0; k < dim; k++) if (pf[k] != q[k]) bad++;
                    }
                }
            } else {
                uint32_t child = wn.w[4] + (m & 31) - 1;
                if ((m & 31) - 1 != nInternal) bad++; // internal children are ranked in slot order
                nInternal++;
                if (child >= bvh.nodes.size() || child <= it.node) { bad++; continue; }
                parentCount[child]++;
                stack.push_back({child, std::move(boxes)});
            }
        }
    }
    for (int i = 0; i < bvh.nPrims; i++) if (seen[i] != 1) bad++;
    for (size_t i = 1; i < bvh.nodes.size(); i++) if (parentCount[i] != 1) bad++;
    return bad;
}

void updateCoefficients(HostBvh& bvh, const float *minCoeff, const float *maxCoeff)
{
    if (bvh.adjacency.empty()) return;
    for (int i = 0; i < bvh.nPrims; i++) {
        int p = bvh.primIndex[i];
        bvh.adjacency[i].minCoeff = minCoeff[p];
        bvh.adjacency[i].maxCoeff = maxCoeff[p];
        if (bvh.dim == 3) bvh.prims[(size_t)12*i + 7] = maxCoeff[p]; // the copy in the vertex record (v1.w)
    }
    // children precede... no: nodes are breadth first, so children have larger indices than parents;
    // sweep backwards accumulating each node's total range into its parent's child slot
    size_t nWide = bvh.nodes.size();
    std::vector<float> nodeMin(nWide, FLT_MAX), nodeMax(nWide, -FLT_MAX);
    for (size_t w = nWide; w-- > 0;) {
        const WideNode& wn = bvh.nodes[w];
        const uint8_t *meta = reinterpret_cast<const uint8_t *>(&wn.w[6]);
        for (int c = 0; c < kWidth; c++) {
            uint8_t m = meta[c];
            if (m == 0) continue;
            float mn = FLT_MAX, mx = -FLT_MAX;
            int count = m >> 5;
            if (count > 0) {
                int first = (int)wn.w[5] + (m & 31);
                for (int i = 0; i < count; i++) {
                    mn = std::min(mn, bvh.adjacency[first + i].minCoeff);
                    mx = std::max(mx, bvh.adjacency[first + i].maxCoeff);
                }
            } else {
                size_t child = wn.w[4] + (m & 31) - 1;
                mn = nodeMin[child];
                mx = nodeMax[child];
            }
            bvh.cones[w*kWidth + c].minCoeff = mn;
            bvh.cones[w*kWidth + c].maxCoeff = mx;
            nodeMin[w] = std::min(nodeMin[w], mn);
            nodeMax[w] = std::max(nodeMax[w], mx);
        }
    }
}

} // namespace zb
