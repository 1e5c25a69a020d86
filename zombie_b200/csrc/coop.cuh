// Warp-cooperative BVH traversal for the two queries whose cost is heavy-tailed: closest point and
// star radius. Measured on the device code compiled for the host (tools/host_probe.cu, 1.3 M triangles):
// the median star-radius query of a walk touches ~5 nodes, the mean 136 nodes + 736 primitives, the 99th
// percentile ~2000 nodes + 9000 primitives, so with one thread per query a warp idles on its slowest lane
// (lane efficiency ~0.1).
//
// Here the (up to) 32 queries of a warp share the work: the query state lives in shared memory and the warp
// keeps TWO work lists there,
//   * a stack of (owner query, node) items, and
//   * a queue of (owner query, primitive) items,
// and runs homogeneous rounds: a NODE round pops up to 32 nodes, every lane tests the 8 children of one node
// against its owner's query and the survivors are appended with warp prefix sums (internal children to the node
// stack, the primitives of leaf children to the primitive queue); a PRIMITIVE round pops 32 primitives, one per
// lane. Rounds never mix the two code paths, so a warp does not pay for both on every iteration, and a leaf
// with 4 primitives occupies 4 lanes for one round instead of one lane for four. All lanes stay busy until
// fewer than 32 items remain, whatever the skew between the 32 queries. Results are combined with shared-memory
// atomicMin on (squared radius, primitive), which makes them independent of the traversal order (ties on
// distance resolve to the smaller leaf-order index).
//
// Node tests only ever PRUNE (they bound what a subtree can contribute), so they are free to use fast
// algebraic forms — the cone test below replaces the reference's acos/asin/atan2 (fcpw/core/bounding_volumes.h:
// 275-321) by products of sines and cosines with explicit slack — as long as they stay conservative; the values
// a query returns come from the primitive tests alone.
#pragma once

#include "queries.cuh"

// Warp intrinsics behind macros so that tools/coop_check.cu can run this file on a CPU with 32 host threads
// standing in for the lanes of a warp (tools/warp_emu.h); the library itself only ever compiles the CUDA side.
#if defined(ZB_EMULATE_WARP)
#include "../../tools/warp_emu.h"
#else
#define ZB_COOP __device__ __forceinline__
#define ZB_COOP_NOINLINE __device__
#define zb_lane() ((int)(threadIdx.x & 31))
#define zb_syncwarp() __syncwarp()
#define zb_ballot(p) __ballot_sync(0xffffffffu, (p))
#define zb_shfl(v, src) __shfl_sync(0xffffffffu, (v), (src))
#define zb_shfl_up(v, o) __shfl_up_sync(0xffffffffu, (v), (o))
#define zb_atomic_min64(p, v) atomicMin((p), (v))
#define zb_status_or(p, v) atomicOr((p), (v))
#define zb_sincos_fast(x, s, c) __sincosf((x), (s), (c))
#define zb_rsqrt(x) rsqrtf(x)
#define zb_exp_fast(x) __expf(x)
#define zb_popc(x) __popc(x)
#endif

namespace zb {

constexpr int kNodeStack = 640;       // node items per warp; the top kNodeHeadroom are depth-first headroom
constexpr int kNodeHeadroom = 128;
constexpr int kPrimQueue = 1056;      // >= 31 + 32 lanes x 8 leaves x 4 primitives
constexpr uint32_t kPrimIndexMask = 0x03ffffffu; // 26 bits: up to 67 M primitives per boundary
enum : int { kStatusCoopOverflow = 4 };

struct WarpShared {
    uint2 nodes[kNodeStack];            // x: lower-bound key (float bits, low 6 bits = noSil | owner lane); y: node index
    uint32_t prims[kPrimQueue];         // owner lane << 27 | noSil << 26 | primitive (leaf order)
    unsigned long long best[32];        // high word: current squared radius (float bits); low word: payload
    float qx[32], qy[32], qz[32];
    unsigned int flags[32];
};

ZB_COOP unsigned long long packBest(float d2, uint32_t payload)
{
    return ((unsigned long long)zb_float_as_uint(d2) << 32) | payload;
}

ZB_COOP float coopRadius2(const WarpShared& ws, int owner)
{
    return zb_uint_as_float((uint32_t)(*(volatile const unsigned long long *)&ws.best[owner] >> 32));
}

// inclusive warp prefix sum
ZB_COOP int warpInclusiveSum(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = zb_shfl_up(v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// how many node items the next node round may pop: bounded by the room their children need on the node stack
// (7 net entries per popped node) — the last kNodeHeadroom entries are reserved for the depth-first tail, where
// one node is popped per round and the stack cannot grow by more than 7 per tree level
ZB_COOP int coopNodeBudget(int nodeTop)
{
    int room = (kNodeStack - kNodeHeadroom - nodeTop)/7;
    return min(min(nodeTop, 32), max(room, 1));
}

// Appends the surviving children of this lane's node (keys[0..nKeep) ascending; low 3 bits = child slot) to the
// work lists: internal children far-to-near on the node stack (the nearest ends on top of the lane's group),
// the primitives of leaf children far-to-near on the primitive queue. tag = noSil bits (per slot) for the
// star-radius query, 0 for closest point. Every lane must call this (lanes without a node pass nKeep = 0).
ZB_COOP void coopPushChildren(WarpShared& ws, int& nodeTop, int& primTop, int lane, int owner,
                                                 const NodeRegs& node, const uint32_t keys[8], int nKeep, uint32_t noSil, int *status)
{
    int nNodes = 0, nPrims = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        if (c < nKeep) {
            uint32_t meta = childMeta(node, keys[c] & 7u);
            uint32_t count = meta >> 5;
            nNodes += count == 0u ? 1 : 0;
            nPrims += (int)count;
        }
    }
    int inclN = warpInclusiveSum(nNodes, lane), inclP = warpInclusiveSum(nPrims, lane);
    int totalN = zb_shfl(inclN, 31), totalP = zb_shfl(inclP, 31);
    if (nodeTop + totalN > kNodeStack || primTop + totalP > kPrimQueue) { // warp-uniform; cannot happen for trees of depth <= 18
        if (lane == 0 && status) zb_status_or(status, kStatusCoopOverflow);
        return;
    }
    int slotN = nodeTop + inclN - nNodes, slotP = primTop + inclP - nPrims;
#pragma unroll
    for (int c = 7; c >= 0; c--) {
        if (c < nKeep) {
            uint32_t key = keys[c], slot = key & 7u;
            uint32_t meta = childMeta(node, slot);
            uint32_t count = meta >> 5;
            uint32_t sil = (noSil >> slot) & 1u;
            if (count == 0u) {
                ws.nodes[slotN++] = make_uint2((key & ~63u) | (sil << 5) | (uint32_t)owner, node.q[1].x + (meta & 31u) - 1u);
            } else {
                uint32_t first = node.q[1].y + (meta & 31u);
                uint32_t head = ((uint32_t)owner << 27) | (sil << 26);
                for (uint32_t i = count; i > 0u; i--) ws.prims[slotP++] = head | (first + i - 1u);
            }
        }
    }
    nodeTop += totalN;
    primTop += totalP;
}

// Aggregate::findClosestPoint for up to 32 queries at once (one per lane with hasQuery). Every lane of
// the warp must call this. r2max: squared search radius. On return hit describes the lane's own query.
template <int DIM>
ZB_COOP_NOINLINE void coopClosestPoint(const BvhView& bvh, WarpShared& ws, bool hasQuery, const Vec<DIM>& x, float r2max,
                                 ClosestHit<DIM>& hit, int *status)
{
    const int lane = zb_lane();
    zb_syncwarp();
    ws.qx[lane] = x[0];
    ws.qy[lane] = x[1];
    ws.qz[lane] = DIM == 3 ? x[DIM - 1] : 0.0f;
    ws.best[lane] = packBest(r2max, 0xffffffffu);
    unsigned m = zb_ballot(hasQuery);
    int nodeTop = zb_popc(m), primTop = 0;
    if (hasQuery) ws.nodes[zb_popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)lane, 0u);
    zb_syncwarp();

    while (nodeTop > 0 || primTop > 0) {
        if (primTop >= 32 || nodeTop == 0) {
            // ---- primitive round: one primitive per lane
            int n = min(primTop, 32);
            if (lane < n) {
                uint32_t item = ws.prims[primTop - 1 - lane];
                int owner = (int)(item >> 27);
                uint32_t idx = item & kPrimIndexMask;
                float r2 = coopRadius2(ws, owner);
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                Prim<DIM> prim;
                prim.load(bvh, idx);
                Vec<DIM> pt;
                float u, v;
                float d = prim.closest(q, pt, u, v);
                float d2 = d*d;
                if (d2 <= r2) zb_atomic_min64(&ws.best[owner], packBest(d2, idx));
            }
            primTop -= n;
            zb_syncwarp();
            continue;
        }

        // ---- node round: one node per lane
        int n = coopNodeBudget(nodeTop);
        bool have = lane < n;
        uint2 item = make_uint2(0u, 0u);
        if (have) item = ws.nodes[nodeTop - 1 - lane];
        nodeTop -= n;
        zb_syncwarp();

        uint32_t keys[8];
#pragma unroll
        for (int c = 0; c < 8; c++) keys[c] = 0xffffffffu;
        int nKeep = 0;
        NodeRegs node;
        int owner = item.x & 31u;
        if (have) {
            float r2 = coopRadius2(ws, owner);
            if (zb_uint_as_float(item.x & ~63u) <= r2) {
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                loadNode(bvh.nodes, item.y, node);
                NodeFrame<DIM> f = nodeFrame<DIM>(node);
                float shrink = r2;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    float lo[DIM], hi[DIM], d2min, d2max;
                    childBox<DIM>(node, f, c, lo, hi);
                    boxDistance2<DIM>(lo, hi, q, d2min, d2max);
                    bool valid = childMeta(node, c) != 0u;
                    if (valid) shrink = fminf(shrink, d2max);
                    // keys are rounded DOWN to a multiple of 64 ulps: still a lower bound of the child's distance
                    keys[c] = valid ? ((zb_float_as_uint(d2min) & ~63u) | (uint32_t)c) : 0xffffffffu;
                }
                if (shrink < r2) {
                    // the search radius also shrinks by each child's farthest-corner distance (bvh.inl:1339-1343)
                    zb_atomic_min64(&ws.best[owner], packBest(shrink, 0xffffffffu));
                    r2 = shrink;
                }
                sort8(keys);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (keys[c] != 0xffffffffu && zb_uint_as_float(keys[c] & ~63u) <= r2) nKeep = c + 1;
                }
            }
        }
        coopPushChildren(ws, nodeTop, primTop, lane, owner, node, keys, nKeep, 0u, status);
        zb_syncwarp();
    }

    unsigned long long b = ws.best[lane];
    uint32_t prim = (uint32_t)b;
    hit.prim = -1;
    hit.d = kMaxFloat;
    hit.p = x;
    if (hasQuery && prim != 0xffffffffu) {
        Prim<DIM> p;
        p.load(bvh, prim);
        float u, v;
        hit.d = p.closest(x, hit.p, u, v);
        hit.prim = (int)prim;
    }
    zb_syncwarp();
}

// ------------------------------------------------------------------------------------------
// star radius
// ------------------------------------------------------------------------------------------

// ReflectanceBvh::visitNode (reflectance_boundary_bvh/bvh.h:370-412) for one child of a wide node, in the
// transcendental-free form described at the top of this file. Same contract as starChildBounds (queries.cuh):
// bounds [lower, upper] on the squared star radius anything inside the child can produce, whether a silhouette
// may hide inside; false when the child cannot shrink the radius at all. Conservative with respect to
// starChildBounds' own tests: the overlap threshold carries 2e-5 of slack on the sine, the Robin cosine bound
// 1e-5, on top of the 1e-4 rad by which the builder pads every cone.
template <int DIM>
ZB_COOP bool starChildBoundsFast(const BvhView& bvh, uint32_t nodeIdx, int c, const Vec<DIM>& x,
                                                    const float lo[DIM], const float hi[DIM], float d2min, float d2max,
                                                    float& lower, float& upper, bool& hasSil)
{
    const float4 *cone = bvh.cones + ((size_t)nodeIdx*8 + c)*2;
    float4 c0 = zb_ldg(cone), c1 = zb_ldg(cone + 1);
    float minCoeff = c1.y, maxCoeff = c1.z;
    lower = d2min;
    upper = d2max;
    hasSil = true;
    if (!(minCoeff < kMaxFloat - kEpsilon)) return true; // perfectly absorbing subtree: plain distance bounds

    upper = kMaxFloat;
    float halfAngle = c0.w;
    if (halfAngle >= kHalfPiF || d2min < kEpsilon) return true;
    float view[DIM], l2 = 0.0f;
#pragma unroll
    for (int k = 0; k < DIM; k++) { view[k] = 0.5f*(lo[k] + hi[k]) - x[k]; l2 = fmaf(view[k], view[k], l2); }
    float radius = c1.x;
    if (!(l2 > radius*radius)) return true; // inside the bounding sphere of the child: treat as overlapping
    float invL = zb_rsqrt(l2);
    float cosTheta = 0.0f;
    const float axis[3] = {c0.x, c0.y, c0.z};
#pragma unroll
    for (int k = 0; k < DIM; k++) cosTheta = fmaf(axis[k], view[k], cosTheta);
    cosTheta = fminf(1.0f, fabsf(cosTheta*invL));
    float sinA, cosA;
    zb_sincos_fast(halfAngle, &sinA, &cosA);
    float sinB = fminf(1.0f, radius*invL), cosB = sqrtf(fmaxf(0.0f, 1.0f - sinB*sinB));
    float sinG = sinA*cosB + cosA*sinB, cosG = cosA*cosB - sinA*sinB; // gamma = cone half angle + view half angle
    if (cosG <= 1e-4f || cosTheta <= sinG + 2e-5f) return true;        // pi/2 within theta +- gamma: may hold a silhouette

    hasSil = false;
    if (!(maxCoeff > kEpsilon)) return false; // perfectly reflecting and no silhouette inside
    // RobinBvhNodeBound (robin_bounds.h:102-154): radius range over the whole subtree
    float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta*cosTheta));
    float minCos = fmaxf(0.0f, cosTheta*cosG - sinTheta*sinG - 1e-5f); // min |cos| over [theta - gamma, theta + gamma]
    float rMin = sqrtf(d2min), rMax = sqrtf(d2max);
    if (DIM == 3) {
        if (rMax < minCos/maxCoeff) lower = kMaxFloat;
        else { float rb = rMin/(1.0f - minCos/(maxCoeff*rMax)); lower = rb*rb; }
        if (!(rMin < 1.0f/minCoeff)) { float rb = rMax/(1.0f - 1.0f/(minCoeff*rMin)); upper = rb*rb; }
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

// ReflectanceBvh::computeSquaredStarRadius / findClosestSilhouettePoint for up to 32 queries at once.
// Every lane of the warp must call this; returns the squared star radius (<= r2max) of the lane's query.
template <int DIM>
ZB_COOP_NOINLINE float coopStarRadius2(const BvhView& bvh, WarpShared& ws, bool hasQuery, const Vec<DIM>& x, float r2max,
                                 bool flip, float precision, int *status)
{
    const int lane = zb_lane();
    zb_syncwarp();
    ws.qx[lane] = x[0];
    ws.qy[lane] = x[1];
    ws.qz[lane] = DIM == 3 ? x[DIM - 1] : 0.0f;
    ws.best[lane] = packBest(r2max, 0u);
    ws.flags[lane] = flip ? 1u : 0u;
    unsigned m = zb_ballot(hasQuery);
    int nodeTop = zb_popc(m), primTop = 0;
    if (hasQuery) ws.nodes[zb_popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)lane, 0u);
    zb_syncwarp();

    while (nodeTop > 0 || primTop > 0) {
        if (primTop >= 32 || nodeTop == 0) {
            // ---- primitive round
            int n = min(primTop, 32);
            if (lane < n) {
                uint32_t item = ws.prims[primTop - 1 - lane];
                int owner = (int)(item >> 27);
                // bit 26: "no silhouette inside this subtree" (reflectance bvh.h:474-476 sign trick)
                bool silhouetteTests = !((item >> 26) & 1u);
                float r2 = coopRadius2(ws, owner), before = r2;
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                primStarRadius2<DIM>(bvh, item & kPrimIndexMask, q, r2, ws.flags[owner] != 0u, precision, silhouetteTests);
                if (r2 < before) zb_atomic_min64(&ws.best[owner], packBest(r2, 0u));
            }
            primTop -= n;
            zb_syncwarp();
            continue;
        }

        // ---- node round
        int n = coopNodeBudget(nodeTop);
        bool have = lane < n;
        uint2 item = make_uint2(0u, 0u);
        if (have) item = ws.nodes[nodeTop - 1 - lane];
        nodeTop -= n;
        zb_syncwarp();

        uint32_t keys[8];
#pragma unroll
        for (int c = 0; c < 8; c++) keys[c] = 0xffffffffu;
        int nKeep = 0;
        uint32_t noSil = 0u;
        NodeRegs node;
        int owner = item.x & 31u;
        if (have) {
            float r2 = coopRadius2(ws, owner);
            if (zb_uint_as_float(item.x & ~63u) <= r2) {
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                bool parentNoSil = (item.x >> 5) & 1u; // no silhouette above => none below
                loadNode(bvh.nodes, item.y, node);
                NodeFrame<DIM> f = nodeFrame<DIM>(node);
                float shrink = r2;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint32_t meta = childMeta(node, c);
                    if (meta == 0u) continue;
                    float lo[DIM], hi[DIM], d2min, d2max;
                    childBox<DIM>(node, f, c, lo, hi);
                    boxDistance2<DIM>(lo, hi, q, d2min, d2max);
                    if (d2min > r2) continue;
                    float lower, upper;
                    bool hasSil;
                    if (!starChildBoundsFast<DIM>(bvh, item.y, c, q, lo, hi, d2min, d2max, lower, upper, hasSil)) continue;
                    shrink = fminf(shrink, upper);
                    if (!hasSil || parentNoSil) noSil |= 1u << c;
                    keys[c] = (zb_float_as_uint(lower) & ~63u) | (uint32_t)c;
                }
                if (shrink < r2) {
                    zb_atomic_min64(&ws.best[owner], packBest(shrink, 0u));
                    r2 = shrink;
                }
                sort8(keys);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (keys[c] != 0xffffffffu && zb_uint_as_float(keys[c] & ~63u) <= r2) nKeep = c + 1;
                }
            }
        }
        coopPushChildren(ws, nodeTop, primTop, lane, owner, node, keys, nKeep, noSil, status);
        zb_syncwarp();
    }
    float r2 = coopRadius2(ws, lane);
    zb_syncwarp();
    return hasQuery ? r2 : r2max;
}

} // namespace zb
