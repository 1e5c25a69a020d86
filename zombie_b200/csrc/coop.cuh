// Warp-cooperative BVH traversal for the two queries whose cost is heavy-tailed: closest point and
// star radius. Measured on the device path (tools/host_probe.cu): the median star-radius query of a
// walk touches ~1 node, the 99th percentile ~2500 primitives, so with one thread per query a warp
// idles on its slowest lane (lane efficiency ~0.1). Here the 32 queries of a warp share ONE work
// stack in shared memory: every round each lane pops one (owner, node-or-leaf) item, tests it
// against the OWNER's query (query state lives in shared memory), and the surviving children are
// pushed back with a warp prefix sum. All lanes stay busy until fewer than 32 items remain, whatever
// the skew between the 32 queries. Results are combined with shared-memory atomicMin, which makes
// them independent of the traversal order (ties on distance resolve to the smaller leaf index).
#pragma once

#include "queries.cuh"

namespace zb {

constexpr int kCoopStack = 2048;
enum : int { kStatusCoopOverflow = 4 };

struct WarpShared {
    uint2 stack[kCoopStack];            // x: lower-bound key | flags | owner lane; y: node / leaf reference
    unsigned long long best[32];        // high word: current squared radius (float bits); low word: payload
    float qx[32], qy[32], qz[32];
    unsigned int flags[32];
};

__device__ __forceinline__ unsigned long long packBest(float d2, uint32_t payload)
{
    return ((unsigned long long)__float_as_uint(d2) << 32) | payload;
}

__device__ __forceinline__ float coopRadius2(const WarpShared& ws, int owner)
{
    return __uint_as_float((uint32_t)(*(volatile const unsigned long long *)&ws.best[owner] >> 32));
}

// inclusive warp prefix sum
__device__ __forceinline__ int warpInclusiveSum(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// pushes the sorted child keys of a node (ascending in keys[0..7], 0xffffffff = none) far-to-near, so
// that each lane's nearest child ends on top of its group; returns the new stack height
template <typename MakeItem>
__device__ __forceinline__ int coopPush(WarpShared& ws, int top, int lane, const uint32_t keys[8], int pushes,
                                        MakeItem makeItem, int *status)
{
    int incl = warpInclusiveSum(pushes, lane);
    int total = __shfl_sync(0xffffffffu, incl, 31);
    if (top + total > kCoopStack) { // warp-uniform
        if (lane == 0 && status) atomicOr(status, kStatusCoopOverflow);
        return top;
    }
    int slot = top + incl - pushes;
#pragma unroll
    for (int c = 7; c >= 0; c--) {
        if (c < pushes) ws.stack[slot++] = makeItem(keys[c]);
    }
    return top + total;
}

// Aggregate::findClosestPoint for up to 32 queries at once (one per lane with hasQuery). Every lane of
// the warp must call this. r2max: squared search radius. On return hit describes the lane's own query.
template <int DIM>
__device__ void coopClosestPoint(const BvhView& bvh, WarpShared& ws, bool hasQuery, const Vec<DIM>& x, float r2max,
                                 ClosestHit<DIM>& hit, int *status)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    ws.qx[lane] = x[0];
    ws.qy[lane] = x[1];
    ws.qz[lane] = DIM == 3 ? x[DIM - 1] : 0.0f;
    ws.best[lane] = packBest(r2max, 0xffffffffu);
    unsigned m = __ballot_sync(full, hasQuery);
    int top = __popc(m);
    if (hasQuery) ws.stack[__popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)lane, 0u);
    __syncwarp();

    while (top > 0) {
        int n = min(top, 32);
        bool have = lane < n;
        uint2 item = make_uint2(0u, 0u);
        if (have) item = ws.stack[top - 1 - lane];
        top -= n;
        __syncwarp();

        uint32_t keys[8];
#pragma unroll
        for (int c = 0; c < 8; c++) keys[c] = 0xffffffffu;
        int pushes = 0;
        NodeRegs node;
        int owner = item.x & 31u;
        if (have) {
            float r2 = coopRadius2(ws, owner);
            if (__uint_as_float(item.x & ~31u) <= r2) {
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                if (item.y & kLeafFlag) {
                    uint32_t first = item.y & 0x0fffffffu, count = (item.y >> 28) & 7u;
                    for (uint32_t i = 0; i < count; i++) {
                        Prim<DIM> prim;
                        prim.load(bvh, first + i);
                        Vec<DIM> pt;
                        float u, v;
                        float d = prim.closest(q, pt, u, v);
                        float d2 = d*d;
                        if (d2 <= r2) {
                            atomicMin(&ws.best[owner], packBest(d2, first + i));
                            r2 = d2;
                        }
                    }
                } else {
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
                        keys[c] = valid ? ((__float_as_uint(d2min) & ~7u) | (uint32_t)c) : 0xffffffffu;
                    }
                    if (shrink < r2) {
                        atomicMin(&ws.best[owner], packBest(shrink, 0xffffffffu));
                        r2 = shrink;
                    }
                    sort8(keys);
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        if (keys[c] != 0xffffffffu && keyValue(keys[c]) <= r2) pushes = c + 1;
                    }
                }
            }
        }
        top = coopPush(ws, top, lane, keys, pushes, [&](uint32_t key) {
            return make_uint2((key & ~31u) | (uint32_t)owner, childRef(node, childMeta(node, key & 7u)));
        }, status);
        __syncwarp();
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
    __syncwarp();
}

// ReflectanceBvh::computeSquaredStarRadius / findClosestSilhouettePoint for up to 32 queries at once.
// Every lane of the warp must call this; returns the squared star radius (<= r2max) of the lane's query.
template <int DIM>
__device__ float coopStarRadius2(const BvhView& bvh, WarpShared& ws, bool hasQuery, const Vec<DIM>& x, float r2max,
                                 bool flip, float precision, int *status)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    ws.qx[lane] = x[0];
    ws.qy[lane] = x[1];
    ws.qz[lane] = DIM == 3 ? x[DIM - 1] : 0.0f;
    ws.best[lane] = packBest(r2max, 0u);
    ws.flags[lane] = flip ? 1u : 0u;
    unsigned m = __ballot_sync(full, hasQuery);
    int top = __popc(m);
    if (hasQuery) ws.stack[__popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)lane, 0u);
    __syncwarp();

    while (top > 0) {
        int n = min(top, 32);
        bool have = lane < n;
        uint2 item = make_uint2(0u, 0u);
        if (have) item = ws.stack[top - 1 - lane];
        top -= n;
        __syncwarp();

        uint32_t keys[8];
#pragma unroll
        for (int c = 0; c < 8; c++) keys[c] = 0xffffffffu;
        int pushes = 0;
        uint32_t noSil = 0u;
        NodeRegs node;
        int owner = item.x & 31u;
        if (have) {
            float r2 = coopRadius2(ws, owner);
            // bit 5 of the key: "no silhouette inside this subtree" (reflectance bvh.h:474-476 sign trick)
            bool silhouetteTests = !(item.x & 32u);
            if (__uint_as_float(item.x & ~63u) <= r2) {
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                bool ownerFlip = ws.flags[owner] != 0u;
                if (item.y & kLeafFlag) {
                    uint32_t first = item.y & 0x0fffffffu, count = (item.y >> 28) & 7u;
                    float before = r2;
                    for (uint32_t i = 0; i < count; i++) primStarRadius2<DIM>(bvh, first + i, q, r2, ownerFlip, precision, silhouetteTests);
                    if (r2 < before) atomicMin(&ws.best[owner], packBest(r2, 0u));
                } else {
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
                        if (!starChildBounds<DIM>(bvh, item.y, c, q, lo, hi, d2min, d2max, lower, upper, hasSil)) continue;
                        shrink = fminf(shrink, upper);
                        if (!hasSil) noSil |= 1u << c;
                        keys[c] = (__float_as_uint(lower) & ~7u) | (uint32_t)c;
                    }
                    if (shrink < r2) {
                        atomicMin(&ws.best[owner], packBest(shrink, 0u));
                        r2 = shrink;
                    }
                    sort8(keys);
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        if (keys[c] != 0xffffffffu && keyValue(keys[c]) <= r2) pushes = c + 1;
                    }
                }
            }
        }
        top = coopPush(ws, top, lane, keys, pushes, [&](uint32_t key) {
            uint32_t slot = key & 7u;
            return make_uint2((key & ~63u) | (((noSil >> slot) & 1u) ? 32u : 0u) | (uint32_t)owner,
                              childRef(node, childMeta(node, slot)));
        }, status);
        __syncwarp();
    }
    float r2 = coopRadius2(ws, lane);
    __syncwarp();
    return hasQuery ? r2 : r2max;
}

} // namespace zb
