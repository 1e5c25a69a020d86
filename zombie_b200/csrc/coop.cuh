// Warp-cooperative BVH traversal for the two queries whose cost is heavy-tailed: closest point and
// star radius. Measured on the device code compiled for the host (tools/host_probe.cu, 1.3 M triangles):
// the median star-radius query of a walk touches ~5 nodes, the mean 136 nodes + 736 primitives, the 99th
// percentile ~2000 nodes + 9000 primitives, so with one thread per query a warp idles on its slowest lane
// (lane efficiency ~0.1).
//
// Two phases. (1) Every lane runs its own depth-first traversal (queries.cuh: closestPointDfs / starRadiusDfs,
// nearest child first) for a bounded number of stack pops. Most queries finish here with exactly the work of
// a sequential traversal, and the ones that do not have already found a close primitive, so their search
// radius is tight. (2) The unfinished traversals are pooled: the query state moves to shared memory and the
// warp keeps TWO work lists there,
//   * a stack of (owner query, node) items, and
//   * a queue of (owner query, primitive) items,
// refilled from the lanes' private stacks, and runs homogeneous rounds: a NODE round pops up to 32 nodes, every lane tests the 8 children of one node
// against its owner's query and the survivors are appended with warp prefix sums (internal children to the node
// stack, the primitives of leaf children to the primitive queue); a PRIMITIVE round pops 32 primitives, one per
// lane. Rounds never mix the two code paths, so a warp does not pay for both on every iteration, and a leaf
// with 4 primitives occupies 4 lanes for one round instead of one lane for four. All lanes stay busy until
// fewer than 32 items remain, whatever the skew between the 32 queries. Results are combined with shared-memory
// atomicMin on (squared radius, primitive), which makes them independent of the traversal order (ties on
// distance resolve to the smaller leaf-order index).
//
// Primitive rounds come in two stages (3D): a FILTER round runs a cheap conservative test on 32 primitives — can this
// triangle still lower the owner's radius? (plane distance and bounding sphere for the closest point; plane-distance
// bound of the reflectance radius and "no edge to test" for the star radius) — and compacts the survivors into a third
// list, kept at the end of the primitive queue; a FULL round runs the complete test on 32 survivors. In a warp an early
// exit is only worth what the slowest lane saves, so the rejection has to happen between rounds, not inside one.
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
#define zb_popc(x) __popc(x)
#endif

namespace zb {

#if defined(ZB_EMULATE_WARP)
extern long long g_coopCount[10]; // [0] closest node items, [1] closest prims, [2] star node items, [3] star prims (full test), [4..7] rounds, [8] star prims through the filter stage, [9] filter rounds
#define ZB_COOP_COUNT(i, n) __atomic_fetch_add(&g_coopCount[i], (long long)(n), __ATOMIC_RELAXED);
#else
#define ZB_COOP_COUNT(i, n)
#endif

// debugging hook of tools/gpu_star_debug.cu: reports what lowers a given query's radius in the pooled phase
#ifndef ZB_COOP_TRACE
#define ZB_COOP_TRACE(owner, what, value, id)
#endif

#ifndef ZB_NODE_STACK
#define ZB_NODE_STACK 512
#endif
#ifndef ZB_STAR_TWO_STAGE
#define ZB_STAR_TWO_STAGE 1 // 0: every pooled primitive round runs the full star test (A/B builds)
#endif
#ifndef ZB_STAR_DEFER_REDO
#define ZB_STAR_DEFER_REDO 1 // 0: the reference-ordered redo of a star primitive runs inside the full round that asked for it (A/B builds)
#endif
#ifndef ZB_PHASE1_BUDGET
#define ZB_PHASE1_BUDGET 8
#endif
constexpr int kNodeStack = ZB_NODE_STACK; // node items per warp; the top kNodeHeadroom are depth-first headroom
constexpr int kNodeHeadroom = kNodeStack/8;
constexpr int kPhase1Budget = ZB_PHASE1_BUDGET; // private stack pops per lane before the warp pools what is left
#ifndef ZB_PRIM_QUEUE
#define ZB_PRIM_QUEUE 704
#endif
constexpr int kPrimQueue = ZB_PRIM_QUEUE; // primitive items per warp; >= 31 + what one node round may append (coopNodeBudget) + 63 survivors of the cheap stage (kept at the end)
constexpr int kPrimSlack = 31 + 63;       // what the queue may hold when a node round or a refill starts: < 32 unfiltered items + the survivor list
constexpr int kRefillLowWater = kNodeStack*3/16;  // refill when the shared node stack is below this
// private stack entries (<= 4 primitives each) a lane hands over per refill: what the primitive queue can take, and no more
// node items than leave a full node round (32 nodes, 7 net entries each) its room below the headroom
constexpr int kRefillByPrims = (kPrimQueue - kPrimSlack)/128, kRefillByNodes = (kNodeStack - kNodeHeadroom - 7*32 - kRefillLowWater)/32;
constexpr int kRefillPerLane = kRefillByPrims < kRefillByNodes ? (kRefillByPrims < 8 ? kRefillByPrims : 8) : (kRefillByNodes < 8 ? kRefillByNodes : 8);
static_assert(kRefillPerLane >= 1 && kNodeStack >= kRefillLowWater + 32*kRefillPerLane, "work lists too small");
constexpr uint32_t kPrimIndexMask = 0x03ffffffu; // 26 bits: up to 67 M primitives per boundary
enum : int { kStatusCoopOverflow = 4 };

struct WarpShared {
    uint2 nodes[kNodeStack];            // x: lower-bound key (float bits, low 6 bits = noSil | owner lane); y: node index
    uint32_t prims[kPrimQueue];         // owner lane << 27 | noSil << 26 | primitive (leaf order)
    unsigned long long best[32];        // high word: current squared radius (float bits); low word: payload
    float qx[32], qy[32], qz[32];
    unsigned int flags[32];
    uint32_t redo[64];                  // star radius: primitives whose fast test asked for the reference-ordered redo (same item format)
    uint32_t redoRobin[64];             // ... for the redo of their reflectance part only
};

// four 128-thread CTAs per SM (the register budget of the walk kernel) must also fit in the 227 KB of shared memory,
// 1 KB of which the system reserves per CTA
static_assert(4*(4*sizeof(WarpShared) + 1024) <= 227*1024, "WarpShared too large for 4 CTAs per SM");

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
ZB_COOP int coopNodeBudget(int nodeTop, int primTop, int fullTop)
{
    int room = (kNodeStack - kNodeHeadroom - nodeTop)/7;
    int primRoom = (kPrimQueue - primTop - fullTop)/32; // a node has at most 8 leaf children of 4 primitives
    return min(min(nodeTop, 32), max(min(room, primRoom), 1));
}

// Appends the surviving children of this lane's node (keys[0..nKeep) ascending; low 3 bits = child slot) to the
// work lists: internal children far-to-near on the node stack (the nearest ends on top of the lane's group),
// the primitives of leaf children far-to-near on the primitive queue. tag = noSil bits (per slot) for the
// star-radius query, 0 for closest point. Every lane must call this (lanes without a node pass nKeep = 0).
ZB_COOP void coopPushChildren(WarpShared& ws, int& nodeTop, int& primTop, int lane, int owner,
                                                 const NodeRegs& node, const uint32_t keys[8], int nKeep, uint32_t noSil, int *status,
                                                 int primReserved = 0)
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
    // one scan for both counts (nodes in the low half-word, primitives in the high one: totals stay below 2^16)
    int incl = warpInclusiveSum(nNodes | (nPrims << 16), lane), total = zb_shfl(incl, 31);
    int inclN = incl & 0xffff, inclP = incl >> 16, totalN = total & 0xffff, totalP = total >> 16;
    if (nodeTop + totalN > kNodeStack || primTop + totalP + primReserved > kPrimQueue) { // warp-uniform; cannot happen for trees of depth <= 18
        if (lane == 0 && status) zb_status_or(status, kStatusCoopOverflow);
        return;
    }
    // the lane's group is filled from its END while the keys are walked nearest first: one rolled loop of nKeep
    // iterations (the keys rotate through their registers, so none is indexed with a run-time value) instead of eight
    // unrolled copies of the body
    int endN = nodeTop + inclN, endP = primTop + inclP;
    uint32_t k[8];
#pragma unroll
    for (int c = 0; c < 8; c++) k[c] = keys[c];
#pragma unroll 1
    for (int c = 0; c < nKeep; c++) {
        uint32_t key = k[0];
#pragma unroll
        for (int i = 0; i < 7; i++) k[i] = k[i + 1];
        uint32_t slot = key & 7u;
        uint32_t meta = childMeta(node, slot);
        uint32_t count = meta >> 5;
        uint32_t sil = (noSil >> slot) & 1u;
        if (count == 0u) {
            ws.nodes[--endN] = make_uint2((key & ~63u) | (sil << 5) | (uint32_t)owner, node.q[1].x + (meta & 31u) - 1u);
        } else {
            uint32_t first = node.q[1].y + (meta & 31u);
            uint32_t head = ((uint32_t)owner << 27) | (sil << 26);
            endP -= (int)count;
            for (uint32_t i = 0; i < count; i++) ws.prims[endP + i] = head | (first + count - 1u - i);
        }
    }
    nodeTop += totalN;
    primTop += totalP;
}

// Moves up to kRefillPerLane entries from the top of every lane's private traversal stack (queries.cuh format: x =
// key with flag bits in the low 3 bits, y = node index or kLeafFlag | count << 28 | first primitive) to the
// shared lists. starFlags: bit 2 of x is the "no silhouette" mark of the star-radius traversal. Every lane calls.
// Only called with primTop < 32 and nodeTop < kRefillLowWater, so the lists cannot overflow.
//
// STACK: where the lanes' private stacks live. LocalStack: the lane's own array (a query run by one function call);
// StridedStack: rows of a global-memory scratch, entry i of virtual thread t at p[i*stride + t] (the wavefront's star-radius
// query, whose private phase runs in a kernel of its own and hands the unfinished traversals over through HBM, wave.cuh).
struct LocalStack {
    const uint2 *p;
    ZB_COOP uint2 at(int i) const { return p[i]; }
};
struct StridedStack {
    const uint2 *p;
    size_t stride;
    ZB_COOP uint2 at(int i) const { return p[(size_t)i*stride]; }
};

template <class STACK>
ZB_COOP void coopRefill(WarpShared& ws, int& nodeTop, int& primTop, int lane, const STACK& stack, int& sp, bool starFlags)
{
    int take = min(sp, kRefillPerLane);
    int nNodes = 0, nPrims = 0;
#pragma unroll 1
    for (int k = 0; k < take; k++) {
        uint32_t y = stack.at(sp - take + k).y;
        if (y & kLeafFlag) nPrims += (int)((y >> 28) & 7u); else nNodes++;
    }
    int incl = warpInclusiveSum(nNodes | (nPrims << 16), lane), total = zb_shfl(incl, 31);
    int inclN = incl & 0xffff, inclP = incl >> 16, totalN = total & 0xffff, totalP = total >> 16;
    int slotN = nodeTop + inclN - nNodes, slotP = primTop + inclP - nPrims;
#pragma unroll 1
    for (int k = 0; k < take; k++) { // bottom to top: the lane's nearest entry ends on top of its group
        uint2 e = stack.at(sp - take + k);
        uint32_t sil = starFlags ? (e.x >> 2) & 1u : 0u;
        if (e.y & kLeafFlag) {
            uint32_t first = e.y & 0x0fffffffu, count = (e.y >> 28) & 7u;
            uint32_t head = ((uint32_t)lane << 27) | (sil << 26);
            for (uint32_t i = count; i > 0u; i--) ws.prims[slotP++] = head | (first + i - 1u);
        } else {
            ws.nodes[slotN++] = make_uint2((e.x & ~63u) | (sil << 5) | (uint32_t)lane, e.y);
        }
    }
    sp -= take;
    nodeTop += totalN;
    primTop += totalP;
}

#ifndef ZB_CLOSEST_CHILD_PARALLEL
#define ZB_CLOSEST_CHILD_PARALLEL 0 // 1: the distance query's node rounds run one child per lane as well (A/B builds)
#endif
#ifndef ZB_STAR_CHILD_PARALLEL
#define ZB_STAR_CHILD_PARALLEL 0 // 1: node rounds of the pooled star query run one CHILD per lane (coopNodeRoundByChild below); measured slower, see there
#endif

// MEASURED AND NOT SHIPPED (profiles/README.md, r2w): same instruction count as the node-per-lane round (every lane of a group
// decodes the node again, and the ranking costs what the sorted insertion did), 23.7 instead of 17.7 active lanes per
// instruction, but 57 % instead of 67 % busy issue slots (shuffle and shared-memory dependencies): 21.4 ms instead of 17.7 ms
// per round. Kept as a build option (ZB_STAR_CHILD_PARALLEL / ZB_CLOSEST_CHILD_PARALLEL = 1); the emulated-warp test covers it.
//
// Node round of the pooled star-radius query, one CHILD per lane. With one node per lane a round of n nodes keeps n lanes
// busy for eight child tests each, and a pool in its depth-first phase rarely has 32 nodes to give (ncu, one node per lane:
// 13 of 32 lanes active in the child test, 10 in the push, together half the kernel's instructions). Here the n popped nodes
// are processed four at a time: the eight lanes of a group test the eight child slots of one node, the node's radius bound
// is the group's minimum, and the survivors are ranked inside their group by seven shuffles (farther children lower on the
// lists, the nearest on top) — the same list contents as coopPushChildren produces, in a different order of the groups:
// passes run from the farthest popped node to the nearest, so the nearest node's nearest child ends on top.
// STAR = false is the same round for the distance query: the child test is the box distance alone.
template <int DIM, bool STAR>
ZB_COOP void coopNodeRoundByChild(const BvhView& bvh, WarpShared& ws, int& nodeTop, int& primTop, int fullTop, int lane,
                                  float precision, int *status)
{
    const int n = coopNodeBudget(nodeTop, primTop, fullTop);
    if (lane == 0) { if (STAR) { ZB_COOP_COUNT(2, n) ZB_COOP_COUNT(6, 1) } else { ZB_COOP_COUNT(0, n) ZB_COOP_COUNT(4, 1) } }
    // the popped items leave the list before anything is pushed: lane j keeps item j (0 = top of the list)
    uint2 mine = make_uint2(0u, 0u);
    if (lane < n) mine = ws.nodes[nodeTop - 1 - lane];
    nodeTop -= n;
    zb_syncwarp();
    const int g = lane >> 3, c = lane & 7;
    const uint32_t kNone = 0xffffffffu;
#pragma unroll 1
    for (int p = (n + 3)/4 - 1; p >= 0; p--) {
        const int j = 4*p + g;
        uint2 item;
        item.x = (uint32_t)zb_shfl((int)mine.x, j & 31);
        item.y = (uint32_t)zb_shfl((int)mine.y, j & 31);
        bool valid = j < n;
        const int owner = (int)(item.x & 31u);
        float r2 = 0.0f, upper = kMaxFloat;
        uint32_t packed = kNone; // lower bound (rounded down to 64 ulps) | primitive count << 3 | child slot; kNone: nothing to push
        uint32_t ref = 0u, sil = 0u;
        if (valid) {
            r2 = coopRadius2(ws, owner);
            valid = zb_uint_as_float(item.x & ~63u) <= r2;
        }
        if (valid) {
            NodeRegs node;
            loadNode(bvh.nodes, item.y, node);
            const uint32_t meta = childMeta(node, c);
            if (meta != 0u) {
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                NodeFrame<DIM> f = nodeFrame<DIM>(node);
                float lo[DIM], hi[DIM], d2min, d2max;
                childBox<DIM>(node, f, c, lo, hi);
                boxDistance2<DIM>(lo, hi, q, d2min, d2max);
                if (!STAR) {
                    // closest point: every child's farthest corner bounds the radius (bvh.inl:1339-1343), its nearest is the key
                    upper = d2max;
                    const uint32_t count = meta >> 5;
                    packed = (zb_float_as_uint(d2min) & ~63u) | (count << 3) | (uint32_t)c;
                    ref = count ? node.q[1].y + (meta & 31u) : node.q[1].x + (meta & 31u) - 1u;
                } else if (!(d2min > r2)) {
                    const float4 *cone = bvh.cones + (size_t)item.y*16 + 2*c;
                    float4 c0 = zb_ldg(cone), c1 = zb_ldg(cone + 1);
                    float lower, up;
                    bool hasSil;
                    if (starChildBoundsFast<DIM>(c0, c1, q, lo, hi, d2min, d2max, lower, up, hasSil, precision)) {
                        upper = up;
                        const bool parentNoSil = (item.x >> 5) & 1u; // no silhouette above => none below
                        sil = (!hasSil || parentNoSil) ? 1u : 0u;
                        const uint32_t count = meta >> 5;
                        packed = (zb_float_as_uint(lower) & ~63u) | (count << 3) | (uint32_t)c;
                        ref = count ? node.q[1].y + (meta & 31u) : node.q[1].x + (meta & 31u) - 1u;
                    }
                }
            }
        }
        // the node's upper bound of the radius: the minimum over its children (non-negative floats order like their bits)
        uint32_t shrinkBits = zb_float_as_uint(upper);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) shrinkBits = zb_umin(shrinkBits, (uint32_t)zb_shfl((int)shrinkBits, lane ^ o));
        const float shrink = zb_uint_as_float(shrinkBits);
        if (valid && shrink < r2) {
            if (c == 0) zb_atomic_min64(&ws.best[owner], packBest(shrink, STAR ? 0u : 0xffffffffu));
            r2 = shrink;
        }
        if (packed != kNone && !(zb_uint_as_float(packed & ~63u) <= r2)) packed = kNone;

        // rank inside the group (farther survivors below) and the group's totals, from the seven other lanes of the group
        int belowN = 0, belowP = 0, groupN = 0, groupP = 0;
        if (packed != kNone) { uint32_t cnt = (packed >> 3) & 7u; groupN = cnt == 0u ? 1 : 0; groupP = (int)cnt; }
#pragma unroll
        for (int d = 1; d < 8; d++) {
            uint32_t o = (uint32_t)zb_shfl((int)packed, (lane & ~7) | ((c + d) & 7));
            if (o != kNone) {
                uint32_t cnt = (o >> 3) & 7u;
                int isNode = cnt == 0u ? 1 : 0;
                groupN += isNode; groupP += (int)cnt;
                if (packed != kNone && o > packed) { belowN += isNode; belowP += (int)cnt; }
            }
        }
        // groups are stacked 3 (bottom) ... 0 (top): group 0 holds the nearest of the pass's four nodes
        const uint32_t tot = (uint32_t)groupN | ((uint32_t)groupP << 16);
        const uint32_t t0 = (uint32_t)zb_shfl((int)tot, 0), t1 = (uint32_t)zb_shfl((int)tot, 8), t2 = (uint32_t)zb_shfl((int)tot, 16),
                       t3 = (uint32_t)zb_shfl((int)tot, 24);
        const uint32_t below = (g < 3 ? t3 : 0u) + (g < 2 ? t2 : 0u) + (g < 1 ? t1 : 0u), all = t0 + t1 + t2 + t3;
        const int totalN = (int)(all & 0xffffu), totalP = (int)(all >> 16);
        if (nodeTop + totalN > kNodeStack || primTop + totalP + fullTop > kPrimQueue) { // warp-uniform; cannot happen for trees of depth <= 18
            if (lane == 0 && status) zb_status_or(status, kStatusCoopOverflow);
            continue;
        }
        if (packed != kNone) {
            const uint32_t cnt = (packed >> 3) & 7u;
            if (cnt == 0u) {
                ws.nodes[nodeTop + (int)(below & 0xffffu) + belowN] = make_uint2((packed & ~63u) | (sil << 5) | (uint32_t)owner, ref);
            } else {
                const uint32_t head = ((uint32_t)owner << 27) | (sil << 26);
                const int base = primTop + (int)(below >> 16) + belowP;
                for (uint32_t i = 0; i < cnt; i++) ws.prims[base + (int)i] = head | (ref + cnt - 1u - i);
            }
        }
        nodeTop += totalN;
        primTop += totalP;
    }
    zb_syncwarp();
}

// Aggregate::findClosestPoint for up to 32 queries at once (one per lane with hasQuery). Every lane of
// the warp must call this. r2max: squared search radius. On return hit describes the lane's own query.
template <int DIM>
ZB_COOP_NOINLINE void coopClosestPoint(const BvhView& bvh, WarpShared& ws, bool hasQuery, const Vec<DIM>& x, float r2max,
                                 ClosestHit<DIM>& hit, int *status)
{
    const int lane = zb_lane();
    // ---- phase 1: private depth-first traversal with a budget
    uint2 stack[kStackSize];
    int sp = 0;
    float r2own = r2max;
    hit.d = kMaxFloat;
    hit.prim = -1;
    hit.p = x;
    if (hasQuery) {
        stack[sp++] = make_uint2(0u, 0u);
        closestPointDfs<DIM>(bvh, x, r2own, hit, stack, sp, kPhase1Budget, status);
    }
    if (zb_ballot(sp > 0) == 0u) return;

    // ---- phase 2: pool the unfinished traversals
    zb_syncwarp();
    ws.qx[lane] = x[0];
    ws.qy[lane] = x[1];
    ws.qz[lane] = DIM == 3 ? x[DIM - 1] : 0.0f;
    // the phase-1 primitive stays the incumbent only if it lies within the (possibly box-shrunk) radius
    ws.best[lane] = packBest(r2own, hit.prim >= 0 && hit.d*hit.d <= r2own ? (uint32_t)hit.prim : 0xffffffffu);
    int nodeTop = 0, primTop = 0;
    // two-stage primitive rounds as in coopStarRadius2 below: a filter round keeps the triangles whose plane is within
    // the search radius (primClosestNeedsFullTest), full rounds run the closest-point test on 32 survivors
    constexpr bool kTwoStage = DIM == 3 && ZB_CLOSEST_PLANE_FILTER && ZB_STAR_TWO_STAGE;
    int fullTop = 0;
    zb_syncwarp();

    while (true) {
        if (primTop < 32 && nodeTop < kRefillLowWater && zb_ballot(sp > 0) != 0u) {
            coopRefill(ws, nodeTop, primTop, lane, LocalStack{stack}, sp, false);
            zb_syncwarp();
        }
        if (nodeTop == 0 && primTop == 0 && fullTop == 0) break;
        const bool fullRound = kTwoStage && (fullTop >= 32 || (nodeTop == 0 && primTop == 0));
        if (fullRound || primTop >= 32 || nodeTop == 0) {
            // ---- primitive round: one primitive per lane
            int n = fullRound ? min(fullTop, 32) : min(primTop, 32);
            uint32_t item = 0u;
            bool survivor = false;
            if (lane < n) {
                item = fullRound ? ws.prims[kPrimQueue - fullTop + lane] : ws.prims[primTop - 1 - lane];
                int owner = (int)(item >> 27);
                uint32_t idx = item & kPrimIndexMask;
                float r2 = coopRadius2(ws, owner);
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                if (kTwoStage && !fullRound) {
                    survivor = primClosestNeedsFullTest<DIM>(bvh.prims, idx, q, r2);
                } else {
                    float d = primClosestCall<DIM>(bvh.prims, idx, q).w;
                    float d2 = d*d;
                    if (d2 <= r2) zb_atomic_min64(&ws.best[owner], packBest(d2, idx));
                }
            }
            if (fullRound) {
                if (lane == 0) { ZB_COOP_COUNT(1, n) ZB_COOP_COUNT(5, 1) }
                fullTop -= n;
            } else {
                primTop -= n;
                if (kTwoStage) {
                    unsigned int m = zb_ballot(survivor);
                    if (survivor) ws.prims[kPrimQueue - 1 - (fullTop + zb_popc(m & ((1u << lane) - 1u)))] = item;
                    fullTop += zb_popc(m);
                } else if (lane == 0) { ZB_COOP_COUNT(1, n) ZB_COOP_COUNT(5, 1) }
            }
            zb_syncwarp();
            continue;
        }

        // ---- node round
        if (ZB_CLOSEST_CHILD_PARALLEL) {
            coopNodeRoundByChild<DIM, false>(bvh, ws, nodeTop, primTop, fullTop, lane, 0.0f, status);
            continue;
        }
        // one node per lane
        int n = coopNodeBudget(nodeTop, primTop, fullTop);
        if (lane == 0) { ZB_COOP_COUNT(0, n) ZB_COOP_COUNT(4, 1) }
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
                ZB_CHILD_LOOP
                for (int c = 0; c < 8; c++) {
                    if (childMeta(node, c) == 0u) continue;
                    float lo[DIM], hi[DIM], d2min, d2max;
                    childBox<DIM>(node, f, c, lo, hi);
                    boxDistance2<DIM>(lo, hi, q, d2min, d2max);
                    shrink = fminf(shrink, d2max);
                    // keys are rounded DOWN to a multiple of 64 ulps: still a lower bound of the child's distance
                    insertKey(keys, (zb_float_as_uint(d2min) & ~63u) | (uint32_t)c);
                }
                if (shrink < r2) {
                    // the search radius also shrinks by each child's farthest-corner distance (bvh.inl:1339-1343)
                    zb_atomic_min64(&ws.best[owner], packBest(shrink, 0xffffffffu));
                    r2 = shrink;
                }
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (keys[c] != 0xffffffffu && zb_uint_as_float(keys[c] & ~63u) <= r2) nKeep = c + 1;
                }
            }
        }
        coopPushChildren(ws, nodeTop, primTop, lane, owner, node, keys, nKeep, 0u, status, fullTop);
        zb_syncwarp();
    }

    unsigned long long b = ws.best[lane];
    uint32_t prim = (uint32_t)b;
    if (hasQuery && prim != 0xffffffffu && (int)prim != hit.prim) {
        float4 cp = primClosestCall<DIM>(bvh.prims, prim, x);
        hit.d = cp.w;
        hit.p[0] = cp.x; hit.p[1] = cp.y;
        if (DIM == 3) hit.p[DIM - 1] = cp.z;
        hit.prim = (int)prim;
    }
    zb_syncwarp();
}

// ------------------------------------------------------------------------------------------
// star radius
// ------------------------------------------------------------------------------------------

// Phase 2 of the star-radius query: pools the traversals phase 1 left unfinished. Every lane of the warp calls; hasQuery:
// the lane owns a query (x, flip) whose radius so far is r2own and whose unfinished traversal is stack[0..sp). Returns the
// squared star radius of the lane's query (r2max for lanes without one).
template <int DIM, class STACK>
ZB_COOP float coopStarPooled(const BvhView& bvh, WarpShared& ws, bool hasQuery, const Vec<DIM>& x, float r2own, float r2max,
                             bool flip, float precision, int *status, const STACK& stack, int sp)
{
    const int lane = zb_lane();
    zb_syncwarp();
    ws.qx[lane] = x[0];
    ws.qy[lane] = x[1];
    ws.qz[lane] = DIM == 3 ? x[DIM - 1] : 0.0f;
    ws.best[lane] = packBest(r2own, 0u);
    ZB_COOP_TRACE(lane, "phase-1 radius", r2own, sp)
    ws.flags[lane] = flip ? 1u : 0u;
    int nodeTop = 0, primTop = 0;
    // Two-stage primitive rounds (3D): a FILTER round pops 32 primitives and runs the cheap first stage of the test
    // (primStarNeedsFullTest); the survivors are compacted into a second list, kept at the END of ws.prims, and a FULL
    // round runs the whole test on 32 of them. fullTop counts that list (entries ws.prims[kPrimQueue - 1 - k]).
    constexpr bool kTwoStage = DIM == 3 && ZB_STAR_TWO_STAGE;
    int fullTop = 0;
    // Third stage (3D): a primitive whose fast test lands within rounding distance of a discrete decision is redone with
    // the reference-ordered arithmetic (primStarRadius2Redo, ~1,900 instructions). A third of the pooled full tests of a
    // C3 solve ask for it, in more than half of the full rounds at least one lane does (profiles/r2_wave_split_ncu.txt) —
    // and a warp pays for the redo in full whenever one lane takes it. So the full round only marks such primitives; they
    // collect in ws.redo and are redone 32 at a time. Radii combine by minimum, so the order of the tests does not matter.
    constexpr bool kDeferRedo = kTwoStage && ZB_STAR_DEFER_REDO;
    int redoTop = 0, robinTop = 0; // whole-primitive redos (rare: a silhouette decision on its threshold) / reflectance-only redos
    zb_syncwarp();

    while (true) {
        if (primTop < 32 && nodeTop < kRefillLowWater && zb_ballot(sp > 0) != 0u) {
            coopRefill(ws, nodeTop, primTop, lane, stack, sp, true);
            zb_syncwarp();
        }
        const bool drained = nodeTop == 0 && primTop == 0 && fullTop == 0;
        if (drained && redoTop == 0 && robinTop == 0) break;
        if (kDeferRedo && (robinTop >= 32 || (drained && robinTop > 0))) {
            // ---- redo round, reflectance part only
            int n = min(robinTop, 32);
            if (lane < n) {
                uint32_t item = ws.redoRobin[robinTop - n + lane];
                int owner = (int)(item >> 27);
                float before = coopRadius2(ws, owner);
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                float r2 = primStarRobinRedo<DIM>(bvh.prims, item & kPrimIndexMask, q, before);
                if (r2 < before) zb_atomic_min64(&ws.best[owner], packBest(r2, 0u));
            }
            robinTop -= n;
            zb_syncwarp();
            continue;
        }
        if (kDeferRedo && (redoTop >= 32 || (drained && redoTop > 0))) {
            // ---- redo round
            int n = min(redoTop, 32);
            if (lane < n) {
                uint32_t item = ws.redo[redoTop - n + lane];
                int owner = (int)(item >> 27);
                bool silhouetteTests = !((item >> 26) & 1u);
                float before = coopRadius2(ws, owner);
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                float r2 = primStarRadius2Redo<DIM>(bvh.prims, bvh.adjacency, bvh.snchSign, item & kPrimIndexMask, q, before,
                                                    ws.flags[owner] != 0u, precision, silhouetteTests);
                if (r2 < before) zb_atomic_min64(&ws.best[owner], packBest(r2, 0u));
            }
            redoTop -= n;
            zb_syncwarp();
            continue;
        }
        if (kTwoStage && (fullTop >= 32 || (nodeTop == 0 && primTop == 0))) {
            // ---- full round on the survivors
            int n = min(fullTop, 32);
            if (lane == 0) { ZB_COOP_COUNT(3, n) ZB_COOP_COUNT(7, 1) }
            int redoKind = 0;
            uint32_t item = 0u;
            if (lane < n) {
                item = ws.prims[kPrimQueue - fullTop + lane];
                int owner = (int)(item >> 27);
                bool silhouetteTests = !((item >> 26) & 1u);
                float r2 = coopRadius2(ws, owner), before = r2;
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                float2 res = primStarRadius2Call<DIM>(bvh.prims, bvh.adjacency, bvh.snchSign, item & kPrimIndexMask, q, r2, ws.flags[owner] != 0u,
                                                      precision, silhouetteTests, kDeferRedo);
                r2 = res.x;
                redoKind = kDeferRedo ? (int)res.y : 0;
                if (r2 < before) { ZB_COOP_TRACE(owner, silhouetteTests ? "full-round primitive (silhouette tests on)" : "full-round primitive (no silhouette tests)", r2, item & kPrimIndexMask) ZB_COOP_TRACE(owner, "   ... its incoming radius", before, lane) ZB_COOP_TRACE(owner, "   ... query x", q[0], item) ZB_COOP_TRACE(owner, "   ... query y", q[1], ws.flags[owner]) ZB_COOP_TRACE(owner, "   ... query z", q[DIM - 1], fullTop) zb_atomic_min64(&ws.best[owner], packBest(r2, 0u)); }
            }
            fullTop -= n;
            if (kDeferRedo) {
                unsigned int m = zb_ballot(redoKind == 1);
                if (redoKind == 1) ws.redo[redoTop + zb_popc(m & ((1u << lane) - 1u))] = item;
                redoTop += zb_popc(m); // < 32 before the round (else a redo round would have run), <= 63 now
                m = zb_ballot(redoKind == 2);
                if (redoKind == 2) ws.redoRobin[robinTop + zb_popc(m & ((1u << lane) - 1u))] = item;
                robinTop += zb_popc(m);
            }
            zb_syncwarp();
            continue;
        }
        if (primTop >= 32 || nodeTop == 0) {
            // ---- primitive round (two-stage: filter round)
            int n = min(primTop, 32);
            uint32_t item = 0u;
            bool survivor = false;
            if (lane < n) {
                item = ws.prims[primTop - 1 - lane];
                int owner = (int)(item >> 27);
                // bit 26: "no silhouette inside this subtree" (reflectance bvh.h:474-476 sign trick)
                bool silhouetteTests = !((item >> 26) & 1u);
                float r2 = coopRadius2(ws, owner), before = r2;
                Vec<DIM> q;
                q[0] = ws.qx[owner]; q[1] = ws.qy[owner];
                if (DIM == 3) q[DIM - 1] = ws.qz[owner];
                if (kTwoStage) {
                    survivor = primStarNeedsFullTest<DIM>(bvh, item & kPrimIndexMask, q, r2, silhouetteTests);
                } else {
                    r2 = primStarRadius2Call<DIM>(bvh.prims, bvh.adjacency, bvh.snchSign, item & kPrimIndexMask, q, r2,
                                                  ws.flags[owner] != 0u, precision, silhouetteTests).x;
                    if (r2 < before) zb_atomic_min64(&ws.best[owner], packBest(r2, 0u));
                }
            }
            primTop -= n;
            if (kTwoStage) {
                if (lane == 0) { ZB_COOP_COUNT(8, n) ZB_COOP_COUNT(9, 1) }
                unsigned int m = zb_ballot(survivor);
                if (survivor) ws.prims[kPrimQueue - 1 - (fullTop + zb_popc(m & ((1u << lane) - 1u)))] = item;
                fullTop += zb_popc(m);
            } else if (lane == 0) { ZB_COOP_COUNT(3, n) ZB_COOP_COUNT(7, 1) }
            zb_syncwarp();
            continue;
        }

        // ---- node round
        if (ZB_STAR_CHILD_PARALLEL) {
            coopNodeRoundByChild<DIM, true>(bvh, ws, nodeTop, primTop, fullTop, lane, precision, status);
            continue;
        }
        int n = coopNodeBudget(nodeTop, primTop, fullTop);
        if (lane == 0) { ZB_COOP_COUNT(2, n) ZB_COOP_COUNT(6, 1) }
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
                ConePrefetch cones;
                cones.start(bvh, item.y);
                ZB_CHILD_LOOP
                for (int c = 0; c < 8; c++) {
                    float4 c0, c1;
                    cones.next(c, c0, c1);
                    uint32_t meta = childMeta(node, c);
                    if (meta == 0u) continue;
                    float lo[DIM], hi[DIM], d2min, d2max;
                    childBox<DIM>(node, f, c, lo, hi);
                    boxDistance2<DIM>(lo, hi, q, d2min, d2max);
                    if (d2min > r2) continue;
                    float lower, upper;
                    bool hasSil;
                    if (!starChildBoundsFast<DIM>(c0, c1, q, lo, hi, d2min, d2max, lower, upper, hasSil, precision)) continue;
                    shrink = fminf(shrink, upper);
                    if (!hasSil || parentNoSil) noSil |= 1u << c;
                    insertKey(keys, (zb_float_as_uint(lower) & ~63u) | (uint32_t)c);
                }
                if (shrink < r2) {
                    ZB_COOP_TRACE(owner, "node upper bound", shrink, item.y)
                    zb_atomic_min64(&ws.best[owner], packBest(shrink, 0u));
                    r2 = shrink;
                }
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (keys[c] != 0xffffffffu && zb_uint_as_float(keys[c] & ~63u) <= r2) nKeep = c + 1;
                }
            }
        }
        coopPushChildren(ws, nodeTop, primTop, lane, owner, node, keys, nKeep, noSil, status, fullTop);
        zb_syncwarp();
    }
    float r2 = coopRadius2(ws, lane);
    zb_syncwarp();
    return hasQuery ? r2 : r2max;
}

// ReflectanceBvh::computeSquaredStarRadius / findClosestSilhouettePoint for up to 32 queries at once, both phases in one
// call. Every lane of the warp must call this; returns the squared star radius (<= r2max) of the lane's query.
template <int DIM>
ZB_COOP_NOINLINE float coopStarRadius2(const BvhView& bvh, WarpShared& ws, bool hasQuery, const Vec<DIM>& x, float r2max,
                                 bool flip, float precision, int *status)
{
    // ---- phase 1: private depth-first traversal with a budget
    uint2 stack[kStackSize];
    int sp = 0;
    float r2own = r2max;
    if (hasQuery) {
        stack[sp++] = make_uint2(0u, 0u);
        starRadiusDfs<DIM, true>(bvh, x, r2own, flip, precision, stack, sp, kPhase1Budget, status);
    }
    if (zb_ballot(sp > 0) == 0u) return hasQuery ? r2own : r2max;
    // ---- phase 2: pool the unfinished traversals
    return coopStarPooled<DIM>(bvh, ws, hasQuery, x, r2own, r2max, flip, precision, status, LocalStack{stack}, sp);
}

} // namespace zb
