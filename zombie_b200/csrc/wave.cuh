// Wavefront scheduling of the walks (the shipped scheduler).
//
// Round 1 ran every walk inside ONE persistent kernel (walkKernel below in walk.cuh's history): 128 registers, 54 KB of
// shared memory per CTA, 250 KB of SASS. Its ncu capture (profiles/r2_walkKernel_persistent_ncu.txt) shows 70 % of all
// warp-stall samples in `no_instruction`: sixteen warps per SM, each in a different phase of the step (distance query,
// star-radius query, ray casts, Green's-function math), thrash the 32 KB instruction cache, and the issue slots are
// 14 % busy. The fix is structural: the walk state lives in HBM (120 bytes per walk, SoA) and a ROUND of the solve is
// a short sequence of small, homogeneous kernels over all live walks,
//
//   waveClosestKernel   distance to the absorbing boundary at the walk's position (warp-cooperative traversal) for the
//                       walks that need it, then the scalar prologue of the step: inside the epsilon shell? does the
//                       step need a star-radius query?
//   waveStarKernel      star radius (warp-cooperative traversal) for the walks that need one
//   waveStepKernel      the rest of the step, one thread per walk: ray cast, boundary sample + occlusion ray, source
//                       sample, throughput, weight window; walks that ended write their record and the slot takes the
//                       next task from the global queue (compaction of terminated walks happens here, per slot)
//   radix sort          (cub) of the live slots by the Morton code of the walk's new position: the 32 queries of a warp
//                       of the next round are neighbours in space, so they traverse the same BVH nodes (L1 hits) and
//                       take similar numbers of steps through it (lane efficiency); slots without a walk sort to the
//                       end, which is also how the tail of a solve shrinks its grids
//
// Each kernel's hot loop fits the instruction cache, needs half the registers of the fused kernel, and every warp of
// a launch runs the same code. State traffic is ~250 bytes per walk step against the ~10^4..10^5 instructions a step
// costs: noise. Results do not depend on the schedule: a walk's random numbers come from its own pcg32 stream and its
// record goes to a slot indexed by the task.
#pragma once

#include "walk.cuh"

#ifndef ZB_STAR_MIN_BLOCKS
#define ZB_STAR_MIN_BLOCKS 6 // resident 128-thread CTAs per SM the star-radius kernel is compiled for (register cap 85; measured r2m/r2n: 6 > 5, 7, 8)
#endif
#ifndef ZB_STAR_PRIVATE_MIN_BLOCKS
#define ZB_STAR_PRIVATE_MIN_BLOCKS 8 // the private phase of the star-radius query needs no shared memory: registers alone bound its occupancy
#endif
#ifndef ZB_CLOSEST_MIN_BLOCKS
#define ZB_CLOSEST_MIN_BLOCKS 6
#endif

namespace zb {

enum : uint32_t {
    kWfActive = 1u, kWfNeedClosest = 2u, kWfKeepDistA = 4u, kWfFirstStep = 8u, kWfFlip = 16u, kWfOuterFlip = 32u,
    kWfSuccess = 64u, kWfGradient = 128u, kWfOnRefl = 256u, kWfYukawa = 512u, kWfNeedStar = 1024u, kWfReached = 2048u,
    kWfFinalProjection = 4096u, // Walk on Spheres ran out of steps: finish with kMaxLength once the projection of the last position is known
    kWfProjValid = 8192u,       // projStar.xyz is a point of the absorbing boundary (seeds the next distance query)
    kWfAligned = 16384u         // double-sided solves: the walk stands on the normal-aligned side of the absorbing boundary (its closest primitive)
};

struct WaveSlots {
    float4 *ptDist;       // position | distance to the absorbing boundary
    float4 *projStar;     // closest absorbing-boundary point | star radius of the coming step
    float4 *normalPrev;   // boundary normal at the position | length of the previous step
    float4 *dirThr;       // direction of the previous step | throughput
    float4 *acc;          // reflecting-boundary terms, source terms (current walk state) | task total | first sphere radius
    ulonglong2 *rng;      // pcg32 state, increment
    int4 *meta;           // sample point, walk length, task walk length, flags
    int4 *meta2;          // task (low, high word), finished walk states, queued splits
    void *queue;          // WalkState<DIM>[kSplitStack] per slot; null when the settings rule splitting out
    uint32_t *keys;       // Morton key per POSITION of the current order (written by waveStepKernel, sorted by cub)
    const uint32_t *order; // position -> slot
    unsigned int *live;   // walks alive after the current round; live[1], live[2]: warp tickets of the round's two query kernels
    // hand-over of the star-radius query from its private phase to its pooled phase (waveStarPrivateKernel -> waveStarPooledKernel),
    // indexed by VIRTUAL THREAD t = 32*virtual warp + lane: the unfinished traversal's stack entries (entry i at
    // starStack[i*starStride + t]), its radius so far and stack size, and per virtual warp the lanes that have one
    uint2 *starStack;
    uint2 *starPending;   // x: squared radius after the private phase (float bits); y: stack entries left
    uint32_t *starWarpMask;
    size_t starStride;
    unsigned long long *taskCounter; // tasks this group has taken; the group owns the tasks first + k*stride
    int taskFirst, taskStride;
    int nSlots;
    int debugEscapes;     // ZB_WAVE_DEBUG=1: report walks that leave the bounding domain (device printf)
    int sortMode;         // 0 fixed pseudo-random rank of the task (default), 1 Morton code of the position, 2 descending distance to the absorbing boundary
    float sortMin[3], sortScale[3];
};

ZB_DEV uint32_t spreadBits3(uint32_t v) // 10 bits -> every third bit
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
ZB_DEV uint32_t spreadBits2(uint32_t v) // 15 bits -> every second bit
{
    v &= 0x7fffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

template <int DIM>
ZB_DEV uint32_t mortonKey(const WaveSlots& w, const Vec<DIM>& p)
{
    uint32_t q[DIM];
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        float t = (p[k] - w.sortMin[k])*w.sortScale[k];
        const float top = DIM == 3 ? 1023.0f : 32767.0f;
        q[k] = (uint32_t)fminf(fmaxf(t, 0.0f), top);
    }
    if (DIM == 3) return spreadBits3(q[0]) | (spreadBits3(q[1]) << 1) | (spreadBits3(q[DIM - 1]) << 2);
    return spreadBits2(q[0]) | (spreadBits2(q[1]) << 1);
}

template <int DIM> ZB_DEV Vec<DIM> vec3of(const float4& v) { Vec<DIM> r; r[0] = v.x; r[1] = v.y; if (DIM == 3) r[DIM - 1] = v.z; return r; }
template <int DIM> ZB_DEV float4 pack4(const Vec<DIM>& v, float w) { return make_float4(v[0], v[1], DIM == 3 ? v[DIM - 1] : 0.0f, w); }


// Thread -> position of the slot order for the two query kernels: STRIPED, warp w of W takes the positions w, w + W, w + 2W, ...
// Live walks sit at the front of the order, so while every slot holds a walk each warp gets 32 of them, and when a solve
// runs out of tasks the remaining walks spread over all warps instead of filling a few: a warp with one or two queries
// brings all 32 lanes to their traversal (coop.cuh pools the work), which is what the tail of a solve needs — compacting
// the survivors into full warps would leave most SMs idle while a few warps grind through 32 heavy queries each.
ZB_DEV int stripedPosition(int nWarps)
{
#if defined(__CUDA_ARCH__)
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    return (i & 31)*nWarps + (i >> 5);
#else
    return 0;
#endif
}

// The two query kernels are PERSISTENT: the grid holds as many CTAs as the machine keeps resident, and every warp draws
// virtual warps (32 striped positions of the slot order each) from a ticket counter until none is left. A launch with one
// CTA per 128 positions runs in waves — 3.5 of them at 262,144 slots — and the cost of a query is heavy-tailed, so the last
// wave keeps a few SMs busy while the rest idle (ncu: 16.7 % achieved against 25 % theoretical occupancy); with tickets a
// warp that drew light queries simply draws again.
ZB_DEV int nextVirtualWarp(unsigned int *ticket, int nWarps)
{
#if defined(__CUDA_ARCH__)
    unsigned int v = 0u;
    if ((threadIdx.x & 31) == 0) v = atomicAdd(ticket, 1u);
    v = __shfl_sync(0xffffffffu, v, 0);
    return v < (unsigned int)nWarps ? (int)v : -1;
#else
    return -1;
#endif
}

// ------------------------------------------------------------------------------------------
// distance query + scalar prologue of the step
// ------------------------------------------------------------------------------------------
template <int DIM, bool WOST>
ZB_DEV void waveClosestWarp(const SolveArgs& a, const WaveSlots& w, WarpShared& ws, int nBound, int i)
{
#if defined(__CUDA_ARCH__)
    int slot = -1;
    int4 m = make_int4(0, 0, 0, 0);
    if (i < nBound) { slot = (int)w.order[i]; m = w.meta[slot]; }
    uint32_t flags = (uint32_t)m.w;
    const bool active = (flags & kWfActive) != 0u;
    if (!__any_sync(0xffffffffu, active)) return;

    // the part of the lane the prologue reads
    Lane<DIM> L;
    float4 pd = active ? w.ptDist[slot] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    L.st.pt = vec3of<DIM>(pd);
    L.distA = pd.w;
    L.st.walkLength = m.y;
    L.st.onRefl = (flags & kWfOnRefl) != 0u;
    L.firstStep = (flags & kWfFirstStep) != 0u;
    L.keepDistA = (flags & kWfKeepDistA) != 0u;
    L.flip = (flags & kWfFlip) != 0u;
    L.projValid = (flags & kWfProjValid) != 0u;

    const bool qc = active && (flags & kWfNeedClosest) != 0u;
    if (__any_sync(0xffffffffu, qc)) {
        float4 ps = qc ? w.projStar[slot] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        L.proj = vec3of<DIM>(ps);
        Vec<DIM> proj = L.st.pt;
        bool found = false, alignedSide = false;
        if (a.hasAbsorbing) {
            ClosestHit<DIM> hit;
            coopClosestPoint<DIM>(a.absorbing, ws, qc, L.st.pt, qc ? closestSearchRadius2<DIM, WOST>(a, L) : 0.0f, hit, a.status);
            if (qc) {
                if (!L.keepDistA) L.distA = hit.d;
                proj = hit.p;
                found = hit.prim >= 0;
                alignedSide = onNormalAlignedSide<DIM>(a, L.st.pt, hit.p, hit.prim);
            }
        } else if (qc) {
            // no absorbing boundary: default of GeometricQueries::populate (geometric_queries.h:213-218), as distToAbsorbing
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < DIM; k++) { float t = fminf(a.domainMin[k] - L.st.pt[k], L.st.pt[k] - a.domainMax[k]); s += t*t; }
            if (!L.keepDistA) L.distA = sqrtf(s);
        }
        if (qc) {
            flags = (flags & ~(kWfNeedClosest | kWfKeepDistA | kWfProjValid | kWfAligned)) | (found ? kWfProjValid : 0u) | (alignedSide ? kWfAligned : 0u);
            pd.w = L.distA;
            w.ptDist[slot] = pd;
            w.projStar[slot] = pack4<DIM>(proj, ps.w);
        }
    }
    if (!active) return;

    // loop condition of walk() (walk_on_spheres.h:237, walk_on_stars.h:357)
    if (flags & kWfFinalProjection) {
        // the walk already ended: waveStepKernel only waited for the projection
    } else if (!(L.distA > a.ws.epsilonShellForAbsorbingBoundary)) {
        flags |= kWfReached;
    } else if (WOST) {
        // stepStarPrologue reads the normal / previous direction only for double-sided walks on the boundary
        float4 np = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const bool twoSided = a.ws.solveDoubleSided && L.st.onRefl;
        if (twoSided) {
            np = w.normalPrev[slot];
            float4 dt = w.dirThr[slot];
            L.st.normal = vec3of<DIM>(np);
            L.st.prevDist = np.w;
            L.st.prevDir = vec3of<DIM>(dt);
        }
        L.firstSphereRadius = L.firstStep ? w.acc[slot].w : 0.0f;
        float starRadius = 0.0f, r2max = 0.0f;
        bool need = stepStarPrologue<DIM>(a, L, starRadius, r2max);
        if (twoSided && L.flip) w.normalPrev[slot] = pack4<DIM>(L.st.normal, np.w);
        flags = (flags & ~(kWfFlip | kWfNeedStar)) | (L.flip ? kWfFlip : 0u) | (need ? kWfNeedStar : 0u);
        if (!need) {
            float4 ps = w.projStar[slot];
            ps.w = starRadius;
            w.projStar[slot] = ps;
        }
    }
    m.w = (int)flags;
    w.meta[slot] = m;
#endif
}

template <int DIM, bool WOST>
__global__ void __launch_bounds__(128, ZB_CLOSEST_MIN_BLOCKS) waveClosestKernel(SolveArgs a, WaveSlots w, int nBound, int nWarps)
{
#if defined(__CUDA_ARCH__)
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    for (int v = nextVirtualWarp(w.live + 1, nWarps); v >= 0; v = nextVirtualWarp(w.live + 1, nWarps)) {
        waveClosestWarp<DIM, WOST>(a, w, ws, nBound, (int)(threadIdx.x & 31)*nWarps + v);
        __syncwarp();
    }
#endif
}

// ------------------------------------------------------------------------------------------
// star radius
// ------------------------------------------------------------------------------------------
// One kernel runs both phases of the cooperative query (coop.cuh). Kept for A/B measurements (ZB_WAVE_STAR_SPLIT=0) and as
// the statement of what the two kernels below compute together.
template <int DIM>
ZB_DEV void waveStarWarp(const SolveArgs& a, const WaveSlots& w, WarpShared& ws, int nBound, int i)
{
#if defined(__CUDA_ARCH__)
    int slot = -1;
    uint32_t flags = 0u;
    if (i < nBound) { slot = (int)w.order[i]; flags = (uint32_t)w.meta[slot].w; }
    const bool qs = (flags & (kWfActive | kWfNeedStar)) == (kWfActive | kWfNeedStar);
    if (!__any_sync(0xffffffffu, qs)) return;
    float4 pd = qs ? w.ptDist[slot] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float trivial = 0.0f, r2max = 0.0f;
    const float eps = a.ws.epsilonShellForReflectingBoundary;
    if (qs) starQueryNeeded<DIM>(a, eps, pd.w, trivial, r2max);
    float r2 = coopStarRadius2<DIM>(a.reflecting, ws, qs, vec3of<DIM>(pd), r2max, !(flags & kWfFlip), a.ws.silhouettePrecision, a.status);
    if (qs) {
        float r = starQueryResult(a, r2, r2max, eps, pd.w);
        float4 ps = w.projStar[slot];
        ps.w = shrinkStarRadius(a, r, pd.w);
        w.projStar[slot] = ps;
    }
#endif
}

template <int DIM>
__global__ void __launch_bounds__(128, ZB_STAR_MIN_BLOCKS) waveStarKernel(SolveArgs a, WaveSlots w, int nBound, int nWarps)
{
#if defined(__CUDA_ARCH__)
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    for (int v = nextVirtualWarp(w.live + 2, nWarps); v >= 0; v = nextVirtualWarp(w.live + 2, nWarps)) {
        waveStarWarp<DIM>(a, w, ws, nBound, (int)(threadIdx.x & 31)*nWarps + v);
        __syncwarp();
    }
#endif
}

// The shipped form: the two phases of the query are two kernels. The fused kernel is 80 KB of SASS of which 44 KB are hot
// (private depth-first loop, pooled node / filter / full rounds, the primitive test both call); with 24 warps per SM each
// in a different phase, 48 % of its warp-stall samples are `no_instruction` (profiles/r2_wave_steady_ncu.txt: the 32 KB
// instruction cache thrashes) and the issue slots are 42 % busy, against 4 % and 69 % for the distance query's 25 KB loop.
// Split, every warp of a launch runs the same loop, the private phase needs no shared memory (more resident warps), and the
// pooled phase only sees the warps that have something left. The hand-over — radius so far and the unfinished stack, a few
// dozen bytes for the queries that have one — goes through HBM. Virtual warps are composed exactly as in the fused
// kernel, so the two forms return the same bits.
template <int DIM>
__global__ void __launch_bounds__(128, ZB_STAR_PRIVATE_MIN_BLOCKS) waveStarPrivateKernel(SolveArgs a, WaveSlots w, int nBound, int nWarps)
{
#if defined(__CUDA_ARCH__)
    const int t = blockIdx.x*blockDim.x + threadIdx.x; // virtual thread: virtual warp t >> 5 (the grid holds nWarps of them), lane t & 31
    if ((t >> 5) >= nWarps) return;
    const int i = (t & 31)*nWarps + (t >> 5);
    int slot = -1;
    uint32_t flags = 0u;
    if (i < nBound) { slot = (int)w.order[i]; flags = (uint32_t)w.meta[slot].w; }
    const bool qs = (flags & (kWfActive | kWfNeedStar)) == (kWfActive | kWfNeedStar);
    int sp = 0;
    if (qs) {
        float4 pd = w.ptDist[slot];
        float trivial = 0.0f, r2max = 0.0f;
        const float eps = a.ws.epsilonShellForReflectingBoundary;
        starQueryNeeded<DIM>(a, eps, pd.w, trivial, r2max);
        uint2 stack[kStackSize];
        float r2 = r2max;
        stack[sp++] = make_uint2(0u, 0u);
        starRadiusDfs<DIM, true>(a.reflecting, vec3of<DIM>(pd), r2, !(flags & kWfFlip), a.ws.silhouettePrecision, stack, sp, kPhase1Budget, a.status);
        if (sp == 0) {
            float r = starQueryResult(a, r2, r2max, eps, pd.w);
            float4 ps = w.projStar[slot];
            ps.w = shrinkStarRadius(a, r, pd.w);
            w.projStar[slot] = ps;
        } else {
            w.starPending[t] = make_uint2(zb_float_as_uint(r2), (uint32_t)sp);
            for (int k = 0; k < sp; k++) w.starStack[(size_t)k*w.starStride + t] = stack[k];
        }
    }
    const unsigned pending = __ballot_sync(0xffffffffu, sp > 0);
    if ((t & 31) == 0) w.starWarpMask[t >> 5] = pending;
#endif
}

template <int DIM>
__global__ void __launch_bounds__(128, ZB_STAR_MIN_BLOCKS) waveStarPooledKernel(SolveArgs a, WaveSlots w, int nBound, int nWarps)
{
#if defined(__CUDA_ARCH__)
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    for (int v = nextVirtualWarp(w.live + 2, nWarps); v >= 0; v = nextVirtualWarp(w.live + 2, nWarps)) {
        const uint32_t pending = w.starWarpMask[v];
        if (pending == 0u) continue;
        const bool qs = (pending >> lane) & 1u;
        const int t = 32*v + lane;
        int slot = -1;
        uint32_t flags = 0u;
        float4 pd = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float trivial = 0.0f, r2max = 0.0f, r2own = 0.0f;
        int sp = 0;
        const float eps = a.ws.epsilonShellForReflectingBoundary;
        if (qs) {
            slot = (int)w.order[lane*nWarps + v];
            flags = (uint32_t)w.meta[slot].w;
            pd = w.ptDist[slot];
            starQueryNeeded<DIM>(a, eps, pd.w, trivial, r2max);
            uint2 pr = w.starPending[t];
            r2own = zb_uint_as_float(pr.x);
            sp = (int)pr.y;
        }
        float r2 = coopStarPooled<DIM>(a.reflecting, ws, qs, vec3of<DIM>(pd), qs ? r2own : r2max, r2max, !(flags & kWfFlip), a.ws.silhouettePrecision,
                                       a.status, StridedStack{w.starStack + t, w.starStride}, sp);
        if (qs) {
            float r = starQueryResult(a, r2, r2max, eps, pd.w);
            float4 ps = w.projStar[slot];
            ps.w = shrinkStarRadius(a, r, pd.w);
            w.projStar[slot] = ps;
        }
        __syncwarp();
    }
#endif
}

// ------------------------------------------------------------------------------------------
// the rest of the step, termination, next task
// ------------------------------------------------------------------------------------------
template <int DIM>
ZB_DEV void loadLane(const SolveArgs& a, const WaveSlots& w, int slot, const int4& m, Lane<DIM>& L)
{
    const uint32_t flags = (uint32_t)m.w;
    float4 pd = w.ptDist[slot], ps = w.projStar[slot], np = w.normalPrev[slot], dt = w.dirThr[slot], ac = w.acc[slot];
    ulonglong2 rg = w.rng[slot];
    int4 m2 = w.meta2[slot];
    L.st.pt = vec3of<DIM>(pd);
    L.distA = pd.w;
    L.proj = vec3of<DIM>(ps);
    L.st.normal = vec3of<DIM>(np);
    L.st.prevDist = np.w;
    L.st.prevDir = vec3of<DIM>(dt);
    L.st.throughput = dt.w;
    L.st.reflContrib = ac.x;
    L.st.srcContrib = ac.y;
    L.total = ac.z;
    L.firstSphereRadius = ac.w;
    L.rng.state = rg.x;
    L.rng.inc = rg.y;
    L.point = m.x;
    L.st.walkLength = m.y;
    L.totalLength = m.z;
    L.task = (long long)(((unsigned long long)(uint32_t)m2.y << 32) | (unsigned long long)(uint32_t)m2.x);
    L.splits = m2.z;
    L.nQueued = m2.w;
    L.active = true;
    L.needClosest = (flags & kWfNeedClosest) != 0u;
    L.keepDistA = (flags & kWfKeepDistA) != 0u;
    L.firstStep = (flags & kWfFirstStep) != 0u;
    L.flip = (flags & kWfFlip) != 0u;
    L.outerFlip = (flags & kWfOuterFlip) != 0u;
    L.success = (flags & kWfSuccess) != 0u;
    L.gradient = (flags & kWfGradient) != 0u;
    L.st.onRefl = (flags & kWfOnRefl) != 0u;
    L.needsFinalProjection = false;
    L.projValid = (flags & kWfProjValid) != 0u;
    L.dirichletAligned = (flags & kWfAligned) != 0u;
    L.g.init((flags & kWfYukawa) != 0u, a.pde.absorptionCoeff);
    L.firstSource = 0.0f;
    L.gradBoundary = vzero<DIM>();
    L.gradSource = vzero<DIM>();
}

template <int DIM>
ZB_DEV uint32_t storeLane(const WaveSlots& w, int slot, const Lane<DIM>& L)
{
    uint32_t flags = kWfActive | (L.needClosest ? kWfNeedClosest : 0u) | (L.keepDistA ? kWfKeepDistA : 0u) |
                     (L.firstStep ? kWfFirstStep : 0u) | (L.flip ? kWfFlip : 0u) | (L.outerFlip ? kWfOuterFlip : 0u) |
                     (L.success ? kWfSuccess : 0u) | (L.gradient ? kWfGradient : 0u) | (L.st.onRefl ? kWfOnRefl : 0u) |
                     (L.g.yukawa ? kWfYukawa : 0u) | (L.needsFinalProjection ? kWfFinalProjection : 0u) |
                     (L.projValid ? kWfProjValid : 0u) | (L.dirichletAligned ? kWfAligned : 0u);
    w.ptDist[slot] = pack4<DIM>(L.st.pt, L.distA);
    w.projStar[slot] = pack4<DIM>(L.proj, 0.0f);
    w.normalPrev[slot] = pack4<DIM>(L.st.normal, L.st.prevDist);
    w.dirThr[slot] = pack4<DIM>(L.st.prevDir, L.st.throughput);
    w.acc[slot] = make_float4(L.st.reflContrib, L.st.srcContrib, L.total, L.firstSphereRadius);
    ulonglong2 rg;
    rg.x = L.rng.state;
    rg.y = L.rng.inc;
    w.rng[slot] = rg;
    w.meta[slot] = make_int4(L.point, L.st.walkLength, L.totalLength, (int)flags);
    w.meta2[slot] = make_int4((int)(uint32_t)((unsigned long long)L.task & 0xffffffffull), (int)(uint32_t)((unsigned long long)L.task >> 32),
                              L.splits, L.nQueued);
    return flags;
}

// the gradient estimator's per-task factors are known when the task begins (beginTask): they go straight to its record
template <int DIM>
ZB_DEV void storeGradientFactors(const SolveArgs& a, const Lane<DIM>& L)
{
    if (!a.anyGradient) return;
    long long t = L.task;
    a.recFirstSource[t] = L.firstSource;
    for (int k = 0; k < DIM; k++) {
        a.recGradBoundary[(size_t)DIM*t + k] = L.gradBoundary[k];
        a.recGradSource[(size_t)DIM*t + k] = L.gradSource[k];
    }
}

// finishState of the wavefront: as walk.cuh's, but the record's gradient factors were written when the task began
template <int DIM, bool WOST, bool CH = false>
ZB_DEV void finishStateWave(const SolveArgs& a, Lane<DIM>& L, StepResult code, WalkState<DIM> *queue)
{
    PdeEval<DIM> pde{a.pde};
    L.splits++;
    if (code != kEscaped) {
        float terminal = 0.0f;
        bool project = WOST ? code == kReached : (code == kReached || code == kMaxLength);
        if (project && !a.ws.ignoreAbsorbingBoundaryContribution) terminal = pde.dirichlet(a.hasAbsorbing ? L.proj : L.st.pt, L.dirichletAligned);
        L.total += L.st.throughput*terminal + L.st.reflContrib + L.st.srcContrib;
        L.totalLength += L.st.walkLength;
        L.success = true;
        if (CH) extraFinishState<DIM>(a.extra, L.task, true, project && !a.ws.ignoreAbsorbingBoundaryContribution,
                                            a.hasAbsorbing ? L.proj : L.st.pt, L.st.throughput, L.dirichletAligned);
    } else if (CH) {
        extraFinishState<DIM>(a.extra, L.task, false, false, L.st.pt, 0.0f, false);
    }
    if (L.nQueued > 0) {
        L.st = queue[--L.nQueued];
        bool yukawa = a.pde.absorptionCoeff > 0.0f && a.ws.stepsBeforeApplyingTikhonov == 0;
        L.g.init(yukawa, a.pde.absorptionCoeff);
        L.distA = kMaxFloat;
        L.needClosest = true;
        L.keepDistA = false;
        L.firstSphereRadius = 0.0f;
        L.firstStep = true;
        L.flip = L.outerFlip;
        L.needsFinalProjection = false;
        return;
    }
    long long t = L.task;
    a.recSolution[t] = L.total;
    a.recLength[t] = L.totalLength;
    a.recFlags[t] = (L.success ? (int)0x80000000 : 0) | (L.splits & 0x7fffffff);
    L.active = false;
}

#ifndef ZB_STEP_MIN_BLOCKS
#define ZB_STEP_MIN_BLOCKS 1 // resident CTAs per SM the step kernel is compiled for (1: no register cap, 129 registers -> 3 CTAs; A/B builds)
#endif
template <int DIM, bool WOST, bool CH = false>
__global__ void __launch_bounds__(128, ZB_STEP_MIN_BLOCKS) waveStepKernel(SolveArgs a, WaveSlots w, int nBound)
{
#if defined(__CUDA_ARCH__)
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int slot = -1;
    int4 m = make_int4(0, 0, 0, 0);
    if (i < nBound) { slot = (int)w.order[i]; m = w.meta[slot]; }
    const uint32_t flags = (uint32_t)m.w;
    Lane<DIM> L;
    L.active = false;
    WalkState<DIM> *queue = w.queue && slot >= 0 ? reinterpret_cast<WalkState<DIM> *>(w.queue) + (size_t)slot*kSplitStack : nullptr;

    if (flags & kWfActive) {
        loadLane<DIM>(a, w, slot, m, L);
        if (flags & kWfFinalProjection) {
            L.needsFinalProjection = false;
            finishStateWave<DIM, WOST, CH>(a, L, kMaxLength, queue);
        } else if (flags & kWfReached) {
            finishStateWave<DIM, WOST, CH>(a, L, kReached, queue);
        } else if (!L.needClosest) {
            const Vec<DIM> ptBefore = L.st.pt;
            const float distBefore = L.distA;
            StepResult r = stepBody<DIM, WOST, CH>(a, L, queue, w.projStar[slot].w);
            if (r == kEscaped && w.debugEscapes)
                printf("escaped: task %lld point %d step %d from (%.9g %.9g %.9g) distA %.9g ball %.9g proj (%.9g %.9g %.9g) to (%.9g %.9g %.9g)\n", L.task, L.point,
                       L.st.walkLength, ptBefore[0], ptBefore[1], DIM == 3 ? ptBefore[DIM - 1] : 0.0f, distBefore, L.g.R, L.proj[0], L.proj[1],
                       DIM == 3 ? L.proj[DIM - 1] : 0.0f, L.st.pt[0], L.st.pt[1], DIM == 3 ? L.st.pt[DIM - 1] : 0.0f);
            // Walk on Spheres ends a walk that ran out of steps with the Dirichlet value at the projection of its LAST
            // position (walk_on_spheres.h:294-308): that projection is one more distance query (one round) away
            if (!WOST && r == kMaxLength && L.needsFinalProjection && a.hasAbsorbing && !a.ws.ignoreAbsorbingBoundaryContribution) {
                L.needClosest = true;
            } else if (r != kContinue) {
                L.needsFinalProjection = false;
                finishStateWave<DIM, WOST, CH>(a, L, r, queue);
            }
        }
    }

    // ---- compaction: slots whose walk ended take the next tasks of the global queue
    const bool want = slot >= 0 && !L.active;
    const unsigned need = __ballot_sync(full, want);
    if (need) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(w.taskCounter, (unsigned long long)__popc(need));
        base = __shfl_sync(full, base, 0);
        long long task = ((long long)base + __popc(need & ((1u << lane) - 1u)))*w.taskStride + w.taskFirst;
        if (want && task < a.totalTasks) {
            beginTask<DIM, WOST, CH>(a, L, task);
            if (L.gradient) storeGradientFactors<DIM>(a, L);
        }
    }

    uint32_t key = 0xffffffffu;
    if (slot >= 0) {
        if (L.active) {
            storeLane<DIM>(w, slot, L);
            // mode 0: a fixed pseudo-random rank per task (odd multiplier = bijection on 32 bits): which slot a task landed in
            // depends on the order of the atomics above, the ORDER of the walks does not — warps are composed the same way
            // on every run, which is what makes the cooperative queries, hence the estimates, repeatable bit for bit
            key = w.sortMode == 2 ? (~zb_float_as_uint(fminf(L.distA, 1e30f)) & 0x7fffffffu)
                : (w.sortMode == 1 ? mortonKey<DIM>(w, L.st.pt) : ((uint32_t)L.task*2654435761u) >> 1);
        } else if (flags & kWfActive) {
            m.w = 0;
            w.meta[slot] = m;
        }
        w.keys[i] = key;
    }
    const unsigned alive = __ballot_sync(full, L.active);
    if (lane == 0 && alive) atomicAdd(w.live, (unsigned int)__popc(alive));
#endif
}

// ------------------------------------------------------------------------------------------
// the tail of a solve: the last walks run to their end without round barriers
// ------------------------------------------------------------------------------------------
// A round lasts as long as its slowest query: ~1 ms for a query that opens thousands of primitives with a single warp behind
// it, whatever the number of live walks. Once the tasks have run out, the solve's critical path is (steps of the longest walk)
// x (that latency) — measured 0.9 s per solve of C3, independent of the batch size — while the machine is nearly idle. The
// walks are independent, so the last ones need no barrier at all: this kernel takes the live slots (the front of the slot
// order), spreads them over a persistent grid — one walk per warp when there are fewer walks than warps, the 32 lanes pooling
// that walk's traversals — and every warp steps its walks until they end. The critical path becomes the longest walk's OWN
// queries. It is the round-1 persistent loop (walkKernel) resumed from slot state, without a task queue.
template <int DIM, bool WOST, bool CH = false>
__global__ void __launch_bounds__(128, ZB_WALK_MIN_BLOCKS) waveTailKernel(SolveArgs a, WaveSlots w, int nBound, int nWarps)
{
#if defined(__CUDA_ARCH__)
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    const unsigned full = 0xffffffffu;
    const int i = stripedPosition(nWarps);
    int slot = -1;
    int4 m = make_int4(0, 0, 0, 0);
    if (i < nBound) { slot = (int)w.order[i]; m = w.meta[slot]; }
    Lane<DIM> L;
    L.active = false;
    L.needClosest = false;
    L.keepDistA = false;
    L.needsFinalProjection = false;
    if ((uint32_t)m.w & kWfActive) {
        loadLane<DIM>(a, w, slot, m, L);
        L.needsFinalProjection = ((uint32_t)m.w & kWfFinalProjection) != 0u;
    }
    WalkState<DIM> *queue = w.queue && slot >= 0 ? reinterpret_cast<WalkState<DIM> *>(w.queue) + (size_t)slot*kSplitStack : nullptr;

    while (__any_sync(full, L.active)) {
        // ---- pending distance queries of the warp, cooperatively
        bool qc = L.active && L.needClosest;
        if (__any_sync(full, qc)) {
            if (a.hasAbsorbing) {
                ClosestHit<DIM> hit;
                coopClosestPoint<DIM>(a.absorbing, ws, qc, L.st.pt, qc ? closestSearchRadius2<DIM, WOST>(a, L) : 0.0f, hit, a.status);
                if (qc) {
                    if (!L.keepDistA) L.distA = hit.d;
                    L.proj = hit.p;
                    L.projValid = hit.prim >= 0;
                    L.dirichletAligned = onNormalAlignedSide<DIM>(a, L.st.pt, hit.p, hit.prim);
                    L.needClosest = false;
                    L.keepDistA = false;
                }
            } else if (qc) {
                resolveClosest<DIM, WOST>(a, L);
            }
        }
        // ---- loop condition of walk()
        if (L.active && !L.needClosest) {
            if (L.needsFinalProjection) {
                L.needsFinalProjection = false;
                finishStateWave<DIM, WOST, CH>(a, L, kMaxLength, queue);
            } else if (!(L.distA > a.ws.epsilonShellForAbsorbingBoundary)) {
                finishStateWave<DIM, WOST, CH>(a, L, kReached, queue);
            }
        }
        bool ready = L.active && !L.needClosest;

        // ---- star radii of the warp, cooperatively
        float starRadius = 0.0f;
        if (WOST) {
            float r2max = 0.0f;
            bool qs = ready && stepStarPrologue<DIM>(a, L, starRadius, r2max);
            if (__any_sync(full, qs)) {
                float r2 = coopStarRadius2<DIM>(a.reflecting, ws, qs, L.st.pt, r2max, !L.flip, a.ws.silhouettePrecision, a.status);
                if (qs) {
                    float r = starQueryResult(a, r2, r2max, a.ws.epsilonShellForReflectingBoundary, L.distA);
                    starRadius = shrinkStarRadius(a, r, L.distA);
                }
            }
        }

        // ---- the rest of the step, one thread per walk
        if (ready) {
            StepResult r = stepBody<DIM, WOST, CH>(a, L, queue, starRadius);
            if (!WOST && r == kMaxLength && L.needsFinalProjection && a.hasAbsorbing && !a.ws.ignoreAbsorbingBoundaryContribution) {
                L.needClosest = true;
            } else if (r != kContinue) {
                L.needsFinalProjection = false;
                finishStateWave<DIM, WOST, CH>(a, L, r, queue);
            }
        }
    }
#endif
}

// The step kernel runs one thread per walk through per-thread traversals (ray cast, sampling descent, occlusion ray) whose
// trip counts grow with the star radius: a warp lasts as long as the lane with the largest ball (3.9 of 32 lanes per instruction
// in the hash order the cooperative query kernels want). This kernel classes every position of the round's order by the work its
// walk will do there — 0: no step (the walk ends or waits), 1..: eight classes per octave of the star radius — and an 8-bit
// stable radix sort (one pass) of the order by that class gives the step kernel warps of similar walks. The step kernel's
// results do not depend on its order (no cooperative operation in it), and the next round's order is sorted from the task keys
// again, so estimates stay bit-identical.
__global__ void waveStepKeyKernel(WaveSlots w, int nBound, uint32_t *stepKeys)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= nBound) return;
    const int slot = (int)w.order[i];
    const uint32_t flags = (uint32_t)w.meta[slot].w;
    uint32_t key = 255u;
    if (flags & kWfActive) {
        key = 0u;
        if (!(flags & (kWfNeedClosest | kWfReached | kWfFinalProjection))) {
            float l = (log2f(fmaxf(w.projStar[slot].w, 4.8828125e-4f)) + 11.0f)*8.0f;
            key = 1u + (uint32_t)fminf(fmaxf(l, 0.0f), 120.0f);
        }
    }
    stepKeys[i] = key;
}

// slots of a new solve: none holds a walk, order = identity
__global__ void waveResetKernel(WaveSlots w, uint32_t *order)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < w.nSlots) {
        w.meta[i] = make_int4(0, 0, 0, 0);
        order[i] = (uint32_t)i;
    }
}

} // namespace zb
