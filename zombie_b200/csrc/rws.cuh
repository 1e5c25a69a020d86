// Reverse Walk on Stars (BASELINE config 4, second half).
//
// Reference: rws::ReverseWalkOnStarsSolver::solve (zombie/variance_reduction/reverse_walk_solver.h:287-330) starts one
// walk at every boundary / domain sample (ReverseWalkOnStars::solve, zombie/point_estimation/reverse_walk_on_stars.h:
// 89-166, walk :268-393); at every step it searches a KD-tree for the evaluation points inside the current ball, tests
// each for visibility and adds weight * contribution to one of five running sums of that evaluation point under a
// per-point spin mutex (reverse_walk_solver.h:233-283).
//
// Here the scatter becomes an atomic-free gather in two phases:
//   1. the walks run in the persistent walk kernel (walk.cuh) with boundary / source terms switched off — a reverse
//      walk IS a forward WoSt walk without them — and every step appends one 48-byte splat record (rwsEmit): the ball
//      and the weight the walk carries, everything that does not depend on the evaluation point;
//   2. rwsGatherKernel: one thread per evaluation point streams all records through shared memory in tiles, keeps the
//      ones whose ball contains its point and that it can see (an any-hit ray against the reflecting BVH), evaluates the
//      ball's Green's function and accumulates the six sums in registers. No locks, no atomics, and the order of
//      additions per evaluation point is fixed: records are sorted by a key that ends in their identity (task, step).
// The gather culls on both sides (sorted record tiles with bounding boxes against Morton-sorted evaluation points, below).
#pragma once

#include "wave.cuh"

namespace zb {

constexpr int kRwsTile = 128;      // records per shared-memory tile (6 KB)
constexpr int kRwsThreads = 128;

// ---- per-sample contribution (ReverseWalkOnStars::solve, reverse_walk_on_stars.h:95-124) -----------------------
// source samples: pde.source at the sample; absorbing samples: the Dirichlet value at the sample projected back onto the
// boundary along its normal by the signed distance; reflecting samples: the Robin / Neumann data at the sample.
template <int DIM>
ZB_DEV float rwsSampleContribution(const SolveArgs& a, int i, const ClosestHit<DIM>& hit)
{
    PdeEval<DIM> pde{a.pde};
    int type = a.pts.type[i];
    Vec<DIM> x = loadVec<DIM>(a.pts.pt + (size_t)DIM*i);
    Vec<DIM> n = loadVec<DIM>(a.pts.normal + (size_t)DIM*i);
    if (type == ZB_IN_DOMAIN) return a.ws.ignoreSourceContribution ? 0.0f : pde.source(x);
    if (type == ZB_ON_ABSORBING_BOUNDARY) {
        if (a.ws.ignoreAbsorbingBoundaryContribution) return 0.0f;
        // signed distance as FCPW computes it (fcpw_geometric_queries.h:1069-1077): sign of (x - p) . n_p
        float signedDistance = hit.d;
        if (hit.prim >= 0) {
            Prim<DIM> prim;
            prim.load(a.absorbing, (uint32_t)hit.prim);
            if (dot(x - hit.p, unitNormal<DIM>(prim)) < 0.0f) signedDistance = -hit.d;
        }
        return pde.dirichlet(x - n*signedDistance);
    }
    return a.ws.ignoreReflectingBoundaryContribution ? 0.0f : pde.robin(x, n);
}

template <int DIM>
__global__ void __launch_bounds__(128) rwsContributionKernel(SolveArgs a, float *contribution)
{
#if defined(__CUDA_ARCH__)
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    bool has = i < a.pts.n;
    Vec<DIM> x = has ? loadVec<DIM>(a.pts.pt + (size_t)DIM*i) : vzero<DIM>();
    bool project = has && a.pts.type[i] == ZB_ON_ABSORBING_BOUNDARY && a.hasAbsorbing;
    ClosestHit<DIM> hit;
    hit.d = 0.0f; hit.prim = -1; hit.p = x;
    if (__any_sync(0xffffffffu, project)) coopClosestPoint<DIM>(a.absorbing, ws, project, x, kMaxFloat, hit, a.status);
    if (has) contribution[i] = rwsSampleContribution<DIM>(a, i, hit);
#endif
}

// ---- gather -------------------------------------------------------------------------------------------------------
struct RwsGatherArgs {
    BvhView reflecting;
    int hasReflecting;
    const float4 *records;
    long long nRecords;
    int nEval;
    const float *evalPt;       // nEval*DIM
    const float *evalNormal;   // nEval*DIM
    const int *evalType;
    float *totals;             // nEval*6: poisson kernel, absorbing, absorbing aligned, reflecting, reflecting aligned, source; in/out
    float sqrtLambda;
    float radiusClamp, kernelRegularization;
    int useSelfNormalization;
    int *status;
};

// GreensFnBall::evaluate(r) for a ball of radius R (distributions.h:587-590, 892-897; 2D: :486-489)
template <int DIM>
ZB_DEV float rwsBallGreens(bool yukawa, float sqrtLambda, float R, float r)
{
    if (!yukawa) return DIM == 3 ? (1.0f/r - 1.0f/R)/(4.0f*kPiF) : logf(R/r)/(2.0f*kPiF);
    GreensBall<DIM> g;
    g.init(true, sqrtLambda*sqrtLambda);
    Vec<DIM> c = vzero<DIM>();
    g.updateBall(c, R);
    return g.evaluate(r);
}

// contribution of one record to one evaluation point: returns false when the point is outside the ball or occluded
template <int DIM>
ZB_DEV bool rwsPairWeight(const RwsGatherArgs& a, const float4& q0, const float4& q1, uint32_t flags, const Vec<DIM>& e,
                          const Vec<DIM>& eNormal, bool evalOnRefl, float& G)
{
    Vec<DIM> c;
    c[0] = q0.x; c[1] = q0.y;
    if (DIM == 3) c[DIM - 1] = q0.z;
    float R = q0.w;
    Vec<DIM> d = e - c;
    float d2 = dot(d, d);
    if (!(d2 < R*R)) return false; // radiusSearch (nearest_neighbor_finder.h:96-109)
    if (a.hasReflecting) {
        // intersectsWithReflectingBoundary (fcpw_geometric_queries.h:1197-1226): both end points are pushed off the boundary
        // they lie on before the occlusion ray is cast
        Vec<DIM> cn;
        cn[0] = q1.x; cn[1] = q1.y;
        if (DIM == 3) cn[DIM - 1] = q1.z;
        Vec<DIM> p1 = (flags & 8u) ? offsetPoint<DIM>(c, -cn) : c;
        Vec<DIM> p2 = evalOnRefl ? offsetPoint<DIM>(e, -eNormal) : e;
        Vec<DIM> dir = p2 - p1;
        float len = norm(dir);
        if (len > 0.0f) {
            RayHit<DIM> occ;
            if (intersectRay<DIM, true>(a.reflecting, p1, dir/len, len, occ, a.status)) return false;
        }
    }
    float r = fmaxf(a.radiusClamp, sqrtf(d2));
    G = rwsBallGreens<DIM>((flags & 16u) != 0u, a.sqrtLambda, R, r);
    if (a.kernelRegularization > 0.0f) {
        float rr = r/a.kernelRegularization;
        if (DIM == 3) G *= erff(rr); // KernelRegularization<3>::regularizationForGreensFn (distributions.h:44-47); 2D: 1
    }
    return true;
}

// one record onto one evaluation point (reverse_walk_solver.h:247-283)
template <int DIM>
ZB_DEV void rwsAccumulate(const RwsGatherArgs& a, const float4& q0, const float4& q1, const float4& q2, const Vec<DIM>& e,
                          const Vec<DIM>& eNormal, bool evalOnRefl, float sums[6])
{
    uint32_t flags = zb_float_as_uint(q2.y);
    float G;
    if (!rwsPairWeight<DIM>(a, q0, q1, flags, e, eNormal, evalOnRefl, G)) return;
    uint32_t bucket = flags & 7u;
    sums[1 + bucket] += q2.x*G;                                    // weight * contribution (:262-281)
    if (bucket == 0u && a.useSelfNormalization) sums[0] += q1.w*G; // totalPoissonKernelContribution (:268)
}

// ---- spatial culling of the gather ---------------------------------------------------------------------------------
// The reference finds the evaluation points of a ball with a KD-tree radius search per walk step
// (nearest_neighbor_finder.h:96-109, reverse_walk_solver.h:233-246). The gather turns that around — which balls contain
// this evaluation point? — and culls on both sides:
//   * records are sorted by (radius octave, Morton code of the ball's centre, record identity) and cut into tiles of 128
//     with the bounding box of their balls; 64 tiles form a super-tile with its own box. Sorting by radius octave first
//     keeps small balls from sharing a tile (and a box) with large ones; the identity (task, step) as the last key makes
//     the order, hence every float sum, independent of the order the walks appended their records in;
//   * evaluation points are sorted by Morton code, so the 128 points of a CTA are neighbours with a small bounding box.
// A CTA walks the super-tiles, descends into the tiles whose box meets its own, and streams only those through shared
// memory; every decision is uniform over the CTA. Work falls from records x points to about (ball, point) pairs in range.
constexpr int kRwsSuper = 64; // tiles per super-tile

struct RwsCullArgs {
    const float4 *records;       // in emission order
    float4 *sorted;              // in gather order
    unsigned long long *keys;
    uint32_t *index;
    long long nRecords;
    float4 *tileLo, *tileHi;     // per tile, then per super-tile (offset nTiles)
    int nTiles, nSuper;
    float lo[3], scale[3];       // Morton frame
    float invDiag;
    // evaluation points
    const float *evalPt;
    uint32_t *evalKeys, *evalIndex;
    int nEval;
};

template <int DIM>
__global__ void rwsRecordKeysKernel(RwsCullArgs c)
{
    long long r = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (r >= c.nRecords) return;
    float4 q0 = c.records[3*r], q2 = c.records[3*r + 2];
    uint32_t q[3] = {0u, 0u, 0u};
    const float pos[3] = {q0.x, q0.y, q0.z};
    for (int k = 0; k < DIM; k++) q[k] = (uint32_t)fminf(fmaxf((pos[k] - c.lo[k])*c.scale[k], 0.0f), DIM == 3 ? 1023.0f : 32767.0f);
    uint32_t morton = DIM == 3 ? (spreadBits3(q[0]) | (spreadBits3(q[1]) << 1) | (spreadBits3(q[2]) << 2)) : (spreadBits2(q[0]) | (spreadBits2(q[1]) << 1));
    int octave = (int)floorf(-log2f(fmaxf(q0.w*c.invDiag, 1e-30f)));
    octave = octave < 0 ? 0 : (octave > 7 ? 7 : octave);
    uint32_t id = (zb_float_as_uint(q2.z)*2654435761u) ^ (zb_float_as_uint(q2.w)*40503u);
    c.keys[r] = ((unsigned long long)octave << 61) | ((unsigned long long)(morton & 0x3fffffffu) << 31) | (unsigned long long)(id & 0x7fffffffu);
    c.index[r] = (uint32_t)r;
}

__global__ void rwsReorderKernel(RwsCullArgs c)
{
    long long r = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (r >= c.nRecords) return;
    long long src = c.index[r];
    c.sorted[3*r] = c.records[3*src];
    c.sorted[3*r + 1] = c.records[3*src + 1];
    c.sorted[3*r + 2] = c.records[3*src + 2];
}

// one warp per tile: box of the tile's balls
__global__ void rwsTileBoundsKernel(RwsCullArgs c)
{
#if defined(__CUDA_ARCH__)
    int tile = (blockIdx.x*blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (tile >= c.nTiles) return;
    float lo[3] = {kMaxFloat, kMaxFloat, kMaxFloat}, hi[3] = {-kMaxFloat, -kMaxFloat, -kMaxFloat};
    for (int j = lane; j < kRwsTile; j += 32) {
        long long r = (long long)tile*kRwsTile + j;
        if (r >= c.nRecords) break;
        float4 q0 = c.sorted[3*r];
        const float p[3] = {q0.x, q0.y, q0.z};
        for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], p[k] - q0.w); hi[k] = fmaxf(hi[k], p[k] + q0.w); }
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if (lane == 0) {
        c.tileLo[tile] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        c.tileHi[tile] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    }
#endif
}

// one warp per super-tile: box of its tiles' boxes
__global__ void rwsSuperBoundsKernel(RwsCullArgs c)
{
#if defined(__CUDA_ARCH__)
    int super = (blockIdx.x*blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (super >= c.nSuper) return;
    float lo[3] = {kMaxFloat, kMaxFloat, kMaxFloat}, hi[3] = {-kMaxFloat, -kMaxFloat, -kMaxFloat};
    for (int j = lane; j < kRwsSuper; j += 32) {
        int t = super*kRwsSuper + j;
        if (t >= c.nTiles) break;
        float4 a = c.tileLo[t], b = c.tileHi[t];
        lo[0] = fminf(lo[0], a.x); lo[1] = fminf(lo[1], a.y); lo[2] = fminf(lo[2], a.z);
        hi[0] = fmaxf(hi[0], b.x); hi[1] = fmaxf(hi[1], b.y); hi[2] = fmaxf(hi[2], b.z);
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if (lane == 0) {
        c.tileLo[c.nTiles + super] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        c.tileHi[c.nTiles + super] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    }
#endif
}

template <int DIM>
__global__ void rwsEvalKeysKernel(RwsCullArgs c)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= c.nEval) return;
    uint32_t q[3] = {0u, 0u, 0u};
    for (int k = 0; k < DIM; k++) q[k] = (uint32_t)fminf(fmaxf((c.evalPt[(size_t)DIM*i + k] - c.lo[k])*c.scale[k], 0.0f), DIM == 3 ? 1023.0f : 32767.0f);
    c.evalKeys[i] = DIM == 3 ? (spreadBits3(q[0]) | (spreadBits3(q[1]) << 1) | (spreadBits3(q[2]) << 2)) : (spreadBits2(q[0]) | (spreadBits2(q[1]) << 1));
    c.evalIndex[i] = (uint32_t)i;
}

template <int DIM>
__global__ void __launch_bounds__(kRwsThreads) rwsGatherKernel(RwsGatherArgs a, const uint32_t *evalOrder, const float4 *tileLo,
                                                               const float4 *tileHi, int nTiles, int nSuper)
{
#if defined(__CUDA_ARCH__)
    __shared__ float4 tile[kRwsTile*3];
    __shared__ float boxLo[3][kRwsThreads/32], boxHi[3][kRwsThreads/32];
    const int pos = blockIdx.x*blockDim.x + threadIdx.x;
    const bool live = pos < a.nEval;
    const int i = live ? (int)evalOrder[pos] : 0;
    Vec<DIM> e = vzero<DIM>(), eNormal = vzero<DIM>();
    int type = ZB_IN_DOMAIN;
    if (live) {
        e = loadVec<DIM>(a.evalPt + (size_t)DIM*i);
        if (a.evalNormal) eNormal = loadVec<DIM>(a.evalNormal + (size_t)DIM*i);
        type = a.evalType[i];
    }
    // evaluation points on the absorbing boundary take no splats (reverse_walk_solver.h:251)
    const bool accumulate = live && type != ZB_ON_ABSORBING_BOUNDARY;
    const bool evalOnRefl = type == ZB_ON_REFLECTING_BOUNDARY;

    // bounding box of the CTA's evaluation points
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        float v = k < DIM ? e[k < DIM ? k : 0] : 0.0f;
        lo[k] = accumulate ? v : kMaxFloat;
        hi[k] = accumulate ? v : -kMaxFloat;
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) { boxLo[k][threadIdx.x >> 5] = lo[k]; boxHi[k][threadIdx.x >> 5] = hi[k]; }
    }
    __syncthreads();
    for (int k = 0; k < 3; k++) {
        lo[k] = boxLo[k][0]; hi[k] = boxHi[k][0];
        for (int w = 1; w < kRwsThreads/32; w++) { lo[k] = fminf(lo[k], boxLo[k][w]); hi[k] = fmaxf(hi[k], boxHi[k][w]); }
    }
    auto meets = [&](const float4& bl, const float4& bh) {
        return bl.x <= hi[0] && bh.x >= lo[0] && bl.y <= hi[1] && bh.y >= lo[1] && (DIM == 2 || (bl.z <= hi[2] && bh.z >= lo[2]));
    };

    float sums[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (int s = 0; s < nSuper; s++) {
        if (!meets(__ldg(tileLo + nTiles + s), __ldg(tileHi + nTiles + s))) continue;          // uniform over the CTA
        const int tEnd = min(nTiles, (s + 1)*kRwsSuper);
        for (int t = s*kRwsSuper; t < tEnd; t++) {
            if (!meets(__ldg(tileLo + t), __ldg(tileHi + t))) continue;                         // uniform over the CTA
            const long long t0 = (long long)t*kRwsTile;
            const int nt = (int)min((long long)kRwsTile, a.nRecords - t0);
            __syncthreads();
            for (int k = threadIdx.x; k < 3*nt; k += kRwsThreads) tile[k] = __ldg(a.records + 3*t0 + k);
            __syncthreads();
            if (!accumulate) continue;
            for (int j = 0; j < nt; j++) {
                rwsAccumulate<DIM>(a, tile[3*j], tile[3*j + 1], tile[3*j + 2], e, eNormal, evalOnRefl, sums);
            }
        }
    }
    if (accumulate) {
        for (int k = 0; k < 6; k++) a.totals[(size_t)6*i + k] += sums[k];
    }
#endif
}

// ---- after the walks: advance every sample's generator and count the walk steps ------------------------------------
template <int DIM>
__global__ void rwsFinishKernel(SolveArgs a)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    long long steps = 0;
    if (i < a.pts.n) {
        long long begin = a.taskOffset[i], end = a.taskOffset[i + 1];
        for (long long t = begin; t < end; t++) steps += a.recLength[t];
        Pcg32 rng;
        rng.state = a.pointState[i];
        rng.inc = a.pts.rngInc[i];
        rng.advance((uint64_t)(end - begin) + 1u);
        a.pts.rngState[i] = rng.state;
    }
#if defined(__CUDA_ARCH__)
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
    if ((threadIdx.x & 31) == 0 && steps) atomicAdd((unsigned long long *)a.walkSteps, (unsigned long long)steps);
#endif
}

} // namespace zb
