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
//      additions per evaluation point is fixed by the record order.
// Round 1 streams every record past every evaluation point (culling by Morton-sorted record tiles is the next step).
#pragma once

#include "walk.cuh"

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

template <int DIM>
__global__ void __launch_bounds__(kRwsThreads) rwsGatherKernel(RwsGatherArgs a)
{
    __shared__ float4 tile[kRwsTile*3];
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const bool live = i < a.nEval;
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
    float sums[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (long long t0 = 0; t0 < a.nRecords; t0 += kRwsTile) {
        const int nt = (int)min((long long)kRwsTile, a.nRecords - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < 3*nt; k += kRwsThreads) tile[k] = __ldg(a.records + 3*t0 + k);
        __syncthreads();
        if (!accumulate) continue;
        for (int j = 0; j < nt; j++) {
            rwsAccumulate<DIM>(a, tile[3*j], tile[3*j + 1], tile[3*j + 2], e, eNormal, evalOnRefl, sums);
        }
    }
    if (accumulate) {
        for (int k = 0; k < 6; k++) a.totals[(size_t)6*i + k] += sums[k];
    }
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
