// Walk-on-spheres / walk-on-stars estimator kernels.
//
// Reference: WalkOnSpheres::solve / WalkOnStars::solve run a TBB parallel_for over sample points
// and, per point, nWalks sequential walks that share the point's pcg32 and update its Welford
// statistics in place (walk_on_spheres.h:132-168,229-414; walk_on_stars.h:151-187,346-848).
//
// Here the unit of scheduling is one *walk* (a "task"), not one point:
//   1. countTasks      one thread per point: how many walks it launches, how many stratified numbers
//   2. preparePoints   one thread per point: first-sphere radius (a star-radius query) and the
//                      Latin-hypercube samples of the gradient estimator, drawn from the point's rng
//   3. walkKernel      persistent threads; every lane runs one walk at a time and, when it ends,
//                      pulls the next task from a warp-local chunk that a leader lane refills
//                      with one global atomic (warp-level compaction of finished walks); a walk
//                      ends by writing one fixed-size record
//   4. reducePoints    one thread per point folds its walks' records, in walk order, into the
//                      (n, mean, M2) statistics with the reference's Welford update and running
//                      control variates — atomic-free and deterministic for a fixed seed
#pragma once

#include "greens.cuh"
#include "pde.cuh"
#include "coop.cuh"
#include "queries.cuh"

namespace zb {

enum : int { kStatusSplitOverflow = 2 };
constexpr int kSplitStack = 6;
constexpr int kTaskChunk = 64; // tasks a warp reserves per global atomic

// what the extra value channels of a solve need on the device, kept in DEVICE memory and handed around as one pointer
struct ExtraChannelsView {
    int n;
    long long tasks;
    const zb_pde_desc *pde;
    float *refl, *src, *recSolution, *recFirstSource;
};

struct SolveArgs {
    BvhView absorbing, reflecting;
    int hasAbsorbing, hasReflecting, reflectingKind, estimator;
    float domainMin[3], domainMax[3];
    zb_walk_settings ws;
    zb_pde_desc pde;
    zb_sample_points pts; // device pointers
    zb_statistics stats;  // device pointers
    // per-point scratch
    int *taskCount;
    long long *taskOffset;  // exclusive scan of taskCount, n + 1 entries
    int *stratCount;
    long long *stratOffset; // exclusive scan of stratCount, n + 1 entries
    float *strat;
    unsigned long long *pointState; // the point's rng state after drawing its stratified samples
    // per-task records
    float *recSolution, *recFirstSource, *recGradBoundary, *recGradSource;
    int *recLength, *recFlags;
    unsigned long long *taskCounter;
    int *status;
    long long *walkSteps;
    long long totalTasks;
    int anyGradient;
    // Reverse Walk on Stars (rws.cuh): when rwsRecords is set the walks are reverse walks started at sample points,
    // one per point, and every step appends one splat record (the ball and the weight carried by the walk)
    float4 *rwsRecords;               // 3 float4 per record
    unsigned long long *rwsCount;     // records appended so far
    long long rwsCapacity;
    const float *rwsPdf;              // per sample point
    const float *rwsContribution;     // per sample point
    float rwsNormalOffset;            // normalOffsetForAbsorbingBoundary
    // Extra value channels (zb_solve_channels: the reference's vector-valued T, python/zombie_bindings.cpp:37-44). The walk —
    // geometry, random numbers, throughput, weight window — is channel 0's; every further channel only evaluates its own
    // source / Dirichlet / Robin data at the points the walk samples and keeps its own sums, per task:
    int nExtra;
    const zb_pde_desc *extraPde;      // device, nExtra descriptors
    const zb_statistics *extraStats;  // device, nExtra structs of device pointers
    float *xRefl, *xSrc;              // nExtra x totalTasks: reflecting-boundary / source terms of the task's current walk state
    float *xRecSolution, *xRecFirstSource; // nExtra x totalTasks: the task's record
    const ExtraChannelsView *extra;   // the same as one device-resident struct; NULL for scalar solves (what the walk kernels test)
};

template <int DIM>
struct WalkState {
    Vec<DIM> pt, normal, prevDir;
    float prevDist, throughput, reflContrib, srcContrib;
    int walkLength;
    bool onRefl;
};

// SphereSampler<DIM>::sampleUnitSphereUniform (zombie/core/sampling.h:104-107, 164-170)
ZB_DEV Vec<3> uniformDirection(const float *u, Vec<3> *)
{
    float z = 1.0f - 2.0f*u[0];
    float r = sqrtf(fmaxf(0.0f, 1.0f - z*z));
    float phi = 2.0f*kPiF*u[1];
    float s, c;
    sincosf(phi, &s, &c);
    return make3(r*c, r*s, z);
}
ZB_DEV Vec<2> uniformDirection(const float *u, Vec<2> *)
{
    float phi = 2.0f*kPiF*u[0];
    float s, c;
    sincosf(phi, &s, &c);
    return make2(c, s);
}
template <int DIM>
ZB_DEV Vec<DIM> uniformDirection(Pcg32& rng)
{
    float u[2];
    for (int k = 0; k < DIM - 1; k++) u[k] = rng.nextFloat();
    return uniformDirection(u, (Vec<DIM> *)nullptr);
}

// cosine-weighted hemisphere about +e_{DIM-1} (sampling.h:119-127, 184-214)
ZB_DEV Vec<3> cosineHemisphere(const float *u, Vec<3> *)
{
    float u1 = 2.0f*u[0] - 1.0f, u2 = 2.0f*u[1] - 1.0f;
    float dx = 0.0f, dy = 0.0f;
    if (!(u1 == 0.0f && u2 == 0.0f)) {
        float theta, r;
        if (fabsf(u1) > fabsf(u2)) { r = u1; theta = 0.25f*kPiF*(u2/u1); }
        else { r = u2; theta = 0.5f*kPiF*(1.0f - 0.5f*(u1/u2)); }
        dx = r*cosf(theta); dy = r*sinf(theta);
    }
    float z = sqrtf(fmaxf(0.0f, 1.0f - (dx*dx + dy*dy)));
    return make3(dx, dy, z);
}
ZB_DEV Vec<2> cosineHemisphere(const float *u, Vec<2> *)
{
    float u1 = 2.0f*u[0] - 1.0f;
    return make2(u1, sqrtf(fmaxf(0.0f, 1.0f - u1*u1)));
}
// SphereSampler::transformCoordinates (sampling.h:149-156, 236-247)
ZB_DEV void alignWith(const Vec<3>& n, Vec<3>& d)
{
    float sign = copysignf(1.0f, n[2]);
    float a = -1.0f/(sign + n[2]);
    float b = n[0]*n[1]*a;
    Vec<3> b1 = make3(1.0f + sign*n[0]*n[0]*a, sign*b, -sign*n[0]);
    Vec<3> b2 = make3(b, sign + n[1]*n[1]*a, -n[1]);
    d = b1*d[0] + b2*d[1] + n*d[2];
}
ZB_DEV void alignWith(const Vec<2>& n, Vec<2>& d)
{
    Vec<2> s = make2(n[1], -n[0]);
    d = s*d[0] + n*d[1];
}

// offsetPointAlongDirectionImpl (zombie/core/geometric_queries.h:176-207): integer-ULP ray-origin offset
template <int DIM>
ZB_DEV Vec<DIM> offsetPoint(const Vec<DIM>& p, const Vec<DIM>& n)
{
    Vec<DIM> r;
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        int nOff = (int)(n[k]*256.0f);
        float pOff = zb_int_as_float(zb_float_as_int(p[k]) + (p[k] < 0.0f ? -nOff : nOff));
        r[k] = fabsf(p[k]) < (1.0f/32.0f) ? p[k] + (1.0f/65536.0f)*n[k] : pOff;
    }
    return r;
}

template <int DIM>
ZB_DEV bool outsideDomain(const SolveArgs& a, const Vec<DIM>& x)
{
    bool out = false;
#pragma unroll
    for (int k = 0; k < DIM; k++) out = out || x[k] < a.domainMin[k] || x[k] > a.domainMax[k];
    return out;
}

// GeometricQueries::computeDistToAbsorbingBoundary (+ the closest point, reused at termination)
template <int DIM>
ZB_DEV float distToAbsorbing(const SolveArgs& a, const Vec<DIM>& x, Vec<DIM>& proj, int& projPrim, float r2max = kMaxFloat)
{
    if (!a.hasAbsorbing) {
        // default of GeometricQueries::populate (geometric_queries.h:213-218)
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < DIM; k++) { float m = fminf(a.domainMin[k] - x[k], x[k] - a.domainMax[k]); s += m*m; }
        proj = x;
        projPrim = -1;
        return sqrtf(s);
    }
    ClosestHit<DIM> hit;
    closestPoint<DIM>(a.absorbing, x, r2max, hit, a.status);
    proj = hit.p;
    projPrim = hit.prim;
    return hit.d;
}

// GeometricQueries::computeStarRadiusForReflectingBoundary (fcpw_geometric_queries.h:1250-1291)
template <int DIM>
ZB_DEV float starRadiusQuery(const SolveArgs& a, const Vec<DIM>& x, float minRadius, float maxRadius, float precision, bool flip)
{
    if (!a.hasReflecting) return maxRadius;
    if (minRadius > maxRadius) return maxRadius;
    bool flipNormals = !flip; // "FCPW's internal convention requires normals to be flipped"
    float r2max = maxRadius < kMaxFloat ? maxRadius*maxRadius : kMaxFloat;
    float r2 = starRadius2<DIM>(a.reflecting, x, r2max, flipNormals, precision, a.status);
    if (a.reflectingKind == ZB_REFLECTING_NEUMANN && !(r2 < r2max)) return fmaxf(maxRadius, minRadius);
    return fmaxf(sqrtf(r2), minRadius);
}

// the cases of computeStarRadiusForReflectingBoundary that need no traversal; otherwise returns true and
// the squared search radius to hand to the (cooperative) query
template <int DIM>
ZB_DEV bool starQueryNeeded(const SolveArgs& a, float minRadius, float maxRadius, float& trivial, float& r2max)
{
    trivial = maxRadius;
    if (!a.hasReflecting || minRadius > maxRadius) return false;
    r2max = maxRadius < kMaxFloat ? maxRadius*maxRadius : kMaxFloat;
    return true;
}

// result of the query -> radius returned by the reference's lambda (fcpw_geometric_queries.h:1266-1268, 1288-1290)
ZB_DEV float starQueryResult(const SolveArgs& a, float r2, float r2max, float minRadius, float maxRadius)
{
    if (a.reflectingKind == ZB_REFLECTING_NEUMANN && !(r2 < r2max)) return fmaxf(maxRadius, minRadius);
    return fmaxf(sqrtf(r2), minRadius);
}

// the 1% shrink the walk applies to a queried star radius (walk_on_stars.h:392-398, 589-594)
ZB_DEV float shrinkStarRadius(const SolveArgs& a, float r, float distA)
{
    if (a.hasReflecting && a.ws.epsilonShellForReflectingBoundary <= distA) {
        r = fmaxf(0.99f*r, a.ws.epsilonShellForReflectingBoundary);
    }
    return r;
}

// star radius as the walk uses it (walk_on_stars.h:382-398, 577-596): query + 1% shrink
template <int DIM>
ZB_DEV float shrunkStarRadius(const SolveArgs& a, const Vec<DIM>& x, float distA, bool flip)
{
    float r = starRadiusQuery<DIM>(a, x, a.ws.epsilonShellForReflectingBoundary, distA, a.ws.silhouettePrecision, flip);
    if (a.hasReflecting && a.ws.epsilonShellForReflectingBoundary <= distA) {
        r = fmaxf(0.99f*r, a.ws.epsilonShellForReflectingBoundary);
    }
    return r;
}

// ------------------------------------------------------------------------------------------
// 1. task counts
// ------------------------------------------------------------------------------------------
template <int DIM>
ZB_DEV void countTasksPoint(const SolveArgs& a, int i)
{
    int quantity = a.pts.estimationQuantity[i];
    int type = a.pts.type[i];
    int nWalks = a.pts.nWalks[i];
    int tasks = 0, strat = 0;
    if (a.rwsRecords) {
        // ReverseWalkOnStars::solve (reverse_walk_on_stars.h:89-166): one walk per sample point
        tasks = quantity == ZB_SKIP ? 0 : 1;
    } else if (quantity == ZB_SOLUTION_AND_GRADIENT) {
        int nAnti = 1;
        if (a.ws.useGradientAntitheticVariates) { nWalks = max(1, nWalks/2); nAnti = 2; }
        nWalks = max(0, nWalks);
        tasks = nWalks*nAnti;
        strat = 2*nWalks*(DIM - 1);
    } else if (quantity == ZB_SOLUTION) {
        if (type == ZB_ON_ABSORBING_BOUNDARY) tasks = 0;
        else if (a.pts.distToAbsorbingBoundary[i] <= a.ws.epsilonShellForAbsorbingBoundary) tasks = 1;
        else tasks = max(0, nWalks);
    }
    a.taskCount[i] = tasks;
    a.stratCount[i] = strat;
}

template <int DIM>
__global__ void countTasksKernel(SolveArgs a)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < a.pts.n) countTasksPoint<DIM>(a, i);
}

// ------------------------------------------------------------------------------------------
// 2. per-point preparation
// ------------------------------------------------------------------------------------------
// Everything of the per-point preparation except a possibly needed star-radius query: returns true
// when firstSphereRadius[i] still has to be set from a query at the point (with `flip`).
template <int DIM>
ZB_DEV bool preparePointNoQuery(const SolveArgs& a, int i, bool& flip)
{
    bool needStar = false;
    flip = false;
    int quantity = a.pts.estimationQuantity[i];
    Pcg32 rng;
    rng.state = a.pts.rngState[i];
    rng.inc = a.pts.rngInc[i];
    float distA = a.pts.distToAbsorbingBoundary[i];
    bool wost = a.estimator == ZB_WALK_ON_STARS;

    if (a.rwsRecords) {
        a.pts.firstSphereRadius[i] = 0.0f; // a reverse walk queries its first star like any other (reverse_walk_on_stars.h:283-309)
    } else if (quantity == ZB_SOLUTION_AND_GRADIENT) {
        float boundaryDist = wost ? fminf(distA, a.pts.distToReflectingBoundary[i]) : distA;
        a.pts.firstSphereRadius[i] = 0.99f*boundaryDist;
        // generateStratifiedSamples<DIM-1>(2*nWalks) (sampling.h:436-458): jittered diagonal, then one
        // Fisher-Yates shuffle per dimension
        int nSamples = a.stratCount[i]/(DIM - 1);
        float *s = a.strat + a.stratOffset[i];
        if (nSamples > 0) {
            const float oneMinusEps = 1.0f - kEpsilon;
            float inv = 1.0f/nSamples;
            for (int k = 0; k < nSamples; k++) {
                for (int j = 0; j < DIM - 1; j++) {
                    float sj = (k + rng.nextFloat())*inv;
                    s[(DIM - 1)*k + j] = fminf(sj, oneMinusEps);
                }
            }
            for (int j = 0; j < DIM - 1; j++) {
                for (int k = 0; k < nSamples; k++) {
                    int other = k + (int)rng.nextUInt((uint32_t)(nSamples - k));
                    float t = s[(DIM - 1)*k + j];
                    s[(DIM - 1)*k + j] = s[(DIM - 1)*other + j];
                    s[(DIM - 1)*other + j] = t;
                }
            }
        }
    } else if (quantity == ZB_SOLUTION && a.pts.type[i] != ZB_ON_ABSORBING_BOUNDARY) {
        bool hasPrev = a.stats.nSolutionEstimates[i] > 0;
        if (!hasPrev) {
            float radius;
            if (!wost) {
                radius = distA;
            } else if (distA <= a.ws.epsilonShellForAbsorbingBoundary || a.ws.stepsBeforeUsingMaximalSpheres == 0) {
                radius = distA;
            } else {
                PdeEval<DIM> pde{a.pde};
                if (a.pts.type[i] == ZB_ON_REFLECTING_BOUNDARY && pde.hasNonZeroRobinCoeff(loadVec<DIM>(a.pts.pt + (size_t)DIM*i))) {
                    radius = a.ws.epsilonShellForReflectingBoundary;
                } else {
                    flip = a.ws.solveDoubleSided && a.pts.type[i] == ZB_ON_REFLECTING_BOUNDARY &&
                           a.pts.estimateBoundaryNormalAligned && a.pts.estimateBoundaryNormalAligned[i];
                    needStar = true;
                    radius = distA;
                }
            }
            a.pts.firstSphereRadius[i] = radius;
        }
    }
    a.pointState[i] = rng.state;
    return needStar;
}

// one thread per point, thread-per-query star radius (host probe and debugging)
template <int DIM>
ZB_DEV void preparePoint(const SolveArgs& a, int i)
{
    bool flip;
    if (preparePointNoQuery<DIM>(a, i, flip)) {
        Vec<DIM> x = loadVec<DIM>(a.pts.pt + (size_t)DIM*i);
        a.pts.firstSphereRadius[i] = shrunkStarRadius<DIM>(a, x, a.pts.distToAbsorbingBoundary[i], flip);
    }
}

template <int DIM>
__global__ void __launch_bounds__(128) preparePointsKernel(SolveArgs a)
{
#if defined(__CUDA_ARCH__)
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    bool flip = false, needStar = false;
    if (i < a.pts.n) needStar = preparePointNoQuery<DIM>(a, i, flip);
    float eps = a.ws.epsilonShellForReflectingBoundary;
    float distA = needStar ? a.pts.distToAbsorbingBoundary[i] : 0.0f;
    float trivial = 0.0f, r2max = 0.0f;
    bool query = needStar && starQueryNeeded<DIM>(a, eps, distA, trivial, r2max);
    if (__any_sync(0xffffffffu, query)) {
        Vec<DIM> x = query ? loadVec<DIM>(a.pts.pt + (size_t)DIM*i) : vzero<DIM>();
        float r2 = coopStarRadius2<DIM>(a.reflecting, ws, query, x, r2max, !flip, a.ws.silhouettePrecision, a.status);
        if (query) trivial = starQueryResult(a, r2, r2max, eps, distA);
    }
    if (needStar) a.pts.firstSphereRadius[i] = shrinkStarRadius(a, trivial, distA);
#endif
}

// ------------------------------------------------------------------------------------------
// 3. the walks
// ------------------------------------------------------------------------------------------

// everything a lane keeps between steps
// ---- extra value channels -----------------------------------------------------------------------------------------
// Out of line: a scalar solve (nExtra = 0) never executes them and they stay out of the step kernel's hot instruction footprint.
// They take ONE pointer to a device-resident view: a reference to SolveArgs would force every thread of the calling kernel to
// copy the whole parameter block to local memory just to have an address to pass (measured: +1 KB of stack and -4 % walk
// steps/s), and seven by-value fields cost the step kernel nine registers, i.e. one resident CTA per SM.
// Each mirrors the expression of the scalar path it shadows, so channel c of a multi-channel solve equals a scalar solve of
// that channel's PDE on the same generators up to the contraction of multiply-adds.
template <int DIM>
ZB_DEV_COLD void extraSourceTerm(const ExtraChannelsView *vp, long long task, Vec<DIM> p, float throughput, float gnorm)
{
    const ExtraChannelsView v = *vp;
    for (int c = 0; c < v.n; c++) {
        PdeEval<DIM> pde{v.pde[c]};
        v.src[(size_t)c*v.tasks + task] += throughput*(gnorm*pde.source(p));
    }
}
template <int DIM>
ZB_DEV_COLD void extraRobinTerm(const ExtraChannelsView *vp, long long task, Vec<DIM> p, Vec<DIM> n, bool aligned, float weight, float pdf)
{
    const ExtraChannelsView v = *vp;
    for (int c = 0; c < v.n; c++) {
        PdeEval<DIM> pde{v.pde[c]};
        v.refl[(size_t)c*v.tasks + task] += weight*pde.robin(p, n, aligned)/pdf;
    }
}
// the first ball of a gradient estimate (beginTask): the channel's first source term goes to its record and into its sums
template <int DIM>
ZB_DEV_COLD void extraFirstSource(const ExtraChannelsView *vp, long long task, Vec<DIM> p, float throughput, float gnorm)
{
    const ExtraChannelsView v = *vp;
    for (int c = 0; c < v.n; c++) {
        PdeEval<DIM> pde{v.pde[c]};
        float first = gnorm*pde.source(p);
        v.recFirstSource[(size_t)c*v.tasks + task] = first;
        v.src[(size_t)c*v.tasks + task] += throughput*first;
    }
}
// a walk state ended: fold the channel's terminal contribution and sums into the task's record (not when the walk escaped)
template <int DIM>
ZB_DEV_COLD void extraFinishState(const ExtraChannelsView *vp, long long task, bool keep, bool project, Vec<DIM> p, float throughput, bool aligned)
{
    const ExtraChannelsView v = *vp;
    for (int c = 0; c < v.n; c++) {
        size_t k = (size_t)c*v.tasks + task;
        if (keep) {
            float terminal = 0.0f;
            if (project) { PdeEval<DIM> pde{v.pde[c]}; terminal = pde.dirichlet(p, aligned); }
            v.recSolution[k] += throughput*terminal + v.refl[k] + v.src[k];
        }
        v.refl[k] = 0.0f;
        v.src[k] = 0.0f;
    }
}

template <int DIM>
struct Lane {
    WalkState<DIM> st;
    GreensBall<DIM> g;
    Pcg32 rng;
    Vec<DIM> proj;         // closest absorbing-boundary point of st.pt (valid when distA is)
    float distA;
    float firstSphereRadius;
    float total;           // contribution summed over this walk and its splits
    float firstSource;
    Vec<DIM> gradBoundary, gradSource;
    long long task;
    int point;             // index of the sample point the task belongs to
    int totalLength;
    int splits;
    int nQueued;
    bool firstStep, flip, outerFlip, success, active, gradient;
    bool needClosest; // distA / proj must be recomputed at st.pt before the walk continues
    bool keepDistA;   // ... but only proj is wanted: keep the caller-provided distA (sample point inside the epsilon shell)
    bool projValid;   // proj is a point of the absorbing boundary (the last distance query found one)
    bool needsFinalProjection; // Walk on Spheres: the walk ran out of steps and proj is that of the previous position
    bool dirichletAligned;     // double-sided solves: st.pt lies on the side of the absorbing boundary its normal points to (proj's primitive)
};

template <int DIM>
ZB_DEV int findPoint(const long long *offsets, int n, long long task)
{
    int lo = 0, hi = n; // offsets[lo] <= task < offsets[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (offsets[mid] <= task) lo = mid; else hi = mid;
    }
    return lo;
}

// CH: the solve carries extra value channels (zb_solve_channels); the scalar instantiations contain none of their code
template <int DIM, bool WOST, bool CH = false>
ZB_DEV void beginTask(const SolveArgs& a, Lane<DIM>& L, long long task)
{
    PdeEval<DIM> pde{a.pde};
    int i = findPoint<DIM>(a.taskOffset, a.pts.n, task);
    int t = (int)(task - a.taskOffset[i]);
    Vec<DIM> x = loadVec<DIM>(a.pts.pt + (size_t)DIM*i);
    bool yukawa = a.pde.absorptionCoeff > 0.0f && a.ws.stepsBeforeApplyingTikhonov == 0;
    L.task = task;
    L.point = i;
    L.active = true;
    L.total = 0.0f;
    L.totalLength = 0;
    L.splits = -1;
    L.nQueued = 0;
    L.success = false;
    L.firstSource = 0.0f;
    L.gradBoundary = vzero<DIM>();
    L.gradSource = vzero<DIM>();
    L.g.init(yukawa, a.pde.absorptionCoeff);
    L.st.reflContrib = 0.0f;
    L.st.srcContrib = 0.0f;
    L.st.throughput = 1.0f;
    L.st.walkLength = 0;
    L.firstStep = true;
    L.gradient = !a.rwsRecords && a.pts.estimationQuantity[i] == ZB_SOLUTION_AND_GRADIENT;
    L.needClosest = false;
    L.keepDistA = false;
    L.projValid = false;
    L.needsFinalProjection = false;
    L.dirichletAligned = false;

    if (!L.gradient) {
        // estimateSolution (walk_on_spheres.h:354-372; walk_on_stars.h:553-608)
        L.rng = walkGenerator(a.pointState[i], a.pts.rngInc[i], (uint32_t)t);
        L.st.pt = x;
        L.distA = a.pts.distToAbsorbingBoundary[i];
        L.firstSphereRadius = a.pts.firstSphereRadius[i];
        L.flip = false;
        if (WOST) {
            Vec<DIM> n = loadVec<DIM>(a.pts.normal + (size_t)DIM*i);
            L.st.normal = n;
            L.st.prevDir = n;
            L.st.prevDist = kMaxFloat;
            L.st.onRefl = a.pts.type[i] == ZB_ON_REFLECTING_BOUNDARY;
            if (a.ws.solveDoubleSided && L.st.onRefl && a.pts.estimateBoundaryNormalAligned &&
                a.pts.estimateBoundaryNormalAligned[i]) {
                if (a.rwsRecords) {
                    // a reverse walk keeps the sample's normal and only reverses the direction of approach; its first
                    // step then flips the normal and queries the star with flipNormalOrientation = true
                    // (reverse_walk_on_stars.h:127-139, 279-286)
                    L.st.prevDir = -n;
                } else {
                    L.st.normal = -n;
                    L.st.prevDir = -n;
                    L.flip = true;
                }
            }
        } else {
            L.st.normal = vzero<DIM>();
            L.st.prevDir = vzero<DIM>();
            L.st.prevDist = 0.0f;
            L.st.onRefl = false;
        }
        L.outerFlip = L.flip;
        L.proj = x;
        if (!(L.distA > a.ws.epsilonShellForAbsorbingBoundary)) {
            // the walk loop will not run: the terminal projection still needs the closest point
            L.needClosest = true;
            L.keepDistA = true;
        }
        return;
    }

    // estimateSolutionAndGradient (walk_on_spheres.h:417-520; walk_on_stars.h:665-783): the first
    // ball is sampled explicitly; antithetic twins share one generator and mirror the samples
    int nAnti = a.ws.useGradientAntitheticVariates ? 2 : 1;
    int w = t/nAnti, anti = t - w*nAnti;
    L.rng = walkGenerator(a.pointState[i], a.pts.rngInc[i], (uint32_t)w);
    const float *strat = a.strat + a.stratOffset[i];
    float R1 = a.pts.firstSphereRadius[i];
    L.g.updateBall(x, R1);
    L.st.pt = x;
    L.st.normal = vzero<DIM>();
    L.st.prevDir = vzero<DIM>();
    L.st.prevDist = 0.0f;
    L.st.onRefl = false;
    L.flip = false;
    L.outerFlip = false;

    if (!a.ws.ignoreSourceContribution) {
        float gnorm = L.g.norm();
        Vec<DIM> sourcePt = x;
        if (a.pde.isSourceConstant) {
            L.firstSource = gnorm*pde.source(x);
        } else {
            Vec<DIM> dir = uniformDirection(strat + (DIM - 1)*(2*w + 0), (Vec<DIM> *)nullptr);
            float sourceRadius, sourcePdf;
            sourcePt = L.g.sampleVolume(dir, L.rng, sourceRadius, sourcePdf);
            if (anti == 1) sourcePt = x - (sourcePt - x);
            L.firstSource = gnorm*pde.source(sourcePt);
            L.gradSource = L.g.gradient(sourceRadius, sourcePt)/(sourcePdf*gnorm);
        }
        L.st.srcContrib += L.st.throughput*L.firstSource;
        if (CH) extraFirstSource<DIM>(a.extra, task, sourcePt, L.st.throughput, gnorm);
    }

    const float *ub = strat + (DIM - 1)*(2*w + 1);
    Vec<DIM> boundaryDir;
    float boundaryPdf;
    if (a.ws.useCosineSamplingForDerivatives) {
        boundaryDir = cosineHemisphere(ub, (Vec<DIM> *)nullptr);
        if (L.rng.nextFloat() < 0.5f) boundaryDir[DIM - 1] *= -1.0f;
        float cosTheta = fabsf(boundaryDir[DIM - 1]);
        boundaryPdf = 0.5f*(DIM == 3 ? cosTheta/kPiF : cosTheta/2.0f);
        Vec<DIM> axis = a.pts.directionForDerivative ? loadVec<DIM>(a.pts.directionForDerivative + (size_t)DIM*i) : vzero<DIM>();
        if (!a.pts.directionForDerivative) axis[0] = 1.0f;
        alignWith(axis, boundaryDir);
    } else {
        boundaryDir = uniformDirection(ub, (Vec<DIM> *)nullptr);
        boundaryPdf = sphereDirectionPdf<DIM>(1.0f);
    }
    Vec<DIM> boundaryPt = L.g.c + boundaryDir*L.g.R;
    if (anti == 1) boundaryPt = x - (boundaryPt - x);
    if (WOST) {
        L.st.prevDist = L.g.R;
        L.st.prevDir = (boundaryPt - x)/L.g.R;
    }
    L.st.pt = boundaryPt;
    L.st.throughput *= L.g.poissonKernel()/boundaryPdf;
    L.gradBoundary = L.g.poissonKernelGradient(boundaryPt)/(boundaryPdf*L.st.throughput);
    L.firstSphereRadius = 0.0f;
    L.distA = kMaxFloat;
    L.needClosest = true;
}

#ifndef ZB_BOUNDED_CLOSEST
#define ZB_BOUNDED_CLOSEST 1 // 0: every distance query searches without a radius bound (A/B builds)
#endif

// Squared search radius of a pending distance-to-absorbing-boundary query. A Walk on Stars step that starts ON a boundary
// with a non-zero Robin coefficient takes an epsilon-sized ball whatever that distance is (walk_on_stars.h:377-381,
// stepStarPrologue below), so all the walk needs to learn is whether it stands inside the absorbing epsilon shell:
// the search stops at that radius ("nothing found" leaves distA at +inf, the projection unused). 17 % of the steps
// of a C3-like solve and a tenth of its closest-point work (tools/host_probe.cu).
//
// Every other query is seeded with what the walk already knows: the closest point found at the previous position is a
// point OF the absorbing boundary, so its distance from the new position bounds the new distance from above and the
// traversal prunes with it from the root (ZB_SEEDED_CLOSEST; same results — the nearest primitive lies within the bound).
#ifndef ZB_SEEDED_CLOSEST
#define ZB_SEEDED_CLOSEST 1
#endif
template <int DIM, bool WOST>
ZB_DEV float closestSearchRadius2(const SolveArgs& a, const Lane<DIM>& L)
{
    float seeded = kMaxFloat;
    if (ZB_SEEDED_CLOSEST && L.projValid && !L.keepDistA) {
        // |pt - proj| cancels when the walk is close to the boundary (coordinates of order 1, distance of order 1e-3): the
        // slack is absolute, a few ulps of the coordinates, on top of a relative one for the primitive test's own rounding
        Vec<DIM> d = L.st.pt - L.proj;
        float scale = 0.0f;
#pragma unroll
        for (int k = 0; k < DIM; k++) scale = fmaxf(scale, fmaxf(fabsf(L.st.pt[k]), fabsf(L.proj[k])));
        float bound = sqrtf(dot(d, d))*1.0001f + 16.0f*kEpsilon*scale;
        seeded = bound*bound;
    }
    if (!WOST || !ZB_BOUNDED_CLOSEST || !a.hasReflecting) return seeded;
    if (L.keepDistA || L.firstStep || !L.st.onRefl) return seeded;
    if (!(a.ws.stepsBeforeUsingMaximalSpheres > L.st.walkLength)) return seeded; // the maximal-sphere step reads distA
    PdeEval<DIM> pde{a.pde};
    if (!pde.hasNonZeroRobinCoeff(L.st.pt)) return seeded;
    float e = a.ws.epsilonShellForAbsorbingBoundary;
    return fminf(seeded, e*e*1.0001f);
}

// returnBoundaryNormalAlignedValue of the terminal contribution (walk_on_spheres.h:300-308, walk_on_stars.h:499-506):
// solveDoubleSided && signedDistance > 0, the sign being that of (x - p) . n for the closest primitive's normal
// (fcpw Interaction::signedDistance, interaction.h:32-34). prim: leaf-order index of the closest primitive, < 0 = none.
template <int DIM>
ZB_DEV bool onNormalAlignedSide(const SolveArgs& a, const Vec<DIM>& x, const Vec<DIM>& p, int prim)
{
    if (!a.ws.solveDoubleSided || prim < 0) return false;
    Prim<DIM> pr;
    pr.load(a.absorbing, (uint32_t)prim);
    return dot(x - p, unitNormal<DIM>(pr)) > 0.0f;
}

// thread-per-query resolution of a pending distance query (host probe; the kernel does it cooperatively)
template <int DIM, bool WOST = true>
ZB_DEV void resolveClosest(const SolveArgs& a, Lane<DIM>& L)
{
    int projPrim;
    float d = distToAbsorbing<DIM>(a, L.st.pt, L.proj, projPrim, closestSearchRadius2<DIM, WOST>(a, L));
    L.projValid = projPrim >= 0;
    L.dirichletAligned = a.hasAbsorbing && onNormalAlignedSide<DIM>(a, L.st.pt, L.proj, projPrim);
    if (!L.keepDistA) L.distA = d;
    L.needClosest = false;
    L.keepDistA = false;
}

// outcome of one step
enum StepResult { kContinue = 0, kReached, kRoulette, kMaxLength, kEscaped };

// applyWeightWindow (walk_on_spheres.h:193-227; walk_on_stars.h:310-343)
template <int DIM>
ZB_DEV bool weightWindow(const SolveArgs& a, Lane<DIM>& L, WalkState<DIM> *queue)
{
    WalkState<DIM>& st = L.st;
    if (st.throughput > a.ws.splittingThreshold) {
        float left = st.throughput - a.ws.splittingThreshold;
        st.throughput = a.ws.splittingThreshold;
        WalkState<DIM> split = st;
        split.reflContrib = 0.0f;
        split.srcContrib = 0.0f;
        split.walkLength = 0;
        int copies = 0;
        while (left > a.ws.splittingThreshold) { left -= a.ws.splittingThreshold; copies++; }
        if (L.rng.nextFloat() < left/a.ws.splittingThreshold) copies++;
        for (int c = 0; c < copies; c++) {
            if (L.nQueued < kSplitStack) queue[L.nQueued++] = split;
            else zb_atomic_or(a.status, kStatusSplitOverflow);
        }
    } else if (st.throughput < a.ws.russianRouletteThreshold) {
        float survival = st.throughput/a.ws.russianRouletteThreshold;
        if (survival < L.rng.nextFloat()) {
            st.throughput = 0.0f;
            return true;
        }
        st.throughput = a.ws.russianRouletteThreshold;
    }
    return false;
}

// Star radius of the coming WoSt step (walk_on_stars.h:359-399). Returns true when a star-radius
// query at st.pt is needed (r2max = its squared search radius); otherwise starRadius is final.
template <int DIM>
ZB_DEV bool stepStarPrologue(const SolveArgs& a, Lane<DIM>& L, float& starRadius, float& r2max)
{
    PdeEval<DIM> pde{a.pde};
    WalkState<DIM>& st = L.st;
    if (L.firstStep && L.firstSphereRadius > 0.0f) {
        starRadius = L.firstSphereRadius;
        return false;
    }
    L.flip = false;
    if (a.ws.solveDoubleSided && st.onRefl) {
        if (st.prevDist > 0.0f && dot(st.prevDir, st.normal) < 0.0f) {
            st.normal = -st.normal;
            L.flip = true;
        }
    }
    if (a.ws.stepsBeforeUsingMaximalSpheres <= st.walkLength) {
        starRadius = L.distA;
        return false;
    }
    if (st.onRefl && pde.hasNonZeroRobinCoeff(st.pt)) {
        starRadius = a.ws.epsilonShellForReflectingBoundary;
        return false;
    }
    if (!starQueryNeeded<DIM>(a, a.ws.epsilonShellForReflectingBoundary, L.distA, starRadius, r2max)) {
        starRadius = shrinkStarRadius(a, starRadius, L.distA);
        return false;
    }
    return true;
}

// One splat record of a reverse walk: the current ball and the weight the walk carries into it
// (rws::splatContribution, zombie/variance_reduction/reverse_walk_solver.h:233-283, evaluation-point independent part):
//   q0 = (centre, R)   q1 = (current normal, v)   q2 = (v * contribution, flags, task, step)
// v = alpha * throughput / pdf (/ normalOffset for samples of the absorbing boundary); flags: bits 0-2 the statistics
// bucket (0 absorbing, 1 absorbing normal-aligned, 2 reflecting, 3 reflecting normal-aligned, 4 source), bit 3 walk is on
// the reflecting boundary, bit 4 the ball uses the Yukawa Green's function.
template <int DIM>
ZB_DEV void rwsEmit(const SolveArgs& a, const Lane<DIM>& L)
{
#if defined(__CUDA_ARCH__)
    unsigned long long slot = atomicAdd(a.rwsCount, 1ull);
#else
    unsigned long long slot = (*a.rwsCount)++;
#endif
    if ((long long)slot >= a.rwsCapacity) return; // counted, not stored: the host sees count > capacity and retries
    int i = L.point;
    int type = a.pts.type[i];
    bool aligned = a.pts.estimateBoundaryNormalAligned && a.pts.estimateBoundaryNormalAligned[i];
    uint32_t bucket = type == ZB_ON_ABSORBING_BOUNDARY ? (aligned ? 1u : 0u) : (type == ZB_ON_REFLECTING_BOUNDARY ? (aligned ? 3u : 2u) : 4u);
    float v = (L.st.onRefl ? 2.0f : 1.0f)*L.st.throughput/a.rwsPdf[i];
    if (type == ZB_ON_ABSORBING_BOUNDARY) v /= a.rwsNormalOffset;
    uint32_t flags = bucket | (L.st.onRefl ? 8u : 0u) | (L.g.yukawa ? 16u : 0u);
    float4 q0, q1, q2;
    q0.x = L.st.pt[0]; q0.y = L.st.pt[1]; q0.z = DIM == 3 ? L.st.pt[DIM - 1] : 0.0f; q0.w = L.g.R;
    q1.x = L.st.normal[0]; q1.y = L.st.normal[1]; q1.z = DIM == 3 ? L.st.normal[DIM - 1] : 0.0f; q1.w = v;
    // q2.zw: (task, step) — the record's identity, which fixes its place in the gather's order whatever the order of the
    // appends above was
    q2.x = v*a.rwsContribution[i]; q2.y = zb_uint_as_float(flags);
    q2.z = zb_uint_as_float((uint32_t)L.task); q2.w = zb_uint_as_float((uint32_t)L.st.walkLength);
    a.rwsRecords[3*slot] = q0;
    a.rwsRecords[3*slot + 1] = q1;
    a.rwsRecords[3*slot + 2] = q2;
}

// The rest of one iteration of the while loop of WalkOnStars::walk (walk_on_stars.h:401-487) given the
// star radius, or the whole iteration of WalkOnSpheres::walk (walk_on_spheres.h:237-288). The distance
// to the absorbing boundary at the new position is left pending (L.needClosest).
template <int DIM, bool WOST, bool CH = false>
ZB_DEV StepResult stepBody(const SolveArgs& a, Lane<DIM>& L, WalkState<DIM> *queue, float starRadius)
{
    PdeEval<DIM> pde{a.pde};
    WalkState<DIM>& st = L.st;

    if (!WOST) {
        L.g.updateBall(st.pt, L.distA);
        if (!a.ws.ignoreSourceContribution) {
            float contribution;
            Vec<DIM> sourcePt = st.pt;
            if (a.pde.isSourceConstant) {
                contribution = L.g.norm()*pde.source(st.pt);
            } else {
                Vec<DIM> sdir = uniformDirection<DIM>(L.rng);
                float r, pdf;
                sourcePt = L.g.sampleVolume(sdir, L.rng, r, pdf);
                contribution = L.g.norm()*pde.source(sourcePt);
            }
            st.srcContrib += st.throughput*contribution;
            if (CH) extraSourceTerm<DIM>(a.extra, L.task, sourcePt, st.throughput, L.g.norm());
        }
        Vec<DIM> dir = uniformDirection<DIM>(L.rng);
        st.pt = st.pt + dir*L.distA;
        st.throughput *= L.g.directionSampledPoissonKernel(st.pt);
        if (outsideDomain<DIM>(a, st.pt)) return kEscaped;
    } else {
        L.g.updateBall(st.pt, starRadius);
        if (a.rwsRecords) rwsEmit<DIM>(a, L); // splatContribution (reverse_walk_on_stars.h:319)

        // --- direction, hemispherical on the reflecting boundary (:410-416)
        Vec<DIM> dir = uniformDirection<DIM>(L.rng);
        if (st.onRefl && dot(st.normal, dir) > 0.0f) dir = -dir;

        // --- first reflecting-boundary hit inside the star (:419-436)
        Vec<DIM> nextPt, nextNormal = vzero<DIM>();
        float nextDist;
        bool hitRefl = false;
        {
            Vec<DIM> origin = st.onRefl ? offsetPoint<DIM>(st.pt, -st.normal) : st.pt;
            if (a.hasReflecting) {
                RayHit<DIM> hit;
                if (intersectRay<DIM, false>(a.reflecting, origin, dir, starRadius, hit, a.status)) {
                    Prim<DIM> prim;
                    prim.load(a.reflecting, (uint32_t)hit.prim);
                    hitRefl = true;
                    nextPt = prim.pointAt(hit.bv, hit.bw);
                    nextNormal = unitNormal<DIM>(prim);
                    nextDist = hit.t;
                }
            }
            if (!hitRefl) {
                nextPt = origin + dir*starRadius;
                nextDist = starRadius;
            }
        }

        // --- Neumann / Robin data inside the star (:190-250)
        if (a.hasReflecting && !a.ws.ignoreReflectingBoundaryContribution) {
            Vec<DIM> rand;
            for (int k = 0; k < DIM; k++) rand[k] = L.rng.nextFloat();
            SphereSample<DIM> bs;
            if (sampleSphere<DIM>(a.reflecting, st.pt, starRadius*starRadius, rand, bs)) {
                Vec<DIM> toSample = bs.p - st.pt;
                float distToSample = norm(toSample);
                float alpha = st.onRefl ? 2.0f : 1.0f;
                bool aligned = false;
                if (a.ws.solveDoubleSided) {
                    toSample = toSample/distToSample;
                    if (L.flip) {
                        aligned = true;
                    } else if (dot(toSample, bs.n) < -a.ws.silhouettePrecision) {
                        bool flipSample = true;
                        if (alpha > 1.0f) flipSample = dot(toSample, st.normal) < -a.ws.silhouettePrecision;
                        if (flipSample) aligned = true;
                    }
                }
                Vec<DIM> sampleNormal = aligned ? -bs.n : bs.n;
                if (bs.pdf > 0.0f && distToSample < starRadius) {
                    // intersectsWithReflectingBoundary (fcpw_geometric_queries.h:1218-1226)
                    Vec<DIM> p1 = st.onRefl ? offsetPoint<DIM>(st.pt, -st.normal) : st.pt;
                    Vec<DIM> p2 = offsetPoint<DIM>(bs.p, -sampleNormal);
                    Vec<DIM> d12 = p2 - p1;
                    float len = norm(d12);
                    RayHit<DIM> occ;
                    bool blocked = intersectRay<DIM, true>(a.reflecting, p1, d12/len, len, occ, a.status);
                    if (!blocked) {
                        float r = fmaxf(L.g.rClamp, distToSample);
                        float G = L.g.evaluate(r);
                        float h = pde.robin(bs.p, bs.n, aligned); // returnBoundaryNormalAlignedValue (walk_on_stars.h:242-244)
                        st.reflContrib += st.throughput*alpha*G*h/bs.pdf;
                        if (CH) extraRobinTerm<DIM>(a.extra, L.task, bs.p, bs.n, aligned, st.throughput*alpha*G, bs.pdf);
                    }
                }
            }
        }

        // --- source inside the star, sampled along the step direction (:253-275)
        if (!a.ws.ignoreSourceContribution) {
            float sourceRadius, sourcePdf;
            Vec<DIM> sourcePt = L.g.sampleVolume(dir, L.rng, sourceRadius, sourcePdf);
            if (sourceRadius <= nextDist) {
                st.srcContrib += st.throughput*(L.g.norm()*pde.source(sourcePt));
                if (CH) extraSourceTerm<DIM>(a.extra, L.task, sourcePt, st.throughput, L.g.norm());
            }
        }

        // --- move (:446-452) and throughput (:278-307)
        st.prevDist = nextDist;
        st.prevDir = dir;
        st.pt = nextPt;
        st.normal = nextNormal;
        st.onRefl = hitRefl;
        float factor;
        if (st.onRefl && st.prevDist > kEpsilon) {
            float mu = 0.0f;
            Vec<DIM> n = st.normal;
            if (!a.pde.areRobinConditionsPureNeumann) {
                if (a.ws.solveDoubleSided && dot(st.prevDir, st.normal) < 0.0f) n = -n;
                mu = pde.robinCoeff(st.pt, st.normal);
            }
            float reflectance = L.g.reflectance(st.prevDist, st.prevDir, n, mu);
            float maxReflectance = a.pde.areRobinCoeffsNonnegative ? 1.0f : 2.0f;
            factor = reflectance < 0.0f ? 0.0f : (maxReflectance < reflectance ? maxReflectance : reflectance);
        } else {
            factor = L.g.directionSampledPoissonKernel(st.pt);
        }
        st.throughput *= factor;
        if (!st.onRefl && outsideDomain<DIM>(a, st.pt)) return kEscaped;
    }

    if (weightWindow<DIM>(a, L, queue)) return kRoulette;
    st.walkLength++;
    if (st.walkLength >= a.ws.maxWalkLength) { L.needsFinalProjection = !WOST; return kMaxLength; }
    if (a.pde.absorptionCoeff > 0.0f && a.ws.stepsBeforeApplyingTikhonov == st.walkLength) {
        L.g.init(true, a.pde.absorptionCoeff);
    }
    L.needClosest = true;
    L.firstStep = false;
    return kContinue;
}

// thread-per-query version of one whole step (host probe)
template <int DIM, bool WOST, bool CH = false>
ZB_DEV StepResult walkStep(const SolveArgs& a, Lane<DIM>& L, WalkState<DIM> *queue)
{
    if (L.needClosest) resolveClosest<DIM, WOST>(a, L);
    if (!(L.distA > a.ws.epsilonShellForAbsorbingBoundary)) return kReached;
    float starRadius = 0.0f, r2max = 0.0f;
    if (WOST && stepStarPrologue<DIM>(a, L, starRadius, r2max)) {
        starRadius = shrunkStarRadius<DIM>(a, L.st.pt, L.distA, L.flip);
    }
    return stepBody<DIM, WOST, CH>(a, L, queue, starRadius);
}



// a walk state finished: fold it into the task total and start a queued split or emit the record
template <int DIM, bool WOST, bool CH = false>
ZB_DEV void finishState(const SolveArgs& a, Lane<DIM>& L, StepResult code, WalkState<DIM> *queue)
{
    PdeEval<DIM> pde{a.pde};
    L.splits++;
    if (code != kEscaped) {
        // getTerminalContribution (walk_on_spheres.h:294-318; walk_on_stars.h:493-516): the projection
        // onto the absorbing boundary was already computed by the last distance query
        float terminal = 0.0f;
        bool project = WOST ? code == kReached : (code == kReached || code == kMaxLength);
        if (project && !a.ws.ignoreAbsorbingBoundaryContribution) {
            // a Walk on Spheres walk that ran out of steps projects its LAST position (walk_on_spheres.h:294-308); the step
            // that moved it there left the distance query pending (the wavefront defers the finish by one round instead)
            if (!WOST && code == kMaxLength && L.needsFinalProjection) resolveClosest<DIM, WOST>(a, L);
            terminal = pde.dirichlet(a.hasAbsorbing ? L.proj : L.st.pt, L.dirichletAligned);
        }
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
    if (a.anyGradient) {
        a.recFirstSource[t] = L.firstSource;
        for (int k = 0; k < DIM; k++) {
            a.recGradBoundary[(size_t)DIM*t + k] = L.gradBoundary[k];
            a.recGradSource[(size_t)DIM*t + k] = L.gradSource[k];
        }
    }
    L.active = false;
}

#ifndef ZB_WALK_MIN_BLOCKS
#define ZB_WALK_MIN_BLOCKS 4
#endif
template <int DIM, bool WOST>
__global__ void __launch_bounds__(128, ZB_WALK_MIN_BLOCKS) walkKernel(SolveArgs a)
{
#if defined(__CUDA_ARCH__)
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    Lane<DIM> L;
    L.active = false;
    L.needClosest = false;
    L.keepDistA = false;
    L.needsFinalProjection = false;
    WalkState<DIM> queue[kSplitStack];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    long long chunkNext = 0, chunkEnd = 0; // warp-uniform
    bool exhausted = false;                // warp-uniform: the global queue is empty

    while (true) {
        // ---- warp-level compaction: lanes whose walk ended take the next tasks of the warp's chunk
        unsigned need = __ballot_sync(full, !L.active);
        if (need) {
            int want = __popc(need);
            if (!exhausted && chunkEnd - chunkNext < want) {
                // keep what is left of the old chunk and append a fresh one when it is contiguous;
                // otherwise drain the old chunk first
                if (chunkNext == chunkEnd) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(a.taskCounter, (unsigned long long)kTaskChunk);
                    base = __shfl_sync(full, base, 0);
                    chunkNext = (long long)base;
                    chunkEnd = min((long long)base + kTaskChunk, a.totalTasks);
                    if (chunkNext >= a.totalTasks) { exhausted = true; chunkNext = chunkEnd = 0; }
                }
            }
            long long avail = chunkEnd - chunkNext;
            int rank = __popc(need & ((1u << lane) - 1u));
            if (!L.active && rank < avail) beginTask<DIM, WOST>(a, L, chunkNext + rank);
            chunkNext += min((long long)want, avail);
        }
        if (!__any_sync(full, L.active)) {
            if (exhausted) break;
            continue;
        }

        // ---- pending distance-to-absorbing-boundary queries of the warp, cooperatively
        bool qc = L.active && L.needClosest;
        if (__any_sync(full, qc)) {
            if (a.hasAbsorbing) {
                ClosestHit<DIM> hit;
                coopClosestPoint<DIM>(a.absorbing, ws, qc, L.st.pt, closestSearchRadius2<DIM, WOST>(a, L), hit, a.status);
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

        // ---- walks that arrived inside the epsilon shell (loop condition of walk())
        if (L.active && !L.needClosest && !(L.distA > a.ws.epsilonShellForAbsorbingBoundary)) {
            finishState<DIM, WOST>(a, L, kReached, queue);
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
            StepResult r = stepBody<DIM, WOST>(a, L, queue, starRadius);
            if (r != kContinue) finishState<DIM, WOST>(a, L, r, queue);
        }
    }
#endif
}

// ------------------------------------------------------------------------------------------
// 4. per-point reduction (SampleStatistics::update / add* — common.h:201-395)
// ------------------------------------------------------------------------------------------
ZB_DEV void welford(float estimate, float& mean, float& m2, int n)
{
    float delta = estimate - mean;
    mean += delta/n;
    float delta2 = estimate - mean;
    m2 += delta*delta2;
}

// One channel's fold: channel 0 (primary) reads the solve's own PDE, statistics and records; an extra channel (zb_solve_channels)
// its own sums over the same walks.
template <int DIM>
ZB_DEV long long reduceChannel(const SolveArgs& a, int i, const zb_pde_desc& pdeDesc, const zb_statistics& stats, const float *recSolution,
                               const float *recFirstSource, bool primary)
{
    int quantity = a.pts.estimationQuantity[i];
    PdeEval<DIM> pde{pdeDesc};

    float mean = stats.solutionMean[i], m2 = stats.solutionM2[i];
    int nSol = stats.nSolutionEstimates[i], nGrad = stats.nGradientEstimates[i];
    int walkLength = stats.totalWalkLength[i], splits = stats.totalSplits[i];
    float firstSourceTotal = stats.totalFirstSourceContribution[i];
    float derivativeTotal = stats.totalDerivativeContribution[i];
    float gmean[DIM], gm2[DIM];
    for (int k = 0; k < DIM; k++) { gmean[k] = stats.gradientMean[(size_t)DIM*i + k]; gm2[k] = stats.gradientM2[(size_t)DIM*i + k]; }

    if (quantity == ZB_SOLUTION && a.pts.type[i] == ZB_ON_ABSORBING_BOUNDARY) {
        if (nSol == 0) {
            float value = 0.0f;
            // walk_on_spheres.h:331-337, walk_on_stars.h:529-535: the sample's own side flag
            bool aligned = a.ws.solveDoubleSided && a.pts.estimateBoundaryNormalAligned && a.pts.estimateBoundaryNormalAligned[i];
            if (!a.ws.ignoreAbsorbingBoundaryContribution) value = pde.dirichlet(loadVec<DIM>(a.pts.pt + (size_t)DIM*i), aligned);
            nSol += 1;
            welford(value, mean, m2, nSol);
            if (primary) a.pts.firstSphereRadius[i] = 0.0f;
        }
    } else {
        long long begin = a.taskOffset[i], end = a.taskOffset[i + 1];
        bool gradient = quantity == ZB_SOLUTION_AND_GRADIENT;
        int nAnti = gradient && a.ws.useGradientAntitheticVariates ? 2 : 1;
        Vec<DIM> ddir = vzero<DIM>();
        if (gradient) {
            if (a.pts.directionForDerivative) ddir = loadVec<DIM>(a.pts.directionForDerivative + (size_t)DIM*i);
            else ddir[0] = 1.0f;
        }
        float cvBoundary = 0.0f, cvSource = 0.0f;
        for (long long t = begin; t < end; t++) {
            if (gradient && (t - begin)%nAnti == 0 && a.ws.useGradientControlVariates) {
                // control variates are the running means before this antithetic pair (walk_on_stars.h:693-699)
                cvBoundary = mean;
                cvSource = firstSourceTotal/max(1, nSol);
            }
            int flags = a.recFlags[t];
            walkLength += a.recLength[t];
            if (!(flags & (int)0x80000000)) continue;
            float total = recSolution[t];
            if (gradient) {
                float firstSource = recFirstSource[t];
                float boundaryContribution = total - firstSource;
                float derivative = 0.0f;
                float est[DIM];
                for (int k = 0; k < DIM; k++) {
                    float gb = (boundaryContribution - cvBoundary)*a.recGradBoundary[(size_t)DIM*t + k];
                    float gs = (firstSource - cvSource)*a.recGradSource[(size_t)DIM*t + k];
                    derivative += gb*ddir[k];
                    derivative += gs*ddir[k];
                    est[k] = gb + gs;
                }
                nSol += 1;
                welford(total, mean, m2, nSol);
                firstSourceTotal += firstSource;
                nGrad += 1;
                for (int k = 0; k < DIM; k++) welford(est[k], gmean[k], gm2[k], nGrad);
                derivativeTotal += derivative;
            } else {
                nSol += 1;
                welford(total, mean, m2, nSol);
            }
            splits += flags & 0x7fffffff;
        }
        // the point's generator moves on, so a later solve() on the same points draws fresh walks
        if (primary) {
            Pcg32 rng;
            rng.state = a.pointState[i];
            rng.inc = a.pts.rngInc[i];
            rng.advance((uint64_t)(end - begin) + 1u);
            a.pts.rngState[i] = rng.state;
        }
    }

    long long stepsTaken = (long long)walkLength - (long long)stats.totalWalkLength[i];
    stats.solutionMean[i] = mean;
    stats.solutionM2[i] = m2;
    stats.nSolutionEstimates[i] = nSol;
    stats.nGradientEstimates[i] = nGrad;
    stats.totalWalkLength[i] = walkLength;
    stats.totalSplits[i] = splits;
    stats.totalFirstSourceContribution[i] = firstSourceTotal;
    stats.totalDerivativeContribution[i] = derivativeTotal;
    for (int k = 0; k < DIM; k++) { stats.gradientMean[(size_t)DIM*i + k] = gmean[k]; stats.gradientM2[(size_t)DIM*i + k] = gm2[k]; }
    return stepsTaken;
}

template <int DIM>
ZB_DEV long long reducePoint(const SolveArgs& a, int i)
{
    return reduceChannel<DIM>(a, i, a.pde, a.stats, a.recSolution, a.recFirstSource, true);
}

template <int DIM>
__global__ void reducePointsKernel(SolveArgs a)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    long long stepsTaken = 0; // walk steps of this launch = the metric's numerator (sum of walkLength++)
    if (i < a.pts.n && a.pts.estimationQuantity[i] != ZB_SKIP) stepsTaken = reducePoint<DIM>(a, i);
#if defined(__CUDA_ARCH__)
    for (int o = 16; o > 0; o >>= 1) stepsTaken += __shfl_down_sync(0xffffffffu, stepsTaken, o);
    if ((threadIdx.x & 31) == 0 && stepsTaken) atomicAdd((unsigned long long *)a.walkSteps, (unsigned long long)stepsTaken);
#endif
}

// extra value channels: the same fold over the same walks, one thread per (channel, point)
template <int DIM>
__global__ void reduceExtraChannelsKernel(SolveArgs a)
{
    long long k = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= (long long)a.pts.n*a.nExtra) return;
    int c = (int)(k/a.pts.n), i = (int)(k - (long long)c*a.pts.n);
    if (a.pts.estimationQuantity[i] == ZB_SKIP) return;
    reduceChannel<DIM>(a, i, a.extraPde[c], a.extraStats[c], a.xRecSolution + (size_t)c*a.totalTasks,
                       a.anyGradient ? a.xRecFirstSource + (size_t)c*a.totalTasks : nullptr, false);
}

} // namespace zb
