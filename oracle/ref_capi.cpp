// TEST INFRASTRUCTURE — never linked into, imported by, or called from the product.
//
// A C wrapper around the UNMODIFIED reference headers under /root/reference (zombie + FCPW +
// pcg32 + oneTBB), compiled by oracle/Makefile into oracle/_ref/libzombie_ref.so with
// oracle/eigen_shim standing in for the absent Eigen submodule. It exposes:
//   * FCPW's geometric queries exactly as zombie's GeometricQueries lambdas call them
//     (zombie/utils/fcpw_geometric_queries.h:1063-1291) plus the primitive index,
//   * the reference's brute-force Baseline aggregate as a second opinion,
//   * WalkOnSpheres / WalkOnStars::solve with a PDE whose std::function callbacks are built
//     from the same closed-form families the device evaluates (include/zombie_b200/analytic.h).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.

#include <zombie/point_estimation/walk_on_spheres.h>
#include <zombie/point_estimation/walk_on_stars.h>
#include <zombie/utils/fcpw_geometric_queries.h>
#include <zombie/variance_reduction/kelvin_transform.h>
#include <zombie/variance_reduction/boundary_value_caching.h>
#include <zombie/variance_reduction/reverse_walk_solver.h>
#include <zombie/utils/nearest_neighbor_finder.h>
#include "oneapi/tbb/global_control.h"
#include "oneapi/tbb/task_arena.h"

#include <chrono>
#include <cstring>
#include <thread>

#include "../include/zombie_b200.h"
#include "../include/zombie_b200/analytic.h"

using namespace zombie;

namespace {

template <size_t DIM> using Vec = Eigen::Matrix<float, DIM, 1>;
template <size_t DIM> using Veci = Eigen::Matrix<int, DIM, 1>;

// read/write access to the protected accumulators of SampleStatistics (common.h:389-394)
template <size_t DIM>
struct StatsAccess: public SampleStatistics<float, DIM> {
    using Base = SampleStatistics<float, DIM>;
    static float& mean(Base& s) { return s.*(&StatsAccess::solutionMean); }
    static float& m2(Base& s) { return s.*(&StatsAccess::solutionM2); }
    static float *gmean(Base& s) { return s.*(&StatsAccess::gradientMean); }
    static float *gm2(Base& s) { return s.*(&StatsAccess::gradientM2); }
    static float& firstSource(Base& s) { return s.*(&StatsAccess::totalFirstSourceContribution); }
    static float& derivative(Base& s) { return s.*(&StatsAccess::totalDerivativeContribution); }
    static int& nSol(Base& s) { return s.*(&StatsAccess::nSolutionEstimates); }
    static int& nGrad(Base& s) { return s.*(&StatsAccess::nGradientEstimates); }
    static int& walkLen(Base& s) { return s.*(&StatsAccess::totalWalkLength); }
    static int& splits(Base& s) { return s.*(&StatsAccess::totalSplits); }
};

template <size_t DIM>
Vec<DIM> toVec(const float *p) { Vec<DIM> v; for (size_t i = 0; i < DIM; i++) v[i] = p[i]; return v; }
template <size_t DIM>
void fromVec(const Vec<DIM>& v, float *p) { for (size_t i = 0; i < DIM; i++) p[i] = v[i]; }

template <size_t DIM>
struct RefScene {
    std::vector<Vec<DIM>> aPos, rPos;
    std::vector<Veci<DIM>> aIdx, rIdx;
    FcpwDirichletBoundaryHandler<DIM> dirichlet;
    FcpwNeumannBoundaryHandler<DIM> neumann;
    FcpwRobinBoundaryHandler<DIM> robin;
    std::unique_ptr<GeometricQueries<DIM>> queries;
    int reflectingKind = 0;
    bool useBaseline = false;

    fcpw::Aggregate<DIM> *absorbing() { return aPos.empty() ? nullptr : dirichlet.scene.getSceneData()->aggregate.get(); }
    fcpw::Aggregate<DIM> *reflectingFcpw() {
        if (reflectingKind == ZB_REFLECTING_NEUMANN) return neumann.scene.getSceneData()->aggregate.get();
        if (reflectingKind == ZB_REFLECTING_ROBIN) {
            if (robin.bvh) return robin.bvh.get();
            if (robin.baseline) return robin.baseline.get();
        }
        return nullptr;
    }
    fcpw::Aggregate<DIM> *aggregate(int which) { return which == ZB_ABSORBING ? absorbing() : reflectingFcpw(); }
};

struct AnyScene {
    int dim;
    std::unique_ptr<RefScene<2>> s2;
    std::unique_ptr<RefScene<3>> s3;
};

template <size_t DIM>
std::unique_ptr<RefScene<DIM>> buildScene(const float *aP, int nAV, const int32_t *aI, int nAT,
                                          const float *rP, int nRV, const int32_t *rI, int nRT,
                                          int reflectingKind, const float *minRobin, const float *maxRobin,
                                          int solveDoubleSided, float silhouettePrecision,
                                          int watertight, const float *dMin, const float *dMax, int useBaseline)
{
    std::unique_ptr<RefScene<DIM>> s(new RefScene<DIM>());
    s->reflectingKind = nRT > 0 ? reflectingKind : 0;
    s->useBaseline = useBaseline != 0;
    for (int i = 0; i < nAV; i++) s->aPos.push_back(toVec<DIM>(aP + DIM*i));
    for (int i = 0; i < nRV; i++) s->rPos.push_back(toVec<DIM>(rP + DIM*i));
    for (int i = 0; i < nAT; i++) { Veci<DIM> t; for (size_t k = 0; k < DIM; k++) t[k] = aI[DIM*i + k]; s->aIdx.push_back(t); }
    for (int i = 0; i < nRT; i++) { Veci<DIM> t; for (size_t k = 0; k < DIM; k++) t[k] = rI[DIM*i + k]; s->rIdx.push_back(t); }

    s->queries.reset(new GeometricQueries<DIM>(watertight != 0, toVec<DIM>(dMin), toVec<DIM>(dMax)));
    bool buildBvh = !useBaseline;
    if (nAT > 0) {
        s->dirichlet.buildAccelerationStructure(s->aPos, s->aIdx, buildBvh, false, false);
        populateGeometricQueriesForDirichletBoundary<DIM>(s->dirichlet, *s->queries);
    }
    if (s->reflectingKind == ZB_REFLECTING_NEUMANN) {
        s->neumann.buildAccelerationStructure(s->rPos, s->rIdx,
            getIgnoreCandidateSilhouetteCallback(solveDoubleSided != 0, silhouettePrecision), buildBvh, false, false);
        populateGeometricQueriesForNeumannBoundary<DIM>(s->neumann, getBranchTraversalWeightCallback(), *s->queries);

    } else if (s->reflectingKind == ZB_REFLECTING_ROBIN) {
        std::vector<float> mn(minRobin, minRobin + nRT), mx(maxRobin, maxRobin + nRT);
        s->robin.buildAccelerationStructure(s->rPos, s->rIdx,
            getIgnoreCandidateSilhouetteCallback(solveDoubleSided != 0, silhouettePrecision), mn, mx, buildBvh, false, false);
        populateGeometricQueriesForRobinBoundary<DIM>(s->robin, getBranchTraversalWeightCallback(), *s->queries);
    }
    return s;
}

// PDE callbacks from the shared closed-form families
template <size_t DIM>
void buildPde(const zb_pde_desc& d, PDE<float, DIM>& pde)
{
    pde.absorptionCoeff = d.absorptionCoeff;
    pde.isSourceConstant = d.isSourceConstant != 0;
    pde.areRobinConditionsPureNeumann = d.areRobinConditionsPureNeumann != 0;
    pde.areRobinCoeffsNonnegative = d.areRobinCoeffsNonnegative != 0;
    zb_pde_desc desc = d;
    if (d.mode == ZB_PDE_MANUFACTURED) {
        pde.source = [desc](const Vec<DIM>& x) -> float {
            float xx[3] = {0, 0, 0}; for (size_t i = 0; i < DIM; i++) xx[i] = x[i];
            zb_jet j = zb_eval_solution(desc.family, desc.params, xx);
            return desc.absorptionCoeff*j.u - j.lap;
        };
        pde.dirichlet = [desc](const Vec<DIM>& x, bool) -> float {
            float xx[3] = {0, 0, 0}; for (size_t i = 0; i < DIM; i++) xx[i] = x[i];
            return zb_eval_solution(desc.family, desc.params, xx).u;
        };
        pde.robin = [desc](const Vec<DIM>& x, const Vec<DIM>& n, bool) -> float {
            float xx[3] = {0, 0, 0}; for (size_t i = 0; i < DIM; i++) xx[i] = x[i];
            zb_jet j = zb_eval_solution(desc.family, desc.params, xx);
            float dn = 0.0f; for (size_t i = 0; i < DIM; i++) dn += n[i]*j.g[i];
            return dn + desc.robinCoeff*j.u;
        };
    } else {
        pde.source = [desc](const Vec<DIM>&) -> float { return desc.sourceValue; };
        pde.dirichlet = [desc](const Vec<DIM>&, bool aligned) -> float {
            return aligned && desc.hasDirichletAlignedValue ? desc.dirichletAlignedValue : desc.dirichletValue;
        };
        pde.robin = [desc](const Vec<DIM>&, const Vec<DIM>&, bool) -> float { return desc.robinValue; };
    }
    pde.robinCoeff = [desc](const Vec<DIM>&, const Vec<DIM>&, bool) -> float { return desc.robinCoeff; };
    pde.hasReflectingBoundaryConditions = [](const Vec<DIM>&) -> bool { return true; };

    // callbacks backed by sampled data (zb_grid_view with HOST pointers here), the way the reference's demo fills a PDE
    // from textures (demo/model_problem.h:229-266); same sampler as the device (zombie_b200/analytic.h: zb_grid_sample)
    auto at = [](const zb_grid_view& g, const Vec<DIM>& x) -> float {
        float xx[3] = {0, 0, 0}; for (size_t i = 0; i < DIM; i++) xx[i] = x[i];
        return zb_grid_sample(g, xx);
    };
    if (d.robinCoeffGrid.data) {
        pde.robinCoeff = [desc, at](const Vec<DIM>& x, const Vec<DIM>&, bool) -> float { return at(desc.robinCoeffGrid, x); };
        if (d.mode == ZB_PDE_MANUFACTURED && !d.robinGrid.data) {
            pde.robin = [desc, at](const Vec<DIM>& x, const Vec<DIM>& n, bool) -> float {
                float xx[3] = {0, 0, 0}; for (size_t i = 0; i < DIM; i++) xx[i] = x[i];
                zb_jet j = zb_eval_solution(desc.family, desc.params, xx);
                float dn = 0.0f; for (size_t i = 0; i < DIM; i++) dn += n[i]*j.g[i];
                return dn + at(desc.robinCoeffGrid, x)*j.u;
            };
        }
    }
    if (d.sourceGrid.data) pde.source = [desc, at](const Vec<DIM>& x) -> float { return at(desc.sourceGrid, x); };
    if (d.dirichletGrid.data) {
        pde.dirichlet = [desc, at](const Vec<DIM>& x, bool aligned) -> float {
            return at(aligned && desc.dirichletAlignedGrid.data ? desc.dirichletAlignedGrid : desc.dirichletGrid, x);
        };
    }
    if (d.robinGrid.data) {
        pde.robin = [desc, at](const Vec<DIM>& x, const Vec<DIM>&, bool aligned) -> float {
            return at(aligned && desc.robinAlignedGrid.data ? desc.robinAlignedGrid : desc.robinGrid, x);
        };
    }
}

WalkSettings toSettings(const zb_walk_settings& w)
{
    return WalkSettings(w.epsilonShellForAbsorbingBoundary, w.epsilonShellForReflectingBoundary,
                        w.silhouettePrecision, w.russianRouletteThreshold, w.splittingThreshold,
                        w.maxWalkLength, w.stepsBeforeApplyingTikhonov, w.stepsBeforeUsingMaximalSpheres,
                        w.solveDoubleSided != 0, w.useGradientControlVariates != 0,
                        w.useGradientAntitheticVariates != 0, w.useCosineSamplingForDerivatives != 0,
                        w.ignoreAbsorbingBoundaryContribution != 0, w.ignoreReflectingBoundaryContribution != 0,
                        w.ignoreSourceContribution != 0, false);
}

template <size_t DIM, typename F>
void forEach(int n, int nThreads, F f)
{
    if (nThreads <= 1) { for (int i = 0; i < n; i++) f(i); return; }
    tbb::global_control gc(tbb::global_control::max_allowed_parallelism, (size_t)nThreads);
    tbb::parallel_for(tbb::blocked_range<int>(0, n), [&](const tbb::blocked_range<int>& r) {
        for (int i = r.begin(); i < r.end(); ++i) f(i);
    });
}

template <size_t DIM>
int closestPoint(RefScene<DIM>& s, int which, int n, const float *x, const float *maxRadius,
                 int32_t *prim, float *dist, float *pt, float *normal, int nThreads)
{
    fcpw::Aggregate<DIM> *agg = s.aggregate(which);
    if (!agg) return ZB_ERR_INVALID_ARGUMENT;
    forEach<DIM>(n, nThreads, [&](int i) {
        // fcpw_geometric_queries.h:1078-1096 (projectToAbsorbingBoundary), recordNormal as requested
        fcpw::Interaction<DIM> it;
        float r2 = maxRadius && maxRadius[i] < fcpw::maxFloat ? maxRadius[i]*maxRadius[i] : fcpw::maxFloat;
        fcpw::BoundingSphere<DIM> sphere(toVec<DIM>(x + DIM*i), r2);
        bool found = agg->findClosestPoint(sphere, it, normal != nullptr);
        prim[i] = found ? it.primitiveIndex : -1;
        dist[i] = found ? it.d : fcpw::maxFloat;
        if (pt) fromVec<DIM>(it.p, pt + DIM*i);
        if (normal) fromVec<DIM>(it.n, normal + DIM*i);
    });
    return ZB_OK;
}

template <size_t DIM>
int countRayHits(RefScene<DIM>& s, int which, int n, const float *o, const float *d, const float *tMax, int32_t *count, int nThreads)
{
    fcpw::Aggregate<DIM> *agg = s.aggregate(which);
    if (!agg) return ZB_ERR_INVALID_ARGUMENT;
    forEach<DIM>(n, nThreads, [&](int i) {
        // fcpw_geometric_queries.h:1116-1136 (intersectAbsorbingBoundaryAllHits), origin not on the boundary
        fcpw::Ray<DIM> r(toVec<DIM>(o + DIM*i), toVec<DIM>(d + DIM*i), tMax ? tMax[i] : fcpw::maxFloat);
        std::vector<fcpw::Interaction<DIM>> is;
        count[i] = agg->intersect(r, is, false, true);
    });
    return ZB_OK;
}

template <size_t DIM>
int insideDomain(RefScene<DIM>& s, int n, const float *x, uint8_t *inside, int nThreads)
{
    forEach<DIM>(n, nThreads, [&](int i) { inside[i] = s.queries->insideDomain(toVec<DIM>(x + DIM*i)) ? 1 : 0; });
    return ZB_OK;
}

template <size_t DIM>
int intersectRay(RefScene<DIM>& s, int which, int n, const float *o, const float *d, const float *tMax,
                 int32_t *prim, float *dist, float *pt, float *normal, int nThreads)
{
    fcpw::Aggregate<DIM> *agg = s.aggregate(which);
    if (!agg) return ZB_ERR_INVALID_ARGUMENT;
    forEach<DIM>(n, nThreads, [&](int i) {
        // fcpw_geometric_queries.h:1178-1196 without the on-boundary origin offset
        fcpw::Ray<DIM> r(toVec<DIM>(o + DIM*i), toVec<DIM>(d + DIM*i), tMax ? tMax[i] : fcpw::maxFloat);
        fcpw::Interaction<DIM> it;
        bool hit = agg->intersect(r, it, false);
        prim[i] = hit ? it.primitiveIndex : -1;
        dist[i] = hit ? it.d : fcpw::maxFloat;
        if (pt) fromVec<DIM>(it.p, pt + DIM*i);
        if (normal) fromVec<DIM>(it.n, normal + DIM*i);
    });
    return ZB_OK;
}

template <size_t DIM>
int occluded(RefScene<DIM>& s, int which, int n, const float *a, const float *b, uint8_t *out, int nThreads)
{
    fcpw::Aggregate<DIM> *agg = s.aggregate(which);
    if (!agg) return ZB_ERR_INVALID_ARGUMENT;
    forEach<DIM>(n, nThreads, [&](int i) {
        out[i] = agg->hasLineOfSight(toVec<DIM>(a + DIM*i), toVec<DIM>(b + DIM*i)) ? 0 : 1;
    });
    return ZB_OK;
}

template <size_t DIM>
int sampleSphere(RefScene<DIM>& s, int n, const float *x, const float *radius, const float *randNums,
                 int32_t *prim, float *pt, float *normal, float *pdf, int nThreads)
{
    if (!s.queries->hasNonEmptyReflectingBoundary) return ZB_ERR_INVALID_ARGUMENT;
    fcpw::Aggregate<DIM> *agg = s.reflectingFcpw();
    std::function<float(float)> weight = getBranchTraversalWeightCallback();
    forEach<DIM>(n, nThreads, [&](int i) {
        // fcpw_geometric_queries.h:1227-1242
        fcpw::BoundingSphere<DIM> sphere(toVec<DIM>(x + DIM*i), radius[i]*radius[i]);
        fcpw::Interaction<DIM> it;
        Vec<DIM> u = toVec<DIM>(randNums + DIM*i);
        int hits = agg->intersect(sphere, it, u, weight);
        prim[i] = hits > 0 ? it.primitiveIndex : -1;
        if (pt) fromVec<DIM>(it.p, pt + DIM*i);
        if (normal) fromVec<DIM>(it.n, normal + DIM*i);
        pdf[i] = hits > 0 ? it.d : 0.0f;
    });
    return ZB_OK;
}

template <size_t DIM>
int starRadius(RefScene<DIM>& s, int n, const float *x, const float *minR, const float *maxR,
               float precision, const uint8_t *flip, float *out, int nThreads)
{
    forEach<DIM>(n, nThreads, [&](int i) {
        out[i] = s.queries->computeStarRadiusForReflectingBoundary(toVec<DIM>(x + DIM*i), minR[i], maxR[i],
                                                                   precision, flip ? flip[i] != 0 : false);
    });
    return ZB_OK;
}

template <size_t DIM>
int solve(RefScene<DIM>& s, int estimator, const zb_pde_desc *pdeDesc, const zb_walk_settings *ws,
          const zb_sample_points *pts, zb_statistics *st, int64_t *walkSteps, int nThreads, double *seconds)
{
    PDE<float, DIM> pde;
    buildPde<DIM>(*pdeDesc, pde);
    // exterior problems: pdeDesc describes the PDE on the exterior domain, the walk runs on the inverted domain
    // with the callbacks wrapped by KelvinTransform::transformPde (kelvin_transform.h:282-348)
    float zeroOrigin[3] = {0.0f, 0.0f, 0.0f};
    KelvinTransform<float, DIM> kelvin(toVec<DIM>(pdeDesc->kelvin ? pdeDesc->kelvinOrigin : zeroOrigin));
    if (pdeDesc->kelvin) {
        PDE<float, DIM> exterior = pde;
        kelvin.transformPde(exterior, pde);
    }
    WalkSettings settings = toSettings(*ws);
    int n = pts->n;
    std::vector<SamplePoint<float, DIM>> samplePts;
    std::vector<SampleStatistics<float, DIM>> stats(n);
    std::vector<int> nWalks(pts->nWalks, pts->nWalks + n);
    samplePts.reserve(n);
    using SA = StatsAccess<DIM>;
    for (int i = 0; i < n; i++) {
        samplePts.emplace_back(toVec<DIM>(pts->pt + DIM*i), toVec<DIM>(pts->normal + DIM*i),
                               (SampleType)pts->type[i], (EstimationQuantity)pts->estimationQuantity[i], 1.0f,
                               pts->distToAbsorbingBoundary[i], pts->distToReflectingBoundary[i]);
        SamplePoint<float, DIM>& sp = samplePts.back();
        // bypass the clock-seeded seedRng() (core/sampling.h:36-41): the field is public (common.h:182)
        sp.rng.state = pts->rngState[i];
        sp.rng.inc = pts->rngInc[i];
        sp.firstSphereRadius = pts->firstSphereRadius[i];
        if (pts->directionForDerivative) sp.directionForDerivative = toVec<DIM>(pts->directionForDerivative + DIM*i);
        if (pts->estimateBoundaryNormalAligned) sp.estimateBoundaryNormalAligned = pts->estimateBoundaryNormalAligned[i] != 0;
        SA::mean(stats[i]) = st->solutionMean[i];
        SA::m2(stats[i]) = st->solutionM2[i];
        for (size_t k = 0; k < DIM; k++) {
            SA::gmean(stats[i])[k] = st->gradientMean[DIM*i + k];
            SA::gm2(stats[i])[k] = st->gradientM2[DIM*i + k];
        }
        SA::firstSource(stats[i]) = st->totalFirstSourceContribution[i];
        SA::derivative(stats[i]) = st->totalDerivativeContribution[i];
        SA::nSol(stats[i]) = st->nSolutionEstimates[i];
        SA::nGrad(stats[i]) = st->nGradientEstimates[i];
        SA::walkLen(stats[i]) = st->totalWalkLength[i];
        SA::splits(stats[i]) = st->totalSplits[i];
    }

    int64_t before = 0;
    for (int i = 0; i < n; i++) before += SA::walkLen(stats[i]);
    std::unique_ptr<tbb::global_control> gc;
    if (nThreads > 0) gc.reset(new tbb::global_control(tbb::global_control::max_allowed_parallelism, (size_t)nThreads));
    auto t0 = std::chrono::steady_clock::now();
    if (estimator == ZB_WALK_ON_SPHERES) {
        WalkOnSpheres<float, DIM> wos(*s.queries);
        wos.solve(pde, settings, nWalks, samplePts, stats, nThreads == 1);
    } else {
        WalkOnStars<float, DIM> wost(*s.queries);
        wost.solve(pde, settings, nWalks, samplePts, stats, nThreads == 1);
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

    int64_t after = 0;
    for (int i = 0; i < n; i++) {
        pts->rngState[i] = samplePts[i].rng.state;
        pts->rngInc[i] = samplePts[i].rng.inc;
        pts->firstSphereRadius[i] = samplePts[i].firstSphereRadius;
        st->solutionMean[i] = SA::mean(stats[i]);
        st->solutionM2[i] = SA::m2(stats[i]);
        for (size_t k = 0; k < DIM; k++) {
            st->gradientMean[DIM*i + k] = SA::gmean(stats[i])[k];
            st->gradientM2[DIM*i + k] = SA::gm2(stats[i])[k];
        }
        st->totalFirstSourceContribution[i] = SA::firstSource(stats[i]);
        st->totalDerivativeContribution[i] = SA::derivative(stats[i]);
        st->nSolutionEstimates[i] = SA::nSol(stats[i]);
        st->nGradientEstimates[i] = SA::nGrad(stats[i]);
        st->totalWalkLength[i] = SA::walkLen(stats[i]);
        st->totalSplits[i] = SA::splits(stats[i]);
        after += SA::walkLen(stats[i]);
    }
    if (walkSteps) *walkSteps = after - before;
    return ZB_OK;
}

// read access to the five protected SampleStatistics of bvc::EvaluationPoint (boundary_value_caching.h:44-51)
template <size_t DIM>
struct EvalAccess: public bvc::EvaluationPoint<float, DIM> {
    using Base = bvc::EvaluationPoint<float, DIM>;
    static SampleStatistics<float, DIM>& bucket(Base& e, int b) {
        switch (b) {
        case 0: return e.*(&EvalAccess::absorbingBoundaryStatistics);
        case 1: return e.*(&EvalAccess::absorbingBoundaryNormalAlignedStatistics);
        case 2: return e.*(&EvalAccess::reflectingBoundaryStatistics);
        case 3: return e.*(&EvalAccess::reflectingBoundaryNormalAlignedStatistics);
        default: return e.*(&EvalAccess::sourceStatistics);
        }
    }
};

// BoundaryValueCaching::splat(pde, samplePts, ..., evalPts) (boundary_value_caching.h:482-502) on caller-made
// sample points; the GeometricQueries / WalkOnStars the class holds are not touched by splat
template <size_t DIM>
int bvcSplat(float absorptionCoeff, const zb_bvc_samples *s, float radiusClamp, float kernelRegularization,
             float robinCutoff, float cutoffA, float cutoffR, int nEval, const float *evalPt, const int32_t *evalType,
             const float *distA, const float *distR, float *solution, float *gradient, int32_t *bucketCount, float *bucketMean)
{
    GeometricQueries<DIM> queries;
    WalkOnStars<float, DIM> wost(queries);
    bvc::BoundaryValueCaching<float, DIM> cache(queries, wost);
    PDE<float, DIM> pde;
    pde.absorptionCoeff = absorptionCoeff;

    std::vector<SamplePoint<float, DIM>> samplePts;
    for (int i = 0; i < s->n; i++) {
        SamplePoint<float, DIM> p(toVec<DIM>(s->pt + DIM*i), toVec<DIM>(s->normal + DIM*i), (SampleType)s->type[i],
                                  EstimationQuantity::Solution, s->pdf[i], 0.0f, 0.0f);
        p.solution = s->solution[i];
        p.normalDerivative = s->normalDerivative[i];
        p.contribution = s->contribution[i];
        p.robinCoeff = s->robinCoeff[i];
        p.estimateBoundaryNormalAligned = s->estimateBoundaryNormalAligned && s->estimateBoundaryNormalAligned[i];
        samplePts.push_back(p);
    }
    std::vector<bvc::EvaluationPoint<float, DIM>> evalPts;
    for (int i = 0; i < nEval; i++) {
        evalPts.emplace_back(toVec<DIM>(evalPt + DIM*i), Vec<DIM>::Zero(), (SampleType)evalType[i], distA[i], distR[i]);
    }
    cache.splat(pde, samplePts, radiusClamp, kernelRegularization, robinCutoff, cutoffA, cutoffR, evalPts);
    for (int i = 0; i < nEval; i++) {
        solution[i] = evalPts[i].getEstimatedSolution();
        for (size_t k = 0; k < DIM; k++) gradient[DIM*i + k] = evalPts[i].getEstimatedGradient((int)k);
        for (int b = 0; b < 5; b++) {
            SampleStatistics<float, DIM>& st = EvalAccess<DIM>::bucket(evalPts[i], b);
            bucketCount[5*i + b] = st.getSolutionEstimateCount();
            bucketMean[(5*i + b)*(1 + DIM)] = st.getEstimatedSolution();
            for (size_t k = 0; k < DIM; k++) bucketMean[(5*i + b)*(1 + DIM) + 1 + k] = st.getEstimatedGradient((int)k);
        }
    }
    return 0;
}

} // namespace

namespace {
template <size_t DIM>
void kelvinRobin(const float *origin, int nV, const float *tp, int nP, const int32_t *idx, const float *mn, const float *mx,
                 float *outMin, float *outMax)
{
    KelvinTransform<float, DIM> k(toVec<DIM>(origin));
    std::vector<Vec<DIM>> pts;
    std::vector<Veci<DIM>> ind;
    for (int i = 0; i < nV; i++) pts.push_back(toVec<DIM>(tp + DIM*i));
    for (int i = 0; i < nP; i++) { Veci<DIM> t; for (size_t j = 0; j < DIM; j++) t[j] = idx[DIM*i + j]; ind.push_back(t); }
    std::vector<float> a(mn, mn + nP), b(mx, mx + nP), oa, ob;
    k.computeRobinCoefficients(pts, ind, a, b, oa, ob);
    for (int i = 0; i < nP; i++) { outMin[i] = oa[i]; outMax[i] = ob[i]; }
}

template <size_t DIM>
void kelvinBack(const float *origin, int n, const float *y, const float *V, const float *dV, float *u, float *du)
{
    KelvinTransform<float, DIM> k(toVec<DIM>(origin));
    for (int i = 0; i < n; i++) {
        Vec<DIM> yy = toVec<DIM>(y + DIM*i);
        u[i] = k.transformSolutionEstimate(V[i], yy);
        std::vector<float> g(dV + DIM*i, dV + DIM*i + DIM), out;
        k.transformGradientEstimate(V[i], g, yy, out);
        for (size_t j = 0; j < DIM; j++) du[DIM*i + j] = out[j];
    }
}

// rws::ReverseWalkOnStarsSolver with caller-made sample points instead of its samplers' (reverse_walk_solver.h:54-108)
template <size_t DIM>
struct RwsAccess: public rws::ReverseWalkOnStarsSolver<float, DIM, NearestNeighborFinder<DIM>> {
    using Base = rws::ReverseWalkOnStarsSolver<float, DIM, NearestNeighborFinder<DIM>>;
    RwsAccess(const GeometricQueries<DIM>& q): Base(q, nullptr, nullptr, nullptr) {}
    std::vector<SamplePoint<float, DIM>>& group(int type, bool aligned) {
        if (type == ZB_ON_ABSORBING_BOUNDARY) return aligned ? this->absorbingBoundaryNormalAlignedSamplePts : this->absorbingBoundarySamplePts;
        if (type == ZB_ON_REFLECTING_BOUNDARY) return aligned ? this->reflectingBoundaryNormalAlignedSamplePts : this->reflectingBoundarySamplePts;
        return this->domainSamplePts;
    }
};

template <size_t DIM>
int rwsSolve(RefScene<DIM>& s, const zb_pde_desc *pdeDesc, const zb_walk_settings *ws, const zb_rws_samples *smp,
             float normalOffset, float radiusClamp, float kernelRegularization, zb_rws_eval_points *ev, int nThreads,
             int32_t *groupCounts, double *seconds)
{
    PDE<float, DIM> pde;
    buildPde<DIM>(*pdeDesc, pde);
    WalkSettings settings = toSettings(*ws);
    RwsAccess<DIM> solver(*s.queries);
    for (int i = 0; i < smp->n; i++) {
        bool aligned = smp->estimateBoundaryNormalAligned && smp->estimateBoundaryNormalAligned[i];
        SamplePoint<float, DIM> p(toVec<DIM>(smp->pt + DIM*i), toVec<DIM>(smp->normal + DIM*i), (SampleType)smp->type[i],
                                  EstimationQuantity::Solution, smp->pdf[i], smp->distToAbsorbingBoundary[i],
                                  smp->distToReflectingBoundary[i]);
        p.rng.state = smp->rngState[i];
        p.rng.inc = smp->rngInc[i];
        p.estimateBoundaryNormalAligned = aligned;
        solver.group(smp->type[i], aligned).push_back(p);
    }
    std::vector<rws::EvaluationPoint<float, DIM>> evalPts;
    Vec<DIM> zero = Vec<DIM>::Zero();
    for (int i = 0; i < ev->n; i++) {
        evalPts.emplace_back(toVec<DIM>(ev->pt + DIM*i), ev->normal ? toVec<DIM>(ev->normal + DIM*i) : zero,
                             (SampleType)ev->type[i], 0.0f, 0.0f);
        rws::EvaluationPoint<float, DIM>& e = evalPts.back();
        e.totalPoissonKernelContribution = ev->totals[6*i];
        e.totalAbsorbingBoundaryContribution = ev->totals[6*i + 1];
        e.totalAbsorbingBoundaryNormalAlignedContribution = ev->totals[6*i + 2];
        e.totalReflectingBoundaryContribution = ev->totals[6*i + 3];
        e.totalReflectingBoundaryNormalAlignedContribution = ev->totals[6*i + 4];
        e.totalSourceContribution = ev->totals[6*i + 5];
    }
    std::unique_ptr<tbb::global_control> gc;
    if (nThreads > 0) gc.reset(new tbb::global_control(tbb::global_control::max_allowed_parallelism, (size_t)nThreads));
    auto t0 = std::chrono::steady_clock::now();
    solver.solve(pde, settings, normalOffset, radiusClamp, kernelRegularization, evalPts, true, nThreads == 1);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int i = 0; i < ev->n; i++) {
        const rws::EvaluationPoint<float, DIM>& e = evalPts[i];
        ev->totals[6*i] = e.totalPoissonKernelContribution;
        ev->totals[6*i + 1] = e.totalAbsorbingBoundaryContribution;
        ev->totals[6*i + 2] = e.totalAbsorbingBoundaryNormalAlignedContribution;
        ev->totals[6*i + 3] = e.totalReflectingBoundaryContribution;
        ev->totals[6*i + 4] = e.totalReflectingBoundaryNormalAlignedContribution;
        ev->totals[6*i + 5] = e.totalSourceContribution;
    }
    if (groupCounts) {
        groupCounts[0] = solver.getAbsorbingBoundarySampleCount(false);
        groupCounts[1] = solver.getAbsorbingBoundarySampleCount(true);
        groupCounts[2] = solver.getReflectingBoundarySampleCount(false);
        groupCounts[3] = solver.getReflectingBoundarySampleCount(true);
        groupCounts[4] = solver.getDomainSampleCount();
    }
    return ZB_OK;
}

} // namespace

#define DISPATCH(scene, expr2, expr3) \
    (((AnyScene *)(scene))->dim == 2 ? (expr2) : (expr3))

extern "C" {

int zref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
int zref_tbb_max_concurrency() { return tbb::this_task_arena::max_concurrency(); }

void *zref_scene_create(int dim,
                        const float *aP, int nAV, const int32_t *aI, int nAT,
                        const float *rP, int nRV, const int32_t *rI, int nRT,
                        int reflectingKind, const float *minRobin, const float *maxRobin,
                        int solveDoubleSided, float silhouettePrecision,
                        int watertight, const float *dMin, const float *dMax, int useBaseline)
{
    AnyScene *s = new AnyScene();
    s->dim = dim;
    if (dim == 2) s->s2 = buildScene<2>(aP, nAV, aI, nAT, rP, nRV, rI, nRT, reflectingKind, minRobin, maxRobin,
                                        solveDoubleSided, silhouettePrecision, watertight, dMin, dMax, useBaseline);
    else s->s3 = buildScene<3>(aP, nAV, aI, nAT, rP, nRV, rI, nRT, reflectingKind, minRobin, maxRobin,
                               solveDoubleSided, silhouettePrecision, watertight, dMin, dMax, useBaseline);
    return s;
}

void zref_scene_destroy(void *scene) { delete (AnyScene *)scene; }

int zref_closest_point(void *scene, int which, int n, const float *x, const float *maxRadius,
                       int32_t *prim, float *dist, float *pt, float *normal, int nThreads)
{
    AnyScene *s = (AnyScene *)scene;
    return DISPATCH(s, closestPoint<2>(*s->s2, which, n, x, maxRadius, prim, dist, pt, normal, nThreads),
                       closestPoint<3>(*s->s3, which, n, x, maxRadius, prim, dist, pt, normal, nThreads));
}

int zref_intersect_ray(void *scene, int which, int n, const float *o, const float *d, const float *tMax,
                       int32_t *prim, float *dist, float *pt, float *normal, int nThreads)
{
    AnyScene *s = (AnyScene *)scene;
    return DISPATCH(s, intersectRay<2>(*s->s2, which, n, o, d, tMax, prim, dist, pt, normal, nThreads),
                       intersectRay<3>(*s->s3, which, n, o, d, tMax, prim, dist, pt, normal, nThreads));
}

int zref_count_ray_hits(void *scene, int which, int n, const float *o, const float *d, const float *tMax, int32_t *count, int nThreads)
{
    AnyScene *s = (AnyScene *)scene;
    return DISPATCH(s, countRayHits<2>(*s->s2, which, n, o, d, tMax, count, nThreads),
                       countRayHits<3>(*s->s3, which, n, o, d, tMax, count, nThreads));
}

int zref_inside_domain(void *scene, int n, const float *x, uint8_t *inside, int nThreads)
{
    AnyScene *s = (AnyScene *)scene;
    return DISPATCH(s, insideDomain<2>(*s->s2, n, x, inside, nThreads), insideDomain<3>(*s->s3, n, x, inside, nThreads));
}

int zref_occluded(void *scene, int which, int n, const float *a, const float *b, uint8_t *out, int nThreads)
{
    AnyScene *s = (AnyScene *)scene;
    return DISPATCH(s, occluded<2>(*s->s2, which, n, a, b, out, nThreads),
                       occluded<3>(*s->s3, which, n, a, b, out, nThreads));
}

int zref_sample_sphere(void *scene, int n, const float *x, const float *radius, const float *randNums,
                       int32_t *prim, float *pt, float *normal, float *pdf, int nThreads)
{
    AnyScene *s = (AnyScene *)scene;
    return DISPATCH(s, sampleSphere<2>(*s->s2, n, x, radius, randNums, prim, pt, normal, pdf, nThreads),
                       sampleSphere<3>(*s->s3, n, x, radius, randNums, prim, pt, normal, pdf, nThreads));
}

int zref_star_radius(void *scene, int n, const float *x, const float *minR, const float *maxR,
                     float precision, const uint8_t *flip, float *out, int nThreads)
{
    AnyScene *s = (AnyScene *)scene;
    return DISPATCH(s, starRadius<2>(*s->s2, n, x, minR, maxR, precision, flip, out, nThreads),
                       starRadius<3>(*s->s3, n, x, minR, maxR, precision, flip, out, nThreads));
}

int zref_solve(void *scene, int estimator, const zb_pde_desc *pde, const zb_walk_settings *ws,
               const zb_sample_points *pts, zb_statistics *st, int64_t *walkSteps, int nThreads, double *seconds)
{
    AnyScene *s = (AnyScene *)scene;
    return DISPATCH(s, solve<2>(*s->s2, estimator, pde, ws, pts, st, walkSteps, nThreads, seconds),
                       solve<3>(*s->s3, estimator, pde, ws, pts, st, walkSteps, nThreads, seconds));
}

// pcg32 known-answer hook: deps/pcg32/pcg32.h
void zref_pcg32(uint64_t initstate, uint64_t initseq, int n, uint32_t *out)
{
    pcg32 rng(initstate, initseq);
    for (int i = 0; i < n; i++) out[i] = rng.nextUInt();
}

// the whole of deps/pcg32/pcg32-demo.cpp:64-104 per round: 6 x nextUInt, 65 x nextUInt(2), 33 x nextUInt(6),
// shuffle of 52 cards (bounded rejection sampling + Fisher-Yates) — pins nextUInt(bound) and shuffle
void zref_pcg32_demo(uint64_t initstate, uint64_t initseq, int rounds, uint32_t *u32, uint8_t *coins, uint8_t *rolls, uint8_t *cards)
{
    pcg32 rng;
    rng.seed(initstate, initseq);
    for (int r = 0; r < rounds; r++) {
        for (int i = 0; i < 6; i++) u32[6*r + i] = rng.nextUInt();
        for (int i = 0; i < 65; i++) coins[65*r + i] = (uint8_t)rng.nextUInt(2);
        for (int i = 0; i < 33; i++) rolls[33*r + i] = (uint8_t)(rng.nextUInt(6) + 1);
        uint8_t *c = cards + 52*r;
        for (int i = 0; i < 52; i++) c[i] = (uint8_t)i;
        rng.shuffle(c, c + 52);
    }
}

// nextFloat stream (deps/pcg32/pcg32.h:103-112)
void zref_pcg32_floats(uint64_t initstate, uint64_t initseq, int n, float *out)
{
    pcg32 rng(initstate, initseq);
    for (int i = 0; i < n; i++) out[i] = rng.nextFloat();
}

// BVC splat through the reference (seconds = time inside the call)
int zref_bvc_splat(int dim, float absorptionCoeff, const zb_bvc_samples *s, float radiusClamp, float kernelRegularization,
                   float robinCutoff, float cutoffA, float cutoffR, int nEval, const float *evalPt, const int32_t *evalType,
                   const float *distA, const float *distR, float *solution, float *gradient, int32_t *bucketCount,
                   float *bucketMean, double *seconds)
{
    auto t0 = std::chrono::steady_clock::now();
    int rc = dim == 2 ? bvcSplat<2>(absorptionCoeff, s, radiusClamp, kernelRegularization, robinCutoff, cutoffA, cutoffR, nEval,
                                    evalPt, evalType, distA, distR, solution, gradient, bucketCount, bucketMean)
                      : bvcSplat<3>(absorptionCoeff, s, radiusClamp, kernelRegularization, robinCutoff, cutoffA, cutoffR, nEval,
                                    evalPt, evalType, distA, distR, solution, gradient, bucketCount, bucketMean);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

// KelvinTransform host helpers (kelvin_transform.h:109-121, 244-279, 351-388) for the parity tests of
// zombie_b200/kelvin.py
int zref_kelvin_transform_points(int dim, const float *origin, int n, const float *x, float *y)
{
    if (dim == 2) { KelvinTransform<float, 2> k(toVec<2>(origin)); for (int i = 0; i < n; i++) fromVec<2>(k.transformPoint(toVec<2>(x + 2*i)), y + 2*i); }
    else { KelvinTransform<float, 3> k(toVec<3>(origin)); for (int i = 0; i < n; i++) fromVec<3>(k.transformPoint(toVec<3>(x + 3*i)), y + 3*i); }
    return 0;
}

int zref_kelvin_robin_coefficients(int dim, const float *origin, int nV, const float *transformedPoints, int nP,
                                   const int32_t *indices, const float *minCoeff, const float *maxCoeff, float *outMin, float *outMax)
{
    if (dim == 2) kelvinRobin<2>(origin, nV, transformedPoints, nP, indices, minCoeff, maxCoeff, outMin, outMax);
    else kelvinRobin<3>(origin, nV, transformedPoints, nP, indices, minCoeff, maxCoeff, outMin, outMax);
    return 0;
}

int zref_kelvin_back_transform(int dim, const float *origin, int n, const float *y, const float *V, const float *dV, float *u, float *du)
{
    if (dim == 2) kelvinBack<2>(origin, n, y, V, dV, u, du);
    else kelvinBack<3>(origin, n, y, V, dV, u, du);
    return 0;
}

// rws::ReverseWalkOnStarsSolver::solve on caller-made samples; totals (n*6) are in/out
int zref_rws_solve(void *scene, const zb_pde_desc *pde, const zb_walk_settings *ws, const zb_rws_samples *samples,
                   float normalOffset, float radiusClamp, float kernelRegularization, zb_rws_eval_points *evalPts,
                   int nThreads, int32_t *groupCounts, double *seconds)
{
    return DISPATCH(scene, rwsSolve<2>(*((AnyScene *)scene)->s2, pde, ws, samples, normalOffset, radiusClamp, kernelRegularization, evalPts, nThreads, groupCounts, seconds),
                           rwsSolve<3>(*((AnyScene *)scene)->s3, pde, ws, samples, normalOffset, radiusClamp, kernelRegularization, evalPts, nThreads, groupCounts, seconds));
}

} // extern "C"
