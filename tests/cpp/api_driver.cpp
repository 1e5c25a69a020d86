// Test driver for include/zombie_b200/zombie.hpp: user code written against Zombie's C++ surface
// (handlers -> GeometricQueries -> PDE -> WalkOnSpheres/WalkOnStars::solve), reading a problem from a
// binary file written by tests/test_cpp_api.py and writing the statistics back. Built by the tests with
//   g++ -std=c++17 -I include tests/cpp/api_driver.cpp -L zombie_b200 -lzombie_b200
// Exit codes: 0 ok, 3 = zombie_b200::Error (message on stderr), 2 = bad usage / IO.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <zombie_b200/zombie.hpp>

template <typename T>
static bool readv(FILE *f, std::vector<T>& v, size_t n)
{
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}

struct Header {
    int32_t dim, estimator, reflectingKind, nVa, nPa, nVr, nPr, nPts, nWalks, quantity, family, solveDoubleSided;
    float params[16];
    float absorptionCoeff, robinCoeff, epsA, epsR;
    int32_t maxWalkLength, pad;
    uint64_t seed;
};

template <size_t DIM>
static int run(const Header& h, FILE *f, const char *outPath, bool extras)
{
    using namespace zombie; // = zombie_b200: the same names user code uses with the reference
    std::vector<float> ap, rp, mn, mx, pts;
    std::vector<int32_t> ai, ri;
    if (!readv(f, ap, (size_t)h.nVa*DIM) || !readv(f, ai, (size_t)h.nPa*DIM) || !readv(f, rp, (size_t)h.nVr*DIM) ||
        !readv(f, ri, (size_t)h.nPr*DIM) || !readv(f, mn, (size_t)h.nPr) || !readv(f, mx, (size_t)h.nPr) ||
        !readv(f, pts, (size_t)h.nPts*DIM)) return 2;

    auto toVectors = [](const std::vector<float>& flat) {
        std::vector<Vector<DIM>> v(flat.size()/DIM);
        for (size_t i = 0; i < v.size(); i++) for (size_t k = 0; k < DIM; k++) v[i][k] = flat[i*DIM + k];
        return v;
    };
    auto toIndices = [](const std::vector<int32_t>& flat) {
        std::vector<Vectori<DIM>> v(flat.size()/DIM);
        for (size_t i = 0; i < v.size(); i++) for (size_t k = 0; k < DIM; k++) v[i][k] = flat[i*DIM + k];
        return v;
    };

    // bounding box of the domain, padded like the Python binding does
    Vector<DIM> bmin = Vector<DIM>::Constant(1e30f), bmax = Vector<DIM>::Constant(-1e30f);
    for (const std::vector<float> *flat : {&ap, &rp})
        for (size_t i = 0; i < flat->size(); i++) {
            size_t k = i%DIM;
            bmin[k] = std::min(bmin[k], (*flat)[i]);
            bmax[k] = std::max(bmax[k], (*flat)[i]);
        }
    for (size_t k = 0; k < DIM; k++) { bmin[k] -= 1e-3f; bmax[k] += 1e-3f; }

    GeometricQueries<DIM> queries(true, bmin, bmax);
    FcpwDirichletBoundaryHandler<DIM> dirichletHandler;
    FcpwNeumannBoundaryHandler<DIM> neumannHandler;
    FcpwRobinBoundaryHandler<DIM> robinHandler;
    if (h.nPa > 0) {
        dirichletHandler.buildAccelerationStructure(toVectors(ap), toIndices(ai));
        populateGeometricQueriesForDirichletBoundary<DIM>(dirichletHandler, queries);
    }
    auto ignore = getIgnoreCandidateSilhouetteCallback(h.solveDoubleSided != 0, 1e-3f);
    auto weight = getBranchTraversalWeightCallback();
    if (h.nPr > 0 && h.reflectingKind == ZB_REFLECTING_NEUMANN) {
        neumannHandler.buildAccelerationStructure(toVectors(rp), toIndices(ri), ignore);
        populateGeometricQueriesForNeumannBoundary<DIM>(neumannHandler, weight, queries);
    } else if (h.nPr > 0) {
        robinHandler.buildAccelerationStructure(toVectors(rp), toIndices(ri), ignore, mn, mx);
        populateGeometricQueriesForRobinBoundary<DIM>(robinHandler, weight, queries);
    }

    PDE<float, DIM> pde;
    pde.absorptionCoeff = h.absorptionCoeff;
    pde.setManufactured((zb_solution_family)h.family, std::vector<float>(h.params, h.params + 16), h.robinCoeff);

    std::vector<Vector<DIM>> x = toVectors(pts);
    std::vector<float> distA, distR;
    queries.computeDistToAbsorbingBoundary_batch(x, distA);
    queries.computeDistToReflectingBoundary_batch(x, distR);

    std::vector<SamplePoint<float, DIM>> samplePts;
    std::vector<int> nWalks(h.nPts, h.nWalks);
    for (int i = 0; i < h.nPts; i++) {
        SamplePoint<float, DIM> p(x[i], Vector<DIM>::Zero(), SampleType::InDomain, (EstimationQuantity)h.quantity, 1.0f, distA[i], distR[i]);
        p.rng = pcg32(h.seed, (uint64_t)i); // explicit seeding instead of seedRng()'s clock (common.h:173)
        samplePts.push_back(p);
    }
    std::vector<SampleStatistics<float, DIM>> statistics;
    WalkSettings settings(h.epsA, h.epsR, h.maxWalkLength, h.solveDoubleSided != 0);
    const std::vector<SamplePoint<float, DIM>> initialPts = samplePts; // generators before any solve

    int64_t steps = 0;
    if (h.estimator == ZB_WALK_ON_SPHERES) {
        WalkOnSpheres<float, DIM> wos(queries);
        wos.solve(pde, settings, nWalks, samplePts, statistics);
        steps = wos.lastWalkSteps;
    } else {
        WalkOnStars<float, DIM> wost(queries);
        wost.solve(pde, settings, nWalks, samplePts, statistics);
        steps = wost.lastWalkSteps;
    }

    // single-query callbacks agree with the batch
    float d0 = queries.computeDistToAbsorbingBoundary(x[0], false);
    if (h.nPa > 0 && d0 != distA[0]) { fprintf(stderr, "single/batched distance mismatch %g %g\n", d0, distA[0]); return 4; }

    // insideDomain (all-hits ray parity on the device): evaluation points lie inside, a far point outside; single == batched
    {
        std::vector<Vector<DIM>> probe(x.begin(), x.begin() + std::min<size_t>(x.size(), 64));
        Vector<DIM> far = Vector<DIM>::Constant(7.5f);
        probe.push_back(far);
        std::vector<uint8_t> inside;
        queries.insideDomain_batch(probe, inside);
        for (size_t i = 0; i + 1 < probe.size(); i++) {
            if (!inside[i]) { fprintf(stderr, "evaluation point %d reported outside the domain\n", (int)i); return 5; }
        }
        if (inside.back()) { fprintf(stderr, "far point reported inside the domain\n"); return 5; }
        if (queries.insideDomain(probe[0]) != (inside[0] != 0) || queries.insideDomain(far)) { fprintf(stderr, "single/batched insideDomain mismatch\n"); return 5; }
    }

    // Vector-valued solve through the same surface (the reference's T = Array<float, C>): two channels on one set of walks —
    // the problem's own data and constant data; channel 0 must reproduce the scalar solve above on the same generators
    if (extras) {
        PDE<float, DIM> flat = pde;
        flat.setConstant(0.25f, -0.5f, 0.125f, h.robinCoeff);
        flat.isSourceConstant = pde.isSourceConstant;
        flat.areRobinConditionsPureNeumann = pde.areRobinConditionsPureNeumann;
        flat.areRobinCoeffsNonnegative = pde.areRobinCoeffsNonnegative;
        std::vector<SamplePoint<float, DIM>> pts2 = initialPts;
        std::vector<std::vector<SampleStatistics<float, DIM>>> channelStats;
        int64_t steps2 = solveChannels<DIM>(h.estimator, queries, {pde, flat}, settings, nWalks, pts2, channelStats);
        if (steps2 != steps) { fprintf(stderr, "channel solve took %lld steps, scalar solve %lld\n", (long long)steps2, (long long)steps); return 6; }
        for (int i = 0; i < h.nPts; i++) {
            float a = channelStats[0][i].getEstimatedSolution(), b = statistics[i].getEstimatedSolution();
            if (std::fabs(a - b) > 2e-5f*std::max(1.0f, std::fabs(b)) || pts2[i].rng.state != samplePts[i].rng.state ||
                channelStats[1][i].getSolutionEstimateCount() != statistics[i].getSolutionEstimateCount() ||
                !std::isfinite(channelStats[1][i].getEstimatedSolution())) {
                fprintf(stderr, "channel solve deviates at point %d: %g vs %g\n", i, a, b);
                return 6;
            }
        }
    }

    // Boundary Value Caching splat and Reverse Walk on Stars through the same surface: a cache made of the solved
    // points themselves (solution = estimate, no flux), splatted onto / walked towards two interior points
    float bvcSolution = 0.0f, rwsSolution = 0.0f;
    if (extras) {
        std::vector<SamplePoint<float, DIM>> cache;
        for (int i = 0; i < h.nPts && i < 32; i++) {
            SamplePoint<float, DIM> c = samplePts[i];
            c.type = SampleType::InDomain;       // treated as source samples by both solvers
            c.pdf = 1.0f;
            c.contribution = statistics[i].getEstimatedSolution();
            cache.push_back(c);
        }
        WalkOnStars<float, DIM> wostForCache(queries);
        bvc::BoundaryValueCaching<float, DIM> caching(queries, wostForCache);
        std::vector<bvc::EvaluationPoint<float, DIM>> bvcPts;
        bvcPts.emplace_back(x[0], Vector<DIM>::Zero(), SampleType::InDomain, distA[0], distR[0]);
        caching.splat(pde, cache, 1e-3f, 0.0f, std::numeric_limits<float>::max(), 0.0f, 0.0f, bvcPts);
        bvcSolution = bvcPts[0].getEstimatedSolution();
        std::vector<rws::EvaluationPoint<float, DIM>> rwsPts;
        rwsPts.emplace_back(x[0], Vector<DIM>::Zero(), SampleType::InDomain, distA[0], distR[0]);
        rws::solve<float, DIM>(queries, pde, settings, cache, 0.01f, 1e-3f, 0.0f, rwsPts);
        rwsSolution = rwsPts[0].getEstimatedSolution(0, 0, 0, 0, (int)cache.size());
    }

    FILE *o = fopen(outPath, "wb");
    if (!o) return 2;
    fwrite(&steps, sizeof(steps), 1, o);
    fwrite(&bvcSolution, sizeof(float), 1, o);
    fwrite(&rwsSolution, sizeof(float), 1, o);
    for (int i = 0; i < h.nPts; i++) {
        const SampleStatistics<float, DIM>& s = statistics[i];
        float rec[8] = {s.getEstimatedSolution(), s.getEstimatedSolutionVariance(), (float)s.getSolutionEstimateCount(),
                        s.getMeanWalkLength(), samplePts[i].firstSphereRadius, s.getEstimatedGradient(0), s.getEstimatedGradient(1),
                        DIM == 3 ? s.getEstimatedGradient((int)DIM - 1) : 0.0f};
        fwrite(rec, sizeof(float), 8, o);
    }
    fclose(o);
    printf("ok %lld steps\n", (long long)steps);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: api_driver <problem.bin> <out.bin> [extras]\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    Header h;
    if (fread(&h, sizeof(h), 1, f) != 1) return 2;
    try {
        bool extras = argc > 3; // also run the BVC splat and the reverse walks
        int rc = h.dim == 2 ? run<2>(h, f, argv[2], extras) : run<3>(h, f, argv[2], extras);
        fclose(f);
        return rc;
    } catch (const zombie_b200::Error& e) {
        fprintf(stderr, "zombie_b200::Error %d: %s\n", e.code, e.what());
        return 3;
    }
}
