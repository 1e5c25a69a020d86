// DEBUG / TEST TOOL, not part of libzombie_b200.so: runs the *device* query and walk code
// (zombie_b200/csrc/{queries,walk}.cuh, compiled for the host) one walk at a time on a CPU and
// records, per geometric query, how many BVH nodes and primitives it touched. Used to understand
// traversal cost distributions without spending GPU time.
//   nvcc -std=c++17 -O2 -DZB_HOST_STATS -Xcompiler -ffp-contract=off -o /tmp/host_probe tools/host_probe.cu zombie_b200/csrc/wide_bvh.cpp
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <map>

#include "../zombie_b200/csrc/walk.cuh"
#include "../zombie_b200/csrc/wide_bvh.h"

using namespace zb;

struct Rec { int q, nodes, prims; };
static std::vector<Rec> g_recs;
namespace zb { long long g_zbClass[24] = {0}; }
namespace zb { void zbHostRecord(int query, int nodes, int prims) { g_recs.push_back({query, nodes, prims}); } }

static BvhView viewOf(const HostBvh& h, int sign)
{
    BvhView v{};
    v.nodes = reinterpret_cast<const uint4 *>(h.nodes.data());
    v.prims = reinterpret_cast<const float4 *>(h.prims.data());
    v.primIndex = h.primIndex.data();
    v.cones = h.cones.empty() ? nullptr : reinterpret_cast<const float4 *>(h.cones.data());
    v.adjacency = h.adjacency.empty() ? nullptr : reinterpret_cast<const float4 *>(h.adjacency.data());
    v.nPrims = h.nPrims;
    v.snchSign = sign;
    return v;
}

#include "probe_mesh.h"

int main(int argc, char **argv)
{
    int subdiv = argc > 1 ? atoi(argv[1]) : 5;
    int nPts = argc > 2 ? atoi(argv[2]) : 256;
    int W = argc > 3 ? atoi(argv[3]) : 8;
    float mu = argc > 4 ? (float)atof(argv[4]) : 1.0f;       // 0 -> Neumann
    float sigma = argc > 5 ? (float)atof(argv[5]) : 1.0f;
    std::vector<float> pos; std::vector<int> tri;
    icosphere(subdiv, pos, tri);
    // partition by centroid z
    std::vector<int> triA, triR;
    for (size_t i = 0; i < tri.size(); i += 3) {
        float cz = (pos[3*tri[i]+2] + pos[3*tri[i+1]+2] + pos[3*tri[i+2]+2])/3.0f;
        auto& dst = cz > 0 ? triR : triA;
        dst.push_back(tri[i]); dst.push_back(tri[i+1]); dst.push_back(tri[i+2]);
    }
    HostBvh A, R;
    const char *err = nullptr;
    ReflectingOptions none, opt;
    opt.enabled = true;
    std::vector<float> coeff(triR.size()/3, mu);
    if (mu > 0) { opt.minCoeff = coeff.data(); opt.maxCoeff = coeff.data(); }
    if (!buildWideBvh(3, pos.data(), (int)pos.size()/3, triA.data(), (int)triA.size()/3, none, A, &err)) { printf("A: %s\n", err); return 1; }
    if (!buildWideBvh(3, pos.data(), (int)pos.size()/3, triR.data(), (int)triR.size()/3, opt, R, &err)) { printf("R: %s\n", err); return 1; }
    printf("tris A %d R %d | wide nodes A %zu R %zu depth %d %d\n", A.nPrims, R.nPrims, A.nodes.size(), R.nodes.size(), A.maxDepth, R.maxDepth);

    int status = 0;
    SolveArgs a{};
    a.absorbing = viewOf(A, 1);
    a.reflecting = viewOf(R, mu > 0 ? 1 : -1);
    a.hasAbsorbing = a.hasReflecting = 1;
    a.reflectingKind = mu > 0 ? ZB_REFLECTING_ROBIN : ZB_REFLECTING_NEUMANN;
    a.estimator = ZB_WALK_ON_STARS;
    for (int k = 0; k < 3; k++) { a.domainMin[k] = -2; a.domainMax[k] = 2; }
    zb_walk_settings ws{1e-3f, 1e-3f, 1e-3f, 0.0f, kMaxFloat, 1024, 0, 1024, 0, 1, 1, 0, 0, 0, 0, 0};
    a.ws = ws;
    zb_pde_desc pde{};
    pde.absorptionCoeff = sigma;
    pde.areRobinConditionsPureNeumann = mu == 0;
    pde.areRobinCoeffsNonnegative = 1;
    pde.mode = ZB_PDE_MANUFACTURED;
    pde.family = sigma > 0 ? ZB_SOL_EXP_DIR : ZB_SOL_QUADRATIC;
    float k3 = 1.0f/std::sqrt(3.0f);
    if (sigma > 0) { pde.params[0] = pde.params[1] = pde.params[2] = k3; pde.params[3] = 1; }
    else { pde.params[0] = 1; pde.params[1] = -0.5f; pde.params[2] = 0.25f; pde.params[3] = 0.3f; }
    pde.robinCoeff = mu;
    a.pde = pde;
    a.status = &status;

    // points in the ball r < 0.7
    std::vector<float> pt(3*nPts), nrm(3*nPts, 0.0f), dA(nPts), dR(nPts), fsr(nPts, 0.0f);
    std::vector<int> type(nPts, 0), quantity(nPts, ZB_SOLUTION), nWalks(nPts, W);
    std::vector<uint64_t> rs(nPts), ri(nPts);
    Pcg32 gen; gen.seed(7, 1);
    for (int i = 0; i < nPts; i++) {
        float x, y, z;
        do { x = 2*gen.nextFloat()-1; y = 2*gen.nextFloat()-1; z = 2*gen.nextFloat()-1; } while (x*x+y*y+z*z > 1);
        pt[3*i] = 0.7f*x; pt[3*i+1] = 0.7f*y; pt[3*i+2] = 0.7f*z;
        Pcg32 r; r.seed(0xC0FFEE, i); rs[i] = r.state; ri[i] = r.inc;
        ClosestHit<3> h;
        closestPoint<3>(a.absorbing, loadVec<3>(&pt[3*i]), kMaxFloat, h, &status); dA[i] = h.d;
        closestPoint<3>(a.reflecting, loadVec<3>(&pt[3*i]), kMaxFloat, h, &status); dR[i] = h.d;
    }
    g_recs.clear();
    std::vector<float> mean(nPts, 0), m2(nPts, 0), gm(3*nPts, 0), gm2(3*nPts, 0), fs(nPts, 0), dc(nPts, 0);
    std::vector<int> ns(nPts, 0), ng(nPts, 0), wl(nPts, 0), sp(nPts, 0);
    a.pts = zb_sample_points{nPts, pt.data(), nrm.data(), nullptr, type.data(), quantity.data(), dA.data(), dR.data(), nullptr, nWalks.data(), rs.data(), ri.data(), fsr.data()};
    a.stats = zb_statistics{mean.data(), m2.data(), gm.data(), gm2.data(), fs.data(), dc.data(), ns.data(), ng.data(), wl.data(), sp.data()};
    std::vector<int> taskCount(nPts), stratCount(nPts);
    std::vector<long long> taskOffset(nPts + 1, 0), stratOffset(nPts + 1, 0);
    std::vector<unsigned long long> pointState(nPts);
    a.taskCount = taskCount.data(); a.taskOffset = taskOffset.data(); a.stratCount = stratCount.data(); a.stratOffset = stratOffset.data();
    a.pointState = pointState.data();
    for (int i = 0; i < nPts; i++) countTasksPoint<3>(a, i);
    for (int i = 0; i < nPts; i++) { taskOffset[i+1] = taskOffset[i] + taskCount[i]; stratOffset[i+1] = stratOffset[i] + stratCount[i]; }
    std::vector<float> strat(stratOffset[nPts] + 1);
    a.strat = strat.data();
    long long total = taskOffset[nPts];
    a.totalTasks = total;
    std::vector<float> recSol(total), recFS(total), recGB(3*total), recGS(3*total);
    std::vector<int> recLen(total), recFlags(total);
    a.recSolution = recSol.data(); a.recFirstSource = recFS.data(); a.recGradBoundary = recGB.data(); a.recGradSource = recGS.data();
    a.recLength = recLen.data(); a.recFlags = recFlags.data(); a.anyGradient = 1;
    for (int i = 0; i < nPts; i++) preparePoint<3>(a, i);
    g_recs.clear();
    for (int k = 0; k < 24; k++) g_zbClass[k] = 0;

    // run the walks; per step, collect the cost of each query
    struct StepCost { int nodes[5], prims[5]; int stepIndex; bool onRefl; float distA; float ballR, distNow; };
    std::vector<StepCost> steps;
    long long nSteps = 0;
    for (long long t = 0; t < total; t++) {
        Lane<3> L; L.active = false;
        WalkState<3> queue[kSplitStack];
        beginTask<3, true>(a, L, t);
        g_recs.clear();
        while (L.active) {
            bool onRefl = L.st.onRefl; float distA = L.distA; int idx = L.st.walkLength;
            StepResult r = walkStep<3, true>(a, L, queue);
            StepCost c{};
            c.stepIndex = idx; c.onRefl = onRefl; c.distA = distA; c.ballR = L.g.R; c.distNow = L.distA;
            for (auto& rec : g_recs) { c.nodes[rec.q] += rec.nodes; c.prims[rec.q] += rec.prims; }
            g_recs.clear();
            steps.push_back(c);
            if (r != kContinue) finishState<3, true>(a, L, r, queue); else nSteps++;
        }
    }
    printf("walks %lld steps %lld status %d\n", total, nSteps, status);
    printf("star items: with silhouette tests %lld, robin-only %lld | prims rejected by distance %lld, by the plane-distance bound %lld, fully tested %lld\n", g_zbClass[0], g_zbClass[1], g_zbClass[2], g_zbClass[4], g_zbClass[3]);
    printf("star prims (fast test): robin-only skipped by the plane bound %lld, robin-only evaluated %lld, silhouette prims without / with the reflectance part %lld / %lld\n",
           g_zbClass[4], g_zbClass[7], g_zbClass[5], g_zbClass[6]);
    printf("star edges (fast test): skipped by the sign shortcut %lld, tested in full %lld\n", g_zbClass[8], g_zbClass[9]);
    printf("star children: %lld passed the box test, %lld of them went on to the reflectance bounds\n", g_zbClass[10], g_zbClass[11]);
    printf("star prims redone in the reference's order: %lld of %lld full tests (grazing view %lld, bound threshold %lld, conditioning %lld)\n",
           g_zbClass[15], g_zbClass[3], g_zbClass[17], g_zbClass[18], g_zbClass[19]);
    printf("star prims that lowered r2: %lld (robin-only items: %lld); extra rejected by the range bound %lld\n", g_zbClass[12], g_zbClass[13], g_zbClass[14]);
    const char *names[5] = {"closest", "ray", "occlusion", "sample", "star"};
    for (int q = 0; q < 5; q++) {
        std::vector<int> n, p;
        for (auto& s : steps) { n.push_back(s.nodes[q]); p.push_back(s.prims[q]); }
        std::sort(n.begin(), n.end()); std::sort(p.begin(), p.end());
        auto pct = [&](std::vector<int>& v, double f) { return v[(size_t)std::min<double>(v.size()-1, f*v.size())]; };
        double mn = 0, mp = 0; for (int v : n) mn += v; for (int v : p) mp += v;
        printf("%-10s nodes mean %8.1f p50 %6d p90 %6d p99 %6d p99.9 %6d max %7d | prims mean %8.1f p50 %6d p90 %6d p99 %6d p99.9 %6d max %7d\n", names[q],
               mn/n.size(), pct(n, .5), pct(n, .9), pct(n, .99), pct(n, .999), n.back(), mp/p.size(), pct(p, .5), pct(p, .9), pct(p, .99), pct(p, .999), p.back());
    }
    {   // steps that start ON the reflecting boundary: their share of the steps and of the closest-point work
        long long nOn = 0, cn = 0, cp = 0, an = 0, ap = 0;
        for (auto& s : steps) { an += s.nodes[0]; ap += s.prims[0]; if (s.onRefl) { nOn++; cn += s.nodes[0]; cp += s.prims[0]; } }
        printf("steps starting on the reflecting boundary: %.1f %% of the steps, %.1f %% of the closest-point nodes, %.1f %% of its primitives\n",
               100.0*nOn/std::max<size_t>(1, steps.size()), 100.0*cn/std::max(1LL, an), 100.0*cp/std::max(1LL, ap));
    }
    {   // steps whose ball is set by the reflecting boundary (radius well below the absorbing distance): the distance query
        // of such a step could be bounded by the star radius if the star query ran first
        long long n = 0, cn = 0, cp = 0, an = 0, ap = 0;
        for (auto& s : steps) {
            an += s.nodes[0]; ap += s.prims[0];
            if (!s.onRefl && s.distNow < 1e30f && s.ballR < 0.5f*s.distNow) { n++; cn += s.nodes[0]; cp += s.prims[0]; }
        }
        printf("steps (not on the boundary) with ball radius < half the absorbing distance: %.1f %% of the steps, %.1f %% of the closest-point nodes, %.1f %% of its primitives\n",
               100.0*n/std::max<size_t>(1, steps.size()), 100.0*cn/std::max(1LL, an), 100.0*cp/std::max(1LL, ap));
    }
    // the worst star-radius steps
    std::sort(steps.begin(), steps.end(), [](const StepCost& x, const StepCost& y) { return x.prims[4] > y.prims[4]; });
    for (int i = 0; i < 8 && i < (int)steps.size(); i++)
        printf("  worst star: prims %d nodes %d step %d onRefl %d distA %.4f | closest prims %d\n", steps[i].prims[4], steps[i].nodes[4], steps[i].stepIndex, steps[i].onRefl, steps[i].distA, steps[i].prims[0]);
    // warp model: steps grouped 32 at a time in arrival order; cost of a warp step = max over lanes
    double sumMax = 0, sumMean = 0; size_t groups = 0;
    std::vector<StepCost> order = steps; // (sorted) -> use a shuffle-free proxy: original order lost; approximate with random grouping
    std::random_shuffle(order.begin(), order.end());
    for (size_t i = 0; i + 32 <= order.size(); i += 32) {
        double mx = 0, mean_ = 0;
        for (int l = 0; l < 32; l++) { double c = order[i+l].prims[4] + order[i+l].prims[0] + 2.0*(order[i+l].nodes[4] + order[i+l].nodes[0]); mx = std::max(mx, c); mean_ += c/32; }
        sumMax += mx; sumMean += mean_; groups++;
    }
    printf("SIMT model (32 random steps per warp step): mean cost %.1f, max-over-lanes %.1f -> lane efficiency %.3f\n", sumMean/groups, sumMax/groups, sumMean/sumMax);
    return 0;
}
