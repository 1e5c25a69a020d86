// DEBUG / TEST TOOL, not part of libzombie_b200.so: runs the *device* walk code (walk.cuh compiled for the host,
// one walk at a time, thread-per-query traversals) on a problem dumped by a Python script, so that estimator
// questions (bias vs the exact solution / the reference) can be studied without a GPU.
//   nvcc -std=c++17 -O2 -DZB_HOST_STATS -Xcompiler -ffp-contract=off -o /tmp/host_solve tools/host_solve.cu zombie_b200/csrc/wide_bvh.cpp
//   /tmp/host_solve <dir>   reads <dir>/{hdr,apos,aidx,rpos,ridx,pts}.bin and, when present, per-point
//   {type,quantity}.bin (int32) and {normal,dir}.bin (float, dim per point); writes <dir>/out.bin
//   (mean, m2, n, total walk length, derivative sum, firstSphereRadius per point). Optional: sides.bin (two floats: constant
//   Dirichlet data of the two sides of a double-sided absorbing boundary, no source) and channels.bin (16 floats per extra
//   value channel: the parameters of its solution family; its estimates go to out_c<k>.bin as mean, m2, n, gradient mean)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../zombie_b200/csrc/walk.cuh"
#include "../zombie_b200/csrc/wide_bvh.h"

using namespace zb;
namespace zb { long long g_zbClass[24] = {0}; void zbHostRecord(int, int, int) {} }

template <typename T> static std::vector<T> rd(const std::string& p, bool optional = false)
{
    FILE *f = fopen(p.c_str(), "rb");
    if (!f && optional) return std::vector<T>();
    if (!f) { printf("cannot open %s\n", p.c_str()); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(n/sizeof(T));
    if (n && fread(v.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
    return v;
}

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

struct Header { int dim, kind, nWalks, family; float params[16]; float sigma, mu, rr, split; unsigned long long seed; };

template <int DIM>
int run(const std::string& dir, const Header& h)
{
    auto apos = rd<float>(dir + "/apos.bin"); auto aidx = rd<int>(dir + "/aidx.bin");
    auto rpos = rd<float>(dir + "/rpos.bin"); auto ridx = rd<int>(dir + "/ridx.bin");
    auto pt = rd<float>(dir + "/pts.bin");
    int nPts = (int)pt.size()/DIM;
    HostBvh A, R;
    const char *err = nullptr;
    ReflectingOptions none, opt;
    opt.enabled = true;
    auto dsIn = rd<int>(dir + "/double_sided.bin", true); // optional: open reflecting surfaces, walks on both sides
    const bool doubleSided = !dsIn.empty() && dsIn[0] != 0;
    opt.solveDoubleSided = doubleSided;
    std::vector<float> coeff(ridx.size()/DIM, h.mu);
    if (h.kind == ZB_REFLECTING_ROBIN) { opt.minCoeff = coeff.data(); opt.maxCoeff = coeff.data(); }
    // optional per-primitive coefficient bounds (Kelvin-transformed Neumann data: kelvin_transform.h:350-388)
    auto rminIn = rd<float>(dir + "/rmin.bin", true), rmaxIn = rd<float>(dir + "/rmax.bin", true);
    if (h.kind == ZB_REFLECTING_ROBIN && rminIn.size() == ridx.size()/DIM && rmaxIn.size() == rminIn.size()) {
        opt.minCoeff = rminIn.data(); opt.maxCoeff = rmaxIn.data();
    }
    if (!buildWideBvh(DIM, apos.data(), (int)apos.size()/DIM, aidx.data(), (int)aidx.size()/DIM, none, A, &err)) { printf("A: %s\n", err); return 1; }
    bool hasR = !ridx.empty();
    if (hasR && !buildWideBvh(DIM, rpos.data(), (int)rpos.size()/DIM, ridx.data(), (int)ridx.size()/DIM, opt, R, &err)) { printf("R: %s\n", err); return 1; }

    int status = 0;
    SolveArgs a{};
    a.absorbing = viewOf(A, 1);
    if (hasR) a.reflecting = viewOf(R, h.kind == ZB_REFLECTING_NEUMANN && DIM == 3 ? -1 : 1);
    a.hasAbsorbing = 1; a.hasReflecting = hasR ? 1 : 0;
    a.reflectingKind = hasR ? h.kind : ZB_REFLECTING_NONE;
    a.estimator = hasR ? ZB_WALK_ON_STARS : ZB_WALK_ON_SPHERES;
    for (int k = 0; k < 3; k++) { a.domainMin[k] = -3; a.domainMax[k] = 3; }
    zb_walk_settings ws{1e-3f, 1e-3f, 1e-3f, h.rr, h.split, 1024, 0, 1024, 0, 1, 1, 0, 0, 0, 0, 0};
    ws.solveDoubleSided = doubleSided ? 1 : 0;
    a.ws = ws;
    zb_pde_desc pde{};
    pde.absorptionCoeff = h.sigma;
    pde.areRobinConditionsPureNeumann = h.mu == 0;
    pde.areRobinCoeffsNonnegative = 1;
    pde.mode = ZB_PDE_MANUFACTURED;
    pde.family = h.family;
    for (int k = 0; k < 16; k++) pde.params[k] = h.params[k];
    pde.robinCoeff = h.mu;
    // optional Kelvin wrapper: (flag, origin[3]); the flags of the wrapped PDE as zb_solve_dev sets them (capi.cu;
    // kelvin_transform.h:340-347)
    auto kelvinIn = rd<float>(dir + "/kelvin.bin", true);
    if (kelvinIn.size() == 4 && kelvinIn[0] != 0.0f) {
        pde.kelvin = 1;
        for (int k = 0; k < 3; k++) pde.kelvinOrigin[k] = kelvinIn[1 + k];
        pde.isSourceConstant = 0;
        pde.absorptionCoeff = 0.0f;
        if (DIM == 3) { pde.areRobinConditionsPureNeumann = 0; pde.areRobinCoeffsNonnegative = 0; }
    }
    // optional: two-sided constant Dirichlet data (zb_pde_desc.dirichletAlignedValue)
    auto sidesIn = rd<float>(dir + "/sides.bin", true);
    if (sidesIn.size() == 2) {
        pde.mode = ZB_PDE_CONSTANT;
        pde.isSourceConstant = 1;
        pde.dirichletValue = sidesIn[0];
        pde.dirichletAlignedValue = sidesIn[1];
        pde.hasDirichletAlignedValue = 1;
    }
    a.pde = pde;
    a.status = &status;
    // optional: extra value channels riding on the walks (zb_solve_channels)
    auto channelsIn = rd<float>(dir + "/channels.bin", true);
    const int nExtra = (int)channelsIn.size()/16;
    std::vector<zb_pde_desc> extraPde(nExtra, pde);
    for (int c = 0; c < nExtra; c++) for (int k = 0; k < 16; k++) extraPde[c].params[k] = channelsIn[16*c + k];

    std::vector<float> nrm(DIM*nPts, 0.0f), dA(nPts), dR(nPts, kMaxFloat), fsr(nPts, 0.0f);
    std::vector<int> type(nPts, 0), quantity(nPts, ZB_SOLUTION), nWalks(nPts, h.nWalks);
    auto typeIn = rd<int>(dir + "/type.bin", true), quantityIn = rd<int>(dir + "/quantity.bin", true);
    auto normalIn = rd<float>(dir + "/normal.bin", true), dirIn = rd<float>(dir + "/dir.bin", true);
    if ((int)typeIn.size() == nPts) type = typeIn;
    if ((int)quantityIn.size() == nPts) quantity = quantityIn;
    if ((int)normalIn.size() == DIM*nPts) nrm = normalIn;
    std::vector<uint64_t> rs(nPts), ri(nPts);
    for (int i = 0; i < nPts; i++) {
        Pcg32 r; r.seed(h.seed, i); rs[i] = r.state; ri[i] = r.inc;
        ClosestHit<DIM> hit;
        closestPoint<DIM>(a.absorbing, loadVec<DIM>(&pt[DIM*i]), kMaxFloat, hit, &status); dA[i] = hit.d;
        if (hasR) { closestPoint<DIM>(a.reflecting, loadVec<DIM>(&pt[DIM*i]), kMaxFloat, hit, &status); dR[i] = hit.d; }
    }
    std::vector<float> mean(nPts, 0), m2(nPts, 0), gm(DIM*nPts, 0), gm2(DIM*nPts, 0), fs(nPts, 0), dc(nPts, 0);
    std::vector<int> ns(nPts, 0), ng(nPts, 0), wl(nPts, 0), sp(nPts, 0);
    a.pts = zb_sample_points{nPts, pt.data(), nrm.data(), (int)dirIn.size() == DIM*nPts ? dirIn.data() : nullptr, type.data(), quantity.data(), dA.data(), dR.data(), nullptr, nWalks.data(), rs.data(), ri.data(), fsr.data()};
    a.stats = zb_statistics{mean.data(), m2.data(), gm.data(), gm2.data(), fs.data(), dc.data(), ns.data(), ng.data(), wl.data(), sp.data()};
    std::vector<int> taskCount(nPts), stratCount(nPts);
    std::vector<long long> taskOffset(nPts + 1, 0), stratOffset(nPts + 1, 0);
    std::vector<unsigned long long> pointState(nPts);
    a.taskCount = taskCount.data(); a.taskOffset = taskOffset.data(); a.stratCount = stratCount.data(); a.stratOffset = stratOffset.data();
    a.pointState = pointState.data();
    for (int i = 0; i < nPts; i++) countTasksPoint<DIM>(a, i);
    for (int i = 0; i < nPts; i++) { taskOffset[i+1] = taskOffset[i] + taskCount[i]; stratOffset[i+1] = stratOffset[i] + stratCount[i]; }
    std::vector<float> strat(stratOffset[nPts] + 1);
    a.strat = strat.data();
    long long total = taskOffset[nPts];
    a.totalTasks = total;
    std::vector<float> recSol(total), recFS(total), recGB(DIM*total), recGS(DIM*total);
    std::vector<int> recLen(total), recFlags(total);
    a.recSolution = recSol.data(); a.recFirstSource = recFS.data(); a.recGradBoundary = recGB.data(); a.recGradSource = recGS.data();
    a.recLength = recLen.data(); a.recFlags = recFlags.data(); a.anyGradient = 1;
    long long steps = 0;
    a.walkSteps = &steps;
    std::vector<float> xRefl((size_t)nExtra*total, 0.0f), xSrc(xRefl), xRecSol(xRefl), xRecFS(xRefl);
    ExtraChannelsView view{nExtra, total, extraPde.data(), xRefl.data(), xSrc.data(), xRecSol.data(), xRecFS.data()};
    if (nExtra) {
        a.nExtra = nExtra; a.extraPde = extraPde.data();
        a.xRefl = xRefl.data(); a.xSrc = xSrc.data(); a.xRecSolution = xRecSol.data(); a.xRecFirstSource = xRecFS.data();
        a.extra = &view;
    }
    for (int i = 0; i < nPts; i++) preparePoint<DIM>(a, i);
    for (long long t = 0; t < total; t++) {
        Lane<DIM> L; L.active = false;
        WalkState<DIM> queue[kSplitStack];
        if (nExtra && hasR) {
            beginTask<DIM, true, true>(a, L, t);
            while (L.active) { StepResult r = walkStep<DIM, true, true>(a, L, queue); if (r != kContinue) finishState<DIM, true, true>(a, L, r, queue); }
        } else if (nExtra) {
            beginTask<DIM, false, true>(a, L, t);
            while (L.active) { StepResult r = walkStep<DIM, false, true>(a, L, queue); if (r != kContinue) finishState<DIM, false, true>(a, L, r, queue); }
        } else if (hasR) {
            beginTask<DIM, true>(a, L, t);
            while (L.active) { StepResult r = walkStep<DIM, true>(a, L, queue); if (r != kContinue) finishState<DIM, true>(a, L, r, queue); }
        } else {
            beginTask<DIM, false>(a, L, t);
            while (L.active) { StepResult r = walkStep<DIM, false>(a, L, queue); if (r != kContinue) finishState<DIM, false>(a, L, r, queue); }
        }
    }
    // the extra channels fold the same walk records with their own sums (before channel 0 advances the generators)
    for (int c = 0; c < nExtra; c++) {
        std::vector<float> cm(nPts, 0), cm2(nPts, 0), cgm(DIM*nPts, 0), cgm2(DIM*nPts, 0), cfs(nPts, 0), cdc(nPts, 0);
        std::vector<int> cns(nPts, 0), cng(nPts, 0), cwl(nPts, 0), csp(nPts, 0);
        zb_statistics cs{cm.data(), cm2.data(), cgm.data(), cgm2.data(), cfs.data(), cdc.data(), cns.data(), cng.data(), cwl.data(), csp.data()};
        for (int i = 0; i < nPts; i++)
            if (quantity[i] != ZB_SKIP) reduceChannel<DIM>(a, i, extraPde[c], cs, xRecSol.data() + (size_t)c*total, xRecFS.data() + (size_t)c*total, false);
        FILE *fc = fopen((dir + "/out_c" + std::to_string(c) + ".bin").c_str(), "wb");
        for (int i = 0; i < nPts; i++) { float rec[6] = {cm[i], cm2[i], (float)cns[i], cgm[DIM*i], cgm[DIM*i + 1], DIM == 3 ? cgm[DIM*i + DIM - 1] : 0.0f}; fwrite(rec, 4, 6, fc); }
        fclose(fc);
    }
    for (int i = 0; i < nPts; i++) steps += reducePoint<DIM>(a, i);
    printf("points %d walks %lld steps %lld status %d\n", nPts, total, steps, status);
    FILE *f = fopen((dir + "/out.bin").c_str(), "wb");
    for (int i = 0; i < nPts; i++) { float rec[6] = {mean[i], m2[i], (float)ns[i], (float)wl[i], dc[i], fsr[i]}; fwrite(rec, 4, 6, f); }
    fclose(f);
    return 0;
}

int main(int argc, char **argv)
{
    std::string dir = argc > 1 ? argv[1] : "/tmp/hs";
    auto hv = rd<Header>(dir + "/hdr.bin");
    return hv[0].dim == 2 ? run<2>(dir, hv[0]) : run<3>(dir, hv[0]);
}
