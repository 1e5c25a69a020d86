// DEBUG / TEST TOOL, not part of libzombie_b200.so: Reverse Walk on Stars with the DEVICE code compiled for the host —
// the reverse walks (walk.cuh in record-emitting mode), the per-sample contribution and the per-pair gather arithmetic
// of rws.cuh — on a problem dumped by tests/test_rws.py, so that the estimator can be checked against the reference
// without a GPU.
//   nvcc -std=c++17 -O2 -DZB_HOST_STATS -Xcompiler -ffp-contract=off -o host_rws tools/host_rws.cu zombie_b200/csrc/wide_bvh.cpp
//   host_rws <dir>   reads <dir>/{hdr,apos,aidx,rpos,ridx,spt,snormal,stype,spdf,ept}.bin, writes <dir>/totals.bin (nEval*6)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../zombie_b200/csrc/rws.cuh"
#include "../zombie_b200/csrc/wide_bvh.h"

using namespace zb;
namespace zb { long long g_zbClass[24] = {0}; void zbHostRecord(int, int, int) {} }

template <typename T> static std::vector<T> rd(const std::string& p)
{
    FILE *f = fopen(p.c_str(), "rb");
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

struct Header { int dim, kind, family, pad; float params[16]; float sigma, mu, normalOffset, radiusClamp; unsigned long long seed; };

template <int DIM>
int run(const std::string& dir, const Header& h)
{
    auto apos = rd<float>(dir + "/apos.bin"); auto aidx = rd<int>(dir + "/aidx.bin");
    auto rpos = rd<float>(dir + "/rpos.bin"); auto ridx = rd<int>(dir + "/ridx.bin");
    auto spt = rd<float>(dir + "/spt.bin"); auto snormal = rd<float>(dir + "/snormal.bin");
    auto stype = rd<int>(dir + "/stype.bin"); auto spdf = rd<float>(dir + "/spdf.bin");
    auto ept = rd<float>(dir + "/ept.bin");
    int n = (int)stype.size(), m = (int)ept.size()/DIM;
    HostBvh A, R;
    const char *err = nullptr;
    ReflectingOptions none, opt;
    opt.enabled = true;
    std::vector<float> coeff(ridx.size()/DIM, h.mu);
    if (h.kind == ZB_REFLECTING_ROBIN) { opt.minCoeff = coeff.data(); opt.maxCoeff = coeff.data(); }
    if (!buildWideBvh(DIM, apos.data(), (int)apos.size()/DIM, aidx.data(), (int)aidx.size()/DIM, none, A, &err)) { printf("A: %s\n", err); return 1; }
    if (!buildWideBvh(DIM, rpos.data(), (int)rpos.size()/DIM, ridx.data(), (int)ridx.size()/DIM, opt, R, &err)) { printf("R: %s\n", err); return 1; }

    int status = 0;
    SolveArgs a{};
    a.absorbing = viewOf(A, 1);
    a.reflecting = viewOf(R, h.kind == ZB_REFLECTING_NEUMANN && DIM == 3 ? -1 : 1);
    a.hasAbsorbing = a.hasReflecting = 1;
    a.reflectingKind = h.kind;
    a.estimator = ZB_WALK_ON_STARS;
    for (int k = 0; k < 3; k++) { a.domainMin[k] = -3; a.domainMax[k] = 3; }
    zb_walk_settings ws{1e-3f, 1e-3f, 1e-3f, 0.0f, kMaxFloat, 1024, 0, 1024, 0, 1, 1, 0, 0, 0, 0, 0};
    a.ws = ws;
    zb_pde_desc pde{};
    pde.absorptionCoeff = h.sigma;
    pde.areRobinConditionsPureNeumann = h.mu == 0;
    pde.areRobinCoeffsNonnegative = 1;
    pde.mode = ZB_PDE_MANUFACTURED;
    pde.family = h.family;
    for (int k = 0; k < 16; k++) pde.params[k] = h.params[k];
    pde.robinCoeff = h.mu;
    a.pde = pde;
    a.status = &status;

    std::vector<float> dA(n), dR(n), fsr(n, 0.0f), contribution(n);
    std::vector<int> quantity(n, ZB_SOLUTION), nWalks(n, 1);
    std::vector<uint64_t> rs(n), ri(n);
    a.pts = zb_sample_points{n, spt.data(), snormal.data(), nullptr, stype.data(), quantity.data(), dA.data(), dR.data(), nullptr,
                             nWalks.data(), rs.data(), ri.data(), fsr.data()};
    for (int i = 0; i < n; i++) {
        Pcg32 r; r.seed(h.seed, i); rs[i] = r.state; ri[i] = r.inc;
        ClosestHit<DIM> hit;
        closestPoint<DIM>(a.absorbing, loadVec<DIM>(&spt[DIM*i]), kMaxFloat, hit, &status); dA[i] = hit.d;
        contribution[i] = rwsSampleContribution<DIM>(a, i, hit);
        closestPoint<DIM>(a.reflecting, loadVec<DIM>(&spt[DIM*i]), kMaxFloat, hit, &status); dR[i] = hit.d;
    }
    // reverse walks: boundary / source terms off, one splat record per step
    a.ws.ignoreReflectingBoundaryContribution = 1;
    a.ws.ignoreSourceContribution = 1;
    long long capacity = (long long)n*1100;
    std::vector<float4> records(3*(size_t)capacity);
    unsigned long long count = 0;
    a.rwsRecords = records.data(); a.rwsCount = &count; a.rwsCapacity = capacity;
    a.rwsPdf = spdf.data(); a.rwsContribution = contribution.data(); a.rwsNormalOffset = h.normalOffset;
    std::vector<int> taskCount(n), stratCount(n);
    std::vector<long long> taskOffset(n + 1, 0), stratOffset(n + 1, 0);
    std::vector<unsigned long long> pointState(n);
    a.taskCount = taskCount.data(); a.taskOffset = taskOffset.data(); a.stratCount = stratCount.data(); a.stratOffset = stratOffset.data();
    a.pointState = pointState.data();
    for (int i = 0; i < n; i++) countTasksPoint<DIM>(a, i);
    for (int i = 0; i < n; i++) { taskOffset[i+1] = taskOffset[i] + taskCount[i]; stratOffset[i+1] = stratOffset[i] + stratCount[i]; }
    std::vector<float> strat(1);
    a.strat = strat.data();
    long long total = taskOffset[n];
    a.totalTasks = total;
    std::vector<float> recSol(total);
    std::vector<int> recLen(total), recFlags(total);
    a.recSolution = recSol.data(); a.recLength = recLen.data(); a.recFlags = recFlags.data(); a.anyGradient = 0;
    for (int i = 0; i < n; i++) preparePoint<DIM>(a, i);
    long long steps = 0;
    for (long long t = 0; t < total; t++) {
        Lane<DIM> L; L.active = false;
        WalkState<DIM> queue[kSplitStack];
        beginTask<DIM, true>(a, L, t);
        while (L.active) { StepResult r = walkStep<DIM, true>(a, L, queue); if (r != kContinue) finishState<DIM, true>(a, L, r, queue); }
        steps += recLen[t];
    }
    // gather: every record onto every evaluation point
    RwsGatherArgs g{};
    g.reflecting = a.reflecting; g.hasReflecting = 1;
    g.records = records.data(); g.nRecords = (long long)count; g.nEval = m;
    g.sqrtLambda = h.sigma > 0 ? std::sqrt(h.sigma) : 0.0f;
    g.radiusClamp = h.radiusClamp; g.kernelRegularization = 0.0f;
    g.useSelfNormalization = h.sigma == 0.0f && h.mu == 0.0f ? 1 : 0;
    g.status = &status;
    std::vector<float> totals(6*(size_t)m, 0.0f);
    Vec<DIM> zero = vzero<DIM>();
    for (int j = 0; j < m; j++) {
        Vec<DIM> e = loadVec<DIM>(&ept[DIM*j]);
        float sums[6] = {0, 0, 0, 0, 0, 0};
        for (unsigned long long r = 0; r < count; r++) rwsAccumulate<DIM>(g, records[3*r], records[3*r + 1], records[3*r + 2], e, zero, false, sums);
        for (int k = 0; k < 6; k++) totals[6*(size_t)j + k] = sums[k];
    }
    FILE *f = fopen((dir + "/totals.bin").c_str(), "wb");
    fwrite(totals.data(), 4, totals.size(), f);
    fclose(f);
    printf("samples %d walks %lld steps %lld records %llu status %d\n", n, total, steps, count, status);
    return status;
}

int main(int argc, char **argv)
{
    std::string dir = argc > 1 ? argv[1] : "/tmp/hs";
    auto hv = rd<Header>(dir + "/hdr.bin");
    return hv[0].dim == 2 ? run<2>(dir, hv[0]) : run<3>(dir, hv[0]);
}
