// DEBUG / TEST TOOL, not part of libzombie_b200.so: the 2D instantiation of the warp-cooperative traversals
// (zombie_b200/csrc/coop.cuh, DIM = 2: line segments, silhouette vertices) on emulated warps (tools/warp_emu.h),
// checked against the one-thread-per-query traversals of queries.cuh. Companion of tools/coop_check.cu (3D).
//   nvcc -std=c++17 -O2 -DZB_EMULATE_WARP -Xcompiler -ffp-contract=off -Xcompiler -pthread -o /tmp/coop_check2d tools/coop_check2d.cu zombie_b200/csrc/wide_bvh.cpp
//   /tmp/coop_check2d [segments=4096] [nQueries=2048] [mu=1 (0: Neumann)]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../zombie_b200/csrc/rng.cuh"
#include "../zombie_b200/csrc/coop.cuh"
#include "../zombie_b200/csrc/wide_bvh.h"

using namespace zb;
namespace zb { long long g_coopCount[10] = {0}; }

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

int main(int argc, char **argv)
{
    int nSeg = argc > 1 ? atoi(argv[1]) : 4096;
    int nQ = argc > 2 ? atoi(argv[2]) : 2048;
    float mu = argc > 3 ? (float)atof(argv[3]) : 1.0f;
    nQ = (nQ + 31)/32*32;
    // star-shaped polygon r(theta) = 1 + 0.25 sin 5 theta; segments with centroid
    // y > 0 reflect, the others absorb
    const double kTwoPi = 6.283185307179586;
    std::vector<float> pos(2*(size_t)nSeg);
    for (int i = 0; i < nSeg; i++) {
        double th = kTwoPi*i/nSeg, r = 1.0 + 0.25*std::sin(5.0*th);
        pos[2*i] = (float)(r*std::cos(th)); pos[2*i + 1] = (float)(r*std::sin(th));
    }
    std::vector<int> segA, segR;
    for (int i = 0; i < nSeg; i++) {
        int j = (i + 1)%nSeg;
        auto& dst = 0.5f*(pos[2*i + 1] + pos[2*j + 1]) > 0.0f ? segR : segA;
        dst.push_back(i); dst.push_back(j);
    }
    HostBvh A, R;
    const char *err = nullptr;
    ReflectingOptions none, opt;
    opt.enabled = true;
    std::vector<float> coeff(segR.size()/2, mu);
    if (mu > 0) { opt.minCoeff = coeff.data(); opt.maxCoeff = coeff.data(); }
    if (!buildWideBvh(2, pos.data(), nSeg, segA.data(), (int)segA.size()/2, none, A, &err)) { printf("A: %s\n", err); return 1; }
    if (!buildWideBvh(2, pos.data(), nSeg, segR.data(), (int)segR.size()/2, opt, R, &err)) { printf("R: %s\n", err); return 1; }
    BvhView va = viewOf(A, 1), vr = viewOf(R, mu > 0 ? 1 : -1);
    printf("segments A %d R %d, nodes %zu %zu, %d queries\n", A.nPrims, R.nPrims, A.nodes.size(), R.nodes.size(), nQ);

    std::vector<float> x(2*(size_t)nQ);
    Pcg32 gen; gen.seed(13, 5);
    for (int i = 0; i < nQ; i++) {
        float a, b;
        do { a = 2*gen.nextFloat() - 1; b = 2*gen.nextFloat() - 1; } while (a*a + b*b > 1 || a*a + b*b < 1e-4f);
        float s = 0.7f;
        if (i%3 == 0) { // a third of the queries close to the boundary
            double th = std::atan2((double)b, (double)a), r = 1.0 + 0.25*std::sin(5.0*th);
            s = (float)((r - 0.002 - 0.05*gen.nextFloat())/std::sqrt((double)a*a + (double)b*b));
        }
        x[2*i] = s*a; x[2*i + 1] = s*b;
    }

    int status = 0;
    std::vector<ClosestHit<2>> seqHit(nQ), coopHit(nQ);
    std::vector<float> seqStar(nQ), coopStar(nQ), r2max(nQ);
    for (int i = 0; i < nQ; i++) {
        Vec<2> q = loadVec<2>(&x[2*i]);
        closestPoint<2>(va, q, kMaxFloat, seqHit[i], &status);
        r2max[i] = i%5 == 4 ? kMaxFloat : seqHit[i].d*seqHit[i].d;
        seqStar[i] = starRadius2<2>(vr, q, r2max[i], i%7 != 3, 1e-3f, &status);
    }
    int coopStatus = 0;
    for (int w0 = 0; w0 < nQ; w0 += 32) {
        zbemu::Warp warp;
        WarpShared ws;
        std::vector<std::thread> lanes;
        for (int l = 0; l < 32; l++) {
            lanes.emplace_back([&, l]() {
                zbemu::tlsWarp = &warp;
                zbemu::tlsLane = l;
                int i = w0 + l;
                Vec<2> q = loadVec<2>(&x[2*i]);
                bool has = (i%11) != 10;
                coopClosestPoint<2>(va, ws, has, q, kMaxFloat, coopHit[i], &coopStatus);
                if (!has) coopHit[i] = seqHit[i];
                float r2 = coopStarRadius2<2>(vr, ws, has, q, r2max[i], i%7 != 3, 1e-3f, &coopStatus);
                coopStar[i] = has ? r2 : seqStar[i];
            });
        }
        for (auto& t : lanes) t.join();
    }
    printf("sequential status %d\ncooperative status %d\n", status, coopStatus);
    int badDist = 0, primDiff = 0, badStar = 0, shrunk = 0;
    double worst = 0;
    for (int i = 0; i < nQ; i++) {
        if (coopHit[i].d != seqHit[i].d) badDist++;
        if (coopHit[i].prim != seqHit[i].prim) primDiff++;
        if (seqStar[i] < r2max[i]) shrunk++;
        double rel = std::fabs((double)coopStar[i] - seqStar[i])/std::max(1e-30, (double)std::max(coopStar[i], seqStar[i]));
        worst = std::max(worst, rel);
        if (rel > 2e-5) {
            if (badStar < 5) printf("  star mismatch q %d: seq %.9g coop %.9g r2max %.9g\n", i, seqStar[i], coopStar[i], r2max[i]);
            badStar++;
        }
    }
    printf("closest: distance mismatches %d, primitive differs (ties) %d | star: mismatches %d (worst rel %.3g), %d of %d queries shrunk below r2max\n",
           badDist, primDiff, badStar, worst, shrunk, nQ);
    return (badDist || badStar || status || coopStatus) ? 1 : 0;
}
