// DEBUG / TEST TOOL, not part of libzombie_b200.so: runs the warp-cooperative traversals of
// zombie_b200/csrc/coop.cuh on a CPU (32 host threads per emulated warp, tools/warp_emu.h) and checks them
// against the one-thread-per-query traversals of queries.cuh on the same wide BVH.
//   nvcc -std=c++17 -O2 -DZB_EMULATE_WARP -Xcompiler -ffp-contract=off -Xcompiler -pthread -o /tmp/coop_check tools/coop_check.cu zombie_b200/csrc/wide_bvh.cpp
//   /tmp/coop_check [subdiv=6] [nQueries=2048] [mu=1 (0: Neumann)]
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

// ZB_TRACE_QUERY=<index>: prints what lowers that query's radius in the pooled phase (the trace hook of coop.cuh)
static int g_traceWarp = -1, g_traceLane = -1;
static thread_local int tlsWarpIndex = -1;
#define ZB_COOP_TRACE(owner, what, value, id) if (tlsWarpIndex == g_traceWarp && (owner) == g_traceLane) printf("    trace: %s: r2 %.9g id %u\n", what, (double)(value), (unsigned)(id));
#include "../zombie_b200/csrc/rng.cuh"
#include "../zombie_b200/csrc/coop.cuh"
#include "../zombie_b200/csrc/wide_bvh.h"
#include "probe_mesh.h"

using namespace zb;
namespace zb { long long g_coopCount[10] = {0}; }
#if defined(ZB_HOST_STATS)
static long long g_seq[5][2] = {{0}};
namespace zb { long long g_zbClass[24] = {0}; void zbHostRecord(int q, int nodes, int prims) { __atomic_fetch_add(&g_seq[q][0], (long long)nodes, __ATOMIC_RELAXED); __atomic_fetch_add(&g_seq[q][1], (long long)prims, __ATOMIC_RELAXED); } }
#endif

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
    int subdiv = argc > 1 ? atoi(argv[1]) : 6;
    int nQ = argc > 2 ? atoi(argv[2]) : 2048;
    float mu = argc > 3 ? (float)atof(argv[3]) : 1.0f;
    nQ = (nQ + 31)/32*32;
    if (const char *t = getenv("ZB_TRACE_QUERY")) { g_traceWarp = atoi(t)/32; g_traceLane = atoi(t)%32; }
    std::vector<float> pos; std::vector<int> tri;
    icosphere(subdiv, pos, tri);
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
    BvhView va = viewOf(A, 1), vr = viewOf(R, mu > 0 ? 1 : -1);
    printf("tris A %d R %d, nodes %zu %zu, depth %d %d, %d queries\n", A.nPrims, R.nPrims, A.nodes.size(), R.nodes.size(), A.maxDepth, R.maxDepth, nQ);

    // queries: points in the ball |x| < 0.95 (a third of them pulled close to the boundary)
    std::vector<float> x(3*(size_t)nQ);
    Pcg32 gen; gen.seed(11, 3);
    for (int i = 0; i < nQ; i++) {
        float a, b, c;
        do { a = 2*gen.nextFloat()-1; b = 2*gen.nextFloat()-1; c = 2*gen.nextFloat()-1; } while (a*a+b*b+c*c > 1 || a*a+b*b+c*c < 1e-4f);
        float s = 0.95f;
        if (i%3 == 0) { float l = std::sqrt(a*a+b*b+c*c); s = (0.8f + 0.04f*gen.nextFloat())/l; }
        x[3*i] = s*a; x[3*i+1] = s*b; x[3*i+2] = s*c;
    }

    // ---- reference: one thread per query
    int status = 0;
    std::vector<ClosestHit<3>> seqHit(nQ);
    std::vector<float> seqStar(nQ), r2max(nQ);
    for (int i = 0; i < nQ; i++) {
        Vec<3> q = loadVec<3>(&x[3*i]);
        closestPoint<3>(va, q, kMaxFloat, seqHit[i], &status);
        r2max[i] = i%5 == 4 ? kMaxFloat : seqHit[i].d*seqHit[i].d; // the walk searches within the distance to the absorbing boundary
        seqStar[i] = starRadius2<3>(vr, q, r2max[i], i%7 != 3, 1e-3f, &status);
    }
    printf("sequential status %d\n", status);
#if defined(ZB_HOST_STATS)
    printf("sequential work per query: closest %.1f nodes %.1f prims | star %.1f nodes %.1f prims\n", g_seq[0][0]/(double)nQ, g_seq[0][1]/(double)nQ, g_seq[4][0]/(double)nQ, g_seq[4][1]/(double)nQ);
    for (int q = 0; q < 5; q++) g_seq[q][0] = g_seq[q][1] = 0;
#endif

    // ---- emulated warps
    std::vector<ClosestHit<3>> coopHit(nQ);
    std::vector<float> coopStar(nQ);
    int coopStatus = 0;
    for (int w0 = 0; w0 < nQ; w0 += 32) {
        zbemu::Warp warp;
        WarpShared ws;
        std::vector<std::thread> lanes;
        for (int l = 0; l < 32; l++) {
            lanes.emplace_back([&, l]() {
                zbemu::tlsWarp = &warp;
                zbemu::tlsLane = l;
                tlsWarpIndex = w0/32;
                int i = w0 + l;
                Vec<3> q = loadVec<3>(&x[3*i]);
                bool has = (i%11) != 10; // some lanes sit the query out
                coopClosestPoint<3>(va, ws, has, q, kMaxFloat, coopHit[i], &coopStatus);
                if (!has) coopHit[i] = seqHit[i];
                float r2 = coopStarRadius2<3>(vr, ws, has, q, r2max[i], i%7 != 3, 1e-3f, &coopStatus);
                coopStar[i] = has ? r2 : seqStar[i];
            });
        }
        for (auto& t : lanes) t.join();
    }
    printf("cooperative status %d\n", coopStatus);
#if defined(ZB_HOST_STATS)
    printf("phase-1 (private DFS) work per query: closest %.1f nodes %.1f prims | star %.1f nodes %.1f prims\n", g_seq[0][0]/(double)nQ, g_seq[0][1]/(double)nQ, g_seq[4][0]/(double)nQ, g_seq[4][1]/(double)nQ);
#endif
    printf("coop work per query: closest %.1f node items %.1f prims (%.1f + %.1f rounds/warp) | star %.1f node items %.1f prims (%.1f + %.1f rounds/warp)\n", g_coopCount[0]/(double)nQ, g_coopCount[1]/(double)nQ, g_coopCount[4]*32.0/nQ, g_coopCount[5]*32.0/nQ, g_coopCount[2]/(double)nQ, g_coopCount[3]/(double)nQ, g_coopCount[6]*32.0/nQ, g_coopCount[7]*32.0/nQ);
    printf("star filter stage: %.1f prims per query in %.1f filter rounds/warp\n", g_coopCount[8]/(double)nQ, g_coopCount[9]*32.0/nQ);

    int badDist = 0, primDiff = 0, badStar = 0, shrunk = 0;
    double worstStar = 0;
    for (int i = 0; i < nQ; i++) {
        if (coopHit[i].d != seqHit[i].d) badDist++;
        if (coopHit[i].prim != seqHit[i].prim) primDiff++;
        if (seqStar[i] < r2max[i]) shrunk++;
        double rel = std::fabs((double)coopStar[i] - seqStar[i])/std::max(1e-30, (double)std::max(coopStar[i], seqStar[i]));
        worstStar = std::max(worstStar, rel);
        if (rel > 2e-5) {
            if (badStar < 5) {
                // which part differs: the traversal (coop vs one thread) or the arithmetic (fast vs reference-ordered)?
                uint2 stack[kStackSize]; int sp = 0; stack[sp++] = make_uint2(0u, 0u);
                float r2f = r2max[i];
                Vec<3> q = loadVec<3>(&x[3*i]);
                starRadiusDfs<3, true>(vr, q, r2f, i%7 != 3, 1e-3f, stack, sp, 0x7fffffff, &status);
                // brute force over all primitives with both primitive tests
                float be = r2max[i], bf = r2max[i]; int ie = -1, jf = -1;
                for (int p = 0; p < vr.nPrims; p++) {
                    float a = be, b = bf;
                    primStarRadius2<3>(vr, p, q, a, i%7 != 3, 1e-3f, true); if (a < be) { be = a; ie = p; }
                    primStarRadius2Fast<3>(vr, p, q, b, i%7 != 3, 1e-3f, true); if (b < bf) { bf = b; jf = p; }
                }
                // does some primitive report less than both results when it is evaluated with another incoming radius?
                // (the reference's reflectance bound is order dependent on near-coplanar primitives)
                {
                    const float target = 0.99999f*std::max(coopStar[i], seqStar[i]);
                    if (const char *tp = getenv("ZB_TRACE_PRIM")) {
                        int p = atoi(tp); float inc = (float)atof(getenv("ZB_TRACE_R2"));
                        for (int sil = 0; sil < 2; sil++) {
                            float a = inc, b = inc;
                            primStarRadius2<3>(vr, p, q, a, i%7 != 3, 1e-3f, sil != 0);
                            primStarRadius2Fast<3>(vr, p, q, b, i%7 != 3, 1e-3f, sil != 0);
                            float c = primStarRadius2Call<3>(vr.prims, vr.adjacency, vr.snchSign, p, q, inc, i%7 != 3, 1e-3f, sil != 0).x;
                            printf("    prim %d incoming %.9g sil %d: exact %.9g fast %.9g call %.9g needsFull %d\n", p, inc, sil, a, b, c,
                                   (int)primStarNeedsFullTest<3>(vr, p, q, inc, sil != 0));
                        }
                    }
                    float incoming[43] = {r2max[i], seqStar[i], coopStar[i]};
                    for (int k = 0; k < 40; k++) incoming[3 + k] = std::min(r2max[i], 0.98f*std::min(coopStar[i], seqStar[i])*std::pow(1.08f, (float)k));
                    int shown = 0;
                    for (int p = 0; p < vr.nPrims && shown < 4; p++) {
                        for (int k = 0; k < 43; k++) {
                            for (int sil = 0; sil < 2; sil++) {
                                float a = incoming[k], b = incoming[k];
                                primStarRadius2<3>(vr, p, q, a, i%7 != 3, 1e-3f, sil != 0);
                                primStarRadius2Fast<3>(vr, p, q, b, i%7 != 3, 1e-3f, sil != 0);
                                if ((a < target && a < 0.99999f*incoming[k]) || (b < target && b < 0.99999f*incoming[k])) {
                                    printf("    prim %d with incoming %.9g, silhouette tests %d: exact %.9g fast %.9g\n", p, incoming[k], sil, a, b); shown++;
                                }
                            }
                        }
                    }
                }
                printf("  star mismatch q %d: seq %.9g coop %.9g (one-thread fast %.9g) r2max %.9g | brute exact %.9g (prim %d) brute fast %.9g (prim %d)\n",
                       i, seqStar[i], coopStar[i], r2f, r2max[i], be, ie, bf, jf);
            }
            badStar++;
        }
    }
    printf("closest: distance mismatches %d, primitive differs (ties) %d | star: mismatches %d (worst rel %.3g), %d of %d queries shrunk below r2max\n",
           badDist, primDiff, badStar, worstStar, shrunk, nQ);
    return (badDist || badStar || status || coopStatus) ? 1 : 0;
}
