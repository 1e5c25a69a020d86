// DEBUG / TEST TOOL, not part of libzombie_b200.so: the on-device hierarchy builder (zombie_b200/csrc/bvh_build.cuh)
// run on the CPU — the same per-item functions the kernels call, driven by loops, std::stable_sort in place of the
// cub sorts — and checked against the host SAH builder (wide_bvh.cpp) on a mesh dumped by tests/test_host_logic.py:
//   * soundness of the built hierarchy (validateWideBvh);
//   * per-primitive records (vertex record, adjacency normals / flags / coefficients) bit-identical to the host builder's;
//   * closest point (distance, point, tie-broken index), first ray hit and star radius at the query points identical over
//     both trees; nodes / primitives visited by each (traversal cost of the Morton tree relative to the SAH tree).
//   nvcc -std=c++17 -O2 -DZB_HOST_STATS -Xcompiler -ffp-contract=off -o host_build tools/host_build.cu zombie_b200/csrc/wide_bvh.cpp
//   host_build <dir> <dim> <kind: 0 absorbing, 1 Neumann, 2 Robin> <mu> <doubleSided>   reads <dir>/{rpos,ridx,pts}.bin
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../zombie_b200/csrc/walk.cuh"
#include "../zombie_b200/csrc/wide_bvh.h"
#include "../zombie_b200/csrc/bvh_build.cuh"

using namespace zb;
static long long g_visits[2][5][2]; // [tree][query][nodes | prims]
static int g_tree = 0;
namespace zb { long long g_zbClass[24] = {0}; void zbHostRecord(int q, int nodes, int prims) { g_visits[g_tree][q][0] += nodes; g_visits[g_tree][q][1] += prims; } }

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

// bvh_build.cu's buildBvhOnDevice, one loop per kernel
static bool emulatedBuild(int dim, const std::vector<float>& pos, const std::vector<int>& idx, bool refl, const float *minCoeff,
                          const float *maxCoeff, bool doubleSided, float precision, HostBvh& out, std::vector<int>& levelStart)
{
    using namespace lbvh;
    const int n = (int)idx.size()/dim, nVert = (int)pos.size()/dim;
    Mesh mesh{dim, n, nVert, pos.data(), idx.data()};
    std::vector<float> cen((size_t)3*n);
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; i++) {
        primCentroid(mesh, i, &cen[3*(size_t)i]);
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], cen[3*(size_t)i + k]); hi[k] = std::max(hi[k], cen[3*(size_t)i + k]); }
    }
    float inv[3];
    for (int k = 0; k < 3; k++) inv[k] = hi[k] - lo[k] > 0.0f ? 1.0f/(hi[k] - lo[k]) : 0.0f;
    std::vector<uint64_t> keysIn(n), keys(n);
    for (int i = 0; i < n; i++) keysIn[i] = mortonKey(dim, &cen[3*(size_t)i], lo, inv);
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return keysIn[a] < keysIn[b]; });
    for (int i = 0; i < n; i++) keys[i] = keysIn[order[i]];

    const int cap = n/4 + 2;
    std::vector<TopoNode> topo(cap);
    std::vector<int> leafPos(n, -1);
    topo[0] = TopoNode{};
    topo[0].start = 0; topo[0].end = n;
    levelStart.assign(1, 0);
    int first = 0, count = 1, primCursor = 0, leafChildren = 0;
    while (count > 0) {
        if (first + count > cap) { printf("node capacity exceeded\n"); return false; }
        std::vector<int> ni(count), nl(count);
        for (int k = 0; k < count; k++) splitNode(keys.data(), topo[first + k], ni[k], nl[k]);
        int io = 0, lo2 = 0;
        for (int k = 0; k < count; k++) {
            placeChildren(topo.data(), first + k, first + count + io, primCursor + lo2, leafPos.data());
            leafChildren += topo[first + k].nChildren - ni[k];
            io += ni[k]; lo2 += nl[k];
        }
        first += count;
        levelStart.push_back(first);
        count = io;
        primCursor += lo2;
    }
    const int nNodes = first;
    if (primCursor != n) { printf("primitives lost\n"); return false; }

    std::vector<PrimAdjacency> adjInput;
    if (refl) {
        adjInput.resize(n);
        std::vector<int> opposite, startsAt, endsAt;
        if (dim == 3) {
            const int nHalf = 3*n;
            std::vector<uint64_t> hk(nHalf), hks(nHalf);
            std::vector<uint32_t> hv(nHalf), hvs(nHalf);
            for (int he = 0; he < nHalf; he++) halfEdge(mesh, he, hk[he], hv[he]);
            std::vector<int> perm(nHalf);
            std::iota(perm.begin(), perm.end(), 0);
            std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return hk[a] < hk[b]; });
            for (int s = 0; s < nHalf; s++) { hks[s] = hk[perm[s]]; hvs[s] = hv[perm[s]]; }
            opposite.assign(nHalf, -2);
            for (int s = 0; s < nHalf; s++) oppositeFace(hks.data(), hvs.data(), nHalf, s, opposite.data());
        } else {
            startsAt.assign(nVert, -1); endsAt.assign(nVert, -1);
            for (int i = 0; i < n; i++) {
                startsAt[idx[2*(size_t)i]] = std::max(startsAt[idx[2*(size_t)i]], i);
                endsAt[idx[2*(size_t)i + 1]] = std::max(endsAt[idx[2*(size_t)i + 1]], i);
            }
        }
        for (int i = 0; i < n; i++)
            primAdjacency(mesh, i, opposite.data(), startsAt.data(), endsAt.data(), minCoeff, maxCoeff, doubleSided, precision, adjInput[i]);
    }

    out = HostBvh();
    out.dim = dim;
    out.nPrims = n;
    out.nodes.resize(nNodes);
    out.prims.assign((size_t)(dim == 3 ? 12 : 4)*n, 0.0f);
    out.primIndex.assign(n, -1);
    if (refl) { out.adjacency.resize(n); out.cones.resize((size_t)nNodes*kWidth); }
    for (int s = 0; s < n; s++)
        scatterPrim(mesh, order[s], leafPos[s], refl ? adjInput.data() : nullptr, out.prims.data(), out.primIndex.data(),
                    refl ? out.adjacency.data() : nullptr);
    std::vector<Box3> nodeBox(nNodes);
    for (int l = (int)levelStart.size() - 2; l >= 0; l--)
        for (int w = levelStart[l]; w < levelStart[l + 1]; w++) writeNode(dim, topo.data(), w, out.prims.data(), nodeBox.data(), out.nodes.data());
    if (refl) {
        for (int w = 0; w < nNodes; w++) {
            for (int c = 0; c < kWidth; c++) {
                ChildCone cone = emptyCone();
                if (c < topo[w].nChildren) {
                    float blo[3], bhi[3], centre[3];
                    decodeChild(out.nodes[w], c, blo, bhi);
                    for (int k = 0; k < 3; k++) centre[k] = 0.5f*(blo[k] + bhi[k]);
                    // 32 interleaved partial sums merged by a butterfly, like the warp of coneLevelKernel
                    ConeSums lanes[32];
                    for (int l = 0; l < 32; l++) coneInit(lanes[l]);
                    for (int i = topo[w].bounds[c]; i < topo[w].bounds[c + 1]; i++)
                        coneAccumulate(mesh, adjInput.data(), order[i], centre, lanes[(i - topo[w].bounds[c]) & 31]);
                    for (int o = 16; o > 0; o >>= 1) {
                        ConeSums next[32];
                        for (int l = 0; l < 32; l++) { next[l] = lanes[l]; coneMerge(next[l], lanes[l ^ o]); }
                        for (int l = 0; l < 32; l++) lanes[l] = next[l];
                    }
                    float axis[3];
                    if (coneFromSums(lanes[0], cone, axis)) {
                        float half = 0.0f;
                        for (int i = topo[w].bounds[c]; i < topo[w].bounds[c + 1]; i++)
                            half = std::max(half, coneHalfAngle(mesh, adjInput.data(), order[i], axis));
                        coneFinish(cone, axis, half);
                    }
                }
                out.cones[(size_t)w*kWidth + c] = cone;
            }
        }
    }
    out.nLeaves = leafChildren;
    out.maxDepth = (int)levelStart.size() - 1;
    for (int k = 0; k < 3; k++) { out.boxMin[k] = nodeBox[0].lo[k]; out.boxMax[k] = nodeBox[0].hi[k]; }
    return true;
}

static BvhView viewOf(const HostBvh& h, int snchSign)
{
    BvhView v{};
    v.nodes = reinterpret_cast<const uint4 *>(h.nodes.data());
    v.prims = reinterpret_cast<const float4 *>(h.prims.data());
    v.primIndex = h.primIndex.data();
    v.cones = h.cones.empty() ? nullptr : reinterpret_cast<const float4 *>(h.cones.data());
    v.adjacency = h.adjacency.empty() ? nullptr : reinterpret_cast<const float4 *>(h.adjacency.data());
    v.nPrims = h.nPrims;
    v.snchSign = snchSign;
    return v;
}

template <int DIM>
int run(const std::string& dir, int kind, float mu, bool doubleSided)
{
    auto pos = rd<float>(dir + "/rpos.bin"); auto idx = rd<int>(dir + "/ridx.bin");
    auto pt = rd<float>(dir + "/pts.bin");
    const int nPrim = (int)idx.size()/DIM;
    const bool refl = kind != 0;
    std::vector<float> cmin(nPrim), cmax(nPrim);
    for (int i = 0; i < nPrim; i++) { cmin[i] = mu*(0.5f + 0.5f*(float)((i*7919)%13)/13.0f); cmax[i] = cmin[i]*1.5f; }
    ReflectingOptions opt;
    opt.enabled = refl;
    opt.solveDoubleSided = doubleSided;
    if (kind == ZB_REFLECTING_ROBIN) { opt.minCoeff = cmin.data(); opt.maxCoeff = cmax.data(); }
    HostBvh A, B;
    const char *err = nullptr;
    if (!buildWideBvh(DIM, pos.data(), (int)pos.size()/DIM, idx.data(), nPrim, opt, A, &err)) { printf("%s\n", err); return 1; }
    std::vector<int> levels;
    if (!emulatedBuild(DIM, pos, idx, refl, opt.minCoeff, opt.maxCoeff, doubleSided, opt.silhouettePrecision, B, levels)) return 1;

    long long violations = validateWideBvh(B, DIM, pos.data(), idx.data());
    // records, matched through the caller's primitive index
    std::vector<int> slotA(nPrim, -1);
    for (int s = 0; s < nPrim; s++) slotA[A.primIndex[s]] = s;
    const int pf = DIM == 3 ? 12 : 4;
    long long recordMismatch = 0, flagMismatch = 0;
    for (int s = 0; s < nPrim; s++) {
        int p = B.primIndex[s];
        if (p < 0 || p >= nPrim) { recordMismatch++; continue; }
        int a = slotA[p];
        if (memcmp(&A.prims[(size_t)pf*a], &B.prims[(size_t)pf*s], pf*sizeof(float)) != 0) recordMismatch++;
        if (refl) {
            if (memcmp(&A.adjacency[a], &B.adjacency[s], sizeof(PrimAdjacency)) != 0) recordMismatch++;
            if (A.adjacency[a].flags != B.adjacency[s].flags) flagMismatch++;
        }
    }

    const int snch = kind == ZB_REFLECTING_NEUMANN && DIM == 3 ? -1 : 1;
    BvhView views[2] = {viewOf(A, snch), viewOf(B, snch)};
    int n = (int)pt.size()/DIM;
    long long closestMismatch = 0, rayMismatch = 0, starMismatch = 0, starFastMismatch = 0;
    int status = 0;
    Pcg32 rng;
    rng.seed(11u, 5u);
    for (int i = 0; i < n; i++) {
        Vec<DIM> x = loadVec<DIM>(&pt[DIM*i]);
        Vec<DIM> dir = uniformDirection<DIM>(rng);
        float cd[2], ct[2], star[2], starFast[2];
        int cprim[2], rprim[2];
        Vec<DIM> cp[2];
        for (int t = 0; t < 2; t++) {
            g_tree = t;
            const BvhView& v = views[t];
            ClosestHit<DIM> hit;
            closestPoint<DIM>(v, x, kMaxFloat, hit, &status);
            cd[t] = hit.d;
            closestTieBreak<DIM>(v, x, hit, &status);
            cprim[t] = hit.prim >= 0 ? v.primIndex[hit.prim] : -1;
            cp[t] = hit.p;
            RayHit<DIM> rh;
            bool got = intersectRay<DIM, false>(v, x, dir, kMaxFloat, rh, &status);
            ct[t] = got ? rh.t : -1.0f;
            rprim[t] = got ? v.primIndex[rh.prim] : -1;
            star[t] = starFast[t] = 0.0f;
            if (refl) {
                SolveArgs a{};
                a.reflecting = v; a.hasReflecting = 1; a.reflectingKind = kind;
                a.status = &status;
                star[t] = starRadiusQuery<DIM>(a, x, 1e-3f, kMaxFloat, 1e-3f, false);
                float r2 = kMaxFloat;
                uint2 stack[kStackSize];
                int sp = 0;
                stack[sp++] = make_uint2(0u, 0u);
                starRadiusDfs<DIM, true>(v, x, r2, true, 1e-3f, stack, sp, 0x7fffffff, &status);
                starFast[t] = r2;
            }
        }
        if (cd[0] != cd[1] || cprim[0] != cprim[1] || memcmp(&cp[0], &cp[1], sizeof(Vec<DIM>)) != 0) closestMismatch++;
        if (ct[0] != ct[1]) rayMismatch++;
        (void)rprim;
        if (star[0] != star[1]) starMismatch++;
        if (fabsf(starFast[0] - starFast[1]) > 1e-6f*fabsf(starFast[0])) starFastMismatch++;
    }
    printf("nodes %zu %zu leaves %d %d depth %d %d levels %zu\n", A.nodes.size(), B.nodes.size(), A.nLeaves, B.nLeaves, A.maxDepth, B.maxDepth,
           levels.size() - 1);
    printf("violations %lld record_mismatch %lld flag_mismatch %lld closest_mismatch %lld ray_mismatch %lld star_mismatch %lld star_fast_mismatch %lld status %d\n",
           violations, recordMismatch, flagMismatch, closestMismatch, rayMismatch, starMismatch, starFastMismatch, status);
    const char *names[5] = {"closest", "ray", "occlusion", "sample", "star"};
    for (int q = 0; q < 5; q++) {
        if (g_visits[0][q][0] == 0) continue;
        printf("visits %s sah %.1f nodes %.1f prims | morton %.1f nodes %.1f prims per query\n", names[q], (double)g_visits[0][q][0]/n,
               (double)g_visits[0][q][1]/n, (double)g_visits[1][q][0]/n, (double)g_visits[1][q][1]/n);
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 6) { printf("usage: host_build <dir> <dim> <kind> <mu> <doubleSided>\n"); return 2; }
    int dim = atoi(argv[2]), kind = atoi(argv[3]);
    float mu = (float)atof(argv[4]);
    bool ds = atoi(argv[5]) != 0;
    return dim == 2 ? run<2>(argv[1], kind, mu, ds) : run<3>(argv[1], kind, mu, ds);
}
