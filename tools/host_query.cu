// DEBUG / TEST TOOL, not part of libzombie_b200.so: the device query code (queries.cuh, walk.cuh: starRadiusQuery)
// compiled for the host, run on a problem dumped by tests/test_host_walk.py. Writes, per query point, the distance
// to the reflecting boundary and the star radius the walk would use (both the one-thread traversal with the
// reference's trigonometric node test and the one with the algebraic node test + fast primitive test).
//   nvcc -std=c++17 -O2 -DZB_HOST_STATS -Xcompiler -ffp-contract=off -o host_query tools/host_query.cu zombie_b200/csrc/wide_bvh.cpp
//   host_query <dir> <dim> <kind: 1 Neumann, 2 Robin> <mu>    reads <dir>/{rpos,ridx,pts,maxr}.bin, writes <dir>/query_out.bin
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../zombie_b200/csrc/walk.cuh"
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

template <int DIM>
int run(const std::string& dir, int kind, float mu)
{
    auto rpos = rd<float>(dir + "/rpos.bin"); auto ridx = rd<int>(dir + "/ridx.bin");
    auto pt = rd<float>(dir + "/pts.bin"); auto maxr = rd<float>(dir + "/maxr.bin");
    HostBvh R;
    const char *err = nullptr;
    ReflectingOptions opt;
    opt.enabled = true;
    std::vector<float> coeff(ridx.size()/DIM, mu);
    if (kind == ZB_REFLECTING_ROBIN) { opt.minCoeff = coeff.data(); opt.maxCoeff = coeff.data(); }
    if (!buildWideBvh(DIM, rpos.data(), (int)rpos.size()/DIM, ridx.data(), (int)ridx.size()/DIM, opt, R, &err)) { printf("%s\n", err); return 1; }
    BvhView v{};
    v.nodes = reinterpret_cast<const uint4 *>(R.nodes.data());
    v.prims = reinterpret_cast<const float4 *>(R.prims.data());
    v.primIndex = R.primIndex.data();
    v.cones = reinterpret_cast<const float4 *>(R.cones.data());
    v.adjacency = reinterpret_cast<const float4 *>(R.adjacency.data());
    v.nPrims = R.nPrims;
    v.snchSign = kind == ZB_REFLECTING_NEUMANN && DIM == 3 ? -1 : 1;
    SolveArgs a{};
    a.reflecting = v; a.hasReflecting = 1; a.reflectingKind = kind;
    int status = 0;
    a.status = &status;
    int n = (int)pt.size()/DIM;
    std::vector<float> out(3*(size_t)n);
    std::vector<int> prim(n);
    for (int i = 0; i < n; i++) {
        Vec<DIM> x = loadVec<DIM>(&pt[DIM*i]);
        ClosestHit<DIM> hit;
        closestPoint<DIM>(v, x, kMaxFloat, hit, &status);
        out[3*i] = hit.d;
        closestTieBreak<DIM>(v, x, hit, &status); // the index rule of the batched entry point (largest tied input index)
        prim[i] = hit.prim >= 0 && hit.d == out[3*i] ? v.primIndex[hit.prim] : -1;
        out[3*i + 1] = starRadiusQuery<DIM>(a, x, 1e-3f, maxr[i], 1e-3f, false);
        // the traversal the kernels run in their private phase: algebraic node test, fast primitive test
        float r2max = maxr[i] < kMaxFloat ? maxr[i]*maxr[i] : kMaxFloat, r2 = r2max;
        uint2 stack[kStackSize];
        int sp = 0;
        stack[sp++] = make_uint2(0u, 0u);
        starRadiusDfs<DIM, true>(v, x, r2, true, 1e-3f, stack, sp, 0x7fffffff, &status);
        out[3*i + 2] = starQueryResult(a, r2, r2max, 1e-3f, maxr[i]);
    }
    FILE *f = fopen((dir + "/query_out.bin").c_str(), "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    f = fopen((dir + "/query_prim.bin").c_str(), "wb");
    fwrite(prim.data(), 4, prim.size(), f);
    fclose(f);
    printf("queries %d status %d\n", n, status);
    return status;
}

int main(int argc, char **argv)
{
    if (argc < 5) { printf("usage: host_query <dir> <dim> <kind> <mu>\n"); return 2; }
    int dim = atoi(argv[2]), kind = atoi(argv[3]);
    float mu = (float)atof(argv[4]);
    return dim == 2 ? run<2>(argv[1], kind, mu) : run<3>(argv[1], kind, mu);
}
