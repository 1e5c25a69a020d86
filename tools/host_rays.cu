// DEBUG / TEST TOOL, not part of libzombie_b200.so: the device ray, occlusion and sphere-sampling traversals
// (queries.cuh: intersectRay, sampleSphere) compiled for the host and run on a problem dumped by
// tests/test_host_walk.py — the per-query bodies of intersectRayKernel / occludedKernel / sampleSphereKernel (capi.cu).
//   nvcc -std=c++17 -O2 -Xcompiler -ffp-contract=off -o host_rays tools/host_rays.cu zombie_b200/csrc/wide_bvh.cpp
//   host_rays <dir> <dim>   reads <dir>/{pos,idx,o,d,tmax,sx,srad,su}.bin, writes <dir>/rays_out.bin, <dir>/hits_out.bin, <dir>/samples_out.bin
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../zombie_b200/csrc/queries.cuh"
#include "../zombie_b200/csrc/wide_bvh.h"

using namespace zb;

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
int run(const std::string& dir)
{
    auto pos = rd<float>(dir + "/pos.bin"); auto idx = rd<int>(dir + "/idx.bin");
    auto o = rd<float>(dir + "/o.bin"); auto d = rd<float>(dir + "/d.bin"); auto tmax = rd<float>(dir + "/tmax.bin");
    auto sx = rd<float>(dir + "/sx.bin"); auto srad = rd<float>(dir + "/srad.bin"); auto su = rd<float>(dir + "/su.bin");
    HostBvh B;
    const char *err = nullptr;
    ReflectingOptions opt;
    opt.enabled = true; // the sampling descent and the walk's rays run on reflecting boundaries
    if (!buildWideBvh(DIM, pos.data(), (int)pos.size()/DIM, idx.data(), (int)idx.size()/DIM, opt, B, &err)) { printf("%s\n", err); return 1; }
    BvhView v{};
    v.nodes = reinterpret_cast<const uint4 *>(B.nodes.data());
    v.prims = reinterpret_cast<const float4 *>(B.prims.data());
    v.primIndex = B.primIndex.data();
    v.cones = reinterpret_cast<const float4 *>(B.cones.data());
    v.adjacency = reinterpret_cast<const float4 *>(B.adjacency.data());
    v.nPrims = B.nPrims;
    v.snchSign = 1;
    int status = 0;

    // per ray: prim, t, pt[DIM], normal[DIM], occluded(o -> o + d tmax)   (as floats)
    int n = (int)tmax.size(), stride = 3 + 2*DIM;
    std::vector<float> out((size_t)stride*n);
    for (int i = 0; i < n; i++) {
        Vec<DIM> org = loadVec<DIM>(&o[DIM*i]), dir = loadVec<DIM>(&d[DIM*i]);
        RayHit<DIM> hit;
        bool found = intersectRay<DIM, false>(v, org, dir, tmax[i], hit, &status);
        float *r = &out[(size_t)stride*i];
        r[0] = found ? (float)v.primIndex[hit.prim] : -1.0f;
        r[1] = found ? hit.t : kMaxFloat;
        Vec<DIM> p = vzero<DIM>(), nrm = vzero<DIM>();
        if (found) {
            Prim<DIM> tri;
            tri.load(v, (uint32_t)hit.prim);
            p = tri.pointAt(hit.bv, hit.bw);
            nrm = unitNormal<DIM>(tri);
        }
        storeVec<DIM>(r + 2, p);
        storeVec<DIM>(r + 2 + DIM, nrm);
        Vec<DIM> pb = org + dir*tmax[i], dd = pb - org;
        float len = norm(dd);
        RayHit<DIM> occ;
        r[2 + 2*DIM] = intersectRay<DIM, true>(v, org, dd/len, len, occ, &status) ? 1.0f : 0.0f;
    }
    FILE *f = fopen((dir + "/rays_out.bin").c_str(), "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);

    // all-hits counts (countRayHits, the body of countRayHitsKernel / insideDomainKernel): per ray the number of distinct
    // hits within tmax, and of the axis-aligned ray o + t e_0, t < inf
    std::vector<int> hits(2*(size_t)n);
    for (int i = 0; i < n; i++) {
        Vec<DIM> org = loadVec<DIM>(&o[DIM*i]), dir = loadVec<DIM>(&d[DIM*i]), ex = vzero<DIM>();
        ex[0] = 1.0f;
        hits[2*i] = countRayHits<DIM>(v, org, dir, tmax[i], &status);
        hits[2*i + 1] = countRayHits<DIM>(v, org, ex, kMaxFloat, &status);
    }
    f = fopen((dir + "/hits_out.bin").c_str(), "wb");
    fwrite(hits.data(), 4, hits.size(), f);
    fclose(f);

    // per sample: prim, pdf, pt[DIM], normal[DIM]
    int m = (int)srad.size(), sstride = 2 + 2*DIM;
    std::vector<float> sout((size_t)sstride*m);
    for (int i = 0; i < m; i++) {
        Vec<DIM> q = loadVec<DIM>(&sx[DIM*i]), u = loadVec<DIM>(&su[DIM*i]);
        SphereSample<DIM> s;
        bool found = sampleSphere<DIM>(v, q, srad[i]*srad[i], u, s);
        float *r = &sout[(size_t)sstride*i];
        r[0] = found ? (float)v.primIndex[s.prim] : -1.0f;
        r[1] = found ? s.pdf : 0.0f;
        storeVec<DIM>(r + 2, s.p);
        storeVec<DIM>(r + 2 + DIM, s.n);
    }
    f = fopen((dir + "/samples_out.bin").c_str(), "wb");
    fwrite(sout.data(), 4, sout.size(), f);
    fclose(f);
    printf("rays %d samples %d status %d\n", n, m, status);
    return status;
}

int main(int argc, char **argv)
{
    if (argc < 3) { printf("usage: host_rays <dir> <dim>\n"); return 2; }
    return atoi(argv[2]) == 2 ? run<2>(argv[1]) : run<3>(argv[1]);
}
