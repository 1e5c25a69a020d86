// DEBUG TOOL (not part of the library): one star-radius query on the DEVICE, three ways — brute force over all primitives
// with the reference-ordered and the fast primitive test, the one-thread depth-first traversal, and the warp-cooperative
// query — to localise device-only deviations (FMA contraction, approximate intrinsics) the host-side checks cannot see.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -o tools/gpu_star_debug tools/gpu_star_debug.cu zombie_b200/csrc/wide_bvh.cpp
//   tools/gpu_star_debug <dir with rpos.bin ridx.bin> <mu> <x> <y> <z> <maxRadius> [print below]
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
__device__ int g_traceOwner = -1;
__device__ int g_tracePrim = -1;
#define ZB_PRIM_TRACE(prim, what, a, b, c, d) if ((int)(prim) == g_tracePrim) printf("    prim %d lane %d %s: %.9g %.9g %.9g %.9g\n", (int)(prim), (int)(threadIdx.x & 31), what, (double)(a), (double)(b), (double)(c), (double)(d));
#define ZB_COOP_TRACE(owner, what, value, id) if ((owner) == g_traceOwner) printf("    trace: query %d <- %s: r2 %.9g (r %.8f) id %u\n", (int)(owner), what, (double)(value), sqrt((double)(value)), (unsigned)(id));
#include "../zombie_b200/csrc/walk.cuh"
#include "../zombie_b200/csrc/wide_bvh.h"
using namespace zb;
template <typename T> static std::vector<T> rd(const std::string& p) { FILE *f = fopen(p.c_str(), "rb"); if (!f) { printf("cannot open %s\n", p.c_str()); exit(2); } fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); std::vector<T> v(n/sizeof(T)); if (fread(v.data(), 1, n, f) != (size_t)n) exit(2); fclose(f); return v; }

__global__ void bruteKernel(BvhView v, float qx, float qy, float qz, float r2max, float below)
{
    Vec<3> x = make3(qx, qy, qz);
    float be = r2max, bf = r2max; int ie = -1, jf = -1;
    for (int p = 0; p < v.nPrims; p++) {
        float a = r2max, b = r2max;
        primStarRadius2<3>(v, p, x, a, true, 1e-3f, true);
        primStarRadius2Fast<3>(v, p, x, b, true, 1e-3f, true);
        if (a < be) { be = a; ie = p; }
        if (b < bf) { bf = b; jf = p; }
        if (sqrtf(a) < below || sqrtf(b) < below) printf("  prim %d (input %d): exact %.8f fast %.8f needsFull %d\n", p, v.primIndex[p], sqrtf(a), sqrtf(b),
                                                         (int)primStarNeedsFullTest<3>(v, p, x, r2max, true));
    }
    printf("device brute force: exact %.8f (prim %d) fast %.8f (prim %d)\n", sqrtf(be), ie, sqrtf(bf), jf);
    int status = 0; uint2 stack[kStackSize]; int sp = 0; stack[sp++] = make_uint2(0u, 0u); float r2 = r2max;
    starRadiusDfs<3, true>(v, x, r2, true, 1e-3f, stack, sp, 0x7fffffff, &status);
    printf("device one-thread traversal (fast node + primitive tests): %.8f status %d\n", sqrtf(r2), status);
    sp = 0; stack[sp++] = make_uint2(0u, 0u); r2 = r2max;
    starRadiusDfs<3, false>(v, x, r2, true, 1e-3f, stack, sp, 0x7fffffff, &status);
    printf("device one-thread traversal (trigonometric node test, reference-ordered primitive test): %.8f status %d\n", sqrtf(r2), status);
}

__global__ void primKernel(BvhView v, float qx, float qy, float qz, int prim)
{
    Vec<3> x = make3(qx, qy, qz);
    const float r2ins[16] = {0.398416f, 0.146851867f, 0.0883922055f, 0.067240119f, 0.0636010841f, 0.0551435277f, 0.0395059846f, 0.0392959192f, 0.052250728f, 0.0494382009f, 0.04466388f, 0.0535569265f, 0.0540956073f, 0.0440504551f, 0.0526154079f, 0.0503301136f};
    for (int k = 0; k < 16; k++) {
        float a = r2ins[k], b = r2ins[k];
        primStarRadius2<3>(v, prim, x, a, true, 1e-3f, true);
        primStarRadius2Fast<3>(v, prim, x, b, true, 1e-3f, true);
        float c = primStarRadius2Call<3>(v.prims, v.adjacency, v.snchSign, prim, x, r2ins[k], true, 1e-3f, true).x;
        float d = primStarRadius2Call<3>(v.prims, v.adjacency, v.snchSign, prim, x, r2ins[k], true, 1e-3f, false).x;
        printf("prim %d r2in %.9g: exact %.9g (r %.8f) fast %.9g (r %.8f) call %.9g call(noSil) %.9g needsFull %d\n", prim, r2ins[k], a, sqrtf(a), b, sqrtf(b), c, d,
               (int)primStarNeedsFullTest<3>(v, prim, x, r2ins[k], true));
    }
}

// the primitives of one pooled full round, one per lane, all for the same query: does a lane's result depend on its neighbours?
__global__ void divergentKernel(BvhView v, float qx, float qy, float qz, float r2in)
{
    const int prims[15] = {8432, 8459, 8458, 8448, 8447, 8446, 8451, 8450, 8449, 8454, 8355, 8349, 8348, 8347, 8346};
    Vec<3> x = make3(qx, qy, qz);
    int l = threadIdx.x;
    if (l < 15) {
        float r = primStarRadius2Call<3>(v.prims, v.adjacency, v.snchSign, prims[l], x, r2in, true, 1e-3f, true).x;
        printf("divergent lane %d prim %d: %.9g\n", l, prims[l], r);
    }
}

__global__ void __launch_bounds__(32) coopKernel(BvhView v, float qx, float qy, float qz, float r2max, int nQueries, int *status)
{
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[0];
    Vec<3> x = make3(qx, qy, qz);
    bool has = (int)threadIdx.x < nQueries;
    float r2 = coopStarRadius2<3>(v, ws, has, x, r2max, true, 1e-3f, status);
    if (threadIdx.x == 0) printf("device cooperative query (%d identical queries in the warp): %.8f\n", nQueries, sqrtf(r2));
}

// 32 different queries in one warp, as zb_star_radius runs them (a deviation that needs the OTHER queries of the warp)
__global__ void __launch_bounds__(128) coopWarpKernel(BvhView v, const float *x, const float *maxr, float *out, int n, int *status)
{
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    bool has = i < n;
    Vec<3> q = has ? loadVec<3>(x + 3*i) : vzero<3>();
    float r2max = has ? maxr[i]*maxr[i] : 0.0f;
    float r2 = coopStarRadius2<3>(v, ws, has, q, r2max, true, 1e-3f, status);
    if (has) out[i] = fmaxf(sqrtf(r2), 1e-3f);
}

int main(int argc, char **argv)
{
    std::string dir = argv[1]; float mu = atof(argv[2]);
    float qx = atof(argv[3]), qy = atof(argv[4]), qz = atof(argv[5]), maxr = atof(argv[6]);
    float below = argc > 7 ? atof(argv[7]) : 0.0f;
    auto rpos = rd<float>(dir + "/rpos.bin"); auto ridx = rd<int>(dir + "/ridx.bin");
    HostBvh R; const char *err = nullptr; ReflectingOptions opt; opt.enabled = true;
    std::vector<float> coeff(ridx.size()/3, mu); opt.minCoeff = coeff.data(); opt.maxCoeff = coeff.data();
    if (!buildWideBvh(3, rpos.data(), (int)rpos.size()/3, ridx.data(), (int)ridx.size()/3, opt, R, &err)) { printf("%s\n", err); return 1; }
    BvhView v{};
    auto up = [](const void *p, size_t bytes) { void *d; cudaMalloc(&d, bytes); cudaMemcpy(d, p, bytes, cudaMemcpyHostToDevice); return d; };
    v.nodes = (const uint4 *)up(R.nodes.data(), R.nodes.size()*sizeof(WideNode));
    v.prims = (const float4 *)up(R.prims.data(), R.prims.size()*4);
    v.primIndex = (const int *)up(R.primIndex.data(), R.primIndex.size()*4);
    v.cones = (const float4 *)up(R.cones.data(), R.cones.size()*sizeof(ChildCone));
    v.adjacency = (const float4 *)up(R.adjacency.data(), R.adjacency.size()*sizeof(PrimAdjacency));
    v.nPrims = R.nPrims; v.snchSign = 1;
    int *status; cudaMalloc(&status, 4); cudaMemset(status, 0, 4);
    cudaDeviceSetLimit(cudaLimitStackSize, 8192);
    bruteKernel<<<1, 1>>>(v, qx, qy, qz, maxr*maxr, below);
    cudaDeviceSynchronize();
    if (argc > 9) { primKernel<<<1, 1>>>(v, qx, qy, qz, atoi(argv[9])); cudaDeviceSynchronize(); divergentKernel<<<1, 32>>>(v, qx, qy, qz, 0.0551435277f); cudaDeviceSynchronize(); }
    cudaFuncSetAttribute(coopKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpShared));
    for (int n : {1, 32}) { coopKernel<<<1, 32, sizeof(WarpShared)>>>(v, qx, qy, qz, maxr*maxr, n, status); cudaDeviceSynchronize(); }
    {
        FILE *f = fopen((dir + "/warp_pts.bin").c_str(), "rb");
        if (f) {
            fclose(f);
            auto x = rd<float>(dir + "/warp_pts.bin"); auto mr = rd<float>(dir + "/warp_maxr.bin"); auto ref = rd<float>(dir + "/warp_ref.bin");
            int n = (int)mr.size();
            float *dx = (float *)up(x.data(), x.size()*4), *dm = (float *)up(mr.data(), mr.size()*4), *dout;
            cudaMalloc(&dout, n*4);
            cudaFuncSetAttribute(coopWarpKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4*sizeof(WarpShared)));
            int traceOwner = argc > 8 ? atoi(argv[8]) : -1;
            cudaMemcpyToSymbol(g_traceOwner, &traceOwner, sizeof(int));
            int tracePrim = argc > 9 ? atoi(argv[9]) : -1;
            cudaMemcpyToSymbol(g_tracePrim, &tracePrim, sizeof(int));
            for (int threads : {32, 128}) {
                coopWarpKernel<<<(n + threads - 1)/threads, threads, (threads/32)*sizeof(WarpShared)>>>(v, dx, dm, dout, n, status);
                std::vector<float> out(n);
                cudaMemcpy(out.data(), dout, n*4, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int i = 0; i < n; i++) if (fabsf(out[i] - ref[i]) > 1e-3f) { bad++; printf("  %d threads/CTA: query %d coop %.8f ref %.8f\n", threads, i, out[i], ref[i]); }
                printf("warp of %d different queries, %d threads per CTA: %d deviate\n", n, threads, bad);
            }
        }
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
