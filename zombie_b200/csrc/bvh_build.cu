// Host side of the on-device hierarchy builder (bvh_build.cuh): scratch carving, cub sorts / scans, the per-level
// launch loops. Everything runs on the context's stream; the host reads two integers per level (how many internal
// children and leaf primitives the level produced) to size the next level's launches and, once, the node count to
// allocate the final arrays exactly.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "bvh_build.cuh"
#include "scene_storage.h"

namespace zb {

using namespace lbvh;

namespace {

struct Carver {
    char *base = nullptr;
    size_t used = 0;
    template <typename T> T *take(size_t n) {
        used = (used + 255) & ~size_t(255);
        T *p = base ? reinterpret_cast<T *>(base + used) : nullptr;
        used += n*sizeof(T);
        return p;
    }
};

struct Scratch {
    float *centroids;
    uint32_t *bounds; // 6 ordered floats | flags | leaf-children counter
    uint64_t *keys, *keysAlt;
    int *order, *orderAlt, *leafPos;
    TopoNode *topo;
    int *nInternal, *nLeafPrims, *internalIncl, *leafIncl;
    Box3 *nodeBox;
    uint64_t *heKeys, *heKeysAlt;
    uint32_t *heVals, *heValsAlt;
    int *opposite, *startsAt, *endsAt;
    PrimAdjacency *adjInput;
    void *cubTemp;
    size_t cubBytes;
};

int bitsFor(uint64_t v) { int b = 0; while (b < 64 && (v >> b) != 0) b++; return std::max(1, b); }

__global__ void exclusiveFromInclusiveKernel(int count, const int *nInternal, const int *nLeafPrims, int *internalIncl, int *leafIncl,
                                             const TopoNode *topo, int first, uint32_t *leafChildren)
{
    int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= count) return;
    internalIncl[k] -= nInternal[k];
    leafIncl[k] -= nLeafPrims[k];
    atomicAdd(leafChildren, (uint32_t)(topo[first + k].nChildren - nInternal[k]));
}

} // namespace

int buildBvhOnDevice(zb_context *ctx, DeviceBvhStorage& s, int dim, int nPrim, bool doubleSided, float silhouettePrecision, int snchSign)
{
    cudaStream_t st = ctx->stream;
    const int n = nPrim;
    const int cap = n/4 + 2;
    const bool refl = s.reflecting;
    const int nHalf = refl && dim == 3 ? 3*n : 0;
    Mesh mesh{dim, n, s.nVert, s.meshPos.as<float>(), s.meshIdx.as<int32_t>()};

    // ZB_BUILD_TRACE=1: wall clock of every stage (each closed by a stream synchronisation) on stderr
    const bool trace = getenv("ZB_BUILD_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[bvh build] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    cudaEvent_t ev0, ev1;
    ZB_CUDA(cudaEventCreate(&ev0));
    ZB_CUDA(cudaEventCreate(&ev1));
    ZB_CUDA(cudaEventRecord(ev0, st));

    // ---- scratch ----
    size_t cubSort = 0, cubSortHe = 0, cubScan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cubSort, (uint64_t *)nullptr, (uint64_t *)nullptr, (int *)nullptr, (int *)nullptr, n, 0, 64, st);
    if (nHalf) cub::DeviceRadixSort::SortPairs(nullptr, cubSortHe, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, nHalf, 0, 64, st);
    cub::DeviceScan::InclusiveSum(nullptr, cubScan, (int *)nullptr, (int *)nullptr, cap, st);
    Scratch k{};
    DeviceBuffer& scratch = ctx->bvhBuild;
    for (int pass = 0; pass < 2; pass++) {
        Carver c;
        c.base = pass ? scratch.as<char>() : nullptr;
        k.centroids = c.take<float>((size_t)3*n);
        k.bounds = c.take<uint32_t>(8);
        k.keys = c.take<uint64_t>(n); k.keysAlt = c.take<uint64_t>(n);
        k.order = c.take<int>(n); k.orderAlt = c.take<int>(n); k.leafPos = c.take<int>(n);
        k.topo = c.take<TopoNode>(cap);
        k.nInternal = c.take<int>(cap); k.nLeafPrims = c.take<int>(cap);
        k.internalIncl = c.take<int>(cap); k.leafIncl = c.take<int>(cap);
        k.nodeBox = c.take<Box3>(cap);
        if (nHalf) {
            k.heKeys = c.take<uint64_t>(nHalf); k.heKeysAlt = c.take<uint64_t>(nHalf);
            k.heVals = c.take<uint32_t>(nHalf); k.heValsAlt = c.take<uint32_t>(nHalf);
            k.opposite = c.take<int>(nHalf);
        }
        if (refl && dim == 2) { k.startsAt = c.take<int>(s.nVert); k.endsAt = c.take<int>(s.nVert); }
        if (refl) k.adjInput = c.take<PrimAdjacency>(n);
        k.cubBytes = std::max(cubSort, std::max(cubSortHe, cubScan));
        k.cubTemp = c.take<char>(k.cubBytes);
        if (!pass) ZB_CUDA(scratch.reserve(c.used + 256));
    }

    mark("scratch");
    // ---- 1. Morton order ----
    const uint32_t boundsInit[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
    ZB_CUDA(cudaMemcpyAsync(k.bounds, boundsInit, sizeof(boundsInit), cudaMemcpyHostToDevice, st));
    centroidBoundsKernel<<<(n + 255)/256, 256, 0, st>>>(mesh, k.centroids, k.bounds, reinterpret_cast<int *>(k.bounds + 6));
    ZB_CUDA(cudaGetLastError());
    uint32_t hostBounds[8];
    ZB_CUDA(cudaMemcpyAsync(hostBounds, k.bounds, sizeof(hostBounds), cudaMemcpyDeviceToHost, st));
    ZB_CUDA(cudaStreamSynchronize(st));
    if (hostBounds[6] & 1u) return fail(ZB_ERR_INVALID_ARGUMENT, "primitive index out of range");
    if (hostBounds[6] & 2u) return fail(ZB_ERR_INVALID_ARGUMENT, "non-finite vertex position");
    mortonKernel<<<(n + 255)/256, 256, 0, st>>>(dim, n, k.centroids, k.bounds, k.keys, k.order);
    ZB_CUDA(cudaGetLastError());
    ZB_CUDA(cub::DeviceRadixSort::SortPairs(k.cubTemp, k.cubBytes, k.keys, k.keysAlt, k.order, k.orderAlt, n, 0, dim == 3 ? 63 : 62, st));
    mark("bounds, keys, sort");
    const uint64_t *keys = k.keysAlt;
    const int *order = k.orderAlt;

    // ---- 2. topology, level by level ----
    TopoNode root{};
    root.start = 0; root.end = n;
    ZB_CUDA(cudaMemcpyAsync(k.topo, &root, sizeof(TopoNode), cudaMemcpyHostToDevice, st));
    std::vector<int> levelStart(1, 0);
    int first = 0, count = 1, primCursor = 0;
    while (count > 0) {
        if (first + count > cap) return fail(ZB_ERR_UNSUPPORTED, "device BVH build: node capacity exceeded");
        splitLevelKernel<<<(count + 127)/128, 128, 0, st>>>(keys, k.topo, first, count, k.nInternal, k.nLeafPrims);
        ZB_CUDA(cudaGetLastError());
        ZB_CUDA(cub::DeviceScan::InclusiveSum(k.cubTemp, k.cubBytes, k.nInternal, k.internalIncl, count, st));
        ZB_CUDA(cub::DeviceScan::InclusiveSum(k.cubTemp, k.cubBytes, k.nLeafPrims, k.leafIncl, count, st));
        int totals[2];
        ZB_CUDA(cudaMemcpyAsync(&totals[0], k.internalIncl + count - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        ZB_CUDA(cudaMemcpyAsync(&totals[1], k.leafIncl + count - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        exclusiveFromInclusiveKernel<<<(count + 127)/128, 128, 0, st>>>(count, k.nInternal, k.nLeafPrims, k.internalIncl, k.leafIncl,
                                                                       k.topo, first, k.bounds + 7);
        ZB_CUDA(cudaGetLastError());
        placeLevelKernel<<<(count + 127)/128, 128, 0, st>>>(k.topo, first, count, k.internalIncl, k.leafIncl, first + count, primCursor, k.leafPos);
        ZB_CUDA(cudaGetLastError());
        ZB_CUDA(cudaStreamSynchronize(st));
        first += count;
        levelStart.push_back(first);
        count = totals[0];
        primCursor += totals[1];
    }
    mark("topology");
    const int nNodes = first;
    if (primCursor != n) return fail(ZB_ERR_UNSUPPORTED, "device BVH build: primitives lost in the topology pass");

    // ---- final arrays ----
    const int primFloats = dim == 3 ? 12 : 4;
    ZB_CUDA(s.nodes.reserveExact((size_t)nNodes*sizeof(WideNode)));
    ZB_CUDA(s.prims.reserveExact((size_t)primFloats*n*sizeof(float)));
    ZB_CUDA(s.primIndex.reserveExact((size_t)n*sizeof(int32_t)));
    if (refl) {
        ZB_CUDA(s.cones.reserveExact((size_t)nNodes*kWidth*sizeof(ChildCone)));
        ZB_CUDA(s.adjacency.reserveExact((size_t)n*sizeof(PrimAdjacency)));
    }

    mark("final arrays");
    // ---- 3. adjacency ----
    if (refl) {
        if (dim == 3) {
            halfEdgeKernel<<<(nHalf + 255)/256, 256, 0, st>>>(mesh, k.heKeys, k.heVals);
            ZB_CUDA(cudaGetLastError());
            // vertex indices are below nVert: only the bits in use are sorted (the high word first ends at 32 + bits(nVert))
            ZB_CUDA(cub::DeviceRadixSort::SortPairs(k.cubTemp, k.cubBytes, k.heKeys, k.heKeysAlt, k.heVals, k.heValsAlt, nHalf, 0,
                                                    32 + bitsFor((uint64_t)s.nVert), st));
            oppositeFaceKernel<<<(nHalf + 255)/256, 256, 0, st>>>(k.heKeysAlt, k.heValsAlt, nHalf, k.opposite);
            ZB_CUDA(cudaGetLastError());
        } else {
            ZB_CUDA(cudaMemsetAsync(k.startsAt, 0xff, (size_t)s.nVert*sizeof(int), st));
            ZB_CUDA(cudaMemsetAsync(k.endsAt, 0xff, (size_t)s.nVert*sizeof(int), st));
            vertexTablesKernel<<<(n + 255)/256, 256, 0, st>>>(mesh, k.startsAt, k.endsAt);
            ZB_CUDA(cudaGetLastError());
        }
        adjacencyKernel<<<(n + 127)/128, 128, 0, st>>>(mesh, k.opposite, k.startsAt, k.endsAt,
                                                      s.robin ? s.coefMin.as<float>() : nullptr, s.robin ? s.coefMax.as<float>() : nullptr,
                                                      doubleSided ? 1 : 0, silhouettePrecision, k.adjInput);
        ZB_CUDA(cudaGetLastError());
    }

    mark("adjacency");
    // ---- 4. leaf-order records ----
    scatterPrimsKernel<<<(n + 255)/256, 256, 0, st>>>(mesh, order, k.leafPos, refl ? k.adjInput : nullptr, s.prims.as<float>(),
                                                     s.primIndex.as<int32_t>(), refl ? s.adjacency.as<PrimAdjacency>() : nullptr);
    ZB_CUDA(cudaGetLastError());

    mark("scatter");
    // ---- 5. boxes, deepest level first ----
    for (int l = (int)levelStart.size() - 2; l >= 0; l--) {
        const int lf = levelStart[l], lc = levelStart[l + 1] - lf;
        writeLevelKernel<<<(lc + 127)/128, 128, 0, st>>>(dim, k.topo, lf, lc, s.prims.as<float>(), k.nodeBox, s.nodes.as<WideNode>());
        ZB_CUDA(cudaGetLastError());
    }

    mark("boxes");
    // ---- 6. cones ----
    if (refl) {
        for (size_t l = 0; l + 1 < levelStart.size(); l++) {
            const int lf = levelStart[l], lc = levelStart[l + 1] - lf;
            // top levels: few slots of many primitives, one CTA per slot; below: one warp per slot
            if ((long long)lc*kWidth*512 <= (long long)n) {
                coneLevelKernel<16><<<(unsigned)(lc*kWidth), 512, 0, st>>>(mesh, k.topo, lf, lc, s.nodes.as<WideNode>(), order, k.adjInput,
                                                                          s.cones.as<ChildCone>());
            } else {
                const long long threads = (long long)lc*kWidth*32;
                coneLevelKernel<1><<<(unsigned)((threads + 127)/128), 128, 0, st>>>(mesh, k.topo, lf, lc, s.nodes.as<WideNode>(), order,
                                                                                  k.adjInput, s.cones.as<ChildCone>());
            }
            ZB_CUDA(cudaGetLastError());
        }
    }

    mark("cones");
    Box3 rootBox;
    uint32_t leafChildren = 0;
    ZB_CUDA(cudaMemcpyAsync(&rootBox, k.nodeBox, sizeof(Box3), cudaMemcpyDeviceToHost, st));
    ZB_CUDA(cudaMemcpyAsync(&leafChildren, k.bounds + 7, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ZB_CUDA(cudaEventRecord(ev1, st));
    ZB_CUDA(cudaStreamSynchronize(st));
    ZB_CUDA(cudaEventElapsedTime(&s.buildMs, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);

    s.host = HostBvh();
    s.host.dim = dim;
    s.host.nPrims = n;
    s.host.nLeaves = (int)leafChildren;
    s.host.maxDepth = (int)levelStart.size() - 1;
    for (int a = 0; a < 3; a++) { s.host.boxMin[a] = rootBox.lo[a]; s.host.boxMax[a] = rootBox.hi[a]; }
    s.nNodes = nNodes;
    s.levelStart = levelStart;
    s.view.nodes = s.nodes.as<uint4>();
    s.view.prims = s.prims.as<float4>();
    s.view.primIndex = s.primIndex.as<int>();
    s.view.cones = refl ? s.cones.as<float4>() : nullptr;
    s.view.adjacency = refl ? s.adjacency.as<float4>() : nullptr;
    s.view.nPrims = n;
    s.view.snchSign = snchSign;
    s.present = true;
    s.builder = ZB_BVH_BUILDER_DEVICE;
    return ZB_OK;
}

} // namespace zb
