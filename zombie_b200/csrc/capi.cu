// C ABI of libzombie_b200.so (include/zombie_b200.h): contexts, scenes (boundary meshes -> wide BVHs on
// the device), batched geometric queries and the walk-on-spheres / walk-on-stars solve.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/zombie_b200.h"
#include "capi_common.h"
#include "walk.cuh"
#include "wave.cuh"
#include "rws.cuh"
#include "wide_bvh.h"
#include "scene_storage.h"

using namespace zb;

namespace {
// zb_solve knows from the host arrays whether any point asks for gradients and tells zb_solve_dev
// so that solution-only runs do not allocate gradient records; zb_solve_dev alone assumes they may
thread_local bool g_gradientHint = true;
// zb_solve_channels(_dev) hands the extra value channels (descriptors and statistics already on the device) to the
// zb_solve_dev it calls for channel 0
struct ExtraChannels { int n = 0; const zb_pde_desc *pde = nullptr; const zb_statistics *stats = nullptr; };
thread_local ExtraChannels g_extra;
} // namespace

struct zb_scene {
    zb_context *ctx = nullptr;
    int dim = 3;
    int reflectingKind = ZB_REFLECTING_NONE;
    DeviceBvhStorage boundary[2];
    float domainMin[3] = {0, 0, 0}, domainMax[3] = {0, 0, 0};
    int watertight = 1;
    int solveDoubleSided = 0;
    float silhouettePrecision = 1e-3f;
};

namespace {

int uploadBvh(zb_context *ctx, DeviceBvhStorage& s, int snchSign)
{
    const HostBvh& h = s.host;
    ZB_CUDA(s.nodes.reserveExact(h.nodes.size()*sizeof(WideNode)));
    ZB_CUDA(s.prims.reserveExact(h.prims.size()*sizeof(float)));
    ZB_CUDA(s.primIndex.reserveExact(h.primIndex.size()*sizeof(int32_t)));
    ZB_CUDA(cudaMemcpyAsync(s.nodes.ptr, h.nodes.data(), h.nodes.size()*sizeof(WideNode), cudaMemcpyHostToDevice, ctx->stream));
    ZB_CUDA(cudaMemcpyAsync(s.prims.ptr, h.prims.data(), h.prims.size()*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ZB_CUDA(cudaMemcpyAsync(s.primIndex.ptr, h.primIndex.data(), h.primIndex.size()*sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    s.view.nodes = s.nodes.as<uint4>();
    s.view.prims = s.prims.as<float4>();
    s.view.primIndex = s.primIndex.as<int>();
    s.view.cones = nullptr;
    s.view.adjacency = nullptr;
    if (!h.cones.empty()) {
        ZB_CUDA(s.cones.reserveExact(h.cones.size()*sizeof(ChildCone)));
        ZB_CUDA(s.adjacency.reserveExact(h.adjacency.size()*sizeof(PrimAdjacency)));
        ZB_CUDA(cudaMemcpyAsync(s.cones.ptr, h.cones.data(), h.cones.size()*sizeof(ChildCone), cudaMemcpyHostToDevice, ctx->stream));
        ZB_CUDA(cudaMemcpyAsync(s.adjacency.ptr, h.adjacency.data(), h.adjacency.size()*sizeof(PrimAdjacency), cudaMemcpyHostToDevice, ctx->stream));
        s.view.cones = s.cones.as<float4>();
        s.view.adjacency = s.adjacency.as<float4>();
    }
    s.view.nPrims = h.nPrims;
    s.view.snchSign = snchSign;
    s.present = true;
    s.nNodes = (int)h.nodes.size();
    s.builder = ZB_BVH_BUILDER_HOST_SAH;
    s.levelStart.clear();
    ZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZB_OK;
}

int checkStatus(zb_context *ctx, const char *what)
{
    ZB_CUDA(cudaMemcpyAsync(ctx->hostStatus, ctx->status.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ZB_CUDA(cudaStreamSynchronize(ctx->stream));
    int st = *ctx->hostStatus;
    if (st != 0) {
        ZB_CUDA(cudaMemsetAsync(ctx->status.ptr, 0, sizeof(int), ctx->stream));
        std::string msg = std::string(what) + ":";
        if (st & kStatusStackOverflow) msg += " BVH traversal stack overflow";
        if (st & kStatusSplitOverflow) msg += " walk split queue overflow (more than 6 pending splits)";
        if (st & kStatusCoopOverflow) msg += " cooperative traversal stack overflow";
        return fail(ZB_ERR_UNSUPPORTED, msg);
    }
    return ZB_OK;
}

// ---- batched query kernels ---------------------------------------------------------------------
constexpr size_t kCoopSmem = 4*sizeof(WarpShared); // 128-thread CTAs

template <int DIM>
__global__ void __launch_bounds__(128) closestPointKernel(BvhView bvh, int n, const float *x, const float *maxRadius,
                                                          int *prim, float *dist, float *pt, float *normal, int *status)
{
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    bool has = i < n;
    Vec<DIM> q = has ? loadVec<DIM>(x + (size_t)DIM*i) : vzero<DIM>();
    float r2 = kMaxFloat;
    if (has && maxRadius && maxRadius[i] < kMaxFloat) r2 = maxRadius[i]*maxRadius[i];
    ClosestHit<DIM> hit;
    coopClosestPoint<DIM>(bvh, ws, has, q, r2, hit, status);
    if (!has) return;
    closestTieBreak<DIM>(bvh, q, hit, status); // exact ties report the largest input index, as FCPW's Baseline does (queries.cuh)
    bool found = hit.prim >= 0;
    prim[i] = found ? bvh.primIndex[hit.prim] : -1;
    dist[i] = found ? hit.d : kMaxFloat;
    if (pt) storeVec<DIM>(pt + (size_t)DIM*i, found ? hit.p : vzero<DIM>());
    if (normal) {
        Vec<DIM> nrm = vzero<DIM>();
        if (found) { Prim<DIM> p; p.load(bvh, (uint32_t)hit.prim); nrm = unitNormal<DIM>(p); }
        storeVec<DIM>(normal + (size_t)DIM*i, nrm);
    }
}

template <int DIM>
__global__ void intersectRayKernel(BvhView bvh, int n, const float *origin, const float *dir, const float *tMax,
                                   int *prim, float *dist, float *pt, float *normal, int *status)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    Vec<DIM> o = loadVec<DIM>(origin + (size_t)DIM*i), d = loadVec<DIM>(dir + (size_t)DIM*i);
    RayHit<DIM> hit;
    bool found = intersectRay<DIM, false>(bvh, o, d, tMax ? tMax[i] : kMaxFloat, hit, status);
    prim[i] = found ? bvh.primIndex[hit.prim] : -1;
    dist[i] = found ? hit.t : kMaxFloat;
    Vec<DIM> p = vzero<DIM>(), nrm = vzero<DIM>();
    if (found) {
        Prim<DIM> tri;
        tri.load(bvh, (uint32_t)hit.prim);
        p = tri.pointAt(hit.bv, hit.bw);
        nrm = unitNormal<DIM>(tri);
    }
    if (pt) storeVec<DIM>(pt + (size_t)DIM*i, p);
    if (normal) storeVec<DIM>(normal + (size_t)DIM*i, nrm);
}

// Aggregate::intersect(ray, interactions, false, true) reduced to its return value (number of distinct hits)
template <int DIM>
__global__ void countRayHitsKernel(BvhView bvh, int n, const float *origin, const float *dir, const float *tMax, int *count, int *status)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    Vec<DIM> o = loadVec<DIM>(origin + (size_t)DIM*i), d = loadVec<DIM>(dir + (size_t)DIM*i);
    count[i] = countRayHits<DIM>(bvh, o, d, tMax ? tMax[i] : kMaxFloat, status);
}

// GeometricQueries::insideDomain (core/geometric_queries.h:357-371): one all-hits ray along each coordinate axis over both
// boundaries; inside when every count is odd
template <int DIM>
__global__ void insideDomainKernel(BvhView absorbing, BvhView reflecting, int hasAbsorbing, int hasReflecting, int n, const float *x,
                                   uint8_t *inside, int *status)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    Vec<DIM> o = loadVec<DIM>(x + (size_t)DIM*i);
    bool in = true;
    for (int axis = 0; axis < DIM; axis++) {
        Vec<DIM> d = vzero<DIM>();
        d[axis] = 1.0f;
        int hits = 0;
        if (hasAbsorbing) hits += countRayHits<DIM>(absorbing, o, d, kMaxFloat, status);
        if (hasReflecting) hits += countRayHits<DIM>(reflecting, o, d, kMaxFloat, status);
        in = in && (hits & 1);
    }
    inside[i] = in ? 1 : 0;
}

template <int DIM>
__global__ void occludedKernel(BvhView bvh, int n, const float *a, const float *b, uint8_t *out, int *status)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    // Aggregate::hasLineOfSight (fcpw/core/primitive.h:202-212)
    Vec<DIM> pa = loadVec<DIM>(a + (size_t)DIM*i), pb = loadVec<DIM>(b + (size_t)DIM*i);
    Vec<DIM> d = pb - pa;
    float len = norm(d);
    RayHit<DIM> hit;
    out[i] = intersectRay<DIM, true>(bvh, pa, d/len, len, hit, status) ? 1 : 0;
}

template <int DIM>
__global__ void sampleSphereKernel(BvhView bvh, int n, const float *x, const float *radius, const float *rand,
                                   int *prim, float *pt, float *normal, float *pdf)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    Vec<DIM> q = loadVec<DIM>(x + (size_t)DIM*i), u = loadVec<DIM>(rand + (size_t)DIM*i);
    SphereSample<DIM> s;
    bool found = sampleSphere<DIM>(bvh, q, radius[i]*radius[i], u, s);
    prim[i] = found ? bvh.primIndex[s.prim] : -1;
    pdf[i] = found ? s.pdf : 0.0f;
    if (pt) storeVec<DIM>(pt + (size_t)DIM*i, s.p);
    if (normal) storeVec<DIM>(normal + (size_t)DIM*i, s.n);
}

template <int DIM>
__global__ void __launch_bounds__(128) starRadiusKernel(SolveArgs a, int n, const float *x, const float *minR, const float *maxR,
                                                        float precision, const uint8_t *flip, float *out)
{
    extern __shared__ __align__(16) unsigned char zbShared[];
    WarpShared& ws = reinterpret_cast<WarpShared *>(zbShared)[threadIdx.x >> 5];
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    bool has = i < n;
    Vec<DIM> q = has ? loadVec<DIM>(x + (size_t)DIM*i) : vzero<DIM>();
    float mn = has ? minR[i] : 0.0f, mx = has ? maxR[i] : 0.0f;
    bool fl = has && flip ? flip[i] != 0 : false;
    float result = mx, r2max = 0.0f;
    bool query = has && starQueryNeeded<DIM>(a, mn, mx, result, r2max);
    float r2 = coopStarRadius2<DIM>(a.reflecting, ws, query, q, r2max, !fl, precision, a.status);
    if (query) result = starQueryResult(a, r2, r2max, mn, mx);
    if (has) out[i] = result;
}

__global__ void pcg32Kernel(uint64_t initstate, uint64_t initseq, int n, uint32_t *out)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        Pcg32 rng;
        rng.seed(initstate, initseq);
        for (int i = 0; i < n; i++) out[i] = rng.nextUInt();
    }
}

__global__ void pcg32DemoKernel(uint64_t initstate, uint64_t initseq, int rounds, uint32_t *u32, uint8_t *coins, uint8_t *rolls, uint8_t *cards)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Pcg32 rng;
    rng.seed(initstate, initseq);
    for (int r = 0; r < rounds; r++) {
        for (int i = 0; i < 6; i++) u32[6*r + i] = rng.nextUInt();
        for (int i = 0; i < 65; i++) coins[65*r + i] = (uint8_t)rng.nextUInt(2u);
        for (int i = 0; i < 33; i++) rolls[33*r + i] = (uint8_t)(rng.nextUInt(6u) + 1u);
        uint8_t *c = cards + 52*r;
        for (int i = 0; i < 52; i++) c[i] = (uint8_t)i;
        // pcg32::shuffle (deps/pcg32/pcg32.h:169-172)
        for (int it = 51; it > 0; it--) {
            uint32_t j = rng.nextUInt((uint32_t)(it + 1));
            uint8_t t = c[it]; c[it] = c[j]; c[j] = t;
        }
    }
}

__global__ void pcg32FloatKernel(uint64_t initstate, uint64_t initseq, int n, float *out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Pcg32 rng;
    rng.seed(initstate, initseq);
    for (int i = 0; i < n; i++) out[i] = rng.nextFloat();
}

// carves typed device sub-buffers out of one allocation
struct Carver {
    char *base;
    size_t offset = 0;
    explicit Carver(void *p): base(reinterpret_cast<char *>(p)) {}
    template <typename T> T *take(size_t count) {
        offset = (offset + 255) & ~size_t(255);
        T *p = base ? reinterpret_cast<T *>(base + offset) : nullptr;
        offset += count*sizeof(T);
        return p;
    }
};

// staging of host arrays for the batched query entry points
struct IoPlan {
    std::vector<std::pair<const void *, size_t>> inputs;   // host src, bytes
    std::vector<std::pair<void *, size_t>> outputs;        // host dst, bytes
    std::vector<size_t> inOffsets, outOffsets;
    size_t total = 0;
    size_t addIn(const void *src, size_t bytes) {
        total = (total + 255) & ~size_t(255);
        inputs.push_back({src, bytes}); inOffsets.push_back(total);
        size_t o = total; total += bytes; return o;
    }
    size_t addOut(void *dst, size_t bytes) {
        total = (total + 255) & ~size_t(255);
        outputs.push_back({dst, bytes}); outOffsets.push_back(total);
        size_t o = total; total += bytes; return o;
    }
};

int stageIn(zb_context *ctx, IoPlan& plan)
{
    ZB_CUDA(ctx->io.reserve(plan.total + 256));
    for (size_t k = 0; k < plan.inputs.size(); k++) {
        if (plan.inputs[k].first && plan.inputs[k].second)
            ZB_CUDA(cudaMemcpyAsync(ctx->io.as<char>() + plan.inOffsets[k], plan.inputs[k].first, plan.inputs[k].second,
                                    cudaMemcpyHostToDevice, ctx->stream));
    }
    return ZB_OK;
}

int stageOut(zb_context *ctx, IoPlan& plan)
{
    for (size_t k = 0; k < plan.outputs.size(); k++) {
        if (plan.outputs[k].first && plan.outputs[k].second)
            ZB_CUDA(cudaMemcpyAsync(plan.outputs[k].first, ctx->io.as<char>() + plan.outOffsets[k], plan.outputs[k].second,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    }
    ZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZB_OK;
}

template <typename T> T *at(zb_context *ctx, size_t offset, const void *hostPtr)
{
    return hostPtr ? reinterpret_cast<T *>(ctx->io.as<char>() + offset) : nullptr;
}

inline int blocksFor(int n, int threads) { return (n + threads - 1)/threads; }

void fillSceneArgs(const zb_scene *scene, SolveArgs& a)
{
    std::memset(&a, 0, sizeof(a));
    a.absorbing = scene->boundary[ZB_ABSORBING].view;
    a.reflecting = scene->boundary[ZB_REFLECTING].view;
    a.hasAbsorbing = scene->boundary[ZB_ABSORBING].present ? 1 : 0;
    a.hasReflecting = scene->boundary[ZB_REFLECTING].present ? 1 : 0;
    a.reflectingKind = scene->reflectingKind;
    for (int k = 0; k < 3; k++) { a.domainMin[k] = scene->domainMin[k]; a.domainMax[k] = scene->domainMax[k]; }
    a.status = scene->ctx->status.as<int>();
}


// ---- wavefront scheduler (wave.cuh): the host side of one solve's walks ----------------------------------------
int envInt(const char *name, int fallback)
{
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : fallback;
}

constexpr int kLiveRing = 8;       // live-walk counters in flight per group
constexpr int kRoundsPerGroup = 4; // rounds enqueued between two read-backs of a group's live count
constexpr int kMaxWaveGroups = 8;

// One group of slots = an independent sequence of rounds on its own stream. A round's kernels end with a tail (a few warps
// still inside heavy queries while most SMs idle, ncu: achieved occupancy 14 % of a theoretical 25 %); with several groups
// in flight the block scheduler fills that tail with another group's kernels. Group g owns the tasks g, g + G, g + 2G, ...,
// so what a group computes does not depend on how the groups interleave.
struct WaveGroup {
    WaveSlots w;
    uint32_t *orderA = nullptr, *orderB = nullptr, *keysOut = nullptr;
    uint32_t *orderStep = nullptr, *stepKeys = nullptr, *stepKeysOut = nullptr; // the step kernel's own order (waveStepKeyKernel)
    unsigned int *live = nullptr;
    void *sortTemp = nullptr;
    cudaStream_t stream = nullptr;
    int nSlots = 0, nBound = 0, round = 0, batches = 0;
    bool done = false;
};

template <int DIM, bool WOST>
int runWavefront(zb_scene *scene, SolveArgs& a, long long totalTasks)
{
    zb_context *ctx = scene->ctx;
    cudaStream_t stream = ctx->stream;
    // Slots: enough walks in flight to fill the machine several times over, never more than there are tasks.
    // ZB_WAVE_SLOTS / ZB_WAVE_GROUPS / ZB_WAVE_SORT / ZB_WAVE_BLOCK override (measurements).
    const int maxSlots = std::max(128, envInt("ZB_WAVE_SLOTS", 1048576));
    int nGroups = std::min(kMaxWaveGroups, std::max(1, envInt("ZB_WAVE_GROUPS", 1)));
    while (nGroups > 1 && totalTasks < (long long)nGroups*4096) nGroups--; // small solves: one group
    const int sortMode = envInt("ZB_WAVE_SORT", 0);
    const int qThreads = std::min(128, std::max(32, envInt("ZB_WAVE_BLOCK", 128)/32*32)); // CTA size of the two query kernels
    const size_t qSmem = (size_t)(qThreads/32)*sizeof(WarpShared);
    const bool needQueue = a.ws.splittingThreshold < kMaxFloat;
    const int stripeWarps = std::max(1, envInt("ZB_WAVE_STRIPE", 4096)/nGroups); // warps the tail of a solve spreads over (ZB_WAVE_STRIPE=1: compact)
    // live walks at or below which a group hands its walks to the barrier-free tail kernel (wave.cuh: waveTailKernel); 0: never
    const int tailWalks = std::max(0, envInt("ZB_WAVE_TAIL", 8192))/nGroups;
    const bool stepSort = envInt("ZB_WAVE_STEP_SORT", 1) != 0; // the step kernel's own order, by star radius class (A/B: 0)
    int tailPerSm = 0;
    if (tailWalks > 0) {
        ZB_CUDA(cudaFuncSetAttribute(waveTailKernel<DIM, WOST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
        ZB_CUDA(cudaFuncSetAttribute(waveTailKernel<DIM, WOST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
        if (a.extra) ZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tailPerSm, waveTailKernel<DIM, WOST, true>, 128, kCoopSmem));
        else ZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tailPerSm, waveTailKernel<DIM, WOST, false>, 128, kCoopSmem));
        tailPerSm = std::max(1, tailPerSm);
    }
    // the tail kernel steps ONE walk per lane of its resident grid: never hand it more walks than that
    const int tailCapacity = tailWalks > 0 ? std::max(1, ctx->smCount*tailPerSm/nGroups)*128 : 0;

    const bool star = WOST && a.hasReflecting;
    // the star-radius query as two kernels (private phase, pooled phase: wave.cuh) or as one (ZB_WAVE_STAR_SPLIT=0)
    const bool starSplit = star && envInt("ZB_WAVE_STAR_SPLIT", 1) != 0;
    WaveGroup groups[kMaxWaveGroups];
    size_t sortBytes = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) ZB_CUDA(ctx->wave.reserve(sortBytes + 256)); // sortBytes: total of pass 0
        Carver c(pass == 0 ? nullptr : ctx->wave.ptr);
        for (int g = 0; g < nGroups; g++) {
            WaveGroup& G = groups[g];
            long long tasksOfGroup = (totalTasks - g + nGroups - 1)/nGroups;
            const int nSlots = (int)std::min<long long>(maxSlots/nGroups/128*128 > 0 ? maxSlots/nGroups/128*128 : 128, (tasksOfGroup + 127)/128*128);
            G.nSlots = G.nBound = std::max(nSlots, 128);
            WaveSlots& w = G.w;
            std::memset(&w, 0, sizeof(w));
            const size_t n = (size_t)G.nSlots;
            w.ptDist = c.take<float4>(n); w.projStar = c.take<float4>(n); w.normalPrev = c.take<float4>(n);
            w.dirThr = c.take<float4>(n); w.acc = c.take<float4>(n); w.rng = c.take<ulonglong2>(n);
            w.meta = c.take<int4>(n); w.meta2 = c.take<int4>(n);
            w.queue = needQueue ? (void *)c.take<WalkState<DIM>>(n*kSplitStack) : nullptr;
            if (starSplit) {
                // hand-over of the star-radius query between its two kernels: per virtual thread (slots + one CTA of padding)
                const size_t nt = n + 128;
                const size_t rows = std::min<size_t>(kStackSize, 1 + 7*(size_t)kPhase1Budget); // a pop adds at most 7 entries net
                w.starStride = nt;
                w.starStack = c.take<uint2>(rows*nt);
                w.starPending = c.take<uint2>(nt);
                w.starWarpMask = c.take<uint32_t>(nt/32);
            }
            w.keys = c.take<uint32_t>(n);
            G.keysOut = c.take<uint32_t>(n);
            G.orderA = c.take<uint32_t>(n);
            G.orderB = c.take<uint32_t>(n);
            G.orderStep = c.take<uint32_t>(n);
            G.stepKeys = c.take<uint32_t>(n);
            G.stepKeysOut = c.take<uint32_t>(n);
            G.live = c.take<unsigned int>(kLiveRing*4); // per round in flight: live walks, the two query kernels' warp tickets, padding
            w.taskCounter = c.take<unsigned long long>(1);
            size_t bytes = 0;
            ZB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, w.keys, G.keysOut, G.orderA, G.orderB, G.nSlots, 0, 32, stream));
            G.sortTemp = c.take<char>(bytes + 256);
            w.nSlots = G.nSlots;
            w.sortMode = sortMode;
            w.debugEscapes = envInt("ZB_WAVE_DEBUG", 0);
            w.taskFirst = g;
            w.taskStride = nGroups;
        }
        sortBytes = c.offset;
    }
    // Morton frame: the union of the boundary boxes (walks live inside the domain; escapees clamp)
    float lo[3] = {kMaxFloat, kMaxFloat, kMaxFloat}, hi[3] = {-kMaxFloat, -kMaxFloat, -kMaxFloat};
    for (int b = 0; b < 2; b++) {
        if (!scene->boundary[b].present) continue;
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], scene->boundary[b].host.boxMin[k]);
            hi[k] = std::max(hi[k], scene->boundary[b].host.boxMax[k]);
        }
    }
    for (int g = 0; g < nGroups; g++) {
        for (int k = 0; k < 3; k++) {
            float l = hi[k] > lo[k] ? lo[k] : 0.0f, h = hi[k] > lo[k] ? hi[k] : 1.0f;
            groups[g].w.sortMin[k] = l;
            groups[g].w.sortScale[k] = (DIM == 3 ? 1024.0f : 32768.0f)/(h - l);
        }
    }

    if (!ctx->hostLive) ZB_CUDA(cudaMallocHost(&ctx->hostLive, 2*kMaxWaveGroups*sizeof(unsigned int)));
    for (int k = 0; k < 2*kMaxWaveGroups; k++) if (!ctx->evRound[k]) ZB_CUDA(cudaEventCreateWithFlags(&ctx->evRound[k], cudaEventDisableTiming));
    if (!ctx->evFork) ZB_CUDA(cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
    for (int g = 1; g < nGroups; g++) if (!ctx->waveStreams[g]) ZB_CUDA(cudaStreamCreateWithFlags(&ctx->waveStreams[g], cudaStreamNonBlocking));

    ZB_CUDA(cudaFuncSetAttribute(waveClosestKernel<DIM, WOST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
    if (WOST) ZB_CUDA(cudaFuncSetAttribute(waveStarKernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
    if (WOST) ZB_CUDA(cudaFuncSetAttribute(waveStarPooledKernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));

    int closestResident = 1, starResident = 1;
    ZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&closestResident, waveClosestKernel<DIM, WOST>, qThreads, qSmem));
    if (WOST && !starSplit) ZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&starResident, waveStarKernel<DIM>, qThreads, qSmem));
    if (WOST && starSplit) ZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&starResident, waveStarPooledKernel<DIM>, qThreads, qSmem));
    const int residentScale = std::max(1, envInt("ZB_WAVE_PERSIST", 1)); // 0/1: one resident wave; k > 1: k waves' worth of CTAs
    closestResident = std::max(1, closestResident)*ctx->smCount*residentScale/nGroups;
    starResident = std::max(1, starResident)*ctx->smCount*residentScale/nGroups;
    if (envInt("ZB_WAVE_PERSIST", 1) == 0) closestResident = starResident = 1 << 30; // one CTA per 128 positions, as before

    // fork: the group streams start after everything already queued on the context's stream
    ZB_CUDA(cudaEventRecord(ctx->evFork, stream));
    for (int g = 0; g < nGroups; g++) {
        WaveGroup& G = groups[g];
        G.stream = g == 0 ? stream : ctx->waveStreams[g];
        if (g > 0) ZB_CUDA(cudaStreamWaitEvent(G.stream, ctx->evFork, 0));
        ZB_CUDA(cudaMemsetAsync(G.w.taskCounter, 0, sizeof(unsigned long long), G.stream));
        waveResetKernel<<<blocksFor(G.nSlots, 256), 256, 0, G.stream>>>(G.w, G.orderA);
        ZB_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }

    // ZB_WAVE_TRACE=1 (group 0 only): CUDA-event time and position bound of every round on stderr after the solve
    const bool traceRounds = envInt("ZB_WAVE_TRACE", 0) != 0;
    std::vector<cudaEvent_t> traceEvents;
    std::vector<int> traceBound;
    if (traceRounds) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, groups[0].stream); traceEvents.push_back(e); }
    int remaining = nGroups, rounds = 0;
    while (remaining > 0) {
        for (int g = 0; g < nGroups; g++) {
            WaveGroup& G = groups[g];
            if (G.done) continue;
            WaveSlots& w = G.w;
            for (int k = 0; k < kRoundsPerGroup; k++, G.round++) {
                w.order = G.orderA;
                w.live = G.live + 4*(G.round%kLiveRing);
                ZB_CUDA(cudaMemsetAsync(w.live, 0, 4*sizeof(unsigned int), G.stream));
                if (G.round > 0) { // round 0 only hands out the first tasks
                    // striped over all the group's warps while there are enough walks; in the tail over as many warps as
                    // there are walks, but never fewer than a few per SM-resident warp slot
                    const int qWarps = std::max(1, std::min(G.nSlots/32, std::max(G.nBound/32, stripeWarps)));
                    const int qBlocks = blocksFor(qWarps*32, qThreads);
                    const int virtualWarps = qBlocks*(qThreads/32);
                    // persistent grids: the CTAs the machine keeps resident draw the virtual warps from a ticket counter
                    waveClosestKernel<DIM, WOST><<<std::min(qBlocks, closestResident), qThreads, qSmem, G.stream>>>(a, w, G.nBound, virtualWarps);
                    if (starSplit) {
                        waveStarPrivateKernel<DIM><<<blocksFor(virtualWarps*32, 128), 128, 0, G.stream>>>(a, w, G.nBound, virtualWarps);
                        waveStarPooledKernel<DIM><<<std::min(qBlocks, starResident), qThreads, qSmem, G.stream>>>(a, w, G.nBound, virtualWarps);
                    } else if (star) {
                        waveStarKernel<DIM><<<std::min(qBlocks, starResident), qThreads, qSmem, G.stream>>>(a, w, G.nBound, virtualWarps);
                    }
                    ctx->launches += starSplit ? 3 : (star ? 2 : 1);
                }
                // the step kernel takes the round's walks in the order of the work they will do (wave.cuh: waveStepKeyKernel)
                const uint32_t *sortValues = G.orderA;
                if (stepSort && star && G.round > 0) {
                    waveStepKeyKernel<<<blocksFor(G.nBound, 256), 256, 0, G.stream>>>(w, G.nBound, G.stepKeys);
                    ZB_CUDA(cudaGetLastError());
                    size_t tmp = sortBytes;
                    ZB_CUDA(cub::DeviceRadixSort::SortPairs(G.sortTemp, tmp, G.stepKeys, G.stepKeysOut, G.orderA, G.orderStep, G.nBound, 0, 8, G.stream));
                    w.order = G.orderStep;
                    sortValues = G.orderStep;
                    ctx->launches += 2;
                }
                // solves with extra value channels run their own instantiations: the scalar kernels carry none of that code
                if (a.extra) waveStepKernel<DIM, WOST, true><<<blocksFor(G.nBound, 128), 128, 0, G.stream>>>(a, w, G.nBound);
                else waveStepKernel<DIM, WOST, false><<<blocksFor(G.nBound, 128), 128, 0, G.stream>>>(a, w, G.nBound);
                ZB_CUDA(cudaGetLastError());
                size_t bytes = sortBytes; // upper bound of this group's scratch
                ZB_CUDA(cub::DeviceRadixSort::SortPairs(G.sortTemp, bytes, w.keys, G.keysOut, sortValues, G.orderB, G.nBound, 0, 32, G.stream));
                std::swap(G.orderA, G.orderB);
                ctx->launches += 2; // the step kernel and one library call (cub's own kernels are not counted)
                if (traceRounds && g == 0) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, G.stream); traceEvents.push_back(e); traceBound.push_back(G.nBound); }
            }
            rounds = std::max(rounds, G.round);
            const int slot = 2*g + (G.batches & 1);
            ZB_CUDA(cudaMemcpyAsync(ctx->hostLive + slot, G.live + 4*((G.round - 1)%kLiveRing), sizeof(unsigned int), cudaMemcpyDeviceToHost, G.stream));
            ZB_CUDA(cudaEventRecord(ctx->evRound[slot], G.stream));
            if (G.batches >= 1) {
                // one batch of rounds of lag: the device never waits for the host
                const int prev = 2*g + ((G.batches - 1) & 1);
                ZB_CUDA(cudaEventSynchronize(ctx->evRound[prev]));
                unsigned int alive = ctx->hostLive[prev];
                if (alive == 0u) { G.done = true; remaining--; }
                else if ((int)alive <= std::min(tailWalks, tailCapacity) && (int)alive < G.nSlots) {
                    // fewer live walks than slots: the group's tasks have run out, and few enough are left that the round
                    // barriers cost more than they organise. One persistent launch steps them to their end (the walks alive
                    // NOW are a subset of the `alive` ones at the front of the order the last round left behind).
                    const int bound = std::min(G.nBound, (int)((alive + 127u)/128u*128u));
                    const int tailBlocks = std::max(1, std::min(ctx->smCount*tailPerSm/nGroups, blocksFor(bound*32, 128)));
                    w.order = G.orderA;
                    if (a.extra) waveTailKernel<DIM, WOST, true><<<tailBlocks, 128, kCoopSmem, G.stream>>>(a, w, bound, tailBlocks*4);
                    else waveTailKernel<DIM, WOST, false><<<tailBlocks, 128, kCoopSmem, G.stream>>>(a, w, bound, tailBlocks*4);
                    ZB_CUDA(cudaGetLastError());
                    ctx->launches += 1;
                    G.done = true; remaining--;
                }
                // the live walks sit at the front of the order (empty slots sort to the end) and, once the group's tasks have
                // run out, their number only falls: later rounds need not look past it
                else G.nBound = std::min(G.nBound, (int)((alive + 127u)/128u*128u));
            }
            G.batches++;
        }
    }
    // join: the context's stream continues after every group
    for (int g = 1; g < nGroups; g++) {
        ZB_CUDA(cudaEventRecord(ctx->evRound[2*g], groups[g].stream));
        ZB_CUDA(cudaStreamWaitEvent(stream, ctx->evRound[2*g], 0));
    }
    if (traceRounds) {
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, groups[0].stream); traceEvents.push_back(e); traceBound.push_back(-1);
        cudaEventSynchronize(e);
        for (size_t k = 1; k < traceEvents.size(); k++) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, traceEvents[k - 1], traceEvents[k]);
            fprintf(stderr, "[wave round] %4zu bound %8d %9.3f ms%s\n", k - 1, traceBound[k - 1], ms, traceBound[k - 1] < 0 ? " (tail kernel)" : "");
        }
        for (cudaEvent_t ev : traceEvents) cudaEventDestroy(ev);
    }
    ctx->lastRounds = rounds;
    ctx->lastGrid = nGroups;
    return ZB_OK;
}

template <int DIM>
int solveRange(zb_scene *scene, SolveArgs& a, int n, bool anyGradient)
{
    zb_context *ctx = scene->ctx;
    cudaStream_t stream = ctx->stream;

    // per-point scratch
    Carver pp(nullptr);
    pp.take<int>(n); pp.take<long long>(n + 1); pp.take<int>(n); pp.take<long long>(n + 1); pp.take<unsigned long long>(n);
    ZB_CUDA(ctx->perPoint.reserve(pp.offset + 256));
    Carver pc(ctx->perPoint.ptr);
    a.taskCount = pc.take<int>(n);
    a.taskOffset = pc.take<long long>(n + 1);
    a.stratCount = pc.take<int>(n);
    a.stratOffset = pc.take<long long>(n + 1);
    a.pointState = pc.take<unsigned long long>(n);
    a.pts.n = n;
    a.anyGradient = anyGradient ? 1 : 0;

    const int threads = 128;
    ZB_CUDA(cudaEventRecord(ctx->ev[0], stream));
    countTasksKernel<DIM><<<blocksFor(n, threads), threads, 0, stream>>>(a);
    ZB_CUDA(cudaGetLastError());

    // exclusive scans (n + 1 outputs: scanning n + 1 inputs whose last element is ignored)
    // cub accumulates in the type of the explicit init value (long long), not of the int inputs
    size_t tempBytes = 0;
    ZB_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tempBytes, a.taskCount, a.taskOffset, cub::Sum(), (long long)0, n, stream));
    ZB_CUDA(ctx->scanTemp.reserve(tempBytes + 256));
    ZB_CUDA(cub::DeviceScan::ExclusiveScan(ctx->scanTemp.ptr, tempBytes, a.taskCount, a.taskOffset, cub::Sum(), (long long)0, n, stream));
    ZB_CUDA(cub::DeviceScan::ExclusiveScan(ctx->scanTemp.ptr, tempBytes, a.stratCount, a.stratOffset, cub::Sum(), (long long)0, n, stream));
    // totals: last offset + last count
    long long lastOffsets[2];
    int lastCounts[2];
    ZB_CUDA(cudaMemcpyAsync(&lastOffsets[0], a.taskOffset + (n - 1), sizeof(long long), cudaMemcpyDeviceToHost, stream));
    ZB_CUDA(cudaMemcpyAsync(&lastOffsets[1], a.stratOffset + (n - 1), sizeof(long long), cudaMemcpyDeviceToHost, stream));
    ZB_CUDA(cudaMemcpyAsync(&lastCounts[0], a.taskCount + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    ZB_CUDA(cudaMemcpyAsync(&lastCounts[1], a.stratCount + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    ZB_CUDA(cudaStreamSynchronize(stream));
    long long totalTasks = lastOffsets[0] + lastCounts[0];
    long long totalStrat = lastOffsets[1] + lastCounts[1];
    ZB_CUDA(cudaMemcpyAsync(a.taskOffset + n, &totalTasks, sizeof(long long), cudaMemcpyHostToDevice, stream));
    ZB_CUDA(cudaMemcpyAsync(a.stratOffset + n, &totalStrat, sizeof(long long), cudaMemcpyHostToDevice, stream));
    a.totalTasks = totalTasks;

    ZB_CUDA(ctx->strat.reserve((size_t)totalStrat*sizeof(float) + 256));
    a.strat = ctx->strat.as<float>();

    size_t nt = (size_t)std::max<long long>(totalTasks, 1);
    Carver rc0(nullptr);
    rc0.take<float>(nt); rc0.take<int>(nt); rc0.take<int>(nt);
    if (anyGradient) { rc0.take<float>(nt); rc0.take<float>(nt*DIM); rc0.take<float>(nt*DIM); }
    ZB_CUDA(ctx->records.reserve(rc0.offset + 256));
    Carver rc(ctx->records.ptr);
    a.recSolution = rc.take<float>(nt);
    a.recLength = rc.take<int>(nt);
    a.recFlags = rc.take<int>(nt);
    if (anyGradient) {
        a.recFirstSource = rc.take<float>(nt);
        a.recGradBoundary = rc.take<float>(nt*DIM);
        a.recGradSource = rc.take<float>(nt*DIM);
    }
    a.taskCounter = ctx->counters.as<unsigned long long>();
    ZB_CUDA(cudaMemsetAsync(a.taskCounter, 0, sizeof(unsigned long long), stream));
    if (a.nExtra > 0) { // per-task sums and records of the extra value channels
        const size_t per = (size_t)a.nExtra*nt;
        ZB_CUDA(ctx->channelRecords.reserve(256 + 4*per*sizeof(float) + 256));
        a.xRefl = reinterpret_cast<float *>(ctx->channelRecords.as<char>() + 256);
        a.xSrc = a.xRefl + per;
        a.xRecSolution = a.xSrc + per;
        a.xRecFirstSource = a.xRecSolution + per;
        ZB_CUDA(cudaMemsetAsync(a.xRefl, 0, 4*per*sizeof(float), stream));
        ExtraChannelsView view{a.nExtra, totalTasks, a.extraPde, a.xRefl, a.xSrc, a.xRecSolution, a.xRecFirstSource};
        ZB_CUDA(cudaMemcpyAsync(ctx->channelRecords.ptr, &view, sizeof(view), cudaMemcpyHostToDevice, stream)); // staged before the call returns
        a.extra = ctx->channelRecords.as<ExtraChannelsView>();
    }

    ZB_CUDA(cudaFuncSetAttribute(preparePointsKernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
    preparePointsKernel<DIM><<<blocksFor(n, threads), threads, kCoopSmem, stream>>>(a);
    ZB_CUDA(cudaGetLastError());

    ZB_CUDA(cudaEventRecord(ctx->ev[1], stream));
    ctx->lastTasks = totalTasks;
    ctx->lastGrid = 0;
    ctx->launches += 5; // count, 2 scans (init + scan kernels counted once each), prepare, reduce
    if (totalTasks > 0) {
        if (a.extra && envInt("ZB_SCHEDULER_PERSISTENT", 0) != 0)
            return fail(ZB_ERR_UNSUPPORTED, "zb_solve_channels: the persistent debug scheduler does not carry value channels");
        if (envInt("ZB_SCHEDULER_PERSISTENT", 0) == 0) {
            int rc = a.estimator == ZB_WALK_ON_STARS ? runWavefront<DIM, true>(scene, a, totalTasks) : runWavefront<DIM, false>(scene, a, totalTasks);
            if (rc != ZB_OK) return rc;
        } else {
        // persistent grid: a whole number of waves of 128-thread CTAs over all SMs
        int perSm = 0;
        if (a.estimator == ZB_WALK_ON_STARS) {
            ZB_CUDA(cudaFuncSetAttribute(walkKernel<DIM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
            ZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, walkKernel<DIM, true>, 128, kCoopSmem));
        } else {
            ZB_CUDA(cudaFuncSetAttribute(walkKernel<DIM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
            ZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, walkKernel<DIM, false>, 128, kCoopSmem));
        }
        perSm = std::max(1, perSm);
        long long wantBlocks = (totalTasks + 127)/128;
        int grid = (int)std::min<long long>((long long)ctx->smCount*perSm, wantBlocks);
        grid = std::max(1, grid);
        if (a.estimator == ZB_WALK_ON_STARS) walkKernel<DIM, true><<<grid, 128, kCoopSmem, stream>>>(a);
        else walkKernel<DIM, false><<<grid, 128, kCoopSmem, stream>>>(a);
        ZB_CUDA(cudaGetLastError());
        ctx->lastGrid = grid;
        ctx->launches += 1;
        }
    }
    ZB_CUDA(cudaEventRecord(ctx->ev[2], stream));

    if (a.rwsRecords) rwsFinishKernel<DIM><<<blocksFor(n, threads), threads, 0, stream>>>(a);
    else reducePointsKernel<DIM><<<blocksFor(n, threads), threads, 0, stream>>>(a);
    ZB_CUDA(cudaGetLastError());
    if (a.nExtra > 0 && !a.rwsRecords) {
        reduceExtraChannelsKernel<DIM><<<blocksFor(n*a.nExtra, threads), threads, 0, stream>>>(a);
        ZB_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    ZB_CUDA(cudaEventRecord(ctx->ev[3], stream));
    ctx->profiled = true;
    return ZB_OK;
}

} // namespace

extern "C" {

const char *zb_last_error(void) { return lastErrorRef().c_str(); }
int zb_abi_version(void) { return ZB_ABI_VERSION; }

int zb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int zb_context_create(int device, zb_context **out)
{
    if (!out) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_context_create: out is NULL");
    *out = nullptr;
    int n = zb_device_count();
    if (n <= 0) return fail(ZB_ERR_NO_DEVICE, "zb_context_create: no CUDA device is visible (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_context_create: device index out of range");
    ZB_CUDA(cudaSetDevice(device));
    zb_context *ctx = new zb_context();
    ctx->device = device;
    cudaDeviceProp prop;
    ZB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->smCount = prop.multiProcessorCount;
    ZB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->ownsStream = true;
    ZB_CUDA(ctx->status.reserve(256));
    ZB_CUDA(ctx->counters.reserve(256));
    ZB_CUDA(cudaMemset(ctx->status.ptr, 0, 256));
    ZB_CUDA(cudaMemset(ctx->counters.ptr, 0, 256));
    ZB_CUDA(cudaMallocHost(&ctx->hostStatus, 64));
    for (int k = 0; k < 4; k++) ZB_CUDA(cudaEventCreate(&ctx->ev[k]));
    for (int k = 0; k < 2; k++) ZB_CUDA(cudaEventCreate(&ctx->evSplat[k]));
    *out = ctx;
    return ZB_OK;
}

void zb_context_destroy(zb_context *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->ownsStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->hostStatus) cudaFreeHost(ctx->hostStatus);
    if (ctx->hostLive) cudaFreeHost(ctx->hostLive);
    for (int k = 0; k < 16; k++) if (ctx->evRound[k]) cudaEventDestroy(ctx->evRound[k]);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    for (int k = 0; k < 8; k++) if (ctx->waveStreams[k]) cudaStreamDestroy(ctx->waveStreams[k]);
    for (int k = 0; k < 4; k++) if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    for (int k = 0; k < 2; k++) if (ctx->evSplat[k]) cudaEventDestroy(ctx->evSplat[k]);
    delete ctx;
}

int zb_context_set_stream(zb_context *ctx, void *cudaStream)
{
    if (!ctx) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_context_set_stream: ctx is NULL");
    if (ctx->ownsStream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    ctx->stream = reinterpret_cast<cudaStream_t>(cudaStream);
    ctx->ownsStream = false;
    return ZB_OK;
}

int zb_context_synchronize(zb_context *ctx)
{
    if (!ctx) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_context_synchronize: ctx is NULL");
    ZB_CUDA(cudaSetDevice(ctx->device));
    ZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZB_OK;
}

int zb_context_last_solve_profile(zb_context *ctx, float *prepareMs, float *walkMs, float *reduceMs,
                                  int64_t *walks, int32_t *walkGrid, int64_t *kernelLaunches)
{
    if (!ctx) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_context_last_solve_profile: ctx is NULL");
    if (!ctx->profiled) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_context_last_solve_profile: no solve has run on this context");
    ZB_CUDA(cudaSetDevice(ctx->device));
    ZB_CUDA(cudaEventSynchronize(ctx->ev[3]));
    float t01 = 0, t12 = 0, t23 = 0;
    ZB_CUDA(cudaEventElapsedTime(&t01, ctx->ev[0], ctx->ev[1]));
    ZB_CUDA(cudaEventElapsedTime(&t12, ctx->ev[1], ctx->ev[2]));
    ZB_CUDA(cudaEventElapsedTime(&t23, ctx->ev[2], ctx->ev[3]));
    if (prepareMs) *prepareMs = t01;
    if (walkMs) *walkMs = t12;
    if (reduceMs) *reduceMs = t23;
    if (walks) *walks = ctx->lastTasks;
    if (walkGrid) *walkGrid = ctx->lastGrid;
    if (kernelLaunches) *kernelLaunches = ctx->launches;
    return ZB_OK;
}

int zb_context_device(const zb_context *ctx) { return ctx ? ctx->device : -1; }
int zb_context_sm_count(const zb_context *ctx) { return ctx ? ctx->smCount : 0; }

namespace {

// FCPW's SilhouetteEdge test (silhouette_edges.inl:83-111, Neumann handler, 3D) measures the view direction
// with the opposite sign of Zombie's ReflectanceTriangle (reflectance_boundary_bvh/geometry.h:196-206); the 2D
// SilhouetteVertex and ReflectanceLineSegment tests agree (silhouette_vertices.inl:59-85, geometry.h:108-126)
int snchSignOf(const zb_scene *scene, int boundary)
{
    return boundary == ZB_REFLECTING && scene->reflectingKind == ZB_REFLECTING_NEUMANN && scene->dim == 3 ? -1 : 1;
}

// Builds (or rebuilds) one boundary's hierarchy with the chosen builder. positions / indices / coefficients are HOST
// arrays; indices and coefficients may be NULL on a rebuild (the copies kept on the device are used).
int buildBoundary(zb_scene *scene, int boundary, int builder, const float *positions, int nVert, const int32_t *indices, int nPrim,
                  const float *minCoeff, const float *maxCoeff)
{
    zb_context *ctx = scene->ctx;
    DeviceBvhStorage& s = scene->boundary[boundary];
    const int dim = scene->dim;
    const char *which = boundary == ZB_ABSORBING ? "absorbing mesh: " : "reflecting mesh: ";
    s.reflecting = boundary == ZB_REFLECTING;
    s.robin = s.reflecting && scene->reflectingKind == ZB_REFLECTING_ROBIN;
    if (nVert <= 0 || nPrim <= 0) return fail(ZB_ERR_INVALID_ARGUMENT, std::string("zb_scene_create: ") + which + "empty mesh");
    // the mesh stays on the device: vertices that move later are rebuilt there (zb_scene_update_vertices)
    s.nVert = nVert;
    ZB_CUDA(s.meshPos.reserve((size_t)nVert*dim*sizeof(float)));
    ZB_CUDA(cudaMemcpyAsync(s.meshPos.ptr, positions, (size_t)nVert*dim*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (indices) {
        ZB_CUDA(s.meshIdx.reserve((size_t)nPrim*dim*sizeof(int32_t)));
        ZB_CUDA(cudaMemcpyAsync(s.meshIdx.ptr, indices, (size_t)nPrim*dim*sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (s.robin && minCoeff && maxCoeff) {
        ZB_CUDA(s.coefMin.reserve((size_t)nPrim*sizeof(float)));
        ZB_CUDA(s.coefMax.reserve((size_t)nPrim*sizeof(float)));
        ZB_CUDA(cudaMemcpyAsync(s.coefMin.ptr, minCoeff, (size_t)nPrim*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ZB_CUDA(cudaMemcpyAsync(s.coefMax.ptr, maxCoeff, (size_t)nPrim*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (builder == ZB_BVH_BUILDER_DEVICE) {
        int rc = buildBvhOnDevice(ctx, s, dim, nPrim, scene->solveDoubleSided != 0, scene->silhouettePrecision, snchSignOf(scene, boundary));
        if (rc != ZB_OK) lastErrorRef() = std::string("zb_scene_create: ") + which + lastErrorRef();
        return rc;
    }
    if (!indices) return fail(ZB_ERR_INVALID_ARGUMENT, "host BVH builder needs the index array");
    cudaEvent_t ev0, ev1;
    ZB_CUDA(cudaEventCreate(&ev0));
    ZB_CUDA(cudaEventCreate(&ev1));
    ZB_CUDA(cudaEventRecord(ev0, ctx->stream));
    ReflectingOptions opt;
    if (s.reflecting) {
        opt.enabled = true;
        opt.solveDoubleSided = scene->solveDoubleSided != 0;
        opt.silhouettePrecision = scene->silhouettePrecision;
        opt.minCoeff = s.robin ? minCoeff : nullptr;
        opt.maxCoeff = s.robin ? maxCoeff : nullptr;
    }
    const char *err = nullptr;
    if (!buildWideBvh(dim, positions, nVert, indices, nPrim, opt, s.host, &err))
        return fail(ZB_ERR_INVALID_ARGUMENT, std::string("zb_scene_create: ") + which + (err ? err : "?"));
    int rc = uploadBvh(ctx, s, snchSignOf(scene, boundary));
    if (rc != ZB_OK) return rc;
    ZB_CUDA(cudaEventRecord(ev1, ctx->stream));
    ZB_CUDA(cudaEventSynchronize(ev1));
    ZB_CUDA(cudaEventElapsedTime(&s.buildMs, ev0, ev1)); // the stream waits for the host build, so this spans build + upload
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return ZB_OK;
}

} // namespace

int zb_scene_create(zb_context *ctx, int dim,
                    const float *absorbingPositions, int nAbsorbingVertices,
                    const int32_t *absorbingIndices, int nAbsorbingPrimitives,
                    const float *reflectingPositions, int nReflectingVertices,
                    const int32_t *reflectingIndices, int nReflectingPrimitives,
                    int reflectingKind, const float *minRobinCoeff, const float *maxRobinCoeff,
                    int solveDoubleSided, float silhouettePrecision,
                    int domainIsWatertight, const float *domainMin, const float *domainMax,
                    zb_scene **out)
{
    if (!ctx || !out) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_create: ctx/out is NULL");
    *out = nullptr;
    if (dim != 2 && dim != 3) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_create: dim must be 2 or 3");
    if (nAbsorbingPrimitives < 0 || nReflectingPrimitives < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_create: negative primitive count");
    if (nReflectingPrimitives > 0 && reflectingKind != ZB_REFLECTING_NEUMANN && reflectingKind != ZB_REFLECTING_ROBIN)
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_create: reflecting mesh given but reflectingKind is not NEUMANN or ROBIN");
    if (nReflectingPrimitives > 0 && reflectingKind == ZB_REFLECTING_ROBIN && (!minRobinCoeff || !maxRobinCoeff))
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_create: Robin boundary needs min/max coefficient arrays");
    ZB_CUDA(cudaSetDevice(ctx->device));

    zb_scene *scene = new zb_scene();
    scene->ctx = ctx;
    scene->dim = dim;
    scene->watertight = domainIsWatertight;
    scene->solveDoubleSided = solveDoubleSided;
    scene->silhouettePrecision = silhouettePrecision;
    scene->reflectingKind = nReflectingPrimitives > 0 ? reflectingKind : ZB_REFLECTING_NONE;
    for (int k = 0; k < 3; k++) {
        // defaults of GeometricQueries() (geometric_queries.h:128-137)
        scene->domainMin[k] = domainMin && k < dim ? domainMin[k] : -kMaxFloat;
        scene->domainMax[k] = domainMax && k < dim ? domainMax[k] : kMaxFloat;
    }

    if (nAbsorbingPrimitives > 0) {
        int rc = buildBoundary(scene, ZB_ABSORBING, ctx->bvhBuilder, absorbingPositions, nAbsorbingVertices, absorbingIndices,
                               nAbsorbingPrimitives, nullptr, nullptr);
        if (rc != ZB_OK) { delete scene; return rc; }
    }
    if (nReflectingPrimitives > 0) {
        const bool robin = reflectingKind == ZB_REFLECTING_ROBIN;
        int rc = buildBoundary(scene, ZB_REFLECTING, ctx->bvhBuilder, reflectingPositions, nReflectingVertices, reflectingIndices,
                               nReflectingPrimitives, robin ? minRobinCoeff : nullptr, robin ? maxRobinCoeff : nullptr);
        if (rc != ZB_OK) { delete scene; return rc; }
    }
    *out = scene;
    return ZB_OK;
}

void zb_scene_destroy(zb_scene *scene)
{
    if (!scene) return;
    cudaSetDevice(scene->ctx->device);
    cudaStreamSynchronize(scene->ctx->stream);
    delete scene;
}

// ---- Robin coefficient refit on the device -----------------------------------------------------------------
// FcpwRobinBoundaryHandler::updateCoefficientValues -> ReflectanceBvh::updateCoefficients (reflectance_boundary_bvh/bvh.h:304-367):
// new per-primitive coefficient bounds, then every node's range bottom-up. The hierarchy is stored breadth first, so the
// refit is one kernel per level from the deepest up; a node reads the ranges its internal children wrote one launch before.
__global__ void refitPrimCoeffKernel(int nPrims, int dim, const int *primIndex, const float *minCoeff, const float *maxCoeff,
                                     PrimAdjacency *adjacency, float *prims)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= nPrims) return;
    int p = primIndex[i];
    adjacency[i].minCoeff = minCoeff[p];
    adjacency[i].maxCoeff = maxCoeff[p];
    if (dim == 3) prims[(size_t)12*i + 7] = maxCoeff[p]; // the copy the filter stage reads from the vertex record (v1.w)
}

__global__ void refitNodeCoeffKernel(const WideNode *nodes, ChildCone *cones, const PrimAdjacency *adjacency, float2 *nodeRange,
                                     int first, int count)
{
    int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int w = first + k;
    const WideNode wn = nodes[w];
    float nodeMin = kMaxFloat, nodeMax = -kMaxFloat;
    for (int c = 0; c < kWidth; c++) {
        uint32_t m = (wn.w[6 + (c >> 2)] >> (8*(c & 3))) & 0xffu;
        if (m == 0u) continue;
        float mn = kMaxFloat, mx = -kMaxFloat;
        int n = (int)(m >> 5);
        if (n > 0) {
            int firstPrim = (int)wn.w[5] + (int)(m & 31u);
            for (int i = 0; i < n; i++) {
                mn = fminf(mn, adjacency[firstPrim + i].minCoeff);
                mx = fmaxf(mx, adjacency[firstPrim + i].maxCoeff);
            }
        } else {
            float2 r = nodeRange[wn.w[4] + (m & 31u) - 1u];
            mn = r.x; mx = r.y;
        }
        cones[(size_t)w*kWidth + c].minCoeff = mn;
        cones[(size_t)w*kWidth + c].maxCoeff = mx;
        nodeMin = fminf(nodeMin, mn);
        nodeMax = fmaxf(nodeMax, mx);
    }
    nodeRange[w] = make_float2(nodeMin, nodeMax);
}

namespace {

// level boundaries of the breadth-first node array
bool computeLevels(const HostBvh& h, std::vector<int>& levelStart)
{
    const size_t n = h.nodes.size();
    std::vector<int> depth(n, 0);
    for (size_t w = 0; w < n; w++) {
        const uint8_t *meta = reinterpret_cast<const uint8_t *>(&h.nodes[w].w[6]);
        for (int c = 0; c < kWidth; c++) {
            uint8_t m = meta[c];
            if (m != 0 && (m >> 5) == 0) {
                size_t child = h.nodes[w].w[4] + (m & 31) - 1;
                if (child <= w || child >= n) return false;
                depth[child] = depth[w] + 1;
            }
        }
    }
    levelStart.assign(1, 0);
    for (size_t w = 1; w < n; w++) {
        if (depth[w] < depth[w - 1]) return false; // not breadth first
        if (depth[w] > depth[w - 1]) levelStart.push_back((int)w);
    }
    levelStart.push_back((int)n);
    return true;
}

// coefficient arrays already on the device (caller order, one value per input primitive)
int refitRobinCoefficients(zb_scene *scene, const float *devMin, const float *devMax)
{
    DeviceBvhStorage& s = scene->boundary[ZB_REFLECTING];
    zb_context *ctx = scene->ctx;
    if (s.levelStart.empty() && !computeLevels(s.host, s.levelStart))
        return fail(ZB_ERR_UNSUPPORTED, "zb_scene_update_robin_coefficients: node array is not breadth first");
    const int nPrims = s.host.nPrims;
    const size_t nNodes = (size_t)s.nNodes;
    float2 *nodeRange = reinterpret_cast<float2 *>(s.coeffScratch.as<char>() + (((size_t)2*nPrims*sizeof(float) + 255) & ~size_t(255)));
    refitPrimCoeffKernel<<<(nPrims + 255)/256, 256, 0, ctx->stream>>>(nPrims, s.host.dim, s.primIndex.as<int>(), devMin, devMax,
                                                                      s.adjacency.as<PrimAdjacency>(), s.prims.as<float>());
    ZB_CUDA(cudaGetLastError());
    for (int l = (int)s.levelStart.size() - 2; l >= 0; l--) {
        const int first = s.levelStart[l], count = s.levelStart[l + 1] - first;
        refitNodeCoeffKernel<<<(count + 127)/128, 128, 0, ctx->stream>>>(s.nodes.as<WideNode>(), s.cones.as<ChildCone>(),
                                                                        s.adjacency.as<PrimAdjacency>(), nodeRange, first, count);
        ZB_CUDA(cudaGetLastError());
    }
    (void)nNodes;
    // the input-order copies a later rebuild (zb_scene_update_vertices) reads
    if (s.coefMin.ptr && devMin != s.coefMin.as<float>()) {
        ZB_CUDA(cudaMemcpyAsync(s.coefMin.ptr, devMin, (size_t)nPrims*sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        ZB_CUDA(cudaMemcpyAsync(s.coefMax.ptr, devMax, (size_t)nPrims*sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return ZB_OK;
}

int reserveCoeffScratch(DeviceBvhStorage& s)
{
    const size_t primBytes = ((size_t)2*s.host.nPrims*sizeof(float) + 255) & ~size_t(255);
    ZB_CUDA(s.coeffScratch.reserve(primBytes + (size_t)s.nNodes*sizeof(float2) + 256));
    return ZB_OK;
}

} // namespace

int zb_scene_update_robin_coefficients_dev(zb_scene *scene, const float *minRobinCoeff, const float *maxRobinCoeff)
{
    if (!scene || !minRobinCoeff || !maxRobinCoeff) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_robin_coefficients_dev: NULL argument");
    DeviceBvhStorage& s = scene->boundary[ZB_REFLECTING];
    if (!s.present || scene->reflectingKind != ZB_REFLECTING_ROBIN)
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_robin_coefficients_dev: scene has no Robin boundary");
    ZB_CUDA(cudaSetDevice(scene->ctx->device));
    int rc = reserveCoeffScratch(s);
    if (rc != ZB_OK) return rc;
    return refitRobinCoefficients(scene, minRobinCoeff, maxRobinCoeff); // asynchronous on the context's stream
}

int zb_scene_update_robin_coefficients(zb_scene *scene, const float *minRobinCoeff, const float *maxRobinCoeff)
{
    if (!scene || !minRobinCoeff || !maxRobinCoeff) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_robin_coefficients: NULL argument");
    DeviceBvhStorage& s = scene->boundary[ZB_REFLECTING];
    if (!s.present || scene->reflectingKind != ZB_REFLECTING_ROBIN)
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_robin_coefficients: scene has no Robin boundary");
    ZB_CUDA(cudaSetDevice(scene->ctx->device));
    // 8 bytes per primitive go up; the per-primitive records and the hierarchy's coefficient ranges are refitted on the device
    // (round 1 recomputed them on the host and uploaded the cone, adjacency and vertex arrays again: ~100 MB for 655 K triangles)
    int rc = reserveCoeffScratch(s);
    if (rc != ZB_OK) return rc;
    const size_t bytes = (size_t)s.host.nPrims*sizeof(float);
    float *devMin = s.coeffScratch.as<float>(), *devMax = devMin + s.host.nPrims;
    ZB_CUDA(cudaMemcpyAsync(devMin, minRobinCoeff, bytes, cudaMemcpyHostToDevice, scene->ctx->stream));
    ZB_CUDA(cudaMemcpyAsync(devMax, maxRobinCoeff, bytes, cudaMemcpyHostToDevice, scene->ctx->stream));
    rc = refitRobinCoefficients(scene, devMin, devMax);
    if (rc != ZB_OK) return rc;
    ZB_CUDA(cudaStreamSynchronize(scene->ctx->stream));
    return ZB_OK;
}

// ---- PDE data grids (zb_grid_view::data) -------------------------------------------------------------------
struct zb_grid {
    int device = 0;
    float *data = nullptr;
};

int zb_grid_create(zb_context *ctx, const float *hostData, const int32_t shape[3], zb_grid **out)
{
    if (!ctx || !hostData || !shape || !out) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_grid_create: NULL argument");
    *out = nullptr;
    if (shape[0] < 1 || shape[1] < 1 || shape[2] < 1) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_grid_create: every extent must be at least 1");
    ZB_CUDA(cudaSetDevice(ctx->device));
    size_t n = (size_t)shape[0]*(size_t)shape[1]*(size_t)shape[2];
    zb_grid *g = new zb_grid();
    g->device = ctx->device;
    cudaError_t e = cudaMalloc(&g->data, n*sizeof(float));
    if (e != cudaSuccess) { delete g; ZB_CUDA(e); }
    e = cudaMemcpy(g->data, hostData, n*sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(g->data); delete g; ZB_CUDA(e); }
    *out = g;
    return ZB_OK;
}

const float *zb_grid_data(const zb_grid *grid) { return grid ? grid->data : nullptr; }

void zb_grid_destroy(zb_grid *grid)
{
    if (!grid) return;
    cudaSetDevice(grid->device);
    cudaFree(grid->data);
    delete grid;
}

int zb_scene_stats(const zb_scene *scene, int boundary, int64_t *bytes, int32_t *nNodes, int32_t *nLeaves, int32_t *maxDepth)
{
    if (!scene || boundary < 0 || boundary > 1) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_stats: invalid argument");
    const DeviceBvhStorage& s = scene->boundary[boundary];
    if (bytes) *bytes = s.present ? (int64_t)s.bytes() : 0;
    if (nNodes) *nNodes = s.present ? (int32_t)s.nNodes : 0;
    if (nLeaves) *nLeaves = s.present ? s.host.nLeaves : 0;
    if (maxDepth) *maxDepth = s.present ? s.host.maxDepth : 0;
    return ZB_OK;
}

// ---- on-device hierarchy build, moving geometry ------------------------------------------------------------------
int zb_context_set_bvh_builder(zb_context *ctx, int builder)
{
    if (!ctx || (builder != ZB_BVH_BUILDER_HOST_SAH && builder != ZB_BVH_BUILDER_DEVICE))
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_context_set_bvh_builder: invalid argument");
    ctx->bvhBuilder = builder;
    return ZB_OK;
}

int zb_scene_build_profile(const zb_scene *scene, int boundary, float *ms, int32_t *builder)
{
    if (!scene || boundary < 0 || boundary > 1) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_build_profile: invalid argument");
    const DeviceBvhStorage& s = scene->boundary[boundary];
    if (ms) *ms = s.present ? s.buildMs : 0.0f;
    if (builder) *builder = s.builder;
    return ZB_OK;
}

namespace {
int rebuildMoved(zb_scene *scene, int boundary)
{
    DeviceBvhStorage& s = scene->boundary[boundary];
    int rc = buildBvhOnDevice(scene->ctx, s, scene->dim, s.host.nPrims, scene->solveDoubleSided != 0, scene->silhouettePrecision,
                              snchSignOf(scene, boundary));
    if (rc != ZB_OK) lastErrorRef() = "zb_scene_update_vertices: " + lastErrorRef();
    return rc;
}
} // namespace

int zb_scene_update_vertices_dev(zb_scene *scene, int boundary, const float *positions, int nVertices)
{
    if (!scene || boundary < 0 || boundary > 1 || !positions) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_vertices_dev: invalid argument");
    DeviceBvhStorage& s = scene->boundary[boundary];
    if (!s.present) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_vertices_dev: the scene has no such boundary");
    if (nVertices != s.nVert) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_vertices_dev: vertex count differs from the scene's");
    ZB_CUDA(cudaSetDevice(scene->ctx->device));
    ZB_CUDA(cudaMemcpyAsync(s.meshPos.ptr, positions, (size_t)nVertices*scene->dim*sizeof(float), cudaMemcpyDeviceToDevice, scene->ctx->stream));
    return rebuildMoved(scene, boundary);
}

int zb_scene_update_vertices(zb_scene *scene, int boundary, const float *positions, int nVertices)
{
    if (!scene || boundary < 0 || boundary > 1 || !positions) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_vertices: invalid argument");
    DeviceBvhStorage& s = scene->boundary[boundary];
    if (!s.present) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_vertices: the scene has no such boundary");
    if (nVertices != s.nVert) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_update_vertices: vertex count differs from the scene's");
    ZB_CUDA(cudaSetDevice(scene->ctx->device));
    ZB_CUDA(cudaMemcpyAsync(s.meshPos.ptr, positions, (size_t)nVertices*scene->dim*sizeof(float), cudaMemcpyHostToDevice, scene->ctx->stream));
    return rebuildMoved(scene, boundary);
}

int zb_scene_validate(zb_scene *scene, int boundary, const float *positions, int nVertices, const int32_t *indices, int nPrimitives,
                      int64_t *violations)
{
    if (!scene || boundary < 0 || boundary > 1 || !positions || !indices || !violations)
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_validate: invalid argument");
    DeviceBvhStorage& s = scene->boundary[boundary];
    if (!s.present || nPrimitives != s.host.nPrims || nVertices != s.nVert)
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_validate: mesh does not match the scene's boundary");
    ZB_CUDA(cudaSetDevice(scene->ctx->device));
    HostBvh h;
    h.dim = scene->dim;
    h.nPrims = s.host.nPrims;
    h.nodes.resize((size_t)s.nNodes);
    h.prims.resize((size_t)(scene->dim == 3 ? 12 : 4)*h.nPrims);
    h.primIndex.resize((size_t)h.nPrims);
    cudaStream_t st = scene->ctx->stream;
    ZB_CUDA(cudaMemcpyAsync(h.nodes.data(), s.nodes.ptr, h.nodes.size()*sizeof(WideNode), cudaMemcpyDeviceToHost, st));
    ZB_CUDA(cudaMemcpyAsync(h.prims.data(), s.prims.ptr, h.prims.size()*sizeof(float), cudaMemcpyDeviceToHost, st));
    ZB_CUDA(cudaMemcpyAsync(h.primIndex.data(), s.primIndex.ptr, h.primIndex.size()*sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    ZB_CUDA(cudaStreamSynchronize(st));
    long long bad = 0;
    for (int32_t p : h.primIndex) if (p < 0 || p >= nPrimitives) bad++;
    if (bad == 0) bad = validateWideBvh(h, scene->dim, positions, indices);
    // breadth-first level table: level l + 1 holds exactly the internal children of level l (the coefficient refit relies on it)
    if (!s.levelStart.empty()) {
        std::vector<int> levels;
        if (!computeLevels(h, levels) || levels != s.levelStart) bad++;
    }
    *violations = bad;
    return ZB_OK;
}

int zb_scene_download_records(zb_scene *scene, int boundary, int32_t *primIndex, float *prims, float *adjacency)
{
    if (!scene || boundary < 0 || boundary > 1) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_download_records: invalid argument");
    DeviceBvhStorage& s = scene->boundary[boundary];
    if (!s.present) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_download_records: the scene has no such boundary");
    ZB_CUDA(cudaSetDevice(scene->ctx->device));
    cudaStream_t st = scene->ctx->stream;
    const size_t n = (size_t)s.host.nPrims;
    if (primIndex) ZB_CUDA(cudaMemcpyAsync(primIndex, s.primIndex.ptr, n*sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (prims) ZB_CUDA(cudaMemcpyAsync(prims, s.prims.ptr, n*(scene->dim == 3 ? 12 : 4)*sizeof(float), cudaMemcpyDeviceToHost, st));
    if (adjacency) {
        if (!s.view.adjacency) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_scene_download_records: not a reflecting boundary");
        ZB_CUDA(cudaMemcpyAsync(adjacency, s.adjacency.ptr, n*sizeof(PrimAdjacency), cudaMemcpyDeviceToHost, st));
    }
    ZB_CUDA(cudaStreamSynchronize(st));
    return ZB_OK;
}

int zb_host_bvh_check(int dim, const float *positions, int nVertices, const int32_t *indices, int nPrimitives,
                      int reflecting, int32_t *nNodes, int32_t *nLeaves, int32_t *maxDepth, int64_t *violations)
{
    HostBvh h;
    ReflectingOptions opt;
    opt.enabled = reflecting != 0;
    const char *err = nullptr;
    if (!buildWideBvh(dim, positions, nVertices, indices, nPrimitives, opt, h, &err))
        return fail(ZB_ERR_INVALID_ARGUMENT, std::string("zb_host_bvh_check: ") + (err ? err : "?"));
    if (nNodes) *nNodes = (int32_t)h.nodes.size();
    if (nLeaves) *nLeaves = h.nLeaves;
    if (maxDepth) *maxDepth = h.maxDepth;
    if (violations) *violations = validateWideBvh(h, dim, positions, indices);
    return ZB_OK;
}

#define ZB_QUERY_PROLOGUE(name)                                                                         \
    if (!scene) return fail(ZB_ERR_INVALID_ARGUMENT, name ": scene is NULL");                          \
    if (n < 0) return fail(ZB_ERR_INVALID_ARGUMENT, name ": n < 0");                                   \
    if (n == 0) return ZB_OK;                                                                          \
    zb_context *ctx = scene->ctx;                                                                      \
    ZB_CUDA(cudaSetDevice(ctx->device));                                                               \
    const int dim = scene->dim;                                                                        \
    (void)dim;

#define ZB_BOUNDARY(name)                                                                               \
    if (boundary < 0 || boundary > 1 || !scene->boundary[boundary].present)                            \
        return fail(ZB_ERR_INVALID_ARGUMENT, name ": the requested boundary is empty");                \
    const BvhView bvh = scene->boundary[boundary].view;

int zb_closest_point(zb_scene *scene, int boundary, int n, const float *x, const float *maxRadius,
                     int32_t *prim, float *dist, float *pt, float *normal)
{
    ZB_QUERY_PROLOGUE("zb_closest_point")
    ZB_BOUNDARY("zb_closest_point")
    if (!x || !prim || !dist) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_closest_point: x/prim/dist is NULL");
    IoPlan plan;
    size_t oX = plan.addIn(x, sizeof(float)*dim*n), oR = plan.addIn(maxRadius, maxRadius ? sizeof(float)*n : 0);
    size_t oP = plan.addOut(prim, sizeof(int32_t)*n), oD = plan.addOut(dist, sizeof(float)*n);
    size_t oPt = plan.addOut(pt, pt ? sizeof(float)*dim*n : 0), oN = plan.addOut(normal, normal ? sizeof(float)*dim*n : 0);
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    int threads = 128, blocks = blocksFor(n, threads);
    ZB_CUDA(cudaFuncSetAttribute(closestPointKernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
    ZB_CUDA(cudaFuncSetAttribute(closestPointKernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
    if (dim == 3) closestPointKernel<3><<<blocks, threads, kCoopSmem, ctx->stream>>>(bvh, n, at<float>(ctx, oX, x), at<float>(ctx, oR, maxRadius),
        at<int>(ctx, oP, prim), at<float>(ctx, oD, dist), at<float>(ctx, oPt, pt), at<float>(ctx, oN, normal), ctx->status.as<int>());
    else closestPointKernel<2><<<blocks, threads, kCoopSmem, ctx->stream>>>(bvh, n, at<float>(ctx, oX, x), at<float>(ctx, oR, maxRadius),
        at<int>(ctx, oP, prim), at<float>(ctx, oD, dist), at<float>(ctx, oPt, pt), at<float>(ctx, oN, normal), ctx->status.as<int>());
    ZB_CUDA(cudaGetLastError());
    rc = stageOut(ctx, plan);
    if (rc != ZB_OK) return rc;
    return checkStatus(ctx, "zb_closest_point");
}

int zb_intersect_ray(zb_scene *scene, int boundary, int n, const float *origin, const float *dir, const float *tMax,
                     int32_t *prim, float *dist, float *pt, float *normal)
{
    ZB_QUERY_PROLOGUE("zb_intersect_ray")
    ZB_BOUNDARY("zb_intersect_ray")
    if (!origin || !dir || !prim || !dist) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_intersect_ray: origin/dir/prim/dist is NULL");
    IoPlan plan;
    size_t oO = plan.addIn(origin, sizeof(float)*dim*n), oDir = plan.addIn(dir, sizeof(float)*dim*n), oT = plan.addIn(tMax, tMax ? sizeof(float)*n : 0);
    size_t oP = plan.addOut(prim, sizeof(int32_t)*n), oD = plan.addOut(dist, sizeof(float)*n);
    size_t oPt = plan.addOut(pt, pt ? sizeof(float)*dim*n : 0), oN = plan.addOut(normal, normal ? sizeof(float)*dim*n : 0);
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    int threads = 128, blocks = blocksFor(n, threads);
    if (dim == 3) intersectRayKernel<3><<<blocks, threads, 0, ctx->stream>>>(bvh, n, at<float>(ctx, oO, origin), at<float>(ctx, oDir, dir), at<float>(ctx, oT, tMax),
        at<int>(ctx, oP, prim), at<float>(ctx, oD, dist), at<float>(ctx, oPt, pt), at<float>(ctx, oN, normal), ctx->status.as<int>());
    else intersectRayKernel<2><<<blocks, threads, 0, ctx->stream>>>(bvh, n, at<float>(ctx, oO, origin), at<float>(ctx, oDir, dir), at<float>(ctx, oT, tMax),
        at<int>(ctx, oP, prim), at<float>(ctx, oD, dist), at<float>(ctx, oPt, pt), at<float>(ctx, oN, normal), ctx->status.as<int>());
    ZB_CUDA(cudaGetLastError());
    rc = stageOut(ctx, plan);
    if (rc != ZB_OK) return rc;
    return checkStatus(ctx, "zb_intersect_ray");
}

int zb_occluded(zb_scene *scene, int boundary, int n, const float *a, const float *b, uint8_t *occluded)
{
    ZB_QUERY_PROLOGUE("zb_occluded")
    ZB_BOUNDARY("zb_occluded")
    if (!a || !b || !occluded) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_occluded: NULL argument");
    IoPlan plan;
    size_t oA = plan.addIn(a, sizeof(float)*dim*n), oB = plan.addIn(b, sizeof(float)*dim*n);
    size_t oO = plan.addOut(occluded, (size_t)n);
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    int threads = 128, blocks = blocksFor(n, threads);
    if (dim == 3) occludedKernel<3><<<blocks, threads, 0, ctx->stream>>>(bvh, n, at<float>(ctx, oA, a), at<float>(ctx, oB, b), at<uint8_t>(ctx, oO, occluded), ctx->status.as<int>());
    else occludedKernel<2><<<blocks, threads, 0, ctx->stream>>>(bvh, n, at<float>(ctx, oA, a), at<float>(ctx, oB, b), at<uint8_t>(ctx, oO, occluded), ctx->status.as<int>());
    ZB_CUDA(cudaGetLastError());
    rc = stageOut(ctx, plan);
    if (rc != ZB_OK) return rc;
    return checkStatus(ctx, "zb_occluded");
}

int zb_count_ray_hits(zb_scene *scene, int boundary, int n, const float *origin, const float *dir, const float *tMax, int32_t *count)
{
    ZB_QUERY_PROLOGUE("zb_count_ray_hits")
    ZB_BOUNDARY("zb_count_ray_hits")
    if (!origin || !dir || !count) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_count_ray_hits: NULL argument");
    IoPlan plan;
    size_t oO = plan.addIn(origin, sizeof(float)*dim*n), oDir = plan.addIn(dir, sizeof(float)*dim*n), oT = plan.addIn(tMax, tMax ? sizeof(float)*n : 0);
    size_t oC = plan.addOut(count, sizeof(int32_t)*n);
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    int threads = 128, blocks = blocksFor(n, threads);
    if (dim == 3) countRayHitsKernel<3><<<blocks, threads, 0, ctx->stream>>>(bvh, n, at<float>(ctx, oO, origin), at<float>(ctx, oDir, dir), at<float>(ctx, oT, tMax), at<int>(ctx, oC, count), ctx->status.as<int>());
    else countRayHitsKernel<2><<<blocks, threads, 0, ctx->stream>>>(bvh, n, at<float>(ctx, oO, origin), at<float>(ctx, oDir, dir), at<float>(ctx, oT, tMax), at<int>(ctx, oC, count), ctx->status.as<int>());
    ZB_CUDA(cudaGetLastError());
    rc = stageOut(ctx, plan);
    if (rc != ZB_OK) return rc;
    return checkStatus(ctx, "zb_count_ray_hits");
}

int zb_inside_domain(zb_scene *scene, int n, const float *x, uint8_t *inside)
{
    ZB_QUERY_PROLOGUE("zb_inside_domain")
    if (!x || !inside) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_inside_domain: NULL argument");
    if (!scene->watertight) { // geometric_queries.h:358
        memset(inside, 1, (size_t)n);
        return ZB_OK;
    }
    IoPlan plan;
    size_t oX = plan.addIn(x, sizeof(float)*dim*n), oI = plan.addOut(inside, (size_t)n);
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    int threads = 128, blocks = blocksFor(n, threads);
    const int hasA = scene->boundary[ZB_ABSORBING].present ? 1 : 0, hasR = scene->boundary[ZB_REFLECTING].present ? 1 : 0;
    if (dim == 3) insideDomainKernel<3><<<blocks, threads, 0, ctx->stream>>>(scene->boundary[ZB_ABSORBING].view, scene->boundary[ZB_REFLECTING].view, hasA, hasR, n, at<float>(ctx, oX, x), at<uint8_t>(ctx, oI, inside), ctx->status.as<int>());
    else insideDomainKernel<2><<<blocks, threads, 0, ctx->stream>>>(scene->boundary[ZB_ABSORBING].view, scene->boundary[ZB_REFLECTING].view, hasA, hasR, n, at<float>(ctx, oX, x), at<uint8_t>(ctx, oI, inside), ctx->status.as<int>());
    ZB_CUDA(cudaGetLastError());
    rc = stageOut(ctx, plan);
    if (rc != ZB_OK) return rc;
    return checkStatus(ctx, "zb_inside_domain");
}

int zb_sample_sphere(zb_scene *scene, int boundary, int n, const float *x, const float *radius, const float *randNums,
                     int32_t *prim, float *pt, float *normal, float *pdf)
{
    ZB_QUERY_PROLOGUE("zb_sample_sphere")
    ZB_BOUNDARY("zb_sample_sphere")
    if (!x || !radius || !randNums || !prim || !pdf) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_sample_sphere: NULL argument");
    IoPlan plan;
    size_t oX = plan.addIn(x, sizeof(float)*dim*n), oR = plan.addIn(radius, sizeof(float)*n), oU = plan.addIn(randNums, sizeof(float)*dim*n);
    size_t oP = plan.addOut(prim, sizeof(int32_t)*n), oPt = plan.addOut(pt, pt ? sizeof(float)*dim*n : 0);
    size_t oN = plan.addOut(normal, normal ? sizeof(float)*dim*n : 0), oPdf = plan.addOut(pdf, sizeof(float)*n);
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    int threads = 128, blocks = blocksFor(n, threads);
    if (dim == 3) sampleSphereKernel<3><<<blocks, threads, 0, ctx->stream>>>(bvh, n, at<float>(ctx, oX, x), at<float>(ctx, oR, radius), at<float>(ctx, oU, randNums),
        at<int>(ctx, oP, prim), at<float>(ctx, oPt, pt), at<float>(ctx, oN, normal), at<float>(ctx, oPdf, pdf));
    else sampleSphereKernel<2><<<blocks, threads, 0, ctx->stream>>>(bvh, n, at<float>(ctx, oX, x), at<float>(ctx, oR, radius), at<float>(ctx, oU, randNums),
        at<int>(ctx, oP, prim), at<float>(ctx, oPt, pt), at<float>(ctx, oN, normal), at<float>(ctx, oPdf, pdf));
    ZB_CUDA(cudaGetLastError());
    return stageOut(ctx, plan);
}

int zb_star_radius(zb_scene *scene, int n, const float *x, const float *minRadius, const float *maxRadius,
                   float silhouettePrecision, const uint8_t *flipNormalOrientation, float *radius)
{
    ZB_QUERY_PROLOGUE("zb_star_radius")
    if (!x || !minRadius || !maxRadius || !radius) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_star_radius: NULL argument");
    IoPlan plan;
    size_t oX = plan.addIn(x, sizeof(float)*dim*n), oMn = plan.addIn(minRadius, sizeof(float)*n), oMx = plan.addIn(maxRadius, sizeof(float)*n);
    size_t oF = plan.addIn(flipNormalOrientation, flipNormalOrientation ? (size_t)n : 0);
    size_t oR = plan.addOut(radius, sizeof(float)*n);
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    SolveArgs a;
    fillSceneArgs(scene, a);
    int threads = 128, blocks = blocksFor(n, threads);
    ZB_CUDA(cudaFuncSetAttribute(starRadiusKernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
    ZB_CUDA(cudaFuncSetAttribute(starRadiusKernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
    if (dim == 3) starRadiusKernel<3><<<blocks, threads, kCoopSmem, ctx->stream>>>(a, n, at<float>(ctx, oX, x), at<float>(ctx, oMn, minRadius), at<float>(ctx, oMx, maxRadius),
        silhouettePrecision, at<uint8_t>(ctx, oF, flipNormalOrientation), at<float>(ctx, oR, radius));
    else starRadiusKernel<2><<<blocks, threads, kCoopSmem, ctx->stream>>>(a, n, at<float>(ctx, oX, x), at<float>(ctx, oMn, minRadius), at<float>(ctx, oMx, maxRadius),
        silhouettePrecision, at<uint8_t>(ctx, oF, flipNormalOrientation), at<float>(ctx, oR, radius));
    ZB_CUDA(cudaGetLastError());
    rc = stageOut(ctx, plan);
    if (rc != ZB_OK) return rc;
    return checkStatus(ctx, "zb_star_radius");
}

int zb_solve_dev(zb_scene *scene, int estimator, const zb_pde_desc *pde, const zb_walk_settings *settings,
                 const zb_sample_points *pts, zb_statistics *stats, int64_t *walkSteps_dev)
{
    if (!scene || !pde || !settings || !pts || !stats) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve_dev: NULL argument");
    if (estimator != ZB_WALK_ON_SPHERES && estimator != ZB_WALK_ON_STARS) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve_dev: unknown estimator");
    if (pts->n < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve_dev: negative point count");
    if (pts->n == 0) return ZB_OK;
    if (!pts->pt || !pts->normal || !pts->type || !pts->estimationQuantity || !pts->distToAbsorbingBoundary ||
        !pts->distToReflectingBoundary || !pts->nWalks || !pts->rngState || !pts->rngInc || !pts->firstSphereRadius)
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve_dev: a required sample-point array is NULL");
    if (pde->mode != ZB_PDE_MANUFACTURED && pde->mode != ZB_PDE_CONSTANT) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve_dev: unknown pde mode");
    // the closed-form families take their Laplacian over three coordinates: a 2D solve must not carry a z term
    if (scene->dim == 2 && pde->mode == ZB_PDE_MANUFACTURED &&
        ((pde->family == ZB_SOL_QUADRATIC && pde->params[2] != 0.0f) || pde->family == ZB_SOL_TRIG_3D))
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve_dev: this solution family has a z dependence and cannot describe a 2D problem");
    if (pde->kelvin && pde->absorptionCoeff > 0.0f) // kelvin_transform.h:285-288
        return fail(ZB_ERR_UNSUPPORTED, "zb_solve_dev: the Kelvin transform does not support non-zero absorption coefficients");
    if (settings->printLogs) return fail(ZB_ERR_UNSUPPORTED, "zb_solve_dev: printLogs is not supported on the device");
    zb_context *ctx = scene->ctx;
    ZB_CUDA(cudaSetDevice(ctx->device));

    SolveArgs a;
    fillSceneArgs(scene, a);
    a.estimator = estimator;
    a.ws = *settings;
    a.pde = *pde;
    if (pde->kelvin) { // flags of the wrapped PDE (kelvin_transform.h:340-347)
        a.pde.isSourceConstant = 0;
        a.pde.absorptionCoeff = 0.0f;
        if (scene->dim == 3) { a.pde.areRobinConditionsPureNeumann = 0; a.pde.areRobinCoeffsNonnegative = 0; }
    }
    a.pts = *pts;
    a.stats = *stats;
    a.nExtra = g_extra.n;
    a.extraPde = g_extra.pde;
    a.extraStats = g_extra.stats;
    a.walkSteps = ctx->counters.as<long long>() + 1;
    ZB_CUDA(cudaMemsetAsync(a.walkSteps, 0, sizeof(long long), ctx->stream));

    int rc = scene->dim == 3 ? solveRange<3>(scene, a, pts->n, g_gradientHint) : solveRange<2>(scene, a, pts->n, g_gradientHint);
    if (rc != ZB_OK) return rc;
    if (walkSteps_dev) {
        ZB_CUDA(cudaMemcpyAsync(walkSteps_dev, a.walkSteps, sizeof(long long), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return checkStatus(ctx, "zb_solve_dev");
}

int zb_solve(zb_scene *scene, int estimator, const zb_pde_desc *pde, const zb_walk_settings *settings,
             const zb_sample_points *pts, zb_statistics *stats, int64_t *walkSteps)
{
    if (!scene || !pts || !stats) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve: NULL argument");
    if (pts->n < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve: negative point count");
    if (walkSteps) *walkSteps = 0;
    if (pts->n == 0) return ZB_OK;
    if (!pts->pt || !pts->normal || !pts->type || !pts->estimationQuantity || !pts->distToAbsorbingBoundary ||
        !pts->distToReflectingBoundary || !pts->nWalks || !pts->rngState || !pts->rngInc || !pts->firstSphereRadius)
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve: a required sample-point array is NULL");
    if (!stats->solutionMean || !stats->solutionM2 || !stats->gradientMean || !stats->gradientM2 ||
        !stats->totalFirstSourceContribution || !stats->totalDerivativeContribution || !stats->nSolutionEstimates ||
        !stats->nGradientEstimates || !stats->totalWalkLength || !stats->totalSplits)
        return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve: a statistics array is NULL");
    zb_context *ctx = scene->ctx;
    ZB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)pts->n, dim = (size_t)scene->dim;

    IoPlan plan;
    zb_sample_points dp = *pts;
    zb_statistics ds = *stats;
    size_t o;
#define IN_OUT(field, hostPtr, bytes, type)                                                       \
    o = plan.addIn(hostPtr, bytes); plan.outputs.push_back({(void *)(hostPtr), bytes}); plan.outOffsets.push_back(o); field##Off = o;
#define IN_ONLY(hostPtr, bytes) plan.addIn(hostPtr, (hostPtr) ? (bytes) : 0)
    size_t ptOff = IN_ONLY(pts->pt, 4*dim*n), nOff = IN_ONLY(pts->normal, 4*dim*n),
           ddOff = IN_ONLY(pts->directionForDerivative, 4*dim*n), tyOff = IN_ONLY(pts->type, 4*n),
           eqOff = IN_ONLY(pts->estimationQuantity, 4*n), daOff = IN_ONLY(pts->distToAbsorbingBoundary, 4*n),
           drOff = IN_ONLY(pts->distToReflectingBoundary, 4*n), naOff = IN_ONLY(pts->estimateBoundaryNormalAligned, n),
           nwOff = IN_ONLY(pts->nWalks, 4*n), incOff = IN_ONLY(pts->rngInc, 8*n);
    size_t rngStateOff, fsrOff, meanOff, m2Off, gmeanOff, gm2Off, fsOff, dcOff, nsOff, ngOff, wlOff, spOff;
    IN_OUT(rngState, pts->rngState, 8*n, uint64_t)
    IN_OUT(fsr, pts->firstSphereRadius, 4*n, float)
    IN_OUT(mean, stats->solutionMean, 4*n, float)
    IN_OUT(m2, stats->solutionM2, 4*n, float)
    IN_OUT(gmean, stats->gradientMean, 4*dim*n, float)
    IN_OUT(gm2, stats->gradientM2, 4*dim*n, float)
    IN_OUT(fs, stats->totalFirstSourceContribution, 4*n, float)
    IN_OUT(dc, stats->totalDerivativeContribution, 4*n, float)
    IN_OUT(ns, stats->nSolutionEstimates, 4*n, int)
    IN_OUT(ng, stats->nGradientEstimates, 4*n, int)
    IN_OUT(wl, stats->totalWalkLength, 4*n, int)
    IN_OUT(sp, stats->totalSplits, 4*n, int)
#undef IN_OUT
#undef IN_ONLY
    size_t stepsOff = plan.total = (plan.total + 255) & ~size_t(255);
    plan.total += 8;
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    char *base = ctx->io.as<char>();
    dp.pt = (const float *)(base + ptOff);
    dp.normal = (const float *)(base + nOff);
    dp.directionForDerivative = pts->directionForDerivative ? (const float *)(base + ddOff) : nullptr;
    dp.type = (const int32_t *)(base + tyOff);
    dp.estimationQuantity = (const int32_t *)(base + eqOff);
    dp.distToAbsorbingBoundary = (const float *)(base + daOff);
    dp.distToReflectingBoundary = (const float *)(base + drOff);
    dp.estimateBoundaryNormalAligned = pts->estimateBoundaryNormalAligned ? (const uint8_t *)(base + naOff) : nullptr;
    dp.nWalks = (const int32_t *)(base + nwOff);
    dp.rngInc = (uint64_t *)(base + incOff);
    dp.rngState = (uint64_t *)(base + rngStateOff);
    dp.firstSphereRadius = (float *)(base + fsrOff);
    ds.solutionMean = (float *)(base + meanOff);
    ds.solutionM2 = (float *)(base + m2Off);
    ds.gradientMean = (float *)(base + gmeanOff);
    ds.gradientM2 = (float *)(base + gm2Off);
    ds.totalFirstSourceContribution = (float *)(base + fsOff);
    ds.totalDerivativeContribution = (float *)(base + dcOff);
    ds.nSolutionEstimates = (int32_t *)(base + nsOff);
    ds.nGradientEstimates = (int32_t *)(base + ngOff);
    ds.totalWalkLength = (int32_t *)(base + wlOff);
    ds.totalSplits = (int32_t *)(base + spOff);

    bool anyGradient = false;
    for (size_t i = 0; i < n && !anyGradient; i++) anyGradient = pts->estimationQuantity[i] == ZB_SOLUTION_AND_GRADIENT;
    g_gradientHint = anyGradient;
    rc = zb_solve_dev(scene, estimator, pde, settings, &dp, &ds, (int64_t *)(base + stepsOff));
    g_gradientHint = true;
    if (rc != ZB_OK) return rc;
    if (walkSteps) ZB_CUDA(cudaMemcpyAsync(walkSteps, base + stepsOff, 8, cudaMemcpyDeviceToHost, ctx->stream));
    return stageOut(ctx, plan);
}

// ---- vector-valued solves: C value channels on one set of walks -------------------------------------------------
namespace {
// descriptors and statistics structs of channels 1.. -> device; what zb_solve_dev does to a Kelvin-wrapped PDE's flags is done here
int uploadExtraChannels(zb_scene *scene, int nChannels, const zb_pde_desc *pdes, const zb_statistics *statsDev /* device pointers inside */,
                        ExtraChannels& out)
{
    zb_context *ctx = scene->ctx;
    const int nExtra = nChannels - 1;
    out = ExtraChannels();
    if (nExtra == 0) return ZB_OK;
    std::vector<zb_pde_desc> d(pdes + 1, pdes + nChannels);
    for (auto& e : d) {
        if (e.kelvin) {
            e.isSourceConstant = 0;
            e.absorptionCoeff = 0.0f;
            if (scene->dim == 3) { e.areRobinConditionsPureNeumann = 0; e.areRobinCoeffsNonnegative = 0; }
        }
    }
    const size_t descBytes = (nExtra*sizeof(zb_pde_desc) + 255) & ~size_t(255);
    ZB_CUDA(ctx->channelDesc.reserve(descBytes + nExtra*sizeof(zb_statistics) + 256));
    char *base = ctx->channelDesc.as<char>();
    // pageable sources: the copies are staged before the call returns, so the vectors may go out of scope
    ZB_CUDA(cudaMemcpyAsync(base, d.data(), nExtra*sizeof(zb_pde_desc), cudaMemcpyHostToDevice, ctx->stream));
    ZB_CUDA(cudaMemcpyAsync(base + descBytes, statsDev + 1, nExtra*sizeof(zb_statistics), cudaMemcpyHostToDevice, ctx->stream));
    out.n = nExtra;
    out.pde = reinterpret_cast<const zb_pde_desc *>(base);
    out.stats = reinterpret_cast<const zb_statistics *>(base + descBytes);
    return ZB_OK;
}

int checkChannels(const char *who, const zb_scene *scene, int nChannels, const zb_pde_desc *pdes, const zb_statistics *stats)
{
    if (!scene || !pdes || !stats) return fail(ZB_ERR_INVALID_ARGUMENT, std::string(who) + ": NULL argument");
    if (nChannels < 1 || nChannels > 256) return fail(ZB_ERR_INVALID_ARGUMENT, std::string(who) + ": nChannels must be 1..256");
    for (int c = 1; c < nChannels; c++) {
        // the walk is channel 0's: what steers it — screening, the Robin coefficient, the flags — is scalar in the reference
        // as well (PDE<T, DIM>::absorptionCoeff and robinCoeff are floats for every T, core/pde.h:26-44)
        const zb_pde_desc &a = pdes[0], &b = pdes[c];
        if (a.absorptionCoeff != b.absorptionCoeff || a.isSourceConstant != b.isSourceConstant ||
            a.areRobinConditionsPureNeumann != b.areRobinConditionsPureNeumann || a.areRobinCoeffsNonnegative != b.areRobinCoeffsNonnegative ||
            a.kelvin != b.kelvin || a.robinCoeff != b.robinCoeff || a.robinCoeffGrid.data != b.robinCoeffGrid.data)
            return fail(ZB_ERR_INVALID_ARGUMENT, std::string(who) + ": the channels must share absorptionCoeff, Robin coefficient, Kelvin wrapping and PDE flags");
        if (b.mode != ZB_PDE_MANUFACTURED && b.mode != ZB_PDE_CONSTANT) return fail(ZB_ERR_INVALID_ARGUMENT, std::string(who) + ": unknown pde mode");
        if (scene->dim == 2 && b.mode == ZB_PDE_MANUFACTURED && ((b.family == ZB_SOL_QUADRATIC && b.params[2] != 0.0f) || b.family == ZB_SOL_TRIG_3D))
            return fail(ZB_ERR_INVALID_ARGUMENT, std::string(who) + ": this solution family has a z dependence and cannot describe a 2D problem");
    }
    return ZB_OK;
}
} // namespace

int zb_solve_channels_dev(zb_scene *scene, int estimator, int nChannels, const zb_pde_desc *pdes, const zb_walk_settings *settings,
                          const zb_sample_points *pts, zb_statistics *stats, int64_t *walkSteps_dev)
{
    int rc = checkChannels("zb_solve_channels_dev", scene, nChannels, pdes, stats);
    if (rc != ZB_OK) return rc;
    ZB_CUDA(cudaSetDevice(scene->ctx->device));
    ExtraChannels extra;
    rc = uploadExtraChannels(scene, nChannels, pdes, stats, extra);
    if (rc != ZB_OK) return rc;
    g_extra = extra;
    rc = zb_solve_dev(scene, estimator, &pdes[0], settings, pts, &stats[0], walkSteps_dev);
    g_extra = ExtraChannels();
    return rc;
}

int zb_solve_channels(zb_scene *scene, int estimator, int nChannels, const zb_pde_desc *pdes, const zb_walk_settings *settings,
                      const zb_sample_points *pts, zb_statistics *stats, int64_t *walkSteps)
{
    int rc = checkChannels("zb_solve_channels", scene, nChannels, pdes, stats);
    if (rc != ZB_OK) return rc;
    if (!pts) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve_channels: NULL argument");
    if (nChannels == 1 || pts->n <= 0) return zb_solve(scene, estimator, &pdes[0], settings, pts, &stats[0], walkSteps);
    zb_context *ctx = scene->ctx;
    ZB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)pts->n, dim = (size_t)scene->dim;
    const int nExtra = nChannels - 1;
    // statistics of channels 1..: ten arrays each, staged in one device buffer (in/out like channel 0's)
    const size_t words = (8 + 2*dim)*n, stride = (words*4 + 255) & ~size_t(255);
    ZB_CUDA(ctx->channelIo.reserve(stride*nExtra + 256));
    std::vector<zb_statistics> dev(nChannels);
    struct Span { void *host; char *device; size_t bytes; };
    std::vector<Span> spans;
    for (int c = 1; c < nChannels; c++) {
        const zb_statistics& h = stats[c];
        if (!h.solutionMean || !h.solutionM2 || !h.gradientMean || !h.gradientM2 || !h.totalFirstSourceContribution ||
            !h.totalDerivativeContribution || !h.nSolutionEstimates || !h.nGradientEstimates || !h.totalWalkLength || !h.totalSplits)
            return fail(ZB_ERR_INVALID_ARGUMENT, "zb_solve_channels: a statistics array is NULL");
        char *p = ctx->channelIo.as<char>() + stride*(c - 1);
        auto put = [&](void *host, size_t count) { Span sp{host, p, 4*count}; spans.push_back(sp); p += 4*count; return sp.device; };
        zb_statistics& d = dev[c];
        d.solutionMean = (float *)put(h.solutionMean, n);
        d.solutionM2 = (float *)put(h.solutionM2, n);
        d.gradientMean = (float *)put(h.gradientMean, dim*n);
        d.gradientM2 = (float *)put(h.gradientM2, dim*n);
        d.totalFirstSourceContribution = (float *)put(h.totalFirstSourceContribution, n);
        d.totalDerivativeContribution = (float *)put(h.totalDerivativeContribution, n);
        d.nSolutionEstimates = (int32_t *)put(h.nSolutionEstimates, n);
        d.nGradientEstimates = (int32_t *)put(h.nGradientEstimates, n);
        d.totalWalkLength = (int32_t *)put(h.totalWalkLength, n);
        d.totalSplits = (int32_t *)put(h.totalSplits, n);
    }
    for (const Span& sp : spans) ZB_CUDA(cudaMemcpyAsync(sp.device, sp.host, sp.bytes, cudaMemcpyHostToDevice, ctx->stream));
    ExtraChannels extra;
    rc = uploadExtraChannels(scene, nChannels, pdes, dev.data(), extra);
    if (rc != ZB_OK) return rc;
    g_extra = extra;
    rc = zb_solve(scene, estimator, &pdes[0], settings, pts, &stats[0], walkSteps); // stages channel 0, runs the walks, folds every channel
    g_extra = ExtraChannels();
    if (rc != ZB_OK) return rc;
    for (const Span& sp : spans) ZB_CUDA(cudaMemcpyAsync(sp.host, sp.device, sp.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZB_OK;
}

int zb_rws_solve(zb_scene *scene, const zb_pde_desc *pde, const zb_walk_settings *settings, const zb_rws_samples *smp,
                 float normalOffsetForAbsorbingBoundary, float radiusClamp, float kernelRegularization,
                 zb_rws_eval_points *ev, int64_t *walkSteps, int64_t *splatRecords)
{
    if (!scene || !pde || !settings || !smp || !ev) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_rws_solve: NULL argument");
    if (smp->n < 0 || ev->n < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_rws_solve: negative count");
    if (walkSteps) *walkSteps = 0;
    if (splatRecords) *splatRecords = 0;
    if (smp->n == 0 || ev->n == 0) return ZB_OK;
    if (!smp->pt || !smp->normal || !smp->type || !smp->pdf || !smp->distToAbsorbingBoundary || !smp->distToReflectingBoundary ||
        !smp->rngState || !smp->rngInc) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_rws_solve: a sample array is NULL");
    if (!ev->pt || !ev->type || !ev->totals) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_rws_solve: an evaluation-point array is NULL");
    if (pde->kelvin) return fail(ZB_ERR_UNSUPPORTED, "zb_rws_solve: Kelvin-transformed PDEs are not supported");
    if (!(normalOffsetForAbsorbingBoundary > 0.0f)) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_rws_solve: normalOffsetForAbsorbingBoundary must be > 0");
    zb_context *ctx = scene->ctx;
    ZB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)smp->n, m = (size_t)ev->n, dim = (size_t)scene->dim;

    // samples whose contribution the settings switch off take no walk (reverse_walk_on_stars.h:95-124)
    std::vector<int32_t> quantity(n);
    for (size_t i = 0; i < n; i++) {
        int t = smp->type[i];
        bool off = t == ZB_IN_DOMAIN ? settings->ignoreSourceContribution != 0
                 : (t == ZB_ON_ABSORBING_BOUNDARY ? settings->ignoreAbsorbingBoundaryContribution != 0
                                                  : settings->ignoreReflectingBoundaryContribution != 0);
        quantity[i] = off ? ZB_SKIP : ZB_SOLUTION;
    }

    IoPlan plan;
    size_t ptOff = plan.addIn(smp->pt, 4*dim*n), nOff = plan.addIn(smp->normal, 4*dim*n), tyOff = plan.addIn(smp->type, 4*n),
           qOff = plan.addIn(quantity.data(), 4*n), alOff = plan.addIn(smp->estimateBoundaryNormalAligned, smp->estimateBoundaryNormalAligned ? n : 0),
           pdfOff = plan.addIn(smp->pdf, 4*n), daOff = plan.addIn(smp->distToAbsorbingBoundary, 4*n),
           drOff = plan.addIn(smp->distToReflectingBoundary, 4*n), incOff = plan.addIn(smp->rngInc, 8*n);
    size_t stOff = plan.addIn(smp->rngState, 8*n);
    plan.outputs.push_back({(void *)smp->rngState, 8*n}); plan.outOffsets.push_back(stOff);
    size_t ePtOff = plan.addIn(ev->pt, 4*dim*m), eNOff = plan.addIn(ev->normal, ev->normal ? 4*dim*m : 0), eTyOff = plan.addIn(ev->type, 4*m);
    size_t totOff = plan.addIn(ev->totals, 4*6*m);
    plan.outputs.push_back({(void *)ev->totals, 4*6*m}); plan.outOffsets.push_back(totOff);
    // scratch: nWalks (all 1), firstSphereRadius, contribution, record count
    auto scratch = [&](size_t bytes) { plan.total = (plan.total + 255) & ~size_t(255); size_t o = plan.total; plan.total += bytes; return o; };
    size_t nwOff = scratch(4*n), fsrOff = scratch(4*n), conOff = scratch(4*n), cntOff = scratch(8), stepsOff = scratch(8);

    long long capacity = std::max<long long>((long long)n*64, 1 << 18);
    for (int attempt = 0; attempt < 3; attempt++) {
        int rc = stageIn(ctx, plan);
        if (rc != ZB_OK) return rc;
        char *base = ctx->io.as<char>();
        ZB_CUDA(ctx->rwsRecords.reserve((size_t)capacity*48 + 256));
        ZB_CUDA(cudaMemsetAsync(base + cntOff, 0, 8, ctx->stream));
        ZB_CUDA(cudaMemsetAsync(base + stepsOff, 0, 8, ctx->stream));
        ZB_CUDA(cudaMemsetAsync(base + fsrOff, 0, 4*n, ctx->stream));
        std::vector<int32_t> ones(n, 1);
        ZB_CUDA(cudaMemcpyAsync(base + nwOff, ones.data(), 4*n, cudaMemcpyHostToDevice, ctx->stream));

        SolveArgs a;
        fillSceneArgs(scene, a);
        a.estimator = ZB_WALK_ON_STARS;
        a.ws = *settings;
        a.pde = *pde;
        a.pts.n = (int32_t)n;
        a.pts.pt = (const float *)(base + ptOff);
        a.pts.normal = (const float *)(base + nOff);
        a.pts.directionForDerivative = nullptr;
        a.pts.type = (const int32_t *)(base + tyOff);
        a.pts.estimationQuantity = (const int32_t *)(base + qOff);
        a.pts.distToAbsorbingBoundary = (const float *)(base + daOff);
        a.pts.distToReflectingBoundary = (const float *)(base + drOff);
        a.pts.estimateBoundaryNormalAligned = smp->estimateBoundaryNormalAligned ? (const uint8_t *)(base + alOff) : nullptr;
        a.pts.nWalks = (const int32_t *)(base + nwOff);
        a.pts.rngState = (uint64_t *)(base + stOff);
        a.pts.rngInc = (uint64_t *)(base + incOff);
        a.pts.firstSphereRadius = (float *)(base + fsrOff);
        a.walkSteps = (long long *)(base + stepsOff);

        // 1. what every sample carries (needs the caller's ignore flags)
        const int threads = 128;
        if (scene->dim == 3) {
            ZB_CUDA(cudaFuncSetAttribute(rwsContributionKernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
            rwsContributionKernel<3><<<blocksFor((int)n, threads), threads, kCoopSmem, ctx->stream>>>(a, (float *)(base + conOff));
        } else {
            ZB_CUDA(cudaFuncSetAttribute(rwsContributionKernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
            rwsContributionKernel<2><<<blocksFor((int)n, threads), threads, kCoopSmem, ctx->stream>>>(a, (float *)(base + conOff));
        }
        ZB_CUDA(cudaGetLastError());
        ctx->launches += 1;

        // 2. the walks: forward WoSt steps without boundary / source terms, one splat record per step
        a.ws.ignoreReflectingBoundaryContribution = 1;
        a.ws.ignoreSourceContribution = 1;
        a.rwsRecords = ctx->rwsRecords.as<float4>();
        a.rwsCount = (unsigned long long *)(base + cntOff);
        a.rwsCapacity = capacity;
        a.rwsPdf = (const float *)(base + pdfOff);
        a.rwsContribution = (const float *)(base + conOff);
        a.rwsNormalOffset = normalOffsetForAbsorbingBoundary;
        rc = scene->dim == 3 ? solveRange<3>(scene, a, (int)n, false) : solveRange<2>(scene, a, (int)n, false);
        if (rc != ZB_OK) return rc;
        unsigned long long count = 0;
        long long steps = 0;
        ZB_CUDA(cudaMemcpyAsync(&count, base + cntOff, 8, cudaMemcpyDeviceToHost, ctx->stream));
        ZB_CUDA(cudaMemcpyAsync(&steps, base + stepsOff, 8, cudaMemcpyDeviceToHost, ctx->stream));
        ZB_CUDA(cudaStreamSynchronize(ctx->stream));
        rc = checkStatus(ctx, "zb_rws_solve");
        if (rc != ZB_OK) return rc;
        if ((long long)count > capacity) { // the record buffer was too small: run the same walks again with room for all of them
            capacity = (long long)count + (long long)count/8 + 1024;
            continue;
        }

        // 3. gather
        RwsGatherArgs g;
        std::memset(&g, 0, sizeof(g));
        g.reflecting = scene->boundary[ZB_REFLECTING].view;
        g.hasReflecting = scene->boundary[ZB_REFLECTING].present ? 1 : 0;
        g.records = ctx->rwsRecords.as<float4>();
        g.nRecords = (long long)count;
        g.nEval = (int)m;
        g.evalPt = (const float *)(base + ePtOff);
        g.evalNormal = ev->normal ? (const float *)(base + eNOff) : nullptr;
        g.evalType = (const int *)(base + eTyOff);
        g.totals = (float *)(base + totOff);
        g.sqrtLambda = pde->absorptionCoeff > 0.0f ? std::sqrt(pde->absorptionCoeff) : 0.0f;
        g.radiusClamp = radiusClamp;
        g.kernelRegularization = kernelRegularization;
        // reverse_walk_solver.h:244-245
        bool selfNorm = scene->watertight != 0 && pde->absorptionCoeff == 0.0f;
        if (g.hasReflecting && selfNorm) selfNorm = pde->areRobinConditionsPureNeumann != 0;
        g.useSelfNormalization = selfNorm ? 1 : 0;
        g.status = ctx->status.as<int>();

        // spatial culling (rws.cuh): records sorted into tiles with bounding boxes, evaluation points sorted by Morton code
        RwsCullArgs c;
        std::memset(&c, 0, sizeof(c));
        c.records = ctx->rwsRecords.as<float4>();
        c.nRecords = (long long)count;
        c.nTiles = (int)((count + kRwsTile - 1)/kRwsTile);
        c.nSuper = (c.nTiles + kRwsSuper - 1)/kRwsSuper;
        c.nEval = (int)m;
        c.evalPt = g.evalPt;
        float lo[3] = {kMaxFloat, kMaxFloat, kMaxFloat}, hi[3] = {-kMaxFloat, -kMaxFloat, -kMaxFloat};
        for (int b = 0; b < 2; b++) {
            if (!scene->boundary[b].present) continue;
            for (int k = 0; k < 3; k++) {
                lo[k] = std::min(lo[k], scene->boundary[b].host.boxMin[k]);
                hi[k] = std::max(hi[k], scene->boundary[b].host.boxMax[k]);
            }
        }
        float diag2 = 0.0f;
        for (int k = 0; k < 3; k++) {
            float l = hi[k] > lo[k] ? lo[k] : 0.0f, h = hi[k] > lo[k] ? hi[k] : 1.0f;
            c.lo[k] = l;
            c.scale[k] = (scene->dim == 3 ? 1024.0f : 32768.0f)/(h - l);
            if (k < scene->dim) diag2 += (h - l)*(h - l);
        }
        c.invDiag = 1.0f/std::sqrt(std::max(diag2, 1e-30f));
        size_t sortBytesRec = 0, sortBytesEval = 0;
        unsigned long long *k64 = nullptr;
        uint32_t *k32 = nullptr;
        ZB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sortBytesRec, k64, k64, k32, k32, (int)std::max<unsigned long long>(count, 1), 0, 64, ctx->stream));
        ZB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sortBytesEval, k32, k32, k32, k32, (int)m, 0, 32, ctx->stream));
        const size_t nr = (size_t)std::max<unsigned long long>(count, 1);
        Carver sizing(nullptr);
        sizing.take<unsigned long long>(2*nr); sizing.take<uint32_t>(2*nr); sizing.take<uint32_t>(4*m);
        sizing.take<float4>(2*(size_t)(c.nTiles + c.nSuper) + 2); sizing.take<char>(std::max(sortBytesRec, sortBytesEval) + 256);
        ZB_CUDA(ctx->rwsScratch.reserve(sizing.offset + 16*256)); // + alignment padding of the individual arrays carved below
        ZB_CUDA(ctx->rwsSorted.reserve(nr*48 + 256));
        Carver cv(ctx->rwsScratch.ptr);
        unsigned long long *keysA = cv.take<unsigned long long>(nr), *keysB = cv.take<unsigned long long>(nr);
        uint32_t *idxA = cv.take<uint32_t>(nr), *idxB = cv.take<uint32_t>(nr);
        uint32_t *eKeysA = cv.take<uint32_t>(m), *eKeysB = cv.take<uint32_t>(m), *eIdxA = cv.take<uint32_t>(m), *eIdxB = cv.take<uint32_t>(m);
        c.tileLo = cv.take<float4>((size_t)(c.nTiles + c.nSuper) + 1);
        c.tileHi = cv.take<float4>((size_t)(c.nTiles + c.nSuper) + 1);
        char *sortTemp = cv.take<char>(std::max(sortBytesRec, sortBytesEval) + 256);
        c.sorted = ctx->rwsSorted.as<float4>();
        c.keys = keysA; c.index = idxA;
        c.evalKeys = eKeysA; c.evalIndex = eIdxA;
        if (count > 0) {
            if (scene->dim == 3) rwsRecordKeysKernel<3><<<blocksFor((int)count, 256), 256, 0, ctx->stream>>>(c);
            else rwsRecordKeysKernel<2><<<blocksFor((int)count, 256), 256, 0, ctx->stream>>>(c);
            size_t bytes = sortBytesRec;
            ZB_CUDA(cub::DeviceRadixSort::SortPairs(sortTemp, bytes, keysA, keysB, idxA, idxB, (int)count, 0, 64, ctx->stream));
            c.index = idxB;
            rwsReorderKernel<<<blocksFor((int)count, 256), 256, 0, ctx->stream>>>(c);
            rwsTileBoundsKernel<<<blocksFor(c.nTiles*32, 256), 256, 0, ctx->stream>>>(c);
            rwsSuperBoundsKernel<<<blocksFor(c.nSuper*32, 256), 256, 0, ctx->stream>>>(c);
        }
        if (scene->dim == 3) rwsEvalKeysKernel<3><<<blocksFor((int)m, 256), 256, 0, ctx->stream>>>(c);
        else rwsEvalKeysKernel<2><<<blocksFor((int)m, 256), 256, 0, ctx->stream>>>(c);
        {
            size_t bytes = sortBytesEval;
            ZB_CUDA(cub::DeviceRadixSort::SortPairs(sortTemp, bytes, eKeysA, eKeysB, eIdxA, eIdxB, (int)m, 0, 32, ctx->stream));
        }
        g.records = c.sorted;
        if (scene->dim == 3) rwsGatherKernel<3><<<blocksFor((int)m, kRwsThreads), kRwsThreads, 0, ctx->stream>>>(g, eIdxB, c.tileLo, c.tileHi, c.nTiles, c.nSuper);
        else rwsGatherKernel<2><<<blocksFor((int)m, kRwsThreads), kRwsThreads, 0, ctx->stream>>>(g, eIdxB, c.tileLo, c.tileHi, c.nTiles, c.nSuper);
        ZB_CUDA(cudaGetLastError());
        ctx->launches += 8;
        rc = stageOut(ctx, plan);
        if (rc != ZB_OK) return rc;
        if (walkSteps) *walkSteps = steps;
        if (splatRecords) *splatRecords = (int64_t)count;
        return checkStatus(ctx, "zb_rws_solve");
    }
    return fail(ZB_ERR_OUT_OF_MEMORY, "zb_rws_solve: the splat-record buffer kept overflowing");
}

#ifdef ZB_STATS
// debug build only: reads and clears the traversal counters (see queries.cuh)
int zb_debug_counters(zb_context *ctx, unsigned long long *out16)
{
    ZB_CUDA(cudaSetDevice(ctx->device));
    ZB_CUDA(cudaStreamSynchronize(ctx->stream));
    ZB_CUDA(cudaMemcpyFromSymbol(out16, g_zbCounters, sizeof(unsigned long long)*16));
    unsigned long long zero[16] = {0};
    ZB_CUDA(cudaMemcpyToSymbol(g_zbCounters, zero, sizeof(zero)));
    return ZB_OK;
}
#endif

int zb_pcg32_kat(zb_context *ctx, uint64_t initstate, uint64_t initseq, int n, uint32_t *out)
{
    if (!ctx || !out || n < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_pcg32_kat: invalid argument");
    if (n == 0) return ZB_OK;
    ZB_CUDA(cudaSetDevice(ctx->device));
    ZB_CUDA(ctx->io.reserve(sizeof(uint32_t)*(size_t)n));
    pcg32Kernel<<<1, 32, 0, ctx->stream>>>(initstate, initseq, n, ctx->io.as<uint32_t>());
    ZB_CUDA(cudaGetLastError());
    ZB_CUDA(cudaMemcpyAsync(out, ctx->io.ptr, sizeof(uint32_t)*(size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    ZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZB_OK;
}

int zb_pcg32_demo(zb_context *ctx, uint64_t initstate, uint64_t initseq, int rounds,
                  uint32_t *u32, uint8_t *coins, uint8_t *rolls, uint8_t *cards)
{
    if (!ctx || !u32 || !coins || !rolls || !cards || rounds < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_pcg32_demo: invalid argument");
    if (rounds == 0) return ZB_OK;
    ZB_CUDA(cudaSetDevice(ctx->device));
    IoPlan plan;
    size_t oU = plan.addOut(u32, sizeof(uint32_t)*6*(size_t)rounds), oC = plan.addOut(coins, 65*(size_t)rounds),
           oR = plan.addOut(rolls, 33*(size_t)rounds), oK = plan.addOut(cards, 52*(size_t)rounds);
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    pcg32DemoKernel<<<1, 32, 0, ctx->stream>>>(initstate, initseq, rounds, at<uint32_t>(ctx, oU, u32), at<uint8_t>(ctx, oC, coins),
                                               at<uint8_t>(ctx, oR, rolls), at<uint8_t>(ctx, oK, cards));
    ZB_CUDA(cudaGetLastError());
    return stageOut(ctx, plan);
}

int zb_pcg32_floats(zb_context *ctx, uint64_t initstate, uint64_t initseq, int n, float *out)
{
    if (!ctx || !out || n < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_pcg32_floats: invalid argument");
    if (n == 0) return ZB_OK;
    ZB_CUDA(cudaSetDevice(ctx->device));
    IoPlan plan;
    size_t o = plan.addOut(out, sizeof(float)*(size_t)n);
    int rc = stageIn(ctx, plan);
    if (rc != ZB_OK) return rc;
    pcg32FloatKernel<<<1, 32, 0, ctx->stream>>>(initstate, initseq, n, at<float>(ctx, o, out));
    ZB_CUDA(cudaGetLastError());
    return stageOut(ctx, plan);
}

} // extern "C"
