// C ABI of the Boundary Value Caching splat (include/zombie_b200.h: zb_bvc_splat, zb_bvc_splat_dev).
#include <cstring>
#include <vector>

#include "capi_common.h"
#include "bvc.cuh"

using namespace zb;

namespace {

// bucket of a cache sample: which of the evaluation point's five statistics it feeds
// (boundary_value_caching.h:746-768, 808-810)
int bucketOf(int type, bool aligned)
{
    if (type == ZB_ON_ABSORBING_BOUNDARY) return aligned ? 1 : 0;
    if (type == ZB_ON_REFLECTING_BOUNDARY) return aligned ? 3 : 2;
    return 4;
}

// packs the SoA samples into 48-byte records, stably sorted by bucket
void packSamples(int dim, const zb_bvc_samples *s, std::vector<BvcSample>& out, int offsets[kBvcBuckets + 1])
{
    int counts[kBvcBuckets] = {0, 0, 0, 0, 0};
    std::vector<int> bucket(s->n);
    for (int i = 0; i < s->n; i++) {
        bool aligned = s->estimateBoundaryNormalAligned && s->estimateBoundaryNormalAligned[i];
        bucket[i] = bucketOf(s->type[i], aligned);
        counts[bucket[i]]++;
    }
    offsets[0] = 0;
    for (int b = 0; b < kBvcBuckets; b++) offsets[b + 1] = offsets[b] + counts[b];
    int cursor[kBvcBuckets];
    for (int b = 0; b < kBvcBuckets; b++) cursor[b] = offsets[b];
    out.resize((size_t)s->n);
    for (int i = 0; i < s->n; i++) {
        BvcSample r;
        std::memset(&r, 0, sizeof(r));
        bool aligned = s->estimateBoundaryNormalAligned && s->estimateBoundaryNormalAligned[i];
        float sign = aligned ? -1.0f : 1.0f; // :694
        const float *p = s->pt + (size_t)dim*i, *n = s->normal + (size_t)dim*i;
        r.px = p[0]; r.py = p[1]; r.pz = dim == 3 ? p[2] : 0.0f;
        r.nx = n[0]*sign; r.ny = n[1]*sign; r.nz = dim == 3 ? n[2]*sign : 0.0f;
        r.invPdf = 1.0f/s->pdf[i];
        r.solution = s->solution[i];
        r.normalDerivative = s->normalDerivative[i];
        r.robin = s->contribution[i];
        r.robinCoeff = s->robinCoeff[i];
        out[(size_t)cursor[bucket[i]]++] = r;
    }
}

int launchSplat(zb_context *ctx, int dim, bool yukawa, const BvcArgs& a)
{
    int blocks = (a.nEval + kBvcThreads - 1)/kBvcThreads;
    ZB_CUDA(cudaEventRecord(ctx->evSplat[0], ctx->stream));
    if (dim == 3) {
        if (yukawa) bvcSplatKernel<3, true><<<blocks, kBvcThreads, 0, ctx->stream>>>(a);
        else bvcSplatKernel<3, false><<<blocks, kBvcThreads, 0, ctx->stream>>>(a);
    } else {
        if (yukawa) bvcSplatKernel<2, true><<<blocks, kBvcThreads, 0, ctx->stream>>>(a);
        else bvcSplatKernel<2, false><<<blocks, kBvcThreads, 0, ctx->stream>>>(a);
    }
    ZB_CUDA(cudaGetLastError());
    ZB_CUDA(cudaEventRecord(ctx->evSplat[1], ctx->stream));
    ctx->splatProfiled = true;
    ctx->launches += 1;
    return ZB_OK;
}

int validate(const char *name, zb_context *ctx, int dim, float absorptionCoeff, const zb_bvc_samples *s, const zb_bvc_eval_points *e)
{
    if (!ctx || !s || !e) return fail(ZB_ERR_INVALID_ARGUMENT, std::string(name) + ": NULL argument");
    if (dim != 2 && dim != 3) return fail(ZB_ERR_INVALID_ARGUMENT, std::string(name) + ": dim must be 2 or 3");
    if (s->n < 0 || e->n < 0) return fail(ZB_ERR_INVALID_ARGUMENT, std::string(name) + ": negative count");
    if (absorptionCoeff < 0.0f) return fail(ZB_ERR_INVALID_ARGUMENT, std::string(name) + ": absorptionCoeff must be >= 0");
    if (s->n > 0 && (!s->pt || !s->normal || !s->type || !s->pdf || !s->solution || !s->normalDerivative || !s->contribution || !s->robinCoeff))
        return fail(ZB_ERR_INVALID_ARGUMENT, std::string(name) + ": a sample array is NULL");
    if (e->n > 0 && (!e->pt || !e->type || !e->distToAbsorbingBoundary || !e->distToReflectingBoundary || !e->bucketCount ||
                     !e->bucketSum || !e->solution || !e->gradient))
        return fail(ZB_ERR_INVALID_ARGUMENT, std::string(name) + ": an evaluation-point array is NULL");
    return ZB_OK;
}

} // namespace

extern "C" {

int zb_bvc_splat_dev(zb_context *ctx, int dim, float absorptionCoeff, const zb_bvc_samples *samples,
                     float radiusClamp, float kernelRegularization, float robinCoeffCutoffForNormalDerivative,
                     float cutoffDistToAbsorbingBoundary, float cutoffDistToReflectingBoundary,
                     zb_bvc_eval_points *e, int64_t *pairs)
{
    int rc = validate("zb_bvc_splat_dev", ctx, dim, absorptionCoeff, samples, e);
    if (rc != ZB_OK) return rc;
    if (pairs) *pairs = 0;
    if (e->n == 0) return ZB_OK;
    ZB_CUDA(cudaSetDevice(ctx->device));

    std::vector<BvcSample> packed;
    BvcArgs a;
    std::memset(&a, 0, sizeof(a));
    packSamples(dim, samples, packed, a.bucketOffset);
    ZB_CUDA(ctx->bvcSamples.reserve(packed.size()*sizeof(BvcSample) + 256));
    if (!packed.empty()) {
        ZB_CUDA(cudaMemcpyAsync(ctx->bvcSamples.ptr, packed.data(), packed.size()*sizeof(BvcSample), cudaMemcpyHostToDevice, ctx->stream));
        // `packed` is pageable host memory: the copy has been staged by the driver when the call returns
    }
    a.samples = ctx->bvcSamples.as<BvcSample>();
    a.nEval = e->n;
    a.evalPt = e->pt;
    a.evalType = e->type;
    a.distA = e->distToAbsorbingBoundary;
    a.distR = e->distToReflectingBoundary;
    a.bucketCount = e->bucketCount;
    a.bucketSum = e->bucketSum;
    a.solution = e->solution;
    a.gradient = e->gradient;
    a.absorptionCoeff = absorptionCoeff;
    a.radiusClamp = radiusClamp;
    a.kernelRegularization = kernelRegularization;
    a.robinCutoff = robinCoeffCutoffForNormalDerivative;
    a.cutoffA = cutoffDistToAbsorbingBoundary;
    a.cutoffR = cutoffDistToReflectingBoundary;
    rc = launchSplat(ctx, dim, absorptionCoeff > 0.0f, a);
    if (rc != ZB_OK) return rc;
    if (pairs) *pairs = (int64_t)samples->n*(int64_t)e->n;
    return ZB_OK;
}

int zb_bvc_splat(zb_context *ctx, int dim, float absorptionCoeff, const zb_bvc_samples *samples,
                 float radiusClamp, float kernelRegularization, float robinCoeffCutoffForNormalDerivative,
                 float cutoffDistToAbsorbingBoundary, float cutoffDistToReflectingBoundary,
                 zb_bvc_eval_points *e, int64_t *pairs)
{
    int rc = validate("zb_bvc_splat", ctx, dim, absorptionCoeff, samples, e);
    if (rc != ZB_OK) return rc;
    if (pairs) *pairs = 0;
    if (e->n == 0) return ZB_OK;
    ZB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)e->n, d = (size_t)dim;

    // device staging: [pt | type | distA | distR | bucketCount | bucketSum | solution | gradient], 256-byte aligned
    size_t sizes[8] = {4*d*n, 4*n, 4*n, 4*n, 4*kBvcBuckets*n, 4*kBvcBuckets*(1 + d)*n, 4*n, 4*d*n};
    size_t off[8], total = 0;
    for (int k = 0; k < 8; k++) { total = (total + 255) & ~size_t(255); off[k] = total; total += sizes[k]; }
    ZB_CUDA(ctx->bvcIo.reserve(total + 256));
    char *base = ctx->bvcIo.as<char>();
    const void *in[6] = {e->pt, e->type, e->distToAbsorbingBoundary, e->distToReflectingBoundary, e->bucketCount, e->bucketSum};
    for (int k = 0; k < 6; k++) ZB_CUDA(cudaMemcpyAsync(base + off[k], in[k], sizes[k], cudaMemcpyHostToDevice, ctx->stream));

    zb_bvc_eval_points dev = *e;
    dev.pt = (const float *)(base + off[0]);
    dev.type = (const int32_t *)(base + off[1]);
    dev.distToAbsorbingBoundary = (const float *)(base + off[2]);
    dev.distToReflectingBoundary = (const float *)(base + off[3]);
    dev.bucketCount = (int32_t *)(base + off[4]);
    dev.bucketSum = (float *)(base + off[5]);
    dev.solution = (float *)(base + off[6]);
    dev.gradient = (float *)(base + off[7]);
    rc = zb_bvc_splat_dev(ctx, dim, absorptionCoeff, samples, radiusClamp, kernelRegularization, robinCoeffCutoffForNormalDerivative,
                          cutoffDistToAbsorbingBoundary, cutoffDistToReflectingBoundary, &dev, pairs);
    if (rc != ZB_OK) return rc;
    void *out[4] = {e->bucketCount, e->bucketSum, e->solution, e->gradient};
    for (int k = 0; k < 4; k++) ZB_CUDA(cudaMemcpyAsync(out[k], base + off[4 + k], sizes[4 + k], cudaMemcpyDeviceToHost, ctx->stream));
    ZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZB_OK;
}

int zb_context_last_splat_ms(zb_context *ctx, float *ms)
{
    if (!ctx || !ms) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_context_last_splat_ms: NULL argument");
    if (!ctx->splatProfiled) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_context_last_splat_ms: no splat has run on this context");
    ZB_CUDA(cudaSetDevice(ctx->device));
    ZB_CUDA(cudaEventSynchronize(ctx->evSplat[1]));
    ZB_CUDA(cudaEventElapsedTime(ms, ctx->evSplat[0], ctx->evSplat[1]));
    return ZB_OK;
}

} // extern "C"
