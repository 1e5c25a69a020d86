// Boundary Value Caching splat (BASELINE config 4).
//
// Reference: BoundaryValueCaching::splat (zombie/variance_reduction/boundary_value_caching.h:371-502) loops
// SERIALLY over cache samples and, per sample, runs a TBB parallel_for over evaluation points; each (sample,
// evaluation point) pair heap-allocates a free-space Green's function (:385-391), evaluates G, P, grad G,
// grad P (zombie/core/distributions.h:124-321) and folds one estimate into one of five Welford
// SampleStatistics of the evaluation point (splatBoundaryData :682-769, splatSourceData :772-811).
//
// Here the loop nest is turned inside out into an atomic-free gather: one thread owns one evaluation point and
// keeps the running sums of one statistics bucket in registers; the cache samples, sorted by bucket on the
// host, stream through shared memory in tiles that the whole CTA reads by broadcast. All threads of a CTA look
// at the same sample at the same time, so the per-sample case analysis (source / Robin with or without a
// normal-derivative estimate / Neumann-Dirichlet) is branch-uniform. A Welford mean over N estimates equals
// sum/N, so (count, sums) per bucket carry the same information as the reference's statistics and merge by
// addition across calls and across GPUs.
//
// FP32-compute-bound: ~100 flop per pair against 48 B of sample data amortised over the whole CTA.
#pragma once

#include "greens.cuh"

namespace zb {

constexpr int kBvcBuckets = 5; // absorbing, absorbing normal-aligned, reflecting, reflecting normal-aligned, source
constexpr int kBvcTile = 256;  // samples per shared-memory tile (12 KB)
constexpr int kBvcThreads = 256;

// one cache sample, 3 x float4; the normal already carries the -1 of estimateBoundaryNormalAligned (:694)
struct BvcSample {
    float px, py, pz, nx;
    float ny, nz, invPdf, solution;
    float normalDerivative, robin, robinCoeff, pad;
};
static_assert(sizeof(BvcSample) == 48, "BvcSample must be 48 bytes");

struct BvcArgs {
    const BvcSample *samples;      // sorted by bucket
    int bucketOffset[kBvcBuckets + 1];
    int nEval;
    const float *evalPt;           // nEval*DIM
    const int *evalType;
    const float *distA, *distR;
    int *bucketCount;              // nEval*5, in/out
    float *bucketSum;              // nEval*5*(1+DIM), in/out
    float *solution;               // nEval
    float *gradient;               // nEval*DIM
    float absorptionCoeff, radiusClamp, kernelRegularization, robinCutoff, cutoffA, cutoffR;
};

// G, P, grad G, grad P of the free-space Green's function with pole x evaluated at the sample y with normal n
// (Harmonic/YukawaGreensFnFreeSpace<2|3>, distributions.h:124-321); r is already clamped.
template <int DIM, bool YUKAWA>
__device__ __forceinline__ void freeSpaceKernels(const Vec<DIM>& xy, float r, const Vec<DIM>& n, float sqrtLambda,
                                                 float& G, float& P, Vec<DIM>& dG, Vec<DIM>& dP)
{
    float ndotxy = dot(n, xy);
    if (DIM == 3) {
        const float inv4pi = 1.0f/(4.0f*kPiF);
        float r2 = r*r;
        if (!YUKAWA) {
            float invR = 1.0f/r;
            float invR3 = invR*invR*invR*inv4pi; // 1/(4 pi r^3)
            G = inv4pi*invR;
            P = ndotxy*invR3;
            float s = 3.0f*(ndotxy/r2);
            for (int k = 0; k < DIM; k++) {
                dG[k] = -xy[k]*invR3;
                dP[k] = (n[k] - s*xy[k])*invR3;
            }
        } else {
            float mur = r*sqrtLambda;
            float expmur = expf(-mur);
            float Qr1 = sqrtLambda*expmur*(1.0f + 1.0f/mur);
            float Qr2 = sqrtLambda*expmur/mur;
            float inv4pir2 = inv4pi/r2;
            G = expmur*inv4pi/r;
            P = ndotxy*Qr1*inv4pir2;
            float s = (ndotxy/r2)*(2.0f*Qr1 + Qr2);
            for (int k = 0; k < DIM; k++) {
                dG[k] = -xy[k]*Qr1*inv4pir2;
                dP[k] = (n[k]*Qr1 - s*xy[k])*inv4pir2;
            }
        }
    } else if (!YUKAWA) {
        const float inv2pi = 1.0f/(2.0f*kPiF);
        float r2 = r*r;
        float inv2pir2 = inv2pi/r2;
        G = -logf(r)*inv2pi;
        P = ndotxy*inv2pir2;
        float s = 2.0f*(ndotxy/r2);
        for (int k = 0; k < DIM; k++) {
            dG[k] = -xy[k]*inv2pir2;
            dP[k] = (n[k] - s*xy[k])*inv2pir2;
        }
    } else {
        // YukawaGreensFnFreeSpace<2> (distributions.h:215-262): modified Bessel functions K0, K1, K2 evaluated in double like
        // zombie/core/bessel.h, K2 by the recurrence bessk(2, x) uses (bessel.h:345-360)
        const float inv2pi = 1.0f/(2.0f*kPiF);
        float r2 = r*r;
        float mur = r*sqrtLambda;
        float K0 = (float)bessel::k0((double)mur), K1 = (float)bessel::k1((double)mur);
        float K2 = (float)(bessel::k0((double)mur) + (2.0/(double)mur)*bessel::k1((double)mur));
        float lambda = sqrtLambda*sqrtLambda;
        float Qr1 = sqrtLambda*K1, Qr2 = lambda*(K0 + K2)/2.0f;
        float inv2pir = inv2pi/r;
        G = K0*inv2pi;
        P = ndotxy*Qr1*inv2pir;
        float s = (ndotxy/r2)*(Qr1 + r*Qr2);
        for (int k = 0; k < DIM; k++) {
            dG[k] = -xy[k]*Qr1*inv2pir;
            dP[k] = (n[k]*Qr1 - s*xy[k])*inv2pir;
        }
    }
}

// KernelRegularization<DIM> (distributions.h:12-55)
template <int DIM>
__device__ __forceinline__ void regularize(float r, float& G, float& P)
{
    if (DIM == 3) {
        float e = erff(r);
        G *= e;
        P *= e - 2.0f*r*expf(-r*r)/1.7724538509055160273f;
    } else {
        P *= 1.0f - expf(-r*r);
    }
}

template <int DIM, bool YUKAWA>
__global__ void __launch_bounds__(kBvcThreads) bvcSplatKernel(BvcArgs a)
{
    __shared__ float4 tile[kBvcTile*3];
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const bool live = i < a.nEval;
    Vec<DIM> x = vzero<DIM>();
    bool evaluate = false;
    float alpha = 1.0f, alphaGrad = 1.0f;
    if (live) {
        x = loadVec<DIM>(a.evalPt + (size_t)DIM*i);
        // boundary_value_caching.h:381-383: no estimate closer to the boundary than the cutoffs
        evaluate = !(a.distA[i] < a.cutoffA || a.distR[i] < a.cutoffR);
        int type = a.evalType[i];
        if (type == ZB_ON_ABSORBING_BOUNDARY || type == ZB_ON_REFLECTING_BOUNDARY) { alpha = 2.0f; alphaGrad = 0.0f; }
    }
    const float sqrtLambda = YUKAWA ? sqrtf(a.absorptionCoeff) : 0.0f;
    const float invReg = a.kernelRegularization > 0.0f ? 1.0f/a.kernelRegularization : 0.0f;

    float totalU = 0.0f;
    Vec<DIM> totalG = vzero<DIM>();
    for (int b = 0; b < kBvcBuckets; b++) {
        const int begin = a.bucketOffset[b], end = a.bucketOffset[b + 1];
        int count = 0;
        float sumU = 0.0f;
        Vec<DIM> sumG = vzero<DIM>();
        for (int t0 = begin; t0 < end; t0 += kBvcTile) {
            const int nt = min(kBvcTile, end - t0);
            __syncthreads();
            const float4 *src = reinterpret_cast<const float4 *>(a.samples + t0);
            for (int k = threadIdx.x; k < 3*nt; k += kBvcThreads) tile[k] = __ldg(src + k);
            __syncthreads();
            if (!evaluate) continue;
#pragma unroll 2
            for (int j = 0; j < nt; j++) {
                float4 s0 = tile[3*j], s1 = tile[3*j + 1], s2 = tile[3*j + 2];
                Vec<DIM> xy, n;
                xy[0] = x[0] - s0.x; xy[1] = x[1] - s0.y;
                n[0] = s0.w; n[1] = s1.x;
                if (DIM == 3) { xy[DIM - 1] = x[DIM - 1] - s0.z; n[DIM - 1] = s1.y; }
                float r = fmaxf(a.radiusClamp, sqrtf(dot(xy, xy)));
                float G, P;
                Vec<DIM> dG, dP;
                freeSpaceKernels<DIM, YUKAWA>(xy, r, n, sqrtLambda, G, P, dG, dP);
                // :706-709 / :789-791: skip non-finite kernels
                float finiteProbe = b == 4 ? G + dot(dG, dG) : G + P + dot(dG, dG) + dot(dP, dP);
                if (!(fabsf(finiteProbe) <= kMaxFloat)) continue;
                if (a.kernelRegularization > 0.0f) regularize<DIM>(r*invReg, G, P);

                const float invPdf = s1.z, solution = s1.w, normalDerivative = s2.x, robin = s2.y, mu = s2.z;
                float u;
                Vec<DIM> g;
                if (b == 4) { // splatSourceData: contribution holds the source value
                    u = G*robin;
                    for (int k = 0; k < DIM; k++) g[k] = dG[k]*robin;
                } else if (fabsf(mu) > a.robinCutoff) {
                    float invMu = 1.0f/mu;
                    u = (G + P*invMu)*normalDerivative - P*robin*invMu;
                    for (int k = 0; k < DIM; k++) g[k] = (dG[k] + dP[k]*invMu)*normalDerivative - dP[k]*robin*invMu;
                } else if (fabsf(mu) > 0.0f) {
                    u = G*robin - (P + mu*G)*solution;
                    for (int k = 0; k < DIM; k++) g[k] = dG[k]*robin - (dP[k] + mu*dG[k])*solution;
                } else {
                    u = G*normalDerivative - P*solution;
                    for (int k = 0; k < DIM; k++) g[k] = dG[k]*normalDerivative - dP[k]*solution;
                }
                count++;
                sumU += alpha*u*invPdf;
                for (int k = 0; k < DIM; k++) sumG[k] += alphaGrad*g[k]*invPdf;
            }
        }
        if (live) {
            // fold into the caller's running (count, sums) of this bucket and take its mean
            size_t slot = (size_t)i*kBvcBuckets + b;
            int n = a.bucketCount[slot] + count;
            float *sums = a.bucketSum + slot*(1 + DIM);
            float su = sums[0] + sumU;
            a.bucketCount[slot] = n;
            sums[0] = su;
            float invN = n > 0 ? 1.0f/(float)n : 0.0f;
            totalU += su*invN;
            for (int k = 0; k < DIM; k++) {
                float sg = sums[1 + k] + sumG[k];
                sums[1 + k] = sg;
                totalG[k] += sg*invN;
            }
        }
    }
    if (live) {
        a.solution[i] = totalU;                       // EvaluationPoint::getEstimatedSolution (:263-273)
        storeVec<DIM>(a.gradient + (size_t)DIM*i, totalG); // getEstimatedGradient (:276-286)
    }
}

} // namespace zb
