// Green's functions of the Laplace (harmonic) and screened Poisson (Yukawa) operators on a ball,
// centred at the ball centre, in 2D and 3D: the per-step math of the walk
// (zombie/core/distributions.h:454-1005). One value type with a runtime harmonic/Yukawa switch
// replaces the reference's heap-allocated virtual class per walk (walk_on_stars.h:619-625).
#pragma once

#include "rng.cuh"

namespace zb {

// Modified Bessel functions I0, I1, K0, K1: polynomial approximations of Abramowitz & Stegun
// 9.8.1-9.8.8 evaluated in double, as zombie/core/bessel.h:219-343 does.
namespace bessel {
ZB_HOSTDEV double i0(double x) {
    double ax = fabs(x);
    if (ax < 3.75) {
        double y = x/3.75; y *= y;
        return 1.0 + y*(3.5156229 + y*(3.0899424 + y*(1.2067492 + y*(0.2659732 + y*(0.360768e-1 + y*0.45813e-2)))));
    }
    double y = 3.75/ax;
    return (exp(ax)/sqrt(ax))*(0.39894228 + y*(0.1328592e-1 + y*(0.225319e-2 + y*(-0.157565e-2 + y*(0.916281e-2
           + y*(-0.2057706e-1 + y*(0.2635537e-1 + y*(-0.1647633e-1 + y*0.392377e-2))))))));
}
ZB_HOSTDEV double i1(double x) {
    double ax = fabs(x), ans;
    if (ax < 3.75) {
        double y = x/3.75; y *= y;
        ans = ax*(0.5 + y*(0.87890594 + y*(0.51498869 + y*(0.15084934 + y*(0.2658733e-1 + y*(0.301532e-2 + y*0.32411e-3))))));
    } else {
        double y = 3.75/ax;
        ans = 0.2282967e-1 + y*(-0.2895312e-1 + y*(0.1787654e-1 - y*0.420059e-2));
        ans = 0.39894228 + y*(-0.3988024e-1 + y*(-0.362018e-2 + y*(0.163801e-2 + y*(-0.1031555e-1 + y*ans))));
        ans *= exp(ax)/sqrt(ax);
    }
    return x < 0.0 ? -ans : ans;
}
ZB_HOSTDEV double k0(double x) {
    if (x <= 2.0) {
        double y = x*x/4.0;
        return (-log(x/2.0)*i0(x)) + (-0.57721566 + y*(0.42278420 + y*(0.23069756 + y*(0.3488590e-1 + y*(0.262698e-2
               + y*(0.10750e-3 + y*0.74e-5))))));
    }
    double y = 2.0/x;
    return (exp(-x)/sqrt(x))*(1.25331414 + y*(-0.7832358e-1 + y*(0.2189568e-1 + y*(-0.1062446e-1 + y*(0.587872e-2
           + y*(-0.251540e-2 + y*0.53208e-3))))));
}
ZB_HOSTDEV double k1(double x) {
    if (x <= 2.0) {
        double y = x*x/4.0;
        return (log(x/2.0)*i1(x)) + (1.0/x)*(1.0 + y*(0.15443144 + y*(-0.67278579 + y*(-0.18156897 + y*(-0.1919402e-1
               + y*(-0.110404e-2 + y*(-0.4686e-4)))))));
    }
    double y = 2.0/x;
    return (exp(-x)/sqrt(x))*(1.25331414 + y*(0.23498619 + y*(-0.3655620e-1 + y*(0.1504268e-1 + y*(-0.780353e-2
           + y*(0.325614e-2 + y*(-0.68245e-3)))))));
}
} // namespace bessel

template <int DIM>
ZB_HOSTDEV float sphereDirectionPdf(float r) // SphereSampler::pdfSampleSphereUniform (sampling.h:140-142, 226-228)
{
    return DIM == 3 ? 1.0f/(4.0f*kPiF*r*r) : 1.0f/(2.0f*kPiF*r);
}

template <int DIM>
struct GreensBall {
    Vec<DIM> c;
    float R;
    float rClamp;
    float sqrtLambda; // 0 => harmonic
    bool yukawa;
    // 3D Yukawa: muR, exp(-muR), sinh(muR), K_{3/2}(muR), I_{3/2}(muR)   (distributions.h:843-852)
    // 2D Yukawa: muR, K0(muR), I0(muR), K1(muR), I1(muR)                  (distributions.h:685-692)
    float muR, t0, t1, t2, t3;

    ZB_HOSTDEV void init(bool useYukawa, float lambda) {
        yukawa = useYukawa;
        sqrtLambda = useYukawa ? sqrtf(lambda) : 0.0f;
        R = 0.0f; rClamp = 1e-4f;
        c = vzero<DIM>();
        muR = t0 = t1 = t2 = t3 = 0.0f;
    }

    ZB_HOSTDEV void updateBall(const Vec<DIM>& center, float radius) {
        c = center;
        R = radius;
        rClamp = 1e-4f;
        if (!yukawa) return;
        muR = R*sqrtLambda;
        if (DIM == 3) {
            t0 = expf(-muR);
            float e2 = t0*t0;
            float coshmuR = (1.0f + e2)/(2.0f*t0);
            t1 = (1.0f - e2)/(2.0f*t0);
            t2 = t0*(1.0f + 1.0f/muR);
            t3 = coshmuR - t1/muR;
        } else {
            t0 = (float)bessel::k0(muR);
            t1 = (float)bessel::i0(muR);
            t2 = (float)bessel::k1(muR);
            t3 = (float)bessel::i1(muR);
        }
    }

    ZB_HOSTDEV float evaluate(float r) const {
        if (!yukawa) {
            return DIM == 3 ? (1.0f/r - 1.0f/R)/(4.0f*kPiF) : logf(R/r)/(2.0f*kPiF);
        }
        float mur = r*sqrtLambda;
        if (DIM == 3) {
            float e = expf(-mur);
            float sinhmur = (1.0f - e*e)/(2.0f*e);
            return (e - t0*sinhmur/t1)/(4.0f*kPiF*r);
        }
        float K0 = (float)bessel::k0(mur), I0 = (float)bessel::i0(mur);
        return (K0 - I0*t0/t1)/(2.0f*kPiF);
    }

    ZB_HOSTDEV float poissonKernel() const {
        if (!yukawa) return DIM == 3 ? 1.0f/(4.0f*kPiF) : 1.0f/(2.0f*kPiF);
        return DIM == 3 ? muR/(4.0f*kPiF*t1) : 1.0f/(2.0f*kPiF*t1);
    }

    ZB_HOSTDEV float norm() const {
        if (!yukawa) return DIM == 3 ? R*R/6.0f : R*R/4.0f;
        float lambda = sqrtLambda*sqrtLambda;
        return (1.0f - (DIM == 3 ? 4.0f : 2.0f)*kPiF*poissonKernel())/lambda;
    }

    ZB_HOSTDEV float gradientNorm(float r) const {
        if (!yukawa) {
            if (DIM == 3) { float r3 = r*r*r; return (1.0f/r3 - 1.0f/(R*R*R))/(4.0f*kPiF); }
            float r2 = r*r;
            return (1.0f/r2 - 1.0f/(R*R))/(2.0f*kPiF);
        }
        float mur = r*sqrtLambda;
        if (DIM == 3) {
            float e = expf(-mur), e2 = e*e;
            float coshmur = (1.0f + e2)/(2.0f*e), sinhmur = (1.0f - e2)/(2.0f*e);
            float K32 = e*(1.0f + 1.0f/mur), I32 = coshmur - sinhmur/mur;
            float Qr = sqrtLambda*(K32 - I32*t2/t3);
            return Qr/(4.0f*kPiF*r*r);
        }
        float K1 = (float)bessel::k1(mur), I1 = (float)bessel::i1(mur);
        float Qr = sqrtLambda*(K1 - I1*t2/t3);
        return Qr/(2.0f*kPiF*r);
    }

    ZB_HOSTDEV Vec<DIM> gradient(float r, const Vec<DIM>& y) const { return (y - c)*gradientNorm(r); }

    // radial dampening of the centred Poisson kernel (Yukawa only)
    ZB_HOSTDEV float dampening(float r) const {
        float mur = r*sqrtLambda;
        if (DIM == 3) {
            float e = expf(-mur), e2 = e*e;
            float coshmur = (1.0f + e2)/(2.0f*e), sinhmur = (1.0f - e2)/(2.0f*e);
            float K32 = e*(1.0f + 1.0f/mur), I32 = coshmur - sinhmur/mur;
            return mur*(K32 + I32*t0/t1);
        }
        float K1 = (float)bessel::k1(mur), I1 = (float)bessel::i1(mur);
        return mur*(K1 + I1*t0/t1);
    }

    ZB_HOSTDEV float directionSampledPoissonKernel(const Vec<DIM>& y) const {
        if (!yukawa) return 1.0f;
        return dampening(fmaxf(rClamp, zb::norm(y - c)));
    }

    // 1 - mu G / P for Robin boundaries (distributions.h:533-541, 643-651, 796-809, 968-982)
    ZB_HOSTDEV float reflectance(float r, const Vec<DIM>& dir, const Vec<DIM>& n, float robinCoeff) const {
        float Q = yukawa ? dampening(r) : 1.0f;
        if (fabsf(robinCoeff) > 0.0f) {
            float P = DIM == 3 ? dot(n, dir)/(r*r) : dot(n, dir)/r;
            float G;
            if (!yukawa) {
                G = DIM == 3 ? 1.0f/r - 1.0f/R : logf(R/r);
            } else {
                float mur = r*sqrtLambda;
                if (DIM == 3) {
                    float e = expf(-mur);
                    float sinhmur = (1.0f - e*e)/(2.0f*e);
                    G = (e - t0*sinhmur/t1)/r;
                } else {
                    G = (float)bessel::k0(mur) - (float)bessel::i0(mur)*t0/t1;
                }
            }
            return Q - robinCoeff*G/P;
        }
        return Q;
    }

    ZB_HOSTDEV Vec<DIM> poissonKernelGradient(const Vec<DIM>& y) const {
        Vec<DIM> d = y - c;
        if (!yukawa) return DIM == 3 ? d*(3.0f/(4.0f*kPiF*R*R)) : d*(2.0f/(2.0f*kPiF*R*R));
        if (DIM == 3) {
            float lambda = sqrtLambda*sqrtLambda;
            return d*((lambda/t3)/(4.0f*kPiF));
        }
        return d*((sqrtLambda/(R*t3))/(2.0f*kPiF));
    }

    // radius r in (0, R] distributed proportionally to r^{DIM-1} G(r); returns c + r*dir
    ZB_HOSTDEV Vec<DIM> sampleVolume(const Vec<DIM>& dir, Pcg32& rng, float& r, float& pdf) const {
        if (!yukawa) {
            if (DIM == 3) {
                // median of three uniforms has density 6u(1-u) (distributions.h:565-577)
                float u1 = rng.nextFloat(), u2 = rng.nextFloat(), u3 = rng.nextFloat();
                r = R*fmaxf(fminf(u1, u2), fminf(fmaxf(u1, u2), u3));
            } else {
                float u1 = rng.nextFloat(), u2 = rng.nextFloat();
                r = R*sqrtf(u1*u2);
            }
            r = fmaxf(rClamp, r);
            if (r > R) r = R/2.0f;
            pdf = evaluate(r)/norm();
            return c + dir*r;
        }
        // rejection sampling with the reference's envelope constants (distributions.h:419-440, 693-703, 854-864)
        float lambda = sqrtLambda*sqrtLambda;
        float a = DIM == 3 ? 2.0f : 2.2f, b = DIM == 3 ? 0.5f : 0.6f;
        float bound = R <= lambda ? fmaxf(fmaxf(a/R, a/lambda), fmaxf(b*sqrtf(R), b*sqrtLambda))
                                  : fmaxf(fminf(a/R, a/lambda), fminf(b*sqrtf(R), b*sqrtLambda));
        float fnNorm = norm();
        // For small balls (mu R ~ 1e-3: every step taken from a Robin boundary, whose star radius is the epsilon
        // shell) 1 - mu R/sinh(mu R) cancels catastrophically in float and the norm comes out <= 0 about as often as
        // not. Then pdfRadius <= 0 for every candidate, the reference's loop rejects 1000 candidates in a row and
        // keeps the last one (distributions.h:425-436); drawing that one candidate directly gives the same
        // distribution of (r, pdf) without stalling the other 31 walks of the warp for 1000 iterations.
        int iter = fnNorm > 0.0f ? 0 : 999;
        do {
            float u = rng.nextFloat();
            r = rng.nextFloat()*R;
            pdf = evaluate(r)/fnNorm;
            float pdfRadius = pdf/sphereDirectionPdf<DIM>(r);
            iter++;
            if (u < pdfRadius/bound) break;
        } while (iter < 1000);
        r = fmaxf(rClamp, r);
        if (r > R) r = R/2.0f;
        return c + dir*r;
    }
};

} // namespace zb
