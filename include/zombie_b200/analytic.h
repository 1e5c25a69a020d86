// Manufactured-solution PDE families shared by the device PDE evaluator and the test oracle.
//
// The reference describes a PDE through host std::function callbacks (zombie/core/pde.h:31-44),
// which cannot run inside a CUDA kernel. The B200 path therefore takes a *data-described* PDE:
// a closed-form solution family u(x) with its gradient and Laplacian, from which every callback
// of zombie::PDE is derived:
//   dirichlet(x)      = u(x)                                  (pde.h:34)
//   source(x)         = sigma*u(x) - lap u(x)                  (estimator solves lap u - sigma u = -f)
//   robin(x, n)       = n.grad u(x) + mu*u(x)                  (pde.h:37, n = boundary normal passed in)
//   robinCoeff(x, n)  = mu                                     (pde.h:38)
// This header is plain C++ (usable from nvcc and g++): the oracle's wrapper around the reference
// (oracle/ref_capi.cpp) builds std::function callbacks from the same formulas so both sides solve
// the identical problem.
#pragma once

#include <math.h>
#include "../zombie_b200.h"

#if defined(__CUDACC__)
#define ZB_HD __host__ __device__ __forceinline__
#else
#define ZB_HD inline
#endif

enum zb_solution_family {
    ZB_SOL_CONSTANT = 0,   // u = p[0]
    ZB_SOL_EXP_DIR = 1,    // u = p[3]*exp(k.x), k = p[0..2]           (lap u = |k|^2 u)
    ZB_SOL_QUADRATIC = 2,  // u = p[0]x^2 + p[1]y^2 + p[2]z^2 + p[3]    (lap u = 2(p0+p1+p2))
    ZB_SOL_INV_DIST = 3,   // u = p[3]/|x - x0|, x0 = p[0..2]           (harmonic away from x0, 3D)
    ZB_SOL_SIN_EXP_2D = 4, // u = sin(pi x) e^{pi y} + x^2              (lap u = 2, 2D)
    ZB_SOL_CUBIC_2D = 5,   // u = Re((x + iy)^3) = x^3 - 3xy^2          (harmonic, 2D)
    ZB_SOL_LINEAR = 6,     // u = p[0]x + p[1]y + p[2]z + p[3]          (harmonic)
    ZB_SOL_TRIG_3D = 7     // u = sin(p0 x) sin(p1 y) sin(p2 z)         (lap u = -(p0^2+p1^2+p2^2) u)
};

struct zb_jet { float u; float g[3]; float lap; };

// evaluates u, grad u and lap u of the family at x (x[2] ignored by the 2D families)
ZB_HD zb_jet zb_eval_solution(int family, const float *p, const float *x)
{
    zb_jet j;
    j.u = 0.0f; j.g[0] = j.g[1] = j.g[2] = 0.0f; j.lap = 0.0f;
    switch (family) {
    case ZB_SOL_CONSTANT:
        j.u = p[0];
        break;
    case ZB_SOL_EXP_DIR: {
        float e = p[3]*expf(p[0]*x[0] + p[1]*x[1] + p[2]*x[2]);
        j.u = e;
        j.g[0] = p[0]*e; j.g[1] = p[1]*e; j.g[2] = p[2]*e;
        j.lap = (p[0]*p[0] + p[1]*p[1] + p[2]*p[2])*e;
    } break;
    case ZB_SOL_QUADRATIC:
        j.u = p[0]*x[0]*x[0] + p[1]*x[1]*x[1] + p[2]*x[2]*x[2] + p[3];
        j.g[0] = 2.0f*p[0]*x[0]; j.g[1] = 2.0f*p[1]*x[1]; j.g[2] = 2.0f*p[2]*x[2];
        j.lap = 2.0f*(p[0] + p[1] + p[2]);
        break;
    case ZB_SOL_INV_DIST: {
        float dx = x[0] - p[0], dy = x[1] - p[1], dz = x[2] - p[2];
        float r2 = dx*dx + dy*dy + dz*dz;
        float r = sqrtf(r2);
        float ir3 = p[3]/(r2*r);
        j.u = p[3]/r;
        j.g[0] = -dx*ir3; j.g[1] = -dy*ir3; j.g[2] = -dz*ir3;
    } break;
    case ZB_SOL_SIN_EXP_2D: {
        const float pi = 3.14159265358979323846f;
        float s = sinf(pi*x[0]), c = cosf(pi*x[0]), e = expf(pi*x[1]);
        j.u = s*e + x[0]*x[0];
        j.g[0] = pi*c*e + 2.0f*x[0]; j.g[1] = pi*s*e;
        j.lap = 2.0f;
    } break;
    case ZB_SOL_CUBIC_2D:
        j.u = x[0]*x[0]*x[0] - 3.0f*x[0]*x[1]*x[1];
        j.g[0] = 3.0f*x[0]*x[0] - 3.0f*x[1]*x[1]; j.g[1] = -6.0f*x[0]*x[1];
        break;
    case ZB_SOL_LINEAR:
        j.u = p[0]*x[0] + p[1]*x[1] + p[2]*x[2] + p[3];
        j.g[0] = p[0]; j.g[1] = p[1]; j.g[2] = p[2];
        break;
    case ZB_SOL_TRIG_3D: {
        float sx = sinf(p[0]*x[0]), sy = sinf(p[1]*x[1]), sz = sinf(p[2]*x[2]);
        float cx = cosf(p[0]*x[0]), cy = cosf(p[1]*x[1]), cz = cosf(p[2]*x[2]);
        j.u = sx*sy*sz;
        j.g[0] = p[0]*cx*sy*sz; j.g[1] = p[1]*sx*cy*sz; j.g[2] = p[2]*sx*sy*cz;
        j.lap = -(p[0]*p[0] + p[1]*p[1] + p[2]*p[2])*j.u;
    } break;
    default: break;
    }
    return j;
}

// Sample of a zb_grid_view at x (x[2] ignored when shape[2] == 1). Shared by the device evaluator and the oracle's
// std::function callbacks, so both sides read identical values. Nearest: the sample whose cell contains x, like the
// reference demo's Image::get(uv) (demo/image.h); linear: multilinear between the surrounding samples. Positions outside
// the box clamp to its faces.
ZB_HD float zb_grid_sample(const zb_grid_view& g, const float *x)
{
    float t[3];
    int i0[3], i1[3];
    float w[3];
    for (int k = 0; k < 3; k++) {
        int n = g.shape[k];
        if (n <= 1) { i0[k] = i1[k] = 0; w[k] = 0.0f; continue; }
        t[k] = (x[k] - g.origin[k])*g.invCell[k];
        if (g.interp == ZB_GRID_NEAREST) {
            float r = floorf(t[k] + 0.5f);
            r = r < 0.0f ? 0.0f : (r > (float)(n - 1) ? (float)(n - 1) : r);
            i0[k] = i1[k] = (int)r;
            w[k] = 0.0f;
        } else {
            float c = t[k] < 0.0f ? 0.0f : (t[k] > (float)(n - 1) ? (float)(n - 1) : t[k]);
            float f = floorf(c);
            int i = (int)f;
            if (i > n - 2) i = n - 2;
            i0[k] = i; i1[k] = i + 1;
            w[k] = c - (float)i;
        }
    }
    const int sx = g.shape[0], sxy = g.shape[0]*g.shape[1];
    if (g.interp == ZB_GRID_NEAREST) return g.data[(long long)i0[2]*sxy + (long long)i0[1]*sx + i0[0]];
    float acc = 0.0f;
    for (int c = 0; c < 8; c++) {
        int ix = (c & 1) ? i1[0] : i0[0], iy = (c & 2) ? i1[1] : i0[1], iz = (c & 4) ? i1[2] : i0[2];
        float wt = ((c & 1) ? w[0] : 1.0f - w[0])*((c & 2) ? w[1] : 1.0f - w[1])*((c & 4) ? w[2] : 1.0f - w[2]);
        if (wt != 0.0f) acc += wt*g.data[(long long)iz*sxy + (long long)iy*sx + ix];
    }
    return acc;
}
