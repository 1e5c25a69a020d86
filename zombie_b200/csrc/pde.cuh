// Device evaluation of the data-described zombie::PDE (zombie/core/pde.h:20-45). The reference
// evaluates host std::function callbacks; here each callback is derived on the fly from a
// closed-form solution family (include/zombie_b200/analytic.h) or is a constant.
#pragma once

#include "../../include/zombie_b200.h"
#include "../../include/zombie_b200/analytic.h"
#include "vec.cuh"

namespace zb {

// One out-of-line copy of the solution families per kernel: the switch over eight families (with the slow paths of
// sinf / cosf / expf) is ~1000 instructions, and source / dirichlet / robin are evaluated at six places of the walk
// kernel. Every family reads p[0..3] only.
ZB_DEV_COLD zb_jet evalSolutionCall(int family, float p0, float p1, float p2, float p3, float x0, float x1, float x2)
{
    float p[4] = {p0, p1, p2, p3}, x[3] = {x0, x1, x2};
    return zb_eval_solution(family, p, x);
}

template <int DIM>
struct PdeEval {
    const zb_pde_desc& d; // lives in the kernel parameter block

    ZB_HOSTDEV zb_jet jet(const Vec<DIM>& x) const {
        return evalSolutionCall(d.family, d.params[0], d.params[1], d.params[2], d.params[3], x[0], x[1], DIM == 3 ? x[DIM - 1] : 0.0f);
    }
    // ---- the callbacks of the PDE as the caller described it --------------------------------------
    // pde.source: the estimator solves lap u - sigma u = -f
    ZB_HOSTDEV float grid(const zb_grid_view& g, const Vec<DIM>& x) const {
        float p[3] = {x[0], x[1], DIM == 3 ? x[DIM - 1] : 0.0f};
        return zb_grid_sample(g, p);
    }
    ZB_HOSTDEV float baseSource(const Vec<DIM>& x) const {
        if (d.sourceGrid.data) return grid(d.sourceGrid, x);
        if (d.mode == ZB_PDE_CONSTANT) return d.sourceValue;
        zb_jet j = jet(x);
        return d.absorptionCoeff*j.u - j.lap;
    }
    // aligned: returnBoundaryNormalAlignedValue — the side of a double-sided absorbing boundary its normal points to
    ZB_HOSTDEV float baseDirichlet(const Vec<DIM>& x, bool aligned) const {
        if (aligned && d.dirichletAlignedGrid.data) return grid(d.dirichletAlignedGrid, x);
        if (d.dirichletGrid.data) return grid(d.dirichletGrid, x);
        if (d.mode == ZB_PDE_CONSTANT) return aligned && d.hasDirichletAlignedValue ? d.dirichletAlignedValue : d.dirichletValue;
        return jet(x).u;
    }
    // du/dn + mu u with n the boundary normal handed to the callback
    ZB_HOSTDEV float baseRobin(const Vec<DIM>& x, const Vec<DIM>& n, bool aligned) const {
        if (aligned && d.robinAlignedGrid.data) return grid(d.robinAlignedGrid, x);
        if (d.robinGrid.data) return grid(d.robinGrid, x);
        if (d.mode == ZB_PDE_CONSTANT) return d.robinValue;
        zb_jet j = jet(x);
        float dn = 0.0f;
        for (int k = 0; k < DIM; k++) dn += n[k]*j.g[k];
        return dn + baseRobinCoeff(x)*j.u;
    }
    ZB_HOSTDEV float baseRobinCoeff(const Vec<DIM>& x) const {
        return d.robinCoeffGrid.data ? grid(d.robinCoeffGrid, x) : d.robinCoeff;
    }
    // y on the inverted domain -> x on the exterior domain (kelvin_transform.h:288-289)
    ZB_HOSTDEV Vec<DIM> exteriorPoint(const Vec<DIM>& y, float r2) const {
        Vec<DIM> x;
        for (int k = 0; k < DIM; k++) x[k] = y[k]/r2 + d.kelvinOrigin[k];
        return x;
    }
    // ---- what the walk calls: wrapped like KelvinTransform::transformPde (kelvin_transform.h:282-348)
    // when d.kelvin is set, the plain callbacks otherwise
    ZB_HOSTDEV float source(const Vec<DIM>& y) const {
        if (!d.kelvin) return baseSource(y);
        float r2 = dot(y, y);
        float rPow = r2*r2;
        if (DIM == 3) rPow *= sqrtf(r2);
        return baseSource(exteriorPoint(y, r2))/rPow;
    }
    ZB_HOSTDEV float dirichlet(const Vec<DIM>& y, bool aligned = false) const {
        if (!d.kelvin) return baseDirichlet(y, aligned);
        float r2 = dot(y, y);
        float rPow = DIM == 3 ? sqrtf(r2) : 1.0f;
        return baseDirichlet(exteriorPoint(y, r2), aligned)/rPow;
    }
    // aligned: returnBoundaryNormalAlignedValue of the reference's callback (walk_on_stars.h:242-244)
    ZB_HOSTDEV float robin(const Vec<DIM>& y, const Vec<DIM>& N, bool aligned = false) const {
        if (!d.kelvin) return baseRobin(y, N, aligned);
        float r2 = dot(y, y);
        Vec<DIM> n = N - y*(2.0f*dot(N, y))/r2;
        float rPow = r2;
        if (DIM == 3) rPow *= sqrtf(r2);
        return baseRobin(exteriorPoint(y, r2), n, false)/rPow; // the wrapper drops the flag (kelvin_transform.h:311-322)
    }
    ZB_HOSTDEV float robinCoeff(const Vec<DIM>& y, const Vec<DIM>& N) const {
        if (!d.kelvin) return baseRobinCoeff(y);
        float r2 = dot(y, y);
        float P = dot(N, y)/r2;
        float shift = DIM == 3 ? P : 0.0f;
        return baseRobinCoeff(exteriorPoint(y, r2))/r2 + shift;
    }
    // PDE::hasNonZeroRobinCoeff (pde.h:59-76): probes robinCoeff with the normal e_0
    ZB_HOSTDEV bool hasNonZeroRobinCoeff(const Vec<DIM>& y) const {
        if (d.areRobinConditionsPureNeumann) return false;
        Vec<DIM> n = vzero<DIM>();
        n[0] = 1.0f;
        return fabsf(robinCoeff(y, n)) > 0.0f;
    }
};

} // namespace zb
