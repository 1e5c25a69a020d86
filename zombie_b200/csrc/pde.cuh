// Device evaluation of the data-described zombie::PDE (zombie/core/pde.h:20-45). The reference
// evaluates host std::function callbacks; here each callback is derived on the fly from a
// closed-form solution family (include/zombie_b200/analytic.h) or is a constant.
#pragma once

#include "../../include/zombie_b200.h"
#include "../../include/zombie_b200/analytic.h"
#include "vec.cuh"

namespace zb {

template <int DIM>
struct PdeEval {
    const zb_pde_desc& d; // lives in the kernel parameter block

    ZB_HOSTDEV zb_jet jet(const Vec<DIM>& x) const {
        float xx[3] = {x[0], x[1], DIM == 3 ? x[DIM - 1] : 0.0f};
        return zb_eval_solution(d.family, d.params, xx);
    }
    // pde.source: the estimator solves lap u - sigma u = -f
    ZB_HOSTDEV float source(const Vec<DIM>& x) const {
        if (d.mode == ZB_PDE_CONSTANT) return d.sourceValue;
        zb_jet j = jet(x);
        return d.absorptionCoeff*j.u - j.lap;
    }
    // pde.dirichlet
    ZB_HOSTDEV float dirichlet(const Vec<DIM>& x) const {
        if (d.mode == ZB_PDE_CONSTANT) return d.dirichletValue;
        return jet(x).u;
    }
    // pde.robin: du/dn + mu u with n the boundary normal handed to the callback
    ZB_HOSTDEV float robin(const Vec<DIM>& x, const Vec<DIM>& n) const {
        if (d.mode == ZB_PDE_CONSTANT) return d.robinValue;
        zb_jet j = jet(x);
        float dn = 0.0f;
        for (int k = 0; k < DIM; k++) dn += n[k]*j.g[k];
        return dn + d.robinCoeff*j.u;
    }
    ZB_HOSTDEV float robinCoeff() const { return d.robinCoeff; }
    // PDE::hasNonZeroRobinCoeff (pde.h:59-76)
    ZB_HOSTDEV bool hasNonZeroRobinCoeff() const {
        return !d.areRobinConditionsPureNeumann && fabsf(d.robinCoeff) > 0.0f;
    }
};

} // namespace zb
