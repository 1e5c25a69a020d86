// zombie_b200/zombie.hpp — host C++ mirror of Zombie's operator API for the random-walk hot path,
// over the C ABI of libzombie_b200.so (include/zombie_b200.h).
//
// The reference is a header-only template library whose "operator API" for this path is
//   zombie::PDE<T,DIM>                       zombie/core/pde.h:20-45
//   zombie::GeometricQueries<DIM>            zombie/core/geometric_queries.h:59-124
//   zombie::WalkSettings / SamplePoint / SampleStatistics / SampleType / EstimationQuantity
//                                            zombie/point_estimation/common.h:19-395
//   zombie::Fcpw{Dirichlet,Neumann,Robin}BoundaryHandler + populateGeometricQueriesFor*Boundary
//                                            zombie/utils/fcpw_geometric_queries.h:75-150
//   zombie::WalkOnSpheres<T,DIM>::solve      zombie/point_estimation/walk_on_spheres.h:26-44
//   zombie::WalkOnStars<T,DIM>::solve        zombie/point_estimation/walk_on_stars.h:25-46
// This header keeps those names, member names, argument orders and defaults, so user code written
// against <zombie/zombie.h> compiles against it after `namespace zombie = zombie_b200;` (defined
// below unless ZOMBIE_B200_NO_ALIAS). Differences, all forced by "callbacks cannot run on the GPU":
//   * PDE carries a data description (zb_pde_desc) from which its std::function members are
//     generated; a PDE whose callbacks were replaced by arbitrary lambdas is rejected by solve();
//   * solve() throws zombie_b200::Error (what() = zb_last_error()) instead of exit(EXIT_FAILURE);
//   * T is float (the reference also instantiates Array<float,N> channels);
//   * Vector<DIM> is Eigen's when <Eigen/Core> is available, otherwise a minimal POD with the
//     members this surface needs.
// Nothing in this header computes a walk: there is no CPU fallback.
#pragma once

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../zombie_b200.h"
#include "analytic.h"

#if !defined(ZOMBIE_B200_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define ZOMBIE_B200_HAS_EIGEN 1
#endif
#endif

namespace zombie_b200 {

// ------------------------------------------------------------------------------------------------
// vectors (zombie/core/pde.h:14-17)
// ------------------------------------------------------------------------------------------------
#if defined(ZOMBIE_B200_HAS_EIGEN)
template <size_t DIM> using Vector = Eigen::Matrix<float, DIM, 1>;
template <size_t DIM> using Vectori = Eigen::Matrix<int, DIM, 1>;
#else
template <typename S, size_t DIM>
struct VectorBase {
    S v[DIM];
    static VectorBase Zero() { VectorBase r; for (size_t i = 0; i < DIM; i++) r.v[i] = S(0); return r; }
    static VectorBase Constant(S c) { VectorBase r; for (size_t i = 0; i < DIM; i++) r.v[i] = c; return r; }
    S& operator[](size_t i) { return v[i]; }
    const S& operator[](size_t i) const { return v[i]; }
    S& operator()(size_t i) { return v[i]; }
    const S& operator()(size_t i) const { return v[i]; }
    S *data() { return v; }
    const S *data() const { return v; }
    VectorBase operator+(const VectorBase& o) const { VectorBase r; for (size_t i = 0; i < DIM; i++) r.v[i] = v[i] + o.v[i]; return r; }
    VectorBase operator-(const VectorBase& o) const { VectorBase r; for (size_t i = 0; i < DIM; i++) r.v[i] = v[i] - o.v[i]; return r; }
    VectorBase operator-() const { VectorBase r; for (size_t i = 0; i < DIM; i++) r.v[i] = -v[i]; return r; }
    VectorBase operator*(S s) const { VectorBase r; for (size_t i = 0; i < DIM; i++) r.v[i] = v[i]*s; return r; }
    VectorBase operator/(S s) const { VectorBase r; for (size_t i = 0; i < DIM; i++) r.v[i] = v[i]/s; return r; }
    S dot(const VectorBase& o) const { S s = S(0); for (size_t i = 0; i < DIM; i++) s += v[i]*o.v[i]; return s; }
    S squaredNorm() const { return dot(*this); }
    S norm() const { return std::sqrt(squaredNorm()); }
};
template <size_t DIM> using Vector = VectorBase<float, DIM>;
template <size_t DIM> using Vectori = VectorBase<int, DIM>;
#endif
using Vector2 = Vector<2>;
using Vector3 = Vector<3>;
using Vector2i = Vectori<2>;
using Vector3i = Vectori<3>;

// errors: the reference prints to std::cerr and exit()s; the C ABI returns codes, this layer throws
class Error : public std::runtime_error {
public:
    Error(int code_, const std::string& what_): std::runtime_error(what_), code(code_) {}
    int code;
};

namespace detail {
inline void check(int rc) { if (rc != ZB_OK) throw Error(rc, zb_last_error()); }

// one zb_context per device, shared by everything built through this header
struct ContextHolder {
    zb_context *ctx = nullptr;
    explicit ContextHolder(int device) { check(zb_context_create(device, &ctx)); }
    ~ContextHolder() { zb_context_destroy(ctx); }
    ContextHolder(const ContextHolder&) = delete;
    ContextHolder& operator=(const ContextHolder&) = delete;
};
inline int& deviceIndex() { static int device = 0; return device; }
inline std::shared_ptr<ContextHolder> context()
{
    static std::shared_ptr<ContextHolder> holder;
    static int heldDevice = -1;
    if (!holder || heldDevice != deviceIndex()) {
        holder = std::make_shared<ContextHolder>(deviceIndex());
        heldDevice = deviceIndex();
    }
    return holder;
}
} // namespace detail

// selects the CUDA device used by everything created afterwards (one process drives one GPU; evaluation
// points are sharded over processes — see INTEGRATION.md)
inline void setDevice(int device) { detail::deviceIndex() = device; }

// ------------------------------------------------------------------------------------------------
// pcg32 (deps/pcg32/pcg32.h:40-207) — the published PCG-XSH-RR 64/32 generator; SamplePoint owns one
// ------------------------------------------------------------------------------------------------
struct pcg32 {
    uint64_t state, inc;
    pcg32(): state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) {}
    pcg32(uint64_t initstate, uint64_t initseq = 1u) { seed(initstate, initseq); }
    void seed(uint64_t initstate, uint64_t initseq = 1u) {
        state = 0u;
        inc = (initseq << 1u) | 1u;
        nextUInt();
        state += initstate;
        nextUInt();
    }
    uint32_t nextUInt() {
        uint64_t old = state;
        state = old*0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    uint32_t nextUInt(uint32_t bound) {
        uint32_t threshold = (~bound + 1u)%bound;
        for (;;) { uint32_t r = nextUInt(); if (r >= threshold) return r%bound; }
    }
    float nextFloat() {
        union { uint32_t u; float f; } x;
        x.u = (nextUInt() >> 9) | 0x3f800000u;
        return x.f - 1.0f;
    }
};

// zombie/core/sampling.h:36-41
inline pcg32 seedRng()
{
    static std::atomic<uint64_t> streamCounter{0};
    auto now = std::chrono::high_resolution_clock::now();
    uint64_t seed = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(now.time_since_epoch()).count();
    return pcg32(seed, streamCounter.fetch_add(1, std::memory_order_relaxed));
}

// ------------------------------------------------------------------------------------------------
// Sampled PDE data: what the reference's applications put behind the PDE callbacks (demo/model_problem.h:229-266 reads
// source / boundary values from textures). A DataGrid owns the samples on the host (for the std::function members) and
// on the device (zb_grid, for the walk); share it between PDEs with shared_ptr and keep it alive while they are in use.
// ------------------------------------------------------------------------------------------------
class DataGrid {
public:
    // values: nx*ny*nz floats, x fastest (nz = 1 in 2D), covering the box [lo, hi] with both ends included
    DataGrid(const std::vector<float>& values, int nx, int ny, int nz, const float lo[3], const float hi[3],
             zb_grid_interp interp = ZB_GRID_LINEAR): host(values), holder(detail::context()) {
        if ((size_t)nx*(size_t)ny*(size_t)nz != values.size()) throw Error(ZB_ERR_INVALID_ARGUMENT, "DataGrid: size mismatch");
        int32_t shape[3] = {nx, ny, nz};
        detail::check(zb_grid_create(holder->ctx, host.data(), shape, &grid));
        std::memset(&hostView, 0, sizeof(hostView));
        for (int k = 0; k < 3; k++) {
            hostView.shape[k] = shape[k];
            hostView.origin[k] = lo[k];
            hostView.invCell[k] = shape[k] > 1 ? (float)(shape[k] - 1)/(hi[k] - lo[k]) : 0.0f;
        }
        hostView.interp = (int32_t)interp;
        hostView.data = host.data();
        deviceView = hostView;
        deviceView.data = zb_grid_data(grid);
    }
    ~DataGrid() { zb_grid_destroy(grid); }
    DataGrid(const DataGrid&) = delete;
    DataGrid& operator=(const DataGrid&) = delete;
    template <size_t DIM> float operator()(const Vector<DIM>& x) const {
        float xx[3] = {x[0], x[1], DIM == 3 ? x[DIM - 1] : 0.0f};
        return zb_grid_sample(hostView, xx);
    }
    const zb_grid_view& device() const { return deviceView; }
private:
    std::vector<float> host;
    std::shared_ptr<detail::ContextHolder> holder;
    zb_grid *grid = nullptr;
    zb_grid_view hostView, deviceView;
};
enum class PdeData { Source, Dirichlet, Robin, RobinNormalAligned, RobinCoeff, DirichletNormalAligned };

// ------------------------------------------------------------------------------------------------
// PDE (zombie/core/pde.h:20-79)
// ------------------------------------------------------------------------------------------------
template <typename T, size_t DIM>
struct PDE {
    static_assert(std::is_same<T, float>::value, "zombie_b200::PDE: T must be float");
    PDE(): absorptionCoeff(0.0f), isSourceConstant(false), areRobinConditionsPureNeumann(true),
           areRobinCoeffsNonnegative(true), described(false) {
        std::memset(&description, 0, sizeof(description));
        bindHasNonZeroRobinCoeff();
    }
    PDE(const PDE& o) { *this = o; }
    PDE& operator=(const PDE& o) {
        absorptionCoeff = o.absorptionCoeff; isSourceConstant = o.isSourceConstant;
        areRobinConditionsPureNeumann = o.areRobinConditionsPureNeumann;
        areRobinCoeffsNonnegative = o.areRobinCoeffsNonnegative;
        source = o.source; dirichlet = o.dirichlet; robin = o.robin; robinCoeff = o.robinCoeff;
        hasReflectingBoundaryConditions = o.hasReflectingBoundaryConditions;
        description = o.description; described = o.described;
        for (int k = 0; k < 6; k++) grids[k] = o.grids[k];
        bindHasNonZeroRobinCoeff(); // captures `this` (pde.h:60)
        return *this;
    }

    // members (pde.h:26-44)
    float absorptionCoeff;
    bool isSourceConstant;
    bool areRobinConditionsPureNeumann;
    bool areRobinCoeffsNonnegative;
    std::function<T(const Vector<DIM>&)> source;
    std::function<T(const Vector<DIM>&, bool)> dirichlet;
    std::function<T(const Vector<DIM>&, const Vector<DIM>&, bool)> robin;
    std::function<float(const Vector<DIM>&, const Vector<DIM>&, bool)> robinCoeff;
    std::function<bool(const Vector<DIM>&)> hasReflectingBoundaryConditions;
    std::function<bool(const Vector<DIM>&)> hasNonZeroRobinCoeff; // set automatically

    // ---- the data description the device evaluates ------------------------------------------------
    // Closed-form solution family (zombie_b200/analytic.h): dirichlet = u, source = sigma*u - lap u,
    // robin = du/dn + mu*u, robinCoeff = mu; the std::function members are generated to match.
    void setManufactured(zb_solution_family family, const std::vector<float>& params, float robinCoeffValue = 0.0f) {
        std::memset(&description, 0, sizeof(description));
        description.mode = ZB_PDE_MANUFACTURED;
        description.family = (int32_t)family;
        for (size_t i = 0; i < params.size() && i < 16; i++) description.params[i] = params[i];
        description.robinCoeff = robinCoeffValue;
        areRobinConditionsPureNeumann = robinCoeffValue == 0.0f;
        areRobinCoeffsNonnegative = robinCoeffValue >= 0.0f;
        described = true;
        bindCallbacks();
    }
    // constant source / Dirichlet / Robin data
    void setConstant(float sourceValue, float dirichletValue, float robinValue, float robinCoeffValue = 0.0f) {
        std::memset(&description, 0, sizeof(description));
        description.mode = ZB_PDE_CONSTANT;
        description.sourceValue = sourceValue;
        description.dirichletValue = dirichletValue;
        description.robinValue = robinValue;
        description.robinCoeff = robinCoeffValue;
        isSourceConstant = true;
        areRobinConditionsPureNeumann = robinCoeffValue == 0.0f;
        areRobinCoeffsNonnegative = robinCoeffValue >= 0.0f;
        described = true;
        bindCallbacks();
    }
    // constant Dirichlet data for the normal-aligned side of a double-sided absorbing boundary (with setConstant(); the other side
    // keeps dirichletValue): pde.dirichlet(x, true) of walk_on_spheres.h:300-308
    void setDirichletNormalAlignedValue(float value) {
        description.dirichletAlignedValue = value;
        description.hasDirichletAlignedValue = 1;
        bindCallbacks();
    }
    // Backs one callback with sampled data (takes precedence over setManufactured()/setConstant() for that callback; a PDE
    // posed only through grids needs neither). Robin data for the normal-aligned side of a double-sided boundary goes to
    // RobinNormalAligned, Dirichlet data for that side (walk_on_spheres.h:300-308) to DirichletNormalAligned.
    void setGrid(PdeData which, std::shared_ptr<const DataGrid> grid) {
        if (!described) { std::memset(&description, 0, sizeof(description)); description.mode = ZB_PDE_CONSTANT; }
        grids[(int)which] = grid;
        zb_grid_view *views[6] = {&description.sourceGrid, &description.dirichletGrid, &description.robinGrid,
                                  &description.robinAlignedGrid, &description.robinCoeffGrid, &description.dirichletAlignedGrid};
        if (grid) *views[(int)which] = grid->device(); else std::memset(views[(int)which], 0, sizeof(zb_grid_view));
        if (which == PdeData::Source && grid) isSourceConstant = false;
        if (which == PdeData::RobinCoeff && grid) areRobinConditionsPureNeumann = false;
        described = true;
        bindCallbacks();
    }
    bool isDescribed() const { return described; }
    // the descriptor handed to zb_solve, with the public flags folded in
    zb_pde_desc descriptor() const {
        if (!described) throw Error(ZB_ERR_UNSUPPORTED, "zombie_b200::PDE: host callbacks cannot run on the device; "
                                                        "describe the PDE with setManufactured()/setConstant()");
        zb_pde_desc d = description;
        d.absorptionCoeff = absorptionCoeff;
        d.isSourceConstant = isSourceConstant ? 1 : 0;
        d.areRobinConditionsPureNeumann = areRobinConditionsPureNeumann ? 1 : 0;
        d.areRobinCoeffsNonnegative = areRobinCoeffsNonnegative ? 1 : 0;
        return d;
    }

private:
    zb_pde_desc description;
    bool described;
    std::shared_ptr<const DataGrid> grids[6]; // indexed by PdeData

    zb_jet jet(const Vector<DIM>& x) const {
        float xx[3] = {x[0], x[1], DIM == 3 ? x[DIM - 1] : 0.0f};
        return zb_eval_solution(description.family, description.params, xx);
    }
    void bindCallbacks() {
        source = [this](const Vector<DIM>& x) -> T {
            if (grids[(int)PdeData::Source]) return (*grids[(int)PdeData::Source])(x);
            if (description.mode == ZB_PDE_CONSTANT) return description.sourceValue;
            zb_jet j = jet(x);
            return absorptionCoeff*j.u - j.lap;
        };
        dirichlet = [this](const Vector<DIM>& x, bool aligned) -> T {
            if (aligned && grids[(int)PdeData::DirichletNormalAligned]) return (*grids[(int)PdeData::DirichletNormalAligned])(x);
            if (grids[(int)PdeData::Dirichlet]) return (*grids[(int)PdeData::Dirichlet])(x);
            if (description.mode == ZB_PDE_CONSTANT)
                return aligned && description.hasDirichletAlignedValue ? description.dirichletAlignedValue : description.dirichletValue;
            return jet(x).u;
        };
        robin = [this](const Vector<DIM>& x, const Vector<DIM>& n, bool aligned) -> T {
            if (aligned && grids[(int)PdeData::RobinNormalAligned]) return (*grids[(int)PdeData::RobinNormalAligned])(x);
            if (grids[(int)PdeData::Robin]) return (*grids[(int)PdeData::Robin])(x);
            if (description.mode == ZB_PDE_CONSTANT) return description.robinValue;
            zb_jet j = jet(x);
            float dn = 0.0f;
            for (size_t k = 0; k < DIM; k++) dn += n[k]*j.g[k];
            return dn + this->robinCoeff(x, n, aligned)*j.u;
        };
        robinCoeff = [this](const Vector<DIM>& x, const Vector<DIM>&, bool) -> float {
            if (grids[(int)PdeData::RobinCoeff]) return (*grids[(int)PdeData::RobinCoeff])(x);
            return description.robinCoeff;
        };
    }
    void bindHasNonZeroRobinCoeff() {
        if (described) bindCallbacks();
        hasNonZeroRobinCoeff = [this](const Vector<DIM>& x) { // pde.h:60-78
            if (this->areRobinConditionsPureNeumann) return false;
            if (this->robinCoeff) {
                Vector<DIM> n = Vector<DIM>::Zero();
                n(0) = 1.0f;
                return std::fabs(this->robinCoeff(x, n, true)) > 0.0f || std::fabs(this->robinCoeff(x, n, false)) > 0.0f;
            }
            return false;
        };
    }
};

// ------------------------------------------------------------------------------------------------
// settings and per-point data (zombie/point_estimation/common.h)
// ------------------------------------------------------------------------------------------------
struct WalkSettings { // common.h:19-91
    WalkSettings(float epsilonShellForAbsorbingBoundary_, float epsilonShellForReflectingBoundary_,
                 int maxWalkLength_, bool solveDoubleSided_):
        epsilonShellForAbsorbingBoundary(epsilonShellForAbsorbingBoundary_),
        epsilonShellForReflectingBoundary(epsilonShellForReflectingBoundary_),
        silhouettePrecision(1e-3f), russianRouletteThreshold(0.0f),
        splittingThreshold(std::numeric_limits<float>::max()), maxWalkLength(maxWalkLength_),
        stepsBeforeApplyingTikhonov(0), stepsBeforeUsingMaximalSpheres(maxWalkLength_),
        solveDoubleSided(solveDoubleSided_), useGradientControlVariates(true),
        useGradientAntitheticVariates(true), useCosineSamplingForDerivatives(false),
        ignoreAbsorbingBoundaryContribution(false), ignoreReflectingBoundaryContribution(false),
        ignoreSourceContribution(false), printLogs(false) {}
    WalkSettings(float epsilonShellForAbsorbingBoundary_, float epsilonShellForReflectingBoundary_,
                 float silhouettePrecision_, float russianRouletteThreshold_, float splittingThreshold_,
                 int maxWalkLength_, int stepsBeforeApplyingTikhonov_, int stepsBeforeUsingMaximalSpheres_,
                 bool solveDoubleSided_, bool useGradientControlVariates_, bool useGradientAntitheticVariates_,
                 bool useCosineSamplingForDerivatives_, bool ignoreAbsorbingBoundaryContribution_,
                 bool ignoreReflectingBoundaryContribution_, bool ignoreSourceContribution_, bool printLogs_):
        epsilonShellForAbsorbingBoundary(epsilonShellForAbsorbingBoundary_),
        epsilonShellForReflectingBoundary(epsilonShellForReflectingBoundary_),
        silhouettePrecision(silhouettePrecision_), russianRouletteThreshold(russianRouletteThreshold_),
        splittingThreshold(splittingThreshold_), maxWalkLength(maxWalkLength_),
        stepsBeforeApplyingTikhonov(stepsBeforeApplyingTikhonov_),
        stepsBeforeUsingMaximalSpheres(stepsBeforeUsingMaximalSpheres_), solveDoubleSided(solveDoubleSided_),
        useGradientControlVariates(useGradientControlVariates_),
        useGradientAntitheticVariates(useGradientAntitheticVariates_),
        useCosineSamplingForDerivatives(useCosineSamplingForDerivatives_),
        ignoreAbsorbingBoundaryContribution(ignoreAbsorbingBoundaryContribution_),
        ignoreReflectingBoundaryContribution(ignoreReflectingBoundaryContribution_),
        ignoreSourceContribution(ignoreSourceContribution_), printLogs(printLogs_) {}

    float epsilonShellForAbsorbingBoundary;
    float epsilonShellForReflectingBoundary;
    float silhouettePrecision;
    float russianRouletteThreshold;
    float splittingThreshold;
    int maxWalkLength;
    int stepsBeforeApplyingTikhonov;
    int stepsBeforeUsingMaximalSpheres;
    bool solveDoubleSided;
    bool useGradientControlVariates;
    bool useGradientAntitheticVariates;
    bool useCosineSamplingForDerivatives;
    bool ignoreAbsorbingBoundaryContribution;
    bool ignoreReflectingBoundaryContribution;
    bool ignoreSourceContribution;
    bool printLogs;

    zb_walk_settings toAbi() const {
        zb_walk_settings s;
        s.epsilonShellForAbsorbingBoundary = epsilonShellForAbsorbingBoundary;
        s.epsilonShellForReflectingBoundary = epsilonShellForReflectingBoundary;
        s.silhouettePrecision = silhouettePrecision;
        s.russianRouletteThreshold = russianRouletteThreshold;
        s.splittingThreshold = splittingThreshold;
        s.maxWalkLength = maxWalkLength;
        s.stepsBeforeApplyingTikhonov = stepsBeforeApplyingTikhonov;
        s.stepsBeforeUsingMaximalSpheres = stepsBeforeUsingMaximalSpheres;
        s.solveDoubleSided = solveDoubleSided;
        s.useGradientControlVariates = useGradientControlVariates;
        s.useGradientAntitheticVariates = useGradientAntitheticVariates;
        s.useCosineSamplingForDerivatives = useCosineSamplingForDerivatives;
        s.ignoreAbsorbingBoundaryContribution = ignoreAbsorbingBoundaryContribution;
        s.ignoreReflectingBoundaryContribution = ignoreReflectingBoundaryContribution;
        s.ignoreSourceContribution = ignoreSourceContribution;
        s.printLogs = printLogs;
        return s;
    }
};

enum class WalkCompletionCode { ReachedAbsorbingBoundary, TerminatedWithRussianRoulette, ExceededMaxWalkLength, EscapedDomain };
enum class SampleType { InDomain, OnAbsorbingBoundary, OnReflectingBoundary };   // common.h:139-143
enum class EstimationQuantity { Solution, SolutionAndGradient, Skip };          // common.h:145-149

template <typename T, size_t DIM>
struct SamplePoint { // common.h:152-195
    SamplePoint(const Vector<DIM>& pt_, const Vector<DIM>& normal_, SampleType type_,
                EstimationQuantity estimationQuantity_, float pdf_, float distToAbsorbingBoundary_,
                float distToReflectingBoundary_):
        pt(pt_), normal(normal_), type(type_), estimationQuantity(estimationQuantity_), pdf(pdf_),
        distToAbsorbingBoundary(distToAbsorbingBoundary_), distToReflectingBoundary(distToReflectingBoundary_),
        estimateBoundaryNormalAligned(false) {
        directionForDerivative = Vector<DIM>::Zero();
        directionForDerivative(0) = 1.0f;
        reset();
    }
    void reset() {
        rng = seedRng();
        firstSphereRadius = 0.0f;
        robinCoeff = 0.0f;
        solution = T(0.0f);
        normalDerivative = T(0.0f);
        contribution = T(0.0f);
    }

    pcg32 rng;
    Vector<DIM> pt;
    Vector<DIM> normal;
    Vector<DIM> directionForDerivative;
    SampleType type;
    EstimationQuantity estimationQuantity;
    float pdf;
    float distToAbsorbingBoundary;
    float distToReflectingBoundary;
    float firstSphereRadius;
    float robinCoeff;
    T solution, normalDerivative, contribution;
    bool estimateBoundaryNormalAligned;
};

template <typename T, size_t DIM> class WalkOnSpheres;
template <typename T, size_t DIM> class WalkOnStars;
namespace detail { template <typename T, size_t DIM> struct Marshal; }

template <typename T, size_t DIM>
class SampleStatistics { // common.h:201-395
public:
    SampleStatistics() { reset(); }
    void reset() {
        solutionMean = T(0.0f); solutionM2 = T(0.0f);
        for (size_t i = 0; i < DIM; i++) { gradientMean[i] = T(0.0f); gradientM2[i] = T(0.0f); }
        totalFirstSourceContribution = T(0.0f); totalDerivativeContribution = T(0.0f);
        nSolutionEstimates = 0; nGradientEstimates = 0; totalWalkLength = 0; totalSplits = 0;
    }
    void addSolutionEstimate(const T& estimate) { nSolutionEstimates += 1; update(estimate, solutionMean, solutionM2, nSolutionEstimates); }
    void addGradientEstimate(const T *boundaryEstimate, const T *sourceEstimate) {
        nGradientEstimates += 1;
        for (size_t i = 0; i < DIM; i++) update(boundaryEstimate[i] + sourceEstimate[i], gradientMean[i], gradientM2[i], nGradientEstimates);
    }
    void addGradientEstimate(const T *estimate) {
        nGradientEstimates += 1;
        for (size_t i = 0; i < DIM; i++) update(estimate[i], gradientMean[i], gradientM2[i], nGradientEstimates);
    }
    void addFirstSourceContribution(const T& c) { totalFirstSourceContribution += c; }
    void addDerivativeContribution(const T& c) { totalDerivativeContribution += c; }
    void addWalkLength(int length) { totalWalkLength += length; }
    void addSplits(int nSplits) { totalSplits += nSplits; }
    // parallel-variance merge (common.h:268-310); also how shards / repeated solves combine
    void merge(const SampleStatistics<T, DIM>& other) {
        if (other.nSolutionEstimates > 0) {
            int N1 = nSolutionEstimates, N2 = other.nSolutionEstimates, N = N1 + N2;
            if (N1 == 0) { solutionMean = other.solutionMean; solutionM2 = other.solutionM2; }
            else {
                float fN1 = (float)N1, fN2 = (float)N2, fN = (float)N;
                T delta = other.solutionMean - solutionMean;
                solutionMean = solutionMean + delta*(fN2/fN);
                solutionM2 = solutionM2 + other.solutionM2 + delta*delta*(fN1*fN2/fN);
            }
            nSolutionEstimates = N;
        }
        if (other.nGradientEstimates > 0) {
            int N1 = nGradientEstimates, N2 = other.nGradientEstimates, N = N1 + N2;
            float fN1 = (float)N1, fN2 = (float)N2, fN = (float)N;
            for (size_t i = 0; i < DIM; i++) {
                if (N1 == 0) { gradientMean[i] = other.gradientMean[i]; gradientM2[i] = other.gradientM2[i]; }
                else {
                    T delta = other.gradientMean[i] - gradientMean[i];
                    gradientMean[i] = gradientMean[i] + delta*(fN2/fN);
                    gradientM2[i] = gradientM2[i] + other.gradientM2[i] + delta*delta*(fN1*fN2/fN);
                }
            }
            nGradientEstimates = N;
        }
        totalFirstSourceContribution = totalFirstSourceContribution + other.totalFirstSourceContribution;
        totalDerivativeContribution = totalDerivativeContribution + other.totalDerivativeContribution;
        totalWalkLength += other.totalWalkLength;
        totalSplits += other.totalSplits;
    }
    T getEstimatedSolution() const { return solutionMean; }
    T getEstimatedSolutionVariance() const { return solutionM2/(float)std::max(1, nSolutionEstimates - 1); }
    const T *getEstimatedGradient() const { return gradientMean; }
    T getEstimatedGradient(int channel) const { return gradientMean[channel]; }
    std::vector<T> getEstimatedGradientVariance() const {
        int N = std::max(1, nGradientEstimates - 1);
        std::vector<T> variance(DIM);
        for (size_t i = 0; i < DIM; i++) variance[i] = gradientM2[i]/(float)N;
        return variance;
    }
    T getMeanFirstSourceContribution() const { return totalFirstSourceContribution/(float)std::max(1, nSolutionEstimates); }
    T getEstimatedDerivative() const { return totalDerivativeContribution/(float)std::max(1, nSolutionEstimates); }
    int getSolutionEstimateCount() const { return nSolutionEstimates; }
    int getGradientEstimateCount() const { return nGradientEstimates; }
    float getMeanWalkLength() const { return (float)totalWalkLength/(float)std::max(1, nSolutionEstimates); }
    float getMeanSplits() const { return (float)totalSplits/(float)std::max(1, nSolutionEstimates); }

protected:
    void update(const T& estimate, T& mean, T& M2, int N) {
        T delta = estimate - mean;
        mean += delta/(float)N;
        T delta2 = estimate - mean;
        M2 += delta*delta2;
    }
    T solutionMean, solutionM2;
    T gradientMean[DIM], gradientM2[DIM];
    T totalFirstSourceContribution;
    T totalDerivativeContribution;
    int nSolutionEstimates, nGradientEstimates;
    int totalWalkLength, totalSplits;
    friend struct detail::Marshal<T, DIM>;
};

// ------------------------------------------------------------------------------------------------
// geometry (zombie/core/geometric_queries.h, zombie/utils/fcpw_geometric_queries.h)
// ------------------------------------------------------------------------------------------------
template <size_t DIM>
struct IntersectionPoint { // geometric_queries.h:21-35
    IntersectionPoint(): pt(Vector<DIM>::Zero()), normal(Vector<DIM>::Zero()), dist(std::numeric_limits<float>::max()) {}
    IntersectionPoint(const Vector<DIM>& pt_, const Vector<DIM>& normal_, float dist_): pt(pt_), normal(normal_), dist(dist_) {}
    Vector<DIM> pt, normal;
    float dist;
};

template <size_t DIM>
struct BoundarySample { // geometric_queries.h:37-56
    BoundarySample(): pt(Vector<DIM>::Zero()), normal(Vector<DIM>::Zero()), pdf(0.0f) {}
    BoundarySample(const Vector<DIM>& pt_, const Vector<DIM>& normal_, float pdf_): pt(pt_), normal(normal_), pdf(pdf_) {}
    Vector<DIM> pt, normal;
    float pdf;
};

namespace detail {
struct MeshData {
    std::vector<float> positions;
    std::vector<int32_t> indices;
    std::vector<float> minRobin, maxRobin;
    int nVertices = 0, nPrimitives = 0;
    uint64_t version = 0;
};

// the device scene behind one GeometricQueries: both boundary meshes live in one zb_scene
struct SceneBinding {
    int dim = 3;
    std::shared_ptr<MeshData> absorbing, reflecting;
    int reflectingKind = ZB_REFLECTING_NONE;
    bool solveDoubleSided = false;
    float silhouettePrecision = 1e-3f;
    bool watertight = true;
    float domainMin[3], domainMax[3];
    std::shared_ptr<ContextHolder> ctx;
    zb_scene *sceneHandle = nullptr;
    uint64_t builtAbsorbing = ~0ull, builtReflecting = ~0ull;

    ~SceneBinding() { if (sceneHandle) zb_scene_destroy(sceneHandle); }
    void invalidate() { if (sceneHandle) { zb_scene_destroy(sceneHandle); sceneHandle = nullptr; } }
    zb_scene *scene() {
        uint64_t va = absorbing ? absorbing->version : 0, vr = reflecting ? reflecting->version : 0;
        if (sceneHandle && va == builtAbsorbing && vr == builtReflecting) return sceneHandle;
        invalidate();
        if (!ctx) ctx = context();
        const MeshData none;
        const MeshData& a = absorbing ? *absorbing : none;
        const MeshData& r = reflecting ? *reflecting : none;
        check(zb_scene_create(ctx->ctx, dim, a.positions.data(), a.nVertices, a.indices.data(), a.nPrimitives,
                              r.positions.data(), r.nVertices, r.indices.data(), r.nPrimitives, reflectingKind,
                              r.minRobin.empty() ? nullptr : r.minRobin.data(), r.maxRobin.empty() ? nullptr : r.maxRobin.data(),
                              solveDoubleSided ? 1 : 0, silhouettePrecision, watertight ? 1 : 0, domainMin, domainMax,
                              &sceneHandle));
        builtAbsorbing = va;
        builtReflecting = vr;
        return sceneHandle;
    }
};

template <size_t DIM>
inline std::shared_ptr<MeshData> flatten(const std::vector<Vector<DIM>>& positions, const std::vector<Vectori<DIM>>& indices)
{
    static std::atomic<uint64_t> versions{1};
    auto m = std::make_shared<MeshData>();
    m->nVertices = (int)positions.size();
    m->nPrimitives = (int)indices.size();
    m->positions.resize(positions.size()*DIM);
    m->indices.resize(indices.size()*DIM);
    for (size_t i = 0; i < positions.size(); i++) for (size_t k = 0; k < DIM; k++) m->positions[i*DIM + k] = positions[i][k];
    for (size_t i = 0; i < indices.size(); i++) for (size_t k = 0; k < DIM; k++) m->indices[i*DIM + k] = indices[i][k];
    m->version = versions.fetch_add(1);
    return m;
}
} // namespace detail

// Which builder the scenes created afterwards use (zb_context_set_bvh_builder): ZB_BVH_BUILDER_HOST_SAH (default; the
// role of the reference's serial SAH build, fcpw/aggregates/bvh.inl:143-234) or ZB_BVH_BUILDER_DEVICE (on the GPU, in
// milliseconds; for meshes that change often).
inline void setBvhBuilder(int builder) { detail::check(zb_context_set_bvh_builder(detail::context()->ctx, builder)); }

// getIgnoreCandidateSilhouetteCallback (fcpw_geometric_queries.h:1029-1038): the rule is data here
struct IgnoreCandidateSilhouette {
    bool solveDoubleSided;
    float silhouettePrecision;
    bool operator()(float dihedralAngle, int /*index*/) const {
        if (solveDoubleSided) return false;
        return dihedralAngle < silhouettePrecision;
    }
};
inline IgnoreCandidateSilhouette getIgnoreCandidateSilhouetteCallback(bool solveDoubleSided = false, float silhouettePrecision = 1e-3f)
{
    return IgnoreCandidateSilhouette{solveDoubleSided, silhouettePrecision};
}

// getBranchTraversalWeightCallback (fcpw_geometric_queries.h:1391-1398): |G_harmonic(max(r, minRadialDist))|
struct BranchTraversalWeight {
    float minRadialDist;
    float operator()(float r2) const {
        float r = std::max(std::sqrt(r2), minRadialDist);
        return std::fabs(1.0f/(4.0f*3.14159265358979323846f*r));
    }
};
inline BranchTraversalWeight getBranchTraversalWeightCallback(float minRadialDist = 1e-2f) { return BranchTraversalWeight{minRadialDist}; }

template <size_t DIM>
class FcpwDirichletBoundaryHandler { // fcpw_geometric_queries.h:75-88
public:
    FcpwDirichletBoundaryHandler() {}
    void buildAccelerationStructure(const std::vector<Vector<DIM>>& positions, const std::vector<Vectori<DIM>>& indices,
                                    bool /*buildBvh*/ = true, bool /*enableBvhVectorization*/ = false, bool /*printStats*/ = true) {
        mesh = detail::flatten<DIM>(positions, indices);
    }
    std::shared_ptr<detail::MeshData> mesh;
};

template <size_t DIM>
class FcpwNeumannBoundaryHandler { // fcpw_geometric_queries.h:90-105
public:
    FcpwNeumannBoundaryHandler() {}
    void buildAccelerationStructure(const std::vector<Vector<DIM>>& positions, const std::vector<Vectori<DIM>>& indices,
                                    IgnoreCandidateSilhouette ignoreCandidateSilhouette_,
                                    bool /*buildBvh*/ = true, bool /*enableBvhVectorization*/ = false, bool /*printStats*/ = true) {
        mesh = detail::flatten<DIM>(positions, indices);
        ignoreCandidateSilhouette = ignoreCandidateSilhouette_;
    }
    std::shared_ptr<detail::MeshData> mesh;
    IgnoreCandidateSilhouette ignoreCandidateSilhouette{false, 1e-3f};
};

template <size_t DIM>
class FcpwRobinBoundaryHandler { // fcpw_geometric_queries.h:107-133
public:
    FcpwRobinBoundaryHandler() {}
    void buildAccelerationStructure(const std::vector<Vector<DIM>>& positions, const std::vector<Vectori<DIM>>& indices,
                                    IgnoreCandidateSilhouette ignoreCandidateSilhouette_,
                                    const std::vector<float>& minRobinCoeffValues, const std::vector<float>& maxRobinCoeffValues,
                                    bool /*buildBvh*/ = true, bool /*enableBvhVectorization*/ = false, bool /*printStats*/ = true) {
        if (minRobinCoeffValues.size() != indices.size() || maxRobinCoeffValues.size() != indices.size())
            throw Error(ZB_ERR_INVALID_ARGUMENT, "FcpwRobinBoundaryHandler: one min and max Robin coefficient per primitive is required");
        mesh = detail::flatten<DIM>(positions, indices);
        mesh->minRobin = minRobinCoeffValues;
        mesh->maxRobin = maxRobinCoeffValues;
        ignoreCandidateSilhouette = ignoreCandidateSilhouette_;
    }
    // fcpw_geometric_queries.h:997-1010
    void updateCoefficientValues(const std::vector<float>& minRobinCoeffValues, const std::vector<float>& maxRobinCoeffValues) {
        if (!mesh || (int)minRobinCoeffValues.size() != mesh->nPrimitives || (int)maxRobinCoeffValues.size() != mesh->nPrimitives)
            throw Error(ZB_ERR_INVALID_ARGUMENT, "FcpwRobinBoundaryHandler::updateCoefficientValues: size mismatch");
        mesh->minRobin = minRobinCoeffValues;
        mesh->maxRobin = maxRobinCoeffValues;
        for (auto& weak : bindings) {
            if (auto b = weak.lock()) {
                if (b->sceneHandle) detail::check(zb_scene_update_robin_coefficients(b->sceneHandle, mesh->minRobin.data(), mesh->maxRobin.data()));
            }
        }
    }
    std::shared_ptr<detail::MeshData> mesh;
    IgnoreCandidateSilhouette ignoreCandidateSilhouette{false, 1e-3f};
    std::vector<std::weak_ptr<detail::SceneBinding>> bindings;
};

template <size_t DIM>
struct GeometricQueries { // geometric_queries.h:59-124
    GeometricQueries(): hasNonEmptyAbsorbingBoundary(false), hasNonEmptyReflectingBoundary(false), domainIsWatertight(true),
        domainMin(Vector<DIM>::Constant(std::numeric_limits<float>::lowest())),
        domainMax(Vector<DIM>::Constant(std::numeric_limits<float>::max())) { init(); }
    GeometricQueries(bool domainIsWatertight_, const Vector<DIM>& domainMin_, const Vector<DIM>& domainMax_):
        hasNonEmptyAbsorbingBoundary(false), hasNonEmptyReflectingBoundary(false), domainIsWatertight(domainIsWatertight_),
        domainMin(domainMin_), domainMax(domainMax_) { init(); }

    bool hasNonEmptyAbsorbingBoundary;
    bool hasNonEmptyReflectingBoundary;
    bool domainIsWatertight;
    Vector<DIM> domainMin, domainMax;

    // The single-query callbacks of the reference, each one batched query of size 1 through the C ABI (for
    // user code that prepares sample points; the walks themselves never call back to the host).
    std::function<float(const Vector<DIM>&, bool)> computeDistToAbsorbingBoundary;
    std::function<float(const Vector<DIM>&, bool)> computeDistToReflectingBoundary;
    std::function<float(const Vector<DIM>&, bool)> computeDistToBoundary;
    std::function<bool(Vector<DIM>&, Vector<DIM>&, float&, bool)> projectToAbsorbingBoundary;
    std::function<bool(Vector<DIM>&, Vector<DIM>&, float&, bool)> projectToReflectingBoundary;
    std::function<bool(const Vector<DIM>&, const Vector<DIM>&, const Vector<DIM>&, float, bool, IntersectionPoint<DIM>&)> intersectReflectingBoundary;
    std::function<bool(const Vector<DIM>&, const Vector<DIM>&, const Vector<DIM>&, const Vector<DIM>&, bool, bool)> intersectsWithReflectingBoundary;
    std::function<bool(const Vector<DIM>&, float, const Vector<DIM>&, BoundarySample<DIM>&)> sampleReflectingBoundary;
    std::function<float(const Vector<DIM>&, float, float, float, bool)> computeStarRadiusForReflectingBoundary;
    std::function<bool(const Vector<DIM>&)> insideDomain;
    std::function<bool(const Vector<DIM>&)> insideBoundingDomain;
    std::function<bool(const Vector<DIM>&)> outsideBoundingDomain;

    // batched insideDomain (zb_inside_domain): the inside test of a domain sampler over many candidates at once
    void insideDomain_batch(const std::vector<Vector<DIM>>& x, std::vector<uint8_t>& inside) const {
        inside.assign(x.size(), 1);
        if (!domainIsWatertight || x.empty()) return;
        if (!(hasNonEmptyAbsorbingBoundary || hasNonEmptyReflectingBoundary)) { inside.assign(x.size(), 0); return; } // no crossing: even count
        std::vector<float> flat(x.size()*DIM);
        for (size_t i = 0; i < x.size(); i++) for (size_t k = 0; k < DIM; k++) flat[i*DIM + k] = x[i][k];
        detail::check(zb_inside_domain(scene(), (int)x.size(), flat.data(), inside.data()));
    }

    // batched forms (what a caller preparing 10^6 sample points should use): unsigned distances
    void computeDistToAbsorbingBoundary_batch(const std::vector<Vector<DIM>>& x, std::vector<float>& dist) const { distances(ZB_ABSORBING, x, dist); }
    void computeDistToReflectingBoundary_batch(const std::vector<Vector<DIM>>& x, std::vector<float>& dist) const { distances(ZB_REFLECTING, x, dist); }

    // device scene shared by every solver constructed from these queries
    std::shared_ptr<detail::SceneBinding> binding;
    zb_scene *scene() const {
        for (size_t k = 0; k < DIM; k++) { binding->domainMin[k] = domainMin[k]; binding->domainMax[k] = domainMax[k]; }
        binding->watertight = domainIsWatertight;
        return binding->scene();
    }

private:
    void distances(int which, const std::vector<Vector<DIM>>& x, std::vector<float>& dist) const {
        dist.assign(x.size(), std::numeric_limits<float>::max());
        bool present = which == ZB_ABSORBING ? hasNonEmptyAbsorbingBoundary : hasNonEmptyReflectingBoundary;
        if (!present || x.empty()) return;
        std::vector<float> flat(x.size()*DIM);
        for (size_t i = 0; i < x.size(); i++) for (size_t k = 0; k < DIM; k++) flat[i*DIM + k] = x[i][k];
        std::vector<int32_t> prim(x.size());
        detail::check(zb_closest_point(scene(), which, (int)x.size(), flat.data(), nullptr, prim.data(), dist.data(), nullptr, nullptr));
    }
    float distance1(int which, const Vector<DIM>& x) const {
        float xx[DIM], d = std::numeric_limits<float>::max();
        int32_t prim;
        for (size_t k = 0; k < DIM; k++) xx[k] = x[k];
        detail::check(zb_closest_point(scene(), which, 1, xx, nullptr, &prim, &d, nullptr, nullptr));
        return d;
    }
    bool project1(int which, Vector<DIM>& x, Vector<DIM>& normal, float& dist) const {
        float xx[DIM], p[DIM], n[DIM], d;
        int32_t prim;
        for (size_t k = 0; k < DIM; k++) xx[k] = x[k];
        detail::check(zb_closest_point(scene(), which, 1, xx, nullptr, &prim, &d, p, n));
        if (prim < 0) return false;
        for (size_t k = 0; k < DIM; k++) { x[k] = p[k]; normal[k] = n[k]; }
        dist = d;
        return true;
    }
    void init() {
        binding = std::make_shared<detail::SceneBinding>();
        binding->dim = (int)DIM;
        // defaults of GeometricQueries::populate (geometric_queries.h:210-390) for the members kept here
        computeDistToAbsorbingBoundary = [this](const Vector<DIM>& x, bool) -> float {
            if (hasNonEmptyAbsorbingBoundary) return distance1(ZB_ABSORBING, x);
            float s = 0.0f; // distance to the bounding box, as the reference's default (geometric_queries.h:213-218)
            for (size_t k = 0; k < DIM; k++) { float m = std::min(domainMin[k] - x[k], x[k] - domainMax[k]); s += m*m; }
            return std::sqrt(s);
        };
        computeDistToReflectingBoundary = [this](const Vector<DIM>& x, bool) -> float {
            return hasNonEmptyReflectingBoundary ? distance1(ZB_REFLECTING, x) : std::numeric_limits<float>::max();
        };
        computeDistToBoundary = [this](const Vector<DIM>& x, bool s) -> float {
            return std::min(computeDistToAbsorbingBoundary(x, s), computeDistToReflectingBoundary(x, s));
        };
        projectToAbsorbingBoundary = [this](Vector<DIM>& x, Vector<DIM>& n, float& d, bool) -> bool {
            return hasNonEmptyAbsorbingBoundary && project1(ZB_ABSORBING, x, n, d);
        };
        projectToReflectingBoundary = [this](Vector<DIM>& x, Vector<DIM>& n, float& d, bool) -> bool {
            return hasNonEmptyReflectingBoundary && project1(ZB_REFLECTING, x, n, d);
        };
        intersectReflectingBoundary = [this](const Vector<DIM>& origin, const Vector<DIM>& /*normal*/, const Vector<DIM>& dir,
                                             float tMax, bool /*onReflectingBoundary*/, IntersectionPoint<DIM>& out) -> bool {
            if (!hasNonEmptyReflectingBoundary) return false;
            float o[DIM], d[DIM], p[DIM], n[DIM], t;
            int32_t prim;
            for (size_t k = 0; k < DIM; k++) { o[k] = origin[k]; d[k] = dir[k]; }
            detail::check(zb_intersect_ray(scene(), ZB_REFLECTING, 1, o, d, &tMax, &prim, &t, p, n));
            if (prim < 0) return false;
            for (size_t k = 0; k < DIM; k++) { out.pt[k] = p[k]; out.normal[k] = n[k]; }
            out.dist = t;
            return true;
        };
        intersectsWithReflectingBoundary = [this](const Vector<DIM>& xi, const Vector<DIM>& xj, const Vector<DIM>&, const Vector<DIM>&,
                                                  bool, bool) -> bool {
            if (!hasNonEmptyReflectingBoundary) return false;
            float a[DIM], b[DIM];
            uint8_t occluded = 0;
            for (size_t k = 0; k < DIM; k++) { a[k] = xi[k]; b[k] = xj[k]; }
            detail::check(zb_occluded(scene(), ZB_REFLECTING, 1, a, b, &occluded));
            return occluded != 0;
        };
        sampleReflectingBoundary = [this](const Vector<DIM>& x, float radius, const Vector<DIM>& randNums, BoundarySample<DIM>& out) -> bool {
            if (!hasNonEmptyReflectingBoundary) return false;
            float xx[DIM], u[DIM], p[DIM], n[DIM], pdf;
            int32_t prim;
            for (size_t k = 0; k < DIM; k++) { xx[k] = x[k]; u[k] = randNums[k]; }
            detail::check(zb_sample_sphere(scene(), ZB_REFLECTING, 1, xx, &radius, u, &prim, p, n, &pdf));
            if (prim < 0) return false;
            for (size_t k = 0; k < DIM; k++) { out.pt[k] = p[k]; out.normal[k] = n[k]; }
            out.pdf = pdf;
            return true;
        };
        computeStarRadiusForReflectingBoundary = [this](const Vector<DIM>& x, float minRadius, float maxRadius, float silhouettePrecision,
                                                        bool flipNormalOrientation) -> float {
            if (!hasNonEmptyReflectingBoundary) return maxRadius;
            float xx[DIM], r;
            uint8_t flip = flipNormalOrientation ? 1 : 0;
            for (size_t k = 0; k < DIM; k++) xx[k] = x[k];
            detail::check(zb_star_radius(scene(), 1, xx, &minRadius, &maxRadius, silhouettePrecision, &flip, &r));
            return r;
        };
        insideDomain = [this](const Vector<DIM>& x) -> bool { // geometric_queries.h:357-371
            if (!domainIsWatertight) return true;
            if (!(hasNonEmptyAbsorbingBoundary || hasNonEmptyReflectingBoundary)) return false; // no crossing: even count
            float xx[DIM];
            uint8_t in = 0;
            for (size_t k = 0; k < DIM; k++) xx[k] = x[k];
            detail::check(zb_inside_domain(scene(), 1, xx, &in));
            return in != 0;
        };
        insideBoundingDomain = [this](const Vector<DIM>& x) -> bool {
            for (size_t k = 0; k < DIM; k++) if (x[k] < domainMin[k] || x[k] > domainMax[k]) return false;
            return true;
        };
        outsideBoundingDomain = [this](const Vector<DIM>& x) -> bool { return !insideBoundingDomain(x); };
    }
    // the callbacks capture `this`: keep the object where it was populated (the reference has the same
    // constraint for its `[this]` lambdas, geometric_queries.h:210-390)
    GeometricQueries(const GeometricQueries&) = delete;
    GeometricQueries& operator=(const GeometricQueries&) = delete;
};

// fcpw_geometric_queries.h:139-150
template <size_t DIM>
void populateGeometricQueriesForDirichletBoundary(FcpwDirichletBoundaryHandler<DIM>& handler, GeometricQueries<DIM>& queries)
{
    queries.binding->absorbing = handler.mesh;
    queries.hasNonEmptyAbsorbingBoundary = handler.mesh && handler.mesh->nPrimitives > 0;
}
template <size_t DIM>
void populateGeometricQueriesForNeumannBoundary(FcpwNeumannBoundaryHandler<DIM>& handler, BranchTraversalWeight weight,
                                                GeometricQueries<DIM>& queries)
{
    if (weight.minRadialDist != 1e-2f) throw Error(ZB_ERR_UNSUPPORTED, "branch traversal weight: only minRadialDist = 1e-2 is built into the device traversal");
    queries.binding->reflecting = handler.mesh;
    queries.binding->reflectingKind = ZB_REFLECTING_NEUMANN;
    queries.binding->solveDoubleSided = handler.ignoreCandidateSilhouette.solveDoubleSided;
    queries.binding->silhouettePrecision = handler.ignoreCandidateSilhouette.silhouettePrecision;
    queries.hasNonEmptyReflectingBoundary = handler.mesh && handler.mesh->nPrimitives > 0;
}
template <size_t DIM>
void populateGeometricQueriesForRobinBoundary(FcpwRobinBoundaryHandler<DIM>& handler, BranchTraversalWeight weight,
                                              GeometricQueries<DIM>& queries)
{
    if (weight.minRadialDist != 1e-2f) throw Error(ZB_ERR_UNSUPPORTED, "branch traversal weight: only minRadialDist = 1e-2 is built into the device traversal");
    queries.binding->reflecting = handler.mesh;
    queries.binding->reflectingKind = ZB_REFLECTING_ROBIN;
    queries.binding->solveDoubleSided = handler.ignoreCandidateSilhouette.solveDoubleSided;
    queries.binding->silhouettePrecision = handler.ignoreCandidateSilhouette.silhouettePrecision;
    queries.hasNonEmptyReflectingBoundary = handler.mesh && handler.mesh->nPrimitives > 0;
    handler.bindings.push_back(queries.binding);
}

// ------------------------------------------------------------------------------------------------
// estimators
// ------------------------------------------------------------------------------------------------
namespace detail {
// AoS (the reference's vectors of structs) <-> SoA (the C ABI)
template <typename T, size_t DIM>
struct Marshal {
    std::vector<float> pt, normal, dir, distA, distR, firstSphereRadius;
    std::vector<int32_t> type, quantity, nWalks;
    std::vector<uint8_t> aligned;
    std::vector<uint64_t> rngState, rngInc;
    std::vector<float> mean, m2, gmean, gm2, firstSource, derivative;
    std::vector<int32_t> nSol, nGrad, walkLength, splits;

    void gather(const std::vector<int>& walks, const std::vector<SamplePoint<T, DIM>>& pts, const std::vector<SampleStatistics<T, DIM>>& st) {
        size_t n = pts.size();
        pt.resize(n*DIM); normal.resize(n*DIM); dir.resize(n*DIM); distA.resize(n); distR.resize(n); firstSphereRadius.resize(n);
        type.resize(n); quantity.resize(n); nWalks.resize(n); aligned.resize(n); rngState.resize(n); rngInc.resize(n);
        mean.resize(n); m2.resize(n); gmean.resize(n*DIM); gm2.resize(n*DIM); firstSource.resize(n); derivative.resize(n);
        nSol.resize(n); nGrad.resize(n); walkLength.resize(n); splits.resize(n);
        for (size_t i = 0; i < n; i++) {
            const SamplePoint<T, DIM>& p = pts[i];
            const SampleStatistics<T, DIM>& s = st[i];
            for (size_t k = 0; k < DIM; k++) {
                pt[i*DIM + k] = p.pt[k]; normal[i*DIM + k] = p.normal[k]; dir[i*DIM + k] = p.directionForDerivative[k];
                gmean[i*DIM + k] = s.gradientMean[k]; gm2[i*DIM + k] = s.gradientM2[k];
            }
            distA[i] = p.distToAbsorbingBoundary; distR[i] = p.distToReflectingBoundary; firstSphereRadius[i] = p.firstSphereRadius;
            type[i] = (int32_t)p.type; quantity[i] = (int32_t)p.estimationQuantity; nWalks[i] = walks[i];
            aligned[i] = p.estimateBoundaryNormalAligned ? 1 : 0;
            rngState[i] = p.rng.state; rngInc[i] = p.rng.inc;
            mean[i] = s.solutionMean; m2[i] = s.solutionM2; firstSource[i] = s.totalFirstSourceContribution;
            derivative[i] = s.totalDerivativeContribution; nSol[i] = s.nSolutionEstimates; nGrad[i] = s.nGradientEstimates;
            walkLength[i] = s.totalWalkLength; splits[i] = s.totalSplits;
        }
    }
    zb_sample_points points() {
        zb_sample_points p;
        p.n = (int32_t)distA.size();
        p.pt = pt.data(); p.normal = normal.data(); p.directionForDerivative = dir.data();
        p.type = type.data(); p.estimationQuantity = quantity.data();
        p.distToAbsorbingBoundary = distA.data(); p.distToReflectingBoundary = distR.data();
        p.estimateBoundaryNormalAligned = aligned.data(); p.nWalks = nWalks.data();
        p.rngState = rngState.data(); p.rngInc = rngInc.data(); p.firstSphereRadius = firstSphereRadius.data();
        return p;
    }
    zb_statistics statistics() {
        zb_statistics s;
        s.solutionMean = mean.data(); s.solutionM2 = m2.data(); s.gradientMean = gmean.data(); s.gradientM2 = gm2.data();
        s.totalFirstSourceContribution = firstSource.data(); s.totalDerivativeContribution = derivative.data();
        s.nSolutionEstimates = nSol.data(); s.nGradientEstimates = nGrad.data();
        s.totalWalkLength = walkLength.data(); s.totalSplits = splits.data();
        return s;
    }
    void scatter(std::vector<SamplePoint<T, DIM>>& pts, std::vector<SampleStatistics<T, DIM>>& st) const {
        for (size_t i = 0; i < pts.size(); i++) {
            pts[i].rng.state = rngState[i];
            pts[i].firstSphereRadius = firstSphereRadius[i];
            SampleStatistics<T, DIM>& s = st[i];
            s.solutionMean = mean[i]; s.solutionM2 = m2[i];
            for (size_t k = 0; k < DIM; k++) { s.gradientMean[k] = gmean[i*DIM + k]; s.gradientM2[k] = gm2[i*DIM + k]; }
            s.totalFirstSourceContribution = firstSource[i]; s.totalDerivativeContribution = derivative[i];
            s.nSolutionEstimates = nSol[i]; s.nGradientEstimates = nGrad[i];
            s.totalWalkLength = walkLength[i]; s.totalSplits = splits[i];
        }
    }
};

template <typename T, size_t DIM>
int64_t solveVector(int estimator, const GeometricQueries<DIM>& queries, const PDE<T, DIM>& pde, const WalkSettings& settings,
                    const std::vector<int>& nWalks, std::vector<SamplePoint<T, DIM>>& samplePts,
                    std::vector<SampleStatistics<T, DIM>>& statistics, const std::function<void(int, int)>& reportProgress)
{
    int nPoints = (int)samplePts.size();
    if ((int)nWalks.size() != nPoints) throw Error(ZB_ERR_INVALID_ARGUMENT, "solve: nWalks and samplePts differ in length");
    if ((int)statistics.size() != nPoints) { // walk_on_stars.h:160-164
        statistics.clear();
        statistics.resize(nPoints);
    }
    if (nPoints == 0) return 0;
    zb_pde_desc desc = pde.descriptor();
    zb_walk_settings ws = settings.toAbi();
    Marshal<T, DIM> m;
    m.gather(nWalks, samplePts, statistics);
    zb_sample_points p = m.points();
    zb_statistics s = m.statistics();
    int64_t steps = 0;
    check(zb_solve(queries.scene(), estimator, &desc, &ws, &p, &s, &steps));
    m.scatter(samplePts, statistics);
    if (reportProgress) reportProgress(nPoints, 0);
    return steps;
}
} // namespace detail

// Vector-valued problems — the reference's T = Array<float, C>, the `channels` of its bindings (python/zombie_bindings.cpp:37-44;
// core/pde.h:26-44: source / dirichlet / robin return T, absorptionCoeff and robinCoeff stay floats): one scalar PDE description
// and one statistics vector per channel, ONE set of walks (zb_solve_channels). pdes must agree in everything that steers the
// walk (absorptionCoeff, Robin coefficient, the three flags); samplePts are shared and their generators advance once.
// Channel c of the result equals solve(pdes[c], ...) on the same generators up to the contraction of multiply-adds.
template <size_t DIM>
inline int64_t solveChannels(int estimator, const GeometricQueries<DIM>& queries, const std::vector<PDE<float, DIM>>& pdes,
                             const WalkSettings& settings, const std::vector<int>& nWalks, std::vector<SamplePoint<float, DIM>>& samplePts,
                             std::vector<std::vector<SampleStatistics<float, DIM>>>& statistics)
{
    const size_t nChannels = pdes.size(), nPoints = samplePts.size();
    if (nChannels == 0) throw Error(ZB_ERR_INVALID_ARGUMENT, "solveChannels: no PDE given");
    if (nWalks.size() != nPoints) throw Error(ZB_ERR_INVALID_ARGUMENT, "solveChannels: nWalks and samplePts differ in length");
    statistics.resize(nChannels);
    for (auto& s : statistics) if (s.size() != nPoints) { s.clear(); s.resize(nPoints); }
    if (nPoints == 0) return 0;
    std::vector<zb_pde_desc> descs;
    for (const auto& p : pdes) descs.push_back(p.descriptor());
    zb_walk_settings ws = settings.toAbi();
    std::vector<detail::Marshal<float, DIM>> m(nChannels);
    std::vector<zb_statistics> st(nChannels);
    for (size_t c = 0; c < nChannels; c++) { m[c].gather(nWalks, samplePts, statistics[c]); st[c] = m[c].statistics(); }
    zb_sample_points p = m[0].points();
    int64_t steps = 0;
    detail::check(zb_solve_channels(queries.scene(), estimator, (int)nChannels, descs.data(), &ws, &p, st.data(), &steps));
    for (size_t c = nChannels; c-- > 0;) {                    // channel 0 last: it carries the advanced generators
        std::vector<SamplePoint<float, DIM>> scratch = samplePts;
        m[c].scatter(c == 0 ? samplePts : scratch, statistics[c]);
    }
    return steps;
}

template <typename T, size_t DIM>
class WalkOnSpheres { // walk_on_spheres.h:26-44
public:
    WalkOnSpheres(const GeometricQueries<DIM>& queries_): queries(queries_) {}
    void solve(const PDE<T, DIM>& pde, const WalkSettings& walkSettings, int nWalks, SamplePoint<T, DIM>& samplePt,
               SampleStatistics<T, DIM>& statistics) const {
        std::vector<int> w(1, nWalks);
        std::vector<SamplePoint<T, DIM>> p(1, samplePt);
        std::vector<SampleStatistics<T, DIM>> s(1, statistics);
        solve(pde, walkSettings, w, p, s);
        samplePt = p[0];
        statistics = s[0];
    }
    void solve(const PDE<T, DIM>& pde, const WalkSettings& walkSettings, const std::vector<int>& nWalks,
               std::vector<SamplePoint<T, DIM>>& samplePts, std::vector<SampleStatistics<T, DIM>>& statistics,
               bool /*runSingleThreaded*/ = false, std::function<void(int, int)> reportProgress = {}) const {
        lastWalkSteps = detail::solveVector<T, DIM>(ZB_WALK_ON_SPHERES, queries, pde, walkSettings, nWalks, samplePts, statistics, reportProgress);
    }
    // walk steps (iterations of walk()'s while loop) taken by the last solve(): numerator of the steps/s metric
    mutable int64_t lastWalkSteps = 0;
protected:
    const GeometricQueries<DIM>& queries;
};

template <typename T, size_t DIM>
class WalkOnStars { // walk_on_stars.h:25-46
public:
    WalkOnStars(const GeometricQueries<DIM>& queries_): queries(queries_) {}
    void solve(const PDE<T, DIM>& pde, const WalkSettings& walkSettings, int nWalks, SamplePoint<T, DIM>& samplePt,
               SampleStatistics<T, DIM>& statistics) const {
        std::vector<int> w(1, nWalks);
        std::vector<SamplePoint<T, DIM>> p(1, samplePt);
        std::vector<SampleStatistics<T, DIM>> s(1, statistics);
        solve(pde, walkSettings, w, p, s);
        samplePt = p[0];
        statistics = s[0];
    }
    void solve(const PDE<T, DIM>& pde, const WalkSettings& walkSettings, const std::vector<int>& nWalks,
               std::vector<SamplePoint<T, DIM>>& samplePts, std::vector<SampleStatistics<T, DIM>>& statistics,
               bool /*runSingleThreaded*/ = false, std::function<void(int, int)> reportProgress = {}) const {
        lastWalkSteps = detail::solveVector<T, DIM>(ZB_WALK_ON_STARS, queries, pde, walkSettings, nWalks, samplePts, statistics, reportProgress);
    }
    mutable int64_t lastWalkSteps = 0;
protected:
    const GeometricQueries<DIM>& queries;
};

// ------------------------------------------------------------------------------------------------
// Boundary Value Caching: the splat (zombie/variance_reduction/boundary_value_caching.h)
// ------------------------------------------------------------------------------------------------
namespace bvc {

template <typename T, size_t DIM> class BoundaryValueCaching;

template <typename T, size_t DIM>
struct EvaluationPoint { // boundary_value_caching.h:18-56. The five SampleStatistics are kept as (count, sums): a Welford mean is sum/N
    EvaluationPoint(const Vector<DIM>& pt_, const Vector<DIM>& normal_, SampleType type_, float distToAbsorbingBoundary_,
                    float distToReflectingBoundary_):
        pt(pt_), normal(normal_), type(type_), distToAbsorbingBoundary(distToAbsorbingBoundary_),
        distToReflectingBoundary(distToReflectingBoundary_) { reset(); }
    T getEstimatedSolution() const { T s = T(0.0f); for (int b = 0; b < ZB_BVC_BUCKETS; b++) s += mean(b, 0); return s; }
    void getEstimatedGradient(std::vector<T>& gradient) const {
        gradient.resize(DIM);
        for (size_t k = 0; k < DIM; k++) gradient[k] = getEstimatedGradient((int)k);
    }
    T getEstimatedGradient(int channel) const { T s = T(0.0f); for (int b = 0; b < ZB_BVC_BUCKETS; b++) s += mean(b, 1 + channel); return s; }
    void reset() { for (int b = 0; b < ZB_BVC_BUCKETS; b++) { count[b] = 0; for (size_t k = 0; k < 1 + DIM; k++) sums[b][k] = 0.0f; } }

    Vector<DIM> pt;
    Vector<DIM> normal;
    SampleType type;
    float distToAbsorbingBoundary;
    float distToReflectingBoundary;

protected:
    T mean(int b, int k) const { return count[b] > 0 ? sums[b][k]/(float)count[b] : T(0.0f); }
    int32_t count[ZB_BVC_BUCKETS];
    float sums[ZB_BVC_BUCKETS][1 + DIM];
    template <typename A, size_t B> friend class BoundaryValueCaching;
};

template <typename T, size_t DIM>
class BoundaryValueCaching { // boundary_value_caching.h:58-193 (the splat overloads)
public:
    BoundaryValueCaching(const GeometricQueries<DIM>& queries_, const WalkOnStars<T, DIM>& walkOnStars_):
        queries(queries_), walkOnStars(walkOnStars_) {}

    // splats sample pt data to the input evaluation pts (:482-502): one zb_bvc_splat
    void splat(const PDE<T, DIM>& pde, const std::vector<SamplePoint<T, DIM>>& samplePts, float radiusClamp,
               float kernelRegularization, float robinCoeffCutoffForNormalDerivative, float cutoffDistToAbsorbingBoundary,
               float cutoffDistToReflectingBoundary, std::vector<EvaluationPoint<T, DIM>>& evalPts,
               std::function<void(int, int)> reportProgress = {}) const {
        size_t n = samplePts.size(), m = evalPts.size();
        if (n == 0 || m == 0) return;
        std::vector<float> spt(n*DIM), snormal(n*DIM), pdf(n), solution(n), normalDerivative(n), contribution(n), robinCoeff(n);
        std::vector<int32_t> stype(n);
        std::vector<uint8_t> aligned(n);
        for (size_t i = 0; i < n; i++) {
            const SamplePoint<T, DIM>& p = samplePts[i];
            for (size_t k = 0; k < DIM; k++) { spt[i*DIM + k] = p.pt[k]; snormal[i*DIM + k] = p.normal[k]; }
            stype[i] = (int32_t)p.type; aligned[i] = p.estimateBoundaryNormalAligned ? 1 : 0;
            pdf[i] = p.pdf; solution[i] = p.solution; normalDerivative[i] = p.normalDerivative; contribution[i] = p.contribution;
            robinCoeff[i] = p.robinCoeff;
        }
        zb_bvc_samples s{(int32_t)n, spt.data(), snormal.data(), stype.data(), aligned.data(), pdf.data(), solution.data(),
                         normalDerivative.data(), contribution.data(), robinCoeff.data()};
        std::vector<float> ept(m*DIM), da(m), dr(m), bucketSum(m*ZB_BVC_BUCKETS*(1 + DIM)), sol(m), grad(m*DIM);
        std::vector<int32_t> etype(m), bucketCount(m*ZB_BVC_BUCKETS);
        for (size_t i = 0; i < m; i++) {
            const EvaluationPoint<T, DIM>& e = evalPts[i];
            for (size_t k = 0; k < DIM; k++) ept[i*DIM + k] = e.pt[k];
            etype[i] = (int32_t)e.type; da[i] = e.distToAbsorbingBoundary; dr[i] = e.distToReflectingBoundary;
            for (int b = 0; b < ZB_BVC_BUCKETS; b++) {
                bucketCount[i*ZB_BVC_BUCKETS + b] = e.count[b];
                for (size_t k = 0; k < 1 + DIM; k++) bucketSum[(i*ZB_BVC_BUCKETS + b)*(1 + DIM) + k] = e.sums[b][k];
            }
        }
        zb_bvc_eval_points ev{(int32_t)m, ept.data(), etype.data(), da.data(), dr.data(), bucketCount.data(), bucketSum.data(),
                              sol.data(), grad.data()};
        detail::check(zb_bvc_splat(detail::context()->ctx, (int)DIM, pde.absorptionCoeff, &s, radiusClamp, kernelRegularization,
                                   robinCoeffCutoffForNormalDerivative, cutoffDistToAbsorbingBoundary, cutoffDistToReflectingBoundary,
                                   &ev, nullptr));
        for (size_t i = 0; i < m; i++) {
            EvaluationPoint<T, DIM>& e = evalPts[i];
            for (int b = 0; b < ZB_BVC_BUCKETS; b++) {
                e.count[b] = bucketCount[i*ZB_BVC_BUCKETS + b];
                for (size_t k = 0; k < 1 + DIM; k++) e.sums[b][k] = bucketSum[(i*ZB_BVC_BUCKETS + b)*(1 + DIM) + k];
            }
        }
        if (reportProgress) reportProgress((int)n, 0);
    }

protected:
    const GeometricQueries<DIM>& queries;
    const WalkOnStars<T, DIM>& walkOnStars;
};

} // namespace bvc

// ------------------------------------------------------------------------------------------------
// Reverse Walk on Stars (zombie/variance_reduction/reverse_walk_solver.h)
// ------------------------------------------------------------------------------------------------
namespace rws {

template <typename T, size_t DIM>
struct EvaluationPoint { // reverse_walk_solver.h:20-52 (no mutex: the device gather needs none)
    EvaluationPoint(const Vector<DIM>& pt_, const Vector<DIM>& normal_, SampleType type_, float distToAbsorbingBoundary_,
                    float distToReflectingBoundary_):
        pt(pt_), normal(normal_), type(type_), distToAbsorbingBoundary(distToAbsorbingBoundary_),
        distToReflectingBoundary(distToReflectingBoundary_) { reset(); }
    T getEstimatedSolution(int nAbsorbingBoundarySamples, int nAbsorbingBoundaryNormalAlignedSamples, int nReflectingBoundarySamples,
                           int nReflectingBoundaryNormalAlignedSamples, int nSourceSamples) const { // :129-166
        if (type == SampleType::OnAbsorbingBoundary) return totalAbsorbingBoundaryContribution;
        T solution(0.0f);
        if (nAbsorbingBoundarySamples > 0) {
            if (totalPoissonKernelContribution > 0.0f) solution += totalAbsorbingBoundaryContribution/totalPoissonKernelContribution;
            else solution += totalAbsorbingBoundaryContribution/(float)nAbsorbingBoundarySamples;
        }
        if (nAbsorbingBoundaryNormalAlignedSamples > 0) solution += totalAbsorbingBoundaryNormalAlignedContribution/(float)nAbsorbingBoundaryNormalAlignedSamples;
        if (nReflectingBoundarySamples > 0) solution += totalReflectingBoundaryContribution/(float)nReflectingBoundarySamples;
        if (nReflectingBoundaryNormalAlignedSamples > 0) solution += totalReflectingBoundaryNormalAlignedContribution/(float)nReflectingBoundaryNormalAlignedSamples;
        if (nSourceSamples > 0) solution += totalSourceContribution/(float)nSourceSamples;
        return solution;
    }
    void reset() {
        totalPoissonKernelContribution = 0.0f;
        totalAbsorbingBoundaryContribution = totalAbsorbingBoundaryNormalAlignedContribution = T(0.0f);
        totalReflectingBoundaryContribution = totalReflectingBoundaryNormalAlignedContribution = totalSourceContribution = T(0.0f);
    }

    Vector<DIM> pt;
    Vector<DIM> normal;
    SampleType type;
    float distToAbsorbingBoundary;
    float distToReflectingBoundary;
    float totalPoissonKernelContribution;
    T totalAbsorbingBoundaryContribution;
    T totalAbsorbingBoundaryNormalAlignedContribution;
    T totalReflectingBoundaryContribution;
    T totalReflectingBoundaryNormalAlignedContribution;
    T totalSourceContribution;
};

// rws::ReverseWalkOnStarsSolver::solve (:287-330) on caller-made sample points (the reference draws them with its
// BoundarySampler / DomainSampler; zombie_b200/bvc.py has those samplers): one zb_rws_solve. samplePts' rng advances.
template <typename T, size_t DIM>
void solve(const GeometricQueries<DIM>& queries, const PDE<T, DIM>& pde, const WalkSettings& walkSettings,
           std::vector<SamplePoint<T, DIM>>& samplePts, float normalOffsetForAbsorbingBoundary, float radiusClamp,
           float kernelRegularization, std::vector<EvaluationPoint<T, DIM>>& evalPts)
{
    size_t n = samplePts.size(), m = evalPts.size();
    if (n == 0 || m == 0) return;
    std::vector<float> spt(n*DIM), snormal(n*DIM), pdf(n), da(n), dr(n);
    std::vector<int32_t> stype(n);
    std::vector<uint8_t> aligned(n);
    std::vector<uint64_t> state(n), inc(n);
    for (size_t i = 0; i < n; i++) {
        const SamplePoint<T, DIM>& p = samplePts[i];
        for (size_t k = 0; k < DIM; k++) { spt[i*DIM + k] = p.pt[k]; snormal[i*DIM + k] = p.normal[k]; }
        stype[i] = (int32_t)p.type; aligned[i] = p.estimateBoundaryNormalAligned ? 1 : 0; pdf[i] = p.pdf;
        da[i] = p.distToAbsorbingBoundary; dr[i] = p.distToReflectingBoundary; state[i] = p.rng.state; inc[i] = p.rng.inc;
    }
    zb_rws_samples s{(int32_t)n, spt.data(), snormal.data(), stype.data(), aligned.data(), pdf.data(), da.data(), dr.data(),
                     state.data(), inc.data()};
    std::vector<float> ept(m*DIM), enormal(m*DIM), totals(m*6);
    std::vector<int32_t> etype(m);
    for (size_t i = 0; i < m; i++) {
        const EvaluationPoint<T, DIM>& e = evalPts[i];
        for (size_t k = 0; k < DIM; k++) { ept[i*DIM + k] = e.pt[k]; enormal[i*DIM + k] = e.normal[k]; }
        etype[i] = (int32_t)e.type;
        totals[6*i] = e.totalPoissonKernelContribution; totals[6*i + 1] = e.totalAbsorbingBoundaryContribution;
        totals[6*i + 2] = e.totalAbsorbingBoundaryNormalAlignedContribution; totals[6*i + 3] = e.totalReflectingBoundaryContribution;
        totals[6*i + 4] = e.totalReflectingBoundaryNormalAlignedContribution; totals[6*i + 5] = e.totalSourceContribution;
    }
    zb_rws_eval_points ev{(int32_t)m, ept.data(), enormal.data(), etype.data(), totals.data()};
    zb_pde_desc desc = pde.descriptor();
    zb_walk_settings ws = walkSettings.toAbi();
    detail::check(zb_rws_solve(queries.scene(), &desc, &ws, &s, normalOffsetForAbsorbingBoundary, radiusClamp, kernelRegularization,
                               &ev, nullptr, nullptr));
    for (size_t i = 0; i < n; i++) samplePts[i].rng.state = state[i];
    for (size_t i = 0; i < m; i++) {
        EvaluationPoint<T, DIM>& e = evalPts[i];
        e.totalPoissonKernelContribution = totals[6*i]; e.totalAbsorbingBoundaryContribution = totals[6*i + 1];
        e.totalAbsorbingBoundaryNormalAlignedContribution = totals[6*i + 2]; e.totalReflectingBoundaryContribution = totals[6*i + 3];
        e.totalReflectingBoundaryNormalAlignedContribution = totals[6*i + 4]; e.totalSourceContribution = totals[6*i + 5];
        // evaluation points on the absorbing boundary take the Dirichlet data (:324-329)
        if (e.type == SampleType::OnAbsorbingBoundary)
            e.totalAbsorbingBoundaryContribution = !walkSettings.ignoreAbsorbingBoundaryContribution ? pde.dirichlet(e.pt, false) : T(0.0f);
    }
}

} // namespace rws

} // namespace zombie_b200

#if !defined(ZOMBIE_B200_NO_ALIAS)
namespace zombie = zombie_b200;
#endif
