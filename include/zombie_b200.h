/* zombie_b200 C ABI — the drop-in boundary for Zombie's random-walk hot path on B200.
 *
 * The reference has no FFI: its operator API is the header-only C++ template surface
 * (zombie/point_estimation/walk_on_stars.h:40-46, walk_on_spheres.h:38-44,
 * zombie/utils/fcpw_geometric_queries.h:84-150). Each entry point below names the reference
 * interface it replaces. All pointers are plain host pointers unless the function name ends
 * in _dev, in which case they are CUDA device pointers on the context's device. No function
 * ever calls exit(); errors are returned as codes and described by zb_last_error().
 *
 * Conventions: vectors are packed float[dim] per item (AoS of dim floats), dim is 2 or 3 and
 * fixed per scene. Primitive indices are the caller's input-order indices, like
 * fcpw::Interaction::primitiveIndex (fcpw/fcpw.inl:188-204).
 */
#ifndef ZOMBIE_B200_H
#define ZOMBIE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZB_ABI_VERSION 3 /* 3: on-device builder, moving geometry, value channels, per-side Dirichlet data (zb_pde_desc grew at its end) */

enum zb_status {
    ZB_OK = 0,
    ZB_ERR_INVALID_ARGUMENT = 1,
    ZB_ERR_CUDA = 2,
    ZB_ERR_OUT_OF_MEMORY = 3,
    ZB_ERR_UNSUPPORTED = 4,
    ZB_ERR_NO_DEVICE = 5
};

/* zombie::SampleType (point_estimation/common.h:139-143) */
enum zb_sample_type { ZB_IN_DOMAIN = 0, ZB_ON_ABSORBING_BOUNDARY = 1, ZB_ON_REFLECTING_BOUNDARY = 2 };
/* zombie::EstimationQuantity (common.h:145-149) */
enum zb_estimation_quantity { ZB_SOLUTION = 0, ZB_SOLUTION_AND_GRADIENT = 1, ZB_SKIP = 2 };
/* which FcpwXBoundaryHandler the reflecting mesh would have used (fcpw_geometric_queries.h:84-128) */
enum zb_reflecting_kind { ZB_REFLECTING_NONE = 0, ZB_REFLECTING_NEUMANN = 1, ZB_REFLECTING_ROBIN = 2 };
/* estimator: WalkOnSpheres::solve or WalkOnStars::solve */
enum zb_estimator { ZB_WALK_ON_SPHERES = 0, ZB_WALK_ON_STARS = 1 };
enum zb_boundary_id { ZB_ABSORBING = 0, ZB_REFLECTING = 1 };

/* Field-for-field mirror of zombie::WalkSettings (point_estimation/common.h:75-90);
 * bools widened to int32 so the layout is the same from C, C++ and ctypes. */
typedef struct zb_walk_settings {
    float epsilonShellForAbsorbingBoundary;
    float epsilonShellForReflectingBoundary;
    float silhouettePrecision;
    float russianRouletteThreshold;
    float splittingThreshold;
    int32_t maxWalkLength;
    int32_t stepsBeforeApplyingTikhonov;
    int32_t stepsBeforeUsingMaximalSpheres;
    int32_t solveDoubleSided;
    int32_t useGradientControlVariates;
    int32_t useGradientAntitheticVariates;
    int32_t useCosineSamplingForDerivatives;
    int32_t ignoreAbsorbingBoundaryContribution;
    int32_t ignoreReflectingBoundaryContribution;
    int32_t ignoreSourceContribution;
    int32_t printLogs;
} zb_walk_settings;

/* Dense regular samples of one PDE callback over a box: how the reference's applications fill the four callbacks of
 * zombie::PDE in practice (demo/model_problem.h:229-266 reads source / dirichlet / robin values from textures by
 * uv = (x - bMin)/maxLength). `data` is DEVICE memory for zb_solve / zb_solve_dev / zb_rws_solve (upload with
 * zb_grid_create); sample (i, j, k) is data[(k*shape[1] + j)*shape[0] + i] and sits at origin + (i, j, k)/invCell.
 * interp 0: nearest sample (the demo's Image::get); 1: multilinear between the eight (four in 2D) surrounding samples;
 * positions outside the box clamp to its faces. data == NULL: the callback falls back to `mode`. */
enum zb_grid_interp { ZB_GRID_NEAREST = 0, ZB_GRID_LINEAR = 1 };
typedef struct zb_grid_view {
    const float *data;
    int32_t shape[3];                    /* samples per axis; shape[2] = 1 in 2D */
    int32_t interp;                      /* zb_grid_interp */
    float origin[3];                     /* position of sample (0, 0, 0) */
    float invCell[3];                    /* samples per unit length along each axis */
} zb_grid_view;

/* Data-described zombie::PDE (core/pde.h:20-45). std::function callbacks cannot run on the
 * device, so each callback is a dense grid of samples when its zb_grid_view has data != NULL, and otherwise what `mode`
 * says: derived from a closed-form solution family (see zombie_b200/analytic.h) or a constant. */
enum zb_pde_mode { ZB_PDE_MANUFACTURED = 0, ZB_PDE_CONSTANT = 1 };
typedef struct zb_pde_desc {
    float absorptionCoeff;               /* pde.h:26 */
    int32_t isSourceConstant;            /* pde.h:27 */
    int32_t areRobinConditionsPureNeumann; /* pde.h:28 */
    int32_t areRobinCoeffsNonnegative;   /* pde.h:29 */
    int32_t mode;                        /* zb_pde_mode */
    int32_t family;                      /* zb_solution_family when mode == MANUFACTURED */
    float params[16];
    float robinCoeff;                    /* constant mu returned by pde.robinCoeff */
    float sourceValue;                   /* CONSTANT mode: pde.source */
    float dirichletValue;                /* CONSTANT mode: pde.dirichlet */
    float robinValue;                    /* CONSTANT mode: pde.robin */
    int32_t kelvin;                      /* 1: callbacks wrapped like KelvinTransform::transformPde (kelvin_transform.h:282-348) */
    float kelvinOrigin[3];
    /* grids (pde.h:31-44 filled from sampled data, demo/model_problem.h:229-266); each optional */
    zb_grid_view sourceGrid;             /* pde.source(x) */
    zb_grid_view dirichletGrid;          /* pde.dirichlet(x, false) */
    zb_grid_view robinGrid;              /* pde.robin(x, n, false) */
    zb_grid_view robinAlignedGrid;       /* pde.robin(x, n, true): the normal-aligned side of a double-sided boundary; NULL = same as robinGrid */
    zb_grid_view robinCoeffGrid;         /* pde.robinCoeff(x, n, .) */
    /* pde.dirichlet(x, true): the normal-aligned side of a double-sided absorbing boundary (walk_on_spheres.h:300-308,
     * walk_on_stars.h:499-506: the side the closest primitive's normal points to, when WalkSettings::solveDoubleSided);
     * grid with data == NULL and hasDirichletAlignedValue == 0: both sides read the same data */
    zb_grid_view dirichletAlignedGrid;
    float dirichletAlignedValue;         /* CONSTANT mode, when hasDirichletAlignedValue != 0 */
    int32_t hasDirichletAlignedValue;
} zb_pde_desc;

/* SoA mirror of std::vector<zombie::SamplePoint<float,DIM>> (common.h:152-195). rngState/rngInc
 * are pcg32::state / pcg32::inc and are in/out, as is firstSphereRadius. */
typedef struct zb_sample_points {
    int32_t n;
    const float *pt;                     /* n*dim */
    const float *normal;                 /* n*dim */
    const float *directionForDerivative; /* n*dim or NULL (= e_0) */
    const int32_t *type;                 /* zb_sample_type */
    const int32_t *estimationQuantity;   /* zb_estimation_quantity */
    const float *distToAbsorbingBoundary;
    const float *distToReflectingBoundary;
    const uint8_t *estimateBoundaryNormalAligned; /* or NULL (= false) */
    const int32_t *nWalks;               /* the nWalks vector of solve() */
    uint64_t *rngState;                  /* in/out */
    uint64_t *rngInc;                    /* in/out */
    float *firstSphereRadius;            /* in/out */
} zb_sample_points;

/* SoA mirror of std::vector<zombie::SampleStatistics<float,DIM>> (common.h:201-395), in/out so
 * that repeated solves refine progressively exactly like the reference (common.h:268-310). */
typedef struct zb_statistics {
    float *solutionMean;
    float *solutionM2;
    float *gradientMean;                 /* n*dim */
    float *gradientM2;                   /* n*dim */
    float *totalFirstSourceContribution;
    float *totalDerivativeContribution;
    int32_t *nSolutionEstimates;
    int32_t *nGradientEstimates;
    int32_t *totalWalkLength;
    int32_t *totalSplits;
} zb_statistics;

typedef struct zb_context zb_context;
typedef struct zb_scene zb_scene;

/* ---- context ---------------------------------------------------------------------------- */
const char *zb_last_error(void);
int zb_abi_version(void);
int zb_device_count(void);
/* one context per process and device; owns its CUDA stream and scratch buffers */
int zb_context_create(int device, zb_context **out);
void zb_context_destroy(zb_context *ctx);
/* runs subsequent work of this context on a caller-owned cudaStream_t (e.g. torch's current stream) */
int zb_context_set_stream(zb_context *ctx, void *cudaStream);
int zb_context_synchronize(zb_context *ctx);
int zb_context_device(const zb_context *ctx);
/* Measurement hook (no reference counterpart; the reference only prints whole seconds,
 * zombie/utils/progress.h:32-39): CUDA-event durations, on the context's stream, of the last
 * zb_solve/zb_solve_dev — the per-point preparation kernels, the persistent walk kernel and the
 * per-point reduction — plus the number of walks it ran, the walk kernel's grid size and the
 * number of kernels this context has launched so far. Blocks until that solve has finished. */
int zb_context_last_solve_profile(zb_context *ctx, float *prepareMs, float *walkMs, float *reduceMs,
                                  int64_t *walks, int32_t *walkGrid, int64_t *kernelLaunches);
int zb_context_sm_count(const zb_context *ctx);

/* ---- geometry: replaces Fcpw{Dirichlet,Neumann,Robin}BoundaryHandler::buildAccelerationStructure
 * + populateGeometricQueriesFor*Boundary (fcpw_geometric_queries.h:569-586, 650-676, 870-995,
 * 1063-1291) and the GeometricQueries(domainIsWatertight, domainMin, domainMax) constructor
 * (core/geometric_queries.h:135-147). Either mesh may be empty (n == 0). Indices are dim per
 * primitive (line segments in 2D, triangles in 3D). minRobin/maxRobin are per reflecting
 * primitive and only read for ZB_REFLECTING_ROBIN. ------------------------------------------ */
int zb_scene_create(zb_context *ctx, int dim,
                    const float *absorbingPositions, int nAbsorbingVertices,
                    const int32_t *absorbingIndices, int nAbsorbingPrimitives,
                    const float *reflectingPositions, int nReflectingVertices,
                    const int32_t *reflectingIndices, int nReflectingPrimitives,
                    int reflectingKind, const float *minRobinCoeff, const float *maxRobinCoeff,
                    int solveDoubleSided, float silhouettePrecision,
                    int domainIsWatertight, const float *domainMin, const float *domainMax,
                    zb_scene **out);
void zb_scene_destroy(zb_scene *scene);
/* FcpwRobinBoundaryHandler::updateCoefficientValues (fcpw_geometric_queries.h:997-1010) */
int zb_scene_update_robin_coefficients(zb_scene *scene, const float *minRobinCoeff, const float *maxRobinCoeff);
/* The same with the two coefficient arrays (one value per reflecting primitive, input order) already in DEVICE memory;
 * asynchronous on the context's stream. Both forms refit on the device: per-primitive records, then the coefficient
 * range of every node bottom-up (ReflectanceBvh::updateCoefficients, reflectance_boundary_bvh/bvh.h:304-367), one
 * kernel per level of the hierarchy; geometry and topology are untouched. */
int zb_scene_update_robin_coefficients_dev(zb_scene *scene, const float *minRobinCoeff, const float *maxRobinCoeff);

/* ---- PDE data grids ------------------------------------------------------------------------
 * Uploads shape[0]*shape[1]*shape[2] floats (x fastest) to the context's device and returns a handle; zb_grid_data
 * is the device pointer to put into a zb_grid_view. Replaces the reference's texture-backed callbacks
 * (demo/model_problem.h:229-266, demo/image.h). */
typedef struct zb_grid zb_grid;
int zb_grid_create(zb_context *ctx, const float *hostData, const int32_t shape[3], zb_grid **out);
const float *zb_grid_data(const zb_grid *grid);
void zb_grid_destroy(zb_grid *grid);
/* bytes of device memory held by the scene's acceleration structures, and node counts */
int zb_scene_stats(const zb_scene *scene, int boundary, int64_t *bytes, int32_t *nNodes, int32_t *nLeaves, int32_t *maxDepth);

/* ---- building the hierarchy on the device, moving geometry -----------------------------------------------------
 * Which builder zb_scene_create uses for scenes created through this context afterwards. HOST_SAH (default): binned
 * SAH on the host, as the reference builds its hierarchies (fcpw/aggregates/bvh.inl:143-234) — the best trees, 0.5-1 s
 * per million triangles. DEVICE: Morton order + level-synchronous 8-ary cuts on the GPU (zombie_b200/csrc/bvh_build.cuh),
 * including the adjacency tables, normal cones and Robin-coefficient ranges of reflecting boundaries
 * (fcpw_geometric_queries.h:728-795, 892-973; reflectance_boundary_bvh/bvh.h:77-186) — milliseconds, somewhat slower
 * traversals. Both produce the same layout and the same query results (per-primitive records are bit-identical; the
 * quantised boxes and cones of either tree are conservative). */
enum zb_bvh_builder { ZB_BVH_BUILDER_HOST_SAH = 0, ZB_BVH_BUILDER_DEVICE = 1 };
int zb_context_set_bvh_builder(zb_context *ctx, int builder);
/* Moving geometry: new positions for ALL vertices of one boundary (same vertex count and connectivity). Replaces
 * fcpw Scene::updateObjectVertex + Scene::refit (fcpw/fcpw.h:76-107, aggregates/bvh.inl:549-600) and ReflectanceBvh::refit
 * (reflectance_boundary_bvh/bvh.h:271-312): the hierarchy — boxes, adjacency normals and silhouette marks, cones — is
 * rebuilt on the device from the mesh kept there (a rebuild costs about what a refit pass would and keeps tree quality
 * when the motion is large). Current Robin coefficients are kept, and so is the bounding domain given at creation
 * (domainMin / domainMax: make it large enough for the motion). The _dev form reads positions from device memory. */
int zb_scene_update_vertices(zb_scene *scene, int boundary, const float *positions, int nVertices);
int zb_scene_update_vertices_dev(zb_scene *scene, int boundary, const float *positions, int nVertices);
/* duration of the last build of this boundary's hierarchy (HOST_SAH: host build + upload; DEVICE: the kernels, by
 * CUDA events) and the builder that produced it */
int zb_scene_build_profile(const zb_scene *scene, int boundary, float *ms, int32_t *builder);
/* Test hooks: soundness check of the hierarchy as it sits on the device (downloaded and checked like zb_host_bvh_check:
 * every primitive once, inside the quantised box of every enclosing child slot, internal children contiguous, leaves of
 * 1..4 primitives, breadth-first level table) against the caller's mesh; and the leaf-order per-primitive records
 * (primIndex: n; prims: n*12 floats in 3D, n*4 in 2D; adjacency: n*12 floats, reflecting boundaries only; NULL skips). */
int zb_scene_validate(zb_scene *scene, int boundary, const float *positions, int nVertices, const int32_t *indices, int nPrimitives,
                      int64_t *violations);
int zb_scene_download_records(zb_scene *scene, int boundary, int32_t *primIndex, float *prims, float *adjacency);

/* Host-only self-check of the acceleration-structure builder (no device needed; used by the CPU tests):
 * builds the wide quantised hierarchy for the mesh exactly as zb_scene_create does and verifies that every
 * primitive is referenced once and lies inside the quantised box of every enclosing child slot.
 * Outputs: node / leaf counts, depth, and the number of violations (0 = sound). */
int zb_host_bvh_check(int dim, const float *positions, int nVertices, const int32_t *indices, int nPrimitives,
                      int reflecting, int32_t *nNodes, int32_t *nLeaves, int32_t *maxDepth, int64_t *violations);

/* ---- batched geometric queries (host buffers). These are what the GeometricQueries callbacks
 * bind to; they also expose the primitive index so parity with FCPW can be checked exactly. ---- */
/* Aggregate::findClosestPoint with BoundingSphere(x, maxR2) (fcpw_geometric_queries.h:1069-1096;
 * fcpw/aggregates/bvh.inl:1391-1418). prim = -1 when nothing lies within the radius. normal may be NULL.
 * Exact ties: when several primitives are equidistant in FCPW's sense (equal d*d; the closest point lies on a
 * shared edge or vertex), FCPW reports the one it visits last — its BVH by tree topology, its brute-force
 * Baseline aggregate (baseline.inl:349-372) by input order. This function reports the Baseline's answer on
 * every query: the LARGEST input index among the tied primitives, with that primitive's point and normal. */
int zb_closest_point(zb_scene *scene, int boundary, int n, const float *x, const float *maxRadius /* NULL = inf */,
                     int32_t *prim, float *dist, float *pt, float *normal);
/* Aggregate::intersect(Ray(origin, dir, tMax)) closest hit (bvh.inl:753-771). prim = -1 on miss. */
int zb_intersect_ray(zb_scene *scene, int boundary, int n, const float *origin, const float *dir, const float *tMax,
                     int32_t *prim, float *dist, float *pt, float *normal);
/* !Aggregate::hasLineOfSight(a, b) (fcpw/core/primitive.h:202-212): 1 = occluded */
int zb_occluded(zb_scene *scene, int boundary, int n, const float *a, const float *b, uint8_t *occluded);
/* Aggregate::intersect(ray, interactions, checkForOcclusion = false, recordAllHits = true)
 * (fcpw/aggregates/bvh.inl:890-1028, what intersect{Absorbing,Reflecting}BoundaryAllHits bind to,
 * fcpw_geometric_queries.h:1116-1136) reduced to its return value: the number of distinct hits within tMax — hits
 * sorted by distance, neighbours merged when Interaction::operator== holds (interaction.h:20-25). */
int zb_count_ray_hits(zb_scene *scene, int boundary, int n, const float *origin, const float *dir, const float *tMax /* NULL = inf */,
                      int32_t *count);
/* GeometricQueries::insideDomain (core/geometric_queries.h:357-371): 1 when the all-hits counts over both boundaries
 * along +x, +y (, +z) are all odd; always 1 for a scene created with domainIsWatertight = 0. The inside test of
 * the domain samplers (variance_reduction/domain_samplers.h:96) and of the demo's solve-region masks. */
int zb_inside_domain(zb_scene *scene, int n, const float *x, uint8_t *inside);
/* Aggregate::intersect(sphere, randNums, branchTraversalWeight) (bvh.inl:1257-1288) with the
 * harmonic branch weight of getBranchTraversalWeightCallback (fcpw_geometric_queries.h:1391-1398).
 * prim = -1 when no primitive was sampled. */
int zb_sample_sphere(zb_scene *scene, int boundary, int n, const float *x, const float *radius, const float *randNums /* n*dim */,
                     int32_t *prim, float *pt, float *normal, float *pdf);
/* GeometricQueries::computeStarRadiusForReflectingBoundary (fcpw_geometric_queries.h:1250-1291) */
int zb_star_radius(zb_scene *scene, int n, const float *x, const float *minRadius, const float *maxRadius,
                   float silhouettePrecision, const uint8_t *flipNormalOrientation /* or NULL */, float *radius);
/* GeometricQueries::insideDomain is not on the walk path; outsideBoundingDomain is evaluated inline. */

/* ---- estimators: replace WalkOnSpheres::solve / WalkOnStars::solve (vector overloads,
 * walk_on_spheres.h:132-168, walk_on_stars.h:151-187). stats and pts->rng* / firstSphereRadius
 * are in/out. walkSteps (optional) receives the number of walk steps taken by this call. ------ */
int zb_solve(zb_scene *scene, int estimator, const zb_pde_desc *pde, const zb_walk_settings *settings,
             const zb_sample_points *pts, zb_statistics *stats, int64_t *walkSteps);
/* same, all arrays already resident on the context's device; no host<->device copy of point or statistics data.
 * The call enqueues its work on the context's stream in batches of walk rounds and reads one 4-byte live-walk
 * counter per batch back (with one batch of lag, so the device never idles) to know when the walks have ended: it
 * returns when the last batch has been enqueued, NOT when it has run — results are ordered behind it on the
 * context's stream like those of any asynchronous call, but the call itself lasts about as long as the walks.
 * walkSteps_dev points to one int64 on the device. */
int zb_solve_dev(zb_scene *scene, int estimator, const zb_pde_desc *pde, const zb_walk_settings *settings,
                 const zb_sample_points *pts_dev, zb_statistics *stats_dev, int64_t *walkSteps_dev);

/* ---- vector-valued solves: WalkOnSpheres<T, DIM> / WalkOnStars<T, DIM>::solve with T = Array<float, C>, the channels of
 * the reference's bindings (python/zombie_bindings.cpp:37-44: 1, 3 or more value channels; core/pde.h:31-44: source,
 * dirichlet and robin return T, absorptionCoeff and robinCoeff stay floats). C PDE descriptors, C statistics structs,
 * ONE set of walks: geometry, random numbers, throughput and the weight window are shared exactly as the reference
 * shares them between the components of T; every channel evaluates its own source / Dirichlet / Robin data at the
 * points the walk samples and keeps its own estimates. Channel c of the result equals a scalar zb_solve of pdes[c] on
 * the same generators (up to the contraction of multiply-adds); the cost is that of one scalar solve plus a few
 * loads and stores per channel and step. The descriptors must agree in everything that steers the walk
 * (absorptionCoeff, robinCoeff / robinCoeffGrid, the three flags, kelvin); sample-point arrays are shared, rng state
 * advances once. nChannels 1..256. ---- */
int zb_solve_channels(zb_scene *scene, int estimator, int nChannels, const zb_pde_desc *pdes /* nChannels */,
                      const zb_walk_settings *settings, const zb_sample_points *pts, zb_statistics *stats /* nChannels, in/out */,
                      int64_t *walkSteps);
/* same with the sample-point arrays and the arrays inside every stats[c] resident on the context's device (pdes and the
 * stats structs themselves are host memory); stream semantics of zb_solve_dev */
int zb_solve_channels_dev(zb_scene *scene, int estimator, int nChannels, const zb_pde_desc *pdes, const zb_walk_settings *settings,
                          const zb_sample_points *pts_dev, zb_statistics *stats, int64_t *walkSteps_dev);

/* ---- Boundary Value Caching: replaces the splat stage of bvc::BoundaryValueCaching
 * (zombie/variance_reduction/boundary_value_caching.h:371-502, 682-811): every cache sample is splatted onto
 * every evaluation point with the free-space Green's function of the PDE (Harmonic / Yukawa,
 * zombie/core/distributions.h:124-321). ------------------------------------------------------------------ */
/* SoA of the zombie::SamplePoint fields the splat reads (common.h:182-194). type: ZB_ON_ABSORBING_BOUNDARY /
 * ZB_ON_REFLECTING_BOUNDARY = boundary sample (splatBoundaryData), ZB_IN_DOMAIN = source sample
 * (splatSourceData, `contribution` = pde.source at the sample). Always host arrays (caches are small). */
typedef struct zb_bvc_samples {
    int32_t n;
    const float *pt;                              /* n*dim */
    const float *normal;                          /* n*dim */
    const int32_t *type;                          /* zb_sample_type */
    const uint8_t *estimateBoundaryNormalAligned; /* or NULL (= false) */
    const float *pdf;
    const float *solution;
    const float *normalDerivative;
    const float *contribution;
    const float *robinCoeff;
} zb_bvc_samples;

/* SoA of bvc::EvaluationPoint (boundary_value_caching.h:18-56). Its five SampleStatistics (absorbing,
 * absorbing normal-aligned, reflecting, reflecting normal-aligned, source) are kept as (count, sums): a
 * Welford mean over N estimates is sum/N. bucketCount / bucketSum are in/out, so repeated splats accumulate
 * like the reference's statistics and shards merge by addition; solution / gradient are outputs. */
#define ZB_BVC_BUCKETS 5
typedef struct zb_bvc_eval_points {
    int32_t n;
    const float *pt;                       /* n*dim */
    const int32_t *type;                   /* zb_sample_type */
    const float *distToAbsorbingBoundary;
    const float *distToReflectingBoundary;
    int32_t *bucketCount;                  /* n*5, in/out */
    float *bucketSum;                      /* n*5*(1+dim): solution, gradient[dim]; in/out */
    float *solution;                       /* n: EvaluationPoint::getEstimatedSolution (:263-273) */
    float *gradient;                       /* n*dim: EvaluationPoint::getEstimatedGradient (:276-286) */
} zb_bvc_eval_points;

/* BoundaryValueCaching::splat(pde, samplePts, radiusClamp, kernelRegularization,
 * robinCoeffCutoffForNormalDerivative, cutoffDistToAbsorbingBoundary, cutoffDistToReflectingBoundary,
 * evalPts) (:482-502). absorptionCoeff = pde.absorptionCoeff selects Yukawa (> 0) or harmonic kernels. pairs
 * (optional) receives the number of (sample, evaluation point) pairs evaluated. Evaluation-point arrays are host pointers. */
int zb_bvc_splat(zb_context *ctx, int dim, float absorptionCoeff, const zb_bvc_samples *samples,
                 float radiusClamp, float kernelRegularization, float robinCoeffCutoffForNormalDerivative,
                 float cutoffDistToAbsorbingBoundary, float cutoffDistToReflectingBoundary,
                 zb_bvc_eval_points *evalPts, int64_t *pairs);
/* same with the evaluation-point arrays resident on the context's device; asynchronous on its stream */
int zb_bvc_splat_dev(zb_context *ctx, int dim, float absorptionCoeff, const zb_bvc_samples *samples,
                     float radiusClamp, float kernelRegularization, float robinCoeffCutoffForNormalDerivative,
                     float cutoffDistToAbsorbingBoundary, float cutoffDistToReflectingBoundary,
                     zb_bvc_eval_points *evalPts_dev, int64_t *pairs);
/* CUDA-event duration of the last zb_bvc_splat(_dev) kernel on this context (measurement hook) */
int zb_context_last_splat_ms(zb_context *ctx, float *ms);

/* ---- Reverse Walk on Stars: replaces rws::ReverseWalkOnStarsSolver::solve
 * (zombie/variance_reduction/reverse_walk_solver.h:287-330) = ReverseWalkOnStars::solve at every sample point
 * (zombie/point_estimation/reverse_walk_on_stars.h:89-195) + rws::splatContribution at every step (:233-283). ---- */
/* SoA of the zombie::SamplePoint fields a reverse walk reads (common.h:182-194). type: ZB_ON_ABSORBING_BOUNDARY (sample
 * displaced inward by normalOffsetForAbsorbingBoundary), ZB_ON_REFLECTING_BOUNDARY, ZB_IN_DOMAIN (source sample). The
 * contribution of every sample (Dirichlet / Robin / source value) is evaluated from the PDE like the reference does.
 * rngState is in/out. Host arrays. */
typedef struct zb_rws_samples {
    int32_t n;
    const float *pt;                              /* n*dim */
    const float *normal;                          /* n*dim */
    const int32_t *type;                          /* zb_sample_type */
    const uint8_t *estimateBoundaryNormalAligned; /* or NULL */
    const float *pdf;
    const float *distToAbsorbingBoundary;
    const float *distToReflectingBoundary;
    uint64_t *rngState;                           /* in/out */
    uint64_t *rngInc;
} zb_rws_samples;

/* SoA of rws::EvaluationPoint (reverse_walk_solver.h:20-52). totals: 6 running sums per point — totalPoissonKernel-
 * Contribution, totalAbsorbingBoundary-, -NormalAligned-, totalReflectingBoundary-, -NormalAligned-, totalSource-
 * Contribution — in/out, so repeated solves accumulate like the reference's; EvaluationPoint::getEstimatedSolution
 * (:129-166) divides them by the sample counts. */
typedef struct zb_rws_eval_points {
    int32_t n;
    const float *pt;      /* n*dim */
    const float *normal;  /* n*dim or NULL */
    const int32_t *type;  /* zb_sample_type; points on the absorbing boundary take no splats (:251) */
    float *totals;        /* n*6, in/out */
} zb_rws_eval_points;

/* walkSteps / splatRecords (optional): walk steps taken and splat records (= ball visits) gathered. */
int zb_rws_solve(zb_scene *scene, const zb_pde_desc *pde, const zb_walk_settings *settings, const zb_rws_samples *samples,
                 float normalOffsetForAbsorbingBoundary, float radiusClamp, float kernelRegularization,
                 zb_rws_eval_points *evalPts, int64_t *walkSteps, int64_t *splatRecords);

/* pcg32 known-answer hook (deps/pcg32/pcg32.h:53-100): seeds a device generator with
 * (initstate, initseq) and returns n successive nextUInt() values */
int zb_pcg32_kat(zb_context *ctx, uint64_t initstate, uint64_t initseq, int n, uint32_t *out);
/* the device generator run through the whole sequence of deps/pcg32/pcg32-demo.cpp:64-104 (per round: 6
 * nextUInt, 65 nextUInt(2), 33 nextUInt(6)+1, pcg32::shuffle of 52 cards), so that the committed output of
 * the reference's demo (deps/pcg32/pcg32-demo.out) pins nextUInt(bound) and the shuffle as well.
 * u32: rounds*6, coins: rounds*65, rolls: rounds*33, cards: rounds*52 */
int zb_pcg32_demo(zb_context *ctx, uint64_t initstate, uint64_t initseq, int rounds,
                  uint32_t *u32, uint8_t *coins, uint8_t *rolls, uint8_t *cards);
/* n successive nextFloat() values (deps/pcg32/pcg32.h:103-112) */
int zb_pcg32_floats(zb_context *ctx, uint64_t initstate, uint64_t initseq, int n, float *out);

#ifdef __cplusplus
}
#endif

#endif /* ZOMBIE_B200_H */
