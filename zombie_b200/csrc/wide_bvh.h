// Wide (8-ary), 8-bit-quantised bounding volume hierarchy over line segments (2D) or triangles (3D).
//
// This replaces the binary, pointer-chasing fcpw::Bvh (fcpw/aggregates/bvh.h:19-48,69-225) and
// zombie's ReflectanceBvh (zombie/utils/reflectance_boundary_bvh/bvh.h:13-27) with one flat layout
// designed for one-thread-per-walk traversal on B200:
//
//   * a node is 80 bytes = 5 x 16-byte loads and describes up to 8 children: a float3 origin, one
//     power-of-two scale per axis, and for every child an 8-bit [lo,hi] box per axis stored
//     "SoA inside the node" (8 lo_x bytes, 8 lo_y bytes, ...), so a thread decodes all children
//     of an axis from two 32-bit registers;
//   * children that are internal nodes are contiguous (childBase + rank); children that are leaves
//     reference 1..4 primitives stored contiguously in leaf order (primBase + offset);
//   * primitives are stored de-indexed in leaf order (3 x float4 per triangle, 1 x float4 per
//     segment), so a leaf test is a contiguous 48/16-byte read instead of index -> vertex gathers;
//   * reflecting boundaries carry, per child, a normal cone and Robin-coefficient range (32 bytes)
//     used only by the star-radius query.
// The quantised boxes are conservative supersets of the exact boxes, so every query returns the
// same primitive a brute-force scan returns; only the number of visited nodes changes.
#pragma once

#include <cstdint>
#include <vector>

namespace zb {

struct WideNode {
    // w[0..2] origin xyz (float bits); w[3] = ex | ey<<8 | ez<<16 (biased exponents of the scales)
    // w[4] childBase; w[5] primBase; w[6..7] meta bytes of children 0..7
    // w[8..9] lo_x[8]; w[10..11] lo_y[8]; w[12..13] lo_z[8]; w[14..15] hi_x[8]; w[16..17] hi_y[8]; w[18..19] hi_z[8]
    uint32_t w[20];
};
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 bytes");

// meta byte of a child slot: 0 = empty; (count << 5) | offset = leaf with `count` primitives at
// primBase + offset; count == 0 && meta != 0 = internal child number (meta & 31) - 1
constexpr int kWidth = 8;
constexpr int kMaxLeafSize = 4;

// per-child data for the star-radius query of reflecting boundaries (2 x float4)
struct ChildCone {
    float axis[3];
    float halfAngle;     // >= pi/2 means "may contain a silhouette from anywhere"
    float radius;        // max distance from the (decoded) box centroid to an edge midpoint / vertex
    float minCoeff;
    float maxCoeff;
    float noTestableEdge; // 1: every edge below is marked ignoreAdjacentFace (single-sided solves: convex edges), else 0
};
static_assert(sizeof(ChildCone) == 32, "ChildCone must be 32 bytes");

// per-primitive data of reflecting boundaries, leaf order. 3D: adjacent face normals across edges
// (v0v1, v1v2, v2v0); 2D: adjacent segment normals at vertices (v0, v1).
struct PrimAdjacency {
    float n[3][3];
    float minCoeff;
    float maxCoeff;
    uint32_t flags;      // bit j: hasAdjacentFace[j]; bit 4 + j: ignoreAdjacentFace[j]
};
static_assert(sizeof(PrimAdjacency) == 48, "PrimAdjacency must be 48 bytes");

struct HostBvh {
    int dim = 3;
    int nPrims = 0;
    std::vector<WideNode> nodes;        // node 0 is the root
    std::vector<float> prims;           // leaf order; 3D: 12 floats / triangle (v0,idx | v1,0 | v2,0); 2D: 4 floats / segment
    std::vector<int32_t> primIndex;     // leaf order -> caller's primitive index
    std::vector<ChildCone> cones;       // nodes.size()*8 when built for a reflecting boundary
    std::vector<PrimAdjacency> adjacency; // leaf order, reflecting only
    int nLeaves = 0;
    int maxDepth = 0;
    float boxMin[3] = {0, 0, 0}, boxMax[3] = {0, 0, 0};
};

struct ReflectingOptions {
    bool enabled = false;
    bool solveDoubleSided = false;       // getIgnoreCandidateSilhouetteCallback (fcpw_geometric_queries.h:1029-1038)
    float silhouettePrecision = 1e-3f;
    const float *minCoeff = nullptr;     // per input primitive, may be null (=0: pure Neumann)
    const float *maxCoeff = nullptr;
};

// Builds the hierarchy. positions: nV*dim floats; indices: nPrim*dim ints. Returns false and fills
// err on invalid input.
bool buildWideBvh(int dim, const float *positions, int nV, const int32_t *indices, int nPrim,
                  const ReflectingOptions& refl, HostBvh& out, const char **err);

// Re-derives per-child coefficient ranges after the per-primitive coefficients changed
// (FcpwRobinBoundaryHandler::updateCoefficientValues, fcpw_geometric_queries.h:997-1010).
void updateCoefficients(HostBvh& bvh, const float *minCoeff, const float *maxCoeff);

// Self-check of a built hierarchy (host only, used by the CPU test-suite): every primitive appears exactly
// once, lies inside the decoded (quantised) box of every ancestor child slot, internal children are
// contiguous, leaves hold 1..4 primitives. Returns the number of violations found.
long long validateWideBvh(const HostBvh& bvh, int dim, const float *positions, const int32_t *indices);

// decodes child `slot` of `node` exactly as the device does
void decodeChildBox(const WideNode& node, int slot, float lo[3], float hi[3]);

} // namespace zb
