// Device-side storage of one boundary's acceleration structure, shared by capi.cu (scenes, queries, solves) and
// bvh_build.cu (the on-device builder).
#pragma once

#include <vector>

#include "capi_common.h"
#include "queries.cuh"
#include "wide_bvh.h"

struct DeviceBvhStorage {
    zb::HostBvh host;                    // host builder: the whole hierarchy; device builder: dim, counts and bounds only
    zb::DeviceBuffer nodes, prims, primIndex, cones, adjacency;
    zb::DeviceBuffer coeffScratch;       // Robin coefficient refit: the caller's per-primitive values and a (min, max) per node
    // the caller's mesh, kept on the device so that the hierarchy can be rebuilt there when vertices move
    zb::DeviceBuffer meshPos, meshIdx, coefMin, coefMax;
    int nVert = 0;
    int nNodes = 0;
    bool reflecting = false, robin = false;
    int builder = 0;                     // zb_bvh_builder that produced the current hierarchy
    float buildMs = 0.0f;                // host builder: wall clock of build + upload; device builder: CUDA events
    std::vector<int> levelStart;         // nodes are stored breadth first: level l is [levelStart[l], levelStart[l + 1])
    zb::BvhView view{};
    bool present = false;
    size_t bytes() const { return nodes.capacity + prims.capacity + primIndex.capacity + cones.capacity + adjacency.capacity; }
};

namespace zb {
// Builds the hierarchy of s.meshPos / s.meshIdx (nPrim primitives, already on the device) on the device: fills the
// node / primitive / cone / adjacency buffers, the view, the level table and the counts. Asynchronous pieces run on
// the context's stream; the call returns after the build has finished (it reads the level sizes back as it goes).
int buildBvhOnDevice(zb_context *ctx, DeviceBvhStorage& s, int dim, int nPrim, bool doubleSided, float silhouettePrecision,
                     int snchSign);
} // namespace zb
