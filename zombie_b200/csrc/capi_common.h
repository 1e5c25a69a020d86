// Pieces shared by the translation units that implement the C ABI (capi.cu, bvc_capi.cu): error
// reporting, the grow-only device buffer and the context object.
#pragma once

#include <cuda_runtime.h>

#include <string>

#include "../../include/zombie_b200.h"

namespace zb {

inline std::string& lastErrorRef()
{
    static thread_local std::string err;
    return err;
}

inline int fail(int code, const std::string& msg)
{
    lastErrorRef() = msg;
    return code;
}

#define ZB_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t err__ = (expr);                                                                \
        if (err__ != cudaSuccess) {                                                                \
            return zb::fail(err__ == cudaErrorMemoryAllocation ? ZB_ERR_OUT_OF_MEMORY : ZB_ERR_CUDA, \
                            std::string(#expr) + ": " + cudaGetErrorString(err__));                \
        }                                                                                          \
    } while (0)

// grow-only device allocation
struct DeviceBuffer {
    void *ptr = nullptr;
    size_t capacity = 0;
    ~DeviceBuffer() { if (ptr) cudaFree(ptr); }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= capacity) return cudaSuccess;
        if (ptr) { cudaFree(ptr); ptr = nullptr; capacity = 0; }
        size_t want = bytes + bytes/8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) capacity = want;
        return e;
    }
    // exact size (long-lived arrays: the hierarchies of a scene), still grow-only
    cudaError_t reserveExact(size_t bytes) {
        if (bytes <= capacity) return cudaSuccess;
        if (ptr) { cudaFree(ptr); ptr = nullptr; capacity = 0; }
        cudaError_t e = cudaMalloc(&ptr, bytes);
        if (e == cudaSuccess) capacity = bytes;
        return e;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

} // namespace zb

struct zb_context {
    int device = 0;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    bool ownsStream = false;
    zb::DeviceBuffer status, counters, perPoint, strat, records, scanTemp, io;
    zb::DeviceBuffer bvcSamples, bvcIo, rwsRecords, rwsSorted, rwsScratch;
    zb::DeviceBuffer channelRecords, channelDesc, channelIo; // extra value channels of zb_solve_channels: per-task sums, descriptors, staged statistics
    zb::DeviceBuffer bvhBuild;             // scratch of the on-device hierarchy builder (grow-only: rebuilds of moving geometry allocate nothing)
    zb::DeviceBuffer wave, waveSort;       // walk state of the wavefront scheduler, radix-sort scratch
    unsigned int *hostLive = nullptr;      // pinned: live-walk counts read back every few rounds
    cudaEvent_t evRound[16] = {};          // two per slot group
    cudaEvent_t evFork = nullptr;
    cudaStream_t waveStreams[8] = {};      // streams of the slot groups 1.. (group 0 runs on the context's stream)
    int lastRounds = 0;
    int *hostStatus = nullptr; // pinned
    int bvhBuilder = 0;        // zb_bvh_builder used by zb_scene_create
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; // solve start | walk start | walk end | solve end
    cudaEvent_t evSplat[2] = {nullptr, nullptr};
    bool profiled = false, splatProfiled = false;
    long long lastTasks = 0;
    int lastGrid = 0, launches = 0;
};
