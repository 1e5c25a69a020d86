// DEBUG / TEST TOOL: a 32-lane warp emulated by 32 host threads, so that the warp-cooperative traversal of
// zombie_b200/csrc/coop.cuh can be executed and checked on a machine without a GPU (tools/coop_check.cu).
// Every full-mask warp intrinsic becomes "publish my value, barrier, read the others, barrier".
#pragma once

#include <atomic>
#include <cmath>
#include <cstdint>
#include <thread>

namespace zbemu {

struct Warp {
    std::atomic<int> arrived{0};
    std::atomic<int> phase{0};
    int slots[32];
    void barrier() {
        int p = phase.load(std::memory_order_acquire);
        if (arrived.fetch_add(1, std::memory_order_acq_rel) == 31) {
            arrived.store(0, std::memory_order_relaxed);
            phase.store(p + 1, std::memory_order_release);
        } else {
            int spins = 0;
            while (phase.load(std::memory_order_acquire) == p) { if (++spins > 64) { std::this_thread::yield(); spins = 0; } }
        }
    }
};

inline thread_local Warp *tlsWarp = nullptr;
inline thread_local int tlsLane = 0;

inline int lane() { return tlsLane; }
inline void syncwarp() { tlsWarp->barrier(); }
inline unsigned ballot(bool p)
{
    Warp& w = *tlsWarp;
    w.slots[tlsLane] = p ? 1 : 0;
    w.barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)w.slots[i] << i;
    w.barrier();
    return m;
}
inline int shfl(int v, int src)
{
    Warp& w = *tlsWarp;
    w.slots[tlsLane] = v;
    w.barrier();
    int r = w.slots[src & 31];
    w.barrier();
    return r;
}
inline int shflUp(int v, int o)
{
    Warp& w = *tlsWarp;
    w.slots[tlsLane] = v;
    w.barrier();
    int r = tlsLane >= o ? w.slots[tlsLane - o] : v;
    w.barrier();
    return r;
}
inline void atomicMin64(unsigned long long *p, unsigned long long v)
{
    unsigned long long cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}

} // namespace zbemu

#define ZB_COOP inline
#define ZB_COOP_NOINLINE inline
#define zb_lane() zbemu::lane()
#define zb_syncwarp() zbemu::syncwarp()
#define zb_ballot(p) zbemu::ballot(p)
#define zb_shfl(v, src) zbemu::shfl((v), (src))
#define zb_shfl_up(v, o) zbemu::shflUp((v), (o))
#define zb_atomic_min64(p, v) zbemu::atomicMin64((p), (v))
#define zb_status_or(p, v) __atomic_fetch_or((p), (v), __ATOMIC_RELAXED)
#define zb_popc(x) __builtin_popcount(x)
