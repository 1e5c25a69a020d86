// pcg32 (XSH-RR 64/32) on the device, bit-exact with deps/pcg32/pcg32.h:40-161 of the reference:
// same LCG multiplier, seeding sequence, output permutation, bounded-integer rejection and float
// construction. Pinned by the reference's own known-answer vector (deps/pcg32/pcg32-demo.out).
#pragma once

#include "vec.cuh"

namespace zb {

struct Pcg32 {
    uint64_t state, inc;

    static constexpr uint64_t kMult = 0x5851f42d4c957f2dULL;

    ZB_HOSTDEV void seed(uint64_t initstate, uint64_t initseq) {
        state = 0u;
        inc = (initseq << 1u) | 1u;
        nextUInt();
        state += initstate;
        nextUInt();
    }

    ZB_HOSTDEV uint32_t nextUInt() {
        uint64_t old = state;
        state = old*kMult + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31u));
    }

    ZB_HOSTDEV uint32_t nextUInt(uint32_t bound) {
        uint32_t threshold = (~bound + 1u) % bound;
        for (;;) {
            uint32_t r = nextUInt();
            if (r >= threshold) return r % bound;
        }
    }

    ZB_HOSTDEV float nextFloat() {
        uint32_t u = (nextUInt() >> 9) | 0x3f800000u;
#if defined(__CUDA_ARCH__)
        return zb_uint_as_float(u) - 1.0f;
#else
        float f;
        memcpy(&f, &u, 4);
        return f - 1.0f;
#endif
    }

    // jump ahead by delta draws in O(log delta) (Brown 1994), as pcg32::advance
    ZB_HOSTDEV void advance(uint64_t delta) {
        uint64_t curMult = kMult, curPlus = inc, accMult = 1u, accPlus = 0u;
        while (delta > 0) {
            if (delta & 1) {
                accMult *= curMult;
                accPlus = accPlus*curMult + curPlus;
            }
            curPlus = (curMult + 1)*curPlus;
            curMult *= curMult;
            delta >>= 1;
        }
        state = accMult*state + accPlus;
    }
};

// Per-walk generator. The reference draws all walks of a sample point sequentially from the point's
// own generator (zombie/point_estimation/common.h:182, walk_on_stars.h:602-637); walks here run
// concurrently, so walk w of a point owns a distinct pcg32 *stream* derived from the point's
// (state, inc): same generator, same seeding routine, stream id = point stream XOR (w+1) << 32.
ZB_HOSTDEV Pcg32 walkGenerator(uint64_t pointState, uint64_t pointInc, uint32_t walk)
{
    Pcg32 rng;
    uint64_t stream = (pointInc >> 1) ^ ((uint64_t)(walk + 1u) << 32);
    rng.seed(pointState + 0x9E3779B97F4A7C15ULL*(uint64_t)(walk + 1u), stream);
    return rng;
}

} // namespace zb
