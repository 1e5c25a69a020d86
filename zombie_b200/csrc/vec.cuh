// Minimal fixed-size float vectors for device and host code (the reference uses Eigen::Matrix<float,DIM,1>).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

// Query and walk code is host+device so that tools/host_probe.cu can single-step the exact device
// logic on a CPU for debugging and traversal statistics; the shipped library only ever calls it from
// kernels (there is no CPU execution path behind the C ABI).
#define ZB_DEV __host__ __device__ __forceinline__
#define ZB_HOSTDEV __host__ __device__ __forceinline__
// rarely executed paths that must stay OUT of the hot loops' instruction footprint: the walk kernel is bound by
// instruction fetch (profiles/README.md), so what is cold is kept out of line on the device
#define ZB_DEV_COLD __host__ __device__ __noinline__ inline

#if !defined(__CUDA_ARCH__)
#include <string.h>
#include <algorithm>
// host stand-ins for the device intrinsics used below (compile host code with -ffp-contract=off)
template <typename T> static inline T zb_ldg(const T *p) { return *p; }
static inline float zb_uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t zb_float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int zb_float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float zb_int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float zb_fmul(float a, float b) { volatile float r = a*b; return r; }
static inline float zb_fadd(float a, float b) { volatile float r = a + b; return r; }
static inline float zb_fsub(float a, float b) { volatile float r = a - b; return r; }
static inline int zb_atomic_or(int *p, int v) { int o = *p; *p |= v; return o; }
static inline uint32_t zb_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t zb_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline void zb_sincos_fast(float x, float *s, float *c) { *s = sinf(x); *c = cosf(x); }
static inline float zb_rsqrt(float x) { return 1.0f/sqrtf(x); }
static inline float zb_exp_fast(float x) { return expf(x); }
static inline float zb_rcp_fast(float x) { return 1.0f/x; }
static inline float zb_sqrt_fast(float x) { return sqrtf(x); }
static inline float zb_rsqrt_fast(float x) { return 1.0f/sqrtf(x); }
#else
#define zb_ldg __ldg
#define zb_uint_as_float __uint_as_float
#define zb_float_as_uint __float_as_uint
#define zb_float_as_int __float_as_int
#define zb_int_as_float __int_as_float
#define zb_fmul __fmul_rn
#define zb_fadd __fadd_rn
#define zb_fsub __fsub_rn
#define zb_atomic_or atomicOr
#define zb_umin min
#define zb_umax max
#define zb_sincos_fast(x, s, c) __sincosf((x), (s), (c))
#define zb_rsqrt(x) rsqrtf(x)
#define zb_exp_fast(x) __expf(x)
// 2-ulp reciprocal / square root for PRUNING bounds only (they carry explicit slack); values a query returns never
// go through these
// (single MUFU instructions: the library forms rsqrtf / __fdividef add denormal and range handling these operands —
// distances and cosines of order 1e-6 .. 1e3 — never need)
static __device__ __forceinline__ float zb_rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
static __device__ __forceinline__ float zb_sqrt_fast(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
static __device__ __forceinline__ float zb_rsqrt_fast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif

namespace zb {

constexpr float kMaxFloat = 3.402823466e+38f;
constexpr float kEpsilon = 1.192092896e-07f; // fcpw::epsilon = numeric_limits<float>::epsilon()
constexpr float kPiF = 3.14159265358979323846f;
constexpr float kHalfPiF = 1.57079632679489661923f;

template <int D>
struct Vec {
    float v[D];
    ZB_HOSTDEV float& operator[](int i) { return v[i]; }
    ZB_HOSTDEV const float& operator[](int i) const { return v[i]; }
};

template <int D> ZB_HOSTDEV Vec<D> vzero() { Vec<D> r; for (int i = 0; i < D; i++) r.v[i] = 0.0f; return r; }
template <int D> ZB_HOSTDEV Vec<D> operator+(const Vec<D>& a, const Vec<D>& b) { Vec<D> r; for (int i = 0; i < D; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
template <int D> ZB_HOSTDEV Vec<D> operator-(const Vec<D>& a, const Vec<D>& b) { Vec<D> r; for (int i = 0; i < D; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
template <int D> ZB_HOSTDEV Vec<D> operator-(const Vec<D>& a) { Vec<D> r; for (int i = 0; i < D; i++) r.v[i] = -a.v[i]; return r; }
template <int D> ZB_HOSTDEV Vec<D> operator*(const Vec<D>& a, float s) { Vec<D> r; for (int i = 0; i < D; i++) r.v[i] = a.v[i]*s; return r; }
template <int D> ZB_HOSTDEV Vec<D> operator*(float s, const Vec<D>& a) { return a*s; }
template <int D> ZB_HOSTDEV Vec<D> operator/(const Vec<D>& a, float s) { Vec<D> r; for (int i = 0; i < D; i++) r.v[i] = a.v[i]/s; return r; }
template <int D> ZB_HOSTDEV float dot(const Vec<D>& a, const Vec<D>& b) { float s = a.v[0]*b.v[0]; for (int i = 1; i < D; i++) s += a.v[i]*b.v[i]; return s; }
template <int D> ZB_HOSTDEV float norm2(const Vec<D>& a) { return dot(a, a); }
template <int D> ZB_HOSTDEV float norm(const Vec<D>& a) { return sqrtf(dot(a, a)); }
ZB_HOSTDEV Vec<3> cross(const Vec<3>& a, const Vec<3>& b) {
    Vec<3> r;
    r.v[0] = a.v[1]*b.v[2] - a.v[2]*b.v[1];
    r.v[1] = a.v[2]*b.v[0] - a.v[0]*b.v[2];
    r.v[2] = a.v[0]*b.v[1] - a.v[1]*b.v[0];
    return r;
}
ZB_HOSTDEV float cross2(const Vec<2>& a, const Vec<2>& b) { return a.v[0]*b.v[1] - a.v[1]*b.v[0]; }
ZB_HOSTDEV Vec<3> make3(float x, float y, float z) { Vec<3> r; r.v[0] = x; r.v[1] = y; r.v[2] = z; return r; }
ZB_HOSTDEV Vec<2> make2(float x, float y) { Vec<2> r; r.v[0] = x; r.v[1] = y; return r; }

template <int D> ZB_HOSTDEV Vec<D> loadVec(const float *p) { Vec<D> r; for (int i = 0; i < D; i++) r.v[i] = p[i]; return r; }
template <int D> ZB_HOSTDEV void storeVec(float *p, const Vec<D>& a) { for (int i = 0; i < D; i++) p[i] = a.v[i]; }

} // namespace zb
