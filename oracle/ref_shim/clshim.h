/*
 * clshim.h — TEST INFRASTRUCTURE (oracle side).  Just enough of OpenCL C 1.2
 * expressed in C++17 for g++ to compile the reference's own kernel sources
 * (/root/reference/src/kernels/cl/*.cl and kernels/common/*.h) for the host CPU,
 * so that the reference's arithmetic itself — not a restatement — can be run
 * here.  See oracle/build_ref.py for the recipe and the (purely syntactic)
 * rewrites applied to the sources on their way into oracle/_ref/gen/.
 *
 * What the shim supplies: vector types float2/float3/float4/int2 with the
 * swizzles the kernels use (.xy .xyz, also as l-values), address-space and
 * kernel qualifiers, get_global_id, atomic_add, image2d_t + read_imagef
 * (CLK_NORMALIZED_COORDS_TRUE | CLK_ADDRESS_REPEAT | CLK_FILTER_LINEAR, the
 * OpenCL 1.2 spec formula, section 8.2) / write_imagef, and the math built-ins.
 *
 * The OpenCL driver's math library is the one un-vendored dependency of the
 * path: its results are implementation-defined at the ulp level.
 * RT_REF_LIBM selects which implementation stands in for it:
 *   RT_REF_LIBM=0 (default): include/rt_math.h — the same correctly-rounded
 *       functions the CUDA kernels and the oracle restatement use;
 *   RT_REF_LIBM=1: glibc libm (sinf, cosf, ...), to measure how sensitive the
 *       image is to that choice.
 * OpenCL leaves min/max undefined for NaN operands; fmin/fmax semantics are used.
 */
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "rt_math.h"

#ifndef RT_REF_LIBM
#define RT_REF_LIBM 0
#endif

namespace clc
{

typedef unsigned int uint;

#define __kernel
#define __global
#define __constant const
#define __local
#define __private
#define __read_only
#define __write_only

static thread_local size_t t_global_id = 0;
inline size_t get_global_id(uint) { return t_global_id; }

// ---------------------------------------------------------------- vectors
struct float2
{
    float x, y;
    float2() = default;
    float2(float a) : x(a), y(a) {}
    float2(float a, float b) : x(a), y(b) {}
};

struct int2
{
    int x, y;
    int2() = default;
    int2(int a, int b) : x(a), y(b) {}
};

struct float3;

// swizzle views: trivially copyable so they can live in the unions below
struct swz_xy
{
    float v[2];
    operator float2() const { return float2(v[0], v[1]); }
    swz_xy& operator=(const float2& o) { v[0] = o.x; v[1] = o.y; return *this; }
};

struct swz_xyz
{
    float v[3];
    inline operator float3() const;
    inline swz_xyz& operator=(const float3& o);
    inline swz_xyz& operator+=(const float3& o);
    inline swz_xyz& operator*=(const float3& o);
};

// OpenCL float3 occupies 16 bytes
struct alignas(16) float3
{
    union
    {
        struct { float x, y, z; };
        swz_xyz xyz;
        swz_xy xy;
        float v4[4];
    };
    float3() = default;
    float3(float a) : v4{ a, a, a, 0.0f } {}
    float3(float a, float b, float c) : v4{ a, b, c, 0.0f } {}
};

struct alignas(16) float4
{
    union
    {
        struct { float x, y, z, w; };
        swz_xyz xyz;
        swz_xy xy;
        float v4[4];
    };
    float4() = default;
    float4(float a) : v4{ a, a, a, a } {}
    float4(float a, float b, float c, float d) : v4{ a, b, c, d } {}
    float4(const float3& a, float d) : v4{ a.x, a.y, a.z, d } {}
    float4(const float2& a, float c, float d) : v4{ a.x, a.y, c, d } {}
};

static_assert(sizeof(float2) == 8 && sizeof(float3) == 16 && sizeof(float4) == 16, "OpenCL vector sizes");

inline swz_xyz::operator float3() const { return float3(v[0], v[1], v[2]); }
inline swz_xyz& swz_xyz::operator=(const float3& o) { v[0] = o.x; v[1] = o.y; v[2] = o.z; return *this; }
inline swz_xyz& swz_xyz::operator+=(const float3& o) { v[0] += o.x; v[1] += o.y; v[2] += o.z; return *this; }
inline swz_xyz& swz_xyz::operator*=(const float3& o) { v[0] *= o.x; v[1] *= o.y; v[2] *= o.z; return *this; }

#define CLC_BINOP3(op) \
    inline float3 operator op(float3 a, float3 b) { return float3(a.x op b.x, a.y op b.y, a.z op b.z); } \
    inline float3 operator op(float3 a, float b) { return float3(a.x op b, a.y op b, a.z op b); } \
    inline float3 operator op(float a, float3 b) { return float3(a op b.x, a op b.y, a op b.z); } \
    inline float3& operator op##=(float3& a, float3 b) { a.x op##= b.x; a.y op##= b.y; a.z op##= b.z; return a; } \
    inline float3& operator op##=(float3& a, float b) { a.x op##= b; a.y op##= b; a.z op##= b; return a; }
CLC_BINOP3(+) CLC_BINOP3(-) CLC_BINOP3(*) CLC_BINOP3(/)
#undef CLC_BINOP3
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }

#define CLC_BINOP2(op) \
    inline float2 operator op(float2 a, float2 b) { return float2(a.x op b.x, a.y op b.y); } \
    inline float2 operator op(float2 a, float b) { return float2(a.x op b, a.y op b); } \
    inline float2 operator op(float a, float2 b) { return float2(a op b.x, a op b.y); } \
    inline float2& operator op##=(float2& a, float2 b) { a.x op##= b.x; a.y op##= b.y; return a; } \
    inline float2& operator op##=(float2& a, float b) { a.x op##= b; a.y op##= b; return a; }
CLC_BINOP2(+) CLC_BINOP2(-) CLC_BINOP2(*) CLC_BINOP2(/)
#undef CLC_BINOP2

#define CLC_BINOP4(op) \
    inline float4 operator op(float4 a, float4 b) { return float4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    inline float4 operator op(float4 a, float b) { return float4(a.x op b, a.y op b, a.z op b, a.w op b); }
CLC_BINOP4(+) CLC_BINOP4(-) CLC_BINOP4(*) CLC_BINOP4(/)
#undef CLC_BINOP4

// ---------------------------------------------------------------- math built-ins
inline float sqrt(float x) { return ::sqrtf(x); }
inline double sqrt(double x) { return ::sqrt(x); }
inline float fabs(float x) { return ::fabsf(x); }
inline float floor(float x) { return ::floorf(x); }
inline float2 floor(float2 a) { return float2(::floorf(a.x), ::floorf(a.y)); }
inline float ldexp(float x, int e) { return ::ldexpf(x, e); }
#if RT_REF_LIBM
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float tan(float x) { return ::tanf(x); }
inline float acos(float x) { return ::acosf(x); }
inline float atan2(float y, float x) { return ::atan2f(y, x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float exp(float x) { return ::expf(x); }
#else
inline float sin(float x) { return rt_sinf(x); }
inline float cos(float x) { return rt_cosf(x); }
inline float tan(float x) { return rt_tanf(x); }
inline float acos(float x) { return rt_acosf(x); }
inline float atan2(float y, float x) { return rt_atan2f(y, x); }
inline float pow(float x, float y) { return rt_powf(x, y); }
inline float exp(float x) { return (float)rt_exp_d((double)x); }
#endif
inline float3 pow(float3 a, float b) { return float3(pow(a.x, b), pow(a.y, b), pow(a.z, b)); }

inline float min(float a, float b) { return rt_fminf(a, b); }
inline float max(float a, float b) { return rt_fmaxf(a, b); }
inline float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline int clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline float3 clamp(float3 v, float lo, float hi) { return float3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline float3 mix(float3 a, float3 b, float3 t) { return a + (b - a) * t; }
inline float3 mix(float3 a, float3 b, float t) { return a + (b - a) * t; }

// geometric built-ins: left-to-right sums, no fused multiply-add (the whole
// oracle is compiled with -ffp-contract=off); normalize multiplies by one
// correctly rounded reciprocal of the length.
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(float3 a) { return ::sqrtf(dot(a, a)); }
inline float3 normalize(float3 a) { float inv = 1.0f / ::sqrtf(dot(a, a)); return a * inv; }

// ---------------------------------------------------------------- atomics
inline uint atomic_add(uint* p, uint v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// ---------------------------------------------------------------- images
typedef int sampler_t;
enum { CLK_NORMALIZED_COORDS_TRUE = 1, CLK_ADDRESS_REPEAT = 2, CLK_FILTER_LINEAR = 4 };

struct image2d_t
{
    float* data;   // RGBA32F, row-major, row 0 first
    int width;
    int height;
};

// OpenCL 1.2 spec 8.2: normalized coordinates, CLK_ADDRESS_REPEAT, CLK_FILTER_LINEAR
inline float4 read_imagef(image2d_t img, sampler_t, float2 coords)
{
    float wt = (float)img.width, ht = (float)img.height;
    float u = (coords.x - ::floorf(coords.x)) * wt;
    float v = (coords.y - ::floorf(coords.y)) * ht;
    float fu = ::floorf(u - 0.5f), fv = ::floorf(v - 0.5f);
    // NaN / out-of-range guard (float->int of NaN is not portable): texel 0
    int i0 = (fu >= -1.0f && fu <= wt) ? (int)fu : 0;
    int j0 = (fv >= -1.0f && fv <= ht) ? (int)fv : 0;
    int i1 = i0 + 1, j1 = j0 + 1;
    if (i0 < 0) i0 += img.width;
    if (i1 > img.width - 1) i1 -= img.width;
    if (j0 < 0) j0 += img.height;
    if (j1 > img.height - 1) j1 -= img.height;
    float a = (u - 0.5f) - fu;
    float b = (v - 0.5f) - fv;
    const float4* t = (const float4*)img.data;
    float4 t00 = t[(size_t)j0 * img.width + i0], t10 = t[(size_t)j0 * img.width + i1];
    float4 t01 = t[(size_t)j1 * img.width + i0], t11 = t[(size_t)j1 * img.width + i1];
    return t00 * ((1.0f - a) * (1.0f - b)) + t10 * (a * (1.0f - b)) + t01 * ((1.0f - a) * b) + t11 * (a * b);
}

inline void write_imagef(image2d_t img, int2 c, float4 v)
{
    ((float4*)img.data)[(size_t)c.y * img.width + c.x] = v;
}

} // namespace clc
