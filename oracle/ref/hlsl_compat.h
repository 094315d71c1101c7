// hlsl_compat.h — CPU ORACLE (test infrastructure): the small HLSL/GLSL surface that the
// reference's kernel.glsl needs in order to compile as host C++ (SURVEY §8c "recommended
// construction"). Vector types with the swizzles the shader uses, and every intrinsic mapped
// onto the pinned definitions of tb_math.h / tb_vec.h, so that the compiled reference text and
// the hand-restated core (core.cpp) are comparable bit for bit.
#pragma once
#include <cstdint>
#include "../../tracerboy_b200/csrc/common/tb_vec.h"

namespace refcore {

typedef uint32_t uint;

struct float2 {
    union { struct { float x, y; }; struct { float r, g; }; };
    float2() : x(0), y(0) {}
    float2(float s) : x(s), y(s) {}
    float2(float x_, float y_) : x(x_), y(y_) {}
    float2(tbm::f2 v) : x(v.x), y(v.y) {}
    tbm::f2 t() const { return tbm::mk2(x, y); }
    float2 xy() const { return *this; }
};
struct float4;
struct float3 {
    union { struct { float x, y, z; }; struct { float r, g, b; }; };
    float3() : x(0), y(0), z(0) {}
#if defined(RC_POST) || defined(RC_TEMPORAL) || defined(RC_MATERIAL)
    float3(const float4& v); // HLSL implicit truncation float4 -> float3 (PostProcessCS.hlsl:26, 69, 118; TemporalAccumulationCS.hlsl:106, 184)
#endif
#ifdef RC_TEMPORAL
    float2 rg() const { return float2(x, y); } // PrevMomentData.rg (TemporalAccumulationCS.hlsl:222)
#endif
    float3(float s) : x(s), y(s), z(s) {}
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float3(float2 a, float z_) : x(a.x), y(a.y), z(z_) {}
    float3(tbm::f3 v) : x(v.x), y(v.y), z(v.z) {}
    tbm::f3 t() const { return tbm::mk3(x, y, z); }
    float2 xy() const { return float2(x, y); }
    float3 xyz() const { return *this; }
    float3 rgb() const { return *this; }
    float3 yzx() const { return float3(y, z, x); }
#ifdef RC_LOAD
    float2 yz() const { return float2(y, z); }                        // tri.v1.yz (RayTracingHlslCompat.h:113)
#endif
#ifdef RC_TRAVERSE
    float3(float x_, float2 yz) : x(x_), y(yz.x), z(yz.y) {}          // float3(a.w, b.xy) (RayTracingHelper.hlsli:223)
    void set_xy(float2 v) { x = v.x; y = v.y; }                      // `A.xy = ...` (TraverseFunction.hlsli:257-259)
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); } // `v[swizzleOrder.x]` (:225)
#endif
};
struct float4 {
    union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; };
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float s) : x(s), y(s), z(s), w(s) {}
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(float3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
#ifdef RC_LOAD
    float4(float2 a, float2 b) : x(a.x), y(a.y), z(b.x), w(b.y) {}  // float4(tri.v1.yz, tri.v2.xy) (RayTracingHlslCompat.h:113)
#endif
    float2 xy() const { return float2(x, y); }
    float3 xyz() const { return float3(x, y, z); }
    float3 rgb() const { return float3(x, y, z); }
#ifdef RC_FRAME
    float2 zw() const { return float2(z, w); }                       // BlueNoise0Texture[...].zw (RayGenCommon.h:116)
#endif
#ifdef RC_MATERIAL
    void set_rgb(float3 v) { x = v.x; y = v.y; z = v.z; }           // `data.rgb = ...` (SharedRaytracing.h:110)
#endif
};
#if defined(RC_POST) || defined(RC_TEMPORAL) || defined(RC_MATERIAL)
inline float3::float3(const float4& v) : x(v.x), y(v.y), z(v.z) {}
#endif
typedef float2 vec2;
typedef float3 vec3;
typedef float4 vec4;
struct bool3 { bool x, y, z; operator bool() const { return x; } }; // HLSL truncates bool3 -> bool to .x

#define RC_OPS(T, EXPR)                                                                       \
    inline T operator+(T a, T b) { return EXPR(+); }                                          \
    inline T operator-(T a, T b) { return EXPR(-); }                                          \
    inline T operator*(T a, T b) { return EXPR(*); }                                          \
    inline T operator/(T a, T b) { return EXPR(/); }                                          \
    inline T operator+(T a, float s) { return a + T(s); }                                     \
    inline T operator-(T a, float s) { return a - T(s); }                                     \
    inline T operator*(T a, float s) { return a * T(s); }                                     \
    inline T operator/(T a, float s) { return a / T(s); }                                     \
    inline T operator+(float s, T a) { return T(s) + a; }                                     \
    inline T operator-(float s, T a) { return T(s) - a; }                                     \
    inline T operator*(float s, T a) { return T(s) * a; }                                     \
    inline T operator/(float s, T a) { return T(s) / a; }                                     \
    inline T& operator+=(T& a, T b) { a = a + b; return a; }                                  \
    inline T& operator-=(T& a, T b) { a = a - b; return a; }                                  \
    inline T& operator*=(T& a, T b) { a = a * b; return a; }                                  \
    inline T& operator/=(T& a, T b) { a = a / b; return a; }                                  \
    inline T& operator*=(T& a, float s) { a = a * s; return a; }                              \
    inline T& operator/=(T& a, float s) { a = a / s; return a; }
#define RC_E2(op) float2(a.x op b.x, a.y op b.y)
#define RC_E3(op) float3(a.x op b.x, a.y op b.y, a.z op b.z)
#define RC_E4(op) float4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w)
RC_OPS(float2, RC_E2)
RC_OPS(float3, RC_E3)
RC_OPS(float4, RC_E4)
inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline bool3 operator<(float3 a, float s) { return bool3{a.x < s, a.y < s, a.z < s}; }

// ---- intrinsics, pinned (SURVEY §8c trap 18)
inline float sin(float x) { return tbm::sin_(x); }
inline float cos(float x) { return tbm::cos_(x); }
inline float acos(float x) { return tbm::acos_(x); }
inline float exp(float x) { return tbm::exp_(x); }
inline float log(float x) { return tbm::log_(x); }
inline float pow(float x, float y) { return tbm::pow_(x, y); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float abs(float x) { return ::fabsf(x); }
inline float floor(float x) { return ::floorf(x); }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline float clamp(float x, float lo, float hi) { return tbm::clamp_(x, lo, hi); }
inline float saturate(float x) { return tbm::saturate(x); }
inline float frac(float x) { return tbm::frac(x); }
inline float lerp(float a, float b, float s) { return tbm::lerp(a, b, s); }
inline float3 abs(float3 v) { return float3(abs(v.x), abs(v.y), abs(v.z)); }
inline float3 exp(float3 v) { return float3(exp(v.x), exp(v.y), exp(v.z)); }
inline float3 pow(float3 v, float e) { return float3(pow(v.x, e), pow(v.y, e), pow(v.z, e)); }
inline float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float3 min(float3 a, float s) { return min(a, float3(s)); }
inline float3 max(float3 a, float s) { return max(a, float3(s)); }
inline float3 frac(float3 v) { return float3(frac(v.x), frac(v.y), frac(v.z)); }
inline float3 lerp(float3 a, float3 b, float s) { return float3(lerp(a.x, b.x, s), lerp(a.y, b.y, s), lerp(a.z, b.z, s)); }
inline float dot(float3 a, float3 b) { return tbm::dot(a.t(), b.t()); }
inline float3 cross(float3 a, float3 b) { return float3(tbm::cross(a.t(), b.t())); }
inline float length(float3 a) { return tbm::length(a.t()); }
inline float3 normalize(float3 a) { return float3(tbm::normalize(a.t())); }
inline float3 reflect(float3 i, float3 n) { return float3(tbm::reflect(i.t(), n.t())); }
inline bool any(float3 v) { return v.x != 0.0f || v.y != 0.0f || v.z != 0.0f; }
#define fract(f) frac(f)                 // GLSLCompat.h:5
#define mix(x0, x1, a) lerp(x0, x1, a)   // GLSLCompat.h:6

// row-major 3x3 built from 9 scalars (float3x3 constructor order); mul(v, M) = v.x*row0 + v.y*row1 + v.z*row2
struct mat3 {
    float3 r0, r1, r2;
    mat3(float a, float b, float c, float d, float e, float f, float g, float h, float i) : r0(a, b, c), r1(d, e, f), r2(g, h, i) {}
};
inline float3 mul(float3 v, const mat3& m) { return (v.x * m.r0 + v.y * m.r1) + v.z * m.r2; }

} // namespace refcore
