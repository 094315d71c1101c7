// tb_math.h — the pinned numeric definition of the HLSL intrinsics the tracer uses.
//
// HLSL leaves sin/cos/acos/atan2/exp/log/pow/rcp to the driver (SURVEY §8c traps 4,
// 6, 18), so "the reference's result" only exists once these are pinned. This header
// is that pin. It is written in plain IEEE-754 binary32 operations (+ - * / sqrt,
// explicit fmaf, float<->int conversions) that g++ (-ffp-contract=off) and nvcc
// (-fmad=false) evaluate identically, so the CPU oracle and the CUDA kernels get
// bit-identical values. Polynomials follow the classic Cephes single-precision
// kernels (published algorithm, Moshier 1992); error is <= ~2 ulp over the ranges the
// tracer uses. Both the product (tracerboy_b200/csrc/cuda) and the checker (oracle/)
// include this file: it is a specification shared by both, not an implementation of
// the path.
#ifndef TB_MATH_H
#define TB_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
#endif

namespace tbm {

TB_HD float as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
TB_HD uint32_t as_uint(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

TB_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
TB_HD float rcp(float x) { return 1.0f / x; }                 // HLSL rcp := IEEE 1/x
TB_HD float min_(float a, float b) { return fminf(a, b); }     // IEEE minNum/maxNum
TB_HD float max_(float a, float b) { return fmaxf(a, b); }
TB_HD float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); } // saturate(NaN)=0
TB_HD float clamp_(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
TB_HD float frac(float x) { return x - floorf(x); }
TB_HD float lerp(float a, float b, float s) { return a + s * (b - a); }
TB_HD bool isnan_(float x) { return x != x; }

// ---- sin / cos -----------------------------------------------------------
// Quadrant reduction with a 3-part pi/2 (Cody-Waite, fmaf), then Cephes sinf/cosf
// minimax kernels on [-pi/4, pi/4]. Accurate for |x| < ~1e5 (the tracer's
// arguments are < ~1e3); beyond that precision degrades gracefully.
TB_HD void sincos_reduce(float x, float& r, int& q) {
    const float TWO_OVER_PI = 0.636619772367581343f;
    const float P1 = 1.5703125f;                  // pi/2 split, 3 parts
    const float P2 = 4.837512969970703125e-4f;
    const float P3 = 7.54978995489188e-8f;
    float k = rintf(x * TWO_OVER_PI);
    r = fmaf(-k, P1, x);
    r = fmaf(-k, P2, r);
    r = fmaf(-k, P3, r);
    q = (int)k;
}
TB_HD float sin_kernel(float r) {
    float z = r * r;
    float p = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    p = fmaf(p, z, -1.6666654611e-1f);
    return fmaf(p * z, r, r);
}
TB_HD float cos_kernel(float r) {
    float z = r * r;
    float p = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    p = fmaf(p, z, 4.166664568298827e-2f);
    return fmaf(p * z, z, fmaf(-0.5f, z, 1.0f));
}
TB_HD float sin_(float x) {
    if (!(fabsf(x) <= 1.0e8f)) return as_float(0x7fc00000u); // out of pinned range / inf / nan
    float r; int q; sincos_reduce(x, r, q);
    float s = (q & 1) ? cos_kernel(r) : sin_kernel(r);
    return (q & 2) ? -s : s;
}
TB_HD float cos_(float x) {
    if (!(fabsf(x) <= 1.0e8f)) return as_float(0x7fc00000u);
    float r; int q; sincos_reduce(x, r, q);
    float c = (q & 1) ? sin_kernel(r) : cos_kernel(r);
    return ((q + 1) & 2) ? -c : c;
}

// ---- asin / acos ---------------------------------------------------------
TB_HD float asin_poly(float z, float x) { // z = x*x (or the half-angle square)
    float p = fmaf(4.2163199048e-2f, z, 2.4181311049e-2f);
    p = fmaf(p, z, 4.5470025998e-2f);
    p = fmaf(p, z, 7.4953002686e-2f);
    p = fmaf(p, z, 1.6666752422e-1f);
    return fmaf(p * z, x, x);
}
TB_HD float acos_(float x) {
    const float PI_F = 3.14159265358979323846f;
    const float PIO2_F = 1.57079632679489661923f;
    float a = fabsf(x);
    if (!(a <= 1.0f)) return as_float(0x7fc00000u); // |x|>1 or nan -> nan (HLSL acos)
    if (a > 0.5f) {
        float z = 0.5f * (1.0f - a);
        float s = sqrtf(z);
        float r = 2.0f * asin_poly(z, s);      // acos(|x|)
        return x < 0.0f ? PI_F - r : r;
    }
    return PIO2_F - asin_poly(x * x, x);
}

// ---- atan / atan2 --------------------------------------------------------
TB_HD float atan_(float xx) {
    const float PIO2_F = 1.57079632679489661923f;
    const float PIO4_F = 0.78539816339744830962f;
    float x = fabsf(xx), y;
    if (x > 2.414213562373095f) { y = PIO2_F; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = PIO4_F; x = (x - 1.0f) / (x + 1.0f); }
    else y = 0.0f;
    float z = x * x;
    float p = fmaf(8.05374449538e-2f, z, -1.38776856032e-1f);
    p = fmaf(p, z, 1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    y += fmaf(p * z, x, x);
    return xx < 0.0f ? -y : y;
}
TB_HD float atan2_(float y, float x) {
    const float PI_F = 3.14159265358979323846f;
    const float PIO2_F = 1.57079632679489661923f;
    if (x != x || y != y) return x + y;
    if (x == 0.0f) {
        if (y == 0.0f) return 0.0f;
        return y > 0.0f ? PIO2_F : -PIO2_F;
    }
    float z = atan_(y / x);
    if (x < 0.0f) z = (y < 0.0f) ? z - PI_F : z + PI_F;
    return z;
}

// ---- exp / log / pow -----------------------------------------------------
TB_HD float exp_(float x) {
    if (x != x) return x;
    if (x > 88.72283905206835f) return as_float(0x7f800000u);
    if (x < -103.972084045410f) return 0.0f;
    const float LOG2E = 1.44269504088896341f;
    const float C1 = 0.693359375f, C2 = -2.12194440e-4f;
    float n = rintf(x * LOG2E);
    float r = fmaf(-n, C1, x);
    r = fmaf(-n, C2, r);
    float z = r * r;
    float p = fmaf(1.9875691500e-4f, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float e = fmaf(p, z, r) + 1.0f;
    // scale by 2^n in two steps so that results in the denormal range stay exact-ish
    int ni = (int)n;
    int n1 = ni / 2, n2 = ni - n1;
    float s1 = as_float((uint32_t)(n1 + 127) << 23);
    float s2 = as_float((uint32_t)(n2 + 127) << 23);
    return e * s1 * s2;
}
TB_HD float log_(float x) {
    if (x != x) return x;
    if (x < 0.0f) return as_float(0x7fc00000u);
    if (x == 0.0f) return as_float(0xff800000u);
    if (x == as_float(0x7f800000u)) return x;
    int e = 0;
    uint32_t u = as_uint(x);
    if (u < 0x00800000u) { x *= 8388608.0f; u = as_uint(x); e = -23; } // denormal
    e += (int)(u >> 23) - 126;
    float m = as_float((u & 0x007fffffu) | 0x3f000000u); // [0.5,1)
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float p = fmaf(7.0376836292e-2f, m, -1.1514610310e-1f);
    p = fmaf(p, m, 1.1676998740e-1f);
    p = fmaf(p, m, -1.2420140846e-1f);
    p = fmaf(p, m, 1.4249322787e-1f);
    p = fmaf(p, m, -1.6668057665e-1f);
    p = fmaf(p, m, 2.0000714765e-1f);
    p = fmaf(p, m, -2.4999993993e-1f);
    p = fmaf(p, m, 3.3333331174e-1f);
    float y = p * m * z;
    float fe = (float)e;
    y = fmaf(-2.12194440e-4f, fe, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    return fmaf(0.693359375f, fe, r);
}
// HLSL pow(x,y) = exp2(y*log2(x)): NaN for x<0, which the tracer's call sites avoid
// (AbsPow) or rely on. Same structure here: exp(y*log(x)).
TB_HD float pow_(float x, float y) {
    if (y == 0.0f) return 1.0f;
    if (y == 2.0f) return x * x; // shader compilers fold pow(x, 2.0) to a multiply
    if (x == 0.0f) return y > 0.0f ? 0.0f : as_float(0x7f800000u);
    return exp_(y * log_(x));
}

// log2 / exp2 / smoothstep (post-processing: Tonemap.h agx + GTTonemap, GenerateHistogramCS.hlsl:27,
// CalculateAveragedLuminanceCS.hlsl:36), pinned on top of log_/exp_
TB_HD float log2_(float x) { return log_(x) * 1.44269504088896341f; }
TB_HD float exp2_(float x) { return exp_(x * 0.693147180559945309f); }
TB_HD float smoothstep_(float a, float b, float x) {
    float t = saturate((x - a) / (b - a));
    return (t * t) * (3.0f - 2.0f * t);
}

} // namespace tbm
#endif
