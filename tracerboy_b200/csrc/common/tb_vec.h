// tb_vec.h — pinned definitions of the HLSL vector intrinsics (SURVEY §8c trap 18).
// Same role as tb_math.h: a numeric specification shared by the CUDA kernels and the
// CPU oracle so that "dot", "normalize", "reflect" ... mean one sequence of IEEE
// binary32 operations everywhere. No implicit contraction: compile with
// -ffp-contract=off (g++) / -fmad=false (nvcc).
#ifndef TB_VEC_H
#define TB_VEC_H
#include "tb_math.h"

namespace tbm {

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

TB_HD f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
TB_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
TB_HD f3 mk3(float s) { return mk3(s, s, s); }
TB_HD f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
TB_HD f4 mk4(f3 v, float w) { return mk4(v.x, v.y, v.z, w); }

TB_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
TB_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
TB_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
TB_HD f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
TB_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
TB_HD f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
TB_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
TB_HD f3 operator+(f3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
TB_HD f3 operator+(float s, f3 a) { return mk3(s + a.x, s + a.y, s + a.z); }
TB_HD f3 operator-(f3 a, float s) { return mk3(a.x - s, a.y - s, a.z - s); }
TB_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
TB_HD f3& operator+=(f3& a, f3 b) { a = a + b; return a; }
TB_HD f3& operator*=(f3& a, f3 b) { a = a * b; return a; }
TB_HD f3& operator*=(f3& a, float s) { a = a * s; return a; }
TB_HD f3& operator/=(f3& a, float s) { a = a / s; return a; }

TB_HD f2 operator+(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
TB_HD f2 operator-(f2 a, f2 b) { return mk2(a.x - b.x, a.y - b.y); }
TB_HD f2 operator*(f2 a, f2 b) { return mk2(a.x * b.x, a.y * b.y); }
TB_HD f2 operator*(f2 a, float s) { return mk2(a.x * s, a.y * s); }
TB_HD f2 operator*(float s, f2 a) { return mk2(s * a.x, s * a.y); }
TB_HD f2 operator/(f2 a, f2 b) { return mk2(a.x / b.x, a.y / b.y); }

TB_HD f4 operator+(f4 a, f4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
TB_HD f4 operator*(f4 a, f4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
TB_HD f4 operator*(f4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }

// dot3 = ((x*x)+(y*y))+(z*z), unfused
TB_HD float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
TB_HD f3 cross(f3 a, f3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
TB_HD float length(f3 a) { return sqrtf(dot(a, a)); }
TB_HD f3 normalize(f3 a) { float l = sqrtf(dot(a, a)); return mk3(a.x / l, a.y / l, a.z / l); }
TB_HD f3 reflect(f3 i, f3 n) { float d2 = 2.0f * dot(i, n); return mk3(i.x - d2 * n.x, i.y - d2 * n.y, i.z - d2 * n.z); }
TB_HD f3 lerp(f3 a, f3 b, float s) { return mk3(lerp(a.x, b.x, s), lerp(a.y, b.y, s), lerp(a.z, b.z, s)); }
TB_HD f3 min3(f3 a, f3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
TB_HD f3 max3(f3 a, f3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
TB_HD f3 min3(f3 a, float s) { return mk3(fminf(a.x, s), fminf(a.y, s), fminf(a.z, s)); }
TB_HD f3 abs3(f3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
TB_HD f3 exp3(f3 a) { return mk3(exp_(a.x), exp_(a.y), exp_(a.z)); }
TB_HD f3 pow3(f3 a, float e) { return mk3(pow_(a.x, e), pow_(a.y, e), pow_(a.z, e)); }
TB_HD f3 frac3(f3 a) { return mk3(frac(a.x), frac(a.y), frac(a.z)); }
TB_HD f2 frac2(f2 a) { return mk2(frac(a.x), frac(a.y)); }
TB_HD float comp(f3 v, int i) {
#if defined(__CUDA_ARCH__)
    // two predicated selects; the plain ternary is compiled to divergent branch regions by nvcc 12.9
    float r;
    asm("{\n\t.reg .pred p0, p1;\n\tsetp.eq.s32 p0, %4, 0;\n\tsetp.eq.s32 p1, %4, 1;\n\tselp.f32 %0, %2, %3, p1;\n\tselp.f32 %0, %1, %0, p0;\n\t}"
        : "=f"(r) : "f"(v.x), "f"(v.y), "f"(v.z), "r"(i));
    return r;
#else
    return i == 0 ? v.x : (i == 1 ? v.y : v.z);
#endif
}

} // namespace tbm
#endif
