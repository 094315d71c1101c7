// traverse.cpp — CPU ORACLE (test infrastructure): restatement of the software ray query
// (D3D12RaytracingFallback/src/TraverseFunction.hlsli:203-313, 316-429, 473-495, 537-779)
// with FAST_PATH=1, DISABLE_ANYHIT, DISABLE_PROCEDURAL_GEOMETRY (RayGenCommon.h:355-362).
//
// Pinned semantics: rcp(x) = 1/x; the slab test's a*b+-c are single fmaf;
// the watertight triangle test is evaluated unfused (`precise`); the traversal stack is
// unbounded (reference: 16 entries, unchecked); on exactly equal t the hit with the lower
// (geometryIndex, primitiveIndex) wins (reference: first found).
//
// Pinned against the reference's own text: the three pure functions (RayBoxTest, GetRayData, RayTriangleIntersect) and
// the whole loop (Traverse / SoftwareRayQuery / TestLeafNodeIntersections) are compiled from the mount
// (oracle/ref/ref_traverse_*.cpp -> oracle/_ref/libref_traverse.so) and tests/test_cpu_oracle.py requires this file, in
// its literal mode, to produce bit-identical hit records and counters on the oracle's own BVH bytes.
#include <cfloat>
#include <cstring>
#include <vector>
#include "../tracerboy_b200/csrc/common/tb_vec.h"
#include "oracle.h"

using namespace tbm;

namespace oracle {

namespace {
struct AABBNode { float c[3]; uint32_t flags; float h[3]; uint32_t right; };

// RayBoxTest, TraverseFunction.hlsli:203-221
inline bool zero_axis_inside(float c, float h, float o) { // D6: |c - o| <= h + 1e-5 (|c| + h)
    float t = fabsf(c) + h;
    float tol = h + 1.0e-5f * t;
    return fabsf(c - o) <= tol;
}
inline bool ray_box(float& resultT, float closestT, f3 org, int zmask, f3 oinv, f3 inv, const AABBNode& b) {
    f3 ainv = abs3(inv);
    float rx = fmaf(b.c[0], inv.x, -oinv.x), ry = fmaf(b.c[1], inv.y, -oinv.y), rz = fmaf(b.c[2], inv.z, -oinv.z);
    float maxx = fmaf(b.h[0], ainv.x, rx), maxy = fmaf(b.h[1], ainv.y, ry), maxz = fmaf(b.h[2], ainv.z, rz);
    float minx = fmaf(-b.h[0], ainv.x, rx), miny = fmaf(-b.h[1], ainv.y, ry), minz = fmaf(-b.h[2], ainv.z, rz);
    if (zmask) {
        const float inf = as_float(0x7f800000u);
        bool inside = true;
        if (zmask & 1) { minx = -inf; maxx = inf; inside = inside && zero_axis_inside(b.c[0], b.h[0], org.x); }
        if (zmask & 2) { miny = -inf; maxy = inf; inside = inside && zero_axis_inside(b.c[1], b.h[1], org.y); }
        if (zmask & 4) { minz = -inf; maxz = inf; inside = inside && zero_axis_inside(b.c[2], b.h[2], org.z); }
        float minTz = fmaxf(fmaxf(minx, miny), minz);
        float maxTz = fminf(fminf(maxx, maxy), maxz);
        resultT = fmaxf(minTz, 0.0f);
        return inside && fmaxf(minTz, 0.0f) < fminf(maxTz, closestT);
    }
    float minT = fmaxf(fmaxf(minx, miny), minz);
    float maxT = fminf(fminf(maxx, maxy), maxz);
    resultT = fmaxf(minT, 0.0f);
    return fmaxf(minT, 0.0f) < fminf(maxT, closestT);
}

// RayTriangleIntersect, TraverseFunction.hlsli:231-313 (RAY_FLAG_NONE, instanceFlags 0 => two-sided)
inline bool ray_tri(float& hitT, float& b1, float& b2, f3 org, int kx, int ky, int kz, f3 shear,
                    const float* v) {
    f3 v0 = mk3(v[0], v[1], v[2]) - org, v1 = mk3(v[3], v[4], v[5]) - org, v2 = mk3(v[6], v[7], v[8]) - org;
    float Ax = comp(v0, kx), Ay = comp(v0, ky), Az = comp(v0, kz);
    float Bx = comp(v1, kx), By = comp(v1, ky), Bz = comp(v1, kz);
    float Cx = comp(v2, kx), Cy = comp(v2, ky), Cz = comp(v2, kz);
    Ax = Ax - shear.x * Az; Ay = Ay - shear.y * Az;
    Bx = Bx - shear.x * Bz; By = By - shear.y * Bz;
    Cx = Cx - shear.x * Cz; Cy = Cy - shear.y * Cz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    float det = (U + V) + W;
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    if (det == 0.0f) return false;
    Az = shear.z * Az; Bz = shear.z * Bz; Cz = shear.z * Cz;
    float T = (U * Az + V * Bz) + W * Cz;
    float signCorrectedT = fabsf(T);
    if ((T > 0.0f) != (det > 0.0f)) signCorrectedT = -signCorrectedT;
    if (signCorrectedT < 0.0f || signCorrectedT > hitT * fabsf(det)) return false;
    float rcpDet = 1.0f / det;
    b1 = V * rcpDet;
    b2 = W * rcpDet;
    hitT = T * rcpDet;
    return true;
}
} // namespace

// test / analysis hook: when set, trace_ray appends one byte per node visit (0 internal, 1 leaf) for the calling thread
static thread_local std::vector<uint8_t>* g_visitLog = nullptr;
void set_visit_log(std::vector<uint8_t>* log) { g_visitLog = log; }
static bool g_literalRcp = false; // D6 off: literal rcp(0) = inf
static bool g_literalNaN = false; // D7 off: NaN rays walk the tree
void set_literal_rcp(bool on) { g_literalRcp = on; g_literalNaN = on; }
void set_literal_mode(int mask) { g_literalRcp = (mask & 1) != 0; g_literalNaN = (mask & 2) != 0; }

void trace_ray(const Scene& s, const TbRay& ray, TbHit& hit) {
    const uint8_t* bvh = s.bvh.data();
    uint32_t header[4];
    memcpy(header, bvh, 16);
    const AABBNode* nodes = (const AABBNode*)(bvh + header[0]);
    const uint8_t* prims = bvh + header[1];
    const uint32_t* meta = (const uint32_t*)(bvh + header[2]);

    f3 org = mk3(ray.Origin[0], ray.Origin[1], ray.Origin[2]);
    f3 dir = mk3(ray.Direction[0], ray.Direction[1], ray.Direction[2]);
    // Deviation D7: a ray with a NaN origin or direction component can never commit a hit (every
    // U/V/W of the watertight test is NaN, so t0 is NaN and `t0 < resultT` is false), but literally its NaN
    // slabs are dropped by min/max, so it walks every node overlapping the remaining axes — the whole tree
    // (2.7 M box tests on the vw-van scene) when the direction is all NaN, which the glass walk produces a
    // few times per million paths. Such a ray is reported as the miss it is, with zero tests.
    if (!g_literalNaN && (org.x != org.x || org.y != org.y || org.z != org.z || dir.x != dir.x || dir.y != dir.y || dir.z != dir.z)) {
        hit.TrianglesTested = hit.BoxesTested = 0; hit.InstanceIndex = 0;
        hit.t = -1.0f; hit.b1 = hit.b2 = 0; hit.PrimitiveIndex = hit.GeometryIndex = 0xffffffffu;
        return;
    }
    // GetRayData, TraverseFunction.hlsli:473-495
    f3 inv = mk3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    // Deviation D6: exactly-zero direction components. Literally, rcp(0) = inf turns
    // that axis' slab arithmetic into NaN, which min/max drop: the axis is ignored and the ray
    // walks every node overlapping the other two slabs. The reference's rand() returns exactly 0
    // about once in 300 draws, so such rays are common. Here a zero axis is tested by containment
    // of the ray's (constant) coordinate in the box, widened by 1e-5 relative so that every node
    // the literal form needs for its hits is still visited; hits are identical
    // (tests/test_cpu_oracle.py::test_zero_axis_rule_keeps_hits_and_radiance), visits drop sharply.
    int zmask = 0;
    if (!g_literalRcp) zmask = (dir.x == 0.0f ? 1 : 0) | (dir.y == 0.0f ? 2 : 0) | (dir.z == 0.0f ? 4 : 0);
    f3 oinv = org * inv;
    f3 ad = abs3(dir);
    int kz = (ad.x > ad.y && ad.x > ad.z) ? 0 : (ad.y > ad.z ? 1 : 2);
    int kx = (kz + 1) % 3, ky = (kz + 2) % 3;
    if (comp(dir, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
    f3 shear = mk3(comp(dir, kx) / comp(dir, kz), comp(dir, ky) / comp(dir, kz), 1.0f / comp(dir, kz));

    float committedT = ray.TMax;
    bool haveHit = false;
    uint32_t hitGeom = 0, hitPrim = 0;
    float hb1 = 0, hb2 = 0;
    uint32_t trisTested = 0, boxesTested = 0;

    static thread_local std::vector<uint32_t> stack; // unbounded (reference: 16, unchecked)
    stack.clear();
    float unusedT;
    if (ray_box(unusedT, committedT, org, zmask, oinv, inv, nodes[0])) stack.push_back(0);
    while (!stack.empty()) {
        uint32_t ni = stack.back();
        stack.pop_back();
        const AABBNode& nd = nodes[ni];
        if (g_visitLog) g_visitLog->push_back((nd.flags & 0x80000000u) ? 1 : 0);
        if (nd.flags & 0x80000000u) {
            uint32_t leaf = nd.flags & 0x3fffffffu;
            const uint32_t* m = meta + 3 * (size_t)leaf;
            trisTested++;
            float t0 = committedT, b1, b2;
            const float* v = (const float*)(prims + 40 * (size_t)leaf + 4);
            bool ok = ray_tri(t0, b1, b2, org, kx, ky, kz, shear, v);
            // TestLeafNodeIntersections :420 plus the equal-t tie-break
            bool closer = t0 < committedT;
            bool tie = haveHit && t0 == committedT && (m[0] < hitGeom || (m[0] == hitGeom && m[1] < hitPrim));
            if (ok && (closer || tie) && t0 > ray.TMin) {
                committedT = t0; hb1 = b1; hb2 = b2; hitGeom = m[0]; hitPrim = m[1]; haveHit = true;
            }
        } else {
            uint32_t l = nd.flags & 0x3fffffffu, r = nd.right;
            float lt, rt;
            bool lh = ray_box(lt, committedT, org, zmask, oinv, inv, nodes[l]);
            bool rh = ray_box(rt, committedT, org, zmask, oinv, inv, nodes[r]);
            boxesTested += 2;
            if (lh && rh) { // far first, near last; on equal t left is near (:754-765)
                bool rightFirst = rt < lt;
                stack.push_back(rightFirst ? l : r);
                stack.push_back(rightFirst ? r : l);
            } else if (lh || rh) stack.push_back(rh ? r : l);
        }
    }
    hit.TrianglesTested = trisTested;
    hit.BoxesTested = boxesTested;
    hit.InstanceIndex = 0;
    if (haveHit && committedT < ray.TMax) {
        hit.t = committedT; hit.b1 = hb1; hit.b2 = hb2; hit.PrimitiveIndex = hitPrim; hit.GeometryIndex = hitGeom;
    } else {
        hit.t = -1.0f; hit.b1 = hit.b2 = 0; hit.PrimitiveIndex = hit.GeometryIndex = 0xffffffffu;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-level ray query: Traverse with FAST_PATH 0 (TraverseFunction.hlsli:537-785). The top-level tree is walked with the
// world-space ray; at an instance leaf (:603-638) whose mask passes, the ray is taken to object space with the stored
// WorldToObject (origin as a point, direction as a vector: t keeps its meaning), GetRayData is redone for it, and the
// instance's bottom-level tree is walked from its root WITHOUT a root box test (the instance's world box stood for it)
// until it is exhausted; then the walk returns to the top level with the world-space ray data (:770-774). One
// committed t, both counters run across the levels. D3 extends to (instance, geometry, primitive): on exactly equal t
// the lower triple wins. D6 / D7 apply to each level's ray as in the single-level query.
namespace {
struct RayData { f3 org, inv, oinv, shear; int kx, ky, kz, zmask; };
inline RayData ray_data(f3 org, f3 dir) {
    RayData r;
    r.org = org;
    r.inv = mk3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    r.zmask = g_literalRcp ? 0 : ((dir.x == 0.0f ? 1 : 0) | (dir.y == 0.0f ? 2 : 0) | (dir.z == 0.0f ? 4 : 0));
    r.oinv = org * r.inv;
    f3 ad = abs3(dir);
    r.kz = (ad.x > ad.y && ad.x > ad.z) ? 0 : (ad.y > ad.z ? 1 : 2);
    r.kx = (r.kz + 1) % 3; r.ky = (r.kz + 2) % 3;
    if (comp(dir, r.kz) < 0.0f) { int t = r.kx; r.kx = r.ky; r.ky = t; }
    r.shear = mk3(comp(dir, r.kx) / comp(dir, r.kz), comp(dir, r.ky) / comp(dir, r.kz), 1.0f / comp(dir, r.kz));
    return r;
}
inline bool is_nan3(f3 v) { return v.x != v.x || v.y != v.y || v.z != v.z; }
struct TlasMeta { float worldToObject[12]; uint32_t idAndMask, contribAndFlags, asLo, asHi; float objectToWorld[12]; uint32_t instanceIndex; };
} // namespace

void trace_ray_tlas(const uint8_t* tlas, const std::vector<const uint8_t*>& blasList, const TbRay& ray, TbHit& hit) {
    uint32_t th[4];
    memcpy(th, tlas, 16);
    const AABBNode* tnodes = (const AABBNode*)(tlas + th[0]);
    const uint8_t* tmeta = tlas + th[1]; // GetOffsetToInstanceDesc: OffsetToLeafNodeMetaDataOffset = 4
    f3 worg = mk3(ray.Origin[0], ray.Origin[1], ray.Origin[2]), wdir = mk3(ray.Direction[0], ray.Direction[1], ray.Direction[2]);
    hit.TrianglesTested = hit.BoxesTested = 0; hit.InstanceIndex = 0;
    hit.t = -1.0f; hit.b1 = hit.b2 = 0; hit.PrimitiveIndex = hit.GeometryIndex = 0xffffffffu;
    if (!g_literalNaN && (is_nan3(worg) || is_nan3(wdir))) return; // D7
    const RayData W = ray_data(worg, wdir);
    float committedT = ray.TMax;
    bool haveHit = false;
    uint32_t hitInst = 0, hitGeom = 0, hitPrim = 0;
    float hb1 = 0, hb2 = 0;
    uint32_t trisTested = 0, boxesTested = 0;
    std::vector<uint32_t> tstack, bstack;
    float unusedT;
    if (ray_box(unusedT, committedT, W.org, W.zmask, W.oinv, W.inv, tnodes[0])) tstack.push_back(0);
    while (!tstack.empty()) {
        const uint32_t ni = tstack.back();
        tstack.pop_back();
        const AABBNode& nd = tnodes[ni];
        if (nd.flags & 0x80000000u) {
            TlasMeta m;
            memcpy(&m, tmeta + 116 * (size_t)(nd.flags & 0x3fffffffu), 116);
            if (((m.idAndMask >> 24) & 0xffu) == 0) continue; // GetInstanceMask & InstanceInclusionMask (~0)
            const float* w = m.worldToObject;
            f3 oorg = mk3(((w[0] * worg.x + w[1] * worg.y) + w[2] * worg.z) + w[3] * 1.0f,
                          ((w[4] * worg.x + w[5] * worg.y) + w[6] * worg.z) + w[7] * 1.0f,
                          ((w[8] * worg.x + w[9] * worg.y) + w[10] * worg.z) + w[11] * 1.0f);
            f3 odir = mk3(((w[0] * wdir.x + w[1] * wdir.y) + w[2] * wdir.z) + w[3] * 0.0f,
                          ((w[4] * wdir.x + w[5] * wdir.y) + w[6] * wdir.z) + w[7] * 0.0f,
                          ((w[8] * wdir.x + w[9] * wdir.y) + w[10] * wdir.z) + w[11] * 0.0f);
            if (!g_literalNaN && (is_nan3(oorg) || is_nan3(odir))) continue; // D7 for the object-space ray
            const RayData O = ray_data(oorg, odir);
            const uint8_t* blas = blasList[m.asLo];
            uint32_t bh[4];
            memcpy(bh, blas, 16);
            const AABBNode* nodes = (const AABBNode*)(blas + bh[0]);
            const uint8_t* prims = blas + bh[1];
            const uint32_t* meta = (const uint32_t*)(blas + bh[2]);
            bstack.clear();
            bstack.push_back(0); // :621 StackPush(0): no box test of the bottom-level root
            while (!bstack.empty()) {
                const uint32_t bi = bstack.back();
                bstack.pop_back();
                const AABBNode& bn = nodes[bi];
                if (bn.flags & 0x80000000u) {
                    const uint32_t leaf = bn.flags & 0x3fffffffu;
                    const uint32_t* pm = meta + 3 * (size_t)leaf;
                    trisTested++;
                    float t0 = committedT, b1, b2;
                    const bool ok = ray_tri(t0, b1, b2, O.org, O.kx, O.ky, O.kz, O.shear, (const float*)(prims + 40 * (size_t)leaf + 4));
                    const bool closer = t0 < committedT;
                    const bool lower = m.instanceIndex < hitInst || (m.instanceIndex == hitInst && (pm[0] < hitGeom || (pm[0] == hitGeom && pm[1] < hitPrim)));
                    const bool tie = haveHit && t0 == committedT && lower;
                    if (ok && (closer || tie) && t0 > ray.TMin) {
                        committedT = t0; hb1 = b1; hb2 = b2; hitInst = m.instanceIndex; hitGeom = pm[0]; hitPrim = pm[1]; haveHit = true;
                    }
                } else {
                    const uint32_t l = bn.flags & 0x3fffffffu, r = bn.right;
                    float lt, rt;
                    const bool lh = ray_box(lt, committedT, O.org, O.zmask, O.oinv, O.inv, nodes[l]);
                    const bool rh = ray_box(rt, committedT, O.org, O.zmask, O.oinv, O.inv, nodes[r]);
                    boxesTested += 2;
                    if (lh && rh) { const bool rightFirst = rt < lt; bstack.push_back(rightFirst ? l : r); bstack.push_back(rightFirst ? r : l); }
                    else if (lh || rh) bstack.push_back(rh ? r : l);
                }
            }
        } else {
            const uint32_t l = nd.flags & 0x3fffffffu, r = nd.right;
            float lt, rt;
            const bool lh = ray_box(lt, committedT, W.org, W.zmask, W.oinv, W.inv, tnodes[l]);
            const bool rh = ray_box(rt, committedT, W.org, W.zmask, W.oinv, W.inv, tnodes[r]);
            boxesTested += 2;
            if (lh && rh) { const bool rightFirst = rt < lt; tstack.push_back(rightFirst ? l : r); tstack.push_back(rightFirst ? r : l); }
            else if (lh || rh) tstack.push_back(rh ? r : l);
        }
    }
    hit.TrianglesTested = trisTested;
    hit.BoxesTested = boxesTested;
    if (haveHit && committedT < ray.TMax) {
        hit.t = committedT; hit.b1 = hb1; hit.b2 = hb2; hit.PrimitiveIndex = hitPrim; hit.GeometryIndex = hitGeom; hit.InstanceIndex = hitInst;
    }
}

} // namespace oracle

// ---- test hooks: the three pure functions of the ray query, so that tests can compare them with the reference's
// own text compiled from the mount (oracle/ref/ref_traverse_*.cpp)
extern "C" __attribute__((visibility("default")))
void oracle_ray_data(const float* org, const float* dir, float* inv, float* oinv, float* shear, int* swz) {
    using namespace oracle;
    f3 o = mk3(org[0], org[1], org[2]), d = mk3(dir[0], dir[1], dir[2]);
    f3 i = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    f3 oi = o * i;
    f3 ad = abs3(d);
    int kz = (ad.x > ad.y && ad.x > ad.z) ? 0 : (ad.y > ad.z ? 1 : 2);
    int kx = (kz + 1) % 3, ky = (kz + 2) % 3;
    if (comp(d, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
    f3 sh = mk3(comp(d, kx) / comp(d, kz), comp(d, ky) / comp(d, kz), 1.0f / comp(d, kz));
    inv[0] = i.x; inv[1] = i.y; inv[2] = i.z; oinv[0] = oi.x; oinv[1] = oi.y; oinv[2] = oi.z;
    shear[0] = sh.x; shear[1] = sh.y; shear[2] = sh.z; swz[0] = kx; swz[1] = ky; swz[2] = kz;
}
extern "C" __attribute__((visibility("default")))
int oracle_ray_box(float closestT, const float* oinv, const float* inv, const float* c, const float* h, float* resultT) {
    using namespace oracle;
    AABBNode b;
    b.c[0] = c[0]; b.c[1] = c[1]; b.c[2] = c[2]; b.h[0] = h[0]; b.h[1] = h[1]; b.h[2] = h[2]; b.flags = 0; b.right = 0;
    return ray_box(*resultT, closestT, mk3(0.0f), 0, mk3(oinv[0], oinv[1], oinv[2]), mk3(inv[0], inv[1], inv[2]), b) ? 1 : 0;
}
extern "C" __attribute__((visibility("default")))
int oracle_ray_tri(float* hitT, const float* org, const int* swz, const float* shear, const float* v9, float* bary) {
    using namespace oracle;
    return ray_tri(*hitT, bary[0], bary[1], mk3(org[0], org[1], org[2]), swz[0], swz[1], swz[2], mk3(shear[0], shear[1], shear[2]), v9) ? 1 : 0;
}
