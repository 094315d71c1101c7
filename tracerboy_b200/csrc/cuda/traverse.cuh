// traverse.cuh — software ray query on the PairNode/WideTri layout (device_types.h).
//
// Role: SoftwareRayQuery::TraceRayInline + Proceed (TraverseFunction.hlsli:537-785) with
// FAST_PATH=1, DISABLE_ANYHIT, DISABLE_PROCEDURAL_GEOMETRY (RayGenCommon.h:355-362):
// single-level BVH2, world-space ray, closest hit, near child first (left on equal t),
// watertight ray/triangle test (Woop/Benthin/Wald 2013, :231-313), slab box test on
// centre/half boxes (:203-221). Same visit order and the same BoxesTested/TrianglesTested
// counters as the reference; what changes is the memory layout: one 64-byte PairNode
// (4 x LDG.128) carries both child boxes and both child references, so an internal visit
// costs one dependent fetch instead of the reference's three (popped node, left, right),
// and a leaf visit is one 48-byte WideTri (3 x LDG.128) that already holds the
// geometry/primitive ids (the reference reads 40 B primitive + 12 B metadata + 2 header
// offsets).
#pragma once
#include "../common/tb_vec.h"
#include "device_types.h"

namespace tbd {

struct HitRec {
    float t, b1, b2;
    uint32_t prim, geom;
    uint32_t tris, boxes;
};

// TB_STACK_DEPTH (device_types.h) bounds the waiting far children: at most one per level of the tree, and the host
// refuses to hand out a BVH deeper than that (DeviceBvh::depth, measured by the builder), so a push never overflows.
// The memory stack is `uint32_t stack[TB_STACK_WORDS]`: word 0 is a sentinel (TB_NO_NODE, written by
// begin()/resume()), words 1..depth are the waiting far children. Popping the sentinel ends the traversal,
// so pop is one unconditional load with no emptiness test.
#define TB_STACK_WORDS (TB_STACK_DEPTH + 1)

// RayBoxTest (TraverseFunction.hlsli:203-221) as an interval: the box is hit iff enter < exit, and `enter` is the
// reference's resultT. Returned by value so that neither end ever has its address taken (a by-reference result
// shared with the out-of-line zero-axis variant used to pin both entry distances in local memory).
struct SlabRange { float enter, exit; };
__device__ __forceinline__ SlabRange slab(float closestT, tbm::f3 oinv, tbm::f3 inv, tbm::f3 ainv,
                                          float cx, float cy, float cz, float hx, float hy, float hz) {
    float rx = fmaf(cx, inv.x, -oinv.x), ry = fmaf(cy, inv.y, -oinv.y), rz = fmaf(cz, inv.z, -oinv.z);
    float maxx = fmaf(hx, ainv.x, rx), maxy = fmaf(hy, ainv.y, ry), maxz = fmaf(hz, ainv.z, rz);
    float minx = fmaf(-hx, ainv.x, rx), miny = fmaf(-hy, ainv.y, ry), minz = fmaf(-hz, ainv.z, rz);
    float minT = fmaxf(fmaxf(minx, miny), minz);
    float maxT = fminf(fminf(maxx, maxy), maxz);
    SlabRange r;
    r.enter = fmaxf(minT, 0.0f);
    r.exit = fminf(maxT, closestT);
    return r;
}

// 256-bit read-only global load (32-byte aligned)
__device__ __forceinline__ void ldg256(const float4* __restrict__ p, float4& lo, float4& hi) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "l"(p));
}

// D6: slab test for a ray with exactly-zero direction components: those axes constrain nothing in t
// and are tested by containment |c - o| <= h + 1e-5 (|c| + h).
__device__ __forceinline__ bool zero_axis_inside(float c, float h, float o) {
    float t = fabsf(c) + h;
    float tol = h + 1.0e-5f * t;
    return fabsf(c - o) <= tol;
}
__device__ __noinline__ SlabRange slab_zero(float closestT, tbm::f3 org, int zmask, tbm::f3 oinv, tbm::f3 inv, tbm::f3 ainv,
                                            float cx, float cy, float cz, float hx, float hy, float hz) {
    float rx = fmaf(cx, inv.x, -oinv.x), ry = fmaf(cy, inv.y, -oinv.y), rz = fmaf(cz, inv.z, -oinv.z);
    float maxx = fmaf(hx, ainv.x, rx), maxy = fmaf(hy, ainv.y, ry), maxz = fmaf(hz, ainv.z, rz);
    float minx = fmaf(-hx, ainv.x, rx), miny = fmaf(-hy, ainv.y, ry), minz = fmaf(-hz, ainv.z, rz);
    const float inf = __uint_as_float(0x7f800000u);
    bool inside = true;
    if (zmask & 1) { minx = -inf; maxx = inf; inside = inside && zero_axis_inside(cx, hx, org.x); }
    if (zmask & 2) { miny = -inf; maxy = inf; inside = inside && zero_axis_inside(cy, hy, org.y); }
    if (zmask & 4) { minz = -inf; maxz = inf; inside = inside && zero_axis_inside(cz, hz, org.z); }
    float minT = fmaxf(fmaxf(minx, miny), minz);
    float maxT = fminf(fminf(maxx, maxy), maxz);
    SlabRange r;
    r.enter = fmaxf(minT, 0.0f);
    r.exit = inside ? fminf(maxT, closestT) : -inf; // outside on a zero axis: empty interval
    return r;
}

// Resumable traversal: begin() once per ray, then step_internal()/step_leaf() on the current
// node until done(). The node to visit next is held in a register (`cur`), the rest of the
// stack in a per-thread local array passed in by the caller, so only the far children of
// two-hit nodes ever touch local memory. Splitting the step lets the persistent kernel
// schedule a whole warp onto ONE of the two code paths per iteration (k_extend) while the
// inline queries of the shading stage simply run to completion. Per-ray visit order and
// counters are identical either way and identical to the reference's stack discipline.
#define TB_NO_NODE 0x7fffffffu
struct Traversal {
    tbm::f3 org, inv, oinv, shear;
    int kx, ky, kz;
    float tmin, tmax, committedT, hb1, hb2;
    uint32_t hitGeom, hitPrim, trisTested, pairsTested; // BoxesTested = 2 x pairsTested (:751 counts both children)
    bool haveHit;
    int zmask;     // bit a set: direction component a is exactly zero (deviation D6)
    uint32_t cur;  // node reference to process next (TB_NO_NODE = traversal finished)
    int sp;        // index of the top entry of the memory stack (0: only the sentinel is left)

    __device__ __forceinline__ uint32_t boxes_tested() const { return 2u * pairsTested; }
    __device__ __forceinline__ uint32_t steps() const { return trisTested + pairsTested; } // node visits so far
    __device__ __forceinline__ void idle(uint32_t* stack) { stack[0] = TB_NO_NODE; sp = 0; cur = TB_NO_NODE; }

    __device__ __forceinline__ void begin(const DeviceBvh& bvh, uint32_t* stack, tbm::f3 o, tbm::f3 dir, float tmin_, float tmax_) {
        using namespace tbm;
        org = o;
        inv = mk3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z); // GetRayData, TraverseFunction.hlsli:473-495
        // Deviation D6: exactly-zero direction components (the reference's rand() returns
        // exactly 0 about once in 300 draws). Literally rcp(0) = inf makes that axis' slab test NaN, which
        // min/max drop, so the ray would walk every node overlapping the other two slabs. A zero axis is
        // tested by containment instead (slab_zero); same hits, orders of magnitude fewer visits.
        zmask = (dir.x == 0.0f ? 1 : 0) | (dir.y == 0.0f ? 2 : 0) | (dir.z == 0.0f ? 4 : 0);
        oinv = org * inv;
        f3 ad = abs3(dir);
        kz = (ad.x > ad.y && ad.x > ad.z) ? 0 : (ad.y > ad.z ? 1 : 2);
        kx = kz == 2 ? 0 : kz + 1; ky = kx == 2 ? 0 : kx + 1;
        if (comp(dir, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
        float dz = comp(dir, kz);
        shear = mk3(comp(dir, kx) / dz, comp(dir, ky) / dz, 1.0f / dz);
        tmin = tmin_; tmax = tmax_; committedT = tmax_;
        haveHit = false; hitGeom = hitPrim = 0xffffffffu; hb1 = hb2 = 0.0f;
        trisTested = pairsTested = 0;
        stack[0] = TB_NO_NODE;
        sp = 0;
        cur = TB_NO_NODE;
        const SlabRange rr = zmask ? slab_zero(committedT, org, zmask, oinv, inv, abs3(inv), bvh.root.c[0], bvh.root.c[1], bvh.root.c[2], bvh.root.h[0], bvh.root.h[1], bvh.root.h[2])
                                   : slab(committedT, oinv, inv, abs3(inv), bvh.root.c[0], bvh.root.c[1], bvh.root.c[2], bvh.root.h[0], bvh.root.h[1], bvh.root.h[2]);
        const bool rootHit = rr.enter < rr.exit;
        // Deviation D7: a ray with a NaN origin or direction component can never commit a hit (t0 is
        // NaN), but its NaN slabs are dropped by min/max, so literally it walks every node overlapping the other
        // axes (the whole tree for an all-NaN direction). It is reported as the miss it is, with zero tests.
        const bool nanRay = o.x != o.x || o.y != o.y || o.z != o.z || dir.x != dir.x || dir.y != dir.y || dir.z != dir.z;
        if (rootHit && !nanRay)
            cur = (bvh.root.flags & 0x80000000u) ? (0x80000000u | (bvh.root.flags & 0x3fffffffu)) : 0u;
    }
    // The bottom level of a two-level query (TraverseFunction.hlsli:621): the walk starts at the instance's root
    // WITHOUT a root box test (the instance's world box stood for it), with the t committed so far and the id triple
    // `seedGeom / seedPrim` that an equal-t hit has to beat (0xffffffff: any hit wins the tie, 0 / 0: none does).
    __device__ __forceinline__ void begin_bottom_level(const DeviceBvh& unusedRootBox, uint32_t rootRef, uint32_t* stack, tbm::f3 o, tbm::f3 dir,
                                                       float tmin_, float tmax_, float committedSoFar, bool haveHitSoFar, uint32_t seedGeom, uint32_t seedPrim) {
        begin(unusedRootBox, stack, o, dir, tmin_, tmax_);
        const bool nanRay = o.x != o.x || o.y != o.y || o.z != o.z || dir.x != dir.x || dir.y != dir.y || dir.z != dir.z;
        cur = nanRay ? TB_NO_NODE : rootRef;
        committedT = committedSoFar; haveHit = haveHitSoFar; hitGeom = seedGeom; hitPrim = seedPrim;
    }
    __device__ __forceinline__ bool done() const { return cur == TB_NO_NODE; }
    __device__ __forceinline__ bool at_leaf() const { return (cur & 0x80000000u) != 0; }
    __device__ __forceinline__ void pop(const uint32_t* stack) { cur = stack[sp]; --sp; } // the sentinel ends the traversal

    // cur is a leaf: RayTriangleIntersect, TraverseFunction.hlsli:231-313 (two-sided, `precise` => unfused)
    __device__ __forceinline__ void step_leaf(const uint32_t* stack, const float4* __restrict__ tris) {
        using namespace tbm;
        uint32_t slot = cur & 0x3fffffffu;
        float4 q0 = __ldg(tris + 3 * (size_t)slot), q1 = __ldg(tris + 3 * (size_t)slot + 1), q2 = __ldg(tris + 3 * (size_t)slot + 2);
        trisTested++;
        f3 v0 = mk3(q0.x, q0.y, q0.z) - org, v1 = mk3(q1.x, q1.y, q1.z) - org, v2 = mk3(q2.x, q2.y, q2.z) - org;
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
        bool ok = !((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) && det != 0.0f;
        if (ok) {
            Az = shear.z * Az; Bz = shear.z * Bz; Cz = shear.z * Cz;
            float T = (U * Az + V * Bz) + W * Cz;
            float sct = fabsf(T);
            if ((T > 0.0f) != (det > 0.0f)) sct = -sct;
            if (!(sct < 0.0f || sct > committedT * fabsf(det))) {
                float rcpDet = 1.0f / det;
                float t0 = T * rcpDet;
                uint32_t g = __float_as_uint(q0.w), p = __float_as_uint(q1.w);
                bool closer = t0 < committedT; // TestLeafNodeIntersections :420 + equal-t tie-break
                bool tie = haveHit && t0 == committedT && (g < hitGeom || (g == hitGeom && p < hitPrim));
                if ((closer || tie) && t0 > tmin) {
                    committedT = t0; hb1 = V * rcpDet; hb2 = W * rcpDet; hitGeom = g; hitPrim = p; haveHit = true;
                }
            }
        }
        pop(stack);
    }

    // cur is an internal node: both child boxes, push far then near; left is near on equal t (:754-765)
    __device__ __forceinline__ void step_internal(uint32_t* stack, const float4* __restrict__ pairs) {
        using namespace tbm;
        uint32_t ref = cur;
        // one 64-byte node = two 256-bit loads (LDG.E.256, sm_100): the kernel is bound by L1 wavefronts
        // (one per lane per load instruction for scattered nodes), so halving the load count matters
        float4 a, b, c, d;
        ldg256(pairs + 4 * (size_t)ref, a, b);
        ldg256(pairs + 4 * (size_t)ref + 2, c, d);
        f3 ainv = abs3(inv);
        SlabRange L, R;
        if (zmask) { // rare path, kept out of line
            L = slab_zero(committedT, org, zmask, oinv, inv, ainv, a.x, a.y, a.z, b.x, b.y, b.z);
            R = slab_zero(committedT, org, zmask, oinv, inv, ainv, c.x, c.y, c.z, d.x, d.y, d.z);
        } else {
            L = slab(committedT, oinv, inv, ainv, a.x, a.y, a.z, b.x, b.y, b.z);
            R = slab(committedT, oinv, inv, ainv, c.x, c.y, c.z, d.x, d.y, d.z);
        }
        const bool lh = L.enter < L.exit, rh = R.enter < R.exit;
        const float lt = L.enter, rt = R.enter;
        pairsTested++;
        const uint32_t lref = __float_as_uint(a.w), rref = __float_as_uint(b.w);
        // Branch-free child selection (the three outcomes used to serialise as divergent branches):
        // the near child (or the only one hit) is visited next, the far child of a two-hit node waits in
        // memory (predicated store), a node with no hit pops (predicated load).
        const bool both = lh && rh;
        const bool takeRight = rh && (!lh || rt < lt); // both hit: right only when strictly nearer (:754-765)
        const uint32_t nearRef = takeRight ? rref : lref;
        const uint32_t farRef = takeRight ? lref : rref;
        if (both) { ++sp; stack[sp] = farRef; } // never overflows: sp <= tree depth <= TB_STACK_DEPTH (checked by the host after the build)
        cur = nearRef;
        if (!(lh || rh)) pop(stack);
    }
    __device__ __forceinline__ void step(uint32_t* stack, const float4* __restrict__ pairs, const float4* __restrict__ tris) {
        if (at_leaf()) step_leaf(stack, tris); else step_internal(stack, pairs);
    }

    // ---- suspension: a long ray can be parked in global memory and resumed by a later kernel
    // in a warp of similarly long rays. The state is exactly the registers below plus the live
    // part of the stack, so pausing never changes the visit order or the counters.
    static constexpr int kRecordWords = 112; // 12-word header + TB_STACK_DEPTH entries, 16 B aligned
    __device__ __forceinline__ void suspend(uint32_t* __restrict__ rec, const uint32_t* stack, uint32_t pi) const {
        uint4* r4 = (uint4*)rec;
        r4[0] = make_uint4(pi, (uint32_t)sp, __float_as_uint(committedT), __float_as_uint(hb1));
        r4[1] = make_uint4(__float_as_uint(hb2), hitGeom, hitPrim, haveHit ? 1u : 0u);
        r4[2] = make_uint4(trisTested, pairsTested, cur, 0u);
        for (int i = 0; i < sp; i++) rec[12 + i] = stack[1 + i];
    }
    // ray setup is recomputed from the ray (same arithmetic => same values), the rest is restored
    __device__ __forceinline__ uint32_t resume(const DeviceBvh& bvh, const uint32_t* __restrict__ rec, uint32_t* stack,
                                               const float4* __restrict__ rayO, const float4* __restrict__ rayD, float tmin_, float tmax_) {
        const uint4* r4 = (const uint4*)rec;
        uint4 a = r4[0], b = r4[1], c = r4[2];
        uint32_t pi = a.x;
        float4 o = rayO[pi], d = rayD[pi];
        begin(bvh, stack, tbm::mk3(o.x, o.y, o.z), tbm::mk3(d.x, d.y, d.z), tmin_, tmax_);
        sp = (int)a.y; committedT = __uint_as_float(a.z); hb1 = __uint_as_float(a.w);
        hb2 = __uint_as_float(b.x); hitGeom = b.y; hitPrim = b.z; haveHit = b.w != 0;
        trisTested = c.x; pairsTested = c.y; cur = c.z;
        for (int i = 0; i < sp; i++) stack[1 + i] = rec[12 + i];
        return pi;
    }

    __device__ __forceinline__ void result(HitRec& out) const {
        out.tris = trisTested;
        out.boxes = boxes_tested();
        if (haveHit && committedT < tmax) { out.t = committedT; out.b1 = hb1; out.b2 = hb2; out.prim = hitPrim; out.geom = hitGeom; }
        else { out.t = -1.0f; out.b1 = out.b2 = 0.0f; out.prim = out.geom = 0xffffffffu; }
    }
};

__device__ __forceinline__ void trace_ray(const DeviceBvh& bvh, tbm::f3 org, tbm::f3 dir, float tmin, float tmax, HitRec& out) {
    Traversal tr;
    uint32_t stack[TB_STACK_WORDS];
    tr.begin(bvh, stack, org, dir, tmin, tmax);
    const float4* __restrict__ pairs = (const float4*)bvh.pairs;
    const float4* __restrict__ tris = (const float4*)bvh.tris;
    while (!tr.done()) tr.step(stack, pairs, tris);
    tr.result(out);
}

} // namespace tbd
