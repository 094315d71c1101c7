// bvh.cpp — CPU ORACLE (test infrastructure): sequential restatement of the reference's
// GPU LBVH builder (D3D12RaytracingFallback/src/GpuBVH2Builder.cpp:167-356 and the HLSL
// kernels it dispatches). Output is the reference's BVH byte layout.
//
// Deliberate, documented deviations (both also made by the CUDA builder):
//  D1  child order on equal subtree sizes: "swap iff leftCount > rightCount". The
//      reference's rule (ComputeAABBs.hlsli:145-158) depends on thread arrival order.
//  D2  treelet climbing is not capped. The reference stops every thread group after 33
//      levels (TreeletReorder.hlsl:293-311), which of two merging groups continues is
//      arrival-order dependent. Scene::maxTreeletClimb reports the longest chain so a
//      test can tell when the cap would have mattered.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <numeric>
#include "../tracerboy_b200/csrc/common/tb_vec.h"
#include "oracle.h"

using namespace tbm;

namespace oracle {

namespace {

struct Prim { uint32_t type; float v[9]; };                 // RayTracingHlslCompat.h:122-176 (40 B)
struct Meta { uint32_t geom, prim, flags; };                // :182-188 (12 B)
struct HNode { uint32_t parent, left, right; };             // :33-38
struct Box { f3 mn, mx; };                                  // AABB :40-58
struct AABBNode { float c[3]; uint32_t flags; float h[3]; uint32_t right; }; // :344-385 (32 B)
static_assert(sizeof(Prim) == 40 && sizeof(Meta) == 12 && sizeof(AABBNode) == 32, "layout");

const uint32_t kLeafFlag = 0x80000000u; // RayTracingHelper.hlsli:24

inline f3 pv(const Prim& p, int k) { return mk3(p.v[3 * k], p.v[3 * k + 1], p.v[3 * k + 2]); }

// CalculateMortonCodesBindings.h:117-162
uint32_t morton_code(f3 centroid, f3 smin, f3 smax) {
    const float epsilon = 0.00001f;
    f3 dim = max3(smax - smin, mk3(epsilon));
    f3 unit = (centroid - smin) / dim;
    const float maxCoord = 1024.0f;
    f3 adj = min3(max3(unit * maxCoord, mk3(0.0f)), mk3(maxCoord - 1.0f));
    uint32_t coords[3] = {(uint32_t)adj.y, (uint32_t)adj.x, (uint32_t)adj.z};
    uint32_t code = 0;
    for (uint32_t bit = 0; bit < 10; bit++)
        for (uint32_t axis = 0; axis < 3; axis++)
            if ((1u << bit) & coords[axis]) code |= 1u << (bit * 3 + axis);
    return code;
}

// BuildBVHSplits.hlsli:18-141
struct Karras {
    const uint32_t* codes;
    uint32_t n;
    static int clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
    int lcp(uint32_t a, uint32_t b) const {
        if (a >= n || b >= n) return -1;
        uint32_t ca = codes[a], cb = codes[b];
        if (ca != cb) return clz(ca ^ cb);
        return clz(a ^ b) + 31;
    }
    void range(uint32_t idx, uint32_t& first, uint32_t& last) const {
        int d = lcp(idx, idx + 1) - lcp(idx, idx - 1);
        d = d < -1 ? -1 : (d > 1 ? 1 : d);
        int minPrefix = lcp(idx, idx - d);
        int maxLength = 2;
        while (lcp(idx, idx + (uint32_t)(maxLength * d)) > minPrefix) maxLength *= 4;
        int length = 0;
        for (int t = maxLength / 2; t > 0; t /= 2)
            if (lcp(idx, idx + (uint32_t)((length + t) * d)) > minPrefix) length += t;
        uint32_t j = idx + (uint32_t)(length * d);
        first = std::min(idx, j);
        last = std::max(idx, j);
    }
    uint32_t split(uint32_t first, uint32_t last) const {
        int common = lcp(first, last);
        int sp = (int)first;
        int step = (int)(last - first);
        do {
            step = (step + 1) >> 1;
            int ns = sp + step;
            if ((uint32_t)ns < last) {
                if (lcp(first, (uint32_t)ns) > common) sp = ns;
            }
        } while (step > 1);
        return (uint32_t)sp;
    }
};

inline float surface_area(const Box& b) { // TreeletReorderBindings.h:101-105
    f3 d = b.mx - b.mn;
    return 2.0f * ((d.x * d.y + d.x * d.z) + d.y * d.z);
}
inline Box combine(const Box& a, const Box& b) { return {min3(a.mn, b.mn), max3(a.mx, b.mx)}; }

// RayTracingHelper.hlsli:251-263 (GetBoxDataFromTriangle) -> center/halfDim
inline void leaf_box(const Prim& p, f3& center, f3& half) {
    f3 v0 = pv(p, 0), v1 = pv(p, 1), v2 = pv(p, 2);
    f3 mn = min3(min3(v0, v1), v2);
    f3 mx = max3(max3(v0, v1), v2);
    mn = min3(mn, mx - 0.001f);
    center = (mn + mx) * 0.5f;
    half = mx - center;
}

// GetBoxFromChildBoxes + AABBtoBoundingBox, RayTracingHelper.hlsli:229-235, 275-285: the parent box is fitted around the
// children's centre +- half-extent corners (not around their original min / max), then stored as centre / half-extent
inline void parent_box(f3 ac, f3 ah, f3 bc, f3 bh, f3& c, f3& h) {
    f3 mn = min3(ac - ah, bc - bh);
    f3 mx = max3(ac + ah, bc + bh);
    c = (mn + mx) * 0.5f;
    h = mx - c;
}

// One treelet: FormTreelet, FindOptimalPartitions, ReformTree (TreeletReorder.hlsl:38-236) for the treelet rooted at `root`.
void optimize_treelet(std::vector<HNode>& H, std::vector<Box>& aabb, uint32_t nInternal, uint32_t root) {
    auto isLeaf = [&](uint32_t i) { return i >= nInternal; };
    // FormTreelet (TreeletReorder.hlsl:38-80)
    uint32_t leaves[7], internals[6];
    internals[0] = root;
    leaves[0] = H[root].left;
    leaves[1] = H[root].right;
    for (uint32_t size = 2; size < 7; size++) {
        float largest = 0.0f;
        uint32_t pick = 0, pickIdx = 0;
        for (uint32_t i = 0; i < size; i++) {
            uint32_t t = leaves[i];
            if (!isLeaf(t)) {
                float sa = surface_area(aabb[t]);
                if (sa > largest) { largest = sa; pick = t; pickIdx = i; }
            }
        }
        HNode nt = H[pick];
        internals[size - 1] = pick;
        leaves[pickIdx] = nt.left;
        leaves[size] = nt.right;
    }
    // FindOptimalPartitions (:82-171)
    float cost[128];
    uint32_t part[128];
    memset(part, 0, sizeof(part));
    cost[0] = 0.0f;
    for (uint32_t mask = 1; mask < 128; mask++) {
        Box b = {mk3(FLT_MAX), mk3(-FLT_MAX)};
        for (uint32_t i = 0; i < 7; i++)
            if ((1u << i) & mask) b = combine(b, aabb[leaves[i]]);
        cost[mask] = surface_area(b);
    }
    float rootSA = surface_area(aabb[root]);
    for (uint32_t i = 0; i < 7; i++) cost[1u << i] = 1.0f * surface_area(aabb[leaves[i]]) / rootSA;
    for (uint32_t subset = 2; subset <= 7; subset++) {
        for (uint32_t mask = 1; mask < 128; mask++) {
            if ((uint32_t)__builtin_popcount(mask) != subset) continue;
            float lowest = FLT_MAX;
            uint32_t best = 0;
            uint32_t delta = (mask - 1) & mask;
            uint32_t p = (0u - delta) & mask;
            do {
                float c = cost[p] + cost[mask ^ p];
                if (c < lowest) { lowest = c; best = p; }
                p = (p - delta) & mask;
            } while (p != 0);
            cost[mask] = 1.0f * cost[mask] + lowest;
            part[mask] = best;
        }
    }
    // ReformTree (:173-236)
    struct Entry { uint32_t mask, node; };
    Entry stack[7];
    uint32_t allocated = 1, sp = 1;
    stack[0] = {127u, internals[0]};
    while (sp > 0) {
        Entry e = stack[--sp];
        Entry l, r;
        l.mask = part[e.mask];
        if (__builtin_popcount(l.mask) > 1) { l.node = internals[allocated++]; stack[sp++] = l; }
        else l.node = leaves[__builtin_ctz(l.mask)];
        r.mask = e.mask ^ l.mask;
        if (__builtin_popcount(r.mask) > 1) { r.node = internals[allocated++]; stack[sp++] = r; }
        else r.node = leaves[__builtin_ctz(r.mask)];
        H[e.node].left = l.node;
        H[e.node].right = r.node;
        H[l.node].parent = e.node;
        H[r.node].parent = e.node;
    }
    for (int j = 5; j >= 0; j--) {
        uint32_t in = internals[j];
        aabb[in] = combine(aabb[H[in].left], aabb[H[in].right]);
    }
}

// One treelet-reorder pass (ClearBuffers.hlsl, FindTreelets.hlsl, TreeletReorder.hlsl).
void treelet_pass(std::vector<HNode>& H, const std::vector<Prim>& prims, uint32_t n, uint32_t minTris,
                  uint32_t& maxClimb) {
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    std::vector<Box> aabb(total);
    std::vector<uint32_t> count(total, 1), depth(total, 0);
    // post-order over the current topology (children before parents)
    std::vector<uint32_t> order;
    order.reserve(nInternal);
    {
        std::vector<std::pair<uint32_t, int>> st;
        st.push_back({0, 0});
        while (!st.empty()) {
            auto& top = st.back();
            uint32_t node = top.first;
            if (node >= nInternal) { st.pop_back(); continue; }
            if (top.second == 0) { top.second = 1; depth[H[node].left] = depth[node] + 1; st.push_back({H[node].left, 0}); }
            else if (top.second == 1) { top.second = 2; depth[H[node].right] = depth[node] + 1; st.push_back({H[node].right, 0}); }
            else { order.push_back(node); st.pop_back(); }
        }
    }
    for (uint32_t i = 0; i < n; i++) { // FindTreelets.hlsl:16-29 ComputeLeafAABB
        f3 c, h;
        leaf_box(prims[i], c, h);
        aabb[nInternal + i] = {c - h, c + h};
    }
    for (uint32_t node : order) {
        count[node] = count[H[node].left] + count[H[node].right];
        aabb[node] = combine(aabb[H[node].left], aabb[H[node].right]);
    }
    for (uint32_t root : order) {
        if (count[root] < minTris) continue;
        bool base = count[H[root].left] < minTris && count[H[root].right] < minTris;
        if (base) maxClimb = std::max(maxClimb, depth[root] + 1);
        optimize_treelet(H, aabb, nInternal, root);
    }
}

} // namespace

// test hook: the Karras hierarchy over sorted codes (parent, left, right per node; 2n-1 nodes), for the comparison with
// the reference's own BuildBVHSplits.hlsli text (oracle/ref/ref_karras.cpp)
void karras_public(const uint32_t* codes, uint32_t n, uint32_t* out3) {
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    std::vector<HNode> H(total, HNode{0xffffffffu, 0, 0});
    Karras K{codes, n};
    for (uint32_t idx = 0; idx < nInternal; idx++) {
        uint32_t first, last;
        K.range(idx, first, last);
        uint32_t split = K.split(first, last);
        uint32_t a = (split == first) ? nInternal + split : split;
        uint32_t b = (split + 1 == last) ? nInternal + split + 1 : split + 1;
        H[idx].left = a; H[idx].right = b; H[a].parent = idx; H[b].parent = idx;
    }
    memcpy(out3, H.data(), sizeof(HNode) * total);
}

// test hook: one treelet optimisation on a caller-provided hierarchy (3 words per node) and boxes (min, max: 6 floats per
// node), for the comparison with the reference's own TreeletReorder.hlsl text (oracle/ref/ref_treelet.cpp)
void treelet_public(uint32_t* H3, float* aabb6, uint32_t n, uint32_t root) {
    const uint32_t total = 2 * n - 1;
    std::vector<HNode> H(total);
    std::vector<Box> boxes(total);
    memcpy(H.data(), H3, sizeof(HNode) * total);
    for (uint32_t i = 0; i < total; i++) boxes[i] = {mk3(aabb6[6 * i], aabb6[6 * i + 1], aabb6[6 * i + 2]), mk3(aabb6[6 * i + 3], aabb6[6 * i + 4], aabb6[6 * i + 5])};
    optimize_treelet(H, boxes, n - 1, root);
    memcpy(H3, H.data(), sizeof(HNode) * total);
    for (uint32_t i = 0; i < total; i++) {
        aabb6[6 * i] = boxes[i].mn.x; aabb6[6 * i + 1] = boxes[i].mn.y; aabb6[6 * i + 2] = boxes[i].mn.z;
        aabb6[6 * i + 3] = boxes[i].mx.x; aabb6[6 * i + 4] = boxes[i].mx.y; aabb6[6 * i + 5] = boxes[i].mx.z;
    }
}

// test hook: one whole treelet pass (ClearBuffers + FindTreelets + TreeletReorder) on a caller-provided hierarchy and the
// sorted primitives (40 bytes each)
void treelet_pass_public(uint32_t* H3, const void* prims40, uint32_t n, uint32_t minTris, uint32_t* maxClimb) {
    const uint32_t total = 2 * n - 1;
    std::vector<HNode> H(total);
    memcpy(H.data(), H3, sizeof(HNode) * total);
    std::vector<Prim> prims((const Prim*)prims40, (const Prim*)prims40 + n);
    uint32_t climb = 0;
    treelet_pass(H, prims, n, minTris, climb);
    memcpy(H3, H.data(), sizeof(HNode) * total);
    if (maxClimb) *maxClimb = climb;
}

// test hooks: the two box constructors of the node writer (leaf from a triangle, parent from two children)
void leaf_box_public(const float* v9, float* c3, float* h3) {
    Prim p; p.type = 1; memcpy(p.v, v9, 36);
    f3 c, h;
    leaf_box(p, c, h);
    c3[0] = c.x; c3[1] = c.y; c3[2] = c.z; h3[0] = h.x; h3[1] = h.y; h3[2] = h.z;
}
void parent_box_public(const float* ac, const float* ah, const float* bc, const float* bh, float* c3, float* h3) {
    f3 c, h;
    parent_box(mk3(ac[0], ac[1], ac[2]), mk3(ah[0], ah[1], ah[2]), mk3(bc[0], bc[1], bc[2]), mk3(bh[0], bh[1], bh[2]), c, h);
    c3[0] = c.x; c3[1] = c.y; c3[2] = c.z; h3[0] = h.x; h3[1] = h.y; h3[2] = h.z;
}

uint32_t morton_public(const float* c, const float* mn, const float* mx) {
    return morton_code(mk3(c[0], c[1], c[2]), mk3(mn[0], mn[1], mn[2]), mk3(mx[0], mx[1], mx[2]));
}

// LoadPrimitives (LoadPrimitivesPass.cpp:56-169, BottomLevelLoadTriangles.hlsli:88-126): triangle soup in geometry order
static void load_primitives(const Scene& s, std::vector<Prim>& prims, std::vector<Meta>& meta) {
    for (size_t g = 0; g < s.geoms.size(); g++) {
        const TbGeometryRecord& G = s.geoms[g];
        for (uint32_t t = 0; t < G.IndexCount / 3; t++) {
            Prim p;
            p.type = 1;
            for (int k = 0; k < 3; k++) {
                const TbFloat3& v = s.positions[G.VertexFirst + s.indices[G.IndexFirst + 3 * t + k]];
                p.v[3 * k] = v.x; p.v[3 * k + 1] = v.y; p.v[3 * k + 2] = v.z;
            }
            prims.push_back(p);
            meta.push_back({(uint32_t)g, t, G.GeometryFlags});
        }
    }
}
void load_primitives_public(const Scene& s, void* prims40, void* meta12) { // test hook
    std::vector<Prim> prims; std::vector<Meta> meta;
    load_primitives(s, prims, meta);
    memcpy(prims40, prims.data(), sizeof(Prim) * prims.size());
    memcpy(meta12, meta.data(), sizeof(Meta) * meta.size());
}
// CalculateSceneAABBFromPrimitives.hlsl:16-40 (the reduction tree of SceneAABBCalculator.cpp:36-84 is min / max: order free)
static void scene_box(const Prim* prims, uint32_t n, f3& smin, f3& smax) {
    smin = mk3(FLT_MAX); smax = mk3(-FLT_MAX);
    for (uint32_t i = 0; i < n; i++) {
        const Prim& p = prims[i];
        smin = min3(min3(min3(pv(p, 0), smin), pv(p, 1)), pv(p, 2));
        smax = max3(max3(max3(pv(p, 0), smax), pv(p, 1)), pv(p, 2));
    }
}
// GetCentroid, CalculateMortonCodesForPrimitives.hlsl:17-24
static f3 centroid(const Prim& p) { return ((pv(p, 0) + pv(p, 1)) + pv(p, 2)) / 3.0f; }
// the order the sort produces: ascending by (code, index) (BitonicSortCommon.hlsli:37-47 with NullItem = 0xffffffff)
static bool sorts_before(uint32_t codeA, uint32_t indexA, uint32_t codeB, uint32_t indexB) {
    return codeA != codeB ? codeA < codeB : indexA < indexB;
}
// test hooks
void scene_box_public(const void* prims40, uint32_t n, float* out6) {
    f3 a, b;
    scene_box((const Prim*)prims40, n, a, b);
    out6[0] = a.x; out6[1] = a.y; out6[2] = a.z; out6[3] = b.x; out6[4] = b.y; out6[5] = b.z;
}
void centroid_public(const void* prim40, float* out3) { f3 c = centroid(*(const Prim*)prim40); out3[0] = c.x; out3[1] = c.y; out3[2] = c.z; }
int sorts_before_public(uint32_t codeA, uint32_t indexA, uint32_t codeB, uint32_t indexB) { return sorts_before(codeA, indexA, codeB, indexB) ? 1 : 0; }

bool build_bvh(Scene& s, int treeletPasses, std::string& err) {
    std::vector<Prim> prims;
    std::vector<Meta> meta;
    load_primitives(s, prims, meta);
    const uint32_t n = (uint32_t)prims.size();
    if (n == 0) { err = "scene has no triangles"; return false; }
    // The reference keeps 24-bit child / leaf indices (RayTracingHelper.hlsli:97-103), which
    // silently wrap above 2^24 nodes. We mask with 30 bits instead (bit 31 = leaf, bit 30 =
    // procedural): byte-identical for every scene the reference can represent, and valid
    // up to 2^30 nodes (needed by the 20 M-triangle config).
    if (2ull * n - 1 > (1ull << 30)) { err = "too many triangles for 30-bit node indices"; return false; }
    // CalculateSceneAABBFromPrimitives.hlsl:16-40
    f3 smin, smax;
    scene_box(prims.data(), n, smin, smax);
    // Morton codes (CalculateMortonCodesForPrimitives.hlsl:17-24)
    std::vector<uint32_t> codes(n), order(n);
    for (uint32_t i = 0; i < n; i++) {
        codes[i] = morton_code(centroid(prims[i]), smin, smax);
        order[i] = i;
    }
    // Bitonic sort == stable sort by (code, index) (BitonicSortCommon.hlsli:37-47)
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return sorts_before(codes[a], a, codes[b], b); });
    // RearrangeTriangles.hlsl:29-36
    std::vector<Prim> sp(n);
    std::vector<Meta> sm(n);
    std::vector<uint32_t> sc(n);
    for (uint32_t i = 0; i < n; i++) { sp[i] = prims[order[i]]; sm[i] = meta[order[i]]; sc[i] = codes[order[i]]; }

    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    std::vector<HNode> H(total, HNode{0xffffffffu, 0, 0});
    if (n > 1) {
        Karras K{sc.data(), n};
        for (uint32_t idx = 0; idx < nInternal; idx++) {
            uint32_t first, last;
            K.range(idx, first, last);
            uint32_t split = K.split(first, last);
            uint32_t a = (split == first) ? nInternal + split : split;
            uint32_t b = (split + 1 == last) ? nInternal + split + 1 : split + 1;
            H[idx].left = a;
            H[idx].right = b;
            H[a].parent = idx;
            H[b].parent = idx;
        }
        // TreeletReorder.cpp:63-108
        uint32_t minTris = 7;
        s.maxTreeletClimb = 0;
        for (int pass = 0; pass < treeletPasses; pass++) {
            if (minTris > n) break;
            treelet_pass(H, sp, n, minTris, s.maxTreeletClimb);
            minTris *= 2;
        }
    }
    s.hierarchy.assign((const uint32_t*)H.data(), (const uint32_t*)H.data() + 3 * (size_t)total);
    // PrepareForComputeAABBs + ComputeAABBs (ComputeAABBs.hlsli:69-172)
    const uint32_t offBoxes = 16;
    const uint32_t offPrims = offBoxes + 32 * total;
    const uint32_t offMeta = offPrims + 40 * n;
    const uint64_t totalSize = (uint64_t)offMeta + 12ull * n;
    s.bvh.assign(totalSize, 0);
    uint32_t header[4] = {offBoxes, offPrims, offMeta, (uint32_t)totalSize};
    memcpy(s.bvh.data(), header, 16);
    AABBNode* nodes = (AABBNode*)(s.bvh.data() + offBoxes);
    memcpy(s.bvh.data() + offPrims, sp.data(), 40ull * n);
    memcpy(s.bvh.data() + offMeta, sm.data(), 12ull * n);
    for (uint32_t i = 0; i < n; i++) {
        f3 c, h;
        leaf_box(sp[i], c, h);
        AABBNode& nd = nodes[nInternal + i];
        nd.c[0] = c.x; nd.c[1] = c.y; nd.c[2] = c.z;
        nd.h[0] = h.x; nd.h[1] = h.y; nd.h[2] = h.z;
        nd.flags = i | kLeafFlag;
        nd.right = 1;
    }
    if (n > 1) {
        // bottom-up in post-order; counts decide the child swap (deviation D1)
        std::vector<uint32_t> cnt(total, 1);
        std::vector<std::pair<uint32_t, int>> st;
        st.push_back({0, 0});
        while (!st.empty()) {
            auto& top = st.back();
            uint32_t node = top.first;
            if (node >= nInternal) { st.pop_back(); continue; }
            if (top.second == 0) { top.second = 1; st.push_back({H[node].left, 0}); }
            else if (top.second == 1) { top.second = 2; st.push_back({H[node].right, 0}); }
            else {
                uint32_t l = H[node].left, r = H[node].right;
                if (cnt[l] > cnt[r]) std::swap(l, r);
                cnt[node] = cnt[l] + cnt[r];
                const AABBNode& A = nodes[l];
                const AABBNode& B = nodes[r];
                f3 ac = mk3(A.c[0], A.c[1], A.c[2]), ah = mk3(A.h[0], A.h[1], A.h[2]);
                f3 bc = mk3(B.c[0], B.c[1], B.c[2]), bh = mk3(B.h[0], B.h[1], B.h[2]);
                f3 c, h;
                parent_box(ac, ah, bc, bh, c, h);
                AABBNode& nd = nodes[node];
                nd.c[0] = c.x; nd.c[1] = c.y; nd.c[2] = c.z;
                nd.h[0] = h.x; nd.h[1] = h.y; nd.h[2] = h.z;
                nd.flags = l & 0x3fffffffu;
                nd.right = r;
                st.pop_back();
            }
        }
    }
    s.numPrims = n;
    return true;
}

// BuildRaytracingAccelerationStructure with PERFORM_UPDATE (GpuBVH2Builder.cpp:165-234): the hierarchy is kept, the
// primitives are reloaded straight into their sorted slots and ComputeAABBs refits every box bottom-up reading the
// child indices from the node flags (ComputeAABBs.hlsli:39-67: GetLeftNodeIndex / GetRightNodeIndex of the stored node).
// The reference finds a primitive's slot through the sort cache it wrote at build time; the slot's own metadata
// (geometry, primitive index) names the same triangle, which is what is used here. `s.positions` holds the moved vertices.
bool update_bvh(Scene& s, std::string& err) {
    const uint32_t n = s.numPrims;
    if (n == 0 || s.bvh.empty()) { err = "no acceleration structure to update"; return false; }
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    const uint32_t* header = (const uint32_t*)s.bvh.data();
    AABBNode* nodes = (AABBNode*)(s.bvh.data() + header[0]);
    Prim* sp = (Prim*)(s.bvh.data() + header[1]);
    const Meta* sm = (const Meta*)(s.bvh.data() + header[2]);
    for (uint32_t i = 0; i < n; i++) {
        const TbGeometryRecord& G = s.geoms[sm[i].geom];
        Prim p;
        p.type = 1;
        for (int k = 0; k < 3; k++) {
            const TbFloat3& v = s.positions[G.VertexFirst + s.indices[G.IndexFirst + 3 * sm[i].prim + k]];
            p.v[3 * k] = v.x; p.v[3 * k + 1] = v.y; p.v[3 * k + 2] = v.z;
        }
        sp[i] = p;
        f3 c, h;
        leaf_box(p, c, h);
        AABBNode& nd = nodes[nInternal + i];
        nd.c[0] = c.x; nd.c[1] = c.y; nd.c[2] = c.z;
        nd.h[0] = h.x; nd.h[1] = h.y; nd.h[2] = h.z;
    }
    if (n > 1) { // post-order over the stored topology; flags and child order stay as built
        std::vector<std::pair<uint32_t, int>> st;
        st.push_back({0, 0});
        while (!st.empty()) {
            auto& top = st.back();
            const uint32_t node = top.first;
            if (node >= nInternal) { st.pop_back(); continue; }
            const uint32_t l = nodes[node].flags & 0x3fffffffu, r = nodes[node].right;
            if (top.second == 0) { top.second = 1; st.push_back({l, 0}); }
            else if (top.second == 1) { top.second = 2; st.push_back({r, 0}); }
            else {
                const AABBNode& A = nodes[l];
                const AABBNode& B = nodes[r];
                f3 c, h;
                parent_box(mk3(A.c[0], A.c[1], A.c[2]), mk3(A.h[0], A.h[1], A.h[2]), mk3(B.c[0], B.c[1], B.c[2]), mk3(B.h[0], B.h[1], B.h[2]), c, h);
                AABBNode& nd = nodes[node];
                nd.c[0] = c.x; nd.c[1] = c.y; nd.c[2] = c.z;
                nd.h[0] = h.x; nd.h[1] = h.y; nd.h[2] = h.z;
                st.pop_back();
            }
        }
    }
    (void)total;
    return true;
}


// ---------------------------------------------------------------------------------------------------------------------
// Top-level acceleration structure: BuildRaytracingAccelerationStructure for
// D3D12_RAYTRACING_ACCELERATION_STRUCTURE_TYPE_TOP_LEVEL (GpuBVH2Builder.cpp:116-146, SceneType::BottomLevelBVHs).
//   load        TopLevelLoadAABBs.hlsli:58-100: per instance, the bottom-level root box transformed to world space
//               (TransformAABB, RayTracingHelper.hlsli:318-344: the 8 corners through ObjectToWorld), the desc's
//               transform replaced by its inverse (InverseAffineTransform :297-316), ObjectToWorld kept beside it
//   scene box   CalculateSceneAABBFromBVHs.hlsl (min / max of the instance boxes, read back as centre -+ half)
//   morton      CalculateMortonCodesForAABBs.hlsl: GetCentroid = the stored box centre
//   sort, rearrange (RearrangeBVHs.hlsl), Karras hierarchy: as for triangles; NO treelet pass (GpuBVH2Builder.cpp:343)
//   boxes       TopLevelComputeAABBs.hlsl: a leaf's box is recomputed from its instance's bottom-level root box and
//               ObjectToWorld; internal nodes as for triangles (smaller subtree left, D1 on ties)
// Layout: 16-byte header {offsetToBoxes = 16, offsetToLeafNodeMetaData, 0, totalSize} (TopLevelPrepareForComputeAABBs.hlsl
// stores three of the four words; OffsetToLeafNodeMetaDataOffset = 4, RayTracingHelper.hlsli:46), 32-byte nodes (internal [0, N-1), leaves [N-1, 2N-1)), 116-byte BVHMetadata per sorted
// leaf (RayTracingHlslCompat.h:217-236: the instance desc with WorldToObject, ObjectToWorld, the original instance index).
namespace {
struct Mat34 { float m[3][4]; };
// mul(float3x4, float4): one dot product per row, left to right (the pin of the primitive load's transform)
inline f3 mul_point(const Mat34& a, f3 p, float w) {
    return mk3(((a.m[0][0] * p.x + a.m[0][1] * p.y) + a.m[0][2] * p.z) + a.m[0][3] * w,
               ((a.m[1][0] * p.x + a.m[1][1] * p.y) + a.m[1][2] * p.z) + a.m[1][3] * w,
               ((a.m[2][0] * p.x + a.m[2][1] * p.y) + a.m[2][2] * p.z) + a.m[2][3] * w);
}
// Determinant, RayTracingHelper.hlsli:287-295 (left to right)
inline float determinant(const Mat34& t) {
    return ((((t.m[0][0] * t.m[1][1] * t.m[2][2] -
               t.m[0][0] * t.m[2][1] * t.m[1][2]) -
              t.m[1][0] * t.m[0][1] * t.m[2][2]) +
             t.m[1][0] * t.m[2][1] * t.m[0][2]) +
            t.m[2][0] * t.m[0][1] * t.m[1][2]) -
           t.m[2][0] * t.m[1][1] * t.m[0][2];
}
// InverseAffineTransform, RayTracingHelper.hlsli:297-316, term by term (the 0.0f / 1.0f factors of the implicit fourth
// row stay in the text: they only matter for the sign of zeros and for non-finite entries)
inline Mat34 inverse_affine(const Mat34& a) {
    const float (*t)[4] = a.m;
    const float invDet = 1.0f / determinant(a);
    Mat34 r;
    r.m[0][0] = invDet * ((t[1][1] * (t[2][2] * 1.0f - 0.0f * t[2][3]) + t[2][1] * (0.0f * t[1][3] - t[1][2] * 1.0f)) + 0.0f * (t[1][2] * t[2][3] - t[2][2] * t[1][3]));
    r.m[1][0] = invDet * ((t[1][2] * (t[2][0] * 1.0f - 0.0f * t[2][3]) + t[2][2] * (0.0f * t[1][3] - t[1][0] * 1.0f)) + 0.0f * (t[1][0] * t[2][3] - t[2][0] * t[1][3]));
    r.m[2][0] = invDet * ((t[1][3] * (t[2][0] * 0.0f - 0.0f * t[2][1]) + t[2][3] * (0.0f * t[1][1] - t[1][0] * 0.0f)) + 1.0f * (t[1][0] * t[2][1] - t[2][0] * t[1][1]));
    r.m[0][1] = invDet * ((t[2][1] * (t[0][2] * 1.0f - 0.0f * t[0][3]) + 0.0f * (t[2][2] * t[0][3] - t[0][2] * t[2][3])) + t[0][1] * (0.0f * t[2][3] - t[2][2] * 1.0f));
    r.m[1][1] = invDet * ((t[2][2] * (t[0][0] * 1.0f - 0.0f * t[0][3]) + 0.0f * (t[2][0] * t[0][3] - t[0][0] * t[2][3])) + t[0][2] * (0.0f * t[2][3] - t[2][0] * 1.0f));
    r.m[2][1] = invDet * ((t[2][3] * (t[0][0] * 0.0f - 0.0f * t[0][1]) + 1.0f * (t[2][0] * t[0][1] - t[0][0] * t[2][1])) + t[0][3] * (0.0f * t[2][1] - t[2][0] * 0.0f));
    r.m[0][2] = invDet * ((0.0f * (t[0][2] * t[1][3] - t[1][2] * t[0][3]) + t[0][1] * (t[1][2] * 1.0f - 0.0f * t[1][3])) + t[1][1] * (0.0f * t[0][3] - t[0][2] * 1.0f));
    r.m[1][2] = invDet * ((0.0f * (t[0][0] * t[1][3] - t[1][0] * t[0][3]) + t[0][2] * (t[1][0] * 1.0f - 0.0f * t[1][3])) + t[1][2] * (0.0f * t[0][3] - t[0][0] * 1.0f));
    r.m[2][2] = invDet * ((1.0f * (t[0][0] * t[1][1] - t[1][0] * t[0][1]) + t[0][3] * (t[1][0] * 0.0f - 0.0f * t[1][1])) + t[1][3] * (0.0f * t[0][1] - t[0][0] * 0.0f));
    r.m[0][3] = invDet * ((t[0][1] * (t[2][2] * t[1][3] - t[1][2] * t[2][3]) + t[1][1] * (t[0][2] * t[2][3] - t[2][2] * t[0][3])) + t[2][1] * (t[1][2] * t[0][3] - t[0][2] * t[1][3]));
    r.m[1][3] = invDet * ((t[0][2] * (t[2][0] * t[1][3] - t[1][0] * t[2][3]) + t[1][2] * (t[0][0] * t[2][3] - t[2][0] * t[0][3])) + t[2][2] * (t[1][0] * t[0][3] - t[0][0] * t[1][3]));
    r.m[2][3] = invDet * ((t[0][3] * (t[2][0] * t[1][1] - t[1][0] * t[2][1]) + t[1][3] * (t[0][0] * t[2][1] - t[2][0] * t[0][1])) + t[2][3] * (t[1][0] * t[0][1] - t[0][0] * t[1][1]));
    return r;
}
// TransformAABB, RayTracingHelper.hlsli:318-344 (min / max are order independent)
inline Box transform_aabb(const Box& b, const Mat34& m) {
    Box r{mk3(FLT_MAX), mk3(-FLT_MAX)};
    for (int i = 0; i < 8; i++) {
        f3 v = mul_point(m, mk3((i & 4) ? b.mx.x : b.mn.x, (i & 2) ? b.mx.y : b.mn.y, (i & 1) ? b.mx.z : b.mn.z), 1.0f);
        r.mn = min3(r.mn, v); r.mx = max3(r.mx, v);
    }
    return r;
}
// the world-space box of an instance: bottom-level root (centre / half -> min / max), transformed, back to centre / half
inline void instance_box(const uint8_t* blas, const Mat34& objectToWorld, f3& c, f3& h) {
    const AABBNode& root = *(const AABBNode*)(blas + 16);
    f3 rc = mk3(root.c[0], root.c[1], root.c[2]), rh = mk3(root.h[0], root.h[1], root.h[2]);
    Box w = transform_aabb(Box{rc - rh, rc + rh}, objectToWorld); // BoundingBoxToAABB :237-243
    c = (w.mn + w.mx) * 0.5f;                                     // AABBtoBoundingBox :229-235
    h = w.mx - c;
}
} // namespace

bool build_tlas(const TbInstanceDesc* inst, uint32_t n, const std::vector<const uint8_t*>& blas, std::vector<uint8_t>& out, std::string& err) {
    if (n == 0) { err = "no instances"; return false; }
    struct MetaRec { float worldToObject[12]; uint32_t idAndMask, contribAndFlags, asLo, asHi; float objectToWorld[12]; uint32_t instanceIndex; };
    static_assert(sizeof(MetaRec) == 116, "BVHMetadata is 116 bytes");
    std::vector<AABBNode> leaf(n);
    std::vector<MetaRec> meta(n);
    for (uint32_t i = 0; i < n; i++) {
        if (inst[i].AccelerationStructure >= blas.size()) { err = "instance references a bottom-level structure that does not exist"; return false; }
        Mat34 o2w;
        memcpy(o2w.m, inst[i].Transform, 48);
        Mat34 w2o = inverse_affine(o2w);
        f3 c, h;
        instance_box(blas[inst[i].AccelerationStructure], o2w, c, h);
        leaf[i] = AABBNode{{c.x, c.y, c.z}, kLeafFlag | i, {h.x, h.y, h.z}, 0};
        MetaRec& m = meta[i];
        memcpy(m.worldToObject, w2o.m, 48);
        m.idAndMask = inst[i].InstanceIDAndMask; m.contribAndFlags = inst[i].InstanceContributionToHitGroupIndexAndFlags;
        m.asLo = (uint32_t)inst[i].AccelerationStructure; m.asHi = (uint32_t)(inst[i].AccelerationStructure >> 32);
        memcpy(m.objectToWorld, o2w.m, 48);
        m.instanceIndex = i;
    }
    f3 smin = mk3(FLT_MAX), smax = mk3(-FLT_MAX);
    for (uint32_t i = 0; i < n; i++) { // CalculateSceneAABBFromBVHs.hlsl: RawDataToAABB of the stored box
        f3 c = mk3(leaf[i].c[0], leaf[i].c[1], leaf[i].c[2]), h = mk3(leaf[i].h[0], leaf[i].h[1], leaf[i].h[2]);
        smin = min3(c - h, smin); smax = max3(c + h, smax);
    }
    std::vector<uint32_t> codes(n), order(n);
    for (uint32_t i = 0; i < n; i++) { codes[i] = morton_code(mk3(leaf[i].c[0], leaf[i].c[1], leaf[i].c[2]), smin, smax); order[i] = i; }
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return sorts_before(codes[a], a, codes[b], b); });
    std::vector<uint32_t> sc(n);
    std::vector<MetaRec> sm(n);
    for (uint32_t i = 0; i < n; i++) { sc[i] = codes[order[i]]; sm[i] = meta[order[i]]; }
    const uint32_t nInternal = n - 1, total = 2 * n - 1;
    std::vector<HNode> H(total, HNode{0xffffffffu, 0, 0});
    if (n > 1) {
        Karras K{sc.data(), n};
        for (uint32_t idx = 0; idx < nInternal; idx++) {
            uint32_t first, last;
            K.range(idx, first, last);
            uint32_t split = K.split(first, last);
            uint32_t a = (split == first) ? nInternal + split : split;
            uint32_t b = (split + 1 == last) ? nInternal + split + 1 : split + 1;
            H[idx].left = a; H[idx].right = b; H[a].parent = idx; H[b].parent = idx;
        }
    }
    const uint32_t offBoxes = 16, offMeta = offBoxes + 32 * total, totalSize = offMeta + 116 * n;
    out.assign(totalSize, 0);
    uint32_t header[4] = {offBoxes, offMeta, 0, totalSize}; // OffsetToLeafNodeMetaDataOffset = 4 (RayTracingHelper.hlsli:46), word 2 is not written
    memcpy(out.data(), header, 16);
    AABBNode* nodes = (AABBNode*)(out.data() + offBoxes);
    memcpy(out.data() + offMeta, sm.data(), 116ull * n);
    for (uint32_t i = 0; i < n; i++) { // TopLevelComputeAABBs.hlsl ComputeLeafAABB: from the metadata of the sorted leaf
        Mat34 o2w;
        memcpy(o2w.m, sm[i].objectToWorld, 48);
        f3 c, h;
        instance_box(blas[sm[i].asLo], o2w, c, h);
        nodes[nInternal + i] = AABBNode{{c.x, c.y, c.z}, i | kLeafFlag, {h.x, h.y, h.z}, 1};
    }
    if (n > 1) {
        std::vector<uint32_t> cnt(total, 1);
        std::vector<std::pair<uint32_t, int>> st;
        st.push_back({0, 0});
        while (!st.empty()) {
            auto& top = st.back();
            uint32_t node = top.first;
            if (node >= nInternal) { st.pop_back(); continue; }
            if (top.second == 0) { top.second = 1; st.push_back({H[node].left, 0}); }
            else if (top.second == 1) { top.second = 2; st.push_back({H[node].right, 0}); }
            else {
                uint32_t l = H[node].left, r = H[node].right;
                if (cnt[l] > cnt[r]) std::swap(l, r);
                cnt[node] = cnt[l] + cnt[r];
                f3 c, h;
                parent_box(mk3(nodes[l].c[0], nodes[l].c[1], nodes[l].c[2]), mk3(nodes[l].h[0], nodes[l].h[1], nodes[l].h[2]),
                           mk3(nodes[r].c[0], nodes[r].c[1], nodes[r].c[2]), mk3(nodes[r].h[0], nodes[r].h[1], nodes[r].h[2]), c, h);
                nodes[node] = AABBNode{{c.x, c.y, c.z}, l & 0x3fffffffu, {h.x, h.y, h.z}, r};
                st.pop_back();
            }
        }
    }
    return true;
}
// test hooks: the pure functions of the instance load
void inverse_affine_public(const float* m12, float* out12) { Mat34 a; memcpy(a.m, m12, 48); Mat34 r = inverse_affine(a); memcpy(out12, r.m, 48); }
void transform_aabb_public(const float* mn, const float* mx, const float* m12, float* out6) {
    Mat34 a; memcpy(a.m, m12, 48);
    Box r = transform_aabb(Box{mk3(mn[0], mn[1], mn[2]), mk3(mx[0], mx[1], mx[2])}, a);
    out6[0] = r.mn.x; out6[1] = r.mn.y; out6[2] = r.mn.z; out6[3] = r.mx.x; out6[4] = r.mx.y; out6[5] = r.mx.z;
}

} // namespace oracle
