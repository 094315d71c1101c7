// ref_load.cpp — CPU ORACLE (test infrastructure): the builder's first stage compiled from the mount —
// BottomLevelLoadTriangles.hlsli (the index readers for 32-bit, 16-bit and absent index buffers, GetVertex,
// TransformVertex, main()), GetOutputIndex / StorePrimitiveMetadata of LoadPrimitivesBindings.h and TriangleToRawData /
// NullPrimitive / CreateTrianglePrimitive of RayTracingHlslCompat.h — pre-passed into oracle/_ref/load_*.inc by
// prepass.run_load. GetIndex and main() are included three times, with INDEX_BUFFER_32_BIT, INDEX_BUFFER_16_BIT and
// NO_INDEX_BUFFER, as the three LoadTriangles*.hlsl files do. Restated: the resources, the constants, the host's
// dispatch loop over the geometry descs (LoadPrimitivesPass.cpp:70-166: PrimitiveOffset accumulates,
// GeometryContributionToHitGroupIndex = element index, the adjustment for a 2-byte-aligned 16-bit index buffer), and
// mul(float3x4, float4) as the product's own transform (one dot product per row, left to right).
#define RC_TRAVERSE 1
#define RC_LOAD 1
#include "hlsl_compat.h"
#include <cstring>
#include <vector>

namespace refcore {

struct uint2 { uint x, y; uint operator[](uint i) const { return i == 0 ? x : y; } };
struct uint3 {
    uint x, y, z;
    uint3() : x(0), y(0), z(0) {}
    uint3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    uint& operator[](uint i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct uint4 { uint x, y, z, w; uint4& operator=(uint v) { x = y = z = w = v; return *this; } };
struct float3x4 { float4 r[3]; float4& operator[](int i) { return r[i]; } };
inline float3 mul(float3x4 m, float4 v) {
    return float3(((m.r[0].x * v.x + m.r[0].y * v.y) + m.r[0].z * v.z) + m.r[0].w * v.w,
                  ((m.r[1].x * v.x + m.r[1].y * v.y) + m.r[1].z * v.z) + m.r[1].w * v.w,
                  ((m.r[2].x * v.x + m.r[2].y * v.y) + m.r[2].z * v.z) + m.r[2].w * v.w);
}
inline float asfloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline float3 asfloat(uint3 u) { return float3(asfloat(u.x), asfloat(u.y), asfloat(u.z)); }
inline uint asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline uint4 asuint(float4 f) { uint4 u; u.x = asuint(f.x); u.y = asuint(f.y); u.z = asuint(f.z); u.w = asuint(f.w); return u; }

struct ByteAddressBuffer {
    const uint8_t* bytes = nullptr;
    uint2 Load2(uint o) const { uint2 v; memcpy(&v, bytes + o, 8); return v; }
    uint3 Load3(uint o) const { uint3 v; memcpy(&v, bytes + o, 12); return v; }
};
struct Triangle { float3 v0, v1, v2; };                                         // RayTracingHlslCompat.h:90-95
struct Primitive { uint PrimitiveType; uint4 data0; uint4 data1; uint data2; }; // :122-136 (HLSL side), 40 bytes
struct PrimitiveMetaData { uint GeometryContributionToHitGroupIndex; uint PrimitiveIndex; uint GeometryFlags; };
struct LoadPrimitivesInputConstants { // LoadPrimitivesBindings.h:19-33
    uint ElementBufferStride, IndexBufferOffset, NumPrimitivesBound, PrimitiveOffset, TotalPrimitiveCount,
        GeometryContributionToHitGroupIndex, HasValidTransform, GeometryFlags, PerformUpdate;
};
static Primitive* PrimitiveBuffer;
static PrimitiveMetaData* MetadataBuffer;
static uint* CachedSortBuffer;
static ByteAddressBuffer elementBuffer, indexBuffer;
static const float4* TransformBuffer;
static LoadPrimitivesInputConstants Constants;
#define TRIANGLE_TYPE 0x1
#define SizeOfUINT16 2
#define SizeOfUINT32 4
#define NumberOfVerticesPerTriangle 3

#include "../_ref/load_common_gen.inc"

namespace idx32 {
#define INDEX_BUFFER_32_BIT
#include "../_ref/load_variant_gen.inc"
#undef INDEX_BUFFER_32_BIT
}
namespace idx16 {
#define INDEX_BUFFER_16_BIT
#include "../_ref/load_variant_gen.inc"
#undef INDEX_BUFFER_16_BIT
}
namespace noidx {
#define NO_INDEX_BUFFER
#include "../_ref/load_variant_gen.inc"
#undef NO_INDEX_BUFFER
}

} // namespace refcore

// One geometry desc, as tb_bvh_build / D3D12_RAYTRACING_GEOMETRY_TRIANGLES_DESC give it.
struct RefGeometryDesc {
    const void* positions; uint32_t strideBytes; uint32_t vertexCount;
    const void* indices; uint32_t indexFormat; /* 0, 2 or 4 bytes */ uint32_t indexCount;
    const float* transform3x4; uint32_t geometryFlags;
};

// LoadPrimitivesPass::LoadPrimitives over all descs. Index "GPU addresses" are byte offsets here: `indices` may be
// 2-byte aligned for 16-bit indices, which the host loop turns into an aligned base + IndexBufferOffset = 2.
extern "C" __attribute__((visibility("default")))
int ref_load_primitives(const RefGeometryDesc* descs, uint32_t numDescs, void* prims40, void* meta12) {
    using namespace refcore;
    static_assert(sizeof(Primitive) == 40 && sizeof(PrimitiveMetaData) == 12, "layouts");
    uint32_t total = 0;
    for (uint32_t e = 0; e < numDescs; e++) total += (descs[e].indexFormat == 0 ? descs[e].vertexCount : descs[e].indexCount) / 3;
    PrimitiveBuffer = (Primitive*)prims40; MetadataBuffer = (PrimitiveMetaData*)meta12; CachedSortBuffer = nullptr;
    uint32_t loaded = 0;
    for (uint32_t e = 0; e < numDescs; e++) {
        const RefGeometryDesc& d = descs[e];
        const bool nullIndex = d.indexFormat == 0;
        if (d.indices == nullptr && !nullIndex) return -1; // E_INVALIDARG, LoadPrimitivesPass.cpp:77-80
        const uint32_t count = (nullIndex ? d.vertexCount : d.indexCount) / 3;
        uintptr_t indexVA = (uintptr_t)d.indices;
        uint32_t indexOffset = 0;
        if (indexVA % 4 == 2) { indexVA -= 2; indexOffset = 2; }
        memset(&Constants, 0, sizeof(Constants));
        Constants.IndexBufferOffset = indexOffset; Constants.NumPrimitivesBound = count; Constants.TotalPrimitiveCount = total;
        Constants.PrimitiveOffset = loaded; Constants.ElementBufferStride = d.strideBytes;
        Constants.GeometryContributionToHitGroupIndex = e; Constants.HasValidTransform = d.transform3x4 != nullptr;
        Constants.GeometryFlags = d.geometryFlags; Constants.PerformUpdate = 0;
        elementBuffer.bytes = (const uint8_t*)d.positions; indexBuffer.bytes = (const uint8_t*)indexVA;
        float4 rows[3];
        if (d.transform3x4) {
            for (int r = 0; r < 3; r++) rows[r] = float4(d.transform3x4[4 * r], d.transform3x4[4 * r + 1], d.transform3x4[4 * r + 2], d.transform3x4[4 * r + 3]);
            TransformBuffer = rows;
        }
        for (uint32_t t = 0; t < count; t++) {
            if (nullIndex) noidx::load_main(uint3(t, 0, 0));
            else if (d.indexFormat == 2) idx16::load_main(uint3(t, 0, 0));
            else idx32::load_main(uint3(t, 0, 0));
        }
        loaded += count;
    }
    return (int)loaded;
}
