// device_types.h — device-side views of the scene, BVH and render state shared by the
// CUDA translation units and the host API glue.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "tracerboy_b200.h"

namespace tbd {

// Reference BVH layout (RayTracingHlslCompat.h:344-398): 16 B header, 32 B nodes
// (internal [0,N-1), leaves [N-1,2N-1)), 40 B primitives, 12 B metadata.
struct RefNode { float c[3]; uint32_t flags; float h[3]; uint32_t right; };
struct HNode { uint32_t parent, left, right; };

// Traversal layout, derived 1:1 from the reference BVH2 after the build (same boxes,
// same child order, same leaf order), arranged for 128-bit loads:
//   PairNode (64 B, one per internal node): both child boxes + both child references.
//     A child reference is the reference node index for an internal child, or
//     0x80000000 | sortedPrimSlot for a leaf child.
//   WideTri (48 B, one per sorted primitive): v0|geomIndex, v1|primIndex, v2|0.
struct PairNode {
    float4 lc;  // left centre xyz,  w = left child ref (bits)
    float4 lh;  // left half  xyz,   w = right child ref (bits)
    float4 rc;  // right centre xyz, w unused
    float4 rh;  // right half xyz,   w unused
};
struct WideTri { float4 v0, v1, v2; };

// Entries of the per-ray traversal stack. A near-first BVH2 traversal keeps at most one waiting far child per tree
// level, so a tree of depth <= TB_STACK_DEPTH can never overflow it; deeper trees are rejected after the build.
#define TB_STACK_DEPTH 96

struct DeviceBvh {
    uint8_t* ref = nullptr;      // reference-layout bytes
    uint64_t refBytes = 0;
    PairNode* pairs = nullptr;   // numPrims-1 entries (>=1 allocated)
    WideTri* tris = nullptr;     // numPrims entries
    RefNode root;                // root node copy (box for the initial test; leaf flag if N==1)
    uint32_t numPrims = 0;
    uint32_t depth = 0;          // height of the tree (0 = a single leaf): bounds the traversal stack
};

// Two-level query: one record per SORTED leaf of a top-level structure (tlas.cpp writes them behind the reference bytes)
#define TB_TLAS_STACK_DEPTH 64
struct TlasInstanceRecord {
    float worldToObject[12];   // row-major 3x4 (the inverted instance transform)
    const PairNode* pairs;     // the instance's bottom-level structure, traversal layout
    const WideTri* tris;
    uint32_t rootRef;          // 0 (internal root) or 0x80000000 | slot (a single-triangle bottom level)
    uint32_t instanceIndex;    // index of the instance in the caller's array (SoftwareHitData::InstanceIndex)
    uint32_t mask;             // InstanceMask (8 bit): 0 = never hit
    uint32_t pad;
};

// What the top-level build needs to know about the bottom-level structure of one instance (resolved on the host from the
// structure's address, uploaded with the instance descs).
struct TlasBlasInfo {
    float c[3]; uint32_t rootRef;  // root box centre; 0 (internal root) or 0x80000000 | slot (a single-triangle bottom level)
    float h[3]; uint32_t pad;      // root box half extent
    const PairNode* pairs;
    const WideTri* tris;
};

// One geometry of a build as the load kernel reads it: the caller's device buffers
// (D3D12_RAYTRACING_GEOMETRY_TRIANGLES_DESC: vertex buffer + stride, optional 16 / 32 bit indices, optional 3x4 transform).
struct BuildGeometry {
    const uint8_t* positions;  // first vertex; float3 at every strideBytes
    const void* indices;       // nullptr when indexFormat == 0
    const float* transform;    // 12 floats, row-major 3x4, or nullptr
    uint32_t strideBytes;
    uint32_t indexFormat;      // 0 none, 2 uint16, 4 uint32
    uint32_t flags;            // D3D12_RAYTRACING_GEOMETRY_FLAGS
    uint32_t pad;
};

struct DeviceScene {
    const TbGeometryRecord* geoms = nullptr;
    const float* positions = nullptr; // float3 pooled
    const TbVertex* vertices = nullptr;
    const uint32_t* indices = nullptr;
    const TbMaterial* materials = nullptr;
    const TbLight* lights = nullptr;
    const TbTextureData* textures = nullptr;
    // images: table of {ptr, w, h, format}
    struct ImageRef { const void* data; uint32_t width, height, format; };
    const ImageRef* images = nullptr;
    const uint8_t* geomClass = nullptr; // per geometry: material class of the shading stage's queue (pathtrace.cu: material_class)
    const uint8_t* blueNoise = nullptr; // 2 x 256 x 256 x RGBA8
    uint32_t numGeoms = 0, numMaterials = 0, numLights = 0, numTextures = 0, numImages = 0;
    int32_t envImage = -1;
    uint32_t flipTextureUVs = 0;
    float envTransform[3][4];
    float envColorScale[3];
};

struct FrameConstants {
    TbOutputSettings settings;
    TbCamera camera;
    float time;
    uint32_t frame;
    uint32_t width, height;
    int32_t selectedX, selectedY;
    float halton2, halton3; // Halton23(frame), RayGenCommon.h:79-82 (per-frame constant)
    uint32_t rowOffset, rowStride; // row-band shard: bands of 8 rows, band b is rendered iff b % rowStride == rowOffset
    uint32_t worldPosSlot;  // which of the two world-position buffers this frame writes (local sample index & 1)
    uint32_t aovMask;       // AOV_FULL / AOV_WORLDPOS ownership of this frame (frames run concurrently)
    uint32_t clearAccum;    // 1 on the first frame this handle renders after an invalidate. Equals
                            // (GlobalFrameCount == 0) on one GPU; differs only under sample sharding.
};

} // namespace tbd
