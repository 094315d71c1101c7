"""Pre-pass that turns the reference's TracerBoy/kernel.glsl (read from the mount, never copied
into the repository) into text a C++ compiler accepts against hlsl_compat.h:

  1. C preprocessor with IS_SHADER_TOY=0 (the configuration RayGenCommon.h:6 selects);
  2. GLSL/HLSL parameter qualifiers:  out T x / inout T x -> T& x,  in T x -> T x;
  3. multi-component swizzles become method calls (.xy -> .xy());
  4. the two call sites that draw two rand() values inside one argument list are sequenced
     left to right, as DXC evaluates them (g++ evaluates right to left; SURVEY §8c trap 17).

The output goes to oracle/_ref/ (git-ignored) and is compiled there; nothing is written anywhere else.
"""
import os
import re
import subprocess
import sys


def run(src, dst):
    text = subprocess.run(["gcc", "-E", "-P", "-x", "c", "-DIS_SHADER_TOY=0", src], check=True,
                          stdout=subprocess.PIPE, text=True).stdout
    text = re.sub(r"\b(?:inout|out)\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1& \2", text)
    text = re.sub(r"\bin\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1 \2", text)
    text = re.sub(r"\.(xyz|rgb|xy)\b(?!\s*\()", r".\1()", text)
    hoists = {
        "GenerateCosineWeightedDirection(normal, rand(), rand(), pdfValue)":
            "[&]{ float r0_ = rand(); float r1_ = rand(); return GenerateCosineWeightedDirection(normal, r0_, r1_, pdfValue); }()",
        "GenerateImportanceSampledDirection(normal, roughness, rand(), rand(), PDFValue)":
            "[&]{ float r0_ = rand(); float r1_ = rand(); return GenerateImportanceSampledDirection(normal, roughness, r0_, r1_, PDFValue); }()",
    }
    for a, b in hoists.items():
        if a not in text:
            raise SystemExit("prepass: expected call site not found: " + a)
        text = text.replace(a, b)
    if re.search(r"\([^()]*rand\(\)[^()]*rand\(\)[^()]*\)", text):
        raise SystemExit("prepass: an unsequenced pair of rand() calls is left in one argument list")
    open(dst, "w").write(text)


def run_post(tonemap_h, postprocess_hlsl, dst):
    """Tonemap.h whole + the Process* functions of PostProcessCS.hlsl (everything between the root-signature
    macro and the entry point; the resource declarations and main()'s switch cannot be compiled and are
    restated in ref_post.cpp)."""
    t = open(tonemap_h).read().replace("#pragma once", "")
    p = open(postprocess_hlsl).read()
    a, b = p.index("float3 ProcessLit(float4 color)"), p.index("[numthreads(8, 8, 1)]")
    text = t + "\n" + p[a:b]
    text = re.sub(r"\.(xyz|rgb|xy)\b(?!\s*\()", r".\1()", text)
    open(dst, "w").write(text)


def run_temporal(tonemap_h, temporal_hlsl, dst):
    """ColorToLuma of Tonemap.h + PlaneIntersection and the whole body of main() of TemporalAccumulationCS.hlsl (with its
    NEIGHBORHOOD_CLAMPING / WORLD_POSITION_HISTORY_REJECTION switches). The resource declarations, the root signature
    and the entry-point attributes cannot be compiled; the resources are shims in ref_temporal.cpp. The unused
    SampleTextureCatmullRom is skipped."""
    t = open(tonemap_h).read()
    a = t.index("float ColorToLuma(float3 color)")
    luma = t[a:t.index("}", a) + 1]
    h = open(temporal_hlsl).read()
    a, b = h.index("float PlaneIntersection("), h.index("#define ComputeRS")
    plane = h[a:b]
    a = h.index("#define NEIGHBORHOOD_CLAMPING 0")
    body = h[a:]
    for attr in ("[RootSignature(ComputeRS)]", "[numthreads(TEMPORAL_ACCUMULATION_THREAD_GROUP_WIDTH, TEMPORAL_ACCUMULATION_THREAD_GROUP_HEIGHT, 1)]"):
        if attr not in body:
            raise SystemExit("prepass: expected attribute not found: " + attr)
        body = body.replace(attr, "")
    entry = "void main( uint3 DTid : SV_DispatchThreadID )"
    if entry not in body:
        raise SystemExit("prepass: entry point not found")
    body = body.replace(entry, "static void shader_main(uint3 DTid)")
    text = luma + "\n" + plane + "\n" + body
    text = re.sub(r"\.(xyz|rgb|xy|rg)\b(?!\s*\()", r".\1()", text)
    open(dst, "w").write(text)


def run_hist(tonemap_h, generate_hlsl, average_hlsl, dst_gen, dst_avg):
    """GenerateHistogramCS.hlsl from the groupshared declaration to the end (LuminanceToHistogramIndex + main()) with
    Tonemap.h's ColorToLuma in front, and CalculateAveragedLuminanceCS.hlsl from its groupshared declaration to the end.
    Resource declarations are shims in ref_hist.cpp; the entry points become plain functions."""
    t = open(tonemap_h).read()
    a = t.index("float ColorToLuma(float3 color)")
    luma = t[a:t.index("}", a) + 1]
    g = open(generate_hlsl).read()
    body = g[g.index("groupshared uint GroupHistogram[NUM_HISTOGRAM_BINS];"):]
    for old, new in (("[numthreads(GENERATE_HISTOGRAM_THREAD_GROUP_WIDTH, GENERATE_HISTOGRAM_THREAD_GROUP_HEIGHT, 1)]", ""),
                     ("void main( uint3 DTid : SV_DispatchThreadID, uint Gid : SV_GroupIndex )", "static void shader_main(uint3 DTid, uint Gid)")):
        if old not in body:
            raise SystemExit("prepass: expected text not found: " + old)
        body = body.replace(old, new)
    text = luma + "\n" + body
    text = re.sub(r"\.(xyz|rgb|xy)\b(?!\s*\()", r".\1()", text)
    open(dst_gen, "w").write(text)
    c = open(average_hlsl).read()
    body = c[c.index("groupshared uint AveragedHistogramCount;"):]
    for old, new in (("[numthreads(CALCULATE_AVERAGED_LUMINANCE_THREAD_GROUP_WIDTH, CALCULATE_AVERAGED_LUMINANCE_THREAD_GROUP_HEIGHT, 1)]", ""),
                     ("void main(uint Gid : SV_GroupIndex )", "static void shader_main(uint Gid)")):
        if old not in body:
            raise SystemExit("prepass: expected text not found: " + old)
        body = body.replace(old, new)
    open(dst_avg, "w").write(body)


def run_traverse(src, dst_box, dst_rest):
    """The three pure functions of the fallback layer's ray query (TraverseFunction.hlsli): RayBoxTest into one
    file (compiled with contraction on: the pinned slab test is one fma per product), GetRayData and the watertight
    RayTriangleIntersect into another (compiled unfused: `precise`)."""
    t = open(src).read()
    a, b = t.index("inline\nbool RayBoxTest("), t.index("float3 Swizzle(float3 v, int3 swizzleOrder)")
    c = t.index("#define MULTIPLE_LEAVES_PER_NODE")
    d, e = t.index("int GetIndexOfBiggestChannel(float3 vec)"), t.index("#define TOP_LEVEL_INDEX")
    f, g = t.index("struct RayData"), t.index("bool Cull(bool opaque, uint rayFlags)")

    def fix(text):
        text = re.sub(r"\b(?:inout|out)\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1& \2", text)
        text = re.sub(r"\bin\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1 \2", text)
        text = re.sub(r"\b(\w+)\.xy\s*=\s*([^;]+);", r"\1.set_xy(\2);", text)   # swizzle on the left-hand side
        text = re.sub(r"\.(xyz|rgb|xy)\b(?!\s*\()", r".\1()", text)
        return text
    open(dst_box, "w").write(fix(t[a:b]))
    open(dst_rest, "w").write(fix(t[d:e] + t[f:g] + t[b:c]))


def run_traverse_loop(helper_hlsli, traverse_hlsli, dst):
    """The fallback layer's ray query LOOP: from RayTracingHelper.hlsli the node / primitive readers (IsLeafFlag ...
    BVHReadTriangle), from TraverseFunction.hlsli everything except the three pure functions that are compiled on their
    own (RayBoxTest, RayTriangleIntersect with Swizzle / IsPositive, GetRayData with its helpers): SoftwareRayDesc /
    SoftwareHitData / SoftwareRayQuery, the stack helpers, IsOpaque, TestLeafNodeIntersections, struct RayData, Cull,
    the flag helpers, Traverse and SoftwareRayQuery::Proceed. The #if FAST_PATH / DISABLE_ANYHIT /
    DISABLE_PROCEDURAL_GEOMETRY switches stay in the text and are set by ref_traverse_loop.cpp the way RayGenCommon.h:355-362
    sets them."""
    h = open(helper_hlsli).read()
    helper = h[h.index("static const int IsLeafFlag = 0x80000000;"):h.index("BoundingBox AABBtoBoundingBox(AABB aabb)")]
    t = open(traverse_hlsli).read()
    box = t.index("inline\nbool RayBoxTest(")
    leaves = t.index("#define MULTIPLE_LEAVES_PER_NODE")
    biggest = t.index("int GetIndexOfBiggestChannel(float3 vec)")
    levels = t.index("#define TOP_LEVEL_INDEX")
    raydata = t.index("struct RayData")
    getraydata = t.index("RayData GetRayData(float3 rayOrigin, float3 rayDirection)")
    cull = t.index("bool Cull(bool opaque, uint rayFlags)")
    head = t[t.index("#define INSTANCE_FLAG_NONE"):box]
    # strip the comment banner that introduces RayBoxTest from the tail of `head` (plain text, harmless) -- kept as is
    text = helper + "\n" + head + "\n" + t[leaves:biggest] + "\n" + t[levels:raydata] + "\n" + t[raydata:getraydata] + \
        "\nRayData GetRayData(float3 rayOrigin, float3 rayDirection);\n" + t[cull:]
    for old, new in (("const GpuVA nullptr = GpuVA(0, 0);", "const GpuVA nullptr_va = GpuVA(0, 0);"),
                     ("Traverse(this);", "Traverse(*this);"),
                     ("uint stackPointer = 0;", "int stackPointer = 0;"),
                     ("Declare_Fallback_SetPendingAttr(BuiltInTriangleIntersectionAttributes);", "")):
        if old not in text:
            raise SystemExit("prepass: expected text not found: " + old)
        text = text.replace(old, new)
    text = re.sub(r"\b(?:inout|out)\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1& \2", text)
    text = re.sub(r"\bin\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1 \2", text)
    text = re.sub(r"\.(xyz|rgb|xy|zw)\b(?!\s*\()", r".\1()", text)
    open(dst, "w").write(text)


def run_intersect(structs_h, hitgroup_h, raygen_h, dst):
    """What sits between the ray query and the path tracer: struct RayPayload (SharedShaderStructs.h), the geometry
    fetch of SharedHitGroup.h (HitGroupShaderRecord ... GetHitInfo: shader record -> indices -> three vertices ->
    interpolated, normalised normal / tangent and uv) and IntersectWithMaxDistance of RayGenCommon.h (the text keeps its
    #if USE_INLINE_RAYTRACING / USE_SW_RAYTRACING branches; ref_traverse_loop.cpp selects the software one). Only the
    three resource declarations lose their register bindings."""
    sh = open(structs_h).read()
    a = sh.index("struct RayPayload")
    payload = sh[a:sh.index("};", a) + 2]
    hg = open(hitgroup_h).read()
    geo = hg[hg.index("#define VertexStride 8"):hg.index("Material GetMaterial_NonRecursive(int MaterialID);")]
    for old, new in (("StructuredBuffer<HitGroupShaderRecord> ShaderTable: register(t11);", "static thread_local StructuredBuffer<HitGroupShaderRecord> ShaderTable;"),
                     ("Buffer<uint> IndexBuffers[] : register(t0, space2);", "static thread_local Buffer<uint> IndexBuffers[1];"),
                     ("Buffer<float> VertexBuffers[] : register(t0, space3);", "static thread_local Buffer<float> VertexBuffers[1];")):
        if old not in geo:
            raise SystemExit("prepass: expected resource declaration not found: " + old)
        geo = geo.replace(old, new)
    rg = open(raygen_h).read()
    isect = rg[rg.index("#define MIN_T 0.001f"):rg.index("float2 IntersectAnything(")]
    text = payload + "\n" + geo + "\n" + isect
    text = re.sub(r"\b(?:inout|out)\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1& \2", text)
    text = re.sub(r"\.(xyz|rgb|xy)\b(?!\s*\()", r".\1()", text)
    open(dst, "w").write(text)


def run_morton(src, dst):
    """GetMortonCodesFromUnitCoord(float3) + CalculateMortonCode(float3): from `#define BIT(x)` to the entry point."""
    t = open(src).read()
    a = t.index("#define BIT(x) (1 << (x))")
    b = t.index("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]", a)
    text = t[a:b]
    # `uint coords[numAxis] = { adjustedCoord.y, ... }`: HLSL converts float -> uint implicitly (truncation); C++ list
    # initialisation forbids the narrowing, so the conversion is spelled out
    text = re.sub(r"\{ adjustedCoord\.y, adjustedCoord\.x, adjustedCoord\.z \}",
                  "{ (uint)adjustedCoord.y, (uint)adjustedCoord.x, (uint)adjustedCoord.z }", text)
    open(dst, "w").write(text)


def run_karras(src, dst):
    """BuildBVHSplits.hlsli from CountLeadingZeroes to GenerateHierarchy (everything between the include and main())."""
    t = open(src).read()
    a, b = t.index("int CountLeadingZeroes(uint num)"), t.index("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]")
    open(dst, "w").write(t[a:b])


def run_treelet(bindings_h, treelet_hlsl, dst):
    """TreeletReorderBindings.h: `FullTreeletSize` and everything from `#define BIT` to the end of CombineAABB (tables,
    GetBitPermutation, IsLeaf, ComputeBoxSurfaceArea, CombineAABB); TreeletReorder.hlsl: CalculateCost, the groupshared
    declarations, FormTreelet, FindOptimalPartitions, ReformTree (up to TraverseToParent)."""
    b = open(bindings_h).read()
    t = open(treelet_hlsl).read()
    b0 = "static const uint FullTreeletSize = 7;\n"
    assert b0 in b
    a1 = b.index("#define BIT(x) (1 << (x))")
    e1 = b.index("#endif", a1)
    a2, e2 = t.index("static const float CostOfRayBoxIntersection"), t.index("void TraverseToParent(")
    text = b0 + b[a1:e1] + "\n" + t[a2:e2]
    text = re.sub(r"\bin\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1 \2", text)
    text = text.replace("[unroll]", "")
    # The group is 32 threads = one wave, which executes in lockstep: every thread has written its share of the union
    # surface areas before any thread overwrites the single-leaf entries (TreeletReorder.hlsl:96-127 has no barrier
    # between the two). Host threads are not in lockstep, so the implicit ordering is made explicit.
    marker = "    AABB nodeAABB = AABBBuffer[nodeIndex];"
    assert text.count(marker) == 1
    text = text.replace(marker, "    GroupMemoryBarrierWithGroupSync(); /* wave lockstep made explicit */\n" + marker)
    open(dst, "w").write(text)


def run_boxes(src, dst):
    """RayTracingHelper.hlsli: CreateFlag, AABBtoBoundingBox, BoundingBoxToAABB, GetBoxDataFromTriangle ... GetBoxFromChildBoxes."""
    t = open(src).read()
    a0, e0 = t.index("uint2 CreateFlag(uint leftNodeIndex, uint rightNodeIndex)"), t.index("uint GetLeftNodeIndex(uint2 flag)")
    a1, e1 = t.index("BoundingBox AABBtoBoundingBox(AABB aabb)"), t.index("AABB RawDataToAABB(int4 a, int4 b)")
    a2, e2 = t.index("BoundingBox GetBoxDataFromTriangle("), t.index("float Determinant(")
    text = t[a0:e0] + t[a1:e1] + t[a2:e2]
    text = re.sub(r"\b(?:inout|out)\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1& \2", text)
    open(dst, "w").write(text)


def run_raygen(structs_h, raygen_h, dst_light, dst):
    """`struct Light` (with its HLSL-only GetPosition) and the pure functions of RayGenCommon.h: SampleEnvironmentMap,
    Halton*, GetRandomBarycentric ... GetOneLightSample, hash13."""
    sh = open(structs_h).read()
    a, b = sh.index("struct Light\n{"), sh.index("struct AreaLightData") if "struct AreaLightData" in sh else None
    e = sh.index("#define LIGHT_TYPE_DIRECTIONAL 1") + len("#define LIGHT_TYPE_DIRECTIONAL 1")
    open(dst_light, "w").write(sh[a:e] + "\n")
    t = open(raygen_h).read()
    s1, e1 = t.index("float3 SampleEnvironmentMap(float3 v)"), t.index("float rand();")
    s2, e2 = t.index("float Halton(int b, int i)"), t.index("struct BlueNoiseData")
    s3 = t.index("float3 GetRandomBarycentric()")
    e3 = t.index("float4 GetLastFrameData()") - 1                  # the function after GetOneLightSample
    s4, e4 = t.index("float hash13(vec3 p3)"), t.index("bool ShouldSkipRay()")
    text = t[s1:e1] + t[s2:e2] + t[s3:e3] + "\n" + t[s4:e4]
    text = re.sub(r"\b(?:inout|out)\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1& \2", text)
    text = re.sub(r"\bin\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1 \2", text)
    text = re.sub(r"\.(xyz|rgb|xy|yzx)\b(?!\s*\()", r".\1()", text)
    open(dst, "w").write(text)


def run_material(structs_h, tonemap_h, sharedrt_h, raygen_h, dst):
    """Material and texture fetch: the material / texture flag defines, struct Material and struct TextureData
    (SharedShaderStructs.h), GammaToLinear (Tonemap.h), GetMaterial_NonRecursive ... GetTextureData_Recursive
    (SharedRaytracing.h:55-137: image / checker / scale textures, the uv flip, the gamma flag) and GetDetailNormal +
    GetMaterialInternal (RayGenCommon.h:273-341: mix materials, albedo / emissive / specular-map overrides)."""
    sh = open(structs_h).read()
    a = sh.index("#define DEFAULT_MATERIAL_FLAG 0x0")
    b = sh.index("struct TextureData")
    structs = sh[a:sh.index("};", b) + 2]
    t = open(tonemap_h).read()
    a = t.index("float3 GammaToLinear(float3 color)")
    gamma = t[a:t.index("}", a) + 1]
    rt = open(sharedrt_h).read()
    tex = rt[rt.index("Material GetMaterial_NonRecursive(int MaterialID)"):]
    rg = open(raygen_h).read()
    mat = rg[rg.index("float3 GetDetailNormal(Material mat, float3 normal, float3 tangent, float2 uv)"):rg.index("struct Ray\n{")]
    text = structs + "\n" + gamma + "\n" + tex + "\n" + mat
    if "data.rgb = GammaToLinear(data.rgb);" not in text:
        raise SystemExit("prepass: expected swizzle assignment not found")
    text = re.sub(r"\b(\w+)\.rgb\s*=\s*([^;]+);", r"\1.set_rgb(\2);", text)   # swizzle on the left-hand side
    text = re.sub(r"\.(xyz|rgb|xy)\b(?!\s*\()", r".\1()", text)
    open(dst, "w").write(text)


def run_refit(helper_hlsli, prepare_hlsl, bottom_hlsl, compute_hlsli, dst_prepare, dst_compute):
    """The last builder stage: main() of BottomLevelPrepareForComputeAABBs.hlsl (header offsets, the thread -> node map,
    counter clear) into one file; RayTracingHelper.hlsli's helpers (AABB_Min_Padding ... GetBoxFromChildBoxes), ComputeLeafAABB of
    BottomLevelComputeAABBs.hlsl and the whole ComputeAABBs.hlsli (node encoding, the bottom-up climb with its
    InterlockedAdd hand-over and the smaller-subtree-left rule) into another."""
    h = open(helper_hlsli).read()
    helper = h[h.index("#define AABB_Min_Padding 0.001"):h.index("float Determinant(in AffineMatrix transform)")]
    p = open(prepare_hlsl).read()
    prep = p[p.index("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]"):]
    prep = prep.replace("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]", "").replace("void main(uint3 DTid : SV_DispatchThreadID)", "static void prepare_main(uint3 DTid)")
    b = open(bottom_hlsl).read()
    leaf = b[b.index("BoundingBox ComputeLeafAABB("):b.index("#define BOTTOM_LEVEL 1")]
    c = open(compute_hlsli).read()
    comp = c[c.index("uint DivideAndRoundUp("):]
    comp = comp.replace("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]", "").replace("void main(uint3 DTid : SV_DispatchThreadID)", "static void compute_main(uint3 DTid)")
    if "static void compute_main(uint3 DTid)" not in comp or "static void prepare_main(uint3 DTid)" not in prep:
        raise SystemExit("prepass: entry points not found")

    def fix(text):
        text = text.replace("[unroll]", "")
        text = re.sub(r"\b(?:inout|out)\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1& \2", text)
        text = re.sub(r"\.(xyz|rgb|xy|zw)\b(?!\s*\()", r".\1()", text)
        return text
    open(dst_prepare, "w").write(fix(prep))
    open(dst_compute, "w").write(fix(helper + "\n" + leaf + "\n" + comp))


def run_flatten(tracerboy_cpp, tracerboy_h, structs_h, dst, dst_light):
    """The scene flatten's material and light rules: struct Light ... struct Material with the flag constants
    (SharedShaderStructs.h), struct MaterialTracker (TracerBoy.h), ConvertFloat3, ChannelAverage, ConvertSpecularToIOR,
    GetAreaLightColor, reciprocol and CreateMaterial (TracerBoy.cpp) into one file; the per-triangle body of LoadScene's
    area-light loop (from `Light light = {};` to `lightList.push_back(light);`) into another."""
    sh = open(structs_h).read()
    structs = sh[sh.index("struct Light\n"):sh.index("#define IMAGE_TEXTURE_TYPE 0")]
    structs = re.sub(r"#ifdef HLSL\n.*?#endif\n", "", structs, flags=re.S)
    th = open(tracerboy_h).read()
    tracker = th[th.index("struct MaterialTracker"):th.index("class TracerBoy\n")]
    cpp = open(tracerboy_cpp).read()
    conv3 = cpp[cpp.index("float3 ConvertFloat3(const pbrt::vec3f& v)"):cpp.index("float4 ConvertFloat4(")]
    avg = cpp[cpp.index("float ChannelAverage(const pbrt::vec3f& v)"):cpp.index("bool IsNormalizedFormat(DXGI_FORMAT format)")]
    light = cpp[cpp.index("pbrt::vec3f GetAreaLightColor(pbrt::AreaLight::SP pAreaLight)"):cpp.index("Material CreateMaterial(")]
    a = cpp.index("Material CreateMaterial(")
    create = cpp[a:cpp.index("TracerBoy::TracerBoy(", a)]
    open(dst, "w").write("\n".join([structs, tracker, conv3, avg, light, create]))
    b = cpp.index("\t\t\t\t\tLight light = {};")
    body = cpp[b:cpp.index("lightList.push_back(light);", b) + len("lightList.push_back(light);")]
    open(dst_light, "w").write(body + "\n")
    # the two geometry loops of LoadScene: per-vertex attributes (:1638-1661) and indices + flat normals (:1704-1730)
    v0 = cpp.index("for (UINT v = 0; v < pTriangleMesh->vertex.size(); v++)")
    v1 = cpp.index("auto SRVIndexIter = ResourceToSRVIndex.find(pVertexBuffer->GetGPUVirtualAddress());", v0)
    i0 = cpp.index("for (UINT i = 0; i < pTriangleMesh->index.size(); i++)", v1)
    i1 = cpp.index("auto SRVIndexIter = ResourceToSRVIndex.find(pIndexBuffer->GetGPUVirtualAddress());", i0)
    vertex_struct = sh[sh.index("struct Vertex\n"):sh.index("struct Light\n")]
    base = os.path.dirname(dst)
    open(os.path.join(base, "flatten_vertex_struct_gen.inc"), "w").write(vertex_struct)
    open(os.path.join(base, "flatten_vertices_gen.inc"), "w").write(cpp[v0:v1] + "\n")
    open(os.path.join(base, "flatten_indices_gen.inc"), "w").write(cpp[i0:i1] + "\n")
    # which geometry and transform loop iteration i of LoadScene takes (top-level shape or object instance, :1358-1376)
    # and whether the transform is baked into the vertex data (:1623-1624)
    s0 = cpp.index("std::shared_ptr<pbrt::Object> pObject;")
    s1 = cpp.index("bool bNeedToCreateGlobalBLAS = bInsertIntoGlobalBLAS && !GlobalBLASKey;", s0)
    k0 = cpp.index("bool bBakeTransformIntoVertexBuffer = bInsertIntoGlobalBLAS;")
    k1 = cpp.index("bool bNormalsProvided = pTriangleMesh->normal.size();", k0)
    open(os.path.join(base, "flatten_select_gen.inc"), "w").write(cpp[s0:s1] + "\n")
    open(os.path.join(base, "flatten_bake_gen.inc"), "w").write(cpp[k0:k1] + "\n")


def run_instance_desc(compat_h, dst):
    """The instance-desc readers the two-level walk uses (RayTracingHlslCompat.h): CreateMatrix, struct
    RaytracingInstanceDesc, struct BVHMetadata, RawDataToRaytracingInstanceDesc, LoadBVHMetadata and the Get* accessors."""
    c = open(compat_h).read()
    create = c[c.index("AffineMatrix CreateMatrix(float4 rows[3])"):c.index("static const uint D3D12_RAYTRACING_INSTANCE_FLAG_NONE")]
    structs = c[c.index("struct RaytracingInstanceDesc"):c.index("#define Store4StrideInBytes 16")]
    readers = c[c.index("RaytracingInstanceDesc RawDataToRaytracingInstanceDesc("):c.index("RaytracingInstanceDesc LoadRaytracingInstanceDesc(RWByteAddressBuffer buffer, uint offset)")]
    getters = c[c.index("uint GetInstanceContributionToHitGroupIndex(RaytracingInstanceDesc desc)"):c.index("#else\nstatic_assert(sizeof(BVHMetadata) == SizeOfBVHMetadata")]
    text = create + "\n" + structs + "\n" + readers + "\n" + getters
    # the slices cut through the header's #ifdef HLSL / #else / #endif bracketing: resolve it for HLSL by hand
    text = re.sub(r"#ifdef HLSL\n(.*?)#else\n.*?#endif\s*\n", r"\1", text, flags=re.S)
    text = "\n".join(ln for ln in text.split("\n") if ln.strip() not in ("#endif", "#ifdef HLSL"))
    text = text.replace("[unroll]", "")
    text = re.sub(r"\.(zw)\b(?!\s*\()", r".\1()", text)
    open(dst, "w").write(text)


def run_tlas(helper_hlsli, dst):
    """The arithmetic of the top-level instance load (RayTracingHelper.hlsli): AABBtoBoundingBox, BoundingBoxToAABB,
    Determinant, InverseAffineTransform, TransformAABB. float4(...) constructors become make4(...) overloads."""
    h = open(helper_hlsli).read()
    boxes = h[h.index("BoundingBox AABBtoBoundingBox(AABB aabb)"):h.index("AABB RawDataToAABB(int4 a, int4 b)")]
    affine = h[h.index("float Determinant(in AffineMatrix transform)"):h.index("static const uint OffsetToAnyHitStateId")]
    text = boxes + "\n" + affine
    text = re.sub(r"\bin\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1 \2", text)
    text = re.sub(r"\.(xyz|xy|yz)\b(?!\s*\()", r".\1()", text)
    text = text.replace("float4 boxVertices[verticesPerAABB];", "@@DECL@@")
    text = re.sub(r"\bfloat4\(", "make4(", text).replace("@@DECL@@", "float4 boxVertices[verticesPerAABB];")
    if "InverseAffineTransform" not in text or "TransformAABB" not in text:
        raise SystemExit("prepass: top-level helpers not found")
    open(dst, "w").write(text)


def run_treelet_pass(bindings_h, treelet_hlsl, clear_hlsl, find_hlsl, helper_hlsli, compat_h, dst):
    """One whole treelet-reorder pass: TreeletReorderBindings.h tables and helpers, RawDataToTriangle / GetTriangle
    (RayTracingHlslCompat.h), the box helpers of RayTracingHelper.hlsli, ClearBuffers.hlsl main(), FindTreelets.hlsl
    (ComputeLeafAABB + main()) and ALL of TreeletReorder.hlsl from CalculateCost to the end: FormTreelet,
    FindOptimalPartitions, ReformTree, TraverseToParent and main() with its 33-iteration loop."""
    b = open(bindings_h).read()
    b0 = "static const uint FullTreeletSize = 7;\n"
    assert b0 in b
    a1 = b.index("#define BIT(x) (1 << (x))")
    bind = b0 + b[a1:b.index("#endif", a1)]
    c = open(compat_h).read()
    tri = c[c.index("Triangle RawDataToTriangle(uint4 a, uint4 b, uint c)"):c.index("void TriangleToRawData(")]
    gt = c[c.index("Triangle GetTriangle(Primitive prim)"):c.index("AABB GetProceduralPrimitiveAABB(Primitive prim)")]
    h = open(helper_hlsli).read()
    helper = h[h.index("#define AABB_Min_Padding 0.001"):h.index("float Determinant(in AffineMatrix transform)")]
    cl = open(clear_hlsl).read()
    clear = cl[cl.index("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]"):]
    f = open(find_hlsl).read()
    find = f[f.index("AABB ComputeLeafAABB(uint triangleIndex)"):]
    t = open(treelet_hlsl).read()
    reorder = t[t.index("static const float CostOfRayBoxIntersection"):]
    marker = "    AABB nodeAABB = AABBBuffer[nodeIndex];"
    assert reorder.count(marker) == 1   # wave lockstep made explicit, as in run_treelet
    reorder = reorder.replace(marker, "    GroupMemoryBarrierWithGroupSync(); /* wave lockstep made explicit */\n" + marker)
    for old, new, where in (("void main(uint3 DTid : SV_DispatchThreadID)", "static void clear_main(uint3 DTid)", "clear"),
                            ("void main(uint3 DTid : SV_DispatchThreadID)", "static void find_main(uint3 DTid)", "find"),
                            ("void main(uint3 Gid : SV_GroupID, uint3 GTid : SV_GroupThreadId)", "static void reorder_main(uint3 Gid, uint3 GTid)", "reorder")):
        src = {"clear": clear, "find": find, "reorder": reorder}[where]
        if old not in src:
            raise SystemExit("prepass: entry point not found in " + where)
        src = src.replace(old, new)
        if where == "clear": clear = src
        elif where == "find": find = src
        else: reorder = src
    # the builder's front, riding along because this translation unit has the shims: the sort's comparator, the
    # centroid the Morton code is taken of, and the per-thread part of the scene box
    srcdir = os.path.dirname(find_hlsl)
    bs = open(os.path.join(srcdir, "BitonicSortCommon.hlsli")).read()
    swap = bs[bs.index("bool ShouldSwap(uint A, uint B, uint indexA, uint indexB)"):]
    swap = swap[:swap.index("\n}\n") + 3]
    mc = open(os.path.join(srcdir, "CalculateMortonCodesForPrimitives.hlsl")).read()
    cen = mc[mc.index("float3 GetCentroid(uint elementIndex)"):]
    cen = cen[:cen.index("\n}\n") + 3]
    sa = open(os.path.join(srcdir, "CalculateSceneAABBFromPrimitives.hlsl")).read()
    box = sa[sa.index("AABB CalculateSceneAABB(uint baseElementIndex)"):]
    box = box[:box.index("\n}\n") + 3]
    text = bind + "\n" + helper + "\n" + tri + "\n" + gt + "\n" + clear + "\n" + find + "\n" + reorder + "\n" + swap + "\n" + cen + "\n" + box
    text = text.replace("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]", "").replace("[numthreads(NumThreadsInGroup, 1, 1)]", "").replace("[unroll]", "")
    text = re.sub(r"\b(?:inout|out)\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1& \2", text)
    text = re.sub(r"([(,]\s*)in[ \t]+([A-Za-z_]\w*)[ \t]+([A-Za-z_]\w*)", r"\1\2 \3", text)   # parameter lists only (comments say "in", too)
    text = re.sub(r"\.(xyz|rgb|xy|zw)\b(?!\s*\()", r".\1()", text)
    open(dst, "w").write(text)


def run_load(load_hlsli, bindings_h, compat_h, dst_common, dst_variant):
    """The builder's first stage, BottomLevelLoadTriangles.hlsli: the index readers (32-bit, 16-bit with its 2-byte-aligned
    start, none), GetVertex, TransformVertex and main(); GetOutputIndex / StorePrimitiveMetadata of
    LoadPrimitivesBindings.h; TriangleToRawData / NullPrimitive / CreateTrianglePrimitive of RayTracingHlslCompat.h.
    dst_common holds what does not depend on the index format; dst_variant holds GetIndex + main(), which the three
    LoadTriangles*.hlsl files compile with INDEX_BUFFER_32_BIT / INDEX_BUFFER_16_BIT / NO_INDEX_BUFFER."""
    c = open(compat_h).read()
    raw = c[c.index("void TriangleToRawData("):c.index("#else", c.index("void TriangleToRawData("))]
    nullp = c[c.index("Primitive NullPrimitive()"):c.index("Primitive CreateProceduralGeometryPrimitive(AABB aabb)")]
    create = c[c.index("Primitive CreateTrianglePrimitive(Triangle tri)"):c.index("Triangle GetTriangle(Primitive prim)")]
    b = open(bindings_h).read()
    bind = b[b.index("uint GetOutputIndex(uint inputIndex)"):b.rindex("#endif")]
    t = open(load_hlsli).read()
    readers = t[t.index("uint3 GetUint32Index3("):t.index("uint3 GetIndex(uint threadIndex)")]
    rest = t[t.index("float3 GetVertex(ByteAddressBuffer VertexBuffer, uint index, uint stride)"):t.index("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]")]
    getindex = t[t.index("uint3 GetIndex(uint threadIndex)"):t.index("float3 GetVertex(ByteAddressBuffer VertexBuffer, uint index, uint stride)")]
    main = t[t.index("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]"):]
    main = main.replace("[numthreads(THREAD_GROUP_1D_WIDTH, 1, 1)]", "")
    if "void main( uint3 DTid : SV_DispatchThreadID )" not in main:
        raise SystemExit("prepass: load entry point not found")
    main = main.replace("void main( uint3 DTid : SV_DispatchThreadID )", "static void load_main(uint3 DTid)")

    def fix(text):
        text = re.sub(r"([(,]\s*)(?:inout|out)\s+([A-Za-z_]\w*)\s+([A-Za-z_]\w*)", r"\1\2& \3", text)
        text = re.sub(r"([(,]\s*)in[ \t]+([A-Za-z_]\w*)[ \t]+([A-Za-z_]\w*)", r"\1\2 \3", text)
        text = re.sub(r"\.(xyz|rgb|xy|zw|yz)\b(?!\s*\()", r".\1()", text)
        return text
    open(dst_common, "w").write(fix(raw + "\n" + nullp + "\n" + create + "\n" + bind + "\n" + readers + "\n" + rest))
    open(dst_variant, "w").write(fix(getindex + "\n" + main))


def run_frame(raygen_h, entry_hlsl, dst):
    """The per-pixel wrapper around PathTrace: Halton / Halton23, struct BlueNoiseData, ApplyLDSToNoise, the
    Resolution / DispatchIndex accessors and GetBlueNoise (RayGenCommon.h:48-122), the AOV writers OutputPrimaryAlbedo,
    OutputPrimaryEmissive, OutputRayStats, OutputPrimaryNormal, OutputPrimaryWorldPosition (with its statics),
    IsSelectedPixel, OutputDistanceToFirstHit, OutputMaterial, ClearAOVs (:524-654), hash13 (:662-667), RayTraceCommon
    (:690-728) and the per-pixel part of SoftwareRayTraceCS.hlsl's main() (ClearAOVs ... RayTraceCommon, :37-51)."""
    t = open(raygen_h).read()

    def between(a, b):
        i = t.index(a)
        return t[i:t.index(b, i)]
    parts = [between("float Halton(int b, int i)", "float GetRotationFactor()") if "float GetRotationFactor()" in t[t.index("float Halton(int b, int i)"):t.index("BlueNoiseData GetBlueNoise()")] else between("float Halton(int b, int i)", "float4 GetMouse()"),
             between("BlueNoiseData GetBlueNoise()", "\n}\n") + "\n}\n"]
    defs = t.index("void OutputPrimaryAlbedo(float3 albedo, float DiffuseContribution)\n{")
    tail = t[defs:]

    def fn(sig, src=tail):
        i = src.index(sig)
        return src[i:src.index("\n}\n", i) + 3]
    for sig in ("void OutputPrimaryAlbedo(float3 albedo, float DiffuseContribution)", "void OutputPrimaryEmissive(float3 emissive)",
                "void OutputRayStats(uint TrianglesTested, uint BoxesTested)", "void OutputPrimaryNormal(float3 normal)"):
        parts.append(fn(sig))
    parts.append(between("static float3 WorldPosition;", "bool IsSelectedPixel()"))
    for sig in ("bool IsSelectedPixel()", "void OutputDistanceToFirstHit(float Distance)", "void OutputMaterial(int MaterialID)", "void ClearAOVs()"):
        parts.append(fn(sig))
    parts.append(between("float hash13(vec3 p3)", "bool ShouldSkipRay()"))
    parts.append(t[t.index("void RayTraceCommon()"):])
    e = open(entry_hlsl).read()
    body = e[e.index("\tClearAOVs();"):e.index("\tRayTraceCommon();") + len("\tRayTraceCommon();")]
    parts.append("\nstatic void pixel_main()\n{\n" + body + "\n}\n")
    text = "\n".join(parts)
    # the four float2(rand(), rand()) of GetBlueNoise: DXC evaluates arguments left to right, g++ right to left
    # (SURVEY 8c trap 17); a braced initialiser list is sequenced left to right by the language
    if text.count("float2(rand(), rand())") != 4:
        raise SystemExit("prepass: expected four float2(rand(), rand()) in GetBlueNoise")
    text = text.replace("float2(rand(), rand())", "float2{rand(), rand()}")
    if re.search(r"\([^()]*rand\(\)[^()]*rand\(\)[^()]*\)", text):
        raise SystemExit("prepass: an unsequenced pair of rand() calls is left in one argument list")
    text = re.sub(r"\.(xyz|rgb|xy|zw|yzx)\b(?!\s*\()", r".\1()", text)
    open(dst, "w").write(text)


if __name__ == "__main__":
    run(sys.argv[1], sys.argv[2])
