// pbrt_import.cpp — optional scene importer: pbrt::Scene (the reference's vendored
// ingowald/pbrt-parser, a third-party dependency that is compiled in place from the
// reference mount, never copied) -> tb::Scene. Restates the behaviour of
// CreateMaterial (TracerBoy.cpp:273-505), TextureAllocator::CreateTexture (:177-251)
// and LoadScene steps 2-4 (:1243-1272, 1356-1835, 1896-1944).
//
// Built into libtb_pbrtimport.so only when the parser sources are available; the main
// library dlopen()s it for *.pbrt / *.pbf paths and reports TB_ERR_NOT_IMPL otherwise.
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <unordered_map>
#include "pbrtParser/Scene.h"
#include "scene.h"

using namespace tb;

namespace {

TbFloat3 cv3(const pbrt::vec3f& v) { return {v.x, v.y, v.z}; }
float channel_average(const pbrt::vec3f& v) { return (float)((v.x + v.y + v.z) / 3.0); } // :117-120
float specular_to_ior(float s) { return (float)((sqrt(s) + 1.0) / (1.0 - sqrt(s))); }   // :122-125

struct Importer {
    Scene& out;
    std::string dir;
    std::unordered_map<pbrt::Material*, uint32_t> matIndex; // MaterialTracker, TracerBoy.h:130-156
    bool insertInstancesIntoBLAS; // LoadScene's bInsertInstancesIntoBLAS (TracerBoy.cpp:1355; false in the reference build)
    Importer(Scene& s, const std::string& d, bool insertInstances) : out(s), dir(d), insertInstancesIntoBLAS(insertInstances) {}

    uint32_t add_material(pbrt::Material* key, const TbMaterial& m) {
        uint32_t i = (uint32_t)out.materials.size();
        matIndex[key] = i;
        out.materials.push_back(m);
        out.materialNames.push_back(key ? key->name : std::string());
        return i;
    }

    uint32_t load_image(const std::string& fileName, bool* hasAlpha) {
        // InitializeTexture, TracerBoy.cpp:2186-2246: .pfm is renamed to .hdr; .hdr / .tga have their own loaders,
        // everything else goes through WIC (PNG here; image_decode.cpp follows the loaders' format rules)
        std::string full = dir + fileName;
        std::string ext = full.size() >= 4 ? full.substr(full.size() - 4) : "";
        if (ext == ".pfm") full = full.substr(0, full.size() - 4) + ".hdr";
        Image img;
        std::string err;
        bool alpha = false;
        if (!load_image_file(full, img, &alpha, err)) throw std::runtime_error(err);
        if (hasAlpha) *hasAlpha = alpha;
        out.images.push_back(std::move(img));
        return (uint32_t)out.images.size() - 1;
    }

    // TextureAllocator::CreateTexture, TracerBoy.cpp:177-251
    uint32_t create_texture(pbrt::Texture::SP tex, bool gammaCorrect = false, bool* hasAlpha = nullptr) {
        if (!tex) return TB_INVALID_TEXTURE;
        TbTextureData t;
        memset(&t, 0, sizeof(t));
        auto img = std::dynamic_pointer_cast<pbrt::ImageTexture>(tex);
        auto chk = std::dynamic_pointer_cast<pbrt::CheckerTexture>(tex);
        auto scl = std::dynamic_pointer_cast<pbrt::ScaleTexture>(tex);
        if (img) {
            t.TextureType = TB_IMAGE_TEXTURE_TYPE;
            t.DescriptorHeapIndex = load_image(img->fileName, hasAlpha);
            // NEEDS_GAMMA_CORRECTION only for "normalized" formats (IsNormalizedFormat, :128-160, :205-209): the 8-bit
            // ones, sRGB included (the sampler's own sRGB decode and the shader's pow 2.2 then both apply, as in the
            // reference), and R16G16B16A16_UNORM (16-bit PNG, stored as float here); never the float .hdr
            t.TextureFlags = 0;
            const Image& loaded = out.images[t.DescriptorHeapIndex];
            if (gammaCorrect && (loaded.format != 0 || loaded.unorm16)) t.TextureFlags |= TB_NEEDS_GAMMA_CORRECTION_TEXTURE_FLAG;
        } else if (chk) {
            t.TextureType = TB_CHECKER_TEXTURE_TYPE;
            t.UScale = chk->uScale;
            t.VScale = chk->vScale;
            t.CheckerColor1 = cv3(chk->tex1);
            t.CheckerColor2 = cv3(chk->tex2);
        } else if (scl) {
            bool a1 = false, a2 = false;
            if (std::dynamic_pointer_cast<pbrt::ScaleTexture>(scl->tex1) ||
                std::dynamic_pointer_cast<pbrt::ScaleTexture>(scl->tex2))
                throw std::runtime_error("nested scale textures are not supported (TracerBoy.cpp:233)");
            t.TextureType = TB_SCALE_TEXTURE_TYPE;
            t.TextureIndex1 = create_texture(scl->tex1, gammaCorrect, &a1);
            t.TextureIndex2 = create_texture(scl->tex2, gammaCorrect, &a2);
            t.ScaleColor1 = cv3(scl->scale1);
            t.ScaleColor2 = cv3(scl->scale2);
            if (hasAlpha) *hasAlpha = a1 || a2;
        } else {
            throw std::runtime_error("unsupported pbrt texture type: " + tex->toString());
        }
        out.textures.push_back(t);
        return (uint32_t)out.textures.size() - 1;
    }

    // CreateMaterial, TracerBoy.cpp:273-505
    TbMaterial create_material(pbrt::Material::SP mat, pbrt::Texture::SP* alphaTex, pbrt::vec3f emissive) {
        TbMaterial m = default_material(cv3(emissive));
        bool hasAlpha = false;
        if (alphaTex) { m.alphaIndex = create_texture(*alphaTex); hasAlpha = true; }
        auto substrate = std::dynamic_pointer_cast<pbrt::SubstrateMaterial>(mat);
        auto uber = std::dynamic_pointer_cast<pbrt::UberMaterial>(mat);
        auto mix = std::dynamic_pointer_cast<pbrt::MixMaterial>(mat);
        auto mirror = std::dynamic_pointer_cast<pbrt::MirrorMaterial>(mat);
        auto metal = std::dynamic_pointer_cast<pbrt::MetalMaterial>(mat);
        auto fourier = std::dynamic_pointer_cast<pbrt::FourierMaterial>(mat);
        auto glass = std::dynamic_pointer_cast<pbrt::GlassMaterial>(mat);
        auto matte = std::dynamic_pointer_cast<pbrt::MatteMaterial>(mat);
        auto disney = std::dynamic_pointer_cast<pbrt::DisneyMaterial>(mat);
        auto plastic = std::dynamic_pointer_cast<pbrt::PlasticMaterial>(mat);
        auto sss = std::dynamic_pointer_cast<pbrt::SubSurfaceMaterial>(mat);
        auto translucent = std::dynamic_pointer_cast<pbrt::TranslucentMaterial>(mat);
        if (!mat) {
        } else if (disney) {
            m.albedo = cv3(disney->color);
            if (m.albedo.x > 0.7) m.albedo = {0.2f, 0.2f, 0.2f};
            m.roughness = disney->roughness;
            m.IOR = disney->eta;
            if (disney->metallic > 0.5) m.Flags |= TB_METALLIC_MATERIAL_FLAG;
            if (disney->specTrans > 0.001) {
                m.Flags |= TB_SUBSURFACE_SCATTER_MATERIAL_FLAG;
                m.absorption = {0, 0, 0};
                m.roughness = 0;
            }
        } else if (uber) {
            if (uber->map_kd) { bool a = false; m.albedoIndex = create_texture(uber->map_kd, true, &a); hasAlpha |= a; }
            if (uber->map_normal) m.normalMapIndex = create_texture(uber->map_normal);
            if (uber->map_emissive) m.emissiveIndex = create_texture(uber->map_emissive);
            if (uber->map_specular) m.specularMapIndex = create_texture(uber->map_specular);
            m.albedo = cv3(uber->kd);
            m.roughness = uber->uRoughness > 0.0 ? uber->uRoughness : uber->roughness;
            if (channel_average(uber->opacity) < 1.0) {
                m.Flags |= TB_SUBSURFACE_SCATTER_MATERIAL_FLAG | TB_SINGLE_SIDED_MATERIAL_FLAG;
                m.IOR = uber->index;
                m.absorption = cv3(uber->kt);
            }
        } else if (mix) {
            uint32_t i0 = add_material(mix->material0.get(), create_material(mix->material0, nullptr, emissive));
            uint32_t i1 = add_material(mix->material1.get(), create_material(mix->material1, nullptr, emissive));
            m.Flags = TB_MIX_MATERIAL_FLAG;
            m.albedo = {(float)i0, (float)i1, channel_average(mix->amount)};
        } else if (mirror) {
            m.albedo = cv3(mirror->kr);
            m.SpecularCoef = 1.0f;
            m.roughness = 0.0f;
            m.Flags |= TB_METALLIC_MATERIAL_FLAG;
        } else if (metal) {
            m.albedo = {1.0f, 1.0f, 1.0f};
            m.IOR = channel_average(metal->eta);
            m.roughness = metal->uRoughness;
            m.Flags |= TB_METALLIC_MATERIAL_FLAG;
        } else if (substrate) {
            if (substrate->map_kd) { bool a = false; m.albedoIndex = create_texture(substrate->map_kd, false, &a); hasAlpha |= a; }
            m.albedo = cv3(substrate->kd);
            m.IOR = specular_to_ior(channel_average(substrate->ks));
            m.SpecularCoef = channel_average(substrate->ks);
            m.roughness = substrate->uRoughness;
        } else if (glass) {
            m.albedo = {0, 0, 0};
            m.absorption = {0, 0, 0};
            m.IOR = glass->index;
            m.Flags |= TB_SUBSURFACE_SCATTER_MATERIAL_FLAG;
        } else if (fourier) {
            m.albedo = {0.6f, 0.6f, 0.6f};
            m.roughness = 0.2f;
        } else if (matte) {
            m.roughness = matte->sigma;
            if (matte->map_kd) { bool a = false; m.albedoIndex = create_texture(matte->map_kd, false, &a); hasAlpha |= a; }
            m.albedo = cv3(matte->kd);
            m.Flags |= TB_NO_SPECULAR_MATERIAL_FLAG;
        } else if (plastic) {
            m.roughness = plastic->roughness;
            if (plastic->map_kd) { bool a = false; m.albedoIndex = create_texture(plastic->map_kd, false, &a); hasAlpha |= a; }
            m.albedo = cv3(plastic->kd);
            m.IOR = specular_to_ior(channel_average(plastic->ks));
            m.SpecularCoef = channel_average(plastic->ks);
        } else if (sss) {
            throw std::runtime_error("pbrt 'subsurface' material is a HANDLE_FAILURE() in the reference (TracerBoy.cpp:451)");
        } else if (translucent) {
            if (translucent->map_kd) { bool a = false; m.albedoIndex = create_texture(translucent->map_kd, false, &a); hasAlpha |= a; }
            else {
                m.albedo = {0, 0, 0};
                m.absorption = {0.001f, 0.001f, 0.001f};
                m.Flags |= TB_SUBSURFACE_SCATTER_MATERIAL_FLAG;
            }
        } else {
            m.albedo = {(float)(153.0 / 255.0f), (float)(102.0f / 255.0), 58.0f / 255.0f};
            m.roughness = 0.2f;
        }
        if (!hasAlpha) m.Flags |= TB_NO_ALPHA_MATERIAL_FLAG;
        return m;
    }

    // Hair / curve shapes -> triangle tubes, TracerBoy.cpp:1426-1524 with Curves.cpp:3-52. Every quirk is kept because it
    // decides the geometry the tracer sees: three rings of three vertices per cubic segment and none at the segment's
    // end; ring faces index the *first* segment's rings whatever the segment (loopStartIndex ignores curveIndex);
    // the tangent's first term is multiplied by (1.0 - 1); the radius lerp runs from width0 to width1 *widths*, not
    // half widths; and the merge loop runs ten times whether or not the next shape is a matching curve, so a curve
    // with no mergeable successor is tessellated again on every remaining pass.
    static pbrt::vec3f quad_bezier(const pbrt::vec3f& a, const pbrt::vec3f& b, const pbrt::vec3f& c, float t) {
        return (a * (1.0f - t) + b * t) * (1.0f - t) + (b * (1.0f - t) + c * t) * t;
    }
    pbrt::TriangleMesh::SP tessellate_curves(std::vector<pbrt::Shape::SP>& shapes, size_t& si, pbrt::Curve::SP curve) {
        using pbrt::vec3f;
        auto mesh = std::make_shared<pbrt::TriangleMesh>();
        const uint32_t ringVerts = 3, ringsPerSegment = 3;
        const float radiansPerVert = 3.14 * 2.0 / (float)(ringVerts);
        for (int pass = 0; pass < 10; pass++) {
            if (pass > 0 && si + 1 < shapes.size()) {
                auto nextCurve = std::dynamic_pointer_cast<pbrt::Curve>(shapes[si + 1]);
                if (nextCurve && nextCurve->material == curve->material && nextCurve->areaLight == curve->areaLight) { curve = nextCurve; si++; }
            }
            const uint32_t vertexOffset = (uint32_t)mesh->vertex.size();
            mesh->material = curve->material;
            mesh->areaLight = curve->areaLight;
            if (curve->P.size() < 4) throw std::runtime_error("curve needs at least 4 control points (TracerBoy.cpp:1467)");
            const uint32_t segments = (uint32_t)curve->P.size() - 3;
            const float stepPerSegment = 1.0 / (float)segments;
            const float stepPerRing = stepPerSegment / (float)ringsPerSegment;
            for (uint32_t seg = 0; seg < segments; seg++) {
                const vec3f p0 = curve->P[seg], p1 = curve->P[seg + 1], p2 = curve->P[seg + 2], p3 = curve->P[seg + 3];
                for (uint32_t ring = 0; ring < ringsPerSegment; ring++) {
                    const float t = seg * stepPerSegment + ring * stepPerRing;
                    const float radius = t * curve->width1 + (1.0 - t) * curve->width0;
                    const vec3f centre = quad_bezier(p0, p1, p2, t) * (1.0f - t) + quad_bezier(p1, p2, p3, t) * t;
                    const vec3f tangent = (p1 - p0) * 3.0 * (1.0 - t) * (1.0 - 1) + (p2 - p1) * 6.0 * t * (1.0 - t) + (p3 - p2) * 3.0 * t * t;
                    const vec3f forward = pbrt::math::normalize(tangent);
                    const vec3f up = forward.y < 0.99f ? vec3f(0, 1, 0) : vec3f(0, 0, 1);
                    const vec3f n0 = pbrt::math::normalize(pbrt::math::cross(forward, up));
                    const vec3f n1 = pbrt::math::normalize(pbrt::math::cross(n0, forward));
                    for (uint32_t v = 0; v < ringVerts; v++) {
                        const float theta = v * radiansPerVert;
                        const vec3f normal = std::cos(theta) * n0 + std::sin(theta) * n1;
                        mesh->vertex.push_back(centre + normal * radius);
                        mesh->normal.push_back(normal);
                        mesh->tangents.push_back(forward);
                    }
                    if (ring == 0) continue; // the first ring has no previous ring to form faces with
                    const uint32_t ringStart = ringVerts * ring + vertexOffset, prevStart = ringVerts * (ring - 1) + vertexOffset;
                    for (uint32_t f = 0; f < ringVerts; f++) {
                        const uint32_t l = f, r = f == ringVerts - 1 ? 0 : f + 1;
                        mesh->index.push_back(pbrt::vec3i(ringStart + l, ringStart + r, prevStart + l));
                        mesh->index.push_back(pbrt::vec3i(ringStart + r, prevStart + l, prevStart + r));
                    }
                }
            }
        }
        return mesh;
    }

    void run(pbrt::Scene::SP scene) {
        if (scene->cameras.empty()) throw std::runtime_error("scene has no camera");
        // camera, TracerBoy.cpp:1243-1272
        auto& cam = scene->cameras[0];
        pbrt::vec3f pos = cam->frame * pbrt::vec3f(0.f);
        pbrt::vec3f view = pbrt::math::normalize(pbrt::math::xfmVector(cam->frame, pbrt::vec3f(0.f, 0.f, 1.f)));
        pbrt::vec3f right = pbrt::math::normalize(pbrt::math::xfmVector(cam->frame, pbrt::vec3f(1.f, 0.f, 0.f)));
        pbrt::vec3f up = pbrt::math::xfmVector(cam->frame, pbrt::vec3f(0.f, 1.f, 0.f));
        out.camera.LensHeight = (float)(2.0 * sqrtf(pbrt::math::dot(up, up)));
        up = pbrt::math::normalize(up);
        float fovAngle = (float)(cam->fov * M_PI / 180.0);
        out.camera.FocalDistance = (float)((out.camera.LensHeight / 2.0) / tan(fovAngle / 2.0));
        pos = pos + (out.camera.FocalDistance + 0.01f) * view;
        out.camera.Position = cv3(pos);
        out.camera.LookAt = cv3(pos + view);
        out.camera.Right = cv3(right);
        out.camera.Up = cv3(up);

        // shapes: every top-level shape goes into the one global BLAS (:1361-1366). The SW path renders only that BLAS
        // (TracerBoy.cpp:2862), so object instances are dropped unless the caller asks for LoadScene's other mode
        // (bInsertInstancesIntoBLAS, :1355, 1367-1376): then every instance joins the global BLAS as the FIRST shape of
        // its object with the instance transform baked into positions, normals and tangents (:1623-1624).
        auto& world = scene->world;
        const size_t topLevelShapes = world->shapes.size();
        const size_t totalSceneInstances = topLevelShapes + (insertInstancesIntoBLAS ? world->instances.size() : 0);
        for (size_t si = 0; si < totalSceneInstances; si++) {
            pbrt::Shape::SP geometry;
            pbrt::affine3f transform = pbrt::affine3f::identity();
            if (si < topLevelShapes) geometry = world->shapes[si];
            else {
                auto& instance = world->instances[si - topLevelShapes];
                // the reference indexes object->shapes[0] unconditionally; an object without shapes (instances of instances) has nothing to insert
                if (!instance || !instance->object || instance->object->shapes.empty()) continue;
                transform = instance->xfm;
                geometry = instance->object->shapes[0];
            }
            auto mesh = std::dynamic_pointer_cast<pbrt::TriangleMesh>(geometry);
            // (an instanced curve never merges: its loop index is past the top-level shapes, :1437-1438)
            if (auto curve = std::dynamic_pointer_cast<pbrt::Curve>(geometry)) mesh = tessellate_curves(world->shapes, si, curve);
            if (!mesh) continue; // spheres, disks, ...: skipped as in the reference
            pbrt::vec3f emissive(0.f);
            std::vector<uint32_t> idx(mesh->index.size() * 3);
            for (size_t i = 0; i < mesh->index.size(); i++) {
                idx[3 * i] = mesh->index[i].x; idx[3 * i + 1] = mesh->index[i].y; idx[3 * i + 2] = mesh->index[i].z;
            }
            const TbFloat3* P = (const TbFloat3*)mesh->vertex.data();
            const TbFloat3* N = mesh->normal.size() ? (const TbFloat3*)mesh->normal.data() : nullptr;
            if (mesh->areaLight) {
                auto dl = std::dynamic_pointer_cast<pbrt::DiffuseAreaLightRGB>(mesh->areaLight);
                if (!dl) throw std::runtime_error("unsupported area light type (TracerBoy.cpp:253-266)");
                emissive = dl->L;
                append_area_lights(out, P, N, idx.data(), (uint32_t)idx.size(), cv3(emissive));
            }
            uint32_t matId;
            auto it = matIndex.find(mesh->material.get());
            if (it != matIndex.end()) matId = it->second;
            else {
                auto at = mesh->textures.find("alpha");
                TbMaterial m = create_material(mesh->material, at != mesh->textures.end() ? &at->second : nullptr, emissive);
                matId = add_material(mesh->material.get(), m);
            }
            // The reference pushes every position through `vertexBufferTransform * v` and every normal and tangent
            // through normalize(xfmNormal(vertexBufferTransform, n)) with the transform at identity
            // (TracerBoy.cpp:1636-1650); an inserted instance carries its own transform. The identity products are kept
            // literally: they turn -0 into +0. (Area lights above read the UNtransformed vertices, as :1538-1540 does.)
            const pbrt::affine3f vertexBufferTransform = transform;
            size_t nv = mesh->vertex.size();
            std::vector<TbFloat3> pos(nv), nrm(nv), tan(nv);
            std::vector<TbFloat2> uv(nv);
            for (size_t v = 0; v < nv; v++) {
                pos[v] = cv3(vertexBufferTransform * mesh->vertex[v]);
                nrm[v] = N ? cv3(pbrt::math::normalize(pbrt::math::xfmNormal(vertexBufferTransform, mesh->normal[v]))) : TbFloat3{0, 1, 0};
                tan[v] = v < mesh->tangents.size() ? cv3(pbrt::math::normalize(pbrt::math::xfmNormal(vertexBufferTransform, mesh->tangents[v]))) : TbFloat3{0, 0, 1};
                uv[v] = v < mesh->texcoord.size() ? TbFloat2{mesh->texcoord[v].x, mesh->texcoord[v].y} : TbFloat2{0, 0};
            }
            if (!N) {
                // flat normals from the untransformed positions, written into the shared vertices, last face wins (:1710-1729)
                for (size_t i = 0; i < mesh->index.size(); i++) {
                    auto t = mesh->index[i];
                    pbrt::vec3f edge1 = mesh->vertex[t.z] - mesh->vertex[t.x];
                    pbrt::vec3f edge2 = mesh->vertex[t.z] - mesh->vertex[t.y];
                    pbrt::vec3f n = pbrt::math::cross(edge1, edge2);
                    if (pbrt::math::dot(n, n) <= 0.0000000001f) n = pbrt::vec3f(0, 1, 0);
                    else n = pbrt::math::normalize(pbrt::math::xfmNormal(vertexBufferTransform, n));
                    nrm[t.x] = nrm[t.y] = nrm[t.z] = cv3(n);
                }
            }
            append_geometry(out, pos.data(), nrm.data(), uv.data(), tan.data(), (uint32_t)nv, idx.data(),
                            (uint32_t)idx.size(), matId);
        }
        // lights / environment, TracerBoy.cpp:1896-1934
        for (auto& ls : world->lightSources) {
            auto inf = std::dynamic_pointer_cast<pbrt::InfiniteLightSource>(ls);
            auto dist = std::dynamic_pointer_cast<pbrt::DistantLightSource>(ls);
            if (inf) {
                // the reference takes the last four characters of the map name unconditionally (TracerBoy.cpp:1903-1904,
                // 2200): an infinite light without a map is an out_of_range exception there, an error here
                if (inf->mapName.size() < 4) throw std::runtime_error("infinite light source without a \"mapname\" is not supported (TracerBoy.cpp:1903)");
                out.envImage = (int32_t)load_image(inf->mapName, nullptr);
                auto& l = inf->transform.l;
                out.envTransform[0] = {l.vx.x, l.vx.y, l.vx.z, 0};
                out.envTransform[1] = {l.vy.x, l.vy.y, l.vy.z, 0};
                out.envTransform[2] = {l.vz.x, l.vz.y, l.vz.z, 0};
                out.envColorScale = cv3(inf->scale);
            } else if (dist) {
                TbLight l;
                memset(&l, 0, sizeof(l));
                l.LightColor = cv3(dist->L);
                l.LightType = TB_LIGHT_TYPE_DIRECTIONAL;
                l.Direction = cv3(pbrt::math::normalize(dist->to - dist->from));
                out.lights.push_back(l);
            }
        }
        out.flipTextureUVs = 1; // PBRT uses GL-style texture sampling (:1208)
    }
};

} // namespace

extern "C" __attribute__((visibility("default")))
int tb_pbrt_import_ex(const char* path, uint32_t instanceMode, void* sceneOut, char* err, size_t errCap) {
    Scene& s = *(Scene*)sceneOut;
    try {
        std::string p(path);
        std::string ext = p.size() >= 4 ? p.substr(p.size() - 4) : "";
        pbrt::Scene::SP scene;
        if (ext == "pbrt") scene = pbrt::importPBRT(p);
        else if (ext == ".pbf") scene = pbrt::Scene::loadFrom(p);
        else throw std::runtime_error("unsupported scene extension");
        size_t slash = p.find_last_of('/');
        std::string dir = slash == std::string::npos ? "" : p.substr(0, slash + 1);
        s.clear();
        Importer imp(s, dir, instanceMode == TB_INSTANCES_INSERT_INTO_BLAS);
        imp.run(scene);
        return 0;
    } catch (const std::exception& e) {
        if (err && errCap) { strncpy(err, e.what(), errCap - 1); err[errCap - 1] = 0; }
        return -3;
    }
}

extern "C" __attribute__((visibility("default")))
int tb_pbrt_import(const char* path, void* sceneOut, char* err, size_t errCap) { return tb_pbrt_import_ex(path, TB_INSTANCES_SKIP, sceneOut, err, errCap); }

// Import + write the .tbscene cache in one call (used by the build step that converts the
// bundled scenes, and by tb_load_scene when it wants to cache).
extern "C" __attribute__((visibility("default")))
int tb_pbrt_convert(const char* path, const char* outTbscene, char* err, size_t errCap) {
    Scene s;
    int rc = tb_pbrt_import(path, &s, err, errCap);
    if (rc != 0) return rc;
    std::string e;
    if (!save_tbscene(s, outTbscene, e)) {
        if (err && errCap) { strncpy(err, e.c_str(), errCap - 1); err[errCap - 1] = 0; }
        return -3;
    }
    return 0;
}
