/*
 * reference_api.cpp — implementation of the mirrored reference-facing classes declared in reference_api.hpp:
 *   Scene (OBJ/MTL loader, material packing, lights, emissive list, Radiance .hdr reader)
 *   Integrator (the per-frame wavefront schedule and the base-class setters)
 */
#include "reference_api.hpp"
#include "obj_reader.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace rt_host
{

namespace
{

float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// scene.cpp:53-61
std::uint32_t PackAlbedo(float r, float g, float b, std::uint32_t tex)
{
    r = clampf(r, 0.0f, 1.0f); g = clampf(g, 0.0f, 1.0f); b = clampf(b, 0.0f, 1.0f);
    return ((std::uint32_t)(r * 255.0f)) | ((std::uint32_t)(g * 255.0f) << 8) | ((std::uint32_t)(b * 255.0f) << 16) | (tex << 24);
}

// scene.cpp:63-85
std::uint32_t PackRGBE(float r, float g, float b)
{
    r = std::fmax(r, 0.0f); g = std::fmax(g, 0.0f); b = std::fmax(b, 0.0f);
    float v = r;
    if (g > v) v = g;
    if (b > v) v = b;
    if (v < 1e-32f) return 0;
    int e;
    v = (float)(std::frexp(v, &e) * 256.0f / v);
    return ((std::uint32_t)(r * v)) | ((std::uint32_t)(g * v) << 8) | ((std::uint32_t)(b * v) << 16) | ((std::uint32_t)(e + 128) << 24);
}

// scene.cpp:87-103 (host-side unpack used only to find emissive triangles)
float EmissionSum(std::uint32_t rgbe)
{
    int e = (int)(rgbe >> 24);
    if (!e) return 0.0f;
    float f = std::ldexp(1.0f, e - (128 + 8));
    float r = (float)((rgbe >> 0) & 0xFF) * f, g = (float)((rgbe >> 8) & 0xFF) * f, b = (float)((rgbe >> 16) & 0xFF) * f;
    return r + g + b;
}

// scene.cpp:105-124
std::uint32_t PackRoughnessMetalness(float roughness, std::uint32_t ri, float metalness, std::uint32_t mi)
{
    roughness = clampf(roughness, 0.0f, 1.0f); metalness = clampf(metalness, 0.0f, 1.0f);
    return ((std::uint32_t)(roughness * 255.0f)) | (ri << 8) | ((std::uint32_t)(metalness * 255.0f) << 16) | (mi << 24);
}
std::uint32_t PackIorEmissionIdxTransparency(float ior, std::uint32_t ei, float transparency, std::uint32_t ti)
{
    ior = clampf(ior, 0.0f, 10.0f); transparency = clampf(transparency, 0.0f, 1.0f);
    return ((std::uint32_t)(ior * 25.5f)) | (ei << 8) | ((std::uint32_t)(transparency * 255.0f) << 16) | (ti << 24);
}

std::string dirname_of(const std::string& p)
{
    size_t s = p.find_last_of("/\\");
    return s == std::string::npos ? std::string() : p.substr(0, s);
}

} // namespace

Scene::Scene(const char* filename, float scale, bool flip_yz) { Load(filename, scale, flip_yz); }

Scene::Scene(std::vector<Triangle> triangles, std::vector<PackedMaterial> materials, std::vector<Light> lights,
             std::vector<Texture> textures, std::vector<std::uint32_t> texture_data)
    : triangles_(std::move(triangles)), materials_(std::move(materials)), lights_(std::move(lights)),
      textures_(std::move(textures)), texture_data_(std::move(texture_data))
{
    if (triangles_.empty() || materials_.empty()) throw std::runtime_error("Scene: triangles and materials are required");
    for (Triangle const& t : triangles_)
        if (t.mtlIndex >= materials_.size()) throw std::runtime_error("Scene: triangle references a missing material");
}

// scene.cpp:127-274: the OBJ reader's output (obj_reader.cpp stands in for tinyobjloader) -> packed materials + triangles
void Scene::Load(const char* filename, float scale, bool flip_yz)
{
    const std::string folder = dirname_of(filename);
    obj::Mesh mesh;
    std::string why;
    if (!obj::Read(filename, folder, mesh, why)) throw std::runtime_error("Failed to load the scene! " + why);

    const float kGamma = 2.2f;
    const std::uint32_t kInvalidTextureIndex = 0xFF;
    materials_.resize(mesh.materials.size());
    for (size_t i = 0; i < mesh.materials.size(); ++i)
    {
        const obj::Material& m = mesh.materials[i];
        // scene.cpp:155-186: the order in which the textures are loaded fixes their indices.  Diffuse and specular are separate
        // statements; the roughness / metallic pair and the emissive / alpha pair are each two arguments of ONE call
        // (scene.cpp:172-184), and both g++ and MSVC evaluate the later argument first: metallic before roughness, alpha before
        // emissive
        auto tex = [&](const std::string& name) -> std::uint32_t {
            return name.empty() ? kInvalidTextureIndex : (std::uint32_t)LoadTexture(folder + "/" + name);
        };
        PackedMaterial& o = materials_[i];
        o.diffuse_albedo = PackAlbedo(std::pow(m.diffuse[0], kGamma), std::pow(m.diffuse[1], kGamma), std::pow(m.diffuse[2], kGamma), tex(m.diffuse_tex));
        o.specular_albedo = PackAlbedo(std::pow(m.specular[0], kGamma), std::pow(m.specular[1], kGamma), std::pow(m.specular[2], kGamma), tex(m.specular_tex));
        o.emission = PackRGBE(m.emission[0], m.emission[1], m.emission[2]);
        const std::uint32_t mt = tex(m.metallic_tex), rt = tex(m.roughness_tex);
        o.roughness_metalness = PackRoughnessMetalness(m.roughness, rt, m.metallic, mt);
        const std::uint32_t at = tex(m.alpha_tex), et = tex(m.emissive_tex);
        o.ior_emission_idx_transparency = PackIorEmissionIdxTransparency(m.ior, et, m.transmittance[0], at);
    }

    auto flip = [flip_yz](float3& v) { if (flip_yz) { float t = v.y; v.y = v.z; v.z = t; v.y = -v.y; } };
    const std::vector<float>& positions = mesh.positions;
    const std::vector<float>& normals = mesh.normals;
    const std::vector<float>& texcoords = mesh.texcoords;
    const size_t n_faces = mesh.indices.size() / 3;
    triangles_.reserve(n_faces);
    for (size_t face = 0; face < n_faces; ++face)
    {
        Triangle t;
        memset(&t, 0, sizeof(t));
        Vertex* vs[3] = { &t.v1, &t.v2, &t.v3 };
        for (int k = 0; k < 3; ++k)
        {
            const obj::Index& c = mesh.indices[face * 3 + k];
            // (the reference indexes its arrays unchecked here, scene.cpp:203-241: a missing vertex or normal is undefined behaviour
            // there and an error here)
            if (c.v < 0 || (size_t)c.v * 3 + 2 >= positions.size()) throw std::runtime_error("OBJ face references a missing vertex");
            if (c.vn < 0 || (size_t)c.vn * 3 + 2 >= normals.size()) throw std::runtime_error("OBJ face has no normal (normals are required, scene.cpp:222-224)");
            if (c.vt >= 0 && (size_t)c.vt * 2 + 1 >= texcoords.size()) throw std::runtime_error("OBJ face references a missing texture coordinate");
            Vertex& v = *vs[k];
            v.position = make_float3(positions[c.v * 3 + 0] * scale, positions[c.v * 3 + 1] * scale, positions[c.v * 3 + 2] * scale);
            v.normal = make_float3(normals[c.vn * 3 + 0], normals[c.vn * 3 + 1], normals[c.vn * 3 + 2]);
            v.texcoord = make_float3(c.vt < 0 ? 0.0f : texcoords[c.vt * 2 + 0], c.vt < 0 ? 0.0f : texcoords[c.vt * 2 + 1], 0.0f);
            flip(v.position); flip(v.normal);
        }
        const int material = mesh.material_ids[face];
        t.mtlIndex = (material >= 0 && (size_t)material < materials_.size()) ? (std::uint32_t)material : 0u;
        triangles_.push_back(t);
    }
}

void Scene::CollectEmissiveTriangles()
{
    emissive_indices_.clear();
    for (std::uint32_t i = 0; i < triangles_.size(); ++i)
        if (EmissionSum(materials_[triangles_[i].mtlIndex].emission) > 0.0f) emissive_indices_.push_back(i);
    scene_info_.emissive_count = (std::uint32_t)emissive_indices_.size();
}

// scene.cpp:276-323
std::size_t Scene::LoadTexture(const std::string& filename)
{
    auto it = loaded_textures_.find(filename);
    if (it != loaded_textures_.end()) return it->second;
    TextureImage image; std::string why;
    if (!LoadTextureImage(filename.c_str(), image, why)) throw std::runtime_error("Failed to load file " + filename + ": " + why);
    if (textures_.size() >= 0xFF) throw std::runtime_error("more than 255 textures: the packed material format holds 8-bit texture indices (0xFF = none)");
    Texture t; memset(&t, 0, sizeof(t));
    t.width = (int)image.width; t.height = (int)image.height; t.data_start = (int)texture_data_.size();
    std::size_t idx = textures_.size();
    textures_.push_back(t);
    texture_data_.insert(texture_data_.end(), image.data.begin(), image.data.end());
    loaded_textures_.emplace(filename, idx);
    return idx;
}

void Scene::AddPointLight(float3 origin, float3 radiance)
{
    Light l; memset(&l, 0, sizeof(l));
    l.origin = origin; l.radiance = radiance; l.type = RT_LIGHT_TYPE_POINT;
    lights_.push_back(l);
}

void Scene::AddDirectionalLight(float3 d, float3 radiance)
{
    // float3::Normalize (mathlib.hpp:47-48): each component divided by the length
    float len = std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
    Light l; memset(&l, 0, sizeof(l));
    l.origin = make_float3(d.x / len, d.y / len, d.z / len); l.radiance = radiance; l.type = RT_LIGHT_TYPE_DIRECTIONAL;
    lights_.push_back(l);
}

void Scene::Finalize(const char* env_map_path)
{
    CollectEmissiveTriangles();
    scene_info_.analytic_light_count = (std::uint32_t)lights_.size();
    if (!LoadHDR(env_map_path, env_image_)) throw std::runtime_error(std::string("Failed to load environment map ") + env_map_path);
}

void Scene::Finalize(const float* env_rgba, std::uint32_t env_width, std::uint32_t env_height)
{
    CollectEmissiveTriangles();
    scene_info_.analytic_light_count = (std::uint32_t)lights_.size();
    env_image_.width = env_width; env_image_.height = env_height;
    env_image_.data.assign(env_rgba, env_rgba + (size_t)env_width * env_height * 4);
}

// ------------------------------------------------------------------------------------------------ Radiance HDR
namespace
{
struct Rgbe { unsigned char c[4]; };

bool read_flat(std::vector<Rgbe>& line, size_t from, FILE* f)
{   // old-style scanline with run markers (1,1,1,count)
    int rshift = 0;
    size_t i = from;
    while (i < line.size())
    {
        int r = fgetc(f), g = fgetc(f), b = fgetc(f), e = fgetc(f);
        if (e == EOF) return false;
        if (r == 1 && g == 1 && b == 1)
        {
            if (i == 0) return false;             // a run marker needs a previous pixel to repeat
            // the reference keeps the repeat count in an unsigned char (hdr_loader.cpp:192): a second marker in a row, whose count
            // the format shifts left by 8, repeats nothing there — and so nothing here
            for (int n = (unsigned char)(e << rshift); n > 0 && i < line.size(); --n) { line[i] = line[i - 1]; ++i; }
            rshift += 8;
        }
        else { line[i].c[0] = (unsigned char)r; line[i].c[1] = (unsigned char)g; line[i].c[2] = (unsigned char)b; line[i].c[3] = (unsigned char)e; ++i; rshift = 0; }
    }
    return true;
}

bool read_scanline(std::vector<Rgbe>& line, FILE* f)
{
    size_t len = line.size();
    if (len < 8 || len > 32767) return read_flat(line, 0, f);
    int c0 = fgetc(f);
    if (c0 != 2) { ungetc(c0, f); return read_flat(line, 0, f); }
    int g = fgetc(f), b = fgetc(f), e = fgetc(f);
    if (g != 2 || (b & 128))
    {
        line[0].c[0] = 2; line[0].c[1] = (unsigned char)g; line[0].c[2] = (unsigned char)b; line[0].c[3] = (unsigned char)e;
        return read_flat(line, 1, f);
    }
    for (int comp = 0; comp < 4; ++comp)
        for (size_t j = 0; j < len;)
        {
            int code = fgetc(f);
            if (code == EOF) return false;
            if (code > 128) { int n = code & 127, val = fgetc(f); while (n-- && j < len) line[j++].c[comp] = (unsigned char)val; }
            else { int n = code; while (n-- && j < len) line[j++].c[comp] = (unsigned char)fgetc(f); }
        }
    return !feof(f);
}
} // namespace

bool LoadHDR(const char* filename, Image& res)
{
    FILE* f = fopen(filename, "rb");
    if (!f) return false;
    char magic[10];
    if (fread(magic, 10, 1, f) != 1 || memcmp(magic, "#?RADIANCE", 10) != 0) { fclose(f); return false; }
    // header lines up to the empty line, then the resolution line "-Y h +X w"
    int prev = 0, c = 0;
    fgetc(f);
    for (;;) { prev = c; c = fgetc(f); if (c == EOF) { fclose(f); return false; } if (c == '\n' && prev == '\n') break; }
    char reso[200]; int n = 0;
    for (;;) { c = fgetc(f); if (c == EOF || n >= 199) { fclose(f); return false; } reso[n++] = (char)c; if (c == '\n') break; }
    reso[n] = 0;
    long w = 0, h = 0;
    if (sscanf(reso, "-Y %ld +X %ld", &h, &w) != 2 || w <= 0 || h <= 0) { fclose(f); return false; }
    res.width = (std::uint32_t)w; res.height = (std::uint32_t)h;
    res.data.assign((size_t)w * h * 4, 0.0f);
    std::vector<Rgbe> line((size_t)w);
    float* out = res.data.data();
    for (long y = 0; y < h; ++y)
    {
        if (!read_scanline(line, f)) break;
        for (long x = 0; x < w; ++x)
        {   // hdr_loader.cpp:102-120: v = val / 256, d = 2^(E - 128)
            float d = std::pow(2.0f, (float)((int)line[x].c[3] - 128));
            out[0] = (line[x].c[0] / 256.0f) * d; out[1] = (line[x].c[1] / 256.0f) * d; out[2] = (line[x].c[2] / 256.0f) * d;
            out += 4;
        }
    }
    fclose(f);
    return true;
}

// ================================================================================================ Integrator (base class)
// The per-frame wavefront schedule, integrator.cpp:27-59: the bounce loop is INCLUSIVE of max_bounces.
void Integrator::Integrate()
{
    if (request_reset_ || enable_denoiser_)
    {
        Reset();
        request_reset_ = false;
    }
    GenerateRays();
    for (std::uint32_t bounce = 0; bounce <= max_bounces_; ++bounce)
    {
        IntersectRays(bounce);
        if (bounce == 0) ComputeAOVs();
        ShadeMissedRays(bounce);
        ClearOutgoingRayCounter(bounce);
        ClearShadowRayCounter();
        ShadeSurfaceHits(bounce);
        IntersectShadowRays();
        AccumulateDirectSamples();
    }
    AdvanceSampleCount();
    if (enable_denoiser_)
    {
        Denoise();
        CopyHistoryBuffers();
    }
    ResolveRadiance();
}

void Integrator::SetMaxBounces(std::uint32_t max_bounces)
{
    max_bounces_ = max_bounces;
    RequestReset();
}

void Integrator::EnableWhiteFurnace(bool enable)
{
    if (enable == enable_white_furnace_) return;
    enable_white_furnace_ = enable;
    CreateKernels();
    RequestReset();
}

} // namespace rt_host
