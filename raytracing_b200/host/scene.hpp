/*
 * Mirror of the reference's Scene (/root/reference/src/scene/scene.hpp:34-67, scene.cpp:46-361):
 * Wavefront OBJ + MTL  ->  Triangle[], PackedMaterial[], Light[], Texture[], texel words, emissive
 * index list, SceneInfo, environment image.  Host-only producer of the arrays the integrator uploads.
 * The OBJ/MTL reader is a small purpose-built parser (triangulated faces with v//vn or v/vt/vn
 * indices, the MTL keys the reference consumes: Kd Ks Ke Ni Tf Pr Pm map_*), with tinyobjloader's
 * defaults for absent keys (tiny_obj_loader.h:1331-1340).  Image textures (map_*) need an image
 * decoder and are not supported here (none of the shipped scenes has one): loading fails loudly.
 */
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "types.hpp"

namespace rt_host
{

struct Image
{
    std::uint32_t width = 0, height = 0;
    std::vector<float> data;      // RGBA32F
};

// Radiance .hdr reader with the reference's conversion (loaders/hdr_loader.cpp:29-120):
// rows in file order, value = (mantissa / 256) * 2^(e - 128), alpha left 0.
bool LoadHDR(const char* filename, Image& result);

class Scene
{
public:
    Scene(const char* filename, float scale, bool flip_yz);

    std::vector<Triangle>& GetTriangles() { return triangles_; }
    std::vector<Triangle> const& GetTriangles() const { return triangles_; }
    std::vector<std::uint32_t> const& GetEmissiveIndices() const { return emissive_indices_; }
    std::vector<PackedMaterial> const& GetMaterials() const { return materials_; }
    std::vector<Texture> const& GetTextures() const { return textures_; }
    std::vector<std::uint32_t> const& GetTextureData() const { return texture_data_; }
    std::vector<Light> const& GetLights() const { return lights_; }
    SceneInfo const& GetSceneInfo() const { return scene_info_; }
    Image const& GetEnvImage() const { return env_image_; }
    // The reference hard-codes "assets/ibl/CGSkies_0036_free.hdr" relative to the CWD (scene.cpp:360);
    // the path is a parameter here, with that default.
    void Finalize(const char* env_map_path = "assets/ibl/CGSkies_0036_free.hdr");
    // Headless variant for callers that already hold the decoded environment image.
    void Finalize(const float* env_rgba, std::uint32_t env_width, std::uint32_t env_height);
    void AddPointLight(float3 origin, float3 radiance);
    void AddDirectionalLight(float3 direction, float3 radiance);

private:
    void Load(const char* filename, float scale, bool flip_yz);
    void CollectEmissiveTriangles();

    std::vector<Triangle> triangles_;
    std::vector<std::uint32_t> emissive_indices_;
    std::vector<PackedMaterial> materials_;
    std::vector<Light> lights_;
    std::vector<Texture> textures_;
    std::vector<std::uint32_t> texture_data_;
    SceneInfo scene_info_ = {};
    Image env_image_;
};

} // namespace rt_host
