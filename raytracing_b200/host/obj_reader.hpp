/*
 * obj_reader.hpp — Wavefront OBJ / MTL reader of the scene loader.
 *
 * The reference reads scenes with the tinyobjloader v2.0 it vendors (3rdparty/tinyobjloader, called from scene.cpp:138 with
 * triangulation on).  What reaches the renderer depends on details that the OBJ format leaves open and that loader decides:
 * how a decimal string becomes a float, how quads and polygons are cut into triangles, which faces are dropped, how material
 * names and texture options are tokenised.  This reader follows those decisions (each one is named where it is made in
 * obj_reader.cpp) so that the triangle and material arrays are the reference's, bit for bit; tests/test_host.py compares it
 * with the reference's own loader on the inputs of tests/obj_cases.py (committed expectations, and live where the reference exists).
 */
#pragma once

#include <string>
#include <vector>

namespace rt_host
{
namespace obj
{

struct Material
{
    std::string name;
    float diffuse[3] = { 0, 0, 0 }, specular[3] = { 0, 0, 0 }, transmittance[3] = { 0, 0, 0 }, emission[3] = { 0, 0, 0 };
    float ior = 1.0f, roughness = 0.0f, metallic = 0.0f;
    std::string diffuse_tex, specular_tex, roughness_tex, metallic_tex, emissive_tex, alpha_tex;
};

struct Index { int v = -1, vt = -1, vn = -1; };          // zero-based; -1 = absent

struct Mesh
{
    std::vector<float> positions, normals, texcoords;    // 3, 3, 2 floats per element
    std::vector<Index> indices;                          // 3 per triangle, in file order
    std::vector<int> material_ids;                       // per triangle, -1 = none
    std::vector<Material> materials;
};

// false + `error` when the file cannot be opened or a face has an index the format does not allow (0, or no number)
bool Read(const char* filename, const std::string& mtl_dir, Mesh& out, std::string& error);

// one decimal number in the reader's arithmetic (exposed for the tests): false when [s, end) is not a number
bool ParseDouble(const char* s, const char* end, double* result);

} // namespace obj
} // namespace rt_host
