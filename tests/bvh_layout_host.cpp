/* tests/bvh_layout_host.cpp — TEST INFRASTRUCTURE: raytracing_b200/csrc/rt_bvh_layout.h (the host-side record layouts that
 * rt_upload_scene builds) behind a C entry point for tests/test_bvh_layout.py. */
#include <cstring>
#include <string>
#include "rt_bvh_layout.h"

extern "C" int bvh_layout_build(const RtLinearBVHNode* nodes, uint64_t n_nodes, const RtTriangle* tris, uint64_t n_tris,
                                float* wnodes_out /* 16 floats per record, n_nodes records of room */, float* wtris_out /* 12 floats per triangle */,
                                int* root_ref, int* max_depth, uint32_t* top_n, uint64_t* n_records, char* err, int err_len)
{
    rtbvh::WideLayout wl; std::string e;
    if (!rtbvh::build_layout(nodes, n_nodes, tris, n_tris, wl, e)) { strncpy(err, e.c_str(), err_len - 1); err[err_len - 1] = 0; return 1; }
    memcpy(wnodes_out, wl.nodes.data(), wl.nodes.size() * 16);
    memcpy(wtris_out, wl.tris.data(), wl.tris.size() * 16);
    *root_ref = wl.root_ref; *max_depth = wl.max_depth; *top_n = wl.top_n; *n_records = wl.nodes.size() / 4;
    return 0;
}

extern "C" void shade_records_build(const RtTriangle* tris, uint64_t n_tris, const RtPackedMaterial* mats, uint64_t n_mats,
                                    const RtLight* lights, uint64_t n_lights, float* tri_out /* 28 per triangle */,
                                    float* mat_out /* 16 per material */, float* light_out /* 8 per light */)
{
    std::vector<rtbvh::F4> rec;
    rtbvh::build_tri_shade(tris, n_tris, rec); memcpy(tri_out, rec.data(), rec.size() * 16);
    rtbvh::build_mat_rec(mats, n_mats, rec); memcpy(mat_out, rec.data(), rec.size() * 16);
    rtbvh::build_light_rec(lights, n_lights, rec); memcpy(light_out, rec.data(), rec.size() * 16);
}
