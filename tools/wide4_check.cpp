/* tools/wide4_check.cpp — ANALYSIS ONLY: runs raytracing_b200/csrc/rt_wide4.h (the function body the GPU kernels of
 * RT_OPT_TRAVERSAL = 3 compile) on the CPU over the layouts rt_upload_scene would build, for tools/wide4_check.py. */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "rt_math.h"
#include "rt_types.h"
#include "../raytracing_b200/csrc/rt_bvh_layout.h"
#include "../raytracing_b200/csrc/rt_wide4.h"

struct HostOps
{
    static inline rtw4::F4 ld(const rtw4::F4* p) { return *p; }
    static inline float fmin(float a, float b) { return rt_fminf(a, b); }
    static inline float fmax(float a, float b) { return rt_fmaxf(a, b); }
    static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
};

extern "C" int wide4_trace(const RtLinearBVHNode* nodes, uint32_t n_nodes, const RtTriangle* tris, uint32_t n_tris, const RtRay* rays, uint32_t n_rays,
                           int any_hit, RtHit* hits, uint8_t* status, uint64_t* n_wide_out)
{
    rtbvh::WideLayout wl; std::string err;
    if (!rtbvh::build_layout(nodes, n_nodes, tris, n_tris, wl, err)) return -1;
    std::vector<rtw4::F4> w4;
    int root = rtw4::build_wide4(nodes, n_nodes, w4);
    *n_wide_out = w4.size() / RT_W4_NODE_F4;
    const rtw4::F4* wt = (const rtw4::F4*)wl.tris.data();
    rtw4::F4 rmin = { nodes[0].bounds_min.x, nodes[0].bounds_min.y, nodes[0].bounds_min.z, 0 }, rmax = { nodes[0].bounds_max.x, nodes[0].bounds_max.y, nodes[0].bounds_max.z, 0 };
#pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n_rays; ++i)
    {
        const RtRay& r = rays[i];
        RtHit h; h.bc.x = h.bc.y = 0.0f; h.t = 0.0f; h.primitive_id = RT_INVALID_ID;
        float s = ((r.origin.x + r.origin.y) + r.origin.z) + ((r.direction.x + r.direction.y) + r.direction.z);
        bool ok = std::fabs(s) <= 3.0e38f && r.direction.x != 0.0f && r.direction.y != 0.0f && r.direction.z != 0.0f;
        status[i] = ok ? 0 : 1;
        if (ok)
        {
            float bu = 0, bv = 0, bt = 0;
            uint32_t p = any_hit ? trace_wide4<true, rtw4::F4, HostOps>(w4.data(), wt, root, rmin, rmax, r.origin.x, r.origin.y, r.origin.z, r.direction.x, r.direction.y,
                                                                       r.direction.z, r.origin.w, r.direction.w, bu, bv, bt)
                                 : trace_wide4<false, rtw4::F4, HostOps>(w4.data(), wt, root, rmin, rmax, r.origin.x, r.origin.y, r.origin.z, r.direction.x, r.direction.y,
                                                                        r.direction.z, r.origin.w, r.direction.w, bu, bv, bt);
            h.primitive_id = p; h.bc.x = bu; h.bc.y = bv; h.t = bt;
        }
        hits[i] = h;
    }
    return 0;
}
