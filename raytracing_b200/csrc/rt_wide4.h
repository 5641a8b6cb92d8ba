/*
 * rt_wide4.h — EXPERIMENTAL 4-wide traversal layout (RT_OPT_TRAVERSAL = 3), host + device.
 *
 * The reference's binary LinearBVHNode[] (bvh.cpp:223-245) is collapsed so that every interior node N takes the children
 * of its interior children (a leaf child stays as it is): 2..4 children per wide node.  A wide node is 128 bytes
 * (8 x float4): the four child boxes slot-major (slot s = floats 6s .. 6s+5 = min.xyz, max.xyz), then the four child
 * references, then a meta word (split axes of N and of its two children, valid-slot mask).  Slots 0,1 belong to N's
 * first child, slots 2,3 to its second child; a leaf child occupies the even slot of its pair.
 *
 * trace_wide4 visits the subtrees in EXACTLY the order of the reference's binary traversal (trace_bvh.cl:99-211: near
 * child by the ray sign on the split axis first, far child deferred, inclusive slab test, later equal-t hit wins,
 * back-face culling) with the same IEEE operations: all children of a wide node are slab-tested when the node is
 * visited, deferred children carry their entry distance and are re-tested against the current t_max when popped.
 * Not testing the collapsed intermediate node's own box is exact because a child's slab interval lies inside its
 * parent's under monotone rounding — unless 0 * inf produces a NaN: callers route rays with a zero or non-finite
 * direction component to the literal traversal.
 *
 * The same function body is compiled for the GPU (rt_kernels.cu) and for the CPU model that pins it against the
 * oracle's literal traversal (tools/wide4_check.py): V4 is float4 / rtbvh::F4, OPS supplies the load and min/max.
 */
#pragma once

#include <cstdint>

#include "rt_types.h"

#if defined(__CUDACC__)
#define RT_W4_HD __host__ __device__ __forceinline__
#else
#define RT_W4_HD inline
#endif

#define RT_W4_NODE_F4 8

template <bool ANY, class V4, class OPS>
RT_W4_HD uint32_t trace_wide4(const V4* w4, const V4* wtris, int root_ref, V4 root_min, V4 root_max,
                              float ox, float oy, float oz, float dx, float dy, float dz, float t_min, float t_max,
                              float& bu, float& bv, float& bt)
{
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    const uint32_t sign_bits = (ix < 0 ? 1u : 0u) | (iy < 0 ? 2u : 0u) | (iz < 0 ? 4u : 0u);
    uint32_t prim = RT_INVALID_ID;
    {   // the reference tests every node it visits, including the root
        float ax = (root_min.x - ox) * ix, ay = (root_min.y - oy) * iy, az = (root_min.z - oz) * iz;
        float bx = (root_max.x - ox) * ix, by = (root_max.y - oy) * iy, bz = (root_max.z - oz) * iz;
        float lo = OPS::fmax(OPS::fmax(OPS::fmin(ax, bx), OPS::fmin(ay, by)), OPS::fmin(az, bz));
        float hi = OPS::fmin(OPS::fmin(OPS::fmax(ax, bx), OPS::fmax(ay, by)), OPS::fmax(az, bz));
        if (!(OPS::fmin(hi, t_max) >= OPS::fmax(lo, t_min))) return prim;
    }
    int stack_ref[64];
    float stack_lo[64];
    int sp = 0;
    int cur = root_ref;
    for (;;)
    {
        while (cur >= 0)
        {
            const V4* np = w4 + (size_t)cur * RT_W4_NODE_F4;
            const V4 q0 = OPS::ld(np), q1 = OPS::ld(np + 1), q2 = OPS::ld(np + 2), q3 = OPS::ld(np + 3), q4 = OPS::ld(np + 4), q5 = OPS::ld(np + 5);
            const V4 qr = OPS::ld(np + 6), qm = OPS::ld(np + 7);
            // slot 0: min q0.xyz max (q0.w, q1.x, q1.y); slot 1: min (q1.z, q1.w, q2.x) max q2.yzw; slot 2: min q3.xyz max (q3.w, q4.x, q4.y); slot 3: ...
#define RT_W4_SLAB(mnx, mny, mnz, mxx, mxy, mxz, LO, HIT)                                                             \
            {                                                                                                         \
                float ax = ((mnx) - ox) * ix, ay = ((mny) - oy) * iy, az = ((mnz) - oz) * iz;                           \
                float bx = ((mxx) - ox) * ix, by = ((mxy) - oy) * iy, bz = ((mxz) - oz) * iz;                           \
                LO = OPS::fmax(OPS::fmax(OPS::fmax(OPS::fmin(ax, bx), OPS::fmin(ay, by)), OPS::fmin(az, bz)), t_min); \
                float hi = OPS::fmin(OPS::fmin(OPS::fmax(ax, bx), OPS::fmax(ay, by)), OPS::fmax(az, bz));             \
                HIT = OPS::fmin(hi, t_max) >= LO;                                                                     \
            }
            float lo0, lo1, lo2, lo3; bool h0, h1, h2, h3;
            RT_W4_SLAB(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, lo0, h0)
            RT_W4_SLAB(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, lo1, h1)
            RT_W4_SLAB(q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, lo2, h2)
            RT_W4_SLAB(q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, lo3, h3)
#undef RT_W4_SLAB
            const uint32_t meta = OPS::bits(qm.x);
            const uint32_t valid = meta >> 8;
            h0 = h0 && (valid & 1u); h1 = h1 && (valid & 2u); h2 = h2 && (valid & 4u); h3 = h3 && (valid & 8u);
            const int r0 = (int)OPS::bits(qr.x), r1 = (int)OPS::bits(qr.y), r2 = (int)OPS::bits(qr.z), r3 = (int)OPS::bits(qr.w);
            // reference visiting order: the pair of N's near child first, inside each pair the near grandchild first
            const uint32_t g = (sign_bits >> (meta & 3u)) & 1u;               // 1: N's second child (slots 2,3) is nearer
            const uint32_t fl = (sign_bits >> ((meta >> 2) & 3u)) & 1u;       // 1: slot 1 before slot 0
            const uint32_t fr = (sign_bits >> ((meta >> 4) & 3u)) & 1u;       // 1: slot 3 before slot 2
            // hit / lo / ref of the four slots in visiting order v0..v3
            const bool pl0 = fl ? h1 : h0, pl1 = fl ? h0 : h1, pr0 = fr ? h3 : h2, pr1 = fr ? h2 : h3;
            const float ll0 = fl ? lo1 : lo0, ll1 = fl ? lo0 : lo1, lr0 = fr ? lo3 : lo2, lr1 = fr ? lo2 : lo3;
            const int rl0 = fl ? r1 : r0, rl1 = fl ? r0 : r1, rr0 = fr ? r3 : r2, rr1 = fr ? r2 : r3;
            const bool v0h = g ? pr0 : pl0, v1h = g ? pr1 : pl1, v2h = g ? pl0 : pr0, v3h = g ? pl1 : pr1;
            const float v0l = g ? lr0 : ll0, v1l = g ? lr1 : ll1, v2l = g ? ll0 : lr0, v3l = g ? ll1 : lr1;
            const int v0r = g ? rr0 : rl0, v1r = g ? rr1 : rl1, v2r = g ? rl0 : rr0, v3r = g ? rl1 : rr1;
            // the first hit child in visiting order is entered, the later ones are deferred (pushed last-visited first)
            bool have = false; int next = 0; float next_lo = 0.0f;
            if (v3h) { next = v3r; next_lo = v3l; have = true; }
            if (v2h) { if (have) { stack_ref[sp] = next; stack_lo[sp] = next_lo; ++sp; } next = v2r; next_lo = v2l; have = true; }
            if (v1h) { if (have) { stack_ref[sp] = next; stack_lo[sp] = next_lo; ++sp; } next = v1r; next_lo = v1l; have = true; }
            if (v0h) { if (have) { stack_ref[sp] = next; stack_lo[sp] = next_lo; ++sp; } next = v0r; next_lo = v0l; have = true; }
            if (have) cur = next;
            else
            {   // pop: a deferred child is re-tested against the (possibly shrunk) t_max
                bool found = false;
                while (sp > 0) { --sp; if (t_max >= stack_lo[sp]) { cur = stack_ref[sp]; found = true; break; } }
                if (!found) return prim;
            }
        }
        // leaf: triangles [~cur ...] until the end-of-leaf flag (wtris of rt_bvh_layout.h)
        uint32_t ti = (uint32_t)(~cur);
        for (;;)
        {
            const V4* tp = wtris + (size_t)ti * 3;
            const V4 a = OPS::ld(tp), b = OPS::ld(tp + 1), c = OPS::ld(tp + 2);
            const float p1x = a.x, p1y = a.y, p1z = a.z, e1x = a.w, e1y = b.x, e1z = b.y, e2x = b.z, e2y = b.w, e2z = c.x;
            const bool last = OPS::bits(c.y) != 0u;
            const float pvx = dy * e2z - dz * e2y, pvy = dz * e2x - dx * e2z, pvz = dx * e2y - dy * e2x;      // cross(d, e2)
            const float det = e1x * pvx + e1y * pvy + e1z * pvz;
            if (!(det < 1e-8f || -det > 1e-8f))
            {
                const float inv_det = 1.0f / det;
                const float tvx = ox - p1x, tvy = oy - p1y, tvz = oz - p1z;
                const float u = (tvx * pvx + tvy * pvy + tvz * pvz) * inv_det;
                if (!(u < 0.0f || u > 1.0f))
                {
                    const float qx = tvy * e1z - tvz * e1y, qy = tvz * e1x - tvx * e1z, qz = tvx * e1y - tvy * e1x;   // cross(tvec, e1)
                    const float v = (dx * qx + dy * qy + dz * qz) * inv_det;
                    if (!(v < 0.0f || u + v > 1.0f))
                    {
                        const float t = (e2x * qx + e2y * qy + e2z * qz) * inv_det;
                        if (!(t < t_min || t > t_max))
                        {
                            bu = u; bv = v; bt = t; prim = ti; t_max = t;
                            if (ANY) return 0u;
                        }
                    }
                }
            }
            if (last) break;
            ++ti;
        }
        bool found = false;
        while (sp > 0) { --sp; if (t_max >= stack_lo[sp]) { cur = stack_ref[sp]; found = true; break; } }
        if (!found) return prim;
    }
}

#include <cstring>
#include <string>
#include <vector>

namespace rtw4
{
struct F4 { float x, y, z, w; };
inline float ibits(uint32_t v) { float f; memcpy(&f, &v, 4); return f; }

/* Host: collapse the reference nodes into wide nodes.  Leaf references are ~(first triangle index), the convention of
 * rt_bvh_layout.h's wtris.  Returns the root reference (>= 0 wide node, < 0 leaf). */
inline int build_wide4(const RtLinearBVHNode* nodes, uint64_t n_nodes, std::vector<F4>& out)
{
    out.clear();
    if ((nodes[0].num_primitives_axis >> 16) > 0) { out.assign(RT_W4_NODE_F4, F4{ 0, 0, 0, 0 }); return ~(int)nodes[0].offset; }
    // iterative: (binary interior node, wide index); children wide indices are assigned when pushed
    struct Item { uint32_t node; int wide; };
    std::vector<Item> todo;
    int n_wide = 1;
    todo.push_back({ 0u, 0 });
    out.resize(RT_W4_NODE_F4);
    while (!todo.empty())
    {
        Item it = todo.back(); todo.pop_back();
        const RtLinearBVHNode& N = nodes[it.node];
        float box[24]; int ref[4] = { 0, 0, 0, 0 }; uint32_t valid = 0, axes[3] = { N.num_primitives_axis & 0xFFFFu, 0, 0 };
        for (int k = 0; k < 24; ++k) box[k] = 0.0f;
        const uint32_t kids[2] = { it.node + 1, N.offset };
        auto set_slot = [&](int s, uint32_t child) {
            const RtLinearBVHNode& C = nodes[child];
            box[6 * s + 0] = C.bounds_min.x; box[6 * s + 1] = C.bounds_min.y; box[6 * s + 2] = C.bounds_min.z;
            box[6 * s + 3] = C.bounds_max.x; box[6 * s + 4] = C.bounds_max.y; box[6 * s + 5] = C.bounds_max.z;
            valid |= 1u << s;
            if ((C.num_primitives_axis >> 16) > 0) ref[s] = ~(int)C.offset;
            else
            {
                ref[s] = n_wide++;
                out.resize((size_t)n_wide * RT_W4_NODE_F4);
                todo.push_back({ child, ref[s] });
            }
        };
        for (int c = 0; c < 2; ++c)
        {
            const RtLinearBVHNode& K = nodes[kids[c]];
            if ((K.num_primitives_axis >> 16) > 0) set_slot(2 * c, kids[c]);
            else
            {
                axes[1 + c] = K.num_primitives_axis & 0xFFFFu;
                set_slot(2 * c, kids[c] + 1);
                set_slot(2 * c + 1, K.offset);
            }
        }
        F4* w = &out[(size_t)it.wide * RT_W4_NODE_F4];
        for (int k = 0; k < 6; ++k) w[k] = F4{ box[4 * k], box[4 * k + 1], box[4 * k + 2], box[4 * k + 3] };
        w[6] = F4{ ibits((uint32_t)ref[0]), ibits((uint32_t)ref[1]), ibits((uint32_t)ref[2]), ibits((uint32_t)ref[3]) };
        w[7] = F4{ ibits(axes[0] | (axes[1] << 2) | (axes[2] << 4) | (valid << 8)), 0.0f, 0.0f, 0.0f };
    }
    (void)n_nodes;
    return 0;
}
} // namespace rtw4
