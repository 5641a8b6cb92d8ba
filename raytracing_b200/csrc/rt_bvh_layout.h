/*
 * rt_bvh_layout.h — host-side conversion (done once in rt_upload_scene) of the reference's
 * LinearBVHNode[] / Triangle[] into the device layout the optimised traversal kernel reads.
 * Analogous to the RTTriangle[] array the reference derives at upload
 * (/root/reference/src/integrator/cl_pt_integrator.cpp:392-402).
 *
 * Reference layout (acceleration input, bvh.cpp:223-245): depth-first array of 48-byte nodes;
 * an interior node's first child is the next node, its second child is `offset`; leaves hold
 * [offset, offset + count) into the leaf-ordered triangle array.
 *
 * Device layout:
 *   wnodes: one 64-byte record (4 x float4) per INTERIOR node holding BOTH children's boxes, so
 *           one record fetch decides both children (the reference fetches a 48-byte node per
 *           visited child, 8 bytes of which are padding):
 *             [0] c0.min.xyz, c0.max.x   [1] c0.max.yz, c1.min.xy   [2] c1.min.z, c1.max.xyz
 *             [3] (ref0, ref1, split axis, 0) as int bits;  ref >= 0: interior record index,
 *                 ref < 0: leaf, first triangle = ~ref
 *   wtris:  48 bytes (3 x float4) per triangle in the same (leaf) order, i.e. index == primitive_id:
 *             p1.xyz, e1.x | e1.yz, e2.xy | e2.z, end_of_leaf flag, 0, 0     (e1 = p2-p1, e2 = p3-p1,
 *           the two subtractions trace_bvh.cl:30-31 performs per test, done once here in the same
 *           IEEE arithmetic)
 * The visiting order of the traversal is unchanged (near child by ray sign of the split axis,
 * far child deferred), so hit results are identical to the reference order; see trace_fast().
 */
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "rt_types.h"

namespace rtbvh
{

struct F4 { float x, y, z, w; };

struct WideLayout
{
    std::vector<F4> nodes;   // 4 per interior node
    std::vector<F4> tris;    // 3 per triangle
    int root_ref = 0;
    int max_depth = 0;
};

inline float int_bits(int v) { float f; memcpy(&f, &v, 4); return f; }

inline bool build_layout(const RtLinearBVHNode* nodes, uint64_t n_nodes, const RtTriangle* tris, uint64_t n_tris,
                         WideLayout& out, std::string& err)
{
    if (n_nodes == 0 || n_tris == 0) { err = "empty BVH"; return false; }
    std::vector<int> wide_index(n_nodes, -1);
    std::vector<uint8_t> seen(n_nodes, 0), covered(n_tris, 0), last_flag(n_tris, 0);
    // iterative DFS from the root: validates indices, acyclicity, leaf coverage, depth
    struct Item { uint32_t node; int depth; };
    std::vector<Item> stack;
    stack.push_back({ 0u, 1 });
    int n_interior = 0, max_depth = 0;
    std::vector<uint32_t> interior_order;
    while (!stack.empty())
    {
        Item it = stack.back(); stack.pop_back();
        if (it.node >= n_nodes) { err = "child index out of range"; return false; }
        if (seen[it.node]) { err = "node reachable twice (not a tree)"; return false; }
        seen[it.node] = 1;
        if (it.depth > max_depth) max_depth = it.depth;
        const RtLinearBVHNode& n = nodes[it.node];
        uint32_t count = n.num_primitives_axis >> 16;
        if (count > 0)
        {
            if ((uint64_t)n.offset + count > n_tris) { err = "leaf range outside the triangle array"; return false; }
            for (uint32_t i = 0; i < count; ++i)
            {
                if (covered[n.offset + i]) { err = "triangle referenced by two leaves"; return false; }
                covered[n.offset + i] = 1;
            }
            last_flag[n.offset + count - 1] = 1;
        }
        else
        {
            if ((n.num_primitives_axis & 0xFFFFu) > 2) { err = "split axis out of range"; return false; }
            wide_index[it.node] = n_interior++;
            interior_order.push_back(it.node);
            // push second child first so the first child (node+1) is numbered next: DFS order is kept
            stack.push_back({ n.offset, it.depth + 1 });
            stack.push_back({ it.node + 1, it.depth + 1 });
        }
    }
    // the reference kernel's private stack holds 64 entries (trace_bvh.cl:142)
    if (max_depth > 64) { err = "BVH deeper than the 64-entry traversal stack of trace_bvh.cl:142"; return false; }
    out.max_depth = max_depth;

    auto ref_of = [&](uint32_t child) -> int {
        const RtLinearBVHNode& c = nodes[child];
        if ((c.num_primitives_axis >> 16) > 0) return ~(int)c.offset;
        return wide_index[child];
    };
    out.nodes.assign((size_t)(n_interior ? n_interior : 1) * 4, F4{ 0, 0, 0, 0 });
    for (uint32_t node : interior_order)
    {
        const RtLinearBVHNode& n = nodes[node];
        const RtLinearBVHNode& c0 = nodes[node + 1];
        const RtLinearBVHNode& c1 = nodes[n.offset];
        F4* w = &out.nodes[(size_t)wide_index[node] * 4];
        w[0] = F4{ c0.bounds_min.x, c0.bounds_min.y, c0.bounds_min.z, c0.bounds_max.x };
        w[1] = F4{ c0.bounds_max.y, c0.bounds_max.z, c1.bounds_min.x, c1.bounds_min.y };
        w[2] = F4{ c1.bounds_min.z, c1.bounds_max.x, c1.bounds_max.y, c1.bounds_max.z };
        w[3] = F4{ int_bits(ref_of(node + 1)), int_bits(ref_of(n.offset)), int_bits((int)(n.num_primitives_axis & 0xFFFFu)), 0.0f };
    }
    const RtLinearBVHNode& root = nodes[0];
    out.root_ref = ((root.num_primitives_axis >> 16) > 0) ? ~(int)root.offset : 0;

    out.tris.resize((size_t)n_tris * 3);
    for (uint64_t i = 0; i < n_tris; ++i)
    {
        const RtFloat3 &p1 = tris[i].v1.position, &p2 = tris[i].v2.position, &p3 = tris[i].v3.position;
        // volatile-free plain float subtractions; x86-64 (no FMA in the baseline ISA) rounds each once
        float e1x = p2.x - p1.x, e1y = p2.y - p1.y, e1z = p2.z - p1.z;
        float e2x = p3.x - p1.x, e2y = p3.y - p1.y, e2z = p3.z - p1.z;
        // triangles not referenced by any reachable leaf can never be visited; flag them as leaf ends too
        uint32_t flag = (last_flag[i] || !covered[i]) ? 1u : 0u;
        out.tris[i * 3 + 0] = F4{ p1.x, p1.y, p1.z, e1x };
        out.tris[i * 3 + 1] = F4{ e1y, e1z, e2x, e2y };
        float fl; memcpy(&fl, &flag, 4);
        out.tris[i * 3 + 2] = F4{ e2z, fl, 0.0f, 0.0f };
    }
    return true;
}

} // namespace rtbvh
