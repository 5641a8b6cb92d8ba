#include "bvh.hpp"

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cstring>
#include <stdexcept>

namespace rt_host
{

namespace
{

struct Box
{
    float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX);
    float3 hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    void grow(const Box& b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); }
    void grow(const float3& p) { lo = vmin(lo, p); hi = vmax(hi, p); }
    // 2 * (dx*dy + dx*dz + dy*dz), Bounds3::SurfaceArea (mathlib.hpp:176)
    float area() const { float3 d = hi - lo; return 2 * (d.x * d.y + d.x * d.z + d.y * d.z); }
    unsigned widest() const
    {   // Bounds3::MaximumExtent (mathlib.hpp:179-188)
        float3 d = hi - lo;
        if (d.x > d.y && d.x > d.z) return 0;
        return d.y > d.z ? 1 : 2;
    }
};

struct Prim { unsigned index; Box box; float3 centroid; };

// build record (arena): children by index, -1 for leaves
struct Rec { Box box; int child[2]; int axis, first, count; };

struct Builder
{
    const std::vector<Triangle>& tris;
    std::vector<Prim> prims;
    std::vector<Rec> recs;
    std::vector<Triangle> ordered;

    explicit Builder(const std::vector<Triangle>& t) : tris(t) {}

    int leaf(int rec, unsigned start, unsigned end, const Box& box)
    {
        recs[rec].first = (int)ordered.size();
        recs[rec].count = (int)(end - start);
        recs[rec].box = box;
        recs[rec].child[0] = recs[rec].child[1] = -1;
        for (unsigned i = start; i < end; ++i) ordered.push_back(tris[prims[i].index]);
        return rec;
    }

    // bucket of a centroid along `dim` inside the centroid bounds cb (bvh.cpp:141-146, Bounds3::Offset mathlib.hpp:190-196)
    static int bucket_of(const Box& cb, const float3& c, unsigned dim)
    {
        float lo = component(cb.lo, dim), hi = component(cb.hi, dim);
        float o = component(c, dim) - lo;
        if (hi > lo) o /= hi - lo;
        int b = (int)(12u * o);
        return b == 12 ? 11 : b;
    }

    int build(unsigned start, unsigned end)
    {
        int rec = (int)recs.size();
        recs.push_back(Rec());
        Box box;
        for (unsigned i = start; i < end; ++i) box.grow(prims[i].box);
        unsigned n = end - start;
        if (n == 1) return leaf(rec, start, end, box);

        Box cb;
        for (unsigned i = start; i < end; ++i) cb.grow(prims[i].centroid);
        unsigned dim = cb.widest();
        if (component(cb.hi, dim) == component(cb.lo, dim)) return leaf(rec, start, end, box);   // all centroids coincide

        unsigned mid = (start + end) / 2;
        if (n <= 2)
        {
            std::nth_element(&prims[start], &prims[mid], &prims[end - 1] + 1,
                             [dim](const Prim& a, const Prim& b) { return component(a.centroid, dim) < component(b.centroid, dim); });
        }
        else
        {
            const int kBuckets = 12;
            int count[kBuckets] = {};
            Box bbox[kBuckets];
            for (unsigned i = start; i < end; ++i)
            {
                int b = bucket_of(cb, prims[i].centroid, dim);
                ++count[b];
                bbox[b].grow(prims[i].box);
            }
            float cost[kBuckets - 1];
            for (int i = 0; i < kBuckets - 1; ++i)
            {
                Box b0, b1;
                int c0 = 0, c1 = 0;
                for (int j = 0; j <= i; ++j) { b0.grow(bbox[j]); c0 += count[j]; }
                for (int j = i + 1; j < kBuckets; ++j) { b1.grow(bbox[j]); c1 += count[j]; }
                cost[i] = 1.0f + (c0 * b0.area() + c1 * b1.area()) / box.area();
            }
            float min_cost = cost[0];
            int split = 0;
            for (int i = 1; i < kBuckets - 1; ++i)
                if (cost[i] < min_cost) { min_cost = cost[i]; split = i; }
            if (n > 4 || min_cost < (float)n)
            {
                Prim* pm = std::partition(&prims[start], &prims[end - 1] + 1,
                                          [&](const Prim& p) { return bucket_of(cb, p.centroid, dim) <= split; });
                mid = (unsigned)(pm - &prims[0]);
            }
            else return leaf(rec, start, end, box);
        }
        // The reference builds both children as arguments of one call (bvh.cpp:212-216); with g++ (and MSVC)
        // the SECOND argument is evaluated first, which decides the leaf order of the triangle array.
        int c1 = build(mid, end);
        int c0 = build(start, mid);
        recs[rec].child[0] = c0; recs[rec].child[1] = c1;
        recs[rec].axis = (int)dim; recs[rec].count = 0;
        Box u = recs[c0].box; u.grow(recs[c1].box);
        recs[rec].box = u;
        return rec;
    }
};

} // namespace

void Bvh::BuildCPU(std::vector<Triangle>& triangles)
{
    if (triangles.empty()) throw std::runtime_error("Bvh::BuildCPU: no triangles");
    Builder b(triangles);
    b.prims.resize(triangles.size());
    for (unsigned i = 0; i < triangles.size(); ++i)
    {   // Triangle::GetBounds (shared_structures.h:134-137) and the centroid of bvh.hpp:52
        Box bx;
        bx.lo = vmin(triangles[i].v1.position, triangles[i].v2.position); bx.hi = vmax(triangles[i].v1.position, triangles[i].v2.position);
        bx.grow(triangles[i].v3.position);
        b.prims[i].index = i; b.prims[i].box = bx;
        b.prims[i].centroid = bx.lo * 0.5f + bx.hi * 0.5f;
    }
    b.recs.reserve(2 * triangles.size());
    b.ordered.reserve(triangles.size());
    // explicit stack instead of recursion depth problems: build() recurses at most tree depth (<= ~64 for sane input)
    int root = b.build(0, (unsigned)triangles.size());
    triangles.swap(b.ordered);

    // depth-first flattening (bvh.cpp:223-245): first child follows its parent, `offset` = second child
    nodes_.assign(b.recs.size(), LinearBVHNode());
    memset(nodes_.data(), 0, nodes_.size() * sizeof(LinearBVHNode));
    max_depth_ = 0;
    struct Item { int rec; int parent_slot; unsigned depth; };   // parent_slot: node whose `offset` must receive our index (-1: none)
    std::vector<Item> stack;
    stack.push_back({ root, -1, 1 });
    unsigned next = 0;
    while (!stack.empty())
    {
        Item it = stack.back(); stack.pop_back();
        const Rec& r = b.recs[it.rec];
        unsigned me = next++;
        if (it.parent_slot >= 0) nodes_[it.parent_slot].offset = me;
        if (it.depth > max_depth_) max_depth_ = it.depth;
        LinearBVHNode& n = nodes_[me];
        n.bounds_min = r.box.lo; n.bounds_max = r.box.hi;
        n.bounds_min.w = 0.0f; n.bounds_max.w = 0.0f;
        if (r.count > 0)
        {
            if (r.count >= 65536) throw std::runtime_error("Bvh::BuildCPU: leaf with >= 65536 primitives (bvh.cpp:231)");
            n.offset = (uint32_t)r.first;
            n.num_primitives_axis = (uint32_t)r.count << 16;
        }
        else
        {
            n.num_primitives_axis = (uint32_t)r.axis;
            stack.push_back({ r.child[1], (int)me, it.depth + 1 });   // second child: numbered after the whole first subtree
            stack.push_back({ r.child[0], -1, it.depth + 1 });        // first child: next index
        }
    }
    assert(next == nodes_.size());
}

} // namespace rt_host
