#include "bvh.hpp"

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
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
    std::vector<Rec> recs;            // pre-sized arena; records are claimed with an atomic cursor
    std::vector<Triangle> ordered;    // pre-sized; every leaf knows its offset before it is built
    std::atomic<int> next_rec{ 0 };

    explicit Builder(const std::vector<Triangle>& t) : tris(t) {}

    int leaf(int rec, unsigned start, unsigned end, const Box& box, unsigned first)
    {
        recs[rec].first = (int)first;
        recs[rec].count = (int)(end - start);
        recs[rec].box = box;
        recs[rec].child[0] = recs[rec].child[1] = -1;
        for (unsigned i = start; i < end; ++i) ordered[first + (i - start)] = tris[prims[i].index];
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

    // `first`: where this subtree's triangles start in the leaf-ordered output.  The reference appends leaves in
    // construction order and constructs the SECOND child first (see below), so the second child's triangles
    // occupy [first, first + (end-mid)) and the first child's follow — known before either child is built,
    // which is what lets the two children be built as independent OpenMP tasks with an identical result.
    int build(unsigned start, unsigned end, unsigned first)
    {
        int rec = next_rec.fetch_add(1);
        Box box;
        for (unsigned i = start; i < end; ++i) box.grow(prims[i].box);
        unsigned n = end - start;
        if (n == 1) return leaf(rec, start, end, box, first);

        Box cb;
        for (unsigned i = start; i < end; ++i) cb.grow(prims[i].centroid);
        unsigned dim = cb.widest();
        if (component(cb.hi, dim) == component(cb.lo, dim)) return leaf(rec, start, end, box, first);   // all centroids coincide

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
            // cost[i] = 1 + (n0 * area(U b[0..i]) + n1 * area(U b[i+1..11])) / area(node)   (bvh.cpp:149-166).
            // The reference recomputes both unions from scratch for every i (O(buckets^2) per node, which
            // dominates the build of the many small nodes); min/max are exact and associative, so prefix and
            // suffix unions give the same boxes in O(buckets).
            Box left[kBuckets], right[kBuckets];
            int nleft[kBuckets], nright[kBuckets];
            left[0] = bbox[0]; nleft[0] = count[0];
            for (int j = 1; j < kBuckets; ++j) { left[j] = left[j - 1]; left[j].grow(bbox[j]); nleft[j] = nleft[j - 1] + count[j]; }
            right[kBuckets - 1] = bbox[kBuckets - 1]; nright[kBuckets - 1] = count[kBuckets - 1];
            for (int j = kBuckets - 2; j >= 0; --j) { right[j] = right[j + 1]; right[j].grow(bbox[j]); nright[j] = nright[j + 1] + count[j]; }
            float cost[kBuckets - 1];
            const float node_area = box.area();
            for (int i = 0; i < kBuckets - 1; ++i)
                cost[i] = 1.0f + (nleft[i] * left[i].area() + nright[i + 1] * right[i + 1].area()) / node_area;
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
            else return leaf(rec, start, end, box, first);
        }
        // The reference builds both children as arguments of one call (bvh.cpp:212-216); with g++ (and MSVC)
        // the SECOND argument is evaluated first, which decides the leaf order of the triangle array.
        int c0 = -1, c1 = -1;
        const unsigned kTaskThreshold = 8192;
        if (n > kTaskThreshold)
        {
#pragma omp task shared(c1) firstprivate(mid, end, first)
            c1 = build(mid, end, first);
#pragma omp task shared(c0) firstprivate(start, mid, end, first)
            c0 = build(start, mid, first + (end - mid));
#pragma omp taskwait
        }
        else
        {
            c1 = build(mid, end, first);
            c0 = build(start, mid, first + (end - mid));
        }
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
    const bool verbose = getenv("RT_BVH_VERBOSE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!verbose) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[bvh] %-10s %.2f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
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
    lap("prims");
    b.recs.resize(2 * triangles.size());          // a binary tree over n >= 1 primitives has at most 2n-1 nodes
    b.ordered.resize(triangles.size(), triangles[0]);
    int root = 0;
#pragma omp parallel
#pragma omp single
    root = b.build(0, (unsigned)triangles.size(), 0);
    lap("build");
    b.recs.resize((size_t)b.next_rec.load());
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
    lap("flatten");
}

} // namespace rt_host
