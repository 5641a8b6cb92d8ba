#include "bvh.hpp"

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <omp.h>

/*
 * Parallel SAH build with the reference builder's result (bvh.cpp:36-245), organised for 10 M-triangle scenes:
 *  - no shared state on the way down: a subtree over n primitives owns the 2n-1 records rec .. rec + 2n - 2 of one arena and the
 *    slots first .. first + n - 1 of the leaf order, both known before it is built, so the two children are independent OpenMP
 *    tasks without an atomic or a shared cache line between them;
 *  - nodes above `par_node` primitives (the top of the tree, where there are fewer nodes than threads) run their three passes —
 *    bounds, SAH buckets, partition — as chunked tasks whose partial results are merged in chunk order; the merges keep the
 *    reference's "first of equal values wins" (its Union is std::min / std::max over the primitives in order, which fixes the
 *    sign of a zero bound), and the partition reproduces the permutation of std::partition (k-th misplaced element from the left
 *    swaps with the k-th from the right), so the tree and the leaf order do not depend on the thread count;
 *  - the depth-first numbering needs each subtree's node count, which the build returns: flattening is parallel too.
 */
namespace rt_host
{

namespace
{

struct Box
{
    float3 lo, hi;
    void clear() { lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX); hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX); }
    // Union (mathlib.hpp:198-213): on equal values the left operand is kept
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
inline Box empty_box() { Box b; b.clear(); return b; }

struct Prim { unsigned index; int bucket; Box box; float3 centroid; };

// build record: children by arena index, -1 for leaves; `nodes` / `depth` of the subtree for the numbering pass
struct Rec { Box box; int child[2]; int axis, first, count; unsigned nodes, depth; };

constexpr int kBuckets = 12;

struct BucketSums
{
    int count[kBuckets];
    Box box[kBuckets];
    void clear() { for (int b = 0; b < kBuckets; ++b) { count[b] = 0; box[b].clear(); } }
    void merge(const BucketSums& later) { for (int b = 0; b < kBuckets; ++b) { count[b] += later.count[b]; box[b].grow(later.box[b]); } }
};

struct Builder
{
    const Triangle* tris;
    size_t n;
    std::unique_ptr<Prim[]> prims;
    std::unique_ptr<Rec[]> recs;              // 2n - 1 records; a subtree's region is fixed before it is built
    std::unique_ptr<unsigned[]> order;        // leaf order: order[slot] = index of the triangle that goes there
    std::unique_ptr<unsigned[]> scratch;      // positions to swap in the chunked partition
    unsigned task_node = 8192;                // children of larger nodes become tasks
    unsigned par_node = 1u << 18;             // larger nodes run their passes as chunked tasks
    int chunks = 1;
    std::atomic<bool> leaf_too_large{ false };

    Builder(const Triangle* t, size_t count) : tris(t), n(count) {}

    void leaf(int rec, unsigned start, unsigned end, const Box& box, unsigned first)
    {
        Rec& r = recs[rec];
        r.first = (int)first; r.count = (int)(end - start); r.box = box;
        r.child[0] = r.child[1] = -1; r.axis = 0; r.nodes = 1; r.depth = 1;
        for (unsigned i = start; i < end; ++i) order[first + (i - start)] = prims[i].index;
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

    // [start, end) cut into `chunks` pieces; piece k
    void piece(unsigned start, unsigned end, int k, unsigned& a, unsigned& b) const
    {
        unsigned long long len = end - start;
        a = start + (unsigned)(len * (unsigned long long)k / (unsigned)chunks);
        b = start + (unsigned)(len * (unsigned long long)(k + 1) / (unsigned)chunks);
    }

    void bounds_pass(unsigned start, unsigned end, Box& box, Box& cb)
    {
        box.clear(); cb.clear();
        if (end - start <= par_node || chunks == 1)
        {
            for (unsigned i = start; i < end; ++i) { box.grow(prims[i].box); cb.grow(prims[i].centroid); }
            return;
        }
        std::vector<Box> pb((size_t)chunks), pc((size_t)chunks);
        for (int k = 0; k < chunks; ++k)
        {
#pragma omp task shared(pb, pc) firstprivate(k, start, end)
            {
                unsigned a, b; piece(start, end, k, a, b);
                Box x = empty_box(), c = empty_box();
                for (unsigned i = a; i < b; ++i) { x.grow(prims[i].box); c.grow(prims[i].centroid); }
                pb[(size_t)k] = x; pc[(size_t)k] = c;
            }
        }
#pragma omp taskwait
        for (int k = 0; k < chunks; ++k) { box.grow(pb[(size_t)k]); cb.grow(pc[(size_t)k]); }     // in order: the first of equal values stays
    }

    void bucket_pass(unsigned start, unsigned end, const Box& cb, unsigned dim, BucketSums& sums)
    {
        sums.clear();
        auto run = [&](unsigned a, unsigned b, BucketSums& s) {
            for (unsigned i = a; i < b; ++i)
            {
                int k = bucket_of(cb, prims[i].centroid, dim);
                prims[i].bucket = k;
                ++s.count[k];
                s.box[k].grow(prims[i].box);
            }
        };
        if (end - start <= par_node || chunks == 1) { run(start, end, sums); return; }
        std::vector<BucketSums> part((size_t)chunks);
        for (int k = 0; k < chunks; ++k)
        {
#pragma omp task shared(part, run) firstprivate(k, start, end)
            {
                unsigned a, b; piece(start, end, k, a, b);
                part[(size_t)k].clear();
                run(a, b, part[(size_t)k]);
            }
        }
#pragma omp taskwait
        for (int k = 0; k < chunks; ++k) sums.merge(part[(size_t)k]);
    }

    // std::partition(prims[start..end), bucket <= split): returns the first position of the second group.  The chunked form yields
    // the same permutation: with mid = start + #selected, the k-th unselected element of [start, mid) from the left changes places
    // with the k-th selected element of [mid, end) from the right.
    unsigned partition_pass(unsigned start, unsigned end, int split)
    {
        if (end - start <= par_node || chunks == 1)
        {
            Prim* pm = std::partition(&prims[start], &prims[start] + (end - start), [split](const Prim& p) { return p.bucket <= split; });
            return (unsigned)(pm - &prims[0]);
        }
        std::vector<unsigned> selected((size_t)chunks);
        for (int k = 0; k < chunks; ++k)
        {
#pragma omp task shared(selected) firstprivate(k, start, end, split)
            {
                unsigned a, b, c = 0; piece(start, end, k, a, b);
                for (unsigned i = a; i < b; ++i) c += prims[i].bucket <= split;
                selected[(size_t)k] = c;
            }
        }
#pragma omp taskwait
        unsigned total = 0;
        for (int k = 0; k < chunks; ++k) total += selected[(size_t)k];
        const unsigned mid = start + total;
        // misplaced elements: unselected left of mid (listed left to right), selected right of mid (listed right to left)
        std::vector<unsigned> left_before((size_t)chunks + 1, 0), right_before((size_t)chunks + 1, 0);
        for (int k = 0; k < chunks; ++k)
        {
            unsigned a, b; piece(start, end, k, a, b);
            // per chunk: unselected in [a,b) ∩ [start,mid) and selected in [a,b) ∩ [mid,end) — counted from the chunk totals where the
            // chunk lies on one side of mid, by a scan where it straddles it
            unsigned lo_a = a, lo_b = b < mid ? b : (a < mid ? mid : a);          // part of the chunk left of mid
            unsigned hi_a = a > mid ? a : (b > mid ? mid : b), hi_b = b;          // part right of mid
            unsigned unsel_left = 0, sel_right = 0;
            if (lo_b > lo_a)
            {
                if (lo_b == b) unsel_left = (b - a) - selected[(size_t)k];
                else for (unsigned i = lo_a; i < lo_b; ++i) unsel_left += !(prims[i].bucket <= split);
            }
            if (hi_b > hi_a)
            {
                if (hi_a == a) sel_right = selected[(size_t)k];
                else for (unsigned i = hi_a; i < hi_b; ++i) sel_right += prims[i].bucket <= split;
            }
            left_before[(size_t)k + 1] = left_before[(size_t)k] + unsel_left;
            right_before[(size_t)k + 1] = right_before[(size_t)k] + sel_right;
        }
        const unsigned pairs = left_before[(size_t)chunks];
        assert(pairs == right_before[(size_t)chunks]);
        unsigned* left_pos = scratch.get() + start;                  // pairs <= (end - start) / 2: both lists fit the node's own range
        unsigned* right_pos = left_pos + pairs;
        for (int k = 0; k < chunks; ++k)
        {
#pragma omp task shared(left_before, right_before) firstprivate(k, start, end, split, mid, pairs, left_pos, right_pos)
            {
                unsigned a, b; piece(start, end, k, a, b);
                unsigned w = left_before[(size_t)k];
                for (unsigned i = a; i < b && i < mid; ++i)
                    if (!(prims[i].bucket <= split)) left_pos[w++] = i;
                // selected elements right of mid are listed from the right: this chunk's come after those of all later chunks
                unsigned r = pairs - right_before[(size_t)k + 1];
                for (unsigned i = b; i > a && i > mid; --i)
                    if (prims[i - 1].bucket <= split) right_pos[r++] = i - 1;
            }
        }
#pragma omp taskwait
        for (int k = 0; k < chunks; ++k)
        {
#pragma omp task firstprivate(k, pairs, left_pos, right_pos)
            {
                unsigned a = (unsigned)((unsigned long long)pairs * (unsigned)k / (unsigned)chunks), b = (unsigned)((unsigned long long)pairs * (unsigned)(k + 1) / (unsigned)chunks);
                for (unsigned i = a; i < b; ++i) std::swap(prims[left_pos[i]], prims[right_pos[i]]);
            }
        }
#pragma omp taskwait
        return mid;
    }

    // `rec`: first record of this subtree's region; `first`: where its triangles start in the leaf order.  The reference appends
    // leaves in construction order and constructs the SECOND child first (bvh.cpp:212-216: both children are arguments of one
    // call, g++ and MSVC evaluate the last argument first), so the second child's triangles occupy [first, first + (end - mid)) and
    // the first child's follow.
    void build(int rec, unsigned start, unsigned end, unsigned first)
    {
        Box box, cb;
        bounds_pass(start, end, box, cb);
        unsigned count = end - start;
        if (count == 1) { leaf(rec, start, end, box, first); return; }
        unsigned dim = cb.widest();
        if (component(cb.hi, dim) == component(cb.lo, dim)) { leaf(rec, start, end, box, first); return; }   // all centroids coincide

        unsigned mid = (start + end) / 2;
        if (count <= 2)
        {
            std::nth_element(&prims[start], &prims[mid], &prims[start] + count,
                             [dim](const Prim& a, const Prim& b) { return component(a.centroid, dim) < component(b.centroid, dim); });
        }
        else
        {
            BucketSums sums;
            bucket_pass(start, end, cb, dim, sums);
            // cost[i] = 1 + (n0 * area(U b[0..i]) + n1 * area(U b[i+1..11])) / area(node)   (bvh.cpp:149-166).  The reference
            // recomputes both unions for every i; prefix and suffix unions give the same boxes up to the sign of a zero bound,
            // which an area does not see.
            Box left[kBuckets], right[kBuckets];
            int nleft[kBuckets], nright[kBuckets];
            left[0] = sums.box[0]; nleft[0] = sums.count[0];
            for (int j = 1; j < kBuckets; ++j) { left[j] = left[j - 1]; left[j].grow(sums.box[j]); nleft[j] = nleft[j - 1] + sums.count[j]; }
            right[kBuckets - 1] = sums.box[kBuckets - 1]; nright[kBuckets - 1] = sums.count[kBuckets - 1];
            for (int j = kBuckets - 2; j >= 0; --j) { right[j] = right[j + 1]; right[j].grow(sums.box[j]); nright[j] = nright[j + 1] + sums.count[j]; }
            float cost[kBuckets - 1];
            const float node_area = box.area();
            for (int i = 0; i < kBuckets - 1; ++i)
                cost[i] = 1.0f + (nleft[i] * left[i].area() + nright[i + 1] * right[i + 1].area()) / node_area;
            float min_cost = cost[0];
            int split = 0;
            for (int i = 1; i < kBuckets - 1; ++i)
                if (cost[i] < min_cost) { min_cost = cost[i]; split = i; }
            if (count > 4 || min_cost < (float)count) mid = partition_pass(start, end, split);
            else { leaf(rec, start, end, box, first); return; }
        }
        const unsigned n1 = end - mid;                       // second child
        const int c1 = rec + 1, c0 = rec + (int)(2 * n1);    // regions: [rec+1, rec+2*n1-1], then the first child's
        if (count > task_node)
        {
#pragma omp task firstprivate(c1, mid, end, first)
            build(c1, mid, end, first);
#pragma omp task firstprivate(c0, start, mid, first, n1)
            build(c0, start, mid, first + n1);
#pragma omp taskwait
        }
        else
        {
            build(c1, mid, end, first);
            build(c0, start, mid, first + n1);
        }
        Rec& r = recs[rec];
        r.child[0] = c0; r.child[1] = c1;
        r.axis = (int)dim; r.first = 0; r.count = 0;
        Box u = recs[c0].box; u.grow(recs[c1].box);
        r.box = u;
        r.nodes = 1 + recs[c0].nodes + recs[c1].nodes;
        r.depth = 1 + std::max(recs[c0].depth, recs[c1].depth);
    }

    // depth-first numbering (bvh.cpp:223-245): the first child follows its parent, `offset` = index of the second child
    void number(int rec, unsigned index, LinearBVHNode* nodes)
    {
        for (;;)
        {
            const Rec& r = recs[rec];
            LinearBVHNode& out = nodes[index];
            out.bounds_min = r.box.lo; out.bounds_max = r.box.hi;
            out.bounds_min.w = 0.0f; out.bounds_max.w = 0.0f;
            if (r.child[0] < 0)
            {
                if (r.count >= 65536) leaf_too_large = true;        // the count has 16 bits (bvh.cpp:231); reported by the caller
                out.offset = (uint32_t)r.first;
                out.num_primitives_axis = (uint32_t)r.count << 16;
                return;
            }
            const unsigned second = index + 1 + recs[r.child[0]].nodes;
            out.offset = second;
            out.num_primitives_axis = (uint32_t)r.axis;
            if (r.nodes > task_node)
            {
                const int c1 = r.child[1];
#pragma omp task firstprivate(c1, second, nodes)
                number(c1, second, nodes);
            }
            else number(r.child[1], second, nodes);
            rec = r.child[0]; index = index + 1;
        }
    }
};

unsigned env_unsigned(const char* name, unsigned fallback)
{
    const char* v = getenv(name);
    return v && *v ? (unsigned)strtoul(v, nullptr, 10) : fallback;
}

} // namespace

void Bvh::BuildCPU(std::vector<Triangle>& triangles)
{
    if (triangles.empty()) throw std::runtime_error("Bvh::BuildCPU: no triangles");
    Build(triangles.data(), triangles.size());
}

void Bvh::Build(Triangle* triangles, size_t count)
{
    if (!triangles || count == 0) throw std::runtime_error("Bvh::BuildCPU: no triangles");
    if (count >= (1u << 30)) throw std::runtime_error("Bvh::BuildCPU: more than 2^30 triangles");
    const bool verbose = getenv("RT_BVH_VERBOSE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!verbose) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[bvh] %-10s %.2f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    Builder b(triangles, count);
    b.task_node = env_unsigned("RT_BVH_TASK_NODE", b.task_node);
    b.par_node = env_unsigned("RT_BVH_PARALLEL_NODE", b.par_node);
    const int threads = (int)env_unsigned("RT_BVH_THREADS", (unsigned)omp_get_max_threads());
    b.chunks = threads > 1 ? 4 * threads : 1;
    b.prims.reset(new Prim[count]);
    b.recs.reset(new Rec[2 * count]);
    b.order.reset(new unsigned[count]);
    b.scratch.reset(new unsigned[count]);
    const long long n = (long long)count;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long long i = 0; i < n; ++i)
    {   // Triangle::GetBounds (shared_structures.h:134-137) and the centroid of bvh.hpp:52
        Box bx;
        bx.lo = vmin(triangles[i].v1.position, triangles[i].v2.position); bx.hi = vmax(triangles[i].v1.position, triangles[i].v2.position);
        bx.grow(triangles[i].v3.position);
        Prim& p = b.prims[(size_t)i];
        p.index = (unsigned)i; p.bucket = 0; p.box = bx;
        p.centroid = bx.lo * 0.5f + bx.hi * 0.5f;
    }
    lap("prims");
#pragma omp parallel num_threads(threads)
#pragma omp single
    b.build(0, 0, (unsigned)count, 0);
    lap("build");
    b.prims.reset();
    b.scratch.reset();

    // triangles into leaf order: through a copy, both passes parallel
    {
        std::unique_ptr<Triangle[]> copy(new Triangle[count]);
#pragma omp parallel for schedule(static) num_threads(threads)
        for (long long i = 0; i < n; ++i) memcpy(&copy[(size_t)i], &triangles[i], sizeof(Triangle));
#pragma omp parallel for schedule(static) num_threads(threads)
        for (long long i = 0; i < n; ++i) memcpy(&triangles[i], &copy[b.order[(size_t)i]], sizeof(Triangle));
    }
    b.order.reset();
    lap("reorder");

    const Rec& root = b.recs[0];
    nodes_.clear();
    nodes_.resize(root.nodes);                   // zero-initialised: the padding words of a node are 0
    max_depth_ = root.depth;
#pragma omp parallel num_threads(threads)
#pragma omp single
    b.number(0, 0, nodes_.data());
    if (b.leaf_too_large) throw std::runtime_error("Bvh::BuildCPU: leaf with >= 65536 primitives (bvh.cpp:231)");
    lap("number");
}

} // namespace rt_host
