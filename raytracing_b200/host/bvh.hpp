/*
 * Host SAH BVH builder producing the reference's LinearBVHNode[] layout
 * (/root/reference/src/bvh.hpp:30-103, src/bvh.cpp:36-245).  The build stays on the host
 * (BASELINE.json north_star).  Same decisions as the reference builder — 12-bucket SAH on the
 * centroid-bounds' widest axis, median split for <= 2 primitives, leaf when <= 4 primitives and
 * not worth splitting, leaf of any size when all centroids coincide — but an arena of build
 * records and index ranges instead of a `new` per node (the reference never frees them,
 * bvh.cpp:77), and every pass parallel (bvh.cpp here), which is what makes a 10 M-triangle
 * build practical.  The tree does not depend on the number of threads.
 */
#pragma once

#include <vector>

#include "reference_api.hpp"

namespace rt_host
{

class Bvh : public AccelerationStructure
{
public:
    void BuildCPU(std::vector<Triangle>& triangles) override;
    // the same build over caller-owned memory: `triangles` are reordered in place into leaf order
    void Build(Triangle* triangles, size_t count);
    std::vector<LinearBVHNode> const& GetNodes() const override { return nodes_; }
    unsigned MaxDepth() const { return max_depth_; }

private:
    std::vector<LinearBVHNode> nodes_;
    unsigned max_depth_ = 0;
};

} // namespace rt_host
