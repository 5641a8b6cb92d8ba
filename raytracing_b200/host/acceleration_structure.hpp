/*
 * Mirror of the reference's AccelerationStructure interface
 * (/root/reference/src/acceleration_structure.hpp:31-38): BuildCPU reorders the
 * triangle array in place into leaf order; GetNodes returns the depth-first node array.
 */
#pragma once

#include <vector>

#include "types.hpp"

namespace rt_host
{

class AccelerationStructure
{
public:
    virtual ~AccelerationStructure() = default;
    virtual void BuildCPU(std::vector<Triangle>& triangles) = 0;
    virtual std::vector<LinearBVHNode> const& GetNodes() const = 0;
};

} // namespace rt_host
