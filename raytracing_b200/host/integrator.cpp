#include "integrator.hpp"

namespace rt_host
{

// The per-frame wavefront schedule, integrator.cpp:27-59: the bounce loop is INCLUSIVE of max_bounces.
void Integrator::Integrate()
{
    if (request_reset_ || enable_denoiser_)
    {
        Reset();
        request_reset_ = false;
    }
    GenerateRays();
    for (std::uint32_t bounce = 0; bounce <= max_bounces_; ++bounce)
    {
        IntersectRays(bounce);
        if (bounce == 0) ComputeAOVs();
        ShadeMissedRays(bounce);
        ClearOutgoingRayCounter(bounce);
        ClearShadowRayCounter();
        ShadeSurfaceHits(bounce);
        IntersectShadowRays();
        AccumulateDirectSamples();
    }
    AdvanceSampleCount();
    if (enable_denoiser_)
    {
        Denoise();
        CopyHistoryBuffers();
    }
    ResolveRadiance();
}

void Integrator::SetMaxBounces(std::uint32_t max_bounces)
{
    max_bounces_ = max_bounces;
    RequestReset();
}

void Integrator::EnableWhiteFurnace(bool enable)
{
    if (enable == enable_white_furnace_) return;
    enable_white_furnace_ = enable;
    CreateKernels();
    RequestReset();
}

} // namespace rt_host
