/*
 * Mirror of the reference's abstract Integrator (/root/reference/src/integrator/integrator.hpp:34-100,
 * integrator.cpp:27-77): same public interface, same protected virtual steps, same fixed wavefront
 * schedule in Integrate().  A backend plugs in by overriding the steps; CUDAPathTraceIntegrator
 * (cuda_pt_integrator.hpp) is the B200 one, next to the reference's OpenCL and OpenGL backends.
 */
#pragma once

#include <cstdint>

#include "acceleration_structure.hpp"
#include "types.hpp"

namespace rt_host
{

class Scene;

class Integrator
{
public:
    enum class SamplerType { kRandom, kBlueNoise };
    enum AOV { kShadedColor, kDiffuseAlbedo, kDepth, kNormal, kMotionVectors };

    Integrator(std::uint32_t width, std::uint32_t height, AccelerationStructure& acc_structure)
        : width_(width), height_(height), acc_structure_(acc_structure) {}
    virtual ~Integrator() = default;

    void Integrate();
    virtual void UploadGPUData(Scene const& scene, AccelerationStructure const& acc_structure) = 0;
    virtual void SetCameraData(Camera const& camera) = 0;
    void RequestReset() { request_reset_ = true; }
    void EnableWhiteFurnace(bool enable);
    void SetMaxBounces(std::uint32_t max_bounces);
    virtual void SetSamplerType(SamplerType sampler_type) = 0;
    virtual void SetAOV(AOV aov) = 0;
    virtual void EnableDenoiser(bool enable) = 0;

protected:
    virtual void CreateKernels() = 0;
    virtual void Reset() = 0;
    virtual void AdvanceSampleCount() = 0;
    virtual void GenerateRays() = 0;
    virtual void IntersectRays(std::uint32_t bounce) = 0;
    virtual void ComputeAOVs() = 0;
    virtual void ShadeMissedRays(std::uint32_t bounce) = 0;
    virtual void ShadeSurfaceHits(std::uint32_t bounce) = 0;
    virtual void IntersectShadowRays() = 0;
    virtual void AccumulateDirectSamples() = 0;
    virtual void ClearOutgoingRayCounter(std::uint32_t bounce) = 0;
    virtual void ClearShadowRayCounter() = 0;
    virtual void Denoise() = 0;
    virtual void CopyHistoryBuffers() = 0;
    virtual void ResolveRadiance() = 0;

    std::uint32_t width_;
    std::uint32_t height_;
    AccelerationStructure& acc_structure_;
    Camera camera_ = {};
    Camera prev_camera_ = {};
    std::uint32_t max_bounces_ = 3u;
    SamplerType sampler_type_ = SamplerType::kRandom;
    AOV aov_ = AOV::kShadedColor;
    bool request_reset_ = false;
    bool enable_white_furnace_ = false;
    bool enable_denoiser_ = false;
};

} // namespace rt_host
