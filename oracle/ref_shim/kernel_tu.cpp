/*
 * kernel_tu.cpp — TEST INFRASTRUCTURE (oracle side).  One translation unit per
 * reference kernel file / -D variant.  oracle/build_ref.py compiles it with
 *   -DRT_KERNEL_FILE="src/kernels/cl/<file>.cl"   (the syntactically rewritten
 *                                                  copy under oracle/_ref/gen/)
 *   -DRT_KERNEL_ID=<n>  -DRT_SUFFIX=<variant suffix>  -DRT_NS=<unique namespace>  [-DSHADOW_RAYS ...]
 * The reference kernel is included verbatim inside a private namespace (the .cl
 * files define their helper functions without `static`, so each needs its own),
 * and an extern "C" launcher plays the part of clEnqueueNDRangeKernel
 * (cl_context.cpp:115-119): 1-D range, work_size items, OpenMP over work-items.
 */
#include "clshim.h"

#define RT_CAT2(a, b) a##b
#define RT_CAT(a, b) RT_CAT2(a, b)
#define RT_LAUNCH(name) extern "C" void RT_CAT(RT_CAT(refk_, name), RT_SUFFIX)

#define RT_FOR_EACH_ITEM(work_size, call)                                      \
    _Pragma("omp parallel for schedule(dynamic, 1024)")                        \
    for (long long rt_i = 0; rt_i < (long long)(work_size); ++rt_i)            \
    {                                                                          \
        t_global_id = (size_t)rt_i;                                            \
        call;                                                                  \
    }

namespace clc
{
namespace RT_NS
{

#include RT_KERNEL_FILE

#if RT_KERNEL_ID == 1   // raygeneration.cl
RT_LAUNCH(RayGeneration)(size_t work_size, uint width, uint height, const void* camera, void* sample_counter,
    void* rays, void* ray_counter, void* pixel_indices, void* throughputs,
    void* diffuse_albedo, void* depth, void* normal, void* velocity)
{
    Camera cam; memcpy(&cam, camera, sizeof(Camera));   // caller storage may be only 4-byte aligned
    RT_FOR_EACH_ITEM(work_size, RayGeneration(width, height, cam, (uint*)sample_counter, (Ray*)rays,
        (uint*)ray_counter, (uint*)pixel_indices, (float3*)throughputs, (float3*)diffuse_albedo,
        (float*)depth, (float3*)normal, (float2*)velocity))
}
#elif RT_KERNEL_ID == 2 // trace_bvh.cl (closest hit, or any hit with -DSHADOW_RAYS)
RT_LAUNCH(TraceBvh)(size_t work_size, void* rays, void* ray_counter, void* triangles, void* nodes, void* out)
{
#ifdef SHADOW_RAYS
    RT_FOR_EACH_ITEM(work_size, TraceBvh((Ray*)rays, (uint*)ray_counter, (RTTriangle*)triangles,
        (LinearBVHNode*)nodes, (uint*)out))
#else
    RT_FOR_EACH_ITEM(work_size, TraceBvh((Ray*)rays, (uint*)ray_counter, (RTTriangle*)triangles,
        (LinearBVHNode*)nodes, (Hit*)out))
#endif
}
#elif RT_KERNEL_ID == 3 // miss.cl
RT_LAUNCH(Miss)(size_t work_size, void* rays, void* ray_counter, void* hits, void* pixel_indices,
    void* throughputs, float* env_data, int env_width, int env_height, void* radiance)
{
    image2d_t tex = { env_data, env_width, env_height };
    RT_FOR_EACH_ITEM(work_size, Miss((Ray*)rays, (uint*)ray_counter, (Hit*)hits, (uint*)pixel_indices,
        (float3*)throughputs, tex, (float3*)radiance))
}
#elif RT_KERNEL_ID == 4 // hit_surface.cl
RT_LAUNCH(HitSurface)(size_t work_size, void* incoming_rays, void* incoming_ray_counter, void* incoming_pixel_indices,
    void* hits, void* triangles, void* analytic_lights, void* emissive_indices, void* materials,
    void* textures, void* texture_data, uint bounce, uint width, uint height, void* sample_counter,
    const void* scene_info, void* sobol, void* scrambling, void* ranking,
    void* throughputs, void* outgoing_rays, void* outgoing_ray_counter, void* outgoing_pixel_indices,
    void* shadow_rays, void* shadow_ray_counter, void* shadow_pixel_indices, void* direct_light_samples,
    void* radiance)
{
    SceneInfo info; memcpy(&info, scene_info, sizeof(SceneInfo));
    RT_FOR_EACH_ITEM(work_size, HitSurface((Ray*)incoming_rays, (uint*)incoming_ray_counter,
        (uint*)incoming_pixel_indices, (Hit*)hits, (Triangle*)triangles, (Light*)analytic_lights,
        (uint*)emissive_indices, (PackedMaterial*)materials, (Texture*)textures, (uint*)texture_data,
        bounce, width, height, (uint*)sample_counter, info, (int*)sobol, (int*)scrambling, (int*)ranking,
        (float3*)throughputs, (Ray*)outgoing_rays, (uint*)outgoing_ray_counter, (uint*)outgoing_pixel_indices,
        (Ray*)shadow_rays, (uint*)shadow_ray_counter, (uint*)shadow_pixel_indices,
        (float3*)direct_light_samples, (float4*)radiance))
}
#elif RT_KERNEL_ID == 5 // accumulate_direct_samples.cl
RT_LAUNCH(AccumulateDirectSamples)(size_t work_size, void* shadow_hits, void* shadow_ray_counter,
    void* shadow_pixel_indices, void* direct_light_samples, void* radiance)
{
    RT_FOR_EACH_ITEM(work_size, AccumulateDirectSamples((uint*)shadow_hits, (uint*)shadow_ray_counter,
        (uint*)shadow_pixel_indices, (float3*)direct_light_samples, (float4*)radiance))
}
#elif RT_KERNEL_ID == 6 // clear_counter.cl
RT_LAUNCH(ClearCounter)(size_t work_size, void* counter)
{
    RT_FOR_EACH_ITEM(work_size, ClearCounter((uint*)counter))
}
#elif RT_KERNEL_ID == 7 // increment_counter.cl
RT_LAUNCH(IncrementCounter)(size_t work_size, void* counter)
{
    RT_FOR_EACH_ITEM(work_size, IncrementCounter((uint*)counter))
}
#elif RT_KERNEL_ID == 8 // reset_radiance.cl
RT_LAUNCH(ResetRadiance)(size_t work_size, uint width, uint height, void* radiance)
{
    RT_FOR_EACH_ITEM(work_size, ResetRadiance(width, height, (float4*)radiance))
}
#elif RT_KERNEL_ID == 9 // aov.cl
RT_LAUNCH(GenerateAOV)(size_t work_size, void* rays, void* ray_counter, void* pixel_indices, void* hits,
    void* triangles, void* materials, void* textures, void* texture_data, uint width, uint height,
    const void* camera, const void* prev_camera, void* diffuse_albedo, void* depth, void* normal, void* velocity)
{
    Camera cam; memcpy(&cam, camera, sizeof(Camera));   // caller storage may be only 4-byte aligned
    Camera prev; memcpy(&prev, prev_camera, sizeof(Camera));
    RT_FOR_EACH_ITEM(work_size, GenerateAOV((Ray*)rays, (uint*)ray_counter, (uint*)pixel_indices, (Hit*)hits,
        (Triangle*)triangles, (PackedMaterial*)materials, (Texture*)textures, (uint*)texture_data,
        width, height, cam, prev, (float3*)diffuse_albedo, (float*)depth, (float3*)normal, (float2*)velocity))
}
#elif RT_KERNEL_ID == 10 // denoiser.cl
RT_LAUNCH(TemporalAccumulation)(size_t work_size, uint width, uint height, void* radiance, void* prev_radiance,
    void* depth, void* prev_depth, void* motion_vectors)
{
    RT_FOR_EACH_ITEM(work_size, TemporalAccumulation(width, height, (float4*)radiance, (float4*)prev_radiance,
        (float*)depth, (float*)prev_depth, (float2*)motion_vectors))
}
#elif RT_KERNEL_ID == 11 // resolve_radiance.cl
RT_LAUNCH(ResolveRadiance)(size_t work_size, uint width, uint height, uint aov_index, void* radiance,
    void* diffuse_albedo, void* depth, void* normal, void* motion_vectors, void* sample_counter, float* result)
{
    image2d_t img = { result, (int)width, (int)height };
    RT_FOR_EACH_ITEM(work_size, ResolveRadiance(width, height, aov_index, (float4*)radiance,
        (float3*)diffuse_albedo, (float*)depth, (float3*)normal, (float2*)motion_vectors,
        (uint*)sample_counter, img))
}
#else
#error "unknown RT_KERNEL_ID"
#endif

} // namespace k_<suffix>
} // namespace clc
