/*
 * rt_kernels.cu — the sm_100a kernels of the wavefront path tracer and the C ABI
 * (include/rt_b200.h) that launches them.  One translation unit so that the shading
 * code inlines into both the stepwise kernels (one per reference kernel) and the fused
 * ones.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (see
 * __graft_entry__.build()).  There is no CPU fallback anywhere in this file.
 *
 * Data layout in HBM (all SoA streams of float4, indexed by compacted ray slot):
 *   extension-ray queue q[b&1]:  A = (origin.xyz, pixel_idx bits)
 *                                B = (dir.xyz, t_max)
 *                                C = (path throughput.xyz, -)
 *   shadow-ray queue:            A = (origin.xyz, pixel_idx bits)
 *                                B = (dir.xyz, t_max = distance to light)
 *                                C = (light sample.xyz, -)
 *   hit queue:                   (bc.x, bc.y, primitive_id bits, ray slot bits)   — rays that hit, compacted
 *   miss queue:                  ray slot (uint32)                                — rays that missed, compacted
 *   hits (stepwise path only):   (bc.x, bc.y, primitive_id bits, t)
 *   radiance:                    float4 per LOCAL pixel (scanline partition, see rt_set_partition)
 * "pixel_idx bits" is the pixel packed as x | y << 16 (the RNG and the radiance index need x and y, never the
 * linear index, and this saves two integer divisions per path vertex); the API converts back.
 * The reference keeps throughput in a per-pixel buffer that every bounce scatters to
 * (hit_surface.cl:105,166); there is at most one live path per pixel, so carrying it in
 * the ray stream is equivalent and turns the scatter into a coalesced stream.
 * Ray counters are per bounce (q_count[b], shadow_count[b]): nothing has to be cleared
 * between kernels (the reference launches two 1-thread clear kernels per bounce,
 * integrator.cpp:45-46) and the per-bounce statistics fall out for free.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rt_b200.h"
#include "rt_bvh_layout.h"
#include "rt_device.cuh"

using namespace rt;

// minimum resident CTAs per SM requested from ptxas for the hot kernels (register budget = 65536 / (256 * N)):
// tuned on B200, see DESIGN.md section 6
#ifndef RT_MINB_TRACE
#define RT_MINB_TRACE 5
#endif
#ifndef RT_MINB_SHADE
#define RT_MINB_SHADE 4
#endif
// The whole-frame kernel (k_frame) is compiled for up to 1024 threads per CTA (64 registers): the CTA size is chosen at launch
// (RT_OPT_FRAME_THREADS) — 1024 threads = one CTA per SM, all 32 warps of an SM in the same phase; 256 threads = 4 CTAs per SM
// in different phases
#define RT_FRAME_MAX_THREADS 1024

namespace
{

#include "rt_kernel_types.cuh"
#include "rt_traverse.cuh"

// ------------------------------------------------------------------------------------ shading
struct ShadeOut
{
    bool spawn_shadow, spawn_next, emissive;
    f3 emission_add;
    f3 s_origin, s_dir, s_sample; float s_tmax;
    f3 n_origin, n_dir, n_throughput;
};

// kernels/cl/miss.cl:41-77: radiance += sky * throughput
__device__ __forceinline__ void shade_miss(const DevScene& sc, const FrameParams& p, float4* radiance, uint32_t pxy, f3 dir, f3 throughput)
{
    f3 sky = p.white_furnace ? splat(0.5f) : sample_sky(sc, dir);
    f3 add = sky * throughput;
    uint32_t li = local_index(p, pxy);
    float4 r = radiance[li];
    r.x += add.x; r.y += add.y; r.z += add.z;
    radiance[li] = r;
    if (p.gather) p.gather[li] = r;        // the path ends here: its pixel is final (its last shadow ray was accumulated a phase ago)
}

// kernels/cl/hit_surface.cl:30-186
__device__ __forceinline__ void shade_hit(const DevScene& sc, const FrameParams& p, const AovParams& aov, uint32_t bounce, uint32_t pxy,
                                          f3 ray_origin, f3 ray_dir, f3 hit_throughput, uint32_t prim, float u, float v, ShadeOut& out)
{
    f3 incoming = -ray_dir;
    uint32_t px = pxy & 0xFFFFu, py = pxy >> 16;
    const float4* tp = sc.tri_shade + (size_t)prim * 7;
    float4 r0 = __ldg(tp), r1 = __ldg(tp + 1), r2 = __ldg(tp + 2), r3 = __ldg(tp + 3), r4 = __ldg(tp + 4), r5 = __ldg(tp + 5);
    f3 p1 = mk3(r0), p2 = mk3(r1), p3 = mk3(r2), n1 = mk3(r3), n2 = mk3(r4), n3 = mk3(r5);
    uint32_t mtl = __float_as_uint(r0.w);
    f3 geometry_normal = mk3(r1.w, r2.w, r3.w);       // normalize(cross(p2-p1, p3-p1)), precomputed at upload
    float w0 = 1.0f - u - v;
    f3 position = p1 * w0 + p2 * u + p3 * v;
    f3 normal = normalize(n1 * w0 + n2 * u + n3 * v);
    f2 texcoord; texcoord.x = 0.0f; texcoord.y = 0.0f;
    if (__float_as_uint(__ldg(sc.mat_rec + (size_t)mtl * 4 + 3).y) != 0u)
    {   // texture coordinates are only needed by textured materials (hit_surface.cl:96-97)
        float4 r6 = __ldg(tp + 6);
        texcoord.x = r4.w * w0 + r6.x * u + r6.z * v; texcoord.y = r5.w * w0 + r6.y * u + r6.w * v;
    }
    Material material = load_material(sc, mtl, texcoord);

    if (bounce == 0 && aov.enabled)
    {   // GenerateAOV, aov.cl:106-109
        uint32_t li = local_index(p, pxy);
        aov.albedo[li] = make_float4(material.diffuse_albedo.x, material.diffuse_albedo.y, material.diffuse_albedo.z, 0.0f);
        aov.depth[li] = length(ray_origin - position);
        aov.normal[li] = make_float4(normal.x, normal.y, normal.z, 0.0f);
        f2 s0 = project_screen(position, p.dyn->cam), s1 = project_screen(position, p.dyn->prev);
        aov.velocity[li] = make_float2(s0.x - s1.x, s0.y - s1.y);
    }

    out.emissive = false;
    if (!p.white_furnace && dot(material.emission, splat(1.0f)) > 0.0f)
    {
        out.emissive = true;
        out.emission_add = hit_throughput * material.emission;
    }
    uint32_t pixel_seed = sample_seed_pixel(px, py, p.dyn->sample_idx);
    {   // direct lighting (next-event estimation on analytic lights)
        float s_light = p.bn ? sample_blue_noise(p.bn, px, py, p.dyn->sample_idx, bounce * 5u + SAMPLE_LIGHT)
                             : sample_random(pixel_seed, bounce, SAMPLE_LIGHT);
        f3 outgoing; float pdf, distance_to_light;
        f3 light_radiance = light_sample(sc, position, s_light, outgoing, distance_to_light, pdf);
        // light_sample = L * throughput * brdf / pdf * max(n.l, 0); a shadow ray is spawned iff pdf > 0 and
        // dot(sample, sample) > 0 (hit_surface.cl:127-129).  When max(n.l, 0) is 0 every component of the sample is
        // +-0 or NaN whatever the BRDF evaluates to, so the ray is never spawned: the BRDF evaluation is skipped then.
        const float cos_l = fmaxf(dot(outgoing, normal), 0.0f);
        out.spawn_shadow = false;
        if (cos_l > 0.0f)
        {
            f3 brdf = evaluate_material(material, normal, incoming, outgoing);
            f3 ls = light_radiance * hit_throughput * brdf / pdf * cos_l;
            out.spawn_shadow = (pdf > 0.0f) && (dot(ls, ls) > 0.0f);
            out.s_origin = position + normal * RT_EPS;
            out.s_dir = outgoing; out.s_tmax = distance_to_light; out.s_sample = ls;
        }
    }
    {   // BSDF sampling
        f2 s; float s1;
        if (p.bn)
        {
            const uint32_t sample_idx = p.dyn->sample_idx;
            s.x = sample_blue_noise(p.bn, px, py, sample_idx, bounce * 5u + SAMPLE_U);
            s.y = sample_blue_noise(p.bn, px, py, sample_idx, bounce * 5u + SAMPLE_V);
            s1 = sample_blue_noise(p.bn, px, py, sample_idx, bounce * 5u + SAMPLE_LAYER);
        }
        else
        {
            s.x = sample_random(pixel_seed, bounce, SAMPLE_U); s.y = sample_random(pixel_seed, bounce, SAMPLE_V);
            s1 = sample_random(pixel_seed, bounce, SAMPLE_LAYER);
        }
        float pdf = 0.0f, offset = 1.0f;
        f3 outgoing = mk3(0.0f, 0.0f, 0.0f);
        f3 bxdf = sample_bxdf(s1, s, material, normal, incoming, p.white_furnace != 0, outgoing, pdf, offset);
        f3 throughput = splat(0.0f);
        if (pdf > 0.0f) throughput = bxdf / pdf;
        out.n_throughput = hit_throughput * throughput;
        out.spawn_next = pdf > 0.0f;
        out.n_origin = position + geometry_normal * RT_EPS * offset;
        out.n_dir = outgoing;
    }
}

// ------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256) k_reset(float4* radiance, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) radiance[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// raygeneration.cl:65-139.  One thread per LOCAL pixel; slot i of queue 0 = local pixel i.
__global__ void __launch_bounds__(256) k_raygen(FrameParams p, Queues q, DevCounters* ctr, AovParams aov)
{
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li == 0) ctr->n_primary = p.n_local;
    if (li >= p.n_local) return;
    uint32_t px = li % p.width, py = (li / p.width) * p.world + p.rank;
    uint32_t pixel = py * p.width + px;
    f3 o, d;
    generate_primary_ray(p.dyn->raygen, pixel, px, py, p.dyn->sample_idx, o, d);
    q.A[0][li] = make_float4(o.x, o.y, o.z, __uint_as_float(pack_pixel(px, py)));
    q.B[0][li] = make_float4(d.x, d.y, d.z, RT_MAX_RENDER_DIST);
    q.C[0][li] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    if (aov.enabled)
    {   // raygeneration.cl:129-133
        aov.albedo[li] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        aov.depth[li] = RT_MAX_RENDER_DIST;
        aov.normal[li] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        aov.velocity[li] = make_float2(0.0f, 0.0f);
    }
}

// IntersectRays (stepwise): closest hit of every live ray of bounce b -> hits[]
template <bool COUNT>
__global__ void __launch_bounds__(256) k_intersect(FrameParams p, DevScene sc, int mode, Queues q, DevCounters* ctr, uint32_t bounce)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = *in_count_ptr(ctr, bounce);
    int in = bounce & 1;
    uint32_t nv = 0, nt = 0;
    if (i < n)
    {
        float4 a = q.A[in][i], b = q.B[in][i];
        float bu = 0.0f, bv = 0.0f, bt = 0.0f;
        uint32_t prim = trace<false, COUNT>(sc, mode, mk3(a), mk3(b), 0.0f, b.w, bu, bv, bt, nv, nt);
        q.hits[i] = make_float4(bu, bv, __uint_as_float(prim), bt);
    }
    if (COUNT) { warp_sum64(&ctr->nodes_ext[bounce], nv); warp_sum64(&ctr->tris_ext[bounce], nt); }
}

// ShadeMissedRays (stepwise)
__global__ void __launch_bounds__(256) k_shade_miss(FrameParams p, DevScene sc, Queues q, DevCounters* ctr, float4* radiance, uint32_t bounce)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = *in_count_ptr(ctr, bounce);
    int in = bounce & 1;
    bool miss = false;
    if (i < n)
    {
        float4 h = q.hits[i];
        miss = __float_as_uint(h.z) == RT_INVALID_ID;
        if (miss)
        {
            float4 a = q.A[in][i], b = q.B[in][i], c = q.C[in][i];
            shade_miss(sc, p, radiance, __float_as_uint(a.w), mk3(b), mk3(c));
        }
    }
    warp_count(&ctr->hm[bounce].miss, miss);
}

__device__ __forceinline__ void emit_rays(const FrameParams& p, Queues& q, DevCounters* ctr, float4* radiance, uint32_t bounce,
                                          uint32_t pixel, bool hit, const ShadeOut& so)
{
    int out = (bounce + 1) & 1;
    if (hit && so.emissive)
    {
        uint32_t li = local_index(p, pixel);
        float4 r = radiance[li];
        r.x += so.emission_add.x; r.y += so.emission_add.y; r.z += so.emission_add.z;
        radiance[li] = r;
    }
    warp_count(&ctr->n_emissive[bounce], hit && so.emissive);
    // hit/miss stream compaction of the two output streams: __ballot + __popc give every lane its slot, ONE 64-bit
    // atomic per warp reserves the slots of both queues (shadow count in the low word, continuation count in the high)
    const bool ss = hit && so.spawn_shadow, sn = hit && so.spawn_next;
    const unsigned smask = __ballot_sync(0xffffffffu, ss), nmask = __ballot_sync(0xffffffffu, sn);
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long slot = 0ull;
    if ((smask | nmask) != 0u)
    {
        if (lane == 0)
            slot = atomicAdd((unsigned long long*)&ctr->emit[bounce], (unsigned long long)__popc(smask) | ((unsigned long long)__popc(nmask) << 32));
        slot = __shfl_sync(0xffffffffu, slot, 0);
    }
    if (ss)
    {
        uint32_t si = (uint32_t)slot + __popc(smask & lt_mask);
        q.sA[si] = make_float4(so.s_origin.x, so.s_origin.y, so.s_origin.z, __uint_as_float(pixel));
        q.sB[si] = make_float4(so.s_dir.x, so.s_dir.y, so.s_dir.z, so.s_tmax);
        q.sC[si] = make_float4(so.s_sample.x, so.s_sample.y, so.s_sample.z, 0.0f);
    }
    if (sn)
    {
        uint32_t ni = (uint32_t)(slot >> 32) + __popc(nmask & lt_mask);
        q.A[out][ni] = make_float4(so.n_origin.x, so.n_origin.y, so.n_origin.z, __uint_as_float(pixel));
        q.B[out][ni] = make_float4(so.n_dir.x, so.n_dir.y, so.n_dir.z, RT_MAX_RENDER_DIST);
        q.C[out][ni] = make_float4(so.n_throughput.x, so.n_throughput.y, so.n_throughput.z, 0.0f);
    }
}

// ShadeSurfaceHits (stepwise)
__global__ void __launch_bounds__(256) k_shade_hits(FrameParams p, DevScene sc, Queues q, DevCounters* ctr, float4* radiance, uint32_t bounce, AovParams aov)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = *in_count_ptr(ctr, bounce);
    int in = bounce & 1;
    bool hit = false;
    uint32_t pixel = 0;
    ShadeOut so;
    so.emissive = so.spawn_next = so.spawn_shadow = false;
    if (i < n)
    {
        float4 h = q.hits[i];
        uint32_t prim = __float_as_uint(h.z);
        hit = prim != RT_INVALID_ID;
        if (hit)
        {
            float4 a = q.A[in][i], b = q.B[in][i], c = q.C[in][i];
            pixel = __float_as_uint(a.w);
            shade_hit(sc, p, aov, bounce, pixel, mk3(a), mk3(b), mk3(c), prim, h.x, h.y, so);
        }
    }
    emit_rays(p, q, ctr, radiance, bounce, pixel, hit, so);
}

// IntersectShadowRays (stepwise): any-hit -> flags (0 = occluded, INVALID_ID = unoccluded)
template <bool COUNT>
__global__ void __launch_bounds__(256) k_intersect_shadow(FrameParams p, DevScene sc, int mode, Queues q, DevCounters* ctr, uint32_t bounce)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = ctr->emit[bounce].shadow;
    uint32_t nv = 0, nt = 0;
    if (i < n)
    {
        float4 a = q.sA[i], b = q.sB[i];
        float bu, bv, bt;
        q.shadow_flags[i] = trace<true, COUNT>(sc, mode, mk3(a), mk3(b), 0.0f, b.w, bu, bv, bt, nv, nt);
    }
    if (COUNT) { warp_sum64(&ctr->nodes_shadow[bounce], nv); warp_sum64(&ctr->tris_shadow[bounce], nt); }
}

// AccumulateDirectSamples (stepwise): accumulate_direct_samples.cl:27-53
__global__ void __launch_bounds__(256) k_accumulate(FrameParams p, Queues q, DevCounters* ctr, float4* radiance, uint32_t bounce)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = ctr->emit[bounce].shadow;
    bool un = false;
    if (i < n)
    {
        un = q.shadow_flags[i] == RT_INVALID_ID;
        if (un)
        {
            float4 a = q.sA[i], c = q.sC[i];
            uint32_t pixel = __float_as_uint(a.w);
            uint32_t li = local_index(p, pixel);
            float4 r = radiance[li];
            r.x += c.x; r.y += c.y; r.z += c.z;
            radiance[li] = r;
        }
    }
    warp_count(&ctr->n_unoccluded[bounce], un);
}

// Fused IntersectShadowRays + AccumulateDirectSamples: persistent warps drain the bounce's shadow-ray queue through a
// global atomic cursor (work_shadow[bounce]), 32 consecutive rays per grab.
template <bool COUNT, int SMEM>
__device__ __forceinline__ void shadow_phase(const FrameParams& p, const DevScene& sc, int mode, const Queues& q, DevCounters* ctr, float4* radiance,
                                             uint32_t bounce, const float4* s_bvh)
{
    const uint32_t n = ctr->emit[bounce].shadow;
    const int lane = threadIdx.x & 31;
    uint32_t nv = 0, nt = 0;
    for (;;)
    {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctr->work_shadow[bounce], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
      {
        uint32_t i = base + lane;
        bool un = false;
        if (i < n)
        {
            float4 a = q.sA[i], b = q.sB[i];
            float bu, bv, bt;
            if (SMEM == 1) un = trace_fast<true, false, 1>(sc, s_bvh, s_bvh + sc.wnodes_f4, mk3(a), mk3(b), 0.0f, b.w, bu, bv, bt, nv, nt) == RT_INVALID_ID;
            else if (SMEM == 2) un = trace_fast<true, false, 2>(sc, s_bvh, sc.wtris, mk3(a), mk3(b), 0.0f, b.w, bu, bv, bt, nv, nt) == RT_INVALID_ID;
            else un = trace<true, COUNT>(sc, mode, mk3(a), mk3(b), 0.0f, b.w, bu, bv, bt, nv, nt) == RT_INVALID_ID;
            if (un)
            {
                float4 c = q.sC[i];
                uint32_t pixel = __float_as_uint(a.w);
                uint32_t li = local_index(p, pixel);
                float4 r = radiance[li];
                r.x += c.x; r.y += c.y; r.z += c.z;
                radiance[li] = r;
            }
        }
        warp_count(&ctr->n_unoccluded[bounce], un);
      }
    }
    if (COUNT) { warp_sum64(&ctr->nodes_shadow[bounce], nv); warp_sum64(&ctr->tris_shadow[bounce], nt); }
}

template <bool COUNT, int SMEM>
__global__ void __launch_bounds__(256, RT_MINB_TRACE) k_shadow_accumulate(FrameParams p, DevScene sc, int mode, Queues q, DevCounters* ctr, float4* radiance, uint32_t bounce)
{
    extern __shared__ __align__(128) float4 s_bvh[];
    __shared__ uint64_t s_mbar;
    if (SMEM == 1) tma_stage_bvh(s_bvh, sc, &s_mbar);
    if (SMEM == 2) tma_stage_top(s_bvh, sc.wnodes, sc.top_k, &s_mbar);
    pdl_wait(); pdl_launch_dependents();
    shadow_phase<COUNT, SMEM>(p, sc, mode, q, ctr, radiance, bounce, s_bvh);
}

// ---- queue-split schedule (default): hit/miss stream compaction between traversal and shading --------------
// Traversal and shading are separate persistent kernels: the traversal kernel stays small in registers
// (more resident warps to hide the dependent node fetches), and it compacts its results into a HIT queue and a
// MISS queue with one warp-aggregated atomic each (__ballot + __popc), so that the shading kernel runs full
// warps of hits (Lambert/GGX + NEE) and full warps of misses (environment lookup) instead of warps that
// mix the two and idle through each other's code.  Both kernels drain their queues through a global atomic
// cursor, 32 entries per grab.
template <bool COUNT, int SMEM>
__device__ __forceinline__ void closest_phase(const FrameParams& p, const DevScene& sc, int mode, const Queues& q, DevCounters* ctr, uint32_t bounce,
                                              const float4* s_bvh)
{
    const uint32_t n = *in_count_ptr(ctr, bounce);
    const int in = bounce & 1;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t nv = 0, nt = 0;
    // 32 rays per cursor grab.  Measured alternatives that were slower: 64 rays per grab (-4 %), and issuing the next
    // round's grab before the current round's work to hide the atomic's latency (-5 %): both make a warp own more
    // work at a time, and the coarser tail costs more than the hidden latency saves.
    for (;;)
    {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctr->work_ext[bounce], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        {
            uint32_t i = base + lane;
            bool live = i < n, hit = false;
            float bu = 0.0f, bv = 0.0f, bt = 0.0f;
            uint32_t prim = RT_INVALID_ID;
            if (live)
            {
                float4 a = q.A[in][i], b = q.B[in][i];
                if (SMEM == 1) prim = trace_fast<false, false, 1>(sc, s_bvh, s_bvh + sc.wnodes_f4, mk3(a), mk3(b), 0.0f, b.w, bu, bv, bt, nv, nt);
                else if (SMEM == 2) prim = trace_fast<false, false, 2>(sc, s_bvh, sc.wtris, mk3(a), mk3(b), 0.0f, b.w, bu, bv, bt, nv, nt);
                else prim = trace<false, COUNT>(sc, mode, mk3(a), mk3(b), 0.0f, b.w, bu, bv, bt, nv, nt);
                hit = prim != RT_INVALID_ID;
            }
            const unsigned hmask = __ballot_sync(0xffffffffu, hit);
            const unsigned mmask = __ballot_sync(0xffffffffu, live && !hit);
            unsigned long long slot = 0ull;
            if (lane == 0)
                slot = atomicAdd((unsigned long long*)&ctr->hm[bounce], (unsigned long long)__popc(hmask) | ((unsigned long long)__popc(mmask) << 32));
            slot = __shfl_sync(0xffffffffu, slot, 0);
#ifdef RT_HITQ_CARRY
            if (hit)
            {
                const uint32_t k = (uint32_t)slot + __popc(hmask & lt_mask);
                q.hitq[k] = make_float4(bu, bv, __uint_as_float(prim), __uint_as_float(i));
                q.hA[k] = q.A[in][i]; q.hB[k] = q.B[in][i]; q.hC[k] = q.C[in][i];
            }
#else
            if (hit) q.hitq[(uint32_t)slot + __popc(hmask & lt_mask)] = make_float4(bu, bv, __uint_as_float(prim), __uint_as_float(i));
#endif
            else if (live) q.missq[(uint32_t)(slot >> 32) + __popc(mmask & lt_mask)] = i;
        }
    }
    if (COUNT) { warp_sum64(&ctr->nodes_ext[bounce], nv); warp_sum64(&ctr->tris_ext[bounce], nt); }
}

template <bool COUNT, int SMEM>
__global__ void __launch_bounds__(256, RT_MINB_TRACE) k_trace_closest(FrameParams p, DevScene sc, int mode, Queues q, DevCounters* ctr, uint32_t bounce)
{
    extern __shared__ __align__(128) float4 s_bvh[];
    __shared__ uint64_t s_mbar;
    if (SMEM == 1) tma_stage_bvh(s_bvh, sc, &s_mbar);
    if (SMEM == 2) tma_stage_top(s_bvh, sc.wnodes, sc.top_k, &s_mbar);
    pdl_wait(); pdl_launch_dependents();
    closest_phase<COUNT, SMEM>(p, sc, mode, q, ctr, bounce, s_bvh);
}

// Closest-hit traversal of bounce b and the shadow pass of bounce b-1 in ONE persistent kernel (RT_OPT_OVERLAP = 2, the
// default): the two are independent (the shadow pass only shares the radiance buffer with LATER shading passes), every
// warp drains the extension queue and then the shadow queue, so the short shadow rays fill the tail of the long
// extension rays without a second launch, a second staging of the BVH or cross-stream events.
template <int SMEM>
__global__ void __launch_bounds__(256, RT_MINB_TRACE) k_trace_both(FrameParams p, DevScene sc, int mode, Queues q, DevCounters* ctr, float4* radiance,
                                                                   uint32_t bounce, uint32_t shadow_bounce)
{
    extern __shared__ __align__(128) float4 s_bvh[];
    __shared__ uint64_t s_mbar;
    if (SMEM == 1) tma_stage_bvh(s_bvh, sc, &s_mbar);
    if (SMEM == 2) tma_stage_top(s_bvh, sc.wnodes, sc.top_k, &s_mbar);
    pdl_wait(); pdl_launch_dependents();
    closest_phase<false, SMEM>(p, sc, mode, q, ctr, bounce, s_bvh);
    shadow_phase<false, SMEM>(p, sc, mode, q, ctr, radiance, shadow_bounce, s_bvh);
}


// ShadeSurfaceHits over the hit queue, then ShadeMissedRays over the miss queue (independent pixels).
__global__ void __launch_bounds__(256, RT_MINB_SHADE) k_shade_queues(FrameParams p, DevScene sc, Queues q, DevCounters* ctr, float4* radiance, uint32_t bounce, AovParams aov)
{
    pdl_wait(); pdl_launch_dependents();
    const uint32_t n_hit = ctr->hm[bounce].hit, n_miss = ctr->hm[bounce].miss;
    const uint32_t hit_span = (n_hit + 31u) & ~31u;            // warps never mix hits and misses
    const uint32_t total = hit_span + n_miss;
    const int in = bounce & 1;
    const int lane = threadIdx.x & 31;
    for (;;)
    {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctr->work_shade[bounce], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        if (base < hit_span)
        {
            uint32_t k = base + lane;
            bool hit = k < n_hit;
            uint32_t pixel = 0;
            ShadeOut so;
            so.emissive = so.spawn_next = so.spawn_shadow = false;
            if (hit)
            {
                float4 h = q.hitq[k];
#ifdef RT_HITQ_CARRY
                float4 a = q.hA[k], b = q.hB[k], c = q.hC[k];
#else
                uint32_t i = __float_as_uint(h.w);
                float4 a = q.A[in][i], b = q.B[in][i], c = q.C[in][i];
#endif
                pixel = __float_as_uint(a.w);
                shade_hit(sc, p, aov, bounce, pixel, mk3(a), mk3(b), mk3(c), __float_as_uint(h.z), h.x, h.y, so);
            }
            emit_rays(p, q, ctr, radiance, bounce, pixel, hit, so);
        }
        else
        {
            uint32_t k = base - hit_span + lane;
            if (k < n_miss)
            {
                uint32_t i = q.missq[k];
                float4 a = q.A[in][i], b = q.B[in][i], c = q.C[in][i];
                shade_miss(sc, p, radiance, __float_as_uint(a.w), mk3(b), mk3(c));
            }
        }
    }
}

#include "rt_frame_kernel.cuh"


// resolve_radiance.cl:31-86: shaded colour (radiance / sample_count unless the denoiser is on, then Reinhard x/(1+x))
// or one of the AOV views
__global__ void __launch_bounds__(256) k_resolve(const float4* radiance, float4* out, uint32_t n, uint32_t sample_count, int denoiser,
                                                 int aov_index, AovParams aov)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (aov_index == 1) { float4 a = aov.albedo[i]; out[i] = make_float4(a.x, a.y, a.z, 1.0f); }
    else if (aov_index == 2) { float d = aov.depth[i] * 0.1f; out[i] = make_float4(d, d, d, 1.0f); }
    else if (aov_index == 3) { float4 nn = aov.normal[i]; out[i] = make_float4(nn.x * 0.5f + 0.5f, nn.y * 0.5f + 0.5f, nn.z * 0.5f + 0.5f, 1.0f); }
    else if (aov_index == 4) { float2 v = aov.velocity[i]; out[i] = make_float4(v.x, v.y, 0.0f, 1.0f); }
    else
    {
        float4 r = radiance[i];
        f3 hdr = mk3(r);
        if (!denoiser) hdr = hdr / (float)sample_count;
        f3 ldr = hdr / (mk3(hdr.x + 1.0f, hdr.y + 1.0f, hdr.z + 1.0f));
        out[i] = make_float4(ldr.x, ldr.y, ldr.z, 1.0f);
    }
}

// ResolveRadiance of a gathered multi-GPU frame: pixel (x, y) lives in slab y % world at local row y / world
__global__ void __launch_bounds__(256) k_resolve_gathered(const float4* slabs, size_t stride_f4, float4* out, uint32_t width, uint32_t height,
                                                          uint32_t world, uint32_t sample_count)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width * height) return;
    uint32_t y = i / width, x = i - y * width;
    float4 r = slabs[(size_t)(y % world) * stride_f4 + (size_t)(y / world) * width + x];
    f3 hdr = mk3(r) / (float)sample_count;
    f3 ldr = hdr / (mk3(hdr.x + 1.0f, hdr.y + 1.0f, hdr.z + 1.0f));
    out[i] = make_float4(ldr.x, ldr.y, ldr.z, 1.0f);
}

// TemporalAccumulation, denoiser.cl:27-79 (single-GPU only: the reprojected pixel may be any pixel of the image)
__global__ void __launch_bounds__(256) k_temporal(uint32_t width, uint32_t height, float4* radiance, const float4* prev_radiance,
                                                  const float* depth, const float* prev_depth, const float2* velocity)
{
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= width * height) return;
    int x = (int)(idx % width), y = (int)(idx / width);
    float dv = depth[idx];
    if (dv == RT_MAX_RENDER_DIST) return;
    float2 mv = velocity[idx];
    float pu = ((float)x + 0.5f) / (float)width - mv.x;
    float pv = ((float)y + 0.5f) / (float)height - mv.y;
    float fx = pu * (float)width, fy = pv * (float)height;
    if (!(fx == fx) || !(fy == fy) || fabsf(fx) > 1.0e9f || fabsf(fy) > 1.0e9f) return;      // NaN / overflow: out of range
    int px = (int)fx, py = (int)fy;
    if (px < 0 || px >= (int)width || py < 0 || py >= (int)height) return;
    int pidx = py * (int)width + px;
    float pd = prev_depth[pidx];
    if (fabsf(dv - pd) / dv > 0.1f) return;
    float4 cur = radiance[idx], prev = prev_radiance[pidx];
    cur.x = cur.x + (prev.x - cur.x) * 0.9f; cur.y = cur.y + (prev.y - cur.y) * 0.9f; cur.z = cur.z + (prev.z - cur.z) * 0.9f;
    radiance[idx] = cur;
}

// Completion flags of the fused gather (one uint32 per rank, in the presenting device's memory, behind the slabs): a rank stores its
// frame number after its frame kernel has completed (stream order) — everything it pushed is then visible system-wide — and the
// presenting rank's stream waits until every flag has reached its own frame number.
__global__ void k_gather_signal(uint32_t* flag, uint32_t frame)
{
    __threadfence_system();
    *(volatile uint32_t*)flag = frame;
    __threadfence_system();
}
__global__ void k_gather_wait(const uint32_t* flags, uint32_t world, uint32_t frame)
{
    if (threadIdx.x < world)
        while ((int32_t)(*(volatile const uint32_t*)(flags + threadIdx.x) - frame) < 0) __nanosleep(200);
    __threadfence_system();
}

// rt_math_eval: the functions of include/rt_math.h evaluated on the device (parity tap: tests compare them bitwise with the
// same header compiled for the host)
__global__ void k_math_eval(int fn, const float* a, const float* b, float* out, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = a[i], y = b[i];
    float r = 0.0f;
    switch (fn)
    {
    case 0: r = rt_sinf(x); break;
    case 1: r = rt_cosf(x); break;
    case 2: r = rt_tanf(x); break;
    case 3: r = rt_atan2f(x, y); break;
    case 4: r = rt_acosf(x); break;
    case 5: r = rt_powf(x, y); break;
    case 6: r = rt_fminf(x, y); break;
    case 7: r = rt_fmaxf(x, y); break;
    case 8: r = 1.0f / sqrtf(x); break;          // the normalize() building blocks: IEEE sqrt and division
    case 9: r = x / y; break;
    }
    out[i] = r;
}

__global__ void k_unpack_rays(const float4* A, const float4* B, const uint32_t* count, uint32_t width, RtRay* rays, uint32_t* pixels)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    float4 a = A[i], b = B[i];
    rays[i].origin = RtFloat3{ a.x, a.y, a.z, 0.0f };
    rays[i].direction = RtFloat3{ b.x, b.y, b.z, b.w };
    uint32_t pxy = __float_as_uint(a.w);
    pixels[i] = (pxy >> 16) * width + (pxy & 0xFFFFu);
}

} // namespace

// ====================================================================================== host side
struct rt_ctx
{
    int device = 0;
    // A multi-device context (rt_create_multi) owns no device memory itself: it holds one child context per device (child i
    // renders rank i of an n-way scanline partition) and every call on it fans out to the children from the caller's thread.
    std::vector<rt_ctx*> children;
    int present = 0;                               // RT_OPT_PRESENT (groups): 0 parallel read-back, 1 gather to device 0 over NVLink
    float4* gather_buf = nullptr; size_t gather_stride_f4 = 0;   // device 0: the gathered radiance slabs (rt_gather_radiance)
    std::vector<cudaEvent_t> gather_events;
    uint32_t width = 0, height = 0, rank = 0, world = 1, n_local = 0, local_rows = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    std::string error;

    // options
    int white_furnace = 0, sampler = 0, aov = 0, denoiser = 0, count_traversal = 0, kernel_timing = 0, traversal = 1, smem_bvh = 1;
    uint32_t top_smem_records = 0;                 // RT_OPT_TOP_SMEM
    int bvh_depth = 0;

    // per-pixel buffers
    Queues q = {};
    float4* radiance = nullptr;
    float4* resolved = nullptr;
    // whole-frame CUDA graph (RT_OPT_GRAPH): rt_integrate replays the captured launch sequence; only k_set_frame's
    // by-value argument (sample index, camera constants) is updated per frame
    int use_graph = 1;
    uint64_t config_gen = 1, graph_gen = 0;     // config_gen changes whenever a launch argument other than FrameDyn may change
    uint32_t graph_max_bounces = 0;
    cudaGraph_t graph = nullptr;                // kept alive: node handles used for parameter updates belong to it
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraphNode_t graph_set_frame_node = nullptr;
    void* graph_set_frame_func = nullptr;
    uint64_t graph_launches = 0;
    // shadow pass on a second stream (RT_OPT_OVERLAP): k_shadow_accumulate(b) runs concurrently with k_trace_closest(b+1)
    int overlap = 2;               // RT_OPT_OVERLAP
    bool shadow_deferred = false;  // overlap 2: rt_shadow_accumulate(shadow_deferred_bounce) has been requested but not launched yet
    uint32_t shadow_deferred_bounce = 0;
    cudaStream_t shadow_stream = nullptr;
    cudaEvent_t ev_shaded = nullptr, ev_shadowed = nullptr;
    bool shadow_pending = false;
    // pipelined read-back (rt_resolve_async): second resolve buffer, copy stream, events
    float4* resolved2 = nullptr;
    float4* resolved_full[2] = { nullptr, nullptr };   // whole-frame resolve targets of rt_resolve_gathered
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t resolve_done[2] = { nullptr, nullptr }, copy_done[2] = { nullptr, nullptr };
    bool copy_pending[2] = { false, false };
    uint32_t async_index = 0;
    // AOV + denoiser buffers (allocated on first use)
    float4* aov_albedo = nullptr; float* aov_depth = nullptr; float4* aov_normal = nullptr; float2* aov_velocity = nullptr;
    float4* prev_radiance = nullptr; float* prev_depth = nullptr;
    int aov_always = 0;
    RtCamera prev_camera = {}, aov_prev_camera = {};
    DevCounters* counters = nullptr;
    FrameDyn* d_dyn = nullptr;
    bool pdl = true;               // RT_OPT_PDL
    int frame_kernel = 2;          // RT_OPT_FRAME_KERNEL: 1 rt_integrate launches ONE persistent kernel per frame (k_frame), 0 one kernel per phase, 2 by partition size
    // fused gather (rt_set_gather_target): this rank's slab and flag in the presenting device's memory, frames signalled so far
    float4* gather_slab = nullptr; uint32_t* gather_flag = nullptr; uint32_t gather_frame = 0;
    uint32_t gather_sample = 0xFFFFFFFFu;          // sample_count after the last frame that was pushed
    void* own_gather = nullptr; size_t own_gather_stride = 0;      // the buffer this context allocated (rt_gather_buffer)
    std::vector<void*> ipc_opened;
    int grid_div = 1;              // contexts that share this device (rt_create_multi lists a device k times): each launches 1/k of the resident CTAs
    int frame_threads = 0;         // RT_OPT_FRAME_THREADS: threads per CTA of k_frame (0 = by partition size)
    size_t n_alloc = 0;            // entries of every per-pixel queue: n_local + one group of 32 per CTA that can be resident (k_frame's regions)
    int* d_bn = nullptr;           // sobol | scrambling | ranking (rt_upload_sampler_tables)
    struct Occupancy { const void* kernel; size_t smem; int per_sm; };
    std::vector<Occupancy> occupancy;   // resident CTAs per SM of each persistent kernel (persistent_grid)
    void* scratch = nullptr; size_t scratch_bytes = 0;

    // scene
    bool scene_ready = false, camera_ready = false;
    DevScene scene = {};
    std::vector<void*> scene_allocs;
    RtCamera camera = {};
    RayGenConsts raygen = {};

    uint32_t sample_count = 0;
    uint32_t cur_bounce = 0;
    bool frame_started = false;
    uint64_t launches = 0;

    // kernel timing
    struct Timed { cudaEvent_t a, b; int cls; };
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> event_pool;
    float ms[RT_K_CLASS_COUNT] = {};
    uint32_t nlaunch[RT_K_CLASS_COUNT] = {};
};

static std::string g_create_error;

#define RT_FAIL(ctx, code, ...)                                   \
    do {                                                          \
        char rt_buf_[512];                                        \
        snprintf(rt_buf_, sizeof(rt_buf_), __VA_ARGS__);          \
        (ctx)->error = rt_buf_;                                   \
        return (code);                                            \
    } while (0)

#define RT_CUDA(ctx, expr)                                                                         \
    do {                                                                                           \
        cudaError_t rt_e_ = (expr);                                                                \
        if (rt_e_ != cudaSuccess)                                                                  \
            RT_FAIL(ctx, RT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(rt_e_), __FILE__, __LINE__); \
    } while (0)

#define RT_CHECK_CTX(ctx) do { if (!(ctx)) return RT_ERR_INVALID_ARGUMENT; } while (0)

// Multi-device contexts: forward the call to every child (variable `k`), in rank order, from the caller's thread.  Kernel
// launches and copies are asynchronous, so the devices work concurrently; the first failing child's message is kept.
#define RT_FANOUT(ctx, call)                                                             \
    do {                                                                                 \
        if (!(ctx)->children.empty())                                                    \
        {                                                                                \
            for (rt_ctx* k : (ctx)->children)                                            \
            {                                                                            \
                int rt_frc_ = (call);                                                    \
                if (rt_frc_ != RT_OK) { (ctx)->error = k->error; return rt_frc_; }       \
            }                                                                            \
            return RT_OK;                                                                \
        }                                                                                \
    } while (0)
#define RT_NOT_ON_GROUP(ctx, what)                                                       \
    do { if (!(ctx)->children.empty()) RT_FAIL(ctx, RT_ERR_UNSUPPORTED, what " is a single-device call: use it on a context of rt_create"); } while (0)

struct rt_ctx;
extern "C" {   // defined with rt_shadow_accumulate
static int launch_shadow_pass(rt_ctx* c, uint32_t bounce, cudaStream_t st);
}

namespace
{

struct TimedLaunch
{
    rt_ctx* c; int cls; cudaEvent_t a = nullptr, b = nullptr; cudaStream_t st;
    TimedLaunch(rt_ctx* ctx, int k, cudaStream_t stream = nullptr) : c(ctx), cls(k), st(stream ? stream : ctx->stream)
    {
        ++c->launches;
        if (!c->kernel_timing) return;
        auto get = [&]() { cudaEvent_t e; if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); } else cudaEventCreate(&e); return e; };
        a = get(); b = get();
        cudaEventRecord(a, st);
    }
    ~TimedLaunch()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        c->timed.push_back({ a, b, cls });
    }
};

inline dim3 grid_for(uint32_t n, uint32_t block = 256) { uint32_t g = (n + block - 1) / block; return dim3(g ? g : 1u); }   // n = 0: one idle block

// Work submitted to the shadow stream must be ordered before anything on the render stream that touches the radiance
// buffer, the shadow queue or the counters again.
int join_shadow(rt_ctx* c)
{
    // callers reach this before their own cudaSetDevice; with several devices in one process (rt_create_multi) the current device
    // may be another context's
    if (c->shadow_deferred || c->shadow_pending) RT_CUDA(c, cudaSetDevice(c->device));
    if (c->shadow_deferred)
    {   // nobody merged the deferred shadow pass into a traversal kernel: it runs on its own, in stream order
        c->shadow_deferred = false;
        int rc = launch_shadow_pass(c, c->shadow_deferred_bounce, c->stream); if (rc) return rc;
    }
    if (c->shadow_pending)
    {
        RT_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_shadowed, 0));
        c->shadow_pending = false;
    }
    return RT_OK;
}

int post_launch(rt_ctx* c, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) RT_FAIL(c, RT_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return RT_OK;
}

FrameParams frame_params(const rt_ctx* c)
{
    FrameParams p;
    p.width = c->width; p.height = c->height; p.rank = c->rank; p.world = c->world; p.n_local = c->n_local;
    p.white_furnace = c->white_furnace; p.bn = c->sampler ? c->d_bn : nullptr; p.dyn = c->d_dyn;
    p.gather = nullptr;            // set by the one-kernel frame's launch only
    return p;
}

int require_ready(rt_ctx* c)
{
    if (!c->scene_ready) RT_FAIL(c, RT_ERR_NOT_READY, "rt_upload_scene has not been called");
    if (!c->camera_ready) RT_FAIL(c, RT_ERR_NOT_READY, "rt_set_camera has not been called");
    return RT_OK;
}

void free_aov_buffers(rt_ctx* c);
bool aov_wanted(const rt_ctx* c);
int ensure_aov_buffers(rt_ctx* c);

int alloc_frame_buffers(rt_ctx* c)
{
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->shadow_stream) cudaStreamSynchronize(c->shadow_stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream); // a pipelined read-back may still be reading the resolve buffers
    c->shadow_pending = false; c->shadow_deferred = false;     // the frame in flight (if any) is abandoned with its buffers
    ++c->config_gen;
    auto freep = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    for (int i = 0; i < 2; ++i) { freep(c->q.A[i]); freep(c->q.B[i]); freep(c->q.C[i]); }
    freep(c->q.sA); freep(c->q.sB); freep(c->q.sC); freep(c->q.hits); freep(c->q.shadow_flags);
    freep(c->q.hitq); freep(c->q.missq);
    freep(c->radiance); freep(c->resolved); freep(c->resolved2);
    free_aov_buffers(c);
    c->local_rows = (c->height > c->rank) ? (c->height - c->rank + c->world - 1) / c->world : 0;
    c->n_local = c->local_rows * c->width;
    // k_frame gives every resident CTA its own region of each queue, rounded up to whole groups of 32 slots
    size_t n = (size_t)c->n_local + 32u + (size_t)c->num_sms * 32u * 32u;       // up to 32 resident CTAs per SM
    c->n_alloc = n;
    for (int i = 0; i < 2; ++i)
    {
        RT_CUDA(c, cudaMalloc(&c->q.A[i], n * 16)); RT_CUDA(c, cudaMalloc(&c->q.B[i], n * 16)); RT_CUDA(c, cudaMalloc(&c->q.C[i], n * 16));
    }
    RT_CUDA(c, cudaMalloc(&c->q.sA, n * 16)); RT_CUDA(c, cudaMalloc(&c->q.sB, n * 16)); RT_CUDA(c, cudaMalloc(&c->q.sC, n * 16));
    RT_CUDA(c, cudaMalloc(&c->q.hits, n * 16)); RT_CUDA(c, cudaMalloc(&c->q.shadow_flags, n * 4));
    RT_CUDA(c, cudaMalloc(&c->q.hitq, n * 16)); RT_CUDA(c, cudaMalloc(&c->q.missq, n * 4));
#ifdef RT_HITQ_CARRY
    RT_CUDA(c, cudaMalloc(&c->q.hA, n * 16)); RT_CUDA(c, cudaMalloc(&c->q.hB, n * 16)); RT_CUDA(c, cudaMalloc(&c->q.hC, n * 16));
#endif
    RT_CUDA(c, cudaMalloc(&c->radiance, n * 16)); RT_CUDA(c, cudaMalloc(&c->resolved, n * 16));
    RT_CUDA(c, cudaMemsetAsync(c->radiance, 0, n * 16, c->stream));
    c->copy_pending[0] = c->copy_pending[1] = false;          // both streams were drained above
    if (aov_wanted(c)) return ensure_aov_buffers(c);           // a selected view / the denoiser keeps its buffers across a re-partition
    return RT_OK;
}

// Grid of a persistent kernel: exactly the CTAs that are resident at once (occupancy x SMs) — CTAs of a second wave
// would only start, find the cursor exhausted and exit — and no more CTAs than there can be work for (a rank of a
// multi-GPU partition owns few pixels).  Measured: -1.4 % frame time on one GPU, -5 % on a 1/8 partition.
int persistent_grid(rt_ctx* c, const void* kernel, size_t dyn_smem)
{
    int per_sm = 0;
    for (auto& e : c->occupancy) if (e.kernel == kernel && e.smem == dyn_smem) { per_sm = e.per_sm; break; }
    if (!per_sm)
    {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        c->occupancy.push_back({ kernel, dyn_smem, per_sm });
    }
    int g = c->num_sms * per_sm / c->grid_div;
    if (g < c->num_sms) g = c->num_sms;
    int need = (int)((c->n_local + 255u) / 256u);
    if (need < 1) need = 1;
    return g < need ? g : need;
}
#define RT_PGRID(c, kern, smem) persistent_grid(c, (const void*)(kern), smem)

// Can the traversal kernel of the next bounce also run a deferred shadow pass (k_trace_both)?
bool merged_trace_available(const rt_ctx* c) { return !c->count_traversal && !c->kernel_timing; }

// <<<grid, 256, smem, stream>>> with the programmatic-dependent-launch attribute when RT_OPT_PDL is on
template <class... KArgs, class... Args>
cudaError_t launch_chain(rt_ctx* c, void (*kernel)(KArgs...), int grid, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    if (c->pdl && !c->kernel_timing) { cfg.attrs = &attr; cfg.numAttrs = 1; }
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// How a traversal kernel stages the BVH (TMA bulk copies into dynamic shared memory), decided per launch:
//   mode 1  the whole structure (interior records + triangle records) when it fits 40 KB (5 resident CTAs x 40 KB stay under
//           the 227 KB of an SM): CornellBox;
//   mode 2  only the top of the tree — the first top_k interior records in breadth-first order (RT_OPT_TOP_SMEM records,
//           default 0 = off, see DESIGN.md for the measurements) — when the structure does not fit;
//   mode 0  nothing staged: literal traversal (RT_OPT_TRAVERSAL 0), RT_OPT_SMEM_BVH off.
struct Stage { int mode; size_t bytes; uint32_t top_k; };
Stage bvh_stage(const rt_ctx* c)
{
    Stage st = { 0, 0, 0 };
    if (!c->smem_bvh || c->traversal != 1) return st;
    size_t bytes = ((size_t)c->scene.wnodes_f4 + c->scene.wtris_f4) * 16;
    if (bytes <= 40 * 1024) { st.mode = 1; st.bytes = bytes; return st; }
    uint32_t k = c->top_smem_records < c->scene.top_n ? c->top_smem_records : c->scene.top_n;
    if (k >= 8) { st.mode = 2; st.bytes = (size_t)k * 64; st.top_k = k; }
    return st;
}
size_t smem_stage_bytes(const rt_ctx* c) { return bvh_stage(c).bytes; }
DevScene staged_scene(const rt_ctx* c, const Stage& st) { DevScene sc = c->scene; sc.top_k = st.top_k; sc.stack_off = (uint32_t)((st.bytes + 127) & ~(size_t)127); return sc; }
// kernels that ask for more than 48 KB of dynamic shared memory need the limit raised first
void set_smem_limit(const void* kernel, size_t bytes) { if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); }
#ifdef RT_SMEM_STACK
// experiment: dynamic shared memory of the per-phase traversal kernels = staged BVH + one stack column per thread
size_t with_stack(const rt_ctx* c, size_t stage) { return ((stage + 127) & ~(size_t)127) + (size_t)c->bvh_depth * 256 * 8; }
#else
size_t with_stack(const rt_ctx*, size_t stage) { return stage; }
#endif

AovCam aov_cam(const RtCamera& cam);

FrameDyn frame_dyn(const rt_ctx* c)
{
    FrameDyn dyn;
    memset(&dyn, 0, sizeof(dyn));
    dyn.sample_idx = c->sample_count; dyn.raygen = c->raygen;
    dyn.cam = aov_cam(c->camera); dyn.prev = aov_cam(c->aov_prev_camera);
    return dyn;
}

bool aov_wanted(const rt_ctx* c) { return c->aov != 0 || c->denoiser != 0 || c->aov_always != 0; }

AovCam aov_cam(const RtCamera& cam)
{
    AovCam a;
    a.position = f3{ cam.position.x, cam.position.y, cam.position.z };
    a.front = f3{ cam.front.x, cam.front.y, cam.front.z };
    a.up = f3{ cam.up.x, cam.up.y, cam.up.z };
    a.right = f3{ a.front.y * a.up.z - a.front.z * a.up.y, a.front.z * a.up.x - a.front.x * a.up.z, a.front.x * a.up.y - a.front.y * a.up.x };
    a.angle = rt_tanf(0.5f * cam.fov);          // aov.cl:35
    a.aspect_ratio = cam.aspect_ratio;
    return a;
}

AovParams aov_params(const rt_ctx* c)
{
    AovParams a;
    memset(&a, 0, sizeof(a));
    a.enabled = (aov_wanted(c) && c->aov_albedo) ? 1 : 0;
    a.albedo = c->aov_albedo; a.depth = c->aov_depth; a.normal = c->aov_normal; a.velocity = c->aov_velocity;
    return a;
}

int ensure_aov_buffers(rt_ctx* c)
{
    if (c->aov_albedo) return RT_OK;
    ++c->config_gen;
    size_t n = c->n_local ? c->n_local : 1;
    RT_CUDA(c, cudaMalloc(&c->aov_albedo, n * 16)); RT_CUDA(c, cudaMalloc(&c->aov_depth, n * 4));
    RT_CUDA(c, cudaMalloc(&c->aov_normal, n * 16)); RT_CUDA(c, cudaMalloc(&c->aov_velocity, n * 8));
    RT_CUDA(c, cudaMalloc(&c->prev_radiance, n * 16)); RT_CUDA(c, cudaMalloc(&c->prev_depth, n * 4));
    RT_CUDA(c, cudaMemsetAsync(c->aov_albedo, 0, n * 16, c->stream)); RT_CUDA(c, cudaMemsetAsync(c->aov_depth, 0, n * 4, c->stream));
    RT_CUDA(c, cudaMemsetAsync(c->aov_normal, 0, n * 16, c->stream)); RT_CUDA(c, cudaMemsetAsync(c->aov_velocity, 0, n * 8, c->stream));
    RT_CUDA(c, cudaMemsetAsync(c->prev_radiance, 0, n * 16, c->stream)); RT_CUDA(c, cudaMemsetAsync(c->prev_depth, 0, n * 4, c->stream));
    return RT_OK;
}

void free_aov_buffers(rt_ctx* c)
{
    cudaFree(c->aov_albedo); cudaFree(c->aov_depth); cudaFree(c->aov_normal); cudaFree(c->aov_velocity);
    cudaFree(c->prev_radiance); cudaFree(c->prev_depth);
    c->aov_albedo = nullptr; c->aov_depth = nullptr; c->aov_normal = nullptr; c->aov_velocity = nullptr;
    c->prev_radiance = nullptr; c->prev_depth = nullptr;
}   // 8 CTAs x 256 threads = full occupancy target per SM

} // namespace

extern "C" {

const char* rt_last_error(const rt_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int rt_create(uint32_t width, uint32_t height, int device, rt_ctx** out_ctx)
{
    if (!out_ctx || width == 0 || height == 0 || width > 65535u || height > 65535u)
    {
        g_create_error = "rt_create: bad arguments (need 1 <= width, height <= 65535)"; return RT_ERR_INVALID_ARGUMENT;
    }
    *out_ctx = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
    {
        g_create_error = std::string("rt_create: no CUDA device (") + cudaGetErrorString(e) + "); this backend has no CPU fallback";
        return RT_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) { g_create_error = "rt_create: device index out of range"; return RT_ERR_INVALID_ARGUMENT; }
    rt_ctx* c = new rt_ctx;
    c->device = device; c->width = width; c->height = height;
    auto fail = [&](const char* what, cudaError_t err) { g_create_error = std::string("rt_create: ") + what + ": " + cudaGetErrorString(err); delete c; return RT_ERR_CUDA; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
    c->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaMalloc(&c->counters, sizeof(DevCounters))) != cudaSuccess) return fail("cudaMalloc(counters)", e);
    if ((e = cudaMalloc(&c->d_dyn, sizeof(FrameDyn))) != cudaSuccess) return fail("cudaMalloc(frame constants)", e);
    cudaMemsetAsync(c->d_dyn, 0, sizeof(FrameDyn), c->stream);
    cudaMemsetAsync(c->counters, 0, sizeof(DevCounters), c->stream);
    int rc = alloc_frame_buffers(c);
    if (rc != RT_OK) { g_create_error = c->error; rt_destroy(c); return rc; }
    *out_ctx = c;
    return RT_OK;
}

int rt_destroy(rt_ctx* c)
{
    if (!c) return RT_ERR_INVALID_ARGUMENT;
    if (!c->children.empty())
    {   // multi-device context: the gather buffer lives on the first child's device
        cudaSetDevice(c->children[0]->device);
        for (rt_ctx* k : c->children) if (k->stream) cudaStreamSynchronize(k->stream);
        cudaFree(c->gather_buf);
        for (cudaEvent_t e : c->gather_events) cudaEventDestroy(e);
        for (rt_ctx* k : c->children) rt_destroy(k);
        delete c;
        return RT_OK;
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->shadow_stream) cudaStreamSynchronize(c->shadow_stream);
    for (int i = 0; i < 2; ++i) { cudaFree(c->q.A[i]); cudaFree(c->q.B[i]); cudaFree(c->q.C[i]); }
    cudaFree(c->q.sA); cudaFree(c->q.sB); cudaFree(c->q.sC); cudaFree(c->q.hits); cudaFree(c->q.shadow_flags);
    cudaFree(c->q.hitq); cudaFree(c->q.missq);
    cudaFree(c->radiance); cudaFree(c->resolved); cudaFree(c->resolved2); cudaFree(c->counters); cudaFree(c->scratch); cudaFree(c->d_dyn); cudaFree(c->d_bn); cudaFree(c->resolved_full[0]); cudaFree(c->resolved_full[1]);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    if (c->shadow_stream) { cudaStreamSynchronize(c->shadow_stream); cudaStreamDestroy(c->shadow_stream); }
    if (c->ev_shaded) cudaEventDestroy(c->ev_shaded);
    if (c->ev_shadowed) cudaEventDestroy(c->ev_shadowed);
    for (int i = 0; i < 2; ++i) { if (c->resolve_done[i]) cudaEventDestroy(c->resolve_done[i]); if (c->copy_done[i]) cudaEventDestroy(c->copy_done[i]); }
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    free_aov_buffers(c);
    cudaFree(c->own_gather);
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void* p : c->scene_allocs) cudaFree(p);
    for (auto& t : c->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return RT_OK;
}

int rt_set_partition(rt_ctx* c, uint32_t rank, uint32_t world)
{
    RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_set_partition");
    if (world == 0 || rank >= world) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_set_partition: need rank < world, world >= 1");
    if (world > 1 && c->denoiser)
        RT_FAIL(c, RT_ERR_UNSUPPORTED, "the temporal denoiser reprojects across the whole image and is not available with a multi-GPU partition");
    RT_CUDA(c, cudaSetDevice(c->device));
    c->rank = rank; c->world = world;
    c->gather_slab = nullptr; c->gather_flag = nullptr;       // a gather target belongs to the partition it was set for
    return alloc_frame_buffers(c);
}

int rt_upload_scene(rt_ctx* c, const RtSceneDesc* s)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_upload_scene(k, s));
    if (!s) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: null scene");
    // cl_pt_integrator.cpp:385,404 assert non-empty triangles/materials; light.h:46 divides by the light count
    if (!s->triangles || s->n_triangles == 0) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: scene has no triangles");
    if (!s->materials || s->n_materials == 0) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: scene has no materials");
    if (!s->nodes || s->n_nodes == 0) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: scene has no BVH nodes");
    if (!s->lights || s->n_lights == 0 || s->scene_info.analytic_light_count == 0)
        RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: at least one analytic light is required (light.h:46 divides by the count)");
    if (s->scene_info.analytic_light_count > s->n_lights) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: analytic_light_count exceeds n_lights");
    if (!s->env_image || s->env_width == 0 || s->env_height == 0) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: environment image required");
    if (s->n_triangles >= 0x7FFFFFFFull || s->n_nodes >= 0x7FFFFFFFull) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: scene too large");
    for (uint64_t i = 0; i < s->n_triangles; ++i)
        if (s->triangles[i].mtlIndex >= s->n_materials) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: triangle %llu has material index out of range", (unsigned long long)i);

    for (uint64_t i = 0; i < s->n_materials; ++i)
    {   // texture indices of the packed materials must exist (0xFF = none)
        const RtPackedMaterial& m = s->materials[i];
        uint32_t idx[6] = { m.diffuse_albedo >> 24, m.specular_albedo >> 24, (m.roughness_metalness >> 8) & 0xFF, m.roughness_metalness >> 24,
                            (m.ior_emission_idx_transparency >> 8) & 0xFF, m.ior_emission_idx_transparency >> 24 };
        for (uint32_t t : idx)
            if (t != RT_INVALID_TEXTURE_IDX && t >= s->n_textures) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: material %llu references texture %u of %llu", (unsigned long long)i, t, (unsigned long long)s->n_textures);
    }
    for (uint64_t i = 0; i < s->n_textures; ++i)
    {
        const RtTexture& t = s->textures[i];
        if (t.width <= 0 || t.height <= 0 || t.data_start < 0 || (uint64_t)t.data_start + (uint64_t)t.width * t.height > s->n_texture_data)
            RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: texture %llu lies outside the texel array", (unsigned long long)i);
    }

    RT_CUDA(c, cudaSetDevice(c->device));
    { int rc = join_shadow(c); if (rc) return rc; }     // a shadow pass of an unfinished frame may still read the old scene
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    for (void* p : c->scene_allocs) cudaFree(p);
    c->scene_allocs.clear();
    c->scene_ready = false;

    auto upload = [&](const void* src, size_t bytes, const void** dst) -> int {
        void* d = nullptr;
        size_t alloc = bytes ? bytes : 16;
        RT_CUDA(c, cudaMalloc(&d, alloc));
        c->scene_allocs.push_back(d);
        if (bytes) RT_CUDA(c, cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        *dst = d;
        return RT_OK;
    };
    int rc;
    DevScene& ds = c->scene;
    if ((rc = upload(s->nodes, s->n_nodes * sizeof(RtLinearBVHNode), (const void**)&ds.nodes_ref))) return rc;
    {
        std::vector<rtbvh::F4> rec;
        rtbvh::build_tri_shade(s->triangles, s->n_triangles, rec);
        if ((rc = upload(rec.data(), rec.size() * 16, (const void**)&ds.tri_shade))) return rc;
        rtbvh::build_mat_rec(s->materials, s->n_materials, rec);
        if ((rc = upload(rec.data(), rec.size() * 16, (const void**)&ds.mat_rec))) return rc;
        rtbvh::build_light_rec(s->lights, s->n_lights, rec);
        if ((rc = upload(rec.data(), rec.size() * 16, (const void**)&ds.light_rec))) return rc;
    }
    {   // RTTriangle[] derived at upload, cl_pt_integrator.cpp:392-402
        std::vector<RtRTTriangle> rt(s->n_triangles);
        for (uint64_t i = 0; i < s->n_triangles; ++i)
        {
            rt[i].position1 = s->triangles[i].v1.position; rt[i].position2 = s->triangles[i].v2.position; rt[i].position3 = s->triangles[i].v3.position;
        }
        if ((rc = upload(rt.data(), rt.size() * sizeof(RtRTTriangle), (const void**)&ds.tris_ref))) return rc;
    }
    if ((rc = upload(s->materials, s->n_materials * sizeof(RtPackedMaterial), (const void**)&ds.materials))) return rc;
    if ((rc = upload(s->textures, s->n_textures * sizeof(RtTexture), (const void**)&ds.textures))) return rc;
    if ((rc = upload(s->texture_data, s->n_texture_data * 4, (const void**)&ds.texels))) return rc;
    if ((rc = upload(s->env_image, (size_t)s->env_width * s->env_height * 16, (const void**)&ds.env))) return rc;
    ds.env_w = (int)s->env_width; ds.env_h = (int)s->env_height;
    ds.light_count = s->scene_info.analytic_light_count;
    {   // optimised traversal layout
        rtbvh::WideLayout wl;
        std::string err;
        if (!rtbvh::build_layout(s->nodes, s->n_nodes, s->triangles, s->n_triangles, wl, err))
            RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_scene: malformed BVH: %s", err.c_str());
        if ((rc = upload(wl.nodes.data(), wl.nodes.size() * 16, (const void**)&ds.wnodes))) return rc;
        if ((rc = upload(wl.tris.data(), wl.tris.size() * 16, (const void**)&ds.wtris))) return rc;
        ds.root_ref = wl.root_ref;
        ds.wnodes_f4 = (uint32_t)wl.nodes.size(); ds.wtris_f4 = (uint32_t)wl.tris.size();
        ds.top_n = wl.top_n; ds.top_k = 0; ds.stack_off = 0;
        c->bvh_depth = wl.max_depth;
    }
    c->scene_ready = true;
    ++c->config_gen;
    return RT_OK;
}

int rt_set_camera(rt_ctx* c, const RtCamera* cam)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_set_camera(k, cam));
    if (!cam) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_set_camera: null camera");
    c->camera = *cam;
    c->aov_prev_camera = c->prev_camera;        // the AOV kernel sees the camera of the previous SetCameraData call
    c->prev_camera = *cam;                      // (cl_pt_integrator.cpp:365-371; zero-initialised before the first call)
    RayGenConsts& r = c->raygen;
    r.position = f3{ cam->position.x, cam->position.y, cam->position.z };
    r.front = f3{ cam->front.x, cam->front.y, cam->front.z };
    r.up = f3{ cam->up.x, cam->up.y, cam->up.z };
    // cross(front, up), raygeneration.cl:112 — plain float ops on the host (no contraction: -fmad=false
    // applies to host code too, and x86-64 has no implicit FMA)
    r.right = f3{ r.front.y * r.up.z - r.front.z * r.up.y, r.front.z * r.up.x - r.front.x * r.up.z, r.front.x * r.up.y - r.front.y * r.up.x };
    r.tan_half_fov = rt_tanf(0.5f * cam->fov);        // raygeneration.cl:108, uniform over the frame
    r.aspect_ratio = cam->aspect_ratio;
    r.aperture = cam->aperture; r.focus_distance = cam->focus_distance;
    r.inv_width = 1.0f / (float)c->width; r.inv_height = 1.0f / (float)c->height;
    c->camera_ready = true;
    return RT_OK;
}

int rt_set_option(rt_ctx* c, int key, uint32_t value)
{
    RT_CHECK_CTX(c);
    if (key == RT_OPT_PRESENT)
    {
        if (value > 1) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "present mode must be 0 (parallel read-back) or 1 (gather to the first device)");
        c->present = (int)value;
        if (!c->children.empty())
        {   // gather presentation: the children push their pixels into a buffer on the first device while they render (fused gather)
            rt_ctx* k0 = c->children[0];
            void* buf = nullptr; uint64_t stride = 0;
            if (value == 1) { int rc = rt_gather_buffer(k0, &buf, &stride, nullptr); if (rc) { c->error = k0->error; return rc; } }
            for (rt_ctx* k : c->children) { int rc = rt_set_gather_target(k, value == 1 ? buf : nullptr, stride); if (rc) { c->error = k->error; return rc; } }
        }
        return RT_OK;
    }
    RT_FANOUT(c, rt_set_option(k, key, value));
    ++c->config_gen;
    switch (key)
    {
    case RT_OPT_WHITE_FURNACE: c->white_furnace = value != 0; return RT_OK;
    case RT_OPT_SAMPLER:
        if (value > 1) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "sampler type must be 0 (kRandom) or 1 (kBlueNoise)");
        if (value == 1 && !c->d_bn)
            RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "kBlueNoise needs the sampler tables: call rt_upload_sampler_tables first");
        c->sampler = (int)value; return RT_OK;
    case RT_OPT_AOV:
        if (value > 4) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "AOV index out of range");
        c->aov = (int)value;
        if (aov_wanted(c)) { RT_CUDA(c, cudaSetDevice(c->device)); return ensure_aov_buffers(c); }
        return RT_OK;
    case RT_OPT_DENOISER:
        if (value != 0 && c->world != 1)
            RT_FAIL(c, RT_ERR_UNSUPPORTED, "the temporal denoiser reprojects across the whole image and is not available with a multi-GPU partition");
        c->denoiser = value != 0;
        if (aov_wanted(c)) { RT_CUDA(c, cudaSetDevice(c->device)); return ensure_aov_buffers(c); }
        return RT_OK;
    case RT_OPT_AOV_ALWAYS:
        c->aov_always = value != 0;
        if (aov_wanted(c)) { RT_CUDA(c, cudaSetDevice(c->device)); return ensure_aov_buffers(c); }
        return RT_OK;
    case RT_OPT_COUNT_TRAVERSAL: c->count_traversal = value != 0; return RT_OK;
    case RT_OPT_KERNEL_TIMING: c->kernel_timing = value != 0; return RT_OK;
    case RT_OPT_GRAPH: c->use_graph = value != 0; return RT_OK;
    case RT_OPT_PDL: c->pdl = value != 0; return RT_OK;
    case RT_OPT_FRAME_THREADS:
        if (value != 0 && (value % 32 != 0 || value > RT_FRAME_MAX_THREADS)) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "frame-kernel CTA size must be 0 (automatic) or a multiple of 32 up to 1024");
        c->frame_threads = (int)value; return RT_OK;
    case RT_OPT_FRAME_KERNEL:
        if (value > 2) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "frame-kernel mode must be 0 (per-phase kernels), 1 (one kernel per frame) or 2 (by partition size)");
        c->frame_kernel = (int)value; return RT_OK;
    case RT_OPT_SMEM_BVH: c->smem_bvh = value != 0; return RT_OK;
    case RT_OPT_TOP_SMEM:
        if (value > 640u) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "at most 640 top-of-tree records (40 KB per CTA) can be staged");
        c->top_smem_records = value; return RT_OK;
    case RT_OPT_OVERLAP:
    {
        if (value > 2) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "overlap mode must be 0, 1 or 2");
        int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_;
        c->overlap = (int)value; return RT_OK;
    }
    case RT_OPT_TRAVERSAL:
        if (value > 1) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "traversal mode must be 0 (literal) or 1 (child-box layout)");
        { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
        c->traversal = (int)value; return RT_OK;
    }
    RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "unknown option key %d", key);
}

int rt_upload_sampler_tables(rt_ctx* c, const int32_t* sobol, const int32_t* scrambling, const int32_t* ranking)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_upload_sampler_tables(k, sobol, scrambling, ranking));
    if (!sobol || !scrambling || !ranking) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_sampler_tables: null table");
    for (int i = 0; i < RT_BN_TILE_COUNT; ++i)
        if (ranking[i] < 0 || ranking[i] > 255)
            RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_upload_sampler_tables: ranking entry %d is %d, outside 0..255", i, ranking[i]);
    RT_CUDA(c, cudaSetDevice(c->device));
    int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_;
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)RT_BN_SOBOL_COUNT + 2 * (size_t)RT_BN_TILE_COUNT;
    if (!c->d_bn) RT_CUDA(c, cudaMalloc(&c->d_bn, n * sizeof(int)));
    RT_CUDA(c, cudaMemcpy(c->d_bn, sobol, RT_BN_SOBOL_COUNT * sizeof(int), cudaMemcpyHostToDevice));
    RT_CUDA(c, cudaMemcpy(c->d_bn + RT_BN_SOBOL_COUNT, scrambling, RT_BN_TILE_COUNT * sizeof(int), cudaMemcpyHostToDevice));
    RT_CUDA(c, cudaMemcpy(c->d_bn + RT_BN_SOBOL_COUNT + RT_BN_TILE_COUNT, ranking, RT_BN_TILE_COUNT * sizeof(int), cudaMemcpyHostToDevice));
    ++c->config_gen;
    return RT_OK;
}

int rt_reset(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_reset(k));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    RT_CUDA(c, cudaSetDevice(c->device));
    if (!c->denoiser) c->sample_count = 0;                                   // cl_pt_integrator.cpp:499-504
    TimedLaunch t(c, RT_K_MISC);
    k_reset<<<grid_for(c->n_local), 256, 0, c->stream>>>(c->radiance, c->n_local);
    return post_launch(c, "k_reset");
}

int rt_advance_sample_count(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_advance_sample_count(k));
    int rc = join_shadow(c); if (rc) return rc;     // end of the bounce loop: everything of this frame is ordered on the render stream
    ++c->sample_count;
    return RT_OK;
}

int rt_generate_rays(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_generate_rays(k));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    int rc = require_ready(c); if (rc) return rc;
    RT_CUDA(c, cudaSetDevice(c->device));
    RT_CUDA(c, cudaMemsetAsync(c->counters, 0, sizeof(DevCounters), c->stream));
    TimedLaunch t(c, RT_K_RAYGEN);
    {
        FrameDyn dyn = frame_dyn(c);
        k_set_frame<<<1, 1, 0, c->stream>>>(c->d_dyn, dyn);
        ++c->launches;
    }
    k_raygen<<<grid_for(c->n_local), 256, 0, c->stream>>>(frame_params(c), c->q, c->counters, aov_params(c));
    c->frame_started = true;
    return post_launch(c, "k_raygen");
}

#define RT_BOUNCE_CHECK(c, b)                                                                                   \
    do {                                                                                                        \
        int rt_rc_ = require_ready(c); if (rt_rc_) return rt_rc_;                                               \
        if ((b) > RT_MAX_BOUNCES) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "bounce %u exceeds RT_MAX_BOUNCES", (b)); \
        if (!(c)->frame_started) RT_FAIL(c, RT_ERR_NOT_READY, "rt_generate_rays has not been called");          \
        RT_CUDA(c, cudaSetDevice((c)->device));                                                                 \
    } while (0)

int rt_intersect(rt_ctx* c, uint32_t bounce)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_intersect(k, bounce));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; } RT_BOUNCE_CHECK(c, bounce);
    TimedLaunch t(c, RT_K_INTERSECT);
    if (c->count_traversal) k_intersect<true><<<grid_for(c->n_local), 256, 0, c->stream>>>(frame_params(c), c->scene, c->traversal, c->q, c->counters, bounce);
    else k_intersect<false><<<grid_for(c->n_local), 256, 0, c->stream>>>(frame_params(c), c->scene, c->traversal, c->q, c->counters, bounce);
    return post_launch(c, "k_intersect");
}

/* ComputeAOVs: the AOV outputs are produced by the bounce-0 shading pass (it already holds the hit, the interpolated
 * normal and the unpacked material), so this step only makes sure the buffers exist when a view or the denoiser
 * needs them; with the default view (shaded colour, no denoiser) no AOV work is done at all. */
int rt_compute_aovs(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_compute_aovs(k));
    if (!aov_wanted(c)) return RT_OK;
    RT_CUDA(c, cudaSetDevice(c->device));
    return ensure_aov_buffers(c);
}

int rt_shade_miss(rt_ctx* c, uint32_t bounce)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_shade_miss(k, bounce));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; } RT_BOUNCE_CHECK(c, bounce);
    TimedLaunch t(c, RT_K_MISS);
    k_shade_miss<<<grid_for(c->n_local), 256, 0, c->stream>>>(frame_params(c), c->scene, c->q, c->counters, c->radiance, bounce);
    return post_launch(c, "k_shade_miss");
}

int rt_clear_outgoing_counter(rt_ctx* c, uint32_t) { RT_CHECK_CTX(c); return RT_OK; }
int rt_clear_shadow_counter(rt_ctx* c) { RT_CHECK_CTX(c); return RT_OK; }

int rt_shade_hits(rt_ctx* c, uint32_t bounce)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_shade_hits(k, bounce));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; } RT_BOUNCE_CHECK(c, bounce);
    c->cur_bounce = bounce;
    TimedLaunch t(c, RT_K_HIT);
    k_shade_hits<<<grid_for(c->n_local), 256, 0, c->stream>>>(frame_params(c), c->scene, c->q, c->counters, c->radiance, bounce, aov_params(c));
    return post_launch(c, "k_shade_hits");
}

int rt_intersect_shadow(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_intersect_shadow(k));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; } RT_BOUNCE_CHECK(c, c->cur_bounce);
    TimedLaunch t(c, RT_K_INTERSECT_SHADOW);
    if (c->count_traversal) k_intersect_shadow<true><<<grid_for(c->n_local), 256, 0, c->stream>>>(frame_params(c), c->scene, c->traversal, c->q, c->counters, c->cur_bounce);
    else k_intersect_shadow<false><<<grid_for(c->n_local), 256, 0, c->stream>>>(frame_params(c), c->scene, c->traversal, c->q, c->counters, c->cur_bounce);
    return post_launch(c, "k_intersect_shadow");
}

int rt_accumulate_direct(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_accumulate_direct(k));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; } RT_BOUNCE_CHECK(c, c->cur_bounce);
    TimedLaunch t(c, RT_K_ACCUMULATE);
    k_accumulate<<<grid_for(c->n_local), 256, 0, c->stream>>>(frame_params(c), c->q, c->counters, c->radiance, c->cur_bounce);
    return post_launch(c, "k_accumulate");
}

int rt_denoise(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_denoise(k));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    if (!c->denoiser) return RT_OK;
    if (c->world != 1) RT_FAIL(c, RT_ERR_UNSUPPORTED, "temporal denoiser is single-GPU only");
    RT_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_aov_buffers(c); if (rc) return rc;
    TimedLaunch t(c, RT_K_MISC);
    k_temporal<<<grid_for(c->n_local), 256, 0, c->stream>>>(c->width, c->height, c->radiance, c->prev_radiance, c->aov_depth, c->prev_depth, c->aov_velocity);
    return post_launch(c, "k_temporal");
}

int rt_copy_history(rt_ctx* c)
{   // cl_pt_integrator.cpp:670-675
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_copy_history(k));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    if (!c->denoiser) return RT_OK;
    RT_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_aov_buffers(c); if (rc) return rc;
    RT_CUDA(c, cudaMemcpyAsync(c->prev_radiance, c->radiance, (size_t)c->n_local * 16, cudaMemcpyDeviceToDevice, c->stream));
    RT_CUDA(c, cudaMemcpyAsync(c->prev_depth, c->aov_depth, (size_t)c->n_local * 4, cudaMemcpyDeviceToDevice, c->stream));
    return RT_OK;
}

int rt_extend_shade(rt_ctx* c, uint32_t bounce)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_extend_shade(k, bounce)); RT_BOUNCE_CHECK(c, bounce);
    c->cur_bounce = bounce;
    if (c->shadow_deferred && merged_trace_available(c))
    {   // this bounce's closest-hit traversal + the previous bounce's shadow pass in one kernel
        c->shadow_deferred = false;
        TimedLaunch t(c, RT_K_TRACE_BOTH);
        const Stage sg = bvh_stage(c);
        const DevScene sc = staged_scene(c, sg);
        const size_t sm = with_stack(c, sg.bytes);
        set_smem_limit((const void*)k_trace_both<0>, sm); set_smem_limit((const void*)k_trace_both<1>, sm); set_smem_limit((const void*)k_trace_both<2>, sm);
        if (sg.mode == 1) launch_chain(c, k_trace_both<1>, RT_PGRID(c, k_trace_both<1>, sm), sm, c->stream, frame_params(c), sc, c->traversal, c->q, c->counters, c->radiance, bounce, c->shadow_deferred_bounce);
        else if (sg.mode == 2) launch_chain(c, k_trace_both<2>, RT_PGRID(c, k_trace_both<2>, sm), sm, c->stream, frame_params(c), sc, c->traversal, c->q, c->counters, c->radiance, bounce, c->shadow_deferred_bounce);
        else launch_chain(c, k_trace_both<0>, RT_PGRID(c, k_trace_both<0>, sm), sm, c->stream, frame_params(c), sc, c->traversal, c->q, c->counters, c->radiance, bounce, c->shadow_deferred_bounce);
        int rc = post_launch(c, "k_trace_both"); if (rc) return rc;
    }
    else
    {
        TimedLaunch t(c, RT_K_TRACE_CLOSEST);
        const Stage sg = bvh_stage(c);
        const DevScene sc = staged_scene(c, sg);
        const size_t sm = with_stack(c, sg.bytes);
        set_smem_limit((const void*)k_trace_closest<false, 0>, sm); set_smem_limit((const void*)k_trace_closest<false, 1>, sm); set_smem_limit((const void*)k_trace_closest<false, 2>, sm);
        if (c->count_traversal) k_trace_closest<true, 0><<<RT_PGRID(c, (k_trace_closest<true, 0>), 0), 256, 0, c->stream>>>(frame_params(c), c->scene, c->traversal, c->q, c->counters, bounce);
        else if (sg.mode == 1) launch_chain(c, k_trace_closest<false, 1>, RT_PGRID(c, (k_trace_closest<false, 1>), sm), sm, c->stream, frame_params(c), sc, c->traversal, c->q, c->counters, bounce);
        else if (sg.mode == 2) launch_chain(c, k_trace_closest<false, 2>, RT_PGRID(c, (k_trace_closest<false, 2>), sm), sm, c->stream, frame_params(c), sc, c->traversal, c->q, c->counters, bounce);
        else launch_chain(c, k_trace_closest<false, 0>, RT_PGRID(c, (k_trace_closest<false, 0>), sm), sm, c->stream, frame_params(c), sc, c->traversal, c->q, c->counters, bounce);
        int rc = post_launch(c, "k_trace_closest"); if (rc) return rc;
    }
    { int rc = join_shadow(c); if (rc) return rc; }     // the shading pass accumulates into radiance and refills the shadow queue
    TimedLaunch t(c, RT_K_SHADE_QUEUES);
    launch_chain(c, k_shade_queues, RT_PGRID(c, k_shade_queues, 0), 0, c->stream, frame_params(c), c->scene, c->q, c->counters, c->radiance, bounce, aov_params(c));
    return post_launch(c, "k_shade_queues");
}

static int launch_shadow_pass(rt_ctx* c, uint32_t bounce, cudaStream_t st)
{
    TimedLaunch t(c, RT_K_SHADOW_ACCUMULATE, st);
    const Stage sg = bvh_stage(c);
    const DevScene sc = staged_scene(c, sg);
    const size_t sm = with_stack(c, sg.bytes);
    set_smem_limit((const void*)k_shadow_accumulate<false, 0>, sm); set_smem_limit((const void*)k_shadow_accumulate<false, 1>, sm); set_smem_limit((const void*)k_shadow_accumulate<false, 2>, sm);
    if (c->count_traversal) k_shadow_accumulate<true, 0><<<RT_PGRID(c, (k_shadow_accumulate<true, 0>), 0), 256, 0, st>>>(frame_params(c), c->scene, c->traversal, c->q, c->counters, c->radiance, bounce);
    else if (sg.mode == 1) k_shadow_accumulate<false, 1><<<RT_PGRID(c, (k_shadow_accumulate<false, 1>), sm), 256, sm, st>>>(frame_params(c), sc, c->traversal, c->q, c->counters, c->radiance, bounce);
    else if (sg.mode == 2) k_shadow_accumulate<false, 2><<<RT_PGRID(c, (k_shadow_accumulate<false, 2>), sm), 256, sm, st>>>(frame_params(c), sc, c->traversal, c->q, c->counters, c->radiance, bounce);
    else k_shadow_accumulate<false, 0><<<RT_PGRID(c, (k_shadow_accumulate<false, 0>), sm), 256, sm, st>>>(frame_params(c), sc, c->traversal, c->q, c->counters, c->radiance, bounce);
    return post_launch(c, "k_shadow_accumulate");
}

int rt_shadow_accumulate(rt_ctx* c, uint32_t bounce)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_shadow_accumulate(k, bounce)); RT_BOUNCE_CHECK(c, bounce);
    int rc = join_shadow(c); if (rc) return rc;
    // The shadow pass of bounce b only shares the radiance buffer with LATER shading passes, so it may overlap the
    // closest-hit traversal of bounce b+1 (which touches neither).
    if (c->overlap == 2 && merged_trace_available(c))
    {   // deferred: rt_extend_shade(b+1) runs it inside its traversal kernel; join_shadow() launches it for anyone else
        c->shadow_deferred = true; c->shadow_deferred_bounce = bounce;
        return RT_OK;
    }
    if (c->overlap != 1) return launch_shadow_pass(c, bounce, c->stream);
    // overlap 1: second stream; join_shadow() orders it before the next kernel that needs its results.  Both kernels are
    // persistent, so the second fills the SMs as the first drains.
    if (!c->shadow_stream)
    {
        RT_CUDA(c, cudaStreamCreateWithFlags(&c->shadow_stream, cudaStreamNonBlocking));
        RT_CUDA(c, cudaEventCreateWithFlags(&c->ev_shaded, cudaEventDisableTiming));
        RT_CUDA(c, cudaEventCreateWithFlags(&c->ev_shadowed, cudaEventDisableTiming));
    }
    cudaStream_t st = c->shadow_stream;
    RT_CUDA(c, cudaEventRecord(c->ev_shaded, c->stream));
    RT_CUDA(c, cudaStreamWaitEvent(st, c->ev_shaded, 0));
    rc = launch_shadow_pass(c, bounce, st); if (rc) return rc;
    RT_CUDA(c, cudaEventRecord(c->ev_shadowed, st));
    c->shadow_pending = true;
    return RT_OK;
}

static int integrate_body(rt_ctx* c, uint32_t max_bounces)
{
    int rc = rt_generate_rays(c); if (rc) return rc;
    for (uint32_t b = 0; b <= max_bounces; ++b)          // inclusive, integrator.cpp:37
    {
        if ((rc = rt_extend_shade(c, b))) return rc;
        if ((rc = rt_shadow_accumulate(c, b))) return rc;
    }
    return rt_advance_sample_count(c);      // joins the shadow stream: the frame is complete in render-stream order
}

static void drop_graph(rt_ctx* c)
{
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    c->graph_exec = nullptr; c->graph = nullptr; c->graph_set_frame_node = nullptr; c->graph_gen = 0;
}

// Captures one frame (both streams) into a graph.  Nothing executes during capture; host-side frame state is restored.
static int capture_frame_graph(rt_ctx* c, uint32_t max_bounces)
{
    drop_graph(c);
    int rc = join_shadow(c); if (rc) return rc;
    if (c->overlap == 1 && !c->shadow_stream)
    {   // create the second stream and its events outside the capture
        RT_CUDA(c, cudaStreamCreateWithFlags(&c->shadow_stream, cudaStreamNonBlocking));
        RT_CUDA(c, cudaEventCreateWithFlags(&c->ev_shaded, cudaEventDisableTiming));
        RT_CUDA(c, cudaEventCreateWithFlags(&c->ev_shadowed, cudaEventDisableTiming));
    }
    const uint32_t saved_samples = c->sample_count, saved_bounce = c->cur_bounce;
    const uint64_t saved_launches = c->launches;
    const bool saved_started = c->frame_started;
    RT_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    rc = integrate_body(c, max_bounces);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    c->graph_launches = c->launches - saved_launches;
    c->sample_count = saved_samples; c->cur_bounce = saved_bounce; c->launches = saved_launches; c->frame_started = saved_started;
    c->shadow_pending = false; c->shadow_deferred = false;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) RT_FAIL(c, RT_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    size_t n_nodes = 0;
    RT_CUDA(c, cudaGraphGetNodes(graph, nullptr, &n_nodes));
    std::vector<cudaGraphNode_t> nodes(n_nodes);
    RT_CUDA(c, cudaGraphGetNodes(graph, nodes.data(), &n_nodes));
    for (cudaGraphNode_t nd : nodes)
    {
        cudaGraphNodeType ty;
        RT_CUDA(c, cudaGraphNodeGetType(nd, &ty));
        if (ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        RT_CUDA(c, cudaGraphKernelNodeGetParams(nd, &kp));
        if (kp.gridDim.x == 1 && kp.blockDim.x == 1 && !c->graph_set_frame_node)
        {   // k_set_frame is the only <<<1,1>>> launch of the frame and the first kernel of it
            c->graph_set_frame_node = nd; c->graph_set_frame_func = kp.func;
        }
    }
    if (!c->graph_set_frame_node) { cudaGraphDestroy(graph); RT_FAIL(c, RT_ERR_CUDA, "frame graph has no k_set_frame node"); }
    e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); c->graph_exec = nullptr; RT_FAIL(c, RT_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    c->graph = graph;
    c->graph_gen = c->config_gen; c->graph_max_bounces = max_bounces;
    return RT_OK;
}

// Which schedule rt_integrate runs, and the CTA shape of the one-kernel frame.  Measured on B200 (profiles/r02_frame_kernel_ab.txt,
// ms per frame of a 1/world partition of the 1080p frame, per-phase kernels vs k_frame at its best CTA size):
//   CornellBox   world 1: 2.141 / 2.181   2: 1.157 / 1.134   4: 0.657 / 0.611   8: 0.419 / 0.350
//   ShaderBalls  world 1: 3.985 / 4.041   2: 2.485 / 2.452   4: 1.768 / 1.571   8: 1.406 / 1.106
//   Dragon 4K    world 1: 24.74 / 26.59                                         8: 4.273 / 4.102
// With a whole frame per GPU the per-phase kernels win by 2-7 % (specialised register budgets — 5 CTAs/SM of 48 registers for
// traversal — and one kind of code per SM at a time); from half a 1080p frame down the launch boundaries and kernel tails of
// 20+ dependent launches cost more and the one-kernel frame wins (1.2-1.3x on a 1/8 partition).  RT_OPT_FRAME_KERNEL = 2
// switches at 7168 pixels per SM (1.06 M pixels on 148 SMs).
static bool frame_kernel_selected(const rt_ctx* c)
{
    if (c->frame_kernel == 2) return (size_t)c->n_local <= (size_t)c->num_sms * 7168u;
    return c->frame_kernel == 1;
}

// CTA size of k_frame (same measurements).  One 1024-thread CTA per SM keeps all 32 warps of an SM in the same phase — one kind
// of code in the instruction caches — and is the best shape when the BVH lives in shared memory (CornellBox, every partition
// size) or every warp still has several items per phase; scenes traversed through L1/L2 on small partitions (<= 4096 pixels
// per SM) prefer 4 CTAs of 256 threads per SM whose phases interleave (ShaderBalls, 1/8 partition: 1.106 vs 1.252 ms).
static int frame_kernel_threads(const rt_ctx* c)
{
    if (c->frame_threads) return c->frame_threads;
    if (bvh_stage(c).mode == 1) return 1024;
    return (size_t)c->n_local > (size_t)c->num_sms * 4096u ? 1024 : 256;
}

// Fused gather, end of a frame: k_frame has pushed every pixel already (copy_slab false); the per-phase schedule pushes its whole
// slab with one device-to-device copy.  Then the completion flag.
static int gather_signal(rt_ctx* c, bool copy_slab)
{
    if (!c->gather_slab) return RT_OK;
    if (copy_slab && c->n_local)
        RT_CUDA(c, cudaMemcpyAsync(c->gather_slab, c->radiance, (size_t)c->n_local * 16, cudaMemcpyDefault, c->stream));
    ++c->gather_frame;
    c->gather_sample = c->sample_count;
    k_gather_signal<<<1, 1, 0, c->stream>>>(c->gather_flag, c->gather_frame);
    ++c->launches;
    return post_launch(c, "k_gather_signal");
}

// The whole frame as ONE persistent kernel (k_frame): a counter clear and a launch.
static int integrate_frame_kernel(rt_ctx* c, uint32_t max_bounces)
{
    int rc = require_ready(c); if (rc) return rc;
    if ((rc = join_shadow(c))) return rc;
    RT_CUDA(c, cudaSetDevice(c->device));
    RT_CUDA(c, cudaMemsetAsync(c->counters, 0, sizeof(DevCounters), c->stream));
    const Stage sg = bvh_stage(c);
    const DevScene sc = staged_scene(c, sg);
    const size_t stage = sg.bytes;
    const void* kern = sg.mode == 1 ? (const void*)k_frame<1> : (sg.mode == 2 ? (const void*)k_frame<2> : (const void*)k_frame<0>);
    const int threads = frame_kernel_threads(c);
    int per_sm = 0;
    for (auto& e : c->occupancy) if (e.kernel == kern && e.smem == stage + ((size_t)threads << 32)) { per_sm = e.per_sm; break; }
    if (!per_sm)
    {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, stage) != cudaSuccess || per_sm < 1) per_sm = 1;
        if (per_sm > 32) per_sm = 32;
        c->occupancy.push_back({ kern, stage + ((size_t)threads << 32), per_sm });
    }
    const uint32_t n_groups = (c->n_local + 31u) / 32u;
    uint32_t grid = (uint32_t)(c->num_sms * per_sm / c->grid_div);
    if (grid < (uint32_t)c->num_sms) grid = (uint32_t)c->num_sms;
    if (grid > n_groups) grid = n_groups;
    if (grid < 1) grid = 1;
    const uint32_t slots_per_cta = ((n_groups + grid - 1u) / grid) * 32u;
    if ((size_t)grid * slots_per_cta > c->n_alloc) RT_FAIL(c, RT_ERR_CUDA, "frame kernel: queue regions exceed the allocation");
    const FrameDyn dyn = frame_dyn(c);
    FrameParams fp = frame_params(c);
    fp.gather = c->gather_slab;
    {
        TimedLaunch t(c, RT_K_MISC);
        if (sg.mode == 1) k_frame<1><<<grid, threads, stage, c->stream>>>(fp, sc, c->traversal, c->q, c->counters, c->radiance, aov_params(c), max_bounces, slots_per_cta, dyn);
        else if (sg.mode == 2) k_frame<2><<<grid, threads, stage, c->stream>>>(fp, sc, c->traversal, c->q, c->counters, c->radiance, aov_params(c), max_bounces, slots_per_cta, dyn);
        else k_frame<0><<<grid, threads, 0, c->stream>>>(fp, sc, c->traversal, c->q, c->counters, c->radiance, aov_params(c), max_bounces, slots_per_cta, dyn);
    }
    if ((rc = post_launch(c, "k_frame"))) return rc;
    c->frame_started = true; c->cur_bounce = max_bounces;
    ++c->sample_count;
    return gather_signal(c, false);
}

int rt_integrate(rt_ctx* c, uint32_t max_bounces)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_integrate(k, max_bounces));
    if (max_bounces > RT_MAX_BOUNCES) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "max_bounces %u exceeds RT_MAX_BOUNCES", max_bounces);
    if (frame_kernel_selected(c) && !c->kernel_timing && !c->count_traversal) return integrate_frame_kernel(c, max_bounces);
    if (!c->use_graph || c->kernel_timing || c->count_traversal)
    {
        int rc = integrate_body(c, max_bounces); if (rc) return rc;
        return gather_signal(c, true);
    }
    int rc = require_ready(c); if (rc) return rc;
    RT_CUDA(c, cudaSetDevice(c->device));
    if (!c->graph_exec || c->graph_gen != c->config_gen || c->graph_max_bounces != max_bounces)
        if ((rc = capture_frame_graph(c, max_bounces))) return rc;
    if ((rc = join_shadow(c))) return rc;
    // the one per-frame update: k_set_frame's by-value argument (sample index + camera constants)
    FrameDyn dyn = frame_dyn(c);
    FrameDyn* dst = c->d_dyn;
    void* args[2] = { &dst, &dyn };
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.func = c->graph_set_frame_func; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.sharedMemBytes = 0; kp.kernelParams = args; kp.extra = nullptr;
    RT_CUDA(c, cudaGraphExecKernelNodeSetParams(c->graph_exec, c->graph_set_frame_node, &kp));
    RT_CUDA(c, cudaGraphLaunch(c->graph_exec, c->stream));
    c->launches += c->graph_launches;
    c->frame_started = true; c->cur_bounce = max_bounces;
    ++c->sample_count;
    return gather_signal(c, true);
}

// resolve kernel + device->host copy of this context's rows, enqueued on its stream (no synchronisation)
static int resolve_enqueue(rt_ctx* c, float* dst)
{
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    RT_CUDA(c, cudaSetDevice(c->device));
    if (c->copy_pending[0]) RT_CUDA(c, cudaStreamWaitEvent(c->stream, c->copy_done[0], 0));   // rt_resolve_async may still copy out of c->resolved
    {
        TimedLaunch t(c, RT_K_RESOLVE);
        k_resolve<<<grid_for(c->n_local), 256, 0, c->stream>>>(c->radiance, c->resolved, c->n_local, c->sample_count, c->denoiser,
                                                                  aov_params(c).enabled ? c->aov : 0, aov_params(c));
        int rc = post_launch(c, "k_resolve"); if (rc) return rc;
    }
    if (dst && c->local_rows)
        RT_CUDA(c, cudaMemcpy2DAsync(dst + (size_t)c->rank * c->width * 4, (size_t)c->world * c->width * 16, c->resolved,
                                     (size_t)c->width * 16, (size_t)c->width * 16, c->local_rows, cudaMemcpyDeviceToHost, c->stream));
    return RT_OK;
}

int rt_resolve(rt_ctx* c, float* dst)
{
    RT_CHECK_CTX(c);
    if (!c->children.empty())
    {
        if (c->present == 1 && dst)
        {   // the frame's ONE collective: radiance slabs -> first device over NVLink, resolved and read back there
            rt_ctx* k0 = c->children[0];
            bool pushed = true;              // did the frame kernels push this frame already (fused gather)?
            for (rt_ctx* k : c->children) pushed = pushed && k->gather_slab && k->gather_sample == k->sample_count;
            int rc;
            if (pushed)
            {
                rc = rt_gather_wait(k0);
                if (!rc) rc = rt_resolve_gathered(k0, k0->own_gather, (uint64_t)k0->own_gather_stride, dst);
            }
            else
            {   // frames rendered before the option was set: explicit peer copies
                rc = rt_gather_radiance(c); if (rc) return rc;
                rc = rt_resolve_gathered(k0, c->gather_buf, (uint64_t)c->gather_stride_f4 * 16, dst);
            }
            if (rc) c->error = k0->error;
            return rc;
        }
        // parallel read-back: every device resolves its rows and copies them into the caller's image over its own PCIe
        // link (page-lock the image, rt_host_register, for the copies to overlap); all enqueued first, then awaited
        for (rt_ctx* k : c->children) { int rc = resolve_enqueue(k, dst); if (rc) { c->error = k->error; return rc; } }
        for (rt_ctx* k : c->children)
        {
            if (cudaSetDevice(k->device) != cudaSuccess || cudaStreamSynchronize(k->stream) != cudaSuccess)
                RT_FAIL(c, RT_ERR_CUDA, "rt_resolve: device %d failed: %s", k->device, cudaGetErrorString(cudaGetLastError()));
        }
        return RT_OK;
    }
    int rc = resolve_enqueue(c, dst); if (rc) return rc;
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    return RT_OK;
}

/* Pipelined variant of rt_resolve for frame loops that do not need the image before submitting the next frame:
 * resolves on the render stream, then copies device->host on a second stream so that the PCIe transfer of frame i
 * overlaps the kernels of frame i+1.  Two resolve buffers alternate; the host buffer handed to call i must stay
 * untouched until rt_resolve_wait() (or the second-next rt_resolve_async) returns. */
int rt_resolve_async(rt_ctx* c, float* dst)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_resolve_async(k, dst));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    if (!dst) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_resolve_async: null destination");
    RT_CUDA(c, cudaSetDevice(c->device));
    if (!c->copy_stream)
    {
        RT_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i)
        {
            RT_CUDA(c, cudaEventCreateWithFlags(&c->resolve_done[i], cudaEventDisableTiming));
            RT_CUDA(c, cudaEventCreateWithFlags(&c->copy_done[i], cudaEventDisableTiming));
        }
    }
    if (!c->resolved2) RT_CUDA(c, cudaMalloc(&c->resolved2, (size_t)(c->n_local ? c->n_local : 1) * 16));
    const int k = (int)(c->async_index++ & 1u);
    float4* buf = k ? c->resolved2 : c->resolved;
    if (c->copy_pending[k]) RT_CUDA(c, cudaStreamWaitEvent(c->stream, c->copy_done[k], 0));   // buffer k is being re-used
    {
        TimedLaunch t(c, RT_K_RESOLVE);
        AovParams ap = aov_params(c);
        k_resolve<<<grid_for(c->n_local), 256, 0, c->stream>>>(c->radiance, buf, c->n_local, c->sample_count, c->denoiser, ap.enabled ? c->aov : 0, ap);
        int rc = post_launch(c, "k_resolve"); if (rc) return rc;
    }
    RT_CUDA(c, cudaEventRecord(c->resolve_done[k], c->stream));
    RT_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->resolve_done[k], 0));
    if (c->local_rows)
        RT_CUDA(c, cudaMemcpy2DAsync(dst + (size_t)c->rank * c->width * 4, (size_t)c->world * c->width * 16, buf,
                                     (size_t)c->width * 16, (size_t)c->width * 16, c->local_rows, cudaMemcpyDeviceToHost, c->copy_stream));
    RT_CUDA(c, cudaEventRecord(c->copy_done[k], c->copy_stream));
    c->copy_pending[k] = true;
    return RT_OK;
}

static int resolve_gathered_impl(rt_ctx* c, const void* slabs, uint64_t stride_bytes, float* dst, bool async)
{
    RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_resolve_gathered");
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    if (!slabs || !dst) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_resolve_gathered: null pointer");
    const size_t rows_max = ((size_t)c->height + c->world - 1) / c->world;
    if (stride_bytes % 16 != 0 || stride_bytes < rows_max * c->width * 16)
        RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_resolve_gathered: slab stride must be a multiple of 16 and hold %zu rows", rows_max);
    if (c->aov != 0 || c->denoiser) RT_FAIL(c, RT_ERR_UNSUPPORTED, "rt_resolve_gathered resolves the shaded colour only (AOV views and the denoiser are single-GPU)");
    RT_CUDA(c, cudaSetDevice(c->device));
    const size_t n = (size_t)c->width * c->height;
    int k = 0;
    if (async)
    {
        if (!c->copy_stream)
        {
            RT_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i)
            {
                RT_CUDA(c, cudaEventCreateWithFlags(&c->resolve_done[i], cudaEventDisableTiming));
                RT_CUDA(c, cudaEventCreateWithFlags(&c->copy_done[i], cudaEventDisableTiming));
            }
        }
        k = (int)(c->async_index++ & 1u);
        if (c->copy_pending[k]) RT_CUDA(c, cudaStreamWaitEvent(c->stream, c->copy_done[k], 0));
    }
    if (!c->resolved_full[k]) RT_CUDA(c, cudaMalloc(&c->resolved_full[k], n * 16));
    {
        TimedLaunch t(c, RT_K_RESOLVE);
        k_resolve_gathered<<<grid_for((uint32_t)n), 256, 0, c->stream>>>((const float4*)slabs, (size_t)(stride_bytes / 16), c->resolved_full[k],
                                                                            c->width, c->height, c->world, c->sample_count);
        int rc = post_launch(c, "k_resolve_gathered"); if (rc) return rc;
    }
    if (!async)
    {
        RT_CUDA(c, cudaMemcpyAsync(dst, c->resolved_full[k], n * 16, cudaMemcpyDeviceToHost, c->stream));
        RT_CUDA(c, cudaStreamSynchronize(c->stream));
        return RT_OK;
    }
    RT_CUDA(c, cudaEventRecord(c->resolve_done[k], c->stream));
    RT_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->resolve_done[k], 0));
    RT_CUDA(c, cudaMemcpyAsync(dst, c->resolved_full[k], n * 16, cudaMemcpyDeviceToHost, c->copy_stream));
    RT_CUDA(c, cudaEventRecord(c->copy_done[k], c->copy_stream));
    c->copy_pending[k] = true;
    return RT_OK;
}
int rt_resolve_gathered(rt_ctx* c, const void* slabs, uint64_t stride_bytes, float* dst) { return resolve_gathered_impl(c, slabs, stride_bytes, dst, false); }
int rt_resolve_gathered_async(rt_ctx* c, const void* slabs, uint64_t stride_bytes, float* dst) { return resolve_gathered_impl(c, slabs, stride_bytes, dst, true); }

int rt_resolve_wait(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_resolve_wait(k));
    RT_CUDA(c, cudaSetDevice(c->device));
    if (c->copy_stream) RT_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    c->copy_pending[0] = c->copy_pending[1] = false;
    return RT_OK;
}

int rt_sync(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_sync(k));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    RT_CUDA(c, cudaSetDevice(c->device));
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    return RT_OK;
}

static int ensure_scratch(rt_ctx* c, size_t bytes)
{
    if (c->scratch_bytes >= bytes) return RT_OK;
    cudaFree(c->scratch); c->scratch = nullptr; c->scratch_bytes = 0;
    RT_CUDA(c, cudaMalloc(&c->scratch, bytes));
    c->scratch_bytes = bytes;
    return RT_OK;
}

int rt_read_hits(rt_ctx* c, uint32_t bounce, RtHit* hits, uint32_t* pixels, uint32_t* n_out)
{
    RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_read_hits");
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    if (bounce > RT_MAX_BOUNCES || !n_out) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_read_hits: bad arguments");
    RT_CUDA(c, cudaSetDevice(c->device));
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    uint32_t n = 0;
    RT_CUDA(c, cudaMemcpy(&n, in_count_ptr(c->counters, bounce), 4, cudaMemcpyDeviceToHost));
    *n_out = n;
    if (hits && n) RT_CUDA(c, cudaMemcpy(hits, c->q.hits, (size_t)n * 16, cudaMemcpyDeviceToHost));   // same 16-byte layout as RtHit
    if (pixels && n)
    {
        std::vector<float4> a(n);
        RT_CUDA(c, cudaMemcpy(a.data(), c->q.A[bounce & 1], (size_t)n * 16, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < n; ++i) { uint32_t pxy; memcpy(&pxy, &a[i].w, 4); pixels[i] = (pxy >> 16) * c->width + (pxy & 0xFFFFu); }
    }
    return RT_OK;
}

int rt_read_rays(rt_ctx* c, uint32_t bounce, RtRay* rays, uint32_t* pixels, uint32_t* n_out)
{
    RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_read_rays");
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    if (bounce > RT_MAX_BOUNCES || !n_out || !rays || !pixels) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_read_rays: bad arguments");
    RT_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_scratch(c, (size_t)c->n_local * (sizeof(RtRay) + 4) + 64); if (rc) return rc;
    RtRay* drays = (RtRay*)c->scratch;
    uint32_t* dpix = (uint32_t*)((char*)c->scratch + (size_t)c->n_local * sizeof(RtRay));
    k_unpack_rays<<<grid_for(c->n_local), 256, 0, c->stream>>>(c->q.A[bounce & 1], c->q.B[bounce & 1], in_count_ptr(c->counters, bounce), c->width, drays, dpix);
    ++c->launches;
    if ((rc = post_launch(c, "k_unpack_rays"))) return rc;
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    uint32_t n = 0;
    RT_CUDA(c, cudaMemcpy(&n, in_count_ptr(c->counters, bounce), 4, cudaMemcpyDeviceToHost));
    *n_out = n;
    if (n)
    {
        RT_CUDA(c, cudaMemcpy(rays, drays, (size_t)n * sizeof(RtRay), cudaMemcpyDeviceToHost));
        RT_CUDA(c, cudaMemcpy(pixels, dpix, (size_t)n * 4, cudaMemcpyDeviceToHost));
    }
    return RT_OK;
}

int rt_read_radiance(rt_ctx* c, float* dst)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_read_radiance(k, dst));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    if (!dst) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_read_radiance: null destination");
    RT_CUDA(c, cudaSetDevice(c->device));
    if (c->local_rows)
        RT_CUDA(c, cudaMemcpy2DAsync(dst + (size_t)c->rank * c->width * 4, (size_t)c->world * c->width * 16, c->radiance,
                                     (size_t)c->width * 16, (size_t)c->width * 16, c->local_rows, cudaMemcpyDeviceToHost, c->stream));
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    return RT_OK;
}

int rt_read_frame_stats(rt_ctx* c, RtFrameStats* out)
{
    RT_CHECK_CTX(c);
    if (!c->children.empty())
    {   // whole-frame counters = sums over the partitions
        if (!out) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_read_frame_stats: null destination");
        static thread_local RtFrameStats part;
        memset(out, 0, sizeof(*out));
        for (rt_ctx* k : c->children)
        {
            int rc = rt_read_frame_stats(k, &part); if (rc) { c->error = k->error; return rc; }
            for (uint32_t b = 0; b <= RT_MAX_BOUNCES; ++b)
            {
                out->n_ext[b] += part.n_ext[b]; out->n_miss[b] += part.n_miss[b]; out->n_emissive_hits[b] += part.n_emissive_hits[b];
                out->n_shadow[b] += part.n_shadow[b]; out->n_cont[b] += part.n_cont[b]; out->n_unoccluded[b] += part.n_unoccluded[b];
                out->nodes_ext[b] += part.nodes_ext[b]; out->tris_ext[b] += part.tris_ext[b];
                out->nodes_shadow[b] += part.nodes_shadow[b]; out->tris_shadow[b] += part.tris_shadow[b];
            }
        }
        return RT_OK;
    }
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    if (!out) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_read_frame_stats: null destination");
    RT_CUDA(c, cudaSetDevice(c->device));
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    static thread_local DevCounters h;
    RT_CUDA(c, cudaMemcpy(&h, c->counters, sizeof(DevCounters), cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof(*out));
    for (uint32_t b = 0; b <= RT_MAX_BOUNCES; ++b)
    {
        out->n_ext[b] = (b == 0) ? h.n_primary : h.emit[b - 1].next; out->n_miss[b] = h.hm[b].miss; out->n_emissive_hits[b] = h.n_emissive[b];
        out->n_shadow[b] = h.emit[b].shadow; out->n_cont[b] = h.emit[b].next; out->n_unoccluded[b] = h.n_unoccluded[b];
        out->nodes_ext[b] = h.nodes_ext[b]; out->tris_ext[b] = h.tris_ext[b];
        out->nodes_shadow[b] = h.nodes_shadow[b]; out->tris_shadow[b] = h.tris_shadow[b];
    }
    return RT_OK;
}

int rt_read_sample_count(rt_ctx* c, uint32_t* out) { RT_CHECK_CTX(c); if (!out) return RT_ERR_INVALID_ARGUMENT; *out = c->children.empty() ? c->sample_count : c->children[0]->sample_count; return RT_OK; }

int rt_read_aovs(rt_ctx* c, float* albedo, float* depth, float* normal, float* velocity)
{
    RT_CHECK_CTX(c); RT_FANOUT(c, rt_read_aovs(k, albedo, depth, normal, velocity));
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    if (!c->aov_albedo) RT_FAIL(c, RT_ERR_NOT_READY, "AOV buffers do not exist: select an AOV view, enable the denoiser or set RT_OPT_AOV_ALWAYS first");
    RT_CUDA(c, cudaSetDevice(c->device));
    auto rows = [&](void* dst, const void* src, size_t elem) -> int {
        if (dst && c->local_rows)
            RT_CUDA(c, cudaMemcpy2DAsync((char*)dst + (size_t)c->rank * c->width * elem, (size_t)c->world * c->width * elem, src,
                                         (size_t)c->width * elem, (size_t)c->width * elem, c->local_rows, cudaMemcpyDeviceToHost, c->stream));
        return RT_OK;
    };
    int rc;
    if ((rc = rows(albedo, c->aov_albedo, 16)) || (rc = rows(depth, c->aov_depth, 4)) || (rc = rows(normal, c->aov_normal, 16)) || (rc = rows(velocity, c->aov_velocity, 8))) return rc;
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    return RT_OK;
}

int rt_kernel_times(rt_ctx* c, float* ms, uint32_t* launches)
{
    RT_CHECK_CTX(c);
    if (!c->children.empty())
    {   // the devices run concurrently: per class the slowest device's time, launches summed
        float m[RT_K_CLASS_COUNT]; uint32_t n[RT_K_CLASS_COUNT];
        for (int i = 0; i < RT_K_CLASS_COUNT; ++i) { if (ms) ms[i] = 0.0f; if (launches) launches[i] = 0; }
        for (rt_ctx* k : c->children)
        {
            int rc = rt_kernel_times(k, m, n); if (rc) { c->error = k->error; return rc; }
            for (int i = 0; i < RT_K_CLASS_COUNT; ++i) { if (ms && m[i] > ms[i]) ms[i] = m[i]; if (launches) launches[i] += n[i]; }
        }
        return RT_OK;
    }
    { int rt_j_ = join_shadow(c); if (rt_j_) return rt_j_; }
    RT_CUDA(c, cudaSetDevice(c->device));
    RT_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& t : c->timed)
    {
        float m = 0.0f;
        RT_CUDA(c, cudaEventElapsedTime(&m, t.a, t.b));
        c->ms[t.cls] += m; c->nlaunch[t.cls] += 1;
        c->event_pool.push_back(t.a); c->event_pool.push_back(t.b);
    }
    c->timed.clear();
    for (int k = 0; k < RT_K_CLASS_COUNT; ++k)
    {
        if (ms) ms[k] = c->ms[k];
        if (launches) launches[k] = c->nlaunch[k];
        c->ms[k] = 0.0f; c->nlaunch[k] = 0;
    }
    return RT_OK;
}

int rt_launch_count(rt_ctx* c, uint64_t* out)
{
    RT_CHECK_CTX(c); if (!out) return RT_ERR_INVALID_ARGUMENT;
    *out = c->launches;
    for (rt_ctx* k : c->children) *out += k->launches;
    return RT_OK;
}
int rt_local_pixel_count(rt_ctx* c, uint32_t* out)
{
    RT_CHECK_CTX(c); if (!out) return RT_ERR_INVALID_ARGUMENT;
    *out = c->n_local;
    for (rt_ctx* k : c->children) *out += k->n_local;
    return RT_OK;
}

int rt_radiance_device_ptr(rt_ctx* c, void** out_ptr, uint64_t* out_bytes)
{
    RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_radiance_device_ptr");
    if (!out_ptr || !out_bytes) return RT_ERR_INVALID_ARGUMENT;
    *out_ptr = c->radiance; *out_bytes = (uint64_t)c->n_local * 16;
    return RT_OK;
}

int rt_stream_handle(rt_ctx* c, void** out) { RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_stream_handle"); if (!out) return RT_ERR_INVALID_ARGUMENT; *out = (void*)c->stream; return RT_OK; }

/* One context over several devices of the node ("multi-GPU fan-out internal to the shim"; the reference's CLContext
 * enumerates every device of its platform, cl_context.cpp:64-89, and uses one).  devices[i] renders rank i of an n-way scanline
 * partition with the scene replicated; the caller stays single-threaded and every call fans out.  A device may be listed more
 * than once (two partitions time-share it). */
int rt_create_multi(uint32_t width, uint32_t height, const int* devices, uint32_t n_devices, rt_ctx** out_ctx)
{
    if (!out_ctx || !devices || n_devices == 0 || n_devices > 64)
    {
        g_create_error = "rt_create_multi: bad arguments (need 1..64 devices)"; return RT_ERR_INVALID_ARGUMENT;
    }
    *out_ctx = nullptr;
    if (n_devices == 1) return rt_create(width, height, devices[0], out_ctx);
    rt_ctx* g = new rt_ctx;
    g->width = width; g->height = height; g->world = n_devices; g->device = devices[0];
    for (uint32_t i = 0; i < n_devices; ++i)
    {
        rt_ctx* k = nullptr;
        int rc = rt_create(width, height, devices[i], &k);
        if (rc == RT_OK) { rc = rt_set_partition(k, i, n_devices); if (rc != RT_OK) { g_create_error = k->error; rt_destroy(k); } }
        if (rc != RT_OK) { for (rt_ctx* o : g->children) rt_destroy(o); delete g; return rc; }
        g->children.push_back(k);
    }
    // contexts that time-share one device split its resident CTA slots, so that their persistent kernels run side by side
    for (rt_ctx* k : g->children)
    {
        int same = 0;
        for (rt_ctx* o : g->children) same += o->device == k->device;
        k->grid_div = same;
    }
    // peer access towards the first device for the NVLink gather (ignored where it is already on / not available:
    // cudaMemcpyPeerAsync then stages through the host)
    for (uint32_t i = 1; i < n_devices; ++i)
        if (devices[i] != devices[0])
        {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can)
            {
                cudaSetDevice(devices[i]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e != cudaSuccess) cudaGetLastError();       // cudaErrorPeerAccessAlreadyEnabled included
            }
        }
    *out_ctx = g;
    return RT_OK;
}

/* The single collective of a multi-device frame (north_star): every device's radiance slab is copied to the first device
 * (cudaMemcpyPeerAsync on the source device's stream: NVLink / NVSwitch between peers), into world slabs of stride
 * rows_max * width float4 — the layout rt_resolve_gathered reads.  The first device's stream then waits for all copies. */
int rt_gather_radiance(rt_ctx* c)
{
    RT_CHECK_CTX(c);
    if (c->children.empty()) RT_FAIL(c, RT_ERR_UNSUPPORTED, "rt_gather_radiance needs a multi-device context (rt_create_multi)");
    rt_ctx* k0 = c->children[0];
    const size_t rows_max = ((size_t)c->height + c->world - 1) / c->world;
    const size_t stride = rows_max * c->width;
    RT_CUDA(c, cudaSetDevice(k0->device));
    if (!c->gather_buf)
    {
        RT_CUDA(c, cudaMalloc(&c->gather_buf, stride * c->world * 16));
        RT_CUDA(c, cudaMemsetAsync(c->gather_buf, 0, stride * c->world * 16, k0->stream));
        RT_CUDA(c, cudaStreamSynchronize(k0->stream));
        c->gather_stride_f4 = stride;
        c->gather_events.resize(c->world, nullptr);
    }
    for (uint32_t i = 0; i < c->world; ++i)
    {
        rt_ctx* k = c->children[i];
        { int rc = join_shadow(k); if (rc) { c->error = k->error; return rc; } }
        RT_CUDA(c, cudaSetDevice(k->device));
        if (!c->gather_events[i]) RT_CUDA(c, cudaEventCreateWithFlags(&c->gather_events[i], cudaEventDisableTiming));
        if (k->n_local)
            RT_CUDA(c, cudaMemcpyPeerAsync(c->gather_buf + (size_t)i * stride, k0->device, k->radiance, k->device, (size_t)k->n_local * 16, k->stream));
        RT_CUDA(c, cudaEventRecord(c->gather_events[i], k->stream));
    }
    RT_CUDA(c, cudaSetDevice(k0->device));
    for (uint32_t i = 1; i < c->world; ++i) RT_CUDA(c, cudaStreamWaitEvent(k0->stream, c->gather_events[i], 0));
    return RT_OK;
}

int rt_device_count(int* out)
{
    if (!out) return RT_ERR_INVALID_ARGUMENT;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    *out = n;
    return RT_OK;
}

/* Page-locks / releases a caller-owned host buffer (the image handed to rt_resolve): device->host copies into page-locked
 * memory are asynchronous, which is what lets the devices of a multi-device context read back in parallel. */
int rt_host_register(void* ptr, uint64_t bytes)
{
    if (!ptr || !bytes) return RT_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("rt_host_register: ") + cudaGetErrorString(e); return RT_ERR_CUDA; }
    return RT_OK;
}
int rt_host_unregister(void* ptr)
{
    if (!ptr) return RT_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("rt_host_unregister: ") + cudaGetErrorString(e); return RT_ERR_CUDA; }
    return RT_OK;
}

/* ---- fused gather over peer memory ------------------------------------------------------------------------------------------
 * The frame's one collective done by the frame kernel itself: every rank pushes the radiance of a pixel into its slab of a buffer
 * on the presenting device (NVLink peer stores) the moment the pixel's path ends, so the transfer is spread over the whole frame
 * instead of following it; a per-rank flag signals completion.  The buffer holds world slabs of rows_max * width float4 (the layout
 * rt_resolve_gathered reads) followed by the flags.  One process per GPU: the owner exports the buffer with rt_ipc_export, the
 * other ranks map it with rt_ipc_open; one process over several devices (rt_create_multi): peer access, no handles needed. */
int rt_gather_buffer(rt_ctx* c, void** dev_ptr, uint64_t* stride_bytes, uint64_t* total_bytes)
{
    RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_gather_buffer");
    if (!dev_ptr || !stride_bytes) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_gather_buffer: null output");
    const size_t rows_max = ((size_t)c->height + c->world - 1) / c->world;
    const size_t stride = rows_max * c->width * 16;
    const size_t total = stride * c->world + 256;
    RT_CUDA(c, cudaSetDevice(c->device));
    if (c->own_gather && c->own_gather_stride != stride) { cudaFree(c->own_gather); c->own_gather = nullptr; }
    if (!c->own_gather)
    {
        RT_CUDA(c, cudaMalloc(&c->own_gather, total));
        RT_CUDA(c, cudaMemset(c->own_gather, 0, total));
        c->own_gather_stride = stride;
    }
    *dev_ptr = c->own_gather; *stride_bytes = stride;
    if (total_bytes) *total_bytes = total;
    return RT_OK;
}

int rt_ipc_export(const void* dev_ptr, void* handle64)
{
    if (!dev_ptr || !handle64) return RT_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr));
    if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("rt_ipc_export: ") + cudaGetErrorString(e); return RT_ERR_CUDA; }
    memcpy(handle64, &h, 64);
    return RT_OK;
}

int rt_ipc_open(rt_ctx* c, const void* handle64, void** dev_ptr)
{
    RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_ipc_open");
    if (!handle64 || !dev_ptr) RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_ipc_open: null argument");
    RT_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h; memcpy(&h, handle64, 64);
    void* p = nullptr;
    RT_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->ipc_opened.push_back(p);
    *dev_ptr = p;
    return RT_OK;
}

/* base = the buffer of rt_gather_buffer as THIS process addresses it (own pointer, peer pointer or rt_ipc_open mapping);
 * NULL turns the fused gather off.  Takes effect with the next frame. */
int rt_set_gather_target(rt_ctx* c, void* base, uint64_t stride_bytes)
{
    RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_set_gather_target");
    { int rc = join_shadow(c); if (rc) return rc; }
    if (!base) { c->gather_slab = nullptr; c->gather_flag = nullptr; return RT_OK; }
    const size_t rows_max = ((size_t)c->height + c->world - 1) / c->world;
    if (stride_bytes % 16 != 0 || stride_bytes < rows_max * c->width * 16)
        RT_FAIL(c, RT_ERR_INVALID_ARGUMENT, "rt_set_gather_target: slab stride must be a multiple of 16 and hold %zu rows", rows_max);
    c->gather_slab = (float4*)((char*)base + (size_t)c->rank * stride_bytes);
    c->gather_flag = (uint32_t*)((char*)base + (size_t)c->world * stride_bytes) + c->rank;
    return RT_OK;
}

/* On the context that presents (it must have a gather target itself, i.e. take part in the frames): makes its stream wait until
 * every rank has signalled the frame this context rendered last. */
int rt_gather_wait(rt_ctx* c)
{
    RT_CHECK_CTX(c); RT_NOT_ON_GROUP(c, "rt_gather_wait");
    if (!c->gather_flag) RT_FAIL(c, RT_ERR_NOT_READY, "rt_gather_wait: no gather target set");
    RT_CUDA(c, cudaSetDevice(c->device));
    k_gather_wait<<<1, 64, 0, c->stream>>>(c->gather_flag - c->rank, c->world, c->gather_frame);
    ++c->launches;
    return post_launch(c, "k_gather_wait");
}

/* Parity tap: out[i] = f(a[i], b[i]) for a function of include/rt_math.h, evaluated on `device` (host pointers; blocking).
 * fn: 0 sin 1 cos 2 tan 3 atan2(a, b) 4 acos 5 pow(a, b) 6 fmin 7 fmax 8 1/sqrt(a) 9 a / b. */
int rt_math_eval(int device, int fn, const float* a, const float* b, float* out, uint64_t n)
{
    if (!a || !b || !out || fn < 0 || fn > 9) return RT_ERR_INVALID_ARGUMENT;
    if (n == 0) return RT_OK;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); g_create_error = "rt_math_eval: no such CUDA device"; return RT_ERR_NO_DEVICE; }
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    cudaError_t e = cudaMalloc(&da, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&db, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dout, n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(da, a, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, b, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        k_math_eval<<<(unsigned)((n + 255) / 256), 256>>>(fn, da, db, dout, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dout);
    if (e != cudaSuccess) { g_create_error = std::string("rt_math_eval: ") + cudaGetErrorString(e); return RT_ERR_CUDA; }
    return RT_OK;
}

} // extern "C"
