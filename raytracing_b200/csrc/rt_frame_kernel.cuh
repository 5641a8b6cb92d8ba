/*
 * rt_frame_kernel.cuh — the one-kernel frame (k_frame): every CTA an independent wavefront over its own pixels.
 * Part of the single translation unit rt_kernels.cu (included there, inside its anonymous namespace).
 */
#pragma once

// ---- whole-frame kernel (RT_OPT_FRAME_KERNEL, default) ---------------------------------------------------------
// ONE launch per frame.  Every CTA is an independent wavefront path tracer over its own subset of the partition's pixels
// (groups of 32 consecutive local pixels, dealt round-robin to the resident CTAs), with its own region of every queue:
// a path never leaves the CTA that generated its primary ray.  The per-bounce schedule is the one of rt_extend_shade /
// rt_shadow_accumulate (integrator.cpp:27-59 fused the same way):
//     T(b): closest-hit traversal of bounce b  +  shadow pass of bounce b-1   -> hit queue / miss queue
//     S(b): shading of the hit queue, then of the miss queue                   -> shadow queue, ray queue of bounce b+1
// but the queue cursors and counters live in SHARED memory and the phases are separated by __syncthreads(): no global
// atomics on the ray path, no grid-wide synchronisation, no launch boundary between phases (the ~10-15 us floor that 22
// dependent launches per frame cost a small multi-GPU partition), and the 4 CTAs of an SM are in different phases at any
// time, so a CTA waiting for its slowest warp at a barrier leaves the issue slots to its neighbours.  Primary rays are
// generated inside T(0) (no separate pass over the queue).  Results are bit-identical: every per-pixel quantity is a
// function of (pixel, sample index, bounce) only, and the order of the additions into a pixel's radiance is unchanged.
struct CtaFrame
{
    uint32_t cur_trace, cur_shade;           // work cursors of the current T / S phase
    uint32_t ext_n[2], shadow_n[2];          // rays entering bounce b (parity b & 1), shadow rays spawned by S(b) (parity b & 1)
    unsigned long long hm[2];                // hits (low word) and misses (high word) of T(b), parity b & 1
    unsigned long long emit[2];              // shadow rays (low word) and continuation rays (high word) spawned by S(b): slot reservation
    uint32_t n_emissive, n_unoccluded;
    uint32_t ended_n[2];                     // fused gather: pixels whose path ended at a hit in S(b) (no continuation), parity b & 1
};

template <int SMEM>
__global__ void __launch_bounds__(RT_FRAME_MAX_THREADS, 1) k_frame(FrameParams p, DevScene sc, int mode, Queues q, DevCounters* ctr, float4* radiance,
                                                              AovParams aov, uint32_t max_bounces, uint32_t slots_per_cta,
                                                              const __grid_constant__ FrameDyn dyn)
{
    extern __shared__ __align__(128) float4 s_bvh[];
    __shared__ uint64_t s_mbar;
    __shared__ CtaFrame s;
    p.dyn = &dyn;                                  // per-frame constants (sample index, camera) arrive as a kernel parameter
    if (SMEM == 1) tma_stage_bvh(s_bvh, sc, &s_mbar);
    if (SMEM == 2) tma_stage_top(s_bvh, sc.wnodes, sc.top_k, &s_mbar);
    const int lane = threadIdx.x & 31;
    const uint32_t base = blockIdx.x * slots_per_cta;          // this CTA's region of every queue: slots [base, base + slots_per_cta)
    // (queue pointers are re-read from the kernel parameters where they are used: nothing but `base` stays live across the
    // traversal and shading loops)
#define FQ(plane, i) (q.plane)[base + (i)]
    if (threadIdx.x == 0)
    {
        const uint32_t n_groups = (p.n_local + 31u) / 32u;
        const uint32_t cta = blockIdx.x, n_cta = gridDim.x;
        const uint32_t my_groups = cta < n_groups ? (n_groups - cta + n_cta - 1u) / n_cta : 0u;
        s.cur_trace = 0; s.cur_shade = 0; s.ext_n[0] = my_groups * 32u; s.ext_n[1] = 0; s.shadow_n[0] = s.shadow_n[1] = 0;
        s.hm[0] = s.hm[1] = 0ull; s.emit[0] = s.emit[1] = 0ull; s.n_emissive = 0; s.n_unoccluded = 0; s.ended_n[0] = s.ended_n[1] = 0;
        if (blockIdx.x == 0) ctr->n_primary = p.n_local;
    }
    __syncthreads();
    const uint32_t sample_idx = p.dyn->sample_idx;
    uint32_t nv = 0, nt = 0;     // (not counted here: RT_OPT_COUNT_TRAVERSAL uses the per-phase kernels)

    for (uint32_t b = 0; b <= max_bounces + 1u; ++b)
    {
        // ---------------------------------------------------------------- T(b): extension rays of bounce b, then shadow rays of bounce b-1
        {
            const int in = b & 1;
            const uint32_t n_ext = b <= max_bounces ? s.ext_n[in] : 0u;          // the last round is the shadow pass of the last bounce only
            const uint32_t ext_span = (n_ext + 31u) & ~31u;
            const uint32_t n_sh = b == 0 ? 0u : s.shadow_n[(b - 1u) & 1];
            const uint32_t total = ext_span + n_sh;
            if (threadIdx.x == 0)
            {   // state of the NEXT phases that nobody reads during this one
                s.cur_shade = 0; s.ext_n[(b + 1u) & 1] = 0; s.shadow_n[in] = 0; s.emit[in] = 0ull; s.ended_n[in] = 0;
            }
            for (;;)
            {
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(&s.cur_trace, 32u);
                at = __shfl_sync(0xffffffffu, at, 0);
                if (at >= total) break;
                if (at < ext_span)
                {
                    const uint32_t i = at + lane;
                    bool live = i < n_ext, hit = false;
                    float bu = 0.0f, bv = 0.0f, bt = 0.0f;
                    uint32_t prim = RT_INVALID_ID;
                    float4 a, bb;
                    if (b == 0)
                    {   // RayGeneration (raygeneration.cl:65-139) fused into the first traversal pass; slot i <-> local pixel li
                        const uint32_t li = (blockIdx.x + (i >> 5) * gridDim.x) * 32u + (uint32_t)lane;
                        live = li < p.n_local;
                        if (live)
                        {
                            const uint32_t px = li % p.width, py = (li / p.width) * p.world + p.rank;
                            f3 o, d;
                            generate_primary_ray(p.dyn->raygen, py * p.width + px, px, py, sample_idx, o, d);
                            a = make_float4(o.x, o.y, o.z, __uint_as_float(pack_pixel(px, py)));
                            bb = make_float4(d.x, d.y, d.z, RT_MAX_RENDER_DIST);
                            FQ(A[0], i) = a; FQ(B[0], i) = bb; FQ(C[0], i) = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
                            if (aov.enabled)
                            {   // raygeneration.cl:129-133
                                aov.albedo[li] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                aov.depth[li] = RT_MAX_RENDER_DIST;
                                aov.normal[li] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                aov.velocity[li] = make_float2(0.0f, 0.0f);
                            }
                        }
                    }
                    else if (live) { a = FQ(A[in], i); bb = FQ(B[in], i); }
                    if (live)
                    {
                        if (SMEM == 1) prim = trace_fast<false, false, 1, true>(sc, s_bvh, s_bvh + sc.wnodes_f4, mk3(a), mk3(bb), 0.0f, bb.w, bu, bv, bt, nv, nt);
                        else if (SMEM == 2) prim = trace_fast<false, false, 2, true>(sc, s_bvh, sc.wtris, mk3(a), mk3(bb), 0.0f, bb.w, bu, bv, bt, nv, nt);
                        else prim = trace<false, false>(sc, mode, mk3(a), mk3(bb), 0.0f, bb.w, bu, bv, bt, nv, nt);
                        hit = prim != RT_INVALID_ID;
                    }
                    const unsigned hmask = __ballot_sync(0xffffffffu, hit);
                    const unsigned mmask = __ballot_sync(0xffffffffu, live && !hit);
                    unsigned long long slot = 0ull;
                    if (lane == 0)
                        slot = atomicAdd(&s.hm[in], (unsigned long long)__popc(hmask) | ((unsigned long long)__popc(mmask) << 32));
                    slot = __shfl_sync(0xffffffffu, slot, 0);
                    const unsigned lt_mask = (1u << lane) - 1u;
                    if (hit) FQ(hitq, (uint32_t)slot + __popc(hmask & lt_mask)) = make_float4(bu, bv, __uint_as_float(prim), __uint_as_float(i));
                    else if (live) FQ(missq, (uint32_t)(slot >> 32) + __popc(mmask & lt_mask)) = i;
                }
                else
                {   // IntersectShadowRays + AccumulateDirectSamples of bounce b-1
                    const uint32_t i = at - ext_span + lane;
                    bool un = false;
                    if (i < n_sh)
                    {
                        const float4 a = FQ(sA, i), bb = FQ(sB, i);
                        float bu, bv, bt;
                        if (SMEM == 1) un = trace_fast<true, false, 1, true>(sc, s_bvh, s_bvh + sc.wnodes_f4, mk3(a), mk3(bb), 0.0f, bb.w, bu, bv, bt, nv, nt) == RT_INVALID_ID;
                        else if (SMEM == 2) un = trace_fast<true, false, 2, true>(sc, s_bvh, sc.wtris, mk3(a), mk3(bb), 0.0f, bb.w, bu, bv, bt, nv, nt) == RT_INVALID_ID;
                        else un = trace<true, false>(sc, mode, mk3(a), mk3(bb), 0.0f, bb.w, bu, bv, bt, nv, nt) == RT_INVALID_ID;
                        if (un)
                        {
                            const float4 c = FQ(sC, i);
                            const uint32_t li = local_index(p, __float_as_uint(a.w));
                            float4 r = radiance[li];
                            r.x += c.x; r.y += c.y; r.z += c.z;
                            radiance[li] = r;
                        }
                    }
                    const unsigned umask = __ballot_sync(0xffffffffu, un);
                    if (lane == 0 && umask) atomicAdd(&s.n_unoccluded, (uint32_t)__popc(umask));
                }
            }
            __syncthreads();
            if (p.gather && b > 0)
            {   // fused gather: the shadow rays of bounce b-1 are accumulated, so the pixels whose path ended at a hit of S(b-1) are
                // final; after the last round so are the paths that were still alive (their rays sit in the queue S(max) filled)
                const int pin = (b - 1u) & 1;
                const uint32_t n_end = s.ended_n[pin];
                const uint32_t* list = pin ? (const uint32_t*)q.hits + base : q.shadow_flags + base;
                for (uint32_t k = threadIdx.x; k < n_end; k += blockDim.x) { const uint32_t li = list[k]; p.gather[li] = radiance[li]; }
                if (b > max_bounces)
                {
                    const uint32_t n_alive = s.ext_n[in];
                    for (uint32_t k = threadIdx.x; k < n_alive; k += blockDim.x)
                    {
                        const uint32_t li = local_index(p, __float_as_uint(FQ(A[in], k).w));
                        p.gather[li] = radiance[li];
                    }
                }
            }
            if (threadIdx.x == 0)
            {   // per-bounce statistics (rt_read_frame_stats): one fire-and-forget global add per CTA and counter
                const unsigned long long hm = s.hm[in];
                if (b <= max_bounces && hm) atomicAdd((unsigned long long*)&ctr->hm[b], hm);
                if (b > 0 && s.n_unoccluded) atomicAdd(&ctr->n_unoccluded[b - 1u], s.n_unoccluded);
                s.n_unoccluded = 0;
            }
        }
        if (b > max_bounces) break;
        // ---------------------------------------------------------------- S(b): ShadeSurfaceHits over the hit queue, ShadeMissedRays over the miss queue
        {
            const int in = b & 1, out = (b + 1u) & 1;
            const unsigned long long hm = s.hm[in];
            const uint32_t n_hit = (uint32_t)hm, n_miss = (uint32_t)(hm >> 32);
            const uint32_t hit_span = (n_hit + 31u) & ~31u;            // warps never mix hits and misses
            const uint32_t total = hit_span + n_miss;
            if (threadIdx.x == 0) { s.cur_trace = 0; s.hm[out] = 0ull; }
            for (;;)
            {
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(&s.cur_shade, 32u);
                at = __shfl_sync(0xffffffffu, at, 0);
                if (at >= total) break;
                if (at < hit_span)
                {
                    const uint32_t k = at + lane;
                    const bool hit = k < n_hit;
                    uint32_t pixel = 0;
                    ShadeOut so;
                    so.emissive = so.spawn_next = so.spawn_shadow = false;
                    if (hit)
                    {
                        const float4 h = FQ(hitq, k);
                        const uint32_t i = __float_as_uint(h.w);
                        const float4 a = FQ(A[in], i), bb = FQ(B[in], i), c = FQ(C[in], i);
                        pixel = __float_as_uint(a.w);
                        shade_hit(sc, p, aov, b, pixel, mk3(a), mk3(bb), mk3(c), __float_as_uint(h.z), h.x, h.y, so);
                        if (so.emissive)
                        {
                            const uint32_t li = local_index(p, pixel);
                            float4 r = radiance[li];
                            r.x += so.emission_add.x; r.y += so.emission_add.y; r.z += so.emission_add.z;
                            radiance[li] = r;
                        }
                    }
                    const bool ss = hit && so.spawn_shadow, sn = hit && so.spawn_next;
                    const unsigned smask = __ballot_sync(0xffffffffu, ss), nmask = __ballot_sync(0xffffffffu, sn);
                    const unsigned emask = __ballot_sync(0xffffffffu, hit && so.emissive);
                    unsigned long long slot = 0ull;
                    if ((smask | nmask | emask) != 0u)
                    {
                        if (lane == 0)
                        {
                            slot = atomicAdd(&s.emit[in], (unsigned long long)__popc(smask) | ((unsigned long long)__popc(nmask) << 32));
                            if (emask) atomicAdd(&s.n_emissive, (uint32_t)__popc(emask));
                        }
                        slot = __shfl_sync(0xffffffffu, slot, 0);
                    }
                    const unsigned lt_mask = (1u << lane) - 1u;
                    if (p.gather)
                    {   // paths that end at this hit (no continuation ray): final once their shadow ray is accumulated in T(b+1)
                        const unsigned dmask = __ballot_sync(0xffffffffu, hit && !so.spawn_next);
                        if (dmask)
                        {
                            uint32_t at0 = 0;
                            if (lane == 0) at0 = atomicAdd(&s.ended_n[in], (uint32_t)__popc(dmask));
                            at0 = __shfl_sync(0xffffffffu, at0, 0);
                            if (hit && !so.spawn_next)
                            {
                                uint32_t* list = in ? (uint32_t*)q.hits + base : q.shadow_flags + base;
                                list[at0 + __popc(dmask & lt_mask)] = local_index(p, pixel);
                            }
                        }
                    }
                    if (ss)
                    {
                        const uint32_t si = (uint32_t)slot + __popc(smask & lt_mask);
                        FQ(sA, si) = make_float4(so.s_origin.x, so.s_origin.y, so.s_origin.z, __uint_as_float(pixel));
                        FQ(sB, si) = make_float4(so.s_dir.x, so.s_dir.y, so.s_dir.z, so.s_tmax);
                        FQ(sC, si) = make_float4(so.s_sample.x, so.s_sample.y, so.s_sample.z, 0.0f);
                    }
                    if (sn)
                    {
                        const uint32_t ni = (uint32_t)(slot >> 32) + __popc(nmask & lt_mask);
                        FQ(A[out], ni) = make_float4(so.n_origin.x, so.n_origin.y, so.n_origin.z, __uint_as_float(pixel));
                        FQ(B[out], ni) = make_float4(so.n_dir.x, so.n_dir.y, so.n_dir.z, RT_MAX_RENDER_DIST);
                        FQ(C[out], ni) = make_float4(so.n_throughput.x, so.n_throughput.y, so.n_throughput.z, 0.0f);
                    }
                }
                else
                {
                    const uint32_t k = at - hit_span + lane;
                    if (k < n_miss)
                    {
                        const uint32_t i = FQ(missq, k);
                        const float4 a = FQ(A[in], i), bb = FQ(B[in], i), c = FQ(C[in], i);
                        shade_miss(sc, p, radiance, __float_as_uint(a.w), mk3(bb), mk3(c));
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0)
            {
                const unsigned long long em = s.emit[in];
                s.shadow_n[in] = (uint32_t)em; s.ext_n[out] = (uint32_t)(em >> 32);
                if (em) atomicAdd((unsigned long long*)&ctr->emit[b], em);
                if (s.n_emissive) atomicAdd(&ctr->n_emissive[b], s.n_emissive);
                s.n_emissive = 0;
            }
            __syncthreads();
        }
    }
}
#undef FQ
