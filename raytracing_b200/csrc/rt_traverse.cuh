/*
 * rt_traverse.cuh — BVH staging (TMA bulk copies) and the traversal functions: literal reference-order traversal and the child-box layout traversal.
 * Part of the single translation unit rt_kernels.cu (included there, inside its anonymous namespace).
 */
#pragma once

// ------------------------------------------------------------------------------------ TMA staging of a small BVH
// When the whole traversal structure (interior records + triangle records) is small enough, every CTA of a
// traversal kernel copies it ONCE into shared memory with two TMA bulk copies (cp.async.bulk, completion signalled
// through an mbarrier transaction count) issued by one elected thread, and all node / triangle fetches of the
// kernel become shared-memory loads: the L1 data pipe is the second-busiest unit of the traversal kernels (ncu:
// l1tex data-pipe wavefronts ~58 % of peak, a divergent LDG.128 touches one 128-byte line per active lane), while a
// 16-byte LDS from 32 different records needs 4 conflict-free wavefronts.  The kernels are persistent, so the copy
// is amortised over every ray the CTA traces.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_stage_bvh(float4* dst, const DevScene& sc, uint64_t* mbar)
{
    const uint32_t bar = smem_u32(mbar);
    const uint32_t nodes_bytes = sc.wnodes_f4 * 16u, tris_bytes = sc.wtris_f4 * 16u;
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nodes_bytes + tris_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(sc.wnodes), "r"(nodes_bytes), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst + sc.wnodes_f4)), "l"(sc.wtris), "r"(tris_bytes), "r"(bar) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(0) : "memory");
}

// Scenes that do not fit: only the top of the tree — the first `n_records` interior records, breadth-first (rt_bvh_layout.h) — is
// staged, with one TMA bulk copy per CTA; deeper records and all triangles stay behind L1/L2 (RT_OPT_TOP_SMEM).
__device__ __forceinline__ void tma_stage_top(float4* dst, const float4* wnodes, uint32_t n_records, uint64_t* mbar)
{
    const uint32_t bar = smem_u32(mbar);
    const uint32_t bytes = n_records * 64u;
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(wnodes), "r"(bytes), "r"(bar) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(0) : "memory");
}

// Programmatic dependent launch (RT_OPT_PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start (launch its CTAs, stage the BVH) while the previous kernel of the stream is still draining; pdl_wait() blocks
// until that kernel has completed and its memory is visible, and is a no-op for a normal launch.  Every persistent
// kernel lets ITS dependent start as early as possible: its CTAs are all resident by then (persistent_grid), so
// the dependent's CTAs only take the slots that exiting CTAs free.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// SMEM: 0 records behind L1/L2 (ld.global.nc), 1 all records in shared memory, 2 the first sc.top_k interior records in shared
// memory and the rest in global memory (one generic load serves both)
template <int SMEM>
__device__ __forceinline__ float4 ld_bvh(const float4* p) { return SMEM ? *p : __ldg(p); }

// ------------------------------------------------------------------------------------ traversal
// Literal restatement of kernels/cl/trace_bvh.cl:99-211 on the reference node layout: per-ray
// DFS, 64-entry private stack, far child pushed unconditionally and box-tested when popped,
// near child chosen by ray_sign[axis], inclusive tests, later equal-t hit overwrites, back-face
// culling (det < 1e-8 rejects).  ANY = the -D SHADOW_RAYS variant (returns 0 on first hit).
template <bool ANY, bool COUNT>
__device__ __forceinline__ uint32_t trace_literal(const DevScene& sc, f3 o, f3 d, float t_min, float t_max,
                                                  float& bu, float& bv, float& bt, uint32_t& nv, uint32_t& nt)
{
    f3 inv = splat(1.0f) / d;
    int sx = inv.x < 0, sy = inv.y < 0, sz = inv.z < 0;
    uint32_t prim = RT_INVALID_ID;
    int to_visit = 0, cur = 0;
    int stack[64];
    for (;;)
    {
        float4 n0 = __ldg(sc.nodes_ref + (size_t)cur * 3), n1 = __ldg(sc.nodes_ref + (size_t)cur * 3 + 1), n2 = __ldg(sc.nodes_ref + (size_t)cur * 3 + 2);
        if (COUNT) ++nv;
        f3 t0 = (mk3(n0) - o) * inv, t1 = (mk3(n1) - o) * inv;
        float lo = fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z));
        float hi = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
        bool box = fminf(hi, t_max) >= fmaxf(lo, t_min);
        uint32_t offset = __float_as_uint(n2.x), npa = __float_as_uint(n2.y);
        if (box)
        {
            int nprims = (int)(npa >> 16);
            if (nprims > 0)
            {
                for (int i = 0; i < nprims; ++i)
                {
                    const float4* tp = sc.tris_ref + (size_t)(offset + i) * 3;
                    f3 p1 = mk3(__ldg(tp)), p2 = mk3(__ldg(tp + 1)), p3 = mk3(__ldg(tp + 2));
                    if (COUNT) ++nt;
                    f3 e1 = p2 - p1, e2 = p3 - p1;
                    f3 pvec = cross(d, e2);
                    float det = dot(e1, pvec);
                    if (det < 1e-8f || -det > 1e-8f) continue;
                    float inv_det = 1.0f / det;
                    f3 tvec = o - p1;
                    float u = dot(tvec, pvec) * inv_det;
                    if (u < 0.0f || u > 1.0f) continue;
                    f3 qvec = cross(tvec, e1);
                    float v = dot(d, qvec) * inv_det;
                    if (v < 0.0f || u + v > 1.0f) continue;
                    float t = dot(e2, qvec) * inv_det;
                    if (t < t_min || t > t_max) continue;
                    bu = u; bv = v; bt = t;
                    prim = offset + i;
                    t_max = t;
                    if (ANY) return 0u;
                }
                if (to_visit == 0) break;
                cur = stack[--to_visit];
            }
            else
            {
                uint32_t axis = npa & 0xFFFFu;
                int s = axis == 0 ? sx : (axis == 1 ? sy : sz);
                if (s) { stack[to_visit++] = cur + 1; cur = (int)offset; }
                else   { stack[to_visit++] = (int)offset; cur = cur + 1; }
            }
        }
        else
        {
            if (to_visit == 0) break;
            cur = stack[--to_visit];
        }
    }
    return prim;
}

// Optimised traversal on the child-box node layout (rt_bvh_layout.h).  Same visiting order and
// the same arithmetic per box / triangle test as trace_literal, so results are bit-identical
// for finite rays; non-finite rays (NaN/inf components; their traversal is garbage-in but must
// still match) take the literal path.
// PIN: the whole-frame kernel shares its register budget with the shading code and ptxas then re-derives the sign bits on
// every step and recomputes the determinant after its branch; an empty asm makes both values opaque (kept in registers).
template <bool ANY, bool COUNT, int SMEM, bool PIN = false>
__device__ __forceinline__ uint32_t trace_fast(const DevScene& sc, const float4* wnodes, const float4* wtris, f3 o, f3 d, float t_min, float t_max,
                                               float& bu, float& bv, float& bt, uint32_t& nv, uint32_t& nt)
{
    float fin = ((o.x + o.y) + o.z) + ((d.x + d.y) + d.z);
    if (!(fabsf(fin) <= 3.0e38f) || COUNT)
        return trace_literal<ANY, COUNT>(sc, o, d, t_min, t_max, bu, bv, bt, nv, nt);

    f3 inv = splat(1.0f) / d;
    const bool sx = inv.x < 0, sy = inv.y < 0, sz = inv.z < 0;
    uint32_t sign_bits = (sx ? 1u : 0u) | (sy ? 2u : 0u) | (sz ? 4u : 0u);
    if (PIN) asm volatile("" : "+r"(sign_bits));
    uint32_t prim = RT_INVALID_ID;
    int sp = 0;
    int cur = sc.root_ref;
    // Deferred far children: (node reference, entry distance).  Two code shapes, chosen by where the BVH records live
    // (measured, same results): with the records in shared memory the kernel is purely issue-bound and the packed
    // 64-bit stack entry + the sign-bit axis test win (CornellBox frame -2.7 %); with the records behind L1/L2 the two
    // 32-bit arrays (a discarded pop costs one load) and predicate selects are faster (ShaderBalls +1 %, Dragon +3.5 %).
#ifdef RT_SMEM_STACK
    // experiment: the traversal stack in shared memory (entry i of thread t at [i * 256 + t]: conflict-free), per-phase kernels only
    extern __shared__ __align__(128) float4 rt_dyn_smem[];
    int2* const sstk = (int2*)((char*)rt_dyn_smem + sc.stack_off) + threadIdx.x;
    constexpr bool SSTK = !PIN;
#else
    int2* const sstk = nullptr;
    constexpr bool SSTK = false;
#endif
    int2 stack[(SMEM == 1 && !SSTK) ? 64 : 1];
    int stack_ref[(SMEM == 1 || SSTK) ? 1 : 64];
    float stack_t[(SMEM == 1 || SSTK) ? 1 : 64];
    if (cur < 0)
    {   // single-leaf tree: the root box is tested like any visited node
        float4 r0 = __ldg(sc.nodes_ref), r1 = __ldg(sc.nodes_ref + 1);
        f3 t0 = (mk3(r0) - o) * inv, t1 = (mk3(r1) - o) * inv;
        float lo = fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z));
        float hi = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
        if (!(fminf(hi, t_max) >= fmaxf(lo, t_min))) return prim;
    }
    else
    {   // root box test (the reference tests every node it visits, including the root)
        float4 r0 = __ldg(sc.nodes_ref), r1 = __ldg(sc.nodes_ref + 1);
        f3 t0 = (mk3(r0) - o) * inv, t1 = (mk3(r1) - o) * inv;
        float lo = fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z));
        float hi = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
        if (!(fminf(hi, t_max) >= fmaxf(lo, t_min))) return prim;
    }
    for (;;)
    {
        while (cur >= 0)
        {
            // SMEM == 2: `wnodes` is the staged top of the tree (records [0, sc.top_k)), deeper records come from sc.wnodes
            const float4* np = (SMEM == 2 && (uint32_t)cur >= sc.top_k) ? sc.wnodes + (size_t)cur * 4 : wnodes + (size_t)cur * 4;
            float4 a = ld_bvh<SMEM>(np), b = ld_bvh<SMEM>(np + 1), c = ld_bvh<SMEM>(np + 2), m = ld_bvh<SMEM>(np + 3);
            // child 0 box: min (a.x,a.y,a.z) max (a.w,b.x,b.y); child 1 box: min (b.z,b.w,c.x) max (c.y,c.z,c.w)
#ifdef RT_FMA_TRAVERSAL
            // experiment (NOT bit-exact): contracted slab test, plane * inv - origin * inv in one FFMA per plane
            const f3 noi = mk3(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
            f3 t00 = mk3(__fmaf_rn(a.x, inv.x, noi.x), __fmaf_rn(a.y, inv.y, noi.y), __fmaf_rn(a.z, inv.z, noi.z));
            f3 t01 = mk3(__fmaf_rn(a.w, inv.x, noi.x), __fmaf_rn(b.x, inv.y, noi.y), __fmaf_rn(b.y, inv.z, noi.z));
            f3 t10 = mk3(__fmaf_rn(b.z, inv.x, noi.x), __fmaf_rn(b.w, inv.y, noi.y), __fmaf_rn(c.x, inv.z, noi.z));
            f3 t11 = mk3(__fmaf_rn(c.y, inv.x, noi.x), __fmaf_rn(c.z, inv.y, noi.y), __fmaf_rn(c.w, inv.z, noi.z));
#else
            f3 t00 = (mk3(a.x, a.y, a.z) - o) * inv, t01 = (mk3(a.w, b.x, b.y) - o) * inv;
            f3 t10 = (mk3(b.z, b.w, c.x) - o) * inv, t11 = (mk3(c.y, c.z, c.w) - o) * inv;
#endif
            float lo0 = fmaxf(fmaxf(fmaxf(fminf(t00.x, t01.x), fminf(t00.y, t01.y)), fminf(t00.z, t01.z)), t_min);
            float hi0 = fminf(fminf(fmaxf(t00.x, t01.x), fmaxf(t00.y, t01.y)), fmaxf(t00.z, t01.z));
            float lo1 = fmaxf(fmaxf(fmaxf(fminf(t10.x, t11.x), fminf(t10.y, t11.y)), fminf(t10.z, t11.z)), t_min);
            float hi1 = fminf(fminf(fmaxf(t10.x, t11.x), fmaxf(t10.y, t11.y)), fmaxf(t10.z, t11.z));
            bool h0 = fminf(hi0, t_max) >= lo0, h1 = fminf(hi1, t_max) >= lo1;
            int r0 = __float_as_int(m.x), r1 = __float_as_int(m.y);
            uint32_t axis = __float_as_uint(m.z);
            bool swap = SMEM == 1 ? ((sign_bits >> axis) & 1u) != 0u : (axis == 0 ? sx : (axis == 1 ? sy : sz));   // near child = second iff inv_dir[axis] < 0
            int near_ref = swap ? r1 : r0, far_ref = swap ? r0 : r1;
            bool near_hit = swap ? h1 : h0, far_hit = swap ? h0 : h1;
            float far_lo = swap ? lo0 : lo1;
            if (near_hit)
            {
                if (far_hit)
                {
                    if (SSTK) sstk[sp * 256] = make_int2(far_ref, __float_as_int(far_lo));
                    else if (SMEM == 1) stack[sp] = make_int2(far_ref, __float_as_int(far_lo));
                    else { stack_ref[sp] = far_ref; stack_t[sp] = far_lo; }
                    ++sp;
                }
                cur = near_ref;
            }
            else if (far_hit) cur = far_ref;
            else
            {   // pop: a pushed far child is re-tested against the (possibly shrunk) t_max, as the
                // reference does when it pops it; its slab interval was already valid at push time
                bool found = false;
                while (sp > 0)
                {
                    --sp;
                    if (SSTK) { int2 e = sstk[sp * 256]; if (t_max >= __int_as_float(e.y)) { cur = e.x; found = true; break; } }
                    else if (SMEM == 1) { int2 e = stack[sp]; if (t_max >= __int_as_float(e.y)) { cur = e.x; found = true; break; } }
                    else if (t_max >= stack_t[sp]) { cur = stack_ref[sp]; found = true; break; }
                }
                if (!found) return prim;
            }
        }
        // leaf: triangles [~cur ...] until the end-of-leaf flag
        uint32_t ti = (uint32_t)(~cur);
        for (;;)
        {
            const float4* tp = wtris + (size_t)ti * 3;
            float4 q0 = ld_bvh<(SMEM == 1)>(tp), q1 = ld_bvh<(SMEM == 1)>(tp + 1), q2 = ld_bvh<(SMEM == 1)>(tp + 2);
            f3 p1 = mk3(q0.x, q0.y, q0.z), e1 = mk3(q0.w, q1.x, q1.y), e2 = mk3(q1.z, q1.w, q2.x);
            bool last = __float_as_uint(q2.y) != 0u;
#ifdef RT_FMA_TRAVERSAL
#define RT_CROSS(a, b) mk3(__fmaf_rn((a).y, (b).z, -((a).z * (b).y)), __fmaf_rn((a).z, (b).x, -((a).x * (b).z)), __fmaf_rn((a).x, (b).y, -((a).y * (b).x)))
#define RT_DOT(a, b) __fmaf_rn((a).x, (b).x, __fmaf_rn((a).y, (b).y, (a).z * (b).z))
#else
#define RT_CROSS(a, b) cross(a, b)
#define RT_DOT(a, b) dot(a, b)
#endif
            f3 pvec = RT_CROSS(d, e2);
            float det = RT_DOT(e1, pvec);
            if (PIN) asm volatile("" : "+f"(det));
            if (!(det < 1e-8f || -det > 1e-8f))
            {
                float inv_det = 1.0f / det;
                f3 tvec = o - p1;
                float u = RT_DOT(tvec, pvec) * inv_det;
                if (!(u < 0.0f || u > 1.0f))
                {
                    f3 qvec = RT_CROSS(tvec, e1);
                    float v = RT_DOT(d, qvec) * inv_det;
                    if (!(v < 0.0f || u + v > 1.0f))
                    {
                        float t = RT_DOT(e2, qvec) * inv_det;
                        if (!(t < t_min || t > t_max))
                        {
                            bu = u; bv = v; bt = t; prim = ti; t_max = t;
                            if (ANY) return 0u;
                        }
                    }
                }
            }
            if (last) break;
            ++ti;
        }
        bool found = false;
        while (sp > 0)
                {
                    --sp;
                    if (SSTK) { int2 e = sstk[sp * 256]; if (t_max >= __int_as_float(e.y)) { cur = e.x; found = true; break; } }
                    else if (SMEM == 1) { int2 e = stack[sp]; if (t_max >= __int_as_float(e.y)) { cur = e.x; found = true; break; } }
                    else if (t_max >= stack_t[sp]) { cur = stack_ref[sp]; found = true; break; }
                }
        if (!found) return prim;
    }
}

template <bool ANY, bool COUNT>
__device__ __forceinline__ uint32_t trace(const DevScene& sc, int mode, f3 o, f3 d, float t_min, float t_max,
                                          float& bu, float& bv, float& bt, uint32_t& nv, uint32_t& nt)
{
    if (mode == 0) return trace_literal<ANY, COUNT>(sc, o, d, t_min, t_max, bu, bv, bt, nv, nt);
    return trace_fast<ANY, COUNT, 0>(sc, sc.wnodes, sc.wtris, o, d, t_min, t_max, bu, bv, bt, nv, nt);
}
