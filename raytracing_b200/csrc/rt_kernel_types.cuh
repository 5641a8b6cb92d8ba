/*
 * rt_kernel_types.cuh — device-side state of the wavefront path tracer: per-bounce counters, queue planes, per-frame constants, warp helpers.
 * Part of the single translation unit rt_kernels.cu (included there, inside its anonymous namespace).
 */
#pragma once

// ------------------------------------------------------------------------------------ device state
struct DevCounters
{
    uint32_t n_primary;                          // rays entering bounce 0
    uint32_t pad0;
    // rays spawned by the shading pass of bounce b: shadow rays and continuation rays (= the rays entering bounce b+1).
    // Adjacent + 8-byte aligned so that ONE 64-bit atomic reserves slots in both output queues.
    struct alignas(8) Emit { uint32_t shadow, next; };
    Emit emit[RT_MAX_BOUNCES + 1];
    struct alignas(8) HitMiss { uint32_t hit, miss; };   // adjacent + 8-byte aligned: one 64-bit atomic advances both
    HitMiss hm[RT_MAX_BOUNCES + 1];              // hit-queue entries / misses of bounce b
    uint32_t n_emissive[RT_MAX_BOUNCES + 1];
    uint32_t n_unoccluded[RT_MAX_BOUNCES + 1];
    uint32_t work_ext[RT_MAX_BOUNCES + 1];       // persistent-kernel work cursors
    uint32_t work_shade[RT_MAX_BOUNCES + 1];
    uint32_t work_shadow[RT_MAX_BOUNCES + 1];
    unsigned long long nodes_ext[RT_MAX_BOUNCES + 1], tris_ext[RT_MAX_BOUNCES + 1];
    unsigned long long nodes_shadow[RT_MAX_BOUNCES + 1], tris_shadow[RT_MAX_BOUNCES + 1];
};

__host__ __device__ __forceinline__ const uint32_t* in_count_ptr(const DevCounters* c, uint32_t bounce)
{
    return bounce == 0 ? &c->n_primary : &c->emit[bounce - 1].next;
}

struct Queues
{
    float4* A[2]; float4* B[2]; float4* C[2];
    float4* sA; float4* sB; float4* sC;
    float4* hits;
    uint32_t* shadow_flags;
    float4* hitq;
    uint32_t* missq;
#ifdef RT_HITQ_CARRY
    float4* hA; float4* hB; float4* hC;      // experiment: the hit queue carries its ray (coalesced reads in the shading kernel)
#endif
};

// Per-frame constants that change from frame to frame (sample index, camera): kept in a small device buffer that a
// 1-thread kernel refreshes at the start of every frame, so that the rest of the frame's launches have frame-invariant
// arguments and the whole frame can be replayed as ONE CUDA graph (rt_integrate) with a single node-parameter update.
struct FrameDyn;

struct FrameParams
{
    uint32_t width, height, rank, world, n_local;
    int white_furnace;
    const int* bn;                 // blue-noise sampler tables (kBlueNoise) or nullptr (kRandom)
    const FrameDyn* dyn;
    float4* gather;                // k_frame only: this rank's slab of the gathered radiance on the presenting device (peer memory,
                                   // rt_set_gather_target) — a pixel's radiance is pushed there the moment its path ends; else nullptr
};

// AOV outputs of bounce 0 (kernels/cl/aov.cl:44-110), written by the bounce-0 shading pass when enabled
struct AovCam { f3 position, front, up, right; float angle, aspect_ratio; };
struct AovParams
{
    int enabled;
    float4* albedo; float* depth; float4* normal; float2* velocity;
};

struct FrameDyn
{
    uint32_t sample_idx, pad[3];
    RayGenConsts raygen;
    AovCam cam, prev;
};

__global__ void k_set_frame(FrameDyn* dst, FrameDyn value) { *dst = value; }

// kernels/cl/aov.cl:30-42
__device__ __forceinline__ f2 project_screen(f3 position, const AovCam& c)
{
    f3 d = normalize(position - c.position);
    f3 ipd = d / dot(c.front, d);
    float u = dot(c.right, ipd) / (c.angle * c.aspect_ratio);
    float v = dot(c.up, ipd) / (c.angle);
    f2 r; r.x = u * 0.5f + 0.5f; r.y = v * 0.5f + 0.5f;
    return r;
}

__device__ __forceinline__ uint32_t pack_pixel(uint32_t px, uint32_t py) { return px | (py << 16); }
__device__ __forceinline__ uint32_t local_index(const FrameParams& p, uint32_t pxy)
{
    uint32_t px = pxy & 0xFFFFu, py = pxy >> 16;
    return (p.world == 1 ? py : py / p.world) * p.width + px;
}

__device__ __forceinline__ void warp_count(uint32_t* counter, bool pred)
{
    unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask != 0 && (threadIdx.x & 31) == __ffs(mask) - 1) atomicAdd(counter, (uint32_t)__popc(mask));
}

__device__ __forceinline__ void warp_sum64(unsigned long long* counter, uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counter, (unsigned long long)v);
}
