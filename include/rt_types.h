/*
 * rt_types.h — byte-exact data contract between host producers, the C-ABI shim
 * (rt_b200.h), the CPU oracle and the CUDA kernels.
 *
 * Every struct here has the size and field offsets of the corresponding struct
 * in the reference's device/host shared header
 *   /root/reference/src/kernels/common/shared_structures.h:56-181
 * where the host `float3` is 16 bytes (x,y,z + hidden pad,
 *   /root/reference/src/mathlib/mathlib.hpp:40-77)
 * so that the arrays a reference `Scene` / `Bvh` produces can be handed to
 * rt_upload_scene() without conversion.  Plain C; no CUDA, no torch types.
 */
#ifndef RT_TYPES_H
#define RT_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 16-byte padded 3-vector == reference host float3 / OpenCL float3. */
typedef struct RtFloat3 { float x, y, z, w; } RtFloat3;
typedef struct RtFloat2 { float x, y; } RtFloat2;

/* shared_structures.h:56-59 — origin.w = t_min, direction.w = t_max */
typedef struct RtRay { RtFloat3 origin; RtFloat3 direction; } RtRay;

/* shared_structures.h:61-66 — primitive_id == 0xFFFFFFFF means miss;
 * bc and t are undefined for misses (trace_bvh.cl:135-138). */
typedef struct RtHit { RtFloat2 bc; uint32_t primitive_id; float t; } RtHit;

/* shared_structures.h:68-73 */
typedef struct RtSceneInfo {
    uint32_t analytic_light_count;
    uint32_t emissive_count;
    uint32_t environment_map_index;
    uint32_t padding;
} RtSceneInfo;

/* shared_structures.h:75-81 — array stride is 20 bytes. */
typedef struct RtPackedMaterial {
    uint32_t diffuse_albedo;                /* 24 bit RGB, 8 bit texture index */
    uint32_t specular_albedo;               /* 24 bit RGB, 8 bit texture index */
    uint32_t emission;                      /* RGBE */
    uint32_t roughness_metalness;           /* 8 rough, 8 tex, 8 metal, 8 tex  */
    uint32_t ior_emission_idx_transparency; /* 8 ior, 8 emis tex, 8 transp, 8 tex */
} RtPackedMaterial;

#define RT_LIGHT_TYPE_POINT 0u
#define RT_LIGHT_TYPE_DIRECTIONAL 1u

/* shared_structures.h:83-88 */
typedef struct RtLight {
    RtFloat3 origin;   /* point: position; directional: normalised direction */
    RtFloat3 radiance;
    uint32_t type;
    uint32_t padding[3];
} RtLight;

/* shared_structures.h:90-95 */
typedef struct RtTexture { int32_t data_start, width, height, padding; } RtTexture;

/* shared_structures.h:97-110 */
typedef struct RtVertex { RtFloat3 position, texcoord, normal; } RtVertex;

/* shared_structures.h:112-141 */
typedef struct RtTriangle {
    RtVertex v1, v2, v3;
    uint32_t mtlIndex;
    uint32_t padding[3];
} RtTriangle;

/* shared_structures.h:143-153 */
typedef struct RtRTTriangle { RtFloat3 position1, position2, position3; } RtRTTriangle;

/* shared_structures.h:160-171 — depth-first order (bvh.cpp:223-245).
 * leaf:     offset = first primitive, num_primitives_axis >> 16 = count (>0)
 * interior: first child = self+1, second child = offset,
 *           num_primitives_axis & 0xFFFF = split axis */
typedef struct RtLinearBVHNode {
    RtFloat3 bounds_min;
    RtFloat3 bounds_max;
    uint32_t offset;
    uint32_t num_primitives_axis;
    uint32_t padding[2];
} RtLinearBVHNode;

/* shared_structures.h:173-181 */
typedef struct RtCamera {
    RtFloat3 position, front, up;
    float fov, aspect_ratio, aperture, focus_distance;
} RtCamera;

/* kernels/common/constants.h:28-36 */
#define RT_MAX_RENDER_DIST 20000.0f
#define RT_EPS 1e-3f
#define RT_PI 3.14159265359f
#define RT_TWO_PI 6.28318530718f
#define RT_INV_PI 0.31830988618f
#define RT_INV_TWO_PI 0.15915494309f
#define RT_INVALID_ID 0xFFFFFFFFu
#define RT_INVALID_TEXTURE_IDX 0xFFu
/* blue-noise sampler tables (kernels/common/sampling.h:40-61; the reference ships them in utils/blue_noise_sampler.hpp) */
#define RT_BN_SOBOL_COUNT 65536      /* sobol_256spp_256d: 256 samples x 256 dimensions */
#define RT_BN_TILE_COUNT 131072      /* scramblingTile / rankingTile: 128 x 128 pixels x 8 dimensions */

#ifdef __cplusplus
}
#define RT_SA(T, n) static_assert(sizeof(T) == (n), #T " size mismatch with reference layout")
RT_SA(RtFloat3, 16); RT_SA(RtRay, 32); RT_SA(RtHit, 16); RT_SA(RtSceneInfo, 16);
RT_SA(RtPackedMaterial, 20); RT_SA(RtLight, 48); RT_SA(RtTexture, 16); RT_SA(RtVertex, 48);
RT_SA(RtTriangle, 160); RT_SA(RtRTTriangle, 48); RT_SA(RtLinearBVHNode, 48); RT_SA(RtCamera, 64);
#undef RT_SA
#endif

#endif /* RT_TYPES_H */
