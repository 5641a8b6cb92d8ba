/*
 * types.hpp — host-side names of the shared structs.  The reference's host code uses
 * `Triangle`, `LinearBVHNode`, `Camera`, ... from kernels/common/shared_structures.h
 * (with the 16-byte `float3` of mathlib/mathlib.hpp:40-77); here they are aliases of
 * the byte-identical PODs in include/rt_types.h.
 */
#pragma once

#include <cstdint>
#include <cmath>

#include "rt_types.h"

namespace rt_host
{

using float3 = RtFloat3;        // x, y, z + 4 bytes of padding (w), as the reference's host float3
using float2 = RtFloat2;
using Ray = RtRay;
using Hit = RtHit;
using SceneInfo = RtSceneInfo;
using PackedMaterial = RtPackedMaterial;
using Light = RtLight;
using Texture = RtTexture;
using Vertex = RtVertex;
using Triangle = RtTriangle;
using RTTriangle = RtRTTriangle;
using LinearBVHNode = RtLinearBVHNode;
using Camera = RtCamera;

inline float3 make_float3(float x, float y, float z) { return float3{ x, y, z, 0.0f }; }
inline float3 operator+(const float3& a, const float3& b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(const float3& a, const float3& b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(const float3& a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float component(const float3& v, unsigned i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
// std::min / std::max semantics, as the reference's Min/Max (mathlib.hpp:125-133); these inline to minss/maxss
inline float fmin2(float a, float b) { return (b < a) ? b : a; }
inline float fmax2(float a, float b) { return (a < b) ? b : a; }
inline float3 vmin(const float3& a, const float3& b) { return make_float3(fmin2(a.x, b.x), fmin2(a.y, b.y), fmin2(a.z, b.z)); }
inline float3 vmax(const float3& a, const float3& b) { return make_float3(fmax2(a.x, b.x), fmax2(a.y, b.y), fmax2(a.z, b.z)); }
inline float3 cross(const float3& a, const float3& b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

} // namespace rt_host
