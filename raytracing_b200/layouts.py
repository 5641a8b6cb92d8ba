"""
numpy views of the byte-exact structs in include/rt_types.h (== the reference's
kernels/common/shared_structures.h:56-181 with 16-byte float3).
"""
import numpy as np

RAY_DT = np.dtype([("origin", "<f4", 4), ("direction", "<f4", 4)])
HIT_DT = np.dtype([("bc", "<f4", 2), ("primitive_id", "<u4"), ("t", "<f4")])
VERTEX_DT = np.dtype([("position", "<f4", 4), ("texcoord", "<f4", 4), ("normal", "<f4", 4)])
TRIANGLE_DT = np.dtype([("v1", VERTEX_DT), ("v2", VERTEX_DT), ("v3", VERTEX_DT), ("mtlIndex", "<u4"), ("padding", "<u4", 3)])
NODE_DT = np.dtype([("bounds_min", "<f4", 4), ("bounds_max", "<f4", 4), ("offset", "<u4"),
                    ("num_primitives_axis", "<u4"), ("padding", "<u4", 2)])
MATERIAL_DT = np.dtype([("diffuse_albedo", "<u4"), ("specular_albedo", "<u4"), ("emission", "<u4"),
                        ("roughness_metalness", "<u4"), ("ior_emission_idx_transparency", "<u4")])
LIGHT_DT = np.dtype([("origin", "<f4", 4), ("radiance", "<f4", 4), ("type", "<u4"), ("padding", "<u4", 3)])
TEXTURE_DT = np.dtype([("data_start", "<i4"), ("width", "<i4"), ("height", "<i4"), ("padding", "<i4")])
SCENE_INFO_DT = np.dtype([("analytic_light_count", "<u4"), ("emissive_count", "<u4"),
                          ("environment_map_index", "<u4"), ("padding", "<u4")])
CAMERA_DT = np.dtype([("position", "<f4", 4), ("front", "<f4", 4), ("up", "<f4", 4), ("fov", "<f4"),
                      ("aspect_ratio", "<f4"), ("aperture", "<f4"), ("focus_distance", "<f4")])

assert RAY_DT.itemsize == 32 and HIT_DT.itemsize == 16 and TRIANGLE_DT.itemsize == 160 and NODE_DT.itemsize == 48
assert MATERIAL_DT.itemsize == 20 and LIGHT_DT.itemsize == 48 and TEXTURE_DT.itemsize == 16
assert SCENE_INFO_DT.itemsize == 16 and CAMERA_DT.itemsize == 64

INVALID_ID = 0xFFFFFFFF

# name -> dtype of every array of a scene dump (the inputs of rt_upload_scene)
SCENE_ARRAYS = [("triangles", TRIANGLE_DT), ("nodes", NODE_DT), ("materials", MATERIAL_DT), ("lights", LIGHT_DT),
                ("textures", TEXTURE_DT), ("texels", np.dtype("<u4")), ("emissive", np.dtype("<u4")),
                ("env", np.dtype("<f4")), ("scene_info", SCENE_INFO_DT)]
