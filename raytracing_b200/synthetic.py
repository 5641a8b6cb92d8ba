"""
Synthetic "Bistro-scale" scene of BASELINE.json configs[4] / SURVEY 8(d) C5: K copies of the ShaderBalls triangles
(183 copies = 10 026 570 triangles) on a 14 x 14 grid (first K cells), pitch = 1.1 x the model's XY extent, every copy
rotated about Z by an angle drawn from the LCG x = 1103515245 x + 12345 (seed 1234), same 13 materials and the
reference's directional light, built directly as a Triangle[] (no OBJ round trip).  The BVH is built on the host by
raytracing_b200/host/bvh.cpp (the build stays on the host, north_star).  The camera is raised and pulled back so
that the whole grid is in view; its exact pose is part of the returned scene dict (key "camera_pose").
"""
import math

import numpy as np

from .camera import default_camera
from .layouts import TRIANGLE_DT


def replicate(base_triangles: np.ndarray, copies: int = 183, grid: int = 14, seed: int = 1234):
    """Returns (triangles[copies * n], (xmin, xmax, ymin, ymax, zmax) of the whole field)."""
    t = np.ascontiguousarray(base_triangles, dtype=TRIANGLE_DT)
    pos = np.stack([t[v]["position"][:, :3] for v in ("v1", "v2", "v3")], axis=1).astype(np.float64)      # n,3,3
    nrm = np.stack([t[v]["normal"][:, :3] for v in ("v1", "v2", "v3")], axis=1).astype(np.float64)
    lo, hi = pos.reshape(-1, 3).min(0), pos.reshape(-1, 3).max(0)
    centre = (lo + hi) * 0.5
    pitch = 1.1 * max(hi[0] - lo[0], hi[1] - lo[1])
    out = np.zeros(copies * len(t), dtype=TRIANGLE_DT)
    x = seed
    for k in range(copies):
        x = (1103515245 * x + 12345) & 0xFFFFFFFF
        ang = (x / 4294967296.0) * 2.0 * math.pi
        c, s = math.cos(ang), math.sin(ang)
        rot = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        cell = np.array([(k % grid) * pitch, (k // grid) * pitch, 0.0])
        p = (pos - centre * [1, 1, 0]) @ rot.T + cell
        n = nrm @ rot.T
        blk = out[k * len(t):(k + 1) * len(t)]
        blk[:] = t
        for j, v in enumerate(("v1", "v2", "v3")):
            blk[v]["position"][:, :3] = p[:, j].astype(np.float32)
            blk[v]["normal"][:, :3] = n[:, j].astype(np.float32)
    rows = (copies + grid - 1) // grid
    half = pitch * 0.5
    bounds = (-half, (min(copies, grid) - 1) * pitch + half, -half, (rows - 1) * pitch + half, float(hi[2]))
    return out, bounds


def field_camera(width: int, height: int, bounds, pitch_angle: float = 2.0944) -> np.ndarray:
    """Camera above and in front of the field looking 30 degrees down at it (yaw = pi/2 as the default camera,
    pitch = pi/2 + 30 deg): the far edge of the field sits ~20 % below the top of the frame, the near edge ~20 %
    above the bottom."""
    xmin, xmax, ymin, ymax, zmax = bounds
    depth = ymax - ymin
    position = ((xmin + xmax) * 0.5, ymin - 0.125 * depth, 0.16 * depth + zmax)
    return default_camera(width, height, position=position, pitch=pitch_angle)


def bistro_scale_scene(base_scene: dict, copies: int = 183, width: int = 1920, height: int = 1080) -> dict:
    """Full scene dict (arrays in the reference layout) with the BVH built by the host builder."""
    from . import hostapi
    tris, bounds = replicate(base_scene["triangles"], copies)
    ordered, nodes, depth = hostapi.build_bvh(tris)
    scene = dict(base_scene)
    scene["triangles"], scene["nodes"] = ordered, nodes
    mats = np.ascontiguousarray(base_scene["materials"])
    emis = (mats["emission"][ordered["mtlIndex"]] >> 24) != 0          # RGBE exponent byte 0 <=> no emission (scene.cpp:87-103)
    scene["emissive"] = np.nonzero(emis)[0].astype("<u4")
    info = np.array(base_scene["scene_info"], copy=True)
    info["emissive_count"] = len(scene["emissive"])
    scene["scene_info"] = info
    scene["camera_pose"] = field_camera(width, height, bounds)
    scene["bvh_depth"] = depth
    return scene
