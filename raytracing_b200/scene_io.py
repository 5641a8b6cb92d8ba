"""
Scene dumps: every array rt_upload_scene() takes (triangles in BVH leaf order, nodes,
materials, lights, textures, texels, emissive list, scene info) in the reference's byte
layout, stored as an xz-compressed .npz so that the oracle, oracle/_ref and the CUDA
backend all consume IDENTICAL bytes (the BVH triangle order depends on the C++ library's
std::partition/nth_element).  The environment map is shared by all scenes
(scene.cpp:360 always loads assets/ibl/CGSkies_0036_free.hdr) and lives in its own file.
"""
import io
import lzma
import os

import numpy as np

from .layouts import SCENE_ARRAYS

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE_DIR = os.path.join(REPO, "tests", "golden", "scenes")
ENV_NAME = "env_CGSkies_0036"


def save_npz_xz(path: str, arrays: dict, preset: int = 9):
    bio = io.BytesIO()
    np.savez(bio, **arrays)
    with open(path, "wb") as fh:
        fh.write(lzma.compress(bio.getvalue(), preset=preset))


def load_npz_xz(path: str) -> dict:
    with open(path, "rb") as fh:
        raw = lzma.decompress(fh.read())
    with np.load(io.BytesIO(raw), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def save_scene(path: str, scene: dict):
    arrays = {k: np.ascontiguousarray(scene[k], dtype=dt) for k, dt in SCENE_ARRAYS if k != "env"}
    save_npz_xz(path, arrays)


def save_env(path: str, env: np.ndarray, width: int, height: int):
    e = np.ascontiguousarray(env, dtype="<f4").reshape(height, width, 4)
    assert (e[..., 3] == 0).all()        # LoadHDR leaves alpha untouched (hdr_loader.cpp:109-120)
    save_npz_xz(path, {"rgb": e[..., :3].copy(), "size": np.array([width, height], dtype="<u4")}, preset=6)


def load_env(path: str = None):
    d = load_npz_xz(path or os.path.join(SCENE_DIR, ENV_NAME + ".npz.xz"))
    w, h = int(d["size"][0]), int(d["size"][1])
    env = np.zeros((h, w, 4), dtype="<f4")
    env[..., :3] = d["rgb"]
    return env.reshape(-1), w, h


def load_scene(name_or_path: str, with_env: bool = True) -> dict:
    """name: CornellBox | ShaderBalls | CornellBox_Dragon (fixtures made from the reference's assets by
    tests/golden/make_fixtures.py) or a path to a .npz.xz dump."""
    path = name_or_path if os.path.exists(name_or_path) else os.path.join(SCENE_DIR, name_or_path + ".npz.xz")
    d = load_npz_xz(path)
    scene = {k: np.ascontiguousarray(d[k], dtype=dt) for k, dt in SCENE_ARRAYS if k != "env"}
    if with_env:
        scene["env"], scene["env_width"], scene["env_height"] = load_env()
    else:
        scene["env"], scene["env_width"], scene["env_height"] = np.zeros(4, dtype="<f4"), 1, 1
    return scene
