"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

from raytracing_b200 import scene_io

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")

_scene_cache = {}


def scene(name):
    if name not in _scene_cache:
        _scene_cache[name] = scene_io.load_scene(name)
    return _scene_cache[name]


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "golden_*.npz.xz")))


def load_golden(path):
    g = scene_io.load_npz_xz(path)
    base = os.path.basename(path)[len("golden_"):]
    # file name: golden_<scene>_<w>x<h>_b<mb>[_wf][_<tag>].npz.xz
    stem = base.replace(".npz.xz", "")
    parts = stem.split("_")
    size = max(i for i, p in enumerate(parts) if "x" in p and p.replace("x", "").isdigit())
    g["scene_name"] = "_".join(parts[:size])
    return g


def bits(a):
    """Bit pattern of a float32 array for exact comparisons.  NaNs are canonicalised first: the payload/sign of a NaN
    produced by an invalid operation is not specified by IEEE 754 (x86 gives 0xFFC00000, sm_100a 0x7FFFFFFF)."""
    u = np.ascontiguousarray(a, dtype="<f4").copy().view("<u4")
    u[(u & 0x7FFFFFFF) > 0x7F800000] = 0x7FC00000
    return u


def synthetic_sampler_tables(seed=20260923):
    """Seeded stand-ins for the blue-noise sampler's tables (same shapes and value ranges as the reference's
    utils/blue_noise_sampler.hpp, which is reference data and not committed here): any tables exercise the same lookups."""
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, 65536, dtype=np.int32), rng.integers(0, 256, 131072, dtype=np.int32),
            rng.integers(0, 256, 131072, dtype=np.int32))


def reference_sampler_tables():
    """The reference's own tables, read out of oracle/_ref/libref.so when that library was built; else None."""
    from oracle import refbind
    if not refbind.available():
        return None
    return refbind.RefRenderer().sampler_tables()
