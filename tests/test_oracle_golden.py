"""
CPU tests (-m "not gpu"): the oracle restatement (oracle/oracle.cpp) against the golden
vectors committed under tests/golden/, which are outputs of the reference's own kernels
(oracle/_ref, see tests/golden/make_fixtures.py).  Bar: BIT-EXACT — primitive ids,
per-bounce ray counters and every radiance float.
"""
import numpy as np
import pytest

from oracle.orcbind import Oracle
from tests.helpers import bits, golden_files, load_golden, scene


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("golden_")[-1].replace(".npz.xz", ""))
def test_oracle_matches_reference_golden(path):
    g = load_golden(path)
    o = Oracle(scene(g["scene_name"]))
    w, h, mb = int(g["width"]), int(g["height"]), int(g["max_bounces"])
    rad = np.zeros((h, w, 4), dtype="<f4")
    for sample in range(int(g["sample_count"])):          # progressive fixtures: hits and counters are the last sample's
        rad, hits, st = o.render(g["camera"], w, h, mb, sample_idx=sample, white_furnace=bool(g["white_furnace"]), radiance=rad)
    assert np.array_equal(hits["primitive_id"], g["primitive_id"])
    for k in ("n_ext", "n_miss", "n_hit", "n_shadow", "n_cont", "n_unoccluded"):
        assert np.array_equal(st[k][: mb + 1], g[k]), k
    assert np.array_equal(bits(rad[..., :3]), bits(g["radiance_rgb"]))
    if "hit_bc" in g:
        hit = g["primitive_id"] != 0xFFFFFFFF
        assert np.array_equal(bits(hits["bc"][hit]), bits(g["hit_bc"][hit]))
        assert np.array_equal(bits(hits["t"][hit]), bits(g["hit_t"][hit]))
        rays = o.generate_rays(g["camera"], w, h)
        assert np.array_equal(bits(rays["origin"]), bits(g["ray_origin"]))
        assert np.array_equal(bits(rays["direction"]), bits(g["ray_direction"]))


def test_white_furnace_bounded(golden_dir):
    """Energy conservation aid of the reference: with all albedos 1, no emission and a 0.5 sky,
    one-sample radiance estimates stay finite and the image mean stays at or below ~0.5 grey."""
    import glob, os
    path = glob.glob(os.path.join(golden_dir, "golden_*_wf.npz.xz"))[0]
    g = load_golden(path)
    rgb = g["radiance_rgb"]
    assert np.isfinite(rgb).all()
    assert rgb.mean() <= 0.55


def test_rng_streams_are_exact_integer_functions():
    """WangHash (utils.h:113-121) and SampleRandom (sampling.h:64-82) restated in pure Python."""
    o = Oracle(scene("CornellBox"))

    def wang(x):
        x = ((x ^ 61) ^ (x >> 16)) & 0xFFFFFFFF
        x = (x + (x << 3)) & 0xFFFFFFFF
        x = x ^ (x >> 4)
        x = (x * 0x27d4eb2d) & 0xFFFFFFFF
        return x ^ (x >> 15)
    rng = np.random.default_rng(7)
    for x in [0, 1, 61, 0xFFFFFFFF, 0x80000000] + list(rng.integers(0, 2**32, 200)):
        assert o.wang_hash(int(x)) == wang(int(x))
    for (px, py, s, b, t) in rng.integers(0, 4096, size=(200, 5)):
        seed = wang(int(px))
        seed = wang((seed + wang(int(py))) & 0xFFFFFFFF)
        seed = wang((seed + wang(int(s))) & 0xFFFFFFFF)
        seed = wang((seed + wang(int(b) * 5 + int(t) % 5)) & 0xFFFFFFFF)
        expect = np.float32(seed) * np.float32(2.3283064365386963e-10)
        assert o.sample_random(int(px), int(py), int(s), int(b), int(t) % 5) == expect
