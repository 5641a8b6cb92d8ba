"""
Host-side C++ (raytracing_b200/host: Scene loader, SAH BVH builder, HDR reader, Integrator mirror,
CUDAPathTraceIntegrator, headless Render).  CPU tests run everywhere; the Render tests need a GPU.
"""
import os

import numpy as np
import pytest

from oracle.orcbind import Oracle
from raytracing_b200 import hostapi, scene_io
from raytracing_b200.camera import default_camera
from tests.helpers import bits, scene, synthetic_sampler_tables

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROC_OBJ = os.path.join(REPO, "tests", "golden", "scenes", "procedural.obj")
REF_ASSETS = "/root/reference/assets"


def write_hdr(path, rgbe, rle=True):
    """Minimal Radiance .hdr writer (new-style RLE scanlines of literal chunks, or flat pixels)."""
    h, w, _ = rgbe.shape
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        for y in range(h):
            if rle and 8 <= w <= 32767:
                f.write(bytes([2, 2, w >> 8, w & 255]))
                for c in range(4):
                    x = 0
                    while x < w:
                        n = min(128, w - x)
                        if n >= 3 and (x // 128) % 2 == 1:      # alternate: a run of the first value, then literals
                            f.write(bytes([128 + 2, int(rgbe[y, x, c])]))
                            row = rgbe[y, x:x + 2, c].copy(); row[:] = rgbe[y, x, c]
                            rgbe[y, x:x + 2, c] = row
                            x += 2
                            continue
                        f.write(bytes([n])); f.write(rgbe[y, x:x + n, c].astype(np.uint8).tobytes())
                        x += n
            else:
                f.write(rgbe[y].astype(np.uint8).tobytes())
    return rgbe


def decode_rgbe(rgbe):
    """hdr_loader.cpp:102-120: (mantissa / 256) * 2^(E - 128), alpha 0."""
    e = rgbe[..., 3].astype(np.int32) - 128
    d = np.ldexp(np.float32(1.0), e).astype(np.float32)
    out = np.zeros(rgbe.shape[:2] + (4,), dtype=np.float32)
    for c in range(3):
        out[..., c] = (rgbe[..., c].astype(np.float32) / np.float32(256.0)) * d
    return out


def make_env(tmp_path, w=48, h=24, rle=True):
    rng = np.random.default_rng(3)
    rgbe = rng.integers(0, 256, size=(h, w, 4)).astype(np.int64)
    rgbe[..., 3] = rng.integers(120, 134, size=(h, w))
    path = str(tmp_path / f"env_{w}x{h}_{int(rle)}.hdr")
    rgbe = write_hdr(path, rgbe, rle)
    return path, decode_rgbe(rgbe)


def test_default_camera_matches_python_mirror():
    for (w, h) in [(1920, 1080), (256, 256), (3840, 2160), (333, 127)]:
        assert hostapi.default_camera(w, h).tobytes() == default_camera(w, h).tobytes()


@pytest.mark.parametrize("w,rle", [(48, True), (300, True), (5, False), (48, False)])
def test_hdr_reader(tmp_path, w, rle):
    path, expect = make_env(tmp_path, w=w, h=7, rle=rle)
    s = hostapi.HostScene(PROC_OBJ)
    s.add_directional_light((0, 0, 1), (1, 1, 1))
    s.finalize(env_path=path)
    a = s.arrays()
    assert (a["env_width"], a["env_height"]) == (w, 7)
    assert np.array_equal(bits(a["env"].reshape(7, w, 4)), bits(expect))
    s.close()


def test_procedural_obj_loads_and_bvh_is_a_valid_tree():
    s = hostapi.HostScene(PROC_OBJ)
    s.add_directional_light((-0.6, -1.5, 3.5), (15, 10, 5))
    s.build_bvh()
    s.finalize(env=np.zeros(4, dtype=np.float32), env_width=1, env_height=1)
    a = s.arrays()
    tris, nodes = a["triangles"], a["nodes"]
    assert len(tris) == 458 and len(a["materials"]) == 7 and len(a["lights"]) == 1
    assert a["scene_info"]["analytic_light_count"][0] == 1 and a["scene_info"]["emissive_count"][0] == 2 == len(a["emissive"])
    # directional light is normalised (scene.cpp:347-351)
    assert abs(np.linalg.norm(a["lights"]["origin"][0, :3]) - 1) < 1e-6
    # every triangle is covered by exactly one leaf, child boxes are inside parents, leaves hold <= 4 prims
    count = nodes["num_primitives_axis"] >> 16
    leaves = count > 0
    cover = np.zeros(len(tris), dtype=int)
    for off, c in zip(nodes["offset"][leaves], count[leaves]):
        cover[off:off + c] += 1
    assert (cover == 1).all() and count.max() <= 4
    for i in np.nonzero(~leaves)[0]:
        for child in (i + 1, nodes["offset"][i]):
            assert (nodes["bounds_min"][child, :3] >= nodes["bounds_min"][i, :3]).all()
            assert (nodes["bounds_max"][child, :3] <= nodes["bounds_max"][i, :3]).all()
    # materials: the Tf 0 sheet is pass-through, the lamp is emissive, roughness/metalness are 8-bit packed
    mats = a["materials"]
    assert ((mats["ior_emission_idx_transparency"] >> 16) & 0xFF).min() == 0
    assert (mats["emission"] != 0).sum() == 1
    s.close()


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference assets only exist in the build container")
@pytest.mark.parametrize("name", ["CornellBox", "ShaderBalls", "CornellBox_Dragon"])
def test_loader_and_builder_reproduce_reference_dumps(name):
    """Same OBJ/MTL/HDR in -> the arrays the reference's own Scene + Bvh::BuildCPU + LoadHDR produced (fixtures)."""
    s = hostapi.HostScene(os.path.join(REF_ASSETS, name + ".obj"))
    s.add_directional_light((-0.6, -1.5, 3.5), (15, 10, 5))      # main.cpp:58
    s.build_bvh()
    s.finalize(env_path=os.path.join(REF_ASSETS, "ibl", "CGSkies_0036_free.hdr"))
    a, g = s.arrays(), scene(name)
    assert len(a["triangles"]) == len(g["triangles"]) and len(a["nodes"]) == len(g["nodes"])
    for v in ("v1", "v2", "v3"):
        for f in ("position", "texcoord", "normal"):
            assert np.array_equal(bits(a["triangles"][v][f][:, :3]), bits(g["triangles"][v][f][:, :3])), (v, f)
    assert np.array_equal(a["triangles"]["mtlIndex"], g["triangles"]["mtlIndex"])
    for k in ("bounds_min", "bounds_max"):
        assert np.array_equal(bits(a["nodes"][k][:, :3]), bits(g["nodes"][k][:, :3]))
    for k in ("offset", "num_primitives_axis"):
        assert np.array_equal(a["nodes"][k], g["nodes"][k])
    assert a["materials"].tobytes() == g["materials"].tobytes()
    assert np.array_equal(bits(a["lights"]["origin"][:, :3]), bits(g["lights"]["origin"][:, :3]))
    assert np.array_equal(a["emissive"], g["emissive"]) and a["scene_info"].tobytes() == g["scene_info"].tobytes()
    assert np.array_equal(bits(a["env"]), bits(g["env"]))
    s.close()


def test_standalone_bvh_build_on_fixture_triangles():
    g = scene("ShaderBalls")
    rng = np.random.default_rng(0)
    shuffled = g["triangles"][rng.permutation(len(g["triangles"]))][:5000]
    tris, nodes, depth = hostapi.build_bvh(shuffled)
    assert len(tris) == 5000 and depth <= 64
    assert sorted(map(bytes, tris)) == sorted(map(bytes, shuffled))       # a permutation of the input
    count = nodes["num_primitives_axis"] >> 16
    assert count[count > 0].sum() == 5000


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", ["frame", "fused", "stepwise"])
def test_render_kcuda_over_several_devices_matches_oracle(tmp_path, schedule):
    """Render(kCUDA, devices): ONE CUDAPathTraceIntegrator over several GPUs behind the unchanged Integrator interface (the
    devices of the box, or device 0 listed twice when it has one) — Integrate() x3 in every schedule, resolved image against
    the oracle.  "frame" is the deferred schedule: the virtuals only check their order and AdvanceSampleCount submits the frame."""
    from raytracing_b200 import capi
    w, h, mb = 180, 101, 5
    env_path, _ = make_env(tmp_path, w=64, h=32)
    s = hostapi.HostScene(PROC_OBJ)
    s.add_directional_light((-0.6, -1.5, 3.5), (15, 10, 5))
    n_dev = capi.device_count()
    r = hostapi.HostRender(s, w, h, env_path, devices=list(range(n_dev)) if n_dev >= 2 else [0, 0], schedule=schedule)
    r.set_max_bounces(mb)
    a = s.arrays(); a["nodes"] = r.nodes()
    o = Oracle(a)
    cam = hostapi.default_camera(w, h)
    acc = np.zeros((h, w, 4), dtype="<f4")
    for sample in range(3):
        r.render_frame()
        acc, _, _ = o.render(cam, w, h, mb, sample_idx=sample, radiance=acc)
        hdr = acc[..., :3] / np.float32(sample + 1)
        assert np.array_equal(bits(r.image()[..., :3]), bits(hdr / (hdr + np.float32(1.0)))), (schedule, sample)
    r.close(); s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("stepwise", [False, True], ids=["frame", "stepwise"])
def test_render_kcuda_backend_matches_oracle(tmp_path, stepwise):
    """The whole C++ host path — Scene(OBJ) -> Render(kCUDA) -> Bvh::BuildCPU -> Finalize -> CUDAPathTraceIntegrator
    -> UploadGPUData -> RenderFrame() x2 (Integrator::Integrate schedule) — against the oracle on the same arrays."""
    w, h, mb = 200, 120, 6
    env_path, _ = make_env(tmp_path, w=64, h=32)
    s = hostapi.HostScene(PROC_OBJ)
    s.add_directional_light((-0.6, -1.5, 3.5), (15, 10, 5))
    s.add_point_light((0.5, 0.2, 2.0), (4, 4, 6))
    r = hostapi.HostRender(s, w, h, env_path, stepwise=stepwise)
    r.set_max_bounces(mb)
    a = s.arrays()                       # after Render built the BVH (reordering the triangles) and finalized the scene
    a["nodes"] = r.nodes()               # Render owns the acceleration structure
    o = Oracle(a)
    cam = hostapi.default_camera(w, h)
    acc = np.zeros((h, w, 4), dtype="<f4")
    for sample in range(2):
        r.render_frame()
        acc, _, _ = o.render(cam, w, h, mb, sample_idx=sample, radiance=acc)
        hdr = acc[..., :3] / np.float32(sample + 1)              # resolve_radiance.cl:80-84
        expect = hdr / (hdr + np.float32(1.0))
        img = r.image()
        assert np.array_equal(bits(img[..., :3]), bits(expect)), sample
        assert (img[..., 3] == 1.0).all()
    # option setters keep the reference behaviour: kBlueNoise without its tables fails loudly; with them the next frame
    # restarts the accumulation with the other sampler; white furnace re-renders from scratch
    with pytest.raises(hostapi.HostError):
        r.set_blue_noise(True)
    tables = synthetic_sampler_tables()
    r.set_blue_noise_tables(*tables)
    r.set_blue_noise(True)
    r.render_frame()
    try:
        o.set_sampler_tables(tables)
        bn, _, _ = o.render(cam, w, h, mb, sample_idx=0)
    finally:
        o.set_sampler_tables(None)
    assert np.array_equal(bits(r.image()[..., :3]), bits(bn[..., :3] / (bn[..., :3] + np.float32(1.0))))
    r.set_blue_noise(False)
    r.enable_white_furnace(True)
    r.render_frame()
    wf, _, _ = o.render(cam, w, h, mb, sample_idx=0, white_furnace=True)
    expect = wf[..., :3] / (wf[..., :3] + np.float32(1.0))
    assert np.array_equal(bits(r.image()[..., :3]), bits(expect))
    r.close(); s.close()
