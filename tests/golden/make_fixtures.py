#!/usr/bin/env python3
"""
tests/golden/make_fixtures.py — regenerates every committed fixture from the
reference itself.  Run where /root/reference exists (after python oracle/build_ref.py):

    python tests/golden/make_fixtures.py

Scenes  (tests/golden/scenes/*.npz.xz): produced by the reference's OWN Scene loader,
    Bvh::BuildCPU and LoadHDR (compiled in place into oracle/_ref/libref.so) from
    assets/{CornellBox,ShaderBalls,CornellBox_Dragon}.obj + the directional light
    main.cpp:58 adds; env map assets/ibl/CGSkies_0036_free.hdr.
Golden outputs (tests/golden/golden_*.npz.xz): produced by the reference's OWN kernels
    (src/kernels/cl/*.cl compiled for the CPU, math built-ins = include/rt_math.h) driven
    by the reference's Integrator::Integrate(): primary rays/hits, radiance, per-bounce
    ray counters, AOVs, resolved image.  sample_idx = 0, default camera, kRandom sampler.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from oracle.refbind import RefRenderer  # noqa: E402
from raytracing_b200.camera import default_camera  # noqa: E402
from raytracing_b200 import scene_io  # noqa: E402

REFERENCE = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")

# (scene, width, height, max_bounces, white_furnace, full dump of rays/hits/AOVs?[, extras])
# extras: "samples" = progressive samples accumulated (sample_idx 0..n-1), "camera" = default_camera keyword arguments,
# "tag" = file-name suffix.  Primary hits and counters of a multi-sample fixture are those of its LAST sample.
GOLDEN = [
    ("CornellBox", 128, 128, 2, False, True),
    ("CornellBox", 256, 256, 2, False, False),       # BASELINE config C1
    ("CornellBox", 128, 128, 4, True, False),        # white furnace
    ("ShaderBalls", 320, 180, 8, False, False),      # C3 at reduced resolution
    ("CornellBox_Dragon", 240, 135, 16, False, False),  # C4 at reduced resolution
    ("CornellBox", 160, 96, 3, False, False, {"samples": 3, "tag": "lens3spp",
                                              "camera": {"position": (0.15, -1.3, 0.9), "aperture": 0.04, "focus_distance": 1.7}}),
    ("ShaderBalls", 200, 112, 5, False, False, {"samples": 2, "tag": "moved2spp", "camera": {"position": (0.6, -1.6, 0.8), "yaw": 1.9, "pitch": 1.4}}),
]


def main():
    os.makedirs(scene_io.SCENE_DIR, exist_ok=True)
    renderers = {}
    env_saved = False
    for name in ("CornellBox", "ShaderBalls", "CornellBox_Dragon"):
        r = RefRenderer().open_obj(REFERENCE, f"assets/{name}.obj")
        sc = r.scene()
        keep = "--only-new" in sys.argv and os.path.exists(os.path.join(scene_io.SCENE_DIR, name + ".npz.xz"))
        if not keep:       # (struct padding bytes are uninitialised in the reference: a re-saved scene differs in those bytes only)
            scene_io.save_scene(os.path.join(scene_io.SCENE_DIR, name + ".npz.xz"), sc)
        if not env_saved and not keep:
            scene_io.save_env(os.path.join(scene_io.SCENE_DIR, scene_io.ENV_NAME + ".npz.xz"), sc["env"], sc["env_width"], sc["env_height"])
            env_saved = True
        renderers[name] = r
        print(name, len(sc["triangles"]), "triangles", len(sc["nodes"]), "nodes")

    only_new = "--only-new" in sys.argv
    for entry in GOLDEN:
        (name, w, h, mb, wf, full), extra = entry[:6], (entry[6] if len(entry) > 6 else {})
        fn = f"golden_{name}_{w}x{h}_b{mb}{'_wf' if wf else ''}{'_' + extra['tag'] if extra.get('tag') else ''}.npz.xz"
        if only_new and os.path.exists(os.path.join(OUT, fn)):
            continue
        r = renderers[name]
        r.begin(w, h)
        cam = default_camera(w, h, **extra.get("camera", {}))
        r.set_camera(cam)
        r.set_max_bounces(mb)
        r.enable_white_furnace(wf)
        for _ in range(extra.get("samples", 1)):
            r.integrate()
        hits = r.primary_hits()
        st = r.stats()
        out = {
            "camera": cam, "width": np.uint32(w), "height": np.uint32(h), "max_bounces": np.uint32(mb),
            "white_furnace": np.uint32(wf), "sample_count": np.uint32(r.sample_count()),
            "primitive_id": hits["primitive_id"].copy(),
            "radiance_rgb": r.radiance()[..., :3].copy(),
        }
        for k, v in st.items():
            out[k] = v[: mb + 1].copy()
        if full:
            rays = r.primary_rays()
            out["ray_origin"] = rays["origin"].copy()
            out["ray_direction"] = rays["direction"].copy()
            out["hit_bc"] = hits["bc"].copy()
            out["hit_t"] = hits["t"].copy()
            out["resolved"] = r.resolved().copy()
            out["aov_albedo"] = r.aov_albedo()[..., :3].copy()
            out["aov_depth"] = r.aov_depth().copy()
            out["aov_normal"] = r.aov_normal()[..., :3].copy()
            out["aov_velocity"] = r.aov_velocity().copy()
        scene_io.save_npz_xz(os.path.join(OUT, fn), out)
        print(fn, os.path.getsize(os.path.join(OUT, fn)), "bytes")


if __name__ == "__main__":
    main()
