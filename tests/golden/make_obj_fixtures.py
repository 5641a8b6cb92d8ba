#!/usr/bin/env python3
"""
tests/golden/make_obj_fixtures.py — what the REFERENCE's own scene loader (tinyobjloader + Scene::Load + Bvh::BuildCPU,
compiled into oracle/_ref) makes of the OBJ / MTL inputs of tests/obj_cases.py.

    python tests/golden/make_obj_fixtures.py          (where /root/reference and oracle/_ref/libref.so exist)

Writes tests/golden/obj/expected.npz: per case the triangle array (leaf order of the reference's BVH), the packed materials
and the texture table; and bvh_soups.json: digests of the trees Bvh::BuildCPU built over ten generated triangle soups
(uniform, clustered, repeated centroids, regular grid, one axis; 3 000 and 20 000 triangles); and hdr_files.json: digests of
the environment images LoadHDR read from the files of tests/hdr_cases.py.  tests/test_host.py::test_obj_reader_cases_load_like_the_reference loads the same inputs (re-created
from tests/obj_cases.py, nothing but the expectations is stored) through host/obj_reader.cpp and compares.
"""
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from tests import obj_cases  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "obj")


# cases loaded with the reference's --scale / --flip_yz options (main.cpp:45-47; the Bistro command line uses 0.01 and 1)
LOAD_OPTIONS = {"scaled_flipped_quads_polys": (0.01, True), "scaled_groups": (3.5, False), "flipped_plain": (1.0, True)}


def all_cases():
    for name, files in obj_cases.QUIRK_CASES.items():
        yield name, files
    for seed, features in obj_cases.RANDOM_CASES:
        yield "random_%d_%s" % (seed, "_".join(features) or "plain"), obj_cases.random_scene(seed, features)
    yield "scaled_flipped_quads_polys", obj_cases.random_scene(50, ("quads", "polys", "neg"))
    yield "scaled_groups", obj_cases.random_scene(51, ("groups", "tabs", "quads"))
    yield "flipped_plain", obj_cases.random_scene(52, ())


def reference_env(tmp, hdr_bytes):
    """(env, width, height) as the reference reads `hdr_bytes`: Scene::Finalize opens assets/ibl/CGSkies_0036_free.hdr relative to the
    working directory (scene.cpp:360), so the file is placed under that name in a scratch root."""
    from oracle.refbind import RefRenderer
    os.makedirs(os.path.join(tmp, "assets", "ibl"))
    with open(os.path.join(tmp, "assets", "ibl", "CGSkies_0036_free.hdr"), "wb") as f:
        f.write(hdr_bytes)
    obj = obj_cases.write_case(tmp, obj_cases.QUIRK_CASES["plain"])
    sc = RefRenderer().open_obj(tmp, obj).scene()
    return sc["env"], int(sc["env_width"]), int(sc["env_height"])


def main():
    from oracle.refbind import RefRenderer
    os.makedirs(OUT, exist_ok=True)
    arrays = {}
    for name, files in all_cases():
        with tempfile.TemporaryDirectory() as tmp:
            obj = obj_cases.write_case(tmp, files)
            scale, flip = LOAD_OPTIONS.get(name, (1.0, False))
            sc = RefRenderer().open_obj("/root/reference", obj, scale=scale, flip_yz=flip).scene()
        arrays[name + ":triangles"] = sc["triangles"]
        arrays[name + ":materials"] = sc["materials"]
        arrays[name + ":textures"] = sc["textures"]
        print(f"{name:60s} {len(sc['triangles']):4d} triangles {len(sc['materials'])} materials {len(sc['textures'])} textures")
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **arrays)
    # BVH builder inputs: only a digest of the reference's tree is stored
    import json
    digests = {}
    for mode, seed, n in obj_cases.SOUP_CASES:
        with tempfile.TemporaryDirectory() as tmp:
            obj = obj_cases.write_case(tmp, obj_cases.triangle_soup(mode, seed, n))
            digests["%s_%d_%d" % (mode, seed, n)] = obj_cases.tree_digest(RefRenderer().open_obj("/root/reference", obj).scene())
        print(mode, seed, n, digests["%s_%d_%d" % (mode, seed, n)][:24])
    with open(os.path.join(OUT, "bvh_soups.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)
    # environment maps: digests of what the reference's LoadHDR (through Scene::Finalize) read
    from tests import hdr_cases
    digests = {}
    for name, data in hdr_cases.cases().items():
        with tempfile.TemporaryDirectory() as tmp:
            digests[name] = hdr_cases.digest(*reference_env(tmp, data))
        print(name, digests[name][:28])
    with open(os.path.join(OUT, "hdr_files.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
