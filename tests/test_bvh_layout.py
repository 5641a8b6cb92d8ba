"""
The traversal record layout that rt_upload_scene derives from the reference's LinearBVHNode[] / Triangle[]
(raytracing_b200/csrc/rt_bvh_layout.h), checked on the CPU against the reference arrays it was built from: every interior node
becomes one record holding exactly its two children's boxes, child references form the same tree, leaves point at the same
triangle runs, edges are the same float subtractions, and the first `top_n` records are a breadth-first top of the tree (any
prefix is closed under "parent of") — what the TMA top-of-tree staging relies on.  Malformed trees are refused.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from raytracing_b200.layouts import NODE_DT, TRIANGLE_DT
from tests.helpers import scene

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(tempfile.gettempdir(), f"libbvh_layout_host_{os.getpid()}.so")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I" + os.path.join(REPO, "include"),
                    "-I" + os.path.join(REPO, "raytracing_b200", "csrc"), "-o", so, os.path.join(HERE, "bvh_layout_host.cpp")], check=True)
    L = C.CDLL(so)
    L.bvh_layout_build.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
    L.shade_records_build.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    yield L
    os.remove(so)


def build(L, nodes, tris):
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DT); tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    wn = np.zeros((max(len(nodes), 1), 16), "<f4"); wt = np.zeros((len(tris), 12), "<f4")
    root, depth, top, nrec = C.c_int(), C.c_int(), C.c_uint32(), C.c_uint64()
    err = C.create_string_buffer(256)
    rc = L.bvh_layout_build(nodes.ctypes.data, len(nodes), tris.ctypes.data, len(tris), wn.ctypes.data, wt.ctypes.data,
                            C.byref(root), C.byref(depth), C.byref(top), C.byref(nrec), err, 256)
    return rc, err.value.decode(), wn[: nrec.value], wt, root.value, depth.value, top.value


@pytest.mark.parametrize("name", ["CornellBox", "ShaderBalls", "CornellBox_Dragon"])
def test_record_layout_matches_the_reference_tree(lib, name):
    sc = scene(name)
    nodes, tris = sc["nodes"], sc["triangles"]
    rc, err, wn, wt, root, depth, top_n = build(lib, nodes, tris)
    assert rc == 0, err
    count = nodes["num_primitives_axis"] >> 16
    interior = np.nonzero(count == 0)[0]
    assert len(wn) == len(interior) and root == 0 and 1 <= top_n <= min(len(interior), 1024)
    refs = wn[:, 12:14].copy().view("<i4")                      # (ref0, ref1) per record
    axis = wn[:, 14].copy().view("<u4")
    # walk both trees together from the root
    rec_of = {0: 0}
    parent_of_rec = {0: -1}
    todo = [0]
    seen_recs = set()
    while todo:
        n = todo.pop()
        r = rec_of[n]
        assert r not in seen_recs; seen_recs.add(r)
        kids = (n + 1, int(nodes["offset"][n]))
        boxes = wn[r, :12]
        want = np.concatenate([nodes["bounds_min"][kids[0], :3], nodes["bounds_max"][kids[0], :3], nodes["bounds_min"][kids[1], :3], nodes["bounds_max"][kids[1], :3]])
        assert np.array_equal(boxes.view("<u4"), want.astype("<f4").view("<u4")), (name, n)
        assert axis[r] == (nodes["num_primitives_axis"][n] & 0xFFFF)
        for k, child in enumerate(kids):
            ref = int(refs[r, k])
            if count[child] > 0:
                assert ref == ~int(nodes["offset"][child])       # leaf: first triangle of the run
                last = int(nodes["offset"][child]) + int(count[child]) - 1
                flags = wt[int(nodes["offset"][child]): last + 1, 9].copy().view("<u4")
                assert flags[-1] == 1 and (flags[:-1] == 0).all()   # end-of-leaf flag on the last triangle only
            else:
                assert 0 <= ref < len(wn)
                rec_of[child] = ref; parent_of_rec[ref] = r
                todo.append(child)
    assert len(seen_recs) == len(wn)
    # breadth-first top: every record of the prefix has its parent earlier in the prefix
    for r in range(1, top_n):
        assert 0 <= parent_of_rec[r] < r
    # triangle records: p1, e1 = p2 - p1, e2 = p3 - p1 as single float subtractions
    p1 = tris["v1"]["position"][:, :3]; p2 = tris["v2"]["position"][:, :3]; p3 = tris["v3"]["position"][:, :3]
    assert np.array_equal(wt[:, 0:3].view("<u4"), p1.view("<u4"))
    assert np.array_equal(wt[:, 3:6].view("<u4"), (p2 - p1).astype("<f4").view("<u4"))
    assert np.array_equal(wt[:, 6:9].view("<u4"), (p3 - p1).astype("<f4").view("<u4"))
    assert depth <= 64


def test_malformed_trees_are_refused(lib):
    sc = scene("CornellBox")
    nodes = sc["nodes"].copy(); tris = sc["triangles"]
    bad = nodes.copy(); bad["offset"][0] = len(nodes) + 5                   # child index out of range
    assert build(lib, bad, tris)[0] == 1
    bad = nodes.copy()
    first_interior = int(np.nonzero((nodes["num_primitives_axis"] >> 16) == 0)[0][0])
    bad["offset"][first_interior] = first_interior + 1                      # both children the same node: not a tree
    rc, err, *_ = build(lib, bad, tris)
    assert rc == 1 and "twice" in err
    bad = nodes.copy()
    leaf = int(np.nonzero((nodes["num_primitives_axis"] >> 16) > 0)[0][0])
    bad["offset"][leaf] = len(tris)                                         # leaf run outside the triangle array
    assert build(lib, bad, tris)[0] == 1
    bad = nodes.copy(); bad["num_primitives_axis"][first_interior] = 3      # split axis 3
    assert build(lib, bad, tris)[0] == 1


@pytest.mark.parametrize("name", ["CornellBox", "ShaderBalls"])
def test_shading_records_are_the_reference_unpacks(lib, name):
    """tri_shade / mat_rec / light_rec (rt_bvh_layout.h): per-upload evaluations of what the reference computes per hit —
    geometric normal = normalize(cross(p2 - p1, p3 - p1)) (hit_surface.cl:91), the packed-material unpack (utils.h:133-190),
    the directional light's direction and distance (light.h:52-60, hit_surface.cl:122-123) — checked against the same float32
    operations in numpy."""
    from raytracing_b200.layouts import LIGHT_DT, MATERIAL_DT
    sc = scene(name)
    if name == "ShaderBalls":
        from tests.scenes_extra import many_lights_scene
        sc = many_lights_scene(name)                      # point + directional lights
    tris = np.ascontiguousarray(sc["triangles"], dtype=TRIANGLE_DT); mats = np.ascontiguousarray(sc["materials"], dtype=MATERIAL_DT)
    lights = np.ascontiguousarray(sc["lights"], dtype=LIGHT_DT)
    to = np.zeros((len(tris), 28), "<f4"); mo = np.zeros((len(mats), 16), "<f4"); lo = np.zeros((len(lights), 8), "<f4")
    lib.shade_records_build(tris.ctypes.data, len(tris), mats.ctypes.data, len(mats), lights.ctypes.data, len(lights), to.ctypes.data, mo.ctypes.data, lo.ctypes.data)
    f = np.float32
    p1, p2, p3 = (tris[v]["position"][:, :3] for v in ("v1", "v2", "v3"))
    e1, e2 = (p2 - p1).astype(f), (p3 - p1).astype(f)
    cx = (e1[:, 1] * e2[:, 2]).astype(f) - (e1[:, 2] * e2[:, 1]).astype(f)
    cy = (e1[:, 2] * e2[:, 0]).astype(f) - (e1[:, 0] * e2[:, 2]).astype(f)
    cz = (e1[:, 0] * e2[:, 1]).astype(f) - (e1[:, 1] * e2[:, 0]).astype(f)
    d2 = ((cx * cx).astype(f) + (cy * cy).astype(f)).astype(f) + (cz * cz).astype(f)
    with np.errstate(all="ignore"):
        inv = (f(1.0) / np.sqrt(d2.astype(f))).astype(f)
    gn = np.stack([(cx * inv).astype(f), (cy * inv).astype(f), (cz * inv).astype(f)], 1)
    bits = lambda a: np.ascontiguousarray(a, dtype="<f4").view("<u4")
    same = lambda a, b: np.array_equal(bits(a), bits(b)) or np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
    assert np.array_equal(bits(to[:, 0:3]), bits(p1)) and np.array_equal(bits(to[:, 4:7]), bits(p2)) and np.array_equal(bits(to[:, 8:11]), bits(p3))
    assert np.array_equal(to[:, 3].copy().view("<u4"), tris["mtlIndex"])
    assert same(np.stack([to[:, 7], to[:, 11], to[:, 15]], 1), gn)
    assert np.array_equal(bits(to[:, 12:15]), bits(tris["v1"]["normal"][:, :3])) and np.array_equal(bits(to[:, 16:19]), bits(tris["v2"]["normal"][:, :3]))
    assert np.array_equal(bits(to[:, 20:23]), bits(tris["v3"]["normal"][:, :3]))
    uv = np.stack([to[:, 19], to[:, 23], to[:, 24], to[:, 25], to[:, 26], to[:, 27]], 1)
    want_uv = np.concatenate([tris["v1"]["texcoord"][:, :2], tris["v2"]["texcoord"][:, :2], tris["v3"]["texcoord"][:, :2]], 1)
    assert np.array_equal(bits(uv), bits(want_uv))
    # materials
    def rgb(w):
        return np.stack([((w >> s) & 0xFF).astype(f) / f(255.0) for s in (0, 8, 16)], 1).astype(f)
    assert np.array_equal(bits(mo[:, 0:3]), bits(rgb(mats["diffuse_albedo"]))) and np.array_equal(bits(mo[:, 4:7]), bits(rgb(mats["specular_albedo"])))
    rm, it, em = mats["roughness_metalness"], mats["ior_emission_idx_transparency"], mats["emission"]
    assert np.array_equal(bits(mo[:, 3]), bits((rm & 0xFF).astype(f) / f(255.0))) and np.array_equal(bits(mo[:, 7]), bits(((rm >> 16) & 0xFF).astype(f) / f(255.0)))
    scale = np.ldexp(f(1.0), (em >> 24).astype(np.int32) - 136).astype(f)
    want_em = np.stack([((em >> s) & 0xFF).astype(f) * scale for s in (0, 8, 16)], 1).astype(f)
    assert np.array_equal(bits(mo[:, 8:11]), bits(want_em))
    assert np.array_equal(bits(mo[:, 11]), bits((it & 0xFF).astype(f) / f(25.5))) and np.array_equal(bits(mo[:, 12]), bits(((it >> 16) & 0xFF).astype(f) / f(255.0)))
    textured = ((mats["diffuse_albedo"] >> 24) != 0xFF) | ((mats["specular_albedo"] >> 24) != 0xFF) | (((rm >> 8) & 0xFF) != 0xFF) | ((rm >> 24) != 0xFF) | \
               (((it >> 8) & 0xFF) != 0xFF) | ((it >> 24) != 0xFF)
    assert np.array_equal(mo[:, 13].copy().view("<u4"), textured.astype("<u4"))
    # lights
    for i, l in enumerate(lights):
        assert np.array_equal(bits(lo[i, 4:7]), bits(l["radiance"][:3])) and lo[i, 7:8].copy().view("<u4")[0] == l["type"]
        if l["type"] == 0:                                   # point light: origin as is
            assert np.array_equal(bits(lo[i, 0:3]), bits(l["origin"][:3]))
        else:                                                # directional: normalize(origin * MAX_RENDER_DIST) and its length
            v = (l["origin"][:3].astype(f) * f(20000.0)).astype(f)   # RT_MAX_RENDER_DIST
            dd = f(f(f(v[0] * v[0]) + f(v[1] * v[1])) + f(v[2] * v[2]))
            assert np.array_equal(bits(lo[i, 0:3]), bits((v * (f(1.0) / np.sqrt(dd))).astype(f))) and bits(lo[i, 3:4])[0] == bits(np.sqrt(dd))[0]
