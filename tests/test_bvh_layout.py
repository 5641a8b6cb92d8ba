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
