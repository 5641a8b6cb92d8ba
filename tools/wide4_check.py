"""Pins raytracing_b200/csrc/rt_wide4.h (the traversal the RT_OPT_TRAVERSAL = 3 kernels compile, run here on the CPU through
tools/wide4_check.cpp over the layouts rt_upload_scene builds) against the oracle's literal binary traversal: primitive id,
t, barycentrics (closest hit) and occlusion (any hit), bit for bit, on the rays entering bounces 0..3 of a frame."""
import ctypes as C, os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle.orcbind import Oracle
from raytracing_b200 import scene_io
from raytracing_b200.camera import default_camera
from raytracing_b200.layouts import RAY_DT, HIT_DT

lib_path = os.path.join(tempfile.gettempdir(), "libwide4_check.so")
subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-I" + os.path.join(REPO, "include"),
                "-shared", "-o", lib_path, os.path.join(HERE, "wide4_check.cpp")], check=True)
W = C.CDLL(lib_path)
W.wide4_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

w, h, mb = 480, 270, 4
total_bad = total = 0
for name in ("CornellBox", "ShaderBalls", "CornellBox_Dragon"):
    sc = scene_io.load_scene(name); o = Oracle(sc)
    cams = [default_camera(w, h), default_camera(w, h, position=(0.4, -1.5, 0.7), yaw=1.8, pitch=1.45, aperture=0.03, focus_distance=2.0)]
    nodes = np.ascontiguousarray(sc["nodes"]); tris = np.ascontiguousarray(sc["triangles"])
    for ci, cam in enumerate(cams):
        for bounce in range(4):
            rays = np.zeros(w * h, dtype=RAY_DT); work = np.zeros((w * h, 2), np.uint32); cnt = C.c_uint32(0)
            o.lib.orc_set_dump(bounce, rays.ctypes.data_as(C.c_void_p), work.ctypes.data_as(C.c_void_p), C.byref(cnt))
            o.render(cam, w, h, mb, want_hits=False)
            o.lib.orc_set_dump(-1, None, None, None)
            n = cnt.value
            if n == 0:
                continue
            r = np.ascontiguousarray(rays[:n])
            for any_hit in (0, 1):
                bh = np.zeros(n, dtype=HIT_DT); bf = np.zeros(n, dtype=np.uint32); bc = np.zeros(2, dtype=np.uint64)
                o.lib.orc_trace(C.byref(o.scene), r.ctypes.data, n, any_hit, bh.ctypes.data, bf.ctypes.data, bc.ctypes.data)
                wh = np.zeros(n, dtype=HIT_DT); st = np.zeros(n, dtype=np.uint8); nw = C.c_uint64(0)
                rc = W.wide4_trace(nodes.ctypes.data, len(nodes), tris.ctypes.data, len(tris), r.ctypes.data, n, any_hit, wh.ctypes.data, st.ctypes.data, C.byref(nw))
                assert rc == 0
                ok = st == 0
                if any_hit:
                    same = bf[ok] == np.where(wh["primitive_id"][ok] == 0, 0, 0xFFFFFFFF).astype(np.uint32)
                else:
                    hit = bh["primitive_id"][ok] != 0xFFFFFFFF
                    same = bh["primitive_id"][ok] == wh["primitive_id"][ok]
                    same &= ~hit | ((bh["t"][ok].view(np.uint32) == wh["t"][ok].view(np.uint32)) &
                                    (bh["bc"][ok].view(np.uint32).reshape(-1, 2) == wh["bc"][ok].view(np.uint32).reshape(-1, 2)).all(1))
                bad = int((~same).sum()); total_bad += bad; total += int(ok.sum())
                print(f"{name:18s} cam {ci} bounce {bounce} {'any    ' if any_hit else 'closest'} rays {n:7d} literal-only {int((~ok).sum()):4d} "
                      f"wide nodes {nw.value:6d} (binary nodes {len(nodes)}) mismatches {bad}")
print("RAYS", total, "TOTAL MISMATCHES", total_bad)
sys.exit(1 if total_bad else 0)
