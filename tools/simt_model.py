"""Issue-slot model of the traversal kernels' warp schedules (analysis only, CPU; see tools/simt_model.cpp).
usage: python tools/simt_model.py [scene] [width height max_bounces]"""
import ctypes as C, os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle.orcbind import Oracle
from raytracing_b200 import scene_io
from raytracing_b200.camera import default_camera
from raytracing_b200.layouts import RAY_DT

lib_path = os.path.join(tempfile.gettempdir(), "libsimt_model.so")
subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-I" + os.path.join(REPO, "include"),
                "-shared", "-o", lib_path, os.path.join(HERE, "simt_model.cpp")], check=True)
M = C.CDLL(lib_path)
M.simt_model.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]

name = sys.argv[1] if len(sys.argv) > 1 else "CornellBox"
w, h, mb = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (960, 540, 3)
sc = scene_io.load_scene(name); o = Oracle(sc); cam = default_camera(w, h)
nodes = np.ascontiguousarray(sc["nodes"]); tris = np.ascontiguousarray(sc["triangles"])
# issue slots: interior record step, triangle test by exit stage (det / u / v / full), ray fetch + setup + result write
costs = np.array([75.0, 24.0, 48.0, 66.0, 78.0, 90.0])

def run(rays, any_hit, schedule, K, per):
    out = np.zeros(4)
    rc = M.simt_model(nodes.ctypes.data, len(nodes), tris.ctypes.data, len(tris), rays.ctypes.data, len(rays), any_hit, schedule, K, per,
                      costs.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out

for bounce in range(0, mb + 1):
    rays = np.zeros(w * h, dtype=RAY_DT); work = np.zeros((w * h, 2), np.uint32); cnt = C.c_uint32(0)
    o.lib.orc_set_dump(bounce, rays.ctypes.data_as(C.c_void_p), work.ctypes.data_as(C.c_void_p), C.byref(cnt))
    o.render(cam, w, h, mb, want_hits=False)
    o.lib.orc_set_dump(-1, None, None, None)
    n = cnt.value
    if n == 0:
        continue
    r = np.ascontiguousarray(rays[:n])
    base = run(r, 0, 0, 1, 32)
    print(f"{name} bounce {bounce}: {n} rays, {base[2] / n:.1f} record steps + {base[3] / n:.1f} triangle tests per ray")
    print(f"   {'schedule':34s} {'slots/ray':>10s} {'lanes':>6s} {'speed-up':>8s}")
    rows = [("1 ray / lane, batch of 32 (round 1)", 0, 1, 32)]
    for K in (2, 4, 8, 16):
        rows.append((f"{K} rays / lane, static", 1, K, 32 * K))
    rows.append(("ideal dynamic refill", 2, 1, 32 * 256))
    for label, sch, K, per in rows:
        out = run(r, 0, sch, K, per)
        print(f"   {label:34s} {out[0] / n:10.1f} {out[1] / out[0]:6.1f} {base[0] / out[0]:8.2f}x")

# ---- effect of re-ordering the bounce's ray queue on the round-1 schedule (phase coherence, not only length variance)
def octant(d): return (d[:, 0] < 0).astype(np.int64) | ((d[:, 1] < 0).astype(np.int64) << 1) | ((d[:, 2] < 0).astype(np.int64) << 2)
def morton_cells(org, bits):
    mn, mx = org.min(0), org.max(0)
    cell = np.clip(((org - mn) / np.maximum(mx - mn, 1e-9) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    m = np.zeros(len(org), np.int64)
    for b in range(bits):
        for a in range(3):
            m |= ((cell[:, a] >> b) & 1) << (3 * b + a)
    return m
def dir_cells(d, bits):
    # octahedral-ish: quantised direction
    q = np.clip(((d * 0.5 + 0.5) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    m = np.zeros(len(d), np.int64)
    for b in range(bits):
        for a in range(3):
            m |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return m

print("\nre-ordering (schedule: 1 ray / lane, batch of 32)")
for bounce in (1, 2):
    rays = np.zeros(w * h, dtype=RAY_DT); work = np.zeros((w * h, 2), np.uint32); cnt = C.c_uint32(0)
    o.lib.orc_set_dump(bounce, rays.ctypes.data_as(C.c_void_p), work.ctypes.data_as(C.c_void_p), C.byref(cnt))
    o.render(cam, w, h, mb, want_hits=False)
    o.lib.orc_set_dump(-1, None, None, None)
    n = cnt.value
    r = np.ascontiguousarray(rays[:n])
    org = r["origin"][:, :3].astype(np.float64); d = r["direction"][:, :3].astype(np.float64)
    base = run(r, 0, 0, 1, 32)
    oc = octant(d)
    orders = {"queue order": np.arange(n),
              "octant (8 sub-queues, stable)": np.argsort(oc, kind="stable"),
              "octant + origin cell 4b": np.argsort(oc * (1 << 12) + morton_cells(org, 4), kind="stable"),
              "origin cell 4b + octant": np.argsort(morton_cells(org, 4) * 8 + oc, kind="stable"),
              "origin cell 3b + dir cell 3b": np.argsort(morton_cells(org, 3) * 512 + dir_cells(d, 3), kind="stable"),
              "dir cell 3b + origin cell 3b": np.argsort(dir_cells(d, 3) * 512 + morton_cells(org, 3), kind="stable"),
              "dir cell 4b + origin cell 4b": np.argsort(dir_cells(d, 4) * 4096 + morton_cells(org, 4), kind="stable"),
              "random": np.random.default_rng(1).permutation(n)}
    print(f"{name} bounce {bounce}")
    for label, order in orders.items():
        rr = np.ascontiguousarray(r[order])
        out = run(rr, 0, 0, 1, 32)
        print(f"   {label:34s} {out[0] / n:10.1f} {out[1] / out[0]:6.1f} {base[0] / out[0]:8.2f}x")
