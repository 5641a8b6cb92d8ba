"""Cost model for ray re-ordering between bounces (analysis only; uses the oracle's per-ray traversal work).
A warp costs max over its 32 rays of (nodes visited + c * triangles tested); SIMD efficiency = sum(work) / (32 * sum(warp max)).
Compares queue order with a few orderings that a GPU could produce cheaply."""
import ctypes as C, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.orcbind import Oracle
from raytracing_b200 import scene_io
from raytracing_b200.camera import default_camera
from raytracing_b200.layouts import RAY_DT

def eff(work, order=None):
    w = work if order is None else work[order]
    n = (len(w) + 31) // 32 * 32
    p = np.zeros(n); p[: len(w)] = w
    m = p.reshape(-1, 32).max(axis=1)
    return w.sum() / (32.0 * m.sum()), m.sum()

def octant(d): return (d[:, 0] < 0).astype(np.int64) | ((d[:, 1] < 0).astype(np.int64) << 1) | ((d[:, 2] < 0).astype(np.int64) << 2)

def blocked_sort(key, block):
    n = len(key); order = np.arange(n)
    for s in range(0, n, block):
        e = min(n, s + block)
        order[s:e] = s + np.argsort(key[s:e], kind="stable")
    return order

name = sys.argv[1] if len(sys.argv) > 1 else "ShaderBalls"
w, h, mb = 960, 540, 4
sc = scene_io.load_scene(name); o = Oracle(sc); cam = default_camera(w, h)
lo = sc["nodes"]["min"][0][:3] if "min" in sc["nodes"].dtype.names else None
for bounce in (1, 2, 3):
    rays = np.zeros(w * h, dtype=RAY_DT); work = np.zeros((w * h, 2), np.uint32); cnt = C.c_uint32(0)
    o.lib.orc_set_dump(bounce, rays.ctypes.data_as(C.c_void_p), work.ctypes.data_as(C.c_void_p), C.byref(cnt))
    o.render(cam, w, h, mb, want_hits=False)
    o.lib.orc_set_dump(-1, None, None, None)
    n = cnt.value
    r = rays[:n]; cost = work[:n, 0].astype(np.float64) + 1.0 * work[:n, 1]
    org = np.stack([r["origin"][k] for k in ("x", "y", "z")], 1) if r["origin"].dtype.names else r["origin"][:, :3]
    d = np.stack([r["direction"][k] for k in ("x", "y", "z")], 1) if r["direction"].dtype.names else r["direction"][:, :3]
    base, m0 = eff(cost)
    oc = octant(d)
    mn, mx = org.min(0), org.max(0)
    cell = np.clip(((org - mn) / np.maximum(mx - mn, 1e-9) * 16).astype(np.int64), 0, 15)
    morton = np.zeros(n, np.int64)
    for b in range(4):
        for a in range(3):
            morton |= ((cell[:, a] >> b) & 1) << (3 * b + a)
    res = {"queue order": base}
    for blk in (256, 1024, 8192, n):
        res[f"octant, block {blk}"] = eff(cost, blocked_sort(oc, blk))[0]
        res[f"octant+cell, block {blk}"] = eff(cost, blocked_sort(oc * 4096 + morton, blk))[0]
        res[f"cell+octant, block {blk}"] = eff(cost, blocked_sort(morton * 8 + oc, blk))[0]
    res["by cost (oracle knowledge, upper bound)"] = eff(cost, np.argsort(cost))[0]
    print(f"{name} bounce {bounce}: {n} rays, mean work {cost.mean():.1f}")
    for k, v in res.items():
        print(f"   {k:42s} eff {v:.3f}  speed-up vs queue order {v / base:.2f}x")
