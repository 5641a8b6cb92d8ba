"""Same-box A/B of library variants (tools/build_variant.py): ms/frame of a 1/world partition for every library given.
usage: ab_probe.py scene worlds lib[:fk[:threads[:top]]] ...   (lib = path, variant name or "cur" for the in-tree library; fk = RT_OPT_FRAME_KERNEL
value, threads = RT_OPT_FRAME_THREADS value, top = RT_OPT_TOP_SMEM records; default: the library's defaults)"""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__)); REPO = os.path.dirname(HERE)
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from raytracing_b200 import capi, scene_io
from raytracing_b200.camera import default_camera
name, worlds, fk, thr, top = sys.argv[1], [int(x) for x in sys.argv[2].split(",")], sys.argv[3], sys.argv[4], sys.argv[5]
w, h, mb = (3840, 2160, 16) if name == "CornellBox_Dragon" else (1920, 1080, 8)
if name == "Synthetic10M":
    import bench
    sc, _ = bench.load_workload_scene(name, w, h)          # cached per box
else:
    sc = scene_io.load_scene(name)
out = []
for world in worlds:
    ctx = capi.Context(w, h, device=0, rank=0, world=world)
    if fk != "-":
        ctx.set_option(26, int(fk))
    if thr != "-":
        ctx.set_option(28, int(thr))
    if top != "-":
        ctx.set_option(29, int(top))
    ctx.upload_scene(sc); ctx.set_camera(sc["camera_pose"] if "camera_pose" in sc else default_camera(w, h))
    stream = torch.cuda.ExternalStream(ctx.stream_handle())
    best = 1e9
    for rep in range(3):
        for _ in range(5):
            ctx.reset(); ctx.integrate(mb)
        ctx.sync()
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            ctx.reset(); ctx.integrate(mb)
        e1.record(stream); ctx.sync(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    out.append("%%.3f" %% best)
    ctx.destroy()
print(" ".join(out))
''' % REPO
name, worlds = sys.argv[1], sys.argv[2]
print(f"{name}: ms/frame (best of 3 x 20 frames) for world = {worlds}")
for spec in sys.argv[3:]:
    lib, fk, thr, top = (spec.split(":") + ["", "", ""])[:4]
    path = lib if os.path.sep in lib else (os.path.join(REPO, "raytracing_b200", "librt_b200.so") if lib == "cur" else os.path.join(REPO, "raytracing_b200", "variants", f"librt_b200_{lib}.so"))
    env = dict(os.environ, RT_B200_LIB=path)
    r = subprocess.run([sys.executable, "-c", CHILD, name, worlds, fk or "-", thr or "-", top or "-"], capture_output=True, text=True, env=env)
    print(f"  {spec:24s} {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
