"""Renders a few frames of a workload through the C ABI and nothing else (the command ncu profiles: profiles/run_ncu.sh).
usage: render_frames.py scene [frames] [frame_kernel] [world] [copies] [overlap]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from raytracing_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else "CornellBox"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
fk = int(sys.argv[3]) if len(sys.argv) > 3 else 1
world = int(sys.argv[4]) if len(sys.argv) > 4 else 1
copies = int(sys.argv[5]) if len(sys.argv) > 5 else 183
overlap = int(sys.argv[6]) if len(sys.argv) > 6 else 2
_, w, h, mb = bench.WORKLOADS[name]
scene, cam = bench.load_workload_scene(name, w, h, copies)
ctx = capi.Context(w, h, device=0, rank=0, world=world)
ctx.set_option(capi.OPT_FRAME_KERNEL, fk)
ctx.set_option(capi.OPT_OVERLAP, overlap)
ctx.upload_scene(scene); ctx.set_camera(cam)
for _ in range(frames):
    ctx.reset(); ctx.integrate(mb)
ctx.sync()
st = ctx.frame_stats()
print(name, "rays/frame", int(st["n_ext"][: mb + 1].sum() + st["n_shadow"][: mb + 1].sum()))
ctx.destroy()
