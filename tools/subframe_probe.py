"""One GPU, k sub-frames: a multi-device context that lists device 0 k times renders k scanline partitions side by side (k streams,
each persistent kernel sized to 1/k of the resident CTA slots), so that the tail of one partition's kernel overlaps the body of
another's.  ms per whole frame for k = 1, 2, 4, 8.   usage: subframe_probe.py scene [fk] [threads]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from raytracing_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else "Synthetic10M"
fk = int(sys.argv[2]) if len(sys.argv) > 2 else -1
thr = int(sys.argv[3]) if len(sys.argv) > 3 else -1
_, w, h, mb = bench.WORKLOADS[name]
scene, cam = bench.load_workload_scene(name, w, h)
out = []
for k in (1, 2, 4, 8):
    ctx = capi.Context(w, h, devices=[0] * k) if k > 1 else capi.Context(w, h)
    if fk >= 0:
        ctx.set_option(capi.OPT_FRAME_KERNEL, fk)
    if thr >= 0:
        ctx.set_option(capi.OPT_FRAME_THREADS, thr)
    ctx.upload_scene(scene); ctx.set_camera(cam)
    best = 1e9
    for rep in range(3):
        for _ in range(4):
            ctx.reset(); ctx.integrate(mb)
        ctx.sync()
        n = 15
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            ctx.reset(); ctx.integrate(mb)
        ctx.sync(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    out.append(f"k={k}: {best:.3f}")
    ctx.destroy()
print(f"{name} fk={fk} threads={thr}: ms/frame  " + "  ".join(out), flush=True)
