"""One GPU renders 1/world of the frame (scanline partition, rank 0): RT_OPT_FRAME_KERNEL 1 (one persistent kernel per frame,
CTA-private queues) against 0 (one kernel per phase, graph + PDL).  usage: frame_kernel_probe.py [scene] [worlds]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from raytracing_b200 import capi, scene_io
from raytracing_b200.camera import default_camera
name = sys.argv[1] if len(sys.argv) > 1 else "CornellBox"
worlds = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
w, h, mb = (3840, 2160, 16) if name == "CornellBox_Dragon" else (1920, 1080, 8)
sc = scene_io.load_scene(name)
for world in worlds:
    ctx = capi.Context(w, h, device=0, rank=0, world=world)
    ctx.upload_scene(sc); ctx.set_camera(default_camera(w, h))
    stream = torch.cuda.ExternalStream(ctx.stream_handle())
    res = {}
    for fk in (0, 1):
        ctx.set_option(capi.OPT_FRAME_KERNEL, fk)
        for _ in range(5):
            ctx.reset(); ctx.integrate(mb)
        ctx.sync()
        n = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            ctx.reset(); ctx.integrate(mb)
        e1.record(stream); ctx.sync(); torch.cuda.synchronize()
        res[fk] = e0.elapsed_time(e1) / n
    print(f"{name} world {world:2d}: phase kernels {res[0]:.3f} ms/frame, frame kernel {res[1]:.3f} ms/frame, ratio {res[0] / res[1]:.2f}x", flush=True)
    ctx.destroy()
