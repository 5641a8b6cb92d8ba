"""One GPU renders 1/world of the frame (scanline partition) to see what the fixed per-frame cost is."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from raytracing_b200 import capi, scene_io
from raytracing_b200.camera import default_camera
w, h, mb = 1920, 1080, 8
sc = scene_io.load_scene("CornellBox")
for world in (1, 2, 4, 8, 16):
    ctx = capi.Context(w, h, device=0, rank=0, world=world)
    ctx.upload_scene(sc); ctx.set_camera(default_camera(w, h))
    stream = torch.cuda.ExternalStream(ctx.stream_handle())
    for graph in (1, 0):
        ctx.set_option(capi.OPT_GRAPH, graph)
        for _ in range(5):
            ctx.reset(); ctx.integrate(mb)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(50):
            ctx.reset(); ctx.integrate(mb)
        e1.record(stream); ctx.sync(); torch.cuda.synchronize()
        print(f"world {world:2d} graph {graph}: {e0.elapsed_time(e1)/50:.3f} ms/frame", flush=True)
    ctx.set_option(capi.OPT_KERNEL_TIMING, 1); ctx.set_option(capi.OPT_OVERLAP, 0)
    for _ in range(3):
        ctx.reset(); ctx.integrate(mb)
    ctx.sync(); ctx.kernel_times()
    for _ in range(10):
        ctx.reset(); ctx.integrate(mb)
    ctx.sync()
    kt = ctx.kernel_times()
    print("   per-kernel ms/frame:", {k: (round(v[0] / 10, 4), v[1] // 10) for k, v in kt.items() if v[1]}, flush=True)
    ctx.destroy()
