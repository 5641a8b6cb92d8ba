import sys, os, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from raytracing_b200 import capi, scene_io
from raytracing_b200.camera import default_camera
w, h, mb = 1920, 1080, 8
for name in ("CornellBox", "ShaderBalls"):
    sc = scene_io.load_scene(name)
    ref = {}
    for world in (1, 8):
        for ov in (0, 1, 2):
            ctx = capi.Context(w, h, device=0, rank=0, world=world)
            ctx.upload_scene(sc); ctx.set_camera(default_camera(w, h))
            ctx.set_option(capi.OPT_OVERLAP, ov)
            stream = torch.cuda.ExternalStream(ctx.stream_handle())
            for _ in range(5):
                ctx.reset(); ctx.integrate(mb)
            ctx.sync()
            rad = ctx.read_radiance()
            if world in ref:
                same = np.array_equal(rad.view(np.uint32), ref[world].view(np.uint32))
            else:
                ref[world] = rad; same = True
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(50):
                ctx.reset(); ctx.integrate(mb)
            e1.record(stream); ctx.sync(); torch.cuda.synchronize()
            print(f"{name} world {world} overlap {ov}: {e0.elapsed_time(e1)/50:.3f} ms/frame same={same}", flush=True)
            ctx.destroy()
