"""Development aid (GPU box): repeatability of the C3 static render (Cornell + smoke, 1920x1080, 1024 spp, depth 50)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
import torch  # noqa: E402
from rt_b200 import engine, host  # noqa: E402

ctx = engine.Context(0)
hs = host.HostScene.builtin("cornell_smoke", 1234, 0)
scene = engine.Scene(ctx, hs.desc)
cfg = hs.camera_config(1920, 1024, 50)
cfg.aspect_ratio = 16.0 / 9.0
cam = engine.camera_from_config(cfg)
film = engine.Film(ctx, cam.image_width, cam.image_height)
stream = torch.cuda.ExternalStream(ctx.stream)
engine.render_static(scene, cam, film, 2, 50, 1)
ctx.synchronize()
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    film.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    engine.render_static(scene, cam, film, 32, 50, 3)
    t1 = time.perf_counter()
    e1.record(stream)
    ctx.synchronize()
    t2 = time.perf_counter()
    c = ctx.counters()
    print("run %d: device %.1f ms, host submit %.1f ms, host total %.1f ms; graph launches %d instantiations %d" %
          (k, e0.elapsed_time(e1), (t1 - t0) * 1e3, (t2 - t0) * 1e3, c.graph_launches, c.graph_instantiations), flush=True)
