"""Development aid: stage times of a static render as a function of the depth limit."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
from rt_b200 import engine, host  # noqa: E402

name, p0, width, root = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctx = engine.Context(0)
hs = host.HostScene.builtin(name, 1234, p0)
scene = engine.Scene(ctx, hs.desc)
for depth in [int(d) for d in sys.argv[5:]]:
    cam = engine.camera_from_config(hs.camera_config(width, root * root, depth))
    film = engine.Film(ctx, cam.image_width, cam.image_height)
    engine.render_static(scene, cam, film, 1, depth, 1)
    ctx.synchronize()
    ctx.set_stage_timing(True)
    ctx.reset_counters()
    engine.render_static(scene, cam, film, root, depth, 2)
    ms, n = ctx.stage_times()
    ctx.set_stage_timing(False)
    c = ctx.counters()
    print(f"{name} depth {depth:3d}: total {sum(ms):9.2f} ms  " +
          " ".join(f"{k}={v:.2f}({int(m)})" for k, v, m in zip(["gen", "ext", "shade", "acc", "tail"], ms, n)) +
          f"  seg/path {c.segments / c.paths:.3f} tail segs {c.tail_segments}", flush=True)
    film.close()
