"""Small end-to-end run for compute-sanitizer: every kernel once on tiny inputs (build, trace, render with media,
lights, image textures, multi-rank films, resolve, scatter)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
from rt_b200 import abi, engine, host  # noqa: E402

ctx = engine.Context(0)
for name, p0, p1 in [("spheres", 6, -1), ("cornell_smoke", 0, -1), ("final", 3, 30), ("earth", 0, -1)]:
    hs = host.HostScene.builtin(name, 1234, p0, p1)
    scene = engine.Scene(ctx, hs.desc)
    cfg = hs.camera_config(64, 4, 12)
    cam = engine.camera_from_config(cfg)
    rays = (abi.rt_ray * 256)()
    for i, r in enumerate(rays):
        r.origin[:] = tuple(cfg.lookfrom)
        r.direction[:] = (cfg.lookat[0] - cfg.lookfrom[0] + 0.01 * (i % 16), cfg.lookat[1] - cfg.lookfrom[1] + 0.01 * (i // 16),
                          cfg.lookat[2] - cfg.lookfrom[2])
        r.t_min, r.t_max = 0.001, float("inf")
    for mode in (abi.RT_TRACE_EXACT_F64, abi.RT_TRACE_FAST_F32):
        scene.trace(rays, mode, 1)
    for rank, n_ranks in [(0, 1), (1, 3)]:
        film = engine.Film(ctx, cam.image_width, cam.image_height, rank, n_ranks, 4)
        engine.render_static(scene, cam, film, 2, 12, 5)
        engine.render_accumulate(scene, cam, film, 0, 1, 2, 12, 6)
        film.resolve_rgb8(0.2)
        img = film.read_rgb(0.2)
        assert np.isfinite(img).all()
        film.close()
    scene.close()
    hs.close()
    print(name, "ok", flush=True)
ctx.close()
