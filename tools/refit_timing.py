"""Development aid: time rt_scene_update_spheres (re-bake + scatter + BVH4 refit) against a full rebuild."""
import ctypes as C
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
from rt_b200 import abi, engine, host  # noqa: E402

ctx = engine.Context(0)
for name, p0 in [("spheres", 11), ("spheres_textured", 500)]:
    hs = host.HostScene.builtin(name, 1234, p0)
    d = hs.desc.contents
    scene = engine.Scene(ctx, hs.desc)
    for count in (100, d.n_spheres - 3):
        count = min(count, d.n_spheres - 3)
        sph = (abi.rt_sphere * count)(*[d.spheres[3 + i] for i in range(count)])
        for s in sph:
            s.center0[1] += 0.05
        scene.update_spheres(3, sph)  # first call allocates the links
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            scene.update_spheres(3, sph)
        ms = (time.perf_counter() - t0) / reps * 1e3
        print(f"{name}: {d.n_spheres} spheres, update of {count}: {ms:.3f} ms per call (host re-bake + upload + device refit)")
    t0 = time.perf_counter()
    s2 = engine.Scene(ctx, hs.desc)
    print(f"{name}: full rebuild {1e3 * (time.perf_counter() - t0):.1f} ms wall (device build {s2.info().build_ms:.2f} ms)")
    s2.close()
    scene.close()
    hs.close()
ctx.close()
