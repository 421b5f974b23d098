"""Development aid (GPU box): wall time of rt_scene_create, three times in one process (the first call also pays
the one-off module loads), with RT_BUILD_TIMING=1 phase lines on stderr.

  RT_BUILD_TIMING=1 python tools/build_timing.py [scene p0]      (default: spheres_textured 500 = 1,000,002 spheres)
"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
from rt_b200 import engine, host  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "spheres_textured"
p0 = int(sys.argv[2]) if len(sys.argv) > 2 else 500
ctx = engine.Context(0)
t0 = time.perf_counter()
hs = host.HostScene.builtin(name, 1234, p0)
print("host scene description: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
for k in range(3):
    sys.stderr.write("--- build %d\n" % k)
    t0 = time.perf_counter()
    scene = engine.Scene(ctx, hs.desc)
    ctx.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    info = scene.info()
    print("rt_scene_create #%d: %.1f ms wall, %d primitives, %d nodes, device build %.2f ms" % (k, wall, info.n_prims, info.n_nodes, info.build_ms), flush=True)
    scene.close()
