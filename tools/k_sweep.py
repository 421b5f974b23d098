"""Development aid: frame time of the spheres scene at several sizes against the wavefront depth (RT_WAVE_BOUNCES)."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys
sys.path.insert(0, os.path.join(%r, "real-time-ray-tracing-engine_b200"))
sys.path.insert(0, os.path.join(%r, "tools"))
from rt_b200 import engine, host
from quick_gpu import time_frames
ctx = engine.Context(0)
out = []
for p0 in (11, 16, 22, 32, 64, 128):
    hs = host.HostScene.builtin("spheres", 1234, p0)
    scene = engine.Scene(ctx, hs.desc)
    cfg = hs.camera_config(1920, 1, 8); cam = engine.camera_from_config(cfg)
    film = engine.Film(ctx, cam.image_width, cam.image_height)
    ms, segs = time_frames(ctx, scene, cam, film, 1, 8, 10)
    out.append("%%d:%%.3f" %% (scene.info().n_prims, ms))
    film.close(); scene.close(); hs.close()
print(" ".join(out))
""" % (REPO, REPO)
for k in sys.argv[1:] or ["1", "2", "3"]:
    env = dict(os.environ, RT_WAVE_BOUNCES=k)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print("K=" + k, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:])
