"""Development aid: the two-rays-per-lane extend kernel (RT_EXTEND_MUX=1) must render the default kernel's bits."""
import os, subprocess, sys, hashlib
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, hashlib
sys.path.insert(0, "%s/real-time-ray-tracing-engine_b200")
from rt_b200 import engine, host
ctx = engine.Context(0)
for name, p0, p1, w, d in (("spheres", 11, -1, 640, 8), ("cornell_smoke", 0, -1, 320, 12), ("final", 20, 1000, 640, 8)):
    hs = host.HostScene.builtin(name, 1234, p0, p1)
    scene = engine.Scene(ctx, hs.desc)
    cam = engine.camera_from_config(hs.camera_config(w, 4, d))
    film = engine.Film(ctx, cam.image_width, cam.image_height)
    engine.render_static(scene, cam, film, 2, d, 3)
    engine.render_accumulate(scene, cam, film, 0, 0, 1, d, 4)
    print(name, hashlib.sha256(film.read_rgb(1.0).tobytes()).hexdigest()[:16], ctx.counters().segments)
''' % REPO
outs = []
for mux in ("0", "1"):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RT_EXTEND_MUX=mux), capture_output=True, text=True, timeout=600)
    print("mux", mux, r.stdout.strip().replace("\n", " | "), r.stderr[-300:])
    outs.append(r.stdout)
print("IDENTICAL" if outs[0] == outs[1] and outs[0] else "DIFFERENT")
