"""Development aid (GPU box): the same renders under different environment settings must give the same bits.

  python tools/env_check.py RT_TAIL=couple RT_TAIL=regroup [RT_B200_LIB=build/variants/librt_x.so ...]
Every argument is one run (comma-separated KEY=VALUE pairs) in its own process; prints a SHA of each image and
IDENTICAL / DIFFERENT.
"""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, hashlib
sys.path.insert(0, "%s/real-time-ray-tracing-engine_b200")
from rt_b200 import engine, host
ctx = engine.Context(0)
for name, p0, p1, w, d in (("spheres", 11, -1, 640, 8), ("cornell_smoke", 0, -1, 320, 50), ("final", 20, 1000, 640, 50),
                           ("spheres_textured", 100, -1, 480, 8), ("cornell", 0, -1, 200, 3)):
    hs = host.HostScene.builtin(name, 1234, p0, p1)
    scene = engine.Scene(ctx, hs.desc)
    cam = engine.camera_from_config(hs.camera_config(w, 4, d))
    film = engine.Film(ctx, cam.image_width, cam.image_height)
    engine.render_static(scene, cam, film, 2, d, 3)
    engine.render_accumulate(scene, cam, film, 0, 0, 1, d, 4)
    print(name, hashlib.sha256(film.read_rgb(1.0).tobytes()).hexdigest()[:16], ctx.counters().segments)
''' % REPO
outs = []
for setting in sys.argv[1:]:
    env = dict(os.environ)
    for kv in setting.split(","):
        k, v = kv.split("=", 1)
        env[k] = os.path.join(REPO, v) if k == "RT_B200_LIB" and not os.path.isabs(v) else v
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    print(setting, "::", r.stdout.strip().replace("\n", " | "), r.stderr[-300:])
    outs.append(r.stdout)
print("IDENTICAL" if outs and outs[0] and all(o == outs[0] for o in outs) else "DIFFERENT")
