"""Development aid: render a few progressive frames of one built-in scene (for ncu captures)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
from rt_b200 import engine, host  # noqa: E402

name, p0, width, depth, frames = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
ctx = engine.Context(0)
hs = host.HostScene.builtin(name, 1234, p0)
scene = engine.Scene(ctx, hs.desc)
cam = engine.camera_from_config(hs.camera_config(width, 1, depth))
film = engine.Film(ctx, cam.image_width, cam.image_height)
for f in range(frames):
    engine.render_accumulate(scene, cam, film, 0, 0, 1, depth, 100 + f)
ctx.synchronize()
print("frames", film.samples, "mean", float(film.read_rgb(1.0 / film.samples).mean()))
