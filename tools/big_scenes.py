import os, sys
sys.path.insert(0, 'real-time-ray-tracing-engine_b200'); sys.path.insert(0, 'tools')
from rt_b200 import engine, host
from quick_gpu import time_frames
ctx = engine.Context(0)
for name, p0 in (("spheres_textured", 500), ("spheres", 128), ("final", 20)):
    hs = host.HostScene.builtin(name, 1234, p0)
    scene = engine.Scene(ctx, hs.desc)
    cam = engine.camera_from_config(hs.camera_config(1920, 1, 8))
    film = engine.Film(ctx, cam.image_width, cam.image_height)
    ms, _ = time_frames(ctx, scene, cam, film, 1, 8, 10)
    print(os.environ.get("RT_B200_LIB", "default").split("/")[-1], name, scene.info().n_prims, round(ms, 3))
