"""Development aid: quick timings of the main configurations on one GPU."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
from rt_b200 import abi, engine, host  # noqa: E402


def time_frames(ctx, scene, cam, film, sqrt_spp, depth, frames, warm=3):
    stream = torch.cuda.ExternalStream(ctx.stream)
    for i in range(warm):
        engine.render_accumulate(scene, cam, film, 0, 0, sqrt_spp, depth, 100 + i)
    ctx.synchronize()
    ctx.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(frames):
        engine.render_accumulate(scene, cam, film, 0, 0, sqrt_spp, depth, 1000 + i)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / frames
    c = ctx.counters()
    return ms, c.segments / frames


def main():
    ctx = engine.Context(0)
    for name, p0, W, depth in [("spheres", 11, 1920, 8), ("cornell", 0, 1080, 8), ("cornell_smoke", 0, 1080, 8),
                               ("final", 20, 1920, 8), ("spheres_textured", 500, 1920, 8)]:
        t0 = time.time()
        hs = host.HostScene.builtin(name, 1234, p0)
        t1 = time.time()
        scene = engine.Scene(ctx, hs.desc)
        t2 = time.time()
        info = scene.info()
        cfg = hs.camera_config(W, 1, depth)
        cam = engine.camera_from_config(cfg)
        film = engine.Film(ctx, cam.image_width, cam.image_height)
        ms, segs = time_frames(ctx, scene, cam, film, 1, depth, 10)
        npix = cam.image_width * cam.image_height
        img = film.read_rgb(1.0 / max(film.samples, 1))
        print(f"{name:18s} prims {info.n_prims:8d} nodes {info.n_nodes:8d} host-gen {t1 - t0:6.2f}s upload+build "
              f"{t2 - t1:6.2f}s (device build {info.build_ms:7.2f} ms) | {cam.image_width}x{cam.image_height} d{depth}: "
              f"{ms:8.3f} ms/frame {npix / ms / 1e3:9.1f} Mpath/s  {segs / npix:5.2f} seg/path  mean {img.mean():.4f}",
              flush=True)
        film.close()
        scene.close()
        hs.close()
    ctx.close()


if __name__ == "__main__":
    main()
