"""Measures the BASELINE.json configurations on the GPU(s) and prints one JSON line per configuration.

Single process:      python tools/run_configs.py [c1 c2 c3 c4 c5 ...]
One process per GPU: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_configs.py c5

  c1  spheres scene 400x225, 100 spp, depth 50 (static render) + RMSE against the reference's own CPU render
  c2  spheres scene 1920x1080, 1 spp, depth 8 (one dynamic-mode frame)
  c3  Cornell box with two smoke volumes and the quad light, 1920x1080, 1024 spp, depth 50
  c4  1,000,002-sphere scene (Perlin / checker textures, motion blur, depth of field) 1920x1080: 1 spp depth 8
      and 64 spp depth 50; BVH build time
  c5  "final" scene 3840x2160, 4096 spp, depth 50, image tiles over all ranks + framebuffer gather
Times are device times (CUDA events on the context stream), max over ranks; scene upload / BVH build excluded
and reported separately.  A measurement aid, not product code: the c1 leg also times the reference's own CPU
render (oracle/_ref through tests/oracle_lib.py) and reports the RMSE against it - the checker role the oracle
has in tests/ and in bench.py's cpu_baseline, nothing the GPU path depends on.  `--quick` divides the sample counts of c3/c5 by 16 for smoke runs.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
from rt_b200 import abi, distributed, engine, host  # noqa: E402

TILE_ROWS = 8


def setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def render(ctx, stream, scene, cam, sqrt_spp, depth, seed, world, rank, frames=1, dynamic=False):
    """Renders and returns (ms, full image float32 [H,W,3] on rank 0 or None, segments/path)."""
    W, H = cam.image_width, cam.image_height
    owned = distributed.owned_pixels(W, H, rank, world, TILE_ROWS)
    accum = torch.zeros((owned, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    film = engine.Film(ctx, W, H, rank, world, TILE_ROWS, external_accum=accum.data_ptr())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ctx.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if dynamic:
        for f in range(frames):
            engine.render_accumulate(scene, cam, film, 0, 0, 1, depth, seed + f)
        n_samples = frames
    else:
        engine.render_static(scene, cam, film, sqrt_spp, depth, seed)
        n_samples = sqrt_spp * sqrt_spp
    full = None
    with torch.cuda.stream(stream):
        if world > 1:
            gathered = distributed.gather_film(accum, W, H, TILE_ROWS, dst=0)
            if rank == 0:
                full = torch.empty((W * H, 4), dtype=torch.float32, device="cuda")
                abi.check(ctx.lib, ctx.lib.rt_film_scatter_gathered(ctx._h, W, H, world, TILE_ROWS, gathered.data_ptr(),
                                                                    full.data_ptr()), "rt_film_scatter_gathered")
        else:
            full = accum
    e1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    c = ctx.counters()
    seg = torch.tensor([float(c.segments), float(c.paths)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(seg)
    img = None
    if rank == 0:
        img = (full[:, :3] / n_samples).reshape(H, W, 3).cpu().numpy()
    film.close()
    return float(ms.item()), img, float(seg[0] / max(seg[1], 1))


def reference_render(name, p0, width, spp, depth, seed):
    """The reference's CPU render of the same configuration (oracle/_ref, all host threads)."""
    import oracle_lib as ol

    if not ol.have_ref():
        return None, None, None
    r = ol.ref()
    h = r.ref_scene_build(name.encode(), 1234, p0, 0)
    cfg = abi.rt_camera_config()
    r.ref_scene_camera_config(h, width, spp, depth, cfg)
    cam = abi.rt_camera()
    r.ref_camera_init(cfg, cam)
    n = cam.image_width * cam.image_height
    img = (C.c_double * (n * 3))()
    seg = C.c_uint64()
    cores = r.ref_hardware_threads()
    r.ref_render(h, 32, 1, 2, 1, 1, 1, 0, -1, -1, img, C.byref(seg))  # build the BVH outside the timed call
    secs = r.ref_render(h, width, spp, depth, seed, 1, cores, 0, cam.image_height, -1, img, C.byref(seg))
    out = np.nan_to_num(np.frombuffer(img, dtype=np.float64).reshape(cam.image_height, cam.image_width, 3).copy())
    r.ref_scene_free(h)
    return out, secs, cores


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    quick = "--quick" in sys.argv
    which = args or ["c1", "c2", "c3", "c4"]
    world, rank, local = setup()
    ctx = engine.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    lum = np.array([0.2126, 0.7152, 0.0722])

    def emit(d):
        if rank == 0:
            d["n_gpus"] = world
            print(json.dumps(d), flush=True)

    def load(name, p0=0, p1=-1):
        t0 = time.time()
        hs = host.HostScene.builtin(name, 1234, p0, p1)
        t1 = time.time()
        sc = engine.Scene(ctx, hs.desc)
        info = sc.info()
        return hs, sc, {"primitives": info.n_prims, "nodes4": info.n_nodes, "host_generate_s": round(t1 - t0, 3),
                        "upload_build_s": round(time.time() - t1, 3), "device_build_ms": round(info.build_ms, 3)}

    for cfg_name in which:
        if cfg_name == "c1":
            hs, sc, binfo = load("spheres", 11)
            cam = engine.camera_from_config(hs.camera_config(400, 100, 50))
            render(ctx, stream, sc, cam, 10, 50, 1, world, rank)  # warm-up
            ms, img, spp_seg = render(ctx, stream, sc, cam, 10, 50, 7, world, rank)
            out = {"config": "c1 spheres 400x225 100spp depth50 static", "ms": ms, "mpath_s": 400 * 225 * 100 / ms / 1e3,
                   "segments_per_path": spp_seg, "bvh": binfo}
            if rank == 0:
                ref_a, secs, cores = reference_render("spheres", 11, 400, 100, 50, 11)
                if ref_a is not None:
                    ref_b, _, _ = reference_render("spheres", 11, 400, 100, 50, 100011)
                    clip = lambda x: np.minimum(x, 4.0)
                    floor = float(np.sqrt(np.mean((clip(ref_a) - clip(ref_b)) ** 2)))
                    rmse = float(np.sqrt(np.mean((clip(img.astype(np.float64)) - clip(ref_a)) ** 2)))
                    l_ref, l_gpu = float((ref_a @ lum).mean()), float((img.astype(np.float64) @ lum).mean())
                    out.update({"reference_cpu_s": secs, "reference_cores": cores,
                                "reference_mpath_s": 400 * 225 * 100 / secs / 1e6, "rmse_vs_reference": rmse,
                                "rmse_reference_vs_reference": floor, "rmse_ratio": rmse / floor,
                                "mean_luminance_gpu": l_gpu, "mean_luminance_reference": l_ref,
                                "luminance_rel_diff": (l_gpu - l_ref) / l_ref})
            emit(out)
            sc.close()
        elif cfg_name == "c2":
            hs, sc, binfo = load("spheres", 11)
            cam = engine.camera_from_config(hs.camera_config(1920, 1, 8))
            render(ctx, stream, sc, cam, 1, 8, 1, world, rank, frames=5, dynamic=True)
            ms, img, spp_seg = render(ctx, stream, sc, cam, 1, 8, 100, world, rank, frames=50, dynamic=True)
            emit({"config": "c2 spheres 1920x1080 1spp depth8 dynamic frame", "ms_per_frame": ms / 50,
                  "mpath_s": 1920 * 1080 / (ms / 50) / 1e3, "segments_per_path": spp_seg, "bvh": binfo})
            sc.close()
        elif cfg_name == "c3":
            hs, sc, binfo = load("cornell_smoke")
            root = 8 if quick else 32
            cfg = hs.camera_config(1920, root * root, 50)
            cfg.aspect_ratio = 16.0 / 9.0  # BASELINE config 3: 1080p (the scene's native aspect is 1.0)
            cam = engine.camera_from_config(cfg)
            # warm-up with full-size passes (36 strata: a long render, 64 M-path passes): the queue storage must not grow -
            # cudaFree + cudaMalloc of several GB, 20 ... 470 ms depending on the box - inside the timed render
            render(ctx, stream, sc, cam, 3 if quick else 6, 50, 1, world, rank)
            ms, img, spp_seg = render(ctx, stream, sc, cam, root, 50, 3, world, rank)
            paths = cam.image_width * cam.image_height * root * root
            emit({"config": f"c3 cornell+smoke {cam.image_width}x{cam.image_height} {root * root}spp depth50 static", "ms": ms,
                  "mpath_s": paths / ms / 1e3, "segments_per_path": spp_seg, "bvh": binfo,
                  "mean_luminance": float((img @ lum).mean()) if img is not None else None})
            sc.close()
        elif cfg_name == "c4":
            hs, sc, binfo = load("spheres_textured", 500)
            cam = engine.camera_from_config(hs.camera_config(1920, 1, 8))
            render(ctx, stream, sc, cam, 1, 8, 1, world, rank, frames=3, dynamic=True)
            ms, img, spp_seg = render(ctx, stream, sc, cam, 1, 8, 100, world, rank, frames=20, dynamic=True)
            cam2 = engine.camera_from_config(hs.camera_config(1920, 64, 50))
            render(ctx, stream, sc, cam2, 2, 50, 1, world, rank)  # warm-up: sizes the queues for full passes
            runs = [render(ctx, stream, sc, cam2, 8, 50, 5 + k, world, rank) for k in range(3)]
            ms2, img2, spp_seg2 = min(runs, key=lambda r: r[0])  # first-use allocation / box noise: report the best of 3
            emit({"config": "c4 1M textured moving spheres 1920x1080", "frame_1spp_depth8_ms": ms / 20,
                  "frame_mpath_s": 1920 * 1080 / (ms / 20) / 1e3, "segments_per_path": spp_seg,
                  "static_64spp_depth50_ms": ms2, "static_64spp_depth50_ms_all_runs": [r[0] for r in runs],
                  "static_mpath_s": 1920 * 1080 * 64 / ms2 / 1e3,
                  "static_segments_per_path": spp_seg2, "bvh": binfo})
            sc.close()
        elif cfg_name == "c5":
            hs, sc, binfo = load("final", 20, 1000)
            root = 16 if quick else 64
            cam = engine.camera_from_config(hs.camera_config(3840, root * root, 50))
            render(ctx, stream, sc, cam, 4 if quick else 8, 50, 1, world, rank)  # warm-up: 64 strata = full-size passes on up to 8 GPUs
            ms, img, spp_seg = render(ctx, stream, sc, cam, root, 50, 9, world, rank)
            paths = cam.image_width * cam.image_height * root * root
            emit({"config": f"c5 final scene {cam.image_width}x{cam.image_height} {root * root}spp depth50 static, tiles over "
                            f"{world} GPU(s) + gather", "ms": ms, "mpath_s": paths / ms / 1e3, "segments_per_path": spp_seg,
                  "bvh": binfo, "mean_luminance": float((img @ lum).mean()) if img is not None else None})
            sc.close()
        hs.close()
    del stream
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
