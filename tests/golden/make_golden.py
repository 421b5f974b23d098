"""Generates the golden fixtures in this directory from the UNMODIFIED reference
(oracle/_ref/libref_harness.so, built by oracle/Makefile from /root/reference/src).

Run in the build container, where /root/reference exists:   python tests/golden/make_golden.py
The fixtures are what lets machines without the reference (the GPU box) check that the oracle still
reproduces the reference bit for bit:
  rng.npz            std::mt19937 -> random_double / random_int sequences (Utility.hpp:16-37)
  tonemap.npz        to_byte over a sweep of inputs (ColorUtility.hpp:18-26)
  scenes.json        SHA-256 of every array of the built-in scenes' flat descriptions + counts
  <scene>.npz        camera (CameraConfig -> derived camera), primary rays of one stratum, closest hits of
                     the primary rays and of every segment of a small render (list and BVH world), and the
                     render itself (linear FP64 RGB) - all produced by reference code
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
from rt_b200 import abi  # noqa: E402

SCENES = {
    # name: (builtin, p0, p1, width, spp, depth)
    "spheres": ("spheres", 11, 0, 64, 4, 8),
    "spheres_textured": ("spheres_textured", 6, 0, 48, 4, 8),
    "cornell": ("cornell", 0, 0, 40, 4, 8),
    "cornell_smoke": ("cornell_smoke", 0, 0, 40, 4, 8),
    "final": ("final", 4, 40, 48, 4, 8),
}
SCENE_SEED = 1234
RAY_SEED = 77
TRACE_SEED = 5
RENDER_SEED = 99


def sha_parts(desc):
    return [hashlib.sha256(p).hexdigest() for p in ol.desc_bytes(desc)]


def main():
    r = ol.ref()
    o = ol.oracle()
    o.ora_ray_log.argtypes = [C.POINTER(abi.rt_ray), C.c_int64]
    o.ora_ray_log_count.restype = C.c_int64

    # --- RNG ---
    r.ref_seed(SCENE_SEED)
    canon = np.array([r.ref_random_double() for _ in range(256)])
    ints = {}
    for lo, hi in [(0, 1), (0, 6), (0, 255), (3, 1000), (0, 2**31 - 1)]:
        r.ref_seed(SCENE_SEED + hi)
        ints[f"int_{lo}_{hi}"] = np.array([r.ref_random_int(lo, hi) for _ in range(128)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "rng.npz"), canonical=canon, **ints)

    # --- tonemap ---
    xs = np.concatenate([np.linspace(-0.5, 1.5, 4001), (np.arange(0, 257) / 256.0) ** 2,
                         np.nextafter((np.arange(1, 257) / 256.0) ** 2, 0), [np.nan, np.inf, -np.inf, 0.0, -0.0]])
    np.savez_compressed(os.path.join(HERE, "tonemap.npz"), x=xs,
                        byte=np.array([r.ref_to_byte(float(x)) for x in xs], dtype=np.uint8))

    # --- scenes ---
    index = {}
    for key, (name, p0, p1, W, spp, depth) in SCENES.items():
        h = r.ref_scene_build(name.encode(), SCENE_SEED, p0, p1)
        desc = r.ref_scene_desc(h).contents
        index[key] = {"builtin": name, "seed": SCENE_SEED, "p0": p0, "p1": p1, "width": W, "spp": spp, "depth": depth,
                      "counts": {f: getattr(desc, f) for f, _ in abi.rt_scene_desc._fields_ if f.startswith("n_")},
                      "sha256": sha_parts(desc)}
        cfg = abi.rt_camera_config()
        r.ref_scene_camera_config(h, W, spp, depth, cfg)
        cam = abi.rt_camera()
        r.ref_camera_init(cfg, cam)
        n = cam.image_width * cam.image_height
        rays = (abi.rt_ray * n)()
        r.ref_primary_rays(h, W, spp, RAY_SEED, 1, 0, rays)
        out = {"camera_config": np.frombuffer(bytes(cfg), dtype=np.uint8), "camera": np.frombuffer(bytes(cam), dtype=np.uint8),
               "primary_rays": np.frombuffer(bytes(rays), dtype=np.uint8)}
        for use_bvh in (0, 1):
            hits = (abi.rt_hit * n)()
            r.ref_seed(TRACE_SEED)
            r.ref_trace(h, rays, n, use_bvh, hits)
            out[f"primary_hits_bvh{use_bvh}"] = np.frombuffer(bytes(hits), dtype=np.uint8)
        # every segment of a small render: rays logged by the oracle (bit-identical to the reference's
        # render, which make_golden checks below), hits answered by the reference
        sc = o.ora_scene_create(C.byref(desc))
        cap = n * spp * depth
        log = (abi.rt_ray * cap)()
        o.ora_ray_log(log, cap)
        img_o = (C.c_double * (n * 3))()
        o.ora_render(sc, cfg, ol.ORA_RNG_MT19937, ol.ORA_SAMPLER_REJECTION, RENDER_SEED, 0, 0, cam.image_height, -1, img_o, None)
        n_seg = o.ora_ray_log_count()
        o.ora_ray_log(None, 0)
        o.ora_scene_destroy(sc)
        for use_bvh in (0, 1):
            img = (C.c_double * (n * 3))()
            seg = C.c_uint64()
            r.ref_render(h, W, spp, depth, RENDER_SEED, use_bvh, 1, 0, cam.image_height, -1, img, C.byref(seg))
            out[f"render_bvh{use_bvh}"] = np.frombuffer(bytes(img), dtype=np.float64)
            out[f"render_segments_bvh{use_bvh}"] = np.array([seg.value])
        assert np.array_equal(out["render_bvh0"], np.frombuffer(bytes(img_o), dtype=np.float64), equal_nan=True), key
        seg_rays = (abi.rt_ray * n_seg).from_buffer_copy(bytes(log)[: n_seg * C.sizeof(abi.rt_ray)])
        has_media = desc.n_media > 0
        if not has_media:  # media consume the sequential RNG inside hit(): segment hits are not replayable
            hits = (abi.rt_hit * n_seg)()
            r.ref_trace(h, seg_rays, n_seg, 0, hits)
            out["segment_rays"] = np.frombuffer(bytes(seg_rays), dtype=np.uint8)
            out["segment_hits"] = np.frombuffer(bytes(hits), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, f"{key}.npz"), **out)
        r.ref_scene_free(h)
        print(key, "ok:", n, "primary rays,", n_seg, "segments")

    # the large-scene generators are pinned by checksum only
    for key, (name, p0, p1) in {"spheres_30": ("spheres", 30, 0), "spheres_textured_40": ("spheres_textured", 40, 0),
                                "final_full": ("final", 20, 1000)}.items():
        h = r.ref_scene_build(name.encode(), SCENE_SEED, p0, p1)
        desc = r.ref_scene_desc(h).contents
        index[key] = {"builtin": name, "seed": SCENE_SEED, "p0": p0, "p1": p1,
                      "counts": {f: getattr(desc, f) for f, _ in abi.rt_scene_desc._fields_ if f.startswith("n_")},
                      "sha256": sha_parts(desc)}
        r.ref_scene_free(h)
    with open(os.path.join(HERE, "scenes.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
