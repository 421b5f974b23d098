"""Parity audit of the FP32 render path at the BASELINE configurations' own sizes (GPU box).

  python tools/audit_configs.py [c2 c3 c4 c5 ...] [--samples out.npz]

For every ray segment of one frame the FP64 parity traversal (the reference's arithmetic, bit-exact against the
reference itself in tests/) answers the very ray the FP32 extend kernel traces; prints one JSON line per
configuration with the mismatch tallies (rt_get_audit).  A measurement aid, not product code.
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
from rt_b200 import abi, engine, host  # noqa: E402

CONFIGS = {
    # name: (builtin scene, p0, p1, width, depth, aspect override, description)
    "c1": ("spheres", 11, -1, 400, 50, None, "spheres 400x225, depth 50"),
    "c2": ("spheres", 11, -1, 1920, 8, None, "spheres 1920x1080, depth 8"),
    "c3": ("cornell_smoke", 0, -1, 1920, 50, 16.0 / 9.0, "Cornell + smoke 1920x1080, depth 50"),
    "c3n": ("cornell", 0, -1, 1080, 50, None, "Cornell (boxes, no smoke) 1080x1080, depth 50"),
    "c4": ("spheres_textured", 500, -1, 1920, 8, None, "1M textured moving spheres 1920x1080, depth 8"),
    "c5": ("final", 20, 1000, 3840, 50, None, "final scene 3840x2160, depth 50"),
}


def audit_config(ctx, name, frames=1):
    scene_name, p0, p1, width, depth, aspect, what = CONFIGS[name]
    hs = host.HostScene.builtin(scene_name, 1234, p0, p1)
    cfg = hs.camera_config(width, 1, depth)
    if aspect:
        cfg.aspect_ratio = aspect
    cam = engine.camera_from_config(cfg)
    scene = engine.Scene(ctx, hs.desc)
    film = engine.Film(ctx, cam.image_width, cam.image_height)
    ctx.set_audit(True)
    for f in range(frames):
        engine.render_accumulate(scene, cam, film, 0, 0, 1, depth, 1000 + f)
    a = ctx.audit()
    samples = ctx.audit_samples()
    ctx.set_audit(False)
    d = hs.desc.contents
    n_sph, n_quad = d.n_spheres, d.n_quads

    def kind(i):
        return "miss" if i < 0 else ("sphere" if i < n_sph else ("quad" if i < n_sph + n_quad else "medium"))

    pairs = {}
    for s in samples:
        k = f"{kind(s.fast_prim)}->{kind(s.exact_prim)}" + ("" if s.bounce else " (primary)")
        pairs[k] = pairs.get(k, 0) + 1
    out = {"config": name, "what": what, "frames": frames, "segments": a.segments, "prim_mismatch": a.prim_mismatch,
           "mismatch_rate": a.prim_mismatch / max(1, a.segments), "primary_segments": a.primary_segments,
           "primary_mismatch": a.primary_mismatch, "primary_mismatch_rate": a.primary_mismatch / max(1, a.primary_segments),
           "hit_miss_flips": a.hit_miss_flips, "t_rel_above_1e-4": a.t_rel_above_1e4, "max_rel_t_error": a.max_rel_t_error,
           "rechecked": a.rechecked, "sampled_mismatch_kinds (fast->exact)": pairs}
    film.close()
    scene.close()
    hs.close()
    return out, samples


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    save = None
    if "--samples" in sys.argv:
        save = sys.argv[sys.argv.index("--samples") + 1]
        args = [a for a in args if a != save]
    ctx = engine.Context(0)
    kept = {}
    for name in args or ["c2", "c3", "c4", "c5"]:
        out, samples = audit_config(ctx, name)
        print(json.dumps(out), flush=True)
        if save and samples:
            kept[name] = np.frombuffer(b"".join(bytes(s) for s in samples), dtype=np.uint8)
    if save and kept:
        np.savez_compressed(save, **kept)
    ctx.close()


if __name__ == "__main__":
    main()
