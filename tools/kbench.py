"""Development aid (GPU box): frame times of the main scenes with the stage split, for one library.

  python tools/kbench.py [--frames 30] [--scenes c2,cornell,final,c4] [--stats]
  RT_B200_LIB=build/variants/librt_x.so python tools/kbench.py     (one variant)
  python tools/kbench.py --all                                      (default library + every build/variants/librt_*.so,
                                                                     each in its own process)
Prints one line per (library, scene): ms per 1080p depth-8 frame (CUDA events on the context stream, L2 flushed by a
192 MiB write before every frame) and the per-stage milliseconds of a second, stage-timed run.
"""
import glob
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))

SCENES = {"c2": ("spheres", 11, -1, 1920, 8), "cornell": ("cornell", 0, -1, 1080, 8),
          "smoke": ("cornell_smoke", 0, -1, 1080, 8), "final": ("final", 20, 1000, 1920, 8),
          "c4": ("spheres_textured", 500, -1, 1920, 8), "s14k": ("spheres", 60, -1, 1920, 8),
          "c1": ("spheres", 11, -1, 400, 50)}


def run_one(args):
    import torch

    from rt_b200 import engine, host

    frames = int(args[args.index("--frames") + 1]) if "--frames" in args else 30
    names = args[args.index("--scenes") + 1].split(",") if "--scenes" in args else ["c2", "cornell", "final", "c4"]
    stats = "--stats" in args
    ctx = engine.Context(0)
    if "--graph" in args:
        ctx.set_graph(True)
        lib_suffix = "+graph"
    else:
        lib_suffix = ""
    stream = torch.cuda.ExternalStream(ctx.stream)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    lib = os.path.basename(os.environ.get("RT_B200_LIB", "default")).replace("librt_", "").replace(".so", "") + lib_suffix
    for name in names:
        scene_name, p0, p1, width, depth = SCENES[name]
        hs = host.HostScene.builtin(scene_name, 1234, p0, p1)
        scene = engine.Scene(ctx, hs.desc)
        sqrt_spp = 10 if name == "c1" else 1
        cam = engine.camera_from_config(hs.camera_config(width, sqrt_spp * sqrt_spp, depth))
        film = engine.Film(ctx, cam.image_width, cam.image_height)

        def frame(f):
            if name == "c1":
                engine.render_static(scene, cam, film, sqrt_spp, depth, 1000 + f)
            else:
                engine.render_accumulate(scene, cam, film, 0, 0, 1, depth, 1000 + f)

        for f in range(3):
            frame(f)
        ctx.synchronize()
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
        ctx.reset_counters()
        for f, (e0, e1) in enumerate(pairs):
            with torch.cuda.stream(stream):
                flush.fill_(f & 255)
            e0.record(stream)
            frame(f)
            e1.record(stream)
        ctx.synchronize()
        # host-blocking frames: submit, wait for the GPU, submit the next (what a viewer with no frame in flight pays)
        import time as _time
        t0 = _time.perf_counter()
        for f in range(frames):
            frame(f)
            ctx.synchronize()
        blocking_ms = (_time.perf_counter() - t0) / frames * 1e3
        times = sorted(e0.elapsed_time(e1) for e0, e1 in pairs)
        ms, best = sum(times) / len(times), times[0]
        c = ctx.counters()
        ctx.set_stage_timing(True)
        for f in range(10):
            frame(f)
        st, sn = ctx.stage_times()
        ctx.set_stage_timing(False)
        info = scene.info()
        out = {"lib": lib, "bvh": os.environ.get("RT_BVH", "auto") + ":" + ["none", "sah", "ploc", "lbvh"][info.builder],
               "build_ms": round(info.build_ms, 2), "depth": info.depth, "scene": name, "ms": round(ms, 4), "min_ms": round(best, 4), "blocking_ms": round(blocking_ms, 4),
               "seg_per_path": round(c.segments / max(1, c.paths), 3),
               "stages": {k: round(st[i] / 10, 4) for i, k in enumerate(["gen", "extend", "shade", "accum", "tail"])}}
        if stats:
            ctx.set_stats(True)
            ctx.reset_counters()
            frame(0)
            c = ctx.counters()
            ctx.set_stats(False)
            out["nodes_per_seg"] = round(c.nodes_visited / max(1, c.segments), 3)
            out["tests_per_seg"] = round(c.prim_tests / max(1, c.segments), 3)
            out["queues"] = ctx.queue_lengths(depth + 1)[:6]
        print(json.dumps(out), flush=True)
        film.close()
        scene.close()
        hs.close()
    del stream, flush
    ctx.close()


def main():
    args = sys.argv[1:]
    if "--all" in args:
        rest = [a for a in args if a != "--all"]
        for lib in [""] + sorted(glob.glob(os.path.join(REPO, "build", "variants", "librt_*.so"))):
            env = dict(os.environ)
            if lib:
                env["RT_B200_LIB"] = lib
            else:
                env.pop("RT_B200_LIB", None)
            subprocess.run([sys.executable, os.path.abspath(__file__)] + rest, env=env)
        return
    run_one(args)


if __name__ == "__main__":
    main()
