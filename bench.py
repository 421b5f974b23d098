#!/usr/bin/env python
"""Benchmark of the path-tracing hot path (BASELINE.json: Mpath-samples/s, ms/frame at 1 spp).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the reference's random-spheres scene (485 objects, seed 1234),
1920x1080, depth 8, progressive dynamic-mode frames.  One step = one frame: one new stratum for every
pixel.  With N GPUs the image is cut into 8-scanline tiles interleaved over the ranks and per-GPU work is
kept constant (weak scaling): a step renders N strata of the whole frame, each rank tracing N strata of
its own 1/N of the tiles, and the compact films are gathered to rank 0 (NCCL over NVLink) and scattered
into the full frame at the end of every step (RGB8, after the device tone map).

Printed JSON (one line, rank 0):
  value      whole-job Mpath-samples/s with the scene resident in HBM, device-timed (CUDA events, max over ranks)
  e2e        the same metric through the reference-facing call sequence with HOST buffers: camera parameters
             in (kernel arguments), render, device to_byte resolve, RGB8 frame copied back to host memory
  roofline   the extend kernel (BVH traversal + primitive tests): algorithmic bytes / its measured duration
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, all host threads) on the same frame
`--impl reference` times only that CPU implementation.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))

WIDTH, DEPTH, SCENE, SCENE_SEED, TILE_ROWS = 1920, 8, "spheres", 1234, 8
METRIC = "Mpath-samples/s (spheres scene, 1920x1080, 1 spp per frame, depth 8)"
# Algorithmic work per ray segment on the reference's own binary BVH (oracle counters on
# this frame): 26.0 box tests x 32 B + 1.71 sphere tests x 32 B + 128 B of ray/hit/path-state queue traffic.
NODE_TESTS_PER_SEGMENT, SPHERE_TESTS_PER_SEGMENT = 26.0, 1.71
BYTES_PER_SEGMENT = 32.0 * NODE_TESTS_PER_SEGMENT + 32.0 * SPHERE_TESTS_PER_SEGMENT + 128.0
FLOPS_PER_SEGMENT = 20.0 * NODE_TESTS_PER_SEGMENT + 30.0 * SPHERE_TESTS_PER_SEGMENT + 150.0


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region, through NVML (the same counters nvidia-smi's
    clocks.sm / clocks_event_reasons.* columns print, B200_PROFILING.md), sampled every 2 ms."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.max_mhz = None
        self.error = None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag.is_set():
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(reasons_fn(h))))
                self.stop_flag.wait(0.002)
            nv.nvmlShutdown()
        except Exception as e:  # NVML unavailable: report it instead of inventing clocks
            self.error = repr(e)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled: " + str(self.error)]}
        sm = sorted(s[0] for s in self.samples)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        reasons = [n for n, b in bits.items() if any(s[1] & b for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}


def cpu_reference_frame(n_frames, threads=None):
    """The reference's CPU implementation of the path on this workload: Camera::get_ray + Camera::ray_color
    on the reference's own BVH world (oracle/_ref/libref_harness.so, unmodified reference sources), one
    dynamic-mode frame (stratum 0 of 1 spp, depth 8) of the full 1920x1080 image, scanlines handed to all
    host threads.  Falls back to the single-thread C restatement (oracle/liboracle.so) when the harness has not
    been built.  Returns (Mpath/s, description dict)."""
    import oracle_lib as ol

    npix = None
    if ol.have_ref():
        r = ol.ref()
        cores = threads or r.ref_hardware_threads()
        h = r.ref_scene_build(SCENE.encode(), SCENE_SEED, 11, 0)
        height = int(WIDTH / (16.0 / 9.0))
        npix = WIDTH * height
        img = (C.c_double * (npix * 3))()
        seg = C.c_uint64()
        r.ref_render(h, 64, 1, DEPTH, 1, 1, 1, 0, -1, 0, img, C.byref(seg))  # builds the BVH outside the timed region
        secs = 0.0
        for f in range(n_frames):
            secs += r.ref_render(h, WIDTH, 1, DEPTH, 100 + f, 1, cores, 0, height, 0, img, C.byref(seg))
        r.ref_scene_free(h)
        return npix * n_frames / secs / 1e6, {
            "kind": "reference", "cores": cores,
            "sample": f"{n_frames} full frame(s) 1920x1080, 1 spp, depth 8, reference BVH (-b), one scanline per job"}
    from rt_b200 import host

    o = ol.oracle()
    hs = host.HostScene.builtin(SCENE, SCENE_SEED, 11)
    cfg = hs.camera_config(WIDTH, 1, DEPTH)
    height = int(WIDTH / (16.0 / 9.0))
    rows = 270  # a quarter of the frame, single thread
    img = (C.c_double * (WIDTH * rows * 3))()
    osc = o.ora_scene_create(hs.desc)
    secs = o.ora_render(osc, C.byref(cfg), ol.ORA_RNG_MT19937, ol.ORA_SAMPLER_REJECTION, 100, 1, (height - rows) // 2,
                        (height - rows) // 2 + rows, 0, img, None)
    o.ora_scene_destroy(osc)
    return WIDTH * rows / secs / 1e6, {"kind": "port", "cores": 1,
                                       "sample": f"rows {(height - rows) // 2}..{(height - rows) // 2 + rows} of one 1920x1080 frame, 1 spp, depth 8"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, desc = 0.0, {}
    for _ in range(args.warmup):
        cpu_reference_frame(1)
    t0 = time.time()
    value, desc = cpu_reference_frame(args.steps)
    elapsed = time.time() - t0
    height = int(WIDTH / (16.0 / 9.0))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpath-samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": WIDTH * height / value / 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "spheres scene (485 objects, seed 1234) 1920x1080, 1 spp per frame, depth 8; "
                                   "reference CPU path on the host cores"},
            "cpu_baseline": dict(desc, value=value, unit="Mpath-samples/s"),
            "e2e": {"value": value, "unit": "Mpath-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": elapsed}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from rt_b200 import abi, distributed, engine, host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA path is the product, there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = engine.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    hs = host.HostScene.builtin(SCENE, SCENE_SEED, 11)
    scene = engine.Scene(ctx, hs.desc)
    info = scene.info()
    cfg = hs.camera_config(WIDTH, 1, DEPTH)
    cam = engine.camera_from_config(cfg)
    W, H = cam.image_width, cam.image_height
    npix = W * H
    strata_per_step = world  # weak scaling: N strata of the whole frame per step
    sqrt_spp = 1
    while sqrt_spp * sqrt_spp < strata_per_step:  # smallest stratification grid holding N strata
        sqrt_spp += 1

    owned = distributed.owned_pixels(W, H, rank, world, TILE_ROWS)
    accum = torch.zeros((owned, 4), dtype=torch.float32, device="cuda")
    own_rgb8 = torch.zeros((owned, 3), dtype=torch.uint8, device="cuda")  # this rank's tiles, tone-mapped
    rgb8_dev = torch.zeros((npix, 3), dtype=torch.uint8, device="cuda") if rank == 0 else None
    rgb8_host = torch.zeros((npix, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None
    torch.cuda.synchronize()
    film = engine.Film(ctx, W, H, rank, world, TILE_ROWS, external_accum=accum.data_ptr())
    l2_flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def render_step(step):
        """One progressive step on the context stream: `strata_per_step` strata for this rank's tiles."""
        if strata_per_step == 1:
            engine.render_accumulate(scene, cam, film, 0, 0, 1, DEPTH, 1000 + step)
        else:  # the step's strata in ONE wavefront pass (same launch count as a single-GPU frame)
            engine.render_strata(scene, cam, film, 0, strata_per_step, sqrt_spp, DEPTH, 1000 + step)

    def present(step, frame_dev):
        """The displayed frame: every rank tone-maps its own tiles to RGB8 on the device (DynamicCamera::
        update_texture, to_byte); with several GPUs the RGB8 tiles (3 bytes per pixel) are gathered to rank 0 over
        NCCL / NVLink and scattered into the row-major frame."""
        scale = 1.0 / max(1, film.samples)
        if world == 1:
            abi.check(ctx.lib, ctx.lib.rt_film_resolve_rgb8_device(film._h, scale, frame_dev.data_ptr()),
                      "rt_film_resolve_rgb8_device")
            return
        abi.check(ctx.lib, ctx.lib.rt_film_resolve_rgb8_device(film._h, scale, own_rgb8.data_ptr()),
                  "rt_film_resolve_rgb8_device")
        with torch.cuda.stream(stream):
            gathered = distributed.gather_film(own_rgb8, W, H, TILE_ROWS, dst=0)
            if rank == 0:
                abi.check(ctx.lib, ctx.lib.rt_film_scatter_gathered_rgb8(ctx._h, W, H, world, TILE_ROWS,
                                                                         gathered.data_ptr(), frame_dev.data_ptr()),
                          "rt_film_scatter_gathered_rgb8")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_only(s):
        with torch.cuda.stream(stream):
            l2_flush.fill_(s & 0xFF)  # evict the scene and queues from L2 between timed steps

    def device_step(s):
        flush_only(s)
        render_step(s)
        present(s, rgb8_dev)

    def timed(n_steps):
        """Device time of n_steps steps on the context stream (ms per step, max over ranks).  Every step is
        bracketed by its own CUDA-event pair recorded after the L2 flush, so the flush is not in the timed
        region; the whole loop is bracketed by barrier + synchronize."""
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        barrier()
        for s, (e0, e1) in enumerate(pairs):
            flush_only(s)
            e0.record(stream)
            render_step(s)
            present(s, rgb8_dev)
            e1.record(stream)
        barrier()
        ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in pairs) / n_steps], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput ----
    for s in range(args.warmup):
        device_step(s)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.reset_counters()
    ms_step = timed(args.steps)
    counters = ctx.counters()
    paths_per_step = npix * strata_per_step
    value = paths_per_step / ms_step / 1e3

    # ---- end to end through the reference-facing calls, host buffers ----
    # (what DynamicCamera::render_gpu does per frame: launch, synchronise, copy the frame back, tone-map -
    #  here the tone map runs on the device and 3 bytes per pixel cross PCIe instead of 24)
    def e2e_step(s):
        render_step(s)
        present(s, rgb8_dev)
        if rank == 0:
            with torch.cuda.stream(stream):
                rgb8_host.copy_(rgb8_dev, non_blocking=True)
        ctx.synchronize()

    def wall_ms_per_step(step_fn):
        for s in range(args.warmup):
            step_fn(s)
        ctx.synchronize()
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            step_fn(s)
        ctx.synchronize()
        barrier()
        wall = torch.tensor([(time.perf_counter() - t0) / args.steps * 1e3], device="cuda")
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        return float(wall.item())

    e2e_sync_ms = wall_ms_per_step(e2e_step)  # frame latency: host blocks until the frame is in host memory

    # Pipelined presentation (what an interactive viewer does): the RGB8 frame of step s is copied to pinned
    # host memory on a copy stream while step s+1 renders; two device + two host buffers, every frame still
    # reaches the host.
    copy_stream = torch.cuda.Stream()
    if rank == 0:
        rgb8_dev2 = [rgb8_dev, torch.zeros_like(rgb8_dev)]
        rgb8_host2 = [rgb8_host, torch.zeros((npix, 3), dtype=torch.uint8).pin_memory()]
        resolved = [torch.cuda.Event(), torch.cuda.Event()]
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        for ev in copied:
            ev.record(copy_stream)

    def e2e_pipelined_step(s):
        render_step(s)
        k = s & 1
        if rank == 0:
            stream.wait_event(copied[k])  # the buffer's previous frame has left the device
        present(s, rgb8_dev2[k] if rank == 0 else None)
        if rank == 0:
            resolved[k].record(stream)
            copy_stream.wait_event(resolved[k])
            with torch.cuda.stream(copy_stream):
                rgb8_host2[k].copy_(rgb8_dev2[k], non_blocking=True)
            copied[k].record(copy_stream)

    e2e_ms = wall_ms_per_step(e2e_pipelined_step)
    copy_stream.synchronize()
    e2e_value = paths_per_step / e2e_ms / 1e3

    # ---- roofline of the dominant kernel (k_extend), measured live with per-launch events ----
    ctx.set_stage_timing(True)
    ctx.reset_counters()
    roof_steps = max(3, min(args.steps, 10))
    for s in range(roof_steps):
        flush_only(s)
        render_step(s)
    stage_ms, stage_n = ctx.stage_times()
    ctx.set_stage_timing(False)
    roof_counters = ctx.counters()
    seg = roof_counters.segments - roof_counters.tail_segments  # segments of the k_extend launches (tail excluded)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()
    extend_ms_per_launch = stage_ms[1] / max(1, stage_n[1])
    seg_per_launch = seg / max(1, stage_n[1])
    peaks = {}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = seg_per_launch * BYTES_PER_SEGMENT / (extend_ms_per_launch * 1e-3) / 1e9 if extend_ms_per_launch > 0 else 0.0
    fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    flops = seg_per_launch * FLOPS_PER_SEGMENT / (extend_ms_per_launch * 1e-3) / 1e12 if extend_ms_per_launch > 0 else 0.0
    cache = {}
    try:  # L1- / L2-resident read bandwidth measured on this pool (tools/cache_peaks): the levels that serve the BVH
        with open(os.path.join(REPO, "profiles", "cache_peaks.json")) as f:
            cache = json.load(f)
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(REPO, "profiles", "extend_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    line = None
    if rank == 0:
        cpu_value, cpu_desc = cpu_reference_frame(3)
        line = {
            "metric": METRIC, "value": value, "unit": "Mpath-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"spheres scene (485 objects, seed 1234) 1920x1080, depth 8, dynamic-mode frames: "
                            f"{strata_per_step} stratum/strata of the whole frame per step (= 1 per GPU), "
                            f"{TILE_ROWS}-scanline tiles interleaved over {world} GPU(s), scene replicated; every step ends with "
                            f"the displayed frame: device to_byte resolve of each rank's tiles, RGB8 tiles gathered to rank 0 "
                            f"(NCCL) and scattered into the row-major frame",
                "paths_per_step": paths_per_step, "segments_per_path": counters.segments / max(1, counters.paths),
                "bvh": {"primitives": info.n_prims, "nodes4": info.n_nodes, "device_build_ms": info.build_ms},
                "l2": "192 MiB buffer written before every timed step (outside the step's CUDA-event pair)"},
            "frame_ms": ms_step / strata_per_step,
            "e2e": {"value": e2e_value, "unit": "Mpath-samples/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": C.sizeof(abi.rt_camera), "d2h_bytes_per_step": npix * 3,
                    "path": "rt_render_accumulate / rt_render_strata -> rt_film_resolve_rgb8_device -> pinned host RGB8 "
                            "(copy of frame s overlaps the render of frame s+1; every frame reaches the host)",
                    "frame_latency_ms": e2e_sync_ms,
                    "frame_latency_note": "same call sequence with the host blocking until each frame is in host memory"},
            "gpu_launches": int(counters.kernel_launches),
            "roofline": {"bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_segment": BYTES_PER_SEGMENT, "segments_per_launch": seg_per_launch,
                         "launch_ms": extend_ms_per_launch,
                         "note": "working set is L1/L2 resident (0.06 MB scene); the algorithmic bytes are node/primitive "
                                 "fetches + queue traffic, so this is an L1/L2 figure set against the HBM copy peak",
                         "fp32": {"achieved_tflops": flops, "peak_tflops": fp32_peak, "frac": flops / fp32_peak},
                         "on_chip": ({"note": "the same algorithmic bytes against the measured L1- and L2-resident read "
                                              "bandwidth (profiles/cache_peaks.json, tools/cache_peaks): the BVH is served by L1",
                                      "l1_peak": cache["l1_read_gbs"], "l1_frac": achieved / cache["l1_read_gbs"],
                                      "l2_peak": cache["l2_read_gbs"], "l2_frac": achieved / cache["l2_read_gbs"]}
                                     if cache.get("l1_read_gbs") and cache.get("l2_read_gbs") else None),
                         "hbm_only": {"note": "bytes that must come from HBM per segment: the ray queue entry read (32 B) "
                                              "+ skip primitive read and hit written (16 B); the BVH stays in L1/L2",
                                      "bytes_per_segment": 48.0,
                                      "achieved": seg_per_launch * 48.0 / (extend_ms_per_launch * 1e-3) / 1e9
                                      if extend_ms_per_launch > 0 else 0.0,
                                      "frac": (seg_per_launch * 48.0 / (extend_ms_per_launch * 1e-3) / 1e9 / hbm_peak)
                                      if extend_ms_per_launch > 0 else 0.0},
                         "stage_ms_per_frame": {k: stage_ms[i] / roof_steps / strata_per_step
                                                for i, k in enumerate(["generate", "extend", "shade", "accumulate", "tail"])},
                         "tail_segments_per_frame": roof_counters.tail_segments / roof_steps / strata_per_step},
            "cpu_baseline": dict(cpu_desc, value=cpu_value, unit="Mpath-samples/s"),
            "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    # tear down in dependency order: torch buffers that were used on the context's stream go first (the
    # pinned-memory allocator records an event on that stream when a block is released)
    barrier()
    film.close()
    if rank == 0:
        del rgb8_dev2, rgb8_host2, resolved, copied
    del accum, own_rgb8, rgb8_dev, rgb8_host, l2_flush, copy_stream
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    scene.close()
    hs.close()
    del stream
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
