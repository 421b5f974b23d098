#!/usr/bin/env python
"""Benchmark of the path-tracing hot path (BASELINE.json: Mpath-samples/s, ms/frame at 1 spp, RMSE vs ref).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the reference's random-spheres scene (485 objects, seed 1234),
1920x1080, depth 8, progressive dynamic-mode frames.  One step = one frame: one new stratum for every
pixel.  With N GPUs the image is cut into 8-scanline tiles interleaved over the ranks and per-GPU work is
kept constant (weak scaling): a step renders N strata of the whole frame, each rank tracing N strata of
its own 1/N of the tiles; every step ends with the displayed frame: each rank tone-maps its tiles and stores
them straight into rank 0's row-major RGB8 frame over NVLink (rt_film_present, CUDA IPC peer stores + one
arrival flag per rank), rank 0 waits for the flags on the device.

Printed JSON (one line, rank 0):
  value        whole-job Mpath-samples/s with the scene resident in HBM, device-timed (CUDA events, max over ranks)
  e2e          the same metric through the reference-facing call sequence with HOST buffers: camera parameters
               in (kernel arguments), render, present, RGB8 frame copied back to pinned host memory
  roofline     the extend kernel (BVH traversal + primitive tests): algorithmic flops / its measured duration
               against the FP32 issue peak (the kernel is issue / divergence bound, not HBM bound), with the
               issue-slot, HBM, L1 and L2 views beside it
  rmse         BASELINE config 1 (400x225, 100 spp, depth 50) rendered on the GPU against the reference's own CPU
               render of it (oracle/_ref): RMSE, the reference-vs-reference RMSE floor, ratio, luminance
  parity       fast_vs_exact: the FP32 product traversal against the FP64 parity traversal (the reference's
               arithmetic) on EVERY segment of one 1080p frame; how often they name a different primitive
  static_4k    strong scaling: the final scene at 3840x2160, 64 spp, depth 50 - fixed total work over the N GPUs -
               with the SHA-256 of the assembled RGB8 frame (identical for every N)
  frame_sha    SHA-256 of one assembled 1080p frame of the benchmark scene (stratum 0, fixed seed; identical for every N)
  gpu_reference  the reference's own GPU kernels (unmodified, compiled for sm_100) on the same frame and GPU (N=1 only)
  cpu_baseline the reference's own CPU implementation (oracle/_ref, all host threads) on the same frame
`--impl reference` times only that CPU implementation.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
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
LUM = (0.2126, 0.7152, 0.0722)


def workload_config():
    """`config` of the JSON line: the SAME object from both arms (ours and --impl reference) and for every N, so that
    the two lines can be compared key by key; what differs between the arms and with N is in `run`."""
    return {"workload": "spheres scene (BASELINE config 2: 485 objects, seed 1234) 1920x1080, 1 spp per frame, depth 8, "
                        "dynamic-mode (progressive) frames",
            "scene": "spheres", "objects": 485, "width": WIDTH, "height": int(WIDTH / (16.0 / 9.0)), "spp_per_frame": 1,
            "max_depth": DEPTH, "paths_per_frame": WIDTH * int(WIDTH / (16.0 / 9.0)),
            "l2": "GPU arm: 192 MiB buffer written before every timed step (outside the step's CUDA-event pair); "
                  "CPU reference arm: not applicable"}


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


def reference_render(name, p0, p1, width, spp, depth, seed, aspect=None):
    """The reference's own CPU render (oracle/_ref, all host threads) of a built-in scene: (image [H,W,3] f64, s, cores)."""
    import numpy as np

    import oracle_lib as ol
    from rt_b200 import abi

    r = ol.ref()
    h = r.ref_scene_build(name.encode(), SCENE_SEED, p0, p1)
    if aspect:
        r.ref_scene_set_aspect(h, aspect)
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


def image_distance(got, ref_a, ref_b):
    """RMSE of `got` against the reference render `ref_a`, beside the RMSE between two reference renders with
    different seeds (the Monte-Carlo noise floor at this sample count), per pixel and on 8x8-pixel block means
    (noise / 8, so a spatially coherent bias shows); values clipped at 4.0 (fireflies)."""
    import numpy as np

    def clip(x):
        return np.minimum(x, 4.0)

    def blocks(x, b=8):
        h, w = (x.shape[0] // b) * b, (x.shape[1] // b) * b
        return x[:h, :w].reshape(h // b, b, w // b, b, 3).mean(axis=(1, 3))

    def rmse(a, b):
        return float(np.sqrt(np.mean((a - b) ** 2)))

    g, a, b = clip(got.astype(np.float64)), clip(ref_a), clip(ref_b)
    lum = np.array(LUM)
    l_ref = float(0.5 * ((ref_a @ lum).mean() + (ref_b @ lum).mean()))
    l_got = float((got.astype(np.float64) @ lum).mean())
    floor, floor_b = rmse(a, b), rmse(blocks(a), blocks(b))
    return {"rmse": rmse(g, a), "rmse_floor_ref_vs_ref": floor, "ratio": rmse(g, a) / floor,
            "block8_rmse": rmse(blocks(g), blocks(a)), "block8_floor": floor_b, "block8_ratio": rmse(blocks(g), blocks(a)) / floor_b,
            "mean_luminance": l_got, "mean_luminance_reference": l_ref, "luminance_rel_diff": (l_got - l_ref) / l_ref}


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
            "config": workload_config(),
            "run": {"what": "the reference's CPU path (oracle/_ref, unmodified sources) on the host cores, one full frame per step"},
            "cpu_baseline": dict(desc, value=value, unit="Mpath-samples/s"),
            "e2e": {"value": value, "unit": "Mpath-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": elapsed}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from rt_b200 import abi, engine, host

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x):
        t = torch.tensor([float(x)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_sum(values):
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def shared_frames(width, height, count):
        """`count` displayed frames in rank 0's memory; every other rank maps them through CUDA IPC."""
        if rank == 0:
            made = [engine.Frame(ctx, width, height, world) for _ in range(count)]
            handles = [f.export() for f in made] if world > 1 else []
        else:
            made, handles = [], None
        if world > 1:
            box = [handles]
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                made = [engine.Frame(ctx, width, height, world, ipc_handle=h) for h in box[0]]
            dist.barrier()
        return made

    def close_frames(made):
        barrier()
        if rank != 0:  # mappings go before the allocation they map
            for f in made:
                f.close()
        barrier()
        if rank == 0:
            for f in made:
                f.close()

    film = engine.Film(ctx, W, H, rank, world, TILE_ROWS)
    frames = shared_frames(W, H, 2)  # two displayed frames: the copy of one overlaps the render of the next
    rgb8_host = [torch.zeros((npix, 3), dtype=torch.uint8).pin_memory() for _ in range(2)] if rank == 0 else None
    l2_flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def render_step(step):
        """One progressive step on the context stream: `strata_per_step` strata for this rank's tiles."""
        if strata_per_step == 1:
            engine.render_accumulate(scene, cam, film, 0, 0, 1, DEPTH, 1000 + step)
        else:  # the step's strata in ONE wavefront pass (same launch count as a single-GPU frame)
            engine.render_strata(scene, cam, film, 0, strata_per_step, sqrt_spp, DEPTH, 1000 + step)

    def present(frame):
        """The displayed frame: to_byte of this rank's tiles (DynamicCamera::update_texture) stored into their rows of
        rank 0's frame - local stores on rank 0, NVLink stores elsewhere - and the rank's arrival flag; rank 0's
        stream then waits for every rank's flag."""
        frame.present(film, 1.0 / max(1, film.samples))
        if rank == 0:
            frame.wait()

    def flush_only(s):
        with torch.cuda.stream(stream):
            l2_flush.fill_(s & 0xFF)  # evict the scene and queues from L2 between timed steps

    def present_on_device(frame):
        """A frame that stays on the device (a window that displays from GPU memory): present, then rank 0 waits for
        all ranks and marks the frame consumed in one launch."""
        frame.present(film, 1.0 / max(1, film.samples))
        if rank == 0:
            frame.wait_release()

    def device_step(s):
        flush_only(s)
        render_step(s)
        present_on_device(frames[s & 1])

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
            present_on_device(frames[s & 1])  # two displayed frames alternate (double buffering): a rank can present
            e1.record(stream)                 # frame s+1 while the owner is still consuming frame s
        barrier()
        return all_max(sum(e0.elapsed_time(e1) for e0, e1 in pairs) / n_steps)

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
        present(frames[0])
        if rank == 0:
            frames[0].download(rgb8_host[0].data_ptr())
            frames[0].download_wait()
        else:
            ctx.synchronize()

    def wall_ms_per_step(step_fn, finish=None):
        for s in range(args.warmup):
            step_fn(s)
        if finish:
            finish()
        ctx.synchronize()
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            step_fn(s)
        if finish:
            finish()
        ctx.synchronize()
        barrier()
        return all_max((time.perf_counter() - t0) / args.steps * 1e3)

    e2e_sync_ms = wall_ms_per_step(e2e_step)  # frame latency: host blocks until the frame is in host memory

    # Pipelined presentation (what an interactive viewer does): the RGB8 frame of step s goes to pinned host memory
    # on the frame's copy stream while step s+1 renders; two frames + two host buffers, every frame reaches the host.
    def e2e_pipelined_step(s):
        k = s & 1
        render_step(s)
        if rank == 0 and s >= 2:
            frames[k].download_wait()  # host buffer k (frame s-2) is complete before it is reused
        present(frames[k])
        if rank == 0:
            frames[k].download(rgb8_host[k].data_ptr())

    def drain():
        if rank == 0:
            frames[0].download_wait()
            frames[1].download_wait()

    e2e_ms = wall_ms_per_step(e2e_pipelined_step, drain)
    e2e_value = paths_per_step / e2e_ms / 1e3
    frame_error = (frames[0].error() | frames[1].error()) if rank == 0 else 0

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

    # ---- traversal statistics measured on the GPU (instrumented kernels, outside any timed region) ----
    ctx.set_stats(True)
    ctx.reset_counters()
    render_step(0)
    stat = ctx.counters()
    queues = ctx.queue_lengths(DEPTH + 1)
    ctx.set_stats(False)
    stat_sum = all_sum([stat.nodes_visited, stat.prim_tests, stat.segments])

    # ---- parity of the FP32 product traversal: every segment of one frame against the FP64 parity traversal ----
    film.clear()
    ctx.set_audit(True)
    engine.render_accumulate(scene, cam, film, 0, 0, 1, DEPTH, 1000)
    a = ctx.audit()
    ctx.set_audit(False)
    audit_sum = all_sum([a.segments, a.prim_mismatch, a.primary_segments, a.primary_mismatch, a.hit_miss_flips])
    audit_max_t = all_max(a.max_rel_t_error)

    # ---- SHA of one assembled frame (stratum 0, fixed seed): the same bytes for every GPU count ----
    film.clear()
    engine.render_accumulate(scene, cam, film, 0, 0, 1, DEPTH, 4242)
    present(frames[1])
    frame_sha = None
    if rank == 0:
        frames[1].download(rgb8_host[1].data_ptr())
        frames[1].download_wait()
        frame_sha = hashlib.sha256(rgb8_host[1].numpy().tobytes()).hexdigest()
    barrier()

    # ---- strong scaling: the final scene at 4K, fixed total work over the N GPUs ----
    static_4k = static_4k_leg(ctx, stream, world, rank, shared_frames, close_frames, barrier, all_max, all_sum)

    # ---- BASELINE config 1 against the reference's own render (rank 0) ----
    rmse = rmse_leg(ctx) if rank == 0 else None

    # ---- the reference's own GPU kernels on this GPU, same frame (informational comparator, own process) ----
    gpu_reference = gpu_reference_leg(ms_step) if rank == 0 and world == 1 else None

    peaks = {}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass

    if rank == 0:
        cpu_value, cpu_desc = cpu_reference_frame(3)
        clocks = sampler.summary()
        line = {
            "metric": METRIC, "value": value, "unit": "Mpath-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(),
            "run": {
                "what": f"{strata_per_step} stratum/strata of the whole frame per step (= 1 per GPU), "
                        f"{TILE_ROWS}-scanline tiles interleaved over {world} GPU(s), scene replicated; every step ends with "
                        f"the displayed frame: each rank's to_byte tiles stored into rank 0's row-major RGB8 frame "
                        f"(NVLink peer stores through CUDA IPC + arrival flags; no collective); two frames alternate",
                "paths_per_step": paths_per_step, "segments_per_path": counters.segments / max(1, counters.paths),
                "bvh": {"primitives": info.n_prims, "nodes4": info.n_nodes, "device_build_ms": info.build_ms}},
            "frame_ms": ms_step / strata_per_step,
            "e2e": {"value": e2e_value, "unit": "Mpath-samples/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": C.sizeof(abi.rt_camera), "d2h_bytes_per_step": npix * 3,
                    "path": "rt_render_accumulate / rt_render_strata -> rt_film_present -> rt_frame_wait -> rt_frame_download "
                            "into pinned host RGB8 (the copy of frame s overlaps the render of frame s+1; every frame reaches the host)",
                    "frame_latency_ms": e2e_sync_ms,
                    "frame_latency_note": "same call sequence with the host blocking until each frame is in host memory",
                    "frame_wait_timeouts": int(frame_error)},
            "gpu_launches": int(counters.kernel_launches),
            "roofline": roofline_block(peaks, seg_per_launch, extend_ms_per_launch, stage_ms, roof_steps,
                                       strata_per_step, roof_counters, clocks),
            "traversal": {"note": "counted on the GPU by the instrumented kernels (rt_context_set_stats), one step",
                          "node_visits_per_segment": stat_sum[0] / max(1.0, stat_sum[2]),
                          "box_tests_per_segment": 4.0 * stat_sum[0] / max(1.0, stat_sum[2]),
                          "primitive_tests_per_segment": stat_sum[1] / max(1.0, stat_sum[2]),
                          "queue_lengths_rank0": queues},
            "parity": {"fast_vs_exact": {
                "what": "every ray segment of one 1920x1080 depth-8 frame: the FP32 product traversal (k_extend) against the "
                        "FP64 parity traversal (the reference's arithmetic, bit-exact against the reference in tests/) of the same ray",
                "segments": int(audit_sum[0]), "prim_mismatch": int(audit_sum[1]),
                "mismatch_rate": audit_sum[1] / max(1.0, audit_sum[0]),
                "primary_segments": int(audit_sum[2]), "primary_mismatch": int(audit_sum[3]),
                "primary_mismatch_rate": audit_sum[3] / max(1.0, audit_sum[2]),
                "hit_miss_flips": int(audit_sum[4]), "max_rel_t_error_same_prim": audit_max_t}},
            "rmse": rmse,
            "gpu_reference": gpu_reference,
            "frame_sha": frame_sha,
            "static_4k": static_4k,
            "cpu_baseline": dict(cpu_desc, value=cpu_value, unit="Mpath-samples/s"),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    # tear down in dependency order: torch buffers that were used on the context's stream go first (the
    # pinned-memory allocator records an event on that stream when a block is released)
    close_frames(frames)
    film.close()
    del rgb8_host, l2_flush
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    scene.close()
    hs.close()
    del stream
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def roofline_block(peaks, seg_per_launch, launch_ms, stage_ms, roof_steps, strata_per_step, roof_counters, clocks):
    """The dominant kernel (k_extend) against the bound that limits it.  BVH traversal of an L1/L2-resident scene is
    bound by instruction issue at the lane utilisation incoherent rays allow (ncu: 2.4-2.8 of 4 issue slots, 14-27
    of 32 lanes, DRAM at 4-8 % of peak), so the headline fraction is the algorithmic FP32 rate over the FP32 issue
    peak; the issue-slot, HBM, L1 and L2 views are sub-blocks.  ncu-derived constants come from the tracked
    profiles/extend_ncu.json (tools/summarize_ncu.py); everything else is measured live."""
    secs = launch_ms * 1e-3
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    flops = seg_per_launch * FLOPS_PER_SEGMENT / secs / 1e12 if secs > 0 else 0.0
    gbs = seg_per_launch * BYTES_PER_SEGMENT / secs / 1e9 if secs > 0 else 0.0
    ncu, cache = {}, {}
    for name, target in (("extend_ncu.json", ncu), ("cache_peaks.json", cache)):
        try:
            with open(os.path.join(REPO, "profiles", name)) as f:
                target.update(json.load(f))
        except Exception:
            pass
    sm_mhz = float(clocks.get("sm_mhz") or sm_max)
    inst_per_seg = ncu.get("warp_inst_per_segment")
    issue = None
    if inst_per_seg and secs > 0:
        per_clk = inst_per_seg * seg_per_launch / secs / (sm_mhz * 1e6) / 148
        issue = {"warp_inst_per_segment": inst_per_seg, "issue_slots_per_clk_per_sm": per_clk, "peak": 4.0, "frac": per_clk / 4.0,
                 "lanes_per_inst": ncu.get("lanes_per_inst"),
                 "useful_lane_issue_frac": per_clk / 4.0 * (ncu.get("lanes_per_inst") or 0) / 32.0,
                 "source": "warp instructions and lanes per instruction: ncu capture summarised in profiles/extend_ncu.json; "
                           "duration, segments and SM clock: this run"}
    dram_per_seg = ncu.get("dram_bytes_per_segment")
    return {
        "bound": "issue", "kernel": "k_extend", "achieved": flops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": flops / fp32_peak,
        "traffic": ncu.get("dram_bytes_per_launch"),
        "peak_source": (f"FP32 issue peak 148 SMs x 128 lanes x 2 flop x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json; "
                        "that file holds no FP32 figure)") if peaks else "FP32 issue peak at the fallback 1965 MHz",
        "algorithmic_flops_per_segment": FLOPS_PER_SEGMENT, "algorithmic_bytes_per_segment": BYTES_PER_SEGMENT,
        "segments_per_launch": seg_per_launch, "launch_ms": launch_ms,
        "issue": issue,
        "hbm": {"note": "the scene (0.06 MB) is L1/L2 resident: the algorithmic bytes are node / primitive fetches served on chip, so "
                        "their rate may exceed the HBM peak; what HBM actually moves is dram_bytes_per_segment (queue entries)",
                "algorithmic_gbs": gbs, "peak": hbm_peak, "algorithmic_frac_of_hbm": gbs / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "dram_bytes_per_segment": dram_per_seg,
                "dram_gbs": dram_per_seg * seg_per_launch / secs / 1e9 if dram_per_seg and secs > 0 else None,
                "dram_frac": dram_per_seg * seg_per_launch / secs / 1e9 / hbm_peak if dram_per_seg and secs > 0 else None},
        "on_chip": ({"note": "the same algorithmic bytes against the L1- and L2-resident read bandwidth measured on this pool "
                             "(profiles/cache_peaks.json, tools/cache_peaks; not in MEASURED_PEAKS.json)",
                     "l1_peak": cache["l1_read_gbs"], "l1_frac": gbs / cache["l1_read_gbs"],
                     "l2_peak": cache["l2_read_gbs"], "l2_frac": gbs / cache["l2_read_gbs"]}
                    if cache.get("l1_read_gbs") and cache.get("l2_read_gbs") else None),
        "stage_ms_per_frame": {k: stage_ms[i] / roof_steps / strata_per_step
                               for i, k in enumerate(["generate", "extend", "shade", "accumulate", "tail"])},
        "tail_segments_per_frame": roof_counters.tail_segments / roof_steps / strata_per_step}


def static_4k_leg(ctx, stream, world, rank, shared_frames, close_frames, barrier, all_max, all_sum):
    """Strong scaling (north_star: "near-linear 8-GPU scaling on 4K high-spp static renders"): the final scene
    (BASELINE config 5's scene) at 3840x2160, depth 50, 64 spp instead of 4096 so that the default run stays short -
    FIXED total work, tiles interleaved over the N GPUs, every rank's tiles tone-mapped into rank 0's 4K frame.
    Device time of render + present (max over ranks), best of 2, and the SHA-256 of the assembled frame."""
    import torch

    from rt_b200 import engine, host

    WIDTH4K, ROOT, DEPTH4K, SEED = 3840, 8, 50, 77
    hs = host.HostScene.builtin("final", SCENE_SEED, 20, 1000)
    scene = engine.Scene(ctx, hs.desc)
    n_prims = scene.info().n_prims
    cam = engine.camera_from_config(hs.camera_config(WIDTH4K, ROOT * ROOT, DEPTH4K))
    W, H = cam.image_width, cam.image_height
    film = engine.Film(ctx, W, H, rank, world, TILE_ROWS)
    made = shared_frames(W, H, 1)
    frame = made[0]
    host_rgb8 = torch.zeros((W * H, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None

    def one(root):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        engine.render_static(scene, cam, film, root, DEPTH4K, SEED)
        frame.present(film, 1.0 / (root * root))
        if rank == 0:
            frame.wait_release()
        e1.record(stream)
        barrier()
        return all_max(e0.elapsed_time(e1))

    one(1)  # warm-up: sizes the queues
    ctx.reset_counters()
    times = [one(ROOT) for _ in range(2)]
    c = ctx.counters()
    seg = all_sum([c.segments, c.paths])
    sha = None
    frame.present(film, 1.0 / (ROOT * ROOT))
    if rank == 0:
        frame.wait()
        frame.download(host_rgb8.data_ptr())
        frame.download_wait()
        sha = hashlib.sha256(host_rgb8.numpy().tobytes()).hexdigest()
    ms = min(times)
    paths = W * H * ROOT * ROOT
    out = {"workload": f"final scene ({n_prims} primitives, two media) {W}x{H}, {ROOT * ROOT} spp, depth {DEPTH4K}, static "
                       f"render, fixed total work: {TILE_ROWS}-scanline tiles interleaved over {world} GPU(s), tiles tone-mapped into "
                       f"rank 0's frame over NVLink", "scaling": "strong",
           "ms": ms, "ms_all_runs": times, "value": paths / ms / 1e3, "unit": "Mpath-samples/s", "paths": paths,
           "segments_per_path": seg[0] / max(1.0, seg[1]), "frame_sha": sha}
    close_frames(made)
    film.close()
    scene.close()
    hs.close()
    del host_rgb8
    return out


def rmse_leg(ctx):
    """BASELINE config 1 (spheres scene 400x225, 100 spp, depth 50): this library's render against the reference's
    own CPU render of the same configuration (oracle/_ref, all host threads), with the reference-vs-reference RMSE
    (two seeds) as the noise floor.  The checker role of oracle/; null where the harness was not built."""
    import oracle_lib as ol
    from rt_b200 import engine, host

    if not ol.have_ref():
        return None
    hs = host.HostScene.builtin(SCENE, SCENE_SEED, 11)
    scene = engine.Scene(ctx, hs.desc)
    cam = engine.camera_from_config(hs.camera_config(400, 100, 50))
    film = engine.Film(ctx, cam.image_width, cam.image_height)
    engine.render_static(scene, cam, film, 10, 50, 7)
    got = film.read_rgb(1.0 / 100).reshape(cam.image_height, cam.image_width, 3)
    ref_a, secs, cores = reference_render(SCENE, 11, 0, 400, 100, 50, 11)
    ref_b, _, _ = reference_render(SCENE, 11, 0, 400, 100, 50, 100011)  # far from seed 11: thread t is seeded seed + 1 + t
    out = {"config": "BASELINE config 1: spheres scene 400x225, 100 spp, depth 50; reference = oracle/_ref (unmodified reference "
                     "sources, all host threads), seeds 11 and 100011",
           "reference_cpu_s": secs, "reference_cores": cores}
    out.update(image_distance(got, ref_a, ref_b))
    film.close()
    scene.close()
    hs.close()
    return out


def gpu_reference_leg(our_frame_ms):
    """The reference's OWN GPU path (its .cu kernels compiled unmodified for sm_100, oracle/_ref_gpu) on the benchmark
    frame, on this GPU, in a process of its own (tools/ref_gpu_frame.py): the list world (its -b configuration faults
    on this scene, profiles/r02_reference_gpu.md, and is not launched here), one whole-frame launch of its kernel.
    Informational: the headline ratio stays the driver's, against the CPU arm."""
    import subprocess

    try:
        # whole-frame launch only: the reference's per-tile launch loop takes 44 s per 1080p frame (profiles/r02_reference_gpu.md)
        r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ref_gpu_frame.py"), SCENE, "11", str(WIDTH), str(DEPTH), "1",
                            "--whole-frame-only", "--list-world-only"],  # a bench run launches no kernel known to fault
                           capture_output=True, text=True, timeout=900)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        out = json.loads(lines[-1]) if lines else {"error": (r.stderr or "no output")[-300:]}
    except Exception as e:  # a comparator that cannot run is reported, never fatal
        out = {"error": repr(e)}
    out["what"] = ("the reference's own CUDA path (FP64, recursive megakernel, XORWOW state per pixel), unmodified sources compiled "
                   "for sm_100, one 1920x1080 1-spp depth-8 frame of the same scene on this GPU; the harness raises the device "
                   "stack limit to 32 KB (the reference never does)")
    out["ours_frame_ms"] = our_frame_ms
    for world in ("bvh_world", "list_world"):
        for key in ("whole_frame_launch", "tile32_loop"):
            leg = out.get(world, {}).get(key) if isinstance(out.get(world), dict) else None
            if isinstance(leg, dict) and leg.get("ms_per_frame"):
                leg["ratio_over_ours"] = leg["ms_per_frame"] / our_frame_ms
    return out


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
