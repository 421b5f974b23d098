"""Development aid (GPU box): frame-time distribution of one scene, frame by frame, and - for the slowest and fastest
frames - the longest single traversal the instrumented kernels see (RT_DEBUG_STATS line on stderr).

  python tools/frame_spread.py final 20 1000 1920 8 [frames]
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
os.environ["RT_DEBUG_STATS"] = "1"
import torch  # noqa: E402
from rt_b200 import engine, host  # noqa: E402

name, p0, p1, width, depth = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
frames = int(sys.argv[6]) if len(sys.argv) > 6 else 40
ctx = engine.Context(0)
hs = host.HostScene.builtin(name, 1234, p0, p1)
scene = engine.Scene(ctx, hs.desc)
cam = engine.camera_from_config(hs.camera_config(width, 1, depth))
film = engine.Film(ctx, cam.image_width, cam.image_height)
stream = torch.cuda.ExternalStream(ctx.stream)
for f in range(3):
    engine.render_accumulate(scene, cam, film, 0, 0, 1, depth, 1000 + f)
ctx.synchronize()
pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
for f, (e0, e1) in enumerate(pairs):
    e0.record(stream)
    engine.render_accumulate(scene, cam, film, 0, 0, 1, depth, 1000 + f)
    e1.record(stream)
ctx.synchronize()
times = [e0.elapsed_time(e1) for e0, e1 in pairs]
print("ms per frame:", " ".join("%.2f" % t for t in times))
order = sorted(range(frames), key=lambda f: times[f])
ctx.set_stats(True)
for f in (order[0], order[-1], order[-2]):
    ctx.reset_counters()
    engine.render_accumulate(scene, cam, film, 0, 0, 1, depth, 1000 + f)
    ctx.synchronize()
    sys.stderr.write("frame %d (%.2f ms): " % (f, times[f]))
    sys.stderr.flush()
    c = ctx.counters()
    print("frame", f, "ms", round(times[f], 3), "segments", c.segments, "nodes/seg", round(c.nodes_visited / max(1, c.segments), 2))
