"""Development aid (GPU box): how much of a C2 frame is kernel-end idling?  Renders frames of the same scene from TWO
contexts (two streams) at once and compares with one context doing the same number of frames back to back: the gap
is what a frame split into two concurrently submitted halves could recover at best."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))
import torch  # noqa: E402
from rt_b200 import engine, host  # noqa: E402

frames = 40
hs = host.HostScene.builtin("spheres", 1234, 11)
ctxs = [engine.Context(0), engine.Context(0)]
scenes = [engine.Scene(c, hs.desc) for c in ctxs]
cam = engine.camera_from_config(hs.camera_config(1920, 1, 8))
films = [engine.Film(c, cam.image_width, cam.image_height) for c in ctxs]
streams = [torch.cuda.ExternalStream(c.stream) for c in ctxs]


def run(n_ctx):
    for c in ctxs:
        c.synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event() for _ in range(n_ctx)]
    e0.record(streams[0])
    if n_ctx == 2:
        streams[1].wait_event(e0)
    for f in range(frames):
        for k in range(n_ctx):
            engine.render_accumulate(scenes[k], cam, films[k], 0, 0, 1, 8, 1000 + f)
    for k in range(1, n_ctx):
        ends[k].record(streams[k])
        streams[0].wait_event(ends[k])
    e1.record(streams[0])
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (frames * n_ctx)


for _ in range(2):
    run(1), run(2)
one = min(run(1) for _ in range(3))
two = min(run(2) for _ in range(3))
print("ms per frame: one context %.4f, two contexts interleaved %.4f (%.1f %% less)" % (one, two, 100 * (1 - two / one)))
