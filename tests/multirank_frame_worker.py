"""Worker of tests/test_gpu_multirank.py (one process per rank, launched by torch.distributed.run).

Every rank renders its own tiles of a few progressive frames and stores them into rank 0's displayed frame through
CUDA IPC (rt_frame_export / rt_frame_open / rt_film_present); rank 0 waits for the arrival flags, downloads the frame
and compares it byte for byte with the frame a single rank renders.  The rendezvous uses gloo, so the ranks may share
one GPU (rank % device_count): the product path under test needs no collective.
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "real-time-ray-tracing-engine_b200"))
from rt_b200 import engine, host  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    device = rank % torch.cuda.device_count()
    torch.cuda.set_device(device)
    ctx = engine.Context(device)
    hs = host.HostScene.builtin("spheres", 1234, 11)
    scene = engine.Scene(ctx, hs.desc)
    width, depth, tile_rows, n_frames = int(os.environ.get("RT_TEST_WIDTH", "320")), 8, 8, 5
    cam = engine.camera_from_config(hs.camera_config(width, 1, depth))
    W, H = cam.image_width, cam.image_height
    film = engine.Film(ctx, W, H, rank, world, tile_rows)
    # two displayed frames in rank 0's memory, mapped by the other ranks
    if rank == 0:
        frames = [engine.Frame(ctx, W, H, world) for _ in range(2)]
        box = [[f.export() for f in frames]]
    else:
        frames, box = [], [None]
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        frames = [engine.Frame(ctx, W, H, world, ipc_handle=h) for h in box[0]]
    dist.barrier()
    host_buf = [np.zeros((W * H, 3), dtype=np.uint8) for _ in range(2)]
    shas = []
    for f in range(n_frames):
        k = f & 1
        engine.render_accumulate(scene, cam, film, 0, 0, 1, depth, 100 + f)
        frames[k].present(film, 1.0 / film.samples)
        if rank == 0:
            frames[k].wait()
            frames[k].download(host_buf[k].ctypes.data)
            frames[k].download_wait()
            shas.append(hashlib.sha256(host_buf[k].tobytes()).hexdigest())
    ctx.synchronize()
    ok = True
    if rank == 0:
        # the same frames from one rank
        solo = engine.Film(ctx, W, H)
        frame = engine.Frame(ctx, W, H, 1)
        buf = np.zeros((W * H, 3), dtype=np.uint8)
        for f in range(n_frames):
            engine.render_accumulate(scene, cam, solo, 0, 0, 1, depth, 100 + f)
            frame.present(solo, 1.0 / solo.samples)
            frame.wait()
            frame.download(buf.ctypes.data)
            frame.download_wait()
            want = hashlib.sha256(buf.tobytes()).hexdigest()
            assert want == hashlib.sha256(solo.resolve_rgb8(1.0 / solo.samples).tobytes()).hexdigest()
            if want != shas[f]:
                ok = False
                print(f"frame {f}: {world}-rank frame differs from the single-rank frame", flush=True)
        assert frames[0].error() == 0 and frames[1].error() == 0 and frame.error() == 0
        assert buf.max() > 0
        frame.close()
        solo.close()
    dist.barrier()
    if rank != 0:
        for fr in frames:
            fr.close()
    dist.barrier()
    if rank == 0:
        for fr in frames:
            fr.close()
        print("MULTIRANK_FRAMES_OK" if ok else "MULTIRANK_FRAMES_DIFFER", world, shas[-1], flush=True)
    film.close()
    scene.close()
    hs.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
