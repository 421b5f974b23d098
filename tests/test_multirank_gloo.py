"""World-size-2 (and 3) run of the framebuffer gather over torch.distributed's gloo backend on CPU: every
rank fills its compact film with the global pixel index of each owned pixel, the films are gathered to
rank 0 and assembled; the result must be the row-major index image."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rt_b200 import distributed


def _worker(rank, world, port, width, height, tile_rows, out_path, rgb8=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = distributed.owned_rows(height, rank, world, tile_rows)
    pix = (torch.from_numpy(rows)[:, None] * width + torch.arange(width)[None, :]).reshape(-1)
    if rgb8:  # the tone-mapped tiles bench.py gathers per displayed frame: 3 bytes per pixel
        film = torch.zeros((len(rows) * width, 3), dtype=torch.uint8)
        film[:, 0] = (pix % 251).to(torch.uint8)
        film[:, 1] = rank
        film[:, 2] = (pix // 251 % 256).to(torch.uint8)
    else:
        film = torch.zeros((len(rows) * width, 4), dtype=torch.float32)
        film[:, 0] = pix.float()
        film[:, 1] = float(rank)
    gathered = distributed.gather_film(film, width, height, tile_rows, dst=0)
    if rank == 0:
        counts = [distributed.owned_pixels(width, height, r, world, tile_rows) for r in range(world)]
        parts = list(torch.split(gathered, counts))
        full = distributed.assemble_host([p.numpy() for p in parts], width, height, tile_rows)
        np.save(out_path, full)
    else:
        assert gathered is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,width,height,tile_rows", [(2, 16, 37, 4), (3, 8, 10, 8), (2, 5, 3, 1)])
def test_gather_assembles_the_frame(tmp_path, world, width, height, tile_rows):
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), width, height, tile_rows, out), nprocs=world, join=True)
    full = np.load(out)
    assert np.array_equal(full[..., 0].reshape(-1), np.arange(width * height, dtype=np.float32))
    owner = (np.arange(height) // tile_rows) % world
    assert np.array_equal(full[..., 1], np.repeat(owner[:, None], width, axis=1).astype(np.float32))


@pytest.mark.parametrize("world,width,height,tile_rows", [(2, 16, 37, 4), (3, 24, 50, 8)])
def test_gather_of_rgb8_tiles(tmp_path, world, width, height, tile_rows):
    """The per-frame exchange of bench.py / the dynamic camera: uint8 RGB tiles, ranks owning different pixel counts."""
    out = str(tmp_path / "full8.npy")
    mp.spawn(_worker, args=(world, _free_port(), width, height, tile_rows, out, True), nprocs=world, join=True)
    full = np.load(out)
    assert full.dtype == np.uint8 and full.shape == (height, width, 3)
    pix = np.arange(width * height).reshape(height, width)
    assert np.array_equal(full[..., 0], (pix % 251).astype(np.uint8))
    assert np.array_equal(full[..., 2], (pix // 251 % 256).astype(np.uint8))
    owner = (np.arange(height) // tile_rows) % world
    assert np.array_equal(full[..., 1], np.repeat(owner[:, None], width, axis=1).astype(np.uint8))
