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


def _worker(rank, world, port, width, height, tile_rows, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = distributed.owned_rows(height, rank, world, tile_rows)
    film = torch.zeros((len(rows) * width, 4), dtype=torch.float32)
    pix = (torch.from_numpy(rows)[:, None] * width + torch.arange(width)[None, :]).reshape(-1).float()
    film[:, 0] = pix
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
