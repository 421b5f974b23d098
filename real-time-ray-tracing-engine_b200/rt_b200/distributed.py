"""Image-space partition over GPUs: scanline tiles interleaved over ranks, one framebuffer gather.

Per-pixel samples need no communication while rendering (the Philox stream is keyed by the global
pixel index), so every rank renders into a compact film holding only its own tiles and the frame is
assembled once: all ranks' compact films are gathered to rank 0 (torch.distributed over NCCL / NVLink,
or gloo in the CPU tests) and scattered into the row-major image (k_scatter_gathered on the GPU).
"""
import numpy as np


def owned_rows(height, rank, n_ranks, tile_rows):
    """Global scanline indices rank `rank` owns, in its storage order (mirrors csrc/rt_device.h owned_rows)."""
    rows = []
    n_tiles = (height + tile_rows - 1) // tile_rows
    for t in range(rank, n_tiles, n_ranks):
        rows.extend(range(t * tile_rows, min((t + 1) * tile_rows, height)))
    return np.asarray(rows, dtype=np.int64)


def owned_pixels(width, height, rank, n_ranks, tile_rows):
    return int(len(owned_rows(height, rank, n_ranks, tile_rows))) * width


def assemble_host(parts, width, height, tile_rows):
    """numpy reference of the scatter: parts[r] = rank r's compact film, shape (owned_pixels, C)."""
    n_ranks = len(parts)
    channels = parts[0].shape[-1]
    full = np.zeros((height, width, channels), dtype=parts[0].dtype)
    for r, part in enumerate(parts):
        rows = owned_rows(height, r, n_ranks, tile_rows)
        full[rows] = np.asarray(part).reshape(len(rows), width, channels)
    return full


def gather_film(film_tensor, width, height, tile_rows, dst=0, group=None):
    """Gather every rank's compact film (torch tensor [owned_pixels, 4]) to rank `dst`.

    Returns on `dst` a tensor [sum(owned_pixels), 4] in rank-major order (the layout
    rt_film_scatter_gathered consumes) and None elsewhere.  Ranks may own different pixel counts, so the
    exchange pads every part to the largest one."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [owned_pixels(width, height, r, world, tile_rows) for r in range(world)]
    biggest = max(counts)
    padded = film_tensor
    if film_tensor.shape[0] != biggest:
        padded = torch.zeros((biggest, film_tensor.shape[1]), dtype=film_tensor.dtype, device=film_tensor.device)
        padded[: film_tensor.shape[0]] = film_tensor
    if rank == dst:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, parts, dst=dst, group=group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    dist.gather(padded, None, dst=dst, group=group)
    return None
