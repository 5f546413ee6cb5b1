"""Host-side description of the multi-GPU frame layout (same arithmetic as dust_shard_layout /
resolve_kernel in csrc): the frame is cut into 32x32 tiles, tile t (row-major) belongs to shard
t % N as that shard's local tile t // N, and each shard owns a compact buffer of
tiles_per_shard * 1024 pixels.  numpy only; used to assemble host-side shard buffers and by the
CPU (gloo) tests of the N > 1 path.
"""
from __future__ import annotations

import numpy as np

TILE = 32


def layout(width: int, height: int, shard_count: int):
    tiles_x, tiles_y = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
    per = (tiles_x * tiles_y + shard_count - 1) // shard_count
    return tiles_x, tiles_y, per, per * TILE * TILE


def pixel_to_shard(width: int, height: int, shard_count: int):
    """(shard[H,W], index[H,W]): where every pixel of the row-major frame lives."""
    tiles_x, _, _, _ = layout(width, height, shard_count)
    y, x = np.mgrid[0:height, 0:width]
    t = (y // TILE) * tiles_x + (x // TILE)
    shard, lt = t % shard_count, t // shard_count
    return shard.astype(np.int64), (lt * TILE * TILE + (y % TILE) * TILE + (x % TILE)).astype(np.int64)


def resolve(gathered: np.ndarray, width: int, height: int) -> np.ndarray:
    """gathered[shard_count, pixels_per_shard, C] -> frame[H, W, C] (what resolve_kernel does on the GPU)."""
    shard, idx = pixel_to_shard(width, height, gathered.shape[0])
    return gathered[shard, idx]


def scatter(frame: np.ndarray, shard_count: int) -> np.ndarray:
    """frame[H, W, C] -> gathered[shard_count, pixels_per_shard, C] (inverse of resolve; padding = 0)."""
    h, w = frame.shape[:2]
    _, _, _, px = layout(w, h, shard_count)
    out = np.zeros((shard_count, px) + frame.shape[2:], frame.dtype)
    shard, idx = pixel_to_shard(w, h, shard_count)
    out[shard, idx] = frame
    return out
