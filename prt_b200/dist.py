"""Multi-GPU driver: one process per GPU (torch.distributed), BVH replicated, vertex ranges sharded, coefficient rows
all-gathered (NCCL over NVLink on GPUs; gloo in the CPU test-suite).

The reference has no multi-process path at all; its per-vertex loop
(``std::for_each(par, verts...)``, reference src/raytracing/raytracing.cpp:328) is embarrassingly parallel over vertices,
which is what is sharded here.  There is no data-path collective other than the final all-gather of the
``[n_verts, order^2]`` coefficient rows.
"""
from __future__ import annotations

import numpy as np

CHUNK = 2048


def shard_indices(n_verts: int, world: int, rank: int, chunk: int = CHUNK):
    """Interleaved fixed-size chunks of a (Morton-ordered) vertex list, padded so every rank owns the same count.

    Occlusion cost varies over the surface, so contiguous ranges would be unbalanced; round-robin chunks of a
    space-filling-curve order give every rank a statistically identical slice.  Returns ``(indices, valid, padded_total)``:
    ``indices`` (clamped to ``n_verts - 1`` for padding slots) and ``valid`` have the per-rank length ``padded_total / world``.
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    n_chunks = (n_verts + chunk - 1) // chunk
    n_chunks_pad = ((n_chunks + world - 1) // world) * world
    idx = np.arange(n_chunks_pad * chunk, dtype=np.int64).reshape(n_chunks_pad, chunk)
    mine = idx[rank::world].reshape(-1)
    return np.minimum(mine, max(n_verts - 1, 0)), mine < n_verts, n_chunks_pad * chunk


def unshard_rows(gathered: np.ndarray, n_verts: int, world: int, chunk: int = CHUNK) -> np.ndarray:
    """Inverse of the sharding for an all-gathered ``[world, per_rank, n2]`` array -> ``[n_verts, n2]`` in list order."""
    world_, per_rank, n2 = gathered.shape
    assert world_ == world
    out = np.empty((n_verts, n2), gathered.dtype)
    for r in range(world):
        idx, valid, _ = shard_indices(n_verts, world, r, chunk)
        out[idx[valid]] = gathered[r][valid]
    return out


def sharded_bake(bake_fn, pos: np.ndarray, nrm: np.ndarray, n2: int, group=None, chunk: int = CHUNK) -> np.ndarray:
    """Bakes ``pos/nrm`` (already in the order to be sharded) across the ranks of ``group`` and returns all rows on
    every rank.  ``bake_fn(pos_shard, nrm_shard, vertex_ids) -> [n, n2] float32`` is the per-rank compute (the CUDA
    bake on a GPU rank).  Host-side reference implementation of the collective plumbing; bench.py keeps the rows on
    the device and calls ``all_gather_into_tensor`` directly."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    idx, valid, _ = shard_indices(len(pos), world, rank, chunk)
    rows = np.ascontiguousarray(bake_fn(pos[idx], nrm[idx], idx), dtype=np.float32)
    assert rows.shape == (len(idx), n2)
    if world == 1:
        return unshard_rows(rows[None], len(pos), 1, chunk)
    mine = torch.from_numpy(rows)
    allrows = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype)
    dist.all_gather_into_tensor(allrows.view(-1), mine.reshape(-1), group=group)
    return unshard_rows(allrows.numpy(), len(pos), world, chunk)
