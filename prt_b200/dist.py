"""Multi-GPU driver: one process per GPU (torch.distributed), BVH replicated, vertex ranges sharded, coefficient rows
all-gathered (NCCL over NVLink on GPUs; gloo in the CPU test-suite).

The reference has no multi-process path at all; its per-vertex loop
(``std::for_each(par, verts...)``, reference src/raytracing/raytracing.cpp:328) is embarrassingly parallel over vertices,
which is what is sharded here.  There is no data-path collective other than the final all-gather of the
``[n_verts, order^2]`` coefficient rows.

Probe capture (``SH_volume::precompute``, reference src/sh/volume.cpp:149-316) shards by contiguous probe ranges; the
variable-length CSR slices, key sets and surfel accumulators are all-gathered and merged on the host
(``merge_probe_csr``): surfel ids are the rank of the cluster key, so the union of the ranks' key sets defines them
globally.
"""
from __future__ import annotations

import numpy as np

CHUNK = 64        # vertices per interleaved chunk: 2048 left the slowest of 8 ranks 4 % behind the mean, 64 leaves 0.5 %
                  # (tools/shard_balance.py); a chunk is still two 32-vertex Morton runs, enough for L1/L2 locality


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
    every rank.  ``bake_fn(pos_shard, nrm_shard, vertex_ids) -> [n, n2] float32`` is the per-rank compute; ``vertex_ids``
    are the positions of the shard's vertices in the whole list, which key the bounce RNG of an interreflection bake -- on a
    GPU rank the compute is ``prt_bake_transfer_device_shard(..., world, rank, ...)``, which derives exactly these ids from
    (world, rank) itself (``global_row`` in kernels.h), so a sharded bake does not depend on the number of ranks.
    Host-side reference implementation of the collective plumbing; bench.py keeps the rows on the device and calls
    ``all_gather_into_tensor`` directly."""
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


# ---- probe capture ---------------------------------------------------------------------------------------------------------

def probe_shard_range(n_probes: int, world: int, rank: int):
    """Contiguous probe range ``[lo, hi)`` of ``rank`` (probes keep the x-fastest order of volume.cpp:83-90, so the
    concatenation of the ranks' CSR slices is the whole CSR)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = (n_probes + world - 1) // world
    return min(rank * per, n_probes), min((rank + 1) * per, n_probes)


def merge_probe_csr(parts):
    """Merges per-rank captures of consecutive probe ranges.  ``parts``: list (rank order) of dicts with ``range`` [P_r,2]
    u32, ``ids`` [nnz_r] u32, ``transfer`` [nnz_r,9] f32, ``keys`` [S_r] u64 (ascending) and ``sums`` [S_r,7] f64.
    Returns ``(range, ids, transfer, surfels, keys)`` laid out exactly like a single capture of all probes."""
    keys = np.unique(np.concatenate([np.asarray(p["keys"], np.uint64) for p in parts]))
    sums = np.zeros((len(keys), 7), np.float64)
    rng_out, ids_out, tr_out = [], [], []
    base = 0
    for p in parts:
        remap = np.searchsorted(keys, np.asarray(p["keys"], np.uint64)).astype(np.uint32)
        ids = np.asarray(p["ids"], np.uint32)
        ids_out.append(remap[ids] if len(ids) else ids)
        tr_out.append(np.asarray(p["transfer"], np.float32).reshape(-1, 9))
        rng_out.append(np.asarray(p["range"], np.uint32).reshape(-1, 2) + np.uint32(base))
        if len(remap):
            np.add.at(sums, remap, np.asarray(p["sums"], np.float64).reshape(-1, 7))
        base += len(ids)
        if base >= 0xFFFFFFFF:
            raise ValueError("merged CSR exceeds 2^32 entries")
    cnt = sums[:, 6:7]
    nm = sums[:, 3:6] / cnt
    nm /= np.sqrt((nm * nm).sum(1, keepdims=True))                                   # volume.cpp:307-308
    surfels = np.concatenate([sums[:, 0:3] / cnt, nm], 1).astype(np.float32)
    return (np.concatenate(rng_out).astype(np.uint32), np.concatenate(ids_out).astype(np.uint32),
            np.concatenate(tr_out).astype(np.float32), surfels, keys)


def _all_gather_bytes(arr: np.ndarray, group, device):
    """all-gather of one variable-length array per rank (as bytes, padded to the longest) -> list of uint8 arrays."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    raw = np.frombuffer(np.ascontiguousarray(arr).tobytes(), np.uint8)
    n = torch.tensor([len(raw)], dtype=torch.int64, device=device)
    sizes = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    sizes = sizes.cpu().numpy()
    cap = max(int(sizes.max()), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=device)
    if len(raw):
        mine[:len(raw)] = torch.from_numpy(raw.copy()).to(device)
    out = torch.empty(world * cap, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    out = out.cpu().numpy().reshape(world, cap)
    return [out[r, :int(sizes[r])] for r in range(world)]


def sharded_probe_capture(capture_fn, probe_pos: np.ndarray, group=None, device="cpu"):
    """Captures ``probe_pos`` across the ranks of ``group`` and returns the merged CSR on every rank.
    ``capture_fn(probe_pos_slice) -> dict(range, ids, transfer, keys, sums)`` is the per-rank compute
    (``ProbeTransfer(...)`` + ``download()`` + ``surfel_sums()`` on a GPU rank).  ``device``: where the collective's
    buffers live ("cpu" for gloo, the rank's CUDA device for NCCL)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = probe_shard_range(len(probe_pos), world, rank)
    empty = dict(range=np.zeros((0, 2), np.uint32), ids=np.zeros(0, np.uint32), transfer=np.zeros((0, 9), np.float32),
                 keys=np.zeros(0, np.uint64), sums=np.zeros((0, 7), np.float64))
    part = capture_fn(probe_pos[lo:hi]) if hi > lo else empty
    if world == 1:
        return merge_probe_csr([part])
    dtypes = dict(range=np.uint32, ids=np.uint32, transfer=np.float32, keys=np.uint64, sums=np.float64)
    gathered = {k: _all_gather_bytes(np.asarray(part[k], dt), group, device) for k, dt in dtypes.items()}
    parts = [{k: np.frombuffer(gathered[k][r].tobytes(), dtypes[k]) for k in dtypes} for r in range(world)]
    return merge_probe_csr(parts)
