#!/usr/bin/env python
"""tools/probe_multi.py -- BASELINE config 3 (probe capture) sharded over the ranks of a torchrun job; rank 0 prints one JSON
line.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe_multi.py [--res 32]
With --check the merged CSR is compared with a single-GPU capture of all probes on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import prt_b200  # noqa: E402
from prt_b200 import dist as pdist, meshes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=32)
ap.add_argument("--dirs", type=int, default=4096)
ap.add_argument("--torus", type=int, default=737)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

half = 6.0
c = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * half
f = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2]], np.uint32)
tp, _, tt = meshes.bumpy_torus(a.torus, a.torus)
pos = np.concatenate([c, tp]).astype(np.float32)
tri = np.concatenate([f, tt + np.uint32(len(c))]).astype(np.uint32)
ctx = prt_b200.Context(local)
scene = prt_b200.RTScene(pos, tri, ctx)
probes = prt_b200.probe_positions([a.res] * 3, [half] * 3)
d, w = prt_b200.fibonacci_dirs(a.dirs)
kernel_ms = [0.0]


def capture(pp):
    pt = prt_b200.ProbeTransfer(scene, pp, d, w)
    kernel_ms[0] = pt.capture_ms
    rng, ids, tr, _, keys = pt.download()
    return dict(range=rng, ids=ids, transfer=tr, keys=keys, sums=pt.surfel_sums())


capture(probes[:8])                                     # warm-up (module load, allocator)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
rng, ids, tr, sf, keys = pdist.sharded_probe_capture(capture, probes, device=torch.device("cuda", local))
torch.cuda.synchronize()
t1 = time.perf_counter()
km = torch.tensor([kernel_ms[0], (t1 - t0) * 1e3], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(km, op=dist.ReduceOp.MAX)
ok = None
if a.check and rank == 0:
    whole = prt_b200.ProbeTransfer(scene, probes, d, w)
    wr, wi, wt, wsf, wk = whole.download()
    ok = bool(np.array_equal(rng, wr) and np.array_equal(ids, wi) and np.array_equal(tr, wt) and np.array_equal(keys, wk)
              and np.abs(sf - wsf).max() <= 1e-6)
if rank == 0:
    n_rays = len(probes) * a.dirs
    print(json.dumps({"metric": "probe_closest_hit_rays_per_sec", "n_gpus": world, "probes": len(probes), "dirs": a.dirs,
                      "capture_kernel_ms_max": float(km[0]), "rays_per_sec_kernel": n_rays / (float(km[0]) * 1e-3),
                      "wall_ms_incl_gather_merge": float(km[1]), "nnz": int(len(ids)), "surfels": int(len(sf)),
                      "merged_equals_single": ok}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
