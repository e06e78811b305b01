#!/usr/bin/env python
"""tools/shard_balance.py -- load balance of bench.py's vertex sharding, emulated on ONE GPU: times every rank's shard of a
world-way split (interleaved chunks of the Morton-ordered vertex list) one after the other and prints max / mean per chunk size.
Usage: python tools/shard_balance.py --world 8 --chunks 2048,512,256,128"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import prt_b200  # noqa: E402
from prt_b200 import meshes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nu", type=int, default=737)
ap.add_argument("--nv", type=int, default=737)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--chunks", default="2048,512,256,128")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tune", default="")
a = ap.parse_args()

pos, nrm, tri = meshes.bumpy_torus(a.nu, a.nv)
morton = meshes.morton_order(pos)
ctx = prt_b200.Context(0)
for kv in filter(None, a.tune.split(",")):
    k, v = kv.split("=")
    ctx.set_tuning(**{k: int(v)})
scene = prt_b200.RTScene(pos, tri, ctx)
dev = torch.device("cuda", 0)
params = prt_b200.BakeParams.make(order=3, samples_u=32, samples_v=32)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()
L = ctx.L
for chunk in [int(x) for x in a.chunks.split(",")]:
    n_chunks = (len(morton) + chunk - 1) // chunk
    per_rank = []
    for rank in range(a.world):
        sel = np.concatenate([morton[c * chunk:(c + 1) * chunk] for c in range(rank, n_chunks, a.world)])
        d_pos = torch.from_numpy(np.ascontiguousarray(pos[sel])).to(dev)
        d_nrm = torch.from_numpy(np.ascontiguousarray(nrm[sel])).to(dev)
        d_out = torch.zeros((len(sel), 9), dtype=torch.float32, device=dev)
        ms = []
        for _ in range(a.reps):
            flush.zero_()
            rc = L.prt_bake_transfer_device(ctx.h, scene.h, C.c_void_p(d_pos.data_ptr()), C.c_void_p(d_nrm.data_ptr()), 12, len(sel), 0,
                                            C.byref(params), C.c_void_p(d_out.data_ptr()), None, C.c_void_p(stream.cuda_stream))
            assert rc == 0, L.prt_last_error()
            torch.cuda.synchronize()
            ms.append(ctx.last_bake_stats().kernel_ms)
        per_rank.append(min(ms))
    print(json.dumps({"world": a.world, "chunk": chunk, "tune": a.tune, "max_ms": max(per_rank), "mean_ms": float(np.mean(per_rank)),
                      "per_rank_ms": [round(x, 3) for x in per_rank]}), flush=True)
