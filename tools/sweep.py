#!/usr/bin/env python
"""tools/sweep.py -- times the bake kernel over tuning-knob combinations on one GPU (kernel_ms from the library's
own CUDA events).  Usage: python tools/sweep.py [--nu 737 --nv 737] [--order 3] [--su 32 --sv 32] knob=v1,v2 ..."""
import argparse
import ctypes as C
import itertools
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
ap.add_argument("--order", type=int, default=3)
ap.add_argument("--su", type=int, default=32)
ap.add_argument("--sv", type=int, default=32)
ap.add_argument("--mode", type=int, default=1)
ap.add_argument("--bounces", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--mesh", default="torus", help="torus | folds (the heavily self-occluding twin of bench.py --workload folds) | icoK")
ap.add_argument("--world", type=int, default=1, help="time the shard rank --rank of a --world-way split (bench.py's interleaved chunks)")
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--chunk", type=int, default=64)
ap.add_argument("--flush", action="store_true", help="flush L2 between repetitions, as bench.py does")
ap.add_argument("knobs", nargs="*")
a = ap.parse_args()

if a.mesh == "torus":
    pos, nrm, tri = meshes.bumpy_torus(a.nu, a.nv)
elif a.mesh == "folds":
    pos, nrm, tri = meshes.bumpy_torus(a.nu, a.nv, amp=0.25, fscale=3)
else:
    pos, nrm, tri = meshes.icosphere(int(a.mesh.replace("ico", "")))
order = meshes.morton_order(pos)
if a.world > 1:
    n_chunks = (len(order) + a.chunk - 1) // a.chunk
    order = np.concatenate([order[c * a.chunk:(c + 1) * a.chunk] for c in range(a.rank, n_chunks, a.world)])
ctx = prt_b200.Context(0)
scene = prt_b200.RTScene(pos, tri, ctx)
dev = torch.device("cuda", 0)
d_pos = torch.from_numpy(np.ascontiguousarray(pos[order])).to(dev)
d_nrm = torch.from_numpy(np.ascontiguousarray(nrm[order])).to(dev)
n = len(order)
params = prt_b200.BakeParams.make(order=a.order, samples_u=a.su, samples_v=a.sv, mode=a.mode, bounces=a.bounces,
                                  albedo=(0.5, 0.5, 0.5) if a.bounces else (1, 1, 1))
d_out = torch.zeros((n, a.order ** 2), dtype=torch.float32, device=dev)
L = ctx.L
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if a.flush else None
stream = torch.cuda.current_stream()
names = [k.split("=")[0] for k in a.knobs]
values = [[int(x) for x in k.split("=")[1].split(",")] for k in a.knobs]
S = a.su * a.sv
for combo in itertools.product(*values) if values else [()]:
    kw = dict(zip(names, combo))
    ctx.set_tuning(**kw)
    ms = []
    for _ in range(a.reps):
        if a.flush:
            flush_buf.zero_()
        rc = L.prt_bake_transfer_device(ctx.h, scene.h, C.c_void_p(d_pos.data_ptr()), C.c_void_p(d_nrm.data_ptr()), 12, n, 0,
                                        C.byref(params), C.c_void_p(d_out.data_ptr()), None, C.c_void_p(stream.cuda_stream))
        assert rc == 0, L.prt_last_error()
        torch.cuda.synchronize()
        ms.append(ctx.last_bake_stats().kernel_ms)
    ctx.set_tuning(count_work=1)
    L.prt_bake_transfer_device(ctx.h, scene.h, C.c_void_p(d_pos.data_ptr()), C.c_void_p(d_nrm.data_ptr()), 12, n, 0,
                               C.byref(params), C.c_void_p(d_out.data_ptr()), None, C.c_void_p(stream.cuda_stream))
    torch.cuda.synchronize()
    st = ctx.last_bake_stats()
    ctx.set_tuning(count_work=0)
    best = min(ms)
    print(json.dumps({"knobs": kw, "kernel_ms": best, "horizon_ms": round(st.horizon_ms, 2), "grays_per_s": n * S / best / 1e6, "grid": st.grid, "block": st.block,
                      "nodes_per_ray": round(st.node_visits / (n * S), 2), "tris_per_ray": round(st.tri_tests / (n * S), 2),
                      "cands_per_ray": round(st.cand_tests / (n * S), 2), "traversed_frac": round(st.rays_traversed / (n * S), 3),
                      "all_ms": [round(m, 2) for m in ms]}), flush=True)
