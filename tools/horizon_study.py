#!/usr/bin/env python
"""tools/horizon_study.py -- CPU study of the horizon map's tightness (no GPU): runs the product's horizon builder on the warp
emulator (tests/hostcheck) for a sample of vertices of the bench mesh and compares the share of rays it leaves to trace with
  * the share that is really occluded (oracle), and
  * the best ANY 32-bin map could do on these samples (per bin: the highest occluded sample),
so that the slack of the bounds can be told from the cost of the per-bin structure itself.
Usage: python tools/horizon_study.py [--nu 737 --nv 737] [--n 200] [--near 30,24] [--budget 64]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402  (builds tests/hostcheck/libhostcheck.so)
from oracle import pyoracle as oracle  # noqa: E402
from prt_b200 import meshes  # noqa: E402
from test_horizon_math import BINS, _free_mask, _maps, pang  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nu", type=int, default=737)
ap.add_argument("--nv", type=int, default=737)
ap.add_argument("--n", type=int, default=200)
ap.add_argument("--near", default="157")
ap.add_argument("--budget", type=int, default=64)
ap.add_argument("--folds", action="store_true", help="the heavily self-occluding twin of the bench mesh (amp 0.25, fscale 3)")
ap.add_argument("--mid", default="24", help="mid-size refinement rule: angular radius x 100 (0 = off), comma list")
ap.add_argument("--gain", default="0.2", help="... its threshold in units of S / 32 samples, comma list")
ap.add_argument("--slabs", default="1", help="oriented slabs of the nodes in the builder: 1, 0 or 0,1 to compare")
a = ap.parse_args()

hc = conftest.load_hostcheck()
pos, nrm, tri = meshes.bumpy_torus(a.nu, a.nv, amp=0.25, fscale=3) if a.folds else meshes.bumpy_torus(a.nu, a.nv)
order = meshes.morton_order(pos)
sel = order[:: max(1, len(order) // a.n)][: a.n]
h = hc.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
op = oracle.make_params(order=3, samples_u=32, samples_v=32)
_, vis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), pos[sel], nrm[sel], op, want_vis=True)
visible = np.unpackbits(vis.view(np.uint8), axis=1, bitorder="little")[:, :1024].astype(bool)
_, dirs = oracle.sample_table(op)
den = np.abs(dirs[:, 0].astype(np.float64)) + np.abs(dirs[:, 1].astype(np.float64))
bins = np.clip(np.floor(np.where(den > 0, pang(dirs[:, 0].astype(np.float64), dirs[:, 1].astype(np.float64)), 0.0) * (BINS / 4)).astype(int), 0, BINS - 1)
# best possible per-bin map on these samples: the highest occluded sample of each bin
ideal = np.zeros((len(sel), BINS))
for b in range(BINS):
    m = bins == b
    zz = np.where(~visible[:, m], dirs[None, m, 2], 0.0)
    ideal[:, b] = zz.max(axis=1) if m.any() else 0.0
ideal_traversed = (dirs[None, :, 2] <= ideal[:, bins]).mean()
out = {"vertices": len(sel), "occluded": float(1.0 - visible.mean()), "best_32_bin_map_traversed": float(ideal_traversed)}
print(json.dumps(out))
for near in [int(x) for x in a.near.split(",")]:
  for mid in [int(x) for x in a.mid.split(",")]:
   for gain in [float(x) for x in a.gain.split(",")]:
    for slabs in [int(x) for x in a.slabs.split(",")]:
        hc.hc_use_slabs(slabs)
        hc.hc_horizon_mid(mid, gain)
        st = np.zeros(4, np.uint64)
        hz, ncand = _maps(hc, h, pos[sel], nrm[sel], budget=a.budget, near=near, stats=st, slabs=slabs, mid=mid, gain=gain)
        free = _free_mask(dirs, hz)
        assert not (free & ~visible).any()
        print(json.dumps({"near": near, "budget": a.budget, "slabs": slabs, "mid": mid, "gain": gain, "traversed": round(float(1.0 - free.mean()), 4),
                          "map_minus_ideal": round(float(np.mean(hz - ideal)), 4),
                          "per_vertex": {"iterations": round(float(st[0]) / len(sel), 2), "nodes_expanded": round(float(st[1]) / len(sel), 1),
                                         "boxes_bounded": round(float(st[2]) / len(sel), 1), "triangle_rounds": round(float(st[3]) / len(sel), 2)}}), flush=True)
hc.hc_free(h)
