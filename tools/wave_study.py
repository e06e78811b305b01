#!/usr/bin/env python
"""tools/wave_study.py -- CPU model of the two-pass shadow bake (no GPU): runs the product's horizon builder and traversal code on
the warp emulator (tests/hostcheck) for a Morton-strided sample of the bench mesh and prints the work counters bench.py reports
from an instrumented GPU launch (node visits, triangle tests, entry-list box tests per ray, share of rays traversed), so that
algorithmic changes can be scored before any GPU time is spent.
Usage: python tools/wave_study.py [--nu 737 --nv 737] [--n 96] [--near 157 --mid 24 --gain 0.2 --slabs 1] [--budget 64]  (old builder: --near 30 --mid 0 --slabs 0)"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402
from oracle import pyoracle as oracle  # noqa: E402
from prt_b200 import meshes  # noqa: E402
from test_horizon_math import _maps  # noqa: E402
from test_wave_emulated import processing_table, run_wave  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nu", type=int, default=737)
ap.add_argument("--nv", type=int, default=737)
ap.add_argument("--n", type=int, default=96)
ap.add_argument("--near", type=int, default=157)
ap.add_argument("--mid", type=int, default=24)
ap.add_argument("--gain", type=float, default=0.2)
ap.add_argument("--slabs", type=int, default=1)
ap.add_argument("--wave-dop", type=int, default=1, help="fourth slab axis in the node test of the traversal pass (product default: on)")
ap.add_argument("--dop", type=int, default=0, help="study: fourth slab axis per node in the child test (1 quantised, 2 exact extents)")
ap.add_argument("--budget", type=int, default=64)
ap.add_argument("--folds", action="store_true", help="the heavily self-occluding twin of the bench mesh (amp 0.25, fscale 3)")
a = ap.parse_args()

hc = conftest.load_hostcheck()
pos, nrm, tri = meshes.bumpy_torus(a.nu, a.nv, amp=0.25, fscale=3) if a.folds else meshes.bumpy_torus(a.nu, a.nv)
order = meshes.morton_order(pos)
sel = order[:: max(1, len(order) // a.n)][: a.n]
h = hc.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
op = oracle.make_params(order=3, samples_u=32, samples_v=32)
tab, bins = processing_table(oracle, op)
hz, _ = _maps(hc, h, pos[sel], nrm[sel], budget=a.budget, near=a.near, slabs=a.slabs, mid=a.mid, gain=a.gain)
need = ~(tab[None, :, 2] > hz[:, bins])
need_words = np.ascontiguousarray(np.packbits(need, axis=1, bitorder="little")).view(np.uint32).copy()
keep = need.any(axis=1)                         # vertices the horizon pass finishes are never seen by the traversal pass
work = np.zeros(4, np.uint64)
hc.hc_dop_study.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_int, __import__("ctypes").c_void_p]
hc.hc_dop_study(h, a.dop, None)
hc.hc_wave_dop(a.wave_dop)
hc.hc_wave_step_stats.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_int]
_ss = np.zeros(8, np.uint64); hc.hc_wave_step_stats(_ss.ctypes.data, 1)
got, vis = run_wave(hc, h, pos[sel][keep], nrm[sel][keep], tab, 3, need=np.ascontiguousarray(need_words[keep]), work=work)
hc.hc_wave_step_stats(_ss.ctypes.data, 0)
_ds = np.zeros(3, np.uint64); hc.hc_dop_study(h, 0, _ds.ctypes.data)
if a.dop:
    print(json.dumps({"dop_study": {"mode": a.dop, "hit_children_tested": int(_ds[0]), "culled_inner_share": float(_ds[1]) / max(1.0, float(_ds[0])),
                                    "culled_leaf_share": float(_ds[2]) / max(1.0, float(_ds[0]))}}))
print(json.dumps({"steps_per_vertex": {k: round(float(_ss[2 * i]) / len(sel), 1) for i, k in enumerate(["leaf", "node", "scan"])},
                  "stack_overflows_per_vertex": {"subtrees": round(float(_ss[6]) / len(sel), 3), "leaves": round(float(_ss[7]) / len(sel), 3)},
                  "lanes_per_step": {k: round(float(_ss[2 * i + 1]) / max(1.0, float(_ss[2 * i])), 1) for i, k in enumerate(["leaf", "node", "scan"])}}))
rays = float(len(sel) * len(tab))
trav = float(work[3])
print(json.dumps({"vertices": len(sel), "finished_by_horizon_pass": int((~keep).sum()), "rays_traversed_frac": trav / rays,
                  "node_visits_per_ray": float(work[0]) / rays, "tri_tests_per_ray": float(work[1]) / rays,
                  "entry_list_box_tests_per_ray": float(work[2]) / rays,
                  "per_traversed_ray": {"node_visits": float(work[0]) / trav, "tri_tests": float(work[1]) / trav, "box_tests": float(work[2]) / trav},
                  "gpu_reference_r1": {"rays_traversed_frac": 0.2816, "node_visits_per_ray": 2.93, "tri_tests_per_ray": 3.35,
                                       "entry_list_box_tests_per_ray": 13.78}}))
hc.hc_free(h)
