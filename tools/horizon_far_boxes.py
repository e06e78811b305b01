#!/usr/bin/env python
"""tools/horizon_far_boxes.py -- CPU study (warp emulator): the far boxes whose cheap bound the horizon builder merges, compared with
the exact bound of the same axis-aligned box (hz_box).  Tells how much of the map's slack a better bound of the SAME box could remove."""
import sys, os, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import conftest
from prt_b200 import meshes
hc = conftest.load_hostcheck()
pos, nrm, tri = meshes.bumpy_torus(737, 737)
order = meshes.morton_order(pos)
sel = order[:: len(order)//60][:60]
h = hc.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
rows=[]
for v in sel:
    buf = np.zeros((4096, 9), np.float32)
    p = np.ascontiguousarray(pos[v]); n = np.ascontiguousarray(nrm[v])
    k = hc.hc_horizon_trace_far(h, p.ctypes.data, n.ctypes.data, 1e-4, 64, 30, buf.ctypes.data, 4096)
    b = buf[:k]
    c = np.ascontiguousarray(b[:, 0:3]); e = np.ascontiguousarray(b[:, 3:6]); nn = np.ascontiguousarray(np.tile(n, (k, 1)))
    bins = np.zeros((k, 2), np.int32); val = np.zeros(k, np.float32)
    hc.hc_hz_box(c.ctypes.data, e.ctypes.data, nn.ctypes.data, k, 1, bins.ctypes.data, val.ctypes.data)
    rows.append(np.stack([b[:, 6], val, b[:, 8] - b[:, 7] + 1, (bins[:, 1] - bins[:, 0] + 1).astype(np.float32)], 1))
r = np.concatenate(rows)
print(json.dumps({"far_boxes_merged_per_vertex": len(r) / len(sel), "cheap_v_mean": float(r[:, 0].mean()), "exact_aabb_v_mean": float(r[:, 1].mean()),
                  "mean_excess_cheap_over_exact_aabb": float((r[:, 0] - r[:, 1]).mean()), "p90_excess": float(np.percentile(r[:, 0] - r[:, 1], 90)),
                  "bins_cheap_mean": float(r[:, 2].mean()), "bins_exact_mean": float(r[:, 3].mean())}))
