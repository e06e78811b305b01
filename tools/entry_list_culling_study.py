#!/usr/bin/env python
"""tools/entry_list_culling_study.py -- CPU study (warp emulator): how many of the traversal pass's entry-list box tests could be
skipped if every candidate carried the horizon builder's cheap elevation bound -- per lane (the ray's z above the bound), per
32-ray scan round (the round's lowest z above the bound: a warp-uniform skip), and with the candidate's azimuth range as well."""
import sys, os, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import conftest
from oracle import pyoracle as oracle
from prt_b200 import meshes
from test_horizon_math import BINS, _maps, pang
from test_wave_emulated import processing_table
hc = conftest.load_hostcheck()
pos, nrm, tri = meshes.bumpy_torus(737, 737)
order = meshes.morton_order(pos)
n=80
sel = order[:: len(order)//n][:n]
h = hc.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
op = oracle.make_params(order=3, samples_u=32, samples_v=32)
tab, bins = processing_table(oracle, op)
hz, _ = _maps(hc, h, pos[sel], nrm[sel])
need = ~(tab[None, :, 2] > hz[:, bins])
tot=0; lane_skip=0; batch_skip=0; batch_skip_az=0
for vi, v in enumerate(sel):
    buf = np.zeros((96, 8), np.float32)
    p = np.ascontiguousarray(pos[v]); nn = np.ascontiguousarray(nrm[v])
    k = hc.hc_entry_list(h, p.ctypes.data, nn.ctypes.data, 1e-4, buf.ctypes.data)
    c = np.ascontiguousarray(buf[:k, 0:3]); e = np.ascontiguousarray(buf[:k, 3:6]); N = np.ascontiguousarray(np.tile(nn, (k, 1)))
    b01 = np.zeros((k, 2), np.int32); val = np.zeros(k, np.float32)
    hc.hc_hz_box(c.ctypes.data, e.ctypes.data, N.ctypes.data, k, 0, b01.ctypes.data, val.ctypes.data)     # cheap bound per candidate
    idx = np.where(need[vi])[0]                    # flagged samples in processing order
    for s0 in range(0, len(idx), 32):
        batch = idx[s0:s0+32]
        z = tab[batch, 2]; bb = bins[batch]
        zmin = z.min()
        tot += len(batch) * k
        lane_skip += (z[:, None] > val[None, :]).sum()
        bs = val < zmin
        batch_skip += len(batch) * bs.sum()
        # azimuth: candidate's bin range vs the batch's bins
        inb = np.zeros(k, bool)
        for j in range(k):
            if b01[j,1]-b01[j,0] >= BINS-1: inb[j]=True
            else: inb[j] = (((bb - b01[j,0]) % BINS) <= (b01[j,1]-b01[j,0])).any()
        batch_skip_az += len(batch) * (bs | ~inb).sum()
print(json.dumps({"box_tests": int(tot), "skippable_per_lane_by_elevation": lane_skip/tot, "skippable_per_batch_by_elevation": batch_skip/tot, "skippable_per_batch_elevation_or_azimuth": batch_skip_az/tot}))
