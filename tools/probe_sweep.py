#!/usr/bin/env python
"""tools/probe_sweep.py -- config-3 probe capture kernel time over refill_thresh values on one GPU."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prt_b200
from prt_b200 import meshes
ctx = prt_b200.Context(0)
rp = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * 6.18
rt = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2]], np.uint32)
tp, _, tt = meshes.bumpy_torus(737, 737)
sc = prt_b200.RTScene(np.concatenate([rp, tp]).astype(np.float32), np.concatenate([rt, tt + np.uint32(8)]).astype(np.uint32), ctx)
probes = prt_b200.probe_positions([32] * 3, [6.18] * 3)
d, w = prt_b200.fibonacci_dirs(4096)
prt_b200.ProbeTransfer(sc, probes[:64], d, w).close()
for thr in [int(x) for x in (sys.argv[1:] or ["0", "8", "16", "24"])]:
    ctx.set_tuning(refill_thresh=thr)
    ms = []
    for _ in range(3):
        pt = prt_b200.ProbeTransfer(sc, probes, d, w); ms.append(pt.capture_ms); nnz = pt.nnz; pt.close()
    print(json.dumps({"refill_thresh": thr, "capture_kernel_ms": min(ms), "grays_per_s": len(probes) * 4096 / min(ms) / 1e6, "nnz": nnz}), flush=True)
