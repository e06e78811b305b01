import ctypes as C, sys, os, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
import prt_b200
from prt_b200 import meshes
pos, nrm, tri = meshes.bumpy_torus(737, 737)
order = meshes.morton_order(pos)
dev = torch.device("cuda", 0)
d_pos = torch.from_numpy(np.ascontiguousarray(pos[order])).to(dev)
d_nrm = torch.from_numpy(np.ascontiguousarray(nrm[order])).to(dev)
n = len(pos)
d_out = torch.zeros((n, 9), dtype=torch.float32, device=dev)
params = prt_b200.BakeParams.make()
ctxs = [prt_b200.Context(0), prt_b200.Context(0)]
scenes = [prt_b200.RTScene(pos, tri, c) for c in ctxs]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def bake(ci, lo, hi):
    c, sc = ctxs[ci], scenes[ci]
    rc = c.L.prt_bake_transfer_device(c.h, sc.h, C.c_void_p(d_pos.data_ptr() + 12 * lo), C.c_void_p(d_nrm.data_ptr() + 12 * lo), 12, hi - lo, 0,
                                      C.byref(params), C.c_void_p(d_out.data_ptr() + 36 * lo), None, C.c_void_p(streams[ci].cuda_stream))
    assert rc == 0
def run(K, cps, concurrent):
    for c in ctxs: c.set_tuning(ctas_per_sm=cps)
    bounds = [(i * n // K // 4 * 4, ((i + 1) * n // K // 4 * 4) if i < K - 1 else n) for i in range(K)]
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        for s in streams: s.wait_event(e0)
        for i, (lo, hi) in enumerate(bounds): bake(i % 2 if concurrent else 0, lo, hi)
        ev = [torch.cuda.Event() for _ in streams]
        for s, e in zip(streams, ev): e.record(s); torch.cuda.current_stream().wait_event(e)
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ref = None
for K, cps, conc in [(1, 0, False), (2, 0, False), (2, 6, True), (4, 6, True), (8, 6, True), (8, 5, True), (16, 6, True), (8, 0, True)]:
    ms = run(K, cps, conc)
    out = d_out.clone()
    if ref is None: ref = out
    print(f"chunks {K} ctas_per_sm {cps} concurrent {conc}: {ms:.2f} ms  same={bool(torch.equal(out, ref))}", flush=True)
