#!/usr/bin/env python
"""tools/sanitize_small.py -- one small call of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prt_b200
from prt_b200 import meshes, hdr

pos, nrm, tri = meshes.bumpy_torus(40, 28)
sc = prt_b200.RTScene(pos, tri)
sel = np.arange(0, len(pos), 5)
for kw in (dict(), dict(order=4, mode=prt_b200.INTERREFLECT, bounces=2, albedo=(0.5, 0.5, 0.5)), dict(mode=prt_b200.UNSHADOWED), dict(order=5, samples_u=8, samples_v=8)):
    out, vis = prt_b200.bake_transfer(sc, pos[sel], nrm[sel], prt_b200.BakeParams.make(samples_u=kw.pop("samples_u", 16), samples_v=kw.pop("samples_v", 16), **kw), want_vis=True)
    assert np.isfinite(out).all()
for knobs in (dict(horizon=0), dict(pair_queue=0)):
    sc.ctx.set_tuning(**knobs)
    prt_b200.bake_transfer(sc, pos[sel], nrm[sel], prt_b200.BakeParams.make(samples_u=16, samples_v=16))
    prt_b200.bake_transfer(sc, pos[sel], nrm[sel], prt_b200.BakeParams.make(samples_u=8, samples_v=8, mode=prt_b200.INTERREFLECT, bounces=1))
    sc.ctx.set_tuning(horizon=1, pair_queue=2)
rays = prt_b200.RTScene.pack_rays(pos[:64] + 1e-3 * nrm[:64], nrm[:64])
sc.any_hit(rays); sc.first_hit(rays)
probes = prt_b200.probe_positions([2, 2, 2], [3, 3, 3])
d, w = prt_b200.fibonacci_dirs(700)
pt = prt_b200.ProbeTransfer(sc, probes, d, w)
pt.project(np.ones((pt.n_surfels, 4), np.float32))
wts = prt_b200.calculate_weight(sc, [2, 2, 2], [4, 4, 4], [3, 3, 3])
sky, M = prt_b200.paral_shadow_matrix(0.17, 0.84)
vol = prt_b200.SHVolume(pt, [2, 2, 2], [4, 4, 4], [3, 3, 3], wts)
vol.set_shadow_map(prt_b200.shadow_map(sc, M, 64))
vol.step(prt_b200.RelightParams.make(sky, M), 3)
film = prt_b200.Film(40, 24, sc.ctx)
prt_b200.raytrace(sc, film, prt_b200.Camera.look_at((5, 3, 4), (0, 0, 0)), n_frames=2)
lp = prt_b200.LightProbe(hdr.synthetic_env(64, 32), 16, sc.ctx)
lp.irradiance(4); lp.prefilter(8, 2, 32); lp.project_sh(3, 0); lp.project_sh(3, 1)
prt_b200.brdf_lut(16, 16, 32, sc.ctx)
# multi-GPU driver (the same device twice on a one-GPU box): sharded bake with fused P2P row stores, sorted work list, device-side CSR merge
grp = prt_b200.Group([0, 0])
grp.set_scene(pos, tri)
order = meshes.morton_order(pos)
rows, st = grp.bake_transfer(pos[order], nrm[order], prt_b200.BakeParams.make(samples_u=16, samples_v=16))
assert np.isfinite(rows).all() and st.n_devices == 2
sc.ctx.set_tuning(work_list=1)
prt_b200.bake_transfer(sc, pos[sel], nrm[sel], prt_b200.BakeParams.make(samples_u=16, samples_v=16))
sc.ctx.set_tuning(work_list=-1)
gpt, _, _ = grp.probe_capture(prt_b200.probe_positions([3, 2, 2], [3, 3, 3]), d, w)
gpt.close(); grp.close()
print("sanitize_small: all kernels ran")
