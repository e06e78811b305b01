#!/usr/bin/env python
"""tools/bench_configs.py -- one timing line per BASELINE config other than the headline (configs[1..4]), on one GPU.
Sizes are the named ones unless --quick. Output: JSON lines (copied to profiles/)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import prt_b200  # noqa: E402
from prt_b200 import hdr, meshes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--only", default="")
ap.add_argument("--budget", type=int, default=0, help="horizon_budget for configs 4/5 (0: library default)")
ap.add_argument("--near", default="", help="comma list of horizon_near values to time configs 4/5 with (default: the library default)")
a = ap.parse_args()
ctx = prt_b200.Context(0)
dev = torch.device("cuda", 0)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def want(name):
    return not a.only or name in a.only.split(",")


if want("c2"):
    eq = hdr.synthetic_env(1600, 800)
    t_create = timed(lambda: prt_b200.LightProbe(eq, 512, ctx).close(), 2)
    lp = prt_b200.LightProbe(eq, 512, ctx)
    t_irr = timed(lambda: lp.irradiance(32))
    t_pf = timed(lambda: lp.prefilter(256, 5, 1024))
    t_lut = timed(lambda: prt_b200.brdf_lut(512, 512, 1024, ctx))
    t_sh = timed(lambda: lp.project_sh(3, 0))
    emit(config="2: env SH + split-sum prefilter + BRDF LUT (1600x800 synthetic HDR, reference sizes)", seconds_incl_copies=dict(
        equirect_to_cube_and_mips=t_create, irradiance_32=t_irr, prefilter_256x5_1024spp=t_pf, brdf_lut_512_1024spp=t_lut, env_sh9_latlong=t_sh),
        prefilter_gsamples_per_s=523776 * 1024 / t_pf / 1e9, lut_gsamples_per_s=512 * 512 * 1024 / t_lut / 1e9,
        irradiance_gfetches_per_s=6144 * 15876 / t_irr / 1e9)

if want("c3"):
    rp = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * 6.18
    rt = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2]], np.uint32)
    nu = 200 if a.quick else 737
    tp, _, tt = meshes.bumpy_torus(nu, nu)
    pos = np.concatenate([rp, tp]).astype(np.float32)
    tri = np.concatenate([rt, tt + np.uint32(8)]).astype(np.uint32)
    sc = prt_b200.RTScene(pos, tri, ctx)
    res = 16 if a.quick else 32
    probes = prt_b200.probe_positions([res] * 3, [6.18] * 3)
    d, w = prt_b200.fibonacci_dirs(4096)
    prt_b200.ProbeTransfer(sc, probes[:64], d, w).close()                  # warm-up: module load, allocator
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        pt = prt_b200.ProbeTransfer(sc, probes, d, w)
        wall = time.perf_counter() - t0
        if best is None or pt.capture_ms < best[0].capture_ms:
            if best is not None:
                best[0].close()
            best = (pt, wall)
        else:
            pt.close()
    pt, wall = best
    rad = np.ones((pt.n_surfels, 4), np.float32)
    t_proj = timed(lambda: pt.project(rad))
    emit(config=f"3: probe capture {res}^3 probes x 4096 rays, room + buddha-scale torus ({len(tri)} triangles)", probes=len(probes),
         closest_hit_rays=len(probes) * 4096, capture_kernel_ms=pt.capture_ms, capture_wall_s=wall,
         grays_per_s=len(probes) * 4096 / pt.capture_ms / 1e6, nnz=pt.nnz, surfels=pt.n_surfels, project_s_incl_copies=t_proj)
    pt.close()

for name, order, su, sv, mode, bounces, nu, nv, nsub in [
        ("4: 3-bounce interreflection, order 4, 2.1M-vertex torus, 4096 samples/vertex", 4, 64, 64, 2, 3, 1448, 1448, 65536),
        ("5: shadowed, order 5, 20M-vertex torus, 8192 samples/vertex", 5, 128, 64, 1, 0, 5000, 4000, 131072)]:
    key = "c4" if name.startswith("4") else "c5"
    if not want(key):
        continue
    if a.quick:
        nu, nv, nsub = nu // 4, nv // 4, nsub // 8
    pos, nrm, tri = meshes.bumpy_torus(nu, nv)
    t0 = time.perf_counter()
    sc = prt_b200.RTScene(pos, tri, ctx)
    info = sc.info()
    order_idx = meshes.morton_order(pos)
    stride = max(1, len(pos) // nsub)
    sel = order_idx[::stride][:nsub]
    d_pos = torch.from_numpy(np.ascontiguousarray(pos[sel])).to(dev)
    d_nrm = torch.from_numpy(np.ascontiguousarray(nrm[sel])).to(dev)
    d_out = torch.zeros((len(sel), order * order), dtype=torch.float32, device=dev)
    params = prt_b200.BakeParams.make(order=order, samples_u=su, samples_v=sv, mode=mode, bounces=bounces,
                                      albedo=(0.5, 0.5, 0.5) if bounces else (1, 1, 1))
    stream = torch.cuda.current_stream()
    S = su * sv
    for near in ([int(x) for x in a.near.split(",")] if a.near else [None]):
        if near is not None:
            ctx.set_tuning(horizon_near=near)
        if a.budget:
            ctx.set_tuning(horizon_budget=a.budget)
        ms = []
        for _ in range(2):
            rc = ctx.L.prt_bake_transfer_device(ctx.h, sc.h, C.c_void_p(d_pos.data_ptr()), C.c_void_p(d_nrm.data_ptr()), 12, len(sel), 0,
                                                C.byref(params), C.c_void_p(d_out.data_ptr()), None, C.c_void_p(stream.cuda_stream))
            assert rc == 0, ctx.L.prt_last_error()
            torch.cuda.synchronize()
            ms.append(ctx.last_bake_stats().kernel_ms)
        emit(config=name, horizon_near=near, horizon_budget=a.budget or None, horizon_ms=ctx.last_bake_stats().horizon_ms, vertices_total=len(pos), triangles=len(tri), bvh_build_s=info.build_seconds,
             bvh_mb=(info.node_bytes + info.tri_bytes) / 1e6,
             bvh_depth=info.max_depth, vertices_timed=len(sel), sample=f"every {stride}-th vertex in Morton order", samples_per_vertex=S,
             kernel_ms=min(ms), primary_grays_per_s=len(sel) * S / min(ms) / 1e6, vertices_per_s=len(sel) / min(ms) * 1e3,
             full_mesh_seconds_extrapolated=len(pos) / (len(sel) / min(ms) * 1e3), finite=bool(torch.isfinite(d_out).all().item()))
    sc.close()

if want("f"):
    # SURVEY 8 rows f1 / f2 / f4 at the reference's default sizes (app.cpp:50-51: scene_size 12, probes 3.0 apart -> 8^3, volume 0.25 -> 96^3;
    # gl.h:314: 4096^2 shadow map; 1920x1080 preview)
    rp = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * 6.18
    rt = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2]], np.uint32)
    nu = 200 if a.quick else 737
    tp, _, tt = meshes.bumpy_torus(nu, nu)
    pos = np.concatenate([rp, tp]).astype(np.float32)
    tri = np.concatenate([rt, tt + np.uint32(8)]).astype(np.uint32)
    sc = prt_b200.RTScene(pos, tri, ctx)
    pres, vres, size = [8] * 3, [24 if a.quick else 96] * 3, [12.0] * 3
    prt_b200.calculate_weight(sc, [2] * 3, [4] * 3, size)
    t_w = timed(lambda: prt_b200.calculate_weight(sc, pres, vres, size), 2)
    weights = prt_b200.calculate_weight(sc, pres, vres, size)
    nvox = int(np.prod(vres))
    emit(config=f"f1: calculate_weight, {pres[0]}^3 probes, {vres[0]}^3 voxels x (100 closest-hit + 8 any-hit rays)", rays=nvox * 108,
         seconds_incl_copies=t_w, grays_per_s=nvox * 108 / t_w / 1e9)
    d, w = prt_b200.cube_dirs(26)
    pt = prt_b200.ProbeTransfer(sc, prt_b200.probe_positions(pres, size), d, w)
    sky, M = prt_b200.paral_shadow_matrix(0.17, 0.84)
    smap = 1024 if a.quick else 4096
    prt_b200.shadow_map(sc, M, 64)
    t_sm = timed(lambda: prt_b200.shadow_map(sc, M, smap), 2)
    vol = prt_b200.SHVolume(pt, pres, vres, size, weights)
    vol.set_shadow_map(prt_b200.shadow_map(sc, M, smap))
    P = prt_b200.RelightParams.make(sky, M)
    vol.step(P, 1)
    t_gi = timed(lambda: vol.step(P, 100), 2) / 100
    emit(config=f"f2: per-frame probe pipeline, {pt.n_surfels} surfels, {pt.nnz} CSR entries, {pres[0]}^3 probes -> {vres[0]}^3 voxels",
         shadow_map=f"{smap}^2 texels by ray casting", shadow_map_seconds_incl_copy=t_sm, shadow_grays_per_s=smap * smap / t_sm / 1e9,
         seconds_per_round_relight_project_blend=t_gi, rounds_per_second=1.0 / t_gi)
    W_, H_ = (480, 270) if a.quick else (1920, 1080)
    film = prt_b200.Film(W_, H_, ctx)
    cam = prt_b200.Camera.look_at((5.5, 3.0, 5.0), (0, -0.3, 0), zoom_deg=45)
    prt_b200.raytrace(sc, film, cam, max_path_length=3, albedo=(0.7, 0.7, 0.7), n_frames=1)
    t_rt = timed(lambda: prt_b200.raytrace(sc, film, cam, max_path_length=3, albedo=(0.7, 0.7, 0.7), n_frames=16), 2) / 16
    emit(config=f"f4: AO preview {W_}x{H_}, max_path_length 3 (inside the closed room: every path runs its full length)",
         seconds_per_frame=t_rt, mpaths_per_s=W_ * H_ / t_rt / 1e6)
