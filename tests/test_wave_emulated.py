"""The headline traversal kernel's per-vertex code -- bake_wave_vertex (prt_b200/csrc/bake_wave.cuh: entry list, lockstep candidate
scan, node / leaf wavefront steps, projection, visibility permutation), exactly what one persistent warp of bake_wave_kernel runs --
executed UNMODIFIED on the CPU through the warp emulator (tests/hostcheck/warp_emu.h) and compared with the oracle: visibility bits
bit for bit, SH rows within 1e-4 relative L2 (north_star).  Also the two-pass pipeline: horizon maps from the emulated horizon
builder -> need bits -> traversal of the flagged samples only, which must give the same bits."""
import numpy as np
import pytest

from prt_b200 import meshes
from test_horizon_math import BINS, _maps, pang

REL_L2_TOL = 1e-4


def processing_table(oracle, params):
    """The sample table abi.cu (ensure_samples) uploads: strata in Morton order over (i, j); w = reference index | bin << 24."""
    _, dirs = oracle.sample_table(params)
    ru, rv = params.samples_u, params.samples_v

    def spread(x):
        x &= 0xFFFF; x = (x | (x << 8)) & 0x00FF00FF; x = (x | (x << 4)) & 0x0F0F0F0F
        x = (x | (x << 2)) & 0x33333333; x = (x | (x << 1)) & 0x55555555
        return x
    keys = sorted((spread(j) | (spread(i) << 1), i * rv + j) for i in range(ru) for j in range(rv))
    sref = np.array([s for _, s in keys], np.uint32)
    d = dirs[sref]
    den = np.abs(d[:, 0].astype(np.float64)) + np.abs(d[:, 1].astype(np.float64))
    pa = np.where(den > 0, pang(d[:, 0].astype(np.float64), d[:, 1].astype(np.float64)), 0.0)
    bins = np.clip(np.floor(pa * (BINS / 4)).astype(np.int64), 0, BINS - 1).astype(np.uint32)
    tab = np.zeros((len(sref), 4), np.float32)
    tab[:, :3] = d
    tab[:, 3] = (sref | (bins << 24)).view(np.float32)
    return tab, bins


def run_wave(hostcheck, h, pos, nrm, tab, order, need=None, want_vis=True, cs_phase=0, work=None):
    n, S = len(pos), len(tab)
    words = (S + 31) // 32
    out = np.zeros((n, order * order), np.float32)
    vis = np.zeros((n, words), np.uint32) if want_vis else None
    p32, n32 = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nrm, np.float32)
    rc = hostcheck.hc_bake_wave(h, p32.ctypes.data, n32.ctypes.data, n, tab.ctypes.data, S, order,
                                need.ctypes.data if need is not None else None, 1e-4, cs_phase, out.ctypes.data,
                                vis.ctypes.data if vis is not None else None, work.ctypes.data if work is not None else None)
    assert rc == 0
    return out, vis


def rel_l2(a, b):
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-20)


@pytest.fixture(scope="module")
def scene(hostcheck, oracle):
    pos, nrm, tri = meshes.bumpy_torus(96, 64)
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    yield pos, nrm, tri, h, oracle.Scene(pos, tri)
    hostcheck.hc_free(h)


@pytest.mark.parametrize("order,su,sv", [(3, 16, 16), (5, 8, 32), (2, 5, 7), (3, 32, 32)])
def test_wave_vertex_matches_oracle(hostcheck, oracle, scene, order, su, sv):
    pos, nrm, tri, h, osc = scene
    sel = np.arange(5, len(pos), 53)[:40 if su * sv <= 256 else 16]
    op = oracle.make_params(order=order, samples_u=su, samples_v=sv)
    tab, _ = processing_table(oracle, op)
    got, gvis = run_wave(hostcheck, h, pos[sel], nrm[sel], tab, order)
    ref, ovis, _ = oracle.bake_transfer(osc, pos[sel], nrm[sel], op, want_vis=True)
    assert np.array_equal(gvis, ovis), "per-ray visibility bits must agree exactly"
    assert rel_l2(got, ref).max() <= REL_L2_TOL
    frac = np.unpackbits(ovis.view(np.uint8)).sum() / (len(sel) * su * sv)
    assert 0.3 < frac < 0.99
    # rows do not depend on whether the caller asks for the visibility words
    got2, _ = run_wave(hostcheck, h, pos[sel], nrm[sel], tab, order, want_vis=False)
    assert np.array_equal(got.view(np.uint32), got2.view(np.uint32))


def test_two_pass_pipeline_on_the_emulator(hostcheck, oracle, scene):
    """horizon maps (emulated horizon builder) -> need bits in processing order -> traversal of the flagged samples only"""
    pos, nrm, tri, h, osc = scene
    sel = np.arange(11, len(pos), 67)[:24]
    op = oracle.make_params(order=3, samples_u=32, samples_v=32)
    tab, bins = processing_table(oracle, op)
    hz, _ = _maps(hostcheck, h, pos[sel], nrm[sel])
    need = ~(tab[None, :, 2] > hz[:, bins])                               # horizon_kernel's classification (horizon.cu)
    need_words = np.ascontiguousarray(np.packbits(need, axis=1, bitorder="little")).view(np.uint32).copy()
    assert 0.05 < need.mean() < 0.95
    got, gvis = run_wave(hostcheck, h, pos[sel], nrm[sel], tab, 3, need=need_words)
    ref, ovis, _ = oracle.bake_transfer(osc, pos[sel], nrm[sel], op, want_vis=True)
    assert np.array_equal(gvis, ovis)
    assert rel_l2(got, ref).max() <= REL_L2_TOL
    full, fvis = run_wave(hostcheck, h, pos[sel], nrm[sel], tab, 3)
    assert np.array_equal(fvis, gvis) and np.array_equal(full.view(np.uint32), got.view(np.uint32))


# ---- interreflection: bake_inter_vertex (prt_b200/csrc/bake_inter.cuh) on the warp emulator -----------------------------------------
def run_inter(hostcheck, h, pos, nrm, tab, order, need_words, bounces, albedo, seed, vid_base=0):
    n, S = len(pos), len(tab)
    words = (S + 31) // 32
    out = np.zeros((n, order * order), np.float32)
    vis = np.zeros((n, words), np.uint32)
    need_words = np.ascontiguousarray(need_words, np.uint32)
    need_count = np.ascontiguousarray(np.unpackbits(need_words.view(np.uint8), axis=1, bitorder="little")[:, :S].sum(axis=1), np.uint32)
    p32, n32 = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nrm, np.float32)
    alb = np.ascontiguousarray(albedo, np.float32)
    rc = hostcheck.hc_bake_inter(h, p32.ctypes.data, n32.ctypes.data, n, vid_base, tab.ctypes.data, S, order, need_words.ctypes.data,
                                 need_count.ctypes.data, seed, bounces, alb.ctypes.data, 1e-4, 1e-5, out.ctypes.data, vis.ctypes.data)
    assert rc == 0
    return out, vis, need_count


@pytest.mark.parametrize("bounces,albedo,use_horizon", [(1, (1.0, 1.0, 1.0), False), (3, (0.5, 0.5, 0.5), True), (6, (0.5, 0.3, 0.2), True)])
def test_inter_vertex_matches_oracle(hostcheck, oracle, scene, bounces, albedo, use_horizon):
    """Asynchronous wavefront over 64 path slots, closest hit by 64-bit atomicMin, the reference's bounce (raytracing.cpp:263-275):
    visibility bits of the primary rays bit for bit, rows within 1e-4 -- with every sample flagged and with the need bits the
    (emulated) horizon pass produces.  Vertices the horizon pass finishes are skipped by the kernel and checked by the GPU tests."""
    pos, nrm, tri, h, osc = scene
    sel = np.arange(9, len(pos), 59)[:24]
    kw = dict(order=4, samples_u=16, samples_v=16, bounces=bounces, albedo=albedo)
    op = oracle.make_params(mode=oracle.INTERREFLECT, **kw)
    tab, bins = processing_table(oracle, op)
    S = len(tab)
    if use_horizon:
        hz, _ = _maps(hostcheck, h, pos[sel], nrm[sel])
        need = ~(tab[None, :, 2] > hz[:, bins])
    else:
        need = np.ones((len(sel), S), bool)
    need_words = np.ascontiguousarray(np.packbits(need, axis=1, bitorder="little")).view(np.uint32).copy()
    got, gvis, need_count = run_inter(hostcheck, h, pos[sel], nrm[sel], tab, 4, need_words, bounces, albedo, op.seed, vid_base=1000)
    ref, ovis, _ = oracle.bake_transfer(osc, pos[sel], nrm[sel], op, want_vis=True, vertex_id_base=1000)
    live = need_count > 0
    assert live.sum() >= len(sel) // 2
    assert np.array_equal(gvis[live], ovis[live])
    assert rel_l2(got[live], ref[live]).max() <= REL_L2_TOL
    # interreflection only adds energy to the DC term of the shadowed transfer
    sh, _, _ = oracle.bake_transfer(osc, pos[sel], nrm[sel], oracle.make_params(order=4, samples_u=16, samples_v=16))
    assert (got[live, 0] >= sh[live, 0] - 1e-6).all()


def test_wave_cull_variant_matches_oracle(oracle):
    """Experimental build variant PRT_WAVE_CULL=1 (candidate scan skips, warp-uniformly, the candidates whose elevation bound lies below
    the lowest ray of the round): results must not change.  CPU-validated here so that a later round only has to time it."""
    import conftest
    hc = conftest.load_hostcheck(["-DPRT_WAVE_CULL=1"])
    pos, nrm, tri = meshes.bumpy_torus(96, 64)
    h = hc.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    try:
        sel = np.arange(2, len(pos), 47)[:40]
        op = oracle.make_params(order=3, samples_u=16, samples_v=16)
        tab, bins = processing_table(oracle, op)
        hz, _ = _maps(hc, h, pos[sel], nrm[sel])
        need = ~(tab[None, :, 2] > hz[:, bins])
        need_words = np.ascontiguousarray(np.packbits(need, axis=1, bitorder="little")).view(np.uint32).copy()
        got, gvis = run_wave(hc, h, pos[sel], nrm[sel], tab, 3, need=need_words)
        full, fvis = run_wave(hc, h, pos[sel], nrm[sel], tab, 3)
    finally:
        hc.hc_free(h)
    ref, ovis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), pos[sel], nrm[sel], op, want_vis=True)
    assert np.array_equal(gvis, ovis) and np.array_equal(fvis, ovis)
    assert rel_l2(got, ref).max() <= REL_L2_TOL and rel_l2(full, ref).max() <= REL_L2_TOL


def test_wave_vertex_on_triangle_soup(hostcheck, oracle):
    """unstructured soup, arbitrary normals, ray origins on the geometry itself: deep entry chains, rays that start inside many
    boxes, hits at t ~ 0 -- the traversal pass must still agree with the oracle bit for bit (full and with need bits)"""
    rng = np.random.RandomState(12)
    k = 500
    c = rng.uniform(-1.5, 1.5, size=(k, 1, 3))
    pos = (c + rng.normal(size=(k, 3, 3)) * 0.3).reshape(-1, 3).astype(np.float32)
    tri = np.arange(3 * k, dtype=np.uint32).reshape(k, 3)
    nrm = rng.normal(size=(3 * k, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    sel = np.arange(1, 3 * k, 37)[:40]
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    try:
        op = oracle.make_params(order=3, samples_u=16, samples_v=16)
        tab, bins = processing_table(oracle, op)
        full, fvis = run_wave(hostcheck, h, pos[sel], nrm[sel], tab, 3)
        hz, _ = _maps(hostcheck, h, pos[sel], nrm[sel])
        need = ~(tab[None, :, 2] > hz[:, bins])
        need_words = np.ascontiguousarray(np.packbits(need, axis=1, bitorder="little")).view(np.uint32).copy()
        got, gvis = run_wave(hostcheck, h, pos[sel], nrm[sel], tab, 3, need=need_words)
    finally:
        hostcheck.hc_free(h)
    ref, ovis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), pos[sel], nrm[sel], op, want_vis=True)
    assert np.array_equal(fvis, ovis) and np.array_equal(gvis, ovis)
    assert rel_l2(full, ref).max() <= REL_L2_TOL and rel_l2(got, ref).max() <= REL_L2_TOL
    assert 0.2 < np.unpackbits(ovis.view(np.uint8)).mean() < 0.95


@pytest.mark.parametrize("order,su,sv", [(3, 32, 32), (4, 5, 7)])
def test_both_passes_as_the_kernels_run_them(hostcheck, oracle, order, su, sv):
    """horizon_vertex (horizon.cuh) writes need bits / counts and finishes the fully visible vertices; bake_wave_vertex traces the
    rest from those bits -- the bake exactly as the two kernels run it, on the emulator, against the oracle for EVERY vertex.  A
    smooth sphere next to the torus makes sure some vertices are finished by the first pass."""
    p0, n0, t0 = meshes.bumpy_torus(64, 48)
    p1, n1, t1 = meshes.icosphere(3)
    pos = np.concatenate([p0, p1 * 0.4 + np.float32([6.0, 0.0, 0.0])]).astype(np.float32)
    nrm = np.concatenate([n0, n1]).astype(np.float32)
    tri = np.concatenate([t0, t1 + len(p0)]).astype(np.uint32)
    sel = np.concatenate([np.arange(3, len(p0), 61)[:30], len(p0) + np.arange(0, len(p1), 13)[:20]])
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    op = oracle.make_params(order=order, samples_u=su, samples_v=sv)
    tab, _ = processing_table(oracle, op)
    S, n, words = len(tab), len(sel), (len(tab) + 31) // 32
    out = np.zeros((n, order * order), np.float32)
    vis = np.zeros((n, words), np.uint32)
    need_bits = np.zeros((n, words), np.uint32)
    need_count = np.zeros(n, np.uint32)
    p32, n32 = np.ascontiguousarray(pos[sel]), np.ascontiguousarray(nrm[sel])
    try:
        rc = hostcheck.hc_horizon_pass(h, p32.ctypes.data, n32.ctypes.data, n, tab.ctypes.data, S, order, 64, 30, 1e-4, 0, out.ctypes.data,
                                       vis.ctypes.data, need_bits.ctypes.data, need_count.ctypes.data)
        assert rc == 0
        live = need_count > 0
        assert live.any() and (~live).any(), "the scene is meant to exercise both ways a vertex can be finished"
        assert np.array_equal(need_count, np.unpackbits(need_bits.view(np.uint8), axis=1, bitorder="little")[:, :S].sum(axis=1))
        o2, v2 = run_wave(hostcheck, h, p32[live], n32[live], tab, order, need=np.ascontiguousarray(need_bits[live]))
        out[live], vis[live] = o2, v2
    finally:
        hostcheck.hc_free(h)
    ref, ovis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), pos[sel], nrm[sel], op, want_vis=True)
    assert np.array_equal(vis, ovis)
    assert rel_l2(out, ref).max() <= REL_L2_TOL


@pytest.mark.parametrize("scale", [0.5, 1.00008])
def test_non_unit_normals_through_both_passes(hostcheck, oracle, scale):
    """The reference hands assimp's normals to frame(N) un-normalised (model.cpp:27, raytracing.cpp:340): with |N| != 1 the frame is
    sheared against the sample table, so horizon_vertex must not build a map (all samples flagged) -- visibility bits stay exact.
    (Advisor finding of round 1: 1657 of 24576 bits differed with normals scaled by 0.5.)"""
    pos, nrm, tri = meshes.bumpy_torus(96, 64)
    sel = np.arange(7, len(pos), 251)[:24]
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    op = oracle.make_params(order=3, samples_u=32, samples_v=32)
    tab, _ = processing_table(oracle, op)
    S, n, words = len(tab), len(sel), (len(tab) + 31) // 32
    out, vis = np.zeros((n, 9), np.float32), np.zeros((n, words), np.uint32)
    need_bits, need_count = np.zeros((n, words), np.uint32), np.zeros(n, np.uint32)
    p32 = np.ascontiguousarray(pos[sel])
    n32 = np.ascontiguousarray(nrm[sel] * np.float32(scale))
    n32[::5] = nrm[sel][::5]                                         # unit and non-unit normals mixed
    try:
        assert hostcheck.hc_horizon_pass(h, p32.ctypes.data, n32.ctypes.data, n, tab.ctypes.data, S, 3, 64, 30, 1e-4, 0, out.ctypes.data,
                                         vis.ctypes.data, need_bits.ctypes.data, need_count.ctypes.data) == 0
        unit = np.zeros(n, bool); unit[::5] = True
        assert (need_count[~unit] == S).all() and (need_count[unit] < S).any()
        out, vis = run_wave(hostcheck, h, p32, n32, tab, 3, need=need_bits)
    finally:
        hostcheck.hc_free(h)
    ref, ovis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), p32, n32, op, want_vis=True)
    assert np.array_equal(vis, ovis)
    assert rel_l2(out, ref).max() <= REL_L2_TOL


def test_work_model_of_the_traversal_pass(hostcheck, oracle):
    """Regression guard on the WORK the traversal pass does (results are covered above): on a dense mesh with 1024 samples the item
    stacks must all but never overflow -- an overflowing push runs a whole per-ray stack traversal on a few lanes; on the GPU a
    node stack of 192 entries turned a 47 ms step into 66 ms (profiles/r2_slab_filter_rejected_ncu_summary.txt) -- and the steps must
    stay full."""
    import ctypes
    pos, nrm, tri = meshes.bumpy_torus(320, 320)
    order_ = meshes.morton_order(pos)
    sel = order_[:: len(order_) // 24][:24]
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    op = oracle.make_params(order=3, samples_u=32, samples_v=32)
    tab, bins = processing_table(oracle, op)
    hostcheck.hc_wave_step_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
    try:
        hz, _ = _maps(hostcheck, h, pos[sel], nrm[sel])
        need = ~(tab[None, :, 2] > hz[:, bins])
        keep = need.any(axis=1)
        words = np.ascontiguousarray(np.packbits(need, axis=1, bitorder="little")).view(np.uint32).copy()
        st = np.zeros(8, np.uint64)
        work = np.zeros(4, np.uint64)
        hostcheck.hc_wave_step_stats(st.ctypes.data, 1)
        got, vis = run_wave(hostcheck, h, pos[sel][keep], nrm[sel][keep], tab, 3, need=np.ascontiguousarray(words[keep]), work=work)
        hostcheck.hc_wave_step_stats(st.ctypes.data, 0)
    finally:
        hostcheck.hc_free(h)
    _, ovis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), pos[sel][keep], nrm[sel][keep], op, want_vis=True)
    assert np.array_equal(vis, ovis)
    n = int(keep.sum())
    leaf_steps, leaf_lanes, node_steps, node_lanes, scan_steps, scan_lanes, ovf_sub, ovf_leaf = (float(x) for x in st)
    assert node_steps > 20 * n
    assert ovf_sub <= 1.0 * n and ovf_leaf <= 1.0 * n, (ovf_sub / n, ovf_leaf / n)
    assert node_lanes / node_steps > 28 and leaf_lanes / leaf_steps > 28
    assert node_lanes == float(work[0]) or node_lanes >= float(work[0])              # items of occluded rays are dropped when popped


def test_fourth_slab_axis_only_removes_work(hostcheck, oracle):
    """The node test's fourth slab axis (Dop32) is a pure culling device: with it on or off the visibility words equal the oracle's bit
    for bit and the rows agree, and on the rays the horizon pass leaves to trace it removes a good share of the node visits and
    triangle tests (CPU work model of the bench mesh: -21 % / -20 %)."""
    pos, nrm, tri = meshes.bumpy_torus(320, 320)
    order_ = meshes.morton_order(pos)
    sel = order_[:: len(order_) // 20][:20]
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    op = oracle.make_params(order=3, samples_u=32, samples_v=32)
    tab, bins = processing_table(oracle, op)
    out = {}
    try:
        hz, _ = _maps(hostcheck, h, pos[sel], nrm[sel])
        need = ~(tab[None, :, 2] > hz[:, bins])
        keep = need.any(axis=1)
        words = np.ascontiguousarray(np.packbits(need, axis=1, bitorder="little")).view(np.uint32).copy()
        for dop in (1, 0):
            hostcheck.hc_wave_dop(dop)
            work = np.zeros(4, np.uint64)
            out[dop] = run_wave(hostcheck, h, pos[sel][keep], nrm[sel][keep], tab, 3, need=np.ascontiguousarray(words[keep]), work=work) + (work.copy(),)
    finally:
        hostcheck.hc_wave_dop(1)
        hostcheck.hc_free(h)
    ref, ovis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), pos[sel][keep], nrm[sel][keep], op, want_vis=True)
    for dop in (1, 0):
        assert np.array_equal(out[dop][1], ovis)
        assert rel_l2(out[dop][0], ref).max() <= REL_L2_TOL
    assert float(out[1][2][0]) < 0.92 * float(out[0][2][0])          # node visits
    assert float(out[1][2][1]) < 0.92 * float(out[0][2][1])          # triangle tests


@pytest.mark.parametrize("scene_kind", ["flat_room", "triangle_soup", "far_from_origin", "degenerate"])
def test_fourth_slab_axis_on_awkward_geometry(hostcheck, oracle, scene_kind):
    """The fourth slab axis of the node test and the oriented slabs of the horizon pass on geometry that is NOT a smooth bumpy surface:
    axis-aligned flat walls (slabs of zero thickness: the extent is floored), an unstructured triangle soup (mean normals mean nothing),
    a mesh far from the coordinate origin (large |M . o|: the float evaluation of the slab coordinate is what the padding is for) and
    zero-area / duplicated triangles (degenerate mean normals).  Both passes, emulated, against the oracle, bit for bit."""
    rng = np.random.RandomState(5)
    if scene_kind == "flat_room":
        # a closed box seen from inside with a floating slab in it
        def quad(a, b, c, d):
            return [a, b, c, a, c, d]
        L = 2.0
        corners = np.array([[x, y, z] for x in (-L, L) for y in (-L, L) for z in (-L, L)], np.float32)
        faces = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
        verts = []
        for f in faces:                                        # every wall tessellated 6 x 6
            p = corners[list(f)]
            for i in range(6):
                for j in range(6):
                    u0, u1, v0, v1 = i / 6, (i + 1) / 6, j / 6, (j + 1) / 6
                    bil = lambda u, v: (1 - u) * (1 - v) * p[0] + u * (1 - v) * p[1] + u * v * p[2] + (1 - u) * v * p[3]
                    verts += quad(bil(u0, v0), bil(u1, v0), bil(u1, v1), bil(u0, v1))
        verts += quad(np.float32([-1, -1, 0.3]), np.float32([1, -1, 0.3]), np.float32([1, 1, 0.3]), np.float32([-1, 1, 0.3]))
        pos = np.array(verts, np.float32)
        tri = np.arange(len(pos), dtype=np.uint32).reshape(-1, 3)
        org = np.float32([[0.1, 0.2, -1.9], [1.5, -0.3, 0.0], [0.0, 0.0, 0.31], [-1.99, 0.5, 0.5]])
        nrm = np.float32([[0, 0, 1], [-1, 0, 0], [0, 0, 1], [1, 0, 0]])
    elif scene_kind == "triangle_soup":
        k = 500
        c = rng.uniform(-2, 2, size=(k, 1, 3))
        pos = (c + rng.normal(size=(k, 3, 3)) * 0.3).reshape(-1, 3).astype(np.float32)
        tri = np.arange(3 * k, dtype=np.uint32).reshape(k, 3)
        org = rng.uniform(-1.5, 1.5, size=(8, 3)).astype(np.float32)
        nrm = rng.normal(size=(8, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    elif scene_kind == "far_from_origin":
        pos, n0, tri = meshes.bumpy_torus(64, 48)
        pos = (pos * 0.01 + np.float32([250.0, -180.0, 90.0])).astype(np.float32)
        sel = np.arange(7, len(pos), 211)[:10]
        org, nrm = pos[sel], n0[sel]
    else:
        pos, n0, tri = meshes.bumpy_torus(48, 32)
        extra = np.concatenate([tri[:50], np.stack([tri[:50, 0], tri[:50, 0], tri[:50, 1]], axis=1)])     # duplicates and zero-area triangles
        tri = np.concatenate([tri, extra]).astype(np.uint32)
        sel = np.arange(3, len(pos), 97)[:12]
        org, nrm = pos[sel], n0[sel]
    org, nrm = np.ascontiguousarray(org, np.float32), np.ascontiguousarray(nrm, np.float32)
    eps = 1e-4 if scene_kind != "far_from_origin" else 1e-4
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    op = oracle.make_params(order=3, samples_u=16, samples_v=32)
    tab, bins = processing_table(oracle, op)
    try:
        hz, _ = _maps(hostcheck, h, org, nrm, eps=eps)
        need = ~(tab[None, :, 2] > hz[:, bins])
        words = np.ascontiguousarray(np.packbits(need, axis=1, bitorder="little")).view(np.uint32).copy()
        ref, ovis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), org, nrm, oracle.make_params(order=3, samples_u=16, samples_v=32, origin_eps=eps), want_vis=True)
        visible = np.unpackbits(ovis.view(np.uint8), axis=1, bitorder="little")[:, :len(tab)].astype(bool)
        # (1) the horizon pass never frees an occluded ray (need bits are in processing order, the oracle's words in reference order)
        sref = (tab[:, 3].view(np.uint32) & 0xFFFFFF).astype(np.int64)
        assert not (~need & ~visible[:, sref]).any()
        # (2) the traversal pass with the fourth axis on, all rays and the flagged ones only
        for nb in (None, words):
            got, gvis = run_wave(hostcheck, h, org, nrm, tab, 3, need=nb)
            assert np.array_equal(gvis, ovis), scene_kind
            assert rel_l2(got, ref)[np.linalg.norm(ref, axis=1) > 1e-3].max(initial=0) <= REL_L2_TOL
    finally:
        hostcheck.hc_free(h)
