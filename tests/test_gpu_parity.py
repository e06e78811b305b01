"""GPU parity tests (run on a B200 with `pytest -m gpu`): the sm_100a path, called through the C ABI, against the
CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): visibility bits and hit records bit-exact (pinned arithmetic),
SH coefficients <= 1e-4 relative L2 per vertex.
"""
import numpy as np
import pytest

from prt_b200 import meshes

pytestmark = pytest.mark.gpu

REL_L2_TOL = 1e-4  # north_star: "SH coefficients must match to <= 1e-4 relative L2 per vertex"


def rel_l2(a, b):
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-20)


@pytest.fixture(scope="module")
def torus():
    return meshes.bumpy_torus(96, 64)


@pytest.fixture(scope="module")
def torus_scenes(torus, prt, oracle):
    pos, nrm, tri = torus
    return prt.RTScene(pos, tri), oracle.Scene(pos, tri)


def _hemisphere_rays(pos, nrm, n, seed, eps=1e-4):
    rng = np.random.RandomState(seed)
    vi = rng.randint(0, len(pos), n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[np.sum(d * nrm[vi], 1) < 0] *= -1
    return (pos[vi] + np.float32(eps) * nrm[vi]).astype(np.float32), d.astype(np.float32)


def test_any_hit_bit_exact(torus, torus_scenes, prt):
    pos, nrm, _ = torus
    gs, os_ = torus_scenes
    org, d = _hemisphere_rays(pos, nrm, 30000, 1)
    got = gs.any_hit(prt.RTScene.pack_rays(org, d))
    ref = np.array([os_.any_hit(org[i], d[i:i + 1])[0] for i in range(len(d))])
    assert 0.05 < ref.mean() < 0.95
    assert np.array_equal(got.astype(np.int32), ref)


def test_closest_hit_bit_exact(torus, torus_scenes, prt):
    pos, nrm, _ = torus
    gs, os_ = torus_scenes
    org, d = _hemisphere_rays(pos, nrm, 20000, 2)
    t, prim, ng = gs.first_hit(prt.RTScene.pack_rays(org, d))
    hit, t2, prim2, ng2 = os_.closest_hit(org, d)
    assert np.array_equal(prim, prim2)
    assert np.array_equal(t.view(np.uint32), t2.view(np.uint32))
    assert np.array_equal(ng.view(np.uint32)[hit == 1], ng2.view(np.uint32)[hit == 1])


def test_segment_rays(torus, torus_scenes, prt):
    """unnormalised direction + tfar = 1 segment test (light_probe.cpp:250)."""
    pos, _, _ = torus
    gs, os_ = torus_scenes
    rng = np.random.RandomState(3)
    a = rng.uniform(-3.5, 3.5, (5000, 3)).astype(np.float32)
    b = rng.uniform(-3.5, 3.5, (5000, 3)).astype(np.float32)
    got = gs.any_hit(prt.RTScene.pack_rays(a, b - a, 0.0, 1.0))
    ref = np.array([os_.any_hit(a[i], (b - a)[i:i + 1], 0.0, 1.0)[0] for i in range(len(a))])
    assert np.array_equal(got.astype(np.int32), ref)


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_bake_shadowed_orders(torus, torus_scenes, prt, oracle, order):
    pos, nrm, _ = torus
    gs, os_ = torus_scenes
    sel = np.arange(0, len(pos), 7)[:600]
    gp = prt.BakeParams.make(order=order, samples_u=16, samples_v=16)
    op = oracle.make_params(order=order, samples_u=16, samples_v=16)
    got, gvis = prt.bake_transfer(gs, pos[sel], nrm[sel], gp, want_vis=True)
    ref, ovis, _ = oracle.bake_transfer(os_, pos[sel], nrm[sel], op, want_vis=True)
    assert np.array_equal(gvis, ovis), "per-ray visibility bits must agree exactly"
    assert rel_l2(got, ref).max() <= REL_L2_TOL


def test_bake_reference_defaults(torus, torus_scenes, prt, oracle):
    """order 3 (9 coeffs), 32x32 jittered strata: the reference's default bake (raytracing.cpp:320, app.h:70-71)."""
    pos, nrm, _ = torus
    gs, os_ = torus_scenes
    sel = np.arange(0, len(pos), 11)[:400]
    got, gvis = prt.bake_transfer(gs, pos[sel], nrm[sel], prt.BakeParams.make(), want_vis=True)
    ref, ovis, _ = oracle.bake_transfer(os_, pos[sel], nrm[sel], oracle.make_params(), want_vis=True)
    assert np.array_equal(gvis, ovis)
    assert rel_l2(got, ref).max() <= REL_L2_TOL
    frac = np.unpackbits(ovis.view(np.uint8)).mean()
    assert 0.3 < frac < 0.98  # the torus really is self-occluding


@pytest.mark.parametrize("knobs", [dict(refill_thresh=0), dict(refill_thresh=24), dict(refill_thresh=32), dict(horizon=0), dict(horizon=1, horizon_budget=0), dict(horizon=1, horizon_near=60, horizon_budget=4),
                                   dict(horizon_near=30, horizon_mid=0, horizon_slabs=0), dict(horizon_slabs=0), dict(wave_dop=0), dict(horizon_mid=40, horizon_gain=0, horizon_budget=256),
                                   dict(horizon_mid=5, horizon_gain=500), dict(work_list=0), dict(work_list=1), dict(pair_queue=0),
                                   dict(pair_queue=0, refill_thresh=0), dict(entry_list=0), dict(entry_list=0, refill_thresh=16)])
def test_bake_tuning_invariant(torus, torus_scenes, prt, oracle, knobs):
    """ray compaction, per-origin entry lists and the pair queues only reorganise work: results must not change."""
    pos, nrm, _ = torus
    gs, os_ = torus_scenes
    sel = np.arange(0, len(pos), 13)[:300]
    gs.ctx.set_tuning(**knobs)
    try:
        got, gvis = prt.bake_transfer(gs, pos[sel], nrm[sel], prt.BakeParams.make(samples_u=16, samples_v=16), want_vis=True)
    finally:
        gs.ctx.set_tuning(refill_thresh=8, entry_list=1, pair_queue=2, horizon=1, horizon_budget=64, work_list=-1, horizon_near=157, horizon_mid=24, horizon_gain=64, horizon_slabs=1, wave_dop=1)
    ref, ovis, _ = oracle.bake_transfer(os_, pos[sel], nrm[sel], oracle.make_params(samples_u=16, samples_v=16), want_vis=True)
    assert np.array_equal(gvis, ovis)
    assert rel_l2(got, ref).max() <= REL_L2_TOL


def test_work_list_order_independent(torus, torus_scenes, prt):
    """The sorted work list (heaviest vertices first) only changes which warp takes which vertex and when: rows and visibility words are bit-identical
    with it on or off (every vertex is reduced by one warp in a fixed order), also without visibility output."""
    pos, nrm, _ = torus
    gs, _ = torus_scenes
    gp = prt.BakeParams.make(order=3, samples_u=32, samples_v=32)
    try:
        gs.ctx.set_tuning(work_list=0, l2_prefetch=0)
        off, voff = prt.bake_transfer(gs, pos, nrm, gp, want_vis=True)
        assert gs.ctx.last_bake_stats().launches == 2
        gs.ctx.set_tuning(work_list=1)
        on, von = prt.bake_transfer(gs, pos, nrm, gp, want_vis=True)
        assert gs.ctx.last_bake_stats().launches == 5
        on_novis = prt.bake_transfer(gs, pos, nrm, gp)[0]
    finally:
        gs.ctx.set_tuning(work_list=-1, l2_prefetch=0)         # auto: on for small vertex counts (a shard of a multi-GPU bake)
    assert np.array_equal(von, voff)
    assert np.array_equal(on.view(np.uint32), off.view(np.uint32))
    assert np.array_equal(on.view(np.uint32), on_novis.view(np.uint32))


@pytest.mark.parametrize("bounces,albedo", [(1, (1.0, 1.0, 1.0)), (3, (0.5, 0.5, 0.5)), (8, (0.5, 0.3, 0.2))])
def test_bake_interreflect(torus, torus_scenes, prt, oracle, bounces, albedo):
    pos, nrm, _ = torus
    gs, os_ = torus_scenes
    sel = np.arange(0, len(pos), 17)[:250]
    kw = dict(order=4, samples_u=16, samples_v=16, bounces=bounces, albedo=albedo)
    got, gvis = prt.bake_transfer(gs, pos[sel], nrm[sel], prt.BakeParams.make(mode=prt.INTERREFLECT, **kw), want_vis=True, vertex_id_base=1000)
    ref, ovis, _ = oracle.bake_transfer(os_, pos[sel], nrm[sel], oracle.make_params(mode=oracle.INTERREFLECT, **kw), want_vis=True, vertex_id_base=1000)
    assert np.array_equal(gvis, ovis)
    assert rel_l2(got, ref).max() <= REL_L2_TOL
    # interreflection only adds energy to the DC term
    sh, _ = prt.bake_transfer(gs, pos[sel], nrm[sel], prt.BakeParams.make(order=4, samples_u=16, samples_v=16))
    assert (got[:, 0] >= sh[:, 0] - 1e-6).all()
    # the horizon pre-pass only skips primary rays that provably escape: same visibility bits, same rows up to summation order
    gs.ctx.set_tuning(horizon=0, l2_prefetch=0)
    try:
        got0, gvis0 = prt.bake_transfer(gs, pos[sel], nrm[sel], prt.BakeParams.make(mode=prt.INTERREFLECT, **kw), want_vis=True, vertex_id_base=1000)
    finally:
        gs.ctx.set_tuning(horizon=1, l2_prefetch=0)
    assert np.array_equal(gvis0, gvis) and rel_l2(got0, got).max() <= 1e-5
    assert gs.ctx.last_bake_stats().launches == 1


@pytest.mark.parametrize("su,sv,mode,bounces", [(128, 64, "shadowed", 0), (96, 96, "shadowed", 0), (64, 64, "interreflect", 2), (96, 96, "interreflect", 1)])
def test_bake_large_sample_counts(torus, torus_scenes, prt, oracle, su, sv, mode, bounces):
    """BASELINE configs 4/5 sample counts: 8192 = the wavefront kernels' occlusion-bitset limit (config 5), 9216 takes the
    per-ray fallback kernel, 4096-sample interreflection (config 4) runs the asynchronous wavefront."""
    pos, nrm, _ = torus
    gs, os_ = torus_scenes
    sel = np.arange(7, len(pos), 97)[:24]
    kw = dict(order=5 if mode == "shadowed" else 4, samples_u=su, samples_v=sv, bounces=bounces, albedo=(0.6, 0.5, 0.4))
    gm = prt.SHADOWED if mode == "shadowed" else prt.INTERREFLECT
    om = oracle.SHADOWED if mode == "shadowed" else oracle.INTERREFLECT
    got, gvis = prt.bake_transfer(gs, pos[sel], nrm[sel], prt.BakeParams.make(mode=gm, **kw), want_vis=True, vertex_id_base=5)
    ref, ovis, _ = oracle.bake_transfer(os_, pos[sel], nrm[sel], oracle.make_params(mode=om, **kw), want_vis=True, vertex_id_base=5)
    assert gvis.shape == ovis.shape == (len(sel), (su * sv + 31) // 32)
    assert np.array_equal(gvis, ovis)
    assert rel_l2(got, ref).max() <= REL_L2_TOL


def test_bake_unshadowed_modes(torus, prt, oracle):
    pos, nrm, _ = torus
    sel = np.arange(0, len(pos), 29)[:200]
    for mode_g, mode_o in [(prt.UNSHADOWED, oracle.UNSHADOWED), (prt.UNSHADOWED_ANALYTIC, oracle.UNSHADOWED_ANALYTIC)]:
        got, _ = prt.bake_transfer(None, pos[sel], nrm[sel], prt.BakeParams.make(order=5, mode=mode_g))
        ref, _, _ = oracle.bake_transfer(None, pos[sel], nrm[sel], oracle.make_params(order=5, mode=mode_o))
        assert rel_l2(got, ref).max() <= REL_L2_TOL


def test_cs_phase_and_centres(torus, torus_scenes, prt, oracle):
    pos, nrm, _ = torus
    gs, os_ = torus_scenes
    sel = np.arange(0, len(pos), 31)[:150]
    kw = dict(order=5, samples_u=8, samples_v=32, cs_phase=1, jitter=0)
    got, gvis = prt.bake_transfer(gs, pos[sel], nrm[sel], prt.BakeParams.make(**kw), want_vis=True)
    ref, ovis, _ = oracle.bake_transfer(os_, pos[sel], nrm[sel], oracle.make_params(**kw), want_vis=True)
    assert np.array_equal(gvis, ovis)
    assert rel_l2(got, ref).max() <= REL_L2_TOL


def test_convex_sphere_known_answer(prt):
    """SURVEY 8c KAT 2: on a convex mesh shadowed == unshadowed -> (A_l/pi) Y_lm(perm(n)); MC tolerance 1e-2."""
    pos, nrm, tri = meshes.icosphere(4)
    sc = prt.RTScene(pos, tri)
    sh, vis = prt.bake_transfer(sc, pos, nrm, prt.BakeParams.make(), want_vis=True)
    un, _ = prt.bake_transfer(None, pos, nrm, prt.BakeParams.make(mode=prt.UNSHADOWED))
    ana, _ = prt.bake_transfer(None, pos, nrm, prt.BakeParams.make(mode=prt.UNSHADOWED_ANALYTIC))
    assert np.unpackbits(vis.view(np.uint8)).all()
    assert rel_l2(sh, un).max() <= 1e-6
    assert rel_l2(sh, ana).max() <= 1e-2


def test_closed_room_all_occluded(prt):
    """SURVEY 8c KAT 8: rays from inside a closed box never escape."""
    c = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * 3
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2]], np.uint32)
    sc = prt.RTScene(c, f)
    rng = np.random.RandomState(5)
    p = rng.uniform(-2.5, 2.5, (500, 3)).astype(np.float32)
    n = rng.normal(size=(500, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    co, vis = prt.bake_transfer(sc, p, n, prt.BakeParams.make(samples_u=8, samples_v=8), want_vis=True)
    assert not vis.any()
    assert np.abs(co).max() == 0.0


def test_edge_cases(prt):
    pos, nrm, tri = meshes.icosphere(1)
    sc = prt.RTScene(pos, tri)
    # empty vertex range
    co, _ = prt.bake_transfer(sc, pos[:0], nrm[:0], prt.BakeParams.make())
    assert co.shape == (0, 9)
    # single vertex, single sample
    co, vis = prt.bake_transfer(sc, pos[:1], nrm[:1], prt.BakeParams.make(samples_u=1, samples_v=1), want_vis=True)
    assert co.shape == (1, 9) and vis.shape == (1, 1)
    # ragged sample count (not a multiple of 32) and degenerate triangles in the scene
    tri2 = np.concatenate([tri, np.array([[0, 0, 1], [2, 2, 2]], np.uint32)])
    sc2 = prt.RTScene(pos, tri2)
    a, va = prt.bake_transfer(sc2, pos, nrm, prt.BakeParams.make(samples_u=5, samples_v=7), want_vis=True)
    b, vb = prt.bake_transfer(sc, pos, nrm, prt.BakeParams.make(samples_u=5, samples_v=7), want_vis=True)
    assert np.array_equal(va, vb) and np.allclose(a, b, atol=1e-7)
    # bad arguments fail loudly
    with pytest.raises(prt.PRTError):
        prt.bake_transfer(sc, pos, nrm, prt.BakeParams.make(order=6))
    with pytest.raises(prt.PRTError):
        prt.bake_transfer(None, pos, nrm, prt.BakeParams.make())
    with pytest.raises(prt.PRTError):
        prt.RTScene(pos, np.array([[0, 1, 9999]], np.uint32))


def test_bake_SH_mesh_vert_layout(prt, oracle):
    """bake_SH(Mesh&): interleaved 60-byte Mesh::Vert in, sh_coeff[9] written in place (gl.h:76-80)."""
    pos, nrm, tri = meshes.bumpy_torus(48, 32)
    verts = np.zeros((len(pos), 15), np.float32)
    verts[:, 0:3], verts[:, 3:6] = pos, nrm
    out = prt.bake_SH(verts, tri)
    assert np.array_equal(out, verts[:, 6:15])
    ref, _, _ = oracle.bake_transfer(oracle.Scene(pos, tri), pos, nrm, oracle.make_params(cs_phase=1))   # sh::EvalSH's sign
    assert rel_l2(verts[:, 6:15], ref).max() <= REL_L2_TOL
    # the sign-free convention of SH_function.h / SH.glsl differs exactly by (-1)^m
    out0 = prt.bake_SH(verts.copy(), tri, prt.BakeParams.make(cs_phase=0))
    assert np.array_equal(out0 * np.array([1, -1, 1, -1, 1, -1, 1, -1, 1], np.float32), out)


def test_full_size_properties(prt):
    """BASELINE config 1 size (buddha-scale torus, S = 1024): size-independent properties instead of the oracle."""
    pos, nrm, tri = meshes.bumpy_torus(737, 737)
    sc = prt.RTScene(pos, tri)
    p = prt.BakeParams.make()
    co, vis = prt.bake_transfer(sc, pos, nrm, p, want_vis=True)
    assert np.isfinite(co).all()
    frac = np.unpackbits(vis.view(np.uint8)).reshape(len(pos), -1).mean(1)
    # DC coefficient is exactly 0.282095 * visible fraction (cosine-weighted sampling: every sample weighs 1/S)
    assert np.abs(co[:, 0] - 0.282095 * frac).max() < 2e-6
    # shadowed <= unshadowed in the DC term, and idempotence of a second run
    co2, vis2 = prt.bake_transfer(sc, pos, nrm, p, want_vis=True)
    assert np.array_equal(vis, vis2)
    assert rel_l2(co, co2).max() < 1e-5
    # sharded bake == whole bake (vertex ranges are independent)
    half = len(pos) // 2
    a, _ = prt.bake_transfer(sc, pos[:half], nrm[:half], p)
    assert rel_l2(a, co[:half]).max() < 1e-5


def _soup(n_tri, seed, scale=1.0, offset=(0, 0, 0), big=False):
    rs = np.random.RandomState(seed)
    c = rs.uniform(-1, 1, (n_tri, 1, 3))
    size = rs.uniform(0.02, 0.6 if big else 0.15, (n_tri, 1, 1))
    v = (c + size * rs.normal(size=(n_tri, 3, 3))).reshape(-1, 3)
    pos = (v * scale + np.asarray(offset)).astype(np.float32)
    tri = np.arange(3 * n_tri, dtype=np.uint32).reshape(-1, 3)
    return pos, tri


@pytest.mark.parametrize("case", ["soup", "soup_big", "soup_offset", "soup_tiny", "room_inside", "sphere_shell"])
def test_horizon_map_is_conservative_on_adversarial_scenes(prt, oracle, case):
    """The horizon pass may only skip rays that provably hit nothing.  Random triangle soups with origins anywhere (not on
    a surface), big overlapping triangles, large coordinate offsets, tiny scales, a closed room and a sphere shell seen from
    inside: visibility words must still equal the oracle's bit for bit, with every tuning of the horizon builder."""
    rs = np.random.RandomState(11)
    if case == "soup":
        pos, tri = _soup(3000, 1)
    elif case == "soup_big":
        pos, tri = _soup(800, 2, big=True)
    elif case == "soup_offset":
        pos, tri = _soup(2000, 3, scale=5.0, offset=(300.0, -150.0, 80.0))
    elif case == "soup_tiny":
        pos, tri = _soup(2000, 4, scale=1e-3)
    elif case == "room_inside":
        from test_oracle_probe import room
        pos, tri = room(2.0)
        extra, et = _soup(300, 5, scale=1.5)
        pos, tri = np.concatenate([pos, extra]).astype(np.float32), np.concatenate([tri, et + np.uint32(len(pos))]).astype(np.uint32)
    else:
        p, n, t = meshes.icosphere(4)
        pos, tri = (p * 2).astype(np.float32), t[:, ::-1].copy()
    lo, hi = pos.min(0), pos.max(0)
    n_org = 160
    org = rs.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo), (n_org, 3)).astype(np.float32)
    # a third of the origins sit exactly on triangle vertices (the bake's real use), with random normals
    org[: n_org // 3] = pos[rs.randint(0, len(pos), n_org // 3)]
    nrm = rs.normal(size=(n_org, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[:8] = np.eye(3, dtype=np.float32)[[0, 1, 2, 0, 1, 2, 0, 1]] * np.array([1, 1, 1, -1, -1, -1, 1, -1], np.float32)[:, None]
    eps = 1e-4 * (1e-3 if case == "soup_tiny" else 1.0)
    gs, os_ = prt.RTScene(pos, tri), oracle.Scene(pos, tri)
    kw = dict(samples_u=16, samples_v=32, origin_eps=eps)
    ref, ovis, _ = oracle.bake_transfer(os_, org, nrm, oracle.make_params(**kw), want_vis=True)
    frac = np.unpackbits(ovis.view(np.uint8)).mean()
    assert 0.02 < frac < 0.99 or case in ("room_inside", "sphere_shell")
    for knobs in (dict(), dict(horizon=0), dict(horizon_budget=0), dict(horizon_budget=3, horizon_near=80), dict(horizon_budget=200, horizon_near=10),
                  dict(horizon_near=30, horizon_mid=0, horizon_slabs=0), dict(horizon_mid=40, horizon_gain=0, horizon_budget=256), dict(horizon_slabs=0), dict(wave_dop=0)):
        gs.ctx.set_tuning(**knobs)
        try:
            got, gvis = prt.bake_transfer(gs, org, nrm, prt.BakeParams.make(**kw), want_vis=True)
        finally:
            gs.ctx.set_tuning(horizon=1, horizon_budget=64, horizon_near=157, horizon_mid=24, horizon_gain=64, horizon_slabs=1, wave_dop=1)
        assert np.array_equal(gvis, ovis), f"{case} {knobs}: {np.count_nonzero(gvis != ovis)} visibility words differ"
        assert rel_l2(got, ref)[np.linalg.norm(ref, axis=1) > 1e-3].max(initial=0) <= REL_L2_TOL
