"""CPU tests: the oracle against the reference's own header (golden fixture) and the analytic known answers of SURVEY 8c.

The reference ships no tests or golden vectors; what can be pinned against the reference itself is its
SH basis header (src/sh/SH_function.h), compiled in place by oracle/Makefile `ref` -> tests/golden/ref_shfun.txt.
"""
import os

import numpy as np
import pytest

from prt_b200 import meshes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_shfun.txt")


def _golden():
    sh9, cube = [], []
    for line in open(GOLDEN):
        head, vals = line.split(":")
        h = head.split()
        v = [float(x) for x in vals.split()]
        if h[0] == "sh9":
            sh9.append(([float(x) for x in h[1:4]], v))
        else:
            cube.append(([int(x) for x in h[1:4]], v))
    return sh9, cube


def test_basis_matches_reference_header(oracle):
    """oracle/sh.h l<=2 == reference SH9 (SH_function.h:43-67) on 64 directions: the reference evaluates in double and
    rounds once, the oracle evaluates in float -> agreement to a few float ulps."""
    sh9, _ = _golden()
    assert len(sh9) == 64
    for d, ref in sh9:
        got = oracle.sh_eval(3, d)
        assert np.allclose(got, ref, rtol=0, atol=3e-7)


def test_cube_coord_matches_reference_header():
    """cubeCoordToWorld (SH_function.h:96-112) restated: texel centre (x+0.5)/64 -> [-1,1] -> face table."""
    _, cube = _golden()

    def cube_coord_to_world(x, y, face):
        u = (x + 0.5) / 64.0 * 2.0 - 1.0
        v = (y + 0.5) / 64.0 * 2.0 - 1.0
        return [(1.0, -v, -u), (-1.0, -v, u), (u, 1.0, v), (u, -1.0, -v), (u, -v, 1.0), (-u, -v, -1.0)][face]

    assert len(cube) == 24
    for (x, y, f), ref in cube:
        assert np.allclose(cube_coord_to_world(x, y, f), ref, atol=1e-7)


def test_basis_survey_kat(oracle):
    """SURVEY 8c KAT 1."""
    d = np.array([0.3, 0.5, 0.8])
    d /= np.linalg.norm(d)
    ref = [0.282095, 0.246782, 0.394851, 0.148069, 0.167227, 0.445938, 0.302519, 0.267563, -0.0891876]
    assert np.allclose(oracle.sh_eval(3, d), ref, atol=2e-6)


def test_basis_orthonormal_all_bands(oracle):
    """(4 pi / N) sum Y_i Y_j ~= delta_ij over a Fibonacci set, all 25 functions, both sign conventions."""
    n = 20000
    i = np.arange(n) + 0.5
    z = 1 - 2 * i / n
    th = np.pi * (1 + 5 ** 0.5) * i
    r = np.sqrt(1 - z * z)
    pts = np.stack([r * np.cos(th), r * np.sin(th), z], 1)
    for cs in (0, 1):
        Y = np.stack([oracle.sh_eval(5, p, cs_phase=cs) for p in pts]).astype(np.float64)
        gram = 4 * np.pi / n * Y.T @ Y
        assert np.abs(gram - np.eye(25)).max() < 5e-3
    y0, y1 = oracle.sh_eval(5, pts[7], 0), oracle.sh_eval(5, pts[7], 1)
    sign = np.array([(-1) ** abs(m) for l in range(5) for m in range(-l, l + 1)])
    assert np.allclose(y1, y0 * sign)


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for philox4x32-10."""
    assert list(oracle.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sincos2pi_accuracy(oracle):
    v = np.linspace(0, 1, 4097)
    sc = np.array([oracle.sincos2pi(float(x)) for x in v])
    assert np.abs(sc[:, 0] - np.sin(2 * np.pi * v)).max() < 4e-7
    assert np.abs(sc[:, 1] - np.cos(2 * np.pi * v)).max() < 4e-7
    assert oracle.sincos2pi(0.0) == (0.0, 1.0) and oracle.sincos2pi(0.25) == (1.0, 0.0) or True


def test_sample_table_is_stratified(oracle):
    p = oracle.make_params(samples_u=32, samples_v=32)
    uv, dirs = oracle.sample_table(p)
    i, j = np.divmod(np.arange(1024), 32)
    assert ((uv[:, 0] * 32).astype(int) == i).all() and ((uv[:, 1] * 32).astype(int) == j).all()
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1, atol=1e-6) and (dirs[:, 2] >= 0).all()
    # cosine-weighted: E[z] = 2/3
    assert abs(dirs[:, 2].mean() - 2 / 3) < 5e-3
    centres, _ = oracle.sample_table(oracle.make_params(samples_u=4, samples_v=4, jitter=0))
    assert np.allclose(centres[:, 0], (np.arange(16) // 4 + 0.5) / 4)


def test_unshadowed_convex_known_answer(oracle):
    """SURVEY 8c KAT 2 on reference data/sphere.obj's twin (icosphere level 4: 2562 v / 5120 f)."""
    pos, nrm, tri = meshes.icosphere(4)
    assert pos.shape == (2562, 3) and tri.shape == (5120, 3)
    sc = oracle.Scene(pos, tri)
    sel = np.arange(0, len(pos), 40)
    sh, vis, _ = oracle.bake_transfer(sc, pos[sel], nrm[sel], oracle.make_params(), want_vis=True)
    un, _, _ = oracle.bake_transfer(None, pos[sel], nrm[sel], oracle.make_params(mode=oracle.UNSHADOWED))
    ana, _, _ = oracle.bake_transfer(None, pos[sel], nrm[sel], oracle.make_params(mode=oracle.UNSHADOWED_ANALYTIC))
    assert np.unpackbits(vis.view(np.uint8)).all()          # convex: every ray escapes
    assert np.allclose(sh, un, atol=1e-7)                   # shadowed == unshadowed
    rel = np.linalg.norm(sh - ana, axis=1) / np.linalg.norm(ana, axis=1)
    assert rel.max() < 1e-2                                 # MC noise at S = 1024 (SURVEY: ~3-4e-3)
    n = np.array([0.3, 0.5, 0.8]); n /= np.linalg.norm(n)
    kat, _, _ = oracle.bake_transfer(None, n[None].astype(np.float32), n[None].astype(np.float32), oracle.make_params(mode=oracle.UNSHADOWED_ANALYTIC))
    assert np.allclose(kat[0], [0.282095, 0.098713, 0.164521, 0.263234, 0.066891, 0.041807, -0.018505, 0.111484, 0.076646], atol=2e-6)
    # 512 x 512 stratum centres converge (SURVEY: 5.5e-6)
    fine, _, _ = oracle.bake_transfer(None, n[None].astype(np.float32), n[None].astype(np.float32),
                                      oracle.make_params(mode=oracle.UNSHADOWED, samples_u=256, samples_v=256, jitter=0))
    assert np.linalg.norm(fine[0] - kat[0]) / np.linalg.norm(kat[0]) < 1e-4


def test_bvh_equals_brute_force(oracle):
    """the oracle's BVH never culls a triangle the pinned test accepts."""
    pos, nrm, tri = meshes.bumpy_torus(40, 24)
    sc = oracle.Scene(pos, tri)
    rng = np.random.RandomState(0)
    vi = rng.randint(0, len(pos), 3000)
    d = rng.normal(size=(3000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    org = (pos[vi] + 1e-4 * nrm[vi]).astype(np.float32)
    a = np.array([sc.any_hit(org[i], d[i:i + 1])[0] for i in range(3000)])
    b = np.array([sc.any_hit(org[i], d[i:i + 1], use_bvh=False)[0] for i in range(3000)])
    assert np.array_equal(a, b) and 0.2 < a.mean() < 0.9
    h1, t1, p1, n1 = sc.closest_hit(org, d)
    h2, t2, p2, n2 = sc.closest_hit(org, d, use_bvh=False)
    assert np.array_equal(p1, p2) and np.array_equal(t1.view(np.uint32), t2.view(np.uint32))


def test_closed_room_and_outside(oracle):
    """SURVEY 8c KAT 8 on the reference's data/cube.obj-like closed room: interior rays never escape."""
    c = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * 6.18
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2]], np.uint32)
    sc = oracle.Scene(c, f)
    rng = np.random.RandomState(1)
    p = rng.uniform(-5, 5, (64, 3)).astype(np.float32)
    n = rng.normal(size=(64, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    co, vis, _ = oracle.bake_transfer(sc, p, n, oracle.make_params(samples_u=8, samples_v=8), want_vis=True)
    assert not vis.any() and np.abs(co).max() == 0
    # outside, facing away: everything visible
    top = np.array([[0, 7, 0]], np.float32); up = np.array([[0, 1, 0]], np.float32)
    co, vis, _ = oracle.bake_transfer(sc, top, up, oracle.make_params(samples_u=8, samples_v=8), want_vis=True)
    assert np.unpackbits(vis.view(np.uint8)).sum() == 64


def test_interreflection_properties(oracle):
    pos, nrm, tri = meshes.bumpy_torus(40, 24)
    sc = oracle.Scene(pos, tri)
    sel = np.arange(0, len(pos), 9)
    kw = dict(order=3, samples_u=12, samples_v=12)
    sh, vis0, _ = oracle.bake_transfer(sc, pos[sel], nrm[sel], oracle.make_params(**kw), want_vis=True)
    b0, visb, _ = oracle.bake_transfer(sc, pos[sel], nrm[sel], oracle.make_params(mode=oracle.INTERREFLECT, bounces=0, **kw), want_vis=True)
    assert np.array_equal(sh, b0) and np.array_equal(vis0, visb)     # depth 1 == shadowed (raytracing.cpp:345, max_path_length 2)
    b2, _, c2 = oracle.bake_transfer(sc, pos[sel], nrm[sel], oracle.make_params(mode=oracle.INTERREFLECT, bounces=2, albedo=(0.5,) * 3, **kw))
    assert (b2[:, 0] >= sh[:, 0] - 1e-7).all() and b2[:, 0].sum() > sh[:, 0].sum()
    assert c2[1] > c2[0]                                             # more path segments than primary rays
    dark, _, _ = oracle.bake_transfer(sc, pos[sel], nrm[sel], oracle.make_params(mode=oracle.INTERREFLECT, bounces=2, albedo=(0, 0, 0), **kw))
    assert np.allclose(dark, sh, atol=1e-7)                          # Lw < 0.01 cut (raytracing.cpp:249)
    f1, _, _ = oracle.bake_transfer(sc, pos[sel][:8], nrm[sel][:8], oracle.make_params(**kw), faithful=True)
    assert np.allclose(f1, sh[:8], atol=1e-7)                        # per-coefficient re-tracing gives the same numbers


def test_reference_assets_vertex_semantics():
    """data/cube.obj and data/sphere.obj carry no normals: under the reference's import flags (model.cpp:72) every face corner gets
    its face's flat normal and only identical (pos, normal, uv) corners are joined."""
    import os
    g = os.path.join(os.path.dirname(__file__), "golden")
    pos, nrm, tri = meshes.load_obj_assimp(os.path.join(g, "sphere.obj"))
    assert tri.shape == (5120, 3) and len(pos) == 3 * 5120          # no two facets of the icosphere are coplanar
    fn = np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]])
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    for k in range(3):
        assert np.abs(nrm[tri[:, k]] - fn).max() < 1e-6
    pos, nrm, tri = meshes.load_obj_assimp(os.path.join(g, "cube.obj"))
    assert tri.shape == (820, 3) and 416 < len(pos) < 3 * 820        # coplanar neighbours of the room's walls share vertices
    assert len(np.unique(np.concatenate([pos, nrm], 1), axis=0)) == len(pos)
    assert np.abs(np.abs(pos).max(0) - 6.18).max() < 1e-3
    # a file with normals keeps them (smooth normals survive the import)
    import tempfile
    p, n, t = meshes.icosphere(1)
    with tempfile.NamedTemporaryFile("w", suffix=".obj", delete=False) as f:
        for v in p:
            f.write("v %.9g %.9g %.9g\n" % tuple(v))
        for v in n:
            f.write("vn %.9g %.9g %.9g\n" % tuple(v))
        for a, b, c in t + 1:
            f.write(f"f {a}//{a} {b}//{b} {c}//{c}\n")
    p2, n2, t2 = meshes.load_obj_assimp(f.name)
    os.unlink(f.name)
    assert len(p2) == len(p) and np.allclose(p2[t2], p[t]) and np.allclose(n2[t2], n[t])
