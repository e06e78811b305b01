"""CPU tests: known answers for the oracle's probe capture / projection (SURVEY 8c KATs 3 and 4)."""
import numpy as np

from prt_b200 import meshes


def room(half=6.0):
    """closed box with INWARD facing triangles (a room, like the reference's data/cube.obj)."""
    c = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * half
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2]], np.uint32)
    return c, f      # (v1-v0)x(v2-v0) of these triangles points into the box


def test_probe_positions_and_direction_sets(oracle):
    pos = oracle.probe_positions([8, 8, 8], [12, 12, 12])           # app.cpp:50-51: 3.0 spacing over +-12
    assert pos.shape == (512, 3) and np.allclose(pos[0], -10.5) and np.allclose(pos[1], [-7.5, -10.5, -10.5]) and np.allclose(pos[-1], 10.5)
    d, w = oracle.fibonacci_dirs(100)                                # light_probe.cpp:137: get_dirs(100)
    assert np.allclose(np.linalg.norm(d, axis=1), 1, atol=1e-6) and abs(d[:, 2].mean()) < 1e-6 and d[0, 2] == 1 and d[-1, 2] == -1
    cd, cw = oracle.cube_dirs(64)
    assert cd.shape == (24576, 3) and abs(cw.sum() - 4 * np.pi) < 1e-3   # volume.cpp:264 "solid angle {} PI" expects 4


def test_closed_room_known_answers(oracle):
    """KAT 3: a probe inside a closed room sees 4 pi; KAT 4: constant radiance 1 -> SH_Irad = pi for every normal."""
    pos, tri = room()
    sc = oracle.Scene(pos, tri)
    probes = np.array([[0.3, -0.2, 0.1], [3.0, 2.0, -4.0]], np.float32)
    for dirs, w in (oracle.fibonacci_dirs(4096), oracle.cube_dirs(26)):
        pt = oracle.ProbeTransfer(sc, probes, dirs, w)
        rng, ids, tr, sf, keys = pt.download()
        assert pt.nnz == rng[-1, 1] and (np.diff(ids[rng[0, 0]:rng[0, 1]].astype(int)) > 0).all()
        for p in range(2):
            s = tr[rng[p, 0]:rng[p, 1]].sum(0)
            assert abs(s[0] - 3.54491) < 2e-3 and np.abs(s[1:]).max() < 3e-2
        # surfels lie on the walls, normals are axis aligned and point inward
        on_wall = np.isclose(np.abs(sf[:, :3]).max(1), 6.0, atol=1e-3)
        assert on_wall.all() and np.allclose(np.abs(sf[:, 3:]).max(1), 1.0, atol=1e-5)
        assert (np.sum(sf[:, :3] * sf[:, 3:], 1) < 0).all()
        out = pt.project(np.ones((pt.n_surfels, 4), np.float32))
        assert np.abs(out[:, 0:3, 3] - np.pi).max() < 5e-3           # A.w = c4 L00 - c5 L20 = pi
        assert np.abs(out[:, 0:3, 0:3]).max() < 5e-2 and np.allclose(out[:, 6, 3], 1.0)


def test_backface_and_sky_are_skipped(oracle):
    """outward-facing box seen from inside: every hit is a back face (volume.cpp:249); open scene: sky rays are skipped (:246)."""
    pos, tri = room()
    sc = oracle.Scene(pos, tri[:, ::-1].copy())
    d, w = oracle.fibonacci_dirs(512)
    pt = oracle.ProbeTransfer(sc, np.zeros((1, 3), np.float32), d, w)
    assert pt.nnz == 0 and pt.n_surfels == 0
    floor_only = oracle.Scene(pos, tri[:2])
    pt = oracle.ProbeTransfer(floor_only, np.zeros((1, 3), np.float32), d, w)
    rng, ids, tr, sf, keys = pt.download()
    assert 0 < tr[:, 0].sum() < 0.282095 * 2 * np.pi


def test_calculate_weight_known_answers(oracle):
    """SURVEY 8c KAT 5: empty-ish scene -> pure trilinear weights that sum to 1 (light_probe.cpp:315-316); a wall between the
    voxel and some probes zeroes exactly those weights and renormalises (:341-348)."""
    # a tiny far-away triangle: nothing is ever occluded, no ray hits -> score NaN -> no relocation
    far = np.array([[100, 100, 100], [101, 100, 100], [100, 101, 100]], np.float32)
    sc = oracle.Scene(far, np.array([[0, 1, 2]], np.uint32))
    w0, w1, score = oracle.volume_weights(sc, [2, 2, 2], [4, 4, 4], [2, 2, 2])
    w = np.concatenate([w0, w1], 1)
    assert np.isnan(score).all()
    interior = np.array([(z * 4 + y) * 4 + x for z in (1, 2) for y in (1, 2) for x in (1, 2)])
    assert np.allclose(w[interior].sum(1), 1.0, atol=1e-6)
    # voxel (1,1,1) of a 4^3 volume over 2^3 probes: fract = 0.25 on every axis
    f = 0.25
    expect = [(1 - f) * (1 - f) * f, f * (1 - f) * f, f * (1 - f) * (1 - f), (1 - f) ** 3, (1 - f) * f * (1 - f), (1 - f) * f * f, f ** 3, f * f * (1 - f)]
    assert np.allclose(w[(1 * 4 + 1) * 4 + 1], expect, atol=1e-6)
    # border voxels only see the probes that exist; weights still renormalise to 1
    assert np.allclose(w[0], [0, 0, 0, 0, 0, 0, 1, 0], atol=1e-6)
    # a big wall at x = 0 separates the x<0 voxels from the x>0 probes
    wall = np.array([[0, -50, -50], [0, 50, -50], [0, 50, 50], [0, -50, 50]], np.float32)
    sc2 = oracle.Scene(wall, np.array([[0, 1, 2], [0, 2, 3]], np.uint32))
    w0, w1, score = oracle.volume_weights(sc2, [2, 2, 2], [4, 4, 4], [2, 2, 2])
    w = np.concatenate([w0, w1], 1)
    # voxel (2,1,1) at x = +0.5 faces the wall's front side (score 0, stays put): the -x probes (corners 0,3,4,5) are hidden
    v = w[(1 * 4 + 1) * 4 + 2]
    assert score[(1 * 4 + 1) * 4 + 2] == 0.0
    assert np.allclose(v[[0, 3, 4, 5]], 0.0) and abs(v.sum() - 1.0) < 1e-6
    fx = 0.75
    e = np.array([(1 - fx) * (1 - f) * f, fx * (1 - f) * f, fx * (1 - f) * (1 - f), (1 - fx) * (1 - f) ** 2, (1 - fx) * f * (1 - f), (1 - fx) * f * f, fx * f * f, fx * f * (1 - f)])
    e[[0, 3, 4, 5]] = 0; e /= e.sum()
    assert np.allclose(v, e, atol=1e-6)
    # voxel (1,1,1) at x = -0.5 sees the wall's back side (score 1 > 0.2) and is moved to its least-inside neighbour across the
    # wall (light_probe.cpp:320-332), so it ends up with the +x probes as well
    u = w[(1 * 4 + 1) * 4 + 1]
    assert score[(1 * 4 + 1) * 4 + 1] == 1.0 and np.allclose(u[[0, 3, 4, 5]], 0.0) and abs(u.sum() - 1.0) < 1e-6
