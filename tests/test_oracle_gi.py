"""CPU tests: known answers for the oracle's per-frame probe pipeline (oracle/gi.c; reference relight.comp, transfer2volume.comp,
Paral_Shadow).  The reference holds no golden output for these GLSL passes (parity unpinned), so the anchors are analytic."""
import numpy as np

import prt_b200
from test_oracle_probe import room


def test_paral_shadow_matrix_is_glm_ortho_lookat(oracle):
    d, m = oracle.paral_shadow_matrix(0.17, 0.84)                         # app.h:67 defaults
    M = m.reshape(4, 4).T                                                  # column-major -> rows
    th, ph = np.pi * 0.17, 2 * np.pi * 0.84
    assert np.allclose(d, [np.sin(th) * np.sin(ph), np.cos(th), np.sin(th) * np.cos(ph)], atol=1e-6)
    assert np.allclose(M[3], [0, 0, 0, 1])
    eye = np.append(30 * d, 1)                                             # the light's eye maps to ndc z = -(f+n)/(f-n) (view z = 0)
    assert np.allclose(M @ eye, [0, 0, -60.1 / 59.9, 1], atol=1e-5)
    origin = M @ np.array([0, 0, 0, 1.0])                                  # the look-at point: centre of the frustum, 30 along the axis
    assert np.allclose(origin[:2], 0, atol=1e-6) and abs(origin[2] - (2 * 30 / 59.9 - 60.1 / 59.9)) < 1e-6
    R = M[:3, :3] * np.array([[30], [30], [-59.9 / 2]])                    # undo the ortho scale -> rotation
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-5)
    d2, m2 = oracle.paral_shadow_matrix(0.05, 0.3)                         # up < 0.1 switches the up vector (gl.cpp:628)
    assert np.isfinite(m2).all()


def test_shadow_map_depth_of_a_floor(oracle):
    """light straight down (up = 0 -> direction +y): a floor at y = -6 is 36 from the eye plane -> depth (36-0.1)/59.9."""
    pos, tri = room()
    sc = oracle.Scene(pos, tri)
    d, m = oracle.paral_shadow_matrix(0.0, 0.0)
    assert np.allclose(d, [0, 1, 0], atol=1e-6)
    dm = oracle.shadow_map(sc, m, 32)
    # texels whose ray passes through the room (|x|,|z| < 6 of +-30) see the ceiling at distance 24; all others see nothing
    inside = np.abs((np.arange(32) + 0.5) / 32 * 60 - 30) < 6
    mask = np.outer(inside, inside)
    assert (dm[~mask] == 1.0).all()
    # rays exactly on the shared diagonal of the two ceiling triangles may slip through (Moeller-Trumbore is not watertight)
    # and then see the floor at 36; everywhere else the ceiling
    ij = np.arange(32)
    diag = (ij[:, None] == ij[None, :]) | (ij[:, None] == 31 - ij[None, :])
    assert np.allclose(dm[mask & ~diag], (24 - 0.1) / 59.9, atol=1e-5)
    assert np.all(np.isclose(dm[mask & diag], (24 - 0.1) / 59.9, atol=1e-5) | np.isclose(dm[mask & diag], (36 - 0.1) / 59.9, atol=1e-5))


def test_relight_known_answers(oracle):
    d, m = oracle.paral_shadow_matrix(0.0, 0.0)
    sf = np.array([[0, -6, 0, 0, 1, 0],          # floor, facing the sky light
                   [0, 6, 0, 0, -1, 0],          # ceiling, facing away
                   [5.95, 0, 0, -1, 0, 0]], np.float32)     # coloured wall (x > 5.9), lit by the spot light only
    P = prt_b200.RelightParams.make(d, m, sky_intensity=(5, 4, 3), cast_intensity=100.0, cast_position=(4, 0, 0), multi_bounce=False, temp_weight=1.0)
    rad = oracle.relight(P, sf, np.zeros((3, 4), np.float32))
    assert np.allclose(rad[0, :3], 0.4 * np.array([5, 4, 3]), rtol=1e-6)            # albedo 0.4 * sky * cos 0, no spot (outside the cone)
    assert np.allclose(rad[1, :3], 0)
    dist = 1.95
    sel = (int(0 / 6 + 100) + int(0 / 6 + 100)) % 3
    alb = np.full(3, 0.1, np.float32); alb[sel] += 0.7
    assert np.allclose(rad[2, :3], alb * 100.0 / dist ** 2, rtol=1e-5)               # theta = 1 -> soft edge 1, cos 1
    assert (rad[:, 3] == 1).all()
    # temporal blend: relight.comp:80
    P.temp_weight = 0.1
    rad2 = oracle.relight(P, sf, rad)
    assert np.allclose(rad2, rad, rtol=1e-6)
    rad3 = oracle.relight(P, sf, np.zeros((3, 4), np.float32))
    assert np.allclose(rad3[:, :3], 0.1 * rad[:, :3], rtol=1e-6)
    # shadow map: a blocker above the floor surfel
    dm = np.full((8, 8), 0.2, np.float32)
    assert np.allclose(oracle.relight(P, sf, np.zeros((3, 4), np.float32), depth=dm)[0, :3], 0)
    assert np.allclose(oracle.relight(P, sf, np.zeros((3, 4), np.float32), depth=np.ones((8, 8), np.float32))[0, :3], rad3[0, :3])
    # SH feedback: a volume holding constant irradiance E (only Ar.w etc. = c4*L00 - c5*L20 -> x.w term) adds albedo*atten*E/pi
    vol = np.zeros((2 * 2 * 2, 7, 4), np.float32); vol[:, 0:3, 3] = [2.0, 3.0, 4.0]
    P.multi_bounce = 1; P.temp_weight = 1.0; P.atten = 0.5
    fb = oracle.relight(P, sf[:1], np.zeros((1, 4), np.float32), volumes=vol, volume_res=[2, 2, 2], scene_size=[6, 6, 6])
    assert np.allclose(fb[0, :3], 0.4 * np.array([5, 4, 3]) + 0.4 * 0.5 * np.array([2, 3, 4]) / np.float32(3.14159265359), rtol=1e-6)


def test_transfer_to_volume_partition_of_unity(oracle):
    """constant probe SH + weights that sum to one -> the same constant in every voxel; one-hot weights select the right corner."""
    pr, vr = [3, 2, 2], [6, 4, 4]
    psh = np.tile(np.arange(28, dtype=np.float32).reshape(1, 7, 4), (12, 1, 1))
    rs = np.random.RandomState(1)
    w = rs.rand(96, 8).astype(np.float32)
    # corners outside the probe grid contribute nothing: give them zero weight like calculate_weight does
    out = oracle.transfer_to_volume(psh, pr, w[:, :4], w[:, 4:], vr)
    off = np.array([[0, 0, 1], [1, 0, 1], [1, 0, 0], [0, 0, 0], [0, 1, 0], [0, 1, 1], [1, 1, 1], [1, 1, 0]])
    for v in range(96):
        x, y, z = v % 6, (v // 6) % 4, v // 24
        anchor = np.floor((np.array([x, y, z]) + 0.5) / np.array(vr) * np.array(pr) - 0.5).astype(int)
        ok = [(0 <= anchor + o).all() and (anchor + o < pr).all() for o in off]
        assert np.allclose(out[v], psh[0] * w[v][ok].sum(), rtol=1e-5)
    # distinct probes, one-hot weight on corner 6 (+1,+1,+1)
    psh2 = np.arange(12, dtype=np.float32).reshape(12, 1, 1) * np.ones((1, 7, 4), np.float32)
    w1 = np.zeros((96, 8), np.float32); w1[:, 6] = 1
    out2 = oracle.transfer_to_volume(psh2, pr, w1[:, :4], w1[:, 4:], vr)
    v = (1 * 4 + 1) * 6 + 2                                                # voxel (2,1,1): anchor = floor((2.5/6*3-0.5, 1.5/4*2-0.5, .)) = (0,0,0)
    assert np.allclose(out2[v], (1 * 2 + 1) * 3 + 1)                       # probe (1,1,1)
