"""CPU tests: known answers for the oracle's preview tracer (oracle/raytrace.c; reference raytracing.cpp:162-222,280-317)."""
import numpy as np

import prt_b200
from prt_b200 import meshes
from test_oracle_probe import room


def test_normal_view_of_a_sphere(oracle):
    pos, nrm, tri = meshes.icosphere(3)
    sc = oracle.Scene(pos, tri)
    cam = prt_b200.Camera.look_at((0, 0, 4), (0, 0, 0), zoom_deg=45)
    acc, px = oracle.raytrace(sc, cam, 33, 33, mode=1, gamma=False)
    assert (acc[..., 3] == 1).all()
    c = acc[16, 16, :3]                                  # centre ray hits the pole facing the camera: Ng ~ (0,0,1)
    assert np.allclose(c, [0.5, 0.5, 1.0], atol=0.06) and tuple(px[16, 16]) == (int(255 * c[0]), int(255 * c[1]), int(255 * c[2]), 255)
    assert (acc[0, 0, :3] == 0).all() and tuple(px[0, 0]) == (0, 0, 0, 255)       # corner rays miss
    # row 0 is the bottom of the image (raytracing.cpp:297-299): lower half sees normals with y < 0
    assert acc[10, 16, 1] < 0.5 < acc[22, 16, 1] and acc[16, 10, 0] < 0.5 < acc[16, 22, 0]


def test_ao_known_answers(oracle):
    pos, nrm, tri = meshes.icosphere(3)
    sc = oracle.Scene(pos, tri)
    cam = prt_b200.Camera.look_at((0, 0, 4), (0, 0, 0))
    # convex object, white albedo: every bounce leaves the surface and escapes -> L = 1 wherever the sphere is hit, and where it is missed
    acc, px = oracle.raytrace(sc, cam, 24, 24, max_path_length=3, albedo=(1, 1, 1), gamma=False)
    assert np.allclose(acc[..., :3], 1.0)
    # albedo 0.5: hit pixels carry exactly 0.5 after one bounce, misses 1
    acc2, _ = oracle.raytrace(sc, cam, 24, 24, max_path_length=3, albedo=(0.5, 0.25, 1.0), gamma=False)
    hit = acc2[..., 0] < 1
    assert hit.sum() > 50 and np.allclose(acc2[hit][:, :3], [0.5, 0.25, 1.0]) and np.allclose(acc2[~hit][:, :3], 1.0)
    # max_path_length 1: the primary hit consumes the only segment -> black object
    acc3, _ = oracle.raytrace(sc, cam, 24, 24, max_path_length=1, gamma=False)
    assert np.array_equal(acc3[..., 0] == 0, hit)
    # inside a closed room nothing escapes
    rp, rt = room()
    acc4, px4 = oracle.raytrace(oracle.Scene(rp, rt), prt_b200.Camera.look_at((0, 0, 0), (1, 0.2, 0.1)), 16, 12, max_path_length=4)
    assert (acc4[..., :3] == 0).all() and (px4[..., :3] == 0).all() and acc4.shape == (12, 16, 4)
    # accumulation: second frame adds to the first; gamma 1/2.2 on the running mean
    accA, _ = oracle.raytrace(sc, cam, 24, 24, albedo=(0.5, 0.5, 0.5), frame=0)
    accB, pxB = oracle.raytrace(sc, cam, 24, 24, accum=accA, albedo=(0.5, 0.5, 0.5), frame=1)
    assert (accB[..., 3] == 2).all() and np.allclose(accB[hit][:, 0], 1.0)
    assert abs(int(pxB[12, 12, 0]) - int(255 * 0.5 ** (1 / 2.2))) <= 1
