"""CPU tests of the product's host code: the 8-wide compressed BVH builder (prt_b200/csrc/bvh_build.cpp) and the traversal
header compiled as plain C++ (tests/hostcheck, test tooling) must give the oracle's answers bit for bit."""
import ctypes as C

import numpy as np
import pytest

from prt_b200 import meshes


def _rays(pos, nrm, n, seed):
    rng = np.random.RandomState(seed)
    vi = rng.randint(0, len(pos), n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[np.sum(d * nrm[vi], 1) < 0] *= -1
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = pos[vi] + np.float32(1e-4) * nrm[vi]
    rays[:, 4:7] = d
    rays[:, 7] = np.inf
    return rays


@pytest.mark.parametrize("mesh", ["torus", "sphere", "tiny"])
def test_bvh8_traversal_matches_oracle(hostcheck, oracle, mesh):
    if mesh == "torus":
        pos, nrm, tri = meshes.bumpy_torus(72, 48)
    elif mesh == "sphere":
        pos, nrm, tri = meshes.icosphere(3)
    else:
        pos, nrm, tri = meshes.icosphere(0)
        tri = tri[:2]            # two triangles: root node is a leaf-only node
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    info = np.zeros(3, np.uint32)
    hostcheck.hc_info(h, info.ctypes.data)
    assert info[1] == len(tri) and info[2] <= 46
    sc = oracle.Scene(pos, tri)
    rays = _rays(pos, nrm, 6000, 3)
    # also rays with exactly zero direction components and segment rays
    rays[:50, 4] = 0.0
    rays[50:100, 5] = -0.0
    rays[100:200, 7] = 0.7
    got = np.zeros(len(rays), np.uint8)
    hostcheck.hc_any_hit(h, rays.ctypes.data, len(rays), got.ctypes.data)
    ref = np.array([sc.any_hit(rays[i, 0:3], rays[i:i + 1, 4:7], 0.0, float(rays[i, 7]))[0] for i in range(len(rays))])
    assert np.array_equal(got.astype(np.int32), ref)
    t = np.zeros(len(rays), np.float32); prim = np.zeros(len(rays), np.uint32); ng = np.zeros((len(rays), 3), np.float32)
    hostcheck.hc_closest_hit(h, rays.ctypes.data, len(rays), t.ctypes.data, prim.ctypes.data, ng.ctypes.data)
    sel = slice(200, None)
    hit, t2, prim2, ng2 = sc.closest_hit(rays[sel, 0:3], rays[sel, 4:7])
    assert np.array_equal(prim[sel], prim2) and np.array_equal(t[sel].view(np.uint32), t2.view(np.uint32))
    assert np.array_equal(ng[sel].view(np.uint32), ng2.view(np.uint32))
    hostcheck.hc_free(h)


def test_bvh8_rejects_bad_input(hostcheck):
    pos = np.zeros((3, 3), np.float32)
    tri = np.array([[0, 1, 7]], np.uint32)
    assert not hostcheck.hc_build(pos.ctypes.data, 12, 3, tri.ctypes.data, 1)
    pos[1, 0] = np.nan
    tri = np.array([[0, 1, 2]], np.uint32)
    assert not hostcheck.hc_build(pos.ctypes.data, 12, 3, tri.ctypes.data, 1)
