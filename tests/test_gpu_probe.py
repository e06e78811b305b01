"""GPU parity tests for BASELINE config 3 (per-probe SH radiance-transfer capture + projection) against the oracle."""
import numpy as np
import pytest

from prt_b200 import meshes
from test_oracle_probe import room

pytestmark = pytest.mark.gpu


def scene_with_occluder():
    """room (+-6) with a bumpy torus inside, like the reference's cube + buddha scene."""
    rp, rt = room()
    tp, _, tt = meshes.bumpy_torus(48, 32)
    pos = np.concatenate([rp, tp * np.float32(1.2)]).astype(np.float32)
    tri = np.concatenate([rt, tt + np.uint32(len(rp))]).astype(np.uint32)
    return pos, tri


@pytest.mark.parametrize("dirset", ["fibonacci", "cube"])
def test_probe_capture_matches_oracle(prt, oracle, dirset):
    pos, tri = scene_with_occluder()
    gs, os_ = prt.RTScene(pos, tri), oracle.Scene(pos, tri)
    probes = prt.probe_positions([4, 3, 2], [6, 6, 6])
    assert np.array_equal(probes, oracle.probe_positions([4, 3, 2], [6, 6, 6]))
    if dirset == "fibonacci":
        d, w = prt.fibonacci_dirs(1500)
        d2, w2 = oracle.fibonacci_dirs(1500)
    else:
        d, w = prt.cube_dirs(16)
        d2, w2 = oracle.cube_dirs(16)
    assert np.array_equal(d, d2) and np.array_equal(w, w2)
    g, o = prt.ProbeTransfer(gs, probes, d, w), oracle.ProbeTransfer(os_, probes, d, w)
    assert (g.nnz, g.n_surfels) == (o.nnz, o.n_surfels) and g.nnz > 1000
    gr, gi, gt, gsf, gk = g.download()
    orr, oi, ot, osf, ok = o.download()
    assert np.array_equal(gr, orr) and np.array_equal(gi, oi) and np.array_equal(gk, ok)     # structure is exact
    assert np.abs(gt - ot).max() <= 1e-5                                                        # SH9 * dOmega sums
    assert np.abs(gsf - osf).max() <= 1e-4                                                      # surfel means
    rs = np.random.RandomState(0)
    rad = rs.rand(g.n_surfels, 4).astype(np.float32)
    assert np.abs(g.project(rad) - o.project(rad)).max() <= 1e-4


def test_probe_known_answers(prt):
    """SURVEY 8c KATs 3+4 on the GPU: closed room -> sum dOmega = 4 pi; constant radiance -> SH_Irad = pi."""
    pos, tri = room()
    sc = prt.RTScene(pos, tri)
    d, w = prt.fibonacci_dirs(4096)
    g = prt.ProbeTransfer(sc, prt.probe_positions([3, 3, 3], [6, 6, 6]), d, w)
    rng, ids, tr, sf, keys = g.download()
    for p in range(27):
        s = tr[rng[p, 0]:rng[p, 1]].sum(0)
        assert abs(s[0] - 3.54491) < 2e-3 and np.abs(s[1:]).max() < 3e-2
        assert (np.diff(ids[rng[p, 0]:rng[p, 1]].astype(np.int64)) > 0).all()
    out = g.project(np.ones((g.n_surfels, 4), np.float32))
    assert np.abs(out[:, 0:3, 3] - np.pi).max() < 5e-3
    # empty result: outward facing box -> everything is a back face
    g2 = prt.ProbeTransfer(prt.RTScene(pos, tri[:, ::-1].copy()), np.zeros((2, 3), np.float32), d, w)
    assert g2.nnz == 0 and g2.n_surfels == 0
    with pytest.raises(prt.PRTError):
        prt.ProbeTransfer(sc, np.zeros((1, 3), np.float32), *prt.fibonacci_dirs(5000))


def test_probe_grid_scale_properties(prt):
    """a 16^3 grid (4096 probes x 4096 rays = 1.7e7 rays): every probe inside the room sees 4 pi minus nothing."""
    pos, tri = scene_with_occluder()
    sc = prt.RTScene(pos, tri)
    d, w = prt.fibonacci_dirs(4096)
    probes = prt.probe_positions([16, 16, 16], [6, 6, 6])
    g = prt.ProbeTransfer(sc, probes, d, w)
    rng, ids, tr, sf, keys = g.download()
    assert rng[0, 0] == 0 and (rng[1:, 0] == rng[:-1, 1]).all() and rng[-1, 1] == g.nnz      # probe-major, contiguous
    dc = np.add.reduceat(tr[:, 0], rng[:, 0].astype(np.int64)) if g.nnz else np.zeros(len(probes))
    # probes inside the torus tube can see less (back faces are skipped); nobody sees more than 4 pi
    assert dc.max() < 3.54491 * 1.001 and np.median(dc) > 3.5
    assert ids.max() == g.n_surfels - 1 and (np.diff(keys.astype(np.int64)) > 0).all()


def test_calculate_weight_matches_oracle(prt, oracle):
    """calculate_weight (light_probe.cpp:156-367): scores and weights against the oracle on a room + occluder scene."""
    pos, tri = scene_with_occluder()
    gs, os_ = prt.RTScene(pos, tri), oracle.Scene(pos, tri)
    args = ([4, 4, 4], [12, 10, 8], [6, 6, 6])
    g0, g1, gsc = prt.calculate_weight(gs, *args)
    o0, o1, osc = oracle.volume_weights(os_, *args)
    assert np.array_equal(np.isnan(gsc), np.isnan(osc)) and np.allclose(np.nan_to_num(gsc), np.nan_to_num(osc), atol=1e-6)
    assert np.abs(g0 - o0).max() <= 1e-6 and np.abs(g1 - o1).max() <= 1e-6
    s = np.concatenate([g0, g1], 1).sum(1)
    assert np.all(np.isclose(s, 1.0, atol=1e-5) | (s == 0.0)) and (s == 0).mean() < 0.5
    assert (osc[~np.isnan(osc)] > 0.2).any()          # the relocation branch is exercised (voxels inside the torus tube)


def test_split_capture_merge_upload_project(prt, oracle):
    """multi-GPU probe path on one GPU: capture two probe slices, merge (dist.merge_probe_csr), upload the merged CSR
    (prt_csr_upload) and project -- identical to the single capture."""
    from prt_b200 import dist as pdist
    pos, tri = scene_with_occluder()
    gs = prt.RTScene(pos, tri)
    probes = prt.probe_positions([4, 3, 2], [6, 6, 6])
    d, w = prt.fibonacci_dirs(1200)
    whole = prt.ProbeTransfer(gs, probes, d, w)
    wr, wi, wt, wsf, wk = whole.download()
    parts = []
    for r in range(3):
        lo, hi = pdist.probe_shard_range(len(probes), 3, r)
        pt = prt.ProbeTransfer(gs, probes[lo:hi], d, w)
        rng, ids, tr, _, keys = pt.download()
        parts.append(dict(range=rng, ids=ids, transfer=tr, keys=keys, sums=pt.surfel_sums()))
    mr, mi, mt, msf, mk = pdist.merge_probe_csr(parts)
    assert np.array_equal(mr, wr) and np.array_equal(mi, wi) and np.array_equal(mt, wt) and np.array_equal(mk, wk)
    assert np.abs(msf - wsf).max() <= 1e-6
    sums = whole.surfel_sums()
    assert sums.shape == (whole.n_surfels, 7) and (sums[:, 6] >= 1).all() and np.allclose(sums[:, :3] / sums[:, 6:7], wsf[:, :3], atol=1e-5)
    up = prt.ProbeTransfer.from_arrays(gs.ctx, mr, mi, mt, msf, mk)
    assert (up.n_probes, up.nnz, up.n_surfels) == (whole.n_probes, whole.nnz, whole.n_surfels)
    rad = np.random.RandomState(3).rand(whole.n_surfels, 4).astype(np.float32)
    assert np.array_equal(up.project(rad), whole.project(rad))
    ur = up.download()
    assert all(np.array_equal(a, b) for a, b in zip(ur, (mr, mi, mt, msf, mk)))
    with pytest.raises(prt.PRTError):
        up.surfel_sums()                                   # uploaded CSRs carry no accumulators
    bad = mi.copy(); bad[0] = whole.n_surfels
    with pytest.raises(prt.PRTError):
        prt.ProbeTransfer.from_arrays(gs.ctx, mr, bad, mt, msf, mk)
