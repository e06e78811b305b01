"""GPU parity tests for the per-frame probe pipeline (Paral_Shadow, relight.comp, precomp_projectSH.comp,
transfer2volume.comp) against the oracle (oracle/gi.c)."""
import numpy as np
import pytest

from test_gpu_probe import scene_with_occluder

pytestmark = pytest.mark.gpu


def test_paral_shadow_matrix_and_map_match_oracle(prt, oracle):
    pos, tri = scene_with_occluder()
    gs, os_ = prt.RTScene(pos, tri), oracle.Scene(pos, tri)
    for up, dr in ((0.17, 0.84), (0.05, 0.3), (0.5, 0.0)):
        gd, gm = prt.paral_shadow_matrix(up, dr)
        od, om = oracle.paral_shadow_matrix(up, dr)
        assert np.array_equal(gd, od) and np.array_equal(gm, om)
    gd, gm = prt.paral_shadow_matrix(0.17, 0.84)
    g, o = prt.shadow_map(gs, gm, 96), oracle.shadow_map(os_, gm, 96)
    assert np.array_equal(g, o)                                            # pinned ray set-up and triangle test: bit-exact depths
    assert (g < 1).sum() > 300 and (g == 1).sum() > 300
    bad = gm.copy(); bad[3] = 0.1                                          # perspective matrix: rejected
    with pytest.raises(prt.PRTError):
        prt.shadow_map(gs, bad, 16)


def _setup(prt, oracle, pres, vres, n_dirs=1500):
    pos, tri = scene_with_occluder()
    gs, os_ = prt.RTScene(pos, tri), oracle.Scene(pos, tri)
    size = [6.0, 6.0, 6.0]
    probes = prt.probe_positions(pres, size)
    d, w = prt.fibonacci_dirs(n_dirs)
    pt = prt.ProbeTransfer(gs, probes, d, w)
    weights = prt.calculate_weight(gs, pres, vres, size)
    return gs, os_, size, pt, weights


def test_gi_rounds_match_oracle(prt, oracle):
    pres, vres = [4, 3, 3], [8, 6, 6]
    gs, os_, size, pt, weights = _setup(prt, oracle, pres, vres)
    rng, ids, tr, sf, keys = pt.download()
    sky_dir, M = prt.paral_shadow_matrix(0.17, 0.84)
    depth = prt.shadow_map(gs, M, 128)
    P = prt.RelightParams.make(sky_dir, M, cast_position=(2.5, 0.5, 0.2), cast_cutoff=0.5, ambient_intensity=(3, 2, 1), ambient_position=(0, 4, 0),
                               multi_bounce=True, atten=0.8, sh_shift=0.1)
    vol = prt.SHVolume(pt, pres, vres, size, weights)
    vol.set_shadow_map(depth)
    # oracle state, advanced with the same stages
    o_rad = np.zeros((pt.n_surfels, 4), np.float32)
    o_vol = np.zeros((int(np.prod(vres)), 7, 4), np.float32)
    o_pt = oracle.ProbeTransfer(os_, prt.probe_positions(pres, size), *prt.fibonacci_dirs(1500))
    assert np.array_equal(o_pt.download()[1], ids)
    for rnd in range(3):
        vol.step(P, 1)
        g_rad, g_psh, g_vol = vol.download()
        o_rad_next = oracle.relight(P, sf, o_rad, depth=depth, volumes=o_vol, volume_res=vres, scene_size=size)
        if rnd == 0:
            assert np.array_equal(g_rad, o_rad_next)                       # same inputs -> bit-identical relight
            assert (g_rad[:, :3] > 0).any() and (g_rad[:, 3] == 1).all()
        assert np.abs(g_rad - o_rad_next).max() <= 1e-5 * max(1.0, np.abs(o_rad_next).max())
        # the projection and the blend, each checked on the GPU's own input (bit-exact blend) and end to end
        o_psh_from_g = o_pt.project(g_rad)
        assert np.abs(g_psh - o_psh_from_g).max() <= 2e-5 * max(1.0, np.abs(o_psh_from_g).max())
        assert np.array_equal(g_vol, oracle.transfer_to_volume(g_psh, pres, weights[0], weights[1], vres))
        o_rad = o_rad_next
        o_vol = oracle.transfer_to_volume(o_pt.project(o_rad), pres, weights[0], weights[1], vres)
        assert np.abs(g_vol - o_vol).max() <= 5e-5 * max(1.0, np.abs(o_vol).max())
    assert np.abs(g_vol).max() > 0


def test_gi_multi_round_call_equals_single_rounds(prt, oracle):
    pres, vres = [3, 3, 2], [5, 6, 4]
    gs, os_, size, pt, weights = _setup(prt, oracle, pres, vres, n_dirs=800)
    sky_dir, M = prt.paral_shadow_matrix(0.3, 0.1)
    P = prt.RelightParams.make(sky_dir, M)
    alb = np.random.RandomState(5).rand(pt.n_surfels, 3).astype(np.float32)
    a = prt.SHVolume(pt, pres, vres, size, weights); b = prt.SHVolume(pt, pres, vres, size, weights)
    a.set_albedo(alb); b.set_albedo(alb)                                   # no shadow map: unshadowed sky
    a.step(P, 5)
    for _ in range(5):
        b.step(P, 1)
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x, y)
    rad = a.download()[0]
    o = oracle.relight(P, pt.download()[3], np.zeros_like(rad), albedo=alb)          # first round has zero feedback
    c = prt.SHVolume(pt, pres, vres, size, weights); c.set_albedo(alb); c.step(P, 1)
    assert np.array_equal(c.download()[0], o)
    c.set_radiance(rad); c.set_albedo(None); c.step(P, 0)
    assert np.array_equal(c.download()[0], rad)
    with pytest.raises(prt.PRTError):
        prt.SHVolume(pt, [2, 2, 2], vres, size, weights)                  # probe grid does not match the capture
