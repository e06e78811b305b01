"""CPU tests: known answers for the oracle's image-based-lighting restatement (SURVEY 8c KATs 6, 7) and the RGBE reader on the
reference's own data/hdr/newport_loft.hdr (tests/golden/) against the statistics SURVEY section 0.5 records for stb_image."""
import os

import numpy as np
import pytest

from prt_b200 import hdr


def test_hdr_reader_matches_reference_asset_stats():
    path = os.path.join(os.path.dirname(__file__), "golden", "newport_loft.hdr")
    img = hdr.load_hdr(path)
    assert img.shape == (800, 1600, 3)
    assert abs(img.mean() - 0.2516) < 1e-3 and abs(img.max() - 15.25) < 1e-6


def test_constant_environment_known_answers(oracle):
    """KAT 6: constant environment 1 -> L00 = 0.282095 * 4 pi = 3.54491 with both quadratures; irradiance = pi."""
    env = oracle.EnvCube(np.ones((32, 64, 3), np.float32), 32)
    assert np.allclose(env.cube(0), 1.0) and np.allclose(env.cube(3), 1.0)
    for method, tol in ((0, 5e-4), (1, 1e-3)):
        L = env.project_sh(3, method)
        assert abs(L[0, 0] - 3.54491) < tol * 3.5
        # the cube rule samples texel CORNERS (image_projectSH.comp:66), which biases band 1 by ~2/size
        assert np.abs(L[1:]).max() < (2e-2 if method == 0 else 4e-2)
    # R-H pack: SH_Irad of a constant environment = c4 * L00 = pi for every normal (KAT 4 / SH.glsl:17-36)
    L = np.zeros((9, 3), np.float32); L[0] = 3.54491
    packed = oracle.sh_pack_rh(L)
    assert np.allclose(packed[[3, 7, 11]], 0.886227 * 3.54491, atol=1e-5) and abs(packed[3] - np.pi) < 1e-3
    irr = env.irradiance(4)
    assert np.abs(irr - np.pi).max() < 2e-2          # 0.025-step Riemann sum of irradiance.frag


def test_cube_sampling_pins(oracle):
    """face table == cubeCoordToWorld, bilinear at texel centres is exact, seams are continuous."""
    rs = np.random.RandomState(0)
    eq = rs.rand(16, 32, 3).astype(np.float32)
    env = oracle.EnvCube(eq, 8)
    table = [lambda u, v: (1, -v, -u), lambda u, v: (-1, -v, u), lambda u, v: (u, 1, v), lambda u, v: (u, -1, -v),
             lambda u, v: (u, -v, 1), lambda u, v: (-u, -v, -1)]
    c0 = env.cube(0)
    for f in range(6):
        for (i, j) in [(0, 0), (3, 5), (7, 7)]:
            d = table[f](2 * (i + 0.5) / 8 - 1, 2 * (j + 0.5) / 8 - 1)
            assert np.allclose(env.sample(d, 0.0), c0[f, j, i], atol=1e-6)
    # continuity across the +X/+Z edge
    a = env.sample([1.0, 0.2, 0.999], 0.0); b = env.sample([0.999, 0.2, 1.0], 0.0)
    assert np.abs(a - b).max() < 0.05
    # mip 1 is the 2x2 box of mip 0; trilinear lerps
    assert np.allclose(env.cube(1)[2, 1, 2], c0[2, 2:4, 4:6].mean((0, 1)), atol=1e-6)
    d = [0.3, -0.5, 0.8]
    assert np.allclose(env.sample(d, 0.5), 0.5 * (env.sample(d, 0.0) + env.sample(d, 1.0)), atol=1e-6)


def test_brdf_lut_known_answers(oracle):
    """KAT 7: A + B <= 1; (NdotV -> 1, roughness -> 0) -> (~1, ~0); A decreases with roughness at fixed NdotV."""
    lut = oracle.brdf_lut(32, 32, 1024)
    assert (lut.sum(-1) <= 1.0 + 1e-4).all() and (lut >= 0).all()
    assert lut[0, 31, 0] > 0.99 and lut[0, 31, 1] < 1e-3
    assert (np.diff(lut[:, 24, 0]) <= 1e-3).all()


def test_prefilter_roughness_zero_is_the_environment(oracle):
    eq = hdr.synthetic_env(128, 64)
    env = oracle.EnvCube(eq, 32)
    pf = env.prefilter(32, 5, 64)
    assert np.allclose(pf[0], env.cube(0), atol=1e-5)       # roughness 0: every sample reflects to N, mip 0
    assert pf[4].shape == (6, 2, 2, 3) and np.isfinite(pf[4]).all()
    # energy is roughly preserved, detail is lost with roughness
    assert abs(pf[2].mean() / env.cube(0).mean() - 1) < 0.25 and pf[3].std() < pf[1].std()
