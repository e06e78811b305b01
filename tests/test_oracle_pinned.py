"""CPU tests: the oracle against outputs of the reference's OWN source (tests/golden/ref_*.txt, produced by oracle/_ref/ref_slices
from line ranges of /root/reference compiled in place -- see oracle/ref_slices.cpp and tests/golden/make_golden.py).

What is pinned here: sampling + frame (a3), get_dirs (a9 / f1 ray set), the BRDF LUT (a12, completely), prefilter / irradiance
shader arithmetic up to the texture sampler (a12; the sampler is GL-driver behaviour, pinned in oracle/env.c), equirect uv (a12).
Tolerances are stated per test: the reference evaluates libm sin/cos on float(2*PI*v), the oracle its pinned polynomial.
"""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")
PREF_OUT, IRR_OUT = 16, 8
ENV_W, ENV_H, ENV_CUBE = 256, 128, 64          # the environment tests/golden/make_golden.py fed to the reference's shaders


def rows(name, tag):
    out = []
    for line in open(os.path.join(G, name)):
        head, _, vals = line.partition(":")
        h = head.split()
        if h[0] != tag:
            continue
        out.append(([float(x) for x in h[1:]], [float(x) for x in vals.split()]))
    return out


def test_frame_matches_reference_source(oracle):
    """frame(N) (raytracing.cpp:101-107), incl. the |N.z| >= 0.99 branch and the axes: <= 2e-7 abs (normalize rounding)."""
    fr = rows("ref_sampling.txt", "frame")
    assert len(fr) == 40
    branch = 0
    for (n, ref) in fr:
        _, _, f, _ = oracle.cosine_world(0.5, 0.5, n)
        assert np.allclose(f.reshape(-1), ref, rtol=0, atol=2e-7), (n, f, ref)
        branch += abs(n[2]) >= 0.99
    assert branch >= 3


def test_cosine_sample_matches_reference_source(oracle):
    """cosineSampleHemisphere(u, v) / (u, v, N) + pdf (raytracing.cpp:130-160).  The reference calls libm cos/sin on the float
    product 2*PI*v (angle error up to 2.4e-7 * 2pi), the oracle a pinned polynomial on exact quadrants: <= 1.5e-6 abs on
    directions, 1e-6 on the pdf."""
    sm = rows("ref_sampling.txt", "sample")
    assert len(sm) == 480
    worst = 0.0
    for (h, ref) in sm:
        n, u, v = h[0:3], h[3], h[4]
        l, w, _, pdf = oracle.cosine_world(u, v, n)
        worst = max(worst, np.abs(l - ref[0:3]).max(), np.abs(w - ref[3:6]).max())
        assert abs(pdf - ref[6]) <= 1e-6
    assert worst <= 1.5e-6, worst


def test_get_dirs_matches_reference_source(oracle):
    """get_dirs (light_probe.cpp:136-152), n = 100 (calculate_weight) and 4096 (config 3): computed in double, stored float --
    bit-identical up to libm's last double ulp (<= 1 float ulp allowed)."""
    ref = np.array([[float(x) for x in line.split()[1:4]] for line in open(os.path.join(G, "ref_get_dirs.txt"))], np.float32)
    assert len(ref) == 100 + 4096
    for n, sl in ((100, slice(0, 100)), (4096, slice(100, 4196))):
        d, _ = oracle.fibonacci_dirs(n)
        assert np.abs(d - ref[sl]).max() <= 1.2e-7
        assert (d == ref[sl]).mean() > 0.99


def test_brdf_lut_matches_reference_source(oracle):
    """IntegrateBRDF (brdf.frag:69-107) on a 32x32 LUT, 1024 samples: <= 2e-5 abs (gate of the path: 1e-3)."""
    ref = rows("ref_brdf.txt", "lut")
    lut = oracle.brdf_lut(32, 32, 1024)
    worst = max(np.abs(lut[int(h[1]), int(h[0])] - v).max() for h, v in ref)
    assert len(ref) == 1024 and worst <= 2e-5, worst


def _env_cube(oracle):
    from prt_b200 import hdr
    return oracle.EnvCube(hdr.synthetic_env(ENV_W, ENV_H), ENV_CUBE)


def test_prefilter_matches_reference_source(oracle):
    """prefilter.frag main() (8-107) with the pinned sampler at texel centres of a 16^2 cube, roughness 0, .25, .5, .75, 1:
    <= 1e-4 relative to the texel's magnitude (gate of the path: 1e-3 abs)."""
    pre = _env_cube(oracle).prefilter(PREF_OUT, 5, 1024)
    ref = rows("ref_prefilter.txt", "prefilter")
    assert len(ref) == 6 * (4 + 4 + 4 + 4 + 1)
    worst = 0.0
    for h, v in ref:
        mip, f, i, j = (int(x) for x in h)
        worst = max(worst, np.abs(pre[mip][f, j, i] - v).max() / max(1.0, np.abs(v).max()))
    assert worst <= 1e-4, worst


def test_irradiance_matches_reference_source(oracle):
    """irradiance.frag main() (7-43) with the pinned sampler: the float loop counters must give the same 252 x 63 samples."""
    irr = _env_cube(oracle).irradiance(IRR_OUT)
    ref = rows("ref_irradiance.txt", "irradiance")
    assert len(ref) == 24
    worst = max(np.abs(irr[int(h[0]), int(h[2]), int(h[1])] - v).max() / max(1.0, np.abs(v).max()) for h, v in ref)
    assert worst <= 1e-4, worst


def test_equirect_uv_matches_reference_source(oracle):
    """SampleSphericalMap (rectangle2cube.frag:7-15): <= 2e-7."""
    ref = rows("ref_rect2cube.txt", "uv")
    worst = max(np.abs(oracle.equirect_uv(h) - v).max() for h, v in ref)
    assert len(ref) == 64 and worst <= 2e-7, worst
