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


# ---- the reference's WHOLE files, compiled unmodified with Embree replaced by the oracle's tracer (oracle/ref_bake.cpp, ref_weight.cpp) ----
BAKE_CASES = {"shadow": (32, 24, 16, 2, 1.0), "bounce": (32, 24, 8, 4, 0.5), "default": (16, 12, 32, 2, 1.0)}   # = tests/golden/make_golden.py


@pytest.mark.parametrize("case", ["shadow", "bounce", "default"])
def test_bake_SH_matches_the_reference_binary(oracle, case):
    """bake_SH(Mesh&) of raytracing.cpp (RTScene set-up, frame, cosine sampling, renderSH path logic with its cut-offs and offsets,
    estimator) executed from the reference's own source -> sh_coeff[9] of every vertex; the oracle's literal restatement, fed with the
    same std::mt19937 random() sequence, must reproduce it.  Tolerance 5e-6 relative L2 per vertex (libm sin/cos vs the pinned polynomial
    at 1e-7 in the directions; measured <= 9e-7) -- a single flipped visibility decision would show as >= 4e-4."""
    from prt_b200 import meshes
    nu, nv, res, mpl, alb = BAKE_CASES[case]
    pos, nrm, tri = meshes.bumpy_torus(nu, nv)
    ref = np.loadtxt(os.path.join(G, f"ref_bake_SH_{case}.txt"), dtype=np.float32)
    assert ref.shape == (nu * nv, 9)
    rnd = oracle.mt19937_floats(len(pos) * 9 * res * res * 2 + len(pos) * 9 * res * res)          # jitter pairs + room for the bounce draws
    got, used = oracle.bake_transfer_ref_order(oracle.Scene(pos, tri), pos, nrm, res=res, max_path_length=mpl, albedo=(alb,) * 3, rnd=rnd, u_first=0)
    assert len(pos) * 9 * res * res * 2 < used <= len(rnd)                                           # hits consumed random() pairs as in renderSH
    rel = np.linalg.norm(got - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-20)
    assert rel.max() <= 5e-6, rel.max()
    assert 0.05 < (np.abs(ref[:, 0]) < 0.28).mean()                                                  # the mesh really shadows itself


@pytest.mark.parametrize("case", ["shadow", "bounce"])
def test_production_oracle_equals_the_literal_restatement(oracle, case):
    """The oracle the GPU is tested against (one trace per sample shared by all coefficients, any-hit on the last segment, double
    accumulation, Philox draws) against the literal reference-order loop drawing the same Philox numbers: <= 5e-6 (float vs double
    accumulation; measured 2e-6)."""
    from prt_b200 import meshes
    nu, nv, res, mpl, alb = BAKE_CASES[case]
    pos, nrm, tri = meshes.bumpy_torus(nu, nv)
    sc = oracle.Scene(pos, tri)
    kw = dict(order=3, samples_u=res, samples_v=res, cs_phase=1, albedo=(alb,) * 3)
    if mpl > 2:
        kw.update(mode=oracle.INTERREFLECT, bounces=mpl - 2)
    prod, _, _ = oracle.bake_transfer(sc, pos, nrm, oracle.make_params(**kw), vertex_id_base=11)
    lit, _ = oracle.bake_transfer_ref_order(sc, pos, nrm, res=res, max_path_length=mpl, albedo=(alb,) * 3, rnd=None, vertex_id_base=11)
    rel = np.linalg.norm(prod - lit, axis=1) / np.maximum(np.linalg.norm(lit, axis=1), 1e-20)
    assert rel.max() <= 5e-6, rel.max()


def weight_scene():
    from prt_b200 import meshes
    pos, _, tri = meshes.load_obj_assimp(os.path.join(G, "cube.obj"))
    tp, _, tt = meshes.bumpy_torus(40, 28)
    return (np.concatenate([pos, tp * np.float32(1.3)]).astype(np.float32), np.concatenate([tri, tt + np.uint32(len(pos))]).astype(np.uint32))


def test_calculate_weight_matches_the_reference_binary(oracle):
    """calculate_weight of light_probe.cpp executed from the reference's own source (inside scores from 100 get_dirs rays, relocation to
    the least-inside neighbour, 8 segment rays, masking, renormalisation) on data/cube.obj + a torus: bit-identical weights."""
    pos, tri = weight_scene()
    ref = np.loadtxt(os.path.join(G, "ref_volume_weight.txt"), dtype=np.float32)
    w0, w1, _ = oracle.volume_weights(oracle.Scene(pos, tri), [4] * 3, [12] * 3, [6.18] * 3)
    got = np.concatenate([w0, w1], 1)
    assert ref.shape == got.shape == (1728, 8)
    assert 0.1 < (ref == 0).mean() < 0.9                       # occluded probes are masked, others are not
    assert np.array_equal(got, ref, equal_nan=True) and np.isnan(ref).mean() < 0.05


def test_probe_capture_matches_the_reference_loop(oracle):
    """The CPU half of SH_volume::precompute executed from the reference's own source (volume.cpp:185-315; G-buffer = one oracle ray per
    texel centre of its 64^2 x 6 cubemap): same CSR structure (ranges, number of entries and surfels, id sets per probe up to the
    reference's first-seen numbering), transfer rows <= 1e-6, surfel table <= 5e-4 (the reference averages in float)."""
    L = open(os.path.join(G, "ref_probe_capture.txt")).read().splitlines()
    npb, nnz, nprim = (int(x) for x in L[0].split()[1:])
    probes = np.array([[float(x) for x in ln.split()[1:4]] for ln in L[1:1 + npb]], np.float32)
    rng = np.array([[int(x) for x in ln.split()[5:7]] for ln in L[1:1 + npb]], np.uint32)
    ent = L[1 + npb:1 + npb + nnz]
    ids = np.array([int(ln.split()[1]) for ln in ent], np.int64)
    tr = np.array([[float(x) for x in ln.split()[3:12]] for ln in ent], np.float32)
    sf = np.array([[float(x) for x in ln.split()[3:9]] for ln in L[1 + npb + nnz:]], np.float32)
    assert len(sf) == nprim and npb == 3 and nnz > 500
    pos, tri = weight_scene()
    op = oracle.probe_positions([4] * 3, [6.18] * 3)[:npb]
    assert np.array_equal(op, probes)                                            # volume.cpp:83-90
    d, w = oracle.cube_dirs(64)
    o = oracle.ProbeTransfer(oracle.Scene(pos, tri), op, d, w)
    orng, oids, otr, osf, _ = o.download()
    assert (o.nnz, o.n_surfels) == (nnz, nprim) and np.array_equal(orng, rng)
    # reference id (first seen) -> oracle id (rank of the cluster key): match the surfel tables
    dist = np.abs(sf[:, None, :] - osf[None, :, :]).max(-1)
    perm = dist.argmin(1)
    assert dist.min(1).max() <= 5e-4 and len(set(perm.tolist())) == nprim
    for p in range(npb):
        a = slice(rng[p, 0], rng[p, 1])
        rid = perm[ids[a]]
        order = np.argsort(rid)
        assert np.array_equal(rid[order], oids[a].astype(np.int64))
        assert np.abs(tr[a][order] - otr[a]).max() <= 1e-6


def test_project_sh_matches_the_reference_kernel(oracle):
    """precomp_projectSH.comp main() (CSR SpMV over 128 invocations, shared-memory tree reduction, sinc window, Ramamoorthi-Hanrahan
    pack) compiled from the reference and run on host threads: <= 2e-6 (summation order)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    rng, ids, tr, rad = mg.project_case()
    ref = np.array([[float(x) for x in ln.split()[3:]] for ln in open(os.path.join(G, "ref_project.txt"))], np.float32).reshape(len(rng), 7, 4)
    got = oracle.project_arrays(rng, ids, tr, rad)
    assert np.abs(got - ref).max() <= 2e-6, np.abs(got - ref).max()
    assert np.array_equal(ref[0, :6], np.zeros((6, 4), np.float32)) and ref[0, 6, 3] == 1.0        # empty range: zeros, C.w = 1
