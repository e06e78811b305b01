"""CPU property tests of the horizon map's per-item bounds (prt_b200/csrc/horizon_math.cuh compiled as plain C++ by
tests/hostcheck): every bound must be CONSERVATIVE -- for any point p of the triangle / box that lies above the tangent plane,
sin(elevation) = p.z / |p| does not exceed the item's value and p's azimuth bin lies inside the item's bin range.  A ray the
horizon pass frees is never traced, so a bound that is too low would silently turn occluded rays into visible ones; the GPU
parity tests catch that end to end, these tests pin the formulas themselves without a GPU.

The host build uses exact division / square roots where the device uses __fdividef / rsqrtf (2 ulp); the items carry margins
(2e-4 in sin-elevation, 0.02 in pseudo-angle) that are orders of magnitude above that difference."""
import numpy as np
import pytest

BINS = 32


def pang(x, y):
    """diamond pseudo-angle in [0,4) -- the formula of hz_pang / the host sample table (abi.cu)"""
    p = y / (np.abs(x) + np.abs(y))
    return np.where(x < 0, 2.0 - p, np.where(y < 0, 4.0 + p, p))


def check_points(pts, b0, b1, v, what):
    """pts: [k,3] local-frame points of one primitive; asserts the conservativeness of the item (b0, b1, v)."""
    z = pts[:, 2]
    r = np.linalg.norm(pts, axis=1)
    up = (z > 0) & (r > 0)
    if not up.any():
        return
    sin_el = z[up] / r[up]
    assert v > 0, f"{what}: geometry above the tangent plane but the item is empty"
    assert sin_el.max() <= v + 1e-6, f"{what}: sin(elevation) {sin_el.max():.6f} above the bound {v:.6f}"
    if b1 - b0 >= BINS - 1:
        return
    q = pts[up]
    planar = np.abs(q[:, 0]) + np.abs(q[:, 1])
    ok = planar > 1e-3 * r[up]           # azimuth is undefined on the vertical axis
    if not ok.any():
        return
    bins = np.floor(pang(q[ok, 0].astype(np.float64), q[ok, 1].astype(np.float64)) * (BINS / 4)).astype(int) % BINS
    inside = ((bins - b0) % BINS) <= (b1 - b0)
    assert inside.all(), f"{what}: azimuth bins {sorted(set(bins[~inside]))} outside [{b0},{b1}]"


def test_pseudo_angle_matches_python(hostcheck):
    rng = np.random.RandomState(0)
    xy = rng.normal(size=(2000, 2)).astype(np.float32)
    got = np.array([hostcheck.hc_hz_pang(float(x), float(y)) for x, y in xy])
    ref = pang(xy[:, 0].astype(np.float64), xy[:, 1].astype(np.float64))
    assert np.abs(got - ref).max() < 1e-5
    assert (got >= 0).all() and (got <= 4.0).all()
    # monotone in the true angle
    a = np.linspace(-np.pi, np.pi, 4001)[1:-1]
    p = pang(np.cos(a), np.sin(a))
    p = np.where(a < 0, p - 4.0, p)
    assert (np.diff(p) > 0).all()


def _tri_points(t, n=24):
    """barycentric grid over one triangle [3,3] (edges and interior)"""
    u, w = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n))
    m = u + w <= 1.0 + 1e-12
    u, w = u[m], w[m]
    return (1 - u - w)[:, None] * t[0] + u[:, None] * t[1] + w[:, None] * t[2]


@pytest.mark.parametrize("kind", ["generic", "near_axis", "crossing_plane", "sliver", "far_small"])
def test_triangle_bound_is_conservative(hostcheck, kind):
    rng = np.random.RandomState({"generic": 1, "near_axis": 2, "crossing_plane": 3, "sliver": 4, "far_small": 5}[kind])
    n = 400
    tri = rng.normal(size=(n, 3, 3))
    if kind == "near_axis":            # triangles around / close to the vertical through the origin
        tri[:, :, :2] *= 0.2
        tri[:, :, 2] = np.abs(tri[:, :, 2]) + 0.05
    elif kind == "crossing_plane":     # one vertex below, two above the tangent plane
        tri[:, 0, 2] = -np.abs(tri[:, 0, 2])
        tri[:, 1:, 2] = np.abs(tri[:, 1:, 2])
    elif kind == "sliver":             # long thin triangles at grazing elevation
        c = rng.normal(size=(n, 1, 3)) * 3
        d = rng.normal(size=(n, 1, 3))
        tri = c + d * np.array([0.0, 1.0, 1.0])[None, :, None] * np.array([[0], [5.0], [5.001]])[None] + rng.normal(size=(n, 3, 3)) * 1e-3
        tri[:, :, 2] = np.abs(tri[:, :, 2]) * 0.02
    elif kind == "far_small":
        tri = rng.normal(size=(n, 1, 3)) * 20 + rng.normal(size=(n, 3, 3)) * 0.05
    tri = tri.astype(np.float32)
    bins = np.zeros((n, 2), np.int32)
    val = np.zeros(n, np.float32)
    q = np.ascontiguousarray(tri.reshape(n, 9))
    hostcheck.hc_hz_triangle(q.ctypes.data, n, bins.ctypes.data, val.ctypes.data)
    for i in range(n):
        check_points(_tri_points(tri[i].astype(np.float64)), int(bins[i, 0]), int(bins[i, 1]), float(val[i]), f"{kind} triangle {i}")
    if kind == "far_small":            # and the bound is tight where it can be: within 2e-3 of the true maximum
        for i in range(0, n, 7):
            p = _tri_points(tri[i].astype(np.float64), 64)
            true = max(0.0, (p[:, 2] / np.linalg.norm(p, axis=1)).max())
            if true > 0:
                assert val[i] - true < 2e-3


def _box_points(c, e, rng, k=400):
    s = rng.uniform(-1, 1, size=(k, 3))
    # push a third of the samples onto faces / edges / corners, where the extrema are
    s[: k // 3] = np.sign(s[: k // 3])
    s[k // 3: 2 * k // 3, rng.randint(0, 3)] = np.sign(s[k // 3: 2 * k // 3, 0])
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float64)
    s = np.concatenate([s, corners])
    return c + s * e


@pytest.mark.parametrize("which,name", [(0, "cheap (cone clamped by the slab bound)"), (1, "exact box"), (2, "cone")])
@pytest.mark.parametrize("kind", ["far", "near", "flat_on_surface", "overhead"])
def test_box_bounds_are_conservative(hostcheck, which, name, kind):
    rng = np.random.RandomState(10 * which + {"far": 1, "near": 2, "flat_on_surface": 3, "overhead": 4}[kind])
    n = 500
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    c = rng.normal(size=(n, 3))
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    if kind == "far":
        c *= rng.uniform(3, 30, size=(n, 1)); e = rng.uniform(0.01, 1.0, size=(n, 3))
    elif kind == "near":
        c *= rng.uniform(0.3, 2.0, size=(n, 1)); e = rng.uniform(0.05, 1.0, size=(n, 3))
    elif kind == "flat_on_surface":    # thin slabs hugging the tangent plane: the case the slab bound is made for
        nrm[:] = 0.0
        ax = rng.randint(0, 3, n)
        nrm[np.arange(n), ax] = rng.choice([-1.0, 1.0], n)
        c *= rng.uniform(1.0, 8.0, size=(n, 1))
        c[np.arange(n), ax] = rng.uniform(-0.02, 0.05, n) * nrm[np.arange(n), ax]
        e = rng.uniform(0.2, 1.5, size=(n, 3))
        e[np.arange(n), ax] = rng.uniform(0.005, 0.05, n)
    else:                              # boxes over the vertex
        c = nrm * rng.uniform(0.5, 5.0, size=(n, 1)) + rng.normal(size=(n, 3)) * 0.1
        e = rng.uniform(0.05, 1.0, size=(n, 3))
    c32, e32, n32 = (np.ascontiguousarray(a, np.float32) for a in (c, e, nrm))
    bins = np.zeros((n, 2), np.int32)
    val = np.zeros(n, np.float32)
    hostcheck.hc_hz_box(c32.ctypes.data, e32.ctypes.data, n32.ctypes.data, n, which, bins.ctypes.data, val.ctypes.data)
    fr = np.zeros(9, np.float32)
    tighter = 0
    for i in range(n):
        hostcheck.hc_frame(n32[i].ctypes.data, fr.ctypes.data)
        R = fr.reshape(3, 3).astype(np.float64)                      # rows: right, up, n
        pts = _box_points(c32[i].astype(np.float64), e32[i].astype(np.float64), rng) @ R.T
        check_points(pts, int(bins[i, 0]), int(bins[i, 1]), float(val[i]), f"{name}, {kind} box {i}")
        tighter += 1
    assert tighter == n


def test_slab_clamp_tightens_flat_boxes(hostcheck):
    """The reason the cheap bound exists: for a thin box lying in the tangent plane the cone around the bounding sphere lifts
    the horizon by the box's angular radius, z_max / d_min does not."""
    n = 200
    rng = np.random.RandomState(7)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (n, 1))
    c = np.zeros((n, 3), np.float32)
    c[:, 0] = rng.uniform(3, 6, n)
    e = np.zeros((n, 3), np.float32)
    e[:, 0] = e[:, 1] = rng.uniform(0.5, 1.0, n)
    e[:, 2] = 0.01
    out = {}
    for which in (0, 2):
        bins = np.zeros((n, 2), np.int32)
        val = np.zeros(n, np.float32)
        hostcheck.hc_hz_box(c.ctypes.data, e.ctypes.data, nrm.ctypes.data, n, which, bins.ctypes.data, val.ctypes.data)
        out[which] = val.copy()
    assert (out[0] <= out[2] + 1e-7).all()
    assert np.median(out[2] / out[0]) > 10          # cone: ~0.2, slab: ~0.005


# ---- oriented slabs (bvh8.h Slab32, hz_slab_value) ---------------------------------------------------------------------------------
def _slab_cases(rng, n, kind):
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    c = rng.normal(size=(n, 3))
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    c *= rng.uniform(1.5, 12.0, size=(n, 1))
    e = rng.uniform(0.05, 1.0, size=(n, 3))
    m = rng.normal(size=(n, 3))
    if kind == "sheet_near_tangent":          # a sheet almost parallel to the tangent plane: the case the slab is made for
        m = nrm + rng.normal(size=(n, 3)) * 0.15
    elif kind == "axis":                      # slab normal along a coordinate axis (zero components: no breakpoint there)
        m[:] = 0.0
        m[np.arange(n), rng.randint(0, 3, n)] = rng.choice([-1.0, 1.0], n)
    m /= np.linalg.norm(m, axis=1, keepdims=True)
    # slab through the box: centre offset within the box's extent along m, random thickness
    ext = (np.abs(m) * e).sum(axis=1)
    mid = (m * c).sum(axis=1) + rng.uniform(-0.8, 0.8, n) * ext
    half = rng.uniform(0.01, 0.5, n) * ext
    if kind == "wide":                        # slab wider than the box: must reduce to the plain box bound
        half = 3.0 * ext
    return nrm, c, e, m, mid - half, mid + half


@pytest.mark.parametrize("kind", ["generic", "sheet_near_tangent", "axis", "wide"])
def test_slab_bound_is_conservative_and_solves_the_lp(hostcheck, kind):
    """hz_slab_value bounds sin(elevation) of everything inside box AND slab: never below brute-force samples of that set, and its
    height term equals the optimum of the linear programme (scipy) -- the dual breakpoints it tries include the minimiser."""
    from scipy.optimize import linprog
    rng = np.random.RandomState({"generic": 21, "sheet_near_tangent": 22, "axis": 23, "wide": 24}[kind])
    n = 300
    nrm, c, e, m, L0, U0 = _slab_cases(rng, n, kind)
    c32, e32, n32 = (np.ascontiguousarray(a, np.float32) for a in (c, e, nrm))
    slab = np.ascontiguousarray(np.concatenate([m, L0[:, None], U0[:, None]], axis=1), np.float32)
    val = np.zeros(n, np.float32)
    hostcheck.hc_hz_slab(c32.ctypes.data, e32.ctypes.data, n32.ctypes.data, slab.ctypes.data, n, val.ctypes.data)
    box = np.zeros(n, np.float32)
    bins = np.zeros((n, 2), np.int32)
    hostcheck.hc_hz_box(c32.ctypes.data, e32.ctypes.data, n32.ctypes.data, n, 0, bins.ctypes.data, box.ctypes.data)
    fr = np.zeros(9, np.float32)
    tight = eligible = 0
    for i in range(n):
        cc, ee, mm = c32[i].astype(np.float64), e32[i].astype(np.float64), slab[i, :3].astype(np.float64)
        lo, hi = float(slab[i, 3]), float(slab[i, 4])
        pts = _box_points(cc, ee, rng, k=3000)
        # project a share of the samples onto the two slab planes (clipped to the box afterwards), where the extrema lie
        d = pts @ mm
        k3 = len(pts) // 3
        pts[:k3] += np.outer(hi - d[:k3], mm) / (mm @ mm)
        pts[k3:2 * k3] += np.outer(lo - d[k3:2 * k3], mm) / (mm @ mm)
        d = pts @ mm
        keep = (d >= lo - 1e-9) & (d <= hi + 1e-9) & (np.abs(pts - cc) <= ee + 1e-9).all(axis=1)
        pts = pts[keep]
        hostcheck.hc_frame(n32[i].ctypes.data, fr.ctypes.data)
        nn = fr.reshape(3, 3).astype(np.float64)[2]
        if len(pts):
            z = pts @ nn
            r = np.linalg.norm(pts, axis=1)
            up = (z > 0) & (r > 0)
            if up.any():
                assert (z[up] / r[up]).max() <= val[i] + 1e-6, f"{kind} {i}: a point of box and slab lies above the bound"
        # exactness of the height term: max n . x over the polytope
        res = linprog(-nn, A_ub=np.stack([mm, -mm]), b_ub=[hi, -lo], bounds=list(zip(cc - ee, cc + ee)), method="highs")
        if res.status == 0:
            zmax = -res.fun
            mxyz = np.maximum(np.abs(cc) - ee, 0.0)
            gap = max(lo, -hi, 0.0)
            dm = max(np.linalg.norm(mxyz), gap * np.sqrt(0.999))
            if dm > 0 and zmax > 1e-3:
                want = zmax / dm * 1.0001 + (1e-6 + 2e-4)
                eligible += 1
                assert val[i] >= want - 2e-5 * max(1.0, want), f"{kind} {i}: bound {val[i]} below the programme's optimum {want}"
                if abs(val[i] - want) <= 1e-4 * max(1.0, want):
                    tight += 1
        if kind == "wide":
            assert val[i] >= min(box[i], 2.0) - 1e-5 or box[i] >= 1.0       # a slab that does not cut cannot beat the box's z_max / d_min
    assert eligible > n // 4
    assert tight >= 0.98 * eligible, f"only {tight} of {eligible} bounds equal the optimum of the linear programme"


def test_builder_slabs_contain_their_triangles(hostcheck):
    """Slab32 of every 8-wide node (bvh_build.cpp): all triangles below the node lie between its two planes; on a surface mesh the
    slabs of the lower levels are thin compared with their boxes."""
    from prt_b200 import meshes
    pos, nrm, tri = meshes.bumpy_torus(96, 64)
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    info = np.zeros(3, np.uint32)
    hostcheck.hc_info(h, info.ctypes.data)
    n_nodes, n_tris = int(info[0]), int(info[1])
    slabs = np.zeros((n_nodes, 8), np.float32)
    assert hostcheck.hc_slabs(h, slabs.ctypes.data, n_nodes) == n_nodes
    rng_ = np.zeros((n_nodes, 2), np.uint32)
    hostcheck.hc_node_tri_ranges(h, rng_.ctypes.data)
    tv = np.zeros((n_tris, 9), np.float32)
    hostcheck.hc_tris(h, tv.ctypes.data)
    hostcheck.hc_free(h)
    assert rng_[0, 0] == 0 and rng_[0, 1] == n_tris
    thin = []
    for x in range(n_nodes):
        m, d0, d1 = slabs[x, :3].astype(np.float64), float(slabs[x, 3]), float(slabs[x, 4])
        v = tv[rng_[x, 0]: rng_[x, 0] + rng_[x, 1]].reshape(-1, 3).astype(np.float64)
        assert len(v)
        if not m.any():
            continue
        assert abs(np.linalg.norm(m) - 1.0) < 1e-5
        d = v @ m
        assert d.min() >= d0 and d.max() <= d1, f"node {x}: triangles outside the slab"
        if rng_[x, 1] <= 64:
            thin.append((d1 - d0) / max(np.linalg.norm(v.max(axis=0) - v.min(axis=0)), 1e-30))
    assert np.median(thin) < 0.25


def test_builder_dops_contain_their_children(hostcheck):
    """Dop32 of every 8-wide node (bvh_build.cpp): the triangles of every child lie inside the child's quantised extent along the node's
    scaled mean normal, with at least one step to spare either side (the padding the float evaluation on the device lives on), and the
    extents are worth having: on a surface mesh the median child spans well under half of the 255 steps."""
    from prt_b200 import meshes
    pos, nrm, tri = meshes.bumpy_torus(96, 64)
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    info = np.zeros(3, np.uint32)
    hostcheck.hc_info(h, info.ctypes.data)
    n_nodes, n_tris = int(info[0]), int(info[1])
    raw = np.zeros((n_nodes, 32), np.uint8)
    assert hostcheck.hc_dops(h, raw.ctypes.data, n_nodes) == n_nodes
    M = raw[:, :16].copy().view(np.float32).reshape(n_nodes, 4)
    qlo, qhi = raw[:, 16:24].astype(np.float64), raw[:, 24:32].astype(np.float64)
    ranges = np.zeros((n_nodes, 8, 2), np.uint32)
    hostcheck.hc_child_tri_ranges(h, ranges.ctypes.data)
    tv = np.zeros((n_tris, 9), np.float32)
    hostcheck.hc_tris(h, tv.ctypes.data)
    hostcheck.hc_free(h)
    spans = []
    checked = 0
    for x in range(n_nodes):
        m, d0 = M[x, :3].astype(np.float64), float(M[x, 3])
        if not m.any():
            assert (qlo[x] == 0).all() and (qhi[x] == 255).all()
            continue
        for s_ in range(8):
            t0, cnt = int(ranges[x, s_, 0]), int(ranges[x, s_, 1])
            if cnt == 0:
                continue
            v = tv[t0:t0 + cnt].reshape(-1, 3).astype(np.float64)
            sv = v @ m - d0
            lo_ok = sv.min() >= qlo[x, s_] + 1.0 or qlo[x, s_] == 0
            hi_ok = sv.max() <= qhi[x, s_] - 1.0 or qhi[x, s_] == 255
            assert lo_ok and hi_ok, (x, s_, sv.min(), sv.max(), qlo[x, s_], qhi[x, s_])
            assert -1e-3 <= sv.min() and sv.max() <= 255.001            # inside the node's own range
            spans.append(qhi[x, s_] - qlo[x, s_])
            checked += 1
    assert checked > n_nodes
    assert np.median(spans) < 160 and np.percentile(spans, 25) < 110


# ---- the warp-cooperative builder itself, run on the CPU through the warp emulator (tests/hostcheck/warp_emu.h) --------------------
def _maps(hostcheck, h, pos, nrm, budget=64, near=157, eps=1e-4, stats=None, slabs=1, mid=24, gain=0.2):
    """defaults = the product's (abi.cu: horizon_near 157, horizon_mid 24, horizon_gain 6.4 samples = 0.2 x 1024 / 32, slabs on)"""
    n = len(pos)
    hostcheck.hc_use_slabs(slabs)
    hostcheck.hc_horizon_mid(mid, gain)
    hz = np.zeros((n, BINS), np.float32)
    ncand = np.zeros(n, np.int32)
    p32, n32 = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nrm, np.float32)
    hostcheck.hc_horizon_maps(h, p32.ctypes.data, n32.ctypes.data, n, eps, budget, near, hz.ctypes.data, ncand.ctypes.data,
                              stats.ctypes.data if stats is not None else None)
    return hz, ncand


def _free_mask(dirs, hz):
    """classification of horizon_kernel (horizon.cu): a sample is free iff its local z exceeds the map in its azimuth bin; the
    bin is the one abi.cu's ensure_samples stores in the sample table"""
    den = np.abs(dirs[:, 0].astype(np.float64)) + np.abs(dirs[:, 1].astype(np.float64))
    pa = np.where(den > 0, pang(dirs[:, 0].astype(np.float64), dirs[:, 1].astype(np.float64)), 0.0)
    bins = np.clip(np.floor(pa * (BINS / 4)).astype(int), 0, BINS - 1)
    return dirs[None, :, 2] > hz[:, bins]


@pytest.mark.parametrize("knobs", [dict(), dict(budget=0), dict(near=60, budget=4), dict(near=12, budget=256), dict(near=30, slabs=0, mid=0),
                                   dict(near=30, mid=0), dict(slabs=0), dict(mid=40, gain=0.01, budget=200), dict(near=157, mid=5, gain=1.5)])
def test_horizon_map_never_frees_an_occluded_ray(hostcheck, oracle, knobs):
    """build_entry_list + build_horizon (the product's device code, unmodified, on the warp emulator) against the oracle's
    per-ray visibility on a self-occluding mesh: every freed sample must be visible -- and the map must be worth having."""
    from prt_b200 import meshes
    pos, nrm, tri = meshes.bumpy_torus(96, 64)
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    try:
        sel = np.arange(3, len(pos), 41)[:150]
        hz, ncand = _maps(hostcheck, h, pos[sel], nrm[sel], **knobs)
    finally:
        hostcheck.hc_free(h)
    assert (ncand > 0).all() and (ncand <= 96).all()
    assert np.isfinite(hz).all() and (hz >= 0).all() and (hz <= 2.0).all()
    op = oracle.make_params(order=3, samples_u=32, samples_v=32)
    _, vis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), pos[sel], nrm[sel], op, want_vis=True)
    visible = np.unpackbits(vis.view(np.uint8), axis=1, bitorder="little")[:, :1024].astype(bool)
    _, dirs = oracle.sample_table(op)
    free = _free_mask(dirs, hz)
    assert not (free & ~visible).any(), "the horizon pass would free an occluded ray"
    occluded = 1.0 - visible.mean()
    traversed = 1.0 - free.mean()
    assert 0.05 < occluded < 0.6
    # tightness (regression guard, not a correctness bar): the default builder leaves fewer than 2.2x the occluded share to trace
    if not knobs:
        assert traversed < 2.2 * occluded + 0.05, (traversed, occluded)
    assert traversed >= occluded - 1e-9


def test_horizon_map_adversarial_geometry(hostcheck, oracle):
    """thin sliver at grazing elevation, an overhang, a wall crossing the tangent plane and a far big occluder around one vertex"""
    rng = np.random.RandomState(5)
    base = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32) * 4
    verts = [base]
    tris = [np.array([[0, 1, 2], [0, 2, 3]], np.uint32)]

    def add(tv):
        off = sum(len(v) for v in verts)
        verts.append(np.asarray(tv, np.float32))
        tris.append(np.array([[0, 1, 2]], np.uint32) + off)
    add([[1.0, -2.0, 0.02], [1.0, 2.0, 0.02], [1.0005, 0.0, 0.06]])            # sliver just above the plane
    add([[-0.5, -0.5, 0.8], [0.5, -0.5, 0.8], [0.0, 0.6, 0.8]])                # overhang above the vertex
    add([[0.3, 0.4, -0.5], [0.9, 0.4, 0.7], [0.3, 1.0, 0.7]])                  # crosses the tangent plane
    add([[-30.0, -40.0, 1.0], [-30.0, 40.0, 1.0], [-30.0, 0.0, 25.0]])         # far, large
    for _ in range(40):                                                        # clutter so that the BVH has inner nodes
        c = rng.uniform(-3, 3, 3); c[2] = rng.uniform(0.05, 2.0)
        add(c + rng.normal(size=(3, 3)) * 0.15)
    pos = np.concatenate(verts).astype(np.float32)
    tri = np.concatenate(tris).astype(np.uint32)
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    p = np.array([[0.0, 0.0, 0.0], [0.5, 0.2, 0.0], [-1.0, 1.0, 0.0]], np.float32)
    n = np.array([[0, 0, 1]] * 3, np.float32)
    op = oracle.make_params(order=3, samples_u=32, samples_v=32)
    _, dirs = oracle.sample_table(op)
    _, vis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), p, n, op, want_vis=True)
    visible = np.unpackbits(vis.view(np.uint8), axis=1, bitorder="little")[:, :1024].astype(bool)
    try:
        for knobs in (dict(), dict(budget=0), dict(near=90, budget=1), dict(near=10, budget=256), dict(near=30, slabs=0, mid=0), dict(mid=40, gain=0.0, budget=256)):
            hz, _ = _maps(hostcheck, h, p, n, **knobs)
            free = _free_mask(dirs, hz)
            assert not (free & ~visible).any(), knobs
    finally:
        hostcheck.hc_free(h)
    assert 0.02 < 1.0 - visible.mean() < 0.9


@pytest.mark.parametrize("scene_kind", ["sphere_in_sphere", "triangle_soup"])
def test_horizon_map_other_scenes(hostcheck, oracle, scene_kind):
    """the same end-to-end check on geometry that is not a smooth surface seen from itself: a small sphere inside a big inverted
    one (everything is far field, every ray is occluded or escapes through nothing) and an unstructured triangle soup with
    arbitrary vertex normals (origins inside bounding boxes, geometry on both sides of every tangent plane)"""
    from prt_b200 import meshes
    rng = np.random.RandomState(11)
    if scene_kind == "sphere_in_sphere":
        p0, n0, t0 = meshes.icosphere(3)
        p1, n1, t1 = meshes.icosphere(2)
        pos = np.concatenate([p0 * 0.5, p1 * 3.0]).astype(np.float32)
        nrm = np.concatenate([n0, -n1]).astype(np.float32)
        tri = np.concatenate([t0, t1[:, ::-1] + len(p0)]).astype(np.uint32)
        sel = np.concatenate([np.arange(0, len(p0), 9)[:40], len(p0) + np.arange(0, len(p1), 5)[:20]])
    else:
        k = 600
        c = rng.uniform(-2, 2, size=(k, 1, 3))
        pos = (c + rng.normal(size=(k, 3, 3)) * 0.25).reshape(-1, 3).astype(np.float32)
        tri = np.arange(3 * k, dtype=np.uint32).reshape(k, 3)
        nrm = rng.normal(size=(3 * k, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        sel = np.arange(0, 3 * k, 31)[:60]
    h = hostcheck.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
    assert h
    op = oracle.make_params(order=3, samples_u=32, samples_v=32)
    _, dirs = oracle.sample_table(op)
    _, vis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), pos[sel], nrm[sel], op, want_vis=True)
    visible = np.unpackbits(vis.view(np.uint8), axis=1, bitorder="little")[:, :1024].astype(bool)
    try:
        for knobs in (dict(), dict(near=15, budget=200), dict(budget=1), dict(near=30, slabs=0, mid=0), dict(mid=40, gain=0.0, budget=256)):
            hz, _ = _maps(hostcheck, h, pos[sel], nrm[sel], **knobs)
            free = _free_mask(dirs, hz)
            assert not (free & ~visible).any(), (scene_kind, knobs)
    finally:
        hostcheck.hc_free(h)
    if scene_kind == "sphere_in_sphere":
        assert visible.mean() < 0.05            # a closed room: (almost) nothing escapes
