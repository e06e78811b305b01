"""ctypes binding of the CPU oracle (oracle/libprt_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package prt_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

UNSHADOWED, SHADOWED, INTERREFLECT, UNSHADOWED_ANALYTIC = 0, 1, 2, 3


class BakeParams(C.Structure):
    _fields_ = [("order", C.c_int32), ("samples_u", C.c_int32), ("samples_v", C.c_int32), ("seed", C.c_uint32),
                ("bounces", C.c_int32), ("albedo", C.c_float * 3), ("origin_eps", C.c_float),
                ("bounce_eps", C.c_float), ("mode", C.c_int32), ("cs_phase", C.c_int32), ("jitter", C.c_int32)]


def make_params(order=3, samples_u=32, samples_v=32, seed=0x50525400, bounces=0, albedo=(1.0, 1.0, 1.0),
                origin_eps=1e-4, bounce_eps=1e-5, mode=SHADOWED, cs_phase=0, jitter=1) -> BakeParams:
    p = BakeParams()
    p.order, p.samples_u, p.samples_v, p.seed, p.bounces = order, samples_u, samples_v, seed, bounces
    p.albedo[:] = albedo
    p.origin_eps, p.bounce_eps, p.mode, p.cs_phase, p.jitter = origin_eps, bounce_eps, mode, cs_phase, jitter
    return p


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "libprt_oracle.so")
    if force or not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libprt_oracle.so"])
    return path


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, f32p, u32p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L.prt_o_scene_create.restype = vp
        L.prt_o_scene_create.argtypes = [vp, C.c_size_t, C.c_uint32, vp, C.c_uint32]
        L.prt_o_scene_destroy.argtypes = [vp]
        L.prt_o_any_hit.restype = C.c_int
        L.prt_o_any_hit.argtypes = [vp, f32p, f32p, C.c_float, C.c_float, C.c_int]
        L.prt_o_closest_hit.restype = C.c_int
        L.prt_o_closest_hit.argtypes = [vp, f32p, f32p, C.c_float, C.c_float, C.c_int, f32p, u32p, f32p]
        L.prt_o_sample_table.argtypes = [C.POINTER(BakeParams), vp, vp]
        L.prt_o_bake_transfer.restype = C.c_int
        L.prt_o_bake_transfer.argtypes = [vp, vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.POINTER(BakeParams),
                                          vp, vp, C.c_int, C.c_int, vp]
        L.prt_o_bake_transfer_ref_order.restype = C.c_uint64
        L.prt_o_bake_transfer_ref_order.argtypes = [vp, vp, vp, C.c_size_t, C.c_uint32, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_uint64, C.c_int,
                                                    C.c_uint32, C.c_uint32, vp]
        L.prt_o_sh_eval.argtypes = [C.c_int, C.c_int, f32p, f32p]
        L.prt_o_philox.argtypes = [u32p, u32p, u32p]
        L.prt_o_sincos2pi.argtypes = [C.c_float, f32p, f32p]
        L.prt_o_hw_threads.restype = C.c_int
        L.prt_o_cosine_world.argtypes = [C.c_float, C.c_float, f32p, f32p, f32p, f32p, f32p]
        L.prt_o_equirect_uv.argtypes = [f32p, f32p]
        L.prt_o_env_irradiance_dir.argtypes = [vp, C.c_int, C.c_int, f32p, f32p]
        L.prt_o_env_prefilter_dir.argtypes = [vp, C.c_int, C.c_int, f32p, C.c_float, C.c_int, f32p]
        L.prt_o_cube_floats.restype = C.c_size_t
        L.prt_o_cube_floats.argtypes = [C.c_int, C.c_int]
        L.prt_o_cube_levels.restype = C.c_int
        L.prt_o_cube_levels.argtypes = [C.c_int]
        L.prt_o_cube_sample.argtypes = [vp, C.c_int, C.c_int, f32p, C.c_float, f32p]
        L.prt_o_env_equirect_to_cube.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.prt_o_env_irradiance.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.prt_o_env_prefilter.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.prt_o_brdf_lut.argtypes = [C.c_int, C.c_int, C.c_int, vp]
        L.prt_o_env_project_sh.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.prt_o_sh_pack_rh.argtypes = [vp, vp]
        L.prt_o_fibonacci_dirs.argtypes = [C.c_int, vp]
        L.prt_o_cube_dirs.argtypes = [C.c_int, vp, vp]
        L.prt_o_probe_positions.argtypes = [vp, vp, vp]
        L.prt_o_probe_capture.restype = vp
        L.prt_o_probe_capture.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_uint32]
        L.prt_o_csr_sizes.argtypes = [vp, u32p, u32p]
        L.prt_o_csr_get.argtypes = [vp, vp, vp, vp, vp, vp]
        L.prt_o_csr_get_sums.argtypes = [vp, vp]
        L.prt_o_csr_destroy.argtypes = [vp]
        L.prt_o_probe_project.argtypes = [vp, vp, vp]
        L.prt_o_project_arrays.argtypes = [vp, C.c_uint32, vp, vp, vp, vp]
        L.prt_o_volume_weights.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.prt_o_paral_shadow_matrix.argtypes = [C.c_float, C.c_float, vp, vp]
        L.prt_o_shadow_map.argtypes = [vp, vp, C.c_int, vp]
        L.prt_o_relight.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_int, vp, vp, vp, vp]
        L.prt_o_transfer_to_volume.argtypes = [vp, vp, vp, vp, vp, vp]
        L.prt_o_raytrace.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_uint32, C.c_uint32, vp, vp]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Scene:
    """RTScene (reference raytracing.cpp:58-99) on the CPU oracle."""

    def __init__(self, pos: np.ndarray, tri: np.ndarray):
        self.pos = np.ascontiguousarray(pos, dtype=np.float32)
        self.tri = np.ascontiguousarray(tri, dtype=np.uint32)
        self.h = lib().prt_o_scene_create(_ptr(self.pos), 12, len(self.pos), _ptr(self.tri), len(self.tri))
        if not self.h:
            raise RuntimeError("oracle scene_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().prt_o_scene_destroy(self.h)
            self.h = None

    def any_hit(self, org, dirs, tnear=0.0, tfar=np.inf, use_bvh=True) -> np.ndarray:
        org = np.ascontiguousarray(np.broadcast_to(np.asarray(org, np.float32), np.asarray(dirs).shape), np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        out = np.zeros(len(dirs), np.int32)
        f32p = C.POINTER(C.c_float)
        L = lib()
        for i in range(len(dirs)):
            out[i] = L.prt_o_any_hit(self.h, org[i].ctypes.data_as(f32p), dirs[i].ctypes.data_as(f32p),
                                     C.c_float(tnear), C.c_float(tfar), int(use_bvh))
        return out

    def closest_hit(self, org, dirs, tnear=0.0, tfar=np.inf, use_bvh=True):
        org = np.ascontiguousarray(np.broadcast_to(np.asarray(org, np.float32), np.asarray(dirs).shape), np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        n = len(dirs)
        hit, t, prim, ng = np.zeros(n, np.int32), np.full(n, np.inf, np.float32), np.full(n, 0xFFFFFFFF, np.uint32), np.zeros((n, 3), np.float32)
        f32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L = lib()
        for i in range(n):
            tt, pp = C.c_float(), C.c_uint32()
            hit[i] = L.prt_o_closest_hit(self.h, org[i].ctypes.data_as(f32p), dirs[i].ctypes.data_as(f32p), C.c_float(tnear),
                                         C.c_float(tfar), int(use_bvh), C.byref(tt), C.byref(pp), ng[i].ctypes.data_as(f32p))
            if hit[i]:
                t[i], prim[i] = tt.value, pp.value
        return hit, t, prim, ng


def sample_table(params: BakeParams):
    S = params.samples_u * params.samples_v
    uv, dirs = np.zeros((S, 2), np.float32), np.zeros((S, 3), np.float32)
    lib().prt_o_sample_table(C.byref(params), _ptr(uv), _ptr(dirs))
    return uv, dirs


def bake_transfer(scene: Scene | None, pos, nrm, params: BakeParams, want_vis=False, vertex_id_base=0,
                  n_threads=0, faithful=False):
    """bake_SH (reference raytracing.cpp:320-360). Returns (coeffs[n, order^2], vis_words or None, counters)."""
    pos = np.ascontiguousarray(pos, np.float32)
    nrm = np.ascontiguousarray(nrm, np.float32)
    n = len(pos)
    n2 = params.order ** 2
    S = params.samples_u * params.samples_v
    out = np.zeros((n, n2), np.float32)
    vis = np.zeros((n, (S + 31) // 32), np.uint32) if want_vis else None
    counters = np.zeros(2, np.uint64)
    rc = lib().prt_o_bake_transfer(scene.h if scene is not None else None, _ptr(pos), _ptr(nrm), 12, n, vertex_id_base,
                                   C.byref(params), _ptr(out), _ptr(vis), n_threads, int(faithful), _ptr(counters))
    if rc != 0:
        raise RuntimeError(f"oracle bake_transfer failed rc={rc}")
    return out, vis, counters


def mt19937_floats(n: int) -> np.ndarray:
    """The first n values of the reference's ``random()`` (raytracing.cpp:14-18): a default-seeded std::mt19937 (seed 5489; numpy's
    legacy RandomState(5489) is the same init_genrand + genrand_int32) through libstdc++'s uniform_real_distribution<float>(0, 1) =
    generate_canonical<float, 24>: float(u32) / 2^32, a value that rounds up to 1 replaced by the float below 1."""
    u = np.random.RandomState(5489).randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    f = u.astype(np.float32) / np.float32(4294967296.0)
    f[f >= 1] = np.nextafter(np.float32(1), np.float32(0))
    return f


def bake_transfer_ref_order(scene: Scene, pos, nrm, order=3, res=32, max_path_length=2, albedo=(1.0, 1.0, 1.0), cs_phase=1, rnd=None, u_first=0,
                            seed=0x50525400, vertex_id_base=0):
    """bake_SH exactly in the reference's loop order (oracle/bake.c prt_o_bake_transfer_ref_order).  rnd: the random() sequence to
    consume (``mt19937_floats``), or None for the production oracle's Philox draws.  u_first = 0 is what g++ makes of
    ``cosineSampleHemisphere(random(), random(), normal)`` (arguments evaluated right to left).  -> (rows, values consumed)"""
    pos = np.ascontiguousarray(pos, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
    out = np.zeros((len(pos), order * order), np.float32)
    a = np.asarray(albedo, np.float32)
    r = None if rnd is None else np.ascontiguousarray(rnd, np.float32)
    used = lib().prt_o_bake_transfer_ref_order(scene.h, _ptr(pos), _ptr(nrm), 12, len(pos), order, res, max_path_length, _ptr(a), cs_phase,
                                               _ptr(r), 0 if r is None else len(r), u_first, seed & 0xFFFFFFFF, vertex_id_base, _ptr(out))
    return out, int(used)


def sh_eval(order: int, d_sh, cs_phase=0) -> np.ndarray:
    d = np.asarray(d_sh, np.float32)
    out = np.zeros(order * order, np.float32)
    f32p = C.POINTER(C.c_float)
    lib().prt_o_sh_eval(order, cs_phase, d.ctypes.data_as(f32p), out.ctypes.data_as(f32p))
    return out


def philox(ctr, key) -> np.ndarray:
    c, k, o = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
    u32p = C.POINTER(C.c_uint32)
    lib().prt_o_philox(c.ctypes.data_as(u32p), k.ctypes.data_as(u32p), o.ctypes.data_as(u32p))
    return o


def sincos2pi(v: float):
    s, c = C.c_float(), C.c_float()
    lib().prt_o_sincos2pi(C.c_float(v), C.byref(s), C.byref(c))
    return s.value, c.value


def cosine_world(u: float, v: float, N):
    """frame(N) + cosineSampleHemisphere(u, v, N) (reference raytracing.cpp:101-107,130-160) -> (local, world, frame [3,3] rows =
    right, up, N, pdf)."""
    f32p = C.POINTER(C.c_float)
    n = np.asarray(N, np.float32); l = np.zeros(3, np.float32); w = np.zeros(3, np.float32); fr = np.zeros(9, np.float32); pdf = C.c_float()
    lib().prt_o_cosine_world(C.c_float(u), C.c_float(v), n.ctypes.data_as(f32p), l.ctypes.data_as(f32p), w.ctypes.data_as(f32p),
                             fr.ctypes.data_as(f32p), C.byref(pdf))
    return l, w, fr.reshape(3, 3), pdf.value


def equirect_uv(d) -> np.ndarray:
    """SampleSphericalMap (reference rectangle2cube.frag:7-15) of a normalised direction."""
    f32p = C.POINTER(C.c_float)
    v = np.asarray(d, np.float32); uv = np.zeros(2, np.float32)
    lib().prt_o_equirect_uv(v.ctypes.data_as(f32p), uv.ctypes.data_as(f32p))
    return uv


def hw_threads() -> int:
    return lib().prt_o_hw_threads()


class EnvCube:
    """oracle twin of prt_b200.LightProbe (oracle/env.c)."""

    def __init__(self, equirect: np.ndarray, cube_size: int = 512):
        eq = np.ascontiguousarray(equirect, np.float32)
        self.n0 = cube_size
        self.levels = lib().prt_o_cube_levels(cube_size)
        self.data = np.zeros(lib().prt_o_cube_floats(cube_size, self.levels), np.float32)
        lib().prt_o_env_equirect_to_cube(_ptr(eq), eq.shape[1], eq.shape[0], cube_size, self.levels, _ptr(self.data))

    def cube(self, level=0) -> np.ndarray:
        o = lib().prt_o_cube_floats(self.n0, level)
        n = self.n0 >> level
        return self.data[o:o + 6 * n * n * 3].reshape(6, n, n, 3)

    def sample(self, d, lod=0.0) -> np.ndarray:
        d = np.asarray(d, np.float32)
        out = np.zeros(3, np.float32)
        f32p = C.POINTER(C.c_float)
        lib().prt_o_cube_sample(_ptr(self.data), self.n0, self.levels, d.ctypes.data_as(f32p), C.c_float(lod), out.ctypes.data_as(f32p))
        return out

    def irradiance_dir(self, P) -> np.ndarray:
        """irradiance.frag main() for one fragment, P = CubeTexPos."""
        f32p = C.POINTER(C.c_float)
        p = np.asarray(P, np.float32); out = np.zeros(3, np.float32)
        lib().prt_o_env_irradiance_dir(_ptr(self.data), self.n0, self.levels, p.ctypes.data_as(f32p), out.ctypes.data_as(f32p))
        return out

    def prefilter_dir(self, P, roughness: float, n_samples: int = 1024) -> np.ndarray:
        """prefilter.frag main() for one fragment, P = CubeTexPos."""
        f32p = C.POINTER(C.c_float)
        p = np.asarray(P, np.float32); out = np.zeros(3, np.float32)
        lib().prt_o_env_prefilter_dir(_ptr(self.data), self.n0, self.levels, p.ctypes.data_as(f32p), C.c_float(roughness), n_samples,
                                      out.ctypes.data_as(f32p))
        return out

    def irradiance(self, n_out=32) -> np.ndarray:
        out = np.zeros((6, n_out, n_out, 3), np.float32)
        lib().prt_o_env_irradiance(_ptr(self.data), self.n0, self.levels, n_out, _ptr(out))
        return out

    def prefilter(self, n_out=256, mips=5, n_samples=1024) -> list:
        sizes = [n_out >> m for m in range(mips)]
        flat = np.zeros(sum(6 * n * n * 3 for n in sizes), np.float32)
        lib().prt_o_env_prefilter(_ptr(self.data), self.n0, self.levels, n_out, mips, n_samples, _ptr(flat))
        out, o = [], 0
        for n in sizes:
            out.append(flat[o:o + 6 * n * n * 3].reshape(6, n, n, 3))
            o += 6 * n * n * 3
        return out

    def project_sh(self, order=3, method=0, size=None) -> np.ndarray:
        size = size or (256 if method == 0 else 64)
        out = np.zeros((order * order, 3), np.float32)
        lib().prt_o_env_project_sh(_ptr(self.data), self.n0, self.levels, order, method, size, _ptr(out))
        return out


def brdf_lut(w=512, h=512, n_samples=1024) -> np.ndarray:
    out = np.zeros((h, w, 2), np.float32)
    lib().prt_o_brdf_lut(w, h, n_samples, _ptr(out))
    return out


def sh_pack_rh(L9) -> np.ndarray:
    L9 = np.ascontiguousarray(L9, np.float32)
    out = np.zeros(28, np.float32)
    lib().prt_o_sh_pack_rh(_ptr(L9), _ptr(out))
    return out


def probe_positions(res, scene_size) -> np.ndarray:
    res = np.asarray(res, np.int32); size = np.asarray(scene_size, np.float32)
    out = np.zeros((int(np.prod(res)), 3), np.float32)
    lib().prt_o_probe_positions(_ptr(res), _ptr(size), _ptr(out))
    return out


def fibonacci_dirs(n):
    d = np.zeros((n, 3), np.float32)
    lib().prt_o_fibonacci_dirs(n, _ptr(d))
    return d, np.full(n, 4 * np.pi / n, np.float32)


def cube_dirs(res):
    d, w = np.zeros((6 * res * res, 3), np.float32), np.zeros(6 * res * res, np.float32)
    lib().prt_o_cube_dirs(res, _ptr(d), _ptr(w))
    return d, w


class ProbeTransfer:
    """oracle twin of prt_b200.ProbeTransfer (oracle/probe.c)."""

    def __init__(self, scene: Scene, probe_pos, dirs, weights):
        pp = np.ascontiguousarray(probe_pos, np.float32); d = np.ascontiguousarray(dirs, np.float32); w = np.ascontiguousarray(weights, np.float32)
        self.h = lib().prt_o_probe_capture(scene.h, _ptr(pp), len(pp), _ptr(d), _ptr(w), len(d))
        nnz, ns = C.c_uint32(), C.c_uint32()
        lib().prt_o_csr_sizes(self.h, C.byref(nnz), C.byref(ns))
        self.n_probes, self.nnz, self.n_surfels = len(pp), nnz.value, ns.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().prt_o_csr_destroy(self.h)
            self.h = None

    def download(self):
        rng = np.zeros((self.n_probes, 2), np.uint32); ids = np.zeros(self.nnz, np.uint32); tr = np.zeros((self.nnz, 9), np.float32)
        sf = np.zeros((self.n_surfels, 6), np.float32); keys = np.zeros(self.n_surfels, np.uint64)
        lib().prt_o_csr_get(self.h, _ptr(rng), _ptr(ids), _ptr(tr), _ptr(sf), _ptr(keys))
        return rng, ids, tr, sf, keys

    def surfel_sums(self) -> np.ndarray:
        sums = np.zeros((self.n_surfels, 7), np.float64)
        lib().prt_o_csr_get_sums(self.h, _ptr(sums))
        return sums

    def project(self, radiance_rgba) -> np.ndarray:
        rad = np.ascontiguousarray(radiance_rgba, np.float32)
        out = np.zeros((self.n_probes, 7, 4), np.float32)
        lib().prt_o_probe_project(self.h, _ptr(rad), _ptr(out))
        return out


def project_arrays(rng, ids, transfer, radiance_rgba) -> np.ndarray:
    """SH_volume::project_sh's kernel (precomp_projectSH.comp) on a CSR given as arrays -> [n_probes, 7, 4]."""
    rng = np.ascontiguousarray(rng, np.uint32); ids = np.ascontiguousarray(ids, np.uint32)
    tr = np.ascontiguousarray(transfer, np.float32); rad = np.ascontiguousarray(radiance_rgba, np.float32)
    out = np.zeros((len(rng), 7, 4), np.float32)
    lib().prt_o_project_arrays(_ptr(rng), len(rng), _ptr(ids), _ptr(tr), _ptr(rad), _ptr(out))
    return out


def volume_weights(scene: Scene, probe_res, volume_res, scene_size):
    """calculate_weight (reference light_probe.cpp:156-367) -> (weight0123, weight4567, inside_score)."""
    pr = np.asarray(probe_res, np.int32); vr = np.asarray(volume_res, np.int32); sz = np.asarray(scene_size, np.float32)
    n = int(np.prod(vr))
    w0, w1, sc = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros(n, np.float32)
    lib().prt_o_volume_weights(scene.h, _ptr(pr), _ptr(vr), _ptr(sz), _ptr(w0), _ptr(w1), _ptr(sc))
    return w0, w1, sc


def paral_shadow_matrix(up: float, direction: float):
    """Paral_Shadow::set_dir (reference gl.cpp:620-631)."""
    d = np.zeros(3, np.float32); m = np.zeros(16, np.float32)
    lib().prt_o_paral_shadow_matrix(float(up), float(direction), _ptr(d), _ptr(m))
    return d, m


def shadow_map(scene: Scene, matrix, size: int) -> np.ndarray:
    m = np.ascontiguousarray(np.asarray(matrix, np.float32).reshape(16))
    out = np.zeros((size, size), np.float32)
    if lib().prt_o_shadow_map(scene.h, _ptr(m), size, _ptr(out)) != 0:
        raise RuntimeError("oracle shadow_map: matrix must be affine")
    return out


def relight(params, surfels, radiance, albedo=None, depth=None, volumes=None, volume_res=None, scene_size=None) -> np.ndarray:
    """relight.comp:68-82 on a surfel table; ``params`` is any ctypes struct laid out like prt_o_relight_params
    (prt_b200.RelightParams is).  Returns the new radiance [n,4]."""
    sf = np.ascontiguousarray(surfels, np.float32); rad = np.array(radiance, np.float32, copy=True, order="C")
    alb = None if albedo is None else np.ascontiguousarray(albedo, np.float32)
    dep = None if depth is None else np.ascontiguousarray(depth, np.float32)
    vol = None if volumes is None else np.ascontiguousarray(volumes, np.float32)
    vr = np.asarray(volume_res if volume_res is not None else [1, 1, 1], np.int32)
    sz = np.asarray(scene_size if scene_size is not None else [1, 1, 1], np.float32)
    lib().prt_o_relight(C.cast(C.pointer(params), C.c_void_p), len(sf), _ptr(sf), _ptr(alb), _ptr(dep), 0 if dep is None else dep.shape[0],
                        _ptr(vol), _ptr(vr), _ptr(sz), _ptr(rad))
    return rad


def transfer_to_volume(probe_sh, probe_res, w0123, w4567, volume_res) -> np.ndarray:
    """transfer2volume.comp:36-147 -> [n_voxels, 7, 4]"""
    ps = np.ascontiguousarray(probe_sh, np.float32); pr = np.asarray(probe_res, np.int32); vr = np.asarray(volume_res, np.int32)
    w0 = np.ascontiguousarray(w0123, np.float32); w1 = np.ascontiguousarray(w4567, np.float32)
    out = np.zeros((int(np.prod(vr)), 7, 4), np.float32)
    lib().prt_o_transfer_to_volume(_ptr(ps), _ptr(pr), _ptr(w0), _ptr(w1), _ptr(vr), _ptr(out))
    return out


def raytrace(scene: Scene, camera, w: int, h: int, accum=None, max_path_length=3, albedo=(1.0, 1.0, 1.0), gamma=True, mode=0,
             seed=0x50525400, frame=0):
    """raytrace() (reference raytracing.cpp:280-317), one frame; ``camera`` is a ctypes struct laid out like prt_o_camera
    (prt_b200.Camera is).  Returns (accum [h,w,4], pixels [h,w,4] uint8)."""
    acc = np.zeros((h, w, 4), np.float32) if accum is None else np.array(accum, np.float32, copy=True, order="C")
    px = np.zeros((h, w, 4), np.uint8)
    a = np.asarray(albedo, np.float32)
    lib().prt_o_raytrace(scene.h, C.cast(C.pointer(camera), C.c_void_p), w, h, int(max_path_length), _ptr(a), int(bool(gamma)), int(mode),
                         int(seed) & 0xFFFFFFFF, int(frame), _ptr(acc), _ptr(px))
    return acc, px
