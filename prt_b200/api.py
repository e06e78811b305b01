"""ctypes binding of libprt_b200.so (include/prt_b200.h) and the reference-shaped host interface."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

UNSHADOWED, SHADOWED, INTERREFLECT, UNSHADOWED_ANALYTIC = 0, 1, 2, 3

# every symbol include/prt_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "prt_last_error", "prt_abi_version", "prt_ctx_create", "prt_ctx_destroy", "prt_ctx_device", "prt_ctx_last_kernel_ms", "prt_ctx_set_tuning",
    "prt_scene_create", "prt_scene_destroy", "prt_scene_get_info", "prt_trace_any_hit", "prt_trace_closest_hit",
    "prt_bake_params_default", "prt_bake_transfer", "prt_bake_transfer_device", "prt_bake_transfer_device_strided", "prt_bake_transfer_device_shard", "prt_bake_transfer_device_shard_fused",
    "prt_device_alloc", "prt_device_free", "prt_ipc_export", "prt_ipc_open", "prt_ipc_close", "prt_scatter_sh9",
    "prt_bake_sample_table", "prt_ctx_last_bake_stats",
    "prt_env_create", "prt_env_destroy", "prt_env_levels", "prt_env_get_cube", "prt_env_irradiance", "prt_env_prefilter",
    "prt_brdf_lut", "prt_env_project_sh", "prt_sh_pack_rh",
    "prt_probe_capture", "prt_csr_destroy", "prt_csr_sizes", "prt_csr_download", "prt_csr_surfel_sums", "prt_csr_upload", "prt_probe_project", "prt_probe_positions",
    "prt_fibonacci_dirs", "prt_cube_dirs", "prt_volume_weights",
    "prt_paral_shadow_matrix", "prt_shadow_map", "prt_gi_create", "prt_gi_destroy", "prt_gi_set_shadow_map", "prt_gi_set_albedo",
    "prt_gi_set_radiance", "prt_gi_step", "prt_gi_download",
    "prt_film_create", "prt_film_destroy", "prt_film_reset", "prt_raytrace", "prt_film_download",
    "prt_hash_bytes", "prt_mesh_hash", "prt_cache_save_transfer", "prt_cache_load_transfer", "prt_cache_save_csr", "prt_cache_csr_sizes",
    "prt_cache_load_csr",
    "prt_group_create", "prt_group_destroy", "prt_group_size", "prt_group_ctx", "prt_group_capabilities", "prt_group_set_tuning",
    "prt_group_scene_create", "prt_group_scene_destroy", "prt_group_scene_get_info", "prt_group_scene_member", "prt_group_bake_transfer",
    "prt_group_rows_device", "prt_group_download_rows", "prt_group_probe_capture",
]


class PRTError(RuntimeError):
    pass


class BakeParams(C.Structure):
    """prt_bake_params: the App singleton fields bake_SH reads (reference app.h:55,70-71; raytracing.cpp:320,343)."""
    _fields_ = [("order", C.c_int32), ("samples_u", C.c_int32), ("samples_v", C.c_int32), ("seed", C.c_uint32),
                ("bounces", C.c_int32), ("albedo", C.c_float * 3), ("origin_eps", C.c_float),
                ("bounce_eps", C.c_float), ("mode", C.c_int32), ("cs_phase", C.c_int32), ("jitter", C.c_int32)]

    @classmethod
    def make(cls, order=3, samples_u=32, samples_v=32, seed=0x50525400, bounces=0, albedo=(1.0, 1.0, 1.0),
             origin_eps=1e-4, bounce_eps=1e-5, mode=SHADOWED, cs_phase=0, jitter=1) -> "BakeParams":
        p = cls()
        p.order, p.samples_u, p.samples_v, p.seed, p.bounces = order, samples_u, samples_v, seed, bounces
        p.albedo[:] = albedo
        p.origin_eps, p.bounce_eps, p.mode, p.cs_phase, p.jitter = origin_eps, bounce_eps, mode, cs_phase, jitter
        return p

    @property
    def n_samples(self) -> int:
        return self.samples_u * self.samples_v

    @property
    def n_coeffs(self) -> int:
        return self.order * self.order


class RelightParams(C.Structure):
    """``prt_relight_params``: the uniforms SH_volume::relight sets (reference src/sh/volume.cpp:362-377)."""
    _fields_ = [("cast_intensity", C.c_float * 3), ("cast_position", C.c_float * 3), ("cast_direction", C.c_float * 3), ("cast_cutoff", C.c_float),
                ("ambient_intensity", C.c_float * 3), ("ambient_position", C.c_float * 3),
                ("sky_intensity", C.c_float * 3), ("sky_direction", C.c_float * 3), ("light_space_matrix", C.c_float * 16),
                ("multi_bounce", C.c_int32), ("atten", C.c_float), ("sh_shift", C.c_float), ("temp_weight", C.c_float)]

    @classmethod
    def make(cls, sky_direction, light_space_matrix, sky_intensity=(5, 5, 5), cast_intensity=100.0, cast_position=(4, 0, 0),
             cast_direction=(1, 0, 0), cast_cutoff=0.9, ambient_intensity=(0, 0, 0), ambient_position=(0, 0, 0),
             multi_bounce=True, atten=1.0, sh_shift=0.0, temp_weight=0.1):
        """Defaults are the reference's (app.h:31-37,67-68; volume.cpp:362-367; relight.comp:12)."""
        p = cls()
        ci = (cast_intensity,) * 3 if np.isscalar(cast_intensity) else cast_intensity
        p.cast_intensity[:] = [float(x) for x in ci]; p.cast_position[:] = [float(x) for x in cast_position]
        p.cast_direction[:] = [float(x) for x in cast_direction]; p.cast_cutoff = float(cast_cutoff)
        p.ambient_intensity[:] = [float(x) for x in ambient_intensity]; p.ambient_position[:] = [float(x) for x in ambient_position]
        p.sky_intensity[:] = [float(x) for x in sky_intensity]; p.sky_direction[:] = [float(x) for x in sky_direction]
        p.light_space_matrix[:] = [float(x) for x in np.asarray(light_space_matrix, np.float32).reshape(-1)]
        p.multi_bounce = int(bool(multi_bounce)); p.atten = float(atten); p.sh_shift = float(sh_shift); p.temp_weight = float(temp_weight)
        return p


class Camera(C.Structure):
    """``prt_camera``: the fields of the reference's Camera that raytrace() reads (raytracing.cpp:293-299)."""
    _fields_ = [("position", C.c_float * 3), ("front", C.c_float * 3), ("up", C.c_float * 3), ("right", C.c_float * 3), ("zoom_deg", C.c_float)]

    @classmethod
    def look_at(cls, position, target, world_up=(0, 1, 0), zoom_deg=45.0):
        p = np.asarray(position, np.float64); f = np.asarray(target, np.float64) - p
        f /= np.linalg.norm(f)
        r = np.cross(f, np.asarray(world_up, np.float64)); r /= np.linalg.norm(r)
        u = np.cross(r, f)
        c = cls()
        c.position[:] = [float(x) for x in p]; c.front[:] = [float(x) for x in f]; c.up[:] = [float(x) for x in u]; c.right[:] = [float(x) for x in r]
        c.zoom_deg = float(zoom_deg)
        return c


class SceneInfo(C.Structure):
    _fields_ = [("n_tris", C.c_uint32), ("n_nodes", C.c_uint32), ("max_depth", C.c_uint32), ("reserved", C.c_uint32),
                ("node_bytes", C.c_uint64), ("tri_bytes", C.c_uint64), ("build_seconds", C.c_double),
                ("upload_seconds", C.c_double), ("sah_cost", C.c_double)]


class BakeStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("rays", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("launches", C.c_uint32), ("grid", C.c_uint32),
                ("block", C.c_uint32), ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64), ("cand_tests", C.c_uint64), ("rays_traversed", C.c_uint64), ("horizon_ms", C.c_double)]


class GroupStats(C.Structure):
    _fields_ = [("n_devices", C.c_uint32), ("gather_mode", C.c_int32), ("wall_ms", C.c_double), ("kernel_ms_max", C.c_double),
                ("gather_ms_max", C.c_double), ("h2d_ms", C.c_double * 8), ("kernel_ms", C.c_double * 8), ("gather_ms", C.c_double * 8),
                ("vertices", C.c_uint32 * 8), ("gather_bytes_per_gpu", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


GATHER_NONE, GATHER_NCCL, GATHER_P2P, GATHER_AUTO = 0, 1, 2, -1


def lib_path() -> str:
    # PRT_B200_LIB selects an alternative in-tree build (kernel tuning experiments)
    return os.environ.get("PRT_B200_LIB") or os.path.join(_HERE, "csrc", "libprt_b200.so")


def load_library():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise PRTError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
    L.prt_last_error.restype = C.c_char_p
    L.prt_abi_version.restype = i32
    L.prt_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.prt_ctx_destroy.argtypes = [vp]
    L.prt_ctx_destroy.restype = None
    L.prt_ctx_device.argtypes = [vp]
    L.prt_ctx_set_tuning.argtypes = [vp, C.c_char_p, i32]
    L.prt_ctx_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.prt_scene_create.argtypes = [vp, vp, sz, u32, vp, u32, C.POINTER(vp)]
    L.prt_scene_destroy.argtypes = [vp]
    L.prt_scene_destroy.restype = None
    L.prt_scene_get_info.argtypes = [vp, C.POINTER(SceneInfo)]
    L.prt_trace_any_hit.argtypes = [vp, vp, u32, vp]
    L.prt_trace_closest_hit.argtypes = [vp, vp, u32, vp, vp, vp]
    L.prt_bake_params_default.argtypes = [C.POINTER(BakeParams)]
    L.prt_bake_params_default.restype = None
    L.prt_bake_transfer.argtypes = [vp, vp, vp, vp, sz, u32, u32, C.POINTER(BakeParams), vp, vp]
    L.prt_bake_transfer_device.argtypes = [vp, vp, vp, vp, sz, u32, u32, C.POINTER(BakeParams), vp, vp, vp]
    L.prt_bake_transfer_device_shard.argtypes = [vp, vp, vp, vp, sz, u32, u32, u32, C.POINTER(BakeParams), vp, vp, vp]
    L.prt_bake_transfer_device_shard_fused.argtypes = [vp, vp, vp, vp, sz, u32, u32, u32, C.POINTER(BakeParams), vp, C.POINTER(vp), C.c_int32, vp]
    L.prt_device_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.prt_device_free.argtypes = [vp, vp]
    L.prt_ipc_export.argtypes = [vp, vp, vp]
    L.prt_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.prt_ipc_close.argtypes = [vp, vp]
    L.prt_bake_transfer_device_strided.argtypes = [vp, vp, vp, vp, sz, u32, u32, C.POINTER(BakeParams), vp, sz, vp, vp]
    L.prt_scatter_sh9.argtypes = [vp, C.c_int32, u32, vp, sz, sz]
    L.prt_bake_sample_table.argtypes = [C.POINTER(BakeParams), vp, vp]
    L.prt_ctx_last_bake_stats.argtypes = [vp, C.POINTER(BakeStats)]
    L.prt_env_create.argtypes = [vp, vp, i32, i32, i32, C.POINTER(vp)]
    L.prt_env_destroy.argtypes = [vp]
    L.prt_env_destroy.restype = None
    L.prt_env_levels.argtypes = [vp]
    L.prt_env_get_cube.argtypes = [vp, i32, vp]
    L.prt_env_irradiance.argtypes = [vp, i32, vp]
    L.prt_env_prefilter.argtypes = [vp, i32, i32, i32, vp]
    L.prt_brdf_lut.argtypes = [vp, i32, i32, i32, vp]
    L.prt_env_project_sh.argtypes = [vp, i32, i32, i32, vp]
    L.prt_sh_pack_rh.argtypes = [vp, vp]
    L.prt_probe_capture.argtypes = [vp, vp, u32, vp, vp, u32, C.POINTER(vp)]
    L.prt_csr_destroy.argtypes = [vp]
    L.prt_csr_destroy.restype = None
    L.prt_csr_sizes.argtypes = [vp, C.POINTER(u32), C.POINTER(C.c_uint64), C.POINTER(u32), C.POINTER(C.c_double)]
    L.prt_csr_download.argtypes = [vp, vp, vp, vp, vp, vp]
    L.prt_csr_surfel_sums.argtypes = [vp, vp]
    L.prt_csr_upload.argtypes = [vp, u32, C.c_uint64, u32, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.prt_probe_project.argtypes = [vp, vp, vp]
    L.prt_probe_positions.argtypes = [vp, vp, vp]
    L.prt_fibonacci_dirs.argtypes = [i32, vp]
    L.prt_cube_dirs.argtypes = [i32, vp, vp]
    L.prt_volume_weights.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.prt_paral_shadow_matrix.argtypes = [C.c_float, C.c_float, vp, vp]
    L.prt_shadow_map.argtypes = [vp, vp, i32, vp]
    L.prt_gi_create.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.prt_gi_destroy.argtypes = [vp]
    L.prt_gi_destroy.restype = None
    L.prt_gi_set_shadow_map.argtypes = [vp, vp, i32]
    L.prt_gi_set_albedo.argtypes = [vp, vp]
    L.prt_gi_set_radiance.argtypes = [vp, vp]
    L.prt_gi_step.argtypes = [vp, C.POINTER(RelightParams), i32]
    L.prt_gi_download.argtypes = [vp, vp, vp, vp]
    L.prt_film_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.prt_film_destroy.argtypes = [vp]
    L.prt_film_destroy.restype = None
    L.prt_film_reset.argtypes = [vp]
    L.prt_raytrace.argtypes = [vp, vp, C.POINTER(Camera), i32, vp, i32, i32, u32, i32]
    L.prt_film_download.argtypes = [vp, vp, vp]
    u64 = C.c_uint64
    L.prt_hash_bytes.restype = u64
    L.prt_hash_bytes.argtypes = [vp, sz, u64]
    L.prt_mesh_hash.restype = u64
    L.prt_mesh_hash.argtypes = [vp, vp, sz, u32, vp, u32]
    L.prt_group_create.argtypes = [vp, i32, C.POINTER(vp)]
    L.prt_group_destroy.argtypes = [vp]
    L.prt_group_destroy.restype = None
    L.prt_group_size.argtypes = [vp]
    L.prt_group_ctx.argtypes = [vp, i32]
    L.prt_group_ctx.restype = vp
    L.prt_group_capabilities.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.c_char_p, sz]
    L.prt_group_set_tuning.argtypes = [vp, C.c_char_p, i32]
    L.prt_group_scene_create.argtypes = [vp, vp, sz, u32, vp, u32, C.POINTER(vp)]
    L.prt_group_scene_destroy.argtypes = [vp]
    L.prt_group_scene_destroy.restype = None
    L.prt_group_scene_get_info.argtypes = [vp, C.POINTER(SceneInfo)]
    L.prt_group_scene_member.argtypes = [vp, i32]
    L.prt_group_scene_member.restype = vp
    L.prt_group_bake_transfer.argtypes = [vp, vp, vp, vp, sz, u32, C.POINTER(BakeParams), vp, i32, C.POINTER(GroupStats)]
    L.prt_group_rows_device.argtypes = [vp, i32]
    L.prt_group_rows_device.restype = vp
    L.prt_group_download_rows.argtypes = [vp, i32, vp]
    L.prt_group_probe_capture.argtypes = [vp, vp, vp, u32, vp, vp, u32, i32, C.POINTER(vp), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.prt_cache_save_transfer.argtypes = [C.c_char_p, u64, u32, C.POINTER(BakeParams), vp]
    L.prt_cache_load_transfer.argtypes = [C.c_char_p, u64, u32, C.POINTER(BakeParams), vp]
    L.prt_cache_save_csr.argtypes = [C.c_char_p, u64, u64, u32, u64, u32, vp, vp, vp, vp, vp]
    L.prt_cache_csr_sizes.argtypes = [C.c_char_p, u64, u64, C.POINTER(u32), C.POINTER(u64), C.POINTER(u32)]
    L.prt_cache_load_csr.argtypes = [C.c_char_p, u64, u64, u32, u64, u32, vp, vp, vp, vp, vp]
    _LIB = L
    return L


def _check(rc: int, what: str):
    if rc != 0:
        raise PRTError(f"{what} failed (rc={rc}): {load_library().prt_last_error().decode(errors='replace')}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One GPU = one context (the reference's process-global RTCDevice, raytracing.cpp:42-52)."""

    def __init__(self, device: int = -1):
        self.L = load_library()
        h = C.c_void_p()
        _check(self.L.prt_ctx_create(device, C.byref(h)), "prt_ctx_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.prt_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def device(self) -> int:
        return self.L.prt_ctx_device(self.h)

    def set_tuning(self, **kw):
        for k, v in kw.items():
            _check(self.L.prt_ctx_set_tuning(self.h, k.encode(), int(v)), f"prt_ctx_set_tuning({k})")

    def last_kernel_ms(self) -> float:
        ms = C.c_double()
        _check(self.L.prt_ctx_last_kernel_ms(self.h, C.byref(ms)), "prt_ctx_last_kernel_ms")
        return ms.value

    def last_bake_stats(self) -> BakeStats:
        s = BakeStats()
        _check(self.L.prt_ctx_last_bake_stats(self.h, C.byref(s)), "prt_ctx_last_bake_stats")
        return s


_DEFAULT_CTX = {}


def default_context(device: int = -1) -> Context:
    if device not in _DEFAULT_CTX:
        _DEFAULT_CTX[device] = Context(device)
    return _DEFAULT_CTX[device]


class RTScene:
    """reference RTScene(Mesh&) (raytracing.cpp:58-99): positions + index triples -> GPU-resident BVH."""

    def __init__(self, pos: np.ndarray, tri: np.ndarray, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.L = self.ctx.L
        pos = np.ascontiguousarray(pos, np.float32)
        tri = np.ascontiguousarray(tri, np.uint32)
        if pos.ndim != 2 or pos.shape[1] != 3 or tri.ndim != 2 or tri.shape[1] != 3:
            raise PRTError("RTScene: pos must be [V,3] float32 and tri [F,3] uint32")
        h = C.c_void_p()
        _check(self.L.prt_scene_create(self.ctx.h, _ptr(pos), 12, len(pos), _ptr(tri), len(tri), C.byref(h)), "prt_scene_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.L.prt_scene_destroy(self.h)
        self.h = None

    __del__ = close

    def info(self) -> SceneInfo:
        s = SceneInfo()
        _check(self.L.prt_scene_get_info(self.h, C.byref(s)), "prt_scene_get_info")
        return s

    @staticmethod
    def pack_rays(org, dirs, tnear=0.0, tfar=np.inf) -> np.ndarray:
        dirs = np.asarray(dirs, np.float32)
        rays = np.empty((len(dirs), 8), np.float32)
        rays[:, 0:3] = np.asarray(org, np.float32)
        rays[:, 3] = tnear
        rays[:, 4:7] = dirs
        rays[:, 7] = tfar
        return rays

    def any_hit(self, rays: np.ndarray) -> np.ndarray:
        """struct Ray::any_hit (light_probe.cpp:125-130) for a batch: rays [n,8] = org, tnear, dir, tfar."""
        rays = np.ascontiguousarray(rays, np.float32)
        out = np.zeros(len(rays), np.uint8)
        _check(self.L.prt_trace_any_hit(self.h, _ptr(rays), len(rays), _ptr(out)), "prt_trace_any_hit")
        return out

    def first_hit(self, rays: np.ndarray):
        """struct Ray::first_hit + hit_normal (light_probe.cpp:115-124): returns (t, prim, Ng)."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = len(rays)
        t, prim, ng = np.zeros(n, np.float32), np.zeros(n, np.uint32), np.zeros((n, 3), np.float32)
        _check(self.L.prt_trace_closest_hit(self.h, _ptr(rays), n, _ptr(t), _ptr(prim), _ptr(ng)), "prt_trace_closest_hit")
        return t, prim, ng


def bake_transfer(scene: RTScene | None, pos, nrm, params: BakeParams, want_vis=False, vertex_id_base=0,
                  ctx: Context | None = None):
    """prt_bake_transfer with host buffers (the e2e path). Returns (coeffs [n, order^2], vis words or None)."""
    ctx = ctx or (scene.ctx if scene is not None else default_context())
    pos = np.ascontiguousarray(pos, np.float32)
    nrm = np.ascontiguousarray(nrm, np.float32)
    n = len(pos)
    out = np.zeros((n, params.n_coeffs), np.float32)
    vis = np.zeros((n, (params.n_samples + 31) // 32), np.uint32) if want_vis else None
    _check(ctx.L.prt_bake_transfer(ctx.h, scene.h if scene is not None else None, _ptr(pos), _ptr(nrm), 12, n,
                                   vertex_id_base, C.byref(params), _ptr(out), _ptr(vis)), "prt_bake_transfer")
    return out, vis


def bake_SH(verts: np.ndarray, indices: np.ndarray, params: BakeParams | None = None, ctx: Context | None = None):
    """reference ``void bake_SH(Mesh& gl_mesh)`` (raytracing.cpp:320-360).

    ``verts`` is the reference's interleaved Mesh::Vert array, shape [V, 15] float32 = pos(3) norm(3) sh_coeff(9)
    (gl.h:76-80); it is updated in place (sh_coeff columns) exactly like ``edit_verts()`` and also returned.
    Default parameters are ``prt_bake_params_default`` (the reference's, with the Condon-Shortley sign of ``sh::EvalSH``).
    """
    if params is None:
        params = BakeParams()
        load_library().prt_bake_params_default(C.byref(params))    # incl. cs_phase = 1: bake_SH evaluates sh::EvalSH
    verts = np.asarray(verts)
    if verts.dtype != np.float32 or verts.ndim != 2 or verts.shape[1] != 15 or not verts.flags.c_contiguous:
        raise PRTError("bake_SH: verts must be a C-contiguous [V,15] float32 Mesh::Vert array")
    ctx = ctx or default_context()
    scene = RTScene(verts[:, 0:3], np.asarray(indices, np.uint32).reshape(-1, 3), ctx)
    n = len(verts)
    out = np.zeros((n, params.n_coeffs), np.float32)
    L = ctx.L
    base = verts.ctypes.data
    _check(L.prt_bake_transfer(ctx.h, scene.h, C.c_void_p(base), C.c_void_p(base + 12), 60, n, 0, C.byref(params),
                               _ptr(out), None), "prt_bake_transfer")
    if params.order >= 3:
        _check(L.prt_scatter_sh9(_ptr(out), params.order, n, C.c_void_p(base), 60, 24), "prt_scatter_sh9")
    scene.close()
    return out


class DeviceBuffer:
    """A plain device allocation of the library (``prt_device_alloc``: cudaMalloc, zeroed) that other processes can map through CUDA
    IPC (``export`` -> 64-byte handle, ``DeviceBuffer.open(ctx, handle)`` in the peer).  ``__cuda_array_interface__`` lets
    ``torch.as_tensor(buf, device=...)`` view it as float32 [n_floats] without a copy."""

    def __init__(self, ctx: Context, n_floats: int, _ptr_=None):
        self.ctx, self.L, self.n_floats, self.opened = ctx, ctx.L, int(n_floats), _ptr_ is not None
        if _ptr_ is None:
            p = C.c_void_p()
            _check(self.L.prt_device_alloc(ctx.h, self.n_floats * 4, C.byref(p)), "prt_device_alloc")
            _ptr_ = p.value
        self.ptr = _ptr_

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.n_floats,), "typestr": "<f4", "data": (self.ptr, False), "version": 2}

    def export(self) -> bytes:
        h = (C.c_uint8 * 64)()
        _check(self.L.prt_ipc_export(self.ctx.h, C.c_void_p(self.ptr), h), "prt_ipc_export")
        return bytes(h)

    @classmethod
    def open(cls, ctx: Context, handle: bytes, n_floats: int) -> "DeviceBuffer":
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _check(ctx.L.prt_ipc_open(ctx.h, h, C.byref(p)), "prt_ipc_open")
        return cls(ctx, n_floats, _ptr_=p.value)

    def close(self):
        if getattr(self, "ptr", None) and getattr(self.ctx, "h", None):
            (self.L.prt_ipc_close if self.opened else self.L.prt_device_free)(self.ctx.h, C.c_void_p(self.ptr))
        self.ptr = None

    __del__ = close


class Group:
    """Multi-GPU driver in ONE process (``prt_group_*``): the reference's ``bake_SH`` vertex loop (raytracing.cpp:328) sharded over the
    listed GPUs, BVH replicated, rows gathered on every GPU (fused P2P stores or NCCL all-gather)."""

    def __init__(self, devices):
        self.L = load_library()
        ids = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        _check(self.L.prt_group_create(ids, len(devices), C.byref(h)), "prt_group_create")
        self.h, self.n = h, len(devices)
        self.scene_h = None

    def close(self):
        if getattr(self, "scene_h", None):
            self.L.prt_group_scene_destroy(self.scene_h)
            self.scene_h = None
        if getattr(self, "h", None):
            self.L.prt_group_destroy(self.h)
            self.h = None

    __del__ = close

    def capabilities(self) -> dict:
        p2p, nccl, ver = C.c_int(), C.c_int(), C.c_int()
        why = C.create_string_buffer(512)
        _check(self.L.prt_group_capabilities(self.h, C.byref(p2p), C.byref(nccl), C.byref(ver), why, 512), "prt_group_capabilities")
        return {"p2p": bool(p2p.value), "nccl": bool(nccl.value), "nccl_version": ver.value, "why": why.value.decode()}

    def set_tuning(self, **kw):
        for k, v in kw.items():
            _check(self.L.prt_group_set_tuning(self.h, k.encode(), int(v)), f"prt_group_set_tuning({k})")

    def set_scene(self, pos: np.ndarray, tri: np.ndarray) -> SceneInfo:
        pos = np.ascontiguousarray(pos, np.float32); tri = np.ascontiguousarray(tri, np.uint32)
        if self.scene_h:
            self.L.prt_group_scene_destroy(self.scene_h)
            self.scene_h = None
        h = C.c_void_p()
        _check(self.L.prt_group_scene_create(self.h, _ptr(pos), 12, len(pos), _ptr(tri), len(tri), C.byref(h)), "prt_group_scene_create")
        self.scene_h = h
        info = SceneInfo()
        _check(self.L.prt_group_scene_get_info(h, C.byref(info)), "prt_group_scene_get_info")
        return info

    def bake_transfer(self, pos, nrm, params: BakeParams, gather: int = GATHER_AUTO, out: np.ndarray | None = None, want_host: bool = True):
        """-> (rows [n, order^2] in the order of ``pos`` (or None), GroupStats)."""
        pos = np.ascontiguousarray(pos, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
        n = len(pos)
        if want_host and out is None:
            out = np.zeros((n, params.n_coeffs), np.float32)
        st = GroupStats()
        _check(self.L.prt_group_bake_transfer(self.h, self.scene_h, _ptr(pos), _ptr(nrm), 12, n, C.byref(params),
                                              _ptr(out) if want_host else None, gather, C.byref(st)), "prt_group_bake_transfer")
        return (out if want_host else None), st

    def member(self, i: int):
        """(context, scene) of member ``i`` as non-owning views: the group's own BVH copy on that GPU, usable with ``bake_transfer`` etc."""
        class _View:
            pass
        ctx, sc = _View(), _View()
        ctx.L, ctx.h = self.L, C.c_void_p(self.L.prt_group_ctx(self.h, i))
        sc.L, sc.ctx, sc.h = self.L, ctx, C.c_void_p(self.L.prt_group_scene_member(self.scene_h, i))
        sc._group = self                                   # keeps the group alive
        return ctx, sc

    def probe_capture(self, probe_pos, dirs, weights, target: int = 0):
        """SH_volume::precompute over the group -> (ProbeTransfer on member ``target``, capture kernel ms (slowest GPU), merge ms)."""
        pp = np.ascontiguousarray(probe_pos, np.float32); d = np.ascontiguousarray(dirs, np.float32); w = np.ascontiguousarray(weights, np.float32)
        h, cap, mrg = C.c_void_p(), C.c_double(), C.c_double()
        _check(self.L.prt_group_probe_capture(self.h, self.scene_h, _ptr(pp), len(pp), _ptr(d), _ptr(w), len(d), target, C.byref(h), C.byref(cap),
                                              C.byref(mrg)), "prt_group_probe_capture")
        pt = ProbeTransfer.__new__(ProbeTransfer)
        pt.scene, pt.L, pt.h = self, self.L, h              # keeps the group alive; `.h` doubles as the liveness check
        pt._sizes()
        return pt, cap.value, mrg.value

    def download_rows(self, member: int, n: int, n2: int) -> np.ndarray:
        out = np.zeros((n, n2), np.float32)
        _check(self.L.prt_group_download_rows(self.h, member, _ptr(out)), "prt_group_download_rows")
        return out


def sample_table(params: BakeParams):
    S = params.n_samples
    uv, dirs = np.zeros((S, 2), np.float32), np.zeros((S, 3), np.float32)
    _check(load_library().prt_bake_sample_table(C.byref(params), _ptr(uv), _ptr(dirs)), "prt_bake_sample_table")
    return uv, dirs


class LightProbe:
    """reference ``LightProbe`` passes (src/opengl/gl.h:273-298, gl.cpp:543-591) + ``load_hdr`` upload (util.cpp:6-24) on the GPU.

    ``equirect`` is the [h, w, 3] float32 image stb_image returns for data/hdr/newport_loft.hdr (top row first)."""

    def __init__(self, equirect: np.ndarray, cube_size: int = 512, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.L = self.ctx.L
        eq = np.ascontiguousarray(equirect, np.float32)
        if eq.ndim != 3 or eq.shape[2] != 3:
            raise PRTError("LightProbe: equirect must be [h, w, 3] float32")
        h = C.c_void_p()
        _check(self.L.prt_env_create(self.ctx.h, _ptr(eq), eq.shape[1], eq.shape[0], cube_size, C.byref(h)), "prt_env_create")
        self.h, self.cube_size = h, cube_size
        self.levels = self.L.prt_env_levels(h)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.L.prt_env_destroy(self.h)
        self.h = None

    __del__ = close

    def cube(self, level: int = 0) -> np.ndarray:
        """equirectangular_to_cubemap (+ generateMipmap) result, [6, n, n, 3]."""
        n = self.cube_size >> level
        out = np.zeros((6, n, n, 3), np.float32)
        _check(self.L.prt_env_get_cube(self.h, level, _ptr(out)), "prt_env_get_cube")
        return out

    def irradiance(self, n_out: int = 32) -> np.ndarray:
        out = np.zeros((6, n_out, n_out, 3), np.float32)
        _check(self.L.prt_env_irradiance(self.h, n_out, _ptr(out)), "prt_env_irradiance")
        return out

    def prefilter(self, n_out: int = 256, mips: int = 5, n_samples: int = 1024) -> list:
        sizes = [n_out >> m for m in range(mips)]
        flat = np.zeros(sum(6 * n * n * 3 for n in sizes), np.float32)
        _check(self.L.prt_env_prefilter(self.h, n_out, mips, n_samples, _ptr(flat)), "prt_env_prefilter")
        out, o = [], 0
        for n in sizes:
            out.append(flat[o:o + 6 * n * n * 3].reshape(6, n, n, 3))
            o += 6 * n * n * 3
        return out

    def project_sh(self, order: int = 3, method: int = 0, size: int | None = None) -> np.ndarray:
        size = size or (256 if method == 0 else 64)
        out = np.zeros((order * order, 3), np.float32)
        _check(self.L.prt_env_project_sh(self.h, order, method, size, _ptr(out)), "prt_env_project_sh")
        return out


def brdf_lut(w: int = 512, h: int = 512, n_samples: int = 1024, ctx: Context | None = None) -> np.ndarray:
    """reference ``brdfLUT.render_to(screen_quad)`` with brdf.frag (app.cpp:61-63): [h, w, 2] = (A, B)."""
    ctx = ctx or default_context()
    out = np.zeros((h, w, 2), np.float32)
    _check(ctx.L.prt_brdf_lut(ctx.h, w, h, n_samples, _ptr(out)), "prt_brdf_lut")
    return out


def sh_pack_rh(L9: np.ndarray) -> np.ndarray:
    L9 = np.ascontiguousarray(L9, np.float32)
    out = np.zeros(28, np.float32)
    _check(load_library().prt_sh_pack_rh(_ptr(L9), _ptr(out)), "prt_sh_pack_rh")
    return out


def probe_positions(res, scene_size) -> np.ndarray:
    """SH_volume::init probe grid (volume.cpp:83-90)."""
    res = np.asarray(res, np.int32)
    size = np.asarray(scene_size, np.float32)
    out = np.zeros((int(np.prod(res)), 3), np.float32)
    _check(load_library().prt_probe_positions(_ptr(res), _ptr(size), _ptr(out)), "prt_probe_positions")
    return out


def fibonacci_dirs(n: int):
    """get_dirs (light_probe.cpp:137-152) with uniform solid angle 4 pi / n."""
    d = np.zeros((n, 3), np.float32)
    _check(load_library().prt_fibonacci_dirs(n, _ptr(d)), "prt_fibonacci_dirs")
    return d, np.full(n, 4 * np.pi / n, np.float32)


def cube_dirs(res: int):
    d, w = np.zeros((6 * res * res, 3), np.float32), np.zeros(6 * res * res, np.float32)
    _check(load_library().prt_cube_dirs(res, _ptr(d), _ptr(w)), "prt_cube_dirs")
    return d, w


class ProbeTransfer:
    """SH_volume::precompute result (volume.cpp:149-316) on the GPU: CSR probe -> surfel transfer + surfel table."""

    def __init__(self, scene: RTScene, probe_pos: np.ndarray, dirs: np.ndarray, weights: np.ndarray):
        self.scene, self.L = scene, scene.L
        pp = np.ascontiguousarray(probe_pos, np.float32)
        d = np.ascontiguousarray(dirs, np.float32)
        w = np.ascontiguousarray(weights, np.float32)
        h = C.c_void_p()
        _check(self.L.prt_probe_capture(scene.h, _ptr(pp), len(pp), _ptr(d), _ptr(w), len(d), C.byref(h)), "prt_probe_capture")
        self.h = h
        self._sizes()

    def _sizes(self):
        npb, nnz, ns, ms = C.c_uint32(), C.c_uint64(), C.c_uint32(), C.c_double()
        _check(self.L.prt_csr_sizes(self.h, C.byref(npb), C.byref(nnz), C.byref(ns), C.byref(ms)), "prt_csr_sizes")
        self.n_probes, self.nnz, self.n_surfels, self.capture_ms = npb.value, nnz.value, ns.value, ms.value

    @classmethod
    def from_arrays(cls, ctx: Context, rng, ids, transfer, surfels, keys=None) -> "ProbeTransfer":
        """Device CSR from host arrays (a merged multi-GPU capture, ``prt_b200.dist.merge_probe_csr``, or a cached one)."""
        self = cls.__new__(cls)
        self.scene, self.L = ctx, ctx.L                      # keeps the context alive; `.h` doubles as the liveness check
        rng = np.ascontiguousarray(rng, np.uint32); ids = np.ascontiguousarray(ids, np.uint32)
        tr = np.ascontiguousarray(transfer, np.float32); sf = np.ascontiguousarray(surfels, np.float32)
        k = None if keys is None else np.ascontiguousarray(keys, np.uint64)
        if rng.ndim != 2 or rng.shape[1] != 2 or tr.shape != (len(ids), 9) or sf.ndim != 2 or sf.shape[1] != 6 or (k is not None and len(k) != len(sf)):
            raise PRTError("from_arrays: expected range [P,2], ids [nnz], transfer [nnz,9], surfels [S,6], keys [S]")
        h = C.c_void_p()
        _check(self.L.prt_csr_upload(ctx.h, len(rng), len(ids), len(sf), _ptr(rng), _ptr(ids), _ptr(tr), _ptr(sf),
                                     None if k is None else _ptr(k), C.byref(h)), "prt_csr_upload")
        self.h = h
        self._sizes()
        return self

    def close(self):
        if getattr(self, "h", None) and getattr(self.scene, "h", None):
            self.L.prt_csr_destroy(self.h)
        self.h = None

    __del__ = close

    def download(self):
        rng = np.zeros((self.n_probes, 2), np.uint32)
        ids = np.zeros(self.nnz, np.uint32)
        tr = np.zeros((self.nnz, 9), np.float32)
        sf = np.zeros((self.n_surfels, 6), np.float32)
        keys = np.zeros(self.n_surfels, np.uint64)
        _check(self.L.prt_csr_download(self.h, _ptr(rng), _ptr(ids), _ptr(tr), _ptr(sf), _ptr(keys)), "prt_csr_download")
        return rng, ids, tr, sf, keys

    def surfel_sums(self) -> np.ndarray:
        """[n_surfels, 7] float64: sum of hit positions, sum of hit normals, hit count (for merging partial captures)."""
        sums = np.zeros((self.n_surfels, 7), np.float64)
        _check(self.L.prt_csr_surfel_sums(self.h, _ptr(sums)), "prt_csr_surfel_sums")
        return sums

    def project(self, radiance_rgba: np.ndarray) -> np.ndarray:
        """SH_volume::project_sh (precomp_projectSH.comp): [n_surfels,4] radiance -> [n_probes,7,4] packed SH volumes."""
        rad = np.ascontiguousarray(radiance_rgba, np.float32)
        if rad.shape != (self.n_surfels, 4):
            raise PRTError("project: radiance must be [n_surfels, 4]")
        out = np.zeros((self.n_probes, 7, 4), np.float32)
        _check(self.L.prt_probe_project(self.h, _ptr(rad), _ptr(out)), "prt_probe_project")
        return out


def calculate_weight(scene: RTScene, probe_res, volume_res, scene_size):
    """reference ``Volume_weight calculate_weight(Model&, probe_res, volume_res, scene_size)`` (light_probe.cpp:156-367).
    Returns (weight0123, weight4567, inside_score), each indexed (z*ry + y)*rx + x."""
    pr = np.asarray(probe_res, np.int32); vr = np.asarray(volume_res, np.int32); sz = np.asarray(scene_size, np.float32)
    n = int(np.prod(vr))
    w0, w1, sc = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros(n, np.float32)
    _check(scene.L.prt_volume_weights(scene.h, _ptr(pr), _ptr(vr), _ptr(sz), _ptr(w0), _ptr(w1), _ptr(sc)), "prt_volume_weights")
    return w0, w1, sc


def paral_shadow_matrix(up: float, direction: float):
    """reference ``Paral_Shadow::set_dir`` (src/opengl/gl.cpp:620-631) -> (sky direction [3], lightSpaceMatrix [4,4] column-major rows)."""
    d = np.zeros(3, np.float32); m = np.zeros(16, np.float32)
    _check(load_library().prt_paral_shadow_matrix(float(up), float(direction), _ptr(d), _ptr(m)), "prt_paral_shadow_matrix")
    return d, m


def shadow_map(scene: RTScene, matrix, size: int = 4096) -> np.ndarray:
    """reference ``Paral_Shadow::render`` (src/opengl/gl.cpp:633-648) by ray casting: [size, size] depths in [0,1]."""
    m = np.ascontiguousarray(np.asarray(matrix, np.float32).reshape(16))
    out = np.zeros((size, size), np.float32)
    _check(scene.L.prt_shadow_map(scene.h, _ptr(m), size, _ptr(out)), "prt_shadow_map")
    return out


class SHVolume:
    """Per-frame half of the reference's ``SH_volume`` (src/sh/volume.cpp:357-452): ``relight()`` + ``project_sh()`` on device-resident
    state.  ``transfer``: a ``ProbeTransfer`` of prod(probe_res) probes; ``weights``: ``calculate_weight`` output."""

    def __init__(self, transfer: ProbeTransfer, probe_res, volume_res, scene_size, weights):
        self.transfer, self.L = transfer, transfer.L
        self.probe_res = np.asarray(probe_res, np.int32); self.volume_res = np.asarray(volume_res, np.int32)
        sz = np.asarray(scene_size, np.float32)
        w0 = np.ascontiguousarray(weights[0], np.float32); w1 = np.ascontiguousarray(weights[1], np.float32)
        n = int(np.prod(self.volume_res))
        if w0.shape != (n, 4) or w1.shape != (n, 4):
            raise PRTError("SHVolume: weights must be two [prod(volume_res), 4] arrays")
        h = C.c_void_p()
        _check(self.L.prt_gi_create(transfer.h, _ptr(self.probe_res), _ptr(self.volume_res), _ptr(sz), _ptr(w0), _ptr(w1), C.byref(h)), "prt_gi_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.transfer, "h", None):
            self.L.prt_gi_destroy(self.h)
        self.h = None

    __del__ = close

    def set_shadow_map(self, depth):
        if depth is None:
            _check(self.L.prt_gi_set_shadow_map(self.h, None, 0), "prt_gi_set_shadow_map")
            return
        d = np.ascontiguousarray(depth, np.float32)
        if d.ndim != 2 or d.shape[0] != d.shape[1]:
            raise PRTError("set_shadow_map: depth must be square")
        _check(self.L.prt_gi_set_shadow_map(self.h, _ptr(d), d.shape[0]), "prt_gi_set_shadow_map")

    def set_albedo(self, albedo_rgb):
        a = None if albedo_rgb is None else np.ascontiguousarray(albedo_rgb, np.float32)
        if a is not None and a.shape != (self.transfer.n_surfels, 3):
            raise PRTError("set_albedo: albedo must be [n_surfels, 3]")
        _check(self.L.prt_gi_set_albedo(self.h, None if a is None else _ptr(a)), "prt_gi_set_albedo")

    def set_radiance(self, radiance_rgba):
        r = np.ascontiguousarray(radiance_rgba, np.float32)
        if r.shape != (self.transfer.n_surfels, 4):
            raise PRTError("set_radiance: radiance must be [n_surfels, 4]")
        _check(self.L.prt_gi_set_radiance(self.h, _ptr(r)), "prt_gi_set_radiance")

    def step(self, params: RelightParams, n_rounds: int = 1):
        """n_rounds x (relight(); project_sh()) -- reference app.cpp:164-166 runs one round per frame."""
        _check(self.L.prt_gi_step(self.h, C.byref(params), int(n_rounds)), "prt_gi_step")

    def download(self):
        """-> radiance [n_surfels,4], probe_sh [n_probes,7,4], volumes [n_voxels,7,4]"""
        rad = np.zeros((self.transfer.n_surfels, 4), np.float32)
        psh = np.zeros((self.transfer.n_probes, 7, 4), np.float32)
        vol = np.zeros((int(np.prod(self.volume_res)), 7, 4), np.float32)
        _check(self.L.prt_gi_download(self.h, _ptr(rad), _ptr(psh), _ptr(vol)), "prt_gi_download")
        return rad, psh, vol


AO, NORMAL = 0, 1


class Film:
    """App::pixels_w / App::pixels of the reference's preview tracer (raytracing.cpp:280-317), resident on the GPU."""

    def __init__(self, width: int, height: int, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.L = self.ctx.L
        self.width, self.height = int(width), int(height)
        h = C.c_void_p()
        _check(self.L.prt_film_create(self.ctx.h, self.width, self.height, C.byref(h)), "prt_film_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.L.prt_film_destroy(self.h)
        self.h = None

    __del__ = close

    def reset(self):
        _check(self.L.prt_film_reset(self.h), "prt_film_reset")

    def download(self):
        acc = np.zeros((self.height, self.width, 4), np.float32); px = np.zeros((self.height, self.width, 4), np.uint8)
        _check(self.L.prt_film_download(self.h, _ptr(acc), _ptr(px)), "prt_film_download")
        return acc, px


def raytrace(scene: RTScene, film: Film, camera: Camera, max_path_length: int = 3, albedo=(1.0, 1.0, 1.0), gamma: bool = True,
             mode: int = AO, seed: int = 0x50525400, n_frames: int = 1):
    """reference ``raytrace(const RTScene&)`` (raytracing.cpp:280-317): n_frames more samples per pixel into ``film``."""
    a = np.asarray(albedo, np.float32)
    _check(scene.L.prt_raytrace(scene.h, film.h, C.byref(camera), int(max_path_length), _ptr(a), int(bool(gamma)), int(mode),
                                int(seed) & 0xFFFFFFFF, int(n_frames)), "prt_raytrace")


# ---- on-disk cache (host only) -----------------------------------------------------------------------------
ERR_IO, ERR_CACHE_MISS = -6, -7


def mesh_hash(pos: np.ndarray, tri: np.ndarray, nrm: np.ndarray | None = None) -> int:
    """Key of a baked result: positions, triangles and -- for per-vertex transfer, whose rays start at P + eps N in the frame of N --
    the vertex normals (pass ``nrm``); a probe capture does not depend on them."""
    p = np.ascontiguousarray(pos, np.float32); t = np.ascontiguousarray(tri, np.uint32)
    n = None if nrm is None else np.ascontiguousarray(nrm, np.float32)
    if n is not None and n.shape != p.shape:
        raise PRTError("mesh_hash: nrm must have the shape of pos")
    return int(load_library().prt_mesh_hash(_ptr(p), _ptr(n), 12, len(p), _ptr(t), len(t)))


def hash_arrays(*arrays) -> int:
    """FNV-1a over the bytes of the arrays, chained (the ``config_hash`` of a probe capture: positions, directions, weights)."""
    h = 0
    for a in arrays:
        a = np.ascontiguousarray(a)
        h = int(load_library().prt_hash_bytes(_ptr(a), a.nbytes, h))
    return h


def cache_save_transfer(path: str, mhash: int, params: BakeParams, coeffs: np.ndarray):
    c = np.ascontiguousarray(coeffs, np.float32)
    if c.ndim != 2 or c.shape[1] != params.n_coeffs:
        raise PRTError("cache_save_transfer: coeffs must be [n_verts, order^2]")
    _check(load_library().prt_cache_save_transfer(os.fsencode(path), mhash, len(c), C.byref(params), _ptr(c)), "prt_cache_save_transfer")


def cache_load_transfer(path: str, mhash: int, n_verts: int, params: BakeParams):
    """-> [n_verts, order^2] rows, or None on a cache miss (no file, or baked from another mesh / with other parameters)."""
    out = np.zeros((n_verts, params.n_coeffs), np.float32)
    rc = load_library().prt_cache_load_transfer(os.fsencode(path), mhash, n_verts, C.byref(params), _ptr(out))
    if rc == ERR_CACHE_MISS:
        return None
    _check(rc, "prt_cache_load_transfer")
    return out


def cache_save_csr(path: str, mhash: int, config_hash: int, rng, ids, transfer, surfels, keys):
    rng = np.ascontiguousarray(rng, np.uint32); ids = np.ascontiguousarray(ids, np.uint32); tr = np.ascontiguousarray(transfer, np.float32)
    sf = np.ascontiguousarray(surfels, np.float32); k = np.ascontiguousarray(keys, np.uint64)
    _check(load_library().prt_cache_save_csr(os.fsencode(path), mhash, config_hash, len(rng), len(ids), len(sf), _ptr(rng), _ptr(ids), _ptr(tr),
                                             _ptr(sf), _ptr(k)), "prt_cache_save_csr")


def cache_load_csr(path: str, mhash: int, config_hash: int):
    """-> (range, ids, transfer, surfels, keys) or None on a cache miss; feed to ``ProbeTransfer.from_arrays``."""
    L = load_library()
    npb, nnz, ns = C.c_uint32(), C.c_uint64(), C.c_uint32()
    rc = L.prt_cache_csr_sizes(os.fsencode(path), mhash, config_hash, C.byref(npb), C.byref(nnz), C.byref(ns))
    if rc == ERR_CACHE_MISS:
        return None
    _check(rc, "prt_cache_csr_sizes")
    rng = np.zeros((npb.value, 2), np.uint32); ids = np.zeros(nnz.value, np.uint32); tr = np.zeros((nnz.value, 9), np.float32)
    sf = np.zeros((ns.value, 6), np.float32); keys = np.zeros(ns.value, np.uint64)
    _check(L.prt_cache_load_csr(os.fsencode(path), mhash, config_hash, npb.value, nnz.value, ns.value, _ptr(rng), _ptr(ids), _ptr(tr), _ptr(sf),
                                _ptr(keys)), "prt_cache_load_csr")
    return rng, ids, tr, sf, keys
