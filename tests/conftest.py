import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


def load_hostcheck(extra_flags=None):
    """Host build of the product's BVH8 builder + traversal header (test tooling, tests/hostcheck)."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "hostcheck")
    # PRT_HOSTCHECK_FLAGS: extra compile flags (e.g. -DPRT_WAVE_CULL=1 to run the suite on an experimental build variant of the kernels)
    extra = list(extra_flags) if extra_flags else os.environ.get("PRT_HOSTCHECK_FLAGS", "").split()
    import hashlib
    so = os.path.join(d, "libhostcheck.so" if not extra else "libhostcheck_variant_%s.so" % hashlib.md5(" ".join(extra).encode()).hexdigest()[:8])
    csrc = os.path.join(ROOT, "prt_b200", "csrc")
    srcs = [os.path.join(d, "hostcheck.cpp"), os.path.join(d, "hostcheck_warp.cpp"), os.path.join(csrc, "bvh_build.cpp"),
            os.path.join(d, "warp_emu.h")] + [os.path.join(csrc, f) for f in ("traverse.cuh", "prt_math.cuh", "horizon_math.cuh",
                                                                              "entry_list.cuh", "bvh8.h", "bake_wave.cuh", "bake_inter.cuh", "horizon.cuh", "kernels.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")      # float4 & co. for entry_list.cuh
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-fPIC", "-ffp-contract=off", "-march=x86-64-v3", "-shared", "-I", cuda_inc, *extra,
                               "-o", so, srcs[0], srcs[1], srcs[2], "-lpthread"])
    L = C.CDLL(so)
    L.hc_build.restype = C.c_void_p
    L.hc_build.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_uint32]
    L.hc_free.argtypes = [C.c_void_p]
    L.hc_info.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_any_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.hc_closest_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hc_frame.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_hz_pang.restype = C.c_float
    L.hc_hz_pang.argtypes = [C.c_float, C.c_float]
    L.hc_hz_triangle.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.hc_hz_box.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    L.hc_horizon_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hc_bake_wave.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p]
    L.hc_bake_inter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_uint32, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.hc_horizon_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hc_entry_list.restype = C.c_int
    L.hc_entry_list.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    L.hc_use_slabs.argtypes = [C.c_int]
    L.hc_hz_slab.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.hc_slabs.restype = C.c_uint32
    L.hc_slabs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.hc_node_tri_ranges.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_tris.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_dops.restype = C.c_uint32
    L.hc_dops.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.hc_child_tri_ranges.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_wave_dop.argtypes = [C.c_int]
    L.hc_horizon_mid.argtypes = [C.c_int, C.c_float]
    L.hc_horizon_trace_far.restype = C.c_uint32
    L.hc_horizon_trace_far.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
    return L


@pytest.fixture(scope="session")
def hostcheck():
    return load_hostcheck()


@pytest.fixture(scope="session")
def prt():
    """The product package with the CUDA library loaded (GPU tests)."""
    import prt_b200
    prt_b200.load_library()
    return prt_b200
