"""CPU tests: the C-ABI library loads, exports every symbol include/prt_b200.h declares, fails loudly without a GPU, and
its host-side sample table equals the oracle's bit for bit.  No GPU compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "prt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(prt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(prt):
    L = prt.load_library()
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/prt_b200.h but not exported"
    from prt_b200 import api
    assert set(api.ABI_SYMBOLS) <= set(names)
    assert L.prt_abi_version() == 2


def test_library_is_sm100a_only():
    """the in-tree .so carries sm_100a SASS and nothing else (no multi-arch fallback)."""
    import subprocess
    import prt_b200
    out = subprocess.run(["cuobjdump", "-lelf", prt_b200.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(prt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(prt.PRTError, match="no CUDA device|CUDA"):
        prt.Context()


def test_sample_table_bit_identical_to_oracle(prt, oracle):
    from prt_b200 import api
    for kw in [dict(), dict(samples_u=64, samples_v=64), dict(samples_u=128, samples_v=64, jitter=0), dict(samples_u=5, samples_v=7, seed=3)]:
        uv, d = api.sample_table(prt.BakeParams.make(**kw))
        uv2, d2 = oracle.sample_table(oracle.make_params(**kw))
        assert np.array_equal(uv.view(np.uint32), uv2.view(np.uint32))
        assert np.array_equal(d.view(np.uint32), d2.view(np.uint32))


def test_param_struct_layout_matches(prt, oracle):
    assert C.sizeof(prt.BakeParams) == C.sizeof(oracle.BakeParams) == 52
    p = prt.BakeParams()
    prt.load_library().prt_bake_params_default(C.byref(p))
    assert (p.order, p.samples_u, p.samples_v, p.bounces, p.mode, p.jitter) == (3, 32, 32, 0, 1, 1)   # reference defaults
    assert abs(p.origin_eps - 1e-4) < 1e-10 and abs(p.bounce_eps - 1e-5) < 1e-11 and tuple(p.albedo) == (1.0, 1.0, 1.0)
    assert p.cs_phase == 1                  # bake_SH evaluates the basis with sh::EvalSH (raytracing.cpp:226): Condon-Shortley sign
    from prt_b200 import api
    assert C.sizeof(api.GroupStats) == 8 + 8 * 3 + 8 * 8 * 3 + 4 * 8 + 8 * 3


def test_group_api_fails_loudly_without_gpu(prt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(prt.PRTError, match="no CUDA device|CUDA"):
        prt.Group([0, 1])
    L = prt.load_library()
    h = C.c_void_p()
    assert L.prt_group_create(None, 2, C.byref(h)) == -1 and L.prt_group_create((C.c_int * 9)(*range(9)), 9, C.byref(h)) == -1


def test_scatter_sh9_mesh_vert_layout(prt):
    """prt_scatter_sh9 writes rows into the reference's 60-byte Mesh::Vert (gl.h:76-80) without touching pos/norm."""
    L = prt.load_library()
    n = 5
    verts = np.arange(n * 15, dtype=np.float32).reshape(n, 15)
    before = verts.copy()
    co = np.random.RandomState(0).rand(n, 16).astype(np.float32)
    rc = L.prt_scatter_sh9(co.ctypes.data_as(C.c_void_p), 4, n, verts.ctypes.data_as(C.c_void_p), 60, 24)
    assert rc == 0
    assert np.array_equal(verts[:, :6], before[:, :6]) and np.array_equal(verts[:, 6:], co[:, :9])
    assert L.prt_scatter_sh9(co.ctypes.data_as(C.c_void_p), 2, n, verts.ctypes.data_as(C.c_void_p), 60, 24) != 0
