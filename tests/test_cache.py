"""CPU tests of the on-disk cache: host-only C-ABI calls, no GPU needed."""
import numpy as np
import pytest

import prt_b200
from prt_b200 import meshes


def test_transfer_cache_roundtrip_miss_and_corruption(tmp_path):
    pos, nrm, tri = meshes.bumpy_torus(24, 16)
    mh = prt_b200.mesh_hash(pos, tri)
    assert mh == prt_b200.mesh_hash(pos.copy(), tri.copy()) and mh != 0
    pos2 = pos.copy(); pos2[5, 1] += 1e-3
    assert prt_b200.mesh_hash(pos2, tri) != mh and prt_b200.mesh_hash(pos, tri[::-1].copy()) != mh
    # the transfer rows depend on the vertex normals (ray origin P + eps N, cosine frame): same geometry with other normals is
    # another key -- a re-smoothed or flipped mesh must MISS, never load stale rows
    mhn = prt_b200.mesh_hash(pos, tri, nrm)
    assert mhn != mh and mhn == prt_b200.mesh_hash(pos, tri, nrm.copy())
    nrm2 = nrm.copy(); nrm2[7] = -nrm2[7]
    assert prt_b200.mesh_hash(pos, tri, nrm2) != mhn and prt_b200.mesh_hash(pos, tri, -nrm) != mhn
    p = prt_b200.BakeParams.make(order=4, samples_u=16, samples_v=8)
    rows = np.random.RandomState(0).randn(len(pos), 16).astype(np.float32)
    path = str(tmp_path / "torus.prt")
    assert prt_b200.cache_load_transfer(path, mh, len(pos), p) is None                     # no file yet: miss
    prt_b200.cache_save_transfer(path, mh, p, rows)
    assert np.array_equal(prt_b200.cache_load_transfer(path, mh, len(pos), p), rows)
    # any change of the key is a miss, never stale data
    assert prt_b200.cache_load_transfer(path, mh + 1, len(pos), p) is None
    assert prt_b200.cache_load_transfer(path, mhn, len(pos), p) is None                    # keyed without normals != with normals
    for kw in (dict(order=3), dict(samples_u=32), dict(seed=7), dict(mode=prt_b200.INTERREFLECT, bounces=2), dict(albedo=(0.5, 1, 1)), dict(cs_phase=1)):
        q = prt_b200.BakeParams.make(**{**dict(order=4, samples_u=16, samples_v=8), **kw})
        assert prt_b200.cache_load_transfer(path, mh, len(pos), q) is None, kw
    assert prt_b200.cache_load_transfer(path, mh, len(pos) - 1, p) is None
    with pytest.raises(prt_b200.PRTError):
        prt_b200.cache_save_transfer(path, mh, p, rows[:, :9])
    # damage: flipped payload byte, truncation, foreign file
    raw = bytearray(open(path, "rb").read())
    assert raw[:7] == b"PRTB200" and len(raw) == 80 + rows.nbytes
    bad = bytearray(raw); bad[200] ^= 0x40
    open(path, "wb").write(bad)
    with pytest.raises(prt_b200.PRTError, match="corrupt"):
        prt_b200.cache_load_transfer(path, mh, len(pos), p)
    open(path, "wb").write(raw[:-8])
    with pytest.raises(prt_b200.PRTError, match="corrupt"):
        prt_b200.cache_load_transfer(path, mh, len(pos), p)
    open(path, "wb").write(b"not a cache file, just eighty-plus bytes of text" * 4)
    with pytest.raises(prt_b200.PRTError, match="not a prt_b200 cache"):
        prt_b200.cache_load_transfer(path, mh, len(pos), p)


def test_csr_cache_roundtrip(tmp_path, oracle):
    from test_oracle_probe import room
    pos, tri = room()
    probes = oracle.probe_positions([2, 2, 2], [5, 5, 5])
    d, w = oracle.fibonacci_dirs(300)
    csr = oracle.ProbeTransfer(oracle.Scene(pos, tri), probes, d, w).download()
    mh, ch = prt_b200.mesh_hash(pos, tri), prt_b200.hash_arrays(probes, d, w)
    assert ch != prt_b200.hash_arrays(probes, d, w * 2) and ch == prt_b200.hash_arrays(probes.copy(), d, w)
    path = str(tmp_path / "room.csr")
    assert prt_b200.cache_load_csr(path, mh, ch) is None
    prt_b200.cache_save_csr(path, mh, ch, *csr)
    back = prt_b200.cache_load_csr(path, mh, ch)
    assert all(np.array_equal(a, b) for a, b in zip(back, csr))
    assert prt_b200.cache_load_csr(path, mh, ch + 1) is None
    assert prt_b200.cache_load_transfer(path, mh, 8, prt_b200.BakeParams.make()) is None     # a CSR file is not a transfer file
