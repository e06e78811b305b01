"""CPU tests of the multi-GPU host logic with gloo, world_size 2: vertex sharding + all-gather of coefficient rows must
reproduce the single-process bake exactly (the per-rank compute is the oracle here; on the GPU box it is the CUDA bake)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from prt_b200 import dist as pdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,world", [(1, 1), (5000, 2), (4096, 2), (10000, 4), (2049, 8)])
def test_shard_indices_partition(n, world):
    seen = np.zeros(n, int)
    per = None
    for r in range(world):
        idx, valid, total = pdist.shard_indices(n, world, r, chunk=512)
        per = per or len(idx)
        assert len(idx) == per == total // world and (idx < n).all()
        np.add.at(seen, idx[valid], 1)
    assert (seen == 1).all()
    fake = np.stack([np.stack([pdist.shard_indices(n, world, r, 512)[0]] * 2, 1).astype(np.float32) for r in range(world)])
    back = pdist.unshard_rows(fake, n, world, 512)
    assert np.array_equal(back[:, 0], np.arange(n, dtype=np.float32))


WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["PRT_ROOT"])
import torch.distributed as dist
from prt_b200 import meshes, dist as pdist
from oracle import pyoracle as O
dist.init_process_group("gloo")
pos, nrm, tri = meshes.bumpy_torus(40, 26)
order = meshes.morton_order(pos)
pm, nm = pos[order], nrm[order]
sc = O.Scene(pos, tri)
p = O.make_params(order=3, samples_u=8, samples_v=8, mode=O.INTERREFLECT, bounces=1, albedo=(0.5, 0.5, 0.5))
def bake(ps, ns, ids):
    # vertex ids key the bounce RNG, so each row is baked with its global id
    out = np.zeros((len(ps), 9), np.float32)
    for k in range(len(ps)):
        out[k] = O.bake_transfer(sc, ps[k:k+1], ns[k:k+1], p, vertex_id_base=int(ids[k]))[0][0]
    return out
rows = pdist.sharded_bake(bake, pm, nm, 9, chunk=128)
if dist.get_rank() == 0:
    np.save(os.environ["PRT_OUT"], rows)
dist.barrier()
dist.destroy_process_group()
'''


def test_sharded_bake_gloo_world2(tmp_path, oracle):
    from prt_b200 import meshes
    out = tmp_path / "rows.npy"
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PRT_ROOT=ROOT, PRT_OUT=str(out), OMP_NUM_THREADS="1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                    "--master-port", "29611", str(script)], check=True, env=env, timeout=300)
    rows = np.load(out)
    pos, nrm, tri = meshes.bumpy_torus(40, 26)
    order = meshes.morton_order(pos)
    sc = oracle.Scene(pos, tri)
    p = oracle.make_params(order=3, samples_u=8, samples_v=8, mode=oracle.INTERREFLECT, bounces=1, albedo=(0.5, 0.5, 0.5))
    ref, _, _ = oracle.bake_transfer(sc, pos[order], nrm[order], p)      # ids = 0..n-1 in list order
    assert rows.shape == ref.shape and np.array_equal(rows, ref)


def _oracle_capture_fn(oracle, sc, d, w):
    def capture(pp):
        pt = oracle.ProbeTransfer(sc, pp, d, w)
        rng, ids, tr, _, keys = pt.download()
        return dict(range=rng, ids=ids, transfer=tr, keys=keys, sums=pt.surfel_sums())
    return capture


def test_merge_probe_csr_equals_whole_capture(oracle):
    """captures of consecutive probe slices merged on the host == one capture of all probes (ids via the union of keys)."""
    from test_oracle_probe import room
    pos, tri = room()
    sc = oracle.Scene(pos, tri)
    probes = oracle.probe_positions([3, 3, 2], [5, 5, 5])
    d, w = oracle.fibonacci_dirs(600)
    whole = oracle.ProbeTransfer(sc, probes, d, w)
    wr, wi, wt, wsf, wk = whole.download()
    cap = _oracle_capture_fn(oracle, sc, d, w)
    for world in (1, 2, 3, 5):
        parts = [cap(probes[slice(*pdist.probe_shard_range(len(probes), world, r))]) for r in range(world)]
        assert sum(len(p["range"]) for p in parts) == len(probes)
        mr, mi, mt, msf, mk = pdist.merge_probe_csr(parts)
        assert np.array_equal(mr, wr) and np.array_equal(mi, wi) and np.array_equal(mt, wt) and np.array_equal(mk, wk)
        assert np.abs(msf - wsf).max() <= 1e-6
    assert pdist.probe_shard_range(5, 8, 7) == (5, 5) and pdist.probe_shard_range(5, 2, 0) == (0, 3)


PROBE_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["PRT_ROOT"]); sys.path.insert(0, os.path.join(os.environ["PRT_ROOT"], "tests"))
import torch.distributed as dist
from prt_b200 import dist as pdist
from oracle import pyoracle as O
from test_oracle_probe import room
dist.init_process_group("gloo")
pos, tri = room()
sc = O.Scene(pos, tri)
probes = O.probe_positions([3, 3, 1], [5, 5, 5])        # 9 probes: ranks own 5 and 4
d, w = O.fibonacci_dirs(500)
def capture(pp):
    pt = O.ProbeTransfer(sc, pp, d, w)
    rng, ids, tr, _, keys = pt.download()
    return dict(range=rng, ids=ids, transfer=tr, keys=keys, sums=pt.surfel_sums())
rng, ids, tr, sf, keys = pdist.sharded_probe_capture(capture, probes)
if dist.get_rank() == 1:                                 # every rank holds the merged CSR; check the non-zero one
    np.savez(os.environ["PRT_OUT"], rng=rng, ids=ids, tr=tr, sf=sf, keys=keys)
dist.barrier()
dist.destroy_process_group()
'''


def test_sharded_probe_capture_gloo_world2(tmp_path, oracle):
    from test_oracle_probe import room
    out = tmp_path / "csr.npz"
    script = tmp_path / "probe_worker.py"
    script.write_text(PROBE_WORKER)
    env = dict(os.environ, PRT_ROOT=ROOT, PRT_OUT=str(out), OMP_NUM_THREADS="1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                    "--master-port", "29613", str(script)], check=True, env=env, timeout=300)
    z = np.load(out)
    pos, tri = room()
    sc = oracle.Scene(pos, tri)
    probes = oracle.probe_positions([3, 3, 1], [5, 5, 5])
    d, w = oracle.fibonacci_dirs(500)
    wr, wi, wt, wsf, wk = oracle.ProbeTransfer(sc, probes, d, w).download()
    assert np.array_equal(z["rng"], wr) and np.array_equal(z["ids"], wi) and np.array_equal(z["tr"], wt) and np.array_equal(z["keys"], wk)
    assert np.abs(z["sf"] - wsf).max() <= 1e-6
