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
