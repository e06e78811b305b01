"""GPU tests of the multi-GPU driver behind the C ABI (prt_group_*, prt_b200/csrc/group.cu) and of the row-placement options of the
bake kernels.  On a one-GPU box the group lists device 0 several times: two contexts on one GPU exercise the sharding, the global
RNG keys, the strided copies to and from the host and the fused P2P row stores (NCCL needs distinct devices: its path is covered
when >= 2 GPUs are visible, e.g. `gpurun --gpus 2`)."""
import ctypes as C

import numpy as np
import pytest

from prt_b200 import meshes

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def mesh():
    pos, nrm, tri = meshes.bumpy_torus(96, 64)
    order = meshes.morton_order(pos)
    return pos, tri, pos[order].copy(), nrm[order].copy()


@pytest.mark.parametrize("world,n_verts", [(1, 1000), (2, 6144), (3, 1000), (4, 6100), (2, 64), (4, 65), (8, 6144)])
def test_group_bake_matches_single_gpu_bitwise(prt, mesh, world, n_verts):
    """Sharded bake == single-context bake of the same list, bit for bit: ragged counts (partial last chunk, ranks without work),
    every gather mode the box supports, rows gathered on every member."""
    pos, tri, vp, vn = mesh
    vp, vn = vp[:n_verts], vn[:n_verts]
    devs = [i % max(n_gpus(), 1) for i in range(world)]
    grp = prt.Group(devs)
    grp.set_scene(pos, tri)
    caps = grp.capabilities()
    params = prt.BakeParams.make(order=3, samples_u=16, samples_v=16)
    ref, _ = prt.bake_transfer(prt.RTScene(pos, tri), vp, vn, params)
    modes = [prt.GATHER_NONE] + ([prt.GATHER_P2P] if caps["p2p"] else []) + ([prt.GATHER_NCCL] if caps["nccl"] and world > 1 else [])
    assert world == 1 or len(modes) >= 2, caps
    for mode in modes:
        got, st = grp.bake_transfer(vp, vn, params, gather=mode)
        assert st.gather_mode == mode and st.n_devices == world and sum(st.vertices[:world]) == n_verts
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"mode {mode}"
        if mode != prt.GATHER_NONE:
            for m in range(world):
                assert np.array_equal(grp.download_rows(m, n_verts, 9).view(np.uint32), ref.view(np.uint32)), f"mode {mode} member {m}"
    got, st = grp.bake_transfer(vp, vn, params)                  # AUTO
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    grp.close()


def test_group_interreflection_keys_rng_by_global_vertex(prt, oracle, mesh):
    """The bounce RNG is keyed by the position in the caller's list, whatever GPU bakes the vertex (sharded interreflection
    must not depend on the world size): 3 members == 1 context == oracle."""
    pos, tri, vp, vn = mesh
    vp, vn = vp[:700], vn[:700]
    kw = dict(order=4, samples_u=16, samples_v=16, bounces=2, albedo=(0.5, 0.5, 0.5))
    grp = prt.Group([0] * 3 if n_gpus() < 3 else [0, 1, 2])
    grp.set_scene(pos, tri)
    got, _ = grp.bake_transfer(vp, vn, prt.BakeParams.make(mode=prt.INTERREFLECT, **kw))
    one, _ = prt.bake_transfer(prt.RTScene(pos, tri), vp, vn, prt.BakeParams.make(mode=prt.INTERREFLECT, **kw))
    ref, _, _ = oracle.bake_transfer(oracle.Scene(pos, tri), vp, vn, oracle.make_params(mode=oracle.INTERREFLECT, **kw))
    assert np.array_equal(got.view(np.uint32), one.view(np.uint32))
    rel = np.linalg.norm(got - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-20)
    assert rel.max() <= 1e-4
    grp.close()


def test_group_mesh_vert_layout_and_errors(prt, mesh):
    """Interleaved Mesh::Vert input (60-byte stride, one upload) through the group entry point; bad arguments fail loudly."""
    pos, tri, vp, vn = mesh
    n = 3000
    verts = np.zeros((n, 15), np.float32)
    verts[:, 0:3], verts[:, 3:6] = vp[:n], vn[:n]
    grp = prt.Group([0, 0])
    grp.set_scene(pos, tri)
    params = prt.BakeParams.make(samples_u=8, samples_v=8)
    out = np.zeros((n, 9), np.float32)
    st = prt.GroupStats()
    base = verts.ctypes.data
    rc = grp.L.prt_group_bake_transfer(grp.h, grp.scene_h, C.c_void_p(base), C.c_void_p(base + 12), 60, n, C.byref(params),
                                       out.ctypes.data_as(C.c_void_p), prt.GATHER_AUTO, C.byref(st))
    assert rc == 0, grp.L.prt_last_error()
    ref, _ = prt.bake_transfer(prt.RTScene(pos, tri), vp[:n], vn[:n], params)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert st.h2d_bytes == n * 60
    with pytest.raises(prt.PRTError):
        grp.bake_transfer(vp[:10], vn[:10], prt.BakeParams.make(order=7))
    if not grp.capabilities()["nccl"]:
        with pytest.raises(prt.PRTError, match="NCCL"):
            grp.bake_transfer(vp[:10], vn[:10], params, gather=prt.GATHER_NCCL)
    grp.close()


def test_strided_device_output_fills_mesh_vert_in_place(prt, mesh):
    """prt_bake_transfer_device_strided: rows written 60 bytes apart at sh_coeff of a device-resident Mesh::Vert array (gl.h:76-80) ==
    packed rows; position / normal columns untouched (the interop half without GL)."""
    import torch
    pos, tri, vp, vn = mesh
    n = 2000
    ctx = prt.Context(0)
    sc = prt.RTScene(pos, tri, ctx)
    verts = np.full((n, 15), 7.0, np.float32)
    verts[:, 0:3], verts[:, 3:6] = vp[:n], vn[:n]
    d = torch.from_numpy(verts).cuda()
    params = prt.BakeParams.make()
    st = torch.cuda.current_stream()
    rc = ctx.L.prt_bake_transfer_device_strided(ctx.h, sc.h, C.c_void_p(d.data_ptr()), C.c_void_p(d.data_ptr() + 12), 60, n, 0, C.byref(params),
                                                C.c_void_p(d.data_ptr() + 24), 60, None, C.c_void_p(st.cuda_stream))
    assert rc == 0, ctx.L.prt_last_error()
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    ref, _ = prt.bake_transfer(sc, vp[:n], vn[:n], params)
    assert np.array_equal(got[:, 6:15].view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(got[:, 0:6], verts[:, 0:6])
    assert ctx.L.prt_bake_transfer_device_strided(ctx.h, sc.h, C.c_void_p(d.data_ptr()), C.c_void_p(d.data_ptr() + 12), 60, n, 0, C.byref(params),
                                                  C.c_void_p(d.data_ptr() + 24), 32, None, C.c_void_p(st.cuda_stream)) != 0      # stride < row


@pytest.mark.parametrize("scale", [0.5, 0.25, 1.00008, 3.0])
def test_non_unit_normals_match_oracle(prt, oracle, mesh, scale):
    """the reference passes assimp's normals through un-normalised (model.cpp:27); frame(N) is then sheared and the
    horizon map (built for an orthonormal frame) must not be used.  Visibility bits exact with the horizon pass on."""
    pos, tri, vp, vn = mesh
    sel = np.arange(0, len(vp), 61)[:100]
    p, n = vp[sel], (vn[sel] * np.float32(scale)).astype(np.float32)
    n[::7] = vn[sel][::7]                                            # a mix of unit and non-unit normals in one launch
    kw = dict(samples_u=32, samples_v=32)
    got, gvis = prt.bake_transfer(prt.RTScene(pos, tri), p, n, prt.BakeParams.make(**kw), want_vis=True)
    ref, ovis, _ = oracle.bake_transfer(oracle.Scene(pos, tri), p, n, oracle.make_params(**kw), want_vis=True)
    assert np.array_equal(gvis, ovis)
    rel = np.linalg.norm(got - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-20)
    assert rel.max() <= 1e-4


def test_two_contexts_on_one_device_large_smem_kernels(prt, mesh):
    """kernels that need > 48 KB of dynamic shared memory must get the attribute on every device / context."""
    pos, tri, vp, vn = mesh
    ndev = n_gpus()
    for dev in range(min(ndev, 2)):
        ctx = prt.Context(dev)
        sc = prt.RTScene(pos, tri, ctx)
        a, _ = prt.bake_transfer(sc, vp[:200], vn[:200], prt.BakeParams.make(order=4, samples_u=16, samples_v=16, mode=prt.INTERREFLECT, bounces=1))
        assert np.isfinite(a).all()
        d, w = prt.fibonacci_dirs(512)
        pt = prt.ProbeTransfer(sc, prt.probe_positions([2, 2, 2], [3, 3, 3]), d, w)
        assert pt.n_probes == 8


@pytest.mark.parametrize("world", [2, 3])
def test_group_probe_capture_equals_single_gpu_capture(prt, world):
    """SH_volume::precompute sharded over the group and merged on the device == one capture of all probes: ranges, ids, keys and
    36-byte transfer rows bit-identical, surfel table to 1e-6 (its means are re-derived from the summed accumulators)."""
    from test_gpu_probe import scene_with_occluder
    pos, tri = scene_with_occluder()
    devs = [i % max(n_gpus(), 1) for i in range(world)]
    grp = prt.Group(devs)
    grp.set_scene(pos, tri)
    probes = prt.probe_positions([5, 3, 3], [6, 6, 6])                 # 45 probes: uneven split
    d, w = prt.fibonacci_dirs(2048)
    pt, cap_ms, merge_ms = grp.probe_capture(probes, d, w, target=world - 1)
    whole = prt.ProbeTransfer(prt.RTScene(pos, tri), probes, d, w)
    assert (pt.n_probes, pt.nnz, pt.n_surfels) == (whole.n_probes, whole.nnz, whole.n_surfels) and pt.nnz > 1000
    a, b = pt.download(), whole.download()
    for k in (0, 1, 2, 4):
        assert np.array_equal(a[k], b[k]), k
    assert np.abs(a[3] - b[3]).max() <= 1e-6
    rad = np.random.RandomState(1).rand(pt.n_surfels, 4).astype(np.float32)
    assert np.array_equal(pt.project(rad), whole.project(rad))
    assert cap_ms > 0 and merge_ms >= 0
    pt.close(); grp.close()


IPC_WORKER = r'''
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.environ["PRT_ROOT"])
import torch, torch.distributed as dist
import prt_b200
from prt_b200 import meshes, dist as pdist
dist.init_process_group("gloo")                      # handles travel through gloo: both ranks may sit on ONE GPU (NCCL would refuse that)
world, rank = dist.get_world_size(), dist.get_rank()
dev = rank % torch.cuda.device_count()
torch.cuda.set_device(dev)
pos, nrm, tri = meshes.bumpy_torus(96, 64)
order = meshes.morton_order(pos)
pm, nm = pos[order][:6100], nrm[order][:6100]        # ragged: the last chunk is partial
V, n2 = len(pm), 16
ctx = prt_b200.Context(dev)
scene = prt_b200.RTScene(pos, tri, ctx)
params = prt_b200.BakeParams.make(order=4, samples_u=16, samples_v=16, mode=prt_b200.INTERREFLECT, bounces=1, albedo=(0.5, 0.5, 0.5))
mine, valid, v_pad = pdist.shard_indices(V, world, rank)
d_pos = torch.from_numpy(np.ascontiguousarray(pm[mine])).cuda(); d_nrm = torch.from_numpy(np.ascontiguousarray(nm[mine])).cuda()
full = prt_b200.DeviceBuffer(ctx, v_pad * n2)
handles = [None] * world
dist.all_gather_object(handles, full.export())
peers = [prt_b200.DeviceBuffer.open(ctx, handles[r], v_pad * n2) for r in range(world) if r != rank]
arr = (C.c_void_p * len(peers))(*[p.ptr for p in peers])
st = torch.cuda.current_stream()
rc = ctx.L.prt_bake_transfer_device_shard_fused(ctx.h, scene.h, C.c_void_p(d_pos.data_ptr()), C.c_void_p(d_nrm.data_ptr()), 12, int(valid.sum()), world, rank,
                                                C.byref(params), C.c_void_p(full.ptr), arr, len(peers), C.c_void_p(st.cuda_stream))
assert rc == 0, ctx.L.prt_last_error()
torch.cuda.synchronize()
dist.barrier()                                       # every rank's stores have landed
rows = torch.as_tensor(full, device="cuda").view(v_pad, n2)[:V].cpu().numpy()
np.save(os.environ["PRT_OUT"] + f".{rank}.npy", rows)
dist.barrier()
for p in peers: p.close()
dist.barrier()
full.close()
dist.destroy_process_group()
'''


def test_fused_gather_between_processes_over_cuda_ipc(prt, tmp_path):
    """One process per GPU (torchrun): every rank stores its finished rows straight into the full-size buffer of every peer through
    CUDA-IPC-mapped pointers (prt_ipc_export / prt_ipc_open, prt_bake_transfer_device_shard_fused).  After a barrier EVERY rank
    holds all rows, bit-identical to an unsharded bake (the bounce RNG is keyed by the list position).  Two ranks; on a one-GPU box
    both sit on GPU 0."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ipc_worker.py"
    script.write_text(IPC_WORKER)
    out = str(tmp_path / "rows")
    env = dict(os.environ, PRT_ROOT=root, PRT_OUT=out)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                    "--master-port", "29631", str(script)], check=True, env=env, timeout=600)
    pos, nrm, tri = meshes.bumpy_torus(96, 64)
    order = meshes.morton_order(pos)
    pm, nm = pos[order][:6100], nrm[order][:6100]
    params = prt.BakeParams.make(order=4, samples_u=16, samples_v=16, mode=prt.INTERREFLECT, bounces=1, albedo=(0.5, 0.5, 0.5))
    ref, _ = prt.bake_transfer(prt.RTScene(pos, tri), pm, nm, params)
    for r in range(2):
        rows = np.load(out + f".{r}.npy")
        assert np.array_equal(rows.view(np.uint32), ref.view(np.uint32)), f"rank {r}"
