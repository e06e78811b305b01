#!/usr/bin/env python
"""bench.py -- shadow rays/s and transfer vertices/s of the per-vertex diffuse SH transfer bake (BASELINE.json).

Workload (config.workload): shadowed PRT transfer, SH order 3 (9 coefficients), 32x32 = 1024 jittered stratified
samples per vertex, on a synthetic buddha-scale mesh (bumpy_torus 737x737: 543 169 vertices / 1 086 338 triangles;
the reference's data/buddha.obj is a missing blob) -- BASELINE.json configs[0], the configuration the metric is
quoted on.  One step = one bake of every vertex of the mesh.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

N > 1: one process per GPU, BVH replicated, vertices sharded in Morton-ordered interleaved chunks (strong scaling on
the fixed mesh), coefficient rows all-gathered with NCCL.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from prt_b200 import meshes  # noqa: E402

METRIC = "shadow_rays_per_sec"
UNIT = "rays/s"


# dram__bytes_read.sum + dram__bytes_write.sum of one bake_wave_kernel<3,true> launch on the bench workload, from the ncu --set full
# capture summarised in profiles/r1_final_ncu_summary.txt (259.8 MB read + 46.5 MB written).
NCU_TRAFFIC_BYTES = 305.0e6
# the resource that actually binds that kernel (same capture): smsp__issue_active.avg.pct_of_peak_sustained_active
NCU_ISSUE_ACTIVE_FRAC = 0.728

def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nu", type=int, default=737)
    ap.add_argument("--nv", type=int, default=737)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--samples-u", type=int, default=32)
    ap.add_argument("--samples-v", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=196608, help="vertices of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tune", default="", help="comma list name=value of prt_ctx_set_tuning knobs")
    return ap.parse_args()


def workload_name(a):
    return (f"shadowed diffuse PRT transfer, bumpy_torus {a.nu}x{a.nv} ({a.nu * a.nv} vertices / {2 * a.nu * a.nv} triangles, "
            f"buddha-scale), SH order {a.order} ({a.order ** 2} coeffs), {a.samples_u}x{a.samples_v}={a.samples_u * a.samples_v} "
            f"jittered stratified samples/vertex")


def base_config(a, n_gpus):
    return {"workload": workload_name(a), "vertices": a.nu * a.nv, "triangles": 2 * a.nu * a.nv, "sh_order": a.order,
            "samples_per_vertex": a.samples_u * a.samples_v, "parallelism": f"vertex-sharded x{n_gpus}, BVH replicated",
            "l2": "flushed between timed steps with a 512 MiB memset (untimed)"}


# ----------------------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle (the reference's own path cannot be built here: Embree/Eigen/GL absent) on all cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_bake_sample(a, scene_pos, tri, pos, nrm, n_sample, oscene=None):
    """Bakes a strided sample of the (Morton-ordered) vertices pos/nrm against the mesh (scene_pos, tri)."""
    from oracle import pyoracle
    if oscene is None:
        oscene = pyoracle.Scene(scene_pos, tri)
    stride = max(1, len(pos) // n_sample)
    sel = np.arange(0, len(pos), stride)[:n_sample]
    p = pyoracle.make_params(order=a.order, samples_u=a.samples_u, samples_v=a.samples_v)
    t0 = time.perf_counter()
    _, _, counters = pyoracle.bake_transfer(oscene, pos[sel], nrm[sel], p)
    dt = time.perf_counter() - t0
    return float(counters[0]) / dt, len(sel), dt, pyoracle.hw_threads(), oscene


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene_pos, nrm, tri = meshes.bumpy_torus(a.nu, a.nv)
    order = meshes.morton_order(scene_pos)
    pos, nrm = scene_pos[order], nrm[order]
    n_sample = min(len(pos), 32768)
    oscene = None
    for _ in range(max(1, min(a.warmup, 1))):
        _, _, _, cores, oscene = cpu_bake_sample(a, scene_pos, tri, pos, nrm, n_sample, oscene)
    times, rays = [], 0.0
    for _ in range(a.steps):
        rps, n_sel, dt, cores, oscene = cpu_bake_sample(a, scene_pos, tri, pos, nrm, n_sample, oscene)
        times.append(dt)
        rays += rps * dt
    total = sum(times)
    value = rays / total
    sample = f"{n_sample} Morton-strided vertices x {a.samples_u * a.samples_v} rays per step (of {len(pos)} vertices)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": base_config(a, a.gpus),
            "vertices_per_sec": value / (a.samples_u * a.samples_v),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU oracle (binary SAH BVH + scalar pinned Moeller-Trumbore, one trace per sample); "
                                     "the reference's Embree path cannot be built here"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev, period_ms=100):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(period_ms), "-i", str(dev)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "power_w_max": float(max(power)), "samples": len(sm)}
        return out


def shard_indices(n_verts, world, rank, chunk=64):
    """Interleaved chunks of the Morton-ordered vertex list; padded so every rank owns the same count."""
    n_chunks = (n_verts + chunk - 1) // chunk
    n_chunks_pad = ((n_chunks + world - 1) // world) * world
    idx = np.arange(n_chunks_pad * chunk, dtype=np.int64).reshape(n_chunks_pad, chunk)
    mine = idx[rank::world].reshape(-1)
    return np.minimum(mine, n_verts - 1), mine < n_verts, n_chunks_pad * chunk


def run_ours(a):
    import torch
    import torch.distributed as dist
    import prt_b200
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # this NCCL build prints its version banner on stdout unless NCCL_DEBUG is NONE: keep stdout to the one JSON line of the
        # contract (an NCCL_DEBUG set by the caller, e.g. INFO to look for NVLS, is respected)
        os.environ.setdefault("NCCL_DEBUG", "NONE")
        dist.init_process_group("nccl", device_id=dev)

    pos, nrm, tri = meshes.bumpy_torus(a.nu, a.nv)
    order = meshes.morton_order(pos)          # tri indices refer to the original numbering; the scene uses that
    ctx = prt_b200.Context(local)
    for kv in filter(None, a.tune.split(",")):
        k, v = kv.split("=")
        ctx.set_tuning(**{k: int(v)})
    scene = prt_b200.RTScene(pos, tri, ctx)
    info = scene.info()
    pos_m, nrm_m = pos[order], nrm[order]
    V = len(pos_m)
    S = a.samples_u * a.samples_v
    n2 = a.order ** 2
    params = prt_b200.BakeParams.make(order=a.order, samples_u=a.samples_u, samples_v=a.samples_v)

    mine, valid, v_pad = shard_indices(V, world, rank)
    n_mine = len(mine)
    h_pos = torch.from_numpy(np.ascontiguousarray(pos_m[mine])).pin_memory()
    h_nrm = torch.from_numpy(np.ascontiguousarray(nrm_m[mine])).pin_memory()
    d_pos, d_nrm = h_pos.to(dev), h_nrm.to(dev)
    d_all = torch.zeros((world, n_mine, n2), dtype=torch.float32, device=dev)
    d_out = d_all[rank]
    h_out = torch.zeros((world * n_mine, n2), dtype=torch.float32).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    L = ctx.L
    stream = torch.cuda.current_stream()

    def step_device():
        rc = L.prt_bake_transfer_device(ctx.h, scene.h, C.c_void_p(d_pos.data_ptr()), C.c_void_p(d_nrm.data_ptr()), 12, n_mine,
                                        0, C.byref(params), C.c_void_p(d_out.data_ptr()), None, C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(L.prt_last_error().decode())
        if world > 1:
            dist.all_gather_into_tensor(d_all.view(-1), d_out.reshape(-1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- algorithmic work of one launch (instrumented, untimed) ------------------------------------------------
    ctx.set_tuning(count_work=1)
    step_device(); torch.cuda.synchronize()
    st = ctx.last_bake_stats()
    visits, tests, cands, traversed = int(st.node_visits), int(st.tri_tests), int(st.cand_tests), int(st.rays_traversed)
    launches_per_step = int(st.launches)
    ctx.set_tuning(count_work=0)

    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()

    # --- device-resident timing: K steps, CUDA events per step, L2 flushed between steps ----------------------
    # every NVML query costs the running kernel ~0.1 ms (measured: 20 ms sampling made a 50 ms step 0.6 % slower), so the period is
    # 100 ms where the timed regions are long enough for that and 30 ms for the short steps of 4 and 8 GPUs
    clocks = ClockSampler(local, 100 if world <= 2 else 30) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for s0, s1 in ev:
        flush.zero_()
        if world > 1:
            dist.barrier()
        s0.record()
        step_device()
        s1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [s0.elapsed_time(s1) for s0, s1 in ev]
    # the dominant kernel's own duration (events inside the library, same stream) of the last step
    last = ctx.last_bake_stats()
    k_ms = last.kernel_ms - last.horizon_ms        # the dominant kernel (traversal + projection) alone
    hz_ms = last.horizon_ms
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    kms = torch.tensor([k_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    rays_per_step = float(V) * S
    value = rays_per_step * a.steps / (total_ms * 1e-3)

    # --- e2e: host buffers in, host rows out, through the C ABI ------------------------------------------------
    def step_e2e():
        if world == 1:
            rc = L.prt_bake_transfer(ctx.h, scene.h, C.c_void_p(h_pos.data_ptr()), C.c_void_p(h_nrm.data_ptr()), 12, n_mine, 0,
                                     C.byref(params), C.c_void_p(h_out.data_ptr()), None)
            if rc != 0:
                raise RuntimeError(L.prt_last_error().decode())
        else:
            d_pos.copy_(h_pos, non_blocking=True); d_nrm.copy_(h_nrm, non_blocking=True)
            step_device()
            if rank == 0:
                h_out.copy_(d_all.view(-1, n2), non_blocking=True)
            else:
                h_out[:n_mine].copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = rays_per_step * a.steps / float(e2e_s.item())
    # the clock sampler ran through both timed regions: at N = 8 the device-timed one alone lasts 35 ms
    clk = clocks.stop() if clocks else None
    h2d = world * n_mine * 24
    d2h = (world * n_mine * n2 * 4) + (world - 1) * n_mine * n2 * 4 if world > 1 else n_mine * n2 * 4

    # sanity: finished rows are finite and the DC term is a visibility fraction
    res = d_all.view(-1, n2)[:, 0]
    ok = bool(torch.isfinite(d_all).all().item()) and float(res.min().item()) >= -1e-6 and float(res.max().item()) <= 0.2821

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        k_ms_max = float(kms.item())
        rays_launch = float(n_mine) * S
        need_bytes = n_mine * (4.0 * ((S + 31) // 32) + 4.0) if launches_per_step >= 2 else 0.0
        alg_bytes = visits * 80.0 + tests * 48.0 + cands * 32.0 + n_mine * (24.0 + 4.0 * n2) + need_bytes
        achieved = alg_bytes / (k_ms_max * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": f"bake_wave_kernel<{a.order},true> (traversal + projection; the horizon pass ran {hz_ms:.2f} ms before it)", "achieved": achieved, "peak": hbm_peak,
                    "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": NCU_TRAFFIC_BYTES,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                    "kernel_ms": k_ms_max, "rays_per_launch": rays_launch,
                    "node_visits_per_ray": visits / rays_launch, "tri_tests_per_ray": tests / rays_launch,
                    "entry_list_box_tests_per_ray": cands / rays_launch, "rays_traversed_frac": traversed / rays_launch,
                    "horizon_pass_ms": hz_ms,
                    "binding_resource": {"name": "instruction issue slots", "frac": NCU_ISSUE_ACTIVE_FRAC,
                                         "source": "ncu smsp__issue_active of the same launch, profiles/r1_final_ncu_summary.txt"},
                    "algorithmic_bytes_per_ray": alg_bytes / rays_launch,
                    "hbm_floor_bytes_per_launch": n_mine * (24.0 + 4.0 * n2) + float(info.node_bytes + info.tri_bytes),
                    "note": "algorithmic bytes = node fetches x 80 B + triangle fetches x 48 B + entry-list boxes x 32 B (shared memory) + 60 B/vertex I/O + need bits; per-ray figures are averages over ALL rays (rays above the horizon map cost none); the BVH "
                            f"({(info.node_bytes + info.tri_bytes) / 1e6:.1f} MB) is L2-resident, so this traffic is served by L1/L2, "
                            "not HBM: the kernel is latency/issue bound, see profiles/"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": base_config(a, world),
                "vertices_per_sec": value / S, "results_ok": ok,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "vertices_per_sec": e2e_value / S},
                "gpu_launches": a.steps * world * launches_per_step, "roofline": roofline, "clocks": clk,
                "wall_s_timed_region": t_wall,
                "scene": {"nodes": int(info.n_nodes), "node_bytes": int(info.node_bytes), "tri_bytes": int(info.tri_bytes),
                          "max_depth": int(info.max_depth), "build_s": info.build_seconds, "upload_s": info.upload_seconds},
                "launch": {"grid": int(st.grid), "block": int(st.block)}}
        if world == 1 and not a.no_cpu_baseline:
            rps, n_sel, dt, cores, _ = cpu_bake_sample(a, pos, tri, pos_m, nrm_m, a.cpu_sample)
            line["cpu_baseline"] = {"value": rps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_sel} Morton-strided vertices x {S} rays ({dt:.1f} s of CPU work)"}
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = None


def emit_line(line):
    """The one JSON line of the contract goes to the process's ORIGINAL stdout (see main)."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line.  Libraries write there behind Python's back (NCCL prints its version banner with
    # printf when NCCL_DEBUG is VERSION/INFO in the environment), so file descriptor 1 is pointed at stderr for the whole run and
    # the JSON line is written to a private duplicate of the original stdout.
    global _JSON_OUT
    a = parse()
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
