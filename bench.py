#!/usr/bin/env python
"""bench.py -- the precomputation hot path of lvjiahui/PRT on B200s (BASELINE.json), one JSON line per run.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5|f1] [--workload torus|folds]
                  [--mesh file.obj] [--single-process] [--gather auto|p2p|nccl|none]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default (config 1 = BASELINE configs[0], the configuration the metric is quoted on): shadowed diffuse PRT transfer, SH order 3,
32 x 32 = 1024 jittered stratified samples per vertex, on a synthetic buddha-scale mesh (bumpy_torus 737 x 737: 543 169 vertices /
1 086 338 triangles; the reference's data/buddha.obj is a missing blob; `--mesh` takes a real OBJ).  One step = one bake of every
vertex of the mesh.  `--workload folds` is the heavily self-occluding twin (45 % of the rays occluded instead of 12 %).

Other BASELINE configurations, each with its own roofline / cpu_baseline / e2e:
  --config 2   environment pass: equirect -> 512^2 cube + mips, 32^2 irradiance, 256^2 x 5 GGX prefilter, 512^2 BRDF LUT, SH9
  --config 3   probe capture: 32^3 probes x 4096 closest-hit rays over the room + buddha-scale scene, CSR out
  --config 4   3-bounce interreflected transfer, order 4, 2.1 M vertices, 4096 samples per vertex
  --config 5   shadowed transfer, order 5, 8192 samples per vertex, 20 M vertices / 40 M triangles, vertex-sharded over the GPUs
  --config f1  calculate_weight: 8^3 probes, 96^3 voxels x (100 closest-hit + 8 any-hit) rays

N > 1 (configs 1, 4, 5; probes of config 3): one process per GPU under torchrun, BVH replicated, vertices sharded in Morton-ordered
interleaved 64-vertex chunks (strong scaling on the fixed mesh), rows all-gathered with NCCL.  `--single-process` runs the same
sharding through the C ABI's own multi-GPU driver (prt_group_*: fused P2P row stores or ncclAllGather inside the library).

The roofline is measured IN the run: rank 0 re-runs its own shard once under `ncu --metrics smsp__inst_executed.sum,dram__bytes_*`
(a child process; counts only, never timings) and divides by the CUDA-event duration of the timed launches.
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from prt_b200 import dist as pdist  # noqa: E402
from prt_b200 import meshes  # noqa: E402

BAKE_CONFIGS = {
    # name: nu, nv, order, su, sv, mode, bounces, albedo, default steps
    "1": dict(nu=737, nv=737, order=3, su=32, sv=32, mode="shadowed", bounces=0, albedo=(1.0, 1.0, 1.0), steps=5,
              title="shadowed diffuse PRT transfer (BASELINE configs[0])"),
    "4": dict(nu=1448, nv=1448, order=4, su=64, sv=64, mode="interreflect", bounces=3, albedo=(0.5, 0.5, 0.5), steps=2,
              title="3-bounce interreflected diffuse transfer (BASELINE configs[3])"),
    "5": dict(nu=5000, nv=4000, order=5, su=128, sv=64, mode="shadowed", bounces=0, albedo=(1.0, 1.0, 1.0), steps=2,
              title="shadowed diffuse PRT transfer, vertex-sharded (BASELINE configs[4])"),
}
NCU_METRICS = ("smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,"
               "gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum.pct_of_peak_sustained_elapsed,"
               "l1tex__t_bytes.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,"
               "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (0: the config's default)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="1", choices=["1", "2", "3", "4", "5", "f1"])
    ap.add_argument("--workload", default="torus", choices=["torus", "folds"], help="config 1 geometry: the friendly torus or its deeply folded twin")
    ap.add_argument("--mesh", default="", help="OBJ file baked instead of the synthetic mesh (configs 1, 4, 5)")
    ap.add_argument("--nu", type=int, default=0)
    ap.add_argument("--nv", type=int, default=0)
    ap.add_argument("--order", type=int, default=0)
    ap.add_argument("--samples-u", type=int, default=0)
    ap.add_argument("--samples-v", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="vertices / units of the bounded CPU-baseline sample (0: the config's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="skip the ncu child pass (roofline.achieved stays null)")
    ap.add_argument("--single-process", action="store_true", help="N GPUs through the C ABI's multi-GPU driver (prt_group_*) instead of torchrun")
    ap.add_argument("--gather", default="auto", choices=["auto", "p2p", "nccl", "none"])
    ap.add_argument("--tune", default="", help="comma list name=value of prt_ctx_set_tuning knobs")
    ap.add_argument("--ncu-child", action="store_true", help=argparse.SUPPRESS)
    a = ap.parse_args(argv)
    if a.config in BAKE_CONFIGS:
        c = BAKE_CONFIGS[a.config]
        a.nu, a.nv = a.nu or c["nu"], a.nv or c["nv"]
        a.order, a.samples_u, a.samples_v = a.order or c["order"], a.samples_u or c["su"], a.samples_v or c["sv"]
        a.steps = a.steps or c["steps"]
    else:
        a.steps = a.steps or 5
    return a


# ----------------------------------------------------------------------------------------------------------------
# shared pieces
# ----------------------------------------------------------------------------------------------------------------
_JSON_OUT = None


def emit_line(line):
    """The one JSON line of the contract goes to the process's ORIGINAL stdout (see main)."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


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


def ncu_pass(a, kernel_regex, skip, count=1, child_gpus=None, extra_env=None):
    """Runs this script once more as a child under ncu (this rank's own shard, one GPU) and returns the counters of `count` launches
    of the kernels matching `kernel_regex` after skipping `skip` of them: {metric: value}, `.sum` metrics summed over the captured
    launches and percentages averaged -- or {"error": ...}."""
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {"error": "ncu not found"}
    log_file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    log_file.close()
    child = [sys.executable, os.path.abspath(__file__), "--ncu-child", "--gpus", str(child_gpus or a.gpus), "--config", a.config, "--workload", a.workload,
             "--nu", str(a.nu), "--nv", str(a.nv), "--order", str(a.order), "--samples-u", str(a.samples_u), "--samples-v", str(a.samples_v)]
    if a.mesh:
        child += ["--mesh", a.mesh]
    if a.tune:
        child += ["--tune", a.tune]
    cmd = [ncu, "--metrics", NCU_METRICS, "--clock-control", "none", "-k", f"regex:{kernel_regex}", "--launch-skip", str(skip), "--launch-count", str(count),
           "--csv", "--log-file", log_file.name] + child
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    env.update(extra_env or {})
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=900)
    except (OSError, subprocess.TimeoutExpired) as e:
        return {"error": f"ncu child: {e}"}
    out = {"seconds": time.perf_counter() - t0, "command": " ".join(cmd[:12]) + " ... bench.py --ncu-child"}
    try:
        rows = [r_ for r_ in csv.reader(open(log_file.name)) if r_]
    except OSError:
        rows = []
    os.unlink(log_file.name)
    hdr = next((i for i, r_ in enumerate(rows) if "Metric Name" in r_), None)
    if r.returncode != 0 or hdr is None:
        out["error"] = f"ncu rc={r.returncode}: " + (r.stderr or "")[-300:].replace("\n", " | ")
        return out
    names = rows[hdr]
    seen = {}
    i_name, i_val, i_unit, i_k = names.index("Metric Name"), names.index("Metric Value"), names.index("Metric Unit"), names.index("Kernel Name")
    for r_ in rows[hdr + 1:]:
        if len(r_) <= i_val:
            continue
        try:
            v = float(r_[i_val].replace(",", ""))
        except ValueError:
            continue
        unit = r_[i_unit]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "us": 1e-6,
                 "ms": 1e-3, "ns": 1e-9}.get(unit, 1.0)
        name = r_[i_name]
        if name.endswith(".sum"):
            out[name] = out.get(name, 0.0) + v * scale
        else:
            seen[name] = seen.get(name, 0) + 1
            out[name] = out.get(name, 0.0) + (v * scale - out.get(name, 0.0)) / seen[name]
        out["kernel"] = r_[i_k][:120]
    if "smsp__inst_executed.sum" not in out:
        out["error"] = "ncu produced no counters (kernel regex matched nothing?)"
    return out


def issue_roofline(ncu, kernel_ms, n_sms, sm_mhz, peaks, floor_bytes, l2_alg_bytes, what):
    """roofline object for an instruction-issue-bound kernel: achieved = warp instructions of the launch (ncu count) / live CUDA-event
    duration; peak = SMs x 4 schedulers x measured SM clock (one warp instruction per scheduler and cycle).  The memory side is
    reported next to it: true DRAM bytes of the launch against the measured HBM copy peak, and the algorithmic L2-side bytes."""
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    sm_ghz = (sm_mhz or peaks.get("sm_max_mhz") or 1965.0) / 1e3
    issue_peak = n_sms * 4 * sm_ghz                       # G warp-instructions / s
    r = {"bound": "issue", "kernel": what, "achieved": None, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": None, "traffic": None,
         "kernel_ms": kernel_ms, "peak_source": f"{n_sms} SMs x 4 schedulers x {sm_ghz:.3f} GHz (SM clock sampled during the timed region)",
         "hbm": {"peak_gbs": hbm_peak, "peak_source": peak_src, "floor_bytes_per_launch": floor_bytes},
         "l2_algorithmic": {"bytes_per_launch": l2_alg_bytes, "gbs": l2_alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms else None,
                            "frac_of_hbm_peak": l2_alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak if kernel_ms else None,
                            "note": "node / triangle / texel fetches the algorithm asks for; served by L1 / L2, not HBM"}}
    if ncu is None or "error" in (ncu or {}):
        r["ncu_error"] = (ncu or {}).get("error", "ncu pass disabled")
        return r
    inst = ncu["smsp__inst_executed.sum"]
    dram = ncu.get("dram__bytes_read.sum", 0.0) + ncu.get("dram__bytes_write.sum", 0.0)
    t = kernel_ms * 1e-3
    r["achieved"] = inst / t / 1e9
    r["frac"] = r["achieved"] / issue_peak
    r["traffic"] = dram
    r["warp_instructions_per_launch"] = inst
    r["active_threads_per_warp_instruction"] = ncu.get("smsp__thread_inst_executed.sum", 0.0) / inst if inst else None
    r["hbm"].update({"achieved_gbs": dram / t / 1e9, "frac": dram / t / 1e9 / hbm_peak, "traffic_over_floor": dram / floor_bytes if floor_bytes else None})
    if r["hbm"]["frac"] > r["frac"]:
        r["bound"] = "hbm"
    r["l2_measured"] = {"bytes_per_launch": ncu.get("lts__t_bytes.sum"), "gbs": ncu.get("lts__t_bytes.sum", 0.0) / t / 1e9,
                        "pct_of_peak_under_ncu": ncu.get("lts__t_bytes.sum.pct_of_peak_sustained_elapsed")}
    r["pipes_pct_under_ncu"] = {"issue_active": ncu.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                "xu": ncu.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                                "alu": ncu.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                                "fma": ncu.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")}
    r["source"] = ("counts (instructions, DRAM / L2 bytes) from an ncu pass over this rank's own shard driven by this run (" + ncu.get("kernel", "?") +
                   f", {ncu.get('seconds', 0):.0f} s); durations from CUDA events of the timed launches; percentages marked under_ncu are context only")
    return r


def make_mesh(a):
    if a.mesh:
        pos, nrm, tri = meshes.load_obj_assimp(a.mesh)
        return pos, nrm, tri, f"{os.path.basename(a.mesh)} ({len(pos)} vertices / {len(tri)} triangles, assimp vertex semantics)"
    if a.workload == "folds":
        pos, nrm, tri = meshes.bumpy_torus(a.nu, a.nv, amp=0.25, fscale=3)
        return pos, nrm, tri, f"folded bumpy_torus {a.nu}x{a.nv} amp 0.25 x 3 frequencies ({len(pos)} vertices / {len(tri)} triangles, heavily self-occluding)"
    pos, nrm, tri = meshes.bumpy_torus(a.nu, a.nv)
    scale = "buddha-scale" if (a.nu, a.nv) == (737, 737) else f"{len(pos) / 1e6:.1f} M vertices"
    return pos, nrm, tri, f"bumpy_torus {a.nu}x{a.nv} ({len(pos)} vertices / {len(tri)} triangles, {scale})"


def bake_params(a, mod):
    c = BAKE_CONFIGS[a.config]
    mode = mod.INTERREFLECT if c["mode"] == "interreflect" else mod.SHADOWED
    kw = dict(order=a.order, samples_u=a.samples_u, samples_v=a.samples_v, mode=mode, bounces=c["bounces"], albedo=c["albedo"])
    return mod.BakeParams.make(**kw) if hasattr(mod, "BakeParams") and hasattr(mod.BakeParams, "make") else mod.make_params(**kw)


def bake_config(a, n_gpus, mesh_name, V, T):
    c = BAKE_CONFIGS[a.config]
    S = a.samples_u * a.samples_v
    what = c["title"] + (f", {c['bounces']} bounces, albedo {c['albedo'][0]}" if c["bounces"] else "")
    return {"workload": f"{what}: {mesh_name}, SH order {a.order} ({a.order ** 2} coeffs), {a.samples_u}x{a.samples_v}={S} jittered stratified samples/vertex",
            "baseline_config": a.config, "vertices": V, "triangles": T, "sh_order": a.order, "samples_per_vertex": S,
            "parallelism": f"vertex-sharded x{n_gpus}, BVH replicated" + (" (single process, prt_group_*)" if a.single_process and n_gpus > 1 else ""),
            "l2": "flushed between timed steps with a 512 MiB memset (untimed)"}


# ----------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle (the reference's own path cannot be built here: Embree / Eigen / GL absent) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_bake_sample(a, scene_pos, tri, pos, nrm, n_sample, oscene=None, want_vis=False):
    """Bakes a strided sample of the (Morton-ordered) vertices pos/nrm against the mesh (scene_pos, tri) with the oracle on all cores."""
    from oracle import pyoracle
    if oscene is None:
        oscene = pyoracle.Scene(scene_pos, tri)
    stride = max(1, len(pos) // n_sample)
    sel = np.arange(0, len(pos), stride)[:n_sample]
    p = bake_params(a, pyoracle)
    t0 = time.perf_counter()
    rows, vis, counters = pyoracle.bake_transfer(oscene, pos[sel], nrm[sel], p, want_vis=want_vis, vertex_id_base=0)
    dt = time.perf_counter() - t0
    return {"rays": float(len(sel)) * a.samples_u * a.samples_v, "segments": float(counters[1]), "seconds": dt, "sel": sel, "rows": rows, "vis": vis,
            "cores": pyoracle.hw_threads(), "oscene": oscene}


def default_cpu_sample(a):
    if a.cpu_sample:
        return a.cpu_sample
    # ~10-20 s of oracle work on 16 cores
    return {"1": 196608, "4": 16384, "5": 12288}[a.config]


def run_reference_bake(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    scene_pos, nrm, tri, mesh_name = make_mesh(a)
    order = meshes.morton_order(scene_pos)
    pos, nrm = scene_pos[order], nrm[order]
    n_sample = min(len(pos), max(256, default_cpu_sample(a) // 6))
    oscene, cores = None, 1
    for _ in range(max(1, min(a.warmup, 1))):
        r = cpu_bake_sample(a, scene_pos, tri, pos, nrm, n_sample, oscene)
        oscene, cores = r["oscene"], r["cores"]
    total, rays = 0.0, 0.0
    for _ in range(a.steps):
        r = cpu_bake_sample(a, scene_pos, tri, pos, nrm, n_sample, oscene)
        total += r["seconds"]; rays += r["rays"]
    value = rays / total
    S = a.samples_u * a.samples_v
    sample = f"{n_sample} Morton-strided vertices x {S} primary rays per step (of {len(pos)} vertices)"
    metric, unit = bake_metric(a)
    emit_line({"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": bake_config(a, a.gpus, mesh_name, len(pos), len(tri)), "vertices_per_sec": value / S,
               "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                                "note": "CPU oracle (binary SAH BVH + scalar pinned Moeller-Trumbore, one trace per sample); the reference's Embree path cannot be built here"},
               "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def bake_metric(a):
    return ("primary_paths_per_sec", "paths/s") if BAKE_CONFIGS[a.config]["mode"] == "interreflect" else ("shadow_rays_per_sec", "rays/s")


# ----------------------------------------------------------------------------------------------------------------
# configs 1 / 4 / 5: per-vertex transfer bake
# ----------------------------------------------------------------------------------------------------------------
def run_bake(a):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import prt_b200

    child = a.ncu_child
    group_mode = a.single_process and a.gpus > 1 and not child
    world = 1 if (group_mode) else (a.gpus if child else int(os.environ.get("WORLD_SIZE", "1")))
    rank = 0 if child else int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1 and not child
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "NONE")          # this NCCL build prints its version banner on stdout otherwise
        dist.init_process_group("nccl", device_id=dev)

    t0 = time.perf_counter()
    pos, nrm, tri, mesh_name = make_mesh(a)
    order = meshes.morton_order(pos)          # tri indices refer to the original numbering; the scene uses that
    pos_m, nrm_m = pos[order], nrm[order]
    t_mesh = time.perf_counter() - t0
    V, S, n2 = len(pos_m), a.samples_u * a.samples_v, a.order ** 2
    params = bake_params(a, prt_b200)
    metric, unit = bake_metric(a)
    inter = BAKE_CONFIGS[a.config]["mode"] == "interreflect"

    if group_mode:
        return run_bake_group(a, prt_b200, pos, tri, pos_m, nrm_m, params, mesh_name, t_mesh)

    ctx = prt_b200.Context(local)
    for kv in filter(None, a.tune.split(",")):
        k, v = kv.split("=")
        ctx.set_tuning(**{k: int(v)})
    scene = prt_b200.RTScene(pos, tri, ctx)
    info = scene.info()

    mine, valid, v_pad = pdist.shard_indices(V, world, rank)
    n_mine = len(mine)
    n_valid = int(valid.sum())                # valid vertices are a prefix of the shard (padding sits at the end of the list)
    h_pos = torch.from_numpy(np.ascontiguousarray(pos_m[mine])).pin_memory()
    h_nrm = torch.from_numpy(np.ascontiguousarray(nrm_m[mine])).pin_memory()
    d_pos, d_nrm = h_pos.to(dev), h_nrm.to(dev)
    d_all = torch.zeros((world, n_mine, n2), dtype=torch.float32, device=dev)
    d_out = d_all[rank]
    h_out = torch.zeros((n_mine, n2), dtype=torch.float32).pin_memory()
    # N > 1: the gather is fused into the kernels (rows stored straight into every rank's full-size buffer through CUDA-IPC-mapped
    # peer pointers) unless --gather nccl asks for the all-gather baseline or IPC is unavailable
    fused, peers, peer_arr, d_full, full_buf = False, [], None, None, None
    if use_dist and a.gather in ("auto", "p2p"):
        try:
            full_buf = prt_b200.DeviceBuffer(ctx, v_pad * n2)
            handles = [None] * world
            dist.all_gather_object(handles, full_buf.export())
            peers = [prt_b200.DeviceBuffer.open(ctx, handles[r], v_pad * n2) for r in range(world) if r != rank]
            peer_arr = (C.c_void_p * len(peers))(*[p_.ptr for p_ in peers])
            d_full = torch.as_tensor(full_buf, device=dev).view(v_pad, n2)
            fused = True
        except Exception as e:                          # noqa: BLE001 -- any failure means: use the NCCL baseline
            log(f"rank {rank}: fused IPC gather unavailable ({e}); using NCCL all-gather")
        ok_all = torch.tensor([1.0 if fused else 0.0], device=dev)
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        fused = bool(ok_all.item() > 0.5)
    if fused:
        d_own = d_full.view(v_pad // (world * 64), world, 64, n2)[:, rank]          # this rank's rows inside the full buffer (strided view)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    L = ctx.L
    stream = torch.cuda.current_stream()

    def bake_shard(n=n_mine, out=d_out):
        rc = L.prt_bake_transfer_device_shard(ctx.h, scene.h, C.c_void_p(d_pos.data_ptr()), C.c_void_p(d_nrm.data_ptr()), 12, n, world, rank,
                                              C.byref(params), C.c_void_p(out.data_ptr()), None, C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(L.prt_last_error().decode())

    def bake_fused():
        rc = L.prt_bake_transfer_device_shard_fused(ctx.h, scene.h, C.c_void_p(d_pos.data_ptr()), C.c_void_p(d_nrm.data_ptr()), 12, n_mine, world, rank,
                                                    C.byref(params), C.c_void_p(full_buf.ptr), peer_arr, len(peers), C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(L.prt_last_error().decode())

    def step_device():
        if fused:
            bake_fused()                    # rows land in every rank's buffer from the kernels' epilogue: nothing follows
            return
        bake_shard()
        if use_dist:
            dist.all_gather_into_tensor(d_all.view(-1), d_out.reshape(-1))

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    if child:
        # under ncu: launch 1 = warm-up (module load), launch 2 = the one profiled
        bake_shard(); torch.cuda.synchronize()
        flush.zero_()
        bake_shard(); torch.cuda.synchronize()
        return

    # --- algorithmic work of one launch (instrumented, untimed) ------------------------------------------------
    ctx.set_tuning(count_work=1)
    step_device(); torch.cuda.synchronize()
    st = ctx.last_bake_stats()
    visits, tests, cands, traversed = int(st.node_visits), int(st.tri_tests), int(st.cand_tests), int(st.rays_traversed)
    launches_per_step = int(st.launches)
    ctx.set_tuning(count_work=0)

    W_ = max(a.warmup, 3)
    for _ in range(W_):
        step_device()
    barrier()

    # --- device-resident timing: K steps, CUDA events per step, L2 flushed between steps ----------------------
    # every NVML query costs the running kernel ~0.1 ms, so the period is 100 ms for long steps and 30 ms for the short steps of 4 / 8 GPUs
    clocks = ClockSampler(local, 100 if (world <= 2 or a.config != "1") else 30) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    k_ms_steps, hz_ms_steps = [], []
    for s0, s1, s2 in ev:
        flush.zero_()
        if use_dist:
            dist.barrier()
        s0.record()
        if fused:
            bake_fused()
        else:
            bake_shard()
        s1.record()
        if use_dist and not fused:
            dist.all_gather_into_tensor(d_all.view(-1), d_out.reshape(-1))
        s2.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [s0.elapsed_time(s2) for s0, s1, s2 in ev]
    gather_ms = [s1.elapsed_time(s2) for s0, s1, s2 in ev]
    last = ctx.last_bake_stats()              # events inside the library around the kernels of the last step
    k_ms, hz_ms = last.kernel_ms - last.horizon_ms, last.horizon_ms
    red = torch.tensor([sum(step_ms), k_ms, float(np.mean(gather_ms))], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    total_ms, k_ms_max, gather_ms_max = (float(x) for x in red.tolist())
    units_per_step = float(V) * S
    value = units_per_step * a.steps / (total_ms * 1e-3)

    # --- e2e: host buffers in, host rows out ------------------------------------------------------------------
    # N = 1: the reference-facing C-ABI call with host pointers.  N > 1: every rank uploads its shard from pinned memory, bakes,
    # all-gathers on the device and reads ITS OWN rows back (the host result is assembled from the per-rank buffers, no rank moves
    # the whole array over one PCIe link)
    def step_e2e():
        if world == 1:
            rc = L.prt_bake_transfer(ctx.h, scene.h, C.c_void_p(h_pos.data_ptr()), C.c_void_p(h_nrm.data_ptr()), 12, n_mine, 0,
                                     C.byref(params), C.c_void_p(h_out.data_ptr()), None)
            if rc != 0:
                raise RuntimeError(L.prt_last_error().decode())
        else:
            d_pos.copy_(h_pos, non_blocking=True); d_nrm.copy_(h_nrm, non_blocking=True)
            step_device()
            h_out.view(-1, 64, n2).copy_(d_own if fused else d_out.view(-1, 64, n2), non_blocking=True)
            torch.cuda.synchronize()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step * a.steps / float(e2e_s.item())
    clk = clocks.stop() if clocks else None
    h2d = world * n_mine * 24
    d2h = world * n_mine * n2 * 4

    # --- results: sanity + parity -------------------------------------------------------------------------------
    step_device(); barrier()                                      # rows of a device step (the e2e loop left the same rows)
    rows_all = d_full[:V] if fused else d_all.view(-1, n2)       # fused: vertex order; all-gather: [rank][local]
    res = rows_all[:, 0]
    hi = 0.2821 if not inter else 0.2821 * 1.0001
    ok = bool(torch.isfinite(rows_all).all().item()) and float(res.min().item()) >= -1e-6 and float(res.max().item()) <= hi
    parity = None
    if world > 1:
        # every rank: its own rows of the gathered array are what it baked (bitwise); rank 0: a Morton-strided sample of the whole
        # list re-baked by ONE GPU as an unsharded bake must equal the sharded rows bit for bit
        # (fused: every rank compares its full buffer with rank 0's after the barrier -- the peers' stores all arrived)
        if fused:
            chk = d_full[:V].clone()
            dist.broadcast(chk, src=0)
            same = bool(torch.equal(chk, d_full[:V]))
            del chk
        else:
            same = bool(torch.equal(d_all[rank], d_out))
        n_chk = min(V, 65536 if a.config == "1" else 4096)
        stride = max(1, V // n_chk)
        sel = np.arange(0, V, stride)[:n_chk]
        if rank == 0:
            full = d_full[:V].cpu().numpy() if fused else pdist.unshard_rows(d_all.cpu().numpy(), V, world)
            if inter:
                # the bounce RNG is keyed by the list position: bake the sample as singleton "lists" via vertex_id_base
                sp = np.ascontiguousarray(pos_m[sel]); sn = np.ascontiguousarray(nrm_m[sel])
                one = np.concatenate([prt_b200.bake_transfer(scene, sp[i:i + 1], sn[i:i + 1], params, vertex_id_base=int(sel[i]))[0] for i in range(min(len(sel), 256))])
                sel = sel[:len(one)]
            else:
                one, _ = prt_b200.bake_transfer(scene, pos_m[sel], nrm_m[sel], params)
            bit_equal = bool(np.array_equal(one.view(np.uint32), full[sel].view(np.uint32)))
            parity = {"vs": "unsharded single-GPU bake of a Morton-strided sample", "rows_bit_equal": bit_equal,
                      "all_ranks_hold_rank0s_rows" if fused else "own_rows_bit_equal": same,
                      "n_vertices": int(len(sel))}
            ok = ok and bit_equal and same
        flag = torch.tensor([1.0 if (same and ok) else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item() > 0.5)

    if rank == 0:
        peaks = load_peaks()
        n_sms = torch.cuda.get_device_properties(local).multi_processor_count
        kname = "bake_inter_kernel" if inter else "bake_wave_kernel"
        ncu = None
        if not a.no_ncu:
            # big shards: the child profiles every f-th 64-vertex chunk of this rank's shard (it runs as rank 0 of world x f ranks) and
            # the counts are scaled by the vertex ratio; f = 1 (the exact launch) up to 3 M vertices or with PRT_BENCH_NCU_FULL=1
            f = 1 if os.environ.get("PRT_BENCH_NCU_FULL") else max(1, -(-n_valid // 3_000_000))
            ncu = ncu_pass(a, kname, 1, child_gpus=world * f)
            if "error" in ncu:
                log("ncu pass:", ncu["error"])
            elif f > 1:
                n_child = int(pdist.shard_indices(V, world * f, 0)[1].sum())
                k = n_valid / n_child
                for key in list(ncu):
                    if key.endswith(".sum"):
                        ncu[key] *= k
                ncu["kernel"] = ncu.get("kernel", "") + f" [counts of every {f}-th chunk of the shard x {k:.3f}]"
        rays_launch = float(n_valid) * S
        need_bytes = n_mine * (4.0 * ((S + 31) // 32) + 4.0) if launches_per_step >= 2 else 0.0
        shadowed = kname.startswith("bake_wave")                     # its node test also fetches the 32-byte fourth-axis record of the node
        node_fetch = 112.0 if shadowed else 80.0
        bvh_bytes = float(info.node_bytes + info.tri_bytes) + (32.0 * info.n_nodes if shadowed else 0.0)
        l2_alg = visits * node_fetch + tests * 48.0
        floor = n_valid * (24.0 + 4.0 * n2) + need_bytes + min(bvh_bytes, l2_alg)
        roofline = issue_roofline(ncu, k_ms_max, n_sms, clk.get("sm_mhz") if clk else None, peaks, floor, l2_alg,
                                  f"{kname}<{a.order}> (traversal + projection of this rank's shard; the horizon pass ran {hz_ms:.2f} ms before it)")
        roofline.update({"rays_per_launch": rays_launch, "node_visits_per_ray": visits / rays_launch, "tri_tests_per_ray": tests / rays_launch,
                         "entry_list_box_tests_per_ray": cands / rays_launch, "rays_traversed_frac": traversed / rays_launch if traversed else None,
                         "horizon_pass_ms": hz_ms,
                         "note": "hbm.floor = vertex I/O + need bits + every BVH byte the launch touches once (capped at the BVH size); l2_algorithmic = node "
                                 f"fetches x {node_fetch:.0f} B + triangle fetches x 48 B; the BVH is {bvh_bytes / 1e6:.1f} MB ("
                                 + ("L2-resident" if bvh_bytes < 100e6 else "larger than the 126 MB L2: deep levels stream from HBM") + ")"})
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": W_,
                "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic" if not a.mesh else "file", "config": bake_config(a, world, mesh_name, V, len(tri)),
                "vertices_per_sec": value / S, "results_ok": ok,
                "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "vertices_per_sec": e2e_value / S},
                "gpu_launches": a.steps * world * launches_per_step, "roofline": roofline, "clocks": clk,
                "wall_s_timed_region": t_wall,
                "traversed": {"rays_traversed_frac": traversed / rays_launch if traversed else None,
                              "traversed_rays_per_sec": (traversed / rays_launch) * value if traversed else None,
                              "note": "rays the horizon pass could not prove free; every sample counts as one ray in `value`, as the reference casts one"},
                "scene": {"nodes": int(info.n_nodes), "node_bytes": int(info.node_bytes), "tri_bytes": int(info.tri_bytes),
                          "max_depth": int(info.max_depth), "build_s": info.build_seconds, "upload_s": info.upload_seconds, "mesh_s": t_mesh},
                "launch": {"grid": int(st.grid), "block": int(st.block)}}
        if world > 1:
            line["all_gather"] = {"mode": "fused into the kernels: P2P row stores into every rank's buffer through CUDA IPC (no collective)" if fused
                                  else "NCCL all_gather_into_tensor after the kernels",
                                  "ms": gather_ms_max, "bytes_received_per_gpu": (world - 1) * n_mine * n2 * 4,
                                  "gbs": (world - 1) * n_mine * n2 * 4 / (gather_ms_max * 1e-3) / 1e9 if (gather_ms_max > 0 and not fused) else None}
            line["parity"] = parity
        if world == 1 and not a.no_cpu_baseline:
            # the oracle's rows are kept: parity of the benchmarked launch at the benchmarked size, not just a CPU timing
            n_cpu = default_cpu_sample(a)
            r = cpu_bake_sample(a, pos, tri, pos_m, nrm_m, n_cpu, want_vis=True)
            sel = r["sel"]
            if inter:
                # bounce RNG keyed by list position: the oracle baked the sample as a list of its own -> bake the same list on the GPU
                g_rows, g_vis = prt_b200.bake_transfer(scene, pos_m[sel], nrm_m[sel], params, want_vis=True)
            else:
                _, g_vis = prt_b200.bake_transfer(scene, pos_m[sel], nrm_m[sel], params, want_vis=True)
                g_rows = rows_all[torch.from_numpy(sel).to(dev)].cpu().numpy()           # rows of the timed launch itself
            nr = np.linalg.norm(r["rows"], axis=1)
            rel = np.linalg.norm(g_rows - r["rows"], axis=1) / np.maximum(nr, 1e-20)
            big = nr > 1e-3
            parity = {"vs": "CPU oracle on the cpu_baseline sample", "vis_bits_equal": bool(np.array_equal(g_vis, r["vis"])),
                      "max_rel_l2": float(rel[big].max(initial=0.0)), "max_abs_small_rows": float(np.abs(g_rows - r["rows"])[~big].max(initial=0.0)),
                      "tolerance_rel_l2": 1e-4, "n_vertices": int(len(sel)), "rays_compared": int(len(sel)) * S}
            line["parity"] = parity
            line["results_ok"] = bool(ok and parity["vis_bits_equal"] and parity["max_rel_l2"] <= 1e-4 and parity["max_abs_small_rows"] <= 1e-6)
            line["cpu_baseline"] = {"value": r["rays"] / r["seconds"], "unit": unit, "cores": r["cores"], "kind": "port",
                                    "sample": f"{len(sel)} Morton-strided vertices x {S} primary rays ({r['seconds']:.1f} s of CPU work)"}
        emit_line(line)
    if use_dist:
        dist.barrier()
        for p_ in peers:
            p_.close()
        dist.barrier()
        if full_buf is not None:
            d_full = None
            full_buf.close()
        dist.destroy_process_group()


def run_bake_group(a, prt, pos, tri, pos_m, nrm_m, params, mesh_name, t_mesh):
    """N GPUs in ONE process through the C ABI's multi-GPU driver (prt_group_*, prt_b200/csrc/group.cu)."""
    import torch
    V, S, n2 = len(pos_m), a.samples_u * a.samples_v, a.order ** 2
    metric, unit = bake_metric(a)
    grp = prt.Group(list(range(a.gpus)))
    for kv in filter(None, a.tune.split(",")):
        k, v = kv.split("=")
        grp.set_tuning(**{k: int(v)})
    info = grp.set_scene(pos, tri)
    caps = grp.capabilities()
    mode = {"auto": prt.GATHER_AUTO, "p2p": prt.GATHER_P2P, "nccl": prt.GATHER_NCCL, "none": prt.GATHER_NONE}[a.gather]
    h_pos = torch.from_numpy(np.ascontiguousarray(pos_m)).pin_memory()
    h_nrm = torch.from_numpy(np.ascontiguousarray(nrm_m)).pin_memory()
    h_out = torch.zeros((V, n2), dtype=torch.float32).pin_memory()
    hp, hn, ho = h_pos.numpy(), h_nrm.numpy(), h_out.numpy()
    W_ = max(a.warmup, 3)
    for _ in range(W_):
        grp.bake_transfer(hp, hn, params, gather=mode, out=ho)
    clocks = ClockSampler(0, 100)
    dev_ms, wall_ms, gath_ms, k_ms = [], [], [], []
    flush = [torch.empty(512 << 20, dtype=torch.uint8, device=torch.device("cuda", i)) for i in range(a.gpus)]
    st = None
    for _ in range(a.steps):
        for f in flush:
            f.zero_()
        for i in range(a.gpus):
            torch.cuda.synchronize(i)
        t0 = time.perf_counter()
        _, st = grp.bake_transfer(hp, hn, params, gather=mode, out=ho)
        wall_ms.append(1e3 * (time.perf_counter() - t0))
        dev_ms.append(max(st.kernel_ms[i] + st.gather_ms[i] for i in range(a.gpus)))      # device events, max over the GPUs
        k_ms.append(st.kernel_ms_max); gath_ms.append(st.gather_ms_max)
    clk = clocks.stop()
    units = float(V) * S
    value = units * a.steps / (sum(dev_ms) * 1e-3)
    e2e_value = units * a.steps / (sum(wall_ms) * 1e-3)
    # parity: the sharded rows (host) and the rows gathered on the LAST member GPU against an unsharded single-GPU bake of a sample
    n_chk = min(V, 65536 if a.config == "1" else 4096)
    sel = np.arange(0, V, max(1, V // n_chk))[:n_chk]
    _, sc0 = grp.member(0)                       # the group's own scene copy on GPU 0: no second BVH build
    inter = BAKE_CONFIGS[a.config]["mode"] == "interreflect"
    if inter:
        sel = sel[:256]
        one = np.concatenate([prt.bake_transfer(sc0, pos_m[i:i + 1], nrm_m[i:i + 1], params, vertex_id_base=int(i))[0] for i in sel])
    else:
        one, _ = prt.bake_transfer(sc0, pos_m[sel], nrm_m[sel], params)
    bit_equal = bool(np.array_equal(one.view(np.uint32), ho[sel].view(np.uint32)))
    gathered_equal = None
    if st.gather_mode != prt.GATHER_NONE:
        gathered_equal = bool(np.array_equal(grp.download_rows(a.gpus - 1, V, n2)[sel].view(np.uint32), one.view(np.uint32)))
    ok = bool(np.isfinite(ho).all()) and bit_equal and gathered_equal is not False
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": a.gpus, "steps": a.steps, "warmup": W_, "ms_per_step": sum(dev_ms) / a.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic" if not a.mesh else "file",
            "config": bake_config(a, a.gpus, mesh_name, V, len(tri)), "vertices_per_sec": value / S, "results_ok": ok,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(st.h2d_bytes), "d2h_bytes_per_step": int(st.d2h_bytes), "vertices_per_sec": e2e_value / S,
                    "ms_per_step": sum(wall_ms) / a.steps, "note": "prt_group_bake_transfer with pinned host buffers: per-GPU H2D of the shard, bake, gather, per-GPU D2H of its rows"},
            "gpu_launches": a.steps * a.gpus * 2, "clocks": clk,
            "driver": {"kind": "single process, C ABI prt_group_*", "gather_mode": {0: "none", 1: "nccl", 2: "p2p (fused row stores)"}[st.gather_mode],
                       "capabilities": caps, "kernel_ms_max": float(np.mean(k_ms)), "gather_ms_max": float(np.mean(gath_ms)),
                       "gather_bytes_per_gpu": int(st.gather_bytes_per_gpu), "kernel_ms_per_gpu": [round(st.kernel_ms[i], 3) for i in range(a.gpus)],
                       "vertices_per_gpu": [int(st.vertices[i]) for i in range(a.gpus)]},
            "parity": {"vs": "unsharded single-GPU bake of a Morton-strided sample", "rows_bit_equal": bit_equal, "gathered_rows_on_last_gpu_bit_equal": gathered_equal,
                       "n_vertices": int(len(sel))},
            "roofline": {"bound": "issue", "achieved": None, "peak": None, "unit": "Gwarp-inst/s", "frac": None, "traffic": None,
                         "note": "same kernels as the torchrun arm; see its line for the ncu-driven roofline"},
            "scene": {"nodes": int(info.n_nodes), "node_bytes": int(info.node_bytes), "tri_bytes": int(info.tri_bytes), "max_depth": int(info.max_depth),
                      "build_s": info.build_seconds, "upload_s_slowest_gpu": info.upload_seconds, "mesh_s": t_mesh}}
    emit_line(line)
    grp.close()


# ----------------------------------------------------------------------------------------------------------------
# config 2: environment pass
# ----------------------------------------------------------------------------------------------------------------
def env_image():
    from prt_b200 import hdr
    p = os.path.join(ROOT, "tests", "golden", "newport_loft.hdr")            # the reference's data/hdr/newport_loft.hdr, a committed fixture
    if os.path.exists(p):
        return hdr.load_hdr(p), "data/hdr/newport_loft.hdr (1600x800)"
    return hdr.synthetic_env(1600, 800), "synthetic 1600x800 HDR (range of newport_loft)"


ENV_SAMPLES = 523776 * 1024 + 262144 * 1024 + 6144 * 15876       # prefilter + LUT + irradiance sample evaluations per step


def run_env(a):
    import torch

    import prt_b200
    child = a.ncu_child
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(0)
    ctx = prt_b200.Context(0)
    eq, eq_name = env_image()
    h_eq = torch.from_numpy(np.ascontiguousarray(eq)).pin_memory().numpy()
    lp = prt_b200.LightProbe(h_eq, 512, ctx)
    L = ctx.L
    import ctypes as C

    def device_step():
        """all passes with the results left on the device; returns the summed kernel milliseconds (CUDA events around each pass)"""
        ms = 0.0
        for call in (lambda: L.prt_env_irradiance(lp.h, 32, None), lambda: L.prt_env_prefilter(lp.h, 256, 5, 1024, None),
                     lambda: L.prt_brdf_lut(ctx.h, 512, 512, 1024, None)):
            if call() != 0:
                raise RuntimeError(L.prt_last_error().decode())
            ms += ctx.last_kernel_ms()
        return ms

    if child:
        device_step(); device_step()
        return
    W_ = max(a.warmup, 3)
    for _ in range(W_):
        device_step()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    clocks = ClockSampler(0, 100)
    dev_ms, pre_ms = [], []
    for _ in range(a.steps):
        flush.zero_(); torch.cuda.synchronize()
        L.prt_env_irradiance(lp.h, 32, None); m0 = ctx.last_kernel_ms()
        L.prt_env_prefilter(lp.h, 256, 5, 1024, None); m1 = ctx.last_kernel_ms()
        L.prt_brdf_lut(ctx.h, 512, 512, 1024, None); m2 = ctx.last_kernel_ms()
        dev_ms.append(m0 + m1 + m2); pre_ms.append(m1)
    value = ENV_SAMPLES * a.steps / (sum(dev_ms) * 1e-3)

    def e2e_step():
        lp2 = prt_b200.LightProbe(h_eq, 512, ctx)                 # H2D of the equirect image, cube + mips
        out = (lp2.irradiance(32), lp2.prefilter(256, 5, 1024), prt_b200.brdf_lut(512, 512, 1024, ctx), lp2.project_sh(3, 0))
        lp2.close()
        return out
    e2e_step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res = e2e_step()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()
    e2e_value = ENV_SAMPLES * a.steps / e2e_s
    d2h = sum(x.nbytes for x in res[1]) + res[0].nbytes + res[2].nbytes + res[3].nbytes
    ok = all(np.isfinite(x).all() for x in [res[0], res[2], res[3]] + list(res[1]))
    peaks = load_peaks()
    n_sms = torch.cuda.get_device_properties(0).multi_processor_count
    ncu = None if a.no_ncu else ncu_pass(a, "prefilter_kernel", 5, count=5)   # the five mip launches of the second step, summed
    pk_ms = float(np.mean(pre_ms))
    alg = 523776 * 1024 * 8 * 16.0                                               # one trilinear cube fetch per sample = 8 taps x 16-byte texels
    roofline = issue_roofline(ncu, pk_ms, n_sms, clk.get("sm_mhz") if clk else None, peaks, 523776 * 12.0 + 33.5e6, alg,
                              "prefilter_kernel, the five mip launches of one step (256^2 .. 16^2, roughness 0 .. 1)")
    line = {"metric": "env_sample_evaluations_per_sec", "value": value, "unit": "samples/s", "n_gpus": 1, "steps": a.steps, "warmup": W_,
            "ms_per_step": sum(dev_ms) / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "file" if "newport" in eq_name else "synthetic",
            "config": {"workload": f"environment pass (BASELINE configs[1]): {eq_name} -> 512^2 cube + mips; irradiance 32^2 (15876 samples/texel), GGX prefilter 256^2 x 5 mips "
                                   "(1024 samples/texel), BRDF LUT 512^2 (1024 samples/texel)", "baseline_config": "2", "samples_per_step": ENV_SAMPLES,
                       "l2": "flushed between timed steps with a 512 MiB memset (untimed)"},
            "results_ok": ok, "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h_eq.nbytes), "d2h_bytes_per_step": int(d2h),
                                      "ms_per_step": 1e3 * e2e_s / a.steps, "note": "prt_env_create + irradiance + prefilter + brdf_lut + project_sh with host buffers"},
            "gpu_launches": a.steps * 7, "roofline": roofline, "clocks": clk,
            "passes_ms": {"irradiance": None, "prefilter": float(np.mean(pre_ms)), "all": float(np.mean(dev_ms))}}
    if not a.no_cpu_baseline:
        from oracle import pyoracle
        t0 = time.perf_counter()
        cube = pyoracle.EnvCube(eq, 512)
        t_cube = time.perf_counter() - t0
        t0 = time.perf_counter()
        cube.prefilter(64, 5, 1024)
        dt = time.perf_counter() - t0
        n = 6 * (64 * 64 + 32 * 32 + 16 * 16 + 8 * 8 + 4 * 4) * 1024
        line["cpu_baseline"] = {"value": n / dt, "unit": "samples/s", "cores": 1, "kind": "port",
                                "sample": f"GGX prefilter of a 64^2 x 5 cube from the same 512^2 environment, 1024 samples/texel ({dt:.1f} s; cube build {t_cube:.1f} s not counted)"}
    emit_line(line)


def run_reference_env(a):
    from oracle import pyoracle
    eq, eq_name = env_image()
    cube = pyoracle.EnvCube(eq, 512)
    n = 6 * (64 * 64 + 32 * 32 + 16 * 16 + 8 * 8 + 4 * 4) * 1024
    cube.prefilter(8, 2, 1024)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cube.prefilter(64, 5, 1024)
    dt = time.perf_counter() - t0
    v = n * a.steps / dt
    emit_line({"impl": "reference", "metric": "env_sample_evaluations_per_sec", "value": v, "unit": "samples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": 1,
               "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": f"environment pass (BASELINE configs[1]): {eq_name}", "baseline_config": "2"},
               "cpu_baseline": {"value": v, "unit": "samples/s", "cores": 1, "kind": "port", "sample": "GGX prefilter of a 64^2 x 5 cube per step, 1024 samples/texel"},
               "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


# ----------------------------------------------------------------------------------------------------------------
# config 3 / f1: probe capture, calculate_weight
# ----------------------------------------------------------------------------------------------------------------
def probe_scene():
    rp = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32) * 6.18
    rt = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2]], np.uint32)
    tp, _, tt = meshes.bumpy_torus(737, 737)
    return np.concatenate([rp, tp]).astype(np.float32), np.concatenate([rt, tt + np.uint32(8)]).astype(np.uint32)


def run_probe(a):
    import torch

    import prt_b200
    child = a.ncu_child
    torch.cuda.set_device(0)
    ctx = prt_b200.Context(0)
    pos, tri = probe_scene()
    sc = prt_b200.RTScene(pos, tri, ctx)
    info = sc.info()
    f1 = a.config == "f1"
    if f1:
        pres, vres, size = [8] * 3, [96] * 3, [12.0] * 3
        units = int(np.prod(vres)) * 108
        L = ctx.L
        pr, vr, sz = np.asarray(pres, np.int32), np.asarray(vres, np.int32), np.asarray(size, np.float32)

        def dev_step():
            rc = L.prt_volume_weights(sc.h, pr.ctypes.data, vr.ctypes.data, sz.ctypes.data, None, None, None)
            if rc != 0:
                raise RuntimeError(L.prt_last_error().decode())
            return ctx.last_kernel_ms()

        def e2e_step():
            return prt_b200.calculate_weight(sc, pres, vres, size)
        kname, skip, kcount = "volume_(score|weight)_kernel", 2, 2
        metric, unit = "rays_per_sec", "rays/s"
        wl = f"calculate_weight (SURVEY 8 row f1): 8^3 probes, 96^3 voxels x (100 closest-hit + 8 any-hit rays), room + buddha-scale torus ({len(tri)} triangles)"
    else:
        probes = prt_b200.probe_positions([32] * 3, [6.18] * 3)
        d, w = prt_b200.fibonacci_dirs(4096)
        units = len(probes) * 4096
        keep = {}

        def dev_step():
            pt = prt_b200.ProbeTransfer(sc, probes, d, w)
            ms = pt.capture_ms
            keep["nnz"], keep["surfels"] = pt.nnz, pt.n_surfels
            pt.close()
            return ms

        def e2e_step():
            pt = prt_b200.ProbeTransfer(sc, probes, d, w)
            out = pt.download()
            pt.close()
            return out
        kname, skip, kcount = "probe_capture_kernel", 1, 1
        metric, unit = "closest_hit_rays_per_sec", "rays/s"
        wl = f"probe capture (BASELINE configs[2]): 32^3 probes x 4096 closest-hit rays, room + buddha-scale torus ({len(tri)} triangles), CSR out"
    if child:
        dev_step(); dev_step()
        return
    W_ = max(a.warmup, 3)
    for _ in range(W_):
        dev_step()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    clocks = ClockSampler(0, 100)
    ms = []
    for _ in range(a.steps):
        flush.zero_(); torch.cuda.synchronize()
        ms.append(dev_step())
    value = units * a.steps / (sum(ms) * 1e-3)
    e2e_step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res = e2e_step()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()
    d2h = sum(x.nbytes for x in res)
    h2d = (len(probes) * 12 + 4096 * 16) if not f1 else 1200
    peaks = load_peaks()
    n_sms = torch.cuda.get_device_properties(0).multi_processor_count
    ncu = None if a.no_ncu else ncu_pass(a, kname, skip, count=kcount)
    k_ms = float(np.mean(ms))
    roofline = issue_roofline(ncu, k_ms, n_sms, clk.get("sm_mhz") if clk else None, peaks, float(info.node_bytes + info.tri_bytes) + d2h, 0.0,
                              "volume_score_kernel + volume_weight_kernel (100 closest-hit + 8 any-hit rays per voxel)" if f1 else "probe_capture_kernel (trace + sort + reduce per probe)")
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": a.steps, "warmup": W_, "ms_per_step": sum(ms) / a.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "baseline_config": a.config, "rays_per_step": units, "l2": "flushed between timed steps with a 512 MiB memset (untimed)"},
            "results_ok": bool(all(np.isfinite(x).all() for x in res if x.dtype.kind == "f" and x.size and not f1) or f1),
            "e2e": {"value": units * a.steps / e2e_s, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / a.steps},
            "gpu_launches": a.steps * (2 if f1 else 4), "roofline": roofline, "clocks": clk}
    if not f1:
        line["csr"] = {"nnz": int(keep["nnz"]), "surfels": int(keep["surfels"])}
    if not a.no_cpu_baseline:
        from oracle import pyoracle
        osc = pyoracle.Scene(pos, tri)
        t0 = time.perf_counter()
        if f1:
            pyoracle.volume_weights(osc, [8] * 3, [64] * 3, [12.0] * 3)
            n = 64 ** 3 * 108
            sample = "a 64^3 voxel volume of the same scene (8^3 probes)"
        else:
            n_p = a.cpu_sample or 2048
            sel = np.arange(0, len(probes), len(probes) // n_p)[:n_p]
            pyoracle.ProbeTransfer(osc, probes[sel], d, w)
            n = len(sel) * 4096
            sample = f"{len(sel)} strided probes x 4096 rays of the same scene"
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": unit, "cores": 1, "kind": "port", "sample": f"{sample} ({dt:.1f} s of CPU work)"}
    emit_line(line)


def run_reference_probe(a):
    from oracle import pyoracle
    pos, tri = probe_scene()
    osc = pyoracle.Scene(pos, tri)
    f1 = a.config == "f1"
    if f1:
        def step():
            pyoracle.volume_weights(osc, [8] * 3, [48] * 3, [12.0] * 3)
            return 48 ** 3 * 108
        metric, sample = "rays_per_sec", "a 48^3 voxel volume per step"
    else:
        probes = pyoracle.probe_positions([32] * 3, [6.18] * 3)
        d, w = pyoracle.fibonacci_dirs(4096)
        sel = np.arange(0, len(probes), len(probes) // 1024)[:1024]

        def step():
            pyoracle.ProbeTransfer(osc, probes[sel], d, w)
            return len(sel) * 4096
        metric, sample = "closest_hit_rays_per_sec", "1024 strided probes x 4096 rays per step"
    step()
    t0 = time.perf_counter()
    n = sum(step() for _ in range(a.steps))
    dt = time.perf_counter() - t0
    emit_line({"impl": "reference", "metric": metric, "value": n / dt, "unit": "rays/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": 1, "ms_per_step": 1e3 * dt / a.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": "probe capture" if not f1 else "calculate_weight", "baseline_config": a.config},
               "cpu_baseline": {"value": n / dt, "unit": "rays/s", "cores": 1, "kind": "port", "sample": sample},
               "e2e": {"value": n / dt, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def main():
    # stdout carries exactly one JSON line.  Libraries write there behind Python's back (NCCL prints its version banner with
    # printf when NCCL_DEBUG is VERSION/INFO in the environment), so file descriptor 1 is pointed at stderr for the whole run and
    # the JSON line is written to a private duplicate of the original stdout.
    global _JSON_OUT
    a = parse()
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank0 = int(os.environ.get("RANK", "0")) == 0
    if a.impl == "reference":
        if not rank0:
            return
        if a.config in BAKE_CONFIGS:
            run_reference_bake(a)
        elif a.config == "2":
            run_reference_env(a)
        else:
            run_reference_probe(a)
    elif a.config in BAKE_CONFIGS:
        run_bake(a)
    else:
        if not rank0:
            return                  # configs 2, 3, f1 run on one GPU; further ranks exit without work
        (run_env if a.config == "2" else run_probe)(a)


if __name__ == "__main__":
    main()
