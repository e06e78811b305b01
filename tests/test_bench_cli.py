"""CPU tests of bench.py's host logic: configuration table, reference arm on a bounded sample (the JSON contract), and the parser
of the ncu child pass (a canned --csv log)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_config_table_matches_baseline_json():
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    a = bench.parse([])
    assert (a.config, a.nu, a.nv, a.order, a.samples_u * a.samples_v, a.gpus) == ("1", 737, 737, 3, 1024, 1)      # configs[0]: order 3, 1024 samples
    a = bench.parse(["--config", "4"])
    assert (a.order, a.samples_u * a.samples_v, a.nu * a.nv) == (4, 4096, 1448 * 1448) and bench.BAKE_CONFIGS["4"]["bounces"] == 3
    a = bench.parse(["--config", "5"])
    assert (a.order, a.samples_u * a.samples_v, a.nu * a.nv) == (5, 8192, 20_000_000)
    assert bench.ENV_SAMPLES == 523776 * 1024 + 262144 * 1024 + 6144 * 15876


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--nu", "60", "--nv", "40",
                        "--cpu-sample", "1800"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "shadow_rays_per_sec" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_ncu_csv_parser(tmp_path, monkeypatch):
    """ncu_pass sums `.sum` metrics over the captured launches, averages percentages and converts units."""
    import bench
    log = '''==PROF== Connected to process 1
"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
"0","1","python","h","k<3>(A)","1","7","(128, 1, 1)","(1036, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Mbyte","250.5"
"0","1","python","h","k<3>(A)","1","7","(128, 1, 1)","(1036, 1, 1)","0","10.0","Command line profiler metrics","smsp__inst_executed.sum","inst","1,000,000,000"
"0","1","python","h","k<3>(A)","1","7","(128, 1, 1)","(1036, 1, 1)","0","10.0","Command line profiler metrics","smsp__issue_active.avg.pct_of_peak_sustained_active","%","70"
"1","1","python","h","k<3>(A)","1","7","(128, 1, 1)","(1036, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Kbyte","500"
"1","1","python","h","k<3>(A)","1","7","(128, 1, 1)","(1036, 1, 1)","0","10.0","Command line profiler metrics","smsp__inst_executed.sum","inst","3,000,000,000"
"1","1","python","h","k<3>(A)","1","7","(128, 1, 1)","(1036, 1, 1)","0","10.0","Command line profiler metrics","smsp__issue_active.avg.pct_of_peak_sustained_active","%","80"
'''

    class R:
        returncode, stderr = 0, ""

    def fake_run(cmd, **kw):
        open(cmd[cmd.index("--log-file") + 1], "w").write(log)
        return R()
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    monkeypatch.setattr(bench.shutil, "which", lambda _: sys.executable)
    out = bench.ncu_pass(bench.parse([]), "k", 1, count=2)
    assert "error" not in out, out
    assert out["smsp__inst_executed.sum"] == 4e9 and abs(out["dram__bytes_read.sum"] - 251.0e6) < 1
    assert abs(out["smsp__issue_active.avg.pct_of_peak_sustained_active"] - 75.0) < 1e-9
    r = bench.issue_roofline(out, 2.0, 148, 1965.0, {"hbm_gbs": 6466.5}, 1e6, 2e6, "k")
    assert r["bound"] == "issue" and abs(r["achieved"] - 4e9 / 2e-3 / 1e9) < 1e-6 and abs(r["peak"] - 148 * 4 * 1.965) < 1e-9 and r["frac"] == r["achieved"] / r["peak"]
    assert r["traffic"] == out["dram__bytes_read.sum"]
