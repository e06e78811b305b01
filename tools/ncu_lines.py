#!/usr/bin/env python
"""tools/ncu_lines.py REPORT.ncu-rep [N [KERNEL_REGEX]] -- per-source-line instruction / stall-sample shares and key raw metrics."""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kf = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, *kf, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    if k in hdr: print(f"{k:75s} {vals[hdr.index(k)]} {rows[1][hdr.index(k)]}")
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
        try:
            if float(vals[i]) >= 2.0: print(f"{h:75s} {vals[i]}")
        except ValueError: pass
src = subprocess.run(["ncu", "-i", rep, *kf, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; lines = []; tot = tots = 0
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iE = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iS = hdr.index('# Samples'); continue
    if cur and hdr and r[0].isdigit():
        try: n, t, s = int(r[iE]), int(r[iT]), int(r[iS])
        except ValueError: continue
        if n == 0: continue
        tot += n; tots += s
        lines.append((n, t, s, cur.split('/')[-1], int(r[0]), r[1].strip()[:80]))
print("total warp inst", tot, "samples", tots)
pf = collections.Counter(); pfs = collections.Counter(); pft = collections.Counter()
for n, t, s, f, l, c in lines: pf[f] += n; pfs[f] += s; pft[f] += t
for f, n in pf.most_common(): print(f"{f:28s} {100*n/tot:5.1f}% inst {100*pfs[f]/tots:5.1f}% samples  thr/inst {pft[f]/n:5.1f}")
for n, t, s, f, l, c in sorted(lines, reverse=True)[:topn]:
    print(f"{100*n/tot:5.1f}% {100*s/tots:5.1f}%s thr/inst {t/n:5.1f}  {f}:{l}  {c}")
