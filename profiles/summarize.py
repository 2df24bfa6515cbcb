#!/usr/bin/env python
"""Summarise gpurun_out/*.ncu-rep + launches_*.csv into profiles/<tag>_*.{csv,md} (text, committed).
usage: python profiles/summarize.py r01"""
import csv, io, os, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
out = [f"# ncu summaries, tag {tag} (profiles/run_ncu.sh; B200, --clock-control none)\n"]
for name in ("trace", "rerender", "rrwalk", "conv"):
    rep = os.path.join(G, f"{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        out.append(f"\n## {name}: {d.get('Kernel Name')}\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                out.append(f"| {k} | {d[k]} | {u[k]} |")
        stalls = sorted(((float(d[h].replace(',', '')), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and d[h]), reverse=True)[:6]
        if stalls:
            out.append("\nTop warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in stalls))
lc = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lc):
    txt = "".join(l for l in open(lc) if not l.startswith("=="))
    agg = defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(io.StringIO(txt)):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        agg[r["Kernel Name"]][0] += 1; agg[r["Kernel Name"]][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append(f"\n## launch list ({lc.split('/')[-1]}): gpu__time_duration.sum per kernel, cold-cache and serialised\n\n| launches | total ms | share | kernel |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% | `{k[:100]}` |")
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as fh:
        fh.write(txt)
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
# per-launch DRAM traffic of the captured kernels, read by bench.py for roofline.traffic
import json, re
traffic, figures = {}, {}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
for name in ("trace", "rerender", "rrwalk", "conv"):
    rep = os.path.join(G, f"{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
    val = lambda k: float(d[k].replace(",", "")) * scale.get(u[k], 1)
    tot = sum(val(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in d)
    traffic[name] = {"kernel": d.get("Kernel Name"), "dram_bytes_per_launch": tot, "tag": tag}
    if name == "trace":
        # the figures north_star asks for: warp-execution efficiency and achieved L2 / HBM GB/s of the traversal
        t = val("gpu__time_duration.sum")
        figures["wave_kernel"] = {
            "source": f"ncu --set full capture {tag} (profiles/{tag}_ncu_summary.md), C2, one launch, cold L2",
            "warp_exec_efficiency": val("smsp__thread_inst_executed_per_inst_executed.ratio") / 32.0,
            "lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "l2_gbs": val("lts__t_sectors.sum") * 32 / t / 1e9 if "lts__t_sectors.sum" in d else None,
            "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
            "hbm_gbs": tot / t / 1e9,
            "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"),
            "kernel_ms": t * 1e3}
if traffic:
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
if figures:
    json.dump(figures, open(os.path.join(ROOT, "profiles", "ncu_figures.json"), "w"), indent=1)
print("\n".join(out))
