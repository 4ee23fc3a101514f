"""Summarise an .ncu-rep (one k_step launch, `ncu --set full`) + a launch list CSV into profiles/*.md|json.

    python tools/summarize_ncu.py gpurun_out/prof_step_v3.ncu-rep gpurun_out/launches_v3.csv profiles/r01_step

Run in the build container (ncu is installed, no GPU needed to read reports)."""
import csv
import io
import json
import subprocess
import sys

rep, launches, out = sys.argv[1], sys.argv[2], sys.argv[3]
desc = sys.argv[4] if len(sys.argv) > 4 else "`ncu --set full --clock-control none --import-source on -k regex:k_step` on `python bench.py` (10x20, queue 7)"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
def unit_scale(v, u):
    v = float(v)
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(u, 1)
summary = {k: {"value": m[k][0], "unit": m[k][1]} for k in keys if k in m}
stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
          for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and "not_issued" not in h and v not in ("", "n/a")}
rd = unit_scale(*m["dram__bytes_read.sum"]); wr = unit_scale(*m["dram__bytes_write.sum"]); dur = unit_scale(*m["gpu__time_duration.sum"])
summary["derived"] = {"dram_bytes_per_launch": rd + wr, "duration_s": dur, "dram_GBps_under_ncu": (rd + wr) / dur / 1e9}
summary["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
summary["kernel"] = m.get("Kernel Name", ("", ""))[0]
# launch list ("-" = none)
tot = {}
if launches != "-":
    ls = list(csv.reader(open(launches)))
    hi = next(i for i, r in enumerate(ls) if r and r[0] == "ID")
    kn, mv = ls[hi].index("Kernel Name"), ls[hi].index("Metric Value")
    for r in ls[hi + 1:]:
        if len(r) > mv:
            name = r[kn].split("(")[0][-70:]
            t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += float(r[mv].replace(",", ""))
allns = sum(t[1] for t in tot.values()) or 1.0
summary["launch_list"] = [{"kernel": k, "launches": t[0], "total_us": t[1] / 1e3, "share": t[1] / allns} for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1])[:10]]
json.dump(summary, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write(f"# ncu summary: {rep}\n\nKernel: `{summary['kernel'][:160]}`\n\nSource: {desc}.\n"
            "Numbers under ncu are serialised / cold-cache: use shares and byte counts, not absolute times.\n\n| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in m:
            f.write(f"| {k} | {m[k][0]} | {m[k][1]} |\n")
    f.write(f"| dram bytes per launch (read+write) | {rd + wr:.4g} | byte |\n| dram GB/s under ncu | {(rd + wr) / dur / 1e9:.1f} | GB/s |\n")
    f.write("\n## stall reasons (warps per issue-active cycle)\n\n" + "\n".join(f"- {k}: {v:.2f}" for k, v in summary["stalls_per_issue"].items()))
    if tot:
        f.write(f"\n\n## launch list ({launches}, `--metrics gpu__time_duration.sum`)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for r in summary["launch_list"]:
        f.write(f"| `{r['kernel']}` | {r['launches']} | {r['total_us']:.1f} | {100 * r['share']:.1f}% |\n")
print(open(out + ".md").read())
