#!/usr/bin/env python
"""Turns the ncu captures a gpurun call left in gpurun_out/ into the small, tracked summaries under profiles/.

  python profiles/make_summaries.py            (reads gpurun_out/r01_k_update_{win,flat}.ncu-rep, launches_r01.csv)

Nothing here is measured by this script: it only reformats `ncu -i ... --page raw --csv` output."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
SRC = ROOT / "gpurun_out"
ROUND = "r01"
VER = "v6"  # bumps when the profiled kernel changed (v5: persistent queue; v6: conflict-free matrix planes, FFMA2 sums)

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


def summarise(rep, name, what):
    m = raw(rep)
    lines = [f"# ncu --set full --clock-control none --import-source on, one launch; {what}", "metric,unit,value"]
    for k in KEEP:
        if k in m:
            lines.append(f"{k},{m[k][0]},{m[k][1]}")
    for k in sorted(m):
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            lines.append(f"{k},{m[k][0]},{m[k][1]}")
    (OUT / name).write_text("\n".join(lines) + "\n")
    return m


def gb(m, key):
    unit, val = m[key]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(val) * scale


def main():
    n, views = 16773120, 5
    win = summarise(SRC / f"{ROUND}_k_update_win.ncu-rep", f"{ROUND}_k_update_win_{VER}_ncu_raw.csv",
                    f"k_update_win<5>, {n} instances in depth-4 groups, {views} views, all dirty (bench.py default workload)")
    if "--flat" in sys.argv:  # the flat kernel's capture is only refreshed when it was re-profiled
        summarise(SRC / f"{ROUND}_k_update_flat.ncu-rep", f"{ROUND}_k_update_flat_{VER}_ncu_raw.csv",
                  f"k_update_flat<5>, {n} flat instances, {views} views, all dirty (SCGPU_BENCH_WORKLOAD=flat)")
    rd, wr = gb(win, "dram__bytes_read.sum"), gb(win, "dram__bytes_write.sum")
    (OUT / "traffic.json").write_text(json.dumps({
        "kernel": "k_update_win<5 views>", "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "algorithmic_bytes_per_launch": 132 * n,
        "source": f"profiles/{ROUND}_k_update_win_{VER}_ncu_raw.csv (ncu --set full, one launch, {n} instances x {views} views, all dirty)",
    }, indent=1))
    # launch list
    rows = list(csv.reader(open(SRC / f"launches_{ROUND}.csv", errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 80, python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline",
           "# cold-cache serialised per-launch device times (us); compare SHARES not absolutes", "index,kernel,grid,block,us"]
    col = {c: i for i, c in enumerate(rows[hdr])}
    for r in rows[hdr + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        out.append(f'{r[col["ID"]]},{name},"{r[col["Grid Size"]]}","{r[col["Block Size"]]}",{float(r[col["Metric Value"]]) / 1e3:.3f}')
    (OUT / f"{ROUND}_launches.csv").write_text("\n".join(out) + "\n")
    print("written", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    sys.exit(main())
