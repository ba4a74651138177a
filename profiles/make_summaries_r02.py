#!/usr/bin/env python
"""Round 2: turns the ncu captures a gpurun call left in gpurun_out/ into the small, tracked summaries under profiles/.

  python profiles/make_summaries_r02.py     (reads gpurun_out/r02_prof.ncu-rep, r02_launches.csv, r02_bench_n*.json)

Nothing here is measured by this script: it only reformats `ncu -i ... --page raw --csv` output. The captures came from
  ncu --set full --clock-control none --import-source on -k regex:"^k_update_win$|k_cull_only|k_compact|k_resolve_lists" \
      -s 12 -c 12 -o gpurun_out/r02_prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-churn
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --churn-frames 4"""
import collections
import csv
import io
import json
import re
import shutil
import subprocess
from pathlib import Path

from make_summaries import KEEP

ROOT = Path(__file__).resolve().parent.parent
OUT, SRC = ROOT / "profiles", ROOT / "gpurun_out"
N, VIEWS = 16773120, 5


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def val(unit, v):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    return float(v) * scale


def main():
    hdr, units, rows = raw_rows(SRC / "r02_prof.ncu-rep")
    ik = hdr.index("Kernel Name")
    seen = collections.Counter()
    picked = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
        seen[name] += 1
        picked.setdefault((name, seen[name]), dict(zip(hdr, zip(units, r))))
    # which launch of each kernel: the all-dirty leg comes first (k_update_win #1), the clean leg's k_cull_only #1,
    # the compaction / resolve kernels of the all-dirty leg
    wanted = {"k_update_win<5>": ("k_update_win", "all dirty: every instance re-transformed, culled against 5 views"),
              "k_cull_only<5>": ("k_cull_only", "nothing dirty: stored matrices, cull only"),
              "k_compact<5>": ("k_compact", "bitmaps -> rank-ordered lists, one CTA per 32 Ki-rank chunk"),
              "k_resolve_lists": ("k_resolve_lists", "rank -> slot -> entity handle")}
    traffic = None
    for (name, k), m in picked.items():
        if k != 1 or name not in wanted:
            continue
        short, what = wanted[name]
        lines = [f"# ncu --set full --clock-control none --import-source on, one launch; {name}, {N} instances in depth-4 groups, {VIEWS} views; {what}",
                 "metric,unit,value"]
        for key in KEEP:
            if key in m:
                lines.append(f"{key},{m[key][0]},{m[key][1]}")
        for key in sorted(m):
            if key.startswith("smsp__average_warps_issue_stalled") and key.endswith("per_issue_active.ratio"):
                lines.append(f"{key},{m[key][0]},{m[key][1]}")
        (OUT / f"r02_{short}_ncu_raw.csv").write_text("\n".join(lines) + "\n")
        if short == "k_update_win":
            rd, wr = val(*m["dram__bytes_read.sum"]), val(*m["dram__bytes_write.sum"])
            traffic = {"kernel": "k_update_win<5 views>", "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "algorithmic_bytes_per_launch": 132 * N,
                       "source": f"profiles/r02_k_update_win_ncu_raw.csv (ncu --set full, one launch, {N} instances x {VIEWS} views, all dirty)"}
    if traffic:
        (OUT / "traffic.json").write_text(json.dumps(traffic, indent=1))
    # launch list: per launch and aggregated per kernel
    t = open(SRC / "r02_launches.csv", errors="replace").read()
    t = t[t.index('"ID"'):]
    by = {}
    for r in csv.DictReader(io.StringIO(t)):
        e = by.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), "grid": r["Grid Size"], "block": r["Block Size"]})
        e[r["Metric Name"]] = float(r["Metric Value"])
    out = ["# ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400,",
           "# python bench.py --steps 2 --warmup 1 --no-cpu-baseline --churn-frames 4  (legs in order: all dirty, gather check (N>1), e2e x3, clean, 30 % dirty, churn)",
           "# cold-cache serialised per-launch device times; compare SHARES not absolutes", "index,kernel,grid,block,us,warp_instructions,dram_bytes"]
    agg = collections.OrderedDict()
    for i in sorted(by):
        e = by[i]
        us, inst = e.get("gpu__time_duration.sum", 0.0) / 1e3, e.get("smsp__inst_executed.sum", 0.0)
        dram = e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
        out.append(f'{i},{e["name"]},"{e["grid"]}","{e["block"]}",{us:.3f},{inst:.0f},{dram:.0f}')
        a = agg.setdefault(e["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += inst; a[3] += dram
    (OUT / "r02_launches.csv").write_text("\n".join(out) + "\n")
    lines = ["# per-kernel averages over profiles/r02_launches.csv", "kernel,launches,avg_us,avg_warp_instructions,avg_dram_bytes"]
    for k, a in agg.items():
        lines.append(f"{k},{a[0]},{a[1] / a[0]:.2f},{a[2] / a[0]:.0f},{a[3] / a[0]:.0f}")
    (OUT / "r02_launches_by_kernel.csv").write_text("\n".join(lines) + "\n")
    for f in SRC.glob("r02_bench_n*.json"):
        shutil.copy(f, OUT / f.name)
    print("written", sorted(p.name for p in OUT.iterdir() if p.name.startswith("r02") or p.name == "traffic.json"))


if __name__ == "__main__":
    main()
