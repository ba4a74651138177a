"""The tracked bench lines (profiles/r02_bench_n*.json, written by bench.py on the B200) carry what the measurement
contract asks for, and their derived figures follow from their own inputs. CPU only: nothing is measured here."""
import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parent.parent / "profiles"


def _line(name):
    return json.loads((PROFILES / name).read_text().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r02_bench_n1.json", "r02_bench_n2.json", "r02_bench_n4.json"])
def test_bench_line_is_consistent(name):
    d = _line(name)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in d, key
    assert d["unit"] == "instances/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f32"
    assert d["warmup"] >= 3 and "workload" in d["config"] and "l2" in d["config"]
    n = d["config"]["instances_per_gpu"] * d["n_gpus"]
    assert d["value"] == pytest.approx(n / (d["ms_per_step"] * 1e-3), rel=1e-9)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert r["algorithmic_bytes_per_launch"] == 132 * d["config"]["instances_per_gpu"]  # dirty instance
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_launch"] / (r["kernel_ms_avg"] * 1e-3) / 1e9, rel=1e-9)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert 0.5 < r["kernel_share_of_step"] < 1.0
    # DRAM traffic of the dominant kernel (one ncu --set full capture) within 2 % of the algorithmic bytes: no re-reads
    assert r["traffic"] == pytest.approx(r["algorithmic_bytes_per_launch"], rel=0.02)
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] >= 4 * d["steps"]  # k_update_win, k_update_win_slow, k_compact, k_resolve_lists per step
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    else:
        g = d["gather"]
        assert g["gather_checked"] is True
        assert g["gathered_visible_per_view"] != g["rank0_local_visible_per_view"]  # the lists come from more than one GPU


def test_traffic_file_matches_the_capture_summary():
    t = json.loads((PROFILES / "traffic.json").read_text())
    rows = dict(line.split(",")[0::2] for line in (PROFILES / "r02_k_update_win_ncu_raw.csv").read_text().splitlines()
                if line.startswith("dram__bytes_"))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    units = dict(line.split(",")[0:2] for line in (PROFILES / "r02_k_update_win_ncu_raw.csv").read_text().splitlines()
                 if line.startswith("dram__bytes_"))
    total = sum(float(rows[k]) * scale[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    assert t["dram_bytes_per_launch"] == pytest.approx(total, rel=1e-6)
    assert t["algorithmic_bytes_per_launch"] == 132 * 16773120
