#!/usr/bin/env python
"""SURVEY.md 8(f) N4 on one B200: the on-rails traffic pass as a device-resident dirty producer
(python tools/bench_traffic.py > gpurun_out/traffic.json).

n vehicles on the procedural lanes of a 128 x 128 sector city. Timed, median of 10 (wall clock around the call + a
synchronise): scgpuTrafficAdvance (nothing crosses PCIe), the host path it replaces on the device side (scgpuSetLocal of
the n moved TRS from a host buffer), the plain-C oracle on one host core, and — when oracle/_ref travelled — the
reference's own TrafficAISystem over the same agents."""
import ctypes as C
import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import scgpu  # noqa: E402
from scgpu import scenes  # noqa: E402
import oracle_bind  # noqa: E402
from oracle_bind import LANE_KEYS, port_traffic_step  # noqa: E402


def med(fn, reps=10):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    g = scenes.lane_grid(128, 128)
    agents, trs = scenes.traffic_agents(g, n, seed=1, hostile=False)
    e = np.arange(n, dtype=np.uint32)
    s = scgpu.Scene(n, max_views=1, max_entity_index=n)
    s.spawn(e, trs)
    s.traffic_set_lanes(*[g[k] for k in LANE_KEYS], default_speed=float(g["default_speed"]))
    s.traffic_set_agents(e, agents["lane"], agents["s"], agents["speed"], agents["look"])
    s.set_views(scenes.standard_views(1, center=(4096.0, 6.0, 4096.0)))
    s.update()
    s.synchronize()

    def gpu_step():
        s.traffic_advance(1 / 60, want_moved=False)
        s.synchronize()
    gpu_step()
    out = {"agents": n, "lane_segments": int(len(g["seg_len"])), "traffic_advance_ms": med(gpu_step),
           "algorithmic_bytes_per_agent": 16 + 4 + 16 + 16 + 8 + 4 + 16 + 4,
           "what": "k_traffic_advance: agent record 16+4 B read and written back, localPos 16 B read + 16 B written, yaw 8 B, "
                   "dirty stamp 4 B r/w; lane records come from L2"}
    moved = s.traffic_advance(1 / 60)
    out["moved"] = int(moved)

    def upd():
        s.update()
        s.synchronize()
    s.traffic_advance(1 / 60, want_moved=False)
    upd()
    ts = []
    for _ in range(5):
        s.traffic_advance(1 / 60, want_moved=False)
        s.synchronize()
        t0 = time.perf_counter()
        upd()
        ts.append(time.perf_counter() - t0)
    out["update_after_traffic_ms"] = statistics.median(ts) * 1e3
    host_trs = s.read_local(e)

    def host_path():
        s.set_local(e, host_trs)
        s.synchronize()
    host_path()
    out["host_set_local_ms"] = med(host_path, 5)
    out["host_set_local_h2d_bytes"] = int(n * 40)
    s.close()

    pa = {k: v.copy() for k, v in agents.items()}
    pt = trs.copy()
    out["oracle_port_1core_ms"] = med(lambda: port_traffic_step(g, pa, pt, 1 / 60), 3)
    if oracle_bind.ref_available() and n <= 2_000_000:
        lanes = oracle_bind.RefLanes(3.5, 12.0)
        for sx in range(128):
            for sz in range(128):
                lanes.build_sector(sx, sz, 64.0)
        R = oracle_bind.RefScene(1)
        ents = R.create_entities(n + 1)
        R.spawn(ents[:1], np.array([[0, 0, 0, 0, 0, 0, 1, 1, 1]], np.float32))
        ag = np.ascontiguousarray(ents[1:])
        R.spawn(ag, trs)
        f = lambda a: a.ctypes.data_as(C.c_void_p)
        R.L.screfTrafficSetPlayer(R.w, int(ents[0]))
        R.L.screfTrafficAddAgents(R.w, n, f(ag), f(agents["lane"]), f(agents["s"]), f(agents["speed"]), f(agents["look"]))
        step = lambda: R.L.screfRunTrafficAI(R.w, lanes.h, 1 / 60, 0, 0.0, 0.0)
        step()
        out["reference_TrafficAISystem_ms"] = med(step, 3)
        R.close()
        lanes.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
