"""SURVEY.md 8(f) N4, traffic on rails — CPU side of the parity chain:

  host libm (glibc 2.39)  ==  oracle/scoracle.c (sco_expf / sco_atanf / sco_atan2f)  ==  device math compiled for the
  host (tests/hostsim, scgpu_math.cuh);
  the reference's own TrafficAISystem + TrafficLaneGraph (oracle/_ref)  ==  sco_traffic_ai_on_rails  ==  the device
  routine traffic_agent_on_rails compiled for the host;
  scenes.lane_grid  ==  the reference's buildProceduralForSector.
The GPU kernel around the device routine is checked in tests/test_gpu_traffic.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_bind
from oracle_bind import LANE_KEYS, port_traffic_step
from scgpu import scenes

HS_DIR = Path(__file__).resolve().parent / "hostsim"
GOLDEN = Path(__file__).resolve().parent / "golden" / "traffic.npz"
f = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def hs():
    subprocess.run(["make", "-C", str(HS_DIR)], check=True, capture_output=True)
    L = C.CDLL(str(HS_DIR / "libhostsim.so"))
    L.hs_unary_sweep.restype = C.c_uint64
    L.hs_unary_sweep.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p]
    L.hs_atan2_pairs.restype = C.c_uint64
    L.hs_atan2_pairs.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hs_traffic_on_rails.restype = None
    L.hs_traffic_on_rails.argtypes = [C.c_uint32, C.c_uint32] + [C.c_void_p] * 8 + [C.c_float, C.c_uint32] + [C.c_void_p] * 7 + \
                                     [C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p]
    return L


@pytest.fixture(scope="module")
def libm():
    L = C.CDLL("libm.so.6")
    return L


def hs_traffic_step(hs, g, agents, trs9, dt, brake=None, skip=None, debug=None):
    n = len(agents["lane"])
    moved = np.zeros(n, np.uint8)
    hs.hs_traffic_on_rails(len(g["node_speed"]), len(g["seg_len"]), *[f(g[k]) for k in LANE_KEYS], float(g["default_speed"]), n,
                           f(agents["lane"]), f(agents["s"]), f(agents["speed"]), f(agents["look"]), f(trs9), f(brake), f(skip),
                           dt, 1 if debug else 0, debug[0] if debug else 0.0, debug[1] if debug else 0.0, f(moved))
    return moved


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def copy_agents(a):
    return {k: v.copy() for k, v in a.items()}


# ---------------------------------------------------------------------------------------------------------------
# libm
# ---------------------------------------------------------------------------------------------------------------

EXPF_FMA_ONLY = (0x4202422F, 0xC27C65D9)  # the two inputs where glibc's FMA ifunc variant of expf differs


def test_oracle_libm_matches_host_glibc(port, libm):
    """strided sweeps (4.2 M inputs each) + dense windows at the branch thresholds; the full 2^32 sweeps of expf and
    atanf and 4e8 atan2f pairs were run once in the build container: 0 mismatches against the generic variants."""
    n = (1 << 32) // 1021 + 1
    bad = port.sco_unary_sweep(0, 0, n, 1021, C.cast(libm.expf, C.c_void_p))
    assert bad == 0
    assert port.sco_unary_sweep(1, 0, n, 1021, C.cast(libm.atanf, C.c_void_p)) == 0
    for centre in (0x42B00000, 0x42B17218, 0xC2CFF1B4, 0x7F800000, 0x00800000, 0x3F800000):  # 88, ln(2^128), ln(2^-150)
        for c in (centre, centre | 0x80000000):
            assert port.sco_unary_sweep(0, (c - 20000) & 0xFFFFFFFF, 40000, 1, C.cast(libm.expf, C.c_void_p)) == 0
    for centre in (0x4C000000, 0x3EE00000, 0x31000000, 0x3F980000, 0x3F300000, 0x401C0000):
        for c in (centre, centre | 0x80000000):
            assert port.sco_unary_sweep(1, c - 20000, 40000, 1, C.cast(libm.atanf, C.c_void_p)) == 0
    assert port.sco_atan2_sweep(12345, 3_000_000, C.cast(libm.atan2f, C.c_void_p)) == 0


def test_device_libm_on_host_matches_oracle(hs, port):
    n = (1 << 32) // 1021 + 1
    assert hs.hs_unary_sweep(0, 0, n, 1021, C.cast(port.sco_expf, C.c_void_p)) == 0
    assert hs.hs_unary_sweep(1, 0, n, 1021, C.cast(port.sco_atanf, C.c_void_p)) == 0
    for x in EXPF_FMA_ONLY:
        assert hs.hs_unary_sweep(0, x, 1, 1, C.cast(port.sco_expf, C.c_void_p)) == 0
    rng = np.random.default_rng(3)
    m = 2_000_000
    y = rng.integers(0, 1 << 32, m, dtype=np.uint64).astype(np.uint32).view(np.float32).copy()
    x = rng.integers(0, 1 << 32, m, dtype=np.uint64).astype(np.uint32).view(np.float32).copy()
    ang = (rng.random(m // 2) * 6.3).astype(np.float32)
    y[: m // 2], x[: m // 2] = np.sin(ang), np.cos(ang)            # directions, as lane graphs hold them
    x[m // 2: m // 2 + 1000] = np.float32(1.0)                       # the x == 1 shortcut
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, 1e38, 1e-38], np.float32)
    yy, xx = np.meshgrid(sp, sp)
    y[-100:], x[-100:] = yy.ravel(), xx.ravel()
    assert hs.hs_atan2_pairs(m, f(y), f(x), C.cast(port.sco_atan2f, C.c_void_p)) == 0


# ---------------------------------------------------------------------------------------------------------------
# lane graph + the AI step
# ---------------------------------------------------------------------------------------------------------------

def ref_graph_from_arrays(g):
    """feeds an arbitrary array graph to the reference's TrafficLaneGraph through addNode / addSegment; returns the
    RefLanes or None when the reference would build a different graph (addNode merges nodes with equal quantised keys)"""
    lanes = oracle_bind.RefLanes(3.5, float(g["default_speed"]))
    for i in range(len(g["node_speed"])):
        # a unique direction per node keeps addNode's quantised (pos, dir) key unique
        d = np.array([((i % 1000) - 500) / 1000.0, ((i // 1000) % 1000 - 500) / 1000.0, 0.25], np.float32)
        if lanes.add_node(g["node_pos"][i], d, float(g["node_speed"][i])) != i:
            lanes.close()
            return None
    for s in range(len(g["seg_len"])):
        a, b = g["seg_nodes"][s]
        lanes.add_segment(int(a), int(b), g["seg_dir"][s])
    return lanes


def run_chain(hs, g, agents, trs, dts, ref_frames=None, brake=None, skip=None, debug=None):
    """port and hostsim stepped side by side over dts; optionally compared with the reference's frames"""
    pa, pt = copy_agents(agents), trs.copy()
    ha, ht = copy_agents(agents), trs.copy()
    moved_total = 0
    for k, dt in enumerate(dts):
        pm = port_traffic_step(g, pa, pt, dt, brake, skip, debug)
        hm = hs_traffic_step(hs, g, ha, ht, dt, brake, skip, debug)
        assert np.array_equal(pm, hm), f"frame {k}: moved masks differ"
        assert np.array_equal(pa["lane"], ha["lane"]), f"frame {k}"
        for key in ("s", "speed", "look"):
            assert bits_equal(pa[key], ha[key]), f"frame {k}: {key}"
        assert bits_equal(pt, ht), f"frame {k}: local TRS"
        if ref_frames is not None:
            lane, s, v, look, t, dirty = ref_frames[k]
            assert np.array_equal(lane, pa["lane"]), f"frame {k}: lane vs reference"
            assert bits_equal(s, pa["s"]) and bits_equal(v, pa["speed"]) and bits_equal(look, pa["look"]), f"frame {k}"
            assert bits_equal(t, pt), f"frame {k}: local TRS vs reference"
            assert np.array_equal(dirty, pm), f"frame {k}: dirty flags vs reference"
        moved_total += int(pm.sum())
    return moved_total, pa, pt


def test_lane_grid_equals_reference_procedural_lanes(ref):
    lanes = oracle_bind.RefLanes(3.5, 12.0)
    for sx in range(-2, 3):
        for sz in range(-1, 3):
            lanes.build_sector(sx, sz, 64.0)
    r = lanes.export()
    lanes.close()
    g = scenes.lane_grid(5, 4, x0=-2, z0=-1)
    for k in LANE_KEYS:
        assert r[k].shape == g[k].shape, k
        assert np.array_equal(r[k].view(np.uint8), g[k].view(np.uint8)), k
    assert r["default_speed"] == g["default_speed"]


@pytest.mark.parametrize("debug", [None, (9.0, 1.7)])
def test_on_rails_step_equals_reference_on_the_procedural_grid(ref, hs, debug):
    lanes = oracle_bind.RefLanes(3.5, 12.0)
    for sx in range(4):
        for sz in range(4):
            lanes.build_sector(sx, sz, 64.0)
    lanes.remove_sector(2, 1)  # an unloaded sector in the middle: inactive lanes
    g = lanes.export()
    agents, trs = scenes.traffic_agents(g, 600, seed=5)
    dts = [1 / 60, 1 / 60, 0.0, 1 / 30, 0.25, 1.5, 1 / 144] + [1 / 60] * 25 + [3.0] * 4
    frames = oracle_bind.ref_traffic_frames(lanes, agents, trs, dts, debug)
    lanes.close()
    moved, _, _ = run_chain(hs, g, agents, trs, dts, frames, debug=debug)
    assert moved > 600 * 10


def test_on_rails_step_equals_reference_on_random_graphs(ref, hs):
    for seed in range(4):
        g = scenes.lane_random(300, 700, seed=seed, hostile=False)
        lanes = ref_graph_from_arrays(g)
        assert lanes is not None
        # hostile lane state through the reference's own switches
        rng = np.random.default_rng(seed)
        for s in np.nonzero(rng.random(700) < 0.08)[0]:
            lanes.set_active(int(s), False)
        r = lanes.export()
        for k in ("node_pos", "node_speed", "seg_nodes", "conn_offset", "conn"):
            assert np.array_equal(r[k], g[k]), k
        agents, trs = scenes.traffic_agents(r, 800, seed=seed + 40)
        dts = [1 / 60] * 20 + [0.5, 2.0, 10.0, 1 / 60]
        frames = oracle_bind.ref_traffic_frames(lanes, agents, trs, dts)
        lanes.close()
        moved, _, _ = run_chain(hs, r, agents, trs, dts, frames)
        assert moved > 800 * 5


def test_device_routine_equals_oracle_on_hostile_graphs(hs):
    """no reference needed: inactive / degenerate segments, dangling connection ids, zero and negative speed limits,
    agents without or with invalid lanes, obstacle brakes, skipped agents, NaN and huge dt"""
    for seed in range(6):
        g = scenes.lane_random(200, 500, seed=100 + seed, hostile=True)
        agents, trs = scenes.traffic_agents(g, 700, seed=seed)
        rng = np.random.default_rng(seed)
        brake = rng.random(700).astype(np.float32)
        brake[rng.random(700) < 0.5] = 0
        skip = (rng.random(700) < 0.1).astype(np.uint8)
        dts = [1 / 60] * 10 + [0.0, 5.0, 1e-8, 40.0, float("nan"), 1 / 60, 1e30, 1 / 60]
        run_chain(hs, g, agents, trs, dts, None, brake, skip, debug=(20.0, 0.5) if seed & 1 else None)


def test_golden_traffic_frames(hs):
    """frames recorded from the reference's TrafficAISystem (tests/golden/make_golden.py) — holds when oracle/_ref
    is absent"""
    z = np.load(GOLDEN)
    g = {k: z["g_" + k] for k in LANE_KEYS}
    g["default_speed"] = np.float32(z["g_default_speed"])
    agents = dict(lane=z["a_lane"].copy(), s=z["a_s"].copy(), speed=z["a_speed"].copy(), look=z["a_look"].copy())
    dts = [float(x) for x in z["dts"]]
    frames = [(z["f_lane"][k], z["f_s"][k], z["f_speed"][k], z["f_look"][k], z["f_trs"][k], z["f_dirty"][k]) for k in range(len(dts))]
    run_chain(hs, g, agents, z["trs"].copy(), dts, frames)
    # known answers of the three libm routines, taken from the reference build (std::exp / std::atan / std::atan2)
    port = oracle_bind.port_lib()
    for name, fn in (("expf", port.sco_expf), ("atanf", port.sco_atanf)):
        got = np.array([fn(float(x)) for x in z["kat_x"]], np.float32)
        assert bits_equal(got, z["kat_" + name]), name
    got = np.array([port.sco_atan2f(float(y), float(x)) for y, x in zip(z["kat_y2"], z["kat_x2"])], np.float32)
    assert bits_equal(got, z["kat_atan2f"])
