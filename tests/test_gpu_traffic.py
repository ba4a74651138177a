"""SURVEY.md 8(f) N4 on the GPU: k_traffic_advance through the C ABI (scgpuTraffic*) against the plain-C oracle
(sco_traffic_ai_on_rails, pinned to the reference's TrafficAISystem by tests/test_traffic_oracle.py) and against the
frames recorded from the reference itself (tests/golden/traffic.npz). Bit-exact: lane ids, laneS, targetSpeed,
lookAheadDist, the local TRS written into HBM, the dirty set, and — after scgpuUpdate — the world matrices and
visible lists the moved vehicles (and the wheels parented to them) produce."""
import numpy as np
import pytest

from oracle_bind import LANE_KEYS, PortScene, port_traffic_step
from scenarios import GpuAdapter, compare_frame, load_golden
from scgpu import scenes

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def set_lanes(scene, g):
    scene.traffic_set_lanes(*[g[k] for k in LANE_KEYS], default_speed=float(g["default_speed"]))


def check_agents(s, agents, what):
    lane, ls, v, look = s.traffic_read_agents()
    assert np.array_equal(lane, agents["lane"]), what + ": lane ids"
    assert bits_equal(ls, agents["s"]), what + ": laneS"
    assert bits_equal(v, agents["speed"]), what + ": targetSpeed"
    assert bits_equal(look, agents["look"]), what + ": lookAheadDist"


def test_golden_frames_of_the_reference():
    z = load_golden("traffic.npz")
    g = {k: z["g_" + k] for k in LANE_KEYS}
    g["default_speed"] = np.float32(z["g_default_speed"])
    n = len(z["a_lane"])
    e = np.arange(1, n + 1, dtype=np.uint32)
    s = GpuAdapter(n + 4, max_views=1).s
    s.spawn(e, z["trs"])
    set_lanes(s, g)
    s.traffic_set_agents(e, z["a_lane"], z["a_s"], z["a_speed"], z["a_look"])
    for k, dt in enumerate(z["dts"]):
        moved = s.traffic_advance(float(dt))
        assert moved == int(z["f_dirty"][k].sum()), f"frame {k}: moved count"
        check_agents(s, dict(lane=z["f_lane"][k], s=z["f_s"][k], speed=z["f_speed"][k], look=z["f_look"][k]), f"frame {k}")
        assert bits_equal(s.read_local(e), z["f_trs"][k]), f"frame {k}: local TRS"
    s.close()


@pytest.mark.parametrize("seed,debug", [(0, None), (1, (20.0, 0.5)), (2, None)])
def test_hostile_graphs_and_inputs_match_oracle(seed, debug):
    g = scenes.lane_random(300, 800, seed=200 + seed, hostile=True)
    n = 5000
    agents, trs = scenes.traffic_agents(g, n, seed=seed)
    rng = np.random.default_rng(seed)
    e = (np.arange(n, dtype=np.uint32) + 1) | (np.uint32(3) << 24)  # generation 3 handles
    brake = rng.random(n).astype(np.float32)
    brake[rng.random(n) < 0.5] = 0
    skip = (rng.random(n) < 0.1).astype(np.uint8)
    s = GpuAdapter(n + 4, max_views=1).s
    s.spawn(e[: n - 50], trs[: n - 50])  # the last 50 agents own no Transform: skipped by the ForEach
    set_lanes(s, g)
    s.traffic_set_agents(e, agents["lane"], agents["s"], agents["speed"], agents["look"])
    live = np.arange(n) < n - 50
    dts = [1 / 60] * 6 + [0.0, 5.0, 1e-8, 40.0, float("nan"), 1 / 60, 1e30, 1 / 60]
    for k, dt in enumerate(dts):
        pa = {key: v[live].copy() for key, v in agents.items()}
        pt = trs[live].copy()
        pm = port_traffic_step(g, pa, pt, dt, brake[live].copy(), skip[live].copy(), debug)
        for key in agents:
            agents[key][live] = pa[key]
        trs[live] = pt
        moved = s.traffic_advance(dt, brake, skip, debug)
        assert moved == int(pm.sum()), f"frame {k}"
        check_agents(s, agents, f"frame {k}")
        assert bits_equal(s.read_local(e[live]), trs[live]), f"frame {k}: local TRS"
        if k == 7:  # a sector unloads: TrafficLaneGraph::removeSector
            off = np.nonzero(rng.random(800) < 0.2)[0].astype(np.uint32)
            g["seg_active"][off] = 0
            s.traffic_set_lane_active(off, np.zeros(len(off), np.uint8))
    s.close()


def test_moved_vehicles_are_dirty_for_the_frame_update():
    """vehicles with four wheels parented to them drive over the procedural lanes of a 12 x 12 sector city among
    static props: after every traffic pass the frame update must recompute exactly what the reference's
    TransformSystem recomputes (vehicle + inherited-dirty wheels) and cull / list them identically."""
    grid = scenes.lane_grid(12, 12)
    nv = 3000
    agents, vtrs = scenes.traffic_agents(grid, nv, seed=3, hostile=False)
    props = scenes.city_flat(20_000, seed=8)
    rng = np.random.default_rng(1)
    wheel = np.zeros((nv * 4, 9), np.float32)
    wheel[:, 0:3] = np.tile(np.float32([[0.9, -0.4, 1.4], [-0.9, -0.4, 1.4], [0.9, -0.4, -1.4], [-0.9, -0.4, -1.4]]), (nv, 1))
    wheel[:, 3] = rng.uniform(0, 6.28, nv * 4)
    wheel[:, 6:9] = 0.35
    n = nv * 5 + 20_000
    e = np.arange(1, n + 1, dtype=np.uint32)
    ve, we, pe = e[:nv], e[nv: nv * 5], e[nv * 5:]
    g = GpuAdapter(n + 8, max_views=5)
    p = PortScene()
    vps = scenes.standard_views(5, center=(384.0, 6.0, 400.0))
    for s in (g, p):
        s.spawn(ve, vtrs)
        s.spawn(we, wheel, np.repeat(ve, 4))
        s.spawn(pe, props["trs9"], None, props["aabb6"], props["mesh_mat"], props["flags"])
        s.update(vps)
    compare_frame(g, p, e, 5, "traffic city, frame 0")
    g.s.traffic_set_lanes(*[grid[k] for k in LANE_KEYS], default_speed=float(grid["default_speed"]))
    g.s.traffic_set_agents(ve, agents["lane"], agents["s"], agents["speed"], agents["look"])
    for k in range(6):
        dt = 1 / 60 if k < 4 else 0.8
        pt = p.trs[:nv].copy()
        pm = port_traffic_step(grid, agents, pt, dt)
        p.set_local(ve[pm != 0], pt[pm != 0])
        assert g.s.traffic_advance(dt) == int(pm.sum())
        for s in (g, p):
            s.update(vps)
        assert g.recomputed == p.recomputed == int(pm.sum()) * 5, f"frame {k + 1}: recomputed {g.recomputed} vs {p.recomputed}"
        compare_frame(g, p, e, 5, f"traffic city, frame {k + 1}")
        check_agents(g.s, agents, f"frame {k + 1}")
    g.close()
