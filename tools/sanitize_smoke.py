#!/usr/bin/env python
"""Small scenes through every kernel of the frame and of the delta path, meant to run under compute-sanitizer:

  compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py      (shared-memory hazards: k_update_win's per-warp
                                                                          staging + cp.async, k_update_flat's TMA ring,
                                                                          k_compact's scans, the sort's histograms)
  compute-sanitizer --tool synccheck python tools/sanitize_smoke.py

Results are also checked against the plain-C oracle, so a run that passes did the real work."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import scgpu  # noqa: E402
from oracle_bind import PortScene  # noqa: E402
from scenarios import GpuAdapter, compare_draws, compare_frame, random_trs  # noqa: E402
from scgpu import scenes  # noqa: E402


def flat(n=5000):
    sc = scenes.city_flat(n, seed=3)
    e = np.arange(n, dtype=np.uint32)
    g, p = GpuAdapter(n + 64, max_views=5), PortScene()
    vps = scenes.standard_views(5)
    for s in (g, p):
        s.spawn(e, sc["trs9"], None, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        s.update(vps)
    compare_frame(g, p, e, 5, "flat")
    g.update(vps)                       # clean frame
    compare_frame(g, p, e, 5, "flat clean")
    g.close()


def hier_and_churn(n=12000):
    rng = np.random.default_rng(1)
    sc = scenes.city_hier(n, seed=4)
    e = np.arange(n, dtype=np.uint32)
    par = scenes.parent_handles(sc["parent"], e)
    g, p = GpuAdapter(2 * n, max_views=5, max_entity_index=8 * n), PortScene()
    vps = scenes.standard_views(5)
    for s in (g, p):
        s.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        s.update(vps)
    compare_frame(g, p, e, 5, "hier all dirty")
    alive, nxt = e.copy(), n
    for frame in range(3):
        dead = rng.choice(alive, len(alive) // 10, replace=False)
        alive = np.setdiff1d(alive, dead)
        m = len(dead)
        fresh = scenes.city_hier(m + 16, seed=50 + frame)
        tp = np.where(fresh["parent"][:m] < m, fresh["parent"][:m], -1)
        fe = np.arange(nxt, nxt + m, dtype=np.uint32)
        nxt += m
        moved = rng.choice(alive, len(alive) // 3, replace=False)
        trs = random_trs(rng, len(moved), spread=300.0)
        for s in (g, p):
            s.despawn(dead)
            s.spawn(fe, fresh["trs9"][:m], scenes.parent_handles(tp, fe), fresh["aabb6"][:m], fresh["mesh_mat"][:m], fresh["flags"][:m])
            s.set_local(moved, trs)
            s.update(vps)
        alive = np.concatenate([alive, fe])
        assert np.array_equal(g.dense_entities(), p.entity)
        compare_frame(g, p, p.entity, 5, f"churn frame {frame}")
    compare_draws(g, p, 0, "draws")
    g.s.sorted_draws(0, np.array([0, 1, 1, 0, 1], np.uint32), 4)
    g.s.set_local_pos_rot(alive[:100], np.zeros((100, 6), np.float32))
    g.s.set_local_range(0, np.zeros((50, 3), np.float32), 3)
    g.s.set_render(alive[:10], np.ones((10, 2), np.uint32), None, np.full(10, 3, np.uint32))
    g.update(vps, freeze=True)
    g.update(vps, skip_transform=True)
    g.close()


def sectors():
    s = scgpu.Scene(4096, max_views=1)
    gen = scgpu.SectorGen(sectorSizeMeters=64.0, seed=424242, propsPerSectorMin=18, propsPerSectorMax=34, includeGroundPlane=1,
                          meshCube=1, meshTriangle=2, matUnlit=1, matChecker=2, matTest=3)
    coords = np.array([[0, 0], [1, 0], [0, 1]], np.int32)
    n = sum(s.lib.scgpuSectorSpawnCount(gen, int(x), int(z)) for x, z in coords)
    s.spawn_sectors(gen, coords, np.arange(n, dtype=np.uint32))
    s.set_views(scenes.standard_views(1))
    s.update()
    assert s.counts().transforms == n
    s.close()


if __name__ == "__main__":
    flat()
    hier_and_churn()
    sectors()
    print("SANITIZE SMOKE OK")
