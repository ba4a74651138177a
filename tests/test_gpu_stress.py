"""Repetition stress in place of compute-sanitizer's racecheck (closed on the GPU pool): the same frame computed a hundred
and more times must give the same bits every time. The one shared-memory race this code base has had (a cross-proxy WAR
hazard in k_update_flat's TMA ring, round 1) showed up about once in 30 frames; the per-warp cp.async double buffers of
k_update_win, the asynchronous stored-matrix copies of partly dirty windows, k_compact's scans and the counting sort are
exercised the same way here. The parity suite itself also runs against the checked build (device-side assertions on
every data-derived index, `make -C sc-gameengine_b200 libscgpu_checked.so`, SCGPU_LIB=...)."""
import zlib

import numpy as np
import pytest

import scgpu
from scgpu import scenes

pytestmark = pytest.mark.gpu


def _digest(s, e, views):
    h = zlib.crc32(s.read_world(e).tobytes())
    for v in range(views):
        h = zlib.crc32(s.read_visible(v).tobytes(), h)
    c = s.counts()
    return h, int(c.recomputed), int(c.renderablesTotal)


@pytest.mark.parametrize("kind,n,frames", [("flat", 300_000, 120), ("hier", 300_000, 120)])
def test_all_dirty_frames_repeat_bit_for_bit(kind, n, frames):
    sc = scenes.city_flat(n, seed=12) if kind == "flat" else scenes.city_hier(n, seed=12)
    e = np.arange(n, dtype=np.uint32)
    s = scgpu.Scene(n, max_views=5, max_entity_index=n)
    s.spawn(e, sc["trs9"], scenes.parent_handles(sc["parent"], e), sc["aabb6"], sc["mesh_mat"], sc["flags"])
    s.set_views(scenes.standard_views(5))
    s.update()
    first = _digest(s, e, 5)
    assert first[1] == n
    for f in range(frames):
        s.mark_all_dirty()
        s.update()
        assert _digest(s, e, 5) == first, f"frame {f} differs from frame 0"
    s.close()


def test_partly_dirty_frames_repeat_bit_for_bit():
    """the same 30 % get the same TRS again every frame: recomputed count, matrices and lists must not move"""
    n, frames = 300_000, 120
    rng = np.random.default_rng(3)
    sc = scenes.city_hier(n, seed=13)
    e = np.arange(n, dtype=np.uint32)
    s = scgpu.Scene(n, max_views=5, max_entity_index=n)
    s.spawn(e, sc["trs9"], scenes.parent_handles(sc["parent"], e), sc["aabb6"], sc["mesh_mat"], sc["flags"])
    s.set_views(scenes.standard_views(5))
    s.update()
    idx = np.sort(rng.choice(n, (3 * n) // 10, replace=False)).astype(np.uint32)
    trs = sc["trs9"][idx].copy()
    trs[:, 0] += np.float32(0.5)
    s.set_local(idx, trs)
    s.update()
    first = _digest(s, e, 5)
    for f in range(frames):
        s.set_local(idx, trs)
        s.update()
        assert _digest(s, e, 5) == first, f"frame {f} differs"
    s.close()


def test_sorted_draws_repeat_bit_for_bit():
    n = 200_000
    rng = np.random.default_rng(4)
    sc = scenes.city_flat(n, seed=14)
    mm = np.stack([rng.integers(0, 200, n), rng.integers(0, 900, n)], axis=1).astype(np.uint32)
    e = np.arange(n, dtype=np.uint32)
    s = scgpu.Scene(n, max_views=1, max_entity_index=n)
    s.spawn(e, sc["trs9"], None, sc["aabb6"], mm, sc["flags"])
    s.set_views(scenes.standard_views(1))
    s.update(scgpu.UPDATE_FREEZE_CULLING)
    mp = rng.integers(0, 2, 900).astype(np.uint32)
    items0, runs0 = s.sorted_draws(0, mp, 200)
    for f in range(40):
        items, runs = s.sorted_draws(0, mp, 200)
        assert np.array_equal(items["entity"], items0["entity"]) and np.array_equal(runs, runs0), f"sort {f} differs"
    s.close()


def test_dirty_stamp_wrap_keeps_clean_nodes_clean(monkeypatch):
    """The dirty stamp is the 24-bit id of the update that must recompute an instance. Across the wrap of that id every
    stored stamp is reset once (k_clear_stamps), so that no stamp of 16.7 M updates ago resurrects as 'dirty now':
    the recomputed count must stay what the oracle's dirty flags give, frame by frame, through the wrap."""
    from oracle_bind import PortScene
    from scenarios import GpuAdapter, compare_frame
    monkeypatch.setenv("SCGPU_TEST_FIRST_FRAME", str(0xFFFFFF - 6))
    n = 30_000
    rng = np.random.default_rng(8)
    sc = scenes.city_hier(n, seed=15)
    e = np.arange(n, dtype=np.uint32)
    par = scenes.parent_handles(sc["parent"], e)
    g, p = GpuAdapter(n, max_views=3), PortScene()
    vps = scenes.standard_views(3)
    for s in (g, p):
        s.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        s.update(vps)
    for frame in range(16):   # ids 0xFFFFFA .. wrap .. 0x00000A
        idx = rng.choice(n, 500 + 37 * frame, replace=False)
        trs = sc["trs9"][idx].copy()
        trs[:, 0] += np.float32(0.5 + frame)
        for s in (g, p):
            s.set_local(e[idx], trs)
            s.update(vps)
        assert g.recomputed == p.recomputed, (frame, g.recomputed, p.recomputed)
        compare_frame(g, p, e, 3, f"wrap frame {frame}")
    g.close()
