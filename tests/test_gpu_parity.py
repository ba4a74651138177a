"""GPU parity proper: the CUDA path through the C ABI (libscgpu.so) against the plain-C oracle, and against the
compiled reference (oracle/_ref) when it travelled to the box. Bit-exact: ordered visible/culled lists, counts,
draw items and world matrices (NaNs compare equal to NaNs)."""
import numpy as np
import pytest

import oracle_bind
from oracle_bind import PortScene, RefScene
from scenarios import (INVALID, GpuAdapter, assert_same_bits, compare_draws, compare_frame, random_aabb,
                       random_forest, random_trs)
from scgpu import scenes

pytestmark = pytest.mark.gpu


def _checkers():
    out = [("port", PortScene())]
    if oracle_bind.ref_available():
        out.append(("ref", RefScene()))
    return out


def _entities(checkers, n):
    for name, s in checkers:
        if name == "ref":
            return s.create_entities(n)
    return np.arange(n, dtype=np.uint32)


def test_create_reports_device():
    import scgpu
    s = scgpu.Scene(1024, max_views=2)
    assert s.lib.scgpuGetApiVersion() == 2
    s.close()


@pytest.mark.parametrize("n,seed", [(1, 1), (255, 2), (1024, 3), (1025, 4), (50_000, 5)])
def test_flat_city_matches_oracle(n, seed):
    sc = scenes.city_flat(n, seed=seed)
    chk = _checkers()
    e = _entities(chk, n)
    g = GpuAdapter(n + 8, max_views=5)
    vps = scenes.standard_views(5)
    for _, s in chk + [("gpu", g)]:
        s.spawn(e, sc["trs9"], None, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        s.update(vps)
    for name, s in chk:
        compare_frame(g, s, e, 5, f"flat n={n} vs {name}")
        compare_draws(g, s, 0, f"flat n={n} vs {name}")
        compare_draws(g, s, 5, f"flat n={n} budget vs {name}")
    assert g.recomputed == n
    # second frame, nothing dirty: nothing recomputed, same lists
    g.update(vps)
    assert g.recomputed == 0
    compare_frame(g, chk[0][1], e, 5, "flat clean frame")
    g.close()


@pytest.mark.parametrize("n,seed", [(10, 1), (3000, 2), (40_000, 3)])
def test_depth4_city_matches_oracle(n, seed):
    sc = scenes.city_hier(n, seed=seed)
    chk = _checkers()
    e = _entities(chk, n)
    par = scenes.parent_handles(sc["parent"], e)
    g = GpuAdapter(n, max_views=5)
    vps = scenes.standard_views(5)
    rng = np.random.default_rng(seed)
    for _, s in chk + [("gpu", g)]:
        s.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        s.update(vps)
    for name, s in chk:
        compare_frame(g, s, e, 5, f"hier n={n} frame0 vs {name}")
    # 30 % dirty
    idx = rng.choice(n, max(1, n * 3 // 10), replace=False)
    t2 = sc["trs9"][idx].copy()
    t2[:, 0:3] += rng.normal(size=(len(idx), 3)).astype(np.float32)
    t2[:, 4] += np.float32(0.1)
    for _, s in chk + [("gpu", g)]:
        s.set_local(e[idx], t2)
        s.update(vps)
    for name, s in chk:
        compare_frame(g, s, e, 5, f"hier n={n} frame1 vs {name}")
        compare_draws(g, s, 0, f"hier n={n} vs {name}")
        assert g.recomputed == s.recomputed if name == "port" else True
    g.close()


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16])
def test_random_forest_edge_cases(seed):
    rng = np.random.default_rng(seed)
    chk = _checkers()
    n = 1500
    e = _entities(chk, n)
    g = GpuAdapter(4096, max_views=3)
    parent_idx = random_forest(rng, n, max_back=600 if seed % 2 else 40)  # long jumps cross sub-tiles
    trs = random_trs(rng, n, spread=60.0)
    flags = rng.choice([0, 1, 2, 3], size=n, p=[0.05, 0.1, 0.15, 0.7]).astype(np.uint32)
    trs[5, 6:9] = 0.0
    trs[6, 6:8] = 0.0
    trs[7, 0] = np.nan
    trs[8, 4] = np.inf
    trs[9, 3:6] = [1e6, -3e9, 1e-30]
    trs[10, 3:6] = [120.0, -119.99, 0.78539819]
    trs[11, 6] = -2.0
    par = scenes.parent_handles(parent_idx, e)
    par[20] = e[20]
    par[21] = 0x00ABCDEF
    par[30], par[31] = e[31], e[30]
    par[40], par[41], par[42] = e[41], e[42], e[40]
    par[43] = e[40]
    par[300], par[900] = e[900], e[300]   # cycle across sub-tiles
    par[901] = e[300]
    aabb = random_aabb(rng, n)
    mm = rng.integers(0, 50, size=(n, 2)).astype(np.uint32)
    everyone = chk + [("gpu", g)]
    vps = scenes.standard_views(3, center=(0.0, 10.0, 80.0))
    for _, s in everyone:
        s.spawn(e, trs, par, aabb, mm, flags)
        s.update(vps)
    for name, s in chk:
        compare_frame(g, s, e, 3, f"frame0 vs {name}")
    assert np.array_equal(g.s.read_parents(e), chk[0][1].parent), "parent fix-ups differ"

    idx = rng.choice(n, 200, replace=False)
    t2 = random_trs(rng, 200, spread=60.0)
    for _, s in everyone:
        s.set_local(e[idx], t2)
        s.set_parent(e[[40, 50, 51]], np.array([INVALID, e[52], e[20]], np.uint32))
        s.update(vps)
    for name, s in chk:
        compare_frame(g, s, e, 3, f"frame1 vs {name}")

    dead = rng.choice(n, 300, replace=False)
    dead_handles = np.concatenate([e[dead], np.array([0x00FFFFF0, e[dead[0]]], np.uint32)])
    for _, s in everyone:
        s.despawn(dead_handles)
    live = np.setdiff1d(np.arange(n), dead)
    if len(chk) > 1:
        e2 = chk[1][1].create_entities(100)
    else:
        e2 = np.arange(n, n + 100, dtype=np.uint32)
    trs3 = random_trs(rng, 100, spread=60.0)
    par3 = e[rng.choice(live, 100)]
    bb3 = random_aabb(rng, 100)
    for _, s in everyone:
        s.spawn(e2, trs3, par3, bb3, None, None)
        s.update(vps)
    assert np.array_equal(g.dense_entities(), chk[0][1].entity), "pool order after swap-remove differs"
    alive = np.concatenate([e[live], e2])
    for name, s in chk:
        compare_frame(g, s, alive, 3, f"frame2 vs {name}")
        compare_draws(g, s, 0, f"frame2 vs {name}")
        compare_draws(g, s, 13, f"frame2 budget vs {name}")
    for _, s in everyone:
        s.update(vps, freeze=True)
    for name, s in chk:
        compare_frame(g, s, alive, 3, f"freeze vs {name}")
    # mark-all-dirty must reproduce the same matrices (idempotence)
    before = g.read_world(alive)
    g.mark_all_dirty()
    g.update(vps)
    assert_same_bits(before, g.read_world(alive), "idempotent recompute")
    g.close()


def test_views_zero_and_identity_matrix():
    """Zero VP => zero planes => everything visible; identity VP => unit cube planes"""
    n = 2000
    sc = scenes.city_flat(n, seed=9)
    chk = _checkers()
    e = _entities(chk, n)
    g = GpuAdapter(n, max_views=2)
    vps = np.stack([np.zeros(16, np.float32), np.eye(4, dtype=np.float32).ravel()])
    for _, s in chk + [("gpu", g)]:
        s.spawn(e, sc["trs9"], None, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        s.update(vps)
    assert len(g.visible[0]) == n
    for name, s in chk:
        compare_frame(g, s, e, 2, f"special views vs {name}")
    pl = g.s.get_view_planes(1)
    assert np.array_equal(pl, np.array([[1, 0, 0, 1], [-1, 0, 0, 1], [0, 1, 0, 1], [0, -1, 0, 1], [0, 0, 1, 1], [0, 0, -1, 1]], np.float32))
    g.close()


def test_empty_scene_and_errors():
    import scgpu
    s = scgpu.Scene(16, max_views=1)
    with pytest.raises(scgpu.ScGpuError):
        s.update()  # no views
    s.set_views(np.eye(4, dtype=np.float32).ravel())
    s.update()
    c = s.counts()
    assert (c.transforms, c.renderablesTotal, c.visible[0]) == (0, 0, 0)
    assert len(s.read_visible(0)) == 0
    with pytest.raises(scgpu.ScGpuError):
        s.spawn(np.arange(17, dtype=np.uint32), np.zeros((17, 9), np.float32))  # over capacity
    s.spawn(np.array([3], np.uint32), np.array([[0, 0, 0, 0, 0, 0, 1, 1, 1]], np.float32))
    with pytest.raises(scgpu.ScGpuError):
        s.spawn(np.array([3], np.uint32), np.zeros((1, 9), np.float32))  # duplicate Transform
    s.close()


@pytest.mark.parametrize("hier", [False, True])
def test_frustum_boundary_shell(hier):
    """Whole warps of instances whose bounding spheres graze a frustum plane (signed distance + radius within a few
    1e-6 of zero, relative): the conservative warp prefilter and the early-outs must never change a single decision."""
    rng = np.random.default_rng(77)
    vps = scenes.standard_views(5)
    chk = _checkers()[:1]
    port = chk[0][1]
    planes = np.zeros(24, np.float32)
    n = 64 * 512
    trs = np.zeros((n, 9), np.float32)
    trs[:, 6:9] = 1.0
    half_diag = np.float32(np.sqrt(0.75))  # unit cube: |extent| = sqrt(3)/2, scale 1
    for blk in range(n // 64):
        v, pl = blk % 5, (blk // 5) % 6
        oracle_bind.port_lib().sco_frustum_from_viewproj(vps[v].ctypes.data_as(oracle_bind.C.c_void_p),
                                                         planes.ctypes.data_as(oracle_bind.C.c_void_p))
        nrm, d = planes[pl * 4: pl * 4 + 3].astype(np.float64), float(planes[pl * 4 + 3])
        if not np.any(nrm):
            continue
        # a point on the plane shifted to signed distance -(radius) * (1 + eps), eps in +-3e-6, jittered along the plane
        base = -nrm * d / np.dot(nrm, nrm)
        t1 = np.cross(nrm, [0.3, 0.9, 0.1]); t1 /= np.linalg.norm(t1)
        for k in range(64):
            eps = rng.uniform(-3e-6, 3e-6)
            c = base - nrm * half_diag * (1.0 + eps) + t1 * rng.uniform(-0.5, 0.5)
            trs[blk * 64 + k, 0:3] = c.astype(np.float32)
    e = _entities(chk, n)
    par = None
    if hier:  # pairs: odd slots are children of the even slot before them, with a tiny local offset
        parent_idx = np.full(n, -1, np.int64)
        parent_idx[1::2] = np.arange(0, n, 2)
        child = trs[1::2].copy()
        trs[1::2, 0:3] = rng.uniform(-1e-3, 1e-3, size=(n // 2, 3)).astype(np.float32)
        trs[1::2, 3:9] = child[:, 3:9]
        par = scenes.parent_handles(parent_idx, e)
    g = GpuAdapter(n, max_views=5)
    for s in (port, g):
        s.spawn(e, trs, par, None, None, None)
        s.update(vps)
    compare_frame(g, port, e, 5, "boundary shell")
    tot = sum(len(v) for v in port.visible)
    assert 0 < tot < 5 * n, "the shell must straddle the planes"
    g.close()


@pytest.mark.parametrize("n,seed", [(60_000, 21)])
def test_streaming_churn_matches_oracle(n, seed):
    """BASELINE.json configs[4] in miniature: per frame 10 % of the instances despawn (any node of a group, so
    children lose their parents and parents their children), 10 % new groups spawn, 30 % get a new local TRS.
    After every frame: pool order, ordered visible lists and all world matrices against the oracle. Exercises
    the window builder on topologies that swap-remove has shuffled (long links, windows flagged for the generic
    path, partially dirty windows -> run-time level schedule)."""
    rng = np.random.default_rng(seed)
    sc = scenes.city_hier(n, seed=seed)
    p = PortScene()
    g = GpuAdapter(2 * n, max_views=5, max_entity_index=4 * n)
    vps = scenes.standard_views(5)
    e = np.arange(n, dtype=np.uint32)
    par = scenes.parent_handles(sc["parent"], e)
    for s in (g, p):
        s.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        s.update(vps)
    compare_frame(g, p, e, 5, "churn frame 0")
    alive = e.copy()
    next_id = n
    for frame in range(1, 6):
        dead = rng.choice(alive, len(alive) // 10, replace=False)
        alive = np.setdiff1d(alive, dead)
        m = n // 10
        fresh = scenes.city_hier(m, seed=seed + 100 * frame)
        fe = np.arange(next_id, next_id + m, dtype=np.uint32)
        next_id += m
        fpar = scenes.parent_handles(fresh["parent"], fe)
        moved = rng.choice(alive, (3 * len(alive)) // 10, replace=False)
        trs = random_trs(rng, len(moved), spread=400.0)
        for s in (g, p):
            s.despawn(dead)
            s.spawn(fe, fresh["trs9"], fpar, fresh["aabb6"], fresh["mesh_mat"], fresh["flags"])
            s.set_local(moved, trs)
            s.update(vps)
        alive = np.concatenate([alive, fe])
        assert np.array_equal(g.dense_entities(), p.entity), f"pool order differs in frame {frame}"
        compare_frame(g, p, alive, 5, f"churn frame {frame}")
        assert g.recomputed == p.recomputed, (frame, g.recomputed, p.recomputed)
    compare_draws(g, p, 0, "churn draws")
    g.close()


@pytest.mark.parametrize("max_back,p_child,seed", [(6, 0.9, 31), (12, 0.8, 32), (200, 0.7, 33), (3, 0.97, 34)])
def test_random_forest_large(max_back, p_child, seed):
    """40k-node random forests: dense short links (no uncrossed cut position for long stretches -> forced cuts,
    parents outside the window), long links, chains deeper than a window, levels wider than 16 children. Two frames:
    everything dirty, then 20 % dirty (children inherit)."""
    rng = np.random.default_rng(seed)
    n = 40_000
    e = np.arange(n, dtype=np.uint32)
    parent_idx = random_forest(rng, n, p_child=p_child, max_back=max_back)
    trs = random_trs(rng, n, spread=300.0)
    trs[:, 6:9] = rng.uniform(0.7, 1.3, size=(n, 3)).astype(np.float32)  # keep deep chains finite
    par = scenes.parent_handles(parent_idx, e)
    aabb = random_aabb(rng, n)
    p = PortScene()
    g = GpuAdapter(n, max_views=3)
    vps = scenes.standard_views(3, center=(0.0, 10.0, 80.0))
    for s in (g, p):
        s.spawn(e, trs, par, aabb, None, None)
        s.update(vps)
    compare_frame(g, p, e, 3, f"forest max_back={max_back} frame 0")
    assert g.recomputed == p.recomputed == n
    idx = rng.choice(n, n // 5, replace=False)
    t2 = random_trs(rng, len(idx), spread=300.0)
    t2[:, 6:9] = rng.uniform(0.7, 1.3, size=(len(idx), 3)).astype(np.float32)
    for s in (g, p):
        s.set_local(e[idx], t2)
        s.update(vps)
    compare_frame(g, p, e, 3, f"forest max_back={max_back} frame 1")
    assert g.recomputed == p.recomputed
    g.close()


@pytest.mark.parametrize("max_draws,n_mesh,n_mat", [(0, 10, 37), (700, 10, 37), (0, 300, 1500), (0, 1, 1)])
def test_sorted_draws_and_runs_match_renderer_sort(max_draws, n_mesh, n_mat):
    """scgpuBuildSortedDraws (hand-written counting sort, one to three 8-bit passes depending on the
    size of the asset tables) against the restated renderer loop AND, when oracle/_ref is there, against the reference's
    own block of sc_vk.cpp (:1841-1912): same kept set, same (pipeline, material, mesh) order, runs exactly at the
    loop's bind points; ties in visible order; models bit-identical."""
    from oracle_bind import check_against_renderer, renderer_sorted_draws
    rng = np.random.default_rng(77)
    n = 30_000
    sc = scenes.city_flat(n, seed=9)
    mm = np.stack([rng.integers(0, n_mesh + 2, n), rng.integers(0, n_mat + 3, n)], axis=1).astype(np.uint32)  # meshId, materialId
    e = np.arange(n, dtype=np.uint32)
    g = GpuAdapter(n, max_views=2)
    vps = scenes.standard_views(2)
    g.spawn(e, sc["trs9"], None, sc["aabb6"], mm, sc["flags"])
    g.update(vps, freeze=True)  # every candidate visible: a long draw list
    mesh_count = n_mesh                                  # the two largest meshIds are out of range -> dropped
    mat_pipe = rng.integers(0, 2, n_mat).astype(np.uint32)  # the three largest materialIds are beyond the table
    if n_mat > 20:
        mat_pipe[[3, 17]] = 0xFFFFFFFF                   # getMaterial() == nullptr
    draws, emitted, dropped = g.read_draw_items(0, max_draws)
    order, runs = renderer_sorted_draws(draws, mat_pipe, mesh_count)
    items, gruns = g.s.sorted_draws(0, mat_pipe, mesh_count, max_draws)
    assert len(items) == len(order) and len(order) < emitted
    want = draws[order]
    for f in ("entity", "meshId", "materialId"):
        assert np.array_equal(items[f], want[f]), f
    assert_same_bits(items["model"], want["model"], "sorted draw models")
    assert [tuple(int(x) for x in r) for r in gruns] == runs
    assert sum(r[4] for r in runs) == len(items)
    if oracle_bind.ref_available():
        # ours -> positions in `draws` through the entity handle (unique per draw), then against the reference's loop
        where = {int(h): i for i, h in enumerate(draws["entity"])}
        mine = np.array([where[int(h)] for h in items["entity"]], np.int64)
        check_against_renderer(draws, mat_pipe, mesh_count, mine, [tuple(int(x) for x in r) for r in gruns], "gpu vs sc_vk.cpp")
    # an empty table drops everything
    items0, runs0 = g.s.sorted_draws(0, np.zeros(0, np.uint32), mesh_count, max_draws)
    assert len(items0) == 0 and len(runs0) == 0
    g.close()


@pytest.mark.parametrize("hier", [False, True])
def test_radius_overflow_is_never_culled_early(hier):
    """The early 'certainly culled' exit of sphere_cull_warp works with an upper bound of the radius. Where the
    reference's own radius overflows (column norms beyond sqrt(FLT_MAX) -> radius Inf -> 'd < -radius' is false ->
    VISIBLE) the bound must give up too, however far outside the frustum the instance lies."""
    n = 64
    trs = np.zeros((n, 9), np.float32)
    trs[:, 6:9] = 1.0
    trs[:, 0] = np.linspace(1e3, 1e6, n)          # ordinary far-away instances: culled
    big = [3, 17, 40]
    trs[big, 0] = [1e25, -3e30, 5e21]
    trs[big, 6] = [1e20, 4e19, 2.5e19]              # column norm^2 overflows
    trs[41, 0] = 1e25; trs[41, 6] = 1e17            # large but finite radius: culled like the oracle says
    e = np.arange(n, dtype=np.uint32)
    par = None
    if hier:
        idx = np.full(n, -1, np.int64)
        idx[1:8] = 0                                 # a group at the front, the rest roots
        par = scenes.parent_handles(idx, e)
    p, g = PortScene(), GpuAdapter(n, max_views=2)
    vps = scenes.standard_views(2)
    for s in (g, p):
        s.spawn(e, trs, par, None, None, None)
        s.update(vps)
    compare_frame(g, p, e, 2, "radius overflow")
    assert all(b in g.visible[0] for b in big)
    g.close()


@pytest.mark.parametrize("views", [1, 2, 4, 6, 7, 8])
def test_every_view_count_matches_oracle(views):
    """Every instantiation of the fused kernels (V = 1..8 views, flat and hierarchical) against the oracle: the
    adaptive plane order keeps 3 bits per view, the scatter packs 5 views per 64-bit scan."""
    n = 30_000
    vps = scenes.standard_views(views, center=(40.0, 8.0, 60.0))
    for kind in ("flat", "hier"):
        sc = scenes.city_flat(n, seed=50 + views) if kind == "flat" else scenes.city_hier(n, seed=60 + views)
        e = np.arange(n, dtype=np.uint32)
        par = scenes.parent_handles(sc["parent"], e)
        p, g = PortScene(), GpuAdapter(n, max_views=views)
        for s in (g, p):
            s.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
            s.update(vps)
        compare_frame(g, p, e, views, f"{kind}, {views} views")
        assert sum(len(v) for v in g.visible) > 0
        g.close()
