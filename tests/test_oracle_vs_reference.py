"""Pins the plain-C oracle (oracle/scoracle.c) to the reference's own compiled code (oracle/_ref/libscref.so).
CPU only. Skipped when the reference build is absent."""
import numpy as np
import pytest

import oracle_bind
from oracle_bind import PortScene, RefScene
from scenarios import (INVALID, assert_same_bits, compare_draws, compare_frame, random_aabb, random_forest,
                       random_trs)
from scgpu import scenes


def _mk(ref_needed=True):
    if not oracle_bind.ref_available():
        pytest.skip("reference build absent")
    return RefScene(), PortScene()


def test_default_sandbox_scene_counts(ref):
    """SURVEY.md §8c scene golden: 25 sectors / 721 transforms / 719 renderables / 166 visible / 553 culled"""
    import ctypes as C
    w = ref.screfCreate(0)
    sectors = ref.screfBuildDefaultScene(w, 60)
    t, v, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    ref.screfGetCullStats(w, C.byref(t), C.byref(v), C.byref(c))
    assert (sectors, ref.screfTransformCount(w), t.value, v.value, c.value) == (25, 721, 719, 166, 553)
    ref.screfDestroy(w)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_flat_scene_port_equals_reference(seed):
    r, p = _mk()
    sc = scenes.city_flat(3000, seed=seed)
    e = r.create_entities(sc["n"])
    for s in (r, p):
        s.spawn(e, sc["trs9"], None, sc["aabb6"], sc["mesh_mat"], sc["flags"])
    vps = scenes.standard_views(5)
    for s in (r, p):
        s.update(vps)
    compare_frame(r, p, e, 5, "flat")
    compare_draws(r, p, 0, "flat")
    compare_draws(r, p, 7, "flat budget")
    assert sum(len(v) for v in r.visible) > 0


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_random_forest_edge_cases_port_equals_reference(seed):
    rng = np.random.default_rng(seed)
    r, p = _mk()
    n = 700
    e = r.create_entities(n)
    parent_idx = random_forest(rng, n)
    trs = random_trs(rng, n, spread=60.0)
    flags = rng.choice([0, 1, 2, 3], size=n, p=[0.05, 0.1, 0.15, 0.7]).astype(np.uint32)
    # edge cases: zero scale, partial zero scale, NaN/Inf, huge angles, tiny angles
    trs[5, 6:9] = 0.0
    trs[6, 6:8] = 0.0
    trs[7, 0] = np.nan
    trs[8, 4] = np.inf
    trs[9, 3:6] = [1e6, -3e9, 1e-30]
    trs[10, 3:6] = [120.0, -119.99, 0.78539819]
    trs[11, 6] = -2.0
    par = scenes.parent_handles(parent_idx, e)
    par[20] = e[20]            # self parent
    par[21] = 0x00ABCDEF       # never-created entity
    par[30], par[31] = e[31], e[30]  # 2-cycle
    par[40], par[41], par[42] = e[41], e[42], e[40]  # 3-cycle
    par[43] = e[40]            # hangs off a cycle
    aabb = random_aabb(rng, n)
    mm = rng.integers(0, 50, size=(n, 2)).astype(np.uint32)
    for s in (r, p):
        s.spawn(e, trs, par, aabb, mm, flags)
    vps = scenes.standard_views(3, center=(0.0, 10.0, 80.0))
    for s in (r, p):
        s.update(vps)
    compare_frame(r, p, e, 3, "frame0")
    assert np.array_equal(r.read_parents(e), p.parent), "parent fix-ups differ"

    # frame 1: partial dirty, re-parenting (incl. breaking a cycle), stale-clean nodes
    idx = rng.choice(n, 90, replace=False)
    t2 = random_trs(rng, 90, spread=60.0)
    for s in (r, p):
        s.set_local(e[idx], t2)
        s.set_parent(e[[40, 50, 51]], np.array([INVALID, e[52], e[20]], np.uint32))
        s.update(vps)
    compare_frame(r, p, e, 3, "frame1")

    # frame 2: despawn (swap-remove order) incl. parents of live children and stale handles, then spawn more
    dead = rng.choice(n, 120, replace=False)
    dead_handles = np.concatenate([e[dead], np.array([0x00FFFFF0, e[dead[0]]], np.uint32)])
    for s in (r, p):
        s.despawn(dead_handles)
    e2 = r.create_entities(60)
    trs3 = random_trs(rng, 60, spread=60.0)
    live = np.setdiff1d(np.arange(n), dead)
    par3 = e[rng.choice(live, 60)]
    for s in (r, p):
        s.spawn(e2, trs3, par3, random_aabb(np.random.default_rng(5), 60), None, None)
        s.update(vps)
    assert np.array_equal(r.dense_entities(), p.entity), "pool order after swap-remove differs"
    alive = np.concatenate([e[live], e2])
    compare_frame(r, p, alive, 3, "frame2")
    compare_draws(r, p, 0, "frame2")
    # freeze culling: every candidate visible
    for s in (r, p):
        s.update(vps, freeze=True)
    compare_frame(r, p, alive, 3, "freeze")


def test_math_kats_port_equals_reference(ref, port):
    import ctypes as C
    rng = np.random.default_rng(7)
    f = lambda a: a.ctypes.data_as(C.c_void_p)
    for i in range(300):
        a = rng.normal(size=16).astype(np.float32)
        b = rng.normal(size=16).astype(np.float32)
        o1, o2 = np.zeros(16, np.float32), np.zeros(16, np.float32)
        ref.screfMat4Mul(f(a), f(b), f(o1)); port.sco_mat4_mul(f(a), f(b), f(o2))
        assert_same_bits(o1, o2, "mat4_mul")
        ref.screfMat4Inverse(f(a), f(o1)); port.sco_mat4_inverse(f(a), f(o2))
        assert_same_bits(o1, o2, "mat4_inverse")
        p1, p2 = np.zeros(24, np.float32), np.zeros(24, np.float32)
        ref.screfFrustumFromViewProj(f(a), f(p1)); port.sco_frustum_from_viewproj(f(a), f(p2))
        assert_same_bits(p1, p2, "frustum")
        bb = np.sort(rng.normal(size=(2, 3)).astype(np.float32), axis=0).ravel()
        c1, c2 = np.zeros(4, np.float32), np.zeros(4, np.float32)
        ref.screfWorldBoundsSphere(f(a), f(bb), f(c1), f(c1[3:])); port.sco_world_bounds_sphere(f(a), f(bb), f(c2), f(c2[3:]))
        assert_same_bits(c1, c2, "sphere")
    ref.screfMat4Perspective(C.c_float(1.0471975), C.c_float(16 / 9), C.c_float(0.1), C.c_float(1000.0), 1, f(o1))
    port.sco_mat4_perspective_rh_zo(C.c_float(1.0471975), C.c_float(16 / 9), C.c_float(0.1), C.c_float(1000.0), 1, f(o2))
    assert_same_bits(o1, o2, "perspective")


def test_long_streaming_churn_port_equals_reference():
    """30 frames of group-wise churn on a depth-4 city (the pattern of bench.py's churn leg in miniature): whole groups
    despawn, fresh groups spawn, a fifth of the live instances get a new local TRS. Swap-with-last scrambles the pool
    frame by frame (tools/model_churn_order.py: after 30 frames a quarter of the children sit 32 or more slots from their
    parent, a third of the parents FOLLOW their children) — the regime the GPU path's window builder and generic path
    live in. Pool order, ordered lists and world matrices of the oracle against the reference's own systems."""
    rng = np.random.default_rng(77)
    r, p = _mk()
    n = 3000
    sc = scenes.city_hier(n, seed=31)
    e = r.create_entities(n)
    par = scenes.parent_handles(sc["parent"], e)
    vps = scenes.standard_views(5)
    for s in (r, p):
        s.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        s.update(vps)
    compare_frame(r, p, e, 5, "churn frame 0")
    roots = np.nonzero(sc["parent"] < 0)[0]
    groups = [e[a:b] for a, b in zip(roots, np.append(roots[1:], n))]
    far_seen = 0
    for frame in range(1, 31):
        pick = rng.choice(len(groups), max(1, len(groups) // 12), replace=False)
        dead = np.concatenate([groups[g] for g in pick])
        groups = [g for i, g in enumerate(groups) if i not in set(pick.tolist())]
        m = len(dead)
        tmpl = scenes.city_hier(m + 8, seed=500 + frame)
        tp = np.where(tmpl["parent"][:m] < m, tmpl["parent"][:m], -1)
        fe = r.create_entities(m)                      # the reference recycles freed indices with a new generation
        fpar = scenes.parent_handles(tp, fe)
        tr = np.nonzero(tp < 0)[0]
        groups += [fe[a:b] for a, b in zip(tr, np.append(tr[1:], m))]
        live = np.concatenate(groups)
        moved = rng.choice(live, len(live) // 5, replace=False)
        trs = random_trs(rng, len(moved), spread=300.0)
        for s in (r, p):
            s.despawn(dead)
            s.spawn(fe, tmpl["trs9"][:m], fpar, tmpl["aabb6"][:m], tmpl["mesh_mat"][:m], tmpl["flags"][:m])
            s.set_local(moved, trs)
            s.update(vps)
        dense = r.dense_entities()
        assert np.array_equal(dense, p.entity), f"pool order differs in frame {frame}"
        compare_frame(r, p, dense, 5, f"churn frame {frame}")
        if frame % 10 == 0:
            compare_draws(r, p, 0, f"churn frame {frame}")
            slot = {int(h): i for i, h in enumerate(dense)}
            ph = r.read_parents(dense)
            far_seen = sum(1 for i, h in enumerate(ph) if int(h) in slot and abs(slot[int(h)] - i) >= 32)
    assert far_seen > n // 20, "the scene did not get scrambled: the test does not reach the regime it is about"


def test_renderer_sort_restatement_equals_the_reference_loop():
    """SURVEY 8(f) N1 oracle pinned: the numpy restatement of the renderer's filter + sort + bind-on-change loop
    (oracle_bind.renderer_sorted_draws) against the reference's own block of sc_vk.cpp (:1841-1912) compiled into
    oracle/_ref: kept set, key order, bind points. Unknown materials, null materials, out-of-range meshes, empty input."""
    from oracle_bind import DRAW_ITEM_DTYPE, check_against_renderer, ref_renderer_submit, renderer_sorted_draws
    rng = np.random.default_rng(5)
    for n, n_mat, n_mesh in ((0, 4, 3), (1, 1, 1), (257, 9, 5), (20_000, 41, 12), (3000, 300, 70)):
        d = np.zeros(n, DRAW_ITEM_DTYPE)
        d["entity"] = np.arange(n)
        d["meshId"] = rng.integers(0, n_mesh + 3, n)       # some out of range
        d["materialId"] = rng.integers(0, n_mat + 2, n)    # some beyond the table
        d["model"] = rng.normal(size=(n, 16))
        mp = rng.integers(0, 2, n_mat).astype(np.uint32)   # PipelineId::UnlitColor / ::Textured
        mp[rng.integers(0, n_mat, max(1, n_mat // 8))] = 0xFFFFFFFF  # getMaterial() == nullptr
        order, runs = renderer_sorted_draws(d, mp, n_mesh)
        check_against_renderer(d, mp, n_mesh, order, runs, f"n={n}")
        ro, rb = ref_renderer_submit(d, mp, n_mesh)
        assert len(ro) == 0 or rb[0] == 7  # the first draw binds pipeline, material and mesh
