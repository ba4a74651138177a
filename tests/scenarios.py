"""Scenario builders replayed on the GPU path, the plain-C oracle and the compiled reference alike."""
from __future__ import annotations

import numpy as np

import scgpu
from scgpu import scenes

INVALID = 0xFFFFFFFF


class GpuAdapter:
    """scgpu.Scene behind the same interface as oracle_bind.PortScene / RefScene."""

    def __init__(self, max_instances, max_views=8, max_entity_index=0):
        self.s = scgpu.Scene(max_instances, max_views=max_views, max_entity_index=max_entity_index)
        self.visible, self.culled = [], []

    def close(self):
        self.s.close()

    def spawn(self, entity, trs9, parent=None, aabb6=None, mesh_mat=None, flags=None):
        self.s.spawn(entity, trs9, parent, aabb6, mesh_mat, flags)

    def despawn(self, entity):
        self.s.despawn(entity)

    def set_local(self, entity, trs9):
        self.s.set_local(entity, trs9)

    def set_parent(self, entity, parent):
        self.s.set_parent(entity, parent)

    def mark_dirty(self, entity):
        self.s.mark_dirty(entity)

    def mark_all_dirty(self):
        self.s.mark_all_dirty()

    def update(self, view_projs, freeze=False, skip_transform=False):
        vps = np.ascontiguousarray(view_projs, np.float32).reshape(-1, 16)
        self.s.set_views(vps)
        flags = scgpu.UPDATE_CULLED_LISTS
        if freeze:
            flags |= scgpu.UPDATE_FREEZE_CULLING
        if skip_transform:
            flags |= scgpu.UPDATE_SKIP_TRANSFORM
        self.s.update(flags)
        self.visible = [self.s.read_visible(v) for v in range(vps.shape[0])]
        self.culled = [self.s.read_culled(v) for v in range(vps.shape[0])]
        self.recomputed = self.s.counts().recomputed

    def read_world(self, entity):
        return self.s.read_world(entity)

    def read_draw_items(self, view=0, max_draws=0):
        return self.s.read_draw_items(view, max_draws)

    def dense_entities(self):
        return self.s.read_dense_entities()


def same_bits(a, b):
    """bit equality of float arrays, with all NaNs considered equal (libm / GPU NaN payloads differ)"""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape:
        return False
    eq = a.view(np.uint32) == b.view(np.uint32)
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(eq | both_nan))


def assert_same_bits(a, b, what=""):
    """Same IEEE value in every float: identical bits, except that NaN matches any NaN (payloads differ between
    libm and the GPU) and +0 matches -0 (the structured products of scgpu_math.cuh drop x*0 terms, which can
    only change the sign of a zero entry; see DESIGN.md "Parity definition")."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    eq = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))
    if not eq.all():
        idx = np.argwhere(~eq)
        i = tuple(idx[0])
        raise AssertionError(f"{what}: {len(idx)} of {eq.size} floats differ; first at {i}: "
                             f"{a[i]!r} ({a.view(np.uint32)[i]:08x}) vs {b[i]!r} ({b.view(np.uint32)[i]:08x})")


def compare_frame(a, b, entities, n_views, what=""):
    """a, b: scene objects after update(); checks ordered visible / culled lists per view and world matrices"""
    for v in range(n_views):
        assert np.array_equal(a.visible[v], b.visible[v]), f"{what}: visible list of view {v} differs " \
            f"({len(a.visible[v])} vs {len(b.visible[v])})"
        assert np.array_equal(a.culled[v], b.culled[v]), f"{what}: culled list of view {v} differs"
    if len(entities):
        assert_same_bits(a.read_world(entities), b.read_world(entities), what + " world matrices")


def compare_draws(a, b, max_draws=0, what=""):
    da, ea, xa = a.read_draw_items(0, max_draws)
    db, eb, xb = b.read_draw_items(0, max_draws)
    assert (ea, xa) == (eb, xb), f"{what}: emitted/dropped {(ea, xa)} vs {(eb, xb)}"
    assert np.array_equal(da["entity"], db["entity"]), what
    assert np.array_equal(da["meshId"], db["meshId"]), what
    assert np.array_equal(da["materialId"], db["materialId"]), what
    assert_same_bits(da["model"], db["model"], what + " draw models")


def random_forest(rng, n, p_child=0.6, max_back=40):
    """parent index per node (-1 none): parents precede or FOLLOW children (spawn order is arbitrary in an ECS)"""
    parent = np.full(n, -1, np.int64)
    for i in range(n):
        if rng.random() < p_child and n > 1:
            lo = max(0, i - max_back)
            hi = min(n, i + max_back // 4 + 1)
            j = int(rng.integers(lo, hi))
            if j != i:
                parent[i] = j
    # break cycles deterministically so that this helper yields a forest (cycles are added explicitly by tests)
    state = np.zeros(n, np.int8)
    for i in range(n):
        path = []
        j = i
        while j >= 0 and state[j] == 0:
            state[j] = 1
            path.append(j)
            j = parent[j]
        if j >= 0 and state[j] == 1:
            parent[path[-1]] = -1
        for k in path:
            state[k] = 2
    return parent


def random_trs(rng, n, spread=200.0):
    t = np.zeros((n, 9), np.float32)
    t[:, 0:3] = rng.normal(size=(n, 3)) * spread
    t[:, 3:6] = rng.uniform(-7.0, 7.0, size=(n, 3))
    t[:, 6:9] = rng.uniform(0.2, 3.0, size=(n, 3))
    return t.astype(np.float32)


def random_aabb(rng, n):
    c = rng.normal(size=(n, 3)) * 0.5
    e = rng.uniform(0.1, 2.0, size=(n, 3))
    return np.concatenate([c - e, c + e], axis=1).astype(np.float32)


# ---- golden fixtures (tests/golden/*.npz, produced by the reference itself; see tests/golden/make_golden.py) ----

def load_golden(name):
    from pathlib import Path
    return np.load(Path(__file__).resolve().parent / "golden" / name)


def check_snapshot(scene, g, prefix, n_views, what):
    """scene: after update(); g: golden npz; compares everything the reference produced for that frame"""
    e = g[prefix + "entity"]
    for v in range(n_views):
        assert np.array_equal(scene.visible[v], g[f"{prefix}visible_{v}"]), f"{what}: visible list, view {v}"
        assert np.array_equal(scene.culled[v], g[f"{prefix}culled_{v}"]), f"{what}: culled list, view {v}"
    assert_same_bits(scene.read_world(e), g[prefix + "world"], what + " world")
    for key in g.files:
        if key.startswith(prefix + "draws_") and key.endswith("_stats"):
            md = int(key[len(prefix + "draws_"):-len("_stats")])
            items, em, dr = scene.read_draw_items(0, md)
            assert [em, dr] == list(g[key]), f"{what}: draw stats budget {md}"
            assert np.array_equal(items["entity"], g[f"{prefix}draws_{md}_entity"]), what
            assert np.array_equal(items["meshId"], g[f"{prefix}draws_{md}_mesh"]), what
            assert np.array_equal(items["materialId"], g[f"{prefix}draws_{md}_mat"]), what
            assert_same_bits(items["model"], g[f"{prefix}draws_{md}_model"], what + " draw models")


def replay_default_scene(scene, g):
    """Config 1: the sandbox's default scene. The golden holds the settled state; re-marking everything dirty
    must reproduce the reference's matrices, lists and draws."""
    scene.spawn(g["entity"], g["trs_after"], g["parent_after"], g["aabb"], g["mesh_mat"], g["flags"])
    scene.update(g["view_proj"])
    check_snapshot(scene, g, "", 1, "default scene")
    assert len(scene.visible[0]) == 166 and len(scene.culled[0]) == 553


def replay_forest_scene(scene, g):
    vps = g["view_proj"]
    scene.spawn(g["in_entity"], g["in_trs"], g["in_parent"], g["in_aabb"], g["in_mesh_mat"], g["in_flags"])
    scene.update(vps)
    check_snapshot(scene, g, "f0_", 3, "forest frame 0")
    scene.set_local(g["f1_set_entity"], g["f1_set_trs"])
    scene.set_parent(g["f1_setparent_entity"], g["f1_setparent_parent"])
    scene.update(vps)
    check_snapshot(scene, g, "f1_", 3, "forest frame 1")
    scene.despawn(g["f2_despawn"])
    scene.spawn(g["f2_spawn_entity"], g["f2_spawn_trs"], g["f2_spawn_parent"], g["f2_spawn_aabb"], None, None)
    scene.update(vps)
    check_snapshot(scene, g, "f2_", 3, "forest frame 2")
