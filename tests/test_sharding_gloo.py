"""Multi-GPU host logic on CPU: world-cell sharding + gather of the per-view visible lists, world_size 2 over gloo.
Each rank runs the plain-C oracle on its shard (standing in for the per-GPU context); the concatenation in rank
order must equal the unsharded result as sets, with identical counts."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))
sys.path.insert(0, str(ROOT / "tests"))

N = 12000
VIEWS = 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    from scgpu import scenes
    sc = scenes.city_hier(N, seed=77)
    e = np.arange(N, dtype=np.uint32)
    return sc, e, scenes.parent_handles(sc["parent"], e), scenes.standard_views(VIEWS)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_bind import PortScene
    from scgpu import scenes
    sc, e, par, vps = _scene()
    owner = scenes.shard_by_sector(sc["sector"], world)
    mine = np.nonzero(owner == rank)[0]
    s = PortScene()
    s.spawn(e[mine], sc["trs9"][mine], par[mine], sc["aabb6"][mine], sc["mesh_mat"][mine], sc["flags"][mine])
    s.update(vps)
    # the exchange: counts of every rank to every rank, lists to the submitting rank (rank 0), rank order
    counts = torch.tensor([len(v) for v in s.visible], dtype=torch.int64)
    all_counts = [torch.zeros(VIEWS, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    gathered = []
    for v in range(VIEWS):
        mx = int(max(c[v] for c in all_counts))
        buf = torch.zeros(max(mx, 1), dtype=torch.int64)
        buf[: len(s.visible[v])] = torch.from_numpy(s.visible[v].astype(np.int64))
        outs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, outs, dst=0)
        if rank == 0:
            gathered.append(np.concatenate([outs[r][: int(all_counts[r][v])].numpy() for r in range(world)]).astype(np.uint32))
    if rank == 0:
        q.put(dict(gathered=gathered, counts=torch.stack(all_counts).numpy(), owned=len(mine)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_union_equals_unsharded_world_size_2():
    from oracle_bind import PortScene
    from scgpu import scenes
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc, e, par, vps = _scene()
    ref = PortScene()
    ref.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
    ref.update(vps)
    for v in range(VIEWS):
        assert res["counts"][:, v].sum() == len(ref.visible[v])
        assert np.array_equal(np.sort(res["gathered"][v]), np.sort(ref.visible[v])), f"view {v}"
    assert sum(len(x) for x in ref.visible) > 0


def test_shard_by_sector_keeps_cells_and_groups_together():
    from scgpu import scenes
    sc = scenes.city_hier(20000, seed=5)
    for world in (2, 4, 8):
        owner = scenes.shard_by_sector(sc["sector"], world)
        assert owner.min() == 0 and owner.max() == world - 1
        key = sc["sector"][:, 1].astype(np.int64) * 100003 + sc["sector"][:, 0]
        for k in np.unique(key)[:200]:
            assert len(np.unique(owner[key == k])) == 1            # a world cell never splits
        child = sc["parent"] >= 0
        assert np.array_equal(owner[child], owner[sc["parent"][child]])  # a hierarchy group lives in its root's cell
        share = np.bincount(owner, minlength=world) / len(owner)
        assert share.max() < 1.5 / world                           # balanced by instance count


# ---- churn across shards (SURVEY.md §8e "Churn"): ShardRouter replicated on every rank ---------------------------------

CHURN_FRAMES = 4


def _churn_batches():
    """Deterministic per-frame deltas, identical on every rank and in the unsharded run: whole groups despawn (plus a
    few stale and repeated handles), fresh groups spawn — half of them in cells nobody has seen, beyond the rim of the
    initial city — and random live instances get a new local TRS."""
    from scgpu import scenes
    sc, e, par, vps = _scene()
    rng = np.random.default_rng(123)
    roots = np.nonzero(sc["parent"] < 0)[0]
    ends = np.append(roots[1:], N)
    alive = np.ones(len(roots), bool)
    fresh_alive = []          # (entity handles) of spawned groups still alive
    next_index = N
    frames = []
    for f in range(CHURN_FRAMES):
        pick = rng.choice(np.nonzero(alive)[0], 60, replace=False)
        alive[pick] = False
        dead = np.concatenate([np.arange(roots[g], ends[g], dtype=np.uint32) for g in pick])
        if fresh_alive and f % 2 == 1:      # also drop a group spawned earlier
            dead = np.concatenate([dead, fresh_alive.pop(0)])
        dead = np.concatenate([dead, dead[:5], np.array([0x00FFFFF0, 0x07000001], np.uint32)])  # repeats + stale handles
        m = 400
        tmpl = scenes.city_hier(m, seed=900 + f)
        ent = np.arange(next_index, next_index + m, dtype=np.uint32)
        next_index += m
        troots = tmpl["parent"] < 0
        trs = tmpl["trs9"].copy()
        sector = tmpl["sector"].copy()
        # every second group moves out beyond the rim: cells no rank has seen yet
        gid = np.cumsum(troots) - 1
        far = (gid % 2 == 1)
        shift = np.int32(40 + 3 * f)
        trs[troots & far, 0] += np.float32(shift * 64.0)
        sector[far, 0] += shift
        spawn = dict(entity=ent, trs9=trs, parent=scenes.parent_handles(tmpl["parent"], ent), aabb6=tmpl["aabb6"],
                     mesh_mat=tmpl["mesh_mat"], flags=tmpl["flags"], sector=sector)
        tr = np.nonzero(troots)[0]
        tend = np.append(tr[1:], m)
        for a, b in list(zip(tr, tend))[:3]:
            fresh_alive.append(ent[a:b].copy())
        live = np.concatenate([np.repeat(alive, ends - roots).nonzero()[0].astype(np.uint32)] + [ent])
        moved = rng.choice(live, len(live) // 5, replace=False).astype(np.uint32)
        mtrs = np.tile(np.array([0, 0, 0, 0, 0, 0, 1, 1, 1], np.float32), (len(moved), 1))
        mtrs[:, 0:3] = rng.uniform(-300, 300, (len(moved), 3)).astype(np.float32)
        mtrs[:, 4] = rng.uniform(0, 6.28, len(moved)).astype(np.float32)
        frames.append(dict(dead=dead, spawn=spawn, moved=moved, mtrs=mtrs))
    return frames


def _apply(scene, fr, keep_dead=None, keep_spawn=None, keep_moved=None):
    d = fr["dead"] if keep_dead is None else fr["dead"][keep_dead]
    if len(d):
        scene.despawn(d)
    sp = fr["spawn"]
    k = slice(None) if keep_spawn is None else keep_spawn
    if len(sp["entity"][k]):
        scene.spawn(sp["entity"][k], sp["trs9"][k], sp["parent"][k], sp["aabb6"][k], sp["mesh_mat"][k], sp["flags"][k])
    k = slice(None) if keep_moved is None else keep_moved
    if len(fr["moved"][k]):
        scene.set_local(fr["moved"][k], fr["mtrs"][k])


def _churn_worker(rank, world, port, q):
    try:
        _churn_worker_body(rank, world, port, q)
    except BaseException as ex:   # surface the failure at once instead of letting the parent wait for its timeout
        import traceback
        q.put("rank %d: %s\n%s" % (rank, ex, traceback.format_exc()))
        raise


def _churn_worker_body(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_bind import PortScene
    from scgpu import scenes
    from scgpu.sharding import ShardRouter, NOWHERE
    sc, e, par, vps = _scene()
    owner = scenes.shard_by_sector(sc["sector"], world)
    router = ShardRouter(world, e, sc["sector"], owner, max_entity_index=1 << 16)
    mine = np.nonzero(owner == rank)[0]
    s = PortScene()
    s.spawn(e[mine], sc["trs9"][mine], par[mine], sc["aabb6"][mine], sc["mesh_mat"][mine], sc["flags"][mine])
    s.update(vps)
    per_frame = []
    for fr in _churn_batches():
        # every rank routes the whole batch (replicated router) and applies its share; the order of the three calls is
        # the unsharded one: despawn, spawn, edits
        r_dead = router.route_despawn(fr["dead"])
        r_spawn = router.route_spawn(fr["spawn"]["entity"], fr["spawn"]["sector"])
        r_moved = router.rank_of(fr["moved"])
        assert not np.any(r_moved == NOWHERE)
        _apply(s, fr, r_dead == rank, r_spawn == rank, r_moved == rank)
        s.update(vps)
        counts = torch.tensor([len(v) for v in s.visible] + [len(s.entity)], dtype=torch.int64)
        all_counts = [torch.zeros(VIEWS + 1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(all_counts, counts)
        gathered = []
        for v in range(VIEWS):
            mx = int(max(c[v] for c in all_counts))
            buf = torch.zeros(max(mx, 1), dtype=torch.int64)
            buf[: len(s.visible[v])] = torch.from_numpy(s.visible[v].astype(np.int64))
            outs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
            dist.gather(buf, outs, dst=0)
            if rank == 0:
                gathered.append(np.concatenate([outs[r][: int(all_counts[r][v])].numpy() for r in range(world)]).astype(np.uint32))
        if rank == 0:
            per_frame.append(dict(gathered=gathered, counts=torch.stack(all_counts).numpy(), router_counts=router.counts()))
    if rank == 0:
        q.put(per_frame)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_churn_routed_by_cell_equals_unsharded_world_size_2():
    """Spawns follow their root's cell (new cells join the neighbouring block), despawns and edits follow the entity:
    after every frame the union of the ranks' visible lists equals the unsharded world's, counts included, and every
    rank's instance count is what the replicated router says it is."""
    from oracle_bind import PortScene
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_churn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=500)
    if isinstance(res, str):
        for p in procs:
            p.kill()
        pytest.fail(res)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc, e, par, vps = _scene()
    ref = PortScene()
    ref.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
    ref.update(vps)
    seen_visible = 0
    for f, fr in enumerate(_churn_batches()):
        _apply(ref, fr)
        ref.update(vps)
        got = res[f]
        assert got["counts"][:, VIEWS].sum() == len(ref.entity), f"frame {f}: instances"
        assert np.array_equal(got["counts"][:, VIEWS], got["router_counts"]), f"frame {f}: router bookkeeping"
        assert got["counts"][:, VIEWS].min() > 0
        for v in range(VIEWS):
            assert got["counts"][:, v].sum() == len(ref.visible[v]), f"frame {f} view {v}: count"
            assert np.array_equal(np.sort(got["gathered"][v]), np.sort(ref.visible[v])), f"frame {f} view {v}"
            seen_visible += len(ref.visible[v])
    assert seen_visible > 0


def test_router_cells_handles_and_rim_growth():
    from scgpu import scenes
    from scgpu.sharding import ShardRouter, NOWHERE
    sc = scenes.city_hier(6000, seed=3)
    e = (np.arange(6000, dtype=np.uint32) | np.uint32(2 << 24))
    owner = scenes.shard_by_sector(sc["sector"], 4)
    r = ShardRouter(4, e, sc["sector"], owner, max_entity_index=1 << 15)
    assert np.array_equal(r.rank_of(e), owner)
    assert np.array_equal(r.counts(), np.bincount(owner, minlength=4))
    # wrong generation / never spawned -> nowhere
    assert np.all(r.rank_of(np.array([5 | (3 << 24), 7000], np.uint32)) == NOWHERE)
    # known cells keep their owner, cells beyond the rim join the outer blocks and are remembered
    assert np.array_equal(r.rank_of_cell(sc["sector"][::97]), owner[::97])
    zmax, zmin = sc["sector"][:, 1].max(), sc["sector"][:, 1].min()
    assert r.rank_of_cell(np.array([[0, zmax + 5]], np.int32))[0] == 3
    assert r.rank_of_cell(np.array([[0, zmin - 5]], np.int32))[0] == 0
    new = np.array([9000 | (1 << 24), 9001 | (1 << 24)], np.uint32)
    got = r.route_spawn(new, np.array([[2, zmax + 5], [2, zmax + 5]], np.int32))
    assert list(got) == [3, 3] and list(r.rank_of(new)) == [3, 3]
    with pytest.raises(ValueError):
        r.route_spawn(new[:1], np.array([[0, 0]], np.int32))          # index still owned
    d = r.route_despawn(np.array([new[0], new[0], e[10]], np.uint32))
    assert list(d) == [3, NOWHERE, owner[10]]
    assert r.rank_of(new[:1])[0] == NOWHERE
    # a split cell is refused
    bad = owner.copy()
    k = np.nonzero((sc["sector"] == sc["sector"][0]).all(axis=1))[0]
    if len(k) > 1:
        bad[k[0]] = (bad[k[0]] + 1) % 4
        with pytest.raises(ValueError):
            ShardRouter(4, e, sc["sector"], bad, max_entity_index=1 << 15)


def test_cpp_router_agrees_with_the_harness_router():
    """sc-gameengine_b200/host/sc_gpu_shard_router.h (what an engine process links) against scgpu/sharding.py over random
    spawn / despawn / lookup batches with known cells, unseen cells, stale, repeated and out-of-range handles."""
    import ctypes as C
    import subprocess
    from scgpu import scenes
    from scgpu.sharding import ShardRouter
    hs_dir = ROOT / "tests" / "hostsim"
    subprocess.run(["make", "-C", str(hs_dir)], check=True, capture_output=True)
    L = C.CDLL(str(hs_dir / "libhostsim.so"))
    L.hs_router_create.restype = C.c_void_p
    L.hs_router_create.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    for f in ("hs_router_rank_of_cell", "hs_router_rank_of", "hs_router_despawn"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.hs_router_spawn.restype = C.c_int
    L.hs_router_spawn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hs_router_counts.argtypes = [C.c_void_p, C.c_void_p]
    L.hs_router_destroy.argtypes = [C.c_void_p]

    world, cap = 4, 1 << 15
    sc = scenes.city_hier(8000, seed=21)
    e = np.arange(8000, dtype=np.uint32) | np.uint32(1 << 24)
    owner = scenes.shard_by_sector(sc["sector"], world)
    py = ShardRouter(world, e, sc["sector"], owner, max_entity_index=cap)
    cells, first = np.unique(sc["sector"], axis=0, return_index=True)
    cells = np.ascontiguousarray(cells, np.int32)
    cown = np.ascontiguousarray(owner[first], np.int32)
    cpp = L.hs_router_create(world, len(cells), cells.ctypes.data, cown.ctypes.data, cap)
    assert cpp
    # a split cell is refused by both
    bad = np.concatenate([cells, cells[:1]])
    bown = np.concatenate([cown, (cown[:1] + 1) % world]).astype(np.int32)
    assert not L.hs_router_create(world, len(bad), np.ascontiguousarray(bad).ctypes.data, bown.ctypes.data, cap)

    def cpp_call(fn, ent):
        ent = np.ascontiguousarray(ent, np.uint32)
        out = np.zeros(len(ent), np.int32)
        fn(cpp, len(ent), ent.ctypes.data, out.ctypes.data)
        return out

    def cpp_spawn(ent, sect):
        ent = np.ascontiguousarray(ent, np.uint32)
        sect = np.ascontiguousarray(sect, np.int32)
        out = np.zeros(len(ent), np.int32)
        ok = L.hs_router_spawn(cpp, len(ent), ent.ctypes.data, sect.ctypes.data, out.ctypes.data)
        return ok, out

    ok, got = cpp_spawn(e, sc["sector"])          # the initial scene, registered through the spawn path
    assert ok and np.array_equal(got, owner)
    rng = np.random.default_rng(8)
    live = list(e)
    next_index = 8000
    lo, hi = sc["sector"].min() - 6, sc["sector"].max() + 6
    for it in range(25):
        m = int(rng.integers(1, 300))
        ent = (np.arange(next_index, next_index + m, dtype=np.uint32) | np.uint32((it % 200) << 24))
        next_index += m
        sect = rng.integers(lo, hi, (m, 2)).astype(np.int32)      # known cells, holes inside the city, cells beyond the rim
        sect[m // 2:] = sect[m // 2]                              # several spawns share one (possibly new) cell
        want = py.route_spawn(ent, sect)
        ok, got = cpp_spawn(ent, sect)
        assert ok and np.array_equal(got, want), f"spawn batch {it}"
        live.extend(ent)
        probe = rng.integers(lo - 3, hi + 3, (200, 2)).astype(np.int32)
        out = np.zeros(200, np.int32)
        L.hs_router_rank_of_cell(cpp, 200, probe.ctypes.data, out.ctypes.data)
        assert np.array_equal(out, py.rank_of_cell(probe)), f"cells after batch {it}"
        k = int(rng.integers(1, 400))
        pick = rng.choice(len(live), min(k, len(live)), replace=False)
        victims = np.array([live[i] for i in pick], np.uint32)
        noise = np.array([0x00FFFFF0, 7 | (250 << 24), int(victims[0]), 40000], np.uint32)
        batch = np.concatenate([victims, noise])
        rng.shuffle(batch)
        assert np.array_equal(cpp_call(L.hs_router_rank_of, batch), py.rank_of(batch)), f"lookup {it}"
        assert np.array_equal(cpp_call(L.hs_router_despawn, batch), py.route_despawn(batch)), f"despawn {it}"
        gone = set(int(v) for v in victims)
        live = [h for h in live if int(h) not in gone]
        cnt = np.zeros(world, np.uint64)
        L.hs_router_counts(cpp, cnt.ctypes.data)
        assert np.array_equal(cnt.astype(np.int64), py.counts())
    # a spawn onto a live index or a repeat inside the batch fails and changes nothing
    before = cpp_call(L.hs_router_rank_of, np.array(live[:50], np.uint32))
    ok, _ = cpp_spawn(np.array([live[0], 30000], np.uint32), np.zeros((2, 2), np.int32))
    assert not ok
    ok, _ = cpp_spawn(np.array([30001, 30001], np.uint32), np.zeros((2, 2), np.int32))
    assert not ok
    ok, _ = cpp_spawn(np.array([30002, cap + 5], np.uint32), np.zeros((2, 2), np.int32))
    assert not ok
    assert np.array_equal(cpp_call(L.hs_router_rank_of, np.array(live[:50], np.uint32)), before)
    assert list(cpp_call(L.hs_router_rank_of, np.array([30000, 30001, 30002], np.uint32))) == [-1, -1, -1]
    L.hs_router_destroy(cpp)
