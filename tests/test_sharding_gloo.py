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
