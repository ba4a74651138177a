"""torchrun worker for tests/test_gpu_golden_and_scale.py::test_nccl_gather_across_gpus: shards a depth-4 city by
world cell, runs one scgpu context per GPU, gathers the visible lists to rank 0 over NCCL (scgpuGatherVisible) and
checks the union against the plain-C oracle on the whole scene."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))
sys.path.insert(0, str(ROOT / "tests"))

import scgpu  # noqa: E402
from scgpu import scenes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, views = 300_000, 5
    sc = scenes.city_hier(n, seed=31)
    e = np.arange(n, dtype=np.uint32)
    par = scenes.parent_handles(sc["parent"], e)
    vps = scenes.standard_views(views)
    owner = scenes.shard_by_sector(sc["sector"], world)
    mine = np.nonzero(owner == rank)[0]
    s = scgpu.Scene(len(mine) + 16, max_views=views, device=local, max_entity_index=n)
    s.spawn(e[mine], sc["trs9"][mine], par[mine], sc["aabb6"][mine], sc["mesh_mat"][mine], sc["flags"][mine])
    s.set_views(vps)
    uid = [scgpu.Scene.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(world, rank, uid[0])
    for frame in range(2):
        s.mark_all_dirty()
        s.update()
        s.gather_visible(0)
    counts = s.gathered_counts()
    local_counts = s.counts()
    assert all(counts[rank][v] == local_counts.visible[v] for v in range(views))
    if rank == 0:
        from oracle_bind import PortScene
        p = PortScene()
        p.spawn(e, sc["trs9"], par, sc["aabb6"], sc["mesh_mat"], sc["flags"])
        p.update(vps)
        for v in range(views):
            got = s.read_gathered_visible(v)
            assert len(got) == counts[:, v].sum() == len(p.visible[v]), (v, len(got), len(p.visible[v]))
            assert np.array_equal(np.sort(got), np.sort(p.visible[v])), f"view {v}"
            # shard-major stable order: rank r's slice is its own pool-order list
            off = int(counts[:rank, v].sum())
            assert np.array_equal(got[off: off + counts[0, v]], s.read_visible(v))
        nccl_lists = [s.read_gathered_visible(v) for v in range(views)]
        print("MULTIGPU OK", world, "ranks, visible per view", [int(counts[:, v].sum()) for v in range(views)], flush=True)
    dist.barrier()
    # the same gather through NVLink peer memory (no NCCL, no host sync per frame): several frames back to back so
    # that both mailbox parities and the root-progress handshake are exercised, then the lists must be identical
    s.enable_peer_gather(0)
    for frame in range(5):
        s.mark_all_dirty()
        s.update()
        s.gather_visible(0)
    if rank == 0:
        pc = s.gathered_counts()
        assert np.array_equal(pc, counts), (pc, counts)
        for v in range(views):
            assert np.array_equal(s.read_gathered_visible(v), nccl_lists[v]), f"peer gather, view {v}"
        print("MULTIGPU PEER OK", flush=True)
    else:
        try:
            s.gathered_counts()
            raise AssertionError("peer gather counts must be root-only")
        except scgpu.ScGpuError:
            pass
    dist.barrier()
    s.close()

    # ---- a frame whose lists overflow the mailbox fails THAT frame only: the error word is per gather ----
    t = scgpu.Scene(len(mine) + 16, max_views=views, device=local, max_entity_index=n)
    t.spawn(e[mine], sc["trs9"][mine], par[mine], sc["aabb6"][mine], sc["mesh_mat"][mine], sc["flags"][mine])
    uid = [scgpu.Scene.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    t.comm_init(world, rank, uid[0])
    t.enable_peer_gather(0, cap_entries=64)          # far too small for the standard views
    t.set_views(vps)
    t.update()
    t.gather_visible(0)
    if rank == 0:
        try:
            t.gathered_counts()
            raise AssertionError("an overflowing frame must be reported")
        except scgpu.ScGpuError as ex:
            assert "exceed the mailbox capacity" in str(ex), str(ex)
    t.synchronize()
    dist.barrier()
    away = scenes.standard_views(views, center=(1.0e6, 6.0, 1.0e6))   # nothing of the city is near these frusta
    t.set_views(away)
    for frame in range(3):
        t.update()
        t.gather_visible(0)
    if rank == 0:
        pc = t.gathered_counts()                         # the earlier overflow must not fail this frame
        assert pc.sum() <= 64 * world, pc
        for v in range(views):
            assert len(t.read_gathered_visible(v)) == pc[:, v].sum()
        print("MULTIGPU PEER RECOVERY OK", flush=True)
    t.synchronize()
    dist.barrier()
    t.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
