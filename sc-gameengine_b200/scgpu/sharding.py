"""Host-side routing of ECS deltas to the GPUs of a cell-sharded world ("Churn").

The instance set is sharded by world cell — `SectorCoord`, 64 m, the key the reference tags every streamed entity with
(`WorldSector`, /root/reference/src/engine/world/sc_world_partition.h:292-296; `worldToSector`, .cpp:268-275) — in
contiguous blocks of cells in row-major (z, x) order, balanced by instance count (`scenes.shard_by_sector`). There is no
collective on the data path, so the only multi-GPU logic a frame needs besides the gather of the visible lists is this:
every rank looks at the same batch of spawns / despawns / TRS edits and keeps its own share.

`ShardRouter` is replicated, not distributed: every rank builds it from the same initial scene and feeds it the same
batches, so all ranks agree on every owner without exchanging a byte.
  * a spawn goes to the rank that owns its ROOT's cell (a hierarchy group lives in its root's cell, so parent links
    never cross GPUs); a cell nobody has seen yet joins the block of the nearest preceding known cell in (z, x) order —
    blocks stay contiguous, and a world that grows at its rim extends the outer blocks;
  * despawns and edits go to the rank the entity was spawned on (entity index -> rank table, cleared on despawn; a
    stale handle routes nowhere, like `World::destroy` returning false).
It is plain numpy on the host: O(batch) per call plus one `searchsorted` for cells.
"""
from __future__ import annotations

import numpy as np

INDEX_MASK = 0xFFFFFF  # 24-bit entity index, /root/reference/src/core/include/sc_ecs.h:18-20
NOWHERE = -1


def cell_key(sector):
    """(z, x) row-major sort key of a SectorCoord array [n, 2] = (x, z)."""
    sector = np.asarray(sector)
    return (sector[:, 1].astype(np.int64) << 32) + (sector[:, 0].astype(np.int64) & 0xFFFFFFFF)


class ShardRouter:
    def __init__(self, n_ranks, entity, sector, owner, max_entity_index=1 << 24):
        """entity [n] u32 handles, sector [n, 2] i32 cell of each entity's ROOT, owner [n] rank of each entity
        (scenes.shard_by_sector). Entities of one cell must share an owner."""
        self.n_ranks = int(n_ranks)
        entity = np.asarray(entity, np.uint32)
        owner = np.asarray(owner, np.int32)
        key = cell_key(sector)
        uniq, first = np.unique(key, return_index=True)
        cell_owner = owner[first]
        # one owner per cell, or the shard map is not a cell map
        check = cell_owner[np.searchsorted(uniq, key)]
        if not np.array_equal(check, owner):
            raise ValueError("ShardRouter: a world cell is split across ranks")
        self._keys = uniq
        self._cell_owner = cell_owner.astype(np.int32)
        self._rank_of_index = np.full(int(max_entity_index), NOWHERE, np.int8 if n_ranks < 128 else np.int32)
        self._handle = np.zeros(int(max_entity_index), np.uint32)
        idx = entity & np.uint32(INDEX_MASK)
        self._rank_of_index[idx] = owner
        self._handle[idx] = entity

    # ---- cells ---------------------------------------------------------------------------------------------------
    def rank_of_cell(self, sector):
        """Owner of each cell; unknown cells take the owner of the nearest preceding known cell in (z, x) order (the
        first block before the first known cell). Pure function of the map: it does not learn the cell."""
        key = cell_key(sector)
        pos = np.searchsorted(self._keys, key, side="right") - 1
        return self._cell_owner[np.maximum(pos, 0)]

    def add_cells(self, sector):
        """Pins unknown cells to the owner rank_of_cell gives them now, so that later neighbours are placed relative
        to them. Called by route_spawn; idempotent for known cells."""
        key = np.unique(cell_key(sector))
        new = key[~np.isin(key, self._keys)]
        if len(new) == 0:
            return
        pos = np.searchsorted(self._keys, new, side="right") - 1
        own = self._cell_owner[np.maximum(pos, 0)]
        keys = np.concatenate([self._keys, new])
        owners = np.concatenate([self._cell_owner, own])
        order = np.argsort(keys, kind="stable")
        self._keys, self._cell_owner = keys[order], owners[order]

    # ---- entities ------------------------------------------------------------------------------------------------
    def route_spawn(self, entity, root_sector):
        """Registers new entities; returns their rank. root_sector [n, 2]: the cell of each entity's hierarchy ROOT
        (for a root: its own cell). Raises if an index is still owned (World::create never hands out a live index)."""
        entity = np.asarray(entity, np.uint32)
        idx = (entity & np.uint32(INDEX_MASK)).astype(np.int64)
        if len(idx) and idx.max() >= len(self._rank_of_index):
            raise ValueError("ShardRouter.route_spawn: entity index beyond max_entity_index")
        if len(np.unique(idx)) != len(idx) or np.any(self._rank_of_index[idx] != NOWHERE):
            raise ValueError("ShardRouter.route_spawn: an entity index is already owned")
        rank = self.rank_of_cell(root_sector)
        self.add_cells(root_sector)
        self._rank_of_index[idx] = rank
        self._handle[idx] = entity
        return rank.astype(np.int32)

    def rank_of(self, entity):
        """Rank each handle lives on, NOWHERE for stale / unknown handles (wrong generation included)."""
        entity = np.asarray(entity, np.uint32)
        idx = (entity & np.uint32(INDEX_MASK)).astype(np.int64)
        inside = idx < len(self._rank_of_index)   # an index beyond max_entity_index was never handed out
        safe = np.where(inside, idx, 0)
        rank = self._rank_of_index[safe].astype(np.int32)
        rank[~inside | (self._handle[safe] != entity)] = NOWHERE
        return rank

    def route_despawn(self, entity):
        """Rank of each handle (NOWHERE for stale ones and for repeats after the first), and forgets the entities."""
        entity = np.asarray(entity, np.uint32)
        rank = self.rank_of(entity)
        idx = entity & np.uint32(INDEX_MASK)
        # a handle repeated in the batch is stale the second time
        _, first = np.unique(entity, return_index=True)
        repeat = np.ones(len(entity), bool)
        repeat[first] = False
        rank[repeat] = NOWHERE
        live = rank != NOWHERE
        self._rank_of_index[idx[live]] = NOWHERE
        return rank

    def counts(self):
        """Instances per rank."""
        r = self._rank_of_index[self._rank_of_index != NOWHERE]
        return np.bincount(r.astype(np.int64), minlength=self.n_ranks)
