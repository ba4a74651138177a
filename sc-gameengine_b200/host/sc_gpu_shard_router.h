// sc_gpu_shard_router.h — which GPU owns an entity of a cell-sharded world ("Churn").
//
// The instance set is sharded by world cell: sc::SectorCoord, 64 m, the key the reference tags every streamed entity
// with (WorldSector, src/engine/world/sc_world_partition.h:292-296; worldToSector, .cpp:268-275), in contiguous blocks of
// cells in row-major (z, x) order. One process per GPU, one scgpu context per process, no collective on the data path:
// the only multi-GPU logic a frame needs besides the gather of the visible lists is that every process looks at the
// same batch of spawns / despawns / edits and keeps its own share. The router is therefore REPLICATED — every process
// builds it from the same cell map and feeds it the same batches, so all agree on every owner without exchanging a byte.
//
//   * a spawn goes to the rank that owns its hierarchy ROOT's cell (parent links never cross GPUs); a cell nobody has
//     seen yet joins the block of the nearest preceding known cell in (z, x) order, so blocks stay contiguous while the
//     world grows at its rim;
//   * despawns and edits go to the rank the entity was spawned on (entity index -> rank table, cleared on despawn); a
//     stale, repeated or out-of-range handle routes nowhere, like World::destroy returning false (sc_ecs.cpp:33-45).
//
// Header only, no engine headers (coordinates come as int32 pairs, handles as uint32), so the CPU suite compiles it
// through tests/hostsim and checks it against scgpu/sharding.py, the harness's numpy form of the same rules.
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

namespace sc::gpu
{
  class ShardRouter
  {
  public:
    static constexpr int32_t kNowhere = -1;
    static constexpr uint32_t kIndexMask = 0xFFFFFFu;  // 24-bit entity index (src/core/include/sc_ecs.h:18-20)

    // cellsXZ: nCells (x, z) pairs, all distinct; owner: the rank of each. maxEntityIndex: size of the handle table.
    // Returns false (router left empty) when a cell appears twice with different owners or an owner is out of range.
    bool init(uint32_t nRanks, uint32_t nCells, const int32_t* cellsXZ, const int32_t* owner, uint32_t maxEntityIndex)
    {
      m_ranks = nRanks;
      m_cells.clear();
      m_rankOfIndex.assign(maxEntityIndex, (int16_t)kNowhere);
      m_handle.assign(maxEntityIndex, 0u);
      m_cells.reserve(nCells);
      for (uint32_t i = 0; i < nCells; ++i)
      {
        if (owner[i] < 0 || (uint32_t)owner[i] >= nRanks) { m_cells.clear(); return false; }
        m_cells.push_back(Cell{key(cellsXZ[2 * i], cellsXZ[2 * i + 1]), owner[i]});
      }
      std::sort(m_cells.begin(), m_cells.end(), [](const Cell& a, const Cell& b) { return a.key < b.key || (a.key == b.key && a.owner < b.owner); });
      size_t w = 0;
      for (size_t i = 0; i < m_cells.size(); ++i)
      {
        if (w && m_cells[w - 1].key == m_cells[i].key)
        {
          if (m_cells[w - 1].owner != m_cells[i].owner) { m_cells.clear(); return false; }  // a split cell
          continue;
        }
        m_cells[w++] = m_cells[i];
      }
      m_cells.resize(w);
      return !m_cells.empty();
    }

    uint32_t ranks() const { return m_ranks; }

    // Owner of a cell; an unknown cell takes the owner of the nearest preceding known cell in (z, x) order (the first
    // block before the first known cell). Does not learn the cell.
    int32_t rankOfCell(int32_t x, int32_t z) const
    {
      if (m_cells.empty()) return kNowhere;
      const int64_t k = key(x, z);
      auto it = std::upper_bound(m_cells.begin(), m_cells.end(), k, [](int64_t v, const Cell& c) { return v < c.key; });
      return it == m_cells.begin() ? m_cells.front().owner : (it - 1)->owner;
    }

    // Registers n new entities and writes their ranks. rootCellXZ: the cell of each entity's hierarchy root (for a
    // root: its own). All owners are decided against the map as it was BEFORE the batch, then the batch's unknown cells
    // are pinned. Returns false and changes nothing if an index is out of range, repeated, or still owned
    // (World::create never hands out a live index).
    bool routeSpawn(uint32_t n, const uint32_t* entity, const int32_t* rootCellXZ, int32_t* outRank)
    {
      for (uint32_t i = 0; i < n; ++i)
      {
        const uint32_t idx = entity[i] & kIndexMask;
        if (idx >= m_rankOfIndex.size() || m_rankOfIndex[idx] != kNowhere)
        {
          for (uint32_t j = 0; j < i; ++j) m_rankOfIndex[entity[j] & kIndexMask] = (int16_t)kNowhere;
          return false;
        }
        m_rankOfIndex[idx] = (int16_t)m_ranks;  // provisional mark: catches a repeat inside the batch
      }
      std::vector<Cell> fresh;
      for (uint32_t i = 0; i < n; ++i)
      {
        const int32_t x = rootCellXZ[2 * i], z = rootCellXZ[2 * i + 1];
        const int32_t r = rankOfCell(x, z);
        outRank[i] = r;
        const uint32_t idx = entity[i] & kIndexMask;
        m_rankOfIndex[idx] = (int16_t)r;
        m_handle[idx] = entity[i];
        const int64_t k = key(x, z);
        auto it = std::lower_bound(m_cells.begin(), m_cells.end(), k, [](const Cell& c, int64_t v) { return c.key < v; });
        if (it == m_cells.end() || it->key != k) fresh.push_back(Cell{k, r});
      }
      if (!fresh.empty())
      {
        std::sort(fresh.begin(), fresh.end(), [](const Cell& a, const Cell& b) { return a.key < b.key; });
        fresh.erase(std::unique(fresh.begin(), fresh.end(), [](const Cell& a, const Cell& b) { return a.key == b.key; }), fresh.end());
        const size_t mid = m_cells.size();
        m_cells.insert(m_cells.end(), fresh.begin(), fresh.end());
        std::inplace_merge(m_cells.begin(), m_cells.begin() + (std::ptrdiff_t)mid, m_cells.end(),
                           [](const Cell& a, const Cell& b) { return a.key < b.key; });
      }
      return true;
    }

    // Rank a handle lives on; kNowhere for stale / unknown handles (wrong generation included).
    int32_t rankOf(uint32_t entity) const
    {
      const uint32_t idx = entity & kIndexMask;
      if (idx >= m_rankOfIndex.size() || m_handle[idx] != entity) return kNowhere;
      return m_rankOfIndex[idx];
    }
    void rankOf(uint32_t n, const uint32_t* entity, int32_t* outRank) const
    {
      for (uint32_t i = 0; i < n; ++i) outRank[i] = rankOf(entity[i]);
    }

    // Ranks of a despawn batch (a handle repeated in the batch is stale the second time), and forgets the entities.
    void routeDespawn(uint32_t n, const uint32_t* entity, int32_t* outRank)
    {
      for (uint32_t i = 0; i < n; ++i)
      {
        const int32_t r = rankOf(entity[i]);
        outRank[i] = r;
        if (r != kNowhere) m_rankOfIndex[entity[i] & kIndexMask] = (int16_t)kNowhere;
      }
    }

    // Instances per rank (out: ranks() counters).
    void counts(uint64_t* out) const
    {
      for (uint32_t r = 0; r < m_ranks; ++r) out[r] = 0;
      for (int16_t r : m_rankOfIndex) if (r >= 0 && (uint32_t)r < m_ranks) ++out[r];
    }

  private:
    struct Cell
    {
      int64_t key;
      int32_t owner;
    };
    static int64_t key(int32_t x, int32_t z) { return ((int64_t)z << 32) + (int64_t)(uint32_t)x; }

    uint32_t m_ranks = 0;
    std::vector<Cell> m_cells;           // sorted by key
    std::vector<int16_t> m_rankOfIndex;  // entity index -> rank, kNowhere = not alive
    std::vector<uint32_t> m_handle;      // entity index -> full handle (generation check)
  };
}
