// scgpu_layout.h — where a Transform lives in HBM, decoupled from where it stands in the reference's pool.
//
// The reference's ComponentPool<Transform> (src/core/include/sc_ecs.h:199-277) defines an ORDER (dense index = "rank":
// append on add, swap-with-last on remove, sc_ecs.h:240-262) and every list the path emits follows it. It does not
// have to define the storage order of the device arrays, and under streaming churn it must not: swap-with-last drags
// members of tail groups into holes all over the pool, so after a few dozen frames a third of the children sit more than
// a warp away from their parent (tools/model_churn_order.py) and the hierarchy windows of k_update_win fall apart.
//
// So a Transform gets a device SLOT when it is spawned and keeps it until it is despawned; only its rank changes
// (rank[slot], perm[rank] on the device, maintained by the same `dst <- src` moves the pool mirror derives). Slots are
// handed out per hierarchy GROUP — a run of a spawn batch in which every element's parent is an earlier element of the
// same run — so that parent and children stay inside one window for as long as they live, and the holes that whole-group
// despawns leave are reused by later groups of the same size. The visible lists are put back into rank order by the
// compaction kernel (a bitmap indexed by rank), which is ~1 % of the scene.
//
// Pure host C++ (no CUDA types): compiled into libscgpu.so and, for the CPU suite, into tests/hostsim.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace scgpu
{

constexpr uint32_t kLayoutMaxGroup = 32;  // a hierarchy window of k_update_win holds at most 32 slots

class SlotLayout
{
public:
  void reset(uint32_t capacity)
  {
    m_capacity = capacity;
    m_extent = 0;
    m_freeSlots = 0;
    for (auto& b : m_bucket) b.clear();
    m_big.clear();
  }
  uint32_t extent() const { return m_extent; }        // slots [0, extent) are in use or free holes
  uint32_t freeSlots() const { return m_freeSlots; }  // holes below the extent
  uint32_t capacity() const { return m_capacity; }
  bool hasHoles() const { return m_freeSlots != 0; }

  // n new slots at the end of the extent (no holes are looked at); returns the first one, or UINT32_MAX
  uint32_t appendRun(uint32_t n)
  {
    if ((uint64_t)m_extent + n > m_capacity) return 0xFFFFFFFFu;
    const uint32_t s = m_extent;
    m_extent += n;
    return s;
  }

  // g <= kLayoutMaxGroup consecutive slots: a hole of exactly that size, else the front of the smallest larger hole,
  // else fresh slots at the end. UINT32_MAX when nothing contiguous of that size is left.
  uint32_t allocGroup(uint32_t g)
  {
    if (g == 0 || g > kLayoutMaxGroup) return 0xFFFFFFFFu;
    for (uint32_t len = g; len <= kLayoutMaxGroup; ++len)
    {
      std::vector<uint32_t>& b = m_bucket[len];
      if (b.empty()) continue;
      const uint32_t s = b.back();
      b.pop_back();
      m_freeSlots -= g;
      if (len > g) m_bucket[len - g].push_back(s + g);
      return s;
    }
    if (!m_big.empty())
    {
      std::pair<uint32_t, uint32_t>& r = m_big.back();
      const uint32_t s = r.first;
      r.first += g;
      r.second -= g;
      m_freeSlots -= g;
      if (r.second <= kLayoutMaxGroup)
      {
        const std::pair<uint32_t, uint32_t> rest = r;
        m_big.pop_back();
        if (rest.second) m_bucket[rest.second].push_back(rest.first);
      }
      return s;
    }
    return appendRun(g);
  }

  // Assigns a slot to every element of a spawn batch. entity / parent: the batch as scgpuSpawn receives it (parent may
  // be null). A group is cut where an element's parent is not among the elements of the current group (linear look
  // back over at most kLayoutMaxGroup handles: in practice the parent is one to four positions back), and placed as a
  // whole. If nothing contiguous is left for a group its elements are placed one by one — always possible while the
  // pool itself has room — and resolve their parent through the generic path. Returns false when slots run out
  // (nothing is allocated then).
  bool placeBatch(uint32_t n, const uint32_t* entity, const uint32_t* parent, uint32_t* slotOut)
  {
    if ((uint64_t)(m_extent - m_freeSlots) + n > m_capacity) return false;
    uint32_t j = 0;
    while (j < n)
    {
      uint32_t end = j + 1;
      if (parent)
      {
        while (end < n && end - j < kLayoutMaxGroup)
        {
          const uint32_t ph = parent[end];
          if (ph == 0xFFFFFFFFu) break;
          bool inside = false;
          for (uint32_t k = end; k-- > j;)
            if (entity[k] == ph) { inside = true; break; }
          if (!inside) break;
          ++end;
        }
      }
      const uint32_t g = end - j;
      const uint32_t s = allocGroup(g);
      if (s != 0xFFFFFFFFu)
      {
        for (uint32_t k = 0; k < g; ++k) slotOut[j + k] = s + k;
      }
      else
      {
        for (uint32_t k = 0; k < g; ++k) slotOut[j + k] = allocGroup(1u);  // cannot fail: free + tail room >= n was checked
      }
      j = end;
    }
    return true;
  }

  // Slots of despawned Transforms, in batch order. Consecutive ascending slots (a group destroyed as a whole) become one
  // hole; a run that ends at the extent shrinks the extent instead.
  void release(uint32_t n, const uint32_t* slot)
  {
    uint32_t j = 0;
    while (j < n)
    {
      uint32_t end = j + 1;
      while (end < n && slot[end] == slot[end - 1] + 1u) ++end;
      addHole(slot[j], end - j);
      j = end;
    }
  }

  // every hole as (start, length), unordered (tests)
  void holes(std::vector<std::pair<uint32_t, uint32_t>>& out) const
  {
    out.clear();
    for (uint32_t len = 1; len <= kLayoutMaxGroup; ++len)
      for (uint32_t s : m_bucket[len]) out.emplace_back(s, len);
    for (const auto& r : m_big) out.push_back(r);
  }

private:
  void addHole(uint32_t start, uint32_t len)
  {
    if (start + len == m_extent) { m_extent = start; return; }
    m_freeSlots += len;
    if (len <= kLayoutMaxGroup) m_bucket[len].push_back(start);
    else m_big.emplace_back(start, len);
  }

  uint32_t m_capacity = 0, m_extent = 0, m_freeSlots = 0;
  std::vector<uint32_t> m_bucket[kLayoutMaxGroup + 1];   // [len]: starts of the holes of exactly that length
  std::vector<std::pair<uint32_t, uint32_t>> m_big;      // holes longer than a window
};

}  // namespace scgpu
