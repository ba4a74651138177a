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
// same run — so that parent and children stay inside one window for as long as they live. The visible lists are put
// back into rank order by the compaction kernel (a bitmap indexed by rank), which is ~1 % of the scene.
//
// The allocator is a free-slot BITMAP walked by a roving cursor (next fit): a spawn batch takes the free runs it meets
// in address order, each group the first run from the cursor on that holds it. What spawns together therefore lands
// close together — a world sector that streams in fills the holes a sector that streamed out left, one after the other
// — and that SPATIAL coherence of neighbouring slots is what the warp-wide early-outs of the culling code live on (a
// warp whose 32 instances come from all over the city has somebody near a frustum most of the time). Adjacent holes
// coalesce by construction (they are adjacent bits), fragments too small for a group are passed over and taken by a
// smaller group or a later lap.
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
    for (uint32_t& c : m_cursor) c = 0;
    m_binCursor = 0;
    m_free.assign(((size_t)capacity + 63u) / 64u + 1u, 0ull);
    for (uint32_t& f : m_noRunSince) f = 0;
    for (uint32_t& f : m_noExactSince) f = 0;
    m_epoch = 1;
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

  // g <= kLayoutMaxGroup consecutive slots. Every group size has its own roving cursor and sweeps the bitmap forward
  // (one lap around the extent at most), first for a run of EXACTLY g slots or a long run (>= a window: a sector-sized
  // hole, carved group after group), then — when a whole lap holds no such run — for any run that is large enough;
  // else fresh slots at the end. Preferring the exact fit matters when holes and groups come in a few sizes (vehicles
  // of 10, peds of 4): plain first fit lets the small groups eat the large holes and leaves the large groups crumbs.
  // A hole is looked at once per lap and size class, so a batch costs O(holes it passes), not O(groups x look-ahead).
  // UINT32_MAX when nothing contiguous of that size is left.
  uint32_t allocGroup(uint32_t g)
  {
    if (g == 0 || g > kLayoutMaxGroup) return 0xFFFFFFFFu;
    uint32_t s = allocInHoles(g);
    if (s != 0xFFFFFFFFu) return s;
    s = appendRun(g);
    if (s != 0xFFFFFFFFu) m_cursor[g] = m_extent;
    return s;
  }

  // the hole half of allocGroup: UINT32_MAX when no hole below the extent holds g consecutive slots
  uint32_t allocInHoles(uint32_t g)
  {
    if (g == 0 || g > kLayoutMaxGroup) return 0xFFFFFFFFu;
    if (m_freeSlots >= g)
    {
      if (m_noExactSince[g] != m_epoch)
      {
        const uint32_t s = sweep(g, true);
        if (s != 0xFFFFFFFFu) return s;
        m_noExactSince[g] = m_epoch;
      }
      if (m_noRunSince[g] != m_epoch)
      {
        const uint32_t s = sweep(g, false);
        if (s != 0xFFFFFFFFu) return s;
        // no run of g anywhere: do not lap again for this size (or a larger one) until something is released
        for (uint32_t k = g; k <= kLayoutMaxGroup; ++k) m_noRunSince[k] = m_epoch;
      }
    }
    return 0xFFFFFFFFu;
  }

  // true when the batch holds at least one parent link, i.e. when placeBatch has groups to keep together and to pack
  static bool batchHasHierarchy(uint32_t n, const uint32_t* parent)
  {
    if (!parent) return false;
    for (uint32_t j = 0; j < n; ++j)
      if (parent[j] != 0xFFFFFFFFu) return true;
    return false;
  }

  // Assigns a slot to every element of a spawn batch. entity / parent: the batch as scgpuSpawn receives it (parent may
  // be null). A group is cut where an element's parent is not among the elements of the current group (linear look
  // back over at most kLayoutMaxGroup handles: in practice the parent is one to four positions back), and placed as a
  // whole. If nothing contiguous is left for a group its elements are placed one by one — always possible while the
  // pool itself has room — and resolve their parent through the generic path. Returns false when slots run out
  // (nothing is allocated then).
  //
  // WINDOW PACKING. k_update_win gives a warp to every window of <= 32 consecutive slots that no parent link crosses,
  // and its cost is per window, not per instance. Groups placed in arrival order leave windows 89 % full when vehicles
  // of 10 and peds of 4 alternate at random (10 + 10 + 4 + 4 = 28, the next 10 does not fit). So the groups of a batch
  // are first packed into BINS of at most 32 slots and a bin is placed as a whole: the oldest group not yet placed
  // opens a bin, a bounded knapsack over the sizes that are waiting (128 slots' worth of look-ahead, so what spawns
  // together still lands together) fills the rest - 10 + 10 + 4 + 4 + 4 - and the window cut of k_build_windows finds
  // its boundaries at multiples of 32: 98 % full windows, 9 % fewer of them. The device never sees any of this: slots
  // are its storage order only, every list it emits is in rank (pool) order.
  bool placeBatch(uint32_t n, const uint32_t* entity, const uint32_t* parent, uint32_t* slotOut)
  {
    if ((uint64_t)(m_extent - m_freeSlots) + n > m_capacity) return false;
    if (!parent)
    {
      for (uint32_t j = 0; j < n; ++j) slotOut[j] = allocGroup(1u);  // cannot fail: free + tail room >= n was checked
      return true;
    }
    m_gStart.clear();
    m_gSize.clear();
    m_gTaken.clear();
    for (auto& q : m_queue) { q.clear(); }
    for (uint32_t& h : m_qHead) h = 0;
    m_sizeMask = 0;
    m_oldest = 0;
    m_pendingSlots = 0;
    beginRuns();
    uint32_t j = 0;
    while (j < n)
    {
      uint32_t end = j + 1;
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
      const uint32_t g = end - j;
      m_queue[g].push_back((uint32_t)m_gStart.size());
      m_sizeMask |= 1ull << g;
      m_gStart.push_back(j);
      m_gSize.push_back((uint8_t)g);
      m_gTaken.push_back(0);
      m_pendingSlots += g;
      if (m_pendingSlots >= kPackKeep) packPending(false, slotOut);
      j = end;
    }
    packPending(true, slotOut);
    return true;
  }

  // Slots of despawned Transforms, in batch order. Free slots at the very end of the extent give it back.
  void release(uint32_t n, const uint32_t* slot)
  {
    if (n == 0) return;
    for (uint32_t j = 0; j < n; ++j)
    {
      const uint32_t s = slot[j];
      m_free[s >> 6] |= 1ull << (s & 63u);
    }
    m_freeSlots += n;
    ++m_epoch;
    if (m_epoch == 0) { ++m_epoch; for (uint32_t& f : m_noRunSince) f = 0; for (uint32_t& f : m_noExactSince) f = 0; }
    while (m_extent > 0u && ((m_free[(m_extent - 1u) >> 6] >> ((m_extent - 1u) & 63u)) & 1ull))
    {
      // whole trailing words at once where possible
      const uint32_t last = m_extent - 1u;
      if ((last & 63u) == 63u && m_free[last >> 6] == ~0ull) { m_free[last >> 6] = 0ull; m_extent -= 64u; m_freeSlots -= 64u; continue; }
      m_free[last >> 6] &= ~(1ull << (last & 63u));
      --m_extent;
      --m_freeSlots;
    }
  }

  // every hole as (start, length), ascending (tests)
  void holes(std::vector<std::pair<uint32_t, uint32_t>>& out) const
  {
    out.clear();
    uint32_t pos = 0;
    for (;;)
    {
      const uint32_t f = nextFree(pos, m_extent);
      if (f >= m_extent) break;
      uint32_t e = f;
      while (e < m_extent && ((m_free[e >> 6] >> (e & 63u)) & 1ull)) ++e;
      out.emplace_back(f, e - f);
      pos = e;
    }
  }

private:
  // Slots' worth of groups that wait for a bin (the look-ahead of the packing). 128 is enough for vehicles of 10 and
  // peds of 4 in random order to fill windows to 31.5 of 32 slots (64: 31.3, 32: 30.4, unbounded: 31.5), and small
  // against a world sector (~1000 slots): a sector that streams out still leaves ONE hole, not crumbs among its
  // neighbours - a first version with a look-ahead of a thousand groups leaked 2 % of the pool in eight churn frames.
  static constexpr uint32_t kPackKeep = 4 * kLayoutMaxGroup;

  // One group of size g placed on its own (the packing found no room for a whole bin): as a run, else slot by slot.
  void placeGroup(uint32_t gi, uint32_t* slotOut)
  {
    const uint32_t j = m_gStart[gi], g = m_gSize[gi];
    const uint32_t s = allocGroup(g);
    for (uint32_t k = 0; k < g; ++k) slotOut[j + k] = s != 0xFFFFFFFFu ? s + k : allocGroup(1u);
  }

  uint32_t popQueue(uint32_t g)
  {
    const uint32_t gi = m_queue[g][m_qHead[g]++];
    if (m_qHead[g] == m_queue[g].size()) { m_queue[g].clear(); m_qHead[g] = 0; m_sizeMask &= ~(1ull << g); }
    m_gTaken[gi] = 1;
    m_pendingSlots -= g;
    return gi;
  }
  uint32_t waiting(uint32_t g) const { return (uint32_t)m_queue[g].size() - m_qHead[g]; }

  // ---- where the bins go: the free runs of the pool in address order, one lap from a roving cursor, then the tail ----
  // A run is carved bin after bin; what is left of it when nothing that waits fits any more stays a hole for a later
  // batch. One walk of the bitmap per batch, whatever the number of bins.
  void beginRuns()
  {
    m_tailMode = !hasHoles();
    m_runPos = m_binCursor < m_extent ? m_binCursor : 0u;
    m_lapStart = m_runPos;
    m_wrapped = false;
    m_runAt = 0;
    m_runLeft = 0;
  }
  // the next free run into (m_runAt, m_runLeft); false when neither holes nor tail room are left
  bool nextRun()
  {
    while (!m_tailMode)
    {
      const uint32_t limit = m_wrapped ? m_lapStart : m_extent;
      const uint32_t f = nextFree(m_runPos, limit);
      if (f >= limit)
      {
        if (m_wrapped || m_lapStart == 0u) { m_tailMode = true; break; }
        m_wrapped = true;
        m_runPos = 0u;
        continue;
      }
      const uint32_t e = runEnd(f, m_extent);
      m_runAt = f;
      m_runLeft = e - f;
      m_runPos = e;
      return true;
    }
    m_runAt = m_extent;
    m_runLeft = m_capacity - m_extent;
    return m_runLeft != 0u;
  }

  // bit s set: some of the waiting groups (one of size `less` left aside, if any) add up to exactly s, s < 64
  uint64_t reachable(uint32_t less) const
  {
    uint64_t reach = 1ull;
    for (uint64_t m = m_sizeMask; m; m &= m - 1ull)  // the sizes that wait, ascending
    {
      const uint32_t g = (uint32_t)__builtin_ctzll(m);
      uint32_t c = waiting(g);
      if (g == less && c) --c;
      if (c > 63u / g) c = 63u / g;
      for (uint32_t k = 0; k < c; ++k) reach |= reach << g;
    }
    return reach;
  }

  // Largest sum <= cap of waiting group sizes (bounded knapsack; reachBefore[i] = sums reachable before copy i); the
  // groups that make it up are popped into bin[nb..). Returns the sum.
  uint32_t fillFromWaiting(uint32_t cap, uint32_t* bin, uint32_t& nb)
  {
    if (cap == 0u) return 0u;
    uint8_t copySize[160];
    uint64_t reachBefore[160];
    uint32_t copies = 0;
    uint64_t reach = 1ull;
    for (uint64_t m = m_sizeMask & ((2ull << cap) - 1ull); m; m &= m - 1ull)  // the sizes that wait and fit, ascending
    {
      const uint32_t g = (uint32_t)__builtin_ctzll(m);
      const uint32_t w = waiting(g);
      const uint32_t c = w < cap / g ? w : cap / g;
      for (uint32_t k = 0; k < c && copies < 160u; ++k)
      {
        copySize[copies] = (uint8_t)g;
        reachBefore[copies] = reach;
        reach |= reach << g;
        ++copies;
      }
    }
    reach &= (2ull << cap) - 1ull;
    uint32_t best = 63u - (uint32_t)__builtin_clzll(reach);
    const uint32_t sum = best;
    // walk the copies backwards (largest sizes first): a copy is taken when the sum is not reachable without it
    for (uint32_t i = copies; i-- > 0u && best != 0u;)
    {
      if ((reachBefore[i] >> best) & 1ull) continue;
      bin[nb++] = popQueue(copySize[i]);
      best -= copySize[i];
    }
    return sum;
  }

  // Bins out of the waiting groups until fewer than kPackKeep slots wait (or none, when `final`). A bin is at most a
  // window and at most what is left of the current free run; the oldest waiting group goes first whenever it fits.
  void packPending(bool final, uint32_t* slotOut)
  {
    uint32_t bin[kLayoutMaxGroup];
    for (;;)
    {
      while (m_oldest < m_gTaken.size() && m_gTaken[m_oldest]) ++m_oldest;
      if (m_oldest >= m_gTaken.size()) break;
      if (!final && m_pendingSlots < kPackKeep) break;
      if (m_runLeft == 0u && !nextRun())
      {
        // no contiguous room anywhere (a pool that is all but full): group by group, then slot by slot. (Only in tail
        // mode, i.e. when this walk holds no hole of its own: allocGroup takes its slots from the bitmap directly.)
        placeGroup(popQueue(m_gSize[m_oldest]), slotOut);
        continue;
      }
      uint32_t cap = m_runLeft < kLayoutMaxGroup ? m_runLeft : kLayoutMaxGroup;
      bool oldestFirst = true;
      if (!m_tailMode && m_runLeft > kLayoutMaxGroup && m_runLeft < 2u * kLayoutMaxGroup)
      {
        // the last two bins of a hole: the first one is sized so that the second can fill what is left exactly (with
        // groups of 10 and 4 a full bin of 32 in front of a rest of 2 or 6 would leave a crumb nothing ever fits)
        const uint64_t reach = reachable(0u);
        for (uint32_t first = kLayoutMaxGroup; first >= 1u; --first)
          if (((reach >> first) & 1ull) && ((reach >> (m_runLeft - first)) & 1ull)) { cap = first; oldestFirst = false; break; }
      }
      uint32_t nb = 0, used = 0;
      const uint32_t g0 = m_gSize[m_oldest];  // by FIFO order the oldest waiting group heads the queue of its size
      // (in front of the end of a run the oldest group goes first only if the rest can still be filled exactly: a hole
      // of 12 takes 4 + 4 + 4, not the 10 that has waited longest and a crumb of 2)
      if (oldestFirst && g0 <= cap && (cap == kLayoutMaxGroup || ((reachable(g0) >> (cap - g0)) & 1ull)))
      {
        bin[nb++] = popQueue(g0);
        used = g0;
      }
      used += fillFromWaiting(cap - used, bin, nb);
      if (used == 0u)
      {
        // nothing that waits fits what is left of this run: a hole stays a hole (for a later batch); at the END of the
        // pool (the tail has less room than the smallest waiting group, the walk of the holes is over) the oldest group
        // is placed on its own - into any hole that holds it, else slot by slot - and the tail is looked at afresh
        if (m_tailMode) placeGroup(popQueue(g0), slotOut);
        m_runLeft = 0u;
        continue;
      }
      const uint32_t at0 = m_runAt;
      if (m_tailMode) m_extent += used;
      else take(at0, used);
      m_runAt += used;
      m_runLeft -= used;
      m_binCursor = m_runAt;
      uint32_t at = at0;
      for (uint32_t b = 0; b < nb; ++b)
      {
        const uint32_t j = m_gStart[bin[b]], g = m_gSize[bin[b]];
        for (uint32_t k = 0; k < g; ++k) slotOut[j + k] = at + k;
        at += g;
      }
    }
  }

  void take(uint32_t f, uint32_t g)
  {
    clearBits(f, g);
    m_freeSlots -= g;
  }

  // one lap from this size's cursor: the first run that qualifies, or UINT32_MAX
  uint32_t sweep(uint32_t g, bool exactOrLong)
  {
    uint32_t pos = m_cursor[g] < m_extent ? m_cursor[g] : 0u;
    const uint32_t startedAt = pos;
    bool wrapped = false;
    for (;;)
    {
      const uint32_t limit = wrapped ? startedAt : m_extent;  // the second half of the lap ends where the first began
      ++m_steps;
      const uint32_t f = nextFree(pos, limit);
      if (f >= limit)
      {
        if (wrapped || startedAt == 0u) return 0xFFFFFFFFu;
        wrapped = true;
        pos = 0u;
        continue;
      }
      const uint32_t len = runLength(f, 64u);  // free slots from f on, counted up to 64 (a run may reach across `limit`: fine)
      if (f + g <= m_extent && (exactOrLong ? (len == g || len >= kLayoutMaxGroup) : len >= g))
      {
        take(f, g);
        m_cursor[g] = f + g;
        return f;
      }
      pos = len < 64u ? f + len : runEnd(f + 64u, limit);  // on to the end of this run
    }
  }

  // first free slot in [from, limit), or limit
  uint32_t nextFree(uint32_t from, uint32_t limit) const
  {
    if (from >= limit) return limit;
    size_t w = from >> 6;
    uint64_t bits = m_free[w] & (~0ull << (from & 63u));
    const size_t lastWord = (limit - 1u) >> 6;
    for (;;)
    {
      if (bits)
      {
        const uint32_t f = (uint32_t)(w << 6) + (uint32_t)__builtin_ctzll(bits);
        return f < limit ? f : limit;
      }
      if (++w > lastWord) return limit;
      bits = m_free[w];
    }
  }

  // first slot in [from, limit) that is NOT free, or limit
  uint32_t runEnd(uint32_t from, uint32_t limit) const
  {
    if (from >= limit) return limit;
    size_t w = from >> 6;
    uint64_t bits = ~m_free[w] & (~0ull << (from & 63u));
    const size_t lastWord = (limit - 1u) >> 6;
    for (;;)
    {
      if (bits)
      {
        const uint32_t e = (uint32_t)(w << 6) + (uint32_t)__builtin_ctzll(bits);
        return e < limit ? e : limit;
      }
      if (++w > lastWord) return limit;
      bits = ~m_free[w];
    }
  }

  // number of consecutive free slots starting at f, counted up to `cap` (<= 64)
  uint32_t runLength(uint32_t f, uint32_t cap) const
  {
    // 64 bits starting at f, from two words
    const size_t w = f >> 6;
    const uint32_t sh = f & 63u;
    uint64_t bits = m_free[w] >> sh;
    if (sh) bits |= m_free[w + 1] << (64u - sh);
    const uint64_t inv = ~bits;
    const uint32_t len = inv ? (uint32_t)__builtin_ctzll(inv) : 64u;
    return len < cap ? len : cap;
  }

  void clearBits(uint32_t f, uint32_t g)  // g <= 32: at most two words
  {
    const size_t w = f >> 6;
    const uint32_t sh = f & 63u;
    const uint64_t mask = g >= 64u ? ~0ull : ((1ull << g) - 1ull);
    m_free[w] &= ~(mask << sh);
    if (sh + g > 64u) m_free[w + 1] &= ~(mask >> (64u - sh));
  }

  uint32_t m_capacity = 0, m_extent = 0, m_freeSlots = 0;
  uint32_t m_cursor[kLayoutMaxGroup + 1] = {};        // per group size: where its next sweep starts
  std::vector<uint64_t> m_free;                       // one bit per slot: set = a hole below the extent
  uint32_t m_noExactSince[kLayoutMaxGroup + 1] = {};  // epoch in which a full lap found no exact or long run for that size
  uint32_t m_noRunSince[kLayoutMaxGroup + 1] = {};    // epoch in which a full lap found no run of that size at all
  uint32_t m_epoch = 1;                               // advanced by every release
  // placeBatch: the groups cut so far (first element, size, placed?), per size the groups still waiting (FIFO)
  std::vector<uint32_t> m_gStart;
  std::vector<uint8_t> m_gSize, m_gTaken;
  std::vector<uint32_t> m_queue[kLayoutMaxGroup + 1];
  uint32_t m_qHead[kLayoutMaxGroup + 1] = {};
  size_t m_oldest = 0;
  uint32_t m_pendingSlots = 0;
  uint64_t m_sizeMask = 0;                             // bit g: groups of size g are waiting
  uint32_t m_binCursor = 0;                            // where the previous batch stopped carving
  uint32_t m_runPos = 0, m_lapStart = 0, m_runAt = 0, m_runLeft = 0;
  bool m_wrapped = false, m_tailMode = false;
public:
  uint64_t m_steps = 0;                               // diagnostics: free runs examined so far
};

}  // namespace scgpu
