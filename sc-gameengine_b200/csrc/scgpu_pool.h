// scgpu_pool.h — host mirror of the reference's ComponentPool<Transform> (src/core/include/sc_ecs.h:199-277).
//
// Pure host C++ (no CUDA types) so that tests/hostsim can compile it and the CPU suite can check it against a naive
// replay and against the reference's own pool. The device keeps the same two arrays (entity[slot], sparse[index] ->
// slot + 1); the mirror exists because ComponentPool::remove is order dependent: destroying entities one after the
// other swaps the LAST dense element into each hole (sc_ecs.h:228-247), so the final dense order depends on the
// sequence. A batch of despawns is replayed here handle by handle and reduced to the net `dst <- src` slot moves the
// device has to apply (k_despawn_apply).
//
// The replay is a chain of dependent random accesses into arrays of tens of megabytes (sparse[index] -> slot ->
// entity[slot], then sparse[index of the moved tail element]): unassisted it costs ~45 ns per handle, all of it
// cache misses. The loop therefore runs a two-stage software prefetch pipeline over the batch. Prefetches are
// hints computed from possibly stale values; the replay itself is unchanged, so the result is independent of them.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace scgpu
{

struct PoolMove
{
  uint32_t dst, src;  // same layout as the uint2 k_despawn_apply reads
};

constexpr uint32_t kPoolInvalidEntity = 0xFFFFFFFFu;  // SCGPU_INVALID_ENTITY / sc::kInvalidEntity (sc_ecs.h:22)
constexpr uint32_t kPoolIndexMask = 0xFFFFFFu;        // 24-bit entity index (sc_ecs.h:18-20)

// Replays ComponentPool::remove for entity[0..n) in order.
//   dense   : handles in pool order (size == count on entry, shrunk on return)
//   sparse  : entity index -> slot + 1, 0 = no component
//   origin  : scratch, slot -> slot its content came from; identity outside a call (grown on demand)
//   moves   : out, net moves for the surviving slots (sources always lie in the vacated tail [count_out, count_in),
//             so no source is also a destination and the moves can be applied in parallel)
//   removed : out, entity indices whose sparse entry must be cleared on the device
// Stale, unknown, repeated and invalid handles are skipped, like World::destroy returning false.
inline void poolReplayDespawn(std::vector<uint32_t>& dense, std::vector<uint32_t>& sparse, std::vector<uint32_t>& origin,
                              uint32_t& count, uint32_t n, const uint32_t* entity, std::vector<PoolMove>& moves,
                              std::vector<uint32_t>& removed)
{
  moves.clear();
  removed.clear();
  if (origin.size() < count)
  {
    const size_t old = origin.size();
    origin.resize(count);
    for (size_t i = old; i < origin.size(); ++i) origin[i] = (uint32_t)i;
  }
  std::vector<uint32_t> touched;
  touched.reserve(n);
  removed.reserve(n);
  const uint32_t count0 = count;
  uint32_t cnt = count;
  const size_t sparseSize = sparse.size();
  uint32_t* const pd = dense.data();
  uint32_t* const ps = sparse.data();
  uint32_t* const po = origin.data();

  constexpr uint32_t kFar = 24, kNear = 12;  // prefetch distances in handles
  for (uint32_t j = 0; j < n; ++j)
  {
    // stage 1: the sparse entry of a handle far ahead
    if (j + kFar < n)
    {
      const uint32_t i2 = entity[j + kFar] & kPoolIndexMask;
      if (i2 < sparseSize) __builtin_prefetch(ps + i2, 1);
    }
    // stage 2: its dense slot (sparse entry has arrived by now), the tail element that will probably fill the hole,
    // and that element's sparse entry
    if (j + kNear < n)
    {
      const uint32_t i1 = entity[j + kNear] & kPoolIndexMask;
      if (i1 < sparseSize)
      {
        const uint32_t sp1 = ps[i1];
        if (sp1 != 0u && sp1 <= cnt)
        {
          __builtin_prefetch(pd + (sp1 - 1u), 1);
          __builtin_prefetch(po + (sp1 - 1u), 1);
        }
      }
      if (cnt > kNear)
      {
        const uint32_t tail = pd[cnt - 1u - kNear] & kPoolIndexMask;  // sequential, cached
        if (tail < sparseSize) __builtin_prefetch(ps + tail, 1);
      }
    }

    const uint32_t e = entity[j];
    const uint32_t idx = e & kPoolIndexMask;
    if (e == kPoolInvalidEntity || idx >= sparseSize) continue;
    const uint32_t sp = ps[idx];
    if (sp == 0u || pd[sp - 1u] != e) continue;  // stale or unknown handle
    const uint32_t s = sp - 1u, last = cnt - 1u;
    if (s != last)
    {
      const uint32_t moved = pd[last];
      pd[s] = moved;
      ps[moved & kPoolIndexMask] = s + 1u;
      po[s] = po[last];
      touched.push_back(s);
    }
    ps[idx] = 0u;
    removed.push_back(idx);
    --cnt;
  }
  dense.resize(cnt);

  // net moves: final content of every touched slot that survived
  moves.reserve(touched.size());
  for (uint32_t s : touched)
  {
    if (s < cnt && po[s] != s)
    {
      moves.push_back(PoolMove{s, po[s]});
      po[s] = s;  // also dedups slots touched more than once
    }
  }
  for (uint32_t s : touched) po[s] = s;
  for (uint32_t s = cnt; s < count0; ++s) po[s] = s;
  count = cnt;
}

// Host mirror of World::create + add<Transform> for a batch: validates every handle first so that a failed call
// leaves the pool untouched, then appends in order. Returns 0 on success, else 1 + the reason and the offending
// position: 1 invalid handle, 2 index out of range, 3 index already owns a Transform.
inline int poolRegisterSpawn(std::vector<uint32_t>& dense, std::vector<uint32_t>& sparse, uint32_t count, uint32_t n,
                             const uint32_t* entity, uint32_t* badAt)
{
  const size_t sparseSize = sparse.size();
  uint32_t* const ps = sparse.data();
  constexpr uint32_t kAhead = 16;
  int why = 0;
  uint32_t j = 0;
  for (; j < n; ++j)
  {
    if (j + kAhead < n)
    {
      const uint32_t ia = entity[j + kAhead] & kPoolIndexMask;
      if (ia < sparseSize) __builtin_prefetch(ps + ia, 1);
    }
    const uint32_t idx = entity[j] & kPoolIndexMask;
    if (entity[j] == kPoolInvalidEntity) { why = 1; break; }
    if (idx >= sparseSize) { why = 2; break; }
    if (ps[idx] != 0u) { why = 3; break; }
    ps[idx] = count + j + 1u;
  }
  if (why)
  {
    for (uint32_t k = 0; k < j; ++k) ps[entity[k] & kPoolIndexMask] = 0u;
    if (badAt) *badAt = j;
    return why;
  }
  dense.insert(dense.end(), entity, entity + n);
  return 0;
}

}  // namespace scgpu
