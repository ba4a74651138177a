// scgpu_pool.h — host mirror of the reference's ComponentPool<Transform> (src/core/include/sc_ecs.h:199-277).
//
// Pure host C++ (no CUDA types) so that tests/hostsim can compile it and the CPU suite can check it against a naive
// replay and against the reference's own pool. The device keeps the same two arrays (entity[slot], sparse[index] ->
// slot + 1); the mirror exists because ComponentPool::remove is order dependent: destroying entities one after the
// other swaps the LAST dense element into each hole (sc_ecs.h:228-247), so the final dense order depends on the
// sequence. A batch of despawns is reduced here to the net `dst <- src` slot moves the device has to apply
// (k_despawn_apply).
//
// A literal handle-by-handle replay is a chain of dependent cache misses into arrays of tens of megabytes (sparse[index]
// -> slot -> entity[slot] -> sparse[index of the moved tail element]; ~45 ns per handle). poolReplayDespawn computes
// the SAME final state in three passes whose misses are independent of each other:
//   A  gather   every handle's slot (sparse, then entity for the liveness check); iterations do not depend on each
//               other: software prefetch keeps many misses in flight and the range can be cut across host threads.
//               A bitmap over the slots tells whether the batch names an element twice; if not (the normal case) the
//               victims are compacted in parallel too, else one sequential walk lets the first occurrence win;
//   B  simulate the swap-with-last sequence on the vacated tail only. With k valid victims the pool shrinks from
//               count to base = count - k, and three facts make the tail [base, count) self-contained:
//                 - the element moved by a removal is the current last one, and last >= base throughout;
//                 - so only elements that START in the tail ever move, and every one of them that survives ends up
//                   in a slot below base (the tail is exactly what gets vacated);
//                 - a victim that starts below base is removed from the slot it started in.
//               The state is two k-sized arrays (which tail element sits in a tail slot / where a tail element is
//               now), walked sequentially and cache resident;
//   C  scatter  the surviving tail elements to their final slots (entity and sparse writes, again independent and
//               cut across host threads).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace scgpu
{

struct PoolMove
{
  uint32_t dst, src;  // same layout as the uint2 k_despawn_apply reads
};

constexpr uint32_t kPoolInvalidEntity = 0xFFFFFFFFu;  // SCGPU_INVALID_ENTITY / sc::kInvalidEntity (sc_ecs.h:22)
constexpr uint32_t kPoolIndexMask = 0xFFFFFFu;        // 24-bit entity index (sc_ecs.h:18-20)

// Work arrays of poolReplayDespawn, kept between calls: a batch needs ~25 bytes per handle, and taking megabytes from
// the allocator per call means fresh pages (one fault per 4 KiB) every frame.
struct PoolScratch
{
  std::vector<uint32_t> raw, slot0, content, where;
  std::vector<uint8_t> dead;
  std::vector<uint64_t> seen;  // one bit per slot, all zero between calls
  std::vector<uint32_t> chunkMoves, partCount;
};

// Helper threads that live as long as their owner (one set per scgpu context): run(parts, fn) executes fn(part) for
// every part in [0, parts), the caller taking parts like any helper, and returns when all are done. Creating and
// joining std::threads per call cost 50-100 us a time, several times per frame; these sleep on a condition variable
// between calls. No exception leaves run(); a helper that could not be created is simply absent.
class HostWorkers
{
public:
  explicit HostWorkers(uint32_t helpers)
  {
    m_threads.reserve(helpers);
    for (uint32_t i = 0; i < helpers; ++i)
    {
      try { m_threads.emplace_back([this] { loop(); }); }
      catch (...) { break; }  // thread limit reached: fewer helpers
    }
  }
  ~HostWorkers()
  {
    {
      std::lock_guard<std::mutex> lk(m_mutex);
      m_quit = true;
      ++m_generation;
    }
    m_wake.notify_all();
    for (std::thread& t : m_threads) t.join();
  }
  HostWorkers(const HostWorkers&) = delete;
  HostWorkers& operator=(const HostWorkers&) = delete;
  uint32_t helpers() const { return (uint32_t)m_threads.size(); }

  void run(uint32_t parts, const std::function<void(uint32_t)>& fn)
  {
    if (parts == 0) return;
    if (parts == 1 || m_threads.empty()) { for (uint32_t p = 0; p < parts; ++p) fn(p); return; }
    {
      std::lock_guard<std::mutex> lk(m_mutex);
      m_fn = &fn;
      m_parts = parts;
      m_next.store(0, std::memory_order_relaxed);
      m_done = 0;
      ++m_generation;
    }
    m_wake.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_mutex);
    m_finished.wait(lk, [&] { return m_done == m_parts; });
    m_fn = nullptr;
  }

private:
  void work()
  {
    uint32_t mine = 0;
    for (;;)
    {
      const uint32_t p = m_next.fetch_add(1, std::memory_order_relaxed);
      if (p >= m_parts) break;
      (*m_fn)(p);
      ++mine;
    }
    if (mine)
    {
      std::lock_guard<std::mutex> lk(m_mutex);
      m_done += mine;
      if (m_done == m_parts) m_finished.notify_all();
    }
  }
  void loop()
  {
    uint64_t seen = 0;
    for (;;)
    {
      {
        std::unique_lock<std::mutex> lk(m_mutex);
        m_wake.wait(lk, [&] { return m_generation != seen; });
        seen = m_generation;
        if (m_quit) return;
        if (!m_fn) continue;  // woke up after the job was already finished by the others
      }
      work();
    }
  }

  std::vector<std::thread> m_threads;
  std::mutex m_mutex;
  std::condition_variable m_wake, m_finished;
  const std::function<void(uint32_t)>* m_fn = nullptr;
  std::atomic<uint32_t> m_next{0};
  uint32_t m_parts = 0, m_done = 0;
  uint64_t m_generation = 0;
  bool m_quit = false;
};

// Runs fn(part, begin, end) over [0, n) cut into `parts` contiguous ranges: on the owner's helper threads when it has
// some, else on short-lived threads (tests/hostsim), part 0 on the calling thread.
template <class Fn>
inline void poolParallelFor(uint32_t parts, uint32_t n, Fn fn, HostWorkers* workers = nullptr)
{
  if (parts <= 1u || n < 32768u) { fn(0u, 0u, n); return; }
  const uint32_t per = (n + parts - 1u) / parts;
  if (workers)
  {
    workers->run(parts, [&](uint32_t p) {
      const uint32_t b = std::min(n, p * per), e = std::min(n, b + per);
      fn(p, b, e);
    });
    return;
  }
  std::vector<std::thread> th;
  th.reserve(parts - 1u);
  uint32_t started = 1;  // ranges [0, started) have an owner; a thread that cannot be created leaves its range to the caller
  for (; started < parts; ++started)
  {
    const uint32_t p = started, b = std::min(n, p * per), e = std::min(n, b + per);
    try { th.emplace_back([=] { fn(p, b, e); }); }
    catch (...) { break; }  // no exception may cross the C ABI: the rest runs on the calling thread
  }
  fn(0u, 0u, std::min(n, per));
  for (uint32_t p = started; p < parts; ++p)
  {
    const uint32_t b = std::min(n, p * per), e = std::min(n, b + per);
    fn(p, b, e);
  }
  for (std::thread& t : th) t.join();
}

// Final state of ComponentPool::remove applied to entity[0..n) in order.
//   dense   : handles in pool order (size == count on entry, shrunk on return)
//   sparse  : entity index -> slot + 1, 0 = no component
//   moves   : out, one `dst <- src` per surviving element of the vacated tail: src in [count_out, count_in)
//             ascending, dst < count_out, all dst distinct — they can be applied in parallel and in any order
//   removed : out, entity indices whose sparse entry must be cleared on the device, in batch order
//   threads : host threads for the gather and scatter passes (their iterations are independent; each is a chain of
//             two cache misses, so this is latency hiding more than arithmetic); 1 = the calling thread only. The
//             result does not depend on it.
// Stale, unknown, repeated and invalid handles are skipped, like World::destroy returning false.
inline void poolReplayDespawn(std::vector<uint32_t>& dense, std::vector<uint32_t>& sparse, uint32_t& count, uint32_t n,
                              const uint32_t* entity, std::vector<PoolMove>& moves, std::vector<uint32_t>& removed,
                              PoolScratch& scratch, uint32_t threads = 1, HostWorkers* workers = nullptr)
{
  moves.clear();
  removed.clear();
  if (n == 0) return;
  if (removed.capacity() < n) removed.reserve(n + n / 4);
  const uint32_t count0 = count;
  const size_t sparseSize = sparse.size();
  uint32_t* const pd = dense.data();
  uint32_t* const ps = sparse.data();
  threads = std::max(1u, std::min(threads, 64u));

  // ---- A1 (parallel): slot + 1 of every handle that names a live element, else 0. Every live slot named is also
  // marked in a bitmap (one bit per slot); finding the bit already set means the batch names an element twice.
  if (scratch.raw.size() < n) scratch.raw.resize(n + n / 4);
  uint32_t* const raw = scratch.raw.data();
  if (scratch.seen.size() < ((size_t)count0 + 63u) / 64u) scratch.seen.resize(((size_t)count0 + 63u) / 64u + 1024u, 0ull);
  uint64_t* const seen = scratch.seen.data();
  const uint32_t partsA = (threads <= 1u || n < 32768u) ? 1u : threads;
  scratch.partCount.assign(partsA + 1u, 0u);
  uint32_t* const partCount = scratch.partCount.data();
  std::atomic<uint32_t> repeats{0u};
  poolParallelFor(partsA, n, [=, &repeats](uint32_t part, uint32_t b, uint32_t e_) {
    constexpr uint32_t kFar = 32, kNear = 16;  // prefetch distances in handles: sparse entry, then the dense slot
    uint32_t valid = 0, again = 0;
    for (uint32_t j = b; j < e_; ++j)
    {
      if (j + kFar < e_)
      {
        const uint32_t i2 = entity[j + kFar] & kPoolIndexMask;
        if (i2 < sparseSize) __builtin_prefetch(ps + i2, 0);
      }
      if (j + kNear < e_)
      {
        const uint32_t i1 = entity[j + kNear] & kPoolIndexMask;
        if (i1 < sparseSize)
        {
          const uint32_t sp1 = ps[i1];
          if (sp1 != 0u && sp1 <= count0) __builtin_prefetch(pd + (sp1 - 1u), 0);
        }
      }
      const uint32_t e = entity[j];
      const uint32_t idx = e & kPoolIndexMask;
      uint32_t r = 0u;
      if (e != kPoolInvalidEntity && idx < sparseSize)
      {
        const uint32_t sp = ps[idx];
        if (sp != 0u && sp <= count0 && pd[sp - 1u] == e) r = sp;
      }
      raw[j] = r;
      if (r)
      {
        const uint32_t s = r - 1u;
        const uint64_t bit = 1ull << (s & 63u);
        uint64_t old;
        if (partsA == 1u) { old = seen[s >> 6]; seen[s >> 6] = old | bit; }
        else old = __atomic_fetch_or(seen + (s >> 6), bit, __ATOMIC_RELAXED);
        again += (old & bit) ? 1u : 0u;
        ++valid;
      }
    }
    partCount[part + 1u] = valid;
    if (again) repeats.fetch_add(again, std::memory_order_relaxed);
  }, workers);

  // ---- A2: the victims in batch order (slot0) and their entity indices (removed)
  std::vector<uint32_t>& slot0 = scratch.slot0;
  if (repeats.load() == 0u)
  {
    // no element is named twice: every non-zero raw entry is a victim, and the ranges compact in parallel
    for (uint32_t p = 0; p < partsA; ++p) partCount[p + 1u] += partCount[p];
    const uint32_t kAll = partCount[partsA];
    slot0.resize(kAll);
    removed.resize(kAll);
    uint32_t* const s0 = slot0.data();
    uint32_t* const rm = removed.data();
    poolParallelFor(partsA, n, [=](uint32_t part, uint32_t b, uint32_t e_) {
      uint32_t w = partCount[part];
      for (uint32_t j = b; j < e_; ++j)
      {
        if (raw[j] == 0u) continue;
        s0[w] = raw[j] - 1u;
        rm[w] = entity[j] & kPoolIndexMask;
        ++w;
      }
    }, workers);
  }
  else
  {
    // a handle repeated in the batch is stale the second time: walk the batch in order, first occurrence wins
    for (uint32_t j = 0; j < n; ++j) if (raw[j]) seen[(raw[j] - 1u) >> 6] = 0ull;  // every set bit is this batch's
    slot0.clear();
    if (slot0.capacity() < n) slot0.reserve(n + n / 4);
    for (uint32_t j = 0; j < n; ++j)
    {
      if (raw[j] == 0u) continue;
      const uint32_t s = raw[j] - 1u;
      const uint64_t bit = 1ull << (s & 63u);
      if (seen[s >> 6] & bit) continue;
      seen[s >> 6] |= bit;
      slot0.push_back(s);
      removed.push_back(entity[j] & kPoolIndexMask);
    }
  }
  const uint32_t k = (uint32_t)slot0.size();
  if (k == 0) return;
  const uint32_t base = count0 - k;

  // ---- B (sequential, cache resident): the swap-with-last sequence on the tail [base, count0) only.
  // content: tail slot -> tail element in it; where: tail element -> its current slot; dead: tail element is a victim
  if (scratch.content.size() < k)
  {
    scratch.content.resize(k + k / 4);
    scratch.where.resize(k + k / 4);
    scratch.dead.resize(k + k / 4);
  }
  uint32_t* const content = scratch.content.data();
  uint32_t* const where = scratch.where.data();
  uint8_t* const dead = scratch.dead.data();
  for (uint32_t t = 0; t < k; ++t) { content[t] = where[t] = base + t; dead[t] = 0; }
  for (uint32_t v = 0; v < k; ++v)
  {
    const uint32_t s0 = slot0[v];
    seen[s0 >> 6] = 0ull;  // leaves the bitmap clean for the next call
    if (s0 >= base) dead[s0 - base] = 1;
  }
  uint32_t cnt = count0;
  for (uint32_t v = 0; v < k; ++v)
  {
    const uint32_t s0 = slot0[v];
    const uint32_t s = s0 >= base ? where[s0 - base] : s0;
    const uint32_t last = cnt - 1u;
    if (s != last)
    {
      const uint32_t o = content[last - base];  // the live element in the last slot (always a tail element)
      if (s >= base) content[s - base] = o;
      where[o - base] = s;
    }
    --cnt;
  }

  // ---- C (parallel): the victims' sparse entries are cleared, surviving tail elements go to their final slots.
  // Writes are disjoint: cleared indices belong to victims, set indices to survivors, dst slots are distinct and
  // below base, sources at or above it. Moves are collected per range and concatenated in range order.
  const uint32_t parts = (threads <= 1u || k < 32768u) ? 1u : threads;
  if (moves.capacity() < k) moves.reserve(k + k / 4);
  moves.resize(k);
  PoolMove* const mv = moves.data();
  scratch.chunkMoves.assign(parts, 0u);
  uint32_t* const chunkMoves = scratch.chunkMoves.data();
  const uint32_t* const rem = removed.data();
  poolParallelFor(parts, k, [=](uint32_t part, uint32_t b, uint32_t e_) {
    for (uint32_t v = b; v < e_; ++v)
    {
      if (v + 16u < e_) __builtin_prefetch(ps + rem[v + 16u], 1);
      ps[rem[v]] = 0u;
    }
    uint32_t m = b;  // range-local output position: at most e_ - b moves, written in place at [b, ...)
    for (uint32_t t = b; t < e_; ++t)
    {
      if (t + 16u < e_ && !dead[t + 16u]) __builtin_prefetch(pd + where[t + 16u], 1);
      if (dead[t]) continue;
      const uint32_t dst = where[t], h = pd[base + t];
      pd[dst] = h;
      ps[h & kPoolIndexMask] = dst + 1u;
      mv[m++] = PoolMove{dst, base + t};
    }
    chunkMoves[part] = m - b;
  }, workers);
  // close the gaps between the ranges' move lists
  uint32_t total = chunkMoves[0];
  if (parts > 1u)
  {
    const uint32_t per = (k + parts - 1u) / parts;
    for (uint32_t p = 1; p < parts; ++p)
    {
      const uint32_t b = std::min(k, p * per);
      if (total != b) std::memmove(mv + total, mv + b, (size_t)chunkMoves[p] * sizeof(PoolMove));
      total += chunkMoves[p];
    }
  }
  moves.resize(total);
  dense.resize(base);
  count = base;
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
    const uint32_t idx = entity[j] & kPoolIndexMask;
    if (j + kAhead < n)
    {
      // scattered handles (a recycled free list) only: for runs of consecutive indices the hardware prefetcher does
      // the work and the extra instruction costs more than the loop body
      const uint32_t ia = entity[j + kAhead] & kPoolIndexMask;
      if (ia - idx > 4u * kAhead && ia < sparseSize) __builtin_prefetch(ps + ia, 1);
    }
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
