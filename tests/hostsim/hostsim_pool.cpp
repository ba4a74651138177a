// Host mirror of the Transform pool (sc-gameengine_b200/csrc/scgpu_pool.h), exported for ctypes (tests only).
#include "../../sc-gameengine_b200/csrc/scgpu_pool.h"
#include <chrono>
#include <cstring>
using namespace scgpu;

struct HsPool
{
  std::vector<uint32_t> dense, sparse;
  uint32_t count = 0;
  std::vector<PoolMove> moves;
  std::vector<uint32_t> removed;
  PoolScratch scratch;
  uint32_t threads = 1;
};

extern "C" {
HsPool* hs_pool_create(uint32_t sparseSize)
{
  HsPool* p = new HsPool;
  p->sparse.assign(sparseSize, 0u);
  return p;
}
void hs_pool_destroy(HsPool* p) { delete p; }
void hs_pool_set_threads(HsPool* p, uint32_t threads) { p->threads = threads; }
// 0 = ok, else 1 + reason (scgpu_pool.h), *badAt = offending position
int hs_pool_spawn(HsPool* p, uint32_t n, const uint32_t* entity, uint32_t* badAt)
{
  const int r = poolRegisterSpawn(p->dense, p->sparse, p->count, n, entity, badAt);
  if (r == 0) p->count += n;
  return r;
}
// returns the replay's wall time in seconds; results are read with the getters below
double hs_pool_despawn(HsPool* p, uint32_t n, const uint32_t* entity)
{
  const auto t0 = std::chrono::steady_clock::now();
  poolReplayDespawn(p->dense, p->sparse, p->count, n, entity, p->moves, p->removed, p->scratch, p->threads);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
uint32_t hs_pool_count(const HsPool* p) { return p->count; }
uint32_t hs_pool_num_moves(const HsPool* p) { return (uint32_t)p->moves.size(); }
uint32_t hs_pool_num_removed(const HsPool* p) { return (uint32_t)p->removed.size(); }
void hs_pool_read(const HsPool* p, uint32_t* dense, uint32_t* sparse, uint32_t* moves2, uint32_t* removed)
{
  if (dense) std::memcpy(dense, p->dense.data(), p->dense.size() * 4);
  if (sparse) std::memcpy(sparse, p->sparse.data(), p->sparse.size() * 4);
  if (moves2) std::memcpy(moves2, p->moves.data(), p->moves.size() * 8);
  if (removed) std::memcpy(removed, p->removed.data(), p->removed.size() * 4);
}
}
