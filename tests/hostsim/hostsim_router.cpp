// The C++ shard router (sc-gameengine_b200/host/sc_gpu_shard_router.h), exported for ctypes (tests only).
#include "../../sc-gameengine_b200/host/sc_gpu_shard_router.h"
using sc::gpu::ShardRouter;
extern "C" {
ShardRouter* hs_router_create(uint32_t nRanks, uint32_t nCells, const int32_t* cellsXZ, const int32_t* owner, uint32_t maxEntityIndex)
{
  ShardRouter* r = new ShardRouter;
  if (!r->init(nRanks, nCells, cellsXZ, owner, maxEntityIndex)) { delete r; return nullptr; }
  return r;
}
void hs_router_destroy(ShardRouter* r) { delete r; }
void hs_router_rank_of_cell(const ShardRouter* r, uint32_t n, const int32_t* xz, int32_t* out)
{
  for (uint32_t i = 0; i < n; ++i) out[i] = r->rankOfCell(xz[2 * i], xz[2 * i + 1]);
}
int hs_router_spawn(ShardRouter* r, uint32_t n, const uint32_t* entity, const int32_t* rootCellXZ, int32_t* out)
{
  return r->routeSpawn(n, entity, rootCellXZ, out) ? 1 : 0;
}
void hs_router_rank_of(const ShardRouter* r, uint32_t n, const uint32_t* entity, int32_t* out) { r->rankOf(n, entity, out); }
void hs_router_despawn(ShardRouter* r, uint32_t n, const uint32_t* entity, int32_t* out) { r->routeDespawn(n, entity, out); }
void hs_router_counts(const ShardRouter* r, uint64_t* out) { r->counts(out); }
}
