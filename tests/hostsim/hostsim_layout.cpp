// Host half of scgpuSpawn / scgpuDespawn (pool mirror + slot layout: sc-gameengine_b200/csrc/scgpu_pool.h,
// scgpu_layout.h), exported for ctypes (tests only). The CPU suite plays the device half (k_spawn / k_despawn_apply /
// k_compact's rank -> slot -> entity walk) in numpy on top of it.
#include "../../sc-gameengine_b200/csrc/scgpu_layout.h"
#include "../../sc-gameengine_b200/csrc/scgpu_pool.h"
#include <cstring>
using namespace scgpu;

struct HsScene
{
  std::vector<uint32_t> dense, sparse, slotOf, scratch;
  uint32_t count = 0;
  std::vector<PoolMove> moves;
  std::vector<uint32_t> removed;
  PoolScratch pool;
  SlotLayout layout;
};

extern "C" {
HsScene* hs_scene_create(uint32_t capacity, uint32_t sparseSize)
{
  HsScene* s = new HsScene;
  s->sparse.assign(sparseSize, 0u);
  s->slotOf.assign(sparseSize, 0u);
  s->layout.reset(capacity);
  return s;
}
void hs_scene_destroy(HsScene* s) { delete s; }
uint32_t hs_scene_count(const HsScene* s) { return s->count; }
uint32_t hs_scene_extent(const HsScene* s) { return s->layout.extent(); }
uint32_t hs_scene_free(const HsScene* s) { return s->layout.freeSlots(); }
// the host half of scgpuSpawn (registerSpawn in scgpu_api.cu): 0 = ok and slotOut[j] = device slot, rank0 = first rank
int hs_scene_spawn(HsScene* s, uint32_t n, const uint32_t* entity, const uint32_t* parent, uint32_t* slotOut, uint32_t* rank0)
{
  if ((uint64_t)s->count + n > s->layout.capacity()) return 9;
  uint32_t at = 0;
  const int r = poolRegisterSpawn(s->dense, s->sparse, s->count, n, entity, &at);
  if (r) return r;
  if (!s->layout.hasHoles() && !SlotLayout::batchHasHierarchy(n, parent))
  {
    const uint32_t s0 = s->layout.appendRun(n);
    if (s0 == 0xFFFFFFFFu) return 8;
    for (uint32_t j = 0; j < n; ++j) slotOut[j] = s0 + j;
  }
  else if (!s->layout.placeBatch(n, entity, parent, slotOut)) return 8;
  for (uint32_t j = 0; j < n; ++j) s->slotOf[entity[j] & 0xFFFFFFu] = slotOut[j];
  *rank0 = s->count;
  s->count += n;
  return 0;
}
// the host half of scgpuDespawn: returns the number of victims; moves (dst, src pairs in rank space) and the victims'
// slots are read with hs_scene_read
uint32_t hs_scene_despawn(HsScene* s, uint32_t n, const uint32_t* entity)
{
  poolReplayDespawn(s->dense, s->sparse, s->count, n, entity, s->moves, s->removed, s->pool, 1);
  const uint32_t k = (uint32_t)s->removed.size();
  s->scratch.resize(k);
  for (uint32_t v = 0; v < k; ++v) s->scratch[v] = s->slotOf[s->removed[v]];
  s->layout.release(k, s->scratch.data());
  return k;
}
uint32_t hs_scene_num_moves(const HsScene* s) { return (uint32_t)s->moves.size(); }
void hs_scene_read(const HsScene* s, uint32_t* dense, uint32_t* moves2, uint32_t* removedSlot)
{
  if (dense) std::memcpy(dense, s->dense.data(), s->dense.size() * 4);
  if (moves2) std::memcpy(moves2, s->moves.data(), s->moves.size() * 8);
  if (removedSlot) std::memcpy(removedSlot, s->scratch.data(), s->removed.size() * 4);
}
uint32_t hs_scene_num_holes(const HsScene* s)
{
  std::vector<std::pair<uint32_t, uint32_t>> h;
  s->layout.holes(h);
  return (uint32_t)h.size();
}
void hs_scene_holes(const HsScene* s, uint32_t* startLen2)
{
  std::vector<std::pair<uint32_t, uint32_t>> h;
  s->layout.holes(h);
  for (size_t i = 0; i < h.size(); ++i) { startLen2[2 * i] = h[i].first; startLen2[2 * i + 1] = h[i].second; }
}
}
