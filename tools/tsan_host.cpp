// ThreadSanitizer run of the host half of scgpuDespawn / scgpuSpawn with the persistent helper threads (HostWorkers,
// scgpu_pool.h) and the slot layout: group-wise churn on a 1 M-entity pool, four threads.
//   g++ -std=c++17 -O1 -g -fsanitize=thread -pthread tools/tsan_host.cpp -o /tmp/tsan_host && /tmp/tsan_host
#include "../sc-gameengine_b200/csrc/scgpu_layout.h"
#include "../sc-gameengine_b200/csrc/scgpu_pool.h"
#include <cstdio>
#include <random>
using namespace scgpu;

int main()
{
  const uint32_t n = 1u << 20;
  HostWorkers workers(3);
  std::vector<uint32_t> dense, sparse(n, 0), slotOf(n, 0), scratch;
  uint32_t count = 0, bad = 0;
  SlotLayout lay;
  lay.reset(n + n / 8);
  std::mt19937 rng(1);
  std::vector<uint32_t> ent(n), par(n, 0xFFFFFFFFu), cohort(n);
  for (uint32_t i = 0; i < n;)
  {
    const uint32_t g = (rng() & 1) ? 4 : 10, c = rng() % 10;
    for (uint32_t k = 0; k < g && i < n; ++k, ++i) { ent[i] = i; par[i] = k ? i - k : 0xFFFFFFFFu; cohort[i] = c; }
  }
  poolRegisterSpawn(dense, sparse, count, n, ent.data(), &bad);
  const uint32_t s0 = lay.appendRun(n);
  for (uint32_t j = 0; j < n; ++j) slotOf[j] = s0 + j;
  count = n;
  std::vector<std::vector<uint32_t>> members(10);
  for (uint32_t j = 0; j < n; ++j) members[cohort[j]].push_back(j);
  uint32_t gen[10] = {0};
  std::vector<PoolMove> moves;
  std::vector<uint32_t> removed;
  PoolScratch ps;
  for (int f = 0; f < 12; ++f)
  {
    const int c = f % 10;
    const std::vector<uint32_t>& ix = members[c];
    const uint32_t m = (uint32_t)ix.size();
    std::vector<uint32_t> dead(m), fresh(m), fp(m), so(m);
    for (uint32_t k = 0; k < m; ++k) dead[k] = ix[k] | (gen[c] << 24);
    ++gen[c];
    for (uint32_t k = 0; k < m; ++k)
    {
      fresh[k] = ix[k] | (gen[c] << 24);
      fp[k] = par[ix[k]] == 0xFFFFFFFFu ? 0xFFFFFFFFu : (par[ix[k]] | (gen[c] << 24));
    }
    poolReplayDespawn(dense, sparse, count, m, dead.data(), moves, removed, ps, 4, &workers);
    scratch.resize(removed.size());
    uint32_t* sl = scratch.data();
    const uint32_t* so_ = slotOf.data();
    const uint32_t* ri = removed.data();
    poolParallelFor(4, (uint32_t)removed.size(), [=](uint32_t, uint32_t b, uint32_t e) { for (uint32_t v = b; v < e; ++v) sl[v] = so_[ri[v]]; }, &workers);
    lay.release((uint32_t)removed.size(), sl);
    if (poolRegisterSpawn(dense, sparse, count, m, fresh.data(), &bad)) { std::printf("spawn refused\n"); return 1; }
    lay.placeBatch(m, fresh.data(), fp.data(), so.data());
    uint32_t* sw = slotOf.data();
    const uint32_t* fr = fresh.data();
    const uint32_t* sp = so.data();
    poolParallelFor(4, m, [=](uint32_t, uint32_t b, uint32_t e) { for (uint32_t j = b; j < e; ++j) sw[fr[j] & 0xFFFFFFu] = sp[j]; }, &workers);
    count += m;
  }
  std::printf("TSAN HOST OK: %u live, extent %u, %u holes\n", count, lay.extent(), lay.freeSlots());
  return 0;
}
