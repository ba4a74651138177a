// scgpu_draws.cuh — SURVEY.md §8(f) N1: the step right after the hot path, on the device.
//
// The reference's renderer takes RenderFrameData::draws and, on the CPU, every frame
//   (src/engine/src/sc_vk.cpp:1843-1852)  drops draws whose meshId is out of range or whose material is unknown,
//   (src/engine/src/sc_vk.cpp:1854-1864)  std::sorts the rest by (pipelineId of the material, materialId, meshId),
//   (src/engine/src/sc_vk.cpp:1866-1905)  walks the sorted list binding pipeline / material / mesh on change and
//                                         issuing one draw per item.
// Here, all hand-written (no library sort):
//   k_draw_keys        one key per emitted draw, packed into exactly the bits the asset tables need —
//                      [dropped | pipeline | material | mesh] with bits(meshCount) + bits(nMaterials) + bits(pipelines) + 1
//                      bits: an engine with a few hundred meshes and materials sorts on ~18 bits, the sandbox on 7;
//   k_radix_hist / k_scan_tiles / k_radix_scatter
//                      stable least-significant-digit counting sort of (key, position) pairs, 8 key bits per pass and
//                      only as many passes as the key has digits (one pass for the sandbox's key);
//   k_draw_heads / k_scan_tiles / k_draw_runs / k_draw_run_counts
//                      one RUN per (pipeline, material, mesh): what the bind-on-change loop derives item by item,
//                      i.e. the instanced batches;
//   k_gather_sorted_draws  the 80-byte DrawItems in sorted order.
// std::sort is not stable; this sort is, so its output is one of the orders the reference may produce (ties keep
// CullingState::visible order).
#pragma once
#include "scgpu_kernels.cuh"

namespace scgpu
{

constexpr uint32_t kDrawIdBits = 29;  // materialId / meshId must be < 2^29, pipelineId < 63
constexpr uint32_t kSortItems = 8;    // keys per thread and block
constexpr uint32_t kSortTile = kBlock * kSortItems;
constexpr uint32_t kRadixBits = 8, kRadix = 1u << kRadixBits;

struct DrawRun  // == ScGpuDrawRun
{
  uint32_t pipelineId, materialId, meshId, first, count;
};

struct DrawKeyLayout
{
  uint32_t meshBits, matBits, pipeBits;  // key = dropped << (sum) | pipe << (meshBits + matBits) | material << meshBits | mesh
  __host__ __device__ uint32_t total() const { return meshBits + matBits + pipeBits; }
};

// one key per emitted draw (the first `emitted` entries of the view's visible list, i.e. after maxDrawsBudget)
__global__ void __launch_bounds__(kBlock) k_draw_keys(const uint32_t* __restrict__ visSlot, const uint2* __restrict__ meshMat,
                                                      const uint32_t* __restrict__ materialPipeline, uint32_t nMaterials,
                                                      uint32_t meshCount, uint32_t emitted, DrawKeyLayout lay,
                                                      uint64_t* __restrict__ keys, uint32_t* __restrict__ pos,
                                                      uint32_t* __restrict__ kept)
{
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  bool valid = false;
  if (i < emitted)
  {
    const uint2 mm = meshMat[visSlot[i]];  // x = meshId, y = materialId
    uint32_t pipe = 0xFFFFFFFFu;
    if (mm.y < nMaterials) pipe = materialPipeline[mm.y];
    valid = mm.x < meshCount && pipe != 0xFFFFFFFFu;  // sc_vk.cpp:1847-1850
    const uint64_t key = ((uint64_t)pipe << (lay.meshBits + lay.matBits)) | ((uint64_t)mm.y << lay.meshBits) | (uint64_t)mm.x;
    keys[i] = valid ? key : (1ull << lay.total());  // dropped draws sort behind every kept one
    pos[i] = i;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31u) == 0 && m) atomicAdd(kept, (uint32_t)__popc(m));
}

// ---- stable counting sort, one 8-bit digit per pass ---------------------------------------------------------------------
// hist[digit][block] = number of keys of the block's tile with that digit; an exclusive scan over the array in this
// (digit-major) order gives every (digit, block) its first output position.
__global__ void __launch_bounds__(kBlock) k_radix_hist(const uint64_t* __restrict__ keys, uint32_t n, uint32_t shift,
                                                       uint32_t* __restrict__ hist, uint32_t nBlocks)
{
  __shared__ uint32_t sHist[kRadix];
  sHist[threadIdx.x] = 0u;
  static_assert(kRadix == kBlock, "one bin per thread");
  __syncthreads();
  const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll
  for (uint32_t r = 0; r < kSortItems; ++r)
  {
    const uint32_t i = base + r * kBlock + threadIdx.x;
    if (i < n) atomicAdd(&sHist[(uint32_t)(keys[i] >> shift) & (kRadix - 1u)], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nBlocks + blockIdx.x] = sHist[threadIdx.x];
}

// Scatter in tile order. A tile is walked in kSortItems rounds of kBlock consecutive keys; inside a round the rank of a
// key among the equal digits before it = (lanes before it in its warp: match.any) + (equal digits in the warps before
// it: per-warp counts in shared memory, prefixed by one thread per digit) + (equal digits in earlier rounds: the
// running per-digit position).
__global__ void __launch_bounds__(kBlock) k_radix_scatter(const uint64_t* __restrict__ keysIn, const uint32_t* __restrict__ posIn,
                                                          uint64_t* __restrict__ keysOut, uint32_t* __restrict__ posOut,
                                                          uint32_t n, uint32_t shift, const uint32_t* __restrict__ histScan,
                                                          uint32_t nBlocks)
{
  constexpr uint32_t kWarps = kBlock / 32;
  __shared__ uint32_t sNext[kRadix];            // next output position per digit for this tile
  __shared__ uint16_t sWarpCnt[kWarps][kRadix]; // this round: keys per (warp, digit), then their start within the round
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  sNext[tid] = histScan[(size_t)tid * nBlocks + blockIdx.x];
  const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll 1
  for (uint32_t r = 0; r < kSortItems; ++r)
  {
    for (uint32_t w = 0; w < kWarps; ++w) sWarpCnt[w][tid] = 0;
    __syncthreads();
    const uint32_t i = base + r * kBlock + tid;
    const bool have = i < n;
    uint64_t key = 0;
    uint32_t p = 0, d = 0;
    if (have) { key = keysIn[i]; p = posIn[i]; d = (uint32_t)(key >> shift) & (kRadix - 1u); }
    // lanes of this warp with the same digit (lanes without a key form their own group under a digit nobody has)
    const uint32_t peers = __match_any_sync(0xffffffffu, have ? d : kRadix);
    const uint32_t before = __popc(peers & ((1u << lane) - 1u));
    if (have && before == 0u) sWarpCnt[warp][d] = (uint16_t)__popc(peers);
    __syncthreads();
    {
      // digit `tid`: exclusive prefix over the warps, on top of the running position
      uint32_t run = 0;
#pragma unroll
      for (uint32_t w = 0; w < kWarps; ++w)
      {
        const uint32_t c = sWarpCnt[w][tid];
        sWarpCnt[w][tid] = (uint16_t)run;
        run += c;
      }
      // (sNext is read below by other threads: keep this round's base, advance after the barrier)
      __syncthreads();
      if (have)
      {
        const uint32_t dst = sNext[d] + sWarpCnt[warp][d] + before;
        SC_ASSERT(dst < n);
        keysOut[dst] = key;
        posOut[dst] = p;
      }
      __syncthreads();
      sNext[tid] += run;
    }
  }
}

// ---- runs of equal keys among the first *kept sorted keys ------------------------------------------------------------------
__device__ __forceinline__ bool draw_is_head(const uint64_t* __restrict__ keys, uint32_t i, uint32_t kept)
{
  return i < kept && (i == 0u || keys[i] != keys[i - 1u]);
}

__global__ void __launch_bounds__(kBlock) k_draw_heads(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ kept,
                                                       uint32_t* __restrict__ tileHeads)
{
  __shared__ uint32_t sSum;
  const uint32_t n = *kept, base = blockIdx.x * kSortTile;
  if (threadIdx.x == 0) sSum = 0u;
  __syncthreads();
  uint32_t c = 0;
#pragma unroll
  for (uint32_t r = 0; r < kSortItems; ++r) c += draw_is_head(keys, base + r * kBlock + threadIdx.x, n) ? 1u : 0u;
  const uint32_t w = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31u) == 0 && w) atomicAdd(&sSum, w);
  __syncthreads();
  if (threadIdx.x == 0) tileHeads[blockIdx.x] = sSum;
}

// run index of item i = heads before it in earlier tiles (tileBase) + heads up to and including i in its tile. Heads
// write the run's key fields and start, the last item of a run its end; k_draw_run_counts turns ends into counts.
__global__ void __launch_bounds__(kBlock) k_draw_runs(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ kept,
                                                      const uint32_t* __restrict__ tileBase, DrawKeyLayout lay,
                                                      DrawRun* __restrict__ runs, uint32_t* __restrict__ nRuns)
{
  __shared__ uint32_t sWarp[kBlock / 32];
  const uint32_t n = *kept, base = blockIdx.x * kSortTile;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  // thread t owns kSortItems CONSECUTIVE items here (order matters for the scan)
  const uint32_t first = base + tid * kSortItems;
  uint32_t headBits = 0;
#pragma unroll
  for (uint32_t k = 0; k < kSortItems; ++k) headBits |= draw_is_head(keys, first + k, n) ? (1u << k) : 0u;
  const uint32_t mine = __popc(headBits);
  uint32_t x = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((int)lane >= o) x += y;
  }
  if (lane == 31) sWarp[warp] = x;
  __syncthreads();
  uint32_t before = tileBase[blockIdx.x] + x - mine;
  for (uint32_t w = 0; w < warp; ++w) before += sWarp[w];
  const uint32_t meshMask = (1u << lay.meshBits) - 1u, matMask = (1u << lay.matBits) - 1u;
#pragma unroll
  for (uint32_t k = 0; k < kSortItems; ++k)
  {
    const uint32_t i = first + k;
    if (i >= n) break;
    const uint64_t key = keys[i];
    if (headBits & (1u << k))
    {
      DrawRun& r = runs[before];
      r.pipelineId = (uint32_t)(key >> (lay.meshBits + lay.matBits));
      r.materialId = (uint32_t)(key >> lay.meshBits) & matMask;
      r.meshId = (uint32_t)key & meshMask;
      r.first = i;
      ++before;
    }
    if (i + 1u == n || keys[i + 1u] != key) runs[before - 1u].count = i + 1u;  // end, for now
    if (i + 1u == n) *nRuns = before;
  }
}

__global__ void __launch_bounds__(kBlock) k_draw_run_counts(DrawRun* __restrict__ runs, const uint32_t* __restrict__ nRuns)
{
  const uint32_t r = blockIdx.x * kBlock + threadIdx.x;
  if (r < *nRuns) runs[r].count -= runs[r].first;
}

// DrawItems in sorted order (sc::DrawItem 80 B, sc_ecs.h:159-165): 5 threads per item, one 16-byte chunk each
__global__ void __launch_bounds__(kBlock) k_gather_sorted_draws(const uint32_t* __restrict__ visSlot, const uint32_t* __restrict__ pos,
                                                                const uint32_t* __restrict__ kept, const uint32_t* __restrict__ entity,
                                                                const uint2* __restrict__ meshMat, const float4* __restrict__ w0,
                                                                const float4* __restrict__ w1, const float4* __restrict__ w2,
                                                                const float4* __restrict__ w3, float4* __restrict__ out)
{
  const uint64_t g = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
  if (g >= (uint64_t)(*kept) * 5ull) return;
  const uint32_t item = (uint32_t)(g / 5ull), chunk = (uint32_t)(g % 5ull);
  const uint32_t s = visSlot[pos[item]];
  float4 v;
  if (chunk == 0)
  {
    const uint2 mm = meshMat[s];
    v = make_float4(__uint_as_float(entity[s]), __uint_as_float(mm.x), __uint_as_float(mm.y), 0.f);
  }
  else if (chunk == 1) v = w0[s];
  else if (chunk == 2) v = w1[s];
  else if (chunk == 3) v = w2[s];
  else v = w3[s];
  out[g] = v;
}

}  // namespace scgpu
