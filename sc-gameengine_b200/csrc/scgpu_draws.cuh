// scgpu_draws.cuh — SURVEY.md §8(f) N1: the step right after the hot path, on the device.
//
// The reference's renderer takes RenderFrameData::draws and, on the CPU, every frame
//   (src/engine/src/sc_vk.cpp:1843-1852)  drops draws whose meshId is out of range or whose material is unknown,
//   (src/engine/src/sc_vk.cpp:1854-1864)  std::sorts the rest by (pipelineId of the material, materialId, meshId),
//   (src/engine/src/sc_vk.cpp:1866-1905)  walks the sorted list binding pipeline / material / mesh on change and
//                                         issuing one draw per item.
// Here: one 64-bit key per emitted draw, a stable radix sort of (key, position) pairs (cub::DeviceRadixSort — a plain
// library sort, restricted to the key bits that are in use), then hand-written kernels that gather the 80-byte
// DrawItems in sorted order and emit one RUN per (pipeline, material, mesh): what the bind-on-change loop derives
// item by item, i.e. the instanced batches. std::sort is not stable; this sort is, so its output is one of the
// orders the reference may produce (ties keep CullingState::visible order).
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "scgpu_kernels.cuh"

namespace scgpu
{

constexpr uint64_t kDrawKeyInvalid = ~0ull;
constexpr uint32_t kDrawIdBits = 29;  // materialId / meshId must be < 2^29, pipelineId < 2^6

struct DrawRun  // == ScGpuDrawRun
{
  uint32_t pipelineId, materialId, meshId, first, count;
};

__device__ __forceinline__ uint64_t draw_key(uint32_t pipe, uint32_t material, uint32_t mesh)
{
  return ((uint64_t)pipe << (2 * kDrawIdBits)) | ((uint64_t)material << kDrawIdBits) | (uint64_t)mesh;
}

// one key per emitted draw (the first `emitted` entries of the view's visible list, i.e. after maxDrawsBudget)
__global__ void __launch_bounds__(kBlock) k_draw_keys(const uint32_t* __restrict__ visSlot, const uint2* __restrict__ meshMat,
                                                      const uint32_t* __restrict__ materialPipeline, uint32_t nMaterials,
                                                      uint32_t meshCount, uint32_t emitted, uint64_t* __restrict__ keys,
                                                      uint32_t* __restrict__ pos, uint32_t* __restrict__ kept)
{
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  bool valid = false;
  if (i < emitted)
  {
    const uint2 mm = meshMat[visSlot[i]];  // x = meshId, y = materialId
    uint32_t pipe = 0xFFFFFFFFu;
    if (mm.y < nMaterials) pipe = materialPipeline[mm.y];
    valid = mm.x < meshCount && pipe != 0xFFFFFFFFu;  // sc_vk.cpp:1847-1850
    keys[i] = valid ? draw_key(pipe, mm.y, mm.x) : kDrawKeyInvalid;
    pos[i] = i;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31u) == 0 && m) atomicAdd(kept, (uint32_t)__popc(m));
}

// head flags of the runs of equal keys among the first *kept sorted keys
__global__ void __launch_bounds__(kBlock) k_draw_run_flags(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ kept,
                                                           uint32_t emitted, uint32_t* __restrict__ flags)
{
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= emitted) return;
  flags[i] = (i < *kept && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
}

// runIndex = inclusive scan of the head flags - 1. Heads write the run's key fields and start, the last item of a
// run its end; k_draw_run_counts turns ends into counts.
__global__ void __launch_bounds__(kBlock) k_draw_runs(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ runIncl,
                                                      const uint32_t* __restrict__ kept, DrawRun* __restrict__ runs,
                                                      uint32_t* __restrict__ nRuns)
{
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  const uint32_t n = *kept;
  if (i >= n) return;
  const uint32_t r = runIncl[i] - 1u;
  const uint64_t k = keys[i];
  if (i == 0 || keys[i - 1] != k)
  {
    const uint32_t idMask = (1u << kDrawIdBits) - 1u;
    runs[r].pipelineId = (uint32_t)(k >> (2 * kDrawIdBits));
    runs[r].materialId = (uint32_t)(k >> kDrawIdBits) & idMask;
    runs[r].meshId = (uint32_t)k & idMask;
    runs[r].first = i;
  }
  if (i + 1 == n || keys[i + 1] != k) runs[r].count = i + 1u;  // end, for now
  if (i + 1 == n) *nRuns = r + 1u;
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
