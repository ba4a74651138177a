// scgpu_peer.cuh — gather of the compacted visible lists to the submitting rank through NVLink PEER MEMORY.
//
// The only exchange of the sharded path is "every rank's per-view visible lists and counts -> the submitting rank"
// (SURVEY.md §8e). Over NCCL that costs an allgather of the counts, a host synchronisation (receive sizes must be
// known to post the receives) and nRanks x nViews send/recv pairs per frame: 0.07 ms at 2 GPUs, 0.4 ms at 8 — more
// than half of the 0.59 ms frame. Here the producing rank writes its lists straight into a mailbox in the ROOT's HBM
// (cudaIpc-mapped, stores travel over NVLink/NVSwitch) from the kernel that packs them, then raises a flag; the root
// runs one small kernel that waits for the flags. No host round trip, no NCCL on the per-frame path; NCCL remains the
// bootstrap (it carries the IPC handle) and the fallback (scgpuGatherVisible without scgpuCommEnablePeerGather).
//
// Mailbox layout in the root's memory (one allocation):
//   word 0                     rootProgress: sequence number of the gather the root's stream has reached
//   header[rank][parity]       64 B: flag (sequence number when complete), counts[kMaxViews+2], overflow
//   payload[rank][parity]      cap entries: the rank's lists, view after view
// Two parities: a producer may run one frame ahead of the root's consumers. It waits (bounded) until
// rootProgress >= seq-1, i.e. until everything the root had enqueued before its gather seq-1... has consumed the
// buffer this sequence number reuses.
#pragma once
#include "scgpu_kernels.cuh"

namespace scgpu
{

constexpr uint32_t kPeerHeaderWords = 16;  // 64 B
constexpr uint32_t kPeerFlag = 0, kPeerCounts = 1, kPeerOverflow = kPeerCounts + kMaxViews + 2;
constexpr long long kPeerSpinClocks = 4000000000ll;  // ~2 s at 1.9 GHz: a dead peer must not hang the box
// error word of one gather (ScGpuScene::dPeerState[1], cleared at the start of every gather): 1 = this (non-root) rank
// timed out waiting for the root's progress and delivered nothing; on the root: 2 = a rank's flag never came,
// 4 = a rank's lists exceed the mailbox, 8 = a rank reported that it delivered nothing

struct PeerBox
{
  uint32_t* base;      // mailbox in the root's memory (local pointer on the root, IPC mapping elsewhere)
  uint32_t nRanks;
  uint32_t cap;        // payload entries per (rank, parity)
  __host__ __device__ uint32_t* progress() const { return base; }
  __host__ __device__ uint32_t* header(uint32_t rank, uint32_t parity) const
  {
    return base + kPeerHeaderWords * (1u + rank * 2u + parity);
  }
  __host__ __device__ uint32_t* payload(uint32_t rank, uint32_t parity) const
  {
    return base + kPeerHeaderWords * (1u + 2u * nRanks) + (size_t)(rank * 2u + parity) * cap;
  }
  static size_t bytes(uint32_t nRanks, uint32_t cap)
  {
    return ((size_t)kPeerHeaderWords * (1u + 2u * nRanks) + (size_t)nRanks * 2u * cap) * 4u;
  }
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct PeerPackParams
{
  PeerBox box;
  const uint32_t* totals;              // this rank's device counts row (kMaxViews+2 words)
  const uint32_t* visEntity[kMaxViews];
  uint32_t* done;                      // local ticket counter (zero between launches)
  uint32_t* error;                     // local error word (1 = root progress timeout)
  uint32_t seq, rank, nViews, isRoot;
};

// Every rank: copy this rank's lists into its slot of the root's mailbox, then publish counts + flag.
__global__ void __launch_bounds__(kBlock) k_peer_pack(const __grid_constant__ PeerPackParams q)
{
  __shared__ uint32_t sOff[kMaxViews + 1];
  __shared__ uint32_t sLast, sStale;
  const uint32_t parity = q.seq & 1u;
  if (threadIdx.x == 0)
  {
    sStale = 0u;
    if (q.isRoot)
    {
      // everything the root enqueued before this gather (the consumers of gather seq-1 included) has completed
      if (blockIdx.x == 0) st_release_sys(q.box.progress(), q.seq);
    }
    else
    {
      // the buffer of this parity was last used by gather seq-2: wait until the root is past gather seq-1's start
      const long long t0 = clock64();
      while ((int32_t)(ld_acquire_sys(q.box.progress()) - (q.seq - 1u)) < 0)
      {
        // the root may still be reading this buffer: write NOTHING into it, report the frame as failed instead
        if (clock64() - t0 > kPeerSpinClocks) { atomicOr(q.error, 1u); sStale = 1u; break; }
        __nanosleep(200);
      }
    }
    uint32_t off = 0;
    for (uint32_t v = 0; v < q.nViews; ++v) { sOff[v] = off; off += q.totals[v]; }
    sOff[q.nViews] = off;
  }
  __syncthreads();
  const uint32_t total = sOff[q.nViews];
  const bool overflow = total > q.box.cap;
  const bool stale = sStale != 0u;
  uint32_t* dst = q.box.payload(q.rank, parity);
  if (!overflow && !stale)
  {
    for (uint32_t v = 0; v < q.nViews; ++v)
    {
      const uint32_t n = sOff[v + 1] - sOff[v];
      const uint32_t* src = q.visEntity[v];
      uint32_t* d = dst + sOff[v];
      for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) d[i] = src[i];
    }
  }
  __threadfence_system();  // this thread's remote stores are visible system-wide before the ticket
  __syncthreads();
  if (threadIdx.x == 0) sLast = (atomicAdd(q.done, 1u) == gridDim.x - 1u) ? 1u : 0u;
  __syncthreads();
  if (sLast)
  {
    uint32_t* h = q.box.header(q.rank, parity);
    // (the header is 64 B that the root reads only after the flag: safe to write even when the payload was not)
    if (threadIdx.x < kMaxViews + 2) h[kPeerCounts + threadIdx.x] = q.totals[threadIdx.x];
    if (threadIdx.x == 0) h[kPeerOverflow] = stale ? 2u : (overflow ? 1u : 0u);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
      *q.done = 0u;
      st_release_sys(h + kPeerFlag, q.seq);
    }
  }
}

// Root: wait until every rank's flag carries this sequence number; collect the counts rows.
__global__ void __launch_bounds__(64) k_peer_wait(PeerBox box, uint32_t seq, uint32_t* __restrict__ allCounts, uint32_t* __restrict__ error)
{
  const uint32_t parity = seq & 1u;
  for (uint32_t r = threadIdx.x; r < box.nRanks; r += blockDim.x)
  {
    const uint32_t* h = box.header(r, parity);
    const long long t0 = clock64();
    bool ok = true;
    while (ld_acquire_sys(h + kPeerFlag) != seq)
    {
      if (clock64() - t0 > kPeerSpinClocks) { ok = false; break; }
      __nanosleep(100);
    }
    if (!ok) atomicOr(error, 2u);
    else if (h[kPeerOverflow] == 1u) atomicOr(error, 4u);
    else if (h[kPeerOverflow] == 2u) { atomicOr(error, 8u); ok = false; }  // the producer gave up: its payload is not this frame's
    for (uint32_t k = 0; k < kMaxViews + 2; ++k) allCounts[r * (kMaxViews + 2) + k] = ok ? h[kPeerCounts + k] : 0u;
  }
}

}  // namespace scgpu
