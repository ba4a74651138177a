// scgpu_peer.cuh — gather of the compacted visible lists to the submitting rank through NVLink PEER MEMORY.
//
// The only exchange of the sharded path is "every rank's per-view visible lists and counts -> the submitting rank"
// (SURVEY.md §8e). Over NCCL that costs an allgather of the counts, a host synchronisation (receive sizes must be
// known to post the receives) and nRanks x nViews send/recv pairs per frame: 0.07 ms at 2 GPUs, 0.4 ms at 8 — more
// than half of the 0.59 ms frame. Here the producing rank writes its lists straight into a mailbox in the ROOT's HBM
// (cudaIpc-mapped, stores travel over NVLink/NVSwitch) from the kernel that PRODUCES them (k_resolve_lists, the last
// kernel of the frame: compute and peer store fused), then raises a flag; the root runs one small kernel that waits
// for the flags. No host round trip, no NCCL on the per-frame path; NCCL remains the
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

// error word of one gather (ScGpuScene::dPeerState[1 + (seq & 1)], cleared one gather ahead): 1 = this (non-root) rank
// timed out waiting for the root's progress and delivered nothing; on the root: 2 = a rank's flag never came,
// 4 = a rank's lists exceed the mailbox, 8 = a rank reported that it delivered nothing.
// The PRODUCER side is k_resolve_lists (scgpu_kernels.cuh), the last kernel of every frame: with the peer gather
// enabled it stores every entity handle into the root's mailbox as it resolves it and its last CTA raises the flag.

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
