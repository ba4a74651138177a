// scgpu_kernels.cuh — sm_100a kernels of the scene-update hot path.
//
// HBM layout (one "slot" per Transform, slot order == the reference's Transform-pool dense order, so every
// output list comes out in the reference's order without a sort):
//   rec0[slot] = { pos.x, pos.y, pos.z, rot.x }            float4 planes: one 128-bit load per thread, a warp
//   rec1[slot] = { rot.y, rot.z, scale.x, scale.y }        reads 512 contiguous bytes per instruction.
//   rec2[slot] = { scale.z, aabbMin.x, aabbMin.y, aabbMin.z }   64 B per instance = TRS 36 + AABB 24 + flags 4,
//   rec3[slot] = { aabbMax.x, aabbMax.y, aabbMax.z, flags }     exactly SURVEY.md §8(d)'s read set.
//   world0..3[slot] = world matrix columns (float4 planes, coalesced 128-bit stores)
//   parentSlot[slot] = resolved parent slot or kNone (maintained by k_resolve_parents on topology changes)
//   flags: bit0 HAS_BOUNDS, bit1 HAS_MESH, bits 8..31 = dirty stamp (id of the update that must recompute
//          the instance). A stamp instead of a dirty bit means the frame kernel never writes the records.
//
// Frame = k_update (transform + sphere + V-view plane tests + per-tile counts, one pass over the records)
//         -> k_scan_tiles (exclusive scan of the per-tile counts, V+1 rows)
//         -> k_scatter_visible (stable per-view compaction of entity handles / slots)
#pragma once
#include "scgpu_math.cuh"

namespace scgpu
{

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kMaxViews = 8;
constexpr uint32_t kBlock = 256;       // threads per CTA
constexpr uint32_t kSubTiles = 4;      // sub-tiles of kBlock slots per CTA
constexpr uint32_t kTile = kBlock * kSubTiles;
constexpr uint32_t kFlagBounds = 1u, kFlagMesh = 2u;
constexpr uint32_t kStampShift = 8;

constexpr uint32_t kUpdForceDirty = 1u, kUpdFreeze = 2u, kUpdSkipTransform = 4u;

struct ViewPlanes
{
  float4 planes[kMaxViews][6];
};

struct UpdateParams
{
  const float4* rec0;
  const float4* rec1;
  const float4* rec2;
  const float4* rec3;
  float4* w0;
  float4* w1;
  float4* w2;
  float4* w3;
  const uint32_t* parentSlot;
  uint8_t* vismask;
  uint32_t* tileCounts;  // [(nViews+1)][numTiles]; row nViews = culling candidates
  uint32_t* recomputed;  // single counter
  uint32_t count;
  uint32_t numTiles;
  uint32_t stamp;
  uint32_t nViews;
  uint32_t flags;
};

// ---- loads / stores ---------------------------------------------------------------------------------

__device__ __forceinline__ float4 ld_stream(const float4* p)
{
  // read-once streaming data: non-coherent path, do not keep in L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ Mat4 load_world(const UpdateParams& p, uint32_t s)
{
  Mat4 m;
  m.c0 = p.w0[s]; m.c1 = p.w1[s]; m.c2 = p.w2[s]; m.c3 = p.w3[s];
  return m;
}

__device__ __forceinline__ void store_world(const UpdateParams& p, uint32_t s, const Mat4& m)
{
  p.w0[s] = m.c0; p.w1[s] = m.c1; p.w2[s] = m.c2; p.w3[s] = m.c3;
}

// local TRS matrix of a slot (mat4_trs, sc_math.cpp:130-142)
__device__ __forceinline__ Mat4 local_of(const UpdateParams& p, uint32_t s)
{
  const float4 a = p.rec0[s];
  const float4 b = p.rec1[s];
  const float sz = p.rec2[s].x;
  return mat4_trs_dense(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, sz);
}

__device__ __forceinline__ bool slot_dirty(const UpdateParams& p, uint32_t s)
{
  if (p.flags & kUpdSkipTransform) return false;
  if (p.flags & kUpdForceDirty) return true;
  return (__float_as_uint(p.rec3[s].w) >> kStampShift) == p.stamp;
}

// ---- ancestor walk for a parent outside the CTA's sub-tile --------------------------------------------
// Reproduces what the reference's DFS (sc_ecs.cpp:167-210) would have produced for slot `ps` this frame without
// depending on any other thread: finds the ancestor closest to the root that is dirty, starts from the stored
// world matrix of ITS parent (clean with clean ancestors => not written by anyone this frame) and multiplies
// down. Returns false when the chain never reaches a root (cycle): such nodes are never visited by the DFS.
__device__ __noinline__ bool walk_up(const UpdateParams& p, uint32_t ps, bool needWorld, Mat4& outW, bool& outDirty)
{
  // pass 1: Brent cycle detection + index of the dirty ancestor closest to the root
  int lastDirty = -1;
  {
    uint32_t cur = ps, tortoise = ps;
    int steps = 0, power = 1, lam = 0;
    for (;;)
    {
      if (slot_dirty(p, cur)) lastDirty = steps;
      const uint32_t nxt = p.parentSlot[cur];
      if (nxt == kNone) break;
      cur = nxt;
      ++steps;
      ++lam;
      if (cur == tortoise) return false;
      if (lam == power) { tortoise = cur; power <<= 1; lam = 0; }
    }
  }
  outDirty = lastDirty >= 0;
  if (!needWorld && !outDirty) return true;
  if (!outDirty)
  {
    outW = load_world(p, ps);
    return true;
  }
  // pass 2: recompute ps's world from the topmost dirty ancestor down
  Mat4 W;
  for (int d = lastDirty; d >= 0; --d)
  {
    uint32_t node = ps;
    for (int k = 0; k < d; ++k) node = p.parentSlot[node];
    const Mat4 L = local_of(p, node);
    if (d == lastDirty)
    {
      const uint32_t up = p.parentSlot[node];
      W = (up == kNone) ? L : mat4_mul(load_world(p, up), L);
    }
    else
    {
      W = mat4_mul(W, L);
    }
  }
  outW = W;
  return true;
}

// ---- K1+K2: fused transform + cull, all views in one pass ------------------------------------------------
// One thread per slot, kSubTiles sub-tiles of kBlock consecutive slots per CTA.
// kHier=false: no instance has a parent (flat scene): pure streaming.
// kHier=true : parents inside the sub-tile are resolved level by level through shared memory (the parent's
//              fresh world matrix is staged there by its own thread); parents outside it by walk_up().
template <bool kHier>
__global__ void __launch_bounds__(kBlock) k_update(const __grid_constant__ UpdateParams p,
                                                   const __grid_constant__ ViewPlanes vp)
{
  __shared__ uint32_t sCounts[kMaxViews + 2];
  __shared__ float4 sW[kHier ? 4 : 1][kHier ? kBlock : 1];
  __shared__ uint8_t sState[kHier ? kBlock : 1];

  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  if (tid < kMaxViews + 2) sCounts[tid] = 0;
  __syncthreads();

  const uint32_t allMask = (1u << p.nViews) - 1u;
  uint32_t nRecomputed = 0;

  for (uint32_t sub = 0; sub < kSubTiles; ++sub)
  {
    const uint32_t base = blockIdx.x * kTile + sub * kBlock;
    if (base >= p.count) break;  // block-uniform
    const uint32_t i = base + tid;
    const bool live = i < p.count;

    float4 r2 = make_float4(0.f, 0.f, 0.f, 0.f), r3 = r2;
    uint32_t fl = 0;
    if (live)
    {
      r3 = ld_stream(p.rec3 + i);
      r2 = ld_stream(p.rec2 + i);
      fl = __float_as_uint(r3.w);
    }
    bool ownDirty = false;
    if (live && !(p.flags & kUpdSkipTransform))
      ownDirty = (p.flags & kUpdForceDirty) || ((fl >> kStampShift) == p.stamp);

    Mat4 W = mat4_identity();
    uint32_t ps = kNone;
    bool hier = false;
    if (kHier)
    {
      if (live) ps = p.parentSlot[i];
      hier = __syncthreads_or(ps != kNone) != 0;
    }

    if (!hier)
    {
      if (live)
      {
        if (ownDirty)
        {
          const float4 a = ld_stream(p.rec0 + i);
          const float4 b = ld_stream(p.rec1 + i);
          W = mat4_trs_dense(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, r2.x);
          store_world(p, i, W);
          ++nRecomputed;
        }
        else
        {
          W = load_world(p, i);
        }
      }
    }
    else if constexpr (kHier)
    {
      // sState: 0 pending, 1 done & clean, 2 done & recomputed, 3 dead (unreachable from any root)
      enum { PENDING = 0, DONE_CLEAN = 1, DONE_DIRTY = 2, DEAD = 3 };
      enum { SRC_ROOT = 0, SRC_WALKED = 1, SRC_TILE = 2 };
      int state = live ? PENDING : DEAD;
      int src = SRC_ROOT;
      Mat4 PW = mat4_identity();
      bool pDirtyWalked = false;
      if (live && ps != kNone)
      {
        if (ps - base < kBlock)
        {
          src = SRC_TILE;  // parent's thread is in this CTA: wait for it to publish through shared memory
        }
        else
        {
          src = SRC_WALKED;  // parent lives in another sub-tile: resolve it from global memory alone
          if (!walk_up(p, ps, ownDirty, PW, pDirtyWalked)) state = DEAD;
        }
      }
      sState[tid] = (uint8_t)state;
      for (;;)
      {
        __syncthreads();
        bool progressed = false;
        if (state == PENDING)
        {
          bool go = false, pDirty = false, hasParent = false;
          if (src == SRC_ROOT)
          {
            go = true;
          }
          else if (src == SRC_WALKED)
          {
            go = true; pDirty = pDirtyWalked; hasParent = true;
          }
          else
          {
            const uint32_t pt = ps - base;
            const int pst = sState[pt];
            if (pst == DONE_CLEAN || pst == DONE_DIRTY)
            {
              go = true; pDirty = (pst == DONE_DIRTY); hasParent = true;
              if (ownDirty || pDirty)
              {
                PW.c0 = sW[0][pt]; PW.c1 = sW[1][pt]; PW.c2 = sW[2][pt]; PW.c3 = sW[3][pt];
              }
            }
            else if (pst == DEAD)
            {
              state = DEAD;
              progressed = true;
            }
          }
          if (go)
          {
            const bool nodeDirty = ownDirty || pDirty;
            if (nodeDirty)
            {
              const Mat4 L = local_of(p, i);
              W = hasParent ? mat4_mul(PW, L) : L;
            }
            else
            {
              W = load_world(p, i);
            }
            state = nodeDirty ? DONE_DIRTY : DONE_CLEAN;
            progressed = true;
          }
        }
        const int any = __syncthreads_or(progressed);
        if (progressed)
        {
          sState[tid] = (uint8_t)state;
          if (state != DEAD)
          {
            sW[0][tid] = W.c0; sW[1][tid] = W.c1; sW[2][tid] = W.c2; sW[3][tid] = W.c3;
          }
        }
        if (!any) break;
      }
      if (state == DONE_DIRTY)
      {
        store_world(p, i, W);
        ++nRecomputed;
      }
      else if (live && state != DONE_CLEAN)
      {
        W = load_world(p, i);  // cycle members and their descendants: never visited by the DFS, world stays
      }
    }

    // ---- bounding sphere + 6*V plane tests in registers (CullingSystem, .cpp:1240-1270) ----
    uint32_t mask = 0;
    const bool cand = live && (fl & kFlagMesh);
    if (cand)
    {
      if ((p.flags & kUpdFreeze) || !(fl & kFlagBounds))
      {
        mask = allMask;
      }
      else
      {
        float cx, cy, cz, radius;
        world_bounds_sphere(W, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, cx, cy, cz, radius);
#pragma unroll 1
        for (uint32_t v = 0; v < p.nViews; ++v)
          if (sphere_in_frustum(vp.planes[v], cx, cy, cz, radius)) mask |= 1u << v;
      }
    }
    if (live) p.vismask[i] = (uint8_t)mask;

    // per-view visible counts of this tile (warp ballot + popc, one shared atomic per warp and view)
#pragma unroll 1
    for (uint32_t v = 0; v < p.nViews; ++v)
    {
      const uint32_t b = __ballot_sync(0xffffffffu, (mask >> v) & 1u);
      if (lane == 0 && b) atomicAdd(&sCounts[v], __popc(b));
    }
    {
      const uint32_t b = __ballot_sync(0xffffffffu, cand);
      if (lane == 0 && b) atomicAdd(&sCounts[p.nViews], __popc(b));
    }
  }

  // recomputed counter: warp reduce, one shared atomic per warp
  {
    uint32_t r = nRecomputed;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0 && r) atomicAdd(&sCounts[kMaxViews + 1], r);
  }
  __syncthreads();
  if (tid <= p.nViews) p.tileCounts[tid * p.numTiles + blockIdx.x] = sCounts[tid];
  if (tid == 0 && sCounts[kMaxViews + 1]) atomicAdd(p.recomputed, sCounts[kMaxViews + 1]);
}

// ---- K3a: exclusive scan of the per-tile counts, one CTA per row (view) -----------------------------------
__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t* __restrict__ tileCounts,
                                                     uint32_t* __restrict__ tileOffsets, uint32_t* __restrict__ totals,
                                                     uint32_t numTiles)
{
  __shared__ uint32_t sWarp[32];
  __shared__ uint32_t sCarry;
  const uint32_t row = blockIdx.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t* in = tileCounts + (size_t)row * numTiles;
  uint32_t* out = tileOffsets + (size_t)row * numTiles;
  if (tid == 0) sCarry = 0;
  __syncthreads();
  for (uint32_t start = 0; start < numTiles; start += 1024)
  {
    const uint32_t idx = start + tid;
    const uint32_t v = idx < numTiles ? in[idx] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((int)lane >= o) x += y;
    }
    if (lane == 31) sWarp[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
      uint32_t w = sWarp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if ((int)lane >= o) w += y;
      }
      sWarp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t carry = sCarry;
    const uint32_t warpExcl = warp ? sWarp[warp - 1] : 0u;
    if (idx < numTiles) out[idx] = carry + warpExcl + x - v;
    __syncthreads();
    if (tid == 1023) sCarry = carry + warpExcl + x;
    __syncthreads();
  }
  if (tid == 0) totals[row] = sCarry;
}

// ---- K3b: stable compaction of the visible sets (CullingState::visible, .cpp:1273-1280) --------------------
// One CTA per tile, thread t owns slots base+4t..base+4t+3 (their 4 mask bytes are one 32-bit load). Up to five
// views share one 64-bit block scan (12 bits per view, a tile holds at most 1024 instances).
struct ScatterParams
{
  const uint8_t* vismask;
  const uint32_t* entity;
  const uint32_t* tileCounts;
  const uint32_t* tileOffsets;
  uint32_t* outEntity[kMaxViews];
  uint32_t* outSlot[kMaxViews];
  uint32_t count;
  uint32_t numTiles;
  uint32_t nViews;
};

__device__ __forceinline__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t* sWarp, uint32_t tid)
{
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  uint64_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((int)lane >= o) x += y;
  }
  if (lane == 31) sWarp[warp] = x;
  __syncthreads();
  if (warp == 0)
  {
    uint64_t w = lane < (kBlock / 32) ? sWarp[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const uint64_t y = __shfl_up_sync(0xffffffffu, w, o);
      if ((int)lane >= o) w += y;
    }
    if (lane < (kBlock / 32)) sWarp[lane] = w;
  }
  __syncthreads();
  const uint64_t r = (warp ? sWarp[warp - 1] : 0ull) + x - v;
  __syncthreads();  // sWarp is reused by the caller's next scan
  return r;
}

__global__ void __launch_bounds__(kBlock) k_scatter_visible(const __grid_constant__ ScatterParams p)
{
  __shared__ uint32_t sCnt[kMaxViews];
  __shared__ uint64_t sWarp[kBlock / 32];
  const uint32_t tile = blockIdx.x, tid = threadIdx.x;
  if (tid < kMaxViews) sCnt[tid] = tid < p.nViews ? p.tileCounts[tid * p.numTiles + tile] : 0u;
  __syncthreads();
  uint32_t anyVis = 0;
#pragma unroll
  for (uint32_t v = 0; v < kMaxViews; ++v) anyVis |= sCnt[v];
  if (!anyVis) return;  // block-uniform: most tiles of an open world are fully culled

  const uint32_t slot0 = tile * kTile + tid * 4u;
  uint32_t m4 = 0;
  if (slot0 < p.count) m4 = reinterpret_cast<const uint32_t*>(p.vismask)[slot0 >> 2];
  // mask bytes of slots >= count are never written: drop them
  if (slot0 + 3u >= p.count)
  {
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k)
      if (slot0 + k >= p.count) m4 &= ~(0xFFu << (8u * k));
  }

  for (uint32_t v0 = 0; v0 < p.nViews; v0 += 5)
  {
    const uint32_t vEnd = min(p.nViews, v0 + 5u);
    uint64_t packed = 0;
    for (uint32_t v = v0; v < vEnd; ++v)
    {
      const uint32_t bits = (m4 >> v) & 0x01010101u;
      const uint32_t c = __popc(bits);
      packed |= (uint64_t)c << (12u * (v - v0));
    }
    const uint64_t excl = block_exclusive_scan_u64(packed, sWarp, tid);
    for (uint32_t v = v0; v < vEnd; ++v)
    {
      if (sCnt[v] == 0) continue;
      const uint32_t bits = (m4 >> v) & 0x01010101u;
      if (!bits) continue;
      uint32_t dst = p.tileOffsets[v * p.numTiles + tile] + (uint32_t)((excl >> (12u * (v - v0))) & 0xFFFu);
#pragma unroll
      for (uint32_t k = 0; k < 4; ++k)
      {
        if (bits & (1u << (8u * k)))
        {
          const uint32_t s = slot0 + k;
          p.outSlot[v][dst] = s;
          p.outEntity[v][dst] = p.entity[s];
          ++dst;
        }
      }
    }
  }
}

// culled lists (CullingState::culled): candidates whose view bit is clear. Offsets follow from the candidate and
// visible offsets, so no second scan is needed.
struct CulledParams
{
  const uint8_t* vismask;
  const float4* rec3;
  const uint32_t* entity;
  const uint32_t* tileOffsets;  // rows 0..nViews-1 visible, row nViews candidates
  uint32_t* outEntity[kMaxViews];
  uint32_t count;
  uint32_t numTiles;
  uint32_t nViews;
};

__global__ void __launch_bounds__(kBlock) k_scatter_culled(const __grid_constant__ CulledParams p)
{
  __shared__ uint64_t sWarp[kBlock / 32];
  const uint32_t tile = blockIdx.x, tid = threadIdx.x;
  const uint32_t slot0 = tile * kTile + tid * 4u;
  uint32_t m4 = 0, cand4 = 0;
  if (slot0 < p.count)
  {
    m4 = reinterpret_cast<const uint32_t*>(p.vismask)[slot0 >> 2];
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k)
      if (slot0 + k < p.count && (__float_as_uint(p.rec3[slot0 + k].w) & kFlagMesh)) cand4 |= 1u << (8u * k);
  }
  for (uint32_t v0 = 0; v0 < p.nViews; v0 += 5)
  {
    const uint32_t vEnd = min(p.nViews, v0 + 5u);
    uint64_t packed = 0;
    for (uint32_t v = v0; v < vEnd; ++v)
    {
      const uint32_t bits = cand4 & ~((m4 >> v) & 0x01010101u);
      packed |= (uint64_t)__popc(bits) << (12u * (v - v0));
    }
    const uint64_t excl = block_exclusive_scan_u64(packed, sWarp, tid);
    for (uint32_t v = v0; v < vEnd; ++v)
    {
      const uint32_t bits = cand4 & ~((m4 >> v) & 0x01010101u);
      if (!bits) continue;
      uint32_t dst = p.tileOffsets[p.nViews * p.numTiles + tile] - p.tileOffsets[v * p.numTiles + tile] +
                     (uint32_t)((excl >> (12u * (v - v0))) & 0xFFFu);
#pragma unroll
      for (uint32_t k = 0; k < 4; ++k)
        if (bits & (1u << (8u * k))) p.outEntity[v][dst++] = p.entity[slot0 + k];
    }
  }
}

// ---- K4: draw items (RenderPrepStreamingSystem, .cpp:1308-1328; sc::DrawItem 80 B, sc_ecs.h:159-165) ------
// 5 threads per item, one 16-byte chunk each, so a warp writes contiguous 16-byte pieces.
__global__ void __launch_bounds__(kBlock) k_build_draw_items(const uint32_t* __restrict__ visSlot,
                                                             const uint32_t* __restrict__ entity,
                                                             const uint2* __restrict__ meshMat,
                                                             const float4* __restrict__ w0, const float4* __restrict__ w1,
                                                             const float4* __restrict__ w2, const float4* __restrict__ w3,
                                                             uint32_t emitted, float4* __restrict__ out)
{
  const uint64_t g = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
  const uint64_t total = (uint64_t)emitted * 5ull;
  if (g >= total) return;
  const uint32_t item = (uint32_t)(g / 5ull), chunk = (uint32_t)(g % 5ull);
  const uint32_t s = visSlot[item];
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

// ---- K5: ECS deltas --------------------------------------------------------------------------------------

struct SceneArrays
{
  float4* rec[4];
  float4* world[4];
  uint32_t* parent;      // parent entity handle
  uint32_t* parentSlot;  // resolved slot or kNone
  uint32_t* entity;      // slot -> handle
  uint2* meshMat;
  uint32_t* sparse;      // Entity::index() -> slot+1 (ComponentPool sparse array, sc_ecs.h:199-277)
  uint32_t sparseSize;
};

// World::add<Transform> + setLocal (+Bounds/RenderMesh) for n new slots [slot0, slot0+n)
__global__ void __launch_bounds__(kBlock) k_spawn(SceneArrays a, uint32_t slot0, uint32_t n,
                                                  const uint32_t* __restrict__ entity, const uint32_t* __restrict__ parent,
                                                  const float* __restrict__ trs9, const float* __restrict__ aabb6,
                                                  const uint32_t* __restrict__ meshMat2, const uint32_t* __restrict__ flags,
                                                  uint32_t stamp)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t s = slot0 + j;
  const float* t = trs9 + (size_t)j * 9;
  float sx = t[6], sy = t[7], sz = t[8];
  // TransformSystem's zero-scale patch (sc_ecs.cpp:143-149); the instance is dirty anyway
  if (sx == 0.0f && sy == 0.0f && sz == 0.0f) { sx = sy = sz = 1.0f; }
  float bmin[3] = { -0.5f, -0.5f, -0.5f }, bmax[3] = { 0.5f, 0.5f, 0.5f };
  if (aabb6)
  {
    const float* b = aabb6 + (size_t)j * 6;
    bmin[0] = b[0]; bmin[1] = b[1]; bmin[2] = b[2]; bmax[0] = b[3]; bmax[1] = b[4]; bmax[2] = b[5];
  }
  const uint32_t f = (flags ? (flags[j] & 0xFFu) : (kFlagBounds | kFlagMesh)) | (stamp << kStampShift);
  a.rec[0][s] = make_float4(t[0], t[1], t[2], t[3]);
  a.rec[1][s] = make_float4(t[4], t[5], sx, sy);
  a.rec[2][s] = make_float4(sz, bmin[0], bmin[1], bmin[2]);
  a.rec[3][s] = make_float4(bmax[0], bmax[1], bmax[2], __uint_as_float(f));
  a.world[0][s] = make_float4(1.f, 0.f, 0.f, 0.f);
  a.world[1][s] = make_float4(0.f, 1.f, 0.f, 0.f);
  a.world[2][s] = make_float4(0.f, 0.f, 1.f, 0.f);
  a.world[3][s] = make_float4(0.f, 0.f, 0.f, 1.f);
  const uint32_t e = entity[j];
  a.entity[s] = e;
  a.parent[s] = parent ? parent[j] : kNone;
  a.parentSlot[s] = kNone;
  a.meshMat[s] = meshMat2 ? make_uint2(meshMat2[(size_t)j * 2], meshMat2[(size_t)j * 2 + 1]) : make_uint2(0u, 0u);
  const uint32_t idx = e & 0xFFFFFFu;
  if (idx < a.sparseSize) a.sparse[idx] = s + 1u;
}

__device__ __forceinline__ uint32_t find_slot(const SceneArrays& a, uint32_t handle)
{
  if (handle == kNone) return kNone;
  const uint32_t idx = handle & 0xFFFFFFu;
  if (idx >= a.sparseSize) return kNone;
  const uint32_t s = a.sparse[idx];
  if (s == 0u || a.entity[s - 1u] != handle) return kNone;
  return s - 1u;
}

// setLocal (sc_ecs.h:78-84) for n entities; unknown handles are skipped
__global__ void __launch_bounds__(kBlock) k_set_local(SceneArrays a, uint32_t n, const uint32_t* __restrict__ entity,
                                                      const float* __restrict__ trs9, uint32_t stamp)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t s = find_slot(a, entity[j]);
  if (s == kNone) return;
  const float* t = trs9 + (size_t)j * 9;
  float sx = t[6], sy = t[7], sz = t[8];
  if (sx == 0.0f && sy == 0.0f && sz == 0.0f) { sx = sy = sz = 1.0f; }
  a.rec[0][s] = make_float4(t[0], t[1], t[2], t[3]);
  a.rec[1][s] = make_float4(t[4], t[5], sx, sy);
  reinterpret_cast<float*>(a.rec[2] + s)[0] = sz;
  uint32_t* fw = reinterpret_cast<uint32_t*>(a.rec[3] + s) + 3;
  *fw = (*fw & 0xFFu) | (stamp << kStampShift);
}

__global__ void __launch_bounds__(kBlock) k_mark_dirty(SceneArrays a, uint32_t n, const uint32_t* __restrict__ entity,
                                                       uint32_t stamp)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t s = find_slot(a, entity[j]);
  if (s == kNone) return;
  uint32_t* fw = reinterpret_cast<uint32_t*>(a.rec[3] + s) + 3;
  *fw = (*fw & 0xFFu) | (stamp << kStampShift);
}

// setParent (sc_ecs.h:86-90): stores the handle, marks dirty; validity is judged by k_resolve_parents
__global__ void __launch_bounds__(kBlock) k_set_parent(SceneArrays a, uint32_t n, const uint32_t* __restrict__ entity,
                                                       const uint32_t* __restrict__ parent, uint32_t stamp)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t s = find_slot(a, entity[j]);
  if (s == kNone) return;
  a.parent[s] = parent[j];
  uint32_t* fw = reinterpret_cast<uint32_t*>(a.rec[3] + s) + 3;
  *fw = (*fw & 0xFFu) | (stamp << kStampShift);
}

// World::destroy batches: the host replays ComponentPool::remove's swap-with-last (sc_ecs.h:240-262) on its
// entity mirror and hands over the net result: slots to fill (dst <- src, src always in the vacated tail, so
// sources and destinations never overlap) and the sparse entries to clear.
__global__ void __launch_bounds__(kBlock) k_despawn_apply(SceneArrays a, uint32_t nMoves, const uint2* __restrict__ moves,
                                                          uint32_t nRemoved, const uint32_t* __restrict__ removedIndex)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j < nMoves)
  {
    const uint32_t dst = moves[j].x, src = moves[j].y;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      a.rec[k][dst] = a.rec[k][src];
      a.world[k][dst] = a.world[k][src];
    }
    const uint32_t e = a.entity[src];
    a.entity[dst] = e;
    a.parent[dst] = a.parent[src];
    a.meshMat[dst] = a.meshMat[src];
    a.sparse[e & 0xFFFFFFu] = dst + 1u;
  }
  else if (j < nMoves + nRemoved)
  {
    a.sparse[removedIndex[j - nMoves]] = 0u;
  }
}

// TransformSystem's per-frame parent validation (sc_ecs.cpp:151-164), run only when the topology changed:
// a parent is valid iff it is not the entity itself and the pool holds exactly that handle (alive && has
// Transform); otherwise the node becomes a root and, if it had a parent handle, dirty.
__global__ void __launch_bounds__(kBlock) k_resolve_parents(SceneArrays a, uint32_t count, uint32_t stamp)
{
  const uint32_t s = blockIdx.x * kBlock + threadIdx.x;
  if (s >= count) return;
  const uint32_t ph = a.parent[s];
  uint32_t ps = kNone;
  if (ph != kNone)
  {
    if (ph != a.entity[s]) ps = find_slot(a, ph);
    if (ps == kNone)
    {
      a.parent[s] = kNone;
      uint32_t* fw = reinterpret_cast<uint32_t*>(a.rec[3] + s) + 3;
      *fw = (*fw & 0xFFu) | (stamp << kStampShift);
    }
  }
  a.parentSlot[s] = ps;
}

// gathers for read-back by entity handle
__global__ void __launch_bounds__(kBlock) k_gather_world(SceneArrays a, uint32_t n, const uint32_t* __restrict__ entity,
                                                         float4* __restrict__ out, uint32_t* __restrict__ missing)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t s = find_slot(a, entity[j]);
  if (s == kNone)
  {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    out[j * 4 + 0] = z; out[j * 4 + 1] = z; out[j * 4 + 2] = z; out[j * 4 + 3] = z;
    atomicAdd(missing, 1u);
    return;
  }
  out[j * 4 + 0] = a.world[0][s]; out[j * 4 + 1] = a.world[1][s];
  out[j * 4 + 2] = a.world[2][s]; out[j * 4 + 3] = a.world[3][s];
}

__global__ void __launch_bounds__(kBlock) k_gather_parent(SceneArrays a, uint32_t n, const uint32_t* __restrict__ entity,
                                                          uint32_t* __restrict__ out)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t s = find_slot(a, entity[j]);
  out[j] = (s == kNone) ? kNone : a.parent[s];
}

}  // namespace scgpu
