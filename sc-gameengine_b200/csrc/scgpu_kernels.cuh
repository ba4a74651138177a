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
//   slotInfo[slot]   = depth of the slot inside its hierarchy window (+ flags), winStart[tile][..] = window starts:
//                      hierarchical scenes are cut into windows of <= 32 consecutive slots that no parent link
//                      crosses; one warp resolves one window with shuffles (k_build_windows / k_update_win).
//   flags: bit0 HAS_BOUNDS, bit1 HAS_MESH, bits 8..31 = dirty stamp (id of the update that must recompute
//          the instance). A stamp instead of a dirty bit means the frame kernel never writes the records.
//
// Frame = k_update_flat | k_update_win (transform + sphere + V-view plane tests + per-tile counts, one pass)
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
constexpr uint32_t kUpdateSmemFlat = 3 * 4 * kBlock * 16;  // dynamic shared memory of k_update_flat
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

// ---- out-of-line slow paths --------------------------------------------------------------------------------
// The rare paths (dense fallbacks for non-finite input, ancestor walks) are real function calls. They exchange
// matrices with the caller through a per-thread exchange slot (4 consecutive float4) instead of through
// reference parameters, so that no matrix of the hot path ever has its address taken (which would pin it to
// local memory).
__device__ __forceinline__ Mat4 xs_load(const float4* x)
{
  Mat4 m;
  m.c0 = x[0]; m.c1 = x[1]; m.c2 = x[2]; m.c3 = x[3];
  return m;
}

__device__ __forceinline__ void xs_store(float4* x, const Mat4& m)
{
  x[0] = m.c0; x[1] = m.c1; x[2] = m.c2; x[3] = m.c3;
}

__device__ __noinline__ void trs_dense_to(float px, float py, float pz, float rx, float ry, float rz, float sx, float sy,
                                          float sz, float4* out)
{
  xs_store(out, mat4_trs_dense(px, py, pz, rx, ry, rz, sx, sy, sz));
}

// out = a * b, dense (sc_math.cpp:52-68); out may alias a or b
__device__ __noinline__ void mul_dense_to(const float4* a, const float4* b, float4* out)
{
  const Mat4 A = xs_load(a), B = xs_load(b);
  xs_store(out, mat4_mul(A, B));
}

// mat4_trs (sc_math.cpp:130-142): structured fast path where it is value-exact, dense call otherwise
__device__ __forceinline__ Mat4 trs_any(float4 a, float4 b, float sz, bool& affine, float4* scratch)
{
  affine = trs_inputs_tame(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, sz);
  if (affine) return mat4_trs_fast(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, sz);
  trs_dense_to(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, sz, scratch);
  return xs_load(scratch);
}

// parent.world * local (sc_ecs.cpp:191-195). scratchA/scratchB: two exchange slots owned by this thread.
__device__ __forceinline__ Mat4 compose_any(const Mat4& pw, const Mat4& l, bool localAffine, float4* scratchA,
                                            float4* scratchB)
{
  const float mag = fabsf(pw.c3.x) + fabsf(pw.c3.y) + fabsf(pw.c3.z) + fabsf(pw.c3.w);
  if (localAffine && mag < __int_as_float(0x7f800000)) return mat4_mul_affine(pw, l);
  xs_store(scratchA, pw);
  xs_store(scratchB, l);
  mul_dense_to(scratchA, scratchB, scratchA);
  return xs_load(scratchA);
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
// down. The parent's world matrix is left in the exchange slot `out`.
// Returns bit0 = a root is reachable (false: cycle, such nodes are never visited by the DFS), bit1 = parent dirty.
__device__ __noinline__ uint32_t walk_up(const UpdateParams& p, uint32_t ps, bool needWorld, float4* out, float4* tmp)
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
      if (cur == tortoise) return 0u;
      if (lam == power) { tortoise = cur; power <<= 1; lam = 0; }
    }
  }
  const bool dirty = lastDirty >= 0;
  if (!dirty)
  {
    if (needWorld) { out[0] = p.w0[ps]; out[1] = p.w1[ps]; out[2] = p.w2[ps]; out[3] = p.w3[ps]; }
    return 1u;
  }
  // pass 2: recompute ps's world from the topmost dirty ancestor down
  Mat4 W = mat4_identity();
  for (int d = lastDirty; d >= 0; --d)
  {
    uint32_t node = ps;
    for (int k = 0; k < d; ++k) node = p.parentSlot[node];
    bool affine;
    const Mat4 L = trs_any(p.rec0[node], p.rec1[node], p.rec2[node].x, affine, tmp);
    if (d == lastDirty)
    {
      const uint32_t up = p.parentSlot[node];
      if (up == kNone) W = L;
      else
      {
        Mat4 U;
        U.c0 = p.w0[up]; U.c1 = p.w1[up]; U.c2 = p.w2[up]; U.c3 = p.w3[up];
        W = compose_any(U, L, affine, out, tmp);
      }
    }
    else
    {
      W = compose_any(W, L, affine, out, tmp);
    }
  }
  xs_store(out, W);
  return 3u;
}

// ---- TMA bulk copies (cp.async.bulk + mbarrier): stage a sub-tile's planes in shared memory ahead of use ------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t done;
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// signed plane distance in the reference's order: ((n0*c0 + n1*c1) + n2*c2) + d
__device__ __forceinline__ float plane_dist(const float4 pl, float cx, float cy, float cz)
{
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pl.x, cx), __fmul_rn(pl.y, cy)), __fmul_rn(pl.z, cz)), pl.w);
}

// ---- K1+K2, flat scenes: fused transform + cull, all views in one pass ------------------------------------------------
// No instance has a parent: pure streaming. One thread per slot, kSubTiles sub-tiles of kBlock consecutive slots per
// CTA; the four record planes of a sub-tile are staged in shared memory by TMA bulk copies two sub-tiles ahead, so the
// DRAM round trip overlaps the arithmetic. kViews is a compile-time view count, so the 6*V plane tests read their
// planes straight from the constant bank; views are tested in plane pairs with a warp-uniform early-out.
template <int kViews>
__global__ void __launch_bounds__(kBlock, 4) k_update_flat(const __grid_constant__ UpdateParams p,
                                                           const __grid_constant__ ViewPlanes vp)
{
  __shared__ uint32_t sCounts[kMaxViews + 2];
  __shared__ __align__(8) uint64_t sFull[2];
  // dynamic shared memory (kUpdateSmemFlat bytes, opted in by the host):
  //   sRec[2][4][kBlock]  TMA staging, double buffered
  //   sX[kBlock][4]       per-thread exchange slot of the out-of-line dense fallback
  extern __shared__ __align__(128) unsigned char sDyn[];
  float4(*sRec)[4][kBlock] = reinterpret_cast<float4(*)[4][kBlock]>(sDyn);
  float4(*sX)[4] = reinterpret_cast<float4(*)[4]>(sDyn + 2 * 4 * kBlock * sizeof(float4));

  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  if (tid < kMaxViews + 2) sCounts[tid] = 0;
  const uint32_t tileBase = blockIdx.x * kTile;
  const uint32_t nSub = min(kSubTiles, (p.count - tileBase + kBlock - 1) / kBlock);
  // the arrays are padded to a multiple of kTile, so a partial last sub-tile is still copied whole
  auto stage = [&](uint32_t sub)
  {
    const uint32_t b = sub & 1u, base = tileBase + sub * kBlock;
    constexpr uint32_t kPlane = kBlock * sizeof(float4);
    mbar_expect_tx(&sFull[b], 4u * kPlane);
    bulk_g2s(&sRec[b][0][0], p.rec0 + base, kPlane, &sFull[b]);
    bulk_g2s(&sRec[b][1][0], p.rec1 + base, kPlane, &sFull[b]);
    bulk_g2s(&sRec[b][2][0], p.rec2 + base, kPlane, &sFull[b]);
    bulk_g2s(&sRec[b][3][0], p.rec3 + base, kPlane, &sFull[b]);
  };
  if (tid == 0)
  {
    mbar_init(&sFull[0], 1);
    mbar_init(&sFull[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
  {
    stage(0);
    if (nSub > 1) stage(1);
  }

  constexpr uint32_t allMask = (1u << kViews) - 1u;
  const bool skip = (p.flags & kUpdSkipTransform) != 0;
  const bool force = (p.flags & kUpdForceDirty) != 0;
  const bool freeze = (p.flags & kUpdFreeze) != 0;
  uint32_t nRecomputed = 0, nCand = 0;
  uint32_t nVis[kViews];
#pragma unroll
  for (int v = 0; v < kViews; ++v) nVis[v] = 0;

#pragma unroll 1
  for (uint32_t sub = 0; sub < nSub; ++sub)
  {
    const uint32_t buf = sub & 1u;
    const uint32_t s = tileBase + sub * kBlock + tid;
    const bool live = s < p.count;
    mbar_wait(&sFull[buf], (sub >> 1) & 1u);  // this sub-tile's planes have landed in shared memory

    float4 r0, r1, r2, r3;
    r0 = r1 = r2 = r3 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live)
    {
      r3 = sRec[buf][3][tid]; r2 = sRec[buf][2][tid];
      r0 = sRec[buf][0][tid]; r1 = sRec[buf][1][tid];
    }
    __syncthreads();  // everybody has copied its record out of the staging buffer: refill it
    if (tid == 0 && sub + 2 < nSub) stage(sub + 2);

    const uint32_t fl = __float_as_uint(r3.w);
    const bool ownDirty = live && !skip && (force || ((fl >> kStampShift) == p.stamp));
    Mat4 W = mat4_identity();
    if (ownDirty)
    {
      bool affine;
      W = trs_any(r0, r1, r2.x, affine, &sX[tid][0]);
      store_world(p, s, W);
      ++nRecomputed;
    }
    else if (live)
    {
      W = load_world(p, s);
    }

    // ---- bounding sphere + 6*V plane tests in registers (CullingSystem, .cpp:1240-1270) ----
    const bool cand = live && (fl & kFlagMesh);
    const bool test = cand && !freeze && (fl & kFlagBounds);
    uint32_t mask = 0;
    if (__any_sync(0xffffffffu, test))
    {
      float cx, cy, cz, radius;
      world_bounds_sphere(W, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, cx, cy, cz, radius);
      const float negR = -radius;
#pragma unroll
      for (int v = 0; v < kViews; ++v)
      {
        bool alive = test;
        // plane pairs (left,right) (bottom,top) (near,far); stop as soon as the whole warp is outside
#pragma unroll
        for (int pp = 0; pp < 3; ++pp)
        {
          if (!__any_sync(0xffffffffu, alive)) break;
          const float d0 = plane_dist(vp.planes[v][2 * pp], cx, cy, cz);
          const float d1 = plane_dist(vp.planes[v][2 * pp + 1], cx, cy, cz);
          alive = alive && !(d0 < negR) && !(d1 < negR);  // NaN compares false => stays visible
        }
        if (alive) mask |= 1u << v;
      }
    }
    if (cand && !test) mask = allMask;  // frozen culling or no Bounds component: always visible
    if (live) p.vismask[s] = (uint8_t)mask;
#pragma unroll
    for (int v = 0; v < kViews; ++v) nVis[v] += (mask >> v) & 1u;
    nCand += cand ? 1u : 0u;
  }

  // per-tile counts: one warp reduction (REDUX) per counter, one shared atomic per warp
#pragma unroll
  for (int v = 0; v < kViews; ++v)
  {
    const uint32_t r = __reduce_add_sync(0xffffffffu, nVis[v]);
    if (lane == 0 && r) atomicAdd(&sCounts[v], r);
  }
  {
    const uint32_t r = __reduce_add_sync(0xffffffffu, nCand);
    if (lane == 0 && r) atomicAdd(&sCounts[kViews], r);
    const uint32_t q = __reduce_add_sync(0xffffffffu, nRecomputed);
    if (lane == 0 && q) atomicAdd(&sCounts[kMaxViews + 1], q);
  }
  __syncthreads();
  if (tid <= (uint32_t)kViews) p.tileCounts[tid * p.numTiles + blockIdx.x] = sCounts[tid];
  if (tid == 0 && sCounts[kMaxViews + 1]) atomicAdd(p.recomputed, sCounts[kMaxViews + 1]);
}

// ---- hierarchy windows ------------------------------------------------------------------------------------------
// A window is a run of <= 32 consecutive slots owned by one warp. k_build_windows (topology changes only) cuts every
// kTile-slot tile into windows at positions that NO parent link crosses, so a hierarchy group never straddles two
// warps and k_update_win resolves it with shuffles alone. Per slot it records the depth inside the window
// (0 = root or parent outside the window), per tile the window starts.
constexpr uint32_t kMaxWin = 128;          // windows per tile (greedy cuts give ~36; forced 32-slot cuts bound it)
constexpr uint32_t kWinUnreachable = 0x40; // slotInfo: node on / below a cycle closed inside its window
constexpr uint32_t kWinExternal = 0x20;    // slotInfo: parent lives outside the window (resolved by walk_up)

__global__ void __launch_bounds__(kBlock) k_build_windows(const uint32_t* __restrict__ parentSlot, uint8_t* __restrict__ slotInfo,
                                                          uint16_t* __restrict__ winStart, uint32_t count)
{
  __shared__ int sCross[kTile + 2];
  __shared__ uint16_t sStart[kMaxWin + 1];
  __shared__ uint16_t sWinOf[kTile];  // window start (tile-relative) of every slot
  __shared__ uint32_t sWarpSum[kBlock / 32];
  __shared__ uint32_t sNumWin;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t tileBase = blockIdx.x * kTile;
  const uint32_t n = min(kTile, count - tileBase);

  for (uint32_t k = tid; k < kTile + 2; k += kBlock) sCross[k] = 0;
  __syncthreads();
  // a link between slots lo < hi (both in this tile) crosses every cut position c with lo < c <= hi
  for (uint32_t k = tid; k < n; k += kBlock)
  {
    const uint32_t ps = parentSlot[tileBase + k];
    if (ps != kNone && ps >= tileBase && ps < tileBase + n && ps != tileBase + k)
    {
      const uint32_t q = ps - tileBase, lo = min(k, q), hi = max(k, q);
      atomicAdd(&sCross[lo + 1], 1);
      atomicAdd(&sCross[hi + 1], -1);
    }
  }
  __syncthreads();
  // inclusive prefix sum over kTile+1 positions: thread t owns positions 4t..4t+3 (+ the last one)
  {
    int v[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { sum += sCross[tid * 4 + j]; v[j] = sum; }
    int x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((int)lane >= o) x += y;
    }
    if (lane == 31) sWarpSum[warp] = (uint32_t)x;
    __syncthreads();
    int off = x - sum;
    for (uint32_t w = 0; w < warp; ++w) off += (int)sWarpSum[w];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) sCross[tid * 4 + j] = v[j] + off;
    if (tid == kBlock - 1) sCross[kTile] += v[3] + off;
  }
  __syncthreads();
  // greedy cuts by warp 0: the next window ends at the LARGEST uncrossed position within 32 slots (forced at 32)
  if (warp == 0)
  {
    uint32_t start = 0, nw = 0;
    while (start < n)
    {
      if (lane == 0) sStart[nw] = (uint16_t)start;
      ++nw;
      const uint32_t c = start + 1 + lane;  // candidate end
      const bool ok = c <= n && (c == n || sCross[c] == 0);
      uint32_t m = __ballot_sync(0xffffffffu, ok);
      if (nw >= kMaxWin - 32) m = 0;  // too many small windows: finish with forced cuts so that nw <= kMaxWin
      start = m ? start + 32u - __clz(m) : min(start + 32u, n);
    }
    if (lane == 0) { sStart[nw] = (uint16_t)n; sNumWin = nw; }
  }
  __syncthreads();
  const uint32_t nw = sNumWin;
  for (uint32_t k = tid; k < kMaxWin + 2; k += kBlock)
    winStart[(size_t)blockIdx.x * (kMaxWin + 2) + k] = (k == kMaxWin + 1) ? (uint16_t)nw : (k <= nw ? sStart[k] : (uint16_t)0xFFFF);
  for (uint32_t w = warp; w < nw; w += kBlock / 32)
    for (uint32_t k = sStart[w] + lane; k < sStart[w + 1]; k += 32) sWinOf[k] = sStart[w];
  __syncthreads();
  // depth inside the window (0: root, or parent outside the window => kWinExternal)
  for (uint32_t k = tid; k < n; k += kBlock)
  {
    uint32_t cur = k, depth = 0, info = 0;
    for (;;)
    {
      const uint32_t ps = parentSlot[tileBase + cur];
      if (ps == kNone) break;
      const bool inside = ps >= tileBase && (ps - tileBase) < n && sWinOf[ps - tileBase] == sWinOf[k];
      if (!inside)
      {
        if (depth == 0) info = kWinExternal;  // deeper nodes hang off an ancestor that carries the flag itself
        break;
      }
      cur = ps - tileBase;
      if (++depth > 32u) { info = kWinUnreachable; depth = 0; break; }  // a cycle closed inside the window
    }
    slotInfo[tileBase + k] = (uint8_t)(info | (depth & 31u));
  }
}

// ---- K1+K2, hierarchical scenes: one window per warp, levels resolved with shuffles ----------------------------
// Loads and stores are the flat kernel's (coalesced 128-bit planes, slot = window start + lane). Inside the warp:
//   1. dirty / never-visited bits are propagated parent -> child level by level (one shuffle per level),
//   2. every lane that must be recomputed builds its local matrix (all lanes busy, FP64 sincos in parallel),
//   3. for level 1..max, the lanes of that level fetch their parent's world matrix from the parent's lane and
//      multiply: parent-before-child, parent matrices staged in registers of the same warp.
// No shared memory traffic, no CTA barrier, no dependency between warps: they free-run like in the flat kernel.
constexpr uint32_t kWinTilesPerCta = 4;  // tiles per CTA of k_update_win: ~146 windows shared by 8 warps

template <int kViews>
__global__ void __launch_bounds__(kBlock, 4) k_update_win(const __grid_constant__ UpdateParams p,
                                                          const __grid_constant__ ViewPlanes vp,
                                                          const uint8_t* __restrict__ slotInfo,
                                                          const uint16_t* __restrict__ winStart)
{
  __shared__ uint16_t sStart[kWinTilesPerCta][kMaxWin + 2];
  __shared__ uint32_t sNext;  // next unclaimed (tile, window) of this CTA: warps pull work dynamically
  // per-warp double buffer: the record planes of the NEXT claimed window are fetched with cp.async while the
  // current window is being computed
  __shared__ __align__(16) float4 sPre[kBlock / 32][2][4][32];

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t firstTile = blockIdx.x * kWinTilesPerCta;
  const uint32_t nTiles = min(kWinTilesPerCta, p.numTiles - firstTile);
  for (uint32_t k = tid; k < nTiles * (kMaxWin + 2); k += kBlock)
    (&sStart[0][0])[k] = winStart[(size_t)firstTile * (kMaxWin + 2) + k];
  if (tid == 0) sNext = 0;
  __syncthreads();

  constexpr uint32_t allMask = (1u << kViews) - 1u;
  const bool skip = (p.flags & kUpdSkipTransform) != 0;
  const bool force = (p.flags & kUpdForceDirty) != 0;
  const bool freeze = (p.flags & kUpdFreeze) != 0;
  // exchange slots of the out-of-line slow paths: thread-local, touched only when a slow path runs
  float4 xaBuf[4], xbBuf[4];
  float4* const xa = xaBuf;
  float4* const xb = xbBuf;
  uint32_t nRecomputed = 0;

  // No barrier below this line: a warp that finishes a window claims another one, whichever tile it belongs to.
  uint32_t t = 0, wBase = 0;  // claimed indices only grow, so the tile cursor moves forward
  // claim + locate + start fetching one window; returns false when the CTA's windows are exhausted
  struct Win { uint32_t tile, a, len, info, ps; };
  auto claim_and_fetch = [&](uint32_t buf, Win& o) -> bool
  {
    uint32_t claim = 0;
    if (lane == 0) claim = atomicAdd(&sNext, 1u);
    claim = __shfl_sync(0xffffffffu, claim, 0);
    while (t < nTiles && claim >= wBase + sStart[t][kMaxWin + 1]) { wBase += sStart[t][kMaxWin + 1]; ++t; }
    if (t >= nTiles) return false;
    const uint32_t w = claim - wBase;
    o.tile = firstTile + t;
    o.a = o.tile * kTile + sStart[t][w];
    o.len = sStart[t][w + 1] - sStart[t][w];
    o.info = kWinUnreachable;
    o.ps = kNone;
    if (lane < o.len)
    {
      const uint32_t q = o.a + lane;
      cp_async16(&sPre[warp][buf][0][lane], p.rec0 + q);
      cp_async16(&sPre[warp][buf][1][lane], p.rec1 + q);
      cp_async16(&sPre[warp][buf][2][lane], p.rec2 + q);
      cp_async16(&sPre[warp][buf][3][lane], p.rec3 + q);
      o.info = slotInfo[q];
      o.ps = p.parentSlot[q];
    }
    cp_async_commit();
    return true;
  };
  Win cur, nxt;
  uint32_t buf = 0;
  bool have = claim_and_fetch(0, cur);
#pragma unroll 1
  while (have)
  {
    const bool haveNext = claim_and_fetch(buf ^ 1u, nxt);  // in flight while this window is computed
    if (haveNext) cp_async_wait<1>();
    else cp_async_wait<0>();
    const uint32_t tile = cur.tile;
    uint32_t nCand = 0;
    uint32_t nVis[kViews];
#pragma unroll
    for (int v = 0; v < kViews; ++v) nVis[v] = 0;
  {
    const uint32_t a = cur.a;
    const uint32_t len = cur.len;
    const bool live = lane < len;
    const uint32_t s = a + lane;

    float4 r0, r1, r2, r3;
    r0 = r1 = r2 = r3 = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t info = cur.info, ps = cur.ps;
    if (live)
    {
      r3 = sPre[warp][buf][3][lane]; r2 = sPre[warp][buf][2][lane];
      r0 = sPre[warp][buf][0][lane]; r1 = sPre[warp][buf][1][lane];
    }
    const uint32_t fl = __float_as_uint(r3.w);
    const bool ownDirty = live && !skip && (force || ((fl >> kStampShift) == p.stamp));
    const uint32_t wl = info & 31u;
    const bool external = (info & kWinExternal) != 0;
    bool dead = !live || (info & kWinUnreachable);  // never visited by the DFS: world matrix stays as stored
    const uint32_t parentLane = (wl != 0) ? (ps - a) & 31u : lane;
    const uint32_t maxL = __reduce_max_sync(0xffffffffu, live ? wl : 0u);

    // ---- parents outside the window: resolved from global memory alone (walk_up leaves the matrix in xa) ----
    bool nodeDirty = ownDirty;
    if (live && external)
    {
      const uint32_t walk = walk_up(p, ps, true, xa, xb);
      if (!(walk & 1u)) dead = true;
      nodeDirty = ownDirty || (walk & 2u) != 0;
    }
    // ---- 1. inherit dirtiness / deadness down the levels ----
    for (uint32_t l = 1; l <= maxL; ++l)
    {
      const bool pd = __shfl_sync(0xffffffffu, nodeDirty ? 1 : 0, parentLane) != 0;
      const bool pdead = __shfl_sync(0xffffffffu, dead ? 1 : 0, parentLane) != 0;
      if (live && wl == l)
      {
        nodeDirty = nodeDirty || pd;
        dead = dead || pdead;
      }
    }
    if (dead) nodeDirty = false;

    // ---- 2. local matrices of everything that is recomputed; stored world matrices of the rest ----
    Mat4 W = mat4_identity();
    bool affine = true;
    if (nodeDirty)
    {
      W = trs_any(r0, r1, r2.x, affine, xb);  // roots: world == local
      if (external) W = compose_any(xs_load(xa), W, affine, xa, xb);
    }
    else if (live)
    {
      W = load_world(p, s);
    }

    // ---- 3. parent.world * local, one level at a time, parents read from their lanes ----
    // Common case: every matrix involved is affine (bottom row (0,0,0,1)) with a finite translation, so only the
    // upper 3x4 travels through the shuffles and the product skips the bottom row. Anything else: full 4x4.
    bool wAff = live && (nodeDirty ? (affine && !external) : mat4_is_affine(W));
    if (nodeDirty && external) wAff = mat4_is_affine(W);
    for (uint32_t l = 1; l <= maxL; ++l)
    {
      const bool mine = nodeDirty && wl == l;
      const float mag = fabsf(W.c3.x) + fabsf(W.c3.y) + fabsf(W.c3.z);
      const bool parentOk = __shfl_sync(0xffffffffu, (wAff && mag < __int_as_float(0x7f800000)) ? 1 : 0, parentLane) != 0;
      Mat4 PW;
      PW.c0.x = __shfl_sync(0xffffffffu, W.c0.x, parentLane); PW.c0.y = __shfl_sync(0xffffffffu, W.c0.y, parentLane);
      PW.c0.z = __shfl_sync(0xffffffffu, W.c0.z, parentLane);
      PW.c1.x = __shfl_sync(0xffffffffu, W.c1.x, parentLane); PW.c1.y = __shfl_sync(0xffffffffu, W.c1.y, parentLane);
      PW.c1.z = __shfl_sync(0xffffffffu, W.c1.z, parentLane);
      PW.c2.x = __shfl_sync(0xffffffffu, W.c2.x, parentLane); PW.c2.y = __shfl_sync(0xffffffffu, W.c2.y, parentLane);
      PW.c2.z = __shfl_sync(0xffffffffu, W.c2.z, parentLane);
      PW.c3.x = __shfl_sync(0xffffffffu, W.c3.x, parentLane); PW.c3.y = __shfl_sync(0xffffffffu, W.c3.y, parentLane);
      PW.c3.z = __shfl_sync(0xffffffffu, W.c3.z, parentLane);
      if (__all_sync(0xffffffffu, !mine || (parentOk && affine)))
      {
        if (mine) W = mat4_mul_affine3(PW, W);  // wAff stays true
      }
      else
      {
        PW.c0.w = __shfl_sync(0xffffffffu, W.c0.w, parentLane); PW.c1.w = __shfl_sync(0xffffffffu, W.c1.w, parentLane);
        PW.c2.w = __shfl_sync(0xffffffffu, W.c2.w, parentLane); PW.c3.w = __shfl_sync(0xffffffffu, W.c3.w, parentLane);
        if (mine)
        {
          W = compose_any(PW, W, affine, xa, xb);
          wAff = mat4_is_affine(W);
        }
      }
    }
    if (nodeDirty)
    {
      store_world(p, s, W);
      ++nRecomputed;
    }

    // ---- bounding sphere + 6*V plane tests in registers (CullingSystem, .cpp:1240-1270) ----
    const bool cand = live && (fl & kFlagMesh);
    const bool test = cand && !freeze && (fl & kFlagBounds);
    uint32_t mask = 0;
    if (__any_sync(0xffffffffu, test))
    {
      float cx, cy, cz, radius;
      world_bounds_sphere(W, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, cx, cy, cz, radius);
      const float negR = -radius;
#pragma unroll
      for (int v = 0; v < kViews; ++v)
      {
        bool alive = test;
#pragma unroll
        for (int pp = 0; pp < 3; ++pp)
        {
          if (!__any_sync(0xffffffffu, alive)) break;
          const float d0 = plane_dist(vp.planes[v][2 * pp], cx, cy, cz);
          const float d1 = plane_dist(vp.planes[v][2 * pp + 1], cx, cy, cz);
          alive = alive && !(d0 < negR) && !(d1 < negR);  // NaN compares false => stays visible
        }
        if (alive) mask |= 1u << v;
      }
    }
    if (cand && !test) mask = allMask;
    if (live) p.vismask[s] = (uint8_t)mask;
#pragma unroll
    for (int v = 0; v < kViews; ++v) nVis[v] += (mask >> v) & 1u;
    nCand += cand ? 1u : 0u;
  }
    // per-tile counts (zeroed by the host before the launch): ballots, one global atomic per warp and counter
#pragma unroll
    for (int v = 0; v < kViews; ++v)
    {
      const uint32_t c = __reduce_add_sync(0xffffffffu, nVis[v]);
      if (lane == 0 && c) atomicAdd(&p.tileCounts[v * p.numTiles + tile], c);
    }
    {
      const uint32_t c = __reduce_add_sync(0xffffffffu, nCand);
      if (lane == 0 && c) atomicAdd(&p.tileCounts[kViews * p.numTiles + tile], c);
    }
    have = haveNext;
    cur = nxt;
    buf ^= 1u;
  }
  const uint32_t q = __reduce_add_sync(0xffffffffu, nRecomputed);
  if (lane == 0 && q) atomicAdd(p.recomputed, q);
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
