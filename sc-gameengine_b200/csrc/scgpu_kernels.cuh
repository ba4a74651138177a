// scgpu_kernels.cuh — sm_100a kernels of the scene-update hot path.
//
// HBM layout. One device SLOT per Transform, given out when the Transform is spawned and kept until it is despawned
// (scgpu_layout.h): a hierarchy group stays contiguous for as long as it lives, whatever swap-with-last does to the
// ORDER of the reference's Transform pool. That order (the dense index, "rank") is carried beside the data:
//   rec0[slot] = { pos.x, pos.y, pos.z, rot.x }            float4 planes: one 128-bit load per thread, a warp
//   rec1[slot] = { rot.y, rot.z, scale.x, scale.y }        reads 512 contiguous bytes per instruction.
//   rec2[slot] = { scale.z, aabbMin.x, aabbMin.y, aabbMin.z }   64 B per instance = TRS 36 + AABB 24 + flags 4,
//   rec3[slot] = { aabbMax.x, aabbMax.y, aabbMax.z, flags }     exactly SURVEY.md §8(d)'s read set.
//   world0..3[slot] = world matrix columns (float4 planes, coalesced 128-bit stores)
//   rank[slot] = dense index of the Transform in the reference's pool; perm[rank] = slot   (k_spawn, k_despawn_apply)
//   parentSlot[slot] = resolved parent slot or kNone (maintained by k_resolve_parents on topology changes)
//   slotInfo[slot]   = depth + parent lane of the slot inside its hierarchy window, flags, 3-level work schedule;
//   winList[w]       = start slot of window w (+ flags): hierarchical scenes are cut into windows of <= 32 consecutive
//                      slots that no parent link crosses; one warp resolves one window (k_build_windows ->
//                      k_scan_tiles -> k_flatten_windows on topology changes; k_update_win every frame).
//   flags: bit0 HAS_BOUNDS, bit1 HAS_MESH, bit2 LIVE (clear: a hole left by a despawn), bits 8..31 = dirty stamp (id of
//          the update that must recompute the instance). A stamp instead of a dirty bit means the frame kernel never
//          writes the records.
//   visBits[view][rank / 32] = one bit per pool rank: the instance is visible in that view (+ one plane of culling
//          candidates when the culled lists are wanted). Written by the frame kernels for the ~1 % of instances that
//          survive the plane tests, read AND cleared by k_compact.
//
// Frame = k_update_flat | k_update_win + k_update_win_slow (transform + sphere + V-view plane tests, one pass over the
//         slots; visible instances set their bit)
//         -> k_compact (one pass over the bitmaps in rank order = the reference's pool order: per-view lists of entity
//            handles and slots, totals; leaves bitmaps, counters and the work queue zeroed for the next frame)
#pragma once
#include "scgpu_math.cuh"

// Checked build (make libscgpu_checked.so, -DSCGPU_CHECKED): every index the kernels derive from data — ranks, slots,
// chunk numbers, output positions, window geometry — is asserted against the bounds of the array it goes into. A
// violated assertion traps the kernel and the next API call fails. compute-sanitizer is closed on the GPU pool; the
// parity suite runs against this build instead (SCGPU_LIB=.../libscgpu_checked.so).
#ifdef SCGPU_CHECKED
#include <cassert>
#define SC_ASSERT(cond) assert(cond)
#else
#define SC_ASSERT(cond) ((void)0)
#endif

namespace scgpu
{

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kMaxViews = 8;
constexpr uint32_t kBlock = 256;       // threads per CTA
constexpr uint32_t kSubTiles = 4;      // sub-tiles of kBlock slots per CTA
constexpr uint32_t kTile = kBlock * kSubTiles;
constexpr uint32_t kFlagBounds = 1u, kFlagMesh = 2u, kFlagLive = 4u;
constexpr uint32_t kStampShift = 8;
constexpr uint32_t kUpdateSmemFlat = 3 * 4 * kBlock * 16;  // dynamic shared memory of k_update_flat
constexpr uint32_t kUpdForceDirty = 1u, kUpdFreeze = 2u, kUpdSkipTransform = 4u, kUpdCandBits = 8u;
constexpr uint32_t kChunkShift = 15, kChunkRanks = 1u << kChunkShift;  // compaction chunk: 32 Ki pool ranks = 1024 bitmap words
// accumulators of one frame (ScGpuScene::acc), all zero between frames (k_compact leaves them so)
constexpr uint32_t kAccCand = 0, kAccRecomputed = 1, kAccQueueNext = 2, kAccQueueSlow = 3, kAccWords = 4;

struct ViewPlanes
{
  float4 planes[kMaxViews][6];
  // 2^-19 * the largest |normal component| of any plane in use (host, scgpuSetViews / scgpuSetViewPlanes): scale of
  // the rounding slack of the fused favourite-plane pre-test, see sphere_cull_warp_fav
  float slackK;
  float pad_[3];
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
  const uint32_t* rank;  // slot -> dense index in the reference's pool
  uint32_t* visBits;     // [nViews (+1 with kUpdCandBits)][bitWords], indexed by rank
  uint32_t* chunkCounts; // [nViews (+1)][chunkStride]: set bits per chunk of kChunkRanks ranks (this frame's parity)
  uint32_t chunkStride;
  uint32_t* acc;         // kAcc* counters
  uint32_t count;        // slots in use (extent): live Transforms + holes
  uint32_t live;         // live Transforms: every rank is below it
  uint32_t bitWords;     // words per bitmap plane
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
  const uint32_t fl = __float_as_uint(p.rec3[s].w);
  if (p.flags & kUpdForceDirty) return (fl & kFlagLive) != 0u;
  return (fl >> kStampShift) == p.stamp;
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
      SC_ASSERT(cur < p.count);
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

__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
  const float2 xy = fmul2_rn(pl.x, pl.y, cx, cy);
  return __fadd_rn(__fadd_rn(__fadd_rn(xy.x, xy.y), __fmul_rn(pl.z, cz)), pl.w);
}

// ---- warp-cooperative pieces shared by the two frame kernels --------------------------------------------------------

// Sines and cosines of the three Euler angles of every lane. Angles below 2^-12 (zero above all) cost nothing in libm
// and nothing here; the others are worked off one PER ITERATION, each lane taking its next non-trivial angle
// whichever axis it belongs to. A warp whose lanes rotate about different single axes (wheels about X, props about Y,
// roots about Y) therefore evaluates one range reduction + two polynomials, not one per axis.
__device__ __forceinline__ void sincos3_warp(bool active, float rx, float ry, float rz, float& sx, float& cx, float& sy,
                                             float& cy, float& sz, float& cz)
{
  const bool nx = active && !sincos_is_trivial(rx), ny = active && !sincos_is_trivial(ry), nz = active && !sincos_is_trivial(rz);
  sx = rx; cx = 1.0f; sy = ry; cy = 1.0f; sz = rz; cz = 1.0f;
  if (!__any_sync(0xffffffffu, nx || ny || nz)) return;
  // First round without a per-lane branch: EVERY lane evaluates one angle - its first non-trivial one; a lane that has
  // none evaluates 0 and drops the result - and the results are put in place with selects. (The loop form below costs
  // ~25 more instructions per round in divergence bookkeeping: BSSY / BSYNC, the three-way assignment as branches.)
  {
    const float y = nx ? rx : (ny ? ry : (nz ? rz : 0.0f));
    float sn, cs;
    sincosf_glibc_nt(y, sn, cs);
    const bool ay = !nx && ny, az = !nx && !ny && nz;
    sx = nx ? sn : sx; cx = nx ? cs : cx;
    sy = ay ? sn : sy; cy = ay ? cs : cy;
    sz = az ? sn : sz; cz = az ? cs : cz;
  }
  // lanes with a second / third non-trivial angle (none in a city of wheels, props and roots)
  uint32_t need = ((nx && ny) ? 2u : 0u) | (((nx || ny) && nz) ? 4u : 0u);
  while (__any_sync(0xffffffffu, need != 0u))
  {
    if (need)
    {
      const float y = (need & 1u) ? rx : ((need & 2u) ? ry : rz);
      float sn, cs;
      sincosf_glibc_nt(y, sn, cs);
      if (need & 1u) { sx = sn; cx = cs; }
      else if (need & 2u) { sy = sn; cy = cs; }
      else { sz = sn; cz = cs; }
      need &= need - 1u;
    }
  }
}

// sphereInFrustum for all views of one warp (CullingSystem, .cpp:1240-1270; cull iff d < -radius on any plane, NaN
// keeps). The six tests of a view are independent, so their ORDER is free: `order` remembers, per view (3 bits
// each), the plane that culled this warp's previous instances. That plane is tested first for all views; if it
// culls every lane of the warp again - the normal case in an open world, where whole neighbourhoods lie on the same
// side of a frustum - the view is finished after one plane. Only when some lane survives are the other planes
// tested, one at a time with a warp-uniform early-out, and the plane that finished the job becomes the new favourite.
// The predicate evaluated per (instance, plane) is exactly the reference's; no instance is ever decided by a
// neighbour's result.
template <int kViews>
__device__ __forceinline__ uint32_t cull_views_warp(const ViewPlanes& vp, bool test, float cx, float cy, float cz,
                                                    float negR, uint32_t& order)
{
  uint32_t alive = 0;
#pragma unroll
  for (int v = 0; v < kViews; ++v)
  {
    const uint32_t k0 = (order >> (3 * v)) & 7u;
    const float d = plane_dist(vp.planes[v][k0], cx, cy, cz);
    alive |= (d < negR) ? 0u : (1u << v);
  }
  alive = test ? alive : 0u;
  if (__any_sync(0xffffffffu, alive != 0u))
  {
#pragma unroll 1
    for (uint32_t v = 0; v < (uint32_t)kViews; ++v)
    {
      if (!__any_sync(0xffffffffu, (alive >> v) & 1u)) continue;
      const uint32_t k0 = (order >> (3u * v)) & 7u;
#pragma unroll 1
      for (uint32_t k = 0; k < 6u; ++k)
      {
        if (k == k0) continue;
        const float d = plane_dist(vp.planes[v][k], cx, cy, cz);
        if (d < negR) alive &= ~(1u << v);
        if (!__any_sync(0xffffffffu, (alive >> v) & 1u))
        {
          order = (order & ~(7u << (3u * v))) | (k << (3u * v));
          break;
        }
      }
    }
  }
  return alive;
}

// Bounding sphere + plane tests of one warp. Far from every frustum - the normal case - the decision needs no square
// root: with the cheap UPPER bound of the radius, "the favourite plane culls every lane in every view" is already
// certain (d < -bound implies the reference's d < -radius). Only warps near a frustum compute the exact radius and run
// the exact tests; the result is the reference's in both cases.
template <int kViews>
__device__ __forceinline__ uint32_t sphere_cull_warp(const ViewPlanes& vp, bool test, const Mat4& W, float4 r2, float4 r3,
                                                     uint32_t& order)
{
  float ox, oy, oz, ex, ey, ez;
  world_bounds_centre(W, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, ox, oy, oz, ex, ey, ez);
  const float negBound = -world_bounds_radius_bound(W, ex, ey, ez);
  bool certain = true;
#pragma unroll
  for (int v = 0; v < kViews; ++v)
  {
    const uint32_t k0 = (order >> (3 * v)) & 7u;
    certain = certain && (plane_dist(vp.planes[v][k0], ox, oy, oz) < negBound);
  }
  if (__all_sync(0xffffffffu, certain || !test)) return 0u;
  return cull_views_warp<kViews>(vp, test, ox, oy, oz, -world_bounds_radius(W, ex, ey, ez), order);
}

// The same for the window kernel, whose issue slots are the bound: the pre-test costs ~24 instructions per warp
// instead of ~68 for five views.
//  * The favourite plane of every view is kept in a per-warp shared-memory cache (rewritten when `order` changes:
//    only in warps near a frustum), two views side by side: { nx_a nx_b ny_a ny_b } { nz_a nz_b d'_a d'_b } - two
//    broadcast LDS.128 per pair of views, no index arithmetic on `order`, no constant-bank loads.
//  * Two views per instruction: three fma.rn.f32x2 give n . o + d' for a pair of views.
//  * FUSED multiply-adds are allowed here although the reference's distance is ((n0*o0 + n1*o1) + n2*o2) + d with every
//    operation rounded: the pre-test only PROVES "culled", and it does so with a margin that covers the difference.
//    Both evaluations approximate the real number X = n . o + d; the reference's D with |D - X| <= 4u(1+e) S, the fused
//    chain D' with |D' - X| <= 3u(1+e) S, where u = 2^-24 and S = |n0 o0| + |n1 o1| + |n2 o2| + |d| <= K |o|_1 + |d|,
//    K = the largest |normal component| of any plane (ViewPlanes::slackK = 2^-19 K, from the host). The chain starts
//    from d' = d + 2^-19 |d| (rounded up) and is compared with negBound - 2^-19 K |o|_1: a slack of
//    2^-19 S >= 4 x the 7u S the two roundings can differ by. (Underflow adds at most a few 2^-149, the bound's
//    absolute 1e-30 covers that; rounding negBound - slack moves it by u |negBound|, the bound's factor 1.0001 covers
//    that.) So "D' + slack < negBound" implies D < negBound <= -radius: exactly the instances the reference culls by
//    that plane. NaN / Inf anywhere make a comparison false and the warp takes the exact path.
__device__ __forceinline__ float2 ffma2_rn(float ax, float ay, float b, float2 c)
{
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
  asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float4 lds128(uint32_t a);
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v);

// (re)writes the warp's cache of favourite planes from `order`; lane i writes float i of the layout above
template <int kViews>
__device__ __forceinline__ void fav_refresh(const ViewPlanes& vp, uint32_t favAddr, uint32_t lane, uint32_t order)
{
  constexpr uint32_t nC = kViews < 6 ? kViews : 6;
  __syncwarp();
  if (lane < ((nC + 1u) / 2u) * 8u)
  {
    const uint32_t c = (lane >> 1) & 3u;
    const uint32_t v = min(2u * (lane >> 3) + (lane & 1u), nC - 1u);  // an odd view count: the last view twice
    const float4 pl = vp.planes[v][(order >> (3u * v)) & 7u];
    const float val = c == 0u ? pl.x : (c == 1u ? pl.y : (c == 2u ? pl.z : __fmaf_ru(fabsf(pl.w), 0x1p-19f, pl.w)));
    sts32(favAddr + lane * 4u, __float_as_uint(val));
  }
  __syncwarp();
}

template <int kViews>
__device__ __forceinline__ uint32_t sphere_cull_warp_fav(const ViewPlanes& vp, uint32_t favAddr, uint32_t lane, bool test,
                                                         const Mat4& W, float4 r2, float4 r3, uint32_t& order)
{
  constexpr int nC = kViews < 6 ? kViews : 6;
  float ox, oy, oz, ex, ey, ez;
  world_bounds_centre(W, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, ox, oy, oz, ex, ey, ez);
  const float negBound = -world_bounds_radius_bound(W, ex, ey, ez);
  const float nb = __fmaf_rn(-vp.slackK, fabsf(ox) + fabsf(oy) + fabsf(oz), negBound);
  bool certain = true;
#pragma unroll
  for (int j = 0; j < (nC + 1) / 2; ++j)
  {
    const float4 A = lds128(favAddr + 32u * j), B = lds128(favAddr + 32u * j + 16u);
    float2 t = ffma2_rn(B.x, B.y, oz, make_float2(B.z, B.w));
    t = ffma2_rn(A.z, A.w, oy, t);
    t = ffma2_rn(A.x, A.y, ox, t);
    certain = certain & (t.x < nb) & (t.y < nb);
  }
#pragma unroll
  for (int v = nC; v < kViews; ++v)  // a seventh and eighth view: the plain form
  {
    const uint32_t k0 = (order >> (3 * v)) & 7u;
    certain = certain & (plane_dist(vp.planes[v][k0], ox, oy, oz) < negBound);
  }
  if (__all_sync(0xffffffffu, certain || !test)) return 0u;
  const uint32_t before = order;
  const uint32_t alive = cull_views_warp<kViews>(vp, test, ox, oy, oz, -world_bounds_radius(W, ex, ey, ez), order);
  if (order != before) fav_refresh<kViews>(vp, favAddr, lane, order);  // warp-uniform
  return alive;
}

// ---- visible bits -> the per-view bitmaps, indexed by POOL RANK (the reference's order) ------------------------------
// Called by a whole warp when at least one of its lanes has something to report (~1 % of the warps of an open world).
// Only those lanes read their rank. When the ranks are the lane numbers plus a common base — the layout of a scene that
// has seen no churn, and of most groups in one that has — the warp's 32 bits per view are two word-sized ORs by two
// lanes; otherwise every lane sets its own bits.
__device__ __forceinline__ void red_global_or(uint32_t* p, uint32_t v)
{
  asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void red_global_add_u32(uint32_t* p, uint32_t v)
{
  asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int kViews>
__device__ __forceinline__ void emit_visible_warp(const UpdateParams& p, uint32_t s, uint32_t lane, uint32_t mask, bool cand)
{
  const bool wantCand = (p.flags & kUpdCandBits) != 0;
  const bool need = mask != 0u || (wantCand && cand);
  const uint32_t needM = __ballot_sync(0xffffffffu, need);
  if (needM == 0u) return;
  SC_ASSERT(!need || s < p.count);
  const uint32_t r = need ? p.rank[s] : 0u;
  SC_ASSERT(r < p.live || !need);
  SC_ASSERT((r >> 5) < p.bitWords && (r >> kChunkShift) < p.chunkStride);
  const uint32_t first = __ffs(needM) - 1u;
  const uint32_t rFirst = __shfl_sync(0xffffffffu, r, first);
  const uint32_t base = rFirst - first;  // rank lane 0 would have
  const bool consecutive = __all_sync(0xffffffffu, !need || r == base + lane) && rFirst >= first;
  if (consecutive)
  {
    // lanes 0, 1: the two bitmap words the 32 ranks fall into; lanes 2, 3: the one or two chunks they fall into
    const uint32_t word = base >> 5, sh = base & 31u;
    const uint32_t chunk = base >> kChunkShift;
    const uint32_t room = ((chunk + 1u) << kChunkShift) - base;  // ranks left in the first chunk
    const uint32_t lowMask = room >= 32u ? 0xffffffffu : ((1u << room) - 1u);
#pragma unroll
    for (int v = 0; v <= kViews; ++v)
    {
      if (v == kViews && !wantCand) break;
      const uint32_t m = __ballot_sync(0xffffffffu, v < kViews ? ((mask >> v) & 1u) != 0u : cand);
      if (m == 0u) continue;
      uint32_t val = 0;
      uint32_t* dst = nullptr;
      if (lane < 2u)
      {
        val = lane == 0u ? (m << sh) : (sh ? (m >> (32u - sh)) : 0u);
        dst = p.visBits + (size_t)v * p.bitWords + word + lane;
        if (val) red_global_or(dst, val);
      }
      else if (lane < 4u)
      {
        val = __popc(lane == 2u ? (m & lowMask) : (m & ~lowMask));
        dst = p.chunkCounts + (size_t)v * p.chunkStride + chunk + (lane - 2u);
        if (val) red_global_add_u32(dst, val);
      }
    }
  }
  else
  {
    // Ranks scattered (a window whose groups were spawned at different times, after churn): still one reduction per
    // bitmap word and per chunk, by the lanes that share it — a warp near the camera would otherwise send up to 64
    // single-bit reductions per view at a handful of addresses, and those serialise in the L2.
#pragma unroll
    for (int v = 0; v <= kViews; ++v)
    {
      if (v == kViews && !wantCand) break;
      const bool bit = v < kViews ? ((mask >> v) & 1u) != 0u : cand;
      const uint32_t m = __ballot_sync(0xffffffffu, bit);
      if (m == 0u) continue;
      if (bit)
      {
        const uint32_t sameWord = __match_any_sync(m, r >> 5);
        const uint32_t bits = __reduce_or_sync(sameWord, 1u << (r & 31u));
        if (lane == __ffs(sameWord) - 1u) red_global_or(p.visBits + (size_t)v * p.bitWords + (r >> 5), bits);
        const uint32_t sameChunk = __match_any_sync(m, r >> kChunkShift);
        if (lane == __ffs(sameChunk) - 1u)
          red_global_add_u32(p.chunkCounts + (size_t)v * p.chunkStride + (r >> kChunkShift), __popc(sameChunk));
      }
    }
  }
}

// ---- K1+K2, flat scenes: fused transform + cull, all views in one pass ------------------------------------------------
// No instance has a parent: pure streaming. One thread per slot, kSubTiles sub-tiles of kBlock consecutive slots per
// CTA; the four record planes of a sub-tile are staged in shared memory by TMA bulk copies two sub-tiles ahead, so the
// DRAM round trip overlaps the arithmetic. kViews is a compile-time view count, so the plane tests read their planes
// straight from the constant bank.
template <int kViews>
__global__ void __launch_bounds__(kBlock, 4) k_update_flat(const __grid_constant__ UpdateParams p,
                                                           const __grid_constant__ ViewPlanes vp)
{
  __shared__ uint32_t sCounts[2];  // candidates, recomputed
  __shared__ __align__(8) uint64_t sFull[2];
  // dynamic shared memory (kUpdateSmemFlat bytes, opted in by the host):
  //   sRec[2][4][kBlock]  TMA staging, double buffered
  //   sX[kBlock][4]       per-thread exchange slot of the out-of-line dense fallback
  extern __shared__ __align__(128) unsigned char sDyn[];
  float4(*sRec)[4][kBlock] = reinterpret_cast<float4(*)[4][kBlock]>(sDyn);
  float4(*sX)[4] = reinterpret_cast<float4(*)[4]>(sDyn + 2 * 4 * kBlock * sizeof(float4));

  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  if (tid < 2) sCounts[tid] = 0;
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
  uint32_t order = 0;  // favourite plane per view, see cull_views_warp

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
    // WAR across proxies: the reads above go through the generic proxy, the refill below through the async proxy
    // (TMA). A CTA barrier alone does not order the two - a queued ld.shared was observed to return the NEXT
    // stage's bytes (tests: rows of sub-tile 0 carrying sub-tile 2's TRS). The proxy fence orders this thread's
    // reads before every async-proxy access that follows the barrier.
    fence_proxy_async_shared();
    __syncthreads();  // everybody has copied its record out of the staging buffer: refill it
    if (tid == 0 && sub + 2 < nSub) stage(sub + 2);

    const uint32_t fl = __float_as_uint(r3.w);
    // (a hole left by a despawn has flags == 0: stamp 0 is never the current one, and it owns no LIVE bit)
    const bool ownDirty = live && !skip && ((force && (fl & kFlagLive)) || ((fl >> kStampShift) == p.stamp));
    const bool tame = trs_inputs_tame(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x);
    float sx, cx, sy, cy, sz, cz;
    sincos3_warp(ownDirty && tame, r0.w, r1.x, r1.y, sx, cx, sy, cy, sz, cz);
    Mat4 W = mat4_identity();
    if (ownDirty)
    {
      if (tame) W = mat4_trs_from_sincos(r0.x, r0.y, r0.z, sx, cx, sy, cy, sz, cz, r1.z, r1.w, r2.x);
      else
      {
        trs_dense_to(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, &sX[tid][0]);
        W = xs_load(&sX[tid][0]);
      }
      store_world(p, s, W);
      ++nRecomputed;
    }
    else if (live)
    {
      W = load_world(p, s);
    }

    // ---- bounding sphere + plane tests in registers ----
    const bool cand = live && (fl & kFlagMesh);
    const bool test = cand && !freeze && (fl & kFlagBounds);
    uint32_t mask = 0;
    if (__any_sync(0xffffffffu, test))
    {
      mask = sphere_cull_warp<kViews>(vp, test, W, r2, r3, order);
    }
    if (cand && !test) mask = allMask;  // frozen culling or no Bounds component: always visible
    emit_visible_warp<kViews>(p, s, lane, mask, cand);
    nCand += cand ? 1u : 0u;
  }

  // frame totals: one warp reduction (REDUX) per counter, one shared atomic per warp, one global atomic per CTA
  {
    const uint32_t r = __reduce_add_sync(0xffffffffu, nCand);
    if (lane == 0 && r) atomicAdd(&sCounts[0], r);
    const uint32_t q = __reduce_add_sync(0xffffffffu, nRecomputed);
    if (lane == 0 && q) atomicAdd(&sCounts[1], q);
  }
  __syncthreads();
  if (tid < 2 && sCounts[tid]) atomicAdd(p.acc + (tid == 0 ? kAccCand : kAccRecomputed), sCounts[tid]);
}

// ---- cull-only frames: nothing to transform -----------------------------------------------------------------------------
// A frame in which no Transform is dirty (a static city under a moving camera; every SCGPU_UPDATE_SKIP_TRANSFORM update)
// needs neither the TRS records nor the hierarchy: the stored world matrices are valid, whatever their parents are. Pure
// streaming over the slots: AABB + flags (32 B) and world matrix (64 B) per instance = the 96 algorithmic bytes of a
// clean instance, six independent 128-bit loads per thread in flight at once, sphere + plane tests in
// registers, visible bits out. A CTA walks kCullSubTiles consecutive sub-tiles so that the favourite-plane memory of
// its warps (see cull_views_warp) carries over.
constexpr uint32_t kCullSubTiles = 8;

template <int kViews>
__global__ void __launch_bounds__(kBlock, 6) k_cull_only(const __grid_constant__ UpdateParams p, const __grid_constant__ ViewPlanes vp)
{
  __shared__ uint32_t sCand;
  constexpr uint32_t allMask = (1u << kViews) - 1u;
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const bool freeze = (p.flags & kUpdFreeze) != 0;
  if (tid == 0) sCand = 0u;
  __syncthreads();
  uint32_t order = 0, nCand = 0;
#pragma unroll 1
  for (uint32_t sub = 0; sub < kCullSubTiles; ++sub)
  {
    const uint32_t s = (blockIdx.x * kCullSubTiles + sub) * kBlock + tid;
    if (s - tid >= p.count) break;  // block-uniform
    const bool live = s < p.count;
    float4 r2 = make_float4(0.f, 0.f, 0.f, 0.f), r3 = r2;
    Mat4 W = mat4_identity();
    if (live)
    {
      r3 = ld_stream(p.rec3 + s); r2 = ld_stream(p.rec2 + s);
      W.c0 = ld_stream(p.w0 + s); W.c1 = ld_stream(p.w1 + s); W.c2 = ld_stream(p.w2 + s); W.c3 = ld_stream(p.w3 + s);
    }
    const uint32_t fl = __float_as_uint(r3.w);
    const bool cand = live && (fl & kFlagMesh);
    const bool test = cand && !freeze && (fl & kFlagBounds);
    uint32_t mask = 0;
    if (__any_sync(0xffffffffu, test)) mask = sphere_cull_warp<kViews>(vp, test, W, r2, r3, order);
    if (cand && !test) mask = allMask;  // frozen culling or no Bounds component: always visible
    emit_visible_warp<kViews>(p, s, lane, mask, cand);
    nCand += cand ? 1u : 0u;
  }
  const uint32_t r = __reduce_add_sync(0xffffffffu, nCand);
  if (lane == 0 && r) atomicAdd(&sCand, r);
  __syncthreads();
  if (tid == 0 && sCand) atomicAdd(p.acc + kAccCand, sCand);
}

// ---- hierarchy windows ------------------------------------------------------------------------------------------
// A window is a run of <= 32 consecutive slots owned by one warp. k_build_windows (topology changes only) cuts the
// slot range into windows at positions that NO parent link crosses, so a hierarchy group never straddles two warps
// and k_update_win resolves it inside the warp. Tile boundaries are not special: a window may run up to 31 slots
// into the next tile when a group straddles the boundary (the tile where a window STARTS owns it). Per slot the
// build records depth and parent lane inside the window, per window the start slot; links longer than a window
// (or chains deeper than 32) are flagged and resolved from global memory by the generic path (window_slow).
constexpr uint32_t kMaxWin = 96;              // windows per tile: two consecutive greedy windows span > 32 slots => <= 66
constexpr uint32_t kHalo = 32;                // slots looked at on either side of a tile
constexpr uint32_t kInfoDepthMask = 31u;      // slotInfo bits 0..4: depth inside the window
constexpr uint32_t kInfoParentShift = 5;      // bits 5..9: lane of the parent (depth > 0)
constexpr uint32_t kInfoExternal = 1u << 10;  // parent lives outside the window (resolved by walk_up)
constexpr uint32_t kInfoUnreachable = 1u << 11;  // node on / below a cycle closed inside its window
constexpr uint32_t kInfoSchedShift = 12;      // bits 12..29: static schedule of the level loop, 6 bits per level 1..3:
                                              //   bit 5 = this lane works in that level, bits 0..4 = lane of the child it
                                              //   helps with (lanes 2j, 2j+1 take the j-th node of the level)
constexpr uint32_t kWinSlow = 0x80000000u;    // window list entry: some lane is external / unreachable
constexpr uint32_t kWinNoStatic = 0x40000000u;  // window list entry: no static schedule (deeper than 3 levels, or a
                                                // level with more nodes than half the window's lanes)
constexpr uint32_t kWinSlotMask = 0x3FFFFFFFu;

// first position in [from, from+kHalo) (and <= count) that no short link crosses; `from` itself when there is none
// (forced cut through a chain longer than a window). sCross is indexed relative to `origin`.
__device__ __forceinline__ uint32_t first_uncrossed(const int* sCross, uint32_t origin, uint32_t from, uint32_t count)
{
  if (from >= count) return count;
  for (uint32_t c = from; c < from + kHalo && c <= count; ++c)
    if (c == count || sCross[c - origin] == 0) return c;
  return from;
}

// INCREMENTAL: the cut of a tile depends on the parentSlot words of [tileBase - kHalo, tileBase + kTile + 2 kHalo) and on
// `count` alone, slots never move, and the per-tile results (slotInfo, winLocal, tileWinCount) stay in place between
// calls. So a tile is cut again only if a link changed in it, in the last kHalo slots of the tile before it or in the
// first 2 kHalo slots of the tile behind it (tileDirty, see SceneArrays; the host marks the tiles around a changed
// `count`); every other CTA returns at once. The cost of a topology change is
// proportional to what changed, not to the size of the scene.
__global__ void __launch_bounds__(kBlock) k_build_windows(const uint32_t* __restrict__ parentSlot, uint32_t* __restrict__ slotInfo,
                                                          uint16_t* __restrict__ winLocal, uint32_t* __restrict__ tileWinCount,
                                                          uint32_t count, const uint8_t* __restrict__ tileDirty)
{
  static_assert(kHalo == 32, "mark_tile's edge zones");
  if (!(tileDirty[blockIdx.x * 4u + 2u] | tileDirty[(blockIdx.x + 1u) * 4u] | tileDirty[(blockIdx.x + 2u) * 4u + 1u])) return;
  constexpr uint32_t kSpan = kTile + 3 * kHalo;  // slots [tileBase - kHalo, tileBase + kTile + 2*kHalo)
  __shared__ int sCross[kSpan + 8];              // sCross[c - origin]: number of short links crossing position c
  __shared__ uint16_t sStart[kMaxWin + 1];       // window starts relative to tileBase (<= kTile + kHalo)
  __shared__ uint16_t sWinOf[kTile + kHalo];     // window index of every owned slot (relative to tileBase)
  __shared__ uint32_t sWinFlag[kMaxWin];         // bit 0: generic path, bit 1: no static schedule
  __shared__ uint32_t sLvl[kMaxWin][3];          // lanes of the window at depth 1, 2, 3
  __shared__ uint8_t sChild[kMaxWin][3][16];     // the j-th lane of that depth (static schedule)
  __shared__ uint32_t sInfo[kTile + kHalo];
  __shared__ uint16_t sPar[kSpan];               // parent of every slot of the span, relative to `origin` (see kPar*)
  __shared__ uint32_t sWarpSum[kBlock / 32];
  __shared__ uint32_t sNumWin, sBeg, sEnd;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t tileBase = blockIdx.x * kTile;
  const uint32_t origin = tileBase >= kHalo ? tileBase - kHalo : 0u;  // slot / position of sCross[0]
  const uint32_t spanEnd = min(count, tileBase + kTile + 2 * kHalo);  // slots [origin, spanEnd)

  for (uint32_t k = tid; k < kSpan + 8; k += kBlock) sCross[k] = 0;
  for (uint32_t k = tid; k < kMaxWin; k += kBlock) { sWinFlag[k] = 0; sLvl[k][0] = sLvl[k][1] = sLvl[k][2] = 0; }
  __syncthreads();
  // a link between slots lo < hi shorter than a window crosses every cut position c with lo < c <= hi
  // (parentSlot is read from global memory ONCE, here; the depth walk below follows the links in shared memory)
  constexpr uint16_t kParNone = 0xFFFFu, kParFar = 0xFFFEu;  // no parent / parent too far away to share a window
  for (uint32_t k = origin + tid; k < spanEnd; k += kBlock)
  {
    const uint32_t ps = parentSlot[k];
    if (ps == kNone) { sPar[k - origin] = kParNone; continue; }
    const uint32_t lo = min(k, ps), hi = max(k, ps);
    const bool near = ps != k && ps < count && hi - lo < 32u && ps >= origin && ps < spanEnd;
    sPar[k - origin] = near ? (uint16_t)(ps - origin) : kParFar;
    if (ps == k || ps >= count) continue;
    if (hi - lo >= 32u) continue;  // can never sit inside one window: the child is flagged external below
    // positions lo+1 .. hi, clipped to the span (difference array, summed below)
    const uint32_t from = max(lo + 1u, origin), to = min(hi + 1u, origin + kSpan + 4u);
    if (from < to)
    {
      atomicAdd(&sCross[from - origin], 1);
      atomicAdd(&sCross[to - origin], -1);
    }
  }
  __syncthreads();
  // inclusive prefix sum over the span: thread t owns 5 consecutive positions (256 * 5 >= kSpan + 8)
  {
    constexpr int kPer = 5;
    static_assert(kBlock * kPer >= kSpan + 8, "prefix sum coverage");
    int v[kPer], sum = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j)
    {
      const uint32_t idx = tid * kPer + j;
      sum += (idx < kSpan + 8) ? sCross[idx] : 0;
      v[j] = sum;
    }
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
#pragma unroll
    for (int j = 0; j < kPer; ++j)
    {
      const uint32_t idx = tid * kPer + j;
      if (idx < kSpan + 8) sCross[idx] = v[j] + off;
    }
  }
  __syncthreads();
  // this tile owns the windows that start in [beg, end): beg / end = first uncrossed position at or after the tile's
  // first slot / the next tile's first slot. Both neighbours derive the shared boundary from the same data.
  if (tid == 0)
  {
    sBeg = first_uncrossed(sCross, origin, tileBase, count);
    sEnd = first_uncrossed(sCross, origin, tileBase + kTile, count);
  }
  __syncthreads();
  const uint32_t beg = sBeg, end = sEnd;
  // greedy cuts by warp 0: the next window ends at the LARGEST uncrossed position within 32 slots (forced at 32)
  if (warp == 0)
  {
    uint32_t start = beg, nw = 0;
    while (start < end && nw < kMaxWin)
    {
      if (lane == 0) sStart[nw] = (uint16_t)(start - tileBase);
      const uint32_t c = start + 1 + lane;  // candidate end
      const bool ok = c <= end && (c == end || sCross[c - origin] == 0);
      const uint32_t m = __ballot_sync(0xffffffffu, ok);
      const uint32_t next = m ? start + 32u - __clz(m) : min(start + 32u, end);
      if (start + lane < next) sWinOf[start + lane - tileBase] = (uint16_t)nw;  // the window's slots learn their window here
      start = next;
      ++nw;
    }
    if (lane == 0) { sStart[nw] = (uint16_t)(end - tileBase); sNumWin = nw; }
  }
  __syncthreads();
  const uint32_t nw = sNumWin;
  // per owned slot: depth and parent lane inside the window
  for (uint32_t k = (beg - tileBase) + tid; k < end - tileBase; k += kBlock)
  {
    const uint32_t w = sWinOf[k], wa = sStart[w], wb = sStart[w + 1];
    uint32_t cur = k, depth = 0, info = 0;
    for (;;)
    {
      const uint32_t code = sPar[tileBase + cur - origin];
      if (code == kParNone) break;
      const uint32_t ps = origin + code;  // meaningless for kParFar, which fails the test below
      const bool inside = code != kParFar && ps >= tileBase + wa && ps < tileBase + wb;
      if (!inside)
      {
        if (depth == 0) info = kInfoExternal;  // deeper nodes hang off an ancestor that carries the flag itself
        break;
      }
      if (depth == 0) info |= (ps - tileBase - wa) << kInfoParentShift;
      cur = ps - tileBase;
      if (++depth > 32u) { info = kInfoUnreachable; depth = 0; break; }  // a cycle closed inside the window
    }
    if (info & (kInfoExternal | kInfoUnreachable)) atomicOr(&sWinFlag[w], 1u);
    if (depth > 3u) atomicOr(&sWinFlag[w], 2u);
    else if (depth != 0u) atomicOr(&sLvl[w][depth - 1u], 1u << (k - wa));
    sInfo[k] = info | (depth & kInfoDepthMask);
  }
  __syncthreads();
  // static schedule of the level loop (used by k_update_win when every node of the window is recomputed): in
  // level l, lanes 2j and 2j+1 of the window take the j-th node of that level. First every node writes its lane at
  // its position in the order of its level, then every lane looks up the node it helps with.
  for (uint32_t k = (beg - tileBase) + tid; k < end - tileBase; k += kBlock)
  {
    const uint32_t d = sInfo[k] & kInfoDepthMask;
    if (d >= 1u && d <= 3u && !(sInfo[k] & kInfoUnreachable))
    {
      // Order of the nodes of one level: by (rank inside the residue class of the lane index mod 4, class), not by
      // lane index. The four lane pairs of a quarter warp then work on children from different classes wherever the
      // level has them, and two children collide in the shared-memory banks of the matrix planes exactly when they
      // are in the same class (plane rows are 16 B per lane; the odd lane of a pair is skewed by 64 B, kMatC2).
      const uint32_t w = sWinOf[k], h = k - sStart[w];
      const uint32_t M = sLvl[w][d - 1u], r = h & 3u;
      const uint32_t t = __popc(M & (0x11111111u << r) & ((1u << h) - 1u));
      uint32_t j = 0;
#pragma unroll
      for (uint32_t q = 0; q < 4u; ++q)
      {
        const uint32_t c = __popc(M & (0x11111111u << q));
        j += min(c, t) + ((q < r && c > t) ? 1u : 0u);
      }
      if (j < 16u) sChild[w][d - 1u][j] = (uint8_t)h;
    }
  }
  __syncthreads();
  for (uint32_t k = (beg - tileBase) + tid; k < end - tileBase; k += kBlock)
  {
    const uint32_t w = sWinOf[k], wa = sStart[w], len = sStart[w + 1] - wa, h = k - wa;
    uint32_t info = sInfo[k];
#pragma unroll
    for (uint32_t l = 0; l < 3u; ++l)
    {
      const uint32_t cnt = __popc(sLvl[w][l]);
      if (2u * cnt > len) { if (h == 0u) atomicOr(&sWinFlag[w], 2u); continue; }
      const uint32_t j = h >> 1;
      if (j >= cnt) continue;
      info |= (32u | (uint32_t)sChild[w][l][j]) << (kInfoSchedShift + 6u * l);
    }
    slotInfo[tileBase + k] = info;
  }
  __syncthreads();
  for (uint32_t k = tid; k <= nw; k += kBlock)
    winLocal[(size_t)blockIdx.x * (kMaxWin + 1) + k] = (uint16_t)(sStart[k] | ((k < nw) ? (sWinFlag[k] & 3u) << 14 : 0u));
  if (tid == 0) tileWinCount[blockIdx.x] = nw;
}

// per-tile window lists -> one list of absolute start slots (bit 31: generic path), winList[total] = count
__global__ void __launch_bounds__(128) k_flatten_windows(const uint16_t* __restrict__ winLocal, const uint32_t* __restrict__ tileWinCount,
                                                         const uint32_t* __restrict__ tileWinBase, uint32_t* __restrict__ winList,
                                                         uint32_t numTiles, uint32_t count, uint8_t* __restrict__ tileDirty)
{
  const uint32_t tile = blockIdx.x, nw = tileWinCount[tile], base = tileWinBase[tile];
  if (threadIdx.x < 4) tileDirty[(tile + 1u) * 4u + threadIdx.x] = 0;  // every k_build_windows CTA of this rebuild has finished (stream order)
  for (uint32_t k = threadIdx.x; k < nw; k += 128)
  {
    const uint32_t e = winLocal[(size_t)tile * (kMaxWin + 1) + k];
    winList[base + k] = (tile * kTile + (e & 0x3FFFu)) | ((e & 0x4000u) ? kWinSlow : 0u) | ((e & 0x8000u) ? kWinNoStatic : 0u);
  }
  if (tile == numTiles - 1 && threadIdx.x == 0) winList[base + nw] = count;
}

// ---- generic window resolution (rare) ------------------------------------------------------------------------------
// Exact for ANY input: full 4x4 matrices, dense products where the structured ones are not value-exact, parents
// outside the window resolved from global memory (walk_up), cycles left untouched. Called by all lanes of the warp.
// Leaves this lane's world matrix in out[0..3]; returns 1 if the node was recomputed this frame.
__device__ __noinline__ uint32_t window_slow(const UpdateParams& p, uint32_t a, uint32_t len, uint32_t info, float4* out)
{
  float4 xa[4], xb[4];
  const uint32_t lane = threadIdx.x & 31u;
  const bool live = lane < len;
  const uint32_t s = a + lane;
  const bool skip = (p.flags & kUpdSkipTransform) != 0;
  const bool force = (p.flags & kUpdForceDirty) != 0;
  float4 r0, r1, r2, r3;
  r0 = r1 = r2 = r3 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) { r0 = p.rec0[s]; r1 = p.rec1[s]; r2 = p.rec2[s]; r3 = p.rec3[s]; }
  const uint32_t fl = __float_as_uint(r3.w);
  const bool ownDirty = live && !skip && ((force && (fl & kFlagLive)) || ((fl >> kStampShift) == p.stamp));
  const uint32_t wl = info & kInfoDepthMask;
  const bool external = (info & kInfoExternal) != 0;
  bool dead = !live || (info & kInfoUnreachable);  // never visited by the DFS: world matrix stays as stored
  const uint32_t parentLane = (wl != 0) ? ((info >> kInfoParentShift) & 31u) : lane;
  const uint32_t maxL = __reduce_max_sync(0xffffffffu, live ? wl : 0u);

  bool nodeDirty = ownDirty;
  if (live && external)
  {
    const uint32_t walk = walk_up(p, p.parentSlot[s], true, xa, xb);
    if (!(walk & 1u)) dead = true;
    nodeDirty = ownDirty || (walk & 2u) != 0;
  }
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
  for (uint32_t l = 1; l <= maxL; ++l)
  {
    Mat4 PW;
    PW.c0.x = __shfl_sync(0xffffffffu, W.c0.x, parentLane); PW.c0.y = __shfl_sync(0xffffffffu, W.c0.y, parentLane);
    PW.c0.z = __shfl_sync(0xffffffffu, W.c0.z, parentLane); PW.c0.w = __shfl_sync(0xffffffffu, W.c0.w, parentLane);
    PW.c1.x = __shfl_sync(0xffffffffu, W.c1.x, parentLane); PW.c1.y = __shfl_sync(0xffffffffu, W.c1.y, parentLane);
    PW.c1.z = __shfl_sync(0xffffffffu, W.c1.z, parentLane); PW.c1.w = __shfl_sync(0xffffffffu, W.c1.w, parentLane);
    PW.c2.x = __shfl_sync(0xffffffffu, W.c2.x, parentLane); PW.c2.y = __shfl_sync(0xffffffffu, W.c2.y, parentLane);
    PW.c2.z = __shfl_sync(0xffffffffu, W.c2.z, parentLane); PW.c2.w = __shfl_sync(0xffffffffu, W.c2.w, parentLane);
    PW.c3.x = __shfl_sync(0xffffffffu, W.c3.x, parentLane); PW.c3.y = __shfl_sync(0xffffffffu, W.c3.y, parentLane);
    PW.c3.z = __shfl_sync(0xffffffffu, W.c3.z, parentLane); PW.c3.w = __shfl_sync(0xffffffffu, W.c3.w, parentLane);
    if (nodeDirty && wl == l) W = compose_any(PW, W, affine, xa, xb);
  }
  xs_store(out, W);
  return nodeDirty ? 1u : 0u;
}

// ---- shared-memory access by 32-bit shared address ------------------------------------------------------------------
// The window kernel addresses its shared memory through plain 32-bit shared-window addresses computed once per
// warp; going through generic pointers makes the compiler rebuild the address (S2R CgaCtaId, LEA ...) at every use.
__device__ __forceinline__ float4 lds128(uint32_t a)
{
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v)
{
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v)
{
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t a)
{
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v)
{
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory");
}
// One elected lane of the (converged) warp does the atomic. Written with elect.sync so that ptxas emits a bare
// ATOMS: behind an `if (lane == 0)` it wraps every shared atomic in its warp-aggregation sequence (VOTE, FLO, two
// POPC, S2R LTMASK, SHFL), ~15 instructions that are pure overhead when one lane is active by construction.
__device__ __forceinline__ uint32_t warp_atoms_add(uint32_t a, uint32_t v)
{
  uint32_t old = 0, leader;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync %1|p, 0xffffffff;\n\t@p atom.shared.add.u32 %0, [%2], %3;\n\t}"
               : "+r"(old), "=r"(leader)
               : "r"(a), "r"(v)
               : "memory");
  return __shfl_sync(0xffffffffu, old, leader);
}
// the two halves of warp_atoms_add, for callers that have work to do while the atomic is in flight
__device__ __forceinline__ void warp_atoms_add_issue(uint32_t a, uint32_t v, uint32_t& old, uint32_t& leader)
{
  old = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync %1|p, 0xffffffff;\n\t@p atom.shared.add.u32 %0, [%2], %3;\n\t}"
               : "+r"(old), "=r"(leader)
               : "r"(a), "r"(v)
               : "memory");
}
__device__ __forceinline__ void warp_reds_add(uint32_t a, uint32_t v)
{
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" ::"r"(a), "r"(v) : "memory");
}
// next free entry of the deferred-window list (single caller lane)
__device__ __forceinline__ uint32_t warp_slow_slot(uint32_t sBase);
__device__ __forceinline__ void cp_async4s(uint32_t smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16s(uint32_t smem, const void* gmem)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}

// ---- K1+K2, hierarchical scenes: one window per warp -----------------------------------------------------------------
// Loads and stores are the flat kernel's (coalesced 128-bit planes, slot = window start + lane). Inside the warp:
//   1. dirty bits are propagated parent -> child with ballots (skipped when everything is dirty anyway),
//   2. every lane that must be recomputed builds its local matrix (all lanes busy, FP64 sincos in parallel),
//   3. the matrices go to a per-warp shared-memory area; for level 1..max the products parent.world * local of that
//      level are computed by lane PAIRS (one lane columns 0-1, the other columns 2-3 + translation), 16 children per
//      round, whichever lanes the children themselves occupy: parent-before-child, parent matrices staged in shared
//      memory, and the lanes stay busy although a level holds only a fraction of the window's nodes.
// The fast path assumes what holds for every sane scene - tame TRS values, affine stored matrices, finite
// translations, parents inside the window - and VERIFIES it per window (two warp votes); a window that fails the
// check is redone by window_slow(), which is exact for any input. No CTA barrier inside the loop, no dependency
// between warps: a warp that finishes a window claims the next unclaimed one of its CTA.
#ifndef SCGPU_WIN_BLOCK
#define SCGPU_WIN_BLOCK 128
#define SCGPU_WIN_MINBLOCKS 8
#endif
constexpr uint32_t kWinBlock = SCGPU_WIN_BLOCK;          // threads per CTA of k_update_win
constexpr uint32_t kWinWarps = kWinBlock / 32;
#ifndef SCGPU_WIN_CHUNK
#define SCGPU_WIN_CHUNK 8
#endif
constexpr uint32_t kWinChunk = SCGPU_WIN_CHUNK;                        // consecutive windows per claim from the global queue
// shared memory of k_update_win (byte offsets). Everything a warp touches in the loop sits in ONE per-warp block, so
// that every address is "lane base + constant" and folds into the instruction's immediate offset.
constexpr uint32_t kWsBuf = 4 * 512 + 128;              // one prefetch buffer: 4 record planes + 32 slotInfo words
constexpr uint32_t kWwMat = 2 * kWsBuf;                 // [4][32] float4: matrix columns of the level loop
// Byte offsets of the four column planes inside kWwMat. Columns 2 and 3 are skewed by 64 B: in a level product the even
// lane of a pair reads / writes columns 0, 1 of its child and the odd lane columns 2, 3 of the SAME child; with plain
// 512-byte planes both halves of every pair hit the same banks (a 2-way conflict on every LDS.128 / STS.128 of the
// level loop, 7.7 wavefronts per instruction instead of 3 in ncu). With the skew the four pairs of a quarter warp,
// which work on consecutive children, cover 128 distinct bytes.
constexpr uint32_t kMatC1 = 512, kMatC2 = 1024 + 64, kMatC3 = 1536 + 64;
constexpr uint32_t kWwSched = kWwMat + 4 * 512 + 64;    // [32] u16: children of the current level
constexpr uint32_t kWwListSlot = 48;                    // bytes per list slot (kWinChunk + 1 words, rounded up to 16)
constexpr uint32_t kWwList = kWwSched + 64;             // 2 x [kWinChunk+1] u32: window starts of the current and of the
                                                        // next claimed chunk
constexpr uint32_t kWwFav = kWwList + 2 * kWwListSlot;  // 3 x 32 B: favourite planes of up to 6 views, two views per
                                                        // f32x2 lane pair (sphere_cull_warp_fav)
constexpr uint32_t kWwSize = kWwFav + 96;               // per-warp block
static_assert((kWinChunk + 1) * 4 <= kWwListSlot && kWwFav % 16 == 0 && kWwSize % 16 == 0, "per-warp block layout");
constexpr uint32_t kWsRecomputed = kWinWarps * kWwSize; // u32: world matrices rewritten by this CTA; +4: its candidates
constexpr uint32_t kUpdateSmemWin = kWsRecomputed + 16;
// device-side work queue of k_update_win: acc[kAccQueueNext] = next unclaimed chunk, acc[kAccQueueSlow] = number of
// windows handed to k_update_win_slow (k_compact leaves both zeroed for the next frame)
// frame totals: [0..nViews) visible, [nViews] candidates, [kMaxViews+1] recomputed, [kMaxViews+2] windows that took the generic path
constexpr uint32_t kTotalsWords = kMaxViews + 3;

// Programmatic dependent launch (the chain k_update_win -> k_update_win_slow -> k_scan_tiles -> k_scatter_visible of
// one frame): pdl_trigger() lets the NEXT kernel of the stream be scheduled while this one still runs, pdl_wait() at
// the top of that kernel blocks until this one has completed and its writes are visible. Launch latency and CTA
// ramp-up then overlap the predecessor's tail. Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void red_global_add(uint32_t* p, uint32_t v)
{
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q red.global.add.u32 [%0], %1;\n\t}" ::"l"(p), "r"(v) : "memory");
}
// claim of one chunk of windows: issue now (one elected lane), read the answer later. The queue counts CLAIMS;
// claim_to_chunk maps a claim index to its windows.
__device__ __forceinline__ void queue_claim_issue(uint32_t* queue, uint32_t& old)
{
  old = 0;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q atom.global.add.u32 %0, [%1], 1;\n\t}"
               : "+r"(old) : "l"(queue) : "memory");
}
// Guided schedule: the first k1 claims are chunks of kWinChunk windows, the rest of the list is handed out in chunks
// of kWinFine. A warp works ~3.6 us per window and owns up to three chunks at a time (current, staged, in flight),
// so with coarse chunks only the warps finish up to a chunk (~29 us of a 450 us launch) apart; the fine region is long
// enough (24 windows per resident warp) for every warp to reach it before the list ends. kWinFine >= 3: the list of
// the next chunk is staged while window 1 of the current chunk is processed and must have landed (wait at the loop
// top) before the last window of the chunk fetches from it.
constexpr uint32_t kWinFine = 3;
__device__ __forceinline__ void claim_to_chunk(uint32_t k, uint32_t total, uint32_t k1, uint32_t& cs, uint32_t& cnt)
{
  const bool coarse = k < k1;
  cs = coarse ? k * kWinChunk : k1 * kWinChunk + (k - k1) * kWinFine;
  cnt = cs < total ? min(coarse ? kWinChunk : kWinFine, total - cs) : 0u;
}
__device__ __forceinline__ uint32_t queue_claim_result(uint32_t old)
{
  uint32_t leader;  // elect.sync is deterministic for a given member mask: the same lane that issued the atomic
  asm volatile("{\n\t.reg .pred q;\n\telect.sync %0|q, 0xffffffff;\n\t}" : "=r"(leader));
  return __shfl_sync(0xffffffffu, old, leader);
}

// store + bounding sphere + plane tests of one resolved window (records in shared memory at recAddr); visible lanes set
// their bits in the rank-indexed bitmaps, candidates are counted warp-uniformly and flushed once per warp
// favAddr: the warp's favourite-plane cache (k_update_win), 0 = none (k_update_win_slow)
template <int kViews, bool kFav>
__device__ __forceinline__ void finish_window(const UpdateParams& p, const ViewPlanes& vp, uint32_t a, uint32_t lane, uint32_t recAddr,
                                              bool live, bool nodeDirty, const Mat4& W, uint32_t& order, uint32_t& nRecomputed,
                                              uint32_t& accCand, uint32_t favAddr)
{
  constexpr uint32_t allMask = (1u << kViews) - 1u;
  const bool freeze = (p.flags & kUpdFreeze) != 0;
  if (nodeDirty) store_world(p, a + lane, W);
  nRecomputed += __popc(__ballot_sync(0xffffffffu, nodeDirty));  // warp-uniform running count, flushed once per warp
  // lanes beyond the window read whatever their part of the buffer holds (their own shared memory; `live` masks
  // everything derived from it): no branch around the loads
  const float4 r3 = lds128(recAddr + 1536);
  const uint32_t fl = __float_as_uint(r3.w);
  const bool cand = live && (fl & kFlagMesh);
  const bool test = cand && !freeze && (fl & kFlagBounds);
  uint32_t mask = 0;
  if (__any_sync(0xffffffffu, test))
  {
    const float4 r2 = lds128(recAddr + 1024);
    if (kFav) mask = sphere_cull_warp_fav<kViews>(vp, favAddr, lane, test, W, r2, r3, order);
    else mask = sphere_cull_warp<kViews>(vp, test, W, r2, r3, order);
  }
  if (cand && !test) mask = allMask;
  accCand += __popc(__ballot_sync(0xffffffffu, cand));
  emit_visible_warp<kViews>(p, a + lane, lane, mask, cand);
}

}  // namespace scgpu
// The window kernel's shared memory as a file-scope array with an unmangled name: its shared-window address is then a
// link-time IMMEDIATE ("mov.u32 r, scgpu_win_smem") that costs no register, where the address of a dynamic array is
// a computed value that ptxas kept spilling to local memory across the window loop.
extern "C" __shared__ __align__(128) unsigned char scgpu_win_smem[scgpu::kUpdateSmemWin];
namespace scgpu
{

// PERSISTENT: the grid is 8 CTAs per SM; warps pull chunks of kWinChunk consecutive windows from a device-wide queue
// (one global atomic per chunk, issued a whole chunk ahead of its use), so there is no CTA-level barrier, no per-CTA
// prologue per tile, and the tail of the launch is one chunk per warp. Windows that need the generic path are appended
// to a global list and handled by k_update_win_slow afterwards.
template <int kViews>
__global__ void __launch_bounds__(kWinBlock, SCGPU_WIN_MINBLOCKS) k_update_win(const __grid_constant__ UpdateParams p,
                                                          const __grid_constant__ ViewPlanes vp,
                                                          const uint32_t* __restrict__ slotInfo,
                                                          const uint32_t* __restrict__ winList,
                                                          const uint32_t* __restrict__ totalWindows,
                                                          uint32_t* __restrict__ queue,  // &acc[kAccQueueNext], [1] = slow count
                                                          uint32_t* __restrict__ slowList)
{
  uint32_t sBase;
  asm("mov.u32 %0, scgpu_win_smem;" : "=r"(sBase));
  asm volatile("" ::"l"(scgpu_win_smem));  // keeps the array emitted: every other access goes through sBase
  const uint32_t tid = threadIdx.x;
  // lane / warp ids through volatile asm: the compiler then keeps them in registers instead of re-reading the
  // special register (S2R, ~20 cycles each) wherever register pressure makes rematerialisation look cheap
  uint32_t lane, warp;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  asm volatile("shr.u32 %0, %1, 5;" : "=r"(warp) : "r"(tid));
  if (tid < 2) reinterpret_cast<uint32_t*>(scgpu_win_smem + kWsRecomputed)[tid] = 0u;  // (also keeps the symbol referenced)
  __syncthreads();
  pdl_trigger();
  const uint32_t total = *totalWindows;

  // "this slot must be recomputed" as ONE masked compare of its flags word: the stamp equals this update's id; with every
  // instance forced dirty: the LIVE bit is set (a hole is never dirty); in a cull-only update: never (0 != 1 under mask 0)
  const bool skip = (p.flags & kUpdSkipTransform) != 0;
  const bool force = (p.flags & kUpdForceDirty) != 0;
  const uint32_t dirtyMask = skip ? 0u : (force ? kFlagLive : 0xFFFFFF00u);
  const uint32_t dirtyWant = skip ? 1u : (force ? kFlagLive : (p.stamp << kStampShift));
  uint32_t order = 0;  // favourite plane per view, see cull_views_warp
  uint32_t nRecomputed = 0, accCand = 0;
  const uint32_t warpBase = sBase + warp * kWwSize;
  const uint32_t laneBase = warpBase + lane * 16;  // this lane's float4 of plane 0, buffer 0

  // window starts of chunk [cs, cs + kWinChunk] -> list slot `slot` (0/1) of this warp, asynchronously
  auto stage_list = [&](uint32_t cs, uint32_t slot)
  {
    if (lane <= kWinChunk) cp_async4s(warpBase + kWwList + slot * kWwListSlot + lane * 4u, winList + min(cs + lane, total));
  };
  // records + slotInfo words of the window whose list entry sits at listAddr -> prefetch buffer at byte offset off
  auto fetch = [&](uint32_t listAddr, uint32_t off)
  {
    const uint32_t e0 = lds32(listAddr) & kWinSlotMask, e1 = lds32(listAddr + 4) & kWinSlotMask;
    if (lane < e1 - e0)
    {
      const uint32_t q = e0 + lane;
      const uint32_t d = laneBase + off;
      cp_async16s(d, p.rec0 + q);
      cp_async16s(d + 512, p.rec1 + q);
      cp_async16s(d + 1024, p.rec2 + q);
      cp_async16s(d + 1536, p.rec3 + q);
      cp_async4s(d + 2048 - lane * 12, slotInfo + q);
    }
  };

  fav_refresh<kViews>(vp, warpBase + kWwFav, lane, order);
  // ---- prologue: first chunk (its list is waited for), second chunk claimed and staged, third claim in flight ----
  uint32_t claimOld;
  queue_claim_issue(queue, claimOld);
  const uint32_t fineWindows = gridDim.x * kWinWarps * 24u;
  const uint32_t k1 = total > fineWindows ? (total - fineWindows) / kWinChunk : 0u;  // claims below k1 are coarse
  uint32_t cs, cnt;  // first window and number of windows of the current chunk
  claim_to_chunk(queue_claim_result(claimOld), total, k1, cs, cnt);
  uint32_t pos = 0;                                             // index of the current window in it
  uint32_t cntNext = 0;                                         // windows in the staged next chunk
  uint32_t listAddr = warpBase + kWwList;                       // shared address of the current window's list entry
  uint32_t bufOff = 0;
  if (cnt)
  {
    stage_list(cs, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    fetch(listAddr, 0);
  }
  cp_async_commit();
  queue_claim_issue(queue, claimOld);
#pragma unroll 1
  while (cnt)
  {
    cp_async_wait<0>();
    __syncwarp();  // the staged list words were copied by other lanes
    const uint32_t recAddr = laneBase + bufOff;
    bool live, nodeDirty = false, fast;
    uint32_t info;
    Mat4 W;  // assigned on every path that reads it
    {
      const uint32_t e0 = lds32(listAddr);
      const uint32_t len = (lds32(listAddr + 4) & kWinSlotMask) - (e0 & kWinSlotMask);
      SC_ASSERT(len <= 32u && (e0 & kWinSlotMask) + len <= p.count);
      live = lane < len;
      fast = (e0 & kWinSlow) == 0u;
      info = live ? lds32(recAddr + 2048 - lane * 12) : 0u;
    }
    const uint32_t wl = info & kInfoDepthMask;
    const uint32_t maxL = __reduce_max_sync(0xffffffffu, wl);
    uint32_t dirtyM = 0, liveMask = 0, parentLane = 0;
    if (fast)
    {
      // (lanes beyond the window read stale bytes of their own buffer part; `live` masks everything derived from them)
      const float4 r0 = lds128(recAddr), r1 = lds128(recAddr + 512);
      const float sclZ = __uint_as_float(lds32(recAddr + 1024));
      const uint32_t fl = lds32(recAddr + 1536 + 12);
      nodeDirty = live && ((fl & dirtyMask) == dirtyWant);
      parentLane = (info >> kInfoParentShift) & 31u;
      // ---- 1. children inherit dirtiness level by level ----
      liveMask = __ballot_sync(0xffffffffu, live);
      dirtyM = __ballot_sync(0xffffffffu, nodeDirty);
      if (dirtyM != liveMask && dirtyM != 0u)
      {
        for (uint32_t l = 1; l <= maxL; ++l)
        {
          if (wl == l && ((dirtyM >> parentLane) & 1u)) nodeDirty = true;
          dirtyM = __ballot_sync(0xffffffffu, nodeDirty);
        }
      }
      // ---- 2. local matrices of everything that is recomputed; stored world matrices of the rest ----
      // A window with NOTHING dirty (static props; every window of a cull-only update) keeps its stored matrices: they
      // are copied global -> shared straight into this warp's matrix area, asynchronously, and read back - and
      // checked - below. A window that is PARTLY dirty is computed like a fully dirty one: every mutation of a
      // Transform (setLocal*, setParent, spawn, the fix-ups) stamps it dirty and dirtiness is inherited, so a clean
      // node's stored matrix IS what its record and its (clean) ancestors yield - recomputing it reproduces the same
      // value (up to the sign of a zero, see DESIGN.md "Parity"), costs no 64-byte read of the stored matrix and lets
      // the window use the static schedule of its level loop. Only the dirty lanes are written back and counted.
      const bool loadStored = dirtyM == 0u;
      const bool compute = loadStored ? false : live;  // this lane builds its matrix
      if (loadStored)
      {
        if (live)
        {
          const uint32_t q = (lds32(listAddr) & kWinSlotMask) + lane;
          const uint32_t own = laneBase + kWwMat;
          cp_async16s(own, p.w0 + q); cp_async16s(own + kMatC1, p.w1 + q);
          cp_async16s(own + kMatC2, p.w2 + q); cp_async16s(own + kMatC3, p.w3 + q);
        }
        cp_async_commit();
      }
      if (dirtyM != 0u)  // warp-uniform
      {
        const bool tame = trs_inputs_tame(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, sclZ);
        float sx, cx, sy, cy, sz, cz;
        sincos3_warp(compute && tame, r0.w, r1.x, r1.y, sx, cx, sy, cy, sz, cz);
        // every lane builds a matrix (lanes beyond the window from stale bytes, never used): no per-lane branch around sixteen
        // live registers. Roots: world == local.
        W = mat4_trs_from_sincos(r0.x, r0.y, r0.z, sx, cx, sy, cy, sz, cz, r1.z, r1.w, sclZ);
        fast = __all_sync(0xffffffffu, tame || !compute);
      }
      else
        W = mat4_identity();  // replaced by the stored matrices below (lanes beyond the window never use theirs)
    }
    // ---- what comes next for this warp: the next window of the chunk, or the first one of the staged next chunk ----
    const uint32_t a = lds32(listAddr) & kWinSlotMask;  // this window's start: the list slot may be recycled below
    uint32_t nextList = listAddr + 4u;
    if (pos == 1u)
    {
      // the other list slot is dead since the previous chunk's last window finished: take the claim that has been
      // in flight for a whole chunk, stage that chunk's window starts there, and put the next claim in flight
      uint32_t csNext;
      claim_to_chunk(queue_claim_result(claimOld), total, k1, csNext, cntNext);
      if (cntNext) stage_list(csNext, ((listAddr - warpBase - kWwList) < kWwListSlot) ? 1u : 0u);
      queue_claim_issue(queue, claimOld);
    }
    ++pos;
    if (pos == cnt)
    {
      nextList = warpBase + kWwList + (((listAddr - warpBase - kWwList) < kWwListSlot) ? kWwListSlot : 0u);
      cnt = cntNext;
      cntNext = 0u;
      pos = 0u;
    }
    if (cnt) fetch(nextList, bufOff ^ kWsBuf);  // in flight while this window is composed and culled
    cp_async_commit();
    if (fast)
    {
      // ---- 3. parent.world * local level by level, one lane PAIR per child ----
      const bool compose = maxL != 0u && dirtyM != 0u;
      const bool loadStored = dirtyM == 0u;
      if (compose || loadStored)
      {
        const uint32_t own = laneBase + kWwMat;
        if (!loadStored)
        {
          sts128(own, W.c0); sts128(own + kMatC1, W.c1); sts128(own + kMatC2, W.c2); sts128(own + kMatC3, W.c3);
        }
        else
          cp_async_wait<1>();  // the stored matrices have landed; the next window's records may still fly
        __syncwarp();
        // one product: this lane computes columns 2h, 2h+1 (h = lane & 1) of child = parent * child, in place
        auto level_item = [&](uint32_t child, uint32_t par)
        {
          const uint32_t cAddr = warpBase + child * 16u + (lane & 1u) * kMatC2;
          const uint32_t pAddr = warpBase + par * 16u;
          const float4 P0 = lds128(pAddr + kWwMat), P1 = lds128(pAddr + kWwMat + kMatC1), P2 = lds128(pAddr + kWwMat + kMatC2);
          const float4 La = lds128(cAddr + kWwMat), Lb = lds128(cAddr + kWwMat + 512);
          float4 oa, ob;
          {
            // rows x, y of both columns as register pairs (FMUL2 products, FFMA2 sums: fadd2_rn); row z stays scalar
            // (pairing it across the two columns costs more register moves than it saves). Every sum is
            // (p0 + p1) + p2 per element like sum3_ref
            const float2 a0 = fmul2_rn(P0.x, P0.y, La.x), a1 = fmul2_rn(P1.x, P1.y, La.y), a2 = fmul2_rn(P2.x, P2.y, La.z);
            const float2 b0 = fmul2_rn(P0.x, P0.y, Lb.x), b1 = fmul2_rn(P1.x, P1.y, Lb.y), b2 = fmul2_rn(P2.x, P2.y, Lb.z);
            const float2 sa = fadd2_rn(fadd2_rn(a0, a1), a2);
            float2 sb = fadd2_rn(fadd2_rn(b0, b1), b2);
            float2 sz;
            sz.x = __fadd_rn(__fadd_rn(__fmul_rn(P0.z, La.x), __fmul_rn(P1.z, La.y)), __fmul_rn(P2.z, La.z));
            sz.y = __fadd_rn(__fadd_rn(__fmul_rn(P0.z, Lb.x), __fmul_rn(P1.z, Lb.y)), __fmul_rn(P2.z, Lb.z));
            if (lane & 1u)
            {
              const float4 P3 = lds128(pAddr + kWwMat + kMatC3);  // column 3: + parent translation (p[r][3] * 1)
              sb = fadd2_rn(sb, make_float2(P3.x, P3.y));
              sz.y = __fadd_rn(sz.y, P3.z);
            }
            oa = make_float4(sa.x, sa.y, sz.x, La.w);
            ob = make_float4(sb.x, sb.y, sz.y, Lb.w);
          }
          sts128(cAddr + kWwMat, oa);
          sts128(cAddr + kWwMat + 512, ob);
        };
        if (!compose)
        {
          // nothing to multiply (no hierarchy in this window, or nothing dirty): the matrices only pass through
        }
        else if ((lds32(listAddr) & kWinNoStatic) == 0u)
        {
          // every node is computed: who multiplies what is a function of the topology alone and was laid down by
          // k_build_windows in the slotInfo words (6 bits per level)
#pragma unroll
          for (uint32_t l = 0; l < 3u; ++l)
          {
            if (l < maxL)
            {
              // the child's parent lane comes from the child lane's register (one SHFL by the whole warp instead of
              // address + LDS + shift + mask by the scheduled lanes)
              const uint32_t sch = info >> (kInfoSchedShift + 6u * l);
              const uint32_t child = sch & 31u;
              const uint32_t par = __shfl_sync(0xffffffffu, parentLane, child);
              if (sch & 32u) level_item(child, par);
              __syncwarp();
            }
          }
        }
        else
        {
          // windows without a static schedule (deeper than three levels, or a level wider than half the window): the
          // nodes of each level are compacted into a schedule at run time, 16 per round
          for (uint32_t l = 1; l <= maxL; ++l)
          {
            const bool mine = live && wl == l;
            const uint32_t m = __ballot_sync(0xffffffffu, mine);
            if (m == 0u) continue;
            const uint32_t cnt = __popc(m);
            if (mine) sts16(warpBase + kWwSched + __popc(m & ((1u << lane) - 1u)) * 2u, lane | (parentLane << 8));
            __syncwarp();
            for (uint32_t r = 0; r < cnt; r += 16u)
            {
              const uint32_t idx = r + (lane >> 1);
              if (idx < cnt)
              {
                const uint32_t e = lds16(warpBase + kWwSched + idx * 2u);
                level_item(e & 31u, e >> 8);
              }
              __syncwarp();
            }
          }
        }
        W.c0 = lds128(own); W.c1 = lds128(own + kMatC1); W.c2 = lds128(own + kMatC2); W.c3 = lds128(own + kMatC3);
        // the structured products are value-exact iff every parent translation was finite; a non-finite one
        // propagates into the translation of all its descendants, so one test of the results covers all levels
        // (a stored matrix that is kept must be affine for the structured sphere and products to be exact)
        const float mag = fabsf(W.c3.x) + fabsf(W.c3.y) + fabsf(W.c3.z);
        fast = __all_sync(0xffffffffu, !live || (loadStored ? mat4_is_affine(W) : mag < __int_as_float(0x7f800000)));
      }
    }
    if (fast) finish_window<kViews, true>(p, vp, a, lane, recAddr, live, nodeDirty, W, order, nRecomputed, accCand, warpBase + kWwFav);
    else
    {
      // redone by k_update_win_slow (generic path), which also culls and counts it
      uint32_t slot = 0;
      if (lane == 0) slot = atomicAdd(queue + 1, 1u);
      SC_ASSERT(slot < total);
      if (lane == 0) slowList[slot] = a | (((lds32(listAddr + 4) & kWinSlotMask) - a) << 24);
    }
    listAddr = nextList;
    bufOff ^= kWsBuf;
  }
  if (nRecomputed) warp_reds_add(sBase + kWsRecomputed, nRecomputed);
  if (accCand) warp_reds_add(sBase + kWsRecomputed + 4, accCand);
  __syncthreads();
  if (tid < 2)
  {
    const uint32_t r = lds32(sBase + kWsRecomputed + 4 * tid);
    if (r) atomicAdd(p.acc + (tid == 0 ? kAccRecomputed : kAccCand), r);
  }
}

// The generic path for the windows k_update_win put aside (parents outside the window, cycles, non-affine or
// non-finite matrices, hostile TRS values): exact for any input, not tuned. One warp per window, strided over the list.
template <int kViews>
__global__ void __launch_bounds__(kWinBlock) k_update_win_slow(const __grid_constant__ UpdateParams p,
                                                               const __grid_constant__ ViewPlanes vp,
                                                               const uint32_t* __restrict__ slotInfo,
                                                               const uint32_t* __restrict__ slowCount,
                                                               const uint32_t* __restrict__ slowList)
{
  uint32_t sBase;
  asm("mov.u32 %0, scgpu_win_smem;" : "=r"(sBase));
  asm volatile("" ::"l"(scgpu_win_smem));  // keeps the array emitted: every other access goes through sBase
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid < 2) reinterpret_cast<uint32_t*>(scgpu_win_smem + kWsRecomputed)[tid] = 0u;  // (also keeps the symbol referenced)
  __syncthreads();
  pdl_trigger();
  pdl_wait();  // the list and its length are written by k_update_win
  const uint32_t nSlow = *slowCount;
  const uint32_t laneBase = sBase + warp * kWwSize + lane * 16;
  uint32_t order = 0, nRecomputed = 0, accCand = 0;
#pragma unroll 1
  for (uint32_t k = blockIdx.x * kWinWarps + warp; k < nSlow; k += gridDim.x * kWinWarps)
  {
    const uint32_t e = slowList[k];
    const uint32_t a = e & 0xFFFFFFu, len = e >> 24;
    const bool live = lane < len;
    __syncwarp();
    if (live)
    {
      sts128(laneBase, p.rec0[a + lane]); sts128(laneBase + 512, p.rec1[a + lane]);
      sts128(laneBase + 1024, p.rec2[a + lane]); sts128(laneBase + 1536, p.rec3[a + lane]);
    }
    const uint32_t info = live ? slotInfo[a + lane] : 0u;
    float4 wb[4];
    const bool nodeDirty = window_slow(p, a, len, info, wb) != 0u;
    finish_window<kViews, false>(p, vp, a, lane, laneBase, live, nodeDirty, xs_load(wb), order, nRecomputed, accCand, 0u);
  }
  if (nRecomputed) warp_reds_add(sBase + kWsRecomputed, nRecomputed);
  if (accCand) warp_reds_add(sBase + kWsRecomputed + 4, accCand);
  __syncthreads();
  if (tid < 2)
  {
    const uint32_t r = lds32(sBase + kWsRecomputed + 4 * tid);
    if (r) atomicAdd(p.acc + (tid == 0 ? kAccRecomputed : kAccCand), r);
  }
}

// ---- SURVEY.md 8(f) N4 (second mat4_trs caller): the world editor's BuildDrawItems --------------------------------
// tools/world_editor/editor_core/editor_core.cpp:242-264: for every entity with a mesh and a material handle,
// ScRenderDrawItem{mesh, material (64-bit handles), model = mat4_trs(position, rotation, scale), flags = 0}, in
// document order.
// Pass 1 counts the kept entities per 256-entity block, k_scan_tiles turns the counts into offsets, pass 2 builds the
// matrices (same warp-cooperative sincos + structured TRS as the frame kernels, dense fallback for hostile values)
// and writes the 88-byte items at their stable positions.
__global__ void __launch_bounds__(kBlock) k_editor_count(const uint64_t* __restrict__ mesh, const uint64_t* __restrict__ material,
                                                         uint32_t n, uint32_t* __restrict__ blockCounts)
{
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  const bool keep = i < n && mesh[i] != 0ull && material[i] != 0ull;
  const uint32_t c = __syncthreads_count(keep);
  if (threadIdx.x == 0) blockCounts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(kBlock) k_editor_write(const float* __restrict__ trs9, const uint64_t* __restrict__ mesh,
                                                         const uint64_t* __restrict__ material, uint32_t n,
                                                         const uint32_t* __restrict__ blockOffsets, uint32_t* __restrict__ out22)
{
  __shared__ uint32_t sWarp[kBlock / 32];
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const bool keep = i < n && mesh[i] != 0ull && material[i] != 0ull;
  float t[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) t[k] = keep ? trs9[(size_t)i * 9 + k] : 0.0f;
  const bool tame = trs_inputs_tame(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]);
  float sx, cx, sy, cy, sz, cz;
  sincos3_warp(keep && tame, t[3], t[4], t[5], sx, cx, sy, cy, sz, cz);
  Mat4 W = mat4_identity();
  if (keep)
  {
    if (tame) W = mat4_trs_from_sincos(t[0], t[1], t[2], sx, cx, sy, cy, sz, cz, t[6], t[7], t[8]);
    else W = mat4_trs_dense_call(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]);
  }
  // stable position: block offset + rank inside the block
  const uint32_t m = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) sWarp[warp] = __popc(m);
  __syncthreads();
  uint32_t pos = blockOffsets[blockIdx.x] + __popc(m & ((1u << lane) - 1u));
  for (uint32_t w = 0; w < warp; ++w) pos += sWarp[w];
  if (!keep) return;
  uint32_t* o = out22 + (size_t)pos * 22u;  // ScRenderDrawItem: u64 mesh, u64 material, model[16], flags, pad (88 bytes)
  const uint64_t mh = mesh[i], th = material[i];
  o[0] = (uint32_t)mh; o[1] = (uint32_t)(mh >> 32); o[2] = (uint32_t)th; o[3] = (uint32_t)(th >> 32);
  const float4 c[4] = { W.c0, W.c1, W.c2, W.c3 };
#pragma unroll
  for (int k = 0; k < 4; ++k)
  {
    o[4 + 4 * k] = __float_as_uint(c[k].x); o[5 + 4 * k] = __float_as_uint(c[k].y);
    o[6 + 4 * k] = __float_as_uint(c[k].z); o[7 + 4 * k] = __float_as_uint(c[k].w);
  }
  o[20] = 0u; o[21] = 0u;
}

// ---- K3a: exclusive scan of the per-tile counts, one CTA per row (view) -----------------------------------
__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t* __restrict__ tileCounts,
                                                     uint32_t* __restrict__ tileOffsets, uint32_t* __restrict__ totals,
                                                     uint32_t numTiles)
{
  // 16 consecutive counts per thread: 16 Ki tiles (a full context) are scanned in ONE pass with two barriers
  constexpr uint32_t kPer = 16;
  __shared__ uint32_t sWarp[32];
  __shared__ uint32_t sCarry;
  const uint32_t row = blockIdx.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  pdl_trigger();
  pdl_wait();
  const uint32_t* in = tileCounts + (size_t)row * numTiles;
  uint32_t* out = tileOffsets + (size_t)row * numTiles;
  if (tid == 0) sCarry = 0;
  __syncthreads();
  for (uint32_t start = 0; start < numTiles; start += 1024 * kPer)
  {
    const uint32_t base = start + tid * kPer;
    uint32_t v[kPer], sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < kPer; ++j)
    {
      v[j] = (base + j < numTiles) ? in[base + j] : 0u;
      sum += v[j];
    }
    uint32_t x = sum;
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
    uint32_t run = carry + (warp ? sWarp[warp - 1] : 0u) + x - sum;  // exclusive prefix of this thread's first count
#pragma unroll
    for (uint32_t j = 0; j < kPer; ++j)
    {
      if (base + j < numTiles) out[base + j] = run;
      run += v[j];
    }
    __syncthreads();
    if (tid == 1023) sCarry = run;
    __syncthreads();
  }
  if (tid == 0) totals[row] = sCarry;
}

// ---- K3: compaction of the visible sets in POOL ORDER (CullingState::visible / ::culled, .cpp:1273-1280) ----------
// The frame kernels left one bit per (view, pool rank) and, per CHUNK of 32 Ki ranks, the number of bits they set. One
// pass over the bitmaps in rank order yields the reference's lists whatever the device layout is:
// rank -> perm[rank] = slot -> entity[slot]. A single launch, one CTA per chunk (1024 words of every plane, one uint4
// per thread and plane, held in registers from the load to the last use), no dependency between CTAs:
//   1. output position of the chunk per row (row = view, + one culled row per view when the candidate plane is
//      present) = sum of the chunk counts before it (<= 511 values per row, L2 resident),
//   2. block scan of the per-thread popcounts; the set bit positions (= ranks) are written at chunk position + scan,
//      one non-empty word per warp step with the 32 lanes taking its 32 bits (coalesced stores, and a dense word costs
//      what a sparse one costs); k_resolve_lists then maps rank -> slot -> entity handle for all entries in parallel,
//   3. the words found set are cleared and the chunk counts of the OTHER frame parity are zeroed, so the next frame
//      starts clean without a memset; the CTA of the last chunk also writes the frame totals and zeroes the
//      accumulators and the work queue of k_update_win.
// Chunks of 32 Ki ranks (not one long segment per SM) because the visible instances cluster around the camera: the few
// hundred thousand ranks that hold nearly all set bits spread over ten or more SMs this way.
// Replaces round 1's per-tile counts + k_scan_tiles + k_scatter_visible + k_scatter_culled + two memsets.
constexpr uint32_t kCompactThreads = 256;
constexpr uint32_t kCompactRows = 2 * kMaxViews;
constexpr uint32_t kCompactChunkWords = kCompactThreads * 4;  // one uint4 per thread and plane
static_assert(kCompactChunkWords * 32u == kChunkRanks, "one CTA per chunk");

struct CompactParams
{
  uint32_t* bits;                        // [nViews (+1)][bitWords]
  const uint32_t* chunkCounts;           // [nViews (+1)][chunkStride], this frame's parity
  uint32_t* chunkCountsNext;             // the other parity: zeroed here for the next frame
  uint32_t* outSlot[kMaxViews];          // receives the visible RANKS in ascending order (k_resolve_lists makes slots of them)
  uint32_t* culledEntity[kMaxViews];     // valid when culled != 0; receives ranks likewise
  uint32_t* acc;                         // kAcc*
  uint32_t* totals;                      // out, kTotalsWords
  uint32_t* totalsHost;                  // the same, in pinned host memory: written over PCIe by the last chunk's CTA, so
                                         // that no copy-engine operation sits in the frame's kernel chain
  uint32_t bitWords, chunkStride;        // words per plane / counts per row
  uint32_t nWords;                       // words that can hold a set bit: ceil(live count / 32), rounded up to 4
  uint32_t nViews, culled;
  uint32_t listCap;                      // entries every output list can hold
};

__device__ __forceinline__ uint32_t popc4(uint4 w) { return __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w); }
__device__ __forceinline__ uint4 andn4(uint4 a, uint4 b) { return make_uint4(a.x & ~b.x, a.y & ~b.y, a.z & ~b.z, a.w & ~b.w); }

// exclusive scan of one 64-bit value per thread over the CTA (kCompactThreads threads); *total = sum over the CTA
__device__ __forceinline__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t* sWarp, uint32_t tid, uint64_t* total)
{
  constexpr uint32_t kWarps = kCompactThreads / 32;
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
  uint64_t before = 0, all = 0;
#pragma unroll
  for (uint32_t k = 0; k < kWarps; ++k)
  {
    const uint64_t t = sWarp[k];
    if (k < warp) before += t;
    all += t;
  }
  *total = all;
  __syncthreads();  // sWarp is reused by the next scan
  return before + x - v;
}

// the set bits of one plane row held by this warp (4 words per lane, consecutive lanes = consecutive words) as ranks at
// out[first position of the word + rank order inside the word]; one warp step per non-empty word
__device__ __forceinline__ void emit_ranks_warp(uint32_t* __restrict__ out, uint4 b, uint32_t pos, uint32_t w, uint32_t lane, uint32_t cap)
{
  (void)cap;
  const uint32_t word[4] = { b.x, b.y, b.z, b.w };
  const uint32_t ltMask = (1u << lane) - 1u;
#pragma unroll
  for (uint32_t k = 0; k < 4u; ++k)
  {
    uint32_t nz = __ballot_sync(0xffffffffu, word[k] != 0u);
    while (nz)
    {
      const uint32_t src = __ffs(nz) - 1u;
      nz &= nz - 1u;
      const uint32_t wv = __shfl_sync(0xffffffffu, word[k], src);
      const uint32_t wp = __shfl_sync(0xffffffffu, pos, src);
      const uint32_t wr = (__shfl_sync(0xffffffffu, w, src) + k) * 32u;
      SC_ASSERT(wp + __popc(wv) <= cap);
      if ((wv >> lane) & 1u) out[wp + __popc(wv & ltMask)] = wr + lane;
    }
    pos += __popc(word[k]);
  }
}

#ifdef SCGPU_COMPACT_TIMING
__device__ unsigned long long g_compactStamps[1024 * 8];
#define SC_STAMP(k)                                                                                                   \
  do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_compactStamps[blockIdx.x * 8 + (k)] = t_; } } while (0)
#else
#define SC_STAMP(k) do { } while (0)
#endif

template <int kViews>
__global__ void __launch_bounds__(kCompactThreads, 4) k_compact(const __grid_constant__ CompactParams p)
{
  constexpr uint32_t V = (uint32_t)kViews;
  __shared__ uint32_t sBase[2 * V];    // output position of this chunk's first entry per row: visible rows 0..V-1, culled rows V..2V-1
  __shared__ uint32_t sCount[2 * V];   // set bits of this chunk per row
  __shared__ uint64_t sWarp[kCompactThreads / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint32_t seg = blockIdx.x;
  const bool culled = p.culled != 0u;
  const uint32_t nRows = culled ? 2u * V : V;
  const uint32_t w = seg * kCompactChunkWords + tid * 4u;
  const bool mine = w < p.nWords;
  if (tid < 2u * V) { sBase[tid] = 0u; sCount[tid] = 0u; }
  SC_STAMP(0);
  pdl_wait();  // the frame kernels' bits and counters are complete and visible
  __syncthreads();
  pdl_trigger();  // k_resolve_lists may take its place on the SMs once every CTA of this grid is running
  SC_STAMP(1);

  // ---- this thread's four words of every plane (in flight while the chunk positions are added up)
  uint4 vis[V];
  uint4 cand = make_uint4(0u, 0u, 0u, 0u);
  if (culled && mine) cand = reinterpret_cast<const uint4*>(p.bits + (size_t)V * p.bitWords)[w >> 2];
#pragma unroll
  for (uint32_t v = 0; v < V; ++v)
  {
    vis[v] = make_uint4(0u, 0u, 0u, 0u);
    if (mine) vis[v] = reinterpret_cast<const uint4*>(p.bits + (size_t)v * p.bitWords)[w >> 2];
  }

  // ---- 1. sum of the chunk counts before this chunk, per row (culled = candidates - visible; the candidates are row V
  //         of the counts). A 16 Mi-slot context has 512 chunks: at most three counts per thread and row, all loads
  //         independent. The thread that meets t == seg contributes the chunk's own counts.
  {
    static_assert((1u << 24) / kChunkRanks <= 3u * kCompactThreads, "three rounds cover every chunk");
    uint32_t before[V + 1], own[V + 1];
#pragma unroll
    for (uint32_t v = 0; v <= V; ++v) { before[v] = 0u; own[v] = 0u; }
#pragma unroll
    for (uint32_t round = 0; round < 3u; ++round)
    {
      const uint32_t t = tid + round * kCompactThreads;
#pragma unroll
      for (uint32_t v = 0; v <= V; ++v)
      {
        if (t <= seg && (v < V || culled))
        {
          const uint32_t c = p.chunkCounts[(size_t)v * p.chunkStride + t];
          if (t < seg) before[v] += c; else own[v] = c;
        }
      }
    }
#pragma unroll
    for (uint32_t v = 0; v < V; ++v)
    {
      const uint32_t b = __reduce_add_sync(0xffffffffu, before[v]);
      const uint32_t o = __reduce_add_sync(0xffffffffu, own[v]);
      if (lane == 0 && b) atomicAdd(&sBase[v], b);
      if (lane == 0 && o) atomicAdd(&sCount[v], o);
      if (culled)
      {
        const uint32_t cb = __reduce_add_sync(0xffffffffu, before[V] - before[v]);
        const uint32_t co = __reduce_add_sync(0xffffffffu, own[V] - own[v]);
        if (lane == 0 && cb) atomicAdd(&sBase[V + v], cb);
        if (lane == 0 && co) atomicAdd(&sCount[V + v], co);
      }
    }
  }
  __syncthreads();
  SC_STAMP(2);
#ifdef SCGPU_CHECKED
  {
    // the per-chunk counts of the frame kernels must be the popcounts of their bits
    __shared__ uint32_t sPop[2 * V];
    if (tid < 2u * V) sPop[tid] = 0u;
    __syncthreads();
#pragma unroll
    for (uint32_t v = 0; v < V; ++v)
    {
      const uint32_t c = __reduce_add_sync(0xffffffffu, popc4(vis[v]));
      if (lane == 0 && c) atomicAdd(&sPop[v], c);
      const uint32_t d = __reduce_add_sync(0xffffffffu, culled ? popc4(andn4(cand, vis[v])) : 0u);
      if (lane == 0 && d) atomicAdd(&sPop[V + v], d);
    }
    __syncthreads();
    if (tid < nRows) SC_ASSERT(sPop[tid] == sCount[tid]);
    __syncthreads();
  }
#endif
  if (seg == gridDim.x - 1u)
  {
    // frame totals; accumulators and the window queue are left zeroed for the next frame
    uint32_t val = 0, at = 0xFFFFFFFFu;
    if (tid < V) { at = tid; val = sBase[tid] + sCount[tid]; }
    if (tid == 32) { at = V; val = p.acc[kAccCand]; p.acc[kAccCand] = 0u; }
    if (tid == 33) { at = kMaxViews + 1; val = p.acc[kAccRecomputed]; p.acc[kAccRecomputed] = 0u; }
    if (tid == 34) { at = kMaxViews + 2; val = p.acc[kAccQueueSlow]; p.acc[kAccQueueNext] = 0u; p.acc[kAccQueueSlow] = 0u; }
    if (at != 0xFFFFFFFFu)
    {
      p.totals[at] = val;
      p.totalsHost[at] = val;
      __threadfence_system();
    }
  }

  // ---- 2. ranks out: three rows share one 64-bit block scan (21-bit fields: a chunk holds 2^15 ranks)
#pragma unroll
  for (uint32_t r0 = 0; r0 < 2u * V; r0 += 3u)
  {
    if (r0 >= nRows) break;
    uint32_t any = 0;
#pragma unroll
    for (uint32_t j = 0; j < 3u; ++j) any |= (r0 + j < nRows) ? sCount[r0 + j] : 0u;
    if (!any) continue;  // block-uniform: nothing set in this chunk for these rows
    uint4 bitsOf[3];
    uint64_t packed = 0;
#pragma unroll
    for (uint32_t j = 0; j < 3u; ++j)
    {
      const uint32_t r = r0 + j;  // compile-time after unrolling: vis[] is indexed statically
      uint4 b = make_uint4(0u, 0u, 0u, 0u);
      if (r < 2u * V && r < nRows) b = r < V ? vis[r < V ? r : 0] : andn4(cand, vis[r >= V && r < 2u * V ? r - V : 0]);
      bitsOf[j] = b;
      packed |= (uint64_t)popc4(b) << (21u * j);
    }
    uint64_t total;
    const uint64_t excl = block_exclusive_scan_u64(packed, sWarp, tid, &total);
#pragma unroll
    for (uint32_t j = 0; j < 3u; ++j)
    {
      const uint32_t r = r0 + j;
      if (r < nRows && ((total >> (21u * j)) & 0x1FFFFFu) != 0u)  // warp-uniform (block-uniform)
      {
        uint32_t* out = r < V ? p.outSlot[r < V ? r : 0] : p.culledEntity[r >= V && r < 2u * V ? r - V : 0];
        emit_ranks_warp(out, bitsOf[j], sBase[r] + (uint32_t)((excl >> (21u * j)) & 0x1FFFFFu), w, lane, p.listCap);
      }
    }
  }

  SC_STAMP(3);
  // ---- 3. clean bitmaps and chunk counts for the next frame: only words that were set are written
  if (mine)
  {
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (uint32_t v = 0; v < V; ++v)
      if ((vis[v].x | vis[v].y | vis[v].z | vis[v].w) != 0u)
        reinterpret_cast<uint4*>(p.bits + (size_t)v * p.bitWords)[w >> 2] = zero;
    if (culled && (cand.x | cand.y | cand.z | cand.w) != 0u)
      reinterpret_cast<uint4*>(p.bits + (size_t)V * p.bitWords)[w >> 2] = zero;
  }
  // the other parity was read by the previous frame's k_compact (complete) and is written by the next frame's kernels
  // (not started): every row of it, ALL chunks (the pool may have been larger then than it is now, and another view count)
  if (tid <= kMaxViews)
    for (uint32_t c = seg; c < p.chunkStride; c += gridDim.x) p.chunkCountsNext[(size_t)tid * p.chunkStride + c] = 0u;
  SC_STAMP(4);
}

// Mailbox in the gather root's HBM (scgpu_peer.cuh documents the protocol; the struct lives here because the last
// kernel of the frame writes into it).
constexpr uint32_t kPeerHeaderWords = 16;  // 64 B
constexpr uint32_t kPeerFlag = 0, kPeerCounts = 1, kPeerOverflow = kPeerCounts + kMaxViews + 2;
constexpr long long kPeerSpinClocks = 4000000000ll;  // ~2 s at 1.9 GHz: a dead peer must not hang the box

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

struct ResolveParams
{
  const uint32_t* perm;                  // rank -> slot
  const uint32_t* entity;                // slot -> handle
  const uint32_t* totals;                // frame totals written by k_compact
  uint32_t* outEntity[kMaxViews];
  uint32_t* outSlot[kMaxViews];          // in: ranks, out: slots
  uint32_t* culledEntity[kMaxViews];     // in: ranks, out: entity handles
  uint32_t nViews, culled;
  uint32_t live, extent;                 // checked build: bounds of ranks and slots
  // multi-GPU, peer gather enabled: the entity handles go straight into this rank's slot of the ROOT's mailbox too
  // (NVLink stores from the kernel that produces them), and the last CTA publishes counts + flag
  PeerBox box;
  uint32_t peer, seq, rank, isRoot;
  uint32_t* done;                        // local ticket counter (zero between launches)
  uint32_t* error;                       // local error words, one per gather parity
};

// rank -> (slot, entity handle) for every entry of every list, grid-stride: one independent chain of two loads per
// entry. With the peer gather enabled this is also the producer side of the gather — compute and peer store in one
// kernel: every handle is stored locally AND into the root's mailbox the moment it is known.
__global__ void __launch_bounds__(kBlock) k_resolve_lists(const __grid_constant__ ResolveParams p)
{
  __shared__ uint32_t sOff[kMaxViews + 1];
  __shared__ uint32_t sLast, sStale;
  pdl_trigger();  // (the root's k_peer_wait only polls the mailbox: it may take its place already)
  pdl_wait();
  const uint32_t stride = gridDim.x * kBlock;
  const uint32_t cand = p.totals[p.nViews];
  const uint32_t parity = p.seq & 1u;
  bool remote = false;
  uint32_t* mail = nullptr;
  if (p.peer)
  {
    if (threadIdx.x == 0)
    {
      sStale = 0u;
      if (blockIdx.x == 0) p.error[parity ^ 1u] = 0u;  // the error word describes ONE gather: clear the next one's
      if (p.isRoot)
      {
        // everything the root enqueued before this frame's kernels (the readers of gather seq-1 included) has completed
        if (blockIdx.x == 0) st_release_sys(p.box.progress(), p.seq);
      }
      else
      {
        // the buffer of this parity was last used by gather seq-2: wait until the root is past gather seq-1's start
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(p.box.progress()) - (p.seq - 1u)) < 0)
        {
          // the root may still be reading this buffer: write NOTHING into it, report the frame as failed instead
          if (clock64() - t0 > kPeerSpinClocks) { atomicOr(p.error + parity, 1u); sStale = 1u; break; }
          __nanosleep(200);
        }
      }
      uint32_t off = 0;
      for (uint32_t v = 0; v < p.nViews; ++v) { sOff[v] = off; off += p.totals[v]; }
      sOff[p.nViews] = off;
    }
    __syncthreads();
    remote = sStale == 0u && sOff[p.nViews] <= p.box.cap;
    mail = p.box.payload(p.rank, parity);
  }
  for (uint32_t v = 0; v < p.nViews; ++v)
  {
    const uint32_t nVis = p.totals[v];
    uint32_t* __restrict__ outS = p.outSlot[v];
    uint32_t* __restrict__ outE = p.outEntity[v];
    uint32_t* __restrict__ outM = remote ? mail + sOff[v] : nullptr;
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < nVis; i += stride)
    {
      SC_ASSERT(outS[i] < p.live);
      const uint32_t slot = __ldg(p.perm + outS[i]);
      SC_ASSERT(slot < p.extent);
      outS[i] = slot;
      const uint32_t e = __ldg(p.entity + slot);
      outE[i] = e;
      if (outM) outM[i] = e;
    }
    if (p.culled)
    {
      const uint32_t nCul = cand - nVis;
      uint32_t* __restrict__ outC = p.culledEntity[v];
      for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < nCul; i += stride)
        outC[i] = __ldg(p.entity + __ldg(p.perm + outC[i]));
    }
  }
  if (p.peer)
  {
    __threadfence_system();  // this thread's remote stores are visible system-wide before the ticket
    __syncthreads();
    if (threadIdx.x == 0) sLast = (atomicAdd(p.done, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (sLast)
    {
      // (the header is 64 B that the root reads only after the flag: safe to write even when the payload was not)
      uint32_t* h = p.box.header(p.rank, parity);
      if (threadIdx.x < kMaxViews + 2) h[kPeerCounts + threadIdx.x] = p.totals[threadIdx.x];
      if (threadIdx.x == 0) h[kPeerOverflow] = sStale ? 2u : (sOff[p.nViews] > p.box.cap ? 1u : 0u);
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0)
      {
        *p.done = 0u;
        st_release_sys(h + kPeerFlag, p.seq);
      }
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
  uint32_t* sparse;      // Entity::index() -> slot+1 (ComponentPool sparse array, sc_ecs.h:199-277; here: DEVICE slot)
  uint32_t sparseSize;
  uint32_t* rank;        // slot -> dense index in the reference's Transform pool (ComponentPool::m_denseEntities)
  uint32_t* perm;        // dense index -> slot
  uint8_t* tileDirty;    // [1 + tiles + 1][4] bytes, tile t at [t + 1]: since k_build_windows last cut the tile into windows
                         // a parentSlot changed (or a slot was spawned) [0] anywhere in it, [1] in its first 2 kHalo slots,
                         // [2] in its last kHalo slots; set by whoever writes parentSlot, cleared by k_flatten_windows
};
// marks the tile of slot s for the next window cut (benign races: everybody stores the same bytes)
__device__ __forceinline__ void mark_tile(const SceneArrays& a, uint32_t s)
{
  uint8_t* d = a.tileDirty + ((size_t)(s / kTile) + 1u) * 4u;
  const uint32_t off = s % kTile;
  d[0] = 1;
  if (off < 64u) d[1] = 1;          // 2 * kHalo: what the tile before this one looks at
  if (off >= kTile - 32u) d[2] = 1;  // kHalo: what the tile behind this one looks at
}

// slot of the j-th element of a spawn batch: a run of fresh slots, or wherever the host's layout put it (scgpu_layout.h)
__device__ __forceinline__ uint32_t spawn_slot(uint32_t slot0, const uint32_t* __restrict__ slotOf, uint32_t j)
{
  return slotOf ? slotOf[j] : slot0 + j;
}

// World::add<Transform> + setLocal (+Bounds/RenderMesh) for n new slots [slot0, slot0+n)
__global__ void __launch_bounds__(kBlock) k_spawn(SceneArrays a, uint32_t slot0, const uint32_t* __restrict__ slotOf, uint32_t rank0,
                                                  uint32_t n, const uint32_t* __restrict__ entity, const uint32_t* __restrict__ parent,
                                                  const float* __restrict__ trs9, const float* __restrict__ aabb6,
                                                  const uint32_t* __restrict__ meshMat2, const uint32_t* __restrict__ flags,
                                                  uint32_t stamp)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t s = spawn_slot(slot0, slotOf, j);
  SC_ASSERT(a.entity[s] == kNone || a.entity[s] == 0u);  // a hole or a never-used slot
  a.rank[s] = rank0 + j;
  a.perm[rank0 + j] = s;
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
  const uint32_t f = (flags ? (flags[j] & (kFlagBounds | kFlagMesh)) : (kFlagBounds | kFlagMesh)) | kFlagLive | (stamp << kStampShift);
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
  mark_tile(a, s);
  a.meshMat[s] = meshMat2 ? make_uint2(meshMat2[(size_t)j * 2], meshMat2[(size_t)j * 2 + 1]) : make_uint2(0u, 0u);
  const uint32_t idx = e & 0xFFFFFFu;
  if (idx < a.sparseSize) a.sparse[idx] = s + 1u;
}

// ---- procedural sectors spawned on the device ----------------------------------------------------
// generateSectorSpawnsStatic (src/engine/world/sc_world_partition.cpp:105-169) + the per-record World::add loop of
// pumpCompletedLoads (:923-954), writing the SoA records directly: the host only creates the entity handles. Same
// integer hash (mix32 / hashCoordSeed / rand01, :34-57) and the same float expressions, unfused.
struct SectorGen  // == ScGpuSectorGen without struct_size
{
  float sectorSizeMeters;
  uint32_t seed, propsMin, propsMax, includeGround;
  uint32_t meshCube, meshTriangle, matUnlit, matChecker, matTest;
};

__host__ __device__ __forceinline__ uint32_t sg_mix32(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t sg_hash_coord(uint32_t seed, int32_t cx, int32_t cz)
{
  uint32_t h = seed;
  h ^= sg_mix32((uint32_t)cx * 73856093u);
  h ^= sg_mix32((uint32_t)cz * 19349663u);
  return sg_mix32(h + 0x9e3779b9u);
}
__host__ __device__ __forceinline__ uint32_t sg_prop_count(const SectorGen& g, int32_t cx, int32_t cz)
{
  const uint32_t rng = sg_hash_coord(g.seed, cx, cz);
  const uint32_t range = g.propsMax - g.propsMin + 1u;
  return g.propsMin + (range > 0 ? (sg_mix32(rng) % range) : 0u);
}
__device__ __forceinline__ float sg_rand01(uint32_t& state)
{
  state = sg_mix32(state + 0x6d2b79f5u);
  return __fdiv_rn((float)(state & 0x00FFFFFFu), 16777215.0f);
}
__device__ __forceinline__ float sg_lerp(float a, float b, float t) { return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t)); }

// one CTA per sector, one thread per SpawnRecord (ground plane first, then the props in generation order)
__global__ void __launch_bounds__(64) k_spawn_sectors(SceneArrays a, SectorGen g, uint32_t slot0, const uint32_t* __restrict__ slotOf,
                                                      uint32_t rank0, const int32_t* __restrict__ coordXZ,
                                                      const uint32_t* __restrict__ first, const uint32_t* __restrict__ entity,
                                                      uint32_t stamp)
{
  const uint32_t sec = blockIdx.x;
  const int32_t cx = coordXZ[2 * sec], cz = coordXZ[2 * sec + 1];
  const uint32_t base = first[sec], n = first[sec + 1] - base;
  const float size = g.sectorSizeMeters;
  const float minX = __fmul_rn((float)cx, size), minZ = __fmul_rn((float)cz, size);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
  {
    float px, py, pz, ry = 0.0f, sx, sy, sz;
    uint32_t mesh, mat;
    if (g.includeGround && i == 0)
    {
      px = __fadd_rn(minX, __fmul_rn(size, 0.5f)); py = -0.55f; pz = __fadd_rn(minZ, __fmul_rn(size, 0.5f));
      sx = size; sy = 0.10f; sz = size;
      mesh = g.meshCube; mat = g.matUnlit;
    }
    else
    {
      const uint32_t prop = i - (g.includeGround ? 1u : 0u);
      uint32_t rng = sg_hash_coord(g.seed, cx, cz);
      for (uint32_t k = 0; k < prop * 8u; ++k) rng = sg_mix32(rng + 0x6d2b79f5u);  // 8 draws per earlier prop
      const float pad = 1.0f;
      px = sg_lerp(__fadd_rn(minX, pad), __fsub_rn(__fadd_rn(minX, size), pad), sg_rand01(rng));
      pz = sg_lerp(__fadd_rn(minZ, pad), __fsub_rn(__fadd_rn(minZ, size), pad), sg_rand01(rng));
      sx = sg_lerp(0.4f, 1.9f, sg_rand01(rng));
      sy = sg_lerp(0.5f, 3.2f, sg_rand01(rng));
      sz = sg_lerp(0.4f, 1.9f, sg_rand01(rng));
      py = __fmul_rn(sy, 0.5f);
      ry = __fmul_rn(sg_rand01(rng), 3.1415926535f * 2.0f);
      const float m = sg_rand01(rng);
      mat = (m < 0.40f) ? g.matChecker : ((m < 0.80f) ? g.matTest : g.matUnlit);
      mesh = (sg_rand01(rng) < 0.90f) ? g.meshCube : g.meshTriangle;
    }
    const uint32_t s = spawn_slot(slot0, slotOf, base + i);
    a.rank[s] = rank0 + base + i;
    a.perm[rank0 + base + i] = s;
    const uint32_t f = (kFlagBounds | kFlagMesh | kFlagLive) | (stamp << kStampShift);
    a.rec[0][s] = make_float4(px, py, pz, 0.0f);
    a.rec[1][s] = make_float4(ry, 0.0f, sx, sy);
    a.rec[2][s] = make_float4(sz, -0.5f, -0.5f, -0.5f);
    a.rec[3][s] = make_float4(0.5f, 0.5f, 0.5f, __uint_as_float(f));
    a.world[0][s] = make_float4(1.f, 0.f, 0.f, 0.f);
    a.world[1][s] = make_float4(0.f, 1.f, 0.f, 0.f);
    a.world[2][s] = make_float4(0.f, 0.f, 1.f, 0.f);
    a.world[3][s] = make_float4(0.f, 0.f, 0.f, 1.f);
    const uint32_t e = entity[base + i];
    a.entity[s] = e;
    a.parent[s] = kNone;
    a.parentSlot[s] = kNone;
    mark_tile(a, s);
    a.meshMat[s] = make_uint2(mesh, mat);
    const uint32_t idx = e & 0xFFFFFFu;
    if (idx < a.sparseSize) a.sparse[idx] = s + 1u;
  }
}

// ---- .scsector INST chunk -> SoA ------------------------------------------------------------------
// The raw INST payload (tools/shared/world_format.cpp:92-125 writes it, :207-281 reads it) is uploaded as it lies in
// the file; one thread per record unpacks mesh id, material id and the 9-float transform straight into the record
// planes — what WorldPartition::readSectorFile (sc_world_partition.cpp:695-732) + the World::add loop (:923-954) do on
// the CPU. Asset ids resolve through a small table (resolveMeshHandle / resolveMaterialHandle, :746-800: id 0 -> handle
// 0, unknown id -> the default asset).
struct AssetBinding  // == ScGpuAssetBinding
{
  uint32_t idLo, idHi, handle, pad;
};

__device__ __forceinline__ uint32_t resolve_asset(const AssetBinding* __restrict__ tab, uint32_t n, uint32_t dflt, uint32_t lo, uint32_t hi)
{
  if ((lo | hi) == 0u) return 0u;
  for (uint32_t k = 0; k < n; ++k)
    if (tab[k].idLo == lo && tab[k].idHi == hi) return tab[k].handle;
  return dflt;
}

__global__ void __launch_bounds__(kBlock) k_spawn_sector_file(SceneArrays a, uint32_t slot0, const uint32_t* __restrict__ slotOf, uint32_t rank0,
                                                              uint32_t n, const uint32_t* __restrict__ payload,
                                                              uint32_t recordWords, uint32_t meshWord, const uint32_t* __restrict__ entity,
                                                              const AssetBinding* __restrict__ meshes, uint32_t nMeshes, uint32_t defaultMesh,
                                                              const AssetBinding* __restrict__ materials, uint32_t nMaterials,
                                                              uint32_t defaultMaterial, uint32_t stamp)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t* r = payload + (size_t)j * recordWords + meshWord;  // mesh id (u64), material id (u64), Transform (9 floats)
  const uint32_t mesh = resolve_asset(meshes, nMeshes, defaultMesh, r[0], r[1]);
  const uint32_t mat = resolve_asset(materials, nMaterials, defaultMaterial, r[2], r[3]);
  float t[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) t[k] = __uint_as_float(r[4 + k]);
  float sx = t[6], sy = t[7], sz = t[8];
  if (sx == 0.0f && sy == 0.0f && sz == 0.0f) { sx = sy = sz = 1.0f; }  // TransformSystem's zero-scale patch, as k_spawn
  const uint32_t s = spawn_slot(slot0, slotOf, j);
  a.rank[s] = rank0 + j;
  a.perm[rank0 + j] = s;
  const uint32_t f = (kFlagBounds | kFlagMesh | kFlagLive) | (stamp << kStampShift);
  a.rec[0][s] = make_float4(t[0], t[1], t[2], t[3]);
  a.rec[1][s] = make_float4(t[4], t[5], sx, sy);
  a.rec[2][s] = make_float4(sz, -0.5f, -0.5f, -0.5f);  // kUnitCubeBounds (sc_world_partition.cpp:27, 725)
  a.rec[3][s] = make_float4(0.5f, 0.5f, 0.5f, __uint_as_float(f));
  a.world[0][s] = make_float4(1.f, 0.f, 0.f, 0.f);
  a.world[1][s] = make_float4(0.f, 1.f, 0.f, 0.f);
  a.world[2][s] = make_float4(0.f, 0.f, 1.f, 0.f);
  a.world[3][s] = make_float4(0.f, 0.f, 0.f, 1.f);
  const uint32_t e = entity[j];
  a.entity[s] = e;
  a.parent[s] = kNone;
  a.parentSlot[s] = kNone;
  mark_tile(a, s);
  a.meshMat[s] = make_uint2(mesh, mat);
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

// The setLocal family, one thread per instance. kFloats = 9: setLocal (sc_ecs.h:78-84: position, rotation, scale);
// 6: position + rotation, all that the physics sync and the traffic tiers write (sc_physics.cpp:1178-1184,
// sc_traffic_ai.cpp:449-457); 3: setLocalPosition (sc_ecs.h:92-96). Addressed by entity handle (unknown handles are
// skipped) or, with entity == nullptr, by position in the pool: element j is the Transform at dense index
// firstRank + j — no handle crosses PCIe and no sparse lookup is made.
template <int kFloats>
__global__ void __launch_bounds__(kBlock) k_set_local(SceneArrays a, uint32_t n, const uint32_t* __restrict__ entity, uint32_t firstRank,
                                                      const float* __restrict__ data, uint32_t stamp)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  uint32_t s;
  if (entity)
  {
    s = find_slot(a, entity[j]);
    if (s == kNone) return;
  }
  else
    s = a.perm[firstRank + j];
  const float* t = data + (size_t)j * kFloats;
  if (kFloats == 3)
  {
    float* r0 = reinterpret_cast<float*>(a.rec[0] + s);
    r0[0] = t[0]; r0[1] = t[1]; r0[2] = t[2];
  }
  else
  {
    a.rec[0][s] = make_float4(t[0], t[1], t[2], t[3]);
    if (kFloats == 6)
      *reinterpret_cast<float2*>(a.rec[1] + s) = make_float2(t[4], t[5]);
    else
    {
      float sx = t[6], sy = t[7], sz = t[8];
      if (sx == 0.0f && sy == 0.0f && sz == 0.0f) { sx = sy = sz = 1.0f; }  // TransformSystem's zero-scale patch (sc_ecs.cpp:143-149)
      a.rec[1][s] = make_float4(t[4], t[5], sx, sy);
      reinterpret_cast<float*>(a.rec[2] + s)[0] = sz;
    }
  }
  uint32_t* fw = reinterpret_cast<uint32_t*>(a.rec[3] + s) + 3;
  *fw = (*fw & 0xFFu) | (stamp << kStampShift);
}

// World::add / remove of RenderMesh or Bounds on an entity that already owns a Transform, and edits of their fields
// (sc_traffic_lod.cpp:47-70 adds both late and rewrites meshId / materialId on every LOD change; sc_imgui.cpp:720 edits
// materialId): mesh + material ids, local AABB and the HAS_* bits. Does not touch the dirty stamp — the reference's
// culling reads these components every frame and they have nothing to do with Transform::dirty.
__global__ void __launch_bounds__(kBlock) k_set_render(SceneArrays a, uint32_t n, const uint32_t* __restrict__ entity,
                                                       const uint32_t* __restrict__ meshMat2, const float* __restrict__ aabb6,
                                                       const uint32_t* __restrict__ flags)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t s = find_slot(a, entity[j]);
  if (s == kNone) return;
  if (meshMat2) a.meshMat[s] = make_uint2(meshMat2[(size_t)j * 2], meshMat2[(size_t)j * 2 + 1]);
  if (aabb6)
  {
    const float* b = aabb6 + (size_t)j * 6;
    float* r2 = reinterpret_cast<float*>(a.rec[2] + s);
    float* r3 = reinterpret_cast<float*>(a.rec[3] + s);
    r2[1] = b[0]; r2[2] = b[1]; r2[3] = b[2];
    r3[0] = b[3]; r3[1] = b[4]; r3[2] = b[5];
  }
  if (flags)
  {
    uint32_t* fw = reinterpret_cast<uint32_t*>(a.rec[3] + s) + 3;
    *fw = (*fw & ~(kFlagBounds | kFlagMesh)) | (flags[j] & (kFlagBounds | kFlagMesh));
  }
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

// World::destroy batches. The host replays ComponentPool::remove's swap-with-last (sc_ecs.h:240-262) on its mirror of
// the pool and hands over the net result in RANK space: `dst <- src` (the element at dense index src ends up at dst;
// src always in the vacated tail, dst below it, so sources and destinations never overlap) and the device slots of
// the destroyed Transforms. Nothing moves in HBM: a survivor only learns its new rank, a victim's slot becomes a hole
// (flags 0: not live, not a candidate, never dirty) that scgpu_layout.h hands to a later spawn.
__global__ void __launch_bounds__(kBlock) k_despawn_apply(SceneArrays a, uint32_t nMoves, const uint2* __restrict__ moves,
                                                          uint32_t nRemoved, const uint32_t* __restrict__ removedSlot)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j < nMoves)
  {
    const uint32_t dst = moves[j].x, src = moves[j].y;
    SC_ASSERT(dst < src);
    const uint32_t s = a.perm[src];
    SC_ASSERT(a.rank[s] == src && a.entity[s] != kNone);
    a.perm[dst] = s;
    a.rank[s] = dst;
  }
  else if (j < nMoves + nRemoved)
  {
    const uint32_t s = removedSlot[j - nMoves];
    const uint32_t e = a.entity[s];
    SC_ASSERT(e != kNone && a.sparse[e & 0xFFFFFFu] == s + 1u);
    a.sparse[e & 0xFFFFFFu] = 0u;
    a.entity[s] = kNone;
    a.parent[s] = kNone;
    if (a.parentSlot[s] != kNone) mark_tile(a, s);  // a link that goes away changes the window cut; a dead root does not
    a.parentSlot[s] = kNone;
    // an all-zero record: no LIVE bit, never dirty, no RenderMesh; and tame, should its window be computed as a whole
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    a.rec[0][s] = z; a.rec[1][s] = z; a.rec[2][s] = z; a.rec[3][s] = z;
  }
}

// Dirty stamps are 24 bits wide. When the update counter wraps, every stamp still stored is older than 16.7 M updates
// and would alias a future one: reset them all to 0 ("never") once, right after the update that used the last id.
__global__ void __launch_bounds__(kBlock) k_clear_stamps(SceneArrays a, uint32_t extent)
{
  const uint32_t s = blockIdx.x * kBlock + threadIdx.x;
  if (s >= extent) return;
  uint32_t* fw = reinterpret_cast<uint32_t*>(a.rec[3] + s) + 3;
  *fw &= 0xFFu;
}

// ComponentPool::denseEntities(): entity handles in pool order
__global__ void __launch_bounds__(kBlock) k_gather_dense(SceneArrays a, uint32_t n, uint32_t* __restrict__ out)
{
  const uint32_t r = blockIdx.x * kBlock + threadIdx.x;
  if (r < n) out[r] = a.entity[a.perm[r]];
}

// TransformSystem's per-frame parent validation (sc_ecs.cpp:151-164), run only when the topology changed:
// a parent is valid iff it is not the entity itself and the pool holds exactly that handle (alive && has
// Transform); otherwise the node becomes a root and, if it had a parent handle, dirty.
__global__ void __launch_bounds__(kBlock) k_resolve_parents(SceneArrays a, uint32_t count, uint32_t stamp)
{
  // four consecutive slots per thread: two 128-bit loads, then four independent neighbour look-ups in flight at once
  // (the kernel is bound by the latency of those gathers, not by bytes); count is padded to a multiple of four slots
  const uint32_t s0 = (blockIdx.x * kBlock + threadIdx.x) * 4u;
  if (s0 >= count) return;
  const uint4 ph4 = reinterpret_cast<const uint4*>(a.parent)[s0 >> 2];
  const uint4 old4 = reinterpret_cast<const uint4*>(a.parentSlot)[s0 >> 2];
  const uint32_t ph_[4] = { ph4.x, ph4.y, ph4.z, ph4.w }, old_[4] = { old4.x, old4.y, old4.z, old4.w };
  uint32_t held[4];
#pragma unroll
  for (uint32_t k = 0; k < 4u; ++k)
  {
    // A link resolved earlier stays valid for as long as that slot holds that very handle (index AND generation: the
    // parent is alive and owns its Transform) - a look at a neighbouring slot instead of the random walk through the
    // sparse table; a slot never moves. Only new or broken links go the long way.
    const bool look = s0 + k < count && ph_[k] != kNone && old_[k] != kNone && old_[k] != s0 + k && old_[k] < count;
    held[k] = look ? a.entity[old_[k]] : kNone;
  }
#pragma unroll
  for (uint32_t k = 0; k < 4u; ++k)
  {
    const uint32_t s = s0 + k, ph = ph_[k];
    if (s >= count) break;
    uint32_t ps = kNone;
    if (ph != kNone)
    {
      if (held[k] == ph) continue;  // (ph != kNone, so a skipped look-up never matches)
      if (ph != a.entity[s]) ps = find_slot(a, ph);
      if (ps == kNone)
      {
        a.parent[s] = kNone;
        uint32_t* fw = reinterpret_cast<uint32_t*>(a.rec[3] + s) + 3;
        *fw = (*fw & 0xFFu) | (stamp << kStampShift);
      }
    }
    if (old_[k] != ps)
    {
      a.parentSlot[s] = ps;
      mark_tile(a, s);
    }
  }
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
