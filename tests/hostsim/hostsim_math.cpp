// Host build of the device math header, exported for ctypes (tests only).
#include "cuda_runtime.h"
#include "../../sc-gameengine_b200/csrc/scgpu_math.cuh"
#include "../../sc-gameengine_b200/csrc/scgpu_traffic.cuh"
#include <vector>
using namespace scgpu;
extern "C" {
void hs_sincos(float y, float* s, float* c) { sincosf_glibc(y, *s, *c); }
void hs_trs(const float* t, float* out16)
{
  Mat4 m = mat4_trs_dense(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]);
  std::memcpy(out16, &m, 64);
}
int hs_trs_fast(const float* t, float* out16)
{
  bool affine = false;
  Mat4 m = mat4_trs(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], affine);
  std::memcpy(out16, &m, 64);
  return affine ? 1 : 0;
}
void hs_compose(const float* a, const float* b, int localAffine, float* out16)
{
  Mat4 A, B; std::memcpy(&A, a, 64); std::memcpy(&B, b, 64);
  Mat4 m = compose(A, B, localAffine != 0);
  std::memcpy(out16, &m, 64);
}
void hs_mul(const float* a, const float* b, float* out16)
{
  Mat4 A, B; std::memcpy(&A, a, 64); std::memcpy(&B, b, 64);
  Mat4 m = mat4_mul(A, B);
  std::memcpy(out16, &m, 64);
}
void hs_sphere(const float* w16, const float* bb, float* out4)
{
  Mat4 W; std::memcpy(&W, w16, 64);
  world_bounds_sphere(W, bb[0], bb[1], bb[2], bb[3], bb[4], bb[5], out4[0], out4[1], out4[2], out4[3]);
}
int hs_in_frustum(const float* planes24, const float* c, float r)
{
  return sphere_in_frustum((const float4*)planes24, c[0], c[1], c[2], r) ? 1 : 0;
}
// The favourite-plane PRE-TEST of k_update_win (sphere_cull_warp_fav, csrc/scgpu_kernels.cuh) restated with the same
// operations - fmaf is the per-lane semantics of fma.rn.f32x2 - on top of the REAL header functions for the sphere
// centre, the radius bound and the exact radius. Fuzzes the claim the kernel relies on: whenever the fused chain with
// its slack says "certainly culled by this plane", the reference's own predicate (every operation rounded, exact
// radius) culls too. cityLike != 0 keeps the magnitudes to those of a city scene. out[0] = cases, out[1] = "certain" verdicts, out[2] = cases the reference culls by that plane,
// out[3] = VIOLATIONS (must be 0), out[4] = culled by the old pre-test (rounded distance < -bound) but not certain now.
void hs_fav_pretest_fuzz(uint64_t seed, uint64_t n, int cityLike, uint64_t* out)
{
  uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
  auto next = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  auto uni = [&]() { return (float)((next() >> 40) * (1.0 / 16777216.0)); };                  // [0, 1)
  auto sym = [&]() { return uni() * 2.0f - 1.0f; };
  auto mag = [&](float lo, float hi) { return std::pow(10.0f, lo + (hi - lo) * uni()); };  // log-uniform
  for (int k = 0; k < 5; ++k) out[k] = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    // an affine world matrix (rotation-ish 3x3 with scales over four decades, translation up to 1e5 m), an AABB
    Mat4 W;
    // cityLike: scales 0.1 .. 10, positions up to 10 km, half extents 0.1 .. 5 m; else four decades more of everything
    const float sc = cityLike ? mag(-1.f, 1.f) : mag(-2.f, 2.f), tr = cityLike ? mag(0.f, 4.f) : mag(-1.f, 5.f);
    W.c0 = make_float4(sym() * sc, sym() * sc, sym() * sc, 0.f);
    W.c1 = make_float4(sym() * sc, sym() * sc, sym() * sc, 0.f);
    W.c2 = make_float4(sym() * sc, sym() * sc, sym() * sc, 0.f);
    W.c3 = make_float4(sym() * tr, sym() * tr, sym() * tr, 1.f);
    const float he = cityLike ? mag(-1.f, 0.7f) : mag(-2.f, 1.f);
    const float cxl = sym(), cyl = sym(), czl = sym();
    float ox, oy, oz, ex, ey, ez;
    world_bounds_centre(W, cxl - he * uni(), cyl - he * uni(), czl - he * uni(), cxl + he * uni(), cyl + he * uni(), czl + he * uni(),
                        ox, oy, oz, ex, ey, ez);
    const float bound = world_bounds_radius_bound(W, ex, ey, ez), radius = world_bounds_radius(W, ex, ey, ez);
    // a plane: normalised like frustumFromViewProj does it (a * invLen), or left tiny / unnormalised now and then
    float a = sym(), b = sym(), c = sym();
    const float lenSq = a * a + b * b + c * c;
    const uint64_t kind = next() % 16;
    if (kind == 0) { a *= 1e-5f; b *= 1e-5f; c *= 1e-5f; }
    else if (kind == 1) { a *= 3.f; b *= 3.f; c *= 3.f; }  // scgpuSetViewPlanes takes whatever it is given
    else if (lenSq > 1e-8f) { const float inv = 1.0f / std::sqrt(lenSq); a *= inv; b *= inv; c *= inv; }
    // d so that the sphere sits at a chosen signed distance in units of its radius bound: mostly around the threshold
    const float want = (kind < 12 ? -(0.9f + 0.3f * uni()) : -mag(0.f, 3.f) * (next() & 1 ? 1.f : -0.1f)) * bound;
    const float d = want - (a * ox + b * oy + c * oz);
    const float K = std::fmax(std::fabs(a), std::fmax(std::fabs(b), std::fabs(c)));  // refreshPlaneSlack, one plane
    const float slackK = K * 0x1p-19f;
    // reference predicate: sphereInFrustum's distance, every operation rounded (plane_dist)
    const float D = (((a * ox) + (b * oy)) + (c * oz)) + d;
    const bool refCulls = D < -radius;
    // the kernel's pre-test
    const float negBound = -bound;
    const float nb = fmaf(-slackK, std::fabs(ox) + std::fabs(oy) + std::fabs(oz), negBound);
    // the device starts the chain from __fmaf_ru(|d|, 2^-19, d); the round-to-nearest value used here is never larger,
    // the chain is monotone in its start value, so a "certain" here is a superset of the device's: the weaker claim is checked
    const float dStart = fmaf(std::fabs(d), 0x1p-19f, d);
    float t = fmaf(c, oz, dStart); t = fmaf(b, oy, t); t = fmaf(a, ox, t);
    const bool certain = t < nb;
    out[0] += 1;
    out[1] += certain ? 1 : 0;
    out[2] += refCulls ? 1 : 0;
    out[3] += (certain && !refCulls) ? 1 : 0;
    out[4] += (D < negBound && !certain) ? 1 : 0;
  }
}
// exhaustive-capable sweep: counts bit mismatches of device sincos vs reference fns over [first, first+count)
uint64_t hs_sincos_sweep(uint32_t first, uint64_t count, uint32_t stride, float (*rs)(float), float (*rc)(float))
{
  uint64_t bad = 0;
  uint32_t bits = first;
  for (uint64_t i = 0; i < count; ++i, bits += stride)
  {
    float x = __uint_as_float(bits), s, c;
    sincosf_glibc(x, s, c);
    const float es = rs(x), ec = rc(x);
    const bool okS = (__float_as_uint(s) == __float_as_uint(es)) || (s != s && es != es);
    const bool okC = (__float_as_uint(c) == __float_as_uint(ec)) || (c != c && ec != ec);
    bad += (okS ? 0 : 1) + (okC ? 0 : 1);
  }
  return bad;
}

// ---- traffic on rails (scgpu_traffic.cuh) -------------------------------------------------------------------
float hs_expf(float x) { return expf_glibc(x); }
float hs_atanf(float x) { return atanf_glibc(x); }
float hs_atan2f(float y, float x) { return atan2f_glibc(y, x); }
static bool hs_same(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }
uint64_t hs_unary_sweep(int which, uint32_t first, uint64_t count, uint32_t stride, float (*ref)(float))
{
  uint64_t bad = 0;
  uint32_t bits = first;
  for (uint64_t i = 0; i < count; ++i, bits += stride)
  {
    const float x = __uint_as_float(bits);
    bad += hs_same(which == 0 ? expf_glibc(x) : atanf_glibc(x), ref(x)) ? 0 : 1;
  }
  return bad;
}
uint64_t hs_atan2_pairs(uint64_t n, const float* y, const float* x, float (*ref)(float, float))
{
  uint64_t bad = 0;
  for (uint64_t i = 0; i < n; ++i) bad += hs_same(atan2f_glibc(y[i], x[i]), ref(y[i], x[i])) ? 0 : 1;
  return bad;
}
// the per-agent device routine over the flat arrays of ScGpuLaneGraph, packed like scgpuTrafficSetLanes packs them
void hs_traffic_on_rails(uint32_t nNodes, uint32_t nSegs, const float* nodePos3, const float* nodeSpeed, const uint32_t* connOffset,
                         const uint32_t* conn, const uint32_t* segNodes2, const float* segDir3, const float* segLen,
                         const uint8_t* segActive, float defaultSpeed, uint32_t n, uint32_t* laneId, float* laneS,
                         float* targetSpeed, float* look, float* trs9, const float* brake, const uint8_t* skip, float dt,
                         int hasDebug, float dbgLook, float dbgMul, uint8_t* outMoved)
{
  std::vector<float4> np(nNodes), sd(nSegs);
  std::vector<uint2> nc(nNodes);
  std::vector<uint4> sn(nSegs);
  for (uint32_t i = 0; i < nNodes; ++i)
  {
    np[i] = make_float4(nodePos3[3 * i], nodePos3[3 * i + 1], nodePos3[3 * i + 2], nodeSpeed[i]);
    nc[i] = make_uint2(connOffset[i], connOffset[i + 1] - connOffset[i]);
  }
  for (uint32_t i = 0; i < nSegs; ++i)
  {
    sd[i] = make_float4(segDir3[3 * i], segDir3[3 * i + 1], segDir3[3 * i + 2], segLen[i]);
    sn[i] = make_uint4(segNodes2[2 * i], segNodes2[2 * i + 1], segActive[i] ? 1u : 0u, 0u);
  }
  LaneGraphView g{ np.data(), nc.data(), conn, sd.data(), sn.data(), nNodes, nSegs, defaultSpeed };
  TrafficStepParams st{ dt, dbgMul, dbgLook, hasDebug ? 1u : 0u };
  for (uint32_t i = 0; i < n; ++i)
  {
    outMoved[i] = 0;
    if (skip && skip[i]) continue;
    float* t = trs9 + (size_t)i * 9;
    float pos[3] = { t[0], t[1], t[2] };
    float yaw = 0.0f;
    if (traffic_agent_on_rails(g, st, brake ? brake[i] : 0.0f, laneId[i], laneS[i], targetSpeed[i], look[i], pos, yaw))
    {
      t[0] = pos[0]; t[1] = pos[1]; t[2] = pos[2];
      t[3] = 0.0f; t[4] = yaw; t[5] = 0.0f;
      outMoved[i] = 1;
    }
  }
}
}
