// Host build of the device math header, exported for ctypes (tests only).
#include "cuda_runtime.h"
#include "../../sc-gameengine_b200/csrc/scgpu_math.cuh"
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
}
