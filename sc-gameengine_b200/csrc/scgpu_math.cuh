// scgpu_math.cuh — device arithmetic of the scene-update hot path, bit-matched to the reference's CPU code.
//
// Every parity-critical operation is written with the explicit round-to-nearest intrinsics
// (__fmul_rn/__fadd_rn/__dmul_rn/__dadd_rn ...): nvcc never contracts those into FMA, so the results do not
// depend on -fmad. Denormals are kept (no -ftz), sqrt and division are the IEEE variants.
//
// Reference (relative to /root/reference):
//   mat4_mul           src/core/src/sc_math.cpp:52-68   ((a0*b0 + a1*b1) + a2*b2) + a3*b3, separate mul/add
//   mat4_rotation_xyz  src/core/src/sc_math.cpp:100-128 R = (Rz*Ry)*Rx, dense products
//   mat4_trs           src/core/src/sc_math.cpp:130-142 M = T*(R*S), dense products
//   computeWorldBoundsSphere / sphereInFrustum  src/engine/world/sc_world_partition.cpp:1105-1144
//   std::sin/std::cos(float) -> glibc 2.39 sinf/cosf generic variant (see oracle/scoracle.c header)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scgpu
{

struct Mat4
{
  float4 c0, c1, c2, c3;  // columns; element (row r, col c) = c<c>.<xyzw[r]>
};

// ---------------------------------------------------------------------------------------------------
// glibc sinf/cosf (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h) in IEEE double, no contraction
// ---------------------------------------------------------------------------------------------------

__constant__ uint32_t kInvPio4[24] = {
  0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
  0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
  0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041
};

// s_sincosf_data.c coefficients, kept in constant memory so that the FP64 instructions take them as constant-bank
// operands instead of re-materialising 64-bit immediates through uniform registers at every use
__constant__ double kSinCosCoef[11] = {
  -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13,                                  // S1 S2 S3
  0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16,    // C0..C4
  0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1.921FB54442D18p-62                                     // 2/pi*2^24, pi/2, pi*2^-63
};

// sine polynomial of sincosf.h:sinf_poly (n even): x + x^3*s1 + x^7*(s2 + x^2*s3)
__device__ __forceinline__ float sin_poly(double x, double x2)
{
  const double S1 = kSinCosCoef[0], S2 = kSinCosCoef[1], S3 = kSinCosCoef[2];
  const double x3 = __dmul_rn(x, x2);
  const double s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
  const double x7 = __dmul_rn(x3, x2);
  const double s = __dadd_rn(x, __dmul_rn(x3, S1));
  return __double2float_rn(__dadd_rn(s, __dmul_rn(x7, s1)));
}

// cosine polynomial of sincosf.h:sinf_poly (n odd), table 0 coefficients. Table 1 holds the negated
// coefficients; round-to-nearest is sign-symmetric, so its result is exactly the negation of this one.
__device__ __forceinline__ float cos_poly(double x2)
{
  const double C0 = kSinCosCoef[3], C1 = kSinCosCoef[4], C2 = kSinCosCoef[5], C3 = kSinCosCoef[6], C4 = kSinCosCoef[7];
  const double x4 = __dmul_rn(x2, x2);
  const double c2 = __dadd_rn(C3, __dmul_rn(x2, C4));
  const double c1 = __dadd_rn(C0, __dmul_rn(x2, C1));
  const double x6 = __dmul_rn(x4, x2);
  const double c = __dadd_rn(c1, __dmul_rn(x4, C2));
  return __double2float_rn(__dadd_rn(c, __dmul_rn(x6, c2)));
}

// sincosf.h:reduce_large — |x| >= 120, 32x96 -> 128 bit fixed-point product with 4/pi
__device__ __forceinline__ double reduce_large(uint32_t xi, int* np)
{
  const uint32_t* arr = &kInvPio4[(xi >> 26) & 15];
  const int shift = (xi >> 23) & 7;
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  res0 = (uint32_t)(xi * arr[0]);
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  const double x = __ll2double_rn((long long)res0);
  *np = (int)n;
  return __dmul_rn(x, kSinCosCoef[10]);
}

// abstop12-style classes of sincosf.h: pio4f = 0x3f490fdb, 2^-12 = 0x39800000, 120.0f = 0x42f00000, inf = 0x7f800000
constexpr uint32_t kTopTiny = 0x398u, kTop120 = 0x42fu, kTopInf = 0x7f8u;

// |y| < 2^-12 (zero and denormals included): sinf returns y and cosf returns 1.0f without any arithmetic
__device__ __forceinline__ bool sincos_is_trivial(float y) { return ((__float_as_uint(y) >> 20) & 0x7ffu) < kTopTiny; }

// sinf(y) and cosf(y) of glibc 2.39 (generic variant) for |y| >= 2^-12, sharing one range reduction; bit-identical
// to the two separate libm calls of sc_math.cpp:102-107.
// glibc evaluates |y| < pi/4 without a reduction. Here that class goes through reduce_fast as well: for |y| < 0.75
// it yields n = 0 exactly (|y * 2/pi * 2^24| < 2^23) and x - 0*pi/2 = x exactly, after which both routes evaluate
// the same two polynomials with sign +1 and table 0. One route instead of two keeps a warp whose angles straddle
// pi/4 from evaluating the polynomials twice.
__device__ __forceinline__ void sincosf_glibc_nt(float y, float& sn, float& cs)
{
  const uint32_t yi = __float_as_uint(y);
  const uint32_t top = (yi >> 20) & 0x7ffu;
  double x = (double)y;
  int n;
  int q;  // quadrant used for the sign / table selection
  if (top < kTop120)
  {
    // sincosf.h:reduce_fast, !TOINT_INTRINSICS: hpi_inv prescaled by 2^24
    const double r = __dmul_rn(x, kSinCosCoef[8]);
    n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __dsub_rn(x, __dmul_rn((double)n, kSinCosCoef[9]));
    q = n;
  }
  else if (top < kTopInf)
  {
    x = reduce_large(yi, &n);
    q = n + (int)(yi >> 31);
  }
  else
  {
    // __math_invalidf: (y - y) / (y - y)
    const float d = __fsub_rn(y, y);
    sn = cs = __fdiv_rn(d, d);
    return;
  }

  // sign[q & 3] = {1, -1, -1, 1}; multiplication by +-1.0 is exact
  const double xs = (((q + 1) & 2) != 0) ? -x : x;
  const double x2 = __dmul_rn(x, x);
  const float S = sin_poly(xs, x2);
  float C = cos_poly(x2);
  if (q & 2) C = -C;  // table 1
  if (n & 1) { sn = C; cs = S; }
  else       { sn = S; cs = C; }
}

__device__ __forceinline__ void sincosf_glibc(float y, float& sn, float& cs)
{
  if (sincos_is_trivial(y))
  {
    sn = y;
    cs = 1.0f;
    return;
  }
  sincosf_glibc_nt(y, sn, cs);
}

// Two IEEE single-precision products in ONE instruction (Blackwell FMUL2, PTX mul.rn.f32x2): each lane is rounded to
// nearest like a scalar mul.rn.f32, so results are bit-identical; the kernels are issue-bound, so halving the
// multiply count matters. Only multiplies are paired: ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even with
// -fmad=false and explicit .rn, which would break parity, whereas the scalar add.rn.f32 that consumes these products
// is never contracted.
__device__ __forceinline__ float2 fmul2_rn(float ax, float ay, float b)
{
#if defined(__CUDA_ARCH__)
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
  asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(b));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
#else
  return make_float2(__fmul_rn(ax, b), __fmul_rn(ay, b));
#endif
}
// Two IEEE single-precision SUMS in one instruction: fma.rn.f32x2(a, 1.0, b). a * 1.0 is exact, so the single
// rounding of the fused operation is the rounding of a + b: bit-identical to add.rn.f32 per lane (signed zeros, NaN,
// Inf and denormals included). The pair of ones comes from constant memory on purpose: with a literal 1.0 ptxas
// simplifies the fma to an add and then contracts it with the mul.f32x2 that produced its operand (a real FMA,
// which breaks parity); a value it cannot see through leaves 3 x FMUL2 + 2 x FFMA2(UR) for a 3-term dot-product
// pair. The ones live in a uniform register pair, no general register is spent.
__constant__ unsigned long long kOnes2 = 0x3f8000003f800000ull;
__device__ __forceinline__ float2 fadd2_rn(float2 a, float2 b)
{
#if defined(__CUDA_ARCH__)
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(kOnes2), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
#else
  return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
#endif
}
// (ax*bx, ay*by)
__device__ __forceinline__ float2 fmul2_rn(float ax, float ay, float bx, float by)
{
#if defined(__CUDA_ARCH__)
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(bx), "f"(by));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
#else
  return make_float2(__fmul_rn(ax, bx), __fmul_rn(ay, by));
#endif
}

// ---------------------------------------------------------------------------------------------------
// dense 4x4 products, reference operation order
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ float dot4_ref(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3)
{
  float r = __fmul_rn(a0, b0);
  r = __fadd_rn(r, __fmul_rn(a1, b1));
  r = __fadd_rn(r, __fmul_rn(a2, b2));
  r = __fadd_rn(r, __fmul_rn(a3, b3));
  return r;
}

__device__ __forceinline__ float4 mul_col(const Mat4& a, float4 b)
{
  float4 r;
  r.x = dot4_ref(a.c0.x, a.c1.x, a.c2.x, a.c3.x, b.x, b.y, b.z, b.w);
  r.y = dot4_ref(a.c0.y, a.c1.y, a.c2.y, a.c3.y, b.x, b.y, b.z, b.w);
  r.z = dot4_ref(a.c0.z, a.c1.z, a.c2.z, a.c3.z, b.x, b.y, b.z, b.w);
  r.w = dot4_ref(a.c0.w, a.c1.w, a.c2.w, a.c3.w, b.x, b.y, b.z, b.w);
  return r;
}

// mat4_mul, sc_math.cpp:52-68
__device__ __forceinline__ Mat4 mat4_mul(const Mat4& a, const Mat4& b)
{
  Mat4 r;
  r.c0 = mul_col(a, b.c0);
  r.c1 = mul_col(a, b.c1);
  r.c2 = mul_col(a, b.c2);
  r.c3 = mul_col(a, b.c3);
  return r;
}

__device__ __forceinline__ Mat4 mat4_identity()
{
  Mat4 m;
  m.c0 = make_float4(1.f, 0.f, 0.f, 0.f);
  m.c1 = make_float4(0.f, 1.f, 0.f, 0.f);
  m.c2 = make_float4(0.f, 0.f, 1.f, 0.f);
  m.c3 = make_float4(0.f, 0.f, 0.f, 1.f);
  return m;
}

// mat4_trs, sc_math.cpp:130-142, all four products dense (zeros and ones take part), so that signed zeros,
// NaN and Inf propagate exactly as on the CPU.
__device__ __forceinline__ Mat4 mat4_trs_dense(float px, float py, float pz, float rx, float ry, float rz, float sx_,
                                               float sy_, float sz_)
{
  float sx, cx, sy, cy, sz, cz;
  sincosf_glibc(rx, sx, cx);
  sincosf_glibc(ry, sy, cy);
  sincosf_glibc(rz, sz, cz);

  Mat4 rxm = mat4_identity();
  rxm.c1.y = cx; rxm.c1.z = sx; rxm.c2.y = -sx; rxm.c2.z = cx;
  Mat4 rym = mat4_identity();
  rym.c0.x = cy; rym.c0.z = -sy; rym.c2.x = sy; rym.c2.z = cy;
  Mat4 rzm = mat4_identity();
  rzm.c0.x = cz; rzm.c0.y = sz; rzm.c1.x = -sz; rzm.c1.y = cz;

  const Mat4 r = mat4_mul(mat4_mul(rzm, rym), rxm);

  Mat4 s;
  s.c0 = make_float4(sx_, 0.f, 0.f, 0.f);
  s.c1 = make_float4(0.f, sy_, 0.f, 0.f);
  s.c2 = make_float4(0.f, 0.f, sz_, 0.f);
  s.c3 = make_float4(0.f, 0.f, 0.f, 1.f);

  Mat4 t = mat4_identity();
  t.c3 = make_float4(px, py, pz, 1.f);

  return mat4_mul(t, mat4_mul(r, s));
}

// ---------------------------------------------------------------------------------------------------
// structured products: same values as the dense reference products, without the zero/one terms
//
// For FINITE operands every skipped term is x*(+-0) = +-0 or x*1 = x, and adding +-0 to a sum never changes
// its value; the kept terms are evaluated in the reference's order with the same roundings. The results are
// therefore equal to the dense ones as real numbers - the only representable difference is the SIGN OF A ZERO
// entry, which no consumer can observe (every later use multiplies, adds or compares with <). Non-finite
// operands (where 0*inf = NaN would contaminate the dense result) take the dense path.
// ---------------------------------------------------------------------------------------------------

// true when mat4_trs_fast() is valid: all nine inputs finite and small enough that rotation*scale cannot
// overflow (|R| <= 2, so |scale| < 2^120 keeps |R*s| finite). One comparison on the sum of magnitudes.
__device__ __forceinline__ bool trs_inputs_tame(float px, float py, float pz, float rx, float ry, float rz, float sx,
                                                float sy, float sz)
{
  const float sum = fabsf(px) + fabsf(py) + fabsf(pz) + fabsf(rx) + fabsf(ry) + fabsf(rz) + fabsf(sx) + fabsf(sy) +
                    fabsf(sz);
  return sum < 0x1p120f;  // false for NaN
}

// mat4_trs for tame inputs, given the six sines / cosines: R = (Rz*Ry)*Rx, M = T*(R*S) with the structural zeros
// and ones removed.
__device__ __forceinline__ Mat4 mat4_trs_from_sincos(float px, float py, float pz, float sx, float cx, float sy, float cy,
                                                     float sz, float cz, float sx_, float sy_, float sz_)
{
  // A = Rz*Ry  (products in pairs, FMUL2; every sum stays a scalar add in the reference's order)
  const float2 a0 = fmul2_rn(cz, sz, cy);          // a00, a10
  const float a20 = -sy;
  const float a01 = -sz, a11 = cz;                 // a21 = 0
  const float2 a2 = fmul2_rn(cz, sz, sy);          // a02, a12
  const float a22 = cy;
  // R = A*Rx : col0 = A.col0 ; col1 = A.col1*cx + A.col2*sx ; col2 = A.col1*(-sx) + A.col2*cx
  const float nsx = -sx;
  const float2 p1 = fmul2_rn(a01, a11, cx), q1 = fmul2_rn(a2.x, a2.y, sx);
  const float2 r1 = fadd2_rn(p1, q1);
  const float r01 = r1.x, r11 = r1.y;
  const float2 p2 = fmul2_rn(a01, a11, nsx), q2 = fmul2_rn(a2.x, a2.y, cx);
  const float2 r2_ = fadd2_rn(p2, q2);
  const float r02 = r2_.x, r12 = r2_.y;
  const float2 r2 = fmul2_rn(sx, cx, a22);         // r21, r22
  Mat4 m;
  const float2 c0 = fmul2_rn(a0.x, a0.y, sx_), c1 = fmul2_rn(r01, r11, sy_), c2 = fmul2_rn(r02, r12, sz_);
  m.c0 = make_float4(c0.x, c0.y, __fmul_rn(a20, sx_), 0.f);
  m.c1 = make_float4(c1.x, c1.y, __fmul_rn(r2.x, sy_), 0.f);
  m.c2 = make_float4(c2.x, c2.y, __fmul_rn(r2.y, sz_), 0.f);
  m.c3 = make_float4(px, py, pz, 1.f);
  return m;
}

__device__ __forceinline__ Mat4 mat4_trs_fast(float px, float py, float pz, float rx, float ry, float rz, float sx_,
                                              float sy_, float sz_)
{
  float sx, cx, sy, cy, sz, cz;
  sincosf_glibc(rx, sx, cx);
  sincosf_glibc(ry, sy, cy);
  sincosf_glibc(rz, sz, cz);
  return mat4_trs_from_sincos(px, py, pz, sx, cx, sy, cy, sz, cz, sx_, sy_, sz_);
}

__device__ __noinline__ Mat4 mat4_trs_dense_call(float px, float py, float pz, float rx, float ry, float rz, float sx,
                                                 float sy, float sz)
{
  return mat4_trs_dense(px, py, pz, rx, ry, rz, sx, sy, sz);
}

// mat4_trs with the fast path where it is exact; `affine` reports that the bottom row is exactly (0,0,0,1)
__device__ __forceinline__ Mat4 mat4_trs(float px, float py, float pz, float rx, float ry, float rz, float sx, float sy,
                                         float sz, bool& affine)
{
  affine = trs_inputs_tame(px, py, pz, rx, ry, rz, sx, sy, sz);
  if (affine) return mat4_trs_fast(px, py, pz, rx, ry, rz, sx, sy, sz);
  return mat4_trs_dense_call(px, py, pz, rx, ry, rz, sx, sy, sz);
}

__device__ __noinline__ Mat4 mat4_mul_dense_call(const Mat4& a, const Mat4& b) { return mat4_mul(a, b); }

// parent.world * local where local's bottom row is exactly (0,0,0,1): the term p[r][3]*0 is dropped for the
// first three columns and p[r][3]*1 = p[r][3] for the last. Requires the parent's last column to be finite.
__device__ __forceinline__ Mat4 mat4_mul_affine(const Mat4& p, const Mat4& l)
{
  Mat4 r;
#define SC_COL3(dst, L)                                                                                      \
  dst.x = __fadd_rn(__fadd_rn(__fmul_rn(p.c0.x, L.x), __fmul_rn(p.c1.x, L.y)), __fmul_rn(p.c2.x, L.z));     \
  dst.y = __fadd_rn(__fadd_rn(__fmul_rn(p.c0.y, L.x), __fmul_rn(p.c1.y, L.y)), __fmul_rn(p.c2.y, L.z));     \
  dst.z = __fadd_rn(__fadd_rn(__fmul_rn(p.c0.z, L.x), __fmul_rn(p.c1.z, L.y)), __fmul_rn(p.c2.z, L.z));     \
  dst.w = __fadd_rn(__fadd_rn(__fmul_rn(p.c0.w, L.x), __fmul_rn(p.c1.w, L.y)), __fmul_rn(p.c2.w, L.z));
  SC_COL3(r.c0, l.c0)
  SC_COL3(r.c1, l.c1)
  SC_COL3(r.c2, l.c2)
  SC_COL3(r.c3, l.c3)
#undef SC_COL3
  r.c3.x = __fadd_rn(r.c3.x, p.c3.x);
  r.c3.y = __fadd_rn(r.c3.y, p.c3.y);
  r.c3.z = __fadd_rn(r.c3.z, p.c3.z);
  r.c3.w = __fadd_rn(r.c3.w, p.c3.w);
  return r;
}

// the same with an affine parent too (bottom row (0,0,0,1)): the bottom row of the product is (0,0,0,1) again
__device__ __forceinline__ Mat4 mat4_mul_affine3(const Mat4& p, const Mat4& l)
{
  Mat4 r;
#define SC_COL3R(dst, L)                                                                                   \
  dst.x = __fadd_rn(__fadd_rn(__fmul_rn(p.c0.x, L.x), __fmul_rn(p.c1.x, L.y)), __fmul_rn(p.c2.x, L.z));     \
  dst.y = __fadd_rn(__fadd_rn(__fmul_rn(p.c0.y, L.x), __fmul_rn(p.c1.y, L.y)), __fmul_rn(p.c2.y, L.z));     \
  dst.z = __fadd_rn(__fadd_rn(__fmul_rn(p.c0.z, L.x), __fmul_rn(p.c1.z, L.y)), __fmul_rn(p.c2.z, L.z));     \
  dst.w = 0.0f;
  SC_COL3R(r.c0, l.c0)
  SC_COL3R(r.c1, l.c1)
  SC_COL3R(r.c2, l.c2)
  SC_COL3R(r.c3, l.c3)
#undef SC_COL3R
  r.c3.x = __fadd_rn(r.c3.x, p.c3.x);
  r.c3.y = __fadd_rn(r.c3.y, p.c3.y);
  r.c3.z = __fadd_rn(r.c3.z, p.c3.z);
  r.c3.w = 1.0f;
  return r;
}

__device__ __forceinline__ bool mat4_is_affine(const Mat4& m)
{
  return m.c0.w == 0.0f && m.c1.w == 0.0f && m.c2.w == 0.0f && m.c3.w == 1.0f;
}

// world = parent.world * local (sc_ecs.cpp:191-195)
__device__ __forceinline__ Mat4 compose(const Mat4& parentWorld, const Mat4& local, bool localAffine)
{
  const float mag = fabsf(parentWorld.c3.x) + fabsf(parentWorld.c3.y) + fabsf(parentWorld.c3.z) + fabsf(parentWorld.c3.w);
  if (localAffine && mag < __int_as_float(0x7f800000)) return mat4_mul_affine(parentWorld, local);
  return mat4_mul_dense_call(parentWorld, local);
}

// ---------------------------------------------------------------------------------------------------
// bounds + plane tests, sc_world_partition.cpp:1105-1144
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ float sum3_ref(float a, float b, float c)
{
  return __fadd_rn(__fadd_rn(a, b), c);
}

__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }  // std::max

// computeWorldBoundsSphere (sc_world_partition.cpp:1119-1144) in two halves: centre = M * aabb-centre (affine,
// left-to-right sum) + the AABB half extents, and radius = |extent| * max column norm.
__device__ __forceinline__ void world_bounds_centre(const Mat4& m, float bminx, float bminy, float bminz, float bmaxx,
                                                    float bmaxy, float bmaxz, float& ox, float& oy, float& oz, float& ex,
                                                    float& ey, float& ez)
{
  const float cx = __fmul_rn(__fadd_rn(bminx, bmaxx), 0.5f);
  const float cy = __fmul_rn(__fadd_rn(bminy, bmaxy), 0.5f);
  const float cz = __fmul_rn(__fadd_rn(bminz, bmaxz), 0.5f);
  ex = __fmul_rn(__fsub_rn(bmaxx, bminx), 0.5f);
  ey = __fmul_rn(__fsub_rn(bmaxy, bminy), 0.5f);
  ez = __fmul_rn(__fsub_rn(bmaxz, bminz), 0.5f);
  // products in pairs (x, y) per column; the sums keep the reference's order and stay scalar
  const float2 a0 = fmul2_rn(m.c0.x, m.c0.y, cx), a1 = fmul2_rn(m.c1.x, m.c1.y, cy), a2 = fmul2_rn(m.c2.x, m.c2.y, cz);
  const float2 oxy = fadd2_rn(fadd2_rn(fadd2_rn(a0, a1), a2), make_float2(m.c3.x, m.c3.y));  // sum3_ref per lane, + c3
  ox = oxy.x;
  oy = oxy.y;
  oz = __fadd_rn(sum3_ref(__fmul_rn(m.c0.z, cx), __fmul_rn(m.c1.z, cy), __fmul_rn(m.c2.z, cz)), m.c3.z);
}

__device__ __forceinline__ float world_bounds_radius(const Mat4& m, float ex, float ey, float ez)
{
  // std::max(sx, std::max(sy, sz)) of the three column norms. sqrt is monotonic and correctly rounded, so the
  // maximum of the square roots is the square root of the maximum of the squares (same comparison structure:
  // NaN operands and ties select the same side in both domains): one sqrt instead of three.
  const float2 s0 = fmul2_rn(m.c0.x, m.c0.y, m.c0.x, m.c0.y), s1 = fmul2_rn(m.c1.x, m.c1.y, m.c1.x, m.c1.y);
  const float2 s2 = fmul2_rn(m.c2.x, m.c2.y, m.c2.x, m.c2.y);
  const float qx = sum3_ref(s0.x, s0.y, __fmul_rn(m.c0.z, m.c0.z));
  const float qy = sum3_ref(s1.x, s1.y, __fmul_rn(m.c1.z, m.c1.z));
  const float qz = sum3_ref(s2.x, s2.y, __fmul_rn(m.c2.z, m.c2.z));
  const float maxScale = __fsqrt_rn(std_max(qx, std_max(qy, qz)));
  const float localRadius = __fsqrt_rn(sum3_ref(__fmul_rn(ex, ex), __fmul_rn(ey, ey), __fmul_rn(ez, ez)));
  return __fmul_rn(localRadius, maxScale);
}

// A cheap UPPER bound of world_bounds_radius(): sqrt(a^2+b^2+c^2) <= |a|+|b|+|c| for the extent, and every column
// norm <= the sum of the magnitudes of all nine entries. The factor and the absolute term swallow every rounding error
// on either side (about twenty roundings of 2^-24 each, denormal results included), so for finite inputs
// bound >= radius holds for the VALUES THE REFERENCE COMPUTES. NaN, Inf and magnitudes above 1e18 give +Inf, which
// makes every "d < -bound" test false. Used only to prove "culled" early: d < -bound implies d < -radius.
__device__ __forceinline__ float world_bounds_radius_bound(const Mat4& m, float ex, float ey, float ez)
{
  const float e = fabsf(ex) + fabsf(ey) + fabsf(ez);
  const float a = ((fabsf(m.c0.x) + fabsf(m.c0.y)) + (fabsf(m.c0.z) + fabsf(m.c1.x))) +
                  ((fabsf(m.c1.y) + fabsf(m.c1.z)) + (fabsf(m.c2.x) + fabsf(m.c2.y))) + fabsf(m.c2.z);
  // Beyond 1e18 the reference's squares can overflow (radius = Inf or NaN: "never culled") while these sums stay
  // finite: give up there. NaN fails both comparisons and ends up at +Inf as well.
  return (e < 1e18f && a < 1e18f) ? e * a * 1.0001f + 1e-30f : __int_as_float(0x7f800000);
}

__device__ __forceinline__ void world_bounds_sphere(const Mat4& m, float bminx, float bminy, float bminz, float bmaxx,
                                                    float bmaxy, float bmaxz, float& ox, float& oy, float& oz,
                                                    float& radius)
{
  float ex, ey, ez;
  world_bounds_centre(m, bminx, bminy, bminz, bmaxx, bmaxy, bmaxz, ox, oy, oz, ex, ey, ez);
  radius = world_bounds_radius(m, ex, ey, ez);
}

// sphereInFrustum for one view: culled iff any plane has ((n0*c0 + n1*c1) + n2*c2) + d < -radius (NaN => kept)
__device__ __forceinline__ bool sphere_in_frustum(const float4* __restrict__ planes, float cx, float cy, float cz,
                                                  float radius)
{
  const float negR = -radius;
  bool inside = true;
#pragma unroll
  for (int p = 0; p < 6; ++p)
  {
    const float4 pl = planes[p];
    const float d = __fadd_rn(sum3_ref(__fmul_rn(pl.x, cx), __fmul_rn(pl.y, cy), __fmul_rn(pl.z, cz)), pl.w);
    inside = inside && !(d < negR);
  }
  return inside;
}

// ---------------------------------------------------------------------------------------------------
// glibc 2.39 expf / atanf / atan2f (generic variants) for the traffic on-rails producer (scgpu_traffic.cuh):
//   smoothExp   src/engine/traffic/sc_traffic_ai.cpp:58-62  1.0f - std::exp(-response * dt)
//   yawFromDir  src/engine/traffic/sc_traffic_ai.cpp:72-75  std::atan2(dir[0], dir[2])
// expf: sysdeps/ieee754/flt-32/e_expf.c + e_exp2f_data.c (ARM optimized-routines: N = 32 table, degree-3
// polynomial in double, one rounding to float). atanf / atan2f: sysdeps/ieee754/flt-32/s_atanf.c, e_atan2f.c
// (fdlibm, float arithmetic). Checked against the host libm (hwcaps=-FMA,-AVX2): expf and atanf over all 2^32
// inputs, atan2f over 4e8 pairs, 0 mismatches (oracle/scoracle.c header; tests/test_traffic_oracle.py repeats
// strided sweeps).
// ---------------------------------------------------------------------------------------------------

// e_exp2f_data.c: tab[i] = asuint64(2^(i/32)) - (i << 47)
__constant__ uint64_t kExp2fTab[32] = {
  0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
  0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
  0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
  0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
  0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
  0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
  0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
  0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull
};

__device__ __forceinline__ float expf_glibc(float x)
{
  const uint32_t xi = __float_as_uint(x);
  const uint32_t abstop = (xi >> 20) & 0x7ffu;
  if (abstop >= 0x42bu)  // |x| >= 88.0f or NaN
  {
    if (xi == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8u) return __fadd_rn(x, x);
    if (x > 0x1.62e42ep6f) return __int_as_float(0x7f800000);  // __math_oflowf
    if (x < -0x1.9fe368p6f) return 0.0f;                        // __math_uflowf
  }
  // x * N/ln2 = k + r, r in [-1/2, 1/2]; the poly_scaled coefficients carry the powers of 1/N (exact)
  const double z = __dmul_rn(0x1.71547652b82fep+5, (double)x);
  double kd = __dadd_rn(z, 0x1.8p+52);
  const uint64_t ki = (uint64_t)__double_as_longlong(kd);
  kd = __dsub_rn(kd, 0x1.8p+52);
  const double r = __dsub_rn(z, kd);
  const uint64_t t = kExp2fTab[ki & 31u] + (ki << 47);
  const double s = __longlong_as_double((long long)t);
  const double p = __dadd_rn(__dmul_rn(0x1.c6af84b912394p-20, r), 0x1.ebfce50fac4f3p-13);
  const double r2 = __dmul_rn(r, r);
  double y = __dadd_rn(__dmul_rn(0x1.62e42ff0c52d6p-6, r), 1.0);
  y = __dadd_rn(__dmul_rn(p, r2), y);
  return __double2float_rn(__dmul_rn(y, s));
}

__device__ __forceinline__ float atanf_glibc(float x)
{
  // atanhi/atanlo for 0.5, 1.0, 1.5, inf and the 11 polynomial coefficients of s_atanf.c
  const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
  const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
  const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f, a3 = -1.1111110449e-01f,
              a4 = 9.0908870101e-02f, a5 = -7.6918758452e-02f, a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f,
              a8 = 4.9768779427e-02f, a9 = -3.6531571299e-02f, a10 = 1.6285819933e-02f;
  const int32_t hx = (int32_t)__float_as_uint(x);
  const int32_t ix = hx & 0x7fffffff;
  if (ix >= 0x4c000000)  // |x| >= 2^25
  {
    if (ix > 0x7f800000) return __fadd_rn(x, x);
    return (hx > 0) ? __fadd_rn(hi3, lo3) : __fsub_rn(-hi3, lo3);
  }
  int id = -1;
  float hi = 0.0f, lo = 0.0f;
  if (ix < 0x3ee00000)  // |x| < 0.4375
  {
    if (ix < 0x31000000) return x;  // |x| < 2^-29
  }
  else
  {
    x = fabsf(x);
    if (ix < 0x3f980000)
    {
      if (ix < 0x3f300000) { id = 0; hi = hi0; lo = lo0; x = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, x), 1.0f), __fadd_rn(2.0f, x)); }
      else                 { id = 1; hi = hi1; lo = lo1; x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f)); }
    }
    else
    {
      if (ix < 0x401c0000) { id = 2; hi = hi2; lo = lo2; x = __fdiv_rn(__fsub_rn(x, 1.5f), __fadd_rn(1.0f, __fmul_rn(1.5f, x))); }
      else                 { id = 3; hi = hi3; lo = lo3; x = __fdiv_rn(-1.0f, x); }
    }
  }
  const float z = __fmul_rn(x, x);
  const float w = __fmul_rn(z, z);
  float e = __fadd_rn(a8, __fmul_rn(w, a10));
  e = __fadd_rn(a6, __fmul_rn(w, e));
  e = __fadd_rn(a4, __fmul_rn(w, e));
  e = __fadd_rn(a2, __fmul_rn(w, e));
  e = __fadd_rn(a0, __fmul_rn(w, e));
  const float s1 = __fmul_rn(z, e);
  float o = __fadd_rn(a7, __fmul_rn(w, a9));
  o = __fadd_rn(a5, __fmul_rn(w, o));
  o = __fadd_rn(a3, __fmul_rn(w, o));
  o = __fadd_rn(a1, __fmul_rn(w, o));
  const float s2 = __fmul_rn(w, o);
  const float xs = __fmul_rn(x, __fadd_rn(s1, s2));
  if (id < 0) return __fsub_rn(x, xs);
  const float r = __fsub_rn(hi, __fsub_rn(__fsub_rn(xs, lo), x));
  return (hx < 0) ? -r : r;
}

__device__ __forceinline__ float atan2f_glibc(float y, float x)
{
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  const int32_t hx = (int32_t)__float_as_uint(x), hy = (int32_t)__float_as_uint(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return __fadd_rn(x, y);
  if (hx == 0x3f800000) return atanf_glibc(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);  // 2*sign(x) + sign(y)
  const float piT = __fadd_rn(pi, tiny), hpiT = __fadd_rn(pi_o_2, tiny);
  if (iy == 0) return (m < 2) ? y : ((m == 2) ? piT : __fsub_rn(-pi, tiny));
  if (ix == 0) return (hy < 0) ? __fsub_rn(-pi_o_2, tiny) : hpiT;
  if (ix == 0x7f800000)
  {
    if (iy == 0x7f800000)
    {
      if (m == 0) return __fadd_rn(pi_o_4, tiny);
      if (m == 1) return __fsub_rn(-pi_o_4, tiny);
      if (m == 2) return __fadd_rn(__fmul_rn(3.0f, pi_o_4), tiny);
      return __fsub_rn(__fmul_rn(-3.0f, pi_o_4), tiny);
    }
    if (m == 0) return 0.0f;
    if (m == 1) return -0.0f;
    if (m == 2) return piT;
    return __fsub_rn(-pi, tiny);
  }
  if (iy == 0x7f800000) return (hy < 0) ? __fsub_rn(-pi_o_2, tiny) : hpiT;
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = __fadd_rn(pi_o_2, __fmul_rn(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_glibc(fabsf(__fdiv_rn(y, x)));
  if (m == 0) return z;
  if (m == 1) return __uint_as_float(__float_as_uint(z) ^ 0x80000000u);
  if (m == 2) return __fsub_rn(pi, __fsub_rn(z, pi_lo));
  return __fsub_rn(__fsub_rn(z, pi_lo), pi);
}

}  // namespace scgpu
