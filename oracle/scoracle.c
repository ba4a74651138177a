/* scoracle.c — plain-C CPU restatement of the reference's scene-update hot path.
 * TEST INFRASTRUCTURE ONLY (see scoracle.h). Build: oracle/Makefile, -O2 -ffp-contract=off, no -mfma.
 * All file:line citations are relative to /root/reference.
 */
#include "scoracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * libm: glibc 2.39 sinf / cosf, generic variant (no FMA contraction).
 *
 * Third-party dependency of the reference (std::sin/std::cos on float in src/core/src/sc_math.cpp:102-107),
 * not vendored in /root/reference. Pinned version: GNU libc 2.39 (Ubuntu 2.39-0ubuntu8.5), x86_64,
 * __sinf_sse2/__cosf_sse2 ifunc variants (GLIBC_TUNABLES=glibc.cpu.hwcaps=-FMA,-AVX2 selects them on
 * FMA-capable hosts). Published algorithm: Szabolcs Nagy's single-precision sincos from ARM
 * optimized-routines as imported in glibc 2.28 (sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,sincosf.h,
 * s_sincosf_data.c}): double-precision range reduction by pi/2 and two degree-7/8 polynomials, result
 * rounded once to float. tests/test_oracle_libm.py sweeps this against the host libm.
 * ---------------------------------------------------------------------------------------------- */

typedef struct
{
  double sign[4];
  double hpi_inv; /* 2/pi * 2^24 */
  double hpi;     /* pi/2 */
  double c0, c1, c2, c3, c4;
  double s1, s2, s3;
} sco_sincos_t;

static const sco_sincos_t k_sincos[2] = {
  { { 1.0, -1.0, -1.0, 1.0 },
    0x1.45F306DC9C883p+23,
    0x1.921FB54442D18p0,
    0x1p0,
    -0x1.ffffffd0c621cp-2,
    0x1.55553e1068f19p-5,
    -0x1.6c087e89a359dp-10,
    0x1.99343027bf8c3p-16,
    -0x1.555545995a603p-3,
    0x1.1107605230bc4p-7,
    -0x1.994eb3774cf24p-13 },
  { { 1.0, -1.0, -1.0, 1.0 },
    0x1.45F306DC9C883p+23,
    0x1.921FB54442D18p0,
    -0x1p0,
    0x1.ffffffd0c621cp-2,
    -0x1.55553e1068f19p-5,
    0x1.6c087e89a359dp-10,
    -0x1.99343027bf8c3p-16,
    -0x1.555545995a603p-3,
    0x1.1107605230bc4p-7,
    -0x1.994eb3774cf24p-13 }
};

/* 4/pi to 192 bits, 8 new bits per entry */
static const uint32_t k_inv_pio4[24] = {
  0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
  0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
  0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041
};

static const double k_pi63 = 0x1.921FB54442D18p-62; /* 2PI * 2^-64 */

static inline uint32_t asuint(float f)
{
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

static inline uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }

/* n even: sine polynomial; n odd: cosine polynomial. Returns float (single rounding of the double sum). */
static inline float sinf_poly(double x, double x2, const sco_sincos_t* p, int n)
{
  double x3, x4, x6, x7, s, c, c1, c2, s1;
  if ((n & 1) == 0)
  {
    x3 = x * x2;
    s1 = p->s2 + x2 * p->s3;
    x7 = x3 * x2;
    s = x + x3 * p->s1;
    return (float)(s + x7 * s1);
  }
  else
  {
    x4 = x2 * x2;
    c2 = p->c3 + x2 * p->c4;
    c1 = p->c0 + x2 * p->c1;
    x6 = x4 * x2;
    c = c1 + x4 * p->c2;
    return (float)(c + x6 * c2);
  }
}

/* |x| < 120: quotient from a scaled float->int conversion with explicit rounding */
static inline double reduce_fast(double x, const sco_sincos_t* p, int* np)
{
  double r = x * p->hpi_inv;
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return x - n * p->hpi;
}

/* |x| >= 120: 32x96 -> 128-bit fixed-point multiply by 4/pi */
static inline double reduce_large(uint32_t xi, int* np)
{
  const uint32_t* arr = &k_inv_pio4[(xi >> 26) & 15];
  int shift = (xi >> 23) & 7;
  uint64_t n, res0, res1, res2;

  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;

  res0 = xi * arr[0];
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;

  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  double x = (double)(int64_t)res0;
  *np = (int)n;
  return x * k_pi63;
}

static const float k_pio4f = 0x1.921FB6p-1f;

float sco_sinf(float y)
{
  double x = y;
  double s;
  int n;
  const sco_sincos_t* p = &k_sincos[0];

  if (abstop12(y) < abstop12(k_pio4f))
  {
    s = x * x;
    if (abstop12(y) < abstop12(0x1p-12f))
      return y;
    return sinf_poly(x, s, p, 0);
  }
  else if (abstop12(y) < abstop12(120.0f))
  {
    x = reduce_fast(x, p, &n);
    s = p->sign[n & 3];
    if (n & 2)
      p = &k_sincos[1];
    return sinf_poly(x * s, x * x, p, n);
  }
  else if (abstop12(y) < abstop12(INFINITY))
  {
    uint32_t xi = asuint(y);
    int sign = (int)(xi >> 31);
    x = reduce_large(xi, &n);
    s = p->sign[(n + sign) & 3];
    if ((n + sign) & 2)
      p = &k_sincos[1];
    return sinf_poly(x * s, x * x, p, n);
  }
  return (y - y) / (y - y); /* __math_invalidf: NaN */
}

float sco_cosf(float y)
{
  double x = y;
  double s;
  int n;
  const sco_sincos_t* p = &k_sincos[0];

  if (abstop12(y) < abstop12(k_pio4f))
  {
    double x2 = x * x;
    if (abstop12(y) < abstop12(0x1p-12f))
      return 1.0f;
    return sinf_poly(x, x2, p, 1);
  }
  else if (abstop12(y) < abstop12(120.0f))
  {
    x = reduce_fast(x, p, &n);
    s = p->sign[n & 3];
    if (n & 2)
      p = &k_sincos[1];
    return sinf_poly(x * s, x * x, p, n ^ 1);
  }
  else if (abstop12(y) < abstop12(INFINITY))
  {
    uint32_t xi = asuint(y);
    int sign = (int)(xi >> 31);
    x = reduce_large(xi, &n);
    s = p->sign[(n + sign) & 3];
    if ((n + sign) & 2)
      p = &k_sincos[1];
    return sinf_poly(x * s, x * x, p, n ^ 1);
  }
  return (y - y) / (y - y);
}

void sco_sincos_sweep(uint32_t first, uint64_t count, uint32_t stride, uint64_t* outSinHash, uint64_t* outCosHash)
{
  uint64_t hs = 0, hc = 0;
  uint32_t bits = first;
  for (uint64_t i = 0; i < count; ++i, bits += stride)
  {
    float x;
    memcpy(&x, &bits, 4);
    const float s = sco_sinf(x), c = sco_cosf(x);
    uint32_t sb = asuint(s), cb = asuint(c);
    if (s != s) sb = 0x7fc00000u;
    if (c != c) cb = 0x7fc00000u;
    hs = (hs ^ sb) * 0x100000001b3ull + bits;
    hc = (hc ^ cb) * 0x100000001b3ull + bits;
  }
  *outSinHash = hs;
  *outCosHash = hc;
}

/* ------------------------------------------------------------------------------------------------
 * Matrices (column-major, m[col*4+row]) — src/core/include/sc_math.h:8-51, src/core/src/sc_math.cpp
 * ---------------------------------------------------------------------------------------------- */

static void mat4_identity(float* m)
{
  memset(m, 0, 64);
  m[0] = m[5] = m[10] = m[15] = 1.0f;
}

/* sc_math.cpp:52-68: res = a0*b[0]; res += a1*b[1]; res += a2*b[2]; res += a3*b[3] per column, separate
 * mul and add (no FMA). The AVX path (:13-51) is arithmetically identical. */
void sco_mat4_mul(const float* a, const float* b, float* out)
{
  float r[16];
  for (int col = 0; col < 4; ++col)
  {
    const float* bc = b + col * 4;
    for (int row = 0; row < 4; ++row)
    {
      float res = a[0 + row] * bc[0];
      res = res + a[4 + row] * bc[1];
      res = res + a[8 + row] * bc[2];
      res = res + a[12 + row] * bc[3];
      r[col * 4 + row] = res;
    }
  }
  memcpy(out, r, 64);
}

/* sc_math.cpp:100-128: R = (Rz * Ry) * Rx with dense products */
void sco_mat4_rotation_xyz(float rx, float ry, float rz, float* out)
{
  const float cx = sco_cosf(rx), sx = sco_sinf(rx);
  const float cy = sco_cosf(ry), sy = sco_sinf(ry);
  const float cz = sco_cosf(rz), sz = sco_sinf(rz);

  float rxm[16], rym[16], rzm[16], t[16];
  mat4_identity(rxm);
  rxm[5] = cx; rxm[6] = sx; rxm[9] = -sx; rxm[10] = cx;
  mat4_identity(rym);
  rym[0] = cy; rym[2] = -sy; rym[8] = sy; rym[10] = cy;
  mat4_identity(rzm);
  rzm[0] = cz; rzm[1] = sz; rzm[4] = -sz; rzm[5] = cz;

  sco_mat4_mul(rzm, rym, t);
  sco_mat4_mul(t, rxm, out);
}

/* sc_math.cpp:130-142: M = T * (R * S) */
void sco_mat4_trs(const float* pos, const float* rot, const float* scale, float* out)
{
  float t[16], r[16], s[16], rs[16];
  mat4_identity(t);
  t[12] = pos[0]; t[13] = pos[1]; t[14] = pos[2];
  sco_mat4_rotation_xyz(rot[0], rot[1], rot[2], r);
  memset(s, 0, 64);
  s[0] = scale[0]; s[5] = scale[1]; s[10] = scale[2]; s[15] = 1.0f;
  sco_mat4_mul(r, s, rs);
  sco_mat4_mul(t, rs, out);
}

/* sc_math.cpp:144-207: cofactor expansion, left-to-right sums, identity when |det| <= 1e-6 */
void sco_mat4_inverse(const float* m, float* out)
{
  float o[16];
  o[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  o[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  o[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  o[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  o[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  o[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  o[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  o[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  o[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  o[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  o[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  o[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  o[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  o[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  o[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  o[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];

  const float det = m[0] * o[0] + m[1] * o[4] + m[2] * o[8] + m[3] * o[12];
  if (fabsf(det) <= 1e-6f)
  {
    mat4_identity(out);
    return;
  }
  const float inv_det = 1.0f / det;
  for (int i = 0; i < 16; ++i)
    out[i] = o[i] * inv_det;
}

/* sc_math.cpp:209-232 */
void sco_mat4_perspective_rh_zo(float fovYRadians, float aspect, float zNear, float zFar, int flipY, float* out)
{
  const float EPS = 1e-6f;
  if (fovYRadians <= EPS || aspect <= EPS || zNear <= EPS || zFar <= zNear + EPS)
  {
    mat4_identity(out);
    return;
  }
  memset(out, 0, 64);
  const float f = 1.0f / tanf(fovYRadians * 0.5f);
  out[0] = f / aspect;
  out[5] = flipY ? -f : f;
  out[10] = zFar / (zNear - zFar);
  out[14] = (zFar * zNear) / (zNear - zFar);
  out[11] = -1.0f;
}

/* ------------------------------------------------------------------------------------------------
 * Culling math — src/engine/world/sc_world_partition.cpp:1071-1144
 * ---------------------------------------------------------------------------------------------- */

static void normalize_plane(float a, float b, float c, float d, float* p)
{
  p[0] = p[1] = p[2] = p[3] = 0.0f;
  const float lenSq = a * a + b * b + c * c;
  if (lenSq > 1e-8f)
  {
    const float invLen = 1.0f / sqrtf(lenSq);
    p[0] = a * invLen;
    p[1] = b * invLen;
    p[2] = c * invLen;
    p[3] = d * invLen;
  }
}

void sco_frustum_from_viewproj(const float* m, float* pl)
{
  const float r0x = m[0], r0y = m[4], r0z = m[8], r0w = m[12];
  const float r1x = m[1], r1y = m[5], r1z = m[9], r1w = m[13];
  const float r2x = m[2], r2y = m[6], r2z = m[10], r2w = m[14];
  const float r3x = m[3], r3y = m[7], r3z = m[11], r3w = m[15];
  normalize_plane(r3x + r0x, r3y + r0y, r3z + r0z, r3w + r0w, pl + 0);  /* left   */
  normalize_plane(r3x - r0x, r3y - r0y, r3z - r0z, r3w - r0w, pl + 4);  /* right  */
  normalize_plane(r3x + r1x, r3y + r1y, r3z + r1z, r3w + r1w, pl + 8);  /* bottom */
  normalize_plane(r3x - r1x, r3y - r1y, r3z - r1z, r3w - r1w, pl + 12); /* top    */
  normalize_plane(r3x + r2x, r3y + r2y, r3z + r2z, r3w + r2w, pl + 16); /* near   */
  normalize_plane(r3x - r2x, r3y - r2y, r3z - r2z, r3w - r2w, pl + 20); /* far    */
}

int sco_sphere_in_frustum(const float* pl, const float* c, float radius)
{
  for (int i = 0; i < 6; ++i)
  {
    const float* p = pl + i * 4;
    const float d = p[0] * c[0] + p[1] * c[1] + p[2] * c[2] + p[3];
    if (d < -radius)
      return 0;
  }
  return 1;
}

void sco_world_bounds_sphere(const float* m, const float* bb, float* center, float* radius)
{
  const float cx = (bb[0] + bb[3]) * 0.5f, cy = (bb[1] + bb[4]) * 0.5f, cz = (bb[2] + bb[5]) * 0.5f;
  const float ex = (bb[3] - bb[0]) * 0.5f, ey = (bb[4] - bb[1]) * 0.5f, ez = (bb[5] - bb[2]) * 0.5f;
  center[0] = m[0] * cx + m[4] * cy + m[8] * cz + m[12];
  center[1] = m[1] * cx + m[5] * cy + m[9] * cz + m[13];
  center[2] = m[2] * cx + m[6] * cy + m[10] * cz + m[14];
  const float sx = sqrtf(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
  const float sy = sqrtf(m[4] * m[4] + m[5] * m[5] + m[6] * m[6]);
  const float sz = sqrtf(m[8] * m[8] + m[9] * m[9] + m[10] * m[10]);
  /* std::max(a,b) = (a < b) ? b : a */
  const float syz = (sy < sz) ? sz : sy;
  const float maxScale = (sx < syz) ? syz : sx;
  const float localRadius = sqrtf(ex * ex + ey * ey + ez * ez);
  *radius = localRadius * maxScale;
}

/* ------------------------------------------------------------------------------------------------
 * TransformSystem — src/core/src/sc_ecs.cpp:118-211
 * ---------------------------------------------------------------------------------------------- */

/* entity handle -> dense slot; mirrors ComponentPool's sparse array (sc_ecs.h:199-277) plus the generation
 * check of EntityManager::isAlive (sc_ecs.cpp:46-52): a handle is a valid parent iff the pool holds exactly
 * that handle. */
typedef struct
{
  uint32_t* sparse; /* index -> slot+1 */
  uint32_t size;
} sco_sparse;

static int sparse_build(sco_sparse* sp, uint32_t n, const uint32_t* entity)
{
  uint32_t maxIndex = 0;
  for (uint32_t i = 0; i < n; ++i)
  {
    const uint32_t idx = entity[i] & 0xFFFFFFu;
    if (idx > maxIndex) maxIndex = idx;
  }
  sp->size = maxIndex + 1u;
  sp->sparse = (uint32_t*)calloc(sp->size, sizeof(uint32_t));
  if (!sp->sparse) return 0;
  for (uint32_t i = 0; i < n; ++i)
    sp->sparse[entity[i] & 0xFFFFFFu] = i + 1u;
  return 1;
}

static uint32_t sparse_find(const sco_sparse* sp, const uint32_t* entity, uint32_t handle)
{
  if (handle == SCO_INVALID_ENTITY) return 0xFFFFFFFFu;
  const uint32_t idx = handle & 0xFFFFFFu;
  if (idx >= sp->size) return 0xFFFFFFFFu;
  const uint32_t s = sp->sparse[idx];
  if (s == 0 || entity[s - 1u] != handle) return 0xFFFFFFFFu;
  return s - 1u;
}

int64_t sco_transform_system(uint32_t n, const uint32_t* entity, uint32_t* parent, float* trs, float* world,
                             uint8_t* dirty)
{
  if (n == 0) return 0;
  sco_sparse sp;
  if (!sparse_build(&sp, n, entity)) return -1;

  uint32_t* pslot = (uint32_t*)malloc((size_t)n * 4);
  uint32_t* childStart = (uint32_t*)calloc((size_t)n + 1, 4);
  uint32_t* childList = (uint32_t*)malloc((size_t)n * 4);
  uint32_t* stack = (uint32_t*)malloc((size_t)n * 4);
  uint8_t* stackDirty = (uint8_t*)malloc((size_t)n);
  if (!pslot || !childStart || !childList || !stack || !stackDirty) return -1;

  /* sc_ecs.cpp:139-165: zero-scale patch, parent validity fix-ups, roots / child lists */
  uint32_t nRoots = 0;
  for (uint32_t i = 0; i < n; ++i)
  {
    float* s = trs + (size_t)i * 9 + 6;
    if (s[0] == 0.0f && s[1] == 0.0f && s[2] == 0.0f)
    {
      s[0] = s[1] = s[2] = 1.0f;
      dirty[i] = 1;
    }
    uint32_t ps = 0xFFFFFFFFu;
    if (parent[i] != SCO_INVALID_ENTITY && parent[i] != entity[i])
      ps = sparse_find(&sp, entity, parent[i]);
    if (ps == 0xFFFFFFFFu)
    {
      if (parent[i] != SCO_INVALID_ENTITY) dirty[i] = 1;
      parent[i] = SCO_INVALID_ENTITY;
    }
    else
    {
      childStart[ps + 1]++;
    }
    pslot[i] = ps;
  }
  for (uint32_t i = 0; i < n; ++i) childStart[i + 1] += childStart[i];
  {
    uint32_t* fill = (uint32_t*)malloc((size_t)n * 4);
    if (!fill) return -1;
    memcpy(fill, childStart, (size_t)n * 4);
    for (uint32_t i = 0; i < n; ++i)
      if (pslot[i] != 0xFFFFFFFFu) childList[fill[pslot[i]]++] = i;
    free(fill);
  }

  /* sc_ecs.cpp:167-210: explicit-stack DFS from the roots; nodes never reached (cycles) stay untouched */
  uint32_t top = 0;
  for (uint32_t i = 0; i < n; ++i)
    if (pslot[i] == 0xFFFFFFFFu) { stack[top] = i; stackDirty[top] = 0; ++top; ++nRoots; }

  int64_t recomputed = 0;
  while (top > 0)
  {
    --top;
    const uint32_t i = stack[top];
    const int nodeDirty = dirty[i] || stackDirty[top];
    if (nodeDirty)
    {
      float local[16];
      const float* t = trs + (size_t)i * 9;
      sco_mat4_trs(t, t + 3, t + 6, local);
      if (pslot[i] != 0xFFFFFFFFu)
        sco_mat4_mul(world + (size_t)pslot[i] * 16, local, world + (size_t)i * 16);
      else
        memcpy(world + (size_t)i * 16, local, 64);
      dirty[i] = 0;
      ++recomputed;
    }
    for (uint32_t k = childStart[i]; k < childStart[i + 1]; ++k)
    {
      stack[top] = childList[k];
      stackDirty[top] = (uint8_t)nodeDirty;
      ++top;
    }
  }
  (void)nRoots;
  free(sp.sparse); free(pslot); free(childStart); free(childList); free(stack); free(stackDirty);
  return recomputed;
}

/* ------------------------------------------------------------------------------------------------
 * CullingSystem — src/engine/world/sc_world_partition.cpp:1199-1284
 * ---------------------------------------------------------------------------------------------- */

void sco_culling_system(uint32_t n, const uint32_t* entity, const uint32_t* flags, const float* world,
                        const float* aabb, const float* viewProj, int freezeCulling, uint32_t* outVisible,
                        uint32_t* outVisibleCount, uint32_t* outCulled, uint32_t* outCulledCount,
                        uint32_t* outVisibleSlot)
{
  float planes[24];
  sco_frustum_from_viewproj(viewProj, planes);
  uint32_t nv = 0, nc = 0;
  for (uint32_t i = 0; i < n; ++i)
  {
    if (!(flags[i] & SCO_HAS_MESH)) continue; /* candidates: Transform && RenderMesh (:1206-1210) */
    int vis;
    if (freezeCulling) vis = 1;                       /* :1227-1233 */
    else if (!(flags[i] & SCO_HAS_BOUNDS)) vis = 1;   /* :1252-1256 */
    else
    {
      float c[3], r;
      sco_world_bounds_sphere(world + (size_t)i * 16, aabb + (size_t)i * 6, c, &r);
      vis = sco_sphere_in_frustum(planes, c, r);
    }
    if (vis)
    {
      if (outVisible) outVisible[nv] = entity[i];
      if (outVisibleSlot) outVisibleSlot[nv] = i;
      ++nv;
    }
    else
    {
      if (outCulled) outCulled[nc] = entity[i];
      ++nc;
    }
  }
  if (outVisibleCount) *outVisibleCount = nv;
  if (outCulledCount) *outCulledCount = nc;
}

/* ------------------------------------------------------------------------------------------------
 * RenderPrepStreamingSystem — src/engine/world/sc_world_partition.cpp:1286-1359, pushDrawItem :201-209
 * ---------------------------------------------------------------------------------------------- */

void sco_render_prep(uint32_t nVisible, const uint32_t* visibleSlot, const uint32_t* entity,
                     const uint32_t* meshMat2, const float* world, uint32_t maxDraws, void* outDraws80,
                     uint32_t* outEmitted, uint32_t* outDropped)
{
  uint32_t emitted = 0, dropped = 0;
  uint8_t* out = (uint8_t*)outDraws80;
  for (uint32_t k = 0; k < nVisible; ++k)
  {
    if (maxDraws > 0 && emitted >= maxDraws) { ++dropped; continue; }
    const uint32_t s = visibleSlot[k];
    uint8_t* d = out + (size_t)emitted * 80;
    memset(d, 0, 80);
    memcpy(d + 0, &entity[s], 4);
    memcpy(d + 4, &meshMat2[(size_t)s * 2 + 0], 4);
    memcpy(d + 8, &meshMat2[(size_t)s * 2 + 1], 4);
    memcpy(d + 16, world + (size_t)s * 16, 64);
    ++emitted;
  }
  *outEmitted = emitted;
  *outDropped = dropped;
}

int64_t sco_frame(uint32_t n, const uint32_t* entity, uint32_t* parent, float* trs, float* world, uint8_t* dirty,
                  const uint32_t* flags, const float* aabb, uint32_t nViews, const float* viewProjs,
                  uint32_t* visCounts, uint32_t* visScratch)
{
  const int64_t r = sco_transform_system(n, entity, parent, trs, world, dirty);
  for (uint32_t v = 0; v < nViews; ++v)
    sco_culling_system(n, entity, flags, world, aabb, viewProjs + (size_t)v * 16, 0, visScratch, &visCounts[v], NULL,
                       NULL, NULL);
  return r;
}

/* ------------------------------------------------------------------------------------------------
 * Traffic on rails.
 *
 * libm: glibc 2.39 expf / atanf / atan2f, generic variants. Third-party dependency of the reference
 * (std::exp in smoothExp, src/engine/traffic/sc_traffic_ai.cpp:58-62; std::atan2 in yawFromDir, :72-75), not
 * vendored in /root/reference. Published algorithms: expf = ARM optimized-routines single-precision exp as
 * imported in glibc 2.27 (sysdeps/ieee754/flt-32/e_expf.c, e_exp2f_data.c: N = 32 table of 2^(i/N), degree-3
 * polynomial in double, one rounding); atanf / atan2f = the fdlibm float routines (s_atanf.c, e_atan2f.c).
 * Checked here against the host's glibc 2.39 with GLIBC_TUNABLES=glibc.cpu.hwcaps=-FMA,-AVX2: expf and atanf
 * over all 2^32 inputs and atan2f over 4e8 pairs, 0 mismatches (the FMA ifunc variant of expf differs on 2 of
 * 2^32 inputs: 0x4202422f and 0xc27c65d9; atanf / atan2f have no FMA variant that differs).
 * ---------------------------------------------------------------------------------------------- */

static inline uint32_t tr_fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float tr_bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* e_exp2f_data.c: tab[i] = asuint64(2^(i/32)) - (i << 47) */
static const uint64_t k_exp2f_tab[32] = {
  0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
  0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
  0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
  0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
  0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
  0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
  0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
  0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull
};

float sco_expf(float x)
{
  const double N = 32.0;
  const double inv_ln2_n = 0x1.71547652b82fep+0 * N;
  const double shift = 0x1.8p+52;
  const double c0 = 0x1.c6af84b912394p-5 / N / N / N, c1 = 0x1.ebfce50fac4f3p-3 / N / N, c2 = 0x1.62e42ff0c52d6p-1 / N;
  const uint32_t abstop = (tr_fbits(x) >> 20) & 0x7ff;
  if (abstop >= (tr_fbits(88.0f) >> 20))
  {
    if (tr_fbits(x) == tr_fbits(-INFINITY)) return 0.0f;
    if (abstop >= (tr_fbits(INFINITY) >> 20)) return x + x;
    if (x > 0x1.62e42ep6f) return INFINITY; /* __math_oflowf */
    if (x < -0x1.9fe368p6f) return 0.0f;    /* __math_uflowf */
  }
  const double xd = (double)x;
  double z = inv_ln2_n * xd;
  double kd = z + shift;
  uint64_t ki;
  memcpy(&ki, &kd, 8);
  kd -= shift;
  const double r = z - kd;
  uint64_t t = k_exp2f_tab[ki % 32];
  t += ki << (52 - 5);
  double s;
  memcpy(&s, &t, 8);
  z = c0 * r + c1;
  const double r2 = r * r;
  double y = c2 * r + 1;
  y = z * r2 + y;
  y = y * s;
  return (float)y;
}

static const float k_atanhi[4] = { 4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f };
static const float k_atanlo[4] = { 5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f };
static const float k_aT[11] = { 3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                                9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                                4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f };

float sco_atanf(float x)
{
  float w, s1, s2, z;
  int32_t id;
  const int32_t hx = (int32_t)tr_fbits(x);
  const int32_t ix = hx & 0x7fffffff;
  if (ix >= 0x4c000000) /* |x| >= 2^25 */
  {
    if (ix > 0x7f800000) return x + x;
    if (hx > 0) return k_atanhi[3] + k_atanlo[3];
    return -k_atanhi[3] - k_atanlo[3];
  }
  if (ix < 0x3ee00000) /* |x| < 0.4375 */
  {
    if (ix < 0x31000000) return x; /* |x| < 2^-29 */
    id = -1;
  }
  else
  {
    x = fabsf(x);
    if (ix < 0x3f980000) /* |x| < 1.1875 */
    {
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
      else                 { id = 1; x = (x - 1.0f) / (x + 1.0f); }
    }
    else
    {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
      else                 { id = 3; x = -1.0f / x; }
    }
  }
  z = x * x;
  w = z * z;
  s1 = z * (k_aT[0] + w * (k_aT[2] + w * (k_aT[4] + w * (k_aT[6] + w * (k_aT[8] + w * k_aT[10])))));
  s2 = w * (k_aT[1] + w * (k_aT[3] + w * (k_aT[5] + w * (k_aT[7] + w * k_aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  z = k_atanhi[id] - ((x * (s1 + s2) - k_atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}

float sco_atan2f(float y, float x)
{
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  float z;
  const int32_t hx = (int32_t)tr_fbits(x), hy = (int32_t)tr_fbits(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return sco_atanf(y);
  const int32_t m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0)
  {
    switch (m)
    {
      case 0: case 1: return y;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000)
  {
    if (iy == 0x7f800000)
    {
      switch (m)
      {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    }
    switch (m)
    {
      case 0: return 0.0f;
      case 1: return -0.0f;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int32_t k = (iy - ix) >> 23;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = sco_atanf(fabsf(y / x));
  switch (m)
  {
    case 0: return z;
    case 1: return tr_bitsf(tr_fbits(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

static int tr_same(float a, float b) { return tr_fbits(a) == tr_fbits(b) || (a != a && b != b); }

uint64_t sco_unary_sweep(int which, uint32_t first, uint64_t count, uint32_t stride, float (*ref)(float))
{
  uint64_t bad = 0;
  uint32_t bits = first;
  for (uint64_t i = 0; i < count; ++i, bits += stride)
  {
    const float x = tr_bitsf(bits);
    const float mine = which == 0 ? sco_expf(x) : sco_atanf(x);
    bad += tr_same(mine, ref(x)) ? 0 : 1;
  }
  return bad;
}

uint64_t sco_atan2_sweep(uint32_t seed, uint64_t count, float (*ref)(float, float))
{
  uint64_t bad = 0;
  uint32_t r = seed ? seed : 1u;
#define TR_NEXT() (r ^= r << 13, r ^= r >> 17, r ^= r << 5, r)
  for (uint64_t i = 0; i < count; ++i)
  {
    const uint32_t a = TR_NEXT(), b = TR_NEXT();
    float y = tr_bitsf(a), x = tr_bitsf(b);
    if ((i & 3) == 1)
    { /* directions as the lane graph holds them */
      const float ang = (float)(TR_NEXT() & 0xffffff) * 3.7e-7f;
      y = sco_sinf(ang);
      x = sco_cosf(ang);
      if (i & 4) y *= 0.5f;
    }
    else if ((i & 3) == 2)
    { /* exponents within +-16 of each other */
      x = tr_bitsf((b & 0x807fffffu) | ((a & 0x0f800000u) + 0x38000000u));
      y = tr_bitsf((a & 0x807fffffu) | (((b >> 3) & 0x0f800000u) + 0x38000000u));
    }
    bad += tr_same(sco_atan2f(y, x), ref(y, x)) ? 0 : 1;
  }
#undef TR_NEXT
  return bad;
}

/* chooseNextSegment, src/engine/traffic/sc_traffic_lanes.cpp:150-169 */
static uint32_t tr_choose_next(const ScoLaneGraph* g, const float* dir, uint32_t node)
{
  uint32_t best = 0xFFFFFFFFu;
  float bestDot = -1.0f;
  for (uint32_t k = g->nodeConnOffset[node]; k < g->nodeConnOffset[node + 1]; ++k)
  {
    const uint32_t segId = g->nodeConn[k];
    if (segId >= g->nSegments) continue;
    if (!g->segActive[segId]) continue;
    const float* sd = g->segDir + (size_t)segId * 3;
    const float d = dir[0] * sd[0] + dir[1] * sd[1] + dir[2] * sd[2];
    if (d > bestDot)
    {
      bestDot = d;
      best = segId;
    }
  }
  return best;
}

int sco_lane_advance(const ScoLaneGraph* g, uint32_t* laneId, float* s, float distance, float* outPos, float* outDir)
{
  if (*laneId == 0xFFFFFFFFu || *laneId >= g->nSegments) return 0;
  float remaining = distance;
  uint32_t current = *laneId;
  float currentS = *s;
  for (uint32_t guard = 0; guard < 8; ++guard)
  {
    if (!g->segActive[current]) return 0;
    const float len = g->segLength[current];
    if (len <= 1e-5f) return 0;
    const float* dir = g->segDir + (size_t)current * 3;
    const float available = len - currentS;
    if (remaining <= available)
    {
      currentS += remaining;
      const float* a = g->nodePos + (size_t)g->segNodes[(size_t)current * 2] * 3;
      outPos[0] = a[0] + dir[0] * currentS;
      outPos[1] = a[1] + dir[1] * currentS;
      outPos[2] = a[2] + dir[2] * currentS;
      outDir[0] = dir[0]; outDir[1] = dir[1]; outDir[2] = dir[2];
      *laneId = current;
      *s = currentS;
      return 1;
    }
    remaining -= available;
    currentS = 0.0f;
    const uint32_t endNode = g->segNodes[(size_t)current * 2 + 1];
    const uint32_t next = tr_choose_next(g, dir, endNode);
    if (next == 0xFFFFFFFFu)
    {
      const float* e = g->nodePos + (size_t)endNode * 3;
      outPos[0] = e[0]; outPos[1] = e[1]; outPos[2] = e[2];
      outDir[0] = dir[0]; outDir[1] = dir[1]; outDir[2] = dir[2];
      *laneId = current;
      *s = len;
      return 1;
    }
    current = next;
  }
  return 0;
}

uint32_t sco_lane_query_nearest(const ScoLaneGraph* g, const float* pos, float* outS)
{
  uint32_t best = 0xFFFFFFFFu;
  float bestDist = 0.0f;
  int hasBest = 0;
  for (uint32_t i = 0; i < g->nSegments; ++i)
  {
    if (!g->segActive[i] || g->segLength[i] <= 1e-5f) continue;
    const float* a = g->nodePos + (size_t)g->segNodes[(size_t)i * 2] * 3;
    const float* dir = g->segDir + (size_t)i * 3;
    const float toP[3] = { pos[0] - a[0], pos[1] - a[1], pos[2] - a[2] };
    const float proj = toP[0] * dir[0] + toP[1] * dir[1] + toP[2] * dir[2];
    const float mn = (proj < g->segLength[i]) ? proj : g->segLength[i]; /* std::min(seg.length, proj) */
    const float s = (0.0f < mn) ? mn : 0.0f;                             /* std::max(0.0f, .) */
    const float dx = pos[0] - (a[0] + dir[0] * s);
    const float dy = pos[1] - (a[1] + dir[1] * s);
    const float dz = pos[2] - (a[2] + dir[2] * s);
    const float distSq = dx * dx + dy * dy + dz * dz;
    if (!hasBest || distSq < bestDist)
    {
      hasBest = 1;
      bestDist = distSq;
      best = i;
      *outS = s;
    }
  }
  return best;
}

/* laneSpeedLimit, src/engine/traffic/sc_traffic_lanes.cpp:392-400 */
static float tr_speed_limit(const ScoLaneGraph* g, uint32_t laneId)
{
  if (laneId == 0xFFFFFFFFu || laneId >= g->nSegments) return g->defaultSpeedLimit;
  const uint32_t a = g->segNodes[(size_t)laneId * 2];
  if (a >= g->nNodes) return g->defaultSpeedLimit;
  return g->nodeSpeedLimit[a];
}

void sco_traffic_ai_on_rails(const ScoLaneGraph* g, uint32_t n, uint32_t* laneId, float* laneS, float* targetSpeed,
                             float* lookAheadDist, float* trs9, const float* obstacleBrake, const uint8_t* skip, float dt,
                             int hasDebug, float dbgLookAheadDist, float dbgSpeedMultiplier, uint8_t* outMoved)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    float* tr = trs9 + (size_t)i * 9; /* localPos 0..2, localRot 3..5, localScale 6..8 */
    if (outMoved) outMoved[i] = 0;
    if (skip && skip[i]) continue;                    /* sector not Active, :218-226 */
    if (hasDebug) lookAheadDist[i] = dbgLookAheadDist; /* :237-238 */
    if (laneId[i] == 0xFFFFFFFFu)                      /* :264-272 */
    {
      float qs = 0.0f;
      const uint32_t q = sco_lane_query_nearest(g, tr, &qs);
      if (q != 0xFFFFFFFFu)
      {
        laneId[i] = q;
        laneS[i] = qs;
      }
    }
    if (laneId[i] == 0xFFFFFFFFu || laneId[i] >= g->nSegments || !g->segActive[laneId[i]]) continue; /* :274-276 */

    float target[3], dir[3];
    {
      uint32_t id = laneId[i];
      float ss = laneS[i];
      if (!sco_lane_advance(g, &id, &ss, lookAheadDist[i], target, dir)) continue; /* getLookAheadPoint :278-280 */
    }
    const float toTarget[3] = { target[0] - tr[0], 0.0f, target[2] - tr[2] };
    if (sqrtf(toTarget[0] * toTarget[0] + toTarget[1] * toTarget[1] + toTarget[2] * toTarget[2]) < 1e-4f) continue;

    float desiredSpeed = tr_speed_limit(g, laneId[i]); /* :295-298 */
    if (hasDebug) desiredSpeed *= dbgSpeedMultiplier;
    desiredSpeed = (0.0f < desiredSpeed) ? desiredSpeed : 0.0f;

    const float brake = obstacleBrake ? obstacleBrake[i] : 0.0f;
    /* on-rails branch :434-458 */
    const float desired = desiredSpeed * (1.0f - brake);
    const float t = 1.0f - sco_expf(-2.5f * dt); /* smoothExp(current, target, 2.5f, dt) */
    targetSpeed[i] = targetSpeed[i] + (desired - targetSpeed[i]) * t;
    const float travel = targetSpeed[i] * dt;
    uint32_t id = laneId[i];
    float ss = laneS[i];
    float pos[3];
    if (sco_lane_advance(g, &id, &ss, travel, pos, dir))
    {
      laneId[i] = id;
      laneS[i] = ss;
      tr[0] = pos[0];
      tr[2] = pos[2];
      tr[3] = 0.0f;
      tr[4] = sco_atan2f(dir[0], dir[2]);
      tr[5] = 0.0f;
      if (outMoved) outMoved[i] = 1;
    }
  }
}
