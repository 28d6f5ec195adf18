/* TEST INFRASTRUCTURE ONLY -- see oracle_port.h.  Scalar C restatement of the reference hot path.
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/SoftwareRasterizer).  Plain float arithmetic, fmaf() only where the reference
 * has an explicit fmadd/fmsub/fnmadd, x86 semantics for min/max/cvtt written out.
 * Build with -ffp-contract=off (oracle/Makefile) so the compiler never fuses on its own. */
#include "oracle_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>

/* ---------------------------------------------------------------- bit helpers */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float fxor(float a, uint32_t m) { return u2f(f2u(a) ^ m); }
static inline float fabs_bits(float a) { return u2f(f2u(a) & 0x7fffffffu); }
#define SIGN 0x80000000u

/* x86 minps/maxps: second operand when unordered or equal */
static inline float minps(float a, float b) { return a < b ? a : b; }
static inline float maxps(float a, float b) { return a > b ? a : b; }
/* cvttps2dq: "integer indefinite" 0x80000000 for NaN and out-of-range */
static inline int32_t cvtt(float f) {
  if (!(f > -2147483904.0f && f < 2147483648.0f)) return (int32_t)0x80000000u;
  return (int32_t)f;
}
/* packDepthPremultiplied, Rasterizer.cpp:508-525: srai 12 then packus_epi32 */
static inline uint32_t pack16(float f) {
  int32_t v = ((int32_t)f2u(f)) >> 12;
  return v < 0 ? 0u : (v > 65535 ? 65535u : (uint32_t)v);
}
static inline uint32_t avg16(uint32_t a, uint32_t b) { return (a + b + 1) >> 1; }
/* dpps with all four products: (p0+p1)+(p2+p3), each product rounded */
static inline float dp4(const float* a, const float* b) {
  return (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3]);
}

/* ---------------------------------------------------------------- rcpps / rsqrtps models */
static uint32_t g_rcp_default[2048];
static uint32_t g_rsqrt_default[2 * 1024];
static const uint32_t* g_rcp = NULL;
static int g_rcp_bits = 11;
static const uint32_t* g_rsqrt = NULL;
static int g_rsqrt_bits = 10;
static uint32_t* g_rcp_owned = NULL;
static uint32_t* g_rsqrt_owned = NULL;

void orc_probe_host_rcp(uint32_t* table, int bits) {
  for (uint32_t i = 0; i < (1u << bits); ++i) {
    float x = u2f(0x3f800000u | (i << (23 - bits)));
    float y;
    _mm_store_ss(&y, _mm_rcp_ss(_mm_set_ss(x)));
    table[i] = f2u(y);
  }
}
void orc_probe_host_rsqrt(uint32_t* table, int bits) {
  for (uint32_t p = 0; p < 2; ++p)
    for (uint32_t i = 0; i < (1u << bits); ++i) {
      float x = u2f(((127u + p) << 23) | (i << (23 - bits)));
      float y;
      _mm_store_ss(&y, _mm_rsqrt_ss(_mm_set_ss(x)));
      table[(p << bits) + i] = f2u(y);
    }
}
void orc_set_rcp_table(const uint32_t* table, int bits) {
  free(g_rcp_owned);
  g_rcp_owned = (uint32_t*)malloc(sizeof(uint32_t) << bits);
  memcpy(g_rcp_owned, table, sizeof(uint32_t) << bits);
  g_rcp = g_rcp_owned;
  g_rcp_bits = bits;
}
void orc_set_rsqrt_table(const uint32_t* table, int bits) {
  free(g_rsqrt_owned);
  g_rsqrt_owned = (uint32_t*)malloc(sizeof(uint32_t) << (bits + 1));
  memcpy(g_rsqrt_owned, table, sizeof(uint32_t) << (bits + 1));
  g_rsqrt = g_rsqrt_owned;
  g_rsqrt_bits = bits;
}
void orc_use_host_tables(void) {
  orc_probe_host_rcp(g_rcp_default, 11);
  orc_probe_host_rsqrt(g_rsqrt_default, 10);
  g_rcp = g_rcp_default; g_rcp_bits = 11;
  g_rsqrt = g_rsqrt_default; g_rsqrt_bits = 10;
}
/* rcpps as a function of (sign, exponent, leading mantissa bits) -- SURVEY 7.1, verified there
 * against the hardware on all 2^32 inputs and re-verified by tests/test_oracle_port.py */
float orc_rcp(float x) {
  if (!g_rcp) orc_use_host_tables();
  uint32_t in = f2u(x), s = in & SIGN, e = (in >> 23) & 0xff, m = in & 0x7fffffu;
  if (e == 0) return u2f(s | 0x7f800000u);                  /* 0 and denormals -> +-inf */
  if (e == 255) return m ? u2f(in | 0x00400000u) : u2f(s);  /* NaN quieted, inf -> +-0 */
  uint32_t base = g_rcp[m >> (23 - g_rcp_bits)];
  int32_t re = (int32_t)((base >> 23) & 0xff) + 127 - (int32_t)e;
  if (re <= 0) return u2f(s);                               /* result underflows -> +-0 */
  return u2f(s | ((uint32_t)re << 23) | (base & 0x7fffffu));
}
/* rsqrtps: x = 2^(2k+p) * 1.m  ->  2^-k * table[p][m]   */
float orc_rsqrt(float x) {
  if (!g_rsqrt) orc_use_host_tables();
  uint32_t in = f2u(x), e = (in >> 23) & 0xff, m = in & 0x7fffffu;
  if (e == 255) return m ? u2f(in | 0x00400000u) : ((in & SIGN) ? u2f(0xffc00000u) : 0.0f);
  if (e == 0) return u2f((in & SIGN) | 0x7f800000u);        /* +-0 / denormal -> +-inf */
  if (in & SIGN) return u2f(0xffc00000u);                   /* negative -> default NaN */
  int32_t ue = (int32_t)e - 127;
  int32_t p = ue & 1, k = (ue - p) / 2;
  uint32_t base = g_rsqrt[((uint32_t)p << g_rsqrt_bits) + (m >> (23 - g_rsqrt_bits))];
  int32_t re = (int32_t)((base >> 23) & 0xff) - k;
  return u2f(((uint32_t)re << 23) | (base & 0x7fffffu));
}

/* ---------------------------------------------------------------- constants */
/* Rasterizer.cpp:9 floatCompressionBias = 0x0FFFF000 as float */
#define DEPTH_BIAS_BITS 0x0FFFF000u
enum { MODE_CULLED = 0, MODE_TRI0, MODE_TRI1, MODE_CONCAVE_RIGHT, MODE_CONCAVE_LEFT, MODE_CONCAVE_CENTER, MODE_CONVEX };
/* Rasterizer.cpp:30-64 modeTable, one nibble per entry, entry i in nibble (i & 7) of word (i >> 3) */
static const uint32_t kModeNibbles[32] = {
    0x01012426u, 0x01012023u, 0x01012426u, 0x01012520u, 0x01012406u, 0x00012523u, 0x01010426u, 0x01002523u,
    0x01012026u, 0x01012523u, 0x01012520u, 0x01012525u, 0x00012426u, 0x01012503u, 0x00002222u, 0x00000222u,
    0x01002426u, 0x01010523u, 0x00012426u, 0x01012503u, 0x01012420u, 0x01012523u, 0x01012023u, 0x01012523u,
    0x01010426u, 0x01002523u, 0x01010101u, 0x00010101u, 0x01012424u, 0x01012520u, 0x00000000u, 0x00000000u};
static inline uint32_t mode_of(uint32_t config) { return (kModeNibbles[config >> 3] >> ((config & 7) * 4)) & 7u; }

struct OrcRasterizer {
  uint32_t width, height, blocksX, blocksY;
  float baked[16]; /* prebaked columns, Rasterizer.cpp:98-104 */
  float raw[16];   /* transposed raw rows, Rasterizer.cpp:86-89 */
  int64_t* lut;
  uint16_t* depth; /* [block][y][x] */
  uint16_t* hiz;
};

/* ---------------------------------------------------------------- LUT, Rasterizer.cpp:470-480, 496-506, 527-604 */
static uint64_t transpose_mask(uint64_t mask) { /* Rasterizer.cpp:527-545 */
  uint64_t out = 0;
  for (uint32_t g = 0; g < 8; ++g)
    for (uint32_t b = 0; b < 4; ++b) {
      out |= ((mask >> (8 * g + 2 * b + 0)) & 1) << (4 + g * 8 + b);
      out |= ((mask >> (8 * g + 2 * b + 1)) & 1) << (0 + g * 8 + b);
    }
  return out;
}
void orc_build_lut(int64_t* lut) {
  memset(lut, 0, 4096 * sizeof(int64_t));
  const float offMul = (64 - 1) / (0.45f - (-0.45f)); /* Rasterizer.cpp:501 */
  const float offAdd = 0.5f - (-0.45f) * offMul;      /* Rasterizer.cpp:502 */
  for (uint32_t i = 0; i < 2000; ++i) {
    float angle = -0.1f + 6.4f * (float)i / (2000 - 1);
    float nx = cosf(angle), ny = sinf(angle);
    float l = 1.0f / (fabsf(nx) + fabsf(ny));
    nx *= l; ny *= l;
    /* builder's slope quantiser uses ny < 0 (Rasterizer.cpp:472), the runtime one ny <= 0 (:484) */
    const float mul = (64 / 2 - 1) * 0.5f, add = mul + 0.5f;
    uint32_t slope = (uint32_t)(((cvtt(fmaf(nx, mul, add)) << 1) + (ny < 0.0f ? 1 : 0)) << 6);
    for (uint32_t j = 0; j < 2000; ++j) {
      float offset = -0.6f + 1.2f * (float)j / (2000 - 1);
      float lookup = offset * offMul + offAdd;
      int32_t q = (int32_t)lookup; /* Rasterizer.cpp:504-505 */
      q = q < 0 ? 0 : (q > 63 ? 63 : q);
      uint64_t block = 0;
      for (int x = 0; x < 8; ++x)
        for (int y = 0; y < 8; ++y) {
          float d = offset + (x - 3.5f) / 8.0f * nx + (y - 3.5f) / 8.0f * ny; /* Rasterizer.cpp:581 */
          if (d <= 0.0f) block |= (uint64_t)1 << (8 * x + y);
        }
      lut[slope | (uint32_t)q] |= (int64_t)transpose_mask(block);
    }
  }
}

/* ---------------------------------------------------------------- lifecycle */
OrcRasterizer* orc_create(uint32_t width, uint32_t height, const int64_t* lut4096) {
  OrcRasterizer* r = (OrcRasterizer*)calloc(1, sizeof(*r));
  r->width = width; r->height = height; r->blocksX = width / 8; r->blocksY = height / 8;
  r->lut = (int64_t*)malloc(4096 * sizeof(int64_t));
  if (lut4096) memcpy(r->lut, lut4096, 4096 * sizeof(int64_t)); else orc_build_lut(r->lut);
  size_t blocks = (size_t)r->blocksX * r->blocksY;
  r->depth = (uint16_t*)calloc(blocks * 64, 2);
  r->hiz = (uint16_t*)calloc(blocks + 8, 2);
  return r;
}
void orc_destroy(OrcRasterizer* r) { if (r) { free(r->lut); free(r->depth); free(r->hiz); free(r); } }
const uint16_t* orc_depth(const OrcRasterizer* r) { return r->depth; }
const uint16_t* orc_hiz(const OrcRasterizer* r) { return r->hiz; }
const int64_t* orc_lut(const OrcRasterizer* r) { return r->lut; }
void orc_get_matrices(const OrcRasterizer* r, float* baked16, float* raw16) {
  memcpy(baked16, r->baked, 64); memcpy(raw16, r->raw, 64);
}

/* Rasterizer.cpp:76-105 */
void orc_set_mvp(OrcRasterizer* r, const float* m) {
  float row[4][4]; /* row[k] = coefficients producing clip component k = column k of the input */
  for (int k = 0; k < 4; ++k) for (int i = 0; i < 4; ++i) row[k][i] = m[4 * i + k];
  memcpy(r->raw, row, 64);
  float sx = r->width * 0.5f - 4.0f, sy = r->height * 0.5f - 4.0f, sz = 0.5f * u2f(DEPTH_BIAS_BITS);
  for (int i = 0; i < 4; ++i) {
    float x = (row[0][i] + row[3][i]) * sx;
    float y = (row[1][i] + row[3][i]) * sy;
    float z = (row[3][i] - row[2][i]) * sz;
    r->baked[4 * i + 0] = x; r->baked[4 * i + 1] = y; r->baked[4 * i + 2] = z; r->baked[4 * i + 3] = row[3][i];
  }
}

/* Rasterizer.cpp:107-121; depth zeroed as well = fresh-state semantics */
void orc_clear(OrcRasterizer* r) {
  size_t blocks = (size_t)r->blocksX * r->blocksY;
  for (size_t i = 0; i < blocks + 8; ++i) r->hiz[i] = 1;
  memset(r->depth, 0, blocks * 128);
}

/* ---------------------------------------------------------------- per-call matrix prep, Rasterizer.cpp:616-655 */
typedef struct { float rx[4], ry[4], rw[4], c0, c1; } CallMat;
static void prep_call(const OrcRasterizer* r, const float* refMin, const float* refMax, CallMat* cm) {
  float c[4][4]; /* c[i] = baked column i = (X,Y,Z,W) contribution of input component i */
  memcpy(c, r->baked, 64);
  float ext[3] = {refMax[0] - refMin[0], refMax[1] - refMin[1], refMax[2] - refMin[2]};
  const float kx = 1.0f / (float)(2047ull << 21), ky = 1.0f / (float)(2047 << 10), kz = 1.0f / 1023;
  float sx = ext[0] * kx, sy = ext[1] * ky, sz = ext[2] * kz;
  for (int k = 0; k < 4; ++k) {
    c[3][k] = fmaf(c[0][k], refMin[0], fmaf(c[1][k], refMin[1], fmaf(c[2][k], refMin[2], c[3][k]))); /* :625-629 */
    c[0][k] = c[0][k] * sx; c[1][k] = c[1][k] * sy; c[2][k] = c[2][k] * sz;                        /* :631-633 */
    c[3][k] = fmaf(c[0][k], (float)(1024ull << 21), c[3][k]);                                        /* :636 */
    c[1][k] = c[1][k] - c[0][k]; c[2][k] = c[2][k] - c[0][k];                                        /* :639-640 */
  }
  float rz[4];
  for (int i = 0; i < 4; ++i) { cm->rx[i] = c[i][0]; cm->ry[i] = c[i][1]; rz[i] = c[i][2]; cm->rw[i] = c[i][3]; }
  const float w[4] = {(float)(1 << 21), (float)(1 << 10), 1.0f, 1.0f};
  float Za = rz[3], Zb = dp4(rz, w), Wa = cm->rw[3], Wb = dp4(cm->rw, w);  /* :645-655 */
  cm->c0 = (Za - Zb) / (Wa - Wb);
  cm->c1 = fmaf(-cm->c0, Wa, Za);
}

/* ---------------------------------------------------------------- per-quad setup, Rasterizer.cpp:660-1063 */
static void setup_quad(const OrcRasterizer* r, const CallMat* cm, const uint32_t word[4], int clipped, OrcPrim* P) {
  float X[4], Y[4], W[4], invW[4], x[4], y[4];
  uint32_t wSign[4];
  memset(P, 0, sizeof(*P));
  for (int i = 0; i < 4; ++i) {
    float Xi = (float)(int32_t)word[i];                   /* :666, whole word, Y/Z bleed corrected by the skew */
    float Yi = (float)(int32_t)(word[i] & (2047u << 10)); /* :671 */
    float Zi = (float)(int32_t)(word[i] & 1023u);         /* :676 */
    X[i] = fmaf(Xi, cm->rx[0], fmaf(Yi, cm->rx[1], fmaf(Zi, cm->rx[2], cm->rx[3])));
    Y[i] = fmaf(Xi, cm->ry[0], fmaf(Yi, cm->ry[1], fmaf(Zi, cm->ry[2], cm->ry[3])));
    W[i] = fmaf(Xi, cm->rw[0], fmaf(Yi, cm->rw[1], fmaf(Zi, cm->rw[2], cm->rw[3])));
    if (clipped) {                                         /* :713-721, maxInvW = sqrt(FLT_MAX) */
      const float M = u2f(0x5f7fffffu);
      invW[i] = minps(M, maxps(-M, orc_rcp(W[i])));
    } else {
      invW[i] = orc_rcp(W[i]);
    }
    x[i] = rintf(X[i] * invW[i]) * 0.125f;                 /* :731-739 */
    y[i] = rintf(Y[i] * invW[i]) * 0.125f;
    wSign[i] = clipped ? (f2u(invW[i]) & SIGN) : 0u;       /* :759-773 */
  }
  float eX[5], eY[5];
  for (int i = 0; i < 4; ++i) { eX[i] = y[(i + 1) & 3] - y[i]; eY[i] = x[i] - x[(i + 1) & 3]; } /* :742-750 */
  float area0 = fmaf(eX[0], eY[1], -(eX[1] * eY[0]));    /* :752-755 */
  float area1 = fmaf(eX[1], eY[2], -(eX[2] * eY[1]));
  float area2 = fmaf(eX[2], eY[3], -(eX[3] * eY[2]));
  float area3 = (area0 + area2) - area1;
  uint32_t config =                                        /* :776-803 */
      ((fxor(area0, wSign[0] ^ wSign[1] ^ wSign[2]) <= 0.0f) ? 1u : 0u) |
      ((fxor(area1, wSign[1] ^ wSign[2] ^ wSign[3]) <= 0.0f) ? 2u : 0u) |
      ((fxor(area2, wSign[0] ^ wSign[2] ^ wSign[3]) <= 0.0f) ? 4u : 0u) |
      ((fxor(area3, wSign[1] ^ wSign[0] ^ wSign[3]) <= 0.0f) ? 8u : 0u) |
      (wSign[0] >> 27) | (wSign[1] >> 26) | (wSign[2] >> 25) | (wSign[3] >> 24);
  uint32_t mode = mode_of(config);                         /* :805 */
  if (mode == MODE_CULLED) return;

  float minFx, minFy, maxFx, maxFy;
  if (clipped) {                                           /* clipless bbox, :818-910 */
    const float infP = 10000.0f, infN = -10000.0f;
    float mnP[2][4], mxP[2][4], mnN[2][4], mxN[2][4];
    for (int i = 0; i < 4; ++i) {
      const float v[2] = {x[i], y[i]};
      for (int a = 0; a < 2; ++a) {
        mnP[a][i] = wSign[i] ? infP : v[a];
        mxP[a][i] = fxor(mnP[a][i], wSign[i]);
        mnN[a][i] = wSign[i] ? v[a] : infP;
        mxN[a][i] = wSign[i] ? v[a] : infN;
      }
    }
    float inc[2][2];
    for (int a = 0; a < 2; ++a) {
      float minP = minps(minps(mnP[a][0], mnP[a][1]), minps(mnP[a][2], mnP[a][3]));
      float maxP = maxps(maxps(mxP[a][0], mxP[a][1]), maxps(mxP[a][2], mxP[a][3]));
      float minN = minps(minps(mnN[a][0], mnN[a][1]), minps(mnN[a][2], mnN[a][3]));
      float maxN = maxps(maxps(mxN[a][0], mxN[a][1]), maxps(mxN[a][2], mxN[a][3]));
      inc[a][0] = (maxN > minP) ? infN : minP;             /* :899-900 */
      inc[a][1] = (maxP > minN) ? infP : maxP;             /* :902-903 */
    }
    minFx = minps(inc[0][0], inc[0][1]); maxFx = maxps(inc[0][0], inc[0][1]);
    minFy = minps(inc[1][0], inc[1][1]); maxFy = maxps(inc[1][0], inc[1][1]);
  } else {                                                 /* :913-918 */
    minFx = minps(minps(x[0], x[1]), minps(x[2], x[3])); maxFx = maxps(maxps(x[0], x[1]), maxps(x[2], x[3]));
    minFy = minps(minps(y[0], y[1]), minps(y[2], y[3])); maxFy = maxps(maxps(y[0], y[1]), maxps(y[2], y[3]));
  }
  const float loAdd = 4.9999f / 8.0f, hiAdd = 11.0f / 8.0f; /* :923-926 */
  int32_t minX = cvtt(minFx + loAdd); if (minX < 0) minX = 0;
  int32_t minY = cvtt(minFy + loAdd); if (minY < 0) minY = 0;
  int32_t maxX = cvtt(maxFx + hiAdd); if (maxX > (int32_t)r->blocksX) maxX = (int32_t)r->blocksX;
  int32_t maxY = cvtt(maxFy + hiAdd); if (maxY > (int32_t)r->blocksY) maxY = (int32_t)r->blocksY;
  if (!(maxX > minX && maxY > minY)) return;               /* :929-930 */

  float z[4];
  for (int i = 0; i < 4; ++i) z[i] = fmaf(invW[i], cm->c1, cm->c0); /* :945-948 */
  float maxZ = maxps(maxps(z[0], z[1]), maxps(z[2], z[3]));
  if (clipped && (wSign[0] | wSign[1] | wSign[2] | wSign[3])) maxZ = 1.0f; /* :953-956 */

  int tri0 = mode == MODE_TRI0, tri1 = mode == MODE_TRI1;
  int ga = fabs_bits(area0) < fabs_bits(area2);            /* :964 */
  ga = !tri0 && (tri1 || ga);                              /* :969 */
  float sel = ga ? area2 : area0;
  float invArea = clipped ? 1.0f / sel : orc_rcp(sel);     /* :972-982 */
  float z12 = z[1] - z[2], z20 = z[2] - z[0], z30 = z[3] - z[0];
  eX[4] = y[0] - y[2]; eY[4] = x[2] - x[0];               /* :989-990 */
  float dzdx = invArea * (ga ? fmaf(-z20, eX[3], z30 * eX[4]) : fmaf(z20, eX[1], -(z12 * eX[4]))); /* :993 */
  float dzdy = invArea * (ga ? fmaf(-z20, eY[3], z30 * eY[4]) : fmaf(z20, eY[1], -(z12 * eY[4]))); /* :994 */
  float fminX = (float)minX, fminY = (float)minY;
  float x0r = x[0] - fminX, y0r = y[0] - fminY;            /* :996-997 */
  float plane0 = fmaf(-x0r, dzdx, fmaf(-y0r, dzdy, z[0])); /* :999 */

  float nx[4] = {eX[0], eX[1], eX[2], eX[3]}, ny[4] = {eY[0], eY[1], eY[2], eY[3]};
  if (tri0) { nx[2] = eX[4]; ny[2] = eY[4]; }              /* :1002-1005 */
  if (tri1) { nx[0] = fxor(eX[4], SIGN); ny[0] = fxor(eY[4], SIGN); }
  uint32_t flip[4] = {0, 0, 0, 0};
  if (clipped) {                                           /* :1009-1015 */
    flip[0] = wSign[0] ^ (tri1 ? wSign[2] : wSign[1]);
    flip[1] = wSign[1] ^ wSign[2];
    flip[2] = wSign[2] ^ (tri0 ? wSign[0] : wSign[3]);
    flip[3] = wSign[0] ^ wSign[3];
  }
  const float scale = (64 - 1) / (0.45f - (-0.45f));       /* normalizeEdge, :452-468 */
  const float add = 0.5f - (-0.45f) * (64 - 1) / (0.45f - (-0.45f)); /* :1031 */
  const float smul = (64 / 2 - 1) * 0.5f / ((64 - 1) / (0.45f - (-0.45f))); /* :488 */
  const float sadd = (64 / 2 - 1) * 0.5f + 0.5f;           /* :489 */
  const float vx[4] = {x0r, x[1], x[2], x[3]}, vy[4] = {y0r, y[1], y[2], y[3]};
  for (int e = 0; e < 4; ++e) {
    float inv = orc_rcp(fabs_bits(nx[e]) + fabs_bits(ny[e]));
    inv = fxor(scale, flip[e]) * inv;
    nx[e] = nx[e] * inv; ny[e] = ny[e] * inv;
    float off = fmaf(-vx[e], nx[e], fmaf(-vy[e], ny[e], add)); /* :1034-1037 */
    if (e > 0) {                                           /* :1039-1045 */
      off = fmaf(fminX, nx[e], off);
      off = fmaf(fminY, ny[e], off);
    }
    P->nx[e] = nx[e]; P->ny[e] = ny[e]; P->off[e] = off;
    P->slope[e] = (uint32_t)(((cvtt(fmaf(nx[e], smul, sadd)) << 1) + (ny[e] <= 0.0f ? 1 : 0)) << 6); /* :482-493 */
  }
  P->mode = mode; P->minX = minX; P->minY = minY; P->rangeX = maxX - minX; P->rangeY = maxY - minY;
  P->maxZ = pack16(maxZ); P->dzdx = dzdx; P->dzdy = dzdy; P->plane0 = plane0;
}

void orc_setup_quad(const OrcRasterizer* r, const uint32_t word[4], const float* refMin4, const float* refMax4,
                    int clipped, OrcPrim* out) {
  CallMat cm;
  prep_call(r, refMin4, refMax4, &cm);
  setup_quad(r, &cm, word, clipped, out);
}

/* ---------------------------------------------------------------- block traversal, Rasterizer.cpp:1098-1292 */
/* Optional work statistics of the block loop (tools/lane_model.py, planning only): block visits, visits the
   reference's own HiZ test rejects, depth + HiZ updates, updates that leave every pixel unchanged, and updates a
   per-block bound could have skipped exactly: the largest of the four corner samples (rows 0 / 9, pixels 0 / 7)
   bounds every generated pixel -- all steps from the samples to the pixels are monotone, avg16 never exceeds its
   larger operand -- so when it is <= the block's HiZ (the smallest stored pixel) the max-merge changes nothing. */
static uint64_t g_stats[5];
static int g_block_bound_skip = 0;
/* Planning switch: skip the updates the corner bound proves to be no-ops.  The results must not
   change -- tests/test_oracle_port.py runs the port with the switch on against the unmodified reference. */
void orc_set_block_bound_skip(int on) { g_block_bound_skip = on; }
void orc_stats(uint64_t* out5, int reset) {
  if (out5) memcpy(out5, g_stats, sizeof g_stats);
  if (reset) memset(g_stats, 0, sizeof g_stats);
}

static void traverse(OrcRasterizer* r, const OrcPrim* P, int clipped) {
  const uint32_t blocksX = r->blocksX;
  /* _mm256_mullo_epi16 at :1054: the row offset wraps mod 65536 */
  uint32_t firstBlock = (((uint32_t)P->minY * blocksX) & 0xffffu) + (uint32_t)P->minX;
  const float s = -0.5f + 1.0f / 16.0f;                   /* :1103-1107 */
  float lineDepth[8], lineOff[4];
  for (int l = 0; l < 8; ++l) {
    float sx = s + 0.125f * (float)(l & 3), sy = s + ((l >> 2) ? 0.125f : 0.0f);
    lineDepth[l] = fmaf(P->dzdx, sx, fmaf(P->dzdy, sy, P->plane0));
  }
  for (int e = 0; e < 4; ++e) lineOff[e] = P->off[e];
  for (int32_t by = 0; by < P->rangeY; ++by) {
    float d[8], o[4];
    memcpy(d, lineDepth, sizeof d); memcpy(o, lineOff, sizeof o);
    for (int32_t bx = 0; bx < P->rangeX; ++bx) {
      size_t b = (size_t)firstBlock + (size_t)by * blocksX + (size_t)bx;
      uint32_t h = r->hiz[b];
      ++g_stats[0];
      if (!(h < P->maxZ)) ++g_stats[1];
      if (h < P->maxZ) {                                   /* :1148-1152 */
        uint64_t mask;
        int update;                                        /* does this visit write depth + HiZ? */
        if (P->mode == MODE_CONVEX) {                      /* :1155-1187; break == continue */
          if (o[0] >= 63.0f || o[1] >= 63.0f || o[2] >= 63.0f || o[3] >= 63.0f) {
            mask = 0; update = 0;
          } else {
            update = 1; /* the convex path has NO `mask == 0` test (:1186): an empty mask still turns a
                           cleared block (HiZ 1) into depth 0 / HiZ 0 */
            mask = ~(uint64_t)0;
            for (int e = 0; e < 4; ++e) {
              int32_t q = cvtt(o[e]); if (q < 0) q = 0;
              mask &= (uint64_t)r->lut[P->slope[e] | (uint32_t)q];
            }
          }
        } else {                                           /* :1188-1239 */
          uint64_t L[4];
          for (int e = 0; e < 4; ++e) {
            int32_t q = cvtt(o[e]); if (q < 0) q = 0; if (q > 63) q = 63;
            L[e] = (uint64_t)r->lut[P->slope[e] | (uint32_t)q];
          }
          switch (P->mode) {
            case MODE_TRI0: mask = L[0] & L[1] & L[2]; break;
            case MODE_TRI1: mask = L[0] & L[2] & L[3]; break;
            case MODE_CONCAVE_RIGHT: mask = (L[0] | L[3]) & (L[1] & L[2]); break;
            case MODE_CONCAVE_LEFT: mask = (L[0] & L[3]) & (L[1] | L[2]); break;
            default: /* ConcaveCenter; in the unclipped build the default falls through to ConcaveLeft */
              mask = clipped ? ((L[0] & L[1]) | (L[2] & L[3])) : ((L[0] & L[3]) & (L[1] | L[2]));
              break;
          }
          update = mask != 0;                              /* :1234-1238 */
        }
        if (update) {
          uint32_t row[10][8];                             /* rows 0,1 exact; 8,9 = next block's 0,1; :1241-1254 */
          for (int l = 0; l < 8; ++l) {
            int rr = l >> 2, px = l & 3;
            float d0 = d[l], d1 = fmaf(P->dzdx, 0.5f, d0), d8 = P->dzdy + d0, d9 = P->dzdy + d1;
            row[rr][px] = pack16(d0); row[rr][px + 4] = pack16(d1);
            row[8 + rr][px] = pack16(d8); row[8 + rr][px + 4] = pack16(d9);
          }
          for (int px = 0; px < 8; ++px) {
            row[4][px] = avg16(row[0][px], row[8][px]); row[5][px] = avg16(row[1][px], row[9][px]);
            row[2][px] = avg16(row[0][px], row[4][px]); row[3][px] = avg16(row[1][px], row[5][px]);
            row[6][px] = avg16(row[4][px], row[8][px]); row[7][px] = avg16(row[5][px], row[9][px]);
          }
          uint16_t* D = r->depth + 64 * b;
          uint32_t mn = 0xffffu;
          int changed = 0;
          {
            uint32_t c0 = row[0][0], c1 = row[0][7], c2 = row[9][0], c3 = row[9][7];
            uint32_t bound = c0 > c1 ? c0 : c1;
            if (c2 > bound) bound = c2;
            if (c3 > bound) bound = c3;
            ++g_stats[2];
            if (h != 1 && bound <= h) {
              ++g_stats[4];
              if (g_block_bound_skip) goto next_block;
            }
          }
          for (int yy = 0; yy < 8; ++yy)
            for (int px = 0; px < 8; ++px) {
              uint32_t bit = 8u * (uint32_t)px + ((yy & 1) ? 0u : 4u) + ((uint32_t)yy >> 1); /* :1257-1268 */
              uint32_t v = ((mask >> bit) & 1) ? row[yy][px] : 0u;
              if (h != 1) { uint32_t old = D[8 * yy + px]; if (old > v) v = old; } /* :1271-1278 */
              if (h == 1 || D[8 * yy + px] != (uint16_t)v) changed = 1;
              D[8 * yy + px] = (uint16_t)v;
              if (v < mn) mn = v;
            }
          r->hiz[b] = (uint16_t)mn;                        /* :1287-1290 */
          if (!changed) ++g_stats[3];
        }
      }
    next_block:
      for (int l = 0; l < 8; ++l) d[l] = P->dzdx + d[l];   /* :1145-1146, every block, hit or not */
      for (int e = 0; e < 4; ++e) o[e] = P->nx[e] + o[e];
    }
    for (int l = 0; l < 8; ++l) lineDepth[l] = lineDepth[l] + P->dzdy; /* :1130-1131 */
    for (int e = 0; e < 4; ++e) lineOff[e] = lineOff[e] + P->ny[e];
  }
}

void orc_rasterize(OrcRasterizer* r, const uint32_t* packets, uint32_t packetCount, const float* refMin4,
                   const float* refMax4, int clipped) {
  CallMat cm;
  prep_call(r, refMin4, refMax4, &cm);
  for (uint32_t p = 0; p < packetCount; p += 4)            /* 4 x __m256i = 8 quads, lane q (Occluder.cpp:146-156) */
    for (uint32_t q = 0; q < 8; ++q) {
      uint32_t word[4];
      for (int j = 0; j < 4; ++j) word[j] = packets[(size_t)(p + j) * 8 + q];
      OrcPrim P;
      setup_quad(r, &cm, word, clipped, &P);
      if (P.mode != MODE_CULLED) traverse(r, &P, clipped);
    }
}

/* ---------------------------------------------------------------- queries, Rasterizer.cpp:123-349 */
int orc_query2d(const OrcRasterizer* r, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ) {
  uint32_t bx0 = minX / 8, bx1 = maxX / 8, by0 = minY / 8, by1 = maxY / 8;
  for (uint32_t by = by0; by <= by1; ++by) {
    int32_t sY = (int32_t)(minY - 8 * by); if (sY < 0) sY = 0;
    int32_t eY = (int32_t)(maxY - 8 * by); if (eY > 7) eY = 7;
    for (uint32_t bx = bx0; bx <= bx1; ++bx) {
      size_t b = (size_t)by * r->blocksX + bx;
      if (maxZ <= r->hiz[b]) continue;                     /* :310 */
      int32_t sX = (int32_t)(minX - 8 * bx); if (sX < 0) sX = 0;
      int32_t eX = (int32_t)(maxX - 8 * bx); if (eX > 7) eX = 7;
      if (sX == 0 && eX == 7 && sY == 0 && eY == 7) return 1; /* :319-325 */
      const uint16_t* D = r->depth + 64 * b;               /* depth is 0 in cleared blocks (fresh state) */
      for (int32_t yy = sY; yy <= eY; ++yy)
        for (int32_t xx = sX; xx <= eX; ++xx)
          if (D[8 * yy + xx] < maxZ) return 1;             /* :327-343: min(depth,maxZ) != maxZ */
    }
  }
  return 0;
}

int orc_query_visibility(OrcRasterizer* r, const float* mn, const float* mx) {
  float ext[4], cen[4];
  for (int i = 0; i < 4; ++i) { ext[i] = mx[i] - mn[i]; cen[i] = mx[i] + mn[i]; } /* :126-127 */
  const float* raw = r->raw;
  for (int k = 0; k < 3; ++k)                              /* :136-167, planes row3 +- row k */
    for (int sgn = 0; sgn < 2; ++sgn) {
      float plane[4], off[4];
      for (int i = 0; i < 4; ++i) {
        plane[i] = sgn ? raw[12 + i] - raw[4 * k + i] : raw[12 + i] + raw[4 * k + i];
        off[i] = cen[i] + fxor(ext[i], f2u(plane[i]) & SIGN);
      }
      if (f2u(dp4(plane, off)) & SIGN) return 0;
    }
  float c[8][4];                                           /* :170-198 */
  const float* col = r->baked;
  float e0[4], e1[4], e2[4];
  for (int k = 0; k < 4; ++k) {
    e0[k] = col[0 + k] * ext[0]; e1[k] = col[4 + k] * ext[1]; e2[k] = col[8 + k] * ext[2];
    c[0][k] = fmaf(col[0 + k], mn[0], fmaf(col[4 + k], mn[1], fmaf(col[8 + k], mn[2], col[12 + k])));
  }
  for (int k = 0; k < 4; ++k) {
    c[1][k] = c[0][k] + e0[k]; c[2][k] = c[0][k] + e1[k]; c[4][k] = c[0][k] + e2[k];
    c[3][k] = c[1][k] + e1[k]; c[5][k] = c[4][k] + e0[k]; c[6][k] = c[2][k] + e2[k];
    c[7][k] = c[6][k] + e0[k];
  }
  float maxExt = maxps(maxps(ext[0], ext[2]), maxps(ext[1], ext[3])); /* :205-206 (lane 0 of the shuffle tree) */
  float eps = maxExt * 0.001f;
  for (int k = 0; k < 8; ++k) if (c[k][3] < eps) return 3; /* :208-213 needsClipping, visible */
  float X[8], Y[8], Z[8];
  for (int k = 0; k < 8; ++k) {                            /* :218-226 */
    float iw = orc_rcp(c[k][3]);
    X[k] = c[k][0] * iw; Y[k] = c[k][1] * iw; Z[k] = c[k][2] * iw;
  }
  float mnX[4], mxX[4], mnY[4], mxY[4];                    /* :229-233 */
  for (int i = 0; i < 4; ++i) {
    mnX[i] = minps(X[i], X[i + 4]); mxX[i] = maxps(X[i], X[i + 4]);
    mnY[i] = minps(Y[i], Y[i + 4]); mxY[i] = maxps(Y[i], Y[i + 4]);
  }
  /* :236-241 lanes (x: 0|2, y: 0|2, x: 1|3, y: 1|3), clamp, then :244-247 */
  float a0 = maxps(minps(mnX[0], mnX[2]), 0.0f), a1 = maxps(minps(mnY[0], mnY[2]), 0.0f);
  float a2 = maxps(minps(mnX[1], mnX[3]), 0.0f), a3 = maxps(minps(mnY[1], mnY[3]), 0.0f);
  float wl = (float)(r->width - 1), hl = (float)(r->height - 1);
  float b0 = minps(maxps(mxX[0], mxX[2]), wl), b1 = minps(maxps(mxY[0], mxY[2]), hl);
  float b2 = minps(maxps(mxX[1], mxX[3]), wl), b3 = minps(maxps(mxY[1], mxY[3]), hl);
  b0 = fxor(b0, SIGN); b1 = fxor(b1, SIGN); b2 = fxor(b2, SIGN); b3 = fxor(b3, SIGN);
  /* unpacklo(mins,maxs) = (a0,b0,a1,b1), unpackhi = (a2,b2,a3,b3) */
  float f0 = minps(a0, a2), f1 = minps(b0, b2), f2 = minps(a1, a3), f3 = minps(b1, b3);
  int32_t i0 = cvtt(floorf(f0)), i1 = -cvtt(floorf(f1)), i2 = cvtt(floorf(f2)), i3 = -cvtt(floorf(f3)); /* :250-258 */
  if (i0 >= i1 || i2 >= i3) return 0;                      /* :261 */
  uint32_t maxZ = 0;                                       /* :271-273 */
  for (int k = 0; k < 8; ++k) { uint32_t p = pack16(Z[k]); if (p > maxZ) maxZ = p; }
  return orc_query2d(r, (uint32_t)i0, (uint32_t)i1, (uint32_t)i2, (uint32_t)i3, maxZ) ? 1 : 0;
}

/* ---------------------------------------------------------------- readBackDepth, Rasterizer.cpp:351-399 */
void orc_readback_depth(const OrcRasterizer* r, uint8_t* target) {
  const float bias = 3.9623753e+28f;
  for (uint32_t by = 0; by < r->blocksY; ++by)
    for (uint32_t bx = 0; bx < r->blocksX; ++bx) {
      size_t b = (size_t)by * r->blocksX + bx;
      for (uint32_t yy = 0; yy < 8; ++yy) {
        uint8_t* dest = target + 4 * ((size_t)8 * bx + (size_t)r->width * (8 * by + yy));
        if (r->hiz[b] == 1) { memset(dest, 0, 32); continue; }
        for (uint32_t xx = 0; xx < 8; ++xx) {
          float depth = u2f((uint32_t)r->depth[64 * b + 8 * yy + xx] << 12) * bias;
          float lin = (2 * 0.25f) / ((0.25f + 1000.0f) - (1.0f - depth) * (1000.0f - 0.25f));
          uint32_t dd = (uint32_t)(100 * 256 * lin);
          dest[4 * xx + 0] = (uint8_t)(dd / 100); dest[4 * xx + 1] = (uint8_t)(dd % 256);
          dest[4 * xx + 2] = 0; dest[4 * xx + 3] = 255;
        }
      }
    }
}

/* ---------------------------------------------------------------- Occluder::bake, Occluder.cpp:7-181 */
static void v_normal(const float* v0, const float* v1, const float* v2, float* n) { /* VectorMath.h:6-18 */
  float a[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, b[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
  n[0] = a[1] * b[2] - a[2] * b[1]; n[1] = a[2] * b[0] - a[0] * b[2]; n[2] = a[0] * b[1] - a[1] * b[0];
}
static inline float dp3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; } /* dpps 0x7F */
static void v_normalize(float* v) { /* VectorMath.h:20-23 */
  float s = orc_rsqrt(dp3(v, v));
  v[0] *= s; v[1] *= s; v[2] *= s;
}
uint32_t orc_bake(const float* vertices, uint32_t nVerts, const float* refMin, const float* refMax,
                  uint32_t* packets, float* center4, float* bmin4, float* bmax4) {
  uint32_t nQuads = nVerts / 4;
  float* normals = (float*)malloc(sizeof(float) * 3 * nQuads);
  uint32_t* assign = (uint32_t*)calloc(nQuads, sizeof(uint32_t));
  for (uint32_t q = 0; q < nQuads; ++q) {                  /* :12-21 */
    const float* v = vertices + 16 * (size_t)q;
    float n0[3], n1[3];
    v_normal(v, v + 4, v + 8, n0); v_normal(v, v + 8, v + 12, n1);
    float* n = normals + 3 * q;
    n[0] = n0[0] + n1[0]; n[1] = n0[1] + n1[1]; n[2] = n0[2] + n1[2];
    v_normalize(n);
  }
  float cen[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, -1, 0}, {0, 0, -1}, {-1, 0, 0}}; /* :25-30 */
  int changed = 1;
  for (int iter = 0; iter < 10 && changed; ++iter) {       /* :35-78 */
    changed = 0;
    for (uint32_t q = 0; q < nQuads; ++q) {
      float best = -INFINITY; uint32_t bestK = 0;
      for (uint32_t k = 0; k < 6; ++k) {
        float d = dp3(cen[k], normals + 3 * q);
        if (d >= best) { best = d; bestK = k; }            /* comige: false on NaN */
      }
      if (assign[q] != bestK) { assign[q] = bestK; changed = 1; }
    }
    memset(cen, 0, sizeof cen);
    for (uint32_t q = 0; q < nQuads; ++q) for (int i = 0; i < 3; ++i) cen[assign[q]][i] = cen[assign[q]][i] + normals[3 * q + i];
    for (uint32_t k = 0; k < 6; ++k) v_normalize(cen[k]);
  }
  float* ordered = (float*)malloc(sizeof(float) * 4 * nVerts); /* :80-93 */
  uint32_t n = 0;
  for (uint32_t k = 0; k < 6; ++k)
    for (uint32_t q = 0; q < nQuads; ++q)
      if (assign[q] == k) { memcpy(ordered + 16 * (size_t)n, vertices + 16 * (size_t)q, 64); ++n; }
  float inv[3];
  for (int i = 0; i < 3; ++i) inv[i] = 1.0f / (refMax[i] - refMin[i]); /* :97 */
  const float scale[3] = {2047.0f, 2047.0f, 1023.0f};
  uint32_t packetCount = 0;
  for (uint32_t g = 0; g < nQuads / 8; ++g) {              /* :108-156 */
    for (uint32_t j = 0; j < 4; ++j)
      for (uint32_t q = 0; q < 8; ++q) {
        const float* v = ordered + 4 * ((size_t)g * 32 + 4 * q + j);
        int32_t c[3];
        for (int i = 0; i < 3; ++i) c[i] = cvtt(fmaf((v[i] - refMin[i]) * inv[i], scale[i], 0.5f));
        uint32_t X = (uint32_t)c[0] - 1024u;
        packets[(size_t)(4 * g + j) * 8 + q] = (X << 21) | ((uint32_t)c[1] << 10) | (uint32_t)c[2];
      }
    packetCount += 4;
  }
  float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (uint32_t i = 0; i < nVerts; ++i)                    /* :162-166 (reads `vertices`, not the ordered copy) */
    for (int k = 0; k < 4; ++k) { mn[k] = minps(vertices[4 * (size_t)i + k], mn[k]); mx[k] = maxps(vertices[4 * (size_t)i + k], mx[k]); }
  mn[3] = mx[3] = 1.0f;                                    /* :169-170 */
  for (int k = 0; k < 4; ++k) { bmin4[k] = mn[k]; bmax4[k] = mx[k]; center4[k] = (mx[k] + mn[k]) * 0.5f; }
  free(normals); free(assign); free(ordered);
  return packetCount;
}
