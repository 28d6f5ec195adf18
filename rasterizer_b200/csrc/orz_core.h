// Scalar cores of the occlusion-culling hot path, written once for device and host:
// per-quad transform / cull / edge + depth-plane setup, the state-independent front half of
// queryVisibility, the per-view matrix baking, and the x86 instruction models (rcpps, cvttps2dq,
// minps/maxps, packus) the results are defined by.  Kernels in orz_kernels.cu call these per
// lane; tests compile the same header with g++ (tests/core_host_shim.cpp) to check every field
// against the oracle without spending GPU time.
//
// Arithmetic contract (SURVEY 8a "Arithmetic spec", Appendix A): IEEE binary32, round to nearest
// even, denormals kept, NO implicit contraction -- build device code with -fmad=false and host
// code with -ffp-contract=off; ORZ_FMA marks the places where the reference has an explicit
// fmadd / fmsub / fnmadd (Rasterizer.cpp file:line cited at each use).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ORZ_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define ORZ_HD static inline
#endif

namespace orz {

#if defined(__CUDA_ARCH__)
ORZ_HD float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
ORZ_HD uint32_t f2u(float f) { return __float_as_uint(f); }
ORZ_HD float u2f(uint32_t u) { return __uint_as_float(u); }
ORZ_HD float rint_rn(float f) { return rintf(f); }
ORZ_HD float floor_f(float f) { return floorf(f); }
#else
ORZ_HD float fma_rn(float a, float b, float c) { return fmaf(a, b, c); }
ORZ_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
ORZ_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
ORZ_HD float rint_rn(float f) { return rintf(f); }
ORZ_HD float floor_f(float f) { return floorf(f); }
#endif
#define ORZ_FMA(a, b, c) ::orz::fma_rn((a), (b), (c))

constexpr uint32_t kSign = 0x80000000u;
ORZ_HD float fxor(float a, uint32_t m) { return u2f(f2u(a) ^ m); }
ORZ_HD float fabs_bits(float a) { return u2f(f2u(a) & 0x7fffffffu); }

// minps / maxps return the SECOND operand when unordered or equal
ORZ_HD float min_x86(float a, float b) { return a < b ? a : b; }
ORZ_HD float max_x86(float a, float b) { return a > b ? a : b; }
// cvttps2dq: 0x80000000 ("integer indefinite") for NaN and anything outside int32
ORZ_HD int32_t cvtt_x86(float f) {
  return (f >= -2147483648.0f && f < 2147483648.0f) ? (int32_t)f : (int32_t)0x80000000u;
}
// packDepthPremultiplied (Rasterizer.cpp:508-525): arithmetic >> 12 then unsigned saturation.
// A NaN depth (inf * 0 from a degenerate depth plane) is the NEGATIVE default NaN on x86 and packs
// to 0; GPUs produce a positive canonical NaN, so NaN is mapped to -inf first (fmaxf returns the
// non-NaN operand and leaves every other value unchanged).
ORZ_HD uint32_t pack16(float f) {
  f = fmaxf(f, u2f(0xff800000u));
  int32_t v = ((int32_t)f2u(f)) >> 12;
  return v < 0 ? 0u : (v > 65535 ? 65535u : (uint32_t)v);
}

// rcpps model: table[m >> shift] holds rcpps(1.m) for every leading-mantissa class of the host
// the table was probed on (orz_host.cpp); exponent handling as on x86.
struct RcpTable {
  const uint32_t* base;
  int shift;  // 23 - bits
};
ORZ_HD float rcp_x86(float x, const RcpTable& t) {
  uint32_t in = f2u(x), s = in & kSign, e = (in >> 23) & 0xffu, m = in & 0x7fffffu;
  if (e == 0) return u2f(s | 0x7f800000u);
  if (e == 255) return m ? u2f(in | 0x00400000u) : u2f(s);
  uint32_t b = t.base[m >> t.shift];
  int32_t re = (int32_t)((b >> 23) & 0xffu) + 127 - (int32_t)e;
  return re <= 0 ? u2f(s) : u2f(s | ((uint32_t)re << 23) | (b & 0x7fffffu));
}

// rsqrtps model (VectorMath.h:20-23 uses it in normalize): x = 2^(2k+p) * 1.m ->
// 2^-k * table[p][m >> shift]; table probed on the host over [1, 4) (orz_host.cpp).  Zero and
// denormals give +-inf, negative inputs the default NaN.
struct RsqrtTable {
  const uint32_t* base;  // [2][1 << bits]
  int bits;
};
ORZ_HD float rsqrt_x86(float x, const RsqrtTable& t) {
  const uint32_t in = f2u(x), e = (in >> 23) & 0xffu, m = in & 0x7fffffu;
  if (e == 255) return m ? u2f(in | 0x00400000u) : ((in & kSign) ? u2f(0xffc00000u) : 0.0f);
  if (e == 0) return u2f((in & kSign) | 0x7f800000u);
  if (in & kSign) return u2f(0xffc00000u);
  const int32_t ue = (int32_t)e - 127, p = ue & 1, k = (ue - p) / 2;
  const uint32_t b = t.base[((uint32_t)p << t.bits) + (m >> (23 - t.bits))];
  const int32_t re = (int32_t)((b >> 23) & 0xffu) - k;
  return u2f(((uint32_t)re << 23) | (b & 0x7fffffu));
}

// Primitive modes, numbering of Rasterizer.cpp:19-28 (ordered by frequency)
enum : uint32_t { kCulled = 0, kTriangle0, kTriangle1, kConcaveRight, kConcaveLeft, kConcaveCenter, kConvex };
// modeTable of Rasterizer.cpp:30-64 as nibbles: entry i = nibble (i & 7) of word (i >> 3);
// index bits 0-3 = area_i <= 0, bits 4-7 = W_i < 0 (Rasterizer.cpp:793-803)
#define ORZ_MODE_NIBBLES                                                                                   \
  0x01012426u, 0x01012023u, 0x01012426u, 0x01012520u, 0x01012406u, 0x00012523u, 0x01010426u, 0x01002523u, \
  0x01012026u, 0x01012523u, 0x01012520u, 0x01012525u, 0x00012426u, 0x01012503u, 0x00002222u, 0x00000222u, \
  0x01002426u, 0x01010523u, 0x00012426u, 0x01012503u, 0x01012420u, 0x01012523u, 0x01012023u, 0x01012523u, \
  0x01010426u, 0x01002523u, 0x01010101u, 0x00010101u, 0x01012424u, 0x01012520u, 0x00000000u, 0x00000000u

// ---------------------------------------------------------------------------------------------
// setModelViewProjection, Rasterizer.cpp:76-105.  `raw` = transposed input rows (frustum
// planes), `baked` = columns with the viewport (pixels, shifted half a block) and the depth
// remap [-1,1] -> [bias,0] folded in.
struct ViewMatrices {
  float baked[16];
  float raw[16];
};
ORZ_HD void bake_view_matrices(const float* m, uint32_t width, uint32_t height, ViewMatrices& vm) {
  const float sx = width * 0.5f - 4.0f, sy = height * 0.5f - 4.0f;
  const float sz = 0.5f * u2f(0x0FFFF000u);  // floatCompressionBias, Rasterizer.cpp:9
  for (int i = 0; i < 4; ++i) {
    float r0 = m[4 * i + 0], r1 = m[4 * i + 1], r2 = m[4 * i + 2], r3 = m[4 * i + 3];
    vm.raw[0 + i] = r0; vm.raw[4 + i] = r1; vm.raw[8 + i] = r2; vm.raw[12 + i] = r3;
    vm.baked[4 * i + 0] = (r0 + r3) * sx;
    vm.baked[4 * i + 1] = (r1 + r3) * sy;
    vm.baked[4 * i + 2] = (r3 - r2) * sz;
    vm.baked[4 * i + 3] = r3;
  }
}

// ---------------------------------------------------------------------------------------------
// Per rasterize() call: fold the occluder's dequantisation (refAabb, 11/11/10 packing, X bias,
// Y/Z bleed skew) into the matrix and derive z = c0 + c1 / W.  Rasterizer.cpp:616-655.
struct CallMatrix {
  float rx[4], ry[4], rw[4];
  float c0, c1;
};
ORZ_HD void prepare_call(const float* baked, const float* refMin, const float* refMax, CallMatrix& cm) {
  const float kx = 1.0f / (float)(2047ull << 21), ky = 1.0f / (float)(2047 << 10), kz = 1.0f / 1023;
  const float sx = (refMax[0] - refMin[0]) * kx, sy = (refMax[1] - refMin[1]) * ky, sz = (refMax[2] - refMin[2]) * kz;
  float rz[4], o[4][4];
  for (int k = 0; k < 4; ++k) {  // k = output component X,Y,Z,W
    float c0 = baked[0 + k], c1 = baked[4 + k], c2 = baked[8 + k], c3 = baked[12 + k];
    c3 = ORZ_FMA(c0, refMin[0], ORZ_FMA(c1, refMin[1], ORZ_FMA(c2, refMin[2], c3)));  // :625-629
    c0 = c0 * sx; c1 = c1 * sy; c2 = c2 * sz;                                            // :631-633
    c3 = ORZ_FMA(c0, (float)(1024ull << 21), c3);                                        // :636
    c1 = c1 - c0; c2 = c2 - c0;                                                          // :639-640
    o[k][0] = c0; o[k][1] = c1; o[k][2] = c2; o[k][3] = c3;
  }
  for (int i = 0; i < 4; ++i) { cm.rx[i] = o[0][i]; cm.ry[i] = o[1][i]; rz[i] = o[2][i]; cm.rw[i] = o[3][i]; }
  const float w0 = (float)(1 << 21), w1 = (float)(1 << 10);  // dpps (p0+p1)+(p2+p3), :645-655
  float Zb = (rz[0] * w0 + rz[1] * w1) + (rz[2] * 1.0f + rz[3] * 1.0f);
  float Wb = (cm.rw[0] * w0 + cm.rw[1] * w1) + (cm.rw[2] * 1.0f + cm.rw[3] * 1.0f);
  float Za = rz[3], Wa = cm.rw[3];
  cm.c0 = (Za - Zb) / (Wa - Wb);
  cm.c1 = ORZ_FMA(-cm.c0, Wa, Za);
}

// ---------------------------------------------------------------------------------------------
// What crosses from setup to traversal for one valid primitive ("What crosses
// the K1 -> K3 boundary").  All of it is independent of the depth-buffer state.
struct Prim {
  uint32_t mode;
  int32_t minX, minY, rangeX, rangeY;  // 8x8-block units, half-open
  uint32_t maxZ;                       // packed 16-bit upper bound
  float dzdx, dzdy, plane0;
  float nx[4], ny[4], off[4];          // normalised (x70), flipped edge normals; offsets at the bbox origin (+32)
  uint32_t slope[4];                   // slope index << 6
};

// Rasterizer.cpp:660-1063 for one quad (one lane of the reference's 8-wide packet).
// Returns false when the quad is culled (backface / degenerate / outside), true with `P` filled.
template <bool kClipped>
ORZ_HD bool setup_quad(const uint32_t word[4], const CallMatrix& cm, const RcpTable& rt, const uint32_t* modeNibbles,
                       int32_t blocksX, int32_t blocksY, Prim& P) {
  float x[4], y[4], invW[4];
  uint32_t wSign[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float Xi = (float)(int32_t)word[i];                    // :666 (whole word; skewed matrix undoes the bleed)
    const float Yi = (float)(int32_t)(word[i] & (2047u << 10));  // :671
    const float Zi = (float)(int32_t)(word[i] & 1023u);          // :676
    const float X = ORZ_FMA(Xi, cm.rx[0], ORZ_FMA(Yi, cm.rx[1], ORZ_FMA(Zi, cm.rx[2], cm.rx[3])));  // :686-709
    const float Y = ORZ_FMA(Xi, cm.ry[0], ORZ_FMA(Yi, cm.ry[1], ORZ_FMA(Zi, cm.ry[2], cm.ry[3])));
    const float W = ORZ_FMA(Xi, cm.rw[0], ORZ_FMA(Yi, cm.rw[1], ORZ_FMA(Zi, cm.rw[2], cm.rw[3])));
    float iw = rcp_x86(W, rt);
    if (kClipped) {  // :713-721, +-sqrt(FLT_MAX)
      const float M = u2f(0x5f7fffffu);
      iw = min_x86(M, max_x86(-M, iw));
    }
    invW[i] = iw;
    x[i] = rint_rn(X * iw) * 0.125f;  // :731-739
    y[i] = rint_rn(Y * iw) * 0.125f;
    wSign[i] = kClipped ? (f2u(iw) & kSign) : 0u;  // :759-773
  }
  float eX[5], eY[5];
#pragma unroll
  for (int i = 0; i < 4; ++i) { eX[i] = y[(i + 1) & 3] - y[i]; eY[i] = x[i] - x[(i + 1) & 3]; }  // :742-750
  const float area0 = ORZ_FMA(eX[0], eY[1], -(eX[1] * eY[0]));  // :752-755
  const float area1 = ORZ_FMA(eX[1], eY[2], -(eX[2] * eY[1]));
  const float area2 = ORZ_FMA(eX[2], eY[3], -(eX[3] * eY[2]));
  const float area3 = (area0 + area2) - area1;
  const uint32_t config =  // :776-803
      ((fxor(area0, wSign[0] ^ wSign[1] ^ wSign[2]) <= 0.0f) ? 1u : 0u) |
      ((fxor(area1, wSign[1] ^ wSign[2] ^ wSign[3]) <= 0.0f) ? 2u : 0u) |
      ((fxor(area2, wSign[0] ^ wSign[2] ^ wSign[3]) <= 0.0f) ? 4u : 0u) |
      ((fxor(area3, wSign[1] ^ wSign[0] ^ wSign[3]) <= 0.0f) ? 8u : 0u) |
      (wSign[0] >> 27) | (wSign[1] >> 26) | (wSign[2] >> 25) | (wSign[3] >> 24);
  const uint32_t mode = (modeNibbles[config >> 3] >> ((config & 7u) * 4u)) & 7u;  // :805
  if (mode == kCulled) return false;

  float minFx, minFy, maxFx, maxFy;
  if (kClipped) {  // clipless bounding box, :818-910
    const float infP = 10000.0f, infN = -10000.0f;
    float lo[2], hi[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const float* v = a ? y : x;
      float mnP[4], mxP[4], mnN[4], mxN[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        mnP[i] = wSign[i] ? infP : v[i];
        mxP[i] = fxor(mnP[i], wSign[i]);
        mnN[i] = wSign[i] ? v[i] : infP;
        mxN[i] = wSign[i] ? v[i] : infN;
      }
      const float minP = min_x86(min_x86(mnP[0], mnP[1]), min_x86(mnP[2], mnP[3]));
      const float maxP = max_x86(max_x86(mxP[0], mxP[1]), max_x86(mxP[2], mxP[3]));
      const float minN = min_x86(min_x86(mnN[0], mnN[1]), min_x86(mnN[2], mnN[3]));
      const float maxN = max_x86(max_x86(mxN[0], mxN[1]), max_x86(mxN[2], mxN[3]));
      const float incA = (maxN > minP) ? infN : minP;  // :899-900
      const float incB = (maxP > minN) ? infP : maxP;  // :902-903
      lo[a] = min_x86(incA, incB);
      hi[a] = max_x86(incA, incB);
    }
    minFx = lo[0]; maxFx = hi[0]; minFy = lo[1]; maxFy = hi[1];
  } else {  // :913-918
    minFx = min_x86(min_x86(x[0], x[1]), min_x86(x[2], x[3]));
    maxFx = max_x86(max_x86(x[0], x[1]), max_x86(x[2], x[3]));
    minFy = min_x86(min_x86(y[0], y[1]), min_x86(y[2], y[3]));
    maxFy = max_x86(max_x86(y[0], y[1]), max_x86(y[2], y[3]));
  }
  const float loAdd = 4.9999f / 8.0f, hiAdd = 11.0f / 8.0f;  // :923-926
  int32_t minX = cvtt_x86(minFx + loAdd); minX = minX < 0 ? 0 : minX;
  int32_t minY = cvtt_x86(minFy + loAdd); minY = minY < 0 ? 0 : minY;
  int32_t maxX = cvtt_x86(maxFx + hiAdd); maxX = maxX > blocksX ? blocksX : maxX;
  int32_t maxY = cvtt_x86(maxFy + hiAdd); maxY = maxY > blocksY ? blocksY : maxY;
  if (!(maxX > minX && maxY > minY)) return false;  // :929-935

  float z[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) z[i] = ORZ_FMA(invW[i], cm.c1, cm.c0);  // :945-948
  float maxZ = max_x86(max_x86(z[0], z[1]), max_x86(z[2], z[3]));
  if (kClipped && (wSign[0] | wSign[1] | wSign[2] | wSign[3])) maxZ = 1.0f;  // :953-956

  const bool tri0 = mode == kTriangle0, tri1 = mode == kTriangle1;
  bool ga = fabs_bits(area0) < fabs_bits(area2);  // :964
  ga = !tri0 && (tri1 || ga);                     // :969
  const float sel = ga ? area2 : area0;
  const float invArea = kClipped ? 1.0f / sel : rcp_x86(sel, rt);  // :972-982
  const float z12 = z[1] - z[2], z20 = z[2] - z[0], z30 = z[3] - z[0];
  eX[4] = y[0] - y[2]; eY[4] = x[2] - x[0];  // :989-990
  const float dzdx = invArea * (ga ? ORZ_FMA(-z20, eX[3], z30 * eX[4]) : ORZ_FMA(z20, eX[1], -(z12 * eX[4])));  // :993
  const float dzdy = invArea * (ga ? ORZ_FMA(-z20, eY[3], z30 * eY[4]) : ORZ_FMA(z20, eY[1], -(z12 * eY[4])));  // :994
  const float fminX = (float)minX, fminY = (float)minY;
  const float x0r = x[0] - fminX, y0r = y[0] - fminY;                         // :996-997
  P.plane0 = ORZ_FMA(-x0r, dzdx, ORZ_FMA(-y0r, dzdy, z[0]));                  // :999
  P.dzdx = dzdx; P.dzdy = dzdy;

  float nx[4] = {eX[0], eX[1], eX[2], eX[3]}, ny[4] = {eY[0], eY[1], eY[2], eY[3]};
  if (tri0) { nx[2] = eX[4]; ny[2] = eY[4]; }  // :1002-1005
  if (tri1) { nx[0] = fxor(eX[4], kSign); ny[0] = fxor(eY[4], kSign); }
  uint32_t flip[4] = {0u, 0u, 0u, 0u};
  if (kClipped) {  // :1009-1015
    flip[0] = wSign[0] ^ (tri1 ? wSign[2] : wSign[1]);
    flip[1] = wSign[1] ^ wSign[2];
    flip[2] = wSign[2] ^ (tri0 ? wSign[0] : wSign[3]);
    flip[3] = wSign[0] ^ wSign[3];
  }
  const float scale = (64 - 1) / (0.45f - (-0.45f));                          // normalizeEdge, :452-468
  const float add = 0.5f - (-0.45f) * (64 - 1) / (0.45f - (-0.45f));          // :1031
  const float smul = (64 / 2 - 1) * 0.5f / ((64 - 1) / (0.45f - (-0.45f)));   // :488
  const float sadd = (64 / 2 - 1) * 0.5f + 0.5f;                              // :489
  const float vx[4] = {x0r, x[1], x[2], x[3]}, vy[4] = {y0r, y[1], y[2], y[3]};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float inv = rcp_x86(fabs_bits(nx[e]) + fabs_bits(ny[e]), rt);
    inv = fxor(scale, flip[e]) * inv;
    const float ex = nx[e] * inv, ey = ny[e] * inv;
    float off = ORZ_FMA(-vx[e], ex, ORZ_FMA(-vy[e], ey, add));  // :1034-1037
    if (e > 0) {                                                 // :1039-1045
      off = ORZ_FMA(fminX, ex, off);
      off = ORZ_FMA(fminY, ey, off);
    }
    P.nx[e] = ex; P.ny[e] = ey; P.off[e] = off;
    P.slope[e] = (uint32_t)(((cvtt_x86(ORZ_FMA(ex, smul, sadd)) << 1) + (ey <= 0.0f ? 1 : 0)) << 6);  // :482-493
  }
  P.mode = mode; P.minX = minX; P.minY = minY; P.rangeX = maxX - minX; P.rangeY = maxY - minY;
  P.maxZ = pack16(maxZ);
  return true;
}

// ---------------------------------------------------------------------------------------------
// State-independent front half of queryVisibility (Rasterizer.cpp:123-273): frustum test, the 8
// projected corners, the near-plane epsilon, the pixel rectangle and the 16-bit max depth.
struct BoxFront {
  uint32_t status;  // 0 = not visible (outside frustum / empty rect), 1 = needsClipping (visible), 2 = test rect
  uint32_t minX, maxX, minY, maxY;  // inclusive pixels
  uint32_t maxZ;
};
enum : uint32_t { kBoxCulled = 0, kBoxNearClip = 1, kBoxRect = 2 };

ORZ_HD BoxFront box_front_half(const ViewMatrices& vm, const float* mn, const float* mx, uint32_t width, uint32_t height,
                               const RcpTable& rt) {
  BoxFront out;
  out.status = kBoxCulled; out.minX = out.maxX = out.minY = out.maxY = out.maxZ = 0;
  float ext[4], cen[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { ext[i] = mx[i] - mn[i]; cen[i] = mx[i] + mn[i]; }  // :126-127
  const float* raw = vm.raw;
  bool outside = false;
#pragma unroll
  for (int k = 0; k < 3; ++k)  // :136-167: planes row3 +- row k, dpps = (p0+p1)+(p2+p3)
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
      float p[4], o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p[i] = sgn ? raw[12 + i] - raw[4 * k + i] : raw[12 + i] + raw[4 * k + i];
        o[i] = cen[i] + fxor(ext[i], f2u(p[i]) & kSign);
      }
      const float d = (p[0] * o[0] + p[1] * o[1]) + (p[2] * o[2] + p[3] * o[3]);
      // movemask of the distances (:161-164).  With finite inputs a NaN distance can only be a
      // generated one (inf - inf, 0 * inf), which x86 creates with the sign bit set.
      outside = outside || (f2u(d) & kSign) || (d != d);
    }
  if (outside) return out;

  const float* col = vm.baked;  // :170-198
  float c[8][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float e0 = col[0 + k] * ext[0], e1 = col[4 + k] * ext[1], e2 = col[8 + k] * ext[2];
    const float c0 = ORZ_FMA(col[0 + k], mn[0], ORZ_FMA(col[4 + k], mn[1], ORZ_FMA(col[8 + k], mn[2], col[12 + k])));
    c[0][k] = c0;
    c[1][k] = c0 + e0; c[2][k] = c0 + e1; c[4][k] = c0 + e2;
    c[3][k] = c[1][k] + e1; c[5][k] = c[4][k] + e0; c[6][k] = c[2][k] + e2;
    c[7][k] = c[6][k] + e0;
  }
  const float maxExt = max_x86(max_x86(ext[0], ext[2]), max_x86(ext[1], ext[3]));  // :205-206
  const float eps = maxExt * 0.001f;
  bool nearClip = false;
#pragma unroll
  for (int k = 0; k < 8; ++k) nearClip = nearClip || (c[k][3] < eps);  // :208-213
  if (nearClip) { out.status = kBoxNearClip; return out; }

  float X[8], Y[8];
  uint32_t maxZ = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {  // :218-226, :271-273
    const float iw = rcp_x86(c[k][3], rt);
    X[k] = c[k][0] * iw; Y[k] = c[k][1] * iw;
    const uint32_t pz = pack16(c[k][2] * iw);
    maxZ = pz > maxZ ? pz : maxZ;
  }
  float mnX[4], mxX[4], mnY[4], mxY[4];  // :229-233
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mnX[i] = min_x86(X[i], X[i + 4]); mxX[i] = max_x86(X[i], X[i + 4]);
    mnY[i] = min_x86(Y[i], Y[i + 4]); mxY[i] = max_x86(Y[i], Y[i + 4]);
  }
  // :236-247 -- lanes (x:0|2, y:0|2, x:1|3, y:1|3), clamp, negate maxes, second reduction
  const float a0 = max_x86(min_x86(mnX[0], mnX[2]), 0.0f), a1 = max_x86(min_x86(mnY[0], mnY[2]), 0.0f);
  const float a2 = max_x86(min_x86(mnX[1], mnX[3]), 0.0f), a3 = max_x86(min_x86(mnY[1], mnY[3]), 0.0f);
  const float wl = (float)(width - 1), hl = (float)(height - 1);
  const float b0 = fxor(min_x86(max_x86(mxX[0], mxX[2]), wl), kSign), b1 = fxor(min_x86(max_x86(mxY[0], mxY[2]), hl), kSign);
  const float b2 = fxor(min_x86(max_x86(mxX[1], mxX[3]), wl), kSign), b3 = fxor(min_x86(max_x86(mxY[1], mxY[3]), hl), kSign);
  // :250-258 (negation in unsigned arithmetic: the indefinite value must wrap like the x86 `neg`)
  const int32_t i0 = cvtt_x86(floor_f(min_x86(a0, a2))), i1 = (int32_t)(0u - (uint32_t)cvtt_x86(floor_f(min_x86(b0, b2))));
  const int32_t i2 = cvtt_x86(floor_f(min_x86(a1, a3))), i3 = (int32_t)(0u - (uint32_t)cvtt_x86(floor_f(min_x86(b1, b3))));
  if (i0 >= i1 || i2 >= i3) return out;  // :261
  out.status = kBoxRect;
  out.minX = (uint32_t)i0; out.maxX = (uint32_t)i1; out.minY = (uint32_t)i2; out.maxY = (uint32_t)i3;
  out.maxZ = maxZ;
  return out;
}

}  // namespace orz
