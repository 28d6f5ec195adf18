// TEST-ONLY: the eighth-of-a-block update (rasterizer_b200/csrc/orz_pixel.h, what the cluster kernel runs on eight
// lanes per block) compiled for the host, next to the lane-per-block form of the round-1 kernels (restated below with
// plain C for the packed-u16 intrinsics; that form is pinned to the unmodified reference by the GPU parity tests and
// follows Rasterizer.cpp:1241-1290 line by line).  Not part of the product.
#include "../rasterizer_b200/csrc/orz_pixel.h"

using namespace orz;

static uint32_t byte_perm(uint32_t x, uint32_t sel) {  // __byte_perm(x, 0, sel), no sign modes
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t s = (sel >> (4 * i)) & 7u;
    const uint32_t byte = s < 4 ? (x >> (8 * s)) & 0xffu : 0u;
    r |= byte << (8 * i);
  }
  return r;
}

extern "C" {
// round-1 form: one lane owns the block.  dv[l] = the eight depth lanes at this block, d = 8 rows x 4 words
void pixel_block_lane(const float* dv8, float dzdx, float dzdy, uint32_t mkx, uint32_t mky, uint32_t hOld, uint32_t* d, uint32_t* hOut) {
  const uint32_t keep = hOld != 1u ? 0xffffffffu : 0u;
  uint32_t r0[2][4], r4[2][4], r8[2][4];
  for (int rr = 0; rr < 2; ++rr) {
    const float* dv = dv8 + 4 * rr;
    for (int i = 0; i < 4; ++i) {
      float a = dv[(2 * i) & 3], b = dv[(2 * i + 1) & 3];
      if (i >= 2) { a = ORZ_FMA(dzdx, 0.5f, a); b = ORZ_FMA(dzdx, 0.5f, b); }
      const float a8 = dzdy + a, b8 = dzdy + b;
      r0[rr][i] = pack16(a) | (pack16(b) << 16);
      r8[rr][i] = pack16(a8) | (pack16(b8) << 16);
      r4[rr][i] = avg_u16x2(r0[rr][i], r8[rr][i]);
    }
  }
  uint32_t mnAcc = 0xffffffffu;
  for (int k = 0; k < 4; ++k)
    for (int rr = 0; rr < 2; ++rr) {
      const int y = 2 * k + rr;
      uint32_t w[4];
      for (int i = 0; i < 4; ++i)
        w[i] = k == 0 ? r0[rr][i] : k == 2 ? r4[rr][i] : k == 1 ? avg_u16x2(r0[rr][i], r4[rr][i]) : avg_u16x2(r4[rr][i], r8[rr][i]);
      const int ky = (rr ? 0 : 4) + k;
      const uint32_t lo = ((mkx >> ky) & 0x01010101u) * 0xffu, hi = ((mky >> ky) & 0x01010101u) * 0xffu;
      uint32_t v[4];
      v[0] = max_u16x2(w[0] & byte_perm(lo, 0x1100), d[4 * y + 0] & keep);
      v[1] = max_u16x2(w[1] & byte_perm(lo, 0x3322), d[4 * y + 1] & keep);
      v[2] = max_u16x2(w[2] & byte_perm(hi, 0x1100), d[4 * y + 2] & keep);
      v[3] = max_u16x2(w[3] & byte_perm(hi, 0x3322), d[4 * y + 3] & keep);
      for (int i = 0; i < 4; ++i) { d[4 * y + i] = v[i]; mnAcc = min_u16x2(mnAcc, v[i]); }
    }
  const uint32_t lo16 = mnAcc & 0xffffu, hi16 = mnAcc >> 16;
  *hOut = lo16 < hi16 ? lo16 : hi16;
}

// new form: eight independent items; a cleared block (hOld == 1) enters with zero depth
void pixel_block_items(const float* dv8, float dzdx, float dzdy, uint32_t mkx, uint32_t mky, uint32_t hOld, uint32_t* d, uint32_t* hOut) {
  if (hOld == 1u)
    for (int j = 0; j < 32; ++j) d[j] = 0u;
  uint32_t mn = 0xffffu;
  for (uint32_t rr = 0; rr < 2; ++rr)
    for (uint32_t i = 0; i < 4; ++i) {
      const float a = dv8[item_lane0(rr, i)], b = dv8[item_lane1(rr, i)];
      const uint32_t t = (i < 2 ? mkx : mky) >> item_mask_shift(rr, i);
      const uint32_t m = update_item(a, b, dzdx, dzdy, i >= 2, t, d[4 * (0 + rr) + i], d[4 * (2 + rr) + i], d[4 * (4 + rr) + i], d[4 * (6 + rr) + i]);
      const uint32_t lo16 = m & 0xffffu, hi16 = m >> 16;
      mn = lo16 < mn ? lo16 : mn;
      mn = hi16 < mn ? hi16 : mn;
    }
  *hOut = mn;
}
}
