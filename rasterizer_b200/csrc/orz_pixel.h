// Block update arithmetic of the traversal (Rasterizer.cpp:1241-1290) cut into EIGHTHS of an 8x8 block, written once
// for device and host (tests/pixel_host_shim.cpp checks it on the CPU against the lane-per-block form the round-1
// kernels used, which is pinned to the reference).
//
// The reference builds a block's 64 pixels from the eight depth lanes d[l] (l = 4 rr + c: pixel-column pair c of the
// row pair rr): rows rr and 8 + rr are packed to 16 bits, every other row is an avg_epu16 of those two IN PACKED SPACE.
// Rows of parity rr only need lanes 4 rr .. 4 rr + 3, and the 32-bit word i of a row (pixels 2i, 2i + 1) only lanes
// (2i) & 3 and (2i + 1) & 3 -- so a block splits into 8 independent items (rr, i) of 2 pixels x 4 rows with NO
// arithmetic shared between them: eight lanes finish one block in an eighth of the instructions, and a warp can pack
// the covered blocks of a tile four at a time instead of leaving the lanes of uncovered blocks idle.
#pragma once
#include "orz_core.h"

namespace orz {

// avg_epu16 on two packed halves: (a + b + 1) >> 1 without overflow
ORZ_HD uint32_t avg_u16x2(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7fff7fffu); }

#if defined(__CUDA_ARCH__)
ORZ_HD uint32_t max_u16x2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
ORZ_HD uint32_t min_u16x2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
// packDepthPremultiplied (Rasterizer.cpp:508-525) of two depths into one word: NaN -> -inf (see pack16), arithmetic
// >> 12, then ONE cvt.pack.sat.u16.s32 does both unsigned saturations and the merge (upper half = first source)
ORZ_HD uint32_t pack16x2(float lo, float hi) {
  const int32_t a = (int32_t)f2u(fmaxf(lo, u2f(0xff800000u))) >> 12, b = (int32_t)f2u(fmaxf(hi, u2f(0xff800000u))) >> 12;
  uint32_t r;
  asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(r) : "r"(b), "r"(a));
  return r;
}
// bit k of byte 0 of t -> 0xffff in the low half, bit k of byte 1 -> 0xffff in the high half: the bits are moved to
// the sign positions of the two bytes and prmt replicates the signs (selector nibbles 8 = sign of byte 0, 9 = of byte 1)
ORZ_HD uint32_t mask_u16x2(uint32_t t, int k) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(t << (7 - k)), "r"(0u), "r"(0x9988u));
  return r;
}
#else
ORZ_HD uint32_t max_u16x2(uint32_t a, uint32_t b) {
  const uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
  return (al > bl ? al : bl) | ((ah > bh ? ah : bh) << 16);
}
ORZ_HD uint32_t min_u16x2(uint32_t a, uint32_t b) {
  const uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
  return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
}
ORZ_HD uint32_t pack16x2(float lo, float hi) { return pack16(lo) | (pack16(hi) << 16); }
ORZ_HD uint32_t mask_u16x2(uint32_t t, int k) { return (((t >> k) & 1u) ? 0x0000ffffu : 0u) | (((t >> (8 + k)) & 1u) ? 0xffff0000u : 0u); }
#endif

// Which depth lanes / mask bits item (rr, i) of a block needs.
//   lanes:  l0 = 4 rr + ((2 i) & 3), l1 = 4 rr + ((2 i + 1) & 3); items i >= 2 take the half-step depth1 (:1243)
//   mask:   pixel px of row y <-> bit 8 px + (y odd ? 0 : 4) + (y >> 1) of the 64-bit coverage mask (:1257-1268), i.e.
//           for row 2 k + rr and pixels 2 i, 2 i + 1: bits k of bytes 0 and 1 of  (i < 2 ? mask.lo : mask.hi) >> shift
ORZ_HD uint32_t item_lane0(uint32_t rr, uint32_t i) { return 4u * rr + ((2u * i) & 3u); }
ORZ_HD uint32_t item_lane1(uint32_t rr, uint32_t i) { return 4u * rr + ((2u * i + 1u) & 3u); }
ORZ_HD uint32_t item_mask_shift(uint32_t rr, uint32_t i) { return 16u * (i & 1u) + (rr ? 0u : 4u); }

// One item: a, b = depth lanes l0, l1 at this block; t = mask word >> item_mask_shift; d0..d3 = the stored word
// (pixels 2 i, 2 i + 1) of rows rr, 2 + rr, 4 + rr, 6 + rr -- ZERO for a cleared block (then max() is the overwrite of
// Rasterizer.cpp:1271-1278).  Returns min over the item's 8 merged pixels as two packed halves (Rasterizer.cpp:1287-1290
// takes the min over all 64; the caller folds the 8 items and the two halves).
ORZ_HD uint32_t update_item(float a, float b, const float dzdx, const float dzdy, const bool halfStep, const uint32_t t, uint32_t& d0,
                            uint32_t& d1, uint32_t& d2, uint32_t& d3) {
  if (halfStep) { a = ORZ_FMA(dzdx, 0.5f, a); b = ORZ_FMA(dzdx, 0.5f, b); }  // depth1, :1243
  const float a8 = dzdy + a, b8 = dzdy + b;                                    // depth8/9, :1244-1245
  const uint32_t r0 = pack16x2(a, b), r8 = pack16x2(a8, b8);                   // rows rr, 8 + rr
  const uint32_t r4 = avg_u16x2(r0, r8);                                       // :1252
  const uint32_t r2 = avg_u16x2(r0, r4), r6 = avg_u16x2(r4, r8);               // :1253-1254
  d0 = max_u16x2(r0 & mask_u16x2(t, 0), d0);
  d1 = max_u16x2(r2 & mask_u16x2(t, 1), d1);
  d2 = max_u16x2(r4 & mask_u16x2(t, 2), d2);
  d3 = max_u16x2(r6 & mask_u16x2(t, 3), d3);
  return min_u16x2(min_u16x2(d0, d1), min_u16x2(d2, d3));
}

}  // namespace orz
