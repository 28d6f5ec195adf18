// CUDA kernels (sm_100a) + C ABI of the B200-native occlusion-culling rasterizer.
//
// Execution model (DESIGN.md has the long form):
//   * A *view group* of GW warps (one CTA) owns one camera view at a time and walks its
//     occluders front to back exactly like Main.cpp:192-206: gate query -> setup -> traversal.
//     Views are independent, so a persistent grid pulls views from an atomic counter and the
//     machine is filled by views, not by splitting one view.
//   * Setup (Rasterizer.cpp:660-1063): one lane per quad, one 128-bit coalesced load of the four
//     packed vertices, valid primitives compacted in order (ballot + popc prefix) into
//     shared-memory records.
//   * Binning: screen block-rows are interleaved over the warps of the group
//     (row r belongs to warp r mod GW); every warp walks the compacted list in order and touches
//     only its own rows, so per-block primitive order -- which the HiZ early-out makes
//     observable -- is preserved without atomics or sorting.
//   * Traversal (Rasterizer.cpp:1098-1292): the whole warp works on one 8x8 block at a time, one
//     32-bit word (2 pixels) of the 128-byte block per lane: coalesced 128 B read-modify-write,
//     edge masks from the 32 KB table through 4 lanes + shuffle-AND, packed-u16 SIMD depth
//     (vavg/vmax), HiZ by one warp-wide REDUX min.  The 12 iterated float add chains
//     (4 edge offsets, 8 depth lanes) are distributed over the lanes (3 FADDs per block step) and
//     advanced exactly as the reference does (y chain, then x chain from the row start).
//   * Queries (Rasterizer.cpp:123-349): one thread per box; the occluder gate uses all threads of
//     the group on the blocks of one rectangle.
// No tensor cores: nothing here is a contraction.  Build: -fmad=false (see orz_core.h).
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/orz.h"
#include "orz_core.h"
#include "orz_host.h"

#ifndef ORZ_PREFETCH_LEVEL
#define ORZ_PREFETCH_LEVEL 2  // cache level the depth prefetch targets (0 = off)
#endif
#ifndef ORZ_V2_PRELOAD
#define ORZ_V2_PRELOAD 1  // lane-per-block update: 1 = load all 8 rows up front (32 registers), 0 = L1 prefetch + row-wise loads
#endif
#ifndef ORZ_VAR_UNROLL
#define ORZ_VAR_UNROLL 4  // unroll of the chain stepping loops (1, 2, 8 measured: 4 is best)
#endif
#ifndef ORZ_SPIN_NAP
#define ORZ_SPIN_NAP 32  // ns a warp sleeps between two looks at a gate decision (0 = pure spin)
#endif
#ifndef ORZ_THREADS_PER_SM_V2
#define ORZ_THREADS_PER_SM_V2 512  // same, for the lane-per-block traversal (register cap 128)
#endif
#ifndef ORZ_THREADS_PER_SM
#define ORZ_THREADS_PER_SM 1024  // resident threads per SM the view-batch kernel is compiled for (register cap = 65536 / this)
#endif

namespace orz {
namespace cg = cooperative_groups;

__constant__ uint32_t c_modeNibbles[32] = {ORZ_MODE_NIBBLES};

constexpr uint32_t kFull = 0xffffffffu;
constexpr int kChainUnroll = ORZ_VAR_UNROLL;
constexpr int kRecStride = 21;  // odd stride: conflict-free lane-per-record stores

struct OccMeta {
  uint32_t quadOffset, quadCount, pad0, pad1;
  float refMin[4], refMax[4], boundsMin[4], boundsMax[4], center[4];
};

struct Target {
  uint16_t* depth;  // [block][row][px], 128 B per 8x8 block
  uint16_t* hiz;    // [block]
  uint32_t width, height, blocksX, blocksY;
};

// ---------------------------------------------------------------------------------------------
// primitive record <-> registers
__device__ __forceinline__ void store_record(uint32_t* rec, const Prim& P) {
  rec[0] = (uint32_t)P.minX | ((uint32_t)P.minY << 16);
  rec[1] = (uint32_t)P.rangeX | ((uint32_t)P.rangeY << 16);
  rec[2] = P.maxZ | (P.mode << 16);
  rec[3] = f2u(P.dzdx); rec[4] = f2u(P.dzdy); rec[5] = f2u(P.plane0);
#pragma unroll
  for (int e = 0; e < 4; ++e) { rec[6 + e] = f2u(P.nx[e]); rec[10 + e] = f2u(P.ny[e]); rec[14 + e] = f2u(P.off[e]); }
  rec[18] = (P.slope[0] & 0xfc0u) | ((P.slope[1] & 0xfc0u) << 16);
  rec[19] = (P.slope[2] & 0xfc0u) | ((P.slope[3] & 0xfc0u) << 16);
  rec[20] = (1024u + (uint32_t)P.rangeX - 1u) / (uint32_t)P.rangeX;  // lane / rangeX == (lane * this) >> 10 for lane < 32
}

__device__ __forceinline__ void prefetch_line(const void* p) {
#if ORZ_PREFETCH_LEVEL == 1
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif ORZ_PREFETCH_LEVEL == 2
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// avg_epu16 on two packed halves: (a + b + 1) >> 1 without overflow
__device__ __forceinline__ uint32_t avg_u16x2(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7fff7fffu); }

// ---------------------------------------------------------------------------------------------
// Traversal of one primitive by one warp, restricted to the screen block-rows the warp owns
// (row % rowStride == rowPhase).  Rasterizer.cpp:1098-1292.
//
// Lane roles: lane = 4*y + w addresses word w (pixels 2w, 2w+1) of row y of the 8x8 block.
//   edge e = lane & 3 (offset chain e is replicated in the 8 lanes with that residue)
//   depth chains: row parity r = y & 1 -> reference lanes 4r + 2(w&1) and 4r + 2(w&1) + 1;
//   words 2,3 (pixels 4-7) use depth1 = depth0 + dzdx/2 (Rasterizer.cpp:1243).
// Block addressing is linear with the reference's 16-bit wrap of the first-row offset
// (Rasterizer.cpp:1054); ownership follows the linear index, so a wrapped row that
// straddles two screen rows is split between their owners.
template <int kStride, bool kWindowed = false>
__device__ __forceinline__ void raster_prim(const uint32_t* __restrict__ rec, const int lane, const uint32_t rowPhase,
                                            const uint32_t rowStrideRt, const Target& T, const uint2* __restrict__ lut,
                                            const uint32_t colLo = 0u, const uint32_t colHi = 0xffffffffu) {
  const uint32_t rowStride = kStride > 0 ? (uint32_t)kStride : rowStrideRt;
  const uint32_t w0 = rec[0], w1 = rec[1], w2 = rec[2];
  const uint32_t minX = w0 & 0xffffu, minY = w0 >> 16, rangeX = w1 & 0xffffu, rangeY = w1 >> 16;
  const uint32_t maxZ = w2 & 0xffffu, mode = w2 >> 16;
  const uint32_t blocksX = T.blocksX;

  const uint32_t fb = ((minY * blocksX) & 0xffffu) + minX;
  uint32_t r0 = minY, c0 = minX;
  if (blocksX * T.blocksY > 65536u) { r0 = fb / blocksX; c0 = fb - r0 * blocksX; }
  const uint32_t split = min(rangeX, blocksX - c0);  // blocks of a primitive row inside screen row r0 + by
  const bool crossing = split < rangeX;
  {  // any row of mine in [r0, r0 + rangeY + crossing) ?
    const uint32_t first = r0 + (rowPhase + rowStride - r0 % rowStride) % rowStride;
    if (first >= r0 + rangeY + (crossing ? 1u : 0u)) return;
  }

  const int e = lane & 3;
  const float nxe = u2f(rec[6 + e]), nye = u2f(rec[10 + e]);
  float lineOff = u2f(rec[14 + e]);
  const uint32_t slope = (rec[18 + (e >> 1)] >> ((e & 1) * 16)) & 0xffffu;
  const float dzdx = u2f(rec[3]), dzdy = u2f(rec[4]), plane0 = u2f(rec[5]);
  const int rpar = (lane >> 2) & 1, wIdx = lane & 3, k2 = lane >> 3;
  const float s = -0.5f + 1.0f / 16.0f;  // Rasterizer.cpp:1103
  const float sy = rpar ? s + 0.125f : s;
  const float sxA = s + 0.125f * (float)(2 * (wIdx & 1)), sxB = s + 0.125f * (float)(2 * (wIdx & 1) + 1);
  const float base = ORZ_FMA(dzdy, sy, plane0);  // Rasterizer.cpp:1104-1107
  float lineA = ORZ_FMA(dzdx, sxA, base), lineB = ORZ_FMA(dzdx, sxB, base);
  const bool upperHalf = (wIdx & 2) != 0;
  const uint32_t sh0 = (uint32_t)(wIdx & 1) * 16u + (rpar ? 0u : 4u) + (uint32_t)k2;  // mask bit of pixel 2w (Rasterizer.cpp:1257-1268)
  const uint32_t sh1 = sh0 + 8u;
  const bool convex = mode == kConvex;
  const uint32_t selHi = (k2 & 2) ? 0xffffffffu : 0u, selOdd = (k2 & 1) ? 0xffffffffu : 0u;
  uint32_t* const depthWords = reinterpret_cast<uint32_t*>(T.depth) + lane;

  uint32_t rowMod = r0 % rowStride;  // (r0 + by) % rowStride, kept incrementally
  for (uint32_t by = 0; by < rangeY; ++by) {
    const bool mineA = rowMod == rowPhase;
    rowMod = rowMod + 1u == rowStride ? 0u : rowMod + 1u;
    const bool mineB = crossing && rowMod == rowPhase;
    if (mineA || mineB) {
      // The x chain restarts from the row start (Rasterizer.cpp:1136-1137).  Steps are applied
      // lazily: `owed` counts the adds still to do before the next block that is really visited,
      // so blocks behind the last HiZ candidate of the row cost nothing.
      float o = lineOff, dA = lineA, dB = lineB;
      uint32_t owed = 0;
      bool hitInRow = false, rowDone = false;
      const uint32_t L = fb + by * blocksX;
      uint32_t a = 0;
#pragma unroll 1
      for (int piece = 0; piece < 2 && !rowDone; ++piece) {
        const uint32_t b = piece == 0 ? split : rangeX;
        const bool mine = piece == 0 ? mineA : mineB;
        if (!mine) { owed += b - a; a = b; continue; }
        uint32_t xLo = a, xHi = b;  // blocks of this piece inside my column window [colLo, colHi)
        if (kWindowed) {
          const uint32_t pieceCol = piece == 0 ? c0 : 0u;  // screen column of block `a`
          xLo = a + (colLo > pieceCol ? colLo - pieceCol : 0u);
          xHi = colHi > pieceCol ? min(b, a + (colHi - pieceCol)) : a;
          if (xLo >= xHi) { owed += b - a; a = b; continue; }
          owed += xLo - a;
        }
#pragma unroll 1
        for (uint32_t s0 = xLo; s0 < xHi && !rowDone; s0 += 32) {
          const uint32_t m = min(32u, xHi - s0);
          const uint32_t hv = (uint32_t)lane < m ? (uint32_t)T.hiz[L + s0 + lane] : 0xffffu;
          uint32_t cand = __ballot_sync(kFull, hv < maxZ);  // Rasterizer.cpp:1148-1152
          const uint32_t cleared = __ballot_sync(kFull, hv == 1u);
          // One instruction pulls the stored depth of every candidate block of the segment towards
          // the SM (lane j -> block j): the blocks are then visited one after the other, and
          // without this each visit would expose a full HBM round trip (memory-level parallelism
          // per warp would be 1).
          if (hv < maxZ && hv != 1u) prefetch_line(depthWords + (size_t)(L + s0 + (uint32_t)lane) * 32u - lane);
          uint32_t pos = 0;
          while (cand) {
            const uint32_t j = (uint32_t)__ffs((int)cand) - 1u;
            cand &= cand - 1u;
            const uint32_t steps = owed + j - pos;
#pragma unroll 1
            for (uint32_t i = 0; i < steps; ++i) { o = nxe + o; dA = dzdx + dA; dB = dzdx + dB; }  // Rasterizer.cpp:1145-1146
            owed = 0; pos = j;
            const uint32_t blk = L + s0 + j;
            uint2 mk;
            if (convex) {  // Rasterizer.cpp:1155-1187
              if (__any_sync(kFull, o >= 63.0f)) {
                if (hitInRow) { rowDone = true; break; }  // convexity: nothing further in this row (:1161-1165)
                continue;
              }
              hitInRow = true;
              // max(cvtt(o), 0) for o < 63 or NaN: NaN and negatives give 0
              const uint32_t q = (uint32_t)__float2int_rz(fmaxf(o, 0.0f));
              // A & B & C & D (Rasterizer.cpp:1184): the four edge masks sit in lanes e = 0..3 (replicated
              // 8x), so a warp-wide AND reduction combines them in two REDUX instructions
              const uint2 t = lut[slope | q];
              mk.x = __reduce_and_sync(kFull, t.x);
              mk.y = __reduce_and_sync(kFull, t.y);  // no empty-mask test on this path (Rasterizer.cpp:1186)
            } else {  // Rasterizer.cpp:1188-1239
              // min(max(cvtt(o), 0), 63): NaN and anything >= 2^31 convert to 0x80000000 -> 0
              const uint32_t q = o < 2147483648.0f ? (uint32_t)__float2int_rz(fminf(fmaxf(o, 0.0f), 63.0f)) : 0u;
              const uint2 t = lut[slope | q];
              const int g = lane & ~3;
              uint2 A, B, C, D;
              A.x = __shfl_sync(kFull, t.x, g + 0); A.y = __shfl_sync(kFull, t.y, g + 0);
              B.x = __shfl_sync(kFull, t.x, g + 1); B.y = __shfl_sync(kFull, t.y, g + 1);
              C.x = __shfl_sync(kFull, t.x, g + 2); C.y = __shfl_sync(kFull, t.y, g + 2);
              D.x = __shfl_sync(kFull, t.x, g + 3); D.y = __shfl_sync(kFull, t.y, g + 3);
              if (mode == kTriangle0) { mk.x = A.x & B.x & C.x; mk.y = A.y & B.y & C.y; }
              else if (mode == kTriangle1) { mk.x = A.x & C.x & D.x; mk.y = A.y & C.y & D.y; }
              else if (mode == kConcaveRight) { mk.x = (A.x | D.x) & (B.x & C.x); mk.y = (A.y | D.y) & (B.y & C.y); }
              else if (mode == kConcaveCenter) { mk.x = (A.x & B.x) | (C.x & D.x); mk.y = (A.y & B.y) | (C.y & D.y); }
              else { mk.x = (A.x & D.x) & (B.x | C.x); mk.y = (A.y & D.y) & (B.y | C.y); }
              if ((mk.x | mk.y) == 0u) continue;
            }
            uint32_t* dptr = depthWords + (size_t)blk * 32u;
            uint32_t old = 0u;
            if (((cleared >> j) & 1u) == 0u) old = *dptr;  // Rasterizer.cpp:1271-1278
            // ---- depth of this lane's two pixels, Rasterizer.cpp:1241-1254
            float a0 = dA, b0 = dB;
            if (upperHalf) { a0 = ORZ_FMA(dzdx, 0.5f, a0); b0 = ORZ_FMA(dzdx, 0.5f, b0); }
            const float a8 = dzdy + a0, b8 = dzdy + b0;
            const uint32_t v0 = pack16(a0) | (pack16(b0) << 16);  // row rpar
            const uint32_t v8 = pack16(a8) | (pack16(b8) << 16);  // row 8 + rpar
            const uint32_t mid = avg_u16x2(v0, v8);               // row 4 + rpar
            const uint32_t near8 = v0 ^ ((v0 ^ v8) & selHi);      // k2 >= 2 ? v8 : v0
            const uint32_t quarter = avg_u16x2(near8, mid);       // rows 2 + rpar / 6 + rpar
            const uint32_t even = v0 ^ ((v0 ^ mid) & selHi);      // k2 == 0 ? v0 : mid   (for even k2)
            uint32_t val = even ^ ((even ^ quarter) & selOdd);    // odd k2 -> the quarter rows
            // ---- coverage of the two pixels, Rasterizer.cpp:1257-1268
            const uint32_t mw = upperHalf ? mk.y : mk.x;
            const uint32_t selMask = ((0u - ((mw >> sh0) & 1u)) & 0x0000ffffu) | ((0u - ((mw >> sh1) & 1u)) & 0xffff0000u);
            val &= selMask;
            // ---- merge, store, HiZ; Rasterizer.cpp:1271-1290
            val = __vmaxu2(val, old);
            *dptr = val;
            uint32_t mn = min(val & 0xffffu, val >> 16);
            mn = __reduce_min_sync(kFull, mn);
            if (lane == 0) T.hiz[blk] = (uint16_t)mn;
          }
          owed += m - pos;
        }
        if (kWindowed) owed += b - xHi;
        a = b;
      }
    }
    lineA = lineA + dzdy; lineB = lineB + dzdy; lineOff = lineOff + nye;  // Rasterizer.cpp:1130-1131
  }
  // HiZ is written by lane 0 and prefetched by other lanes for the next primitive: order the
  // warp's memory accesses (each block is visited at most once per primitive, so once is enough)
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Traversal, second mapping: ONE LANE PER 8x8 BLOCK (Rasterizer.cpp:1098-1292).
//
// The warp-per-block mapping above spends ~100 warp instructions per updated block, most of them
// uniform bookkeeping replicated over 32 lanes, and keeps one block (one HBM round trip) in flight
// per warp.  Here up to 32 blocks of one primitive are processed at once, one per lane: each lane
// builds the whole 64-pixel block in registers (packed u16x2 arithmetic) and read-modify-writes its
// own 128 bytes.  Blocks of one primitive are distinct, so no two lanes touch the same block; order
// between primitives is kept because a warp finishes one primitive before it starts the next.
//
// The 12 iterated add chains still have to be stepped exactly as the reference does (y chain,
// then x chain restarted at every row start).  Lanes 0-11 each own one chain (one FADD advances
// all 12) and publish the value at every block position of the chunk through shared memory;
// afterwards lane j picks up the 12 values of its own block.
struct BlockWork {
  uint32_t blk;    // linear block index
  uint32_t hiz;    // HiZ read for the candidate test
};

// pack16 without the NaN guard: valid when the depth plane is finite (then no chain value can be NaN)
__device__ __forceinline__ uint32_t pack16_finite(float f) {
  const int32_t v = ((int32_t)f2u(f)) >> 12;
  return (uint32_t)min(max(v, 0), 65535);
}

// 64 pixels of one block for one lane: depth rows, coverage, merge, HiZ.  Rasterizer.cpp:1241-1290
template <bool kFinite>
__device__ __forceinline__ void update_block_lane(const Target& T, const uint32_t blk, const bool merge, const uint2 mk,
                                                  const float* __restrict__ smd /* this lane's 8 depth chain values, stride 32 */,
                                                  const float dzdx, const float dzdy) {
  uint4* dp = reinterpret_cast<uint4*>(T.depth) + (size_t)blk * 8u;
#if ORZ_V2_PRELOAD
  uint4 old[8];  // all eight rows are requested before any arithmetic: one HBM round trip per block
#pragma unroll
  for (int y = 0; y < 8; ++y) old[y] = merge ? dp[y] : make_uint4(0u, 0u, 0u, 0u);  // Rasterizer.cpp:1271-1278
#endif

  uint32_t r0[2][4], r4[2][4], r8[2][4];  // rows 0/1, 4/5, 8/9 as u16x2 words (pixels 2i, 2i+1)
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    float d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = smd[(4 * rr + k) * 32];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = d[(2 * i) & 3], b = d[(2 * i + 1) & 3];
      if (i >= 2) { a = ORZ_FMA(dzdx, 0.5f, a); b = ORZ_FMA(dzdx, 0.5f, b); }  // depth1, Rasterizer.cpp:1243
      const float a8 = dzdy + a, b8 = dzdy + b;                                // depth8/9, :1244-1245
      r0[rr][i] = kFinite ? pack16_finite(a) | (pack16_finite(b) << 16) : pack16(a) | (pack16(b) << 16);
      r8[rr][i] = kFinite ? pack16_finite(a8) | (pack16_finite(b8) << 16) : pack16(a8) | (pack16(b8) << 16);
      r4[rr][i] = avg_u16x2(r0[rr][i], r8[rr][i]);                             // :1252
    }
  }
  uint32_t mnAcc = 0xffffffffu;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int y = 2 * k + rr;
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        w[i] = k == 0 ? r0[rr][i] : k == 2 ? r4[rr][i] : k == 1 ? avg_u16x2(r0[rr][i], r4[rr][i]) : avg_u16x2(r4[rr][i], r8[rr][i]);  // :1253-1254
      // coverage of row y: pixel px <-> bit 8 px + ky (Rasterizer.cpp:1257-1268)
      const int ky = (rr ? 0 : 4) + k;
      const uint32_t lo = ((mk.x >> ky) & 0x01010101u) * 0xffu, hi = ((mk.y >> ky) & 0x01010101u) * 0xffu;
#if ORZ_V2_PRELOAD
      const uint4 o = old[y];
#else
      const uint4 o = merge ? dp[y] : make_uint4(0u, 0u, 0u, 0u);  // the line was prefetched when the block passed HiZ
#endif
      uint4 v;
      v.x = __vmaxu2(w[0] & __byte_perm(lo, 0u, 0x1100), o.x);
      v.y = __vmaxu2(w[1] & __byte_perm(lo, 0u, 0x3322), o.y);
      v.z = __vmaxu2(w[2] & __byte_perm(hi, 0u, 0x1100), o.z);
      v.w = __vmaxu2(w[3] & __byte_perm(hi, 0u, 0x3322), o.w);
      dp[y] = v;
      mnAcc = __vminu2(mnAcc, __vminu2(__vminu2(v.x, v.y), __vminu2(v.z, v.w)));
    }
  T.hiz[blk] = (uint16_t)min(mnAcc & 0xffffu, mnAcc >> 16);  // Rasterizer.cpp:1287-1290
}

// coverage + update for the (up to 32) blocks whose chain values sit in `sm`; `pass` = HiZ candidate
__device__ __forceinline__ void process_chunk_lanes(const Target& T, const uint2* __restrict__ lut, const float* __restrict__ sm,
                                                    const int lane, const bool pass, const uint32_t blk, const uint32_t h,
                                                    const uint32_t mode, const uint32_t slope01, const uint32_t slope23,
                                                    const float dzdx, const float dzdy, const bool finitePlane) {
  bool upd = false;
  uint2 mk = make_uint2(0u, 0u);
  if (pass) {
    const float o0 = sm[0 * 32 + lane], o1 = sm[1 * 32 + lane], o2 = sm[2 * 32 + lane], o3 = sm[3 * 32 + lane];
    const uint32_t s0 = slope01 & 0xffffu, s1 = slope01 >> 16, s2 = slope23 & 0xffffu, s3 = slope23 >> 16;
    if (mode == kConvex) {  // Rasterizer.cpp:1155-1187
      if (!(o0 >= 63.0f || o1 >= 63.0f || o2 >= 63.0f || o3 >= 63.0f)) {
        const uint2 A = lut[s0 | (uint32_t)__float2int_rz(fmaxf(o0, 0.0f))], B = lut[s1 | (uint32_t)__float2int_rz(fmaxf(o1, 0.0f))];
        const uint2 C = lut[s2 | (uint32_t)__float2int_rz(fmaxf(o2, 0.0f))], D = lut[s3 | (uint32_t)__float2int_rz(fmaxf(o3, 0.0f))];
        mk.x = (A.x & B.x) & (C.x & D.x); mk.y = (A.y & B.y) & (C.y & D.y);
        upd = true;  // no empty-mask test on this path (Rasterizer.cpp:1186)
      }
    } else {  // Rasterizer.cpp:1188-1239
      const uint32_t q0 = o0 < 2147483648.0f ? (uint32_t)__float2int_rz(fminf(fmaxf(o0, 0.0f), 63.0f)) : 0u;
      const uint32_t q1 = o1 < 2147483648.0f ? (uint32_t)__float2int_rz(fminf(fmaxf(o1, 0.0f), 63.0f)) : 0u;
      const uint32_t q2 = o2 < 2147483648.0f ? (uint32_t)__float2int_rz(fminf(fmaxf(o2, 0.0f), 63.0f)) : 0u;
      const uint32_t q3 = o3 < 2147483648.0f ? (uint32_t)__float2int_rz(fminf(fmaxf(o3, 0.0f), 63.0f)) : 0u;
      const uint2 A = lut[s0 | q0], B = lut[s1 | q1], C = lut[s2 | q2], D = lut[s3 | q3];
      if (mode == kTriangle0) { mk.x = A.x & B.x & C.x; mk.y = A.y & B.y & C.y; }
      else if (mode == kTriangle1) { mk.x = A.x & C.x & D.x; mk.y = A.y & C.y & D.y; }
      else if (mode == kConcaveRight) { mk.x = (A.x | D.x) & (B.x & C.x); mk.y = (A.y | D.y) & (B.y & C.y); }
      else if (mode == kConcaveCenter) { mk.x = (A.x & B.x) | (C.x & D.x); mk.y = (A.y & B.y) | (C.y & D.y); }
      else { mk.x = (A.x & D.x) & (B.x | C.x); mk.y = (A.y & D.y) & (B.y | C.y); }
      upd = (mk.x | mk.y) != 0u;
    }
  }
  if (upd) {
    if (finitePlane) update_block_lane<true>(T, blk, h != 1u, mk, sm + 4 * 32 + lane, dzdx, dzdy);
    else update_block_lane<false>(T, blk, h != 1u, mk, sm + 4 * 32 + lane, dzdx, dzdy);
  }
}

template <int kStride>
__device__ __forceinline__ void raster_prim_blocks(const uint32_t* __restrict__ rec, const int lane, const uint32_t rowPhase,
                                                   const Target& T, const uint2* __restrict__ lut, float* __restrict__ sm) {
  const uint32_t w0 = rec[0], w1 = rec[1], w2 = rec[2];
  const uint32_t minX = w0 & 0xffffu, minY = w0 >> 16, W = w1 & 0xffffu, rangeY = w1 >> 16;
  const uint32_t maxZ = w2 & 0xffffu, mode = w2 >> 16;
  const uint32_t blocksX = T.blocksX;
  const uint32_t b0 = ((uint32_t)kStride + rowPhase - minY % (uint32_t)kStride) % (uint32_t)kStride;  // first row of mine
  if (b0 >= rangeY) return;
  const uint32_t nRows = (rangeY - b0 + (uint32_t)kStride - 1u) / (uint32_t)kStride;
  const float dzdx = u2f(rec[3]), dzdy = u2f(rec[4]);
  const uint32_t slope01 = rec[18], slope23 = rec[19];
  // a finite depth plane cannot produce NaN depths (sums of finite terms overflow to inf at worst)
  const bool finitePlane = ((rec[3] & 0x7f800000u) != 0x7f800000u) && ((rec[4] & 0x7f800000u) != 0x7f800000u) &&
                           ((rec[5] & 0x7f800000u) != 0x7f800000u);

  // chain lane c: 0-3 edge offsets, 4-11 the eight depth lanes (Rasterizer.cpp:1103-1112)
  float cur = 0.0f, incX = 0.0f, incY = 0.0f;
  if (lane < 4) { cur = u2f(rec[14 + lane]); incX = u2f(rec[6 + lane]); incY = u2f(rec[10 + lane]); }
  else if (lane < 12) {
    const int l = lane - 4;
    const float s = -0.5f + 1.0f / 16.0f;
    cur = ORZ_FMA(dzdx, s + 0.125f * (float)(l & 3), ORZ_FMA(dzdy, (l >> 2) ? s + 0.125f : s, u2f(rec[5])));
    incX = dzdx; incY = dzdy;
  }
  for (uint32_t i = 0; i < b0; ++i) cur = cur + incY;  // Rasterizer.cpp:1130-1131

  if (W <= 32u) {
    // several rows per chunk: lane -> (row r of the chunk, column c)
    const uint32_t magic = rec[20];  // ceil(1024 / W), exact for lane < 32 (stored by store_record)
    const uint32_t rpc = (32u * magic) >> 10;  // == 32 / W for every W <= 32
    const uint32_t r = ((uint32_t)lane * magic) >> 10, c = (uint32_t)lane - r * W;
    // HiZ of the next chunk is requested while the current one is processed
    const uint32_t blk0 = (minY + b0 + (uint32_t)kStride * r) * blocksX + minX + c;
    const uint32_t blkStep = (uint32_t)kStride * rpc * blocksX;
    uint32_t hNext = r < min(rpc, nRows) ? (uint32_t)T.hiz[blk0] : 0xffffu;
    for (uint32_t row0 = 0; row0 < nRows; row0 += rpc) {
      const uint32_t rowsHere = min(rpc, nRows - row0);
      const uint32_t blk = blk0 + (row0 / rpc) * blkStep;
      const uint32_t h = hNext;
      if (row0 + rpc < nRows) hNext = r < min(rpc, nRows - row0 - rpc) ? (uint32_t)T.hiz[blk + blkStep] : 0xffffu;
      const bool pass = h < maxZ;  // Rasterizer.cpp:1148-1152
      if (!__any_sync(kFull, pass)) {
        for (uint32_t i = 0; i < rowsHere * (uint32_t)kStride; ++i) cur = cur + incY;
        continue;
      }
#if !ORZ_V2_PRELOAD
      if (pass && h != 1u) prefetch_l1(T.depth + (size_t)blk * 64u);  // one 128 B line = the whole block
#endif
      uint32_t j = 0;
      for (uint32_t rr = 0; rr < rowsHere; ++rr) {
        float run = cur;  // x chain restarts at the row start (Rasterizer.cpp:1136-1137)
#pragma unroll 4
        for (uint32_t bx = 0; bx < W; ++bx, ++j) {
          if (lane < 12) sm[lane * 32 + j] = run;
          run = incX + run;  // Rasterizer.cpp:1145-1146
        }
#pragma unroll
        for (int k = 0; k < kStride; ++k) cur = cur + incY;
      }
      __syncwarp();
      process_chunk_lanes(T, lut, sm, lane, pass, blk, h, mode, slope01, slope23, dzdx, dzdy, finitePlane);
      __syncwarp();
    }
  } else {
    for (uint32_t row = 0; row < nRows; ++row) {
      const uint32_t by = b0 + (uint32_t)kStride * row;
      const uint32_t rowBlk = (minY + by) * blocksX + minX;
      float run = cur;
      for (uint32_t s0 = 0; s0 < W; s0 += 32u) {
        const uint32_t m = min(32u, W - s0);
        const uint32_t blk = rowBlk + s0 + (uint32_t)lane;
        const uint32_t h = (uint32_t)lane < m ? (uint32_t)T.hiz[blk] : 0xffffu;
        const bool pass = h < maxZ;
        if (!__any_sync(kFull, pass)) {
          if (s0 + 32u < W) for (uint32_t i = 0; i < 32u; ++i) run = incX + run;
          continue;
        }
#if !ORZ_V2_PRELOAD
        if (pass && h != 1u) prefetch_l1(T.depth + (size_t)blk * 64u);
#endif
#pragma unroll 4
        for (uint32_t j = 0; j < m; ++j) {
          if (lane < 12) sm[lane * 32 + j] = run;
          run = incX + run;
        }
        __syncwarp();
        process_chunk_lanes(T, lut, sm, lane, pass, blk, h, mode, slope01, slope23, dzdx, dzdy, finitePlane);
        __syncwarp();
      }
#pragma unroll
      for (int k = 0; k < kStride; ++k) cur = cur + incY;
    }
  }
  __syncwarp();  // order this primitive's depth/HiZ stores before the next primitive's loads (other lanes)
}

// ---------------------------------------------------------------------------------------------
// query2D, Rasterizer.cpp:283-349.  Depth of cleared blocks is zero (fresh state), so no HiZ==1
// special case is needed on the read side.
__device__ __forceinline__ bool block_fine_test(const uint16_t* __restrict__ depth, uint32_t b, uint32_t maxZ, int sX, int eX,
                                                int sY, int eY) {
  const uint4* rows = reinterpret_cast<const uint4*>(depth + (size_t)b * 64u);
  const uint32_t mz = maxZ | (maxZ << 16);
  uint32_t sel[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    sel[i] = ((2 * i >= sX && 2 * i <= eX) ? 0x0000ffffu : 0u) | ((2 * i + 1 >= sX && 2 * i + 1 <= eX) ? 0xffff0000u : 0u);
  uint32_t any = 0;
  for (int y = sY; y <= eY; ++y) {
    const uint4 r = rows[y];  // visible where depth < maxZ (Rasterizer.cpp:335-339)
    any |= (__vcmpltu2(r.x, mz) & sel[0]) | (__vcmpltu2(r.y, mz) & sel[1]) | (__vcmpltu2(r.z, mz) & sel[2]) |
           (__vcmpltu2(r.w, mz) & sel[3]);
  }
  return any != 0u;
}

__device__ __forceinline__ bool query_block(const Target& T, uint32_t bx, uint32_t by, uint32_t minX, uint32_t maxX,
                                            uint32_t minY, uint32_t maxY, uint32_t maxZ) {
  const uint32_t b = by * T.blocksX + bx;
  const uint32_t h = T.hiz[b];
  if (maxZ <= h) return false;  // Rasterizer.cpp:310
  if (h == 1u) return true;     // cleared block: depth reads as 0 < maxZ (fresh state), stored bytes are not valid yet
  const int sX = max((int)minX - (int)(8u * bx), 0), eX = min((int)maxX - (int)(8u * bx), 7);
  const int sY = max((int)minY - (int)(8u * by), 0), eY = min((int)maxY - (int)(8u * by), 7);
  if (sX == 0 && eX == 7 && sY == 0 && eY == 7) return true;  // Rasterizer.cpp:319-325
  return block_fine_test(T.depth, b, maxZ, sX, eX, sY, eY);
}

// one thread walks the whole rectangle (occludee queries)
__device__ bool query2d_serial(const Target& T, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ) {
  const uint32_t bx0 = minX >> 3, bx1 = maxX >> 3, by0 = minY >> 3, by1 = maxY >> 3;
  for (uint32_t by = by0; by <= by1; ++by)
    for (uint32_t bx = bx0; bx <= bx1; ++bx)
      if (query_block(T, bx, by, minX, maxX, minY, maxY, maxZ)) return true;
  return false;
}

// One box per lane, whole warp converged: small rectangles are walked by their own lane, large ones
// (which would leave 31 lanes idle for hundreds of iterations) are taken one at a time by the whole
// warp, 32 blocks per step with coalesced HiZ reads.  query2D is an OR over blocks, so the visiting
// order does not matter.  Returns this lane's visibility.
__device__ __forceinline__ bool query2d_warp(const Target& T, const BoxFront& f, const int lane) {
  bool vis = false, big = false;
  if (f.status == kBoxRect) {
    const uint32_t nb = ((f.maxX >> 3) - (f.minX >> 3) + 1u) * ((f.maxY >> 3) - (f.minY >> 3) + 1u);
    if (nb <= 6u) vis = query2d_serial(T, f.minX, f.maxX, f.minY, f.maxY, f.maxZ);
    else big = true;
  }
  uint32_t pending = __ballot_sync(kFull, big);
  while (pending) {
    const int src = __ffs((int)pending) - 1;
    pending &= pending - 1u;
    const uint32_t minX = __shfl_sync(kFull, f.minX, src), maxX = __shfl_sync(kFull, f.maxX, src);
    const uint32_t minY = __shfl_sync(kFull, f.minY, src), maxY = __shfl_sync(kFull, f.maxY, src);
    const uint32_t maxZ = __shfl_sync(kFull, f.maxZ, src);
    const uint32_t bx0 = minX >> 3, by0 = minY >> 3;
    const uint32_t cols = (maxX >> 3) - bx0 + 1u, n = cols * ((maxY >> 3) - by0 + 1u);
    const uint32_t magic = (65536u + cols - 1u) / cols;  // i / cols ~ (i * magic) >> 16, at most one too large (i < 65536)
    bool found = false;
    // 128 blocks per step: the four HiZ reads of a lane are in flight together (a fully occluded
    // large box is a chain of dependent L2 round trips otherwise); fine tests only where needed
    for (uint32_t base = 0; base < n && !found; base += 128u) {
      uint32_t h[4], bxs[4], bys[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t i = base + (uint32_t)u * 32u + (uint32_t)lane;
        h[u] = 0xffffu;  // maxZ <= 0xffff: skipped
        bxs[u] = bys[u] = 0u;
        if (i < n) {
          uint32_t ry = n <= 65536u ? (i * magic) >> 16 : i / cols;
          if (ry * cols > i) --ry;
          bxs[u] = bx0 + (i - ry * cols); bys[u] = by0 + ry;
          h[u] = T.hiz[bys[u] * T.blocksX + bxs[u]];
        }
      }
      bool hit = false;
      uint32_t fine = 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (maxZ > h[u]) {  // Rasterizer.cpp:310
          const int sX = max((int)minX - (int)(8u * bxs[u]), 0), eX = min((int)maxX - (int)(8u * bxs[u]), 7);
          const int sY = max((int)minY - (int)(8u * bys[u]), 0), eY = min((int)maxY - (int)(8u * bys[u]), 7);
          if (h[u] == 1u || (sX == 0 && eX == 7 && sY == 0 && eY == 7)) hit = true;  // cleared block / Rasterizer.cpp:319-325
          else fine |= 1u << u;
        }
      if (__any_sync(kFull, hit)) { found = true; break; }
      if (__any_sync(kFull, fine != 0u)) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if ((fine >> u) & 1u) {
            const int sX = max((int)minX - (int)(8u * bxs[u]), 0), eX = min((int)maxX - (int)(8u * bxs[u]), 7);
            const int sY = max((int)minY - (int)(8u * bys[u]), 0), eY = min((int)maxY - (int)(8u * bys[u]), 7);
            hit = hit || block_fine_test(T.depth, bys[u] * T.blocksX + bxs[u], maxZ, sX, eX, sY, eY);
          }
        if (__any_sync(kFull, hit)) { found = true; break; }
      }
    }
    if (lane == src) vis = found;
  }
  return vis;
}

// all threads of a group share one rectangle (occluder gate): every warp takes 32 blocks per
// step; `flag` is a shared-memory word a finder sets so the other warps can stop early (read and
// written with atomics only -- the value is consumed after the group barrier that follows)
__device__ __forceinline__ void query2d_coop(const Target& T, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY,
                                             uint32_t maxZ, uint32_t tid, uint32_t nThreads, uint32_t* flag) {
  const uint32_t lane = tid & 31u;
  const uint32_t bx0 = minX >> 3, by0 = minY >> 3;
  const uint32_t cols = (maxX >> 3) - bx0 + 1u, rows = (maxY >> 3) - by0 + 1u;
  const uint32_t n = cols * rows;
  for (uint32_t base = tid - lane; base < n; base += nThreads) {
    uint32_t stop = 0u;
    if (lane == 0) stop = atomicOr(flag, 0u);
    if (__shfl_sync(kFull, stop, 0)) return;
    const uint32_t i = base + lane;
    bool hit = false;
    if (i < n) {
      const uint32_t ry = i / cols, rx = i - ry * cols;
      hit = query_block(T, bx0 + rx, by0 + ry, minX, maxX, minY, maxY, maxZ);
    }
    if (__any_sync(kFull, hit)) {
      if (lane == 0) atomicExch(flag, 1u);
      return;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Setup of up to 32 * nWarps quads starting at quad `q0`, compacted in order into `recs`
// (warp w writes slots [32 w, 32 w + count[w])).  Rasterizer.cpp:657-1086.
__device__ __forceinline__ void setup_chunk(const uint4* __restrict__ quads, uint32_t q0, uint32_t nq, bool clipped,
                                            const CallMatrix& cm, const RcpTable& rt, const Target& T, int warp, int lane,
                                            uint32_t* recs, uint32_t* counts) {
  const uint32_t qi = q0 + (uint32_t)warp * 32u + (uint32_t)lane;
  bool ok = false;
  Prim P;
  if (qi < nq) {
    const uint4 v = quads[qi];  // 128-bit coalesced load: the four packed vertices of this lane's quad
    const uint32_t word[4] = {v.x, v.y, v.z, v.w};
    ok = clipped ? setup_quad<true>(word, cm, rt, c_modeNibbles, (int32_t)T.blocksX, (int32_t)T.blocksY, P)
                 : setup_quad<false>(word, cm, rt, c_modeNibbles, (int32_t)T.blocksX, (int32_t)T.blocksY, P);
  }
  const uint32_t valid = __ballot_sync(kFull, ok);
  if (ok) store_record(recs + ((uint32_t)warp * 32u + (uint32_t)__popc(valid & ((1u << lane) - 1u))) * kRecStride, P);
  if (lane == 0) counts[warp] = (uint32_t)__popc(valid);
}

// ---------------------------------------------------------------------------------------------
// View-batch path: Main.cpp:181-206 for many independent views, three launches per batch:
//   k_prepare_views  per (view, occluder): everything that does not depend on the depth buffer --
//                    matrices, front-to-back order, query front half, per-call matrix
//   k_render_views   one CTA of GW warps per view at a time: clear, then gate -> setup -> traversal
//                    per occluder in order
//   k_query_views    one thread per (view, occludee box) on the finished buffers
constexpr int kFrontWords = 20;  // status, minX, maxX, minY, maxY, maxZ, CallMatrix (14 floats)

struct FrameParams {
  const uint4* quads;
  const OccMeta* occ;
  uint32_t nOcc;
  const float4* boxes;
  uint32_t nBoxes;
  const uint32_t* rcp;
  int rcpShift;
  const uint2* lut;
  uint32_t width, height, nViews, flags;
  const float* mvps;
  const uint32_t* orders;  // caller's order, or NULL: computed from camPos into orderBuf
  const float* camPos;
  uint32_t* orderBuf;      // nViews x nOcc
  ViewMatrices* vmBuf;     // nViews
  uint32_t* frontBuf;      // nViews x nOcc x kFrontWords
  uint16_t* depth;
  uint16_t* hiz;
  unsigned long long depthStride, hizStride;  // elements between consecutive views
  uint32_t* visBits;
  uint32_t* clipBits;
  uint32_t bitWords;
  uint8_t* gate;
  uint32_t* quadsSubmitted;
  uint32_t* viewCounter;
  uint32_t* viewCost;   // nViews: quads of the occluders that survive the frustum test (scheduling estimate)
  uint32_t* viewOrder;  // nViews: views sorted by descending cost (longest first), or NULL
  uint32_t viewBase, groupViews;  // this launch handles sorted ranks [viewBase, viewBase + groupViews)
  int exportDepth;      // 1: the caller reads depth back -> zero-fill blocks that stayed cleared
  uint32_t clusterK;    // cluster kernel: tiles per warp
  // cluster path: speculative setup output per view (k_setup_views)
  uint32_t* recBuf;     // [nViews][totalQuads][kRecStride] records, each occluder's at its quadOffset
  uint2* hdrBuf;        // [nViews][totalQuads] bounding boxes of the records
  uint4* recInfo;       // [nViews][nOcc][2]: {records, quadOffset, quadCount, -}, {block rectangle of all records, half open}
  uint32_t totalQuads;
};

__global__ void __launch_bounds__(128) k_prepare_views(const FrameParams p) {
  __shared__ ViewMatrices s_vm;
  const uint32_t view = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  const RcpTable rt{p.rcp, p.rcpShift};
  if (tid == 0) {  // setModelViewProjection, Rasterizer.cpp:76-105
    bake_view_matrices(p.mvps + 16 * (size_t)view, p.width, p.height, s_vm);
    p.vmBuf[view] = s_vm;
  }
  // front-to-back order (Main.cpp:185-190) when the caller did not supply one: rank sort on
  // dp(c - p, c - p) in the dpps 0x7f sum order, stable by index
  const uint32_t* order = p.orders ? p.orders + (size_t)view * p.nOcc : nullptr;
  if (!order) {
    uint32_t* mine = p.orderBuf + (size_t)view * p.nOcc;
    const float cx = p.camPos[3 * (size_t)view + 0], cy = p.camPos[3 * (size_t)view + 1], cz = p.camPos[3 * (size_t)view + 2];
    for (uint32_t i = tid; i < p.nOcc; i += NT) {
      const float* ci = p.occ[i].center;
      const float dxi = ci[0] - cx, dyi = ci[1] - cy, dzi = ci[2] - cz;
      const float ki = (dxi * dxi + dyi * dyi) + dzi * dzi;
      uint32_t rank = 0;
      for (uint32_t j = 0; j < p.nOcc; ++j) {
        const float* cj = p.occ[j].center;
        const float dxj = cj[0] - cx, dyj = cj[1] - cy, dzj = cj[2] - cz;
        const float kj = (dxj * dxj + dyj * dyj) + dzj * dzj;
        rank += (kj < ki || (kj == ki && j < i)) ? 1u : 0u;
      }
      mine[rank] = i;
    }
    order = mine;
  }
  __syncthreads();
  const bool useGate = (p.flags & ORZ_BATCH_NO_GATE) == 0u;
  uint32_t cost = 0;
  for (uint32_t slot = tid; slot < p.nOcc; slot += NT) {
    const OccMeta& om = p.occ[order[slot]];
    BoxFront f;
    if (useGate) {
      f = box_front_half(s_vm, om.boundsMin, om.boundsMax, p.width, p.height, rt);  // Rasterizer.cpp:123-273
    } else {
      f.status = kBoxNearClip; f.minX = f.maxX = f.minY = f.maxY = f.maxZ = 0;
    }
    CallMatrix cm;
    prepare_call(s_vm.baked, om.refMin, om.refMax, cm);  // Rasterizer.cpp:616-655
    uint32_t* out = p.frontBuf + ((size_t)view * p.nOcc + slot) * kFrontWords;
    out[0] = f.status; out[1] = f.minX; out[2] = f.maxX; out[3] = f.minY; out[4] = f.maxY; out[5] = f.maxZ;
#pragma unroll
    for (int k = 0; k < 4; ++k) { out[6 + k] = f2u(cm.rx[k]); out[10 + k] = f2u(cm.ry[k]); out[14 + k] = f2u(cm.rw[k]); }
    out[18] = f2u(cm.c0); out[19] = f2u(cm.c1);
    if (f.status != kBoxCulled) cost += om.quadCount;
  }
  // per-view work estimate for longest-first scheduling of the render kernel
  __shared__ uint32_t s_cost;
  if (tid == 0) s_cost = 0u;
  __syncthreads();
  cost = __reduce_add_sync(kFull, cost);
  if ((tid & 31u) == 0 && cost) atomicAdd(&s_cost, cost);
  __syncthreads();
  if (tid == 0) p.viewCost[view] = s_cost;
}

// views by descending cost, ties by index (rank sort; nViews is at most a few thousand per chunk)
__global__ void __launch_bounds__(256) k_sort_views(const uint32_t* __restrict__ cost, uint32_t n, uint32_t* __restrict__ order) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t ci = cost[i];
  uint32_t rank = 0;
  for (uint32_t j = 0; j < n; ++j) {
    const uint32_t cj = cost[j];
    rank += (cj > ci || (cj == ci && j < i)) ? 1u : 0u;
  }
  order[rank] = i;
}

template <int GW, int kTrav>
struct FrameSmem {
  static constexpr int kBufs = GW >= 16 ? 1 : 2;  // record buffers (double buffering saves one barrier per chunk)
  static constexpr size_t kBytes = (size_t)kBufs * GW * 32 * kRecStride * 4 + (kTrav == 2 ? (size_t)GW * 12 * 32 * 4 : 0);
};

// kTrav selects the traversal mapping: 1 = one warp per block (raster_prim), 2 = one lane per block
// (raster_prim_blocks; needs more registers, so it is compiled for fewer resident threads per SM)
template <int GW, int kTrav>
__global__ void __launch_bounds__(GW * 32, (kTrav == 2 ? ORZ_THREADS_PER_SM_V2 : ORZ_THREADS_PER_SM) / (GW * 32)) k_render_views(const FrameParams p) {
  constexpr uint32_t NT = GW * 32;
  // dynamic shared memory (may exceed the 48 KB static limit): [2][NT][21] records, then [GW][12][32] chain slots
  constexpr int kBufs = FrameSmem<GW, kTrav>::kBufs;
  extern __shared__ __align__(16) uint32_t s_dyn[];
  uint32_t (*s_recs)[NT * kRecStride] = reinterpret_cast<uint32_t (*)[NT * kRecStride]>(s_dyn);
  float* s_chain = reinterpret_cast<float*>(s_dyn + kBufs * NT * kRecStride);
  __shared__ uint32_t s_count[kBufs][GW];
  __shared__ uint32_t s_flag[3];
  __shared__ uint32_t s_view;

  const uint32_t tid = threadIdx.x;
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31u);
  const RcpTable rt{p.rcp, p.rcpShift};
  Target T;
  T.width = p.width; T.height = p.height; T.blocksX = p.width >> 3; T.blocksY = p.height >> 3;
  const uint32_t blocks = T.blocksX * T.blocksY;
  const bool useGate = (p.flags & ORZ_BATCH_NO_GATE) == 0u;
  const bool forceClip = (p.flags & ORZ_BATCH_FORCE_CLIPPED) != 0u;
  uint32_t buf = 0;

  for (;;) {
    if (tid == 0) { s_view = atomicAdd(p.viewCounter, 1u); s_flag[0] = s_flag[1] = s_flag[2] = 0u; }
    __syncthreads();
    if (s_view >= p.groupViews) break;
    const uint32_t view = p.viewOrder ? p.viewOrder[p.viewBase + s_view] : p.viewBase + s_view;
    T.depth = p.depth + (size_t)view * p.depthStride;
    T.hiz = p.hiz + (size_t)view * p.hizStride;
    const uint32_t* order = p.orders ? p.orders + (size_t)view * p.nOcc : p.orderBuf + (size_t)view * p.nOcc;
    const uint32_t* front = p.frontBuf + (size_t)view * p.nOcc * kFrontWords;

    // ---- clear (Rasterizer.cpp:107-121): HiZ := 1.  Depth is NOT touched here: a block whose
    // HiZ is 1 is overwritten by its first update (Rasterizer.cpp:1271) and reads as zero in
    // queries, so the zero fill of never-touched blocks is deferred to the end of the view and
    // every depth byte is written to HBM once instead of twice.
    for (uint32_t i = tid; i < blocks; i += NT) T.hiz[i] = 1;
    __syncthreads();

    uint32_t gateIdx = 0, quadsSubmitted = 0;
    for (uint32_t slot = 0; slot < p.nOcc; ++slot) {
      const uint32_t* fr = front + (size_t)slot * kFrontWords;
      const uint32_t status = fr[0];
      bool visible = false, clipped = false;
      if (status == kBoxNearClip) {
        visible = true;
        clipped = useGate ? true : forceClip;
      } else if (status == kBoxRect) {
        // ---- gate: query2D on the buffers as built so far (Main.cpp:195)
        uint32_t* flag = &s_flag[gateIdx % 3u];
        if (tid == 0) s_flag[(gateIdx + 1u) % 3u] = 0u;
        query2d_coop(T, fr[1], fr[2], fr[3], fr[4], fr[5], tid, NT, flag);
        __syncthreads();
        visible = *flag != 0u;  // after the barrier: plain read
        ++gateIdx;
      }
      if (p.gate && tid == 0) p.gate[(size_t)view * p.nOcc + slot] = (uint8_t)((visible ? 1 : 0) | (clipped && useGate ? 2 : 0));
      if (!visible) continue;

      // ---- rasterize<clipped>(occluder): setup chunk -> records -> traversal of my rows.
      // Records are double buffered: one barrier per chunk (after its setup) is enough, because
      // a warp can only start overwriting buffer b two barriers after the traversal that read it.
      const OccMeta& om = p.occ[order[slot]];
      const uint4* quads = p.quads + om.quadOffset;
      const uint32_t nq = om.quadCount;
      quadsSubmitted += nq;
      CallMatrix cm;
#pragma unroll
      for (int k = 0; k < 4; ++k) { cm.rx[k] = u2f(fr[6 + k]); cm.ry[k] = u2f(fr[10 + k]); cm.rw[k] = u2f(fr[14 + k]); }
      cm.c0 = u2f(fr[18]); cm.c1 = u2f(fr[19]);
      for (uint32_t q0 = 0; q0 < nq; q0 += NT) {
        setup_chunk(quads, q0, nq, clipped, cm, rt, T, warp, lane, s_recs[buf], s_count[buf]);
        __syncthreads();
#pragma unroll 1
        for (int w2 = 0; w2 < GW; ++w2) {
          const uint32_t cnt = s_count[buf][w2];
          for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t* rec = s_recs[buf] + ((uint32_t)w2 * 32u + i) * kRecStride;
            if (kTrav == 2 && blocks <= 65536u) raster_prim_blocks<GW>(rec, lane, (uint32_t)warp, T, p.lut, s_chain + warp * (12 * 32));
            else raster_prim<GW>(rec, lane, (uint32_t)warp, GW, T, p.lut);
          }
        }
        if (kBufs == 2) buf ^= 1u;
        else __syncthreads();  // single buffer: records are rewritten by the next chunk
      }
      __syncthreads();  // depth/HiZ of this occluder visible to the whole group before the next gate
    }
    if (p.quadsSubmitted && tid == 0) p.quadsSubmitted[view] = quadsSubmitted;
    if (p.exportDepth) {  // canonical depth for the caller: cleared blocks read as zero
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (uint32_t i = tid; i < blocks; i += NT)
        if (T.hiz[i] == 1) {
          uint4* d4 = reinterpret_cast<uint4*>(T.depth) + (size_t)i * 8u;
#pragma unroll
          for (int k = 0; k < 8; ++k) d4[k] = z;
        }
    }
    __syncthreads();  // s_view is rewritten by the next view
  }
}

// ---------------------------------------------------------------------------------------------
// Few views (BASELINE configs 1 and 2 are ONE view): the latency path.  A single view is a chain
// of dependent gate -> setup -> traversal steps (Main.cpp:192-206) and, inside one occluder,
// primitives stack on the same blocks (Castle, default camera: ~300 in-order updates of one
// block per frame), so what counts is the length of the dependency chain, not throughput.
//
//   k_setup_views           everything that does not depend on the depth buffer, at full width:
//                           one CTA per (occluder that survives the frustum, view) sets up its
//                           quads (Rasterizer.cpp:657-1086) and writes the valid primitives in
//                           order as records + 8-byte bounding-box headers (speculative: the
//                           gate may still reject the occluder)
//   k_raster_views_cluster  one thread-block CLUSTER of C CTAs x 16 warps per view, run as a
//                           DATAFLOW machine with no barrier in its main loop:
//     * the screen is cut into TILES of 8x4 blocks; tile t belongs to warp t mod (16 C) of the
//       cluster for the whole view, lane <-> block.  A tile is only ever read or written by its
//       owner (gate included): no cross-SM traffic on depth / HiZ, order preserved per block;
//     * every warp walks the occluders front to back at ITS OWN pace.  For a rectangle candidate
//       it tests the part of the rectangle that lies on its tiles (query2D, Rasterizer.cpp:283-349)
//       -- at that point it has applied every earlier visible occluder to those tiles, which is
//       all the test depends on -- and either raises the candidate's `visible` flag in the shared
//       memory of every CTA (DSMEM stores) or adds itself to the candidate's `done` count
//       (per-CTA count, forwarded to every CTA by the CTA's last warp).  Warps whose tiles do not
//       meet the rectangle are counted before the walk starts.  A candidate is visible as soon as
//       ONE warp says so, invisible when all have said no: fast warps run ahead and only the true
//       dependencies remain (sum over occluders of the slowest warp -> slowest warp's own total:
//       660 -> 176 primitive-tile steps on the Castle default view);
//     * TILE-MAJOR traversal: for each of its tiles a warp walks the occluder's primitives in
//       order with the tile's depth held in REGISTERS (one 8x8 block = 8 x uint4 per lane) and
//       its HiZ in a register + shared-memory mirror: a stacked primitive costs shared-memory and
//       ALU latency only; the L2 round trip (load at first touch, store at the end) is paid once
//       per (occluder, tile) instead of once per (primitive, block);
//     * the edge-mask table (32 KB) lives in shared memory; records are gathered from L2 into a
//       per-warp staging area 32 at a time with all loads in flight together.
#ifndef ORZ_CLUSTER_GW
#define ORZ_CLUSTER_GW 16
#endif
#ifndef ORZ_CLUSTER_LUT_SMEM
#define ORZ_CLUSTER_LUT_SMEM 1  // edge-mask table staged in shared memory (0: read through L1)
#endif
#ifndef ORZ_CLUSTER_CTAS_PER_SM
#define ORZ_CLUSTER_CTAS_PER_SM 0  // > 0: compile with __launch_bounds__(threads, this) instead of the register cap
#endif
#ifndef ORZ_CLUSTER_REGS
#define ORZ_CLUSTER_REGS 96  // 16 warps x 96 registers leave room for one CTA of the query kernel on the same SM
#endif
constexpr int kClusterGW = ORZ_CLUSTER_GW;  // warps per CTA of the cluster kernel; registers per thread capped so that they fit one SM
constexpr uint32_t kTileW = 8, kTileH = 4;   // blocks per tile: lane = 8 * (row in tile) + column in tile
constexpr uint32_t kChainStride = 33;        // words between two chains' slots: the publishing lanes (chain, tile row) hit 32 different banks
constexpr uint32_t kStageCap = 32;           // records a warp stages at a time
constexpr uint32_t kClusterMaxOcc = 2048;    // occluders per scene the cluster path accepts (shared-memory decision arrays)
constexpr uint32_t kHeadWords = 6;           // status, minX, maxX, minY, maxY, maxZ of kFrontWords

struct ClusterSmem {
  static constexpr uint32_t kLutWords = ORZ_CLUSTER_LUT_SMEM ? 4096 * 2 : 0;
  static constexpr uint32_t kStageWords = kClusterGW * kStageCap * kRecStride;
  static constexpr uint32_t kIdxWords = kClusterGW * kStageCap;
  static constexpr uint32_t kChainWords = kClusterGW * 12 * kChainStride;
  static constexpr uint32_t kFixedWords = kLutWords + kStageWords + kIdxWords + kChainWords;
  // + [nOcc][6] gate heads, 3 x [nOcc] decision words, [GW][K][32] u16 HiZ mirror
  static size_t bytes(uint32_t tilesPerWarp, uint32_t nOcc) {
    return (size_t)(kFixedWords + nOcc * (kHeadWords + 3u)) * 4 + (size_t)kClusterGW * tilesPerWarp * 32 * 2;
  }
};

// ---- speculative setup of every occluder that survives the frustum, Rasterizer.cpp:657-1086
__global__ void __launch_bounds__(256) k_setup_views(const FrameParams p) {
  __shared__ uint32_t s_cnt[8];
  __shared__ uint32_t s_box[4];
  const uint32_t slot = blockIdx.x, view = blockIdx.y, tid = threadIdx.x;
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31u);
  const uint32_t* fr = p.frontBuf + ((size_t)view * p.nOcc + slot) * kFrontWords;
  const uint32_t status = fr[0];
  if (status == kBoxCulled) {
    if (tid == 0) p.recInfo[((size_t)view * p.nOcc + slot) * 2u] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const bool useGate = (p.flags & ORZ_BATCH_NO_GATE) == 0u;
  const bool clipped = status == kBoxNearClip ? (useGate ? true : (p.flags & ORZ_BATCH_FORCE_CLIPPED) != 0u) : false;
  const uint32_t* order = p.orders ? p.orders + (size_t)view * p.nOcc : p.orderBuf + (size_t)view * p.nOcc;
  const OccMeta& om = p.occ[order[slot]];
  const RcpTable rt{p.rcp, p.rcpShift};
  CallMatrix cm;
#pragma unroll
  for (int k = 0; k < 4; ++k) { cm.rx[k] = u2f(fr[6 + k]); cm.ry[k] = u2f(fr[10 + k]); cm.rw[k] = u2f(fr[14 + k]); }
  cm.c0 = u2f(fr[18]); cm.c1 = u2f(fr[19]);
  const int32_t blocksX = (int32_t)(p.width >> 3), blocksY = (int32_t)(p.height >> 3);
  const size_t recBase = (size_t)view * p.totalQuads + om.quadOffset;  // records of this (view, occluder) start here
  uint32_t* recs = p.recBuf + recBase * kRecStride;
  uint2* hdrs = p.hdrBuf + recBase;
  if (tid < 4) s_box[tid] = tid < 2 ? 0xffffffffu : 0u;
  uint32_t written = 0;
  uint32_t bx0 = 0xffffffffu, by0 = 0xffffffffu, bx1 = 0u, by1 = 0u;
  for (uint32_t q0 = 0; q0 < om.quadCount; q0 += 256u) {
    const uint32_t qi = q0 + tid;
    bool ok = false;
    Prim P;
    if (qi < om.quadCount) {
      const uint4 v = p.quads[om.quadOffset + qi];  // 128-bit coalesced load: the four packed vertices of this lane's quad
      const uint32_t word[4] = {v.x, v.y, v.z, v.w};
      ok = clipped ? setup_quad<true>(word, cm, rt, c_modeNibbles, blocksX, blocksY, P)
                   : setup_quad<false>(word, cm, rt, c_modeNibbles, blocksX, blocksY, P);
    }
    const uint32_t valid = __ballot_sync(kFull, ok);
    __syncthreads();  // s_cnt of the previous chunk has been read
    if (lane == 0) s_cnt[warp] = (uint32_t)__popc(valid);
    __syncthreads();
    uint32_t base = written, total = 0;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) { const uint32_t c = s_cnt[w2]; base += w2 < warp ? c : 0u; total += c; }
    if (ok) {  // in order: binning by prefix-sum compaction
      const uint32_t at = base + (uint32_t)__popc(valid & ((1u << lane) - 1u));
      store_record(recs + (size_t)at * kRecStride, P);
      hdrs[at] = make_uint2((uint32_t)P.minX | ((uint32_t)P.minY << 16), (uint32_t)P.rangeX | ((uint32_t)P.rangeY << 16));
      bx0 = min(bx0, (uint32_t)P.minX); by0 = min(by0, (uint32_t)P.minY);
      bx1 = max(bx1, (uint32_t)(P.minX + P.rangeX)); by1 = max(by1, (uint32_t)(P.minY + P.rangeY));
    }
    written += total;
  }
  bx0 = __reduce_min_sync(kFull, bx0); by0 = __reduce_min_sync(kFull, by0);
  bx1 = __reduce_max_sync(kFull, bx1); by1 = __reduce_max_sync(kFull, by1);
  __syncthreads();
  if (lane == 0) { atomicMin(&s_box[0], bx0); atomicMin(&s_box[1], by0); atomicMax(&s_box[2], bx1); atomicMax(&s_box[3], by1); }
  __syncthreads();
  if (tid == 0) {
    p.recInfo[((size_t)view * p.nOcc + slot) * 2u] = make_uint4(written, om.quadOffset, om.quadCount, 0u);
    // block rectangle that holds every primitive of the occluder, half open (lo > hi when there is none)
    p.recInfo[((size_t)view * p.nOcc + slot) * 2u + 1u] = make_uint4(s_box[0], s_box[1], s_box[2], s_box[3]);
  }
}

// Decision words of the cluster kernel are read and written concurrently by design (monotonic
// flags / counters): strong relaxed accesses at cluster scope, which the PTX memory model allows
// to race (no data is published through them, only the decision itself).  compute-sanitizer's
// racecheck still lists exactly these two accesses (it only exempts atomics); polling with
// atomics instead was tried and starves the remote updates it is waiting for -- the kernel hangs.
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag_remote(uint32_t* localPtr, uint32_t ctaRank, uint32_t val) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(localPtr)), "r"(ctaRank));
  asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(val) : "memory");
}

// one block of query2D (Rasterizer.cpp:305-343) with the block's HiZ already at hand
__device__ __forceinline__ bool query_block_h(const Target& T, uint32_t bx, uint32_t by, uint32_t h, uint32_t minX, uint32_t maxX,
                                              uint32_t minY, uint32_t maxY, uint32_t maxZ) {
  if (maxZ <= h) return false;  // Rasterizer.cpp:310
  if (h == 1u) return true;     // cleared block: depth reads as 0 < maxZ (fresh state)
  const int sX = max((int)minX - (int)(8u * bx), 0), eX = min((int)maxX - (int)(8u * bx), 7);
  const int sY = max((int)minY - (int)(8u * by), 0), eY = min((int)maxY - (int)(8u * by), 7);
  if (sX == 0 && eX == 7 && sY == 0 && eY == 7) return true;  // Rasterizer.cpp:319-325
  return block_fine_test(T.depth, by * T.blocksX + bx, maxZ, sX, eX, sY, eY);
}

// One iterated chain for one tile row: nyCommon + nyExtra y steps, nPre x steps up to tile column
// cA, then the values at tile columns [cA, cB] go to out[c].  Trip counts are warp uniform except
// nyExtra (0-3, the row inside the tile); every add is the reference's own (same operands, same
// order), only lanes differ in what they own.
__device__ __forceinline__ void step_chain(float cur, const float incX, const float incY, const uint32_t nyCommon, const uint32_t nyExtra,
                                           const uint32_t nPre, const uint32_t cA, const uint32_t cB, const bool active, float* out) {
#pragma unroll kChainUnroll
  for (uint32_t i = 0; i < nyCommon; ++i) cur = cur + incY;  // Rasterizer.cpp:1130-1131
#pragma unroll
  for (uint32_t i = 0; i < kTileH - 1u; ++i) cur = i < nyExtra ? cur + incY : cur;
#pragma unroll kChainUnroll
  for (uint32_t i = 0; i < nPre; ++i) cur = incX + cur;      // Rasterizer.cpp:1145-1146
  for (uint32_t c = cA; c <= cB; ++c) {
    if (active) out[c] = cur;
    cur = incX + cur;
  }
}

// One primitive on the tile a warp has open (Rasterizer.cpp:1098-1292 restricted to the tile's
// blocks).  d[8] / h are the lane's block and its HiZ, kept in registers between primitives.
__device__ __forceinline__ void tile_prim(const uint32_t* __restrict__ rec, const int lane, const uint32_t x0, const uint32_t y0,
                                          const uint32_t x1, const uint32_t y1, const uint2* __restrict__ lut, float* __restrict__ sm,
                                          uint4 (&d)[8], uint32_t& h, bool& dirty) {
  const uint32_t w0 = rec[0], w1 = rec[1], w2 = rec[2];
  const uint32_t minX = w0 & 0xffffu, minY = w0 >> 16, maxZ = w2 & 0xffffu, mode = w2 >> 16;
  const uint32_t xa = max(minX, x0), xb = min(minX + (w1 & 0xffffu), x1), ya = max(minY, y0), yb = min(minY + (w1 >> 16), y1);
  const uint32_t bx = x0 + ((uint32_t)lane & 7u), by = y0 + ((uint32_t)lane >> 3);
  const bool pass = bx >= xa && bx < xb && by >= ya && by < yb && h < maxZ;  // Rasterizer.cpp:1148-1152
  const uint32_t passMask = __ballot_sync(kFull, pass);
  if (!passMask) return;  // the whole tile is behind its HiZ: no chain has to be stepped at all

  // ---- the iterated add chains, stepped exactly as the reference does: y chain from the
  // primitive's first row (Rasterizer.cpp:1130), x chain restarted at every row start (:1136,
  // :1145).  One lane per (chain, tile row): first the 4 edge offsets x 4 rows (16 lanes); the
  // 8 depth chains x 4 rows (32 lanes) only when some block is really covered.
  const float dzdx = u2f(rec[3]), dzdy = u2f(rec[4]);
  const uint32_t rFirst = ya - y0;
  {
    const uint32_t rLast = (31u - (uint32_t)__clz((int)passMask)) >> 3;
    const uint32_t cols = (passMask | (passMask >> 8) | (passMask >> 16) | (passMask >> 24)) & 0xffu;
    const uint32_t cA = (uint32_t)__ffs((int)cols) - 1u, cB = 31u - (uint32_t)__clz((int)cols);
    const uint32_t r = (uint32_t)lane >> 2, e = (uint32_t)lane & 3u;
    const bool active = lane < 16 && r >= rFirst && r <= rLast;
    float cur = 0.0f, incX = 0.0f, incY = 0.0f;
    if (active) { cur = u2f(rec[14 + e]); incX = u2f(rec[6 + e]); incY = u2f(rec[10 + e]); }
    step_chain(cur, incX, incY, ya - minY, r - rFirst, x0 + cA - minX, cA, cB, active, sm + e * kChainStride + r * 8u);
  }
  __syncwarp();

  // ---- coverage (Rasterizer.cpp:1155-1239)
  bool upd = false;
  uint2 mk = make_uint2(0u, 0u);
  if (pass) {
    const float o0 = sm[0 * kChainStride + lane], o1 = sm[1 * kChainStride + lane], o2 = sm[2 * kChainStride + lane], o3 = sm[3 * kChainStride + lane];
    const uint32_t slope01 = rec[18], slope23 = rec[19];
    const uint32_t s0 = slope01 & 0xffffu, s1 = slope01 >> 16, s2 = slope23 & 0xffffu, s3 = slope23 >> 16;
    if (mode == kConvex) {
      if (!(o0 >= 63.0f || o1 >= 63.0f || o2 >= 63.0f || o3 >= 63.0f)) {
        const uint2 A = lut[s0 | (uint32_t)__float2int_rz(fmaxf(o0, 0.0f))], B = lut[s1 | (uint32_t)__float2int_rz(fmaxf(o1, 0.0f))];
        const uint2 C2 = lut[s2 | (uint32_t)__float2int_rz(fmaxf(o2, 0.0f))], D = lut[s3 | (uint32_t)__float2int_rz(fmaxf(o3, 0.0f))];
        mk.x = (A.x & B.x) & (C2.x & D.x); mk.y = (A.y & B.y) & (C2.y & D.y);
        upd = true;  // no empty-mask test on this path (Rasterizer.cpp:1186)
      }
    } else {
      const uint32_t q0 = o0 < 2147483648.0f ? (uint32_t)__float2int_rz(fminf(fmaxf(o0, 0.0f), 63.0f)) : 0u;
      const uint32_t q1 = o1 < 2147483648.0f ? (uint32_t)__float2int_rz(fminf(fmaxf(o1, 0.0f), 63.0f)) : 0u;
      const uint32_t q2 = o2 < 2147483648.0f ? (uint32_t)__float2int_rz(fminf(fmaxf(o2, 0.0f), 63.0f)) : 0u;
      const uint32_t q3 = o3 < 2147483648.0f ? (uint32_t)__float2int_rz(fminf(fmaxf(o3, 0.0f), 63.0f)) : 0u;
      const uint2 A = lut[s0 | q0], B = lut[s1 | q1], C2 = lut[s2 | q2], D = lut[s3 | q3];
      if (mode == kTriangle0) { mk.x = A.x & B.x & C2.x; mk.y = A.y & B.y & C2.y; }
      else if (mode == kTriangle1) { mk.x = A.x & C2.x & D.x; mk.y = A.y & C2.y & D.y; }
      else if (mode == kConcaveRight) { mk.x = (A.x | D.x) & (B.x & C2.x); mk.y = (A.y | D.y) & (B.y & C2.y); }
      else if (mode == kConcaveCenter) { mk.x = (A.x & B.x) | (C2.x & D.x); mk.y = (A.y & B.y) | (C2.y & D.y); }
      else { mk.x = (A.x & D.x) & (B.x | C2.x); mk.y = (A.y & D.y) & (B.y | C2.y); }
      upd = (mk.x | mk.y) != 0u;
    }
  }
  const uint32_t updMask = __ballot_sync(kFull, upd);
  __syncwarp();  // orders this primitive's reads of the edge slots before the next primitive's writes (free: the warp is converged)
  if (!updMask) return;
  {  // the eight depth lanes (Rasterizer.cpp:1103-1112) at the covered blocks
    const uint32_t rLast = (31u - (uint32_t)__clz((int)updMask)) >> 3, rLo = ((uint32_t)__ffs((int)updMask) - 1u) >> 3;
    const uint32_t cols = (updMask | (updMask >> 8) | (updMask >> 16) | (updMask >> 24)) & 0xffu;
    const uint32_t cA = (uint32_t)__ffs((int)cols) - 1u, cB = 31u - (uint32_t)__clz((int)cols);
    const uint32_t r = (uint32_t)lane >> 3, l = (uint32_t)lane & 7u;
    const bool active = r >= rLo && r <= rLast;
    const float s = -0.5f + 1.0f / 16.0f;
    const float cur = ORZ_FMA(dzdx, s + 0.125f * (float)(l & 3u), ORZ_FMA(dzdy, (l >> 2) ? s + 0.125f : s, u2f(rec[5])));
    step_chain(cur, dzdx, dzdy, y0 + rLo - minY, r - rLo, x0 + cA - minX, cA, cB, active, sm + (4u + l) * kChainStride + r * 8u);
  }
  __syncwarp();
  // ---- depth rows, merge into the registers, HiZ (Rasterizer.cpp:1241-1290)
  if (upd) {
    const float* smd = sm + 4 * kChainStride + lane;
    const uint32_t keep = h != 1u ? 0xffffffffu : 0u;  // a cleared block is overwritten (:1271-1278)
    uint32_t r0[2][4], r4[2][4], r8[2][4];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) dv[k] = smd[(4 * rr + k) * kChainStride];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = dv[(2 * i) & 3], b = dv[(2 * i + 1) & 3];
        if (i >= 2) { a = ORZ_FMA(dzdx, 0.5f, a); b = ORZ_FMA(dzdx, 0.5f, b); }  // depth1, :1243
        const float a8 = dzdy + a, b8 = dzdy + b;                                // depth8/9, :1244-1245
        r0[rr][i] = pack16(a) | (pack16(b) << 16);  // (a run-time "finite plane" shortcut for the NaN guard was measured slower)
        r8[rr][i] = pack16(a8) | (pack16(b8) << 16);
        r4[rr][i] = avg_u16x2(r0[rr][i], r8[rr][i]);                             // :1252
      }
    }
    uint32_t mnAcc = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int y = 2 * k + rr;
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w[i] = k == 0 ? r0[rr][i] : k == 2 ? r4[rr][i] : k == 1 ? avg_u16x2(r0[rr][i], r4[rr][i]) : avg_u16x2(r4[rr][i], r8[rr][i]);  // :1253-1254
        const int ky = (rr ? 0 : 4) + k;  // pixel px of row y <-> bit 8 px + ky (:1257-1268)
        const uint32_t lo = ((mk.x >> ky) & 0x01010101u) * 0xffu, hi = ((mk.y >> ky) & 0x01010101u) * 0xffu;
        uint4 v;
        v.x = __vmaxu2(w[0] & __byte_perm(lo, 0u, 0x1100), d[y].x & keep);
        v.y = __vmaxu2(w[1] & __byte_perm(lo, 0u, 0x3322), d[y].y & keep);
        v.z = __vmaxu2(w[2] & __byte_perm(hi, 0u, 0x1100), d[y].z & keep);
        v.w = __vmaxu2(w[3] & __byte_perm(hi, 0u, 0x3322), d[y].w & keep);
        d[y] = v;
        mnAcc = __vminu2(mnAcc, __vminu2(__vminu2(v.x, v.y), __vminu2(v.z, v.w)));
      }
    h = min(mnAcc & 0xffffu, mnAcc >> 16);  // Rasterizer.cpp:1287-1290
    dirty = true;
  }
  __syncwarp();  // chain slots are rewritten by the next primitive
}

template <int C>
#if ORZ_CLUSTER_CTAS_PER_SM
__global__ void __launch_bounds__(kClusterGW * 32, ORZ_CLUSTER_CTAS_PER_SM) k_raster_views_cluster(const FrameParams p) {
#else
__global__ void __maxnreg__(ORZ_CLUSTER_REGS) k_raster_views_cluster(const FrameParams p) {
#endif
  constexpr uint32_t GW = kClusterGW, NT = GW * 32, kWarps = C * GW;
  extern __shared__ __align__(16) uint32_t s_dyn[];
  uint2* s_lut = reinterpret_cast<uint2*>(s_dyn);
  uint32_t* s_stageAll = s_dyn + ClusterSmem::kLutWords;
  uint32_t* s_idxAll = s_stageAll + ClusterSmem::kStageWords;
  float* s_chain = reinterpret_cast<float*>(s_idxAll + ClusterSmem::kIdxWords);
  uint32_t* s_head = s_dyn + ClusterSmem::kFixedWords;     // [nOcc][6]: status + gate rectangle of every order slot
  uint32_t* s_vis = s_head + p.nOcc * kHeadWords;          // [nOcc]: some warp saw a visible pixel
  uint32_t* s_doneLocal = s_vis + p.nOcc;                  // [nOcc]: warps of THIS CTA that answered "not on my tiles"
  uint32_t* s_doneCta = s_doneLocal + p.nOcc;              // [nOcc]: CTAs of the cluster whose 16 warps all answered
  uint16_t* s_hiz = reinterpret_cast<uint16_t*>(s_doneCta + p.nOcc);  // [GW][K][32]: HiZ of the tiles my warps own

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const uint32_t vrank = p.viewBase + blockIdx.x / (uint32_t)C;
  const uint32_t view = p.viewOrder ? p.viewOrder[vrank] : vrank;  // longest first: clusters are scheduled in grid order
  const uint32_t tid = threadIdx.x;
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31u);
  const uint32_t gw = (uint32_t)warp * (uint32_t)C + rank;  // my tiles: t % kWarps == gw
  const uint32_t K = p.clusterK;
  const uint32_t nOcc = p.nOcc;

  const uint32_t* front = p.frontBuf + (size_t)view * nOcc * kFrontWords;
  if (ORZ_CLUSTER_LUT_SMEM) for (uint32_t i = tid; i < 4096u; i += NT) s_lut[i] = p.lut[i];
  for (uint32_t i = tid; i < nOcc * kHeadWords; i += NT) s_head[i] = front[(size_t)(i / kHeadWords) * kFrontWords + i % kHeadWords];
  for (uint32_t i = tid; i < nOcc * 3u; i += NT) s_vis[i] = 0u;

  Target T;
  T.width = p.width; T.height = p.height; T.blocksX = p.width >> 3; T.blocksY = p.height >> 3;
  T.depth = p.depth + (size_t)view * p.depthStride;
  T.hiz = p.hiz + (size_t)view * p.hizStride;
  const uint32_t tilesX = (T.blocksX + kTileW - 1u) / kTileW, tilesY = (T.blocksY + kTileH - 1u) / kTileH, nTiles = tilesX * tilesY;
  const bool useGate = (p.flags & ORZ_BATCH_NO_GATE) == 0u;
  const bool forceClip = (p.flags & ORZ_BATCH_FORCE_CLIPPED) != 0u;
  const uint4* recInfo = p.recInfo + (size_t)view * nOcc * 2u;
  uint16_t* myHiz = s_hiz + (size_t)warp * K * 32u + lane;  // + 32 k
  float* myChain = s_chain + warp * (12 * kChainStride);
  uint32_t* myStage = s_stageAll + (uint32_t)warp * kStageCap * kRecStride;
  uint32_t* myIdx = s_idxAll + (uint32_t)warp * kStageCap;
  const uint32_t lx = (uint32_t)lane & 7u, ly = (uint32_t)lane >> 3;
  const bool reporter = rank == 0u && warp == 0 && lane == 0;  // writes the per-slot outputs of the view

  // lane k keeps the origin (in blocks) of my k-th tile; 0xffff = none
  uint32_t tileX0 = 0xffffu, tileY0 = 0xffffu;
  if ((uint32_t)lane < K) {
    const uint32_t t = gw + (uint32_t)lane * kWarps;
    if (t < nTiles) { const uint32_t ty = t / tilesX; tileX0 = (t - ty * tilesX) * kTileW; tileY0 = ty * kTileH; }
  }
  const uint32_t allTiles = __ballot_sync(kFull, tileX0 != 0xffffu);
  // clear (Rasterizer.cpp:107-121): HiZ := 1 on my tiles; depth is overwritten by the first update
  for (uint32_t m = allTiles; m; m &= m - 1u) {
    const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
    const uint32_t bx = __shfl_sync(kFull, tileX0, (int)k) + lx, by = __shfl_sync(kFull, tileY0, (int)k) + ly;
    myHiz[32u * k] = 1;
    if (bx < T.blocksX && by < T.blocksY) T.hiz[by * T.blocksX + bx] = 1;
  }
  __syncwarp();
  cluster.sync();  // tables staged, decision words zero in every CTA before the first remote access

  // "no visible pixel on my tiles" for candidate s: per-CTA count, forwarded by the CTA's last warp
  auto answer_no = [&](uint32_t s) {
    uint32_t old = 0u;
    if (lane == 0) old = atomicAdd(&s_doneLocal[s], 1u);
    old = __shfl_sync(kFull, old, 0);
    if (old == GW - 1u && lane < C) atomicAdd(cluster.map_shared_rank(&s_doneCta[s], (unsigned)lane), 1u);
  };
  // my tiles that meet the block rectangle [bx0, bx1] x [by0, by1] (inclusive), as a mask over k
  auto tiles_meeting = [&](uint32_t bx0, uint32_t bx1, uint32_t by0, uint32_t by1) -> uint32_t {
    return __ballot_sync(kFull, tileX0 != 0xffffu && tileX0 <= bx1 && tileX0 + kTileW > bx0 && tileY0 <= by1 && tileY0 + kTileH > by0);
  };

  // ---- candidates whose rectangle does not touch my tiles: answered before the walk starts
  for (uint32_t s = 0; s < nOcc; ++s) {
    const uint32_t* hd = s_head + s * kHeadWords;
    if (hd[0] != kBoxRect) continue;
    if (!tiles_meeting(hd[1] >> 3, hd[2] >> 3, hd[3] >> 3, hd[4] >> 3)) answer_no(s);
  }

  uint32_t quadsSubmitted = 0;
  for (uint32_t s = 0; s < nOcc; ++s) {
    const uint32_t* hd = s_head + s * kHeadWords;
    const uint32_t status = hd[0];
    if (status == kBoxCulled) {
      if (p.gate && reporter) p.gate[(size_t)view * nOcc + s] = 0;
      continue;
    }
    const uint4 info = recInfo[2u * s], box = recInfo[2u * s + 1u];  // requested now, needed after the gate
    bool visible = true, clipped = false;
    if (status == kBoxNearClip) {
      clipped = useGate ? true : forceClip;
    } else {
      // ---- gate: query2D (Rasterizer.cpp:283-349) on the part of the rectangle that lies on my tiles
      const uint32_t minX = hd[1], maxX = hd[2], minY = hd[3], maxY = hd[4], maxZ = hd[5];
      const uint32_t bx0 = minX >> 3, bx1 = maxX >> 3, by0 = minY >> 3, by1 = maxY >> 3;
      const uint32_t* vis = s_vis + s;
      uint32_t tm = tiles_meeting(bx0, bx1, by0, by1);
      if (tm && !ld_flag(vis)) {
        bool found = false;
        for (; tm; tm &= tm - 1u) {
          const uint32_t k = (uint32_t)__ffs((int)tm) - 1u;
          const uint32_t bx = __shfl_sync(kFull, tileX0, (int)k) + lx, by = __shfl_sync(kFull, tileY0, (int)k) + ly;
          if (ld_flag(vis)) break;  // another warp already found a visible pixel
          const bool hit = bx >= bx0 && bx <= bx1 && by >= by0 && by <= by1 && bx < T.blocksX && by < T.blocksY &&
                           query_block_h(T, bx, by, (uint32_t)myHiz[32u * k], minX, maxX, minY, maxY, maxZ);
          if (__any_sync(kFull, hit)) { found = true; break; }
        }
        if (found) { if (lane < C) st_flag_remote(s_vis + s, (uint32_t)lane, 1u); }
        else if (!ld_flag(vis)) answer_no(s);
      }
      // visible as soon as ONE warp says so, invisible when all 16 C warps have said no
      const uint32_t* done = s_doneCta + s;
      for (;;) {
        if (ld_flag(vis)) break;
        if (ld_flag(done) >= (uint32_t)C) { visible = ld_flag(vis) != 0u; break; }
#if ORZ_SPIN_NAP
        __nanosleep(ORZ_SPIN_NAP);  // (a longer or growing nap was measured slower: the wake-up delay sits on the dependency chain)
#endif
      }
    }
    if (reporter) {
      if (p.gate) p.gate[(size_t)view * nOcc + s] = (uint8_t)((visible ? 1 : 0) | (clipped && useGate ? 2 : 0));
      if (visible) quadsSubmitted += info.z;
    }
    if (!visible || info.x == 0u) continue;

    // ---- rasterize<clipped>(occluder): the records k_setup_views wrote, on my tiles
    uint32_t tmOcc = 0u;
    if (box.x < box.z) tmOcc = tiles_meeting(box.x, box.z - 1u, box.y, box.w - 1u);
    if (!tmOcc) continue;
    const uint32_t cnt = info.x;
    const size_t recBase = (size_t)view * p.totalQuads + info.y;
    const uint32_t* recs = p.recBuf + recBase * kRecStride;
    const uint2* hdrs = p.hdrBuf + recBase;
    if ((uint32_t)lane * 16u < cnt) prefetch_l1(hdrs + (uint32_t)lane * 16u);  // <= 504 headers = 32 lines

    uint32_t nStaged = 0;
    // staged records -> my tiles, tile-major, each tile's primitives in order
    auto flush = [&]() {
      __syncwarp();
#pragma unroll 8
      for (uint32_t i = 0; i < nStaged; ++i)
        if (lane < kRecStride) myStage[i * kRecStride + lane] = recs[(size_t)myIdx[i] * kRecStride + lane];
      __syncwarp();
      for (uint32_t m = tmOcc; m; m &= m - 1u) {
        const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
        const uint32_t x0 = __shfl_sync(kFull, tileX0, (int)k), y0 = __shfl_sync(kFull, tileY0, (int)k);
        const uint32_t x1 = min(x0 + kTileW, T.blocksX), y1 = min(y0 + kTileH, T.blocksY);
        bool touches = false;
        if ((uint32_t)lane < nStaged) {
          const uint32_t a = myStage[lane * kRecStride], b = myStage[lane * kRecStride + 1];
          const uint32_t minX = a & 0xffffu, minY = a >> 16;
          touches = minX < x1 && minX + (b & 0xffffu) > x0 && minY < y1 && minY + (b >> 16) > y0;
        }
        uint32_t hits = __ballot_sync(kFull, touches);
        if (!hits) continue;
        // bring the tile into registers
        const uint32_t bx = x0 + lx, by = y0 + ly;
        const bool inScreen = bx < x1 && by < y1;
        uint4* dp = reinterpret_cast<uint4*>(T.depth) + (size_t)(by * T.blocksX + bx) * 8u;
        uint32_t h = inScreen ? (uint32_t)myHiz[32u * k] : 0xffffu;  // off-screen lanes never pass
        const bool load = inScreen && h != 1u;
        uint4 d[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) d[y] = load ? dp[y] : make_uint4(0u, 0u, 0u, 0u);
        bool dirty = false;
        for (; hits; hits &= hits - 1u)
          tile_prim(myStage + ((uint32_t)__ffs((int)hits) - 1u) * kRecStride, lane, x0, y0, x1, y1, ORZ_CLUSTER_LUT_SMEM ? s_lut : p.lut, myChain, d, h, dirty);
        if (dirty) {
#pragma unroll
          for (int y = 0; y < 8; ++y) dp[y] = d[y];
          myHiz[32u * k] = (uint16_t)h;
          T.hiz[by * T.blocksX + bx] = (uint16_t)h;
        }
      }
      __syncwarp();
      nStaged = 0;
    };
    for (uint32_t r0 = 0; r0 < cnt; r0 += 32u) {
      uint32_t hx0 = 0, hx1 = 0, hy0 = 0, hy1 = 0;  // empty
      if (r0 + (uint32_t)lane < cnt) {
        const uint2 hdr = hdrs[r0 + (uint32_t)lane];
        hx0 = hdr.x & 0xffffu; hy0 = hdr.x >> 16; hx1 = hx0 + (hdr.y & 0xffffu); hy1 = hy0 + (hdr.y >> 16);
      }
      bool touches = false;
      for (uint32_t m = tmOcc; m; m &= m - 1u) {
        const int k = __ffs((int)m) - 1;
        const uint32_t x0 = __shfl_sync(kFull, tileX0, k), y0 = __shfl_sync(kFull, tileY0, k);
        touches = touches || (hx0 < x0 + kTileW && hx1 > x0 && hy0 < y0 + kTileH && hy1 > y0);
      }
      uint32_t hits = __ballot_sync(kFull, touches);
      while (hits) {
        const uint32_t take = min(kStageCap - nStaged, (uint32_t)__popc(hits));
        const uint32_t myRank = (uint32_t)__popc(hits & ((1u << lane) - 1u));
        if (((hits >> lane) & 1u) && myRank < take) myIdx[nStaged + myRank] = r0 + (uint32_t)lane;
        for (uint32_t i = 0; i < take; ++i) hits &= hits - 1u;
        nStaged += take;
        if (nStaged == kStageCap) flush();
      }
    }
    if (nStaged) flush();
  }
  if (p.quadsSubmitted && reporter) p.quadsSubmitted[view] = quadsSubmitted;
  if (p.exportDepth) {  // canonical depth for the caller: blocks that stayed cleared read as zero
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t m = allTiles; m; m &= m - 1u) {
      const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
      const uint32_t bx = __shfl_sync(kFull, tileX0, (int)k) + lx, by = __shfl_sync(kFull, tileY0, (int)k) + ly;
      if (bx < T.blocksX && by < T.blocksY && myHiz[32u * k] == 1) {
        uint4* d4 = reinterpret_cast<uint4*>(T.depth) + (size_t)(by * T.blocksX + bx) * 8u;
#pragma unroll
        for (int y = 0; y < 8; ++y) d4[y] = z;
      }
    }
  }
  cluster.sync();  // no CTA may leave while another one can still write its decision words
}

// queryVisibility for every (view, occludee box) on the finished buffers; Rasterizer.cpp:123-349
__global__ void __launch_bounds__(256) k_query_views(const FrameParams p) {
  __shared__ ViewMatrices s_vm;
  const uint32_t view = p.viewOrder ? p.viewOrder[p.viewBase + blockIdx.y] : p.viewBase + blockIdx.y, tid = threadIdx.x;
  if (tid < 32) reinterpret_cast<float*>(&s_vm)[tid] = reinterpret_cast<const float*>(p.vmBuf + view)[tid];
  __syncthreads();
  const RcpTable rt{p.rcp, p.rcpShift};
  Target T;
  T.width = p.width; T.height = p.height; T.blocksX = p.width >> 3; T.blocksY = p.height >> 3;
  T.depth = p.depth + (size_t)view * p.depthStride;
  T.hiz = p.hiz + (size_t)view * p.hizStride;
  const uint32_t i = blockIdx.x * blockDim.x + tid;
  BoxFront f;
  f.status = kBoxCulled; f.minX = f.maxX = f.minY = f.maxY = f.maxZ = 0;
  if (i < p.nBoxes) {
    const float4 mn = p.boxes[2 * (size_t)i], mx = p.boxes[2 * (size_t)i + 1];
    const float bmn[4] = {mn.x, mn.y, mn.z, mn.w}, bmx[4] = {mx.x, mx.y, mx.z, mx.w};
    f = box_front_half(s_vm, bmn, bmx, p.width, p.height, rt);
  }
  const bool clip = f.status == kBoxNearClip;
  const bool seen = query2d_warp(T, f, (int)(tid & 31u));  // every lane must take part (warp collectives inside)
  const bool vis = clip || seen;
  const uint32_t vb = __ballot_sync(kFull, vis), cb = __ballot_sync(kFull, clip);
  const uint32_t word = i >> 5;
  if ((tid & 31u) == 0 && word < p.bitWords) {
    if (p.visBits) p.visBits[(size_t)view * p.bitWords + word] = vb;
    if (p.clipBits) p.clipBits[(size_t)view * p.bitWords + word] = cb;
  }
}

// ---------------------------------------------------------------------------------------------
// "Wide" path for few views over very many ungated occluders (BASELINE config 4: 5 M near-clipped
// quads at 3840x2160, every batch through rasterize<true>, no gate).  A view group of 4-8 warps
// cannot fill the GPU with one view, so the view is split the other way:
//   k_slot_prefix    quads before each order slot (one thread, nOcc is ~10^4)
//   k_setup_wide     ALL quads of the view set up in parallel, one lane per quad, records
//                    compacted in order per 32-quad chunk into global memory (+ the rows a chunk touches)
//   k_raster_wide    one warp per screen block-row walks the chunk list in order and traverses
//                    the primitives that touch its row -- per-block order preserved, no atomics
__global__ void __launch_bounds__(1024) k_slot_prefix(const FrameParams p, uint32_t view, uint32_t* __restrict__ slotStart) {
  // exclusive prefix sum of quadCount over the order slots: contiguous ranges per thread + block scan
  __shared__ uint32_t s_warp[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t* order = p.orders ? p.orders + (size_t)view * p.nOcc : p.orderBuf + (size_t)view * p.nOcc;
  const uint32_t per = (p.nOcc + blockDim.x - 1) / blockDim.x;
  const uint32_t s0 = min(tid * per, p.nOcc), s1 = min(s0 + per, p.nOcc);
  uint32_t sum = 0;
  for (uint32_t s = s0; s < s1; ++s) sum += p.occ[order[s]].quadCount;
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, d); if (lane >= (uint32_t)d) incl += t; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(kFull, wi, d); if (lane >= (uint32_t)d) wi += t; }
    s_warp[lane] = wi - w;  // exclusive offset of each warp
  }
  __syncthreads();
  uint32_t acc = s_warp[warp] + incl - sum;
  for (uint32_t s = s0; s < s1; ++s) {
    slotStart[s] = acc;
    acc += p.occ[order[s]].quadCount;
    if (p.gate) p.gate[(size_t)view * p.nOcc + s] = 1;
  }
  if (tid == blockDim.x - 1) {
    slotStart[p.nOcc] = acc;
    if (p.quadsSubmitted) p.quadsSubmitted[view] = acc;
  }
}

constexpr int kWideRecWords = 21;  // same record as store_record writes (20 words + the division magic)
__global__ void __launch_bounds__(256) k_setup_wide(const FrameParams p, uint32_t view, const uint32_t* __restrict__ slotStart,
                                                     uint32_t totalQuads, uint32_t* __restrict__ recs, uint32_t* __restrict__ chunkCount,
                                                     uint32_t* __restrict__ chunkRows) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;  // quad index in submission order
  const int lane = (int)(threadIdx.x & 31u);
  const uint32_t chunk = g >> 5;
  const RcpTable rt{p.rcp, p.rcpShift};
  const uint32_t* order = p.orders ? p.orders + (size_t)view * p.nOcc : p.orderBuf + (size_t)view * p.nOcc;
  const bool forceClip = (p.flags & ORZ_BATCH_FORCE_CLIPPED) != 0u;
  bool ok = false;
  Prim P;
  if (g < totalQuads) {
    uint32_t lo = 0, hi = p.nOcc;  // last slot with slotStart <= g
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (slotStart[mid] <= g) lo = mid; else hi = mid; }
    const OccMeta& om = p.occ[order[lo]];
    const uint32_t* fr = p.frontBuf + ((size_t)view * p.nOcc + lo) * kFrontWords;
    CallMatrix cm;
#pragma unroll
    for (int k = 0; k < 4; ++k) { cm.rx[k] = u2f(fr[6 + k]); cm.ry[k] = u2f(fr[10 + k]); cm.rw[k] = u2f(fr[14 + k]); }
    cm.c0 = u2f(fr[18]); cm.c1 = u2f(fr[19]);
    const uint4 v = p.quads[om.quadOffset + (g - slotStart[lo])];
    const uint32_t word[4] = {v.x, v.y, v.z, v.w};
    const int32_t bx = (int32_t)(p.width >> 3), by = (int32_t)(p.height >> 3);
    ok = forceClip ? setup_quad<true>(word, cm, rt, c_modeNibbles, bx, by, P) : setup_quad<false>(word, cm, rt, c_modeNibbles, bx, by, P);
  }
  const uint32_t valid = __ballot_sync(kFull, ok);
  // rows (in linear block space, with the 16-bit wrap of Rasterizer.cpp:1054) this chunk touches
  uint32_t rLo = 0xffffffffu, rHi = 0u;
  if (ok) {
    store_record(recs + ((size_t)chunk * 32u + (uint32_t)__popc(valid & ((1u << lane) - 1u))) * kWideRecWords, P);
    const uint32_t blocksX = p.width >> 3;
    const uint32_t fb = (((uint32_t)P.minY * blocksX) & 0xffffu) + (uint32_t)P.minX;
    const bool wrap = blocksX * (p.height >> 3) > 65536u;
    const uint32_t r0 = wrap ? fb / blocksX : (uint32_t)P.minY, c0 = wrap ? fb - r0 * blocksX : (uint32_t)P.minX;
    rLo = r0;
    rHi = r0 + (uint32_t)P.rangeY - 1u + ((c0 + (uint32_t)P.rangeX > blocksX) ? 1u : 0u);
  }
  rLo = __reduce_min_sync(kFull, rLo);
  rHi = __reduce_max_sync(kFull, rHi);
  if (lane == 0 && chunk * 32u < ((totalQuads + 31u) & ~31u)) {
    chunkCount[chunk] = (uint32_t)__popc(valid);
    chunkRows[chunk] = valid ? (rLo | (rHi << 16)) : 0xffffu;  // empty chunk: lo > hi
  }
}

__global__ void k_clear_hiz(uint16_t* hiz, uint32_t blocks) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < blocks; k += gridDim.x * blockDim.x) hiz[k] = 1;
}
__global__ void k_zero_cleared(uint16_t* depth, const uint16_t* hiz, uint32_t blocks) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < blocks; i += gridDim.x * blockDim.x)
    if (hiz[i] == 1) {
      uint4* d4 = reinterpret_cast<uint4*>(depth) + (size_t)i * 8u;
#pragma unroll
      for (int k = 0; k < 8; ++k) d4[k] = z;
    }
}

__global__ void __launch_bounds__(128) k_raster_wide(Target T, const uint2* __restrict__ lut, const uint32_t* __restrict__ recs,
                                                      const uint32_t* __restrict__ chunkCount, const uint32_t* __restrict__ chunkRows,
                                                      uint32_t nChunks, uint32_t nSeg, uint32_t segWidth) {
  const int lane = (int)(threadIdx.x & 31u);
  // this warp owns the blocks of screen block-row `row` whose column lies in [colLo, colHi)
  const uint32_t wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t row = wid / nSeg, seg = wid - row * nSeg;
  if (row >= T.blocksY) return;
  const uint32_t blocksX = T.blocksX;
  const uint32_t colLo = seg * segWidth, colHi = seg + 1u == nSeg ? blocksX : colLo + segWidth;
  const bool wrap = blocksX * T.blocksY > 65536u;
  for (uint32_t c0 = 0; c0 < nChunks; c0 += 32) {
    // 32 chunk summaries at a time: which of them touch my row?
    const uint32_t ci = c0 + (uint32_t)lane;
    const uint32_t rr = ci < nChunks ? chunkRows[ci] : 0xffffu;
    uint32_t hitChunks = __ballot_sync(kFull, (rr & 0xffffu) <= row && row <= (rr >> 16));
    while (hitChunks) {
      const uint32_t cj = c0 + (uint32_t)__ffs((int)hitChunks) - 1u;
      hitChunks &= hitChunks - 1u;
      const uint32_t cnt = chunkCount[cj];
      const uint32_t* base = recs + (size_t)cj * 32u * kWideRecWords;
      bool mine = false;
      if ((uint32_t)lane < cnt) {  // does primitive `lane` of this chunk touch my row?
        const uint32_t w0 = base[(size_t)lane * kWideRecWords + 0], w1 = base[(size_t)lane * kWideRecWords + 1];
        const uint32_t minX = w0 & 0xffffu, minY = w0 >> 16, rangeX = w1 & 0xffffu, rangeY = w1 >> 16;
        const uint32_t fb = ((minY * blocksX) & 0xffffu) + minX;
        const uint32_t r0 = wrap ? fb / blocksX : minY, cc = wrap ? fb - r0 * blocksX : minX;
        const bool crossing = cc + rangeX > blocksX;
        mine = row >= r0 && row <= r0 + rangeY - 1u + (crossing ? 1u : 0u) && (crossing || (cc < colHi && cc + rangeX > colLo));
      }
      uint32_t hits = __ballot_sync(kFull, mine);
      while (hits) {
        const uint32_t k = (uint32_t)__ffs((int)hits) - 1u;
        hits &= hits - 1u;
        raster_prim<0, true>(base + (size_t)k * kWideRecWords, lane, row, T.blocksY + 1u, T, lut, colLo, colHi);  // stride > rows: one row is mine
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Single-view kernels behind the per-call API (Rasterizer.h:13-26)
__global__ void k_clear(uint16_t* depth, uint16_t* hiz, uint32_t blocks) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
  uint4* d4 = reinterpret_cast<uint4*>(depth);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t k = i; k < blocks * 8u; k += n) d4[k] = z;
  for (uint32_t k = i; k < blocks; k += n) hiz[k] = 1;
}

// rasterize<clipped>(occluder) for one view: every CTA sets up all quads of the batch (cheap,
// <= 504 quads) and traverses only the block rows its warps own, so no inter-CTA ordering is needed.
template <int GW>
__global__ void __launch_bounds__(GW * 32) k_rasterize_single(const ViewMatrices vm, const uint4* quads, uint32_t nq,
                                                               const float4 refMin, const float4 refMax, int clipped, Target T,
                                                               const uint32_t* rcp, int rcpShift, const uint2* lut) {
  constexpr uint32_t NT = GW * 32;
  __shared__ uint32_t s_recs[NT * kRecStride];
  __shared__ uint32_t s_count[GW];
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
  const RcpTable rt{rcp, rcpShift};
  CallMatrix cm;
  const float rmn[4] = {refMin.x, refMin.y, refMin.z, refMin.w}, rmx[4] = {refMax.x, refMax.y, refMax.z, refMax.w};
  prepare_call(vm.baked, rmn, rmx, cm);
  const uint32_t rowStride = gridDim.x * GW, rowPhase = blockIdx.x * GW + (uint32_t)warp;
  for (uint32_t q0 = 0; q0 < nq; q0 += NT) {
    setup_chunk(quads, q0, nq, clipped != 0, cm, rt, T, warp, lane, s_recs, s_count);
    __syncthreads();
#pragma unroll 1
    for (int w2 = 0; w2 < GW; ++w2) {
      const uint32_t cnt = s_count[w2];
      for (uint32_t i = 0; i < cnt; ++i)
        raster_prim<0>(s_recs + ((uint32_t)w2 * 32u + i) * kRecStride, lane, rowPhase, rowStride, T, lut);
    }
    __syncthreads();
  }
}

// setup records of every quad, uncompacted (parity tests of the setup stage)
__global__ void k_debug_setup(const ViewMatrices vm, const uint4* quads, uint32_t nq, const float4 refMin, const float4 refMax,
                              int clipped, Target T, const uint32_t* rcp, int rcpShift, orz_prim_record* out) {
  const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const RcpTable rt{rcp, rcpShift};
  CallMatrix cm;
  const float rmn[4] = {refMin.x, refMin.y, refMin.z, refMin.w}, rmx[4] = {refMax.x, refMax.y, refMax.z, refMax.w};
  prepare_call(vm.baked, rmn, rmx, cm);
  const uint4 v = quads[qi];
  const uint32_t word[4] = {v.x, v.y, v.z, v.w};
  Prim P;
  const bool ok = clipped ? setup_quad<true>(word, cm, rt, c_modeNibbles, (int32_t)T.blocksX, (int32_t)T.blocksY, P)
                          : setup_quad<false>(word, cm, rt, c_modeNibbles, (int32_t)T.blocksX, (int32_t)T.blocksY, P);
  orz_prim_record r;
  memset(&r, 0, sizeof r);
  if (ok) {
    r.mode = P.mode; r.minX = P.minX; r.minY = P.minY; r.rangeX = P.rangeX; r.rangeY = P.rangeY; r.maxZ = P.maxZ;
    r.dzdx = P.dzdx; r.dzdy = P.dzdy; r.plane0 = P.plane0;
    for (int e = 0; e < 4; ++e) { r.nx[e] = P.nx[e]; r.ny[e] = P.ny[e]; r.off[e] = P.off[e]; r.slope[e] = P.slope[e]; }
  }
  out[qi] = r;
}

// queryVisibility for n boxes, one thread each; out[i] bit0 visible, bit1 needsClipping
__global__ void k_query_boxes(const ViewMatrices vm, const float4* boxes, uint32_t n, Target T, const uint32_t* rcp, int rcpShift,
                              uint8_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const RcpTable rt{rcp, rcpShift};
  BoxFront f;
  f.status = kBoxCulled; f.minX = f.maxX = f.minY = f.maxY = f.maxZ = 0;
  if (i < n) {
    const float4 mn = boxes[2 * (size_t)i], mx = boxes[2 * (size_t)i + 1];
    const float bmn[4] = {mn.x, mn.y, mn.z, mn.w}, bmx[4] = {mx.x, mx.y, mx.z, mx.w};
    f = box_front_half(vm, bmn, bmx, T.width, T.height, rt);
  }
  const bool vis = query2d_warp(T, f, (int)(threadIdx.x & 31u));
  if (i < n) out[i] = f.status == kBoxNearClip ? 3 : (vis ? 1 : 0);
}

__global__ void k_query2d(Target T, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ, uint32_t* out) {
  __shared__ uint32_t s_flag;
  if (threadIdx.x == 0) s_flag = 0u;
  __syncthreads();
  query2d_coop(T, minX, maxX, minY, maxY, maxZ, threadIdx.x, blockDim.x, &s_flag);
  __syncthreads();
  if (threadIdx.x == 0) *out = s_flag;
}

// readBackDepth, Rasterizer.cpp:351-399: one thread per pixel, BGRA8 row-major
__global__ void k_readback(Target T, uint8_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T.width * T.height) return;
  const uint32_t x = i % T.width, y = i / T.width;
  const uint32_t b = (y >> 3) * T.blocksX + (x >> 3);
  uchar4 px = make_uchar4(0, 0, 0, 0);
  if (T.hiz[b] != 1) {
    const float bias = 3.9623753e+28f;
    const float depth = u2f((uint32_t)T.depth[(size_t)b * 64u + (y & 7u) * 8u + (x & 7u)] << 12) * bias;
    const float lin = (2 * 0.25f) / ((0.25f + 1000.0f) - (1.0f - depth) * (1000.0f - 0.25f));
    const uint32_t d = (uint32_t)(100 * 256 * lin);
    px = make_uchar4((uint8_t)(d / 100u), (uint8_t)(d % 256u), 0, 255);
  }
  reinterpret_cast<uchar4*>(out)[i] = px;
}

// canonical export: cleared blocks (HiZ == 1) read as zero -- already true by construction since
// clear zeroes depth; kept as a copy kernel so downloads never expose garbage after a natural HiZ==1
__global__ void k_canonical_depth(const uint16_t* depth, const uint16_t* hiz, uint32_t blocks, uint4* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= blocks * 8u) return;
  const uint4 v = reinterpret_cast<const uint4*>(depth)[i];
  out[i] = hiz[i >> 3] == 1 ? make_uint4(0u, 0u, 0u, 0u) : v;
}


// ---------------------------------------------------------------------------------------------
// Occluder::bake (Occluder.cpp:7-181) on the GPU, one CTA per batch: quad normals -> k-means by
// facing (6 axis seeds, at most 10 rounds) -> stable regroup by cluster -> 11/11/10 quantisation
// -> one uint4 per quad (and, when asked, the reference's packet layout) -> bounds and centre.
// Bit-exact with the host bake: every float sum runs in the reference's order (the cluster sums
// are accumulated quad by quad by one thread per (cluster, component)), rsqrtps through its
// table model (rsqrt_x86), products rounded separately (-fmad=false).
struct BakeJob {
  uint32_t vertOffset;  // first vertex (float4) of the batch
  uint32_t nQuads;
  uint32_t quadOffset;  // first output quad
  uint32_t pad;
};

__device__ __forceinline__ void bake_normal(const float4 v0, const float4 v1, const float4 v2, float& x, float& y, float& z) {
  // normal() of VectorMath.h:6-18: cross(v1 - v0, v2 - v0)
  const float ax = v1.x - v0.x, ay = v1.y - v0.y, az = v1.z - v0.z, bx = v2.x - v0.x, by = v2.y - v0.y, bz = v2.z - v0.z;
  x = ay * bz - az * by; y = az * bx - ax * bz; z = ax * by - ay * bx;
}

__global__ void __launch_bounds__(256) k_bake(const float4* __restrict__ verts, const BakeJob* __restrict__ jobs, const float4 refMin,
                                               const float4 refMax, const RsqrtTable rs, uint4* __restrict__ outQuads, OccMeta* __restrict__ meta,
                                               uint32_t* __restrict__ outPackets) {
  extern __shared__ __align__(16) float s_bake[];
  __shared__ float s_seed[6][3], s_sum[6][3];
  __shared__ float s_mn[8][4], s_mx[8][4];
  __shared__ uint32_t s_mnI[8][4], s_mxI[8][4];
  const BakeJob job = jobs[blockIdx.x];
  const uint32_t n = job.nQuads, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  float* nx = s_bake; float* ny = nx + n; float* nz = ny + n;
  uint32_t* cl = reinterpret_cast<uint32_t*>(nz + n);
  uint32_t* pos = cl + n;
  const float4* v = verts + job.vertOffset;

  // quad normals (Occluder.cpp:12-21) and the bounds over all four lanes (Occluder.cpp:159-170).
  // minps / maxps keep the EARLIER vertex when two compare equal (+0 / -0), so the reduction
  // carries the vertex index and breaks ties towards the lower one: same result as the serial loop.
  float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  uint32_t mnI[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, mxI[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
  for (uint32_t q = tid; q < n; q += 256u) {
    const float4 v0 = v[4 * q], v1 = v[4 * q + 1], v2 = v[4 * q + 2], v3 = v[4 * q + 3];
    float ax, ay, az, bx, by, bz;
    bake_normal(v0, v1, v2, ax, ay, az);
    bake_normal(v0, v2, v3, bx, by, bz);
    const float sx = ax + bx, sy = ay + by, sz = az + bz;
    const float r = rsqrt_x86((sx * sx + sy * sy) + sz * sz, rs);  // normalize(), VectorMath.h:20-23; dpps 0x7F sum order
    nx[q] = sx * r; ny[q] = sy * r; nz[q] = sz * r;
    cl[q] = 0u;
    const float vv[4][4] = {{v0.x, v0.y, v0.z, v0.w}, {v1.x, v1.y, v1.z, v1.w}, {v2.x, v2.y, v2.z, v2.w}, {v3.x, v3.y, v3.z, v3.w}};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (vv[j][k] < mn[k]) { mn[k] = vv[j][k]; mnI[k] = 4u * q + (uint32_t)j; }
        if (vv[j][k] > mx[k]) { mx[k] = vv[j][k]; mxI[k] = 4u * q + (uint32_t)j; }
      }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const float om = __shfl_xor_sync(kFull, mn[k], d), oM = __shfl_xor_sync(kFull, mx[k], d);
      const uint32_t omI = __shfl_xor_sync(kFull, mnI[k], d), oMI = __shfl_xor_sync(kFull, mxI[k], d);
      if (om < mn[k] || (om == mn[k] && omI < mnI[k])) { mn[k] = om; mnI[k] = omI; }
      if (oM > mx[k] || (oM == mx[k] && oMI < mxI[k])) { mx[k] = oM; mxI[k] = oMI; }
    }
    if (lane == 0) { s_mn[warp][k] = mn[k]; s_mx[warp][k] = mx[k]; s_mnI[warp][k] = mnI[k]; s_mxI[warp][k] = mxI[k]; }
  }
  if (tid < 18) s_seed[tid / 3][tid % 3] = 0.0f;
  __syncthreads();
  if (tid == 0) { s_seed[0][0] = 1.0f; s_seed[1][1] = 1.0f; s_seed[2][2] = 1.0f; s_seed[3][1] = -1.0f; s_seed[4][2] = -1.0f; s_seed[5][0] = -1.0f; }
  __syncthreads();

  // k-means by facing (Occluder.cpp:23-78)
  for (int round = 0; round < 10; ++round) {
    int moved = 0;
    for (uint32_t q = tid; q < n; q += 256u) {
      float best = -INFINITY;
      uint32_t pick = 0;
#pragma unroll
      for (uint32_t k = 0; k < 6; ++k) {
        const float d = (s_seed[k][0] * nx[q] + s_seed[k][1] * ny[q]) + s_seed[k][2] * nz[q];
        if (d >= best) { best = d; pick = k; }  // _mm_comige_ss: false when unordered
      }
      if (cl[q] != pick) { cl[q] = pick; moved = 1; }
    }
    if (!__syncthreads_or(moved)) break;  // the seeds are not used after the last round
    if (tid < 18) {  // cluster sums in quad order, one thread per (cluster, component)
      const uint32_t k = tid / 3u;
      const float* comp = tid % 3u == 0 ? nx : (tid % 3u == 1 ? ny : nz);
      float acc = 0.0f;
      for (uint32_t q = 0; q < n; ++q)
        if (cl[q] == k) acc = acc + comp[q];
      s_sum[k][tid % 3u] = acc;
    }
    __syncthreads();
    if (tid < 6) {
      const float x = s_sum[tid][0], y = s_sum[tid][1], z = s_sum[tid][2];
      const float r = rsqrt_x86((x * x + y * y) + z * z, rs);
      s_seed[tid][0] = x * r; s_seed[tid][1] = y * r; s_seed[tid][2] = z * r;
    }
    __syncthreads();
  }

  // stable regroup by cluster (Occluder.cpp:80-93): slot of quad q = quads of lower clusters + earlier quads of its own
  if (warp == 0) {
    uint32_t count[6] = {0, 0, 0, 0, 0, 0};
    for (uint32_t q0 = 0; q0 < n; q0 += 32u) {
      const uint32_t c = q0 + lane < n ? cl[q0 + lane] : 7u;
#pragma unroll
      for (uint32_t k = 0; k < 6; ++k) {
        const uint32_t m = __ballot_sync(kFull, c == k);
        if (c == k) pos[q0 + lane] = count[k] + (uint32_t)__popc(m & ((1u << lane) - 1u));
        count[k] += (uint32_t)__popc(m);
      }
    }
    uint32_t base[6];
    base[0] = 0;
#pragma unroll
    for (int k = 1; k < 6; ++k) base[k] = base[k - 1] + count[k - 1];
    for (uint32_t q = lane; q < n; q += 32u) {
      const uint32_t c = cl[q];
#pragma unroll
      for (uint32_t k = 0; k < 6; ++k) if (c == k) pos[q] += base[k];
    }
  }
  __syncthreads();

  // quantise and pack (Occluder.cpp:97-156): word = (X - 1024) << 21 | Y << 10 | Z
  const float ivx = 1.0f / (refMax.x - refMin.x), ivy = 1.0f / (refMax.y - refMin.y), ivz = 1.0f / (refMax.z - refMin.z);
  for (uint32_t q = tid; q < n; q += 256u) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 p = v[4 * q + j];
      const uint32_t cx = (uint32_t)cvtt_x86(ORZ_FMA((p.x - refMin.x) * ivx, 2047.0f, 0.5f));
      const uint32_t cy = (uint32_t)cvtt_x86(ORZ_FMA((p.y - refMin.y) * ivy, 2047.0f, 0.5f));
      const uint32_t cz = (uint32_t)cvtt_x86(ORZ_FMA((p.z - refMin.z) * ivz, 1023.0f, 0.5f));
      w[j] = ((cx - 1024u) << 21) | (cy << 10) | cz;
    }
    const uint32_t at = pos[q];
    outQuads[job.quadOffset + at] = make_uint4(w[0], w[1], w[2], w[3]);
    if (outPackets) {  // the reference's own layout: group of 8 quads = 4 x 8 words
      uint32_t* pk = outPackets + (size_t)job.quadOffset * 4u + (size_t)(at >> 3) * 32u + (at & 7u);
      pk[0] = w[0]; pk[8] = w[1]; pk[16] = w[2]; pk[24] = w[3];
    }
  }
  if (tid < 4 && meta) {  // bounds, w := 1 (Occluder.cpp:172-173), centre
    float lo = s_mn[0][tid], hi = s_mx[0][tid];
    uint32_t loI = s_mnI[0][tid], hiI = s_mxI[0][tid];
    for (int w2 = 1; w2 < 8; ++w2) {
      if (s_mn[w2][tid] < lo || (s_mn[w2][tid] == lo && s_mnI[w2][tid] < loI)) { lo = s_mn[w2][tid]; loI = s_mnI[w2][tid]; }
      if (s_mx[w2][tid] > hi || (s_mx[w2][tid] == hi && s_mxI[w2][tid] < hiI)) { hi = s_mx[w2][tid]; hiI = s_mxI[w2][tid]; }
    }
    if (tid == 3) { lo = 1.0f; hi = 1.0f; }
    OccMeta& om = meta[blockIdx.x];
    om.boundsMin[tid] = lo; om.boundsMax[tid] = hi; om.center[tid] = (hi + lo) * 0.5f;
    om.refMin[tid] = tid == 0 ? refMin.x : tid == 1 ? refMin.y : tid == 2 ? refMin.z : refMin.w;
    om.refMax[tid] = tid == 0 ? refMax.x : tid == 1 ? refMax.y : tid == 2 ? refMax.z : refMax.w;
    if (tid == 0) { om.quadOffset = job.quadOffset; om.quadCount = n; om.pad0 = om.pad1 = 0u; }
  }
}

}  // namespace orz

// =================================================================================================
// C ABI
using namespace orz;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define ORZ_CUDA(x)                                                                                   \
  do {                                                                                                \
    cudaError_t _e = (x);                                                                             \
    if (_e != cudaSuccess) return fail(ORZ_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct orz_context {
  int device = 0;
  int numSMs = 0;
  cudaStream_t stream = nullptr;
  uint2* d_lut = nullptr;
  uint32_t* d_rcp = nullptr;
  int rcpBits = 0;
  bool rcpExact = true;
  std::vector<uint32_t> h_rcp;
  uint64_t launches = 0;
  int groupWarps = 0;
  int traversal = 2;  // 1 = warp per block, 2 = lane per block (default)
  int clusterViews = 1024;  // batches of at most this many views run one thread-block cluster per view (0 = never)
  int clusterSize = 0;     // CTAs per cluster (2, 4, 8, 16); 0 = automatic
  // grow-only device scratch
  uint32_t* d_counter = nullptr;
  void* d_scratch[12] = {nullptr};
  size_t scratchBytes[12] = {0};
  uint32_t* h_pinned = nullptr;  // small pinned mailbox for scalar results
  static constexpr int kGroups = 4;          // sub-batches pipelined on auxiliary streams
  cudaStream_t aux[kGroups] = {nullptr};
  cudaEvent_t evFork = nullptr, evJoin[kGroups] = {nullptr};
  size_t arenaBudget = size_t(8) << 30;  // bytes of internal per-view depth+HiZ targets (views are chunked to fit)
};
struct orz_occluder {
  orz_context* ctx;
  uint4* d_quads;
  uint32_t nQuads;
  float refMin[4], refMax[4];
};
struct orz_rasterizer {
  orz_context* ctx;
  Target T;
  ViewMatrices vm;
  uint8_t* d_boxOut = nullptr;
};
struct orz_scene {
  orz_context* ctx;
  uint4* d_quads = nullptr;
  OccMeta* d_occ = nullptr;
  uint32_t nOcc = 0, totalQuads = 0;
  float4* d_boxes = nullptr;
  uint32_t nBoxes = 0;
};

extern "C" const char* orz_last_error(void) { return g_err.c_str(); }
extern "C" int orz_version(void) { return 100; }

static int ensure_scratch(orz_context* ctx, int idx, size_t bytes) {
  if (ctx->scratchBytes[idx] >= bytes) return ORZ_OK;
  if (ctx->d_scratch[idx]) cudaFree(ctx->d_scratch[idx]);
  ctx->d_scratch[idx] = nullptr;
  ctx->scratchBytes[idx] = 0;
  ORZ_CUDA(cudaMalloc(&ctx->d_scratch[idx], bytes));
  ctx->scratchBytes[idx] = bytes;
  return ORZ_OK;
}

extern "C" int orz_context_create(int device, orz_context** out) {
  if (!out) return fail(ORZ_ERR_ARG, "orz_context_create: out is NULL");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return fail(ORZ_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
  if (device < 0 || device >= n) return fail(ORZ_ERR_ARG, "orz_context_create: bad device index");
  ORZ_CUDA(cudaSetDevice(device));
  orz_context* ctx = new orz_context();
  ctx->device = device;
  cudaDeviceProp prop;
  ORZ_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->numSMs = prop.multiProcessorCount;
  ctx->arenaBudget = std::min<size_t>(size_t(24) << 30, prop.totalGlobalMem / 6);
  ORZ_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ORZ_CUDA(cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
  for (int g = 0; g < orz_context::kGroups; ++g) {
    ORZ_CUDA(cudaStreamCreateWithFlags(&ctx->aux[g], cudaStreamNonBlocking));
    ORZ_CUDA(cudaEventCreateWithFlags(&ctx->evJoin[g], cudaEventDisableTiming));
  }
  ORZ_CUDA(cudaMalloc(&ctx->d_lut, 4096 * sizeof(uint2)));
  ORZ_CUDA(cudaMemcpy(ctx->d_lut, edge_mask_table(), 4096 * sizeof(uint2), cudaMemcpyHostToDevice));
  probe_host_rcp(ctx->h_rcp, ctx->rcpBits, ctx->rcpExact);
  ORZ_CUDA(cudaMalloc(&ctx->d_rcp, ctx->h_rcp.size() * 4));
  ORZ_CUDA(cudaMemcpy(ctx->d_rcp, ctx->h_rcp.data(), ctx->h_rcp.size() * 4, cudaMemcpyHostToDevice));
  ORZ_CUDA(cudaMalloc(&ctx->d_counter, 64));
  ORZ_CUDA(cudaHostAlloc(&ctx->h_pinned, 64, cudaHostAllocDefault));
  *out = ctx;
  return ORZ_OK;
}
extern "C" void orz_context_destroy(orz_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& p : ctx->d_scratch) if (p) cudaFree(p);
  cudaFree(ctx->d_lut); cudaFree(ctx->d_rcp); cudaFree(ctx->d_counter);
  cudaFreeHost(ctx->h_pinned);
  for (int g = 0; g < orz_context::kGroups; ++g) { cudaStreamSynchronize(ctx->aux[g]); cudaStreamDestroy(ctx->aux[g]); cudaEventDestroy(ctx->evJoin[g]); }
  cudaEventDestroy(ctx->evFork);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}
extern "C" int orz_context_synchronize(orz_context* ctx) {
  ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return ORZ_OK;
}
extern "C" void* orz_context_stream(orz_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t orz_context_launch_count(orz_context* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int orz_context_set_group_warps(orz_context* ctx, int warps) {
  if (warps != 0 && warps != 1 && warps != 2 && warps != 4 && warps != 8 && warps != 16) return fail(ORZ_ERR_ARG, "group warps must be 0, 1, 2, 4, 8 or 16");
  ctx->groupWarps = warps;
  return ORZ_OK;
}
extern "C" int orz_context_set_traversal(orz_context* ctx, int mapping) {
  if (!ctx || (mapping != 1 && mapping != 2)) return fail(ORZ_ERR_ARG, "traversal mapping must be 1 (warp per block) or 2 (lane per block)");
  ctx->traversal = mapping;
  return ORZ_OK;
}
extern "C" int orz_context_set_cluster_views(orz_context* ctx, int maxViews) {
  if (!ctx || maxViews < 0) return fail(ORZ_ERR_ARG, "orz_context_set_cluster_views: bad arguments");
  ctx->clusterViews = maxViews;
  return ORZ_OK;
}
extern "C" int orz_context_set_cluster_size(orz_context* ctx, int ctas) {
  if (!ctx || (ctas != 0 && ctas != 1 && ctas != 2 && ctas != 4 && ctas != 8 && ctas != 16)) return fail(ORZ_ERR_ARG, "cluster size must be 0, 1, 2, 4, 8 or 16");
  ctx->clusterSize = ctas;
  return ORZ_OK;
}
extern "C" int orz_context_set_arena_bytes(orz_context* ctx, size_t bytes) {
  if (!ctx) return fail(ORZ_ERR_ARG, "null context");
  ctx->arenaBudget = bytes;
  return ORZ_OK;
}
extern "C" int orz_context_set_rcp_table(orz_context* ctx, const uint32_t* table, int bits) {
  if (!ctx || !table || bits < 1 || bits > 23) return fail(ORZ_ERR_ARG, "orz_context_set_rcp_table: bad arguments");
  ORZ_CUDA(cudaSetDevice(ctx->device));
  ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->h_rcp.assign(table, table + (size_t(1) << bits));
  ctx->rcpBits = bits;
  cudaFree(ctx->d_rcp);
  ORZ_CUDA(cudaMalloc(&ctx->d_rcp, ctx->h_rcp.size() * 4));
  ORZ_CUDA(cudaMemcpy(ctx->d_rcp, ctx->h_rcp.data(), ctx->h_rcp.size() * 4, cudaMemcpyHostToDevice));
  return ORZ_OK;
}
extern "C" int orz_context_get_rcp_table(orz_context* ctx, uint32_t* table, int* bits) {
  if (!ctx) return fail(ORZ_ERR_ARG, "null context");
  if (bits) *bits = ctx->rcpBits;
  if (table) memcpy(table, ctx->h_rcp.data(), ctx->h_rcp.size() * 4);
  return ORZ_OK;
}
extern "C" int orz_context_get_lut(orz_context* ctx, int64_t* lut4096) {
  if (!ctx || !lut4096) return fail(ORZ_ERR_ARG, "bad arguments");
  memcpy(lut4096, edge_mask_table(), 4096 * 8);
  return ORZ_OK;
}

// reference packet layout (Occluder.cpp:146-156) -> one uint4 (v0..v3) per quad
static void relayout_packets(const uint32_t* packets, uint32_t packetCount, uint4* out) {
  const uint32_t nQuads = packetCount * 2;
  for (uint32_t q = 0; q < nQuads; ++q) {
    const uint32_t g = q >> 3, l = q & 7;
    const uint32_t* base = packets + (size_t)g * 32 + l;
    out[q] = make_uint4(base[0], base[8], base[16], base[24]);
  }
}

extern "C" int orz_occluder_create(orz_context* ctx, const uint32_t* packets, uint32_t packetCount, const float* refMin4,
                                   const float* refMax4, orz_occluder** out) {
  if (!ctx || !packets || !out || packetCount % 4 != 0) return fail(ORZ_ERR_ARG, "orz_occluder_create: bad arguments");
  ORZ_CUDA(cudaSetDevice(ctx->device));
  orz_occluder* o = new orz_occluder();
  o->ctx = ctx;
  o->nQuads = packetCount * 2;
  memcpy(o->refMin, refMin4, 16);
  memcpy(o->refMax, refMax4, 16);
  std::vector<uint4> tmp(o->nQuads);
  relayout_packets(packets, packetCount, tmp.data());
  o->d_quads = nullptr;
  if (o->nQuads) {
    ORZ_CUDA(cudaMalloc(&o->d_quads, tmp.size() * sizeof(uint4)));
    ORZ_CUDA(cudaMemcpy(o->d_quads, tmp.data(), tmp.size() * sizeof(uint4), cudaMemcpyHostToDevice));
  }
  *out = o;
  return ORZ_OK;
}
extern "C" void orz_occluder_destroy(orz_occluder* o) {
  if (!o) return;
  cudaSetDevice(o->ctx->device);
  cudaStreamSynchronize(o->ctx->stream);
  cudaFree(o->d_quads);
  delete o;
}

extern "C" int orz_rasterizer_create(orz_context* ctx, uint32_t width, uint32_t height, orz_rasterizer** out) {
  if (!ctx || !out || width == 0 || height == 0 || width % 8 || height % 8 || width > 65535u * 8u || height > 65535u * 8u)
    return fail(ORZ_ERR_ARG, "orz_rasterizer_create: width and height must be positive multiples of 8");  // Rasterizer.cpp:68
  ORZ_CUDA(cudaSetDevice(ctx->device));
  orz_rasterizer* r = new orz_rasterizer();
  r->ctx = ctx;
  r->T.width = width; r->T.height = height; r->T.blocksX = width / 8; r->T.blocksY = height / 8;
  const size_t blocks = (size_t)r->T.blocksX * r->T.blocksY;
  ORZ_CUDA(cudaMalloc(&r->T.depth, blocks * 128));
  ORZ_CUDA(cudaMalloc(&r->T.hiz, (blocks + 8) * 2));
  ORZ_CUDA(cudaMemsetAsync(r->T.depth, 0, blocks * 128, ctx->stream));
  ORZ_CUDA(cudaMemsetAsync(r->T.hiz, 0, (blocks + 8) * 2, ctx->stream));  // Rasterizer.cpp:71: HiZ starts at 0 until clear()
  memset(&r->vm, 0, sizeof r->vm);
  *out = r;
  return ORZ_OK;
}
extern "C" void orz_rasterizer_destroy(orz_rasterizer* r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  cudaStreamSynchronize(r->ctx->stream);
  cudaFree(r->T.depth); cudaFree(r->T.hiz); cudaFree(r->d_boxOut);
  delete r;
}
extern "C" int orz_rasterizer_set_mvp(orz_rasterizer* r, const float* m) {
  if (!r || !m) return fail(ORZ_ERR_ARG, "orz_rasterizer_set_mvp: bad arguments");
  bake_view_matrices(m, r->T.width, r->T.height, r->vm);
  return ORZ_OK;
}
extern "C" int orz_rasterizer_clear(orz_rasterizer* r) {
  if (!r) return fail(ORZ_ERR_ARG, "null rasterizer");
  ORZ_CUDA(cudaSetDevice(r->ctx->device));
  const uint32_t blocks = r->T.blocksX * r->T.blocksY;
  k_clear<<<r->ctx->numSMs * 2, 256, 0, r->ctx->stream>>>(r->T.depth, r->T.hiz, blocks);
  r->ctx->launches++;
  ORZ_CUDA(cudaGetLastError());
  return ORZ_OK;
}
extern "C" int orz_rasterizer_rasterize(orz_rasterizer* r, const orz_occluder* occ, int clipped) {
  if (!r || !occ) return fail(ORZ_ERR_ARG, "orz_rasterizer_rasterize: bad arguments");
  if (occ->nQuads == 0) return ORZ_OK;
  ORZ_CUDA(cudaSetDevice(r->ctx->device));
  constexpr int GW = 4;
  const uint32_t grid = (r->T.blocksY + GW - 1) / GW;  // one screen block-row per warp
  const float4 mn = make_float4(occ->refMin[0], occ->refMin[1], occ->refMin[2], occ->refMin[3]);
  const float4 mx = make_float4(occ->refMax[0], occ->refMax[1], occ->refMax[2], occ->refMax[3]);
  k_rasterize_single<GW><<<grid, GW * 32, 0, r->ctx->stream>>>(r->vm, occ->d_quads, occ->nQuads, mn, mx, clipped, r->T,
                                                                 r->ctx->d_rcp, 23 - r->ctx->rcpBits, r->ctx->d_lut);
  r->ctx->launches++;
  ORZ_CUDA(cudaGetLastError());
  return ORZ_OK;
}
extern "C" int orz_rasterizer_query2d(orz_rasterizer* r, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ,
                                      int* visible) {
  if (!r || !visible) return fail(ORZ_ERR_ARG, "orz_rasterizer_query2d: bad arguments");
  if (maxX >= r->T.width || maxY >= r->T.height || minX > maxX || minY > maxY) return fail(ORZ_ERR_ARG, "orz_rasterizer_query2d: rectangle outside the buffer");
  ORZ_CUDA(cudaSetDevice(r->ctx->device));
  k_query2d<<<1, 256, 0, r->ctx->stream>>>(r->T, minX, maxX, minY, maxY, maxZ, r->ctx->d_counter + 8);
  r->ctx->launches++;
  ORZ_CUDA(cudaMemcpyAsync(r->ctx->h_pinned, r->ctx->d_counter + 8, 4, cudaMemcpyDeviceToHost, r->ctx->stream));
  ORZ_CUDA(cudaStreamSynchronize(r->ctx->stream));
  *visible = r->ctx->h_pinned[0] ? 1 : 0;
  return ORZ_OK;
}
extern "C" int orz_rasterizer_query_visibility(orz_rasterizer* r, const float* bmin, const float* bmax, int* visible, int* needsClipping) {
  if (!r || !bmin || !bmax || !visible) return fail(ORZ_ERR_ARG, "orz_rasterizer_query_visibility: bad arguments");
  // the front half is state independent and scalar: evaluate it here with the same core the
  // kernels use, then ask the GPU only for the rectangle test (Rasterizer.cpp:275)
  const RcpTable rt{r->ctx->h_rcp.data(), 23 - r->ctx->rcpBits};
  const BoxFront f = box_front_half(r->vm, bmin, bmax, r->T.width, r->T.height, rt);
  if (f.status == kBoxCulled) { *visible = 0; return ORZ_OK; }  // needsClipping untouched, as in the reference
  if (f.status == kBoxNearClip) { *visible = 1; if (needsClipping) *needsClipping = 1; return ORZ_OK; }
  if (needsClipping) *needsClipping = 0;
  return orz_rasterizer_query2d(r, f.minX, f.maxX, f.minY, f.maxY, f.maxZ, visible);
}
extern "C" int orz_rasterizer_query_boxes(orz_rasterizer* r, const float* boxes, uint32_t n, uint8_t* out) {
  if (!r || (n && (!boxes || !out))) return fail(ORZ_ERR_ARG, "orz_rasterizer_query_boxes: bad arguments");
  if (n == 0) return ORZ_OK;
  orz_context* ctx = r->ctx;
  ORZ_CUDA(cudaSetDevice(ctx->device));
  if (int e = ensure_scratch(ctx, 0, (size_t)n * 32)) return e;
  if (int e = ensure_scratch(ctx, 1, n)) return e;
  ORZ_CUDA(cudaMemcpyAsync(ctx->d_scratch[0], boxes, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
  k_query_boxes<<<(n + 127) / 128, 128, 0, ctx->stream>>>(r->vm, (const float4*)ctx->d_scratch[0], n, r->T, ctx->d_rcp, 23 - ctx->rcpBits,
                                                          (uint8_t*)ctx->d_scratch[1]);
  ctx->launches++;
  ORZ_CUDA(cudaGetLastError());
  ORZ_CUDA(cudaMemcpyAsync(out, ctx->d_scratch[1], n, cudaMemcpyDeviceToHost, ctx->stream));
  ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return ORZ_OK;
}
extern "C" int orz_rasterizer_readback_depth(orz_rasterizer* r, void* target) {
  if (!r || !target) return fail(ORZ_ERR_ARG, "orz_rasterizer_readback_depth: bad arguments");
  orz_context* ctx = r->ctx;
  ORZ_CUDA(cudaSetDevice(ctx->device));
  const size_t px = (size_t)r->T.width * r->T.height;
  if (int e = ensure_scratch(ctx, 2, px * 4)) return e;
  k_readback<<<(uint32_t)((px + 255) / 256), 256, 0, ctx->stream>>>(r->T, (uint8_t*)ctx->d_scratch[2]);
  ctx->launches++;
  ORZ_CUDA(cudaGetLastError());
  ORZ_CUDA(cudaMemcpyAsync(target, ctx->d_scratch[2], px * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return ORZ_OK;
}
extern "C" int orz_rasterizer_download(orz_rasterizer* r, uint16_t* depth, uint16_t* hiz) {
  if (!r) return fail(ORZ_ERR_ARG, "null rasterizer");
  orz_context* ctx = r->ctx;
  ORZ_CUDA(cudaSetDevice(ctx->device));
  const uint32_t blocks = r->T.blocksX * r->T.blocksY;
  if (depth) {
    if (int e = ensure_scratch(ctx, 2, (size_t)blocks * 128)) return e;
    k_canonical_depth<<<(blocks * 8 + 255) / 256, 256, 0, ctx->stream>>>(r->T.depth, r->T.hiz, blocks, (uint4*)ctx->d_scratch[2]);
    ctx->launches++;
    ORZ_CUDA(cudaGetLastError());
    ORZ_CUDA(cudaMemcpyAsync(depth, ctx->d_scratch[2], (size_t)blocks * 128, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (hiz) ORZ_CUDA(cudaMemcpyAsync(hiz, r->T.hiz, (size_t)blocks * 2, cudaMemcpyDeviceToHost, ctx->stream));
  ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return ORZ_OK;
}
extern "C" int orz_rasterizer_debug_setup(orz_rasterizer* r, const orz_occluder* occ, int clipped, orz_prim_record* out) {
  if (!r || !occ || !out) return fail(ORZ_ERR_ARG, "orz_rasterizer_debug_setup: bad arguments");
  orz_context* ctx = r->ctx;
  ORZ_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)occ->nQuads * sizeof(orz_prim_record);
  if (int e = ensure_scratch(ctx, 3, bytes)) return e;
  const float4 mn = make_float4(occ->refMin[0], occ->refMin[1], occ->refMin[2], occ->refMin[3]);
  const float4 mx = make_float4(occ->refMax[0], occ->refMax[1], occ->refMax[2], occ->refMax[3]);
  k_debug_setup<<<(occ->nQuads + 127) / 128, 128, 0, ctx->stream>>>(r->vm, occ->d_quads, occ->nQuads, mn, mx, clipped, r->T, ctx->d_rcp,
                                                                    23 - ctx->rcpBits, (orz_prim_record*)ctx->d_scratch[3]);
  ctx->launches++;
  ORZ_CUDA(cudaGetLastError());
  ORZ_CUDA(cudaMemcpyAsync(out, ctx->d_scratch[3], bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return ORZ_OK;
}

// ---- scenes + view batches ---------------------------------------------------------------------
extern "C" int orz_scene_create(orz_context* ctx, const uint32_t* packets, const uint32_t* packetCounts, uint32_t nOcc,
                                const float* refMin, const float* refMax, const float* boundsMin, const float* boundsMax,
                                const float* centers, orz_scene** out) {
  if (!ctx || !packets || !packetCounts || !refMin || !refMax || !boundsMin || !boundsMax || !centers || !out || nOcc == 0)
    return fail(ORZ_ERR_ARG, "orz_scene_create: bad arguments");
  ORZ_CUDA(cudaSetDevice(ctx->device));
  orz_scene* s = new orz_scene();
  s->ctx = ctx;
  s->nOcc = nOcc;
  std::vector<OccMeta> meta(nOcc);
  size_t totalPackets = 0;
  for (uint32_t i = 0; i < nOcc; ++i) {
    if (packetCounts[i] % 4 != 0) { delete s; return fail(ORZ_ERR_ARG, "orz_scene_create: packet counts must be multiples of 4"); }
    meta[i].quadOffset = (uint32_t)(totalPackets * 2);
    meta[i].quadCount = packetCounts[i] * 2;
    meta[i].pad0 = meta[i].pad1 = 0;
    memcpy(meta[i].refMin, refMin + 4 * i, 16); memcpy(meta[i].refMax, refMax + 4 * i, 16);
    memcpy(meta[i].boundsMin, boundsMin + 4 * i, 16); memcpy(meta[i].boundsMax, boundsMax + 4 * i, 16);
    memcpy(meta[i].center, centers + 4 * i, 16);
    totalPackets += packetCounts[i];
  }
  s->totalQuads = (uint32_t)(totalPackets * 2);
  std::vector<uint4> quads(s->totalQuads);
  size_t pofs = 0;
  for (uint32_t i = 0; i < nOcc; ++i) {
    relayout_packets(packets + pofs * 8, packetCounts[i], quads.data() + meta[i].quadOffset);
    pofs += packetCounts[i];
  }
  ORZ_CUDA(cudaMalloc(&s->d_quads, std::max<size_t>(quads.size(), 1) * sizeof(uint4)));
  ORZ_CUDA(cudaMemcpy(s->d_quads, quads.data(), quads.size() * sizeof(uint4), cudaMemcpyHostToDevice));
  ORZ_CUDA(cudaMalloc(&s->d_occ, meta.size() * sizeof(OccMeta)));
  ORZ_CUDA(cudaMemcpy(s->d_occ, meta.data(), meta.size() * sizeof(OccMeta), cudaMemcpyHostToDevice));
  *out = s;
  return ORZ_OK;
}
// Occluder::bake for every batch of a scene on the GPU; the scene is created directly in HBM
extern "C" int orz_scene_bake(orz_context* ctx, const float* vertices, const uint32_t* vertCounts, uint32_t nOcc, const float* refMin4,
                              const float* refMax4, uint32_t* packetsOut, float* centersOut, float* boundsMinOut, float* boundsMaxOut,
                              orz_scene** out) {
  if (!ctx || !vertices || !vertCounts || !refMin4 || !refMax4 || !out || nOcc == 0) return fail(ORZ_ERR_ARG, "orz_scene_bake: bad arguments");
  ORZ_CUDA(cudaSetDevice(ctx->device));
  std::vector<BakeJob> jobs(nOcc);
  size_t totalVerts = 0;
  uint32_t maxQuads = 0;
  for (uint32_t i = 0; i < nOcc; ++i) {
    if (vertCounts[i] % 32 != 0) return fail(ORZ_ERR_ARG, "orz_scene_bake: every batch needs a multiple of 8 quads (32 vertices)");  // Occluder.cpp:108
    jobs[i].vertOffset = (uint32_t)totalVerts; jobs[i].nQuads = vertCounts[i] / 4; jobs[i].quadOffset = (uint32_t)(totalVerts / 4); jobs[i].pad = 0;
    maxQuads = std::max(maxQuads, jobs[i].nQuads);
    totalVerts += vertCounts[i];
  }
  const size_t smem = (size_t)maxQuads * 20;  // 3 floats + cluster + slot per quad
  if (smem > 200 * 1024) return fail(ORZ_ERR_ARG, "orz_scene_bake: batch too large for the device bake (max 10 240 quads); use orz_bake");
  // rsqrtps model: the installed table (orz_set_rsqrt_table), else this CPU's
  std::vector<uint32_t> rsq;
  int rsqBits = 0;
  current_rsqrt_table(rsq, rsqBits);
  orz_scene* s = new orz_scene();
  s->ctx = ctx; s->nOcc = nOcc; s->totalQuads = (uint32_t)(totalVerts / 4);
  float4* d_verts = nullptr; BakeJob* d_jobs = nullptr; uint32_t* d_rsq = nullptr; uint32_t* d_packets = nullptr;
  auto cleanup = [&]() { cudaFree(d_verts); cudaFree(d_jobs); cudaFree(d_rsq); cudaFree(d_packets); };
#define ORZ_BAKE_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { cleanup(); cudaFree(s->d_quads); cudaFree(s->d_occ); delete s; \
    return fail(ORZ_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); } } while (0)
  ORZ_BAKE_CUDA(cudaMalloc(&d_verts, totalVerts * 16));
  ORZ_BAKE_CUDA(cudaMalloc(&d_jobs, nOcc * sizeof(BakeJob)));
  ORZ_BAKE_CUDA(cudaMalloc(&d_rsq, rsq.size() * 4));
  ORZ_BAKE_CUDA(cudaMalloc(&s->d_quads, std::max<size_t>(totalVerts / 4, 1) * sizeof(uint4)));
  ORZ_BAKE_CUDA(cudaMalloc(&s->d_occ, nOcc * sizeof(OccMeta)));
  if (packetsOut) ORZ_BAKE_CUDA(cudaMalloc(&d_packets, totalVerts * 4));
  ORZ_BAKE_CUDA(cudaMemcpyAsync(d_verts, vertices, totalVerts * 16, cudaMemcpyHostToDevice, ctx->stream));
  ORZ_BAKE_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), nOcc * sizeof(BakeJob), cudaMemcpyHostToDevice, ctx->stream));
  ORZ_BAKE_CUDA(cudaMemcpyAsync(d_rsq, rsq.data(), rsq.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (smem > 48 * 1024) ORZ_BAKE_CUDA(cudaFuncSetAttribute(k_bake, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const RsqrtTable rt{d_rsq, rsqBits};
  k_bake<<<nOcc, 256, smem, ctx->stream>>>(d_verts, d_jobs, make_float4(refMin4[0], refMin4[1], refMin4[2], refMin4[3]),
                                          make_float4(refMax4[0], refMax4[1], refMax4[2], refMax4[3]), rt, s->d_quads, s->d_occ, d_packets);
  ctx->launches++;
  ORZ_BAKE_CUDA(cudaGetLastError());
  if (packetsOut) ORZ_BAKE_CUDA(cudaMemcpyAsync(packetsOut, d_packets, totalVerts * 4, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<OccMeta> meta;
  if (centersOut || boundsMinOut || boundsMaxOut) {
    meta.resize(nOcc);
    ORZ_BAKE_CUDA(cudaMemcpyAsync(meta.data(), s->d_occ, nOcc * sizeof(OccMeta), cudaMemcpyDeviceToHost, ctx->stream));
  }
  ORZ_BAKE_CUDA(cudaStreamSynchronize(ctx->stream));
#undef ORZ_BAKE_CUDA
  for (uint32_t i = 0; i < nOcc && !meta.empty(); ++i) {
    if (centersOut) memcpy(centersOut + 4 * i, meta[i].center, 16);
    if (boundsMinOut) memcpy(boundsMinOut + 4 * i, meta[i].boundsMin, 16);
    if (boundsMaxOut) memcpy(boundsMaxOut + 4 * i, meta[i].boundsMax, 16);
  }
  cleanup();
  *out = s;
  return ORZ_OK;
}
extern "C" int orz_scene_set_occludees(orz_scene* s, const float* boxes, uint32_t n) {
  if (!s || (n && !boxes)) return fail(ORZ_ERR_ARG, "orz_scene_set_occludees: bad arguments");
  ORZ_CUDA(cudaSetDevice(s->ctx->device));
  ORZ_CUDA(cudaStreamSynchronize(s->ctx->stream));
  cudaFree(s->d_boxes);
  s->d_boxes = nullptr;
  s->nBoxes = n;
  if (n) {
    ORZ_CUDA(cudaMalloc(&s->d_boxes, (size_t)n * 32));
    ORZ_CUDA(cudaMemcpy(s->d_boxes, boxes, (size_t)n * 32, cudaMemcpyHostToDevice));
  }
  return ORZ_OK;
}
extern "C" void orz_scene_destroy(orz_scene* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->d_quads); cudaFree(s->d_occ); cudaFree(s->d_boxes);
  delete s;
}

template <int GW, int kTrav>
static int launch_views_t(orz_context* ctx, const FrameParams& p, uint32_t grid, cudaStream_t st) {
  static bool configured[64] = {false};
  if (!configured[ctx->device & 63]) {
    ORZ_CUDA(cudaFuncSetAttribute(k_render_views<GW, kTrav>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FrameSmem<GW, kTrav>::kBytes));
    configured[ctx->device & 63] = true;
  }
  k_render_views<GW, kTrav><<<grid, GW * 32, FrameSmem<GW, kTrav>::kBytes, st>>>(p);
  ctx->launches++;
  ORZ_CUDA(cudaGetLastError());
  return ORZ_OK;
}
template <int GW, int kTrav>
static int occupancy_views_t(int* perSM) {
  ORZ_CUDA(cudaFuncSetAttribute(k_render_views<GW, kTrav>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FrameSmem<GW, kTrav>::kBytes));
  ORZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(perSM, k_render_views<GW, kTrav>, GW * 32, FrameSmem<GW, kTrav>::kBytes));
  return ORZ_OK;
}
static int launch_views(orz_context* ctx, int GW, int trav, const FrameParams& p, uint32_t grid, cudaStream_t st) {
  if (trav == 2) {
    switch (GW) {
      case 1: return launch_views_t<1, 2>(ctx, p, grid, st);
      case 2: return launch_views_t<2, 2>(ctx, p, grid, st);
      case 4: return launch_views_t<4, 2>(ctx, p, grid, st);
      default: return launch_views_t<8, 2>(ctx, p, grid, st);
    }
  }
  switch (GW) {
    case 1: return launch_views_t<1, 1>(ctx, p, grid, st);
    case 2: return launch_views_t<2, 1>(ctx, p, grid, st);
    case 4: return launch_views_t<4, 1>(ctx, p, grid, st);
    case 8: return launch_views_t<8, 1>(ctx, p, grid, st);
    default: return launch_views_t<16, 1>(ctx, p, grid, st);
  }
}
static int occupancy_views(int GW, int trav, int* perSM) {
  if (trav == 2) {
    switch (GW) {
      case 1: return occupancy_views_t<1, 2>(perSM);
      case 2: return occupancy_views_t<2, 2>(perSM);
      case 4: return occupancy_views_t<4, 2>(perSM);
      default: return occupancy_views_t<8, 2>(perSM);
    }
  }
  switch (GW) {
    case 1: return occupancy_views_t<1, 1>(perSM);
    case 2: return occupancy_views_t<2, 1>(perSM);
    case 4: return occupancy_views_t<4, 1>(perSM);
    case 8: return occupancy_views_t<8, 1>(perSM);
    default: return occupancy_views_t<16, 1>(perSM);
  }
}

template <int C>
static int launch_cluster_t(orz_context* ctx, FrameParams p, uint32_t nViews, uint32_t nTiles, cudaStream_t st) {
  p.clusterK = (nTiles + (uint32_t)(C * kClusterGW) - 1u) / (uint32_t)(C * kClusterGW);
  const size_t smem = ClusterSmem::bytes(p.clusterK, p.nOcc);
  static size_t configured[64] = {0};
  if (configured[ctx->device & 63] < smem) {
    ORZ_CUDA(cudaFuncSetAttribute(k_raster_views_cluster<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (C > 8) ORZ_CUDA(cudaFuncSetAttribute(k_raster_views_cluster<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured[ctx->device & 63] = smem;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(nViews * (uint32_t)C);
  cfg.blockDim = dim3(kClusterGW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  ORZ_CUDA(cudaLaunchKernelEx(&cfg, k_raster_views_cluster<C>, p));
  ctx->launches++;
  return ORZ_OK;
}
// one cluster per view.  Cluster size: as many CTAs as the view can use (one tile per warp) while all
// views of the batch still fit the GPU in one wave; at least enough that a warp owns <= 32 tiles.
static int launch_cluster(orz_context* ctx, const FrameParams& p, uint32_t nViews, uint32_t nBatch, cudaStream_t st) {
  const uint32_t nTiles = (((p.width >> 3) + kTileW - 1u) / kTileW) * (((p.height >> 3) + kTileH - 1u) / kTileH);
  uint32_t c = 1;
  while (c < 16u && c * kClusterGW < nTiles && nBatch * c * 2u <= (uint32_t)ctx->numSMs) c *= 2u;
  while (c < 16u && (nTiles + c * kClusterGW - 1u) / (c * kClusterGW) > 32u) c *= 2u;
  if (ctx->clusterSize) c = (uint32_t)ctx->clusterSize;
  if ((nTiles + c * kClusterGW - 1u) / (c * kClusterGW) > 32u) return fail(ORZ_ERR_ARG, "cluster path: target too large for this cluster size");
  switch (c) {
    case 16:  // non-portable cluster size: when the device (e.g. a partitioned one) cannot place it, use 8
      if (launch_cluster_t<16>(ctx, p, nViews, nTiles, st) == ORZ_OK) return ORZ_OK;
      (void)cudaGetLastError();
      if ((nTiles + 8u * kClusterGW - 1u) / (8u * kClusterGW) > 32u) return fail(ORZ_ERR_CUDA, "cluster path: 16-CTA clusters are not available on this device");
      return launch_cluster_t<8>(ctx, p, nViews, nTiles, st);
    case 8: return launch_cluster_t<8>(ctx, p, nViews, nTiles, st);
    case 4: return launch_cluster_t<4>(ctx, p, nViews, nTiles, st);
    case 2: return launch_cluster_t<2>(ctx, p, nViews, nTiles, st);
    default: return launch_cluster_t<1>(ctx, p, nViews, nTiles, st);
  }
}

// Device-pointer entry: three launches per chunk of views (prepare, render, query).  When the
// caller does not ask for depth/HiZ, per-view targets live in an internal arena and the batch is
// processed in chunks that fit the arena budget.
extern "C" int orz_render_views_device(orz_context* ctx, orz_scene* scene, const orz_view_batch* b) {
  if (!ctx || !scene || !b || !b->mvps || (!b->orders && !b->camPos)) return fail(ORZ_ERR_ARG, "orz_render_views_device: bad arguments");
  if (b->width == 0 || b->height == 0 || b->width % 8 || b->height % 8) return fail(ORZ_ERR_ARG, "width and height must be positive multiples of 8");
  if ((b->depth != nullptr) != (b->hiz != nullptr)) return fail(ORZ_ERR_ARG, "depth and hiz outputs must be requested together");
  if (b->nViews == 0) return ORZ_OK;
  ORZ_CUDA(cudaSetDevice(ctx->device));
  // warps per view: enough warps in total to fill the machine (few views -> more warps each)
  int GW = ctx->groupWarps;
  if (!GW) {
    if (ctx->traversal == 2) {  // lane-per-block: fewer, fuller warps (measured on Castle 1080p, 1k-4k views)
      const uint32_t want = (uint32_t)ctx->numSMs * 60u / std::max<uint32_t>(b->nViews, 1u);
      GW = want >= 16 ? 8 : want >= 6 ? 4 : want >= 2 ? 2 : 1;
    } else {
      const uint32_t want = (uint32_t)ctx->numSMs * 110u / std::max<uint32_t>(b->nViews, 1u);  // ~16k warps on 148 SMs
      GW = want >= 8 ? 8 : want >= 4 ? 4 : want >= 2 ? 2 : 1;
    }
    if (b->height < 512 && GW > 1) GW /= 2;  // few block rows: less row parallelism to hand out
  }
  int perSM = 0;
  const int trav = ctx->traversal == 2 ? 2 : 1;
  if (trav == 2 && GW > 8) GW = 8;
  int e = occupancy_views(GW, trav, &perSM);
  if (e) return e;
  if (perSM < 1) perSM = 1;
  const size_t blocks = (size_t)(b->width / 8) * (b->height / 8);
  const size_t nOcc = scene->nOcc;
  const size_t hizStride = (blocks + 7) & ~size_t(7);
  const bool ownTargets = b->depth == nullptr;
  if (!ownTargets && blocks % 8 != 0) return fail(ORZ_ERR_ARG, "per-view depth output needs (w/8)*(h/8) to be a multiple of 8");

  // views per chunk: everything when the caller owns the targets, else what fits the arena budget
  size_t chunk = b->nViews;
  if (ownTargets) {
    const size_t perView = blocks * 128 + hizStride * 2;
    const size_t budget = std::max<size_t>(ctx->arenaBudget, perView);
    chunk = std::max<size_t>(1, std::min<size_t>(b->nViews, budget / perView));
    if ((e = ensure_scratch(ctx, 4, chunk * blocks * 128))) return e;
    if ((e = ensure_scratch(ctx, 5, chunk * hizStride * 2))) return e;
  }
  if ((e = ensure_scratch(ctx, 6, chunk * sizeof(ViewMatrices) + chunk * nOcc * 4 + chunk * nOcc * kFrontWords * 4 + chunk * 8))) return e;
  uint8_t* prep = (uint8_t*)ctx->d_scratch[6];

  const size_t bitWords = (scene->nBoxes + 31) / 32;
  for (size_t v0 = 0; v0 < b->nViews; v0 += chunk) {
    const uint32_t nv = (uint32_t)std::min<size_t>(chunk, b->nViews - v0);
    FrameParams p;
    memset(&p, 0, sizeof p);
    p.quads = scene->d_quads; p.occ = scene->d_occ; p.nOcc = scene->nOcc;
    p.boxes = scene->d_boxes; p.nBoxes = scene->nBoxes;
    p.rcp = ctx->d_rcp; p.rcpShift = 23 - ctx->rcpBits; p.lut = ctx->d_lut;
    p.width = b->width; p.height = b->height; p.nViews = nv; p.flags = b->flags;
    p.mvps = b->mvps + 16 * v0;
    p.orders = b->orders ? b->orders + v0 * nOcc : nullptr;
    p.camPos = b->camPos ? b->camPos + 3 * v0 : nullptr;
    p.vmBuf = (ViewMatrices*)prep;
    p.orderBuf = (uint32_t*)(prep + chunk * sizeof(ViewMatrices));
    p.frontBuf = (uint32_t*)(prep + chunk * sizeof(ViewMatrices) + chunk * nOcc * 4);
    p.bitWords = (uint32_t)bitWords;
    p.visBits = (scene->nBoxes && b->visBits) ? b->visBits + v0 * bitWords : nullptr;
    p.clipBits = (scene->nBoxes && b->clipBits) ? b->clipBits + v0 * bitWords : nullptr;
    p.gate = b->gate ? b->gate + v0 * nOcc : nullptr;
    p.quadsSubmitted = b->quadsSubmitted ? b->quadsSubmitted + v0 : nullptr;
    p.exportDepth = ownTargets ? 0 : 1;
    if (ownTargets) {
      p.depth = (uint16_t*)ctx->d_scratch[4]; p.hiz = (uint16_t*)ctx->d_scratch[5];
      p.depthStride = blocks * 64; p.hizStride = hizStride;
    } else {
      p.depth = b->depth + v0 * blocks * 64; p.hiz = b->hiz + v0 * blocks;
      p.depthStride = blocks * 64; p.hizStride = blocks;
    }
    p.viewCounter = ctx->d_counter;
    p.viewCost = (uint32_t*)(prep + chunk * sizeof(ViewMatrices) + chunk * nOcc * 4 + chunk * nOcc * kFrontWords * 4);
    const bool wide = (b->flags & ORZ_BATCH_NO_GATE) && ((b->flags & ORZ_BATCH_WIDE) || (nv <= 8u && scene->totalQuads >= 65536u));
    // (above 65 536 blocks the reference's 16-bit index wrap needs the linear traversal of the batch kernel)
    // measured crossover with the batch kernel: ~2000 views at 1920x1080, ~1500 at 512x256 (profiles/r1_few_views_*)
    const uint32_t clusterLimit = (uint32_t)ctx->clusterViews;
    const bool clusterPath = !wide && ctx->clusterViews > 0 && nv <= clusterLimit && blocks <= 65536u && nOcc <= kClusterMaxOcc &&
                             (size_t)nv * scene->totalQuads * (kRecStride * 4 + 8) <= (size_t(8) << 30);
    p.viewOrder = (nv <= 16384u && !(clusterPath && nv * 2u <= (uint32_t)ctx->numSMs)) ? p.viewCost + chunk : nullptr;
    k_prepare_views<<<nv, 128, 0, ctx->stream>>>(p);
    ctx->launches++;
    ORZ_CUDA(cudaGetLastError());
    if (p.viewOrder) {
      k_sort_views<<<(nv + 255) / 256, 256, 0, ctx->stream>>>(p.viewCost, nv, p.viewOrder);
      ctx->launches++;
      ORZ_CUDA(cudaGetLastError());
    }
    // Few views over many ungated occluders: split each view over the whole GPU instead
    if (wide) {
      const uint32_t total = scene->totalQuads, nChunks = (total + 31u) / 32u;
      if ((e = ensure_scratch(ctx, 8, ((size_t)nOcc + 1) * 4))) return e;
      if ((e = ensure_scratch(ctx, 9, (size_t)nChunks * 32 * kWideRecWords * 4))) return e;
      if ((e = ensure_scratch(ctx, 10, (size_t)nChunks * 8))) return e;
      uint32_t* slotStart = (uint32_t*)ctx->d_scratch[8];
      uint32_t* recs = (uint32_t*)ctx->d_scratch[9];
      uint32_t* chunkCount = (uint32_t*)ctx->d_scratch[10];
      uint32_t* chunkRows = chunkCount + nChunks;
      for (uint32_t v = 0; v < nv; ++v) {
        Target T;
        T.width = b->width; T.height = b->height; T.blocksX = b->width / 8; T.blocksY = b->height / 8;
        T.depth = p.depth + (size_t)v * p.depthStride;
        T.hiz = p.hiz + (size_t)v * p.hizStride;
        k_clear_hiz<<<ctx->numSMs, 256, 0, ctx->stream>>>(T.hiz, (uint32_t)blocks);
        k_slot_prefix<<<1, 1024, 0, ctx->stream>>>(p, v, slotStart);
        k_setup_wide<<<(total + 255) / 256, 256, 0, ctx->stream>>>(p, v, slotStart, total, recs, chunkCount, chunkRows);
        // rows x column segments: about 16 warps per SM
        uint32_t nSeg = std::max<uint32_t>(1u, std::min<uint32_t>(T.blocksX / 32u, (uint32_t)ctx->numSMs * 16u / T.blocksY));
        const uint32_t segWidth = (T.blocksX + nSeg - 1) / nSeg;
        nSeg = (T.blocksX + segWidth - 1) / segWidth;
        k_raster_wide<<<(T.blocksY * nSeg + 3) / 4, 128, 0, ctx->stream>>>(T, ctx->d_lut, recs, chunkCount, chunkRows, nChunks, nSeg, segWidth);
        ctx->launches += 4;
        if (p.exportDepth) { k_zero_cleared<<<ctx->numSMs, 256, 0, ctx->stream>>>(T.depth, T.hiz, (uint32_t)blocks); ctx->launches++; }
        ORZ_CUDA(cudaGetLastError());
      }
      if (p.visBits || p.clipBits) {
        FrameParams pq = p;
        pq.viewOrder = nullptr; pq.viewBase = 0; pq.groupViews = nv;
        k_query_views<<<dim3((scene->nBoxes + 255) / 256, nv), 256, 0, ctx->stream>>>(pq);
        ctx->launches++;
        ORZ_CUDA(cudaGetLastError());
      }
      continue;
    }
    // Few views: one thread-block cluster per view (latency path, BASELINE configs 1 and 2)
    if (clusterPath) {
      FrameParams pc = p;
      pc.viewBase = 0; pc.groupViews = nv;
      const size_t recBytes = (size_t)nv * scene->totalQuads * kRecStride * 4, hdrBytes = (size_t)nv * scene->totalQuads * 8;
      if ((e = ensure_scratch(ctx, 11, recBytes + hdrBytes + (size_t)nv * nOcc * 32 + 64))) return e;
      pc.hdrBuf = (uint2*)ctx->d_scratch[11];
      pc.recInfo = (uint4*)((uint8_t*)ctx->d_scratch[11] + ((hdrBytes + 15) & ~size_t(15)));
      pc.recBuf = (uint32_t*)((uint8_t*)pc.recInfo + (size_t)nv * nOcc * 32);
      pc.totalQuads = scene->totalQuads;
      k_setup_views<<<dim3(pc.nOcc, nv), 256, 0, ctx->stream>>>(pc);
      ctx->launches++;
      ORZ_CUDA(cudaGetLastError());
      // Large batches: sub-batches (by descending cost) on auxiliary streams, so that the occludee queries of a
      // finished sub-batch share the SMs with the cluster kernel of the next one (it leaves room for one
      // query CTA per SM and about half of its issue slots).
      const bool wantQ = p.visBits || p.clipBits;
      const int groupsC = (pc.viewOrder && wantQ && nv >= 256u) ? orz_context::kGroups : 1;
      if (groupsC > 1) ORZ_CUDA(cudaEventRecord(ctx->evFork, ctx->stream));
      for (int g = 0; g < groupsC; ++g) {
        cudaStream_t st = groupsC > 1 ? ctx->aux[g] : ctx->stream;
        if (groupsC > 1) ORZ_CUDA(cudaStreamWaitEvent(st, ctx->evFork, 0));
        FrameParams pg = pc;
        pg.viewBase = (uint32_t)((uint64_t)nv * g / groupsC);
        pg.groupViews = (uint32_t)((uint64_t)nv * (g + 1) / groupsC) - pg.viewBase;
        if ((e = launch_cluster(ctx, pg, pg.groupViews, nv, st))) return e;
        if (wantQ) {
          k_query_views<<<dim3((scene->nBoxes + 255) / 256, pg.groupViews), 256, 0, st>>>(pg);
          ctx->launches++;
          ORZ_CUDA(cudaGetLastError());
        }
        if (groupsC > 1) {
          ORZ_CUDA(cudaEventRecord(ctx->evJoin[g], st));
          ORZ_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evJoin[g], 0));
        }
      }
      continue;
    }
    // Sub-batches (by descending cost) on auxiliary streams: the query kernel of a finished
    // sub-batch and the CTAs of the next one fill the SMs that the drain of the previous render
    // kernel leaves idle.  Everything is fenced by events on the context stream.
    const bool wantQuery = p.visBits || p.clipBits;
    const int groups = (p.viewOrder && nv >= 64u) ? orz_context::kGroups : 1;
    ORZ_CUDA(cudaMemsetAsync(ctx->d_counter, 0, 4 * orz_context::kGroups, ctx->stream));
    if (groups > 1) ORZ_CUDA(cudaEventRecord(ctx->evFork, ctx->stream));
    for (int g = 0; g < groups; ++g) {
      cudaStream_t st = groups > 1 ? ctx->aux[g] : ctx->stream;
      if (groups > 1) ORZ_CUDA(cudaStreamWaitEvent(st, ctx->evFork, 0));
      FrameParams pg = p;
      pg.viewBase = (uint32_t)((uint64_t)nv * g / groups);
      pg.groupViews = (uint32_t)((uint64_t)nv * (g + 1) / groups) - pg.viewBase;
      pg.viewCounter = ctx->d_counter + g;
      const uint32_t grid = std::min<uint32_t>(pg.groupViews, (uint32_t)(ctx->numSMs * perSM));
      e = launch_views(ctx, GW, trav, pg, grid, st);
      if (e) return e;
      if (wantQuery) {
        k_query_views<<<dim3((scene->nBoxes + 255) / 256, pg.groupViews), 256, 0, st>>>(pg);
        ctx->launches++;
        ORZ_CUDA(cudaGetLastError());
      }
      if (groups > 1) {
        ORZ_CUDA(cudaEventRecord(ctx->evJoin[g], st));
        ORZ_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evJoin[g], 0));
      }
    }
  }
  return ORZ_OK;
}

// Host-pointer variant: stage inputs to HBM, render, bring the requested outputs back.
extern "C" int orz_render_views(orz_context* ctx, orz_scene* scene, const orz_view_batch* hb) {
  if (!ctx || !scene || !hb || !hb->mvps || (!hb->orders && !hb->camPos)) return fail(ORZ_ERR_ARG, "orz_render_views: bad arguments");
  if (hb->nViews == 0) return ORZ_OK;
  ORZ_CUDA(cudaSetDevice(ctx->device));
  const size_t nV = hb->nViews, nOcc = scene->nOcc, blocks = (size_t)(hb->width / 8) * (hb->height / 8);
  const size_t bitWords = (scene->nBoxes + 31) / 32;
  // one staging arena: [mvps | orders/camPos | visBits | clipBits | gate | quads] + separate depth/hiz
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  const size_t oMvp = take(nV * 64);
  const size_t oOrd = take(hb->orders ? nV * nOcc * 4 : nV * 12);
  const size_t oVis = take(hb->visBits ? nV * bitWords * 4 : 0);
  const size_t oClip = take(hb->clipBits ? nV * bitWords * 4 : 0);
  const size_t oGate = take(hb->gate ? nV * nOcc : 0);
  const size_t oQuads = take(hb->quadsSubmitted ? nV * 4 : 0);
  int e;
  if ((e = ensure_scratch(ctx, 7, off))) return e;
  uint8_t* arena = (uint8_t*)ctx->d_scratch[7];
  orz_view_batch db = *hb;
  db.mvps = (const float*)(arena + oMvp);
  ORZ_CUDA(cudaMemcpyAsync(arena + oMvp, hb->mvps, nV * 64, cudaMemcpyHostToDevice, ctx->stream));
  if (hb->orders) {
    db.orders = (const uint32_t*)(arena + oOrd);
    ORZ_CUDA(cudaMemcpyAsync(arena + oOrd, hb->orders, nV * nOcc * 4, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    db.camPos = (const float*)(arena + oOrd);
    ORZ_CUDA(cudaMemcpyAsync(arena + oOrd, hb->camPos, nV * 12, cudaMemcpyHostToDevice, ctx->stream));
  }
  db.visBits = hb->visBits ? (uint32_t*)(arena + oVis) : nullptr;
  db.clipBits = hb->clipBits ? (uint32_t*)(arena + oClip) : nullptr;
  db.gate = hb->gate ? (arena + oGate) : nullptr;
  db.quadsSubmitted = hb->quadsSubmitted ? (uint32_t*)(arena + oQuads) : nullptr;
  const bool targetsOnDevice = (hb->flags & ORZ_BATCH_TARGETS_ON_DEVICE) != 0u;
  if ((hb->depth || hb->hiz) && !targetsOnDevice) {
    if (!hb->depth || !hb->hiz) return fail(ORZ_ERR_ARG, "depth and hiz outputs must be requested together");
    if ((e = ensure_scratch(ctx, 2, nV * blocks * 128))) return e;
    if ((e = ensure_scratch(ctx, 3, nV * blocks * 2))) return e;
    db.depth = (uint16_t*)ctx->d_scratch[2];
    db.hiz = (uint16_t*)ctx->d_scratch[3];
  }
  if ((e = orz_render_views_device(ctx, scene, &db))) return e;
  if (hb->visBits) ORZ_CUDA(cudaMemcpyAsync(hb->visBits, db.visBits, nV * bitWords * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (hb->clipBits) ORZ_CUDA(cudaMemcpyAsync(hb->clipBits, db.clipBits, nV * bitWords * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (hb->gate) ORZ_CUDA(cudaMemcpyAsync(hb->gate, db.gate, nV * nOcc, cudaMemcpyDeviceToHost, ctx->stream));
  if (hb->quadsSubmitted) ORZ_CUDA(cudaMemcpyAsync(hb->quadsSubmitted, db.quadsSubmitted, nV * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (hb->depth && !targetsOnDevice) {
    ORZ_CUDA(cudaMemcpyAsync(hb->depth, db.depth, nV * blocks * 128, cudaMemcpyDeviceToHost, ctx->stream));
    ORZ_CUDA(cudaMemcpyAsync(hb->hiz, db.hiz, nV * blocks * 2, cudaMemcpyDeviceToHost, ctx->stream));
  }
  ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return ORZ_OK;
}
