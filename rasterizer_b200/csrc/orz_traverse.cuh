// Block traversal (Rasterizer.cpp:1098-1292) for the large-batch kernel, the wide path and the per-call API: warp per block and lane per block.
// Part of the single translation unit orz_kernels.cu (included inside namespace orz); see DESIGN.md section 4.
#pragma once

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
