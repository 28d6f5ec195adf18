// SurfaceAreaHeuristic::generateBatches (SurfaceAreaHeuristic.cpp:10-104) on the GPU, level by level.
// Part of the single translation unit orz_kernels.cu (included inside namespace orz).
//
// The reference recurses depth first; nodes are disjoint index ranges, so all nodes of one level are
// processed together ("segments" of one working array).  Per level and split axis:
//   keys      centre of every element's box on the axis, as an order-preserving u32 (-0 == +0)
//   sort      STABLE least-significant-digit radix sort by (segment, key), 4 bits per pass: ties keep
//             the order the previous sort left, which is what std::stable_sort on the node's range does
//             (SurfaceAreaHeuristic.cpp:22-24) and what makes batch contents reproducible
//   scans     boxes of the prefix / suffix of every position inside its segment (min / max are exact,
//             so the grouping of a parallel scan cannot change them) -> the two surface areas
//   costs     every multiple of the split granularity: areaLeft * count + areaRight * count in the
//             reference's float operations; the segment's best is an atomicMin over
//             (cost, axis, position) -- the serial loop keeps the first strictly smaller cost
// then one more sort per segment by its best axis (SurfaceAreaHeuristic.cpp:69-72).  The host only
// keeps the list of segments (start, size): it reads one u64 per segment and level.
// Everything is HBM-bound streaming over 12-byte records; no tensor cores (nothing to contract).
#pragma once

constexpr int kSahThreads = 256;
constexpr int kSahItems = 8;                        // consecutive elements per thread in a sort tile
constexpr int kSahTile = kSahThreads * kSahItems;   // 2 048 elements per CTA and radix pass
constexpr int kSahChunk = 256;                      // elements per scan chunk (chunks never straddle segments)

struct SahSeg {
  uint32_t wstart;  // first position in the working array
  uint32_t gstart;  // first position in the global order
  uint32_t n;
  uint32_t axis;    // best split axis (set by the host after the three cost passes)
};
struct SahChunk {
  uint32_t seg, pos0, len;
};
struct SahBox {
  float mn[3], mx[3];
};

// floats in `<` order as unsigned integers; -0 and +0 compare equal in the reference's comparator
__device__ __forceinline__ uint32_t sah_key(float c) {
  uint32_t u = f2u(c);
  if ((u & 0x7fffffffu) == 0) u = 0;
  return (u & kSign) ? ~u : (u | kSign);
}
__device__ __forceinline__ uint32_t sah_seg_of(const SahSeg* __restrict__ segs, uint32_t nSegs, uint32_t j) {
  uint32_t lo = 0, hi = nSegs;  // last segment with wstart <= j
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (segs[mid].wstart <= j) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_sah_gather(const uint32_t* __restrict__ order, const SahSeg* __restrict__ segs, uint32_t nSegs,
                                                    uint32_t M, uint32_t* __restrict__ idx, uint32_t* __restrict__ seg,
                                                    uint32_t* __restrict__ segStatic) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const uint32_t s = sah_seg_of(segs, nSegs, j);
  seg[j] = s;
  segStatic[j] = s;
  idx[j] = order[segs[s].gstart + (j - segs[s].wstart)];
}
__global__ void __launch_bounds__(256) k_sah_scatter_back(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ segStatic,
                                                          const SahSeg* __restrict__ segs, uint32_t M, uint32_t* __restrict__ order) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const SahSeg s = segs[segStatic[j]];
  order[s.gstart + (j - s.wstart)] = idx[j];
}

// Aabb::getCenter (VectorMath.h:50-53) on one axis; axis < 0: the segment's own best axis
__global__ void __launch_bounds__(256) k_sah_keys(const float4* __restrict__ boxes, const uint32_t* __restrict__ idx,
                                                  const uint32_t* __restrict__ segStatic, const SahSeg* __restrict__ segs, int axis,
                                                  uint32_t M, uint32_t* __restrict__ key) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int a = axis >= 0 ? axis : (int)segs[segStatic[j]].axis;
  const float4 mn = boxes[2 * (size_t)idx[j]], mx = boxes[2 * (size_t)idx[j] + 1];
  const float c = a == 0 ? mn.x + mx.x : (a == 1 ? mn.y + mx.y : mn.z + mx.z);
  key[j] = sah_key(c);
}

// ---- stable radix sort, one 4-bit digit per pass: histogram, scan, scatter -----------------------
// hist[digit * numTiles + tile]; after the exclusive scan it is the first output slot of (digit, tile)
__global__ void __launch_bounds__(kSahThreads) k_sah_hist(const uint32_t* __restrict__ src, uint32_t M, int shift, uint32_t numTiles,
                                                          uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[16];
  if (threadIdx.x < 16) h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kSahTile + threadIdx.x * kSahItems;
#pragma unroll
  for (int e = 0; e < kSahItems; ++e)
    if (base + e < M) atomicAdd(&h[(src[base + e] >> shift) & 15u], 1u);
  __syncthreads();
  if (threadIdx.x < 16) hist[threadIdx.x * numTiles + blockIdx.x] = h[threadIdx.x];
}

// One CTA per digit: in-place exclusive scan of the digit's row hist[digit * numTiles ..] (1 024 tiles per round,
// 4 consecutive ones per thread) and the digit's total; k_sah_scatter adds the totals of the smaller digits itself.
__global__ void __launch_bounds__(256) k_sah_scan(uint32_t* __restrict__ hist, uint32_t numTiles, uint32_t* __restrict__ totals) {
  __shared__ uint32_t warpSum[8];
  __shared__ uint32_t carry;
  uint32_t* row = hist + blockIdx.x * numTiles;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < numTiles; base += 1024) {
    const uint32_t i = base + 4 * tid;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e] = i + e < numTiles ? row[i + e] : 0u;
      sum += v[e];
    }
    uint32_t x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(kFull, x, d);
      if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31) warpSum[warp] = x;
    __syncthreads();
    uint32_t run = carry + x - sum;
    for (uint32_t w = 0; w < warp; ++w) run += warpSum[w];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (i + e < numTiles) row[i + e] = run;
      run += v[e];
    }
    __syncthreads();
    if (tid == 255) carry = run;
    __syncthreads();
  }
  if (tid == 0) totals[blockIdx.x] = carry;
}

// Every thread owns kSahItems CONSECUTIVE elements and counts its digits in its own column of
// cnt[digit][thread]; the block scan of cnt in (digit, thread) order then gives every element its
// place in the tile's stable order without any cross-thread ranking.  The tile is put into that order
// in shared memory first, so that the global stores of a warp go to consecutive addresses (one run per
// digit) instead of 32 different lines.  Counters are padded by one word per 16 (ORZ_SAH_CNT): both the
// per-thread column accesses and the 16-consecutive-word accesses of the scan are then conflict-free.
#define ORZ_SAH_CNT(L) cnt[(L) + ((L) >> 4)]
__global__ void __launch_bounds__(kSahThreads) k_sah_scatter(const uint32_t* __restrict__ keyIn, const uint32_t* __restrict__ idxIn,
                                                             const uint32_t* __restrict__ segIn, uint32_t* __restrict__ keyOut,
                                                             uint32_t* __restrict__ idxOut, uint32_t* __restrict__ segOut, uint32_t M,
                                                             int shift, int bySegment, uint32_t numTiles,
                                                             const uint32_t* __restrict__ histScanned,
                                                             const uint32_t* __restrict__ digitTotals) {
  __shared__ uint32_t cnt[16 * kSahThreads + kSahThreads];
  __shared__ uint32_t warpTot[kSahThreads / 32];
  __shared__ uint32_t sKey[kSahTile], sIdx[kSahTile], sSeg[kSahTile];
  __shared__ uint32_t digitStart[16], globalStart[16];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
#pragma unroll
  for (int d = 0; d < 16; ++d) ORZ_SAH_CNT(d * kSahThreads + tid) = 0;
  const uint32_t tileStart = blockIdx.x * kSahTile, base = tileStart + tid * kSahItems;
  const uint32_t tileCount = min((uint32_t)kSahTile, M - tileStart);
  uint32_t k[kSahItems], ix[kSahItems], sg[kSahItems], dg[kSahItems], rk[kSahItems];
#pragma unroll
  for (int e = 0; e < kSahItems; ++e) {
    dg[e] = 16u;
    if (base + e < M) {
      k[e] = keyIn[base + e];
      ix[e] = idxIn[base + e];
      sg[e] = segIn[base + e];
      dg[e] = ((bySegment ? sg[e] : k[e]) >> shift) & 15u;
      rk[e] = ORZ_SAH_CNT(dg[e] * kSahThreads + tid)++;
    }
  }
  __syncthreads();
  // exclusive scan of the 4 096 counters: 16 per thread, then across the block
  uint32_t vals[16], sum = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    vals[i] = ORZ_SAH_CNT(tid * 16 + i);
    sum += vals[i];
  }
  uint32_t x = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(kFull, x, d);
    if (lane >= (uint32_t)d) x += y;
  }
  if (lane == 31) warpTot[warp] = x;
  __syncthreads();
  uint32_t run = x - sum;
  for (uint32_t w = 0; w < warp; ++w) run += warpTot[w];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    ORZ_SAH_CNT(tid * 16 + i) = run;
    run += vals[i];
  }
  __syncthreads();
  if (tid < 16) {
    digitStart[tid] = ORZ_SAH_CNT(tid * kSahThreads);
    uint32_t start = histScanned[tid * numTiles + blockIdx.x];  // elements of this digit in earlier tiles ...
    for (uint32_t d = 0; d < tid; ++d) start += digitTotals[d];  // ... after all elements of smaller digits
    globalStart[tid] = start;
  }
#pragma unroll
  for (int e = 0; e < kSahItems; ++e)
    if (dg[e] < 16u) {
      const uint32_t place = ORZ_SAH_CNT(dg[e] * kSahThreads + tid) + rk[e];  // position in the tile's stable order
      sKey[place] = k[e];
      sIdx[place] = ix[e];
      sSeg[place] = sg[e];
    }
  __syncthreads();
  for (uint32_t j = tid; j < tileCount; j += kSahThreads) {
    const uint32_t key = sKey[j], seg = sSeg[j];
    const uint32_t d = ((bySegment ? seg : key) >> shift) & 15u;
    const uint32_t dest = globalStart[d] + (j - digitStart[d]);
    keyOut[dest] = key;
    idxOut[dest] = sIdx[j];
    segOut[dest] = seg;
  }
}
#undef ORZ_SAH_CNT

// ---- prefix / suffix boxes and their areas --------------------------------------------------------
__device__ __forceinline__ SahBox sah_empty() {
  SahBox b;
#pragma unroll
  for (int k = 0; k < 3; ++k) { b.mn[k] = u2f(0x7f800000u); b.mx[k] = u2f(0xff800000u); }
  return b;
}
// Aabb::include (VectorMath.h:38-42)
__device__ __forceinline__ SahBox sah_merge(const SahBox& a, const SahBox& b) {
  SahBox r;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.mn[k] = min_x86(a.mn[k], b.mn[k]); r.mx[k] = max_x86(a.mx[k], b.mx[k]); }
  return r;
}
// Aabb::surfaceArea (VectorMath.h:60-65): dpps 0x7F of the extents with their yzx rotation
__device__ __forceinline__ float sah_area(const SahBox& b) {
  const float ex = b.mx[0] - b.mn[0], ey = b.mx[1] - b.mn[1], ez = b.mx[2] - b.mn[2];
  return (ex * ey + ey * ez) + ez * ex;
}
__device__ __forceinline__ SahBox sah_shfl_up(const SahBox& b, int d) {
  SahBox r;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.mn[k] = __shfl_up_sync(kFull, b.mn[k], d); r.mx[k] = __shfl_up_sync(kFull, b.mx[k], d); }
  return r;
}
__device__ __forceinline__ SahBox sah_shfl_down(const SahBox& b, int d) {
  SahBox r;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.mn[k] = __shfl_down_sync(kFull, b.mn[k], d); r.mx[k] = __shfl_down_sync(kFull, b.mx[k], d); }
  return r;
}
__device__ __forceinline__ SahBox sah_shfl(const SahBox& b, int srcLane) {
  SahBox r;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.mn[k] = __shfl_sync(kFull, b.mn[k], srcLane); r.mx[k] = __shfl_sync(kFull, b.mx[k], srcLane); }
  return r;
}
// inclusive scans inside a warp: lanes 0..lane (up) / lanes lane..31 (down)
__device__ __forceinline__ SahBox sah_warp_scan_up(SahBox b, uint32_t lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const SahBox o = sah_shfl_up(b, d);
    if (lane >= (uint32_t)d) b = sah_merge(o, b);
  }
  return b;
}
__device__ __forceinline__ SahBox sah_warp_scan_down(SahBox b, uint32_t lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const SahBox o = sah_shfl_down(b, d);
    if (lane + (uint32_t)d < 32u) b = sah_merge(b, o);
  }
  return b;
}
__device__ __forceinline__ SahBox sah_load_box(const float4* __restrict__ boxes, uint32_t element) {
  const float4 mn = boxes[2 * (size_t)element], mx = boxes[2 * (size_t)element + 1];
  SahBox b;
  b.mn[0] = mn.x; b.mn[1] = mn.y; b.mn[2] = mn.z;
  b.mx[0] = mx.x; b.mx[1] = mx.y; b.mx[2] = mx.z;
  return b;
}

// box of every chunk
__global__ void __launch_bounds__(kSahChunk) k_sah_chunk_boxes(const float4* __restrict__ boxes, const uint32_t* __restrict__ idx,
                                                               const SahChunk* __restrict__ chunks, SahBox* __restrict__ chunkBox) {
  __shared__ SahBox warpBox[kSahChunk / 32];
  const SahChunk c = chunks[blockIdx.x];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  SahBox b = tid < c.len ? sah_load_box(boxes, idx[c.pos0 + tid]) : sah_empty();
  b = sah_warp_scan_up(b, lane);
  if (lane == 31) warpBox[warp] = b;
  __syncthreads();
  if (tid == 0) {
    SahBox all = warpBox[0];
    for (int w = 1; w < kSahChunk / 32; ++w) all = sah_merge(all, warpBox[w]);
    chunkBox[blockIdx.x] = all;
  }
}

// per segment: box of all chunks before (warp 0) / after (warp 1) each chunk
__global__ void __launch_bounds__(64) k_sah_chunk_scan(const SahBox* __restrict__ chunkBox, const uint32_t* __restrict__ segChunk0,
                                                       SahBox* __restrict__ before, SahBox* __restrict__ after) {
  const uint32_t c0 = segChunk0[blockIdx.x], c1 = segChunk0[blockIdx.x + 1], lane = threadIdx.x & 31u;
  const bool fromRight = threadIdx.x >= 32;
  SahBox* __restrict__ out = fromRight ? after : before;
  SahBox carry = sah_empty();
  for (uint32_t done = 0; c0 + done < c1; done += 32) {
    const uint32_t step = done + lane;  // distance from the segment's first (last) chunk
    const bool valid = c0 + step < c1;
    const uint32_t c = !valid ? c0 : (fromRight ? c1 - 1 - step : c0 + step);
    const SahBox incl = sah_warp_scan_up(valid ? chunkBox[c] : sah_empty(), lane);
    SahBox excl = sah_shfl_up(incl, 1);
    if (lane == 0) excl = sah_empty();
    if (valid) out[c] = sah_merge(carry, excl);
    carry = sah_merge(carry, sah_shfl(incl, 31));
  }
}

// areaLeft[j] = area of the box of positions segment start .. j, areaRight[j] = of j .. segment end
// (areasFromLeft / areasFromRight, SurfaceAreaHeuristic.cpp:26-44)
__global__ void __launch_bounds__(kSahChunk) k_sah_areas(const float4* __restrict__ boxes, const uint32_t* __restrict__ idx,
                                                         const SahChunk* __restrict__ chunks, const SahBox* __restrict__ before,
                                                         const SahBox* __restrict__ after, float* __restrict__ areaLeft,
                                                         float* __restrict__ areaRight) {
  __shared__ SahBox warpUp[kSahChunk / 32], warpDown[kSahChunk / 32];
  const SahChunk c = chunks[blockIdx.x];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const SahBox own = tid < c.len ? sah_load_box(boxes, idx[c.pos0 + tid]) : sah_empty();
  const SahBox up = sah_warp_scan_up(own, lane), down = sah_warp_scan_down(own, lane);
  if (lane == 31) warpUp[warp] = up;
  if (lane == 0) warpDown[warp] = down;
  __syncthreads();
  SahBox left = before[blockIdx.x], right = after[blockIdx.x];
  for (uint32_t w = 0; w < warp; ++w) left = sah_merge(left, warpUp[w]);
  for (uint32_t w = warp + 1; w < kSahChunk / 32; ++w) right = sah_merge(right, warpDown[w]);
  left = sah_merge(left, up);
  right = sah_merge(right, down);
  if (tid < c.len) {
    areaLeft[c.pos0 + tid] = sah_area(left);
    areaRight[c.pos0 + tid] = sah_area(right);
  }
}

// cost of splitting before position j (SurfaceAreaHeuristic.cpp:46-66); best[segment] = min over
// (cost, axis, position): equal costs keep the earlier axis / position like the serial loop.  A block that
// lies inside one segment (nearly all of them on the upper levels) reduces in shared memory first, so the
// root's quarter of a million candidates do not queue up on one address.
__global__ void __launch_bounds__(256) k_sah_costs(const float* __restrict__ areaLeft, const float* __restrict__ areaRight,
                                                   const uint32_t* __restrict__ segStatic, const SahSeg* __restrict__ segs, uint32_t M,
                                                   uint32_t granularity, uint32_t axis, unsigned long long* __restrict__ best) {
  __shared__ unsigned long long blockBest;
  const uint32_t blockFirst = blockIdx.x * blockDim.x, blockLast = min(blockFirst + blockDim.x, M) - 1u;
  const bool oneSegment = segStatic[blockFirst] == segStatic[blockLast];  // segments are contiguous
  if (threadIdx.x == 0) blockBest = ~0ull;
  __syncthreads();
  const uint32_t j = blockFirst + threadIdx.x;
  if (j < M) {
    const uint32_t sIdx = segStatic[j];
    const SahSeg s = segs[sIdx];
    const uint32_t pos = j - s.wstart;
    if (pos >= granularity && pos < s.n - granularity && pos % granularity == 0) {
      const float scaledLeft = areaLeft[j - 1] * (float)(int)pos;
      const float scaledRight = areaRight[j] * (float)(int)(s.n - pos);
      const float cost = scaledLeft + scaledRight;
      if (cost < u2f(0x7f800000u)) {  // comilt against the initial +inf: never for inf / NaN
        const unsigned long long packed = ((unsigned long long)sah_key(cost) << 32) | ((unsigned long long)axis << 30) | pos;
        atomicMin(oneSegment ? &blockBest : best + sIdx, packed);
      }
    }
  }
  __syncthreads();
  if (oneSegment && threadIdx.x == 0 && blockBest != ~0ull) atomicMin(best + segStatic[blockFirst], blockBest);
}
