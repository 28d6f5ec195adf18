// Wide path (config 4): one view split over the whole GPU, ungated.
// Part of the single translation unit orz_kernels.cu (included inside namespace orz); see DESIGN.md section 4.
#pragma once

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
