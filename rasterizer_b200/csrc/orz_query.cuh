// query2D / queryVisibility (Rasterizer.cpp:123-349): block tests, per-lane / per-warp / per-group rectangle walks, k_query_views.
// Part of the single translation unit orz_kernels.cu (included inside namespace orz); see DESIGN.md section 4.
#pragma once

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

#ifndef ORZ_QUERY_SERIAL_MAX
#define ORZ_QUERY_SERIAL_MAX 48u  // rectangles of at most this many blocks are walked by their own lane (6 / 12 / 24 / 48 / 96 / 192 measured: profiles/r2ag_*, r2ah_*)
#endif
#ifndef ORZ_QUERY_SERIAL_UNROLL
#define ORZ_QUERY_SERIAL_UNROLL 2  // HiZ reads that walk keeps in flight (1 / 2 / 4 measured: profiles/r2ao_*)
#endif
// the same with the block's HiZ already loaded
__device__ __forceinline__ bool query_block_loaded(const Target& T, uint32_t bx, uint32_t by, uint32_t h, uint32_t minX, uint32_t maxX,
                                                   uint32_t minY, uint32_t maxY, uint32_t maxZ) {
  if (maxZ <= h) return false;  // Rasterizer.cpp:310
  if (h == 1u) return true;
  const int sX = max((int)minX - (int)(8u * bx), 0), eX = min((int)maxX - (int)(8u * bx), 7);
  const int sY = max((int)minY - (int)(8u * by), 0), eY = min((int)maxY - (int)(8u * by), 7);
  if (sX == 0 && eX == 7 && sY == 0 && eY == 7) return true;  // Rasterizer.cpp:319-325
  return block_fine_test(T.depth, by * T.blocksX + bx, maxZ, sX, eX, sY, eY);
}

// one thread walks the whole rectangle (occludee queries); query2D is an OR over blocks, so two HiZ reads are kept in
// flight per step (a rectangle behind the occluders is a chain of dependent L1 / L2 round trips otherwise)
__device__ bool query2d_serial(const Target& T, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ) {
  const uint32_t bx0 = minX >> 3, bx1 = maxX >> 3, by0 = minY >> 3, by1 = maxY >> 3;
#if ORZ_QUERY_SERIAL_UNROLL > 1
  // the rectangle's blocks in row-major order, ORZ_QUERY_SERIAL_UNROLL HiZ reads in flight per step
  const uint32_t cols = bx1 - bx0 + 1u, n = cols * (by1 - by0 + 1u);
  uint32_t rx = 0u, ry = 0u;
  for (uint32_t i = 0; i < n; i += ORZ_QUERY_SERIAL_UNROLL) {
    uint32_t hh[ORZ_QUERY_SERIAL_UNROLL], xs[ORZ_QUERY_SERIAL_UNROLL], ys[ORZ_QUERY_SERIAL_UNROLL];
#pragma unroll
    for (int u = 0; u < ORZ_QUERY_SERIAL_UNROLL; ++u) {
      xs[u] = bx0 + rx; ys[u] = by0 + ry;
      hh[u] = i + (uint32_t)u < n ? (uint32_t)T.hiz[ys[u] * T.blocksX + xs[u]] : 0xffffu;  // maxZ <= 0xffff: never passes
      if (++rx == cols) { rx = 0u; ++ry; }
    }
#pragma unroll
    for (int u = 0; u < ORZ_QUERY_SERIAL_UNROLL; ++u)
      if (query_block_loaded(T, xs[u], ys[u], hh[u], minX, maxX, minY, maxY, maxZ)) return true;
  }
#else
  for (uint32_t by = by0; by <= by1; ++by)
    for (uint32_t bx = bx0; bx <= bx1; ++bx)
      if (query_block(T, bx, by, minX, maxX, minY, maxY, maxZ)) return true;
#endif
  return false;
}

// One box per lane, whole warp converged: small rectangles are walked by their own lane, large ones
// (which would leave 31 lanes idle for hundreds of iterations) are taken one at a time by the whole
// warp, 32 blocks per step with coalesced HiZ reads.  query2D is an OR over blocks, so the visiting
// order does not matter.  Returns this lane's visibility.
// `coarse` (optional): per cell of 8 x cellH blocks the SMALLEST HiZ of its blocks (written by the cluster kernel at the end
// of a view).  query2D only passes a block when maxZ > its HiZ (Rasterizer.cpp:310), so a large rectangle none of whose
// cells has maxZ > the cell's minimum cannot pass anywhere: the common case of an occluded box costs one look per cell
// instead of one per block.  Everything else takes the block walk unchanged.
__device__ __forceinline__ bool query2d_warp(const Target& T, const BoxFront& f, const int lane, const uint16_t* __restrict__ coarse = nullptr,
                                             const uint32_t cellH = 4u) {
  bool vis = false, big = false;
  if (f.status == kBoxRect) {
    const uint32_t nb = ((f.maxX >> 3) - (f.minX >> 3) + 1u) * ((f.maxY >> 3) - (f.minY >> 3) + 1u);
    if (nb <= ORZ_QUERY_SERIAL_MAX) vis = query2d_serial(T, f.minX, f.maxX, f.minY, f.maxY, f.maxZ);
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
    const uint32_t cols = (maxX >> 3) - bx0 + 1u, rows = (maxY >> 3) - by0 + 1u;
    if (coarse) {
      const uint32_t cellsX = (T.blocksX + 7u) >> 3;
      const uint32_t cx0 = bx0 >> 3, cy0 = by0 / cellH, ccols = ((maxX >> 3) >> 3) - cx0 + 1u, ncell = ccols * ((maxY >> 3) / cellH - cy0 + 1u);
      bool open = false;
      for (uint32_t c = (uint32_t)lane; c < ncell; c += 32u) {
        const uint32_t cy = c / ccols, cx = c - cy * ccols;
        open = open || maxZ > (uint32_t)coarse[(cy0 + cy) * cellsX + cx0 + cx];
      }
      if (!__any_sync(kFull, open)) continue;  // vis of lane src stays false
    }
    // lane j starts at block j of the rectangle (row major) and advances 32 blocks at a time:
    // (row, column) are stepped incrementally, two divisions per box instead of one per block
    const uint32_t q32 = 32u / cols, r32 = 32u - q32 * cols;
    uint32_t ry = (uint32_t)lane / cols, rx = (uint32_t)lane - ry * cols;
    bool found = false;
    // 128 blocks per step: the four HiZ reads of a lane are in flight together (a fully occluded
    // large box is a chain of dependent L2 round trips otherwise); fine tests only where needed
    while (__any_sync(kFull, ry < rows) && !found) {
      uint32_t h[4], bxs[4], bys[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        bxs[u] = bx0 + rx; bys[u] = by0 + ry;
        h[u] = ry < rows ? (uint32_t)T.hiz[bys[u] * T.blocksX + bxs[u]] : 0xffffu;  // maxZ <= 0xffff: skipped
        rx += r32; ry += q32;
        if (rx >= cols) { rx -= cols; ++ry; }
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


// queryVisibility for every (view, occludee box) on the finished buffers; Rasterizer.cpp:123-349
#ifndef ORZ_QUERY_THREADS
#define ORZ_QUERY_THREADS 128  // 64 / 96 / 128 / 256 / 512 measured (profiles/r2aj_*, r2ak_*)
#endif
constexpr uint32_t kQueryThreads = ORZ_QUERY_THREADS;  // boxes per CTA of k_query_views
#ifndef ORZ_QUERY_CTAS
#define ORZ_QUERY_CTAS 10  // resident CTAs per SM the query kernel is compiled for (48 registers; 9 / 10 / 12 / 16 measured: profiles/r2as_*)
#endif
__global__ void __launch_bounds__(kQueryThreads, ORZ_QUERY_CTAS) k_query_views(const FrameParams p) {
  __shared__ ViewMatrices s_vm;
  // 1-D grid, view major (grid.y would cap a batch at 65 535 views): CTA = (rank of the view in this launch, chunk of 256 boxes)
  const uint32_t vrank = blockIdx.x / p.queryChunks, chunk = blockIdx.x - vrank * p.queryChunks;
  const uint32_t view = p.viewOrder ? p.viewOrder[p.viewBase + vrank] : p.viewBase + vrank, tid = threadIdx.x;
  if (tid < 32) reinterpret_cast<float*>(&s_vm)[tid] = reinterpret_cast<const float*>(p.vmBuf + view)[tid];
  __syncthreads();
  const RcpTable rt{p.rcp, p.rcpShift};
  Target T;
  T.width = p.width; T.height = p.height; T.blocksX = p.width >> 3; T.blocksY = p.height >> 3;
  T.depth = p.depth + (size_t)view * p.depthStride;
  T.hiz = p.hiz + (size_t)view * p.hizStride;
  const uint32_t i = chunk * blockDim.x + tid;
  BoxFront f;
  f.status = kBoxCulled; f.minX = f.maxX = f.minY = f.maxY = f.maxZ = 0;
  if (i < p.nBoxes) {
    const float4 mn = p.boxes[2 * (size_t)i], mx = p.boxes[2 * (size_t)i + 1];
    const float bmn[4] = {mn.x, mn.y, mn.z, mn.w}, bmx[4] = {mx.x, mx.y, mx.z, mx.w};
    f = box_front_half(s_vm, bmn, bmx, p.width, p.height, rt);
  }
  const bool clip = f.status == kBoxNearClip;
  const bool seen = query2d_warp(T, f, (int)(tid & 31u), p.coarseHiz ? p.coarseHiz + (size_t)view * p.coarseStride : nullptr,
                                 p.coarseCellH);  // every lane must take part (warp collectives inside)
  const bool vis = clip || seen;
  const uint32_t vb = __ballot_sync(kFull, vis), cb = __ballot_sync(kFull, clip);
  const uint32_t word = i >> 5;
  if ((tid & 31u) == 0 && word < p.bitWords) {
    if (p.visBits) p.visBits[(size_t)view * p.bitWords + word] = vb;
    if (p.clipBits) p.clipBits[(size_t)view * p.bitWords + word] = cb;
  }
}
