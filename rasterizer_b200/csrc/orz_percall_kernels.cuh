// Kernels behind the reference's per-call API (Rasterizer.h:13-26).
// Part of the single translation unit orz_kernels.cu (included inside namespace orz); see DESIGN.md section 4.
#pragma once

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

// `out` is a word of MAPPED pinned host memory: the answer travels with the tag of the call (sequence number << 1) in one
// 32-bit store, and the host thread that is waiting for this bool sees it without a copy, an event or a stream sync
__global__ void k_query2d(Target T, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ, volatile uint32_t* out, uint32_t tag) {
  __shared__ uint32_t s_flag;
  if (threadIdx.x == 0) s_flag = 0u;
  __syncthreads();
  query2d_coop(T, minX, maxX, minY, maxY, maxZ, threadIdx.x, blockDim.x, &s_flag);
  __syncthreads();
  if (threadIdx.x == 0) {
    *out = tag | (s_flag ? 1u : 0u);
    __threadfence_system();
  }
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
