// Shared device-side types and helpers: targets, occluder meta, primitive records, frame parameters.
// Part of the single translation unit orz_kernels.cu (included inside namespace orz); see DESIGN.md section 4.
#pragma once

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

// (avg_u16x2 and the other packed-u16 helpers: orz_pixel.h)


constexpr int kFrontWords = 22;  // status, minX, maxX, minY, maxY, maxZ, CallMatrix (14 floats), quadOffset, quadCount of the occluder in this order slot

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
  uint32_t queryChunks;           // k_query_views: CTAs (of 256 boxes) per view
  int exportDepth;      // 1: the caller reads depth back -> zero-fill blocks that stayed cleared
  uint32_t clusterK;    // cluster kernel: tiles per warp
  // cluster path: speculative setup output per view (k_setup_views)
  uint32_t* recBuf;     // [nViews][totalQuads][kRecStride] records, each occluder's at its quadOffset
  uint2* hdrBuf;        // [nViews][totalQuads] bounding boxes of the records
  uint4* recInfo;       // [nViews][nOcc][2]: {records, first record slot, quadCount, -}, {block rectangle of all records, half open}
  uint2* occBox;        // [nViews][nOcc]: the same block rectangle packed {x0 | y0 << 16, x1 | y1 << 16}, {~0, 0} when there are no records
  uint32_t totalQuads;  // record slots per view (2 per quad above 65 536 blocks: index wrap)
  uint16_t* coarseHiz;  // [nViews][coarseStride]: smallest HiZ per tile of 8 x coarseCellH blocks, written by the cluster kernel for k_query_views (or NULL)
  uint32_t coarseStride, coarseCellH;
};
