// CUDA kernels (sm_100a) + C ABI of the B200-native occlusion-culling rasterizer: ONE translation
// unit (every kernel shares the scalar cores of orz_core.h and is compiled with -fmad=false);
// the device code lives in the .cuh files included below, the C ABI (include/orz.h) at the end.
//
//   orz_device_common.cuh    targets, occluder meta, primitive records, frame parameters
//   orz_traverse.cuh         block traversal, warp per block / lane per block (Rasterizer.cpp:1098-1292)
//   orz_query.cuh            query2D / queryVisibility (Rasterizer.cpp:123-349), k_query_views
//   orz_batch_kernels.cuh    large batches: k_prepare_views, k_sort_views, k_render_views (one CTA per view)
//   orz_cluster_kernels.cuh  up to 1 024 views: k_setup_views (speculative setup) + k_raster_views_cluster
//                            (one thread-block cluster per view, dataflow gates, tile-major register depth)
//   orz_wide_kernels.cuh     config 4: one ungated view split over the whole GPU
//   orz_percall_kernels.cuh  the reference's per-call API (Rasterizer.h:13-26)
//   orz_bake_kernel.cuh      Occluder::bake (Occluder.cpp:7-181)
//   orz_sah_kernels.cuh      SurfaceAreaHeuristic::generateBatches (SurfaceAreaHeuristic.cpp:10-104), level-synchronous
//
// Execution models, data layout and measurements: DESIGN.md sections 3 and 4.
// No tensor cores: nothing here is a contraction.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <string>
#include <vector>

#include "../../include/orz.h"
#include "orz_core.h"
#include "orz_pixel.h"
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
#ifndef ORZ_SETUP_SMALL_CTAS
#define ORZ_SETUP_SMALL_CTAS 2048  // launches of at least this many (occluder, view) pairs set up with one-warp CTAs
#endif
#ifndef ORZ_GROUPS
#define ORZ_GROUPS 4  // sub-batches of a large batch, each on its own stream (2 / 6 / 8 measured: profiles/r2am_*)
#endif
#ifndef ORZ_THREADS_PER_SM_V2
#define ORZ_THREADS_PER_SM_V2 512  // same, for the lane-per-block traversal (register cap 128)
#endif
#ifndef ORZ_THREADS_PER_SM
#define ORZ_THREADS_PER_SM 1024  // resident threads per SM the view-batch kernel is compiled for (register cap = 65536 / this)
#endif

namespace orz {
#include "orz_device_common.cuh"
#include "orz_traverse.cuh"
#include "orz_query.cuh"
#include "orz_batch_kernels.cuh"
#include "orz_cluster_kernels.cuh"
#include "orz_wide_kernels.cuh"
#include "orz_percall_kernels.cuh"
#include "orz_bake_kernel.cuh"
#include "orz_sah_kernels.cuh"

}  // namespace orz

// =================================================================================================
// C ABI
using namespace orz;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int orz::set_error(int code, const std::string& msg) { return fail(code, msg); }
#define ORZ_CUDA(x)                                                                                   \
  do {                                                                                                \
    cudaError_t _e = (x);                                                                             \
    if (_e != cudaSuccess) return fail(ORZ_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// same, for code that owns a half-built object: run `cleanup` before returning the error
#define ORZ_CUDA_OR(cleanup, x)                                                \
  do {                                                                         \
    cudaError_t _e = (x);                                                      \
    if (_e != cudaSuccess) {                                                   \
      cleanup;                                                                 \
      return fail(ORZ_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); \
    }                                                                          \
  } while (0)

struct orz_context {
  int device = 0;
  int numSMs = 0;
  size_t maxSmemOptin = 0;  // largest dynamic shared memory a CTA may ask for on this device
  cudaStream_t stream = nullptr;
  uint2* d_lut = nullptr;
  uint32_t* d_rcp = nullptr;
  int rcpBits = 0;
  bool rcpExact = true;
  std::vector<uint32_t> h_rcp;
  uint64_t launches = 0;
  int groupWarps = 0;
  int traversal = 2;  // 1 = warp per block, 2 = lane per block (default)
  int clusterViews = 16384;  // batches of at most this many views run one thread-block cluster per view (0 = never); measured faster than the large-batch kernel at every size (profiles/r2_probe_matrix_paths.txt)
  int clusterSize = 0;     // CTAs per cluster (2, 4, 8, 16); 0 = automatic
  // grow-only device scratch
  uint32_t* d_counter = nullptr;
  void* d_scratch[14] = {nullptr};
  size_t scratchBytes[14] = {0};
  // pinned + mapped mailbox for the per-call queries: the answering kernel stores tag | answer straight into a word the
  // waiting host thread polls (no copy, no event, no stream sync); slots are handed out round robin, tags never repeat
  static constexpr uint32_t kMailSlots = 256;
  uint32_t* h_pinned = nullptr;
  uint32_t* d_mail = nullptr;    // the same words as the device sees them
  uint32_t mailSeq = 0;          // sequence number of the last query handed to the GPU (tag = sequence << 2)
  uint32_t mailNext = 0;         // next slot
  size_t smemCall = 0;           // k_rasterize_call
  bool percallLegacy = false;    // ORZ_PERCALL_LEGACY=1: one launch per rasterize (round-1 kernel) and per query, no predicted chains
  static constexpr int kGroups = ORZ_GROUPS;  // sub-batches pipelined on auxiliary streams
  cudaStream_t aux[kGroups] = {nullptr};
  cudaEvent_t evFork = nullptr, evJoin[kGroups] = {nullptr};
  size_t arenaBudget = size_t(8) << 30;  // bytes of internal per-view depth+HiZ targets (views are chunked to fit)
  size_t recordBudget = size_t(8) << 30; // bytes of speculative setup records of the cluster path (views are chunked to fit)
  // dynamic shared memory already granted to a kernel instantiation on this context's device (cudaFuncSetAttribute is
  // per device and idempotent: keeping the record per context avoids process-wide mutable state)
  size_t smemViews[2][5] = {{0}};   // [traversal - 1][log2 GW]
  size_t smemCluster[10] = {0};     // [log2 C (+ 5 for 8 x 1 tiles)]
  uint32_t groupCut[kGroups + 1];   // cumulative per-mille shares of the cluster path's sub-batches (equal; ORZ_GROUP_CUTS="a,b,c" with four groups)
  orz_context() { for (int g = 0; g <= kGroups; ++g) groupCut[g] = 1000u * (uint32_t)g / (uint32_t)kGroups; }
  int clusterTileH = 0;             // tile height of the cluster path: 4, 1, or 0 = automatic = 4 (ORZ_CLUSTER_TILE_H)
  bool coarseQuery = false;         // cluster path: occludee queries look at per-tile HiZ minima first (ORZ_COARSE_QUERY=1; exact, measured neutral: off)
  uint32_t percallTileH = 1;        // tile height of the per-call rasterize (ORZ_PERCALL_TILE_H)
  size_t smemTiles = 0;             // k_raster_tiles
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
  // ---- predicted query chains (orz_percall_kernels.cuh): the frame loop of Main.cpp:192-206 asks the same boxes in
  // (nearly) the same order frame after frame, so every launch also answers the rectangle queries expected next.
  typedef std::array<uint32_t, 8> BoxKey;  // bit patterns of boundsMin / boundsMax
  std::vector<BoxKey> history, current;   // boxes asked in the previous frame / so far in this one (a frame = clear() to clear())
  size_t cursor = 0;                      // history[cursor] is the query expected next
  struct Pending { uint32_t rect[5], slot, tag; };
  std::vector<Pending> pending;           // answers under way for the buffers AS THEY ARE NOW (dropped by rasterize / clear)
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
#if ORZ_WAIT_STATS
extern "C" int orz_debug_wait_stats(unsigned long long* out8) {  // measurement builds only: reads and clears the counters
  unsigned long long zero[8] = {0};
  ORZ_CUDA(cudaDeviceSynchronize());
  ORZ_CUDA(cudaMemcpyFromSymbol(out8, orz::g_waitStats, sizeof zero));
  ORZ_CUDA(cudaMemcpyToSymbol(orz::g_waitStats, zero, sizeof zero));
  return ORZ_OK;
}
#endif
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
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaGetDeviceProperties(&prop, device));
  ctx->numSMs = prop.multiProcessorCount;
  ctx->maxSmemOptin = prop.sharedMemPerBlockOptin;
  ctx->arenaBudget = std::min<size_t>(size_t(24) << 30, prop.totalGlobalMem / 6);
  if (const char* rb = getenv("ORZ_RECORD_BUDGET_GB")) ctx->recordBudget = (size_t)std::max(1, atoi(rb)) << 30;
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
  for (int g = 0; g < orz_context::kGroups; ++g) {
    ORZ_CUDA_OR(orz_context_destroy(ctx), cudaStreamCreateWithFlags(&ctx->aux[g], cudaStreamNonBlocking));
    ORZ_CUDA_OR(orz_context_destroy(ctx), cudaEventCreateWithFlags(&ctx->evJoin[g], cudaEventDisableTiming));
  }
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaMalloc(&ctx->d_lut, 4096 * sizeof(uint2)));
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaMemcpy(ctx->d_lut, edge_mask_table(), 4096 * sizeof(uint2), cudaMemcpyHostToDevice));
  probe_host_rcp(ctx->h_rcp, ctx->rcpBits, ctx->rcpExact);
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaMalloc(&ctx->d_rcp, ctx->h_rcp.size() * 4));
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaMemcpy(ctx->d_rcp, ctx->h_rcp.data(), ctx->h_rcp.size() * 4, cudaMemcpyHostToDevice));
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaMalloc(&ctx->d_counter, 64));
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaHostAlloc(&ctx->h_pinned, orz_context::kMailSlots * 4, cudaHostAllocMapped));
  memset(ctx->h_pinned, 0, orz_context::kMailSlots * 4);
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaMemset(ctx->d_counter, 0, 64));
  if (const char* legacy = getenv("ORZ_PERCALL_LEGACY")) ctx->percallLegacy = legacy[0] == '1';
  if (const char* cq = getenv("ORZ_COARSE_QUERY")) ctx->coarseQuery = cq[0] != '0';
  if (const char* cuts = getenv("ORZ_GROUP_CUTS")) {
    unsigned a = 250, b = 500, c = 750;
    if (orz_context::kGroups == 4 && sscanf(cuts, "%u,%u,%u", &a, &b, &c) == 3 && a <= b && b <= c && c <= 1000) { ctx->groupCut[1] = a; ctx->groupCut[2] = b; ctx->groupCut[3] = c; }
  }
  if (const char* th = getenv("ORZ_PERCALL_TILE_H")) ctx->percallTileH = atoi(th) == 4 ? 4u : 1u;
  if (const char* th = getenv("ORZ_CLUSTER_TILE_H")) ctx->clusterTileH = atoi(th) == 1 ? 1 : atoi(th) == 4 ? 4 : 0;
  ORZ_CUDA_OR(orz_context_destroy(ctx), cudaHostGetDevicePointer((void**)&ctx->d_mail, ctx->h_pinned, 0));
  *out = ctx;
  return ORZ_OK;
}
extern "C" void orz_context_destroy(orz_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& p : ctx->d_scratch) if (p) cudaFree(p);
  cudaFree(ctx->d_lut); cudaFree(ctx->d_rcp); cudaFree(ctx->d_counter);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  for (int g = 0; g < orz_context::kGroups; ++g) {  // a context that failed half-way through creation has null handles
    if (ctx->aux[g]) { cudaStreamSynchronize(ctx->aux[g]); cudaStreamDestroy(ctx->aux[g]); }
    if (ctx->evJoin[g]) cudaEventDestroy(ctx->evJoin[g]);
  }
  if (ctx->evFork) cudaEventDestroy(ctx->evFork);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}
extern "C" int orz_context_synchronize(orz_context* ctx) {
  if (!ctx) return fail(ORZ_ERR_ARG, "orz_context_synchronize: context is NULL");
  ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return ORZ_OK;
}
extern "C" void* orz_context_stream(orz_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int orz_context_device(orz_context* ctx) { return ctx ? ctx->device : -1; }
extern "C" uint64_t orz_context_launch_count(orz_context* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int orz_context_set_group_warps(orz_context* ctx, int warps) {
  if (!ctx || (warps != 0 && warps != 1 && warps != 2 && warps != 4 && warps != 8 && warps != 16)) return fail(ORZ_ERR_ARG, "group warps must be 0, 1, 2, 4, 8 or 16");
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
extern "C" int orz_context_set_tile_height(orz_context* ctx, int clusterTileH, int perCallTileH) {
  if (!ctx || (clusterTileH != 0 && clusterTileH != 1 && clusterTileH != 4) || (perCallTileH != 1 && perCallTileH != 4))
    return fail(ORZ_ERR_ARG, "tile heights: 0 (automatic), 1 or 4 for the cluster path, 1 or 4 for the per-call path");
  ctx->clusterTileH = clusterTileH;
  ctx->percallTileH = (uint32_t)perCallTileH;
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
  if (!ctx || !packets || !refMin4 || !refMax4 || !out || packetCount % 4 != 0) return fail(ORZ_ERR_ARG, "orz_occluder_create: bad arguments");
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
    ORZ_CUDA_OR(orz_occluder_destroy(o), cudaMalloc(&o->d_quads, tmp.size() * sizeof(uint4)));
    ORZ_CUDA_OR(orz_occluder_destroy(o), cudaMemcpy(o->d_quads, tmp.data(), tmp.size() * sizeof(uint4), cudaMemcpyHostToDevice));
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
  r->T.depth = nullptr; r->T.hiz = nullptr;
  ORZ_CUDA_OR(orz_rasterizer_destroy(r), cudaMalloc(&r->T.depth, blocks * 128));
  ORZ_CUDA_OR(orz_rasterizer_destroy(r), cudaMalloc(&r->T.hiz, (blocks + 8) * 2));
  ORZ_CUDA_OR(orz_rasterizer_destroy(r), cudaMemsetAsync(r->T.depth, 0, blocks * 128, ctx->stream));
  ORZ_CUDA_OR(orz_rasterizer_destroy(r), cudaMemsetAsync(r->T.hiz, 0, (blocks + 8) * 2, ctx->stream));  // Rasterizer.cpp:71: HiZ starts at 0 until clear()
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
  r->pending.clear();  // a new frame: what was asked in the last one is the prediction for this one
  r->history.swap(r->current);
  r->current.clear();
  r->cursor = 0;
  return ORZ_OK;
}

static uint32_t tiles_of(uint32_t width, uint32_t height, uint32_t tileH);
// ---- predicted query chains -------------------------------------------------------------------------------------
static void chain_add(orz_rasterizer* r, QueryChain& qc, const uint32_t rect[5]) {
  orz_context* ctx = r->ctx;
  const uint32_t q = qc.n++;
  for (int i = 0; i < 5; ++i) qc.rect[q][i] = rect[i];
  qc.tag[q] = (++ctx->mailSeq) << 2;
  qc.slot[q] = ctx->mailNext;
  ctx->mailNext = (ctx->mailNext + 1u) % orz_context::kMailSlots;
  orz_rasterizer::Pending pe;
  for (int i = 0; i < 5; ++i) pe.rect[i] = rect[i];
  pe.slot = qc.slot[q]; pe.tag = qc.tag[q];
  r->pending.push_back(pe);
}
// the rectangle queries expected after history[cursor - 1], under the current matrix: boxes the frustum culls are answered
// on the host and skipped; a near-clipped box is visible without a test and will be rasterised, which ends the chain
static void chain_predict(orz_rasterizer* r, QueryChain& qc) {
  if (r->ctx->percallLegacy) return;
  const RcpTable rt{r->ctx->h_rcp.data(), 23 - r->ctx->rcpBits};
  for (size_t j = r->cursor; j < r->history.size() && qc.n < kChainMax; ++j) {
    float mn[4], mx[4];
    memcpy(mn, r->history[j].data(), 16);
    memcpy(mx, r->history[j].data() + 4, 16);
    const BoxFront f = box_front_half(r->vm, mn, mx, r->T.width, r->T.height, rt);
    if (f.status == kBoxCulled) continue;
    if (f.status == kBoxNearClip) break;
    const uint32_t rect[5] = {f.minX, f.maxX, f.minY, f.maxY, f.maxZ};
    chain_add(r, qc, rect);
  }
}
// waits for the tagged answer of one mailbox slot: 0 / 1 = the answer, 2 = the chain stopped before this query
static int mail_wait(orz_context* ctx, uint32_t slot, uint32_t tag, uint32_t* answer) {
  volatile uint32_t* mail = ctx->h_pinned + slot;
  for (uint32_t spins = 0;; ++spins) {
    const uint32_t v = *mail;
    if ((v & ~3u) == tag) { *answer = v & 3u; return ORZ_OK; }
    if ((spins & 0x3fffu) == 0x3fffu) {  // now and then: has the stream failed (or finished without our store)?
      const cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q != cudaSuccess && q != cudaErrorNotReady) return fail(ORZ_ERR_CUDA, std::string("per-call query: ") + cudaGetErrorString(q));
      if (q == cudaSuccess && ((*mail) & ~3u) != tag) return fail(ORZ_ERR_CUDA, "per-call query: the kernel finished without an answer");
    }
  }
}
extern "C" int orz_rasterizer_rasterize(orz_rasterizer* r, const orz_occluder* occ, int clipped) {
  if (!r || !occ) return fail(ORZ_ERR_ARG, "orz_rasterizer_rasterize: bad arguments");
  if (occ->nQuads == 0) return ORZ_OK;
  ORZ_CUDA(cudaSetDevice(r->ctx->device));
  r->pending.clear();  // the buffers change: answers under way describe the old ones
  {
    // one launch, tile major over the whole GPU (8 x 1 strips: the stacked updates of a block row are the critical path), the
    // queries expected next answered by its last CTA
    orz_context* ctx = r->ctx;
    const uint32_t blocks = r->T.blocksX * r->T.blocksY;
    const uint32_t tileH = ctx->percallTileH;
    const uint32_t nTiles = tiles_of(r->T.width, r->T.height, tileH);
    const uint32_t grid = std::min<uint32_t>((uint32_t)ctx->numSMs, (nTiles + kClusterGW - 1u) / kClusterGW);
    const uint32_t K = (nTiles + grid * kClusterGW - 1u) / (grid * kClusterGW);
    const size_t smem = CallSmem::bytes(K);
    if (!ctx->percallLegacy && occ->nQuads <= kCallQuadsMax && blocks <= 65536u && K <= 32u && smem <= ctx->maxSmemOptin) {
      if (ctx->smemCall < smem) {
        ORZ_CUDA(cudaFuncSetAttribute(k_rasterize_call<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ORZ_CUDA(cudaFuncSetAttribute(k_rasterize_call<kTileH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->smemCall = smem;
      }
      CallMatrix cm;
      prepare_call(r->vm.baked, occ->refMin, occ->refMax, cm);
      QueryChain qc;
      qc.n = 0;
      chain_predict(r, qc);
      if (tileH == 1u)
        k_rasterize_call<1><<<grid, kClusterGW * 32, smem, ctx->stream>>>(cm, occ->d_quads, occ->nQuads, clipped, r->T, ctx->d_rcp, 23 - ctx->rcpBits, ctx->d_lut, K, qc,
                                                                           ctx->d_mail, ctx->d_counter + 8);
      else
        k_rasterize_call<kTileH><<<grid, kClusterGW * 32, smem, ctx->stream>>>(cm, occ->d_quads, occ->nQuads, clipped, r->T, ctx->d_rcp, 23 - ctx->rcpBits, ctx->d_lut, K,
                                                                                qc, ctx->d_mail, ctx->d_counter + 8);
      ctx->launches++;
      ORZ_CUDA(cudaGetLastError());
      return ORZ_OK;
    }
  }
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
  orz_context* ctx = r->ctx;
  const uint32_t rect[5] = {minX, maxX, minY, maxY, maxZ};
  // already asked for by an earlier launch (a predicted chain)?
  for (size_t i = 0; i < r->pending.size(); ++i) {
    if (memcmp(r->pending[i].rect, rect, sizeof rect) != 0) continue;
    const orz_rasterizer::Pending pe = r->pending[i];
    r->pending.erase(r->pending.begin() + (long)i);
    if (ctx->mailSeq - (pe.tag >> 2) >= orz_context::kMailSlots) break;  // its mailbox word has been handed out again (other rasterizers of this context): ask now
    uint32_t answer = 0;
    if (int e = mail_wait(ctx, pe.slot, pe.tag, &answer)) return e;
    if (answer & 2u) break;  // the chain stopped before it got here: ask now
    *visible = (int)(answer & 1u);
    return ORZ_OK;
  }
  // the answer comes back through mapped pinned memory, tagged with this call's sequence number: no copy, no stream sync;
  // the launch also carries the queries expected after this one
  QueryChain qc;
  qc.n = 0;
  r->pending.clear();  // (entries of an abandoned chain; answers still under way land in slots nobody reads)
  chain_add(r, qc, rect);
  const orz_rasterizer::Pending mine = r->pending.back();
  r->pending.pop_back();
  chain_predict(r, qc);
  k_query_chain<<<1, 256, 0, ctx->stream>>>(r->T, qc, ctx->d_mail);
  ctx->launches++;
  ORZ_CUDA(cudaGetLastError());
  uint32_t answer = 0;
  if (int e = mail_wait(ctx, mine.slot, mine.tag, &answer)) return e;
  *visible = (int)(answer & 1u);
  return ORZ_OK;
}
extern "C" int orz_rasterizer_query_visibility(orz_rasterizer* r, const float* bmin, const float* bmax, int* visible, int* needsClipping) {
  if (!r || !bmin || !bmax || !visible) return fail(ORZ_ERR_ARG, "orz_rasterizer_query_visibility: bad arguments");
  // the front half is state independent and scalar: evaluate it here with the same core the
  // kernels use, then ask the GPU only for the rectangle test (Rasterizer.cpp:275)
  const RcpTable rt{r->ctx->h_rcp.data(), 23 - r->ctx->rcpBits};
  const BoxFront f = box_front_half(r->vm, bmin, bmax, r->T.width, r->T.height, rt);
  {  // where this frame is in the last frame's sequence of queries (the prediction for what is asked next)
    orz_rasterizer::BoxKey key;
    memcpy(key.data(), bmin, 16);
    memcpy(key.data() + 4, bmax, 16);
    if (r->cursor < r->history.size() && r->history[r->cursor] == key) ++r->cursor;
    else {
      for (size_t j = 0; j < r->history.size(); ++j)
        if (r->history[j] == key) { r->cursor = j + 1; break; }
    }
    if (r->current.size() < 65536u) r->current.push_back(key);
  }
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
  ORZ_CUDA_OR(orz_scene_destroy(s), cudaMalloc(&s->d_quads, std::max<size_t>(quads.size(), 1) * sizeof(uint4)));
  ORZ_CUDA_OR(orz_scene_destroy(s), cudaMemcpy(s->d_quads, quads.data(), quads.size() * sizeof(uint4), cudaMemcpyHostToDevice));
  ORZ_CUDA_OR(orz_scene_destroy(s), cudaMalloc(&s->d_occ, meta.size() * sizeof(OccMeta)));
  ORZ_CUDA_OR(orz_scene_destroy(s), cudaMemcpy(s->d_occ, meta.data(), meta.size() * sizeof(OccMeta), cudaMemcpyHostToDevice));
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
// ---- SurfaceAreaHeuristic::generateBatches on the GPU: kernels in orz_sah_kernels.cuh, level loop here
#define ORZ_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#include "orz_sah_driver.inl"

extern "C" int orz_scene_set_occludees(orz_scene* s, const float* boxes, uint32_t n) {
  if (!s || (n && !boxes)) return fail(ORZ_ERR_ARG, "orz_scene_set_occludees: bad arguments");
  ORZ_CUDA(cudaSetDevice(s->ctx->device));
  // the new set is complete in HBM before the old one is released: a failure leaves the scene as it was
  float4* fresh = nullptr;
  if (n) {
    ORZ_CUDA(cudaMalloc(&fresh, (size_t)n * 32));
    ORZ_CUDA_OR(cudaFree(fresh), cudaMemcpy(fresh, boxes, (size_t)n * 32, cudaMemcpyHostToDevice));
  }
  ORZ_CUDA_OR(cudaFree(fresh), cudaStreamSynchronize(s->ctx->stream));  // queries in flight still read the old boxes
  cudaFree(s->d_boxes);
  s->d_boxes = fresh;
  s->nBoxes = n;
  return ORZ_OK;
}
extern "C" int orz_scene_get_occluders(orz_scene* s, uint32_t* nOccluders, float* centers, float* boundsMin, float* boundsMax,
                                       uint32_t* quadCounts) {
  if (!s) return fail(ORZ_ERR_ARG, "orz_scene_get_occluders: scene is NULL");
  if (nOccluders) *nOccluders = s->nOcc;
  if (!centers && !boundsMin && !boundsMax && !quadCounts) return ORZ_OK;
  ORZ_CUDA(cudaSetDevice(s->ctx->device));
  std::vector<OccMeta> meta(s->nOcc);
  ORZ_CUDA(cudaMemcpyAsync(meta.data(), s->d_occ, s->nOcc * sizeof(OccMeta), cudaMemcpyDeviceToHost, s->ctx->stream));
  ORZ_CUDA(cudaStreamSynchronize(s->ctx->stream));
  for (uint32_t i = 0; i < s->nOcc; ++i) {
    if (centers) memcpy(centers + 4 * i, meta[i].center, 16);
    if (boundsMin) memcpy(boundsMin + 4 * i, meta[i].boundsMin, 16);
    if (boundsMax) memcpy(boundsMax + 4 * i, meta[i].boundsMax, 16);
    if (quadCounts) quadCounts[i] = meta[i].quadCount;
  }
  return ORZ_OK;
}
extern "C" uint32_t orz_scene_occludee_count(orz_scene* s) { return s ? s->nBoxes : 0; }
// ---- cached baked scene: the HBM layout written to / read from a file -----------
namespace {
struct SceneFileHeader {
  char magic[8];  // "ORZBAKE1"
  uint32_t nOcc, totalQuads, nBoxes, metaBytes;
};
}  // namespace
extern "C" int orz_scene_save(orz_scene* s, const char* path) {
  if (!s || !path) return fail(ORZ_ERR_ARG, "orz_scene_save: bad arguments");
  ORZ_CUDA(cudaSetDevice(s->ctx->device));
  std::vector<OccMeta> meta(s->nOcc);
  std::vector<uint4> quads(s->totalQuads);
  std::vector<float4> boxes((size_t)s->nBoxes * 2);
  ORZ_CUDA(cudaMemcpyAsync(meta.data(), s->d_occ, meta.size() * sizeof(OccMeta), cudaMemcpyDeviceToHost, s->ctx->stream));
  if (!quads.empty()) ORZ_CUDA(cudaMemcpyAsync(quads.data(), s->d_quads, quads.size() * sizeof(uint4), cudaMemcpyDeviceToHost, s->ctx->stream));
  if (!boxes.empty()) ORZ_CUDA(cudaMemcpyAsync(boxes.data(), s->d_boxes, boxes.size() * sizeof(float4), cudaMemcpyDeviceToHost, s->ctx->stream));
  ORZ_CUDA(cudaStreamSynchronize(s->ctx->stream));
  FILE* f = fopen(path, "wb");
  if (!f) return fail(ORZ_ERR_ARG, std::string("orz_scene_save: cannot open ") + path);
  SceneFileHeader h{{'O', 'R', 'Z', 'B', 'A', 'K', 'E', '1'}, s->nOcc, s->totalQuads, s->nBoxes, (uint32_t)sizeof(OccMeta)};
  bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(meta.data(), sizeof(OccMeta), meta.size(), f) == meta.size() &&
            fwrite(quads.data(), sizeof(uint4), quads.size(), f) == quads.size() &&
            fwrite(boxes.data(), sizeof(float4), boxes.size(), f) == boxes.size();
  ok = (fclose(f) == 0) && ok;
  return ok ? ORZ_OK : fail(ORZ_ERR_ARG, std::string("orz_scene_save: short write to ") + path);
}
extern "C" int orz_scene_load(orz_context* ctx, const char* path, orz_scene** out) {
  if (!ctx || !path || !out) return fail(ORZ_ERR_ARG, "orz_scene_load: bad arguments");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(ORZ_ERR_ARG, std::string("orz_scene_load: cannot open ") + path);
  SceneFileHeader h;
  std::vector<OccMeta> meta;
  std::vector<uint4> quads;
  std::vector<float4> boxes;
  bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "ORZBAKE1", 8) == 0 && h.metaBytes == sizeof(OccMeta) && h.nOcc > 0;
  if (ok) {
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    const uint64_t want = sizeof h + (uint64_t)h.nOcc * sizeof(OccMeta) + (uint64_t)h.totalQuads * sizeof(uint4) + (uint64_t)h.nBoxes * 2 * sizeof(float4);
    ok = size >= 0 && (uint64_t)size == want;  // sizes in the header must describe exactly this file
    fseek(f, sizeof h, SEEK_SET);
  }
  if (ok) {
    meta.resize(h.nOcc); quads.resize(h.totalQuads); boxes.resize((size_t)h.nBoxes * 2);
    ok = fread(meta.data(), sizeof(OccMeta), meta.size(), f) == meta.size() && fread(quads.data(), sizeof(uint4), quads.size(), f) == quads.size() &&
         fread(boxes.data(), sizeof(float4), boxes.size(), f) == boxes.size();
  }
  fclose(f);
  for (size_t i = 0; ok && i < meta.size(); ++i)  // every batch must lie inside the quad array
    ok = (uint64_t)meta[i].quadOffset + meta[i].quadCount <= h.totalQuads;
  if (!ok) return fail(ORZ_ERR_ARG, std::string("orz_scene_load: not a baked scene file (or truncated): ") + path);
  ORZ_CUDA(cudaSetDevice(ctx->device));
  orz_scene* s = new orz_scene();
  s->ctx = ctx; s->nOcc = h.nOcc; s->totalQuads = h.totalQuads; s->nBoxes = h.nBoxes;
  cudaError_t e = cudaMalloc(&s->d_quads, std::max<size_t>(quads.size(), 1) * sizeof(uint4));
  if (e == cudaSuccess) e = cudaMalloc(&s->d_occ, meta.size() * sizeof(OccMeta));
  if (e == cudaSuccess && h.nBoxes) e = cudaMalloc(&s->d_boxes, boxes.size() * sizeof(float4));
  if (e == cudaSuccess && !quads.empty()) e = cudaMemcpy(s->d_quads, quads.data(), quads.size() * sizeof(uint4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(s->d_occ, meta.data(), meta.size() * sizeof(OccMeta), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && h.nBoxes) e = cudaMemcpy(s->d_boxes, boxes.data(), boxes.size() * sizeof(float4), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(s->d_quads); cudaFree(s->d_occ); cudaFree(s->d_boxes);
    delete s;
    return fail(ORZ_ERR_CUDA, std::string("orz_scene_load: ") + cudaGetErrorString(e));
  }
  *out = s;
  return ORZ_OK;
}
// Main.cpp:56-84 + 86-128: the reference's raw scene files (uint32 triangle indices, float4 vertices) -> baked scene in HBM
extern "C" int orz_scene_from_mesh_files(orz_context* ctx, const char* indexPath, const char* vertexPath, uint32_t targetSize,
                                         uint32_t splitGranularity, int occludeesFromQuads, orz_scene** out, orz_mesh_scene_info* info) {
  if (!ctx || !indexPath || !vertexPath || !out) return fail(ORZ_ERR_ARG, "orz_scene_from_mesh_files: bad arguments");
  auto slurp = [](const char* path, std::vector<char>& data) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    data.resize(size > 0 ? (size_t)size : 0);
    const bool ok = fread(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    return ok;
  };
  std::vector<char> idx, vtx;
  if (!slurp(indexPath, idx)) return fail(ORZ_ERR_ARG, std::string("orz_scene_from_mesh_files: cannot read ") + indexPath);
  if (!slurp(vertexPath, vtx)) return fail(ORZ_ERR_ARG, std::string("orz_scene_from_mesh_files: cannot read ") + vertexPath);
  return orz_scene_from_mesh(ctx, reinterpret_cast<const uint32_t*>(idx.data()), idx.size() / 4, reinterpret_cast<const float*>(vtx.data()),
                             vtx.size() / 16, targetSize, splitGranularity, occludeesFromQuads, out, info);
}
extern "C" void orz_scene_destroy(orz_scene* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->d_quads); cudaFree(s->d_occ); cudaFree(s->d_boxes);
  delete s;
}

// queryVisibility of every occludee box for views [pg.viewBase, pg.viewBase + pg.groupViews) of the (sorted) batch:
// 1-D grids of (view, 256-box chunk) CTAs, in slabs of views when one grid cannot hold them all
static int launch_query(orz_context* ctx, FrameParams pg, uint32_t nBoxes, cudaStream_t st) {
  pg.queryChunks = (nBoxes + kQueryThreads - 1u) / kQueryThreads;
  if (pg.queryChunks == 0u || pg.groupViews == 0u) return ORZ_OK;
  const uint32_t slab = std::max<uint32_t>(1u, 0x7fffffffu / pg.queryChunks);
  const uint32_t first = pg.viewBase, last = pg.viewBase + pg.groupViews;
  for (uint32_t v = first; v < last; v += slab) {
    pg.viewBase = v;
    pg.groupViews = std::min(slab, last - v);
    k_query_views<<<pg.groupViews * pg.queryChunks, kQueryThreads, 0, st>>>(pg);
    ctx->launches++;
    ORZ_CUDA(cudaGetLastError());
  }
  return ORZ_OK;
}

template <int GW, int kTrav>
static int launch_views_t(orz_context* ctx, const FrameParams& p, uint32_t grid, cudaStream_t st) {
  constexpr int kLog = GW == 1 ? 0 : GW == 2 ? 1 : GW == 4 ? 2 : GW == 8 ? 3 : 4;
  if (!ctx->smemViews[kTrav - 1][kLog]) {
    ORZ_CUDA(cudaFuncSetAttribute(k_render_views<GW, kTrav>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FrameSmem<GW, kTrav>::kBytes));
    ctx->smemViews[kTrav - 1][kLog] = FrameSmem<GW, kTrav>::kBytes;
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

static uint32_t tiles_of(uint32_t width, uint32_t height, uint32_t tileH) {
  return (((width >> 3) + kTileW - 1u) / kTileW) * (((height >> 3) + tileH - 1u) / tileH);
}
template <int C, uint32_t TH>
static int launch_cluster_t(orz_context* ctx, FrameParams p, uint32_t nViews, cudaStream_t st) {
  const uint32_t nTiles = tiles_of(p.width, p.height, TH);
  p.clusterK = (nTiles + (uint32_t)(C * kClusterGW) - 1u) / (uint32_t)(C * kClusterGW);
  const size_t smem = ClusterSmem::bytes(p.clusterK, p.nOcc, nTiles);
  constexpr int kLog = (C == 1 ? 0 : C == 2 ? 1 : C == 4 ? 2 : C == 8 ? 3 : 4) + (TH == 1 ? 5 : 0);
  if (ctx->smemCluster[kLog] < smem) {
    ORZ_CUDA(cudaFuncSetAttribute(k_raster_views_cluster<C, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (C > 8) ORZ_CUDA(cudaFuncSetAttribute(k_raster_views_cluster<C, TH>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    ctx->smemCluster[kLog] = smem;
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
  ORZ_CUDA(cudaLaunchKernelEx(&cfg, k_raster_views_cluster<C, TH>, p));
  ctx->launches++;
  return ORZ_OK;
}
// one cluster per view.  Cluster size: as many CTAs as the view can use (one tile per warp) while all
// views of the batch still fit the GPU in one wave; at least enough that a warp owns <= 32 tiles and the CTA's
// shared memory (tables, staging, the open tiles, decision words of every occluder, HiZ mirror) fits.
// Tile height: 8 x 4 blocks; 8 x 1 strips (four times the tiles to hand out, a quarter of the stacked updates per tile)
// can be asked for with orz_context_set_tile_height.
// Returns 0 when no cluster size works: the caller stays on the large-batch kernel.
struct ClusterShape { uint32_t c, tileH; };
static ClusterShape pick_cluster_shape(const orz_context* ctx, uint32_t width, uint32_t height, uint32_t nOcc, uint32_t nBatch) {
  auto tilesPerWarp = [&](uint32_t c, uint32_t th) { return (tiles_of(width, height, th) + c * kClusterGW - 1u) / (c * kClusterGW); };
  auto fits = [&](uint32_t c, uint32_t th) { return tilesPerWarp(c, th) <= 32u && ClusterSmem::bytes(tilesPerWarp(c, th), nOcc, tiles_of(width, height, th)) <= ctx->maxSmemOptin; };
  const uint32_t nTiles = tiles_of(width, height, kTileH);
  uint32_t c = 1;
  while (c < 16u && c * kClusterGW < nTiles && nBatch * c * 2u <= (uint32_t)ctx->numSMs) c *= 2u;
  if (ctx->clusterSize) c = (uint32_t)ctx->clusterSize;
  else while (c < 16u && !fits(c, kTileH)) c *= 2u;
  uint32_t th = kTileH;
  if (ctx->clusterTileH == 1) {
    // strips (only on request: measured 7-20 % SLOWER than 8 x 4 tiles for one view -- every strip pays the HiZ test, the
    // chains and the coverage test of a primitive again, and that overhead is what the busiest warp's time is made of),
    // on the largest cluster that still leaves every warp a tile
    uint32_t c1 = ctx->clusterSize ? c : 16u;
    while (!ctx->clusterSize && c1 > 1u && (c1 / 2u) * kClusterGW >= tiles_of(width, height, 1u)) c1 /= 2u;
    if (fits(c1, 1u)) { c = c1; th = 1u; }
  }
  ClusterShape sh = {fits(c, th) ? c : 0u, th};
  return sh;
}
static uint32_t pick_cluster_size(const orz_context* ctx, uint32_t width, uint32_t height, uint32_t nOcc, uint32_t nBatch) {
  return pick_cluster_shape(ctx, width, height, nOcc, nBatch).c;
}
template <uint32_t TH>
static int launch_cluster_c(orz_context* ctx, const FrameParams& p, uint32_t c, uint32_t nViews, cudaStream_t st) {
  const uint32_t nTiles = tiles_of(p.width, p.height, TH);
  switch (c) {
    case 16:  // non-portable cluster size: when the device (e.g. a partitioned one) cannot place it, use 8
      if (launch_cluster_t<16, TH>(ctx, p, nViews, st) == ORZ_OK) return ORZ_OK;
      (void)cudaGetLastError();
      if ((nTiles + 8u * kClusterGW - 1u) / (8u * kClusterGW) > 32u || ClusterSmem::bytes((nTiles + 8u * kClusterGW - 1u) / (8u * kClusterGW), p.nOcc, nTiles) > ctx->maxSmemOptin)
        return fail(ORZ_ERR_CUDA, "cluster path: 16-CTA clusters are not available on this device");
      return launch_cluster_t<8, TH>(ctx, p, nViews, st);
    case 8: return launch_cluster_t<8, TH>(ctx, p, nViews, st);
    case 4: return launch_cluster_t<4, TH>(ctx, p, nViews, st);
    case 2: return launch_cluster_t<2, TH>(ctx, p, nViews, st);
    default: return launch_cluster_t<1, TH>(ctx, p, nViews, st);
  }
}
static int launch_cluster(orz_context* ctx, const FrameParams& p, uint32_t nViews, uint32_t nBatch, cudaStream_t st) {
  const ClusterShape sh = pick_cluster_shape(ctx, p.width, p.height, p.nOcc, nBatch);
  if (!sh.c) return fail(ORZ_ERR_ARG, "cluster path: target too large (or too many occluders) for this cluster size");
  return sh.tileH == 1u ? launch_cluster_c<1>(ctx, p, sh.c, nViews, st) : launch_cluster_c<kTileH>(ctx, p, sh.c, nViews, st);
}

// speculative setup of every (occluder in the frustum, view) of a launch: one CTA each
static void launch_setup(const FrameParams& p, uint32_t nViews, cudaStream_t st) {
  if ((size_t)p.nOcc * nViews >= (size_t)ORZ_SETUP_SMALL_CTAS) k_setup_views<32, 32><<<dim3(p.nOcc, nViews), 32, 0, st>>>(p);
  else k_setup_views<256, 4><<<dim3(p.nOcc, nViews), 256, 0, st>>>(p);
}

// Error paths between the fork of the auxiliary streams and their join must not leave work running on them that later
// calls (which reuse or free the scratch buffers) know nothing about: the guard drains them unless the join was reached.
namespace {
struct AuxDrain {
  orz_context* ctx;
  bool armed = false;
  ~AuxDrain() {
    if (!armed) return;
    for (int g = 0; g < orz_context::kGroups; ++g) cudaStreamSynchronize(ctx->aux[g]);
  }
};
}  // namespace

// Device-pointer entry: three launches per chunk of views (prepare, render, query).  When the
// caller does not ask for depth/HiZ, per-view targets live in an internal arena and the batch is
// processed in chunks that fit the arena budget.
extern "C" int orz_render_views_device(orz_context* ctx, orz_scene* scene, const orz_view_batch* b) {
  if (!ctx || !scene || !b || !b->mvps || (!b->orders && !b->camPos)) return fail(ORZ_ERR_ARG, "orz_render_views_device: bad arguments");
  if (scene->ctx != ctx) return fail(ORZ_ERR_ARG, "orz_render_views_device: the scene belongs to another context");
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

  // views per chunk: everything when the caller owns the targets, else what fits the arena budget; batches that take the
  // cluster path (speculative setup records of every view: ~2.3 MB per Castle view) also stay within the record budget
  size_t chunk = b->nViews;
  const size_t recPerView = (size_t)scene->totalQuads * (blocks > 65536u ? 2 : 1) * (kRecStride * 4 + 8);
  const size_t recBudget = ctx->recordBudget;  // (8 GB; ORZ_RECORD_BUDGET_GB=24 makes 8 192 Castle probes one chunk: measured 1 % faster for 19 GB of scratch)
  if (ctx->clusterViews > 0 && b->nViews <= (size_t)ctx->clusterViews && recPerView > 0 && recBudget / recPerView >= 256)
    chunk = std::min<size_t>(chunk, recBudget / recPerView);
  if (chunk < b->nViews) chunk = (b->nViews + (b->nViews + chunk - 1) / chunk - 1) / ((b->nViews + chunk - 1) / chunk);  // equal chunks
  if (ownTargets) {
    const size_t perView = blocks * 128 + hizStride * 2;
    const size_t budget = std::max<size_t>(ctx->arenaBudget, perView);
    chunk = std::max<size_t>(1, std::min<size_t>(chunk, budget / perView));
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
    // few ungated views over very many quads (BASELINE config 4): ONE view at a time over the whole GPU, tile major
    // (k_raster_tiles); ORZ_BATCH_WIDE keeps the round-1 row walk (k_raster_wide) selectable for comparison
    const bool wide = (b->flags & ORZ_BATCH_NO_GATE) && (b->flags & ORZ_BATCH_WIDE);
    const bool tilesPath = (b->flags & ORZ_BATCH_NO_GATE) && !wide && nv <= 8u && scene->totalQuads >= 65536u &&
                           ((size_t)((b->width / 8 + kTileW - 1) / kTileW) * ((b->height / 8 + kTileH - 1) / kTileH) + (size_t)ctx->numSMs * kClusterGW - 1) /
                                   ((size_t)ctx->numSMs * kClusterGW) <= 32u;
    // (targets above 32 tiles per warp of a 16-CTA cluster -- beyond ~8K x 4K -- stay on the batch kernel)
    // measured crossover with the batch kernel: ~2000 views at 1920x1080, ~1500 at 512x256 (profiles/r1_few_views_*)
    const uint32_t clusterLimit = std::min<uint32_t>((uint32_t)ctx->clusterViews, 65535u);  // k_setup_views puts the view on grid.y
    const size_t nTilesC = (size_t)((b->width / 8 + kTileW - 1) / kTileW) * ((b->height / 8 + kTileH - 1) / kTileH);
    const bool clusterPath = !wide && !tilesPath && ctx->clusterViews > 0 && nv <= clusterLimit && nTilesC <= 32u * 16u * kClusterGW && nOcc <= kClusterMaxOcc &&
                             pick_cluster_size(ctx, b->width, b->height, (uint32_t)nOcc, nv) != 0u &&
                             (size_t)nv * recPerView <= recBudget;
    p.viewOrder = (nv <= 16384u && !tilesPath && !(clusterPath && nv * 2u <= (uint32_t)ctx->numSMs)) ? p.viewCost + chunk : nullptr;
    if (!p.orders && nOcc > 1024u) {  // many occluders: order them on the whole GPU (keys live in the front buffer until k_prepare_views overwrites it)
      float* keys = reinterpret_cast<float*>(p.frontBuf);
      const size_t nKeys = (size_t)nv * nOcc;
      const uint32_t chunksPerView = (uint32_t)((nOcc + 255u) / 256u);
      if (nKeys / 256u + 1u > 0x7fffffffull || (size_t)nv * chunksPerView > 0x7fffffffull) return fail(ORZ_ERR_ARG, "batch too large for the device-side occluder sort");
      k_order_keys<<<(uint32_t)((nKeys + 255u) / 256u), 256, 0, ctx->stream>>>(p, keys);
      k_order_rank<<<nv * chunksPerView, 256, 0, ctx->stream>>>(p, keys, chunksPerView);
      ctx->launches += 2;
      ORZ_CUDA(cudaGetLastError());
      p.orders = p.orderBuf;
    }
    k_prepare_views<<<nv, nv <= 16u ? 512 : 128, 0, ctx->stream>>>(p);
    ctx->launches++;
    ORZ_CUDA(cudaGetLastError());
    if (p.viewOrder) {
      k_sort_views<<<(nv + 7) / 8, 256, 0, ctx->stream>>>(p.viewCost, nv, p.viewOrder);
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
        if ((e = launch_query(ctx, pq, scene->nBoxes, ctx->stream))) return e;
      }
      continue;
    }
    if (tilesPath) {
      const size_t recSlots = (size_t)scene->totalQuads * (blocks > 65536u ? 2 : 1);  // a wrapped primitive is two records
      const size_t recBytes = recSlots * kRecStride * 4, hdrBytes = (recSlots * 8 + 15) & ~size_t(15), infoBytes = nOcc * 32, boxBytes = (nOcc * 8 + 15) & ~size_t(15);
      if ((e = ensure_scratch(ctx, 11, recBytes + hdrBytes + infoBytes + boxBytes + 64))) return e;
      const uint32_t nTiles = (uint32_t)nTilesC;
      const uint32_t grid = std::min<uint32_t>((uint32_t)ctx->numSMs, (nTiles + kClusterGW - 1u) / kClusterGW);
      const uint32_t K = (nTiles + grid * kClusterGW - 1u) / (grid * kClusterGW);
      const size_t smem = TilesSmem::bytes(K);
      if (ctx->smemTiles < smem) {
        ORZ_CUDA(cudaFuncSetAttribute(k_raster_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->smemTiles = smem;
      }
      for (uint32_t v = 0; v < nv; ++v) {  // the scratch holds ONE view's records: the views follow each other on the stream
        FrameParams pv = p;
        pv.nViews = 1; pv.viewOrder = nullptr; pv.viewBase = 0; pv.groupViews = 1; pv.clusterK = K;
        pv.frontBuf = p.frontBuf + (size_t)v * nOcc * kFrontWords;
        if (p.orders) pv.orders = p.orders + (size_t)v * nOcc; else pv.orderBuf = p.orderBuf + (size_t)v * nOcc;
        pv.depth = p.depth + (size_t)v * p.depthStride; pv.hiz = p.hiz + (size_t)v * p.hizStride;
        pv.quadsSubmitted = p.quadsSubmitted ? p.quadsSubmitted + v : nullptr;
        pv.hdrBuf = (uint2*)ctx->d_scratch[11];
        pv.recInfo = (uint4*)((uint8_t*)ctx->d_scratch[11] + hdrBytes);
        pv.occBox = (uint2*)((uint8_t*)pv.recInfo + infoBytes);
        pv.recBuf = (uint32_t*)((uint8_t*)pv.occBox + boxBytes);
        pv.totalQuads = (uint32_t)recSlots;
        launch_setup(pv, 1u, ctx->stream);
        k_raster_tiles<<<grid, kClusterGW * 32, smem, ctx->stream>>>(pv, 0u, scene->totalQuads);
        ctx->launches += 2;
        ORZ_CUDA(cudaGetLastError());
      }
      if (p.gate) ORZ_CUDA(cudaMemsetAsync(p.gate, 1, (size_t)nv * nOcc, ctx->stream));  // no gate: every occluder is submitted
      if (p.visBits || p.clipBits) {
        FrameParams pq = p;
        pq.viewOrder = nullptr; pq.viewBase = 0; pq.groupViews = nv;
        if ((e = launch_query(ctx, pq, scene->nBoxes, ctx->stream))) return e;
      }
      continue;
    }
    // Few views: one thread-block cluster per view (latency path, BASELINE configs 1 and 2)
    if (clusterPath) {
      FrameParams pc = p;
      pc.viewBase = 0; pc.groupViews = nv;
      const size_t recSlots = (size_t)scene->totalQuads * (blocks > 65536u ? 2 : 1);  // a wrapped primitive is two records
      const size_t recBytes = (size_t)nv * recSlots * kRecStride * 4, hdrBytes = (size_t)nv * recSlots * 8;
      if ((e = ensure_scratch(ctx, 11, recBytes + hdrBytes + (size_t)nv * nOcc * 32 + 64))) return e;
      pc.hdrBuf = (uint2*)ctx->d_scratch[11];
      pc.recInfo = (uint4*)((uint8_t*)ctx->d_scratch[11] + ((hdrBytes + 15) & ~size_t(15)));
      pc.recBuf = (uint32_t*)((uint8_t*)pc.recInfo + (size_t)nv * nOcc * 32);
      pc.totalQuads = (uint32_t)recSlots;
      pc.occBox = nullptr;
      if ((p.visBits || p.clipBits) && ctx->coarseQuery) {  // per-tile HiZ minima for the occludee queries
        const ClusterShape sh = pick_cluster_shape(ctx, b->width, b->height, (uint32_t)nOcc, nv);
        const uint32_t nTilesQ = tiles_of(b->width, b->height, sh.tileH);
        if ((e = ensure_scratch(ctx, 12, (size_t)nv * nTilesQ * 2))) return e;
        pc.coarseHiz = (uint16_t*)ctx->d_scratch[12];
        pc.coarseStride = nTilesQ;
        pc.coarseCellH = sh.tileH;
      }
      // Large batches: sub-batches (by descending cost) on auxiliary streams, so that the occludee queries of a
      // finished sub-batch fill the SMs the cluster kernel of the next one leaves idle; each sub-batch sets up its
      // own views first, so the heaviest views start rasterising after a quarter of the setup work.
      const bool wantQ = p.visBits || p.clipBits;
      const int groupsC = (pc.viewOrder && wantQ && nv >= 256u) ? orz_context::kGroups : 1;
      if (groupsC > 1) ORZ_CUDA(cudaEventRecord(ctx->evFork, ctx->stream));
      AuxDrain drain{ctx, groupsC > 1};
      for (int g = 0; g < groupsC; ++g) {
        cudaStream_t st = groupsC > 1 ? ctx->aux[g] : ctx->stream;
        if (groupsC > 1) ORZ_CUDA(cudaStreamWaitEvent(st, ctx->evFork, 0));
        FrameParams pg = pc;
        // cumulative shares of the sub-batches (per mille): the first one's setup and the last one's queries have nothing
        // to overlap with, so the ends are smaller than the middle
        pg.viewBase = groupsC > 1 ? (uint32_t)((uint64_t)nv * ctx->groupCut[g] / 1000u) : 0u;
        pg.groupViews = (groupsC > 1 ? (uint32_t)((uint64_t)nv * ctx->groupCut[g + 1] / 1000u) : nv) - pg.viewBase;
        if (pg.groupViews == 0u) continue;
        launch_setup(pg, pg.groupViews, st);
        ctx->launches++;
        ORZ_CUDA(cudaGetLastError());
        if ((e = launch_cluster(ctx, pg, pg.groupViews, nv, st))) return e;
        if (wantQ && (e = launch_query(ctx, pg, scene->nBoxes, st))) return e;
        if (groupsC > 1) {
          ORZ_CUDA(cudaEventRecord(ctx->evJoin[g], st));
          ORZ_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evJoin[g], 0));
        }
      }
      drain.armed = false;
      continue;
    }
    // Sub-batches (by descending cost) on auxiliary streams: the query kernel of a finished
    // sub-batch and the CTAs of the next one fill the SMs that the drain of the previous render
    // kernel leaves idle.  Everything is fenced by events on the context stream.
    const bool wantQuery = p.visBits || p.clipBits;
    const int groups = (p.viewOrder && nv >= 64u) ? orz_context::kGroups : 1;
    ORZ_CUDA(cudaMemsetAsync(ctx->d_counter, 0, 4 * orz_context::kGroups, ctx->stream));
    if (groups > 1) ORZ_CUDA(cudaEventRecord(ctx->evFork, ctx->stream));
    AuxDrain drain{ctx, groups > 1};
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
      if (wantQuery && (e = launch_query(ctx, pg, scene->nBoxes, st))) return e;
      if (groups > 1) {
        ORZ_CUDA(cudaEventRecord(ctx->evJoin[g], st));
        ORZ_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evJoin[g], 0));
      }
    }
    drain.armed = false;
  }
  return ORZ_OK;
}

// Host-pointer variant: stage inputs to HBM, render, bring the requested outputs back.
extern "C" int orz_render_views(orz_context* ctx, orz_scene* scene, const orz_view_batch* hb) {
  if (!ctx || !scene || !hb || !hb->mvps || (!hb->orders && !hb->camPos)) return fail(ORZ_ERR_ARG, "orz_render_views: bad arguments");
  if (scene->ctx != ctx) return fail(ORZ_ERR_ARG, "orz_render_views: the scene belongs to another context");
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

// ---- multi-GPU: NCCL all-gather of the per-view visibility bitmasks behind the C ABI
#include "orz_comm.inl"
