// Cluster path: speculative setup (k_setup_views) + dataflow cluster rasteriser (k_raster_views_cluster).
// Part of the single translation unit orz_kernels.cu (included inside namespace orz); see DESIGN.md section 4.
#pragma once

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
#ifndef ORZ_ROUNDS_X2
#define ORZ_ROUNDS_X2 0  // 1: the eight-lanes-per-block update takes two covered blocks per group and pass (needs ORZ_HIZ_ATOMIC and ORZ_COVERED_WORD; measured neutral to 1.5 % slower, profiles/r2y_*)
#endif
#ifndef ORZ_CHAIN_MERGED
#define ORZ_CHAIN_MERGED 1  // depth chains are stepped together with the edge chains: one pass, two independent add sequences per lane (0: depth chains only after the coverage test; measured 4-10 % slower, profiles/r2_variants.txt)
#endif
#ifndef ORZ_BLOCK_BOUND_SKIP
#define ORZ_BLOCK_BOUND_SKIP 0  // 1: skip the block updates whose corner bound proves they change nothing (exact, every parity suite passes; measured 1-2 % SLOWER on Castle and Sponza: the test costs more than the skipped passes save)
#endif
#ifndef ORZ_CHAIN_F32X2
#define ORZ_CHAIN_F32X2 1  // 1: the two chains of a lane advance with one packed add.rn.f32x2 (sm_100: two IEEE single adds in one instruction)
#endif
#ifndef ORZ_HIZ_ATOMIC
#define ORZ_HIZ_ATOMIC 1  // 1: a block's new HiZ is folded with one shared-memory atomicMin per lane instead of three shuffles per pass
#endif
#ifndef ORZ_COVERED_WORD
#define ORZ_COVERED_WORD 1  // 1: the covered-block list is stored pass-major so that a group fetches its (up to) eight blocks with ONE 64-bit load
#endif
#ifndef ORZ_LUT_BULK
#define ORZ_LUT_BULK 1  // 1: the 32 KB edge-mask table comes in with ONE cp.async.bulk (TMA engine, mbarrier complete_tx) instead of a copy loop
#endif
#ifndef ORZ_ASYNC_GATHER
#define ORZ_ASYNC_GATHER 1  // 1: staged records are gathered with cp.async (no registers, no scoreboard wait): the L2 round trip overlaps the tile test and the first tile's open
#endif
#ifndef ORZ_LOOKAHEAD
#define ORZ_LOOKAHEAD 0  // candidates further down the order a warp tries to answer "no" early after it has rasterised an occluder (0: off; exact -- every parity suite passes with 8 -- but measured: Sponza 256 views 2-3 % faster, Castle equal, probes 8 % and single views 3 % slower: profiles/r2ay_*)
#endif
#ifndef ORZ_LOOKAHEAD_WAITING
#define ORZ_LOOKAHEAD_WAITING 1  // ... and before it starts waiting for a decision
#endif
#ifndef ORZ_WAIT_STATS
#define ORZ_WAIT_STATS 0
#endif
#ifndef ORZ_TILE_MAP
#define ORZ_TILE_MAP 0  // 1: the header scan looks a record's tile rectangle up in a per-warp bitmap of owned tiles instead of testing every owned tile of the occluder (measured: batches equal, one view 11 % slower -- few tiles per warp there: profiles/r2av_*)
#endif
#ifndef ORZ_TILE_PREFETCH
#define ORZ_TILE_PREFETCH 1  // flush: prefetch the next tile's depth blocks into L1 while the current tile is processed
#endif
#ifndef ORZ_FLAG_VOTE
#define ORZ_FLAG_VOTE 1  // decision words are observed through a warp vote (uniform control flow around the collectives)
#endif
#ifndef ORZ_HDR_PREFETCH_EARLY
#define ORZ_HDR_PREFETCH_EARLY 1  // record headers of a candidate are prefetched before its gate test (0: after the decision)
#endif
#ifndef ORZ_CLUSTER_REGS
#define ORZ_CLUSTER_REGS 128  // 16 warps x 128 registers = the whole register file: measured faster than leaving room for a query CTA (96) in every case (profiles/r2_variants.txt)
#endif
constexpr int kClusterGW = ORZ_CLUSTER_GW;  // warps per CTA of the cluster kernel; registers per thread capped so that they fit one SM
constexpr uint32_t kTileW = 8, kTileH = 4;   // blocks per tile: lane = 8 * (row in tile) + column in tile
constexpr uint32_t kChainStride = 33;        // words between two chains' slots: the publishing lanes (chain, tile row) hit 32 different banks
#ifndef ORZ_STAGE_CAP
#define ORZ_STAGE_CAP 32
#endif
constexpr uint32_t kStageCap = ORZ_STAGE_CAP;  // records a warp stages at a time (<= 32: one lane per staged record)
constexpr uint32_t kClusterMaxOcc = 2048;    // occluders per scene the cluster path accepts (shared-memory decision arrays)
constexpr uint32_t kHeadWords = 6;           // status, minX, maxX, minY, maxY, maxZ of kFrontWords

constexpr uint32_t kTileWords = 32u * 32u;     // the open tile's depth: 32 blocks x 128 B, [block][item (rr, i), swizzled][row pair k]
constexpr uint32_t kTileAuxWords = 64u + 32u + 8u;  // + per block: 64-bit coverage mask of the current primitive, HiZ after the update; covered blocks in order (bytes)

struct ClusterSmem {
  static constexpr uint32_t kLutWords = ORZ_CLUSTER_LUT_SMEM ? 4096 * 2 : 0;
  static constexpr uint32_t kTileAllWords = kClusterGW * (kTileWords + kTileAuxWords);
  static constexpr uint32_t kStageWords = kClusterGW * kStageCap * kRecStride;
  static constexpr uint32_t kIdxWords = kClusterGW * kStageCap;
  static constexpr uint32_t kChainWords = kClusterGW * 12 * kChainStride;
  static constexpr uint32_t kFixedWords = kLutWords + kTileAllWords + kStageWords + kIdxWords + kChainWords;
  // + [nOcc][6] gate heads, 3 x [nOcc] decision words, [GW][K][32] u16 HiZ mirror, [GW][ceil(nTiles / 32)] tile-ownership bitmaps
  static size_t bytes(uint32_t tilesPerWarp, uint32_t nOcc, uint32_t nTiles) {
    return (size_t)(kFixedWords + nOcc * (kHeadWords + 3u)) * 4 + (size_t)kClusterGW * tilesPerWarp * 32 * 2 +
           (ORZ_TILE_MAP ? (size_t)kClusterGW * ((nTiles + 31u) / 32u) * 4 : 0);
  }
};

// ---- speculative setup of every occluder that survives the frustum, Rasterizer.cpp:657-1086
// CTA size: 256 threads when the launch is small (one or a few views: the setup sits on the view's critical path and an
// occluder's ~335 quads should be set up in two rounds), ONE WARP per CTA for the batches (no block barrier at all, 32
// resident CTAs per SM; 64 / 96 / 128 / 256 / 512 measured: profiles/r2ap-r2ar_*)
template <uint32_t kSetupThreads, int kCtasPerSM>
__global__ void __launch_bounds__(kSetupThreads, kCtasPerSM) k_setup_views(const FrameParams p) {
  constexpr uint32_t kSetupWarps = kSetupThreads / 32u;
  __shared__ uint32_t s_cnt[kSetupWarps];
  __shared__ uint32_t s_box[4];
  // grid: (order slot, rank of the view in this launch's group of the cost-sorted batch)
  const uint32_t slot = blockIdx.x, vrank = p.viewBase + blockIdx.y, tid = threadIdx.x;
  const uint32_t view = p.viewOrder ? p.viewOrder[vrank] : vrank;
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31u);
  const uint32_t* fr = p.frontBuf + ((size_t)view * p.nOcc + slot) * kFrontWords;
  const uint32_t status = fr[0];
  if (status == kBoxCulled) {
    if (tid == 0) {
      p.recInfo[((size_t)view * p.nOcc + slot) * 2u] = make_uint4(0u, 0u, 0u, 0u);
      if (p.occBox) p.occBox[(size_t)view * p.nOcc + slot] = make_uint2(0xffffffffu, 0u);
    }
    return;
  }
  const bool useGate = (p.flags & ORZ_BATCH_NO_GATE) == 0u;
  const bool clipped = status == kBoxNearClip ? (useGate ? true : (p.flags & ORZ_BATCH_FORCE_CLIPPED) != 0u) : false;
  struct { uint32_t quadOffset, quadCount; } om = {fr[20], fr[21]};  // the occluder of this order slot (k_prepare_views)
  const RcpTable rt{p.rcp, p.rcpShift};
  CallMatrix cm;
#pragma unroll
  for (int k = 0; k < 4; ++k) { cm.rx[k] = u2f(fr[6 + k]); cm.ry[k] = u2f(fr[10 + k]); cm.rw[k] = u2f(fr[14 + k]); }
  cm.c0 = u2f(fr[18]); cm.c1 = u2f(fr[19]);
  const int32_t blocksX = (int32_t)(p.width >> 3), blocksY = (int32_t)(p.height >> 3);
  // Above 65 536 blocks the reference wraps the first-block index to 16 bits (Rasterizer.cpp:1054): a
  // primitive row then starts at linear block fb + by * blocksX, which can straddle two screen rows.
  // In screen space that is at most TWO rectangles sharing the primitive's chains: the second one
  // starts `split` blocks into every chain row (record word 20 = x steps to skip).
  const bool wrap = (uint32_t)blocksX * (uint32_t)blocksY > 65536u;
  const uint32_t recSlots = wrap ? 2u : 1u;  // records a quad can produce
  const uint32_t slotBase = om.quadOffset * recSlots;
  const size_t recBase = (size_t)view * p.totalQuads + slotBase;  // records of this (view, occluder) start here
  uint32_t* recs = p.recBuf + recBase * kRecStride;
  uint2* hdrs = p.hdrBuf + recBase;
  if (tid < 4) s_box[tid] = tid < 2 ? 0xffffffffu : 0u;
  uint32_t written = 0;
  uint32_t bx0 = 0xffffffffu, by0 = 0xffffffffu, bx1 = 0u, by1 = 0u;
  for (uint32_t q0 = 0; q0 < om.quadCount; q0 += kSetupThreads) {
    const uint32_t qi = q0 + tid;
    bool ok = false;
    Prim P;
    if (qi < om.quadCount) {
      const uint4 v = p.quads[om.quadOffset + qi];  // 128-bit coalesced load: the four packed vertices of this lane's quad
      const uint32_t word[4] = {v.x, v.y, v.z, v.w};
      ok = clipped ? setup_quad<true>(word, cm, rt, c_modeNibbles, blocksX, blocksY, P)
                   : setup_quad<false>(word, cm, rt, c_modeNibbles, blocksX, blocksY, P);
    }
    // screen rectangles of the primitive: {x, y, w, h, chain x offset}
    uint32_t nPieces = ok ? 1u : 0u;
    uint32_t px[2] = {0u, 0u}, py[2] = {0u, 0u}, pw[2] = {0u, 0u}, ph[2] = {0u, 0u}, pskip[2] = {0u, 0u};
    if (ok) {
      px[0] = (uint32_t)P.minX; py[0] = (uint32_t)P.minY; pw[0] = (uint32_t)P.rangeX; ph[0] = (uint32_t)P.rangeY;
      if (wrap) {
        const uint32_t fb = (((uint32_t)P.minY * (uint32_t)blocksX) & 0xffffu) + (uint32_t)P.minX;
        const uint32_t r0 = fb / (uint32_t)blocksX, c0 = fb - r0 * (uint32_t)blocksX;
        const uint32_t split = min((uint32_t)P.rangeX, (uint32_t)blocksX - c0);
        px[0] = c0; py[0] = r0; pw[0] = split;
        if (split < (uint32_t)P.rangeX) { nPieces = 2u; px[1] = 0u; py[1] = r0 + 1u; pw[1] = (uint32_t)P.rangeX - split; ph[1] = ph[0]; pskip[1] = split; }
        // rows past the target would be out-of-bounds writes in the reference: never produced by its own scenes, dropped here
        for (uint32_t k = 0; k < nPieces; ++k) ph[k] = py[k] < (uint32_t)blocksY ? min(ph[k], (uint32_t)blocksY - py[k]) : 0u;
        if (nPieces == 2u && ph[1] == 0u) nPieces = 1u;
        if (ph[0] == 0u) { nPieces = nPieces == 2u ? 1u : 0u; px[0] = px[1]; py[0] = py[1]; pw[0] = pw[1]; ph[0] = ph[1]; pskip[0] = pskip[1]; }
      }
    }
    // in-order slots: inclusive warp scan of the piece counts, block prefix over the 8 warps
    uint32_t incl = nPieces;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, dd); if (lane >= dd) incl += t; }
    __syncthreads();  // s_cnt of the previous chunk has been read
    if (lane == 31) s_cnt[warp] = incl;
    __syncthreads();
    uint32_t base = written, total = 0;
#pragma unroll
    for (int w2 = 0; w2 < (int)kSetupWarps; ++w2) { const uint32_t c = s_cnt[w2]; base += w2 < warp ? c : 0u; total += c; }
    for (uint32_t k = 0; k < nPieces; ++k) {  // binning by prefix-sum compaction
      const uint32_t at = base + incl - nPieces + k;
      uint32_t* rec = recs + (size_t)at * kRecStride;
      store_record(rec, P);
      rec[0] = px[k] | (py[k] << 16); rec[1] = pw[k] | (ph[k] << 16); rec[20] = pskip[k];
      hdrs[at] = make_uint2(rec[0], rec[1]);
      bx0 = min(bx0, px[k]); by0 = min(by0, py[k]);
      bx1 = max(bx1, px[k] + pw[k]); by1 = max(by1, py[k] + ph[k]);
    }
    written += total;
  }
  bx0 = __reduce_min_sync(kFull, bx0); by0 = __reduce_min_sync(kFull, by0);
  bx1 = __reduce_max_sync(kFull, bx1); by1 = __reduce_max_sync(kFull, by1);
  __syncthreads();
  if (lane == 0) { atomicMin(&s_box[0], bx0); atomicMin(&s_box[1], by0); atomicMax(&s_box[2], bx1); atomicMax(&s_box[3], by1); }
  __syncthreads();
  if (tid == 0) {
    p.recInfo[((size_t)view * p.nOcc + slot) * 2u] = make_uint4(written, slotBase, om.quadCount, 0u);
    // block rectangle that holds every primitive of the occluder, half open (lo > hi when there is none)
    p.recInfo[((size_t)view * p.nOcc + slot) * 2u + 1u] = make_uint4(s_box[0], s_box[1], s_box[2], s_box[3]);
    if (p.occBox) p.occBox[(size_t)view * p.nOcc + slot] = written ? make_uint2(s_box[0] | (s_box[1] << 16), s_box[2] | (s_box[3] << 16)) : make_uint2(0xffffffffu, 0u);
  }
}

// ---- asynchronous copies: cp.async (per-lane, 4 bytes: records are 84 bytes at arbitrary slots) and the bulk copy engine
__device__ __forceinline__ void cp_async_word(uint32_t* smemDst, const uint32_t* gmemSrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smemDst)), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(smemDst)),
               "l"(gmemSrc), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(parity) : "memory");
}

#if ORZ_WAIT_STATS  // (measurement builds only: make variant NAME=waitstats DEFS=-DORZ_WAIT_STATS=1, tools/wait_stats.py)
__device__ unsigned long long g_waitStats[8];  // waits entered, ended visible, ended invisible, decided on arrival, spin iterations (visible), (invisible), clocks waited (visible), (invisible)
#endif
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
// the same word as ONE value for the whole warp: every lane's load races with remote stores on its own, so control flow
// that contains full-mask collectives branches on lane 0's observation only
__device__ __forceinline__ uint32_t ld_flag_warp(const uint32_t* p) { return __shfl_sync(kFull, ld_flag(p), 0); }
// ... and as warp votes (one instruction): the words only grow, so "some lane saw it" is a valid observation for all
#if ORZ_FLAG_VOTE
__device__ __forceinline__ bool flag_set_warp(const uint32_t* p) { return __any_sync(kFull, ld_flag(p) != 0u); }
__device__ __forceinline__ bool flag_reached_warp(const uint32_t* p, uint32_t c) { return __any_sync(kFull, ld_flag(p) >= c); }
#else  // (measurement only: every lane branches on its own load)
__device__ __forceinline__ bool flag_set_warp(const uint32_t* p) { return ld_flag(p) != 0u; }
__device__ __forceinline__ bool flag_reached_warp(const uint32_t* p, uint32_t c) { return ld_flag(p) >= c; }
#endif
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
template <uint32_t TH>
__device__ __forceinline__ void step_chain(float cur, const float incX, const float incY, const uint32_t nyCommon, const uint32_t nyExtra,
                                           const uint32_t nPre, const uint32_t cA, const uint32_t cB, const bool active, float* out) {
#pragma unroll kChainUnroll
  for (uint32_t i = 0; i < nyCommon; ++i) cur = cur + incY;  // Rasterizer.cpp:1130-1131
#pragma unroll
  for (uint32_t i = 0; i + 1u < TH; ++i) cur = i < nyExtra ? cur + incY : cur;
#pragma unroll kChainUnroll
  for (uint32_t i = 0; i < nPre; ++i) cur = incX + cur;      // Rasterizer.cpp:1145-1146
  for (uint32_t c = cA; c <= cB; ++c) {
    if (active) out[c] = cur;
    cur = incX + cur;
  }
}

// two IEEE single-precision adds (round to nearest even, denormals kept) in one instruction: sm_100's packed add
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo_f32x2(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_f32x2(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// Two chains of the same (tile row, columns) per lane: independent add sequences that share the loop overhead and hide
// each other's latency (ORZ_CHAIN_MERGED: depth chain l and, in the lanes with l < 4, edge chain l).
template <uint32_t TH>
__device__ __forceinline__ void step_chain2(float a, const float aX, const float aY, float b, const float bX, const float bY, const uint32_t nyCommon,
                                            const uint32_t nyExtra, const uint32_t nPre, const uint32_t cA, const uint32_t cB, const bool actA,
                                            const bool actB, float* outA, float* outB) {
#if ORZ_CHAIN_F32X2
  uint64_t ab = pack_f32x2(a, b);
  const uint64_t incY = pack_f32x2(aY, bY), incX = pack_f32x2(aX, bX);
#pragma unroll kChainUnroll
  for (uint32_t i = 0; i < nyCommon; ++i) ab = add_f32x2(ab, incY);  // Rasterizer.cpp:1130-1131
#pragma unroll
  for (uint32_t i = 0; i + 1u < TH; ++i) ab = i < nyExtra ? add_f32x2(ab, incY) : ab;
#pragma unroll kChainUnroll
  for (uint32_t i = 0; i < nPre; ++i) ab = add_f32x2(incX, ab);      // Rasterizer.cpp:1145-1146
  for (uint32_t c = cA; c <= cB; ++c) {
    if (actA) outA[c] = lo_f32x2(ab);
    if (actB) outB[c] = hi_f32x2(ab);
    ab = add_f32x2(incX, ab);
  }
#else
#pragma unroll kChainUnroll
  for (uint32_t i = 0; i < nyCommon; ++i) { a = a + aY; b = b + bY; }  // Rasterizer.cpp:1130-1131
#pragma unroll
  for (uint32_t i = 0; i + 1u < TH; ++i) { a = i < nyExtra ? a + aY : a; b = i < nyExtra ? b + bY : b; }
#pragma unroll kChainUnroll
  for (uint32_t i = 0; i < nPre; ++i) { a = aX + a; b = bX + b; }      // Rasterizer.cpp:1145-1146
  for (uint32_t c = cA; c <= cB; ++c) {
    if (actA) outA[c] = a;
    if (actB) outB[c] = b;
    a = aX + a; b = bX + b;
  }
#endif
}

// One primitive on the tile a warp has open (Rasterizer.cpp:1098-1292 restricted to the tile's
// blocks).  d[8] / h are the lane's block and its HiZ, kept in registers between primitives.
template <uint32_t TH>
__device__ __forceinline__ void tile_prim(const uint32_t* __restrict__ rec, const int lane, const uint32_t x0, const uint32_t y0,
                                          const uint32_t x1, const uint32_t y1, const uint2* __restrict__ lut, float* __restrict__ sm,
                                          uint4* __restrict__ tile, uint32_t* __restrict__ aux, uint32_t& h, bool& dirty) {
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
#if ORZ_CHAIN_MERGED
  {  // all twelve chains in one pass: lane (row r, l) steps depth chain l and, for l < 4, edge chain l
    const uint32_t rLast = (31u - (uint32_t)__clz((int)passMask)) >> 3;
    const uint32_t cols = (passMask | (passMask >> 8) | (passMask >> 16) | (passMask >> 24)) & 0xffu;
    const uint32_t cA = (uint32_t)__ffs((int)cols) - 1u, cB = 31u - (uint32_t)__clz((int)cols);
    const uint32_t r = (uint32_t)lane >> 3, l = (uint32_t)lane & 7u, e = l & 3u;
    const bool active = r >= rFirst && r <= rLast;
    const float s = -0.5f + 1.0f / 16.0f;
    const float curD = ORZ_FMA(dzdx, s + 0.125f * (float)(l & 3u), ORZ_FMA(dzdy, (l >> 2) ? s + 0.125f : s, u2f(rec[5])));
    const float curE = u2f(rec[14 + e]), eX = u2f(rec[6 + e]), eY = u2f(rec[10 + e]);
    step_chain2<TH>(curD, dzdx, dzdy, curE, eX, eY, ya - minY, r - rFirst, x0 + cA - minX + rec[20], cA, cB, active, active && l < 4u,
                sm + (4u + l) * kChainStride + r * 8u, sm + e * kChainStride + r * 8u);
  }
#else
  {
    const uint32_t rLast = (31u - (uint32_t)__clz((int)passMask)) >> 3;
    const uint32_t cols = (passMask | (passMask >> 8) | (passMask >> 16) | (passMask >> 24)) & 0xffu;
    const uint32_t cA = (uint32_t)__ffs((int)cols) - 1u, cB = 31u - (uint32_t)__clz((int)cols);
    const uint32_t r = (uint32_t)lane >> 2, e = (uint32_t)lane & 3u;
    const bool active = lane < 16 && r >= rFirst && r <= rLast;
    float cur = 0.0f, incX = 0.0f, incY = 0.0f;
    if (active) { cur = u2f(rec[14 + e]); incX = u2f(rec[6 + e]); incY = u2f(rec[10 + e]); }
    step_chain<TH>(cur, incX, incY, ya - minY, r - rFirst, x0 + cA - minX + rec[20], cA, cB, active, sm + e * kChainStride + r * 8u);
  }
#endif
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
  const uint32_t covMask = __ballot_sync(kFull, upd);
  __syncwarp();  // orders this primitive's reads of the edge slots before the next primitive's writes (free: the warp is converged)
  if (!covMask) return;
#if !ORZ_CHAIN_MERGED
  {  // the eight depth lanes (Rasterizer.cpp:1103-1112) at the covered blocks
    const uint32_t rLast = (31u - (uint32_t)__clz((int)covMask)) >> 3, rLo = ((uint32_t)__ffs((int)covMask) - 1u) >> 3;
    const uint32_t cols = (covMask | (covMask >> 8) | (covMask >> 16) | (covMask >> 24)) & 0xffu;
    const uint32_t cA = (uint32_t)__ffs((int)cols) - 1u, cB = 31u - (uint32_t)__clz((int)cols);
    const uint32_t r = (uint32_t)lane >> 3, l = (uint32_t)lane & 7u;
    const bool active = r >= rLo && r <= rLast;
    const float s = -0.5f + 1.0f / 16.0f;
    const float cur = ORZ_FMA(dzdx, s + 0.125f * (float)(l & 3u), ORZ_FMA(dzdy, (l >> 2) ? s + 0.125f : s, u2f(rec[5])));
    step_chain<TH>(cur, dzdx, dzdy, y0 + rLo - minY, r - rLo, x0 + cA - minX + rec[20], cA, cB, active, sm + (4u + l) * kChainStride + r * 8u);
  }
  __syncwarp();
#endif
#if ORZ_BLOCK_BOUND_SKIP
  // Exact early-out per block (the reference only tests the primitive's GLOBAL maximum against the HiZ, Rasterizer.cpp:1149):
  // every pixel the update can produce is bounded by the largest of the block's four corner samples -- rows 0 and 9 at
  // pixels 0 and 7: rounding is monotone, so the samples of a row are ordered like the plane in x (same increments, same
  // number of steps in every lane) and rows 0 / 1 / 8 / 9 like the plane in y, and avg_epu16 never exceeds its larger
  // operand -- so when that bound is <= the HiZ (the smallest stored pixel) the max-merge cannot change a pixel and the
  // block is left alone.  Castle: 1.5 % of the updates, interiors (Sponza): a quarter (profiles/r1_update_stats.json).
  if (upd && h != 1u) {
    const float* smd = sm + 4u * kChainStride + lane;
    const float d0 = smd[0], d3 = smd[3u * kChainStride], d4 = smd[4u * kChainStride], d7 = smd[7u * kChainStride];
    const uint32_t ca = pack16x2(d0, ORZ_FMA(dzdx, 0.5f, d3)), cb = pack16x2(dzdy + d4, dzdy + ORZ_FMA(dzdx, 0.5f, d7));
    const uint32_t m2 = __vmaxu2(ca, cb);
    if (max(m2 & 0xffffu, m2 >> 16) <= h) upd = false;
  }
#endif
  const uint32_t updMask = __ballot_sync(kFull, upd);
  if (upd) {
    reinterpret_cast<uint2*>(aux)[lane] = mk;  // the eight lanes that will share this block read their half of it
    const uint32_t at = (uint32_t)__popc(updMask & ((1u << lane) - 1u));  // covered blocks, compacted
#if ORZ_COVERED_WORD
    reinterpret_cast<uint8_t*>(aux + 96u)[(at & 3u) * 8u + (at >> 2)] = (uint8_t)lane;  // pass-major: group g's blocks are bytes 8 g .. 8 g + 7
#else
    reinterpret_cast<uint8_t*>(aux + 96u)[at] = (uint8_t)lane;
#endif
#if ORZ_HIZ_ATOMIC
    aux[64u + lane] = 0xffffffffu;
#endif
  }
  __syncwarp();
  if (!updMask) return;
  // ---- depth rows, merge, HiZ (Rasterizer.cpp:1241-1290): the covered blocks of the tile FOUR AT A TIME, eight lanes
  // per block -- lane (rr, i) of a group builds pixels 2i, 2i+1 of rows rr, 2+rr, 4+rr, 6+rr (orz_pixel.h: the eight
  // items of a block share no arithmetic), so a primitive that covers n blocks costs ceil(n / 4) short passes at
  // (nearly) full width instead of one long pass with 32 - n idle lanes
  {
    const uint32_t it = (uint32_t)lane & 7u, g = (uint32_t)lane >> 3, rr = it >> 2, i = it & 3u;
    const float* c0 = sm + (4u + item_lane0(rr, i)) * kChainStride;
    const float* c1 = sm + (4u + item_lane1(rr, i)) * kChainStride;
    const uint32_t shift = item_mask_shift(rr, i), half = i >> 1;
    uint32_t* hNew = aux + 64u;
    const uint8_t* covered = reinterpret_cast<const uint8_t*>(aux + 96u);
    const uint32_t n = (uint32_t)__popc(updMask);
#if ORZ_ROUNDS_X2
    // two blocks per group and pass: the two updates are independent, so their shared-memory and ALU latencies overlap
    const uint2 cw2 = reinterpret_cast<const uint2*>(covered)[g];
    uint64_t mine8 = (uint64_t)cw2.x | ((uint64_t)cw2.y << 32);  // pass-major list: bytes 2 t, 2 t + 1 = this group's blocks of pass t
    for (uint32_t j = g; j < n; j += 8u) {
      const uint32_t bA = (uint32_t)mine8 & 0xffu, bB = j + 4u < n ? (uint32_t)(mine8 >> 8) & 0xffu : 32u;
      mine8 >>= 16;
      uint4* slotA = tile + bA * 8u + (it ^ (bA & 7u));
      uint4* slotB = tile + (bB & 31u) * 8u + (it ^ (bB & 7u));
      uint4 dA = *slotA, dB = *slotB;
      const uint32_t mnA = update_item(c0[bA], c1[bA], dzdx, dzdy, i >= 2u, aux[2u * bA + half] >> shift, dA.x, dA.y, dA.z, dA.w);
      *slotA = dA;
      atomicMin(&hNew[bA], min(mnA & 0xffffu, mnA >> 16));  // Rasterizer.cpp:1287-1290
      if (bB < 32u) {
        const uint32_t mnB = update_item(c0[bB], c1[bB], dzdx, dzdy, i >= 2u, aux[2u * bB + half] >> shift, dB.x, dB.y, dB.z, dB.w);
        *slotB = dB;
        atomicMin(&hNew[bB], min(mnB & 0xffffu, mnB >> 16));
      }
    }
#else
#if ORZ_HIZ_ATOMIC && ORZ_COVERED_WORD
    const uint2 cw = reinterpret_cast<const uint2*>(covered)[g];
    uint64_t mine8 = (uint64_t)cw.x | ((uint64_t)cw.y << 32);
    for (uint32_t j = g; j < n; j += 4u) {  // (no collective inside: every group makes just the passes it has blocks for)
      const uint32_t b = (uint32_t)mine8 & 0xffu;  // block (= owner lane) this group of eight lanes takes in this pass
      mine8 >>= 8;
      uint4* slot = tile + b * 8u + (it ^ (b & 7u));
      uint4 d = *slot;
      const uint32_t mn = update_item(c0[b], c1[b], dzdx, dzdy, i >= 2u, aux[2u * b + half] >> shift, d.x, d.y, d.z, d.w);
      *slot = d;
      atomicMin(&hNew[b], min(mn & 0xffffu, mn >> 16));  // Rasterizer.cpp:1287-1290
    }
#else
#if ORZ_COVERED_WORD
    const uint2 cw = reinterpret_cast<const uint2*>(covered)[g];
    uint64_t mine8 = (uint64_t)cw.x | ((uint64_t)cw.y << 32);
#endif
    for (uint32_t j = g; j < n + g; j += 4u) {  // (n + g: every group makes the same number of passes -- the shuffles below are full width)
#if ORZ_COVERED_WORD
      const uint32_t b = j < n ? (uint32_t)mine8 & 0xffu : 32u;
      mine8 >>= 8;
#else
      const uint32_t b = j < n ? (uint32_t)covered[j] : 32u;  // block (= owner lane) this group of eight lanes takes in this pass
#endif
      uint32_t mn = 0xffffffffu;
      if (b < 32u) {
        uint4* slot = tile + b * 8u + (it ^ (b & 7u));
        uint4 d = *slot;
        mn = update_item(c0[b], c1[b], dzdx, dzdy, i >= 2u, aux[2u * b + half] >> shift, d.x, d.y, d.z, d.w);
        *slot = d;
#if ORZ_HIZ_ATOMIC
        atomicMin(&hNew[b], min(mn & 0xffffu, mn >> 16));  // Rasterizer.cpp:1287-1290
#endif
      }
#if !ORZ_HIZ_ATOMIC
      mn = __vminu2(mn, __shfl_xor_sync(kFull, mn, 1));
      mn = __vminu2(mn, __shfl_xor_sync(kFull, mn, 2));
      mn = __vminu2(mn, __shfl_xor_sync(kFull, mn, 4));
      if (it == 0u && b < 32u) hNew[b] = min(mn & 0xffffu, mn >> 16);  // Rasterizer.cpp:1287-1290
#endif
    }
#endif
#endif
    __syncwarp();
    if (upd) {
      h = hNew[lane];
      dirty = true;
      // HiZ 1 is the cleared marker: the reference's next update of such a block overwrites it (Rasterizer.cpp:1271-1278),
      // which for the max-merge means its stored depth is zero from now on
      if (h == 1u) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (uint32_t t = 0; t < 8u; ++t) tile[(uint32_t)lane * 8u + t] = z;
      }
    }
  }
  __syncwarp();  // chain slots, masks and the tile are rewritten / read by the next primitive
}

// Per-warp state of the tile-major traversal: the tiles a warp owns (lane k keeps tile k) and its slices of the CTA's
// shared memory.  Shared by the cluster kernel (one view per cluster, gated) and k_raster_tiles (one ungated view over
// the whole GPU): both hand an occluder's records to rasterize().
template <uint32_t TH>  // tile height in blocks: 4 (lane <-> block, 8 x 4) or 1 (8 x 1 strips, lanes 8-31 own nothing: four times
struct TileWalkerT {    // the tiles -- finer ownership and shorter per-tile chains for the few-view and per-call paths)
  Target T;
  int lane;
  uint32_t lx, ly;          // my block inside a tile
  uint32_t tileX0, tileY0;  // lane k: origin (in blocks) of my k-th tile; 0xffff = none
  uint32_t allTiles;        // mask over k of the tiles that exist
  uint16_t* myHiz;          // + 32 k: HiZ mirror of my block in tile k
  float* myChain;
  uint32_t* myStage;
  uint32_t* myIdx;
  uint4* myTile;
  uint32_t* myAux;
  const uint2* lut;
  uint32_t* myMap;          // bit t set: tile t is mine (or NULL: no bitmap)
  uint32_t tilesX;

  // tile t belongs to warp t mod nWarps of the group that shares the view
  __device__ __forceinline__ void own_tiles(uint32_t gw, uint32_t nWarps, uint32_t K) {
    const uint32_t tilesX = (T.blocksX + kTileW - 1u) / kTileW, tilesY = (T.blocksY + TH - 1u) / TH, nTiles = tilesX * tilesY;
    tileX0 = 0xffffu; tileY0 = 0xffffu;
    if ((uint32_t)lane < K) {
      const uint32_t t = gw + (uint32_t)lane * nWarps;
      if (t < nTiles) { const uint32_t ty = t / tilesX; tileX0 = (t - ty * tilesX) * kTileW; tileY0 = ty * TH; }
    }
    allTiles = __ballot_sync(kFull, tileX0 != 0xffffu);
    this->tilesX = tilesX;
    if (myMap) {
      for (uint32_t i = (uint32_t)lane; i < (nTiles + 31u) / 32u; i += 32u) myMap[i] = 0u;
      __syncwarp();
      if (tileX0 != 0xffffu) { const uint32_t t = gw + (uint32_t)lane * nWarps; atomicOr(&myMap[t >> 5], 1u << (t & 31u)); }
      __syncwarp();
    }
  }
  // does one of my tiles lie in the tile rectangle of the block rectangle [hx0, hx1) x [hy0, hy1)?  (per lane: its own rectangle)
  __device__ __forceinline__ bool owns_tile_in(uint32_t hx0, uint32_t hx1, uint32_t hy0, uint32_t hy1) const {
    if (hx1 <= hx0 || hy1 <= hy0) return false;
    const uint32_t tx0 = hx0 >> 3, tx1 = min((hx1 - 1u) >> 3, tilesX - 1u), ty0 = hy0 / TH, ty1 = (hy1 - 1u) / TH;
    for (uint32_t ty = ty0; ty <= ty1; ++ty) {
      const uint32_t lo = ty * tilesX + tx0, hi = ty * tilesX + tx1;  // inclusive bit range of this tile row
      for (uint32_t w = lo >> 5; w <= (hi >> 5); ++w) {
        uint32_t bits = myMap[w];
        if (w == (lo >> 5)) bits &= 0xffffffffu << (lo & 31u);
        if (w == (hi >> 5)) bits &= 0xffffffffu >> (31u - (hi & 31u));
        if (bits) return true;
      }
    }
    return false;
  }
  // clear (Rasterizer.cpp:107-121): HiZ := 1 on my tiles; depth is overwritten by the first update
  __device__ __forceinline__ void clear_tiles() {
    for (uint32_t m = allTiles; m; m &= m - 1u) {
      const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
      const uint32_t bx = __shfl_sync(kFull, tileX0, (int)k) + lx, by = __shfl_sync(kFull, tileY0, (int)k) + ly;
      myHiz[32u * k] = 1;
      if (bx < T.blocksX && by < T.blocksY && ly < TH) T.hiz[by * T.blocksX + bx] = 1;
    }
    __syncwarp();
  }
  // my tiles that meet the block rectangle [bx0, bx1] x [by0, by1] (inclusive), as a mask over k
  __device__ __forceinline__ uint32_t tiles_meeting(uint32_t bx0, uint32_t bx1, uint32_t by0, uint32_t by1) const {
    return __ballot_sync(kFull, tileX0 != 0xffffu && tileX0 <= bx1 && tileX0 + kTileW > bx0 && tileY0 <= by1 && tileY0 + TH > by0);
  }
  // canonical depth for the caller: blocks that stayed cleared read as zero
  __device__ __forceinline__ void zero_cleared_tiles() {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t m = allTiles; m; m &= m - 1u) {
      const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
      const uint32_t bx = __shfl_sync(kFull, tileX0, (int)k) + lx, by = __shfl_sync(kFull, tileY0, (int)k) + ly;
      if (bx < T.blocksX && by < T.blocksY && ly < TH && myHiz[32u * k] == 1) {
        uint4* d4 = reinterpret_cast<uint4*>(T.depth) + (size_t)(by * T.blocksX + bx) * 8u;
#pragma unroll
        for (int y = 0; y < 8; ++y) d4[y] = z;
      }
    }
  }
  // which of <= 32 staged records (lane i holds the header words a, b of record i, `valid` when it has one) touch which
  // of my tiles `tm`: lane k gets the answer for tile k
  __device__ __forceinline__ uint32_t tile_hits(const uint32_t tm, const bool valid, const uint32_t a, const uint32_t b) const {
    uint32_t myHits = 0u;
    const uint32_t minX = a & 0xffffu, minY = a >> 16;
    for (uint32_t m = tm; m; m &= m - 1u) {
      const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
      const uint32_t x0 = __shfl_sync(kFull, tileX0, (int)k), y0 = __shfl_sync(kFull, tileY0, (int)k);
      const uint32_t x1 = min(x0 + kTileW, T.blocksX), y1 = min(y0 + TH, T.blocksY);
      const bool touches = valid && minX < x1 && minX + (b & 0xffffu) > x0 && minY < y1 && minY + (b >> 16) > y0;
      const uint32_t hitsK = __ballot_sync(kFull, touches);
      if ((uint32_t)lane == k) myHits = hitsK;
    }
    return myHits;
  }
  // staged records (`stage`, kRecStride words each) -> my tiles, tile-major, each tile's primitives in order; myHits from
  // tile_hits.  `gathered` false: the records are still on their way (cp.async) and are waited for after the first open.
  __device__ __forceinline__ void tile_loop(const uint32_t* __restrict__ stage, const uint32_t myHits, bool& gathered) {
    for (uint32_t m = __ballot_sync(kFull, myHits != 0u); m;) {
      const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
      m &= m - 1u;
      uint32_t hits = __shfl_sync(kFull, myHits, (int)k);
      const uint32_t x0 = __shfl_sync(kFull, tileX0, (int)k), y0 = __shfl_sync(kFull, tileY0, (int)k);
      const uint32_t x1 = min(x0 + kTileW, T.blocksX), y1 = min(y0 + TH, T.blocksY);
#if ORZ_TILE_PREFETCH
      if (m) {  // the NEXT tile's stored depth starts its way from L2 / HBM while this one is worked on
        const uint32_t k2 = (uint32_t)__ffs((int)m) - 1u;
        const uint32_t bxn = __shfl_sync(kFull, tileX0, (int)k2) + lx, byn = __shfl_sync(kFull, tileY0, (int)k2) + ly;
        if (bxn < T.blocksX && byn < T.blocksY && ly < TH && myHiz[32u * k2] != 1) prefetch_l1(reinterpret_cast<uint4*>(T.depth) + (size_t)(byn * T.blocksX + bxn) * 8u);
      }
#endif
      // open the tile: its depth goes to shared memory, one block per lane, as 8 items (rr, i) of 4 row pairs each
      // (item slot swizzled by the block so that both this lane-per-block pass and the eight-lanes-per-block
      // update passes are bank-conflict free); cleared blocks (HiZ 1) enter as zero
      const uint32_t bx = x0 + lx, by = y0 + ly;
      const bool inScreen = bx < x1 && by < y1;
      uint4* dp = reinterpret_cast<uint4*>(T.depth) + (size_t)(by * T.blocksX + bx) * 8u;
      uint32_t h = inScreen ? (uint32_t)myHiz[32u * k] : 0xffffu;  // off-screen lanes never pass
      const bool load = inScreen && h != 1u;
      uint4* mine = myTile + (uint32_t)lane * 8u;
      const uint32_t sw = (uint32_t)lane & 7u;
#pragma unroll
      for (uint32_t rr = 0; rr < 2u; ++rr) {
        uint4 R[4];
#pragma unroll
        for (uint32_t kk = 0; kk < 4u; ++kk) R[kk] = load ? dp[2u * kk + rr] : make_uint4(0u, 0u, 0u, 0u);
        mine[(rr * 4u + 0u) ^ sw] = make_uint4(R[0].x, R[1].x, R[2].x, R[3].x);
        mine[(rr * 4u + 1u) ^ sw] = make_uint4(R[0].y, R[1].y, R[2].y, R[3].y);
        mine[(rr * 4u + 2u) ^ sw] = make_uint4(R[0].z, R[1].z, R[2].z, R[3].z);
        mine[(rr * 4u + 3u) ^ sw] = make_uint4(R[0].w, R[1].w, R[2].w, R[3].w);
      }
#if ORZ_ASYNC_GATHER
      if (!gathered) { cp_async_wait_all(); gathered = true; }
#endif
      __syncwarp();
      bool dirty = false;
      for (; hits; hits &= hits - 1u)
        tile_prim<TH>(stage + ((uint32_t)__ffs((int)hits) - 1u) * kRecStride, lane, x0, y0, x1, y1, lut, myChain, myTile, myAux, h, dirty);
      if (dirty) {  // close: the blocks this occluder changed go back to HBM / L2 in the reference's row layout
#pragma unroll
        for (uint32_t rr = 0; rr < 2u; ++rr) {
          const uint4 I0 = mine[(rr * 4u + 0u) ^ sw], I1 = mine[(rr * 4u + 1u) ^ sw], I2 = mine[(rr * 4u + 2u) ^ sw], I3 = mine[(rr * 4u + 3u) ^ sw];
          dp[0u + rr] = make_uint4(I0.x, I1.x, I2.x, I3.x);
          dp[2u + rr] = make_uint4(I0.y, I1.y, I2.y, I3.y);
          dp[4u + rr] = make_uint4(I0.z, I1.z, I2.z, I3.z);
          dp[6u + rr] = make_uint4(I0.w, I1.w, I2.w, I3.w);
        }
        myHiz[32u * k] = (uint16_t)h;
        T.hiz[by * T.blocksX + bx] = (uint16_t)h;
      }
      __syncwarp();  // the next tile's open overwrites the slots other lanes may still be reading
    }
  }
  // rasterize<clipped>(occluder) restricted to my tiles: `cnt` records (k_setup_views) with their bounding-box headers;
  // tmOcc = my tiles that meet the occluder's block rectangle
  __device__ __forceinline__ void rasterize(const uint32_t* __restrict__ recs, const uint2* __restrict__ hdrs, const uint32_t cnt, const uint32_t tmOcc) {
    uint32_t nStaged = 0;
    auto flush = [&]() {
      __syncwarp();
#if ORZ_ASYNC_GATHER
      // the records start their way from L2 / HBM now and are waited for after the first tile has been opened
#pragma unroll 8
      for (uint32_t i = 0; i < nStaged; ++i)
        if (lane < kRecStride) cp_async_word(myStage + i * kRecStride + lane, recs + (size_t)myIdx[i] * kRecStride + lane);
      bool gathered = false;
      uint2 myHdr = make_uint2(0u, 0u);
      if ((uint32_t)lane < nStaged) myHdr = hdrs[myIdx[lane]];  // (L1: the scan has just read it)
#else
#pragma unroll 8
      for (uint32_t i = 0; i < nStaged; ++i)
        if (lane < kRecStride) myStage[i * kRecStride + lane] = recs[(size_t)myIdx[i] * kRecStride + lane];
      __syncwarp();
      bool gathered = true;
      uint2 myHdr = make_uint2(0u, 0u);
      if ((uint32_t)lane < nStaged) myHdr = make_uint2(myStage[lane * kRecStride], myStage[lane * kRecStride + 1]);
#endif
      const uint32_t myHits = tile_hits(tmOcc, (uint32_t)lane < nStaged, myHdr.x, myHdr.y);
      tile_loop(myStage, myHits, gathered);
#if ORZ_ASYNC_GATHER
      if (!gathered) cp_async_wait_all();  // (no tile was opened: the staging area must still be quiet before it is refilled)
#endif
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
      if (ORZ_TILE_MAP && myMap) {
        touches = owns_tile_in(hx0, hx1, hy0, hy1);
      } else {
        for (uint32_t m = tmOcc; m; m &= m - 1u) {
          const int k = __ffs((int)m) - 1;
          const uint32_t x0 = __shfl_sync(kFull, tileX0, k), y0 = __shfl_sync(kFull, tileY0, k);
          touches = touches || (hx0 < x0 + kTileW && hx1 > x0 && hy0 < y0 + TH && hy1 > y0);
        }
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
};
typedef TileWalkerT<kTileH> TileWalker;

template <int C, uint32_t TH>
#if ORZ_CLUSTER_CTAS_PER_SM
__global__ void __launch_bounds__(kClusterGW * 32, ORZ_CLUSTER_CTAS_PER_SM) k_raster_views_cluster(const FrameParams p) {
#else
__global__ void __maxnreg__(ORZ_CLUSTER_REGS) k_raster_views_cluster(const FrameParams p) {
#endif
  constexpr uint32_t GW = kClusterGW, NT = GW * 32, kWarps = C * GW;
  extern __shared__ __align__(16) uint32_t s_dyn[];
  uint2* s_lut = reinterpret_cast<uint2*>(s_dyn);
  uint32_t* s_tileAll = s_dyn + ClusterSmem::kLutWords;  // [GW] tiles (16-byte aligned), then [GW] mask / HiZ scratch
  uint32_t* s_stageAll = s_tileAll + ClusterSmem::kTileAllWords;
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
#if ORZ_LUT_BULK && ORZ_CLUSTER_LUT_SMEM
  __shared__ __align__(8) uint64_t s_lutBar;
  if (tid == 0) mbar_init(&s_lutBar, 1u);
  __syncthreads();
  if (tid == 0) bulk_load(s_lut, p.lut, 4096u * 8u, &s_lutBar);  // one bulk copy, in flight while the heads are staged and the tiles cleared
#else
  if (ORZ_CLUSTER_LUT_SMEM) for (uint32_t i = tid; i < 4096u; i += NT) s_lut[i] = p.lut[i];
#endif
  for (uint32_t i = tid; i < nOcc * kHeadWords; i += NT) s_head[i] = front[(size_t)(i / kHeadWords) * kFrontWords + i % kHeadWords];
  for (uint32_t i = tid; i < nOcc * 3u; i += NT) s_vis[i] = 0u;
  if (tid * 4u < nOcc) prefetch_l1(p.recInfo + (size_t)view * nOcc * 2u + tid * 8u);  // the walk reads two uint4 per slot: four slots per line

  Target T;
  T.width = p.width; T.height = p.height; T.blocksX = p.width >> 3; T.blocksY = p.height >> 3;
  T.depth = p.depth + (size_t)view * p.depthStride;
  T.hiz = p.hiz + (size_t)view * p.hizStride;
  const bool useGate = (p.flags & ORZ_BATCH_NO_GATE) == 0u;
  const uint4* recInfo = p.recInfo + (size_t)view * nOcc * 2u;
  uint16_t* myHiz = s_hiz + (size_t)warp * K * 32u + lane;  // + 32 k
  float* myChain = s_chain + warp * (12 * kChainStride);
  uint32_t* myStage = s_stageAll + (uint32_t)warp * kStageCap * kRecStride;
  uint32_t* myIdx = s_idxAll + (uint32_t)warp * kStageCap;
  uint4* myTile = reinterpret_cast<uint4*>(s_tileAll + (uint32_t)warp * kTileWords);
  uint32_t* myAux = s_tileAll + (uint32_t)GW * kTileWords + (uint32_t)warp * kTileAuxWords;
  const uint32_t lx = (uint32_t)lane & 7u, ly = (uint32_t)lane >> 3;

  TileWalkerT<TH> tw;
  tw.T = T; tw.lane = lane; tw.lx = lx; tw.ly = ly;
  tw.myHiz = myHiz; tw.myChain = myChain; tw.myStage = myStage; tw.myIdx = myIdx; tw.myTile = myTile; tw.myAux = myAux;
  tw.lut = ORZ_CLUSTER_LUT_SMEM ? s_lut : p.lut;
  {
    const uint32_t nTilesAll = ((T.blocksX + kTileW - 1u) / kTileW) * ((T.blocksY + TH - 1u) / TH);
    tw.myMap = ORZ_TILE_MAP ? reinterpret_cast<uint32_t*>(s_hiz + (size_t)GW * K * 32u) + (uint32_t)warp * ((nTilesAll + 31u) / 32u) : nullptr;
  }
  tw.own_tiles(gw, kWarps, K);
  tw.clear_tiles();
  const uint32_t tileX0 = tw.tileX0, tileY0 = tw.tileY0;
#if ORZ_LUT_BULK && ORZ_CLUSTER_LUT_SMEM
  mbar_wait(&s_lutBar, 0u);
#endif
  cluster.sync();  // tables staged, decision words zero in every CTA before the first remote access

  // "no visible pixel on my tiles" for candidate s: per-CTA count, forwarded by the CTA's last warp
  auto answer_no = [&](uint32_t s) {
    uint32_t old = 0u;
    if (lane == 0) old = atomicAdd(&s_doneLocal[s], 1u);
    old = __shfl_sync(kFull, old, 0);
    if (old == GW - 1u && lane < C) atomicAdd(cluster.map_shared_rank(&s_doneCta[s], (unsigned)lane), 1u);
  };
  auto tiles_meeting = [&](uint32_t bx0, uint32_t bx1, uint32_t by0, uint32_t by1) -> uint32_t { return tw.tiles_meeting(bx0, bx1, by0, by1); };

  // Which candidates this warp has answered "no" already: bit (s & 31) of lane (s >> 5) -- 1 024 order slots; scenes with more
  // occluders run without the early answers below.
  uint32_t answeredBits = 0u;
  const bool lookahead = ORZ_LOOKAHEAD > 0 && nOcc <= 1024u;
  auto is_answered = [&](uint32_t s) -> bool { return lookahead && ((__shfl_sync(kFull, answeredBits, (int)(s >> 5)) >> (s & 31u)) & 1u); };
  auto mark_answered = [&](uint32_t s) { if ((uint32_t)lane == (s >> 5)) answeredBits |= 1u << (s & 31u); };
  // EARLY "no".  Depth only grows while a view is drawn (max-merge), so a block's HiZ -- the smallest depth in it -- only
  // grows too, and query2D lets a block pass only when maxZ > HiZ (Rasterizer.cpp:310).  If NOW, for a candidate further down
  // the order, no block of its rectangle on my tiles has maxZ > HiZ (and none is still cleared: HiZ 1 reads as depth 0), then
  // the exact test this warp would make when its walk gets there -- after even more occluders -- fails at that first
  // comparison in every block: the answer can be given today.  A warp that is busy rasterising is the one the others wait
  // for at every candidate it touches; with the early answers they only wait for it where something may really be visible.
  auto look_ahead = [&](uint32_t s) {
    if (!lookahead) return;
    const uint32_t hi = min(nOcc, s + 1u + (uint32_t)ORZ_LOOKAHEAD);
    for (uint32_t s2 = s + 1u; s2 < hi; ++s2) {
      const uint32_t* hd2 = s_head + s2 * kHeadWords;
      if (hd2[0] != kBoxRect || is_answered(s2) || flag_set_warp(s_vis + s2)) continue;
      const uint32_t bx0 = hd2[1] >> 3, bx1 = hd2[2] >> 3, by0 = hd2[3] >> 3, by1 = hd2[4] >> 3, maxZ = hd2[5];
      bool maybe = false;
      for (uint32_t tm = tiles_meeting(bx0, bx1, by0, by1); tm && !maybe; tm &= tm - 1u) {
        const uint32_t k = (uint32_t)__ffs((int)tm) - 1u;
        const uint32_t bx = __shfl_sync(kFull, tileX0, (int)k) + lx, by = __shfl_sync(kFull, tileY0, (int)k) + ly;
        const uint32_t h = (uint32_t)myHiz[32u * k];
        maybe = __any_sync(kFull, ly < TH && bx >= bx0 && bx <= bx1 && by >= by0 && by <= by1 && bx < T.blocksX && by < T.blocksY && (h == 1u || maxZ > h));
      }
      if (!maybe) { answer_no(s2); mark_answered(s2); }
    }
  };

  // ---- candidates whose rectangle does not touch my tiles: answered before the walk starts
  for (uint32_t s = 0; s < nOcc; ++s) {
    const uint32_t* hd = s_head + s * kHeadWords;
    if (hd[0] != kBoxRect) continue;
    if (!tiles_meeting(hd[1] >> 3, hd[2] >> 3, hd[3] >> 3, hd[4] >> 3)) { answer_no(s); if (lookahead) mark_answered(s); }
  }

  uint4 infoNext = recInfo[0], boxNext = recInfo[1];  // {records, first slot, quads}, {block rectangle}: fetched one slot ahead of the walk
  for (uint32_t s = 0; s < nOcc; ++s) {
    const uint32_t* hd = s_head + s * kHeadWords;
    const uint32_t status = hd[0];
    const uint4 info = infoNext, box = boxNext;
    if (s + 1u < nOcc) { infoNext = recInfo[2u * s + 2u]; boxNext = recInfo[2u * s + 3u]; }
    if (status == kBoxCulled) continue;
    // my tiles that the occluder's primitives can touch; their record headers start their way to the SM now, so that
    // the gate test and the wait for the decision hide the L2 round trip (wasted on the candidates the gate rejects)
    uint32_t tmOcc = 0u;
    if (info.x != 0u && box.x < box.z) tmOcc = tiles_meeting(box.x, box.z - 1u, box.y, box.w - 1u);
    const size_t recBase = (size_t)view * p.totalQuads + info.y;
    const uint2* hdrs = p.hdrBuf + recBase;
#if ORZ_HDR_PREFETCH_EARLY
    if (tmOcc && (uint32_t)lane * 16u < info.x) prefetch_l1(hdrs + (uint32_t)lane * 16u);  // <= 504 headers = 32 lines
#endif
    bool visible = true;
    if (status != kBoxNearClip) {
      // ---- gate: query2D (Rasterizer.cpp:283-349) on the part of the rectangle that lies on my tiles
      const uint32_t minX = hd[1], maxX = hd[2], minY = hd[3], maxY = hd[4], maxZ = hd[5];
      const uint32_t bx0 = minX >> 3, bx1 = maxX >> 3, by0 = minY >> 3, by1 = maxY >> 3;
      const uint32_t* vis = s_vis + s;
      uint32_t tm = tiles_meeting(bx0, bx1, by0, by1);
      if (tm && !is_answered(s) && !flag_set_warp(vis)) {
        bool found = false;
        for (; tm; tm &= tm - 1u) {
          const uint32_t k = (uint32_t)__ffs((int)tm) - 1u;
          const uint32_t bx = __shfl_sync(kFull, tileX0, (int)k) + lx, by = __shfl_sync(kFull, tileY0, (int)k) + ly;
          if (flag_set_warp(vis)) break;  // another warp already found a visible pixel
          const bool hit = ly < TH && bx >= bx0 && bx <= bx1 && by >= by0 && by <= by1 && bx < T.blocksX && by < T.blocksY &&
                           query_block_h(T, bx, by, (uint32_t)myHiz[32u * k], minX, maxX, minY, maxY, maxZ);
          if (__any_sync(kFull, hit)) { found = true; break; }
        }
        if (found) { if (lane < C) st_flag_remote(s_vis + s, (uint32_t)lane, 1u); }
        else if (!flag_set_warp(vis)) answer_no(s);
      }
      // Only the warps that would rasterise the occluder -- its primitives touch their tiles -- depend on the decision;
      // every other warp has given its answer and walks on (the view's outputs are read from the decision words after
      // the final barrier, when all of them are final)
      if (!tmOcc) continue;
      // visible as soon as ONE warp says so, invisible when all 16 C warps have said no
      const uint32_t* done = s_doneCta + s;
      if (ORZ_LOOKAHEAD_WAITING && !flag_set_warp(vis) && !flag_reached_warp(done, (uint32_t)C)) look_ahead(s);  // (idle anyway)
#if ORZ_WAIT_STATS
      const long long t0w = clock64();
      unsigned long long spins = 0;
#endif
      for (;;) {
        if (flag_set_warp(vis)) break;
        if (flag_reached_warp(done, (uint32_t)C)) { visible = flag_set_warp(vis); break; }
#if ORZ_WAIT_STATS
        ++spins;
#endif
#if ORZ_SPIN_NAP
        __nanosleep(ORZ_SPIN_NAP);  // (a longer or growing nap was measured slower: the wake-up delay sits on the dependency chain)
#endif
      }
#if ORZ_WAIT_STATS
      if (lane == 0) {
        atomicAdd(&g_waitStats[0], 1ull);
        atomicAdd(&g_waitStats[visible ? 1 : 2], 1ull);
        if (!spins) atomicAdd(&g_waitStats[3], 1ull);
        atomicAdd(&g_waitStats[visible ? 4 : 5], spins);
        atomicAdd(&g_waitStats[visible ? 6 : 7], (unsigned long long)(clock64() - t0w));
      }
#endif
    }
    if (!visible || !tmOcc) continue;

    // ---- rasterize<clipped>(occluder): the records k_setup_views wrote, on my tiles
    const uint32_t cnt = info.x;
    const uint32_t* recs = p.recBuf + recBase * kRecStride;
#if !ORZ_HDR_PREFETCH_EARLY
    if ((uint32_t)lane * 16u < cnt) prefetch_l1(hdrs + (uint32_t)lane * 16u);  // <= 504 headers = 32 lines
#endif

    tw.rasterize(recs, hdrs, cnt, tmOcc);
    look_ahead(s);  // my tiles have just become more opaque: which of the next candidates are certainly hidden on them now?
  }
  if (p.coarseHiz) {  // the view is final on my tiles: their smallest HiZ, for the occludee queries' coarse look (orz_query.cuh)
    uint16_t* coarse = p.coarseHiz + (size_t)view * p.coarseStride;
    for (uint32_t m = tw.allTiles; m; m &= m - 1u) {
      const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
      const uint32_t bx = __shfl_sync(kFull, tileX0, (int)k) + lx, by = __shfl_sync(kFull, tileY0, (int)k) + ly;
      const uint32_t h = (bx < T.blocksX && by < T.blocksY && ly < TH) ? (uint32_t)myHiz[32u * k] : 0xffffu;
      const uint32_t lo = __reduce_min_sync(kFull, h);
      if (lane == 0) coarse[gw + k * kWarps] = (uint16_t)lo;
    }
  }
  if (p.exportDepth) tw.zero_cleared_tiles();
  cluster.sync();  // no CTA may leave while another one can still write its decision words
  // ---- the view's per-slot outputs (Main.cpp:195-204), from the now final decision words
  if (rank == 0u && warp == 0 && (p.gate || p.quadsSubmitted)) {
    uint32_t quadsSubmitted = 0;
    for (uint32_t s = (uint32_t)lane; s < nOcc; s += 32u) {
      const uint32_t status = s_head[s * kHeadWords];
      const bool visible = status == kBoxNearClip || (status == kBoxRect && s_vis[s] != 0u);
      if (p.gate) p.gate[(size_t)view * nOcc + s] = (uint8_t)((visible ? 1 : 0) | (status == kBoxNearClip && useGate ? 2 : 0));
      if (visible) quadsSubmitted += recInfo[2u * s].z;
    }
    quadsSubmitted = __reduce_add_sync(kFull, quadsSubmitted);
    if (p.quadsSubmitted && lane == 0) p.quadsSubmitted[view] = quadsSubmitted;
  }
}

// ---------------------------------------------------------------------------------------------
// One UNGATED view split over the whole GPU (BASELINE config 4: millions of quads in thousands of batches, every
// batch through rasterize<clipped>, no queryVisibility gate).  Without the gate no decision is shared between warps,
// so there is no cluster and no barrier after the prologue: tile t of the screen belongs to warp t mod (all warps of
// the grid) for the whole view, and every warp walks the occluders front to back on its own tiles with the same
// tile-major traversal as the cluster kernel (TileWalker::rasterize).  The walk over thousands of occluders is itself
// lane parallel: 32 packed block rectangles (k_setup_views: occBox) per step, a ballot picks the occluders that meet
// my tiles, and only those are opened -- in order.
struct TilesSmem {
  static constexpr uint32_t kFixedWords = ClusterSmem::kLutWords + ClusterSmem::kTileAllWords + ClusterSmem::kStageWords + ClusterSmem::kIdxWords +
                                          ClusterSmem::kChainWords;
  static size_t bytes(uint32_t tilesPerWarp) { return (size_t)kFixedWords * 4 + (size_t)kClusterGW * tilesPerWarp * 32 * 2; }
};

__global__ void __maxnreg__(ORZ_CLUSTER_REGS) k_raster_tiles(const FrameParams p, const uint32_t view, const uint32_t quadsTotal) {
  constexpr uint32_t GW = kClusterGW, NT = GW * 32;
  extern __shared__ __align__(16) uint32_t s_dyn[];
  uint2* s_lut = reinterpret_cast<uint2*>(s_dyn);
  uint32_t* s_tileAll = s_dyn + ClusterSmem::kLutWords;
  uint32_t* s_stageAll = s_tileAll + ClusterSmem::kTileAllWords;
  uint32_t* s_idxAll = s_stageAll + ClusterSmem::kStageWords;
  float* s_chain = reinterpret_cast<float*>(s_idxAll + ClusterSmem::kIdxWords);
  uint16_t* s_hiz = reinterpret_cast<uint16_t*>(s_dyn + TilesSmem::kFixedWords);  // [GW][K][32]
  const uint32_t tid = threadIdx.x;
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31u);
  const uint32_t K = p.clusterK, nOcc = p.nOcc;
  if (ORZ_CLUSTER_LUT_SMEM) for (uint32_t i = tid; i < 4096u; i += NT) s_lut[i] = p.lut[i];

  TileWalker tw;
  tw.T.width = p.width; tw.T.height = p.height; tw.T.blocksX = p.width >> 3; tw.T.blocksY = p.height >> 3;
  tw.T.depth = p.depth + (size_t)view * p.depthStride;
  tw.T.hiz = p.hiz + (size_t)view * p.hizStride;
  tw.lane = lane; tw.lx = (uint32_t)lane & 7u; tw.ly = (uint32_t)lane >> 3;
  tw.myHiz = s_hiz + (size_t)warp * K * 32u + lane;
  tw.myChain = s_chain + warp * (12 * kChainStride);
  tw.myStage = s_stageAll + (uint32_t)warp * kStageCap * kRecStride;
  tw.myIdx = s_idxAll + (uint32_t)warp * kStageCap;
  tw.myTile = reinterpret_cast<uint4*>(s_tileAll + (uint32_t)warp * kTileWords);
  tw.myAux = s_tileAll + (uint32_t)GW * kTileWords + (uint32_t)warp * kTileAuxWords;
  tw.lut = ORZ_CLUSTER_LUT_SMEM ? s_lut : p.lut;
  tw.myMap = nullptr;
  // consecutive tiles go to the warps of one CTA, then to the next CTA: a CTA's tiles are short horizontal runs all over the screen
  tw.own_tiles(blockIdx.x * GW + (uint32_t)warp, gridDim.x * GW, K);
  tw.clear_tiles();
  __syncthreads();  // table staged
  if (blockIdx.x == 0 && tid == 0 && p.quadsSubmitted) p.quadsSubmitted[view] = quadsTotal;

  const uint4* recInfo = p.recInfo + (size_t)view * nOcc * 2u;
  const uint2* occBox = p.occBox + (size_t)view * nOcc;
  for (uint32_t s0 = 0; s0 < nOcc; s0 += 32u) {
    uint32_t bx0 = 1u, by0 = 1u, bx1 = 0u, by1 = 0u;  // empty
    if (s0 + (uint32_t)lane < nOcc) {
      const uint2 b = occBox[s0 + (uint32_t)lane];
      bx0 = b.x & 0xffffu; by0 = b.x >> 16; bx1 = b.y & 0xffffu; by1 = b.y >> 16;  // half open; {0xffff, 0xffff, 0, 0}: none
    }
    bool meets = false;
    for (uint32_t m = tw.allTiles; m; m &= m - 1u) {
      const int k = __ffs((int)m) - 1;
      const uint32_t x0 = __shfl_sync(kFull, tw.tileX0, k), y0 = __shfl_sync(kFull, tw.tileY0, k);
      meets = meets || (bx0 < x0 + kTileW && bx1 > x0 && by0 < y0 + kTileH && by1 > y0);
    }
    for (uint32_t cand = __ballot_sync(kFull, meets); cand; cand &= cand - 1u) {  // in order: Main.cpp:192-206
      const uint32_t s = s0 + (uint32_t)__ffs((int)cand) - 1u;
      const uint4 info = recInfo[2u * s], box = recInfo[2u * s + 1u];
      const uint32_t tmOcc = tw.tiles_meeting(box.x, box.z - 1u, box.y, box.w - 1u);
      if (!tmOcc || info.x == 0u) continue;
      const size_t recBase = (size_t)view * p.totalQuads + info.y;
      const uint2* hdrs = p.hdrBuf + recBase;
      if ((uint32_t)lane * 16u < info.x) prefetch_l1(hdrs + (uint32_t)lane * 16u);
      tw.rasterize(p.recBuf + recBase * kRecStride, hdrs, info.x, tmOcc);
    }
  }
  if (p.exportDepth) tw.zero_cleared_tiles();
}
