// Large-batch path: k_prepare_views, k_sort_views, k_render_views (one CTA per view at a time).
// Part of the single translation unit orz_kernels.cu (included inside namespace orz); see DESIGN.md section 4.
#pragma once

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

__global__ void __launch_bounds__(512) k_prepare_views(const FrameParams p) {  // 128 threads per view for batches, 512 for a few views (latency)
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
    __shared__ float s_key[1024];  // every occluder's key once (this kernel orders at most 1 024 occluders: larger scenes take k_order_keys / k_order_rank)
    for (uint32_t i = tid; i < p.nOcc; i += NT) {
      const float* ci = p.occ[i].center;
      const float dxi = ci[0] - cx, dyi = ci[1] - cy, dzi = ci[2] - cz;
      s_key[i] = (dxi * dxi + dyi * dyi) + dzi * dzi;
    }
    __syncthreads();
    for (uint32_t i = tid; i < p.nOcc; i += NT) {
      const float ki = s_key[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < p.nOcc; ++j) {
        const float kj = s_key[j];
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
    out[20] = om.quadOffset; out[21] = om.quadCount;  // (k_setup_views: one read instead of order -> occluder meta)
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

// Front-to-back order for scenes with many occluders (config 4: ~10 000 batches): the rank sort of k_prepare_views is
// quadratic inside ONE CTA per view, so it is spread over the GPU instead -- keys once, then one thread per
// (view, occluder) counts the keys in front of it from shared-memory tiles.  Same key arithmetic, same tie rule.
__global__ void __launch_bounds__(256) k_order_keys(const FrameParams p, float* __restrict__ keys) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)p.nViews * p.nOcc) return;
  const uint32_t view = (uint32_t)(i / p.nOcc), o = (uint32_t)(i - (size_t)view * p.nOcc);
  const float cx = p.camPos[3 * (size_t)view + 0], cy = p.camPos[3 * (size_t)view + 1], cz = p.camPos[3 * (size_t)view + 2];
  const float* c = p.occ[o].center;
  const float dx = c[0] - cx, dy = c[1] - cy, dz = c[2] - cz;
  keys[i] = (dx * dx + dy * dy) + dz * dz;
}
__global__ void __launch_bounds__(256) k_order_rank(const FrameParams p, const float* __restrict__ keys, uint32_t chunksPerView) {
  __shared__ float s_keys[1024];
  const uint32_t view = blockIdx.x / chunksPerView, chunk = blockIdx.x - view * chunksPerView, tid = threadIdx.x;
  const float* vk = keys + (size_t)view * p.nOcc;
  const uint32_t i = chunk * 256u + tid;
  const float ki = i < p.nOcc ? vk[i] : 0.0f;
  uint32_t rank = 0;
  for (uint32_t j0 = 0; j0 < p.nOcc; j0 += 1024u) {
    const uint32_t n = min(1024u, p.nOcc - j0);
    __syncthreads();
    for (uint32_t j = tid; j < n; j += 256u) s_keys[j] = vk[j0 + j];
    __syncthreads();
    for (uint32_t j = 0; j < n; ++j) {
      const float kj = s_keys[j];
      rank += (kj < ki || (kj == ki && j0 + j < i)) ? 1u : 0u;
    }
  }
  if (i < p.nOcc) p.orderBuf[(size_t)view * p.nOcc + rank] = i;
}

// views by descending cost, ties by index: rank sort, one WARP per view (lane l counts the costs l, l + 32, ... that come
// first; the costs pass through shared memory 1 024 at a time)
__global__ void __launch_bounds__(256) k_sort_views(const uint32_t* __restrict__ cost, uint32_t n, uint32_t* __restrict__ order) {
  __shared__ uint32_t s_cost[1024];
  const uint32_t lane = threadIdx.x & 31u, i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t ci = i < n ? cost[i] : 0u;
  uint32_t rank = 0;
  for (uint32_t j0 = 0; j0 < n; j0 += 1024u) {
    const uint32_t m = min(1024u, n - j0);
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) s_cost[j] = cost[j0 + j];
    __syncthreads();
    for (uint32_t j = lane; j < m; j += 32u) {
      const uint32_t cj = s_cost[j];
      rank += (cj > ci || (cj == ci && j0 + j < i)) ? 1u : 0u;
    }
  }
  rank = __reduce_add_sync(kFull, rank);
  if (i < n && lane == 0u) order[rank] = i;
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
