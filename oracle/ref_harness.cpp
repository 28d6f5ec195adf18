// TEST INFRASTRUCTURE ONLY -- headless C harness around the UNMODIFIED reference sources.
//
// The four portable reference translation units (Rasterizer.cpp, Occluder.cpp,
// QuadDecomposition.cpp, SurfaceAreaHeuristic.cpp) are compiled where they lie under
// /root/reference/SoftwareRasterizer by oracle/Makefile; nothing from them is copied
// here.  Main.cpp (Win32 + DirectXMath, Main.cpp:10-11) cannot be built, so this file
// restates only what Main.cpp does around the hot path:
//   * scene load / pad / per-quad AABBs / SAH batches / refAabb / bake  (Main.cpp:56-128)
//   * the frame loop clear -> setMVP -> {queryVisibility -> rasterize<clip>}*  (Main.cpp:181-206)
// and exposes the reference objects through a flat C API for ctypes (tests, smoke(),
// bench.py's cpu_baseline / --impl reference legs).  The product never links this.
//
// Fresh-state semantics: Rasterizer::clear() does not zero depth
// (Rasterizer.cpp:107-121) while query2D reads depth of cleared blocks
// (Rasterizer.cpp:310-343); the harness zeroes m_depthBuffer when asked so that a
// view does not depend on its predecessor.
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <new>
#include <thread>
#include <vector>

#define private public
#include "Rasterizer.h"
#undef private
#include "Occluder.h"
#include "QuadDecomposition.h"
#include "SurfaceAreaHeuristic.h"
#include "VectorMath.h"

// The reference stores 32-byte vectors into a std::vector<__m128i> (Rasterizer.cpp:1274-1284);
// glibc only guarantees 16 bytes, so every allocation of this library is 64-byte aligned.
void* operator new(std::size_t n) {
  void* p = nullptr;
  if (posix_memalign(&p, 64, n ? n : 1) != 0) throw std::bad_alloc();
  return p;
}
void* operator new[](std::size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { free(p); }
void operator delete[](void* p) noexcept { free(p); }
void operator delete(void* p, std::size_t) noexcept { free(p); }
void operator delete[](void* p, std::size_t) noexcept { free(p); }

namespace {

struct RefScene {
  std::vector<__m128> vertices;
  std::vector<uint32_t> quadIndices;            // 4 per quad, padded to a multiple of 8 quads
  std::vector<std::vector<__m128>> batchVertices;  // input of Occluder::bake per batch
  std::vector<std::unique_ptr<Occluder>> occluders;
  std::vector<float> quadBoxes;                 // per quad: min xyz,1  max xyz,1
  __m128 refMin, refMax;
};

template <typename T>
bool readFile(const char* path, std::vector<T>& out) {
  std::ifstream in(path, std::ifstream::binary);
  if (!in) return false;
  in.seekg(0, std::ifstream::end);
  size_t size = size_t(in.tellg());
  in.seekg(0);
  out.resize(size / sizeof(T));
  in.read(reinterpret_cast<char*>(out.data()), out.size() * sizeof(T));
  return true;
}

}  // namespace

extern "C" {

// ---- scene preparation: Main.cpp:56-128 -----------------------------------------------------
void* ref_scene_load(const char* indexPath, const char* vertexPath) {
  auto s = std::make_unique<RefScene>();
  std::vector<uint32_t> indices;
  if (!readFile(indexPath, indices) || !readFile(vertexPath, s->vertices)) return nullptr;

  indices = QuadDecomposition::decompose(indices, s->vertices);  // Main.cpp:86
  while (indices.size() % 32 != 0) indices.push_back(indices[0]);  // Main.cpp:91-94

  std::vector<Aabb> quadAabbs;  // Main.cpp:96-105
  for (size_t q = 0; q < indices.size() / 4; ++q) {
    Aabb aabb;
    for (int k = 0; k < 4; ++k) aabb.include(s->vertices[indices[4 * q + k]]);
    quadAabbs.push_back(aabb);
  }
  auto batches = SurfaceAreaHeuristic::generateBatches(quadAabbs, 512, 8);  // Main.cpp:107

  Aabb refAabb;  // Main.cpp:109-113
  for (auto v : s->vertices) refAabb.include(v);
  s->refMin = refAabb.m_min;
  s->refMax = refAabb.m_max;

  for (const auto& batch : batches) {  // Main.cpp:116-128
    std::vector<__m128> bv;
    for (auto q : batch)
      for (int k = 0; k < 4; ++k) bv.push_back(s->vertices[indices[q * 4 + k]]);
    s->occluders.push_back(Occluder::bake(bv, refAabb.m_min, refAabb.m_max));
    s->batchVertices.push_back(std::move(bv));
  }
  // occludee boxes: per-quad AABBs in batch order with w := 1 (as Occluder.cpp:172-173 does)
  for (const auto& bv : s->batchVertices) {
    for (size_t q = 0; q < bv.size() / 4; ++q) {
      Aabb aabb;
      for (int k = 0; k < 4; ++k) aabb.include(bv[4 * q + k]);
      float mn[4], mx[4];
      _mm_storeu_ps(mn, aabb.m_min);
      _mm_storeu_ps(mx, aabb.m_max);
      mn[3] = mx[3] = 1.0f;
      s->quadBoxes.insert(s->quadBoxes.end(), mn, mn + 4);
      s->quadBoxes.insert(s->quadBoxes.end(), mx, mx + 4);
    }
  }
  s->quadIndices = std::move(indices);
  return s.release();
}

// ---- the two preprocessing steps on their own (tests/test_scene_prep.py) -----------------------
// QuadDecomposition::decompose (QuadDecomposition.cpp:346); out: room for 4 * (nIndices / 3) words
size_t ref_quad_decompose(const uint32_t* indices, size_t nIndices, const float* verts, size_t nVerts, uint32_t* out) {
  std::vector<uint32_t> idx(indices, indices + nIndices);
  std::vector<__m128> v(nVerts);
  for (size_t i = 0; i < nVerts; ++i) v[i] = _mm_loadu_ps(verts + 4 * i);
  auto quads = QuadDecomposition::decompose(idx, v);
  memcpy(out, quads.data(), quads.size() * 4);
  return quads.size();
}
// SurfaceAreaHeuristic::generateBatches (SurfaceAreaHeuristic.cpp:96); aabbs: n x (min4, max4);
// indicesOut: n words, batchSizes: room for n entries; returns the number of batches
uint32_t ref_generate_batches(const float* aabbs, uint32_t n, uint32_t targetSize, uint32_t granularity,
                              uint32_t* indicesOut, uint32_t* batchSizes) {
  std::vector<Aabb> boxes(n);
  for (uint32_t i = 0; i < n; ++i) {
    boxes[i].m_min = _mm_loadu_ps(aabbs + 8 * size_t(i));
    boxes[i].m_max = _mm_loadu_ps(aabbs + 8 * size_t(i) + 4);
  }
  auto batches = SurfaceAreaHeuristic::generateBatches(boxes, targetSize, granularity);
  uint32_t w = 0;
  for (size_t b = 0; b < batches.size(); ++b) {
    batchSizes[b] = uint32_t(batches[b].size());
    for (auto q : batches[b]) indicesOut[w++] = q;
  }
  return uint32_t(batches.size());
}

// Scene from caller-made batches (synthetic scenes): bake every batch with the reference bake.
void* ref_scene_from_batches(const float* verts, const uint32_t* batchQuads, uint32_t nBatches,
                             const float* refMin4, const float* refMax4) {
  auto s = std::make_unique<RefScene>();
  s->refMin = _mm_loadu_ps(refMin4);
  s->refMax = _mm_loadu_ps(refMax4);
  const float* p = verts;
  for (uint32_t b = 0; b < nBatches; ++b) {
    std::vector<__m128> bv;
    for (uint32_t i = 0; i < batchQuads[b] * 4; ++i, p += 4) bv.push_back(_mm_loadu_ps(p));
    s->occluders.push_back(Occluder::bake(bv, s->refMin, s->refMax));
    s->batchVertices.push_back(std::move(bv));
  }
  return s.release();
}

void ref_scene_free(void* h) {
  auto* s = static_cast<RefScene*>(h);
  if (!s) return;
  for (auto& o : s->occluders) free(o->m_vertexData);  // the reference leaks this (no destructor)
  delete s;
}

uint32_t ref_scene_num_occluders(void* h) { return uint32_t(static_cast<RefScene*>(h)->occluders.size()); }
uint32_t ref_scene_num_boxes(void* h) { return uint32_t(static_cast<RefScene*>(h)->quadBoxes.size() / 8); }
const float* ref_scene_boxes(void* h) { return static_cast<RefScene*>(h)->quadBoxes.data(); }
void ref_scene_ref_aabb(void* h, float* mn, float* mx) {
  auto* s = static_cast<RefScene*>(h);
  _mm_storeu_ps(mn, s->refMin);
  _mm_storeu_ps(mx, s->refMax);
}
uint32_t ref_scene_batch_quads(void* h, uint32_t i) {
  return uint32_t(static_cast<RefScene*>(h)->batchVertices[i].size() / 4);
}
void ref_scene_batch_vertices(void* h, uint32_t i, float* out) {
  auto& bv = static_cast<RefScene*>(h)->batchVertices[i];
  memcpy(out, bv.data(), bv.size() * 16);
}
// center, boundsMin, boundsMax (4 floats each), packetCount  -- Occluder.h:11-20
void ref_scene_occluder_meta(void* h, uint32_t i, float* center, float* bmin, float* bmax, uint32_t* packets) {
  auto& o = *static_cast<RefScene*>(h)->occluders[i];
  _mm_storeu_ps(center, o.m_center);
  _mm_storeu_ps(bmin, o.m_boundsMin);
  _mm_storeu_ps(bmax, o.m_boundsMax);
  *packets = o.m_packetCount;
}
const void* ref_scene_occluder_packets(void* h, uint32_t i) {
  return static_cast<RefScene*>(h)->occluders[i]->m_vertexData;
}

// ---- rasterizer -----------------------------------------------------------------------------
void* ref_rast_create(uint32_t w, uint32_t h) { return new Rasterizer(w, h); }
void ref_rast_free(void* r) { delete static_cast<Rasterizer*>(r); }
void ref_rast_set_mvp(void* r, const float* m) { static_cast<Rasterizer*>(r)->setModelViewProjection(m); }
void ref_rast_clear(void* r, int zeroDepth) {
  auto* R = static_cast<Rasterizer*>(r);
  R->clear();
  if (zeroDepth) memset(R->m_depthBuffer.data(), 0, R->m_depthBuffer.size() * sizeof(__m128i));
}
void ref_rast_rasterize(void* r, void* scene, uint32_t occ, int clipped) {
  auto* R = static_cast<Rasterizer*>(r);
  auto& o = *static_cast<RefScene*>(scene)->occluders[occ];
  if (clipped) R->rasterize<true>(o); else R->rasterize<false>(o);
}
// bit0 = return value, bit1 = needsClipping (left 0 when the call returns before setting it)
int ref_rast_query(void* r, const float* bmin, const float* bmax) {
  bool clip = false;
  bool vis = static_cast<Rasterizer*>(r)->queryVisibility(_mm_loadu_ps(bmin), _mm_loadu_ps(bmax), clip);
  return (vis ? 1 : 0) | (clip ? 2 : 0);
}
int ref_rast_query2d(void* r, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ) {
  return static_cast<Rasterizer*>(r)->query2D(minX, maxX, minY, maxY, maxZ) ? 1 : 0;
}
void ref_rast_query_boxes(void* r, const float* boxes, uint32_t n, uint8_t* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = uint8_t(ref_rast_query(r, boxes + 8 * i, boxes + 8 * i + 4));
}
void ref_rast_readback(void* r, void* target) { static_cast<Rasterizer*>(r)->readBackDepth(target); }
void ref_rast_get_hiz(void* r, uint16_t* out) {
  auto* R = static_cast<Rasterizer*>(r);
  memcpy(out, R->m_hiZ.data(), size_t(R->m_blocksX) * R->m_blocksY * 2);
}
// native block layout [block][row 0..7][px 0..7] u16; canonical => cleared blocks (HiZ==1) read as 0
void ref_rast_get_depth(void* r, uint16_t* out, int canonical) {
  auto* R = static_cast<Rasterizer*>(r);
  size_t blocks = size_t(R->m_blocksX) * R->m_blocksY;
  memcpy(out, R->m_depthBuffer.data(), blocks * 128);
  if (canonical)
    for (size_t b = 0; b < blocks; ++b)
      if (R->m_hiZ[b] == 1) memset(out + 64 * b, 0, 128);
}
void ref_rast_get_lut(void* r, int64_t* out) {
  auto* R = static_cast<Rasterizer*>(r);
  memcpy(out, R->m_precomputedRasterTables.data(), R->m_precomputedRasterTables.size() * 8);
}
void ref_rast_get_matrices(void* r, float* baked16, float* raw16) {
  auto* R = static_cast<Rasterizer*>(r);
  memcpy(baked16, R->m_modelViewProjection, 64);
  memcpy(raw16, R->m_modelViewProjectionRaw, 64);
}

// One frame with Main.cpp:181-206 semantics; `order` replaces the std::sort of Main.cpp:185-190
// (computed once by the caller and shared with the implementation under test).
// gate[i] for order[i]: bit0 visible, bit1 needsClipping.  Returns quads handed to rasterize.
uint64_t ref_rast_frame(void* r, void* scene, const float* mvp, const uint32_t* order, uint32_t nOrder,
                        int zeroDepth, uint8_t* gate) {
  auto* R = static_cast<Rasterizer*>(r);
  auto* S = static_cast<RefScene*>(scene);
  ref_rast_clear(r, zeroDepth);
  R->setModelViewProjection(mvp);
  uint64_t quads = 0;
  for (uint32_t i = 0; i < nOrder; ++i) {
    const Occluder& o = *S->occluders[order[i]];
    bool clip = false;
    bool vis = R->queryVisibility(o.m_boundsMin, o.m_boundsMax, clip);
    if (gate) gate[i] = uint8_t((vis ? 1 : 0) | (clip ? 2 : 0));
    if (vis) {
      if (clip) R->rasterize<true>(o); else R->rasterize<false>(o);
      quads += 2 * uint64_t(o.m_packetCount);
    }
  }
  return quads;
}

// Submit every occluder in `order` through rasterize<clipped> with no gate (config 4 shape).
void ref_rast_submit_all(void* r, void* scene, const float* mvp, const uint32_t* order, uint32_t nOrder,
                         int clipped, int zeroDepth) {
  auto* R = static_cast<Rasterizer*>(r);
  auto* S = static_cast<RefScene*>(scene);
  ref_rast_clear(r, zeroDepth);
  R->setModelViewProjection(mvp);
  for (uint32_t i = 0; i < nOrder; ++i) {
    const Occluder& o = *S->occluders[order[i]];
    if (clipped) R->rasterize<true>(o); else R->rasterize<false>(o);
  }
}

// 1 when this library's own (64-byte aligned) operator new is the one its code gets -- see -Bsymbolic in the Makefile
int ref_selftest_alignment() {
  bool ok = true;
  for (int i = 0; i < 8; ++i) {
    std::vector<__m128i>* v = new std::vector<__m128i>(size_t(1000 + 37 * i));
    ok = ok && (reinterpret_cast<uintptr_t>(v->data()) % 64 == 0);
    delete v;
  }
  return ok ? 1 : 0;
}

// ---- host instruction probes (rcpps / rsqrtps define the results) -----------
void ref_rcp_ps(const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) _mm_store_ss(out + i, _mm_rcp_ss(_mm_load_ss(in + i)));
}
void ref_rsqrt_ps(const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) _mm_store_ss(out + i, _mm_rsqrt_ss(_mm_load_ss(in + i)));
}

// ---- CPU baseline timing: one Rasterizer per thread, views interleaved over threads ----------
// Timed window per view = Main.cpp:180-208 (clear .. last rasterize) and, separately, the
// occludee-query loop.  The stock path is timed: depth is NOT zeroed between views (that is the
// reference's own behaviour and the cheaper one).  Returns wall seconds for all views; sums per phase in out[0..1]
// (thread-seconds), quads submitted in out[2], visible boxes in out[3].
double ref_bench_views(void* scene, uint32_t w, uint32_t h, const float* mvps, const uint32_t* orders,
                       uint32_t nViews, uint32_t nOrder, const float* boxes, uint32_t nBoxes,
                       uint32_t nThreads, uint32_t reps, double* out) {
  auto* S = static_cast<RefScene*>(scene);
  if (nThreads == 0) nThreads = 1;
  std::vector<std::unique_ptr<Rasterizer>> rast;
  for (uint32_t t = 0; t < nThreads; ++t) rast.push_back(std::make_unique<Rasterizer>(w, h));
  std::vector<double> frameSec(nThreads, 0.0), querySec(nThreads, 0.0);
  std::vector<uint64_t> quads(nThreads, 0), visible(nThreads, 0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](uint32_t t) {
    Rasterizer* R = rast[t].get();
    for (uint32_t rep = 0; rep < reps; ++rep)
      for (uint32_t v = t; v < nViews; v += nThreads) {
        auto a = std::chrono::steady_clock::now();
        quads[t] += ref_rast_frame(R, S, mvps + 16 * size_t(v), orders + size_t(nOrder) * v, nOrder, 0, nullptr);
        auto b = std::chrono::steady_clock::now();
        uint64_t vis = 0;
        for (uint32_t i = 0; i < nBoxes; ++i) vis += ref_rast_query(R, boxes + 8 * size_t(i), boxes + 8 * size_t(i) + 4) & 1;
        auto c = std::chrono::steady_clock::now();
        visible[t] += vis;
        frameSec[t] += std::chrono::duration<double>(b - a).count();
        querySec[t] += std::chrono::duration<double>(c - b).count();
      }
  };
  if (nThreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nThreads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  out[0] = out[1] = out[2] = out[3] = 0;
  for (uint32_t t = 0; t < nThreads; ++t) {
    out[0] += frameSec[t];
    out[1] += querySec[t];
    out[2] += double(quads[t]);
    out[3] += double(visible[t]);
  }
  return wall;
}

// ---- parity at benchmark sizes: render every view with the reference (fresh state, all host threads)
// and compare with what the implementation under test produced.  Any `got*` pointer may be NULL.
// mode bit0: no gate (every occluder submitted), bit1: with bit0, rasterize<true>.
// mismatch[v] bits: 1 gate, 2 HiZ, 4 depth (canonical: cleared blocks read as zero), 8 visible bits,
// 16 needsClipping bits, 32 quads submitted.  Returns the number of views with any mismatch.
uint32_t ref_check_views(void* scene, uint32_t w, uint32_t h, const float* mvps, const uint32_t* orders, uint32_t nViews,
                         uint32_t nOrder, const float* boxes, uint32_t nBoxes, uint32_t nThreads, uint32_t mode,
                         const uint8_t* gotGate, const uint16_t* gotDepth, const uint16_t* gotHiz, const uint32_t* gotVis,
                         const uint32_t* gotClip, const uint32_t* gotQuads, uint32_t* mismatch) {
  auto* S = static_cast<RefScene*>(scene);
  if (nThreads == 0) nThreads = 1;
  nThreads = std::min<uint32_t>(nThreads, std::max<uint32_t>(nViews, 1u));
  const size_t blocks = size_t(w / 8) * (h / 8), words = (size_t(nBoxes) + 31) / 32;
  std::atomic<uint32_t> next{0}, bad{0};
  auto work = [&]() {
    Rasterizer R(w, h);
    std::vector<uint8_t> gate(nOrder);
    std::vector<uint16_t> depth(blocks * 64);
    std::vector<uint32_t> vis(words), clip(words);
    for (;;) {
      const uint32_t v = next.fetch_add(1);
      if (v >= nViews) break;
      const float* mvp = mvps + 16 * size_t(v);
      const uint32_t* order = orders + size_t(nOrder) * v;
      uint64_t quads = 0;
      if (mode & 1u) {
        ref_rast_submit_all(&R, S, mvp, order, nOrder, (mode & 2u) ? 1 : 0, 1);
        for (uint32_t i = 0; i < nOrder; ++i) quads += 2 * uint64_t(S->occluders[order[i]]->m_packetCount);
        std::fill(gate.begin(), gate.end(), uint8_t(1));
      } else {
        quads = ref_rast_frame(&R, S, mvp, order, nOrder, 1, gate.data());
      }
      uint32_t m = 0;
      if (gotGate && !(mode & 1u) && memcmp(gotGate + size_t(nOrder) * v, gate.data(), nOrder) != 0) m |= 1u;
      if (gotHiz && memcmp(gotHiz + blocks * v, R.m_hiZ.data(), blocks * 2) != 0) m |= 2u;
      if (gotDepth) {
        ref_rast_get_depth(&R, depth.data(), 1);
        if (memcmp(gotDepth + blocks * 64 * v, depth.data(), blocks * 128) != 0) m |= 4u;
      }
      if ((gotVis || gotClip) && nBoxes) {
        std::fill(vis.begin(), vis.end(), 0u);
        std::fill(clip.begin(), clip.end(), 0u);
        for (uint32_t i = 0; i < nBoxes; ++i) {
          const int q = ref_rast_query(&R, boxes + 8 * size_t(i), boxes + 8 * size_t(i) + 4);
          if (q & 1) vis[i >> 5] |= 1u << (i & 31);
          if (q & 2) clip[i >> 5] |= 1u << (i & 31);
        }
        if (gotVis && memcmp(gotVis + words * v, vis.data(), words * 4) != 0) m |= 8u;
        if (gotClip && memcmp(gotClip + words * v, clip.data(), words * 4) != 0) m |= 16u;
      }
      if (gotQuads && gotQuads[v] != uint32_t(quads)) m |= 32u;
      if (mismatch) mismatch[v] = m;
      if (m) bad.fetch_add(1);
    }
  };
  std::vector<std::thread> th;
  for (uint32_t t = 1; t < nThreads; ++t) th.emplace_back(work);
  work();
  for (auto& x : th) x.join();
  return bad.load();
}

// ---- persistent bench pool: one Rasterizer per thread, built once (each constructor builds the 520 ms
// edge-mask table, Rasterizer.cpp:547-604) and in parallel; ref_pool_bench times the stock path like
// ref_bench_views on the pool's warm threads' rasterizers.
struct RefPool {
  uint32_t w, h;
  std::vector<std::unique_ptr<Rasterizer>> rast;
};
void* ref_pool_create(uint32_t w, uint32_t h, uint32_t nThreads) {
  auto* P = new RefPool{w, h, {}};
  if (nThreads == 0) nThreads = 1;
  P->rast.resize(nThreads);
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < nThreads; ++t) th.emplace_back([P, t, w, h]() { P->rast[t] = std::make_unique<Rasterizer>(w, h); });
  for (auto& x : th) x.join();
  return P;
}
void ref_pool_free(void* p) { delete static_cast<RefPool*>(p); }
double ref_pool_bench(void* pool, void* scene, const float* mvps, const uint32_t* orders, uint32_t nViews, uint32_t nOrder,
                      const float* boxes, uint32_t nBoxes, uint32_t reps, double* out) {
  auto* P = static_cast<RefPool*>(pool);
  auto* S = static_cast<RefScene*>(scene);
  const uint32_t nThreads = uint32_t(P->rast.size());
  std::vector<double> frameSec(nThreads, 0.0), querySec(nThreads, 0.0);
  std::vector<uint64_t> quads(nThreads, 0), visible(nThreads, 0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](uint32_t t) {
    Rasterizer* R = P->rast[t].get();
    for (uint32_t rep = 0; rep < reps; ++rep)
      for (uint32_t v = t; v < nViews; v += nThreads) {
        auto a = std::chrono::steady_clock::now();
        quads[t] += ref_rast_frame(R, S, mvps + 16 * size_t(v), orders + size_t(nOrder) * v, nOrder, 0, nullptr);
        auto b = std::chrono::steady_clock::now();
        uint64_t vis = 0;
        for (uint32_t i = 0; i < nBoxes; ++i) vis += ref_rast_query(R, boxes + 8 * size_t(i), boxes + 8 * size_t(i) + 4) & 1;
        auto c = std::chrono::steady_clock::now();
        visible[t] += vis;
        frameSec[t] += std::chrono::duration<double>(b - a).count();
        querySec[t] += std::chrono::duration<double>(c - b).count();
      }
  };
  std::vector<std::thread> th;
  for (uint32_t t = 1; t < nThreads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  out[0] = out[1] = out[2] = out[3] = 0;
  for (uint32_t t = 0; t < nThreads; ++t) { out[0] += frameSec[t]; out[1] += querySec[t]; out[2] += double(quads[t]); out[3] += double(visible[t]); }
  return wall;
}

}  // extern "C"
