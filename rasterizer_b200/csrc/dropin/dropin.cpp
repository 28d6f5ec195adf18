// Thin C++ wrappers: reference signatures (Occluder.h:9, Rasterizer.h:13-26, QuadDecomposition.h:10,
// SurfaceAreaHeuristic.h:10) over the C ABI.
// The reference has no error channel (asserts only, Rasterizer.cpp:68, Occluder.cpp:9); a failing
// ABI call prints orz_last_error() and aborts, which keeps the void signatures.
#include "Occluder.h"
#include "QuadDecomposition.h"
#include "Rasterizer.h"
#include "SurfaceAreaHeuristic.h"
#include "VectorMath.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../../include/orz.h"

namespace {
void check(int code, const char* what) {
  if (code != ORZ_OK) {
    std::fprintf(stderr, "rasterizer_b200: %s failed: %s\n", what, orz_last_error());
    std::abort();
  }
}
}  // namespace

// One context per host thread (a context is one stream + scratch and is not re-entrant), shared by every drop-in object
// the thread creates and REFERENCE COUNTED by them: the reference application keeps its Rasterizer and Occluders in
// static-duration globals (Main.cpp:41-44), which outlive the main thread's thread_locals, and objects may be handed to
// or destroyed on another thread -- the context dies with its last user, not with the thread that made it.
namespace {
struct ContextHolder {
  orz_context* ctx = nullptr;
  ~ContextHolder() { if (ctx) orz_context_destroy(ctx); }
};
std::shared_ptr<void> threadContext() {
  static thread_local std::shared_ptr<ContextHolder> tl;
  if (!tl) {
    auto h = std::make_shared<ContextHolder>();
    const char* dev = std::getenv("ORZ_DEVICE");
    check(orz_context_create(dev ? std::atoi(dev) : 0, &h->ctx), "orz_context_create");
    tl = std::move(h);
  }
  return tl;
}
orz_context* raw(const std::shared_ptr<void>& holder) { return static_cast<ContextHolder*>(holder.get())->ctx; }
std::mutex g_uploadMutex;  // lazy device copies of occluders shared between threads (const Occluder& is shareable in the reference)
}  // namespace

orz_context* Rasterizer::context() { return raw(threadContext()); }

std::unique_ptr<Occluder> Occluder::bake(const std::vector<__m128>& vertices, __m128 refMin, __m128 refMax) {
  auto occ = std::make_unique<Occluder>();
  const uint32_t nVerts = uint32_t(vertices.size());
  float mn[4], mx[4], c[4], bmin[4], bmax[4];
  _mm_storeu_ps(mn, refMin);
  _mm_storeu_ps(mx, refMax);
  void* mem = nullptr;
  if (posix_memalign(&mem, 32, size_t(nVerts ? nVerts : 8) * 4) != 0) std::abort();
  occ->m_vertexData = static_cast<__m256i*>(mem);
  occ->m_packetCount = orz_bake(reinterpret_cast<const float*>(vertices.data()), nVerts, mn, mx,
                                reinterpret_cast<uint32_t*>(occ->m_vertexData), c, bmin, bmax);
  if (nVerts && occ->m_packetCount * 8 != nVerts) check(ORZ_ERR_ARG, "orz_bake (vertex count must be a multiple of 32)");
  occ->m_refMin = refMin;
  occ->m_refMax = refMax;
  occ->m_center = _mm_loadu_ps(c);
  occ->m_boundsMin = _mm_loadu_ps(bmin);
  occ->m_boundsMax = _mm_loadu_ps(bmax);
  return occ;
}

Occluder::~Occluder() {
  if (m_device) orz_occluder_destroy(m_device);
  std::free(m_vertexData);
}

Rasterizer::Rasterizer(uint32_t width, uint32_t height) : m_impl(nullptr), m_width(width), m_height(height), m_context(threadContext()) {
  check(orz_rasterizer_create(raw(m_context), width, height, &m_impl), "orz_rasterizer_create");
}
Rasterizer::~Rasterizer() { orz_rasterizer_destroy(m_impl); }  // m_context is released afterwards (member order)

void Rasterizer::setModelViewProjection(const float* matrix) { check(orz_rasterizer_set_mvp(m_impl, matrix), "orz_rasterizer_set_mvp"); }
void Rasterizer::clear() { check(orz_rasterizer_clear(m_impl), "orz_rasterizer_clear"); }

template <bool possiblyNearClipped>
void Rasterizer::rasterize(const Occluder& occluder) {
  {
    std::lock_guard<std::mutex> lock(g_uploadMutex);
    if (!occluder.m_device) {  // the copy lives in (and keeps alive) the context of the rasterizer that first used it
      float mn[4], mx[4];
      _mm_storeu_ps(mn, occluder.m_refMin);
      _mm_storeu_ps(mx, occluder.m_refMax);
      occluder.m_context = m_context;
      check(orz_occluder_create(raw(m_context), reinterpret_cast<const uint32_t*>(occluder.m_vertexData), occluder.m_packetCount, mn, mx,
                                &occluder.m_device), "orz_occluder_create");
    }
  }
  check(orz_rasterizer_rasterize(m_impl, occluder.m_device, possiblyNearClipped ? 1 : 0), "orz_rasterizer_rasterize");
}
template void Rasterizer::rasterize<true>(const Occluder& occluder);
template void Rasterizer::rasterize<false>(const Occluder& occluder);

bool Rasterizer::queryVisibility(__m128 boundsMin, __m128 boundsMax, bool& needsClipping) {
  float mn[4], mx[4];
  _mm_storeu_ps(mn, boundsMin);
  _mm_storeu_ps(mx, boundsMax);
  int vis = 0, clip = needsClipping ? 1 : 0;  // left untouched when frustum culled, as in Rasterizer.cpp:164-167
  check(orz_rasterizer_query_visibility(m_impl, mn, mx, &vis, &clip), "orz_rasterizer_query_visibility");
  needsClipping = clip != 0;
  return vis != 0;
}

bool Rasterizer::query2D(uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ) const {
  int vis = 0;
  check(orz_rasterizer_query2d(m_impl, minX, maxX, minY, maxY, maxZ, &vis), "orz_rasterizer_query2d");
  return vis != 0;
}

void Rasterizer::readBackDepth(void* target) const { check(orz_rasterizer_readback_depth(m_impl, target), "orz_rasterizer_readback_depth"); }

void Rasterizer::queryVisibilityBatch(const float* boxesMinMax, uint32_t count, uint8_t* out) {
  check(orz_rasterizer_query_boxes(m_impl, boxesMinMax, count, out), "orz_rasterizer_query_boxes");
}
void Rasterizer::download(uint16_t* depth, uint16_t* hiZ) const { check(orz_rasterizer_download(m_impl, depth, hiZ), "orz_rasterizer_download"); }

// ---- scene preparation (Main.cpp:86-107) ----------------------------------------------------------
std::vector<uint32_t> QuadDecomposition::decompose(const std::vector<uint32_t>& indices, const std::vector<__m128>& vertices) {
  std::vector<uint32_t> quads(4 * (indices.size() / 3));
  size_t words = 0;
  check(orz_quad_decompose(indices.data(), indices.size(), reinterpret_cast<const float*>(vertices.data()), vertices.size(), quads.data(), &words),
        "orz_quad_decompose");
  quads.resize(words);
  return quads;
}

std::vector<std::vector<uint32_t>> SurfaceAreaHeuristic::generateBatches(const std::vector<Aabb>& aabbs, uint32_t targetSize,
                                                                         uint32_t splitGranularity) {
  const uint32_t n = uint32_t(aabbs.size()), capacity = n / (splitGranularity ? splitGranularity : 1) + 2;
  std::vector<uint32_t> order(n), sizes(capacity);
  uint32_t nBatches = 0;
  const float* boxes = reinterpret_cast<const float*>(aabbs.data());
  const char* onHost = std::getenv("ORZ_PREP_ON_HOST");
  if (onHost && onHost[0] == '1')
    check(orz_generate_batches(boxes, n, targetSize, splitGranularity, order.data(), sizes.data(), capacity, &nBatches), "orz_generate_batches");
  else
    check(orz_generate_batches_device(Rasterizer::context(), boxes, n, targetSize, splitGranularity, order.data(), sizes.data(), capacity, &nBatches),
          "orz_generate_batches_device");
  std::vector<std::vector<uint32_t>> batches(nBatches);
  size_t at = 0;
  for (uint32_t b = 0; b < nBatches; ++b) {
    batches[b].assign(order.begin() + at, order.begin() + at + sizes[b]);
    at += sizes[b];
  }
  return batches;
}
