// Application-side test program written against the reference's C++ API only (Occluder::bake,
// Rasterizer::clear / setModelViewProjection / queryVisibility / rasterize<bool> / readBackDepth):
// the frame loop of Main.cpp:181-206 over a prepared scene file.  Compiled by tests/test_dropin_cpp.py
// against rasterizer_b200/csrc/dropin; the same source would compile against the reference headers.
//   usage: dropin_frame scene.orzscn width height mvp.bin order.bin out.bin
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <vector>

#include "Occluder.h"
#include "Rasterizer.h"

template <typename T>
static std::vector<T> readAll(const char* path) {
  std::ifstream in(path, std::ifstream::binary);
  in.seekg(0, std::ifstream::end);
  size_t n = size_t(in.tellg());
  in.seekg(0);
  std::vector<T> v(n / sizeof(T));
  in.read(reinterpret_cast<char*>(v.data()), v.size() * sizeof(T));
  return v;
}

int main(int argc, char** argv) {
  if (argc != 7) return 2;
  auto file = readAll<char>(argv[1]);
  const uint32_t width = uint32_t(atoi(argv[2])), height = uint32_t(atoi(argv[3]));
  auto mvp = readAll<float>(argv[4]);
  auto order = readAll<uint32_t>(argv[5]);
  if (memcmp(file.data(), "ORZSCN1", 7) != 0) return 3;
  uint32_t nBatches, nQuads;
  memcpy(&nBatches, file.data() + 8, 4);
  memcpy(&nQuads, file.data() + 12, 4);
  __m128 refMin = _mm_loadu_ps(reinterpret_cast<const float*>(file.data() + 16));
  __m128 refMax = _mm_loadu_ps(reinterpret_cast<const float*>(file.data() + 32));
  const uint32_t* counts = reinterpret_cast<const uint32_t*>(file.data() + 48);
  const float* verts = reinterpret_cast<const float*>(file.data() + 48 + 4 * size_t(nBatches));

  std::vector<std::unique_ptr<Occluder>> occluders;  // Main.cpp:116-128
  for (uint32_t b = 0; b < nBatches; ++b) {
    std::vector<__m128> batch;
    for (uint32_t i = 0; i < counts[b] * 4; ++i, verts += 4) batch.push_back(_mm_loadu_ps(verts));
    occluders.push_back(Occluder::bake(batch, refMin, refMax));
  }

  Rasterizer rasterizer(width, height);  // Main.cpp:88
  std::vector<uint8_t> gate(order.size());
  rasterizer.clear();                    // Main.cpp:181-206
  rasterizer.setModelViewProjection(mvp.data());
  for (size_t i = 0; i < order.size(); ++i) {
    const auto& occluder = occluders[order[i]];
    bool needsClipping = false;
    bool visible = rasterizer.queryVisibility(occluder->m_boundsMin, occluder->m_boundsMax, needsClipping);
    gate[i] = uint8_t((visible ? 1 : 0) | (needsClipping ? 2 : 0));
    if (visible) {
      if (needsClipping) rasterizer.rasterize<true>(*occluder);
      else rasterizer.rasterize<false>(*occluder);
    }
  }
  std::vector<uint8_t> image(size_t(width) * height * 4);
  rasterizer.readBackDepth(image.data());  // Main.cpp:229

  std::vector<uint16_t> depth(size_t(width) * height), hiz(size_t(width / 8) * (height / 8));
  rasterizer.download(depth.data(), hiz.data());
  std::ofstream out(argv[6], std::ofstream::binary);
  out.write(reinterpret_cast<const char*>(gate.data()), gate.size());
  out.write(reinterpret_cast<const char*>(hiz.data()), hiz.size() * 2);
  out.write(reinterpret_cast<const char*>(depth.data()), depth.size() * 2);
  out.write(reinterpret_cast<const char*>(image.data()), image.size());
  std::printf("ok %u batches %u quads\n", nBatches, nQuads);
  // optional timing of the same frame loop (ORZ_FRAME_REPS=n): what an unchanged application pays per frame
  if (const char* reps = std::getenv("ORZ_FRAME_REPS")) {
    const int n = atoi(reps);
    unsigned visibleCount = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int rep = 0; rep < n; ++rep) {
      rasterizer.clear();
      rasterizer.setModelViewProjection(mvp.data());
      for (size_t i = 0; i < order.size(); ++i) {
        const auto& occluder = occluders[order[i]];
        bool needsClipping = false;
        if (rasterizer.queryVisibility(occluder->m_boundsMin, occluder->m_boundsMax, needsClipping)) {
          ++visibleCount;
          if (needsClipping) rasterizer.rasterize<true>(*occluder);
          else rasterizer.rasterize<false>(*occluder);
        }
      }
    }
    bool nc = false;  // the last query also waits for the last rasterize
    rasterizer.queryVisibility(occluders[0]->m_boundsMin, occluders[0]->m_boundsMax, nc);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (n > 0 ? n : 1);
    std::printf("frame_ms %.4f visible_per_frame %u\n", ms, n > 0 ? visibleCount / unsigned(n) : 0u);
  }
  return 0;
}
