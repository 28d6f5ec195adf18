// Application-side test program written against the reference's C++ API only: the scene preparation of
// Main.cpp:86-128 (QuadDecomposition::decompose, Aabb, SurfaceAreaHeuristic::generateBatches,
// Occluder::bake) over a raw mesh.  Compiled by tests/test_scene_prep.py against
// rasterizer_b200/csrc/dropin; the same source compiles against the reference headers.
//   usage: dropin_prepare indices.bin vertices.bin out.bin
//   out: u32 nBatches, then per batch u32 nQuads + center/boundsMin/boundsMax (12 floats) + packed words
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <memory>
#include <vector>

#include "Occluder.h"
#include "QuadDecomposition.h"
#include "SurfaceAreaHeuristic.h"
#include "VectorMath.h"

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
  if (argc != 4) return 2;
  auto indices = readAll<uint32_t>(argv[1]);
  auto raw = readAll<float>(argv[2]);
  std::vector<__m128> vertices(raw.size() / 4);
  for (size_t i = 0; i < vertices.size(); ++i) vertices[i] = _mm_loadu_ps(raw.data() + 4 * i);

  indices = QuadDecomposition::decompose(indices, vertices);
  while (indices.size() % 32 != 0) indices.push_back(indices[0]);

  std::vector<Aabb> quadAabbs;
  for (size_t quad = 0; quad < indices.size() / 4; ++quad) {
    Aabb aabb;
    for (int k = 0; k < 4; ++k) aabb.include(vertices[indices[4 * quad + k]]);
    quadAabbs.push_back(aabb);
  }
  auto batches = SurfaceAreaHeuristic::generateBatches(quadAabbs, 512, 8);

  Aabb refAabb;
  for (auto v : vertices) refAabb.include(v);

  std::ofstream out(argv[3], std::ofstream::binary);
  uint32_t nBatches = uint32_t(batches.size());
  out.write(reinterpret_cast<const char*>(&nBatches), 4);
  for (const auto& batch : batches) {
    std::vector<__m128> batchVertices;
    for (auto quad : batch)
      for (int k = 0; k < 4; ++k) batchVertices.push_back(vertices[indices[4 * quad + k]]);
    auto occluder = Occluder::bake(batchVertices, refAabb.m_min, refAabb.m_max);
    uint32_t nQuads = uint32_t(batch.size());
    out.write(reinterpret_cast<const char*>(&nQuads), 4);
    out.write(reinterpret_cast<const char*>(&occluder->m_center), 16);
    out.write(reinterpret_cast<const char*>(&occluder->m_boundsMin), 16);
    out.write(reinterpret_cast<const char*>(&occluder->m_boundsMax), 16);
    out.write(reinterpret_cast<const char*>(occluder->m_vertexData), size_t(occluder->m_packetCount) * 32);
  }
  return 0;
}
