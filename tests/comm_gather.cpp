// A caller of the reference's own kind -- one C++ process -- sharding a batch of independent views (Main.cpp:181-206
// per view) over every visible GPU through the C ABI alone (include/orz.h): one context + one copy of the baked scene
// per GPU, views dealt round-robin, orz_render_views_device per GPU, ONE orz_gather_bits collective (NCCL all-gather,
// ncclCommInitAll flavour) and every GPU ends up with every view's visibility bitmask.
//   usage: comm_gather baked.orzbake width height mvps.bin campos.bin out.bin [maxGpus]
// out.bin: nViews x words u32, de-interleaved back into view order, taken from the LAST GPU's copy of the gathered
// buffer; exit code 4 when the GPUs' copies differ.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>

#include "orz.h"

#define CHECK(x)                                                        \
  do {                                                                  \
    if ((x) != 0) {                                                     \
      std::fprintf(stderr, "%s failed: %s\n", #x, orz_last_error());    \
      return 3;                                                         \
    }                                                                   \
  } while (0)
#define CUDA(x)                                                                         \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) {                                                            \
      std::fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));              \
      return 3;                                                                         \
    }                                                                                   \
  } while (0)

template <typename T>
static std::vector<T> readAll(const char* path) {
  std::ifstream in(path, std::ifstream::binary);
  if (!in) return {};
  in.seekg(0, std::ifstream::end);
  size_t n = size_t(in.tellg());
  in.seekg(0);
  std::vector<T> v(n / sizeof(T));
  in.read(reinterpret_cast<char*>(v.data()), v.size() * sizeof(T));
  return v;
}

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  const uint32_t width = uint32_t(atoi(argv[2])), height = uint32_t(atoi(argv[3]));
  auto mvps = readAll<float>(argv[4]);
  auto pos = readAll<float>(argv[5]);
  const uint32_t nViews = uint32_t(mvps.size() / 16);
  int n = 0;
  CUDA(cudaGetDeviceCount(&n));
  if (argc > 7 && atoi(argv[7]) > 0 && atoi(argv[7]) < n) n = atoi(argv[7]);
  if (n < 1) return 3;
  const uint32_t per = (nViews + n - 1) / n;  // rows every rank owns in the gathered buffer

  std::vector<orz_context*> ctx(n);
  std::vector<orz_scene*> scene(n);
  std::vector<orz_comm*> comm(n);
  for (int i = 0; i < n; ++i) {
    CHECK(orz_context_create(i, &ctx[i]));
    CHECK(orz_scene_load(ctx[i], argv[1], &scene[i]));
  }
  CHECK(orz_comm_create_all(ctx.data(), n, comm.data()));
  const size_t words = (orz_scene_occludee_count(scene[0]) + 31) / 32;

  std::vector<float*> dMvp(n), dPos(n);
  std::vector<uint32_t*> dLocal(n), dAll(n);
  std::vector<uint32_t> mine(n, 0);
  for (int i = 0; i < n; ++i) {
    std::vector<float> m, p;
    for (uint32_t v = uint32_t(i); v < nViews; v += uint32_t(n)) {  // round-robin dealing: camera paths are coherent
      m.insert(m.end(), mvps.begin() + 16 * size_t(v), mvps.begin() + 16 * size_t(v) + 16);
      p.insert(p.end(), pos.begin() + 3 * size_t(v), pos.begin() + 3 * size_t(v) + 3);
    }
    mine[i] = uint32_t(m.size() / 16);
    CUDA(cudaSetDevice(i));
    CUDA(cudaMalloc(&dMvp[i], per * 64));
    CUDA(cudaMalloc(&dPos[i], per * 12));
    CUDA(cudaMalloc(&dLocal[i], per * words * 4));
    CUDA(cudaMalloc(&dAll[i], size_t(n) * per * words * 4));
    CUDA(cudaMemset(dLocal[i], 0, per * words * 4));  // rows of views this rank does not own stay zero
    CUDA(cudaMemcpy(dMvp[i], m.data(), m.size() * 4, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(dPos[i], p.data(), p.size() * 4, cudaMemcpyHostToDevice));
  }
  for (int rep = 0; rep < 2; ++rep) {  // second round: the overlapped flavour
    for (int i = 0; i < n; ++i) {      // every GPU renders its slice, asynchronously on its context's stream
      orz_view_batch b;
      memset(&b, 0, sizeof b);
      b.width = width; b.height = height; b.nViews = mine[i];
      b.mvps = dMvp[i]; b.camPos = dPos[i]; b.visBits = dLocal[i];
      CHECK(orz_render_views_device(ctx[i], scene[i], &b));
    }
    CHECK(orz_comm_group_begin());
    for (int i = 0; i < n; ++i) {
      if (rep == 0) CHECK(orz_gather_bits(comm[i], dLocal[i], per * words, dAll[i]));
      else CHECK(orz_gather_bits_overlapped(comm[i], dLocal[i], per * words, dAll[i]));
    }
    CHECK(orz_comm_group_end());
    for (int i = 0; i < n; ++i) CHECK(orz_comm_synchronize(comm[i]));
  }
  std::vector<uint32_t> first, all(size_t(n) * per * words);
  for (int i = 0; i < n; ++i) {
    CUDA(cudaSetDevice(i));
    CUDA(cudaMemcpy(all.data(), dAll[i], all.size() * 4, cudaMemcpyDeviceToHost));
    if (i == 0) first = all;
    else if (all != first) return 4;
  }
  std::vector<uint32_t> out(size_t(nViews) * words);
  for (uint32_t v = 0; v < nViews; ++v)  // view v sits in row v / n of rank v % n
    memcpy(out.data() + size_t(v) * words, all.data() + (size_t(v % n) * per + v / n) * words, words * 4);
  std::ofstream f(argv[6], std::ofstream::binary);
  f.write(reinterpret_cast<const char*>(out.data()), out.size() * 4);
  for (int i = 0; i < n; ++i) {
    CUDA(cudaSetDevice(i));
    cudaFree(dMvp[i]); cudaFree(dPos[i]); cudaFree(dLocal[i]); cudaFree(dAll[i]);
    orz_comm_destroy(comm[i]);
    orz_scene_destroy(scene[i]);
    orz_context_destroy(ctx[i]);
  }
  std::printf("ok %d gpus %u views %zu words\n", n, nViews, words);
  return 0;
}
