// TEST INFRASTRUCTURE: runs the device batching -- the SAME kernel source (orz_sah_kernels.cuh) and the
// SAME host level loop (orz_sah_driver.inl) the CUDA library is built from -- on the CPU, one OS thread
// per CUDA thread, so their logic can be checked where no GPU exists (tests/test_scene_prep.py compares
// the result with the host batching and the reference).  Only the launch / barrier / shuffle / atomic
// primitives and the cuda* memory calls are replaced.  Slow by construction; meant for a few thousand boxes.
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../include/orz.h"

// ---- CUDA surface ---------------------------------------------------------------------------------
#define __global__
#define __device__
#define __forceinline__ inline
#define __shared__ static  // blocks of one launch run one after the other, so one static copy per kernel is a block's shared memory
#define __launch_bounds__(...)
struct float4 { float x, y, z, w; };
struct Dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim, gridDim;
static pthread_barrier_t g_blockBarrier;
static pthread_barrier_t g_warpBarrier[32];
static uint32_t g_slot[32][32];

static inline void __syncthreads() { pthread_barrier_wait(&g_blockBarrier); }
template <typename T>
static inline T emu_exchange(T v, unsigned src) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  memcpy(&g_slot[w][lane], &v, 4);
  pthread_barrier_wait(&g_warpBarrier[w]);
  T r;
  memcpy(&r, &g_slot[w][src], 4);
  pthread_barrier_wait(&g_warpBarrier[w]);
  return r;
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int d) {
  const unsigned lane = threadIdx.x & 31u;
  return emu_exchange(v, lane >= unsigned(d) ? lane - d : lane);
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int d) {
  const unsigned lane = threadIdx.x & 31u;
  return emu_exchange(v, lane + d < 32u ? lane + d : lane);
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, unsigned(src) & 31u); }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}

template <typename F>
static void emu_launch(unsigned grid, unsigned block, F&& kernel) {
  blockDim.x = block;
  gridDim.x = grid;
  pthread_barrier_init(&g_blockBarrier, nullptr, block);
  for (unsigned w = 0; w < (block + 31) / 32; ++w) pthread_barrier_init(&g_warpBarrier[w], nullptr, 32);
  std::vector<std::thread> threads;
  for (unsigned t = 0; t < block; ++t)
    threads.emplace_back([&, t] {
      threadIdx.x = t;
      for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        kernel();
        pthread_barrier_wait(&g_blockBarrier);  // the next block reuses the static shared memory
      }
    });
  for (auto& th : threads) th.join();
  pthread_barrier_destroy(&g_blockBarrier);
  for (unsigned w = 0; w < (block + 31) / 32; ++w) pthread_barrier_destroy(&g_warpBarrier[w]);
}
#define ORZ_LAUNCH(kernel, grid, block, stream, ...) emu_launch((grid), (block), [&] { kernel(__VA_ARGS__); })

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
template <typename T>
static cudaError_t cudaMalloc(T** p, size_t bytes) { *p = static_cast<T*>(malloc(bytes)); return *p ? cudaSuccess : 2; }
static cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, int, cudaStream_t) { memcpy(dst, src, bytes); return cudaSuccess; }
static cudaError_t cudaMemsetAsync(void* dst, int value, size_t bytes, cudaStream_t) { memset(dst, value, bytes); return cudaSuccess; }
static cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static cudaError_t cudaGetLastError() { return cudaSuccess; }
static cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

struct orz_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
};
static std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define ORZ_CUDA(x)                                                          \
  do {                                                                       \
    if ((x) != cudaSuccess) return fail(ORZ_ERR_CUDA, std::string(#x)); \
  } while (0)

#include "../rasterizer_b200/csrc/orz_core.h"
namespace orz {
constexpr uint32_t kFull = 0xffffffffu;
#include "../rasterizer_b200/csrc/orz_sah_kernels.cuh"
}  // namespace orz
using namespace orz;
#include "../rasterizer_b200/csrc/orz_sah_driver.inl"

extern "C" int emu_generate_batches(const float* aabbs, uint32_t n, uint32_t targetSize, uint32_t granularity, uint32_t* indicesOut,
                                    uint32_t* batchSizes, uint32_t batchCapacity, uint32_t* nBatches, uint64_t* launches) {
  orz_context ctx;
  const int rc = orz_generate_batches_device(&ctx, aabbs, n, targetSize, granularity, indicesOut, batchSizes, batchCapacity, nBatches);
  if (launches) *launches = ctx.launches;
  return rc;
}
extern "C" const char* emu_last_error() { return g_err.c_str(); }
// k_sah_scan on its own: 16 digit rows of numTiles words, scanned in place (carry across 1 024-word rounds) + row totals
extern "C" void emu_scan(uint32_t* hist, uint32_t numTiles, uint32_t* totals) { ORZ_LAUNCH(k_sah_scan, 16, 256, nullptr, hist, numTiles, totals); }
