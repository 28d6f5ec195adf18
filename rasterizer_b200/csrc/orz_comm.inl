// Multi-GPU side of the view-batch path behind the C ABI (include/orz.h "multi-GPU"): views are
// independent, every rank renders its own slice with the static scene replicated, and ONE collective -- an all-gather
// of the per-view visibility bitmasks over NCCL / NVLink -- assembles the result; depth and HiZ stay where they were
// produced.  Included at the end of orz_kernels.cu (needs orz_context).
//
// NCCL is bound at run time (dlopen), so the library loads on hosts without it and inside processes that already
// carry their own copy (PyTorch bundles one: dlopen by soname returns the copy the process has loaded).  Only the
// handful of entry points used here are declared; the types follow nccl.h (2.27/2.28): ncclUniqueId is 128 opaque
// bytes passed BY VALUE, ncclUint32 = 3.
#include <dlfcn.h>

#include <mutex>

namespace {
struct NcclId { char bytes[ORZ_COMM_ID_BYTES]; };
typedef void* NcclComm;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("ORZ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
      api.error = dlerror();
    }
    if (!api.handle) return;
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(api.handle, name);
      if (!p) { api.error = std::string("missing symbol ") + name; }
      return p;
    };
    api.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))sym("ncclCommInitRank");
    api.CommInitAll = (int (*)(NcclComm*, int, const int*))sym("ncclCommInitAll");
    api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllGather || !api.GroupStart || !api.GroupEnd ||
        !api.GetErrorString) {
      dlclose(api.handle);
      api.handle = nullptr;
    }
  });
  return api;
}
int nccl_ready() {
  NcclApi& a = nccl_api();
  if (!a.handle) return fail(ORZ_ERR_NCCL, "NCCL is not available (libnccl.so.2: " + a.error + "); set ORZ_NCCL_LIB to its path");
  return ORZ_OK;
}
}  // namespace
#define ORZ_NCCL(x)                                                                                             \
  do {                                                                                                          \
    int _r = (x);                                                                                               \
    if (_r != 0) return fail(ORZ_ERR_NCCL, std::string(#x) + ": " + nccl_api().GetErrorString(_r));             \
  } while (0)

struct orz_comm {
  orz_context* ctx;
  NcclComm comm;
  int nRanks, rank;
  cudaStream_t side = nullptr;            // overlapped gathers run here
  cudaEvent_t evReady = nullptr, evDone[2] = {nullptr, nullptr};  // completion of overlapped gather k: evDone[k & 1]
  uint64_t issued = 0, joined = 0;        // overlapped gathers issued / the context stream already waits for
};
static int comm_streams(orz_comm* c) {
  ORZ_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  ORZ_CUDA(cudaEventCreateWithFlags(&c->evReady, cudaEventDisableTiming));
  ORZ_CUDA(cudaEventCreateWithFlags(&c->evDone[0], cudaEventDisableTiming));
  ORZ_CUDA(cudaEventCreateWithFlags(&c->evDone[1], cudaEventDisableTiming));
  return ORZ_OK;
}

extern "C" int orz_comm_get_unique_id(void* id) {
  if (!id) return fail(ORZ_ERR_ARG, "orz_comm_get_unique_id: id is NULL");
  if (int e = nccl_ready()) return e;
  NcclId u;
  memset(&u, 0, sizeof u);
  ORZ_NCCL(nccl_api().GetUniqueId(&u));
  memcpy(id, u.bytes, ORZ_COMM_ID_BYTES);
  return ORZ_OK;
}
extern "C" int orz_comm_create(orz_context* ctx, int nRanks, int rank, const void* id, orz_comm** out) {
  if (!ctx || !id || !out || nRanks < 1 || rank < 0 || rank >= nRanks) return fail(ORZ_ERR_ARG, "orz_comm_create: bad arguments");
  if (int e = nccl_ready()) return e;
  ORZ_CUDA(cudaSetDevice(ctx->device));
  NcclId u;
  memcpy(u.bytes, id, ORZ_COMM_ID_BYTES);
  NcclComm c = nullptr;
  ORZ_NCCL(nccl_api().CommInitRank(&c, nRanks, u, rank));
  *out = new orz_comm{ctx, c, nRanks, rank};
  if (int e = comm_streams(*out)) { orz_comm_destroy(*out); *out = nullptr; return e; }
  return ORZ_OK;
}
extern "C" int orz_comm_create_all(orz_context* const* ctxs, int n, orz_comm** outs) {
  if (!ctxs || !outs || n < 1 || n > 64) return fail(ORZ_ERR_ARG, "orz_comm_create_all: bad arguments");
  for (int i = 0; i < n; ++i)
    if (!ctxs[i]) return fail(ORZ_ERR_ARG, "orz_comm_create_all: a context is NULL");
  if (int e = nccl_ready()) return e;
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
  std::vector<NcclComm> comms(n, nullptr);
  ORZ_NCCL(nccl_api().CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) outs[i] = new orz_comm{ctxs[i], comms[i], n, i};
  for (int i = 0; i < n; ++i) {
    ORZ_CUDA(cudaSetDevice(ctxs[i]->device));
    if (int e = comm_streams(outs[i])) return e;
  }
  return ORZ_OK;
}
extern "C" void orz_comm_destroy(orz_comm* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
  if (c->evReady) cudaEventDestroy(c->evReady);
  if (c->evDone[0]) cudaEventDestroy(c->evDone[0]);
  if (c->evDone[1]) cudaEventDestroy(c->evDone[1]);
  if (c->comm && nccl_api().handle) nccl_api().CommDestroy(c->comm);
  delete c;
}
extern "C" int orz_comm_rank(const orz_comm* c) { return c ? c->rank : -1; }
extern "C" int orz_comm_size(const orz_comm* c) { return c ? c->nRanks : 0; }
extern "C" int orz_comm_group_begin(void) {
  if (int e = nccl_ready()) return e;
  ORZ_NCCL(nccl_api().GroupStart());
  return ORZ_OK;
}
extern "C" int orz_comm_group_end(void) {
  if (int e = nccl_ready()) return e;
  ORZ_NCCL(nccl_api().GroupEnd());
  return ORZ_OK;
}
// all-gather of `wordsPerRank` words per rank on the context's stream: ordered after the render calls that produced
// localBits and before whatever the caller enqueues next; rank r's words land at allBits + r * wordsPerRank everywhere
extern "C" int orz_gather_bits(orz_comm* c, const uint32_t* localBits, size_t wordsPerRank, uint32_t* allBits) {
  if (!c || !allBits || (wordsPerRank && !localBits)) return fail(ORZ_ERR_ARG, "orz_gather_bits: bad arguments");
  if (wordsPerRank == 0) return ORZ_OK;
  ORZ_CUDA(cudaSetDevice(c->ctx->device));
  ORZ_NCCL(nccl_api().AllGather(localBits, allBits, wordsPerRank, 3 /* ncclUint32 */, c->comm, c->ctx->stream));
  return ORZ_OK;
}
// The same gather on the communicator's own stream: it starts when everything enqueued on the context stream so far
// has finished and runs BESIDE what the caller enqueues next (the next batch's kernels), so a short step does not
// wait for the slowest rank to arrive at the collective.  The caller must not overwrite localBits or read allBits
// before orz_comm_join (context stream waits) or orz_comm_synchronize (host waits): double-buffer them.
extern "C" int orz_gather_bits_overlapped(orz_comm* c, const uint32_t* localBits, size_t wordsPerRank, uint32_t* allBits) {
  if (!c || !allBits || (wordsPerRank && !localBits)) return fail(ORZ_ERR_ARG, "orz_gather_bits_overlapped: bad arguments");
  if (wordsPerRank == 0) return ORZ_OK;
  ORZ_CUDA(cudaSetDevice(c->ctx->device));
  ORZ_CUDA(cudaEventRecord(c->evReady, c->ctx->stream));
  ORZ_CUDA(cudaStreamWaitEvent(c->side, c->evReady, 0));
  ORZ_NCCL(nccl_api().AllGather(localBits, allBits, wordsPerRank, 3 /* ncclUint32 */, c->comm, c->side));
  ORZ_CUDA(cudaEventRecord(c->evDone[c->issued & 1u], c->side));
  c->issued++;
  return ORZ_OK;
}
// the context stream waits for the overlapped gathers issued so far -- all of them, or (keepNewest) all but the most
// recent one: with two buffer pairs in rotation, "join older" before rendering into a pair again is all the ordering needed
static int comm_join(orz_comm* c, uint64_t upTo) {
  if (upTo <= c->joined) return ORZ_OK;
  ORZ_CUDA(cudaSetDevice(c->ctx->device));
  ORZ_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->evDone[(upTo - 1u) & 1u], 0));  // the side stream is in order: gather upTo-1 done => all earlier done
  c->joined = upTo;
  return ORZ_OK;
}
extern "C" int orz_comm_join(orz_comm* c) {
  if (!c) return fail(ORZ_ERR_ARG, "orz_comm_join: communicator is NULL");
  return comm_join(c, c->issued);
}
extern "C" int orz_comm_join_older(orz_comm* c) {
  if (!c) return fail(ORZ_ERR_ARG, "orz_comm_join_older: communicator is NULL");
  return c->issued ? comm_join(c, c->issued - 1u) : ORZ_OK;
}
extern "C" int orz_comm_synchronize(orz_comm* c) {
  if (!c) return fail(ORZ_ERR_ARG, "orz_comm_synchronize: communicator is NULL");
  ORZ_CUDA(cudaSetDevice(c->ctx->device));
  ORZ_CUDA(cudaStreamSynchronize(c->side));
  ORZ_CUDA(cudaStreamSynchronize(c->ctx->stream));
  c->joined = c->issued;
  return ORZ_OK;
}
