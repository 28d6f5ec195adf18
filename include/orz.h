/* orz.h -- C ABI of the B200-native occlusion-culling rasterizer.
 *
 * Drop-in boundary for the hot path of rawrunprotected/rasterizer (SoftwareRasterizer/):
 * every entry point below replaces one public method of the reference's `Occluder` /
 * `Rasterizer` classes (file:line cited per function), or batches the per-frame loop of
 * Main.cpp:181-206 over many independent camera views.  Plain C types only: pointers, sizes,
 * opaque handles.  The C++ classes in rasterizer_b200/csrc/dropin/{Occluder.h,Rasterizer.h}
 * keep the reference's signatures on top of this ABI; INTEGRATION.md shows the binding.
 *
 * All functions return 0 on success and a non-zero code on failure (orz_last_error() gives
 * the message); there is no CPU fallback -- without a CUDA device every compute entry fails.
 */
#ifndef ORZ_H
#define ORZ_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orz_context orz_context;       /* one per (host thread, GPU): stream, tables */
typedef struct orz_occluder orz_occluder;     /* one baked batch resident in HBM (Occluder.h:7-21) */
typedef struct orz_rasterizer orz_rasterizer; /* one view's depth + HiZ + matrices (Rasterizer.h:10-61) */
typedef struct orz_scene orz_scene;           /* all baked batches + occludee boxes of a scene in HBM */

enum { ORZ_OK = 0, ORZ_ERR_CUDA = 1, ORZ_ERR_ARG = 2, ORZ_ERR_NO_DEVICE = 3, ORZ_ERR_NCCL = 4 };

const char* orz_last_error(void);
int orz_version(void);

/* ---- context ---------------------------------------------------------------------------------
 * Builds the 64x64 edge-mask table (Rasterizer.cpp:547-604, once per process) and probes the
 * host's rcpps into the table the kernels use (Rasterizer.cpp:218,223,456,717-727,981 depend on
 * the CPU's approximate reciprocal). */
int orz_context_create(int device, orz_context** out);
void orz_context_destroy(orz_context* ctx);
int orz_context_synchronize(orz_context* ctx);
void* orz_context_stream(orz_context* ctx); /* cudaStream_t all work of this context is ordered on */
int orz_context_device(orz_context* ctx);   /* CUDA device index the context was created on */
/* Replace the probed rcpps model: table[i] = bits of rcpps(1.0 + i * 2^-bits), 2^bits entries. */
int orz_context_set_rcp_table(orz_context* ctx, const uint32_t* table, int bits);
int orz_context_get_rcp_table(orz_context* ctx, uint32_t* table, int* bits); /* table: room for 2^23 max; pass NULL to get bits */
int orz_context_get_lut(orz_context* ctx, int64_t* lut4096);
/* host-only helpers (no device needed): the edge-mask table and the rcpps probe themselves.
 * orz_probe_host_rcp: table may be NULL; *bits = leading mantissa bits rcpps depends on on this
 * CPU (11 on Intel), *exact = 1 when the exponent/special-value model matched the instruction. */
int orz_edge_mask_table(int64_t* lut4096);
int orz_probe_host_rcp(uint32_t* table, int* bits, int* exact);

/* ---- Occluder::bake (Occluder.h:9, Occluder.cpp:7-181), host side -------------------------------
 * vertices: nVerts float4 (4 per quad, nVerts % 32 == 0).  packets: nVerts uint32 in the
 * reference's packet layout (group of 8 quads = 4 x 8 words, Occluder.cpp:146-156).
 * Returns m_packetCount.  rsqrtps (VectorMath.h:22) is taken from the host CPU unless a table
 * was installed with orz_set_rsqrt_table (2 << bits entries: [exponent parity][mantissa]). */
uint32_t orz_bake(const float* vertices, uint32_t nVerts, const float* refMin4, const float* refMax4,
                  uint32_t* packets, float* center4, float* boundsMin4, float* boundsMax4);
int orz_set_rsqrt_table(const uint32_t* table, int bits); /* NULL: back to the host instruction */

/* ---- scene preparation: the offline steps of Main.cpp:86-107, host side -------------------------
 * orz_quad_decompose = QuadDecomposition::decompose (QuadDecomposition.h:10, QuadDecomposition.cpp:346-445):
 * pairs the triangles of an indexed triangle list into quads by a maximum matching over the pairs whose
 * two planes stay within 0.5 units of the shared quad.  indices: nIndices words (3 per triangle);
 * vertices: nVertices float4.  quadIndices: room for 4 * (nIndices / 3) words; *nQuadIndices = words
 * written, 4 per quad (a lone triangle i0 i1 i2 becomes i0 i2 i1 i0).  Same quads, same order as the
 * reference.  Uses rsqrtps like Occluder::bake (orz_set_rsqrt_table applies). */
int orz_quad_decompose(const uint32_t* indices, size_t nIndices, const float* vertices, size_t nVertices,
                       uint32_t* quadIndices, size_t* nQuadIndices);
/* orz_generate_batches = SurfaceAreaHeuristic::generateBatches (SurfaceAreaHeuristic.h:10,
 * SurfaceAreaHeuristic.cpp:10-104): top-down SAH splits at multiples of splitGranularity until a side
 * is smaller than targetSize.  aabbs: nAabbs x (min4, max4).  indicesOut: nAabbs words = the batches'
 * members concatenated in the reference's batch order; batchSizes[b] = members of batch b (room for
 * batchCapacity entries; nAabbs / splitGranularity always suffices); *nBatches = batches found. */
int orz_generate_batches(const float* aabbs, uint32_t nAabbs, uint32_t targetSize, uint32_t splitGranularity,
                         uint32_t* indicesOut, uint32_t* batchSizes, uint32_t batchCapacity, uint32_t* nBatches);

/* The same batching on the GPU (level-synchronous: stable radix sorts by segment and centre, segmented
 * box scans, atomicMin over (cost, axis, position)); same arguments and bit-identical results. */
int orz_generate_batches_device(orz_context* ctx, const float* aabbs, uint32_t nAabbs, uint32_t targetSize,
                                uint32_t splitGranularity, uint32_t* indicesOut, uint32_t* batchSizes,
                                uint32_t batchCapacity, uint32_t* nBatches);

/* ---- single-view path: the reference's per-call API ------------------------------------------- */
/* upload one baked batch (re-laid out as one 16-byte record per quad for 128-bit coalesced loads) */
int orz_occluder_create(orz_context* ctx, const uint32_t* packets, uint32_t packetCount, const float* refMin4,
                        const float* refMax4, orz_occluder** out);
void orz_occluder_destroy(orz_occluder* occ);

int orz_rasterizer_create(orz_context* ctx, uint32_t width, uint32_t height, orz_rasterizer** out); /* Rasterizer.cpp:66-74 */
void orz_rasterizer_destroy(orz_rasterizer* r);
int orz_rasterizer_set_mvp(orz_rasterizer* r, const float* matrix16);                    /* Rasterizer.cpp:76-105 */
int orz_rasterizer_clear(orz_rasterizer* r);                                             /* Rasterizer.cpp:107-121 (+ depth := 0) */
int orz_rasterizer_rasterize(orz_rasterizer* r, const orz_occluder* occ, int possiblyNearClipped); /* Rasterizer.cpp:606-1295 */
int orz_rasterizer_query_visibility(orz_rasterizer* r, const float* boundsMin4, const float* boundsMax4,
                                    int* visible, int* needsClipping);                   /* Rasterizer.cpp:123-281 */
int orz_rasterizer_query2d(orz_rasterizer* r, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY,
                           uint32_t maxZ, int* visible);                                 /* Rasterizer.cpp:283-349 */
/* many boxes at once (host arrays): boxes = n x (min4, max4); out[i] bit0 visible, bit1 needsClipping */
int orz_rasterizer_query_boxes(orz_rasterizer* r, const float* boxes, uint32_t n, uint8_t* out);
int orz_rasterizer_readback_depth(orz_rasterizer* r, void* targetBGRA8);                 /* Rasterizer.cpp:351-399 */
/* raw buffers in the reference's layout: depth u16 [block][row][px] (cleared blocks zero), HiZ u16 [block] */
int orz_rasterizer_download(orz_rasterizer* r, uint16_t* depth, uint16_t* hiz);
/* debug / parity: setup records of every quad of `occ` (orz_prim_record each, mode 0 = culled) */
typedef struct {
  uint32_t mode;
  int32_t minX, minY, rangeX, rangeY;
  uint32_t maxZ;
  float dzdx, dzdy, plane0;
  float nx[4], ny[4], off[4];
  uint32_t slope[4];
} orz_prim_record;
int orz_rasterizer_debug_setup(orz_rasterizer* r, const orz_occluder* occ, int possiblyNearClipped,
                               orz_prim_record* out /* quadCount records, host */);

/* ---- view-batch path: Main.cpp:181-206 for many independent views -------------------------------
 * Scene: nOccluders baked batches (packets concatenated in the reference layout, packetCounts[i]
 * packets each), per-occluder refMin/refMax/boundsMin/boundsMax/center (4 floats each). */
int orz_scene_create(orz_context* ctx, const uint32_t* packets, const uint32_t* packetCounts, uint32_t nOccluders,
                     const float* refMin, const float* refMax, const float* boundsMin, const float* boundsMax,
                     const float* centers, orz_scene** out);
/* Occluder::bake (Occluder.cpp:7-181) for every batch of a scene ON THE GPU, one CTA per batch, bit-exact with
 * orz_bake; the scene is created directly in HBM.  vertices: all batches' float4 vertices concatenated (4 per
 * quad), vertCounts[i] = vertices of batch i (multiple of 32); one refMin/refMax for all batches (Main.cpp:109-128).
 * Optional host outputs (may be NULL): packetsOut = sum(vertCounts) words in the reference's packet layout,
 * centersOut / boundsMinOut / boundsMaxOut = nOccluders x 4 floats (Occluder.h:13-15). */
int orz_scene_bake(orz_context* ctx, const float* vertices, const uint32_t* vertCounts, uint32_t nOccluders,
                   const float* refMin4, const float* refMax4, uint32_t* packetsOut, float* centersOut,
                   float* boundsMinOut, float* boundsMaxOut, orz_scene** out);
int orz_scene_set_occludees(orz_scene* scene, const float* boxes, uint32_t nBoxes); /* n x (min4, max4) */
/* Main.cpp:86-128 in one call: indexed triangle mesh -> quads (orz_quad_decompose) -> padding to 8 quads -> per-quad
 * boxes -> SAH batches on the GPU (orz_generate_batches_device) -> reference box over all vertices -> Occluder::bake
 * of every batch on the GPU (orz_scene_bake).  occludeesFromQuads != 0 installs the per-quad boxes (w := 1, batch
 * order) as the occludee set (BASELINE configs 1-3).  splitGranularity: a multiple of 8. */
typedef struct {
  uint32_t nOccluders, nQuads; /* batches; quads after padding */
  float refMin[4], refMax[4];  /* Main.cpp:109-113 */
} orz_mesh_scene_info;
int orz_scene_from_mesh(orz_context* ctx, const uint32_t* indices, size_t nIndices, const float* vertices, size_t nVertices,
                        uint32_t targetSize, uint32_t splitGranularity, int occludeesFromQuads, orz_scene** out,
                        orz_mesh_scene_info* info /* may be NULL */);
uint32_t orz_scene_occludee_count(orz_scene* scene); /* boxes installed by orz_scene_set_occludees / from_mesh / load */
/* The same from the reference's raw scene files (Main.cpp:56-84: uint32 triangle indices, float4 vertices). */
int orz_scene_from_mesh_files(orz_context* ctx, const char* indexPath, const char* vertexPath, uint32_t targetSize,
                              uint32_t splitGranularity, int occludeesFromQuads, orz_scene** out, orz_mesh_scene_info* info);
/* Cached baked scene: the scene's HBM layout -- occluder meta, one 16-byte record per quad,
 * occludee boxes -- written to / read from one file; loading is three copies, no preparation and no bake. */
int orz_scene_save(orz_scene* scene, const char* path);
int orz_scene_load(orz_context* ctx, const char* path, orz_scene** out);
/* What the application reads from its occluders (Occluder.h:11-20; Main.cpp:186-195 sorts by m_center and gates with
 * the bounds): nOccluders x 4 floats each, quadCounts: nOccluders words.  Any output may be NULL. */
int orz_scene_get_occluders(orz_scene* scene, uint32_t* nOccluders, float* centers, float* boundsMin, float* boundsMax,
                            uint32_t* quadCounts);
void orz_scene_destroy(orz_scene* scene);

enum {
  ORZ_BATCH_NO_GATE = 1u,      /* submit every occluder, no queryVisibility gate (config 4 shape) */
  ORZ_BATCH_FORCE_CLIPPED = 2u, /* with NO_GATE: use rasterize<true> for every occluder */
  ORZ_BATCH_WIDE = 8u,          /* with NO_GATE: force the one-warp-per-screen-row path that splits ONE view over the
                                    whole GPU (chosen automatically for <= 8 views over >= 65536 quads) */
  ORZ_BATCH_TARGETS_ON_DEVICE = 4u /* orz_render_views only: depth / hiz are DEVICE pointers (the
                                      buffers stay resident in HBM; only bits and gates travel) */
};
typedef struct {
  uint32_t width, height;
  uint32_t nViews;
  uint32_t flags;
  const float* mvps;      /* nViews x 16 (Main.cpp:172-178) */
  const uint32_t* orders; /* nViews x nOccluders front-to-back order (Main.cpp:185-190), or NULL ... */
  const float* camPos;    /* ... then nViews x 3 camera positions: the order is computed on the GPU */
  /* outputs, any may be NULL */
  uint32_t* visBits;      /* nViews x ceil(nBoxes/32): occludee visible bits (needsClipping counts as visible) */
  uint32_t* clipBits;     /* same shape: needsClipping bits */
  uint8_t* gate;          /* nViews x nOccluders, per order slot: bit0 visible, bit1 needsClipping */
  uint16_t* depth;        /* nViews x (w*h) u16, reference block layout */
  uint16_t* hiz;          /* nViews x (w/8*h/8) u16 */
  uint32_t* quadsSubmitted; /* nViews: quads handed to rasterize (2 x packetCount per rasterised occluder) */
} orz_view_batch;

/* Host pointers; copies inputs to the GPU, renders, copies the requested outputs back, synchronises. */
int orz_render_views(orz_context* ctx, orz_scene* scene, const orz_view_batch* batch);
/* Device pointers (inputs already resident in HBM); asynchronous on the context stream. */
int orz_render_views_device(orz_context* ctx, orz_scene* scene, const orz_view_batch* batch);
/* kernels launched by the last render / rasterizer call on this context (for launch accounting) */
uint64_t orz_context_launch_count(orz_context* ctx);
/* tuning: bytes of the internal arena that holds per-view depth + HiZ when the caller does not ask
 * for them (default min(24 GB, HBM/6)); larger batches are rendered in chunks of views */
int orz_context_set_arena_bytes(orz_context* ctx, size_t bytes);
/* tuning: block traversal mapping of the view-batch kernel: 1 = one warp per 8x8 block,
 * 2 = one lane per block (up to 32 blocks of a primitive in flight per warp) */
int orz_context_set_traversal(orz_context* ctx, int mapping);
/* tuning: batches of at most maxViews views (default 16384) run one thread-block cluster (1-16 CTAs x 16 warps)
 * per view -- the latency path for single views (BASELINE configs 1 and 2) and, since round 2, the faster path for
 * every batch size measured; 0 = always use the one-CTA-per-view batch kernel */
int orz_context_set_cluster_views(orz_context* ctx, int maxViews);
/* tuning: CTAs (of 16 warps) per cluster on that path: 1, 2, 4, 8 or 16; 0 = automatic (from the target size and the batch) */
int orz_context_set_cluster_size(orz_context* ctx, int ctas);
/* tuning: height of the screen tiles a warp owns, in 8x8 blocks.  Cluster path: 4 (8 x 4 blocks, lane <-> block; what
 * 0 = automatic picks) or 1 (8 x 1 strips: four times the tiles; measured slower for the view batches, kept selectable);
 * per-call rasterize (orz_rasterizer_rasterize): 1 (default, measured faster there) or 4.  Results are identical for
 * every choice. */
int orz_context_set_tile_height(orz_context* ctx, int clusterTileHeight, int perCallTileHeight);
/* tuning: warps cooperating on one view in the batch kernel (1, 2, 4, 8, 16); 0 = automatic */
int orz_context_set_group_warps(orz_context* ctx, int warps);

/* ---- multi-GPU: batches of independent views sharded over GPUs (Main.cpp:181-206 per view) ----------
 * The reference has no multi-device code; its unit of independent work is the view.  Every rank (one GPU, one
 * context) renders its own slice of the view list with the static scene replicated -- orz_render_views_device with
 * visBits in HBM -- and ONE collective, an NCCL all-gather of the per-view visibility bitmasks over NVLink,
 * assembles the result on every rank; depth and HiZ stay on the GPU that produced them.  A single view has no useful
 * split across GPUs (ordered, shared depth buffer): replicas only.  NCCL is bound at run time (libnccl.so.2, or the
 * path in ORZ_NCCL_LIB); without it these calls fail with ORZ_ERR_NCCL and everything else keeps working.
 *   one process per GPU:  rank 0 calls orz_comm_get_unique_id, hands the 128 bytes to the other ranks by its own means
 *                         (MPI, a file, torch.distributed ...), every rank calls orz_comm_create;
 *   one process, n GPUs:  orz_comm_create_all over n contexts; bracket the per-GPU orz_gather_bits calls of one
 *                         collective with orz_comm_group_begin / orz_comm_group_end. */
typedef struct orz_comm orz_comm;
#define ORZ_COMM_ID_BYTES 128
int orz_comm_get_unique_id(void* id128);                                                       /* ncclGetUniqueId */
int orz_comm_create(orz_context* ctx, int nRanks, int rank, const void* id128, orz_comm** out); /* ncclCommInitRank */
int orz_comm_create_all(orz_context* const* ctxs, int n, orz_comm** outs);                     /* ncclCommInitAll */
void orz_comm_destroy(orz_comm* comm);
int orz_comm_rank(const orz_comm* comm);
int orz_comm_size(const orz_comm* comm);
int orz_comm_group_begin(void);
int orz_comm_group_end(void);
/* localBits: wordsPerRank words in HBM (this rank's views x ceil(nBoxes/32), rows of views the rank does not own zero);
 * allBits: nRanks x wordsPerRank words in HBM, rank r's rows at r * wordsPerRank.  Asynchronous on the context stream,
 * after the render calls enqueued before it. */
int orz_gather_bits(orz_comm* comm, const uint32_t* localBits, size_t wordsPerRank, uint32_t* allBits);
/* The same on the communicator's own stream, beside whatever the context enqueues next (the next batch): the buffers
 * must stay untouched until orz_comm_join (the context stream waits for the gathers issued so far), orz_comm_join_older
 * (... for all but the most recent one: call it before rendering into a buffer pair again when two pairs rotate) or
 * orz_comm_synchronize (the host waits for both streams). */
int orz_gather_bits_overlapped(orz_comm* comm, const uint32_t* localBits, size_t wordsPerRank, uint32_t* allBits);
int orz_comm_join(orz_comm* comm);
int orz_comm_join_older(orz_comm* comm);
int orz_comm_synchronize(orz_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* ORZ_H */
