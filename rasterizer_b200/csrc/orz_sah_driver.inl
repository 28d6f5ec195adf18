// Level loop of the device batching (kernels: orz_sah_kernels.cuh).  Included by orz_kernels.cu; kept in
// its own file so that tests/sah_emulation.cpp can compile the SAME host logic and the SAME kernel
// source against a thread-per-lane CPU emulation of the launch / barrier / shuffle primitives and check
// them here, where no GPU exists (ORZ_LAUNCH and the cuda* calls are the only things it replaces).
namespace {
struct SahBuffers {
  float4* boxes = nullptr;
  uint32_t *order = nullptr, *key[2] = {nullptr, nullptr}, *idx[2] = {nullptr, nullptr}, *seg[2] = {nullptr, nullptr};
  uint32_t *segStatic = nullptr, *hist = nullptr, *digitTotals = nullptr, *segChunk0 = nullptr;
  float *areaLeft = nullptr, *areaRight = nullptr;
  SahSeg* segs = nullptr;
  SahChunk* chunks = nullptr;
  SahBox *chunkBox = nullptr, *before = nullptr, *after = nullptr;
  unsigned long long* best = nullptr;
  int cur = 0;  // which half of key / idx / seg holds the current order
  char* arena = nullptr;  // everything above lives in this one allocation
  ~SahBuffers() { cudaFree(arena); }
};
// stable sort of the working array by (segment, key): 8 key digits, then as many segment digits as the level needs
void sah_sort(orz_context* ctx, SahBuffers& B, uint32_t M, uint32_t nSegs) {
  const uint32_t tiles = (M + kSahTile - 1) / kSahTile;
  int segBits = 0;
  while (nSegs > 1 && (1u << segBits) < nSegs) ++segBits;
  for (int pass = 0; pass < 8 + (segBits + 3) / 4; ++pass) {
    const int bySegment = pass >= 8, shift = 4 * (bySegment ? pass - 8 : pass), c = B.cur;
    ORZ_LAUNCH(k_sah_hist, tiles, kSahThreads, ctx->stream, bySegment ? B.seg[c] : B.key[c], M, shift, tiles, B.hist);
    ORZ_LAUNCH(k_sah_scan, 16, 256, ctx->stream, B.hist, tiles, B.digitTotals);
    ORZ_LAUNCH(k_sah_scatter, tiles, kSahThreads, ctx->stream, B.key[c], B.idx[c], B.seg[c], B.key[c ^ 1], B.idx[c ^ 1], B.seg[c ^ 1], M, shift,
                                                          bySegment, tiles, B.hist, B.digitTotals);
    ctx->launches += 3;
    B.cur ^= 1;
  }
}
}  // namespace

extern "C" int orz_generate_batches_device(orz_context* ctx, const float* aabbs, uint32_t n, uint32_t targetSize, uint32_t granularity,
                                           uint32_t* indicesOut, uint32_t* batchSizes, uint32_t batchCapacity, uint32_t* nBatches) {
  if (!ctx || !aabbs || !indicesOut || !batchSizes || !nBatches || granularity == 0 || n >= (1u << 30))
    return fail(ORZ_ERR_ARG, "orz_generate_batches_device: bad arguments");
  ORZ_CUDA(cudaSetDevice(ctx->device));
  ctx->launches = 0;
  struct Range { uint32_t start, n; };
  std::vector<Range> active{{0u, n}}, leaves;
  const uint32_t maxSegs = n / std::max(1u, std::min(granularity, targetSize)) + 2;  // every node holds >= granularity elements
  const uint32_t maxChunks = n / kSahChunk + maxSegs + 1, maxTiles = n / kSahTile + 1;
  SahBuffers B;
  {  // one device allocation for everything (cudaMalloc / cudaFree cost more than the kernels of a small scene)
    const size_t elems = std::max<size_t>(n, 1);
    size_t total = 0;
    auto reserve = [&](size_t bytes) { const size_t at = total; total += (bytes + 255) & ~size_t(255); return at; };
    const size_t oBoxes = reserve(elems * 32), oOrder = reserve(elems * 4), oKey0 = reserve(elems * 4), oKey1 = reserve(elems * 4),
                 oIdx0 = reserve(elems * 4), oIdx1 = reserve(elems * 4), oSeg0 = reserve(elems * 4), oSeg1 = reserve(elems * 4),
                 oSegStatic = reserve(elems * 4), oAreaLeft = reserve(elems * 4), oAreaRight = reserve(elems * 4),
                 oHist = reserve((size_t)16 * maxTiles * 4), oTotals = reserve(16 * 4), oSegs = reserve((size_t)maxSegs * sizeof(SahSeg)),
                 oSegChunk0 = reserve(((size_t)maxSegs + 1) * 4), oBest = reserve((size_t)maxSegs * 8),
                 oChunks = reserve((size_t)maxChunks * sizeof(SahChunk)), oChunkBox = reserve((size_t)maxChunks * sizeof(SahBox)),
                 oBefore = reserve((size_t)maxChunks * sizeof(SahBox)), oAfter = reserve((size_t)maxChunks * sizeof(SahBox));
    ORZ_CUDA(cudaMalloc(&B.arena, total));
    char* base = B.arena;
    B.boxes = reinterpret_cast<float4*>(base + oBoxes);
    B.order = reinterpret_cast<uint32_t*>(base + oOrder);
    B.key[0] = reinterpret_cast<uint32_t*>(base + oKey0); B.key[1] = reinterpret_cast<uint32_t*>(base + oKey1);
    B.idx[0] = reinterpret_cast<uint32_t*>(base + oIdx0); B.idx[1] = reinterpret_cast<uint32_t*>(base + oIdx1);
    B.seg[0] = reinterpret_cast<uint32_t*>(base + oSeg0); B.seg[1] = reinterpret_cast<uint32_t*>(base + oSeg1);
    B.segStatic = reinterpret_cast<uint32_t*>(base + oSegStatic);
    B.areaLeft = reinterpret_cast<float*>(base + oAreaLeft); B.areaRight = reinterpret_cast<float*>(base + oAreaRight);
    B.hist = reinterpret_cast<uint32_t*>(base + oHist); B.digitTotals = reinterpret_cast<uint32_t*>(base + oTotals);
    B.segs = reinterpret_cast<SahSeg*>(base + oSegs); B.segChunk0 = reinterpret_cast<uint32_t*>(base + oSegChunk0);
    B.best = reinterpret_cast<unsigned long long*>(base + oBest);
    B.chunks = reinterpret_cast<SahChunk*>(base + oChunks); B.chunkBox = reinterpret_cast<SahBox*>(base + oChunkBox);
    B.before = reinterpret_cast<SahBox*>(base + oBefore); B.after = reinterpret_cast<SahBox*>(base + oAfter);
  }
  ORZ_CUDA(cudaMemcpyAsync(B.boxes, aabbs, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
  {
    std::vector<uint32_t> iota(n);
    for (uint32_t i = 0; i < n; ++i) iota[i] = i;
    ORZ_CUDA(cudaMemcpyAsync(B.order, iota.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  std::vector<SahSeg> segs;
  std::vector<SahChunk> chunks;
  std::vector<uint32_t> segChunk0;
  std::vector<unsigned long long> best;
  while (!active.empty()) {  // one level of the recursion (SurfaceAreaHeuristic.cpp:77-94) per turn
    const uint32_t nSegs = (uint32_t)active.size();
    if (nSegs > maxSegs) return fail(ORZ_ERR_ARG, "orz_generate_batches_device: internal segment bound exceeded");
    segs.resize(nSegs);
    chunks.clear();
    segChunk0.assign(1, 0u);
    uint32_t M = 0;
    for (uint32_t k = 0; k < nSegs; ++k) {
      if (active[k].n <= 2 * granularity)
        return fail(ORZ_ERR_ARG, "orz_generate_batches_device: a node has no split position with a finite cost (the reference indexes out of bounds here)");
      segs[k] = {M, active[k].start, active[k].n, 0u};
      for (uint32_t p = 0; p < active[k].n; p += kSahChunk) chunks.push_back({k, M + p, std::min<uint32_t>(kSahChunk, active[k].n - p)});
      segChunk0.push_back((uint32_t)chunks.size());
      M += active[k].n;
    }
    const uint32_t nChunks = (uint32_t)chunks.size(), grid = (M + 255) / 256;
    if (nChunks > maxChunks) return fail(ORZ_ERR_ARG, "orz_generate_batches_device: internal chunk bound exceeded");
    ORZ_CUDA(cudaMemcpyAsync(B.segs, segs.data(), nSegs * sizeof(SahSeg), cudaMemcpyHostToDevice, ctx->stream));
    ORZ_CUDA(cudaMemcpyAsync(B.chunks, chunks.data(), nChunks * sizeof(SahChunk), cudaMemcpyHostToDevice, ctx->stream));
    ORZ_CUDA(cudaMemcpyAsync(B.segChunk0, segChunk0.data(), segChunk0.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    ORZ_CUDA(cudaMemsetAsync(B.best, 0xff, (size_t)nSegs * 8, ctx->stream));
    B.cur = 0;
    ORZ_LAUNCH(k_sah_gather, grid, 256, ctx->stream, B.order, B.segs, nSegs, M, B.idx[0], B.seg[0], B.segStatic);
    ctx->launches++;
    for (int axis = 0; axis < 3; ++axis) {
      ORZ_LAUNCH(k_sah_keys, grid, 256, ctx->stream, B.boxes, B.idx[B.cur], B.segStatic, B.segs, axis, M, B.key[B.cur]);
      sah_sort(ctx, B, M, nSegs);
      ORZ_LAUNCH(k_sah_chunk_boxes, nChunks, kSahChunk, ctx->stream, B.boxes, B.idx[B.cur], B.chunks, B.chunkBox);
      ORZ_LAUNCH(k_sah_chunk_scan, nSegs, 64, ctx->stream, B.chunkBox, B.segChunk0, B.before, B.after);
      ORZ_LAUNCH(k_sah_areas, nChunks, kSahChunk, ctx->stream, B.boxes, B.idx[B.cur], B.chunks, B.before, B.after, B.areaLeft, B.areaRight);
      ORZ_LAUNCH(k_sah_costs, grid, 256, ctx->stream, B.areaLeft, B.areaRight, B.segStatic, B.segs, M, granularity, (uint32_t)axis, B.best);
      ctx->launches += 5;
    }
    best.resize(nSegs);
    ORZ_CUDA(cudaMemcpyAsync(best.data(), B.best, (size_t)nSegs * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
    ORZ_CUDA(cudaGetLastError());
    for (uint32_t k = 0; k < nSegs; ++k) {
      if (best[k] == ~0ull)
        return fail(ORZ_ERR_ARG, "orz_generate_batches_device: a node has no split position with a finite cost (the reference indexes out of bounds here)");
      segs[k].axis = (uint32_t)((best[k] >> 30) & 3u);
      const uint32_t split = (uint32_t)(best[k] & 0x3fffffffu);
      if (segs[k].axis > 2 || split < granularity || split >= active[k].n - granularity || split % granularity != 0)
        return fail(ORZ_ERR_CUDA, "orz_generate_batches_device: internal error (split position out of range)");
    }
    // sort every segment by its best axis once more (SurfaceAreaHeuristic.cpp:69-72) and put it back
    ORZ_CUDA(cudaMemcpyAsync(B.segs, segs.data(), nSegs * sizeof(SahSeg), cudaMemcpyHostToDevice, ctx->stream));
    ORZ_LAUNCH(k_sah_keys, grid, 256, ctx->stream, B.boxes, B.idx[B.cur], B.segStatic, B.segs, -1, M, B.key[B.cur]);
    sah_sort(ctx, B, M, nSegs);
    ORZ_LAUNCH(k_sah_scatter_back, grid, 256, ctx->stream, B.idx[B.cur], B.segStatic, B.segs, M, B.order);
    ctx->launches += 2;
    // children: a side smaller than the target is a batch, the rest is split on the next level
    std::vector<Range> next;
    for (uint32_t k = 0; k < nSegs; ++k) {
      const uint32_t split = (uint32_t)(best[k] & 0x3fffffffu);
      const Range child[2] = {{active[k].start, split}, {active[k].start + split, active[k].n - split}};
      for (const Range& c : child) (c.n < targetSize ? leaves : next).push_back(c);
    }
    active.swap(next);
  }
  ORZ_CUDA(cudaMemcpyAsync(indicesOut, B.order, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ORZ_CUDA(cudaStreamSynchronize(ctx->stream));
  ORZ_CUDA(cudaGetLastError());
  // depth-first order of the reference's result = batches by start position
  std::sort(leaves.begin(), leaves.end(), [](const Range& a, const Range& b) { return a.start < b.start; });
  *nBatches = (uint32_t)leaves.size();
  if (leaves.size() > batchCapacity) return fail(ORZ_ERR_ARG, "orz_generate_batches_device: batchSizes too small");
  for (size_t b = 0; b < leaves.size(); ++b) batchSizes[b] = leaves[b].n;
  return ORZ_OK;
}

