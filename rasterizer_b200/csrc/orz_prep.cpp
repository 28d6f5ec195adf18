// Scene preparation on the host: the offline steps the reference's application runs before the hot
// path (Main.cpp:86-107) -- QuadDecomposition::decompose and SurfaceAreaHeuristic::generateBatches --
// written from their observable behaviour, not from their data structures:
//   * the directed-edge lookup is a sorted edge table (the reference uses a hash map of vectors,
//     QuadDecomposition.cpp:358-366; only look-ups by key are observable, never iteration order);
//   * the candidate graph is built as CSR from a time-ordered event list, candidate tests run on all
//     host threads;
//   * the maximum matching keeps the reference's visiting order (greedy pass, then one breadth-first
//     alternating forest per exposed triangle with blossom shrinking, QuadDecomposition.cpp:28-343),
//     because WHICH maximum matching comes out decides the quads; it runs on flat arrays with an
//     explicit work stack instead of recursion, so mesh size is not limited by the call stack;
//   * the SAH split sorts (key, index) pairs instead of indices through a gathering comparator, and
//     independent subtrees are split on separate threads.
// Results are bit-identical with the reference (tests/test_scene_prep.py): same quads in the same
// order, same batches in the same order.  The device version of the batching is orz_sah_kernels.cuh.
// Build with -ffp-contract=off.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/orz.h"
#include "orz_host.h"

namespace {

using orz::set_error;

unsigned host_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n ? std::min(n, 64u) : 4u;
}

template <typename F>
void parallel_ranges(size_t n, size_t grain, F&& body) {
  const size_t parts = std::max<size_t>(1, std::min<size_t>(host_threads(), n / std::max<size_t>(grain, 1)));
  if (parts <= 1) {
    body(0, size_t(0), n);
    return;
  }
  std::vector<std::thread> th;
  for (size_t p = 0; p < parts; ++p) th.emplace_back([&, p] { body(p, n * p / parts, n * (p + 1) / parts); });
  for (auto& t : th) t.join();
}

// ------------------------------------------------------------------------------------------------
// Quad decomposition

// canMergeTrianglesToQuad, QuadDecomposition.cpp:327-343: both triangle planes of the quad
// (v0 v1 v2) / (v2 v3 v0) keep the diagonal v1 - v3 within 0.5 world units.  comigt is false when
// unordered, so a NaN distance (collapsed triangle) does not reject.
bool can_merge(const float* v0, const float* v1, const float* v2, const float* v3) {
  const orz::V3 n0 = orz::normalized(orz::tri_normal(v0, v1, v2));
  const orz::V3 n2 = orz::normalized(orz::tri_normal(v2, v3, v0));
  const orz::V3 d = orz::sub(v1, v3);
  const float a = fabsf(orz::dot3(n0, d)), b = fabsf(orz::dot3(n2, d));
  return !(a > 0.5f || b > 0.5f);
}

// One directed triangle edge (from -> to) with the triangle's third vertex; seq = 3 * triangle + edge
// is the time the reference inserts it into its map (QuadDecomposition.cpp:364-366).
struct EdgeRec {
  uint64_t key;
  uint32_t seq;
  uint32_t apex;
};
inline uint64_t edge_key(uint32_t from, uint32_t to) { return (uint64_t(from) << 32) | to; }

struct EdgeTable {
  std::vector<EdgeRec> recs;  // sorted by (key, seq)
  // entries of `key` in insertion order: [first, last)
  std::pair<const EdgeRec*, const EdgeRec*> find(uint64_t key) const {
    auto lo = std::lower_bound(recs.begin(), recs.end(), key, [](const EdgeRec& r, uint64_t k) { return r.key < k; });
    auto hi = lo;
    while (hi != recs.end() && hi->key == key) ++hi;
    return {recs.data() + (lo - recs.begin()), recs.data() + (hi - recs.begin())};
  }
};

// Maximum-cardinality matching on the candidate graph.  The reference's visiting order is part of
// the result (which of several maximum matchings), so it is kept: vertices in index order, neighbours
// in adjacency order, breadth-first forest with a FIFO queue, blossoms merged into the common
// ancestor's set with the ancestor as representative.
class Matcher {
 public:
  Matcher(const std::vector<uint32_t>& offsets, const std::vector<int32_t>& adjacency)
      : off_(offsets), adj_(adjacency), n_(int32_t(offsets.size()) - 1), mate_(n_, -1), node_(n_), bridge_(n_) {}

  // false: the search stopped making progress.  Self-paired triangles -- (a, b, a) or (a, a, a) faces
  // are their own neighbour -- can put a vertex into the forest as its own mate, and the reference
  // then walks parent links forever; every loop below runs against a step budget instead.
  bool run() {
    std::vector<int32_t> exposed;
    // greedy maximal matching, QuadDecomposition.cpp:35-57
    for (int32_t v = 0; v < n_; ++v) {
      if (mate_[v] != -1) continue;
      bool found = false;
      for (uint32_t k = off_[v]; k < off_[v + 1]; ++k) {
        const int32_t w = adj_[k];
        if (mate_[w] == -1) {
          mate_[v] = w;
          mate_[w] = v;
          found = true;
          break;
        }
      }
      if (!found) exposed.push_back(v);
    }
    // one alternating forest per triangle that is still exposed, QuadDecomposition.cpp:59-70
    for (int32_t root : exposed) {
      if (mate_[root] != -1) continue;
      path_.clear();
      budget_ = 64 * int64_t(n_) + 4096;
      if (grow_forest(root))
        for (size_t i = 0; i + 1 < path_.size(); i += 2) {
          mate_[path_[i]] = path_[i + 1];
          mate_[path_[i + 1]] = path_[i];
        }
      if (budget_ < 0) return false;
    }
    return true;
  }
  int32_t mate(int32_t v) const { return mate_[v]; }

 private:
  struct Node {
    int32_t depth = 0, parent = -1, set = 0;
    uint32_t epoch = 0;  // forest this node was last put into
  };

  // representative of v's blossom in the current forest; a vertex outside the forest is its own
  int32_t find(int32_t x) {
    int32_t r = x;
    while (node_[r].epoch == epoch_ && node_[r].set != r) r = node_[r].set;
    while (node_[x].epoch == epoch_ && node_[x].set != x) {  // path compression
      const int32_t next = node_[x].set;
      node_[x].set = r;
      x = next;
    }
    return r;
  }
  void unite(int32_t x, int32_t y) {  // QuadDecomposition.cpp:270-274
    const int32_t rx = find(x);
    node_[rx].set = find(y);
  }
  void make_representative(int32_t x) {  // QuadDecomposition.cpp:276-281
    const int32_t rx = find(x);
    node_[rx].set = x;
    node_[x].set = x;
  }
  void enter(int32_t v, int32_t depth, int32_t parent) {
    node_[v].depth = depth;
    node_[v].parent = parent;
    node_[v].epoch = epoch_;
    node_[v].set = v;
  }

  bool grow_forest(int32_t root) {  // QuadDecomposition.cpp:95-131
    ++epoch_;
    enter(root, 0, -1);
    queue_.clear();
    head_ = 0;
    queue_.push_back(root);
    while (head_ < queue_.size() && spend()) {
      const int32_t v = queue_[head_++];
      for (uint32_t k = off_[v]; k < off_[v + 1]; ++k)
        if (examine(root, v, adj_[k])) return true;
    }
    return false;
  }

  bool examine(int32_t root, int32_t v, int32_t w) {  // QuadDecomposition.cpp:133-160
    const int32_t rv = find(v), rw = find(w);
    if (rv == rw) return false;
    if (node_[rw].epoch != epoch_) {
      if (mate_[w] == -1) {  // augmenting path: w, then the alternating path v .. root
        path_.push_back(w);
        trace_path(v, root);
        return true;
      }
      // extend the forest by the matched edge (w, u), QuadDecomposition.cpp:168-187
      const int32_t u = mate_[w];
      const int32_t dw = node_[v].depth + 1 + (node_[v].depth & 1);
      enter(w, dw, v);
      enter(u, dw + 1, w);
      queue_.push_back(u);
    } else if (node_[rw].depth % 2 == 0) {
      // odd cycle: shrink it into the common ancestor's blossom, QuadDecomposition.cpp:189-231
      int32_t a = v, b = w;
      while (b != a && a >= 0 && b >= 0 && spend()) {
        if (node_[a].depth > node_[b].depth) a = node_[a].parent; else b = node_[b].parent;
      }
      if (a != b || a < 0) {
        budget_ = -1;
        return false;
      }
      const int32_t base = find(a);
      shrink_side(base, v, w);
      shrink_side(base, w, v);
    }
    return false;
  }

  void shrink_side(int32_t base, int32_t v, int32_t w) {  // QuadDecomposition.cpp:197-213
    int32_t u = find(v);
    while (u != base && spend()) {
      unite(base, u);
      u = mate_[u];
      unite(base, u);
      make_representative(base);
      queue_.push_back(u);
      bridge_[u] = {v, w};
      if (node_[u].parent < 0) {
        budget_ = -1;
        return;
      }
      u = find(node_[u].parent);
    }
  }

  // alternating path s .. t through the forest (QuadDecomposition.cpp:233-262), appended to path_.
  // An odd vertex is left through its blossom's bridge: the stretch bridge.first .. mate(s) is
  // traced forwards and then reversed in place.
  void trace_path(int32_t s0, int32_t t0) {
    struct Task { int32_t s, t; size_t reverseFrom; bool isReverse; };
    std::vector<Task> stack;
    stack.push_back({s0, t0, 0, false});
    while (!stack.empty() && spend()) {
      Task task = stack.back();
      stack.pop_back();
      if (task.isReverse) {
        std::reverse(path_.begin() + task.reverseFrom, path_.end());
        continue;
      }
      int32_t s = task.s;
      const int32_t t = task.t;
      while (spend()) {
        if (s < 0) {
          budget_ = -1;
          break;
        }
        if (s == t) {
          path_.push_back(s);
          break;
        }
        if (node_[s].depth % 2 == 0) {
          path_.push_back(s);
          path_.push_back(mate_[s]);
          s = node_[mate_[s]].parent;
          continue;
        }
        const std::pair<int32_t, int32_t> br = bridge_[s];
        path_.push_back(s);
        stack.push_back({br.second, t, 0, false});
        stack.push_back({0, 0, path_.size(), true});
        stack.push_back({br.first, mate_[s], 0, false});
        break;
      }
    }
  }

  const std::vector<uint32_t>& off_;
  const std::vector<int32_t>& adj_;
  int32_t n_;
  std::vector<int32_t> mate_;
  std::vector<Node> node_;
  std::vector<std::pair<int32_t, int32_t>> bridge_;
  uint32_t epoch_ = 0;
  int64_t budget_ = 0;  // steps left for the current forest
  bool spend() { return --budget_ >= 0; }
  std::vector<int32_t> queue_;
  size_t head_ = 0;
  std::vector<int32_t> path_;
};

// ------------------------------------------------------------------------------------------------
// SAH batching

struct KeyIdx {
  uint32_t key;  // the centre as an order-preserving integer, -0 == +0 (the same key the device batching sorts by)
  uint32_t idx;
};
inline uint32_t order_key(float c) {
  uint32_t u;
  memcpy(&u, &c, 4);
  if ((u & 0x7fffffffu) == 0) u = 0;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct SahInput {
  const float* boxes;  // n x (min4, max4)
  std::vector<float> center[3];  // Aabb::getCenter (min + max, VectorMath.h:50-53) per axis
  uint32_t target, granularity;
};

// minps / maxps keep the SECOND operand when equal or unordered (Aabb::include, VectorMath.h:38-48)
inline float min_x86(float a, float b) { return a < b ? a : b; }
inline float max_x86(float a, float b) { return a > b ? a : b; }

struct Box3 {
  float mn[3], mx[3];
  Box3() {
    for (int k = 0; k < 3; ++k) { mn[k] = INFINITY; mx[k] = -INFINITY; }
  }
  void include(const float* box) {
    for (int k = 0; k < 3; ++k) { mn[k] = min_x86(mn[k], box[k]); mx[k] = max_x86(mx[k], box[4 + k]); }
  }
  // Aabb::surfaceArea (VectorMath.h:60-65): dpps 0x7F of extents with extents.yzx
  float area() const {
    const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    return (ex * ey + ey * ez) + ez * ex;
  }
};

// Stable sort of a node's range by the centre on one axis: least-significant-digit radix sort over 11 + 11 + 10
// bits (a pass whose digit is the same for every element is skipped); ties keep the order the previous sort left,
// exactly what std::stable_sort with a `<` comparator does in the reference (SurfaceAreaHeuristic.cpp:22-24).
void sort_by_axis(const SahInput& in, int axis, uint32_t* first, uint32_t n, std::vector<KeyIdx>& tmp) {
  tmp.resize(2 * size_t(n));
  KeyIdx *src = tmp.data(), *dst = tmp.data() + n;
  const float* c = in.center[axis].data();
  for (uint32_t i = 0; i < n; ++i) src[i] = {order_key(c[first[i]]), first[i]};
  static const int kShift[3] = {0, 11, 22}, kBits[3] = {11, 11, 10};
  uint32_t count[2048];
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t mask = (1u << kBits[pass]) - 1u, shift = uint32_t(kShift[pass]);
    memset(count, 0, sizeof(uint32_t) << kBits[pass]);
    for (uint32_t i = 0; i < n; ++i) ++count[(src[i].key >> shift) & mask];
    if (count[(src[0].key >> shift) & mask] == n) continue;
    uint32_t run = 0;
    for (uint32_t d = 0; d <= mask; ++d) {
      const uint32_t here = count[d];
      count[d] = run;
      run += here;
    }
    for (uint32_t i = 0; i < n; ++i) dst[count[(src[i].key >> shift) & mask]++] = src[i];
    std::swap(src, dst);
  }
  for (uint32_t i = 0; i < n; ++i) first[i] = src[i].idx;
}

// sahSplit, SurfaceAreaHeuristic.cpp:10-75.  Returns the split position, 0 when no candidate has a
// finite cost (the reference then indexes with -1).
uint32_t sah_split(const SahInput& in, uint32_t* first, uint32_t n) {
  const uint32_t g = in.granularity;
  float bestCost = INFINITY;
  int bestAxis = -1;
  uint32_t bestIndex = 0;
  std::vector<KeyIdx> tmp;
  std::vector<float> areaLeft(n), areaRight(n);
  for (int axis = 0; axis < 3; ++axis) {
    sort_by_axis(in, axis, first, n, tmp);  // stable: ties keep the order the previous sort left
    Box3 acc;
    for (uint32_t i = 0; i < n; ++i) {
      acc.include(in.boxes + 8 * size_t(first[i]));
      areaLeft[i] = acc.area();
    }
    acc = Box3();
    for (uint32_t i = n; i-- > 0;) {
      acc.include(in.boxes + 8 * size_t(first[i]));
      areaRight[i] = acc.area();
    }
    for (uint32_t s = g; s < n - g; s += g) {
      const float cost = areaLeft[s - 1] * float(int32_t(s)) + areaRight[s] * float(int32_t(n - s));
      if (cost < bestCost) {
        bestCost = cost;
        bestAxis = axis;
        bestIndex = s;
      }
    }
  }
  if (bestAxis < 0) return 0;
  sort_by_axis(in, bestAxis, first, n, tmp);
  return bestIndex;
}

struct SahState {
  const SahInput* in;
  uint32_t* base;
  std::mutex lock;
  std::vector<std::pair<uint32_t, uint32_t>> leaves;  // (start, size)
  std::atomic<bool> failed{false};
  std::atomic<int> spareThreads{0};
};

bool take_thread(SahState& st) {
  int spare = st.spareThreads.load();
  while (spare > 0 && !st.spareThreads.compare_exchange_weak(spare, spare - 1)) {}
  return spare > 0;
}

// generateBatchesRecursive, SurfaceAreaHeuristic.cpp:77-94: a side smaller than the target is a
// batch, anything else is split again.  Batches come out in depth-first order = by start position.
void sah_recurse(SahState& st, uint32_t start, uint32_t n) {
  if (st.failed.load()) return;
  if (n <= 2 * st.in->granularity) {  // no candidate position: the reference would read out of bounds
    st.failed = true;
    return;
  }
  const uint32_t split = sah_split(*st.in, st.base + start, n);
  if (split == 0) {
    st.failed = true;
    return;
  }
  const uint32_t childStart[2] = {start, start + split}, childSize[2] = {split, n - split};
  std::thread side;
  for (int c = 0; c < 2; ++c) {
    if (childSize[c] < st.in->target) {
      std::lock_guard<std::mutex> guard(st.lock);
      st.leaves.push_back({childStart[c], childSize[c]});
    } else if (c == 0 && childSize[0] >= 4096 && childSize[1] >= 4096 && take_thread(st)) {
      side = std::thread([&st, s = childStart[0], m = childSize[0]] { sah_recurse(st, s, m); });
    } else {
      sah_recurse(st, childStart[c], childSize[c]);
    }
  }
  if (side.joinable()) {
    side.join();
    st.spareThreads.fetch_add(1);
  }
}

}  // namespace

// QuadDecomposition::decompose, QuadDecomposition.h:10 / QuadDecomposition.cpp:346-445
extern "C" int orz_quad_decompose(const uint32_t* indices, size_t nIndices, const float* vertices, size_t nVertices,
                                  uint32_t* quadIndices, size_t* nQuadIndices) {
  if ((!indices && nIndices) || !vertices || !quadIndices || !nQuadIndices)
    return set_error(ORZ_ERR_ARG, "orz_quad_decompose: bad arguments");
  const size_t nTris = nIndices / 3;
  if (nTris >= (size_t(1) << 30)) return set_error(ORZ_ERR_ARG, "orz_quad_decompose: too many triangles");
  for (size_t i = 0; i < 3 * nTris; ++i)
    if (indices[i] >= nVertices) return set_error(ORZ_ERR_ARG, "orz_quad_decompose: vertex index out of range");

  // directed edges, in the order the reference inserts them
  EdgeTable table;
  table.recs.resize(3 * nTris);
  parallel_ranges(nTris, 1 << 14, [&](size_t, size_t t0, size_t t1) {
    for (size_t t = t0; t < t1; ++t) {
      const uint32_t* i = indices + 3 * t;
      for (uint32_t e = 0; e < 3; ++e) table.recs[3 * t + e] = {edge_key(i[e], i[(e + 1) % 3]), uint32_t(3 * t + e), i[(e + 2) % 3]};
    }
  });
  std::sort(table.recs.begin(), table.recs.end(),
            [](const EdgeRec& a, const EdgeRec& b) { return a.key != b.key ? a.key < b.key : a.seq < b.seq; });

  // candidate pairs in time order: triangle t against every triangle <= t that owns the opposite
  // direction of one of t's edges (QuadDecomposition.cpp:368-386)
  struct Pair { uint32_t t, other; };
  std::vector<std::vector<Pair>> found(host_threads() + 1);
  parallel_ranges(nTris, 1 << 12, [&](size_t part, size_t t0, size_t t1) {
    std::vector<Pair>& out = found[part];
    for (size_t t = t0; t < t1; ++t) {
      const uint32_t* i = indices + 3 * t;
      const uint32_t seqEnd = uint32_t(3 * t + 3);
      for (uint32_t e = 0; e < 3; ++e) {
        const uint32_t a = i[e], b = i[(e + 1) % 3], c = i[(e + 2) % 3];
        auto range = table.find(edge_key(b, a));
        for (const EdgeRec* r = range.first; r != range.second && r->seq < seqEnd; ++r)
          if (can_merge(vertices + 4 * size_t(a), vertices + 4 * size_t(r->apex), vertices + 4 * size_t(b), vertices + 4 * size_t(c)))
            out.push_back({uint32_t(t), r->seq / 3});
      }
    }
  });

  // CSR adjacency: each pair appends `other` to t's list and t to other's list, in time order
  std::vector<uint32_t> offsets(nTris + 1, 0);
  size_t nPairs = 0;
  for (const auto& part : found) {
    nPairs += part.size();
    for (const Pair& p : part) {
      ++offsets[p.t + 1];
      ++offsets[p.other + 1];
    }
  }
  for (size_t t = 0; t < nTris; ++t) offsets[t + 1] += offsets[t];
  std::vector<int32_t> adjacency(2 * nPairs);
  {
    std::vector<uint32_t> cursor(offsets.begin(), offsets.end() - 1);
    for (const auto& part : found)
      for (const Pair& p : part) {
        adjacency[cursor[p.t]++] = int32_t(p.other);
        adjacency[cursor[p.other]++] = int32_t(p.t);
      }
  }
  found.clear();

  Matcher matcher(offsets, adjacency);
  if (!matcher.run())
    return set_error(ORZ_ERR_ARG, "orz_quad_decompose: the matching does not terminate on this mesh (self-paired degenerate triangles; the reference loops forever)");

  // output, QuadDecomposition.cpp:394-441: a lone triangle becomes the quad (i0, i2, i1, i0); a pair is
  // written once, by its lower-numbered triangle, starting at the vertex after the shared edge
  size_t w = 0;
  for (size_t t = 0; t < nTris; ++t) {
    const uint32_t* i = indices + 3 * t;
    const int32_t other = matcher.mate(int32_t(t));
    if (other == -1) {
      quadIndices[w++] = i[0];
      quadIndices[w++] = i[2];
      quadIndices[w++] = i[1];
      quadIndices[w++] = i[0];
    } else if (uint32_t(t) < uint32_t(other)) {
      bool done = false;
      for (uint32_t e = 0; e < 3 && !done; ++e) {
        auto range = table.find(edge_key(i[(e + 1) % 3], i[e]));
        for (const EdgeRec* r = range.first; r != range.second; ++r)
          if (r->seq / 3 == uint32_t(other)) {
            quadIndices[w++] = i[e];
            quadIndices[w++] = i[(e + 2) % 3];
            quadIndices[w++] = i[(e + 1) % 3];
            quadIndices[w++] = r->apex;
            done = true;
            break;
          }
      }
    }
  }
  *nQuadIndices = w;
  return ORZ_OK;
}

// SurfaceAreaHeuristic::generateBatches, SurfaceAreaHeuristic.h:10 / SurfaceAreaHeuristic.cpp:96-104
extern "C" int orz_generate_batches(const float* aabbs, uint32_t nAabbs, uint32_t targetSize, uint32_t splitGranularity,
                                    uint32_t* indicesOut, uint32_t* batchSizes, uint32_t batchCapacity, uint32_t* nBatches) {
  if (!aabbs || !indicesOut || !batchSizes || !nBatches || splitGranularity == 0 || nAabbs >= (1u << 30))
    return set_error(ORZ_ERR_ARG, "orz_generate_batches: bad arguments");
  SahInput in;
  in.boxes = aabbs;
  in.target = targetSize;
  in.granularity = splitGranularity;
  for (int k = 0; k < 3; ++k) {
    in.center[k].resize(nAabbs);
    for (uint32_t i = 0; i < nAabbs; ++i) in.center[k][i] = aabbs[8 * size_t(i) + k] + aabbs[8 * size_t(i) + 4 + k];
  }
  for (uint32_t i = 0; i < nAabbs; ++i) indicesOut[i] = i;
  SahState st;
  st.in = &in;
  st.base = indicesOut;
  st.spareThreads = int(host_threads()) - 1;
  sah_recurse(st, 0, nAabbs);  // the root is always split, whatever its size (SurfaceAreaHeuristic.cpp:102)
  if (st.failed) return set_error(ORZ_ERR_ARG, "orz_generate_batches: a node has no split position with a finite cost (the reference indexes out of bounds here)");
  std::sort(st.leaves.begin(), st.leaves.end());
  *nBatches = uint32_t(st.leaves.size());
  if (st.leaves.size() > batchCapacity) return set_error(ORZ_ERR_ARG, "orz_generate_batches: batchSizes too small");
  for (size_t b = 0; b < st.leaves.size(); ++b) batchSizes[b] = st.leaves[b].second;
  return ORZ_OK;
}

// Main.cpp:86-128 in one call: mesh -> quads (host) -> pad -> per-quad boxes -> SAH batches (GPU) -> reference box
// -> Occluder::bake of every batch (GPU).  The scene never exists on the host in baked form.
extern "C" int orz_scene_from_mesh(orz_context* ctx, const uint32_t* indices, size_t nIndices, const float* vertices, size_t nVertices,
                                   uint32_t targetSize, uint32_t splitGranularity, int occludeesFromQuads, orz_scene** out,
                                   orz_mesh_scene_info* info) {
  if (!ctx || !out || !vertices || nVertices == 0 || nIndices < 3) return set_error(ORZ_ERR_ARG, "orz_scene_from_mesh: bad arguments");
  if (splitGranularity == 0 || splitGranularity % 8 != 0 || targetSize > 10240)
    return set_error(ORZ_ERR_ARG, "orz_scene_from_mesh: splitGranularity must be a multiple of 8 (Occluder::bake packs 8 quads, Occluder.cpp:108) "
                                  "and targetSize at most 10 240 (device bake)");
  std::vector<uint32_t> quads(4 * (nIndices / 3));
  size_t words = 0;
  int rc = orz_quad_decompose(indices, nIndices, vertices, nVertices, quads.data(), &words);
  if (rc != ORZ_OK) return rc;
  quads.resize(words);
  while (quads.size() % 32 != 0) quads.push_back(quads[0]);  // Main.cpp:91-94
  const uint32_t nQuads = uint32_t(quads.size() / 4);
  // Main.cpp:96-105: Aabb::include over the four corners, all four lanes
  std::vector<float> boxes(size_t(nQuads) * 8);
  parallel_ranges(nQuads, 1 << 14, [&](size_t, size_t q0, size_t q1) {
    for (size_t q = q0; q < q1; ++q) {
      float* b = boxes.data() + 8 * q;
      for (int k = 0; k < 4; ++k) { b[k] = INFINITY; b[4 + k] = -INFINITY; }
      for (int c = 0; c < 4; ++c) {
        const float* v = vertices + 4 * size_t(quads[4 * q + c]);
        for (int k = 0; k < 4; ++k) { b[k] = min_x86(b[k], v[k]); b[4 + k] = max_x86(b[4 + k], v[k]); }
      }
    }
  });
  std::vector<uint32_t> order(nQuads), sizes(nQuads / splitGranularity + 2);
  uint32_t nBatches = 0;
  rc = orz_generate_batches_device(ctx, boxes.data(), nQuads, targetSize, splitGranularity, order.data(), sizes.data(), uint32_t(sizes.size()), &nBatches);
  if (rc != ORZ_OK) return rc;
  float refMin[4], refMax[4];  // Main.cpp:109-113: over ALL vertices, referenced or not
  for (int k = 0; k < 4; ++k) { refMin[k] = INFINITY; refMax[k] = -INFINITY; }
  for (size_t v = 0; v < nVertices; ++v)
    for (int k = 0; k < 4; ++k) { refMin[k] = min_x86(refMin[k], vertices[4 * v + k]); refMax[k] = max_x86(refMax[k], vertices[4 * v + k]); }
  // Main.cpp:116-128: the batches' corner positions, then bake
  std::vector<float> batchVerts(size_t(nQuads) * 16), occludees(occludeesFromQuads ? size_t(nQuads) * 8 : 0);
  std::vector<uint32_t> vertCounts(nBatches);
  for (uint32_t b = 0; b < nBatches; ++b) vertCounts[b] = sizes[b] * 4;
  parallel_ranges(nQuads, 1 << 14, [&](size_t, size_t q0, size_t q1) {
    for (size_t q = q0; q < q1; ++q) {
      const uint32_t src = order[q];
      for (int c = 0; c < 4; ++c) memcpy(batchVerts.data() + 16 * q + 4 * c, vertices + 4 * size_t(quads[4 * size_t(src) + c]), 16);
      if (occludeesFromQuads) {
        memcpy(occludees.data() + 8 * q, boxes.data() + 8 * size_t(src), 32);
        occludees[8 * q + 3] = occludees[8 * q + 7] = 1.0f;  // w := 1 as Occluder.cpp:172-173 does for its bounds
      }
    }
  });
  orz_scene* scene = nullptr;
  rc = orz_scene_bake(ctx, batchVerts.data(), vertCounts.data(), nBatches, refMin, refMax, nullptr, nullptr, nullptr, nullptr, &scene);
  if (rc != ORZ_OK) return rc;
  if (occludeesFromQuads) {
    rc = orz_scene_set_occludees(scene, occludees.data(), nQuads);
    if (rc != ORZ_OK) {
      orz_scene_destroy(scene);
      return rc;
    }
  }
  if (info) {
    info->nOccluders = nBatches;
    info->nQuads = nQuads;
    memcpy(info->refMin, refMin, 16);
    memcpy(info->refMax, refMax, 16);
  }
  *out = scene;
  return ORZ_OK;
}
