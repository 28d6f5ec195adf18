// Host-side pieces of the hot path that stay on the CPU by design:
//   * Occluder::bake (Occluder.cpp:7-181): runs once per batch, offline; its result (16 bytes per
//     quad) is what gets uploaded to HBM.  Quad order inside a batch is rasterisation order, so
//     the k-means regrouping has to reproduce the reference's float operations exactly,
//     including rsqrtps (VectorMath.h:20-23) -- taken from this CPU unless a table is installed.
//   * the 64x64 edge-mask table (Rasterizer.cpp:547-604): libm cosf/sinf define it, so it is
//     built on the host (multi-threaded, once per process) and uploaded (32 KB).
//   * the rcpps probe: the kernels evaluate x86 rcpps through a table of this CPU's own results
//     (SURVEY 7.1), because coverage and visibility bits depend on it bit for bit.
// Build with -ffp-contract=off: only the fmaf() calls below may fuse.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>

#include <mutex>
#include <thread>
#include <vector>

#include "../../include/orz.h"
#include "orz_core.h"
#include "orz_host.h"

namespace orz {

// ---------------------------------------------------------------------------------------------
static float host_rcp(float x) {
  float y;
  _mm_store_ss(&y, _mm_rcp_ss(_mm_set_ss(x)));
  return y;
}
static float host_rsqrt(float x) {
  float y;
  _mm_store_ss(&y, _mm_rsqrt_ss(_mm_set_ss(x)));
  return y;
}

// Smallest k such that rcpps(1.m) depends only on the top k mantissa bits on this CPU
// (11 on Intel); the table has 2^k entries.  `exact` reports whether the
// exponent / special-value model of rcp_x86() reproduced the instruction on a sample sweep.
void probe_host_rcp(std::vector<uint32_t>& table, int& bits, bool& exact) {
  const uint32_t N = 1u << 23;
  std::vector<uint32_t> full(N);
  for (uint32_t m = 0; m < N; ++m) full[m] = f2u(host_rcp(u2f(0x3f800000u | m)));
  int k = 0;
  for (; k < 23; ++k) {
    const uint32_t group = N >> k;
    bool ok = true;
    for (uint32_t g = 0; g < (1u << k) && ok; ++g) {
      const uint32_t v = full[g * group];
      for (uint32_t i = 1; i < group; ++i)
        if (full[g * group + i] != v) { ok = false; break; }
    }
    if (ok) break;
  }
  bits = k;
  table.resize(size_t(1) << k);
  for (uint32_t g = 0; g < (1u << k); ++g) table[g] = full[size_t(g) << (23 - k)];
  // sample sweep over exponents, signs and specials
  RcpTable rt{table.data(), 23 - k};
  exact = true;
  uint32_t lcg = 12345u;
  const uint32_t exps[] = {0, 1, 2, 3, 64, 125, 126, 127, 128, 129, 200, 251, 252, 253, 254, 255};
  for (uint32_t e : exps)
    for (int i = 0; i < 2048; ++i) {
      lcg = lcg * 1664525u + 1013904223u;
      uint32_t in = (lcg & 0x807fffffu) | (e << 23);
      if (i < 4) in = (in & 0x80000000u) | (e << 23) | (i == 1 ? 0x7fffffu : (i == 2 ? 1u : 0u));
      if (f2u(host_rcp(u2f(in))) != f2u(rcp_x86(u2f(in), rt))) exact = false;
    }
}

// ---------------------------------------------------------------------------------------------
// Edge-mask table.  Index = slope << 6 | offset; a pixel (x, y) of the 8x8 block is covered
// when offset + (x-3.5)/8 nx + (y-3.5)/8 ny <= 0; entries are OR-accumulated over 2000 angles x
// 2000 offsets and stored with pixel (x, y) at bit 8x + (y odd ? 0 : 4) + (y >> 1), the order
// the depth rows are packed in (Rasterizer.cpp:527-545, 1257-1268).
static void lut_slice(uint32_t i0, uint32_t i1, int64_t* acc) {
  const float offMul = (64 - 1) / (0.45f - (-0.45f));
  const float offAdd = 0.5f - (-0.45f) * offMul;
  const float smul = (64 / 2 - 1) * 0.5f, sadd = smul + 0.5f;
  for (uint32_t i = i0; i < i1; ++i) {
    const float angle = -0.1f + 6.4f * float(i) / (2000 - 1);
    float nx = cosf(angle), ny = sinf(angle);
    const float l = 1.0f / (fabsf(nx) + fabsf(ny));
    nx *= l; ny *= l;
    // the table builder's quantiser tests ny < 0 (Rasterizer.cpp:472); the runtime one ny <= 0
    const uint32_t slope = uint32_t(((cvtt_x86(fmaf(nx, smul, sadd)) << 1) + (ny < 0.0f ? 1 : 0)) << 6);
    float tx[8], ty[8];
    for (int k = 0; k < 8; ++k) { tx[k] = (k - 3.5f) / 8.0f * nx; ty[k] = (k - 3.5f) / 8.0f * ny; }
    for (uint32_t j = 0; j < 2000; ++j) {
      const float offset = -0.6f + 1.2f * float(j) / (2000 - 1);
      int32_t q = int32_t(offset * offMul + offAdd);
      q = q < 0 ? 0 : (q > 63 ? 63 : q);
      uint64_t bits = 0;
      for (int x = 0; x < 8; ++x) {
        const float ox = offset + tx[x];
        for (int y = 0; y < 8; ++y)
          if (ox + ty[y] <= 0.0f) bits |= uint64_t(1) << (8 * x + ((y & 1) ? 0 : 4) + (y >> 1));
      }
      acc[slope | uint32_t(q)] |= int64_t(bits);
    }
  }
}

const int64_t* edge_mask_table() {
  static std::once_flag once;
  static std::vector<int64_t> lut(4096, 0);
  std::call_once(once, [] {
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 1 : (nt > 16 ? 16 : nt);
    std::vector<std::vector<int64_t>> part(nt, std::vector<int64_t>(4096, 0));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back(lut_slice, 2000u * t / nt, 2000u * (t + 1) / nt, part[t].data());
    for (auto& x : th) x.join();
    for (unsigned t = 0; t < nt; ++t)
      for (int i = 0; i < 4096; ++i) lut[i] |= part[t][i];
  });
  return lut.data();
}

// ---------------------------------------------------------------------------------------------
// rsqrtps model for host-independent baking (tests install the table of the host the golden
// vectors were made on): x = 2^(2k+p) * 1.m -> 2^-k * table[p][m >> shift].
static std::vector<uint32_t> g_rsqrtTable;
static int g_rsqrtBits = 0;
static inline float rsqrt_x86(float x) {
  if (!g_rsqrtBits) return host_rsqrt(x);
  const RsqrtTable t{g_rsqrtTable.data(), g_rsqrtBits};
  return orz::rsqrt_x86(x, t);
}

// rsqrtps(y) for y in [1, 4): smallest number of leading mantissa bits it depends on (10 on Intel).
// `exact` = the exponent / special-value model reproduced the instruction on a sample sweep.
void probe_host_rsqrt(std::vector<uint32_t>& table, int& bits, bool& exact) {
  const uint32_t N = 1u << 23;
  std::vector<uint32_t> full(2 * size_t(N));
  for (uint32_t p = 0; p < 2; ++p)
    for (uint32_t m = 0; m < N; ++m) full[size_t(p) * N + m] = f2u(host_rsqrt(u2f(((127u + p) << 23) | m)));
  int k = 0;
  for (; k < 23; ++k) {
    const uint32_t group = N >> k;
    bool ok = true;
    for (uint32_t p = 0; p < 2 && ok; ++p)
      for (uint32_t g = 0; g < (1u << k) && ok; ++g) {
        const uint32_t v = full[size_t(p) * N + size_t(g) * group];
        for (uint32_t i = 1; i < group; ++i)
          if (full[size_t(p) * N + size_t(g) * group + i] != v) { ok = false; break; }
      }
    if (ok) break;
  }
  bits = k;
  table.resize(size_t(2) << k);
  for (uint32_t p = 0; p < 2; ++p)
    for (uint32_t g = 0; g < (1u << k); ++g) table[(size_t(p) << k) + g] = full[size_t(p) * N + (size_t(g) << (23 - k))];
  const RsqrtTable rt{table.data(), k};
  exact = true;
  uint32_t lcg = 777u;
  const uint32_t exps[] = {0, 1, 2, 3, 64, 125, 126, 127, 128, 129, 200, 251, 252, 253, 254, 255};
  for (uint32_t e : exps)
    for (int i = 0; i < 2048; ++i) {
      lcg = lcg * 1664525u + 1013904223u;
      uint32_t in = (lcg & 0x807fffffu) | (e << 23);
      if (i < 4) in = (in & 0x80000000u) | (e << 23) | (i == 1 ? 0x7fffffu : (i == 2 ? 1u : 0u));
      if (f2u(host_rsqrt(u2f(in))) != f2u(orz::rsqrt_x86(u2f(in), rt))) exact = false;
    }
}
// the table the bake currently uses: the installed one, or this CPU's
void current_rsqrt_table(std::vector<uint32_t>& table, int& bits) {
  if (g_rsqrtBits) { table = g_rsqrtTable; bits = g_rsqrtBits; return; }
  static std::vector<uint32_t> host;
  static int hostBits = 0;
  static std::once_flag once;
  std::call_once(once, [] { bool ex; probe_host_rsqrt(host, hostBits, ex); });
  table = host; bits = hostBits;
}

float rsqrt_current(float x) { return rsqrt_x86(x); }

}  // namespace orz

using namespace orz;

extern "C" int orz_set_rsqrt_table(const uint32_t* table, int bits) {
  if (!table) { g_rsqrtTable.clear(); g_rsqrtBits = 0; return ORZ_OK; }
  if (bits < 1 || bits > 23) return ORZ_ERR_ARG;
  g_rsqrtTable.assign(table, table + (size_t(2) << bits));
  g_rsqrtBits = bits;
  return ORZ_OK;
}

extern "C" int orz_edge_mask_table(int64_t* lut4096) {
  if (!lut4096) return ORZ_ERR_ARG;
  memcpy(lut4096, edge_mask_table(), 4096 * sizeof(int64_t));
  return ORZ_OK;
}

extern "C" int orz_probe_host_rcp(uint32_t* table, int* bits, int* exact) {
  std::vector<uint32_t> t;
  int b = 0;
  bool ex = false;
  probe_host_rcp(t, b, ex);
  if (bits) *bits = b;
  if (exact) *exact = ex ? 1 : 0;
  if (table) memcpy(table, t.data(), t.size() * 4);
  return ORZ_OK;
}

// Occluder::bake, Occluder.cpp:7-181
extern "C" uint32_t orz_bake(const float* vertices, uint32_t nVerts, const float* refMin, const float* refMax,
                             uint32_t* packets, float* center4, float* boundsMin4, float* boundsMax4) {
  if (!vertices || !packets || nVerts % 32 != 0) return 0;
  const uint32_t nQuads = nVerts / 4;
  // quad normals (Occluder.cpp:12-21)
  std::vector<V3> normals(nQuads);
  for (uint32_t q = 0; q < nQuads; ++q) {
    const float* v = vertices + 16 * size_t(q);
    const V3 a = tri_normal(v, v + 4, v + 8), b = tri_normal(v, v + 8, v + 12);
    normals[q] = normalized({a.x + b.x, a.y + b.y, a.z + b.z});
  }
  // k-means by facing, 6 axis seeds, at most 10 rounds (Occluder.cpp:23-78)
  V3 seeds[6] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, -1, 0}, {0, 0, -1}, {-1, 0, 0}};
  std::vector<uint32_t> cluster(nQuads, 0);
  bool moved = true;
  for (int round = 0; round < 10 && moved; ++round) {
    moved = false;
    for (uint32_t q = 0; q < nQuads; ++q) {
      float best = -INFINITY;
      uint32_t pick = 0;
      for (uint32_t k = 0; k < 6; ++k) {
        const float d = dot3(seeds[k], normals[q]);
        if (d >= best) { best = d; pick = k; }  // _mm_comige_ss: false when unordered
      }
      if (cluster[q] != pick) { cluster[q] = pick; moved = true; }
    }
    for (auto& s : seeds) s = {0, 0, 0};
    for (uint32_t q = 0; q < nQuads; ++q) {
      V3& s = seeds[cluster[q]];
      s = {s.x + normals[q].x, s.y + normals[q].y, s.z + normals[q].z};
    }
    for (auto& s : seeds) s = normalized(s);
  }
  // stable regroup by cluster (Occluder.cpp:80-93), quantise + pack (Occluder.cpp:97-156)
  std::vector<uint32_t> order;
  order.reserve(nQuads);
  for (uint32_t k = 0; k < 6; ++k)
    for (uint32_t q = 0; q < nQuads; ++q)
      if (cluster[q] == k) order.push_back(q);
  float inv[3];
  for (int i = 0; i < 3; ++i) inv[i] = 1.0f / (refMax[i] - refMin[i]);
  const float scale[3] = {2047.0f, 2047.0f, 1023.0f};
  for (uint32_t n = 0; n < nQuads; ++n) {
    const float* quad = vertices + 16 * size_t(order[n]);
    const uint32_t group = n >> 3, lane = n & 7;
    for (uint32_t j = 0; j < 4; ++j) {
      int32_t c[3];
      for (int i = 0; i < 3; ++i) c[i] = cvtt_x86(fmaf((quad[4 * j + i] - refMin[i]) * inv[i], scale[i], 0.5f));
      packets[size_t(group) * 32 + j * 8 + lane] = ((uint32_t(c[0]) - 1024u) << 21) | (uint32_t(c[1]) << 10) | uint32_t(c[2]);
    }
  }
  // bounds over all four lanes, then w := 1 (Occluder.cpp:159-178)
  float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (uint32_t i = 0; i < nVerts; ++i)
    for (int k = 0; k < 4; ++k) {
      const float v = vertices[4 * size_t(i) + k];
      mn[k] = min_x86(v, mn[k]);
      mx[k] = max_x86(v, mx[k]);
    }
  mn[3] = mx[3] = 1.0f;
  for (int k = 0; k < 4; ++k) {
    if (boundsMin4) boundsMin4[k] = mn[k];
    if (boundsMax4) boundsMax4[k] = mx[k];
    if (center4) center4[k] = (mx[k] + mn[k]) * 0.5f;
  }
  return nQuads / 2;  // m_packetCount: 4 packets per 8 quads
}
