// Host-side helpers shared between orz_host.cpp (g++) and orz_kernels.cu (nvcc host pass).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace orz {
// rcpps(1.m) classes of the CPU this runs on; `bits` = leading mantissa bits the result depends on
void probe_host_rcp(std::vector<uint32_t>& table, int& bits, bool& exact);
void probe_host_rsqrt(std::vector<uint32_t>& table, int& bits, bool& exact);
// rsqrtps table Occluder::bake currently uses (installed by orz_set_rsqrt_table, else this CPU's)
void current_rsqrt_table(std::vector<uint32_t>& table, int& bits);
// 4096-entry edge-mask table (Rasterizer.cpp:547-604), built once per process
const int64_t* edge_mask_table();
// rsqrtps as the host-side geometry code sees it: the installed table, else this CPU's instruction
float rsqrt_current(float x);
// error channel of the C ABI (orz_last_error), shared by the host-only translation units
int set_error(int code, const std::string& msg);

// VectorMath.h:6-23 written out per component (products rounded separately, dpps 0x7F sum order)
struct V3 { float x, y, z; };
static inline V3 sub(const float* a, const float* b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
// normal(): cross(v1 - v0, v2 - v0)
static inline V3 tri_normal(const float* v0, const float* v1, const float* v2) {
  const V3 a = sub(v1, v0), b = sub(v2, v0);
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline float dot3(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }  // dpps 0x7F
static inline V3 normalized(const V3& v) {  // VectorMath.h:20-23
  const float s = rsqrt_current(dot3(v, v));
  return {v.x * s, v.y * s, v.z * s};
}
}  // namespace orz
