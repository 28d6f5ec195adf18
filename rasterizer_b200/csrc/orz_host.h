// Host-side helpers shared between orz_host.cpp (g++) and orz_kernels.cu (nvcc host pass).
#pragma once
#include <stdint.h>

#include <vector>

namespace orz {
// rcpps(1.m) classes of the CPU this runs on; `bits` = leading mantissa bits the result depends on
void probe_host_rcp(std::vector<uint32_t>& table, int& bits, bool& exact);
void probe_host_rsqrt(std::vector<uint32_t>& table, int& bits, bool& exact);
// rsqrtps table Occluder::bake currently uses (installed by orz_set_rsqrt_table, else this CPU's)
void current_rsqrt_table(std::vector<uint32_t>& table, int& bits);
// 4096-entry edge-mask table (Rasterizer.cpp:547-604), built once per process
const int64_t* edge_mask_table();
}  // namespace orz
