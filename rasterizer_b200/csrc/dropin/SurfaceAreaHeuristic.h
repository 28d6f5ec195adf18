// Drop-in replacement for the reference's SoftwareRasterizer/SurfaceAreaHeuristic.h
// (SurfaceAreaHeuristic.h:7-11): same class and signature.  Runs on the GPU of the calling thread's
// context (orz_generate_batches_device); set ORZ_PREP_ON_HOST=1 to use the host implementation
// (orz_generate_batches) -- both return the reference's batches in the reference's order.
#pragma once

#include <cstdint>
#include <vector>

struct Aabb;

class SurfaceAreaHeuristic
{
public:
	static std::vector<std::vector<uint32_t>> generateBatches(const std::vector<Aabb>& aabbs, uint32_t targetSize, uint32_t splitGranularity);
};
