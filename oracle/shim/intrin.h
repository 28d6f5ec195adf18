/* MSVC <intrin.h> stand-in so the UNMODIFIED reference sources under
 * /root/reference/SoftwareRasterizer compile with g++ (test infrastructure only;
 * see oracle/README.md).  Occluder.h:5 includes <intrin.h>; Rasterizer.cpp uses
 * __forceinline / __debugbreak / _BitScanForward and Occluder.cpp uses
 * _aligned_malloc -- all MSVC-isms with direct GNU equivalents. */
#pragma once
#include <immintrin.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>

#ifndef __forceinline
#define __forceinline inline __attribute__((always_inline))
#endif

static inline void __debugbreak() { __builtin_trap(); }

static inline void* _aligned_malloc(size_t size, size_t alignment)
{
  size_t rounded = (size + alignment - 1) / alignment * alignment;
  if (rounded == 0) rounded = alignment;
  return aligned_alloc(alignment, rounded);
}

static inline unsigned char _BitScanForward(unsigned long* index, unsigned long mask)
{
  if (mask == 0) return 0;
  *index = (unsigned long)__builtin_ctzl(mask);
  return 1;
}
