// Occluder::bake on the GPU (Occluder.cpp:7-181).
// Part of the single translation unit orz_kernels.cu (included inside namespace orz); see DESIGN.md section 4.
#pragma once

// ---------------------------------------------------------------------------------------------
// Occluder::bake (Occluder.cpp:7-181) on the GPU, one CTA per batch: quad normals -> k-means by
// facing (6 axis seeds, at most 10 rounds) -> stable regroup by cluster -> 11/11/10 quantisation
// -> one uint4 per quad (and, when asked, the reference's packet layout) -> bounds and centre.
// Bit-exact with the host bake: every float sum runs in the reference's order (the cluster sums
// are accumulated quad by quad by one thread per (cluster, component)), rsqrtps through its
// table model (rsqrt_x86), products rounded separately (-fmad=false).
struct BakeJob {
  uint32_t vertOffset;  // first vertex (float4) of the batch
  uint32_t nQuads;
  uint32_t quadOffset;  // first output quad
  uint32_t pad;
};

__device__ __forceinline__ void bake_normal(const float4 v0, const float4 v1, const float4 v2, float& x, float& y, float& z) {
  // normal() of VectorMath.h:6-18: cross(v1 - v0, v2 - v0)
  const float ax = v1.x - v0.x, ay = v1.y - v0.y, az = v1.z - v0.z, bx = v2.x - v0.x, by = v2.y - v0.y, bz = v2.z - v0.z;
  x = ay * bz - az * by; y = az * bx - ax * bz; z = ax * by - ay * bx;
}

__global__ void __launch_bounds__(256) k_bake(const float4* __restrict__ verts, const BakeJob* __restrict__ jobs, const float4 refMin,
                                               const float4 refMax, const RsqrtTable rs, uint4* __restrict__ outQuads, OccMeta* __restrict__ meta,
                                               uint32_t* __restrict__ outPackets) {
  extern __shared__ __align__(16) float s_bake[];
  __shared__ float s_seed[6][3], s_sum[6][3];
  __shared__ float s_mn[8][4], s_mx[8][4];
  __shared__ uint32_t s_mnI[8][4], s_mxI[8][4];
  const BakeJob job = jobs[blockIdx.x];
  const uint32_t n = job.nQuads, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  float* nx = s_bake; float* ny = nx + n; float* nz = ny + n;
  uint32_t* cl = reinterpret_cast<uint32_t*>(nz + n);
  uint32_t* pos = cl + n;
  const float4* v = verts + job.vertOffset;

  // quad normals (Occluder.cpp:12-21) and the bounds over all four lanes (Occluder.cpp:159-170).
  // minps / maxps keep the EARLIER vertex when two compare equal (+0 / -0), so the reduction
  // carries the vertex index and breaks ties towards the lower one: same result as the serial loop.
  float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  uint32_t mnI[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, mxI[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
  for (uint32_t q = tid; q < n; q += 256u) {
    const float4 v0 = v[4 * q], v1 = v[4 * q + 1], v2 = v[4 * q + 2], v3 = v[4 * q + 3];
    float ax, ay, az, bx, by, bz;
    bake_normal(v0, v1, v2, ax, ay, az);
    bake_normal(v0, v2, v3, bx, by, bz);
    const float sx = ax + bx, sy = ay + by, sz = az + bz;
    const float r = rsqrt_x86((sx * sx + sy * sy) + sz * sz, rs);  // normalize(), VectorMath.h:20-23; dpps 0x7F sum order
    nx[q] = sx * r; ny[q] = sy * r; nz[q] = sz * r;
    cl[q] = 0u;
    const float vv[4][4] = {{v0.x, v0.y, v0.z, v0.w}, {v1.x, v1.y, v1.z, v1.w}, {v2.x, v2.y, v2.z, v2.w}, {v3.x, v3.y, v3.z, v3.w}};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (vv[j][k] < mn[k]) { mn[k] = vv[j][k]; mnI[k] = 4u * q + (uint32_t)j; }
        if (vv[j][k] > mx[k]) { mx[k] = vv[j][k]; mxI[k] = 4u * q + (uint32_t)j; }
      }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const float om = __shfl_xor_sync(kFull, mn[k], d), oM = __shfl_xor_sync(kFull, mx[k], d);
      const uint32_t omI = __shfl_xor_sync(kFull, mnI[k], d), oMI = __shfl_xor_sync(kFull, mxI[k], d);
      if (om < mn[k] || (om == mn[k] && omI < mnI[k])) { mn[k] = om; mnI[k] = omI; }
      if (oM > mx[k] || (oM == mx[k] && oMI < mxI[k])) { mx[k] = oM; mxI[k] = oMI; }
    }
    if (lane == 0) { s_mn[warp][k] = mn[k]; s_mx[warp][k] = mx[k]; s_mnI[warp][k] = mnI[k]; s_mxI[warp][k] = mxI[k]; }
  }
  if (tid < 18) s_seed[tid / 3][tid % 3] = 0.0f;
  __syncthreads();
  if (tid == 0) { s_seed[0][0] = 1.0f; s_seed[1][1] = 1.0f; s_seed[2][2] = 1.0f; s_seed[3][1] = -1.0f; s_seed[4][2] = -1.0f; s_seed[5][0] = -1.0f; }
  __syncthreads();

  // k-means by facing (Occluder.cpp:23-78)
  for (int round = 0; round < 10; ++round) {
    int moved = 0;
    for (uint32_t q = tid; q < n; q += 256u) {
      float best = -INFINITY;
      uint32_t pick = 0;
#pragma unroll
      for (uint32_t k = 0; k < 6; ++k) {
        const float d = (s_seed[k][0] * nx[q] + s_seed[k][1] * ny[q]) + s_seed[k][2] * nz[q];
        if (d >= best) { best = d; pick = k; }  // _mm_comige_ss: false when unordered
      }
      if (cl[q] != pick) { cl[q] = pick; moved = 1; }
    }
    if (!__syncthreads_or(moved)) break;  // the seeds are not used after the last round
    if (tid < 18) {  // cluster sums in quad order, one thread per (cluster, component)
      const uint32_t k = tid / 3u;
      const float* comp = tid % 3u == 0 ? nx : (tid % 3u == 1 ? ny : nz);
      float acc = 0.0f;
      for (uint32_t q = 0; q < n; ++q)
        if (cl[q] == k) acc = acc + comp[q];
      s_sum[k][tid % 3u] = acc;
    }
    __syncthreads();
    if (tid < 6) {
      const float x = s_sum[tid][0], y = s_sum[tid][1], z = s_sum[tid][2];
      const float r = rsqrt_x86((x * x + y * y) + z * z, rs);
      s_seed[tid][0] = x * r; s_seed[tid][1] = y * r; s_seed[tid][2] = z * r;
    }
    __syncthreads();
  }

  // stable regroup by cluster (Occluder.cpp:80-93): slot of quad q = quads of lower clusters + earlier quads of its own
  if (warp == 0) {
    uint32_t count[6] = {0, 0, 0, 0, 0, 0};
    for (uint32_t q0 = 0; q0 < n; q0 += 32u) {
      const uint32_t c = q0 + lane < n ? cl[q0 + lane] : 7u;
#pragma unroll
      for (uint32_t k = 0; k < 6; ++k) {
        const uint32_t m = __ballot_sync(kFull, c == k);
        if (c == k) pos[q0 + lane] = count[k] + (uint32_t)__popc(m & ((1u << lane) - 1u));
        count[k] += (uint32_t)__popc(m);
      }
    }
    uint32_t base[6];
    base[0] = 0;
#pragma unroll
    for (int k = 1; k < 6; ++k) base[k] = base[k - 1] + count[k - 1];
    for (uint32_t q = lane; q < n; q += 32u) {
      const uint32_t c = cl[q];
#pragma unroll
      for (uint32_t k = 0; k < 6; ++k) if (c == k) pos[q] += base[k];
    }
  }
  __syncthreads();

  // quantise and pack (Occluder.cpp:97-156): word = (X - 1024) << 21 | Y << 10 | Z
  const float ivx = 1.0f / (refMax.x - refMin.x), ivy = 1.0f / (refMax.y - refMin.y), ivz = 1.0f / (refMax.z - refMin.z);
  for (uint32_t q = tid; q < n; q += 256u) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 p = v[4 * q + j];
      const uint32_t cx = (uint32_t)cvtt_x86(ORZ_FMA((p.x - refMin.x) * ivx, 2047.0f, 0.5f));
      const uint32_t cy = (uint32_t)cvtt_x86(ORZ_FMA((p.y - refMin.y) * ivy, 2047.0f, 0.5f));
      const uint32_t cz = (uint32_t)cvtt_x86(ORZ_FMA((p.z - refMin.z) * ivz, 1023.0f, 0.5f));
      w[j] = ((cx - 1024u) << 21) | (cy << 10) | cz;
    }
    const uint32_t at = pos[q];
    outQuads[job.quadOffset + at] = make_uint4(w[0], w[1], w[2], w[3]);
    if (outPackets) {  // the reference's own layout: group of 8 quads = 4 x 8 words
      uint32_t* pk = outPackets + (size_t)job.quadOffset * 4u + (size_t)(at >> 3) * 32u + (at & 7u);
      pk[0] = w[0]; pk[8] = w[1]; pk[16] = w[2]; pk[24] = w[3];
    }
  }
  if (tid < 4 && meta) {  // bounds, w := 1 (Occluder.cpp:172-173), centre
    float lo = s_mn[0][tid], hi = s_mx[0][tid];
    uint32_t loI = s_mnI[0][tid], hiI = s_mxI[0][tid];
    for (int w2 = 1; w2 < 8; ++w2) {
      if (s_mn[w2][tid] < lo || (s_mn[w2][tid] == lo && s_mnI[w2][tid] < loI)) { lo = s_mn[w2][tid]; loI = s_mnI[w2][tid]; }
      if (s_mx[w2][tid] > hi || (s_mx[w2][tid] == hi && s_mxI[w2][tid] < hiI)) { hi = s_mx[w2][tid]; hiI = s_mxI[w2][tid]; }
    }
    if (tid == 3) { lo = 1.0f; hi = 1.0f; }
    OccMeta& om = meta[blockIdx.x];
    om.boundsMin[tid] = lo; om.boundsMax[tid] = hi; om.center[tid] = (hi + lo) * 0.5f;
    om.refMin[tid] = tid == 0 ? refMin.x : tid == 1 ? refMin.y : tid == 2 ? refMin.z : refMin.w;
    om.refMax[tid] = tid == 0 ? refMax.x : tid == 1 ? refMax.y : tid == 2 ? refMax.z : refMax.w;
    if (tid == 0) { om.quadOffset = job.quadOffset; om.quadCount = n; om.pad0 = om.pad1 = 0u; }
  }
}
