/* TEST INFRASTRUCTURE ONLY -- scalar C restatement ("port") of the reference's occlusion-culling
 * hot path (SoftwareRasterizer/Rasterizer.cpp, Occluder.cpp).  Used by tests/, smoke() and
 * bench.py's cpu_baseline leg as the checker; the product never links or calls it.
 *
 * Parity status: PINNED -- tests/test_oracle_port.py checks this port bit-for-bit against the
 * unmodified reference sources compiled into oracle/_ref/libref_oracle.so (depth, HiZ, gate
 * decisions, occludee bits, baked words, LUT) and against fixtures generated from that build
 * (tests/golden/).  The reference itself ships no tests or golden vectors.
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcRasterizer OrcRasterizer;

/* rcpps / rsqrtps models.  The reference's results depend on the host's approximate
 * reciprocal instructions (Rasterizer.cpp:218,223,456,717-727,981; VectorMath.h:22).
 * Table form: rcp(x) for x = 2^0 * 1.m is table[m >> (23-bits)]; exponent handling as measured
 * on Intel hosts.  orc_use_host_tables() probes the CPU this runs on. */
void orc_set_rcp_table(const uint32_t* table, int bits);
void orc_set_rsqrt_table(const uint32_t* table, int bits); /* 2 << bits entries: [exponent parity][mantissa] */
void orc_use_host_tables(void);
float orc_rcp(float x);
float orc_rsqrt(float x);
void orc_probe_host_rcp(uint32_t* table, int bits);   /* rcpps(1.m) for the 2^bits leading mantissas */
void orc_probe_host_rsqrt(uint32_t* table, int bits); /* rsqrtps over [1,4) */

/* Rasterizer.cpp:547-604 -- 64 slopes x 64 offsets table of 8x8 coverage masks */
void orc_build_lut(int64_t* lut4096);

OrcRasterizer* orc_create(uint32_t width, uint32_t height, const int64_t* lut4096 /* NULL: build */);
void orc_destroy(OrcRasterizer* r);
void orc_set_mvp(OrcRasterizer* r, const float* m16);                 /* Rasterizer.cpp:76-105 */
void orc_clear(OrcRasterizer* r);                                     /* Rasterizer.cpp:107-121, + depth := 0 (fresh) */
/* packets: reference layout (Occluder.cpp:146-156), packetCount x 8 uint32 */
void orc_rasterize(OrcRasterizer* r, const uint32_t* packets, uint32_t packetCount, const float* refMin4,
                   const float* refMax4, int possiblyNearClipped);    /* Rasterizer.cpp:606-1295 */
int orc_query_visibility(OrcRasterizer* r, const float* bmin4, const float* bmax4); /* bit0 visible, bit1 needsClipping; :123-281 */
int orc_query2d(const OrcRasterizer* r, uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ); /* :283-349 */
void orc_readback_depth(const OrcRasterizer* r, uint8_t* bgra);       /* Rasterizer.cpp:351-399 */
const uint16_t* orc_depth(const OrcRasterizer* r);
const uint16_t* orc_hiz(const OrcRasterizer* r);
const int64_t* orc_lut(const OrcRasterizer* r);
void orc_get_matrices(const OrcRasterizer* r, float* baked16, float* raw16);

/* Occluder.cpp:7-181.  vertices: nVerts x float4 (4 per quad, nVerts % 32 == 0).
 * Outputs: packets (nVerts uint32, reference layout), center/boundsMin/boundsMax (4 floats each).
 * Returns the packet count. */
uint32_t orc_bake(const float* vertices, uint32_t nVerts, const float* refMin4, const float* refMax4,
                  uint32_t* packets, float* center4, float* bmin4, float* bmax4);

/* Per-quad setup record, exposed for field-by-field parity tests of the CUDA setup kernel. */
typedef struct {
  uint32_t mode;            /* 0 = culled / outside */
  int32_t minX, minY, rangeX, rangeY;
  uint32_t maxZ;            /* 16-bit */
  float dzdx, dzdy, plane0;
  float nx[4], ny[4], off[4];
  uint32_t slope[4];        /* already << 6 */
} OrcPrim;
/* planning statistics of the block loop: visits, HiZ-rejected, updates, updates without any change, updates a
   per-block corner bound would have skipped; reset != 0 clears them */
void orc_stats(uint64_t* out5, int reset);
void orc_set_block_bound_skip(int on); /* planning switch: skip updates the per-block corner bound proves to be no-ops */
void orc_setup_quad(const OrcRasterizer* r, const uint32_t word[4], const float* refMin4, const float* refMax4,
                    int possiblyNearClipped, OrcPrim* out);

#ifdef __cplusplus
}
#endif
#endif
