// TEST-ONLY: compiles rasterizer_b200/csrc/orz_core.h (the scalar cores the CUDA kernels run per
// lane) for the host so tests can compare every field with the oracle without a GPU.
// Not part of the product; the product library has no CPU path.
#include "../rasterizer_b200/csrc/orz_core.h"

static const uint32_t kNibbles[32] = {ORZ_MODE_NIBBLES};

extern "C" {
void core_bake_view(const float* m, uint32_t w, uint32_t h, float* baked, float* raw) {
  orz::ViewMatrices vm;
  orz::bake_view_matrices(m, w, h, vm);
  memcpy(baked, vm.baked, 64);
  memcpy(raw, vm.raw, 64);
}
// out: 6 uint32 (mode,minX,minY,rangeX,rangeY,maxZ) + 3 float + 12 float + 4 uint32, same order as OrcPrim
int core_setup_quad(const float* m, uint32_t w, uint32_t h, const float* refMin, const float* refMax, const uint32_t* words,
                    int clipped, const uint32_t* rcpTable, int bits, void* outPrim) {
  orz::ViewMatrices vm;
  orz::bake_view_matrices(m, w, h, vm);
  orz::CallMatrix cm;
  orz::prepare_call(vm.baked, refMin, refMax, cm);
  orz::RcpTable rt{rcpTable, 23 - bits};
  orz::Prim P;
  memset(&P, 0, sizeof P);
  bool ok = clipped ? orz::setup_quad<true>(words, cm, rt, kNibbles, int32_t(w / 8), int32_t(h / 8), P)
                    : orz::setup_quad<false>(words, cm, rt, kNibbles, int32_t(w / 8), int32_t(h / 8), P);
  if (!ok) memset(&P, 0, sizeof P);
  memcpy(outPrim, &P, sizeof P);
  return ok ? 1 : 0;
}
void core_box_front(const float* m, uint32_t w, uint32_t h, const float* mn, const float* mx, const uint32_t* rcpTable,
                    int bits, uint32_t* out6) {
  orz::ViewMatrices vm;
  orz::bake_view_matrices(m, w, h, vm);
  orz::RcpTable rt{rcpTable, 23 - bits};
  orz::BoxFront f = orz::box_front_half(vm, mn, mx, w, h, rt);
  out6[0] = f.status; out6[1] = f.minX; out6[2] = f.maxX; out6[3] = f.minY; out6[4] = f.maxY; out6[5] = f.maxZ;
}
}
