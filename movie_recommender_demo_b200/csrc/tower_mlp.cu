// tower_mlp.cu — placeholder until the fused tcgen05 tower lands.
#include "internal.h"
using namespace b2r;
extern "C" {
int b2r_tower_create(b2r_tower** out, const b2r_tower_weights*, int) {
  if (out) *out = nullptr;
  return fail(B2R_EUNSUPPORTED, "tower: not built yet");
}
int b2r_tower_destroy(b2r_tower*) { return B2R_OK; }
size_t b2r_tower_workspace(const b2r_tower*, int64_t) { return 0; }
int b2r_tower_forward(b2r_tower*, const int64_t*, const float*, int64_t, float*, int32_t*, void*, size_t, void*) {
  return fail(B2R_EUNSUPPORTED, "tower: not built yet");
}
}
