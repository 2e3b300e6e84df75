// tower.cu — categorical-feature embedding gather/concat and the tower MLP.
//
// gather_concat replaces EmbeddingLayer.forward (two_tower_model.py:42-47): for every
// field f, out[:, f*E:(f+1)*E] = W_f[idx[:, f]].  Concat, not sum (bag size 1), bit-exact
// fp32 copy.  HBM-bound random 64-byte row reads: one warp per sample, every lane owns one
// 16-byte piece of one field row, all of a sample's row reads are issued before any store
// (F*E/4 <= 128 -> up to 4 independent 128-bit loads in flight per lane), stores are
// fully coalesced (a sample's output row is contiguous).
#include "internal.h"
#include "ptx.cuh"

namespace b2r {
namespace {

constexpr int kGatherWarps = 8;

__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// E4 = emb_dim / 4 (float4 pieces per field row); W = F * E4 pieces per sample.
__global__ void __launch_bounds__(kGatherWarps * 32)
gather_concat_kernel(const float* const* __restrict__ tables, const int64_t* __restrict__ cards, int F,
                     int E4, const int64_t* __restrict__ idx, int64_t B, float* __restrict__ out,
                     int64_t ld, int32_t* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const int W = F * E4;
  const int64_t warp0 = (int64_t)blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kGatherWarps;
  for (int64_t b = warp0; b < B; b += nwarps) {
    for (int w0 = 0; w0 < W; w0 += 128) {
      float4 v[4];
      bool bad = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int w = w0 + u * 32 + lane;
        if (w < W) {
          const int f = w / E4, part = w - f * E4;
          int64_t r = __ldg(idx + b * F + f);
          if (r < 0 || r >= __ldg(cards + f)) {
            bad = true;
            r = 0;
          }
          v[u] = ldg_nc_f4(reinterpret_cast<const float4*>(tables[f]) + r * E4 + part);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int w = w0 + u * 32 + lane;
        if (w < W) *reinterpret_cast<float4*>(out + b * ld + (int64_t)w * 4) = v[u];
      }
      if (bad && err_flag) atomicOr(err_flag, 1);   // B2R_TOWER_BAD_INDEX
    }
  }
}

}  // namespace
}  // namespace b2r

using namespace b2r;

extern "C" int b2r_gather_concat(const float* const* tables, const int64_t* cards, int F, int emb_dim,
                                 const int64_t* idx, int64_t B, float* out, int64_t ld,
                                 int32_t* err_flag, void* stream) {
  if (B < 0 || F < 1 || !tables || !cards || (B > 0 && (!idx || !out)))
    return fail(B2R_EINVAL, "gather_concat: bad arguments");
  if (emb_dim < 4 || emb_dim % 4 != 0) return fail(B2R_EINVAL, "gather_concat: emb_dim must be a multiple of 4");
  if (ld < (int64_t)F * emb_dim || ld % 4 != 0)
    return fail(B2R_EINVAL, "gather_concat: ld must be >= F*emb_dim and a multiple of 4");
  if (B == 0) return B2R_OK;
  int64_t blocks = ceil_div(B, kGatherWarps);
  const int64_t max_blocks = 148 * 8 * 4;
  if (blocks > max_blocks) blocks = max_blocks;
  gather_concat_kernel<<<(unsigned)blocks, kGatherWarps * 32, 0, (cudaStream_t)stream>>>(
      tables, cards, F, emb_dim / 4, idx, B, out, ld, err_flag);
  B2R_CHECK_LAUNCH("gather_concat_kernel");
  return B2R_OK;
}

// The fused tower (gather -> 3 tcgen05 GEMMs -> normalise) lands in tower_mlp.cu.
