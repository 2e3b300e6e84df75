// ingest.cu — row normalisation + fp32/bf16 materialisation for corpus rows and queries.
//
// Replaces `embeddings.astype('float32'); faiss.normalize_L2(embeddings)` at
// faiss_retrieval.py:114-115 (corpus) and :146-147 (queries): per row,
// if sum(x^2) > 0 then x *= 1/sqrt(sum(x^2)); zero rows stay zero.  The input is never
// modified (the reference normalises a copy).  HBM-bound elementwise work: one warp per
// row, 128-bit loads/stores, fp32 master + bf16 scan copy written in the same pass.
#include <cuda_fp16.h>

#include "internal.h"

namespace b2r {
namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// d % 4 == 0, d <= 1024.  rows_out >= rows_in: rows in [rows_in, rows_out) are zero-filled.
__device__ __forceinline__ uint32_t pack16x2(float a, float b, int fp16) {
  if (fp16) {
    uint32_t r;
    asm("{\n\t.reg .b16 l, h;\n\tcvt.rn.satfinite.f16.f32 l, %1;\n\tcvt.rn.satfinite.f16.f32 h, %2;\n\t"
        "mov.b32 %0, {l, h};\n\t}"
        : "=r"(r)
        : "f"(a), "f"(b));
    return r;
  }
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// scan16_unit: the 16-bit copy is always L2-normalised (queries), independent of `normalize`
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normalize_rows_kernel(const float* __restrict__ x, int64_t rows_in, int64_t rows_out, int d,
                      int normalize, float* __restrict__ out32, __nv_bfloat16* __restrict__ out16,
                      int fp16, int scan16_unit, float* __restrict__ norms, float* __restrict__ maxnorm,
                      float* __restrict__ inf_fill) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows_out) return;
  float4 v[8];
  const int nvec = d >> 2;  // float4 per row
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nvec && row < rows_in) {
      v[i] = __ldg(reinterpret_cast<const float4*>(x + row * d) + c);
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  ss = warp_sum(ss);
  float scale = 1.f;
  float stored_norm = sqrtf(ss);
  const float inv = ss > 0.f ? 1.0f / sqrtf(ss) : 1.0f;
  if (normalize) {
    if (ss > 0.f) {
      scale = inv;
      stored_norm = 1.0f;
    }
  }
  const float scale16 = scan16_unit ? inv : scale;
  if (scan16_unit) stored_norm = ss > 0.f ? 1.0f : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float4 o = make_float4(v[i].x * scale, v[i].y * scale, v[i].z * scale, v[i].w * scale);
      if (out32) reinterpret_cast<float4*>(out32 + row * d)[c] = o;
      if (out16) {
        uint2 pk;
        pk.x = pack16x2(v[i].x * scale16, v[i].y * scale16, fp16);
        pk.y = pack16x2(v[i].z * scale16, v[i].w * scale16, fp16);
        reinterpret_cast<uint2*>(out16 + row * d)[c] = pk;
      }
    }
  }
  if (lane == 0) {
    if (inf_fill) inf_fill[row] = INFINITY;  // per-query threshold starts at +inf (padding rows keep it)
    if (norms) norms[row] = (row < rows_in) ? stored_norm : 0.f;
    if (maxnorm && row < rows_in && stored_norm > 0.f && stored_norm < INFINITY)
      atomicMax(reinterpret_cast<int*>(maxnorm), __float_as_int(stored_norm));
  }
}

__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
__global__ void fill_i32_kernel(int* p, int64_t n, int v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

}  // namespace

uint32_t scan_idesc(int fp16) {
  // kind::f16 descriptor: c_format F32 (bit 4), a/b format BF16 = 1 (bits 7, 10) or F16 = 0, N>>3 at 17, M>>4 at 24
  const uint32_t fmt = fp16 ? 0u : ((1u << 7) | (1u << 10));
  return (1u << 4) | fmt | ((128u >> 3) << 17) | ((128u >> 4) << 24);
}
uint32_t scan_idesc_pair(int fp16) {
  // the same descriptor with M = 256 (both CTAs of a pair, 128 rows each) and N = 256
  const uint32_t fmt = fp16 ? 0u : ((1u << 7) | (1u << 10));
  return (1u << 4) | fmt | ((256u >> 3) << 17) | ((256u >> 4) << 24);
}

int launch_reencode(const float* x32, int64_t n, int d, __nv_bfloat16* out16, int fp16, cudaStream_t stream) {
  if (n <= 0) return B2R_OK;
  const int64_t blocks = ceil_div(n, kWarpsPerBlock);
  normalize_rows_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, stream>>>(x32, n, n, d, 0, nullptr, out16, fp16, 0,
                                                                             nullptr, nullptr, nullptr);
  B2R_CHECK_LAUNCH("normalize_rows_kernel(reencode)");
  return B2R_OK;
}

int launch_ingest(const float* x, int64_t n, int d, int normalize, float* out32,
                  __nv_bfloat16* out16, int fp16, float* maxnorm, cudaStream_t stream) {
  if (n <= 0) return B2R_OK;
  if (d % 4 != 0 || d > 1024) return fail(B2R_EINVAL, "ingest: d must be a multiple of 4, <= 1024");
  const int64_t blocks = ceil_div(n, kWarpsPerBlock);
  normalize_rows_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, stream>>>(
      x, n, n, d, normalize, out32, out16, fp16, 0, nullptr, maxnorm, nullptr);
  B2R_CHECK_LAUNCH("normalize_rows_kernel");
  return B2R_OK;
}

int launch_prep_queries(const float* x, int q, int qpad, int d, int normalize, float* q32,
                        __nv_bfloat16* q16, int fp16, float* qnorm, cudaStream_t stream, float* tau_init) {
  if (qpad <= 0) return B2R_OK;
  if (d % 4 != 0 || d > 1024) return fail(B2R_EINVAL, "queries: d must be a multiple of 4, <= 1024");
  const int64_t blocks = ceil_div(qpad, kWarpsPerBlock);
  normalize_rows_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, stream>>>(
      x, q, qpad, d, normalize, q32, q16, fp16, 1, qnorm, nullptr, tau_init);
  B2R_CHECK_LAUNCH("normalize_rows_kernel(queries)");
  return B2R_OK;
}

int launch_fill_f32(float* p, int64_t n, float v, cudaStream_t stream) {
  if (n <= 0) return B2R_OK;
  const int blocks = (int)(ceil_div(n, 256) < 1184 ? ceil_div(n, 256) : 1184);
  fill_f32_kernel<<<blocks, 256, 0, stream>>>(p, n, v);
  B2R_CHECK_LAUNCH("fill_f32_kernel");
  return B2R_OK;
}
int launch_fill_i32(int* p, int64_t n, int v, cudaStream_t stream) {
  if (n <= 0) return B2R_OK;
  const int blocks = (int)(ceil_div(n, 256) < 1184 ? ceil_div(n, 256) : 1184);
  fill_i32_kernel<<<blocks, 256, 0, stream>>>(p, n, v);
  B2R_CHECK_LAUNCH("fill_i32_kernel");
  return B2R_OK;
}

}  // namespace b2r
