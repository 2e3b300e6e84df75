// ranker.cu — Stage-2 TransformerRanker forward in eval mode.
//
// Replaces transformer_ranker.py:332-380 as the reference calls it on the 500 Stage-1 candidates of a user
// (inference.py:250-255, faiss_retrieval.py:351-355):
//
//   features = [user_emb(6 x 32) || ad_emb(20 x 32) || numerical(13)]            :312-330 (embed_features)
//   x = feature_projection(features) + positional_encoding[:, 0]                  :352-361
//   3 x TransformerEncoderLayer on a sequence of length ONE                       :137-155
//        softmax over a single key is 1  =>  attention(x) = W_o (W_v x + b_v) + b_o  (host-folded: one linear map)
//        x = LayerNorm(x + attention(x));  x = LayerNorm(x + fc2(relu(fc1 x)))
//   3 x cross layer  xl = x0 * (xl W_l + b_l) + xl                                :195-207
//   3 heads  Linear(256) ReLU Linear(64) ReLU Linear(1)                           :283-310, :372-376
//
// Every Linear is a dense contraction -> the TMA-fed tcgen05 GEMM of tower_mlp.cu (16-bit operands, fp32
// accumulation in TMEM, bias / ReLU applied straight out of TMEM).  The residual stream, the LayerNorm statistics
// and the cross-layer products stay fp32 (small row-wise kernels below, a warp per row); only the GEMM operands
// are 16-bit.  The three heads run as ONE stacked GEMM (d_model -> 3 x 256) and ONE block-diagonal GEMM
// (3 x 256 -> 3 x 64); the last Linear(64, 1) of each head is a 64-term dot product per row.
// Dropout is the identity in eval mode.
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "tower_internal.h"

struct b2r_ranker {
  int device = 0, num_sms = 148;
  int Fu = 0, Fa = 0, E = 0, nnum = 0;
  int K1 = 0, K1p = 0;        // projection fan-in and its 64-padding
  int dm = 0, dff = 0, L = 0, C = 0, T = 0, h1 = 0, h2 = 0;
  int dmp = 0, dffp = 0;      // d_model / d_ff padded to 128
  int H1 = 0, H1p = 0;        // stacked head layer 1: T * h1, padded to 128
  int H2 = 0, H2p = 0;        // stacked head layer 2: T * h2, padded to 128
  int bf16 = 0;
  const float** tables = nullptr;
  int64_t* cards = nullptr;
  struct Lin {                // one Linear: 16-bit weights [Np, Kp] (both formats), fp32 bias [Np], tensor maps
    int N = 0, K = 0;         // padded
    __half* w = nullptr;
    __nv_bfloat16* wb = nullptr;
    float* b = nullptr;
    CUtensorMap tm, tmb;
  };
  std::vector<Lin> lin;       // proj | per layer: attn, fc1, fc2 | cross[C] | heads1 | heads2
  float* ln = nullptr;        // [L][4][dm]: ln1_g, ln1_b, ln2_g, ln2_b
  float* w3 = nullptr;        // [T, h2] + [T]
};

namespace b2r {
namespace {

__device__ __forceinline__ float4 ldg_nc4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
template <bool BF16>
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
  uint32_t r;
  if (BF16) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    asm("{\n\t.reg .b16 l, h;\n\t"
        "cvt.rn.satfinite.f16.f32 l, %1;\n\t"
        "cvt.rn.satfinite.f16.f32 h, %2;\n\t"
        "mov.b32 %0, {l, h};\n\t}"
        : "=r"(r)
        : "f"(lo), "f"(hi));
  }
  return r;
}
template <bool BF16>
__device__ __forceinline__ uint2 pk4(float4 v) {
  return make_uint2(pk2<BF16>(v.x, v.y), pk2<BF16>(v.z, v.w));
}
__device__ __forceinline__ float amax4(float4 v) {
  return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kRkWarps = 8;

// features of one candidate row -> 16-bit projection operand [B, K1p]: user tables first, then ad tables (the
// ModuleDict orders of embed_features, transformer_ranker.py:318-327), then the numericals, zero padding behind.
template <bool BF16>
__global__ void __launch_bounds__(kRkWarps * 32)
ranker_gather_kernel(const float* const* __restrict__ tables, const int64_t* __restrict__ cards, int Fu, int Fa,
                     int E4, const int64_t* __restrict__ ucat, const int64_t* __restrict__ acat,
                     const float* __restrict__ num, int nnum, int64_t B, __half* __restrict__ out, int K1p,
                     int32_t* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const int W = (Fu + Fa) * E4;
  const int tail0 = W * 4;
  float amax = 0.f;
  bool bad = false;
  for (int64_t b = (int64_t)blockIdx.x * kRkWarps + (threadIdx.x >> 5); b < B; b += (int64_t)gridDim.x * kRkWarps) {
    __half* orow = out + b * K1p;
    for (int w = lane; w < W; w += 32) {
      const int f = w / E4, part = w - f * E4;
      int64_t r = f < Fu ? __ldg(ucat + b * Fu + f) : __ldg(acat + b * Fa + (f - Fu));
      if (r < 0 || r >= __ldg(cards + f)) {   // torch raises IndexError here: flagged, row 0 read instead
        bad = true;
        r = 0;
      }
      const float4 v = ldg_nc4(reinterpret_cast<const float4*>(tables[f]) + r * E4 + part);
      if (!BF16) amax = fmaxf(amax, amax4(v));
      *reinterpret_cast<uint2*>(orow + (int64_t)w * 4) = pk4<BF16>(v);
    }
    for (int c = tail0 + lane; c < K1p; c += 32) {
      const int j = c - tail0;
      const float nv = j < nnum ? num[b * nnum + j] : 0.f;
      if (!BF16) amax = fmaxf(amax, fabsf(nv));
      orow[c] = __ushort_as_half((unsigned short)(pk2<BF16>(nv, 0.f) & 0xFFFFu));
    }
  }
  int flags = bad ? kTowerErrIndex : 0;
  if (!BF16 && !(amax <= 65504.f)) flags |= kTowerErrSaturate;
  if (flags && err_flag) atomicOr(err_flag, flags);
}

// Row-wise update of the fp32 residual stream + its 16-bit copy for the next GEMM (a warp per row, d % 128 == 0):
//   op 0 :  x = g                                  (after the projection)
//   op 1 :  x = LayerNorm(x + g) * gamma + beta    (nn.LayerNorm: biased variance, eps 1e-5; :147, :151)
//   op 2 :  xl = x0 * g + xl                       (cross layer, g = xl W + b; :204); x0 is read-only
// g: fp32 GEMM output [B, ldg]; x (xl) fp32 [B, d] updated in place; x16 [B, ld16] 16-bit, padding untouched
template <bool BF16, int NV>   // NV = d / 128 float4 per lane
__global__ void __launch_bounds__(kRkWarps * 32)
ranker_rowop_kernel(int op, int64_t B, int d, const float* __restrict__ g, int ldg, float* __restrict__ x,
                    const float* __restrict__ x0, const float* __restrict__ gamma, const float* __restrict__ beta,
                    __half* __restrict__ x16, int ld16, int32_t* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  float amax = 0.f;
  for (int64_t b = (int64_t)blockIdx.x * kRkWarps + (threadIdx.x >> 5); b < B; b += (int64_t)gridDim.x * kRkWarps) {
    float4 v[NV];
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int c4 = u * 32 + lane;
      const float4 gv = *reinterpret_cast<const float4*>(g + b * ldg + c4 * 4);
      if (op == 0) {
        v[u] = gv;
      } else if (op == 1) {
        const float4 xv = *reinterpret_cast<const float4*>(x + b * d + c4 * 4);
        v[u] = make_float4(xv.x + gv.x, xv.y + gv.y, xv.z + gv.z, xv.w + gv.w);
      } else {
        const float4 xv = *reinterpret_cast<const float4*>(x + b * d + c4 * 4);
        const float4 zv = *reinterpret_cast<const float4*>(x0 + b * d + c4 * 4);
        v[u] = make_float4(fmaf(zv.x, gv.x, xv.x), fmaf(zv.y, gv.y, xv.y), fmaf(zv.z, gv.z, xv.z), fmaf(zv.w, gv.w, xv.w));
      }
    }
    if (op == 1) {
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < NV; ++u) s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
      const float mean = warp_sum(s) / (float)d;
      float q = 0.f;
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const float a = v[u].x - mean, bb = v[u].y - mean, c = v[u].z - mean, e = v[u].w - mean;
        q += (a * a + bb * bb) + (c * c + e * e);
      }
      const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int c4 = u * 32 + lane;
        const float4 ga = *reinterpret_cast<const float4*>(gamma + c4 * 4);
        const float4 be = *reinterpret_cast<const float4*>(beta + c4 * 4);
        v[u] = make_float4((v[u].x - mean) * rstd * ga.x + be.x, (v[u].y - mean) * rstd * ga.y + be.y,
                           (v[u].z - mean) * rstd * ga.z + be.z, (v[u].w - mean) * rstd * ga.w + be.w);
      }
    }
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int c4 = u * 32 + lane;
      *reinterpret_cast<float4*>(x + b * d + c4 * 4) = v[u];
      if (!BF16) amax = fmaxf(amax, amax4(v[u]));
      *reinterpret_cast<uint2*>(x16 + b * ld16 + c4 * 4) = pk4<BF16>(v[u]);
    }
  }
  if (!BF16 && !(amax <= 65504.f) && err_flag) atomicOr(err_flag, kTowerErrSaturate);
}

// out[t, b] = h[b, t*h2 .. (t+1)*h2) . w3[t] + b3[t]   (the last Linear(64, 1) of each head; :305-309)
template <bool BF16>
__global__ void ranker_head_out_kernel(const __half* __restrict__ h, int ldh, int64_t B, int T, int h2,
                                        const float* __restrict__ w3, const float* __restrict__ b3,
                                        float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int64_t b = (int64_t)blockIdx.x * kRkWarps + (threadIdx.x >> 5); b < B; b += (int64_t)gridDim.x * kRkWarps) {
    for (int t = 0; t < T; ++t) {
      float acc = 0.f;
      for (int j = lane; j < h2; j += 32) {
        const unsigned short raw = __half_as_ushort(h[b * ldh + t * h2 + j]);
        const float a = BF16 ? __uint_as_float((uint32_t)raw << 16) : __half2float(__ushort_as_half(raw));
        acc = fmaf(a, w3[t * h2 + j], acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) out[(int64_t)t * B + b] = acc + b3[t];
    }
  }
}

uint16_t to_f16_sat(float f) {
  if (f > 65504.f) f = 65504.f;
  if (f < -65504.f) f = -65504.f;
  const __half hv = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &hv, 2);
  return u;
}
uint16_t to_bf16(float f) {
  const __nv_bfloat16 hv = __float2bfloat16_rn(f);
  uint16_t u;
  memcpy(&u, &hv, 2);
  return u;
}
int pad_up(int v, int m) { return (v + m - 1) / m * m; }

// builds one Linear from host fp32 [n, k] (+ bias [n]); kp = row stride of the activation buffer it will read
int make_lin(b2r_ranker::Lin& l, const std::vector<float>& W, const std::vector<float>& bias, int n, int k, int kp,
             float* wmax) {
  l.N = pad_up(n, 128);
  l.K = kp;
  std::vector<uint16_t> wh((size_t)l.N * l.K, 0), wb((size_t)l.N * l.K, 0);
  std::vector<float> bb((size_t)l.N, 0.f);
  for (int o = 0; o < n; ++o) {
    for (int i = 0; i < k; ++i) {
      const float v = W[(size_t)o * k + i];
      if (!(fabsf(v) <= *wmax)) *wmax = fabsf(v);
      wh[(size_t)o * l.K + i] = to_f16_sat(v);
      wb[(size_t)o * l.K + i] = to_bf16(v);
    }
    bb[o] = bias[o];
  }
  if (cudaMalloc(&l.w, wh.size() * 2) != cudaSuccess || cudaMalloc(&l.wb, wb.size() * 2) != cudaSuccess ||
      cudaMalloc(&l.b, bb.size() * 4) != cudaSuccess)
    return fail(B2R_ENOMEM, "ranker_create: cudaMalloc failed");
  cudaMemcpy(l.w, wh.data(), wh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(l.wb, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(l.b, bb.data(), bb.size() * 4, cudaMemcpyHostToDevice);
  const int box = (l.N % 256 == 0) ? 256 : 128;
  int rc = make_tmap_f16_2d(&l.tm, l.w, l.N, l.K, box);
  if (!rc) rc = make_tmap_f16_2d(&l.tmb, l.wb, l.N, l.K, box);
  return rc;
}

template <bool BF16>
int launch_rowop(int op, int64_t B, int d, const float* g, int ldg, float* x, const float* x0, const float* gamma,
                 const float* beta, __half* x16, int ld16, int32_t* err_flag, int num_sms, cudaStream_t stream) {
  int64_t blocks = ceil_div(B, kRkWarps);
  if (blocks > (int64_t)num_sms * 8) blocks = (int64_t)num_sms * 8;
  switch (d / 128) {
#define B2R_ROWOP(NV_)                                                                                          \
  case NV_:                                                                                                     \
    ranker_rowop_kernel<BF16, NV_><<<(unsigned)blocks, kRkWarps * 32, 0, stream>>>(op, B, d, g, ldg, x, x0, gamma, \
                                                                                  beta, x16, ld16, err_flag);     \
    break;
    B2R_ROWOP(1) B2R_ROWOP(2) B2R_ROWOP(3) B2R_ROWOP(4) B2R_ROWOP(6) B2R_ROWOP(8)
#undef B2R_ROWOP
    default:
      return fail(B2R_EUNSUPPORTED, "ranker: d_model must be 128, 256, 384, 512, 768 or 1024");
  }
  B2R_CHECK_LAUNCH("ranker_rowop_kernel");
  return B2R_OK;
}

}  // namespace
}  // namespace b2r

using namespace b2r;

extern "C" {

int b2r_ranker_destroy(b2r_ranker* r) {
  if (!r) return B2R_OK;
  for (auto& l : r->lin) {
    cudaFree(l.w);
    cudaFree(l.wb);
    cudaFree(l.b);
  }
  cudaFree((void*)r->tables);
  cudaFree(r->cards);
  cudaFree(r->ln);
  cudaFree(r->w3);
  delete r;
  return B2R_OK;
}

int b2r_ranker_create(b2r_ranker** out, const b2r_ranker_weights* w, int device) {
  if (!out || !w) return fail(B2R_EINVAL, "ranker_create: NULL argument");
  *out = nullptr;
  if (w->n_user < 0 || w->n_ad < 0 || w->n_user + w->n_ad < 1 || w->emb_dim < 4 || w->emb_dim % 4 != 0 ||
      w->num_numerical < 0)
    return fail(B2R_EINVAL, "ranker_create: bad embedding configuration");
  if (w->d_model < 128 || w->d_model % 128 != 0 || w->d_model > 1024 || w->d_model == 640 || w->d_model == 896)
    return fail(B2R_EUNSUPPORTED, "ranker_create: d_model must be 128, 256, 384, 512, 768 or 1024");
  if (w->d_ff < 1 || w->n_layers < 0 || w->n_cross < 0 || w->n_tasks < 1 || w->head1 < 1 || w->head2 < 1)
    return fail(B2R_EINVAL, "ranker_create: bad layer configuration");
  int ndev = 0;
  B2R_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(B2R_EINVAL, "ranker_create: bad device ordinal");
  cudaDeviceProp prop;
  B2R_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(B2R_EUNSUPPORTED, "ranker_create: this library is sm_100a only");
  B2R_CUDA(cudaSetDevice(device));
  b2r_ranker* r = new b2r_ranker();
  r->device = device;
  r->num_sms = prop.multiProcessorCount;
  r->Fu = w->n_user; r->Fa = w->n_ad; r->E = w->emb_dim; r->nnum = w->num_numerical;
  r->K1 = (r->Fu + r->Fa) * r->E + r->nnum;
  r->K1p = pad_up(r->K1, 64);
  r->dm = w->d_model; r->dff = w->d_ff; r->L = w->n_layers; r->C = w->n_cross;
  r->T = w->n_tasks; r->h1 = w->head1; r->h2 = w->head2;
  r->dmp = pad_up(r->dm, 128); r->dffp = pad_up(r->dff, 128);
  r->H1 = r->T * r->h1; r->H1p = pad_up(r->H1, 128);
  r->H2 = r->T * r->h2; r->H2p = pad_up(r->H2, 128);
  const int dm = r->dm, dff = r->dff;
  float wmax = 0.f;
  int rc = B2R_OK;
  // every activation buffer but the gathered features has its row padded to 128 columns (the GEMM epilogues
  // write whole N tiles), so the consuming layer's K is padded to 128 as well
  auto add = [&](const float* W, const float* b, int n, int k, int kp) {
    if (rc) return;
    r->lin.emplace_back();
    std::vector<float> Wv(W, W + (size_t)n * k), bv(b, b + n);
    rc = make_lin(r->lin.back(), Wv, bv, n, k, kp, &wmax);
  };
  add(w->w_proj, w->b_proj, dm, r->K1, r->K1p);
  for (int l = 0; l < r->L; ++l) {
    add(w->w_attn + (size_t)l * dm * dm, w->b_attn + (size_t)l * dm, dm, dm, r->dmp);
    add(w->w_fc1 + (size_t)l * dff * dm, w->b_fc1 + (size_t)l * dff, dff, dm, r->dmp);
    add(w->w_fc2 + (size_t)l * dm * dff, w->b_fc2 + (size_t)l * dm, dm, dff, r->dffp);
  }
  for (int c = 0; c < r->C; ++c) add(w->w_cross + (size_t)c * dm * dm, w->b_cross + (size_t)c * dm, dm, dm, r->dmp);
  add(w->w_h1, w->b_h1, r->H1, dm, r->dmp);   // the T first head layers stacked along N: [T * h1, dm]
  if (!rc) {                                   // the T second head layers as one block-diagonal map [T * h2, T * h1]
    std::vector<float> Wbd((size_t)r->H2 * r->H1, 0.f), bbd((size_t)r->H2, 0.f);
    for (int t = 0; t < r->T; ++t)
      for (int o = 0; o < r->h2; ++o) {
        for (int i = 0; i < r->h1; ++i)
          Wbd[(size_t)(t * r->h2 + o) * r->H1 + t * r->h1 + i] = w->w_h2[((size_t)t * r->h2 + o) * r->h1 + i];
        bbd[t * r->h2 + o] = w->b_h2[t * r->h2 + o];
      }
    r->lin.emplace_back();
    rc = make_lin(r->lin.back(), Wbd, bbd, r->H2, r->H1, r->H1p, &wmax);
  }
  if (!(wmax <= 65504.f)) r->bf16 = 1;
  if (!rc) {
    const int F = r->Fu + r->Fa;
    if (cudaMalloc((void**)&r->tables, (size_t)F * 8) != cudaSuccess || cudaMalloc(&r->cards, (size_t)F * 8) != cudaSuccess ||
        cudaMalloc(&r->ln, (size_t)(r->L > 0 ? r->L : 1) * 4 * dm * 4) != cudaSuccess ||
        cudaMalloc(&r->w3, (size_t)(r->H2 + r->T) * 4) != cudaSuccess) {
      rc = fail(B2R_ENOMEM, "ranker_create: cudaMalloc failed");
    } else {
      cudaMemcpy((void*)r->tables, w->tables, (size_t)F * 8, cudaMemcpyHostToDevice);
      cudaMemcpy(r->cards, w->cards, (size_t)F * 8, cudaMemcpyHostToDevice);
      for (int l = 0; l < r->L; ++l) {
        const float* src[4] = {w->ln1_g + (size_t)l * dm, w->ln1_b + (size_t)l * dm, w->ln2_g + (size_t)l * dm,
                               w->ln2_b + (size_t)l * dm};
        for (int j = 0; j < 4; ++j)
          cudaMemcpy(r->ln + ((size_t)l * 4 + j) * dm, src[j], (size_t)dm * 4, cudaMemcpyHostToDevice);
      }
      cudaMemcpy(r->w3, w->w_h3, (size_t)r->H2 * 4, cudaMemcpyHostToDevice);
      cudaMemcpy(r->w3 + r->H2, w->b_h3, (size_t)r->T * 4, cudaMemcpyHostToDevice);
      if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(B2R_ECUDA, "ranker_create: upload failed");
    }
  }
  if (rc) {
    b2r_ranker_destroy(r);
    return rc;
  }
  *out = r;
  return B2R_OK;
}

int b2r_ranker_set_param(b2r_ranker* r, const char* name, double value) {
  if (!r || !name) return fail(B2R_EINVAL, "ranker_set_param: NULL argument");
  if (std::string(name) == "operand_dtype") {
    if (value != 0 && value != 1) return fail(B2R_EINVAL, "operand_dtype must be 0 (fp16) or 1 (bf16)");
    r->bf16 = (int)value;
    return B2R_OK;
  }
  return fail(B2R_EINVAL, std::string("ranker_set_param: unknown parameter ") + name);
}

double b2r_ranker_get_param(const b2r_ranker* r, const char* name) {
  if (!r || !name) return NAN;
  if (std::string(name) == "operand_dtype") return r->bf16;
  return NAN;
}

// workspace: two 16-bit operand buffers [B, maxK], three fp32 buffers [B, d_model], one fp32 GEMM output [B, d_model]
static size_t ranker_opk(const b2r_ranker* r) {
  int m = r->K1p;
  if (r->dffp > m) m = r->dffp;
  if (r->dmp > m) m = r->dmp;
  if (r->H1p > m) m = r->H1p;
  if (r->H2p > m) m = r->H2p;
  return (size_t)m;
}
size_t b2r_ranker_workspace(const b2r_ranker* r, int64_t B) {
  if (!r || B <= 0) return 0;
  return 2 * align_up((size_t)B * ranker_opk(r) * 2, 256) + 4 * align_up((size_t)B * r->dmp * 4, 256);
}

int b2r_ranker_forward(b2r_ranker* r, const int64_t* user_cat, const int64_t* ad_cat, const float* num, int64_t B,
                       float* out, int32_t* err_flag, void* workspace, size_t ws_bytes, void* stream_) {
  if (!r) return fail(B2R_EINVAL, "ranker_forward: NULL handle");
  if (B < 0 || (B > 0 && (!out || (r->Fu > 0 && !user_cat) || (r->Fa > 0 && !ad_cat))))
    return fail(B2R_EINVAL, "ranker_forward: bad arguments");
  if (r->nnum > 0 && B > 0 && !num) return fail(B2R_EINVAL, "ranker_forward: numerical features required");
  if (B == 0) return B2R_OK;
  if (B > 0x7FFFFF00ll) return fail(B2R_EUNSUPPORTED, "ranker_forward: batch too large");
  if (!workspace || ws_bytes < b2r_ranker_workspace(r, B)) return fail(B2R_ENOMEM, "ranker_forward: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(r->device);
  if (!guard.ok) return fail(B2R_ECUDA, "ranker_forward: cudaSetDevice failed");
  const bool bf = r->bf16 != 0;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const size_t op_bytes = align_up((size_t)B * ranker_opk(r) * 2, 256), f_bytes = align_up((size_t)B * r->dmp * 4, 256);
  __half* a16 = reinterpret_cast<__half*>(ws);               // operand buffer A (K up to maxK)
  __half* b16 = reinterpret_cast<__half*>(ws + op_bytes);    // operand buffer B
  float* x = reinterpret_cast<float*>(ws + 2 * op_bytes);    // residual stream; x0 of the cross layers
  float* xl = reinterpret_cast<float*>(ws + 2 * op_bytes + f_bytes);
  float* g = reinterpret_cast<float*>(ws + 2 * op_bytes + 2 * f_bytes);
  const int dm = r->dm;
  int rc;
  // every 16-bit operand buffer is written over its whole padded width by its producer (the gather zero-fills its
  // tail, the ReLU epilogues store whole N tiles whose padding weights / biases are zero, d_model % 128 == 0)
  {
    int64_t blocks = ceil_div(B, kRkWarps);
    if (blocks > (int64_t)r->num_sms * 8) blocks = (int64_t)r->num_sms * 8;
    if (bf)
      ranker_gather_kernel<true><<<(unsigned)blocks, kRkWarps * 32, 0, stream>>>(
          r->tables, r->cards, r->Fu, r->Fa, r->E / 4, user_cat, ad_cat, num, r->nnum, B, a16, r->K1p, err_flag);
    else
      ranker_gather_kernel<false><<<(unsigned)blocks, kRkWarps * 32, 0, stream>>>(
          r->tables, r->cards, r->Fu, r->Fa, r->E / 4, user_cat, ad_cat, num, r->nnum, B, a16, r->K1p, err_flag);
    B2R_CHECK_LAUNCH("ranker_gather_kernel");
  }
  auto linear = [&](const b2r_ranker::Lin& l, const __half* act, void* o, int64_t ldo, int n_store, int mode) {
    return launch_linear(act, bf ? l.tmb : l.tm, B, l.N, l.K, l.b, o, ldo, n_store, mode, err_flag, r->num_sms, stream, bf);
  };
  auto rowop = [&](int op, float* xs, const float* x0, const float* gamma, const float* beta, __half* x16, int ld16) {
    return bf ? launch_rowop<true>(op, B, dm, g, r->dmp, xs, x0, gamma, beta, x16, ld16, err_flag, r->num_sms, stream)
              : launch_rowop<false>(op, B, dm, g, r->dmp, xs, x0, gamma, beta, x16, ld16, err_flag, r->num_sms, stream);
  };
  size_t li = 0;
  // x = projection(features) (+ positional encoding, folded into the bias); 16-bit copy in b16 [B, dmp]
  if ((rc = linear(r->lin[li++], a16, g, r->dmp, r->dmp, 2))) return rc;
  if ((rc = rowop(0, x, nullptr, nullptr, nullptr, b16, r->dmp))) return rc;
  for (int l = 0; l < r->L; ++l) {
    const float* lnp = r->ln + (size_t)l * 4 * dm;
    if ((rc = linear(r->lin[li++], b16, g, r->dmp, r->dmp, 2))) return rc;                 // folded attention
    if ((rc = rowop(1, x, nullptr, lnp, lnp + dm, b16, r->dmp))) return rc;               // x = LN1(x + attn)
    if ((rc = linear(r->lin[li++], b16, a16, r->dffp, r->dffp, 0))) return rc;            // relu(fc1 x) -> 16-bit
    if ((rc = linear(r->lin[li++], a16, g, r->dmp, r->dmp, 2))) return rc;                 // fc2
    if ((rc = rowop(1, x, nullptr, lnp + 2 * dm, lnp + 3 * dm, b16, r->dmp))) return rc;  // x = LN2(x + ff)
  }
  // cross layers: x0 = x (kept), xl starts as a copy of x0
  const __half* head_in = b16;
  if (r->C > 0) {
    B2R_CUDA(cudaMemcpyAsync(xl, x, (size_t)B * dm * 4, cudaMemcpyDeviceToDevice, stream));
    for (int c = 0; c < r->C; ++c) {
      if ((rc = linear(r->lin[li++], b16, g, r->dmp, r->dmp, 2))) return rc;               // g = xl W_c + b_c
      if ((rc = rowop(2, xl, x, nullptr, nullptr, b16, r->dmp))) return rc;               // xl = x0 * g + xl
    }
  }
  // heads: relu(stacked layer 1) -> relu(block-diagonal layer 2) -> per-head dot product
  if ((rc = linear(r->lin[li++], head_in, a16, r->H1p, r->H1p, 0))) return rc;
  if ((rc = linear(r->lin[li++], a16, b16, r->H2p, r->H2p, 0))) return rc;
  {
    int64_t blocks = ceil_div(B, kRkWarps);
    if (blocks > (int64_t)r->num_sms * 8) blocks = (int64_t)r->num_sms * 8;
    if (bf)
      ranker_head_out_kernel<true><<<(unsigned)blocks, kRkWarps * 32, 0, stream>>>(b16, r->H2p, B, r->T, r->h2, r->w3,
                                                                                  r->w3 + r->H2, out);
    else
      ranker_head_out_kernel<false><<<(unsigned)blocks, kRkWarps * 32, 0, stream>>>(b16, r->H2p, B, r->T, r->h2, r->w3,
                                                                                   r->w3 + r->H2, out);
    B2R_CHECK_LAUNCH("ranker_head_out_kernel");
  }
  return B2R_OK;
}

}  // extern "C"
