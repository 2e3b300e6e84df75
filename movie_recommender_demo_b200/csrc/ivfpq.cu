// ivfpq.cu — IVF-PQ (product quantisation, ADC with look-up tables) on the B200.
//
// Replaces faiss `IndexIVFPQ(IndexFlatIP(d), d, nlist, m, 8)` as built at
// faiss_retrieval.py:57-63 (no metric argument => faiss default METRIC_L2, by_residual = true):
//   train  : coarse k-means (ivf.cu) + per-sub-space 256-means on the residuals x - c(x)
//            (faiss ProductQuantizer: 25 iterations, <= 256*256 training points)
//   add    : residual to the assigned centroid, each of the m sub-vectors replaced by the
//            index (1 byte) of its nearest codeword
//   search : for every (query, probed list): LUT[s][j] = |(q - c)_s - codeword_{s,j}|^2, then
//            dist(x) = sum_s LUT[s][code_s(x)]  (asymmetric distance), top-k SMALLEST.
//
// Only the codes are stored (m bytes per vector, sorted by list).  The scan is byte/LUT work:
// one CTA per (query, list) pair builds its m x 256 table in shared memory (m*256*dsub FMAs)
// and streams the list's codes with 128-bit loads; one thread per code row does m shared-memory
// look-ups.  Negated distances go to the same per-pair score runs the IVF-Flat scan uses, so the
// threshold/select kernels are shared (rescore off: ADC distances are the result, as in faiss).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <random>

#include "internal.h"

namespace b2r {
namespace {

constexpr int kPqIters = 25;        // faiss ProductQuantizer cp.niter
constexpr int kPqMaxTrain = 65536;  // 256 points per codeword

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  int alloc(size_t bytes) {
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return fail(B2R_ENOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed");
    }
    return B2R_OK;
  }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// r[i] = x[i] - cent[assign[i]]
__global__ void residual_kernel(const float* __restrict__ x, const int64_t* __restrict__ assign,
                                const float* __restrict__ cent, int64_t n, int d, float* __restrict__ r) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const float* c = cent + (size_t)assign[i] * d;
  for (int j = threadIdx.x & 31; j < d; j += 32) r[i * d + j] = x[i * d + j] - c[j];
}

// nearest codeword (squared L2) of sub-vector s of every row; blockIdx.y = sub-space.
// One thread per row; the sub-space's 256 x dsub codebook is staged in shared memory.
template <int DSUB_MAX>
__global__ void __launch_bounds__(256)
pq_assign_kernel(const float* __restrict__ r, int64_t n, int d, int m, const float* __restrict__ cb,
                 uint8_t* __restrict__ codes /* [n, m] */) {
  extern __shared__ float scb[];  // [256, dsub]
  const int s = blockIdx.y, dsub = d / m;
  for (int i = threadIdx.x; i < 256 * dsub; i += blockDim.x) scb[i] = cb[(size_t)s * 256 * dsub + i];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v[DSUB_MAX];
  for (int j = 0; j < dsub; ++j) v[j] = r[i * d + s * dsub + j];
  float best = INFINITY;
  int bj = 0;
  for (int c = 0; c < 256; ++c) {
    float acc = 0.f;
    for (int j = 0; j < dsub; ++j) {
      const float df = v[j] - scb[c * dsub + j];
      acc = fmaf(df, df, acc);
    }
    if (acc < best) {  // ties keep the lower codeword index
      best = acc;
      bj = c;
    }
  }
  codes[i * m + s] = (uint8_t)bj;
}

__global__ void pq_accumulate_kernel(const float* __restrict__ r, const uint8_t* __restrict__ codes, int64_t n,
                                     int d, int m, float* __restrict__ sums /* [m,256,dsub] */,
                                     int* __restrict__ counts /* [m,256] */) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int dsub = d / m, lane = threadIdx.x & 31;
  for (int j = lane; j < d; j += 32) {
    const int s = j / dsub, jj = j - s * dsub;
    atomicAdd(sums + ((size_t)s * 256 + codes[i * m + s]) * dsub + jj, r[i * d + j]);
  }
  if (lane < m) atomicAdd(counts + lane * 256 + codes[i * m + lane], 1);
  if (lane + 32 < m) atomicAdd(counts + (lane + 32) * 256 + codes[i * m + lane + 32], 1);
}

__global__ void pq_finalize_kernel(const float* __restrict__ sums, const int* __restrict__ counts,
                                   const float* __restrict__ r, int64_t n, int d, int m, int iter,
                                   float* __restrict__ cb) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // (s, c) pair
  if (e >= m * 256) return;
  const int s = e / 256, dsub = d / m;
  const int cnt = counts[e];
  if (cnt > 0) {
    const float inv = 1.0f / (float)cnt;
    for (int j = 0; j < dsub; ++j) cb[(size_t)e * dsub + j] = sums[(size_t)e * dsub + j] * inv;
  } else {  // empty codeword: re-seed from a pseudo-random training residual
    uint64_t hsh = (uint64_t)(e + 1) * 0x9E3779B97F4A7C15ull + (uint64_t)(iter + 1) * 0xD1B54A32D192ED03ull;
    hsh ^= hsh >> 31;
    const int64_t row = (int64_t)(hsh % (uint64_t)n);
    for (int j = 0; j < dsub; ++j) cb[(size_t)e * dsub + j] = r[row * d + s * dsub + j];
  }
}

__global__ void pq_init_kernel(const float* __restrict__ r, int d, int m, float* __restrict__ cb) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // codeword c of sub-space s <- training row c
  if (e >= m * 256) return;
  const int s = e / 256, c = e % 256, dsub = d / m;
  for (int j = 0; j < dsub; ++j) cb[(size_t)e * dsub + j] = r[(size_t)c * d + s * dsub + j];
}

__global__ void scatter_codes_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t n,
                                     int m, const int32_t* __restrict__ list_src, const int64_t* __restrict__ assign_src,
                                     const uint32_t* __restrict__ perm_src, uint32_t label0,
                                     uint8_t* __restrict__ ocodes, int32_t* __restrict__ olist,
                                     uint32_t* __restrict__ operm) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t t = dst[r];
  for (int j = 0; j < m; ++j) ocodes[t * m + j] = src[r * m + j];
  olist[t] = list_src ? list_src[r] : (int32_t)assign_src[r];
  operm[t] = perm_src ? perm_src[r] : label0 + (uint32_t)r;
}

// ------------------------------------------------------------------ ADC scan ---
// |(q - c_l)_s - y_sj|^2 = |q_s - y_sj|^2  +  (2 c_ls.y_sj + |c_ls|^2)  -  2 c_ls.q_s
//                          A[q][s][j]          B[l][s][j]                  bias[s] (per pair)
// A is built once per query, B once per list (at training / import time); a (query, list) pair only
// adds the two 32 KB tables instead of re-reading the whole codebook (faiss's "precomputed table").

// B[l][s][j]: one CTA per list, thread j
__global__ void __launch_bounds__(256)
pq_list_tables_kernel(const float* __restrict__ cent, const float* __restrict__ cb, int d, int m,
                      float* __restrict__ B) {
  const int l = blockIdx.x, j = threadIdx.x, dsub = d / m, d4 = dsub >> 2;
  for (int s = 0; s < m; ++s) {
    const float4* w = reinterpret_cast<const float4*>(cb + ((size_t)s * 256 + j) * dsub);
    const float4* c = reinterpret_cast<const float4*>(cent + (size_t)l * d + s * dsub);
    float dot = 0.f, nn = 0.f;
    for (int t = 0; t < d4; ++t) {
      const float4 wv = __ldg(w + t), cv = __ldg(c + t);
      dot = fmaf(cv.x, wv.x, dot); dot = fmaf(cv.y, wv.y, dot); dot = fmaf(cv.z, wv.z, dot); dot = fmaf(cv.w, wv.w, dot);
      nn = fmaf(cv.x, cv.x, nn); nn = fmaf(cv.y, cv.y, nn); nn = fmaf(cv.z, cv.z, nn); nn = fmaf(cv.w, cv.w, nn);
    }
    B[((size_t)l * m + s) * 256 + j] = 2.0f * dot + nn;
  }
}

// A[q][s][j]: one CTA per query, thread j
__global__ void __launch_bounds__(256)
pq_query_tables_kernel(const float* __restrict__ q32, const float* __restrict__ cb, int d, int m,
                       float* __restrict__ A) {
  extern __shared__ float qs[];
  const int q = blockIdx.x, j = threadIdx.x, dsub = d / m, d4 = dsub >> 2;
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = q32[(size_t)q * d + i];
  __syncthreads();
  for (int s = 0; s < m; ++s) {
    const float4* w = reinterpret_cast<const float4*>(cb + ((size_t)s * 256 + j) * dsub);
    const float4* qv = reinterpret_cast<const float4*>(qs + s * dsub);
    float acc = 0.f;
    for (int t = 0; t < d4; ++t) {
      const float4 wv = __ldg(w + t), rv = qv[t];
      float df = rv.x - wv.x; acc = fmaf(df, df, acc);
      df = rv.y - wv.y; acc = fmaf(df, df, acc);
      df = rv.z - wv.z; acc = fmaf(df, df, acc);
      df = rv.w - wv.w; acc = fmaf(df, df, acc);
    }
    A[((size_t)q * m + s) * 256 + j] = acc;
  }
}

// One CTA per (query, probe slot) pair: LUT = A[q] + B[l] + bias in shared memory, then the list's
// codes (128-bit loads, m shared-memory look-ups per row).  Writes NEGATED distances so that
// "larger is better" like the inner-product paths.
__global__ void __launch_bounds__(256)
ivfpq_scan_kernel(const float* __restrict__ q32, int d, int m, const int64_t* __restrict__ coarse, int nprobe,
                  const float* __restrict__ cent, const float* __restrict__ A, const float* __restrict__ B,
                  const int64_t* __restrict__ list_off, const uint8_t* __restrict__ codes,
                  const int64_t* __restrict__ pair_out, float* __restrict__ scorebuf) {
  extern __shared__ float sm[];
  float* lut = sm;              // [m, 256]
  float* bias = sm + m * 256;   // [m]
  const int p = blockIdx.x;
  const int64_t l = coarse[p];
  if (l < 0) return;
  const int q = p / nprobe, dsub = d / m;
  if (threadIdx.x < m) {
    const int s = threadIdx.x;
    const float* qv = q32 + (size_t)q * d + s * dsub;
    const float* cv = cent + (size_t)l * d + s * dsub;
    float dot = 0.f;
    for (int t = 0; t < dsub; ++t) dot = fmaf(cv[t], qv[t], dot);
    bias[s] = -2.0f * dot;
  }
  __syncthreads();
  {
    const float4* a4 = reinterpret_cast<const float4*>(A + (size_t)q * m * 256);
    const float4* b4 = reinterpret_cast<const float4*>(B + (size_t)l * m * 256);
    float4* l4 = reinterpret_cast<float4*>(lut);
    for (int i = threadIdx.x; i < m * 64; i += blockDim.x) {
      const float4 av = __ldg(a4 + i), bv = __ldg(b4 + i);
      const float bs = bias[i >> 6];
      l4[i] = make_float4(av.x + bv.x + bs, av.y + bv.y + bs, av.z + bv.z + bs, av.w + bv.w + bs);
    }
  }
  __syncthreads();
  const int64_t x0 = list_off[l], x1 = list_off[l + 1];
  float* out = scorebuf + pair_out[p];
  for (int64_t i = x0 + threadIdx.x; i < x1; i += blockDim.x) {
    const uint8_t* cp = codes + i * m;
    float acc = 0.f;
    if ((m & 15) == 0) {
      for (int s0 = 0; s0 < m; s0 += 16) {
        const uint4 v = *reinterpret_cast<const uint4*>(cp + s0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int b = 0; b < 4; ++b) acc += lut[(s0 + u * 4 + b) * 256 + ((w[u] >> (8 * b)) & 255u)];
        }
      }
    } else if ((m & 7) == 0) {
      for (int s0 = 0; s0 < m; s0 += 8) {
        const uint2 v = *reinterpret_cast<const uint2*>(cp + s0);
        const uint32_t w[2] = {v.x, v.y};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int b = 0; b < 4; ++b) acc += lut[(s0 + u * 4 + b) * 256 + ((w[u] >> (8 * b)) & 255u)];
        }
      }
    } else {
      for (int s = 0; s < m; ++s) acc += lut[s * 256 + cp[s]];
    }
    out[i - x0] = -acc;
  }
  // the run is padded to a multiple of 4 scores with -inf (runs are 16-byte aligned, see ivf.cu)
  const int64_t len = x1 - x0, padded = (len + 3) & ~(int64_t)3;
  if (threadIdx.x < padded - len) out[len + threadIdx.x] = -INFINITY;
}

int pq_assign(const float* r, int64_t n, int d, int m, const float* cb, uint8_t* codes, cudaStream_t stream) {
  const int dsub = d / m;
  const size_t smem = (size_t)256 * dsub * 4;
  dim3 grid((unsigned)ceil_div(n, 256), (unsigned)m);
  if (dsub <= 8) {
    pq_assign_kernel<8><<<grid, 256, smem, stream>>>(r, n, d, m, cb, codes);
  } else if (dsub <= 32) {
    pq_assign_kernel<32><<<grid, 256, smem, stream>>>(r, n, d, m, cb, codes);
  } else {
    static bool configured[64] = {};
    int dev = 0;
    B2R_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
      B2R_CUDA(cudaFuncSetAttribute(pq_assign_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 256 * 4));
      configured[dev & 63] = true;
    }
    pq_assign_kernel<256><<<grid, 256, smem, stream>>>(r, n, d, m, cb, codes);
  }
  B2R_CHECK_LAUNCH("pq_assign_kernel");
  return B2R_OK;
}

}  // namespace

// residuals of the rows of x (already normalised as the caller wants) w.r.t. their coarse centroid
int pq_residuals(b2r_index* h, int64_t n, const float* x, const int64_t* assign, float* r, cudaStream_t stream) {
  residual_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, stream>>>(x, assign, h->quantizer->x32, n, h->d, r);
  B2R_CHECK_LAUNCH("residual_kernel");
  return B2R_OK;
}

int pq_train(b2r_index* h, int64_t n, const float* resid, uint64_t seed, cudaStream_t stream) {
  (void)seed;
  const int d = h->d, m = h->pq_m, dsub = d / m;
  if (n < 256) return fail(B2R_EINVAL, "index_train: IVF-PQ needs at least 256 training vectors");
  const int64_t nt = n < kPqMaxTrain ? n : kPqMaxTrain;  // caller passes a seeded random subsample first
  if (!h->codebooks && cudaMalloc(&h->codebooks, (size_t)m * 256 * dsub * 4) != cudaSuccess)
    return fail(B2R_ENOMEM, "cudaMalloc codebooks");
  DevBuf codes, sums, counts;
  int rc;
  if ((rc = codes.alloc((size_t)nt * m))) return rc;
  if ((rc = sums.alloc((size_t)m * 256 * dsub * 4))) return rc;
  if ((rc = counts.alloc((size_t)m * 256 * 4))) return rc;
  pq_init_kernel<<<(unsigned)ceil_div(m * 256, 128), 128, 0, stream>>>(resid, d, m, h->codebooks);
  B2R_CHECK_LAUNCH("pq_init_kernel");
  for (int it = 0; it < kPqIters; ++it) {
    if ((rc = pq_assign(resid, nt, d, m, h->codebooks, codes.as<uint8_t>(), stream))) return rc;
    B2R_CUDA(cudaMemsetAsync(sums.p, 0, (size_t)m * 256 * dsub * 4, stream));
    B2R_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)m * 256 * 4, stream));
    pq_accumulate_kernel<<<(unsigned)ceil_div(nt, 8), 256, 0, stream>>>(resid, codes.as<uint8_t>(), nt, d, m,
                                                                       sums.as<float>(), counts.as<int>());
    B2R_CHECK_LAUNCH("pq_accumulate_kernel");
    pq_finalize_kernel<<<(unsigned)ceil_div(m * 256, 128), 128, 0, stream>>>(sums.as<float>(), counts.as<int>(), resid, nt,
                                                                             d, m, it, h->codebooks);
    B2R_CHECK_LAUNCH("pq_finalize_kernel");
  }
  h->pq_trained = true;
  return pq_build_list_tables(h, stream);
}

int pq_encode(b2r_index* h, int64_t n, const float* resid, uint8_t* codes, cudaStream_t stream) {
  return pq_assign(resid, n, h->d, h->pq_m, h->codebooks, codes, stream);
}

int pq_scatter_codes(const uint8_t* src, const int64_t* dst, int64_t n, int m, const int32_t* list_src,
                     const int64_t* assign_src, const uint32_t* perm_src, uint32_t label0, uint8_t* ocodes,
                     int32_t* olist, uint32_t* operm, cudaStream_t stream) {
  if (n <= 0) return B2R_OK;
  scatter_codes_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(src, dst, n, m, list_src, assign_src, perm_src,
                                                                      label0, ocodes, olist, operm);
  B2R_CHECK_LAUNCH("scatter_codes_kernel");
  return B2R_OK;
}

int pq_build_list_tables(b2r_index* h, cudaStream_t stream) {
  if (h->kind != B2R_KIND_IVF_PQ || !h->pq_trained || !h->quantizer || h->quantizer->ntotal != h->nlist) return B2R_OK;
  const size_t bytes = (size_t)h->nlist * h->pq_m * 256 * 4;
  if (!h->pq_list_tab && cudaMalloc(&h->pq_list_tab, bytes) != cudaSuccess) {
    cudaGetLastError();
    return fail(B2R_ENOMEM, "cudaMalloc of the PQ per-list tables (" + std::to_string(bytes) + " bytes) failed");
  }
  pq_list_tables_kernel<<<h->nlist, 256, 0, stream>>>(h->quantizer->x32, h->codebooks, h->d, h->pq_m, h->pq_list_tab);
  B2R_CHECK_LAUNCH("pq_list_tables_kernel");
  return B2R_OK;
}

int pq_scan(b2r_index* h, int nq, int npairs, const float* q32, const int64_t* coarse, int nprobe,
            const int64_t* pair_out, float* qtab, float* scorebuf, cudaStream_t stream) {
  const int d = h->d, m = h->pq_m;
  if (!h->pq_list_tab) return fail(B2R_ESTATE, "IVF-PQ per-list tables are missing");
  pq_query_tables_kernel<<<nq, 256, (size_t)d * 4, stream>>>(q32, h->codebooks, d, m, qtab);
  B2R_CHECK_LAUNCH("pq_query_tables_kernel");
  const size_t smem = (size_t)m * 256 * 4 + (size_t)m * 4;
  static bool configured[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2R_CUDA(cudaFuncSetAttribute(ivfpq_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 256 * 4 + 256));
    configured[dev & 63] = true;
  }
  ivfpq_scan_kernel<<<npairs, 256, smem, stream>>>(q32, d, m, coarse, nprobe, h->quantizer->x32, qtab, h->pq_list_tab,
                                                   h->list_off, h->codes, pair_out, scorebuf);
  B2R_CHECK_LAUNCH("ivfpq_scan_kernel");
  return B2R_OK;
}

}  // namespace b2r

using namespace b2r;

extern "C" {

int b2r_index_export_codebooks(const b2r_index* h, float* out) {
  if (!h || !out) return fail(B2R_EINVAL, "export_codebooks: NULL argument");
  if (h->kind != B2R_KIND_IVF_PQ) return fail(B2R_EUNSUPPORTED, "no PQ codebooks on this index kind");
  if (!h->pq_trained) return fail(B2R_ESTATE, "export_codebooks: index is not trained");
  DeviceGuard g(h->device);
  B2R_CUDA(cudaMemcpy(out, h->codebooks, (size_t)h->pq_m * 256 * (h->d / h->pq_m) * 4, cudaMemcpyDeviceToHost));
  return B2R_OK;
}

int b2r_index_import_codebooks(b2r_index* h, const float* in) {
  if (!h || !in) return fail(B2R_EINVAL, "import_codebooks: NULL argument");
  if (h->kind != B2R_KIND_IVF_PQ) return fail(B2R_EUNSUPPORTED, "no PQ codebooks on this index kind");
  if (h->ntotal > 0) return fail(B2R_ESTATE, "import_codebooks: index already holds vectors");
  DeviceGuard g(h->device);
  const size_t bytes = (size_t)h->pq_m * 256 * (h->d / h->pq_m) * 4;
  if (!h->codebooks && cudaMalloc(&h->codebooks, bytes) != cudaSuccess) return fail(B2R_ENOMEM, "cudaMalloc codebooks");
  B2R_CUDA(cudaMemcpy(h->codebooks, in, bytes, cudaMemcpyHostToDevice));
  h->pq_trained = true;
  h->trained = h->quantizer && h->quantizer->ntotal == h->nlist;
  int rc = pq_build_list_tables(h, 0);
  if (rc) return rc;
  B2R_CUDA(cudaStreamSynchronize(0));
  return B2R_OK;
}

/* PQ codes of the STORED rows [row0,row0+n) (uint8 [n, pq_m], device) — parity plumbing. */
int b2r_index_get_codes(const b2r_index* h, int64_t row0, int64_t n, uint8_t* out, void* stream) {
  if (!h || (n > 0 && !out)) return fail(B2R_EINVAL, "get_codes: NULL argument");
  if (h->kind != B2R_KIND_IVF_PQ) return fail(B2R_EUNSUPPORTED, "no PQ codes on this index kind");
  if (row0 < 0 || n < 0 || row0 + n > h->ntotal) return fail(B2R_EINVAL, "get_codes: range out of bounds");
  if (n == 0) return B2R_OK;
  DeviceGuard g(h->device);
  B2R_CUDA(cudaMemcpyAsync(out, h->codes + (size_t)row0 * h->pq_m, (size_t)n * h->pq_m, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return B2R_OK;
}

}  // extern "C"
