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
#include "ptx.cuh"

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

// t[row] = sum_s B[list(row)][s][code_s(row)]  -- the list-dependent part of the ADC distance, per stored row
__global__ void __launch_bounds__(256)
pq_row_terms_kernel(const uint8_t* __restrict__ codes, const int32_t* __restrict__ row_list,
                    const float* __restrict__ B, int64_t n, int m, float* __restrict__ t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* b = B + (size_t)row_list[i] * m * 256;
  const uint8_t* c = codes + i * m;
  float acc = 0.f;
  for (int s = 0; s < m; ++s) acc += __ldg(b + s * 256 + c[s]);
  t[i] = acc;
}

// Query-major ADC scan (pq_m in {8, 16, 32, 64}).
//
//   dist(q, x) = sum_s A[q][s][code_s(x)]  +  t[x]  -  2 <c_l, q>        (expansion above, regrouped)
//                look-ups, per query          per row    per (query, list)
//
// so the look-up table depends on the QUERY only: one CTA builds it once and scans all the lists its
// query probes (blockIdx.y splits the probe slots when there are few queries).  The scan is bound by
// shared-memory look-ups; with the obvious layout lut[s][code] the 32 lanes of a warp hit random banks
// (measured ~3.5-way conflicts, 4.4 ms at 10M x m32, Q=4096).  Here a lane owns 4 sub-quantisers (one
// 32-bit word of the code row), M/4 lanes share a row, 128/M rows are in flight per warp pass, and the
// table is stored as lut[code][128]: 128/M replicas of the M entries, each replica's entries rotated so
// that in every look-up instruction the 32 lanes touch 32 distinct banks whatever the codes are.
template <int M>
__device__ __forceinline__ int pq_slot(int r, int j, int t) {
  constexpr int LR = M / 4;
  const int rho = (M <= 32) ? r / (32 / (M <= 32 ? M : 32)) : r;
  return r * M + (((t + rho) & 3) * LR + j);
}

constexpr int kPqMaxSlots = 256;   // probe slots one CTA of the query-major scan can hold
constexpr int kPqConsumers = 15;   // look-up warps per group; tile g belongs to group g % kPqGroups
constexpr int kPqGroups = 2;
constexpr int kPqThreads = (kPqGroups * kPqConsumers + 1) * 32;   // + one warp that feeds the ring

template <int M> struct PqTile {
  static constexpr int ROWS = ((M == 32) ? 1 : 2) * kPqConsumers * 32;   // code rows per tile: 1-2 blocks per warp
  static constexpr int CODE_BYTES = ROWS * M + 16;          // <= 16 KB (+16: lists start on M-byte, not 16-byte, bounds)
  static constexpr int TERM_BYTES = (ROWS + 4) * 4;         // the rows' list terms (+4: same for 4-byte bounds)
  static constexpr int BYTES = CODE_BYTES + TERM_BYTES;
  // tiles in flight (TMA bulk-copy ring): as many as fit beside the 128 KB table -- the scan is bound by
  // bytes in flight (measured 3-4 TB/s with 3 tiles ahead)
  static constexpr int BUFS = (M == 32) ? 5 : (M == 16) ? 4 : 7;
};

// Warp-specialised: the last warp streams (list, tile) after (list, tile) of code rows + row terms into a
// ring of shared-memory buffers with 1-D TMA bulk copies (two instructions per tile); two groups of 15
// consumer warps take alternate tiles: wait on the tile's mbarrier, look their 32-row blocks up, release
// the buffer -- no block-wide barrier in the loop, so the warps drift apart and hide each other's
// shared-memory latencies.  The loop is bound by the shared-memory pipe (look-ups + shuffles), so the M/4
// partial sums of a warp pass are combined with a transposing butterfly: M/4 - 1 shuffles per 32 rows
// instead of (M/4) log2(M/4).
template <int M>
__global__ void __launch_bounds__(kPqThreads, 1)
ivfpq_scan_query_kernel(const float* __restrict__ q32, int d, const int64_t* __restrict__ coarse, int nprobe,
                        const float* __restrict__ cent, const float* __restrict__ cb,
                        const float* __restrict__ row_term, const int64_t* __restrict__ list_off,
                        const uint8_t* __restrict__ codes, const int64_t* __restrict__ pair_out,
                        float* __restrict__ scorebuf) {
  using T = PqTile<M>;
  constexpr int kPqBufs = T::BUFS;
  constexpr int LR = M / 4, RP = 32 / LR;   // lanes per row, rows per warp pass (= replicas)
  extern __shared__ __align__(128) uint8_t smraw[];
  float* lut = reinterpret_cast<float*>(smraw);                       // [256][128]
  uint8_t* tiles = smraw + 256 * 128 * 4;                             // kPqBufs x (codes | row terms)
  float* qs = reinterpret_cast<float*>(tiles + kPqBufs * T::BYTES);   // [d]
  int64_t* p_x0 = reinterpret_cast<int64_t*>(qs + d);                 // per probe slot of this CTA:
  int64_t* p_out = p_x0 + kPqMaxSlots;                                //   first stored row, score-run offset,
  int* p_len = reinterpret_cast<int*>(p_out + kPqMaxSlots);           //   list length,
  int* p_t0 = p_len + kPqMaxSlots;                                    //   first tile number (cumulative, +1),
  float* p_bias = reinterpret_cast<float*>(p_t0 + kPqMaxSlots + 2);   //   -2 <c_l, q>
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_bias + kPqMaxSlots); // full[kPqBufs], empty[kPqBufs]
  // Tiles the producer has issued so far.  With an odd ring depth a buffer alternates between the two consumer
  // groups, so a group can reach its wait for use u of a buffer without ever having waited for use u-1 - and
  // `try_wait.parity` cannot tell "use u landed" from "use u-1 has not landed yet" (same parity).  The producer
  // issues use u only after use u-1 was consumed (empty barrier), so a consumer that first sees `issued > g`
  // knows the barrier is in phase u or u+1 and the parity test is unambiguous.  (The old kernel without this
  // check died once per ~2e9 tiles: profiles/r02_ivfpq_ring_fault_old_kernel.log.)
  int* issued = reinterpret_cast<int*>(bars + 2 * kPqBufs);
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int dsub = d / M;
  const int nslots = (nprobe - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;   // pr = y, y + Y, ...
  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](int i) { return bar0 + 8u * i; };
  auto bar_empty = [&](int i) { return bar0 + 8u * (kPqBufs + i); };

  if (tid == 0) {
    for (int i = 0; i < kPqBufs; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), kPqConsumers); }   // one group per tile
    *issued = 0;
    fence_mbar_init();
  }
  for (int i = tid; i < d; i += blockDim.x) qs[i] = q32[(size_t)q * d + i];
  for (int k = tid; k < nslots; k += blockDim.x) {
    const int pr = blockIdx.y + k * gridDim.y;
    const int64_t l = coarse[(size_t)q * nprobe + pr];
    int64_t x0 = 0, len = 0;
    if (l >= 0) { x0 = list_off[l]; len = list_off[l + 1] - x0; }
    p_x0[k] = x0;
    p_len[k] = (int)len;
    p_out[k] = pair_out[(size_t)q * nprobe + pr];
  }
  __syncthreads();
  // look-up table: |q_s - y_sj|^2 for every (sub-quantiser s, code j), written to every replica
  for (int i = tid; i < 256 * M; i += blockDim.x) {
    const int c = i / M, s = i % M;
    const float* w = cb + ((size_t)s * 256 + c) * dsub;
    const float* qv = qs + s * dsub;
    float acc = 0.f;
    for (int t = 0; t < dsub; t += 4) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + t));
      float df = qv[t] - wv.x; acc = fmaf(df, df, acc);
      df = qv[t + 1] - wv.y; acc = fmaf(df, df, acc);
      df = qv[t + 2] - wv.z; acc = fmaf(df, df, acc);
      df = qv[t + 3] - wv.w; acc = fmaf(df, df, acc);
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) lut[c * 128 + pq_slot<M>(r, s >> 2, s & 3)] = acc;
  }
  // -2 <c_l, q> per probe slot: one warp per slot, the centroid's loads issued together (d <= 256)
  for (int k = warp; k < nslots; k += nwarps) {
    const int pr = blockIdx.y + k * gridDim.y;
    const int64_t l = coarse[(size_t)q * nprobe + pr];
    float cv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) cv[u] = (l >= 0 && lane + 32 * u < d) ? __ldg(cent + (size_t)l * d + lane + 32 * u) : 0.f;
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) dot = fmaf(cv[u], lane + 32 * u < d ? qs[lane + 32 * u] : 0.f, dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) p_bias[k] = -2.0f * dot;
  }
  if (tid == 0) {
    int cum = 0;
    for (int k = 0; k < nslots; ++k) {
      p_t0[k] = cum;
      cum += (p_len[k] + T::ROWS - 1) / T::ROWS;
    }
    p_t0[nslots] = cum;
  }
  __syncthreads();
  const int ntiles = p_t0[nslots];
  const uint32_t tiles_s = smem_u32(tiles);
  const uint32_t issued_s = smem_u32(issued);

  if (warp == kPqGroups * kPqConsumers) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      int k = 0;
      for (int g = 0; g < ntiles; ++g) {
        const int buf = g % kPqBufs, use = g / kPqBufs;
        if (use > 0) mbar_wait(bar_empty(buf), (uint32_t)((use - 1) & 1), 41);
        while (g >= p_t0[k + 1]) ++k;
        const int row0 = (g - p_t0[k]) * T::ROWS;
        const int nrows = min(T::ROWS, p_len[k] - row0);
        const int64_t first = p_x0[k] + row0;
        // both sources are aligned DOWN to 16 bytes; the consumers skip the same number of leading bytes
        const uint64_t cbyte = (uint64_t)first * M;
        const uint32_t cskip = (uint32_t)(cbyte & 15u);
        const uint32_t cbytes = (cskip + (uint32_t)nrows * M + 15u) & ~15u;
        const uint32_t tskip = (uint32_t)(first & 3);
        const uint32_t tbytes = ((tskip + (uint32_t)nrows + 3u) & ~3u) * 4u;
        const uint32_t dst = tiles_s + (uint32_t)(buf * T::BYTES);
        mbar_arrive_expect_tx(bar_full(buf), cbytes + tbytes);
        bulk_load_1d(dst, codes + (cbyte - cskip), cbytes, bar_full(buf));
        bulk_load_1d(dst + T::CODE_BYTES, row_term + (first - tskip), tbytes, bar_full(buf));
        st_release_shared_u32(issued_s, (uint32_t)(g + 1));
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers
    const int rsub = lane / LR, j = lane % LR;
    const int grp = warp / kPqConsumers, cw = warp % kPqConsumers;
    const float* lb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) lb[t] = lut + pq_slot<M>(rsub, j, t);
    int k = 0;
#pragma unroll 1
    for (int g = grp; g < ntiles; g += kPqGroups) {
      const int buf = g % kPqBufs, use = g / kPqBufs;
      while (g >= p_t0[k + 1]) ++k;
      const int row0 = (g - p_t0[k]) * T::ROWS;
      const int len = p_len[k];
      const int nrows = min(T::ROWS, len - row0);
      const int64_t first = p_x0[k] + row0;
      const uint8_t* tile = tiles + buf * T::BYTES;
      const uint32_t* tw = reinterpret_cast<const uint32_t*>(tile + (((uint64_t)first * M) & 15u));
      const float* tt = reinterpret_cast<const float*>(tile + T::CODE_BYTES) + (first & 3);
      float* out = scorebuf + p_out[k] + row0;
      const float b = p_bias[k];
      for (uint32_t spin = 0; (int)ld_acquire_shared_u32(issued_s) <= g; ++spin)
        if (spin > (1u << 30)) __trap();        // bounded like mbar_wait: a protocol bug must not hang the GPU
      mbar_wait(bar_full(buf), (uint32_t)(use & 1), 42);
      for (int base = cw * 32; base < nrows; base += kPqConsumers * 32) {
        float a[LR];
        const uint32_t* twl = tw + base * LR + lane;   // word (row base + ps*RP + rsub, j) = twl[ps * 32]
#pragma unroll
        for (int ps = 0; ps < LR; ++ps) {
          const uint32_t w = twl[ps * 32];   // rows past the list end: stale bytes, discarded
          // one PRMT per byte (zero-extended), one IMAD for the address: the loop is issue-bound
          float v = lb[0][__byte_perm(w, 0u, 0x4440u) * 128];
          v += lb[1][__byte_perm(w, 0u, 0x4441u) * 128];
          v += lb[2][__byte_perm(w, 0u, 0x4442u) * 128];
          v += lb[3][__byte_perm(w, 0u, 0x4443u) * 128];
          a[ps] = v;
        }
        // transposing butterfly over the LR lanes of a row group: lane j ends with the total of pass j
#pragma unroll
        for (int o = LR / 2, n = LR; o >= 1; o >>= 1, n >>= 1) {
          const bool up = (j & o) != 0;
#pragma unroll
          for (int i = 0; i < n / 2; ++i) {
            const float send = up ? a[i] : a[i + n / 2];
            const float keep = up ? a[i + n / 2] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        const int my = base + j * RP + rsub;
        if (my < nrows) out[my] = -(a[0] + tt[my] + b);
      }
      if (cw == 0 && row0 + nrows == len) {
        // last tile of the list: the run is padded to a multiple of 4 scores with -inf (16-byte aligned runs, ivf.cu)
        const int padded = (len + 3) & ~3;
        if (lane < padded - len) out[nrows + lane] = -INFINITY;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty(buf));
    }
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
  return pq_update_row_terms(h, stream);   // stored rows (if any) depend on the tables
}

int pq_update_row_terms(b2r_index* h, cudaStream_t stream) {
  if (h->kind != B2R_KIND_IVF_PQ || !h->pq_list_tab || !h->codes || h->ntotal <= 0) return B2R_OK;
  if (h->pq_row_term_cap < h->ntotal) {
    cudaFree(h->pq_row_term);
    h->pq_row_term = nullptr;
    const int64_t cap = (int64_t)align_up((size_t)h->ntotal, 1024);
    if (cudaMalloc(&h->pq_row_term, (size_t)cap * 4 + 64) != cudaSuccess) {   // +64: bulk copies read 16-byte granules
      cudaGetLastError();
      h->pq_row_term_cap = 0;
      return fail(B2R_ENOMEM, "cudaMalloc of the PQ per-row terms failed");
    }
    h->pq_row_term_cap = cap;
  }
  pq_row_terms_kernel<<<(unsigned)ceil_div(h->ntotal, 256), 256, 0, stream>>>(h->codes, h->row_list, h->pq_list_tab,
                                                                             h->ntotal, h->pq_m, h->pq_row_term);
  B2R_CHECK_LAUNCH("pq_row_terms_kernel");
  return B2R_OK;
}

template <int M>
static int launch_scan_query(b2r_index* h, int nq, const float* q32, const int64_t* coarse, int nprobe,
                             const int64_t* pair_out, float* scorebuf, cudaStream_t stream) {
  auto kern = ivfpq_scan_query_kernel<M>;
  constexpr int kPqBufs = PqTile<M>::BUFS;
  const size_t smem = (size_t)256 * 128 * 4 + (size_t)kPqBufs * PqTile<M>::BYTES + (size_t)h->d * 4 +
                      (size_t)kPqMaxSlots * 28 + 16 + 2 * kPqBufs * 8 + 16;
  static bool configured[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  256 * 128 * 4 + kPqBufs * PqTile<M>::BYTES + 1024 * 4 + kPqMaxSlots * 28 + 16 + 2 * kPqBufs * 8 + 16));
    configured[dev & 63] = true;
  }
  // few queries: split a query's probe slots over several CTAs so that the grid still covers the SMs;
  // many probe slots: at most kPqMaxSlots per CTA
  int slices = (int)ceil_div(nprobe, kPqMaxSlots);
  while (slices < nprobe && (int64_t)nq * slices < 2 * h->num_sms) ++slices;
  dim3 grid((unsigned)nq, (unsigned)slices);
  kern<<<grid, kPqThreads, smem, stream>>>(q32, h->d, coarse, nprobe, h->quantizer->x32, h->codebooks, h->pq_row_term,
                                    h->list_off, h->codes, pair_out, scorebuf);
  B2R_CHECK_LAUNCH("ivfpq_scan_query_kernel");
  return B2R_OK;
}

int pq_scan(b2r_index* h, int nq, int npairs, const float* q32, const int64_t* coarse, int nprobe,
            const int64_t* pair_out, float* qtab, float* scorebuf, cudaStream_t stream) {
  const int d = h->d, m = h->pq_m;
  if (!h->pq_list_tab) return fail(B2R_ESTATE, "IVF-PQ per-list tables are missing");
  const bool query_major = (m == 8 || m == 16 || m == 32) && h->pq_row_term && (d / m) % 4 == 0 && h->pq_scan_path != 1;
  if (query_major) {
    switch (m) {
      case 8: return launch_scan_query<8>(h, nq, q32, coarse, nprobe, pair_out, scorebuf, stream);
      case 16: return launch_scan_query<16>(h, nq, q32, coarse, nprobe, pair_out, scorebuf, stream);
      default: return launch_scan_query<32>(h, nq, q32, coarse, nprobe, pair_out, scorebuf, stream);
    }
  }
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
