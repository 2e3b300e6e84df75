// ivf.cu — IVF-Flat (inner product) on the B200: coarse quantiser, inverted lists, list scan.
//
// Replaces faiss `IndexIVFFlat(IndexFlatIP(d), d, nlist, METRIC_INNER_PRODUCT)` as built at
// faiss_retrieval.py:50-55 and driven through .train (:93), .add (:118), .search (:155):
//   train  : k-means with inner-product assignment, 25 iterations, <= 256 points per centroid (IVF-Flat: spherical,
//            a documented deviation - see kKmeansIters)
//   add    : assign each vector to the centroid of maximum inner product; append to that list
//   search : top-nprobe centroids by inner product, exact IP over the vectors of those lists
//
// B200 layout: the corpus rows are stored SORTED BY LIST (fp32 master + bf16 scan copy), so a
// list is a contiguous row range that TMA can tile.  The coarse quantiser is a nested flat
// index over the centroids (exact fp32-rescored top-1 / top-nprobe, like faiss's IndexFlatIP
// quantiser).  Search is LIST-MAJOR: the (query, probed list) pairs are counting-sorted by list,
// the queries of a list are gathered into a contiguous bf16 block, and one tcgen05 pass scores
// 128 queries x a slice of the list, so a list is read from HBM once per 128 probing queries
// instead of once per query.  Pair scores are dumped (4 B/score against 512 B/row read), a
// radix select takes each query's threshold, and the shared select_rescore kernel does the
// exact fp32 rescore + canonical sort.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <random>

#include "internal.h"
#include "ptx.cuh"

namespace b2r {
namespace {

// faiss ClusteringParameters: niter = 25, max_points_per_centroid = 256, spherical = false.  Deviation:
// the IVF-Flat centroids ARE L2-normalised after every update (spherical k-means).  The wrapper stores unit-norm
// rows (faiss_retrieval.py:115), so argmax <c, x> with raw means favours lists whose mean happens to be long; with
// unit centroids the assignment is the nearest direction and the lists come out balanced.  Parity never depends on
// it: faiss's own RNG / subsampling cannot be reproduced, list membership is compared on SHARED centroids.
constexpr int kKmeansIters = 25;
constexpr int kMaxPointsPerCentroid = 256;  // faiss ClusteringParameters.max_points_per_centroid
constexpr int kUnitTilesSmallQ = 8;       // corpus tiles (of 128 rows) per scan unit: few queries -> split lists for parallelism
constexpr int kUnitTilesLargeQ = 64;      // many queries -> whole lists per unit (one query-block load per list, fewer pipeline drains)

// ------------------------------------------------------------------ small kernels ---
__global__ void gather_rows_f32_kernel(const float* __restrict__ src, const int64_t* __restrict__ rows,
                                       int64_t n, int d, float* __restrict__ dst) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const float4* s = reinterpret_cast<const float4*>(src + rows[r] * d);
  float4* o = reinterpret_cast<float4*>(dst + r * d);
  for (int j = threadIdx.x & 31; j < (d >> 2); j += 32) o[j] = s[j];
}

// sums[list] += x[row]; counts[list] += 1   (k-means update, fp32 atomics)
__global__ void kmeans_accumulate_kernel(const float* __restrict__ x, const int64_t* __restrict__ assign,
                                         int64_t n, int d, float* __restrict__ sums, int* __restrict__ counts) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t c = assign[r];
  if (c < 0) return;
  const int lane = threadIdx.x & 31;
  for (int j = lane; j < d; j += 32) atomicAdd(sums + c * d + j, x[r * d + j]);
  if (lane == 0) atomicAdd(counts + c, 1);
}

// centroid = mean (then L2-normalised: spherical k-means for the IP metric); an empty cluster is
// re-seeded from a pseudo-random training point
__global__ void kmeans_finalize_kernel(const float* __restrict__ sums, const int* __restrict__ counts,
                                       const float* __restrict__ x, int64_t n, int d, int nlist, int iter,
                                       int spherical, float* __restrict__ cent) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= nlist) return;
  const int lane = threadIdx.x & 31;
  const int cnt = counts[c];
  const float* src;
  float scale;
  if (cnt > 0) {
    src = sums + (size_t)c * d;
    scale = 1.0f / (float)cnt;
  } else {
    uint64_t hsh = (uint64_t)(c + 1) * 0x9E3779B97F4A7C15ull + (uint64_t)(iter + 1) * 0xD1B54A32D192ED03ull;
    hsh ^= hsh >> 29;
    src = x + (size_t)(hsh % (uint64_t)n) * d;
    scale = 1.0f;
  }
  float ss = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float v = src[j] * scale;
    ss += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  float post = 1.0f;
  if (spherical && ss > 0.f) post = 1.0f / sqrtf(ss);
  for (int j = lane; j < d; j += 32) cent[(size_t)c * d + j] = src[j] * scale * post;
}

__global__ void hist_lists_kernel(const int64_t* __restrict__ a, int64_t n, int nlist, int* __restrict__ hist) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int64_t l = a[i];
    if (l >= 0 && l < nlist) atomicAdd(hist + l, 1);
  }
}

// single block: exclusive scan of hist[0..n) (+ optional old sizes) into off[0..n]; cursor = starts
__global__ void scan_lists_kernel(const int* __restrict__ hist, const int64_t* __restrict__ old_off, int n,
                                  int64_t* __restrict__ off, int64_t* __restrict__ cursor_new) {
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b = 0; b < n; b += blockDim.x) {
    const int i = b + threadIdx.x;
    const int64_t old_sz = (old_off && i < n) ? old_off[i + 1] - old_off[i] : 0;
    int64_t v = i < n ? (int64_t)hist[i] + old_sz : 0;
    // block inclusive scan (warp scans + shared partials)
    __shared__ int64_t wsum[32];
    int64_t incl = v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      int64_t s = lane < (blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const int64_t before = carry + (w > 0 ? wsum[w - 1] : 0) + incl - v;
    if (i < n) {
      off[i] = before;
      if (cursor_new) cursor_new[i] = before + old_sz;  // new rows of list i go after its old rows
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = carry;
}

// destination of every row of the merged (old sorted rows + new rows) corpus
__global__ void place_old_rows_kernel(const int64_t* __restrict__ old_off, const int64_t* __restrict__ new_off,
                                      int nlist, int64_t n_old, const int32_t* __restrict__ row_list,
                                      int64_t* __restrict__ dst_of_old) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_old) return;
  const int l = row_list[i];
  dst_of_old[i] = new_off[l] + (i - old_off[l]);
}
__global__ void place_new_rows_kernel(const int64_t* __restrict__ assign, int64_t n_new,
                                      unsigned long long* __restrict__ cursor, int64_t* __restrict__ dst_of_new) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_new) return;
  dst_of_new[i] = (int64_t)atomicAdd(cursor + assign[i], 1ull);
}
// move rows (fp32 + bf16 + list id + label) to their destinations
__global__ void scatter_rows_kernel(const float* __restrict__ s32, const __nv_bfloat16* __restrict__ s16,
                                    const int64_t* __restrict__ dst, int64_t n, int d,
                                    const int32_t* __restrict__ list_src, const int64_t* __restrict__ assign_src,
                                    const uint32_t* __restrict__ perm_src, uint32_t label0,
                                    float* __restrict__ o32, __nv_bfloat16* __restrict__ o16,
                                    int32_t* __restrict__ olist, uint32_t* __restrict__ operm) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const int lane = threadIdx.x & 31;
  const int64_t t = dst[r];
  const float4* a = reinterpret_cast<const float4*>(s32 + r * d);
  float4* b = reinterpret_cast<float4*>(o32 + t * d);
  for (int j = lane; j < (d >> 2); j += 32) b[j] = a[j];
  const uint2* a2 = reinterpret_cast<const uint2*>(s16 + r * d);
  uint2* b2 = reinterpret_cast<uint2*>(o16 + t * d);
  for (int j = lane; j < (d >> 2); j += 32) b2[j] = a2[j];
  if (lane == 0) {
    olist[t] = list_src ? list_src[r] : (int32_t)assign_src[r];
    operm[t] = perm_src ? perm_src[r] : label0 + (uint32_t)r;
  }
}

// ------------------------------------------------------------- search plumbing ---
// per query: run offsets of its nprobe pair-score runs and the total number of scanned rows
__global__ void pair_runs_kernel(const int64_t* __restrict__ coarse, int Q, int nprobe,
                                 const int64_t* __restrict__ list_off, int64_t smax,
                                 int64_t* __restrict__ pair_out, int* __restrict__ row_len,
                                 int* __restrict__ list_cnt) {
  // one WARP per query, a lane per probe: the coarse -> list_off loads are two dependent round trips each, and a
  // single thread walking nprobe = 32 of them serially was 8 us of a batch-1 search
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  int64_t cum = 0;
  for (int j0 = 0; j0 < nprobe; j0 += 32) {
    const int j = j0 + lane;
    int64_t len = 0;
    if (j < nprobe) {
      const int64_t l = coarse[(size_t)q * nprobe + j];
      if (l >= 0) {
        len = (list_off[l + 1] - list_off[l] + 3) & ~(int64_t)3;   // runs start 16-byte aligned (float4 stores)
        atomicAdd(list_cnt + l, 1);
      }
    }
    int64_t incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (j < nprobe) pair_out[(size_t)q * nprobe + j] = (int64_t)q * smax + cum + (incl - len);
    cum += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) row_len[q] = (int)cum;
}

// counting-sort scatter of pairs by list + gather of the pair's bf16 query row
__global__ void scatter_pairs_kernel(const int64_t* __restrict__ coarse, int npairs, int nprobe,
                                     unsigned long long* __restrict__ cursor, int* __restrict__ pair_sorted,
                                     const __nv_bfloat16* __restrict__ q16, int d,
                                     __nv_bfloat16* __restrict__ gq16, int* __restrict__ pair_pos) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= npairs) return;
  const int lane = threadIdx.x & 31;
  const int64_t l = coarse[p];
  if (l < 0) return;
  unsigned long long pos = 0;
  if (lane == 0) pos = atomicAdd(cursor + l, 1ull);
  pos = __shfl_sync(0xffffffffu, pos, 0);
  if (lane == 0) {
    pair_sorted[pos] = p;
    pair_pos[p] = (int)pos;      // inverse map: where the fused path's candidate gather finds this pair's unit lane
  }
  const uint2* s = reinterpret_cast<const uint2*>(q16 + (size_t)(p / nprobe) * d);
  uint2* o = reinterpret_cast<uint2*>(gq16 + (size_t)pos * d);
  for (int j = lane; j < (d >> 2); j += 32) o[j] = s[j];
}

struct IvfUnit {
  int qrow0;      // first row of the gathered query matrix
  int nq;         // valid lanes (<= 128)
  int xrow0;      // first corpus row of this unit's tile range
  int ntiles;
  int xend;       // list end row (rows >= xend are masked)
  int xlist0;     // list start row (score-run offsets are relative to it)
};

__global__ void build_units_kernel(const int64_t* __restrict__ list_off, const int64_t* __restrict__ pair_off,
                                   int nlist, IvfUnit* __restrict__ units, int* __restrict__ num_units,
                                   int max_units, int kUnitTiles, int* __restrict__ list_unit0) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nlist) return;
  const int64_t q0 = pair_off[l], q1 = pair_off[l + 1];
  const int64_t x0 = list_off[l], x1 = list_off[l + 1];
  if (q1 <= q0 || x1 <= x0) return;
  const int tiles = (int)((x1 - x0 + kTileRows - 1) / kTileRows);
  const int tchunks = (tiles + kUnitTiles - 1) / kUnitTiles;
  const int qblocks = (int)((q1 - q0 + kQBlock - 1) / kQBlock);
  const int base = atomicAdd(num_units, tchunks * qblocks);
  list_unit0[l] = base;          // unit of (query block qb, tile chunk tc) of this list = base + qb * tchunks + tc
  int u = base;
  for (int qb = 0; qb < qblocks; ++qb) {
    for (int tc = 0; tc < tchunks; ++tc, ++u) {
      if (u >= max_units) return;
      IvfUnit un;
      un.qrow0 = (int)(q0 + (int64_t)qb * kQBlock);
      const int64_t rem = q1 - un.qrow0;
      un.nq = rem < kQBlock ? (int)rem : kQBlock;
      un.xrow0 = (int)(x0 + (int64_t)tc * kUnitTiles * kTileRows);
      const int tl = tiles - tc * kUnitTiles;
      un.ntiles = tl < kUnitTiles ? tl : kUnitTiles;
      un.xend = (int)x1;
      un.xlist0 = (int)x0;
      units[u] = un;
    }
  }
}

// ------------------------------------------------------------- list scan kernel ---
// Same TMA -> tcgen05 -> TMEM pipeline as scan_tc_kernel<1, DUMP>, but work units come from a
// device-built descriptor array (one unit = 128 gathered queries x <= 8 tiles of one list).
constexpr int kIvfThreads = 384;
constexpr int kIvfNS = 7, kIvfNB = 4;
constexpr int kIvfABytes = 4 * 16384, kIvfBBytes = kIvfNS * 16384;
constexpr int kIvfNBars = 2 * kIvfNS + 2 * kIvfNB + 2;
// per epilogue warp: a 32 x 32 fp32 transpose buffer, rows padded to 36 floats (conflict-free 128-bit
// accesses both ways), so that the score stores leave the SM as full 128-byte row segments
constexpr int kIvfStageRow = 36, kIvfStageBytes = 32 * kIvfStageRow * 4;
constexpr int kIvfSmem = 1024 + kIvfABytes + kIvfBBytes + kIvfNBars * 8 + 64 + 8 * kIvfStageBytes;

struct IvfScanParams {
  const IvfUnit* units;
  const int* num_units;
  int max_units;
  int d;
  uint32_t idesc;
  const int* pair_sorted;     // gathered row -> pair id
  const int64_t* pair_out;    // pair id -> offset of its score run in scorebuf
  float* scorebuf;
  int debug;                  // profiling experiments only (set_param ivf_debug): 1 = skip the score stores
  // sample = 1: only the FIRST tile of each list (units that start at the list's first row) is scored, dumped
  // into the head of every pair's run: the fused path's threshold estimate
  int sample;
  // FILTER instantiation (fused path): scores >= tau[query] are appended to a private candidate segment per
  // (unit, lane = pair, 64-column half) inside the pair's run of `pool`; nothing else reaches HBM
  const float* tau;           // [Q] per-query candidate threshold
  int nprobe;
  uint2* pool;                // [Q * smax] (score bits, stored row), same run geometry as scorebuf
  int* seg_count;             // [max_units][2][128] entries per segment
};

// tiles of a unit the scan processes (0: the unit is skipped by every warp role alike)
__device__ __forceinline__ int ivf_unit_tiles(const IvfUnit& un, int sample) {
  if (!sample) return un.ntiles;
  return un.xrow0 == un.xlist0 ? 1 : 0;
}
// rows of a unit that fall into the first 64-column half of its tiles (the second half's segment starts there)
__device__ __forceinline__ int ivf_half0_rows(const IvfUnit& un) {
  int rows = un.xend - un.xrow0;
  if (rows > un.ntiles * kTileRows) rows = un.ntiles * kTileRows;
  const int full = rows / kTileRows, rem = rows % kTileRows;
  return full * 64 + (rem < 64 ? rem : 64);
}

template <bool FILTER>
__global__ void __launch_bounds__(kIvfThreads, 1)
ivf_scan_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
                const IvfScanParams p) {
  constexpr int NS = kIvfNS, NB = kIvfNB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + kIvfABytes, bar0 = sB + kIvfBBytes;
  auto bar_full = [&](int i) { return bar0 + 8u * i; };
  auto bar_empty = [&](int i) { return bar0 + 8u * (NS + i); };
  auto bar_tfull = [&](int i) { return bar0 + 8u * (2 * NS + i); };
  auto bar_tempty = [&](int i) { return bar0 + 8u * (2 * NS + NB + i); };
  const uint32_t bar_qfull = bar0 + 8u * (2 * NS + 2 * NB);
  const uint32_t bar_qempty = bar_qfull + 8u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kIvfABytes + kIvfBBytes + kIvfNBars * 8);
  float* stage_all = reinterpret_cast<float*>(smem + kIvfABytes + kIvfBBytes + kIvfNBars * 8 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = p.d / kKChunk;
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(bar_tfull(i), 1); mbar_init(bar_tempty(i), 8); }
    mbar_init(bar_qfull, 1);
    mbar_init(bar_qempty, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  int units = *p.num_units;
  if (units > p.max_units) units = p.max_units;

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0, qe_par = 0;
      bool first = true;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const IvfUnit un = p.units[u];
        const int ntiles = ivf_unit_tiles(un, p.sample);
        if (ntiles == 0) continue;
        if (!first) { mbar_wait(bar_qempty, qe_par, 21); qe_par ^= 1; }
        first = false;
        mbar_arrive_expect_tx(bar_qfull, (uint32_t)(KC * 16384));
        for (int kc = 0; kc < KC; ++kc) tma_load_2d(sA + kc * 16384, &tmQ, kc * kKChunk, un.qrow0, bar_qfull);
        for (int t = 0; t < ntiles; ++t) {
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(bar_empty(slot), ph ^ 1, 22);
            mbar_arrive_expect_tx(bar_full(slot), 16384u);
            tma_load_2d(sB + slot * 16384, &tmX, kc * kKChunk, un.xrow0 + t * kTileRows, bar_full(slot));
            if (++slot == NS) { slot = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: uniform loop for the whole warp, one elected lane issues (see scan_tc.cu)
    {
      const uint32_t idesc = p.idesc;
      const uint64_t adesc0 = umma_desc_kmajor_sw128(sA);
      const uint64_t bdesc0 = umma_desc_kmajor_sw128(sB);
      int slot = 0, tb = 0;
      uint32_t ph = 0, tph = 0, qf_par = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int ntiles = ivf_unit_tiles(p.units[u], p.sample);
        if (ntiles == 0) continue;
        mbar_wait(bar_qfull, qf_par, 23);
        qf_par ^= 1;
        tc_fence_after_sync();
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(bar_tempty(tb), tph ^ 1, 24);
          tc_fence_after_sync();
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(bar_full(slot), ph, 25);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t d_tmem = tmem_base + (uint32_t)(tb * 128);
              const uint64_t ad = adesc0 + (uint64_t)((kc * 16384) >> 4);
              const uint64_t bd = bdesc0 + (uint64_t)((slot * 16384) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (uint32_t)((kc | k) != 0));
              umma_commit(bar_empty(slot));
            }
            __syncwarp();
            if (++slot == NS) { slot = 0; ph ^= 1; }
          }
          if (elect_one()) umma_commit(bar_tfull(tb));
          __syncwarp();
          if (++tb == NB) { tb = 0; tph ^= 1; }
        }
        // the producer waits for this before it overwrites the query block: only if this CTA has a later unit
        bool more = false;
        for (int v = u + (int)gridDim.x; v < units && !more; v += (int)gridDim.x)
          more = ivf_unit_tiles(p.units[v], p.sample) > 0;
        if (more) {
          if (elect_one()) umma_commit(bar_qempty);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // Epilogue.  A TMEM lane is a query, so a thread holds 32 consecutive scores of ITS query's run, and
    // 32 lanes hold 32 different runs: storing straight from registers sends 32 partial lines per
    // instruction (measured: 60 % of the kernel's time).  Each warp therefore transposes its 32 x 32
    // block through shared memory and stores it as 128-byte row segments, 8 lanes per query.
    const int e = warp - 4, quarter = e & 3, g = e >> 2;
    const int col_begin = g * 64;
    const int sub = lane >> 3, l8 = lane & 7;
    float* stage = stage_all + (size_t)e * (kIvfStageBytes / 4);
    constexpr uint32_t kNoRun = 0xFFFFFFFFu;
    int tb = 0;
    uint32_t tph = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const IvfUnit un = p.units[u];
      const int ntiles_u = ivf_unit_tiles(un, p.sample);
      if (ntiles_u == 0) continue;
      const int ql = quarter * 32 + lane;
      const bool warp_active = quarter * 32 < un.nq;
      // offset (in floats, < 2^32: the score buffer is capped at 6 GB) of this lane's query run
      uint32_t my_run = kNoRun;
      float my_tau = INFINITY;
      if (ql < un.nq) {
        const int pair = p.pair_sorted[un.qrow0 + ql];
        my_run = (uint32_t)p.pair_out[pair];
        if (FILTER) my_tau = p.tau[pair / p.nprobe];
      }
      // FILTER: this lane's private candidate segment for (unit, column half g) inside its pair's run
      uint2* seg = nullptr;
      int cnt = 0;
      if (FILTER && my_run != kNoRun)
        seg = p.pool + (size_t)my_run + (size_t)(un.xrow0 - un.xlist0) + (size_t)(g ? ivf_half0_rows(un) : 0);
      for (int t = 0; t < ntiles_u; ++t) {
        mbar_wait(bar_tfull(tb), tph, 26);
        tc_fence_after_sync();
        if (FILTER) {
          if (warp_active) {
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tb * 128 + col_begin);
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(taddr, r0);
            tmem_ld_32x32(taddr + 32, r1);
            tmem_ld_wait_dep(r0);
            tmem_ld_wait_dep(r1);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(tb));
            const int row0 = un.xrow0 + t * kTileRows + col_begin;
            auto scan32 = [&](const uint32_t (&r)[32], int base_row) {
              const int nvalid = un.xend - base_row;     // rows of this chunk that belong to the list
#pragma unroll
              for (int s8 = 0; s8 < 32; s8 += 8) {
                const float a = fmax3(__uint_as_float(r[s8]), __uint_as_float(r[s8 + 1]), __uint_as_float(r[s8 + 2]));
                const float b = fmax3(__uint_as_float(r[s8 + 3]), __uint_as_float(r[s8 + 4]), __uint_as_float(r[s8 + 5]));
                const float c = fmaxf(__uint_as_float(r[s8 + 6]), __uint_as_float(r[s8 + 7]));
                if (fmax3(a, b, c) >= my_tau) {
#pragma unroll
                  for (int i = s8; i < s8 + 8; ++i)
                    if (__uint_as_float(r[i]) >= my_tau && i < nvalid) {
                      seg[cnt] = make_uint2(r[i], (uint32_t)(base_row + i));
                      ++cnt;
                    }
                }
              }
            };
            if (seg) {     // my_tau = +inf for idle lanes, but NaN scores must not reach a null segment either
              scan32(r0, row0);
              scan32(r1, row0 + 32);
            }
          } else {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(tb));
          }
          if (++tb == NB) { tb = 0; tph ^= 1; }
          continue;
        }
        if (warp_active) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tb * 128 + col_begin);
          uint32_t r[32];
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait_dep(r);
            if (c == 1) {
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tempty(tb));
            }
            const int row0 = un.xrow0 + t * kTileRows + col_begin + c * 32;
            const int nvalid = un.xend - row0;                                   // real rows in this chunk
            const int npad = ((un.xend - un.xlist0 + 3) & ~3) - (row0 - un.xlist0);  // incl. the run's -inf padding
            if (npad > 0 && p.debug != 1) {
              float4* srow = reinterpret_cast<float4*>(stage + lane * kIvfStageRow);
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                srow[i >> 2] = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                           __uint_as_float(r[i + 3]));
              __syncwarp();
              const int col = 4 * l8;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int qi = it * 4 + sub;
                const uint32_t run = __shfl_sync(0xffffffffu, my_run, qi);
                if (run != kNoRun && col < npad) {
                  float4 v = *reinterpret_cast<const float4*>(stage + qi * kIvfStageRow + col);
                  if (col + 3 >= nvalid) {
                    if (col >= nvalid) v.x = -INFINITY;
                    if (col + 1 >= nvalid) v.y = -INFINITY;
                    if (col + 2 >= nvalid) v.z = -INFINITY;
                    v.w = -INFINITY;
                  }
                  // 16-byte aligned by construction (runs start on multiples of 4 floats)
                  *reinterpret_cast<float4*>(p.scorebuf + (size_t)run + (size_t)(row0 - un.xlist0) + col) = v;
                }
              }
              __syncwarp();
            }
          }
        } else {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty(tb));
        }
        if (++tb == NB) { tb = 0; tph ^= 1; }
      }
      if (FILTER && seg) p.seg_count[(size_t)u * 256 + g * 128 + ql] = cnt;
    }
  }
  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// per query: threshold = m-th largest of its score run (length row_len[q]) + compaction of the
// scores >= threshold, translating run positions to stored corpus rows
constexpr int kIvfSelThreads = 512;
__global__ void __launch_bounds__(kIvfSelThreads)
ivf_threshold_kernel(const float* __restrict__ scorebuf, int64_t smax, const int* __restrict__ row_len, int m,
                     const int64_t* __restrict__ coarse, int nprobe, const int64_t* __restrict__ list_off,
                     float* __restrict__ tau, int* __restrict__ cand_count, uint2* __restrict__ cand, int cap,
                     const float* __restrict__ qnorm, const float* __restrict__ maxnorm, float eps, int npass,
                     int sample_stride, int c_target) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_rem, s_count;
  extern __shared__ int64_t runs[];  // [nprobe] cumulative (padded) run ends, [nprobe] list starts, [nprobe] lengths
  const int q = blockIdx.x;
  const float* row = scorebuf + (size_t)q * smax;
  const int T = row_len[q];
  int64_t* run_end = runs;
  int64_t* run_x0 = runs + nprobe;
  int64_t* run_len = runs + 2 * nprobe;
  // one thread per probe for the (dependent) coarse -> list_off loads: serialised in one thread they were
  // ~20 us of every query's 28-39 us at nprobe = 32
  for (int j = threadIdx.x; j < nprobe; j += blockDim.x) {
    const int64_t l = coarse[(size_t)q * nprobe + j];
    int64_t len = 0, x0 = 0;
    if (l >= 0) {
      x0 = list_off[l];
      len = list_off[l + 1] - x0;
    }
    run_x0[j] = x0;
    run_len[j] = len;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t cum = 0;
    for (int j = 0; j < nprobe; ++j) {
      cum += (run_len[j] + 3) & ~(int64_t)3;
      run_end[j] = cum;
    }
    s_rem = m;
    s_count = 0;
  }
  __syncthreads();
  const float4* row4 = reinterpret_cast<const float4*>(row);
  const int T4 = T >> 2;

  // MSB-first radix select of the `rank`-th largest among the float4 vectors {row4[i * stride]}, over
  // `passes` bytes of the order-preserving key.  Stopping early leaves the low key bits zero = a LOWER
  // bound of that score (<= 2^-7 / 2^-15 relative below it): still a valid threshold (a few more
  // candidates), one or two sweeps less.  128-bit loads (runs are 16-byte aligned and padded to a multiple
  // of 4), two vectors in flight per thread; run-length aggregation: a query's scores share their leading
  // key bytes, so a thread's consecutive elements mostly hit the same bin -> one shared atomic per run.
  auto radix_rank = [&](int stride, int rank, int passes) -> float {
    const int n4 = T4 / stride;
    uint32_t prefix = 0, mask = 0;
    if (threadIdx.x == 0) s_rem = rank;
    for (int pass = 0; pass < passes; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      {
        int last = -1, run = 0;
        auto feed = [&](float f) {
          const uint32_t key = f2ord(f);
          if ((key & mask) == prefix) {
            const int b = (int)((key >> shift) & 255u);
            if (b == last) {
              ++run;
            } else {
              if (run) atomicAdd(&hist[last], run);
              last = b;
              run = 1;
            }
          }
        };
        for (int i = threadIdx.x; i < n4; i += 2 * blockDim.x) {
          const float4 a = row4[(size_t)i * stride];
          const int i2 = i + blockDim.x;
          float4 b = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
          const bool has_b = i2 < n4;
          if (has_b) b = row4[(size_t)i2 * stride];
          feed(a.x); feed(a.y); feed(a.z); feed(a.w);
          if (has_b) { feed(b.x); feed(b.y); feed(b.z); feed(b.w); }
        }
        if (run) atomicAdd(&hist[last], run);
      }
      __syncthreads();
      if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int part = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) part += hist[255 - (lane * 8 + b)];
        int incl = part;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const int rem = s_rem;
        __syncwarp();
        if (incl >= rem && incl - part < rem) {
          int cum = incl - part, b = 0;
          for (; b < 7; ++b) {
            const int hb = hist[255 - (lane * 8 + b)];
            if (cum + hb >= rem) break;
            cum += hb;
          }
          s_rem = rem - cum;
          s_prefix = prefix | ((uint32_t)(255 - (lane * 8 + b)) << shift);
        }
      }
      __syncthreads();
      prefix = s_prefix;
      mask |= 255u << shift;
    }
    return ord2f(prefix);
  };

  // one sweep: compact the scores >= t into this query's candidate list; returns how many there were
  auto emit_all = [&](float t) -> int {
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    auto emit = [&](float v, int i) {
      if (v >= t && v > -INFINITY) {
        const int slot = atomicAdd(&s_count, 1);
        if (slot < cap) {
          int j = 0;
          while (j < nprobe - 1 && i >= run_end[j]) ++j;
          const int64_t start = j > 0 ? run_end[j - 1] : 0;
          cand[(size_t)q * cap + slot] = make_uint2(__float_as_uint(v), (uint32_t)(run_x0[j] + (i - start)));
        }
      }
    };
    // two vectors in flight per thread (this sweep is the first touch of the run: HBM latency-bound)
    for (int i = threadIdx.x; i < T4; i += 2 * blockDim.x) {
      const float4 a = row4[i];
      const int i2 = i + blockDim.x;
      const bool has_b = i2 < T4;
      float4 b = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (has_b) b = row4[i2];
      emit(a.x, 4 * i); emit(a.y, 4 * i + 1); emit(a.z, 4 * i + 2); emit(a.w, 4 * i + 3);
      if (has_b) { emit(b.x, 4 * i2); emit(b.y, 4 * i2 + 1); emit(b.z, 4 * i2 + 2); emit(b.w, 4 * i2 + 3); }
    }
    __syncthreads();
    return s_count;
  };

  // Sampled threshold: the rank-(c_target / stride) score of every stride-th vector estimates the score of
  // rank c_target (2.5 k) of the whole run; one sweep then collects ~c_target candidates, a superset of
  // the top-k and of its rescore window with a margin of several sigma.  If the count comes out short
  // (or over the buffer) this query falls back to the exact radix passes below -- so a sampled threshold
  // never changes the answer, it only saves sweeps.
  float t = -INFINITY;
  bool done = false;
  if (m >= T) {
    emit_all(t);
    done = true;
  } else if (sample_stride >= 2) {
    const int rank_s = (c_target + sample_stride - 1) / sample_stride;
    if ((int64_t)(T4 / sample_stride) * 4 >= (int64_t)rank_s * 8) {
      t = radix_rank(sample_stride, rank_s, 4);
      const int got = emit_all(t);
      done = got >= m + (c_target - m) / 4 && got <= cap;
    }
  }
  if (!done) {
    t = radix_rank(1, m, npass);
    if (qnorm) t -= rescore_margin(eps, qnorm[q], *maxnorm);   // candidates = the provable rescore window
    emit_all(t);
  }
  if (threadIdx.x == 0) tau[q] = t;
  __syncthreads();
  if (threadIdx.x == 0) cand_count[q] = s_count;
}

// ------------------------------------------------------------ fused path: threshold + gather ---
// Threshold of the fused list scan, one CTA per query.  The sample pass scored the first `S` rows of every probed
// list (a list's rows are in insertion order: a random sample of the list).  Lists differ in length AND in how
// close they are to the query, so every sampled score of list j stands for len_j / s_j rows: tau = the score at
// which the WEIGHTED count of sampled scores reaches c_target (the candidates the filter scan should pass).
// When the query scans no more rows than the candidate buffer holds, tau = -inf (everything is a candidate).
constexpr int kIvfTauThreads = 512;
constexpr int kIvfTauMaxKeys = 8192;
constexpr int kIvfTauWShift = 2;     // fixed-point weights (rows per sampled row, quarter-row steps): 32-bit integer histograms
__global__ void __launch_bounds__(kIvfTauThreads)
ivf_sample_tau_kernel(const float* __restrict__ scorebuf, int64_t smax, const int64_t* __restrict__ coarse, int nprobe,
                      const int64_t* __restrict__ list_off, int S, int c_target, int take_all_below,
                      float* __restrict__ tau) {
  extern __shared__ __align__(16) uint8_t tsm[];
  const int max_keys = nprobe * S;                                                // <= kIvfTauMaxKeys (planner)
  uint32_t* keys = reinterpret_cast<uint32_t*>(tsm);                              // [ns] order-preserving score keys
  uint32_t* wk = keys + max_keys;                                                 // [ns] weight of each key
  uint32_t* wgt = wk + max_keys;                                                  // [nprobe] weight of a sample of list j
  int* sstart = reinterpret_cast<int*>(wgt + nprobe);                             // [nprobe + 1] prefix of sample sizes
  int* rstart = sstart + nprobe + 1;                                              // [nprobe] run start (floats)
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_rem, s_prefix;
  __shared__ int s_total_rows, s_done;
  const int q = blockIdx.x, tid = threadIdx.x;
  // list lengths: one thread per probe (the loads are two dependent round trips each - never serialise them)
  for (int j = tid; j < nprobe; j += blockDim.x) {
    const int64_t l = coarse[(size_t)q * nprobe + j];
    const int len = l >= 0 ? (int)(list_off[l + 1] - list_off[l]) : 0;
    const int sj = len < S ? len : S;
    rstart[j] = len;
    wgt[j] = sj > 0 ? (uint32_t)(((uint32_t)len << kIvfTauWShift) / (uint32_t)sj) : 0u;
  }
  __syncthreads();
  if (tid == 0) {
    int64_t cum = 0, rows = 0;
    int ns = 0;
    for (int j = 0; j < nprobe; ++j) {
      const int len = rstart[j];
      sstart[j] = ns;
      rstart[j] = (int)cum;
      ns += len < S ? len : S;
      rows += len;
      cum += (len + 3) & ~3;
    }
    sstart[nprobe] = ns;
    s_total_rows = (int)(rows < 0x7fffffff ? rows : 0x7fffffff);
  }
  __syncthreads();
  const int ns = sstart[nprobe];
  if (s_total_rows <= take_all_below || ns == 0) {
    if (tid == 0) tau[q] = -INFINITY;
    return;
  }
  const float* row = scorebuf + (size_t)q * smax;
  for (int i = tid; i < ns; i += blockDim.x) {
    int lo = 0, hi = nprobe - 1;          // largest j with sstart[j] <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sstart[mid] <= i) lo = mid; else hi = mid - 1;
    }
    keys[i] = f2ord(row[rstart[lo] + (i - sstart[lo])]);
    wk[i] = wgt[lo];
  }
  // at least ~32 sampled scores above the threshold (a rank-3 estimate would scatter the candidate count by
  // +-60 %), never more candidates than 5/7 of what the buffer takes
  float target = fmaxf((float)c_target, 32.f * (float)s_total_rows / (float)ns);
  target = fminf(target, (float)(take_all_below * 5 / 7));
  if (tid == 0) {
    s_rem = (uint32_t)(target * (float)(1 << kIvfTauWShift));
    s_prefix = 0;
    s_done = 0;
  }
  __syncthreads();
  // MSB-first WEIGHTED radix select: the key at which the cumulative weight (from the top) reaches the target.
  // Three byte passes: the low byte stays zero = a lower bound 2^-15 relative below that score (a few more
  // candidates, one pass less).
  uint32_t prefix = 0, mask = 0;
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    // a query's scores share their leading key bytes, so in the first pass most lanes of a warp hit the SAME bin:
    // two warp-aggregated rounds serve the crowded bins, lanes left after that add on their own
    const int ns_pad = (ns + 31) & ~31;
    for (int i = tid; i < ns_pad; i += blockDim.x) {
      uint32_t key = 0, w = 0;
      bool on = false;
      if (i < ns) {
        key = keys[i];
        on = (key & mask) == prefix;
        w = wk[i];
      }
      const int bin = (int)((key >> shift) & 255u);
      unsigned active = __ballot_sync(0xffffffffu, on);
#pragma unroll 1
      for (int round = 0; round < 2 && active; ++round) {
        const int leader = __ffs(active) - 1;
        const int lbin = __shfl_sync(0xffffffffu, bin, leader);
        const bool mine = on && bin == lbin;
        const uint32_t sum = __reduce_add_sync(0xffffffffu, mine ? w : 0u);
        if ((tid & 31) == leader) atomicAdd(&hist[lbin], sum);
        if (mine) on = false;
        active &= ~__ballot_sync(0xffffffffu, mine);
      }
      if (on) atomicAdd(&hist[bin], w);
    }
    __syncthreads();
    if (tid < 32) {          // bin where the cumulative weight from the TOP reaches s_rem (8 bins per lane)
      const int lane = tid;
      uint32_t part = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) part += hist[255 - (lane * 8 + b)];
      uint32_t incl = part;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t rem = s_rem;
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      __syncwarp();
      if (total < rem) {
        if (lane == 0) s_done = 1;       // the whole (remaining) sample weighs less than the target: take everything
      } else if (incl >= rem && incl - part < rem) {
        uint32_t cum = incl - part;
        int b = 0;
        for (; b < 7; ++b) {
          const uint32_t hb = hist[255 - (lane * 8 + b)];
          if (cum + hb >= rem) break;
          cum += hb;
        }
        s_rem = rem - cum;
        s_prefix = prefix | ((uint32_t)(255 - (lane * 8 + b)) << shift);
      }
    }
    __syncthreads();
    prefix = s_prefix;
    mask |= 255u << shift;
    if (s_done) break;
  }
  if (tid == 0) tau[q] = s_done ? -INFINITY : ord2f(prefix);
}

// Collects a query's candidates from the private segments the filter scan wrote -- one per (unit, pair lane,
// column half), inside the pair's run of `pool` -- into its single candidate list for select_rescore.
constexpr int kIvfGatherThreads = 256;
constexpr int kIvfGatherMaxSub = 4096;
__global__ void __launch_bounds__(kIvfGatherThreads)
ivf_gather_kernel(const int64_t* __restrict__ coarse, int nprobe, const int64_t* __restrict__ list_off,
                  const int64_t* __restrict__ pairoff, const int* __restrict__ pair_pos,
                  const int* __restrict__ list_unit0, int unit_tiles, const uint2* __restrict__ pool, int64_t smax,
                  const int* __restrict__ seg_count, uint2* __restrict__ cand, int* __restrict__ cand_count, int cap) {
  extern __shared__ __align__(16) uint8_t gsm[];
  int* sub0 = reinterpret_cast<int*>(gsm);              // [nprobe + 1] first sub-segment of pair j
  int* rstart = sub0 + nprobe + 1;                      // [nprobe] run start of pair j (entries, relative to the query)
  int* cnt = rstart + nprobe;                           // [kIvfGatherMaxSub + 1] -> exclusive prefix
  uint32_t* src = reinterpret_cast<uint32_t*>(cnt + kIvfGatherMaxSub + 1);   // [kIvfGatherMaxSub] segment start
  __shared__ int s_nsub, s_total;
  const int q = blockIdx.x, tid = threadIdx.x;
  for (int j = tid; j < nprobe; j += blockDim.x) {     // one thread per probe: the loads are dependent round trips
    const int64_t l = coarse[(size_t)q * nprobe + j];
    rstart[j] = l >= 0 ? (int)(list_off[l + 1] - list_off[l]) : 0;
  }
  __syncthreads();
  if (tid == 0) {
    int64_t cum = 0;
    int ns = 0;
    for (int j = 0; j < nprobe; ++j) {
      const int64_t len = rstart[j];
      sub0[j] = ns;
      rstart[j] = (int)cum;
      const int tiles = (int)((len + kTileRows - 1) / kTileRows);
      ns += 2 * ((tiles + unit_tiles - 1) / unit_tiles);
      cum += (len + 3) & ~(int64_t)3;
    }
    sub0[nprobe] = ns;
    s_nsub = ns;
  }
  __syncthreads();
  const int nsub = s_nsub;
  if (nsub > kIvfGatherMaxSub) {       // cannot happen with the planner's limits; reported as an overflow
    if (tid == 0) cand_count[q] = cap + 1;
    return;
  }
  for (int ss = tid; ss < nsub; ss += blockDim.x) {
    int lo = 0, hi = nprobe - 1;       // pair j of this sub-segment
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sub0[mid] <= ss) lo = mid; else hi = mid - 1;
    }
    const int j = lo, rel = ss - sub0[j], tc = rel >> 1, h = rel & 1;
    const int64_t l = coarse[(size_t)q * nprobe + j];
    const int64_t x0 = list_off[l], x1 = list_off[l + 1];
    const int tiles = (int)((x1 - x0 + kTileRows - 1) / kTileRows);
    const int tchunks = (tiles + unit_tiles - 1) / unit_tiles;
    const int pos = pair_pos[(size_t)q * nprobe + j] - (int)pairoff[l];
    const int u = list_unit0[l] + (pos / kQBlock) * tchunks + tc;
    IvfUnit un;
    un.xrow0 = (int)(x0 + (int64_t)tc * unit_tiles * kTileRows);
    const int tl = tiles - tc * unit_tiles;
    un.ntiles = tl < unit_tiles ? tl : unit_tiles;
    un.xend = (int)x1;
    un.xlist0 = (int)x0;
    cnt[ss] = seg_count[(size_t)u * 256 + h * 128 + (pos % kQBlock)];
    src[ss] = (uint32_t)(rstart[j] + (un.xrow0 - un.xlist0) + (h ? ivf_half0_rows(un) : 0));
  }
  __syncthreads();
  if (tid < 32) {                      // exclusive scan of the counts, 32 at a time
    int carry = 0;
    for (int b = 0; b < nsub; b += 32) {
      const int idx = b + tid;
      const int c = idx < nsub ? cnt[idx] : 0;
      int v = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (tid >= o) v += t;
      }
      if (idx < nsub) cnt[idx] = carry + v - c;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
    if (tid == 0) { cnt[nsub] = carry; s_total = carry; }
  }
  __syncthreads();
  const int total = s_total;
  const int n = total < cap ? total : cap;
  const uint2* qpool = pool + (size_t)q * smax;
  for (int i = tid; i < n; i += blockDim.x) {   // one thread per candidate: every load of the gather in flight at once
    int lo = 0, hi = nsub - 1;                  // largest ss with cnt[ss] <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (cnt[mid] <= i) lo = mid; else hi = mid - 1;
    }
    cand[(size_t)q * cap + i] = qpool[src[lo] + (uint32_t)(i - cnt[lo])];
  }
  if (tid == 0) cand_count[q] = total;
}

// --------------------------------------------------------------------- host side ---
struct DevBuf {  // RAII cudaMalloc scratch (build-time paths only; search never allocates)
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

// scoped override of a flat index's path choice (0 auto, 1 dense, 2 filter)
struct ForcePath {
  b2r_index* h;
  int keep;
  ForcePath(b2r_index* h_, int v) : h(h_), keep(h_->force_path) { h->force_path = v; }
  ~ForcePath() { h->force_path = keep; }
};

__global__ void or_status_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] |= src[i];
}

// exact top-1 (or top-k) centroid of every row of x through the nested flat quantiser
int quantizer_assign(b2r_index* h, int64_t n, const float* x, int normalize, int k, int64_t* labels,
                     cudaStream_t stream) {
  b2r_index* qz = h->quantizer;
  const int chunk = 65536;
  const int qmax = (int)(n < chunk ? n : chunk);
  const size_t ws_bytes = flat_search_workspace(qz, qmax, k);
  DevBuf ws, dist, st, tr;
  int rc;
  if ((rc = ws.alloc(ws_bytes))) return rc;
  if ((rc = dist.alloc((size_t)qmax * k * 4))) return rc;
  if ((rc = st.alloc((size_t)qmax * 4))) return rc;
  if ((rc = tr.alloc((size_t)qmax * 4))) return rc;
  std::vector<int32_t> sth(qmax);
  for (int64_t i0 = 0; i0 < n; i0 += chunk) {
    const int qc = (int)((n - i0) < chunk ? (n - i0) : chunk);
    rc = flat_search(qz, qc, x + (size_t)i0 * h->d, normalize, k, dist.as<float>(), labels + (size_t)i0 * k,
                     st.as<int32_t>(), tr.as<float>(), nullptr, ws.p, ws_bytes, stream);
    if (rc) return rc;
    // exactness of the assignment matters (list membership): re-run flagged rows with the
    // threshold the kernel suggests until none is flagged
    for (int attempt = 0; attempt < 8; ++attempt) {
      B2R_CUDA(cudaMemcpyAsync(sth.data(), st.p, (size_t)qc * 4, cudaMemcpyDeviceToHost, stream));
      B2R_CUDA(cudaStreamSynchronize(stream));
      bool any = false;
      for (int i = 0; i < qc; ++i) any |= (sth[i] & (B2R_ST_TOO_FEW | B2R_ST_NEED_LOWER_TAU | B2R_ST_CAND_OVERFLOW)) != 0;
      if (!any) break;
      // whole-chunk retry with the per-row suggested thresholds (rows that were fine keep theirs)
      DevBuf tau;
      if ((rc = tau.alloc((size_t)qc * 4))) return rc;
      B2R_CUDA(cudaMemcpyAsync(tau.p, tr.p, (size_t)qc * 4, cudaMemcpyDeviceToDevice, stream));
      rc = flat_search(qz, qc, x + (size_t)i0 * h->d, normalize, k, dist.as<float>(), labels + (size_t)i0 * k,
                       st.as<int32_t>(), tr.as<float>(), tau.as<float>(), ws.p, ws_bytes, stream);
      if (rc) return rc;
      B2R_CUDA(cudaStreamSynchronize(stream));
    }
  }
  return B2R_OK;
}

int set_centroids(b2r_index* h, const float* cent_dev, cudaStream_t stream) {
  b2r_index* qz = h->quantizer;
  qz->ntotal = 0;
  B2R_CUDA(cudaMemsetAsync(qz->maxnorm, 0, 256, stream));
  int rc = flat_add(qz, h->nlist, cent_dev, 0, stream);
  if (rc) return rc;
  if (!h->list_off) {
    if (cudaMalloc(&h->list_off, (size_t)(h->nlist + 1) * 8) != cudaSuccess) return fail(B2R_ENOMEM, "cudaMalloc list_off");
    B2R_CUDA(cudaMemsetAsync(h->list_off, 0, (size_t)(h->nlist + 1) * 8, stream));
  }
  h->list_sizes_host.assign(h->nlist, 0);
  h->list_sizes_desc_for = -1;
  h->trained = (h->kind != B2R_KIND_IVF_PQ) || h->pq_trained;
  return pq_build_list_tables(h, stream);   // no-op unless this is an IVF-PQ index with codebooks
}

}  // namespace

void ivf_free(b2r_index* h) {
  cudaFree(h->list_off);
  cudaFree(h->row_list);
  cudaFree(h->perm);
  cudaFree(h->codebooks);
  cudaFree(h->codes);
  cudaFree(h->pq_list_tab);
  cudaFree(h->pq_row_term);
  h->list_off = nullptr;
}

int ivf_train(b2r_index* h, int64_t n, const float* x, uint64_t seed, cudaStream_t stream) {
  const int d = h->d, nlist = h->nlist;
  if (n < nlist) return fail(B2R_EINVAL, "index_train: fewer training vectors (" + std::to_string(n) +
                                             ") than centroids (" + std::to_string(nlist) + ")");
  // subsample to <= 256 points per centroid with a seeded permutation (faiss: seed 1234)
  std::mt19937_64 rng(seed);
  int64_t nt = n;
  std::vector<int64_t> pick;
  const int64_t cap = (int64_t)kMaxPointsPerCentroid * nlist;
  {
    std::vector<int64_t> perm(n);
    for (int64_t i = 0; i < n; ++i) perm[i] = i;
    // partial Fisher-Yates: the first max(nt, nlist) entries are a uniform sample without replacement
    nt = n > cap ? cap : n;
    for (int64_t i = 0; i < nt; ++i) {
      std::uniform_int_distribution<int64_t> u(i, n - 1);
      std::swap(perm[i], perm[u(rng)]);
    }
    pick.assign(perm.begin(), perm.begin() + nt);
  }
  DevBuf rows_dev, xt, cent, sums, counts, assign;
  int rc;
  if ((rc = rows_dev.alloc((size_t)nt * 8))) return rc;
  if ((rc = xt.alloc((size_t)nt * d * 4))) return rc;
  if ((rc = cent.alloc((size_t)nlist * d * 4))) return rc;
  if ((rc = sums.alloc((size_t)nlist * d * 4))) return rc;
  if ((rc = counts.alloc((size_t)nlist * 4))) return rc;
  if ((rc = assign.alloc((size_t)nt * 8))) return rc;
  B2R_CUDA(cudaMemcpyAsync(rows_dev.p, pick.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, stream));
  gather_rows_f32_kernel<<<(unsigned)ceil_div(nt, 8), 256, 0, stream>>>(x, rows_dev.as<int64_t>(), nt, d, xt.as<float>());
  B2R_CHECK_LAUNCH("gather_rows_f32_kernel");
  // initial centroids: the first nlist sampled points (distinct rows of a random permutation),
  // L2-normalised for the IP metric (spherical k-means)
  const int spherical = (h->metric == B2R_METRIC_IP && h->kind == B2R_KIND_IVF_FLAT) ? 1 : 0;
  if ((rc = launch_fill_i32(counts.as<int>(), nlist, 1, stream))) return rc;
  kmeans_finalize_kernel<<<(unsigned)ceil_div(nlist, 8), 256, 0, stream>>>(
      xt.as<float>(), counts.as<int>(), xt.as<float>(), nt, d, nlist, -1, spherical, cent.as<float>());
  B2R_CHECK_LAUNCH("kmeans_finalize_kernel(init)");
  for (int it = 0; it < kKmeansIters; ++it) {
    if ((rc = set_centroids(h, cent.as<float>(), stream))) return rc;
    if ((rc = quantizer_assign(h, nt, xt.as<float>(), 0, 1, assign.as<int64_t>(), stream))) return rc;
    B2R_CUDA(cudaMemsetAsync(sums.p, 0, (size_t)nlist * d * 4, stream));
    B2R_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)nlist * 4, stream));
    kmeans_accumulate_kernel<<<(unsigned)ceil_div(nt, 8), 256, 0, stream>>>(
        xt.as<float>(), assign.as<int64_t>(), nt, d, sums.as<float>(), counts.as<int>());
    B2R_CHECK_LAUNCH("kmeans_accumulate_kernel");
    kmeans_finalize_kernel<<<(unsigned)ceil_div(nlist, 8), 256, 0, stream>>>(
        sums.as<float>(), counts.as<int>(), xt.as<float>(), nt, d, nlist, it, spherical, cent.as<float>());
    B2R_CHECK_LAUNCH("kmeans_finalize_kernel");
  }
  if ((rc = set_centroids(h, cent.as<float>(), stream))) return rc;
  if (h->kind == B2R_KIND_IVF_PQ) {
    // product quantiser on the residuals of (a prefix of) the same random training sample
    const int64_t np_train = nt < 65536 ? nt : 65536;
    DevBuf resid;
    if ((rc = resid.alloc((size_t)np_train * d * 4))) return rc;
    if ((rc = quantizer_assign(h, np_train, xt.as<float>(), 0, 1, assign.as<int64_t>(), stream))) return rc;
    if ((rc = pq_residuals(h, np_train, xt.as<float>(), assign.as<int64_t>(), resid.as<float>(), stream))) return rc;
    if ((rc = pq_train(h, np_train, resid.as<float>(), seed, stream))) return rc;
    h->trained = true;
  }
  B2R_CUDA(cudaStreamSynchronize(stream));
  return B2R_OK;
}

// merge n new code rows (with their list ids) into the list-sorted code storage
static int store_codes(b2r_index* h, int64_t n, const uint8_t* codes_new, const int64_t* assign,
                       const int64_t* new_off, unsigned long long* cursor, cudaStream_t stream) {
  const int nlist = h->nlist, m = h->pq_m;
  const int64_t n_old = h->ntotal, n_all = n_old + n;
  int rc;
  DevBuf dst_old, dst_new;
    uint8_t* ncodes = nullptr;
    int32_t* nlistid = nullptr;
    uint32_t* nperm = nullptr;
    const int64_t cap = (int64_t)align_up((size_t)n_all, 1024);
    if (cudaMalloc(&ncodes, (size_t)cap * m + 64) != cudaSuccess ||   /* +64: bulk copies read 16-byte granules */ cudaMalloc(&nlistid, (size_t)cap * 4) != cudaSuccess ||
        cudaMalloc(&nperm, (size_t)cap * 4) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(ncodes); cudaFree(nlistid); cudaFree(nperm);
      return fail(B2R_ENOMEM, "index_add: cudaMalloc of the code storage failed");
    }
    if (n_old > 0) {
      if ((rc = dst_old.alloc((size_t)n_old * 8))) return rc;
      place_old_rows_kernel<<<(unsigned)ceil_div(n_old, 256), 256, 0, stream>>>(
          h->list_off, new_off, nlist, n_old, h->row_list, dst_old.as<int64_t>());
      B2R_CHECK_LAUNCH("place_old_rows_kernel");
      if ((rc = pq_scatter_codes(h->codes, dst_old.as<int64_t>(), n_old, m, h->row_list, nullptr, h->perm, 0u, ncodes,
                                 nlistid, nperm, stream)))
        return rc;
    }
    if ((rc = dst_new.alloc((size_t)n * 8))) return rc;
    place_new_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(
        assign, n, cursor, dst_new.as<int64_t>());
    B2R_CHECK_LAUNCH("place_new_rows_kernel");
    if ((rc = pq_scatter_codes(codes_new, dst_new.as<int64_t>(), n, m, nullptr, assign,
                               nullptr, (uint32_t)n_old, ncodes, nlistid, nperm, stream)))
      return rc;
    B2R_CUDA(cudaMemcpyAsync(h->list_off, new_off, (size_t)(nlist + 1) * 8, cudaMemcpyDeviceToDevice, stream));
    std::vector<int64_t> off_host(nlist + 1);
    B2R_CUDA(cudaMemcpyAsync(off_host.data(), new_off, (size_t)(nlist + 1) * 8, cudaMemcpyDeviceToHost, stream));
    B2R_CUDA(cudaStreamSynchronize(stream));
    cudaFree(h->codes); cudaFree(h->row_list); cudaFree(h->perm);
    h->codes = ncodes; h->row_list = nlistid; h->perm = nperm;
    h->capacity = cap;
    h->ntotal = n_all;
    h->list_sizes_host.resize(nlist);
    for (int l = 0; l < nlist; ++l) h->list_sizes_host[l] = off_host[l + 1] - off_host[l];
    h->list_sizes_desc_for = -1;
    return pq_update_row_terms(h, stream);
}

int ivf_add(b2r_index* h, int64_t n, const float* x, int normalize, cudaStream_t stream) {
  const int d = h->d, nlist = h->nlist;
  const int64_t n_old = h->ntotal, n_all = n_old + n;
  if (n_all > 0x7FFFFF00ll) return fail(B2R_EUNSUPPORTED, "index_add: more than 2^31 rows per shard");
  int rc;
  const bool is_pq = h->kind == B2R_KIND_IVF_PQ;
  // 1. normalised fp32 (+ bf16 scan copy for IVF-Flat) of the new rows
  DevBuf t32, t16, assign, hist, new_off, cursor, dst_old, dst_new;
  if ((rc = t32.alloc((size_t)n * d * 4))) return rc;
  if (!is_pq && (rc = t16.alloc((size_t)n * d * 2))) return rc;
  // scan format: fp16 while every stored row was normalised on ingest (|x_i| <= 1), else bf16
  int want16 = h->scan_dtype_req >= 0 ? h->scan_dtype_req : (normalize ? 1 : 0);
  if (h->ntotal > 0 && h->scan_fp16 == 0 && h->scan_dtype_req < 0) want16 = 0;   // once bf16, stay bf16
  if (!is_pq && h->ntotal > 0 && h->scan_fp16 >= 0 && h->scan_fp16 != want16) {
    if ((rc = launch_reencode(h->x32, h->ntotal, d, h->x16, want16, stream))) return rc;
  }
  h->scan_fp16 = want16;
  if ((rc = launch_ingest(x, n, d, normalize, t32.as<float>(), is_pq ? nullptr : t16.as<__nv_bfloat16>(), want16,
                          h->maxnorm, stream)))
    return rc;
  // 2. list assignment: exact max-inner-product centroid (faiss quantizer->assign)
  if ((rc = assign.alloc((size_t)n * 8))) return rc;
  if ((rc = quantizer_assign(h, n, t32.as<float>(), 0, 1, assign.as<int64_t>(), stream))) return rc;
  // 3. merged list offsets (old rows of a list stay first, new rows are appended behind them)
  if ((rc = hist.alloc((size_t)nlist * 4))) return rc;
  if ((rc = new_off.alloc((size_t)(nlist + 1) * 8))) return rc;
  if ((rc = cursor.alloc((size_t)nlist * 8))) return rc;
  B2R_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)nlist * 4, stream));
  hist_lists_kernel<<<296, 256, 0, stream>>>(assign.as<int64_t>(), n, nlist, hist.as<int>());
  B2R_CHECK_LAUNCH("hist_lists_kernel");
  scan_lists_kernel<<<1, 1024, 0, stream>>>(hist.as<int>(), n_old > 0 ? h->list_off : nullptr, nlist,
                                            new_off.as<int64_t>(), cursor.as<int64_t>());
  B2R_CHECK_LAUNCH("scan_lists_kernel");
  if (is_pq) {
    // 4'. IVF-PQ stores only the codes of the residuals
    const int m = h->pq_m;
    DevBuf resid, ncodes_tmp;
    if ((rc = resid.alloc((size_t)n * d * 4))) return rc;
    if ((rc = ncodes_tmp.alloc((size_t)n * m))) return rc;
    if ((rc = pq_residuals(h, n, t32.as<float>(), assign.as<int64_t>(), resid.as<float>(), stream))) return rc;
    if ((rc = pq_encode(h, n, resid.as<float>(), ncodes_tmp.as<uint8_t>(), stream))) return rc;
    return store_codes(h, n, ncodes_tmp.as<uint8_t>(), assign.as<int64_t>(), new_off.as<int64_t>(),
                       cursor.as<unsigned long long>(), stream);
  }
  // 4. new storage, rows scattered to their list positions
  float* n32 = nullptr;
  __nv_bfloat16* n16 = nullptr;
  int32_t* nlistid = nullptr;
  uint32_t* nperm = nullptr;
  const int64_t cap = (int64_t)align_up((size_t)n_all, 1024);
  if (cudaMalloc(&n32, (size_t)cap * d * 4) != cudaSuccess || cudaMalloc(&n16, (size_t)cap * d * 2) != cudaSuccess ||
      cudaMalloc(&nlistid, (size_t)cap * 4) != cudaSuccess || cudaMalloc(&nperm, (size_t)cap * 4) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(n32); cudaFree(n16); cudaFree(nlistid); cudaFree(nperm);
    return fail(B2R_ENOMEM, "index_add: cudaMalloc of the list-sorted corpus failed");
  }
  if (n_old > 0) {
    if ((rc = dst_old.alloc((size_t)n_old * 8))) return rc;
    place_old_rows_kernel<<<(unsigned)ceil_div(n_old, 256), 256, 0, stream>>>(
        h->list_off, new_off.as<int64_t>(), nlist, n_old, h->row_list, dst_old.as<int64_t>());
    B2R_CHECK_LAUNCH("place_old_rows_kernel");
    scatter_rows_kernel<<<(unsigned)ceil_div(n_old, 8), 256, 0, stream>>>(
        h->x32, h->x16, dst_old.as<int64_t>(), n_old, d, h->row_list, nullptr, h->perm, 0u, n32, n16, nlistid, nperm);
    B2R_CHECK_LAUNCH("scatter_rows_kernel(old)");
  }
  if ((rc = dst_new.alloc((size_t)n * 8))) return rc;
  place_new_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(
      assign.as<int64_t>(), n, cursor.as<unsigned long long>(), dst_new.as<int64_t>());
  B2R_CHECK_LAUNCH("place_new_rows_kernel");
  scatter_rows_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, stream>>>(
      t32.as<float>(), t16.as<__nv_bfloat16>(), dst_new.as<int64_t>(), n, d, nullptr, assign.as<int64_t>(), nullptr,
      (uint32_t)n_old, n32, n16, nlistid, nperm);
  B2R_CHECK_LAUNCH("scatter_rows_kernel(new)");
  B2R_CUDA(cudaMemcpyAsync(h->list_off, new_off.p, (size_t)(nlist + 1) * 8, cudaMemcpyDeviceToDevice, stream));
  std::vector<int64_t> off_host(nlist + 1);
  B2R_CUDA(cudaMemcpyAsync(off_host.data(), new_off.p, (size_t)(nlist + 1) * 8, cudaMemcpyDeviceToHost, stream));
  B2R_CUDA(cudaStreamSynchronize(stream));
  cudaFree(h->x32); cudaFree(h->x16); cudaFree(h->row_list); cudaFree(h->perm);
  h->x32 = n32; h->x16 = n16; h->row_list = nlistid; h->perm = nperm;
  h->capacity = cap;
  h->ntotal = n_all;
  h->list_sizes_host.resize(nlist);
  for (int l = 0; l < nlist; ++l) h->list_sizes_host[l] = off_host[l + 1] - off_host[l];
  h->list_sizes_desc_for = -1;
  return make_tmap_bf16_rows(&h->tmX, h->x16, h->ntotal, d);
}

int ivf_add_codes(b2r_index* h, int64_t n, const uint8_t* codes, const int64_t* list_ids, cudaStream_t stream) {
  const int nlist = h->nlist;
  DevBuf hist, new_off, cursor;
  int rc;
  if ((rc = hist.alloc((size_t)nlist * 4))) return rc;
  if ((rc = new_off.alloc((size_t)(nlist + 1) * 8))) return rc;
  if ((rc = cursor.alloc((size_t)nlist * 8))) return rc;
  B2R_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)nlist * 4, stream));
  hist_lists_kernel<<<296, 256, 0, stream>>>(list_ids, n, nlist, hist.as<int>());
  B2R_CHECK_LAUNCH("hist_lists_kernel");
  scan_lists_kernel<<<1, 1024, 0, stream>>>(hist.as<int>(), h->ntotal > 0 ? h->list_off : nullptr, nlist,
                                            new_off.as<int64_t>(), cursor.as<int64_t>());
  B2R_CHECK_LAUNCH("scan_lists_kernel");
  return store_codes(h, n, codes, list_ids, new_off.as<int64_t>(), cursor.as<unsigned long long>(), stream);
}

// ---- search -------------------------------------------------------------------------
namespace {
struct IvfPlan {
  int nprobe = 1, chunk = 0, qpad = 0, max_units = 0, cap = 4096, c_target = 0, sample_stride = 1;
  int64_t smax = 0, pairs_pad = 0;
  // fused path (IVF-Flat): sample pass -> per-query threshold -> filter scan -> candidate gather
  bool fused = false;
  int sample_rows = 0;      // rows of every probed list the sample pass scores
  size_t off_pair_pos = 0, off_list_unit0 = 0, off_seg_count = 0;
  size_t off_cstatus = 0;
  size_t off_q16, off_q32, off_qnorm, off_coarse, off_cdist, off_pair_out, off_row_len, off_listcnt, off_pairoff,
      off_cursor, off_pair_sorted, off_gq16, off_units, off_nunits, off_tau, off_count, off_cand, off_score,
      off_qws, off_qtab, qws_bytes, total;
};

IvfPlan make_ivf_plan(const b2r_index* h, int q, int k, int nprobe) {
  IvfPlan pl;
  if (nprobe < 1) nprobe = 1;
  if (nprobe > h->nlist) nprobe = h->nlist;
  pl.nprobe = nprobe;
  // worst case rows scanned by one query = the nprobe largest lists
  // worst-case bound needs the list sizes in descending order: sorted once per corpus state, not per call
  // (at nlist = 4096 the sort alone was ~100 us of host time in front of every search)
  if (h->list_sizes_desc_for != h->ntotal || h->list_sizes_desc.size() != h->list_sizes_host.size()) {
    h->list_sizes_desc = h->list_sizes_host;
    std::sort(h->list_sizes_desc.begin(), h->list_sizes_desc.end(), std::greater<int64_t>());
    h->list_sizes_desc_for = h->ntotal;
  }
  const std::vector<int64_t>& sz = h->list_sizes_desc;
  int64_t smax = 0;
  for (int i = 0; i < nprobe && i < (int)sz.size(); ++i) smax += (sz[i] + 3) & ~(int64_t)3;
  if (smax < 4) smax = 4;
  pl.smax = (int64_t)align_up((size_t)smax, 4);
  int ct = (int)llround(h->cand_factor * k);
  if (ct < 64) ct = 64;
  if (ct > 2560) ct = 2560;
  if (ct < k) ct = k;
  pl.c_target = ct;
  // sampled threshold (ivf_threshold_kernel): every stride-th score vector, sample rank >= 64
  pl.sample_stride = 1;
  while (pl.sample_stride < 16 && ct / (pl.sample_stride * 2) >= 64) pl.sample_stride *= 2;
  if (h->ivf_sample == 0) pl.sample_stride = 1;
  // units <= sum over lists ceil(nq/128) * ceil(tiles/8); bound: (pairs/128 + nlist) query blocks,
  // each with <= ceil(max_list_tiles / 8) tile chunks
  const int64_t max_tiles = sz.empty() ? 1 : ceil_div(sz[0] > 0 ? sz[0] : 1, kTileRows);
  // fused path: the filter scan appends candidates to private segments inside the pair runs (8 bytes per slot
  // instead of 4 per dumped score, touched only where a candidate lands).  Needs the sample of every probed list
  // to fit the threshold kernel's sort buffer and the query's segment list to fit the gather kernel's.
  pl.sample_rows = std::max(4, std::min(h->ivf_sample_rows, (kIvfTauMaxKeys / nprobe) & ~3));
  int64_t worst_sub = 0;
  for (int i = 0; i < nprobe && i < (int)sz.size(); ++i)
    worst_sub += 2 * ceil_div(ceil_div(sz[i] > 0 ? sz[i] : 1, kTileRows), kUnitTilesSmallQ);
  pl.fused = h->kind == B2R_KIND_IVF_FLAT && h->ivf_fused && h->rescore && (int64_t)nprobe * pl.sample_rows <= kIvfTauMaxKeys &&
             worst_sub <= kIvfGatherMaxSub;
  const int64_t slot_bytes = pl.fused ? 8 : 4;
  const int64_t budget = ((int64_t)6 << 30) / 4 * slot_bytes;   // runs of one query chunk (fewer chunks = fewer passes over the lists)
  int64_t qc = budget / (pl.smax * slot_bytes);
  if (qc < 1) qc = 1;
  if (qc > 8192) qc = 8192;
  pl.chunk = (int)(q < qc ? q : qc);
  pl.qpad = (int)align_up((size_t)pl.chunk, 128);
  const int64_t pairs = (int64_t)pl.chunk * nprobe;
  pl.pairs_pad = pairs + kQBlock;
  int64_t mu = (pairs / kQBlock + std::min<int64_t>(h->nlist, pairs) + 1) * ceil_div(max_tiles, kUnitTilesSmallQ);
  if (mu > (1 << 24)) mu = 1 << 24;
  pl.max_units = (int)mu;
  if (mu * 1024 > ((int64_t)1 << 30)) pl.fused = false;    // segment counters: 1 KB per unit
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
  pl.off_q16 = take((size_t)pl.qpad * h->d * 2);
  pl.off_q32 = take((size_t)pl.qpad * h->d * 4);
  pl.off_qnorm = take((size_t)pl.qpad * 4);
  pl.off_coarse = take((size_t)pairs * 8);
  pl.off_cdist = take((size_t)pairs * 4);
  pl.off_cstatus = take((size_t)pl.chunk * 4);
  pl.off_pair_out = take((size_t)pairs * 8);
  pl.off_row_len = take((size_t)pl.chunk * 4);
  pl.off_listcnt = take((size_t)h->nlist * 4);
  pl.off_pairoff = take((size_t)(h->nlist + 1) * 8);
  pl.off_cursor = take((size_t)h->nlist * 8);
  pl.off_pair_sorted = take((size_t)pl.pairs_pad * 4);
  pl.off_gq16 = take((size_t)pl.pairs_pad * h->d * 2);
  pl.off_units = take((size_t)pl.max_units * sizeof(IvfUnit));
  pl.off_nunits = take(256);
  pl.off_tau = take((size_t)pl.chunk * 4);
  pl.off_count = take((size_t)pl.chunk * 4);
  pl.off_cand = take((size_t)pl.chunk * pl.cap * 8);
  pl.off_score = take((size_t)pl.chunk * pl.smax * (pl.fused ? 8 : 4));   // dumped scores, or the candidate pool (its head doubles as the sample's score buffer)
  pl.off_pair_pos = take((size_t)pairs * 4);
  pl.off_list_unit0 = take((size_t)h->nlist * 4);
  pl.off_seg_count = take(pl.fused ? (size_t)pl.max_units * 1024 : 0);
  {
    // the search-time coarse quantiser always takes the flat index's DENSE path (dump + exact k-th): nlist <= 65536
    // always fits, and the probed list set must never depend on a sampled threshold (see ivf_search)
    ForcePath dense(h->quantizer, 1);
    pl.qws_bytes = flat_search_workspace(h->quantizer, pl.chunk, nprobe);
  }
  pl.off_qws = take(pl.qws_bytes);
  pl.off_qtab = take(h->kind == B2R_KIND_IVF_PQ ? (size_t)pl.chunk * h->pq_m * 256 * 4 : 0);
  pl.total = off;
  return pl;
}
}  // namespace

size_t ivf_search_workspace(const b2r_index* h, int q, int k, int nprobe) {
  if (!h || !h->trained) return 0;
  return make_ivf_plan(h, q, k, nprobe).total;
}

int ivf_search(b2r_index* h, int q, const float* queries, int normalize, int k, int nprobe, float* D,
               int64_t* I, int32_t* status, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  const IvfPlan pl = make_ivf_plan(h, q, k, nprobe);
  if (!workspace || ws_bytes < pl.total)
    return fail(B2R_ENOMEM, "index_search: workspace too small (need " + std::to_string(pl.total) + " bytes)");
  const bool is_pq = h->kind == B2R_KIND_IVF_PQ;
  const int d = h->d, np = pl.nprobe;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  static bool configured[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2R_CUDA(cudaFuncSetAttribute(ivf_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvfSmem));
    B2R_CUDA(cudaFuncSetAttribute(ivf_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvfSmem));
    B2R_CUDA(cudaFuncSetAttribute(ivf_sample_tau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    B2R_CUDA(cudaFuncSetAttribute(ivf_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured[dev & 63] = true;
  }
  for (int q0 = 0; q0 < q; q0 += pl.chunk) {
    const int qc = (q - q0) < pl.chunk ? (q - q0) : pl.chunk;
    const int qpad = (int)align_up((size_t)qc, 128);
    const int npairs = qc * np;
    __nv_bfloat16* q16 = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_q16);
    float* q32 = reinterpret_cast<float*>(ws + pl.off_q32);
    float* qnorm = reinterpret_cast<float*>(ws + pl.off_qnorm);
    int64_t* coarse = reinterpret_cast<int64_t*>(ws + pl.off_coarse);
    float* cdist = reinterpret_cast<float*>(ws + pl.off_cdist);
    int64_t* pair_out = reinterpret_cast<int64_t*>(ws + pl.off_pair_out);
    int* row_len = reinterpret_cast<int*>(ws + pl.off_row_len);
    int* listcnt = reinterpret_cast<int*>(ws + pl.off_listcnt);
    int64_t* pairoff = reinterpret_cast<int64_t*>(ws + pl.off_pairoff);
    int64_t* cursor = reinterpret_cast<int64_t*>(ws + pl.off_cursor);
    int* pair_sorted = reinterpret_cast<int*>(ws + pl.off_pair_sorted);
    __nv_bfloat16* gq16 = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_gq16);
    IvfUnit* units = reinterpret_cast<IvfUnit*>(ws + pl.off_units);
    int* nunits = reinterpret_cast<int*>(ws + pl.off_nunits);
    float* tau = reinterpret_cast<float*>(ws + pl.off_tau);
    int* count = reinterpret_cast<int*>(ws + pl.off_count);
    uint2* cand = reinterpret_cast<uint2*>(ws + pl.off_cand);
    float* scorebuf = reinterpret_cast<float*>(ws + pl.off_score);
    int rc;
    if ((rc = launch_prep_queries(queries + (size_t)q0 * d, qc, qpad, d, normalize, q32, q16, h->scan_fp16 == 1, qnorm,
                                  stream)))
      return rc;
    // coarse quantiser: exact top-nprobe centroids, best first (faiss quantizer->search).  Dense path: the
    // candidate threshold is the exact k-th scan score minus the rescore margin, so the probed list set cannot
    // depend on a sampled threshold; what the flat search still flags (candidate overflow among near-ties) is
    // OR-ed into the caller's status words below instead of being dropped.
    int32_t* cstatus = reinterpret_cast<int32_t*>(ws + pl.off_cstatus);
    {
      ForcePath dense(h->quantizer, 1);
      rc = flat_search(h->quantizer, qc, q32, 0, np, cdist, coarse, status ? cstatus : nullptr, nullptr, nullptr,
                       ws + pl.off_qws, pl.qws_bytes, stream);
    }
    if (rc) return rc;
    if (h->ntotal == 0) {
      if ((rc = launch_fill_f32(D + (size_t)q0 * k, (int64_t)qc * k, -3.4028234663852886e38f, stream))) return rc;
      B2R_CUDA(cudaMemsetAsync(I + (size_t)q0 * k, 0xFF, (size_t)qc * k * 8, stream));
      if (status) B2R_CUDA(cudaMemsetAsync(status + q0, 0, (size_t)qc * 4, stream));
      continue;
    }
    B2R_CUDA(cudaMemsetAsync(listcnt, 0, (size_t)h->nlist * 4, stream));
    B2R_CUDA(cudaMemsetAsync(nunits, 0, 4, stream));
    pair_runs_kernel<<<(unsigned)ceil_div(qc, 4), 128, 0, stream>>>(coarse, qc, np, h->list_off, pl.smax, pair_out,
                                                                      row_len, listcnt);
    B2R_CHECK_LAUNCH("pair_runs_kernel");
    // Below ~1k queries a list is scanned for a handful of pairs: the dump is small, and the fused path's two
    // extra launches (sample scan, segment gather) cost what the score round trip saves (measured at 10M x 256,
    // nlist 4096, nprobe 32: Q=64 0.53 vs 0.52 ms, Q=1024 a wash, Q=4096 see DESIGN.md 3.6).
    // (a caller that passes no status array cannot react to a flagged query: it gets the dump path, whose fallback is in-kernel)
    const bool fused = pl.fused && !is_pq && qc >= 1024 && status != nullptr;
    if (is_pq) {
      // ADC scan: one CTA per (query, probed list) pair, LUT in shared memory (ivfpq.cu)
      if ((rc = pq_scan(h, qc, npairs, q32, coarse, np, pair_out, reinterpret_cast<float*>(ws + pl.off_qtab), scorebuf,
                        stream)))
        return rc;
    } else {
    scan_lists_kernel<<<1, 1024, 0, stream>>>(listcnt, nullptr, h->nlist, pairoff, cursor);
    B2R_CHECK_LAUNCH("scan_lists_kernel(pairs)");
    int* pair_pos = reinterpret_cast<int*>(ws + pl.off_pair_pos);          // inverse maps for the fused path's gather
    int* list_unit0 = reinterpret_cast<int*>(ws + pl.off_list_unit0);
    const int unit_tiles = qc >= 512 ? kUnitTilesLargeQ : kUnitTilesSmallQ;
    scatter_pairs_kernel<<<(unsigned)ceil_div(npairs, 8), 256, 0, stream>>>(
        coarse, npairs, np, reinterpret_cast<unsigned long long*>(cursor), pair_sorted, q16, d, gq16, pair_pos);
    B2R_CHECK_LAUNCH("scatter_pairs_kernel");
    build_units_kernel<<<(unsigned)ceil_div(h->nlist, 128), 128, 0, stream>>>(
        h->list_off, pairoff, h->nlist, units, nunits, pl.max_units, unit_tiles, list_unit0);
    B2R_CHECK_LAUNCH("build_units_kernel");
    CUtensorMap tmQ;
    if ((rc = make_tmap_bf16_rows(&tmQ, gq16, pl.pairs_pad, d))) return rc;
    IvfScanParams sp;
    sp.units = units;
    sp.num_units = nunits;
    sp.max_units = pl.max_units;
    sp.d = d;
    sp.idesc = scan_idesc(h->scan_fp16 == 1);
    sp.pair_sorted = pair_sorted;
    sp.pair_out = pair_out;
    sp.scorebuf = scorebuf;
    sp.debug = h->ivf_debug;
    sp.sample = 0;
    sp.tau = tau;
    sp.nprobe = np;
    sp.pool = reinterpret_cast<uint2*>(scorebuf);
    sp.seg_count = reinterpret_cast<int*>(ws + pl.off_seg_count);
    if (fused) {
      // 1. sample: the first tile of every probed list, dumped into the head of each pair's run
      sp.sample = 1;
      ivf_scan_kernel<false><<<h->num_sms, kIvfThreads, kIvfSmem, stream>>>(tmQ, h->tmX, sp);
      B2R_CHECK_LAUNCH("ivf_scan_kernel(sample)");
      // 2. per-query threshold from the length-weighted sample
      ivf_sample_tau_kernel<<<qc, kIvfTauThreads, (size_t)np * pl.sample_rows * 8 + (size_t)np * 12 + 16, stream>>>(
          scorebuf, pl.smax, coarse, np, h->list_off, pl.sample_rows, pl.c_target, pl.cap - pl.cap / 8, tau);
      B2R_CHECK_LAUNCH("ivf_sample_tau_kernel");
      // 3. the list scan with the threshold in its epilogue: only candidates leave the SM
      sp.sample = 0;
      ivf_scan_kernel<true><<<h->num_sms, kIvfThreads, kIvfSmem, stream>>>(tmQ, h->tmX, sp);
      B2R_CHECK_LAUNCH("ivf_scan_kernel(filter)");
      // 4. segments -> one candidate list per query
      ivf_gather_kernel<<<qc, kIvfGatherThreads, (size_t)(2 * np + 1) * 4 + (size_t)(2 * kIvfGatherMaxSub + 1) * 4 + 16, stream>>>(
          coarse, np, h->list_off, pairoff, pair_pos, list_unit0, unit_tiles, reinterpret_cast<const uint2*>(scorebuf),
          pl.smax, sp.seg_count, cand, count, pl.cap);
      B2R_CHECK_LAUNCH("ivf_gather_kernel");
    } else {
      ivf_scan_kernel<false><<<h->num_sms, kIvfThreads, kIvfSmem, stream>>>(tmQ, h->tmX, sp);
      B2R_CHECK_LAUNCH("ivf_scan_kernel");
    }
    }
    if (!fused) {
    ivf_threshold_kernel<<<qc, kIvfSelThreads, (size_t)np * 24, stream>>>(
        scorebuf, pl.smax, row_len, k, coarse, np, h->list_off, tau, count, cand, pl.cap,
        (h->rescore && !is_pq) ? qnorm : nullptr, h->maxnorm, (float)(h->scan_fp16 == 1 ? h->eps_fp16 : h->eps),
        is_pq ? 4 : 3,    // IVF-Flat re-scores the window, a 2^-15-relative lower bound suffices; PQ needs the exact k-th
        pl.sample_stride, pl.c_target);
    B2R_CHECK_LAUNCH("ivf_threshold_kernel");
    }
    SelectParams sel;
    memset(&sel, 0, sizeof(sel));
    sel.Q = qc;
    sel.k = k;
    sel.d = d;
    sel.nseg = 1;
    sel.cap_seg = pl.cap;
    sel.N = h->ntotal;
    sel.cand_count = count;
    sel.cand = cand;
    sel.tau = tau;
    sel.q32 = q32;
    sel.qnorm = qnorm;
    sel.x32 = h->x32;
    sel.maxnorm = h->maxnorm;
    sel.eps = (float)(h->scan_fp16 == 1 ? h->eps_fp16 : h->eps);
    sel.rescore = is_pq ? 0 : h->rescore;   // ADC distances are the result (faiss does not re-rank)
    sel.negate_out = is_pq ? 1 : 0;
    sel.perm = h->perm;
    sel.ids = (h->ids && h->n_ids >= h->ntotal) ? h->ids : nullptr;
    sel.label_base = h->label_base;
    sel.D = D + (size_t)q0 * k;
    sel.I = I + (size_t)q0 * k;
    sel.status = status ? status + q0 : nullptr;
    sel.tau_retry = nullptr;
    sel.scanned = row_len;
    if ((rc = launch_select_rescore(sel, stream))) return rc;
    if (status) {
      or_status_kernel<<<(unsigned)ceil_div(qc, 256), 256, 0, stream>>>(status + q0, cstatus, qc);
      B2R_CHECK_LAUNCH("or_status_kernel");
    }
  }
  return B2R_OK;
}

}  // namespace b2r

using namespace b2r;

extern "C" {

int b2r_index_export_centroids(const b2r_index* h, float* out) {
  if (!h || !out) return fail(B2R_EINVAL, "export_centroids: NULL argument");
  if (h->kind == B2R_KIND_FLAT || !h->quantizer) return fail(B2R_EUNSUPPORTED, "no coarse quantiser on this index kind");
  if (!h->trained) return fail(B2R_ESTATE, "export_centroids: index is not trained");
  DeviceGuard g(h->device);
  B2R_CUDA(cudaMemcpy(out, h->quantizer->x32, (size_t)h->nlist * h->d * 4, cudaMemcpyDeviceToHost));
  return B2R_OK;
}

int b2r_index_import_centroids(b2r_index* h, const float* in) {
  if (!h || !in) return fail(B2R_EINVAL, "import_centroids: NULL argument");
  if (h->kind == B2R_KIND_FLAT || !h->quantizer) return fail(B2R_EUNSUPPORTED, "no coarse quantiser on this index kind");
  if (h->ntotal > 0) return fail(B2R_ESTATE, "import_centroids: index already holds vectors");
  DeviceGuard g(h->device);
  DevBuf tmp;
  int rc = tmp.alloc((size_t)h->nlist * h->d * 4);
  if (rc) return rc;
  B2R_CUDA(cudaMemcpy(tmp.p, in, (size_t)h->nlist * h->d * 4, cudaMemcpyHostToDevice));
  rc = set_centroids(h, tmp.as<float>(), 0);
  if (rc) return rc;
  B2R_CUDA(cudaStreamSynchronize(0));
  return B2R_OK;
}

/* Restore pre-encoded vectors (load path): codes uint8 [n, pq_m] and their list ids int64 [n], device. */
int b2r_index_add_codes(b2r_index* h, int64_t n, const uint8_t* codes, const int64_t* list_ids, void* stream) {
  if (!h || (n > 0 && (!codes || !list_ids))) return fail(B2R_EINVAL, "add_codes: NULL argument");
  if (h->kind != B2R_KIND_IVF_PQ) return fail(B2R_EUNSUPPORTED, "add_codes: IVF-PQ only");
  if (!h->trained) return fail(B2R_ESTATE, "add_codes: index is not trained");
  if (n <= 0) return B2R_OK;
  DeviceGuard g(h->device);
  return ivf_add_codes(h, n, codes, list_ids, (cudaStream_t)stream);
}

int b2r_index_list_sizes(const b2r_index* h, int64_t* sizes) {
  if (!h || !sizes) return fail(B2R_EINVAL, "list_sizes: NULL argument");
  if (h->kind == B2R_KIND_FLAT) return fail(B2R_EUNSUPPORTED, "no inverted lists on this index kind");
  for (int l = 0; l < h->nlist; ++l) sizes[l] = l < (int)h->list_sizes_host.size() ? h->list_sizes_host[l] : 0;
  return B2R_OK;
}


int b2r_index_get_labels(const b2r_index* h, int64_t row0, int64_t n, int64_t* out, void* stream) {
  if (!h || (n > 0 && !out)) return fail(B2R_EINVAL, "get_labels: NULL argument");
  if (row0 < 0 || n < 0 || row0 + n > h->ntotal) return fail(B2R_EINVAL, "get_labels: range out of bounds");
  if (n == 0) return B2R_OK;
  DeviceGuard g(h->device);
  std::vector<int64_t> host(n);
  if (h->kind == B2R_KIND_FLAT) {
    for (int64_t i = 0; i < n; ++i) host[i] = row0 + i;
  } else {
    std::vector<uint32_t> p32(n);
    B2R_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    B2R_CUDA(cudaMemcpy(p32.data(), h->perm + row0, (size_t)n * 4, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) host[i] = p32[i];
  }
  B2R_CUDA(cudaMemcpy(out, host.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  return B2R_OK;
}

}  // extern "C"
