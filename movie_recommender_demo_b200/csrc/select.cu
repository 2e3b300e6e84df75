// select.cu — exact top-k selection around the tensor-core scan.
//
// Replaces the heap/reservoir top-k inside faiss `IndexFlatIP::search`
// (faiss_retrieval.py:155) and the python id remap loop (faiss_retrieval.py:159-160).
//
//   kth_value        per-row m-th largest (MSB-first radix select on order-preserving keys)
//                    -> candidate threshold tau; optionally compacts (val >= tau) into the
//                    candidate buffers (dense small-corpus path)
//   select_rescore   per-query: sort candidates by bf16 score, take the provable rescore
//                    window {s >= s_k - 2E}, exact fp32 re-score from the master rows, final
//                    sort by (-score, label), id-map gather, write D / I
//   topk_merge       P sorted per-shard lists -> global top-k (multi-GPU)
//
// Integer/compare work; bounded by shared-memory sort latency and the random 1 KB master-row
// gather, not by tensor throughput.
#include <float.h>

#include "internal.h"
#include "ptx.cuh"

namespace b2r {
namespace {

constexpr int kSelThreads = 512;
constexpr int kRescoreMaxBig = 2048;  // rescore window slots (k > 96)
constexpr int kKeyCapBig = 4096;      // candidates per query the select kernel can hold (k > 96)
constexpr int kRescoreMaxSmall = 256; // small-k variant (coarse quantiser, k-means assignment): 256 threads,
constexpr int kKeyCapSmall = 1024;    // ~12 KB of shared memory -> many CTAs per SM

__device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
  return ((uint64_t)f2ord(score) << 32) | (uint64_t)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float key_score(uint64_t k) { return ord2f((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_idx(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }

// in-place bitonic sort, descending, n a power of two, all threads of the block participate
__device__ void bitonic_desc(uint64_t* s, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if (up ? (a < b) : (a > b)) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// Same sort with one key per thread kept in a register, n <= blockDim.x (a power of two): compare-
// exchange distances below 32 go through warp shuffles, only distances >= 32 touch shared memory.
// Only the first max(n, 32) threads take part (named barrier 1); the caller's next __syncthreads()
// joins everybody again.
__device__ void bitonic_desc_reg(uint64_t* s, int n) {
  const int i = threadIdx.x;
  const int nthr = n < 32 ? 32 : n;
  if (i < nthr) {
    uint64_t v = i < n ? s[i] : 0ull;
    for (int k = 2; k <= n; k <<= 1) {
      const bool up = (i & k) == 0;
      for (int j = k >> 1; j > 0; j >>= 1) {
        uint64_t o;
        if (j >= 32) {
          s[i] = v;
          asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
          o = s[i ^ j];
          asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        } else {
          o = __shfl_xor_sync(0xffffffffu, v, j);
        }
        const bool lower = (i & j) == 0;           // this thread holds the lower index of the pair
        const bool take_max = (lower == up);       // descending run: lower index keeps the larger key
        const uint64_t mx = v > o ? v : o, mn = v > o ? o : v;
        v = take_max ? mx : mn;
      }
    }
    if (i < n) s[i] = v;
  }
  __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Warp-parallel search of the histogram bin (scanning from bin 255 down) that holds rank *s_rem:
// each lane sums 8 bins, a shuffle scan finds the owning lane, which walks its 8 bins.
// Called by the first warp only; updates *s_rem and *s_prefix.
__device__ __forceinline__ void radix_pick_bin(const int* hist, int* s_rem, uint32_t* s_prefix,
                                               uint32_t prefix, int shift) {
  const int lane = threadIdx.x & 31;
  int part = 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) part += hist[255 - (lane * 8 + b)];
  int incl = part;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int rem = *s_rem;
  __syncwarp();
  if ((incl >= rem) && (incl - part < rem)) {  // exactly one lane
    int cum = incl - part, b = 0;
    for (; b < 7; ++b) {
      const int hb = hist[255 - (lane * 8 + b)];
      if (cum + hb >= rem) break;
      cum += hb;
    }
    *s_rem = rem - cum;
    *s_prefix = prefix | ((uint32_t)(255 - (lane * 8 + b)) << shift);
  }
}

// Block-wide: high word (order-preserving score bits) of the m-th largest of keys[0, c), 1 <= m <= c.
// MSB-first radix select, 4 passes of 8 bits over a 256-bin shared histogram.  All threads call it.
__device__ uint32_t radix_kth_hi(const uint64_t* keys, int c, int m, int* hist, int* s_rem, uint32_t* s_prefix) {
  const int tid = threadIdx.x;
  uint32_t prefix = 0, mask = 0;
  if (tid == 0) *s_rem = m;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < c; i += blockDim.x) {
      const uint32_t key = (uint32_t)(keys[i] >> 32);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
    }
    __syncthreads();
    if (tid < 32) radix_pick_bin(hist, s_rem, s_prefix, prefix, shift);
    __syncthreads();
    prefix = *s_prefix;
    mask |= 255u << shift;
  }
  return prefix;
}

// ------------------------------------------------------------- kth_value ---
__global__ void __launch_bounds__(kSelThreads)
kth_value_kernel(const float* __restrict__ vals, int64_t T, int64_t ld, int m, float* __restrict__ tau,
                 int* __restrict__ cand_count, uint2* __restrict__ cand, int cap,
                 const float* __restrict__ qnorm, const float* __restrict__ maxnorm, float eps) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_rem;
  __shared__ int s_count;
  const int r = blockIdx.x;
  const float* row = vals + (size_t)r * ld;
  float t;
  if ((int64_t)m >= T) {
    t = -INFINITY;  // everything qualifies
  } else {
    uint32_t prefix = 0, mask = 0;
    if (threadIdx.x == 0) s_rem = m;
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      {
        // run-length aggregation: the scores of one query share their leading key bytes, so a thread's
        // consecutive elements mostly hit the same bin -> one shared-memory atomic per run, not per element
        int last = -1, run = 0;
        for (int64_t i = threadIdx.x; i < T; i += blockDim.x) {
          const uint32_t key = f2ord(row[i]);
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
        }
        if (run) atomicAdd(&hist[last], run);
      }
      __syncthreads();
      if (threadIdx.x < 32) radix_pick_bin(hist, &s_rem, &s_prefix, prefix, shift);
      __syncthreads();
      prefix = s_prefix;
      mask |= 255u << shift;
    }
    t = ord2f(prefix);
    if (qnorm) t -= rescore_margin(eps, qnorm[r], *maxnorm);
  }
  if (threadIdx.x == 0) tau[r] = t;
  if (cand_count) {
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < T; i += blockDim.x) {
      const float v = row[i];
      if (v >= t) {
        const int slot = atomicAdd(&s_count, 1);
        if (slot < cap) cand[(size_t)r * cap + slot] = make_uint2(__float_as_uint(v), (uint32_t)i);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) cand_count[r] = s_count;
  }
}

// kth_value for short rows (T <= 4096, the sampled group maxima): the row is read ONCE into
// registers (coalesced), the four radix passes run on registers + a shared histogram.
constexpr int kKthSmallThreads = 256;
constexpr int kKthSmallPer = 16;

__global__ void __launch_bounds__(kKthSmallThreads)
kth_value_small_kernel(const float* __restrict__ vals, int T, int64_t ld, int m, float* __restrict__ tau) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_rem;
  const int r = blockIdx.x;
  const float* row = vals + (size_t)r * ld;
  uint32_t key[kKthSmallPer];
#pragma unroll
  for (int u = 0; u < kKthSmallPer; ++u) {
    const int i = u * kKthSmallThreads + threadIdx.x;
    key[u] = i < T ? f2ord(row[i]) : 0u;   // padding keys sort below every real value
  }
  if (m >= T) {
    if (threadIdx.x == 0) tau[r] = -INFINITY;
    return;
  }
  uint32_t prefix = 0, mask = 0;
  if (threadIdx.x == 0) s_rem = m;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kKthSmallPer; ++u) {
      if (u * kKthSmallThreads < T && (key[u] & mask) == prefix && key[u] != 0u)
        atomicAdd(&hist[(key[u] >> shift) & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) radix_pick_bin(hist, &s_rem, &s_prefix, prefix, shift);
    __syncthreads();
    prefix = s_prefix;
    mask |= 255u << shift;
  }
  if (threadIdx.x == 0) tau[r] = ord2f(prefix);
}

// --------------------------------------------------------- select_rescore ---
// One CTA (1024 threads) per query:
//   1. candidates -> smem keys (ord(score) << 32 | ~row)
//   2. k-th largest bf16 score by MSB-first radix select over the smem keys (no full sort)
//   3. rescore window = {score >= kth - 2E}, compacted (order irrelevant)
//   4. exact fp32 dot per window entry: one warp per row, 4 rows in flight per warp
//   5. bitonic sort of the window by (-score, row), gather ids, write top-k
constexpr int kSel2Threads = 1024;

// MINB = CTAs per SM the register budget is sized for: 2 (32 registers) when there are more queries than SMs,
// 1 (64 registers, all eight loads of a 256-d row in flight per lane) for small batches, where a second
// resident CTA has nothing to do and the kernel is a pure latency chain.  Same arithmetic order in both.
template <int MINB>
__global__ void __launch_bounds__(kSel2Threads, MINB)
select_rescore_kernel(const SelectParams p) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int kKeyCap = p.key_cap, kRescoreMax = p.rescore_max;             // per-launch capacities
  uint64_t* keys = reinterpret_cast<uint64_t*>(sm);                       // [key_cap]
  uint64_t* rkeys = keys + kKeyCap;                                        // [rescore_max]
  float* qv = reinterpret_cast<float*>(rkeys + kRescoreMax);               // [d]
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_rem;
  __shared__ int s_R;
  __shared__ int s_c, s_over, s_status;
  __shared__ float s_lim;
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) { s_c = 0; s_over = 0; }
  for (int i = tid; i < p.d; i += blockDim.x) qv[i] = p.q32[(size_t)q * p.d + i];
  __syncthreads();
  {
    // gather this query's candidate segments into the key array.  Segment counts are read with
    // one coalesced load, offsets come from a block-wide exclusive scan (shared memory), then
    // every (segment, lane) pair copies its entries, so all the segment reads are in flight at
    // once instead of one latency per segment.
    int* soff = reinterpret_cast<int*>(qv + p.d);  // [nseg + 1] scratch behind the query vector
    for (int sgi = tid; sgi < p.nseg; sgi += blockDim.x) {
      const int produced = p.cand_count[(size_t)q * p.nseg + sgi];
      if (produced > p.cap_seg) s_over = 1;
      soff[sgi + 1] = produced < p.cap_seg ? produced : p.cap_seg;
    }
    if (tid == 0) soff[0] = 0;
    __syncthreads();
    if (tid < 32) {  // warp scan over nseg (<= 4096) counts, 32 at a time
      int carry = 0;
      for (int b = 0; b < p.nseg; b += 32) {
        const int idx = b + tid;
        int v = idx < p.nseg ? soff[idx + 1] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, v, o);
          if (tid >= o) v += t;
        }
        if (idx < p.nseg) soff[idx + 1] = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
      }
      if (tid == 0) s_c = carry;
    }
    __syncthreads();
    // without a rescore (IVF-PQ) the scan score is final: key on the LABEL right away, so that the
    // selection below is canonical (score desc, label asc) even among exact ties
    const uint32_t* to_label = p.rescore ? nullptr : p.perm;
    const int nwarps = blockDim.x >> 5;
    if (p.nseg > nwarps) {
      // many short segments (flat scan at small batch: ~300 segments of ~4 entries): one thread per
      // CANDIDATE, its segment found by binary search over the offsets -> every load of the gather is in
      // flight at once instead of one dependent round trip per segment
      const int total = s_c < kKeyCap ? s_c : kKeyCap;
      for (int idx = tid; idx < total; idx += blockDim.x) {
        int lo = 0, hi = p.nseg - 1;            // largest sgi with soff[sgi] <= idx
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (soff[mid] <= idx) lo = mid; else hi = mid - 1;
        }
        const uint2 e = p.cand[((size_t)q * p.nseg + lo) * p.cap_seg + (idx - soff[lo])];
        keys[idx] = make_key(__uint_as_float(e.x), to_label ? to_label[e.y] : e.y);
      }
    } else {
      // a group of G threads copies one segment: a warp per segment, or the whole block when there is one
      // (IVF: a single candidate list per query)
      const int G = p.nseg >= nwarps ? 32 : ((int)blockDim.x / p.nseg) & ~31;
      const int groups = (int)blockDim.x / G, gid = tid / G, lig = tid % G;
      if (gid < groups) {
        for (int sgi = gid; sgi < p.nseg; sgi += groups) {
          const int pos0 = soff[sgi], n = soff[sgi + 1] - pos0;
          const uint2* seg = p.cand + ((size_t)q * p.nseg + sgi) * p.cap_seg;
          for (int i = lig; i < n; i += G) {
            if (pos0 + i < kKeyCap) {
              const uint2 e = seg[i];
              keys[pos0 + i] = make_key(__uint_as_float(e.x), to_label ? to_label[e.y] : e.y);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  const int c_total = s_over ? kKeyCap + 1 : s_c;
  const int c = s_c < kKeyCap ? s_c : kKeyCap;

  const int64_t n_avail = p.scanned ? (int64_t)p.scanned[q] : p.N;
  const int kk = (int64_t)p.k < n_avail ? p.k : (int)n_avail;  // results that exist
  const float tau_q = p.tau[q];
  const float margin = rescore_margin(p.eps, p.qnorm[q], *p.maxnorm);
  int status = 0;
  if (c_total > kKeyCap) status |= B2R_ST_CAND_OVERFLOW;
  if (c < kk && tau_q > -INFINITY) status |= B2R_ST_TOO_FEW;

  // ---- k-th largest candidate score (rank m = min(kk, c)), radix select on the high word
  const int m = c < kk ? c : kk;
  float kth = -INFINITY;
  if (m >= 1) kth = ord2f(radix_kth_hi(keys, c, m, hist, &s_rem, &s_prefix));
  const float lim = p.rescore ? kth - margin : kth;
  if (p.rescore && c >= kk && kk > 0 && tau_q > lim) status |= B2R_ST_NEED_LOWER_TAU;
  if (tid == 0) {   // parked in shared memory: the register budget is 32 and these are needed only at the very end
    s_status = status;
    s_lim = lim;
  }

  // ---- compact the rescore window into rkeys (unordered)
  if (tid == 0) s_R = 0;
  __syncthreads();
  for (int i = tid; i < c; i += blockDim.x) {
    const uint64_t key = keys[i];
    if (key_score(key) >= lim) {
      const int pos = atomicAdd(&s_R, 1);
      if (pos < kRescoreMax) rkeys[pos] = key;
    }
  }
  __syncthreads();
  int R = s_R;
  if (R > kRescoreMax) {
    // window larger than the buffer: keep the best kRescoreMax by scan score.  With a rescore that is
    // not provably exact; without one (keys already final) the best kRescoreMax >= k are the answer.
    if (p.rescore) status |= B2R_ST_RESCORE_OVERFLOW;
    if (tid == 0) s_status = status;
    __syncthreads();
    const int P = next_pow2(c > 1 ? c : 1);
    for (int i = c + tid; i < P; i += blockDim.x) keys[i] = 0;
    __syncthreads();
    bitonic_desc(keys, P);
    for (int i = tid; i < kRescoreMax; i += blockDim.x) rkeys[i] = keys[i];
    R = kRescoreMax;
    __syncthreads();
  }

  if (p.rescore) {
    // exact fp32 dot against the master rows: 8 lanes per row (each LDG.128 of a row group reads 128
    // contiguous bytes), 4 rows per warp pass, 4 loads in flight per lane; one accumulator per lane
    // and a 3-step shuffle reduction.  Sized for the 32-register budget of 2 x 1024 threads per SM.
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int sub = lane >> 3, l8 = lane & 7;
    const int nv = p.d >> 2;
    const uint32_t row_bytes = (uint32_t)p.d * 4u;
    const char* xbase = reinterpret_cast<const char*>(p.x32);
    const float4* qv4 = reinterpret_cast<const float4*>(qv);
    const bool full = (nv & 31) == 0;   // d = 128 or 256: every pass is complete, no bounds predicates
    for (int i0 = warp * 4; i0 < R; i0 += nwarps * 4) {
      const int i = i0 + sub;
      const uint32_t ix = key_idx(rkeys[i < R ? i : R - 1]);   // past the window: re-read its last row
      const float4* xp = reinterpret_cast<const float4*>(xbase + (uint64_t)ix * row_bytes) + l8;
      const float4* qp = qv4 + l8;
      float acc = 0.f;
      if (MINB == 1 && nv == 64) {
        float4 xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = __ldg(xp + u * 8);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 qq = qp[u * 8];
          acc = fmaf(xv[u].x, qq.x, acc);
          acc = fmaf(xv[u].y, qq.y, acc);
          acc = fmaf(xv[u].z, qq.z, acc);
          acc = fmaf(xv[u].w, qq.w, acc);
        }
      } else if (full) {
#pragma unroll 1
        for (int t = 0; t < nv; t += 32) {
          float4 xv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) xv[u] = __ldg(xp + t + u * 8);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 qq = qp[t + u * 8];
            acc = fmaf(xv[u].x, qq.x, acc);
            acc = fmaf(xv[u].y, qq.y, acc);
            acc = fmaf(xv[u].z, qq.z, acc);
            acc = fmaf(xv[u].w, qq.w, acc);
          }
        }
      } else {
#pragma unroll 1
        for (int t = l8; t < nv; t += 8) {
          const float4 xv = __ldg(xp + t - l8), qq = qv4[t];
          acc = fmaf(xv.x, qq.x, acc);
          acc = fmaf(xv.y, qq.y, acc);
          acc = fmaf(xv.z, qq.z, acc);
          acc = fmaf(xv.w, qq.w, acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (l8 == 0 && i < R) rkeys[i] = make_key(acc, p.perm ? p.perm[ix] : ix);   // final order is by label
    }
  }
  __syncthreads();

  // ---- final order.  When the window is larger than it has to be for the sort size (typical:
  // k = 500, window ~560 -> a 1024-key sort), first cut it down to the exact top-kk (plus score ties)
  // with a radix select on the exact scores, so that the bitonic sort runs on next_pow2(kk) keys.
  uint64_t* sbuf = rkeys;
  int n_sort = R;
  const int P_need = next_pow2(kk > 1 ? kk : 1);
  // (sorting the whole ~560-key window as 1024 keys instead of cutting it down first was measured for the
  // small-batch variant: 25.3 -> 27.6 us at Q=64, so the cut-down stays for every batch size)
  if (kk >= 1 && R > kk && P_need < next_pow2(R)) {
    const uint32_t kth_hi = radix_kth_hi(rkeys, R, kk, hist, &s_rem, &s_prefix);
    if (tid == 0) s_R = 0;
    __syncthreads();
    for (int i = tid; i < R; i += blockDim.x) {
      const uint64_t key = rkeys[i];
      if ((uint32_t)(key >> 32) >= kth_hi) {
        const int pos = atomicAdd(&s_R, 1);
        if (pos < P_need) keys[pos] = key;
      }
    }
    __syncthreads();
    if (s_R <= P_need) {   // else: more ties at the k-th score than the smaller sort holds
      sbuf = keys;
      n_sort = s_R;
    }
  }
  const int R2 = next_pow2(n_sort > 1 ? n_sort : 1);
  for (int i = n_sort + tid; i < R2; i += blockDim.x) sbuf[i] = 0;
  __syncthreads();
  if (R2 <= (int)blockDim.x) bitonic_desc_reg(sbuf, R2);
  else bitonic_desc(sbuf, R2);

  const int avail = n_sort < kk ? n_sort : kk;
  for (int j = tid; j < p.k; j += blockDim.x) {
    float dv;
    int64_t iv;
    if (j < avail) {
      const uint64_t key = sbuf[j];
      const uint32_t idx = key_idx(key);
      dv = p.negate_out ? -key_score(key) : key_score(key);
      iv = p.ids ? p.ids[idx] : (p.label_base + (int64_t)idx);
    } else {
      dv = p.negate_out ? FLT_MAX : -FLT_MAX;
      iv = (p.ids && p.N > 0) ? p.ids[p.N - 1] : -1;  // reference: id_map[-1] wrap-around
    }
    p.D[(size_t)q * p.k + j] = dv;
    p.I[(size_t)q * p.k + j] = iv;
  }
  if (tid == 0) {
    const int status_f = s_status;
    if (p.status) p.status[q] = status_f;
    if (p.tau_retry) {
      float tr = s_lim;
      if (status_f & B2R_ST_TOO_FEW) tr = -INFINITY;
      else if (status_f & B2R_ST_CAND_OVERFLOW) tr = fmaxf(tr, p.tau[q]);
      p.tau_retry[q] = tr;
    }
  }
}

// -------------------------------------------------------------- topk_merge ---
__global__ void __launch_bounds__(kSelThreads)
topk_merge_kernel(int P, int k, const float* __restrict__ D_all, const int64_t* __restrict__ I_all,
                  size_t shard_stride, float* __restrict__ D_out, int64_t* __restrict__ I_out,
                  int largest) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(sm);
  const int q = blockIdx.x;
  const int n = P * k;
  const int n2 = next_pow2(n > 1 ? n : 1);
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    uint64_t key = 0;
    if (i < n) {
      const int s = i / k, j = i % k;
      const size_t off = (size_t)s * shard_stride + (size_t)q * k + j;
      if (I_all[off] >= 0) {
        const float v = D_all[off];
        key = ((uint64_t)f2ord(largest ? v : -v) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)i);
      }
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_desc(keys, n2);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t key = keys[j];
    if (key != 0) {
      const int i = (int)(0xFFFFFFFFu - (uint32_t)key);
      const size_t off = (size_t)(i / k) * shard_stride + (size_t)q * k + (i % k);
      D_out[(size_t)q * k + j] = D_all[off];
      I_out[(size_t)q * k + j] = I_all[off];
    } else {
      D_out[(size_t)q * k + j] = largest ? -FLT_MAX : FLT_MAX;
      I_out[(size_t)q * k + j] = -1;
    }
  }
}


// ------------------------------------------------- packed shard exchange (multi-GPU) ---
// One row per query: [k scores (fp32 bits)] [k LOCAL labels (int32, -1 = empty slot)] [status], i.e. 8 bytes per
// result instead of the 12 of (fp32 score, int64 global label): the shard's label base is added back after
// the exchange.  One buffer = ONE collective.
__global__ void __launch_bounds__(256)
topk_pack_kernel(int q, int q_rows, int k, const float* __restrict__ D, const int64_t* __restrict__ I,
                 const int32_t* __restrict__ status, int64_t base, int32_t* __restrict__ out, int largest) {
  const int W = 2 * k + 1;
  for (int row = blockIdx.x; row < q_rows; row += gridDim.x) {
    int32_t* o = out + (size_t)row * W;
    if (row < q) {
      for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const int64_t lab = I[(size_t)row * k + j];
        o[j] = __float_as_int(D[(size_t)row * k + j]);
        o[k + j] = lab < 0 ? -1 : (int32_t)(lab - base);
      }
      if (threadIdx.x == 0) o[2 * k] = status ? status[row] : 0;
    } else {   // padding rows of the last query slice: empty lists
      for (int j = threadIdx.x; j < k; j += blockDim.x) {
        o[j] = __float_as_int(largest ? -FLT_MAX : FLT_MAX);
        o[k + j] = -1;
      }
      if (threadIdx.x == 0) o[2 * k] = 0;
    }
  }
}

// Merge P best-first lists of k per query by RANKING instead of sorting: entry (shard s, position j) precedes
// entry (t, i) iff its score is better, or equal with (s, j) < (t, i) - the same total order as the bitonic
// merge above (and as an unsharded search: lower shard = lower labels).  Its output slot is
//   j + sum over t != s of #{entries of list t that precede it},
// each count a binary search in a sorted list: no barriers, every thread independent.
// packed: [P, q_stride rows, 2k+1]; this launch merges rows [0, q) of every shard's block.
__global__ void __launch_bounds__(kSelThreads)
topk_merge_packed_kernel(int P, int k, const int32_t* __restrict__ packed, size_t shard_stride,
                         const int64_t* __restrict__ bases, float* __restrict__ D_out,
                         int64_t* __restrict__ I_out, int32_t* __restrict__ status_out, int largest) {
  extern __shared__ __align__(16) uint8_t sm[];
  float* sc = reinterpret_cast<float*>(sm);   // [P][k] scores, sign-normalised so that larger is better
  const int q = blockIdx.x;
  const int W = 2 * k + 1;
  const int n = P * k;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int s = i / k, j = i - s * k;
    const float v = __int_as_float(packed[(size_t)s * shard_stride + (size_t)q * W + j]);
    sc[i] = largest ? v : -v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int s = i / k, j = i - s * k;
    const float v = sc[i];
    int rank = j;
    for (int t = 0; t < P && rank < k; ++t) {
      if (t == s) continue;
      const float* lt = sc + t * k;
      // lists are descending: count entries > v (t > s) or >= v (t < s)
      int lo = 0, hi = k;
      if (t < s) {
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (lt[mid] >= v) lo = mid + 1; else hi = mid; }
      } else {
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (lt[mid] > v) lo = mid + 1; else hi = mid; }
      }
      rank += lo;
    }
    if (rank < k) {
      const int32_t lab = packed[(size_t)s * shard_stride + (size_t)q * W + k + j];
      D_out[(size_t)q * k + rank] = largest ? v : -v;
      I_out[(size_t)q * k + rank] = lab < 0 ? -1 : bases[s] + (int64_t)lab;
    }
  }
  if (status_out && threadIdx.x == 0) {
    int st = 0;
    for (int s = 0; s < P; ++s) st |= packed[(size_t)s * shard_stride + (size_t)q * W + 2 * k];
    status_out[q] = st;
  }
}

}  // namespace

int launch_kth_value(const float* vals, int rows, int64_t T, int64_t ld, int m, float* tau,
                     int* cand_count, uint2* cand, int cap, const float* qnorm, const float* maxnorm,
                     float eps, cudaStream_t stream) {
  if (rows <= 0) return B2R_OK;
  if (m < 1) m = 1;
  if (!cand_count && !qnorm && T <= kKthSmallThreads * kKthSmallPer) {
    kth_value_small_kernel<<<rows, kKthSmallThreads, 0, stream>>>(vals, (int)T, ld, m, tau);
    B2R_CHECK_LAUNCH("kth_value_small_kernel");
    return B2R_OK;
  }
  kth_value_kernel<<<rows, kSelThreads, 0, stream>>>(vals, T, ld, m, tau, cand_count, cand, cap, qnorm, maxnorm, eps);
  B2R_CHECK_LAUNCH("kth_value_kernel");
  return B2R_OK;
}

int launch_select_rescore(const SelectParams& p_in, cudaStream_t stream) {
  SelectParams p = p_in;
  if (p.Q <= 0) return B2R_OK;
  // k <= 96 (coarse quantiser, k-means assignment): 256 threads, ~12 KB smem -> many queries per SM.
  // Larger k: one 1024-thread CTA per query, two resident per SM (register cap 32); 256-thread CTAs
  // were measured slower here (fewer rows in flight during the HBM-bound rescore gather).
  const bool small = p.k <= 96;
  p.key_cap = small ? kKeyCapSmall : kKeyCapBig;
  p.rescore_max = small ? kRescoreMaxSmall : kRescoreMaxBig;
  const int threads = small ? 256 : kSel2Threads;
  if (p.nseg < 1 || p.cap_seg < 1) return fail(B2R_EINVAL, "select: bad candidate segment layout");
  if (p.k > kRescoreMaxBig / 2) return fail(B2R_EUNSUPPORTED, "select: k must be <= 1024");
  if (p.d % 4 != 0) return fail(B2R_EINVAL, "select: d must be a multiple of 4");
  const size_t smem = (size_t)p.key_cap * 8 + (size_t)p.rescore_max * 8 + (size_t)p.d * 4 + (size_t)(p.nseg + 1) * 4;
  static bool configured[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2R_CUDA(cudaFuncSetAttribute(select_rescore_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kKeyCapBig * 8 + kRescoreMaxBig * 8 + 1024 * 4 + 8192 * 4));
    B2R_CUDA(cudaFuncSetAttribute(select_rescore_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kKeyCapBig * 8 + kRescoreMaxBig * 8 + 1024 * 4 + 8192 * 4));
    configured[dev & 63] = true;
  }
  int sms = 0;
  B2R_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (!small && p.Q <= sms) select_rescore_kernel<1><<<p.Q, threads, smem, stream>>>(p);
  else select_rescore_kernel<2><<<p.Q, threads, smem, stream>>>(p);
  B2R_CHECK_LAUNCH("select_rescore_kernel");
  return B2R_OK;
}

}  // namespace b2r

extern "C" int b2r_topk_merge(int P, int q, int k, const float* D_all, const int64_t* I_all,
                              float* D_out, int64_t* I_out, int largest, void* stream) {
  using namespace b2r;
  if (P < 1 || q < 0 || k < 1) return fail(B2R_EINVAL, "topk_merge: bad sizes");
  if ((int64_t)P * k > 8192) return fail(B2R_EUNSUPPORTED, "topk_merge: P*k must be <= 8192");
  if (q == 0) return B2R_OK;
  int n2 = 1;
  while (n2 < P * k) n2 <<= 1;
  const size_t smem = (size_t)n2 * 8;
  static bool configured[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2R_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  8192 * 8));
    configured[dev & 63] = true;
  }
  topk_merge_kernel<<<q, kSelThreads, smem, (cudaStream_t)stream>>>(
      P, k, D_all, I_all, (size_t)q * k, D_out, I_out, largest);
  B2R_CHECK_LAUNCH("topk_merge_kernel");
  return B2R_OK;
}

extern "C" int b2r_topk_pack(int q, int q_rows, int k, const float* D, const int64_t* I, const int32_t* status,
                             int64_t label_base, int32_t* out, int largest, void* stream) {
  using namespace b2r;
  if (q < 0 || q_rows < q || k < 1 || (q > 0 && (!D || !I)) || (q_rows > 0 && !out))
    return fail(B2R_EINVAL, "topk_pack: bad arguments");
  if (q_rows == 0) return B2R_OK;
  const int grid = q_rows < 148 * 8 ? q_rows : 148 * 8;
  topk_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q, q_rows, k, D, I, status, label_base, out, largest);
  B2R_CHECK_LAUNCH("topk_pack_kernel");
  return B2R_OK;
}

extern "C" int b2r_topk_merge_packed(int P, int q, int q_stride, int k, const int32_t* packed, const int64_t* bases,
                                     float* D_out, int64_t* I_out, int32_t* status_out, int largest, void* stream) {
  using namespace b2r;
  if (P < 1 || q < 0 || q_stride < q || k < 1 || !bases) return fail(B2R_EINVAL, "topk_merge_packed: bad arguments");
  if ((int64_t)P * k > 16384) return fail(B2R_EUNSUPPORTED, "topk_merge_packed: P*k must be <= 16384");
  if (q == 0) return B2R_OK;
  const size_t smem = (size_t)P * k * 4;
  static bool configured[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2R_CUDA(cudaFuncSetAttribute(topk_merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
    configured[dev & 63] = true;
  }
  topk_merge_packed_kernel<<<q, kSelThreads, smem, (cudaStream_t)stream>>>(
      P, k, packed, (size_t)q_stride * (2 * k + 1), bases, D_out, I_out, status_out, largest);
  B2R_CHECK_LAUNCH("topk_merge_packed_kernel");
  return B2R_OK;
}
