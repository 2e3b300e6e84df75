// scan_pair.cu — the FILTER score contraction on CTA PAIRS (tcgen05 cta_group::2).
//
// Same job as scan_tc.cu's scan_tc_kernel<2, SCAN_FILTER> (S = Q · Xᵀ for faiss `IndexFlatIP::search`,
// faiss_retrieval.py:155, thresholded in the epilogue so the score matrix never reaches HBM), but two CTAs on the
// two SMs of a TPC issue ONE 256 x 256 x 16 UMMA together:
//
//   A operand : each CTA keeps ITS OWN 128-query block resident (64 KB at d = 256); the pair is one 256-query group
//   B operand : a 256-row corpus tile; each CTA streams only ITS 128-row half (16 KB per 64-element K chunk) by
//               TMA — the tensor cores of both SMs read both halves, so per SM the shared-memory traffic per
//               flop and the L2 -> SM traffic per flop are half of the one-CTA kernel's (that kernel sits at
//               160 B/clk of shared-memory traffic against the SM's 128 B/clk; this one at ~96 B/clk)
//   D         : fp32 in TMEM, per CTA 128 lanes (its queries) x 256 columns (the tile's rows), double-buffered
//
// Protocol (rank 0 of the pair = leader): both producers load their halves and post the bytes on the LEADER's
// full barrier; the leader's MMA warp issues tcgen05.mma.cta_group::2 and commits with a multicast arrive that
// frees the smem slot / publishes the accumulator in BOTH CTAs; the epilogue warps of both CTAs read their own
// TMEM and release the buffer on the leader's barrier (remote arrive for rank 1).
//
// Epilogue: 16 warps per CTA = (TMEM lane quarter) x (64-column quarter of the tile); a thread owns one query,
// a warp owns one candidate segment (query, corpus split, column quarter): appends are plain stores.
#include "internal.h"
#include "ptx.cuh"

namespace b2r {

namespace {

constexpr int kPNS = 10;                  // B ring slots (16 KB: this CTA's 128 rows x 64 K elements)
constexpr int kPNB = 2;                   // accumulator buffers (256 columns each)
constexpr int kPEW = 16;                  // epilogue warps
constexpr int kPThreads = 128 + 32 * kPEW;
constexpr int kPABytes = 4 * 16384;
constexpr int kPBBytes = kPNS * 16384;
constexpr int kPBars = 2 * kPNS + 2 * kPNB + 2;
constexpr int kPSmem = 1024 + kPABytes + kPBBytes + kPBars * 8 + 64;
static_assert(kPSmem <= 232448, "pair scan exceeds the 227 KB shared-memory limit");
constexpr int kPairRows = 2 * kTileRows;  // corpus rows per pair tile

// append the scores of one 32-row chunk that pass tau to this warp's private segment (see scan_tc.cu epi_chunk)
template <int WALK>
__device__ __forceinline__ void pair_filter_chunk(const uint32_t (&r)[32], float tau, int64_t row0, int rows_valid,
                                                  uint2* seg, int& cnt, int cap_seg) {
#pragma unroll
  for (int s = 0; s < 32; s += 8) {
    const float a = fmax3(__uint_as_float(r[s]), __uint_as_float(r[s + 1]), __uint_as_float(r[s + 2]));
    const float b = fmax3(__uint_as_float(r[s + 3]), __uint_as_float(r[s + 4]), __uint_as_float(r[s + 5]));
    const float c = fmaxf(__uint_as_float(r[s + 6]), __uint_as_float(r[s + 7]));
    const float m = fmax3(a, b, c);
    if (m >= tau) {
      auto take = [&](int i) {
        if (__uint_as_float(r[i]) >= tau) {
          if (i < rows_valid) {
            if (cnt < cap_seg) seg[cnt] = make_uint2(r[i], (uint32_t)(row0 + i));
            ++cnt;
          }
        }
      };
      if (WALK == 0) {
#pragma unroll
        for (int i = s; i < s + 8; ++i) take(i);
      } else {   // only the 3-element sub-groups whose maximum passed (row order is preserved)
        if (a >= tau) { take(s); take(s + 1); take(s + 2); }
        if (b >= tau) { take(s + 3); take(s + 4); take(s + 5); }
        if (c >= tau) { take(s + 6); take(s + 7); }
      }
    }
  }
}

// maximum of one 32-row group (threshold sampling pass); rows past the corpus end do not count
__device__ __forceinline__ float pair_group_max(const uint32_t (&r)[32], int rows_valid) {
  float m = -INFINITY;
  if (rows_valid == 32) {
    float m0 = fmax3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
    float m1 = fmax3(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]));
#pragma unroll
    for (int i = 6; i < 30; i += 4) {
      m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
      m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
    }
    m = fmax3(m0, m1, fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31])));
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < rows_valid) m = fmaxf(m, __uint_as_float(r[i]));
  }
  return m;
}

constexpr int kPairGmax = 2;   // MODE value: 0 / 1 = FILTER with the full / sub-group append walk, 2 = GMAX (sampling pass)

template <int WALK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPThreads, 1)
scan_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
                 const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base;
  const uint32_t sB = base + kPABytes;
  const uint32_t bar0 = sB + kPBBytes;
  auto bar_full = [&](int i) { return bar0 + 8u * i; };                       // used in the leader only
  auto bar_empty = [&](int i) { return bar0 + 8u * (kPNS + i); };             // per CTA
  auto bar_tfull = [&](int i) { return bar0 + 8u * (2 * kPNS + i); };         // per CTA
  auto bar_tempty = [&](int i) { return bar0 + 8u * (2 * kPNS + kPNB + i); }; // used in the leader only
  const uint32_t bar_qfull = bar0 + 8u * (2 * kPNS + 2 * kPNB);               // leader only
  const uint32_t bar_qempty = bar_qfull + 8u;                                 // per CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kPABytes + kPBBytes + kPBars * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int KC = p.d / kKChunk;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kPNS; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    for (int i = 0; i < kPNB; ++i) {
      mbar_init(bar_tfull(i), 1);
      mbar_init(bar_tempty(i), 2 * kPEW);   // every epilogue warp of BOTH CTAs
    }
    mbar_init(bar_qfull, 1);
    mbar_init(bar_qempty, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(tmem_slot));
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmX);
  }
  tc_fence_before_sync();
  cluster_sync_all();          // barriers initialised and TMEM allocated in both CTAs before any remote access
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int units = p.splits * p.QG;   // QG = groups of 256 queries (one per CTA pair and unit)

  if (warp == 0) {
    if (lane == 0) {
      // ================================================ TMA producer (both CTAs, each its own halves)
      int slot = 0;
      uint32_t ph = 0, qe_par = 0;
      int last_qg = -1;
      const uint32_t lead_qfull = mapa_u32(bar_qfull, 0);
      for (int u = cluster_id; u < units; u += num_clusters) {
        const int split = u / p.QG, qg = u % p.QG;
        const int j0 = (int)((int64_t)split * p.tile_count / p.splits);
        const int j1 = (int)((int64_t)(split + 1) * p.tile_count / p.splits);
        if (qg != last_qg) {
          if (last_qg >= 0) {  // the previous unit's MMAs must be done reading the query blocks
            mbar_wait(bar_qempty, qe_par, 41);
            qe_par ^= 1;
          }
          if (rank == 0) mbar_arrive_expect_tx(bar_qfull, (uint32_t)(2 * KC * 16384));
          for (int kc = 0; kc < KC; ++kc)
            tma_load_2d_pair(sA + kc * 16384, &tmQ, kc * kKChunk, (qg * 2 + (int)rank) * kQBlock, lead_qfull);
          last_qg = qg;
        }
        for (int j = j0; j < j1; ++j) {
          const int row = (p.tile_first + j * p.tile_stride) * kPairRows + (int)rank * kTileRows;
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(bar_empty(slot), ph ^ 1, 42);
            if (rank == 0) mbar_arrive_expect_tx(bar_full(slot), 32768u);
            tma_load_2d_pair(sB + slot * 16384, &tmX, kc * kKChunk, row, mapa_u32(bar_full(slot), 0));
            if (++slot == kPNS) { slot = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = p.idesc;   // M = 256, N = 256
      const uint64_t adesc0 = umma_desc_kmajor_sw128(sA);
      const uint64_t bdesc0 = umma_desc_kmajor_sw128(sB);
      int slot = 0, tb = 0;
      uint32_t ph = 0, tph = 0, qf_par = 0;
      int last_qg = -1;
      for (int u = cluster_id; u < units; u += num_clusters) {
        const int split = u / p.QG, qg = u % p.QG;
        const int j0 = (int)((int64_t)split * p.tile_count / p.splits);
        const int j1 = (int)((int64_t)(split + 1) * p.tile_count / p.splits);
        if (qg != last_qg) {
          mbar_wait(bar_qfull, qf_par, 43);
          qf_par ^= 1;
          last_qg = qg;
          tc_fence_after_sync();
        }
        for (int j = j0; j < j1; ++j) {
          mbar_wait(bar_tempty(tb), tph ^ 1, 44);
          tc_fence_after_sync();
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(bar_full(slot), ph, 45);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t d_tmem = tmem_base + (uint32_t)(tb * 256);
              const uint64_t ad = adesc0 + (uint64_t)((kc * 16384) >> 4);
              const uint64_t bd = bdesc0 + (uint64_t)((slot * 16384) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_pair_ss(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (uint32_t)((kc | k) != 0));
              umma_commit_pair(bar_empty(slot));   // frees the slot in both CTAs when these MMAs retire
            }
            __syncwarp();
            if (++slot == kPNS) { slot = 0; ph ^= 1; }
          }
          if (elect_one()) umma_commit_pair(bar_tfull(tb));   // accumulators complete -> both epilogues
          __syncwarp();
          if (++tb == kPNB) { tb = 0; tph ^= 1; }
        }
        const int nu = u + num_clusters;
        if (nu < units && (nu % p.QG) != qg) {
          if (elect_one()) umma_commit_pair(bar_qempty);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================== epilogue (both CTAs)
    const int e = warp - 4;
    const int quarter = e & 3;   // == warp % 4: the TMEM lane quarter this warp may read
    const int cq = e >> 2;       // 64-column quarter of the 256-column tile
    int tb = 0;
    uint32_t tph = 0;
    for (int u = cluster_id; u < units; u += num_clusters) {
      const int split = u / p.QG, qg = u % p.QG;
      const int j0 = (int)((int64_t)split * p.tile_count / p.splits);
      const int j1 = (int)((int64_t)(split + 1) * p.tile_count / p.splits);
      const int qbase = (qg * 2 + (int)rank) * kQBlock + quarter * 32;
      const int q = qbase + lane;
      const int segi = split * 4 + cq;
      uint2* seg = WALK == kPairGmax ? nullptr : p.cand + ((size_t)q * p.nseg + segi) * p.cap_seg;
      int cnt = 0;
      const bool warp_active = qbase < p.Q;
      const float tau = WALK == kPairGmax ? 0.f : p.tau[q];
      for (int j = j0; j < j1; ++j) {
        const uint32_t lead_tempty = mapa_u32(bar_tempty(tb), 0);
        mbar_wait(bar_tfull(tb), tph, 46);
        tc_fence_after_sync();
        if (warp_active) {
          const int64_t row_base = (int64_t)(p.tile_first + j * p.tile_stride) * kPairRows + cq * 64;
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tb * 256 + cq * 64);
          // both 32-column chunks go to registers first and the buffer is released BEFORE any score is looked at:
          // the leader's MMA warp waits for 32 warps of two SMs (remote arrives), so every cycle between
          // "accumulator full" and "accumulator free" is a cycle the tensor cores may idle
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr, r0);
          tmem_ld_32x32(taddr + 32, r1);
          tmem_ld_wait_dep(r0);
          tmem_ld_wait_dep(r1);
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead_tempty);
          const int64_t rv0 = p.N - row_base, rv1 = rv0 - 32;
          const int valid0 = rv0 >= 32 ? 32 : (rv0 < 0 ? 0 : (int)rv0);
          const int valid1 = rv1 >= 32 ? 32 : (rv1 < 0 ? 0 : (int)rv1);
          if (WALK == kPairGmax) {
            // sampling pass: the maxima of this warp's two 32-row groups of pair tile j -> gmax[q, j * 8 + cq * 2 ..]
            float* dst = p.gmax + (size_t)q * p.gstride + j * 8 + cq * 2;      // 8-byte aligned (gstride % 8 == 0)
            *reinterpret_cast<float2*>(dst) = make_float2(pair_group_max(r0, valid0), pair_group_max(r1, valid1));
          } else {
            pair_filter_chunk<WALK == kPairGmax ? 0 : WALK>(r0, tau, row_base, valid0, seg, cnt, p.cap_seg);
            pair_filter_chunk<WALK == kPairGmax ? 0 : WALK>(r1, tau, row_base + 32, valid1, seg, cnt, p.cap_seg);
          }
        } else {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead_tempty);
        }
        if (++tb == kPNB) { tb = 0; tph ^= 1; }
      }
      if (WALK != kPairGmax && warp_active) p.cand_count[(size_t)q * p.nseg + segi] = cnt;
    }
  }
  __syncwarp();
  tc_fence_before_sync();
  cluster_sync_all();            // neither CTA may leave (or free TMEM) while the pair's MMAs / remote arrives are in flight
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base);
  }
}

}  // namespace

// p.QG = groups of 256 queries, p.tile_count = 256-row pair tiles, p.idesc for M = N = 256.
// mode SCAN_FILTER: p.nseg = 4 * p.splits; mode SCAN_GMAX: p.gmax [Qpad, gstride], gstride = 8 * tile_count.
int launch_scan_pair(const CUtensorMap& tmQ, const CUtensorMap& tmX, const ScanParams& p, int num_sms,
                     cudaStream_t stream, int walk, int mode) {
  if (p.d % kKChunk != 0 || p.d < kKChunk || p.d > 256)
    return fail(B2R_EINVAL, "scan: d must be a multiple of 64 in [64,256]");
  if (p.tile_count <= 0 || p.Q <= 0) return B2R_OK;
  if (mode == SCAN_FILTER && p.nseg != 4 * p.splits) return fail(B2R_EINVAL, "pair scan: nseg must be 4 * splits");
  if (mode != SCAN_FILTER && mode != SCAN_GMAX) return fail(B2R_EINVAL, "pair scan: FILTER or GMAX");
  if (mode == SCAN_GMAX && p.gstride < 8 * p.tile_count) return fail(B2R_EINVAL, "pair scan: gstride too small");
  const int64_t units = (int64_t)p.splits * p.QG;
  const int clusters = (int)(units < num_sms / 2 ? units : num_sms / 2);
  const int w = mode == SCAN_GMAX ? kPairGmax : (walk ? 1 : 0);
  static bool configured[3][64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (!configured[w][dev & 63]) {
    if (w == 2) B2R_CUDA(cudaFuncSetAttribute(scan_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
    else if (w == 1) B2R_CUDA(cudaFuncSetAttribute(scan_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
    else B2R_CUDA(cudaFuncSetAttribute(scan_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
    configured[w][dev & 63] = true;
  }
  if (w == 2) scan_pair_kernel<2><<<2 * clusters, kPThreads, kPSmem, stream>>>(tmQ, tmX, p);
  else if (w == 1) scan_pair_kernel<1><<<2 * clusters, kPThreads, kPSmem, stream>>>(tmQ, tmX, p);
  else scan_pair_kernel<0><<<2 * clusters, kPThreads, kPSmem, stream>>>(tmQ, tmX, p);
  B2R_CHECK_LAUNCH("scan_pair_kernel");
  return B2R_OK;
}

}  // namespace b2r
