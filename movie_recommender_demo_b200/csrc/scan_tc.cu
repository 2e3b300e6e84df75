// scan_tc.cu — the query x corpus score contraction on 5th-gen tensor cores.
//
// Replaces the arithmetic of faiss `IndexFlatIP::search` as called at
// faiss_retrieval.py:155: S = Q · Xᵀ, never written to HBM in the
// product modes.
//
//   A operand (UMMA M = 128 TMEM lanes) : 128 queries per block, MQ (1|2) blocks resident in smem
//   B operand (UMMA N = 128 columns)    : 128-row corpus tile, streamed by TMA in 64-element
//                                          K chunks (128 B rows, SWIZZLE_128B) through an mbarrier ring
//   D accumulators                      : fp32 in TMEM, 4/MQ buffers x MQ blocks x 128 columns = 512 cols
//
// Queries sit on the lane side so that each epilogue thread owns ONE query: its candidate
// threshold is a scalar register and the per-score work is a 3-input max plus a rare branch.
//
// Warp roles (128 + 32*EW threads, 1 CTA/SM, persistent over "units" = (corpus split, query group)):
//   warp 0 lane 0 : TMA producer       warp 1 lane 0 : tcgen05.mma issuer
//   warp 2        : TMEM alloc/dealloc  warp 3        : idle
//   warps 4..4+EW : epilogue (tcgen05.ld 32x32b.x32 -> registers), TMEM quarter = warp % 4
// EW = 8 epilogue warps by default.  With two query blocks (MQ = 2) each of those warps walks all 128
// columns of its block; under a realistic hit rate (~0.1 % of the scores pass the threshold) that
// serial instruction stream (~550 instructions per tile at ~6 cycles each) is as long as the tile's MMA
// time, so the FILTER mode also exists with EW = 16: two warps per (query block, lane quarter), each
// owning one 64-column half and its own candidate segment.
//
// Epilogue modes:
//   SCAN_DUMP   : store every score                     (tests, small-corpus dense path)
//   SCAN_GMAX   : max over each 32-row group            (threshold sampling pass)
//   SCAN_FILTER : append (score,row) with score >= tau  (the product scan)
#include "internal.h"
#include "ptx.cuh"

namespace b2r {

namespace {

template <int MQ>
struct ScanCfg {
  static constexpr int NS = (MQ == 1) ? 9 : 6;   // B ring slots (16 KB each)
  static constexpr int NB = 4 / MQ;              // TMEM accumulator buffers
  static constexpr int A_BYTES = MQ * 4 * 16384;
  static constexpr int B_BYTES = NS * 16384;
  static constexpr int NBARS = 2 * NS + 2 * NB + 2;
  static constexpr int SMEM = 1024 + A_BYTES + B_BYTES + NBARS * 8 + 64;
};

constexpr int kEpiWarp0 = 4;

template <int MODE, int WALK>
__device__ __forceinline__ void epi_chunk(const ScanParams& p, const uint32_t (&r)[32], int q,
                                          float tau, int64_t row0, int rows_valid, float& gmax_out,
                                          uint2* seg, int& cnt) {
  // r[i] = score(query q, corpus row row0 + i); rows_valid = number of i with row0+i < N (<=32)
  if (MODE == SCAN_DUMP) {
    if (q < p.Q) {
      float* dst = p.dump + (size_t)q * p.ld + row0;
      if (rows_valid == 32 && ((p.ld & 3) == 0)) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 v = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                 __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
          *reinterpret_cast<float4*>(dst + i) = v;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < rows_valid) dst[i] = __uint_as_float(r[i]);
      }
    }
  } else if (MODE == SCAN_GMAX) {
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
    gmax_out = m;   // stored by the caller, several chunks per store instruction
  } else {  // SCAN_FILTER
    // per-lane test of 8 scores at a time (3-input max tree): ~0.6 instructions per score on the
    // common path.  A lane with a hit walks its 8 values and appends to ITS query's private
    // segment: one 8-byte store and a register increment, no atomics.
#pragma unroll
    for (int s = 0; s < 32; s += 8) {
      const float a = fmax3(__uint_as_float(r[s]), __uint_as_float(r[s + 1]), __uint_as_float(r[s + 2]));
      const float b = fmax3(__uint_as_float(r[s + 3]), __uint_as_float(r[s + 4]), __uint_as_float(r[s + 5]));
      const float c = fmaxf(__uint_as_float(r[s + 6]), __uint_as_float(r[s + 7]));
      const float m = fmax3(a, b, c);
      if (m >= tau) {
        if (WALK == 0) {
#pragma unroll
          for (int i = s; i < s + 8; ++i) {
            if (__uint_as_float(r[i]) >= tau && i < rows_valid) {
              if (cnt < p.cap_seg) seg[cnt] = make_uint2(r[i], (uint32_t)(row0 + i));
              ++cnt;
            }
          }
        } else {
          // walk only the sub-group(s) whose maximum passed (row order is preserved: a, b, c)
          auto take = [&](int i) {
            if (__uint_as_float(r[i]) >= tau) {
              if (i < rows_valid) {
                if (cnt < p.cap_seg) seg[cnt] = make_uint2(r[i], (uint32_t)(row0 + i));
                ++cnt;
              }
            }
          };
          if (a >= tau) { take(s); take(s + 1); take(s + 2); }
          if (b >= tau) { take(s + 3); take(s + 4); take(s + 5); }
          if (c >= tau) { take(s + 6); take(s + 7); }
        }
      }
    }
  }
}

template <int MQ, int MODE, int EW, int WALK>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
               const ScanParams p) {
  using Cfg = ScanCfg<MQ>;
  constexpr int NS = Cfg::NS, NB = Cfg::NB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B operands need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base;
  const uint32_t sB = base + Cfg::A_BYTES;
  const uint32_t bar0 = sB + Cfg::B_BYTES;
  auto bar_full = [&](int i) { return bar0 + 8u * i; };
  auto bar_empty = [&](int i) { return bar0 + 8u * (NS + i); };
  auto bar_tfull = [&](int i) { return bar0 + 8u * (2 * NS + i); };
  auto bar_tempty = [&](int i) { return bar0 + 8u * (2 * NS + NB + i); };
  const uint32_t bar_qfull = bar0 + 8u * (2 * NS + 2 * NB);
  const uint32_t bar_qempty = bar_qfull + 8u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::A_BYTES + Cfg::B_BYTES + Cfg::NBARS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KC = p.d / kKChunk;  // K chunks per tile (<= 4)

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(bar_tfull(i), 1);
      mbar_init(bar_tempty(i), EW);  // one arrive per epilogue warp
    }
    mbar_init(bar_qfull, 1);
    mbar_init(bar_qempty, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmX);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int units = p.splits * p.QG;

  if (warp == 0) {
    if (lane == 0) {
      // ================================================ TMA producer
      int slot = 0;
      uint32_t ph = 0, qe_par = 0;
      int last_qg = -1;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = u / p.QG, qg = u % p.QG;
        const int j0 = (int)((int64_t)split * p.tile_count / p.splits);
        const int j1 = (int)((int64_t)(split + 1) * p.tile_count / p.splits);
        if (qg != last_qg) {
          if (last_qg >= 0) {  // previous unit's MMAs must be done reading the query blocks
            mbar_wait(bar_qempty, qe_par, 1);
            qe_par ^= 1;
          }
          mbar_arrive_expect_tx(bar_qfull, (uint32_t)(MQ * KC * 16384));
          for (int h = 0; h < MQ; ++h)
            for (int kc = 0; kc < KC; ++kc)
              tma_load_2d(sA + (h * 4 + kc) * 16384, &tmQ, kc * kKChunk,
                          (qg * MQ + h) * kQBlock, bar_qfull);
          last_qg = qg;
        }
        for (int j = j0; j < j1; ++j) {
          const int row = (p.tile_first + j * p.tile_stride) * kTileRows;
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(bar_empty(slot), ph ^ 1, 2);
            mbar_arrive_expect_tx(bar_full(slot), 16384u);
            tma_load_2d(sB + slot * 16384, &tmX, kc * kKChunk, row, bar_full(slot));
            if (++slot == NS) { slot = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================== MMA issuer
    // The whole warp walks the loop (uniform control flow, every lane polls the barriers); one
    // elected lane issues the tcgen05.mma / commit.  Descriptors are base + 16-byte-unit offsets.
    {
      const uint32_t idesc = p.idesc;
      const uint64_t adesc0 = umma_desc_kmajor_sw128(sA);
      const uint64_t bdesc0 = umma_desc_kmajor_sw128(sB);
      int slot = 0, tb = 0;
      uint32_t ph = 0, tph = 0, qf_par = 0;
      int last_qg = -1;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = u / p.QG, qg = u % p.QG;
        const int j0 = (int)((int64_t)split * p.tile_count / p.splits);
        const int j1 = (int)((int64_t)(split + 1) * p.tile_count / p.splits);
        if (qg != last_qg) {
          mbar_wait(bar_qfull, qf_par, 3);
          qf_par ^= 1;
          last_qg = qg;
          tc_fence_after_sync();
        }
        for (int j = j0; j < j1; ++j) {
          mbar_wait(bar_tempty(tb), tph ^ 1, 4);
          tc_fence_after_sync();
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(bar_full(slot), ph, 5);
            tc_fence_after_sync();
            if (elect_one()) {
#pragma unroll
              for (int h = 0; h < MQ; ++h) {
                const uint32_t d_tmem = tmem_base + (uint32_t)((tb * MQ + h) * 128);
                const uint64_t ad = adesc0 + (uint64_t)(((h * 4 + kc) * 16384) >> 4);
                const uint64_t bd = bdesc0 + (uint64_t)((slot * 16384) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)  // 4 x K=16 per 64-element chunk: +32 B inside the swizzle span
                  umma_bf16_ss(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc,
                               (uint32_t)((kc | k) != 0));
              }
              umma_commit(bar_empty(slot));  // frees the smem slot when these MMAs retire
            }
            __syncwarp();
            if (++slot == NS) { slot = 0; ph ^= 1; }
          }
          if (elect_one()) umma_commit(bar_tfull(tb));  // accumulator complete -> epilogue
          __syncwarp();
          if (++tb == NB) { tb = 0; tph ^= 1; }
        }
        const int nu = u + gridDim.x;
        if (nu < units && (nu % p.QG) != qg) {
          if (elect_one()) umma_commit(bar_qempty);
          __syncwarp();
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ================================================== epilogue
    const int e = warp - kEpiWarp0;
    const int quarter = e & 3;  // == warp % 4: the TMEM lane quarter this warp may read
    const int g = e >> 2;
    // a warp owns (query block h, 64-column half) when the columns are split, else (block h, all 128 columns)
    constexpr bool kSplitCols = (MQ == 1) || (EW == 16);
    const int h = (MQ == 2) ? (g & 1) : 0;
    const int colhalf = (MQ == 2) ? (g >> 1) : g;
    const int col_begin = kSplitCols ? colhalf * 64 : 0;
    constexpr int NCH = kSplitCols ? 2 : 4;  // 32-column chunks per tile for this warp
    int tb = 0;
    uint32_t tph = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int split = u / p.QG, qg = u % p.QG;
      const int j0 = (int)((int64_t)split * p.tile_count / p.splits);
      const int j1 = (int)((int64_t)(split + 1) * p.tile_count / p.splits);
      const int q = (qg * MQ + h) * kQBlock + quarter * 32 + lane;
      // this warp's private candidate segment for (query q, split[, column half g])
      const int segi = kSplitCols ? split * 2 + colhalf : split;
      uint2* seg = nullptr;
      int cnt = 0;
      if (MODE == SCAN_FILTER) seg = p.cand + ((size_t)q * p.nseg + segi) * p.cap_seg;
      const bool warp_active = ((qg * MQ + h) * kQBlock + quarter * 32) < p.Q;
      float tau = INFINITY;
      if (MODE == SCAN_FILTER) tau = p.tau[q];
      for (int j = j0; j < j1; ++j) {
        mbar_wait(bar_tfull(tb), tph, 6);
        tc_fence_after_sync();
        if (warp_active) {
          const int tile = p.tile_first + j * p.tile_stride;
          const int64_t tile_row0 = (int64_t)tile * kTileRows;
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                 (uint32_t)((tb * MQ + h) * 128 + col_begin);
          uint32_t r0[32], r1[32];
          // GMAX: a lane owns a QUERY, so its group maxima go to its own row of p.gmax - 32 different
          // sectors per store instruction.  Keep the tile's maxima in registers and store them 8 or 16
          // bytes at a time (4x fewer store sectors than one 4-byte store per chunk).
          float gm0 = 0.f, gm1 = 0.f, gm_prev0 = 0.f, gm_prev1 = 0.f;
          if (MODE == SCAN_FILTER && NCH == 2 && p.early_release) {
            // both chunks to registers, release the accumulator buffer, THEN look at the scores: the MMA warp
            // gets the buffer back one chunk-processing time (~800 cycles) earlier
            tmem_ld_32x32(taddr, r0);
            tmem_ld_32x32(taddr + 32, r1);
            tmem_ld_wait_dep(r0);
            tmem_ld_wait_dep(r1);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(tb));
            {
              const int64_t row0 = tile_row0 + col_begin;
              const int64_t rv = p.N - row0;
              epi_chunk<MODE, WALK>(p, r0, q, tau, row0, rv >= 32 ? 32 : (rv < 0 ? 0 : (int)rv), gm0, seg, cnt);
            }
            {
              const int64_t row0 = tile_row0 + col_begin + 32;
              const int64_t rv = p.N - row0;
              epi_chunk<MODE, WALK>(p, r1, q, tau, row0, rv >= 32 ? 32 : (rv < 0 ? 0 : (int)rv), gm1, seg, cnt);
            }
            if (++tb == NB) { tb = 0; tph ^= 1; }
            continue;
          }
          tmem_ld_32x32(taddr, r0);
#pragma unroll 1
          for (int c = 0; c < NCH; c += 2) {
            // chunk c is in flight into r0; overlap chunk c+1 (r1) with its processing
            tmem_ld_wait_dep(r0);
            tmem_ld_32x32(taddr + (c + 1) * 32, r1);
            {
              const int64_t row0 = tile_row0 + col_begin + c * 32;
              const int64_t rv = p.N - row0;
              const int rows_valid = rv >= 32 ? 32 : (rv < 0 ? 0 : (int)rv);
              epi_chunk<MODE, WALK>(p, r0, q, tau, row0, rows_valid, gm0, seg, cnt);
            }
            tmem_ld_wait_dep(r1);
            if (c + 2 < NCH) {
              tmem_ld_32x32(taddr + (c + 2) * 32, r0);
            } else {
              // all TMEM reads of this buffer by this warp are complete: release it
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tempty(tb));
            }
            {
              const int64_t row0 = tile_row0 + col_begin + (c + 1) * 32;
              const int64_t rv = p.N - row0;
              const int rows_valid = rv >= 32 ? 32 : (rv < 0 ? 0 : (int)rv);
              epi_chunk<MODE, WALK>(p, r1, q, tau, row0, rows_valid, gm1, seg, cnt);
            }
            if (MODE == SCAN_GMAX) {
              float* dst = p.gmax + (size_t)q * p.gstride + j * 4 + (col_begin >> 5);   // 16-byte aligned (gstride % 4 == 0)
              if (NCH == 2) {
                *reinterpret_cast<float2*>(dst) = make_float2(gm0, gm1);
              } else if (c == 0) {
                gm_prev0 = gm0;
                gm_prev1 = gm1;
              } else {
                *reinterpret_cast<float4*>(dst) = make_float4(gm_prev0, gm_prev1, gm0, gm1);
              }
            }
          }
        } else {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty(tb));
        }
        if (++tb == NB) { tb = 0; tph ^= 1; }
      }
      if (MODE == SCAN_FILTER && warp_active) p.cand_count[(size_t)q * p.nseg + segi] = cnt;
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

template <int MQ, int MODE, int EW = 8, int WALK = 0>
int launch_one(const CUtensorMap& tmQ, const CUtensorMap& tmX, const ScanParams& p, int grid,
               cudaStream_t stream) {
  auto kern = scan_tc_kernel<MQ, MODE, EW, WALK>;
  static bool configured[64] = {};  // per instantiation, per device
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ScanCfg<MQ>::SMEM));
    configured[dev & 63] = true;
  }
  kern<<<grid, 128 + 32 * EW, ScanCfg<MQ>::SMEM, stream>>>(tmQ, tmX, p);
  B2R_CHECK_LAUNCH("scan_tc_kernel");
  return B2R_OK;
}

}  // namespace

int make_tmap_bf16_rows(CUtensorMap* out, const void* base, int64_t rows, int d) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(B2R_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  if (rows <= 0) rows = 1;
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {(cuuint32_t)kKChunk, (cuuint32_t)kTileRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B2R_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return B2R_OK;
}

void plan_scan(int Q, int tile_count, int num_sms, int* MQ, int* QG, int* splits) {
  const int mq = (Q > kQBlock) ? 2 : 1;
  const int qg = (int)ceil_div(Q, kQBlock * mq);
  // choose the number of corpus splits so that units = qg * splits fills whole waves
  int best_s = 1;
  double best_eff = -1.0;
  const int smax = tile_count < 1 ? 1 : tile_count;
  for (int s = 1; s <= smax && s <= 4 * num_sms; ++s) {
    const int64_t units = (int64_t)qg * s;
    if (s > 1 && tile_count / s < 4) break;  // keep >= 4 tiles per unit (pipeline fill cost)
    const int64_t waves = ceil_div(units, num_sms);
    double eff = (double)units / (double)(waves * num_sms);
    // unit granularity: ceil/floor imbalance of tiles per split
    const double per = (double)tile_count / s;
    eff *= per / (double)ceil_div(tile_count, s);
    // fixed cost per unit (query-block reload + pipeline fill/drain), about 1.5 tile times: irrelevant for
    // the full scan (hundreds of tiles per unit), decisive for the short sampling pass, where 4 waves of
    // 7-tile units took 95 us at Q=4096 against ~55 us for one wave of 28-tile units
    eff *= per / (per + 1.5);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_s = s; }
  }
  *MQ = mq;
  *QG = qg;
  *splits = best_s;
}

int launch_scan(int mode, int MQ, const CUtensorMap& tmQ, const CUtensorMap& tmX,
                const ScanParams& p, int num_sms, cudaStream_t stream, int epi_warps, int walk) {
  if (p.d % kKChunk != 0 || p.d < kKChunk || p.d > 256)
    return fail(B2R_EINVAL, "scan: d must be a multiple of 64 in [64,256]");
  if (p.tile_count <= 0 || p.Q <= 0) return B2R_OK;
  const int64_t units = (int64_t)p.splits * p.QG;
  const int grid = (int)(units < num_sms ? units : num_sms);
  if (MQ == 2 && mode == SCAN_FILTER && epi_warps == 16)
    return walk ? launch_one<2, SCAN_FILTER, 16, 1>(tmQ, tmX, p, grid, stream)
                : launch_one<2, SCAN_FILTER, 16, 0>(tmQ, tmX, p, grid, stream);
  if (MQ == 1 && mode == SCAN_FILTER && walk) return launch_one<1, SCAN_FILTER, 8, 1>(tmQ, tmX, p, grid, stream);
#define B2R_SCAN_CASE(MQ_, MODE_)                                             \
  if (MQ == MQ_ && mode == MODE_) return launch_one<MQ_, MODE_>(tmQ, tmX, p, grid, stream);
  B2R_SCAN_CASE(1, SCAN_DUMP)
  B2R_SCAN_CASE(1, SCAN_GMAX)
  B2R_SCAN_CASE(1, SCAN_FILTER)
  B2R_SCAN_CASE(2, SCAN_DUMP)
  B2R_SCAN_CASE(2, SCAN_GMAX)
  B2R_SCAN_CASE(2, SCAN_FILTER)
#undef B2R_SCAN_CASE
  return fail(B2R_EINVAL, "scan: bad mode/MQ");
}

}  // namespace b2r
