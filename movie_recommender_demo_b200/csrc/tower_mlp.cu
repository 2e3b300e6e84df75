// tower_mlp.cu — UserTower / AdTower forward in eval mode (two_tower_model.py:98-121, :167-184).
//
//   x = [EmbeddingLayer(cat) ‖ num]                          two_tower_model.py:110-113
//   Linear -> BatchNorm1d(eval) -> ReLU -> Dropout(eval = id) (x2), Linear      :83-95
//   F.normalize(p=2, dim=1, eps=1e-12)                                           :119
//
// BatchNorm is folded into the Linear weights by the host before b2r_tower_create, so each
// layer is out = act(A · Wᵀ + b): a dense contraction -> TMA-fed tcgen05.mma (kind::f16),
// fp32 accumulators in TMEM, bias/ReLU (or bias/L2-normalise) applied straight out of TMEM.
// Operands are IEEE fp16 (11-bit significand, saturating conversion), not bf16: same tensor
// throughput, 8x smaller rounding error, which keeps the unit-norm outputs within 1e-3 of the
// fp32 reference; tower activations/weights (standardised inputs, N(0,1) embeddings,
// BN-folded weights) sit far inside the fp16 range.
//
// Kernels
//   tower_gather_f16   one warp per sample: 128-bit row gathers of the F embedding rows
//                       (+ numericals), converted to fp16 and written as the layer-1 operand
//                       [B, K1p] (K padded to a multiple of 64 with zeros)
//   gemm_bias_act<BN>   128 x BN output tile per CTA step, K streamed in 64-element chunks
//                       through a 4-slot TMA/mbarrier ring; TMEM double-buffered (2 x BN cols);
//                       epilogue thread == output row, so the L2 norm of the last layer is a
//                       per-thread reduction over its TMEM lane.  Output rows leave through a
//                       per-warp shared-memory transpose: a thread owns a ROW, so storing from
//                       registers sent 32 half-used sectors per instruction (ncu: 32 sectors per
//                       request, 2x the payload in write traffic, the LSU store path at ~2.5 cycles
//                       per sector was the bound of all three layers); after the transpose 8 lanes
//                       write 128 contiguous bytes of one row (4 full lines per instruction).
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "ptx.cuh"
#include "tower_internal.h"

namespace b2r {
namespace {

// ------------------------------------------------------------------ gather ---
constexpr int kGatherWarps = 8;

__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// fp32 pair -> packed 16-bit pair.  fp16: round-to-nearest, saturating to +-65504 instead of inf (the callers
// track max |v| and raise kTowerErrSaturate, so a clipped result is never silent); bf16: fp32 range.
template <bool BF16>
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
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
__device__ __forceinline__ uint2 pack_h4(float4 v) {
  uint2 pk;
  pk.x = pack_h2<BF16>(v.x, v.y);
  pk.y = pack_h2<BF16>(v.z, v.w);
  return pk;
}
__device__ __forceinline__ float absmax4(float4 v) {
  return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
}

// out: fp16 [B, K1p]; columns [0, F*E) embeddings, [F*E, F*E+nnum) numericals, rest zero.
template <bool BF16>
__global__ void __launch_bounds__(kGatherWarps * 32)
tower_gather_f16_kernel(const float* const* __restrict__ tables, const int64_t* __restrict__ cards,
                         int F, int E4, const int64_t* __restrict__ idx, const float* __restrict__ num,
                         int nnum, int64_t B, __half* __restrict__ out, int K1p,
                         int32_t* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const int W = F * E4;
  const int64_t warp0 = (int64_t)blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kGatherWarps;
  const int tail0 = W * 4;  // first non-embedding column
  float amax = 0.f;
  bool any_bad = false;
  for (int64_t b = warp0; b < B; b += nwarps) {
    __half* orow = out + b * K1p;
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
        if (w < W) {
          if (!BF16) amax = fmaxf(amax, absmax4(v[u]));
          *reinterpret_cast<uint2*>(orow + (int64_t)w * 4) = pack_h4<BF16>(v[u]);
        }
      }
      any_bad |= bad;
    }
    for (int c = tail0 + lane; c < K1p; c += 32) {
      const int j = c - tail0;
      const float nv = j < nnum ? num[b * nnum + j] : 0.f;
      if (!BF16) amax = fmaxf(amax, fabsf(nv));
      const uint32_t h2 = pack_h2<BF16>(nv, 0.f);
      orow[c] = __ushort_as_half((unsigned short)(h2 & 0xFFFFu));
    }
  }
  int flags = any_bad ? kTowerErrIndex : 0;
  if (!BF16 && !(amax <= 65504.f)) flags |= kTowerErrSaturate;
  if (flags && err_flag) atomicOr(err_flag, flags);
}

// -------------------------------------------------------------------- GEMM ---
constexpr int kGemmThreads = 256;  // warps 0..3 control, 4..7 epilogue
constexpr int kGemmSlots = 4;

template <int BN>
struct GemmCfg {
  static constexpr int SLOT_A = 16384;
  static constexpr int SLOT_B = BN * 128;
  static constexpr int SLOT = SLOT_A + SLOT_B;
  static constexpr int NBARS = 2 * kGemmSlots + 4;
  static constexpr int STAGE = 4 * 4096;   // per epilogue warp: 32 rows x 128 B transpose buffer
  static constexpr int SMEM = 1024 + kGemmSlots * SLOT + NBARS * 8 + 64 + BN * 4 + STAGE;
};

struct GemmParams {
  int64_t M;      // rows (samples)
  int N, K;       // padded: N % BN == 0, K % 64 == 0
  const float* bias;  // [N]
  void* out;      // mode 0: fp16 [M, ldo]; mode 1: fp32 [M, ldo]
  int64_t ldo;
  int n_store;    // columns actually stored (mode 1: out_dim <= N)
  int mode;       // 0: bias + ReLU -> 16-bit   1: bias -> L2 normalise -> fp32   2: bias -> fp32 (plain linear)
  int32_t* err_flag;  // kTowerErrSaturate when an fp16 activation clipped
};

template <int BN, bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bias_act_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int NS = kGemmSlots;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar0 = base + NS * Cfg::SLOT;
  auto bar_full = [&](int i) { return bar0 + 8u * i; };
  auto bar_empty = [&](int i) { return bar0 + 8u * (NS + i); };
  auto bar_tfull = [&](int i) { return bar0 + 8u * (2 * NS + i); };
  auto bar_tempty = [&](int i) { return bar0 + 8u * (2 * NS + 2 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + NS * Cfg::SLOT + Cfg::NBARS * 8);
  float* bias_s = reinterpret_cast<float*>(smem + NS * Cfg::SLOT + Cfg::NBARS * 8 + 64);
  // transpose buffers behind the bias slice; 16-byte aligned (SLOT, NBARS*8 + 64 and BN*4 are multiples of 16)
  const uint32_t stage0 = base + NS * Cfg::SLOT + Cfg::NBARS * 8 + 64 + BN * 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull(i), 1);
      mbar_init(bar_tempty(i), 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (int)((p.M + 127) / 128);
  const int n_tiles = p.N / BN;
  const int units = m_tiles * n_tiles;
  const int KC = p.K / 64;

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int mt = u / n_tiles, nt = u % n_tiles;
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(bar_empty(slot), ph ^ 1, 11);
          mbar_arrive_expect_tx(bar_full(slot), (uint32_t)Cfg::SLOT);
          const uint32_t sa = base + slot * Cfg::SLOT;
          tma_load_2d(sa, &tmA, kc * 64, mt * 128, bar_full(slot));
          tma_load_2d(sa + Cfg::SLOT_A, &tmW, kc * 64, nt * BN, bar_full(slot));
          if (++slot == NS) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: uniform loop for the whole warp, one elected lane issues (see scan_tc.cu)
    {
      constexpr uint32_t idesc = BF16 ? umma_idesc_bf16_f32(128, BN) : umma_idesc_f16_f32(128, BN);
      const uint64_t desc0 = umma_desc_kmajor_sw128(base);
      int slot = 0, tb = 0;
      uint32_t ph = 0, tph = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        mbar_wait(bar_tempty(tb), tph ^ 1, 12);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(tb * BN);
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(bar_full(slot), ph, 13);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint64_t ad = desc0 + (uint64_t)((slot * Cfg::SLOT) >> 4);
            const uint64_t bd = ad + (uint64_t)(Cfg::SLOT_A >> 4);
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
        if (++tb == 2) { tb = 0; tph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    const int et = threadIdx.x - 128;  // 0..127
    int tb = 0;
    uint32_t tph = 0;
    float amax = 0.f;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int mt = u / n_tiles, nt = u % n_tiles;
      const int n0 = nt * BN;
      // stage this tile's bias slice (all four epilogue warps, named barrier 1)
      asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's readers are done
      for (int i = et; i < BN; i += 128) bias_s[i] = p.bias[n0 + i];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(bar_tfull(tb), tph, 14);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tb * BN);
      uint32_t r[32];
      // transpose buffer of this warp: row l at stg + l*128, its 16-byte piece j stored at slot j ^ (l & 7)
      // (a quarter-warp of 128-bit accesses then covers all 32 banks once, on the write and on the read side)
      const uint32_t stg = stage0 + (uint32_t)quarter * 4096u;
      const uint32_t st_row = stg + (uint32_t)lane * 128u;
      const int rd_r = lane >> 3, rd_j = lane & 7;    // read side: lane -> (row within a group of 4, piece)
      const int64_t row_base = (int64_t)mt * 128 + quarter * 32;
      if (p.mode == 0) {
        __half* obase = reinterpret_cast<__half*>(p.out) + n0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait_dep(r);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a = fmaxf(__uint_as_float(r[i]) + bias_s[c * 32 + i], 0.f);
            const float b = fmaxf(__uint_as_float(r[i + 1]) + bias_s[c * 32 + i + 1], 0.f);
            if (!BF16) amax = fmaxf(amax, fmaxf(a, b));
            pk[i >> 1] = pack_h2<BF16>(a, b);
          }
          // 32 fp16 = 64 B per row per chunk: pieces (c & 1) * 4 .. + 3 of the 128-byte slab row
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t piece = (uint32_t)(((c & 1) * 4 + j) ^ (lane & 7));
            st_shared_v4(st_row + piece * 16u, pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
          if (c & 1) {   // a 64-column slab is complete: 8 lanes per row, 4 rows per store instruction
            __syncwarp();
#pragma unroll
            for (int pass = 0; pass < 8; ++pass) {
              const int rr = pass * 4 + rd_r;
              const uint4 v = ld_shared_v4(stg + (uint32_t)rr * 128u + (uint32_t)((rd_j ^ (rr & 7)) * 16));
              const int64_t grow = row_base + rr;
              if (grow < p.M) *reinterpret_cast<uint4*>(obase + grow * p.ldo + (c >> 1) * 64 + rd_j * 8) = v;
            }
            __syncwarp();
          }
        }
      } else {
        float inv = 1.0f;        // mode 2: plain linear layer, fp32 out
        if (p.mode == 1) {
          float ss = 0.f;
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait_dep(r);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = __uint_as_float(r[i]) + bias_s[c * 32 + i];
              ss = fmaf(v, v, ss);
            }
          }
          inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
        }
        float* obase = reinterpret_cast<float*>(p.out) + n0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait_dep(r);
          // 32 fp32 = 128 B per row per chunk = one slab row
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int i = 4 * j;
            st_shared_v4(st_row + (uint32_t)((j ^ (lane & 7)) * 16),
                         __float_as_uint((__uint_as_float(r[i]) + bias_s[c * 32 + i]) * inv),
                         __float_as_uint((__uint_as_float(r[i + 1]) + bias_s[c * 32 + i + 1]) * inv),
                         __float_as_uint((__uint_as_float(r[i + 2]) + bias_s[c * 32 + i + 2]) * inv),
                         __float_as_uint((__uint_as_float(r[i + 3]) + bias_s[c * 32 + i + 3]) * inv));
          }
          __syncwarp();
#pragma unroll
          for (int pass = 0; pass < 8; ++pass) {
            const int rr = pass * 4 + rd_r;
            const uint4 v = ld_shared_v4(stg + (uint32_t)rr * 128u + (uint32_t)((rd_j ^ (rr & 7)) * 16));
            const int64_t grow = row_base + rr;
            if (grow < p.M) {
              const int col = n0 + c * 32 + rd_j * 4;
              float* o = obase + grow * p.ldo + c * 32 + rd_j * 4;
              // 16-byte stores need rows that start 16-byte aligned (ldo % 4 == 0, e.g. not output_dim = 50)
              if (col + 4 <= p.n_store && (p.ldo & 3) == 0) *reinterpret_cast<uint4*>(o) = v;
              else {
                if (col < p.n_store) o[0] = __uint_as_float(v.x);
                if (col + 1 < p.n_store) o[1] = __uint_as_float(v.y);
                if (col + 2 < p.n_store) o[2] = __uint_as_float(v.z);
                if (col + 3 < p.n_store) o[3] = __uint_as_float(v.w);
              }
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty(tb));
      if (++tb == 2) { tb = 0; tph ^= 1; }
    }
    if (!BF16 && !(amax <= 65504.f) && p.err_flag) atomicOr(p.err_flag, kTowerErrSaturate);
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

int make_tmap_16_impl(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !sym)
      return fail(B2R_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2R_ECUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return B2R_OK;
}

template <int BN, bool BF16>
int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p, int num_sms,
                  cudaStream_t stream) {
  auto kern = gemm_bias_act_kernel<BN, BF16>;
  static bool configured[64] = {};
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::SMEM));
    configured[dev & 63] = true;
  }
  const int64_t units = ((p.M + 127) / 128) * (p.N / BN);
  const int grid = (int)(units < num_sms ? units : num_sms);
  kern<<<grid, kGemmThreads, GemmCfg<BN>::SMEM, stream>>>(tmA, tmW, p);
  B2R_CHECK_LAUNCH("gemm_bias_act_kernel");
  return B2R_OK;
}
template <int BN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p, int num_sms,
                cudaStream_t stream, bool bf16) {
  return bf16 ? launch_gemm_t<BN, true>(tmA, tmW, p, num_sms, stream)
              : launch_gemm_t<BN, false>(tmA, tmW, p, num_sms, stream);
}

uint16_t f32_to_f16_sat(float f) {
  if (f > 65504.f) f = 65504.f;
  if (f < -65504.f) f = -65504.f;
  const __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

uint16_t f32_to_bf16(float f) {
  const __nv_bfloat16 h = __float2bfloat16_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

int pad_to(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

int make_tmap_f16_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows) {
  return make_tmap_16_impl(out, base, rows, cols, box_rows);
}

int launch_linear(const void* act16, const CUtensorMap& tmW, int64_t M, int N, int K, const float* bias, void* out,
                  int64_t ldo, int n_store, int mode, int32_t* err_flag, int num_sms, cudaStream_t stream, bool bf16) {
  if (N % 128 != 0 || K % 64 != 0) return fail(B2R_EINVAL, "linear: N must be padded to 128 and K to 64");
  CUtensorMap tmA;
  int rc = make_tmap_f16_2d(&tmA, act16, M, K, 128);
  if (rc) return rc;
  GemmParams gp;
  gp.M = M;
  gp.N = N;
  gp.K = K;
  gp.bias = bias;
  gp.out = out;
  gp.ldo = ldo;
  gp.n_store = n_store;
  gp.mode = mode;
  gp.err_flag = err_flag;
  if (mode == 1 && N > 256) return fail(B2R_EUNSUPPORTED, "linear: the normalising epilogue needs N <= 256");
  if (N % 256 == 0 && (mode != 1 || N == 256)) return launch_gemm<256>(tmA, tmW, gp, num_sms, stream, bf16);
  return launch_gemm<128>(tmA, tmW, gp, num_sms, stream, bf16);
}

}  // namespace b2r

using namespace b2r;

extern "C" {

int b2r_tower_destroy(b2r_tower* t) {
  if (!t) return B2R_OK;
  for (int l = 0; l < kTowerMaxLayers; ++l) {
    cudaFree(t->w[l]);
    cudaFree(t->wb[l]);
    cudaFree(t->b[l]);
  }
  cudaFree((void*)t->tables);
  cudaFree(t->cards);
  delete t;
  return B2R_OK;
}

int b2r_tower_create_layers(b2r_tower** out, const b2r_tower_layers* w, int device) {
  if (!out || !w) return fail(B2R_EINVAL, "tower_create: NULL argument");
  *out = nullptr;
  if (w->num_fields < 1 || w->emb_dim < 4 || w->emb_dim % 4 != 0 || w->num_numerical < 0)
    return fail(B2R_EINVAL, "tower_create: bad embedding configuration");
  const int L = w->num_layers;
  if (L < 1 || L > kTowerMaxLayers)
    return fail(B2R_EUNSUPPORTED, "tower_create: 1 to " + std::to_string(kTowerMaxLayers) +
                                      " Linear layers (hidden layers + output layer) are supported");
  if (!w->widths || !w->w || !w->b) return fail(B2R_EINVAL, "tower_create: NULL layer arrays");
  for (int l = 0; l < L; ++l)
    if (w->widths[l] < 1 || !w->w[l] || !w->b[l]) return fail(B2R_EINVAL, "tower_create: bad layer " + std::to_string(l));
  if (w->widths[L - 1] > 256)
    return fail(B2R_EUNSUPPORTED, "tower_create: layer widths must be >= 1 and out_dim <= 256");
  int ndev = 0;
  B2R_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(B2R_EINVAL, "tower_create: bad device ordinal");
  cudaDeviceProp prop;
  B2R_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(B2R_EUNSUPPORTED, "tower_create: this library is sm_100a only");
  B2R_CUDA(cudaSetDevice(device));
  b2r_tower* t = new b2r_tower();
  t->device = device;
  t->num_sms = prop.multiProcessorCount;
  t->F = w->num_fields;
  t->E = w->emb_dim;
  t->nnum = w->num_numerical;
  t->K1 = t->F * t->E + t->nnum;
  t->K1p = pad_to(t->K1, 64);
  t->L = L;
  for (int l = 0; l < L; ++l) {
    t->n[l] = w->widths[l];
    t->np[l] = pad_to(t->n[l], 128);
  }
  // the fused kernel is the two-hidden-layer chain: it keeps H1 (128 x N1p) in 128 KB of shared memory and all
  // accumulators in 512 TMEM columns; its TMA output stores need 16-byte aligned rows
  t->fused_ok = L == 3 && t->np[0] <= 512 && t->np[1] <= 256 && t->np[2] <= 256 && (t->n[2] % 4) == 0;
  memset(t->bias_host, 0, sizeof(t->bias_host));
  int rc = B2R_OK;
  float wmax = 0.f;
  int boff = 0;
  for (int l = 0; l < L && rc == B2R_OK; ++l) {
    const int kin = l == 0 ? t->K1 : t->n[l - 1];
    const int kp = l == 0 ? t->K1p : t->np[l - 1];
    const float* W = w->w[l];
    std::vector<uint16_t> wh((size_t)t->np[l] * kp, 0), wbf((size_t)t->np[l] * kp, 0);
    std::vector<float> bb((size_t)t->np[l], 0.f);
    for (int o = 0; o < t->n[l]; ++o) {
      for (int i = 0; i < kin; ++i) {
        const float v = W[(size_t)o * kin + i];
        if (!(fabsf(v) <= wmax)) wmax = fabsf(v);   // NaN propagates into wmax
        wh[(size_t)o * kp + i] = f32_to_f16_sat(v);
        wbf[(size_t)o * kp + i] = f32_to_bf16(v);
      }
      bb[o] = w->b[l][o];
    }
    if (t->fused_ok) memcpy(t->bias_host + boff, bb.data(), bb.size() * 4);
    boff += t->np[l];
    if (cudaMalloc(&t->w[l], wh.size() * 2) != cudaSuccess || cudaMalloc(&t->wb[l], wbf.size() * 2) != cudaSuccess ||
        cudaMalloc(&t->b[l], bb.size() * 4) != cudaSuccess) {
      rc = fail(B2R_ENOMEM, "tower_create: cudaMalloc failed");
      break;
    }
    cudaMemcpy(t->w[l], wh.data(), wh.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(t->wb[l], wbf.data(), wbf.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(t->b[l], bb.data(), bb.size() * 4, cudaMemcpyHostToDevice);
    const int bn = (t->np[l] % 256 == 0) ? 256 : 128;
    rc = make_tmap_f16_2d(&t->tmW[l], t->w[l], t->np[l], kp, bn);
    if (!rc) rc = make_tmap_f16_2d(&t->tmWb[l], t->wb[l], t->np[l], kp, bn);
    if (!rc && t->fused_ok) rc = make_tmap_f16_2d(&t->tmWf[l], t->w[l], t->np[l], kp, 128);
    if (!rc && t->fused_ok) rc = make_tmap_f16_2d(&t->tmWfb[l], t->wb[l], t->np[l], kp, 128);
  }
  // a BN-folded weight outside the fp16 range (tiny running_var, huge gamma): bf16 operands from the start
  if (!(wmax <= 65504.f)) t->bf16 = 1;
  if (rc == B2R_OK) {
    if (cudaMalloc((void**)&t->tables, (size_t)t->F * 8) != cudaSuccess ||
        cudaMalloc(&t->cards, (size_t)t->F * 8) != cudaSuccess)
      rc = fail(B2R_ENOMEM, "tower_create: cudaMalloc failed");
    else {
      cudaMemcpy((void*)t->tables, w->tables, (size_t)t->F * 8, cudaMemcpyHostToDevice);
      cudaMemcpy(t->cards, w->cards, (size_t)t->F * 8, cudaMemcpyHostToDevice);
    }
  }
  if (rc != B2R_OK) {
    b2r_tower_destroy(t);
    return rc;
  }
  *out = t;
  return B2R_OK;
}

int b2r_tower_create(b2r_tower** out, const b2r_tower_weights* w, int device) {
  if (!out || !w) return fail(B2R_EINVAL, "tower_create: NULL argument");
  *out = nullptr;
  if (w->hidden1 < 1 || w->hidden2 < 1 || w->out_dim < 1 || w->out_dim > 256)
    return fail(B2R_EUNSUPPORTED, "tower_create: layer widths must be >= 1 and out_dim <= 256");
  const int widths[3] = {w->hidden1, w->hidden2, w->out_dim};
  const float* W[3] = {w->w1, w->w2, w->w3};
  const float* Bv[3] = {w->b1, w->b2, w->b3};
  b2r_tower_layers lw;
  lw.num_fields = w->num_fields;
  lw.emb_dim = w->emb_dim;
  lw.num_numerical = w->num_numerical;
  lw.num_layers = 3;
  lw.cards = w->cards;
  lw.tables = w->tables;
  lw.widths = widths;
  lw.w = W;
  lw.b = Bv;
  return b2r_tower_create_layers(out, &lw, device);
}

int b2r_tower_set_param(b2r_tower* t, const char* name, double value) {
  if (!t || !name) return fail(B2R_EINVAL, "tower_set_param: NULL argument");
  const std::string n(name);
  if (n == "operand_dtype") {         // 0 fp16 | 1 bf16
    if (value != 0 && value != 1) return fail(B2R_EINVAL, "operand_dtype must be 0 (fp16) or 1 (bf16)");
    t->bf16 = (int)value;
  } else if (n == "trace_ptr") {      // debug: device pointer (as a double; < 2^53) of a [4][8][64] int64 buffer, 0 = off
    t->trace_ptr = (uintptr_t)value;
  } else if (n == "pair") {           // fused kernel: 1 = CTA pairs (cta_group::2), 0 = one CTA per tile
    t->pair = value != 0;
  } else if (n == "force_path") {     // 0 auto | 1 layer-by-layer | 2 fused
    if (value == 2 && !t->fused_ok) return fail(B2R_EUNSUPPORTED, "this tower's shape does not fit the fused kernel");
    t->force_path = (int)value;
  } else {
    return fail(B2R_EINVAL, "tower_set_param: unknown parameter " + n);
  }
  return B2R_OK;
}

double b2r_tower_get_param(const b2r_tower* t, const char* name) {
  if (!t || !name) return NAN;
  const std::string n(name);
  if (n == "operand_dtype") return t->bf16;
  if (n == "force_path") return t->force_path;
  if (n == "pair") return t->pair;
  if (n == "fused") return (t->force_path == 2 || (t->force_path == 0 && t->fused_ok)) ? 1 : 0;
  return NAN;
}

static bool tower_uses_fused(const b2r_tower* t) {
  return t->fused_ok && t->force_path != 1;
}

size_t b2r_tower_workspace(const b2r_tower* t, int64_t B) {
  if (!t || B <= 0) return 0;
  if (tower_uses_fused(t)) return 0;     // the fused kernel keeps every intermediate on the SM
  size_t s = align_up((size_t)B * t->K1p * 2, 256);               // gathered layer-1 operand
  for (int l = 0; l + 1 < t->L; ++l) s += align_up((size_t)B * t->np[l] * 2, 256);   // one buffer per hidden layer
  return s;
}

int b2r_tower_forward(b2r_tower* t, const int64_t* cat, const float* num, int64_t B, float* out,
                      int32_t* err_flag, void* workspace, size_t ws_bytes, void* stream_) {
  if (!t) return fail(B2R_EINVAL, "tower_forward: NULL handle");
  if (B < 0 || (B > 0 && (!cat || !out))) return fail(B2R_EINVAL, "tower_forward: bad arguments");
  if (t->nnum > 0 && B > 0 && !num) return fail(B2R_EINVAL, "tower_forward: numerical features required");
  if (B == 0) return B2R_OK;
  if (B > 0x7FFFFF00ll) return fail(B2R_EUNSUPPORTED, "tower_forward: batch too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  int prev = 0;
  B2R_CUDA(cudaGetDevice(&prev));
  if (prev != t->device) B2R_CUDA(cudaSetDevice(t->device));
  if (tower_uses_fused(t)) {
    const int rc = launch_tower_fused(t, cat, num, B, out, err_flag, stream);
    if (prev != t->device) cudaSetDevice(prev);
    return rc;
  }
  if (!workspace || ws_bytes < b2r_tower_workspace(t, B)) {
    if (prev != t->device) cudaSetDevice(prev);
    return fail(B2R_ENOMEM, "tower_forward: workspace too small");
  }
  const bool bf = t->bf16 != 0;
  const int L = t->L;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  __half* a1 = reinterpret_cast<__half*>(ws);
  int rc = B2R_OK;
  {
    int64_t blocks = ceil_div(B, kGatherWarps);
    const int64_t maxb = (int64_t)t->num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    if (bf)
      tower_gather_f16_kernel<true><<<(unsigned)blocks, kGatherWarps * 32, 0, stream>>>(
          t->tables, t->cards, t->F, t->E / 4, cat, num, t->nnum, B, a1, t->K1p, err_flag);
    else
      tower_gather_f16_kernel<false><<<(unsigned)blocks, kGatherWarps * 32, 0, stream>>>(
          t->tables, t->cards, t->F, t->E / 4, cat, num, t->nnum, B, a1, t->K1p, err_flag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(B2R_ECUDA, std::string("launch tower_gather_f16_kernel: ") + cudaGetErrorString(e));
    else count_launch();
  }
  // one GEMM launch per Linear: hidden layers write their ReLU'd 16-bit activations (padded columns are
  // relu(0 + 0) = 0, so they contribute nothing as the next layer's K padding); the last one normalises
  const void* act = a1;
  uint8_t* next = ws + align_up((size_t)B * t->K1p * 2, 256);
  for (int l = 0; l < L && rc == B2R_OK; ++l) {
    const bool last = l == L - 1;
    const int kp = l == 0 ? t->K1p : t->np[l - 1];
    const CUtensorMap& tmW = bf ? t->tmWb[l] : t->tmW[l];
    rc = launch_linear(act, tmW, B, t->np[l], kp, t->b[l], last ? (void*)out : (void*)next,
                       last ? t->n[l] : t->np[l], last ? t->n[l] : t->np[l], last ? 1 : 0, err_flag, t->num_sms,
                       stream, bf);
    act = next;
    next += align_up((size_t)B * t->np[l] * 2, 256);
  }
  if (prev != t->device) cudaSetDevice(prev);
  return rc;
}

}  // extern "C"
