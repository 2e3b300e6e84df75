// tower_fused.cu — UserTower / AdTower forward (two_tower_model.py:98-121, :167-184, eval mode) as ONE
// persistent kernel: per 128-sample tile
//
//   gather warps  : 26 x 64-byte embedding rows per sample (+ numericals) -> 16-bit, written straight into
//                   shared memory in the K-major 128-byte-swizzled layout a tcgen05 A operand wants
//                   (two_tower_model.py:110-113: x = [EmbeddingLayer(cat) || num]); 64-column chunks through
//                   a 3-slot ring, so the gather of the next chunks overlaps the MMAs of the current one
//   GEMM1 (tcgen05, A = gathered chunk, B = W1 tiles streamed by TMA) -> fp32 accumulators in TMEM (<= 512 cols)
//   epilogue 1    : TMEM -> +bias -> ReLU -> 16-bit -> H1 in shared memory, again as a K-major swizzled A operand
//   GEMM2 (A = H1 chunks as they appear, B = W2 tiles) -> TMEM -> epilogue 2 -> H2 in shared memory
//   GEMM3 (A = H2, B = W3 tiles) -> TMEM -> epilogue 3: bias, row L2 norm (F.normalize, eps 1e-12), fp32,
//                   staged through shared memory and written with TMA stores (full 128-byte lines)
//
// Nothing but the input ids/numericals, the table rows and the fp32 output touches HBM: the 16-bit layer-1
// operand and H1 / H2 (the layer-by-layer path's [B,448] + [B,512] + [B,256] round trips) never leave the SM.
// BatchNorm is folded into W/b on the host (tower_mlp.cu).  Dropout is the identity in eval mode.
//
// Warp roles (640 threads, 1 CTA / SM, persistent over row tiles):
//   warp 0 lane 0 : TMA producer of the weight tiles (W1, W2, W3 in the exact order the MMAs consume them)
//   warp 1        : tcgen05.mma issuer (one elected lane)
//   warp 2        : TMEM alloc / dealloc            warp 3 : L2 prefetch of the next tile's ids
//   warps 4..11   : epilogue, two warps per TMEM lane quarter (each owns one 32-column half of every 64-column chunk)
//   warps 12..19  : gather
//
// Shared memory (1024-byte aligned): H region 128 KB (H1 = 8 chunks [128 rows x 64 cols]; H2 re-uses chunks 0..3;
// the fp32 output staging re-uses chunks 4..7) | A ring 3 x 16 KB | W ring 3 x 16 KB | barriers.
// TMEM (512 columns): GEMM1 accumulators [0, N1p); GEMM2 re-uses [0, N2p) once epilogue 1 has drained them;
// GEMM3 uses [256, 256 + N3p).
//
// Roofline (BASELINE config 4: B = 65536, 26 x 10M-row tables, 429 -> 512 -> 256 -> 256): algorithmic bytes
// B*(26*64 + 26*8 + 13*4 + 256*4) = 193 MB (30 us at the measured copy bandwidth); 54.6 GFLOP (33-39 us at the
// measured cuBLAS bf16 rate).  The weight tiles (842 KB per 128-sample tile, 431 MB per batch) stream from L2.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "tower_internal.h"

namespace b2r {
namespace {

constexpr int kFtThreads = 640;
constexpr int kFtEpiWarp0 = 4, kFtEpiWarps = 8;
constexpr int kFtGatherWarp0 = 12, kFtGatherWarps = 8;
#ifndef B2R_FT_NA
#define B2R_FT_NA 3
#endif
#ifndef B2R_FT_NW
#define B2R_FT_NW 3
#endif
constexpr int kNA = B2R_FT_NA, kNW = B2R_FT_NW;   // ring depths (16 KB slots); -D overrides are for A/B builds only
constexpr int kHBytes = 131072;
constexpr int kSlot = 16384;
constexpr int kMaxKC = 8;                     // 64-column chunks of H1 (N1p <= 512)

struct FtBars {   // mbarrier indices
  static constexpr int a_full = 0;                    // [kNA]
  static constexpr int a_empty = a_full + kNA;        // [kNA]
  static constexpr int w_full = a_empty + kNA;        // [kNW]
  static constexpr int w_empty = w_full + kNW;        // [kNW]
  static constexpr int acc_full = w_empty + kNW;      // [3]   accumulators of GEMM1/2/3 complete
  static constexpr int h1_ready = acc_full + 3;       // [kMaxKC]
  static constexpr int h2_ready = h1_ready + kMaxKC;  // [kMaxKC / 2]
  static constexpr int tmem_free = h2_ready + kMaxKC / 2;  // epilogues 2 + 3 have read their accumulators
  static constexpr int count = tmem_free + 1;
};
constexpr int kFtSmem = 1024 + kHBytes + (kNA + kNW) * kSlot + FtBars::count * 8 + 64;
static_assert(kFtSmem <= 232448, "fused tower kernel exceeds the 227 KB shared-memory limit");

struct FtBias {
  float v[1024];   // b1 at 0, b2 at N1p, b3 at N1p + N2p (zero padded)
};

struct FtParams {
  const float* const* tables;
  const int64_t* cards;
  const int64_t* cat;     // [B, F]
  const float* num;       // [B, nnum] or null
  int64_t B;
  int F, E4, nnum;
  int KC1;                // layer-1 K chunks (K1p / 64)
  int N1p, N2p, N3p;      // padded widths (multiples of 128; <= 512 / 256 / 256)
  int32_t* err_flag;
  long long* trace;       // debug: [4 CTAs][8 tiles][64 events] clock64 stamps (null in production)
};

// debug timeline (set_param trace_ptr): event `slot` of local tile `it`, first 4 CTAs only
#define FT_TRACE(it_, slot_)                                                                     \
  do {                                                                                           \
    if (p.trace && blockIdx.x < 4 && (it_) < 8 && lane == 0)                                     \
      p.trace[((size_t)blockIdx.x * 8 + (it_)) * 64 + (slot_)] = clock64();                      \
  } while (0)

__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// fp32 pair -> packed 16-bit pair.  fp16: round-to-nearest, saturating to +-65504 (the caller tracks |v| and
// raises kTowerErrSaturate); bf16: round-to-nearest-even, fp32 range.
// ReLU fused into the conversion (cvt.rn.relu): max(x, 0) -> 16-bit pair
template <bool BF16>
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
  uint32_t r;
  if (BF16) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
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
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// TMA store shared -> global (2D tile), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// byte offset of (row r, 16-byte piece j) inside a [128 rows x 128 B] K-major tile with the 128-byte swizzle
// (what TMA SWIZZLE_128B writes and the UMMA descriptor of ptx.cuh reads): 8-row groups of 1024 B, piece j of
// row r stored at position j ^ (r & 7)
__device__ __forceinline__ uint32_t sw128_off(int r, int j) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}

// PAIR: two CTAs of a cluster (the two SMs of a TPC) run ONE tcgen05.mma.cta_group::2 stream: each CTA gathers,
// holds and post-processes ITS OWN 128-sample tile, but a weight tile is 256 output columns of which each CTA
// loads only its 128-row half -- half the weight bytes per sample through L2 -> SM and through the 3-slot ring
// (the ring depth, not the tensor pipe, paces the one-CTA kernel: DESIGN.md 3.5).  Rank 0 issues the MMAs.
template <bool BF16, bool PAIR>
__global__ void __launch_bounds__(kFtThreads, 1)
tower_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                   const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmOut,
                   const __grid_constant__ FtBias bias, const FtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sH = base;
  const uint32_t sA = base + kHBytes;
  const uint32_t sW = sA + kNA * kSlot;
  const uint32_t bar0 = sW + kNW * kSlot;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kHBytes + (kNA + kNW) * kSlot + FtBars::count * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kNTShift = PAIR ? 8 : 7;            // output columns per weight tile: 256 (pair) | 128
  constexpr int kNTCols = 1 << kNTShift;
  const int NT1 = p.N1p >> kNTShift, NT2 = p.N2p >> kNTShift, NT3 = p.N3p >> kNTShift;
  const int KC1 = p.KC1, KC2 = p.N1p >> 6, KC3 = p.N2p >> 6;
  const int tiles = (int)((p.B + 127) >> 7);
  // work distribution: `tile0` = this CTA's first 128-sample tile, `tstep` = distance to its next one.  A pair
  // walks tile pairs (2j, 2j + 1); both CTAs of a pair run the same number of rounds (an odd tile count leaves
  // rank 1 an all-padding last tile: zero rows in, clipped stores out).
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) * 2 + (int)rank : (int)blockIdx.x;
  const int tstep = PAIR ? (int)(gridDim.x >> 1) * 2 : (int)gridDim.x;
  const int tiles_loop = PAIR ? ((tiles + 1) & ~1) : tiles;     // loop bound: `for (tile = tile0; tile < tiles_loop; tile += tstep)`
  const bool leader = rank == 0;
  // barriers the MMA warp waits on live in the leader; the other CTA's producers arrive remotely
  auto lead = [&](uint32_t b) { return PAIR ? mapa_u32(b, 0) : b; };
  auto arrive_lead = [&](uint32_t b) {
    if (PAIR) mbar_arrive_cluster(mapa_u32(b, 0));
    else mbar_arrive(b);
  };
  auto wait_lead = [&](uint32_t b, uint32_t parity, int tag) {     // leader-owned barrier with remote arrivals
    mbar_wait(b, parity, tag);
  };
  auto commit = [&](uint32_t b) {
    if (PAIR) umma_commit_pair(b);
    else umma_commit(b);
  };
  (void)lead;

  if (warp == 1 && lane == 0) {
    constexpr int kCtas = PAIR ? 2 : 1;      // arrivals of both CTAs land on the leader's barriers
    for (int i = 0; i < kNA; ++i) {
      mbar_init(bar(FtBars::a_full + i), kFtGatherWarps * kCtas);
      mbar_init(bar(FtBars::a_empty + i), 1);
    }
    for (int i = 0; i < kNW; ++i) {
      mbar_init(bar(FtBars::w_full + i), 1);
      mbar_init(bar(FtBars::w_empty + i), 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(bar(FtBars::acc_full + i), 1);
    for (int i = 0; i < kMaxKC; ++i) mbar_init(bar(FtBars::h1_ready + i), kFtEpiWarps * kCtas);
    for (int i = 0; i < kMaxKC / 2; ++i) mbar_init(bar(FtBars::h2_ready + i), kFtEpiWarps * kCtas);
    mbar_init(bar(FtBars::tmem_free), kFtEpiWarps * kCtas);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(smem_u32(tmem_slot));
    else tmem_alloc<512>(smem_u32(tmem_slot));
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    prefetch_tmap(&tmW3);
    prefetch_tmap(&tmOut);
  }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all();     // barriers initialised and TMEM allocated in both CTAs before any remote access
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================================================ weight-tile producer
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      auto put = [&](const CUtensorMap* tm, int kc, int nt) {
        mbar_wait(bar(FtBars::w_empty + slot), ph ^ 1, 21);
        if (PAIR) {
          // both CTAs load their 128-row half of the 256-column weight tile; the bytes of both land on the leader's barrier
          if (leader) mbar_arrive_expect_tx(bar(FtBars::w_full + slot), (uint32_t)(2 * kSlot));
          tma_load_2d_pair(sW + slot * kSlot, tm, kc * 64, nt * kNTCols + (int)rank * 128, lead(bar(FtBars::w_full + slot)));
        } else {
          mbar_arrive_expect_tx(bar(FtBars::w_full + slot), (uint32_t)kSlot);
          tma_load_2d(sW + slot * kSlot, tm, kc * 64, nt * 128, bar(FtBars::w_full + slot));
        }
        if (++slot == kNW) { slot = 0; ph ^= 1; }
      };
      for (int tile = tile0; tile < tiles_loop; tile += tstep) {
        for (int kc = 0; kc < KC1; ++kc)
          for (int nt = 0; nt < NT1; ++nt) put(&tmW1, kc, nt);
        for (int kc = 0; kc < KC2; ++kc)
          for (int nt = 0; nt < NT2; ++nt) put(&tmW2, kc, nt);
        for (int kc = 0; kc < KC3; ++kc)
          for (int nt = 0; nt < NT3; ++nt) put(&tmW3, kc, nt);
      }
    }
  } else if (warp == 1 && leader) {
    // ============================================================ MMA issuer (warp-uniform loop, one elected lane issues)
    constexpr int kMmaM = PAIR ? 256 : 128;     // pair: M = 128 rows of each CTA, N = 256 = 128 weight rows of each CTA
    constexpr uint32_t idesc = BF16 ? umma_idesc_bf16_f32(kMmaM, kNTCols) : umma_idesc_f16_f32(kMmaM, kNTCols);
    const uint64_t descH = umma_desc_kmajor_sw128(sH);
    const uint64_t descA = umma_desc_kmajor_sw128(sA);
    const uint64_t descW = umma_desc_kmajor_sw128(sW);
    int ws = 0, as = 0;
    uint32_t wph = 0, aph = 0;
    // one 64-column K chunk of one layer: A operand at `adesc`, NT weight tiles, accumulators at acc_col + nt*128
    auto chunk = [&](uint64_t adesc, int NT, uint32_t acc_col, bool first) {
      for (int nt = 0; nt < NT; ++nt) {
        mbar_wait(bar(FtBars::w_full + ws), wph, 22);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint64_t bd = descW + (uint64_t)((ws * kSlot) >> 4);
          const uint32_t d_tmem = tmem_base + acc_col + (uint32_t)(nt * kNTCols);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (PAIR) umma_pair_ss(d_tmem, adesc + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (uint32_t)(!first || k != 0));
            else umma_bf16_ss(d_tmem, adesc + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (uint32_t)(!first || k != 0));
          }
          commit(bar(FtBars::w_empty + ws));
        }
        __syncwarp();
        if (++ws == kNW) { ws = 0; wph ^= 1; }
      }
    };
    uint32_t it = 0;
    for (int tile = tile0; tile < tiles_loop; tile += tstep, ++it) {
      const uint32_t tpar = it & 1;
      // the previous tile's epilogues 2 and 3 must have drained their accumulators before GEMM1 overwrites them
      FT_TRACE(it, 0);
      wait_lead(bar(FtBars::tmem_free), tpar ^ 1, 23);
      tc_fence_after_sync();
      FT_TRACE(it, 1);
      for (int kc = 0; kc < KC1; ++kc) {
        wait_lead(bar(FtBars::a_full + as), aph, 24);
        tc_fence_after_sync();
        if (kc < 8) FT_TRACE(it, 2 + kc);
        chunk(descA + (uint64_t)((as * kSlot) >> 4), NT1, 0u, kc == 0);
        if (elect_one()) commit(bar(FtBars::a_empty + as));
        __syncwarp();
        if (++as == kNA) { as = 0; aph ^= 1; }
      }
      if (elect_one()) commit(bar(FtBars::acc_full + 0));
      __syncwarp();
      FT_TRACE(it, 10);
      // GEMM2 accumulates into columns [0, N2p): those must have been drained by epilogue 1, i.e. the H1
      // chunks that came out of them are complete
      for (int kc = 0; kc < KC3 && kc < KC2; ++kc) wait_lead(bar(FtBars::h1_ready + kc), tpar, 25);
      FT_TRACE(it, 11);
      for (int kc = 0; kc < KC2; ++kc) {
        wait_lead(bar(FtBars::h1_ready + kc), tpar, 26);
        tc_fence_after_sync();
        chunk(descH + (uint64_t)((kc * kSlot) >> 4), NT2, 0u, kc == 0);
      }
      if (elect_one()) commit(bar(FtBars::acc_full + 1));
      __syncwarp();
      FT_TRACE(it, 12);
      for (int kc = 0; kc < KC3; ++kc) {
        wait_lead(bar(FtBars::h2_ready + kc), tpar, 27);
        tc_fence_after_sync();
        chunk(descH + (uint64_t)((kc * kSlot) >> 4), NT3, 256u, kc == 0);
      }
      if (elect_one()) commit(bar(FtBars::acc_full + 2));
      __syncwarp();
      FT_TRACE(it, 13);
    }
  } else if (warp == 3) {
    // ============================================================ L2 prefetch of the NEXT tile's ids (contiguous block)
    for (int tile = tile0; tile < tiles_loop; tile += tstep) {
      const int64_t nrow0 = ((int64_t)tile + tstep) * 128;
      if (nrow0 < p.B) {
        const int64_t rows = (p.B - nrow0) < 128 ? (p.B - nrow0) : 128;
        const char* c0 = reinterpret_cast<const char*>(p.cat + nrow0 * p.F);
        const int64_t bytes = rows * p.F * 8;
        for (int64_t o = (int64_t)lane * 128; o < bytes; o += 32 * 128) prefetch_l2(c0 + o);
        if (p.num) {
          const char* n0 = reinterpret_cast<const char*>(p.num + nrow0 * p.nnum);
          const int64_t nb = rows * p.nnum * 4;
          for (int64_t o = (int64_t)lane * 128; o < nb; o += 32 * 128) prefetch_l2(n0 + o);
        }
      }
    }
  } else if (warp >= kFtGatherWarp0) {
    // ============================================================ gather: build the layer-1 A operand chunk by chunk
    // A thread owns one 4-float piece (column group) of the chunk for 8 rows: 8 ids -> 8 x 16-byte row loads.
    // The table rows are RANDOM 64-byte reads over 16.6 GB (every access misses the TLB and DRAM: ~3.5k cycles in
    // the phase trace) and the ids are a second DRAM-latency load in front of them, so the ids (and the table
    // pointer) of chunk n+1 are fetched while the row loads of chunk n are in flight: one exposed latency per
    // chunk instead of two.
    const int t = (warp - kFtGatherWarp0) * 32 + lane;   // 0..255
    const int piece = t & 15;                             // 4-float piece of the 64-column chunk
    const int rsub = t >> 4;                              // rows rsub, rsub + 16, ...
    const int FE4 = p.F * p.E4;
    const int my_tiles = tiles_loop > tile0 ? (tiles_loop - tile0 + tstep - 1) / tstep : 0;
    const int items = my_tiles * KC1;                     // (tile, chunk) pairs this CTA gathers, in order
    int slot = 0;
    uint32_t ph = 0;
    bool bad = false;
    float amax = 0.f;
    // what a thread needs to know to issue the row loads of item n (tile n / KC1, chunk n % KC1)
    struct Ids {
      int64_t idx[8];      // table row per owned sample row; -1: row beyond the batch (or not an embedding piece)
      const float4* tab;   // table base + this thread's 16-byte part
      int64_t card;
    };
    auto load_ids = [&](int n, Ids& o) {
      const int w4 = (n % KC1) * 16 + piece;
      const bool emb = n < items && w4 < FE4;
      const int f = emb ? w4 / p.E4 : 0;
      o.tab = reinterpret_cast<const float4*>(p.tables[f]) + (emb ? w4 - f * p.E4 : 0);
      o.card = __ldg(p.cards + f);
      const int64_t row0 = ((int64_t)tile0 + (int64_t)(n / KC1) * tstep) * 128;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t grow = row0 + i * 16 + rsub;
        o.idx[i] = (emb && grow < p.B) ? __ldg(p.cat + grow * p.F + f) : -1;
      }
    };
    Ids cur;
    load_ids(0, cur);
    for (int n = 0; n < items; ++n) {
      const int kc = n % KC1;
      const uint32_t git = (uint32_t)(n / KC1);
      const int64_t row0 = ((int64_t)tile0 + (int64_t)git * tstep) * 128;
      if (warp == kFtGatherWarp0 && kc < 8) FT_TRACE(git, 16 + kc);
      const int w4 = kc * 16 + piece;
      float4 v[8];
      if (w4 < FE4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t grow = row0 + i * 16 + rsub;
          int64_t r = cur.idx[i];
          if (grow < p.B) {
            if (r < 0 || r >= cur.card) { bad = true; r = 0; }
            v[i] = ldg_nc_f4(cur.tab + r * p.E4);
          } else {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      } else {
        const int c0 = (w4 - FE4) * 4;   // first numerical column of this piece
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t grow = row0 + i * 16 + rsub;
          float e[4] = {0.f, 0.f, 0.f, 0.f};
          if (grow < p.B && p.num) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c0 + j < p.nnum) e[j] = __ldg(p.num + grow * p.nnum + c0 + j);
          }
          v[i] = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      load_ids(n + 1, cur);                               // next chunk's ids ride behind this chunk's row loads
      mbar_wait(bar(FtBars::a_empty + slot), ph ^ 1, 28);
      const uint32_t dst = sA + (uint32_t)(slot * kSlot);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 16 + rsub;
        if (!BF16) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
        st_shared_v2(dst + sw128_off(r, piece >> 1) + (uint32_t)((piece & 1) * 8), pack2<BF16>(v[i].x, v[i].y),
                     pack2<BF16>(v[i].z, v[i].w));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive_lead(bar(FtBars::a_full + slot));
      if (warp == kFtGatherWarp0 && kc < 8) FT_TRACE(git, 24 + kc);
      if (++slot == kNA) { slot = 0; ph ^= 1; }
    }
    int flags = bad ? kTowerErrIndex : 0;
    if (!BF16 && !(amax <= 65504.f)) flags |= kTowerErrSaturate;   // also catches +-inf inputs
    if (flags && p.err_flag) atomicOr(p.err_flag, flags);
  } else if (warp >= kFtEpiWarp0) {
    // ============================================================ epilogues
    const int e = warp - kFtEpiWarp0;
    const int quarter = e & 3;       // TMEM lane quarter (== warp % 4)
    const int half = e >> 2;         // which 32-column half of every 64-column chunk
    const int row = quarter * 32 + lane;                      // row of the tile == TMEM lane
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t stage = sH + 65536u + (uint32_t)e * 8192u;  // 2 x 4 KB output staging (H chunks 4..7)
    float* ssx = reinterpret_cast<float*>(smem);               // [8][32] partial sums of squares (H chunk 0, free in epilogue 3)
    uint32_t hmax = 0u;      // running max of the packed fp16 activations (>= 0 after ReLU): 0x7BFF = saturated
    uint32_t it = 0;
    for (int tile = tile0; tile < tiles_loop; tile += tstep, ++it) {
      const uint32_t tpar = it & 1;
      uint32_t r[32];
      // ---- hidden layers: TMEM -> bias -> ReLU -> 16-bit -> swizzled K-major chunk of the next layer's A operand
      for (int layer = 0; layer < 2; ++layer) {
        const int KC = layer == 0 ? KC2 : KC3;
        const int boff = layer == 0 ? 0 : p.N1p;
        mbar_wait(bar(FtBars::acc_full + layer), tpar, 29);
        tc_fence_after_sync();
        if (e == 0) FT_TRACE(it, 32 + layer * 2);
        // one chunk = TMEM -> +bias -> ReLU -> 16-bit -> swizzled smem.  The TMEM read of chunk kc+1 is issued
        // before chunk kc is converted (two register buffers), so the load latency hides behind the math.
        auto emit = [&](const uint32_t (&rr)[32], int kc) {
          const int c = kc * 2 + half;     // 32-column chunk
          uint32_t pk[16];
          const float2* b2 = reinterpret_cast<const float2*>(bias.v + boff + c * 32);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 bb = b2[i];
            pk[i] = pack2_relu<BF16>(__uint_as_float(rr[2 * i]) + bb.x, __uint_as_float(rr[2 * i + 1]) + bb.y);
            if (!BF16) hmax = hmax2_u32(hmax, pk[i]);
          }
          const uint32_t dst = sH + (uint32_t)(kc * kSlot);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(dst + sw128_off(row, half * 4 + j), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        };
        // publish chunks [k0, k1): ONE proxy fence for the group (the fence, not the math, was the cost of a chunk:
        // ~1.6 k cycles each in the phase trace), then one arrive per chunk barrier
        auto publish = [&](int k0, int k1) {
          fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core's async-proxy reads
          tc_fence_before_sync();       // ... and this warp's TMEM reads are done before the MMAs overwrite the columns
          __syncwarp();
          if (lane == 0)
            for (int kc = k0; kc < k1; ++kc) arrive_lead(bar((layer == 0 ? FtBars::h1_ready : FtBars::h2_ready) + kc));
        };
        // Layer 1: GEMM2 re-uses the TMEM columns of H1's first N2p columns, so it cannot start before chunks
        // [0, KC3) are ALL drained: they are published as one group.  Later chunks go out in pairs.
        const int first_group = layer == 0 ? ((KC3 < KC ? KC3 : KC) + 1) & ~1 : 2;
        uint32_t r2[32];
        tmem_ld_32x32(tlane + (uint32_t)(half * 32), r);
        int pub = 0;
#pragma unroll 1
        for (int kc = 0; kc < KC; kc += 2) {     // KC is even (widths are padded to 128 columns)
          tmem_ld_wait_dep(r);
          tmem_ld_32x32(tlane + (uint32_t)(((kc + 1) * 2 + half) * 32), r2);
          emit(r, kc);
          tmem_ld_wait_dep(r2);
          if (kc + 2 < KC) tmem_ld_32x32(tlane + (uint32_t)(((kc + 2) * 2 + half) * 32), r);
          emit(r2, kc + 1);
          if (kc + 2 >= first_group || kc + 2 >= KC) {
            publish(pub, kc + 2);
            pub = kc + 2;
          }
        }
        if (e == 0) FT_TRACE(it, 33 + layer * 2);
      }
      // ---- output layer: bias, L2 norm over the whole row (two warps share a row quarter), fp32, TMA store
      mbar_wait(bar(FtBars::acc_full + 2), tpar, 30);
      tc_fence_after_sync();
      if (e == 0) FT_TRACE(it, 36);
      const int boff3 = p.N1p + p.N2p;
      float ss = 0.f;
#pragma unroll 1
      for (int kc = 0; kc < (p.N3p >> 6); ++kc) {
        const int c = kc * 2 + half;
        tmem_ld_32x32(tlane + 256u + (uint32_t)(c * 32), r);
        tmem_ld_wait_dep(r);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = __uint_as_float(r[i]) + bias.v[boff3 + c * 32 + i];
          ss = fmaf(v, v, ss);
        }
      }
      if (e == 0) FT_TRACE(it, 37);
      ssx[e * 32 + lane] = ss;
      asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");      // the two warps of this lane quarter
      const float other = ssx[(e ^ 4) * 32 + lane];
      const float tot = half == 0 ? ss + other : other + ss;                // same operand order in both warps
      const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);                   // F.normalize: x / max(||x||, eps)
      int buf = 0;
#pragma unroll 1
      for (int kc = 0; kc < (p.N3p >> 6); ++kc) {
        const int c = kc * 2 + half;
        tmem_ld_32x32(tlane + 256u + (uint32_t)(c * 32), r);
        tmem_ld_wait_dep(r);
        if (kc + 1 == (p.N3p >> 6)) {     // last TMEM read of this tile by this warp: release the accumulators
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) arrive_lead(bar(FtBars::tmem_free));
        }
        if (lane == 0) bulk_wait_read<1>();       // the store that last used this staging buffer has read it
        __syncwarp();
        const uint32_t sb = stage + (uint32_t)buf * 4096u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = 4 * j;
          st_shared_v4(sb + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)),
                       __float_as_uint((__uint_as_float(r[i]) + bias.v[boff3 + c * 32 + i]) * inv),
                       __float_as_uint((__uint_as_float(r[i + 1]) + bias.v[boff3 + c * 32 + i + 1]) * inv),
                       __float_as_uint((__uint_as_float(r[i + 2]) + bias.v[boff3 + c * 32 + i + 2]) * inv),
                       __float_as_uint((__uint_as_float(r[i + 3]) + bias.v[boff3 + c * 32 + i + 3]) * inv));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOut, sb, c * 32, tile * 128 + quarter * 32);   // rows / columns past the tensor are clipped
          bulk_commit();
        }
        buf ^= 1;
      }
      // the staging buffers and ssx live in the H region: every epilogue warp must be done with them (stores
      // have READ shared memory) before any warp starts writing the next tile's H1 there
      if (e == 0) FT_TRACE(it, 38);
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (e == 0) FT_TRACE(it, 39);
    }
    // a hidden activation that hit the fp16 ceiling (cvt.satfinite clamps to 65504 = 0x7BFF; NaN stays NaN = 0x7FFF)
    if (!BF16 && ((hmax & 0xFFFFu) >= 0x7BFFu || (hmax >> 16) >= 0x7BFFu) && p.err_flag)
      atomicOr(p.err_flag, kTowerErrSaturate);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // global writes complete before exit
  }
  __syncwarp();
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all();   // neither CTA may leave (or free TMEM) while the pair's MMAs / remote arrives are in flight
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    if (PAIR) tmem_dealloc_pair(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !sym)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

}  // namespace

int make_tmap_f32_out(CUtensorMap* out, const void* base, int64_t rows, int64_t cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(B2R_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2R_ECUDA, "cuTensorMapEncodeTiled (output) failed: " + std::to_string((int)r));
  return B2R_OK;
}

int launch_tower_fused(b2r_tower* t, const int64_t* cat, const float* num, int64_t B, float* out,
                       int32_t* err_flag, cudaStream_t stream) {
  if (!t->fused_ok) return fail(B2R_EUNSUPPORTED, "tower_forward: shape does not fit the fused kernel");
  CUtensorMap tmOut;
  int rc = make_tmap_f32_out(&tmOut, out, B, t->n[2]);
  if (rc) return rc;
  FtParams p;
  p.tables = t->tables;
  p.cards = t->cards;
  p.cat = cat;
  p.num = num;
  p.B = B;
  p.F = t->F;
  p.E4 = t->E / 4;
  p.nnum = t->nnum;
  p.KC1 = t->K1p / 64;
  p.N1p = t->np[0];
  p.N2p = t->np[1];
  p.N3p = t->np[2];
  p.err_flag = err_flag;
  p.trace = reinterpret_cast<long long*>(t->trace_ptr);
  int dev = 0;
  B2R_CUDA(cudaGetDevice(&dev));
  const int bf = t->bf16 ? 1 : 0;
  const int64_t tiles = (B + 127) / 128;
  const FtBias& bias = *reinterpret_cast<const FtBias*>(t->bias_host);
  // CTA pairs (cta_group::2): weight tiles of 256 columns -> every padded width must be a multiple of 256
  const bool pair = t->pair && t->num_sms >= 2 && tiles >= 2 && (p.N1p % 256) == 0 && (p.N2p % 256) == 0 && (p.N3p % 256) == 0;
  static bool configured[2][2][64] = {};
  if (!configured[pair][bf][dev & 63]) {
    const void* fn = pair ? (bf ? (const void*)tower_fused_kernel<true, true> : (const void*)tower_fused_kernel<false, true>)
                          : (bf ? (const void*)tower_fused_kernel<true, false> : (const void*)tower_fused_kernel<false, false>);
    B2R_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtSmem));
    configured[pair][bf][dev & 63] = true;
  }
  if (pair) {
    const int64_t pairs = (tiles + 1) / 2;
    const int clusters = (int)(pairs < t->num_sms / 2 ? pairs : t->num_sms / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kFtThreads);
    cfg.dynamicSmemBytes = kFtSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (bf) B2R_CUDA(cudaLaunchKernelEx(&cfg, tower_fused_kernel<true, true>, t->tmWfb[0], t->tmWfb[1], t->tmWfb[2], tmOut, bias, p));
    else B2R_CUDA(cudaLaunchKernelEx(&cfg, tower_fused_kernel<false, true>, t->tmWf[0], t->tmWf[1], t->tmWf[2], tmOut, bias, p));
  } else {
    const int grid = (int)(tiles < t->num_sms ? tiles : t->num_sms);
    if (bf) tower_fused_kernel<true, false><<<grid, kFtThreads, kFtSmem, stream>>>(t->tmWfb[0], t->tmWfb[1], t->tmWfb[2], tmOut, bias, p);
    else tower_fused_kernel<false, false><<<grid, kFtThreads, kFtSmem, stream>>>(t->tmWf[0], t->tmWf[1], t->tmWf[2], tmOut, bias, p);
  }
  B2R_CHECK_LAUNCH("tower_fused_kernel");
  return B2R_OK;
}

}  // namespace b2r
