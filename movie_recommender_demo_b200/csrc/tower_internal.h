// tower_internal.h — the tower handle shared by tower_mlp.cu (handle + layer-by-layer path) and
// tower_fused.cu (the single fused kernel).
#pragma once
#include <cuda_fp16.h>

#include "internal.h"

// hidden_dims of any length (two_tower_model.py:83-95 loops over the list): up to 15 hidden layers + the output layer
constexpr int kTowerMaxLayers = 16;

struct b2r_tower {
  int device = 0, num_sms = 148;
  int F = 0, E = 0, nnum = 0;
  int K1 = 0, K1p = 0;           // layer-1 fan-in and its 64-padding
  int L = 3;                     // Linear layers (hidden layers + the output layer); the fused kernel serves L == 3
  int n[kTowerMaxLayers] = {};   // true fan-outs
  int np[kTowerMaxLayers] = {};  // padded to a multiple of 128
  // operand format of activations and weights: 0 = IEEE fp16 (default: 8x smaller rounding error),
  // 1 = bf16 (fp32 range; chosen when a folded weight, an input or a hidden activation would exceed
  // the fp16 range).  Both weight copies are kept so the switch costs nothing at run time.
  int bf16 = 0;
  int force_path = 0;            // 0 auto, 1 layer-by-layer kernels, 2 fused kernel (tests)
  int pair = 1;                  // fused kernel on CTA pairs (tcgen05 cta_group::2, 256-column weight tiles) when every padded
                                 // width is a multiple of 256: 0.148 -> 0.139 ms at B = 65536; set_param("pair", 0) = one CTA per tile
  uintptr_t trace_ptr = 0;       // debug: device buffer for the fused kernel's phase timeline (tests/prof_tower.py)
  __half* w[kTowerMaxLayers] = {};          // [np[l], Kp[l]] fp16, zero padded
  __nv_bfloat16* wb[kTowerMaxLayers] = {};  // the same weights in bf16
  float* b[kTowerMaxLayers] = {};           // [np[l]]
  float bias_host[1024];         // b1 | b2 | b3 (padded), passed to the fused kernel by value
  const float** tables = nullptr;  // device array [F]
  int64_t* cards = nullptr;        // device array [F]
  CUtensorMap tmW[kTowerMaxLayers], tmWb[kTowerMaxLayers];   // layer-by-layer path: box rows 256 / 128
  CUtensorMap tmWf[3], tmWfb[3];   // fused path: box {64, 128}
  bool fused_ok = false;           // shape fits the fused kernel (two hidden layers, widths <= 512/256/256)
};

namespace b2r {
// box = {64 columns, box_rows rows}, 128-byte swizzle, 16-bit elements (fp16 or bf16: same layout)
int make_tmap_f16_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows);
// out = epilogue(act16 [M, K] · Wᵀ + bias) on the tcgen05 GEMM of tower_mlp.cu.  act16: 16-bit row-major, K a
// multiple of 64; tmW: tensor map of the 16-bit weights [N, K] with box rows 256 when N % 256 == 0 else 128;
// mode 0: ReLU -> 16-bit [M, ldo]; 1: row L2-normalise -> fp32 (N <= 256); 2: fp32 [M, ldo] (first n_store columns)
int launch_linear(const void* act16, const CUtensorMap& tmW, int64_t M, int N, int K, const float* bias, void* out,
                  int64_t ldo, int n_store, int mode, int32_t* err_flag, int num_sms, cudaStream_t stream, bool bf16);
// fp32 [rows, cols] row-major, box = {32 columns (128 B), 32 rows}, 128-byte swizzle (TMA stores of the output)
int make_tmap_f32_out(CUtensorMap* out, const void* base, int64_t rows, int64_t cols);
// status word bits written by the tower kernels
constexpr int kTowerErrIndex = 1;     // categorical id out of range (torch: IndexError)
constexpr int kTowerErrSaturate = 2;  // an fp16 operand saturated (|v| > 65504): result clipped, rerun in bf16
int launch_tower_fused(b2r_tower* t, const int64_t* cat, const float* num, int64_t B, float* out,
                       int32_t* err_flag, cudaStream_t stream);
}  // namespace b2r
