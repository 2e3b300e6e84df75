// internal.h — shared declarations between the .cu translation units of libb2retr.so
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/b2retr.h"

namespace b2r {

// ---------------------------------------------------------------- errors ---
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define B2R_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::b2r::fail(B2R_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
  } while (0)
#define B2R_CHECK_LAUNCH(name)                                                              \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      return ::b2r::fail(B2R_ECUDA, std::string("launch ") + name + ": " +                  \
                                        cudaGetErrorString(_e));                            \
    ::b2r::count_launch();                                                                  \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
// Width of the rescore window below the k-th bf16 score: 2E, E = eps * |q| * max|x| bounds
// |bf16 score - exact score|.  One definition so every kernel computes the identical float.
__host__ __device__ inline float rescore_margin(float eps, float qnorm, float maxnorm) {
  // + 4e-6 * qnorm: absolute slack for fp16 subnormal flushing (|x_i| < 2^-14) and fp32 accumulation
  return 2.0f * (eps * qnorm * maxnorm * 1.0001f + 4e-6f * qnorm);
}
// 16-bit scan formats (same tcgen05 kind::f16 rate): bf16 = range of fp32, 8-bit significand;
// fp16 = 11-bit significand (8x smaller rounding error) but |x| <= 65504 — used when every stored
// row was L2-normalised on ingest (|x_i| <= 1); the query's scan copy is always unit-norm.
enum ScanDtype { SCAN_BF16 = 0, SCAN_FP16 = 1 };
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------ scan kernel ---
constexpr int kTileRows = 128;   // corpus rows per UMMA tile (N dimension)
constexpr int kQBlock = 128;     // queries per UMMA M block (TMEM lanes)
constexpr int kKChunk = 64;      // bf16 elements per 128-byte swizzle span
constexpr int kGroupCols = 32;   // GMAX group = one tcgen05.ld chunk of 32 corpus rows

enum ScanMode { SCAN_DUMP = 0, SCAN_GMAX = 1, SCAN_FILTER = 2 };

struct ScanParams {
  int Q;            // valid queries
  int QG;           // query groups of 128*MQ rows
  int64_t N;        // valid corpus rows (rows >= N are zero-filled by TMA and masked)
  int d;            // embedding dim (multiple of 64, <= 256)
  int tile_first;   // scanned tile j maps to corpus tile tile_first + j*tile_stride
  int tile_stride;
  int tile_count;
  int splits;       // unit u -> (split = u / QG, qg = u % QG)
  uint32_t idesc;   // tcgen05 instruction descriptor (operand format bf16 or fp16, M=128, N=128)
  // FILTER: candidates go to private segments, one per (query, corpus split[, column half]):
  // exactly one warp ever writes a segment, so an append is a plain store + register counter.
  const float* tau;      // [Qpad] candidate threshold (score >= tau passes); +inf for padding
  int* cand_count;       // [Qpad, nseg] entries produced per segment (may exceed cap_seg)
  uint2* cand;           // [Qpad, nseg, cap_seg] (score bits, local row index)
  int nseg;              // splits * (MQ == 1 ? 2 : 1)
  int cap_seg;
  int early_release;     // FILTER, 2 chunks per warp: free the accumulator buffer before the scores are examined
  // GMAX
  float* gmax;           // [Qpad, gstride]; entry (q, j*4 + chunk)
  int gstride;
  // DUMP
  float* dump;           // [Q, ld]
  int64_t ld;
};

// Builds the 2D bf16 tensor map (rows x d, box 128 rows x 64 cols, 128B swizzle).
int make_tmap_bf16_rows(CUtensorMap* out, const void* base, int64_t rows, int d);
// epi_warps: 8, or 16 for the (MQ = 2, FILTER) variant with one candidate segment per 64-column half
// (p.nseg must then be 2 * splits, as for MQ = 1)
int launch_scan(int mode, int MQ, const CUtensorMap& tmQ, const CUtensorMap& tmX,
                const ScanParams& p, int num_sms, cudaStream_t stream, int epi_warps = 8, int walk = 0);
// picks (MQ, splits) for a (Q, tiles) problem
void plan_scan(int Q, int tile_count, int num_sms, int* MQ, int* QG, int* splits);
// FILTER scan on CTA pairs (scan_pair.cu, tcgen05 cta_group::2): p.QG = groups of 256 queries, p.tile_count =
// 256-row pair tiles, p.nseg = 4 * p.splits, p.idesc = scan_idesc_pair()
int launch_scan_pair(const CUtensorMap& tmQ, const CUtensorMap& tmX, const ScanParams& p, int num_sms,
                     cudaStream_t stream, int walk, int mode = SCAN_FILTER);
uint32_t scan_idesc_pair(int fp16);

// --------------------------------------------------------- ingest / queries ---
// rows fp32 [n,d] -> (optional L2 normalise) -> fp32 master + bf16 copy; updates *maxnorm
// (device float, atomic max of the stored rows' L2 norms).
int launch_ingest(const float* x, int64_t n, int d, int normalize, float* out32,
                  __nv_bfloat16* out16, int fp16, float* maxnorm, cudaStream_t stream);
// queries fp32 [q,d] -> q32 [qpad,d] (normalised iff `normalize`), q16 [qpad,d] = ALWAYS unit-norm
// 16-bit scan copy (zero padded), qnorm [qpad] = norm of the scan copy (1, or 0 for a zero query);
// tau_init (optional) [qpad] is set to +inf in the same launch (the per-query threshold's start value)
int launch_prep_queries(const float* x, int q, int qpad, int d, int normalize, float* q32,
                        __nv_bfloat16* q16, int fp16, float* qnorm, cudaStream_t stream,
                        float* tau_init = nullptr);
// re-encode the 16-bit scan copy of n rows from the fp32 master
int launch_reencode(const float* x32, int64_t n, int d, __nv_bfloat16* out16, int fp16, cudaStream_t stream);
uint32_t scan_idesc(int fp16);

// ------------------------------------------------------------------ select ---
// per-row m-th largest of vals[r, 0..T) (ld stride) -> tau[r]; if cand_* given, also
// appends every (val >= tau[r], column) of row r to the candidate buffers (dense path).
// With qnorm/maxnorm given (dense path, m = k) the threshold is lowered by the rescore margin:
// tau = kth - 2E, i.e. exactly the provable rescore window.
int launch_kth_value(const float* vals, int rows, int64_t T, int64_t ld, int m, float* tau,
                     int* cand_count, uint2* cand, int cap, const float* qnorm, const float* maxnorm,
                     float eps, cudaStream_t stream);
struct SelectParams {
  int Q, k, d;
  int nseg, cap_seg;     // candidate segments per query / slots per segment
  int key_cap, rescore_max;  // set by the launcher: shared-memory key / window capacities
  int64_t N;
  const int* cand_count; // [Q, nseg]
  const uint2* cand;     // [Q, nseg, cap_seg]
  const float* tau;      // [Q] threshold used to build the candidates
  const float* q32;      // [Qpad, d] prepared fp32 queries
  const float* qnorm;    // [Qpad]
  const float* x32;      // [N, d] fp32 master rows
  const float* maxnorm;  // device scalar
  float eps;             // relative score error bound
  int rescore;           // 1: fp32 rescore
  const uint32_t* perm;  // optional stored row -> label (IVF keeps rows sorted by list)
  const int* scanned;    // optional [Q]: rows actually scanned per query (IVF); results that exist = min(k, scanned)
  const int64_t* ids;    // optional id map indexed by label
  int negate_out;        // 1: scores are negated distances (L2 paths): D = -score, empty slots +FLT_MAX
  int64_t label_base;
  float* D;              // [Q,k]
  int64_t* I;            // [Q,k]
  int32_t* status;       // [Q] or null
  float* tau_retry;      // [Q] or null
};
int launch_select_rescore(const SelectParams& p, cudaStream_t stream);
int launch_fill_f32(float* p, int64_t n, float v, cudaStream_t stream);
int launch_fill_i32(int* p, int64_t n, int v, cudaStream_t stream);

}  // namespace b2r

struct b2r_index {
  int kind = 0, d = 0, nlist = 0, pq_m = 0, pq_bits = 0, metric = 0, device = 0;
  int num_sms = 148;
  int64_t ntotal = 0, capacity = 0;
  float* x32 = nullptr;          // [capacity, d] fp32 master rows
  __nv_bfloat16* x16 = nullptr;  // [capacity, d] bf16 scan rows
  float* maxnorm = nullptr;      // device scalar: max stored row norm
  int64_t* ids = nullptr;        // optional id map [n_ids]
  int64_t n_ids = 0;
  int64_t label_base = 0;
  CUtensorMap tmX;
  bool trained = true;
  // tunables
  double eps = 0.00390625 * 1.02;  // bf16: 2^-8 (two roundings of 2^-9 per product), 2% slack
  double eps_fp16 = 0.0009765625 * 1.02;  // fp16: 2^-10
  int scan_dtype_req = -1;  // -1 auto (fp16 while every add was normalised), 0 bf16, 1 fp16
  int scan_fp16 = -1;       // current format of x16 (-1: nothing stored yet)
  double cand_factor_fp16 = 2.5;
  int walk = 1;                  // FILTER hit walk: 1 = only the passing 3-element sub-groups (2-5 % faster), 0 = all 8
  int pair_gmax = 1;             // the threshold sampling pass of batches > 128 on CTA pairs as well (0: one-CTA GMAX scan)
  int early_release = 0;         // filter epilogue: release the TMEM buffer before the scores are examined (A/B: no gain)
  int pair_scan = 1;             // FILTER scan of batches > 128 on CTA pairs (tcgen05 cta_group::2, scan_pair.cu): 17-21 %
                                 // faster than the one-CTA kernel (1.58 -> 1.31 ms at Q=4096); 0 selects the one-CTA kernel
  int epi_warps = 16;            // epilogue warps of the MQ = 2 filter scan (8 | 16); 16 measured 8-9 % faster
  double cand_factor = 4.0;
  int cand_cap = 4096;
  int rescore = 1;
  int force_path = 0;  // 0 auto, 1 dense, 2 filter (tests)
  int pq_scan_path = 0;  // 0 auto (query-major when pq_m allows), 1 force the (query, list)-pair kernel (tests)
  int ivf_debug = 0;   // profiling experiments (never set in production)
  int ivf_sample_rows = 128;  // fused IVF path: rows of every probed list the sample pass scores (<= 128 = one tile)
  int ivf_fused = 1;   // IVF-Flat: threshold filter fused into the list scan (sample pass + filter scan); 0 = dump all pair scores
  int ivf_sample = 1;  // IVF candidate threshold from a score sample (0: always the exact radix passes)
  int64_t dense_budget = (int64_t)1 << 30;  // bytes of dumped scores per query chunk
  // optional CUDA-event timing of the dominant (filter scan) kernel, for bench.py's roofline
  int profile = 0;
  std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
  size_t prof_used = 0;
  // ---- IVF state (kind != FLAT): rows are stored SORTED BY LIST (x32/x16 above), so every
  // inverted list is a contiguous row range [list_off[l], list_off[l+1]) the scan can tile.
  b2r_index* quantizer = nullptr;   // flat IP index over the nlist centroids (faiss: IndexFlatIP quantizer)
  int64_t* list_off = nullptr;      // device [nlist + 1]
  int32_t* row_list = nullptr;      // device [capacity] list id of each stored (sorted) row
  uint32_t* perm = nullptr;         // device [capacity] sorted row -> insertion label
  std::vector<int64_t> list_sizes_host;  // mirrors list sizes (workspace bounds)
  mutable std::vector<int64_t> list_sizes_desc;  // the same, sorted descending: cached by the search planner
  mutable int64_t list_sizes_desc_for = -1;      // ntotal the cache was built for (-1: stale)
  // ---- PQ state (kind == IVF_PQ)
  float* codebooks = nullptr;       // device [pq_m, 256, d/pq_m]
  uint8_t* codes = nullptr;         // device [capacity, pq_m] (sorted by list)
  float* pq_list_tab = nullptr;     // device [nlist, pq_m, 256]: 2 c_ls.y_sj + |c_ls|^2 (faiss "precomputed table")
  float* pq_row_term = nullptr;     // device [pq_row_term_cap]: sum_s pq_list_tab[list(row)][s][code_s(row)] per stored row
  int64_t pq_row_term_cap = 0;
  bool pq_trained = false;
};


namespace b2r {
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};


// flat search entry shared by the C ABI and the IVF coarse quantiser (index.cu)
size_t flat_search_workspace(const b2r_index* h, int q, int k);
int flat_search(b2r_index* h, int q, const float* queries, int normalize, int k, float* D, int64_t* I,
                int32_t* status, float* tau_retry, const float* tau_in, void* workspace, size_t ws_bytes,
                cudaStream_t stream);
int flat_add(b2r_index* h, int64_t n, const float* x, int normalize, cudaStream_t stream);
int flat_create(b2r_index** out, int d, int device);

// IVF family (ivf.cu)
int ivf_train(b2r_index* h, int64_t n, const float* x, uint64_t seed, cudaStream_t stream);
int ivf_add(b2r_index* h, int64_t n, const float* x, int normalize, cudaStream_t stream);
size_t ivf_search_workspace(const b2r_index* h, int q, int k, int nprobe);
int ivf_search(b2r_index* h, int q, const float* queries, int normalize, int k, int nprobe, float* D,
               int64_t* I, int32_t* status, void* workspace, size_t ws_bytes, cudaStream_t stream);
void ivf_free(b2r_index* h);
// product quantiser (ivfpq.cu)
int pq_residuals(b2r_index* h, int64_t n, const float* x, const int64_t* assign, float* r, cudaStream_t stream);
int pq_train(b2r_index* h, int64_t n, const float* resid, uint64_t seed, cudaStream_t stream);
int pq_encode(b2r_index* h, int64_t n, const float* resid, uint8_t* codes, cudaStream_t stream);
int pq_scatter_codes(const uint8_t* src, const int64_t* dst, int64_t n, int m, const int32_t* list_src,
                     const int64_t* assign_src, const uint32_t* perm_src, uint32_t label0, uint8_t* ocodes,
                     int32_t* olist, uint32_t* operm, cudaStream_t stream);
int pq_scan(b2r_index* h, int nq, int npairs, const float* q32, const int64_t* coarse, int nprobe,
            const int64_t* pair_out, float* qtab, float* scorebuf, cudaStream_t stream);
int pq_build_list_tables(b2r_index* h, cudaStream_t stream);
int pq_update_row_terms(b2r_index* h, cudaStream_t stream);   // after the code storage or the quantisers changed
}  // namespace b2r
