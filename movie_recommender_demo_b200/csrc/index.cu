// index.cu — opaque index handle + search orchestration behind the C ABI (include/b2retr.h).
//
// Mirrors the faiss objects the reference builds in FAISSIndex._create_index
// (faiss_retrieval.py:44-81) and drives through .train/.add/.search
// (faiss_retrieval.py:93, :118, :155).
//
// Flat search pipeline (all on the caller's stream, no allocation, no host sync):
//   prep_queries -> [GMAX sample scan -> kth_value => tau]  (large corpus)
//                   [DUMP scan -> kth_value + compaction]    (small corpus, "dense")
//                -> FILTER scan (score >= tau -> candidate buffers)
//                -> select_rescore (sort, provable rescore window, exact fp32, id map)
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <algorithm>

#include <mutex>
#include <vector>

#include "internal.h"

namespace b2r {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { t_error = msg; }
int fail(int code, const std::string& msg) {
  t_error = msg;
  return code;
}

}  // namespace b2r

using namespace b2r;

namespace {

// ---- search plan: path + workspace layout for one query chunk ----------------
struct Plan {
  bool dense = false;
  int chunk = 0;        // queries per chunk
  int MQ = 1, QG = 1, splits = 1, qpad = 0;
  int tiles = 0;        // corpus tiles
  int nseg = 1;         // candidate segments per query
  int cap_seg = 0;      // slots per segment
  int c_target = 0;
  // sampling pass
  int s_tiles = 0, s_stride = 1, s_splits = 1, m_rank = 1;
  int64_t gstride = 0;
  int64_t dump_ld = 0;
  // workspace offsets
  size_t off_q16 = 0, off_q32 = 0, off_qnorm = 0, off_tau = 0, off_count = 0, off_cscore = 0,
         off_cidx = 0, off_aux = 0, total = 0;
};

// Launch geometry of the filter scan for a chunk of q queries over `tiles` 128-row corpus tiles.
// pair: CTA-pair kernel (scan_pair.cu) — query groups of 256, 256-row pair tiles, units spread over num_sms / 2
// clusters, 4 candidate segments per split; else the one-CTA kernel of scan_tc.cu (MQ, 1 or 2 segments per split).
struct ScanGeom {
  int MQ = 1, QG = 1, splits = 1, seg_per_split = 1, qpad = 0;
  bool pair = false;
};
ScanGeom scan_geometry(const b2r_index* h, int q, int tiles, bool filter) {
  ScanGeom g;
  plan_scan(q, tiles, h->num_sms, &g.MQ, &g.QG, &g.splits);
  g.qpad = g.QG * g.MQ * kQBlock;
  g.pair = filter && h->pair_scan && g.MQ == 2 && h->num_sms >= 2 && tiles >= 2;
  if (g.pair) {
    int mq, qg;
    plan_scan(q, (int)ceil_div(tiles, 2), h->num_sms / 2, &mq, &qg, &g.splits);   // qg == g.QG: groups of 256 queries
    g.seg_per_split = 4;
  } else {
    g.seg_per_split = (g.MQ == 1 || h->epi_warps == 16) ? 2 : 1;
  }
  return g;
}

Plan make_plan(const b2r_index* h, int q, int k, bool have_tau) {
  Plan pl;
  const int64_t N = h->ntotal;
  pl.tiles = (int)ceil_div(N > 0 ? N : 1, kTileRows);
  const int cap = h->cand_cap;  // candidates per query the select kernel holds (4096)
  int ct = (int)llround((h->scan_fp16 == 1 ? h->cand_factor_fp16 : h->cand_factor) * k);
  if (ct < 64) ct = 64;
  const int ct_max = (cap * 5) / 8;  // leave head-room for sampling noise
  if (ct > ct_max) ct = ct_max;
  if (ct < k) ct = k;
  pl.c_target = ct;
  // dense path (dump all scores, exact k-th) only when the corpus is so small that the sampled group
  // maxima cannot resolve the threshold rank: each 32-row group would hold about one candidate or more.
  // Above that the filter path wins even at 100k rows (its threshold costs two short launches, the
  // dense path's per-query radix sweeps over N dumped scores do not parallelise at small batch).
  const bool small = N < (int64_t)32 * ct;
  pl.dense = !have_tau && (h->force_path == 1 || (h->force_path == 0 && small));
  if (h->force_path == 2) pl.dense = false;
  int chunk = q;
  if (pl.dense) {
    pl.dump_ld = (int64_t)align_up((size_t)(N > 0 ? N : 1), 4);
    int64_t qc = h->dense_budget / (pl.dump_ld * 4);
    if (qc < 1) qc = 1;
    if (qc >= 256) qc = qc / 256 * 256;
    if (qc < chunk) chunk = (int)qc;
  } else {
    if (chunk > 8192) chunk = 8192;
  }
  pl.chunk = chunk;
  const ScanGeom geo = scan_geometry(h, chunk, pl.tiles, !pl.dense);
  pl.MQ = geo.MQ;
  pl.QG = geo.QG;
  pl.splits = geo.splits;
  pl.qpad = geo.qpad;
  if (pl.dense) {
    pl.nseg = 1;
    pl.cap_seg = cap;
  } else {
    // filter path: one private segment per (query, corpus split, column part); sized for the worst chunk
    // geometry (a smaller last chunk may pick another kernel with fewer segments per split: each of its
    // segments then holds more entries, so the slot count is sized for TWO segments per split)
    const int sps = geo.seg_per_split < 2 ? 2 : geo.seg_per_split;
    pl.nseg = pl.splits * sps;
    int per = (int)ceil_div(4 * (int64_t)ct, pl.splits);  // 8x the mean entries of a half-split segment
    if (sps == 4) per = (int)ceil_div(3 * (int64_t)ct, pl.splits);  // 12x the mean of a quarter-split segment
    int cs = 32;
    while (cs < per) cs <<= 1;
    if (cs > cap) cs = cap;
    pl.cap_seg = cs;
  }
  if (!pl.dense && !have_tau) {
    int64_t st = ceil_div(N, 2 * (int64_t)ct);
    if (st < 256) st = 256;
    if (st > pl.tiles) st = pl.tiles;
    pl.s_stride = (int)(pl.tiles / st);
    if (pl.s_stride < 1) pl.s_stride = 1;
    pl.s_tiles = (int)st;
    pl.gstride = (int64_t)pl.s_tiles * (kTileRows / kGroupCols);
    const double frac = (double)pl.s_tiles * kTileRows / (double)N;
    // expected candidates in the sample = ct * frac, spread over G groups of 32 rows; the sampling pass
    // keeps one maximum per group, so the rank to read off is the expected number of DISTINCT groups hit,
    // G (1 - exp(-ct frac / G)) -- equal to ct * frac when candidates are sparse, lower when they collide.
    const double G = (double)pl.gstride;
    const double hit = ct * frac;
    int m = (int)llround(G * (1.0 - exp(-hit / G)));
    if (m < 1) m = 1;
    pl.m_rank = m;
    int mq, qg;
    plan_scan(chunk, pl.s_tiles, h->num_sms, &mq, &qg, &pl.s_splits);
  }
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  pl.off_q16 = take((size_t)pl.qpad * h->d * 2);
  pl.off_q32 = take((size_t)pl.qpad * h->d * 4);
  pl.off_qnorm = take((size_t)pl.qpad * 4);
  pl.off_tau = take((size_t)pl.qpad * 4);
  pl.off_count = take((size_t)pl.qpad * pl.nseg * 4);
  pl.off_cscore = take((size_t)pl.qpad * pl.nseg * pl.cap_seg * 8);
  pl.off_cidx = pl.off_cscore;
  size_t aux = 0;
  if (pl.dense) aux = (size_t)pl.chunk * pl.dump_ld * 4;
  else if (!have_tau) aux = (size_t)pl.qpad * pl.gstride * 4;
  pl.off_aux = take(aux);
  pl.total = off;
  return pl;
}

int ensure_capacity(b2r_index* h, int64_t need, cudaStream_t stream) {
  if (need <= h->capacity) return B2R_OK;
  int64_t cap = h->capacity > 0 ? h->capacity + h->capacity / 2 : 0;
  if (cap < need) cap = need;
  cap = (int64_t)align_up((size_t)cap, 1024);
  float* n32 = nullptr;
  __nv_bfloat16* n16 = nullptr;
  if (cudaMalloc(&n32, (size_t)cap * h->d * 4) != cudaSuccess) {
    cudaGetLastError();
    return fail(B2R_ENOMEM, "cudaMalloc of fp32 corpus failed");
  }
  if (cudaMalloc(&n16, (size_t)cap * h->d * 2) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(n32);
    return fail(B2R_ENOMEM, "cudaMalloc of bf16 corpus failed");
  }
  if (h->ntotal > 0) {
    B2R_CUDA(cudaMemcpyAsync(n32, h->x32, (size_t)h->ntotal * h->d * 4, cudaMemcpyDeviceToDevice, stream));
    B2R_CUDA(cudaMemcpyAsync(n16, h->x16, (size_t)h->ntotal * h->d * 2, cudaMemcpyDeviceToDevice, stream));
  }
  B2R_CUDA(cudaStreamSynchronize(stream));
  cudaFree(h->x32);
  cudaFree(h->x16);
  h->x32 = n32;
  h->x16 = n16;
  h->capacity = cap;
  return B2R_OK;
}

int flat_search_chunk(b2r_index* h, const Plan& pl, int q, const float* queries, int normalize, int k,
                      float* D, int64_t* I, int32_t* status, float* tau_retry, const float* tau_in,
                      uint8_t* ws, cudaStream_t stream) {
  const int d = h->d;
  __nv_bfloat16* q16 = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_q16);
  float* q32 = reinterpret_cast<float*>(ws + pl.off_q32);
  float* qnorm = reinterpret_cast<float*>(ws + pl.off_qnorm);
  float* tau = reinterpret_cast<float*>(ws + pl.off_tau);
  int* count = reinterpret_cast<int*>(ws + pl.off_count);
  uint2* cand = reinterpret_cast<uint2*>(ws + pl.off_cscore);
  float* aux = reinterpret_cast<float*>(ws + pl.off_aux);
  int rc;
  // plan for THIS chunk size (the last chunk may be smaller)
  const ScanGeom geo = scan_geometry(h, q, pl.tiles, !pl.dense);
  const int MQ = geo.MQ, QG = geo.QG;
  int splits = geo.splits;
  const int qpad = geo.qpad;  // <= pl.qpad
  const int fp16 = h->scan_fp16 == 1;
  const float eps = (float)(fp16 ? h->eps_fp16 : h->eps);
  // tau[0, qpad) = +inf is written by the same launch (padding queries must never produce candidates)
  if ((rc = launch_prep_queries(queries, q, qpad, d, normalize, q32, q16, fp16, qnorm, stream, tau))) return rc;
  CUtensorMap tmQ;
  if ((rc = make_tmap_bf16_rows(&tmQ, q16, qpad, d))) return rc;

  ScanParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.Q = q;
  sp.QG = QG;
  sp.N = h->ntotal;
  sp.d = d;
  sp.tile_first = 0;
  sp.tile_stride = 1;
  sp.tile_count = pl.tiles;
  sp.splits = splits;
  sp.idesc = scan_idesc(fp16);
  sp.tau = tau;
  // the planned split count fixes the segment geometry; this chunk may use fewer splits
  if (splits > pl.splits) splits = pl.splits;
  sp.splits = splits;
  sp.cand_count = count;
  sp.cand = cand;
  sp.early_release = h->early_release;
  const int epi_warps = (MQ == 2 && h->epi_warps == 16) ? 16 : 8;
  // the chunk's segments must fit the planned buffer: nseg <= pl.nseg (cap_seg is the plan's)
  if (!pl.dense) {
    while (splits > 1 && splits * geo.seg_per_split > pl.nseg) --splits;
    sp.splits = splits;
  }
  sp.nseg = splits * geo.seg_per_split;
  sp.cap_seg = pl.cap_seg;
  int sel_nseg = sp.nseg, sel_cap_seg = pl.cap_seg;

  if (tau_in) {
    B2R_CUDA(cudaMemcpyAsync(tau, tau_in, (size_t)q * 4, cudaMemcpyDeviceToDevice, stream));
  } else if (pl.dense) {
    ScanParams dp = sp;
    dp.dump = aux;
    dp.ld = pl.dump_ld;
    if ((rc = launch_scan(SCAN_DUMP, MQ, tmQ, h->tmX, dp, h->num_sms, stream))) return rc;
    // exact k-th bf16 score per query, threshold = k-th - 2E: the candidates ARE the provable window
    const int m = (int64_t)k < h->ntotal ? k : (int)h->ntotal;
    if ((rc = launch_kth_value(aux, q, h->ntotal, pl.dump_ld, m, tau, count, cand, pl.cap_seg,
                               h->rescore ? qnorm : nullptr, h->maxnorm, eps, stream)))
      return rc;
    sel_nseg = 1;
    sel_cap_seg = pl.cap_seg;
  } else {
    ScanParams gp = sp;
    int mq, qg;
    gp.gmax = aux;
    const int s_pairs = pl.s_tiles / 2;
    if (geo.pair && h->pair_gmax && s_pairs >= 1) {
      // the sampling pass on CTA pairs too: s_tiles / 2 strided 256-row pair tiles, 8 group maxima each
      const int tiles2 = (int)ceil_div(pl.tiles, 2);
      gp.tile_count = s_pairs;
      gp.tile_stride = tiles2 / s_pairs > 1 ? tiles2 / s_pairs : 1;
      gp.gstride = s_pairs * 8;                       // <= pl.gstride: fits the planned buffer
      gp.idesc = scan_idesc_pair(fp16);
      plan_scan(q, s_pairs, h->num_sms / 2, &mq, &qg, &gp.splits);
      const double G = (double)gp.gstride, frac = (double)s_pairs * 2 * kTileRows / (double)(h->ntotal > 0 ? h->ntotal : 1);
      int m = (int)llround(G * (1.0 - exp(-(double)pl.c_target * (frac < 1.0 ? frac : 1.0) / G)));
      if (m < 1) m = 1;
      if ((rc = launch_scan_pair(tmQ, h->tmX, gp, h->num_sms, stream, 0, SCAN_GMAX))) return rc;
      if ((rc = launch_kth_value(aux, q, gp.gstride, gp.gstride, m, tau, nullptr, nullptr, 0, nullptr, nullptr, 0.f, stream)))
        return rc;
    } else {
      gp.tile_stride = pl.s_stride;
      gp.tile_count = pl.s_tiles;
      plan_scan(q, pl.s_tiles, h->num_sms, &mq, &qg, &gp.splits);
      gp.gstride = (int)pl.gstride;
      if ((rc = launch_scan(SCAN_GMAX, MQ, tmQ, h->tmX, gp, h->num_sms, stream))) return rc;
      if ((rc = launch_kth_value(aux, q, pl.gstride, pl.gstride, pl.m_rank, tau, nullptr, nullptr, 0, nullptr, nullptr,
                                 0.f, stream)))
        return rc;
    }
  }
  if (tau_in || !pl.dense) {
    const bool prof = h->profile && h->prof_used + 2 <= h->prof_ev.size();
    if (prof) cudaEventRecord(h->prof_ev[h->prof_used], stream);
    if (geo.pair) {
      ScanParams pp = sp;            // pair tiles of 256 rows, UMMA M = N = 256
      pp.tile_count = (int)ceil_div(pl.tiles, 2);
      pp.idesc = scan_idesc_pair(fp16);
      if ((rc = launch_scan_pair(tmQ, h->tmX, pp, h->num_sms, stream, h->walk))) return rc;
    } else if ((rc = launch_scan(SCAN_FILTER, MQ, tmQ, h->tmX, sp, h->num_sms, stream, epi_warps, h->walk))) {
      return rc;
    }
    if (prof) {
      cudaEventRecord(h->prof_ev[h->prof_used + 1], stream);
      h->prof_used += 2;
    }
  }
  SelectParams sel;
  memset(&sel, 0, sizeof(sel));
  sel.Q = q;
  sel.k = k;
  sel.d = d;
  sel.nseg = sel_nseg;
  sel.cap_seg = sel_cap_seg;
  sel.N = h->ntotal;
  sel.cand_count = count;
  sel.cand = cand;
  sel.tau = tau;
  sel.q32 = q32;
  sel.qnorm = qnorm;
  sel.x32 = h->x32;
  sel.maxnorm = h->maxnorm;
  sel.eps = eps;
  sel.rescore = h->rescore;
  sel.ids = (h->ids && h->n_ids >= h->ntotal) ? h->ids : nullptr;
  sel.label_base = h->label_base;
  sel.D = D;
  sel.I = I;
  sel.status = status;
  sel.tau_retry = tau_retry;
  return launch_select_rescore(sel, stream);
}

}  // namespace

namespace b2r {

int flat_create(b2r_index** out, int d, int device) {
  *out = nullptr;
  if (d < 64 || d > 256 || d % 64 != 0)
    return fail(B2R_EINVAL, "index_create: d must be a multiple of 64 in [64, 256]");
  int ndev = 0;
  B2R_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(B2R_EINVAL, "index_create: bad device ordinal");
  cudaDeviceProp prop;
  B2R_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(B2R_EUNSUPPORTED, std::string("index_create: device is sm_") + std::to_string(prop.major) +
                                      std::to_string(prop.minor) + ", this library is sm_100a only");
  DeviceGuard g(device);
  b2r_index* h = new b2r_index();
  h->kind = B2R_KIND_FLAT;
  h->d = d;
  h->metric = B2R_METRIC_IP;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  if (cudaMalloc(&h->maxnorm, 256) != cudaSuccess) {
    delete h;
    return fail(B2R_ENOMEM, "index_create: cudaMalloc failed");
  }
  cudaMemset(h->maxnorm, 0, 256);
  memset(&h->tmX, 0, sizeof(h->tmX));
  *out = h;
  return B2R_OK;
}

int flat_add(b2r_index* h, int64_t n, const float* x, int normalize, cudaStream_t stream) {
  if (n == 0) return B2R_OK;
  if (h->ntotal + n > 0x7FFFFF00ll) return fail(B2R_EUNSUPPORTED, "index_add: more than 2^31 rows per shard");
  int rc = ensure_capacity(h, h->ntotal + n, stream);
  if (rc) return rc;
  // scan format: fp16 while every stored row was normalised on ingest (|x_i| <= 1), else bf16
  int want16 = h->scan_dtype_req >= 0 ? h->scan_dtype_req : (normalize ? 1 : 0);
  if (h->ntotal > 0 && h->scan_fp16 == 0 && h->scan_dtype_req < 0) want16 = 0;   // once bf16, stay bf16
  if (h->ntotal > 0 && h->scan_fp16 >= 0 && h->scan_fp16 != want16) {
    if ((rc = launch_reencode(h->x32, h->ntotal, h->d, h->x16, want16, stream))) return rc;
  }
  h->scan_fp16 = want16;
  rc = launch_ingest(x, n, h->d, normalize, h->x32 + (size_t)h->ntotal * h->d,
                     h->x16 + (size_t)h->ntotal * h->d, want16, h->maxnorm, stream);
  if (rc) return rc;
  h->ntotal += n;
  return make_tmap_bf16_rows(&h->tmX, h->x16, h->ntotal, h->d);
}

size_t flat_search_workspace(const b2r_index* h, int q, int k) {
  if (!h || q <= 0 || k <= 0) return 0;
  const Plan a = make_plan(h, q, k, false);
  const Plan b = make_plan(h, q, k, true);
  return a.total > b.total ? a.total : b.total;
}

int flat_search(b2r_index* h, int q, const float* queries, int normalize, int k, float* D, int64_t* I,
                int32_t* status, float* tau_retry, const float* tau_in, void* workspace, size_t ws_bytes,
                cudaStream_t stream) {
  const Plan pl = make_plan(h, q, k, tau_in != nullptr);
  if (!workspace || ws_bytes < pl.total)
    return fail(B2R_ENOMEM, "index_search: workspace too small (need " + std::to_string(pl.total) + " bytes)");
  if (((uintptr_t)workspace & 255) != 0) return fail(B2R_EINVAL, "index_search: workspace must be 256-byte aligned");
  if (h->ntotal == 0) {
    // empty index: every slot is empty (faiss returns -1 labels)
    int rc = launch_fill_f32(D, (int64_t)q * k, -3.4028234663852886e38f, stream);
    if (rc) return rc;
    B2R_CUDA(cudaMemsetAsync(I, 0xFF, (size_t)q * k * 8, stream));
    if (status) B2R_CUDA(cudaMemsetAsync(status, 0, (size_t)q * 4, stream));
    if (tau_retry) B2R_CUDA(cudaMemsetAsync(tau_retry, 0, (size_t)q * 4, stream));
    return B2R_OK;
  }
  for (int q0 = 0; q0 < q; q0 += pl.chunk) {
    const int qc = (q - q0) < pl.chunk ? (q - q0) : pl.chunk;
    int rc = flat_search_chunk(h, pl, qc, queries + (size_t)q0 * h->d, normalize, k, D + (size_t)q0 * k,
                               I + (size_t)q0 * k, status ? status + q0 : nullptr,
                               tau_retry ? tau_retry + q0 : nullptr, tau_in ? tau_in + q0 : nullptr,
                               reinterpret_cast<uint8_t*>(workspace), stream);
    if (rc) return rc;
  }
  return B2R_OK;
}

}  // namespace b2r

// =============================================================== C ABI =====
extern "C" {

int b2r_version(void) { return B2R_VERSION; }
const char* b2r_last_error(void) { return t_error.c_str(); }
int64_t b2r_debug_launch_count(void) { return (int64_t)g_launches.load(); }

int b2r_index_create(b2r_index** out, int kind, int d, int nlist, int pq_m, int pq_bits, int metric,
                     int device) {
  if (!out) return fail(B2R_EINVAL, "index_create: out is NULL");
  *out = nullptr;
  if (kind != B2R_KIND_FLAT && kind != B2R_KIND_IVF_FLAT && kind != B2R_KIND_IVF_PQ)
    return fail(B2R_EINVAL, "index_create: unknown index kind");
  if (kind == B2R_KIND_FLAT && metric != B2R_METRIC_IP)
    return fail(B2R_EUNSUPPORTED, "index_create: the flat index supports inner product only");
  if (kind == B2R_KIND_IVF_FLAT && metric != B2R_METRIC_IP)
    return fail(B2R_EUNSUPPORTED, "index_create: IVF-Flat supports inner product only");
  if (kind == B2R_KIND_IVF_PQ && metric != B2R_METRIC_L2)
    return fail(B2R_EUNSUPPORTED, "index_create: IVF-PQ supports the L2 metric only (the reference's default)");
  b2r_index* h = nullptr;
  int rc = flat_create(&h, d, device);
  if (rc) return rc;
  h->kind = kind;
  h->metric = metric;
  if (kind != B2R_KIND_FLAT) {
    if (nlist < 1 || nlist > 65536) {
      b2r_index_destroy(h);
      return fail(B2R_EINVAL, "index_create: nlist must be in [1, 65536]");
    }
    h->nlist = nlist;
    h->trained = false;
    rc = flat_create(&h->quantizer, d, device);
    if (rc) {
      b2r_index_destroy(h);
      return rc;
    }
    if (kind == B2R_KIND_IVF_PQ) {
      if (pq_bits != 8 || pq_m < 1 || d % pq_m != 0 || (d / pq_m) % 4 != 0 || pq_m > 64) {
        b2r_index_destroy(h);
        return fail(B2R_EINVAL, "index_create: IVF-PQ needs pq_bits == 8, pq_m <= 64 dividing d with (d/pq_m) % 4 == 0");
      }
      h->pq_m = pq_m;
      h->pq_bits = pq_bits;
    }
  }
  *out = h;
  return B2R_OK;
}

int b2r_index_destroy(b2r_index* h) {
  if (!h) return B2R_OK;
  DeviceGuard g(h->device);
  cudaFree(h->x32);
  cudaFree(h->x16);
  cudaFree(h->maxnorm);
  cudaFree(h->ids);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  ivf_free(h);
  if (h->quantizer) b2r_index_destroy(h->quantizer);
  delete h;
  return B2R_OK;
}

int b2r_index_reserve(b2r_index* h, int64_t rows, void* stream) {
  if (!h) return fail(B2R_EINVAL, "index_reserve: NULL handle");
  if (rows < 0) return fail(B2R_EINVAL, "index_reserve: negative row count");
  if (rows > 0x7FFFFF00ll) return fail(B2R_EUNSUPPORTED, "index_reserve: more than 2^31 rows per shard");
  if (h->kind != B2R_KIND_FLAT) return B2R_OK;
  DeviceGuard g(h->device);
  if (rows <= h->capacity) return B2R_OK;
  // exact size (ensure_capacity over-allocates by 1.5x only when it grows an existing buffer on demand)
  const int64_t keep = h->capacity;
  h->capacity = 0;
  const int rc = ensure_capacity(h, rows, (cudaStream_t)stream);
  if (rc) h->capacity = keep;
  return rc;
}

int b2r_index_reset(b2r_index* h) {
  if (!h) return fail(B2R_EINVAL, "index_reset: NULL handle");
  DeviceGuard g(h->device);
  h->ntotal = 0;
  B2R_CUDA(cudaMemset(h->maxnorm, 0, 256));
  std::fill(h->list_sizes_host.begin(), h->list_sizes_host.end(), 0);
  h->list_sizes_desc_for = -1;
  if (h->list_off) B2R_CUDA(cudaMemset(h->list_off, 0, (size_t)(h->nlist + 1) * 8));
  return B2R_OK;
}

int64_t b2r_index_ntotal(const b2r_index* h) { return h ? h->ntotal : 0; }
int b2r_index_is_trained(const b2r_index* h) { return h ? (h->trained ? 1 : 0) : 0; }

int b2r_index_train(b2r_index* h, int64_t n, const float* x, uint64_t seed, void* stream) {
  if (!h) return fail(B2R_EINVAL, "index_train: NULL handle");
  if (h->kind == B2R_KIND_FLAT || h->trained) return B2R_OK;  // faiss: IndexFlat::train is a no-op
  if (n < 1 || !x) return fail(B2R_EINVAL, "index_train: no training vectors");
  DeviceGuard g(h->device);
  return ivf_train(h, n, x, seed, (cudaStream_t)stream);
}

int b2r_index_add(b2r_index* h, int64_t n, const float* x, int normalize, void* stream_) {
  if (!h) return fail(B2R_EINVAL, "index_add: NULL handle");
  if (n < 0 || (n > 0 && !x)) return fail(B2R_EINVAL, "index_add: bad arguments");
  if (n == 0) return B2R_OK;
  DeviceGuard g(h->device);
  if (h->kind == B2R_KIND_FLAT) return flat_add(h, n, x, normalize, (cudaStream_t)stream_);
  if (!h->trained) return fail(B2R_ESTATE, "index_add: index is not trained");
  return ivf_add(h, n, x, normalize, (cudaStream_t)stream_);
}

int b2r_index_set_ids(b2r_index* h, int64_t n, const int64_t* ids, void* stream_) {
  if (!h) return fail(B2R_EINVAL, "index_set_ids: NULL handle");
  DeviceGuard g(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n <= 0 || !ids) {
    B2R_CUDA(cudaStreamSynchronize(stream));
    cudaFree(h->ids);
    h->ids = nullptr;
    h->n_ids = 0;
    return B2R_OK;
  }
  B2R_CUDA(cudaStreamSynchronize(stream));
  cudaFree(h->ids);
  h->ids = nullptr;
  h->n_ids = 0;
  if (cudaMalloc(&h->ids, (size_t)n * 8) != cudaSuccess) {
    cudaGetLastError();
    return fail(B2R_ENOMEM, "index_set_ids: cudaMalloc failed");
  }
  B2R_CUDA(cudaMemcpyAsync(h->ids, ids, (size_t)n * 8, cudaMemcpyDeviceToDevice, stream));
  h->n_ids = n;
  return B2R_OK;
}

int b2r_index_set_label_base(b2r_index* h, int64_t base) {
  if (!h) return fail(B2R_EINVAL, "NULL handle");
  h->label_base = base;
  return B2R_OK;
}

int b2r_index_set_param(b2r_index* h, const char* name, double value) {
  if (!h || !name) return fail(B2R_EINVAL, "index_set_param: NULL argument");
  const std::string n(name);
  if (n == "rescore_eps") { if (value < 0) return fail(B2R_EINVAL, "rescore_eps < 0"); h->eps = value; h->eps_fp16 = value; }
  else if (n == "scan_dtype") {
    // -1 auto, 0 bf16, 1 fp16 (only for corpora whose rows are normalised on ingest); set before the first add
    if (h->ntotal > 0) return fail(B2R_ESTATE, "scan_dtype must be set before the first add");
    h->scan_dtype_req = (int)value;
  }
  else if (n == "cand_factor") { if (value < 1) return fail(B2R_EINVAL, "cand_factor < 1"); h->cand_factor = value; h->cand_factor_fp16 = value; }
  else if (n == "cand_cap") {
    const int c = (int)value;
    if (c < 64 || c > 4096 || (c & (c - 1))) return fail(B2R_EINVAL, "cand_cap must be a power of two in [64,4096]");
    h->cand_cap = c;
  }
  else if (n == "epi_warps") {
    if (value != 8 && value != 16) return fail(B2R_EINVAL, "epi_warps must be 8 or 16");
    h->epi_warps = (int)value;
  }
  else if (n == "pair_scan") h->pair_scan = value != 0;
  else if (n == "pair_gmax") h->pair_gmax = value != 0;
  else if (n == "early_release") h->early_release = value != 0;
  else if (n == "walk") h->walk = value != 0;
  else if (n == "rescore") h->rescore = value != 0;
  else if (n == "force_path") h->force_path = (int)value;
  else if (n == "dense_budget") h->dense_budget = (int64_t)value;
  else if (n == "ivf_sample") h->ivf_sample = (int)value;
  else if (n == "ivf_fused") h->ivf_fused = value != 0;
  else if (n == "ivf_sample_rows") {
    if (value < 4 || value > 128 || ((int)value & 3)) return fail(B2R_EINVAL, "ivf_sample_rows must be a multiple of 4 in [4,128]");
    h->ivf_sample_rows = (int)value;
  }
  else if (n == "ivf_debug") h->ivf_debug = (int)value;
  else if (n == "pq_scan_path") h->pq_scan_path = (int)value;
  else if (n == "profile") {
    // value > 0: (re)start timing of up to `value` filter-scan launches; 0: stop
    DeviceGuard g(h->device);
    h->profile = value > 0;
    h->prof_used = 0;
    const size_t want = value > 0 ? (size_t)value * 2 : 0;
    while (h->prof_ev.size() < want) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return fail(B2R_ECUDA, "cudaEventCreate failed");
      h->prof_ev.push_back(e);
    }
  }
  else return fail(B2R_EINVAL, "index_set_param: unknown parameter " + n);
  return B2R_OK;
}

double b2r_index_get_param(const b2r_index* h, const char* name) {
  if (!h || !name) return NAN;
  const std::string n(name);
  if (n == "rescore_eps") return h->scan_fp16 == 1 ? h->eps_fp16 : h->eps;
  if (n == "scan_dtype") return h->scan_fp16;
  if (n == "cand_factor") return h->cand_factor;
  if (n == "cand_cap") return h->cand_cap;
  if (n == "epi_warps") return h->epi_warps;
  if (n == "pair_scan") return h->pair_scan;
  if (n == "pair_gmax") return h->pair_gmax;
  if (n == "early_release") return h->early_release;
  if (n == "walk") return h->walk;
  if (n == "rescore") return h->rescore;
  if (n == "force_path") return h->force_path;
  if (n == "dense_budget") return (double)h->dense_budget;
  if (n == "ivf_sample") return h->ivf_sample;
  if (n == "ivf_fused") return h->ivf_fused;
  if (n == "ivf_sample_rows") return h->ivf_sample_rows;
  if (n == "num_sms") return h->num_sms;
  if (n == "scan_ms_avg" || n == "scan_launches") {
    // mean device time of the timed filter-scan launches (synchronises on their events)
    double total = 0;
    size_t cnt = 0;
    for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
      float ms = 0;
      if (cudaEventSynchronize(h->prof_ev[i + 1]) != cudaSuccess) break;
      if (cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]) != cudaSuccess) break;
      total += ms;
      ++cnt;
    }
    if (n == "scan_launches") return (double)cnt;
    return cnt ? total / cnt : NAN;
  }
  return NAN;
}

size_t b2r_index_search_workspace(const b2r_index* h, int q, int k, int nprobe) {
  if (!h || q <= 0 || k <= 0) return 0;
  if (h->kind == B2R_KIND_FLAT) return flat_search_workspace(h, q, k);
  return ivf_search_workspace(h, q, k, nprobe);
}

int b2r_index_search(b2r_index* h, int q, const float* queries, int normalize, int k, int nprobe,
                     float* D, int64_t* I, int32_t* status, float* tau_retry, const float* tau_in,
                     void* workspace, size_t ws_bytes, void* stream_) {
  if (!h) return fail(B2R_EINVAL, "index_search: NULL handle");
  if (q < 0 || k < 1 || (q > 0 && (!queries || !D || !I))) return fail(B2R_EINVAL, "index_search: bad arguments");
  if (k > 1024) return fail(B2R_EUNSUPPORTED, "index_search: k must be <= 1024");
  if (q == 0) return B2R_OK;
  if (!h->trained) return fail(B2R_ESTATE, "index_search: index is not trained");
  DeviceGuard g(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (h->kind == B2R_KIND_FLAT)
    return flat_search(h, q, queries, normalize, k, D, I, status, tau_retry, tau_in, workspace, ws_bytes, stream);
  if (tau_in) return fail(B2R_EUNSUPPORTED, "index_search: caller thresholds apply to the flat index only");
  if (tau_retry) B2R_CUDA(cudaMemsetAsync(tau_retry, 0, (size_t)q * 4, stream));
  return ivf_search(h, q, queries, normalize, k, nprobe, D, I, status, workspace, ws_bytes, stream);
}

int b2r_index_get_vectors(const b2r_index* h, int64_t row0, int64_t n, float* out, void* stream) {
  if (!h) return fail(B2R_EINVAL, "NULL handle");
  if (row0 < 0 || n < 0 || row0 + n > h->ntotal) return fail(B2R_EINVAL, "index_get_vectors: range out of bounds");
  if (n == 0) return B2R_OK;
  DeviceGuard g(h->device);
  B2R_CUDA(cudaMemcpyAsync(out, h->x32 + (size_t)row0 * h->d, (size_t)n * h->d * 4, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return B2R_OK;
}


// ------------------------------------------------------------------ debug ---
int b2r_debug_scores_tc(b2r_index* h, int q, const float* queries, int normalize, float* out,
                        void* workspace, size_t ws_bytes, void* stream_) {
  if (!h || q <= 0 || !queries || !out) return fail(B2R_EINVAL, "debug_scores_tc: bad arguments");
  if (h->ntotal == 0) return B2R_OK;
  DeviceGuard g(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  int MQ, QG, splits;
  const int tiles = (int)ceil_div(h->ntotal, kTileRows);
  plan_scan(q, tiles, h->num_sms, &MQ, &QG, &splits);
  const int qpad = QG * MQ * kQBlock;
  const size_t need = align_up((size_t)qpad * h->d * 2, 256) + align_up((size_t)qpad * h->d * 4, 256) +
                      align_up((size_t)qpad * 4, 256);
  if (!workspace || ws_bytes < need) return fail(B2R_ENOMEM, "debug_scores_tc: workspace too small");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  __nv_bfloat16* q16 = reinterpret_cast<__nv_bfloat16*>(ws);
  float* q32 = reinterpret_cast<float*>(ws + align_up((size_t)qpad * h->d * 2, 256));
  float* qn = reinterpret_cast<float*>(ws + align_up((size_t)qpad * h->d * 2, 256) + align_up((size_t)qpad * h->d * 4, 256));
  int rc;
  if ((rc = launch_prep_queries(queries, q, qpad, h->d, normalize, q32, q16, h->scan_fp16 == 1, qn, stream))) return rc;
  CUtensorMap tmQ;
  if ((rc = make_tmap_bf16_rows(&tmQ, q16, qpad, h->d))) return rc;
  ScanParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.Q = q;
  sp.QG = QG;
  sp.N = h->ntotal;
  sp.d = h->d;
  sp.tile_first = 0;
  sp.tile_stride = 1;
  sp.tile_count = tiles;
  sp.splits = splits;
  sp.idesc = scan_idesc(h->scan_fp16 == 1);
  sp.dump = out;
  sp.ld = h->ntotal;
  return launch_scan(SCAN_DUMP, MQ, tmQ, h->tmX, sp, h->num_sms, stream);
}

}  // extern "C"

// plain CUDA-core reference of the same 16-bit x 16-bit -> fp32 contraction (test-only): the query's
// scan copy is unit-norm (as in the product path), rounded to the index's scan format
namespace {
__device__ __forceinline__ float scan16_to_float(uint16_t bits, int fp16) {
  if (fp16) return __half2float(__ushort_as_half(bits));
  return __uint_as_float((uint32_t)bits << 16);
}
__device__ __forceinline__ float round_scan16(float v, int fp16) {
  if (fp16) return __half2float(__float2half_rn(v));
  return __bfloat162float(__float2bfloat16_rn(v));
}
__global__ void scores_simt_kernel(const uint16_t* __restrict__ x16, int64_t N, int d,
                                   const float* __restrict__ qin, int Q, int fp16, float* __restrict__ out) {
  const int q = blockIdx.y;
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float qs[1024];
  __shared__ float s_scale;
  if (threadIdx.x < 32) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < d; i += 32) ss += qin[(size_t)q * d + i] * qin[(size_t)q * d + i];
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (threadIdx.x == 0) s_scale = ss > 0.f ? 1.0f / sqrtf(ss) : 1.0f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = round_scan16(qin[(size_t)q * d + i] * s_scale, fp16);
  __syncthreads();
  if (row >= N) return;
  float acc = 0.f;
  for (int i = 0; i < d; ++i) acc = fmaf(qs[i], scan16_to_float(x16[(size_t)row * d + i], fp16), acc);
  out[(size_t)q * N + row] = acc;
}
}  // namespace

extern "C" int b2r_debug_scores_simt(b2r_index* h, int q, const float* queries, int normalize, float* out,
                                     void* stream_) {
  (void)normalize;
  if (!h || q <= 0 || !queries || !out) return fail(B2R_EINVAL, "debug_scores_simt: bad arguments");
  if (h->ntotal == 0) return B2R_OK;
  DeviceGuard g(h->device);
  dim3 grid((unsigned)ceil_div(h->ntotal, 256), (unsigned)q);
  scores_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(reinterpret_cast<const uint16_t*>(h->x16), h->ntotal,
                                                              h->d, queries, q, h->scan_fp16 == 1, out);
  B2R_CHECK_LAUNCH("scores_simt_kernel");
  return B2R_OK;
}
