// persist.cu — b2r_index_save / b2r_index_load: the native on-disk container of an index handle.
//
// Replaces `faiss.write_index(index, path)` / `faiss.read_index(path)` as FAISSIndex.save / .load call them
// (faiss_retrieval.py:196-245) for hosts that cannot (or need not) speak faiss's own layout — the faiss
// `IxFI` / `IwFl` / `IwPQ` layouts are read and written by the Python module faiss_io.py; this container is
// what a non-Python host gets through the C ABI.
//
// Layout (little endian): Header | centroids fp32 [nlist, d] (IVF kinds) | codebooks fp32 [pq_m, 256, d/pq_m]
// (IVF_PQ) | payload in LABEL (insertion) order: fp32 rows [ntotal, d] (FLAT, IVF_FLAT) or codes uint8
// [ntotal, pq_m] + list ids int64 [ntotal] (IVF_PQ) | id map int64 [n_ids].
// Loading goes through the same ingest paths as `add` (rows are stored already normalised, so they are added
// with normalize = 0 and the scan format recorded in the header): list membership is recomputed from the
// stored centroids and comes out identical, because assignment is a deterministic function of (row, centroids).
#include <stdio.h>
#include <string.h>

#include <vector>

#include "internal.h"

using namespace b2r;

namespace {

struct Header {
  char magic[8];          // "B2RIDX02"
  int32_t kind, d, nlist, pq_m, pq_bits, metric;
  int32_t scan_fp16;      // format of the 16-bit scan copy (-1: nothing stored)
  int32_t trained;
  int64_t ntotal, n_ids, label_base;
  int64_t reserved[4];
};
const char kMagic[8] = {'B', '2', 'R', 'I', 'D', 'X', '0', '2'};

struct File {
  FILE* f = nullptr;
  ~File() { if (f) fclose(f); }
};

bool put(FILE* f, const void* p, size_t bytes) { return bytes == 0 || fwrite(p, 1, bytes, f) == bytes; }
bool get(FILE* f, void* p, size_t bytes) { return bytes == 0 || fread(p, 1, bytes, f) == bytes; }

constexpr int64_t kChunkRows = 1 << 18;

}  // namespace

extern "C" {

int b2r_index_save(const b2r_index* h, const char* path, void* stream_) {
  if (!h || !path) return fail(B2R_EINVAL, "index_save: NULL argument");
  DeviceGuard g(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  B2R_CUDA(cudaStreamSynchronize(stream));
  File fh;
  fh.f = fopen(path, "wb");
  if (!fh.f) return fail(B2R_EINVAL, std::string("index_save: cannot open ") + path);
  Header hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, kMagic, 8);
  hd.kind = h->kind; hd.d = h->d; hd.nlist = h->nlist; hd.pq_m = h->pq_m; hd.pq_bits = h->pq_bits; hd.metric = h->metric;
  hd.scan_fp16 = h->scan_fp16;
  hd.trained = h->trained ? 1 : 0;
  hd.ntotal = h->ntotal;
  hd.n_ids = h->ids ? h->n_ids : 0;
  hd.label_base = h->label_base;
  if (!put(fh.f, &hd, sizeof(hd))) return fail(B2R_EINVAL, "index_save: write failed");
  const int d = h->d;
  if (h->kind != B2R_KIND_FLAT && h->trained) {
    std::vector<float> cent((size_t)h->nlist * d);
    B2R_CUDA(cudaMemcpy(cent.data(), h->quantizer->x32, cent.size() * 4, cudaMemcpyDeviceToHost));
    if (!put(fh.f, cent.data(), cent.size() * 4)) return fail(B2R_EINVAL, "index_save: write failed");
    if (h->kind == B2R_KIND_IVF_PQ) {
      std::vector<float> cb((size_t)256 * d);
      B2R_CUDA(cudaMemcpy(cb.data(), h->codebooks, cb.size() * 4, cudaMemcpyDeviceToHost));
      if (!put(fh.f, cb.data(), cb.size() * 4)) return fail(B2R_EINVAL, "index_save: write failed");
    }
  }
  const int64_t n = h->ntotal;
  if (n > 0) {
    // stored position -> label (identity for FLAT; IVF keeps rows sorted by inverted list)
    std::vector<uint32_t> perm;
    if (h->kind != B2R_KIND_FLAT) {
      perm.resize(n);
      B2R_CUDA(cudaMemcpy(perm.data(), h->perm, (size_t)n * 4, cudaMemcpyDeviceToHost));
    }
    if (h->kind == B2R_KIND_IVF_PQ) {
      const int m = h->pq_m;
      std::vector<uint8_t> stored((size_t)n * m), codes((size_t)n * m);
      std::vector<int32_t> rl(n);
      std::vector<int64_t> lists(n);
      B2R_CUDA(cudaMemcpy(stored.data(), h->codes, stored.size(), cudaMemcpyDeviceToHost));
      B2R_CUDA(cudaMemcpy(rl.data(), h->row_list, (size_t)n * 4, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < n; ++i) {
        memcpy(&codes[(size_t)perm[i] * m], &stored[(size_t)i * m], m);
        lists[perm[i]] = rl[i];
      }
      if (!put(fh.f, codes.data(), codes.size()) || !put(fh.f, lists.data(), (size_t)n * 8))
        return fail(B2R_EINVAL, "index_save: write failed");
    } else if (h->kind == B2R_KIND_FLAT) {
      std::vector<float> buf((size_t)kChunkRows * d);
      for (int64_t r0 = 0; r0 < n; r0 += kChunkRows) {
        const int64_t rc = (n - r0) < kChunkRows ? (n - r0) : kChunkRows;
        B2R_CUDA(cudaMemcpy(buf.data(), h->x32 + (size_t)r0 * d, (size_t)rc * d * 4, cudaMemcpyDeviceToHost));
        if (!put(fh.f, buf.data(), (size_t)rc * d * 4)) return fail(B2R_EINVAL, "index_save: write failed");
      }
    } else {
      // IVF_FLAT: un-permute on the host (the whole master copy: a save path, not a hot path)
      std::vector<float> stored((size_t)n * d), rows((size_t)n * d);
      B2R_CUDA(cudaMemcpy(stored.data(), h->x32, stored.size() * 4, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < n; ++i) memcpy(&rows[(size_t)perm[i] * d], &stored[(size_t)i * d], (size_t)d * 4);
      if (!put(fh.f, rows.data(), rows.size() * 4)) return fail(B2R_EINVAL, "index_save: write failed");
    }
  }
  if (hd.n_ids > 0) {
    std::vector<int64_t> ids(hd.n_ids);
    B2R_CUDA(cudaMemcpy(ids.data(), h->ids, (size_t)hd.n_ids * 8, cudaMemcpyDeviceToHost));
    if (!put(fh.f, ids.data(), (size_t)hd.n_ids * 8)) return fail(B2R_EINVAL, "index_save: write failed");
  }
  if (fflush(fh.f) != 0) return fail(B2R_EINVAL, "index_save: flush failed");
  return B2R_OK;
}

int b2r_index_load(b2r_index** out, const char* path, int device, void* stream_) {
  if (!out || !path) return fail(B2R_EINVAL, "index_load: NULL argument");
  *out = nullptr;
  cudaStream_t stream = (cudaStream_t)stream_;
  File fh;
  fh.f = fopen(path, "rb");
  if (!fh.f) return fail(B2R_EINVAL, std::string("index_load: cannot open ") + path);
  Header hd;
  if (!get(fh.f, &hd, sizeof(hd)) || memcmp(hd.magic, kMagic, 8) != 0)
    return fail(B2R_EINVAL, std::string("index_load: ") + path + " is not a b2r native index container");
  if (hd.ntotal < 0 || hd.n_ids < 0 || hd.d < 1 || hd.d > 4096) return fail(B2R_EINVAL, "index_load: corrupt header");
  b2r_index* h = nullptr;
  int rc = b2r_index_create(&h, hd.kind, hd.d, hd.nlist, hd.pq_m, hd.pq_bits, hd.metric, device);
  if (rc) return rc;
  DeviceGuard g(device);
  auto bail = [&](int code, const std::string& msg) {
    b2r_index_destroy(h);
    return fail(code, msg);
  };
  const int d = hd.d;
  h->label_base = hd.label_base;
  if (hd.kind != B2R_KIND_FLAT && hd.trained) {
    std::vector<float> cent((size_t)hd.nlist * d);
    if (!get(fh.f, cent.data(), cent.size() * 4)) return bail(B2R_EINVAL, "index_load: truncated file (centroids)");
    if (hd.kind == B2R_KIND_IVF_PQ) {
      std::vector<float> cb((size_t)256 * d);
      if (!get(fh.f, cb.data(), cb.size() * 4)) return bail(B2R_EINVAL, "index_load: truncated file (codebooks)");
      if ((rc = b2r_index_import_codebooks(h, cb.data()))) { b2r_index_destroy(h); return rc; }
    }
    if ((rc = b2r_index_import_centroids(h, cent.data()))) { b2r_index_destroy(h); return rc; }
  }
  const int64_t n = hd.ntotal;
  if (n > 0) {
    if (hd.kind == B2R_KIND_IVF_PQ) {
      const int m = hd.pq_m;
      std::vector<uint8_t> codes((size_t)n * m);
      std::vector<int64_t> lists(n);
      if (!get(fh.f, codes.data(), codes.size()) || !get(fh.f, lists.data(), (size_t)n * 8))
        return bail(B2R_EINVAL, "index_load: truncated file (codes)");
      uint8_t* dc = nullptr;
      int64_t* dl = nullptr;
      if (cudaMalloc(&dc, codes.size()) != cudaSuccess || cudaMalloc(&dl, (size_t)n * 8) != cudaSuccess) {
        cudaFree(dc);
        return bail(B2R_ENOMEM, "index_load: cudaMalloc failed");
      }
      cudaMemcpy(dc, codes.data(), codes.size(), cudaMemcpyHostToDevice);
      cudaMemcpy(dl, lists.data(), (size_t)n * 8, cudaMemcpyHostToDevice);
      rc = b2r_index_add_codes(h, n, dc, dl, stream);
      cudaStreamSynchronize(stream);
      cudaFree(dc);
      cudaFree(dl);
      if (rc) { b2r_index_destroy(h); return rc; }
    } else {
      // the stored rows are what `add` produced (already normalised when the caller asked for it): add them
      // verbatim, with the scan format the saved index used (a loaded index must search like the saved one)
      if (hd.scan_fp16 >= 0) h->scan_dtype_req = hd.scan_fp16;
      if (hd.kind == B2R_KIND_FLAT && (rc = b2r_index_reserve(h, n, stream))) { b2r_index_destroy(h); return rc; }
      // IVF re-sorts its whole storage on every add: feed it in one piece; FLAT streams in chunks
      const int64_t piece = hd.kind == B2R_KIND_FLAT ? kChunkRows : n;
      std::vector<float> buf((size_t)piece * d);
      float* dx = nullptr;
      if (cudaMalloc(&dx, (size_t)piece * d * 4) != cudaSuccess) return bail(B2R_ENOMEM, "index_load: cudaMalloc failed");
      for (int64_t r0 = 0; r0 < n && rc == B2R_OK; r0 += piece) {
        const int64_t rcnt = (n - r0) < piece ? (n - r0) : piece;
        if (!get(fh.f, buf.data(), (size_t)rcnt * d * 4)) { rc = fail(B2R_EINVAL, "index_load: truncated file (rows)"); break; }
        cudaMemcpy(dx, buf.data(), (size_t)rcnt * d * 4, cudaMemcpyHostToDevice);
        rc = b2r_index_add(h, rcnt, dx, 0, stream);
        cudaStreamSynchronize(stream);
      }
      cudaFree(dx);
      h->scan_dtype_req = -1;
      if (rc) { b2r_index_destroy(h); return rc; }
    }
  }
  if (hd.n_ids > 0) {
    std::vector<int64_t> ids(hd.n_ids);
    if (!get(fh.f, ids.data(), (size_t)hd.n_ids * 8)) return bail(B2R_EINVAL, "index_load: truncated file (ids)");
    int64_t* di = nullptr;
    if (cudaMalloc(&di, (size_t)hd.n_ids * 8) != cudaSuccess) return bail(B2R_ENOMEM, "index_load: cudaMalloc failed");
    cudaMemcpy(di, ids.data(), (size_t)hd.n_ids * 8, cudaMemcpyHostToDevice);
    rc = b2r_index_set_ids(h, hd.n_ids, di, stream);
    cudaStreamSynchronize(stream);
    cudaFree(di);
    if (rc) { b2r_index_destroy(h); return rc; }
  }
  *out = h;
  return B2R_OK;
}

}  // extern "C"
