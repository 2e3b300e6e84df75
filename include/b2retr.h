/*
 * b2retr.h — C ABI of libb2retr.so: B200 (sm_100a) Stage-1 retrieval hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its
 * own: its arithmetic lives in faiss-cpu / torch behind a Python surface.  Every
 * entry point below names the reference call site whose arithmetic it replaces
 * (paths relative to the reference repo root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the ABI.
 *   - All data pointers are CALLER-OWNED DEVICE pointers unless the name ends in
 *     `_host`.  The caller keeps them alive until the stream work completes.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  All work is
 *     enqueued asynchronously on it; nothing synchronises unless stated.
 *   - Every function returns B2R_OK (0) or a negative B2R_E* code; the message is
 *     available from b2r_last_error() (thread local).  Nothing throws or exits.
 *   - The search path never allocates: the caller supplies a workspace of at least
 *     b2r_index_search_workspace() bytes (256-byte aligned).
 */
#ifndef B2RETR_H_
#define B2RETR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_VERSION 100

/* error codes */
#define B2R_OK 0
#define B2R_EINVAL (-1)     /* bad argument */
#define B2R_ECUDA (-2)      /* CUDA runtime / driver error (text in b2r_last_error) */
#define B2R_ENOMEM (-3)     /* device allocation failed / workspace too small */
#define B2R_ESTATE (-4)     /* wrong state (e.g. search on an untrained IVF index) */
#define B2R_EUNSUPPORTED (-5)

/* index kinds — faiss_retrieval.py:46-63 (`_create_index`) */
#define B2R_KIND_FLAT 0    /* faiss.IndexFlatIP            faiss_retrieval.py:48  */
#define B2R_KIND_IVF_FLAT 1 /* faiss.IndexIVFFlat(IP)      faiss_retrieval.py:52-55 */
#define B2R_KIND_IVF_PQ 2  /* faiss.IndexIVFPQ (L2)        faiss_retrieval.py:59-63 */

/* metrics */
#define B2R_METRIC_IP 0 /* larger is better, results descending */
#define B2R_METRIC_L2 1 /* squared L2, results ascending (IVFPQ default) */

/* per-query status bits written by b2r_index_search (0 = provably exact) */
#define B2R_ST_TOO_FEW 1       /* fewer than k candidates passed the threshold */
#define B2R_ST_NEED_LOWER_TAU 2 /* threshold above (k-th score - 2E): coverage not proven */
#define B2R_ST_CAND_OVERFLOW 4 /* candidate buffer overflowed */
#define B2R_ST_RESCORE_OVERFLOW 8 /* rescore window larger than its buffer */

typedef struct b2r_index b2r_index;
typedef struct b2r_tower b2r_tower;

int b2r_version(void);
const char* b2r_last_error(void);

/* ------------------------------------------------------------------ index -- */

/* Replaces FAISSIndex._create_index (faiss_retrieval.py:44-81).
 * d must be a multiple of 64 (TMA/UMMA K-chunk) and <= 256 for the tensor-core
 * scan.  nlist/pq_m/pq_bits are ignored for B2R_KIND_FLAT. */
int b2r_index_create(b2r_index** out, int kind, int d, int nlist, int pq_m, int pq_bits,
                     int metric, int device);
int b2r_index_destroy(b2r_index* h);

/* Drop all stored vectors (keeps training state). */
int b2r_index_reset(b2r_index* h);

/* index.ntotal / index.is_trained (read at train.py:231, inference.py:156,
 * faiss_retrieval.py:90,107,251-252). */
int64_t b2r_index_ntotal(const b2r_index* h);
int b2r_index_is_trained(const b2r_index* h);

/* Replaces index.train (faiss_retrieval.py:93): k-means for the coarse
 * quantiser (+ PQ codebooks).  x: fp32 [n,d] device, un-normalised (the
 * reference trains before normalising, faiss_retrieval.py:107-108 vs :114-115).
 * No-op for FLAT.  Synchronises the stream. */
int b2r_index_train(b2r_index* h, int64_t n, const float* x, uint64_t seed, void* stream);

/* Replaces `faiss.normalize_L2(x); index.add(x)` (faiss_retrieval.py:114-118).
 * x: fp32 [n,d] device, never modified.  normalize!=0 applies
 * x *= 1/sqrt(sum x^2) when the sum is > 0 (zero rows stay zero).
 * Stores an fp32 master row and a bf16 scan row per vector.  May (re)allocate;
 * synchronises the stream when it does. */
int b2r_index_add(b2r_index* h, int64_t n, const float* x, int normalize, void* stream);

/* Pre-size a FLAT index for `rows` vectors (the `np.vstack` of the whole corpus before
 * `index.add`, training_pipeline.py:524-531, made explicit): later adds up to that size never
 * reallocate, so the peak footprint stays at one copy of the corpus (a 50M-row shard is
 * 77 GB; growing into it would briefly need two).  No-op when already that large or for the
 * IVF kinds (they re-sort their storage on every add). */
int b2r_index_reserve(b2r_index* h, int64_t rows, void* stream);

/* Replaces the python id remap loop `id_map[idx]` (faiss_retrieval.py:159-160):
 * ids: int64 [ntotal] device copy of id_map (copied into the handle); search then
 * returns ids[label] (and ids[ntotal-1] for empty slots, the reference's
 * id_map[-1] wrap-around) instead of labels.  n==0 / ids==NULL clears the map. */
int b2r_index_set_ids(b2r_index* h, int64_t n, const int64_t* ids, void* stream);

/* Added to every returned label (row-sharded corpora: global = base + local). */
int b2r_index_set_label_base(b2r_index* h, int64_t base);

/* Tunables (all have defaults): scan_dtype = -1 auto | 0 bf16 | 1 fp16 (before the first
 * add; auto = fp16 while every add normalised its rows, else bf16 — same tensor rate, fp16 has
 * an 8x smaller rounding error hence a smaller rescore window); rescore_eps = relative scan
 * score error bound for the rescore window (default 2^-8*1.02 bf16 / 2^-10*1.02 fp16 =
 * rigorous); cand_factor = target candidates per query as a multiple of k (default 4 bf16 /
 * 2.5 fp16); rescore = 0 returns scan scores of the unit-norm query instead of exact fp32;
 * ivf_sample = 1 (default) lets IVF searches estimate the candidate threshold from a score sample
 * with an exact fallback, 0 always runs the exact radix passes (same answers, more sweeps);
 * pq_scan_path = 0 (default) picks the query-major ADC scan when pq_m is 8/16/32, 1 forces the older
 * one-CTA-per-(query,list) kernel (tests compare the two); epi_warps = 16 (default) | 8 epilogue warps of
 * the flat filter scan for batches above 128 queries (16: one candidate segment per 64-column half of a
 * corpus tile, 8-9 % faster; same answers); walk = 1 (default) | 0: the filter epilogue's append walk visits
 * only the 3-element sub-groups whose maximum passed the threshold, or all 8 scores of the group; force_path / dense_budget / profile /
 * ivf_debug are test and measurement hooks (see csrc/index.cu). */
int b2r_index_set_param(b2r_index* h, const char* name, double value);
double b2r_index_get_param(const b2r_index* h, const char* name);

/* Workspace bytes needed by b2r_index_search for (q, k, nprobe). */
size_t b2r_index_search_workspace(const b2r_index* h, int q, int k, int nprobe);

/* Replaces `faiss.normalize_L2(q); index.search(q, k)` (faiss_retrieval.py:146-155)
 * plus the id remap (:159-160).
 *   queries  fp32 [q,d] device, never modified; normalize as in add.
 *   D        fp32 [q,k] device: IP descending (L2 ascending for IVF_PQ);
 *            empty slots hold -3.4028235e38 (IP) / +3.4028235e38 (L2).
 *   I        int64 [q,k] device: labels (or mapped ids); empty slots -1
 *            (or ids[ntotal-1] when an id map is set).
 *   status   int32 [q] device, B2R_ST_* bits per query (0 = provably exact);
 *            tau_retry fp32 [q] device: threshold to pass back in `tau_in` for a
 *            retry of the flagged queries.  Both may be NULL.
 *   tau_in   fp32 [q] device or NULL: caller-provided candidate thresholds
 *            (skips the sampling pass).
 * Flagged queries: FLAT - pass tau_retry back as tau_in; IVF_FLAT with the fused list scan (chunks of >= 1024
 * queries, taken only when `status` is given) - re-run those queries with set_param("ivf_fused", 0), whose
 * threshold fallback is exact and in-kernel.  The Python wrappers do both.
 * Tunables added in round 2 (b2r_index_set_param): pair_scan = 1 (default) | 0: FILTER scan of batches > 128 on
 * CTA pairs (tcgen05 cta_group::2) or on single CTAs; ivf_fused = 1 (default) | 0.
 * Asynchronous on `stream`. */
int b2r_index_search(b2r_index* h, int q, const float* queries, int normalize, int k,
                     int nprobe, float* D, int64_t* I, int32_t* status, float* tau_retry,
                     const float* tau_in, void* workspace, size_t ws_bytes, void* stream);

/* Share IVF / PQ state with the oracle ("same centroids/codebooks" parity,
 * SURVEY.md §8c).  Host pointers. centroids fp32 [nlist,d]; codebooks fp32
 * [pq_m, 2^pq_bits, d/pq_m]. */
int b2r_index_export_centroids(const b2r_index* h, float* centroids_host);
int b2r_index_import_centroids(b2r_index* h, const float* centroids_host);
int b2r_index_export_codebooks(const b2r_index* h, float* codebooks_host);
int b2r_index_import_codebooks(b2r_index* h, const float* codebooks_host);

/* Inverted-list sizes (int64 [nlist], host) — for oracle cross-checks. */
int b2r_index_list_sizes(const b2r_index* h, int64_t* sizes_host);

/* PQ codes of the STORED rows [row0,row0+n): uint8 [n, pq_m] device (IVF_PQ only; parity plumbing). */
int b2r_index_get_codes(const b2r_index* h, int64_t row0, int64_t n, uint8_t* out, void* stream);

/* Restore pre-encoded vectors (load path, replaces faiss.read_index for IVF_PQ): codes uint8
 * [n, pq_m] and their inverted-list ids int64 [n], both device. */
int b2r_index_add_codes(b2r_index* h, int64_t n, const uint8_t* codes, const int64_t* list_ids, void* stream);

/* Insertion label of each STORED row [row0,row0+n) (int64, device): identity for FLAT; IVF
 * keeps rows sorted by inverted list, so label != storage position there. */
int b2r_index_get_labels(const b2r_index* h, int64_t row0, int64_t n, int64_t* out, void* stream);

/* Copy stored fp32 master rows [row0,row0+n) to a device buffer (save path,
 * replaces what faiss.write_index serialises, faiss_retrieval.py:203-206). */
int b2r_index_get_vectors(const b2r_index* h, int64_t row0, int64_t n, float* out, void* stream);

/* Replaces `faiss.write_index(index, path)` / `faiss.read_index(path)` (faiss_retrieval.py:203-206, :226) for
 * hosts that do not speak faiss's own file layout (the Python shim reads/writes that one, faiss_io.py): a
 * self-contained native container - header, coarse centroids, PQ codebooks, rows (or codes + list ids) in
 * insertion order, id map.  `path` is a host string.  save synchronises `stream`; load creates a NEW handle
 * on `device` that searches exactly like the saved one (same rows, same 16-bit scan format, same lists,
 * same id map).  The reference's pickled `.metadata` side-car stays the Python wrapper's business. */
int b2r_index_save(const b2r_index* h, const char* path, void* stream);
int b2r_index_load(b2r_index** out, const char* path, int device, void* stream);

/* ------------------------------------------------------------ multi-GPU -- */

/* Merge P per-shard results (after the NCCL all-gather).
 * D_all fp32 [P,q,k], I_all int64 [P,q,k] (each shard's rows best-first, empty
 * slots I=-1).  largest!=0 for IP.  Writes the global best-first top-k. */
int b2r_topk_merge(int P, int q, int k, const float* D_all, const int64_t* I_all, float* D_out,
                   int64_t* I_out, int largest, void* stream);

/* The same exchange with ONE buffer per rank and 8 instead of 12 bytes per result: b2r_topk_pack writes, per
 * query row, [k scores (fp32 bits)] [k LOCAL labels int32 = I - label_base, -1 for empty slots] [status]
 * (2k+1 int32).  Rows q..q_rows-1 are filled as empty lists (padding so that q_rows splits evenly into per-rank
 * query slices for an all-to-all).  After the collective, b2r_topk_merge_packed merges rows [0,q) of P such
 * blocks (block s starts at packed + s*q_stride*(2k+1), its labels get bases[s] added back; bases is a DEVICE
 * array of P int64) into best-first D_out/I_out and ORs the P status words into status_out (may be NULL).
 * Same result, tie order included, as b2r_topk_merge on the unpacked lists. */
int b2r_topk_pack(int q, int q_rows, int k, const float* D, const int64_t* I, const int32_t* status,
                  int64_t label_base, int32_t* out, int largest, void* stream);
int b2r_topk_merge_packed(int P, int q, int q_stride, int k, const int32_t* packed, const int64_t* bases,
                          float* D_out, int64_t* I_out, int32_t* status_out, int largest, void* stream);

/* ---------------------------------------------------------------- tower -- */

/* Replaces EmbeddingLayer.forward (two_tower_model.py:42-47): F per-field row
 * gathers + concat.  tables: device array of F device pointers, table f is fp32
 * [cards[f], emb_dim]; idx int64 [B,F]; out fp32 [B,ld] with ld >= F*emb_dim,
 * field f written at columns [f*emb_dim, (f+1)*emb_dim).  Bit-exact copy.
 * err_flag: int32 device, set to 1 on an out-of-range index (torch raises
 * IndexError there); may be NULL. */
int b2r_gather_concat(const float* const* tables, const int64_t* cards, int F, int emb_dim,
                      const int64_t* idx, int64_t B, float* out, int64_t ld, int32_t* err_flag,
                      void* stream);

/* BN-folded tower weights (host pointers, row-major [out,in] like nn.Linear).
 * Folding: W' = W*g/sqrt(var+eps), b' = (b-mean)*g/sqrt(var+eps)+beta
 * (two_tower_model.py:83-95 in eval mode). */
typedef struct b2r_tower_weights {
  int num_fields;           /* F */
  int emb_dim;              /* 16 */
  int num_numerical;        /* 13 for UserTower, 0 for AdTower */
  int hidden1, hidden2, out_dim; /* 512, 256, 256 */
  const int64_t* cards;     /* [F] host */
  const float* const* tables; /* [F] DEVICE pointers to fp32 [card,emb_dim] tables */
  const float* w1; const float* b1; /* [hidden1, F*emb_dim+num_numerical], [hidden1] host */
  const float* w2; const float* b2; /* [hidden2, hidden1] host */
  const float* w3; const float* b3; /* [out_dim, hidden2] host */
} b2r_tower_weights;

/* Replaces UserTower/AdTower construction + load_state_dict + eval(). */
int b2r_tower_create(b2r_tower** out, const b2r_tower_weights* w, int device);

/* The same for `hidden_dims` of ANY length (two_tower_model.py:83-95 / :152-164 build one
 * Linear+BatchNorm+ReLU+Dropout block per entry; train.py:350 exposes --hidden_dims with nargs='+').
 * num_layers = len(hidden_dims) + 1 Linear layers, 1..16; widths[l] = fan-out of layer l (the last one is
 * output_dim, <= 256); w[l] = BN-folded fp32 [widths[l], fan_in_l] row-major, b[l] = [widths[l]], all HOST.
 * b2r_tower_create is this call with num_layers = 3.  Towers with two hidden layers whose widths fit
 * (<= 512 / 256 / 256) take the fused kernel; every other shape runs gather + one tcgen05 GEMM launch per
 * layer (hidden activations 16-bit in the caller's workspace, see b2r_tower_workspace). */
typedef struct b2r_tower_layers {
  int num_fields, emb_dim, num_numerical;
  int num_layers;
  const int64_t* cards;       /* [F] host */
  const float* const* tables; /* [F] DEVICE pointers */
  const int* widths;          /* [num_layers] host */
  const float* const* w;      /* [num_layers] host pointers */
  const float* const* b;      /* [num_layers] host pointers */
} b2r_tower_layers;
int b2r_tower_create_layers(b2r_tower** out, const b2r_tower_layers* w, int device);
int b2r_tower_destroy(b2r_tower* t);
/* Workspace bytes b2r_tower_forward needs for a batch of B (0 when the fused kernel serves this tower:
 * it keeps every intermediate activation on the SM). */
size_t b2r_tower_workspace(const b2r_tower* t, int64_t B);

/* Tunables: operand_dtype = 0 fp16 (default: 8x smaller rounding error than bf16, same tensor rate) | 1 bf16
 * (fp32 range; chosen automatically at create time when a BN-folded weight exceeds the fp16 range, and by the
 * caller when a forward reports B2R_TOWER_SATURATED); force_path = 0 auto | 1 layer-by-layer kernels (gather +
 * one GEMM launch per layer, any widths, any depth) | 2 fused kernel (two hidden layers of widths <= 512 / 256,
 * out_dim <= 256 and % 4 == 0).
 * get_param also answers "fused" (1 when the next forward takes the fused kernel). */
int b2r_tower_set_param(b2r_tower* t, const char* name, double value);
double b2r_tower_get_param(const b2r_tower* t, const char* name);

/* bits OR-ed into *err_flag by b2r_tower_forward (the flag is never cleared by the library) */
#define B2R_TOWER_BAD_INDEX 1   /* a categorical id was outside [0, card): torch raises IndexError there */
#define B2R_TOWER_SATURATED 2   /* an fp16 operand (input or hidden activation) exceeded +-65504 and was clipped:
                                   the output of this call is not trustworthy, rerun with operand_dtype = 1 */

/* Replaces UserTower.forward / AdTower.forward (two_tower_model.py:98-121,
 * :167-184) in eval mode: gather+concat -> 3 GEMMs (16-bit operands, fp32
 * accumulate) with bias/ReLU -> F.normalize(p=2, eps=1e-12); one fused kernel
 * when the shape allows (see b2r_tower_set_param).
 * cat int64 [B,F]; num fp32 [B,num_numerical] or NULL; out fp32 [B,out_dim];
 * err_flag int32 device (B2R_TOWER_* bits are OR-ed in) or NULL. */
int b2r_tower_forward(b2r_tower* t, const int64_t* cat, const float* num, int64_t B, float* out,
                      int32_t* err_flag, void* workspace, size_t ws_bytes, void* stream);

/* Peer-memory exchange (csrc/peer.cu): the small-batch alternative to the NCCL all-gather of the packed lists.
 * Every rank creates a context (one cudaMalloc'd buffer: control block + receive area of `cap_bytes`), the 64-byte
 * cudaIpcMemHandle blobs are exchanged by the host (e.g. one torch.distributed all_gather), b2r_peer_connect maps
 * the peers' buffers.  b2r_peer_allgather(send, bytes) then copies `send` into slot [rank] of EVERY rank's receive
 * area with P2P stores over NVLink and waits (on the stream, bounded) until all P slots of the local area are
 * filled: *recv + r * bytes = rank r's data - the layout b2r_topk_merge_packed expects with q_stride = q.  After the
 * consumer kernel, b2r_peer_ack releases the area for the peers' next push.  Collective calls: same order and
 * same `bytes` on every rank.  Epochs live in device memory: the sequence can be captured in a CUDA graph. */
typedef struct b2r_peer b2r_peer;
int b2r_peer_create(b2r_peer** out, int rank, int world, size_t cap_bytes, int device);
int b2r_peer_destroy(b2r_peer* c);
int b2r_peer_handle(b2r_peer* c, void* handle64);
int b2r_peer_connect(b2r_peer* c, const void* handles);
int b2r_peer_allgather(b2r_peer* c, const void* send, size_t bytes, void** recv, void* stream);
int b2r_peer_ack(b2r_peer* c, void* stream);

/* ------------------------------------------------------ Stage-2 ranker -- */
/* TransformerRanker.forward (transformer_ranker.py:332-380) in eval mode, as called with
 * stage1_k = 500 candidate rows per user at inference.py:250-255 and faiss_retrieval.py:351-355.
 * With the sequence length of 1 the reference uses (x.unsqueeze(1), transformer_ranker.py:358) the softmax over
 * a single key is 1, so self-attention is W_o (W_v x + b_v) + b_o: the host folds it into ONE d_model x d_model
 * linear map per layer (w_attn, b_attn) and W_q / W_k drop out.  positional_encoding[:, 0] is folded into b_proj.
 * All weight pointers are HOST fp32, row-major [out, in]. */
typedef struct b2r_ranker b2r_ranker;
typedef struct b2r_ranker_weights {
  int n_user, n_ad;          /* categorical fields of the user / of the ad (6 / 20) */
  int emb_dim;               /* 32 */
  int num_numerical;         /* 13 */
  int d_model, d_ff;         /* 256, 1024 */
  int n_layers, n_cross;     /* 3 encoder layers, 3 cross layers */
  int n_tasks, head1, head2; /* 3 heads (ctr, engagement, revenue): d_model -> 256 -> 64 -> 1 */
  const int64_t* cards;      /* [n_user + n_ad] host: user tables first, in ModuleDict order */
  const float* const* tables;/* [n_user + n_ad] DEVICE pointers to fp32 [card, emb_dim] tables */
  const float* w_proj; const float* b_proj;   /* [d_model, (n_user+n_ad)*emb_dim + num_numerical], [d_model] */
  const float* w_attn; const float* b_attn;   /* [L, d_model, d_model], [L, d_model] (folded, see above) */
  const float* ln1_g; const float* ln1_b;     /* [L, d_model] */
  const float* w_fc1; const float* b_fc1;     /* [L, d_ff, d_model], [L, d_ff] */
  const float* w_fc2; const float* b_fc2;     /* [L, d_model, d_ff], [L, d_model] */
  const float* ln2_g; const float* ln2_b;     /* [L, d_model] */
  const float* w_cross; const float* b_cross; /* [C, d_model(out), d_model(in)] = cross_weights[i] TRANSPOSED, [C, d_model] */
  const float* w_h1; const float* b_h1;       /* [T, head1, d_model], [T, head1] */
  const float* w_h2; const float* b_h2;       /* [T, head2, head1], [T, head2] */
  const float* w_h3; const float* b_h3;       /* [T, head2], [T] */
} b2r_ranker_weights;

int b2r_ranker_create(b2r_ranker** out, const b2r_ranker_weights* w, int device);
int b2r_ranker_destroy(b2r_ranker* r);
size_t b2r_ranker_workspace(const b2r_ranker* r, int64_t B);
/* operand_dtype = 0 fp16 (default) | 1 bf16: as for the towers (B2R_TOWER_SATURATED -> rerun with 1) */
int b2r_ranker_set_param(b2r_ranker* r, const char* name, double value);
double b2r_ranker_get_param(const b2r_ranker* r, const char* name);
/* user_cat int64 [B, n_user]; ad_cat int64 [B, n_ad]; num fp32 [B, num_numerical]; out fp32 [n_tasks, B] = the
 * heads' raw outputs (the callers apply torch.sigmoid, inference.py:258-260); err_flag as in b2r_tower_forward. */
int b2r_ranker_forward(b2r_ranker* r, const int64_t* user_cat, const int64_t* ad_cat, const float* num, int64_t B,
                       float* out, int32_t* err_flag, void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- debug -- */
/* Test-only helpers (never on the product path). */

/* Full 16-bit-operand / fp32-accumulate SCAN score matrix through the SAME tcgen05 scan
 * kernel in dump mode: out fp32 [q, ntotal].  Scan scores are those of the UNIT-NORM
 * query copy against the stored scan rows (the product path thresholds on these and
 * re-scores the survivors exactly in fp32 with the caller's query scale). */
int b2r_debug_scores_tc(b2r_index* h, int q, const float* queries, int normalize, float* out,
                        void* workspace, size_t ws_bytes, void* stream);
/* The same matrix from a plain CUDA-core loop (no TMA / tensor cores). */
int b2r_debug_scores_simt(b2r_index* h, int q, const float* queries, int normalize, float* out,
                          void* stream);
/* Kernel launches issued by this library since load (all streams). */
int64_t b2r_debug_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B2RETR_H_ */
