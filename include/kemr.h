/*
 * kemr.h -- C ABI of the B200-native retrieval-scoring engine (libkemr.so).
 *
 * The reference (REEVALUATE/knowledge_enhanced_multimodal_retrieval) is pure Python and has no
 * FFI of its own; its boundary for this path is a set of Python callables.
 * Every entry point below names the reference call site(s) whose arithmetic it replaces
 * (paths relative to the reference root).  The Python host layer in
 * knowledge_enhanced_multimodal_retrieval_b200/{metrics,fusion,retrieval}.py mirrors those callables
 * 1-for-1 and reaches this library through ctypes (see INTEGRATION.md for the stub a
 * maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - unless a parameter says "host", pointers are DEVICE pointers on the current device;
 *   - embeddings are bf16 bit patterns (uint16_t), row-major [rows, D], D % 8 == 0, D <= 1024,
 *     rows 16-byte aligned;
 *   - calls are enqueued on `stream` and do not synchronise (except the *_host calls);
 *   - every call returns KEMR_OK or an error code; kemr_last_error() describes the last failure
 *     of the calling thread;
 *   - gallery row indices inside one call are LOCAL to the shard passed in (0..M-1); `idx_base`
 *     is added when results are written, so shards of a row-partitioned gallery emit global ids.
 *
 * Scoring contract (the "canonical" score, identical to oracle/oracle.py canon_*):
 *     S_a(q,j)  = sum_d q[d]*gal_a[j][d]   accumulated in binary64 in the fixed 32-lane order
 *     clip(q,j) = fl(fl(w_a*S_a) + fl(w_b*S_b))          (gal_b == NULL: clip = fl(w_a*S_a))
 *     final(q,j)= fl(fl(alpha*clip) + bonus(q,j))        (bonus from the KG-hit CSR, else +0)
 *   ordering: final descending, ties by lowest gallery index.
 * The fast scan kernels compute clip in fp32 to SELECT candidates; every returned index/score and
 * every rank is then decided on canonical binary64 scores, and each query carries a certificate
 * bit saying the selection margin was provably wide enough (given |fp32 - canonical| <= eps).
 */
#ifndef KEMR_H_
#define KEMR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KEMR_ABI_VERSION 4

enum {
  KEMR_OK = 0,
  KEMR_ERR_ARG = 1,          /* bad argument (shape, alignment, NULL) */
  KEMR_ERR_CUDA = 2,         /* CUDA runtime / driver error */
  KEMR_ERR_WORKSPACE = 3,    /* workspace too small */
  KEMR_ERR_UNSUPPORTED = 4   /* not available on this device / build */
};

/* scan kernel selection */
enum {
  KEMR_PATH_AUTO = 0,
  KEMR_PATH_WARP = 1,        /* coalesced, vectorised HBM-streaming warp-dot kernel (small batches) */
  KEMR_PATH_MMA = 2          /* TMA-fed tcgen05/TMEM tensor-core kernel (sm_100a) */
};

/* per-query flag bits written by kemr_scan_topk / kemr_rank_count */
#define KEMR_FLAG_UNCERTIFIED 1   /* selection margin not provably sufficient: retry with larger k_sel/eps */
#define KEMR_FLAG_OVERFLOW 2      /* ambiguous-candidate list overflowed the workspace: retry with a larger one */

typedef void* kemr_stream_t;      /* cudaStream_t */

const char* kemr_last_error(void);
int kemr_abi_version(void);
/* sm count, compute capability, and whether the tcgen05 path can run here */
int kemr_device_info(int* sm_count, int* cc_major, int* cc_minor, int* has_tcgen05);

/* measurement hook: record `cuda_event` (a cudaEvent_t, or NULL to disable) right after the scan
 * kernel inside the following kemr_scan_topk / kemr_rank_count calls of this thread. */
int kemr_set_scan_done_event(void* cuda_event);
/* measurement hook: a DEVICE int64[3] (or NULL to disable) that the fused small-batch search kernel stamps with
 * %globaltimer nanoseconds on the following kemr_scan_topk calls of this thread: [0] first CTA start, [1] last CTA's
 * scan arrival (= scan phase done), [2] selection done. */
int kemr_set_phase_stamps(void* device_int64x3);

/* ---- embedding boundary: fp32 rows -> optional x/||x|| -> bf16 (round-to-nearest-even).
 * Replaces: evaluator.py:120-135 (normalise) + the implicit fp32 storage of the reference. */
int kemr_quantize_rows(const float* src, uint16_t* dst, int64_t rows, int D, int normalize,
                       kemr_stream_t stream);

/* ---- largest Euclidean row norm of a bf16 matrix (device float, never under-reported).  The selection margin of
 * the fp32 scan is eps * max||q|| * max||g|| * (|w_a| + |w_b|): callers that do not normalise
 * their embeddings (the reference does, evaluator.py:120-135) scale eps with it. */
int kemr_row_norm_max(const uint16_t* x, int64_t rows, int D, float* out_max, kemr_stream_t stream);

/* ---- deterministic synthetic gallery, generated on the device (counter-based, keyed by
 * seed and GLOBAL row index row_base+r so any shard can be regenerated anywhere). */
int kemr_synth_rows(uint16_t* dst, int64_t rows, int D, uint64_t seed, int64_t row_base,
                    kemr_stream_t stream);

/* ---- which scan kernel KEMR_PATH_AUTO would run for this shape on the current device
 * (KEMR_PATH_WARP or KEMR_PATH_MMA; the tcgen05 kernel needs enough gallery tiles for the candidate
 * lists k_sel asks for), and how many candidate lists ("parts") per query it would leave.
 * galleries = 1 or 2; equal_weights != 0 when both fusion weights are the same number. */
int kemr_scan_plan(int Q, int64_t M, int D, int galleries, int k_sel, int equal_weights, int* path, int* parts);

/* ---- workspace sizing for kemr_scan_topk / kemr_rank_count / kemr_score_matrix */
size_t kemr_workspace_bytes(int Q, int64_t M, int D, int k_sel, int64_t max_hits_per_query);

/* ---- similarity scan + weighted fusion + KG boost + top-k, score matrix never written.
 * Replaces: metrics.py:102,145-148 (sgemm + weighted sum), fusion.py:83 (alpha*S + w*I),
 *           metrics.py:34 (argsort, of which only the first k columns are used),
 *           and the remote CLIPRetriever.search scan (clip_retrieval.py:39-40).
 *   hit_rowptr/hit_col/hit_bonus: optional CSR of KG hits per query (unique local columns per
 *   query, bonus already aggregated); NULL = no boost.  alpha must be > 0.
 *   k_sel >= k: candidates kept by the fp32 scan (k + margin), <= 128.
 *   out_idx[q][i] = idx_base + local row, -1 past the end of a short gallery. */
int kemr_scan_topk(const uint16_t* q, int Q,
                   const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                   double w_a, double w_b, double alpha,
                   const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                   int64_t max_hits_per_query,
                   int k, int k_sel, double eps, int64_t idx_base,
                   double* out_score64, float* out_score32, int64_t* out_idx, int32_t* out_flags,
                   void* workspace, size_t workspace_bytes, int path, kemr_stream_t stream);

/* ---- canonical final score of arbitrary (query, local row) pairs (target scores, spot checks).
 *   pair_bonus may be NULL.  A pair whose row lies outside [0, M) scores NaN (ranks after everything). */
int kemr_score_pairs(const uint16_t* q, const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                     double w_a, double w_b, double alpha,
                     const int32_t* pair_q, const int64_t* pair_row, const double* pair_bonus,
                     int64_t n_pairs, double* out_score64, kemr_stream_t stream);

/* ---- number of rows of this shard ranked strictly ahead of each query's target:
 *   count[q] = #{ j : final(q,j) > t[q]  or  (final(q,j) == t[q] and idx_base+j < t_gidx[q]) }
 * so that rank = 1 + sum over shards.  Replaces the two full-row argsorts of
 * metrics.py:34,62 + :41,68 (position of the target) without materialising or sorting anything.
 *   t_score64 / t_gidx: canonical final score and GLOBAL index of each query's target. */
int kemr_rank_count(const uint16_t* q, int Q,
                    const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                    double w_a, double w_b, double alpha,
                    const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                    const double* t_score64, const int64_t* t_gidx, double eps, int64_t idx_base,
                    int64_t* out_count, int32_t* out_flags,
                    void* workspace, size_t workspace_bytes, int path, kemr_stream_t stream);

/* ---- per-query fusion weights: the gated fusion heads of the reference (fusion_model.py:9-23 SimpleGatedFusionWithBias,
 * :136-180 GatedFusionHead, :182-196 SimpleGatedFusion: scores = gate(q)*T2I + (1-gate(q))*T2T, evaluated there
 * block by block into an (N,N) host matrix, evaluator_fusion.py:76-121).  Same contract as the scalar calls with
 *     clip(q,j) = fl(fl(w_a_q[q]*S_a) + fl(w_b_q[q]*S_b)),   w_*_q: DEVICE arrays of Q binary64 weights;
 * both galleries are required.  The weights ride in the epilogue of the scan (one TMEM lane = one query). */
int kemr_scan_topk_gated(const uint16_t* q, int Q,
                         const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                         const double* w_a_q, const double* w_b_q, double alpha,
                         const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                         int64_t max_hits_per_query,
                         int k, int k_sel, double eps, int64_t idx_base,
                         double* out_score64, float* out_score32, int64_t* out_idx, int32_t* out_flags,
                         void* workspace, size_t workspace_bytes, int path, kemr_stream_t stream);
int kemr_rank_count_gated(const uint16_t* q, int Q,
                          const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                          const double* w_a_q, const double* w_b_q, double alpha,
                          const int64_t* hit_rowptr, const int32_t* hit_col, const double* hit_bonus,
                          const double* t_score64, const int64_t* t_gidx, double eps, int64_t idx_base,
                          int64_t* out_count, int32_t* out_flags,
                          void* workspace, size_t workspace_bytes, int path, kemr_stream_t stream);
int kemr_score_pairs_gated(const uint16_t* q, const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                           const double* w_a_q, const double* w_b_q, double alpha,
                           const int32_t* pair_q, const int64_t* pair_row, const double* pair_bonus,
                           int64_t n_pairs, double* out_score64, kemr_stream_t stream);
/* gate of the linear gated heads in fp32: gate[q] = sigmoid(sum_d q[q][d]*weight[d] + bias)
 * (fusion_model.py:18-19, :190-191); writes w_a_q = gate, w_b_q = fl32(1 - gate) widened to binary64. */
int kemr_gate_linear(const uint16_t* q, int Q, int D, const float* weight, float bias,
                     double* out_w_a_q, double* out_w_b_q, kemr_stream_t stream);

/* ---- dense fp32 fused similarity matrix out[q*ld + j] = fl32(w_a*S_a + w_b*S_b) as the scan
 * kernels compute it (compatibility with callers that want the matrix: metrics.py:102,145-148). */
int kemr_score_matrix(const uint16_t* q, int Q,
                      const uint16_t* gal_a, const uint16_t* gal_b, int64_t M, int D,
                      float w_a, float w_b, float* out, int64_t ld,
                      void* workspace, size_t workspace_bytes, int path, kemr_stream_t stream);

/* ---- matrix-taking compatibility entry points (metrics.py:13-44, :47-76, :165-185;
 * fusion.py:6-20): stable-descending rank of column target_col[q] in row q, NaN last. */
int kemr_matrix_rank(const float* S, int Q, int64_t M, int64_t ld, const int64_t* target_col,
                     int64_t* out_rank, kemr_stream_t stream);
/* top-k columns per row by (value desc, column asc); k <= 128 */
int kemr_matrix_topk(const float* S, int Q, int64_t M, int64_t ld, int k,
                     int64_t* out_idx, float* out_val, kemr_stream_t stream);

/* ---- dense KG fusion in fp32, bit-identical to numpy (fusion.py:83, :119-130, :180-204):
 *   out = scale_first ? fl32(fl32(alpha32*S) + 0) : S ;  then for every CSR entry of row q, in
 *   order:  out[q][hit_col] = fl32(out[q][hit_col] + hit_add).   out may not alias S. */
int kemr_matrix_fuse(const float* S, float* out, int Q, int64_t M, int64_t ld,
                     int scale_first, float alpha32,
                     const int64_t* hit_rowptr, const int32_t* hit_col, const float* hit_add,
                     kemr_stream_t stream);

/* ---- LinearFusionHead of the reference (fusion_model.py:25-48): out[i] = w2 . relu(W1 [S_a[i], S_b[i]] + b1) + b2 for
 * every element of two contiguous [Q, M] fp32 score matrices (T2I, T2T), in fp32 like torch; W1 is [hidden][2]
 * (nn.Linear layout), b1 / w2 [hidden].  out may alias neither input. */
int kemr_matrix_mlp2(const float* S_a, const float* S_b, float* out, int Q, int64_t M,
                     const float* w1, const float* b1, const float* w2, float b2, int hidden, kemr_stream_t stream);

/* ---- fused InfoNCE rows (train/losses.py:45-55): out_row_loss[i] = logsumexp_j(a_i.b_j / T) - a_i.b_i / T, the
 * per-row cross entropy of logits = A B^T / T against the diagonal, in fp32, without materialising the (B, B)
 * logits.  a, b: fp32 [B, D], D % 4 == 0, D <= 1024.  mean(out) = F.cross_entropy(logits, arange(B)); the
 * symmetric loss is the mean of two calls, (a, b) and (b, a). */
int kemr_infonce_rows(const float* a, const float* b, int B, int D, float temperature, float* out_row_loss,
                      kemr_stream_t stream);

/* ---- fused Recall@K / MRR / Mean-Rank reduction over 1-based ranks (metrics.py:41-42,70-71).
 *   out_hits[i] = #{rank <= k_values[i]};  out_stats[0] = sum(rank) (exact integer as double),
 *   out_stats[1] = sum(1/rank) in numpy's pairwise order, so that
 *   MRR = out_stats[1]/Q*100 is bit-identical to np.mean(1.0/pos)*100. */
int kemr_metrics_reduce(const int64_t* ranks, int Q, const int32_t* k_values, int n_k,
                        int64_t* out_hits, double* out_stats, kemr_stream_t stream);
/* host twin of the reduction above (no GPU needed; used to pin the summation order on CPU) */
int kemr_metrics_reduce_host(const int64_t* ranks_host, int Q, const int32_t* k_values_host, int n_k,
                             int64_t* out_hits_host, double* out_stats_host);

/* ---- merge of R per-shard top-k lists after the all-gather:
 *   in_score64/in_idx: [R][Q][k], every list ordered as kemr_scan_topk writes it ((score desc, idx asc), empty
 *   slots idx < 0 at the end); output top-k by (score desc, idx asc); R <= 64. */
int kemr_merge_topk(const double* in_score64, const int64_t* in_idx, int R, int Q, int k,
                    double* out_score64, int64_t* out_idx, kemr_stream_t stream);
/* the same with list r of a query at in_*[r * rank_stride + q * k] (entries; rank_stride >= Q*k): merges straight
 * out of an all-gather receive buffer laid out [rank][score | idx][Q][k] */
int kemr_merge_topk_strided(const double* in_score64, const int64_t* in_idx, int64_t rank_stride, int R, int Q, int k,
                            double* out_score64, int64_t* out_idx, kemr_stream_t stream);

/* ---- result exchange of the row-sharded search over NVLink peer memory ("each GPU's local top-k is
 * merged with a single all-gather").  Here the gather is FUSED into the selection kernel: every rank owns an exchange
 * buffer; the selection kernel of rank r stores the k result rows of each query straight into slot r of every rank's
 * buffer (peer-mapped stores over NVLink) and releases a per-query flag there; the merge kernel waits for the world's
 * flags in local memory.  No collective launch and no host synchronisation in the step.  Setup (once): every rank
 * calls kemr_peer_create, exchanges the 64-byte IPC handles out of band (e.g. torch.distributed.all_gather) and calls
 * kemr_peer_connect with the world's handles in rank order; ranks that live in ONE process pass the other ranks'
 * kemr_peer_local_buffer pointers to kemr_peer_connect_pointers instead.  A step on every rank:
 *     kemr_peer_begin(peer, stream);                         new epoch; arms the push of this thread's next scan
 *     kemr_scan_topk[_gated](...);                           as usual (idx_base = first global row of the shard)
 *     kemr_peer_merge(peer, Q, k, out_score64, out_idx, stream);   global top-k on every rank
 * All ranks must run the same sequence of steps with the same Q and k (world <= 8, one NVSwitch box). */
typedef struct kemr_peer kemr_peer_t;
int kemr_peer_create(int rank, int world, int max_queries, int max_k, kemr_peer_t** out, void* ipc_handle_host64);
int kemr_peer_connect(kemr_peer_t* peer, const void* ipc_handles_host /* [world][64] */);
int kemr_peer_connect_pointers(kemr_peer_t* peer, void* const* bases_host /* [world] device pointers */);
void* kemr_peer_local_buffer(kemr_peer_t* peer);
int kemr_peer_destroy(kemr_peer_t* peer);
int kemr_peer_begin(kemr_peer_t* peer, kemr_stream_t stream);
int kemr_peer_merge(kemr_peer_t* peer, int Q, int k, double* out_score64, int64_t* out_idx, kemr_stream_t stream);
/* instead of the merge: the ranks' rows as they are, out_*[rank][Q][k] (every rank searched its OWN queries against a
 * replica of the gallery and wants everybody's results: an all-gather without a collective launch) */
int kemr_peer_gather(kemr_peer_t* peer, int Q, int k, double* out_score64, int64_t* out_idx, kemr_stream_t stream);

/* ---- resident gallery handle with HOST-buffer search: the serving-side drop-in for
 * CLIPRetriever.search (clip_retrieval.py:39-40 -> retrieval.py:80,98).  The handle owns its
 * device copies of the galleries, a workspace, pinned staging buffers and a stream.
 *   gal_*_host: bf16 bit patterns [M, D] in host memory (gal_b_host may be NULL). */
typedef struct kemr_index kemr_index_t;
int kemr_index_create(const uint16_t* gal_a_host, const uint16_t* gal_b_host, int64_t M, int D,
                      int max_queries, int max_k, kemr_index_t** out);
int kemr_index_destroy(kemr_index_t* index);
/* queries: fp32 [Q, D] host (quantised to bf16 on the device; set normalize=1 to L2-normalise).
 * Copies in, scans, copies the k results out and synchronises.  hits_* host CSR or NULL.
 * Routes (same results on every one): one or two queries -> ONE host-to-device copy of the request (queries + CSR) and
 * ONE kernel (quantise, scan, KG hits, selection), results stored straight into page-locked host memory; larger
 * batches with page-locked caller buffers -> the quantise kernel reads the queries in place over PCIe and the
 * selection kernel writes the caller's arrays in place; pageable caller buffers -> staged through the handle's
 * page-locked buffers, batches of 512 queries and more in chunks (memcpy, transfer and scan of consecutive chunks
 * overlap).
 * Selection margin: 2e-5 * max(1, |w_a| max||g_a|| + |w_b| max||g_b||), the galleries' largest row norms measured once
 * at kemr_index_create -- un-normalised galleries or weights above one widen the margin instead of voiding the
 * certificate.  Queries are unit rows with normalize=1; with normalize=0 they must satisfy ||q|| <= 1 up to bf16
 * rounding (CLIP embeddings do; otherwise pass normalize=1, which leaves every query's ranking unchanged). */
int kemr_index_search_host(kemr_index_t* index, const float* q_host, int Q, int normalize,
                           double w_a, double w_b, double alpha,
                           const int64_t* hit_rowptr_host, const int32_t* hit_col_host,
                           const double* hit_bonus_host,
                           int k, int64_t* out_idx_host, double* out_score64_host,
                           int32_t* out_flags_host);

/* Pipelined use: kemr_index_submit_host queues the same search and returns; kemr_index_wait blocks until the results
 * are in the caller's arrays.  One search per handle at a time; kemr_index_share makes a second LANE over the same
 * resident galleries (own stream, workspace and staging buffers; the source handle must outlive it), so that the next
 * batch's queries cross PCIe while this batch is scanned.  Page-locked caller buffers are used in place and must stay
 * untouched until the wait returns. */
int kemr_index_share(kemr_index_t* source, kemr_index_t** out_lane);
int kemr_index_submit_host(kemr_index_t* index, const float* q_host, int Q, int normalize,
                           double w_a, double w_b, double alpha,
                           const int64_t* hit_rowptr_host, const int32_t* hit_col_host,
                           const double* hit_bonus_host,
                           int k, int64_t* out_idx_host, double* out_score64_host,
                           int32_t* out_flags_host);
int kemr_index_wait(kemr_index_t* index);

/* the same with queries that are already bf16 bit patterns [Q, D] (e.g. the output of a bf16 encoder): half the bytes
 * cross PCIe and no quantise kernel runs. */
int kemr_index_search_host_bf16(kemr_index_t* index, const uint16_t* q_bf16_host, int Q,
                                double w_a, double w_b, double alpha,
                                const int64_t* hit_rowptr_host, const int32_t* hit_col_host,
                                const double* hit_bonus_host,
                                int k, int64_t* out_idx_host, double* out_score64_host,
                                int32_t* out_flags_host);

/* ---- the data path either side of the scan ------------------------------------

 * KG-hit CSR builder on the device.  Input: per query the LIST of gallery rows the knowledge graph returned, in
 * list order, as GLOBAL row ids (-1 = unknown artefact), CSR-shaped (list_rowptr int64[Q+1], list_rows int64[nnz]),
 * plus the bonus of ONE listing per query (w, delta or delta*omega(|R(q)|): fusion.py:83,130,202).  Output: the
 * CSR kemr_scan_topk / kemr_rank_count take for the shard of rows [row_lo, row_hi): unique LOCAL columns per query
 * in order of first listing; sum_repeats = 0 keeps one bonus per row (indicator, fusion.py:80), 1 adds it once per
 * listing (fusion.py:130).  out_col / out_bonus need room for nnz entries; out_max_per_query is a device int64. */
size_t kemr_hits_workspace_bytes(int Q);
int kemr_hits_build_csr(const int64_t* list_rowptr, const int64_t* list_rows, const double* bonus_per_query,
                        int Q, int64_t row_lo, int64_t row_hi, int sum_repeats,
                        int64_t* out_rowptr, int32_t* out_col, double* out_bonus, int64_t* out_max_per_query,
                        void* workspace, size_t workspace_bytes, kemr_stream_t stream);

/* CSR -> CSR on the device: output query i is input query query_sel[i] (NULL: the same queries), entries whose column
 * lies in [col_lo, col_hi) are kept in order and re-based to col - col_lo.  The per-shard hit lists of a row-sharded
 * gallery and the subset of queries a certificate sends back for a wider selection.
 * out_col / out_bonus need room for the input's nnz; workspace as for kemr_hits_build_csr (Q = Q_out). */
int kemr_hits_filter_csr(const int64_t* rowptr, const int32_t* col, const double* bonus, const int64_t* query_sel,
                         int Q_out, int64_t col_lo, int64_t col_hi,
                         int64_t* out_rowptr, int32_t* out_col, double* out_bonus, int64_t* out_max_per_query,
                         void* workspace, size_t workspace_bytes, kemr_stream_t stream);
/* bonus of every query's own target column (0 where the target is not a hit): the KG term of the target's score in
 * the rank path (fusion.py:83 evaluated at column i of row i). */
int kemr_hits_target_bonus(const int64_t* rowptr, const int32_t* col, const double* bonus, int Q,
                           const int64_t* target_col, double* out_bonus, kemr_stream_t stream);

/* uuid -> gallery row map on the HOST (fusion.py:62 artefact_uuid_to_idx).  Keys are passed as one byte blob plus
 * n+1 offsets.  A repeated uuid keeps its last row, like the reference's dict.  Lookup returns -1 for unknown keys;
 * normalize_uri != 0 first cuts the key to its last '/' segment (fusion.py:76, text2sparql_retrieval.py:57). */
typedef struct kemr_idmap kemr_idmap_t;
int kemr_idmap_create(const char* blob_host, const int64_t* offsets_host, int64_t n, kemr_idmap_t** out);
int kemr_idmap_destroy(kemr_idmap_t* map);
int kemr_idmap_lookup(const kemr_idmap_t* map, const char* blob_host, const int64_t* offsets_host, int64_t n,
                      int normalize_uri, int64_t* out_rows_host);

/* persisted bf16 embedding store (replaces the reference's data/embeddings directory, clip_retrieval.py:28,35):
 * one file = 64-byte header + M*D bf16, row-major, so a rank reads its shard as a byte range.  kemr_store_load
 * streams rows [row_lo, row_hi) into DEVICE memory through two page-locked buffers and synchronises `stream`. */
int kemr_store_write(const char* path, const uint16_t* rows_host, int64_t M, int D);
int kemr_store_info(const char* path, int64_t* rows, int* dim);
int kemr_store_load(const char* path, int64_t row_lo, int64_t row_hi, uint16_t* dst_device, kemr_stream_t stream);

/* ---- debug build only: %globaltimer stamps (ns) of the phases of query 0's selection in the last launch */
int kemr_debug_select_stamps(int64_t* out16_host);

/* ---- introspection of the tcgen05 scan's work plan for a shape on a hypothetical device (no GPU needed): see
 * kemr_api.cu; used by the CPU property tests of the scheduling logic. */
int kemr_debug_mma_plan(int Q, int64_t M, int D, int galleries, int k_sel, int equal_weights, int sms, int quads,
                        int64_t* out16_host);

#ifdef __cplusplus
}
#endif
#endif /* KEMR_H_ */
