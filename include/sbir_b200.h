/*
 * sbir_b200.h — C ABI of the B200-native retrieval / triplet hot path.
 *
 * This is the drop-in boundary for the stage of Peer222/art-sbir that follows the
 * sketch and artwork encoders.  The reference has no FFI of its
 * own: that stage is a handful of Python callables on torch tensors.  Every entry
 * point below names the reference call site it replaces (file:line relative to the
 * reference checkout); INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every buffer (inputs, outputs, workspace) is DEVICE memory owned by the
 *     caller unless the function name ends in `_host`;
 *   - row-major, contiguous `[rows, dim]` matrices; `dtype` selects the element
 *     type of the embedding matrices (SBIR_F32 or SBIR_BF16);
 *   - enqueue-only on `stream` (a cudaStream_t passed as void*); no device-wide
 *     synchronisation inside, re-entrant across streams;
 *   - return value is an sbir_status; 0 = success; nothing throws or aborts.
 */
#ifndef SBIR_B200_H_
#define SBIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBIR_B200_ABI_VERSION 2

typedef enum sbir_status {
  SBIR_OK = 0,
  SBIR_ERR_INVALID_ARG = 1, /* null pointer, negative size, bad enum                */
  SBIR_ERR_UNSUPPORTED = 2, /* shape outside what the kernels handle (see each fn)  */
  SBIR_ERR_CUDA = 3,        /* a CUDA runtime/driver call failed; see last_cuda_err */
  SBIR_ERR_WORKSPACE = 4,   /* workspace pointer null/unaligned or too small        */
  SBIR_ERR_NO_DEVICE = 5    /* no sm_100 device visible                             */
} sbir_status;

typedef enum sbir_dtype { SBIR_F32 = 0, SBIR_BF16 = 1 } sbir_dtype;

/* `loss_type` of the reference: 'euclidean' (utils.py:42) or 'cosine' (utils.py:31-40). */
typedef enum sbir_metric { SBIR_EUCLIDEAN = 0, SBIR_COSINE = 1 } sbir_metric;

/* ---- introspection -------------------------------------------------------- */
int sbir_abi_version(void);
const char* sbir_status_string(int status);
/* cudaError_t of the last failing CUDA call seen by this thread (0 if none). */
int sbir_last_cuda_error(void);
/* 1 if the current device is compute capability 10.x, else 0. */
int sbir_device_supported(void);

/* ---- H9: L2 normalisation (implicit in nn.CosineSimilarity, utils.py:34) ---
 * y[i,:] = x[i,:] / max(||x[i,:]||_2, eps).  In/out dtype identical.  y may alias x. */
int sbir_l2_normalize(const void* x, void* y, int64_t rows, int64_t dim, int dtype,
                      float eps, void* stream);

/* out[i] = ||x[i,:]||_2^2 (fp32, accumulated in fp64). */
int sbir_row_sqnorm(const void* x, int64_t rows, int64_t dim, int dtype, float* out,
                    void* stream);

/* ---- N1: gallery feature build (inference.py:72-92, compute_image_features) ---
 * The reference grows the gallery with torch.cat per batch of 50 encoder outputs (O(N²) copies),
 * moves it to the host and dumps CSV text (inference.py:85-90).  Here every block of encoder
 * output [rows, dim] (block_dtype fp32 or bf16, e.g. under autocast) is written straight into rows
 * [row0, row0 + rows) of a PREALLOCATED device matrix gallery[gallery_rows, dim] in the gallery's
 * own storage type (fp32 or bf16), L2-normalised first when normalize != 0 (H9's x / max(‖x‖,1e-8)),
 * and gallery_sqnorm[row0 + i] = ‖stored row‖² (fp32; may be NULL).  The norms are those of the
 * values AS STORED, which is what sbir_pairwise_topk takes as `g_sqnorm` to skip its own pass over
 * the gallery.  With one encoder replica per GPU each rank appends only its own row range and the
 * row-sharded layout of the multi-GPU path falls out for free. */
int sbir_gallery_append(const void* block, int block_dtype, int64_t rows, int64_t dim, void* gallery,
                        int gallery_dtype, int64_t gallery_rows, int64_t row0, float* gallery_sqnorm,
                        int normalize, void* stream);

/* ---- H1 / H2: row-wise distance, the reference's distance modules ----------
 * utils.euclidean_distance = nn.PairwiseDistance(p=2)  (utils.py:42; called at
 * inference.py:44,62) and utils.cosine_distance (utils.py:31-40; inference.py:46,64).
 * x1 is [rows1, dim], x2 is [rows2, dim]; rows1 and rows2 are equal, or one of
 * them is 1 (torch broadcasting).  out is fp32 [max(rows1, rows2)].
 *   euclidean: out[i] = || x1[i] - x2[i] + 1e-6 ||_2
 *   cosine   : out[i] = 1 - sum_d (x1/max(||x1||,1e-8)) * (x2/max(||x2||,1e-8))   */
int sbir_pairwise_distance(const void* x1, int64_t rows1, const void* x2, int64_t rows2,
                           int64_t dim, int dtype, int metric, float* out, void* stream);

/* Backward of sbir_pairwise_distance for fp32 inputs: given grad_out [rows],
 * writes grad_x1 [rows1, dim] and grad_x2 [rows2, dim] (either may be NULL).
 * A broadcast side (rows == 1) receives the sum over rows. */
int sbir_pairwise_distance_bwd(const float* x1, int64_t rows1, const float* x2, int64_t rows2,
                               int64_t dim, int metric, const float* grad_out,
                               float* grad_x1, float* grad_x2, void* stream);

/* ---- H1+H3+H4 batched: distance + top-K (+ rank of the positive) ------------
 * Replaces the per-query loop inference.py:109-121 → get_ranking_position
 * (inference.py:30-57) and get_topk_images (inference.py:60-69):
 *   for every query row q of Q[num_q, dim] against the gallery G[num_g, dim]
 *     out_dist[q, 0:k]  = the k smallest distances, ascending (fp32),
 *     out_index[q, 0:k] = their gallery row indices + index_offset (int64),
 *                         ties broken by ascending index,
 *     out_rank[q]       = 0-based position of gallery row pos_index[q] in that same
 *                         (distance, index) order over the WHOLE gallery (int64) — what the
 *                         reference finds with topk(len(G)) + nonzero at inference.py:49-52 —
 *                         when pos_index != NULL; pos_index[q] < 0 gives num_g (inference.py:39-41).
 * "distance" in the order is the exact reference distance rounded to fp32, so the order (and
 * therefore top-k, rank and recall@K) does not depend on how the gallery is split into shards.
 * The distance matrix is never materialised: tcgen05 tiles of Q·Gᵀ feed a fused
 * per-query selection, the K(+slack) survivors are re-scored with the exact
 * reference formula, and `out_uncertified[0]` counts queries whose selection could
 * not be proven exact (they are then recomputed by the brute-force exact kernel, so
 * results are exact either way; the counter is diagnostic).  Tile element types: bf16
 * embeddings -> kind::f16; fp32 embeddings -> kind::tf32, or (problems of >= 4e11 FLOP
 * whose rows are multiples of 8 elements) kind::f16 on bf16-rounded copies kept in the
 * workspace, certified with measured rounding residuals.  Behind that pass, device-gated
 * tiers take over what it cannot certify: a few queries are re-selected on kind::tf32
 * tiles as a small batch, more trigger a kind::tf32 pass over all queries, and when more
 * than 2 % are still unresolved a 3xTF32 pass (operand copies in the workspace: 3x the
 * inputs, when that is below 12 GiB; euclidean: centred on the gallery mean).
 * Rows may have ANY width: rows whose byte length is not a multiple of 16 are scored
 * through zero-padded tile copies in the workspace (exact kernels read the caller's rows);
 * when the rows ARE 16-byte multiples, q and g must be 16-byte aligned.
 * g_sqnorm (fp32 [num_g], may be NULL) = ‖g_j‖² of the gallery rows as stored, from
 * sbir_gallery_append / sbir_row_sqnorm or the feature-store sidecar; when given, the pass reads
 * num_g floats instead of the whole gallery to build its epilogue vector (ignored when fp32
 * rows are converted to bf16 selection copies: that pass reads the rows anyway).
 * k <= 116.
 * out_rank, pos_index, out_uncertified may be NULL. */
size_t sbir_pairwise_topk_workspace_bytes(int64_t num_q, int64_t num_g, int64_t dim, int k,
                                          int dtype, int metric, int want_rank);
int sbir_pairwise_topk(const void* q, int64_t num_q, const void* g, const float* g_sqnorm,
                       int64_t num_g, int64_t dim, int dtype, int metric, int k, int64_t index_offset,
                       const int64_t* pos_index, float* out_dist, int64_t* out_index,
                       int64_t* out_rank, int32_t* out_uncertified, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Rank-only partial results for a gallery SHARD (multi-GPU): same
 * as above but the positive's distance is supplied (pos_dist, fp64 [num_q], from
 * sbir_positive_distance on the owning shard, all-reduced by the caller) and the
 * count is local to this shard; the caller sums counts across shards.  pos_index_global
 * (int64 [num_q], may be NULL) is the positive's GLOBAL row, used only to break exact
 * fp32 distance ties by index so that shard counts add up to the single-pass rank. */
int sbir_positive_distance(const void* q, int64_t num_q, const void* g, int64_t num_g,
                           int64_t dim, int dtype, int metric, const int64_t* pos_index_local,
                           double* out_pos_dist, void* stream);
int sbir_pairwise_topk_shard(const void* q, int64_t num_q, const void* g, const float* g_sqnorm,
                             int64_t num_g, int64_t dim, int dtype, int metric, int k, int64_t index_offset,
                             const double* pos_dist, const int64_t* pos_index_global, float* out_dist,
                             int64_t* out_index, int64_t* out_count_less,
                             int32_t* out_uncertified, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ---- K4: merge of per-shard top-K lists (after the NCCL all-gather) ---------
 * dist/index hold num_lists lists of [num_q, k], each ascending; list l starts at
 * dist + l*list_stride_dist and index + l*list_stride_index (in elements; 0 = dense, num_q*k), so the
 * packed per-rank messages of ONE all-gather are merged where they landed.  Writes the k smallest
 * of the union per query, ascending, ties by ascending index. */
int sbir_topk_merge(const float* dist, const int64_t* index, int num_lists, int64_t list_stride_dist,
                    int64_t list_stride_index, int64_t num_q, int k, float* out_dist, int64_t* out_index,
                    void* stream);

/* ---- H5: retrieval metrics (inference.py:95-98,113-134) ---------------------
 * rank0 is the 0-based rank per query.  Writes, as fp64:
 *   out[0]            = mean reciprocal rank  (sum 1/(rank0+1) / num_q)
 *   out[1 .. k]       = topk_acc[0..k-1]      (#{rank0 <= i} / num_q)
 *   out[k+1]          = mean of (rank0+1),  out[k+2] = sample std (ddof=1),
 *   out[k+3]          = min, out[k+4] = max   (of rank0+1)
 * (quartiles need a sort; the host mirror computes them from the rank vector). */
int sbir_retrieval_metrics(const int64_t* rank0, int64_t num_q, int k, double* out, void* stream);

/* ---- H6 / H7: triplet margin loss, forward + backward in one launch ---------
 * nn.TripletMarginLoss(margin) (train.py:169) and
 * nn.TripletMarginWithDistanceLoss(margin, distance_function) (utils.py:56,69):
 *   loss = mean_i max(0, margin + d(a_i,p_i) - d(a_i,n_i)), d per `metric`.
 * a, p, n are fp32 [batch, dim].  out_loss is one fp32.  grad_* (fp32 [batch, dim],
 * any may be NULL) receive d loss / d input (i.e. already scaled by 1/batch).
 * out_per_row (fp32 [batch], may be NULL) receives the un-averaged hinge terms. */
int sbir_triplet_margin_loss(const float* a, const float* p, const float* n, int64_t batch,
                             int64_t dim, float margin, int metric, float* out_loss,
                             float* out_per_row, float* grad_a, float* grad_p, float* grad_n,
                             void* stream);

/* ---- H8: batch-hard triplet loss (north_star extension) ------
 * candidates X = cat(p, n) [2*batch, dim]; D_ij = d(a_i, X_j) on tcgen05 tiles;
 * positives of anchor i: {i} when labels == NULL, else {j : cand_label[j] == anchor_label[i]};
 * hp_i = max over positives, hn_i = min over the rest;
 * loss = mean_i max(0, margin + hp_i - hn_i).  Mining runs on the tensor-core
 * tiles; the selected pairs are re-scored with the exact distance, and gradients
 * flow through the selected pairs only (deterministic two-pass scatter).
 * out_hard_index int64 [batch, 2] (positive, negative candidate index; may be NULL). */
size_t sbir_batch_hard_workspace_bytes(int64_t batch, int64_t dim);
int sbir_batch_hard_triplet_loss(const float* a, const float* p, const float* n, int64_t batch,
                                 int64_t dim, float margin, int metric,
                                 const int64_t* anchor_label, const int64_t* cand_label,
                                 float* out_loss, int64_t* out_hard_index, float* grad_a,
                                 float* grad_p, float* grad_n, void* workspace,
                                 size_t workspace_bytes, void* stream);

/* ---- end-to-end with HOST buffers (bench.py `e2e`) ----------
 * q_host / g_host are host (preferably pinned) matrices; results land in host
 * buffers.  Gallery chunks are uploaded on a copy stream while earlier chunks are
 * scored; device staging memory is allocated internally and cached per device.
 * Synchronous: returns when out_* are valid. */
int sbir_retrieve_host(const void* q_host, int64_t num_q, const void* g_host, int64_t num_g,
                       int64_t dim, int dtype, int metric, int k, const int64_t* pos_index_host,
                       float* out_dist_host, int64_t* out_index_host, int64_t* out_rank_host,
                       int32_t* out_uncertified_host);
/* One rank's part of a gallery-sharded retrieval with the shard's rows in HOST (preferably pinned)
 * memory: sbir_pairwise_topk_shard semantics (global indices = local row + index_offset; local
 * count of rows closer than pos_dist; NaN pos_dist = no positive), with the rows uploaded in
 * chunks on a copy stream and scored as they arrive.  q, pos_dist, pos_index_global and the
 * outputs are DEVICE buffers; the work is ordered after `stream`; synchronous like
 * sbir_retrieve_host.  Used by art_sbir_b200/sharded.py: sharded_retrieve_host. */
int sbir_retrieve_host_shard(const void* q_dev, int64_t num_q, const void* g_host, int64_t num_g,
                             int64_t dim, int dtype, int metric, int k, int64_t index_offset,
                             const double* pos_dist_dev, const int64_t* pos_index_global_dev,
                             float* out_dist_dev, int64_t* out_index_dev, int64_t* out_count_less_dev,
                             int32_t* out_uncertified_host, void* stream);
/* Host utility for the sharded host path: dst[i, :] = src[index[i], :] for i < n (row_bytes each; a zero row where
 * index[i] is outside [0, num_rows)), with `threads` host threads (0: up to 8, one per MiB moved).  Used to collect
 * the rows of the positives a rank owns from its host-resident shard before they are uploaded. */
int sbir_gather_rows_host(const void* src, int64_t num_rows, int64_t row_bytes, const int64_t* index,
                          int64_t n, void* dst, int threads);
/* Frees the cached staging memory of sbir_retrieve_host / sbir_retrieve_host_shard. */
int sbir_release_host_staging(void);

/* ---- measurement hooks (bench.py) -------------------------------------------
 * With profiling enabled every launch of the distance/top-k kernel (K1) is bracketed by
 * CUDA events on its own stream.  sbir_profile_collect waits for them, returns the summed
 * K1 device time in milliseconds, the number of K1 launches, and the number of kernels
 * this library launched since the previous collect, and resets all three. */
int sbir_profile_enable(int on);
int sbir_profile_collect(double* k1_ms_sum, int64_t* k1_launches, int64_t* kernel_launches);

/* ---- debug / self-test ------------------------------------------------------
 * Writes the raw epilogue matrix E[num_q, num_g] (euclidean: ||g||²-2qg, cosine:
 * -q·g/max(||g||,eps)) computed by the tcgen05 tiles; used by tests to validate the
 * tensor-core path in isolation.  out_e is fp32 [num_q, num_g]. */
/* Process-wide tuning / test switches — the library never reads the environment on the launch path.
 * name ∈ { "k1_feed" (-1 auto | 0 off), "k1_pair" (0 auto | 1 single CTAs | 2 CTA pairs), "k1_qres" (-1 auto | 0 off),
 *          "k1_sel_bf16" (fp32 embeddings selected on bf16 copies: -1 auto | 0 never | 1 always),
 *          "k1_pair_coop" (1: CTA-pair launches are cooperative, 0: plain cluster launches — Nsight Compute cannot replay
 *          cooperative cluster launches), "k1_bands" (query tiles walked in n L2 bands: 0 / -1 off, n forced), "k1_q_early" (resident-query form: 1 = the next
 *          unit's query tile is stored while the current unit's last accumulator is worked on (default), 0 = at the unit's start), "k1_l2_hints" (L2 eviction
 *          hints of the resident-query form, bits 1|2|4; 0 = off, the default), "k1_chunk_mb" (0 auto), "host_chunk_rows" (0 auto), "watchdog_cycles" (device-side wait bound, default
 *          4e9, 0 = none: for compute-sanitizer / cuda-gdb), "k1_flags" (diagnostic bits, honoured only by a
 *          -DSBIR_DIAG build: sbir_debug_diag_build() == 1), "reset" (all defaults) }.
 * Set between calls, not while one is running.  Unknown names return SBIR_ERR_INVALID_ARG. */
int sbir_debug_set_option(const char* name, int64_t value);
int sbir_debug_diag_build(void);
/* Host-only: the work decomposition K1 would use (no device access).  out[13] = {cap,
 * lists_per_row, num_q_tiles, num_g_tiles, num_partitions, tiles_per_partition, num_chunks,
 * tiles_per_chunk, num_units, part_fastest, pair, q_tile_stride, tile_dtype (1: the tensor-core tiles read bf16 —
 * bf16 embeddings, or fp32 embeddings selected on their bf16 copies; 0: kind::tf32 on fp32)}. */
int sbir_debug_plan(int64_t num_q, int64_t num_g, int64_t dim, int k, int dtype, int num_sms, int32_t* out);
/* Profiling aid (-DSBIR_DIAG builds): when option k1_flags has bit 64 set, the distance kernel
 * records per CTA (8 uint64 each, 148 CTAs) the cycles its MMA issuer waited for a free
 * accumulator [0] and for operands [1], its whole loop [2], and the cycles epilogue warp 0 waited
 * for finished accumulators [3].  Copies up to n values to the host buffer `out` and clears them. */
int sbir_debug_k1_diag(uint64_t* out, int n);
size_t sbir_debug_dist_matrix_workspace_bytes(int64_t num_q, int64_t num_g, int64_t dim, int dtype);
int sbir_debug_dist_matrix(const void* q, int64_t num_q, const void* g, int64_t num_g,
                           int64_t dim, int dtype, int metric, float* out_e, void* workspace,
                           size_t workspace_bytes, void* stream);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* SBIR_B200_H_ */
