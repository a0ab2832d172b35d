// kernels.h — internal launch interface between api.cu and the kernel translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace sbir {

// ---- api.cu: process-wide tuning / test switches, set through sbir_debug_set_option (never read from
// the environment on the launch path).  0 / -1 = the library's own choice.
struct DebugOptions {
  int k1_feed = -1;               // owner+feeder epilogue for 64/128-entry lists: -1 auto, 0 off
  int k1_bands = 0;               // L2 bands of query tiles (make_k1_plan): 0 auto, -1 never, n forced
  int k1_l2_hints = 0;            // resident-query form: L2 eviction hints (bits 1 gallery evict_last, 2 query tiles / 4 parked lists evict_first); 0 off (default: they cost time)
  int k1_q_early = 1;             // resident-query form: early store of the next unit's query tile (0: at the unit's start, A/B)
  int k1_pair = 0;                // CTA pairs: 0 auto, 1 never, 2 always
  int k1_qres = -1;               // resident-query form: -1 auto, 0 off
  int k1_pair_coop = 1;           // CTA-pair launches cooperative (1) or plain cluster launches (0: profilers that cannot replay them)
  int k1_sel_bf16 = -1;           // fp32 embeddings selected on bf16 copies (kind::f16): -1 auto, 0 never (kind::tf32), 1 always
  int k1_chunk_mb = 0;            // gallery bytes per chunk step (MB): 0 auto
  int k1_flags = 0;               // diagnostic bits, honoured by -DSBIR_DIAG builds only
  long long host_chunk_rows = 0;  // upload chunk of the host-buffer entry points (rows): 0 auto
  long long watchdog_cycles = 4000000000LL;  // bound on device-side waits: 0 = none
};
const DebugOptions& debug_options();

// ---- rowops.cu ----------------------------------------------------------------
// mode 0: ‖x‖² ; mode 1: −1/max(‖x‖,1e-8).  Rows [rows, rows_padded) get pad_value.
// max_sqnorm_out is cleared first unless accumulate_max (then the running maximum is kept).
int launch_row_norm(const void* x, int64_t rows, int64_t rows_padded, int64_t dim, int dtype,
                    int mode, float pad_value, float* out, float* max_sqnorm_out, cudaStream_t st,
                    bool accumulate_max = false);
// fp32 rows → bf16 selection operands + epilogue vector (exact fp32 norms, as launch_row_norm) + the norm of the
// rounding residual per row (res_row, queries) / its running maxima (res_max[0] abs, res_max[1] relative; gallery).
// max_sq and res_max ACCUMULATE (clear them before the first call).
int launch_convert_bf16_norm(const float* x, int64_t rows, int64_t rows_padded, int64_t dim, void* y_bf16, int mode,
                             float pad_value, float* vec, float* max_sq, float* res_row, float* res_max, cudaStream_t st);
// dst[r, 0:dim] = src[r, :], dst[r, dim:dim_pad] = 0 (rows whose byte length is not a 16-byte multiple, api.cu)
int launch_pad_rows(const void* src, int64_t rows, int64_t dim, int dtype, void* dst, int64_t dim_pad, cudaStream_t st);
// The same vector from stored ‖g‖² (gallery built by sbir_gallery_append / reloaded with its sidecar).
int launch_gvec_from_sqnorm(const float* sqnorm, int64_t rows, int64_t rows_padded, int mode, float pad_value,
                            float* out, float* max_out, cudaStream_t st, bool accumulate_max = false);
// N1: rows of encoder output → the gallery's storage type (+ optional L2 normalisation) + ‖stored row‖².
int launch_gallery_append(const void* block, int in_dtype, int64_t rows, int64_t dim, void* out_rows, int out_dtype,
                          float* out_sqnorm, int normalize, float eps, cudaStream_t st);
// out[i] = min(gvec[8 i .. 8 i + 7]) (NaN entries ignored)
int launch_chunk_min(const float* gvec, int64_t num_chunks, float* out, cudaStream_t st);
int launch_l2_normalize(const void* x, void* y, int64_t rows, int64_t dim, int dtype, float eps,
                        cudaStream_t st);
int launch_pairwise_distance(const void* x1, int64_t rows1, const void* x2, int64_t rows2,
                             int64_t dim, int dtype, int metric, float* out, cudaStream_t st);
int launch_pairwise_distance_bwd(const float* x1, int64_t rows1, const float* x2, int64_t rows2,
                                 int64_t dim, int metric, const float* grad_out, float* g1,
                                 float* g2, cudaStream_t st);
int launch_triplet(const float* a, const float* p, const float* n, int64_t batch, int64_t dim,
                   float margin, int metric, float* out_loss, float* per_row, float* ga, float* gp,
                   float* gn, cudaStream_t st);

// ---- dist_topk.cu / dist_topk_kernel.cuh (K1) ----------------------------------------------------------
constexpr int kTileQ = 128;        // query rows per tile (UMMA M, TMEM lanes)
constexpr int kTileG = 256;        // gallery rows per tile (UMMA N, TMEM columns)
constexpr int kUncertainPerQuery = 256;  // uncertain-pool capacity = this × num_q (min 65536)
constexpr int kMaxK = 116;         // largest supported k (list capacity 128 minus slack)

enum K1Mode { kModeTopk = 0, kModeTopkRank = 1, kModeDump = 2 };

// Work decomposition of one K1 launch: `num_splits` gallery partitions of `tiles_per_split` tiles
// (independent candidate lists), each scanned in `num_chunks` serial chunks; a unit is
// (query tile, partition, chunk) — see make_k1_plan (dist_topk.cu) / decode_unit (dist_topk_kernel.cuh).
struct K1Plan {
  int cap;            // per-list capacity (16, 32, 64 or 128)
  int lists_per_row;  // 1, or 2 (8 epilogue warps and cap <= 32: one list per column half)
  int epi_warps;      // 4 or 8 epilogue warps (8 with one list per row: owner + feeder warps, dist_topk_kernel.cuh)
  int qres;           // 1: resident-query form (query tile in tensor memory; bf16 rows of at most 1 KB)
  int num_q_tiles, num_g_tiles, num_splits, tiles_per_split, num_units, num_k_blocks;
  int band_q;         // unit-grid rows per L2 band (unit numbering, see decode_unit in dist_topk_kernel.cuh)
  int num_chunks, tiles_per_chunk;  // every partition is scanned in `num_chunks` serial chunks
  int part_fastest;   // unit numbering inside a chunk step: partitions (1) or query tiles (0) vary fastest
  int pair;           // 1: single-CTA tiles (cta_group::1); 2: CTA-pair tiles (cta_group::2, M = 256)
  int q_tile_stride;  // query-tile stride of candidate slots / shared thresholds (num_q_tiles rounded up to even)
  int lists_per_query() const { return num_splits * lists_per_row; }
};
K1Plan make_k1_plan(int64_t num_q, int64_t num_g, int64_t dim, int k, int dtype, int num_sms, int slack = 0);
// The plan of the first scoring pass of sbir_pairwise_topk for these inputs (api.cu: fp32 embeddings may be selected
// on bf16 copies, which changes tile bytes and chunking).
K1Plan topk_primary_plan(int64_t num_q, int64_t num_g, int64_t dim, int k, int dtype, int num_sms = 0);

struct K1Args {
  const void* q;
  const void* g;
  int64_t num_q, num_g, dim;
  int dtype, metric, mode;
  int chunk_begin, chunk_end;  // chunk steps [begin, end) of the plan this launch runs (0, 0 = all)
  const float* gvec;  // [num_g_tiles * kTileG] epilogue vector (‖g‖² | −1/max(‖g‖,eps)), padded
  const float* gmin;  // [num_g_tiles * kTileG / 8] minimum of gvec over each run of 8 rows (select modes)
  const int32_t* gate;  // optional device flag: every kernel of the launch is a no-op unless *gate != 0
  // top-k candidate lists, layout [partition][q_tile_stride][lists_per_row][cap][kTileQ]
  float* cand_val;
  int32_t* cand_idx;
  float* row_max;          // [partition][q_tile_stride][lists_per_row][kTileQ] carried list maximum
  int32_t* row_maxpos;     // ... and its position
  int32_t* chunk_done;     // [partition][q_tile_stride], zeroed by the caller
  uint32_t* unit_counter;  // [1], zeroed by the caller (dynamic unit hand-out)
  // rank (mode kModeTopkRank): e-space band per query and its outputs
  const float* rank_lo;
  const float* rank_hi;
  int32_t* cnt_less;     // [num_q], zeroed by the caller
  uint32_t* pool_count;  // [1] zeroed by the caller: entries claimed in the uncertain pool
  uint32_t pool_cap;     // capacity of the pool
  int32_t* pool_q;       // [pool_cap] query of each uncertain (query, gallery row) pair
  int32_t* pool_idx;     // [pool_cap] gallery row of each pair
  int32_t* dropped;      // [num_q] zeroed by the caller: pairs that did not fit in the pool
  // cross-split shared threshold per query row, ordered-int encoded, [num_q_tiles*kTileQ],
  // initialised to +inf (0x7f800000) by the caller (select modes)
  int32_t* shared_thr;
  // debug (mode kModeDump): full epilogue matrix [num_q][num_g]
  float* dump;
};
int launch_k1(const K1Args& args, const K1Plan& plan, cudaStream_t st);
// Reads and clears the per-CTA cycle counters the kernel fills when SBIR_K1_FLAGS & 64 (8 per CTA).
int k1_diag_read(unsigned long long* out, int n);

// ---- finalize.cu ----------------------------------------------------------------
struct FinalizeArgs {
  const void* q;
  const void* g;
  int64_t num_q, num_g, dim;
  int dtype, metric, k;
  int64_t index_offset;
  const float* cand_val;
  const int32_t* cand_idx;
  const float* qsq;        // [num_q] ‖q‖² (fp32)
  const float* gsq_max;    // [1] max ‖g‖²
  float kappa;             // error bound of the tensor-core dot product, relative to ‖q‖·‖g‖
  const float* q_res;      // [num_q] ‖q − bf16(q)‖ when fp32 embeddings were selected on bf16 copies, else NULL
  const float* g_res;      // [2] max ‖g − bf16(g)‖ (absolute, relative) in that case, else NULL
  float* out_dist;         // [num_q][k]
  int64_t* out_index;      // [num_q][k]
  int32_t* uncertified;    // [1] counter (may be NULL)
  int32_t* flags;          // [num_q] bit0: top-k selection not certified
  const int32_t* gate;     // optional device flag (escalation pass)
};
int launch_finalize_topk(const FinalizeArgs& a, const K1Plan& plan, cudaStream_t st);
// Brute-force exact top-k for the queries flagged by finalize (flags[q] & 1).
int launch_topk_fallback(const FinalizeArgs& a, cudaStream_t st);

struct RankArgs {
  const void* q;
  const void* g;
  int64_t num_q, num_g, dim;
  int dtype, metric;
  const int64_t* pos_index;  // [num_q] local gallery row of the positive, <0 = none (may be NULL)
  const double* pos_dist_in; // [num_q] externally supplied positive distance (NaN = none) or NULL
  const int64_t* pos_tie;    // [num_q] positive's index in the (local row + tie_offset) space, or NULL
  int64_t tie_offset;
  const float* qsq;
  const float* gsq_max;
  float kappa;
  const float* q_res;        // see FinalizeArgs
  const float* g_res;
  double* pos_dist;          // [num_q] workspace
  float* rank_lo;            // [num_q]
  float* rank_hi;            // [num_q]
  int32_t* cnt_less;
  uint32_t* pool_count;
  uint32_t pool_cap;
  int32_t* pool_q;
  int32_t* pool_idx;
  int32_t* dropped;
  int64_t* out_rank;         // [num_q]
  int64_t missing_rank;      // value for queries without a positive (num_g in the reference)
  const int32_t* gate;       // optional device flag (escalation pass)
};
int launch_rank_band(const RankArgs& a, cudaStream_t st);
int launch_rank_resolve(const RankArgs& a, cudaStream_t st);    // exact comparison of the pooled (query, row) pairs of a pass
int launch_rank_output(const RankArgs& a, cudaStream_t st);     // once at the end: rank from the counters / missing / exact brute force
// (device-gated) start of a scoring pass: counters, scheduler words and shared thresholds reset in one launch
int launch_pass_reset(int32_t* cnt_less, int32_t* dropped, int64_t num_q, uint32_t* pool_count, void* sched, size_t sched_bytes,
                      int32_t* shared_thr, int64_t num_thr, const int32_t* gate, cudaStream_t st);
// Tiers behind the bf16 selection of fp32 embeddings (finalize.cu: tier_decide_kernel): a few certificate failures ->
// those queries re-selected on kind::tf32 tiles as a batch of their own (fq / gather / scatter); more -> gate_full.
int launch_tier_decide(const int32_t* flags, const int32_t* dropped, int64_t num_q, int64_t max_sub, int32_t* fq, int32_t* fq_count,
                       int32_t* gate_sub, int32_t* gate_full, int32_t* uncertified, cudaStream_t st);
int launch_gather_sub(const float* q, int64_t dim, const int32_t* fq, const int32_t* fq_count, int64_t sub_q, float* q_sub,
                      const float* qsq, float* qsq_sub, const int32_t* gate, cudaStream_t st);
int launch_scatter_sub(const int32_t* fq, const int32_t* fq_count, int64_t sub_q, int k, const float* sub_dist, const int64_t* sub_index,
                       const int32_t* sub_flags, float* out_dist, int64_t* out_index, int32_t* flags, int32_t* uncertified,
                       const int32_t* gate, cudaStream_t st);
// Escalation (fp32 inputs): after the TF32 pass, gate[0] = 1 iff more than `max_bad` queries failed
// the top-k certificate or overflowed the rank pool; then the 3xTF32 pass re-does everything.
int launch_escalate_decide(const int32_t* flags, const int32_t* dropped, int64_t num_q, int64_t max_bad,
                           int32_t* gate, int32_t* uncertified, cudaStream_t st);
// x [rows, dim] fp32 -> out [rows, 3*dim]: query layout [hi | hi | lo], gallery layout [hi | lo | hi],
// hi = x with the 13 low mantissa bits cleared (what kind::tf32 reads), lo = x - hi (exact).
int launch_split_tf32(const float* x, int64_t rows, int64_t dim, int gallery_layout, float* out,
                      const int32_t* gate, cudaStream_t st);
// Mean-centred variant for the euclidean escalation pass: µ = column mean of the gallery
// (launch_col_mean; `partial` holds col_mean_workspace_bytes(dim)), then c = fl32(x − µ) is split
// as above and vec[r] = ‖c_r‖² (padded rows: pad_value; *max_out = max, cleared by launch_col_mean).
size_t col_mean_workspace_bytes(int64_t dim);
int launch_col_mean(const float* x, int64_t rows, int64_t dim, float* partial, float* mu, float* max_reset,
                    const int32_t* gate, cudaStream_t st);
int launch_center_split_tf32(const float* x, int64_t rows, int64_t rows_padded, int64_t dim, const float* mu,
                             int gallery_layout, float* out, float* vec, float pad_value, float* max_out,
                             const int32_t* gate, cudaStream_t st);
int launch_positive_distance(const void* q, int64_t num_q, const void* g, int64_t num_g, int64_t dim,
                             int dtype, int metric, const int64_t* pos_index, double* out,
                             cudaStream_t st);
// list l starts at dist + l·list_stride_dist / index + l·list_stride_index (elements; 0 = dense, num_q·k)
int launch_topk_merge(const float* dist, const int64_t* index, int num_lists, int64_t num_q, int k,
                      float* out_dist, int64_t* out_index, cudaStream_t st, int64_t list_stride_dist = 0,
                      int64_t list_stride_index = 0);
int launch_fill_i32(int32_t* out, int64_t n, int32_t value, cudaStream_t st);
int launch_fill_i64(int64_t* out, int64_t n, int64_t value, cudaStream_t st);
int launch_retrieval_metrics(const int64_t* rank0, int64_t num_q, int k, double* out, cudaStream_t st);

// ---- api.cu: one retrieval pass in phases (sbir_pairwise_topk*, sbir_retrieve_host) ----
// Workspace of sbir_pairwise_topk / _shard (all offsets 256-byte aligned).
struct TopkLayout {
  K1Plan plan;
  size_t off_gvec, off_gmax, off_qsq, off_cand_val, off_cand_idx, off_flags, off_uncert, off_shared_thr;
  size_t off_row_max, off_row_maxpos, off_sched, sched_bytes, off_gmin;
  int64_t kdim;        // columns of the rows the tensor-core tiles read (dim rounded up to a 16-byte multiple)
  bool padded_rows;    // kdim != dim: zero-padded operand copies at off_qpad / off_gpad
  size_t off_qpad, off_gpad;
  // tiers behind the bf16 selection (sel_bf16): kind::tf32 plans for all queries / for a subset of `sub_q` rows
  K1Plan plan_tf32, plan_sub;
  int64_t sub_q;
  size_t off_tier_gates, off_fq, off_qsub, off_qsq_sub, off_sub_dist, off_sub_index, off_sub_flags;
  size_t off_sub_cand_val, off_sub_cand_idx, off_sub_row_max, off_sub_row_maxpos, off_sub_sched, sub_sched_bytes, off_sub_thr;
  bool sel_bf16;       // fp32 inputs selected on bf16-rounded copies (kind::f16 tiles) — off_qb / off_gb / off_qres / off_gres
  size_t off_qb, off_gb, off_qres, off_gres;
  bool precise;        // fp32 inputs small enough for the 3xTF32 escalation workspace
  K1Plan plan3;        // plan of the escalation pass (dim' = 3·dim), same partitions / lists as `plan`
  size_t off_gate, off_q3, off_g3;  // off_sched: unit counter + chunk_done (zeroed together)
  size_t off_mu, off_colpart;       // column mean of the gallery + its partial sums (centred escalation pass)
  size_t off_pos_dist, off_lo, off_hi, off_cnt, off_dropped, off_pool_count, off_pool_q, off_pool_idx;
  uint32_t pool_cap;
  size_t total;
};
struct TopkPass {
  TopkLayout L;
  K1Args ka;
  FinalizeArgs fa;
  RankArgs ra;
  bool want_rank, done;  // done: nothing (left) to do (empty inputs, or finish ran)
  uint8_t* ws;
  cudaStream_t st;
  const void* q;
  const void* g;
  const void* kq;         // operands of the tensor-core tiles: q / g, or their zero-padded copies (TopkLayout::padded_rows)
  const void* kg;
  const float* g_sqnorm;  // optional stored ‖g‖² of the gallery rows (else computed from the rows)
  int64_t num_q, num_g, dim, padded, fed_rows;
  int dtype, metric;
  float *gvec, *gmax, *gmin, *qsq;
  int32_t *flags, *uncert;
};
int topk_pass_begin(TopkPass& P, const void* q, int64_t num_q, const void* g, const float* g_sqnorm, int64_t num_g, int64_t dim, int dtype,
                    int metric, int k, int64_t index_offset, const int64_t* pos_index, const double* pos_dist_in,
                    const int64_t* pos_tie, int64_t tie_offset, float* out_dist, int64_t* out_index,
                    int64_t* out_rank, int64_t missing_rank, int32_t* out_uncertified, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);
// Rows per intermediate feed must be a multiple of this (0: the gallery cannot be fed in pieces).
int64_t topk_pass_feed_granule(const TopkPass& P);
int topk_pass_feed(TopkPass& P, int64_t row_end);  // gallery rows [fed_rows, row_end) are resident
int topk_pass_finish(TopkPass& P);

// ---- batch_hard.cu (K3) -----------------------------------------------------------
size_t batch_hard_workspace_bytes(int64_t batch, int64_t dim);
int launch_batch_hard(const float* a, const float* p, const float* n, int64_t batch, int64_t dim,
                      float margin, int metric, const int64_t* anchor_label, const int64_t* cand_label,
                      float* out_loss, int64_t* out_hard_index, float* ga, float* gp, float* gn,
                      void* workspace, size_t workspace_bytes, cudaStream_t st);

// relative error bound of the tensor-core dot product
// Bound on |e_tc − e_exact| relative to (‖q‖² + ‖g‖²):
//   * kind::tf32 on fp32 data truncates both operands to 10 mantissa bits: 2^-9 (×1.01);
//   * every MMA k-step (32 bytes of K) adds into the fp32 accumulator with one rounding of at most
//     2^-23 of the running sum (≤ ‖q‖·‖g‖ ≤ (‖q‖²+‖g‖²)/2); with same-signed data (post-ReLU
//     features) these do not cancel, so the term grows linearly with the number of k-steps;
//   * the 3xTF32 split drops ql·gl and truncates the lo parts: 2^-20.
inline float k1_accum_kappa(int64_t dim, size_t elem_bytes) {
  const double ksteps = (double)(dim * (int64_t)elem_bytes + 31) / 32.0;
  return (float)((ksteps + 16.0) * 1.1920929e-07);  // (k-steps + 16) · 2^-23
}
inline float k1_kappa(int dtype, int64_t dim) {
  return dtype == 0 /*F32→tf32*/ ? 1.0f / 512.0f * 1.01f + k1_accum_kappa(dim, 4) : k1_accum_kappa(dim, 2);
}
inline float k1_kappa_precise(int64_t dim) { return k1_accum_kappa(3 * dim, 4) + 1.0f / 1048576.0f; }
// ... plus the rounding of the centring itself, fl32(x − µ): 2^-24 per element of either operand
inline float k1_kappa_centred(int64_t dim) { return k1_kappa_precise(dim) + 1.0f / 2097152.0f; }

}  // namespace sbir
