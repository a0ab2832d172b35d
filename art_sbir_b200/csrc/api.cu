// api.cu — the extern "C" boundary declared in include/sbir_b200.h: argument checking,
// workspace layout, and the launch sequences.  No kernel code lives here.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace sbir {

static thread_local int g_last_cuda_error = 0;
void set_last_cuda_error(int err) { g_last_cuda_error = err; }

// ---- profiling state (process-wide; bench.py only) ----
static std::atomic<long long> g_kernel_launches{0};
static std::atomic<int> g_profile_on{0};
static std::mutex g_profile_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_k1_events;
static cudaEvent_t g_k1_open = nullptr;

// ---- tuning / test switches (process-wide; set before launching, not thread-safe against running calls) ----
static DebugOptions g_debug_options;
const DebugOptions& debug_options() { return g_debug_options; }

void count_kernel_launch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }
void profile_k1_begin(cudaStream_t st) {
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_profile_mu);
  cudaEvent_t e0 = nullptr;
  if (cudaEventCreate(&e0) != cudaSuccess) return;
  cudaEventRecord(e0, st);
  g_k1_open = e0;
}
void profile_k1_end(cudaStream_t st) {
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_profile_mu);
  if (g_k1_open == nullptr) return;
  cudaEvent_t e1 = nullptr;
  if (cudaEventCreate(&e1) != cudaSuccess) return;
  cudaEventRecord(e1, st);
  g_k1_events.emplace_back(g_k1_open, e1);
  g_k1_open = nullptr;
}

namespace {

int num_sms_cached() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

bool dtype_ok(int d) { return d == SBIR_F32 || d == SBIR_BF16; }
bool metric_ok(int m) { return m == SBIR_EUCLIDEAN || m == SBIR_COSINE; }

// fp32 embeddings are selected on their bf16-rounded copies (kind::f16 tiles: twice the kind::tf32 rate; the
// certificate uses measured rounding-residual norms, rowops.cu: convert_bf16_norm_kernel) when the copies fit
// the workspace budget and bf16 rows satisfy TMA's 16-byte pitch.  Option k1_sel_bf16 = 0 keeps kind::tf32.
// Rows whose byte length is not a multiple of 16 (TMA's pitch granularity; the reference accepts any width): the
// tensor-core tiles then read zero-padded copies of the operands kept in the workspace (`kdim` columns), while every
// exact kernel — norms, re-scoring, rank resolution, fallbacks — keeps reading the caller's rows.  Zero columns add
// nothing to q·g or the norms, so the tiles compute the same values.
int64_t tile_dim(int64_t dim, int dtype) {
  const int64_t per16 = 16 / (int64_t)elem_size(dtype);
  return (dim + per16 - 1) / per16 * per16;
}

// The bf16 rounding band is about twice kind::tf32's: measured on the clustered 12.5k x 75k x 2048 workload, k = 10
// (32-entry lists) and k = 30 (64) certify every query and run 6.1 -> 3.8 ms / 6.5 -> 4.3 ms, while k = 100 (128-entry
// lists, 28 entries of slack) leaves 0.6 % of the queries uncertified.  Those are caught by TIERS behind the pass
// (topk_pass_finish): a few certificate failures are re-selected on kind::tf32 tiles as a small batch of their own;
// more (or overflowing rank pools) trigger one kind::tf32 pass over all queries; only then the 3xTF32 pass and the
// brute-force fallbacks.  Problems of a few 1e10 FLOP (the reference's own 1k x 10k evaluation) are launch-bound and
// stay on the path with fewer kernels and lists.
bool select_on_bf16(int64_t num_q, int64_t num_g, int64_t dim, int k, int dtype) {
  const int opt = debug_options().k1_sel_bf16;
  if (dtype != SBIR_F32 || opt == 0) return false;
  if (dim % 8 != 0 || num_g <= 0) return false;  // (padded rows stay on their own element type)
  if (((size_t)num_q + (size_t)num_g) * (size_t)dim * 2 > (size_t(16) << 30)) return false;
  if (opt > 0) return true;  // forced (tests, A/B runs)
  (void)k;
  return 2.0 * (double)dim * (double)num_q * (double)num_g >= 4e11;
}

TopkLayout topk_layout(int64_t num_q, int64_t num_g, int64_t dim, int k, int dtype, int want_rank) {
  TopkLayout L{};
  L.kdim = tile_dim(dim, dtype);
  L.padded_rows = L.kdim != dim;
  L.sel_bf16 = select_on_bf16(num_q, num_g, dim, k, dtype);
  // fp32 embeddings keep their wider candidate slack (k + 16) whichever tensor path selects them
  L.plan = topk_primary_plan(num_q, num_g, dim, k, dtype, 0);
  // 3xTF32 escalation copies ([rows, 3·dim] fp32) — only when they stay below 12 GiB
  const size_t split_bytes = ((size_t)num_q + (size_t)num_g) * (size_t)L.kdim * 3 * sizeof(float);
  L.precise = dtype == SBIR_F32 && num_g > 0 && split_bytes <= (size_t(12) << 30);
  if (L.precise) {
    L.plan3 = make_k1_plan(num_q, num_g, 3 * L.kdim, k, dtype, num_sms_cached(), 16);
    // the escalation pass re-uses the candidate / list-state buffers: they are sized for the larger of the two
    // plans; the passes must agree on the per-list capacity (finalize's certificate compares like with like)
    L.precise = L.plan3.cap == L.plan.cap;
  }
  auto lists_of = [](const K1Plan& p) { return (size_t)p.num_splits * p.q_tile_stride * p.lists_per_row; };
  size_t lists = lists_of(L.plan), parts = (size_t)L.plan.num_splits;
  if (L.precise) { lists = std::max(lists, lists_of(L.plan3)); parts = std::max(parts, (size_t)L.plan3.num_splits); }
  if (L.sel_bf16) {
    // tiers: the kind::tf32 plan for all queries shares the candidate / list-state buffers; the subset pass has its own
    L.plan_tf32 = make_k1_plan(num_q, num_g, L.kdim, k, SBIR_F32, num_sms_cached(), 16);
    int64_t max_bad = num_q / 50;
    if (max_bad < 4) max_bad = 4;
    L.sub_q = (max_bad + kTileQ - 1) / kTileQ * kTileQ;
    L.plan_sub = make_k1_plan(L.sub_q, num_g, L.kdim, k, SBIR_F32, num_sms_cached(), 16);
    if (L.plan_tf32.cap != L.plan.cap || L.plan_sub.cap != L.plan.cap) L.sel_bf16 = false;  // (cannot happen: same k and slack)
  }
  if (L.sel_bf16) { lists = std::max(lists, lists_of(L.plan_tf32)); parts = std::max(parts, (size_t)L.plan_tf32.num_splits); }
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o = align_up(o + (bytes ? bytes : 1), 256); return r; };
  const size_t nq = (size_t)(num_q > 0 ? num_q : 1);
  L.off_gvec = take((size_t)L.plan.num_g_tiles * kTileG * sizeof(float));
  L.off_gmax = take(sizeof(float));
  L.off_gmin = take((size_t)L.plan.num_g_tiles * (kTileG / 8) * sizeof(float));
  L.off_qsq = take(nq * sizeof(float));
  const size_t cand = lists * L.plan.cap * kTileQ;
  L.off_cand_val = take(cand * sizeof(float));
  L.off_cand_idx = take(cand * sizeof(int32_t));
  L.off_flags = take(nq * sizeof(int32_t));
  L.off_uncert = take(sizeof(int32_t));
  L.off_shared_thr = take((size_t)L.plan.q_tile_stride * kTileQ * sizeof(int32_t));
  L.off_row_max = take(lists * kTileQ * sizeof(float));
  L.off_row_maxpos = take(lists * kTileQ * sizeof(int32_t));
  L.sched_bytes = 256 + parts * L.plan.q_tile_stride * sizeof(int32_t);
  L.off_sched = take(L.sched_bytes);
  if (L.sel_bf16) {
    const size_t sq = (size_t)L.sub_q;
    L.off_tier_gates = take(256);                       // gate_sub, gate_full, fq_count
    L.off_fq = take(sq * sizeof(int32_t));
    L.off_qsub = take(sq * (size_t)dim * sizeof(float));
    L.off_qsq_sub = take(sq * sizeof(float));
    L.off_sub_dist = take(sq * (size_t)k * sizeof(float));
    L.off_sub_index = take(sq * (size_t)k * sizeof(int64_t));
    L.off_sub_flags = take(sq * sizeof(int32_t));
    const size_t sub_lists = lists_of(L.plan_sub);
    L.off_sub_cand_val = take(sub_lists * L.plan_sub.cap * kTileQ * sizeof(float));
    L.off_sub_cand_idx = take(sub_lists * L.plan_sub.cap * kTileQ * sizeof(int32_t));
    L.off_sub_row_max = take(sub_lists * kTileQ * sizeof(float));
    L.off_sub_row_maxpos = take(sub_lists * kTileQ * sizeof(int32_t));
    L.sub_sched_bytes = 256 + (size_t)L.plan_sub.num_splits * L.plan_sub.q_tile_stride * sizeof(int32_t);
    L.off_sub_sched = take(L.sub_sched_bytes);
    L.off_sub_thr = take((size_t)L.plan_sub.q_tile_stride * kTileQ * sizeof(int32_t));
    L.off_qb = take((size_t)num_q * dim * 2);
    L.off_gb = take((size_t)num_g * dim * 2);
    L.off_qres = take(nq * sizeof(float));
    L.off_gres = L.off_gmax + 16;  // the residual maxima share gmax's 256-byte block (cleared together)
  }
  if (L.padded_rows) {
    L.off_qpad = take((size_t)num_q * L.kdim * elem_size(dtype));
    L.off_gpad = take((size_t)num_g * L.kdim * elem_size(dtype));
  }
  if (L.precise) {
    L.off_gate = take(256);
    L.off_q3 = take((size_t)num_q * L.kdim * 3 * sizeof(float));
    L.off_g3 = take((size_t)num_g * L.kdim * 3 * sizeof(float));
    L.off_mu = take((size_t)L.kdim * sizeof(float));
    L.off_colpart = take(col_mean_workspace_bytes(L.kdim));
  }
  if (want_rank) {
    L.off_pos_dist = take(nq * sizeof(double));
    L.off_lo = take(nq * sizeof(float));
    L.off_hi = take(nq * sizeof(float));
    L.off_cnt = take(nq * sizeof(int32_t));
    L.off_dropped = take(nq * sizeof(int32_t));
    L.off_pool_count = take(sizeof(uint32_t));
    const size_t cap = nq * kUncertainPerQuery < 65536 ? 65536 : nq * kUncertainPerQuery;
    L.pool_cap = (uint32_t)(cap > 0x7fffffffu ? 0x7fffffffu : cap);
    L.off_pool_q = take((size_t)L.pool_cap * sizeof(int32_t));
    L.off_pool_idx = take((size_t)L.pool_cap * sizeof(int32_t));
  }
  L.total = o;
  return L;
}

}  // namespace

K1Plan topk_primary_plan(int64_t num_q, int64_t num_g, int64_t dim, int k, int dtype, int num_sms) {
  const bool sel = select_on_bf16(num_q, num_g, dim, k, dtype);
  return make_k1_plan(num_q, num_g, tile_dim(dim, dtype), k, sel ? SBIR_BF16 : dtype, num_sms > 0 ? num_sms : num_sms_cached(),
                      dtype == SBIR_F32 ? 16 : 6);
}

// ---- one retrieval pass in three phases (kernels.h: TopkPass) ----
// begin: argument checks, workspace carve-up, counters zeroed, query norms.
// feed:  gallery rows [fed, row_end) are resident — their norms, then K1 over their chunk steps
//        (the candidate lists carry over from feed to feed like they do from chunk to chunk).
// finish: exact re-scoring + certificate, rank resolution, the device-gated escalation pass
//        (fp32) and the brute-force fallbacks.
// sbir_pairwise_topk / _shard run begin, ONE feed of the whole gallery, finish;
// sbir_retrieve_host (host_path.cu) feeds the gallery as its chunks arrive over PCIe.
int topk_pass_begin(TopkPass& P, const void* q, int64_t num_q, const void* g, const float* g_sqnorm, int64_t num_g, int64_t dim, int dtype,
                    int metric, int k, int64_t index_offset, const int64_t* pos_index, const double* pos_dist_in,
                    const int64_t* pos_tie, int64_t tie_offset, float* out_dist, int64_t* out_index,
                    int64_t* out_rank, int64_t missing_rank, int32_t* out_uncertified, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  P = TopkPass{};
  if (!dtype_ok(dtype) || !metric_ok(metric)) return SBIR_ERR_INVALID_ARG;
  if (num_q < 0 || num_g < 0 || dim <= 0 || k <= 0) return SBIR_ERR_INVALID_ARG;
  if (k > kMaxK) return SBIR_ERR_UNSUPPORTED;
  if (num_q > INT32_MAX / 2 || num_g > INT32_MAX / 2 || dim > (1 << 20)) return SBIR_ERR_UNSUPPORTED;
  P.done = true;
  if (num_q == 0) return SBIR_OK;
  if (q == nullptr || out_dist == nullptr || out_index == nullptr) return SBIR_ERR_INVALID_ARG;
  if (num_g > 0 && g == nullptr) return SBIR_ERR_INVALID_ARG;
  const bool want_rank = out_rank != nullptr && (pos_index != nullptr || pos_dist_in != nullptr);
  if (out_rank != nullptr && !want_rank) return SBIR_ERR_INVALID_ARG;
  // rows of any width are accepted (odd widths go through zero-padded tile copies); only when the rows themselves
  // are 16-byte multiples must the base pointers be 16-byte aligned too (TMA reads them in place)
  if ((dim * (int64_t)elem_size(dtype)) % 16 == 0 && (reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(g)) % 16 != 0)
    return SBIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(g)) % elem_size(dtype) != 0) return SBIR_ERR_UNSUPPORTED;

  P.L = topk_layout(num_q, num_g, dim, k, dtype, want_rank ? 1 : 0);
  const TopkLayout& L = P.L;
  if (workspace == nullptr || reinterpret_cast<uintptr_t>(workspace) % 256 != 0 || workspace_bytes < L.total)
    return SBIR_ERR_WORKSPACE;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  P.ws = ws; P.st = st; P.want_rank = want_rank;
  P.q = q; P.g = g; P.g_sqnorm = g_sqnorm; P.num_q = num_q; P.num_g = num_g; P.dim = dim; P.dtype = dtype; P.metric = metric;
  P.gvec = reinterpret_cast<float*>(ws + L.off_gvec);
  P.gmax = reinterpret_cast<float*>(ws + L.off_gmax);
  P.gmin = reinterpret_cast<float*>(ws + L.off_gmin);
  P.qsq = reinterpret_cast<float*>(ws + L.off_qsq);
  P.flags = reinterpret_cast<int32_t*>(ws + L.off_flags);
  P.uncert = out_uncertified ? out_uncertified : reinterpret_cast<int32_t*>(ws + L.off_uncert);
  P.padded = (int64_t)L.plan.num_g_tiles * kTileG;
  float* cand_val = reinterpret_cast<float*>(ws + L.off_cand_val);
  int32_t* cand_idx = reinterpret_cast<int32_t*>(ws + L.off_cand_idx);
  // scalars the kernels accumulate into (external uncertified counter, running maxima): one 4-byte and one 256-byte memset
  SBIR_CUDA_TRY(cudaMemsetAsync(P.uncert, 0, sizeof(int32_t), st));

  if (num_g == 0) {
    // Empty gallery: no neighbours; every query's positive is missing.
    SBIR_TRY(launch_topk_merge(nullptr, nullptr, 0, num_q, k, out_dist, out_index, st));
    if (want_rank) SBIR_TRY(launch_fill_i64(out_rank, num_q, missing_rank, st));
    return SBIR_OK;
  }
  P.done = false;

  RankArgs& ra = P.ra;
  if (want_rank) {
    ra.q = q; ra.g = g; ra.num_q = num_q; ra.num_g = num_g; ra.dim = dim;
    ra.dtype = dtype; ra.metric = metric;
    ra.pos_index = pos_index; ra.pos_dist_in = pos_dist_in;
    ra.pos_tie = pos_tie; ra.tie_offset = tie_offset;
    ra.qsq = P.qsq; ra.gsq_max = P.gmax;
    ra.pos_dist = reinterpret_cast<double*>(ws + L.off_pos_dist);
    ra.rank_lo = reinterpret_cast<float*>(ws + L.off_lo);
    ra.rank_hi = reinterpret_cast<float*>(ws + L.off_hi);
    ra.cnt_less = reinterpret_cast<int32_t*>(ws + L.off_cnt);
    ra.dropped = reinterpret_cast<int32_t*>(ws + L.off_dropped);
    ra.pool_count = reinterpret_cast<uint32_t*>(ws + L.off_pool_count);
    ra.pool_cap = L.pool_cap;
    ra.pool_q = reinterpret_cast<int32_t*>(ws + L.off_pool_q);
    ra.pool_idx = reinterpret_cast<int32_t*>(ws + L.off_pool_idx);
    ra.out_rank = out_rank;
    ra.missing_rank = missing_rank;
  }
  K1Args& ka = P.ka;
  ka.num_q = num_q; ka.num_g = num_g;
  ka.dtype = dtype; ka.metric = metric;
  ka.mode = want_rank ? kModeTopkRank : kModeTopk;
  ka.gvec = P.gvec;
  ka.gmin = P.gmin;
  ka.cand_val = cand_val; ka.cand_idx = cand_idx;
  ka.row_max = reinterpret_cast<float*>(ws + L.off_row_max);
  ka.row_maxpos = reinterpret_cast<int32_t*>(ws + L.off_row_maxpos);
  ka.unit_counter = reinterpret_cast<uint32_t*>(ws + L.off_sched);
  ka.chunk_done = reinterpret_cast<int32_t*>(ws + L.off_sched + 256);
  ka.rank_lo = ra.rank_lo; ka.rank_hi = ra.rank_hi;
  ka.cnt_less = ra.cnt_less;
  ka.pool_count = ra.pool_count; ka.pool_cap = ra.pool_cap; ka.pool_q = ra.pool_q; ka.pool_idx = ra.pool_idx;
  ka.dropped = ra.dropped;
  ka.shared_thr = reinterpret_cast<int32_t*>(ws + L.off_shared_thr);
  FinalizeArgs& fa = P.fa;
  fa.q = q; fa.g = g; fa.num_q = num_q; fa.num_g = num_g; fa.dim = dim;
  fa.dtype = dtype; fa.metric = metric; fa.k = k; fa.index_offset = index_offset;
  fa.cand_val = cand_val; fa.cand_idx = cand_idx;
  fa.qsq = P.qsq; fa.gsq_max = P.gmax;
  fa.out_dist = out_dist; fa.out_index = out_index;
  fa.uncertified = P.uncert; fa.flags = P.flags;

  // Pass 1 runs the tensor-core tiles on the embeddings as they are (bf16 -> kind::f16; fp32 -> kind::tf32), or —
  // fp32 embeddings, normally — on their bf16-rounded copies (kind::f16 at twice the rate; L.sel_bf16).
  ka.gate = nullptr; ka.dim = L.kdim;
  ra.gate = nullptr; fa.gate = nullptr;
  P.kq = q; P.kg = g;  // what the tensor-core tiles (and the escalation pass's operand split) read
  if (L.padded_rows) {
    P.kq = ws + L.off_qpad; P.kg = ws + L.off_gpad;
    SBIR_TRY(launch_pad_rows(q, num_q, dim, dtype, ws + L.off_qpad, L.kdim, st));
  }
  SBIR_CUDA_TRY(cudaMemsetAsync(ws + L.off_gmax, 0, 256, st));  // gmax and, in the same 256-byte block, the bf16 residual maxima
  SBIR_TRY(launch_pass_reset(want_rank ? ra.cnt_less : nullptr, ra.dropped, num_q, want_rank ? ra.pool_count : nullptr,
                             ws + L.off_sched, L.sched_bytes, reinterpret_cast<int32_t*>(ws + L.off_shared_thr),
                             (int64_t)L.plan.q_tile_stride * kTileQ, nullptr, st));
  if (L.sel_bf16) {
    float* qres = reinterpret_cast<float*>(ws + L.off_qres);
    float* gres = reinterpret_cast<float*>(ws + L.off_gres);
    const float kappa = k1_accum_kappa(dim, 2);  // bf16 products are exact in the fp32 accumulator
    ra.kappa = kappa; fa.kappa = kappa;
    ra.q_res = qres; ra.g_res = gres; fa.q_res = qres; fa.g_res = gres;
    ka.q = ws + L.off_qb; ka.g = ws + L.off_gb; ka.dtype = SBIR_BF16;
    SBIR_TRY(launch_convert_bf16_norm(static_cast<const float*>(q), num_q, num_q, dim, ws + L.off_qb, 0, 0.f, P.qsq, nullptr, qres,
                                      nullptr, st));
  } else {
    const float kappa = k1_kappa(dtype, L.kdim);
    ra.kappa = kappa; fa.kappa = kappa;
    ka.q = P.kq; ka.g = P.kg;
    SBIR_TRY(launch_row_norm(q, num_q, num_q, dim, dtype, 0, 0.f, P.qsq, nullptr, st));
  }
  return SBIR_OK;
}

int64_t topk_pass_feed_granule(const TopkPass& P) {
  // intermediate feeds must end on a chunk-step boundary of the single gallery partition
  if (P.done || P.L.plan.num_splits != 1) return 0;
  return (int64_t)P.L.plan.tiles_per_chunk * kTileG;
}

int topk_pass_feed(TopkPass& P, int64_t row_end) {
  if (P.done) return SBIR_OK;
  const TopkLayout& L = P.L;
  cudaStream_t st = P.st;
  const int64_t row0 = P.fed_rows;
  if (row_end <= row0 || row_end > P.num_g) return SBIR_ERR_INVALID_ARG;
  const bool last = row_end == P.num_g;
  const int64_t granule = (int64_t)L.plan.tiles_per_chunk * kTileG;
  if (!(row0 == 0 && last)) {
    if (L.plan.num_splits != 1 || row0 % granule != 0 || (!last && row_end % granule != 0)) return SBIR_ERR_INVALID_ARG;
  }
  const size_t row_bytes = (size_t)P.dim * elem_size(P.dtype);
  const int64_t pad_end = last ? P.padded : row_end;  // the padding rows of the last tile belong to the last feed
  if (L.padded_rows)
    SBIR_TRY(launch_pad_rows(static_cast<const uint8_t*>(P.g) + (size_t)row0 * row_bytes, row_end - row0, P.dim, P.dtype,
                             P.ws + L.off_gpad + (size_t)row0 * L.kdim * elem_size(P.dtype), L.kdim, st));
  const int vec_mode = P.metric == SBIR_EUCLIDEAN ? 0 : 1;
  const float vec_pad = P.metric == SBIR_EUCLIDEAN ? INFINITY : nanf("");
  if (L.sel_bf16)  // fp32 rows -> bf16 selection operands + exact norms + rounding-residual maxima, one pass over the rows
    SBIR_TRY(launch_convert_bf16_norm(static_cast<const float*>(P.g) + (size_t)row0 * P.dim, row_end - row0, pad_end - row0, P.dim,
                                      P.ws + L.off_gb + (size_t)row0 * P.dim * 2, vec_mode, vec_pad, P.gvec + row0, P.gmax, nullptr,
                                      reinterpret_cast<float*>(P.ws + L.off_gres), st));
  else if (P.g_sqnorm != nullptr)  // gallery built by sbir_gallery_append / reloaded with its sidecar: N floats instead of N·dim elements
    SBIR_TRY(launch_gvec_from_sqnorm(P.g_sqnorm + row0, row_end - row0, pad_end - row0, vec_mode, vec_pad, P.gvec + row0, P.gmax,
                                     st, /*accumulate_max=*/true));
  else
    SBIR_TRY(launch_row_norm(static_cast<const uint8_t*>(P.g) + (size_t)row0 * row_bytes, row_end - row0, pad_end - row0, P.dim,
                             P.dtype, vec_mode, vec_pad, P.gvec + row0, P.gmax, st, /*accumulate_max=*/true));
  SBIR_TRY(launch_chunk_min(P.gvec + row0, (pad_end - row0) / 8, P.gmin + row0 / 8, st));
  // the rank band uses the largest gallery norm seen so far (it only widens from feed to feed)
  if (P.want_rank) SBIR_TRY(launch_rank_band(P.ra, st));
  if (row0 != 0) SBIR_CUDA_TRY(cudaMemsetAsync(P.ws + L.off_sched, 0, 256, st));  // unit counter of this launch (begin zeroed the first one)
  P.ka.chunk_begin = (row0 == 0) ? 0 : (int)(row0 / granule);
  P.ka.chunk_end = last ? 0 : (int)(row_end / granule);
  SBIR_TRY(launch_k1(P.ka, L.plan, st));
  P.fed_rows = row_end;
  return SBIR_OK;
}

int topk_pass_finish(TopkPass& P) {
  if (P.done) return SBIR_OK;
  if (P.fed_rows != P.num_g) return SBIR_ERR_INVALID_ARG;
  const TopkLayout& L = P.L;
  cudaStream_t st = P.st;
  uint8_t* ws = P.ws;
  RankArgs& ra = P.ra;
  K1Args& ka = P.ka;
  FinalizeArgs& fa = P.fa;
  const bool want_rank = P.want_rank;
  const int64_t num_q = P.num_q, num_g = P.num_g, dim = P.dim;
  SBIR_TRY(launch_finalize_topk(fa, L.plan, st));
  if (want_rank) SBIR_TRY(launch_rank_resolve(ra, st));

  // One whole-gallery scoring pass: (rank band) -> K1 -> finalize (+ rank pool resolution).  `gate`
  // makes every kernel of the pass a no-op unless the device flag is set.
  auto run_pass = [&](const void* kq, const void* kg, int64_t kdim, const K1Plan& plan, float kappa,
                      const int32_t* gate) -> int {
    ra.kappa = kappa; ra.gate = gate;
    fa.kappa = kappa; fa.gate = gate;
    ra.q_res = nullptr; ra.g_res = nullptr; fa.q_res = nullptr; fa.g_res = nullptr;  // no bf16 rounding in this pass
    ka.q = kq; ka.g = kg; ka.dim = kdim; ka.dtype = SBIR_F32; ka.gate = gate;
    ka.chunk_begin = 0; ka.chunk_end = 0;
    SBIR_TRY(launch_pass_reset(want_rank ? ra.cnt_less : nullptr, ra.dropped, num_q, want_rank ? ra.pool_count : nullptr,
                               ws + L.off_sched, L.sched_bytes, ka.shared_thr, (int64_t)L.plan.q_tile_stride * kTileQ, gate, st));
    if (want_rank) SBIR_TRY(launch_rank_band(ra, st));
    SBIR_TRY(launch_k1(ka, plan, st));
    SBIR_TRY(launch_finalize_topk(fa, plan, st));
    if (want_rank) SBIR_TRY(launch_rank_resolve(ra, st));
    return SBIR_OK;
  };

  // Tiers behind the bf16 selection of fp32 embeddings (device-gated, no host synchronisation): its rounding band is
  // twice kind::tf32's, so a query may fail the certificate here that kind::tf32 tiles would certify.
  if (L.sel_bf16) {
    int32_t* gates = reinterpret_cast<int32_t*>(ws + L.off_tier_gates);
    int32_t* gate_sub = gates, *gate_full = gates + 1, *fq_count = gates + 2;
    int32_t* fq = reinterpret_cast<int32_t*>(ws + L.off_fq);
    float* q_sub = reinterpret_cast<float*>(ws + L.off_qsub);
    float* qsq_sub = reinterpret_cast<float*>(ws + L.off_qsq_sub);
    float* sub_dist = reinterpret_cast<float*>(ws + L.off_sub_dist);
    int64_t* sub_index = reinterpret_cast<int64_t*>(ws + L.off_sub_index);
    int32_t* sub_flags = reinterpret_cast<int32_t*>(ws + L.off_sub_flags);
    int64_t max_bad = num_q / 50;
    if (max_bad < 4) max_bad = 4;
    SBIR_TRY(launch_tier_decide(P.flags, want_rank ? ra.dropped : nullptr, num_q, max_bad, fq, fq_count, gate_sub, gate_full, P.uncert, st));
    // (a) a few certificate failures: those queries alone, on kind::tf32 tiles against the whole gallery
    SBIR_TRY(launch_gather_sub(static_cast<const float*>(P.q), dim, fq, fq_count, L.sub_q, q_sub, P.qsq, qsq_sub, gate_sub, st));
    K1Args ks{};
    ks.q = q_sub; ks.g = P.kg; ks.num_q = L.sub_q; ks.num_g = num_g; ks.dim = L.kdim;
    ks.dtype = SBIR_F32; ks.metric = P.metric; ks.mode = kModeTopk;
    ks.gvec = P.gvec; ks.gmin = P.gmin; ks.gate = gate_sub;
    ks.cand_val = reinterpret_cast<float*>(ws + L.off_sub_cand_val);
    ks.cand_idx = reinterpret_cast<int32_t*>(ws + L.off_sub_cand_idx);
    ks.row_max = reinterpret_cast<float*>(ws + L.off_sub_row_max);
    ks.row_maxpos = reinterpret_cast<int32_t*>(ws + L.off_sub_row_maxpos);
    ks.unit_counter = reinterpret_cast<uint32_t*>(ws + L.off_sub_sched);
    ks.chunk_done = reinterpret_cast<int32_t*>(ws + L.off_sub_sched + 256);
    ks.shared_thr = reinterpret_cast<int32_t*>(ws + L.off_sub_thr);
    SBIR_TRY(launch_pass_reset(nullptr, nullptr, 0, nullptr, ws + L.off_sub_sched, L.sub_sched_bytes, ks.shared_thr,
                               (int64_t)L.plan_sub.q_tile_stride * kTileQ, gate_sub, st));
    SBIR_TRY(launch_k1(ks, L.plan_sub, st));
    FinalizeArgs fs = fa;
    fs.q = q_sub; fs.num_q = L.sub_q; fs.cand_val = ks.cand_val; fs.cand_idx = ks.cand_idx; fs.qsq = qsq_sub;
    fs.kappa = k1_kappa(SBIR_F32, L.kdim); fs.q_res = nullptr; fs.g_res = nullptr;
    fs.out_dist = sub_dist; fs.out_index = sub_index; fs.uncertified = nullptr; fs.flags = sub_flags; fs.gate = gate_sub;
    SBIR_TRY(launch_finalize_topk(fs, L.plan_sub, st));
    SBIR_TRY(launch_scatter_sub(fq, fq_count, L.sub_q, fa.k, sub_dist, sub_index, sub_flags, fa.out_dist, fa.out_index, P.flags,
                                P.uncert, gate_sub, st));
    // (b) more than that, or rank pools overflowed: one kind::tf32 pass over all queries
    SBIR_TRY(run_pass(P.kq, P.kg, L.kdim, L.plan_tf32, k1_kappa(SBIR_F32, L.kdim), gate_full));
  }

  // Pass 2 (fp32 only, device-gated): when more than 2 % of the queries could not be certified or
  // overflowed the rank pool — embeddings whose norms dwarf their distances, positives deep in an
  // unstructured distribution — the TF32 error band is the problem, so redo the pass with the
  // operands split into TF32 hi/lo parts concatenated along K ([qh|qh|ql]·[gh|gl|gh] = 3xTF32,
  // error ~2^-20): same kernel, 3x the MMA work, far cheaper than brute-forcing every query.
  if (L.precise) {
    const void* q = P.kq;      // zero-padded copies when the rows are not 16-byte multiples (kdim columns)
    const void* g = P.kg;
    const int64_t dim = L.kdim;
    const int metric = P.metric;
    float* gvec = P.gvec; float* gmax = P.gmax; float* gmin = P.gmin; float* qsq = P.qsq;
    const int64_t padded = P.padded;
    int32_t* gate = reinterpret_cast<int32_t*>(ws + L.off_gate);
    float* q3 = reinterpret_cast<float*>(ws + L.off_q3);
    float* g3 = reinterpret_cast<float*>(ws + L.off_g3);
    int64_t max_bad = num_q / 50;
    if (max_bad < 4) max_bad = 4;
    SBIR_TRY(launch_escalate_decide(P.flags, want_rank ? ra.dropped : nullptr, num_q, max_bad, gate, P.uncert, st));
    if (metric == SBIR_EUCLIDEAN && dim % 4 == 0) {
      // Euclidean distances are translation-invariant: centre both operands on the gallery's column
      // mean before the split, so the error band scales with the spread of the embeddings, not with
      // their norms (collapsed / post-ReLU features).  ‖q−µ‖², ‖g−µ‖² and their maximum replace the
      // raw norms for this pass (nothing after it reads them: the fallbacks are exact).
      float* mu = reinterpret_cast<float*>(ws + L.off_mu);
      float* colpart = reinterpret_cast<float*>(ws + L.off_colpart);
      SBIR_TRY(launch_col_mean(static_cast<const float*>(g), num_g, dim, colpart, mu, gmax, gate, st));
      SBIR_TRY(launch_center_split_tf32(static_cast<const float*>(q), num_q, num_q, dim, mu, 0, q3, qsq, 0.f, nullptr, gate, st));
      SBIR_TRY(launch_center_split_tf32(static_cast<const float*>(g), num_g, padded, dim, mu, 1, g3, gvec, INFINITY, gmax, gate, st));
      SBIR_TRY(launch_chunk_min(gvec, padded / 8, gmin, st));  // ungated: recomputes the same values when the pass is off
      SBIR_TRY(run_pass(q3, g3, 3 * dim, L.plan3, k1_kappa_centred(dim), gate));
    } else {
      SBIR_TRY(launch_split_tf32(static_cast<const float*>(q), num_q, dim, 0, q3, gate, st));
      SBIR_TRY(launch_split_tf32(static_cast<const float*>(g), num_g, dim, 1, g3, gate, st));
      SBIR_TRY(run_pass(q3, g3, 3 * dim, L.plan3, k1_kappa_precise(dim), gate));
    }
  }
  // Whatever is still unresolved is recomputed exactly by brute force.
  fa.gate = nullptr;
  ra.gate = nullptr;
  SBIR_TRY(launch_topk_fallback(fa, st));
  if (want_rank) SBIR_TRY(launch_rank_output(ra, st));
  P.done = true;
  return SBIR_OK;
}

namespace {

// Shared implementation of sbir_pairwise_topk (pos_index given, full rank) and
// sbir_pairwise_topk_shard (pos_dist given, local count).
int topk_impl(const void* q, int64_t num_q, const void* g, const float* g_sqnorm, int64_t num_g, int64_t dim, int dtype,
              int metric, int k, int64_t index_offset, const int64_t* pos_index,
              const double* pos_dist_in, const int64_t* pos_tie, int64_t tie_offset, float* out_dist, int64_t* out_index, int64_t* out_rank,
              int64_t missing_rank, int32_t* out_uncertified, void* workspace, size_t workspace_bytes,
              cudaStream_t st) {
  TopkPass P;
  SBIR_TRY(topk_pass_begin(P, q, num_q, g, g_sqnorm, num_g, dim, dtype, metric, k, index_offset, pos_index, pos_dist_in, pos_tie,
                           tie_offset, out_dist, out_index, out_rank, missing_rank, out_uncertified, workspace,
                           workspace_bytes, st));
  if (P.done) return SBIR_OK;
  SBIR_TRY(topk_pass_feed(P, num_g));
  return topk_pass_finish(P);
}

}  // namespace
}  // namespace sbir

using namespace sbir;

extern "C" {

int sbir_abi_version(void) { return SBIR_B200_ABI_VERSION; }

const char* sbir_status_string(int status) {
  switch (status) {
    case SBIR_OK: return "ok";
    case SBIR_ERR_INVALID_ARG: return "invalid argument";
    case SBIR_ERR_UNSUPPORTED: return "unsupported shape, alignment or k";
    case SBIR_ERR_CUDA: return "CUDA call failed (see sbir_last_cuda_error)";
    case SBIR_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
    case SBIR_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    default: return "unknown status";
  }
}

int sbir_last_cuda_error(void) { return g_last_cuda_error; }

int sbir_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int sbir_profile_enable(int on) {
  g_profile_on.store(on ? 1 : 0);
  return SBIR_OK;
}

int sbir_profile_collect(double* k1_ms_sum, int64_t* k1_launches, int64_t* kernel_launches) {
  std::lock_guard<std::mutex> lock(g_profile_mu);
  double sum = 0.0;
  int64_t n = 0;
  for (auto& ev : g_k1_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
      sum += ms;
      ++n;
    }
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  g_k1_events.clear();
  if (k1_ms_sum) *k1_ms_sum = sum;
  if (k1_launches) *k1_launches = n;
  if (kernel_launches) *kernel_launches = g_kernel_launches.exchange(0);
  return SBIR_OK;
}

int sbir_l2_normalize(const void* x, void* y, int64_t rows, int64_t dim, int dtype, float eps, void* stream) {
  if (!dtype_ok(dtype) || rows < 0 || dim <= 0) return SBIR_ERR_INVALID_ARG;
  if (rows == 0) return SBIR_OK;
  if (x == nullptr || y == nullptr) return SBIR_ERR_INVALID_ARG;
  return launch_l2_normalize(x, y, rows, dim, dtype, eps, static_cast<cudaStream_t>(stream));
}

int sbir_gallery_append(const void* block, int block_dtype, int64_t rows, int64_t dim, void* gallery, int gallery_dtype,
                        int64_t gallery_rows, int64_t row0, float* gallery_sqnorm, int normalize, void* stream) {
  if (!dtype_ok(block_dtype) || !dtype_ok(gallery_dtype) || rows < 0 || dim <= 0 || row0 < 0 || gallery_rows < 0)
    return SBIR_ERR_INVALID_ARG;
  if (row0 + rows > gallery_rows) return SBIR_ERR_INVALID_ARG;  // the block must fit the preallocated matrix
  if (rows == 0) return SBIR_OK;
  if (block == nullptr || gallery == nullptr) return SBIR_ERR_INVALID_ARG;
  uint8_t* dst = static_cast<uint8_t*>(gallery) + (size_t)row0 * (size_t)dim * elem_size(gallery_dtype);
  return launch_gallery_append(block, block_dtype, rows, dim, dst, gallery_dtype,
                               gallery_sqnorm ? gallery_sqnorm + row0 : nullptr, normalize ? 1 : 0, kCosineEps,
                               static_cast<cudaStream_t>(stream));
}

int sbir_row_sqnorm(const void* x, int64_t rows, int64_t dim, int dtype, float* out, void* stream) {
  if (!dtype_ok(dtype) || rows < 0 || dim <= 0) return SBIR_ERR_INVALID_ARG;
  if (rows == 0) return SBIR_OK;
  if (x == nullptr || out == nullptr) return SBIR_ERR_INVALID_ARG;
  return launch_row_norm(x, rows, rows, dim, dtype, 0, 0.f, out, nullptr, static_cast<cudaStream_t>(stream));
}

int sbir_pairwise_distance(const void* x1, int64_t rows1, const void* x2, int64_t rows2, int64_t dim,
                           int dtype, int metric, float* out, void* stream) {
  if (!dtype_ok(dtype) || !metric_ok(metric) || rows1 < 0 || rows2 < 0 || dim <= 0) return SBIR_ERR_INVALID_ARG;
  if (rows1 != rows2 && rows1 != 1 && rows2 != 1) return SBIR_ERR_INVALID_ARG;
  if (rows1 == 0 || rows2 == 0) return SBIR_OK;
  if (x1 == nullptr || x2 == nullptr || out == nullptr) return SBIR_ERR_INVALID_ARG;
  return launch_pairwise_distance(x1, rows1, x2, rows2, dim, dtype, metric, out, static_cast<cudaStream_t>(stream));
}

int sbir_pairwise_distance_bwd(const float* x1, int64_t rows1, const float* x2, int64_t rows2, int64_t dim,
                               int metric, const float* grad_out, float* grad_x1, float* grad_x2,
                               void* stream) {
  if (!metric_ok(metric) || rows1 < 0 || rows2 < 0 || dim <= 0) return SBIR_ERR_INVALID_ARG;
  if (rows1 != rows2 && rows1 != 1 && rows2 != 1) return SBIR_ERR_INVALID_ARG;
  if (rows1 == 0 || rows2 == 0) return SBIR_OK;
  if (x1 == nullptr || x2 == nullptr || grad_out == nullptr) return SBIR_ERR_INVALID_ARG;
  return launch_pairwise_distance_bwd(x1, rows1, x2, rows2, dim, metric, grad_out, grad_x1, grad_x2,
                                      static_cast<cudaStream_t>(stream));
}

size_t sbir_pairwise_topk_workspace_bytes(int64_t num_q, int64_t num_g, int64_t dim, int k, int dtype,
                                          int metric, int want_rank) {
  (void)metric;
  if (num_q < 0 || num_g < 0 || dim <= 0 || k <= 0 || k > kMaxK || !dtype_ok(dtype)) return 0;
  return topk_layout(num_q, num_g, dim, k, dtype, want_rank).total;
}

int sbir_pairwise_topk(const void* q, int64_t num_q, const void* g, const float* g_sqnorm, int64_t num_g, int64_t dim, int dtype,
                       int metric, int k, int64_t index_offset, const int64_t* pos_index, float* out_dist,
                       int64_t* out_index, int64_t* out_rank, int32_t* out_uncertified, void* workspace,
                       size_t workspace_bytes, void* stream) {
  return topk_impl(q, num_q, g, g_sqnorm, num_g, dim, dtype, metric, k, index_offset, pos_index, nullptr, pos_index, 0, out_dist,
                   out_index, out_rank, /*missing_rank=*/num_g, out_uncertified, workspace, workspace_bytes,
                   static_cast<cudaStream_t>(stream));
}

int sbir_positive_distance(const void* q, int64_t num_q, const void* g, int64_t num_g, int64_t dim, int dtype,
                           int metric, const int64_t* pos_index_local, double* out_pos_dist, void* stream) {
  if (!dtype_ok(dtype) || !metric_ok(metric) || num_q < 0 || num_g < 0 || dim <= 0) return SBIR_ERR_INVALID_ARG;
  if (num_q == 0) return SBIR_OK;
  if (q == nullptr || pos_index_local == nullptr || out_pos_dist == nullptr) return SBIR_ERR_INVALID_ARG;
  return launch_positive_distance(q, num_q, g, num_g, dim, dtype, metric, pos_index_local, out_pos_dist,
                                  static_cast<cudaStream_t>(stream));
}

int sbir_pairwise_topk_shard(const void* q, int64_t num_q, const void* g, const float* g_sqnorm, int64_t num_g, int64_t dim,
                             int dtype, int metric, int k, int64_t index_offset, const double* pos_dist,
                             const int64_t* pos_index_global, float* out_dist, int64_t* out_index, int64_t* out_count_less,
                             int32_t* out_uncertified, void* workspace, size_t workspace_bytes,
                             void* stream) {
  // A query without a positive anywhere (NaN pos_dist) contributes a local count of 0.
  return topk_impl(q, num_q, g, g_sqnorm, num_g, dim, dtype, metric, k, index_offset, nullptr, pos_dist, pos_index_global,
                   index_offset, out_dist,
                   out_index, out_count_less, /*missing_rank=*/0, out_uncertified, workspace, workspace_bytes,
                   static_cast<cudaStream_t>(stream));
}

int sbir_topk_merge(const float* dist, const int64_t* index, int num_lists, int64_t list_stride_dist, int64_t list_stride_index,
                    int64_t num_q, int k, float* out_dist, int64_t* out_index, void* stream) {
  if (num_lists < 0 || num_q < 0 || k <= 0 || list_stride_dist < 0 || list_stride_index < 0) return SBIR_ERR_INVALID_ARG;
  if ((list_stride_dist > 0 && list_stride_dist < num_q * k) || (list_stride_index > 0 && list_stride_index < num_q * k))
    return SBIR_ERR_INVALID_ARG;
  if (num_q == 0) return SBIR_OK;
  if (out_dist == nullptr || out_index == nullptr) return SBIR_ERR_INVALID_ARG;
  if (num_lists > 0 && (dist == nullptr || index == nullptr)) return SBIR_ERR_INVALID_ARG;
  return launch_topk_merge(dist, index, num_lists, num_q, k, out_dist, out_index, static_cast<cudaStream_t>(stream), list_stride_dist,
                           list_stride_index);
}

int sbir_retrieval_metrics(const int64_t* rank0, int64_t num_q, int k, double* out, void* stream) {
  if (num_q <= 0 || k <= 0 || k > 1024 || rank0 == nullptr || out == nullptr) return SBIR_ERR_INVALID_ARG;
  return launch_retrieval_metrics(rank0, num_q, k, out, static_cast<cudaStream_t>(stream));
}

int sbir_triplet_margin_loss(const float* a, const float* p, const float* n, int64_t batch, int64_t dim,
                             float margin, int metric, float* out_loss, float* out_per_row, float* grad_a,
                             float* grad_p, float* grad_n, void* stream) {
  if (!metric_ok(metric) || batch <= 0 || dim <= 0) return SBIR_ERR_INVALID_ARG;
  if (a == nullptr || p == nullptr || n == nullptr || out_loss == nullptr || out_per_row == nullptr)
    return SBIR_ERR_INVALID_ARG;
  return launch_triplet(a, p, n, batch, dim, margin, metric, out_loss, out_per_row, grad_a, grad_p, grad_n,
                        static_cast<cudaStream_t>(stream));
}

size_t sbir_batch_hard_workspace_bytes(int64_t batch, int64_t dim) {
  if (batch <= 0 || dim <= 0) return 0;
  return batch_hard_workspace_bytes(batch, dim);
}

int sbir_batch_hard_triplet_loss(const float* a, const float* p, const float* n, int64_t batch, int64_t dim,
                                 float margin, int metric, const int64_t* anchor_label,
                                 const int64_t* cand_label, float* out_loss, int64_t* out_hard_index,
                                 float* grad_a, float* grad_p, float* grad_n, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  if (!metric_ok(metric) || batch <= 0 || dim <= 0) return SBIR_ERR_INVALID_ARG;
  if (a == nullptr || p == nullptr || n == nullptr || out_loss == nullptr) return SBIR_ERR_INVALID_ARG;
  if ((anchor_label == nullptr) != (cand_label == nullptr)) return SBIR_ERR_INVALID_ARG;
  if ((dim * 4) % 16 != 0) return SBIR_ERR_UNSUPPORTED;
  return launch_batch_hard(a, p, n, batch, dim, margin, metric, anchor_label, cand_label, out_loss,
                           out_hard_index, grad_a, grad_p, grad_n, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

int sbir_debug_set_option(const char* name, int64_t value) {
  if (name == nullptr) return SBIR_ERR_INVALID_ARG;
  DebugOptions& o = g_debug_options;
  if (!std::strcmp(name, "k1_feed")) o.k1_feed = (int)value;
  else if (!std::strcmp(name, "k1_bands")) o.k1_bands = (int)value;
  else if (!std::strcmp(name, "k1_l2_hints")) o.k1_l2_hints = (int)value;
  else if (!std::strcmp(name, "k1_q_early")) o.k1_q_early = (int)value;
  else if (!std::strcmp(name, "k1_pair")) o.k1_pair = (int)value;
  else if (!std::strcmp(name, "k1_qres")) o.k1_qres = (int)value;
  else if (!std::strcmp(name, "k1_pair_coop")) o.k1_pair_coop = (int)value;
  else if (!std::strcmp(name, "k1_sel_bf16")) o.k1_sel_bf16 = (int)value;
  else if (!std::strcmp(name, "k1_chunk_mb")) o.k1_chunk_mb = (int)value;
  else if (!std::strcmp(name, "k1_flags")) o.k1_flags = (int)value;
  else if (!std::strcmp(name, "host_chunk_rows")) o.host_chunk_rows = value;
  else if (!std::strcmp(name, "watchdog_cycles")) o.watchdog_cycles = value;
  else if (!std::strcmp(name, "reset")) o = DebugOptions{};
  else return SBIR_ERR_INVALID_ARG;
  return SBIR_OK;
}

int sbir_debug_diag_build(void) {
#ifdef SBIR_DIAG
  return 1;
#else
  return 0;
#endif
}

int sbir_debug_plan(int64_t num_q, int64_t num_g, int64_t dim, int k, int dtype, int num_sms, int32_t* out) {
  if (num_q <= 0 || num_g <= 0 || dim <= 0 || k <= 0 || k > kMaxK || !dtype_ok(dtype) || num_sms <= 0 || out == nullptr)
    return SBIR_ERR_INVALID_ARG;
  const K1Plan p = topk_primary_plan(num_q, num_g, dim, k, dtype, num_sms);
  const int32_t v[13] = {p.cap, p.lists_per_row, p.num_q_tiles, p.num_g_tiles, p.num_splits, p.tiles_per_split,
                         p.num_chunks, p.tiles_per_chunk, p.num_units, p.part_fastest, p.pair, p.q_tile_stride,
                         (dtype == SBIR_BF16 || select_on_bf16(num_q, num_g, dim, k, dtype)) ? 1 : 0};
  for (int i = 0; i < 13; ++i) out[i] = v[i];
  return SBIR_OK;
}

int sbir_debug_k1_diag(uint64_t* out, int n) {
  if (out == nullptr || n <= 0) return SBIR_ERR_INVALID_ARG;
  return k1_diag_read(reinterpret_cast<unsigned long long*>(out), n);
}

size_t sbir_debug_dist_matrix_workspace_bytes(int64_t num_q, int64_t num_g, int64_t dim, int dtype) {
  if (num_q <= 0 || num_g <= 0 || dim <= 0 || !dtype_ok(dtype)) return 0;
  const K1Plan plan = make_k1_plan(num_q, num_g, dim, 1, dtype, num_sms_cached());
  return align_up((size_t)plan.num_g_tiles * kTileG * sizeof(float), 256) + 256;
}

int sbir_debug_dist_matrix(const void* q, int64_t num_q, const void* g, int64_t num_g, int64_t dim, int dtype,
                           int metric, float* out_e, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dtype_ok(dtype) || !metric_ok(metric) || num_q <= 0 || num_g <= 0 || dim <= 0) return SBIR_ERR_INVALID_ARG;
  if (q == nullptr || g == nullptr || out_e == nullptr) return SBIR_ERR_INVALID_ARG;
  const size_t need = sbir_debug_dist_matrix_workspace_bytes(num_q, num_g, dim, dtype);
  if (workspace == nullptr || workspace_bytes < need || reinterpret_cast<uintptr_t>(workspace) % 256 != 0)
    return SBIR_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const K1Plan plan = make_k1_plan(num_q, num_g, dim, 1, dtype, num_sms_cached());
  float* gvec = static_cast<float*>(workspace);
  uint32_t* counter = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(workspace) + need - 256);
  SBIR_CUDA_TRY(cudaMemsetAsync(counter, 0, 256, st));
  SBIR_TRY(launch_row_norm(g, num_g, (int64_t)plan.num_g_tiles * kTileG, dim, dtype,
                           metric == SBIR_EUCLIDEAN ? 0 : 1, metric == SBIR_EUCLIDEAN ? INFINITY : nanf(""),
                           gvec, nullptr, st));
  K1Args ka{};
  ka.q = q; ka.g = g; ka.num_q = num_q; ka.num_g = num_g; ka.dim = dim;
  ka.dtype = dtype; ka.metric = metric; ka.mode = kModeDump;
  ka.gvec = gvec; ka.dump = out_e; ka.unit_counter = counter;
  return launch_k1(ka, plan, st);
}

}  // extern "C"
