// dist_topk.cu — host side of K1 (the kernel lives in dist_topk_kernel.cuh): work decomposition
// (make_k1_plan), TMA tensor maps, launch parameters, and the hand-off to the translation unit that
// holds the instantiations for the input type and metric.
#include <cstdlib>

#include <cuda.h>

#include "common.cuh"
#include "dist_topk_params.h"
#include "kernels.h"

namespace sbir {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

}  // namespace

// [rows, dim] row-major matrix → 2-D tensor map with a (128-byte × box_rows) SWIZZLE_128B box.
int make_tmap(CUtensorMap* out, const void* base, int64_t rows, int64_t dim, int dtype, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return SBIR_ERR_CUDA;
  const size_t es = elem_size(dtype);
  const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)dim * es};
  const cuuint32_t box[2] = {(cuuint32_t)(kSwizzleBytes / es), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, dtype == SBIR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                         2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_cuda_error(1000 + (int)r);
    return SBIR_ERR_CUDA;
  }
  return SBIR_OK;
}

// Epilogue warps.  Small lists (cap <= 32): fp32 embeddings (kind::tf32) take 4; bf16 tiles complete
// 2-4× sooner, so two warps share each TMEM lane quarter, each with its own list.  Lists of 64/128
// entries make the epilogue the critical path at any width (one warp per scheduler issues one
// dependent instruction every ~4 cycles): 8 warps, the second of each quarter feeding the first
// (K1Config::kFeed).  Option k1_feed = 0 falls back to 4 warps (A/B runs).
static int epi_warps_for(int dtype, int cap) {
  // (kind::tf32 tiles with 8 warps / two lists per row were measured on the small shapes where the epilogue sets the time:
  // 1k x 10k x 2048 K1 180 vs 177 us — the cost there is the ~cap*ln(columns/cap) cold-list insertions per row and list,
  // which a second list per row does not reduce; profiles/r02_probe_cfg1_ab.log)
  if (cap <= 32) return dtype == SBIR_BF16 ? 8 : 4;
  return debug_options().k1_feed == 0 ? 4 : 8;
}

// Single-CTA tiles or CTA pairs (cta_group::2, M = 256: each CTA loads its own query tile and HALF of the
// gallery tile).  Pairs have the faster mainloop (20k × 1M × 512 bf16, epilogue off: 13.6 vs 14.6 ms) but
// one MMA then waits for the slowest of 16 epilogue warps, so with small lists single-CTA tiles win
// (cfg4: 816 vs 930 ms).  With the 128-entry lists of top-100 on fp32 rows only 3 single-CTA operand
// stages fit and the mainloop starves (operand wait 300 of 680 cycles per k-block); the pair's 32 KB
// stages fit 4: cfg3 K1 7.0 -> 6.2 ms.  bf16 tiles of rows of 4 KB and more (2048-d: fp32 embeddings selected on their
// bf16 copies) run the all-shared-memory form, which re-reads the 512 KB query tile for every gallery tile and is L2-bound
// (12.5k x 75k x 2048: 46 GB of operands in 3 ms = 15 TB/s); a pair moves a third less: K1 3.27 -> 2.97 ms, pass 4.03-4.10 ->
// 3.81-3.84 ms (top-10), 6.7-6.9 -> 6.65-6.75 ms (top-100); 1024-d rows: no difference (profiles/r02_probe_pair_wide_rows.log).
// Option k1_pair = 1 / 2 forces either form (A/B runs, tests).
static int k1_pair_for(int dtype, int cap, int64_t dim) {
  const int forced = debug_options().k1_pair;
  if (forced == 1 || forced == 2) return forced;
  if (dtype == SBIR_BF16 && dim * 2 >= 4096) return 2;
  return (dtype == SBIR_F32 && cap >= 64) ? 2 : 1;
}

int k1_diag_read(unsigned long long* out, int n) {
  // every instantiation unit keeps its own counters; only the one that ran has non-zero entries
  unsigned long long sum[148 * 8] = {};
  SBIR_TRY(k1_launch_f32_euclidean_diag(sum));
  SBIR_TRY(k1_launch_f32_cosine_diag(sum));
  SBIR_TRY(k1_launch_bf16_euclidean_diag(sum));
  SBIR_TRY(k1_launch_bf16_cosine_diag(sum));
  for (int i = 0; i < n && i < 148 * 8; ++i) out[i] = sum[i];
  return SBIR_OK;
}

K1Plan make_k1_plan(int64_t num_q, int64_t num_g, int64_t dim, int k, int dtype, int num_sms, int slack) {
  K1Plan p{};
  // capacity = k + slack; fp32 embeddings carry a wider error band (operand rounding), so they get more slack
  if (slack <= 0) slack = dtype == SBIR_BF16 ? 6 : 16;
  const int want = k + slack;
  p.cap = want <= 16 ? 16 : want <= 32 ? 32 : want <= 64 ? 64 : 128;
  if (k + 12 > 128) p.cap = 128;
  p.epi_warps = epi_warps_for(dtype, p.cap);
  p.lists_per_row = (p.epi_warps == 8 && p.cap <= 32) ? 2 : 1;
  p.num_q_tiles = (int)((num_q + kTileQ - 1) / kTileQ);
  p.num_g_tiles = (int)((num_g + kTileG - 1) / kTileG);
  if (p.num_q_tiles < 1) p.num_q_tiles = 1;
  if (p.num_g_tiles < 1) p.num_g_tiles = 1;
  p.q_tile_stride = (p.num_q_tiles + 1) & ~1;
  p.pair = (p.num_q_tiles >= 2 && num_sms >= 2) ? k1_pair_for(dtype, p.cap, dim) : 1;
  const int row_tiles = (p.num_q_tiles + p.pair - 1) / p.pair;  // rows of the unit grid
  const int workers = num_sms / p.pair;                           // CTAs or CTA pairs
  const size_t es = dtype == SBIR_BF16 ? 2 : 4;
  p.num_k_blocks = (int)((dim * es + kSwizzleBytes - 1) / kSwizzleBytes);

  // Partitions (independent candidate lists; finalize merges them): only as many as it takes to
  // give every chunk step a bit more than one wave of units (every extra partition repeats the
  // list warm-up, ~cap·ln(rows/cap) insertions per query), at most one per gallery tile and within
  // finalize's 4096 candidate entries per query.
  int64_t parts = (5LL * workers / 4 + row_tiles - 1) / row_tiles;
  const int64_t max_parts = 4096 / (p.cap * p.lists_per_row);
  if (parts > max_parts) parts = max_parts;
  if (parts > p.num_g_tiles) parts = p.num_g_tiles;
  if (parts < 1) parts = 1;
  // Small problems (a handful of tiles per worker, e.g. the reference's own 1k x 10k evaluation): the units are
  // so few that wave quantisation decides the time — 160 units of 2 tiles on 148 workers take as long as 4 tiles.
  // Pick the partition count that minimises (waves x tiles per unit), counting a unit's fixed cost (TMEM / barrier
  // set-up, list init and parking) as a fraction of a tile.
  if ((int64_t)row_tiles * p.num_g_tiles <= 16LL * workers) {
    const int64_t limit = max_parts < p.num_g_tiles ? max_parts : p.num_g_tiles;
    double best_cost = 1e30;
    for (int64_t c = 1; c <= limit; ++c) {
      const int64_t tps = (p.num_g_tiles + c - 1) / c;
      const int64_t ns = (p.num_g_tiles + tps - 1) / tps;
      const int64_t waves = (ns * row_tiles + workers - 1) / workers;
      const double cost = (double)waves * ((double)tps + 0.35) + 0.01 * (double)ns;  // ties: fewer partitions (less to merge)
      if (cost < best_cost - 1e-9) { best_cost = cost; parts = c; }
    }
  }
  p.tiles_per_split = (int)((p.num_g_tiles + parts - 1) / parts);
  p.num_splits = (p.num_g_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  // Chunks: ~12 MB of gallery rows per (partition, chunk) so that the rows every CTA of a chunk
  // step streams stay in L2 together with the live query tiles.
  const int64_t tile_bytes = (int64_t)kTileG * dim * (int64_t)es;
  // Resident-query form (dist_topk_kernel<..., kQRes>): bf16 rows of at most 1 KB — the query tile fits
  // 256 TMEM columns — on single-CTA tiles.  cfg4: 795 -> 765 ms (chunk 12 MB) -> 740 ms (48 MB).
  // Option k1_qres = 0 switches it off (A/B runs, tests).
  {
    p.qres = (debug_options().k1_qres != 0 && dtype == SBIR_BF16 && p.pair == 1 && p.epi_warps == 8 && dim * 2 <= 1024 &&
              dim % 8 == 0) ? 1 : 0;
  }
  // Chunk size: with the queries streamed through L2 as well, 12 MB keeps a chunk's gallery rows AND the
  // live query tiles resident; the resident-query form reads only gallery rows through L2, and longer
  // units mean fewer list hand-overs and query-tile loads (cfg4: 12 / 24 / 48 / 96 MB -> 787 / 765 / 740 / 736 ms).
  // 64/128-entry lists (top-100) make every list hand-over four to eight times as large: 96 MB chunks there
  // (cfg4 with K = 100, A/B in one process: 48 MB 830-835 ms, 96 MB 823-827 ms, 144 MB 821-826 ms; top-10: within noise).
  int64_t chunk_mb = p.qres ? (p.cap >= 64 ? 96 : 48) : 12;
  if (debug_options().k1_chunk_mb > 0) chunk_mb = debug_options().k1_chunk_mb;  // experiments / tests of the chunk hand-over
  int64_t tpc = (chunk_mb << 20) / (tile_bytes > 0 ? tile_bytes : 1);
  if (tpc < 1) tpc = 1;
  if (tpc > p.tiles_per_split) tpc = p.tiles_per_split;
  p.num_chunks = (int)((p.tiles_per_split + tpc - 1) / tpc);
  p.tiles_per_chunk = (p.tiles_per_split + p.num_chunks - 1) / p.num_chunks;
  // Wide rows: the query tiles of all concurrently running units must stay L2-resident (each is
  // re-read for every gallery tile); when one tile per worker would not (> 24 MB), run the
  // partitions of the same query tile next to each other instead.
  const int64_t q_tile_bytes = (int64_t)kTileQ * p.pair * dim * (int64_t)es;
  p.part_fastest = (p.num_splits > 1 && q_tile_bytes * workers > (24LL << 20)) ? 1 : 0;
  p.num_units = p.num_chunks * p.num_splits * row_tiles;
  // L2 bands (decode_unit, option k1_bands): walking the query tiles in bands that scan the whole gallery one after the
  // other was built to keep a band's queries L2-resident.  Measured on cfg4 (profiles/r02_probe_l2_bands.log): it is
  // 0.5-1.5 % SLOWER and reads MORE from DRAM (1 band 58 GB, 2 bands 80 GB, 3 bands 95 GB per launch) — the re-reads are
  // gallery-chunk lines evicted by the streaming traffic, not the queries, and every band repeats them.  Off unless forced.
  p.band_q = row_tiles;
  {
    const int forced = debug_options().k1_bands;
    if (forced > 1) p.band_q = (int)((row_tiles + forced - 1) / (forced < row_tiles ? forced : row_tiles));
  }
  return p;
}

int launch_k1(const K1Args& a, const K1Plan& plan, cudaStream_t st) {
  if ((a.dim * (int64_t)elem_size(a.dtype)) % 16 != 0) return SBIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.g)) % 16 != 0) return SBIR_ERR_UNSUPPORTED;
  if (a.num_q <= 0 || a.num_g <= 0) return SBIR_OK;
  const int pair = plan.pair == 2 ? 2 : 1;
  const bool select_mode = a.mode == kModeTopk || a.mode == kModeTopkRank;
  const bool qres = plan.qres == 1 && select_mode && pair == 1 && a.dtype == SBIR_BF16 && plan.epi_warps == 8;
  CUtensorMap tq, tg;
  SBIR_TRY(make_tmap(&tq, a.q, a.num_q, a.dim, a.dtype, kTileQ));
  // pair mode / resident-query form: gallery boxes of half a tile (128 rows)
  SBIR_TRY(make_tmap(&tg, a.g, a.num_g, a.dim, a.dtype, (pair == 2 || qres) ? kTileG / 2 : kTileG));
  K1Params prm{};
  prm.gvec = a.gvec;
  prm.gmin = a.gmin;
  prm.gate = a.gate;
  prm.num_q = (int)a.num_q;
  prm.num_g = (int)a.num_g;
  prm.num_q_tiles = plan.num_q_tiles;
  prm.num_g_tiles = plan.num_g_tiles;
  prm.num_k_blocks = plan.num_k_blocks;
  prm.num_row_tiles = (plan.num_q_tiles + pair - 1) / pair;
  prm.num_parts = plan.num_splits;
  prm.tiles_per_part = plan.tiles_per_split;
  prm.num_chunks = plan.num_chunks > 0 ? plan.num_chunks : 1;
  prm.tiles_per_chunk = plan.tiles_per_chunk > 0 ? plan.tiles_per_chunk : plan.tiles_per_split;
  prm.chunk_begin = a.chunk_begin;
  {
    // a launch may cover only the chunk steps [chunk_begin, chunk_end) of the plan (streamed gallery)
    const int c_end = a.chunk_end > 0 ? a.chunk_end : prm.num_chunks;
    const int per_step = plan.num_units / (plan.num_chunks > 0 ? plan.num_chunks : 1);
    prm.num_units = (c_end - a.chunk_begin) * per_step;
    if (prm.num_units <= 0) return SBIR_OK;
    prm.num_steps = c_end - a.chunk_begin;
    prm.band_rows = (plan.band_q > 0 && plan.band_q < prm.num_row_tiles) ? plan.band_q : prm.num_row_tiles;
  }
  prm.part_fastest = plan.part_fastest;
  prm.q_tile_stride = plan.q_tile_stride;
  prm.elems_per_kblock = (int)(kSwizzleBytes / elem_size(a.dtype));
  prm.q_raw = a.q;
  prm.dim_elems = (int)a.dim;
  prm.flags = debug_options().k1_flags;
  prm.watchdog_cycles = debug_options().watchdog_cycles;
  prm.pair_cooperative = debug_options().k1_pair_coop != 0 ? 1 : 0;
  prm.q_early = debug_options().k1_q_early != 0 ? 1 : 0;
  // L2 eviction hints of the resident-query form, bits: 1 gallery chunk evict_last, 2 query tiles evict_first, 4 parked
  // lists evict_first (option k1_l2_hints).  OFF by default: measured on cfg4 (profiles/r02_probe_l2_hints.log) they cut the
  // DRAM traffic of a launch from 64 GB to 46 GB (all three; the query-tile hint alone: 53 GB) but cost 0.6-1.2 % of time —
  // the query tiles that survive in L2 from one chunk step to the next are what keeps a unit's start short, and DRAM
  // runs at ~1 % of its bandwidth either way.  Time is the metric, so the traffic stays.
  prm.l2_hints = (qres && prm.num_chunks > 1 && debug_options().k1_l2_hints > 0) ? debug_options().k1_l2_hints : 0;
  prm.unit_counter = a.unit_counter;
  prm.chunk_done = a.chunk_done;
  prm.cand_val = a.cand_val;
  prm.cand_idx = a.cand_idx;
  prm.row_max = a.row_max;
  prm.row_maxpos = a.row_maxpos;
  prm.rank_lo = a.rank_lo;
  prm.rank_hi = a.rank_hi;
  prm.cnt_less = a.cnt_less;
  prm.pool_count = a.pool_count;
  prm.pool_cap = a.pool_cap;
  prm.pool_q = a.pool_q;
  prm.pool_idx = a.pool_idx;
  prm.dropped = a.dropped;
  prm.shared_thr = a.shared_thr;
  prm.dump = a.dump;
  if (pair == 1 && prm.unit_counter == nullptr) return SBIR_ERR_INVALID_ARG;
  const bool select = a.mode == kModeTopk || a.mode == kModeTopkRank;
  if (select && (prm.chunk_done == nullptr || prm.row_max == nullptr || prm.row_maxpos == nullptr || prm.gmin == nullptr))
    return SBIR_ERR_INVALID_ARG;

  int dev = 0, num_sms = 148;
  SBIR_CUDA_TRY(cudaGetDevice(&dev));
  SBIR_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const int cap = a.mode == kModeDump ? 16 : plan.cap;

  // dump launches carry no plan of their own: 4 warps for fp32, 8 for bf16
  const int epi = select ? plan.epi_warps : (a.dtype == SBIR_BF16 ? 8 : 4);
  if (a.dtype == SBIR_F32) {
    if (a.metric == SBIR_EUCLIDEAN) return k1_launch_f32_euclidean(epi, a.mode, cap, pair, qres, tq, tg, prm, num_sms, st);
    return k1_launch_f32_cosine(epi, a.mode, cap, pair, qres, tq, tg, prm, num_sms, st);
  }
  if (a.metric == SBIR_EUCLIDEAN) return k1_launch_bf16_euclidean(epi, a.mode, cap, pair, qres, tq, tg, prm, num_sms, st);
  return k1_launch_bf16_cosine(epi, a.mode, cap, pair, qres, tq, tg, prm, num_sms, st);
}

}  // namespace sbir
