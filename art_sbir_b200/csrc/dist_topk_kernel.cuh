// dist_topk_kernel.cuh — K1: pairwise sketch×artwork distance tiles on the 5th-gen tensor cores with
// the per-query selection fused into the epilogue (the distance matrix never reaches HBM).
//
// Replaces the reference's per-query `utils.euclidean_distance(q, G)` / `cosine_distance`
// followed by `distances.topk(...)` (inference.py:44-49, 62-65) for ALL queries at once.
//
// Structure (one persistent CTA per SM, warp-specialised):
//   warp 0      scheduler + TMA producer: claims work units from a global counter, publishes
//               them to the other warps through a small shared-memory ring, and streams the
//               operand k-slices (SWIZZLE_128B boxes of 128 bytes × 128/256 rows, mbarrier
//               complete_tx) through a 3-10-stage ring                     (UTMALDG in SASS)
//   warp 1      MMA issue: tcgen05.mma (kind::tf32 for fp32 embeddings, kind::f16 for bf16) into
//               one of two TMEM accumulators; tcgen05.commit releases smem stages and publishes
//               the accumulator                                        (UTCHMMA / UTCBAR)
//               Both loops run warp-uniform with one lane elected by elect.sync: the operands
//               stay in uniform registers and the instructions issue back to back.
//   warps 2..   epilogue: tcgen05.ld 32 lanes × 32 columns → e = ‖g‖² − 2·q·g (euclidean) or
//               e = −q·g/max(‖g‖,eps) (cosine); a thread owns one query row and keeps that
//               query's running best-`cap` list; only chunks whose minimum beats the row's
//               current threshold take the insertion path                             (LDTM)
// Operand forms (K1Config): all-smem — Q [128 × 128 B] + G [256 × 128 B] per stage, M128 × N256
// MMAs, two 256-column accumulators; resident-query (kQRes, bf16 rows <= 1 KB) — the unit's query
// tile lives in 256 TMEM columns and is the A operand from there, G half-tiles [128 × 128 B] per
// stage, M128 × N128 MMAs, two 128-column accumulators; CTA pairs (kPair = 2, cta_group::2, M = 256).
// e orders gallery rows exactly like the distance does for a fixed query (‖q‖² and the
// query norm are per-row constants); exact distances are recomputed for the survivors by
// finalize.cu, so tensor-core rounding never reaches the caller.
//
// Work decomposition (make_k1_plan).  The gallery is cut into `num_splits` PARTITIONS (scanned
// independently, each with its own candidate lists — parallelism when there are few query
// tiles) and every partition into CHUNKS of a few MB that are scanned one after the other; a
// unit is (query tile, partition, chunk).  Units are numbered chunk-major and handed out
// dynamically, so at any moment all CTAs work on the same one or two chunk steps: the chunk's
// gallery rows are read from HBM once and served to everyone else from L2 (with long units the
// CTAs drift apart and L2 sharing collapses — ncu measured 1.73 TB of DRAM reads per cfg4 pass,
// profiles/r01_ncu_k1_cfg4_full.txt).  The candidate list of a (query tile, partition) is
// carried from chunk to chunk through global memory, ordered by a completion counter.
//
// This header holds the kernel template and its dispatch; it is compiled four times, once per
// (input type, metric) pair, by dist_topk_{f32,bf16}_{euclidean,cosine}.cu (which define
// SBIR_K1_INST_TF32 / SBIR_K1_INST_METRIC / SBIR_K1_INST_NAME before including it), so that the
// ~120 instantiations build in parallel.  Planning and launch set-up live in dist_topk.cu.
#include <cstdlib>

#include <cuda.h>

#include "common.cuh"
#include "dist_topk_params.h"
#include "kernels.h"
#include "ptx.cuh"

// Diagnostic switches (A/B runs, cycle counters) exist only in -DSBIR_DIAG builds (SBIR_BUILD_DIAG=1 at build
// time); in the product build the flag word is the constant 0 and every diagnostic branch is compiled out.
#ifdef SBIR_DIAG
#define K1_DIAG_FLAGS(prm) ((prm).flags)
#else
#define K1_DIAG_FLAGS(prm) 0
#endif

namespace sbir {

namespace {

constexpr int kStageBytesQ = kTileQ * kSwizzleBytes;      // 16 KB
constexpr int kStageBytesGFull = kTileG * kSwizzleBytes;  // 32 KB (halved per CTA in pair mode)
constexpr int kSmemLimit = 227 * 1024;
constexpr int kTmemCols = 512;  // two 256-column fp32 accumulators
constexpr int kSchedDepth = 4;  // unit ring between the scheduler and the other warps

// Diagnostics (SBIR_K1_FLAGS & 64): cycles per CTA spent by the MMA issuer waiting for a free
// accumulator [0] and for operands [1], its whole loop [2], and by epilogue warp 0 waiting for a
// finished accumulator [3] and inside list insertions [4]; read with sbir_debug_k1_diag.
__device__ unsigned long long g_k1_diag[148 * 8];

// Monotone float <-> int32 map so a float minimum can be taken with an integer atomicMin.
__device__ __forceinline__ int32_t float_to_ordered_int(float f) {
  const int32_t b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_int_to_float(int32_t i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}
__device__ __forceinline__ int32_t ld_relaxed(const int32_t* p) {
  int32_t v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int32_t ld_acquire(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int32_t* p, int32_t v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 3-input max / min (FMNMX3 on sm_100a): a 32-value reduction in 16 instructions, depth 4.
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
template <bool kMax, typename F>
__device__ __forceinline__ float reduce32(F get) {
  auto op3 = [](float a, float b, float c) { return kMax ? fmax3(a, b, c) : fmin3(a, b, c); };
  float a[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = op3(get(3 * i), get(3 * i + 1), get(3 * i + 2));
  const float b0 = op3(a[0], a[1], a[2]), b1 = op3(a[3], a[4], a[5]), b2 = op3(a[6], a[7], a[8]);
  const float b3 = op3(a[9], get(30), get(31));
  const float c = op3(b0, b1, b2);
  return kMax ? fmaxf(c, b3) : fminf(c, b3);
}

// kPair = 1: one CTA computes a 128×256 tile (cta_group::1).  kPair = 2: a 2-CTA cluster computes
// a 256×256 tile with one M=256 tcgen05.mma (cta_group::2): each CTA loads its own 128 query rows
// and only HALF of the gallery tile.
// kQRes (bf16 rows of at most 1 KB, single-CTA tiles): the 128-row QUERY tile of a unit stays resident
// in tensor memory (256 of the 512 columns) and is the MMA's A operand from there (tcgen05.mma with
// A in TMEM); only gallery k-slices travel through the shared-memory ring.  The MMA rate of the
// all-smem form follows the operand bytes it reads from shared memory (12 KB per K step for a 128×256
// tile); with A in TMEM that is 8 KB per 256 gallery rows, the L2→SM traffic drops by a third and the
// ring gets deeper.  The accumulators shrink to two buffers of 128 columns: a 256-row gallery tile is
// computed as two half-tiles.
template <int kCap, int kEpiWarps, int kPair, bool kQRes = false>
struct K1Config {
  static constexpr int kAccCols = kQRes ? kTileG / 2 : kTileG;  // columns per accumulator buffer = MMA N
  static constexpr int kSubTiles = kTileG / kAccCols;           // accumulator-sized pieces per gallery tile
  static constexpr int kAccBase = kQRes ? 256 : 0;              // first accumulator column (Q tile below it)
  static constexpr int kStageBytesG = kQRes ? kAccCols * kSwizzleBytes : kStageBytesGFull / kPair;
  static constexpr int kStageBytesA = kQRes ? 0 : kStageBytesQ;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesG;
  static constexpr int kMaxStages = kQRes ? 10 : (kPair == 2 ? 6 : 4);
  static_assert(!kQRes || kPair == 1, "the resident-query form uses single-CTA tiles");
  // Two epilogue warps share every TMEM lane quarter when kEpiWarps == 8 and split the columns.
  // Small lists: each of them keeps its own list (two lists per row).  Lists of 64/128 entries do
  // not fit twice beside the operand ring: the first warp OWNS the row's single list and the
  // second one FEEDS it — it screens its columns against the owner's published threshold and
  // forwards the rare hits through a small per-row queue in shared memory.
  static constexpr int kColSplit = kEpiWarps / 4;
  static constexpr int kListsPerRow = (kEpiWarps == 8 && kCap <= 32) ? 2 : 1;
  static constexpr bool kFeed = kColSplit == 2 && kListsPerRow == 1;
  static constexpr int kFeedDepth = 4;  // queue entries per row
  // distance keys of the running lists always live in shared memory (they are re-scanned on
  // every insertion); the gallery indices are write-only inside the kernel and go straight to
  // the global candidate buffer when they do not fit beside the operand ring.
  static constexpr bool kIdxInSmem = kCap * kListsPerRow <= 64;
  static constexpr int kValBytes = kCap * kListsPerRow * kTileQ * 4;
  static constexpr bool kTwoLevel = kCap >= 64;  // per-group-of-8 maxima beside the keys
  static constexpr int kGroupBytes = kTwoLevel ? (kCap / 8) * kListsPerRow * kTileQ * 4 : 0;
  // feed region: queue values + indices [depth][128], tail / head / published threshold [128], 8 flags
  static constexpr int kFeedBytes = kFeed ? (2 * kFeedDepth + 3) * kTileQ * 4 + 32 : 0;
  static constexpr int kListBytes = kValBytes + (kIdxInSmem ? kValBytes : 0) + kGroupBytes + kFeedBytes;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kStagesFit = (kSmemLimit - 1024 - kListBytes - kBarrierBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > kMaxStages ? kMaxStages : kStagesFit;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kListBytes + kBarrierBytes;
  static constexpr int kThreads = 64 + kEpiWarps * 32;
  static constexpr int kColsPerWarp = kAccCols / kColSplit;
  static_assert(kStages >= 2, "operand ring needs at least two stages");
  static_assert(kEpiWarps == 4 || kEpiWarps == 8, "epilogue warps must cover the 4 TMEM lane quarters");
};


// Unit → (row of the unit grid, partition, chunk) and the gallery tiles it covers.  Chunk-major;
// inside a chunk step either the query row or the partition varies fastest (plan.part_fastest:
// wide fp32 rows keep fewer query tiles live in L2 when partitions of one query tile run together).
struct UnitCoord {
  int row_tile, part, chunk, t_begin, t_end;
};
__device__ __forceinline__ UnitCoord decode_unit(int unit, const K1Params& p) {
  // L2 bands: the rows of the unit grid are walked in bands of `band_rows` query tiles, and a band scans all
  // chunk steps of the launch before the next band starts, so that what is live at any time — the band's query
  // tiles and parked lists plus one or two gallery chunks — fits L2 (one band = the plain chunk-major order).
  const int band_units = p.band_rows * p.num_steps * p.num_parts;  // units of a full band
  const int band = unit / band_units;
  const int row0 = band * p.band_rows;
  const int rows = min(p.band_rows, p.num_row_tiles - row0);
  const int per_step = p.num_parts * rows;
  const int in_band = unit - band * band_units;
  UnitCoord c;
  const int step = in_band / per_step;
  c.chunk = p.chunk_begin + step;
  const int r = in_band - step * per_step;
  if (p.part_fastest) {
    c.row_tile = r / p.num_parts;
    c.part = r - c.row_tile * p.num_parts;
  } else {
    c.part = r / rows;
    c.row_tile = r - c.part * rows;
  }
  c.row_tile += row0;
  const int part_begin = c.part * p.tiles_per_part;
  const int part_end = min(part_begin + p.tiles_per_part, p.num_g_tiles);
  c.t_begin = part_begin + c.chunk * p.tiles_per_chunk;
  c.t_end = min(c.t_begin + p.tiles_per_chunk, part_end);
  if (c.t_begin > c.t_end) c.t_begin = c.t_end;  // empty unit (short last partition)
  return c;
}

template <bool kTF32, int kMetric, int kMode, int kCap, int kEpiWarps, int kPair, bool kQRes = false>
__global__ void __launch_bounds__(K1Config<kCap, kEpiWarps, kPair, kQRes>::kThreads, 1)
dist_topk_kernel(const __grid_constant__ CUtensorMap tmap_q,
                 const __grid_constant__ CUtensorMap tmap_g, const K1Params prm) {
  using Cfg = K1Config<kCap, kEpiWarps, kPair, kQRes>;
  static_assert(!kQRes || !kTF32, "the resident-query form is kind::f16 only");
  constexpr int kAccCols = Cfg::kAccCols;
  constexpr int kSubTiles = Cfg::kSubTiles;
  if (prm.gate != nullptr && *prm.gate == 0) return;  // uniform across the grid: nothing was set up yet
  const long long wd = prm.watchdog_cycles;
  // L2 policies of the resident-query form (prm.l2_hints): the gallery chunk every CTA streams during a chunk step is
  // kept (evict_last); query tiles on their way to TMEM and parked lists are touched once per step (evict_first).
  [[maybe_unused]] const uint64_t pol_keep = l2_policy_evict_last();
  [[maybe_unused]] const uint64_t pol_stream = l2_policy_evict_first();
  constexpr int kStages = Cfg::kStages;
  constexpr int kStageBytesG = Cfg::kStageBytesG;
  constexpr int kStageBytes = Cfg::kStageBytes;
  constexpr bool kSelect = (kMode == kModeTopk || kMode == kModeTopkRank);
  constexpr bool kRank = (kMode == kModeTopkRank);
  constexpr bool kDynamic = (kPair == 1);  // dynamic unit hand-out (pairs walk a static stride)
  // Pair mode: `cta_rank` 0 is the leader (issues the MMAs, owns the full/acc_empty barriers);
  // a unit's query "row" is then a PAIR of query tiles and this CTA works on query tile
  // 2·row + cta_rank and on gallery rows [rank·128, +128) of every 256-row tile.
  const int cta_rank = kPair == 2 ? (int)cluster_ctarank() : 0;
  const int worker = kPair == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = kPair == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_g = smem + kStages * Cfg::kStageBytesA;
  float* list_val_s = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  int32_t* list_idx_s = reinterpret_cast<int32_t*>(list_val_s + kCap * Cfg::kListsPerRow * kTileQ);
  float* list_grp_s = reinterpret_cast<float*>(smem + kStages * kStageBytes + Cfg::kListBytes - Cfg::kFeedBytes - Cfg::kGroupBytes);
  // feeder queue (Cfg::kFeed): SPSC ring per row; counters run on for the whole kernel
  float* fq_val = reinterpret_cast<float*>(smem + kStages * kStageBytes + Cfg::kListBytes - Cfg::kFeedBytes);  // [depth][128]
  int32_t* fq_idx = reinterpret_cast<int32_t*>(fq_val + Cfg::kFeedDepth * kTileQ);                                // [depth][128]
  volatile uint32_t* fq_tail = reinterpret_cast<volatile uint32_t*>(fq_idx + Cfg::kFeedDepth * kTileQ);           // [128] pushed
  volatile uint32_t* fq_head = fq_tail + kTileQ;                                                                  // [128] consumed
  volatile float* thr_pub = reinterpret_cast<volatile float*>(fq_head + kTileQ);                                  // [128] owner's threshold
  volatile int32_t* unit_ready = reinterpret_cast<volatile int32_t*>(thr_pub + kTileQ);                           // [4] per quarter
  volatile int32_t* feeder_done = unit_ready + 4;                                                                 // [4] per quarter
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + Cfg::kListBytes);
  uint64_t* full_bar = bars;                         // [kStages]
  uint64_t* empty_bar = bars + kStages;              // [kStages]
  uint64_t* acc_full_bar = bars + 2 * kStages;       // [2]
  uint64_t* acc_empty_bar = bars + 2 * kStages + 2;  // [2]
  uint64_t* sched_full_bar = bars + 2 * kStages + 4;                 // [kSchedDepth]
  uint64_t* sched_empty_bar = bars + 2 * kStages + 4 + kSchedDepth;  // [kSchedDepth]
  int32_t* sched_unit = reinterpret_cast<int32_t*>(bars + 2 * kStages + 4 + 2 * kSchedDepth);  // [kSchedDepth]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_unit + kSchedDepth);
  uint64_t* q_ready_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);  // resident-query form: Q tile of the unit is in TMEM

  const int warp = __shfl_sync(kFullMask, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_g);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full_bar[a], 1);
      mbar_init(&acc_empty_bar[a], kEpiWarps * kPair);  // epilogue warps of both CTAs of a pair
    }
    for (int s = 0; s < kSchedDepth; ++s) {
      mbar_init(&sched_full_bar[s], 1);
      mbar_init(&sched_empty_bar[s], 1 + kEpiWarps);  // MMA issuer + every epilogue warp
    }
    mbar_init(q_ready_bar, kEpiWarps);  // every epilogue warp stores its share of the query tile
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (kPair == 2) tmem_alloc_pair(tmem_slot, kTmemCols);
    else tmem_alloc(tmem_slot, kTmemCols);
  }
  if constexpr (Cfg::kFeed) {
    if (threadIdx.x < kTileQ) {
      fq_tail[threadIdx.x] = 0;
      fq_head[threadIdx.x] = 0;
      thr_pub[threadIdx.x] = INFINITY;
      if (threadIdx.x < 8) unit_ready[threadIdx.x] = -1;  // unit_ready[0..3], feeder_done[0..3]
    }
  }
  tc_fence_before();
  if constexpr (kPair == 2) cluster_sync_all();  // peer barriers initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Consumers (MMA issuer, epilogue warps) take their next unit from the scheduler ring.
  // Returns -1 when the work is exhausted.  `it` counts units taken by this role.
  auto next_unit_consumer = [&](int it) -> int {
    if constexpr (kDynamic) {
      const int slot = it % kSchedDepth;
      mbar_wait(&sched_full_bar[slot], (uint32_t)(it / kSchedDepth) & 1u, wd);
      return *reinterpret_cast<volatile int32_t*>(&sched_unit[slot]);
    } else {
      const int u = worker + it * num_workers;
      return u < prm.num_units ? u : -1;
    }
  };
  auto release_unit_slot = [&](int it) {
    if constexpr (kDynamic) mbar_arrive(&sched_empty_bar[it % kSchedDepth]);
  };

  if (warp == 0) {
    // ------------------------------------------------- scheduler + TMA producer ----
    // Like the MMA issuer: the whole warp runs the loop with warp-uniform state, lane 0 executes the
    // TMA / barrier instructions (their operands then live in uniform registers; with one divergent
    // thread every UTMALDG is wrapped in an R2UR.BROADCAST waterfall).
    {
      const bool issuer = lane == 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0;; ++it) {
        int unit;
        if constexpr (kDynamic) {
          const int slot = it % kSchedDepth;
          mbar_wait(&sched_empty_bar[slot], ((uint32_t)(it / kSchedDepth) & 1u) ^ 1u, wd);
          uint32_t u = 0;
          if (issuer) u = atomicAdd(prm.unit_counter, 1u);
          u = __shfl_sync(kFullMask, u, 0);
          unit = u < (uint32_t)prm.num_units ? (int)u : -1;
          if (issuer) {
            *reinterpret_cast<volatile int32_t*>(&sched_unit[slot]) = unit;
            mbar_arrive(&sched_full_bar[slot]);  // release semantics: the store above is visible to waiters
          }
        } else {
          unit = worker + it * num_workers;
          if (unit >= prm.num_units) unit = -1;
        }
        if (unit < 0) break;
        const UnitCoord uc = decode_unit(unit, prm);
        const int q_tile = uc.row_tile * kPair + cta_rank;
        if constexpr (kQRes) {
          // The unit was claimed a ring of operand stages ahead of its first MMA: pull its query tile (contiguous
          // rows) into L2 now, so that the epilogue warps' loads into tensor memory — the one step between two
          // units that nothing overlaps — do not wait for DRAM.
          if (issuer) {
            const int rows_q = min(kTileQ, prm.num_q - q_tile * kTileQ);
            if (rows_q > 0)
              l2_prefetch_bulk(static_cast<const uint8_t*>(prm.q_raw) + (size_t)q_tile * kTileQ * prm.dim_elems * 2,
                               (uint32_t)rows_q * (uint32_t)prm.dim_elems * 2u);
          }
        }
        for (int t = uc.t_begin * kSubTiles; t < uc.t_end * kSubTiles; ++t) {
          for (int kb = 0; kb < prm.num_k_blocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1, wd);
            if (!elect_one()) {
              // nothing to issue on this lane
            } else if constexpr (kQRes) {
              // only the gallery half-tile's k-slice: the query tile is already in tensor memory
              mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
              if (prm.l2_hints & 1)
                tma_load_2d_hint(smem_g + stage * kStageBytesG, &tmap_g, &full_bar[stage],
                                 kb * prm.elems_per_kblock, t * kAccCols, pol_keep);
              else
                tma_load_2d(smem_g + stage * kStageBytesG, &tmap_g, &full_bar[stage],
                            kb * prm.elems_per_kblock, t * kAccCols);
            } else if constexpr (kPair == 2) {
              // the leader's barrier collects the bytes of both CTAs' loads
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
              tma_load_2d_pair(smem_q + stage * kStageBytesQ, &tmap_q, &full_bar[stage],
                               kb * prm.elems_per_kblock, q_tile * kTileQ);
              tma_load_2d_pair(smem_g + stage * kStageBytesG, &tmap_g, &full_bar[stage],
                               kb * prm.elems_per_kblock, t * kTileG + cta_rank * (kTileG / 2));
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
              tma_load_2d(smem_q + stage * kStageBytesQ, &tmap_q, &full_bar[stage],
                          kb * prm.elems_per_kblock, q_tile * kTileQ);
              tma_load_2d(smem_g + stage * kStageBytesG, &tmap_g, &full_bar[stage],
                          kb * prm.elems_per_kblock, t * kTileG);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();  // reconverge before the (warp-aligned) teardown barriers
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer ----
    // The WHOLE warp runs this loop and one lane chosen by elect.sync executes the tcgen05.mma /
    // commit instructions.  Their operands (smem descriptors, TMEM addresses, barrier addresses) must
    // be provably warp-uniform — values read from shared memory go through a shuffle — so that the
    // compiler keeps them in uniform registers, and the predicate must come from elect.sync so that it
    // emits the instructions back to back.  With a single divergent thread (or an `if (lane == 0)`)
    // it wraps every UTCHMMA in an ELECT / R2UR.BROADCAST waterfall (~60 cycles per instruction), the
    // issue loop then takes ~600 cycles per k-block of 512 MMA cycles and the tensor pipe starves
    // (tools/gpu_probe.py diag).
    if (cta_rank == 0) {
      constexpr uint32_t idesc = make_instr_desc(kTF32 ? 2u : 1u, kTileQ * kPair, kAccCols);
      const uint32_t tmem_u = __shfl_sync(kFullMask, tmem_base, 0);
      const bool issuer = lane == 0;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const bool diag = (K1_DIAG_FLAGS(prm) & 64) != 0;
      long long w_acc = 0, w_full = 0, w_issue = 0, w_commit = 0, n_kb = 0;
      const long long t_loop0 = clock64();
      for (int it = 0;; ++it) {
        const int unit = __shfl_sync(kFullMask, next_unit_consumer(it), 0);
        __syncwarp();  // every lane has read the ring slot
        if (issuer) release_unit_slot(it);
        if (unit < 0) break;
        const UnitCoord uc = decode_unit(unit, prm);
        if constexpr (kQRes) {
          mbar_wait(q_ready_bar, (uint32_t)it & 1u, wd);  // the epilogue warps stored this unit's query tile
          tc_fence_after();
        }
        for (int t = uc.t_begin * kSubTiles; t < uc.t_end * kSubTiles; ++t) {
          long long tw = diag ? clock64() : 0;
          mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1, wd);  // epilogue drained this accumulator
          if (diag) w_acc += clock64() - tw;
          tc_fence_after();
          const uint32_t d_tmem = tmem_u + Cfg::kAccBase + acc * kAccCols;
          for (int kb = 0; kb < prm.num_k_blocks; ++kb) {
            tw = diag ? clock64() : 0;
            mbar_wait(&full_bar[stage], phase, wd);
            if (diag) w_full += clock64() - tw;
            tc_fence_after();
            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_q + stage * kStageBytesQ));
            const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_g + stage * kStageBytesG));
            tw = diag ? clock64() : 0;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // 4 × 32-byte K steps inside the 128-byte swizzle atom
                if constexpr (kQRes) umma_ts_f16(d_tmem, tmem_u + kb * 32 + k * 8, b_desc + 2 * k, idesc, (kb | k) != 0);
                else if constexpr (kPair == 2) umma_ss_pair<kTF32>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                else umma_ss<kTF32>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              }
            }
            if (diag) { const long long t1 = clock64(); w_issue += t1 - tw; tw = t1; }
            if (elect_one()) {
              if constexpr (kPair == 2) umma_commit_pair(&empty_bar[stage]);
              else umma_commit(&empty_bar[stage]);
            }
            if (diag) { w_commit += clock64() - tw; ++n_kb; }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) {
            if constexpr (kPair == 2) umma_commit_pair(&acc_full_bar[acc]);
            else umma_commit(&acc_full_bar[acc]);
          }
          __syncwarp();
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
      if (diag && issuer && blockIdx.x < 148) {
        g_k1_diag[blockIdx.x * 8 + 0] = (unsigned long long)w_acc;
        g_k1_diag[blockIdx.x * 8 + 1] = (unsigned long long)w_full;
        g_k1_diag[blockIdx.x * 8 + 2] = (unsigned long long)(clock64() - t_loop0);
        g_k1_diag[blockIdx.x * 8 + 5] = (unsigned long long)w_issue;   // cycles inside the 4-MMA issue blocks
        g_k1_diag[blockIdx.x * 8 + 6] = (unsigned long long)w_commit;  // cycles inside tcgen05.commit
        g_k1_diag[blockIdx.x * 8 + 7] = (unsigned long long)n_kb;      // k-blocks issued
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue ----
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the ones this warp can read
    const int half = ew >> 2;      // column half when two warps share a lane quarter
    const int row = quarter * 32 + lane;
    const int col_begin = half * Cfg::kColsPerWarp;
    const int lhalf = Cfg::kListsPerRow == 2 ? half : 0;  // which of the row's lists this warp works on
    const bool feeder = Cfg::kFeed && half == 1;          // forwards its hits to the quarter's list owner
    uint32_t fq_pos = 0;  // owner: entries consumed from this row's queue; feeder: entries pushed
    [[maybe_unused]] float thr_pubbed = INFINITY;  // owner: last threshold published to the feeder
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;

    // Resident-query form: store the query tile of unit `u` into tensor memory.  Every thread stores ITS row (TMEM lane =
    // query row, column j = 32-bit word j of the row, i.e. two bf16 per column, K-major); the two warps of a lane quarter
    // take alternate k-blocks, and the MMA issuer is released once all eight have arrived.  May only run after every MMA
    // of the previous unit has completed (its last accumulator was seen full).
    [[maybe_unused]] bool q_preloaded = false;
    [[maybe_unused]] auto load_q_tile = [&](int u) {
      if constexpr (kQRes) {
        const UnitCoord un = decode_unit(u, prm);
        const int qn = (un.row_tile * kPair + cta_rank) * kTileQ + row;
        const bool valid = qn < prm.num_q;
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(prm.q_raw) + (size_t)qn * prm.dim_elems * 2);
        const int row_words = prm.dim_elems / 2;
        for (int j = half; j < prm.num_k_blocks; j += Cfg::kColSplit) {
          uint32_t w[32];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            uint4 t4 = make_uint4(0u, 0u, 0u, 0u);
            if (valid && j * 32 + v * 4 < row_words) t4 = (prm.l2_hints & 2) ? ld_stream_v4(src + j * 8 + v, pol_stream) : __ldg(src + j * 8 + v);
            w[4 * v] = t4.x; w[4 * v + 1] = t4.y; w[4 * v + 2] = t4.z; w[4 * v + 3] = t4.w;
          }
          tmem_st_32x32b_x32(tmem_base + lane_addr + j * 32, w);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(q_ready_bar);
      }
    };

    for (int it = 0;; ++it) {
      const int unit = next_unit_consumer(it);
      __syncwarp();  // every lane has read the ring slot
      if (lane == 0) release_unit_slot(it);
      if (unit < 0) break;
      const UnitCoord uc = decode_unit(unit, prm);
      const int q_tile = uc.row_tile * kPair + cta_rank;
      const int q = q_tile * kTileQ + row;
      const bool q_valid = q < prm.num_q;

      if constexpr (kQRes) {
        // The unit's query tile normally went into tensor memory while the previous unit's last accumulator was still
        // being worked on (load_q_tile below); only the first unit of this CTA, or one after an empty unit, loads it here.
        if (!q_preloaded) load_q_tile(unit);
        q_preloaded = false;
      }

      // This thread's list: entry p lives at [p * kTileQ + row] (conflict-free / coalesced).
      // Global slot of the list: keyed by (partition, query tile), shared by all its chunks.
      const size_t list_slot = ((size_t)uc.part * prm.q_tile_stride + q_tile) * Cfg::kListsPerRow + lhalf;
      float* lv = list_val_s + lhalf * kCap * kTileQ;
      [[maybe_unused]] float* lg = list_grp_s + lhalf * (kCap / 8) * kTileQ;
      float* gval = prm.cand_val + list_slot * kCap * kTileQ;
      int32_t* gidx = prm.cand_idx + list_slot * kCap * kTileQ;
      int32_t* li;
      if constexpr (Cfg::kIdxInSmem) li = list_idx_s + lhalf * kCap * kTileQ;
      else li = gidx;
      int32_t* done_flag = prm.chunk_done + (size_t)uc.part * prm.q_tile_stride + q_tile;
      float thr = INFINITY;      // insertion threshold = min(own list maximum, shared threshold)
      float own_max = INFINITY;  // maximum of this thread's list (+inf until it is full)
      float published = INFINITY;
      int maxpos = 0;
      float lo = -INFINITY, hi = -INFINITY;
      int cnt = 0;
      constexpr int kPend = 4;  // accepted candidates waiting per lane (flush_pending)
      float pe0 = 0.f, pe1 = 0.f, pe2 = 0.f, pe3 = 0.f;
      int pi0 = 0, pi1 = 0, pi2 = 0, pi3 = 0, pn = 0;
      if constexpr (kSelect) {
        if (feeder) {
          // the owner publishes this unit's starting threshold before the feeder may filter with it
          if (lane == 0) {
            const long long t0 = clock64();
            while (unit_ready[quarter] != it) {
              if (watchdog_expired(t0, wd)) {
                printf("sbir: feeder start timed out (unit %d)\n", unit);
                __trap();
              }
            }
          }
          __syncwarp();
          __threadfence_block();
        } else if (uc.chunk == 0) {
#pragma unroll 4
          for (int p = 0; p < kCap; ++p) lv[p * kTileQ + row] = INFINITY;
          if constexpr (Cfg::kTwoLevel) {
            for (int u = 0; u < kCap / 8; ++u) lg[u * kTileQ + row] = INFINITY;
          }
        } else {
          // continue the list the previous chunk of this (partition, query tile) left behind
          if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire(done_flag) < uc.chunk) {
              if (watchdog_expired(t0, wd)) {
                printf("sbir: chunk hand-over timed out (unit %d)\n", unit);
                __trap();
              }
            }
          }
          __syncwarp();
          (void)ld_acquire(done_flag);
          // latency-bound (one L2 round trip per batch of loads in flight): 16-32 loads per batch
          if (K1_DIAG_FLAGS(prm) & 32) {  // A/B: four loads in flight
#pragma unroll 4
            for (int p = 0; p < kCap; ++p) {
              lv[p * kTileQ + row] = __ldcg(gval + p * kTileQ + row);
              if constexpr (Cfg::kIdxInSmem) li[p * kTileQ + row] = __ldcg(gidx + p * kTileQ + row);
            }
          } else
#pragma unroll
          for (int p0 = 0; p0 < kCap; p0 += 32) {
            constexpr int kBatch = kCap < 32 ? kCap : 32;
            float tv[kBatch];
            [[maybe_unused]] int32_t ti[kBatch];
#pragma unroll
            for (int p = 0; p < kBatch; ++p) {
              if (kQRes && (prm.l2_hints & 4)) {
                tv[p] = ld_cg_hint(gval + (p0 + p) * kTileQ + row, pol_stream);
                if constexpr (Cfg::kIdxInSmem) ti[p] = ld_cg_hint(gidx + (p0 + p) * kTileQ + row, pol_stream);
              } else {
                tv[p] = __ldcg(gval + (p0 + p) * kTileQ + row);
                if constexpr (Cfg::kIdxInSmem) ti[p] = __ldcg(gidx + (p0 + p) * kTileQ + row);
              }
            }
#pragma unroll
            for (int p = 0; p < kBatch; ++p) {
              lv[(p0 + p) * kTileQ + row] = tv[p];
              if constexpr (Cfg::kIdxInSmem) li[(p0 + p) * kTileQ + row] = ti[p];
            }
          }
          if constexpr (Cfg::kTwoLevel) {
            for (int u = 0; u < kCap / 8; ++u) {
              float gm = -INFINITY;
#pragma unroll
              for (int v = 0; v < 8; ++v) gm = fmaxf(gm, lv[(u * 8 + v) * kTileQ + row]);
              lg[u * kTileQ + row] = gm;
            }
          }
          own_max = __ldcg(prm.row_max + list_slot * kTileQ + row);
          maxpos = __ldcg(prm.row_maxpos + list_slot * kTileQ + row);
          thr = own_max;
        }
        if constexpr (Cfg::kFeed) {
          if (!feeder) {
            thr_pub[row] = thr;
            thr_pubbed = thr;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) unit_ready[quarter] = it;
          }
        }
      }
      if constexpr (kRank) {
        if (q_valid) {
          lo = prm.rank_lo[q];
          hi = prm.rank_hi[q];
        }
      }
      int32_t shared_next = 0x7f800000;  // +inf in the ordered-int encoding
      if constexpr (kSelect) shared_next = ld_relaxed(prm.shared_thr + q_tile * kTileQ + row);

      // One candidate of this row enters the list when it beats the threshold (owner side) ...
      auto insert_list = [&](float ej, int gidx_e) {
        if (ej < thr) {
          lv[maxpos * kTileQ + row] = ej;
          li[maxpos * kTileQ + row] = gidx_e;
          if (own_max == INFINITY) {
            // The list is not full yet (its entries are finite, free slots hold +inf): slots fill in order —
            // maxpos is the fill count — and there is no maximum to search for until the last one is taken.
            // A cold list takes ~cap·(1 + ln(columns/cap)) insertions per row; this spares the first cap of them
            // the scan (1k x 10k evaluation: a quarter of all insertions).
            if (maxpos + 1 < kCap) {
              ++maxpos;
              return;
            }
            if constexpr (Cfg::kTwoLevel) {  // full from here on: group maxima of the (now all finite) entries
              for (int u = 0; u < kCap / 8; ++u) {
                float gm = -INFINITY;
#pragma unroll
                for (int v = 0; v < 8; ++v) gm = fmaxf(gm, lv[(u * 8 + v) * kTileQ + row]);
                lg[u * kTileQ + row] = gm;
              }
            }
          }
          float mx = -INFINITY;
          int mp = 0;
          if constexpr (Cfg::kTwoLevel) {
            // lists of 64/128 keep a maximum per group of 8: refresh the touched group,
            // pick the group holding the overall maximum, locate it inside that group
            // (24-32 shared loads instead of kCap).  Maxima come from FMNMX3 trees and the
            // positions from independent equality tests, so no step is a long
            // compare-and-select chain (one epilogue warp per scheduler: latency is exposed).
            const int g = maxpos >> 3;
            float w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = lv[(g * 8 + u) * kTileQ + row];
            const float gm = fmaxf(fmax3(fmax3(w[0], w[1], w[2]), fmax3(w[3], w[4], w[5]), w[6]), w[7]);
            lg[g * kTileQ + row] = gm;
            constexpr int kGroups = kCap / 8;
            float gmx[kGroups];
#pragma unroll
            for (int u = 0; u < kGroups; ++u) gmx[u] = lg[u * kTileQ + row];
            float t8[kGroups / 2];
#pragma unroll
            for (int u = 0; u < kGroups / 2; ++u) t8[u] = fmaxf(gmx[2 * u], gmx[2 * u + 1]);
            if constexpr (kGroups == 16)
              mx = fmax3(fmax3(t8[0], t8[1], t8[2]), fmax3(t8[3], t8[4], t8[5]), fmaxf(t8[6], t8[7]));
            else
              mx = fmaxf(fmax3(t8[0], t8[1], t8[2]), t8[3]);
            int bg = 0;
#pragma unroll
            for (int u = 1; u < kGroups; ++u) bg = (gmx[u] == mx) ? u : bg;
            if (bg != g) {
#pragma unroll
              for (int u = 0; u < 8; ++u) w[u] = lv[(bg * 8 + u) * kTileQ + row];
            }
            mp = bg * 8;
#pragma unroll
            for (int u = 1; u < 8; ++u) mp = (w[u] == mx) ? bg * 8 + u : mp;
          } else {
#pragma unroll 8
            for (int p = 0; p < kCap; ++p) {
              const float v = lv[p * kTileQ + row];
              if (v > mx) { mx = v; mp = p; }
            }
          }
          own_max = mx;
          thr = fminf(thr, mx);
          maxpos = mp;
        }
      };
      // ... and is queued for exact evaluation when it sits in the rank band.
      auto band_check = [&](float ej, int gidx_e) {
        if constexpr (kRank) {
          if (ej >= lo && ej < hi) {
            // approximate comparison against d_pos is not trustworthy: queue the
            // pair for exact evaluation (finalize.cu: rank_resolve_kernel)
            const uint32_t slot = atomicAdd(prm.pool_count, 1u);
            if (slot < prm.pool_cap) {
              prm.pool_q[slot] = q;
              prm.pool_idx[slot] = gidx_e;
            } else {
              atomicAdd(prm.dropped + q, 1);
            }
          }
        }
      };
      // Once the lists have warmed up hits are sparse — one lane of the warp at a time — and an
      // insertion executed for a single lane costs the warp as much as one for all 32.  Accepted
      // candidates therefore wait in a per-lane queue of kPend registers and the lists are only
      // updated when some lane's queue is full (or the unit ends): then every lane inserts its
      // pending candidates side by side, ~15 per round instead of 1.  The thresholds lag by at
      // most kPend insertions per row, which only lets a few more candidates through.
      auto flush_pending = [&]() {
        const int maxn = __reduce_max_sync(kFullMask, pn);
        if (maxn > 0) { if (pn > 0) insert_list(pe0, pi0); __syncwarp(); }
        if (maxn > 1) { if (pn > 1) insert_list(pe1, pi1); __syncwarp(); }
        if (maxn > 2) { if (pn > 2) insert_list(pe2, pi2); __syncwarp(); }
        if (maxn > 3) { if (pn > 3) insert_list(pe3, pi3); __syncwarp(); }
        pn = 0;
      };
      auto pend_push = [&](float ej, int gidx_e) {  // caller made sure pn < kPend
        pe3 = pe2; pi3 = pi2;
        pe2 = pe1; pi2 = pi1;
        pe1 = pe0; pi1 = pi0;
        pe0 = ej; pi0 = gidx_e;
        ++pn;
      };
      // Feeder side of a candidate: forward it to the owner of the row's list.
      auto push_feed = [&](float ej, int gidx_e) {
        if constexpr (Cfg::kFeed) {
          const long long t0 = clock64();
          while (fq_pos - fq_head[row] >= (uint32_t)Cfg::kFeedDepth) {  // queue full: the owner drains it
            if (watchdog_expired(t0, wd)) {
              printf("sbir: feeder queue stuck (unit %d)\n", unit);
              __trap();
            }
          }
          const int slot = (int)(fq_pos % (uint32_t)Cfg::kFeedDepth);
          fq_val[slot * kTileQ + row] = ej;
          fq_idx[slot * kTileQ + row] = gidx_e;
          __threadfence_block();
          fq_tail[row] = ++fq_pos;
        }
      };
      auto consume = [&](float ej, int gidx_e) {
        if (feeder) {
          if (ej < thr) push_feed(ej, gidx_e);
        } else if (ej < thr) {
          pend_push(ej, gidx_e);
        }
        band_check(ej, gidx_e);
      };
      // Owner side of the queue: insert what the feeder forwarded (one entry per row and round).
      auto drain_feed = [&]() {
        if constexpr (Cfg::kFeed) {
          uint32_t tail = fq_tail[row];
          while (__any_sync(kFullMask, tail != fq_pos)) {
            if (__any_sync(kFullMask, pn == kPend)) flush_pending();
            if (tail != fq_pos) {
              __threadfence_block();
              const int slot = (int)(fq_pos % (uint32_t)Cfg::kFeedDepth);
              const float ej = fq_val[slot * kTileQ + row];
              const int gidx_e = fq_idx[slot * kTileQ + row];
              if (ej < thr) pend_push(ej, gidx_e);
              __threadfence_block();
              fq_head[row] = ++fq_pos;
            }
            __syncwarp();
            tail = fq_tail[row];
          }
        }
      };

      for (int t = uc.t_begin * kSubTiles; t < uc.t_end * kSubTiles; ++t) {  // t counts accumulator-sized pieces
        if constexpr (kSelect && Cfg::kFeed) {
          if (!feeder) {
            // The feeder may lag a tile behind and fill its queue while this accumulator's next
            // turn still waits for the feeder's own arrival: keep draining while waiting.
            const long long t0 = clock64();
            while (__shfl_sync(kFullMask, (int)mbar_try_wait(&acc_full_bar[acc], acc_phase), 0) == 0) {
              drain_feed();
              if (watchdog_expired(t0, wd)) {
                printf("sbir: accumulator wait timed out (unit %d)\n", unit);
                __trap();
              }
            }
          }
        }
        const long long tw_e = (K1_DIAG_FLAGS(prm) & 64) ? clock64() : 0;
        mbar_wait(&acc_full_bar[acc], acc_phase, wd);
        if ((K1_DIAG_FLAGS(prm) & 64) && ew == 0 && lane == 0 && blockIdx.x < 148)
          g_k1_diag[blockIdx.x * 8 + 3] += (unsigned long long)(clock64() - tw_e);
        tc_fence_after();
        if constexpr (kQRes) {
          // Last accumulator of the unit: every MMA that reads the query tile has completed, so the NEXT unit's tile
          // goes into tensor memory first and the MMA issuer starts on it while this accumulator's epilogue, the list
          // parking and the next unit's list reload run — the tensor pipe no longer idles through them between units.
          if (prm.q_early && t == uc.t_end * kSubTiles - 1) {
            const int next = next_unit_consumer(it + 1);   // published by the scheduler when it finished feeding this unit
            if (next >= 0) {
              load_q_tile(next);
              q_preloaded = true;
            }
          }
        }
        if constexpr (kSelect) {
          // Another partition (or the other column half) scanning the same query may already
          // hold `cap` candidates below some value: nothing at or above it can reach the final
          // best-`cap`, so adopt it as an upper bound on this list's threshold.  The value was
          // requested before waiting for the accumulator (L2 round trip off the critical path).
          if (!feeder) thr = fminf(thr, ordered_int_to_float(shared_next));
        }
        const float* gv = prm.gvec + (size_t)t * kAccCols + col_begin;
#pragma unroll 1
        for (int c = 0; c < ((K1_DIAG_FLAGS(prm) & 8) ? 0 : Cfg::kColsPerWarp / 32); ++c) {
          const uint32_t taddr = tmem_base + Cfg::kAccBase + lane_addr + acc * kAccCols + col_begin + c * 32;
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr, r);
          tmem_ld_wait();
          const int gcol0 = t * kAccCols + col_begin + c * 32;  // gallery row of column 0 of this chunk
          if constexpr (kSelect && Cfg::kFeed) {
            if (feeder) {
              thr = thr_pub[row];  // may lag behind the owner: then a few extra hits are forwarded
            } else {
              drain_feed();
              if (thr < thr_pubbed) {
                thr_pub[row] = thr;
                thr_pubbed = thr;
              }
            }
          }
          if constexpr (kSelect) {
            // Cheap conservative screen before any per-element work: every e of this chunk is
            // >= bound (rounding is monotone, so this holds for the computed values too);
            // when no row of the warp can beat its threshold the chunk costs ~20 FMNMX + 1 FFMA.
            {
              const float smax = reduce32<true>([&](int j) { return __uint_as_float(r[j]); });
              const float4 gm4 = __ldg(reinterpret_cast<const float4*>(prm.gmin + (size_t)gcol0 / 8));
              const float gmin = fminf(fmin3(gm4.x, gm4.y, gm4.z), gm4.w);
              const float bound = (kMetric == SBIR_EUCLIDEAN) ? fmaf(-2.f, smax, gmin) : fminf(0.f, __fmul_rn(smax, gmin));
              const float lim0 = kRank ? fmaxf(thr, hi) : thr;
              if (!(K1_DIAG_FLAGS(prm) & 16) && !__any_sync(kFullMask, bound < lim0)) continue;
            }
            float e[32];
            const float4* gv4 = reinterpret_cast<const float4*>(gv + c * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 g4 = __ldg(gv4 + j4);
              const float gj[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float s = __uint_as_float(r[j4 * 4 + u]);
                e[j4 * 4 + u] = (kMetric == SBIR_EUCLIDEAN) ? fmaf(-2.f, s, gj[u]) : __fmul_rn(s, gj[u]);
              }
            }
            // exact minima per sub-chunk of 8 columns (FMNMX3: 4 instructions each) and of the chunk
            float mg[4];
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8)
              mg[g8] = fminf(fmin3(fmin3(e[8 * g8], e[8 * g8 + 1], e[8 * g8 + 2]), fmin3(e[8 * g8 + 3], e[8 * g8 + 4], e[8 * g8 + 5]), e[8 * g8 + 6]),
                             e[8 * g8 + 7]);
            const float m = fminf(fmin3(mg[0], mg[1], mg[2]), mg[3]);
            // rank: rows closer than the band (e < lo <= hi <= lim) are counted inside the gated
            // path below — a chunk whose minimum is not below lim has none of them
            const float lim = kRank ? fmaxf(thr, hi) : thr;
            if (!__any_sync(kFullMask, m < lim)) continue;
            // Hit mask of this row, built only for the sub-chunks of 8 columns in which some row of
            // the warp has a hit (once the lists have warmed up: a few lanes, one hit each).
            uint32_t mask = 0;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              if (!__any_sync(kFullMask, mg[g8] < lim)) continue;
#pragma unroll
              for (int j = 8 * g8; j < 8 * g8 + 8; ++j) {
                if constexpr (kRank) cnt += (e[j] < lo) ? 1 : 0;
                mask |= (e[j] < lim ? 1u : 0u) << j;
              }
            }
            // Every lane drains ITS OWN hits, one per round, all lanes side by side: the number of
            // rounds is the largest hit count of any row, not the number of columns that have a hit
            // somewhere (list warm-up: 2-5x fewer rounds).  The value of the lane's next hit column
            // is picked out of the registers with a 5-level select tree.
            while (__any_sync(kFullMask, mask != 0)) {
              if (!feeder && __any_sync(kFullMask, pn == kPend)) flush_pending();
              if (mask != 0) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                float s16[16], s8[8], s4[4];
#pragma unroll
                for (int u = 0; u < 16; ++u) s16[u] = (j & 1) ? e[2 * u + 1] : e[2 * u];
#pragma unroll
                for (int u = 0; u < 8; ++u) s8[u] = (j & 2) ? s16[2 * u + 1] : s16[2 * u];
#pragma unroll
                for (int u = 0; u < 4; ++u) s4[u] = (j & 4) ? s8[2 * u + 1] : s8[2 * u];
                const float s2a = (j & 8) ? s4[1] : s4[0], s2b = (j & 8) ? s4[3] : s4[2];
                consume((j & 16) ? s2b : s2a, gcol0 + j);
              }
              __syncwarp();
            }
          } else {
            float e[32];
            const float4* gv4 = reinterpret_cast<const float4*>(gv + c * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 g4 = __ldg(gv4 + j4);
              const float gj[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float s = __uint_as_float(r[j4 * 4 + u]);
                e[j4 * 4 + u] = (kMetric == SBIR_EUCLIDEAN) ? fmaf(-2.f, s, gj[u]) : __fmul_rn(s, gj[u]);
              }
            }
            if constexpr (kMode == kModeDump) {
              if (q_valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (gcol0 + j < prm.num_g) prm.dump[(size_t)q * prm.num_g + gcol0 + j] = e[j];
              }
            }
          }
        }
        if constexpr (kSelect) {
          if (!feeder && own_max < published) {  // list is full and its maximum dropped: share it
            atomicMin(prm.shared_thr + q_tile * kTileQ + row, float_to_ordered_int(own_max));
            published = own_max;
          }
          shared_next = ld_relaxed(prm.shared_thr + q_tile * kTileQ + row);
        }
        // Accumulator fully read: hand it back to the MMA warp.
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty_bar[acc]), 0));
          else mbar_arrive(&acc_empty_bar[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }

      if constexpr (kSelect) {
        if constexpr (kRank) {
          if (q_valid && cnt) atomicAdd(prm.cnt_less + q, cnt);
        }
        if (feeder) {
          // everything this warp forwarded is in the queue: tell the owner, then run ahead
          __threadfence_block();
          __syncwarp();
          if (lane == 0) feeder_done[quarter] = it;
        } else {
          if constexpr (Cfg::kFeed) {
            // keep draining until the feeder has finished the unit and its queue is empty
            const long long t0 = clock64();
            for (;;) {
              const bool done = __shfl_sync(kFullMask, (int)(feeder_done[quarter] >= it), 0) != 0;
              __threadfence_block();
              drain_feed();
              if (done) break;
              if (watchdog_expired(t0, wd)) {
                printf("sbir: feeder hand-over timed out (unit %d)\n", unit);
                __trap();
              }
            }
          }
          flush_pending();
          // Park the list in global memory: the next chunk of this (partition, query tile) — on
          // whichever SM it lands — or finalize.cu picks it up from there.
#pragma unroll 4
          for (int p = 0; p < kCap; ++p) {
            if (kQRes && (prm.l2_hints & 4)) {
              st_hint(gval + p * kTileQ + row, lv[p * kTileQ + row], pol_stream);
              if constexpr (Cfg::kIdxInSmem) st_hint(gidx + p * kTileQ + row, li[p * kTileQ + row], pol_stream);
            } else {
              gval[p * kTileQ + row] = lv[p * kTileQ + row];
              if constexpr (Cfg::kIdxInSmem) gidx[p * kTileQ + row] = li[p * kTileQ + row];
            }
          }
          prm.row_max[list_slot * kTileQ + row] = own_max;
          prm.row_maxpos[list_slot * kTileQ + row] = maxpos;
          __threadfence();
          // all list-owning epilogue warps parked their lists
          asm volatile("bar.sync 1, %0;" ::"r"((Cfg::kFeed ? 4 : kEpiWarps) * 32) : "memory");
          if (ew == 0 && lane == 0) st_release(done_flag, uc.chunk + 1);
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (kPair == 2) {
    cluster_sync_all();  // the peer may still be arriving on / reading this CTA's shared memory
    if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <bool kTF32, int kMetric, int kMode, int kCap, int kEpiWarps, int kPair, bool kQRes = false>
int launch_inst(const CUtensorMap& tq, const CUtensorMap& tg, const K1Params& prm, int num_sms, cudaStream_t st) {
  using Cfg = K1Config<kCap, kEpiWarps, kPair, kQRes>;
  auto kern = dist_topk_kernel<kTF32, kMetric, kMode, kCap, kEpiWarps, kPair, kQRes>;
  SBIR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  int workers = num_sms / kPair;
  if (workers > prm.num_units) workers = prm.num_units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(workers * kPair));
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // CTA pairs walk the units with a STATIC stride and wait on chunk hand-overs from other pairs, so every
  // launched pair must be resident at the same time (the dynamic hand-out of single-CTA tiles needs no such
  // guarantee: a unit is only ever waited for by CTAs that claimed a LATER unit, and its claimant is running).
  // The grid is therefore clamped to what the device can hold (MIG slices, reduced shared memory carve-outs)
  // and launched cooperatively: if the pairs cannot all be co-resident the launch FAILS with an error status
  // instead of spinning into the watchdog.
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kPair == 2 ? (prm.pair_cooperative ? 2 : 1) : 0;  // option k1_pair_coop = 0: Nsight Compute cannot replay cooperative cluster launches
  if constexpr (kPair == 2) {
    int max_clusters = 0;
    SBIR_CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
    if (max_clusters < 1) return SBIR_ERR_UNSUPPORTED;
    if (workers > max_clusters) {
      workers = max_clusters;
      cfg.gridDim = dim3((unsigned)(workers * kPair));
    }
  }
  profile_k1_begin(st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tq, tg, prm);
  if (e != cudaSuccess && kPair == 2) {
    // Some environments reject the cooperative flavour of a cluster launch (Nsight Compute's kernel replay does);
    // the grid is already clamped to what fits, which is the practical guarantee on an otherwise idle device.
    (void)cudaGetLastError();
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, tq, tg, prm);
  }
  profile_k1_end(st);
  if (e != cudaSuccess) {
    set_last_cuda_error((int)e);
    return SBIR_ERR_CUDA;
  }
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

template <bool kTF32, int kMetric, int kEpiWarps>
int dispatch_mode_cap(int mode, int cap, int pair, bool qres, const CUtensorMap& tq, const CUtensorMap& tg, const K1Params& prm,
                      int num_sms, cudaStream_t st) {
  if constexpr (!kTF32 && kEpiWarps == 8) {
    if (qres) {  // resident-query form (bf16 rows <= 1 KB, small lists)
#define SBIR_K1_QRES_CASE(M, C) \
  if (mode == M && cap == C) return launch_inst<false, kMetric, M, C, 8, 1, true>(tq, tg, prm, num_sms, st);
      SBIR_K1_QRES_CASE(kModeTopk, 16)
      SBIR_K1_QRES_CASE(kModeTopk, 32)
      SBIR_K1_QRES_CASE(kModeTopkRank, 16)
      SBIR_K1_QRES_CASE(kModeTopkRank, 32)
      SBIR_K1_QRES_CASE(kModeTopk, 64)
      SBIR_K1_QRES_CASE(kModeTopk, 128)
      SBIR_K1_QRES_CASE(kModeTopkRank, 64)
      SBIR_K1_QRES_CASE(kModeTopkRank, 128)
#undef SBIR_K1_QRES_CASE
      return SBIR_ERR_UNSUPPORTED;
    }
  }
#define SBIR_K1_CASE(M, C)                                                                            \
  if (mode == M && cap == C) {                                                                        \
    if (pair == 2) return launch_inst<kTF32, kMetric, M, C, kEpiWarps, 2>(tq, tg, prm, num_sms, st);  \
    return launch_inst<kTF32, kMetric, M, C, kEpiWarps, 1>(tq, tg, prm, num_sms, st);                 \
  }
#define SBIR_K1_CASE1(M, C) \
  if (mode == M && cap == C) return launch_inst<kTF32, kMetric, M, C, kEpiWarps, 1>(tq, tg, prm, num_sms, st);
  // instantiated combinations: fp32 small lists run 4 warps, bf16 small lists 8 (two lists per
  // row), large lists 8 (owner + feeder) or 4 (SBIR_K1_FEED=0)
  if constexpr ((kEpiWarps == 8) != kTF32) {
    SBIR_K1_CASE(kModeTopk, 16)
    SBIR_K1_CASE(kModeTopk, 32)
    SBIR_K1_CASE(kModeTopkRank, 16)
    SBIR_K1_CASE(kModeTopkRank, 32)
    SBIR_K1_CASE(kModeDump, 16)
  }
  SBIR_K1_CASE(kModeTopk, 64)
  SBIR_K1_CASE(kModeTopk, 128)
  SBIR_K1_CASE(kModeTopkRank, 64)
  SBIR_K1_CASE(kModeTopkRank, 128)
#undef SBIR_K1_CASE
#undef SBIR_K1_CASE1
  return SBIR_ERR_UNSUPPORTED;
}

}  // namespace

int SBIR_K1_INST_NAME(int epi, int mode, int cap, int pair, bool qres, const CUtensorMap& tq, const CUtensorMap& tg,
                      const K1Params& prm, int num_sms, cudaStream_t st) {
  if (epi == 8) return dispatch_mode_cap<SBIR_K1_INST_TF32, SBIR_K1_INST_METRIC, 8>(mode, cap, pair, qres, tq, tg, prm, num_sms, st);
  return dispatch_mode_cap<SBIR_K1_INST_TF32, SBIR_K1_INST_METRIC, 4>(mode, cap, pair, qres, tq, tg, prm, num_sms, st);
}

#define SBIR_K1_CAT2(a, b) a##b
#define SBIR_K1_CAT(a, b) SBIR_K1_CAT2(a, b)
int SBIR_K1_CAT(SBIR_K1_INST_NAME, _diag)(unsigned long long* out) {
  unsigned long long host[148 * 8];
  if (cudaMemcpyFromSymbol(host, g_k1_diag, sizeof(host)) != cudaSuccess) return SBIR_ERR_CUDA;
  for (int i = 0; i < 148 * 8; ++i) out[i] += host[i];
  static const unsigned long long zeros[148 * 8] = {};
  if (cudaMemcpyToSymbol(g_k1_diag, zeros, sizeof(zeros)) != cudaSuccess) return SBIR_ERR_CUDA;
  return SBIR_OK;
}
#undef SBIR_K1_CAT
#undef SBIR_K1_CAT2

}  // namespace sbir
