// dist_topk_params.h — launch parameters of K1 (dist_topk_kernel.cuh) shared by its host side
// (dist_topk.cu) and the four translation units that instantiate the kernel.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace sbir {

constexpr int kSwizzleBytes = 128;  // one k-block = 128 bytes of features per row

struct K1Params {
  const float* gvec;
  const float* gmin;   // per 8 gallery rows: min of gvec (‖g‖² | −1/max(‖g‖,eps)), NaN padding ignored
  int num_q, num_g;
  int num_q_tiles, num_g_tiles, num_k_blocks;
  int num_row_tiles;   // query tiles (kPair = 1) or query-tile pairs (kPair = 2): rows of the unit grid
  int num_parts, tiles_per_part, num_chunks, tiles_per_chunk, num_units, part_fastest;
  int chunk_begin;     // first chunk step of this launch (streamed galleries: later launches continue the lists)
  int num_steps;       // chunk steps this launch covers
  int band_rows;       // rows of the unit grid per L2 band (decode_unit); num_row_tiles = one band
  int q_tile_stride;   // query-tile stride of candidate slots (num_q_tiles rounded up to even)
  int elems_per_kblock;
  const void* q_raw;   // query matrix in global memory (resident-query form: loaded into TMEM by the epilogue warps)
  int dim_elems;
  const int32_t* gate;     // optional: the kernel is a no-op unless *gate != 0 (escalation pass)
  int flags;               // -DSBIR_DIAG builds only (k1_flags option): 8 = epilogue skips the accumulator (mainloop alone), 16 = no chunk screen, 64 = cycle counters
  int q_early;             // resident-query form: next unit's query tile stored while the current unit's last accumulator is worked on (option k1_q_early = 0: A/B)
  int l2_hints;            // resident-query form: L2 eviction hints (gallery chunk evict_last, query tiles / parked lists evict_first)
  int pair_cooperative;    // CTA-pair launches carry the cooperative attribute (co-residency guaranteed or the launch fails)
  long long watchdog_cycles;  // bound on every spin / barrier wait (0 = none), see ptx.cuh
  uint32_t* unit_counter;  // [1] zeroed by the caller: next unit to hand out (kPair = 1)
  int32_t* chunk_done;     // [num_parts][q_tile_stride] zeroed: chunks finished per (partition, query tile)
  float* cand_val;         // [part][q_tile_stride][lists][cap][128]
  int32_t* cand_idx;
  float* row_max;          // [part][q_tile_stride][lists][128] list maximum carried between chunks
  int32_t* row_maxpos;
  const float* rank_lo;
  const float* rank_hi;
  int32_t* cnt_less;
  uint32_t* pool_count;
  uint32_t pool_cap;
  int32_t* pool_q;
  int32_t* pool_idx;
  int32_t* dropped;
  int32_t* shared_thr;
  float* dump;
};

// [rows, dim] row-major matrix → 2-D tensor map with a (128-byte × box_rows) SWIZZLE_128B box (dist_topk.cu).
int make_tmap(CUtensorMap* out, const void* base, int64_t rows, int64_t dim, int dtype, int box_rows);

// One launcher per (input type, metric) translation unit: dist_topk_{f32,bf16}_{euclidean,cosine}.cu.
// epi = epilogue warps (4 / 8), pair = 1 / 2 (CTA pairs), qres = resident-query form.
#define SBIR_K1_LAUNCHER(NAME)                                                                                        \
  int NAME(int epi, int mode, int cap, int pair, bool qres, const CUtensorMap& tq, const CUtensorMap& tg,           \
           const K1Params& prm, int num_sms, cudaStream_t st);                                                      \
  int NAME##_diag(unsigned long long* out)  /* adds this unit's per-CTA cycle counters to out[148*8] and clears them */
SBIR_K1_LAUNCHER(k1_launch_f32_euclidean);
SBIR_K1_LAUNCHER(k1_launch_f32_cosine);
SBIR_K1_LAUNCHER(k1_launch_bf16_euclidean);
SBIR_K1_LAUNCHER(k1_launch_bf16_cosine);
#undef SBIR_K1_LAUNCHER

}  // namespace sbir
