// batch_hard.cu — K3: batch-hard triplet mining + margin loss + gradients as ONE cooperative kernel
// (H8, a north_star extension; the reference forms triplets in its datasets,
// data_preparation.py:67-69,214-222, and only evaluates nn.TripletMarginLoss /
// TripletMarginWithDistanceLoss on them, train.py:164-175).
//
// Candidates X = cat(p, n) are never concatenated: p and n are read in place through two tensor
// maps.  The kernel runs three phases separated by grid-wide barriers (cooperative launch):
//   1. MINING TILES  tcgen05 (kind::tf32) tiles of A·Xᵀ — 128 anchors × 32 candidates per unit, the
//      K dimension split over several units so that a 256-anchor batch still spreads over all SMs —
//      TMA-fed (8-stage ring), accumulators in TMEM, raw dot products stored to a small L2-resident
//      matrix [k_split][anchors][candidates].  While the first tiles are in flight the epilogue warps
//      compute ‖a‖², ‖x‖² of all rows.
//   2. SELECT + LOSS + ANCHOR GRADIENT  one warp per anchor: e = ‖x‖² − 2·a·x (or the cosine form)
//      from the tiles, the approximate hardest positive / negative, then EVERY candidate whose
//      approximate value lies within twice the tensor-core error bound of it (e_margin, the same
//      certificate as the top-k path) is re-scored with the reference formula in fp32/fp64
//      (common.cuh) — so the selected pair is provably the exact arg-max / arg-min of the exact
//      distances, ties by the smaller index.  Hinge, weight, and the anchor's gradient row.
//   3. CANDIDATE GRADIENTS + MEAN  one warp per candidate row walks the anchors in index order and
//      accumulates the contributions of those that selected it (fixed order → bitwise
//      reproducible, no floating-point atomics); one warp sums the hinge terms in fixed order.
#include <cooperative_groups.h>

#include <cuda.h>

#include "common.cuh"
#include "dist_topk_params.h"
#include "kernels.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace sbir {

namespace {

constexpr int kBhTileA = 128;    // anchors per tile (UMMA M, TMEM lanes)
constexpr int kBhTileC = 32;     // candidates per tile (UMMA N, TMEM columns per accumulator)
constexpr int kBhStages = 4;      // 4 × 20 KB: two CTAs per SM (phases 2 and 3 are latency-bound and want the warps)
constexpr int kBhStageA = kBhTileA * kSwizzleBytes;  // 16 KB
constexpr int kBhStageC = kBhTileC * kSwizzleBytes;  // 4 KB
constexpr int kBhThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kBhWarps = kBhThreads / 32;
constexpr int kBhSmemBytes = 1024 + kBhStages * (kBhStageA + kBhStageC) + 256;
constexpr int kBhTmemCols = 64;  // two accumulators of 32 columns
constexpr int kBhMaxSplits = 8;
// phases 2 / 3 reuse the (idle) operand ring:
constexpr int kBhCache = 8192;   // e values of the CTA's anchor (more candidates: recomputed)
constexpr int kBhBand = 4096;    // band-list entries per anchor (more: every candidate is re-scored)
constexpr int kBhHits = 3072;    // hit-list entries per warp and batch (phase 3)
static_assert(kBhHits * 4 * (kBhThreads / 32) <= kBhStages * (kBhStageA + kBhStageC), "per-warp hit lists must fit the operand ring");
static_assert((kBhCache + kBhBand) * 4 <= kBhStages * (kBhStageA + kBhStageC), "e cache + band list must fit the operand ring");

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct BhParams {
  const float* a;
  const float* p;
  const float* n;
  int batch, dim, metric;
  float margin, kappa;
  const long long* anchor_label;  // [batch] or NULL: positive of anchor i is candidate i
  const long long* cand_label;    // [2*batch]
  int num_q_tiles, tiles_per_half, k_splits, kb_per_split, num_k_blocks, num_units;
  int dot_rows, dot_cols;         // padded extents of one split's dot matrix
  float* dots;                    // [k_splits][dot_rows][dot_cols]
  float* anorm;                   // [batch]   ‖a_i‖²
  float* cnorm;                   // [2*batch] ‖x_j‖²
  float* per_row;                 // [batch] hinge terms
  float* weight;                  // [batch] ∂loss/∂hinge_i (1/batch where the hinge is active)
  int* sel;                       // [batch][2] selected (positive, negative) candidate
  double* pair_d;                 // [batch][2] exact distances of the selected pairs
  float* pair_stat;               // [batch][8] cosine: ca, ma, cxp, sp, mxp, cxn, sn, mxn
  float* out_loss;
  long long* out_hard_index;
  float* ga;
  float* gp;
  float* gn;
  unsigned long long* timing;   // [4] phase boundaries seen by CTA 0 (globaltimer ns; diagnostics)
};

__device__ __forceinline__ const float* cand_row(const BhParams& P, int j) {
  return j < P.batch ? P.p + (size_t)j * P.dim : P.n + (size_t)(j - P.batch) * P.dim;
}
// column of candidate j in the dot matrix (each half is padded to whole tiles)
__device__ __forceinline__ int cand_col(const BhParams& P, int j) {
  return j < P.batch ? j : P.tiles_per_half * kBhTileC + (j - P.batch);
}

struct PairExact {
  double d;        // exact distance (reference formula, fp32 elements, fp64 sum)
  float cx, s, mx; // cosine: clamped candidate norm, similarity, 1 when ‖x‖ > eps
};

// Exact distance of (a_i, x_j) by one warp plus what the gradient needs.  ca = clamped anchor norm.
template <bool kVec>
__device__ __forceinline__ PairExact pair_exact(const float* __restrict__ ar, const float* __restrict__ xr, int dim,
                                                int metric, float ca, int lane) {
  PairExact st{};
  if (metric == SBIR_EUCLIDEAN) {
    st.d = sqrt(warp_sq_l2_eps<float, kVec>(ar, xr, dim, lane));
  } else {
    const double qx = warp_sq_norm<float, kVec>(xr, dim, lane);
    st.cx = clamped_norm(qx);
    st.mx = (float)sqrt(qx) > kCosineEps ? 1.f : 0.f;
    const double s = warp_cos_dot<float, kVec>(ar, xr, ca, st.cx, dim, lane);
    st.s = (float)s;
    st.d = 1.0 - s;
  }
  return st;
}

// diagnostics: per-CTA stage stamps (globaltimer ns) behind the four phase stamps
#define BH_STAMP(k) do { if (P.timing && threadIdx.x == 0) P.timing[8 + blockIdx.x * 8 + (k)] = globaltimer_ns(); } while (0)

template <bool kVec>
__global__ void __launch_bounds__(kBhThreads, 2)
batch_hard_fused_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_p,
                        const __grid_constant__ CUtensorMap tmap_n, const BhParams P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_c = smem + kBhStages * kBhStageA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBhStages * (kBhStageA + kBhStageC));
  uint64_t* full_bar = bars;                        // [kBhStages]
  uint64_t* empty_bar = bars + kBhStages;           // [kBhStages]
  uint64_t* acc_full_bar = bars + 2 * kBhStages;    // [2]
  uint64_t* acc_empty_bar = bars + 2 * kBhStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kBhStages + 4);

  const int warp = __shfl_sync(kFullMask, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool has_units = (int)blockIdx.x < P.num_units;
  if (P.timing && blockIdx.x == 0 && threadIdx.x == 0) P.timing[0] = globaltimer_ns();

  // ------------------------------------------------------------ phase 1: mining tiles ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_p);
    tma_prefetch_desc(&tmap_n);
    for (int s = 0; s < kBhStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&acc_full_bar[x], 1);
      mbar_init(&acc_empty_bar[x], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kBhTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int u, int& q_tile, int& c_tile, int& kb0, int& kb1) {
    const int split = u % P.k_splits;
    const int t = u / P.k_splits;
    c_tile = t % (2 * P.tiles_per_half);
    q_tile = t / (2 * P.tiles_per_half);
    kb0 = split * P.kb_per_split;
    kb1 = min(kb0 + P.kb_per_split, P.num_k_blocks);
    return split;
  };

  if (warp == 0) {
    if (has_units) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < P.num_units; u += gridDim.x) {
        int q_tile, c_tile, kb0, kb1;
        decode(u, q_tile, c_tile, kb0, kb1);
        const bool from_p = c_tile < P.tiles_per_half;
        const int c_row = (from_p ? c_tile : c_tile - P.tiles_per_half) * kBhTileC;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], kBhStageA + kBhStageC);
            tma_load_2d(smem_a + stage * kBhStageA, &tmap_a, &full_bar[stage], kb * 32, q_tile * kBhTileA);
            if (from_p) tma_load_2d(smem_c + stage * kBhStageC, &tmap_p, &full_bar[stage], kb * 32, c_row);
            else tma_load_2d(smem_c + stage * kBhStageC, &tmap_n, &full_bar[stage], kb * 32, c_row);
          }
          if (++stage == kBhStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (has_units) {
      constexpr uint32_t idesc = make_instr_desc(2u, kBhTileA, kBhTileC);
      const uint32_t tmem_u = __shfl_sync(kFullMask, tmem_base, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int u = blockIdx.x; u < P.num_units; u += gridDim.x) {
        int q_tile, c_tile, kb0, kb1;
        decode(u, q_tile, c_tile, kb0, kb1);
        mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + acc * kBhTileC;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * kBhStageA));
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_c + stage * kBhStageC));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss<true>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (uint32_t)((kb - kb0) | k) != 0);
          }
          if (elect_one()) umma_commit(&empty_bar[stage]);
          if (++stage == kBhStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&acc_full_bar[acc]);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // epilogue warps: row norms first (they would otherwise idle until the first accumulator is
    // complete), then the tiles
    {
      const int gw = blockIdx.x * 4 + (warp - 2), nw = gridDim.x * 4;
      for (int r = gw; r < 3 * P.batch; r += nw) {
        const float* row = r < P.batch ? P.a + (size_t)r * P.dim : cand_row(P, r - P.batch);
        const float sq = (float)warp_sq_norm<float, kVec>(row, P.dim, lane);
        if (lane == 0) {
          if (r < P.batch) P.anorm[r] = sq;
          else P.cnorm[r - P.batch] = sq;
        }
      }
    }
    if (has_units) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = blockIdx.x; u < P.num_units; u += gridDim.x) {
        int q_tile, c_tile, kb0, kb1;
        const int split = decode(u, q_tile, c_tile, kb0, kb1);
        mbar_wait(&acc_full_bar[acc], acc_phase);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + acc * kBhTileC, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty_bar[acc]);
        float4* dst = reinterpret_cast<float4*>(P.dots + ((size_t)split * P.dot_rows + (size_t)q_tile * kBhTileA + row) * P.dot_cols +
                                                (size_t)c_tile * kBhTileC);
#pragma unroll
        for (int v = 0; v < 8; ++v)
          dst[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                               __uint_as_float(r[4 * v + 3]));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kBhTmemCols);
  __threadfence();
  grid.sync();

  // ------------------------------------ phase 2: exact selection, hinge, anchor gradient ----
  // One CTA per anchor (a 256-anchor batch has ~1.7 anchors per SM: spreading an anchor over the six
  // warps of a CTA shortens the serial chain of L2 round trips that this phase consists of).
  if (P.timing && blockIdx.x == 0 && threadIdx.x == 0) P.timing[1] = globaltimer_ns();
  const int ncand = 2 * P.batch;
  const size_t split_stride = (size_t)P.dot_rows * P.dot_cols;
  // the operand ring is idle now: e values of the CTA's anchor, then the band list
  float* ecache = reinterpret_cast<float*>(smem);
  int* band = reinterpret_cast<int*>(smem) + kBhCache;   // (candidate << 1 | is_positive)
  __shared__ float s_red[3][kBhWarps];
  __shared__ int s_nband;
  __shared__ double s_bd[2][kBhWarps];      // per warp: exact best positive / negative distance (fp32-rounded)
  __shared__ int s_bi[2][kBhWarps];
  __shared__ PairExact s_bs[2][kBhWarps];
  __shared__ PairExact s_sel[2];
  __shared__ int s_seli[2];
  __shared__ float s_w;
  const int tid = threadIdx.x;
  BH_STAMP(0);
  for (int i = blockIdx.x; i < P.batch; i += gridDim.x) {
    const float* ar = P.a + (size_t)i * P.dim;
    const float qsq = __ldcg(P.anorm + i);
    const float ca = clamped_norm((double)qsq);  // fp32 norm as torch returns it, clamped (cosine)
    const long long my_label = P.anchor_label ? P.anchor_label[i] : 0;
    const float* drow = P.dots + (size_t)i * P.dot_cols;
    auto is_positive = [&](int j) -> bool { return P.anchor_label ? (P.cand_label[j] == my_label) : (j == i); };
    // approximate e of every candidate: 4 per thread with all their loads in flight
    float ap = -INFINITY, an = INFINITY, gmax = 0.f;
    if (tid == 0) s_nband = 0;
    for (int j0 = 0; j0 < ncand; j0 += 4 * kBhThreads) {
      float dot[4], cn[4];
      bool pos[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * kBhThreads + tid;
        dot[u] = 0.f; cn[u] = 0.f; pos[u] = false;
        if (j < ncand) {
          const int col = cand_col(P, j);
#pragma unroll
          for (int sp = 0; sp < kBhMaxSplits; ++sp)
            if (sp < P.k_splits) dot[u] += __ldcg(drow + sp * split_stride + col);
          cn[u] = __ldcg(P.cnorm + j);
          pos[u] = is_positive(j);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * kBhThreads + tid;
        if (j < ncand) {
          const float e = P.metric == SBIR_EUCLIDEAN ? fmaf(-2.f, dot[u], cn[u]) : -dot[u] / fmaxf(sqrtf(cn[u]), kCosineEps);
          gmax = fmaxf(gmax, cn[u]);
          if (pos[u]) ap = fmaxf(ap, e);
          else an = fminf(an, e);
          if (j < kBhCache) ecache[j] = e;
        }
      }
    }
    ap = warp_max(ap);
    an = -warp_max(-an);
    gmax = warp_max(gmax);
    if (lane == 0) { s_red[0][warp] = ap; s_red[1][warp] = an; s_red[2][warp] = gmax; }
    __syncthreads();
    BH_STAMP(1);
    ap = s_red[0][0]; an = s_red[1][0]; gmax = s_red[2][0];
#pragma unroll
    for (int w = 1; w < kBhWarps; ++w) {
      ap = fmaxf(ap, s_red[0][w]); an = fminf(an, s_red[1][w]); gmax = fmaxf(gmax, s_red[2][w]);
    }
    const float m2 = (float)(2.0 * e_margin(P.metric, qsq, gmax, P.kappa, P.dim));
    // band list: every candidate within the error band of the approximate extremum
    for (int j = tid; j < ncand; j += kBhThreads) {
      float e;
      if (j < kBhCache) {
        e = ecache[j];
      } else {  // beyond the cache (batches above 3072): recompute
        float dot = 0.f;
        const int col = cand_col(P, j);
        for (int sp = 0; sp < P.k_splits; ++sp) dot += __ldcg(drow + sp * split_stride + col);
        const float cn = __ldcg(P.cnorm + j);
        e = P.metric == SBIR_EUCLIDEAN ? fmaf(-2.f, dot, cn) : -dot / fmaxf(sqrtf(cn), kCosineEps);
      }
      const bool pos = is_positive(j);
      if (pos ? (e >= ap - m2) : (e <= an + m2)) {
        const int slot = atomicAdd(&s_nband, 1);
        if (slot < kBhBand) band[slot] = (j << 1) | (pos ? 1 : 0);
      }
    }
    __syncthreads();
    BH_STAMP(2);
    const int nband = s_nband;
    if (P.timing && threadIdx.x == 0) P.timing[8 + blockIdx.x * 8 + 5] = (unsigned long long)nband;  // diagnostics: band size
    // Exact re-scoring, one warp per band entry.  The arrival order of the entries does not matter: the
    // winner is chosen by the total order (fp32 distance, candidate index).  If the band overflows the
    // list (degenerate data: everything within rounding of everything), every candidate is re-scored.
    double bd[2] = {INFINITY, -INFINITY};   // [0] negative: smallest, [1] positive: largest
    int bi[2] = {-1, -1};
    PairExact bs[2] = {};
    const int nwork = nband <= kBhBand ? nband : ncand;
    for (int t = warp; t < nwork; t += kBhWarps) {
      int jj, pos;
      if (nband <= kBhBand) { jj = band[t] >> 1; pos = band[t] & 1; }
      else { jj = t; pos = is_positive(t) ? 1 : 0; }
      const PairExact st = pair_exact<kVec>(ar, cand_row(P, jj), P.dim, P.metric, ca, lane);
      const double d32 = (double)(float)st.d;
      const bool better = pos ? (d32 > bd[1] || (d32 == bd[1] && jj < bi[1])) : (d32 < bd[0] || (d32 == bd[0] && jj < bi[0]));
      if (better || bi[pos] < 0) { bd[pos] = d32; bi[pos] = jj; bs[pos] = st; }
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 2; ++c) { s_bd[c][warp] = bd[c]; s_bi[c][warp] = bi[c]; s_bs[c][warp] = bs[c]; }
    }
    __syncthreads();
    BH_STAMP(3);
    if (tid == 0) {
      int ip = -1, in = -1, wp = 0, wn = 0;
      for (int w = 0; w < kBhWarps; ++w) {
        if (s_bi[1][w] >= 0 && (ip < 0 || s_bd[1][w] > s_bd[1][wp] || (s_bd[1][w] == s_bd[1][wp] && s_bi[1][w] < ip))) { ip = s_bi[1][w]; wp = w; }
        if (s_bi[0][w] >= 0 && (in < 0 || s_bd[0][w] < s_bd[0][wn] || (s_bd[0][w] == s_bd[0][wn] && s_bi[0][w] < in))) { in = s_bi[0][w]; wn = w; }
      }
      const PairExact sp = s_bs[1][wp], sn = s_bs[0][wn];
      float hinge = 0.f, w = 0.f;
      if (ip >= 0 && in >= 0) {
        const float arg = __fsub_rn(__fadd_rn(P.margin, (float)sp.d), (float)sn.d);
        hinge = fmaxf(arg, 0.f);
        w = arg >= 0.f ? 1.0f / (float)P.batch : 0.f;  // torch's clamp_min backward passes grad at equality
      }
      const float ma = sqrtf(qsq) > kCosineEps ? 1.f : 0.f;
      P.per_row[i] = hinge;
      P.weight[i] = w;
      P.sel[2 * i] = ip;
      P.sel[2 * i + 1] = in;
      P.pair_d[2 * i] = sp.d;
      P.pair_d[2 * i + 1] = sn.d;
      float* st = P.pair_stat + (size_t)i * 8;
      st[0] = ca; st[1] = ma; st[2] = sp.cx; st[3] = sp.s; st[4] = sp.mx; st[5] = sn.cx; st[6] = sn.s; st[7] = sn.mx;
      if (P.out_hard_index) {
        P.out_hard_index[2 * i] = ip;
        P.out_hard_index[2 * i + 1] = in;
      }
      s_sel[0] = sp; s_sel[1] = sn; s_seli[0] = ip; s_seli[1] = in; s_w = w;
    }
    __syncthreads();
    if (P.ga != nullptr) {
      const float w = s_w;
      const PairExact sp = s_sel[0], sn = s_sel[1];
      const float ma = sqrtf(qsq) > kCosineEps ? 1.f : 0.f;
      float4* out = reinterpret_cast<float4*>(P.ga + (size_t)i * P.dim);
      const int nvec = P.dim / 4;
      if (w == 0.f) {
        for (int v = tid; v < nvec; v += kBhThreads) out[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const float4* a4 = reinterpret_cast<const float4*>(ar);
        const float4* xp4 = reinterpret_cast<const float4*>(cand_row(P, s_seli[0]));
        const float4* xn4 = reinterpret_cast<const float4*>(cand_row(P, s_seli[1]));
        const float cp = sp.d > 0.0 ? (float)((double)w / sp.d) : 0.f;
        const float cn = sn.d > 0.0 ? (float)((double)w / sn.d) : 0.f;
#pragma unroll 4
        for (int v = tid; v < nvec; v += kBhThreads) {
          const float4 av4 = __ldg(a4 + v), p4 = __ldg(xp4 + v), n4 = __ldg(xn4 + v);
          const float av[4] = {av4.x, av4.y, av4.z, av4.w}, pv[4] = {p4.x, p4.y, p4.z, p4.w}, nv[4] = {n4.x, n4.y, n4.z, n4.w};
          float o[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (P.metric == SBIR_EUCLIDEAN) {
              const float u = __fadd_rn(__fsub_rn(av[c], pv[c]), kPairwiseEps) * cp;
              const float vv = __fadd_rn(__fsub_rn(av[c], nv[c]), kPairwiseEps) * cn;
              o[c] = u - vv;
            } else {
              const float ah = av[c] / ca;
              // d = 1 − s  →  ∂d/∂a = −(x̂ − s·â·[‖a‖>eps]) / ca
              const float gp_ = -(pv[c] / sp.cx - sp.s * ah * ma) / ca;
              const float gn_ = -(nv[c] / sn.cx - sn.s * ah * ma) / ca;
              o[c] = w * gp_ + (-w) * gn_;
            }
          }
          out[v] = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    __syncthreads();  // shared selection state is rewritten by the CTA's next anchor
    BH_STAMP(4);
  }
  __threadfence();
  grid.sync();
  BH_STAMP(6);

  // -------------------------------------- phase 3: candidate gradients + mean of the hinges ----
  if (P.timing && blockIdx.x == 0 && threadIdx.x == 0) P.timing[2] = globaltimer_ns();
  const int gwarp = blockIdx.x * kBhWarps + warp, nwarps = gridDim.x * kBhWarps;
  if (gwarp == nwarps - 1) {  // last warp of the grid: deterministic mean (fixed order, fp64)
    double acc = 0.0;
    for (int i0 = 0; i0 < P.batch; i0 += 32) {
      double v = (i0 + lane < P.batch) ? (double)__ldcg(P.per_row + i0 + lane) : 0.0;
      // fixed-shape butterfly, then added chunk by chunk: the same order on every run
      v = warp_sum(v);
      acc += v;
    }
    if (lane == 0) P.out_loss[0] = (float)(acc / (double)P.batch);
  }
  if (P.gp != nullptr || P.gn != nullptr) {
    // Work item = (candidate row, segment of 256 elements): a row that many anchors selected (a candidate
    // close to everyone) is spread over several warps instead of serialising on one.
    int* hits = reinterpret_cast<int*>(smem) + (size_t)warp * kBhHits;  // (anchor << 1 | side), ascending anchors
    const int nvec = P.dim / 4;
    const int nseg = (nvec + 63) / 64;
    const long long nitems = (long long)ncand * nseg;
    for (long long item = gwarp; item < nitems; item += nwarps) {
      const int j = (int)(item / nseg), seg = (int)(item - (long long)j * nseg);
      float* out = j < P.batch ? (P.gp ? P.gp + (size_t)j * P.dim : nullptr)
                               : (P.gn ? P.gn + (size_t)(j - P.batch) * P.dim : nullptr);
      if (out == nullptr) continue;
      float4* out4 = reinterpret_cast<float4*>(out);
      const float4* x4 = reinterpret_cast<const float4*>(cand_row(P, j));
      const int va = seg * 64 + lane, vb = va + 32;
      float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa;
      if (va < nvec) xa = __ldg(x4 + va);
      if (vb < nvec) xb = __ldg(x4 + vb);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
      int i_next = 0;
      do {
        // anchors (ascending) that selected row j — bit 0: as their positive, bit 1: as their negative;
        // 8 anchors per lane with their loads in flight, at most kBhHits hits per batch
        int nhits = 0;
        for (; i_next < P.batch && nhits + 512 <= kBhHits; i_next += 256) {
          int f[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = i_next + u * 32 + lane;
            f[u] = 0;
            if (i < P.batch) {
              const int2 sl = __ldcg(reinterpret_cast<const int2*>(P.sel) + i);
              const float w = __ldcg(P.weight + i);
              if (w != 0.f) f[u] = (sl.x == j ? 1 : 0) | (sl.y == j ? 2 : 0);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const unsigned m1 = __ballot_sync(kFullMask, f[u] != 0);
            if (m1 == 0) continue;
            const unsigned m2b = __ballot_sync(kFullMask, f[u] == 3);  // both sides: two entries
            const unsigned below = (1u << lane) - 1u;
            const int my_off = nhits + __popc(m1 & below) + __popc(m2b & below);
            if (f[u] & 1) hits[my_off] = ((i_next + u * 32 + lane) << 1);
            if (f[u] & 2) hits[my_off + (f[u] & 1)] = ((i_next + u * 32 + lane) << 1) | 1;
            nhits += __popc(m1) + __popc(m2b);
          }
        }
        __syncwarp();
        // groups of 4 hits with their loads in flight, accumulated in list order
        for (int h0 = 0; h0 < nhits; h0 += 4) {
          float4 aa[4], ab[4];
          float coef[4], cca[4], ccx[4], cs[4], cmx[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            aa[g] = make_float4(0.f, 0.f, 0.f, 0.f);
            ab[g] = aa[g];
            coef[g] = 0.f; cca[g] = 1.f; ccx[g] = 1.f; cs[g] = 0.f; cmx[g] = 0.f;
            if (h0 + g < nhits) {
              const int hv = hits[h0 + g];
              const int ii = hv >> 1, side = hv & 1;
              const float w = __ldcg(P.weight + ii);
              const float scale = side == 0 ? w : -w;
              const float4* a4 = reinterpret_cast<const float4*>(P.a + (size_t)ii * P.dim);
              if (va < nvec) aa[g] = __ldg(a4 + va);
              if (vb < nvec) ab[g] = __ldg(a4 + vb);
              if (P.metric == SBIR_EUCLIDEAN) {
                const double d = __ldcg(P.pair_d + 2 * ii + side);
                coef[g] = d > 0.0 ? (float)((double)scale / d) : 0.f;
              } else {
                const float* st = P.pair_stat + (size_t)ii * 8;
                coef[g] = scale;
                cca[g] = __ldcg(st + 0);
                ccx[g] = __ldcg(st + 2 + 3 * side); cs[g] = __ldcg(st + 3 + 3 * side); cmx[g] = __ldcg(st + 4 + 3 * side);
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (h0 + g >= nhits) break;
            const float av[8] = {aa[g].x, aa[g].y, aa[g].z, aa[g].w, ab[g].x, ab[g].y, ab[g].z, ab[g].w};
            if (P.metric == SBIR_EUCLIDEAN) {
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[e] += -(__fadd_rn(__fsub_rn(av[e], xv[e]), kPairwiseEps) * coef[g]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float ah = av[e] / cca[g], xh = xv[e] / ccx[g];
                acc[e] += coef[g] * (-(ah - cs[g] * xh * cmx[g]) / ccx[g]);
              }
            }
          }
        }
        __syncwarp();
      } while (i_next < P.batch);
      if (va < nvec) out4[va] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      if (vb < nvec) out4[vb] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
  if (P.timing && blockIdx.x == 0 && threadIdx.x == 0) P.timing[3] = globaltimer_ns();
  BH_STAMP(7);
}

struct BhLayout {
  int num_q_tiles, tiles_per_half, k_splits, kb_per_split, num_k_blocks, num_units, dot_rows, dot_cols;
  size_t off_dots, off_anorm, off_cnorm, off_per_row, off_weight, off_sel, off_pair_d, off_pair_stat, off_timing, total;
};

BhLayout bh_layout(int64_t batch, int64_t dim) {
  BhLayout L{};
  L.num_q_tiles = (int)((batch + kBhTileA - 1) / kBhTileA);
  L.tiles_per_half = (int)((batch + kBhTileC - 1) / kBhTileC);
  L.num_k_blocks = (int)((dim * 4 + kSwizzleBytes - 1) / kSwizzleBytes);
  // K split: enough units for every resident CTA when the batch is small, at least 4 k-blocks per unit
  const int tiles = L.num_q_tiles * 2 * L.tiles_per_half;
  int want = (256 + tiles - 1) / tiles;
  if (want > kBhMaxSplits) want = kBhMaxSplits;
  if (want > L.num_k_blocks / 4) want = L.num_k_blocks / 4;
  if (want < 1) want = 1;
  L.kb_per_split = (L.num_k_blocks + want - 1) / want;
  L.k_splits = (L.num_k_blocks + L.kb_per_split - 1) / L.kb_per_split;
  L.num_units = tiles * L.k_splits;
  L.dot_rows = L.num_q_tiles * kBhTileA;
  L.dot_cols = 2 * L.tiles_per_half * kBhTileC;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o = align_up(o + bytes, 256); return r; };
  L.off_dots = take((size_t)L.k_splits * L.dot_rows * L.dot_cols * sizeof(float));
  L.off_anorm = take((size_t)batch * sizeof(float));
  L.off_cnorm = take((size_t)2 * batch * sizeof(float));
  L.off_per_row = take((size_t)batch * sizeof(float));
  L.off_weight = take((size_t)batch * sizeof(float));
  L.off_sel = take((size_t)batch * 2 * sizeof(int));
  L.off_pair_d = take((size_t)batch * 2 * sizeof(double));
  L.off_pair_stat = take((size_t)batch * 8 * sizeof(float));
  L.off_timing = take((8 + 8 * 512) * sizeof(unsigned long long));  // phase stamps + per-CTA stage stamps (diagnostics)
  L.total = o;
  return L;
}

}  // namespace

size_t batch_hard_workspace_bytes(int64_t batch, int64_t dim) { return bh_layout(batch, dim).total; }

int launch_batch_hard(const float* a, const float* p, const float* n, int64_t batch, int64_t dim,
                      float margin, int metric, const int64_t* anchor_label, const int64_t* cand_label,
                      float* out_loss, int64_t* out_hard_index, float* ga, float* gp, float* gn,
                      void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (batch > (1 << 20) || dim > (1 << 20)) return SBIR_ERR_UNSUPPORTED;
  const BhLayout L = bh_layout(batch, dim);
  if (workspace == nullptr || workspace_bytes < L.total || reinterpret_cast<uintptr_t>(workspace) % 256 != 0)
    return SBIR_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(n)) % 16 != 0)
    return SBIR_ERR_UNSUPPORTED;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  BhParams P{};
  P.a = a; P.p = p; P.n = n;
  P.batch = (int)batch; P.dim = (int)dim; P.metric = metric;
  P.margin = margin;
  P.kappa = k1_kappa(SBIR_F32, dim) + (float)kBhMaxSplits * 1.1920929e-07f;  // + the fp32 sums over the K splits
  P.anchor_label = reinterpret_cast<const long long*>(anchor_label);
  P.cand_label = reinterpret_cast<const long long*>(cand_label);
  P.num_q_tiles = L.num_q_tiles; P.tiles_per_half = L.tiles_per_half;
  P.k_splits = L.k_splits; P.kb_per_split = L.kb_per_split; P.num_k_blocks = L.num_k_blocks; P.num_units = L.num_units;
  P.dot_rows = L.dot_rows; P.dot_cols = L.dot_cols;
  P.dots = reinterpret_cast<float*>(ws + L.off_dots);
  P.anorm = reinterpret_cast<float*>(ws + L.off_anorm);
  P.cnorm = reinterpret_cast<float*>(ws + L.off_cnorm);
  P.per_row = reinterpret_cast<float*>(ws + L.off_per_row);
  P.weight = reinterpret_cast<float*>(ws + L.off_weight);
  P.sel = reinterpret_cast<int*>(ws + L.off_sel);
  P.pair_d = reinterpret_cast<double*>(ws + L.off_pair_d);
  P.pair_stat = reinterpret_cast<float*>(ws + L.off_pair_stat);
#ifdef SBIR_DIAG
  P.timing = reinterpret_cast<unsigned long long*>(ws + L.off_timing);  // stage stamps: diagnostic builds only (SBIR_BUILD_DIAG=1)
#else
  P.timing = nullptr;
#endif
  P.out_loss = out_loss;
  P.out_hard_index = reinterpret_cast<long long*>(out_hard_index);
  P.ga = ga; P.gp = gp; P.gn = gn;

  CUtensorMap ta, tp, tn;
  SBIR_TRY(make_tmap(&ta, a, batch, dim, SBIR_F32, kBhTileA));
  SBIR_TRY(make_tmap(&tp, p, batch, dim, SBIR_F32, kBhTileC));
  SBIR_TRY(make_tmap(&tn, n, batch, dim, SBIR_F32, kBhTileC));

  // Cooperative launch: every CTA must be resident (the phases are separated by grid barriers): two
  // CTAs per SM (81 KB of shared memory each), fewer if the device cannot hold that many.
  auto kern = batch_hard_fused_kernel<true>;
  SBIR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBhSmemBytes));
  int dev = 0, num_sms = 0, per_sm = 0, coop = 0;
  SBIR_CUDA_TRY(cudaGetDevice(&dev));
  SBIR_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return SBIR_ERR_UNSUPPORTED;
  SBIR_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  SBIR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBhThreads, kBhSmemBytes));
  if (per_sm < 1) return SBIR_ERR_UNSUPPORTED;
  // a CTA per anchor (phase 2) / a warp per (candidate row, segment) item (phase 3), never more CTAs than can be co-resident
  int64_t want = batch > (2 * batch * ((dim + 255) / 256) + kBhWarps - 1) / kBhWarps ? batch : (2 * batch * ((dim + 255) / 256) + kBhWarps - 1) / kBhWarps;
  if (want < L.num_units) want = L.num_units;
  const int64_t resident = (int64_t)num_sms * per_sm;
  int grid = (int)(want < resident ? want : resident);
  if (grid < 1) grid = 1;
  void* args[] = {&ta, &tp, &tn, &P};
  const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3((unsigned)grid), dim3(kBhThreads),
                                                    args, kBhSmemBytes, st);
  if (e != cudaSuccess) {
    set_last_cuda_error((int)e);
    return SBIR_ERR_CUDA;
  }
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

}  // namespace sbir
