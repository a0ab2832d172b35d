// batch_hard.cu — K3: batch-hard triplet mining + margin loss + gradients as ONE cooperative kernel
// (H8, a north_star extension; the reference forms triplets in its datasets,
// data_preparation.py:67-69,214-222, and only evaluates nn.TripletMarginLoss /
// TripletMarginWithDistanceLoss on them, train.py:164-175).
//
// Candidates X = cat(p, n) are never concatenated: p and n are read in place through two tensor
// maps.  The kernel runs three phases separated by grid-wide barriers (cooperative launch):
//   1. MINING TILES  tcgen05 (kind::tf32) tiles of A·Xᵀ — 128 anchors × 32 candidates per unit, the
//      K dimension split over several units so that a 256-anchor batch still spreads over ~128 SMs —
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
constexpr int kBhStages = 8;
constexpr int kBhStageA = kBhTileA * kSwizzleBytes;  // 16 KB
constexpr int kBhStageC = kBhTileC * kSwizzleBytes;  // 4 KB
constexpr int kBhThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kBhWarps = kBhThreads / 32;
constexpr int kBhSmemBytes = 1024 + kBhStages * (kBhStageA + kBhStageC) + 256;
constexpr int kBhTmemCols = 64;  // two accumulators of 32 columns
constexpr int kBhMaxSplits = 8;

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

template <bool kVec>
__global__ void __launch_bounds__(kBhThreads, 1)
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
  const int gwarp = blockIdx.x * kBhWarps + warp, nwarps = gridDim.x * kBhWarps;
  const int ncand = 2 * P.batch;
  const size_t split_stride = (size_t)P.dot_rows * P.dot_cols;
  for (int i = gwarp; i < P.batch; i += nwarps) {
    const float* ar = P.a + (size_t)i * P.dim;
    const float qsq = __ldcg(P.anorm + i);
    const float ca = clamped_norm((double)qsq);  // fp32 norm as torch returns it, clamped (cosine)
    const long long my_label = P.anchor_label ? P.anchor_label[i] : 0;
    const float* drow = P.dots + (size_t)i * P.dot_cols;
    auto e_of = [&](int j, float& cn_out) -> float {
      const int col = cand_col(P, j);
      float dot = 0.f;
      for (int s = 0; s < P.k_splits; ++s) dot += __ldcg(drow + s * split_stride + col);
      const float cn = __ldcg(P.cnorm + j);
      cn_out = cn;
      return P.metric == SBIR_EUCLIDEAN ? fmaf(-2.f, dot, cn) : -dot / fmaxf(sqrtf(cn), kCosineEps);
    };
    auto is_positive = [&](int j) -> bool { return P.anchor_label ? (P.cand_label[j] == my_label) : (j == i); };
    // approximate extrema over the tiles
    float ap = -INFINITY, an = INFINITY, gmax = 0.f;
    for (int j = lane; j < ncand; j += 32) {
      float cn;
      const float e = e_of(j, cn);
      gmax = fmaxf(gmax, cn);
      if (is_positive(j)) ap = fmaxf(ap, e);
      else an = fminf(an, e);
    }
    ap = warp_max(ap);
    an = -warp_max(-an);
    gmax = warp_max(gmax);
    const float m2 = (float)(2.0 * e_margin(P.metric, qsq, gmax, P.kappa, P.dim));
    // every candidate inside the error band of the approximate extremum is re-scored exactly
    double bp = -1.0, bn = INFINITY;  // exact extrema in the canonical order (fp32 distance, index)
    int ip = -1, in = -1;
    PairExact sp{}, sn{};
    for (int j0 = 0; j0 < ncand; j0 += 32) {
      const int j = j0 + lane;
      bool pos_hit = false, neg_hit = false;
      if (j < ncand) {
        float cn;
        const float e = e_of(j, cn);
        if (is_positive(j)) pos_hit = e >= ap - m2;
        else neg_hit = e <= an + m2;
      }
      unsigned mask = __ballot_sync(kFullMask, pos_hit || neg_hit);
      const unsigned pmask = __ballot_sync(kFullMask, pos_hit);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const int jj = j0 + b;
        const PairExact st = pair_exact<kVec>(ar, cand_row(P, jj), P.dim, P.metric, ca, lane);
        const double d32 = (double)(float)st.d;
        if ((pmask >> b) & 1u) {
          if (d32 > bp) { bp = d32; ip = jj; sp = st; }   // ascending j: the first maximum wins ties
        } else {
          if (d32 < bn) { bn = d32; in = jj; sn = st; }
        }
      }
    }
    float hinge = 0.f, w = 0.f;
    if (ip >= 0 && in >= 0) {
      const float arg = __fsub_rn(__fadd_rn(P.margin, (float)sp.d), (float)sn.d);
      hinge = fmaxf(arg, 0.f);
      w = arg >= 0.f ? 1.0f / (float)P.batch : 0.f;  // torch's clamp_min backward passes grad at equality
    }
    const float ma = sqrtf(qsq) > kCosineEps ? 1.f : 0.f;
    if (lane == 0) {
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
    }
    if (P.ga != nullptr) {
      float* out = P.ga + (size_t)i * P.dim;
      if (w == 0.f) {
        for (int e = lane; e < P.dim; e += 32) out[e] = 0.f;
      } else {
        const float* xp = cand_row(P, ip);
        const float* xn = cand_row(P, in);
        if (P.metric == SBIR_EUCLIDEAN) {
          const float cp = sp.d > 0.0 ? (float)((double)w / sp.d) : 0.f;
          const float cn = sn.d > 0.0 ? (float)((double)w / sn.d) : 0.f;
          for (int e = lane; e < P.dim; e += 32) {
            const float av = ar[e];
            const float u = __fadd_rn(__fsub_rn(av, xp[e]), kPairwiseEps) * cp;
            const float v = __fadd_rn(__fsub_rn(av, xn[e]), kPairwiseEps) * cn;
            out[e] = u - v;
          }
        } else {
          for (int e = lane; e < P.dim; e += 32) {
            const float ah = ar[e] / ca;
            // d = 1 − s  →  ∂d/∂a = −(x̂ − s·â·[‖a‖>eps]) / ca
            const float gp_ = -(xp[e] / sp.cx - sp.s * ah * ma) / ca;
            const float gn_ = -(xn[e] / sn.cx - sn.s * ah * ma) / ca;
            out[e] = w * gp_ + (-w) * gn_;
          }
        }
      }
    }
  }
  __threadfence();
  grid.sync();

  // -------------------------------------- phase 3: candidate gradients + mean of the hinges ----
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
  if (P.gp == nullptr && P.gn == nullptr) return;
  for (int j = gwarp; j < ncand; j += nwarps) {
    float* out = j < P.batch ? (P.gp ? P.gp + (size_t)j * P.dim : nullptr)
                             : (P.gn ? P.gn + (size_t)(j - P.batch) * P.dim : nullptr);
    if (out == nullptr) continue;
    const float* xr = cand_row(P, j);
    // which anchors selected row j (bit 0: as their positive, bit 1: as their negative)
    bool any = false;
    for (int i0 = 0; i0 < P.batch && !any; i0 += 32) {
      const int i = i0 + lane;
      const bool f = i < P.batch && __ldcg(P.weight + i) != 0.f && (__ldcg(P.sel + 2 * i) == j || __ldcg(P.sel + 2 * i + 1) == j);
      any = __any_sync(kFullMask, f);
    }
    if (!any) {
      for (int e = lane; e < P.dim; e += 32) out[e] = 0.f;
      continue;
    }
    // segments of 32 lanes × 8 elements, anchors walked in index order inside every segment
    for (int e0 = 0; e0 < P.dim; e0 += 256) {
      float acc[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) acc[v] = 0.f;
      for (int i0 = 0; i0 < P.batch; i0 += 32) {
        const int i = i0 + lane;
        int f = 0;
        if (i < P.batch && __ldcg(P.weight + i) != 0.f)
          f = (__ldcg(P.sel + 2 * i) == j ? 1 : 0) | (__ldcg(P.sel + 2 * i + 1) == j ? 2 : 0);
        unsigned mask = __ballot_sync(kFullMask, f != 0);
        while (mask) {
          const int b = __ffs(mask) - 1;
          mask &= mask - 1;
          const int ii = i0 + b;
          const int fb = __shfl_sync(kFullMask, f, b);
          const float w = __ldcg(P.weight + ii);
          const float* ar = P.a + (size_t)ii * P.dim;
          const float* st = P.pair_stat + (size_t)ii * 8;
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            if (!(fb & (1 << side))) continue;
            const float scale = side == 0 ? w : -w;
            if (P.metric == SBIR_EUCLIDEAN) {
              const double d = __ldcg(P.pair_d + 2 * ii + side);
              const float c = d > 0.0 ? (float)((double)scale / d) : 0.f;
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                const int e = e0 + v * 32 + lane;
                if (e < P.dim) acc[v] += -(__fadd_rn(__fsub_rn(ar[e], xr[e]), kPairwiseEps) * c);
              }
            } else {
              const float ca = __ldcg(st + 0);
              const float cx = __ldcg(st + 2 + 3 * side), s = __ldcg(st + 3 + 3 * side), mx = __ldcg(st + 4 + 3 * side);
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                const int e = e0 + v * 32 + lane;
                if (e < P.dim) {
                  const float ah = ar[e] / ca, xh = xr[e] / cx;
                  acc[v] += scale * (-(ah - s * xh * mx) / cx);
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const int e = e0 + v * 32 + lane;
        if (e < P.dim) out[e] = acc[v];
      }
    }
  }
}

struct BhLayout {
  int num_q_tiles, tiles_per_half, k_splits, kb_per_split, num_k_blocks, num_units, dot_rows, dot_cols;
  size_t off_dots, off_anorm, off_cnorm, off_per_row, off_weight, off_sel, off_pair_d, off_pair_stat, total;
};

BhLayout bh_layout(int64_t batch, int64_t dim) {
  BhLayout L{};
  L.num_q_tiles = (int)((batch + kBhTileA - 1) / kBhTileA);
  L.tiles_per_half = (int)((batch + kBhTileC - 1) / kBhTileC);
  L.num_k_blocks = (int)((dim * 4 + kSwizzleBytes - 1) / kSwizzleBytes);
  // K split: enough units for ~128 SMs when the batch is small, at least 4 k-blocks per unit
  const int tiles = L.num_q_tiles * 2 * L.tiles_per_half;
  int want = (128 + tiles - 1) / tiles;
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
  P.out_loss = out_loss;
  P.out_hard_index = reinterpret_cast<long long*>(out_hard_index);
  P.ga = ga; P.gp = gp; P.gn = gn;

  CUtensorMap ta, tp, tn;
  SBIR_TRY(make_tmap(&ta, a, batch, dim, SBIR_F32, kBhTileA));
  SBIR_TRY(make_tmap(&tp, p, batch, dim, SBIR_F32, kBhTileC));
  SBIR_TRY(make_tmap(&tn, n, batch, dim, SBIR_F32, kBhTileC));

  // Cooperative launch: every CTA must be resident (the phases are separated by grid barriers).  One
  // CTA per SM (164 KB of shared memory); fewer if the device cannot hold that many.
  auto kern = batch_hard_fused_kernel<true>;
  SBIR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBhSmemBytes));
  int dev = 0, num_sms = 0, per_sm = 0, coop = 0;
  SBIR_CUDA_TRY(cudaGetDevice(&dev));
  SBIR_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return SBIR_ERR_UNSUPPORTED;
  SBIR_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  SBIR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBhThreads, kBhSmemBytes));
  if (per_sm < 1) return SBIR_ERR_UNSUPPORTED;
  // enough warps for one per anchor / candidate row, never more CTAs than can be co-resident
  int64_t want = (3 * batch + kBhWarps - 1) / kBhWarps;
  if (want < L.num_units) want = L.num_units;
  int grid = (int)(want < num_sms ? want : num_sms);
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
