// batch_hard.cu — K3: batch-hard triplet mining + margin loss + gradients (H8, a north_star
// extension; the reference forms triplets in its datasets, data_preparation.py:67-69,214-222,
// and only evaluates nn.TripletMarginLoss / TripletMarginWithDistanceLoss on them,
// train.py:164-175).  Mining runs on the same tcgen05 tiles as retrieval (dist_topk_kernel.cuh in
// kModeHard); this file selects across tiles, re-scores the selected pairs exactly and
// scatters gradients deterministically (no floating-point atomics).
#include "common.cuh"
#include "kernels.h"

namespace sbir {

namespace {

constexpr int kBhThreads = 128;

__device__ __forceinline__ double bh_block_sum(double v, double* red4) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red4[threadIdx.x >> 5] = v;
  __syncthreads();
  return red4[0] + red4[1] + red4[2] + red4[3];
}

struct PairStats {
  double d;         // exact distance
  float ca, cx, s;  // cosine: clamped norms and similarity
  float ma, mx;     // cosine: 1 when the norm is above eps (gradient flows through the norm)
};

// Exact distance of (a_row, x_row) by the whole block, plus what the gradient needs.
__device__ PairStats pair_stats(const float* __restrict__ ar, const float* __restrict__ xr, int dim,
                                int metric, double* red4) {
  PairStats st{};
  const int t = threadIdx.x;
  if (metric == SBIR_EUCLIDEAN) {
    double s = 0.0;
    for (int i = t; i < dim; i += kBhThreads) {
      const float u = __fadd_rn(__fsub_rn(ar[i], xr[i]), kPairwiseEps);
      s += (double)u * (double)u;
    }
    st.d = sqrt(bh_block_sum(s, red4));
  } else {
    double qa = 0.0, qx = 0.0;
    for (int i = t; i < dim; i += kBhThreads) {
      qa += (double)ar[i] * (double)ar[i];
      qx += (double)xr[i] * (double)xr[i];
    }
    qa = bh_block_sum(qa, red4);
    qx = bh_block_sum(qx, red4);
    st.ca = clamped_norm(qa);
    st.cx = clamped_norm(qx);
    st.ma = (float)sqrt(qa) > kCosineEps ? 1.f : 0.f;
    st.mx = (float)sqrt(qx) > kCosineEps ? 1.f : 0.f;
    double s = 0.0;
    for (int i = t; i < dim; i += kBhThreads)
      s += (double)__fmul_rn(__fdiv_rn(ar[i], st.ca), __fdiv_rn(xr[i], st.cx));
    s = bh_block_sum(s, red4);
    st.s = (float)s;
    st.d = 1.0 - s;
  }
  return st;
}

// One block per anchor: reduce the per-tile hardest candidates, re-score exactly, hinge.
__global__ void __launch_bounds__(kBhThreads) bh_select_kernel(
    const float* __restrict__ a, const float* __restrict__ x, int batch, int dim, int metric,
    float margin, const float* __restrict__ hard_val, const int32_t* __restrict__ hard_idx,
    int num_slots, int slot_stride_rows, int halves, float inv_batch, float* __restrict__ per_row,
    float* __restrict__ weight, long long* __restrict__ sel, long long* __restrict__ out_hard_index) {
  __shared__ double red[4];
  __shared__ int s_hp, s_hn;
  const int i = blockIdx.x;
  if (threadIdx.x == 0) {
    float hp = -INFINITY, hn = INFINITY;
    int hpi = -1, hni = -1;
    for (int s = 0; s < num_slots; ++s) {
      for (int h = 0; h < halves; ++h) {
        const size_t o = ((size_t)s * slot_stride_rows + i) * halves + h;
        const float vp = hard_val[o * 2], vn = hard_val[o * 2 + 1];
        const int ip = hard_idx[o * 2], in = hard_idx[o * 2 + 1];
        if (ip >= 0 && (vp > hp || (vp == hp && ip < hpi) || hpi < 0)) { hp = vp; hpi = ip; }
        if (in >= 0 && (vn < hn || (vn == hn && in < hni) || hni < 0)) { hn = vn; hni = in; }
      }
    }
    s_hp = hpi;
    s_hn = hni;
  }
  __syncthreads();
  const int hpi = s_hp, hni = s_hn;
  float hinge = 0.f, w = 0.f;
  if (hpi >= 0 && hni >= 0) {
    const PairStats sp = pair_stats(a + (size_t)i * dim, x + (size_t)hpi * dim, dim, metric, red);
    const PairStats sn = pair_stats(a + (size_t)i * dim, x + (size_t)hni * dim, dim, metric, red);
    const float arg = __fsub_rn(__fadd_rn(margin, (float)sp.d), (float)sn.d);
    hinge = fmaxf(arg, 0.f);
    w = arg >= 0.f ? inv_batch : 0.f;
  }
  if (threadIdx.x == 0) {
    per_row[i] = hinge;
    weight[i] = w;
    sel[2 * i] = hpi;
    sel[2 * i + 1] = hni;
    if (out_hard_index) {
      out_hard_index[2 * i] = hpi;
      out_hard_index[2 * i + 1] = hni;
    }
  }
}

// ∂d(a,x)/∂a (sign +1) or ∂d(a,x)/∂x (side 1), scaled by `scale`, accumulated into out[].
__device__ __forceinline__ void accumulate_pair_grad(const float* __restrict__ ar, const float* __restrict__ xr,
                                                     int dim, int metric, const PairStats& st, float scale,
                                                     bool wrt_x, float* __restrict__ out) {
  const int t = threadIdx.x;
  if (metric == SBIR_EUCLIDEAN) {
    const float c = st.d > 0.0 ? (float)((double)scale / st.d) : 0.f;
    for (int i = t; i < dim; i += kBhThreads) {
      const float u = __fadd_rn(__fsub_rn(ar[i], xr[i]), kPairwiseEps) * c;
      out[i] += wrt_x ? -u : u;
    }
  } else {
    for (int i = t; i < dim; i += kBhThreads) {
      const float ah = ar[i] / st.ca, xh = xr[i] / st.cx;
      // d = 1 - s  →  ∂d/∂a = -(x̂ - s·â·[‖a‖>eps]) / ca
      const float gval = wrt_x ? -(ah - st.s * xh * st.mx) / st.cx : -(xh - st.s * ah * st.ma) / st.ca;
      out[i] += scale * gval;
    }
  }
}

__global__ void __launch_bounds__(kBhThreads) bh_grad_anchor_kernel(
    const float* __restrict__ a, const float* __restrict__ x, int dim, int metric,
    const float* __restrict__ weight, const long long* __restrict__ sel, float* __restrict__ ga) {
  __shared__ double red[4];
  const int i = blockIdx.x;
  float* out = ga + (size_t)i * dim;
  for (int e = threadIdx.x; e < dim; e += kBhThreads) out[e] = 0.f;
  const float w = weight[i];
  if (w == 0.f) return;
  const float* ar = a + (size_t)i * dim;
  const float* xp = x + (size_t)sel[2 * i] * dim;
  const float* xn = x + (size_t)sel[2 * i + 1] * dim;
  const PairStats sp = pair_stats(ar, xp, dim, metric, red);
  accumulate_pair_grad(ar, xp, dim, metric, sp, w, false, out);
  const PairStats sn = pair_stats(ar, xn, dim, metric, red);
  accumulate_pair_grad(ar, xn, dim, metric, sn, -w, false, out);
}

// One block per candidate row j: walk the anchors in index order and accumulate the
// contributions of those that selected j (fixed order → bitwise reproducible).  The selections
// are scanned 128 anchors at a time (one per thread, flags through shared memory, so the walk
// itself is a handful of broadcast loads) and the gradient row is accumulated in shared memory
// (`acc`, dim floats; global memory when the row does not fit) and written once.
__global__ void __launch_bounds__(kBhThreads) bh_grad_cand_kernel(
    const float* __restrict__ a, const float* __restrict__ x, int batch, int dim, int metric,
    const float* __restrict__ weight, const long long* __restrict__ sel, float* __restrict__ gp,
    float* __restrict__ gn, int acc_in_smem) {
  extern __shared__ float bh_acc[];
  __shared__ double red[4];
  __shared__ unsigned char s_flag[kBhThreads];
  const int j = blockIdx.x;
  float* out = j < batch ? (gp ? gp + (size_t)j * dim : nullptr) : (gn ? gn + (size_t)(j - batch) * dim : nullptr);
  if (out == nullptr) return;
  float* acc = acc_in_smem ? bh_acc : out;
  for (int e = threadIdx.x; e < dim; e += kBhThreads) acc[e] = 0.f;
  const float* xr = x + (size_t)j * dim;
  for (int base = 0; base < batch; base += kBhThreads) {
    const int i_mine = base + threadIdx.x;
    unsigned char f = 0;
    if (i_mine < batch && weight[i_mine] != 0.f) f = (sel[2 * i_mine] == j ? 1 : 0) | (sel[2 * i_mine + 1] == j ? 2 : 0);
    __syncthreads();  // previous chunk's flags fully consumed (and acc zeroed on the first pass)
    s_flag[threadIdx.x] = f;
    __syncthreads();
    for (int u = 0; u < kBhThreads; ++u) {
      const unsigned char fu = s_flag[u];  // same value for every thread: uniform branch
      if (fu == 0) continue;
      const int i = base + u;
      const float w = weight[i];
      const float* ar = a + (size_t)i * dim;
      const PairStats st = pair_stats(ar, xr, dim, metric, red);
      if (fu & 1) accumulate_pair_grad(ar, xr, dim, metric, st, w, true, acc);
      if (fu & 2) accumulate_pair_grad(ar, xr, dim, metric, st, -w, true, acc);
    }
  }
  if (acc_in_smem) {
    for (int e = threadIdx.x; e < dim; e += kBhThreads) out[e] = acc[e];  // each thread wrote these itself
  }
}

__global__ void __launch_bounds__(256) bh_mean_kernel(const float* __restrict__ per_row, int rows,
                                                      float* __restrict__ out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < rows; i += 256) acc += (double)per_row[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] / (double)rows);
}

struct BhLayout {
  K1Plan plan;
  size_t off_x, off_gvec, off_hard_val, off_hard_idx, off_per_row, off_weight, off_sel, off_counter, total;
};

BhLayout bh_layout(int64_t batch, int64_t dim) {
  BhLayout L{};
  L.plan = make_k1_plan(batch, 2 * batch, dim, 1, SBIR_F32, 148);
  // single-CTA tiles, one unit per gallery tile: the epilogue state is per (tile, row)
  L.plan.pair = 1;
  L.plan.tiles_per_split = 1;
  L.plan.num_splits = L.plan.num_g_tiles;
  L.plan.band_q = L.plan.num_q_tiles;
  L.plan.num_chunks = 1;
  L.plan.tiles_per_chunk = 1;
  L.plan.part_fastest = 0;
  L.plan.num_units = L.plan.num_q_tiles * L.plan.num_splits;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o = align_up(o + bytes, 256); return r; };
  L.off_x = take((size_t)2 * batch * dim * sizeof(float));
  L.off_gvec = take((size_t)L.plan.num_g_tiles * kTileG * sizeof(float));
  const size_t hard_elems = (size_t)L.plan.num_splits * L.plan.q_tile_stride * kTileQ * L.plan.lists_per_row * 2;
  L.off_hard_val = take(hard_elems * sizeof(float));
  L.off_hard_idx = take(hard_elems * sizeof(int32_t));
  L.off_per_row = take((size_t)batch * sizeof(float));
  L.off_weight = take((size_t)batch * sizeof(float));
  L.off_sel = take((size_t)batch * 2 * sizeof(long long));
  L.off_counter = take(256);
  L.total = o;
  return L;
}

}  // namespace

size_t batch_hard_workspace_bytes(int64_t batch, int64_t dim) { return bh_layout(batch, dim).total; }

int launch_batch_hard(const float* a, const float* p, const float* n, int64_t batch, int64_t dim,
                      float margin, int metric, const int64_t* anchor_label, const int64_t* cand_label,
                      float* out_loss, int64_t* out_hard_index, float* ga, float* gp, float* gn,
                      void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const BhLayout L = bh_layout(batch, dim);
  if (workspace == nullptr || workspace_bytes < L.total || reinterpret_cast<uintptr_t>(workspace) % 256 != 0)
    return SBIR_ERR_WORKSPACE;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(ws + L.off_x);
  float* gvec = reinterpret_cast<float*>(ws + L.off_gvec);
  float* hard_val = reinterpret_cast<float*>(ws + L.off_hard_val);
  int32_t* hard_idx = reinterpret_cast<int32_t*>(ws + L.off_hard_idx);
  float* per_row = reinterpret_cast<float*>(ws + L.off_per_row);
  float* weight = reinterpret_cast<float*>(ws + L.off_weight);
  long long* sel = reinterpret_cast<long long*>(ws + L.off_sel);

  const size_t half_bytes = (size_t)batch * dim * sizeof(float);
  SBIR_CUDA_TRY(cudaMemcpyAsync(x, p, half_bytes, cudaMemcpyDeviceToDevice, st));
  SBIR_CUDA_TRY(cudaMemcpyAsync(x + (size_t)batch * dim, n, half_bytes, cudaMemcpyDeviceToDevice, st));
  const int64_t padded = (int64_t)L.plan.num_g_tiles * kTileG;
  SBIR_TRY(launch_row_norm(x, 2 * batch, padded, dim, SBIR_F32, metric == SBIR_EUCLIDEAN ? 0 : 1,
                           metric == SBIR_EUCLIDEAN ? INFINITY : nanf(""), gvec, nullptr, st));
  K1Args ka{};
  ka.q = a; ka.g = x;
  ka.num_q = batch; ka.num_g = 2 * batch; ka.dim = dim;
  ka.dtype = SBIR_F32; ka.metric = metric; ka.mode = kModeHard;
  ka.gvec = gvec;
  ka.row_label = anchor_label; ka.col_label = cand_label;
  ka.hard_val = hard_val; ka.hard_idx = hard_idx;
  ka.unit_counter = reinterpret_cast<uint32_t*>(ws + L.off_counter);
  SBIR_CUDA_TRY(cudaMemsetAsync(ws + L.off_counter, 0, 256, st));
  SBIR_TRY(launch_k1(ka, L.plan, st));

  bh_select_kernel<<<(unsigned)batch, kBhThreads, 0, st>>>(
      a, x, (int)batch, (int)dim, metric, margin, hard_val, hard_idx, L.plan.num_splits,
      L.plan.q_tile_stride * kTileQ, L.plan.lists_per_row, 1.0f / (float)batch, per_row, weight, sel,
      reinterpret_cast<long long*>(out_hard_index));
  SBIR_CHECK_LAUNCH();
  bh_mean_kernel<<<1, 256, 0, st>>>(per_row, (int)batch, out_loss);
  SBIR_CHECK_LAUNCH();
  if (ga) {
    bh_grad_anchor_kernel<<<(unsigned)batch, kBhThreads, 0, st>>>(a, x, (int)dim, metric, weight, sel, ga);
    SBIR_CHECK_LAUNCH();
  }
  if (gp || gn) {
    const int acc_in_smem = (size_t)dim * sizeof(float) <= 40 * 1024 ? 1 : 0;
    bh_grad_cand_kernel<<<(unsigned)(2 * batch), kBhThreads, acc_in_smem ? (size_t)dim * sizeof(float) : 0, st>>>(
        a, x, (int)batch, (int)dim, metric, weight, sel, gp, gn, acc_in_smem);
    SBIR_CHECK_LAUNCH();
  }
  return SBIR_OK;
}

}  // namespace sbir
