// finalize.cu — everything after the tensor-core tiles:
//   * merge of the per-split candidate lists + EXACT re-scoring with the reference formula
//     (utils.py:31-42) + certificate that the selection is provably the exact top-k,
//   * brute-force exact top-k for queries the certificate rejects,
//   * rank of the positive = count(d < d_pos) (inference.py:49-52) from the fused counters
//     plus exact resolution of the uncertain band,
//   * K4 merge of per-shard top-k lists (after the all-gather), retrieval metrics (H5).
#include <algorithm>
#include <climits>

#include "common.cuh"
#include "kernels.h"

namespace sbir {

namespace {

template <typename Key>
__device__ __forceinline__ void bitonic_sort_smem(Key* key, int32_t* idx, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int pos = ((i / stride) * (stride << 1)) + (i % stride);
        const int par = pos + stride;
        const bool asc = (pos & size) == 0;
        const Key ka = key[pos], kb = key[par];
        const int32_t ia = idx[pos], ib = idx[par];
        const bool a_first = ka < kb || (ka == kb && ia <= ib);
        if (a_first != asc) {
          key[pos] = kb; key[par] = ka;
          idx[pos] = ib; idx[par] = ia;
        }
      }
    }
  }
  __syncthreads();
}

constexpr int kFinThreads = 128;

// bound on |approximate e − exact e| for query q in the pass described by p (FinParams / RankParams)
template <typename Params>
__device__ __forceinline__ double query_margin(const Params& p, int q) {
  return e_margin(p.metric, p.qsq[q], p.gsq_max[0], p.kappa, p.dim, p.q_res ? p.q_res[q] : 0.f, p.g_res ? p.g_res[0] : 0.f,
                  p.g_res ? p.g_res[1] : 0.f);
}

struct FinParams {
  const void* q;
  const void* g;
  int num_q, num_g, dim, metric, k;
  long long index_offset;
  const float* cand_val;
  const int32_t* cand_idx;
  int cap, lists_per_row, q_tile_stride, num_splits, m_pow2;
  const float* qsq;
  const float* gsq_max;
  float kappa;
  const float* q_res;   // [num_q] ‖q − bf16(q)‖ or NULL (bf16 selection of fp32 embeddings, see e_margin)
  const float* g_res;   // [2] max ‖g − bf16(g)‖ absolute / relative, or NULL
  float* out_dist;
  long long* out_index;
  int32_t* uncertified;
  int32_t* flags;
  const int32_t* gate;
};

template <typename T, bool kVec>
__global__ void __launch_bounds__(kFinThreads) finalize_topk_kernel(const FinParams p) {
  if (p.gate != nullptr && *p.gate == 0) return;
  extern __shared__ uint8_t fin_smem[];
  float* sv = reinterpret_cast<float*>(fin_smem);
  int32_t* si = reinterpret_cast<int32_t*>(sv + p.m_pow2);
  __shared__ double ex[128];
  __shared__ int32_t exi[128];
  __shared__ int s_nfinite;

  const int q = blockIdx.x;
  const int q_tile = q / kTileQ, row = q % kTileQ;
  const int m_tot = p.num_splits * p.lists_per_row * p.cap;
  if (threadIdx.x == 0) s_nfinite = 0;
  __syncthreads();
  int local_finite = 0;
  for (int i = threadIdx.x; i < p.m_pow2; i += kFinThreads) {
    float v = INFINITY;
    int32_t ix = INT_MAX;
    if (i < m_tot) {
      const int l = i / p.cap, pp = i % p.cap;
      const int split = l / p.lists_per_row, h = l % p.lists_per_row;
      const size_t slot = ((size_t)split * p.q_tile_stride + q_tile) * p.lists_per_row + h;
      const size_t addr = (slot * p.cap + pp) * kTileQ + row;
      const float cv = p.cand_val[addr];
      if (cv < INFINITY) {  // slots never filled keep +inf (their index is unspecified)
        v = cv;
        ix = p.cand_idx[addr];
        ++local_finite;
      }
    }
    sv[i] = v;
    si[i] = ix;
  }
  if (local_finite) atomicAdd(&s_nfinite, local_finite);
  __syncthreads();
  bool selected = false;
  if (p.m_pow2 > 128 && p.m_pow2 <= 1024) {
    // Many partitions (few query tiles: the reference's own 1k x 10k evaluation has 14): only the best `cap` of the
    // m_tot candidates matter, so instead of sorting them all (45 block-wide barrier steps for 512 entries) warp 0 finds
    // the cap-th smallest VALUE by a 32-step radix select on keys held in registers, the block compacts the winners to
    // the front, and only those are sorted.  If equal values straddle the cut (the (value, index) order would have to
    // pick among them) the full sort below decides instead — the result is always the prefix the full sort produces.
    __shared__ uint32_t s_kth;
    __shared__ int s_less, s_equal, s_fill;
    const int want = s_nfinite < p.cap ? s_nfinite : p.cap;
    auto ord_of = [&](int i) -> uint32_t {
      const int32_t b = __float_as_int(sv[i]);
      return (uint32_t)(b >= 0 ? b : b ^ 0x7fffffff) ^ 0x80000000u;   // monotone float -> unsigned
    };
    if (threadIdx.x < 32) {
      uint32_t key[32];   // m_pow2 / 32 <= 32 keys per lane
#pragma unroll
      for (int u = 0; u < 32; ++u) key[u] = (u * 32 + (int)threadIdx.x < p.m_pow2) ? ord_of(u * 32 + threadIdx.x) : 0xffffffffu;
      uint32_t prefix = 0;
      int remaining = want;
      for (int bit = 31; bit >= 0 && want > 0; --bit) {
        const uint32_t mask_hi = bit == 31 ? 0u : (~0u << (bit + 1));
        int zeros = 0;
#pragma unroll
        for (int u = 0; u < 32; ++u) zeros += ((key[u] & mask_hi) == prefix && ((key[u] >> bit) & 1u) == 0u) ? 1 : 0;
        zeros = __reduce_add_sync(kFullMask, zeros);
        if (remaining > zeros) { remaining -= zeros; prefix |= (1u << bit); }
      }
      int less = 0, equal = 0;
#pragma unroll
      for (int u = 0; u < 32; ++u) { less += key[u] < prefix ? 1 : 0; equal += key[u] == prefix ? 1 : 0; }
      less = __reduce_add_sync(kFullMask, less);
      equal = __reduce_add_sync(kFullMask, equal);
      if (threadIdx.x == 0) { s_kth = prefix; s_less = less; s_equal = equal; s_fill = 0; }
    }
    __syncthreads();
    selected = want > 0 && s_less + s_equal == want;   // no tie across the cut (block-uniform)
    if (selected) {
      float* cv = sv + 2 * p.m_pow2;
      int32_t* ci = reinterpret_cast<int32_t*>(cv + 128);
      const uint32_t kth = s_kth;
      for (int i = threadIdx.x; i < p.m_pow2; i += kFinThreads) {
        if (ord_of(i) <= kth) {
          const int slot = atomicAdd(&s_fill, 1);
          cv[slot] = sv[i];
          ci[slot] = si[i];
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < 128; i += kFinThreads) {
        const bool have = i < want;
        const float tv = have ? cv[i] : INFINITY;
        const int32_t ti = have ? ci[i] : INT_MAX;
        sv[i] = tv;
        si[i] = ti;
      }
      int n_sel = 2;
      while (n_sel < want) n_sel <<= 1;
      bitonic_sort_smem<float>(sv, si, n_sel);   // includes the barriers that publish sv / si
    }
  }
  if (!selected) bitonic_sort_smem<float>(sv, si, p.m_pow2);

  const int nfinite = s_nfinite;
  const int R = nfinite < p.cap ? nfinite : p.cap;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Only candidates that can still belong to the exact top-k are re-scored: with a_k the k-th
  // smallest APPROXIMATE value, k candidates have exact e <= a_k + m, so one whose approximate
  // value exceeds a_k + 2m (exact e > a_k + m) is out.  The gather of candidate rows is what this
  // kernel's time goes to (k=100, 2048-d fp32: 1 MB per query), so the cut is worth a third of it.
  int RS = R;
  if (p.k <= R) {
    const double lim = (double)sv[p.k - 1] + 2.0 * query_margin(p, q);
    int lo = p.k, hi = R;  // sv[0..R) ascending: first position whose value exceeds lim
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((double)sv[mid] <= lim) lo = mid + 1; else hi = mid;
    }
    RS = lo;
  }
  const T* qrow = reinterpret_cast<const T*>(p.q) + (size_t)q * p.dim;
  for (int c = threadIdx.x; c < 128; c += kFinThreads) {
    ex[c] = INFINITY;
    exi[c] = INT_MAX;
  }
  __syncthreads();
  for (int c = warp; c < RS; c += kFinThreads / 32) {
    const int32_t gi = si[c];
    const double d = warp_exact_distance<T, kVec>(qrow, reinterpret_cast<const T*>(p.g) + (size_t)gi * p.dim,
                                                  p.dim, p.metric, lane);
    if (lane == 0) {
      ex[c] = (double)(float)d;  // canonical order = (fp32 distance, index): shard-count invariant
      exi[c] = gi;
    }
  }
  // only the re-scored entries need sorting (the rest of ex[] holds +inf): k = 10 sorts 16 entries, not 128
  int n_sort = 2;
  while (n_sort < RS) n_sort <<= 1;
  bitonic_sort_smem<double>(ex, exi, n_sort);

  for (int i = threadIdx.x; i < p.k; i += kFinThreads) {
    const bool have = i < RS;
    p.out_dist[(size_t)q * p.k + i] = have ? (float)ex[i] : INFINITY;
    p.out_index[(size_t)q * p.k + i] = have ? (long long)exi[i] + p.index_offset : -1LL;
  }
  if (threadIdx.x == 0 && p.flags != nullptr) {
    int flag = 0;
    if (nfinite >= p.cap && p.k <= R) {
      // Every gallery row that was NOT re-scored has approx e >= tau; it can only belong
      // to the exact top-k if its exact e is below the k-th exact e, i.e. if
      // tau - margin < e_k.  Otherwise the selection is proven exact.
      const double tau = (double)sv[R - 1];
      const float qsq = p.qsq[q];
      const double ek = e_of_distance(ex[p.k - 1], p.metric, qsq);
      const double m = query_margin(p, q);
      if (!(ek + m < tau)) flag = 1;
    }
    p.flags[q] = flag;
    if (flag && p.uncertified != nullptr) atomicAdd(p.uncertified, 1);
  }
}

// ---- warp-per-query form for small candidate sets (at most 32 entries per query: 16-entry lists, one
// partition — the bf16 headline and its shards).  Same arithmetic, same (distance, index) order and the same
// certificate as finalize_topk_kernel, but the sorts are warp bitonic networks on registers (no shared memory,
// no block barriers) and a block of 8 warps finishes 8 queries: 100k queries take ~0.1 ms instead of ~1 ms,
// which is the part of a sharded step that does not shrink with the number of GPUs.
constexpr int kFinSmallWarps = 8;

template <typename Key>
__device__ __forceinline__ void warp_bitonic_sort32(Key& key, int32_t& idx, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const Key ok = __shfl_xor_sync(kFullMask, key, stride);
      const int32_t oi = __shfl_xor_sync(kFullMask, idx, stride);
      const bool lower = (lane & stride) == 0;          // this lane keeps the smaller of the pair when ascending
      const bool asc = (lane & size) == 0;
      const bool mine_first = key < ok || (key == ok && idx <= oi);
      const bool keep_mine = (lower == asc) ? mine_first : !mine_first;
      if (!keep_mine) { key = ok; idx = oi; }
    }
  }
}

template <typename T, bool kVec>
__global__ void __launch_bounds__(kFinSmallWarps * 32) finalize_topk_small_kernel(const FinParams p) {
  if (p.gate != nullptr && *p.gate == 0) return;
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * kFinSmallWarps + (threadIdx.x >> 5);
  if (q >= p.num_q) return;
  const int q_tile = q / kTileQ, row = q % kTileQ;
  const int m_tot = p.num_splits * p.lists_per_row * p.cap;   // <= 32
  float v = INFINITY;
  int32_t ix = INT_MAX;
  if (lane < m_tot) {
    const int l = lane / p.cap, pp = lane % p.cap;
    const int split = l / p.lists_per_row, h = l % p.lists_per_row;
    const size_t slot = ((size_t)split * p.q_tile_stride + q_tile) * p.lists_per_row + h;
    const size_t addr = (slot * p.cap + pp) * kTileQ + row;
    const float cv = p.cand_val[addr];
    if (cv < INFINITY) { v = cv; ix = p.cand_idx[addr]; }   // slots never filled keep +inf (their index is unspecified)
  }
  const int nfinite = __popc(__ballot_sync(kFullMask, v < INFINITY));
  warp_bitonic_sort32<float>(v, ix, lane);
  const int R = nfinite < p.cap ? nfinite : p.cap;
  const double m = query_margin(p, q);
  int RS = R;
  if (p.k <= R) {  // candidates whose approximate value exceeds a_k + 2m cannot reach the exact top-k
    const double lim = (double)__shfl_sync(kFullMask, v, p.k - 1) + 2.0 * m;
    RS = p.k + __popc(__ballot_sync(kFullMask, lane >= p.k && lane < R && (double)v <= lim));
  }
  const T* qrow = reinterpret_cast<const T*>(p.q) + (size_t)q * p.dim;
  float ex = INFINITY;
  int32_t exi = INT_MAX;
  for (int c = 0; c < RS; ++c) {
    const int32_t gi = __shfl_sync(kFullMask, ix, c);
    const double d = warp_exact_distance<T, kVec>(qrow, reinterpret_cast<const T*>(p.g) + (size_t)gi * p.dim, p.dim, p.metric, lane);
    if (lane == c) { ex = (float)d; exi = gi; }   // canonical order = (fp32 distance, index): shard-count invariant
  }
  warp_bitonic_sort32<float>(ex, exi, lane);
  if (lane < p.k) {
    const bool have = lane < RS;
    p.out_dist[(size_t)q * p.k + lane] = have ? ex : INFINITY;
    p.out_index[(size_t)q * p.k + lane] = have ? (long long)exi + p.index_offset : -1LL;
  }
  for (int i = 32 + lane; i < p.k; i += 32) {  // k above the 32 candidates there can be: padding
    p.out_dist[(size_t)q * p.k + i] = INFINITY;
    p.out_index[(size_t)q * p.k + i] = -1LL;
  }
  if (p.flags != nullptr) {
    int flag = 0;
    const float tau_f = __shfl_sync(kFullMask, v, R > 0 ? R - 1 : 0);
    const float ek_f = __shfl_sync(kFullMask, ex, p.k <= 32 ? p.k - 1 : 31);
    if (nfinite >= p.cap && p.k <= R) {
      const double ek = e_of_distance((double)ek_f, p.metric, p.qsq[q]);
      if (!(ek + m < (double)tau_f)) flag = 1;
    }
    if (lane == 0) {
      p.flags[q] = flag;
      if (flag && p.uncertified != nullptr) atomicAdd(p.uncertified, 1);
    }
  }
}

// Exact brute-force top-k for flagged queries: one 256-thread block per query, each warp
// keeps a sorted best-k list in shared memory, the 8 lists are merged by a bitonic sort.
constexpr int kFbThreads = 256;
constexpr int kFbWarps = kFbThreads / 32;

template <typename T, bool kVec>
__global__ void __launch_bounds__(kFbThreads) topk_fallback_kernel(const FinParams p) {
  __shared__ double wd[kFbWarps][128];
  __shared__ int32_t wi[kFbWarps][128];
  __shared__ double md[1024];
  __shared__ int32_t mi[1024];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = p.k;
  // A few blocks per SM walk the queries in runs of 32: every warp reads the run's flags with one coalesced load (the
  // same 32 words in all warps, so the ballot is block-uniform) and the block brute-forces the flagged ones.  With
  // nothing flagged — the normal case — 100k queries cost three load round trips per block (one dependent load per
  // query and block took 80 us).
  for (int base = blockIdx.x * 32; base < p.num_q; base += gridDim.x * 32) {
  unsigned todo = __ballot_sync(kFullMask, base + lane < p.num_q && (p.flags[base + lane] & 1) != 0);
  while (todo != 0u) {
  const int q = base + __ffs(todo) - 1;
  todo &= todo - 1;
  __syncthreads();  // shared lists of the previous flagged query fully consumed
  for (int i = lane; i < k; i += 32) {
    wd[warp][i] = INFINITY;
    wi[warp][i] = INT_MAX;
  }
  __syncwarp();
  const T* qrow = reinterpret_cast<const T*>(p.q) + (size_t)q * p.dim;
  for (int j = warp; j < p.num_g; j += kFbWarps) {
    const double d = (double)(float)warp_exact_distance<T, kVec>(
        qrow, reinterpret_cast<const T*>(p.g) + (size_t)j * p.dim, p.dim, p.metric, lane);
    // all lanes hold d; the list is sorted ascending, worst at k-1
    if (ranks_before(d, j, wd[warp][k - 1], wi[warp][k - 1])) {
      if (lane == 0) {
        int pos = k - 1;
        while (pos > 0 && ranks_before(d, j, wd[warp][pos - 1], wi[warp][pos - 1])) {
          wd[warp][pos] = wd[warp][pos - 1];
          wi[warp][pos] = wi[warp][pos - 1];
          --pos;
        }
        wd[warp][pos] = d;
        wi[warp][pos] = j;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += kFbThreads) {
    const int w = i / 128, pos = i % 128;
    const bool have = w < kFbWarps && pos < k;
    md[i] = have ? wd[w][pos] : INFINITY;
    mi[i] = have ? wi[w][pos] : INT_MAX;
  }
  bitonic_sort_smem<double>(md, mi, 1024);
  for (int i = threadIdx.x; i < k; i += kFbThreads) {
    const bool have = mi[i] != INT_MAX;
    p.out_dist[(size_t)q * k + i] = have ? (float)md[i] : INFINITY;
    p.out_index[(size_t)q * k + i] = have ? (long long)mi[i] + p.index_offset : -1LL;
  }
  }
  }
}

// ------------------------------------------------------------------- rank ----
struct RankParams {
  const void* q;
  const void* g;
  int num_q, num_g, dim, metric;
  const long long* pos_index;
  const double* pos_dist_in;
  const long long* pos_tie;  // index of the positive in the space (local row + tie_offset), or NULL
  long long tie_offset;
  const float* qsq;
  const float* gsq_max;
  float kappa;
  const float* q_res;
  const float* g_res;
  double* pos_dist;
  float* rank_lo;
  float* rank_hi;
  int32_t* cnt_less;
  uint32_t* pool_count;
  uint32_t pool_cap;
  int32_t* pool_q;
  int32_t* pool_idx;
  int32_t* dropped;
  long long* out_rank;
  long long missing_rank;
  const int32_t* gate;
};

constexpr int kRankWarps = 8;

// Canonical order of the ranked list: (fp32-rounded exact distance, gallery index).  Row j of
// this gallery (shard) precedes the positive of query q iff its distance is smaller, or equal
// with a smaller global index (pos_tie == NULL: equal distances never precede).
__device__ __forceinline__ bool ranks_before_positive(double d_exact, int j_local, const RankParams& p, int q) {
  const float d = (float)d_exact;
  const float dp = (float)p.pos_dist[q];
  if (d < dp) return true;
  if (d > dp || p.pos_tie == nullptr) return false;
  const long long pt = p.pos_tie[q];
  return pt >= 0 && (long long)j_local + p.tie_offset < pt;
}

// d_pos (exact) and the e-space band [lo, hi) in which K1's approximate comparison against
// d_pos cannot be trusted.
template <typename T, bool kVec>
__global__ void __launch_bounds__(kRankWarps * 32) rank_band_kernel(const RankParams p) {
  if (p.gate != nullptr && *p.gate == 0) return;
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * kRankWarps + (threadIdx.x >> 5);
  if (q >= p.num_q) return;
  double dpos;
  if (p.pos_dist_in != nullptr) {
    dpos = p.pos_dist_in[q];
  } else {
    const long long pi = p.pos_index[q];
    if (pi < 0 || pi >= p.num_g) {
      dpos = nan("");
    } else {
      dpos = warp_exact_distance<T, kVec>(reinterpret_cast<const T*>(p.q) + (size_t)q * p.dim,
                                          reinterpret_cast<const T*>(p.g) + (size_t)pi * p.dim, p.dim,
                                          p.metric, lane);
    }
  }
  if (lane == 0) {
    p.pos_dist[q] = dpos;
    if (dpos != dpos) {
      p.rank_lo[q] = -INFINITY;
      p.rank_hi[q] = -INFINITY;
    } else {
      const float qsq = p.qsq[q];
      const double c = e_of_distance(dpos, p.metric, qsq);
      const double m = query_margin(p, q);
      p.rank_lo[q] = __double2float_rd(c - m);
      p.rank_hi[q] = __double2float_ru(c + m);
    }
  }
}

// One warp per pooled (query, gallery row) pair: exact distance, exact comparison with d_pos.
template <typename T, bool kVec>
__global__ void __launch_bounds__(kRankWarps * 32) rank_resolve_kernel(const RankParams p) {
  if (p.gate != nullptr && *p.gate == 0) return;
  const int lane = threadIdx.x & 31;
  const uint32_t claimed = *p.pool_count;
  const uint32_t n = claimed < p.pool_cap ? claimed : p.pool_cap;
  const uint32_t stride = gridDim.x * kRankWarps;
  for (uint32_t e = blockIdx.x * kRankWarps + (threadIdx.x >> 5); e < n; e += stride) {
    const int q = p.pool_q[e];
    const int gi = p.pool_idx[e];
    const double d = warp_exact_distance<T, kVec>(reinterpret_cast<const T*>(p.q) + (size_t)q * p.dim,
                                                  reinterpret_cast<const T*>(p.g) + (size_t)gi * p.dim,
                                                  p.dim, p.metric, lane);
    if (lane == 0 && ranks_before_positive(d, gi, p, q)) atomicAdd(p.cnt_less + q, 1);
  }
}

// Rank output, once, after the last scoring pass: the fused counters (+ the resolved band) give the rank;
// queries without a positive get `missing_rank`; queries whose uncertain band overflowed the pool are
// counted exactly by brute force.  A few blocks per SM walk the queries.
template <typename T, bool kVec>
__global__ void __launch_bounds__(kFbThreads) rank_output_kernel(const RankParams p) {
  __shared__ int red[kFbWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // runs of 32 queries, read by every warp with coalesced loads (block-uniform ballot, see topk_fallback_kernel): warp 0
  // writes the ranks the counters settled, the block counts the overflowed ones exactly
  for (int base = blockIdx.x * 32; base < p.num_q; base += gridDim.x * 32) {
    const int qi = base + lane;
    bool brute = false;
    if (qi < p.num_q) {
      const double dp = p.pos_dist[qi];
      brute = dp == dp && p.dropped[qi] > 0;
      if (!brute && warp == 0) p.out_rank[qi] = (dp != dp) ? p.missing_rank : (long long)p.cnt_less[qi];
    }
    unsigned todo = __ballot_sync(kFullMask, brute);
    while (todo != 0u) {
    const int q = base + __ffs(todo) - 1;
    todo &= todo - 1;
    __syncthreads();
    const T* qrow = reinterpret_cast<const T*>(p.q) + (size_t)q * p.dim;
    int cnt = 0;
    for (int j = warp; j < p.num_g; j += kFbWarps) {
      const double d = warp_exact_distance<T, kVec>(qrow, reinterpret_cast<const T*>(p.g) + (size_t)j * p.dim,
                                                    p.dim, p.metric, lane);
      cnt += ranks_before_positive(d, j, p, q) ? 1 : 0;
    }
    if (lane == 0) red[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long r = 0;
      for (int w = 0; w < kFbWarps; ++w) r += red[w];
      p.out_rank[q] = r;
    }
    }
  }
}

// Start of a (device-gated) scoring pass: counters, scheduler state and shared thresholds back to their
// initial values in ONE launch (memsets cannot be gated on a device flag).
__global__ void __launch_bounds__(256) pass_reset_kernel(int32_t* __restrict__ cnt_less, int32_t* __restrict__ dropped, long long num_q,
                                                         uint32_t* __restrict__ pool_count, uint32_t* __restrict__ sched,
                                                         long long sched_words, int32_t* __restrict__ shared_thr, long long num_thr,
                                                         const int32_t* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  if (cnt_less != nullptr)
    for (long long i = i0; i < num_q; i += stride) { cnt_less[i] = 0; dropped[i] = 0; }
  for (long long i = i0; i < sched_words; i += stride) sched[i] = 0u;
  for (long long i = i0; i < num_thr; i += stride) shared_thr[i] = 0x7f800000;
  if (i0 == 0 && pool_count != nullptr) *pool_count = 0u;
}

template <typename T, bool kVec>
__global__ void __launch_bounds__(kRankWarps * 32) positive_distance_kernel(
    const T* __restrict__ q, int num_q, const T* __restrict__ g, int num_g, int dim, int metric,
    const long long* __restrict__ pos_index, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * kRankWarps + (threadIdx.x >> 5);
  if (qi >= num_q) return;
  const long long pi = pos_index[qi];
  double d = nan("");
  if (pi >= 0 && pi < num_g)
    d = warp_exact_distance<T, kVec>(q + (size_t)qi * dim, g + (size_t)pi * dim, dim, metric, lane);
  if (lane == 0) out[qi] = d;
}

// ------------------------------------------------------------------ K4 merge ----
// k best of `num_lists` ascending lists of length k per query; order = (distance, index), padding
// entries (index < 0, +inf) last.  Deterministic.
//
// Bandwidth form (the one that runs): a block takes `qb` consecutive queries; for every list
// their entries are ONE contiguous run of qb·k values in the gathered [list][query][k] buffer,
// read with coalesced loads into shared memory.  The lists are then merged pairwise in
// ⌈log2 num_lists⌉ rounds; in a round one thread produces one output element by a merge-path
// binary search (≤ log2 k steps), so the work per query is ~num_lists·k·log2 k instead of the
// num_lists²·k·log2 k of rank-by-counting.  The merged rows leave as one contiguous run.
// HBM traffic = the algorithmic num_lists·Q·k·12 bytes in + Q·k·12 out, once.
constexpr int kMergeThreads = 128;
// (da, ia) precedes (db, ib).  Padding entries carry +inf, so the distance alone decides unless
// the two are equal (then valid-before-padding, smaller index first).
__device__ __forceinline__ bool merge_before(float da, float db, const long long* pia, const long long* pib) {
  if (da != db) return da < db;
  const long long ia = *pia, ib = *pib;
  return ia >= 0 && (ib < 0 || ia < ib);
}
// x / d for block-local counters (x < 2^16, d <= a few thousand): exact via one float multiply.
__device__ __forceinline__ int small_div(int x, float inv_d) { return __float2int_rz(((float)x + 0.5f) * inv_d); }

__global__ void __launch_bounds__(kMergeThreads) topk_merge_kernel(
    const float* __restrict__ dist, const long long* __restrict__ index, int num_lists, size_t stride_d, size_t stride_i,
    int num_q, int k, int qb, float* __restrict__ out_dist, long long* __restrict__ out_index) {
  extern __shared__ __align__(16) uint8_t merge_smem[];
  const int half_lists = (num_lists + 1) / 2;
  // ping [qb][num_lists][k], pong [qb][half_lists][k]; indices first (8-byte aligned), then distances
  long long* i_ping = reinterpret_cast<long long*>(merge_smem);
  long long* i_pong = i_ping + (size_t)qb * num_lists * k;
  float* d_ping = reinterpret_cast<float*>(i_pong + (size_t)qb * half_lists * k);
  float* d_pong = d_ping + (size_t)qb * num_lists * k;
  const int q0 = blockIdx.x * qb;
  const int nq = min(qb, num_q - q0);
  const int run = nq * k;  // contiguous entries per list for this block
  const float inv_k = 1.0f / (float)k;
  for (int x = threadIdx.x; x < run; x += kMergeThreads) {
    const int qi = small_div(x, inv_k), i = x - qi * k;
    const int dst = qi * num_lists * k + i;
    const size_t src = (size_t)q0 * k + x;
#pragma unroll 4
    for (int l = 0; l < num_lists; ++l) {
      d_ping[dst + l * k] = __ldg(dist + (size_t)l * stride_d + src);
      i_ping[dst + l * k] = __ldg(index + (size_t)l * stride_i + src);
    }
  }
  __syncthreads();
  const float* sd = d_ping;
  const long long* si = i_ping;
  float* td = d_pong;
  long long* ti = i_pong;
  int n = num_lists, s_stride = num_lists, t_stride = half_lists;
  while (n > 1) {
    const int pairs = n >> 1, n_next = (n + 1) >> 1;
    const int per_q = n_next * k;
    const int outs = nq * per_q;
    const float inv_per_q = 1.0f / (float)per_q;
    for (int x = threadIdx.x; x < outs; x += kMergeThreads) {
      const int qi = small_div(x, inv_per_q), r = x - qi * per_q;
      const int pr = small_div(r, inv_k), o = r - pr * k;
      const int dst = (qi * t_stride + pr) * k + o;
      if (pr >= pairs) {  // odd list out: carried to the next round unchanged
        td[dst] = sd[(qi * s_stride + n - 1) * k + o];
        ti[dst] = si[(qi * s_stride + n - 1) * k + o];
        continue;
      }
      const int a = (qi * s_stride + 2 * pr) * k, b = a + k;
      int lo = 0, hi = o;  // number of A elements among the first o merged elements
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        // A[mid] belongs to the first o elements unless B[o-1-mid] strictly precedes it
        if (!merge_before(sd[b + o - 1 - mid], sd[a + mid], si + b + o - 1 - mid, si + a + mid)) lo = mid + 1;
        else hi = mid;
      }
      const int ia = a + lo, ib = b + o - lo;
      const int pick = merge_before(sd[ib], sd[ia], si + ib, si + ia) ? ib : ia;
      td[dst] = sd[pick];
      ti[dst] = si[pick];
    }
    __syncthreads();
    // the output of this round becomes the input of the next
    const float* nd = td; const long long* ni = ti;
    td = const_cast<float*>(sd); ti = const_cast<long long*>(si);
    sd = nd; si = ni;
    const int tmp = s_stride; s_stride = t_stride; t_stride = tmp;
    n = n_next;
  }
  const size_t obase = (size_t)q0 * k;
  for (int x = threadIdx.x; x < run; x += kMergeThreads) {
    const int qi = small_div(x, inv_k), i = x - qi * k;
    const long long ix = si[(qi * s_stride) * k + i];
    out_dist[obase + x] = ix >= 0 ? sd[(qi * s_stride) * k + i] : INFINITY;
    out_index[obase + x] = ix >= 0 ? ix : -1;
  }
}

// Tournament form (at most 32 lists — the shard counts of one box): the same coalesced gather of the block's runs
// into shared memory, then a GROUP of G = next_pow2(num_lists) lanes per query holds the heads of that query's lists
// and pops the minimum k times (3·log2(G) shuffles per pop); a warp merges 32/G queries side by side and the rows leave
// as one contiguous run.  ~60 instructions per query at 8 lists × k = 10 instead of ~600 for the merge-path rounds
// (which Nsight showed instruction-bound: issue slots 69 % busy, 1.7 TB/s), so the kernel is bound by its loads.
template <int G>
__global__ void __launch_bounds__(kMergeThreads, 8) topk_merge_tournament_kernel(
    const float* __restrict__ dist, const long long* __restrict__ index, int num_lists, size_t stride_d, size_t stride_i,
    int num_q, int k, int qb, float* __restrict__ out_dist, long long* __restrict__ out_index) {
  extern __shared__ __align__(16) uint8_t merge_smem[];
  // lists [qb][num_lists][k] and merged rows [qb][k]; indices first (8-byte aligned), then distances
  long long* i_in = reinterpret_cast<long long*>(merge_smem);
  long long* i_out = i_in + (size_t)qb * num_lists * k;
  float* d_in = reinterpret_cast<float*>(i_out + (size_t)qb * k);
  float* d_out = d_in + (size_t)qb * num_lists * k;
  const int q0 = blockIdx.x * qb;
  const int nq = min(qb, num_q - q0);
  const int run = nq * k;  // contiguous entries per list for this block
  const float inv_k = 1.0f / (float)k;
  for (int x = threadIdx.x; x < run; x += kMergeThreads) {
    const int qi = small_div(x, inv_k), i = x - qi * k;
    const int dst = qi * num_lists * k + i;
    const size_t src = (size_t)q0 * k + x;
#pragma unroll 4
    for (int l = 0; l < num_lists; ++l) {
      d_in[dst + l * k] = __ldg(dist + (size_t)l * stride_d + src);
      i_in[dst + l * k] = __ldg(index + (size_t)l * stride_i + src);
    }
  }
  __syncthreads();
  constexpr int kGroups = 32 / G;                   // queries per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / G, p = lane % G;           // query slot of the warp, list of this lane
  for (int qi0 = warp * kGroups; qi0 < nq; qi0 += (kMergeThreads / 32) * kGroups) {
    const int qi = qi0 + sub;
    const bool live = qi < nq && p < num_lists;
    const int base = live ? (qi * num_lists + p) * k : 0;
    int h = 0;                                      // head of this lane's list
    float hd = live ? d_in[base] : INFINITY;
    long long hi = live ? i_in[base] : -1LL;
    for (int o = 0; o < k; ++o) {
      // minimum of the group's heads in the (distance, valid-before-padding, index) order
      float bd = hd;
      long long bi = hi;
      int bl = p;
#pragma unroll
      for (int s = G >> 1; s > 0; s >>= 1) {
        const float od = __shfl_xor_sync(kFullMask, bd, s);
        const long long oi = __shfl_xor_sync(kFullMask, bi, s);
        const int ol = __shfl_xor_sync(kFullMask, bl, s);
        const bool other_first = (od != bd) ? (od < bd) : ((oi >= 0 && (bi < 0 || oi < bi)) || (oi == bi && ol < bl));
        if (other_first) { bd = od; bi = oi; bl = ol; }
      }
      if (qi < nq && p == 0) {
        d_out[qi * k + o] = bi >= 0 ? bd : INFINITY;
        i_out[qi * k + o] = bi >= 0 ? bi : -1LL;
      }
      if (live && bl == p && bi >= 0) {             // this lane's head won: advance
        ++h;
        hd = h < k ? d_in[base + h] : INFINITY;
        hi = h < k ? i_in[base + h] : -1LL;
      }
    }
  }
  __syncthreads();
  const size_t obase = (size_t)q0 * k;
  for (int x = threadIdx.x; x < run; x += kMergeThreads) {
    out_dist[obase + x] = d_out[x];
    out_index[obase + x] = i_out[x];
  }
}

// Generic form for list sets that do not fit in shared memory (hundreds of lists × large k):
// one warp per query, binary searches straight in global memory.
constexpr int kMergeWarps = 4;
__global__ void __launch_bounds__(kMergeWarps * 32) topk_merge_generic_kernel(
    const float* __restrict__ dist, const long long* __restrict__ index, int num_lists, size_t stride_d, size_t stride_i,
    int num_q, int k, float* __restrict__ out_dist, long long* __restrict__ out_index) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * kMergeWarps + (threadIdx.x >> 5);
  if (q >= num_q) return;
  const int total = num_lists * k;
  for (int x = lane; x < total; x += 32) {
    const int l = x / k, i = x % k;
    const float dx = dist[(size_t)l * stride_d + (size_t)q * k + i];
    const long long ix = index[(size_t)l * stride_i + (size_t)q * k + i];
    if (ix < 0) continue;
    int pos = i;
    for (int l2 = 0; l2 < num_lists; ++l2) {
      if (l2 == l) continue;
      const size_t b2d = (size_t)l2 * stride_d + (size_t)q * k, b2i = (size_t)l2 * stride_i + (size_t)q * k;
      int lo = 0, hi = k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float dm = dist[b2d + mid];
        const long long im = index[b2i + mid];
        const bool before = im >= 0 && (dm < dx || (dm == dx && im < ix));
        if (before) lo = mid + 1; else hi = mid;
      }
      pos += lo;
    }
    if (pos < k) {
      out_dist[(size_t)q * k + pos] = dx;
      out_index[(size_t)q * k + pos] = ix;
    }
  }
}
__global__ void fill_topk_kernel(float* out_dist, long long* out_index, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    out_dist[i] = INFINITY;
    out_index[i] = -1;
  }
}

// gate[0] = 1 iff more than max_bad queries are unresolved after the TF32 pass (top-k certificate
// failed, or rank pool entries dropped).  One block; when escalating, the diagnostic counter
// restarts for the second pass.
__global__ void __launch_bounds__(1024) escalate_decide_kernel(const int32_t* __restrict__ flags,
                                                               const int32_t* __restrict__ dropped, long long num_q,
                                                               long long max_bad, int32_t* __restrict__ gate,
                                                               int32_t* __restrict__ uncertified) {
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  int bad = 0;
  for (long long i = threadIdx.x; i < num_q; i += 1024)
    bad += ((flags != nullptr && (flags[i] & 1)) || (dropped != nullptr && dropped[i] > 0)) ? 1 : 0;
  if (bad) atomicAdd(&s_bad, bad);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int esc = (long long)s_bad > max_bad ? 1 : 0;
    gate[0] = esc;
    if (esc && uncertified != nullptr) uncertified[0] = 0;
  }
}

// ---- tiers behind the bf16 selection of fp32 embeddings (api.cu) ----
// After the kind::f16 pass: how many queries are unresolved (top-k certificate failed / rank pool overflowed)?
//   none                                  -> nothing to do
//   a few certificate failures only       -> gate_sub: those queries (listed in `fq`, ascending) are re-selected on
//                                            kind::tf32 tiles as a small batch of their own
//   more, or rank pools overflowed        -> gate_full: one kind::tf32 pass over all queries
// One block; the list is built by an ordered scan so the subset's row order is deterministic.
__global__ void __launch_bounds__(1024) tier_decide_kernel(const int32_t* __restrict__ flags, const int32_t* __restrict__ dropped,
                                                           int num_q, int max_sub, int32_t* __restrict__ fq, int32_t* __restrict__ fq_count,
                                                           int32_t* __restrict__ gate_sub, int32_t* __restrict__ gate_full,
                                                           int32_t* __restrict__ uncertified) {
  __shared__ int s_flag, s_drop, s_base;
  __shared__ int s_warp[32];
  if (threadIdx.x == 0) { s_flag = 0; s_drop = 0; s_base = 0; }
  __syncthreads();
  int nf = 0, nd = 0;
  for (int i = threadIdx.x; i < num_q; i += 1024) {
    nf += (flags[i] & 1) ? 1 : 0;
    nd += (dropped != nullptr && dropped[i] > 0) ? 1 : 0;
  }
  if (nf) atomicAdd(&s_flag, nf);
  if (nd) atomicAdd(&s_drop, nd);
  __syncthreads();
  const int total_flag = s_flag, total_drop = s_drop;
  const bool sub = total_flag > 0 && total_flag <= max_sub && total_drop == 0;
  const bool full = (total_flag > 0 || total_drop > 0) && !sub;
  if (threadIdx.x == 0) {
    gate_sub[0] = sub ? 1 : 0;
    gate_full[0] = full ? 1 : 0;
    fq_count[0] = sub ? total_flag : 0;
    if (full && uncertified != nullptr) uncertified[0] = 0;  // the full pass counts again
  }
  if (!sub) return;
  // ordered compaction: chunks of 1024 queries, ballot + warp prefix inside a chunk
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < num_q; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const bool f = i < num_q && (flags[i] & 1);
    const unsigned m = __ballot_sync(kFullMask, f);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (f) fq[before + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
}

// q_sub[i, :] = q[fq[i], :] for i < count (zero rows behind), qsq_sub likewise.  One warp per row.
__global__ void __launch_bounds__(256) gather_sub_kernel(const float* __restrict__ q, int dim, const int32_t* __restrict__ fq,
                                                         const int32_t* __restrict__ fq_count, int sub_q, float* __restrict__ q_sub,
                                                         const float* __restrict__ qsq, float* __restrict__ qsq_sub,
                                                         const int32_t* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= sub_q) return;
  const int n = *fq_count;
  const bool have = r < n;
  const float* src = have ? q + (size_t)fq[r] * dim : nullptr;
  for (int c = lane; c < dim; c += 32) q_sub[(size_t)r * dim + c] = have ? src[c] : 0.f;
  if (lane == 0) qsq_sub[r] = have ? qsq[fq[r]] : 0.f;
}

// results of the subset pass back into the caller's rows; certificate flags and the uncertified counter follow
__global__ void __launch_bounds__(256) scatter_sub_kernel(const int32_t* __restrict__ fq, const int32_t* __restrict__ fq_count, int k,
                                                          const float* __restrict__ sub_dist, const long long* __restrict__ sub_index,
                                                          const int32_t* __restrict__ sub_flags, float* __restrict__ out_dist,
                                                          long long* __restrict__ out_index, int32_t* __restrict__ flags,
                                                          int32_t* __restrict__ uncertified, const int32_t* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int n = *fq_count;
  for (int r = blockIdx.x; r < n; r += gridDim.x) {
    const int q = fq[r];
    for (int i = threadIdx.x; i < k; i += 256) {
      out_dist[(size_t)q * k + i] = sub_dist[(size_t)r * k + i];
      out_index[(size_t)q * k + i] = sub_index[(size_t)r * k + i];
    }
    if (threadIdx.x == 0 && (sub_flags[r] & 1) == 0) {
      flags[q] = 0;
      if (uncertified != nullptr) atomicSub(uncertified, 1);
    }
  }
}

// 3xTF32 operand split (see kernels.h).  float4 in, three float4 out per vector.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, long long rows, int dim,
                                                         int gallery_layout, float* __restrict__ out,
                                                         const int32_t* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int v4 = dim / 4;
  const long long n4 = rows * v4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / v4;
    const int c = (int)(i - r * v4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
    float4* o = reinterpret_cast<float4*>(out) + r * (3LL * v4) + c;
    o[0] = h;
    o[v4] = gallery_layout ? l : h;
    o[2 * v4] = gallery_layout ? h : l;
  }
}

// ---- mean-centred 3xTF32 operands (euclidean escalation pass) ----
// Distances are translation-invariant, tensor-core rounding is not: its error scales with
// ‖q‖·‖g‖, so embeddings that share a large common component (post-ReLU features, collapsed or
// untrained encoders) lose the small differences that decide the ranking.  The escalation pass
// therefore subtracts the gallery's column mean µ from both operands before the hi/lo split:
// ‖q−g‖ = ‖(q−µ)−(g−µ)‖ exactly, fl32(x−µ) carries 2^-24 relative rounding per element, and
// the error band then scales with the SPREAD of the embeddings instead of their norms.
constexpr int kColBlocks = 592;  // 4 per SM: partial column sums [kColBlocks][dim]

__global__ void __launch_bounds__(256) col_partial_kernel(const float* __restrict__ x, long long rows, int dim,
                                                          float* __restrict__ partial, const int32_t* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int v4 = dim / 4;
  for (int c = threadIdx.x; c < v4; c += 256) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * dim) + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * dim)[c] = acc;
  }
}

// µ[c] = Σ_b partial[b][c] / rows (fixed order: deterministic); also clears the running maximum
// the centred gallery pass accumulates into.
__global__ void __launch_bounds__(256) col_mean_kernel(const float* __restrict__ partial, int num_blocks, long long rows,
                                                       int dim, float* __restrict__ mu, float* __restrict__ max_reset,
                                                       const int32_t* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c == 0 && max_reset != nullptr) *max_reset = 0.f;
  if (c >= dim) return;
  double acc = 0.0;
  for (int b = 0; b < num_blocks; ++b) acc += (double)partial[(size_t)b * dim + c];
  mu[c] = (float)(acc / (double)rows);
}

// x [rows, dim] fp32 -> out [rows, 3*dim]: c = fl32(x − µ) split like split_tf32_kernel; also
// vec[r] = ‖c_r‖² (what K1's epilogue and the certificate use as ‖q‖² / ‖g‖² in this pass),
// vec[rows..rows_padded) = pad_value, *max_out = max ‖c_r‖².  One warp per row.
__global__ void __launch_bounds__(256) center_split_tf32_kernel(const float* __restrict__ x, long long rows,
                                                                long long rows_padded, int dim,
                                                                const float* __restrict__ mu, int gallery_layout,
                                                                float* __restrict__ out, float* __restrict__ vec,
                                                                float pad_value, float* __restrict__ max_out,
                                                                const int32_t* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int lane = threadIdx.x & 31;
  const int v4 = dim / 4;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  float local_max = 0.f;
  for (long long r = warp0; r < rows_padded; r += nwarps) {
    if (r >= rows) {
      if (lane == 0) vec[r] = pad_value;
      continue;
    }
    const float4* xr = reinterpret_cast<const float4*>(x + r * dim);
    float4* o = reinterpret_cast<float4*>(out + r * 3LL * dim);
    double acc = 0.0;
    for (int i = lane; i < v4; i += 32) {
      const float4 v = __ldg(xr + i);
      const float4 m = __ldg(reinterpret_cast<const float4*>(mu) + i);
      const float c[4] = {__fsub_rn(v.x, m.x), __fsub_rn(v.y, m.y), __fsub_rn(v.z, m.z), __fsub_rn(v.w, m.w)};
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = __uint_as_float(__float_as_uint(c[e]) & 0xFFFFE000u);
        l[e] = c[e] - h[e];
        acc += (double)c[e] * (double)c[e];
      }
      const float4 h4 = make_float4(h[0], h[1], h[2], h[3]), l4 = make_float4(l[0], l[1], l[2], l[3]);
      o[i] = h4;
      o[v4 + i] = gallery_layout ? l4 : h4;
      o[2 * v4 + i] = gallery_layout ? h4 : l4;
    }
    const float sq = (float)warp_sum(acc);
    local_max = fmaxf(local_max, sq);
    if (lane == 0) vec[r] = sq;
  }
  if (max_out != nullptr && lane == 0 && local_max > 0.f)
    atomicMax(reinterpret_cast<int*>(max_out), __float_as_int(local_max));
}

__global__ void fill_i32_kernel(int32_t* out, long long n, int32_t value) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = value;
}
__global__ void fill_i64_kernel(long long* out, long long n, long long value) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = value;
}

// ------------------------------------------------------------------ metrics ----
// Single block: MRR, cumulative top-k accuracy, mean / sample-std / min / max of (rank0+1).
__global__ void __launch_bounds__(256) retrieval_metrics_kernel(const long long* __restrict__ rank0,
                                                                long long num_q, int k,
                                                                double* __restrict__ out) {
  __shared__ double red[256];
  __shared__ double s_mean;
  auto block_sum = [&](double v) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  auto block_minmax = [&](double v, bool want_min) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s)
        red[threadIdx.x] = want_min ? fmin(red[threadIdx.x], red[threadIdx.x + s]) : fmax(red[threadIdx.x], red[threadIdx.x + s]);
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  double mrr = 0.0, sum = 0.0, mn = INFINITY, mx = -INFINITY;
  for (long long i = threadIdx.x; i < num_q; i += 256) {
    const double r1 = (double)(rank0[i] + 1);
    mrr += 1.0 / r1;
    sum += r1;
    mn = fmin(mn, r1);
    mx = fmax(mx, r1);
  }
  mrr = block_sum(mrr);
  sum = block_sum(sum);
  mn = block_minmax(mn, true);
  mx = block_minmax(mx, false);
  if (threadIdx.x == 0) s_mean = sum / (double)num_q;
  __syncthreads();
  const double mean = s_mean;
  double ss = 0.0;
  for (long long i = threadIdx.x; i < num_q; i += 256) {
    const double dlt = (double)(rank0[i] + 1) - mean;
    ss += dlt * dlt;
  }
  ss = block_sum(ss);
  for (int kk = 0; kk < k; ++kk) {
    double c = 0.0;
    for (long long i = threadIdx.x; i < num_q; i += 256) c += rank0[i] <= kk ? 1.0 : 0.0;
    c = block_sum(c);
    if (threadIdx.x == 0) out[1 + kk] = c / (double)num_q;
  }
  if (threadIdx.x == 0) {
    out[0] = mrr / (double)num_q;
    out[k + 1] = mean;
    out[k + 2] = num_q > 1 ? sqrt(ss / (double)(num_q - 1)) : nan("");
    out[k + 3] = mn;
    out[k + 4] = mx;
  }
}

int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

FinParams make_fin_params(const FinalizeArgs& a, const K1Plan* plan) {
  FinParams p{};
  p.q = a.q; p.g = a.g;
  p.num_q = (int)a.num_q; p.num_g = (int)a.num_g; p.dim = (int)a.dim;
  p.metric = a.metric; p.k = a.k;
  p.index_offset = a.index_offset;
  p.cand_val = a.cand_val; p.cand_idx = a.cand_idx;
  if (plan) {
    p.cap = plan->cap; p.lists_per_row = plan->lists_per_row;
    p.q_tile_stride = plan->q_tile_stride; p.num_splits = plan->num_splits;
    const int m = next_pow2(plan->num_splits * plan->lists_per_row * plan->cap);
    p.m_pow2 = m < 2 ? 2 : m;
  }
  p.qsq = a.qsq; p.gsq_max = a.gsq_max; p.kappa = a.kappa;
  p.q_res = a.q_res; p.g_res = a.g_res;
  p.out_dist = a.out_dist; p.out_index = reinterpret_cast<long long*>(a.out_index);
  p.uncertified = a.uncertified; p.flags = a.flags;
  p.gate = a.gate;
  return p;
}

RankParams make_rank_params(const RankArgs& a) {
  RankParams p{};
  p.q = a.q; p.g = a.g;
  p.num_q = (int)a.num_q; p.num_g = (int)a.num_g; p.dim = (int)a.dim; p.metric = a.metric;
  p.pos_index = reinterpret_cast<const long long*>(a.pos_index);
  p.pos_dist_in = a.pos_dist_in;
  p.pos_tie = reinterpret_cast<const long long*>(a.pos_tie);
  p.tie_offset = a.tie_offset;
  p.qsq = a.qsq; p.gsq_max = a.gsq_max; p.kappa = a.kappa;
  p.q_res = a.q_res; p.g_res = a.g_res;
  p.pos_dist = a.pos_dist; p.rank_lo = a.rank_lo; p.rank_hi = a.rank_hi;
  p.cnt_less = a.cnt_less;
  p.pool_count = a.pool_count; p.pool_cap = a.pool_cap; p.pool_q = a.pool_q; p.pool_idx = a.pool_idx;
  p.dropped = a.dropped;
  p.out_rank = reinterpret_cast<long long*>(a.out_rank);
  p.missing_rank = a.missing_rank;
  p.gate = a.gate;
  return p;
}

// Dispatch on (element type, vectorisable) — KERNEL<T, kVec><<<...>>>(args)
#define SBIR_DISPATCH_T(dtype, vec, KERNEL, grid, block, smem, st, ...)                       \
  do {                                                                                        \
    if ((dtype) == SBIR_F32) {                                                                \
      if (vec) KERNEL<float, true><<<grid, block, smem, st>>>(__VA_ARGS__);                   \
      else KERNEL<float, false><<<grid, block, smem, st>>>(__VA_ARGS__);                      \
    } else {                                                                                  \
      if (vec) KERNEL<__nv_bfloat16, true><<<grid, block, smem, st>>>(__VA_ARGS__);           \
      else KERNEL<__nv_bfloat16, false><<<grid, block, smem, st>>>(__VA_ARGS__);              \
    }                                                                                         \
  } while (0)

}  // namespace

int launch_finalize_topk(const FinalizeArgs& a, const K1Plan& plan, cudaStream_t st) {
  if (a.num_q <= 0) return SBIR_OK;
  const FinParams p = make_fin_params(a, &plan);
  const size_t smem = (size_t)p.m_pow2 * 8 + (p.m_pow2 > 128 ? 128 * 8 : 0);   // lists + the select's compaction scratch
  if (smem > 40 * 1024) return SBIR_ERR_UNSUPPORTED;
  const bool vec = rows_vectorizable(a.q, a.dim, a.dtype) && rows_vectorizable(a.g, a.dim, a.dtype);
  if (plan.num_splits * plan.lists_per_row * plan.cap <= 32) {
    const unsigned grid = (unsigned)((a.num_q + kFinSmallWarps - 1) / kFinSmallWarps);
    SBIR_DISPATCH_T(a.dtype, vec, finalize_topk_small_kernel, grid, kFinSmallWarps * 32, 0, st, p);
    SBIR_CHECK_LAUNCH();
    return SBIR_OK;
  }
  SBIR_DISPATCH_T(a.dtype, vec, finalize_topk_kernel, (unsigned)a.num_q, kFinThreads, smem, st, p);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_topk_fallback(const FinalizeArgs& a, cudaStream_t st) {
  if (a.num_q <= 0) return SBIR_OK;
  const FinParams p = make_fin_params(a, nullptr);
  const bool vec = rows_vectorizable(a.q, a.dim, a.dtype) && rows_vectorizable(a.g, a.dim, a.dtype);
  const int64_t runs = (a.num_q + 31) / 32;
  const unsigned grid = (unsigned)(runs < 148 * 8 ? runs : 148 * 8);
  SBIR_DISPATCH_T(a.dtype, vec, topk_fallback_kernel, grid, kFbThreads, 0, st, p);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_rank_band(const RankArgs& a, cudaStream_t st) {
  if (a.num_q <= 0) return SBIR_OK;
  const RankParams p = make_rank_params(a);
  const bool vec = rows_vectorizable(a.q, a.dim, a.dtype) && rows_vectorizable(a.g, a.dim, a.dtype);
  const unsigned grid = (unsigned)((a.num_q + kRankWarps - 1) / kRankWarps);
  SBIR_DISPATCH_T(a.dtype, vec, rank_band_kernel, grid, kRankWarps * 32, 0, st, p);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_rank_resolve(const RankArgs& a, cudaStream_t st) {
  if (a.num_q <= 0) return SBIR_OK;
  const RankParams p = make_rank_params(a);
  const bool vec = rows_vectorizable(a.q, a.dim, a.dtype) && rows_vectorizable(a.g, a.dim, a.dtype);
  SBIR_DISPATCH_T(a.dtype, vec, rank_resolve_kernel, 148 * 8, kRankWarps * 32, 0, st, p);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_rank_output(const RankArgs& a, cudaStream_t st) {
  if (a.num_q <= 0) return SBIR_OK;
  const RankParams p = make_rank_params(a);
  const bool vec = rows_vectorizable(a.q, a.dim, a.dtype) && rows_vectorizable(a.g, a.dim, a.dtype);
  const int64_t runs = (a.num_q + 31) / 32;
  const unsigned grid = (unsigned)(runs < 148 * 8 ? runs : 148 * 8);
  SBIR_DISPATCH_T(a.dtype, vec, rank_output_kernel, grid, kFbThreads, 0, st, p);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_pass_reset(int32_t* cnt_less, int32_t* dropped, int64_t num_q, uint32_t* pool_count, void* sched, size_t sched_bytes,
                      int32_t* shared_thr, int64_t num_thr, const int32_t* gate, cudaStream_t st) {
  const long long n = std::max<long long>(std::max<long long>(num_q, (long long)(sched_bytes / 4)), num_thr);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  pass_reset_kernel<<<(unsigned)blocks, 256, 0, st>>>(cnt_less, dropped, (long long)num_q, pool_count, static_cast<uint32_t*>(sched),
                                                       (long long)(sched_bytes / 4), shared_thr, (long long)num_thr, gate);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_tier_decide(const int32_t* flags, const int32_t* dropped, int64_t num_q, int64_t max_sub, int32_t* fq, int32_t* fq_count,
                       int32_t* gate_sub, int32_t* gate_full, int32_t* uncertified, cudaStream_t st) {
  tier_decide_kernel<<<1, 1024, 0, st>>>(flags, dropped, (int)num_q, (int)max_sub, fq, fq_count, gate_sub, gate_full, uncertified);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_gather_sub(const float* q, int64_t dim, const int32_t* fq, const int32_t* fq_count, int64_t sub_q, float* q_sub,
                      const float* qsq, float* qsq_sub, const int32_t* gate, cudaStream_t st) {
  gather_sub_kernel<<<(unsigned)((sub_q + 7) / 8), 256, 0, st>>>(q, (int)dim, fq, fq_count, (int)sub_q, q_sub, qsq, qsq_sub, gate);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_scatter_sub(const int32_t* fq, const int32_t* fq_count, int64_t sub_q, int k, const float* sub_dist, const int64_t* sub_index,
                       const int32_t* sub_flags, float* out_dist, int64_t* out_index, int32_t* flags, int32_t* uncertified,
                       const int32_t* gate, cudaStream_t st) {
  scatter_sub_kernel<<<(unsigned)(sub_q < 1 ? 1 : sub_q), 256, 0, st>>>(fq, fq_count, k, sub_dist, reinterpret_cast<const long long*>(sub_index),
                                                                      sub_flags, out_dist, reinterpret_cast<long long*>(out_index), flags,
                                                                      uncertified, gate);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_escalate_decide(const int32_t* flags, const int32_t* dropped, int64_t num_q, int64_t max_bad,
                           int32_t* gate, int32_t* uncertified, cudaStream_t st) {
  escalate_decide_kernel<<<1, 1024, 0, st>>>(flags, dropped, (long long)num_q, (long long)max_bad, gate, uncertified);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_split_tf32(const float* x, int64_t rows, int64_t dim, int gallery_layout, float* out,
                      const int32_t* gate, cudaStream_t st) {
  if (rows <= 0) return SBIR_OK;
  const long long n4 = (long long)rows * (dim / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, (long long)rows, (int)dim, gallery_layout, out, gate);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

size_t col_mean_workspace_bytes(int64_t dim) { return (size_t)kColBlocks * (size_t)dim * sizeof(float); }

int launch_col_mean(const float* x, int64_t rows, int64_t dim, float* partial, float* mu, float* max_reset,
                    const int32_t* gate, cudaStream_t st) {
  if (rows <= 0) return SBIR_OK;
  const int blocks = (int)(rows < kColBlocks ? rows : kColBlocks);
  col_partial_kernel<<<blocks, 256, 0, st>>>(x, (long long)rows, (int)dim, partial, gate);
  SBIR_CHECK_LAUNCH();
  col_mean_kernel<<<(unsigned)((dim + 255) / 256), 256, 0, st>>>(partial, blocks, (long long)rows, (int)dim, mu,
                                                                max_reset, gate);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_center_split_tf32(const float* x, int64_t rows, int64_t rows_padded, int64_t dim, const float* mu,
                             int gallery_layout, float* out, float* vec, float pad_value, float* max_out,
                             const int32_t* gate, cudaStream_t st) {
  if (rows_padded <= 0) return SBIR_OK;
  long long blocks = (rows_padded + 7) / 8;
  if (blocks > 148LL * 8 * 16) blocks = 148LL * 8 * 16;
  center_split_tf32_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, (long long)rows, (long long)rows_padded, (int)dim, mu,
                                                             gallery_layout, out, vec, pad_value, max_out, gate);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_positive_distance(const void* q, int64_t num_q, const void* g, int64_t num_g, int64_t dim,
                             int dtype, int metric, const int64_t* pos_index, double* out,
                             cudaStream_t st) {
  if (num_q <= 0) return SBIR_OK;
  const bool vec = rows_vectorizable(q, dim, dtype) && rows_vectorizable(g, dim, dtype);
  const unsigned grid = (unsigned)((num_q + kRankWarps - 1) / kRankWarps);
  const long long* pi = reinterpret_cast<const long long*>(pos_index);
  if (dtype == SBIR_F32) {
    if (vec) positive_distance_kernel<float, true><<<grid, kRankWarps * 32, 0, st>>>((const float*)q, (int)num_q, (const float*)g, (int)num_g, (int)dim, metric, pi, out);
    else positive_distance_kernel<float, false><<<grid, kRankWarps * 32, 0, st>>>((const float*)q, (int)num_q, (const float*)g, (int)num_g, (int)dim, metric, pi, out);
  } else {
    using B = __nv_bfloat16;
    if (vec) positive_distance_kernel<B, true><<<grid, kRankWarps * 32, 0, st>>>((const B*)q, (int)num_q, (const B*)g, (int)num_g, (int)dim, metric, pi, out);
    else positive_distance_kernel<B, false><<<grid, kRankWarps * 32, 0, st>>>((const B*)q, (int)num_q, (const B*)g, (int)num_g, (int)dim, metric, pi, out);
  }
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_topk_merge(const float* dist, const int64_t* index, int num_lists, int64_t num_q, int k,
                      float* out_dist, int64_t* out_index, cudaStream_t st, int64_t list_stride_dist, int64_t list_stride_index) {
  if (num_q <= 0 || k <= 0) return SBIR_OK;
  const size_t stride_d = list_stride_dist > 0 ? (size_t)list_stride_dist : (size_t)num_q * k;
  const size_t stride_i = list_stride_index > 0 ? (size_t)list_stride_index : (size_t)num_q * k;
  const size_t n = (size_t)num_q * k;
  int G = 1;
  while (G < num_lists) G <<= 1;
  // tournament form when one full pass of a block (4 warps × 32/G queries) fits 32 KB of shared memory — small k, the
  // common case; long lists (k = 100 × 8 shards: 10.8 KB per query) keep the block-cooperative merge-path rounds
  // (measured 8 × 100k × 100: 584 µs vs 883 µs with mostly idle tournament groups)
  if (num_lists > 0 && num_lists <= 32 &&
      (size_t)(kMergeThreads / 32) * (32 / G) * (size_t)(num_lists + 1) * k * 12 <= 32 * 1024) {
    const size_t per_q = (size_t)(num_lists + 1) * k * 12;
    int qb = (kMergeThreads / 32) * (32 / G);                    // one pass of the block
    while ((size_t)qb * 2 * per_q <= 16 * 1024) qb *= 2;         // ~16 KB of shared memory per block
    while (qb > 1 && (size_t)qb * per_q > 48 * 1024) qb /= 2;
    const size_t smem = (size_t)qb * per_q;
    const unsigned grid = (unsigned)((num_q + qb - 1) / qb);
    const long long* idx = reinterpret_cast<const long long*>(index);
    long long* oidx = reinterpret_cast<long long*>(out_index);
#define SBIR_TOURNAMENT(GG) \
  topk_merge_tournament_kernel<GG><<<grid, kMergeThreads, smem, st>>>(dist, idx, num_lists, stride_d, stride_i, (int)num_q, k, qb, out_dist, oidx)
    switch (G) {
      case 1: SBIR_TOURNAMENT(1); break;
      case 2: SBIR_TOURNAMENT(2); break;
      case 4: SBIR_TOURNAMENT(4); break;
      case 8: SBIR_TOURNAMENT(8); break;
      case 16: SBIR_TOURNAMENT(16); break;
      default: SBIR_TOURNAMENT(32); break;
    }
#undef SBIR_TOURNAMENT
    SBIR_CHECK_LAUNCH();
    return SBIR_OK;
  }
  const size_t per_q = ((size_t)num_lists + (num_lists + 1) / 2) * k * 12;  // gathered lists + one round of merged lists
  if (num_lists > 0 && per_q <= 96 * 1024) {
    // ~16 KB of shared memory per block (a dozen blocks per SM keep loads in flight while others merge)
    int qb = (int)((16 * 1024) / per_q);
    if (qb < 1) qb = 1;
    if (qb > 32) qb = 32;
    const size_t smem = (size_t)qb * per_q;
    if (smem > 48 * 1024)
      SBIR_CUDA_TRY(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((num_q + qb - 1) / qb);
    topk_merge_kernel<<<grid, kMergeThreads, smem, st>>>(dist, reinterpret_cast<const long long*>(index), num_lists,
                                                         stride_d, stride_i, (int)num_q, k, qb, out_dist,
                                                         reinterpret_cast<long long*>(out_index));
    SBIR_CHECK_LAUNCH();
    return SBIR_OK;
  }
  fill_topk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out_dist, reinterpret_cast<long long*>(out_index), n);
  SBIR_CHECK_LAUNCH();
  if (num_lists <= 0) return SBIR_OK;
  const unsigned grid = (unsigned)((num_q + kMergeWarps - 1) / kMergeWarps);
  topk_merge_generic_kernel<<<grid, kMergeWarps * 32, 0, st>>>(dist, reinterpret_cast<const long long*>(index), num_lists,
                                                               stride_d, stride_i, (int)num_q, k, out_dist,
                                                               reinterpret_cast<long long*>(out_index));
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_fill_i32(int32_t* out, int64_t n, int32_t value, cudaStream_t st) {
  if (n <= 0) return SBIR_OK;
  fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, (long long)n, value);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_fill_i64(int64_t* out, int64_t n, int64_t value, cudaStream_t st) {
  if (n <= 0) return SBIR_OK;
  fill_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<long long*>(out), (long long)n,
                                                              (long long)value);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

int launch_retrieval_metrics(const int64_t* rank0, int64_t num_q, int k, double* out, cudaStream_t st) {
  retrieval_metrics_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const long long*>(rank0), (long long)num_q, k, out);
  SBIR_CHECK_LAUNCH();
  return SBIR_OK;
}

}  // namespace sbir
