// common.cuh — shared definitions for the sbir_b200 kernels: status plumbing, element
// loaders (fp32 / bf16 rows, 16-byte vectors), warp reductions, and the EXACT distance
// functions that define parity with the reference (utils.py:31-42 of the reference).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/sbir_b200.h"

namespace sbir {

constexpr float kPairwiseEps = 1e-6f;  // nn.PairwiseDistance default eps (utils.py:42)
constexpr float kCosineEps = 1e-8f;    // nn.CosineSimilarity default eps (utils.py:34)
constexpr unsigned kFullMask = 0xffffffffu;

// Thread-local record of the last failing CUDA call (sbir_last_cuda_error).
void set_last_cuda_error(int err);
#define SBIR_CUDA_TRY(expr)                                  \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) {                                 \
      ::sbir::set_last_cuda_error(static_cast<int>(_e));     \
      return SBIR_ERR_CUDA;                                  \
    }                                                        \
  } while (0)
#define SBIR_TRY(expr)                  \
  do {                                  \
    int _s = (expr);                    \
    if (_s != SBIR_OK) return _s;       \
  } while (0)
// Kernel launches: catch configuration errors immediately (no device sync) and count them
// (sbir_profile_collect reports the count; bench.py's `gpu_launches`).
void count_kernel_launch();
#define SBIR_CHECK_LAUNCH()              \
  do {                                   \
    ::sbir::count_kernel_launch();       \
    SBIR_CUDA_TRY(cudaGetLastError());   \
  } while (0)
// Profiling hooks around the K1 launch (no-ops unless sbir_profile_enable(1) was called).
void profile_k1_begin(cudaStream_t st);
void profile_k1_end(cudaStream_t st);

inline size_t elem_size(int dtype) { return dtype == SBIR_BF16 ? 2 : 4; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- reductions ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}

// ------------------------------------------------------------- element access ----
// A 16-byte vector of embedding elements, unpacked to fp32.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int kElems = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int kElems = 8;
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void from_f32(float& d, float x) { d = x; }
__device__ __forceinline__ void from_f32(__nv_bfloat16& d, float x) { d = __float2bfloat16_rn(x); }

// True when rows of `dim` elements starting at `base` can be read as 16-byte vectors.
inline bool rows_vectorizable(const void* base, int64_t dim, int dtype) {
  return (reinterpret_cast<uintptr_t>(base) % 16 == 0) && ((dim * (int64_t)elem_size(dtype)) % 16 == 0);
}

// -------------------------------------------------- exact reference distances ----
// These restate, element for element, what torch evaluates for the reference's
//   utils.euclidean_distance = nn.PairwiseDistance(p=2)            (utils.py:42)
//       d = sqrt( sum_i ( fl32( fl32(x_i - y_i) + 1e-6 ) )^2 )
//   utils.cosine_distance = 1 - nn.CosineSimilarity(dim=1)          (utils.py:31-40)
//       d = 1 - sum_i fl32( fl32(x_i / max(||x||,1e-8)) * fl32(y_i / max(||y||,1e-8)) )
// with the per-element arithmetic in fp32 exactly as torch does it and the long sum in
// fp64, so the result is within an ulp or two of torch's fp32 reduction (and closer to
// the real-number value than torch's own fp32 sum).  One warp evaluates one pair;
// every lane returns the full sum.
template <typename T, bool kVec>
__device__ __forceinline__ double warp_sq_l2_eps(const T* __restrict__ x, const T* __restrict__ y,
                                                 int dim, int lane) {
  double acc = 0.0;
  if constexpr (kVec) {
    constexpr int E = Vec16<T>::kElems;
    const int nvec = dim / E;
#pragma unroll 4
    for (int i = lane; i < nvec; i += 32) {
      Vec16<T> a, b;
      a.load(x + (size_t)i * E);
      b.load(y + (size_t)i * E);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float t = __fadd_rn(__fsub_rn(a.v[e], b.v[e]), kPairwiseEps);
        acc += (double)t * (double)t;
      }
    }
  } else {
    for (int i = lane; i < dim; i += 32) {
      const float t = __fadd_rn(__fsub_rn(to_f32(x[i]), to_f32(y[i])), kPairwiseEps);
      acc += (double)t * (double)t;
    }
  }
  return warp_sum(acc);
}

template <typename T, bool kVec>
__device__ __forceinline__ double warp_sq_norm(const T* __restrict__ x, int dim, int lane) {
  double acc = 0.0;
  if constexpr (kVec) {
    constexpr int E = Vec16<T>::kElems;
    const int nvec = dim / E;
#pragma unroll 4
    for (int i = lane; i < nvec; i += 32) {
      Vec16<T> a;
      a.load(x + (size_t)i * E);
#pragma unroll
      for (int e = 0; e < E; ++e) acc += (double)a.v[e] * (double)a.v[e];
    }
  } else {
    for (int i = lane; i < dim; i += 32) {
      const float t = to_f32(x[i]);
      acc += (double)t * (double)t;
    }
  }
  return warp_sum(acc);
}

// Σ fl32(fl32(x_i / cx) * fl32(y_i / cy)) with cx, cy the clamped fp32 norms.
template <typename T, bool kVec>
__device__ __forceinline__ double warp_cos_dot(const T* __restrict__ x, const T* __restrict__ y,
                                               float cx, float cy, int dim, int lane) {
  double acc = 0.0;
  if constexpr (kVec) {
    constexpr int E = Vec16<T>::kElems;
    const int nvec = dim / E;
#pragma unroll 4
    for (int i = lane; i < nvec; i += 32) {
      Vec16<T> a, b;
      a.load(x + (size_t)i * E);
      b.load(y + (size_t)i * E);
#pragma unroll
      for (int e = 0; e < E; ++e)
        acc += (double)__fmul_rn(__fdiv_rn(a.v[e], cx), __fdiv_rn(b.v[e], cy));
    }
  } else {
    for (int i = lane; i < dim; i += 32)
      acc += (double)__fmul_rn(__fdiv_rn(to_f32(x[i]), cx), __fdiv_rn(to_f32(y[i]), cy));
  }
  return warp_sum(acc);
}

// fp32 norm as torch's linalg_vector_norm returns it, clamped like cosine_similarity does.
__device__ __forceinline__ float clamped_norm(double sq) {
  return fmaxf((float)sqrt(sq), kCosineEps);
}

// Exact distance between rows x and y (one warp). metric: SBIR_EUCLIDEAN / SBIR_COSINE.
template <typename T, bool kVec>
__device__ __forceinline__ double warp_exact_distance(const T* __restrict__ x,
                                                      const T* __restrict__ y, int dim, int metric,
                                                      int lane) {
  if (metric == SBIR_EUCLIDEAN) return sqrt(warp_sq_l2_eps<T, kVec>(x, y, dim, lane));
  const float cx = clamped_norm(warp_sq_norm<T, kVec>(x, dim, lane));
  const float cy = clamped_norm(warp_sq_norm<T, kVec>(y, dim, lane));
  return 1.0 - warp_cos_dot<T, kVec>(x, y, cx, cy, dim, lane);
}

// e-space value of an exact distance (what K1's epilogue approximates) and the bound on
// |approx − exact| for one query.
__device__ __forceinline__ double e_of_distance(double d, int metric, float qsq) {
  if (metric == SBIR_EUCLIDEAN) return d * d - (double)qsq;
  return (d - 1.0) * (double)fmaxf(sqrtf(qsq), kCosineEps);
}
// q_res = ‖q − bf16(q)‖, g_res_abs / g_res_rel = max_j ‖g_j − bf16(g_j)‖ (absolute / relative to max(‖g_j‖,eps)):
// non-zero when fp32 embeddings were SELECTED on their bf16-rounded copies (rowops.cu: convert_bf16_norm_kernel);
// then q·g − qh·gh = qh·gl + ql·g is bounded by Cauchy-Schwarz on the measured residual norms.
__device__ __forceinline__ double e_margin(int metric, float qsq, float gsq_max, float kappa, int dim, float q_res = 0.f,
                                           float g_res_abs = 0.f, float g_res_rel = 0.f) {
  const double nq = sqrt((double)qsq);
  if (metric == SBIR_EUCLIDEAN) {
    const double s = (double)qsq + (double)gsq_max;
    const double split = 2.0 * (1.01 * nq * (double)g_res_abs + (double)q_res * sqrt((double)gsq_max) + (double)q_res * (double)g_res_abs);
    // tensor-core rounding of 2·q·g  +  the reference's +1e-6 per component  +  fp32 epilogue rounding
    return (double)kappa * s + split + 4e-6 * sqrt((double)dim * s) + 1e-12 * (double)dim + 4e-7 * s + 1e-30;
  }
  const double split = 1.01 * nq * (double)g_res_rel + (double)q_res * (1.0 + (double)g_res_rel);
  return (double)kappa * nq + split + 1e-6 * nq + 1e-30;
}

// Total order used everywhere a ranked list is produced: ascending distance, ties by
// ascending gallery index (the reference inherits torch.topk's unspecified tie order).
__device__ __forceinline__ bool ranks_before(double da, int64_t ia, double db, int64_t ib) {
  return da < db || (da == db && ia < ib);
}

}  // namespace sbir
