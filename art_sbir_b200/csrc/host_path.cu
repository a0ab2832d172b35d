// host_path.cu — sbir_retrieve_host: the same retrieval pass as sbir_pairwise_topk, called
// with HOST buffers (what a caller holding numpy / torch-CPU embeddings would bind).  The
// gallery is uploaded in row chunks on a copy stream while earlier chunks are scored on the
// compute stream, so PCIe transfer overlaps the tensor-core work.  Normally the chunks are FED
// to one retrieval pass (api.cu: topk_pass_begin / feed / finish): the distance kernel is launched
// per uploaded chunk and continues the same candidate lists, so re-scoring, rank resolution and
// list warm-up happen once, not once per chunk; the first chunk is small (32 MB, doubling up to
// 1 GiB) so that scoring starts early.  When the plan cuts the gallery into several partitions
// (few queries) a multi-chunk gallery is scored chunk by chunk as shards and merged (K4).
// Device / pinned staging buffers, streams and events are cached PER DEVICE and released by
// sbir_release_host_staging; an error return drains the cached streams first.
#include <algorithm>
#include <map>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace sbir {
namespace {

struct Staging {
  void* buf = nullptr;
  size_t bytes = 0;
  void* pinned = nullptr;  // host staging for the gathered positive rows
  size_t pinned_bytes = 0;
  int device = -1;
  cudaStream_t compute = nullptr, copy = nullptr;
  std::vector<cudaEvent_t> events;
  std::mutex call_mu;  // one host-buffer call at a time PER DEVICE (calls on different devices run concurrently)
};
// One staging set PER DEVICE (streams, events and device memory belong to the device that was current when
// they were created); calls on different devices of one process do not share anything but the mutex.
std::map<int, Staging> g_staging_by_device;   // nodes are never moved: pointers / mutexes inside stay valid
std::mutex g_staging_mu;                      // guards the map itself only
thread_local Staging* tl_staging = nullptr;  // the current call's set (selected under the mutex)
#define g_staging (*tl_staging)

int select_staging() {
  int dev = 0;
  SBIR_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_staging_mu);
  Staging& s = g_staging_by_device[dev];
  s.device = dev;
  tl_staging = &s;
  return SBIR_OK;
}

int ensure_staging(size_t bytes) {
  if (g_staging.bytes < bytes) {
    if (g_staging.buf) cudaFree(g_staging.buf);
    g_staging.buf = nullptr;
    g_staging.bytes = 0;
    SBIR_CUDA_TRY(cudaMalloc(&g_staging.buf, bytes));
    g_staging.bytes = bytes;
  }
  return SBIR_OK;
}

// An early error return must not leave copies / kernels in flight on the cached streams: the next call
// would reuse the staging memory underneath them.
void drain_staging() {
  if (tl_staging == nullptr) return;
  if (g_staging.copy) cudaStreamSynchronize(g_staging.copy);
  if (g_staging.compute) cudaStreamSynchronize(g_staging.compute);
}

int ensure_streams() {
  if (!g_staging.compute) SBIR_CUDA_TRY(cudaStreamCreateWithFlags(&g_staging.compute, cudaStreamNonBlocking));
  if (!g_staging.copy) SBIR_CUDA_TRY(cudaStreamCreateWithFlags(&g_staging.copy, cudaStreamNonBlocking));
  return SBIR_OK;
}

int ensure_pinned(size_t bytes) {
  if (g_staging.pinned_bytes < bytes) {
    if (g_staging.pinned) cudaFreeHost(g_staging.pinned);
    g_staging.pinned = nullptr;
    g_staging.pinned_bytes = 0;
    SBIR_CUDA_TRY(cudaHostAlloc(&g_staging.pinned, bytes, cudaHostAllocDefault));
    g_staging.pinned_bytes = bytes;
  }
  return SBIR_OK;
}

int event_at(size_t i, cudaEvent_t* out) {
  while (g_staging.events.size() <= i) {
    cudaEvent_t e;
    SBIR_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_staging.events.push_back(e);
  }
  *out = g_staging.events[i];
  return SBIR_OK;
}

__global__ void add_i64_kernel(long long* acc, const long long* x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] += x[i];
}
__global__ void missing_rank_kernel(long long* rank, const double* pos_dist, long long n, long long missing) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && pos_dist[i] != pos_dist[i]) rank[i] = missing;
}

}  // namespace
}  // namespace sbir

using namespace sbir;

// Host utility of the sharded host path: dst[i, :] = src[index[i], :] (zero row where the index is outside
// [0, num_rows)), with a few threads — the rows of the positives a rank owns are gathered from its host-resident shard
// before they are uploaded (under torchrun the Python process runs single-threaded: the same gather through torch
// indexing took tens of milliseconds per call at cfg4).
extern "C" int sbir_gather_rows_host(const void* src, int64_t num_rows, int64_t row_bytes, const int64_t* index, int64_t n,
                                     void* dst, int threads) {
  if (num_rows < 0 || row_bytes <= 0 || n < 0) return SBIR_ERR_INVALID_ARG;
  if (n == 0) return SBIR_OK;
  if (index == nullptr || dst == nullptr || (num_rows > 0 && src == nullptr)) return SBIR_ERR_INVALID_ARG;
  auto work = [&](int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; ++i) {
      const int64_t r = index[i];
      uint8_t* out = static_cast<uint8_t*>(dst) + (size_t)i * (size_t)row_bytes;
      if (r >= 0 && r < num_rows) std::memcpy(out, static_cast<const uint8_t*>(src) + (size_t)r * (size_t)row_bytes, (size_t)row_bytes);
      else std::memset(out, 0, (size_t)row_bytes);
    }
  };
  int nt = threads > 0 ? threads : 8;
  const int64_t by_size = ((int64_t)n * row_bytes) >> 20;  // one thread per MiB moved, at least one
  if (nt > by_size) nt = (int)std::max<int64_t>(1, by_size);
  if (nt <= 1) {
    work(0, n);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, n * t / nt, n * (t + 1) / nt);
    for (auto& th : pool) th.join();
  }
  return SBIR_OK;
}

extern "C" int sbir_release_host_staging(void) {
  std::lock_guard<std::mutex> map_lock(g_staging_mu);
  int prev = 0;
  const bool have_prev = cudaGetDevice(&prev) == cudaSuccess;
  for (auto& kv : g_staging_by_device) {
    Staging& s = kv.second;
    std::lock_guard<std::mutex> lock(s.call_mu);  // waits for a call in flight on that device
    if (cudaSetDevice(kv.first) != cudaSuccess) continue;
    if (s.copy) cudaStreamSynchronize(s.copy);
    if (s.compute) cudaStreamSynchronize(s.compute);
    if (s.buf) cudaFree(s.buf);
    if (s.pinned) cudaFreeHost(s.pinned);
    for (cudaEvent_t e : s.events) cudaEventDestroy(e);
    if (s.compute) cudaStreamDestroy(s.compute);
    if (s.copy) cudaStreamDestroy(s.copy);
    s.buf = nullptr; s.bytes = 0; s.pinned = nullptr; s.pinned_bytes = 0;
    s.events.clear();
    s.compute = s.copy = nullptr;
  }
  if (have_prev) cudaSetDevice(prev);
  return SBIR_OK;
}

static int retrieve_host_locked(const void* q_host, int64_t num_q, const void* g_host, int64_t num_g,
                                int64_t dim, int dtype, int metric, int k, const int64_t* pos_index_host,
                                float* out_dist_host, int64_t* out_index_host, int64_t* out_rank_host,
                                int32_t* out_uncertified_host);

extern "C" int sbir_retrieve_host(const void* q_host, int64_t num_q, const void* g_host, int64_t num_g,
                                  int64_t dim, int dtype, int metric, int k, const int64_t* pos_index_host,
                                  float* out_dist_host, int64_t* out_index_host, int64_t* out_rank_host,
                                  int32_t* out_uncertified_host) {
  if (num_q <= 0 || num_g <= 0 || dim <= 0 || k <= 0) return SBIR_ERR_INVALID_ARG;
  if (q_host == nullptr || g_host == nullptr || out_dist_host == nullptr || out_index_host == nullptr)
    return SBIR_ERR_INVALID_ARG;
  if (out_rank_host != nullptr && pos_index_host == nullptr) return SBIR_ERR_INVALID_ARG;
  if (dtype != SBIR_F32 && dtype != SBIR_BF16) return SBIR_ERR_INVALID_ARG;
  SBIR_TRY(select_staging());
  std::lock_guard<std::mutex> lock(g_staging.call_mu);
  const int status = retrieve_host_locked(q_host, num_q, g_host, num_g, dim, dtype, metric, k, pos_index_host, out_dist_host,
                                          out_index_host, out_rank_host, out_uncertified_host);
  if (status != SBIR_OK) drain_staging();
  return status;
}

static int retrieve_host_locked(const void* q_host, int64_t num_q, const void* g_host, int64_t num_g,
                                int64_t dim, int dtype, int metric, int k, const int64_t* pos_index_host,
                                float* out_dist_host, int64_t* out_index_host, int64_t* out_rank_host,
                                int32_t* out_uncertified_host) {
  const bool want_rank = out_rank_host != nullptr;
  const size_t es = elem_size(dtype);
  const size_t row_bytes = (size_t)dim * es;

  // Upload schedule: row counts of the gallery chunks.  Streamed feeds need chunk ends on the
  // distance kernel's chunk-step boundaries (`granule` rows; 0 = the plan has several gallery
  // partitions and cannot be fed in pieces).
  const int64_t granule = [&]() -> int64_t {
    const K1Plan plan = topk_primary_plan(num_q, num_g, dim, k, dtype);  // the plan topk_pass_begin will make
    return plan.num_splits == 1 ? (int64_t)plan.tiles_per_chunk * kTileG : 0;
  }();
  std::vector<int64_t> chunk_end;
  {
    int64_t forced = 0;
    forced = debug_options().host_chunk_rows;  // test hook: small chunks
    const int64_t unit = granule > 0 ? granule : kTileG;
    size_t target = granule > 0 ? (size_t(32) << 20) : (size_t(1) << 30);
    int64_t next = 0;
    while (next < num_g) {
      int64_t rows = forced > 0 ? forced : (int64_t)(target / row_bytes);
      rows = std::max<int64_t>(unit, rows / unit * unit);
      int64_t end = next + rows;
      if (end + unit / 2 >= num_g) end = num_g;  // do not leave a sliver for a last chunk
      chunk_end.push_back(end);
      next = end;
      target = std::min<size_t>(target * 2, size_t(1) << 30);
    }
  }
  const int num_chunks = (int)chunk_end.size();
  const bool streamed = granule > 0 || num_chunks == 1;  // one pass fed chunk by chunk (or all at once)

  // Device layout: Q | G (whole gallery, chunks land in place) | gathered positives |
  // per-chunk top-k lists (shard mode) | merged outputs | rank accumulators | workspace.
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o = align_up(o + (bytes ? bytes : 1), 256); return r; };
  const size_t off_q = take((size_t)num_q * row_bytes);
  const size_t off_g = take((size_t)num_g * row_bytes);
  const int num_lists = streamed ? 0 : num_chunks;
  const size_t off_lists_d = take((size_t)num_lists * num_q * k * sizeof(float));
  const size_t off_lists_i = take((size_t)num_lists * num_q * k * sizeof(int64_t));
  const size_t off_out_d = take((size_t)num_q * k * sizeof(float));
  const size_t off_out_i = take((size_t)num_q * k * sizeof(int64_t));
  const size_t off_uncert = take(sizeof(int32_t) * (size_t)(num_chunks + 1));
  size_t off_pos = 0, off_posidx = 0, off_posglobal = 0, off_pos_dist = 0, off_rank = 0, off_cnt = 0;
  if (want_rank) {
    off_pos = take((size_t)num_q * row_bytes);
    off_posidx = take((size_t)num_q * sizeof(int64_t));
    off_posglobal = take((size_t)num_q * sizeof(int64_t));
    off_pos_dist = take((size_t)num_q * sizeof(double));
    off_rank = take((size_t)num_q * sizeof(int64_t));
    off_cnt = take((size_t)num_q * sizeof(int64_t));
  }
  // shard mode scores every uploaded chunk as a gallery of its own: the workspace must fit the LARGEST layout, which
  // is not necessarily the one of the longest chunk (a short last chunk may be planned with more partitions)
  size_t ws_bytes = 0;
  if (streamed) {
    ws_bytes = sbir_pairwise_topk_workspace_bytes(num_q, num_g, dim, k, dtype, metric, want_rank ? 1 : 0);
  } else {
    for (int c = 0; c < num_chunks; ++c)
      ws_bytes = std::max(ws_bytes, sbir_pairwise_topk_workspace_bytes(num_q, chunk_end[c] - (c ? chunk_end[c - 1] : 0), dim, k, dtype,
                                                                        metric, want_rank ? 1 : 0));
  }
  if (ws_bytes == 0) return SBIR_ERR_UNSUPPORTED;
  const size_t off_ws = take(ws_bytes);
  SBIR_TRY(ensure_staging(o));
  SBIR_TRY(ensure_streams());
  uint8_t* base = static_cast<uint8_t*>(g_staging.buf);
  cudaStream_t cs = g_staging.compute, xs = g_staging.copy;
  cudaEvent_t ev = nullptr;

  uint8_t* d_q = base + off_q;
  uint8_t* d_g = base + off_g;
  SBIR_CUDA_TRY(cudaMemcpyAsync(d_q, q_host, (size_t)num_q * row_bytes, cudaMemcpyHostToDevice, xs));

  double* d_pos_dist = nullptr;
  long long* d_rank = nullptr;
  const int64_t* d_pos_global = nullptr;
  if (want_rank) {
    // The positive's row may sit in any chunk: gather those rows on the host once (pinned buffer,
    // a few threads), upload them, and evaluate d(q, pos) up front so every chunk can count
    // against it.
    const size_t gather_bytes = align_up((size_t)num_q * row_bytes, 256);
    SBIR_TRY(ensure_pinned(gather_bytes + (size_t)num_q * sizeof(int64_t)));
    uint8_t* gathered = static_cast<uint8_t*>(g_staging.pinned);
    int64_t* ident = reinterpret_cast<int64_t*>(gathered + gather_bytes);
    auto gather = [&](int64_t i0, int64_t i1) {
      for (int64_t i = i0; i < i1; ++i) {
        const int64_t pi = pos_index_host[i];
        if (pi >= 0 && pi < num_g) {
          std::memcpy(gathered + (size_t)i * row_bytes, static_cast<const uint8_t*>(g_host) + (size_t)pi * row_bytes, row_bytes);
          ident[i] = i;
        } else {
          std::memset(gathered + (size_t)i * row_bytes, 0, row_bytes);
          ident[i] = -1;
        }
      }
    };
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(8, (int64_t)((size_t)num_q * row_bytes >> 22)));
    if (nthreads <= 1) {
      gather(0, num_q);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < nthreads; ++t)
        pool.emplace_back(gather, num_q * t / nthreads, num_q * (t + 1) / nthreads);
      for (auto& th : pool) th.join();
    }
    SBIR_CUDA_TRY(cudaMemcpyAsync(base + off_pos, gathered, (size_t)num_q * row_bytes, cudaMemcpyHostToDevice, xs));
    SBIR_CUDA_TRY(cudaMemcpyAsync(base + off_posidx, ident, (size_t)num_q * sizeof(int64_t), cudaMemcpyHostToDevice, xs));
    SBIR_CUDA_TRY(cudaMemcpyAsync(base + off_posglobal, pos_index_host, (size_t)num_q * sizeof(int64_t), cudaMemcpyHostToDevice, xs));
    d_pos_global = reinterpret_cast<const int64_t*>(base + off_posglobal);
    d_pos_dist = reinterpret_cast<double*>(base + off_pos_dist);
    d_rank = reinterpret_cast<long long*>(base + off_rank);
    SBIR_TRY(launch_positive_distance(d_q, num_q, base + off_pos, num_q, dim, dtype, metric,
                                      reinterpret_cast<const int64_t*>(base + off_posidx), d_pos_dist, xs));
    if (!streamed) SBIR_CUDA_TRY(cudaMemsetAsync(d_rank, 0, (size_t)num_q * sizeof(int64_t), xs));
  }
  SBIR_TRY(event_at((size_t)num_chunks, &ev));
  SBIR_CUDA_TRY(cudaEventRecord(ev, xs));
  SBIR_CUDA_TRY(cudaStreamWaitEvent(cs, ev, 0));

  float* d_out_d = reinterpret_cast<float*>(base + off_out_d);
  int64_t* d_out_i = reinterpret_cast<int64_t*>(base + off_out_i);
  int32_t* d_uncert = reinterpret_cast<int32_t*>(base + off_uncert);
  auto upload_chunk = [&](int c) -> int {
    const int64_t r0 = c ? chunk_end[c - 1] : 0;
    SBIR_CUDA_TRY(cudaMemcpyAsync(d_g + (size_t)r0 * row_bytes, static_cast<const uint8_t*>(g_host) + (size_t)r0 * row_bytes,
                                  (size_t)(chunk_end[c] - r0) * row_bytes, cudaMemcpyHostToDevice, xs));
    SBIR_TRY(event_at((size_t)c, &ev));
    SBIR_CUDA_TRY(cudaEventRecord(ev, xs));
    SBIR_CUDA_TRY(cudaStreamWaitEvent(cs, ev, 0));
    return SBIR_OK;
  };
  if (streamed) {
    TopkPass pass;
    SBIR_TRY(topk_pass_begin(pass, d_q, num_q, d_g, nullptr, num_g, dim, dtype, metric, k, /*index_offset=*/0, nullptr, d_pos_dist,
                             d_pos_global, /*tie_offset=*/0, d_out_d, d_out_i,
                             want_rank ? reinterpret_cast<int64_t*>(d_rank) : nullptr, /*missing_rank=*/num_g,
                             d_uncert + 1, base + off_ws, ws_bytes, cs));
    for (int c = 0; c < num_chunks; ++c) {
      SBIR_TRY(upload_chunk(c));
      SBIR_TRY(topk_pass_feed(pass, chunk_end[c]));
    }
    SBIR_TRY(topk_pass_finish(pass));
  } else {
    float* d_lists_d = reinterpret_cast<float*>(base + off_lists_d);
    int64_t* d_lists_i = reinterpret_cast<int64_t*>(base + off_lists_i);
    for (int c = 0; c < num_chunks; ++c) {
      const int64_t r0 = c ? chunk_end[c - 1] : 0;
      const int64_t rows = chunk_end[c] - r0;
      SBIR_TRY(upload_chunk(c));
      int64_t* d_cnt = want_rank ? reinterpret_cast<int64_t*>(base + off_cnt) : nullptr;
      SBIR_TRY(sbir_pairwise_topk_shard(d_q, num_q, d_g + (size_t)r0 * row_bytes, nullptr, rows, dim, dtype, metric, k, r0,
                                        d_pos_dist, d_pos_global, d_lists_d + (size_t)c * num_q * k,
                                        d_lists_i + (size_t)c * num_q * k, d_cnt, d_uncert + 1 + c,
                                        base + off_ws, ws_bytes, cs));
      if (want_rank) {
        add_i64_kernel<<<(unsigned)((num_q + 255) / 256), 256, 0, cs>>>(d_rank, reinterpret_cast<long long*>(d_cnt), num_q);
        SBIR_CHECK_LAUNCH();
      }
    }
    SBIR_TRY(launch_topk_merge(d_lists_d, d_lists_i, num_chunks, num_q, k, d_out_d, d_out_i, cs));
    if (want_rank) {
      missing_rank_kernel<<<(unsigned)((num_q + 255) / 256), 256, 0, cs>>>(d_rank, d_pos_dist, num_q, num_g);
      SBIR_CHECK_LAUNCH();
    }
  }
  if (want_rank)
    SBIR_CUDA_TRY(cudaMemcpyAsync(out_rank_host, d_rank, (size_t)num_q * sizeof(int64_t), cudaMemcpyDeviceToHost, cs));
  SBIR_CUDA_TRY(cudaMemcpyAsync(out_dist_host, d_out_d, (size_t)num_q * k * sizeof(float), cudaMemcpyDeviceToHost, cs));
  SBIR_CUDA_TRY(cudaMemcpyAsync(out_index_host, d_out_i, (size_t)num_q * k * sizeof(int64_t), cudaMemcpyDeviceToHost, cs));
  std::vector<int32_t> unc((size_t)num_chunks + 1, 0);
  SBIR_CUDA_TRY(cudaMemcpyAsync(unc.data(), d_uncert, unc.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
  SBIR_CUDA_TRY(cudaStreamSynchronize(cs));
  if (out_uncertified_host) {
    int32_t total = 0;
    for (int c = 0; c < (streamed ? 1 : num_chunks); ++c) total += unc[1 + c];
    *out_uncertified_host = total;
  }
  return SBIR_OK;
}

// One rank's part of a gallery-sharded retrieval with the shard in HOST memory: like
// sbir_pairwise_topk_shard, but the shard's rows are uploaded in chunks on a copy stream and fed
// to the pass as they arrive (same staging cache as sbir_retrieve_host).  Queries, the positives'
// distances / global indices and the outputs are DEVICE buffers; work is ordered after `stream`
// and the call returns when the outputs are complete.
static int retrieve_host_shard_locked(const void* q_dev, int64_t num_q, const void* g_host, int64_t num_g,
                                      int64_t dim, int dtype, int metric, int k, int64_t index_offset,
                                      const double* pos_dist_dev, const int64_t* pos_index_global_dev,
                                      float* out_dist_dev, int64_t* out_index_dev, int64_t* out_count_less_dev,
                                      int32_t* out_uncertified_host, void* stream);

extern "C" int sbir_retrieve_host_shard(const void* q_dev, int64_t num_q, const void* g_host, int64_t num_g,
                                        int64_t dim, int dtype, int metric, int k, int64_t index_offset,
                                        const double* pos_dist_dev, const int64_t* pos_index_global_dev,
                                        float* out_dist_dev, int64_t* out_index_dev, int64_t* out_count_less_dev,
                                        int32_t* out_uncertified_host, void* stream) {
  if (num_q <= 0 || num_g < 0 || dim <= 0 || k <= 0) return SBIR_ERR_INVALID_ARG;
  if (q_dev == nullptr || out_dist_dev == nullptr || out_index_dev == nullptr) return SBIR_ERR_INVALID_ARG;
  if (num_g > 0 && g_host == nullptr) return SBIR_ERR_INVALID_ARG;
  if (out_count_less_dev != nullptr && pos_dist_dev == nullptr) return SBIR_ERR_INVALID_ARG;
  if (dtype != SBIR_F32 && dtype != SBIR_BF16) return SBIR_ERR_INVALID_ARG;
  SBIR_TRY(select_staging());
  std::lock_guard<std::mutex> lock(g_staging.call_mu);
  const int status = retrieve_host_shard_locked(q_dev, num_q, g_host, num_g, dim, dtype, metric, k, index_offset, pos_dist_dev,
                                                pos_index_global_dev, out_dist_dev, out_index_dev, out_count_less_dev,
                                                out_uncertified_host, stream);
  if (status != SBIR_OK) drain_staging();
  return status;
}

static int retrieve_host_shard_locked(const void* q_dev, int64_t num_q, const void* g_host, int64_t num_g,
                                      int64_t dim, int dtype, int metric, int k, int64_t index_offset,
                                      const double* pos_dist_dev, const int64_t* pos_index_global_dev,
                                      float* out_dist_dev, int64_t* out_index_dev, int64_t* out_count_less_dev,
                                      int32_t* out_uncertified_host, void* stream) {
  const bool want_rank = out_count_less_dev != nullptr;
  const size_t row_bytes = (size_t)dim * elem_size(dtype);
  int64_t granule = 0;
  if (num_g > 0) {
    const K1Plan plan = topk_primary_plan(num_q, num_g, dim, k, dtype);
    granule = plan.num_splits == 1 ? (int64_t)plan.tiles_per_chunk * kTileG : 0;
  }
  std::vector<int64_t> chunk_end;
  if (granule > 0) {
    size_t target = size_t(32) << 20;
    if (debug_options().host_chunk_rows > 0) target = (size_t)debug_options().host_chunk_rows * row_bytes;  // test hook: small chunks
    int64_t next = 0;
    while (next < num_g) {
      const int64_t rows = std::max<int64_t>(granule, (int64_t)(target / row_bytes) / granule * granule);
      int64_t end = next + rows;
      if (end + granule / 2 >= num_g) end = num_g;
      chunk_end.push_back(end);
      next = end;
      target = std::min<size_t>(target * 2, size_t(1) << 30);
    }
  } else if (num_g > 0) {
    chunk_end.push_back(num_g);  // several gallery partitions: upload everything, then one feed
  }
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o = align_up(o + (bytes ? bytes : 1), 256); return r; };
  const size_t off_g = take((size_t)num_g * row_bytes);
  const size_t off_uncert = take(sizeof(int32_t));
  const size_t ws_bytes = sbir_pairwise_topk_workspace_bytes(num_q, num_g, dim, k, dtype, metric, want_rank ? 1 : 0);
  if (ws_bytes == 0) return SBIR_ERR_UNSUPPORTED;
  const size_t off_ws = take(ws_bytes);
  SBIR_TRY(ensure_staging(o));
  SBIR_TRY(ensure_streams());
  uint8_t* base = static_cast<uint8_t*>(g_staging.buf);
  cudaStream_t cs = g_staging.compute, xs = g_staging.copy;
  cudaEvent_t ev = nullptr;
  // everything here runs after the work already queued on the caller's stream
  SBIR_TRY(event_at(chunk_end.size(), &ev));
  SBIR_CUDA_TRY(cudaEventRecord(ev, static_cast<cudaStream_t>(stream)));
  SBIR_CUDA_TRY(cudaStreamWaitEvent(cs, ev, 0));
  uint8_t* d_g = base + off_g;
  int32_t* d_uncert = reinterpret_cast<int32_t*>(base + off_uncert);
  TopkPass pass;
  // A query without a positive anywhere (NaN pos_dist) contributes a local count of 0.
  SBIR_TRY(topk_pass_begin(pass, q_dev, num_q, d_g, nullptr, num_g, dim, dtype, metric, k, index_offset, nullptr, pos_dist_dev,
                           pos_index_global_dev, index_offset, out_dist_dev, out_index_dev, out_count_less_dev,
                           /*missing_rank=*/0, d_uncert, base + off_ws, ws_bytes, cs));
  for (size_t c = 0; c < chunk_end.size(); ++c) {
    const int64_t r0 = c ? chunk_end[c - 1] : 0;
    SBIR_CUDA_TRY(cudaMemcpyAsync(d_g + (size_t)r0 * row_bytes, static_cast<const uint8_t*>(g_host) + (size_t)r0 * row_bytes,
                                  (size_t)(chunk_end[c] - r0) * row_bytes, cudaMemcpyHostToDevice, xs));
    SBIR_TRY(event_at(c, &ev));
    SBIR_CUDA_TRY(cudaEventRecord(ev, xs));
    SBIR_CUDA_TRY(cudaStreamWaitEvent(cs, ev, 0));
    SBIR_TRY(topk_pass_feed(pass, chunk_end[c]));
  }
  SBIR_TRY(topk_pass_finish(pass));
  int32_t unc = 0;
  SBIR_CUDA_TRY(cudaMemcpyAsync(&unc, d_uncert, sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
  SBIR_CUDA_TRY(cudaStreamSynchronize(cs));
  if (out_uncertified_host) *out_uncertified_host = unc;
  return SBIR_OK;
}
