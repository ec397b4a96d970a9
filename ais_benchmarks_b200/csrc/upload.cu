// Host -> HBM upload of the arrays the reference-facing classes are handed: numpy float64 (README.md:41-43 of the
// reference; CBenchmark.evaluate_method passes float64 everywhere) or float32, in PAGEABLE memory, wanted as float32 in
// HBM. A pageable cudaMemcpy is staged by the driver through one pinned buffer on one thread (measured here: ~7 GB/s,
// 13 ms for the 128 MB an e2e step of BASELINE configs[3] moves, 24 % of the step on a rank of an 8-GPU job). This
// uploader does the staging itself, on several host threads: each narrows its chunks to float32 straight into its own
// pinned staging buffers (so the host link carries 4-byte values and no cast kernel runs) and queues the DMA of one
// buffer while it fills the other. aisb_download_f32 is the way back (float32 in HBM -> the float64 numpy arrays the
// reference's API returns). No kernels here.
#include <stdint.h>

#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace aisb {
namespace {

constexpr int kUploadMaxThreads = 8;
constexpr int64_t kStageElems = 1 << 20;   // floats per staging buffer (4 MB): ~100 us of DMA, ~300 us of narrowing
constexpr int kStageBufs = 2;

struct Lane {
  float* stage[kStageBufs] = {nullptr, nullptr};
  cudaEvent_t drained[kStageBufs] = {nullptr, nullptr};   // the DMA out of stage[b] has completed
  cudaEvent_t finished = nullptr;
  cudaStream_t stream = nullptr;
  bool ready = false;
};

struct Pool {
  Lane lanes[kUploadMaxThreads];
  cudaEvent_t start = nullptr;
};

std::mutex g_upload_mu;
Pool g_pools[64];   // per device, created on first use, kept for the life of the process

cudaError_t lane_init(Lane& l) {
  if (l.ready) return cudaSuccess;
  cudaError_t e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) return e;
  for (int b = 0; b < kStageBufs; ++b) {
    e = cudaHostAlloc(reinterpret_cast<void**>(&l.stage[b]), sizeof(float) * kStageElems, cudaHostAllocDefault);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&l.drained[b], cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  e = cudaEventCreateWithFlags(&l.finished, cudaEventDisableTiming);
  if (e != cudaSuccess) return e;
  l.ready = true;
  return cudaSuccess;
}

// Lane t runs on its own thread (lane 0 on the caller's). No exception may cross the C ABI: a lane whose thread cannot
// be started (resource limits) runs on the calling thread instead, after the others.
template <class Fn>
void run_lanes(int nlanes, cudaError_t* errs, Fn&& work) {
  std::vector<std::thread> workers;
  std::vector<int> inline_lanes;
  workers.reserve(nlanes > 0 ? nlanes - 1 : 0);
  for (int t = 1; t < nlanes; ++t) {
    try {
      workers.emplace_back([&errs, &work, t] { errs[t] = work(t); });
    } catch (...) {
      inline_lanes.push_back(t);
    }
  }
  errs[0] = work(0);
  for (int t : inline_lanes) errs[t] = work(t);
  for (auto& w : workers) w.join();
}

// Elements per chunk: a lane should see several chunks (so that the DMA of one overlaps the narrowing of the next) but
// no chunk smaller than 64 K elements (per-copy overheads) or larger than a staging buffer. A rank of an 8-GPU job
// stages 1M elements with 2 threads: one 1M-element chunk would leave the second thread idle and nothing overlapped.
int64_t chunk_elems(int64_t count, int threads) {
  int64_t c = count / (static_cast<int64_t>(threads < 1 ? 1 : threads) * 4);
  c = (c + 4095) / 4096 * 4096;
  if (c < (1 << 16)) c = 1 << 16;
  if (c > kStageElems) c = kStageElems;
  return c;
}

template <class T>
void narrow(const T* __restrict__ src, float* __restrict__ dst, int64_t count) {
  for (int64_t i = 0; i < count; ++i) dst[i] = static_cast<float>(src[i]);
}

// chunks lane, lane + nlanes, ... of the array
template <class T>
cudaError_t lane_run(Lane& l, int dev, int lane, int nlanes, const T* h_src, int64_t count, int64_t chunk,
                     float* d_dst) {
  cudaError_t e = cudaSetDevice(dev);
  if (e != cudaSuccess) return e;
  const int64_t nchunks = (count + chunk - 1) / chunk;
  int used = 0;
  for (int64_t c = lane; c < nchunks; c += nlanes, ++used) {
    const int b = used % kStageBufs;
    if (used >= kStageBufs) {
      e = cudaEventSynchronize(l.drained[b]);
      if (e != cudaSuccess) return e;
    }
    const int64_t off = c * chunk;
    const int64_t len = count - off < chunk ? count - off : chunk;
    narrow(h_src + off, l.stage[b], len);
    e = cudaMemcpyAsync(d_dst + off, l.stage[b], sizeof(float) * len, cudaMemcpyHostToDevice, l.stream);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(l.drained[b], l.stream);
    if (e != cudaSuccess) return e;
  }
  // the staging buffers must be reusable by the next call without further waits
  return cudaStreamSynchronize(l.stream);
}

template <class T>
int upload_t(const T* h_src, int64_t count, float* d_dst, int threads, cudaStream_t st) {
  int dev = 0;
  AISB_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    set_error("upload: device index %d out of range", dev);
    return AISB_ERR_INVALID;
  }
  int nlanes = threads < 1 ? 1 : (threads > kUploadMaxThreads ? kUploadMaxThreads : threads);
  const int64_t chunk = chunk_elems(count, nlanes);
  const int64_t nchunks = (count + chunk - 1) / chunk;
  if (nlanes > nchunks) nlanes = static_cast<int>(nchunks);
  std::lock_guard<std::mutex> lock(g_upload_mu);
  Pool& pool = g_pools[dev];
  if (!pool.start) AISB_CUDA_CHECK(cudaEventCreateWithFlags(&pool.start, cudaEventDisableTiming));
  for (int t = 0; t < nlanes; ++t) AISB_CUDA_CHECK(lane_init(pool.lanes[t]));
  // d_dst may be memory the caller's stream is still working on: the copies queue behind it ...
  AISB_CUDA_CHECK(cudaEventRecord(pool.start, st));
  for (int t = 0; t < nlanes; ++t) AISB_CUDA_CHECK(cudaStreamWaitEvent(pool.lanes[t].stream, pool.start, 0));
  cudaError_t errs[kUploadMaxThreads];
  for (int t = 0; t < nlanes; ++t) errs[t] = cudaSuccess;
  run_lanes(nlanes, errs, [&](int t) { return lane_run(pool.lanes[t], dev, t, nlanes, h_src, count, chunk, d_dst); });
  for (int t = 0; t < nlanes; ++t)
    if (errs[t] != cudaSuccess) {
      for (int u = 0; u < nlanes; ++u) cudaStreamSynchronize(pool.lanes[u].stream);   // leave no copy in flight
      return cuda_fail(errs[t], "upload lane");
    }
  // ... and whatever the caller queues next on its stream sees the data (the lanes have drained: this only orders it)
  for (int t = 0; t < nlanes; ++t) {
    AISB_CUDA_CHECK(cudaEventRecord(pool.lanes[t].finished, pool.lanes[t].stream));
    AISB_CUDA_CHECK(cudaStreamWaitEvent(st, pool.lanes[t].finished, 0));
  }
  return AISB_OK;
}

template <class T>
void widen(const float* __restrict__ src, T* __restrict__ dst, int64_t count) {
  for (int64_t i = 0; i < count; ++i) dst[i] = static_cast<T>(src[i]);
}

// The way back: chunks lane, lane + nlanes, ... of a device float32 array into a host array (float64 or float32). The
// DMA of chunk k+1 into one staging buffer runs while the lane widens chunk k out of the other -- and touches the
// destination pages for the first time (a fresh numpy array is untouched memory: one thread faulting in 2.5 GB is
// most of a second, which is what the single-threaded copy behind Tensor.cpu() pays).
template <class T>
cudaError_t lane_fetch(Lane& l, int dev, int lane, int nlanes, const float* d_src, int64_t count, int64_t chunk,
                       T* h_dst) {
  cudaError_t e = cudaSetDevice(dev);
  if (e != cudaSuccess) return e;
  const int64_t nchunks = (count + chunk - 1) / chunk;
  auto span = [&](int64_t c, int64_t& off, int64_t& len) {
    off = c * chunk;
    len = count - off < chunk ? count - off : chunk;
  };
  auto fetch = [&](int64_t c, int b) -> cudaError_t {
    int64_t off, len;
    span(c, off, len);
    cudaError_t e2 = cudaMemcpyAsync(l.stage[b], d_src + off, sizeof(float) * len, cudaMemcpyDeviceToHost, l.stream);
    if (e2 != cudaSuccess) return e2;
    return cudaEventRecord(l.drained[b], l.stream);   // here: the buffer has been FILLED
  };
  int used = 0;
  if (lane < nchunks) {
    e = fetch(lane, 0);
    if (e != cudaSuccess) return e;
  }
  for (int64_t c = lane; c < nchunks; c += nlanes, ++used) {
    const int b = used % kStageBufs;
    if (c + nlanes < nchunks) {
      e = fetch(c + nlanes, (used + 1) % kStageBufs);   // that buffer was consumed in the previous round
      if (e != cudaSuccess) return e;
    }
    e = cudaEventSynchronize(l.drained[b]);
    if (e != cudaSuccess) return e;
    int64_t off, len;
    span(c, off, len);
    widen(l.stage[b], h_dst + off, len);
  }
  return cudaSuccess;
}

template <class T>
int download_t(const float* d_src, int64_t count, T* h_dst, int threads, cudaStream_t st) {
  int dev = 0;
  AISB_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    set_error("download: device index %d out of range", dev);
    return AISB_ERR_INVALID;
  }
  int nlanes = threads < 1 ? 1 : (threads > kUploadMaxThreads ? kUploadMaxThreads : threads);
  const int64_t chunk = chunk_elems(count, nlanes);
  const int64_t nchunks = (count + chunk - 1) / chunk;
  if (nlanes > nchunks) nlanes = static_cast<int>(nchunks);
  std::lock_guard<std::mutex> lock(g_upload_mu);
  Pool& pool = g_pools[dev];
  if (!pool.start) AISB_CUDA_CHECK(cudaEventCreateWithFlags(&pool.start, cudaEventDisableTiming));
  for (int t = 0; t < nlanes; ++t) AISB_CUDA_CHECK(lane_init(pool.lanes[t]));
  // the copies queue behind whatever produces d_src on the caller's stream
  AISB_CUDA_CHECK(cudaEventRecord(pool.start, st));
  for (int t = 0; t < nlanes; ++t) AISB_CUDA_CHECK(cudaStreamWaitEvent(pool.lanes[t].stream, pool.start, 0));
  cudaError_t errs[kUploadMaxThreads];
  for (int t = 0; t < nlanes; ++t) errs[t] = cudaSuccess;
  run_lanes(nlanes, errs, [&](int t) { return lane_fetch(pool.lanes[t], dev, t, nlanes, d_src, count, chunk, h_dst); });
  for (int t = 0; t < nlanes; ++t)
    if (errs[t] != cudaSuccess) {
      for (int u = 0; u < nlanes; ++u) cudaStreamSynchronize(pool.lanes[u].stream);   // leave no copy in flight
      return cuda_fail(errs[t], "download lane");
    }
  return AISB_OK;
}

}  // namespace
}  // namespace aisb

using namespace aisb;

extern "C" int aisb_download_f32(const float* d_src, int64_t count, void* h_dst, int32_t dst_is_f64, int32_t threads,
                                 void* stream) {
  if (count == 0) return AISB_OK;
  if (!h_dst || !d_src || count < 0) {
    set_error("download_f32: null pointer or negative count");
    return AISB_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dst_is_f64) return download_t(d_src, count, static_cast<double*>(h_dst), threads, st);
  return download_t(d_src, count, static_cast<float*>(h_dst), threads, st);
}

extern "C" int aisb_upload_as_f32(const void* h_src, int32_t src_is_f64, int64_t count, float* d_dst, int32_t threads,
                                  void* stream) {
  if (count == 0) return AISB_OK;
  if (!h_src || !d_dst || count < 0) {
    set_error("upload_as_f32: null pointer or negative count");
    return AISB_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_is_f64) return upload_t(static_cast<const double*>(h_src), count, d_dst, threads, st);
  return upload_t(static_cast<const float*>(h_src), count, d_dst, threads, st);
}
