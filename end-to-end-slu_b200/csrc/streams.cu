// Side-stream fork/join for independent launches of one layer's backward (weight-gradient GEMMs next to the
// input-gradient GEMM).  Host-side only: a per-device pool of non-blocking streams and timing-free events, so the
// Python host pays two C calls per fork/join instead of a dozen torch stream operations.
// Memory contract (same as for every kernel of this library): the caller keeps all buffers alive until work queued on
// `main` after slu_stream_join has been reached; the side streams never outlive a fork/join pair.
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kMaxSide = 8;
constexpr int kMaxDev = 16;

struct Pool {
  bool ready = false;
  cudaStream_t side[kMaxSide];
  cudaEvent_t forked;
  cudaEvent_t done[kMaxSide];
  cudaStream_t copy;               // host -> device staging of the next batch
  cudaEvent_t copy_after, copy_done;
};

Pool g_pool[kMaxDev];
std::mutex g_mu;

int get_pool(Pool** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= kMaxDev) return (int)cudaErrorInvalidDevice;
  Pool& p = g_pool[dev];
  if (!p.ready) {
    for (int i = 0; i < kMaxSide; ++i) {
      if ((e = cudaStreamCreateWithFlags(&p.side[i], cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
      if ((e = cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
    }
    if ((e = cudaEventCreateWithFlags(&p.forked, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
    if ((e = cudaStreamCreateWithFlags(&p.copy, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
    if ((e = cudaEventCreateWithFlags(&p.copy_after, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
    if ((e = cudaEventCreateWithFlags(&p.copy_done, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
    p.ready = true;
  }
  *out = &p;
  return 0;
}

}  // namespace

extern "C" int slu_stream_fork(void* main_stream, int n, void** side_streams) {
  if (n < 0 || n > kMaxSide) return (int)cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lock(g_mu);
  Pool* p = nullptr;
  if (int err = get_pool(&p)) return err;
  cudaError_t e = cudaEventRecord(p->forked, (cudaStream_t)main_stream);
  if (e != cudaSuccess) return (int)e;
  for (int i = 0; i < n; ++i) {
    if ((e = cudaStreamWaitEvent(p->side[i], p->forked, 0)) != cudaSuccess) return (int)e;
    side_streams[i] = (void*)p->side[i];
  }
  return 0;
}

extern "C" int slu_stream_join(void* main_stream, int n) {
  if (n < 0 || n > kMaxSide) return (int)cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lock(g_mu);
  Pool* p = nullptr;
  if (int err = get_pool(&p)) return err;
  for (int i = 0; i < n; ++i) {
    cudaError_t e = cudaEventRecord(p->done[i], p->side[i]);
    if (e != cudaSuccess) return (int)e;
    if ((e = cudaStreamWaitEvent((cudaStream_t)main_stream, p->done[i], 0)) != cudaSuccess) return (int)e;
  }
  return 0;
}

// Host -> device staging on the library's copy stream.  `after_stream` (may be NULL = no ordering): the copy starts only after
// everything queued there so far -- pass the consumer stream when `dst` is a recycled buffer whose previous contents may still be
// in use.  `src` should be pinned (otherwise the runtime stages it synchronously).
extern "C" int slu_h2d_async(void* dst, const void* src, size_t bytes, void* after_stream, int order_after) {
  std::lock_guard<std::mutex> lock(g_mu);
  Pool* p = nullptr;
  if (int err = get_pool(&p)) return err;
  cudaError_t e;
  if (order_after) {
    if ((e = cudaEventRecord(p->copy_after, (cudaStream_t)after_stream)) != cudaSuccess) return (int)e;
    if ((e = cudaStreamWaitEvent(p->copy, p->copy_after, 0)) != cudaSuccess) return (int)e;
  }
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, p->copy);
}

// Work queued on `consumer_stream` from now on waits for every copy queued with slu_h2d_async so far.
extern "C" int slu_h2d_ready(void* consumer_stream) {
  std::lock_guard<std::mutex> lock(g_mu);
  Pool* p = nullptr;
  if (int err = get_pool(&p)) return err;
  cudaError_t e = cudaEventRecord(p->copy_done, p->copy);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamWaitEvent((cudaStream_t)consumer_stream, p->copy_done, 0);
}

// *pending = 1 while copies queued with slu_h2d_async are still in flight on the copy stream (their pinned sources must stay
// alive until then), 0 once all have completed.  Never blocks.
extern "C" int slu_h2d_pending(int* pending) {
  std::lock_guard<std::mutex> lock(g_mu);
  Pool* p = nullptr;
  if (int err = get_pool(&p)) return err;
  const cudaError_t e = cudaStreamQuery(p->copy);
  if (e == cudaErrorNotReady) {
    (void)cudaGetLastError();
    *pending = 1;
    return 0;
  }
  *pending = 0;
  return (int)e;
}

// Block the calling host thread until every copy queued with slu_h2d_async has completed (end of an epoch / teardown).
extern "C" int slu_h2d_wait(void) {
  cudaStream_t s;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    Pool* p = nullptr;
    if (int err = get_pool(&p)) return err;
    s = p->copy;
  }
  return (int)cudaStreamSynchronize(s);
}
