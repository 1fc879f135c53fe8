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
