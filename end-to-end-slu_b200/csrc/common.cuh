// Shared helpers for the slu_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SLU_H 128            // GRU hidden size (every reference cfg: *_rnn_num_hidden=128)
#define SLU_G3 384           // 3 gates * H, PyTorch order (r, z, n)
#define SLU_NFILT 80         // SincLayer filters   (cnn_N_filt[0])
#define SLU_NTAPS 401        // SincLayer taps      (cnn_len_filt[0])
#define SLU_STRIDE 80        // SincLayer stride    (cnn_stride[0])
#define SLU_PAD 200          // cnn_len_filt[0] // 2

#define SLU_ERR_TOO_LARGE 100001   // same value as in include/slu_b200.h: a size exceeds a kernel's 32-bit index range

#define SLU_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

static inline int slu_set_smem(const void* fn, size_t bytes) {
  return (int)cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// The attribute is PER DEVICE: a process that drives several devices must set it on each.  `SLU_SMEM_ONCE(fn, bytes)` sets it the
// first time the calling site launches on each device (one static mask per site) and returns the error code from the enclosing
// function if it fails.
struct SluSmemOnce { unsigned long long done = 0ull; };
static inline int slu_smem_once(SluSmemOnce& st, const void* fn, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (st.done & bit) return 0;
  const int r = slu_set_smem(fn, bytes);
  if (r == 0) st.done |= bit;
  return r;
}
#define SLU_SMEM_ONCE(fn, bytes) do { static SluSmemOnce once__; if (int e__ = slu_smem_once(once__, (const void*)(fn), (bytes))) return e__; } while (0)
