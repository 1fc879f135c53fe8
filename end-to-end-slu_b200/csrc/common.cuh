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

#define SLU_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

static inline int slu_set_smem(const void* fn, size_t bytes) {
  return (int)cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
