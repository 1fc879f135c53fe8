// Minimal hand-written tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX).
// Operand convention used by every tensor-core kernel in this library:
//   * A and B tiles are K-major bf16 in shared memory, NO swizzle ("interleaved" canonical layout):
//       core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes;
//       byte(r, k) = (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2
//     We use the "k-chunk-major" tile: LBO = rows*16, SBO = 128, i.e. all rows of one 8-element
//     k-chunk are contiguous.  Threads stage operands with 16-byte st.shared (conflict-free).
//   * fp32 data is split on the fly into bf16 hi + bf16 lo; every product is issued as
//     hi*hi + hi*lo + lo*hi (3 MMAs, fp32 accumulate in TMEM)  -> ~2^-17 relative error,
//     which the 1e-3 logit tolerance needs (see oracle/precision_study.py).
//   * D (accumulator) lives in TMEM: lane = M row (0..127), column = N index (fp32).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- warp-uniform helpers ------------------------------------------------------------------------
// Warp index as a value the compiler knows to be warp-uniform (so role branches are not "divergent").
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
// One lane of a fully converged warp (elect.sync); tcgen05.mma / commit are issued under this predicate.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}" : "+r"(pred));
  return pred != 0;
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a protocol bug traps (launch failure) instead of hanging the GPU
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}

// ---- TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP) ------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMA tiled tensor copy (3-D tensor map) global -> shared: one instruction moves a [box2][box1][box0] box; coordinates outside
// the tensor read zeros (SASS: UTMALDG).  `tmap` = address of a CUtensorMap in kernel-parameter (__grid_constant__) space;
// dst 128-byte aligned; the mbarrier receives the full box bytes.
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst_smem),
               "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---- Ampere-style 16-byte async copy global -> shared (SASS: LDGSTS), zero-filling beyond src_bytes, and its
// completion hook: the mbarrier receives one of its expected arrivals once all prior cp.async of this thread have landed.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(src_bytes) : "memory");
}
// same with the destination given as a 32-bit shared-memory address (hoisted out of inner loops)
__device__ __forceinline__ void cp_async16_s(uint32_t dst_smem, const void* src_gmem, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src_gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- fences ---------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp) ---------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE, version 1 (Blackwell).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor: kind::f16, A=B=bf16 (K-major), D=f32, dense.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with fp16 operands (11-bit significand), D=f32: the single-pass mode of the GRU recurrence.
__host__ __device__ constexpr uint32_t idesc_f16_f32(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA issue / commit (single thread) ------------------------------------------------------------
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A operand from TMEM (weights-stationary): lane = M row, 32-bit column c holds elements k=2c (low half), 2c+1.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- registers -> TMEM ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM -> registers (warp w may only touch lanes 32*(w%4) .. +31) -------------------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr));
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float* v) {
  static_assert(N == 1 || N == 2 || N == 4 || N == 8 || N == 16, "tmem_ld width");
  if (N == 1) tmem_ld1(taddr, v); else if (N == 2) tmem_ld2(taddr, v); else if (N == 4) tmem_ld4(taddr, v);
  else if (N == 8) tmem_ld8(taddr, v); else tmem_ld16(taddr, v);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- fp32 -> (bf16 hi, bf16 lo) split helpers --------------------------------------------------------
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  // packed converts (one F2FP per pair): hi = {bf16(b), bf16(a)}, lo = {bf16(b - hi_b), bf16(a - hi_a)}
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}
// 8 consecutive-k fp32 values -> one 16-byte hi chunk and one 16-byte lo chunk
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  split2(v[0], v[1], hi.x, lo.x); split2(v[2], v[3], hi.y, lo.y);
  split2(v[4], v[5], hi.z, lo.z); split2(v[6], v[7], hi.w, lo.w);
}
// byte offset of the 16-byte chunk (row r, k-chunk kc) inside a k-chunk-major tile of `rows` rows
__device__ __forceinline__ uint32_t chunk_off(int rows, int r, int kc) { return (uint32_t)(kc * rows + r) * 16u; }

// D[tmem] (+)= A(hi,lo) . B(hi,lo)^T over `k16` K-steps of 16, 3-pass bf16 split.  Single thread.
// a_hi/a_lo/b_hi/b_lo: shared-memory byte addresses of k-chunk-major tiles with a_rows / b_rows rows.
__device__ __forceinline__ void mma_split3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, int a_rows, uint32_t b_hi, uint32_t b_lo,
                                           int b_rows, int k16, uint32_t idesc, bool accumulate_first) {
  const uint32_t a_lbo = a_rows * 16, b_lbo = b_rows * 16;
  uint32_t acc = accumulate_first ? 1u : 0u;
  for (int kk = 0; kk < k16; ++kk) {
    const uint64_t ah = smem_desc(a_hi + kk * 2 * a_lbo, a_lbo, 128), al = smem_desc(a_lo + kk * 2 * a_lbo, a_lbo, 128);
    const uint64_t bh = smem_desc(b_hi + kk * 2 * b_lbo, b_lbo, 128), bl = smem_desc(b_lo + kk * 2 * b_lbo, b_lbo, 128);
    mma_bf16(d_tmem, ah, bh, idesc, acc); acc = 1u;
    mma_bf16(d_tmem, ah, bl, idesc, 1u);
    mma_bf16(d_tmem, al, bh, idesc, 1u);
  }
}

// Same with the A operand resident in TMEM: a_hi_t / a_lo_t are TMEM addresses of [128 x 16*k16] bf16 (packed pairs,
// 8 columns per K-step).  B tile has byte strides b_lbo (between 8-element k-chunks) / 128 (between 8-row groups).
__device__ __forceinline__ void mma_split3_ts(uint32_t d_tmem, uint32_t a_hi_t, uint32_t a_lo_t, uint32_t b_hi, uint32_t b_lo,
                                              uint32_t b_lbo, int k16, uint32_t idesc, bool accumulate_first) {
  uint32_t acc = accumulate_first ? 1u : 0u;
  for (int kk = 0; kk < k16; ++kk) {
    const uint64_t bh = smem_desc(b_hi + kk * 2 * b_lbo, b_lbo, 128), bl = smem_desc(b_lo + kk * 2 * b_lbo, b_lbo, 128);
    mma_bf16_ts(d_tmem, a_hi_t + kk * 8, bh, idesc, acc); acc = 1u;
    mma_bf16_ts(d_tmem, a_hi_t + kk * 8, bl, idesc, 1u);
    mma_bf16_ts(d_tmem, a_lo_t + kk * 8, bh, idesc, 1u);
  }
}

// Store one M-row (this thread's TMEM lane) of fp32 values as packed bf16 hi / lo A-operands: `n` values (multiple of 16).
__device__ __forceinline__ void tmem_store_row_split(uint32_t t_hi, uint32_t t_lo, const float* row, int n) {
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split2(row[c0 + 2 * i], row[c0 + 2 * i + 1], hi[i], lo[i]);
    tmem_st8(t_hi + c0 / 2, hi);
    tmem_st8(t_lo + c0 / 2, lo);
  }
  tmem_st_wait();
}

// fp16 single-pass variant of tmem_store_row_split: one packed fp16 A-operand.
__device__ __forceinline__ void tmem_store_row_f16(uint32_t t_a, const float* row, int n) {
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __half2 h = __floats2half2_rn(row[c0 + 2 * i], row[c0 + 2 * i + 1]);
      v[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st8(t_a + c0 / 2, v);
  }
  tmem_st_wait();
}

// Descriptor arithmetic for unrolled issue loops: advance the start-address field (16-byte units) of a descriptor.
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

}  // namespace tc05
