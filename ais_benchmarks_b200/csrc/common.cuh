// Shared device/host helpers for libais_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ais_b200.h"

namespace aisb {

constexpr int kThreads = 128;            // threads per CTA of the pairwise kernels
constexpr float kNegBig = -3.0e38f;      // finite stand-in for -inf as the running max (keeps m - m_new NaN-free)
constexpr double kLog2e = 1.4426950408889634074;
constexpr double kLn2 = 0.69314718055994530942;
constexpr double kLog2Pi = 1.8378770664093454836;  // log(2*pi)

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define AISB_CUDA_CHECK(expr)                                  \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::aisb::cuda_fail(_e, #expr); \
  } while (0)

#define AISB_LAUNCH_CHECK(what)                                \
  do {                                                         \
    cudaError_t _e = cudaGetLastError();                       \
    if (_e != cudaSuccess) return ::aisb::cuda_fail(_e, what); \
  } while (0)

int sm_count();  // cached multiprocessor count of the current device

// tensor-core log q launcher (tc.cu); *d_err is incremented by CTAs that aborted on a barrier time-out
int tc_logq_launch(const float* d_x, int64_t n, const aisb_packed_mixture* mix, const int32_t* d_hint,
                   const float* d_image, float* d_out, int32_t* d_err, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// MUFU.EX2 (2^x), flush-to-zero; max rel. error 2^-22.
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// order-preserving unsigned encoding of a float (for atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ uint32_t enc_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// Packed layout traits (must match include/ais_b200.h)
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int row_floats(int kind, int DP) {
  return kind == AISB_COV_ISO_SHARED ? DP : (kind == AISB_COV_DIAG ? 2 * DP : ((DP * (DP + 1) / 2 + DP + 3) / 4) * 4);
}

}  // namespace aisb
