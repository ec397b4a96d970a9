// Block-level reductions shared by the weight and KLD kernels.
//
// Weight partial record (4 doubles): { M, S1 = sum exp(lw - M), S2 = sum exp(2 (lw - M)), n_valid }
// over valid entries (not NaN, not +inf). Replaces the host reductions of
// sampling_methods/dm_ais.py:86 (weights / weights.sum()) and base.py:72-75 (ESS).
#pragma once
#include "common.cuh"

namespace aisb {

#ifdef __CUDACC__

// red: NT/32*4 floats of shared memory. Thread-local values lw[e] with validity ok[e].
template <int E, int NT>
__device__ __forceinline__ void block_weight_partials(const float (&lw)[E], const bool (&ok)[E], float* red,
                                                      double* __restrict__ out4) {
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float lm = kNegBig;
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (ok[e]) lm = fmaxf(lm, lw[e]);
  lm = warp_max(lm);
  if (lane == 0) red[warp] = lm;
  __syncthreads();
  float Mb = red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) Mb = fmaxf(Mb, red[w]);
  __syncthreads();
  float s1 = 0.f, s2 = 0.f, cnt = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (ok[e]) {
      const float u = ex2((lw[e] - Mb) * 1.4426950408889634f);
      s1 += u;
      s2 = fmaf(u, u, s2);
      cnt += 1.f;
    }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  cnt = warp_sum(cnt);
  if (lane == 0) {
    red[warp] = s1;
    red[NW + warp] = s2;
    red[2 * NW + warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      a += red[w];
      b += red[NW + w];
      c += red[2 * NW + w];
    }
    out4[0] = Mb;
    out4[1] = a;
    out4[2] = b;
    out4[3] = c;
  }
}

#endif

// one block of 256 threads folds `nblocks` weight records into one; a non-zero *abort (nullable: the abort counter of a
// tensor-core kernel) poisons the result (n_valid = -1)
__global__ void weight_partials_merge_kernel(const double* __restrict__ block_partials, int64_t nblocks,
                                             double* __restrict__ out4, const int32_t* __restrict__ abort);

}  // namespace aisb
