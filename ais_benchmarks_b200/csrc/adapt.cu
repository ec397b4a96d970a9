// Adaptation-step kernels: M-PMC Rao-Blackwellised moments, raw moments, LAIS MH chains.
#include <stdlib.h>

#include "pairwise.cuh"

namespace aisb {

// log2 of the weight normaliser sum_i w_i: a host value, or derived in the kernel from a weight record {M, S1, ..} that
// lives in device memory (M + log S1), so that an adapt iteration needs no host round trip between the weight kernel
// and the moment kernel
struct LogNorm {
  float l2n;
  const double* rec;
  __device__ __forceinline__ float get() const {
    return rec ? static_cast<float>((rec[0] + log(rec[1])) * 1.4426950408889634) : l2n;
  }
};

// ---------------------------------------------------------------------------------------------
// M-PMC moments (reference sampling_methods/m_pmc.py:57-61,102-105)
//   a_{d,i} = w~_i rho_{d,i} = 2^( t_d(x_i) - log2 q(x_i) + log2 w~_i ),  t_d = c_d - ||A_d x_i + b_d||^2
//   out[d][0] += sum_i a_{d,i},  out[d][1+j] += sum_i a_{d,i} x_ij
// Points live in registers (same tiling as the density kernels); the sum over points is a warp
// shuffle reduction followed by one shared-memory double atomic per (component, column) per warp,
// flushed to global memory with double atomics once per chunk.
// ---------------------------------------------------------------------------------------------
template <int KIND, int DP>
__global__ void __launch_bounds__(kThreads)
    mpmc_moments_kernel(const float* __restrict__ x, int64_t n, int D, int vec_ok, const float* __restrict__ rows,
                        const float* __restrict__ offs, int64_t KP, int64_t K, float a_iso,
                        const float* __restrict__ logq, const float* __restrict__ logw, LogNorm log2_norm,
                        double* __restrict__ out) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[kStages];
  constexpr int E = points_per_thread(DP);
  constexpr int J = chunk_comps(KIND, DP);
  constexpr int ROW = row_floats(KIND, DP);
  // accumulators for one chunk live after the staging ring
  double* sacc = reinterpret_cast<double*>(smem + kStages * stage_floats(KIND, DP, true));

  ring_init(bars);
  float xr[E][DP];
  int64_t pidx[E];
  load_points<DP, E>(x, n, D, vec_ok, static_cast<int64_t>(blockIdx.x) * (kThreads * E), xr, pidx);
  float shift[E];  // log2 w~_i - log2 q(x_i); -inf for out-of-range or zero-weight points
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (pidx[e] < n) {
      const float lw = __ldg(logw + pidx[e]), lq = __ldg(logq + pidx[e]);
      shift[e] = (lw - lq) * 1.4426950408889634f - log2_norm.get();
      if (!(shift[e] == shift[e]) || shift[e] == INFINITY) shift[e] = -INFINITY;
    } else {
      shift[e] = -INFINITY;
    }
  }
  for (int i = threadIdx.x; i < J * (DP + 1); i += kThreads) sacc[i] = 0.0;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  uint32_t it = 0;
  stream_components<KIND, DP, true>(
      rows, offs, 0, KP, smem, bars, it, [&](const float* srows, const float* soffs, int cnt, int64_t k0) {
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
          if (k0 + j >= K) break;  // padding components (uniform across the CTA)
          float t[E];
          comp_eval<KIND, DP, E>(srows + j * ROW, soffs[j], a_iso, xr, t);
          float v[DP + 1];
#pragma unroll
          for (int d = 0; d <= DP; ++d) v[d] = 0.f;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const float a = ex2(t[e] + shift[e]);
            v[0] += a;
#pragma unroll
            for (int d = 0; d < DP; ++d) v[1 + d] = fmaf(a, xr[e][d], v[1 + d]);
          }
#pragma unroll
          for (int d = 0; d <= DP; ++d) v[d] = warp_sum(v[d]);
          if (lane == 0) {
#pragma unroll
            for (int d = 0; d <= DP; ++d) atomicAdd(&sacc[j * (DP + 1) + d], static_cast<double>(v[d]));
          }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * (DP + 1); i += kThreads) {
          const int j = i / (DP + 1), d = i % (DP + 1);
          if (k0 + j < K && d <= D) {
            const double val = sacc[i];
            if (val != 0.0) atomicAdd(&out[(k0 + j) * (D + 1) + d], val);
          }
          sacc[i] = 0.0;
        }
        // the __syncthreads() at the end of the stage protects sacc for the next chunk
      });
}

// ---------------------------------------------------------------------------------------------
// Component-stationary variant for small rows (ISO / DIAG up to DP = 16, FULL up to DP = 8): thread = component, its
// parameters and its D+1 accumulators live in registers, the points of a tile stream through shared memory (every
// lane reads the same point: LDS broadcast). The sum over points is then thread-local -- no warp shuffles, no shared
// atomics; a CTA adds its K x (D+1) float64 partials to global memory once. Measured at config 3 (1e6 points x 256
// proposals, D = 8) against the point-stationary kernel above, whose per-component warp reduction of D+1 values
// dominates its time: see DESIGN.md section 4.4.
// ---------------------------------------------------------------------------------------------
constexpr int kCsThreads = 128;      // components per CTA
constexpr int kCsSub = 128;          // points per staged sub-tile (one per thread)
constexpr int kCsSubTiles = 16;      // sub-tiles per CTA: 2048 points

// SECOND: also accumulate the weighted second moments sum_i a_{d,i} x_i x_i^T (lower triangle, row-major over the real
// D: index r (r + 1) / 2 + c behind the D + 1 first-order columns) -- the sufficient statistic of the covariance update
// of M-PMC as published (Cappe et al. 2008, eq. 14.3), which the reference replaces by the unweighted scatter (quirk Q3).
template <int KIND, int DP, bool SECOND>
__global__ void __launch_bounds__(kCsThreads)
    mpmc_moments_cs_kernel(const float* __restrict__ x, int64_t n, int D, int vec_ok, const float* __restrict__ rows,
                           const float* __restrict__ offs, int64_t K, float a_iso, const float* __restrict__ logq,
                           const float* __restrict__ logw, LogNorm log2_norm, double* __restrict__ out) {
  constexpr int ROW = row_floats(KIND, DP);
  constexpr int TRI2 = SECOND ? DP * (DP + 1) / 2 : 1;
  constexpr int XS = DP >= 4 ? DP : 4;                      // padded point row in shared memory (float4 reads)
  __shared__ __align__(16) float sx[kCsSub * XS];
  __shared__ float ssh[kCsSub];
  const int64_t k = static_cast<int64_t>(blockIdx.y) * kCsThreads + threadIdx.x;
  const bool live = k < K;
  float row[ROW];
  float c = -INFINITY;
  if (live) {
#pragma unroll
    for (int q = 0; q < ROW; ++q) row[q] = __ldg(rows + k * ROW + q);
    c = __ldg(offs + k);
  } else {
#pragma unroll
    for (int q = 0; q < ROW; ++q) row[q] = 0.f;
  }
  double acc[DP + 1];
#pragma unroll
  for (int d = 0; d <= DP; ++d) acc[d] = 0.0;
  double acc2[TRI2];
#pragma unroll
  for (int q = 0; q < TRI2; ++q) acc2[q] = 0.0;

  const int64_t tile0 = static_cast<int64_t>(blockIdx.x) * (kCsSub * kCsSubTiles);
#pragma unroll 1
  for (int st = 0; st < kCsSubTiles; ++st) {
    const int64_t base = tile0 + static_cast<int64_t>(st) * kCsSub;
    if (base >= n) break;  // uniform across the CTA
    // ---- stage: thread t loads point base + t and its shift log2 w~_i - log2 q(x_i) ----
    {
      const int64_t p = base + threadIdx.x;
      float xv[XS];
#pragma unroll
      for (int d = 0; d < XS; ++d) xv[d] = 0.f;
      float sh = -INFINITY;
      if (p < n) {
        if (DP >= 4 && D == DP && vec_ok) {
#pragma unroll
          for (int q = 0; q < DP / 4; ++q) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(x + p * DP) + q);
            xv[4 * q] = f.x;
            xv[4 * q + 1] = f.y;
            xv[4 * q + 2] = f.z;
            xv[4 * q + 3] = f.w;
          }
        } else {
#pragma unroll
          for (int d = 0; d < DP; ++d)
            if (d < D) xv[d] = __ldg(x + p * D + d);
        }
        const float lw = __ldg(logw + p), lq = __ldg(logq + p);
        sh = (lw - lq) * 1.4426950408889634f - log2_norm.get();
        if (!(sh == sh) || sh == INFINITY) sh = -INFINITY;
      }
#pragma unroll
      for (int q = 0; q < XS / 4; ++q)
        reinterpret_cast<float4*>(sx + threadIdx.x * XS)[q] = make_float4(xv[4 * q], xv[4 * q + 1], xv[4 * q + 2], xv[4 * q + 3]);
      ssh[threadIdx.x] = sh;
    }
    __syncthreads();
    // ---- this thread's component against the staged points ----
    const int cnt = n - base < kCsSub ? static_cast<int>(n - base) : kCsSub;
    float v[DP + 1];
#pragma unroll
    for (int d = 0; d <= DP; ++d) v[d] = 0.f;
    float v2[TRI2];
#pragma unroll
    for (int q = 0; q < TRI2; ++q) v2[q] = 0.f;
#pragma unroll 2
    for (int i = 0; i < cnt; ++i) {
      float xi[XS];
#pragma unroll
      for (int q = 0; q < XS / 4; ++q) {
        const float4 f = reinterpret_cast<const float4*>(sx + i * XS)[q];
        xi[4 * q] = f.x;
        xi[4 * q + 1] = f.y;
        xi[4 * q + 2] = f.z;
        xi[4 * q + 3] = f.w;
      }
      float t = c;
      if constexpr (KIND == AISB_COV_ISO_SHARED) {
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const float dl = fmaf(xi[d], a_iso, row[d]);
          t = fmaf(-dl, dl, t);
        }
      } else if constexpr (KIND == AISB_COV_DIAG) {
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const float dl = fmaf(xi[d], row[d], row[DP + d]);
          t = fmaf(-dl, dl, t);
        }
      } else {
        constexpr int TRI = DP * (DP + 1) / 2;
        int idx = 0;
#pragma unroll
        for (int r = 0; r < DP; ++r) {
          float z = row[TRI + r];
#pragma unroll
          for (int cc = 0; cc <= r; ++cc) z = fmaf(row[idx++], xi[cc], z);
          t = fmaf(-z, z, t);
        }
      }
      const float a = ex2(t + ssh[i]);
      v[0] += a;
#pragma unroll
      for (int d = 0; d < DP; ++d) v[1 + d] = fmaf(a, xi[d], v[1 + d]);
      if constexpr (SECOND) {
        int q = 0;
#pragma unroll
        for (int r = 0; r < DP; ++r) {
          const float ar = a * xi[r];
#pragma unroll
          for (int cc = 0; cc <= r; ++cc) {
            v2[q] = fmaf(ar, xi[cc], v2[q]);
            ++q;
          }
        }
      }
    }
#pragma unroll
    for (int d = 0; d <= DP; ++d) acc[d] += static_cast<double>(v[d]);
    if constexpr (SECOND) {
#pragma unroll
      for (int q = 0; q < TRI2; ++q) acc2[q] += static_cast<double>(v2[q]);
    }
    __syncthreads();  // the stage is overwritten by the next sub-tile
  }
  if (live) {
    const int W = SECOND ? (D + 1) + D * (D + 1) / 2 : D + 1;
#pragma unroll
    for (int d = 0; d <= DP; ++d)
      if (d <= D && acc[d] != 0.0) atomicAdd(&out[k * W + d], acc[d]);
    if constexpr (SECOND) {
      int q = 0;
#pragma unroll
      for (int r = 0; r < DP; ++r) {
#pragma unroll
        for (int cc = 0; cc <= r; ++cc) {
          if (r < D && acc2[q] != 0.0) atomicAdd(&out[k * W + (D + 1) + r * (r + 1) / 2 + cc], acc2[q]);
          ++q;
        }
      }
    }
  }
}

template <int KIND, int DP, bool SECOND = false>
static int launch_mpmc_cs_t(const float* x, int64_t n, int D, int vec_ok, const aisb_packed_mixture* mix,
                            const float* logq, const float* logw, LogNorm log2_norm, double* out, cudaStream_t st) {
  const int64_t tiles = (n + kCsSub * kCsSubTiles - 1) / (kCsSub * kCsSubTiles);
  const dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>((mix->K + kCsThreads - 1) / kCsThreads));
  mpmc_moments_cs_kernel<KIND, DP, SECOND><<<grid, kCsThreads, 0, st>>>(
      x, n, D, vec_ok, mix->d_rows, mix->d_offsets, mix->K, mix->iso_scale, logq, logw, log2_norm, out);
  AISB_LAUNCH_CHECK("mpmc_moments_cs_kernel");
  return AISB_OK;
}

// second-order variant: D <= 8 (1 + D + D (D + 1) / 2 <= 45 accumulators per thread)
template <int KIND>
static int launch_mpmc2_dp(int DP, const float* x, int64_t n, int D, int vec_ok, const aisb_packed_mixture* mix,
                           const float* logq, const float* logw, LogNorm log2_norm, double* out, cudaStream_t st) {
  switch (DP) {
    case 1: return launch_mpmc_cs_t<KIND, 1, true>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    case 2: return launch_mpmc_cs_t<KIND, 2, true>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    case 4: return launch_mpmc_cs_t<KIND, 4, true>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    case 8: return launch_mpmc_cs_t<KIND, 8, true>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    default: break;
  }
  set_error("mpmc_moments2: second moments are limited to D <= 8 (got %d)", D);
  return AISB_ERR_UNSUPPORTED;
}

__host__ __device__ constexpr bool mpmc_cs_ok(int kind, int DP) {
  return kind == AISB_COV_FULL ? DP <= 8 : DP <= 16;
}

template <int KIND, int DP>
static int launch_mpmc_t(const float* x, int64_t n, int D, int vec_ok, const aisb_packed_mixture* mix,
                         const float* logq, const float* logw, LogNorm log2_norm, double* out, cudaStream_t st) {
  if constexpr (mpmc_cs_ok(KIND, DP)) {
    static const bool ps = [] {  // AISB_MPMC_POINT_STATIONARY=1 keeps the point-stationary kernel (timing experiments)
      const char* e = getenv("AISB_MPMC_POINT_STATIONARY");
      return e && e[0] == '1';
    }();
    if (!ps) return launch_mpmc_cs_t<KIND, DP>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
  }
  constexpr int E = points_per_thread(DP);
  constexpr size_t smem = pass_smem_bytes(KIND, DP, true) + sizeof(double) * chunk_comps(KIND, DP) * (DP + 1);
  auto kern = mpmc_moments_kernel<KIND, DP>;
  if (smem > 48 * 1024) AISB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t ntiles = (n + kThreads * E - 1) / (kThreads * E);
  kern<<<static_cast<unsigned>(ntiles), kThreads, smem, st>>>(x, n, D, vec_ok, mix->d_rows, mix->d_offsets,
                                                             aisb_padded_components(mix->K), mix->K, mix->iso_scale,
                                                             logq, logw, log2_norm, out);
  AISB_LAUNCH_CHECK("mpmc_moments_kernel");
  return AISB_OK;
}

template <int KIND>
static int launch_mpmc_dp(int DP, const float* x, int64_t n, int D, int vec_ok, const aisb_packed_mixture* mix,
                          const float* logq, const float* logw, LogNorm log2_norm, double* out, cudaStream_t st) {
  switch (DP) {
    case 1: return launch_mpmc_t<KIND, 1>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    case 2: return launch_mpmc_t<KIND, 2>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    case 4: return launch_mpmc_t<KIND, 4>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    case 8: return launch_mpmc_t<KIND, 8>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    case 16: return launch_mpmc_t<KIND, 16>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    case 32: return launch_mpmc_t<KIND, 32>(x, n, D, vec_ok, mix, logq, logw, log2_norm, out, st);
    default: break;
  }
  set_error("mpmc_moments: unsupported D=%d (max 32)", D);
  return AISB_ERR_UNSUPPORTED;
}


// ---------------------------------------------------------------------------------------------
// M-PMC parameter update on the device: alpha / mu / Sigma of m_pmc.py:53-78 from the sufficient statistics the moment
// kernels left in HBM, the 10*I failsafe (:71-72, quirk Q3: the covariance is the UNWEIGHTED scatter about mu_d),
// set_moments' det > 0 check (CMultivariateNormal.py:55-56), set_weights' sanitising (CMixtureModel.py:25-41), and the
// re-packing of the proposal mixture for the next weighting (both the DIAG and the FULL layout of include/ais_b200.h:
// the caller picks DIAG when status says every covariance is diagonal -- the usual case, the failsafe) plus the float32
// means / Cholesky factors the ancestral sampler wants. One thread per proposal, everything in float64 in local arrays
// (D <= 16); N <= a few thousand proposals, so one CTA. Replaces three device->host reads, the host linear algebra and
// the host packing + upload of an iteration.
//   status[0] proposals on the failsafe, [1] proposals with a non-diagonal covariance, [2] proposals whose covariance
//   has det <= 0 or is not positive definite (the reference would raise AssertionError), [3] 1 if every mixture weight
//   was sanitised to 0 (the reference then draws every sample from the last component).
// ---------------------------------------------------------------------------------------------
constexpr int kUpdMaxD = 16;
constexpr int kUpdThreads = 256;

__device__ double upd_block_sum(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = kUpdThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// Gauss-Jordan inverse with partial pivoting and the determinant (host.cu, invert_with_det); false if singular
__device__ bool upd_invert(const double* C, int D, double* M, double* P, double* det) {
  for (int i = 0; i < D * D; ++i) M[i] = C[i];
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) P[i * D + j] = i == j ? 1.0 : 0.0;
  double d = 1.0;
  for (int c = 0; c < D; ++c) {
    int piv = c;
    for (int r = c + 1; r < D; ++r)
      if (fabs(M[r * D + c]) > fabs(M[piv * D + c])) piv = r;
    if (M[piv * D + c] == 0.0 || M[piv * D + c] != M[piv * D + c]) return false;
    if (piv != c) {
      for (int j = 0; j < D; ++j) {
        double t = M[piv * D + j]; M[piv * D + j] = M[c * D + j]; M[c * D + j] = t;
        t = P[piv * D + j]; P[piv * D + j] = P[c * D + j]; P[c * D + j] = t;
      }
      d = -d;
    }
    const double pv = M[c * D + c];
    d *= pv;
    for (int j = 0; j < D; ++j) {
      M[c * D + j] /= pv;
      P[c * D + j] /= pv;
    }
    for (int r = 0; r < D; ++r) {
      if (r == c) continue;
      const double f = M[r * D + c];
      if (f == 0.0) continue;
      for (int j = 0; j < D; ++j) {
        M[r * D + j] -= f * M[c * D + j];
        P[r * D + j] -= f * P[c * D + j];
      }
    }
  }
  *det = d;
  return true;
}

__global__ void __launch_bounds__(kUpdThreads)
    mpmc_update_kernel(const double* __restrict__ mom, const double* __restrict__ raw, double n, int64_t N, int D,
                       int DP, int64_t KP, int RF, double* __restrict__ means, double* __restrict__ covs,
                       double* __restrict__ alpha, double* __restrict__ wmix, float* __restrict__ rows_diag,
                       float* __restrict__ rows_full, float* __restrict__ offsets, float* __restrict__ means32,
                       float* __restrict__ chol32, float* __restrict__ w32, int* __restrict__ status) {
  __shared__ double sh[kUpdThreads];
  const int tid = threadIdx.x;
  const double* s1 = raw;
  const double* s2 = raw + D;
  const int W = D + 1;
  if (tid < 4) status[tid] = 0;
  // wproposals = alpha / alpha.sum() (m_pmc.py:77), then set_weights: NaN / <= 0 -> 0, / sum if the sum is positive
  double part = 0.0;
  for (int64_t k = tid; k < N; k += kUpdThreads) part += mom[k * W];
  const double sum1 = upd_block_sum(part, sh);
  part = 0.0;
  for (int64_t k = tid; k < N; k += kUpdThreads) {
    const double a = mom[k * W] / sum1;
    part += (a != a || a <= 0.0) ? 0.0 : a;
  }
  const double sum2 = upd_block_sum(part, sh);
  if (tid == 0 && !(sum2 > 0.0)) status[3] = 1;
  const float ninf = -INFINITY;
  const double base = -0.5 * D * kLog2Pi, sfac = sqrt(kLog2e / 2.0);
  const int TRI = DP * (DP + 1) / 2;
  for (int64_t k = tid; k < KP; k += kUpdThreads) {
    float* rd = rows_diag + k * 2 * DP;
    float* rf = rows_full + k * RF;
    for (int q = 0; q < 2 * DP; ++q) rd[q] = 0.f;
    for (int q = 0; q < RF; ++q) rf[q] = 0.f;
    if (k >= N) {
      offsets[k] = ninf;
      continue;
    }
    const double a_raw = mom[k * W];
    const double wprop = a_raw / sum1;
    alpha[k] = wprop;
    double w = (wprop != wprop || wprop <= 0.0) ? 0.0 : wprop;
    if (sum2 > 0.0) w /= sum2;
    wmix[k] = w;
    w32[k] = static_cast<float>(w);
    double mu[kUpdMaxD], C[kUpdMaxD * kUpdMaxD], M[kUpdMaxD * kUpdMaxD], P[kUpdMaxD * kUpdMaxD], G[kUpdMaxD * kUpdMaxD];
    for (int d = 0; d < D; ++d) mu[d] = mom[k * W + 1 + d] / a_raw;                      // eq. 14.2 (NaN / inf if alpha = 0)
    bool bad = false;
    for (int r = 0; r < D; ++r)
      for (int c = 0; c < D; ++c) {
        const double v = s2[r * D + c] - mu[r] * s1[c] - s1[r] * mu[c] + n * mu[r] * mu[c];  // raw scatter about mu (Q3)
        C[r * D + c] = v;
        bad |= (v != v) || (v > 10.0);
      }
    if (bad) {
      atomicAdd(&status[0], 1);
      for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) C[r * D + c] = r == c ? 10.0 : 0.0;
    }
    bool offdiag = false;
    for (int r = 0; r < D; ++r)
      for (int c = 0; c < D; ++c) {
        covs[(k * D + r) * D + c] = C[r * D + c];
        if (r != c && C[r * D + c] != 0.0) offdiag = true;
      }
    if (offdiag) atomicAdd(&status[1], 1);
    for (int d = 0; d < D; ++d) {
      means[k * D + d] = mu[d];
      means32[k * D + d] = static_cast<float>(mu[d]);
    }
    bool broken = false;
    // Cholesky factor of sym(C) for the sampler (x = mu + L z)
    for (int i = 0; i < D; ++i)
      for (int j = 0; j <= i; ++j) {
        double acc = 0.5 * (C[i * D + j] + C[j * D + i]);
        for (int t = 0; t < j; ++t) acc -= G[i * D + t] * G[j * D + t];
        if (i == j) {
          if (!(acc > 0.0)) broken = true;
          G[i * D + i] = sqrt(acc);
        } else {
          G[i * D + j] = acc / G[j * D + j];
        }
      }
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) chol32[(k * D + i) * D + j] = j <= i ? static_cast<float>(G[i * D + j]) : 0.f;
    // log det and the packed rows
    double logdet = 0.0, det = 0.0;
    if (!offdiag) {
      for (int d = 0; d < D; ++d) {
        const double v = C[d * D + d];
        if (!(v > 0.0)) broken = true;
        logdet += log(v);
        const float af = static_cast<float>(sqrt(kLog2e / (2.0 * v)));
        rd[d] = af;
        rd[DP + d] = static_cast<float>(-static_cast<double>(af) * mu[d]);
      }
    }
    if (!upd_invert(C, D, M, P, &det) || !(det > 0.0)) {
      broken = true;
    } else {
      if (offdiag) logdet = log(det);
      // Q = J sym(P) J = G G^T;  A = J G^T J  (host.cu, aisb_pack_mixture_f64, FULL)
      for (int i = 0; i < D; ++i)
        for (int j = 0; j <= i; ++j) {
          const int ri = D - 1 - i, rj = D - 1 - j;
          double acc = 0.5 * (P[ri * D + rj] + P[rj * D + ri]);
          for (int t = 0; t < j; ++t) acc -= G[i * D + t] * G[j * D + t];
          if (i == j) {
            if (!(acc > 0.0)) broken = true;
            G[i * D + i] = sqrt(acc);
          } else {
            G[i * D + j] = acc / G[j * D + j];
          }
        }
      for (int r = 0; r < D; ++r) {
        double br = 0.0;
        for (int c = 0; c <= r; ++c) {
          const float af = static_cast<float>(sfac * G[(D - 1 - c) * D + (D - 1 - r)]);
          rf[r * (r + 1) / 2 + c] = af;
          br -= static_cast<double>(af) * mu[c];
        }
        rf[TRI + r] = static_cast<float>(br);
      }
    }
    if (broken) atomicAdd(&status[2], 1);
    offsets[k] = w > 0.0 ? static_cast<float>(kLog2e * (log(w) + base - 0.5 * logdet)) : ninf;
  }
}

// ---------------------------------------------------------------------------------------------
// Raw first and second moments in float64: out = { sum x [D], sum x x^T [D*D] }
// ---------------------------------------------------------------------------------------------
constexpr int kRawChunk = 128;

__global__ void __launch_bounds__(256) raw_moments_kernel(const float* __restrict__ x, int64_t n, int D,
                                                          double* __restrict__ out) {
  __shared__ float tile[kRawChunk * AISB_MAX_DIMS];
  const int64_t nchunks = (n + kRawChunk - 1) / kRawChunk;
  const int nq = D + D * D;
  // each thread owns output entries q = tid, tid+256, ... and keeps their sums across its chunks
  double acc[(AISB_MAX_DIMS + AISB_MAX_DIMS * AISB_MAX_DIMS + 255) / 256];
#pragma unroll
  for (int s = 0; s < (AISB_MAX_DIMS + AISB_MAX_DIMS * AISB_MAX_DIMS + 255) / 256; ++s) acc[s] = 0.0;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t p0 = c * kRawChunk;
    const int cnt = static_cast<int>(n - p0 < kRawChunk ? n - p0 : kRawChunk);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * D; i += 256) tile[i] = x[p0 * D + i];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < (AISB_MAX_DIMS + AISB_MAX_DIMS * AISB_MAX_DIMS + 255) / 256; ++s) {
      const int q = threadIdx.x + s * 256;
      if (q < nq) {
        double a = 0.0;
        if (q < D) {
          for (int p = 0; p < cnt; ++p) a += tile[p * D + q];
        } else {
          const int i = (q - D) / D, j = (q - D) % D;
          for (int p = 0; p < cnt; ++p) a += static_cast<double>(tile[p * D + i]) * static_cast<double>(tile[p * D + j]);
        }
        acc[s] += a;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < (AISB_MAX_DIMS + AISB_MAX_DIMS * AISB_MAX_DIMS + 255) / 256; ++s) {
    const int q = threadIdx.x + s * 256;
    if (q < nq && acc[s] != 0.0) atomicAdd(&out[q], acc[s]);
  }
}

// ---------------------------------------------------------------------------------------------
// Generic (runtime-shaped) natural-log mixture density of one point, parameters read from global
// memory. Used where the batch is tiny (LAIS chains) and launch count, not throughput, matters.
// ---------------------------------------------------------------------------------------------
__device__ float mixture_logpdf_point(const float* xp, int D, int DP, int kind, const float* __restrict__ rows,
                                      const float* __restrict__ offs, int64_t K, float a_iso, float uniform_offset) {
  const int ROW = row_floats(kind, DP);
  float m = kNegBig, S = 0.f;
  for (int64_t k = 0; k < K; ++k) {
    const float* row = rows + k * ROW;
    float t = offs ? offs[k] : 0.f;
    if (kind == AISB_COV_ISO_SHARED) {
      for (int d = 0; d < D; ++d) {
        const float dl = fmaf(xp[d], a_iso, row[d]);
        t = fmaf(-dl, dl, t);
      }
    } else if (kind == AISB_COV_DIAG) {
      for (int d = 0; d < D; ++d) {
        const float dl = fmaf(xp[d], row[d], row[DP + d]);
        t = fmaf(-dl, dl, t);
      }
    } else {
      const int TRI = DP * (DP + 1) / 2;
      for (int r = 0; r < D; ++r) {
        float z = row[TRI + r];
        for (int c = 0; c <= r; ++c) z = fmaf(row[r * (r + 1) / 2 + c], xp[c], z);
        t = fmaf(-z, z, t);
      }
    }
    const float mn = fmaxf(m, t);
    S = S * ex2(m - mn) + ex2(t - mn);
    m = mn;
  }
  return lse_finish(m, S) + (offs ? 0.f : uniform_offset);
}

// One thread per chain; L sequential Metropolis-Hastings steps (reference layered_ais.py:54-68).
__global__ void lais_mh_kernel(float* __restrict__ means, int64_t N, int D, int DP, int kind,
                               const float* __restrict__ rows, const float* __restrict__ offs, int64_t K, float a_iso,
                               float uniform_offset, const float* __restrict__ delta, const float* __restrict__ u,
                               int L, const float* __restrict__ lo, const float* __restrict__ hi) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c >= N) return;
  float xc[AISB_MAX_DIMS], xh[AISB_MAX_DIMS];
  for (int d = 0; d < D; ++d) xc[d] = means[c * D + d];
  for (int l = 0; l < L; ++l) {
    // the reference re-evaluates pi(x) every step (2 target evaluations per step)
    const float lp_old = mixture_logpdf_point(xc, D, DP, kind, rows, offs, K, a_iso, uniform_offset);
    const float* dl = delta + (static_cast<int64_t>(l) * N + c) * D;
    for (int d = 0; d < D; ++d) xh[d] = fminf(fmaxf(xc[d] + dl[d], lo[d]), hi[d]);
    const float lp_new = mixture_logpdf_point(xh, D, DP, kind, rows, offs, K, a_iso, uniform_offset);
    // linear-domain ratio in float64, including its underflow behaviour (0/0 -> NaN -> reject, x/0 -> accept)
    const double ratio = exp(static_cast<double>(lp_new)) / exp(static_cast<double>(lp_old));
    if (ratio > static_cast<double>(u[static_cast<int64_t>(l) * N + c]))
      for (int d = 0; d < D; ++d) xc[d] = xh[d];
  }
  for (int d = 0; d < D; ++d) means[c * D + d] = xc[d];
}

}  // namespace aisb

using namespace aisb;

static int check_mix_basic(const aisb_packed_mixture* mix, const char* who) {
  if (!mix || !mix->d_rows || mix->K < 1 || aisb_packed_row_floats(mix->cov_kind, mix->D) == 0) {
    set_error("%s: invalid mixture", who);
    return AISB_ERR_INVALID;
  }
  return AISB_OK;
}

static int mpmc_moments_impl(const float* d_x, int64_t n, const aisb_packed_mixture* proposals, const float* d_logq,
                             const float* d_logw, double log_norm, const double* d_record, double* d_out,
                             void* stream);

extern "C" int aisb_mpmc_moments_f32(const float* d_x, int64_t n, const aisb_packed_mixture* proposals,
                                     const float* d_logq, const float* d_logw, double log_norm, double* d_out,
                                     void* stream) {
  return mpmc_moments_impl(d_x, n, proposals, d_logq, d_logw, log_norm, nullptr, d_out, stream);
}

extern "C" int aisb_mpmc_moments_dev_f32(const float* d_x, int64_t n, const aisb_packed_mixture* proposals,
                                         const float* d_logq, const float* d_logw, const double* d_record,
                                         double* d_out, void* stream) {
  if (!d_record) {
    set_error("mpmc_moments_dev: null weight record");
    return AISB_ERR_INVALID;
  }
  return mpmc_moments_impl(d_x, n, proposals, d_logq, d_logw, 0.0, d_record, d_out, stream);
}

static int mpmc_moments_impl(const float* d_x, int64_t n, const aisb_packed_mixture* proposals, const float* d_logq,
                             const float* d_logw, double log_norm, const double* d_record, double* d_out,
                             void* stream) {
  int rc = check_mix_basic(proposals, "mpmc_moments");
  if (rc) return rc;
  if (!proposals->d_offsets) {
    set_error("mpmc_moments: proposals need per-component offsets (log2 alpha_d + normaliser)");
    return AISB_ERR_INVALID;
  }
  if (n < 0 || !d_out || (n > 0 && (!d_x || !d_logq || !d_logw))) {
    set_error("mpmc_moments: null pointer or negative n");
    return AISB_ERR_INVALID;
  }
  if (n == 0) return AISB_OK;
  const int D = proposals->D, DP = aisb_padded_dims(D);
  int vec_ok = 0;
  if (D == DP && DP >= 4) vec_ok = reinterpret_cast<uintptr_t>(d_x) % 16 == 0;
  if (D == DP && DP == 2) vec_ok = reinterpret_cast<uintptr_t>(d_x) % 8 == 0;
  const LogNorm l2n{static_cast<float>(log_norm * kLog2e), d_record};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (proposals->cov_kind) {
    case AISB_COV_ISO_SHARED: return launch_mpmc_dp<AISB_COV_ISO_SHARED>(DP, d_x, n, D, vec_ok, proposals, d_logq, d_logw, l2n, d_out, st);
    case AISB_COV_DIAG: return launch_mpmc_dp<AISB_COV_DIAG>(DP, d_x, n, D, vec_ok, proposals, d_logq, d_logw, l2n, d_out, st);
    case AISB_COV_FULL: return launch_mpmc_dp<AISB_COV_FULL>(DP, d_x, n, D, vec_ok, proposals, d_logq, d_logw, l2n, d_out, st);
    default: break;
  }
  set_error("mpmc_moments: unknown cov_kind %d", proposals->cov_kind);
  return AISB_ERR_INVALID;
}

extern "C" int aisb_mpmc_moments2_f32(const float* d_x, int64_t n, const aisb_packed_mixture* proposals,
                                      const float* d_logq, const float* d_logw, double log_norm, double* d_out,
                                      void* stream) {
  int rc = check_mix_basic(proposals, "mpmc_moments2");
  if (rc) return rc;
  if (!proposals->d_offsets) {
    set_error("mpmc_moments2: proposals need per-component offsets (log2 alpha_d + normaliser)");
    return AISB_ERR_INVALID;
  }
  if (n < 0 || !d_out || (n > 0 && (!d_x || !d_logq || !d_logw))) {
    set_error("mpmc_moments2: null pointer or negative n");
    return AISB_ERR_INVALID;
  }
  if (n == 0) return AISB_OK;
  const int D = proposals->D, DP = aisb_padded_dims(D);
  int vec_ok = 0;
  if (D == DP && DP >= 4) vec_ok = reinterpret_cast<uintptr_t>(d_x) % 16 == 0;
  const LogNorm l2n{static_cast<float>(log_norm * kLog2e), nullptr};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (proposals->cov_kind) {
    case AISB_COV_ISO_SHARED: return launch_mpmc2_dp<AISB_COV_ISO_SHARED>(DP, d_x, n, D, vec_ok, proposals, d_logq, d_logw, l2n, d_out, st);
    case AISB_COV_DIAG: return launch_mpmc2_dp<AISB_COV_DIAG>(DP, d_x, n, D, vec_ok, proposals, d_logq, d_logw, l2n, d_out, st);
    case AISB_COV_FULL: return launch_mpmc2_dp<AISB_COV_FULL>(DP, d_x, n, D, vec_ok, proposals, d_logq, d_logw, l2n, d_out, st);
    default: break;
  }
  set_error("mpmc_moments2: unknown cov_kind %d", proposals->cov_kind);
  return AISB_ERR_INVALID;
}


extern "C" int aisb_mpmc_update_f64(const double* d_mom, const double* d_raw, double n_samples, int64_t N, int32_t D,
                                    double* d_means, double* d_covs, double* d_alpha, double* d_wmix,
                                    float* d_rows_diag, float* d_rows_full, float* d_offsets, float* d_means32,
                                    float* d_chol32, float* d_w32, int32_t* d_status, void* stream) {
  const int DP = aisb_padded_dims(D);
  if (!d_mom || !d_raw || !d_means || !d_covs || !d_alpha || !d_wmix || !d_rows_diag || !d_rows_full || !d_offsets ||
      !d_means32 || !d_chol32 || !d_w32 || !d_status || N < 1 || D < 1 || D > kUpdMaxD || DP == 0) {
    set_error("mpmc_update: null pointer or unsupported shape (N=%lld, D=%d; D <= %d)", static_cast<long long>(N), D,
              kUpdMaxD);
    return D > kUpdMaxD ? AISB_ERR_UNSUPPORTED : AISB_ERR_INVALID;
  }
  const int64_t KP = aisb_padded_components(N);
  const int RF = static_cast<int>(aisb_packed_row_floats(AISB_COV_FULL, D));
  mpmc_update_kernel<<<1, kUpdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      d_mom, d_raw, n_samples, N, D, DP, KP, RF, d_means, d_covs, d_alpha, d_wmix, d_rows_diag, d_rows_full, d_offsets,
      d_means32, d_chol32, d_w32, d_status);
  AISB_LAUNCH_CHECK("mpmc_update_kernel");
  return AISB_OK;
}

extern "C" int aisb_raw_moments_f32(const float* d_x, int64_t n, int32_t D, double* d_out, void* stream) {
  if (D < 1 || D > AISB_MAX_DIMS || n < 0 || !d_out || (n > 0 && !d_x)) {
    set_error("raw_moments: invalid argument (n=%lld D=%d)", static_cast<long long>(n), D);
    return AISB_ERR_INVALID;
  }
  if (n == 0) return AISB_OK;
  const int64_t nchunks = (n + kRawChunk - 1) / kRawChunk;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  const unsigned grid = static_cast<unsigned>(nchunks < cap ? nchunks : cap);
  raw_moments_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, n, D, d_out);
  AISB_LAUNCH_CHECK("raw_moments_kernel");
  return AISB_OK;
}

extern "C" int aisb_lais_mh_f32(float* d_means, int64_t N, const aisb_packed_mixture* target, const float* d_delta,
                                const float* d_u, int32_t L, const float* d_lo, const float* d_hi, void* stream) {
  int rc = check_mix_basic(target, "lais_mh");
  if (rc) return rc;
  if (target->cov_kind != AISB_COV_ISO_SHARED && !target->d_offsets) {
    set_error("lais_mh: offsets are required for cov_kind %d", target->cov_kind);
    return AISB_ERR_INVALID;
  }
  if (N < 0 || L < 0 || (N > 0 && L > 0 && (!d_means || !d_delta || !d_u || !d_lo || !d_hi))) {
    set_error("lais_mh: null pointer or negative size");
    return AISB_ERR_INVALID;
  }
  if (N == 0 || L == 0) return AISB_OK;
  const int D = target->D;
  lais_mh_kernel<<<static_cast<unsigned>((N + 63) / 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      d_means, N, D, aisb_padded_dims(D), target->cov_kind, target->d_rows, target->d_offsets, target->K,
      target->iso_scale, target->uniform_offset, d_delta, d_u, L, d_lo, d_hi);
  AISB_LAUNCH_CHECK("lais_mh_kernel");
  return AISB_OK;
}
