// Host-side pieces of the C ABI: error reporting, packed-layout queries, float64 parameter
// packing and the partial-record algebra (merge / finalise). No kernels here.
#include <stdarg.h>
#include <string.h>

#include <cmath>
#include <limits>
#include <vector>

#include "common.cuh"

namespace aisb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return AISB_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

}  // namespace aisb

using namespace aisb;

extern "C" int aisb_abi_version(void) { return AISB_ABI_VERSION; }
extern "C" const char* aisb_last_error(void) { return g_err; }

extern "C" int32_t aisb_padded_dims(int32_t D) {
  if (D < 1 || D > AISB_MAX_DIMS) return 0;
  int p = 1;
  while (p < D) p *= 2;
  return p;
}

extern "C" int64_t aisb_packed_row_floats(int32_t cov_kind, int32_t D) {
  const int DP = aisb_padded_dims(D);
  if (DP == 0 || cov_kind < AISB_COV_ISO_SHARED || cov_kind > AISB_COV_FULL) return 0;
  if (cov_kind == AISB_COV_FULL && DP > 32) return 0;
  return row_floats(cov_kind, DP);
}

extern "C" int64_t aisb_padded_components(int64_t K) {
  return K <= 0 ? 0 : (K + AISB_COMP_ALIGN - 1) / AISB_COMP_ALIGN * AISB_COMP_ALIGN;
}

extern "C" float aisb_iso_scale(double variance) { return static_cast<float>(std::sqrt(kLog2e / (2.0 * variance))); }

// Mirrors CMixtureModel.set_weights (reference CMixtureModel.py:25-41): NaN / <= 0 -> 0; divide by the sum
// when it is positive (otherwise the zeros are kept, i.e. every component has weight 0).
static void sanitize_weights(const double* w_in, int64_t K, std::vector<double>& w) {
  w.resize(K);
  double sum = 0.0;
  for (int64_t k = 0; k < K; ++k) {
    double v = w_in ? w_in[k] : 1.0 / static_cast<double>(K);
    if (std::isnan(v) || v <= 0.0) v = 0.0;
    w[k] = v;
    sum += v;
  }
  if (sum > 0.0)
    for (int64_t k = 0; k < K; ++k) w[k] /= sum;
}

// Gauss-Jordan inverse with partial pivoting; also returns det(C). false if singular.
static bool invert_with_det(const double* C, int D, double* P, double* det) {
  std::vector<double> M(C, C + static_cast<size_t>(D) * D);
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) P[i * D + j] = i == j ? 1.0 : 0.0;
  double d = 1.0;
  for (int c = 0; c < D; ++c) {
    int piv = c;
    for (int r = c + 1; r < D; ++r)
      if (std::fabs(M[r * D + c]) > std::fabs(M[piv * D + c])) piv = r;
    if (M[piv * D + c] == 0.0) return false;
    if (piv != c) {
      for (int j = 0; j < D; ++j) {
        std::swap(M[piv * D + j], M[c * D + j]);
        std::swap(P[piv * D + j], P[c * D + j]);
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

extern "C" int aisb_pack_mixture_f64(const double* h_means, const double* h_cov, const double* h_weights, int64_t K,
                                     int32_t D, int32_t cov_kind, float* h_rows, float* h_offsets, float* iso_scale) {
  const int DP = aisb_padded_dims(D);
  const int64_t RF = aisb_packed_row_floats(cov_kind, D);
  if (!h_means || !h_cov || !h_rows || !h_offsets || K < 1 || DP == 0 || RF == 0) {
    set_error("pack_mixture: invalid argument (K=%lld D=%d cov_kind=%d)", static_cast<long long>(K), D, cov_kind);
    return AISB_ERR_INVALID;
  }
  const int64_t KP = aisb_padded_components(K);
  const float ninf = -std::numeric_limits<float>::infinity();
  std::vector<double> w;
  sanitize_weights(h_weights, K, w);
  memset(h_rows, 0, sizeof(float) * KP * RF);
  for (int64_t k = K; k < KP; ++k) h_offsets[k] = ninf;
  const double base = -0.5 * D * kLog2Pi;

  if (cov_kind == AISB_COV_ISO_SHARED) {
    const double v = h_cov[0];
    if (!(v > 0.0)) {
      set_error("pack_mixture: variance must be positive (got %g)", v);
      return AISB_ERR_INVALID;
    }
    const double a = std::sqrt(kLog2e / (2.0 * v));
    const float af = static_cast<float>(a);
    if (iso_scale) *iso_scale = af;
    const double logdet = D * std::log(v);
    for (int64_t k = 0; k < K; ++k) {
      // b is built from the float-rounded scale so that fma(x, a, b) cancels exactly at x == mu
      for (int d = 0; d < D; ++d) h_rows[k * RF + d] = static_cast<float>(-static_cast<double>(af) * h_means[k * D + d]);
      h_offsets[k] = w[k] > 0.0 ? static_cast<float>(kLog2e * (std::log(w[k]) + base - 0.5 * logdet)) : ninf;
    }
    return AISB_OK;
  }

  if (cov_kind == AISB_COV_DIAG) {
    for (int64_t k = 0; k < K; ++k) {
      double logdet = 0.0;
      for (int d = 0; d < D; ++d) {
        const double v = h_cov[k * D + d];
        if (!(v > 0.0)) {
          set_error("pack_mixture: component %lld has non-positive variance %g (reference asserts det > 0)",
                    static_cast<long long>(k), v);
          return AISB_ERR_INVALID;
        }
        logdet += std::log(v);
        const float af = static_cast<float>(std::sqrt(kLog2e / (2.0 * v)));
        h_rows[k * RF + d] = af;
        h_rows[k * RF + DP + d] = static_cast<float>(-static_cast<double>(af) * h_means[k * D + d]);
      }
      h_offsets[k] = w[k] > 0.0 ? static_cast<float>(kLog2e * (std::log(w[k]) + base - 0.5 * logdet)) : ninf;
    }
    return AISB_OK;
  }

  // FULL. The reference evaluates -0.5 * d^T inv(C) d with inv_cov = np.linalg.inv(C) and log det from
  // np.linalg.det(C) (CMultivariateNormal.py:55-62); C is not required to be symmetric there, so neither here:
  //   P = sym(inv(C));  find lower-triangular A with A^T A = (LOG2E/2) P  (Cholesky of the index-reversed P);
  //   b = -A mu;  log det = log(det(C)) from the elimination.
  std::vector<double> P(static_cast<size_t>(D) * D), G(static_cast<size_t>(D) * D);
  const double s = std::sqrt(kLog2e / 2.0);
  const int TRI = DP * (DP + 1) / 2;
  for (int64_t k = 0; k < K; ++k) {
    double det = 0.0;
    if (!invert_with_det(h_cov + k * D * D, D, P.data(), &det) || !(det > 0.0)) {
      set_error("pack_mixture: component %lld covariance has det <= 0 (reference asserts det > 0)",
                static_cast<long long>(k));
      return AISB_ERR_INVALID;
    }
    const double logdet = std::log(det);
    // Q = J sym(P) J (rows and columns reversed), Q = G G^T
    std::fill(G.begin(), G.end(), 0.0);
    for (int i = 0; i < D; ++i) {
      for (int j = 0; j <= i; ++j) {
        const int ri = D - 1 - i, rj = D - 1 - j;
        double acc = 0.5 * (P[ri * D + rj] + P[rj * D + ri]);
        for (int t = 0; t < j; ++t) acc -= G[i * D + t] * G[j * D + t];
        if (i == j) {
          if (!(acc > 0.0)) {
            set_error("pack_mixture: component %lld precision matrix is not positive definite",
                      static_cast<long long>(k));
            return AISB_ERR_INVALID;
          }
          G[i * D + i] = std::sqrt(acc);
        } else {
          G[i * D + j] = acc / G[j * D + j];
        }
      }
    }
    float* row = h_rows + k * RF;
    for (int r = 0; r < D; ++r) {
      double br = 0.0;
      for (int c = 0; c <= r; ++c) {
        // A = J G^T J  ->  A[r][c] = G[D-1-c][D-1-r]
        const float af = static_cast<float>(s * G[(D - 1 - c) * D + (D - 1 - r)]);
        row[r * (r + 1) / 2 + c] = af;
        br -= static_cast<double>(af) * h_means[k * D + c];
      }
      row[TRI + r] = static_cast<float>(br);
    }
    h_offsets[k] = w[k] > 0.0 ? static_cast<float>(kLog2e * (std::log(w[k]) + base - 0.5 * logdet)) : ninf;
  }
  return AISB_OK;
}

// ---- weight partials {M, S1, S2, n} ---------------------------------------------------------
extern "C" void aisb_weight_partials_merge(const double* a, const double* b, double* out) {
  const double M = a[0] > b[0] ? a[0] : b[0];
  const double fa = std::exp(a[0] - M), fb = std::exp(b[0] - M);
  const double r[4] = {M, a[1] * fa + b[1] * fb, a[2] * fa * fa + b[2] * fb * fb, a[3] + b[3]};
  memcpy(out, r, sizeof(r));
}
extern "C" double aisb_weight_partials_logsum(const double* p) { return p[0] + std::log(p[1]); }
extern "C" double aisb_weight_partials_ess(const double* p) { return p[2] > 0.0 ? p[1] * p[1] / p[2] : 0.0; }

// ---- KLD partials {m_p, S_p, A, m_q, S_q, n_in, n_posinf_q, n_total} ------------------------
extern "C" void aisb_kld_partials_merge(const double* a, const double* b, double* out) {
  const double mp = a[0] > b[0] ? a[0] : b[0], mq = a[3] > b[3] ? a[3] : b[3];
  const double fpa = std::exp(a[0] - mp), fpb = std::exp(b[0] - mp);
  const double fqa = std::exp(a[3] - mq), fqb = std::exp(b[3] - mq);
  const double r[8] = {mp, a[1] * fpa + b[1] * fpb, a[2] * fpa + b[2] * fpb, mq, a[4] * fqa + b[4] * fqb,
                       a[5] + b[5], a[6] + b[6], a[7] + b[7]};
  memcpy(out, r, sizeof(r));
}
extern "C" double aisb_kld_finalize(const double* p) {
  if (p[5] <= 0.0) return 0.0;                                         // np.sum of an empty array
  if (p[6] > 0.0) return std::numeric_limits<double>::infinity();      // res[isnan(q)] = inf
  const double Lp = p[0] + std::log(p[1]), Lq = p[3] + std::log(p[4]);
  return p[2] / p[1] - Lp + Lq;
}
