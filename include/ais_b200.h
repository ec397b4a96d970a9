/*
 * ais_b200.h -- C ABI of libais_b200.so: the B200 (sm_100a) implementation of the
 * ais-benchmarks density / importance-weight / KLD hot path.
 *
 * The reference (IntelLabs/ais-benchmarks) is pure Python/numpy and has no FFI of
 * its own; the binding a maintainer adds is a ctypes stub.
 * Each entry point below names the reference code it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *  - Pointers named d_* are DEVICE pointers owned by the caller. Pointers named
 *    h_* are HOST pointers. The library never allocates or frees caller memory;
 *    scratch space is passed in explicitly (see the *_bytes queries).
 *  - Every device call is asynchronous on `stream` (a cudaStream_t passed as
 *    void*; NULL = legacy default stream).
 *  - Return value: 0 (AISB_OK) or a negative aisb_status. No exception crosses
 *    the ABI. aisb_last_error() returns a thread-local message for the last failure.
 *  - All point sets are batch-first row-major float32 [n, D] (the reference's
 *    (N, D) float64 layout, distributions/distributions.py:35-37, cast to fp32).
 *  - Component parameters are passed PACKED (see "Packed mixture layout");
 *    packing is done once per parameter update in float64 on the host
 *    (aisb_pack_mixture_f64) or on device for KDE centres (aisb_kde_pack_f32).
 */
#ifndef AIS_B200_H
#define AIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AISB_ABI_VERSION 1

typedef enum aisb_status {
  AISB_OK = 0,
  AISB_ERR_INVALID = -1,     /* bad argument (null pointer, D out of range, ...) */
  AISB_ERR_UNSUPPORTED = -2, /* shape/kind combination not built */
  AISB_ERR_CUDA = -3,        /* a CUDA runtime call failed (message has the cudaError string) */
  AISB_ERR_WORKSPACE = -4    /* workspace too small */
} aisb_status;

/* Covariance kinds of a packed Gaussian mixture. */
typedef enum aisb_cov_kind {
  AISB_COV_ISO_SHARED = 0, /* every component has covariance v*I with the SAME v (KDE, DM-AIS/LAIS proposals) */
  AISB_COV_DIAG = 1,       /* per-component diagonal covariance (CGaussianMixtureModel) */
  AISB_COV_FULL = 2        /* per-component full covariance (CMultivariateNormal, M-PMC after adapt) */
} aisb_cov_kind;

#define AISB_MAX_DIMS 64

/*
 * Packed mixture layout (all float32, log2 domain; LOG2E = 1/ln 2).
 *
 * For a component k with mean mu_k, covariance Sigma_k and normalised weight w_k
 * (CMixtureModel.set_weights, distributions/mixture/CMixtureModel.py:25-41):
 *     log2 (w_k N(x; mu_k, Sigma_k)) = c_k - || A_k x + b_k ||^2
 *   c_k = LOG2E * ( log w_k - 0.5 * (D log 2pi + log det Sigma_k) )      (-inf if w_k == 0)
 *   A_k = sqrt(LOG2E / 2) * L_k^{-1},  Sigma_k = L_k L_k^T (Cholesky),  b_k = -A_k mu_k
 * (CMultivariateNormal.log_prob, distributions/parametric/CMultivariateNormal.py:67-71).
 *
 * DP = aisb_padded_dims(D) is D rounded up to {1,2,4,8,16,32,64}; KP = K rounded up
 * to a multiple of AISB_COMP_ALIGN. Padding dims hold 0; padding components hold
 * c = -inf (offsets present) or coordinate 1e18 (no offsets).
 *
 *  ISO_SHARED: rows[KP][DP]    = b_k = -a * mu_k,  a = sqrt(LOG2E / (2 v)) passed as `iso_scale`
 *  DIAG:       rows[KP][2*DP]  = a_k[0..DP) then b_k[0..DP),  a_kd = sqrt(LOG2E / (2 v_kd)), b_kd = -a_kd mu_kd
 *  FULL:       rows[KP][RF], RF = DP*(DP+1)/2 + DP rounded up to a multiple of 4
 *              = lower triangle of A_k row-major (A00, A10, A11, A20, ...), then b_k, then zero padding (D <= 32 only)
 *  offsets[KP] = c_k  (may be NULL for ISO_SHARED: all components share `uniform_offset`)
 */
#define AISB_COMP_ALIGN 32

typedef struct aisb_packed_mixture {
  const float* d_rows;    /* device, layout above, 16-byte aligned */
  const float* d_offsets; /* device [KP] or NULL (ISO_SHARED only) */
  int64_t K;              /* real number of components */
  int32_t D;              /* real dimensionality, 1..AISB_MAX_DIMS */
  int32_t cov_kind;       /* aisb_cov_kind */
  float iso_scale;        /* ISO_SHARED: a = sqrt(LOG2E / (2 v)); ignored otherwise */
  float uniform_offset;   /* natural-log constant added to every result when d_offsets == NULL:
                             log w - 0.5 * D * log(2 pi v); ignored otherwise */
} aisb_packed_mixture;

int aisb_abi_version(void);
const char* aisb_last_error(void);

/* Rounded-up dimensionality used by the packed layouts; 0 if D is unsupported. */
int32_t aisb_padded_dims(int32_t D);
/* Floats per packed row for (kind, D); 0 if unsupported. */
int64_t aisb_packed_row_floats(int32_t cov_kind, int32_t D);
/* K rounded up to AISB_COMP_ALIGN. */
int64_t aisb_padded_components(int64_t K);
/* a = (float) sqrt(LOG2E / (2 v)): the ISO_SHARED scale for kernel variance v (same rounding everywhere). */
float aisb_iso_scale(double variance);

/*
 * HOST-side float64 -> packed float32 conversion (no CUDA). Replaces the per-object
 * precompute of CMultivariateNormal.__init__/set_moments (CMultivariateNormal.py:26-62:
 * det, inv, log_det, term1, term2) and CMixtureModel.set_weights (CMixtureModel.py:25-41,
 * NaN/<=0 -> 0, then / sum if sum > 0).
 *   h_means [K, D]; h_cov: ISO_SHARED -> [1] variance; DIAG -> [K, D] variances;
 *   FULL -> [K, D, D] covariance matrices; h_weights [K] raw weights or NULL (= equal, 1/K).
 *   h_rows [KP * row_floats], h_offsets [KP] out.
 * Returns AISB_ERR_INVALID if a covariance is not positive definite (reference: assert det > 0).
 */
int aisb_pack_mixture_f64(const double* h_means, const double* h_cov, const double* h_weights, int64_t K, int32_t D,
                          int32_t cov_kind, float* h_rows, float* h_offsets, float* iso_scale);

/*
 * KDE model build on device. Replaces CMixtureSamplingMethod._update_model
 * (sampling_methods/base.py:162-179): n_c isotropic kernels N(x_idx[j], bw*I) with equal weights.
 *   d_samples [n, D]; d_idx [n_c] int64 row indices into d_samples, or NULL (identity, n_c <= n);
 *   bw = kernel VARIANCE (reference quirk: "bw" goes on the covariance diagonal);
 *   d_rows out [aisb_padded_components(n_c) * DP] (ISO_SHARED layout).
 * The matching aisb_packed_mixture is {d_rows, NULL, n_c, D, ISO_SHARED, sqrt(LOG2E/(2 bw)),
 *   -log(n_c) - 0.5 * D * log(2 pi bw)}.
 */
int aisb_kde_pack_f32(const float* d_samples, int64_t n, int32_t D, const int64_t* d_idx, int64_t n_c, double bw,
                      float* d_rows, void* stream);

/*
 * Workspace (bytes) needed by aisb_mixture_logpdf_f32 / aisb_dm_weights_f32 for n points
 * against K components (split-component partials + block partials).
 */
int aisb_workspace_bytes(int64_t n, int64_t K, int32_t D, size_t* bytes);

/*
 * Mixture log-density:  out_logp[i] = log sum_k w_k N(x_i; mu_k, Sigma_k)   (natural log).
 * Replaces CMixtureModel.log_prob (CMixtureModel.py:43-57: loop over components +
 * scipy logsumexp(b=weights)), CGaussianMixtureModel.log_prob (CGaussianMixtureModel.py:40-41),
 * CMultivariateNormal.log_prob (K == 1) and the KDE evaluation reached through
 * CMixtureSamplingMethod.log_prob (sampling_methods/base.py:156-159).
 */
int aisb_mixture_logpdf_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix, float* d_out_logp,
                            void* d_workspace, size_t workspace_bytes, void* stream);
/*
 * Same result, for callers that have put BOTH the points and the components of an ISO_SHARED mixture without offsets
 * (a KDE) in a spatially coherent order (e.g. along a Z-order curve): selects kernel variants that skip the
 * log-sum-exp bookkeeping of a group of components when it is negligible (2^-(28 + log2 K) of the running maximum) for
 * every point of a warp. Purely a performance hint: on unordered data the result is the same and the vote costs 2-5 %;
 * other mixture kinds take the ordinary kernels.
 */
int aisb_mixture_logpdf_ordered_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix, float* d_out_logp,
                                    void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * Screen-and-refine evaluation of a spatially ordered KDE (same contract and same result as
 * aisb_mixture_logpdf_ordered_f32 to float32 accuracy; replaces the same reference code, CMixtureModel.py:43-57 reached
 * through sampling_methods/base.py:156-179). Every (point, component) pair is first evaluated in the expanded form
 * ||a x||^2 + 2 a (h_j + x . b_j), h_j = ||b_j||^2 / (2 a), which costs half the FMAs of the float32-safe centred form
 * but is only good enough to DECIDE whether the pair can matter; the exact centred evaluation and the log-sum-exp fold
 * run only for groups of components that are not negligible (2^-(21 + log2 KP) of the running maximum, minus a
 * rounding-error allowance that grows with the distance of the data from the origin) for at least one of the 64
 * consecutive points that share a packed register set. Needs: ISO_SHARED without offsets, D in {2, 4, 8, 16}, d_x
 * 16-byte aligned, points and components in a spatially coherent order (else it is correct but slow).
 *   aisb_kde_screen_floats(K)   floats of the screen vector: h_j for the KP packed components + an 8-float trailer
 *                               (trailer[0] = max_j ||b_j||^2)
 *   aisb_kde_screen_pack_f32    builds it from the packed mixture (once per KDE)
 *   d_stats                     NULL, or 2 zero-initialised uint64 counters that receive {refined, screened} units
 *                               (unit = 64 points x one group of components); selects an instrumented kernel
 */
int64_t aisb_kde_screen_floats(int64_t K);
int aisb_kde_screen_pack_f32(const aisb_packed_mixture* mix, float* d_screen, void* stream);
int aisb_mixture_logpdf_screened_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix, const float* d_screen,
                                     float* d_out_logp, void* d_workspace, size_t workspace_bytes,
                                     unsigned long long* d_stats, void* stream);

/*
 * Fast evaluation of a spatially ordered KDE (same contract as aisb_mixture_logpdf_ordered_f32; replaces the same
 * reference code, CMixtureModel.py:43-57 reached through sampling_methods/base.py:156-179). Every warp shifts its 128
 * consecutive points and each staged chunk of components by the warp's centroid and evaluates the pair exponent in
 * the EXPANDED form ||a x'||^2 + 2 a (h'_j + x' . b'_j) -- half the FMAs of the centred form; after the shift its
 * float32 error only depends on local distances in bandwidth units. The kernel bounds that error per point,
 *     eps = 4 * 2^-24 * (||a x'||^2 + |t_max| + 1)      (measured model with a 1.7x margin),
 * lists the points with eps > 3e-6 * max(1, |log p|) and re-evaluates exactly those with the centred form in a second,
 * normally empty, launch: results stay within the 1e-5 parity bar whatever the data, only the speed depends on it.
 * Groups of components that are negligible (2^-(21 + log2 KP) of the running maximum) for all 64 points of a packed
 * register set are skipped after the screen. Needs: ISO_SHARED without offsets, D in {2, 4, 8, 16}, d_x 16-byte aligned,
 * points and components in a spatially coherent order (else it is correct but slow).
 *   d_scratch  aisb_kde_fast_scratch_bytes(n) bytes, 16-byte aligned (list of points to re-evaluate, per-point shifts)
 *   d_stats    NULL, or 3 zero-initialised uint64: {folded, screened} units (64 points x one group of components) and
 *              the number of points re-evaluated exactly; selects an instrumented kernel
 */
size_t aisb_kde_fast_scratch_bytes(int64_t n);
int aisb_mixture_logpdf_fast_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix, float* d_out_logp,
                                 void* d_workspace, size_t workspace_bytes, void* d_scratch, size_t scratch_bytes,
                                 unsigned long long* d_stats, void* stream);

/*
 * Deterministic-mixture importance weights in one pass over the samples:
 *   logp_i = log pi(x_i) (target mixture, or read from d_logp_in when target == NULL),
 *   logq_i = log q(x_i)  (proposal mixture),  logw_i = logp_i - logq_i.
 * Replaces the per-sample loop `w = target_d.prob(s) / self.prob(s)` of
 * CDeterministicMixtureAIS.importance_sample (sampling_methods/dm_ais.py:72-76),
 * CLayeredAIS.importance_sample (layered_ais.py:90-94) and the pi_x/q_x part of
 * CMixturePMC.importance_sample (m_pmc.py:93-96).
 *   d_out_logw [n]; d_out_logp, d_out_logq [n] or NULL;
 *   d_out_partials [4] doubles: {M = max_i logw_i, sum_i exp(logw_i - M), sum_i exp(2 (logw_i - M)), n_finite}
 *   computed over finite logw_i only (reference weights that are inf/nan are excluded).
 */
int aisb_dm_weights_f32(const float* d_x, int64_t n, const aisb_packed_mixture* target, const float* d_logp_in,
                        const aisb_packed_mixture* proposals, float* d_out_logw, float* d_out_logp, float* d_out_logq,
                        double* d_out_partials, void* d_workspace, size_t workspace_bytes, void* stream);

/*
 * Partials of an existing log-weight vector (same 4 doubles as above). Used when samples
 * accumulate over iterations (dm_ais.py:81-86 normalises over ALL iterations).
 */
int aisb_logw_partials_f32(const float* d_logw, int64_t n, double* d_out_partials, void* d_workspace,
                           size_t workspace_bytes, void* stream);

/*
 * Normalised weights  w_i = exp(logw_i - (M + log S))  from (possibly all-reduced) partials.
 * Replaces `self.weights / self.weights.sum()` (dm_ais.py:86, m_pmc.py:108).
 * h_partials: the 4 doubles on the HOST (after any cross-rank merge).
 */
int aisb_normalize_weights_f32(const float* d_logw, int64_t n, const double* h_partials, float* d_out_w, void* stream);

/*
 * Device-resident variants (no host round trip between the weight kernel and the normalised weights; this is what
 * lets a multi-GPU or streamed weighting run without a synchronisation):
 *   aisb_weight_records_merge / aisb_kld_records_merge: fold n_records records that live in DEVICE memory (one per
 *       rank after an all-gather, or one per streamed chunk) into d_out_partials (device, 4 resp. 8 doubles) with the
 *       same log-sum-exp rescale rule as the host helpers below.
 *   aisb_normalize_weights_dev_f32: aisb_normalize_weights_f32 with the record(s) read from device memory: d_records
 *       holds n_records (1..4096) records, folded inside the kernel (so an all-gathered [ranks, 4] buffer can be passed
 *       as it is); d_out_stats (device, nullable) receives {log sum_i w_i, ESS = (sum w)^2 / sum w^2, n_finite}
 *       (sampling_methods/base.py:72-75 without leaving HBM).
 */
int aisb_weight_records_merge(const double* d_records, int64_t n_records, double* d_out_partials, void* stream);

/*
 * Cross-rank exchange of the weight records over NVLink peer memory, as one tiny kernel instead of a collective launch
 * (the exchange is 32 bytes per rank: latency is everything). Every rank allocates a SYMMETRIC buffer of
 * aisb_weight_exchange_doubles(world) doubles, zero-filled once, and passes d_peer_bufs = device array of the `world`
 * peer-mapped base pointers of those buffers (own buffer at index `rank`). All ranks call with the same `epoch`
 * (1, 2, 3, ...: one per exchange, in lockstep). d_out_records [world, 4] receives every rank's record (rank-major,
 * identical on all ranks; feed it to aisb_normalize_weights_dev_f32). A peer that does not show up within ~17 s
 * poisons its record (n_finite = -1) instead of hanging the GPU.
 */
int64_t aisb_weight_exchange_doubles(int32_t world);
int aisb_weight_records_exchange(const double* d_local_record, void* const* d_peer_bufs, int32_t rank, int32_t world,
                                 uint64_t epoch, double* d_out_records, void* stream);
int aisb_kld_records_merge(const double* d_records, int64_t n_records, double* d_out_partials, void* stream);
int aisb_normalize_weights_dev_f32(const float* d_logw, int64_t n, const double* d_records, int32_t n_records,
                                   float* d_out_w, double* d_out_stats, void* stream);

/*
 * Host-side helpers on partials (no CUDA): merge two partial records with the
 * log-sum-exp rescale rule; ESS = 1 / sum w~^2 (sampling_methods/base.py:72-75).
 */
void aisb_weight_partials_merge(const double* a, const double* b, double* out);
double aisb_weight_partials_ess(const double* partials);
double aisb_weight_partials_logsum(const double* partials);

/*
 * KLD one-pass partials over (lp, lq) pairs:
 *   inliers I = { i : lp_i > -inf and lq_i > -inf }  (metrics/divergences.py:138-139)
 *   d_out_partials [8] doubles: { m_p, S_p = sum_I exp(lp - m_p), A = sum_I exp(lp - m_p) (lp - lq),
 *                                  m_q, S_q = sum_I exp(lq - m_q), n_inlier, n_nan_q, n_total }
 * Replaces CKLDivergence.compute_from_log_probs (divergences.py:136-153).
 */
int aisb_kld_partials_f32(const float* d_lp, const float* d_lq, int64_t n, double* d_out_partials, void* d_workspace,
                          size_t workspace_bytes, void* stream);
/* Host: merge two KLD partial records / finish:  KLD = A/S_p - (m_p + log S_p) + (m_q + log S_q);
 * +inf if any lq in I is NaN; 0.0 if I is empty (divergences.py:147-153, 105). */
void aisb_kld_partials_merge(const double* a, const double* b, double* out);
double aisb_kld_finalize(const double* partials);

/*
 * JSD and EVMSE on the same (lp, lq) device vectors.
 *   aisb_pair_partials_f32: aisb_kld_partials_f32 with the inlier set { lp > floor and lq > floor }. The JSD of the
 *       reference works on linear-domain probabilities and keeps the pairs with p > 0 and q > 0
 *       (metrics/divergences.py:165-173): in the log domain that is floor = AISB_LOG_TINY, below which float64 exp()
 *       returns 0.
 *   aisb_jsd_terms_f32: with a = lp - log_norm_p, b = lq - log_norm_q (the inlier normalisers m + log S of the
 *       merged pair partials), log m = logaddexp(a, b) - log 2:
 *       d_out2 = { sum_I 0.5 e^a (a - log m) + 0.5 e^b (b - log m)  [nats],  number of NaN terms dropped }
 *       (divergences.py:180-192); JSD = d_out2[0] / log 2 (:194-195). d_out2 is zeroed by the call.
 *   aisb_weighted_mean_f32: d_out[D] = sum_i x_i exp(d_logw[i] - log_norm), the self-normalised expectation of
 *       metrics/statistics.py:39-45 when log_norm = LSE(d_logw) (aisb_logw_partials_f32 + aisb_weight_partials_logsum);
 *       EVMSE = || E_p[x] - E_q[x] ||^2 (:46-48). d_out is zeroed by the call. 1 <= D <= AISB_MAX_DIMS.
 */
#define AISB_LOG_TINY (-745.13f)
int aisb_pair_partials_f32(const float* d_lp, const float* d_lq, int64_t n, float floor, double* d_out_partials,
                           void* d_workspace, size_t workspace_bytes, void* stream);
int aisb_jsd_terms_f32(const float* d_lp, const float* d_lq, int64_t n, float floor, double log_norm_p,
                       double log_norm_q, double* d_out2, void* stream);
int aisb_weighted_mean_f32(const float* d_x, int64_t n, int32_t D, const float* d_logw, double log_norm,
                           double* d_out, void* stream);

/*
 * M-PMC Rao-Blackwellised moment accumulation (sampling_methods/m_pmc.py:57-61, 102-105):
 *   rho_{d,i} = exp(log alpha_d + log q_d(x_i) - log q(x_i)),
 *   out[d][0]      = sum_i w~_i rho_{d,i}            (alpha_d numerator)
 *   out[d][1..D]   = sum_i w~_i rho_{d,i} x_i        (mean numerator)
 * w~_i = exp(logw_i - log_norm).  `proposals` offsets must hold log2 (alpha_d N-normaliser).
 *   d_out [K, D+1] doubles, ZEROED by the caller (the kernel accumulates with atomics so
 *   that ranks / batches can be chained).
 */
int aisb_mpmc_moments_f32(const float* d_x, int64_t n, const aisb_packed_mixture* proposals, const float* d_logq,
                          const float* d_logw, double log_norm, double* d_out, void* stream);
/* Same plus the weighted second moments sum_i w~_i rho_di x_i x_i^T (lower triangle, row-major: index r (r+1)/2 + c)
 * behind the D + 1 first-order columns: d_out [K, 1 + D + D (D + 1) / 2], zeroed by the caller, D <= 8. This is the
 * sufficient statistic of the covariance update of M-PMC as published (eq. 14.3 of the paper the reference cites,
 * m_pmc.py:62-68 implements the unweighted scatter instead); used by CMixturePMC(paper_covariance=True). */
int aisb_mpmc_moments2_f32(const float* d_x, int64_t n, const aisb_packed_mixture* proposals, const float* d_logq,
                           const float* d_logw, double log_norm, double* d_out, void* stream);

/* aisb_mpmc_moments_f32 with the normaliser taken from a weight record {M, S1, S2, n} that lives in DEVICE memory
 * (log_norm = M + log S1, e.g. the d_out_partials of aisb_dm_weights_f32 or a merged multi-rank record): no host round
 * trip between the weight kernel and the moment kernel. */
int aisb_mpmc_moments_dev_f32(const float* d_x, int64_t n, const aisb_packed_mixture* proposals, const float* d_logq,
                              const float* d_logw, const double* d_record, double* d_out, void* stream);

/*
 * M-PMC parameter update on the device. Replaces CMixturePMC.adapt (sampling_methods/m_pmc.py:53-78) together with the
 * CMultivariateNormal.set_moments calls (CMultivariateNormal.py:52-62) and CMixtureModel.set_weights
 * (CMixtureModel.py:25-41) it ends in: alpha_d = mom[d][0] / sum, mu_d = mom[d][1..D] / mom[d][0],
 * Sigma_d = raw scatter about mu_d (quirk Q3) or 10*I when any entry is NaN or > 10, weights sanitised, and the
 * proposal mixture RE-PACKED for the next weighting without leaving HBM.
 *   d_mom [N, D+1]      output of aisb_mpmc_moments*_f32 (summed over ranks)
 *   d_raw [D + D*D]     output of aisb_raw_moments_f32 (summed over ranks); n_samples = number of points behind it
 *   d_means [N, D], d_covs [N, D, D], d_alpha [N] (normalised, may hold NaN like the reference's wproposals),
 *   d_wmix [N] (sanitised mixture weights): float64 state, read by the host only when somebody asks for it
 *   d_rows_diag [KP, 2 DP], d_rows_full [KP, RF], d_offsets [KP]: the packed mixture in the DIAG and in the FULL layout
 *       (use DIAG iff status[1] == 0); d_means32 [N, D], d_chol32 [N, D, D] (lower-triangular): sampler parameters;
 *   d_w32 [N]: the mixture weights as float32 (component draws)
 *   d_status [4] int32: {proposals on the 10*I failsafe, proposals with off-diagonal covariance, proposals whose
 *       covariance is not positive definite / det <= 0 (the reference raises AssertionError), all weights zero}
 * D <= 16.
 */
int aisb_mpmc_update_f64(const double* d_mom, const double* d_raw, double n_samples, int64_t N, int32_t D,
                         double* d_means, double* d_covs, double* d_alpha, double* d_wmix, float* d_rows_diag,
                         float* d_rows_full, float* d_offsets, float* d_means32, float* d_chol32, float* d_w32,
                         int32_t* d_status, void* stream);

/*
 * First and second raw moments of a point set, in float64:  out = { sum_i x_i [D], sum_i x_i x_i^T [D*D] }.
 * Used for the (unweighted, quirk Q3) scatter of m_pmc.py:64-68. d_out [D + D*D] doubles, zeroed by the caller.
 */
int aisb_raw_moments_f32(const float* d_x, int64_t n, int32_t D, double* d_out, void* stream);

/*
 * LAIS upper layer: N independent Metropolis-Hastings chains, L steps each, all on device
 * (sampling_methods/layered_ais.py:54-74). Noise is supplied by the caller so that the
 * result is a pure function of its inputs (RNG streams are not part of the contract):
 *   d_means [N, D] in/out (chain states = proposal means); d_delta [L, N, D] ~ N(0, mh_sigma I);
 *   d_u [L, N] ~ U(0,1); proposals are clipped to [d_lo, d_hi] ([D] each); accept when
 *   pi(x_hat) / pi(x) > u  (linear-domain ratio, layered_ais.py:63-66).
 */
int aisb_lais_mh_f32(float* d_means, int64_t N, const aisb_packed_mixture* target, const float* d_delta,
                     const float* d_u, int32_t L, const float* d_lo, const float* d_hi, void* stream);

/*
 * Mixture of axis-aligned boxes ("haar" leaves of TP-AIS and CMultivariateUniform):
 *   out_logp[i] = log sum_k w_k [lo_k < x_i (<|<=) hi_k] / vol_k
 * open_upper != 0: lo < x <  hi  (tree_pyramid.py:595-596);  0: lo < x <= hi (CMultivariateUniform.py:44).
 *   d_lo, d_hi [K, D]; d_logw_over_vol [K] = log(w_k / vol_k).  D <= 32.
 */
int aisb_box_mixture_logpdf_f32(const float* d_x, int64_t n, int32_t D, const float* d_lo, const float* d_hi,
                                const float* d_logw_over_vol, int64_t K, int32_t open_upper, float* d_out_logp,
                                void* stream);

/*
 * Device-side sampling (Philox4x32-10 + Box-Muller). RNG streams are NOT comparable with numpy's.
 *  aisb_sample_iso_f32: out[(k*per + j), :] = means[k] + sqrt(var) * z  for k < K, j < per
 *      (DM-AIS / LAIS: K samples from each of N proposals, dm_ais.py:64-69).
 *  aisb_uniform_box_f32: out[i, d] ~ U(lo_d, hi_d)  (CDivergence.compute eval points, divergences.py:84-85).
 */
/*  aisb_morton_keys_f32: Z-order keys of the rows of d_x [n, D] inside the box [d_lo, d_lo + d_span] (device vectors
 *      [D]): `bits` bits per coordinate interleaved, bits * D <= 32, stored so that a SIGNED 32-bit sort of d_keys [n]
 *      yields the curve order. Used to put large KDEs and their evaluation points in a spatially coherent order (see
 *      aisb_mixture_logpdf_ordered_f32); no reference counterpart. */
int aisb_morton_keys_f32(const float* d_x, int64_t n, int32_t D, int32_t bits, const float* d_lo, const float* d_span,
                         int32_t* d_keys, void* stream);
/*  aisb_morton_keys_auto_f32: the same keys over the points' OWN bounding box, which it computes first in one pass
 *      (d_box: 2*D uint32 of scratch: order-preserving encodings of the per-dimension minima / maxima; NaN coordinates
 *      are ignored). One call instead of two reductions + the key kernel: the Z-order of a point set is on the serial
 *      path of every KDE evaluation on every rank. */
int aisb_morton_keys_auto_f32(const float* d_x, int64_t n, int32_t D, int32_t bits, uint32_t* d_box, int32_t* d_keys,
                              void* stream);
/*  aisb_gather_rows_f32: d_out[i, :] = d_x[d_idx[i], :] for i < m (rows of D floats; d_x has n_src rows). Puts point
 *      sets in Z-order (`x[perm]`) and deals out a rank's block-cyclic share; indices outside [0, n_src) are skipped and
 *      counted in *d_err (device int, may be null). No reference counterpart (numpy fancy indexing on the host). */
int aisb_gather_rows_f32(const float* d_x, int64_t n_src, int32_t D, const int64_t* d_idx, int64_t m, float* d_out,
                         int* d_err, void* stream);
/*  aisb_upload_as_f32: h_src (HOST, pageable or pinned; `count` float64 values if src_is_f64 else float32) -> d_dst
 *      (DEVICE, float32 [count]). The reference-facing classes are handed numpy float64 (README.md:41-43,
 *      benchmark/CBenchmark.py:483-517 pass float64 samples and evaluation points); this replaces the pageable
 *      cudaMemcpy + cast kernel behind `to_device_points`: `threads` host threads (1..8) narrow their chunks into their own
 *      pinned staging buffers and queue the DMA of one buffer while filling the other. Ordered after the work already
 *      queued on `stream`; when the call returns h_src has been read completely and the data is in HBM. */
int aisb_upload_as_f32(const void* h_src, int32_t src_is_f64, int64_t count, float* d_dst, int32_t threads,
                       void* stream);
/*  aisb_download_f32: d_src (DEVICE, float32 [count]) -> h_dst (HOST, pageable or pinned; float64 if dst_is_f64 else
 *      float32). The reference's API returns float64 numpy arrays (`importance_sample` -> samples, weights:
 *      sampling_methods/dm_ais.py:86, base.py:21-35); this replaces device-side widening + a pageable cudaMemcpy into
 *      untouched memory: `threads` host threads (1..8) pull chunks through their pinned staging buffers and widen them
 *      into the destination, touching its pages in parallel. Ordered after the work already queued on `stream`;
 *      synchronous: when the call returns h_dst is complete. */
int aisb_download_f32(const float* d_src, int64_t count, void* h_dst, int32_t dst_is_f64, int32_t threads,
                      void* stream);
/*  aisb_sample_mixture_f32: ancestral draws from a full-covariance Gaussian mixture, given the component of every draw:
 *      out[i, :] = means[idx[i], :] + chol[idx[i]] z_i,  z_i ~ N(0, I)   (chol: lower-triangular factors, row-major
 *      [K, D, D]). Replaces CMixtureModel.sample (CMixtureModel.py:79-89) over CMultivariateNormal.sample
 *      (CMultivariateNormal.py:64-65) once the component indices have been drawn (m_pmc.py:86). */
int aisb_sample_mixture_f32(const float* d_means, const float* d_chol, int64_t K, int32_t D, const int64_t* d_idx,
                            int64_t n, uint64_t seed, uint64_t offset, float* d_out, void* stream);
/*  aisb_sample_iso_ctr_f32: aisb_sample_iso_f32 with the Philox state kept in DEVICE memory: d_state[2] =
 *      {position in the stream, key}. The kernel reads both; a one-thread kernel behind it advances the position. A
 *      captured CUDA graph of a sampler iteration therefore draws fresh numbers on every replay and follows a re-seed
 *      (the owner rewrites d_state) without being captured again. */
int aisb_sample_iso_ctr_f32(const float* d_means, int64_t K, int32_t D, int64_t per, double var, uint64_t* d_state,
                            float* d_out, void* stream);
int aisb_sample_iso_f32(const float* d_means, int64_t K, int32_t D, int64_t per, double var, uint64_t seed,
                        uint64_t offset, float* d_out, void* stream);
int aisb_uniform_box_f32(const float* d_lo, const float* d_hi, int32_t D, int64_t n, uint64_t seed, uint64_t offset,
                         float* d_out, void* stream);

/*
 * Tensor-core (tcgen05 / TMEM) cross-term path for ISO_SHARED mixtures WITHOUT offsets (equal-weight proposals, KDE),
 * 5 <= D <= 32:  log q(x_i) via  -||u_i||^2 + LSE_j( 2 u_i . v_j - ||v_j||^2 )  with the bracket computed as a TF32
 * GEMM on the 5th-generation tensor cores (BASELINE.json configs[4], "tensor-core cross-term path").
 * The tensor-core value (absolute error up to ~0.3 in the exponent) is only a SCREEN: every pair that comes within
 * 2^-30 of the point's reference term is re-evaluated exactly with the centred FP32 form, so the result meets the same
 * 1e-5 bar as aisb_mixture_logpdf_f32. Each point widens that window by twice its own worst-case
 * screen error (from the operand magnitudes), so data far from the origin costs time, not accuracy. The reference term comes from d_hint[i] = index of a
 * component expected to dominate log q(x_i) (DM-AIS: the proposal x_i was drawn from); a poor hint costs time (more exact
 * re-evaluations) or, if it is more than ~2^100 below the true maximum, accuracy -- callers without a natural hint use
 * aisb_mixture_logpdf_f32.
 *   aisb_tc_image_floats: floats needed for the pre-arranged B operand ("image") of a K-component, D-dim mixture.
 *   aisb_tc_pack_f32:     build the image from the packed ISO_SHARED rows (once per parameter update).
 *   aisb_tc_logq_f32:     d_out_logq[n]; d_hint[n] int32; *d_err (int32, zeroed by the caller) is incremented if the
 *                         kernel aborted on a barrier time-out.
 */
int64_t aisb_tc_image_floats(int64_t K, int32_t D);
int aisb_tc_pack_f32(const aisb_packed_mixture* mix, float* d_image, void* stream);
int aisb_tc_logq_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix, const int32_t* d_hint,
                     const float* d_image, float* d_out_logq, int32_t* d_err, void* stream);

/*
 * aisb_dm_weights_f32 with the proposal density on the tensor cores (for DM-AIS / LAIS at 9 <= D <= 32, where every
 * sample has a natural hint: the proposal it was drawn from, dm_ais.py:64-69). Same outputs as aisb_dm_weights_f32;
 * d_tc_image from aisb_tc_pack_f32 for `proposals`, d_hint[n] int32. If the tensor-core kernel aborted,
 * d_out_partials[3] (n_valid) is set to -1.
 */
int aisb_dm_weights_hinted_f32(const float* d_x, int64_t n, const aisb_packed_mixture* target, const float* d_logp_in,
                               const aisb_packed_mixture* proposals, const float* d_tc_image, const int32_t* d_hint,
                               float* d_out_logw, float* d_out_logp, float* d_out_logq, double* d_out_partials,
                               void* d_workspace, size_t workspace_bytes, void* stream);

/* Micro-benchmarks used to measure the FP32-FMA and MUFU.EX2 peaks on the box (bench.py roofline denominators).
 * Each runs `iters` dependent-chain iterations in every thread of a full-chip grid; *ops = FMAs (packed FFMA2 chains,
 * 2 flop each) resp. EX2 evaluations executed. */
int aisb_peak_fma_f32(int64_t iters, float* d_sink /* >= 4 zeroed floats */, double* ops, void* stream);
int aisb_peak_ex2_f32(int64_t iters, float* d_sink, double* ops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIS_B200_H */
