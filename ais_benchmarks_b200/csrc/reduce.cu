// Device-side weight-normaliser / ESS and KLD reductions + their C-ABI entry points.
#include "reduce.cuh"

namespace aisb {

__global__ void weight_partials_merge_kernel(const double* __restrict__ bp, int64_t nblocks,
                                             double* __restrict__ out4, const int32_t* __restrict__ abort) {
  __shared__ double sh[3][8];
  __shared__ double shM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double M = kNegBig;
  for (int64_t b = threadIdx.x; b < nblocks; b += blockDim.x) M = fmax(M, bp[4 * b]);
  M = warp_max(M);
  if (lane == 0) sh[0][warp] = M;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = sh[0][0];
    for (int w = 1; w < 8; ++w) v = fmax(v, sh[0][w]);
    shM = v;
  }
  __syncthreads();
  M = shM;
  double s1 = 0.0, s2 = 0.0, cnt = 0.0;
  for (int64_t b = threadIdx.x; b < nblocks; b += blockDim.x) {
    const double f = exp(bp[4 * b] - M);
    s1 += bp[4 * b + 1] * f;
    s2 += bp[4 * b + 2] * f * f;
    // a record poisoned by an aborted tensor-core kernel (n_valid = -1) poisons the merged record
    cnt += bp[4 * b + 3] < 0.0 ? -1.0e300 : bp[4 * b + 3];
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  cnt = warp_sum(cnt);
  __syncthreads();
  if (lane == 0) {
    sh[0][warp] = s1;
    sh[1][warp] = s2;
    sh[2][warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += sh[0][w];
      b2 += sh[1][w];
      c += sh[2][w];
    }
    out4[0] = M;
    out4[1] = a;
    out4[2] = b2;
    out4[3] = (c < 0.0 || (abort != nullptr && *abort != 0)) ? -1.0 : c;
  }
}

__global__ void __launch_bounds__(256) logw_partials_kernel(const float* __restrict__ logw, int64_t n,
                                                            double* __restrict__ block_partials) {
  __shared__ float red[256 / 32 * 4];
  constexpr int E = 8;
  float lw[E];
  bool ok[E];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * (256 * E);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int64_t p = base + e * 256 + threadIdx.x;
    const bool in = p < n;
    lw[e] = in ? logw[p] : 0.f;
    ok[e] = in && (lw[e] == lw[e]) && (lw[e] != INFINITY);
  }
  block_weight_partials<E, 256>(lw, ok, red, block_partials + 4 * static_cast<int64_t>(blockIdx.x));
}

__global__ void normalize_weights_kernel(const float* __restrict__ logw, int64_t n, float M, float logS,
                                         float* __restrict__ out) {
  const int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (p >= n) return;
  // (lw - M) is exact-ish near the maximum where the weights are largest; logS is O(log n)
  out[p] = expf((logw[p] - M) - logS);
}

// Same, with the record(s) in device memory: thread 0 folds the nrec records (one per rank after an all-gather, or one
// per streamed chunk; log-sum-exp rescale rule, float64) and derives the two constants; block 0 also writes the summary
// statistics {log sum w, ESS, n_finite}.
__global__ void __launch_bounds__(256) normalize_weights_dev_kernel(const float* __restrict__ logw, int64_t n,
                                                                    const double* __restrict__ rec, int nrec,
                                                                    float* __restrict__ out,
                                                                    double* __restrict__ stats, int vec_ok) {
  __shared__ float shc[2];
  if (threadIdx.x == 0) {
    double M = rec[0];
    for (int r = 1; r < nrec; ++r) M = fmax(M, rec[4 * r]);
    double S1 = 0.0, S2 = 0.0, cnt = 0.0;
    bool poisoned = false;
    for (int r = 0; r < nrec; ++r) {
      const double f = exp(rec[4 * r] - M);
      S1 += rec[4 * r + 1] * f;
      S2 += rec[4 * r + 2] * f * f;
      poisoned |= rec[4 * r + 3] < 0.0;
      cnt += rec[4 * r + 3];
    }
    const float Mf = static_cast<float>(M);
    shc[0] = Mf;
    shc[1] = static_cast<float>(log(S1) + (M - static_cast<double>(Mf)));
    if (blockIdx.x == 0 && stats) {
      stats[0] = M + log(S1);
      stats[1] = S2 > 0.0 ? S1 * S1 / S2 : 0.0;
      stats[2] = poisoned ? -1.0 : cnt;
    }
  }
  __syncthreads();
  const float M = shc[0], logS = shc[1];
  const int64_t base = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) * 4;
  if (vec_ok && base + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(logw + base);
    float4 r;
    r.x = expf((v.x - M) - logS);
    r.y = expf((v.y - M) - logS);
    r.z = expf((v.z - M) - logS);
    r.w = expf((v.w - M) - logS);
    *reinterpret_cast<float4*>(out + base) = r;
  } else {
    for (int64_t p = base; p < n && p < base + 4; ++p) out[p] = expf((logw[p] - M) - logS);
  }
}

// ---------------------------------------------------------------------------------------------
// Cross-rank exchange of the 4-double weight records over NVLink peer memory (one CTA, no NCCL launch).
// Every rank owns a symmetric buffer laid out as
//     double rec[2][world][4];  unsigned long long flag[2][world];        (index 0/1 = epoch parity)
// and sees its peers' buffers through peer_bufs[r]. Thread r of the CTA stores this rank's record into slot
// [parity][rank] of peer r, fences, and publishes the epoch number in the matching flag; then it waits for the flag
// of slot [parity][r] in the LOCAL buffer and copies that record out. Two parities make the slots safe to re-use: a
// peer can only reach epoch e+2 after this rank has published epoch e+1, i.e. after everything it enqueued for epoch e
// (including the kernels that read out[]) has completed. The wait is bounded (~17 s); a time-out poisons the record.
// ---------------------------------------------------------------------------------------------
__global__ void weight_records_exchange_kernel(const double* __restrict__ local_rec, double* const* __restrict__ peer_bufs,
                                               int rank, int world, unsigned long long epoch,
                                               double* __restrict__ out) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const int par = static_cast<int>(epoch & 1ull);
  {
    double* pb = peer_bufs[r];
    volatile double* slot = pb + (static_cast<size_t>(par) * world + rank) * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) slot[q] = local_rec[q];
    __threadfence_system();
    volatile unsigned long long* flag =
        reinterpret_cast<unsigned long long*>(pb + static_cast<size_t>(2) * world * 4) + par * world + rank;
    *flag = epoch;
  }
  double* mine = peer_bufs[rank];
  volatile unsigned long long* flag =
      reinterpret_cast<unsigned long long*>(mine + static_cast<size_t>(2) * world * 4) + par * world + r;
  const long long t0 = clock64();
  bool ok = true;
  while (*flag < epoch) {
    if (clock64() - t0 > (1ll << 35)) {  // ~17 s at 2 GHz: a peer never arrived (generous: ranks may be skewed)
      ok = false;
      break;
    }
  }
  __threadfence_system();
  const volatile double* src = mine + (static_cast<size_t>(par) * world + r) * 4;
#pragma unroll
  for (int q = 0; q < 4; ++q) out[4 * r + q] = src[q];
  if (!ok) out[4 * r + 3] = -1.0;
}

// KLD block record (8 doubles): { m_p, S_p, A, m_q, S_q, n_inlier, n_posinf_q, n_total }
// Inliers: lp > floor and lq > floor (floor = -inf for the KLD; the float64 underflow threshold of exp() for the
// linear-domain JSD of the reference, divergences.py:170-173).
__global__ void __launch_bounds__(256) kld_partials_kernel(const float* __restrict__ lp, const float* __restrict__ lq,
                                                           int64_t n, float floor_,
                                                           double* __restrict__ block_partials) {
  constexpr int E = 8, NW = 8;
  __shared__ float redf[2 * NW];
  __shared__ double redd[6 * NW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float p[E], q[E];
  bool in[E];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * (256 * E);
  float mp = kNegBig, mq = kNegBig;
  float ntot = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int64_t i = base + e * 256 + threadIdx.x;
    const bool valid = i < n;
    p[e] = valid ? lp[i] : -INFINITY;
    q[e] = valid ? lq[i] : -INFINITY;
    // reference: inliers = (lp > -inf) & (lq > -inf)  (NaN compares false)  divergences.py:138-139
    in[e] = valid && (p[e] > floor_) && (q[e] > floor_);
    if (valid) ntot += 1.f;
    if (in[e]) {
      if (p[e] < INFINITY) mp = fmaxf(mp, p[e]);
      if (q[e] < INFINITY) mq = fmaxf(mq, q[e]);
    }
  }
  mp = warp_max(mp);
  mq = warp_max(mq);
  if (lane == 0) {
    redf[warp] = mp;
    redf[NW + warp] = mq;
  }
  __syncthreads();
  mp = redf[0];
  mq = redf[NW];
#pragma unroll
  for (int w = 1; w < NW; ++w) {
    mp = fmaxf(mp, redf[w]);
    mq = fmaxf(mq, redf[NW + w]);
  }
  double sp = 0.0, a = 0.0, sq = 0.0, nin = 0.0, ninf = 0.0;
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (in[e]) {
      nin += 1.0;
      if (q[e] == INFINITY) {
        ninf += 1.0;
      } else {
        const float up = expf(p[e] - mp);  // +inf lp -> inf -> NaN downstream, as in the reference
        sp += up;
        a += static_cast<double>(up) * (static_cast<double>(p[e]) - static_cast<double>(q[e]));
        sq += expf(q[e] - mq);
      }
    }
  sp = warp_sum(sp);
  a = warp_sum(a);
  sq = warp_sum(sq);
  nin = warp_sum(nin);
  ninf = warp_sum(ninf);
  double nt = warp_sum(static_cast<double>(ntot));
  if (lane == 0) {
    redd[warp] = sp;
    redd[NW + warp] = a;
    redd[2 * NW + warp] = sq;
    redd[3 * NW + warp] = nin;
    redd[4 * NW + warp] = ninf;
    redd[5 * NW + warp] = nt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double r[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 6; ++k)
      for (int w = 0; w < NW; ++w) r[k] += redd[k * NW + w];
    double* o = block_partials + 8 * static_cast<int64_t>(blockIdx.x);
    o[0] = mp;
    o[1] = r[0];
    o[2] = r[1];
    o[3] = mq;
    o[4] = r[2];
    o[5] = r[3];
    o[6] = r[4];
    o[7] = r[5];
  }
}

__global__ void kld_partials_merge_kernel(const double* __restrict__ bp, int64_t nblocks, double* __restrict__ out8) {
  __shared__ double sh[6][8];
  __shared__ double shM[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double mp = kNegBig, mq = kNegBig;
  for (int64_t b = threadIdx.x; b < nblocks; b += blockDim.x) {
    mp = fmax(mp, bp[8 * b]);
    mq = fmax(mq, bp[8 * b + 3]);
  }
  mp = warp_max(mp);
  mq = warp_max(mq);
  if (lane == 0) {
    sh[0][warp] = mp;
    sh[1][warp] = mq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = sh[0][0], b = sh[1][0];
    for (int w = 1; w < 8; ++w) {
      a = fmax(a, sh[0][w]);
      b = fmax(b, sh[1][w]);
    }
    shM[0] = a;
    shM[1] = b;
  }
  __syncthreads();
  mp = shM[0];
  mq = shM[1];
  double r[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t b = threadIdx.x; b < nblocks; b += blockDim.x) {
    const double fp = exp(bp[8 * b] - mp), fq = exp(bp[8 * b + 3] - mq);
    r[0] += bp[8 * b + 1] * fp;
    r[1] += bp[8 * b + 2] * fp;
    r[2] += bp[8 * b + 4] * fq;
    r[3] += bp[8 * b + 5];
    r[4] += bp[8 * b + 6];
    r[5] += bp[8 * b + 7];
  }
  __syncthreads();
  for (int k = 0; k < 6; ++k) {
    r[k] = warp_sum(r[k]);
    if (lane == 0) sh[k][warp] = r[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 6; ++k)
      for (int w = 0; w < 8; ++w) t[k] += sh[k][w];
    out8[0] = mp;
    out8[1] = t[0];
    out8[2] = t[1];
    out8[3] = mq;
    out8[4] = t[2];
    out8[5] = t[3];
    out8[6] = t[4];
    out8[7] = t[5];
  }
}

// JSD terms (divergences.py:183-195) from log densities and the inlier normalisers Lp, Lq of a kld_partials pass with
// the same floor:  a = lp - Lp, b = lq - Lq, log m = logaddexp(a, b) - log 2,
// term = 0.5 e^a (a - log m) + 0.5 e^b (b - log m); NaN terms are dropped and counted (:192).
// out2 += { sum of terms (nats), dropped terms }   (float64 atomics; HBM-bound stream, 8 B per point)
__global__ void __launch_bounds__(256) jsd_terms_kernel(const float* __restrict__ lp, const float* __restrict__ lq,
                                                        int64_t n, float floor_, double Lp, double Lq,
                                                        double* __restrict__ out2) {
  __shared__ double red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sum = 0.0, dropped = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float p = lp[i], q = lq[i];
    if (!(p > floor_) || !(q > floor_)) continue;
    const double a = static_cast<double>(p) - Lp, b = static_cast<double>(q) - Lq;
    const double hi = fmax(a, b), lo = fmin(a, b);
    const double lm = hi + log1p(exp(lo - hi)) - 0.69314718055994530942;
    const double t = 0.5 * exp(a) * (a - lm) + 0.5 * exp(b) * (b - lm);
    if (t == t) {
      sum += t;
    } else {
      dropped += 1.0;
    }
  }
  sum = warp_sum(sum);
  dropped = warp_sum(dropped);
  if (lane == 0) {
    red[0][warp] = sum;
    red[1][warp] = dropped;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(out2 + threadIdx.x, t);
  }
}

// Self-normalised expectation of a point set (statistics.py:39-45):  out[d] += sum_i x[i][d] exp(logw[i] - log_norm).
// Elements are read flat (coalesced); the per-thread stride is a multiple of D so that a thread always sees the same
// coordinate and keeps ONE float64 accumulator. 4 (D + 1) / D bytes per element, HBM-bound.
__global__ void __launch_bounds__(256) weighted_mean_kernel(const float* __restrict__ x, int64_t n, int D,
                                                            const float* __restrict__ logw, double log_norm,
                                                            double* __restrict__ out) {
  __shared__ double acc[AISB_MAX_DIMS];
  const int T = (256 / D) * D;  // active threads: a multiple of D
  if (threadIdx.x < D) acc[threadIdx.x] = 0.0;
  __syncthreads();
  if (threadIdx.x < T) {
    const int64_t total = n * D;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * T;
    double s = 0.0;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * T + threadIdx.x; e < total; e += stride) {
      const double w = exp(static_cast<double>(__ldg(logw + e / D)) - log_norm);
      s += w * static_cast<double>(__ldg(x + e));
    }
    atomicAdd(&acc[threadIdx.x % D], s);
  }
  __syncthreads();
  if (threadIdx.x < D) atomicAdd(out + threadIdx.x, acc[threadIdx.x]);
}

static int64_t reduce_blocks(int64_t n) { return (n + 2047) / 2048; }

static int check_ws(void* d_ws, size_t ws_bytes, size_t need, double** out, const char* who) {
  char* base = static_cast<char*>(d_ws);
  const size_t mis = base ? (256 - reinterpret_cast<uintptr_t>(base) % 256) % 256 : 0;
  if (!base || need + mis > ws_bytes) {
    set_error("%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, need + mis);
    return AISB_ERR_WORKSPACE;
  }
  *out = reinterpret_cast<double*>(base + mis);
  return AISB_OK;
}

}  // namespace aisb

using namespace aisb;

extern "C" int aisb_logw_partials_f32(const float* d_logw, int64_t n, double* d_out_partials, void* d_workspace,
                                      size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!d_out_partials || n < 0 || (n > 0 && !d_logw)) {
    set_error("logw_partials: null pointer or negative n");
    return AISB_ERR_INVALID;
  }
  const int64_t nblocks = reduce_blocks(n);
  double* bp = nullptr;
  int rc = check_ws(d_workspace, workspace_bytes, sizeof(double) * 4 * (nblocks + 1), &bp, "logw_partials");
  if (rc) return rc;
  if (nblocks > 0) {
    logw_partials_kernel<<<static_cast<unsigned>(nblocks), 256, 0, st>>>(d_logw, n, bp);
    AISB_LAUNCH_CHECK("logw_partials_kernel");
  }
  weight_partials_merge_kernel<<<1, 256, 0, st>>>(bp, nblocks, d_out_partials, nullptr);
  AISB_LAUNCH_CHECK("weight_partials_merge_kernel");
  return AISB_OK;
}

extern "C" int aisb_normalize_weights_f32(const float* d_logw, int64_t n, const double* h_partials, float* d_out_w,
                                          void* stream) {
  if (n < 0 || !h_partials || (n > 0 && (!d_logw || !d_out_w))) {
    set_error("normalize_weights: null pointer or negative n");
    return AISB_ERR_INVALID;
  }
  if (n == 0) return AISB_OK;
  const float M = static_cast<float>(h_partials[0]);
  const float logS = static_cast<float>(log(h_partials[1]) + (h_partials[0] - static_cast<double>(M)));
  normalize_weights_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_logw, n, M, logS, d_out_w);
  AISB_LAUNCH_CHECK("normalize_weights_kernel");
  return AISB_OK;
}

extern "C" int aisb_weight_records_merge(const double* d_records, int64_t n_records, double* d_out_partials,
                                         void* stream) {
  if (!d_records || !d_out_partials || n_records < 1) {
    set_error("weight_records_merge: null pointer or n_records < 1");
    return AISB_ERR_INVALID;
  }
  weight_partials_merge_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_records, n_records, d_out_partials,
                                                                                 nullptr);
  AISB_LAUNCH_CHECK("weight_partials_merge_kernel");
  return AISB_OK;
}

extern "C" int64_t aisb_weight_exchange_doubles(int32_t world) {
  return world < 1 ? 0 : static_cast<int64_t>(2) * world * 4 + static_cast<int64_t>(2) * world;
}

extern "C" int aisb_weight_records_exchange(const double* d_local_record, void* const* d_peer_bufs, int32_t rank,
                                            int32_t world, uint64_t epoch, double* d_out_records, void* stream) {
  if (!d_local_record || !d_peer_bufs || !d_out_records || world < 1 || world > 256 || rank < 0 || rank >= world ||
      epoch == 0) {
    set_error("weight_records_exchange: null pointer, bad rank/world or epoch 0");
    return AISB_ERR_INVALID;
  }
  weight_records_exchange_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_local_record, reinterpret_cast<double* const*>(d_peer_bufs), rank, world,
      static_cast<unsigned long long>(epoch), d_out_records);
  AISB_LAUNCH_CHECK("weight_records_exchange_kernel");
  return AISB_OK;
}

extern "C" int aisb_kld_records_merge(const double* d_records, int64_t n_records, double* d_out_partials,
                                      void* stream) {
  if (!d_records || !d_out_partials || n_records < 1) {
    set_error("kld_records_merge: null pointer or n_records < 1");
    return AISB_ERR_INVALID;
  }
  kld_partials_merge_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_records, n_records, d_out_partials);
  AISB_LAUNCH_CHECK("kld_partials_merge_kernel");
  return AISB_OK;
}

extern "C" int aisb_normalize_weights_dev_f32(const float* d_logw, int64_t n, const double* d_records,
                                              int32_t n_records, float* d_out_w, double* d_out_stats, void* stream) {
  if (n < 0 || !d_records || n_records < 1 || n_records > 4096 || (n > 0 && (!d_logw || !d_out_w))) {
    set_error("normalize_weights_dev: null pointer, negative n or n_records outside 1..4096");
    return AISB_ERR_INVALID;
  }
  // float4 accesses when both vectors are 16-byte aligned (always true for whole tensors), scalar ones for odd views
  const int vec_ok = ((reinterpret_cast<uintptr_t>(d_logw) | reinterpret_cast<uintptr_t>(d_out_w)) & 15) == 0;
  // n == 0 still launches one block so that the statistics are written
  const int64_t blocks = n > 0 ? (n + 1023) / 1024 : 1;
  normalize_weights_dev_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_logw, n, d_records, n_records, d_out_w, d_out_stats, vec_ok);
  AISB_LAUNCH_CHECK("normalize_weights_dev_kernel");
  return AISB_OK;
}

extern "C" int aisb_kld_partials_f32(const float* d_lp, const float* d_lq, int64_t n, double* d_out_partials,
                                     void* d_workspace, size_t workspace_bytes, void* stream) {
  return aisb_pair_partials_f32(d_lp, d_lq, n, -INFINITY, d_out_partials, d_workspace, workspace_bytes, stream);
}

extern "C" int aisb_pair_partials_f32(const float* d_lp, const float* d_lq, int64_t n, float floor_,
                                      double* d_out_partials, void* d_workspace, size_t workspace_bytes,
                                      void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!d_out_partials || n < 0 || (n > 0 && (!d_lp || !d_lq)) || floor_ != floor_) {
    set_error("pair_partials: null pointer, negative n or NaN floor");
    return AISB_ERR_INVALID;
  }
  const int64_t nblocks = reduce_blocks(n);
  double* bp = nullptr;
  int rc = check_ws(d_workspace, workspace_bytes, sizeof(double) * 8 * (nblocks + 1), &bp, "kld_partials");
  if (rc) return rc;
  if (nblocks > 0) {
    kld_partials_kernel<<<static_cast<unsigned>(nblocks), 256, 0, st>>>(d_lp, d_lq, n, floor_, bp);
    AISB_LAUNCH_CHECK("kld_partials_kernel");
  }
  kld_partials_merge_kernel<<<1, 256, 0, st>>>(bp, nblocks, d_out_partials);
  AISB_LAUNCH_CHECK("kld_partials_merge_kernel");
  return AISB_OK;
}

extern "C" int aisb_jsd_terms_f32(const float* d_lp, const float* d_lq, int64_t n, float floor_, double log_norm_p,
                                  double log_norm_q, double* d_out2, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!d_out2 || n < 0 || (n > 0 && (!d_lp || !d_lq)) || floor_ != floor_) {
    set_error("jsd_terms: null pointer, negative n or NaN floor");
    return AISB_ERR_INVALID;
  }
  AISB_CUDA_CHECK(cudaMemsetAsync(d_out2, 0, 2 * sizeof(double), st));
  if (n == 0) return AISB_OK;
  const int64_t want = (n + 2047) / 2048;  // ~8 points per thread
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  jsd_terms_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, st>>>(d_lp, d_lq, n, floor_, log_norm_p,
                                                                                  log_norm_q, d_out2);
  AISB_LAUNCH_CHECK("jsd_terms_kernel");
  return AISB_OK;
}

extern "C" int aisb_weighted_mean_f32(const float* d_x, int64_t n, int32_t D, const float* d_logw, double log_norm,
                                      double* d_out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!d_out || n < 0 || D < 1 || D > AISB_MAX_DIMS || (n > 0 && (!d_x || !d_logw))) {
    set_error("weighted_mean: null pointer, negative n or D=%d outside 1..%d", D, AISB_MAX_DIMS);
    return AISB_ERR_INVALID;
  }
  AISB_CUDA_CHECK(cudaMemsetAsync(d_out, 0, D * sizeof(double), st));
  if (n == 0) return AISB_OK;
  const int T = (256 / D) * D;
  const int64_t want = (n * D + 8 * T - 1) / (8 * T);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  weighted_mean_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, st>>>(d_x, n, D, d_logw, log_norm,
                                                                                      d_out);
  AISB_LAUNCH_CHECK("weighted_mean_kernel");
  return AISB_OK;
}
