// Mixture log-density and fused DM-weight kernels + their C-ABI entry points.
#include <stdlib.h>

#include "pairwise.cuh"
#include "reduce.cuh"

#ifndef AISB_MIN_CTAS_SMALL
#define AISB_MIN_CTAS_SMALL 5  // resident CTAs per SM asked of the compiler for the DP <= 8 pairwise kernels
#endif

namespace aisb {

struct PassArgs {
  const float* rows;
  const float* offs;
  int64_t KP;
  float a_iso;
  float uniform_offset;
  float skip_margin;  // 28 + log2(KP): groups this far (log2) below a warp's running maxima are not folded (pairwise.cuh)
};

// ---------------------------------------------------------------------------------------------
// logp[i] = log sum_k w_k N(x_i; mu_k, Sigma_k)
// grid.x = ntiles * nsplit; CTA b handles point tile (b % ntiles) against component range
// split (b / ntiles). With nsplit > 1 the (m, S) partials go to the workspace and
// lse_merge_kernel finishes.
// ---------------------------------------------------------------------------------------------
// SKIP: the warp-uniform fold skip of pairwise.cuh (points and components spatially ordered by the caller); that
// variant needs ~120 registers at DP <= 8, hence 4 resident CTAs instead of 5.
// (DIAG at DP = 8 holds a and b rows: at 5 CTAs per SM it spilled 296 B, so it is compiled for 4 as well.)
template <int KIND, int DP, bool HAS_C, bool SKIP = false>
__global__ void __launch_bounds__(kThreads,
                                  (KIND != AISB_COV_FULL && DP <= 8)
                                      ? ((SKIP || (KIND == AISB_COV_DIAG && DP == 8)) ? 4 : AISB_MIN_CTAS_SMALL)
                                      : 1)
    logpdf_kernel(const float* __restrict__ x, int64_t n, int D, int vec_ok, PassArgs mix, int nsplit,
                  int64_t comps_per_split, int64_t ntiles, float* __restrict__ out, float* __restrict__ ws_m,
                  float* __restrict__ ws_S) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[kStages];
  __shared__ OuterAcc outer;
  constexpr int E = points_per_thread(DP);

  const int64_t tile = blockIdx.x % ntiles;
  const int64_t split = blockIdx.x / ntiles;
  ring_init(bars);

  float xr[E][DP];
  int64_t pidx[E];
  load_points<DP, E>(x, n, D, vec_ok, tile * (kThreads * E), xr, pidx);

  float m[E], S[E];
  const int64_t k_begin = split * comps_per_split;
  const int64_t k_end = k_begin + comps_per_split < mix.KP ? k_begin + comps_per_split : mix.KP;
  uint32_t it = 0;
  mixture_pass<KIND, DP, E, HAS_C, SKIP>(mix.rows, mix.offs, k_begin, k_end, mix.a_iso, smem, bars, outer, it, xr, m, S,
                                         mix.skip_margin);

  if (nsplit == 1) {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (pidx[e] < n) out[pidx[e]] = lse_finish(m[e], S[e]) + (HAS_C ? 0.f : mix.uniform_offset);
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (pidx[e] < n) {
        ws_m[split * n + pidx[e]] = m[e];
        ws_S[split * n + pidx[e]] = S[e];
      }
  }
}


// ---------------------------------------------------------------------------------------------
// Screen-and-refine variant for a spatially ordered KDE (pairwise.cuh, mixture_pass_screen).
// CTA = ONE warp that owns 32*E consecutive points and its own 2-stage ring of kScreenChunk components: how much of a
// chunk has to be refined differs wildly between neighbouring warps (a chunk of the Z-order curve is either next to a
// warp's points or irrelevant to all of them), so warps that shared a ring and its per-chunk barrier spent a third of
// their time waiting for each other (ncu: stall_barrier the top reason, profiles/r2_ncu_kde_screen_v1.txt). Up to 16 such
// CTAs are resident per SM. grid.x = ntiles * nsplit; CTA b sweeps every nsplit-th chunk of the components, starting at
// chunk b / ntiles (interleaved split), for point tile b % ntiles. COUNT adds the (refined, screened) unit counters to
// stats[0..1].
// ---------------------------------------------------------------------------------------------
constexpr int kScreenThreads = 32;
constexpr int kScreenChunk = 128;
__host__ __device__ constexpr size_t screen_smem_bytes(int DP) {
  return size_t(kStages) * (kScreenChunk * DP + kScreenChunk) * sizeof(float);
}

template <int DP, bool COUNT>
__global__ void __launch_bounds__(kScreenThreads, DP <= 8 ? 16 : 8)
    logpdf_screen_kernel(const float* __restrict__ x, int64_t n, PassArgs mix, const float* __restrict__ hs, float marg,
                         int nsplit, int64_t ntiles, float* __restrict__ out, float* __restrict__ ws_m,
                         float* __restrict__ ws_S, unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[kStages];
  __shared__ OuterAccT<kScreenThreads> outer;
  constexpr int E = points_per_thread(DP);

  const int64_t tile = blockIdx.x % ntiles;
  const int64_t split = blockIdx.x / ntiles;
  ring_init(bars);

  float xr[E][DP];
  int64_t pidx[E];
  load_points_warp<DP, E>(x, n, tile * (kScreenThreads * E), xr, pidx);

  float m[E], S[E];
  uint32_t it = 0;
  unsigned nscr = 0, nref = 0;
  mixture_pass_screen<DP, E, kScreenChunk, COUNT>(mix.rows, hs, mix.KP, split * kScreenChunk, nsplit, mix.a_iso, marg,
                                                  smem, bars, outer, it, xr, m, S, &nscr, &nref);
  if constexpr (COUNT) {
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(stats + 0, static_cast<unsigned long long>(nref));
      atomicAdd(stats + 1, static_cast<unsigned long long>(nscr));
    }
  }
  if (nsplit == 1) {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (pidx[e] < n) out[pidx[e]] = lse_finish(m[e], S[e]) + mix.uniform_offset;
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (pidx[e] < n) {
        ws_m[split * n + pidx[e]] = m[e];
        ws_S[split * n + pidx[e]] = S[e];
      }
  }
}


// ---------------------------------------------------------------------------------------------
// Fast variant for a spatially ordered KDE (pairwise.cuh, mixture_pass_fast): recentred expanded form, single-warp CTAs
// with private rings like logpdf_screen_kernel. Every point's float32 error estimate
//     eps = kFastErrC * 2^-24 * (||a x'||^2 + |t_max| + 1)      (natural-log units; model measured by
//                                                                 scripts/kde_fast_error_model.py, 1.7x margin)
// is compared with kFastTol * max(1, |log p|); points that fail are appended to `list` (their count in *list_n) and
// re-evaluated by logpdf_subset_kernel with the centred form. With nsplit > 1 the test runs in fast_merge_kernel.
// ---------------------------------------------------------------------------------------------
constexpr float kFastErrC = 4.0f;
constexpr float kFastTol = 3.0e-6f;

__device__ __forceinline__ bool fast_needs_exact(float u2, float t_max, float logp) {
  const float eps = kFastErrC * 5.9604645e-8f * (u2 + fabsf(t_max) + 1.f);
  return eps > kFastTol * fmaxf(1.f, fabsf(logp));   // NaN / inf results compare false: nothing to refine there
}

#ifndef AISB_FAST_CTAS
// resident single-warp CTAs per SM the fast kernel is compiled for (DP <= 8). 16 -> 128 registers, no spills. Measured
// with 20 (96 registers, 72 B of spills, 21 resident): 368.1 vs 366.0 ms at configs[3] -- more warps do not help, the
// four warps of a scheduler already queue for its FMA pipe (math_pipe_throttle is the top stall).
#define AISB_FAST_CTAS 16
#endif
template <int DP, bool COUNT>
__global__ void __launch_bounds__(kScreenThreads, DP <= 8 ? AISB_FAST_CTAS : 8)
    logpdf_fast_kernel(const float* __restrict__ x, int64_t n, PassArgs mix, float marg, int nsplit, int64_t ntiles,
                       float* __restrict__ out, float* __restrict__ ws_m, float* __restrict__ ws_S,
                       float* __restrict__ ws_u2, int64_t* __restrict__ list, unsigned long long* __restrict__ list_n,
                       unsigned long long* __restrict__ stats, int dbg, uint32_t* __restrict__ gmin) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[kStages];
  __shared__ OuterAccT<kScreenThreads> outer;
  __shared__ __align__(16) float sh_h[kScreenChunk];
  constexpr int E = points_per_thread(DP);

  const int64_t tile = blockIdx.x % ntiles;
  const int64_t split = blockIdx.x / ntiles;
  ring_init(bars);

  float xr[E][DP];
  int64_t pidx[E];
  load_points_warp<DP, E>(x, n, tile * (kScreenThreads * E), xr, pidx);

  float m[E], S[E], u2[E];
  uint32_t it = 0;
  unsigned nscr = 0, nfold = 0;
  mixture_pass_fast<DP, E, kScreenChunk, COUNT>(mix.rows, mix.KP, split * kScreenChunk, nsplit, mix.a_iso, marg, smem,
                                                sh_h, bars, outer, it, xr, m, S, u2, &nscr, &nfold, dbg,
                                                nsplit > 1 ? gmin : nullptr, pidx, n);
  if constexpr (COUNT) {
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(stats + 0, static_cast<unsigned long long>(nfold));
      atomicAdd(stats + 1, static_cast<unsigned long long>(nscr));
    }
  }
  if (nsplit == 1) {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (pidx[e] < n) {
        const float lp = lse_finish(m[e], S[e]) + mix.uniform_offset;
        out[pidx[e]] = lp;
        if (fast_needs_exact(u2[e], m[e], lp)) list[atomicAdd(list_n, 1ull)] = pidx[e];
      }
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (pidx[e] < n) {
        ws_m[split * n + pidx[e]] = m[e];
        ws_S[split * n + pidx[e]] = S[e];
        if (split == 0) ws_u2[pidx[e]] = u2[e];
      }
  }
}

__global__ void fast_merge_kernel(const float* __restrict__ ws_m, const float* __restrict__ ws_S,
                                  const float* __restrict__ ws_u2, int nsplit, int64_t n, float add,
                                  float* __restrict__ out, int64_t* __restrict__ list,
                                  unsigned long long* __restrict__ list_n) {
  const int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (p >= n) return;
  float M = kNegBig;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, ws_m[s * n + p]);
  float S = 0.f;
  for (int s = 0; s < nsplit; ++s) S += ws_S[s * n + p] * ex2(ws_m[s * n + p] - M);
  const float lp = lse_finish(M, S) + add;
  out[p] = lp;
  if (fast_needs_exact(ws_u2[p], M, lp)) list[atomicAdd(list_n, 1ull)] = p;
}

// Exact (centred-form) re-evaluation of the listed points: out[list[k]] = log p(x[list[k]]) for k < *list_n. Same
// arithmetic as logpdf_kernel<ISO_SHARED, DP, false>. The length of the list is only known on the device (normally a few
// hundred to a few thousand points, far too few point tiles to fill the machine), so the launch has a FIXED grid of
// kSubsetGrid CTAs and every CTA derives the same plan from *list_n: the component range is split nsp ways so that
// tiles * nsp fills the grid (>= 256 components per split, at most kSubsetMaxSplit), work item w = (tile w % tiles,
// split w / tiles); partial (m, S) go to ws[split * tiles * tile_pts + k] and subset_merge_kernel finishes. With more
// tiles than CTAs (nearly everything listed) nsp = 1, the CTAs stride over the tiles and write the result directly.
// History: unsplit / 64 splits / 256 splits of the first 8192 entries took 58 / 2.7 / 1.1 ms for the 387 points a rank
// of an 8-GPU job lists (255 CTAs of 4 warps, 1-2 per SM) -- a fixed cost that did not shrink with the number of ranks.
constexpr int kSubsetGrid = 4096;
constexpr int kSubsetMaxSplit = 1024;
constexpr int64_t kSubsetWs = static_cast<int64_t>(kSubsetGrid) * kThreads * 4;   // floats per partial array

struct SubsetPlan {
  int64_t cnt, tiles, cps;   // list entries, point tiles, components per split
  int nsp;
};

__device__ __forceinline__ SubsetPlan subset_plan(const unsigned long long* __restrict__ list_n, int64_t n, int64_t KP,
                                                  int tile_pts, int grid) {
  SubsetPlan p;
  p.cnt = static_cast<int64_t>(*list_n) < n ? static_cast<int64_t>(*list_n) : n;
  p.tiles = (p.cnt + tile_pts - 1) / tile_pts;
  const int64_t units = KP / AISB_COMP_ALIGN;
  int64_t nsp = p.tiles > 0 ? grid / p.tiles : 1;
  if (nsp > units / 8) nsp = units / 8;
  if (nsp > kSubsetMaxSplit) nsp = kSubsetMaxSplit;
  if (nsp < 1) nsp = 1;
  p.cps = (units + nsp - 1) / nsp * AISB_COMP_ALIGN;
  p.nsp = static_cast<int>((KP + p.cps - 1) / p.cps);
  return p;
}

template <int DP>
__global__ void __launch_bounds__(kThreads, DP <= 8 ? 4 : 1)   // 5 CTAs per SM (96 registers) spilled 164 B at DP = 8
    logpdf_subset_kernel(const float* __restrict__ x, int64_t n, PassArgs mix, const int64_t* __restrict__ list,
                         const unsigned long long* __restrict__ list_n, float* __restrict__ ws_m,
                         float* __restrict__ ws_S, float* __restrict__ out) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[kStages];
  __shared__ OuterAcc outer;
  constexpr int E = points_per_thread(DP);
  const SubsetPlan plan = subset_plan(list_n, n, mix.KP, kThreads * E, static_cast<int>(gridDim.x));
  const int64_t nwork = plan.tiles * plan.nsp;
  if (static_cast<int64_t>(blockIdx.x) >= nwork) return;
  ring_init(bars);
  uint32_t it = 0;
  for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int64_t tile = w % plan.tiles, split = w / plan.tiles;
    const int64_t tile_base = tile * (kThreads * E);
    float xr[E][DP];
    int64_t kidx[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int64_t k = tile_base + e * kThreads + threadIdx.x;
      const int64_t p = list[k < plan.cnt ? k : plan.cnt - 1];
      kidx[e] = k < plan.cnt ? k : -1;
      if constexpr (DP >= 4) {
        const float4* r = reinterpret_cast<const float4*>(x + p * DP);
#pragma unroll
        for (int q = 0; q < DP / 4; ++q) {
          const float4 f = __ldg(r + q);
          xr[e][4 * q + 0] = f.x;
          xr[e][4 * q + 1] = f.y;
          xr[e][4 * q + 2] = f.z;
          xr[e][4 * q + 3] = f.w;
        }
      } else {
        const float2 f = __ldg(reinterpret_cast<const float2*>(x + p * 2));
        xr[e][0] = f.x;
        xr[e][1] = f.y;
      }
    }
    float m[E], S[E];
    const int64_t c_begin = split * plan.cps;
    const int64_t c_end = c_begin + plan.cps < mix.KP ? c_begin + plan.cps : mix.KP;
    mixture_pass<AISB_COV_ISO_SHARED, DP, E, false>(mix.rows, nullptr, c_begin, c_end, mix.a_iso, smem, bars, outer, it,
                                                    xr, m, S);
    const int64_t stride = plan.tiles * (kThreads * E);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (kidx[e] < 0) continue;
      if (plan.nsp == 1) {
        out[list[kidx[e]]] = lse_finish(m[e], S[e]) + mix.uniform_offset;
      } else {
        ws_m[split * stride + kidx[e]] = m[e];
        ws_S[split * stride + kidx[e]] = S[e];
      }
    }
  }
}

__global__ void subset_merge_kernel(const float* __restrict__ ws_m, const float* __restrict__ ws_S, int64_t n,
                                    int64_t KP, int tile_pts, int grid, const int64_t* __restrict__ list,
                                    const unsigned long long* __restrict__ list_n, float add, float* __restrict__ out) {
  const SubsetPlan plan = subset_plan(list_n, n, KP, tile_pts, grid);
  if (plan.nsp == 1) return;   // written directly
  const int64_t stride = plan.tiles * tile_pts;
  // one warp per list entry (up to kSubsetMaxSplit partials each): lanes stride over the splits
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); k < plan.cnt; k += nwarps) {
    float M = kNegBig;
    for (int s = lane; s < plan.nsp; s += 32) M = fmaxf(M, ws_m[s * stride + k]);
    M = warp_max(M);
    float S = 0.f;
    for (int s = lane; s < plan.nsp; s += 32) S += ws_S[s * stride + k] * ex2(ws_m[s * stride + k] - M);
    S = warp_sum(S);
    if (lane == 0) out[list[k]] = lse_finish(M, S) + add;
  }
}

// Interleaved component splits of the fast kernel. A CTA (one warp, 128 points) that swept all 1e6 components would run
// for ~100 ms, and CTAs differ 3-4x in how much they fold: with few, long CTAs the SMs drain unevenly at the end of the
// launch (measured: +3.7 ms on a 190 ms launch at 5 splits; at full size 375.2 / 373.0 / 373.3 / 374.8 ms with 3 / 5 / 8 / 16
// splits). The grid is therefore cut until it has kFastWaves waves of resident CTAs, capped at kFastMaxSplit (the
// splits exchange their running minima, so a split costs a partial (m, S, u2) record per point and little else).
// A rank of an 8-GPU job holds 125 000 points = 981 tiles: at 16 splits that is 6.6 waves of 7 ms CTAs and the SMs
// idle ~2 ms of a 48 ms launch at its end; at 48 splits the waves are 20 and a third as long.
constexpr int kFastMaxSplit = 64;
constexpr int kFastWaves = 20;
constexpr int64_t kFastCtaTarget = 148 * 16 * kFastWaves;   // B200: 148 SMs x 16 resident single-warp CTAs

// Most splits a launch over n points asks for (sizes the scratch buffer; independent of the mixture)
static int64_t fast_split_cap(int64_t n) {
  const int64_t ntiles = (n + kScreenThreads * 4 - 1) / (kScreenThreads * 4);
  int64_t want = ntiles > 0 ? (kFastCtaTarget + ntiles - 1) / ntiles : 1;
  if (want > kFastMaxSplit) want = kFastMaxSplit;
  return want < 1 ? 1 : want;
}

// h_j = ||b_j||^2 / (2 a) per packed component (padding rows included) and, behind the vector, max_{j < K} ||b_j||^2
__global__ void kde_screen_pack_kernel(const float* __restrict__ rows, int64_t K, int64_t KP, int DP, float inv2a,
                                       float* __restrict__ hs) {
  const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  float b2 = 0.f;
  if (j < KP) {
    for (int d = 0; d < DP; ++d) {
      const float v = rows[j * DP + d];
      b2 = fmaf(v, v, b2);
    }
    hs[j] = b2 * inv2a;
  }
  float mx = (j < K) ? b2 : 0.f;
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(hs + KP), __float_as_int(mx));  // non-negative floats
}

__global__ void lse_merge_kernel(const float* __restrict__ ws_m, const float* __restrict__ ws_S, int nsplit, int64_t n,
                                 float add, float* __restrict__ out) {
  const int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (p >= n) return;
  float M = kNegBig;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, ws_m[s * n + p]);
  float S = 0.f;
  for (int s = 0; s < nsplit; ++s) S += ws_S[s * n + p] * ex2(ws_m[s * n + p] - M);
  out[p] = lse_finish(M, S) + add;
}

// ---------------------------------------------------------------------------------------------
// Fused DM weights: both mixtures evaluated against the register-resident points in one launch.
// Target is packed DIAG with offsets (or absent: logp read from memory); proposals ISO/DIAG.
// ---------------------------------------------------------------------------------------------
template <int DP, int KIND_Q, bool HAS_C_Q>
__global__ void __launch_bounds__(kThreads)
    dm_weights_kernel(const float* __restrict__ x, int64_t n, int D, int vec_ok, PassArgs tgt,
                      const float* __restrict__ logp_in, PassArgs prop, float* __restrict__ out_logw,
                      float* __restrict__ out_logp, float* __restrict__ out_logq, double* __restrict__ block_partials) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[kStages];
  __shared__ float red[kThreads / 32 * 4];
  __shared__ OuterAcc outer;
  constexpr int E = points_per_thread(DP);

  ring_init(bars);
  float xr[E][DP];
  int64_t pidx[E];
  load_points<DP, E>(x, n, D, vec_ok, static_cast<int64_t>(blockIdx.x) * (kThreads * E), xr, pidx);

  float m[E], S[E], logp[E], logq[E];
  uint32_t it = 0;
  if (tgt.rows != nullptr) {
    mixture_pass<AISB_COV_DIAG, DP, E, true>(tgt.rows, tgt.offs, 0, tgt.KP, 0.f, smem, bars, outer, it, xr, m, S);
#pragma unroll
    for (int e = 0; e < E; ++e) logp[e] = lse_finish(m[e], S[e]);
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) logp[e] = pidx[e] < n ? __ldg(logp_in + pidx[e]) : 0.f;
  }
  mixture_pass<KIND_Q, DP, E, HAS_C_Q>(prop.rows, prop.offs, 0, prop.KP, prop.a_iso, smem, bars, outer, it, xr, m, S);

  float lw[E];
  bool ok[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    logq[e] = lse_finish(m[e], S[e]) + (HAS_C_Q ? 0.f : prop.uniform_offset);
    lw[e] = logp[e] - logq[e];
    const bool in = pidx[e] < n;
    ok[e] = in && (lw[e] == lw[e]) && (lw[e] != INFINITY);
    if (in) {
      out_logw[pidx[e]] = lw[e];
      if (out_logp) out_logp[pidx[e]] = logp[e];
      if (out_logq) out_logq[pidx[e]] = logq[e];
    }
  }
  block_weight_partials<E, kThreads>(lw, ok, red, block_partials + 4 * static_cast<int64_t>(blockIdx.x));
}

// logw = logp - logq on vectors + block partials (unfused fall-back path of aisb_dm_weights_f32)
__global__ void __launch_bounds__(256) logw_combine_kernel(const float* __restrict__ logp, const float* __restrict__ logq,
                                                           int64_t n, float* __restrict__ out_logw,
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
    lw[e] = in ? logp[p] - logq[p] : 0.f;
    ok[e] = in && (lw[e] == lw[e]) && (lw[e] != INFINITY);
    if (in) out_logw[p] = lw[e];
  }
  block_weight_partials<E, 256>(lw, ok, red, block_partials + 4 * static_cast<int64_t>(blockIdx.x));
}

// ---------------------------------------------------------------------------------------------
// KDE pack: rows[j][d] = -a * samples[idx[j]][d], pad dims 0, pad components 1e18
// ---------------------------------------------------------------------------------------------
__global__ void kde_pack_kernel(const float* __restrict__ samples, int64_t n, int D, int DP,
                                const int64_t* __restrict__ idx, int64_t n_c, int64_t KP, float a,
                                float* __restrict__ rows) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= KP * DP) return;
  const int64_t j = i / DP;
  const int d = static_cast<int>(i % DP);
  float v;
  if (j >= n_c) {
    v = 1.0e18f;
  } else if (d >= D) {
    v = 0.f;
  } else {
    int64_t src = idx ? idx[j] : j;
    src = src < 0 ? 0 : (src >= n ? n - 1 : src);
    v = -a * samples[src * D + d];
  }
  rows[i] = v;
}

// ---------------------------------------------------------------------------------------------
// Host-side planning and dispatch
// ---------------------------------------------------------------------------------------------
struct Plan {
  int DP, E;
  int64_t KP, ntiles, comps_per_split;
  int nsplit;
};

static bool make_plan(int64_t n, int64_t K, int D, Plan* p) {
  p->DP = aisb_padded_dims(D);
  if (p->DP == 0 || n < 0 || K < 1) return false;
  p->E = points_per_thread(p->DP);
  p->KP = aisb_padded_components(K);
  const int64_t tile_pts = static_cast<int64_t>(kThreads) * p->E;
  p->ntiles = n > 0 ? (n + tile_pts - 1) / tile_pts : 0;
  // Split the component range when the point tiles alone cannot fill ~16 waves of CTAs
  // (4 resident CTAs per SM; tail loss <= 1/waves); keep >= 2048 components per split so that the
  // per-CTA prologue (point load, barrier init) and the partial (m, S) traffic stay negligible.
  const int64_t target = static_cast<int64_t>(sm_count()) * 4 * 16;
  int64_t max_split = p->KP / 2048;
  if (max_split < 1) max_split = 1;
  int64_t want = p->ntiles > 0 ? (target + p->ntiles - 1) / p->ntiles : 1;
  if (want > max_split) want = max_split;
  if (want < 1) want = 1;
  if (const char* e = getenv("AISB_NSPLIT")) {  // timing experiments
    const int64_t v = atoll(e);
    if (v >= 1 && v <= max_split) want = v;
  }
  const int64_t units = p->KP / AISB_COMP_ALIGN;
  const int64_t units_per_split = (units + want - 1) / want;
  p->comps_per_split = units_per_split * AISB_COMP_ALIGN;
  p->nsplit = static_cast<int>((p->KP + p->comps_per_split - 1) / p->comps_per_split);
  return true;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  float* ws_m;
  float* ws_S;
  float* tmp_p;
  float* tmp_q;
  double* block_partials;
  size_t bytes;
};

static void carve(void* base, int64_t n, int nsplit, int64_t nblocks, Workspace* w) {
  size_t off = 0;
  char* b = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    void* ptr = b ? b + off : nullptr;
    off += align_up(bytes, 256);
    return ptr;
  };
  w->ws_m = static_cast<float*>(take(sizeof(float) * n * nsplit));
  w->ws_S = static_cast<float*>(take(sizeof(float) * n * nsplit));
  w->tmp_p = static_cast<float*>(take(sizeof(float) * n));
  w->tmp_q = static_cast<float*>(take(sizeof(float) * n));
  w->block_partials = static_cast<double*>(take(sizeof(double) * 8 * (nblocks + 1)));
  w->bytes = off;
}

static int64_t reduce_blocks(int64_t n) { return (n + 2047) / 2048; }

using LogpdfFn = void (*)(const float*, int64_t, int, int, PassArgs, int, int64_t, int64_t, float*, float*, float*);
using DmFn = void (*)(const float*, int64_t, int, int, PassArgs, const float*, PassArgs, float*, float*, float*,
                      double*);

template <int KIND, int DP, bool HAS_C, bool SKIP = false>
static int launch_logpdf_t(const float* x, int64_t n, int D, int vec_ok, const PassArgs& mix, const Plan& plan,
                           float* out, const Workspace& ws, cudaStream_t st) {
  constexpr size_t smem = pass_smem_bytes(KIND, DP, HAS_C);
  const int64_t grid = plan.ntiles * plan.nsplit;
  logpdf_kernel<KIND, DP, HAS_C, SKIP><<<static_cast<unsigned>(grid), kThreads, smem, st>>>(
      x, n, D, vec_ok, mix, plan.nsplit, plan.comps_per_split, plan.ntiles, out, ws.ws_m, ws.ws_S);
  AISB_LAUNCH_CHECK("logpdf_kernel");
  if (plan.nsplit > 1) {
    lse_merge_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(ws.ws_m, ws.ws_S, plan.nsplit, n,
                                                                             HAS_C ? 0.f : mix.uniform_offset, out);
    AISB_LAUNCH_CHECK("lse_merge_kernel");
  }
  return AISB_OK;
}


template <int DP>
static int launch_screen_t(const float* x, int64_t n, const PassArgs& mix, const float* hs, int nsplit, float* out,
                           const Workspace& ws, unsigned long long* stats, cudaStream_t st) {
  constexpr size_t smem = screen_smem_bytes(DP);
  constexpr int E = points_per_thread(DP);
  const int64_t ntiles = (n + kScreenThreads * E - 1) / (kScreenThreads * E);
  const int64_t grid = ntiles * nsplit;
  const float marg = 21.f + log2f(static_cast<float>(mix.KP));
  if (stats)
    logpdf_screen_kernel<DP, true><<<static_cast<unsigned>(grid), kScreenThreads, smem, st>>>(
        x, n, mix, hs, marg, nsplit, ntiles, out, ws.ws_m, ws.ws_S, stats);
  else
    logpdf_screen_kernel<DP, false><<<static_cast<unsigned>(grid), kScreenThreads, smem, st>>>(
        x, n, mix, hs, marg, nsplit, ntiles, out, ws.ws_m, ws.ws_S, nullptr);
  AISB_LAUNCH_CHECK("logpdf_screen_kernel");
  if (nsplit > 1) {
    lse_merge_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(ws.ws_m, ws.ws_S, nsplit, n,
                                                                             mix.uniform_offset, out);
    AISB_LAUNCH_CHECK("lse_merge_kernel");
  }
  return AISB_OK;
}


template <int DP>
static int launch_fast_t(const float* x, int64_t n, const PassArgs& mix, int nsplit, float* out, const Workspace& ws,
                         float* ws_u2, int64_t* list, unsigned long long* list_n, float* sub_m, float* sub_S,
                         uint32_t* gmin, unsigned long long* stats, cudaStream_t st) {
  constexpr size_t smem = size_t(kStages) * kScreenChunk * DP * sizeof(float);
  constexpr int E = points_per_thread(DP);
  const int64_t ntiles = (n + kScreenThreads * E - 1) / (kScreenThreads * E);
  const int64_t grid = ntiles * nsplit;
  const float marg = 21.f + log2f(static_cast<float>(mix.KP));
  // bit 0: fold every group; bit 1: no recentring; bit 2: no exact pass; bit 3: no exchange of minima between splits
  const char* dbg_env = getenv("AISB_FAST_DEBUG");
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  AISB_CUDA_CHECK(cudaMemsetAsync(list_n, 0, sizeof(unsigned long long), st));
  if (dbg & 8) gmin = nullptr;   // bit 3: no exchange of the running minima between the splits
  if (nsplit > 1 && gmin) AISB_CUDA_CHECK(cudaMemsetAsync(gmin, 0xFF, sizeof(uint32_t) * n, st));
  if (stats)
    logpdf_fast_kernel<DP, true><<<static_cast<unsigned>(grid), kScreenThreads, smem, st>>>(
        x, n, mix, marg, nsplit, ntiles, out, ws.ws_m, ws.ws_S, ws_u2, list, list_n, stats, dbg, gmin);
  else
    logpdf_fast_kernel<DP, false><<<static_cast<unsigned>(grid), kScreenThreads, smem, st>>>(
        x, n, mix, marg, nsplit, ntiles, out, ws.ws_m, ws.ws_S, ws_u2, list, list_n, nullptr, dbg, gmin);
  AISB_LAUNCH_CHECK("logpdf_fast_kernel");
  if (nsplit > 1) {
    fast_merge_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(ws.ws_m, ws.ws_S, ws_u2, nsplit, n,
                                                                              mix.uniform_offset, out, list, list_n);
    AISB_LAUNCH_CHECK("fast_merge_kernel");
  }
  // exact re-evaluation of the listed points: one fixed-size launch that plans itself from the list length on the device
  // (CTAs without work exit at once) and its merge
  if (!(dbg & 4)) {
    constexpr size_t smem_x = pass_smem_bytes(AISB_COV_ISO_SHARED, DP, false);
    static_assert(kThreads * E <= kThreads * 4, "kSubsetWs assumes at most 4 points per thread");
    logpdf_subset_kernel<DP><<<kSubsetGrid, kThreads, smem_x, st>>>(x, n, mix, list, list_n, sub_m, sub_S, out);
    AISB_LAUNCH_CHECK("logpdf_subset_kernel");
    subset_merge_kernel<<<256, 256, 0, st>>>(sub_m, sub_S, n, mix.KP, kThreads * E, kSubsetGrid, list, list_n,
                                            mix.uniform_offset, out);
    AISB_LAUNCH_CHECK("subset_merge_kernel");
  }
  if (stats) {
    AISB_CUDA_CHECK(cudaMemcpyAsync(stats + 2, list_n, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
  }
  return AISB_OK;
}

template <int KIND, bool HAS_C>
static int launch_logpdf_dp(int DP, const float* x, int64_t n, int D, int vec_ok, const PassArgs& mix,
                            const Plan& plan, float* out, const Workspace& ws, cudaStream_t st) {
  switch (DP) {
    case 1: return launch_logpdf_t<KIND, 1, HAS_C>(x, n, D, vec_ok, mix, plan, out, ws, st);
    case 2: return launch_logpdf_t<KIND, 2, HAS_C>(x, n, D, vec_ok, mix, plan, out, ws, st);
    case 4: return launch_logpdf_t<KIND, 4, HAS_C>(x, n, D, vec_ok, mix, plan, out, ws, st);
    case 8: return launch_logpdf_t<KIND, 8, HAS_C>(x, n, D, vec_ok, mix, plan, out, ws, st);
    case 16: return launch_logpdf_t<KIND, 16, HAS_C>(x, n, D, vec_ok, mix, plan, out, ws, st);
    case 32: return launch_logpdf_t<KIND, 32, HAS_C>(x, n, D, vec_ok, mix, plan, out, ws, st);
    case 64:
      if constexpr (KIND != AISB_COV_FULL) return launch_logpdf_t<KIND, 64, HAS_C>(x, n, D, vec_ok, mix, plan, out, ws, st);
      break;
    default: break;
  }
  set_error("mixture_logpdf: unsupported (cov_kind=%d, D=%d)", KIND, D);
  return AISB_ERR_UNSUPPORTED;
}

static int check_mixture(const aisb_packed_mixture* mix, const char* who) {
  if (!mix || !mix->d_rows) {
    set_error("%s: null mixture", who);
    return AISB_ERR_INVALID;
  }
  if (mix->D < 1 || mix->D > AISB_MAX_DIMS || mix->K < 1) {
    set_error("%s: bad mixture shape K=%lld D=%d", who, static_cast<long long>(mix->K), mix->D);
    return AISB_ERR_INVALID;
  }
  if (mix->cov_kind != AISB_COV_ISO_SHARED && !mix->d_offsets) {
    set_error("%s: offsets are required for cov_kind %d", who, mix->cov_kind);
    return AISB_ERR_INVALID;
  }
  if (mix->cov_kind == AISB_COV_FULL && mix->D > 32) {
    set_error("%s: full covariance is limited to D <= 32 (got %d)", who, mix->D);
    return AISB_ERR_UNSUPPORTED;
  }
  if (reinterpret_cast<uintptr_t>(mix->d_rows) % 16 || reinterpret_cast<uintptr_t>(mix->d_offsets) % 16) {
    set_error("%s: packed buffers must be 16-byte aligned", who);
    return AISB_ERR_INVALID;
  }
  return AISB_OK;
}

static PassArgs pass_args(const aisb_packed_mixture* mix) {
  PassArgs a;
  a.rows = mix->d_rows;
  a.offs = mix->d_offsets;
  a.KP = aisb_padded_components(mix->K);
  a.a_iso = mix->iso_scale;
  a.uniform_offset = mix->uniform_offset;
  a.skip_margin = 28.f + log2f(static_cast<float>(a.KP));
  return a;
}

static int vec_ok_for(const float* x, int D) {
  const int DP = aisb_padded_dims(D);
  if (D != DP) return 0;
  if (DP >= 4) return reinterpret_cast<uintptr_t>(x) % 16 == 0;
  if (DP == 2) return reinterpret_cast<uintptr_t>(x) % 8 == 0;
  return 0;
}

static int run_logpdf(const float* x, int64_t n, const aisb_packed_mixture* mix, float* out, const Plan& plan,
                      const Workspace& ws, cudaStream_t st, bool ordered = false) {
  const PassArgs a = pass_args(mix);
  const int vec_ok = vec_ok_for(x, mix->D);
  if (ordered && mix->cov_kind == AISB_COV_ISO_SHARED && !mix->d_offsets) {
    // spatially ordered points against a spatially ordered KDE: the fold-skipping variants
    switch (plan.DP) {
      case 1: return launch_logpdf_t<AISB_COV_ISO_SHARED, 1, false, true>(x, n, mix->D, vec_ok, a, plan, out, ws, st);
      case 2: return launch_logpdf_t<AISB_COV_ISO_SHARED, 2, false, true>(x, n, mix->D, vec_ok, a, plan, out, ws, st);
      case 4: return launch_logpdf_t<AISB_COV_ISO_SHARED, 4, false, true>(x, n, mix->D, vec_ok, a, plan, out, ws, st);
      case 8: return launch_logpdf_t<AISB_COV_ISO_SHARED, 8, false, true>(x, n, mix->D, vec_ok, a, plan, out, ws, st);
      case 16: return launch_logpdf_t<AISB_COV_ISO_SHARED, 16, false, true>(x, n, mix->D, vec_ok, a, plan, out, ws, st);
      default: break;  // larger D: Z-order says little, the ordinary kernel below
    }
  }
  switch (mix->cov_kind) {
    case AISB_COV_ISO_SHARED:
      if (mix->d_offsets) return launch_logpdf_dp<AISB_COV_ISO_SHARED, true>(plan.DP, x, n, mix->D, vec_ok, a, plan, out, ws, st);
      return launch_logpdf_dp<AISB_COV_ISO_SHARED, false>(plan.DP, x, n, mix->D, vec_ok, a, plan, out, ws, st);
    case AISB_COV_DIAG:
      return launch_logpdf_dp<AISB_COV_DIAG, true>(plan.DP, x, n, mix->D, vec_ok, a, plan, out, ws, st);
    case AISB_COV_FULL:
      return launch_logpdf_dp<AISB_COV_FULL, true>(plan.DP, x, n, mix->D, vec_ok, a, plan, out, ws, st);
    default: break;
  }
  set_error("mixture_logpdf: unknown cov_kind %d", mix->cov_kind);
  return AISB_ERR_INVALID;
}

template <int DP, int KIND_Q, bool HAS_C_Q>
static int launch_dm_t(const float* x, int64_t n, int D, int vec_ok, const PassArgs& tgt, const float* logp_in,
                       const PassArgs& prop, float* lw, float* lp, float* lq, double* bp, int64_t ntiles,
                       cudaStream_t st) {
  constexpr size_t s1 = pass_smem_bytes(AISB_COV_DIAG, DP, true);
  constexpr size_t s2 = pass_smem_bytes(KIND_Q, DP, HAS_C_Q);
  constexpr size_t smem = s1 > s2 ? s1 : s2;
  dm_weights_kernel<DP, KIND_Q, HAS_C_Q>
      <<<static_cast<unsigned>(ntiles), kThreads, smem, st>>>(x, n, D, vec_ok, tgt, logp_in, prop, lw, lp, lq, bp);
  AISB_LAUNCH_CHECK("dm_weights_kernel");
  return AISB_OK;
}

template <int KIND_Q, bool HAS_C_Q>
static int launch_dm_dp(int DP, const float* x, int64_t n, int D, int vec_ok, const PassArgs& tgt,
                        const float* logp_in, const PassArgs& prop, float* lw, float* lp, float* lq, double* bp,
                        int64_t ntiles, cudaStream_t st) {
  switch (DP) {
    case 1: return launch_dm_t<1, KIND_Q, HAS_C_Q>(x, n, D, vec_ok, tgt, logp_in, prop, lw, lp, lq, bp, ntiles, st);
    case 2: return launch_dm_t<2, KIND_Q, HAS_C_Q>(x, n, D, vec_ok, tgt, logp_in, prop, lw, lp, lq, bp, ntiles, st);
    case 4: return launch_dm_t<4, KIND_Q, HAS_C_Q>(x, n, D, vec_ok, tgt, logp_in, prop, lw, lp, lq, bp, ntiles, st);
    case 8: return launch_dm_t<8, KIND_Q, HAS_C_Q>(x, n, D, vec_ok, tgt, logp_in, prop, lw, lp, lq, bp, ntiles, st);
    case 16: return launch_dm_t<16, KIND_Q, HAS_C_Q>(x, n, D, vec_ok, tgt, logp_in, prop, lw, lp, lq, bp, ntiles, st);
    case 32: return launch_dm_t<32, KIND_Q, HAS_C_Q>(x, n, D, vec_ok, tgt, logp_in, prop, lw, lp, lq, bp, ntiles, st);
    case 64: return launch_dm_t<64, KIND_Q, HAS_C_Q>(x, n, D, vec_ok, tgt, logp_in, prop, lw, lp, lq, bp, ntiles, st);
    default: break;
  }
  set_error("dm_weights: unsupported D=%d", D);
  return AISB_ERR_UNSUPPORTED;
}

}  // namespace aisb

using namespace aisb;

extern "C" int aisb_workspace_bytes(int64_t n, int64_t K, int32_t D, size_t* bytes) {
  if (!bytes) {
    set_error("workspace_bytes: null output");
    return AISB_ERR_INVALID;
  }
  Plan plan;
  if (!make_plan(n, K, D, &plan)) {
    set_error("workspace_bytes: bad shape n=%lld K=%lld D=%d", static_cast<long long>(n), static_cast<long long>(K), D);
    return AISB_ERR_INVALID;
  }
  Workspace w;
  const int64_t nblocks = plan.ntiles > reduce_blocks(n) ? plan.ntiles : reduce_blocks(n);
  carve(nullptr, n, plan.nsplit, nblocks, &w);
  *bytes = w.bytes + 256;
  return AISB_OK;
}

static int carve_checked(void* d_ws, size_t ws_bytes, int64_t n, const Plan& plan, Workspace* w, const char* who) {
  const int64_t nblocks = plan.ntiles > reduce_blocks(n) ? plan.ntiles : reduce_blocks(n);
  char* base = static_cast<char*>(d_ws);
  const size_t mis = base ? (256 - reinterpret_cast<uintptr_t>(base) % 256) % 256 : 0;
  carve(base ? base + mis : nullptr, n, plan.nsplit, nblocks, w);
  if (!base || w->bytes + mis > ws_bytes) {
    set_error("%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, w->bytes + mis);
    return AISB_ERR_WORKSPACE;
  }
  return AISB_OK;
}

static int mixture_logpdf_impl(const float* d_x, int64_t n, const aisb_packed_mixture* mix, float* d_out_logp,
                               void* d_workspace, size_t workspace_bytes, void* stream, bool ordered) {
  int rc = check_mixture(mix, "mixture_logpdf");
  if (rc) return rc;
  if (n == 0) return AISB_OK;
  if (!d_x || !d_out_logp || n < 0) {
    set_error("mixture_logpdf: null pointer or negative n");
    return AISB_ERR_INVALID;
  }
  Plan plan;
  if (!make_plan(n, mix->K, mix->D, &plan)) {
    set_error("mixture_logpdf: unsupported D=%d", mix->D);
    return AISB_ERR_UNSUPPORTED;
  }
  Workspace ws{};
  if (plan.nsplit > 1) {
    rc = carve_checked(d_workspace, workspace_bytes, n, plan, &ws, "mixture_logpdf");
    if (rc) return rc;
  }
  return run_logpdf(d_x, n, mix, d_out_logp, plan, ws, static_cast<cudaStream_t>(stream), ordered);
}

extern "C" int aisb_mixture_logpdf_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix, float* d_out_logp,
                                       void* d_workspace, size_t workspace_bytes, void* stream) {
  return mixture_logpdf_impl(d_x, n, mix, d_out_logp, d_workspace, workspace_bytes, stream, false);
}

extern "C" int aisb_mixture_logpdf_ordered_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix,
                                               float* d_out_logp, void* d_workspace, size_t workspace_bytes,
                                               void* stream) {
  return mixture_logpdf_impl(d_x, n, mix, d_out_logp, d_workspace, workspace_bytes, stream, true);
}


extern "C" int64_t aisb_kde_screen_floats(int64_t K) { return K < 1 ? 0 : aisb_padded_components(K) + 8; }

extern "C" int aisb_kde_screen_pack_f32(const aisb_packed_mixture* mix, float* d_screen, void* stream) {
  int rc = check_mixture(mix, "kde_screen_pack");
  if (rc) return rc;
  if (mix->cov_kind != AISB_COV_ISO_SHARED || mix->d_offsets || !d_screen ||
      reinterpret_cast<uintptr_t>(d_screen) % 16) {
    set_error("kde_screen_pack: needs an ISO_SHARED mixture without offsets and a 16-byte aligned output");
    return AISB_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int DP = aisb_padded_dims(mix->D);
  const int64_t KP = aisb_padded_components(mix->K);
  AISB_CUDA_CHECK(cudaMemsetAsync(d_screen + KP, 0, 8 * sizeof(float), st));
  kde_screen_pack_kernel<<<static_cast<unsigned>((KP + 255) / 256), 256, 0, st>>>(mix->d_rows, mix->K, KP, DP,
                                                                                0.5f / mix->iso_scale, d_screen);
  AISB_LAUNCH_CHECK("kde_screen_pack_kernel");
  return AISB_OK;
}

extern "C" int aisb_mixture_logpdf_screened_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix,
                                                const float* d_screen, float* d_out_logp, void* d_workspace,
                                                size_t workspace_bytes, unsigned long long* d_stats, void* stream) {
  int rc = check_mixture(mix, "mixture_logpdf_screened");
  if (rc) return rc;
  if (n == 0) return AISB_OK;
  if (!d_x || !d_out_logp || n < 0 || !d_screen || reinterpret_cast<uintptr_t>(d_screen) % 16) {
    set_error("mixture_logpdf_screened: null / misaligned pointer or negative n");
    return AISB_ERR_INVALID;
  }
  const int DP = aisb_padded_dims(mix->D);
  if (mix->cov_kind != AISB_COV_ISO_SHARED || mix->d_offsets || mix->D != DP || DP < 2 || DP > 16 ||
      !vec_ok_for(d_x, mix->D)) {
    set_error("mixture_logpdf_screened: needs an ISO_SHARED mixture without offsets, D in {2, 4, 8, 16} and 16-byte "
              "aligned points (got cov_kind=%d D=%d)", mix->cov_kind, mix->D);
    return AISB_ERR_UNSUPPORTED;
  }
  Plan plan;
  if (!make_plan(n, mix->K, mix->D, &plan)) {
    set_error("mixture_logpdf_screened: unsupported D=%d", mix->D);
    return AISB_ERR_UNSUPPORTED;
  }
  // Single-warp CTAs, up to 16 resident per SM: split the component range (interleaved chunks) until the grid has ~8
  // waves of them, never beyond what the shared workspace contract (aisb_workspace_bytes) provides, and so that every
  // CTA still sweeps >= 16 chunks.
  constexpr int E8 = 4;
  const int64_t ntiles = (n + kScreenThreads * E8 - 1) / (kScreenThreads * E8);
  const int64_t nchunks = (plan.KP + kScreenChunk - 1) / kScreenChunk;
  int64_t want = (static_cast<int64_t>(sm_count()) * 16 * 8 + ntiles - 1) / ntiles;
  if (want > nchunks / 16) want = nchunks / 16;
  if (want > plan.nsplit) want = plan.nsplit;
  if (want < 1) want = 1;
  if (const char* e = getenv("AISB_SCREEN_NSPLIT")) {  // timing experiments
    const int64_t v = atoll(e);
    if (v >= 1 && v <= plan.nsplit && v <= nchunks) want = v;
  }
  const int nsplit = static_cast<int>(want);
  Workspace ws{};
  if (nsplit > 1) {
    rc = carve_checked(d_workspace, workspace_bytes, n, plan, &ws, "mixture_logpdf_screened");
    if (rc) return rc;
  }
  const PassArgs a = pass_args(mix);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (DP) {
    case 2: return launch_screen_t<2>(d_x, n, a, d_screen, nsplit, d_out_logp, ws, d_stats, st);
    case 4: return launch_screen_t<4>(d_x, n, a, d_screen, nsplit, d_out_logp, ws, d_stats, st);
    case 8: return launch_screen_t<8>(d_x, n, a, d_screen, nsplit, d_out_logp, ws, d_stats, st);
    case 16: return launch_screen_t<16>(d_x, n, a, d_screen, nsplit, d_out_logp, ws, d_stats, st);
    default: break;
  }
  return AISB_ERR_UNSUPPORTED;
}


extern "C" size_t aisb_kde_fast_scratch_bytes(int64_t n) {
  if (n < 0) return 0;
  return static_cast<size_t>(n) * (sizeof(int64_t) + sizeof(float)) + 2 * sizeof(float) * kSubsetWs +
         2 * sizeof(float) * static_cast<size_t>(fast_split_cap(n)) * static_cast<size_t>(n) +
         sizeof(uint32_t) * static_cast<size_t>(n) + 1024;
}

extern "C" int aisb_mixture_logpdf_fast_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix,
                                            float* d_out_logp, void* d_workspace, size_t workspace_bytes,
                                            void* d_scratch, size_t scratch_bytes, unsigned long long* d_stats,
                                            void* stream) {
  int rc = check_mixture(mix, "mixture_logpdf_fast");
  if (rc) return rc;
  if (n == 0) return AISB_OK;
  if (!d_x || !d_out_logp || n < 0 || !d_scratch || reinterpret_cast<uintptr_t>(d_scratch) % 16 ||
      scratch_bytes < aisb_kde_fast_scratch_bytes(n)) {
    set_error("mixture_logpdf_fast: null / misaligned pointer, negative n or scratch smaller than "
              "aisb_kde_fast_scratch_bytes(n)");
    return AISB_ERR_INVALID;
  }
  const int DP = aisb_padded_dims(mix->D);
  if (mix->cov_kind != AISB_COV_ISO_SHARED || mix->d_offsets || mix->D != DP || DP < 2 || DP > 16 ||
      !vec_ok_for(d_x, mix->D)) {
    set_error("mixture_logpdf_fast: needs an ISO_SHARED mixture without offsets, D in {2, 4, 8, 16} and 16-byte "
              "aligned points (got cov_kind=%d D=%d)", mix->cov_kind, mix->D);
    return AISB_ERR_UNSUPPORTED;
  }
  // single-warp CTAs, up to 16 resident per SM: interleaved split of the component range (see kFastWaves), >= 16 chunks
  // per CTA; the partial (m, S) of the splits live in the scratch buffer
  const int64_t KPc = aisb_padded_components(mix->K);
  const int64_t nchunks = (KPc + kScreenChunk - 1) / kScreenChunk;
  const int64_t cap = fast_split_cap(n);
  int64_t want = cap;
  if (want > nchunks / 16) want = nchunks / 16;
  if (want < 1) want = 1;
  if (const char* e = getenv("AISB_SCREEN_NSPLIT")) {  // timing experiments
    const int64_t v = atoll(e);
    if (v >= 1 && v <= cap && v <= nchunks) want = v;
  }
  const int nsplit = static_cast<int>(want);
  (void)d_workspace;
  (void)workspace_bytes;
  char* sc = static_cast<char*>(d_scratch);
  unsigned long long* list_n = reinterpret_cast<unsigned long long*>(sc);
  int64_t* list = reinterpret_cast<int64_t*>(sc + 64);
  float* ws_u2 = reinterpret_cast<float*>(sc + 64 + sizeof(int64_t) * n);
  float* sub_m = reinterpret_cast<float*>(sc + 64 + (sizeof(int64_t) + sizeof(float)) * n + 64);
  sub_m = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sub_m) + 15) / 16 * 16);
  float* sub_S = sub_m + kSubsetWs;
  Workspace ws{};
  ws.ws_m = sub_S + kSubsetWs;
  ws.ws_S = ws.ws_m + static_cast<size_t>(cap) * n;
  uint32_t* gmin = reinterpret_cast<uint32_t*>(ws.ws_S + static_cast<size_t>(cap) * n);
  const PassArgs a = pass_args(mix);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (DP) {
    case 2: return launch_fast_t<2>(d_x, n, a, nsplit, d_out_logp, ws, ws_u2, list, list_n, sub_m, sub_S, gmin, d_stats, st);
    case 4: return launch_fast_t<4>(d_x, n, a, nsplit, d_out_logp, ws, ws_u2, list, list_n, sub_m, sub_S, gmin, d_stats, st);
    case 8: return launch_fast_t<8>(d_x, n, a, nsplit, d_out_logp, ws, ws_u2, list, list_n, sub_m, sub_S, gmin, d_stats, st);
    case 16: return launch_fast_t<16>(d_x, n, a, nsplit, d_out_logp, ws, ws_u2, list, list_n, sub_m, sub_S, gmin, d_stats, st);
    default: break;
  }
  return AISB_ERR_UNSUPPORTED;
}

extern "C" int aisb_dm_weights_f32(const float* d_x, int64_t n, const aisb_packed_mixture* target,
                                   const float* d_logp_in, const aisb_packed_mixture* proposals, float* d_out_logw,
                                   float* d_out_logp, float* d_out_logq, double* d_out_partials, void* d_workspace,
                                   size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_mixture(proposals, "dm_weights(proposals)");
  if (rc) return rc;
  if (target) {
    rc = check_mixture(target, "dm_weights(target)");
    if (rc) return rc;
    if (target->D != proposals->D) {
      set_error("dm_weights: target D=%d != proposal D=%d", target->D, proposals->D);
      return AISB_ERR_INVALID;
    }
  } else if (!d_logp_in) {
    set_error("dm_weights: need a target mixture or d_logp_in");
    return AISB_ERR_INVALID;
  }
  if (!d_out_partials || n < 0 || (n > 0 && (!d_x || !d_out_logw))) {
    set_error("dm_weights: null pointer or negative n");
    return AISB_ERR_INVALID;
  }
  if (n == 0) {
    const double empty[4] = {static_cast<double>(kNegBig), 0.0, 0.0, 0.0};
    AISB_CUDA_CHECK(cudaMemcpyAsync(d_out_partials, empty, sizeof(empty), cudaMemcpyHostToDevice, st));
    AISB_CUDA_CHECK(cudaStreamSynchronize(st));  // `empty` is a stack buffer
    return AISB_OK;
  }
  Plan plan_q, plan_t;
  if (!make_plan(n, proposals->K, proposals->D, &plan_q)) {
    set_error("dm_weights: unsupported D=%d", proposals->D);
    return AISB_ERR_UNSUPPORTED;
  }
  plan_t = plan_q;
  if (target) make_plan(n, target->K, target->D, &plan_t);
  Plan plan_ws = plan_q.nsplit >= plan_t.nsplit ? plan_q : plan_t;
  Workspace ws{};
  rc = carve_checked(d_workspace, workspace_bytes, n, plan_ws, &ws, "dm_weights");
  if (rc) return rc;

  const int D = proposals->D;
  const int vec_ok = vec_ok_for(d_x, D);
  const bool fusable = plan_q.nsplit == 1 && plan_t.nsplit == 1 && proposals->cov_kind != AISB_COV_FULL &&
                       (!target || target->cov_kind == AISB_COV_DIAG);
  int64_t nblocks;
  if (fusable) {
    PassArgs t{};
    if (target) t = pass_args(target);
    const PassArgs q = pass_args(proposals);
    nblocks = plan_q.ntiles;
    if (proposals->cov_kind == AISB_COV_DIAG)
      rc = launch_dm_dp<AISB_COV_DIAG, true>(plan_q.DP, d_x, n, D, vec_ok, t, d_logp_in, q, d_out_logw, d_out_logp,
                                             d_out_logq, ws.block_partials, nblocks, st);
    else if (proposals->d_offsets)
      rc = launch_dm_dp<AISB_COV_ISO_SHARED, true>(plan_q.DP, d_x, n, D, vec_ok, t, d_logp_in, q, d_out_logw,
                                                   d_out_logp, d_out_logq, ws.block_partials, nblocks, st);
    else
      rc = launch_dm_dp<AISB_COV_ISO_SHARED, false>(plan_q.DP, d_x, n, D, vec_ok, t, d_logp_in, q, d_out_logw,
                                                    d_out_logp, d_out_logq, ws.block_partials, nblocks, st);
    if (rc) return rc;
  } else {
    const float* lp = d_logp_in;
    if (target) {
      float* lp_out = d_out_logp ? d_out_logp : ws.tmp_p;
      rc = run_logpdf(d_x, n, target, lp_out, plan_t, ws, st);
      if (rc) return rc;
      lp = lp_out;
    } else if (d_out_logp) {
      AISB_CUDA_CHECK(cudaMemcpyAsync(d_out_logp, d_logp_in, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    }
    float* lq = d_out_logq ? d_out_logq : ws.tmp_q;
    rc = run_logpdf(d_x, n, proposals, lq, plan_q, ws, st);
    if (rc) return rc;
    nblocks = reduce_blocks(n);
    logw_combine_kernel<<<static_cast<unsigned>(nblocks), 256, 0, st>>>(lp, lq, n, d_out_logw, ws.block_partials);
    AISB_LAUNCH_CHECK("logw_combine_kernel");
  }
  weight_partials_merge_kernel<<<1, 256, 0, st>>>(ws.block_partials, nblocks, d_out_partials, nullptr);
  AISB_LAUNCH_CHECK("weight_partials_merge_kernel");
  return AISB_OK;
}

// DM weights with the proposal density on the tensor cores (hinted screen + exact refinement, tc.cu): target density
// on the CUDA cores, log q from tc_logq_launch, then the same combine / partial-sum kernels as the unfused path.
extern "C" int aisb_dm_weights_hinted_f32(const float* d_x, int64_t n, const aisb_packed_mixture* target,
                                          const float* d_logp_in, const aisb_packed_mixture* proposals,
                                          const float* d_tc_image, const int32_t* d_hint, float* d_out_logw,
                                          float* d_out_logp, float* d_out_logq, double* d_out_partials,
                                          void* d_workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_mixture(proposals, "dm_weights_hinted(proposals)");
  if (rc) return rc;
  if (target) {
    rc = check_mixture(target, "dm_weights_hinted(target)");
    if (rc) return rc;
    if (target->D != proposals->D) {
      set_error("dm_weights_hinted: target D=%d != proposal D=%d", target->D, proposals->D);
      return AISB_ERR_INVALID;
    }
  } else if (!d_logp_in) {
    set_error("dm_weights_hinted: need a target mixture or d_logp_in");
    return AISB_ERR_INVALID;
  }
  if (!d_out_partials || !d_tc_image || !d_hint || n < 1 || !d_x || !d_out_logw) {
    set_error("dm_weights_hinted: null pointer or n < 1");
    return AISB_ERR_INVALID;
  }
  Plan plan_q, plan_t;
  if (!make_plan(n, proposals->K, proposals->D, &plan_q)) {
    set_error("dm_weights_hinted: unsupported D=%d", proposals->D);
    return AISB_ERR_UNSUPPORTED;
  }
  plan_t = plan_q;
  if (target) make_plan(n, target->K, target->D, &plan_t);
  Workspace ws{};
  rc = carve_checked(d_workspace, workspace_bytes, n, plan_q.nsplit >= plan_t.nsplit ? plan_q : plan_t, &ws,
                     "dm_weights_hinted");
  if (rc) return rc;
  const float* lp = d_logp_in;
  if (target) {
    float* lp_out = d_out_logp ? d_out_logp : ws.tmp_p;
    rc = run_logpdf(d_x, n, target, lp_out, plan_t, ws, st);
    if (rc) return rc;
    lp = lp_out;
  } else if (d_out_logp) {
    AISB_CUDA_CHECK(cudaMemcpyAsync(d_out_logp, d_logp_in, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  }
  float* lq = d_out_logq ? d_out_logq : ws.tmp_q;
  const int64_t nblocks = reduce_blocks(n);
  // the abort counter lives behind the block records (carve() reserves nblocks + 1 records)
  int32_t* d_err = reinterpret_cast<int32_t*>(ws.block_partials + 8 * nblocks);
  AISB_CUDA_CHECK(cudaMemsetAsync(d_err, 0, sizeof(int32_t), st));
  rc = tc_logq_launch(d_x, n, proposals, d_hint, d_tc_image, lq, d_err, st);
  if (rc) return rc;
  logw_combine_kernel<<<static_cast<unsigned>(nblocks), 256, 0, st>>>(lp, lq, n, d_out_logw, ws.block_partials);
  AISB_LAUNCH_CHECK("logw_combine_kernel");
  // a tensor-core kernel that aborted (barrier time-out) poisons the record: n_valid = -1
  weight_partials_merge_kernel<<<1, 256, 0, st>>>(ws.block_partials, nblocks, d_out_partials, d_err);
  AISB_LAUNCH_CHECK("weight_partials_merge_kernel");
  return AISB_OK;
}

extern "C" int aisb_kde_pack_f32(const float* d_samples, int64_t n, int32_t D, const int64_t* d_idx, int64_t n_c,
                                 double bw, float* d_rows, void* stream) {
  const int DP = aisb_padded_dims(D);
  if (!d_samples || !d_rows || n < 1 || n_c < 1 || DP == 0 || !(bw > 0.0) || (!d_idx && n_c > n)) {
    set_error("kde_pack: invalid argument (n=%lld n_c=%lld D=%d bw=%g)", static_cast<long long>(n),
              static_cast<long long>(n_c), D, bw);
    return AISB_ERR_INVALID;
  }
  const int64_t KP = aisb_padded_components(n_c);
  const float a = static_cast<float>(sqrt(kLog2e / (2.0 * bw)));
  const int64_t total = KP * DP;
  kde_pack_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_samples, n, D, DP, d_idx, n_c, KP, a, d_rows);
  AISB_LAUNCH_CHECK("kde_pack_kernel");
  return AISB_OK;
}
